"""Vocabulary files of the reference's task (`dict.txt`, `dict.<label>.txt`): the subset of fairseq's Dictionary the
SpeechT5 task, criteria and generator read (fairseq/data/dictionary.py:20-80 index rules, :219-262 file format). When
fairseq is importable its own class is used (`load_dictionary` returns it); this stand-in keeps the plugin usable without
it. Index layout: <s> 0, <pad> 1, </s> 2, <unk> 3, then the file's symbols in file order; a line is
`<symbol> <count>` with an optional trailing ` #fairseq:overwrite`."""
import torch


class Vocabulary:
    def __init__(self, bos="<s>", pad="<pad>", eos="</s>", unk="<unk>"):
        self.symbols, self.count, self.indices = [], [], {}
        self.bos_word, self.pad_word, self.eos_word, self.unk_word = bos, pad, eos, unk
        self.bos_index, self.pad_index = self.add_symbol(bos), self.add_symbol(pad)
        self.eos_index, self.unk_index = self.add_symbol(eos), self.add_symbol(unk)
        self.nspecial = len(self.symbols)

    # ---- construction
    def add_symbol(self, word, n=1, overwrite=False):
        """Index of `word`, appended when new (a known word only gains count unless `overwrite`)."""
        if word in self.indices and not overwrite:
            i = self.indices[word]
            self.count[i] += n
            return i
        i = len(self.symbols)
        self.indices[word] = i
        self.symbols.append(word)
        self.count.append(n)
        return i

    @classmethod
    def load(cls, path):
        d = cls()
        with open(path, "r", encoding="utf-8") as fh:
            for raw in fh:
                body = raw.rstrip()
                if not body:
                    continue
                parts = body.rsplit(" ", 1)
                if len(parts) != 2:
                    raise ValueError("Incorrect dictionary format, expected '<token> <cnt> [flags]'")
                word, field = parts
                overwrite = field == "#fairseq:overwrite"
                if overwrite:
                    word, field = word.rsplit(" ", 1)
                try:
                    n = int(field)
                except ValueError:
                    raise ValueError("Incorrect dictionary format, expected '<token> <cnt> [flags]'") from None
                if word in d.indices and not overwrite:
                    raise RuntimeError(f"Duplicate word found when loading Dictionary: '{word}'")
                d.add_symbol(word, n=n, overwrite=overwrite)
        return d

    # ---- what the task / criteria / generator read
    def __len__(self):
        return len(self.symbols)

    def __getitem__(self, i):
        return self.symbols[i] if i < len(self.symbols) else self.unk_word

    def __contains__(self, word):
        return word in self.indices

    def index(self, word):
        return self.indices.get(word, self.unk_index)

    def bos(self):
        return self.bos_index

    def pad(self):
        return self.pad_index

    def eos(self):
        return self.eos_index

    def unk(self):
        return self.unk_index

    def string(self, tensor, bpe_symbol=None, escape_unk=False, extra_symbols_to_ignore=None, unk_string=None,
               include_eos=False, separator=" "):
        """Token ids -> text as fairseq/data/dictionary.py:65-104 does it: bos, eos (always -- `include_eos` is accepted
        and, as there, has no effect on a 1-D input) and the ignored symbols are dropped; a 2-D input gives one line per
        row with the default separator / unknown string. bpe_symbol post-processing is the tokenizer's and is not built
        (None / 'none' only)."""
        if torch.is_tensor(tensor) and tensor.dim() == 2:
            return "\n".join(self.string(t, bpe_symbol, escape_unk, extra_symbols_to_ignore, include_eos=include_eos)
                             for t in tensor)
        if bpe_symbol not in (None, "none"):
            raise NotImplementedError("bpe_symbol post-processing is the tokenizer's (not built in the stand-in vocabulary)")
        skip = set(extra_symbols_to_ignore or [])
        skip.update((self.eos_index, self.bos_index))
        unk_text = unk_string if unk_string is not None else ("<{}>".format(self.unk_word) if escape_unk else self.unk_word)
        return separator.join(unk_text if i == self.unk_index else self[i] for i in (int(v) for v in tensor) if i not in skip)


def load_dictionary(path):
    """fairseq's Dictionary when the package is importable (the real plugin host), else the stand-in above."""
    try:
        from fairseq.data import Dictionary  # pragma: no cover  (fairseq is not installable in the build container)
        return Dictionary.load(path)
    except Exception:  # noqa: BLE001
        return Vocabulary.load(path)
