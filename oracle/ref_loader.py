"""TEST INFRASTRUCTURE (never imported by the product path): loads the reference's OWN hot-path modules, unmodified, from
`/root/reference/SpeechT5` so the oracle restatements can be pinned against them and golden vectors can be generated
from the reference itself (`tests/golden/make_golden_from_ref.py`).

The reference package cannot be imported as a whole in this image (`import fairseq` needs omegaconf<2.1 / hydra-core<1.1
/ numpy<1.24; espnet, librosa absent), but the files on the hot path only need a handful of names from it.
This loader therefore

* registers empty *namespace* packages `fairseq`, `fairseq.modules`, `fairseq.models`, `fairseq.data`, … in
  `sys.modules` (so none of the packages' heavy `__init__`s run),
* executes the reference's real source files for every leaf the path touches BY PATH under their real module names
  (`fairseq/utils.py`, `fairseq/modules/{layer_norm,fairseq_dropout,quant_noise,gelu,sinusoidal_positional_embedding,
  learned_positional_embedding,positional_embedding,same_pad,transpose_last,fp32_group_norm,grad_multiply,layer_drop,
  gumbel_vector_quantizer}.py`, `fairseq/incremental_decoding_utils.py`, `fairseq/data/data_utils.py`,
  `fairseq/models/{fairseq_encoder,fairseq_decoder,fairseq_incremental_decoder,fairseq_model}.py`, `fairseq/search.py`,
  `fairseq/token_generation_constraints.py`, `fairseq/ngram_repeat_block.py`),
* pulls single definitions out of files that are too entangled to execute (`Embedding`, `Linear` of
  `fairseq/models/transformer.py`, `init_bert_params` of `fairseq/modules/transformer_sentence_encoder.py`,
  `FairseqCriterion` …) by compiling just those AST nodes of the reference source — still the reference's code, nothing
  is copied into this repository,
* stubs what is infrastructure, not arithmetic: the model / criterion registries, `metrics`, omegaconf's `II`, FSDP and
  activation-checkpoint wrappers, the file-system `PathManager`, `Dictionary`,
* provides espnet's five classes (`Prenet`, `Postnet`, `PositionalEncoding`, `ScaledPositionalEncoding`,
  `make_non_pad_mask`, `GuidedAttentionLoss`) from the restatement in `oracle/espnet_standin.py`: espnet is an
  UN-VENDORED dependency of the reference (version unpinned, `SpeechT5/README.md:32`), its source is not under
  /root/reference, so for those classes the pin remains the HuggingFace port (`oracle/hf_crosscheck.py`),
* and finally imports `speecht5/models/modules/*.py`, `speecht5/models/speecht5.py`, `speecht5/criterions/*.py` and
  `speecht5/sequence_generator.py` themselves through the normal import machinery.

`available()` is False when /root/reference is absent (the GPU box): tests that need the live reference skip there and
the committed fixtures `tests/golden/ref_*.npz` take over.
"""
import ast
import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("ST5_REFERENCE_ROOT", "/root/reference")
ST5 = os.path.join(REF_ROOT, "SpeechT5")
FAIRSEQ = os.path.join(ST5, "fairseq", "fairseq")
HIFIGAN = os.path.join(REF_ROOT, "SpeechUT", "fairseq", "fairseq", "models", "text_to_speech", "hifigan.py")

_loaded = {}


def available():
    return os.path.isfile(os.path.join(ST5, "speecht5", "models", "speecht5.py")) and os.path.isdir(FAIRSEQ)


def _namespace(name, path=None):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        m.__package__ = name
        sys.modules[name] = m
        if "." in name:
            parent, leaf = name.rsplit(".", 1)
            setattr(_namespace(parent), leaf, m)
    return m


def _exec_file(name, path):
    """Run a reference source file under its real dotted module name."""
    if name in sys.modules and getattr(sys.modules[name], "__file__", None) == path:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    parent, leaf = name.rsplit(".", 1)
    spec.loader.exec_module(mod)
    setattr(_namespace(parent), leaf, mod)
    return mod


def _extract(path, names, namespace, module_name):
    """Compile only the named top-level definitions (and plain imports) of a reference file into `namespace`."""
    tree = ast.parse(open(path).read(), filename=path)
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in keep}
    assert not missing, f"{path}: {missing} not found"
    code = compile(ast.Module(body=keep, type_ignores=[]), path, "exec")
    namespace.setdefault("__name__", module_name)
    exec(code, namespace)
    return namespace


def _fairseq():
    if "fairseq" in _loaded:
        return
    import torch
    import torch.nn as nn
    for pkg in ("fairseq", "fairseq.modules", "fairseq.models", "fairseq.data", "fairseq.distributed",
                "fairseq.criterions", "fairseq.dataclass", "fairseq.logging", "fairseq.tasks"):
        _namespace(pkg)
    fm = sys.modules["fairseq.modules"]
    fq = sys.modules["fairseq"]

    # --- infrastructure stubs (no arithmetic) ---------------------------------------------------------------------
    om = types.ModuleType("omegaconf")
    om.II = lambda key: None
    om.DictConfig = type("DictConfig", (dict,), {})
    om.open_dict = None
    sys.modules.setdefault("omegaconf", om)
    fio = types.ModuleType("fairseq.file_io")
    fio.PathManager = type("PathManager", (), {})
    sys.modules["fairseq.file_io"] = fio
    fq.file_io = fio
    mha_stub = types.ModuleType("fairseq.modules.multihead_attention")
    mha_stub.MultiheadAttention = type("MultiheadAttention", (nn.Module,), {})  # only an isinstance/annotation target
    sys.modules["fairseq.modules.multihead_attention"] = mha_stub
    fm.MultiheadAttention = mha_stub.MultiheadAttention
    meters = types.ModuleType("fairseq.logging.meters")

    def safe_round(number, ndigits):  # fairseq/logging/meters.py:36-46 semantics for python / torch / numpy scalars
        if hasattr(number, "__round__"):
            return round(number, ndigits)
        if torch.is_tensor(number) and number.numel() == 1:
            return safe_round(number.item(), ndigits)
        return number
    meters.safe_round = safe_round
    sys.modules["fairseq.logging.meters"] = meters
    sys.modules["fairseq.logging"].meters = meters

    class _Metrics(types.ModuleType):
        """Records what reduce_metrics logs so tests can read it back."""
        def __init__(self):
            super().__init__("fairseq.metrics")
            self.logged, self.derived = {}, {}

        def log_scalar(self, key, value, weight=1, priority=10, round=None):
            self.logged[key] = value

        def log_derived(self, key, fn, priority=20):
            self.derived[key] = fn

        def log_scalar_sum(self, key, value, priority=10, round=None):
            self.logged[key] = value
    fq.metrics = _Metrics()
    sys.modules["fairseq.metrics"] = fq.metrics
    sys.modules["fairseq.logging"].metrics = fq.metrics

    # --- real leaves, executed by path ----------------------------------------------------------------------------
    for leaf in ("layer_norm", "fairseq_dropout", "quant_noise", "gelu", "same_pad", "transpose_last",
                 "fp32_group_norm", "grad_multiply", "layer_drop", "gumbel_vector_quantizer", "adaptive_softmax"):
        _exec_file(f"fairseq.modules.{leaf}", os.path.join(FAIRSEQ, "modules", f"{leaf}.py"))
    fm.LayerNorm = fm.layer_norm.LayerNorm
    fm.Fp32LayerNorm = fm.layer_norm.Fp32LayerNorm
    fm.FairseqDropout = fm.fairseq_dropout.FairseqDropout
    fm.gelu, fm.gelu_accurate = fm.gelu.gelu, fm.gelu.gelu_accurate  # fairseq/modules/__init__.py re-exports the functions
    fm.SamePad = fm.same_pad.SamePad
    fm.TransposeLast = fm.transpose_last.TransposeLast
    fm.Fp32GroupNorm = fm.fp32_group_norm.Fp32GroupNorm
    fm.GradMultiply = fm.grad_multiply.GradMultiply
    fm.LayerDropModuleList = fm.layer_drop.LayerDropModuleList
    fm.GumbelVectorQuantizer = fm.gumbel_vector_quantizer.GumbelVectorQuantizer
    fm.AdaptiveSoftmax = fm.adaptive_softmax.AdaptiveSoftmax
    fm.TransformerEncoderLayer = type("TransformerEncoderLayer", (nn.Module,), {})  # imported by encoder.py, never used
    fm.TransformerSentenceEncoderLayer = type("TransformerSentenceEncoderLayer", (nn.Module,), {})
    _exec_file("fairseq.incremental_decoding_utils", os.path.join(FAIRSEQ, "incremental_decoding_utils.py"))
    _exec_file("fairseq.utils", os.path.join(FAIRSEQ, "utils.py"))
    for leaf in ("sinusoidal_positional_embedding", "learned_positional_embedding", "positional_embedding"):
        _exec_file(f"fairseq.modules.{leaf}", os.path.join(FAIRSEQ, "modules", f"{leaf}.py"))
    fm.PositionalEmbedding = fm.positional_embedding.PositionalEmbedding
    fm.SinusoidalPositionalEmbedding = fm.sinusoidal_positional_embedding.SinusoidalPositionalEmbedding
    fm.LearnedPositionalEmbedding = fm.learned_positional_embedding.LearnedPositionalEmbedding
    _exec_file("fairseq.data.data_utils", os.path.join(FAIRSEQ, "data", "data_utils.py"))

    # init_bert_params: fairseq/modules/transformer_sentence_encoder.py (the file also defines an encoder we do not need)
    tse = types.ModuleType("fairseq.modules.transformer_sentence_encoder")
    tse.__dict__.update(nn=nn, torch=torch, MultiheadAttention=fm.MultiheadAttention)
    _extract(os.path.join(FAIRSEQ, "modules", "transformer_sentence_encoder.py"), ["init_bert_params"], tse.__dict__,
             tse.__name__)
    sys.modules[tse.__name__] = tse
    fm.transformer_sentence_encoder = tse

    # checkpoint / FSDP wrappers: identity (never enabled on the path under test)
    ca = types.ModuleType("fairseq.modules.checkpoint_activations")
    ca.checkpoint_wrapper = lambda m, *a, **k: m
    sys.modules[ca.__name__] = ca
    fm.checkpoint_activations = ca
    sys.modules["fairseq.distributed"].fsdp_wrap = lambda m, *a, **k: m

    # fairseq.models: real base classes + Embedding/Linear/LayerNorm helpers of models/transformer.py
    fmod = sys.modules["fairseq.models"]
    enc = _exec_file("fairseq.models.fairseq_encoder", os.path.join(FAIRSEQ, "models", "fairseq_encoder.py"))
    fmod.FairseqEncoder = enc.FairseqEncoder
    dec = _exec_file("fairseq.models.fairseq_decoder", os.path.join(FAIRSEQ, "models", "fairseq_decoder.py"))
    fmod.FairseqDecoder = dec.FairseqDecoder
    inc = _exec_file("fairseq.models.fairseq_incremental_decoder",
                     os.path.join(FAIRSEQ, "models", "fairseq_incremental_decoder.py"))
    fmod.FairseqIncrementalDecoder = inc.FairseqIncrementalDecoder
    sys.modules["fairseq.data"].Dictionary = type("Dictionary", (), {})
    dcu = types.ModuleType("fairseq.dataclass.utils")
    dcu.convert_namespace_to_omegaconf = lambda a: a
    dcu.gen_parser_from_dataclass = lambda *a, **k: None
    sys.modules[dcu.__name__] = dcu
    sys.modules["fairseq.dataclass"].utils = dcu
    sys.modules["fairseq.dataclass"].FairseqDataclass = type("FairseqDataclass", (), {})
    fmm = _exec_file("fairseq.models.fairseq_model", os.path.join(FAIRSEQ, "models", "fairseq_model.py"))
    fmod.BaseFairseqModel = fmm.BaseFairseqModel
    fmod.FairseqEncoderDecoderModel = fmm.FairseqEncoderDecoderModel
    fmod.FairseqLanguageModel = fmm.FairseqLanguageModel
    registry = {"models": {}, "archs": {}}

    def register_model(name, dataclass=None):
        def deco(cls):
            registry["models"][name] = cls
            return cls
        return deco

    def register_model_architecture(model_name, arch_name):
        def deco(fn):
            registry["archs"][arch_name] = (model_name, fn)
            return fn
        return deco
    fmod.register_model, fmod.register_model_architecture, fmod.REGISTRY = register_model, register_model_architecture, registry
    tr = types.ModuleType("fairseq.models.transformer")
    tr.__dict__.update(nn=nn, torch=torch)
    _extract(os.path.join(FAIRSEQ, "models", "transformer.py"), ["Embedding", "Linear"], tr.__dict__, tr.__name__)
    tr.LayerNorm = fm.LayerNorm
    sys.modules[tr.__name__] = tr
    fmod.transformer = tr

    # criterions: real FairseqCriterion base (fairseq/criterions/fairseq_criterion.py), registry stubbed
    fc = sys.modules["fairseq.criterions"]
    crit_ns = {"inspect": __import__("inspect"), "Any": object, "Dict": dict, "List": list, "_Loss": nn.modules.loss._Loss,
               "metrics": fq.metrics, "utils": fq.utils, "gen_parser_from_dataclass": dcu.gen_parser_from_dataclass,
               "FairseqDataclass": sys.modules["fairseq.dataclass"].FairseqDataclass, "torch": torch}
    _extract(os.path.join(FAIRSEQ, "criterions", "fairseq_criterion.py"), ["FairseqCriterion"], crit_ns,
             "fairseq.criterions.fairseq_criterion")
    fc.FairseqCriterion = crit_ns["FairseqCriterion"]
    fc.register_criterion = lambda name, dataclass=None: (lambda cls: cls)
    ls = types.ModuleType("fairseq.criterions.label_smoothed_cross_entropy")
    ls.LabelSmoothedCrossEntropyCriterionConfig = type("LabelSmoothedCrossEntropyCriterionConfig", (), {})
    sys.modules[ls.__name__] = ls
    sys.modules["fairseq.tasks"].FairseqTask = type("FairseqTask", (), {})

    # search / n-gram blocking for the sequence generator
    _exec_file("fairseq.token_generation_constraints", os.path.join(FAIRSEQ, "token_generation_constraints.py"))
    _exec_file("fairseq.search", os.path.join(FAIRSEQ, "search.py"))
    _exec_file("fairseq.ngram_repeat_block", os.path.join(FAIRSEQ, "ngram_repeat_block.py"))
    _loaded["fairseq"] = True


def _espnet():
    """espnet is not vendored by the reference: supply its classes from the restatement (see the module docstring)."""
    if "espnet" in _loaded:
        return
    from oracle import espnet_standin as so
    for pkg in ("espnet", "espnet.nets", "espnet.nets.pytorch_backend", "espnet.nets.pytorch_backend.tacotron2",
                "espnet.nets.pytorch_backend.transformer"):
        _namespace(pkg)
    d = types.ModuleType("espnet.nets.pytorch_backend.tacotron2.decoder")
    d.Prenet, d.Postnet = so.Prenet, so.Postnet
    e = types.ModuleType("espnet.nets.pytorch_backend.transformer.embedding")
    e.ScaledPositionalEncoding, e.PositionalEncoding = so.ScaledPositionalEncoding, so.PositionalEncoding
    n = types.ModuleType("espnet.nets.pytorch_backend.nets_utils")
    n.make_non_pad_mask = so.make_non_pad_mask
    t = types.ModuleType("espnet.nets.pytorch_backend.e2e_tts_tacotron2")
    t.GuidedAttentionLoss = so.GuidedAttentionLoss
    c = types.ModuleType("espnet.nets.ctc_prefix_score")
    c.CTCPrefixScore = type("CTCPrefixScore", (), {})  # joint CTC/attention beam search is out of scope
    for m in (d, e, n, t, c):
        sys.modules[m.__name__] = m
        parent, leaf = m.__name__.rsplit(".", 1)
        setattr(sys.modules[parent], leaf, m)
    _loaded["espnet"] = True


def load():
    """Returns a namespace with the reference's own classes: modules, model, criterions, sequence generator."""
    if "ns" in _loaded:
        return _loaded["ns"]
    assert available(), f"reference tree not found under {REF_ROOT}"
    _fairseq()
    _espnet()
    _namespace("speecht5", os.path.join(ST5, "speecht5"))
    _namespace("speecht5.models", os.path.join(ST5, "speecht5", "models"))
    _namespace("speecht5.models.modules", os.path.join(ST5, "speecht5", "models", "modules"))
    _namespace("speecht5.criterions", os.path.join(ST5, "speecht5", "criterions"))
    ns = types.SimpleNamespace()
    for leaf in ("multihead_attention", "transformer_layer", "encoder", "decoder", "text_encoder_prenet",
                 "text_decoder_prenet", "text_decoder_postnet", "speech_encoder_prenet", "speech_encoder_postnet",
                 "speech_decoder_prenet", "speech_decoder_postnet"):
        setattr(ns, leaf, importlib.import_module(f"speecht5.models.modules.{leaf}"))
    ns.model = importlib.import_module("speecht5.models.speecht5")
    ns.T5TransformerModel = ns.model.T5TransformerModel
    ns.tts_loss = importlib.import_module("speecht5.criterions.text_to_speech_loss")
    ns.asr_loss = importlib.import_module("speecht5.criterions.speech_to_text_loss")
    ns.sequence_generator = importlib.import_module("speecht5.sequence_generator")
    ns.fairseq_utils = sys.modules["fairseq.utils"]
    ns.metrics = sys.modules["fairseq.metrics"]
    ns.compute_mask_indices = sys.modules["fairseq.data.data_utils"].compute_mask_indices
    ns.GumbelVectorQuantizer = sys.modules["fairseq.modules"].GumbelVectorQuantizer
    _loaded["ns"] = ns
    return ns


def load_hifigan():
    """The reference's HiFi-GAN generator (sibling tree SpeechUT/fairseq, plain torch, no fairseq imports)."""
    assert os.path.isfile(HIFIGAN), HIFIGAN
    spec = importlib.util.spec_from_file_location("_ref_hifigan", HIFIGAN)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class RefDictionary:
    """Minimal stand-in for fairseq.data.Dictionary as the model builder uses it (len, pad/eos/unk/bos, index)."""

    def __init__(self, n, extra=("<mask>", "<ctc_blank>")):
        self.symbols = ["<s>", "<pad>", "</s>", "<unk>"] + [f"t{i}" for i in range(n - 4 - len(extra))] + list(extra)
        self.indices = {s: i for i, s in enumerate(self.symbols)}

    def __len__(self):
        return len(self.symbols)

    def bos(self):
        return 0

    def pad(self):
        return 1

    def eos(self):
        return 2

    def unk(self):
        return 3

    def index(self, sym):
        return self.indices.get(sym, 3)

    def __getitem__(self, i):
        return self.symbols[i]

    def string(self, t, *a, **k):
        return " ".join(self.symbols[int(i)] for i in t)


class RefTask:
    """What `T5TransformerModel.build_model(args, task)` and the criterions read from the task
    (tasks/speecht5.py:272-297)."""

    def __init__(self, vocab=81, t5_task="t2s", hubert_classes=None):
        self.dicts = {"text": RefDictionary(vocab)}
        if hubert_classes:
            self.dicts["hubert"] = [RefDictionary(hubert_classes, extra=())]
        self.t5_task = t5_task
        self.blank_symbol_idx = self.dicts["text"].index("<ctc_blank>")
        self.blank_symbol = "<ctc_blank>"

    @property
    def target_dictionary(self):
        return self.dicts["text"]

    @property
    def source_dictionary(self):
        return None


def build_reference_model(args, task=None):
    """`T5TransformerModel.build_model` of the reference on a Namespace (arch defaults filled by the reference's own
    `base_architecture`)."""
    ns = load()
    return ns.T5TransformerModel.build_model(args, task or RefTask())


def reference_args(arch="t5_transformer_base_asr", **overrides):
    """A Namespace the reference's builders accept: the options fairseq's parser would have filled that are NOT arch
    defaults (task options `tasks/speecht5.py:45-270`, a few model options with parser defaults), then the overrides, then
    the reference's OWN arch function (`models/speecht5.py:1252-1447`) for everything else."""
    from argparse import Namespace
    ns = load()
    a = dict(label_rates=50, sample_rate=16000, speech_odim=80, modules_filter=None, reduction_factor=2,
             spk_embed_dim=512, max_target_positions=1024, quant_noise_pq_block_size=8, tie_adaptive_weights=False,
             tie_adaptive_proj=False, adaptive_softmax_factor=4, transformer_dec_dropout_rate=0.1, relu_dropout=0.0,
             sid_pad_prenet=False, sid_t5_postnet=False, t5_task="t2s")
    a.update(overrides)
    args = Namespace(**a)
    ns.model.__dict__[{"t5_transformer": "base_architecture", "t5_transformer_base": "t5_transformer_base",
                       "t5_transformer_large": "t5_transformer_large",
                       "t5_transformer_base_asr": "t5_transformer_base_asr"}[arch]](args)
    return args
