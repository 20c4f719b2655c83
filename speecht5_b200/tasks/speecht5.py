"""@register_task("speecht5"): the train_step / valid_step contract of speecht5/tasks/speecht5.py:519-571 (loss is
normalised locally by its own sample_size and the task reports sample_size 1, so data-parallel averaging is a plain
mean over ranks). Dataset loading (tasks/speecht5.py:324-517) is host I/O outside the hot path: synthetic batches with
the collater's key contract (data/text_to_speech_dataset.py:262-281) come from speecht5_b200.data."""
import torch

from ..fairseq_shim import LegacyFairseqTask, register_task
from ..models.speecht5 import T5TransformerModel, _Dict


@register_task("speecht5")
class SpeechT5Task(LegacyFairseqTask):
    def __init__(self, args, dicts=None, config=None):
        super().__init__(args)
        self.dicts = dicts if dicts is not None else {"text": _Dict(getattr(args, "vocab_size", 81))}
        self.config = config
        self.t5_task = getattr(args, "t5_task", "t2s")

    @classmethod
    def setup_task(cls, args, **kwargs):
        return cls(args)

    def build_model(self, args):
        args.speech_odim = 80  # tasks/speecht5.py:581-597
        return T5TransformerModel.build_model(args, self)

    def train_step(self, sample, model, criterion, optimizer, update_num, ignore_grad=False):
        model.train()
        model.set_num_updates(update_num)
        loss, sample_size, logging_output = criterion(model, sample)
        if ignore_grad:
            loss = loss * 0
        loss = loss / sample_size
        if optimizer is not None:
            optimizer.backward(loss)
        else:
            loss.backward()
        return loss.detach(), 1.0, logging_output

    def valid_step(self, sample, model, criterion):
        model.eval()
        with torch.no_grad():
            loss, sample_size, logging_output = criterion(model, sample)
        return loss, sample_size, logging_output
