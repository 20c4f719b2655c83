"""@register_criterion("speecht5") dispatcher, mirroring speecht5/criterions/speecht5_criterion.py:23-120: routes on
sample['task_name']. Round 1 wires the t2s branch (TexttoSpeechLoss); note the reference dispatcher does not forward
guided_attn_loss_lambda, so the effective guided-attention weight is 1.0 (speecht5_criterion.py:61-71)."""
from dataclasses import dataclass, field

from ..fairseq_shim import FairseqCriterion, register_criterion
from .text_to_speech_loss import TexttoSpeechLoss


@dataclass
class SpeechT5CriterionConfig:
    sentence_avg: bool = field(default=True)
    use_masking: bool = field(default=True)
    loss_type: str = field(default="L1")
    bce_pos_weight: float = field(default=5.0)
    bce_loss_lambda: float = field(default=1.0)
    use_guided_attn_loss: bool = field(default=False)
    guided_attn_loss_sigma: float = field(default=0.4)
    num_heads_applied_guided_attn: int = field(default=2)


@register_criterion("speecht5", dataclass=SpeechT5CriterionConfig)
class SpeechT5Criterion(FairseqCriterion):
    def __init__(self, task, sentence_avg=True, use_masking=True, loss_type="L1", bce_pos_weight=5.0,
                 bce_loss_lambda=1.0, use_guided_attn_loss=False, guided_attn_loss_sigma=0.4,
                 num_heads_applied_guided_attn=2, **unused):
        super().__init__(task)
        self.text_to_speech_loss = TexttoSpeechLoss(
            task, sentence_avg, use_masking, False, loss_type, bce_pos_weight, bce_loss_lambda, use_guided_attn_loss,
            guided_attn_loss_sigma, 1.0, 2, num_heads_applied_guided_attn)

    def forward(self, model, sample, reduce=True):
        task_name = sample["task_name"]
        if task_name in ("t2s", "s2s"):
            return self.text_to_speech_loss(model, sample)
        raise NotImplementedError(f"criterion branch '{task_name}' is not built yet in the B200 path (round 1: t2s)")

    @staticmethod
    def logging_outputs_can_be_summed() -> bool:
        return False
