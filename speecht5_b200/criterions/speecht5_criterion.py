"""@register_criterion("speecht5") dispatcher, mirroring speecht5/criterions/speecht5_criterion.py:23-120: routes on
sample['task_name']. Round 1 wires the t2s branch (TexttoSpeechLoss); note the reference dispatcher does not forward
guided_attn_loss_lambda, so the effective guided-attention weight is 1.0 (speecht5_criterion.py:61-71)."""
from dataclasses import dataclass, field

from ..fairseq_shim import FairseqCriterion, register_criterion
from .speech_to_text_loss import SpeechtoTextLoss
from .text_to_speech_loss import TexttoSpeechLoss


@dataclass
class SpeechT5CriterionConfig:
    """Union of the reference's criterion configs (speecht5_criterion.py:23-30 inherits the text-to-speech,
    speech-to-text, label-smoothed CE and pre-training configs), so every recipe's `--criterion speecht5 ...` flags
    parse. Fields of branches that are not built yet are accepted and only matter once that branch is called."""
    sentence_avg: bool = field(default=True)
    # text_to_speech_loss.py:21-69
    use_masking: bool = field(default=True)
    use_weighted_masking: bool = field(default=False)
    loss_type: str = field(default="L1")
    bce_pos_weight: float = field(default=5.0)
    bce_loss_lambda: float = field(default=1.0)
    use_guided_attn_loss: bool = field(default=False)
    guided_attn_loss_sigma: float = field(default=0.4)
    guided_attn_loss_lambda: float = field(default=10.0)
    num_layers_applied_guided_attn: int = field(default=2)
    num_heads_applied_guided_attn: int = field(default=2)
    # speech_to_text_loss.py:27-91 / label-smoothed CE
    zero_infinity: bool = field(default=False)
    post_process: str = field(default="sentencepiece")
    label_smoothing: float = field(default=0.0)
    report_accuracy: bool = field(default=False)
    ignore_prefix_size: int = field(default=0)
    ce_weight: float = field(default=1.0)
    ctc_weight: float = field(default=0.0)
    # pre-training criteria (speech_pretrain_criterion.py, text_pretrain_criterion.py)
    pred_masked_weight: float = field(default=1.0)
    pred_nomask_weight: float = field(default=0.0)
    dec_weight: float = field(default=0.5)
    bart_weight: float = field(default=1.0)
    hubert_weight: float = field(default=1.0)


@register_criterion("speecht5", dataclass=SpeechT5CriterionConfig)
class SpeechT5Criterion(FairseqCriterion):
    def __init__(self, task, sentence_avg=True, use_masking=True, loss_type="L1", bce_pos_weight=5.0,
                 bce_loss_lambda=1.0, use_guided_attn_loss=False, guided_attn_loss_sigma=0.4,
                 num_heads_applied_guided_attn=2, label_smoothing=0.0, ignore_prefix_size=0, report_accuracy=False,
                 ce_weight=1.0, ctc_weight=0.0, zero_infinity=False, post_process="sentencepiece", **unused):
        super().__init__(task)
        self.text_to_speech_loss = TexttoSpeechLoss(
            task, sentence_avg, use_masking, False, loss_type, bce_pos_weight, bce_loss_lambda, use_guided_attn_loss,
            guided_attn_loss_sigma, 1.0, 2, num_heads_applied_guided_attn)
        # s2t branch (speecht5_criterion.py:72-81); meaningful only with the opt-in speech-input / text-output model
        self.speech_to_text_loss = SpeechtoTextLoss(task, sentence_avg, label_smoothing, ignore_prefix_size,
                                                    report_accuracy, ce_weight, ctc_weight, zero_infinity, post_process)

    def forward(self, model, sample, reduce=True):
        task_name = sample["task_name"]
        if task_name in ("t2s", "s2s"):
            return self.text_to_speech_loss(model, sample)
        if task_name == "s2t":
            return self.speech_to_text_loss(model, sample, reduce)
        raise NotImplementedError(f"criterion branch '{task_name}' is not built yet in the B200 path (round 1: t2s)")

    @staticmethod
    def logging_outputs_can_be_summed() -> bool:
        return False
