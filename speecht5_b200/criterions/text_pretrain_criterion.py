"""Text pre-training criterion (task_name 'text_pretrain'): BART-style denoising cross-entropy of the text decoder over
the target tokens + the weighted codebook-diversity term of the shared quantizer. Mirrors
speecht5/criterions/text_pretrain_criterion.py:36-105 (loss, sample size, logging keys); the scalar reads are one
device->host copy. `reduce_metrics` lives in the dispatcher (speecht5_criterion.py, text_* / bart_* keys)."""
import torch
import torch.nn.functional as F

from ..fairseq_shim import FairseqCriterion
from .speech_pretrain_criterion import weighted_extra_losses


class TextPretrainCriterion(FairseqCriterion):
    def __init__(self, task, sentence_avg=True, bart_weight=1.0, loss_weights=(0.1,)):
        super().__init__(task)
        self.sentence_avg, self.bart_weight = sentence_avg, bart_weight
        self.loss_weights = list(loss_weights) if loss_weights is not None else None
        if not hasattr(self, "padding_idx"):  # (real fairseq: FairseqCriterion.__init__ reads it from the target dictionary)
            d = getattr(task, "target_dictionary", None) if task is not None else None
            self.padding_idx = d.pad() if d is not None else 1

    def forward(self, model, sample, reduce=True):
        net_output, codebook_out, _encoder_output = model(**sample["net_input"])
        bart_loss, _ = self.compute_loss(model, net_output, sample, reduce=reduce)
        sample_size = sample["target"].size(0) if self.sentence_avg else sample["ntokens"]
        loss = self.bart_weight * bart_loss
        scalars = {"bart_loss": bart_loss.detach()}
        if "prob_perplexity" in codebook_out:  # (:64-85)
            assert hasattr(model, "get_extra_losses")
            extras, self.loss_weights = weighted_extra_losses(model, codebook_out, self.loss_weights, sample_size, tail=True)
            for name, term in extras:
                loss = loss + term
                scalars[f"loss_{name}"] = term.detach()
        if "loss_prob_perplexity" in scalars:
            scalars["code_perplexity"] = codebook_out["code_perplexity"].detach()
        scalars["loss"] = (self.bart_weight * bart_loss).detach()  # (:57: logged before the extra terms are added)
        host = torch.stack([v.double().sum() for v in scalars.values()]).tolist()
        logging_output = {"ntokens": sample["ntokens"], "nsentences": sample["target"].size(0), "sample_size": sample_size}
        logging_output.update(zip(scalars.keys(), host))
        return loss, sample_size, logging_output

    def compute_loss(self, model, net_output, sample, reduce=True):
        lprobs = model.get_normalized_probs(net_output, log_probs=True)
        lprobs = lprobs.view(-1, lprobs.size(-1))
        target = model.get_targets(sample, net_output).view(-1)
        loss = F.nll_loss(lprobs, target, ignore_index=self.padding_idx, reduction="sum" if reduce else "none")
        return loss, loss

    @staticmethod
    def logging_outputs_can_be_summed() -> bool:
        return True
