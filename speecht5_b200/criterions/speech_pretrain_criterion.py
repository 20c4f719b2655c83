"""Speech pre-training criterion (task_name 'speech_pretrain'): HuBERT masked / unmasked frame prediction against the
label embeddings + the weighted extra terms (feature penalty, codebook diversity) + the mel reconstruction branch of the
speech decoder. Mirrors speecht5/criterions/speech_pretrain_criterion.py:50-190 (forward) -- same loss, sample size and
logging keys; the per-term `.item()` reads of the reference (:101, :113, :137, :150, :158-189) are gathered into ONE
device->host copy at the end of the step. `reduce_metrics` lives in the dispatcher (speecht5_criterion.py, hubert_* keys)."""
import torch
import torch.nn.functional as F

from ..fairseq_shim import FairseqCriterion
from .text_to_speech_loss import TexttoSpeechLoss


def weighted_extra_losses(model, net_output, loss_weights, sample_size, tail=False):
    """speech_pretrain_criterion.py:118-138 / text_pretrain_criterion.py:66-85: [(name, coef * term * sample_size)].
    One configured weight is broadcast over all terms. With MORE weights than terms the speech criterion keeps the
    first len(terms) of them (:128-129) while the text criterion keeps what follows them (`tail`, :74-75 -- so with the
    recipe's single 0.1 and one term nothing changes, and with two weights and one term the text branch adds nothing)."""
    extra, names = model.get_extra_losses(net_output)
    if torch.is_tensor(extra):
        extra, names = [extra], [names]
    weights = list(loss_weights)
    if len(weights) == 1 and len(extra) != 1:
        weights = weights * len(extra)
    if len(weights) > len(extra):
        weights = weights[len(extra):] if tail else weights[:len(extra)]
    out = []
    for term, name, coef in zip(extra, names, weights):
        if coef != 0 and term is not None:
            out.append((name, coef * term.float() * sample_size))
    return out, weights


class SpeechPretrainCriterion(FairseqCriterion):
    def __init__(self, task, sentence_avg=True, pred_masked_weight=1.0, pred_nomask_weight=0.0, loss_weights=(10.0,),
                 log_keys=None, use_masking=True, use_weighted_masking=False, loss_type="L1", bce_pos_weight=5.0,
                 hubert_weight=1.0, dec_weight=1.0):
        super().__init__(task)
        self.pred_masked_weight, self.pred_nomask_weight = pred_masked_weight, pred_nomask_weight
        self.loss_weights = list(loss_weights) if loss_weights is not None else None
        self.log_keys = list(log_keys) if log_keys else []
        self.hubert_weight, self.dec_weight = hubert_weight, dec_weight
        # (:76-83: bce_loss_lambda and the guided-attention options keep TexttoSpeechLoss's defaults here)
        self.speech_criterion = TexttoSpeechLoss(task, sentence_avg, use_masking, use_weighted_masking, loss_type,
                                                 bce_pos_weight)

    @staticmethod
    def _first_is_extreme(logits):
        """(:153-163) a frame counts as correct when class 0 -- the true label's slot -- is the arg-max and not also
        the arg-min (all-equal rows). Returns (correct, count) as device scalars / python int."""
        if logits.numel() == 0:
            return logits.new_zeros((), dtype=torch.long), 0
        top = logits.argmax(-1) == 0
        low = logits.argmin(-1) == 0
        return (top & ~low).long().sum(), top.numel()

    def forward(self, model, sample, reduce=True, log_pred=False):
        net_input = dict(sample["net_input"])
        if self.dec_weight == 0:  # (:91-92) encoder-only pre-training: the decoder branch is not run at all
            net_input["only_hubert"] = True
        net_output, net_output_dec = model(target_list=sample["target_list"], **net_input)
        reduction = "sum" if reduce else "none"
        scalars = {}  # name -> device scalar, read back together
        loss, sample_size = 0.0, 0
        for tag, masked, weight in (("m", True, self.pred_masked_weight), ("u", False, self.pred_nomask_weight)):
            logits_list = model.get_logits(net_output, masked)
            target_list = model.get_targets(None, net_output, masked)
            assert weight == 0 or len(logits_list) > 0
            terms = []
            for i, (lg, tg) in enumerate(zip(logits_list, target_list)):
                ce = F.cross_entropy(lg, tg, reduction=reduction)
                terms.append(ce)
                if reduce:
                    scalars[f"loss_{tag}_{i}"] = ce.detach()
                with torch.no_grad():
                    corr, count = self._first_is_extreme(lg)
                scalars[f"correct_{tag}_{i}"] = corr
                scalars[f"count_{tag}_{i}"] = count
            if weight > 0:
                loss = loss + weight * sum(terms)
                sample_size += target_list[0].numel()
        if self.loss_weights is not None:
            assert hasattr(model, "get_extra_losses")
            extras, self.loss_weights = weighted_extra_losses(model, net_output, self.loss_weights, sample_size)
            for name, term in extras:
                loss = loss + term
                scalars[f"loss_{name}"] = term.detach()
        if "loss_prob_perplexity" in scalars:
            scalars["code_perplexity"] = net_output["code_perplexity"].detach()
        for lk in self.log_keys:
            if lk in net_output:
                v = net_output[lk]  # (the reference calls .item() on it, :154-156: tensors only; plain numbers pass here too)
                scalars[lk] = v.detach().float() if torch.is_tensor(v) else float(v)
        if self.dec_weight != 0.0:  # (:171-189) reconstruction branch, scaled to the frame count of the HuBERT term
            dec_loss, l1_loss, l2_loss, bce_loss, attn_loss = self.speech_criterion.compute_loss(model, net_output_dec, sample)
            scalars.update(dec_loss=dec_loss.detach(), l1_loss=l1_loss.detach(), l2_loss=l2_loss.detach(),
                           bce_loss=bce_loss.detach())
            if attn_loss is not None:
                scalars["enc_dec_attn_loss"] = attn_loss.detach()
            loss = self.hubert_weight * loss + self.dec_weight * sample_size * dec_loss
        if reduce and torch.is_tensor(loss):
            scalars["loss"] = loss.detach()
        logging_output = {"ntokens": sample_size, "nsentences": sample["id"].numel(), "sample_size": sample_size, "ngpu": 1}
        dev = {k: v for k, v in scalars.items() if torch.is_tensor(v)}
        host = torch.stack([v.double().reshape(()) for v in dev.values()]).tolist() if dev else []
        vals = dict(zip(dev.keys(), host))
        for k, v in scalars.items():
            v = vals.get(k, v)
            logging_output[k] = int(round(v)) if k.startswith(("correct_", "count_")) else v
        if not reduce:
            logging_output["loss"] = loss
        return loss, sample_size, logging_output

    @staticmethod
    def logging_outputs_can_be_summed() -> bool:
        return False
