"""TTS criterion: mirrors speecht5/criterions/text_to_speech_loss.py (TexttoSpeechLoss :72-214, Tacotron2Loss :217-345,
GuidedMultiHeadAttentionLoss :370-427). The masks are built on the device with broadcasting instead of the reference's
per-utterance Python loops (:404-413); arithmetic is unchanged."""
import torch
import torch.nn.functional as F

from ..fairseq_shim import FairseqCriterion


def make_non_pad_mask(lengths, maxlen):
    lengths = lengths.long()
    return torch.arange(maxlen, device=lengths.device)[None, :] < lengths[:, None]


class Tacotron2Loss(torch.nn.Module):
    def __init__(self, use_masking=True, use_weighted_masking=False, bce_pos_weight=20.0):
        super().__init__()
        assert use_masking and not use_weighted_masking, "reference recipe: use_masking=True"
        self.register_buffer("pos_weight", torch.tensor(float(bce_pos_weight)), persistent=False)

    def forward(self, after_outs, before_outs, logits, ys, labels, olens):
        masks = make_non_pad_mask(olens, ys.size(1)).unsqueeze(-1)
        n_el = masks.sum() * ys.size(2)
        m = masks.to(ys.dtype)
        da, db = (after_outs - ys) * m, (before_outs - ys) * m
        l1_loss = (da.abs().sum() + db.abs().sum()) / n_el
        mse_loss = ((da * da).sum() + (db * db).sum()) / n_el
        bce = F.binary_cross_entropy_with_logits(logits, labels, pos_weight=self.pos_weight.to(logits.device),
                                                 reduction="none")
        bce_loss = (bce * m[:, :, 0]).sum() / masks.sum()
        return l1_loss, mse_loss, bce_loss


class GuidedMultiHeadAttentionLoss(torch.nn.Module):
    def __init__(self, sigma=0.4, alpha=1.0):
        super().__init__()
        self.sigma, self.alpha = sigma, alpha

    def forward(self, att_ws, ilens, olens):
        """att_ws (B, H, T_out, T_in)."""
        T_out, T_in = att_ws.size(2), att_ws.size(3)
        dev = att_ws.device
        il, ol = ilens.to(dev).float(), olens.to(dev).float()
        gx = torch.arange(T_out, device=dev).float()[None, :, None] / ol[:, None, None]
        gy = torch.arange(T_in, device=dev).float()[None, None, :] / il[:, None, None]
        w = 1.0 - torch.exp(-((gy - gx) ** 2) / (2 * self.sigma ** 2))
        masks = make_non_pad_mask(olens.to(dev), T_out).unsqueeze(-1) & make_non_pad_mask(ilens.to(dev), T_in).unsqueeze(-2)
        w = (w * masks).unsqueeze(1)
        loss = (w * att_ws).sum() / (masks.sum() * att_ws.size(1))
        return self.alpha * loss


class TexttoSpeechLoss(FairseqCriterion):
    def __init__(self, task, sentence_avg=True, use_masking=True, use_weighted_masking=False, loss_type="L1",
                 bce_pos_weight=5.0, bce_loss_lambda=1.0, use_guided_attn_loss=False, guided_attn_loss_sigma=0.4,
                 guided_attn_loss_lambda=1.0, num_layers_applied_guided_attn=2, num_heads_applied_guided_attn=2,
                 modules_applied_guided_attn=("encoder-decoder",)):
        super().__init__(task)
        self.loss_type, self.bce_loss_lambda = loss_type, bce_loss_lambda
        self.use_guided_attn_loss = use_guided_attn_loss
        self.criterion = Tacotron2Loss(use_masking, use_weighted_masking, bce_pos_weight)
        self.num_heads_applied_guided_attn = num_heads_applied_guided_attn
        self.modules_applied_guided_attn = modules_applied_guided_attn
        if use_guided_attn_loss:
            self.attn_criterion = GuidedMultiHeadAttentionLoss(guided_attn_loss_sigma, guided_attn_loss_lambda)

    def forward(self, model, sample):
        net_output = model(**sample["net_input"])
        loss, l1_loss, l2_loss, bce_loss, enc_dec_attn_loss = self.compute_loss(model, net_output, sample)
        stats = torch.stack([loss.detach(), l1_loss.detach(), l2_loss.detach(), bce_loss.detach(),
                             enc_dec_attn_loss.detach() if enc_dec_attn_loss is not None else loss.new_zeros(()),
                             model.text_encoder_prenet.encoder_prenet[-1].alpha.detach().float(),
                             model.speech_decoder_prenet.decoder_prenet[-1].alpha.detach().float()])
        if getattr(self, "defer_logging", False):
            # no device->host sync inside the step (CUDA-graph capture): the caller reads `_stats` after the update
            return loss, 1, {"_stats": stats, "sample_size": 1, "ntokens": sample["ntokens"],
                             "nsentences": sample["target"].size(0)}
        stats = stats.tolist()  # the reference's ~7 .item() calls (:129-146) batched into ONE device->host copy
        logging_output = {"loss": stats[0], "l1_loss": stats[1], "l2_loss": stats[2], "bce_loss": stats[3],
                          "sample_size": 1, "ntokens": sample["ntokens"], "nsentences": sample["target"].size(0)}
        if enc_dec_attn_loss is not None:
            logging_output["enc_dec_attn_loss"] = stats[4]
        logging_output["encoder_alpha"], logging_output["decoder_alpha"] = stats[5], stats[6]
        return loss, 1, logging_output

    def compute_loss(self, model, net_output, sample):
        before_outs, after_outs, logits, attn = net_output
        labels, ys = sample["labels"], sample["dec_target"]
        olens, ilens = sample["dec_target_lengths"], sample["src_lengths"]
        r = model.reduction_factor
        if r > 1:
            olens_in = torch.div(olens, r, rounding_mode="floor")
            olens = olens - olens % r
            # The reference slices ys/labels to max(olens) (:161-166); frames beyond it are masked out anyway, so we
            # keep the padded length (= the model output length) and avoid a device->host sync in the step.
            L = after_outs.size(1)
            ys, labels = ys[:, :L], labels[:, :L]
            labels = torch.scatter(labels, 1, (olens - 1).unsqueeze(1), 1.0)
        else:
            olens_in = olens
        l1_loss, l2_loss, bce_loss = self.criterion(after_outs, before_outs, logits, ys, labels, olens)
        if self.loss_type == "L1":
            loss = l1_loss + self.bce_loss_lambda * bce_loss if self.bce_loss_lambda > 0.0 else l1_loss
        elif self.loss_type == "L2":
            loss = l2_loss + self.bce_loss_lambda * bce_loss if self.bce_loss_lambda > 0.0 else l2_loss
        elif self.loss_type == "L1+L2":
            loss = l1_loss + l2_loss + self.bce_loss_lambda * bce_loss if self.bce_loss_lambda > 0.0 else l1_loss + l2_loss
        else:
            raise ValueError("unknown --loss-type " + self.loss_type)
        enc_dec_attn_loss = None
        if self.use_guided_attn_loss and "encoder-decoder" in self.modules_applied_guided_attn:
            att_ws = torch.cat([a[:, : self.num_heads_applied_guided_attn] for a in attn], dim=1)
            enc_dec_attn_loss = self.attn_criterion(att_ws, ilens, olens_in)
            loss = loss + enc_dec_attn_loss
        return loss, l1_loss, l2_loss, bce_loss, enc_dec_attn_loss
