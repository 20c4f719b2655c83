"""CPU restatement (oracle) of the pre-training extras of SpeechT5 -- SURVEY.md section 8a row 22: the HuBERT-style
masked-prediction head on the encoder output (speecht5/models/modules/speech_encoder_postnet.py:26-124) and the Gumbel
vector quantizer that mixes code vectors into the encoder states (fairseq/modules/gumbel_vector_quantizer.py:13-202,
used at models/speecht5.py:864-885). TEST INFRASTRUCTURE ONLY (see oracle/speecht5_oracle.py header).

The quantizer is pinned (evaluation mode: hard arg-max codes) against the independent
transformers.models.wav2vec2.Wav2Vec2GumbelVectorQuantizer; the prediction head against its closed form
(tests/test_oracle_cpu.py). In training mode the reference draws Gumbel noise with torch's RNG inside
F.gumbel_softmax; here the noise is an explicit input so that a device path can be fed the same draw."""
import torch
import torch.nn as nn


class SpeechEncoderPostnet(nn.Module):
    """speech_encoder_postnet.py:26-124 (no target GLU): project the masked / unmasked frames, score them against
    every label embedding by cosine similarity / temperature; class 0 of each row is the true label
    (compute_nce :61-74; a negative identical to the positive is set to -inf)."""

    def __init__(self, num_classes, encoder_embed_dim=768, final_dim=256, logit_temp=0.1, untie_final_proj=True,
                 skip_masked=False, skip_nomask=False):
        super().__init__()
        self.num_classes = list(num_classes)
        self.logit_temp, self.untie_final_proj = logit_temp, untie_final_proj
        self.skip_masked, self.skip_nomask = skip_masked, skip_nomask
        self.label_embs_concat = nn.Parameter(torch.empty(sum(self.num_classes), final_dim).uniform_())
        self.final_proj = nn.Linear(encoder_embed_dim, final_dim * (len(self.num_classes) if untie_final_proj else 1))

    def compute_nce(self, x, pos, negs):
        neg_is_pos = (pos == negs).all(-1)
        targets = torch.cat([pos.unsqueeze(0), negs], dim=0)
        logits = torch.cosine_similarity(x.float(), targets.float(), dim=-1).type_as(x) / self.logit_temp
        if neg_is_pos.any():
            logits[1:][neg_is_pos] = float("-inf")
        return logits.transpose(0, 1)  # [frames, 1 + num_classes]

    def forward(self, x, padding_mask, mask_indices, target_list):
        label_embs_list = self.label_embs_concat.split(self.num_classes, 0)

        def head(sel):
            proj = self.final_proj(x[sel])
            chunks = proj.chunk(len(target_list), dim=-1) if self.untie_final_proj else [proj] * len(target_list)
            out = []
            for i, (px, t) in enumerate(zip(chunks, target_list)):
                y = torch.index_select(label_embs_list[i], 0, t[sel].long())
                negs = label_embs_list[i].unsqueeze(1).expand(-1, px.size(0), -1)
                out.append(self.compute_nce(px, y, negs))
            return out

        m = None if self.skip_masked else head(torch.logical_and(~padding_mask, mask_indices))
        u = None if self.skip_nomask else head(torch.logical_and(~padding_mask, ~mask_indices))
        return {"logit_m_list": m if m is not None else [None for _ in target_list],
                "logit_u_list": u if u is not None else [None for _ in target_list], "padding_mask": padding_mask}


class GumbelVectorQuantizer(nn.Module):
    """gumbel_vector_quantizer.py:13-202, time-first input, weight_proj_depth 1, groups not combined (the SpeechT5
    configuration: 100 variables x 2 groups, vq_dim = d; models/speecht5.py:177-190). `gumbel_noise` (shape
    [B*T*groups, num_vars], i.i.d. standard Gumbel) replaces the draw inside F.gumbel_softmax when training."""

    def __init__(self, dim=768, num_vars=100, temp=(2.0, 0.5, 0.999995), groups=2, vq_dim=768):
        super().__init__()
        assert vq_dim % groups == 0
        self.groups, self.num_vars, self.input_dim = groups, num_vars, dim
        self.vars = nn.Parameter(torch.empty(1, groups * num_vars, vq_dim // groups).uniform_())
        self.weight_proj = nn.Linear(dim, groups * num_vars)
        nn.init.normal_(self.weight_proj.weight, mean=0, std=1)
        nn.init.zeros_(self.weight_proj.bias)
        self.max_temp, self.min_temp, self.temp_decay = temp
        self.curr_temp = self.max_temp

    def set_num_updates(self, num_updates):
        self.curr_temp = max(self.max_temp * self.temp_decay ** num_updates, self.min_temp)

    def forward(self, x, gumbel_noise=None):
        result = {"num_vars": self.num_vars * self.groups}
        bsz, tsz, fsz = x.shape
        logits = self.weight_proj(x.reshape(-1, fsz)).view(bsz * tsz * self.groups, -1)
        k = logits.argmax(-1)
        hard_x = torch.zeros_like(logits).scatter_(-1, k.view(-1, 1), 1.0).view(bsz * tsz, self.groups, -1)
        hard_probs = hard_x.float().mean(dim=0)
        result["code_perplexity"] = torch.exp(-torch.sum(hard_probs * torch.log(hard_probs + 1e-7), dim=-1)).sum()
        avg_probs = torch.softmax(logits.view(bsz * tsz, self.groups, -1).float(), dim=-1).mean(dim=0)
        result["prob_perplexity"] = torch.exp(-torch.sum(avg_probs * torch.log(avg_probs + 1e-7), dim=-1)).sum()
        result["temp"] = self.curr_temp
        if self.training:
            if gumbel_noise is None:
                gumbel_noise = -torch.empty_like(logits, dtype=torch.float).exponential_().log()
            y_soft = ((logits.float() + gumbel_noise) / self.curr_temp).softmax(-1)
            idx = y_soft.argmax(-1, keepdim=True)
            y_hard = torch.zeros_like(y_soft).scatter_(-1, idx, 1.0)
            sel = (y_hard - y_soft.detach() + y_soft).type_as(logits)  # straight-through (F.gumbel_softmax hard=True)
        else:
            sel = hard_x
        sel = sel.view(bsz * tsz, -1)
        q = (sel.unsqueeze(-1) * self.vars).view(bsz * tsz, self.groups, self.num_vars, -1).sum(-2)
        result["x"] = q.view(bsz, tsz, -1)
        return result


def mix_codes(encoder_out_btc, q_x, codebook_prob, perm):
    """models/speecht5.py:866-872: a random `codebook_prob` fraction of the TIME steps (perm = the reference's
    torch.randperm(T) draw) is replaced by the quantized vectors."""
    T = q_x.size(1)
    w = q_x.new_zeros(T)
    w[perm[: int(T * codebook_prob)]] = 1.0
    return w.view(-1, 1) * q_x + (1.0 - w).view(-1, 1) * encoder_out_btc
