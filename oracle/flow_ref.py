"""CPU fp32 restatement of S3Gen's token->mel path (TEST INFRASTRUCTURE - the oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this file.  Paths cited are relative to /root/reference/src/chatterbox/models/s3gen/.

The two third-party blocks inside the estimator (`diffusers==0.29.0` Attention + GELU feed-forward,
un-vendored, absent from this image; call sites matcha/transformer.py:196-204,100-118) are restated
from their published behaviour and constrained by the checkpoint key names.

Pinned: tests/golden/flow_*.pt come from the real reference modules (`CausalMaskedDiffWithXvec`,
`UpsampleConformerEncoder`, `ConditionalDecoder`, `CausalConditionalCFM`) run by oracle/make_golden.py.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- encoder
def espnet_rel_pos_emb(T, d_model=512):
    """transformer/embedding.py:229-294: pe = [flip(pe_positive) | pe_negative[1:]], sliced to the
    2T-1 positions centred on 0 -> row p corresponds to relative position (T-1-p)."""
    position = torch.arange(0, T, dtype=torch.float32).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2, dtype=torch.float32) * -(math.log(10000.0) / d_model))
    pe_pos = torch.zeros(T, d_model)
    pe_neg = torch.zeros(T, d_model)
    pe_pos[:, 0::2] = torch.sin(position * div_term)
    pe_pos[:, 1::2] = torch.cos(position * div_term)
    pe_neg[:, 0::2] = torch.sin(-1 * position * div_term)
    pe_neg[:, 1::2] = torch.cos(-1 * position * div_term)
    return torch.cat([torch.flip(pe_pos, [0]), pe_neg[1:]], dim=0)[None]          # [1, 2T-1, d]


def rel_shift(x):
    """transformer/attention.py:225-247."""
    zero_pad = torch.zeros((x.size(0), x.size(1), x.size(2), 1), dtype=x.dtype)
    x_padded = torch.cat([zero_pad, x], dim=-1)
    x_padded = x_padded.view(x.size(0), x.size(1), x.size(3) + 1, x.size(2))
    return x_padded[:, :, 1:].view_as(x)[:, :, :, : x.size(-1) // 2 + 1]


class FlowOracle:
    def __init__(self, sd, meanflow=False):
        self.sd = sd
        self.meanflow = meanflow

    # ---- conformer encoder ---------------------------------------------------------------
    def _embed(self, p, x):
        """transformer/subsampling.py LinearNoSubsampling :82-113 + embedding.py:268 (x*sqrt(512))."""
        sd = self.sd
        x = F.linear(x, sd[p + "out.0.weight"], sd[p + "out.0.bias"])
        x = F.layer_norm(x, (512,), sd[p + "out.1.weight"], sd[p + "out.1.bias"], 1e-5)
        return x * math.sqrt(512.0), espnet_rel_pos_emb(x.shape[1])

    def _enc_layer(self, p, x, pos_emb, mask):
        """transformer/encoder_layer.py:160-236 (normalize_before, no macaron, no conv module) with
        RelPositionMultiHeadedAttention.forward transformer/attention.py:249-330 and
        forward_attention :86-130.  mask [B,1,T] bool."""
        sd = self.sd
        B, T, _ = x.shape
        res = x
        h = F.layer_norm(x, (512,), sd[p + "norm_mha.weight"], sd[p + "norm_mha.bias"], 1e-12)
        a = p + "self_attn."
        q = F.linear(h, sd[a + "linear_q.weight"], sd[a + "linear_q.bias"]).view(B, T, 8, 64)
        k = F.linear(h, sd[a + "linear_k.weight"], sd[a + "linear_k.bias"]).view(B, T, 8, 64).transpose(1, 2)
        v = F.linear(h, sd[a + "linear_v.weight"], sd[a + "linear_v.bias"]).view(B, T, 8, 64).transpose(1, 2)
        pe = F.linear(pos_emb, sd[a + "linear_pos.weight"]).view(1, -1, 8, 64).transpose(1, 2)
        q_u = (q + sd[a + "pos_bias_u"]).transpose(1, 2)
        q_v = (q + sd[a + "pos_bias_v"]).transpose(1, 2)
        ac = torch.matmul(q_u, k.transpose(-2, -1))
        bd = rel_shift(torch.matmul(q_v, pe.transpose(-2, -1)))
        scores = (ac + bd) / math.sqrt(64)
        m = mask.unsqueeze(1).eq(0)
        scores = scores.masked_fill(m, -float("inf"))
        attn = torch.softmax(scores, dim=-1).masked_fill(m, 0.0)
        o = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, T, 512)
        x = res + F.linear(o, sd[a + "linear_out.weight"], sd[a + "linear_out.bias"])
        res = x
        h = F.layer_norm(x, (512,), sd[p + "norm_ff.weight"], sd[p + "norm_ff.bias"], 1e-12)
        h = F.linear(F.silu(F.linear(h, sd[p + "feed_forward.w_1.weight"], sd[p + "feed_forward.w_1.bias"])),
                     sd[p + "feed_forward.w_2.weight"], sd[p + "feed_forward.w_2.bias"])
        return res + h

    def encoder(self, xs, xs_lens):
        """transformer/upsample_encoder.py:237-304."""
        sd, e = self.sd, "encoder."
        T = xs.size(1)
        masks = (torch.arange(T)[None] < xs_lens[:, None]).unsqueeze(1)
        xs, pos_emb = self._embed(e + "embed.", xs)
        # PreLookaheadLayer :84-96
        o = xs.transpose(1, 2)
        o = F.pad(o, (0, 3))
        o = F.leaky_relu(F.conv1d(o, sd[e + "pre_lookahead_layer.conv1.weight"], sd[e + "pre_lookahead_layer.conv1.bias"]))
        o = F.pad(o, (2, 0))
        o = F.conv1d(o, sd[e + "pre_lookahead_layer.conv2.weight"], sd[e + "pre_lookahead_layer.conv2.bias"])
        xs = o.transpose(1, 2) + xs
        for i in range(6):
            xs = self._enc_layer(e + f"encoders.{i}.", xs, pos_emb, masks)
        # Upsample1D :59-63
        o = xs.transpose(1, 2)
        o = F.interpolate(o, scale_factor=2.0, mode="nearest")
        o = F.pad(o, (4, 0))
        o = F.conv1d(o, sd[e + "up_layer.conv.weight"], sd[e + "up_layer.conv.bias"])
        xs = o.transpose(1, 2)
        xs_lens = xs_lens * 2
        T = xs.size(1)
        masks = (torch.arange(T)[None] < xs_lens[:, None]).unsqueeze(1)
        xs, pos_emb = self._embed(e + "up_embed.", xs)
        for i in range(4):
            xs = self._enc_layer(e + f"up_encoders.{i}.", xs, pos_emb, masks)
        xs = F.layer_norm(xs, (512,), sd[e + "after_norm.weight"], sd[e + "after_norm.bias"], 1e-5)
        return xs, masks

    # ---- CFM estimator -------------------------------------------------------------------
    def _time_emb(self, t):
        """matcha/decoder.py SinusoidalPosEmb :20-29 (scale 1000, dim 320, [sin|cos]) +
        TimestepEmbedding :103-117 (Linear, SiLU, Linear)."""
        sd, e = self.sd, "decoder.estimator."
        half = 160
        emb = math.log(10000) / (half - 1)
        emb = torch.exp(torch.arange(half).float() * -emb)
        emb = 1000 * t.unsqueeze(1) * emb.unsqueeze(0)
        emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
        h = F.silu(F.linear(emb, sd[e + "time_mlp.linear_1.weight"], sd[e + "time_mlp.linear_1.bias"]))
        return F.linear(h, sd[e + "time_mlp.linear_2.weight"], sd[e + "time_mlp.linear_2.bias"])

    def _causal_block(self, p, x, mask):
        """decoder.py CausalBlock1D :49-63 (causal conv k3 left-pad 2 -> LayerNorm over C -> Mish)."""
        sd = self.sd
        h = F.conv1d(F.pad(x * mask, (2, 0)), sd[p + "block.0.weight"], sd[p + "block.0.bias"])
        h = F.layer_norm(h.transpose(1, 2), (h.shape[1],), sd[p + "block.2.weight"], sd[p + "block.2.bias"], 1e-5)
        return F.mish(h.transpose(1, 2)) * mask

    def _resnet(self, p, x, mask, t):
        """matcha/decoder.py ResnetBlock1D.forward :56-61 with the causal blocks (decoder.py:66-70)."""
        sd = self.sd
        h = self._causal_block(p + "block1.", x, mask)
        h = h + F.linear(F.mish(t), sd[p + "mlp.1.weight"], sd[p + "mlp.1.bias"]).unsqueeze(-1)
        h = self._causal_block(p + "block2.", h, mask)
        return h + F.conv1d(x * mask, sd[p + "res_conv.weight"], sd[p + "res_conv.bias"])

    def _tfmr(self, p, x, bias):
        """matcha/transformer.py BasicTransformerBlock.forward :243-316 (layer_norm, self-attn only,
        GELU-erf FF); diffusers Attention/AttnProcessor2_0: q,k,v 256->512 no bias, 8x64, SDPA with
        additive mask, out 512->256 with bias."""
        sd = self.sd
        B, T, _ = x.shape
        h = F.layer_norm(x, (256,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        sp = lambda w: F.linear(h, sd[p + w]).view(B, T, 8, 64).transpose(1, 2)
        o = F.scaled_dot_product_attention(sp("attn1.to_q.weight"), sp("attn1.to_k.weight"),
                                           sp("attn1.to_v.weight"), attn_mask=bias)
        o = o.transpose(1, 2).reshape(B, T, 512)
        x = x + F.linear(o, sd[p + "attn1.to_out.0.weight"], sd[p + "attn1.to_out.0.bias"])
        h = F.layer_norm(x, (256,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], 1e-5)
        h = F.gelu(F.linear(h, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"]))
        return x + F.linear(h, sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"])

    def estimator(self, x, mask, mu, t, spks, cond, r=None):
        """decoder.py ConditionalDecoder.forward :243-333 for channels=[256] (one down, 12 mid, one up;
        no resampling; full, non-causal attention with key-padding bias -1e10)."""
        sd, e = self.sd, "decoder.estimator."
        temb = self._time_emb(t)
        if self.meanflow:
            remb = self._time_emb(r)
            temb = F.linear(torch.cat([temb, remb], dim=1), sd[e + "time_embed_mixer.weight"])
        T = x.shape[-1]
        x = torch.cat([x, mu, spks.unsqueeze(-1).expand(-1, -1, T), cond], dim=1)    # 320 ch :270-276
        bias = ((1.0 - mask) * -1.0e10).unsqueeze(1)                                # [B,1,1,T] :26-34
        def stage(pfx, x, n_t=4):
            x = self._resnet(pfx + "0.", x, mask, temb)
            h = x.transpose(1, 2)
            for j in range(n_t):
                h = self._tfmr(pfx + f"1.{j}.", h, bias)
            return h.transpose(1, 2)
        x = stage(e + "down_blocks.0.", x)
        skip = x
        x = F.conv1d(F.pad(x * mask, (2, 0)), sd[e + "down_blocks.0.2.weight"], sd[e + "down_blocks.0.2.bias"])
        for i in range(12):
            x = stage(e + f"mid_blocks.{i}.", x)
        x = torch.cat([x, skip], dim=1)
        x = stage(e + "up_blocks.0.", x)
        x = F.conv1d(F.pad(x * mask, (2, 0)), sd[e + "up_blocks.0.2.weight"], sd[e + "up_blocks.0.2.bias"])
        x = self._causal_block(e + "final_block.", x, mask)
        out = F.conv1d(x * mask, sd[e + "final_proj.weight"], sd[e + "final_proj.bias"])
        return out * mask

    # ---- solver --------------------------------------------------------------------------
    def solve(self, z, mu, mask, spks, cond, n_timesteps=10, cfg_rate=0.7):
        """flow_matching.py CausalConditionalCFM.forward :195-233 + solve_euler :78-145 (cosine grid,
        CFG batch 2B with zeroed mu/spks/cond, x += dt * ((1+w) v_c - w v_u));
        meanflow: basic_euler :235-246 (linear grid, no CFG)."""
        t_span = torch.linspace(0, 1, n_timesteps + 1, dtype=mu.dtype)
        if not self.meanflow:
            t_span = 1 - torch.cos(t_span * 0.5 * torch.pi)
        x = z
        B = mu.size(0)
        for t, r in zip(t_span[:-1], t_span[1:]):
            t1, r1 = t[None], r[None]
            if self.meanflow:
                dxdt = self.estimator(x, mask, mu, t1.expand(B), spks, cond, r1.expand(B))
            else:
                x_in = torch.cat([x, x]); mask_in = torch.cat([mask, mask])
                mu_in = torch.cat([mu, torch.zeros_like(mu)])
                sp_in = torch.cat([spks, torch.zeros_like(spks)])
                c_in = torch.cat([cond, torch.zeros_like(cond)])
                d = self.estimator(x_in, mask_in, mu_in, t1.expand(2 * B), sp_in, c_in)
                dxdt = (1.0 + cfg_rate) * d[:B] - cfg_rate * d[B:]
            x = x + (r - t) * dxdt
        return x

    # ---- flow.inference ------------------------------------------------------------------
    def encode(self, token, prompt_token, prompt_feat, embedding):
        """flow.py:131-185 up to the decoder call. token [1,N], prompt_token [1,Np], prompt_feat [1,2Np,80],
        embedding [1,192] -> (mu [1,80,T], spks [1,80], cond [1,80,T], mask [1,1,T], mel_len1)."""
        sd = self.sd
        emb = F.linear(F.normalize(embedding, dim=1), sd["spk_embed_affine_layer.weight"],
                       sd["spk_embed_affine_layer.bias"])
        tok = torch.cat([prompt_token, token], dim=1)
        tok_len = torch.tensor([tok.shape[1]])
        x = sd["input_embedding.weight"][tok.long()]
        h, h_masks = self.encoder(x, tok_len)
        mel_len1 = prompt_feat.shape[1]
        h = F.linear(h, sd["encoder_proj.weight"], sd["encoder_proj.bias"])
        T = h.shape[1]
        conds = torch.zeros(1, T, 80)
        conds[:, :mel_len1] = prompt_feat
        mask = torch.ones(1, 1, T)
        return h.transpose(1, 2).contiguous(), emb, conds.transpose(1, 2).contiguous(), mask, mel_len1

    @torch.inference_mode()
    def inference(self, token, ref_dict, n_timesteps=None, z=None, noised_mels=None, finalize=True):
        """s3gen.py flow_inference :301-321 -> flow.py inference :131-198.  If `z` is None it is drawn
        with torch.randn_like on the global generator exactly like flow_matching.py:216.
        finalize=False (streaming chunk, flow.py:170-171): the last pre_lookahead_len * token_mel_ratio = 6 frames of the
        encoder output are dropped before the decoder.  The reference leaves `mask` at the untruncated length there (a
        shape error for any input, flow.py:173-183); this restates the evident intent: every tensor is 6 frames shorter."""
        n_timesteps = n_timesteps or (2 if self.meanflow else 10)
        mu, spks, cond, mask, mel_len1 = self.encode(torch.atleast_2d(token), ref_dict["prompt_token"],
                                                     ref_dict["prompt_feat"], ref_dict["embedding"])
        if not finalize:
            mu, cond, mask = mu[..., :-6], cond[..., :-6], mask[..., :-6]
        if z is None:
            z = torch.randn_like(mu)
        if noised_mels is not None:
            z = z.clone()
            z[..., mu.size(2) - noised_mels.size(2):] = noised_mels
        feat = self.solve(z, mu, mask, spks, cond, n_timesteps)
        return feat[:, :, mel_len1:]
