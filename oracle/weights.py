"""Deterministic synthetic checkpoints with the reference's state-dict key names and shapes.

TEST INFRASTRUCTURE (oracle/): used by tests/, bench.py (cpu_baseline / --impl reference leg and to
feed both arms the same weights) and __graft_entry__.smoke().  No checkpoints exist on disk and there
is no network (SURVEY.md 8c "Weights"), so parity is pinned on seeded random-init weights:

* every tensor is drawn from its own `torch.Generator` (CPU, seed = f(seed, key)) so any subset can be
  regenerated bit-identically on any machine with the same torch build;
* matmul / conv weights are rounded once to bf16-representable fp32 values so that the fp32 oracle
  and the bf16-weight engine see *identical* weights (SURVEY.md 7 "Hard parts");
* HiFT weight-norm pairs (g, v) are built so that the folded weight g*v/||v|| is again
  bf16-representable: g = ||v|| * 2^k, k in {-1,0,1} (reference hifigan.py:88-100,
  torch.nn.utils.parametrizations.weight_norm, dim=0).

Key names / shapes follow the reference modules (checked with strict load_state_dict in
oracle/make_golden.py):
  T3      reference src/chatterbox/models/t3/t3.py:49-85, modules/cond_enc.py:41-62,
          modules/perceiver.py:114-196, transformers LlamaModel
  flow    reference src/chatterbox/models/s3gen/flow.py:42-84, transformer/upsample_encoder.py:85-233,
          decoder.py:98-226
  hift    reference src/chatterbox/models/s3gen/hifigan.py:286-394, f0_predictor.py:19-50
"""
import hashlib
import math
from collections import OrderedDict

import torch


def _gen(seed, key):
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _randn(seed, key, shape, std=1.0, mean=0.0, bf16=False):
    t = torch.randn(*shape, generator=_gen(seed, key), dtype=torch.float32) * std + mean
    return bf16_round(t) if bf16 else t


# ----------------------------------------------------------------------------- T3
T3_DIM, T3_LAYERS, T3_HEADS, T3_FFN = 1024, 30, 16, 4096
SPEECH_VOCAB = 8194


def t3_spec(text_vocab=704, n_layers=T3_LAYERS):
    """(key, shape, kind) in a fixed order. kind: w=matmul weight (bf16-representable), n=norm weight,
    b=bias/small vector, e=embedding table, p=position table."""
    D = T3_DIM
    s = []
    s.append(("tfmr.embed_tokens.weight", (8, D), "e"))
    for i in range(n_layers):
        p = f"tfmr.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            s.append((p + f"self_attn.{n}.weight", (D, D), "w"))
        s.append((p + "mlp.gate_proj.weight", (T3_FFN, D), "w"))
        s.append((p + "mlp.up_proj.weight", (T3_FFN, D), "w"))
        s.append((p + "mlp.down_proj.weight", (D, T3_FFN), "w"))
        s.append((p + "input_layernorm.weight", (D,), "n"))
        s.append((p + "post_attention_layernorm.weight", (D,), "n"))
    s.append(("tfmr.norm.weight", (D,), "n"))
    s.append(("cond_enc.spkr_enc.weight", (D, 256), "w"))
    s.append(("cond_enc.spkr_enc.bias", (D,), "b"))
    s.append(("cond_enc.emotion_adv_fc.weight", (D, 1), "w1"))
    s.append(("cond_enc.perceiver.pre_attention_query", (1, 32, D), "q"))
    s.append(("cond_enc.perceiver.attn.norm.weight", (D,), "n"))
    s.append(("cond_enc.perceiver.attn.norm.bias", (D,), "b"))
    for n in ("to_q", "to_k", "to_v", "proj_out"):
        s.append((f"cond_enc.perceiver.attn.{n}.weight", (D, D), "w"))
        s.append((f"cond_enc.perceiver.attn.{n}.bias", (D,), "b"))
    s.append(("text_emb.weight", (text_vocab, D), "e"))
    s.append(("speech_emb.weight", (SPEECH_VOCAB, D), "e"))
    s.append(("text_pos_emb.emb.weight", (2048 + 2, D), "p"))
    s.append(("speech_pos_emb.emb.weight", (4096 + 4, D), "p"))
    s.append(("text_head.weight", (text_vocab, D), "w"))
    s.append(("speech_head.weight", (SPEECH_VOCAB, D), "h"))
    return s


def make_t3_weights(seed=0, text_vocab=704, n_layers=T3_LAYERS, head_std=0.06, bf16=True):
    """bf16=False keeps the matmul weights as drawn (NOT bf16-representable): the shape of a real fp32 checkpoint, used to
    measure what rounding the weights to bf16 (north_star: bf16 tensor-core operands) costs against the fp32 reference."""
    sd = OrderedDict()
    for key, shape, kind in t3_spec(text_vocab, n_layers):
        if kind == "w":
            sd[key] = _randn(seed, key, shape, std=0.7 / math.sqrt(shape[-1]), bf16=bf16)
        elif kind == "w1":
            sd[key] = _randn(seed, key, shape, std=0.3, bf16=bf16)
        elif kind == "h":
            w = _randn(seed, key, shape, std=head_std, bf16=bf16)
            # ids >= 6561 (SOS/EOS/unused) are dropped by generate() (tts.py:257-262); a trained model almost never
            # emits them, so the synthetic head keeps their logits near zero (utterance length is then set by
            # max_new_tokens) -- still bf16-representable (power-of-two scale)
            w[6561:] = w[6561:] * (2.0 ** -6)
            sd[key] = w
        elif kind == "n":
            sd[key] = _randn(seed, key, shape, std=0.1, mean=1.0)
        elif kind == "b":
            sd[key] = _randn(seed, key, shape, std=0.02)
        elif kind == "e":
            sd[key] = _randn(seed, key, shape, std=0.5)
        elif kind == "p":
            sd[key] = _randn(seed, key, shape, std=0.1)
        elif kind == "q":
            sd[key] = _randn(seed, key, shape, std=0.3)
        else:
            raise ValueError(kind)
    return sd


# ----------------------------------------------------------------------------- T3 Turbo (GPT-2 backbone)
TURBO_SPEECH_VOCAB = 6563
TURBO_LAYERS = 24


def t3_turbo_spec(text_vocab=50276, n_layers=TURBO_LAYERS):
    """Turbo T3 (reference tts_turbo.py:151-159: T3Config(text_tokens_dict_size=50276), GPT2_medium backbone
    llama_configs.py:35-68, speech vocab 6563, no learned input position tables, no perceiver, no emotion).
    GPT-2 `Conv1D` weights are stored [in, out] (transformers pytorch_utils.Conv1D)."""
    D = T3_DIM
    s = [("tfmr.wpe.weight", (8196, D), "p")]
    for i in range(n_layers):
        p = f"tfmr.h.{i}."
        s.append((p + "ln_1.weight", (D,), "n")); s.append((p + "ln_1.bias", (D,), "b"))
        s.append((p + "attn.c_attn.weight", (D, 3 * D), "wt")); s.append((p + "attn.c_attn.bias", (3 * D,), "b"))
        s.append((p + "attn.c_proj.weight", (D, D), "wt")); s.append((p + "attn.c_proj.bias", (D,), "b"))
        s.append((p + "ln_2.weight", (D,), "n")); s.append((p + "ln_2.bias", (D,), "b"))
        s.append((p + "mlp.c_fc.weight", (D, 4 * D), "wt")); s.append((p + "mlp.c_fc.bias", (4 * D,), "b"))
        s.append((p + "mlp.c_proj.weight", (4 * D, D), "wt")); s.append((p + "mlp.c_proj.bias", (D,), "b"))
    s.append(("tfmr.ln_f.weight", (D,), "n")); s.append(("tfmr.ln_f.bias", (D,), "b"))
    s.append(("cond_enc.spkr_enc.weight", (D, 256), "w"))
    s.append(("cond_enc.spkr_enc.bias", (D,), "b"))
    s.append(("text_emb.weight", (text_vocab, D), "e"))
    s.append(("speech_emb.weight", (TURBO_SPEECH_VOCAB, D), "e"))
    s.append(("text_head.weight", (text_vocab, D), "w"))
    s.append(("speech_head.weight", (TURBO_SPEECH_VOCAB, D), "h"))
    s.append(("speech_head.bias", (TURBO_SPEECH_VOCAB,), "b"))
    return s


def make_t3_turbo_weights(seed=0, text_vocab=50276, n_layers=TURBO_LAYERS, head_std=0.06):
    """Seeded Turbo checkpoint with the reference key names (`tfmr.wte.weight` is left out: the reference deletes it
    right after loading, tts_turbo.py:166, and it is never read on the inputs_embeds path)."""
    sd = OrderedDict()
    for key, shape, kind in t3_turbo_spec(text_vocab, n_layers):
        if kind == "w":
            sd[key] = _randn(seed, "turbo." + key, shape, std=0.7 / math.sqrt(shape[-1]), bf16=True)
        elif kind == "wt":      # Conv1D [in, out]: fan-in is shape[0]
            sd[key] = _randn(seed, "turbo." + key, shape, std=0.7 / math.sqrt(shape[0]), bf16=True)
        elif kind == "h":
            w = _randn(seed, "turbo." + key, shape, std=head_std, bf16=True)
            w[6561:] = w[6561:] * (2.0 ** -6)      # BOS/EOS logits near zero (see make_t3_weights)
            sd[key] = w
        elif kind == "n":
            sd[key] = _randn(seed, "turbo." + key, shape, std=0.1, mean=1.0)
        elif kind == "b":
            sd[key] = _randn(seed, "turbo." + key, shape, std=0.02)
        elif kind == "e":
            sd[key] = _randn(seed, "turbo." + key, shape, std=0.5)
        elif kind == "p":
            sd[key] = _randn(seed, "turbo." + key, shape, std=0.1)
        else:
            raise ValueError(kind)
    return sd


# ----------------------------------------------------------------------------- flow (encoder + CFM estimator)
def flow_spec(meanflow=False):
    s = []
    s.append(("input_embedding.weight", (6561, 512), "e"))
    s.append(("spk_embed_affine_layer.weight", (80, 192), "w"))
    s.append(("spk_embed_affine_layer.bias", (80,), "b"))

    def embed(prefix):
        s.append((prefix + "out.0.weight", (512, 512), "w"))
        s.append((prefix + "out.0.bias", (512,), "b"))
        s.append((prefix + "out.1.weight", (512,), "n"))
        s.append((prefix + "out.1.bias", (512,), "b"))

    def enc_layer(prefix):
        s.append((prefix + "self_attn.pos_bias_u", (8, 64), "u"))
        s.append((prefix + "self_attn.pos_bias_v", (8, 64), "u"))
        for n in ("linear_q", "linear_k", "linear_v", "linear_out"):
            s.append((prefix + f"self_attn.{n}.weight", (512, 512), "w"))
            s.append((prefix + f"self_attn.{n}.bias", (512,), "b"))
        s.append((prefix + "self_attn.linear_pos.weight", (512, 512), "w"))
        s.append((prefix + "feed_forward.w_1.weight", (2048, 512), "w"))
        s.append((prefix + "feed_forward.w_1.bias", (2048,), "b"))
        s.append((prefix + "feed_forward.w_2.weight", (512, 2048), "w"))
        s.append((prefix + "feed_forward.w_2.bias", (512,), "b"))
        s.append((prefix + "norm_ff.weight", (512,), "n"))
        s.append((prefix + "norm_ff.bias", (512,), "b"))
        s.append((prefix + "norm_mha.weight", (512,), "n"))
        s.append((prefix + "norm_mha.bias", (512,), "b"))

    embed("encoder.embed.")
    s.append(("encoder.after_norm.weight", (512,), "n"))
    s.append(("encoder.after_norm.bias", (512,), "b"))
    s.append(("encoder.pre_lookahead_layer.conv1.weight", (512, 512, 4), "w"))
    s.append(("encoder.pre_lookahead_layer.conv1.bias", (512,), "b"))
    s.append(("encoder.pre_lookahead_layer.conv2.weight", (512, 512, 3), "w"))
    s.append(("encoder.pre_lookahead_layer.conv2.bias", (512,), "b"))
    for i in range(6):
        enc_layer(f"encoder.encoders.{i}.")
    s.append(("encoder.up_layer.conv.weight", (512, 512, 5), "w"))
    s.append(("encoder.up_layer.conv.bias", (512,), "b"))
    embed("encoder.up_embed.")
    for i in range(4):
        enc_layer(f"encoder.up_encoders.{i}.")
    s.append(("encoder_proj.weight", (80, 512), "w"))
    s.append(("encoder_proj.bias", (80,), "b"))

    e = "decoder.estimator."
    s.append((e + "time_mlp.linear_1.weight", (1024, 320), "w"))
    s.append((e + "time_mlp.linear_1.bias", (1024,), "b"))
    s.append((e + "time_mlp.linear_2.weight", (1024, 1024), "w"))
    s.append((e + "time_mlp.linear_2.bias", (1024,), "b"))

    def resnet(prefix, cin):
        s.append((prefix + "mlp.1.weight", (256, 1024), "w"))
        s.append((prefix + "mlp.1.bias", (256,), "b"))
        s.append((prefix + "block1.block.0.weight", (256, cin, 3), "w"))
        s.append((prefix + "block1.block.0.bias", (256,), "b"))
        s.append((prefix + "block1.block.2.weight", (256,), "n"))
        s.append((prefix + "block1.block.2.bias", (256,), "b"))
        s.append((prefix + "block2.block.0.weight", (256, 256, 3), "w"))
        s.append((prefix + "block2.block.0.bias", (256,), "b"))
        s.append((prefix + "block2.block.2.weight", (256,), "n"))
        s.append((prefix + "block2.block.2.bias", (256,), "b"))
        s.append((prefix + "res_conv.weight", (256, cin, 1), "w"))
        s.append((prefix + "res_conv.bias", (256,), "b"))

    def tfmr(prefix):
        s.append((prefix + "norm1.weight", (256,), "n"))
        s.append((prefix + "norm1.bias", (256,), "b"))
        s.append((prefix + "attn1.to_q.weight", (512, 256), "w"))
        s.append((prefix + "attn1.to_k.weight", (512, 256), "w"))
        s.append((prefix + "attn1.to_v.weight", (512, 256), "w"))
        s.append((prefix + "attn1.to_out.0.weight", (256, 512), "w"))
        s.append((prefix + "attn1.to_out.0.bias", (256,), "b"))
        s.append((prefix + "norm3.weight", (256,), "n"))
        s.append((prefix + "norm3.bias", (256,), "b"))
        s.append((prefix + "ff.net.0.proj.weight", (1024, 256), "w"))
        s.append((prefix + "ff.net.0.proj.bias", (1024,), "b"))
        s.append((prefix + "ff.net.2.weight", (256, 1024), "w"))
        s.append((prefix + "ff.net.2.bias", (256,), "b"))

    resnet(e + "down_blocks.0.0.", 320)
    for j in range(4):
        tfmr(e + f"down_blocks.0.1.{j}.")
    s.append((e + "down_blocks.0.2.weight", (256, 256, 3), "w"))
    s.append((e + "down_blocks.0.2.bias", (256,), "b"))
    for i in range(12):
        resnet(e + f"mid_blocks.{i}.0.", 256)
        for j in range(4):
            tfmr(e + f"mid_blocks.{i}.1.{j}.")
    resnet(e + "up_blocks.0.0.", 512)
    for j in range(4):
        tfmr(e + f"up_blocks.0.1.{j}.")
    s.append((e + "up_blocks.0.2.weight", (256, 256, 3), "w"))
    s.append((e + "up_blocks.0.2.bias", (256,), "b"))
    s.append((e + "final_block.block.0.weight", (256, 256, 3), "w"))
    s.append((e + "final_block.block.0.bias", (256,), "b"))
    s.append((e + "final_block.block.2.weight", (256,), "n"))
    s.append((e + "final_block.block.2.bias", (256,), "b"))
    s.append((e + "final_proj.weight", (80, 256, 1), "w"))
    s.append((e + "final_proj.bias", (80,), "b"))
    if meanflow:
        s.append((e + "time_embed_mixer.weight", (1024, 2048), "w"))
    return s


def make_flow_weights(seed=0, meanflow=False):
    sd = OrderedDict()
    for key, shape, kind in flow_spec(meanflow):
        if kind == "w":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[key] = _randn(seed, key, shape, std=1.0 / math.sqrt(fan_in), bf16=True)
        elif kind == "n":
            sd[key] = _randn(seed, key, shape, std=0.1, mean=1.0)
        elif kind == "b":
            sd[key] = _randn(seed, key, shape, std=0.05)
        elif kind == "e":
            sd[key] = _randn(seed, key, shape, std=1.0)
        elif kind == "u":
            sd[key] = _randn(seed, key, shape, std=0.2)
        else:
            raise ValueError(kind)
    return sd


# ----------------------------------------------------------------------------- HiFT
def hift_spec():
    """(key, shape, kind).  kind 'wn' expands into (.bias, .parametrizations.weight.original0/1);
    'wnT' is the ConvTranspose1d flavour (dim 0 = in-channels)."""
    s = []
    s.append(("m_source.l_linear.weight", (1, 9), "lin9"))
    s.append(("m_source.l_linear.bias", (1,), "b"))
    s.append(("conv_pre", (512, 80, 7), "wn"))
    ups = [(512, 256, 16), (256, 128, 11), (128, 64, 7)]
    for i, sh in enumerate(ups):
        s.append((f"ups.{i}", sh, "wnT"))
    sd_shapes = [(256, 18, 30), (128, 18, 6), (64, 18, 1)]
    for i, sh in enumerate(sd_shapes):
        s.append((f"source_downs.{i}.weight", sh, "w"))
        s.append((f"source_downs.{i}.bias", (sh[0],), "b"))

    def resblock(prefix, ch, k):
        for c in ("convs1", "convs2"):
            for j in range(3):
                s.append((prefix + f"{c}.{j}", (ch, ch, k), "wn"))
        for a in ("activations1", "activations2"):
            for j in range(3):
                s.append((prefix + f"{a}.{j}.alpha", (ch,), "alpha"))

    for i, (ch, k) in enumerate(zip((256, 128, 64), (7, 7, 11))):
        resblock(f"source_resblocks.{i}.", ch, k)
    for i, ch in enumerate((256, 128, 64)):
        for j, k in enumerate((3, 7, 11)):
            resblock(f"resblocks.{i * 3 + j}.", ch, k)
    s.append(("conv_post", (18, 64, 7), "wn"))
    s.append(("f0_predictor.condnet.0", (512, 80, 3), "wn"))
    for i in (2, 4, 6, 8):
        s.append((f"f0_predictor.condnet.{i}", (512, 512, 3), "wn"))
    s.append(("f0_predictor.classifier.weight", (1, 512), "f0w"))
    s.append(("f0_predictor.classifier.bias", (1,), "f0b"))
    return s


def make_hift_weights(seed=0, gain=0.6):
    sd = OrderedDict()
    for key, shape, kind in hift_spec():
        if kind in ("wn", "wnT"):
            fan_in = shape[1] * shape[2] if kind == "wn" else shape[0] * shape[2] / 4.0
            g_ = gain if "conv_post" not in key else 0.3
            v = _randn(seed, key + ".v", shape, std=g_ / math.sqrt(fan_in), bf16=True)
            # norm over all dims but 0 -- exactly what weight_norm(dim=0) uses (also for ConvTranspose1d)
            norm = torch.norm_except_dim(v, 2, 0)
            k = torch.randint(-1, 2, (shape[0], 1, 1), generator=_gen(seed, key + ".k")).to(torch.float32)
            g = norm * torch.pow(torch.tensor(2.0), k)
            nb = shape[0] if kind == "wn" else shape[1]
            sd[key + ".bias"] = _randn(seed, key + ".bias", (nb,), std=0.05)
            sd[key + ".parametrizations.weight.original0"] = g
            sd[key + ".parametrizations.weight.original1"] = v
        elif kind == "w":
            fan_in = shape[1] * shape[2]
            sd[key] = _randn(seed, key, shape, std=gain / math.sqrt(fan_in), bf16=True)
        elif kind == "b":
            sd[key] = _randn(seed, key, shape, std=0.05)
        elif kind == "alpha":
            sd[key] = _randn(seed, key, shape, std=0.15, mean=1.0).abs() + 0.05
        elif kind == "lin9":
            sd[key] = _randn(seed, key, shape, std=0.6)
        elif kind == "f0w":
            sd[key] = _randn(seed, key, shape, std=8.0)
        elif kind == "f0b":
            sd[key] = torch.full(shape, 25.0)
        else:
            raise ValueError(kind)
    return sd


def fold_weight_norm(sd):
    """Fold every (original0, original1) pair into a plain `.weight` (load-time transform the engine
    applies too).  w = v * (g / ||v||), norm over all dims except 0 (reference hifigan.py Snake/ResBlock
    convs via torch.nn.utils.parametrizations.weight_norm default dim=0)."""
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith(".parametrizations.weight.original0"):
            base = k[: -len(".parametrizations.weight.original0")]
            g = v
            vv = sd[base + ".parametrizations.weight.original1"]
            out[base + ".weight"] = torch._weight_norm(vv, g, 0)
        elif k.endswith(".parametrizations.weight.original1"):
            continue
        else:
            out[k] = v
    return out


# ----------------------------------------------------------------------------- synthetic conditionals / inputs
def make_conds(seed=1234, n_t3_prompt=150, n_gen_prompt=250):
    """Synthetic voice conditionals: stands in for conds.pt /
    prepare_conditionals() output (reference tts.py:182-206)."""
    spk = torch.nn.functional.normalize(_randn(seed, "speaker_emb", (1, 256)), dim=-1)
    g = _gen(seed, "tokens")
    cond_prompt = torch.randint(0, 6561, (1, n_t3_prompt), generator=g)
    prompt_token = torch.randint(0, 6561, (1, n_gen_prompt), generator=g)
    prompt_feat = _randn(seed, "prompt_feat", (1, 2 * n_gen_prompt, 80), std=0.5, mean=-2.0)
    emb = _randn(seed, "xvec", (1, 192))
    t3 = dict(speaker_emb=spk, cond_prompt_speech_tokens=cond_prompt, emotion_adv=0.5 * torch.ones(1, 1, 1))
    gen = dict(prompt_token=prompt_token, prompt_token_len=torch.tensor([n_gen_prompt]),
               prompt_feat=prompt_feat, prompt_feat_len=None, embedding=emb)
    return t3, gen
