"""Import the UNMODIFIED reference modules from /root/reference (authoring container only).

TEST INFRASTRUCTURE -- never imported by the product path.  Only `oracle/make_golden.py`
uses this file, to pin the CPU restatement in `oracle/` against the real reference code and
to write the fixtures under `tests/golden/`.  `/root/reference` does not exist on the GPU
box, so nothing in tests/, bench.py or smoke() may import this module.

Stubs: the wheels `diffusers`, `conformer`, `omegaconf` are absent here and
there is no network.  The stubbed symbols restate diffusers==0.29.0 behaviour constrained by
the checkpoint key names (`attn1.to_q.weight`, `attn1.to_out.0.{weight,bias}`,
`ff.net.0.proj.*`, `ff.net.2.*`) -- call sites: reference
src/chatterbox/models/s3gen/matcha/transformer.py:5-14,196-204 and matcha/decoder.py:7-8.
"""
import sys
import types
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_SRC = "/root/reference/src"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Attention(nn.Module):
    """diffusers.models.attention_processor.Attention (0.29.0) restricted to what
    matcha/transformer.py:196-204 constructs: self-attention, no bias on q/k/v, bias on out,
    AttnProcessor2_0 => F.scaled_dot_product_attention with additive mask."""

    def __init__(self, query_dim, heads=8, dim_head=64, dropout=0.0, bias=False,
                 cross_attention_dim=None, upcast_attention=False, **kw):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        B, T, _ = hidden_states.shape
        q = self.to_q(hidden_states)
        k = self.to_k(hidden_states)
        v = self.to_v(hidden_states)
        hd = q.shape[-1] // self.heads
        q = q.view(B, T, self.heads, hd).transpose(1, 2)
        k = k.view(B, T, self.heads, hd).transpose(1, 2)
        v = v.view(B, T, self.heads, hd).transpose(1, 2)
        if attention_mask is not None:
            # prepare_attention_mask: (B, 1|T, T) -> repeat_interleave(heads) -> view(B, heads, -1, T)
            am = attention_mask
            if am.dim() == 3:
                am = am.repeat_interleave(self.heads, dim=0).view(B, self.heads, -1, am.shape[-1])
            attention_mask = am
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, T, self.heads * hd)
        o = self.to_out[0](o)
        o = self.to_out[1](o)
        return o


class _GELU(nn.Module):
    """diffusers.models.activations.GELU: Linear + erf-GELU (approximate='none')."""

    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(self.proj(x), approximate=self.approximate)


class _Unreachable(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("unreachable with act_fn='gelu', norm_type='layer_norm'")


def _get_activation(name):
    return {"silu": nn.SiLU(), "swish": nn.SiLU(), "mish": nn.Mish(), "gelu": nn.GELU(), "relu": nn.ReLU()}[name]


_installed = False


def install():
    global _installed
    if _installed:
        return
    _installed = True
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    # empty package objects so chatterbox/__init__.py (needs librosa/perth + dist metadata) is skipped
    pk = _mod("chatterbox"); pk.__path__ = [REF_SRC + "/chatterbox"]
    pm = _mod("chatterbox.models"); pm.__path__ = [REF_SRC + "/chatterbox/models"]
    ps = _mod("chatterbox.models.s3gen"); ps.__path__ = [REF_SRC + "/chatterbox/models/s3gen"]
    # third-party stubs
    _mod("diffusers"); _mod("diffusers.models"); _mod("diffusers.utils")
    _mod("diffusers.models.attention", GEGLU=_Unreachable, GELU=_GELU, AdaLayerNorm=_Unreachable,
         AdaLayerNormZero=_Unreachable, ApproximateGELU=_Unreachable)
    _mod("diffusers.models.attention_processor", Attention=_Attention)
    _mod("diffusers.models.lora", LoRACompatibleLinear=nn.Linear)
    _mod("diffusers.models.activations", get_activation=_get_activation)
    _mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda c: c)
    _mod("conformer", ConformerBlock=_Unreachable)
    _mod("omegaconf", DictConfig=dict)
    # s3tokenizer package (absent): only constants are needed by s3gen imports
    # chatterbox.models.s3tokenizer/__init__ imports .s3tokenizer which imports the PyPI pkg;
    # we never import that path (flow/hifigan modules are imported directly).


def build_t3(multilingual=False):
    install()
    from chatterbox.models.t3.t3 import T3
    from chatterbox.models.t3.modules.t3_config import T3Config
    hp = T3Config.multilingual() if multilingual else T3Config.english_only()
    return T3(hp).eval()


def build_t3_turbo(text_vocab=50276, n_layers=24):
    """Turbo T3 exactly as tts_turbo.py:151-166 constructs it (GPT2_medium backbone).  `n_layers` / `text_vocab`
    only shrink the *config values* handed to the unmodified reference constructors (fewer identical blocks, a
    smaller gather table) so that fixtures stay small; 24 / 50276 are the shipped values."""
    install()
    from chatterbox.models.t3 import llama_configs
    from chatterbox.models.t3.t3 import T3
    from chatterbox.models.t3.modules.t3_config import T3Config
    hp = T3Config(text_tokens_dict_size=text_vocab)
    hp.llama_config_name = "GPT2_medium"
    hp.speech_tokens_dict_size = 6563
    hp.input_pos_emb = None
    hp.speech_cond_prompt_len = 375
    hp.use_perceiver_resampler = False
    hp.emotion_adv = False
    saved = llama_configs.LLAMA_CONFIGS["GPT2_medium"]
    try:
        llama_configs.LLAMA_CONFIGS["GPT2_medium"] = dict(saved, n_layer=n_layers, attn_pdrop=0.0, embd_pdrop=0.0,
                                                          resid_pdrop=0.0)
        t3 = T3(hp)
    finally:
        llama_configs.LLAMA_CONFIGS["GPT2_medium"] = saved
    return t3.eval()


def build_flow(meanflow=False):
    """CausalMaskedDiffWithXvec exactly as constructed in s3gen.py:64-104."""
    install()
    from chatterbox.models.s3gen.flow import CausalMaskedDiffWithXvec
    from chatterbox.models.s3gen.transformer.upsample_encoder import UpsampleConformerEncoder
    from chatterbox.models.s3gen.flow_matching import CausalConditionalCFM
    from chatterbox.models.s3gen.decoder import ConditionalDecoder
    from chatterbox.models.s3gen.configs import CFM_PARAMS
    encoder = UpsampleConformerEncoder(
        output_size=512, attention_heads=8, linear_units=2048, num_blocks=6, dropout_rate=0.1,
        positional_dropout_rate=0.1, attention_dropout_rate=0.1, normalize_before=True,
        input_layer='linear', pos_enc_layer_type='rel_pos_espnet', selfattention_layer_type='rel_selfattn',
        input_size=512, use_cnn_module=False, macaron_style=False)
    estimator = ConditionalDecoder(in_channels=320, out_channels=80, causal=True, channels=[256], dropout=0.0,
                                   attention_head_dim=64, n_blocks=4, num_mid_blocks=12, num_heads=8,
                                   act_fn='gelu', meanflow=meanflow)
    decoder = CausalConditionalCFM(spk_emb_dim=80, cfm_params=CFM_PARAMS, estimator=estimator)
    return CausalMaskedDiffWithXvec(encoder=encoder, decoder=decoder).eval()


def build_hift():
    """HiFTGenerator exactly as constructed in s3gen.py:244-252."""
    install()
    from chatterbox.models.s3gen.hifigan import HiFTGenerator
    from chatterbox.models.s3gen.f0_predictor import ConvRNNF0Predictor
    return HiFTGenerator(sampling_rate=24000, upsample_rates=[8, 5, 3], upsample_kernel_sizes=[16, 11, 7],
                         source_resblock_kernel_sizes=[7, 7, 11],
                         source_resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
                         f0_predictor=ConvRNNF0Predictor()).eval()
