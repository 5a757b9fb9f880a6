"""CPU fp32 restatement of the HiFT vocoder (mel -> 24 kHz waveform) (TEST INFRASTRUCTURE - the oracle).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this file.  Paths cited: /root/reference/src/chatterbox/models/s3gen/{hifigan,f0_predictor,s3gen}.py.

Pinned: tests/golden/hift_*.pt come from the real `HiFTGenerator` run by oracle/make_golden.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .weights import fold_weight_norm

UPS = ((8, 16), (5, 11), (3, 7))     # s3gen.py:247-248 (upsample_rates, upsample_kernel_sizes)
RES_K = (3, 7, 11)                   # hifigan.py:308 default resblock_kernel_sizes
DIL = (1, 3, 5)
SRC_K = (7, 7, 11)                   # s3gen.py:249
SR = 24000
N_FFT, HOP = 16, 4                   # hifigan.py:307


def hann16():
    """scipy.signal.get_window('hann', 16, fftbins=True) (hifigan.py:388) == periodic hann."""
    n = torch.arange(N_FFT, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2 * math.pi * n / N_FFT)).to(torch.float32)


def snake(x, alpha):
    """hifigan.py Snake.forward :73-84."""
    a = alpha[None, :, None]
    return x + (1.0 / (a + 1e-9)) * torch.pow(torch.sin(x * a), 2)


class HiFTOracle:
    def __init__(self, sd):
        self.sd = fold_weight_norm(sd)      # weight-norm is recomputed every call in the reference
        self.window = hann16()

    def f0_predictor(self, mel):
        """f0_predictor.py:52-55."""
        sd, x = self.sd, mel
        for i in (0, 2, 4, 6, 8):
            x = F.elu(F.conv1d(x, sd[f"f0_predictor.condnet.{i}.weight"], sd[f"f0_predictor.condnet.{i}.bias"], padding=1))
        x = x.transpose(1, 2)
        return torch.abs(F.linear(x, sd["f0_predictor.classifier.weight"], sd["f0_predictor.classifier.bias"]).squeeze(-1))

    def source(self, f0, phase_vec=None, noise=None):
        """hifigan.py SineGen.forward :200-231 + SourceModuleHnNSF.forward :267-283.
        f0 [B, T] (Hz per mel frame).  phase_vec [B,9,1] / noise [B,9,L] are drawn from the global
        generator in the reference's order when not given.  Returns s [B,1,L]."""
        sd = self.sd
        f0u = F.interpolate(f0[:, None], scale_factor=480.0, mode="nearest")          # hifigan.py:329 -> [B,1,L]
        B, _, L = f0u.shape
        F_mat = torch.zeros(B, 9, L)
        for i in range(9):
            F_mat[:, i:i + 1, :] = f0u * (i + 1) / SR
        theta = 2 * np.pi * (torch.cumsum(F_mat, dim=-1) % 1)                        # fp64-accumulated on CPU
        if phase_vec is None:
            from torch.distributions.uniform import Uniform
            phase_vec = Uniform(low=-np.pi, high=np.pi).sample(sample_shape=(B, 9, 1))
            phase_vec[:, 0, :] = 0
        sine = 0.1 * torch.sin(theta + phase_vec)
        uv = (f0u > 10).type(torch.float32)
        noise_amp = uv * 0.003 + (1 - uv) * 0.1 / 3
        if noise is None:
            noise = torch.randn_like(sine)
        sine = sine * uv + noise_amp * noise
        s = torch.tanh(F.linear(sine.transpose(1, 2), sd["m_source.l_linear.weight"], sd["m_source.l_linear.bias"]))
        return s.transpose(1, 2)

    def _resblock(self, p, x):
        """hifigan.py ResBlock.forward :154-161."""
        sd = self.sd
        k = sd[p + "convs1.0.weight"].shape[-1]
        for j, d in enumerate(DIL):
            xt = snake(x, sd[p + f"activations1.{j}.alpha"])
            xt = F.conv1d(xt, sd[p + f"convs1.{j}.weight"], sd[p + f"convs1.{j}.bias"], dilation=d, padding=(k * d - d) // 2)
            xt = snake(xt, sd[p + f"activations2.{j}.alpha"])
            xt = F.conv1d(xt, sd[p + f"convs2.{j}.weight"], sd[p + f"convs2.{j}.bias"], padding=(k - 1) // 2)
            x = xt + x
        return x

    def decode(self, mel, s):
        """hifigan.py decode :412-444 with _stft :396-402 and _istft :404-410."""
        sd = self.sd
        spec = torch.stft(s.squeeze(1), N_FFT, HOP, N_FFT, window=self.window, return_complex=True)
        s_stft = torch.cat([spec.real, spec.imag], dim=1)                             # [B,18,120T+1]
        x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
        downs = ((15, 30, 7), (3, 6, 1), (1, 1, 0))                                   # hifigan.py:349-364
        for i, (u, k) in enumerate(UPS):
            x = F.leaky_relu(x, 0.1)
            x = F.conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
            if i == 2:
                x = F.pad(x, (1, 0), mode="reflect")
            st, kk, pd = downs[i]
            si = F.conv1d(s_stft, sd[f"source_downs.{i}.weight"], sd[f"source_downs.{i}.bias"], stride=st, padding=pd)
            si = self._resblock(f"source_resblocks.{i}.", si)
            x = x + si
            xs = None
            for j in range(3):
                r = self._resblock(f"resblocks.{i * 3 + j}.", x)
                xs = r if xs is None else xs + r
            x = xs / 3
        x = F.leaky_relu(x)
        x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
        mag = torch.exp(x[:, :9, :])
        phase = torch.sin(x[:, 9:, :])
        mag = torch.clip(mag, max=1e2)
        wav = torch.istft(torch.complex(mag * torch.cos(phase), mag * torch.sin(phase)), N_FFT, HOP, N_FFT, window=self.window)
        return torch.clamp(wav, -0.99, 0.99)

    @torch.inference_mode()
    def inference(self, mel, s=None, phase_vec=None, noise=None, trim_fade=True):
        """hifigan.py inference :462-474 + s3gen.py:254-258,359-360 (trim-fade of first 960 samples).
        Returns (wav [B, 480T], s [B,1,480T])."""
        if s is None:
            f0 = self.f0_predictor(mel)
            s = self.source(f0, phase_vec, noise)
        wav = self.decode(mel, s)
        if trim_fade:
            n_trim = SR // 50
            fade = torch.zeros(2 * n_trim)
            fade[n_trim:] = (torch.cos(torch.linspace(torch.pi, 0, n_trim)) + 1) / 2
            wav[:, :2 * n_trim] *= fade
        return wav, s
