"""Oracle for the MDX23 window / STFT / iSTFT / stem arithmetic (test infrastructure).

* ``MdxGeometry`` / ``stft`` / ``istft`` restate the external MVSEP-MDX23
  ``Conv_TDF_net_trim_model`` that /root/reference/src/audio_cut/separation/backends.py:260-265,
  355, 376 instantiates and calls (it is NOT in the reference tree; algorithm as
  published).  ``torch.stft``/``torch.istft`` on CPU are the
  arithmetic the reference itself uses, so they are used here directly.
* ``infer_chunk`` restates backends.py:268-281 (``_prepare_input``) and :299-406
  (window build, trim/concat, crop, stem arithmetic, mono mean).

The network itself is passed in as a callable ``net(spec[B,4,dim_f,dim_t]) -> same``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class MdxGeometry:
    n_fft: int = 7680
    hop: int = 1024
    dim_f: int = 3072
    dim_t: int = 256

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def chunk_size(self) -> int:  # samples per model window
        return self.hop * (self.dim_t - 1)

    @property
    def trim(self) -> int:
        return self.n_fft // 2

    @property
    def gen(self) -> int:
        return self.chunk_size - 2 * self.trim


def stft(x: torch.Tensor, g: MdxGeometry) -> torch.Tensor:
    """(B,2,chunk_size) float32 -> (B,4,dim_f,dim_t); channels [L_re,L_im,R_re,R_im]."""
    window = torch.hann_window(g.n_fft, periodic=True, dtype=x.dtype)
    x = x.reshape(-1, g.chunk_size)
    X = torch.stft(x, g.n_fft, g.hop, window=window, center=True, return_complex=True)
    X = torch.view_as_real(X).permute(0, 3, 1, 2)
    X = X.reshape(-1, 2, 2, g.n_bins, g.dim_t).reshape(-1, 4, g.n_bins, g.dim_t)
    return X[:, :, : g.dim_f].contiguous()


def istft(y: torch.Tensor, g: MdxGeometry) -> torch.Tensor:
    """(B,4,dim_f,dim_t) -> (B,2,chunk_size)."""
    window = torch.hann_window(g.n_fft, periodic=True, dtype=y.dtype)
    pad = torch.zeros(y.shape[0], 4, g.n_bins - g.dim_f, g.dim_t, dtype=y.dtype)
    y = torch.cat([y, pad], dim=-2)
    y = y.reshape(-1, 2, 2, g.n_bins, g.dim_t).reshape(-1, 2, g.n_bins, g.dim_t)
    y = y.permute(0, 2, 3, 1).contiguous()
    Y = torch.view_as_complex(y)
    w = torch.istft(Y, g.n_fft, g.hop, window=window, center=True)
    return w.reshape(-1, 2, g.chunk_size)


def prepare_input(mix_chunk: np.ndarray, align_hop: int = 4096) -> Tuple[np.ndarray, int]:
    """backends.py:268-281."""
    if mix_chunk.ndim == 1:
        mix = np.stack([mix_chunk, mix_chunk], axis=0)
    elif mix_chunk.ndim == 2:
        mix = mix_chunk
    else:
        raise ValueError("mix_chunk shape")
    mix = np.ascontiguousarray(mix.astype(np.float32, copy=False))
    hop = max(1, int(align_hop))
    pad = (-mix.shape[-1]) % hop
    if pad:
        mix = np.pad(mix, ((0, 0), (0, pad)), mode="constant")
    return mix, pad


def build_windows(mix_stereo: np.ndarray, g: MdxGeometry) -> np.ndarray:
    """backends.py:306-330: [0_trim | mix | 0_pad | 0_trim], windows at stride gen."""
    L = mix_stereo.shape[-1]
    pad = (g.gen - L % g.gen) % g.gen
    padded = np.concatenate(
        (np.zeros((2, g.trim), np.float32), mix_stereo, np.zeros((2, pad), np.float32), np.zeros((2, g.trim), np.float32)),
        axis=1,
    )
    waves = []
    i = 0
    while i < L + pad:
        waves.append(padded[:, i : i + g.chunk_size])
        i += g.gen
    return np.stack(waves).astype(np.float32)


def n_windows(chunk_len: int, g: MdxGeometry, align_hop: int = 4096) -> int:
    L = chunk_len + ((-chunk_len) % max(1, align_hop))
    pad = (g.gen - L % g.gen) % g.gen
    return (L + pad) // g.gen


def infer_chunk(
    mix_chunk: np.ndarray,
    net: Callable[[torch.Tensor], torch.Tensor],
    g: MdxGeometry,
    *,
    align_hop: int = 4096,
    output_type: str = "vocal",
    dtype=torch.float32,
):
    """backends.py:299-406.  Returns (vocal_mono, instrumental_mono) float32."""
    mix_stereo, align_pad = prepare_input(mix_chunk, align_hop)
    aligned_len = mix_stereo.shape[-1]
    original_len = aligned_len - align_pad
    batch = build_windows(mix_stereo, g)
    with torch.no_grad():
        bt = torch.from_numpy(batch).to(dtype)
        spec = stft(bt, g)
        out = net(spec.to(torch.float32)).to(dtype)
        wave_t = istft(out, g)
        wave = wave_t[:, :, g.trim : -g.trim].transpose(0, 1).reshape(2, -1).to(torch.float32).numpy()
    wave = wave[:, :aligned_len]
    mix_for_sub = mix_stereo[:, :aligned_len]
    if align_pad:
        wave = wave[:, :original_len]
        mix_for_sub = mix_for_sub[:, :original_len]
    if output_type == "vocal":
        vocal = wave
        instrumental = mix_for_sub - vocal
    else:
        instrumental = wave
        vocal = mix_for_sub - instrumental
    return vocal.mean(axis=0).astype(np.float32), instrumental.mean(axis=0).astype(np.float32)
