"""Drop-in for ``PureVocalPauseDetector._extract_vocal_features`` (pure_vocal_pause_detector.py:410-459).

The legacy (non relative-energy) branch of the pause detector extracts seven framewise series from the
separated vocal stem with librosa + scipy; here every one of them is a CUDA kernel behind the C ABI:

    f0_contour / f0_confidence   librosa.pyin(C2..C7, hop)                  ac_pyin
    formant_energies             LPC(12) peak magnitudes, ragged F1/F2/F3   ac_lpc_formants
    spectral_centroid            librosa.feature.spectral_centroid          ac_stft_features
    harmonic_ratio               low-third-bin magnitude ratio              ac_stft_features
    zero_crossing_rate           librosa.feature.zero_crossing_rate         ac_zero_crossing_rate
    rms_energy                   librosa.feature.rms (frame 2048)           ac_frame_rms

The result is the reference's ``VocalFeatures`` shape: host numpy arrays, same lengths and dtypes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from . import ops


@dataclass
class VocalFeatures:  # mirrors pure_vocal_pause_detector.py:39-48
    f0_contour: np.ndarray
    f0_confidence: np.ndarray
    formant_energies: List[np.ndarray]
    spectral_centroid: np.ndarray
    harmonic_ratio: np.ndarray
    zero_crossing_rate: np.ndarray
    rms_energy: np.ndarray


def extract_vocal_features(audio, sample_rate: int = 44100, hop_length: int = 441, device: int = 0) -> VocalFeatures:
    """``audio``: mono float32 numpy array or a CUDA tensor (then no host->device copy is made)."""
    if isinstance(audio, torch.Tensor):
        x = audio.to(torch.float32)
        if not x.is_cuda:
            x = x.cuda(device)
    else:
        x = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).cuda(device)
    n = x.numel()
    n_frames = 1 + n // hop_length
    f0, _flag, vprob = ops.pyin(x, sample_rate, hop_length)
    mags, counts = ops.lpc_formants(x, sample_rate, hop_length, 12)
    st = ops.stft_features(x, [(0, n, 0)], hop_length, sample_rate, total_frames=n_frames, want=("centroid", "low_ratio"))
    zcr = ops.zero_crossing_rate(x, 2048, hop_length)
    rms = ops.frame_rms(x, 2048, hop_length)
    torch.cuda.synchronize(x.device)
    return VocalFeatures(
        f0_contour=f0.cpu().numpy().astype(np.float64),  # librosa.pyin returns float64 Hz with NaN when unvoiced
        f0_confidence=vprob.cpu().numpy().astype(np.float64),
        formant_energies=ops.formant_tracks(mags.cpu().numpy(), counts.cpu().numpy()),
        spectral_centroid=st["centroid"].cpu().numpy(),
        harmonic_ratio=st["low_ratio"].cpu().numpy(),
        zero_crossing_rate=zcr.cpu().numpy(),
        rms_energy=rms.cpu().numpy(),
    )
