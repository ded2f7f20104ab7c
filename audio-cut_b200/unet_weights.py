"""TFC-TDF U-Net parameters: naming, random initialisation, BatchNorm folding.

The reference runs an opaque ``Kim_Vocal_1.onnx`` (backends.py:358, config/expert.yaml:22);
parameter names here follow the public KUIELab ``ConvTDFNet`` module that file was
exported from, so a real checkpoint dumped to ``.npz`` loads unchanged.  There is no
network access for checkpoints, hence ``random_state``: seeded He-uniform weights with
a fixed per-layer gain table (found once by ``scripts/calibrate_gains.py``) so that
activations stay O(1) through all 45 layers and the multiplicative skips.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np

BN_EPS = 1e-5


@dataclass(frozen=True)
class UNetGeometry:
    dim_f: int = 3072
    dim_t: int = 256
    dim_c: int = 4
    g: int = 48
    n: int = 5  # L = 2n+1 = 11
    l: int = 3
    bn: int = 8

    def level(self, i: int) -> Tuple[int, int, int]:
        """(channels, T, F) at encoder depth i (0..n; n is the bottleneck)."""
        return self.g * (i + 1), self.dim_t >> i, self.dim_f >> i

    def validate(self) -> None:
        if self.dim_f % (self.bn << self.n) or self.dim_t % (1 << self.n):
            raise ValueError(f"dim_f must be a multiple of {self.bn << self.n}, dim_t of {1 << self.n}")
        if self.g % 16:
            raise ValueError("growth g must be a multiple of 16 (UMMA K granularity)")


def _bn_names(prefix: str) -> List[str]:
    return [prefix + s for s in (".weight", ".bias", ".running_mean", ".running_var")]


def block_prefixes(geo: UNetGeometry) -> List[Tuple[str, int]]:
    """TFC_TDF blocks in execution order with their encoder depth."""
    out = [(f"encoding_blocks.{i}", i) for i in range(geo.n)]
    out.append(("bottleneck_block", geo.n))
    out += [(f"decoding_blocks.{i}", geo.n - 1 - i) for i in range(geo.n)]
    return out


def param_shapes(geo: UNetGeometry) -> "OrderedDict[str, Tuple[int, ...]]":
    geo.validate()
    sh: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def bn(prefix, c):
        for nm in _bn_names(prefix):
            sh[nm] = (c,)

    def tfc_tdf(prefix, c, f):
        for j in range(geo.l):
            sh[f"{prefix}.tfc.H.{j}.0.weight"] = (c, c, 3, 3)
            sh[f"{prefix}.tfc.H.{j}.0.bias"] = (c,)
            bn(f"{prefix}.tfc.H.{j}.1", c)
        sh[f"{prefix}.tdf.0.weight"] = (f // geo.bn, f)
        bn(f"{prefix}.tdf.1", c)
        sh[f"{prefix}.tdf.3.weight"] = (f, f // geo.bn)
        bn(f"{prefix}.tdf.4", c)

    sh["first_conv.0.weight"] = (geo.g, geo.dim_c, 1, 1)
    sh["first_conv.0.bias"] = (geo.g,)
    bn("first_conv.1", geo.g)
    for i in range(geo.n):
        c, _, f = geo.level(i)
        tfc_tdf(f"encoding_blocks.{i}", c, f)
        sh[f"ds.{i}.0.weight"] = (c + geo.g, c, 2, 2)
        sh[f"ds.{i}.0.bias"] = (c + geo.g,)
        bn(f"ds.{i}.1", c + geo.g)
    c, _, f = geo.level(geo.n)
    tfc_tdf("bottleneck_block", c, f)
    for i in range(geo.n):
        c, _, f = geo.level(geo.n - 1 - i)
        sh[f"us.{i}.0.weight"] = (c + geo.g, c, 2, 2)  # ConvTranspose2d: [in, out, kh, kw]
        sh[f"us.{i}.0.bias"] = (c,)
        bn(f"us.{i}.1", c)
        tfc_tdf(f"decoding_blocks.{i}", c, f)
    sh["final_conv.0.weight"] = (geo.dim_c, geo.g, 1, 1)
    sh["final_conv.0.bias"] = (geo.dim_c,)
    return sh


def bn_prefixes(geo: UNetGeometry) -> List[str]:
    """BatchNorm modules in EXECUTION order (the order ``GAINS`` is indexed in)."""
    out = ["first_conv.1"]

    def blk(p):
        return [f"{p}.tfc.H.{j}.1" for j in range(geo.l)] + [f"{p}.tdf.1", f"{p}.tdf.4"]

    for i in range(geo.n):
        out += blk(f"encoding_blocks.{i}") + [f"ds.{i}.1"]
    out += blk("bottleneck_block")
    for i in range(geo.n):
        out += [f"us.{i}.1"] + blk(f"decoding_blocks.{i}")
    return out


# Per-BatchNorm output gains (execution order) and the final-conv gain, produced by
# scripts/calibrate_gains.py on the full Kim_Vocal geometry with the synthetic track.
GAINS: List[float] = [
    0.1924, 0.9259, 0.8884, 0.8443, 0.9496, 0.452, 0.6603, 0.8186, 1.001, 0.9354, 0.8226, 0.4556, 0.5938,
    1.101, 0.8955, 0.7834, 0.8893, 0.4698, 0.7042, 0.919, 0.867, 1.005, 0.8387, 0.4487, 0.6066, 0.8956,
    0.9438, 0.9537, 0.7746, 0.4388, 0.768, 1.004, 0.8853, 0.8672, 1.237, 0.4437, 0.6636, 0.6757, 0.9161,
    0.9981, 0.8432, 0.5164, 0.7034, 0.7043, 0.8396, 0.9197, 0.8796, 0.4482, 0.6953, 0.6418, 0.8894, 0.9421,
    0.8695, 0.4341, 0.7377, 0.447, 0.9549, 0.9362, 0.9043, 0.4462, 0.7208, 0.07598, 0.8999, 1.054, 0.9304,
    0.4812,
]
FINAL_GAIN: float = 1.321


def random_state(geo: UNetGeometry = UNetGeometry(), seed: int = 1234, gains=None, final_gain=None) -> Dict[str, np.ndarray]:
    """Seeded random parameters (float32), names as in ``param_shapes``."""
    rng = np.random.default_rng(seed)
    st: Dict[str, np.ndarray] = {}
    bnset = set(bn_prefixes(geo))
    for name, shape in param_shapes(geo).items():
        if name.endswith("running_mean"):
            v = 0.1 * rng.standard_normal(shape)
        elif name.endswith("running_var"):
            v = 0.5 + rng.random(shape)
        elif name.rsplit(".", 1)[0] in bnset:
            v = 0.8 + 0.4 * rng.random(shape) if name.endswith("weight") else 0.1 + 0.1 * rng.standard_normal(shape)
        elif len(shape) == 1:
            v = 0.05 * rng.standard_normal(shape)
        else:
            if name.startswith("us."):
                fan_in = shape[0]
            elif len(shape) == 4:
                fan_in = shape[1] * shape[2] * shape[3]
            else:
                fan_in = shape[1]
            v = (2 * rng.random(shape) - 1) * np.sqrt(6.0 / fan_in)
        st[name] = np.ascontiguousarray(v, dtype=np.float32)
    gains = GAINS if gains is None else gains
    final_gain = FINAL_GAIN if final_gain is None else final_gain
    if gains:
        for p, gn in zip(bn_prefixes(geo), gains):
            st[p + ".weight"] = (st[p + ".weight"] * gn).astype(np.float32)
            st[p + ".bias"] = (st[p + ".bias"] * gn).astype(np.float32)
    st["final_conv.0.weight"] = (st["final_conv.0.weight"] * final_gain).astype(np.float32)
    st["final_conv.0.bias"] = (st["final_conv.0.bias"] * final_gain).astype(np.float32)
    return st


def fold_bn(st: Dict[str, np.ndarray], bn_prefix: str, conv_bias=None) -> Tuple[np.ndarray, np.ndarray]:
    """Inference BatchNorm as y = x*scale + shift (conv bias folded into shift)."""
    gmm = st[bn_prefix + ".weight"].astype(np.float64)
    beta = st[bn_prefix + ".bias"].astype(np.float64)
    mean = st[bn_prefix + ".running_mean"].astype(np.float64)
    var = st[bn_prefix + ".running_var"].astype(np.float64)
    scale = gmm / np.sqrt(var + BN_EPS)
    b = 0.0 if conv_bias is None else conv_bias.astype(np.float64)
    shift = (b - mean) * scale + beta
    return scale.astype(np.float32), shift.astype(np.float32)


def save_npz(path: str, st: Dict[str, np.ndarray]) -> None:
    np.savez(path, **st)


def load_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def pack_blob(st: Dict[str, np.ndarray], geo: UNetGeometry = UNetGeometry()) -> np.ndarray:
    """Flatten a state dict into the float32 blob ``ac_unet_create`` takes (include/audiocut_b200.h):
    execution order, BatchNorm folded to (scale, shift), conv bias folded into shift."""
    parts: List[np.ndarray] = []

    def put(*arrs):
        for a in arrs:
            parts.append(np.ascontiguousarray(a, dtype=np.float32).reshape(-1))

    def conv_bn(conv: str, bn: str):
        put(st[conv + ".weight"], *fold_bn(st, bn, st.get(conv + ".bias")))

    def block(p: str):
        for j in range(geo.l):
            conv_bn(f"{p}.tfc.H.{j}.0", f"{p}.tfc.H.{j}.1")
        put(st[f"{p}.tdf.0.weight"], *fold_bn(st, f"{p}.tdf.1"))
        put(st[f"{p}.tdf.3.weight"], *fold_bn(st, f"{p}.tdf.4"))

    conv_bn("first_conv.0", "first_conv.1")
    for i in range(geo.n):
        block(f"encoding_blocks.{i}")
        conv_bn(f"ds.{i}.0", f"ds.{i}.1")
    block("bottleneck_block")
    for i in range(geo.n):
        conv_bn(f"us.{i}.0", f"us.{i}.1")
        block(f"decoding_blocks.{i}")
    put(st["final_conv.0.weight"], st["final_conv.0.bias"])
    return np.concatenate(parts)
