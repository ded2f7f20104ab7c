"""Read the TFC-TDF U-Net parameters straight out of an MDX23 ``.onnx`` file.

The reference hands ``Kim_Vocal_1.onnx`` (config/expert.yaml:22) to onnxruntime
(/root/reference/src/audio_cut/separation/backends.py:137-181 picks the file, :216-253 builds the
session, :358 runs it).  This package has no onnxruntime and no ``onnx`` module: the file is a protobuf
``ModelProto`` whose wire format is stable and tiny to walk, so the initializers are pulled out with the
~100-line reader below and mapped onto the parameter names of ``unet_weights.param_shapes`` by walking
the graph's nodes in their (topological = execution) order:

    Conv 1x1 (4 -> g)                          first_conv        [+ BatchNormalization unless folded]
    per block:  l x Conv 3x3, MatMul, BatchNormalization, MatMul, BatchNormalization
    Conv 2x2 stride 2 / ConvTranspose 2x2      ds.i / us.i       [+ BatchNormalization unless folded]
    Conv 1x1 (g -> 4)                          final_conv

PyTorch's exporter folds an inference BatchNorm into the preceding Conv (the weights then arrive as
``onnx::Conv_123`` with a bias and no BN node follows); a folded layer is represented here by an identity
BatchNorm, so both spellings produce the same network.  The geometry (g, n, l, bn, dim_f) is inferred
from the tensor shapes; ``dim_t`` is not a property of the weights (the reference takes 256 from
``Conv_TDF_net_trim_model``, backends.py:260-265).
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

from .unet_weights import BN_EPS, UNetGeometry, param_shapes


# ---- protobuf wire format ---------------------------------------------------------------------
def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    """(field number, wire type, value) of one message; length-delimited values are memoryviews."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = bytes(buf[pos : pos + 8]), pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos : pos + ln], pos + ln
        elif wt == 5:
            val, pos = bytes(buf[pos : pos + 4]), pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        if pos > end:
            raise ValueError("truncated protobuf message")
        yield num, wt, val


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [int(v)]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(x)
    return out


_ONNX_FLOAT, _ONNX_INT64, _ONNX_FLOAT16, _ONNX_DOUBLE = 1, 7, 10, 11


def _tensor(buf: memoryview) -> Tuple[str, Optional[np.ndarray]]:
    """TensorProto -> (name, float32 ndarray) (None for tensors that are not floating point)."""
    dims: List[int] = []
    dtype, name, raw, floats, doubles = 0, "", None, [], []
    for num, wt, val in _fields(buf):
        if num == 1:
            dims += _packed_varints(val, wt)
        elif num == 2:
            dtype = int(val)
        elif num == 4:  # float_data (packed or not)
            floats.append(bytes(val))
        elif num == 10:  # double_data
            doubles.append(bytes(val))
        elif num == 8:
            name = bytes(val).decode("utf-8")
        elif num == 9:
            raw = bytes(val)
        elif num in (13, 14) and (num == 14 and int(val) == 1):
            raise ValueError(f"initializer {name!r} uses external data; export the model with embedded weights")
    if dtype == _ONNX_FLOAT:
        data = np.frombuffer(raw if raw is not None else b"".join(floats), dtype="<f4")
    elif dtype == _ONNX_FLOAT16:
        data = np.frombuffer(raw, dtype="<f2").astype(np.float32) if raw is not None else None
    elif dtype == _ONNX_DOUBLE:
        data = np.frombuffer(raw if raw is not None else b"".join(doubles), dtype="<f8").astype(np.float32)
    else:
        return name, None
    if data is None:
        return name, None
    shape = tuple(int(d) for d in dims)
    if int(np.prod(shape, dtype=np.int64)) != data.size:
        raise ValueError(f"initializer {name!r}: {data.size} values for shape {shape}")
    return name, np.ascontiguousarray(data.reshape(shape), dtype=np.float32)


def _node(buf: memoryview) -> Dict[str, object]:
    ins, outs, op, name = [], [], "", ""
    for num, wt, val in _fields(buf):
        if num == 1:
            ins.append(bytes(val).decode("utf-8"))
        elif num == 2:
            outs.append(bytes(val).decode("utf-8"))
        elif num == 3:
            name = bytes(val).decode("utf-8")
        elif num == 4:
            op = bytes(val).decode("utf-8")
    return {"op": op, "name": name, "inputs": ins, "outputs": outs}


def read_onnx(path: str) -> Tuple[List[Dict[str, object]], Dict[str, np.ndarray]]:
    """(nodes in file order, {initializer name: float32 array}) of an ONNX model file."""
    with open(path, "rb") as f:
        blob = memoryview(f.read())
    graph = None
    for num, wt, val in _fields(blob):
        if num == 7 and wt == 2:  # ModelProto.graph
            graph = val
    if graph is None:
        raise ValueError(f"{path}: no graph in the ONNX model")
    nodes: List[Dict[str, object]] = []
    inits: Dict[str, np.ndarray] = {}
    for num, wt, val in _fields(graph):
        if num == 1 and wt == 2:
            nodes.append(_node(val))
        elif num == 5 and wt == 2:
            name, arr = _tensor(val)
            if arr is not None:
                inits[name] = arr
    # weights may also sit in Constant nodes' "value" attribute (older exporters): AttributeProto.t = field 5
    return nodes, inits


# ---- graph walk -> state dict -------------------------------------------------------------------
class _Layer:
    def __init__(self, kind: str, w: np.ndarray, b: Optional[np.ndarray]):
        self.kind, self.w, self.b, self.bn = kind, w, b, None


def _collect_layers(nodes, inits) -> List[_Layer]:
    produced: Dict[str, _Layer] = {}  # tensor name -> the layer whose (possibly bias-added) output it is
    layers: List[_Layer] = []
    for nd in nodes:
        op, ins, outs = nd["op"], nd["inputs"], nd["outputs"]
        if op in ("Conv", "ConvTranspose"):
            w = inits.get(ins[1]) if len(ins) > 1 else None
            if w is None:
                raise ValueError(f"{op} node {nd['name']!r}: weight is not an initializer")
            b = inits.get(ins[2]) if len(ins) > 2 else None
            lay = _Layer(op, w, b)
            layers.append(lay)
            produced[outs[0]] = lay
        elif op == "MatMul":
            w = inits.get(ins[1])
            if w is None or w.ndim != 2:
                continue  # an activation x activation product is not a layer
            lay = _Layer("MatMul", w, None)
            layers.append(lay)
            produced[outs[0]] = lay
        elif op == "Add" and len(ins) == 2:  # Linear bias after a MatMul
            for a, bname in ((ins[0], ins[1]), (ins[1], ins[0])):
                if a in produced and bname in inits and produced[a].kind == "MatMul" and produced[a].bn is None:
                    produced[a].b = inits[bname]
                    produced[outs[0]] = produced[a]
        elif op == "BatchNormalization":
            lay = produced.get(ins[0])
            if lay is None:
                raise ValueError(f"BatchNormalization {nd['name']!r} does not follow a Conv / ConvTranspose / MatMul")
            try:
                lay.bn = tuple(inits[n] for n in ins[1:5])  # scale, B, mean, var
            except KeyError as exc:
                raise ValueError(f"BatchNormalization {nd['name']!r}: parameter {exc} is not an initializer") from None
    return layers


def _put_bn(st: Dict[str, np.ndarray], prefix: str, lay: _Layer, c: int) -> None:
    if lay.bn is not None:
        scale, beta, mean, var = lay.bn
    else:  # folded by the exporter: identity BatchNorm (scale / sqrt(var + eps) == 1 exactly in fp64, ~1 ulp in fp32)
        scale, beta, mean = np.ones(c, np.float32), np.zeros(c, np.float32), np.zeros(c, np.float32)
        var = np.full(c, 1.0 - BN_EPS, np.float32)
    for nm, v in zip((".weight", ".bias", ".running_mean", ".running_var"), (scale, beta, mean, var)):
        if v.shape != (c,):
            raise ValueError(f"{prefix}{nm}: expected {c} values, file has shape {v.shape}")
        st[prefix + nm] = np.ascontiguousarray(v, dtype=np.float32)


def load_onnx(path: str, dim_t: int = 256) -> Tuple[Dict[str, np.ndarray], UNetGeometry]:
    """State dict (names of ``unet_weights.param_shapes``) and geometry of an MDX23 TFC-TDF ``.onnx`` file."""
    nodes, inits = read_onnx(path)
    layers = _collect_layers(nodes, inits)
    convs3 = [x for x in layers if x.kind == "Conv" and x.w.shape[2:] == (3, 3)]
    ups = [x for x in layers if x.kind == "ConvTranspose"]
    mms = [x for x in layers if x.kind == "MatMul"]
    if not layers or layers[0].kind != "Conv" or layers[0].w.shape[1:] != (4, 1, 1) or not ups or not mms:
        raise ValueError(f"{path}: not a TFC-TDF U-Net (first layer must be a 4-channel 1x1 Conv, with ConvTranspose and MatMul layers)")
    g = int(layers[0].w.shape[0])
    n = len(ups)
    if len(convs3) % (2 * n + 1) or len(mms) != 2 * (2 * n + 1):
        raise ValueError(f"{path}: {len(convs3)} 3x3 convs / {len(mms)} linear layers do not form {2 * n + 1} TFC-TDF blocks")
    l = len(convs3) // (2 * n + 1)
    dim_f, hidden = int(mms[0].w.shape[0]), int(mms[0].w.shape[1])  # MatMul B operand = Linear.weight^T: [f, f/bn]
    if hidden <= 0 or dim_f % hidden:
        raise ValueError(f"{path}: first TDF layer {mms[0].w.shape} is not an f -> f/bn bottleneck")
    geo = UNetGeometry(dim_f=dim_f, dim_t=int(dim_t), dim_c=4, g=g, n=n, l=l, bn=dim_f // hidden)
    shapes = param_shapes(geo)
    st: Dict[str, np.ndarray] = {}
    it = iter(layers)

    def take(kind: str, what: str) -> _Layer:
        try:
            lay = next(it)
        except StopIteration:
            raise ValueError(f"{path}: the graph ends before {what}") from None
        if lay.kind != kind:
            raise ValueError(f"{path}: expected a {kind} for {what}, found {lay.kind} {lay.w.shape}")
        return lay

    def conv_bn(conv: str, bn: str, kind: str = "Conv") -> None:
        lay = take(kind, conv)
        want = shapes[conv + ".weight"]
        if lay.w.shape != want:
            raise ValueError(f"{conv}.weight: expected {want}, file has {lay.w.shape}")
        c = shapes[conv + ".bias"][0]
        st[conv + ".weight"] = lay.w
        st[conv + ".bias"] = lay.b if lay.b is not None else np.zeros(c, np.float32)
        _put_bn(st, bn, lay, c)

    def block(p: str, c: int) -> None:
        for j in range(geo.l):
            conv_bn(f"{p}.tfc.H.{j}.0", f"{p}.tfc.H.{j}.1")
        for lin, bn in (("tdf.0", "tdf.1"), ("tdf.3", "tdf.4")):
            lay = take("MatMul", f"{p}.{lin}")
            w = np.ascontiguousarray(lay.w.T)
            if w.shape != shapes[f"{p}.{lin}.weight"]:
                raise ValueError(f"{p}.{lin}.weight: expected {shapes[f'{p}.{lin}.weight']}, file has {w.shape}")
            if lay.b is not None and np.any(lay.b):
                raise ValueError(f"{p}.{lin}: the TDF linear layers carry no bias in this architecture")
            if lay.bn is None:
                raise ValueError(f"{p}.{bn}: BatchNormalization after the TDF linear layer is missing")
            st[f"{p}.{lin}.weight"] = w
            _put_bn(st, f"{p}.{bn}", lay, c)

    conv_bn("first_conv.0", "first_conv.1")
    for i in range(geo.n):
        block(f"encoding_blocks.{i}", geo.level(i)[0])
        conv_bn(f"ds.{i}.0", f"ds.{i}.1")
    block("bottleneck_block", geo.level(geo.n)[0])
    for i in range(geo.n):
        conv_bn(f"us.{i}.0", f"us.{i}.1", kind="ConvTranspose")
        block(f"decoding_blocks.{i}", geo.level(geo.n - 1 - i)[0])
    lay = take("Conv", "final_conv.0")
    if lay.w.shape != shapes["final_conv.0.weight"] or lay.bn is not None:
        raise ValueError(f"final_conv.0.weight: expected {shapes['final_conv.0.weight']} without BatchNorm, file has {lay.w.shape}")
    st["final_conv.0.weight"] = lay.w
    st["final_conv.0.bias"] = lay.b if lay.b is not None else np.zeros(4, np.float32)
    if next(it, None) is not None:
        raise ValueError(f"{path}: layers left over after final_conv")
    return st, geo


__all__ = ["read_onnx", "load_onnx"]
