"""Host-side ingest of index arrays: CLI JSON text / nested lists / ndarrays / tensors -> int64 [rows, cols] on the CPU.

The reference turns its inputs into tensors with ``json.loads`` (pro_b_gan_infer.py:485, :493, :501) followed by
``torch.tensor(list)`` (:135-136, :182, :226): ~70 ms of interpreter time at 32768 triplets in front of a 0.18 ms
pass.  JSON text goes through ``pbg_parse_index_rows`` (C, one scan, no Python objects); lists go
through numpy's C converter.  Error types: in lists / arrays / tensors a non-integer id is an IndexError like the
reference's (it indexes with a float tensor, :139) and ragged rows a ValueError (``torch.tensor`` refuses them); in
JSON TEXT anything that is not a well-formed array of integers -- a float id, a ragged row, a stray token -- is a
ValueError carrying the byte offset (the reference's json.loads would have accepted a float and failed later)."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np
import torch

from . import cabi


def parse_index_rows(text, cols: int, pin: bool = False) -> torch.Tensor:
    """JSON text (str / bytes) -> int64 CPU tensor [rows, cols] (cols == 1: "[a, b, ...]" -> [rows, 1])."""
    data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    lib = cabi.load()
    cap = (data.count(b",") + 1) // cols + 1          # a well-formed text holds commas + 1 integers
    out = torch.empty((cap, cols), dtype=torch.int64, pin_memory=pin)
    n = lib.pbg_parse_index_rows(data, len(data), cols, C.c_void_p(out.data_ptr()), cap)
    if n < 0:
        raise ValueError(cabi.last_error(None))
    if n > cap:                                        # cannot happen for well-formed text; never return unwritten rows
        raise ValueError(f"parse_index_rows: {n} rows for an estimate of {cap}")
    return out[:n]


def index_rows(x, cols: int, pin: bool = False) -> torch.Tensor:
    """Anything the reference's methods accept (plus JSON text, ndarrays and tensors) -> int64 CPU [rows, cols]."""
    if isinstance(x, (str, bytes, bytearray)):
        return parse_index_rows(x, cols, pin)
    if isinstance(x, torch.Tensor):
        t = x.detach().cpu()
        if t.dtype.is_floating_point or t.dtype == torch.bool:
            raise IndexError("tensors used as indices must be long, int, byte or bool tensors")
        t = t.to(torch.int64)
    else:
        try:
            a = np.asarray(x)
        except ValueError as e:                        # ragged rows
            raise ValueError(f"expected rows of {cols} ids: {e}") from None
        if a.size == 0:
            a = np.zeros((0, cols), dtype=np.int64)
        if a.dtype == object:
            raise ValueError(f"expected rows of {cols} ids (ragged or non-numeric input)")
        if a.dtype.kind not in "iu":
            raise IndexError("tensors used as indices must be long, int, byte or bool tensors")
        t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64))
    if cols == 1 and t.dim() == 1:
        t = t.unsqueeze(1)
    if t.dim() != 2 or t.shape[1] != cols:
        raise ValueError(f"expected [rows, {cols}] ids, got shape {tuple(t.shape)}")
    t = t.contiguous()
    return t.pin_memory() if pin and t.numel() else t


# ---------------------------------------------------------------------------------------------- results -> JSON text
def _array_json(a, indent: int, depth: int) -> str:
    """JSON text of a 1-D / 2-D float32 or integer array, byte for byte what json.dumps prints for ``a.tolist()``."""
    if isinstance(a, torch.Tensor):
        a = a.detach().cpu().numpy()
    if a.ndim not in (1, 2):
        raise ValueError(f"result arrays are 1-D or 2-D, got shape {a.shape}")
    lib = cabi.load()
    if a.dtype.kind == "f":
        a = np.ascontiguousarray(a, dtype=np.float32)
        fn = lib.pbg_format_f32_json
    elif a.dtype.kind in "iu":
        a = np.ascontiguousarray(a, dtype=np.int64)
        fn = lib.pbg_format_i64_json
    else:
        raise ValueError(f"cannot format dtype {a.dtype}")
    rows, cols = (a.shape[0], 0) if a.ndim == 1 else a.shape
    if a.ndim == 2 and cols == 0:                       # [[], [], ...]: no C fast path needed
        return _indent_tail(json.dumps(a.tolist(), indent=indent if indent >= 0 else None), indent, depth)
    per = 28 + max(indent, 0) * (depth + 3) + 2
    cap = a.size * per + rows * per + 64
    buf = C.create_string_buffer(cap)
    need = fn(C.c_void_p(a.ctypes.data), rows, cols, indent, depth, buf, cap)
    if need < 0:
        raise ValueError(cabi.last_error(None))
    if need > cap:                                      # the estimate is generous; never truncate silently
        buf = C.create_string_buffer(need)
        need = fn(C.c_void_p(a.ctypes.data), rows, cols, indent, depth, buf, need)
    return buf.raw[:need].decode("ascii")


def _indent_tail(text: str, indent: int, depth: int) -> str:
    if indent < 0 or depth == 0:
        return text
    pad = " " * (indent * depth)
    return text.replace("\n", "\n" + pad)


def dumps_results(results: dict, indent: int = 2) -> str:
    """``json.dumps(results, indent=indent)`` (pro_b_gan_infer.py:505-508) for a result dictionary whose top-level
    values may be arrays / tensors (``FusedInference(..., as_arrays=True)``): those are written by the C formatter
    without ever becoming Python lists; everything else goes through json.dumps.  The text is identical to what
    json.dumps prints for the same dictionary with ``.tolist()`` applied to the arrays."""
    import json as _json
    if not results:
        return "{}"
    pad = " " * max(indent, 0)
    items = []
    for k, v in results.items():
        if isinstance(v, (np.ndarray, torch.Tensor)):
            body = _array_json(v, indent, 1)
        else:
            body = _indent_tail(_json.dumps(v, indent=indent if indent >= 0 else None), indent, 1)
        items.append(_json.dumps(k) + ": " + body)
    if indent < 0:
        return "{" + ", ".join(items) + "}"
    return "{\n" + ",\n".join(pad + it for it in items) + "\n}"
