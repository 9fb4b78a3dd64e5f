"""Host-side ingest of index arrays: CLI JSON text / nested lists / ndarrays / tensors -> int64 [rows, cols] on the CPU.

The reference turns its inputs into tensors with ``json.loads`` (pro_b_gan_infer.py:485, :493, :501) followed by
``torch.tensor(list)`` (:135-136, :182, :226): ~70 ms of interpreter time at 32768 triplets in front of a 0.18 ms
pass.  JSON text goes through ``pbg_parse_index_rows`` (C, one scan, no Python objects); lists go
through numpy's C converter.  Like the reference, non-integer ids are an IndexError (it indexes with a float tensor,
:139) and ragged rows a ValueError (``torch.tensor`` refuses them)."""
from __future__ import annotations

import ctypes as C

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
