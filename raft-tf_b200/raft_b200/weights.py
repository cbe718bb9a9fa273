"""Weight loading for the reference's ``.npz`` checkpoints (tensorpack ``get_model_loader``,
infer_raft.py:77): keys are TF variable names, conv kernels HWIO."""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np
import torch

from . import capi


def load_npz(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        out = {}
        for k in z.files:
            key = k[:-2] if k.endswith(":0") else k  # tensorpack may store 'name:0'
            out[key] = np.asarray(z[k], dtype=np.float32)
        return out


FP16_MAX = 65504.0


def check_split_range(params: Dict[str, np.ndarray]) -> None:
    """The tensor-core path carries every weight as an fp16 hi/lo pair (csrc/common.cuh): refuse checkpoints whose
    conv weights do not fit instead of silently saturating them.  (He-initialised or trained RAFT weights are O(1).)"""
    for k, v in params.items():
        if k.endswith("/W") or k.endswith("/b"):
            a = np.asarray(v)
            if not np.isfinite(a).all():
                raise ValueError(f"{k}: non-finite values in the checkpoint")
            m = float(np.abs(a).max()) if a.size else 0.0
            if m > FP16_MAX:
                raise ValueError(f"{k}: |w|max = {m:.3g} exceeds the fp16 range of the split-operand format")


def pack_update_block(params: Dict[str, np.ndarray], small: bool, device) -> torch.Tensor:
    """Pack ``update_block/*`` into the device blob rb_update_step consumes."""
    small_i = int(bool(small))
    check_split_range({k: v for k, v in params.items() if k.startswith("update_block/")})
    n = capi.lib.rb_update_num_convs(small_i)
    Ws, bs, keep = (C.c_void_p * n)(), (C.c_void_p * n)(), []
    for i in range(n):
        name = capi.lib.rb_update_conv_name(small_i, i).decode()
        kh, kw, ci, co = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        capi.check(capi.lib.rb_update_conv_shape(small_i, i, C.byref(kh), C.byref(kw), C.byref(ci), C.byref(co)))
        if name + "/W" not in params or name + "/b" not in params:
            raise KeyError(f"checkpoint is missing {name}/W or {name}/b")
        W = np.ascontiguousarray(params[name + "/W"], dtype=np.float32)
        b = np.ascontiguousarray(params[name + "/b"], dtype=np.float32)
        want = (kh.value, kw.value, ci.value, co.value)
        if tuple(W.shape) != want or tuple(b.shape) != (co.value,):
            raise ValueError(f"{name}: expected W{want} b({co.value},), got W{tuple(W.shape)} b{tuple(b.shape)}")
        keep += [W, b]
        Ws[i] = W.ctypes.data
        bs[i] = b.ctypes.data
    nbytes = capi.size_query(capi.lib.rb_update_weights_bytes, small_i)
    blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
    with torch.cuda.device(blob.device):
        capi.check(capi.lib.rb_update_weights_pack(small_i, Ws, bs, capi.ptr(blob), nbytes, capi.stream()))
    return blob
