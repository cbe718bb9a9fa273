"""Drop-in mirror of the hot-path functions of the reference's ``networks/model_utils.py``
(GetCorrPyramid :199, SampleCorr :224, BasicUpdateBlock :172, SmallUpdateBlock :187) on torch CUDA
tensors, backed by the raft_b200 C ABI.  The reference resolves weights through TF variable scopes;
here ``set_variables(params)`` plays the role of the session initialiser (infer_raft.py:77)."""
import numpy as np
import torch

from raft_b200 import capi
from raft_b200.weights import pack_update_block

_VARS = {"params": None, "blobs": {}, "generation": 0}
_ENC = {}  # packed encoders of the current checkpoint, see _encoder()


def set_variables(params):
    """params: dict of reference variable names -> numpy arrays (the .npz content).

    The functional API below (GetCorrPyramid ... SmallEncoder) is module-level like the reference's TF variable
    scopes; loading another checkpoint drops every packed blob / encoder of the previous one."""
    _VARS["params"] = params
    _VARS["blobs"] = {}
    _VARS["generation"] += 1  # cache key of the packed encoders (id() of a dead dict can be reused)
    _ENC.clear()


def _blob(small, device):
    key = (bool(small), str(device))
    if key not in _VARS["blobs"]:
        if _VARS["params"] is None:
            raise RuntimeError("networks.model_utils.set_variables(params) has not been called")
        _VARS["blobs"][key] = pack_update_block(_VARS["params"], small, device)
    return _VARS["blobs"][key]


class CorrPyramid(list):
    """List of 4 level tensors [B*h*w, h_l, w_l, 1] that are views of one contiguous buffer."""
    buffer = None
    shape_bhw = None


def GetCorrPyramid(fmap1, fmap2, num_levels=4):
    """model_utils.py:199-221.  fmap1/fmap2: [B,h,w,C] fp32 CUDA (NHWC)."""
    assert num_levels == 4, "the reference only ever uses 4 levels (model_utils.py:199)"
    B, h, w, Cc = fmap1.shape
    lib = capi.lib
    fmap1, fmap2 = fmap1.contiguous().float(), fmap2.contiguous().float()
    with torch.cuda.device(fmap1.device):
        nbytes = capi.size_query(lib.rb_corr_pyramid_bytes, B, h, w)
        buf = torch.empty(nbytes // 4, dtype=torch.float32, device=fmap1.device)
        wsb = capi.size_query(lib.rb_corr_workspace_bytes, B, h, w, Cc)
        ws = torch.zeros(wsb, dtype=torch.uint8, device=fmap1.device)
        capi.check(lib.rb_corr_build(capi.ptr(fmap1), capi.ptr(fmap2), capi.ptr(buf), B, h, w, Cc, capi.ptr(ws), wsb,
                                     capi.stream()))
    pyr = CorrPyramid()
    off = 0
    for l in range(4):
        hl, wl = h >> l, w >> l
        n = B * h * w * hl * wl
        pyr.append(buf[off:off + n].view(B * h * w, hl, wl, 1))
        off += n
    pyr.buffer, pyr.shape_bhw = buf, (B, h, w)
    return pyr


def _as_buffer(corr_pyramid):
    if isinstance(corr_pyramid, CorrPyramid) and corr_pyramid.buffer is not None:
        return corr_pyramid.buffer
    parts = [t.reshape(-1).float() for t in corr_pyramid]
    parts.append(torch.zeros(64, dtype=torch.float32, device=parts[0].device))  # tail padding (rb_corr_pyramid_bytes)
    return torch.cat(parts).contiguous()


def SampleCorr(corr_pyramid, coords, num_levels=4, radius=4):
    """model_utils.py:224-249.  coords [b,h,w,2] -> [b,h,w,4*(2r+1)^2]."""
    assert num_levels == 4
    b, h, w, _ = coords.shape
    buf = _as_buffer(corr_pyramid)
    coords = coords.contiguous().float()
    K = (2 * radius + 1) ** 2
    out = torch.empty(b, h, w, 4 * K, dtype=torch.float32, device=coords.device)
    with torch.cuda.device(out.device):
        capi.check(capi.lib.rb_corr_lookup(capi.ptr(buf), capi.ptr(coords), capi.ptr(out), b, h, w, radius,
                                           capi.stream()))
    return out


def _update_block(small, net, inp, corr, flow, with_mask):
    B, h, w, _ = net.shape
    s = int(small)
    lib = capi.lib
    dev = net.device
    with torch.cuda.device(dev):
        blob = _blob(small, dev)
        wsb = capi.size_query(lib.rb_update_workspace_bytes, s, B, h, w)
        ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
        st = capi.stream()
        net, inp = net.contiguous().float(), inp.contiguous().float()
        corr, flow = corr.contiguous().float(), flow.contiguous().float()
        capi.check(lib.rb_update_set_state(s, capi.ptr(blob), capi.ptr(ws), capi.ptr(net), capi.ptr(inp), B, h, w, st))
        capi.check(lib.rb_update_set_corr(s, capi.ptr(ws), capi.ptr(corr), B, h, w, st))
        from .utils import coords_grid
        coords1 = (flow + coords_grid(B, h, w, dev)).contiguous()
        delta = torch.empty(B, h, w, 2, dtype=torch.float32, device=dev)
        mask = torch.empty(B, h, w, 576, dtype=torch.float32, device=dev) if with_mask else None
        capi.check(lib.rb_update_step(s, capi.ptr(blob), capi.ptr(ws), capi.ptr(coords1), capi.ptr(delta),
                                      capi.ptr(mask), B, h, w, st))
        net_out = torch.empty_like(net)
        capi.check(lib.rb_update_get_net(s, capi.ptr(ws), capi.ptr(net_out), B, h, w, st))
    return net_out, mask, delta


def BasicUpdateBlock(net, inp, corr, flow, name="update_block", hidden_dim=128):
    """model_utils.py:172-185 -> (net, mask, delta_flow)."""
    assert name == "update_block" and hidden_dim == 128
    return _update_block(False, net, inp, corr, flow, True)


def SmallUpdateBlock(net, inp, corr, flow, name="update_block", hidden_dim=96):
    """model_utils.py:187-194 -> (net, None, delta_flow)."""
    assert name == "update_block" and hidden_dim == 96
    return _update_block(True, net, inp, corr, flow, False)


# ---- encoders (model_utils.py:61-105) -----------------------------------------------------------


def _encoder(name, small, norm_fn, out_dim, device):
    from raft_b200.encoders import CudaEncoder
    key = (name, bool(small), norm_fn, int(out_dim), str(device), _VARS["generation"])
    if key not in _ENC:
        if _VARS["params"] is None:
            raise RuntimeError("networks.model_utils.set_variables(params) has not been called")
        _ENC[key] = CudaEncoder(_VARS["params"], name, small, norm_fn, out_dim, device)
    return _ENC[key]


def BasicEncoder(inputs, name, output_dim=256, norm_fn='instance', dropout=0.0):
    """model_utils.py:61-82.  ``inputs``: [B,H,W,3] in [-1,1] (already through input_preprocess, RAFT.py:53-59)."""
    return _encoder(name, False, norm_fn, output_dim, inputs.device)((inputs + 1.0) * 0.5)


def SmallEncoder(inputs, name, out_dim, norm_fn='batch', dropout=0.0):
    """model_utils.py:84-105."""
    return _encoder(name, True, norm_fn, out_dim, inputs.device)((inputs + 1.0) * 0.5)
