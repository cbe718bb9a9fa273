"""Host-side orchestration of one RAFT forward pass on one GPU.

encoders (torch/cuDNN, out of scope this round) -> rb_corr_build -> rb_update_set_state ->
rb_raft_iterate (lookup + update block per iteration; optionally replayed from a CUDA graph) ->
rb_upsample_convex / rb_upflow8.  Mirrors RAFT.network_graph (networks/RAFT.py:78-109).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np
import torch

from . import capi
from .encoders import Encoder
from .weights import pack_update_block


class RaftEngine:
    def __init__(self, params: Dict[str, np.ndarray], small: bool = False, iters: int = 20,
                 device: Optional[torch.device] = None, use_graph: bool = True, math_mode: int = capi.RB_MATH_TC):
        if not torch.cuda.is_available():
            raise RuntimeError("raft_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device if device is not None else "cuda:0")
        self.small, self.iters = bool(small), int(iters)
        self.hidden, self.ctx, self.radius, self.fdim = (96, 64, 3, 128) if small else (128, 128, 4, 256)
        self.use_graph = use_graph and not os.environ.get("RAFT_B200_NO_GRAPH")
        self.math_mode = math_mode
        torch.backends.cudnn.allow_tf32 = False  # the reference is fp32 end to end
        torch.backends.cudnn.benchmark = True  # encoders (cuDNN, out of scope): let it pick its best fp32 kernels
        torch.backends.cuda.matmul.allow_tf32 = False
        with torch.cuda.device(self.device):
            self.fnet = Encoder(params, "fnet", small, "instance", self.device)
            self.cnet = Encoder(params, "cnet", small, "none" if small else "batch", self.device)
            self.blob = pack_update_block(params, small, self.device)
        self._shape = None
        self._graph = None

    # ---- buffers -------------------------------------------------------------------------------
    def _ensure(self, B: int, H: int, W: int):
        if self._shape == (B, H, W):
            return
        if H % 8 or W % 8:
            raise ValueError(f"H and W must be multiples of 8 (got {H}x{W}); the reference has the same constraint "
                             "(SURVEY fact 6) -- pad first, see networks.RAFT")
        h, w, s = H // 8, W // 8, int(self.small)
        d = self.device
        lib = capi.lib
        self.h, self.w = h, w
        self.pyr_bytes = capi.size_query(lib.rb_corr_pyramid_bytes, B, h, w)
        self.pyramid = torch.empty(self.pyr_bytes // 4, dtype=torch.float32, device=d)
        self.cws_bytes = capi.size_query(lib.rb_corr_workspace_bytes, B, h, w, self.fdim)
        self.corr_ws = torch.zeros(self.cws_bytes, dtype=torch.uint8, device=d)
        self.ws_bytes = capi.size_query(lib.rb_update_workspace_bytes, s, B, h, w)
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=d)  # zero fill = channel padding
        self.coords1 = torch.empty(B, h, w, 2, dtype=torch.float32, device=d)
        self.mask = None if self.small else torch.empty(B, h, w, 576, dtype=torch.float32, device=d)
        self.flow_up = torch.empty(B, H, W, 2, dtype=torch.float32, device=d)
        self.net_in = torch.empty(B, h, w, self.hidden, dtype=torch.float32, device=d)
        self.inp_in = torch.empty(B, h, w, self.ctx, dtype=torch.float32, device=d)
        self.fmap1 = torch.empty(B, h, w, self.fdim, dtype=torch.float32, device=d)
        self.fmap2 = torch.empty(B, h, w, self.fdim, dtype=torch.float32, device=d)
        self._shape = (B, H, W)
        self._graph = None

    # ---- stages --------------------------------------------------------------------------------
    def encode(self, left: torch.Tensor, right: torch.Tensor):
        """RAFT.py:53-59,79-87: 2x-1, fnet(left), fnet(right), cnet(left) -> split/tanh/relu."""
        B = left.shape[0]
        both = torch.cat([left, right], 0) * 2.0 - 1.0
        fm = self.fnet(both)  # instance norm is per sample, so batching left|right is exact
        self.fmap1.copy_(fm[:B])
        self.fmap2.copy_(fm[B:])
        c = self.cnet(both[:B])
        self.net_in.copy_(torch.tanh(c[..., :self.hidden]))
        self.inp_in.copy_(torch.relu(c[..., self.hidden:]))

    def _hot_path(self):
        """corr build + iterations + upsampling: hand-written kernels only (graph-capturable)."""
        B, H, W = self._shape
        h, w, s, lib, st = self.h, self.w, int(self.small), capi.lib, capi.stream()
        capi.check(lib.rb_set_math_mode(self.math_mode))
        capi.check(lib.rb_corr_build(capi.ptr(self.fmap1), capi.ptr(self.fmap2), capi.ptr(self.pyramid), B, h, w,
                                     self.fdim, capi.ptr(self.corr_ws), self.cws_bytes, st))
        capi.check(lib.rb_update_set_state(s, capi.ptr(self.ws), capi.ptr(self.net_in), capi.ptr(self.inp_in), B, h, w, st))
        capi.check(lib.rb_coords_grid(capi.ptr(self.coords1), B, h, w, st))
        capi.check(lib.rb_raft_iterate(s, capi.ptr(self.blob), capi.ptr(self.ws), capi.ptr(self.pyramid),
                                       capi.ptr(self.coords1), capi.ptr(self.mask), B, h, w, self.iters, st))
        if self.small:
            capi.check(lib.rb_upflow8(capi.ptr(self.coords1), capi.ptr(self.flow_up), B, h, w, 1.0, st))
        else:
            capi.check(lib.rb_upsample_convex(capi.ptr(self.coords1), capi.ptr(self.mask), capi.ptr(self.flow_up),
                                              B, h, w, st))

    def launches_per_forward(self) -> int:
        """Number of raft_b200 kernels one hot-path pass launches (counted by the library)."""
        capi.lib.rb_launch_count_reset()
        with torch.cuda.device(self.device):
            self._hot_path()
            torch.cuda.synchronize()
        return int(capi.lib.rb_launch_count())

    def run_hot_path(self):
        if not self.use_graph:
            self._hot_path()
            return
        if self._graph is None:
            self._hot_path()  # warm-up: function attributes, tensor-map cache
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._hot_path()
            self._graph = g
        self._graph.replay()

    @torch.no_grad()
    def forward(self, left: torch.Tensor, right: torch.Tensor) -> torch.Tensor:
        """left/right: [B,H,W,3] fp32 CUDA tensors in [0,1] (BGR like the reference).  Returns the
        [B,H,W,2] flow (a view of an engine-owned buffer, overwritten by the next call)."""
        assert left.shape == right.shape and left.dim() == 4 and left.shape[-1] == 3
        with torch.cuda.device(self.device):
            self._ensure(left.shape[0], left.shape[1], left.shape[2])
            self.encode(left.float(), right.float())
            self.run_hot_path()
        return self.flow_up

    def lowres_flow(self) -> torch.Tensor:
        g = torch.stack(torch.meshgrid(torch.arange(self.w, device=self.device),
                                       torch.arange(self.h, device=self.device), indexing="xy"), -1).float()
        return self.coords1 - g[None]
