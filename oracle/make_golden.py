"""Generate tests/golden/*.npz from the CPU oracle (seeded synthetic frames + seeded weights).
Run from the repo root:  python oracle/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "raft-tf_b200"))
from oracle.raft_oracle import RAFTOracle  # noqa: E402
from raft_b200 import synth  # noqa: E402

if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    for name, small in (("things_64x96_it4.npz", False), ("small_64x96_it4.npz", True)):
        p = synth.make_weights(small, seed=7)
        l, r = synth.make_batch(1, 64, 96, seed0=1000)
        flow, low = RAFTOracle(p, small=small, iters=4).forward(torch.from_numpy(l), torch.from_numpy(r), return_lowres=True)
        np.savez_compressed(os.path.join(out, name), left=l.astype(np.float32), right=r.astype(np.float32),
                            flow=flow.numpy().astype(np.float32), lowres=low.numpy().astype(np.float32),
                            iters=4, weight_seed=7)
        print(name, flow.shape, float(flow.abs().max()))
