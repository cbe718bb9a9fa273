"""Execute the reference's OWN networks/*.py (from /root/reference, unmodified) on the numpy TF shim
and store golden vectors under tests/golden/refshim_*.npz.  Run in the build container only:
    python oracle/make_refshim_golden.py
"""
import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_shim"))
sys.path.insert(0, REF)  # `networks`, `common` resolve to the REFERENCE packages here

import tensorflow as tf  # the shim  # noqa: E402
from networks import RAFT as REF_RAFT  # noqa: E402
from networks import model_utils as REF_MU  # noqa: E402
from networks import utils as REF_U  # noqa: E402

assert REF_U.__file__.startswith(REF), REF_U.__file__

spec = importlib.util.spec_from_file_location("synth", os.path.join(ROOT, "raft-tf_b200", "raft_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

OUT = os.path.join(ROOT, "tests", "golden")
A = lambda x: np.asarray(x, dtype=np.float32)  # noqa: E731


def main():
    rng = np.random.default_rng(42)
    g = {}
    # ---- utils.py: coords_grid, tf_grid_sample, upflow8 ----
    g["coords_grid_2_3_5"] = A(REF_U.coords_grid(2, 3, 5))
    img = rng.normal(size=(6, 9, 11, 1)).astype(np.float32)
    crd = (rng.random((6, 5, 4, 2)) * np.array([16, 14]) - 3).astype(np.float32)
    crd[0, 0, 0] = [-0.75, -0.25]; crd[0, 0, 1] = [10.0, 8.0]; crd[0, 0, 2] = [3.0, 2.0]; crd[0, 0, 3] = [-1.0, -5.5]
    g["gs_img"], g["gs_coords"] = img, crd
    g["gs_out"] = A(REF_U.bilinear_sampler(tf.T(img), tf.T(crd)))
    fl = rng.normal(size=(1, 3, 4, 2)).astype(np.float32)
    g["upflow8_in"], g["upflow8_out"] = fl, A(REF_U.upflow8(tf.T(fl)))
    # ---- model_utils.py: GetCorrPyramid + SampleCorr (odd dims), r=4 and r=3 ----
    f1 = rng.normal(size=(2, 13, 27, 32)).astype(np.float32)
    f2 = rng.normal(size=(2, 13, 27, 32)).astype(np.float32)
    pyr = REF_MU.GetCorrPyramid(tf.T(f1), tf.T(f2))
    g["corr_f1"], g["corr_f2"] = f1, f2
    for l, p in enumerate(pyr):
        g[f"corr_l{l}"] = A(p)
    base = A(REF_U.coords_grid(2, 13, 27))
    c = base + (rng.random(base.shape) * 16 - 8).astype(np.float32)
    c[0, 0, 0] = [-0.75, -0.25]; c[0, 0, 1] = [26.0, 12.0]; c[0, 0, 2] = [3.0, 2.0]; c[1, 5, 5] = [-40.0, 90.0]
    g["lookup_coords"] = c
    g["lookup_r4"] = A(REF_MU.SampleCorr(pyr, tf.T(c), radius=4))
    g["lookup_r3"] = A(REF_MU.SampleCorr(pyr, tf.T(c), radius=3))
    np.savez_compressed(os.path.join(OUT, "refshim_ops.npz"), **g)
    print("refshim_ops.npz", {k: v.shape for k, v in g.items()})

    # ---- update blocks and the whole network_graph, both variants ----
    for small in (False, True):
        params = synth.make_weights(small, seed=7)
        tf.VARIABLES.clear()
        tf.VARIABLES.update(params)
        hid, ctx, r = (96, 64, 3) if small else (128, 128, 4)
        K = 4 * (2 * r + 1) ** 2
        B, h, w = 1, 6, 10
        net = np.tanh(rng.normal(size=(B, h, w, hid))).astype(np.float32)
        inp = np.maximum(rng.normal(size=(B, h, w, ctx)), 0).astype(np.float32)
        corr = (rng.normal(size=(B, h, w, K)) * 3).astype(np.float32)
        flow = (rng.normal(size=(B, h, w, 2)) * 2).astype(np.float32)
        fn = REF_MU.SmallUpdateBlock if small else REF_MU.BasicUpdateBlock
        n2, m2, d2 = fn(tf.T(net), tf.T(inp), tf.T(corr), tf.T(flow), name="update_block", hidden_dim=hid)
        out = dict(net=net, inp=inp, corr=corr, flow=flow, net_out=A(n2), delta=A(d2), weight_seed=7)
        if m2 is not None:
            out["mask"] = A(m2)
        # full forward through the reference's RAFT.network_graph (RAFT.py:78-109), 3 iterations
        model = REF_RAFT.RAFT((64, 96, 3), SimpleNamespace(small=small))
        model.iters = 3
        l, rr = synth.make_batch(1, 64, 96, seed0=1000)
        li, ri = model.input_preprocess(tf.T(l), tf.T(rr))
        out.update(left=l, right=rr, iters=3, flow_up=A(model.network_graph(li, ri)))
        name = f"refshim_{'small' if small else 'things'}.npz"
        np.savez_compressed(os.path.join(OUT, name), **out)
        print(name, out["flow_up"].shape, float(np.abs(out["flow_up"]).max()))


if __name__ == "__main__":
    main()
