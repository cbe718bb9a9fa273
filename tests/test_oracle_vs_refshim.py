"""Pin the torch oracle against vectors produced by the reference's OWN source files executed on the
numpy TF/tensorpack shim (oracle/ref_shim, oracle/make_refshim_golden.py).  CPU only."""
import os

import numpy as np
import torch

from oracle import raft_oracle as O
from raft_b200 import synth

G = os.path.join(os.path.dirname(__file__), "golden")


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_ops_match_reference_source():
    z = np.load(os.path.join(G, "refshim_ops.npz"))
    assert torch.equal(O.coords_grid(2, 3, 5), _t(z["coords_grid_2_3_5"]))
    # tf_grid_sample: identical fp32 operation order -> bit exact
    assert torch.equal(O.bilinear_sampler(_t(z["gs_img"]), _t(z["gs_coords"])), _t(z["gs_out"]))
    assert torch.allclose(O.upflow8(_t(z["upflow8_in"])), _t(z["upflow8_out"]), atol=1e-6)
    pyr = O.get_corr_pyramid(_t(z["corr_f1"]), _t(z["corr_f2"]))
    for l in range(4):
        ref = _t(z[f"corr_l{l}"])
        assert pyr[l].shape == ref.shape
        assert torch.allclose(pyr[l], ref, atol=2e-5), (l, (pyr[l] - ref).abs().max())
    # lookup on the REFERENCE's pyramid: bit exact (pure gather + elementwise)
    ref_pyr = [_t(z[f"corr_l{l}"]) for l in range(4)]
    assert torch.equal(O.sample_corr(ref_pyr, _t(z["lookup_coords"]), radius=4), _t(z["lookup_r4"]))
    assert torch.equal(O.sample_corr(ref_pyr, _t(z["lookup_coords"]), radius=3), _t(z["lookup_r3"]))


def test_update_blocks_match_reference_source():
    for name, small in (("refshim_things.npz", False), ("refshim_small.npz", True)):
        z = np.load(os.path.join(G, name))
        p = {k: _t(v) for k, v in synth.make_weights(small, seed=int(z["weight_seed"])).items()}
        fn = O.small_update_block if small else O.basic_update_block
        net, mask, delta = fn(_t(z["net"]), _t(z["inp"]), _t(z["corr"]), _t(z["flow"]), p)
        assert torch.allclose(net, _t(z["net_out"]), atol=2e-5), (net - _t(z["net_out"])).abs().max()
        assert torch.allclose(delta, _t(z["delta"]), atol=2e-5)
        if not small:
            assert torch.allclose(mask, _t(z["mask"]), atol=1e-4)


def test_full_network_graph_matches_reference_source():
    """RAFT.network_graph (RAFT.py:78-109) incl. encoders, 3 iterations, both variants."""
    for name, small in (("refshim_things.npz", False), ("refshim_small.npz", True)):
        z = np.load(os.path.join(G, name))
        p = synth.make_weights(small, seed=int(z["weight_seed"]))
        out = O.RAFTOracle(p, small=small, iters=int(z["iters"])).forward(_t(z["left"]), _t(z["right"]))
        err = (out - _t(z["flow_up"])).abs().max().item()
        assert err < 1e-4, (name, err)
