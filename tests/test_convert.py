"""F4: .pth <-> .npz key/layout mapping round trip on the synthetic weights (both variants)."""
import numpy as np
import pytest

from raft_b200 import synth
from raft_b200.convert import npz_to_state_dict, state_dict_to_npz


@pytest.mark.parametrize("small", [False, True])
def test_round_trip(small):
    p = synth.make_weights(small)
    sd = npz_to_state_dict(p)
    assert "update_block.gru.convz1.weight" in sd or small
    assert sd["fnet.conv1.weight"].shape == ((32 if small else 64), 3, 7, 7)  # OIHW
    if not small:
        assert "cnet.layer2.0.downsample.1.running_var" in sd and "update_block.mask.2.bias" in sd
    sd = {"module." + k: v for k, v in sd.items()}
    sd["module.cnet.norm1.num_batches_tracked"] = np.array(0)
    back = state_dict_to_npz(sd)
    assert set(back) == set(p)
    for k in p:
        assert np.array_equal(back[k], p[k]), k
