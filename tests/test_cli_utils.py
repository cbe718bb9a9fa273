"""Host-side CLI helpers (no GPU): flow colour coding and .flo I/O; flags of the drop-in CLI."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "raft-tf_b200", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_flo_round_trip(tmp_path):
    fu = _load("flow_utils")
    f = np.random.default_rng(0).normal(size=(7, 9, 2)).astype(np.float32)
    fu.write_flo(tmp_path / "a.flo", f)
    assert np.array_equal(fu.read_flo(tmp_path / "a.flo"), f)


def test_flow_to_color_basic():
    fu = _load("flow_utils")
    w = fu.make_colorwheel()
    assert w.shape == (55, 3) and w[0, 0] == 255 and w[0, 1] == 0
    f = np.zeros((4, 5, 2), np.float32)
    f[..., 0] = 1.0  # uniform flow to the right -> one colour everywhere, zero flow elsewhere would be white
    img = fu.flow_to_color(f, convert_to_bgr=True)
    assert img.shape == (4, 5, 3) and img.dtype == np.uint8 and (img == img[0, 0]).all()
    assert (fu.flow_to_color(np.zeros((2, 2, 2), np.float32)) == 255).all()


def test_cli_flags_match_reference():
    """infer_raft.py:51-67 flags and defaults."""
    src = open(os.path.join(ROOT, "raft-tf_b200", "infer_raft.py")).read()
    for flag, default in (("--gpu", "'1'"), ("--load", "'release_weight/raft-things.npz'"), ("--out", "'./log'"),
                          ("--im1", "'frame_0010.png'"), ("--im2", "'frame_0011.png'"), ("--batch", "1")):
        assert flag in src and f"default={default}" in src, flag
    assert "'--small', action='store_true'" in src and "'-m', '--mode'" in src and "'-o', '--optimizer'" in src


def test_flow_to_color_matches_reference_function_bit_for_bit():
    """tests/golden/flow_to_color.npz = outputs of the reference's own flow_utils.flow_to_color (flow_utils.py:51-121)
    run in-process by oracle/make_flowcolor_golden.py."""
    fu = _load("flow_utils")
    z = np.load(os.path.join(ROOT, "tests", "golden", "flow_to_color.npz"))
    assert np.array_equal(fu.make_colorwheel(), z["colorwheel"])
    names = sorted({k.split("/")[0] for k in z.files if "/" in k})
    assert len(names) >= 6
    for n in names:
        clip = float(z[n + "/clip"])
        for bgr in (0, 1):
            out = fu.flow_to_color(z[n + "/flow"], clip_flow=None if clip < 0 else clip, convert_to_bgr=bool(bgr))
            assert np.array_equal(out, z[f"{n}/bgr{bgr}/out"]), (n, bgr)
