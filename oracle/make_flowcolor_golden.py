#!/usr/bin/env python
"""Golden vectors for the output edge of the CLI: runs the UNMODIFIED reference function
/root/reference/flow_utils.py:flow_to_color (lines 51-121; it needs only numpy + cv2) in this container and stores its
input/output pairs in tests/golden/flow_to_color.npz.  tests/test_cli_utils.py holds raft-tf_b200/flow_utils.py to them
bit for bit.  Generated with the numpy of this image (2.3: python-float scalars are 'weak', so the reference's
`rad_max + epsilon` stays float32 for float32 flows).  Test infrastructure; /root/reference is only read here."""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RAFT_REFERENCE", "/root/reference")


def main():
    spec = importlib.util.spec_from_file_location("ref_flow_utils", os.path.join(REF, "flow_utils.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(20260921)
    out = {}
    cases = []
    yy, xx = np.mgrid[0:48, 0:64].astype(np.float32)
    cases.append(("smooth", np.stack([np.sin(xx / 7) * 5 + 1, np.cos(yy / 5) * 3 - 2], -1).astype(np.float32), None))
    cases.append(("noise", rng.normal(0, 4, (37, 53, 2)).astype(np.float32), None))
    cases.append(("clip", rng.normal(0, 8, (16, 24, 2)).astype(np.float32), 6.0))
    cases.append(("zero", np.zeros((5, 7, 2), np.float32), None))
    cases.append(("axes", np.array([[[1, 0], [0, 1], [-1, 0], [0, -1], [3, 3], [-2, 5]]], np.float32), None))
    cases.append(("f64", rng.normal(0, 2, (9, 11, 2)), None))
    for name, flow, clip in cases:
        for bgr in (False, True):
            img = ref.flow_to_color(flow.copy(), clip_flow=clip, convert_to_bgr=bgr)
            out[f"{name}/bgr{int(bgr)}/out"] = img
        out[f"{name}/flow"] = flow
        out[f"{name}/clip"] = np.array(-1.0 if clip is None else clip)
    out["colorwheel"] = ref.make_colorwheel()
    path = os.path.join(ROOT, "tests", "golden", "flow_to_color.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; numpy", np.__version__)


if __name__ == "__main__":
    main()
