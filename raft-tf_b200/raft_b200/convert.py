"""F4: convert an upstream PyTorch RAFT checkpoint (princeton-vl/RAFT ``.pth`` state_dict) into the
``.npz`` layout the reference loads (readme.md:28 "converted from the official pytorch *.pth"): TF variable names
= upstream parameter names with '.' -> '/', conv kernels OIHW -> HWIO under ``/W``, biases ``/b``, BatchNorm
``weight/bias/running_mean/running_var`` -> ``gamma/beta/mean/EMA/variance/EMA``.

    python -m raft_b200.convert raft-things.pth raft-things.npz
"""
from __future__ import annotations

import sys
from typing import Dict

import numpy as np

_BN = {"weight": "gamma", "bias": "beta", "running_mean": "mean/EMA", "running_var": "variance/EMA"}


def _tf_scope(pt_scope: str) -> str:
    """'cnet.layer2.0.downsample.1' -> 'cnet/layer2/0/downsample.1': the reference names those two layers literally
    'downsample.0' / 'downsample.1' (model_utils.py:33-34,55-56), every other '.' is a scope separator."""
    return pt_scope.replace(".", "/").replace("downsample/", "downsample.")


def _pt_scope(tf_scope: str) -> str:
    return tf_scope.replace("/", ".")


def state_dict_to_npz(sd: Dict[str, "np.ndarray"]) -> Dict[str, np.ndarray]:
    out: Dict[str, np.ndarray] = {}
    keys = {k[len("module."):] if k.startswith("module.") else k: k for k in sd}
    for k, orig in keys.items():
        v = np.asarray(sd[orig].detach().cpu().numpy() if hasattr(sd[orig], "detach") else sd[orig])
        scope, _, leaf = k.rpartition(".")
        if leaf == "num_batches_tracked":
            continue
        tf_scope = _tf_scope(scope)
        is_norm = any(part.startswith("norm") for part in scope.split(".")[-1:]) or scope.endswith("downsample.1")
        if v.ndim == 4 and leaf == "weight":
            out[tf_scope + "/W"] = np.ascontiguousarray(v.transpose(2, 3, 1, 0)).astype(np.float32)  # OIHW -> HWIO
        elif is_norm and leaf in _BN:
            out[f"{tf_scope}/{_BN[leaf]}"] = v.astype(np.float32)
        elif leaf == "bias":
            out[tf_scope + "/b"] = v.astype(np.float32)
        else:
            raise KeyError(f"don't know how to map parameter '{orig}' of shape {v.shape}")
    return out


def npz_to_state_dict(params: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """Inverse mapping (used by the round-trip test and to export weights for upstream-style tooling)."""
    inv = {v: k for k, v in _BN.items()}
    out = {}
    for k, v in params.items():
        if k.endswith("/W"):
            out[_pt_scope(k[:-2]) + ".weight"] = np.ascontiguousarray(np.asarray(v).transpose(3, 2, 0, 1))
        elif k.endswith("/b"):
            out[_pt_scope(k[:-2]) + ".bias"] = np.asarray(v)
        else:
            for tf_leaf, pt_leaf in inv.items():
                if k.endswith("/" + tf_leaf):
                    out[_pt_scope(k[: -len(tf_leaf) - 1]) + "." + pt_leaf] = np.asarray(v)
                    break
            else:
                raise KeyError(k)
    return out


def main(argv=None):
    import torch
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        print(__doc__)
        return 2
    sd = torch.load(argv[0], map_location="cpu")
    sd = sd.get("state_dict", sd) if isinstance(sd, dict) else sd
    np.savez(argv[1], **state_dict_to_npz(sd))
    print(f"wrote {argv[1]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
