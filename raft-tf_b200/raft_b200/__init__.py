"""raft_b200 -- B200-native RAFT recurrent-inference hot path (host side).

Python is the host language here because the reference (gonglixue/RAFT-tf) is Python; all
arithmetic of the hot path runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/raft_b200.h`` (see ``capi.py``).  torch is used for device memory, streams, the
(out-of-scope) encoders and torch.distributed plumbing only.
"""
__all__ = ["capi", "engine", "weights", "encoders", "synth", "shard"]
