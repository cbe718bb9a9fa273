"""The opt-in kernel variants (environment knobs read once per process) must stay correct: run the conv / update-block
parity tests in a subprocess with each knob set."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("knob", ["RAFT_B200_HALO", "RAFT_B200_PAIR", "RAFT_B200_CTA2", "RAFT_B200_PDL", "RAFT_B200_NO_HOIST", "RAFT_B200_FUSED", "RAFT_B200_FH2_SIMT", "RAFT_B200_NO_PDL"])
def test_variant_passes_conv_and_update_parity(cuda, knob):
    env = dict(os.environ, **{knob: "1"})
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_kernels.py"), "-q", "-x",
                        "-k", "(conv2d or update_block or encoder) and tc", "--timeout", "300", "-p", "no:cacheprovider"],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


def test_fused_update_kernel_full_pipeline(cuda):
    """RAFT_B200_FUSED=1 (all convs of an update step in one persistent kernel with in-kernel grid barriers): the
    end-to-end parity tests and the batched == per-sample property (several tiles per CTA and job) must hold."""
    env = dict(os.environ, RAFT_B200_FUSED="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_e2e.py"),
                        os.path.join(ROOT, "tests", "test_gpu_fullsize.py"), "-q", "-x", "-k", "not cli",
                        "--timeout", "600", "-p", "no:cacheprovider"],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
