"""Environment knobs that select another kernel or schedule (read once per process) must stay correct: the conv /
update-block / encoder parity tests are re-run in a subprocess with each knob set.

Default library: the A/B switches that are still part of it.  The measured-slower round-1 variants (halo tiles,
cta_group::2 pairs, weight multicast, the fused per-iteration kernel) live in csrc/experiments/ and in a SEPARATE
library (`python raft-tf_b200/build.py --experiments` -> libraft_b200_exp.so); their tests run only when that library
has been built and RAFT_B200_TEST_EXPERIMENTS=1 is set, so the default GPU test run does not pay for them."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXP_LIB = os.path.join(ROOT, "raft-tf_b200", "lib", "libraft_b200_exp.so")
experiments = pytest.mark.skipif(not (os.environ.get("RAFT_B200_TEST_EXPERIMENTS") and os.path.exists(EXP_LIB)),
                                 reason="experiment library not built / RAFT_B200_TEST_EXPERIMENTS not set")


def _kernel_parity(env):
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_kernels.py"), "-q", "-x",
                        "-k", "(conv2d or update_block or encoder) and tc", "--timeout", "300", "-p", "no:cacheprovider"],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


@pytest.mark.parametrize("knob", ["RAFT_B200_NO_HOIST", "RAFT_B200_NO_STASH", "RAFT_B200_FH2_SIMT", "RAFT_B200_NO_PDL",
                                  "RAFT_B200_STEM_WINDOWS", "RAFT_B200_NO_SPLITK", "RAFT_B200_NO_SPLITK_CLUSTER", "RAFT_B200_CONVF1_SIMT"])
def test_variant_passes_conv_and_update_parity(cuda, knob):
    _kernel_parity(dict(os.environ, **{knob: "1"}))


def test_warp_per_pixel_lookup_kernel_bit_exact(cuda):
    """RAFT_B200_LOOKUP_V5=1 selects the warp-per-pixel lookup kernel (the one the volume-free path instantiates) for the
    materialised volume: same bit-exact results."""
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_kernels.py"), "-q", "-x",
                        "-k", "lookup", "--timeout", "300", "-p", "no:cacheprovider"],
                       env=dict(os.environ, RAFT_B200_LOOKUP_V5="1"), cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]


@experiments
@pytest.mark.parametrize("knob", ["RAFT_B200_HALO", "RAFT_B200_PAIR", "RAFT_B200_CTA2", "RAFT_B200_FUSED"])
def test_experiment_variant_passes_conv_and_update_parity(cuda, knob):
    _kernel_parity(dict(os.environ, RAFT_B200_LIB=EXP_LIB, **{knob: "1"}))


@experiments
def test_fused_update_kernel_full_pipeline(cuda):
    """RAFT_B200_FUSED=1 (all convs of an update step in one persistent kernel with in-kernel grid barriers): the
    end-to-end parity tests and the batched == per-sample property (several tiles per CTA and job) must hold."""
    env = dict(os.environ, RAFT_B200_LIB=EXP_LIB, RAFT_B200_FUSED="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_e2e.py"),
                        os.path.join(ROOT, "tests", "test_gpu_fullsize.py"), "-q", "-x", "-k", "not cli",
                        "--timeout", "600", "-p", "no:cacheprovider"],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
