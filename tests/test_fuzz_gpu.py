"""GPU: randomised parity -- random canonical ensembles and run shapes through every Metropolis kernel (and the
automatic choice) against the C oracle, bit for bit (scripts/fuzz_parity.py is the generator; 1500 cases were
clean when this was written)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [11, 12])
def test_random_ensembles_all_kernels_bit_exact(seed):
    env = dict(os.environ)
    env.pop("BCP_KERNEL", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_parity.py"), "150", str(seed)],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 mismatches" in r.stdout


@pytest.mark.parametrize("seed", [21])
def test_random_ensembles_large_grids_bit_exact(seed):
    """The shapes the table-driven kernel takes (grids of at least 32 values, at most four scalar restraints, snapshot
    spacings of at least 32 steps or none), next to shapes it declines and hands to the other kernels."""
    env = dict(os.environ)
    env.pop("BCP_KERNEL", None)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "fuzz_parity.py"), "60", str(seed), "big"],
                       capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 mismatches" in r.stdout
