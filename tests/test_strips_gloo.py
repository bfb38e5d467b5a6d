"""N > 1 path on CPU: world_size 2 and 3 over gloo (see strips_gloo_worker.py)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 3])
def test_strip_decomposition_over_gloo(world, oracle_built):
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29700 + world),
                        os.path.join(ROOT, "tests", "strips_gloo_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "gloo strips ok" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
