"""GPU (>= 2 devices): the fused gradient exchange over peer memory.  The N-rank gradient must be
bit-identical to the rank-ordered f32 sum of the per-shard gradients, and the replicated weights
must stay bit-identical on all ranks."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


# mlp_cluster: the opt-in cluster step kernel (per-parameter exchange in its P2); mlp_wide: the wide step (exchange in k_wide_adam)
@pytest.mark.parametrize("net", ["mlp", "lstm", "mlp_cluster", "mlp_wide"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_allreduce(world, net):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(29610 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    env = dict(os.environ, MGPU_NET=net.split("_")[0])
    if net.endswith("_cluster"):
        env["SMB200_CLUSTER"] = "1"
    if net.endswith("_wide"):
        env["SMB200_WIDE"] = "1"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("MGPU_RESULT ")][-1]
    res = json.loads(line[len("MGPU_RESULT "):])
    assert len(res) == world
    for r in res:
        assert r["grad_equal"] and r["grad_maxabs"] > 0
        assert r["weights_identical"] and r["weights_moved"] > 0 and r["finite"]
        assert r["grad_step"] == 41
        assert r["beta0"] == r["beta0_alone"]
    assert all(r["beta"] == res[0]["beta"] and r["scaling_sum"] == res[0]["scaling_sum"] for r in res)
