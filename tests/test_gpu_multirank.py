"""GPU, world size >= 2 (skipped on a single-GPU box; run with `gpurun --gpus 2`): the multi-GPU decomposition on real
devices over NCCL -- device-to-device tree replication (full copy and per-frame deltas), every rank's pool equal to
rank 0's, the gathered row-band image equal to the single-GPU image; and ONE map built by all ranks from one voxel grid
(Morton-range shards, all-gather of the per-pass split counters) bit-identical with the single-GPU build.  ("nothing asserts the gathered image
equals the 1-GPU image or that replicate_tree round-trips a real SVO".)"""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 8])
def test_replicas_and_band_raycast_over_nccl(world):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mr_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("MR_RESULT ")]
    assert line, p.stdout[-2000:]
    res = json.loads(line[0][len("MR_RESULT "):])
    assert len(res) == world
    for r in res:
        assert r["sizes_equal"] and r["pool_equal_rank0"], r
        assert r["pool_equal_after_local_frame"], r
        assert r["geometry_mismatch_refused"], r
        assert all(0 < b < 4_000_000 for b in r["delta_bytes"]), r  # deltas, not pools
        assert r["sharded_build_equal"], r  # one map built by all ranks == the single-GPU build, on every rank
        assert r["sharded_voxels"][0] > 100_000 and r["sharded_build_nodes"][1] > r["sharded_build_nodes"][0]
    assert res[0]["image_equal"] and res[0]["full_copy_nodes"] > 8
