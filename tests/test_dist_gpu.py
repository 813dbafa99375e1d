"""Multi-GPU numerical parity of the data-parallel step on real GPUs (run with `gpurun --gpus 2`): launches
tests/dist_gpu_worker.py under torchrun on 2 ranks (NCCL) and checks its verdict.  Skipped on a 1-GPU box."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlap,sharded", [("0", "0"), ("1", "0"), ("0", "1")],
                         ids=["one_allreduce", "overlapped_buckets", "sharded_optimizer"])
def test_two_rank_step_reduces_gradients_and_keeps_replicas_bit_identical(tmp_path, overlap, sharded):
    """overlap=1: SPMM_DDP_OVERLAP - per-layer all-reduces issued from backward markers on a side stream
    (trainer.GradOverlap); sharded=1: SPMM_DP_SHARDED - reduce-scatter, AdamW on this rank's slice, all-gather of the
    weights.  Eager and captured; the reduced gradient must be the mean of the single-rank ones in every mode."""
    out = str(tmp_path / "dist2.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(REPO, "tests", "dist_gpu_worker.py"), out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO, env=dict(os.environ, SPMM_DDP_OVERLAP=overlap, SPMM_DP_SHARDED=sharded))
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0
    res = json.load(open(out))
    keep = os.environ.get("SPMM_DIST_TEST_LOG")                  # e.g. profiles/r2_dist2_parity.json
    if keep:
        json.dump(res, open(keep.replace(".json", "_overlap%s_sharded%s.json" % (overlap, sharded)), "w"), indent=1)
    assert res["world"] == 2 and res["nccl"] == "nccl"
    assert res["grad_mean_rel"] < 1e-5, res["grad_mean_rel"]     # float-atomic summation order between two backward runs
    assert res["grad_differs_from_local_rel"] > 1e-2              # the check is not vacuous: ranks hold different batches
    assert res["weights_moved"] and res["queue_rank_major"] and res["finite"]
    assert res["queue_ptr_after_1"] == 16 and res["queue_ptr_final"] == (5 * 16) % (8 * 2 * 6) and res["t_dev"] == 5
    for tag in ("identical_after_3_eager_steps", "identical_after_2_graph_steps"):
        own = ("exp_avg", "exp_avg_sq") if sharded == "1" else ()      # moments of a slice live on its owner only
        assert all(v for k_, v in res[tag].items() if k_ not in own), (tag, res[tag])
