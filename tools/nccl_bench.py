"""Times the data-parallel exchange of one step in isolation (run under torchrun, one rank per GPU): all-reduce of the
577.5 MB fp32 gradient arena, reduce-scatter + all-gather of the same bytes (the sharded-optimiser variant), and the
98 KB all-gather of the momentum features.  CUDA events on the launching stream, max over ranks; rank 0 prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29540 tools/nccl_bench.py
"""
import datetime
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
n = 144374064 // (8 * 64) * (8 * 64)                      # the optimiser range of the gradient arena (fp32 elements)
g = torch.randn(n, device=dev)
p = torch.randn(n, device=dev)
feats = torch.randn(2, 96, 256, device=dev)
gathered = torch.empty(world, 2, 96, 256, device=dev)
sh = n // world


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def rs_ag():
    dist.reduce_scatter_tensor(g[rank * sh:(rank + 1) * sh], g)
    dist.all_gather_into_tensor(p, p[rank * sh:(rank + 1) * sh])


res = {"world": world, "bytes": 4 * n,
       "all_reduce_ms": timed(lambda: dist.all_reduce(g)),
       "reduce_scatter_plus_all_gather_ms": timed(rs_ag),
       "feature_all_gather_ms": timed(lambda: dist.all_gather_into_tensor(gathered, feats))}
res["all_reduce_algbw_GBs"] = 4 * n / res["all_reduce_ms"] / 1e6
res["all_reduce_busbw_GBs"] = res["all_reduce_algbw_GBs"] * 2 * (world - 1) / world
if rank == 0:
    print(json.dumps(res), flush=True)
dist.barrier()
os._exit(0)
