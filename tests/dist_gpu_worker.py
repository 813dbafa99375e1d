"""Worker of tests/test_dist_gpu.py - run under torchrun with one rank per GPU (NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/dist_gpu_worker.py <out.json>

Checks the data-parallel step of SPMM_models.py:271-286, 389-399 + DDP's gradient mean (SPMM_pretrain.py:35-36) on real
GPUs: (1) the reduced gradient equals the mean of the ranks' single-rank gradients; (2) after 3 eager steps and 2 more
replayed from the step's CUDA graph (NCCL captured inside), weights, momentum weights, Adam moments, queues and the queue
pointer are BIT-IDENTICAL on every rank; (3) the queue holds the ranks' momentum features in rank-major order, like
torch.cat(all_gather(...)).  Rank 0 writes the verdict as JSON."""
import datetime
import json
import os
import sys

import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main(out_path):
    from spmm_b200 import ops, synth, trainer
    from spmm_b200.SPMM_models import SPMM
    from spmm_b200.optim import FusedClipAdamW
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    cfgd = os.path.join(REPO, "spmm_b200", "configs")
    B = 8
    cfg = synth.pretrain_config(os.path.join(cfgd, "config_bert.json"), os.path.join(cfgd, "config_bert_property.json"),
                                queue_size=B * world * 6, batch_size=B)
    model = SPMM(config=cfg)
    synth.fill_by_name(model)
    model.to(dev)
    model.build_arenas(dev)
    model.eval()                                     # dropout off: the only per-rank difference is the data
    opt = FusedClipAdamW(model, lr=1e-4, weight_decay=0.02)
    A = model.arena()
    pv, ids, mask, _ = synth.synthetic_batch(B, seed=1234 + rank, fixed_len=40)
    g = torch.Generator().manual_seed(55 + rank)
    mpm = (torch.rand(B, 53, generator=g) < 0.5).float().to(dev)
    neg = [((torch.arange(B) + torch.randint(1, B, (B,), generator=g)) % B).tolist() for _ in range(2)]
    pv, ids, mask = pv.to(dev), ids.to(dev), mask.to(dev)
    res = {"world": world, "nccl": dist.get_backend(), "overlap": os.environ.get("SPMM_DDP_OVERLAP", "0"),
           "sharded": os.environ.get("SPMM_DP_SHARDED", "1")}
    stepper = trainer.GraphedTrainStep(model, opt)

    # (1) single-rank gradient of this rank's batch (no reduction), state restored afterwards
    snap = stepper._snapshot()
    opt.zero_grad()
    losses = model(pv, ids, mask, alpha=0.4, mpm_mask=mpm, neg_idx=neg)
    sum(losses).backward()
    g_local = A.G[A.adam_start:].clone()
    stepper._restore(snap)
    # no autograd graph of an eager step may outlive this point: its AccumulateGrad nodes (for `temp`, the anchor) carry
    # the legacy stream they were created on and would be re-used by the capture below ("legacy stream depends on a
    # capturing stream")
    del snap, losses
    # the real step: all-reduce(SUM) inside, 1/W folded into clip + AdamW
    p_before = A.P.clone()
    trainer.train_step(model, opt, pv, ids, mask, 0.4, mpm_mask=mpm, neg_idx=neg)
    g_sum = A.G[A.adam_start:].clone()
    parts = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(parts, g_local)
    mean = torch.stack(parts).sum(0) / world
    if os.environ.get("SPMM_DP_SHARDED", "1") == "1":       # reduce-scatter: the SUM lives in this rank's slice only
        sh = g_sum.numel() // world
        sl = slice(rank * sh, (rank + 1) * sh)
    else:
        sl = slice(0, g_sum.numel())
    res["grad_mean_rel"] = float((g_sum[sl] / world - mean[sl]).norm() / mean[sl].norm())
    res["grad_differs_from_local_rel"] = float((g_sum[sl] / world - g_local[sl]).norm() / mean[sl].norm())   # ranks see different data
    worst = torch.tensor([res["grad_mean_rel"]], device=dev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)            # every rank's slice must pass
    res["grad_mean_rel"] = float(worst)
    res["weights_moved"] = bool(not torch.equal(p_before, A.P))
    # (3) queue rows [r*B, (r+1)*B) = rank r's momentum features of step 1
    f = model.last_aux["feat_prop_m"]
    res["queue_rank_major"] = bool(torch.equal(model.prop_queue_km[rank * B:(rank + 1) * B], f))
    res["queue_ptr_after_1"] = int(model.queue_ptr)

    def identical(t):
        ref = t.detach().clone()
        dist.broadcast(ref, src=0)
        flag = torch.tensor([0.0 if torch.equal(ref, t.detach()) else 1.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        return float(flag) == 0.0

    def check(tag):
        res[tag] = {k: identical(t) for k, t in (("P", A.P), ("M", A.M), ("exp_avg", opt.exp_avg), ("exp_avg_sq", opt.exp_avg_sq),
                                                   ("prop_queue", model.prop_queue_km), ("text_queue", model.text_queue_km),
                                                   ("queue_ptr", model.queue_ptr), ("t_dev", opt.t_dev))}
    # (2) two more eager steps, then two graph replays (NCCL collectives captured inside the graph)
    for _ in range(2):
        trainer.train_step(model, opt, pv, ids, mask, 0.4, mpm_mask=mpm, neg_idx=neg)
    check("identical_after_3_eager_steps")
    hist = []
    for _ in range(2):
        hist.append(stepper(pv, ids, mask, 0.4, mpm_mask=mpm).clone())
    check("identical_after_2_graph_steps")
    res["queue_ptr_final"] = int(model.queue_ptr)
    res["t_dev"] = int(opt.t_dev)
    res["losses_rank%d" % rank] = [float(x) for x in hist[-1]]
    fin = torch.tensor([1.0 if all(bool(torch.isfinite(h).all()) for h in hist) else 0.0], device=dev)
    dist.all_reduce(fin, op=dist.ReduceOp.MIN)
    res["finite"] = bool(float(fin) == 1.0)
    torch.cuda.synchronize()
    if rank == 0:
        with open(out_path, "w") as fo:
            json.dump(res, fo, indent=1)
        print(json.dumps(res), flush=True)
    stepper.graphs.clear()
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)                                      # communicator teardown with captured NCCL graphs can block


if __name__ == "__main__":
    main(sys.argv[1])
