#!/usr/bin/env python
"""SPMM pre-training step benchmark (BASELINE.json metric: pretrain molecules/sec on 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--seq-len L] [--ragged]

One "step" = the body of the reference's training_step (SPMM_models.py:348-362): zero_grad -> SPMM.forward (EMA,
all encoder passes, 4 loss heads, enqueue) -> backward -> gradient all-reduce -> clip_grad_norm_(5.) -> AdamW, on
the reference's config_bert*.json shapes with B=96 molecules per GPU, queue 36864, train mode (dropout on), bf16
tensor-core GEMMs with fp32 master weights, synthetic BPE-300 ids + N(0,1) property vectors, name-seeded weights.
N>1 is launched by torchrun (one rank per GPU, NCCL); weak scaling (per-GPU batch fixed).

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM (replays of the step's CUDA
graph); `e2e` goes through the reference-facing hook `SPMM.training_step((pv, list_of_SMILES_strings), batch_idx)`:
WordPiece tokenisation, H2D copies of the pinned host batch, graph replay, scheduler cadence and a D2H read of the
four losses are all inside the timed region.  `roofline` is for the dominant kernel (the tcgen05 GEMM, tensor-bound):
every GEMM launch of the step is timed INSIDE a replay of the step's graph with in-kernel %globaltimer stamps (first
CTA past its dependency wait -> last CTA exit), so the figure is the kernel's exposed time in the real step, not an
eager launch bracketed by events.  `cpu_baseline` is the oracle port of the reference path timed on the host cores on a
bounded sample; `gpu_eager_baseline` is the same port on this GPU under torch.autocast(bfloat16) - what the reference's
'16-mixed' eager PyTorch path would run.
`--impl reference` times the CPU path alone (the reference's own algorithm; /root/reference is not on the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
CFG = os.path.join(REPO, "spmm_b200", "configs")
H, I, E, P_TOK, V = 768, 3072, 256, 54, 300
GEMM_DRAM_BYTES_PER_LAUNCH = 105.24e6   # profiles/r2_launches_dram.csv.gz: 104 607 MB over the 994 gemm2 launches of two eager steps (ncu, cold caches)


def flops_per_molecule(l, B, Q):
    """SURVEY.md section 8d (validated against torch FlopCounter on the reference to 5 digits)."""
    S = lambda T, K: T * (24 * H * H + 4 * K * H)
    X = lambda Tq, Tk: 4 * H * H * (Tq + Tk) + 4 * Tq * Tk * H
    F_ = lambda Tq, Tk: S(Tq, Tq) + X(Tq, Tk)
    P = P_TOK
    head = 2 * l * (2 * H * H + 2 * H * V) + P * (2 * H * H + 2 * H) + 4 * 2 * H * E + 3 * 2 * (2 * H) * 2 + 8 * 2 * E * (B + Q)
    fwd_all = 6 * (3 * S(P, P) + 4 * S(l, l) + 4 * F_(P, l) + 5 * F_(l, P)) + head
    fwd_grad = 6 * (2 * S(P, P) + 2 * S(l, l) + 4 * F_(P, l) + 4 * F_(l, P)) + head / 2
    return fwd_all + 2 * fwd_grad


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, mx = [], set(), None
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx = float(r[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_reference_steps(steps, warmup, batch, seq_len, ragged, queue=36864, device="cpu"):
    """Times the reference algorithm (oracle/spmm_ref.py, pinned to the unmodified reference by tests/test_oracle.py):
    train-step body = zero_grad + forward + backward + clip + AdamW, eager PyTorch.  device="cpu": fp32 on the host cores.
    device="cuda": the same code on this GPU under torch.autocast(bfloat16) (the reference's '16-mixed' path; 2B
    multinomial(...).item() host syncs and all), CUDA-event timed.  Returns (molecules/s, ms/step, cores)."""
    import contextlib
    import torch
    from oracle import spmm_ref
    from spmm_b200 import synth
    cores = os.cpu_count() or 1
    on_gpu = device != "cpu"
    if not on_gpu:
        torch.set_num_threads(cores)
    ct = json.load(open(os.path.join(CFG, "config_bert.json")))
    cp = json.load(open(os.path.join(CFG, "config_bert_property.json")))
    keys = torch.load(os.path.join(REPO, "tests", "golden", "full_b8.pt"), weights_only=False)["state_dict_keys"]
    keys = [(k, (s if not k.endswith("queue") else (s[0], queue)), d) for k, s, d in keys]
    P0 = synth.state_from_keys(keys)
    P, moved = {}, {}
    for k, v in P0.items():                                   # aliases (tied decoder) keep sharing one tensor
        if id(v) not in moved:
            moved[id(v)] = v.to(device)
        P[k] = moved[id(v)]
    del P0, moved
    frozen = ("property_encoder_m.", "text_encoder_m.", "property_proj_m.", "text_proj_m.")
    leaves, seen = [], set()
    for k, v in P.items():
        if v.is_floating_point() and not k.startswith(frozen) and not k.endswith("queue") and id(v) not in seen:
            seen.add(id(v))
            v.requires_grad_(True)
            leaves.append(v)
    opt = torch.optim.AdamW(leaves, lr=5e-5, weight_decay=0.02)
    pv, ids, mask, _ = synth.synthetic_batch(batch, seed=1234, fixed_len=None if ragged else seq_len)
    pv, ids, mask = pv.to(device), ids.to(device), mask.to(device)
    times, ptr = [], 0
    for it in range(warmup + steps):
        if on_gpu:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        opt.zero_grad()
        mpm = torch.bernoulli(torch.full_like(pv, 0.5))
        with (torch.autocast("cuda", dtype=torch.bfloat16) if on_gpu else contextlib.nullcontext()):
            losses, aux = spmm_ref.forward(P, ct, cp, pv, ids, mask, 0.4, mpm, queue_ptr=ptr,
                                           sampler=lambda wt, wi: ([int(torch.multinomial(w, 1)) for w in wt],
                                                                   [int(torch.multinomial(w, 1)) for w in wi]))
        ptr = aux["queue_ptr"]
        sum(losses).backward()
        torch.nn.utils.clip_grad_norm_(leaves, 5.0)
        opt.step()
        if on_gpu:
            e1.record()
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) / 1e3
        else:
            dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * statistics.median(times)
    return batch / (ms / 1e3), ms, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 8
    steps, warmup = max(5, min(args.steps, 8)), 2          # bounded sample: ~1.5 s per step on 16 cores
    val, ms, cores = cpu_reference_steps(steps, warmup, B, args.seq_len, args.ragged)
    line = {"impl": "reference", "metric": "pretrain molecules/sec", "value": val, "unit": "molecules/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, B, 1),
            "cpu_baseline": {"value": val, "unit": "molecules/s", "cores": cores, "kind": "port",
                             "sample": "median of %d timed steps (after %d warm-ups) of batch %d, full queue 36864, fp32 "
                                       "fwd+bwd+clip+AdamW on the host cores; oracle/spmm_ref.py = the PORT of the reference's "
                                       "algorithm pinned to the unmodified reference by golden vectors (the reference itself needs "
                                       "pre-import shims and is not on the GPU box)" % (steps, warmup, B)},
            "e2e": {"value": val, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def shutdown_distributed(dist, torch, graphed=None):
    """Leaves the process group without hanging the launcher: CUDA graphs that captured NCCL kernels are released first,
    and if communicator teardown still blocks (seen at N=2: both ranks parked in destroy_process_group after the result
    was printed) the process exits anyway once stdout is flushed."""
    import gc
    import threading as th
    if graphed is not None:
        graphed.graphs.clear()
    gc.collect()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    t = th.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout=20)
    if t.is_alive():
        os._exit(0)


def workload_config(args, B, world):
    return {"workload": "SPMM pretrain step, reference config_bert*.json shapes (12L text/fusion + 6L PV encoder, H768), "
                        "batch %d molecules/GPU, %s, 53-dim PV, momentum queue 36864x256, alpha 0.4, train mode (dropout 0.1)"
                        % (B, "ragged SMILES lengths U{12..99}" if args.ragged else "SMILES length %d" % args.seq_len),
            "global_batch": B * world, "seq_len": None if args.ragged else args.seq_len, "parallelism": "dp%d" % world,
            "l2": "working set per step (0.58 GB bf16 weights + 2.9 GB fp32 arenas + activations) >> 126 MB L2; no explicit flush"}


def synthetic_smiles(tok, lens, seed):
    """SMILES-alphabet strings whose WordPiece encoding has exactly lens[b] model-input tokens ([CLS] pieces [SEP]),
    built with the tokenizer itself (greedy longest-match merges make the piece count depend on the characters)."""
    import random
    rnd = random.Random(seed)
    alphabet = "CCCCccccNnOo()()==12345#SFl"
    out = []
    for L in lens:
        s, n_pieces = "[CLS]", int(L) - 2
        while True:
            cand = s + rnd.choice(alphabet)
            n = int(tok([cand], pin_memory=False).attention_mask.sum()) - 3      # minus HF [CLS], '[CLS]' piece, [SEP]
            if n > n_pieces:
                continue
            s = cand
            if n == n_pieces:
                break
        out.append(s)
    return out


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from spmm_b200 import _lib, kernels, ops, synth, trainer
    from spmm_b200.SPMM_models import SPMM
    from spmm_b200.optim import FusedClipAdamW

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a stuck collective aborts after 3 minutes instead of hanging the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    def log(msg):
        if rank == 0:
            print("[bench %.1fs] %s" % (time.perf_counter() - t_start, msg), file=sys.stderr, flush=True)
    t_start = time.perf_counter()
    B = args.batch
    cfg = synth.pretrain_config(os.path.join(CFG, "config_bert.json"), os.path.join(CFG, "config_bert_property.json"),
                                queue_size=36864, batch_size=B)
    torch.manual_seed(0)
    model = SPMM(config=cfg)
    synth.fill_by_name(model)
    model.to(dev)
    model.build_arenas(dev)
    model.train()
    (opt,), (sched,) = model.configure_optimizers()          # FusedClipAdamW(lr 5e-5, wd 0.02, clip 5.0) + cosine schedule
    ops.manual_seed(999 + rank)
    torch.manual_seed(999 + rank)
    pv_h, ids_h, mask_h, lens = synth.synthetic_batch(B, seed=1234 + rank, fixed_len=None if args.ragged else args.seq_len)
    pv_h, ids_h, mask_h = pv_h.pin_memory(), ids_h.pin_memory(), mask_h.pin_memory()
    pv, ids, mask = pv_h.to(dev), ids_h.to(dev), mask_h.to(dev)
    alpha = 0.4

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graphed = None if args.eager else trainer.GraphedTrainStep(model, opt)
    # the public hook: SPMM.training_step((pv, SMILES strings), batch_idx) - reference SPMM_models.py:348-380
    from spmm_b200.tokenizer import WordPieceTokenizer
    model.tokenizer = WordPieceTokenizer(os.path.join(CFG, "vocab_bpe_300.txt"), do_lower_case=False, do_basic_tokenize=False)
    smiles = synthetic_smiles(model.tokenizer, lens, 4321 + rank)
    model.attach(opt, sched, global_rank=rank, log=None)
    model.current_epoch, model.loader_len = 1, 1000          # past the epoch-0 ramp: alpha = 0.4 like the resident loop
    model.use_cuda_graph = graphed is not None
    if graphed is not None:
        model.__dict__["_stepper"] = graphed                 # training_step replays the same graph as the resident loop

    def step_eager():
        return trainer.train_step(model, opt, pv, ids, mask, alpha)

    def step_resident():
        if graphed is None:
            return step_eager()
        return graphed(pv, ids, mask, alpha)

    e2e_idx = [1]

    def step_e2e():
        # strings -> WordPiece ids (pinned) -> H2D -> step (graph replay) -> scheduler cadence -> losses D2H
        e2e_idx[0] += 1
        return model.training_step((pv_h, smiles), e2e_idx[0]).cpu()

    log("model built; capturing / first step")
    if graphed is not None:
        try:
            step_resident()
        except Exception as ex:                               # noqa: BLE001  (e.g. a collective that cannot be captured)
            if rank == 0:
                print("graph capture failed (%s: %s); falling back to eager launches" % (type(ex).__name__, ex), file=sys.stderr)
            graphed = None
            model.use_cuda_graph = False

    log("first step done (graph=%s); eager counting step" % (graphed is not None))
    _lib.reset_launch_count()
    step_eager()                                   # one eager step counts this step's kernel launches
    launches_per_step = _lib.launch_count()
    for _ in range(args.warmup):
        last = step_resident()
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        last = step_resident()
    host_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps    # CPU time to ENQUEUE one step (no sync inside)
    e1.record()
    sync()
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    losses = [float(x) for x in (last.detach() if torch.is_tensor(last) else torch.stack([l.detach() for l in last]))]
    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_total / args.steps, "gpu_launches": launches_per_step * args.steps,
                              "graph": graphed is not None,
                              "host_enqueue_ms_per_step": host_ms}), flush=True)
        if world > 1:
            shutdown_distributed(dist, torch, graphed)
        return

    # end-to-end through the public API with host batches
    log("resident loop done (%.2f ms/step); e2e loop" % (ms_total / args.steps))
    step_e2e()
    sync()
    e0.record()
    for _ in range(args.steps):
        host_losses = step_e2e()
    e1.record()
    sync()
    ms_e2e = e0.elapsed_time(e1)
    model.training_step_outputs.clear()
    e2e_ids = model.tokenizer(smiles).input_ids[:, 1:]
    assert e2e_ids.shape[1] == ids_h.shape[1], (e2e_ids.shape, ids_h.shape)   # same padded width as the resident batch

    t = torch.tensor([ms_total, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])

    # every rank runs the instrumented step (it contains the step's collectives); rank 0 reports it
    log("timed regions done; instrumented step")
    roof, extra = kernel_rooflines(torch, kernels, model, step_eager, B, world,
                                   None if graphed is None else (trainer, opt, pv, ids, mask, alpha))
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            shutdown_distributed(dist, torch, graphed)
        return

    pk = peaks()
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)
    l_eff = statistics.mean(lens)
    fl = sum(flops_per_molecule(l, B, 36864) for l in lens) * world
    line = {"metric": "pretrain molecules/sec", "value": value, "unit": "molecules/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(args, B, world),
            "e2e": {"value": B * world / (ms_e2e / args.steps / 1e3), "unit": "molecules/s",
                    "h2d_bytes_per_step": (pv_h.numel() * 4 + ids_h.numel() * 8 + mask_h.numel() * 8),
                    "d2h_bytes_per_step": 16,
                    "api": "SPMM.training_step((pv, list of %d SMILES strings), batch_idx): native WordPiece tokeniser, H2D of the "
                           "pinned batch, CUDA-graph replay, scheduler cadence, losses read back - all inside the timed region" % B},
            "gpu_launches": launches if graphed is None else launches_per_step * args.steps,
            "launch_mode": "eager" if graphed is None else "cuda_graph (one captured graph of the whole step, replayed)",
            "host_enqueue_ms_per_step": host_ms, "clocks": clocks, "losses_last_step": losses,
            "step_tflops": fl / (ms_step / 1e3) / 1e12,
            "step_frac_of_bf16_sustained": fl / (ms_step / 1e3) / 1e12 / (pk["bf16_tflops_sustained"] * world),
            "roofline": roof, "roofline_extra": extra, "peaks": pk}
    if world == 1 and not args.no_cpu_baseline:
        # free our arm's memory first: the eager baseline keeps ~40 GB of autograd state at B=96
        if graphed is not None:
            graphed.graphs.clear()
        graphed = None
        model.__dict__.pop("_stepper", None)
        torch.cuda.empty_cache()
        try:
            v, ms_gpu, _ = cpu_reference_steps(3, 2, B, args.seq_len, args.ragged, device=dev)
            line["gpu_eager_baseline"] = {"value": v, "unit": "molecules/s", "ms_per_step": ms_gpu, "kind": "port",
                                          "what": "oracle/spmm_ref.py (the reference's algorithm, eager PyTorch) on this GPU under "
                                                  "torch.autocast(bfloat16): batch %d, same shapes / queue / step body, median of 3 "
                                                  "CUDA-event-timed steps after 2 warm-ups; cuBLAS / ATen kernels, the reference's 2B "
                                                  "multinomial().item() syncs included" % B}
        except Exception as ex:                              # noqa: BLE001
            line["gpu_eager_baseline"] = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
        torch.cuda.empty_cache()
        v, ms_cpu, cores = cpu_reference_steps(5, 1, 8, args.seq_len, args.ragged)
        line["cpu_baseline"] = {"value": v, "unit": "molecules/s", "cores": cores, "kind": "port",
                                "sample": "median of 5 timed steps (after 1 warm-up) of batch 8, full queue 36864, fp32 "
                                          "fwd+bwd+clip+AdamW, oracle/spmm_ref.py (port of the reference's algorithm, pinned by golden "
                                          "vectors) on the host cores (%.1f s/step)" % (ms_cpu / 1e3)}
    print(json.dumps(line), flush=True)
    if world > 1:
        shutdown_distributed(dist, torch, graphed)


def graph_timed_gemms(torch, kernels, graph_args, model):
    """Times every GEMM launch INSIDE a replay of the step's CUDA graph: a second graph of the step is captured with the
    GEMM kernel's %globaltimer phase stamps switched on (each launch gets its own slot of a ring, baked into the graph),
    replayed, and read back.  Per launch: [first CTA past griddepcontrol.wait (its predecessor has finished), last CTA
    exit].  The launches run back to back with warm L2 and programmatic dependent launch exactly as in the timed
    region - what CUDA events around eager launches cannot give.  Returns (ms per step, flops per step, launches, shapes)."""
    from spmm_b200 import _lib
    trainer, opt, pv, ids, mask, alpha = graph_args
    SLOTS = 2048
    ring = torch.zeros(SLOTS * 148 * 16, dtype=torch.int64, device=pv.device)
    shapes = []
    orig = kernels.gemm

    def rec(a, b, M, N, K_, **k):
        mode = "+".join(n for n, on in (("f32acc", k.get("accumulate")), ("f32", k.get("out_f32") and not k.get("accumulate")),
                                        ("bias", k.get("bias") is not None), ("gelu", k.get("gelu")),
                                        ("pre", k.get("pre_act_out") is not None), ("drop", k.get("dropout_p", 0) > 0),
                                        ("res", k.get("residual") is not None), ("dgelu", k.get("dgelu_pre") is not None),
                                        ("amn", k.get("a_mn")), ("bmn", k.get("b_mn"))) if on) or "plain"
        shapes.append((M, N, K_, mode))
        return orig(a, b, M, N, K_, **k)
    g2 = trainer.GraphedTrainStep(model, opt, warmup_steps=0)
    kernels.gemm = rec
    _lib.lib().spmm_gemm_debug_trace_ring(ring.data_ptr(), SLOTS)
    try:
        g2(pv, ids, mask, alpha)                  # capture (records the launch order) + first replay
    finally:
        kernels.gemm = orig
        _lib.lib().spmm_gemm_debug_trace(None)
    n = len(shapes)
    if n == 0 or n > SLOTS:
        return None
    best = None
    for _ in range(3):                            # replays: the stamps of the last one are read
        ring.zero_()
        g2(pv, ids, mask, alpha)
        torch.cuda.synchronize()
        tr = ring.view(SLOTS, 148, 16)[:n].cpu()
        used = tr[:, :, 1] > 0                    # slot 1 = past the dependency wait, slot 8 = exit (gemm.cu trace_mark)
        big = torch.iinfo(torch.int64).max
        start = torch.where(used, tr[:, :, 1], torch.full_like(tr[:, :, 1], big)).min(dim=1).values
        end = tr[:, :, 8].max(dim=1).values
        ok = used.any(dim=1)
        dur_us = ((end - start).double() / 1e3) * ok        # launches on the 1-CTA kernel carry no stamps (tiny problems)
        tot = float(dur_us.sum()) / 1e3
        if best is None or tot < best[0]:
            best = (tot, dur_us.clone(), ok.clone())
    g2.graphs.clear()
    fl = sum(2.0 * M * N * K_ for (M, N, K_, _), o in zip(shapes, best[2].tolist()) if o)
    return best[0], fl, int(best[2].sum()), shapes, best[1]


def kernel_rooflines(torch, kernels, model, step_fn, B, world, graph_args=None):
    """GEMM: in-graph timing (graph_timed_gemms); plus one instrumented eager step with CUDA events (current stream) around
    every GEMM / EMA / ITC launch (EMA and ITC figures, and the eager GEMM figure kept for comparison)."""
    pk = peaks()
    in_graph = None
    if graph_args is not None:
        try:
            in_graph = graph_timed_gemms(torch, kernels, graph_args, model)
        except Exception as ex:                   # noqa: BLE001
            print("in-graph GEMM timing failed (%s: %s); using the event-timed eager step" % (type(ex).__name__, ex), file=sys.stderr)
    rec = {"gemm": [], "ema": [], "itc": []}
    orig = {"gemm": kernels.gemm, "ema": kernels.ema, "itc": kernels.itc}

    def wrap(name, work):
        fn = orig[name]

        def inner(*a, **k):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = fn(*a, **k)
            e.record()
            rec[name].append((s, e, work(*a, **k)))
            return r
        return inner
    shapes = []

    def gemm_work(a, b, M, N, K_, **k):
        mode = "+".join(n for n, on in (("f32acc", k.get("accumulate")), ("f32", k.get("out_f32") and not k.get("accumulate")),
                                        ("bias", k.get("bias") is not None), ("gelu", k.get("gelu")),
                                        ("pre", k.get("pre_act_out") is not None), ("drop", k.get("dropout_p", 0) > 0),
                                        ("res", k.get("residual") is not None), ("dgelu", k.get("dgelu_pre") is not None),
                                        ("amn", k.get("a_mn")), ("bmn", k.get("b_mn"))) if on) or "plain"
        shapes.append((M, N, K_, mode))
        return 2.0 * M * N * K_
    kernels.gemm = wrap("gemm", gemm_work)
    kernels.ema = wrap("ema", lambda p, pm, pb, pmb, mom: 16.0 * p.numel())             # 8 B read + 4 B + 2x2 B written
    kernels.itc = wrap("itc", lambda zp, zt, zpm, ztm, pq, tq, temp, alpha: 2.0 * 2 * pq.numel() * 4)  # fwd + bwd scan of both queues
    from spmm_b200 import _lib
    dump = os.environ.get("SPMM_BENCH_GEMM_TABLE")
    ring = None
    if dump:
        ring = torch.zeros(2048 * 148 * 16, dtype=torch.int64, device="cuda")
        _lib.lib().spmm_gemm_debug_trace_ring(ring.data_ptr(), 2048)
    try:
        # the host needs ~60 ms to enqueue an eager step: park the GPU behind a spin kernel so the launches queue up and
        # the events bracket back-to-back kernels, not host gaps
        torch.cuda._sleep(int(0.12 * 1.9e9))
        step_fn()
        torch.cuda.synchronize()
    finally:
        kernels.gemm, kernels.ema, kernels.itc = orig["gemm"], orig["ema"], orig["itc"]
        if dump:
            _lib.lib().spmm_gemm_debug_trace(None)
    tot = {k: (sum(s.elapsed_time(e) for s, e, _ in v), sum(w for _, _, w in v), len(v)) for k, v in rec.items()}
    if dump:                                   # per-shape table of the step's GEMMs: event time and in-kernel (%globaltimer) time
        tr = ring.view(2048, 148, 16).cpu()
        agg = {}
        for i, ((s_, e_, w_), sh) in enumerate(zip(rec["gemm"], shapes)):
            a_ = agg.setdefault(sh, [0, 0.0, 0.0, 0.0])
            t = tr[i % 2048]
            used = t[:, 0] > 0
            kus = (int(t[used, 8].max()) - int(t[used, 0].min())) / 1e3 if bool(used.any()) else float("nan")
            a_[0] += 1; a_[1] += s_.elapsed_time(e_); a_[2] += w_; a_[3] += kus
        with open(dump, "w") as f:
            for sh, (n_, ms_, w_, kus) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write("M=%6d N=%5d K=%5d %-30s n=%4d  events %8.3f ms avg %6.1f us %6.1f TF/s | in-kernel avg %6.1f us %6.1f TF/s\n"
                        % (sh[0], sh[1], sh[2], sh[3], n_, ms_, 1e3 * ms_ / n_, w_ / ms_ / 1e9, kus / n_, w_ / n_ / max(kus / n_, 1e-9) / 1e6))
    ev_ms, ev_fl, ev_n = tot["gemm"]
    if in_graph is not None:
        g_ms, g_fl, g_n = in_graph[0], in_graph[1], in_graph[2]
        how = ("every 2-CTA GEMM launch timed inside a replay of the step's CUDA graph with in-kernel %globaltimer stamps "
               "(first CTA past its dependency wait -> last CTA exit), summed over the step")
    else:
        g_ms, g_fl, g_n = ev_ms, ev_fl, ev_n
        how = "CUDA events around every GEMM launch of one eager step (GPU parked behind a spin kernel first)"
    ach = g_fl / (g_ms / 1e3) / 1e12
    # traffic: dram__bytes_read.sum + dram__bytes_write.sum summed over the gemm2 launches of one step / launches, from the
    # committed ncu pass over a whole step (profiles/r2_launches_dram.csv; None until that capture exists)
    traffic = GEMM_DRAM_BYTES_PER_LAUNCH
    roof = {"kernel": "gemm2_bf16_kernel (tcgen05.mma cta_group::2 + TMA + bulk-store epilogue; all fwd/dgrad/wgrad GEMMs of one step)", "bound": "tensor",
            "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
            "traffic": traffic, "launches": g_n, "ms_per_step_in_kernel": g_ms, "flops_per_step": g_fl, "how": how,
            "flops_per_launch": g_fl / max(g_n, 1), "avg_launch_us": 1e3 * g_ms / max(g_n, 1),
            "peak_source": pk["source"] + ", sustained figure (kernel timed inside a long step)"}
    if in_graph is not None and os.environ.get("SPMM_BENCH_GEMM_TABLE"):
        agg = {}
        for sh, us, ok in zip(in_graph[3], in_graph[4].tolist(), [True] * len(in_graph[3])):
            a_ = agg.setdefault(sh, [0, 0.0])
            a_[0] += 1; a_[1] += us
        with open(os.environ["SPMM_BENCH_GEMM_TABLE"] + ".graph", "w") as f:
            for sh, (n_, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                if us > 0:
                    f.write("M=%6d N=%5d K=%5d %-30s n=%4d  in-graph %8.3f ms avg %6.1f us %7.1f TF/s\n"
                            % (sh[0], sh[1], sh[2], sh[3], n_, us / 1e3, us / n_, 2.0 * sh[0] * sh[1] * sh[2] * n_ / us / 1e6))
    extra = {}
    # `traffic` = dram__bytes_read.sum + dram__bytes_write.sum per call from the committed ncu --set full captures
    # (profiles/r1_ncu_full_itc_ema_r1.csv for the EMA kernel, profiles/r1_ncu_full_itc_tc.csv for the two ITC scans)
    for name, label, traffic in (("ema", "ema_kernel (EMA + bf16 shadows, 16 B/param)", 2.2427e9),
                                 ("itc", "spmm_itc_fwd_bwd: itc_scan_kernel<1>, <2> (tcgen05 kind::tf32, queue streamed twice; "
                                         "pass 2 is tensor-bound) + normalize / in-batch / combine / finish", 160.6e6)):
        ms, by, n = tot[name]
        if n:
            gbs = by / (ms / 1e3) / 1e9
            extra[name] = {"kernel": label, "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                           "frac": gbs / pk["hbm_gbs"], "traffic": traffic, "ms": ms, "algorithmic_bytes": by}
    return roof, extra


# ------------------------------------------------------------------------------------------- configs #4 / #5 (inference)
def run_decode(args):
    """BASELINE.json configs[3] / configs[4]: `--workload smiles2pv` (d_smiles2pv.py: 64 tokenised SMILES of length 64 ->
    53 property values each) and `--workload pv2smiles` (d_pv2smiles_batched.py: 64 property vectors, k = 2 beams, 101
    generated tokens - random-init weights never emit [SEP], so every molecule runs the reference's full 1 + 100
    expansions).  One "step" = one whole generation for the batch.  `value`: inputs resident in HBM; `e2e`: pinned host
    inputs copied in and results read back every step.  Beside it: the reference-shaped loop over the same kernels
    (full prefix recomputed per token, one molecule at a time for the beam search, like the reference) and the oracle
    port (eager PyTorch, autocast bf16) on this GPU, on a bounded sample."""
    import torch
    import torch.distributed as dist
    from spmm_b200 import _lib, generate, ops, synth
    from spmm_b200.SPMM_models import SPMM
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    tj, pj = os.path.join(CFG, "config_bert.json"), os.path.join(CFG, "config_bert_property.json")
    model = SPMM(config=synth.pretrain_config(tj, pj, queue_size=96, batch_size=8))
    synth.fill_by_name(model)
    model.to(dev)
    model.build_arenas(dev)
    model.eval()
    N = args.batch if args.batch != 96 else 64
    pv_h, ids_h, mask_h, _ = synth.synthetic_batch(N, seed=1234 + rank, fixed_len=args.seq_len)
    pv_h, ids_h, mask_h = pv_h.pin_memory(), ids_h.pin_memory(), mask_h.pin_memory()
    pv, ids, mask = pv_h.to(dev), ids_h.to(dev), mask_h.to(dev)
    k = 2
    if args.workload == "smiles2pv":
        fast = lambda host: generate.smiles2pv_fast(model, ids_h if host else ids, mask_h if host else mask)
        slow = lambda n: generate.smiles2pv(model, ids[:n], mask[:n])
        what = "SMILES->PV (d_smiles2pv.py): %d SMILES of %d tokens -> 53 autoregressive property predictions each" % (N, args.seq_len)
        h2d, d2h = ids_h.numel() * 16, N * 53 * 4
        n_slow = N
    else:
        fast = lambda host: generate.pv2smiles_batched(model, pv_h if host else pv, k=k)
        slow = lambda n: [generate.pv2smiles(model, pv[i:i + 1], k=k) for i in range(n)]
        what = "PV->SMILES (d_pv2smiles_batched.py): %d property vectors, k=%d beam search, 1 + 100 token expansions each" % (N, k)
        h2d, d2h = pv_h.numel() * 4, N * 12 * 104 * 8
        n_slow = 2

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync(); e0.record()
        for _ in range(steps):
            out = fn()
        e1.record(); sync()
        return e0.elapsed_time(e1) / steps, out
    for _ in range(max(args.warmup, 1)):
        fast(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    ms, _ = timed(lambda: fast(False), args.steps)
    launches = _lib.launch_count()                 # kernels launched from Python (encoders, cross K/V projections)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, res = timed(lambda: fast(True), args.steps)
    # kernels inside the replayed graphs: count one eager token step and multiply by the replays
    _lib.reset_launch_count()
    if args.workload == "smiles2pv":
        dec = next(iter(generate._S2P.values()))
        dec.t_dev.fill_(1)                           # a finished generation leaves the prefix length past the last slot
        dec._step(56)
        launches += args.steps * 53 * _lib.launch_count()
    else:
        next(iter(generate._DECODERS.values()))._step()
        launches += args.steps * 101 * _lib.launch_count()
    torch.cuda.synchronize()
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            shutdown_distributed(dist, torch)
        return
    line = {"metric": "%s molecules/sec" % args.workload, "value": N * world / (ms / 1e3), "unit": "molecules/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": what + "; reference config_bert*.json shapes, name-seeded random-init weights", "global_batch": N * world,
                       "seq_len": args.seq_len, "parallelism": "replicas x%d (independent molecules, no collective)" % world,
                       "l2": "weights 0.3 GB bf16 + KV caches >> 126 MB L2 per generation; no explicit flush"},
            "e2e": {"value": N * world / (ms_e2e / 1e3), "unit": "molecules/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "launch_mode": "cuda_graph per token step (captured once, replayed)", "clocks": clocks,
            "peaks": peaks()}
    if not args.no_cpu_baseline and world == 1:
        slow(1)
        ms_slow, _ = timed(lambda: slow(n_slow), 1)
        line["same_kernels_full_prefix_loop"] = {"value": n_slow / (ms_slow / 1e3), "unit": "molecules/s",
                                                 "what": "the reference-shaped loop (spmm_b200.generate.%s) over the same sm_100a kernels: whole prefix "
                                                         "recomputed per token, eager launches, %d molecule(s)" % (
                                                             "smiles2pv" if args.workload == "smiles2pv" else "pv2smiles, one molecule at a time", n_slow)}
        try:
            import json as _json
            from oracle import generate_ref, spmm_ref
            P = spmm_ref.state_from_model(model, device=dev, requires_grad=False)
            ct, cp = _json.load(open(tj)), _json.load(open(pj))
            with torch.autocast("cuda", dtype=torch.bfloat16):
                if args.workload == "smiles2pv":
                    f = lambda: generate_ref.smiles2pv(P, ct, cp, ids, mask)
                    n_ref = N
                else:
                    f = lambda: generate_ref.pv2smiles_beam(P, ct, cp, pv[:1], k=k)
                    n_ref = 1
                f()
                ms_ref, _ = timed(f, 1)
            line["gpu_eager_baseline"] = {"value": n_ref / (ms_ref / 1e3), "unit": "molecules/s", "kind": "port",
                                          "what": "oracle/generate_ref.py (the reference's loop, eager PyTorch) on this GPU under "
                                                  "torch.autocast(bfloat16), %d molecule(s)" % n_ref}
        except Exception as ex:                                  # noqa: BLE001
            line["gpu_eager_baseline"] = {"unavailable": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=96)
    ap.add_argument("--seq-len", type=int, default=64)
    ap.add_argument("--ragged", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="launch kernels from Python instead of replaying the step's CUDA graph")
    ap.add_argument("--profile", action="store_true", help="bare loop for ncu: no e2e / instrumented / CPU legs")
    ap.add_argument("--workload", default="pretrain", choices=["pretrain", "smiles2pv", "pv2smiles"],
                    help="pretrain = the headline metric (BASELINE.json configs[1]); smiles2pv / pv2smiles = configs[3] / [4]")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload != "pretrain":
        run_decode(args)
    else:
        if not args.profile:
            args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
