"""The body of the reference's `training_step` (SPMM_models.py:348-380) as a plain function: one process per GPU,
torch.distributed (NCCL over NVLink/NVSwitch) for the two exchange points of the data-parallel step:
  * all_gather of the momentum features for the queue enqueue (inside SPMM.forward),
  * ONE all-reduce over the flat gradient arena (577 MB fp32 at full size) after backward.
The clip + AdamW kernels consume the summed gradients with a 1/world scale, so no separate averaging pass runs.
"""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class GradOverlap:
    """Gradient all-reduce overlapped with backward (what DDPStrategy gives the reference, SPMM_pretrain.py:35-36).

    The gradient arena is one flat buffer; an encoder layer owns a contiguous range of it.  Every layer is used by
    several passes of SPMM.forward, and autograd runs backward in reverse creation order, so a layer's gradients are
    final once the backward of the FIRST forward pass that used it has been enqueued.  That pass carries a marker on
    the layer's input (xbert.py); the marker's backward records an event, and a communication stream that waits on it
    all-reduces the layer's range while the main stream continues with the layers below.  `finish()` joins the
    streams and reduces whatever no marker covered (embeddings, heads - a few MB).  Works eagerly and under CUDA-graph
    capture (fork / join through events).  `check=True` (single rank, tests): instead of communicating, snapshot the
    range at marker time and verify at the end that nothing was added to it afterwards."""

    def __init__(self, arena, check=False):
        self.A, self.check = arena, check
        self.stream = None if check else torch.cuda.Stream(device=arena.device)
        self.done, self.claimed, self.snaps = [], set(), []

    def begin(self):
        self.done, self.claimed, self.snaps = [], set(), []

    def claim(self, rng):
        if rng in self.claimed:
            return False
        self.claimed.add(rng)
        return True

    def callback(self, rng):
        return lambda: self.ready(*rng)

    def ready(self, lo, hi):
        G = self.A.G
        if self.check:
            self.snaps.append((lo, hi, G[lo:hi].clone()))
        else:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.stream.wait_event(ev)
            with torch.cuda.stream(self.stream):
                dist.all_reduce(G[lo:hi], op=dist.ReduceOp.SUM)
        self.done.append((lo, hi))

    def finish(self):
        A = self.A
        if self.check:
            bad = [(lo, hi) for lo, hi, snap in self.snaps if not torch.equal(snap, A.G[lo:hi])]
            assert not bad, "gradient ranges modified after their all-reduce was issued: %s" % bad[:4]
            return len(self.snaps)
        torch.cuda.current_stream().wait_stream(self.stream)
        pos = A.adam_start
        for lo, hi in sorted(self.done) + [(A.n_total, A.n_total)]:      # the complement of the ranges already reduced
            if lo > pos:
                dist.all_reduce(A.G[pos:lo], op=dist.ReduceOp.SUM)
            pos = max(pos, hi)
        return len(self.done)


def _overlap_enabled():
    import os
    return os.environ.get("SPMM_DDP_OVERLAP", "0") == "1"


def _sharded_enabled():
    """Reduce-scatter + per-rank AdamW slice + all-gather instead of all-reduce + replicated AdamW: the default under
    NCCL (measured at N = 8: 42.73 vs 43.05 ms per step); SPMM_DP_SHARDED=0 selects the single all-reduce."""
    import os
    return os.environ.get("SPMM_DP_SHARDED", "1") == "1"


def train_step(model, optimizer, prop, text_input_ids, text_attention_mask, alpha, _prepared=False, **fwd_kw):
    """zero_grad -> forward -> backward -> grad all-reduce -> clip(5.) + AdamW.  Returns the 4 losses (device)."""
    from . import ops, xbert
    rng = ops.step_rng(model.arena().device)
    rng.advance(bump_host=not _prepared)    # fresh dropout / sampler randomness for this step (device-side increment)
    optimizer.zero_grad()
    W = world_size()
    A = model.arena()
    ov = None
    sharded = (W > 1 and _sharded_enabled() and hasattr(optimizer, "step_sharded") and dist.get_backend() == "nccl"
               and (A.n_total - A.adam_start) % W == 0)
    if W > 1 and _overlap_enabled() and not sharded:
        ov = getattr(model, "_grad_overlap", None)
        if ov is None or ov.A is not A:
            ov = GradOverlap(A)
            model.__dict__["_grad_overlap"] = ov
        ov.begin()
    xbert.set_grad_overlap(ov)
    try:
        losses = model(prop, text_input_ids, text_attention_mask, alpha=alpha, **fwd_kw)
        loss = losses[0] + losses[1] + losses[2] + losses[3]
        loss.backward()
    finally:
        xbert.set_grad_overlap(None)
    if sharded:
        optimizer.step_sharded(W, dist.get_rank(), skip_flag=model.last_aux["nan_flag"], prepared=_prepared)
        return losses
    if ov is not None:
        ov.finish()
    elif W > 1:
        dist.all_reduce(A.G[A.adam_start:], op=dist.ReduceOp.SUM)
    if hasattr(optimizer, "prepare_step"):
        optimizer.step(skip_flag=model.last_aux["nan_flag"], grad_scale=1.0 / W, prepared=_prepared)
    else:                                   # a stock torch optimizer (SPMM_models.py:340) also works on the arena views
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None], 5.0)
        optimizer.step()
    return losses


class GraphedTrainStep:
    """The whole training step as ONE CUDA graph per batch SHAPE: ~1500 kernel launches replayed without the Python /
    launch overhead (the eager step spends ~56 ms of host time enqueuing ~45 ms of GPU work).

    Possible because the step has no host sync and everything that changes per step lives on the device: negatives are
    sampled there, queue_ptr and the (world-wide) NaN guard are device scalars, the dropout / sampler salt and Adam's
    step counter are advanced by kernels inside the graph, and `alpha` (ramped every batch of epoch 0,
    SPMM_models.py:355), `lr` and the batch's own padded width are device scalars filled before each replay.

    `padding='longest'` gives every batch its own width L (SPMM_models.py:352): batches are padded further to the next
    multiple of `len_bucket` so that a few graphs cover all widths; the pad columns are masked keys / dead rows, and
    the LM loss ignores them through `valid_len`, so the losses equal those of the [B, L] batch.  Graphs share one
    memory pool (they never run concurrently) and are evicted least-recently-used beyond `max_graphs`.

    Note for callers that also run eager steps on the same model: drop the loss tensors of those steps before the first
    capture - a live autograd graph keeps its AccumulateGrad nodes (for `temp`) bound to the stream they were created
    on, and re-using them inside a capture is a CUDA error (cudaErrorStreamCaptureImplicit).
    """

    def __init__(self, model, optimizer, max_graphs=16, warmup_steps=2, len_bucket=8):
        import collections
        self.model, self.opt = model, optimizer
        self.graphs = collections.OrderedDict()
        self.max_graphs, self.warmup_steps, self.len_bucket = max_graphs, warmup_steps, len_bucket
        dev = model.arena().device
        self.alpha_dev = torch.zeros(1, device=dev, dtype=torch.float32)
        self.valid_len = torch.zeros(1, device=dev, dtype=torch.int32)
        self.losses = torch.zeros(4, device=dev, dtype=torch.float32)     # static output, outside the graphs' pool
        self.pool = None
        self.captures = 0

    def _snapshot(self):
        m, o, A = self.model, self.opt, self.model.arena()
        from . import ops
        return {"P": A.P.clone(), "M": A.M.clone(), "m1": o.exp_avg.clone(), "m2": o.exp_avg_sq.clone(), "t": o.t,
                "t_dev": o.t_dev.clone(),
                "pq": m.prop_queue_km.clone(), "tq": m.text_queue_km.clone(), "ptr": m.queue_ptr.clone(),
                "salt": int(ops.step_rng(A.device).host), "rng": torch.cuda.get_rng_state(A.device)}

    def _restore(self, s):
        m, o, A = self.model, self.opt, self.model.arena()
        from . import ops
        A.P.copy_(s["P"]); A.M.copy_(s["M"]); o.exp_avg.copy_(s["m1"]); o.exp_avg_sq.copy_(s["m2"]); o.t = s["t"]
        o.t_dev.copy_(s["t_dev"])
        m.prop_queue_km.copy_(s["pq"]); m.text_queue_km.copy_(s["tq"]); m.queue_ptr.copy_(s["ptr"])
        ops.step_rng(A.device).reset(s["salt"])
        torch.cuda.set_rng_state(s["rng"], A.device)

    def bucket_len(self, L):
        b = self.len_bucket
        return L if b <= 1 else (L + b - 1) // b * b

    def _fill(self, st, prop, ids, mask, alpha, mpm_mask):
        """Batch -> the graph's static input buffers (async; pinned host tensors are copied without a sync)."""
        L = ids.shape[1]
        st["prop"].copy_(prop, non_blocking=True)
        if L == st["ids"].shape[1]:
            st["ids"].copy_(ids, non_blocking=True)
            st["mask"].copy_(mask, non_blocking=True)
        else:                                   # bucket padding: id 0 ([PAD]) / mask 0 beyond the batch's own width
            st["ids"].zero_(); st["mask"].zero_()
            st["ids"][:, :L].copy_(ids, non_blocking=True)
            st["mask"][:, :L].copy_(mask, non_blocking=True)
        if mpm_mask is not None:
            st["mpm"].copy_(mpm_mask, non_blocking=True)
        self.alpha_dev.fill_(float(alpha))      # kernel argument, stream-ordered: no pinned-memory race
        self.valid_len.fill_(int(L))

    def _capture(self, key, prop, ids, mask, alpha, mpm_mask):
        dev = self.model.arena().device
        B, Lb = key[0], key[1]
        st = {"prop": torch.zeros((B,) + tuple(prop.shape[1:]), device=dev, dtype=torch.float32),
              "ids": torch.zeros((B, Lb), device=dev, dtype=torch.int64),
              "mask": torch.zeros((B, Lb), device=dev, dtype=torch.int64),
              "mpm": None if mpm_mask is None else torch.zeros(tuple(mpm_mask.shape), device=dev, dtype=torch.float32)}
        self._fill(st, prop, ids, mask, alpha, mpm_mask)
        kw = {"valid_len": self.valid_len}
        if mpm_mask is not None:
            kw["mpm_mask"] = st["mpm"]
        if self.captures == 0 and self.warmup_steps > 0:
            # lazy allocations / kernel attributes happen in eager warm-up steps; they are real steps: undo them afterwards
            snap = self._snapshot()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(self.warmup_steps):
                    train_step(self.model, self.opt, st["prop"], st["ids"], st["mask"], self.alpha_dev, **kw)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self._restore(snap)
            del snap
        if self.pool is None:
            self.pool = torch.cuda.graph_pool_handle()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self.pool):
            losses = train_step(self.model, self.opt, st["prop"], st["ids"], st["mask"], self.alpha_dev, _prepared=True, **kw)
            self.losses.copy_(torch.stack([l.detach() for l in losses]))
        st["graph"] = g
        self.captures += 1
        while len(self.graphs) >= self.max_graphs:          # LRU eviction: the pool memory is reused by later captures
            self.graphs.popitem(last=False)
        self.graphs[key] = st
        return st

    def __call__(self, prop, ids, mask, alpha, mpm_mask=None):
        """One training step on the batch; returns the 4 losses (static device tensor, overwritten by the next call)."""
        from . import ops
        # dropout on / off is baked into the captured kernels: train and eval mode get their own graphs
        key = (prop.shape[0], self.bucket_len(ids.shape[1]), mpm_mask is not None, bool(getattr(self.model, "training", True)))
        st = self.graphs.get(key)
        if st is None:
            st = self._capture(key, prop, ids, mask, alpha, mpm_mask)
        else:
            self.graphs.move_to_end(key)
            self._fill(st, prop, ids, mask, alpha, mpm_mask)
        ops.step_rng(self.model.arena().device).host += 1
        self.opt.prepare_step()
        st["graph"].replay()
        return self.losses


def fit(model, loader, max_epochs=1, log=None):
    """What `SPMM_pretrain.py:12-37` asks of `pl.Trainer(...).fit(model, data_loader)`, without Lightning: one process
    per GPU (torchrun), `model` already on its device.  Builds the optimiser / scheduler from `configure_optimizers`,
    then runs the reference's hooks: training_step per batch, on_train_epoch_end per epoch."""
    (optimizer,), (scheduler,) = model.configure_optimizers()
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    model.attach(optimizer, scheduler, global_rank=rank, log=log)
    model.train()
    history = []
    for epoch in range(max_epochs):
        model.current_epoch = epoch
        for batch_idx, batch in enumerate(loader):
            model.training_step(batch, batch_idx)       # one CUDA-graph replay per batch (SPMM._graph_stepper)
        history.append(model.on_train_epoch_end())
    return history
