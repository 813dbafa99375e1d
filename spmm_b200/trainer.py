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


def train_step(model, optimizer, prop, text_input_ids, text_attention_mask, alpha, _prepared=False, **fwd_kw):
    """zero_grad -> forward -> backward -> grad all-reduce -> clip(5.) + AdamW.  Returns the 4 losses (device)."""
    from . import ops
    rng = ops.step_rng(model.arena().device)
    rng.advance(bump_host=not _prepared)    # fresh dropout / sampler randomness for this step (device-side increment)
    optimizer.zero_grad()
    losses = model(prop, text_input_ids, text_attention_mask, alpha=alpha, **fwd_kw)
    loss = losses[0] + losses[1] + losses[2] + losses[3]
    loss.backward()
    W = world_size()
    A = model.arena()
    if W > 1:
        dist.all_reduce(A.G[A.adam_start:], op=dist.ReduceOp.SUM)
    if hasattr(optimizer, "prepare_step"):
        optimizer.step(skip_flag=model.last_aux["nan_flag"], grad_scale=1.0 / W, prepared=_prepared)
    else:                                   # a stock torch optimizer (SPMM_models.py:340) also works on the arena views
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.grad is not None], 5.0)
        optimizer.step()
    return losses


class GraphedTrainStep:
    """The whole training step as ONE CUDA graph per (batch shape, alpha): ~2000 kernel launches replayed without the
    Python / launch overhead (the eager step spends ~56 ms of host time enqueuing 61 ms of GPU work).

    Possible because the step has no host sync: negatives are sampled on the device, queue_ptr and the NaN guard live on
    the device, and everything that changes per step lives there too: the dropout / sampler salt and Adam's step
    counter (with its bias corrections) are advanced by kernels inside the graph; only lr is copied from pinned memory.  New batches are copied into static input
    buffers.  `alpha` is baked into a graph; a new value (epoch-0 ramp, SPMM_models.py:355) captures another graph or,
    with `max_graphs` exceeded, falls back to the eager step.
    """

    def __init__(self, model, optimizer, max_graphs=4, warmup_steps=2):
        self.model, self.opt = model, optimizer
        self.graphs = {}
        self.max_graphs, self.warmup_steps = max_graphs, warmup_steps

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

    def _capture(self, key, prop, ids, mask, alpha, mpm_mask):
        from . import ops
        dev = self.model.arena().device
        st = {"prop": prop.to(dev, copy=True), "ids": ids.to(dev, copy=True), "mask": mask.to(dev, copy=True),
              "mpm": None if mpm_mask is None else mpm_mask.to(dev, copy=True)}
        kw = {} if mpm_mask is None else {"mpm_mask": st["mpm"]}
        snap = self._snapshot()                             # warm-up steps are real steps: undo them afterwards
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self.warmup_steps):              # lazy allocations / kernel attributes happen here
                train_step(self.model, self.opt, st["prop"], st["ids"], st["mask"], alpha, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            losses = train_step(self.model, self.opt, st["prop"], st["ids"], st["mask"], alpha, _prepared=True, **kw)
            st["losses"] = torch.stack([l.detach() for l in losses])
        st["graph"] = g
        self.graphs[key] = st
        return st

    def __call__(self, prop, ids, mask, alpha, mpm_mask=None):
        from . import ops
        key = (tuple(prop.shape), tuple(ids.shape), float(alpha), mpm_mask is not None)
        st = self.graphs.get(key)
        if st is None:
            if len(self.graphs) >= self.max_graphs:
                kw = {} if mpm_mask is None else {"mpm_mask": mpm_mask}
                return torch.stack([l.detach() for l in train_step(self.model, self.opt, prop, ids, mask, alpha, **kw)])
            st = self._capture(key, prop, ids, mask, alpha, mpm_mask)
        st["prop"].copy_(prop, non_blocking=True)
        st["ids"].copy_(ids, non_blocking=True)
        st["mask"].copy_(mask, non_blocking=True)
        if mpm_mask is not None:
            st["mpm"].copy_(mpm_mask, non_blocking=True)
        ops.step_rng(self.model.arena().device).host += 1
        self.opt.prepare_step()
        st["graph"].replay()
        return st["losses"]


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
            model.training_step(batch, batch_idx)
        history.append(model.on_train_epoch_end())
    return history
