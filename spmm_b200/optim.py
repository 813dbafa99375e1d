"""Fused clip_grad_norm_(5.) + AdamW over the flat arenas (reference SPMM_models.py:338-343,361-362).

One reduction kernel (sum g^2 over the whole gradient arena) and one update kernel; the clip coefficient, the
data-parallel 1/world scale and the NaN-guard skip are resolved on the device, so a step issues no host sync.
Semantics follow torch.optim.AdamW with ONE param group over all parameters (weight decay also on biases,
LayerNorm and temp, like the reference); the never-used PV word embedding is excluded exactly as torch skips
parameters whose .grad is None.
"""
import torch

from . import kernels as K


class FusedClipAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=5e-5, weight_decay=0.02, betas=(0.9, 0.999), eps=1e-8, max_norm=5.0):
        self.A = model.arena()
        params = [p for p in model.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, weight_decay=weight_decay, betas=betas, eps=eps))
        n = self.A.n_total - self.A.adam_start
        self.exp_avg = torch.zeros(n, device=self.A.device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=self.A.device, dtype=torch.float32)
        self.max_norm = max_norm
        self.t = 0
        # The step counter lives on the device (t_dev, advanced by a 1-thread kernel that also derives the bias
        # corrections), so a captured CUDA graph of the step picks up fresh values on every replay and a host that runs
        # ahead of the GPU cannot race the per-step scalars.  lr comes from the host (scheduler): whenever it changes,
        # `prepare_step` issues a stream-ordered fill of `lr_dev` (the value travels as a kernel argument, so a host
        # that runs several steps ahead cannot overwrite it before the GPU has read it - no pinned staging buffer).
        # `t` is the host-side mirror (state_dict, schedulers).
        self.lr_dev = torch.zeros(1, dtype=torch.float32, device=self.A.device)
        self._lr_published = None
        self.t_dev = torch.zeros(1, dtype=torch.int64, device=self.A.device)
        self.hyper_dev = torch.zeros(3, dtype=torch.float32, device=self.A.device)

    def zero_grad(self, set_to_none=False):
        self.A.ensure_grads()
        self.A.zero_grad()

    def prepare_step(self):
        """Host side of a step: advance the host mirror of t and, if the scheduler changed it, publish lr to the device
        (stream-ordered fill).  Called once per step BEFORE the (possibly graph-replayed) device work, outside capture."""
        grp = self.param_groups[0]
        self.t += 1
        lr = float(grp["lr"])
        if lr != self._lr_published:
            self.lr_dev.fill_(lr)
            self._lr_published = lr

    @torch.no_grad()
    def step(self, closure=None, skip_flag=None, grad_scale=1.0, prepared=False):
        A, g0 = self.A, self.A.adam_start
        grp = self.param_groups[0]
        if not prepared:
            self.prepare_step()
        K.adam_tick(self.t_dev, self.lr_dev, self.hyper_dev, grp["betas"][0], grp["betas"][1], skip_flag)
        K.grad_sumsq(A.G[g0:], A.sumsq)
        K.adamw(A.P[g0:], A.G[g0:], self.exp_avg, self.exp_avg_sq, grp["lr"], grp["betas"][0], grp["betas"][1], grp["eps"],
                grp["weight_decay"], max(self.t, 1), sumsq=A.sumsq, max_norm=self.max_norm, grad_scale=grad_scale,
                skip_flag=skip_flag, hyper_dev=self.hyper_dev)

    @torch.no_grad()
    def step_sharded(self, world, rank, skip_flag=None, prepared=False):
        """Data-parallel step with the optimiser work split over the ranks (ZeRO-1 style): reduce-scatter of the gradient
        arena (this rank receives the SUM of its 1/world slice), clip + AdamW on that slice only, all-gather of the
        updated fp32 weights.  Same bytes on the wire as the all-reduce it replaces (reduce-scatter + all-gather), but
        the 0.9 ms of sumsq + AdamW shrink by the world size.  Every rank ends with bit-identical weights (the gather
        distributes one copy); the Adam moments of a slice live on its owner only."""
        import torch.distributed as dist
        A, g0 = self.A, self.A.adam_start
        grp = self.param_groups[0]
        if not prepared:
            self.prepare_step()
        n = A.n_total - g0
        assert n % world == 0
        sh = n // world
        lo = g0 + rank * sh
        G, P = A.G[g0:], A.P[g0:]
        gs = A.G[lo:lo + sh]
        dist.reduce_scatter_tensor(gs, G, op=dist.ReduceOp.SUM)              # in place: output = own slice of the input
        K.grad_sumsq(gs, A.sumsq)
        dist.all_reduce(A.sumsq, op=dist.ReduceOp.SUM)                       # global sum g^2 (same bits on every rank)
        K.adam_tick(self.t_dev, self.lr_dev, self.hyper_dev, grp["betas"][0], grp["betas"][1], skip_flag)
        a, b = rank * sh, (rank + 1) * sh
        K.adamw(A.P[lo:lo + sh], gs, self.exp_avg[a:b], self.exp_avg_sq[a:b], grp["lr"], grp["betas"][0], grp["betas"][1],
                grp["eps"], grp["weight_decay"], max(self.t, 1), sumsq=A.sumsq, max_norm=self.max_norm,
                grad_scale=1.0 / world, skip_flag=skip_flag, hyper_dev=self.hyper_dev)
        dist.all_gather_into_tensor(P, A.P[lo:lo + sh])                       # in place: input = own slice of the output
        self.sharded = (world, rank)

    def grad_norm(self, grad_scale=1.0):
        """Total gradient L2 norm of the last step() (device scalar)."""
        return self.A.sumsq.sqrt() * grad_scale

    def _full_moments(self):
        """With a sharded step each rank owns the moments of its slice: gather them for a checkpoint."""
        sharded = getattr(self, "sharded", None)
        if sharded is None:
            return self.exp_avg, self.exp_avg_sq
        import torch.distributed as dist
        world, rank = sharded
        sh = self.exp_avg.numel() // world
        out = []
        for t in (self.exp_avg, self.exp_avg_sq):
            full = torch.empty_like(t)
            dist.all_gather_into_tensor(full, t[rank * sh:(rank + 1) * sh].contiguous())
            out.append(full)
        return out

    def state_dict(self):
        m1, m2 = self._full_moments()
        return {"t": int(self.t_dev), "exp_avg": m1, "exp_avg_sq": m2,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.t = sd["t"]
        self.t_dev.fill_(sd["t"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
