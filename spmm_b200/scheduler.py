"""Host-side LR schedule with the call surface SPMM uses (`create_scheduler(args, optimizer)` -> (sched, epochs),
`sched.step(epoch)`; reference scheduler/scheduler_factory.py:10, cosine_lr.py:69-96).  Scalar math only."""
import math


class CosineLRScheduler:
    def __init__(self, optimizer, t_initial, lr_min=0.0, decay_rate=1.0, warmup_t=0, warmup_lr_init=0.0,
                 cycle_limit=1, warmup_prefix=True):
        self.optimizer = optimizer
        self.t_initial, self.lr_min, self.decay_rate = t_initial, lr_min, decay_rate
        self.warmup_t, self.warmup_lr_init, self.cycle_limit, self.warmup_prefix = warmup_t, warmup_lr_init, cycle_limit, warmup_prefix
        self.base_values = [g["lr"] for g in optimizer.param_groups]
        for g in optimizer.param_groups:
            g.setdefault("initial_lr", g["lr"])
        self.warmup_steps = [(v - warmup_lr_init) / warmup_t for v in self.base_values] if warmup_t else [1] * len(self.base_values)
        if warmup_t:
            self._set([warmup_lr_init] * len(self.base_values))

    def _set(self, values):
        for g, v in zip(self.optimizer.param_groups, values):
            g["lr"] = v

    def _get_lr(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * s for s in self.warmup_steps]
        if self.warmup_prefix:
            t = t - self.warmup_t
        i = t // self.t_initial
        t_curr = t - self.t_initial * i
        gamma = self.decay_rate ** i
        lr_min = self.lr_min * gamma
        if self.cycle_limit == 0 or i < self.cycle_limit:
            return [lr_min + 0.5 * (v * gamma - lr_min) * (1 + math.cos(math.pi * t_curr / self.t_initial))
                    for v in self.base_values]
        return [self.lr_min for _ in self.base_values]

    def step(self, epoch, metric=None):
        self._set(self._get_lr(epoch))

    def get_cycle_length(self):
        return self.t_initial * max(1, self.cycle_limit)

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)


def create_scheduler(args, optimizer):
    if args.sched != "cosine":
        raise NotImplementedError("the SPMM pre-training config uses sched='cosine' (SPMM_pretrain.py:61)")
    s = CosineLRScheduler(optimizer, t_initial=args.epochs, lr_min=args.min_lr, decay_rate=args.decay_rate,
                          warmup_lr_init=args.warmup_lr, warmup_t=args.warmup_epochs, cycle_limit=1)
    return s, s.get_cycle_length() + args.cooldown_epochs
