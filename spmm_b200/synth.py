"""Synthetic inputs and name-seeded weights.

Everything here is a pure function of (name, shape, seed), so the reference model in the
build container and this framework's model on the GPU box can be filled with identical
numbers without shipping a 1 GB checkpoint: tests/golden/*.pt only hold inputs and outputs.
"""
import zlib

import torch


def _gen(name, salt=0):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) + 7919 * salt) & 0x7FFFFFFF)
    return g


def named_tensor(name, shape, kind, salt=0):
    """Deterministic fp32 tensor for a state-dict entry.

    kind: 'weight' N(0, 0.02) (reference init scale, xbert.py:747), 'ln_weight' 1+N(0,0.05),
          'bias' N(0, 0.02), 'queue' unit-norm columns (SPMM_models.py:72-77).
    """
    g = _gen(name, salt)
    x = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    if kind == "weight":
        return x * 0.02
    if kind == "ln_weight":
        return 1.0 + 0.05 * x
    if kind == "bias":
        return x * 0.02
    if kind == "queue":
        return torch.nn.functional.normalize(x, dim=0)
    raise ValueError(kind)


def _kind_of(name):
    if name.endswith("queue"):
        return "queue"
    if "LayerNorm.weight" in name or name.endswith("property_mtr_head.2.weight"):
        return "ln_weight"
    if name.endswith(".bias") or name.endswith("LayerNorm.bias"):
        return "bias"
    return "weight"


def _canonical(name):
    """Resolves the two aliases in the reference state dict (xbert.py:686-691 + weight tying)."""
    if name.endswith("cls.predictions.decoder.weight"):
        return name.replace("cls.predictions.decoder.weight", "bert.embeddings.word_embeddings.weight")
    if name.endswith("cls.predictions.decoder.bias"):
        return name.replace("cls.predictions.decoder.bias", "cls.predictions.bias")
    return name


def _online_name(name):
    for m in ("property_encoder_m.", "text_encoder_m.", "property_proj_m.", "text_proj_m."):
        if name.startswith(m):
            return name.replace("_m.", ".", 1)
    return name


def value_for(name, shape, momentum_noise=1e-3, salt=0):
    """The name-seeded value of state-dict entry `name` (None for integer buffers)."""
    if name.endswith("position_ids") or name == "queue_ptr":
        return None
    if name == "temp":
        return torch.full(tuple(shape), 0.07)
    name = _canonical(name)
    online = _online_name(name)
    if online in ("property_cls", "property_mask"):
        return named_tensor(online, shape, "weight", salt)
    v = named_tensor(online, shape, _kind_of(online), salt)
    if online != name:
        v = v + momentum_noise * torch.randn(tuple(shape), generator=_gen(name, salt))
    return v


@torch.no_grad()
def fill_by_name(model, momentum_noise=1e-3, salt=0):
    """Overwrites every parameter / queue of `model` (reference SPMM or ours; same key names)
    with name-seeded values through load_state_dict.  Momentum twins (`*_m.`) get online value +
    small name-seeded noise so the teacher path is distinguishable from the student path."""
    vals = {}
    for name, t in model.state_dict().items():
        v = value_for(name, t.shape, momentum_noise, salt)
        if v is None:
            v = torch.zeros_like(t) if name == "queue_ptr" else t
        vals[name] = v.to(t.dtype)
    model.load_state_dict(vals, strict=True)
    return model


def state_from_keys(keys, momentum_noise=1e-3, salt=0):
    """Builds a {name: tensor} state straight from a recorded [(name, shape, dtype)] list
    (tests/golden/*.pt 'state_dict_keys'); aliases share one tensor object."""
    P = {}
    for name, shape, dtype in keys:
        c = _canonical(name)
        if c != name and c in P:
            P[name] = P[c]
            continue
        v = value_for(name, shape, momentum_noise, salt)
        if v is None:
            v = torch.zeros(tuple(shape), dtype=torch.long)
            if name.endswith("position_ids"):
                v = torch.arange(shape[-1]).expand(tuple(shape)).clone()
        P[name] = v
    return P


def synthetic_batch(batch, seed=1234, min_len=12, max_len=100, fixed_len=None, vocab=300, n_prop=53):
    """BPE-300 shaped ids (col0=[CLS]=2, last real=[SEP]=3, pad 0) + PV ~ N(0,1).

    Returns (pv[B,53] f32, ids[B,L] i64, mask[B,L] i64, lens list)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    if fixed_len is None:
        lens = torch.randint(min_len, max_len, (batch,), generator=g)
    else:
        lens = torch.full((batch,), int(fixed_len), dtype=torch.long)
    L = int(lens.max())
    ids = torch.randint(4, vocab, (batch, L), generator=g)
    ids[:, 0] = 2
    ar = torch.arange(L)[None, :]
    ids[ar == (lens[:, None] - 1)] = 3
    mask = (ar < lens[:, None]).long()
    ids = ids * mask
    pv = torch.randn(batch, n_prop, generator=g)
    return pv, ids, mask, lens.tolist()


def pretrain_config(text_json, prop_json, queue_size=36864, batch_size=96):
    """The dict SPMM_pretrain.py:51-65 passes to SPMM(...)."""
    return {
        'property_width': 768, 'embed_dim': 256, 'batch_size': batch_size, 'temp': 0.07,
        'mlm_probability': 0.15, 'queue_size': queue_size, 'momentum': 0.995, 'alpha': 0.4,
        'bert_config_text': text_json, 'bert_config_property': prop_json,
        'schedular': {'sched': 'cosine', 'lr': 5e-5, 'epochs': 30, 'min_lr': 1e-5, 'decay_rate': 1,
                      'warmup_lr': 5e-5, 'warmup_epochs': 20, 'cooldown_epochs': 0},
        'optimizer': {'opt': 'adamW', 'lr': 5e-5, 'weight_decay': 0.02},
    }
