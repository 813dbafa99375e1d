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


def train_step(model, optimizer, prop, text_input_ids, text_attention_mask, alpha, **fwd_kw):
    """zero_grad -> forward -> backward -> grad all-reduce -> clip(5.) + AdamW.  Returns the 4 losses (device)."""
    optimizer.zero_grad()
    losses = model(prop, text_input_ids, text_attention_mask, alpha=alpha, **fwd_kw)
    loss = losses[0] + losses[1] + losses[2] + losses[3]
    loss.backward()
    W = world_size()
    A = model.arena()
    if W > 1:
        dist.all_reduce(A.G[A.adam_start:], op=dist.ReduceOp.SUM)
    optimizer.step(skip_flag=model.last_aux["nan_flag"], grad_scale=1.0 / W)
    return losses


def alpha_schedule(config_alpha, epoch, batch_idx, loader_len):
    """SPMM_models.py:355."""
    return config_alpha if epoch > 0 else config_alpha * min(1., batch_idx / max(1, loader_len))
