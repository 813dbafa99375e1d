"""Lightning-compatible checkpoints (reference SPMM_pretrain.py:24-37: `ModelCheckpoint(filename='checkpoint_{epoch}',
every_n_train_steps=10000)`, `torch.load(path)['state_dict']` + `load_state_dict(strict=False)` in SPMM_pretrain.py:25-26 and
d_smiles2pv.py:132-143).  A `.ckpt` is a pickled dict whose 'state_dict' holds the reference's keys, shapes and dtypes
(queues as [E, Q], tied LM decoder, `*_m` momentum twins, `queue_ptr`), so checkpoints written by the reference load
here and the other way round.  Optimiser / scheduler state goes into the same Lightning slots for resume."""
import os

import torch

LIGHTNING_VERSION = "2.0.3"      # the reference's requirements.txt pin; only used as a label


def lightning_checkpoint(model, optimizer=None, scheduler=None, epoch=0, global_step=0):
    sd = {k: v.detach().to("cpu").clone() for k, v in model.state_dict().items()}
    ckpt = {"epoch": int(epoch), "global_step": int(global_step), "pytorch-lightning_version": LIGHTNING_VERSION,
            "state_dict": sd, "loops": None, "callbacks": {}, "optimizer_states": [], "lr_schedulers": []}
    if optimizer is not None:
        osd = optimizer.state_dict()
        ckpt["optimizer_states"] = [{k: (v.detach().to("cpu").clone() if torch.is_tensor(v) else v) for k, v in osd.items()}]
    if scheduler is not None:
        ckpt["lr_schedulers"] = [{k: v for k, v in vars(scheduler).items() if isinstance(v, (int, float, list, bool))}]
    return ckpt


def save(model, dirpath, epoch, global_step, optimizer=None, scheduler=None, filename="checkpoint_{epoch}"):
    """Writes `<dirpath>/checkpoint_epoch=<epoch>.ckpt` (Lightning expands '{epoch}' to 'epoch=<n>')."""
    os.makedirs(dirpath, exist_ok=True)
    name = filename.replace("{epoch}", "epoch=%d" % epoch).replace("{step}", "step=%d" % global_step) + ".ckpt"
    path = os.path.join(dirpath, name)
    torch.save(lightning_checkpoint(model, optimizer, scheduler, epoch, global_step), path)
    return path


def load(model, path, optimizer=None, drop_queues=False, map_location="cpu"):
    """`load_state_dict(checkpoint['state_dict'], strict=False)`; `drop_queues` mirrors the d_*.py scripts, which delete
    the queue entries before loading into a `no_train=True` model.  Returns (incompatible-keys message, checkpoint)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    sd = dict(ckpt["state_dict"])
    if drop_queues:
        for k in [k for k in sd if "queue" in k]:
            del sd[k]
    msg = model.load_state_dict(sd, strict=False)
    if optimizer is not None and ckpt.get("optimizer_states"):
        optimizer.load_state_dict(ckpt["optimizer_states"][0])
    return msg, ckpt
