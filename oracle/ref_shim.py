"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference from /root/reference.

The reference (SPMM_models.py / xbert.py) targets transformers 4.30 + pytorch-lightning 2.0;
this container has transformers 5.5 and no lightning.  `load_reference()` applies pre-import
monkey patches (nothing under /root/reference is edited or copied) so that
`SPMM_models.SPMM` constructs and runs on CPU.  It is used by `oracle/make_golden.py` to
generate the fixtures in tests/golden/ and by `bench.py --impl reference` when the reference
tree travels with the job.  /root/reference does not exist on the GPU box, so nothing in the
`-m gpu` tests may call this.

Patches:
  * transformers.modeling_utils gets the three helper names xbert.py:54-59 imports
  * PreTrainedModel.get_head_mask (xbert.py:1052) returns [None]*n
  * a fake `pytorch_lightning` whose LightningModule is nn.Module (SPMM_models.py:6)
  * BertConfig.from_json_file tolerant of the string "True" (config_bert.json)
  * BertPreTrainedModel.init_weights with transformers-4.30 semantics incl. weight tying
"""
import json
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_DIR = os.environ.get("SPMM_REFERENCE_DIR", "/root/reference")
_loaded = {}


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "SPMM_models.py"))


def load_reference():
    """Returns the reference's `SPMM` class (imported from REFERENCE_DIR, unmodified)."""
    if "SPMM" in _loaded:
        return _loaded["SPMM"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    sys.dont_write_bytecode = True
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    mu.prune_linear_layer = pu.prune_linear_layer

    def _no_prune(*a, **k):
        raise NotImplementedError("head pruning is dead code on this path")

    mu.find_pruneable_heads_and_indices = _no_prune
    if not hasattr(mu.PreTrainedModel, "get_head_mask"):
        mu.PreTrainedModel.get_head_mask = lambda self, hm, n, *a, **k: [None] * n if hm is None else hm

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = type("LightningModule", (nn.Module,), {})
        sys.modules["pytorch_lightning"] = pl

    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import xbert  # noqa: the reference's own file
    from transformers.models.bert.configuration_bert import BertConfig

    def _from_json_file(cls, path):
        with open(path) as f:
            d = json.load(f)
        d = {k: (v == "True") if v in ("True", "False") else v for k, v in d.items()}
        return cls(**d)

    BertConfig.from_json_file = classmethod(_from_json_file)

    def _init_weights_430(self):
        def once(mod):
            if not getattr(mod, "_is_hf_initialized", False):
                self._init_weights(mod)
                mod._is_hf_initialized = True

        self.apply(once)
        out = self.get_output_embeddings() if hasattr(self, "get_output_embeddings") else None
        if out is not None and getattr(self.config, "tie_word_embeddings", True):
            out.weight = self.get_input_embeddings().weight

    xbert.BertPreTrainedModel.init_weights = _init_weights_430
    xbert.BertPreTrainedModel.get_input_embeddings = (
        lambda self: (self.bert if hasattr(self, "bert") else self).embeddings.word_embeddings)
    from SPMM_models import SPMM  # noqa: the reference's own file

    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="file:///tmp/spmm_ref_pg_%d" % os.getpid(),
                                rank=0, world_size=1)
    _loaded["SPMM"] = SPMM
    return SPMM


def build_reference(config, cwd_relative_ok=True):
    """Constructs the reference SPMM.  `config` paths may be absolute."""
    SPMM = load_reference()
    return SPMM(config=config, tokenizer=None, loader_len=1)
