"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = os.path.join(REPO, "spmm_b200", "configs")
CASE_CFG = {
    "tiny_b6": ("config_tiny_text.json", "config_tiny_property.json", 96),
    "full_b8": ("config_bert.json", "config_bert_property.json", 36864),
}


def load_golden(case):
    return torch.load(os.path.join(REPO, "tests", "golden", case + ".pt"), weights_only=False)


def load_cfgs(case):
    tj, pj, q = CASE_CFG[case]
    return json.load(open(os.path.join(CFG, tj))), json.load(open(os.path.join(CFG, pj))), q


def sample_idx(numel, n=256):
    step = max(1, numel // n)
    return torch.arange(0, numel, step)[:n]


def oracle_state(golden, device="cpu", dtype=torch.float32):
    """Name-seeded oracle state with reference key names; trainable leaves require grad."""
    from spmm_b200 import synth
    P = synth.state_from_keys(golden["state_dict_keys"])
    out, seen = {}, {}
    for k, v in P.items():
        if id(v) in seen:
            out[k] = out[seen[id(v)]]
            continue
        seen[id(v)] = k
        t = v.to(device)
        if t.is_floating_point():
            t = t.to(dtype)
            frozen = any(k.startswith(m) for m in ("property_encoder_m.", "text_encoder_m.", "property_proj_m.",
                                                    "text_proj_m.")) or k.endswith("queue")
            if not frozen:
                t.requires_grad_(True)
        out[k] = t
    return out
