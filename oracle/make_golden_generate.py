"""TEST INFRASTRUCTURE ONLY -- tests/golden/generate_tiny.pt from the UNMODIFIED reference's inference code.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_generate        (build container; needs /root/reference)

The reference's d_smiles2pv.py / d_pv2smiles_*.py import rdkit-based modules at import time, which this image lacks, so
the functions are not imported as modules: their `def`s are taken from the reference files with `ast` AT GENERATION TIME
and executed unmodified against the reference `SPMM` (built through oracle/ref_shim.py, name-seeded weights):
  * d_smiles2pv.py `generate` (:14-27) driven by the 53-step loop of `pv_generate` (:44-52, re-typed here: it is
    interleaved with file / tokenizer I/O in the reference),
  * d_pv2smiles_single.py `generate` (:26-51) and d_pv2smiles_batched.py `evaluate` (:17-59), the latter run with a stub
    tokenizer whose detokenisation returns the token ids, so the winning beam comes back as ids."""
import ast
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import ref_shim  # noqa: E402
from spmm_b200 import synth  # noqa: E402

CFG = os.path.join(REPO, "spmm_b200", "configs")


def reference_functions(filename, names, extra_globals):
    src = open(os.path.join(ref_shim.REFERENCE_DIR, filename)).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    mod = ast.Module(body=keep, type_ignores=[])
    ns = dict(extra_globals)
    exec(compile(mod, os.path.join(ref_shim.REFERENCE_DIR, filename), "exec"), ns)
    return [ns[n] for n in names]


class StubTokenizer:
    cls_token_id, sep_token_id = 2, 3

    @staticmethod
    def convert_ids_to_tokens(ids):
        return [str(int(i)) for i in ids]

    @staticmethod
    def convert_tokens_to_string(tokens):
        return " ".join(tokens)


def main():
    cfg = synth.pretrain_config(os.path.join(CFG, "config_tiny_text.json"), os.path.join(CFG, "config_tiny_property.json"),
                                queue_size=96, batch_size=6)
    torch.manual_seed(0)
    model = ref_shim.build_reference(cfg)
    synth.fill_by_name(model)
    model.eval()
    pv, ids, mask, lens = synth.synthetic_batch(6, seed=4321)
    (pv_step,) = reference_functions("d_smiles2pv.py", ["generate"], {"torch": torch})
    with torch.no_grad():
        text_embeds = model.text_encoder.bert(ids, attention_mask=mask, return_dict=True, mode='text').last_hidden_state
        prop_input = model.property_cls.expand(ids.shape[0], -1, -1)
        prediction = []
        for _ in range(53):                                      # d_smiles2pv.py:47-52
            output = pv_step(model, prop_input, text_embeds, mask)
            prediction.append(output)
            output = model.property_embed(output.unsqueeze(2))
            prop_input = torch.cat([prop_input, output], dim=1)
        smiles2pv = torch.stack(prediction, dim=-1).squeeze(1)

    from torch.distributions.categorical import Categorical
    (tok_step,) = reference_functions("d_pv2smiles_single.py", ["generate"], {"torch": torch, "Categorical": Categorical})
    (evaluate,) = reference_functions("d_pv2smiles_batched.py", ["evaluate"],
                                      {"torch": torch, "np": np, "generate": tok_step, "tqdm": lambda x: x})
    # a random-init model would never rank [SEP] among its two best tokens, and the reference's `evaluate` crashes
    # when no beam finishes (candidate_k[0] on an empty list): raise the [SEP] logit so that beams end after a few tokens
    sep_bias = float(os.environ.get("SPMM_GOLDEN_SEP_BIAS", "0.0"))
    with torch.no_grad():
        model.text_encoder.cls.predictions.bias[3] += sep_bias
    beams, first = [], []
    with torch.no_grad():
        for b in range(3):
            prop = pv[b:b + 1]
            property1 = model.property_embed(prop.unsqueeze(2))
            properties = torch.cat([model.property_cls.expand(1, -1, -1), property1], dim=1)
            pe = model.property_encoder(inputs_embeds=properties, return_dict=True).last_hidden_state
            vals, idx = tok_step(model, pe, torch.tensor([[2]]), stochastic=False, k=5)
            first.append((vals[0].clone(), idx[0].clone()))
            _, cand = evaluate(model, [(prop, ["[CLS]"])], StubTokenizer, "cpu", stochastic=False, k=2)
            beams.append([int(t) for t in cand[0].split()] if cand[0] else [])
    out = {"ids": ids, "mask": mask, "pv": pv, "smiles2pv": smiles2pv, "first_step": first, "beam_best": beams,
           "sep_bias": sep_bias}
    path = os.path.join(REPO, "tests", "golden", "generate_tiny.pt")
    torch.save(out, path)
    print("smiles2pv[0,:5]", smiles2pv[0, :5].tolist(), "beams", beams, "size", os.path.getsize(path))


if __name__ == "__main__":
    main()
