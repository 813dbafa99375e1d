"""Generates tests/golden/tokenizer.json from the reference's vocabulary and example molecules with the installed
transformers' WordpieceTokenizer (the class the reference plugs into its BertTokenizer, SPMM_pretrain.py:20).
Run in the build container: /root/reference is not available on the GPU box, the fixture is."""
import json
import os
import random

from transformers.models.bert.tokenization_bert_legacy import WordpieceTokenizer

REF = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tokens = [l.rstrip("\n") for l in open(os.path.join(REF, "vocab_bpe_300.txt"), encoding="utf-8")]
vocab = {}
for i, t in enumerate(tokens):
    vocab.setdefault(t, i)
wp = WordpieceTokenizer(vocab=vocab, unk_token="[UNK]", max_input_chars_per_word=250)
smiles = [l.strip() for l in open(os.path.join(REF, "s2p_input.txt")) if l.strip()]
rng = random.Random(7)
pieces = [t[2:] for t in tokens if t.startswith("##")]
for _ in range(160):                                   # SMILES-like strings assembled from vocabulary pieces and atoms
    n = rng.randint(1, 60)
    s = "".join(rng.choice(pieces) if rng.random() < 0.8 else rng.choice("CNOcn()=#123[]+-@HSFl") for _ in range(n))
    smiles.append(s)
smiles += ["", "C", "Q", "C C", "CC(=O)O  c1ccccc1", "C" * 300, "N" * 97 + "O", "c1ccccc1" * 20, "##C", "[CLS]", "Zr(C)"]
texts = ["[CLS]" + s for s in smiles]
rows = []
for t in texts:
    toks = wp.tokenize(t)
    toks = toks[:98]
    rows.append([vocab["[CLS]"]] + [vocab[x] for x in toks] + [vocab["[SEP]"]])
json.dump({"vocab": tokens, "texts": texts, "ids": rows, "max_length": 100},
          open(os.path.join(REPO, "tests", "golden", "tokenizer.json"), "w"))
print("wrote", len(texts), "cases; known answer:", rows[0])
