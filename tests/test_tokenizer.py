"""WordPiece tokenisation (the host half of the reference's training_step, SPMM_models.py:352): the oracle restatement
and the native C-ABI tokenizer against golden vectors produced with transformers' own WordpieceTokenizer on the
reference vocabulary (oracle/make_golden_tokenizer.py).  CPU only."""
import json
import os

import pytest
import torch

from tests.util import REPO

GOLD = json.load(open(os.path.join(REPO, "tests", "golden", "tokenizer.json")))


@pytest.fixture(scope="module")
def vocab_file(tmp_path_factory):
    p = tmp_path_factory.mktemp("vocab") / "vocab_bpe_300.txt"
    p.write_text("\n".join(GOLD["vocab"]) + "\n", encoding="utf-8")
    return str(p)


def test_oracle_restatement_matches_golden():
    from oracle import wordpiece_ref
    vocab = {}
    for i, t in enumerate(GOLD["vocab"]):
        vocab.setdefault(t, i)
    for text, want in zip(GOLD["texts"], GOLD["ids"]):
        ids, mask = wordpiece_ref.encode_batch([text], vocab, GOLD["max_length"])
        assert ids[0] == want, text
    # the survey's known answer: model input = column 0 dropped
    assert GOLD["ids"][0][1:] == [2, 282, 114, 127, 16, 91, 16, 72, 147, 105, 10, 162, 216, 156, 3]


def test_native_tokenizer_matches_golden_and_pads_like_the_reference(vocab_file):
    from spmm_b200.tokenizer import WordPieceTokenizer
    tok = WordPieceTokenizer(vocab_file=vocab_file, do_lower_case=False, do_basic_tokenize=False)
    assert (tok.cls_token_id, tok.sep_token_id, tok.pad_token_id, tok.unk_token_id) == (2, 3, 0, 1) and len(tok) == 300
    enc = tok(GOLD["texts"], padding="longest", truncation=True, max_length=100, return_tensors="pt", pin_memory=False)
    ids, mask = enc.input_ids, enc.attention_mask
    assert ids.dtype == torch.int64 and ids.shape == mask.shape and ids.shape[1] == max(len(r) for r in GOLD["ids"]) == 100
    for row, m, want in zip(ids.tolist(), mask.tolist(), GOLD["ids"]):
        n = len(want)
        assert row[:n] == want and m[:n] == [1] * n
        assert all(x == 0 for x in row[n:]) and all(x == 0 for x in m[n:])
    # padding='longest' is per batch
    small = tok(GOLD["texts"][:3], padding="longest", truncation=True, max_length=100, return_tensors="pt", pin_memory=False)
    assert small.input_ids.shape[1] == max(len(r) for r in GOLD["ids"][:3])
    # single string, .to(), and the detokenisation used by d_pv2smiles_batched.py:55
    one = tok(GOLD["texts"][0], pin_memory=False).to("cpu")
    toks = tok.convert_ids_to_tokens(one.input_ids[0, 1:-1])
    assert tok.convert_tokens_to_string(toks).replace("[CLS]", "") == GOLD["texts"][0].replace("[CLS]", "")
    assert tok.tokenize(GOLD["texts"][1]) == tok.convert_ids_to_tokens(GOLD["ids"][1][1:-1])
