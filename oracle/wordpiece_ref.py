"""TEST INFRASTRUCTURE (never imported by spmm_b200/): plain-Python restatement of the tokenisation the reference
delegates to transformers==4.30.1 (`BertTokenizer(do_basic_tokenize=False)` + `WordpieceTokenizer`, call sites
SPMM_pretrain.py:19-20, SPMM_models.py:352): whitespace split; per word greedy longest-match-first with "##"
continuation pieces, the whole word becoming [UNK] on any miss or above 250 characters; [CLS] ... [SEP] added, sequence
truncated to max_length, right-padded to the longest of the batch.  Pinned by tests/golden/tokenizer.json, produced by
oracle/make_golden_tokenizer.py with the installed transformers' own WordpieceTokenizer on the reference's vocabulary."""


def wordpiece(word, vocab, unk="[UNK]", max_chars=250):
    if len(word) > max_chars:
        return [unk]
    out, start = [], 0
    while start < len(word):
        end, cur = len(word), None
        while start < end:
            sub = word[start:end]
            if start > 0:
                sub = "##" + sub
            if sub in vocab:
                cur = sub
                break
            end -= 1
        if cur is None:
            return [unk]
        out.append(cur)
        start = end
    return out


def encode_batch(texts, vocab, max_length=100, cls="[CLS]", sep="[SEP]", pad="[PAD]"):
    rows = []
    for t in texts:
        pieces = [p for w in t.split() for p in wordpiece(w, vocab)]
        pieces = pieces[:max_length - 2]
        rows.append([vocab[cls]] + [vocab[p] for p in pieces] + [vocab[sep]])
    width = max(len(r) for r in rows)
    ids = [r + [vocab[pad]] * (width - len(r)) for r in rows]
    mask = [[1] * len(r) + [0] * (width - len(r)) for r in rows]
    return ids, mask
