"""WordPiece tokenizer with the call surface SPMM uses (reference SPMM_pretrain.py:19-20, SPMM_models.py:352,
d_smiles2pv.py:43,61, d_pv2smiles_batched.py:29-55): `tok(list_of_str, padding='longest', truncation=True,
max_length=100, return_tensors="pt").to(device)` -> `.input_ids`, `.attention_mask`; `cls_token_id`, `sep_token_id`,
`convert_ids_to_tokens`, `convert_tokens_to_string`.  The encoding itself runs in the C-ABI library
(`spmm_wordpiece_encode_batch`, csrc/tokenizer.cu): a whole batch per call, straight into pinned int64 buffers."""
import collections
import ctypes as C

import torch

from . import _lib


class BatchEncoding(dict):
    """dict with attribute access and `.to(device)` (what the reference reads from HF's BatchEncoding)."""
    __getattr__ = dict.__getitem__

    def to(self, device, non_blocking=True):
        return BatchEncoding({k: v.to(device, non_blocking=non_blocking) for k, v in self.items()})


class WordPieceTokenizer:
    def __init__(self, vocab_file, do_lower_case=False, do_basic_tokenize=False, unk_token="[UNK]", sep_token="[SEP]",
                 pad_token="[PAD]", cls_token="[CLS]", max_input_chars_per_word=250):
        if do_lower_case or do_basic_tokenize:
            raise NotImplementedError("SPMM tokenises SMILES with do_lower_case=False, do_basic_tokenize=False")
        self.vocab = collections.OrderedDict()
        with open(vocab_file, "r", encoding="utf-8") as f:
            for i, line in enumerate(f.readlines()):
                self.vocab.setdefault(line.rstrip("\n"), i)
        self.ids_to_tokens = collections.OrderedDict((i, t) for t, i in self.vocab.items())
        self.unk_token, self.sep_token, self.pad_token, self.cls_token = unk_token, sep_token, pad_token, cls_token
        self.unk_token_id, self.sep_token_id = self.vocab[unk_token], self.vocab[sep_token]
        self.pad_token_id, self.cls_token_id = self.vocab[pad_token], self.vocab[cls_token]
        self.max_input_chars_per_word = max_input_chars_per_word
        self.wordpiece_tokenizer = None          # the reference assigns a WordpieceTokenizer here; kept as an attribute
        toks = [self.ids_to_tokens.get(i, "") for i in range(max(self.ids_to_tokens) + 1)]
        self._c_tokens = (C.c_char_p * len(toks))(*[t.encode("utf-8") for t in toks])
        self._handle = _lib.lib().spmm_wordpiece_create(self._c_tokens, len(toks), self.unk_token_id, max_input_chars_per_word)
        if not self._handle:
            raise _lib.SpmmKernelError("spmm_wordpiece_create failed")

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                _lib.lib().spmm_wordpiece_destroy(h)
            except Exception:  # noqa: BLE001  (interpreter shutdown)
                pass

    def __len__(self):
        return len(self.vocab)

    def __call__(self, text, padding="longest", truncation=True, max_length=100, return_tensors="pt", pin_memory=None):
        texts = [text] if isinstance(text, str) else list(text)
        if padding not in ("longest", True):
            raise NotImplementedError("only padding='longest' (the reference's setting) is supported")
        limit = max_length if truncation else 1 << 20
        n = len(texts)
        enc = [t.encode("utf-8") for t in texts]
        ld = min(limit, max(len(b) for b in enc) + 2) if n else 2
        pin = torch.cuda.is_available() if pin_memory is None else pin_memory
        ids = torch.empty((n, ld), dtype=torch.int64, pin_memory=pin)
        mask = torch.empty((n, ld), dtype=torch.int64, pin_memory=pin)
        arr = (C.c_char_p * n)(*enc)
        w = _lib.lib().spmm_wordpiece_encode_batch(self._handle, arr, n, limit, self.cls_token_id, self.sep_token_id,
                                                   self.pad_token_id, ids.data_ptr(), mask.data_ptr(), ld)
        if w < 0:
            raise _lib.SpmmKernelError("spmm_wordpiece_encode_batch returned %d" % w)
        out = BatchEncoding(input_ids=ids[:, :w], attention_mask=mask[:, :w],
                            token_type_ids=torch.zeros((n, w), dtype=torch.int64))
        if return_tensors != "pt":
            out = BatchEncoding({k: v.tolist() for k, v in out.items()})
        return out

    def tokenize(self, text):
        e = self(text, truncation=False, pin_memory=False)
        n = int(e["attention_mask"][0].sum())
        return self.convert_ids_to_tokens(e["input_ids"][0, 1:n - 1])

    def convert_ids_to_tokens(self, ids):
        if torch.is_tensor(ids):
            ids = ids.tolist()
        if isinstance(ids, int):
            return self.ids_to_tokens.get(ids, self.unk_token)
        return [self.ids_to_tokens.get(int(i), self.unk_token) for i in ids]

    def convert_tokens_to_ids(self, tokens):
        if isinstance(tokens, str):
            return self.vocab.get(tokens, self.unk_token_id)
        return [self.vocab.get(t, self.unk_token_id) for t in tokens]

    @staticmethod
    def convert_tokens_to_string(tokens):
        return " ".join(tokens).replace(" ##", "").strip()
