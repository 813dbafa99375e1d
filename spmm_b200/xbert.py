"""Host-side mirror of the reference's xbert.py for the SPMM hot path.

Same class names, attribute tree and state-dict keys as the reference (`BertForMaskedLM` xbert.py:1352,
`BertModel` :846, `BertLayer` :454, ...), same keyword interface for the call patterns the SPMM scripts use
(`input_ids` / `inputs_embeds` / `encoder_embeds`, `encoder_hidden_states`, `is_decoder`, `mode`,
`return_logits`).  The nn.Linear / nn.Embedding / nn.LayerNorm objects are parameter containers only: the
forward pass never calls them, it launches the block kernels in spmm_b200/ops.py on views of the flat
parameter arena (spmm_b200/arena.py).  Activations are bf16 [tokens, hidden].
"""
import contextlib
import json
import os
from types import SimpleNamespace

import torch
from torch import nn

from . import ops


class BertConfig:
    """Reads config_bert*.json unchanged (including the string "True" of add_cross_attention)."""

    def __init__(self, **kw):
        self.hidden_dropout_prob = 0.1
        self.attention_probs_dropout_prob = 0.1
        self.layer_norm_eps = 1e-12
        self.initializer_range = 0.02
        self.pad_token_id = 0
        self.type_vocab_size = 2
        self.max_position_embeddings = 512
        self.add_cross_attention = False
        for k, v in kw.items():
            if v in ("True", "False"):
                v = v == "True"
            setattr(self, k, v)
        if not hasattr(self, "encoder_width"):
            self.encoder_width = self.hidden_size
        if not hasattr(self, "fusion_layer"):
            self.fusion_layer = self.num_hidden_layers

    @classmethod
    def from_json_file(cls, path):
        with open(path) as f:
            return cls(**json.load(f))

    def to_dict(self):
        return dict(self.__dict__)


def _validation_enabled():
    """Input validation costs a host sync, so it runs where a sync is acceptable: inference (grad disabled) and
    whenever SPMM_CHECK_INPUTS=1; never inside a CUDA-graph capture.  SPMM_CHECK_INPUTS=0 turns it off everywhere."""
    env = os.environ.get("SPMM_CHECK_INPUTS")
    if env == "0":
        return False
    if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
        return False
    return env == "1" or not torch.is_grad_enabled()


class MaskInfo:
    """A right-padded 0/1 attention mask reduced to what the kernels consume: per-sequence valid lengths.

    The reference honours arbitrary masks (xbert.py:889-948); the kernels here take prefix lengths, which is what
    `padding='longest'` produces.  A mask that is not of prefix form (left padding, holes) would silently mask the wrong
    keys, so it is rejected where validation runs (`_validation_enabled`)."""

    def __init__(self, mask=None, kv_len=None):
        self.mask = mask
        if kv_len is None and mask is not None:
            if mask.dim() != 2:
                raise ValueError("attention masks must be [batch, length] 0/1 tensors, got shape %s" % (tuple(mask.shape),))
            if mask.shape[1] > 1 and _validation_enabled():
                if not bool((mask[:, 1:] <= mask[:, :-1]).all()):
                    raise ValueError("spmm_b200 kernels need right-padded (prefix-form) attention masks: every row must be "
                                     "1...1 0...0; got a mask with a hole or left padding")
            kv_len = mask.sum(dim=1).to(torch.int32)
        self.kv_len = kv_len

    @staticmethod
    def of(m):
        if m is None or isinstance(m, MaskInfo):
            return m
        return MaskInfo(m)


class BertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.config = config


class BertSelfAttention(nn.Module):
    def __init__(self, config, is_cross_attention):
        super().__init__()
        if config.hidden_size % config.num_attention_heads != 0:
            raise ValueError("hidden size %d is not a multiple of the number of heads %d" %
                             (config.hidden_size, config.num_attention_heads))
        kv_in = config.encoder_width if is_cross_attention else config.hidden_size
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(kv_in, config.hidden_size)
        self.value = nn.Linear(kv_in, config.hidden_size)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertAttention(nn.Module):
    def __init__(self, config, is_cross_attention=False):
        super().__init__()
        self.self = BertSelfAttention(config, is_cross_attention)
        self.output = BertSelfOutput(config)
        self.is_cross_attention = is_cross_attention


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)


class BertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertLayer(nn.Module):
    def __init__(self, config, layer_num):
        super().__init__()
        self.attention = BertAttention(config)
        self.has_cross_attention = layer_num >= config.fusion_layer
        if self.has_cross_attention:
            self.layer_num = layer_num
            self.crossattention = BertAttention(config, is_cross_attention=True)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)


class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layer = nn.ModuleList([BertLayer(config, i) for i in range(config.num_hidden_layers)])


class BertPredictionHeadTransform(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)


class BertLMPredictionHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.decoder.bias = self.bias     # same alias as the reference (xbert.py:688-691)


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.predictions = BertLMPredictionHead(config)


def _init_bert_weights(module, std):
    """Reference init (xbert.py:742-752): N(0, 0.02) for Linear/Embedding (pad row included), LN = (1, 0)."""
    for m in module.modules():
        if isinstance(m, (nn.Linear, nn.Embedding)):
            m.weight.data.normal_(mean=0.0, std=std)
        if isinstance(m, nn.LayerNorm):
            m.bias.data.zero_()
            m.weight.data.fill_(1.0)
        if isinstance(m, nn.Linear) and m.bias is not None:
            m.bias.data.zero_()


# Activations are bf16 inside SPMM.forward (raw_outputs()); callers that reach into the sub-modules the way the d_*.py
# scripts do (d_smiles2pv.py:15-25, d_pv2smiles_batched.py:25-27) pass fp32 embeddings and apply fp32 torch heads to the
# result, as with the reference's fp32 modules: they get fp32 `last_hidden_state` / logits back.
_RAW = [False]


@contextlib.contextmanager
def raw_outputs():
    old, _RAW[0] = _RAW[0], True
    try:
        yield
    finally:
        _RAW[0] = old


def _bf16(t):
    return t if t is None or t.dtype == torch.bfloat16 else t.to(torch.bfloat16)


# Gradient all-reduce overlapped with backward (trainer.GradOverlap): while set, the first pass of a step that runs a layer
# marks the layer's input, and the marker's backward hands the layer's gradient range to the communication stream.
_OVERLAP = [None]


def set_grad_overlap(ctx):
    _OVERLAP[0] = ctx


class ModelOutput(SimpleNamespace):
    def __getitem__(self, i):
        return (self.last_hidden_state,)[i]


class BertModel(nn.Module):
    """xbert.py:846.  forward() covers the three input kinds and three layer ranges the SPMM scripts use."""

    def __init__(self, config, add_pooling_layer=False):
        super().__init__()
        if add_pooling_layer:
            raise NotImplementedError("BertPooler is never instantiated on the SPMM path (xbert.py:1359)")
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config)
        self.pooler = None
        _init_bert_weights(self, config.initializer_range)

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def _bundles(self):
        b = getattr(self, "_spmm_bundles", None)
        if b is None:
            raise RuntimeError("parameter arena not built: call SPMM.build_arenas() / run through SPMM (spmm_b200/arena.py)")
        return b

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, encoder_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None,
                past_key_values=None, use_cache=None, output_attentions=None, output_hidden_states=None,
                return_dict=True, is_decoder=False, mode='multi_modal', causal_from=None, encoder_index=None):
        """`causal_from` (SPMM.forward only): the batch holds two passes over the same weights - rows [0, causal_from)
        attend bidirectionally, rows [causal_from, B) causally - so both share every GEMM / LayerNorm launch.
        `encoder_index` (int32 [B] on the device, SPMM.forward only): `encoder_hidden_states` holds DISTINCT states and
        query batch element b cross-attends to states[encoder_index[b]] (`encoder_attention_mask` stays per b)."""
        if past_key_values is not None or output_attentions or output_hidden_states or head_mask is not None:
            raise NotImplementedError("unused on the SPMM hot path")
        cfg = self.config
        bd = self._bundles()
        inputs_embeds, encoder_embeds, encoder_hidden_states = _bf16(inputs_embeds), _bf16(encoder_embeds), _bf16(encoder_hidden_states)
        p_hid = cfg.hidden_dropout_prob if self.training else 0.0
        p_att = cfg.attention_probs_dropout_prob if self.training else 0.0
        if input_ids is not None and inputs_embeds is not None:
            raise ValueError("You cannot specify both input_ids and inputs_embeds at the same time")
        if input_ids is not None:
            B, T = input_ids.shape
            if T > cfg.max_position_embeddings:
                raise ValueError("sequence length %d exceeds max_position_embeddings %d" % (T, cfg.max_position_embeddings))
            if _validation_enabled() and input_ids.numel() > 0:
                lo, hi = int(input_ids.min()), int(input_ids.max())
                if lo < 0 or hi >= cfg.vocab_size:
                    raise IndexError("token id out of range [0, %d): min %d max %d" % (cfg.vocab_size, lo, hi))
            x = ops.embed_text(input_ids.contiguous(), bd.emb, p_hid, bd.anchor)
        elif inputs_embeds is not None:
            B, T = inputs_embeds.shape[:2]
            x = ops.embed_inputs(inputs_embeds.contiguous(), bd.emb, p_hid, bd.anchor)
        elif encoder_embeds is not None:
            B, T = encoder_embeds.shape[:2]
            x = encoder_embeds.contiguous().view(B * T, -1)
        else:
            raise ValueError("You have to specify either input_ids or inputs_embeds or encoder_embeds")
        smask = MaskInfo.of(attention_mask)
        self_geom = SimpleNamespace(B=B, Tq=T, Tk=T, kv_len=None if smask is None else smask.kv_len,
                                    causal=(1 + int(causal_from)) if causal_from is not None else int(bool(is_decoder)))
        enc = None
        cross_geom = None
        if encoder_hidden_states is not None:
            if isinstance(encoder_hidden_states, (list, tuple)):
                raise NotImplementedError("list-valued encoder_hidden_states is unused by the SPMM scripts")
            Be, Te = encoder_hidden_states.shape[:2]
            if Be != B and encoder_index is None:
                raise ValueError("encoder batch %d != query batch %d" % (Be, B))
            enc = encoder_hidden_states.contiguous().view(Be * Te, -1)
            cmask = MaskInfo.of(encoder_attention_mask)
            cross_geom = SimpleNamespace(B=B, Tq=T, Tk=Te, kv_len=None if cmask is None else cmask.kv_len, causal=False,
                                         kv_index=encoder_index)
        fl, nl = cfg.fusion_layer, cfg.num_hidden_layers
        lo, hi = {'text': (0, fl), 'fusion': (fl, nl), 'multi_modal': (0, nl)}[mode]
        ov = _OVERLAP[0]
        for i in range(lo, hi):
            lw = bd.layers[i]
            if ov is not None and lw.grad_range is not None and torch.is_grad_enabled() and x.requires_grad and ov.claim(lw.grad_range):
                x = ops.grad_ready(x, ov.callback(lw.grad_range))
            x = ops.attn_block(x, None, lw.attn, self_geom, p_att, p_hid, bd.anchor)
            if lw.cross is not None:
                if enc is None:
                    raise AssertionError("encoder_hidden_states must be given for cross-attention layers")
                x = ops.attn_block(x, enc, lw.cross, cross_geom, p_att, p_hid, bd.anchor)
            x = ops.ffn_block(x, lw.ffn, p_hid, bd.anchor)
        out = x.view(B, T, -1)
        if not _RAW[0]:
            out = out.float()
        if not return_dict:
            return (out,)
        return ModelOutput(last_hidden_state=out, pooler_output=None, past_key_values=None, hidden_states=None,
                           attentions=None, cross_attentions=None)


class BertForMaskedLM(nn.Module):
    """xbert.py:1352; only the `return_logits=True` path is used by SPMM (SPMM_models.py:215-231, d_pv2smiles)."""

    LOGIT_LD_ALIGN = 64

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.bert = BertModel(config, add_pooling_layer=False)
        self.cls = BertOnlyMLMHead(config)
        _init_bert_weights(self.cls, config.initializer_range)
        self.cls.predictions.bias.data.zero_()
        # weight tying (transformers 4.30 init_weights semantics): decoder.weight IS word_embeddings.weight
        self.cls.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight

    def get_output_embeddings(self):
        return self.cls.predictions.decoder

    def logit_ld(self):
        a = self.LOGIT_LD_ALIGN
        return (self.config.vocab_size + a - 1) // a * a

    def forward(self, input_ids=None, attention_mask=None, encoder_embeds=None, encoder_hidden_states=None,
                encoder_attention_mask=None, inputs_embeds=None, return_dict=True, is_decoder=False,
                mode='multi_modal', return_logits=False, return_hidden=False, **unused):
        with raw_outputs():
            h = self.bert(input_ids, attention_mask=attention_mask, inputs_embeds=inputs_embeds,
                          encoder_embeds=encoder_embeds, encoder_hidden_states=encoder_hidden_states,
                          encoder_attention_mask=encoder_attention_mask, return_dict=True, is_decoder=is_decoder,
                          mode=mode).last_hidden_state
        if return_hidden:
            return h if _RAW[0] else h.float()
        if not return_logits:
            raise NotImplementedError("only return_logits=True is used on the SPMM path (the loss is fused in SPMM.forward)")
        B, T, H = h.shape
        if torch.is_grad_enabled() and h.requires_grad:
            raise NotImplementedError("differentiable logits are produced inside ops.lm_head_loss; call under no_grad")
        V = self.config.vocab_size
        logits = ops.lm_logits(h.view(B * T, H), self.bert._bundles().head, V, self.logit_ld())
        logits = logits.view(B, T, -1)[:, :, :V]
        return logits if _RAW[0] else logits.float()
