"""Drop-in `SPMM` for the reference's SPMM_models.py (jinhojsk515/spmm), B200-native underneath.

Same constructor, `forward(property_original, text_input_ids, text_attention_mask, alpha)` signature, sub-module
attribute tree and state-dict keys as the reference class (SPMM_models.py:16-399), so SPMM_pretrain.py and the
d_*.py scripts can import this module instead.  What differs is everything below the Python surface:

  * parameters live in flat HBM arenas (spmm_b200/arena.py); the momentum EMA is one kernel,
  * every encoder block / loss head is a short sequence of hand-written sm_100a kernels (spmm_b200/ops.py),
  * the 2B+8 host syncs of the reference forward are gone: hard negatives are drawn on the device by a
    counter-based sampler, `queue_ptr` advances on the device, the NaN guard is a device flag consumed by the
    enqueue and optimiser kernels,
  * the momentum queues are stored key-major [Q, E]; `state_dict()` still exposes `prop_queue` / `text_queue` as
    [E, Q] like the reference (SPMM_models.py:72-73).

There is no CPU or PyTorch-eager fallback: without the CUDA library the forward raises.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

from . import arena as arena_mod
from . import kernels as K
from . import ops
from .xbert import BertConfig, BertForMaskedLM, MaskInfo

try:                                       # pytorch_lightning is optional (absent in this image)
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:                          # noqa: BLE001
    _Base = nn.Module


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def world_any(flag):
    """MAX over the data-parallel world of a device flag (in place; a no-op on one rank).  The NaN guard of
    SPMM_models.py:132 must be ONE decision for all replicas: the gathered queue rows and the all-reduced gradients are
    global, so a rank that skipped its step alone would diverge from the others for good."""
    if _world() > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    return flag


def gather_world_feats(feats, flag=None):
    """concat_all_gather (reference SPMM_models.py:389-399) for a stacked [2, B, E] feature tensor: ONE collective for
    both modalities; result [2, W*B, E] with rank r's rows at [r*B, (r+1)*B) exactly like torch.cat(tensors_gather).
    `flag` (0-d device float, optional) rides in the same collective: returns (feats, max of the ranks' flags) - the NaN
    guard is one decision for the whole world without a second synchronisation point in the step."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return feats if flag is None else (feats, flag)
    W_ = dist.get_world_size()
    feats = feats.contiguous()
    n = feats.numel()
    send = feats.reshape(-1) if flag is None else torch.cat([feats.reshape(-1), flag.reshape(1).to(feats.dtype)])
    if dist.get_backend() == "nccl":
        gathered = torch.empty((W_, send.numel()), device=feats.device, dtype=feats.dtype)
        dist.all_gather_into_tensor(gathered, send)
    else:
        parts = [torch.empty_like(send) for _ in range(W_)]
        dist.all_gather(parts, send)
        gathered = torch.stack(parts)
    out = gathered[:, :n].reshape((W_,) + tuple(feats.shape)).permute(1, 0, 2, 3).reshape(2, -1, feats.shape[-1]).contiguous()
    if flag is None:
        return out
    return out, gathered[:, n].max().reshape(flag.shape)


class AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


class _StandaloneHooks:
    """What `pl.LightningModule` would provide to `training_step` (SPMM_models.py:348-386), for runs without
    pytorch_lightning: `trainer.fit` attaches the optimiser / scheduler / logger and advances `current_epoch`."""
    current_epoch = 0
    global_rank = 0

    def attach(self, optimizer, scheduler, global_rank=0, log=None):
        self._opt, self._sched, self.global_rank = optimizer, scheduler, global_rank
        self._log = log
        return self

    def optimizers(self):
        return self._opt

    def lr_schedulers(self):
        return self._sched

    def log(self, name, value, prog_bar=False):
        if getattr(self, "_log", None) is not None:
            self._log(name, value)


class SPMM(*((_Base,) if _Base is not nn.Module else (_StandaloneHooks, nn.Module))):
    def __init__(self, tokenizer=None, config=None, loader_len=0, no_train=False):
        super().__init__()
        self.automatic_optimization = False
        self.config = config
        self.tokenizer = tokenizer
        self.training_step_outputs = []
        embed_dim = config['embed_dim']

        bert_config = BertConfig.from_json_file(config['bert_config_text'])
        self.text_encoder = BertForMaskedLM(config=bert_config)
        text_width = self.text_encoder.config.hidden_size
        property_width = text_width
        self.property_proj = nn.Linear(property_width, embed_dim)
        self.text_proj = nn.Linear(text_width, embed_dim)
        self.itm_head = nn.Linear(text_width * 2, 2)
        self.property_embed = nn.Linear(1, property_width)
        bert_config2 = BertConfig.from_json_file(config['bert_config_property'])
        self.property_encoder = BertForMaskedLM(config=bert_config2).bert
        self.property_mtr_head = nn.Sequential(nn.Linear(property_width, property_width), nn.GELU(),
                                               nn.LayerNorm(property_width, bert_config.layer_norm_eps),
                                               nn.Linear(property_width, 1))
        self.property_cls = nn.Parameter(torch.zeros(1, 1, property_width))
        self.property_mask = nn.Parameter(torch.zeros(1, 1, property_width))

        self.property_encoder_m = BertForMaskedLM(config=bert_config2).bert
        self.property_proj_m = nn.Linear(property_width, embed_dim)
        self.text_encoder_m = BertForMaskedLM(config=bert_config)
        self.text_proj_m = nn.Linear(text_width, embed_dim)
        self.model_pairs = [[self.property_encoder, self.property_encoder_m], [self.property_proj, self.property_proj_m],
                            [self.text_encoder, self.text_encoder_m], [self.text_proj, self.text_proj_m]]
        self.copy_params()

        self.no_train = no_train
        self.embed_dim = embed_dim
        self.momentum = 0.995
        self.queue_size = 0
        if not no_train:
            self.temp = nn.Parameter(torch.ones([]) * config['temp'])
            self.mlm_probability = config['mlm_probability']
            self.warmup_steps = config['schedular']['warmup_epochs']
            self.loader_len = loader_len
            self.momentum = config['momentum']
            self.queue_size = config['queue_size']
            # key-major storage [Q, E]; the reference's [E, Q] view is `self.prop_queue` / state_dict()
            self.register_buffer("prop_queue_km", F.normalize(torch.randn(self.queue_size, embed_dim), dim=1), persistent=False)
            self.register_buffer("text_queue_km", F.normalize(torch.randn(self.queue_size, embed_dim), dim=1), persistent=False)
            self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))
        else:
            self.temp = None
        self._register_state_dict_hook(SPMM._queues_to_state_dict)
        self._register_load_state_dict_pre_hook(self._queues_from_state_dict)
        self.register_load_state_dict_post_hook(SPMM._after_load)
        self._arena = None
        self._step = 0
        self.sampler_seed = 0x5EED
        self.last_aux = {}

    # ------------------------------------------------------------------ reference-compatible queue surface
    @property
    def prop_queue(self):
        return self.prop_queue_km.t()

    @property
    def text_queue(self):
        return self.text_queue_km.t()

    @staticmethod
    def _queues_to_state_dict(module, state_dict, prefix, local_metadata):
        if not module.no_train:
            state_dict[prefix + "prop_queue"] = module.prop_queue_km.t().contiguous()
            state_dict[prefix + "text_queue"] = module.text_queue_km.t().contiguous()
        return state_dict

    def _queues_from_state_dict(self, state_dict, prefix, *args):
        for k in ("prop_queue", "text_queue"):
            v = state_dict.pop(prefix + k, None)
            if v is not None and not self.no_train:
                with torch.no_grad():
                    getattr(self, k + "_km").copy_(v.t())

    @staticmethod
    def _after_load(module, incompatible):
        for k in ("prop_queue", "text_queue"):
            if k in incompatible.unexpected_keys:
                incompatible.unexpected_keys.remove(k)
        if module._arena is not None and module._arena.valid_for(module):
            module._arena.refresh_shadows(ema=False)

    # ------------------------------------------------------------------ arenas
    def build_arenas(self, device=None):
        """Moves every parameter into the flat arenas on `device` and builds the kernel-facing views."""
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != "cuda":
            raise K._lib.SpmmKernelError("spmm_b200 runs on CUDA (sm_100a) only: there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        super().to(device)
        A = arena_mod.ParamArena(self, device)
        anchor = torch.zeros(1, device=device, requires_grad=True)
        arena_mod.bert_bundles(A, "text_encoder.bert", self.text_encoder.bert, False, anchor, "text_encoder.cls.predictions")
        arena_mod.bert_bundles(A, "text_encoder_m.bert", self.text_encoder_m.bert, True, anchor, "text_encoder.cls.predictions")
        arena_mod.bert_bundles(A, "property_encoder", self.property_encoder, False, anchor, is_property=True)
        arena_mod.bert_bundles(A, "property_encoder_m", self.property_encoder_m, True, anchor, is_property=True)
        W = {}
        for n in ("property_proj", "text_proj"):
            W[n] = arena_mod.linear_bundle(A, n)
            W[n + "_m"] = arena_mod.linear_bundle(A, n + "_m", momentum=True)
        ns = arena_mod.SimpleNamespace
        W["pv"] = ns(w=A.f32("property_embed.weight").view(-1), b=A.f32("property_embed.bias"),
                     cls=A.f32("property_cls").view(-1), mask=A.f32("property_mask").view(-1),
                     g_w=A.grad("property_embed.weight").view(-1), g_b=A.grad("property_embed.bias"),
                     g_cls=A.grad("property_cls").view(-1), g_mask=A.grad("property_mask").view(-1))
        W["itm"] = ns(w=A.f32("itm_head.weight"), b=A.f32("itm_head.bias"), g_w=A.grad("itm_head.weight"),
                      g_b=A.grad("itm_head.bias"))
        eps = self.property_mtr_head[2].eps
        W["mtr"] = ns(eps=eps, w0=A.w16("property_mtr_head.0.weight"), b0=A.f32("property_mtr_head.0.bias"),
                      ln_g=A.f32("property_mtr_head.2.weight"), ln_b=A.f32("property_mtr_head.2.bias"),
                      w3=A.f32("property_mtr_head.3.weight"), b3=A.f32("property_mtr_head.3.bias"),
                      g_w0=A.grad("property_mtr_head.0.weight"), g_b0=A.grad("property_mtr_head.0.bias"),
                      g_ln_g=A.grad("property_mtr_head.2.weight"), g_ln_b=A.grad("property_mtr_head.2.bias"),
                      g_w3=A.grad("property_mtr_head.3.weight"), g_b3=A.grad("property_mtr_head.3.bias"))
        self._arena, self._W, self._anchor = A, W, anchor
        return A

    def arena(self):
        if self._arena is None or not self._arena.valid_for(self):
            self.build_arenas(next(self.parameters()).device)
        return self._arena

    # ------------------------------------------------------------------ the hot path
    def forward(self, property_original, text_input_ids, text_attention_mask, alpha=0, mpm_mask=None, neg_idx=None,
                valid_len=None):
        """`alpha` is a python number or a 1-element fp32 CUDA tensor (read by the loss kernels on the device, so one
        captured CUDA graph serves the epoch-0 ramp).  `valid_len` (int32 CUDA scalar, optional): the batch's own
        `padding='longest'` width when `text_input_ids` has been padded further to a graph length bucket."""
        from .xbert import raw_outputs
        with raw_outputs():                         # bf16 activations between the blocks of the hot path
            return self._forward(property_original, text_input_ids, text_attention_mask, alpha, mpm_mask, neg_idx, valid_len)

    @property
    def device(self):                               # LightningModule.device, read by d_smiles2pv.py:34
        return self.arena().device if self._arena is not None else next(self.parameters()).device

    def _forward(self, property_original, text_input_ids, text_attention_mask, alpha=0, mpm_mask=None, neg_idx=None,
                 valid_len=None):
        """Reference SPMM_models.py:79-256.  `mpm_mask` / `neg_idx=(neg_t2i, neg_i2t)` inject the random draws
        (parity tests); otherwise torch.bernoulli on the device and the counter-based sampler are used."""
        A = self.arena()
        A.ensure_grads()
        W = self._W
        with torch.no_grad():
            self.temp.clamp_(0.01, 0.5)
            # _momentum_update (:99,265-269) hoisted to the top: identical result (momentum weights are not read
            # before this point, online weights do not change inside forward) and it refreshes the bf16 shadows
            A.refresh_shadows(ema=True, momentum=self.momentum)
        pv = property_original.float().contiguous()
        B = pv.shape[0]
        H = self.text_encoder.config.hidden_size
        if mpm_mask is None:
            mpm_mask = torch.bernoulli(torch.ones_like(pv) * 0.5)          # :85, same torch generator call
        mpm_mask = mpm_mask.float().contiguous()
        tmask = MaskInfo(text_attention_mask)
        ids = text_input_ids.contiguous()
        te, te_m = self.text_encoder, self.text_encoder_m

        properties = ops.pv_tokens(pv, mpm_mask, W["pv"], self._anchor)                                # :82-88
        # Passes that run the SAME weights on the SAME input are batched into one 2B-row pass (rows are independent; the
        # per-row causal flag lives in the attention kernels): the bidirectional property pass (:90) with the causal one
        # of MPM (:242), and the bidirectional text pass (:94) with the first six - text-only - layers of the causal
        # MLM pass (:224), likewise for the momentum text encoder.  Every GEMM / LayerNorm launch of those layers then
        # carries twice the rows (two waves instead of two single-wave launches) and the weights are read once.
        fl = te.config.fusion_layer
        L = ids.shape[1]
        props2 = torch.cat([properties, properties], dim=0)
        pe2 = self.property_encoder(inputs_embeds=props2, causal_from=B).last_hidden_state
        prop_embeds, pc = torch.split(pe2, B, dim=0)                                                    # :90, :242
        z_prop = ops.proj_f32(prop_embeds[:, 0, :], W["property_proj"])                                 # :92
        ids2 = torch.cat([ids, ids], dim=0)
        tmask2 = MaskInfo(kv_len=torch.cat([tmask.kv_len, tmask.kv_len]))
        te2 = te.bert(ids2, attention_mask=tmask2, mode='text', causal_from=B).last_hidden_state
        text_embeds, mlm_lower = torch.split(te2, B, dim=0)                                             # :94, :224 (layers < fl)
        z_text = ops.proj_f32(text_embeds[:, 0, :], W["text_proj"])                                     # :95
        with torch.no_grad():                                                                           # :98-106
            prop_embeds_m = self.property_encoder_m(inputs_embeds=properties).last_hidden_state
            z_prop_m = ops.proj_f32(prop_embeds_m[:, 0, :], W["property_proj_m"])
            te2_m = te_m.bert(ids2, attention_mask=tmask2, mode='text', causal_from=B).last_hidden_state
            text_embeds_m, mlm_lower_m = te2_m[:B], te2_m[B:]                                           # :104, :215
            z_text_m = ops.proj_f32(text_embeds_m[:, 0, :], W["text_proj_m"])
        side = {}
        if not torch.is_tensor(alpha):
            alpha = float(alpha)
        loss_ita = ops.itc(z_prop, z_text, self.temp, z_prop_m, z_text_m, self.prop_queue_km, self.text_queue_km,
                           alpha, side)                                                                 # :102-131
        nan_flag = side["nan_flag"]        # rank-local here; made world-wide together with the feature gather below

        # ================ ITM (:135-206) ================ #
        if neg_idx is None:
            # per-step variation comes from the device RNG salt (ops.StepRng), so the draw is CUDA-graph safe
            neg_t2i, neg_i2t = ops.sample_negatives(side, self.sampler_seed, 0)                         # :154-178
        else:
            neg_t2i = torch.as_tensor(neg_idx[0], device=pv.device, dtype=torch.int32)
            neg_i2t = torch.as_tensor(neg_idx[1], device=pv.device, dtype=torch.int32)
        prop_neg = ops.gather_rows(prop_embeds, neg_t2i)
        text_neg = ops.gather_rows(text_embeds, neg_i2t)
        # The reference runs the positive pairs (B) and the negative pairs (2B) as four fusion passes (:137-198), then
        # the fusion layers of the causal MLM pass (:224, text queries over property keys) and the causal MPM pass (:245,
        # property queries over text keys) - all through the SAME six fusion layers.  Rows are independent, so they are
        # batched into two passes of 4B rows: [pos | neg-prop | neg-text | causal], the last B rows with the causal flag.
        kv = tmask.kv_len
        tmask4 = MaskInfo(kv_len=torch.cat([kv, kv, kv[neg_i2t.long()], kv]))
        prop_q = torch.cat([prop_embeds, prop_neg, prop_embeds, pc], dim=0)
        # keys/values: every pair reads one of the B distinct encoder states, so their projection runs on B states
        # (not 4B pairs) and the attention kernels follow an index
        own = torch.arange(B, device=pv.device, dtype=torch.int32)
        fo_prop = te.bert(encoder_embeds=prop_q, attention_mask=None, encoder_hidden_states=text_embeds,
                          encoder_index=torch.cat([own, own, neg_i2t, own]), encoder_attention_mask=tmask4,
                          mode='fusion', causal_from=3 * B).last_hidden_state
        out_prop, po = fo_prop[:3 * B, 0, :], fo_prop[3 * B:]                                           # :137-198, :245
        text_q = torch.cat([text_embeds, text_embeds, text_neg, mlm_lower], dim=0)
        fo_text = te.bert(encoder_embeds=text_q, attention_mask=tmask4, encoder_hidden_states=prop_embeds,
                          encoder_index=torch.cat([own, neg_t2i, own, own]), mode='fusion',
                          causal_from=3 * B).last_hidden_state
        out_text, h = fo_text[:3 * B, 0, :], fo_text[3 * B:]                                            # :137-198, :224
        vl = torch.cat([out_prop, out_text], dim=-1)              # rows [0,B) positives, [B,3B) negatives (:199-201)
        loss_itm = ops.itm_loss(vl, W["itm"], B)

        # NaN guard (:132-133): ONE decision for the whole data-parallel world - the gathered queue rows and the reduced
        # gradients are global, so a rank skipping alone would diverge for good.  The flag travels with the features.
        nan_flag = self._dequeue_and_enqueue(side["feat_prop_m"], side["feat_text_m"], nan_flag)         # :208

        # ================ MLM (:210-238) ================ #
        V = te.config.vocab_size
        # the online causal pass already ran above (text-only layers with the text pass, fusion layers with ITM); the
        # momentum one has its text-only layers done
        with torch.no_grad():
            h_m = te_m.bert(encoder_embeds=mlm_lower_m, attention_mask=tmask, encoder_hidden_states=prop_embeds_m,
                            is_decoder=True, mode='fusion').last_hidden_state
            logits_m = ops.lm_logits(h_m.reshape(-1, H), te_m.bert._bundles().head, V, te.logit_ld())
        loss_mlm = ops.lm_head_loss(h.reshape(-1, H), logits_m, ids, te.bert._bundles().head, alpha, V, valid_len)

        # ================ MPM (:240-254) ================ #
        # po: the causal quarter of the batched property-query fusion pass above
        loss_mpm = ops.mtr_head_loss(po.view(-1, H), pv, mpm_mask, W["mtr"])      # already x5 (:256)

        self.last_aux = {"neg_t2i": neg_t2i, "neg_i2t": neg_i2t, "nan_flag": nan_flag, "mpm_mask": mpm_mask,
                         "feat_prop_m": side["feat_prop_m"], "feat_text_m": side["feat_text_m"]}
        # NaN guard (:132-133) without a host sync: zero losses, and the flag disables enqueue + optimiser step
        bad = nan_flag > 0
        zero = torch.zeros((), device=pv.device)
        return tuple(torch.where(bad, zero, l) for l in (loss_mlm, loss_mpm, loss_ita, loss_itm))

    @torch.no_grad()
    def copy_params(self):
        for model_pair in self.model_pairs:
            for param, param_m in zip(model_pair[0].parameters(), model_pair[1].parameters()):
                param_m.data.copy_(param.data)
                param_m.requires_grad = False

    @torch.no_grad()
    def _momentum_update(self):
        """SPMM_models.py:265-269 as one arena kernel (bit-exact fp32)."""
        self.arena().refresh_shadows(ema=True, momentum=self.momentum)

    @torch.no_grad()
    def _dequeue_and_enqueue(self, prop_feat, text_feat, skip_flag=None):
        """SPMM_models.py:271-286.  Rank-local feats are all-gathered (one NCCL call for both modalities)."""
        stacked = torch.stack([prop_feat, text_feat])                       # [2, W*B, E] fp32, rank-major like torch.cat
        if skip_flag is None:
            feats = gather_world_feats(stacked)
        else:
            feats, skip_flag = gather_world_feats(stacked, skip_flag)       # flag -> max over the world
        n = feats.shape[1]
        assert self.queue_size % n == 0                                      # :279
        K.enqueue(self.prop_queue_km, self.text_queue_km, feats[0], feats[1], self.queue_ptr, skip_flag)
        return skip_flag

    # ------------------------------------------------------------------ Lightning-shaped hooks (reference :345-386)
    # With pytorch_lightning installed the class IS a LightningModule and `optimizers()`, `lr_schedulers()`, `log`,
    # `current_epoch`, `global_rank` are Lightning's own; without it (this image) `_StandaloneHooks` supplies them and
    # `trainer.fit` (spmm_b200/trainer.py) plays the Trainer.
    def lr_scheduler_step(self, scheduler, optimizer_idx, metric):       # reference :345-346 (manual optimisation)
        pass

    def _graph_stepper(self, optimizer):
        """The step's CUDA graph runner (trainer.GraphedTrainStep), built on first use; SPMM_EAGER=1 or
        `model.use_cuda_graph = False` keeps the eager launches (debugging)."""
        import os
        from . import trainer
        if not getattr(self, "use_cuda_graph", True) or os.environ.get("SPMM_EAGER") == "1":
            return None
        if not hasattr(optimizer, "prepare_step"):          # a stock torch optimizer cannot be captured (host-side state)
            return None
        st = self.__dict__.get("_stepper")
        if st is None or st.opt is not optimizer:
            st = trainer.GraphedTrainStep(self, optimizer)
            self.__dict__["_stepper"] = st
        return st

    def training_step(self, train_batch, batch_idx):
        """Reference SPMM_models.py:348-380: tokenise, alpha ramp (epoch 0), forward/backward/clip/AdamW - ONE replay of
        the step's CUDA graph (alpha, lr and the batch's padded width are device scalars, so the ramp and
        `padding='longest'` do not multiply graphs) - and the cosine schedule with the reference's warm-up cadence.
        Returns the four losses as ONE device tensor (no host sync; the reference's `loss != 0` NaN check is a device
        flag that skips the optimiser step on every rank)."""
        from . import trainer
        optimizer, scheduler = self.optimizers(), self.lr_schedulers()
        optimizer = getattr(optimizer, "optimizer", optimizer)           # unwrap a LightningOptimizer
        prop, text = train_batch
        dev = self.arena().device
        if isinstance(text, (list, tuple)) and len(text) > 0 and isinstance(text[0], str):      # SMILES strings
            text_input = self.tokenizer(text, padding='longest', truncation=True, max_length=100, return_tensors="pt")
            ids, mask = text_input.input_ids[:, 1:], text_input.attention_mask[:, 1:]
        else:                                      # pre-tokenised (ids, mask) pair
            ids, mask = text
        alpha = self.config['alpha'] if self.current_epoch > 0 else \
            self.config['alpha'] * min(1., batch_idx / max(self.loader_len, 1))
        stepper = self._graph_stepper(optimizer)
        if stepper is not None:
            losses = stepper(prop, ids, mask, alpha).clone()             # static output buffer -> this step's own copy
        else:
            prop, ids, mask = prop.to(dev, non_blocking=True), ids.to(dev, non_blocking=True), mask.to(dev, non_blocking=True)
            losses = torch.stack([l.detach() for l in trainer.train_step(self, optimizer, prop, ids, mask, alpha)])
        if self.global_rank == 0:
            self.log('lr', optimizer.param_groups[0]["lr"], prog_bar=True)
            for name, l in zip(('loss_mlm', 'loss_mpm', 'loss_ita', 'loss_itm'), losses):
                self.log(name, l, prog_bar=True)
        step_size = 100
        warmup_iterations = self.warmup_steps * step_size
        if self.current_epoch > 0 and batch_idx == 0:
            scheduler.step(self.current_epoch + self.warmup_steps)
        elif self.current_epoch == 0 and batch_idx % step_size == 0 and batch_idx <= warmup_iterations:
            scheduler.step(batch_idx // step_size)
        self.training_step_outputs.append(losses)
        return losses

    def on_train_epoch_end(self):                  # reference :382-386
        tmp = torch.stack(self.training_step_outputs[-1000:]).float().mean(dim=0).tolist()
        if self.global_rank == 0:
            print(f'\n mean loss: {tmp[0]:.4f}, {tmp[1]:.4f}, {tmp[2]:.4f}, {tmp[3]:.4f}')
        self.training_step_outputs.clear()
        return tmp

    def configure_optimizers(self):
        from .optim import FusedClipAdamW
        from .scheduler import create_scheduler
        arg_opt = self.config['optimizer']
        optimizer = FusedClipAdamW(self, lr=arg_opt['lr'], weight_decay=arg_opt['weight_decay'], max_norm=5.0)
        scheduler, _ = create_scheduler(AttrDict(self.config['schedular']), optimizer)
        return [optimizer], [scheduler]
