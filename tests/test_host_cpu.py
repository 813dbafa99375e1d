"""CPU-side checks (run with -m "not gpu"): C-ABI surface, host logic of the reference-facing mirror, sampler
replica, data-parallel plumbing over gloo with world_size 2.  No kernel is launched here."""
import ctypes
import json
import math
import os
import re

import numpy as np
import pytest
import torch

from tests.util import CFG, REPO, load_golden


def test_abi_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from spmm_b200 import _lib
    header = open(os.path.join(REPO, "include", "spmm_b200.h")).read()
    declared = set(re.findall(r"\b(spmm_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 28
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, "ctypes binding missing for " + name
    assert _lib.lib().spmm_version() == 100
    # the epilogue struct the Python side passes must match the header field for field
    fields = re.search(r"typedef struct spmm_gemm_epilogue \{(.*?)\} spmm_gemm_epilogue;", header, re.S).group(1)
    names = re.findall(r"(\w+);", re.sub(r"/\*.*?\*/", "", fields, flags=re.S))
    assert names == [f[0] for f in _lib.GemmEpilogue._fields_]


def test_product_path_fails_loudly_without_cuda():
    from spmm_b200 import _lib, synth
    from spmm_b200.SPMM_models import SPMM
    cfg = synth.pretrain_config(os.path.join(CFG, "config_tiny_text.json"), os.path.join(CFG, "config_tiny_property.json"), 96, 6)
    model = SPMM(config=cfg)
    pv, ids, mask, _ = synth.synthetic_batch(6, seed=1)
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SpmmKernelError):
            model(pv, ids, mask, alpha=0.4)
    assert "oracle" not in open(os.path.join(REPO, "spmm_b200", "SPMM_models.py")).read().replace("oracle/", "")
    for f in os.listdir(os.path.join(REPO, "spmm_b200")):
        if f.endswith(".py"):
            src = open(os.path.join(REPO, "spmm_b200", f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_state_dict_surface_matches_reference():
    from spmm_b200 import synth
    from spmm_b200.SPMM_models import SPMM
    g = load_golden("tiny_b6")
    cfg = synth.pretrain_config(os.path.join(CFG, "config_tiny_text.json"), os.path.join(CFG, "config_tiny_property.json"), 96, 6)
    model = SPMM(config=cfg)
    sd = model.state_dict()
    assert sorted((k, tuple(v.shape), str(v.dtype)) for k, v in sd.items()) == sorted(g["state_dict_keys"])
    assert sum(p.numel() for p in model.parameters()) == g["n_all"]
    assert sum(p.numel() for p in model.parameters() if p.requires_grad) == g["n_trainable"]
    te = model.text_encoder
    assert te.cls.predictions.decoder.weight is te.bert.embeddings.word_embeddings.weight
    assert te.cls.predictions.decoder.bias is te.cls.predictions.bias
    # queues: [E, Q] unit-norm columns like SPMM_models.py:72-77; round trip through load_state_dict
    assert torch.allclose(model.prop_queue.norm(dim=0), torch.ones(96), atol=1e-5)
    synth.fill_by_name(model)
    sd2 = model.state_dict()
    assert torch.equal(sd2["prop_queue"], synth.value_for("prop_queue", (256, 96)))
    assert torch.equal(sd2["text_encoder_m.bert.encoder.layer.3.crossattention.self.key.weight"],
                       synth.value_for("text_encoder_m.bert.encoder.layer.3.crossattention.self.key.weight", (128, 128)))
    # no_train=True (d_smiles2pv.py:130): no temp / queues
    m2 = SPMM(config=cfg, no_train=True)
    assert "temp" not in m2.state_dict() and "prop_queue" not in m2.state_dict()
    missing = m2.load_state_dict({k: v for k, v in sd2.items() if "queue" not in k}, strict=False)
    assert missing.missing_keys == []


def test_bert_config_reads_reference_json_unchanged():
    from spmm_b200.xbert import BertConfig
    c = BertConfig.from_json_file(os.path.join(CFG, "config_bert.json"))
    assert c.add_cross_attention is True and c.fusion_layer == 6 and c.num_hidden_layers == 12 and c.vocab_size == 300
    p = BertConfig.from_json_file(os.path.join(CFG, "config_bert_property.json"))
    assert p.num_hidden_layers == 6 and p.vocab_size == 1 and p.fusion_layer == 6
    ref = "/root/reference/config_bert.json"
    if os.path.exists(ref):
        assert json.load(open(ref)) == json.load(open(os.path.join(CFG, "config_bert.json")))


def test_arena_qkv_ordering():
    from spmm_b200.arena import _pad, _qkv_order
    names = ["a.self.query.weight", "a.self.query.bias", "a.self.key.weight", "a.self.key.bias", "a.self.value.weight",
             "a.self.value.bias", "a.output.dense.weight"]
    assert _qkv_order(names) == ["a.self.query.weight", "a.self.key.weight", "a.self.value.weight", "a.self.query.bias",
                                 "a.self.key.bias", "a.self.value.bias", "a.output.dense.weight"]
    assert _pad(300) == 320 and _pad(768 * 768) == 768 * 768


def test_cosine_schedule_matches_reference_formula():
    from spmm_b200.scheduler import create_scheduler
    from spmm_b200.SPMM_models import AttrDict
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], lr=5e-5)
    args = AttrDict({'sched': 'cosine', 'lr': 5e-5, 'epochs': 30, 'min_lr': 1e-5, 'decay_rate': 1, 'warmup_lr': 5e-5,
                     'warmup_epochs': 20, 'cooldown_epochs': 0})
    s, n = create_scheduler(args, opt)
    assert n == 30
    for t in (0, 5, 19):
        s.step(t)
        assert opt.param_groups[0]["lr"] == pytest.approx(5e-5)
    for t in (20, 25, 35, 49):
        s.step(t)
        want = 1e-5 + 0.5 * (5e-5 - 1e-5) * (1 + math.cos(math.pi * (t - 20) / 30))
        assert opt.param_groups[0]["lr"] == pytest.approx(want)
    s.step(50)
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-5)
    if os.path.isdir("/root/reference/scheduler"):
        import sys
        sys.path.insert(0, "/root/reference")
        from scheduler import create_scheduler as ref_create
        opt2 = torch.optim.AdamW([p], lr=5e-5)
        rs, _ = ref_create(args, opt2)
        for t in range(0, 55, 3):
            s.step(t); rs.step(t)
            assert opt.param_groups[0]["lr"] == pytest.approx(opt2.param_groups[0]["lr"], rel=1e-12)


def test_sampler_replica_distribution_and_edge_cases():
    from oracle import sampler_ref
    rng = np.random.default_rng(0)
    sim = (rng.normal(size=(6, 6)) * 2).astype(np.float32)
    w = np.exp(sim[2] - sim[2].max()); w[2] = 0; w /= w.sum()
    counts = np.zeros(6)
    for step in range(6000):
        counts[sampler_ref.sample_row(sim[2], 2, 0, 1234, step)] += 1
    assert counts[2] == 0
    assert np.abs(counts / 6000 - w).max() < 0.02
    # B == 2: the only admissible negative is the other row (multinomial needs B >= 2)
    s2 = np.zeros((2, 2), dtype=np.float32)
    assert sampler_ref.sample_negatives(s2, s2, 1, 1) == ([1, 0], [1, 0])
    # extreme logits: everything but one candidate underflows
    s3 = np.array([[0, -200, 50, -200]], dtype=np.float32)
    assert sampler_ref.sample_row(s3[0], 0, 1, 5, 5) == 2
    xs = np.linspace(-80, 0, 400).astype(np.float32)
    rel = max(abs(float(sampler_ref.exact_exp_neg(x)) - math.exp(float(x))) / math.exp(float(x)) for x in xs)
    assert rel < 5e-7


def _dist_worker(rank, world, path, out):
    import torch.distributed as dist
    from spmm_b200.SPMM_models import gather_world_feats
    dist.init_process_group("gloo", init_method="file://" + path, rank=rank, world_size=world)
    B, E = 3, 4
    feats = torch.arange(2 * B * E, dtype=torch.float32).reshape(2, B, E) + 100 * rank
    g = gather_world_feats(feats)
    # gradient all-reduce + 1/world scale as in trainer.train_step
    grad = torch.full((5,), float(rank + 1))
    dist.all_reduce(grad, op=dist.ReduceOp.SUM)
    # the NaN guard is one decision for the whole world: a flag raised on rank 1 only must be seen by rank 0
    from spmm_b200.SPMM_models import world_any
    flag = world_any(torch.tensor(1.0 if rank == 1 else 0.0))
    g2, flag2 = gather_world_feats(feats, torch.tensor(1.0 if rank == 1 else 0.0))   # flag riding on the feature gather
    assert torch.equal(g2, g) and float(flag2) == 1.0
    if rank == 0:
        torch.save({"g": g, "grad": grad / world, "flag": flag}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_plumbing_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "r0.pt")
    mp.spawn(_dist_worker, args=(2, str(tmp_path / "pg"), out), nprocs=2, join=True)
    r = torch.load(out)
    B, E = 3, 4
    base = torch.arange(2 * B * E, dtype=torch.float32).reshape(2, B, E)
    want = torch.cat([base, base + 100], dim=1)             # torch.cat(tensors_gather, dim=0) per modality
    assert torch.equal(r["g"], want)
    assert torch.equal(r["grad"], torch.full((5,), 1.5))
    assert float(r["flag"]) == 1.0
    # queue divisibility rule of the reference (SPMM_models.py:279) for W in {1,2,4,8} at B=96
    assert all(36864 % (96 * w) == 0 for w in (1, 2, 4, 8))


def test_lightning_shaped_hooks_follow_the_reference_schedule(monkeypatch):
    """training_step / on_train_epoch_end (reference SPMM_models.py:348-386) without Lightning: alpha ramp over epoch 0,
    scheduler stepping every 100 batches of the warm-up and once per later epoch, epoch mean of the last 1000 steps.
    The fused device step itself is stubbed (no GPU here)."""
    from types import SimpleNamespace
    from spmm_b200 import synth, trainer
    from spmm_b200.SPMM_models import SPMM
    cfg = synth.pretrain_config(os.path.join(CFG, "config_tiny_text.json"), os.path.join(CFG, "config_tiny_property.json"), 96, 6)
    cfg["schedular"]["warmup_epochs"] = 2
    model = SPMM(config=cfg, loader_len=400)
    seen = {"alpha": [], "sched": []}

    def fake_step(m, opt, prop, ids, mask, alpha, **kw):
        seen["alpha"].append(alpha)
        n = float(len(seen["alpha"]))
        return [torch.tensor(n), torch.tensor(2 * n), torch.tensor(3 * n), torch.tensor(4 * n)]
    monkeypatch.setattr(trainer, "train_step", fake_step)
    monkeypatch.setattr(SPMM, "arena", lambda self: SimpleNamespace(device=torch.device("cpu")))
    opt = SimpleNamespace(param_groups=[{"lr": 1e-4}])
    sched = SimpleNamespace(step=lambda e: seen["sched"].append(e))
    model.attach(opt, sched, global_rank=0, log=None)
    batch = (torch.zeros(6, 53), (torch.ones(6, 12, dtype=torch.long), torch.ones(6, 12, dtype=torch.long)))
    model.current_epoch = 0
    for i in (0, 50, 100, 200, 300, 399):
        out = model.training_step(batch, i)
        assert out.shape == (4,)
    a = cfg["alpha"]
    assert seen["alpha"] == pytest.approx([0.0, a * 50 / 400, a * 100 / 400, a * 200 / 400, a * 300 / 400, a * 399 / 400])
    assert seen["sched"] == [0, 1, 2]                    # batches 0, 100, 200 (<= warmup_epochs * 100); 300 is past the warm-up
    mean = model.on_train_epoch_end()
    assert mean == pytest.approx([3.5, 7.0, 10.5, 14.0]) and model.training_step_outputs == []
    model.current_epoch = 3
    model.training_step(batch, 0)
    model.training_step(batch, 100)
    assert seen["alpha"][-2:] == [a, a]
    assert seen["sched"] == [0, 1, 2, 3 + 2]             # one step per epoch afterwards: epoch + warmup_steps


def test_lightning_checkpoint_round_trip(tmp_path):
    """`.ckpt` files carry the reference's state-dict surface (SPMM_pretrain.py:24-30, d_smiles2pv.py:132-143): a model
    written and re-read is identical, queues keep the reference's [E, Q] layout on disk, and the d_*.py loading pattern
    (queues dropped, no_train=True model, strict=False) reports no unexpected keys."""
    from spmm_b200 import checkpoint, synth
    from spmm_b200.SPMM_models import SPMM
    g = load_golden("tiny_b6")
    cfg = synth.pretrain_config(os.path.join(CFG, "config_tiny_text.json"), os.path.join(CFG, "config_tiny_property.json"), 96, 6)
    torch.manual_seed(3)
    m1 = SPMM(config=cfg)
    m1.queue_ptr.fill_(12)
    path = checkpoint.save(m1, str(tmp_path), epoch=4, global_step=1234)
    assert os.path.basename(path) == "checkpoint_epoch=4.ckpt"
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert raw["epoch"] == 4 and raw["global_step"] == 1234
    assert sorted((k, tuple(v.shape)) for k, v in raw["state_dict"].items()) == sorted((k, s) for k, s, _ in g["state_dict_keys"])
    assert tuple(raw["state_dict"]["prop_queue"].shape) == (cfg["embed_dim"], 96)
    torch.manual_seed(4)
    m2 = SPMM(config=cfg)
    msg, _ = checkpoint.load(m2, path)
    assert not msg.missing_keys and not msg.unexpected_keys
    s1, s2 = m1.state_dict(), m2.state_dict()
    assert all(torch.equal(s1[k], s2[k]) for k in s1) and int(m2.queue_ptr) == 12
    assert torch.equal(m2.prop_queue_km, m1.prop_queue_km)               # key-major storage restored from the [E, Q] view
    m3 = SPMM(config=cfg, no_train=True)                                  # inference scripts
    msg3, _ = checkpoint.load(m3, path, drop_queues=True)
    assert not msg3.missing_keys and set(msg3.unexpected_keys) <= {"temp", "queue_ptr"}


def test_attention_masks_must_be_right_padded():
    """The kernels consume prefix lengths; the reference honours arbitrary masks (xbert.py:889-948).  A mask with a hole
    or left padding is rejected where validation runs (inference / SPMM_CHECK_INPUTS=1) instead of masking the wrong keys."""
    from spmm_b200.xbert import MaskInfo
    ok = torch.tensor([[1, 1, 1, 0], [1, 0, 0, 0], [1, 1, 1, 1]])
    with torch.no_grad():
        assert MaskInfo(ok).kv_len.tolist() == [3, 1, 4]
        for bad in (torch.tensor([[0, 1, 1, 1]]), torch.tensor([[1, 0, 1, 0]])):
            with pytest.raises(ValueError):
                MaskInfo(bad)
    assert MaskInfo(kv_len=torch.tensor([2, 3])).kv_len.tolist() == [2, 3]
    with pytest.raises(ValueError):
        MaskInfo(torch.ones(2, 3, 4))


def test_graph_stepper_buckets_lengths_and_keeps_lru_order():
    """GraphedTrainStep keys its graphs on (batch, bucketed length, injected-mask flag) - never on alpha or lr, which
    are device scalars - and evicts least-recently-used graphs.  Host logic only (capture is stubbed)."""
    from types import SimpleNamespace
    from spmm_b200 import trainer
    g = trainer.GraphedTrainStep.__new__(trainer.GraphedTrainStep)
    import collections
    g.graphs, g.max_graphs, g.len_bucket = collections.OrderedDict(), 2, 8
    assert [g.bucket_len(L) for L in (1, 8, 9, 63, 64, 65, 99)] == [8, 8, 16, 64, 64, 72, 104]
    captured, filled = [], []

    def fake_capture(key, *a):
        captured.append(key)
        while len(g.graphs) >= g.max_graphs:
            g.graphs.popitem(last=False)
        g.graphs[key] = {"graph": SimpleNamespace(replay=lambda: None)}
        return g.graphs[key]
    g._capture = fake_capture
    g._fill = lambda st, *a: filled.append(a[3])
    g.model = SimpleNamespace(arena=lambda: SimpleNamespace(device="cpu"), training=True)
    g.opt = SimpleNamespace(prepare_step=lambda: None)
    g.losses = torch.zeros(4)
    import spmm_b200.ops as ops
    orig = ops.step_rng
    ops.step_rng = lambda dev: SimpleNamespace(host=0)
    try:
        pv = torch.zeros(6, 53)
        for L, alpha in ((60, 0.0), (64, 0.1), (57, 0.2), (70, 0.3), (62, 0.4), (99, 0.4), (70, 0.4)):
            g(pv, torch.zeros(6, L, dtype=torch.long), torch.zeros(6, L, dtype=torch.long), alpha)
    finally:
        ops.step_rng = orig
    # 60/64/57/62 share the 64-bucket; 70 -> 72; 99 -> 104 evicts the least recently used (72: the 64-bucket was just
    # replayed for L=62); the second L=70 batch captures 72 again and evicts 64
    assert captured == [(6, 64, False, True), (6, 72, False, True), (6, 104, False, True), (6, 72, False, True)]
    assert list(g.graphs) == [(6, 104, False, True), (6, 72, False, True)]
    assert filled == [0.1, 0.2, 0.4]                   # replays refresh the device scalars; alpha never keys a graph
