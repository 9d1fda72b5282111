"""Host-side logic that needs no GPU: flat parameter store / state_dict contract, synthetic batch recipe,
data-parallel sharding and the flat-gradient all-reduce over gloo (world_size 2)."""
import importlib
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
PKG = "fine-grained-emotional-control-of-tts_b200"


def test_state_dict_interchanges_with_oracle(pkg):
    import fs2_oracle as O
    m = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4)
    o = O.build(seed=1)
    sd_o, sd_m = o.state_dict(), m.state_dict()
    assert list(sd_o.keys()) == sorted(sd_o.keys(), key=list(sd_o.keys()).index)
    assert set(sd_o) == set(sd_m)
    assert all(sd_o[k].shape == sd_m[k].shape for k in sd_o)
    m.load_state_dict(sd_o)
    assert m.store.views_intact()
    for k in ("decoder.layers.2.pos_ffn.0.conv.weight", "concat_proj.w.weight", "postnet.ln3.bias"):
        assert torch.equal(m.state_dict()[k], sd_o[k])
    o2 = O.build(seed=2)
    o2.load_state_dict(m.state_dict())         # and back: checkpoints interchange (train.py:253 <-> inference.py:27)
    assert torch.equal(o2.state_dict()["durPred.ln1.norm.weight"], sd_o["durPred.ln1.norm.weight"])
    assert sum(p.numel() for p in m.parameters()) == 85295299


def test_postnet_ln_alt_keys(pkg):
    m = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4)
    sd = m.state_dict()
    for i in (1, 2, 3):
        for leaf in ("weight", "bias"):
            sd[f"postnet.ln{i}.norm.{leaf}"] = sd.pop(f"postnet.ln{i}.{leaf}") + 1.0
    m.load_state_dict(sd)
    assert torch.allclose(m.state_dict()["postnet.ln2.weight"], torch.full((512,), 2.0))


def test_flat_views_survive_module_apply(pkg):
    m = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4)
    w = m.state_dict()["linear.w.weight"].clone()
    m = m.to("cpu").float()
    assert m.store.views_intact()
    assert torch.equal(m.state_dict()["linear.w.weight"], w)
    with pytest.raises(RuntimeError):
        m.double()


def test_constructor_validation(pkg):
    cfg = dict(pkg.DEFAULT_MODEL_CONFIG)
    with pytest.raises(NotImplementedError):
        pkg.FastSpeech2(**{**cfg, "ffn_type": "regularFFN"}, n_speakers=4)
    with pytest.raises(ValueError):
        pkg.FastSpeech2(**{**cfg, "ffn_cnn_kernel_size_list": [11, 1]}, n_speakers=4)
    with pytest.raises(NotImplementedError):
        pkg.Loss(**{**pkg.DEFAULT_LOSS_CONFIG, "log_scale_durations": False})


def test_synthetic_batch_recipe():
    data = importlib.import_module(PKG + ".data")
    batches = data.synthetic_batches(8, 3, seed=1234, rank=0)
    for batch, intensity in batches:
        tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = batch[:8]
        B, Tp = tokens.shape
        assert B == 8 and Tp <= 128 and mel.shape[1] <= 800 and mel.shape[2] == 80
        assert (in_lens[:-1] >= in_lens[1:]).all()                       # sorted by descending Tp (dataset.py:65-67)
        assert torch.equal(dur.sum(1), out_lens)                          # mel_len == sum(dur)
        assert int(out_lens.max()) == mel.shape[1]
        for b in range(B):
            assert (tokens[b, : in_lens[b]] > 0).all() and (tokens[b, in_lens[b]:] == 0).all()
            assert (mel[b, out_lens[b]:] == 0).all() and (intensity[b, in_lens[b]:] == 0).all()
        assert intensity.shape == (B, Tp, 5)
        assert batch[10].shape == (B, 82, mel.shape[1])
    other = data.synthetic_batches(8, 1, seed=1234, rank=1)
    assert not torch.equal(other[0][0][0], batches[0][0][0]) or other[0][0][0].shape != batches[0][0][0].shape


def test_data_parallel_sharding_keeps_ranks_on_similar_rectangles():
    """the ranks of one step draw neighbouring length buckets of one shared pool, so their padded
    rectangles (= step cost) are close and the per-rank mix of short and long batches matches the 1-GPU stream."""
    data = importlib.import_module(PKG + ".data")
    world = 4
    per_rank = [data.synthetic_batches(8, 4, seed=1234, rank=r, world=world) for r in range(world)]
    for step in range(4):
        tms = [per_rank[r][step][0][3].shape[1] for r in range(world)]
        assert tms == sorted(tms)                                  # consecutive groups of the sorted pool
        assert max(tms) - min(tms) <= 0.35 * max(tms) + 40
    firsts = [per_rank[r][0][0][0] for r in range(world)]
    assert any(firsts[0].shape != f.shape or not torch.equal(firsts[0], f) for f in firsts[1:])   # different utterances
    again = data.synthetic_batches(8, 4, seed=1234, rank=2, world=world)
    assert all(torch.equal(a[0][3], b[0][3]) for a, b in zip(again, per_rank[2]))                # deterministic


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    par = importlib.import_module(PKG + ".parallel")
    r, w, _ = par.init_from_env("gloo")
    flat = torch.full((1000,), float(rank + 1))
    par.allreduce_flat_(flat, w)
    p = torch.arange(10.0) * (rank + 1)
    par.broadcast_flat_(p, 0)
    shard = par.shard_for_rank(list(range(11)), r, w)
    q.put((rank, flat[0].item(), p.tolist(), shard))
    dist.destroy_process_group()


def test_dp_allreduce_and_sharding_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 3.0 and res[1][1] == 3.0                       # SUM over ranks of the flat buffer
    assert res[0][2] == res[1][2] == [float(i) for i in range(10)]      # replicas identical after broadcast
    assert res[0][3] == [0, 2, 4, 6, 8] and res[1][3] == [1, 3, 5, 7, 9]


def test_dgrad_split_rule_follows_the_persistent_grid():
    """FastSpeech2._dgrad_split mirrors fs2_gemm_tc's kernel choice for the k = 9 FFN dgrad: split the reduction in two
    where two half-length rounds of the persistent grid beat one (148 CTAs / 74 CTA pairs)."""
    pkg = importlib.import_module(PKG)
    m = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision="bf16")
    rows = lambda B, T: B * (T + 8)
    assert m._dgrad_split(rows(32, 128), 384) == 2       # 68 single-CTA tiles on 148 SMs: one half-length round
    assert m._dgrad_split(rows(32, 376), 384) == 1       # 48 pair tiles fit one round of 74 pairs: nothing to gain
    assert m._dgrad_split(rows(32, 632), 384) == 2       # 80 pair tiles = 2 rounds -> 160 items = 3 half rounds
    assert m._dgrad_split(rows(32, 800), 384) == 2
    assert m._dgrad_split(rows(256, 895), 384) == 1      # long grids: quantisation is already below 15 %
    assert m._dgrad_split(rows(32, 800), 512) == 1       # only the 384-wide FFN dgrad has the two-buffer consumer
    assert pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision="fp32")._dgrad_split(rows(32, 128), 384) == 1


def test_upstream_drop_ins_refuse_cpu_inputs():
    """8f rows: the extractor, the segment mean and the device collate have no CPU path either."""
    pkg = importlib.import_module(PKG)
    ext = pkg.IntensityExtractor(**dict(pkg.DEFAULT_RANK_MODEL_CONFIG, n_encoder_layers=1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ext(torch.zeros(2, 30, 82), torch.tensor([30, 10]), torch.tensor([0, 1]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.intensity_segment_mean(torch.zeros(2, 30, 5), torch.ones(2, 6, dtype=torch.long), torch.tensor([6, 6]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.DeviceCollate("cpu")([])
    with pytest.raises(NotImplementedError):
        pkg.IntensityExtractor(n_mels=80, n_heads=4, n_emotions=5, n_encoder_layers=1, hidden_dim=384, kernel_size=9, dropout=0.1)
