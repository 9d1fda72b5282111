"""End-to-end GPU parity of the FastSpeech2 / Loss drop-ins against the oracle on identical random-init weights
and synthetic inputs (eval mode: dropout parity with torch's Philox stream is infeasible).

Tolerances (north_star): mel_lens / masks bit-exact; mel and postnet outputs within 1e-5 relative in the fp32
path and 1e-2 relative in the bf16 tensor-core path (relative = max |a-b| / max |b|); losses 1e-5 / 5e-3;
gradients: fp32 path 1e-4 relative per parameter, bf16 path cosine >= 0.98 per large parameter and 0.15
relative L2 over the whole flat gradient."""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fs2_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
NAMES = ["mel_post", "postnet_output", "predict_durations", "predict_pitch", "avg_pitch", "predict_energy", "avg_energy"]


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)


def cuda_batch(c):
    g = lambda k: c[k].cuda()
    return g("tokens"), g("speakers"), g("durations"), g("pitch"), g("energy"), g("intensity"), g("mel"), g("mel_len"), g("phon_len")


@pytest.fixture(scope="module")
def oracle64():
    torch.set_num_threads(8)
    return O.build(seed=0, dtype=torch.float64).eval()


def build_model(pkg, oracle, precision):
    m = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision=precision)
    m.load_state_dict({k: v.float() for k, v in oracle.state_dict().items()})
    return m.cuda().eval()


def run_oracle(oracle, c, with_grad=True):
    oracle.zero_grad()
    d = lambda t: t.double()
    preds = oracle(c["tokens"], c["speakers"], c["durations"], d(c["pitch"]), d(c["energy"]), intensity=d(c["intensity"]))
    losses = O.Loss(**O.DEFAULT_LOSS_CONFIG)(preds, (d(c["mel"]), c["durations"], d(c["pitch"]), d(c["energy"]), c["mel_len"], c["phon_len"]), 0)
    if with_grad:
        losses["total_loss"].backward()
    return preds, losses, {k: p.grad.clone() for k, p in oracle.named_parameters()} if with_grad else None


def run_model(pkg, model, c, with_grad=True):
    tokens, speakers, dur, pitch, energy, intensity, mel, mel_len, phon_len = cuda_batch(c)
    preds = model(tokens, speakers, dur, pitch, energy, intensity=intensity)
    losses = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)(preds, (mel, dur, pitch, energy, mel_len, phon_len), 0)
    if with_grad:
        model.zero_grad()
        losses["total_loss"].backward()
    torch.cuda.synchronize()
    return preds, losses


@pytest.mark.parametrize("case", ["docstring_case", "synthetic_b4"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_loss_backward_vs_oracle_and_golden(pkg, lib, oracle64, case, precision):
    g = torch.load(os.path.join(GOLD, case + ".pt"))
    c = g["case"]
    preds_o, losses_o, grads_o = run_oracle(oracle64, c)
    model = build_model(pkg, oracle64, precision)
    preds, losses = run_model(pkg, model, c)
    assert lib.gemm_tc_error_flag() == 0
    # integer artefacts: bit-exact
    assert torch.equal(preds[7], preds_o[7]) and torch.equal(preds[7], g["mel_lens"])
    assert preds[7].dtype == torch.int64 and not preds[7].is_cuda                      # CPU int64, as model.py:440
    tol_out = 1e-5 if precision == "fp32" else 1e-2
    for n, a, b, gold in zip(NAMES, preds[:7], preds_o[:7], g["preds"]):
        assert rel(a, b) <= tol_out, (n, rel(a, b))
        assert rel(a, gold) <= tol_out + 1e-6, (n, "golden", rel(a, gold))             # committed fixture
    assert (preds[0][1, int(preds[7][1]):] == 0).all()                                 # mel_post masked rows are exactly 0
    tol_loss = 1e-5 if precision == "fp32" else 5e-3
    for k, v in g["losses"].items():
        assert abs(float(losses[k]) - v) <= tol_loss * max(1.0, abs(v)), (k, float(losses[k]), v)
    # the same case through the reference's OWN model.py / loss.py (speechbrain leaves = oracle restatements), frozen by
    # tests/golden/make_glue_golden.py: fp64 forward, and the fp32 loss dict the reference's Loss returns
    rg = torch.load(os.path.join(GOLD, "reference_glue.pt"))[case]
    assert torch.equal(preds[7], rg["fwd64"]["mel_lens"])
    for n, a, ref in zip(NAMES, preds[:7], rg["fwd64"]["preds"]):
        assert rel(a, ref) <= tol_out + 1e-6, (n, "reference glue", rel(a, ref))
    for k, v in rg["train"]["losses"].items():
        assert abs(float(losses[k]) - v) <= 2 * tol_loss * max(1.0, abs(v)), (k, "reference glue", float(losses[k]), v)
    # gradients
    flat_a, flat_b = [], []
    for k, p in model.named_parameters():
        ga, gb = p.grad.double().cpu(), grads_o[k]
        flat_a.append(ga.flatten())
        flat_b.append(gb.flatten())
        if precision == "fp32":
            assert rel(ga, gb) <= 1e-4, (k, rel(ga, gb))
        elif gb.numel() >= 4096:
            cos = torch.nn.functional.cosine_similarity(ga.flatten(), gb.flatten(), dim=0).item()
            assert cos >= 0.98, (k, cos)
        assert abs(ga.norm().item() - g["grad_norms"][k]) <= (1e-3 if precision == "fp32" else 0.2) * g["grad_norms"][k] + 1e-7
    fa, fb = torch.cat(flat_a), torch.cat(flat_b)
    rl2 = ((fa - fb).norm() / fb.norm()).item()
    assert rl2 <= (1e-5 if precision == "fp32" else 0.15), rl2


def test_intermediate_stages_fp32(pkg, oracle64):
    """Stage-by-stage: encoder in/out, conditioning, pitch/energy adds, length regulator, decoder (oracle trace)."""
    c = torch.load(os.path.join(GOLD, "synthetic_b4.pt"))["case"]
    oracle64.trace = {}
    run_oracle(oracle64, c, with_grad=False)
    trace_o, oracle64.trace = oracle64.trace, None
    model = build_model(pkg, oracle64, "fp32")
    model.trace = {}
    run_model(pkg, model, c, with_grad=False)
    for k in ["enc_in", "enc_out", "cond", "after_pitch", "after_energy", "dec_in", "dec_out"]:
        assert rel(model.trace[k], trace_o[k]) <= 1e-5, (k, rel(model.trace[k], trace_o[k]))


def test_mask_quirk_q1_affects_outputs_like_the_reference(pkg, oracle64):
    """With B=4, H=2 and ragged lengths the transposed attn_mask changes valid outputs: the CUDA path
    must follow the reference, i.e. differ from an 'intended-mask' model.  Checked by permuting batch order."""
    c = torch.load(os.path.join(GOLD, "synthetic_b4.pt"))["case"]
    model = build_model(pkg, oracle64, "fp32")
    p1, _ = run_model(pkg, model, c, with_grad=False)
    perm = torch.tensor([0, 2, 1, 3])
    c2 = {k: (v[perm] if torch.is_tensor(v) and v.shape[0] == 4 else v) for k, v in c.items()}
    p2, _ = run_model(pkg, model, c2, with_grad=False)
    d = lambda t: t.double()
    o2 = oracle64(c2["tokens"], c2["speakers"], c2["durations"], d(c2["pitch"]), d(c2["energy"]), intensity=d(c2["intensity"]))
    Tm = p2[0].shape[1]
    assert rel(p2[0], o2[0]) <= 1e-5                                 # follows the reference under permutation
    # sample 0 keeps its slot; its output changes only because its mask partner changed (a batch-independent model would not)
    assert (p1[0][0].cpu() - p2[0][0].cpu()).abs().max() > 1e-4


@pytest.mark.parametrize("pace", [0.8, 1.0, 1.2])
def test_inference_path_predicted_durations(pkg, oracle64, pace):
    """durations/pitch/energy = None (inference.py:82): dur = clamp(expm1(pred),0), frames = (pace*dur).long(),
    predicted pitch/energy are embedded; per-utterance intensity prototype broadcast over Tp (inference.py:17-19)."""
    o = O.build(seed=0, dtype=torch.float64).eval()
    with torch.no_grad():
        o.durPred.linear.w.bias.fill_(1.7)           # random init predicts ~0 frames; shift so utterances have length
    B, Tp = 6, 17
    g = torch.Generator().manual_seed(int(pace * 10))
    lens = torch.tensor([17, 15, 12, 9, 6, 5])
    tokens = torch.randint(1, 95, (B, Tp), generator=g)
    for b in range(B):
        tokens[b, lens[b]:] = 0
    speakers = torch.randint(0, 4, (B,), generator=g)
    proto = torch.randn(B, 1, 5, generator=g)
    intensity = proto.expand(B, Tp, 5).contiguous() * (tokens != 0).unsqueeze(-1)
    with torch.no_grad():
        po = o(tokens, speakers, pace=pace, pitch_rate=1.1, energy_rate=0.9, intensity=intensity.double())
    model = build_model(pkg, o, "fp32")
    with torch.no_grad():
        pm = model(tokens.cuda(), speakers.cuda(), pace=pace, pitch_rate=1.1, energy_rate=0.9, intensity=intensity.cuda())
    assert pm[4] is None and pm[6] is None
    assert torch.equal(pm[7], po[7])                                 # predicted frame counts: bit-exact here
    for n, a, b in zip(NAMES, pm[:7], po[:7]):
        if a is not None:
            assert rel(a, b) <= 1e-5, (n, rel(a, b))


def test_train_mode_dropout_runs_and_is_seeded(pkg, oracle64):
    c = torch.load(os.path.join(GOLD, "synthetic_b4.pt"))["case"]
    model = build_model(pkg, oracle64, "bf16").train()
    model.manual_seed(7)
    p1, l1 = run_model(pkg, model, c)
    g1 = model.store.flat_grad.clone()
    model.manual_seed(7)
    p2, l2 = run_model(pkg, model, c)
    assert torch.equal(p1[0], p2[0]) and torch.equal(p1[1], p2[1])                                 # same seed, same masks
    assert abs(float(l1["total_loss"]) - float(l2["total_loss"])) <= 1e-5 * abs(float(l1["total_loss"]))   # atomics order only
    assert torch.isfinite(g1).all() and g1.abs().sum() > 0
    p3, _ = run_model(pkg, model, c)                                                             # next step: new masks
    assert not torch.equal(p1[0], p3[0])
    model.eval()
    pe_, le = run_model(pkg, model, c, with_grad=False)
    # train-mode loss is a noisy version of the eval loss (dropout 0.1/0.5): same order of magnitude
    assert 0.5 < float(l1["total_loss"]) / float(le["total_loss"]) < 2.0


def test_cuda_graph_replay_matches_eager(pkg, oracle64):
    """use_cuda_graphs: eager first step, capture on the second, replay afterwards -- same outputs and gradients
    as the eager path (eval mode; bf16 forward is deterministic, gradients differ only by atomics order)."""
    c = torch.load(os.path.join(GOLD, "synthetic_b4.pt"))["case"]
    eager = build_model(pkg, oracle64, "bf16")
    pe_, le = run_model(pkg, eager, c)
    g_e = eager.store.flat_grad.clone()
    graphed = build_model(pkg, oracle64, "bf16")
    graphed.use_cuda_graphs = True
    for step in range(4):
        pg, lg = run_model(pkg, graphed, c)
        assert torch.equal(pg[0], pe_[0]) and torch.equal(pg[1], pe_[1]) and torch.equal(pg[7], pe_[7])
        assert abs(float(lg["total_loss"]) - float(le["total_loss"])) <= 1e-5 * abs(float(le["total_loss"]))
        rl2 = ((graphed.store.flat_grad - g_e).norm() / g_e.norm()).item()
        assert rl2 <= 1e-4, (step, rl2)
    assert len(graphed._graphs) == 1
    # train mode under replay: the device-side counter gives every step fresh dropout masks
    graphed.train()
    outs = [run_model(pkg, graphed, c)[0][0].clone() for _ in range(4)]
    assert not torch.equal(outs[2], outs[3])


def test_async_mel_lens_has_no_host_sync_and_checks_the_hint(pkg, oracle64):
    """async_mel_lens: Tm comes from pitch.shape[1]; mel_lens arrives in pinned memory (valid after a stream sync);
    a batch whose max(sum(dur)) differs from the hint is reported at the next forward."""
    c = torch.load(os.path.join(GOLD, "synthetic_b4.pt"))["case"]
    ref = build_model(pkg, oracle64, "fp32")
    p_ref, _ = run_model(pkg, ref, c, with_grad=False)
    m = build_model(pkg, oracle64, "fp32")
    m.async_mel_lens = True
    p, _ = run_model(pkg, m, c, with_grad=False)           # run_model synchronises
    assert torch.equal(p[0], p_ref[0]) and torch.equal(p[1], p_ref[1])
    assert p[7].tolist() == p_ref[7].tolist() and p[7].dtype == torch.int64 and not p[7].is_cuda
    bad = dict(c)
    bad["pitch"] = torch.cat([c["pitch"], torch.zeros(4, 3)], 1)
    bad["energy"] = torch.cat([c["energy"], torch.zeros(4, 3)], 1)
    tokens, speakers, dur, pitch, energy, intensity, *_ = cuda_batch(bad)
    m(tokens, speakers, dur, pitch, energy, intensity=intensity)
    torch.cuda.synchronize()
    with pytest.raises(RuntimeError, match="async_mel_lens"):
        run_model(pkg, m, c, with_grad=False)


def test_backward_guard_and_grad_accumulation(pkg, oracle64):
    c = torch.load(os.path.join(GOLD, "docstring_case.pt"))["case"]
    model = build_model(pkg, oracle64, "fp32")
    run_model(pkg, model, c)
    g1 = model.store.flat_grad.clone()
    tokens, speakers, dur, pitch, energy, intensity, mel, mel_len, phon_len = cuda_batch(c)
    preds = model(tokens, speakers, dur, pitch, energy, intensity=intensity)
    losses = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)(preds, (mel, dur, pitch, energy, mel_len, phon_len), 0)
    losses["total_loss"].backward()                     # no zero_grad: accumulates like autograd
    assert torch.allclose(model.store.flat_grad, 2 * g1, rtol=1e-4, atol=1e-7)
    preds = model(tokens, speakers, dur, pitch, energy, intensity=intensity)
    stale = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)(preds, (mel, dur, pitch, energy, mel_len, phon_len), 0)
    model(tokens, speakers, dur, pitch, energy, intensity=intensity)
    with pytest.raises(RuntimeError, match="workspace"):
        stale["total_loss"].backward()


def test_training_step_with_torch_adamw_and_fused_adamw_agree(pkg, oracle64):
    """Drop-in contract of train.py:79-81: optimizer.zero_grad(); loss.backward(); optimizer.step() with a stock
    torch.optim.AdamW over model.parameters(); the fused flat AdamW gives the same update for the same gradients
    (gradients are copied across so that atomics-order noise on near-zero gradients cannot flip Adam's sign)."""
    c = torch.load(os.path.join(GOLD, "docstring_case.pt"))["case"]
    m1, m2 = build_model(pkg, oracle64, "fp32"), build_model(pkg, oracle64, "fp32")
    o1 = torch.optim.AdamW(m1.parameters(), lr=1e-4)
    o2 = pkg.FusedAdamW(m2, lr=1e-4)
    first = None
    for it in range(3):
        tokens, speakers, dur, pitch, energy, intensity, mel, mel_len, phon_len = cuda_batch(c)
        preds = m1(tokens, speakers, dur, pitch, energy, intensity=intensity)
        loss = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)(preds, (mel, dur, pitch, energy, mel_len, phon_len), 0)
        o1.zero_grad()
        loss["total_loss"].backward()
        assert all(p.grad is not None for p in m1.parameters())
        m2.store.ensure_grads()
        m2.store.flat_grad.copy_(m1.store.flat_grad)
        o1.step()
        o2.step()
        first = float(loss["total_loss"]) if first is None else first
    torch.cuda.synchronize()
    assert (m1.store.flat - m2.store.flat).abs().max() <= 1e-6
    assert float(loss["total_loss"]) < first                    # three AdamW steps on one batch reduce its loss


def test_full_size_batch_properties(pkg):
    """BASELINE cfg-3 size (B=32, Tp<=128, Tm<=800), bf16: size-independent properties -- mel_lens == sum(dur),
    masked rows exactly zero, finite loss/gradients, fp32-path and bf16-path agree within the bf16 tolerance."""
    data = importlib.import_module("fine-grained-emotional-control-of-tts_b200.data")
    (batch, intensity), = data.synthetic_batches(32, 1, seed=99)
    tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = [t.cuda() for t in batch[:8]]
    torch.manual_seed(0)
    mb = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision="bf16").cuda().eval()
    mf = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision="fp32")
    mf.load_state_dict(mb.state_dict())
    mf = mf.cuda().eval()
    crit = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)
    outs = {}
    for name, m in (("bf16", mb), ("fp32", mf)):
        preds = m(tokens, speakers, dur, pitch, energy, intensity=intensity.cuda())
        loss = crit(preds, (mel, dur, pitch, energy, out_lens, in_lens), 0)
        loss["total_loss"].backward()
        torch.cuda.synchronize()
        assert torch.equal(preds[7], out_lens.cpu())
        for b in range(32):
            assert (preds[0][b, int(out_lens[b]):] == 0).all()
        assert torch.isfinite(m.store.flat_grad).all()
        outs[name] = (preds, float(loss["total_loss"]), m.store.flat_grad.clone())
    assert rel(outs["bf16"][0][0], outs["fp32"][0][0]) <= 1e-2
    assert rel(outs["bf16"][0][1], outs["fp32"][0][1]) <= 1e-2
    assert abs(outs["bf16"][1] - outs["fp32"][1]) <= 5e-3 * abs(outs["fp32"][1])
    ga, gb = outs["bf16"][2].double(), outs["fp32"][2].double()
    assert ((ga - gb).norm() / gb.norm()).item() <= 0.1


def test_north_star_spelled_call_equals_the_reference_call(pkg):
    """forward_batch(texts, src_lens, mels, durations, pitches, energies, intensity) is the reference call in the
    north_star's argument order: identical outputs, and garbage behind src_lens is ignored."""
    import importlib
    data = importlib.import_module("fine-grained-emotional-control-of-tts_b200.data")
    torch.manual_seed(0)
    m = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision="bf16").cuda().eval()
    (batch, intensity), = data.synthetic_batches(4, 1, seed=3, min_tp=8, max_tp=20, max_frames=120, pool_factor=2)
    tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = [t.cuda() for t in batch[:8]]
    with torch.no_grad():
        a = m(tokens, speakers, dur, pitch, energy, intensity=intensity.cuda())
        dirty = tokens.clone()
        dirty[torch.arange(tokens.shape[1], device="cuda")[None] >= in_lens[:, None]] = 7
        b = m.forward_batch(dirty, in_lens, mel, dur, pitch, energy, intensity.cuda(), speakers=speakers)
    for x, y in zip(a[:7], b[:7]):
        assert torch.equal(x, y)
    assert torch.equal(a[7], b[7])


def test_piecewise_adamw_and_multi_part_backward_equal_the_plain_step(pkg):
    """DataParallelStep's schedule on one GPU -- gradients reported piece by piece (grad_ready_hook after every part of the
    backward), each piece updated by its own FusedAdamW.step(ranges=..., advance=...) on an optimizer stream -- must leave
    exactly the parameters of `loss.backward(); opt.step()`; and any choice of encoder cut points gives the same gradient."""
    import importlib
    data = importlib.import_module("fine-grained-emotional-control-of-tts_b200.data")
    par = importlib.import_module("fine-grained-emotional-control-of-tts_b200.parallel")
    (batch, intensity), = data.synthetic_batches(4, 1, seed=3, min_tp=8, max_tp=20, max_frames=120, pool_factor=2)
    dev_batch = [t.cuda() for t in batch[:8]]
    tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = dev_batch
    crit = pkg.Loss(**pkg.DEFAULT_LOSS_CONFIG)
    flats, grads = [], []
    for mode in ("plain", "trainer", "no_cuts"):
        torch.manual_seed(0)
        m = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4, precision="fp32").cuda().eval()
        if mode == "no_cuts":
            m.enc_grad_splits = ()
        opt = pkg.FusedAdamW(m, lr=1e-3)
        if mode == "trainer":
            tr = par.DataParallelStep(m, crit, opt)
            pieces = []
            orig = tr._piece_ready
            tr._piece_ready = lambda lo, hi: (pieces.append((lo, hi)), orig(lo, hi))[1]
            tr(dev_batch, intensity.cuda())
            assert len(pieces) == 5 and pieces[0][1] == m.store.flat.numel() and pieces[4][0] == 0      # five pieces per step
        else:
            opt.zero_grad()
            preds = m(tokens, speakers, dur, pitch, energy, intensity=intensity.cuda())
            crit(preds, (mel, dur, pitch, energy, out_lens, in_lens), 0)["total_loss"].backward()
            opt.step()
        torch.cuda.synchronize()
        flats.append(m.store.flat.clone())
        grads.append(m.store.flat_grad.clone())
    # same kernels in the same order; only the atomics of the segment sums may order their partial sums differently, and
    # Adam's first step (update = lr * sign(g)) turns a last-ulp difference of a near-zero gradient into a visible one:
    # compare the gradients tightly and the parameters robustly
    gmax = grads[0].abs().max()
    for other in (1, 2):
        assert (grads[0] - grads[other]).abs().max() <= 1e-5 * gmax, other
        bad = ((flats[0] - flats[other]).abs() > 1e-5).float().mean().item()
        assert bad <= 1e-4, (other, bad)
