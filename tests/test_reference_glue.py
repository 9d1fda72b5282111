"""The oracle's FastSpeech2 / Loss glue against the REAL reference glue.

tests/reference_glue.py executes emo_rank_tts/fastspeech2/model.py and loss.py unmodified from /root/reference, with only
the speechbrain leaf modules replaced by the oracle's restatements.  Its outputs on seeded cases are frozen in
tests/golden/reference_glue.pt (tests/golden/make_glue_golden.py), so the first test runs anywhere; the live tests run
where the reference tree is mounted (the build container).  What this pins to the reference's own code: module
construction order and names (state_dict layout and random init), the attn_mask quirk, conditioning, the variance
adaptor order, teacher forcing vs predicted durations with pace, the loss slicing / weighting and its gradients.
What stays a restatement: the speechbrain leaves (checked from first principles in test_oracle_first_principles.py)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fs2_oracle as O  # noqa: E402
import reference_glue as RG  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DUR_BIAS = 1.5                                            # make_glue_golden.py: inference cases


def run_oracle(case, mode):
    model = O.build(seed=0).eval()
    c = case
    if mode == "train":
        preds = model(c["tokens"], c["speakers"], c["durations"], c["pitch"], c["energy"], intensity=c["intensity"])
        losses = O.Loss(**O.DEFAULT_LOSS_CONFIG)(preds, (c["mel"], c["durations"], c["pitch"], c["energy"], c["mel_len"],
                                                           c["phon_len"]), 0)
        losses["total_loss"].backward()
        return dict(preds=[p.detach() if p is not None else None for p in preds[:7]], mel_lens=preds[7],
                    losses={k: float(v.detach()) for k, v in losses.items()},
                    grad_norms={k: p.grad.double().norm().item() for k, p in model.named_parameters()})
    model, d = model.double(), torch.float64
    with torch.no_grad():
        if mode == "fwd64":
            preds = model(c["tokens"], c["speakers"], c["durations"], c["pitch"].to(d), c["energy"].to(d),
                          intensity=c["intensity"].to(d))
        else:
            model.durPred.linear.w.bias.data.fill_(DUR_BIAS)
            preds = model(c["tokens"], c["speakers"], pace=float(mode[5:]), intensity=c["intensity"].to(d))
    return dict(preds=[p.detach() if p is not None else None for p in preds[:7]], mel_lens=preds[7])


def compare(got, want, mode):
    assert got["mel_lens"].tolist() == want["mel_lens"].tolist()               # integer frame counts: exact
    rtol, atol = (2e-5, 2e-5) if mode == "train" else (1e-9, 1e-10)
    for a, b in zip(got["preds"], want["preds"]):
        assert (a is None) == (b is None)
        if a is not None:
            assert a.shape == b.shape and a.dtype == b.dtype
            assert torch.allclose(a, b, rtol=rtol, atol=atol), float((a - b).abs().max())
    if mode == "train":
        assert set(got["losses"]) == set(want["losses"])
        for k, v in want["losses"].items():
            assert abs(got["losses"][k] - v) <= 1e-5 * max(1.0, abs(v)), k
        assert list(got["grad_norms"]) == list(want["grad_norms"])             # same parameters, same order
        for k, v in want["grad_norms"].items():
            assert abs(got["grad_norms"][k] - v) <= 1e-3 * max(v, 1e-6) + 1e-7, k


@pytest.mark.parametrize("case_name", ["docstring_case", "synthetic_b4"])
def test_oracle_glue_matches_frozen_reference_glue(case_name):
    torch.set_num_threads(4)
    gold = torch.load(os.path.join(GOLD, "reference_glue.pt"))[case_name]
    case = torch.load(os.path.join(GOLD, case_name + ".pt"))["case"]
    for mode, want in gold.items():
        compare(run_oracle(case, mode), want, mode)


needs_reference = pytest.mark.skipif(not RG.available(), reason="/root/reference is not mounted on this box")


@needs_reference
def test_reference_constructor_gives_the_oracle_state_dict():
    """model.py:149-276 executed: same keys, same order, same shapes, and -- same seed -- the same initial values, i.e.
    the oracle creates its submodules in the reference's order (SURVEY Appendix B is what the reference really builds)."""
    FS2, _ = RG.load()
    torch.manual_seed(0)
    ref = FS2(**O.DEFAULT_MODEL_CONFIG, n_speakers=4)
    orc = O.build(seed=0)
    rs, os_ = ref.state_dict(), orc.state_dict()
    assert list(rs) == list(os_)
    assert all(torch.equal(rs[k], os_[k]) for k in rs)
    assert sum(p.numel() for p in ref.parameters()) == 85295299
    assert "speechbrain" not in sys.modules                                      # the stubs do not leak


@needs_reference
def test_live_reference_glue_in_training_mode_with_dropout():
    """Dropout on: identical outputs under the same RNG seed mean the same modules are called in the same order on the
    same shapes (every dropout draw lines up), for the model and for the loss."""
    torch.set_num_threads(4)
    FS2, Loss = RG.load()
    case = torch.load(os.path.join(GOLD, "synthetic_b4.pt"))["case"]
    c = case
    outs = []
    for build_model, crit in ((lambda: FS2(**O.DEFAULT_MODEL_CONFIG, n_speakers=4), Loss(**O.DEFAULT_LOSS_CONFIG)),
                              (lambda: O.FastSpeech2(**O.DEFAULT_MODEL_CONFIG, n_speakers=4), O.Loss(**O.DEFAULT_LOSS_CONFIG))):
        torch.manual_seed(0)
        model = build_model().train()
        torch.manual_seed(123)
        preds = model(c["tokens"], c["speakers"], c["durations"], c["pitch"], c["energy"], intensity=c["intensity"])
        losses = crit(preds, (c["mel"], c["durations"], c["pitch"], c["energy"], c["mel_len"], c["phon_len"]), 3)
        losses["total_loss"].backward()
        outs.append((preds, losses, {k: p.grad.clone() for k, p in model.named_parameters()}))
    (pr, lr, gr), (po, lo, go) = outs
    assert pr[7].tolist() == po[7].tolist()
    for a, b in zip(pr[:7], po[:7]):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5)
    assert list(lr) == list(lo)
    for k in lr:
        assert abs(float(lr[k].detach()) - float(lo[k].detach())) <= 1e-5 * max(1.0, abs(float(lr[k].detach())))
    for k in gr:
        scale = max(float(gr[k].abs().max()), 1e-6)
        assert float((gr[k] - go[k]).abs().max()) <= 2e-3 * scale, k


@needs_reference
def test_intensity_segment_mean_against_the_reference_function():
    """oracle.intensity_segment_mean (what fs2_intensity_segment_mean is compared with on the GPU)
    against train.py:16-51 itself, run with a stand-in extractor that returns fixed frame intensities."""
    fn = RG.load_get_intensity_representation()
    batch, frames = RG.intensity_case()
    seen = {}

    def extractor(rank_X, mel_len, emo_ids):
        seen["args"] = (rank_X, mel_len, emo_ids)
        return frames

    ref = fn(extractor, batch, torch.device("cpu"))
    assert seen["args"][0] is batch[10] and seen["args"][1] is batch[7] and seen["args"][2] is batch[11]
    got = O.intensity_segment_mean(frames, batch[2], batch[6], batch[0].shape[1])
    assert ref.shape == got.shape == (5, 23, 5)
    assert torch.equal(ref, got)                               # same torch ops in the same order: bit-identical
    frozen = torch.load(os.path.join(GOLD, "reference_glue.pt"))["intensity_rep"]
    assert torch.equal(ref, frozen)


def test_intensity_segment_mean_against_frozen_reference_output():
    batch, frames = RG.intensity_case()
    frozen = torch.load(os.path.join(GOLD, "reference_glue.pt"))["intensity_rep"]
    assert torch.equal(O.intensity_segment_mean(frames, batch[2], batch[6], batch[0].shape[1]), frozen)


@needs_reference
def test_prototype_binning_procedure_against_the_reference_lines():
    """the list / numpy procedure the GPU test compares fs2_prototype_buckets with
    (tests/prototype_case.restated) against the reference's own statements (rank_model/inference.py:91-110), incl. the
    stable sort on tied scores, NaN for empty buckets and zeros for cells without utterances."""
    import numpy as np
    import prototype_case as PC
    I, length, r, spk, emo, n_spk, n_emo, k = PC.inputs(D=4)               # the reference's feature axis is n_emo wide
    speakers, emotions = [f"s{i}" for i in range(n_spk)], [f"e{i}" for i in range(n_emo)]
    storage = {s: {e: [] for e in emotions} for s in speakers}
    for i in range(I.shape[0]):                                            # inference.py:81-88
        storage[speakers[int(spk[i])]][emotions[int(emo[i])]].append((r.numpy()[i], I.numpy()[i, : int(length[i]), :]))
    ref = RG.run_reference_prototype_binning(storage, speakers, emotions, k, n_emo)
    got = PC.restated(I, length, r, spk, emo, n_spk, n_emo, k)
    assert ref.dtype == got.dtype == np.float32 and ref.shape == got.shape
    assert np.array_equal(ref, got, equal_nan=True)
    assert np.isnan(ref[2, 2]).any() and (ref[:, 3] == 0).all()
