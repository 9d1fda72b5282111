"""DeviceCollate (one staging copy + fs2_collate) against the REAL reference collate's output frozen in
tests/golden/collate_ref.pt (tests/golden/make_collate_golden.py runs TextMelCollateWithAlignment from its source)."""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLD)
import make_collate_golden as MG  # noqa: E402

pytestmark = pytest.mark.gpu


def test_device_collate_is_bit_exact_vs_reference(pkg):
    ref = torch.load(os.path.join(GOLD, "collate_ref.pt"))["out"]
    out = pkg.DeviceCollate("cuda")(MG.samples())
    assert len(out) == len(ref) == 12
    for k, (a, b) in enumerate(zip(out, ref)):
        if torch.is_tensor(b):
            assert a.is_cuda and a.dtype == b.dtype and a.shape == b.shape, (k, a.dtype, b.dtype, a.shape, b.shape)
            assert torch.equal(a.cpu(), b), k                 # copies and zero padding only: bit-exact
        else:
            assert a == b, k


def test_device_collate_feeds_the_model_and_the_extractor(pkg):
    """The 12-tuple is what train.py unpacks (train.py:63-72): run it through get_intensity_representation + forward."""
    batch = pkg.DeviceCollate("cuda")(MG.samples(seed=8, n=4))
    phoneme, spk, phon_len, mel, pitch, energy, dur, mel_len = batch[:8]
    assert mel.shape[1] == int(mel_len.max()) and torch.equal(dur.sum(1), mel_len)
    torch.manual_seed(0)
    ext = pkg.IntensityExtractor(**dict(pkg.DEFAULT_RANK_MODEL_CONFIG, n_encoder_layers=1)).cuda().eval()
    rep = pkg.get_intensity_representation(ext, batch, torch.device("cuda"))
    assert rep.shape == (4, phoneme.shape[1], 5) and torch.isfinite(rep).all()
    model = pkg.FastSpeech2(**pkg.DEFAULT_MODEL_CONFIG, n_speakers=4).cuda().eval()
    with torch.no_grad():
        out = model(phoneme, spk, dur, pitch, energy, intensity=rep)
    assert out[0].shape == mel.shape and torch.equal(out[7], mel_len.cpu())


def test_back_to_back_collates_do_not_overwrite_each_others_staging(pkg):
    """the H2D staging copy is asynchronous, so a second call must not refill pinned bytes whose DMA is still
    queued behind running kernels.  Keep the stream busy, collate six different batches back to back, then compare every one
    with a collate done in isolation."""
    batches = [MG.samples(seed=20 + i, n=4) for i in range(6)]
    solo = []
    for b in batches:
        solo.append(pkg.DeviceCollate("cuda")(b))
        torch.cuda.synchronize()
    coll = pkg.DeviceCollate("cuda", n_staging=2)
    a = torch.randn(8192, 8192, device="cuda")
    for _ in range(20):
        a = (a @ a) * 1e-4                          # ~100 ms of queued work in front of the staging copies
    outs = [coll(b) for b in batches]
    torch.cuda.synchronize()
    for o, s in zip(outs, solo):
        for x, y in zip(o, s):
            if torch.is_tensor(x):
                assert torch.equal(x, y)
