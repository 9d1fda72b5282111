"""SURVEY 8f row 2 -- the intensity extractor.  CPU: the oracle restatement (oracle/rank_oracle.py) against the golden
fixture produced by the REAL reference class (tests/golden/make_rank_golden.py), and against the live reference when
/root/reference is mounted.  GPU: the B200 forward against both."""
import importlib
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, GOLD)
import rank_oracle as RO  # noqa: E402
import rank_weights as RW  # noqa: E402

REF = "/root/reference/emo_rank_tts/rank_model/model.py"


def _golden():
    return torch.load(os.path.join(GOLD, "rank_extractor.pt"))


def test_restatement_matches_the_reference_golden():
    g = _golden()
    out = RO.intensity_extractor_forward(RW.state_dict(), g["x"], g["length"], g["emotions"])
    assert out.shape == g["out"].shape == (3, 45, 5)
    assert (out - g["out"]).abs().max() <= 2e-5 * max(1.0, float(g["out"].abs().max()))     # fp32 vs fp32, op order differs
    # rows beyond an utterance's length carry only the classifier bias (masked_fill before the classifier, model.py:107)
    sd = RW.state_dict()
    assert torch.allclose(g["out"][2, 12:], sd["classifier.bias"].expand(33, 5), atol=1e-6)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
def test_restatement_matches_the_live_reference_on_other_shapes():
    spec = importlib.util.spec_from_file_location("ref_rank_model", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    cfg = dict(RW.CFG, n_encoder_layers=2)
    model = ref.IntensityExtractor(**cfg).eval()
    sd = {k: v for k, v in RW.state_dict(cfg).items()}
    model.load_state_dict(sd, strict=True)
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(4, 23, 82, generator=gen)
    length = torch.tensor([23, 23, 9, 1])
    emotions = torch.tensor([1, 1, 2, 0])
    with torch.no_grad():
        want = model(x, length, emotions)
    got = RO.intensity_extractor_forward(sd, x, length, emotions)
    assert (got - want).abs().max() <= 2e-5 * max(1.0, float(want.abs().max()))


def test_state_dict_layout_is_the_reference_layout():
    pkg = importlib.import_module("fine-grained-emotional-control-of-tts_b200.rank_model")
    m = pkg.IntensityExtractor(**RW.CFG)
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert mine == RW.shapes()                      # RW.shapes() is what the real reference accepted with strict=True
    m.load_state_dict(RW.state_dict(), strict=True)


@pytest.mark.gpu
@pytest.mark.parametrize("channels_first", [False, True])
def test_b200_forward_matches_reference_golden(channels_first):
    pkg = importlib.import_module("fine-grained-emotional-control-of-tts_b200.rank_model")
    g = _golden()
    m = pkg.IntensityExtractor(**RW.CFG)
    m.load_state_dict(RW.state_dict(), strict=True)
    m = m.cuda().eval()
    x = g["x"].cuda()
    if channels_first:                              # the layout the FastSpeech2 collate produces (dataset.py:94, 116)
        x = x.transpose(1, 2).contiguous()
    out = m(x, g["length"].cuda(), g["emotions"].cuda())
    ref = g["out"]
    err = (out.cpu() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 2e-2, err                         # bf16 operands, fp32 accumulation, 6 blocks
    # masked frames: exactly the classifier bias
    assert torch.allclose(out.cpu()[2, 12:], RW.state_dict()["classifier.bias"].expand(33, 5), atol=1e-6)


@pytest.mark.gpu
def test_b200_forward_layout_is_explicit_when_ambiguous():
    """A batch padded to exactly n_mels + 2 = 82 frames: both axes have the feature size.  The layout is then taken from
    `channels_first`, never guessed (get_intensity_representation passes True for the collate's rank_X)."""
    pkg = importlib.import_module("fine-grained-emotional-control-of-tts_b200.rank_model")
    cfg = dict(RW.CFG, n_encoder_layers=1)
    m = pkg.IntensityExtractor(**cfg)
    m.load_state_dict(RW.state_dict(cfg), strict=True)
    m = m.cuda().eval()
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 82, 82, generator=gen).cuda()                 # (B, T, C) with T == C
    length, emo = torch.tensor([82, 40]).cuda(), torch.tensor([1, 3]).cuda()
    with pytest.raises(ValueError, match="channels_first"):
        m(x, length, emo)
    a = m(x, length, emo, channels_first=False)
    b = m(x.transpose(1, 2).contiguous(), length, emo, channels_first=True)
    assert torch.equal(a, b)
    c = m(x, length, emo, channels_first=True)                       # the other reading of the same bytes differs
    assert not torch.allclose(a, c)


@pytest.mark.gpu
def test_b200_forward_ragged_batch_vs_oracle():
    pkg = importlib.import_module("fine-grained-emotional-control-of-tts_b200.rank_model")
    cfg = dict(RW.CFG, n_encoder_layers=2)
    sd = RW.state_dict(cfg)
    m = pkg.IntensityExtractor(**cfg)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    gen = torch.Generator().manual_seed(11)
    B, T = 5, 130
    length = torch.tensor([130, 129, 64, 7, 1])
    x = torch.randn(B, T, 82, generator=gen)
    emotions = torch.tensor([4, 0, 2, 2, 1])
    want = RO.intensity_extractor_forward({k: v.double() for k, v in sd.items()}, x.double(), length, emotions)
    got = m(x.cuda(), length.cuda(), emotions.cuda()).cpu().double()
    err = (got - want).abs().max().item() / want.abs().max().item()
    assert err <= 2e-2, err
