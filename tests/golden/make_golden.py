"""Generates tests/golden/{docstring_case,synthetic_b4}.pt from the oracle (oracle/fs2_oracle.py), in this container.

The reference itself cannot be imported (its arithmetic lives in speechbrain, which is absent and un-pinned:
SURVEY.md 8c), so these vectors freeze the ORACLE's outputs on seeded inputs; the known-answer cases that the
reference does pin (the shape doctest model.py:133-146) and the quirk cases are asserted
in tests/test_oracle.py.  Run:  python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fs2_oracle as O  # noqa: E402

data = importlib.import_module("fine-grained-emotional-control-of-tts_b200.data")


def docstring_case():
    """model.py:102-146: tokens (2,5), durations summing to 15/10."""
    tokens = torch.tensor([[13, 12, 31, 14, 19], [31, 16, 30, 31, 0]])
    durations = torch.tensor([[2, 4, 1, 5, 3], [1, 2, 4, 3, 0]])
    speakers = torch.tensor([1, 3])
    g = torch.Generator().manual_seed(5)
    pitch = torch.randn(2, 15, generator=g)
    energy = torch.randn(2, 15, generator=g)
    pitch[1, 10:] = 0
    energy[1, 10:] = 0
    intensity = torch.randn(2, 5, 5, generator=g)
    intensity[1, 4] = 0
    mel = torch.rand(2, 15, 80, generator=g) * 13.5 - 11.5
    mel[1, 10:] = 0
    return dict(tokens=tokens, durations=durations, speakers=speakers, pitch=pitch, energy=energy,
                intensity=intensity, mel=mel, mel_len=torch.tensor([15, 10]), phon_len=torch.tensor([5, 4]))


def run(case, dtype=torch.float64):
    model = O.build(seed=0, dtype=dtype).eval()
    crit = O.Loss(**O.DEFAULT_LOSS_CONFIG)
    c = case
    preds = model(c["tokens"], c["speakers"], c["durations"], c["pitch"].to(dtype), c["energy"].to(dtype),
                  intensity=c["intensity"].to(dtype))
    losses = crit(preds, (c["mel"].to(dtype), c["durations"], c["pitch"].to(dtype), c["energy"].to(dtype),
                          c["mel_len"], c["phon_len"]), 0)
    losses["total_loss"].backward()
    gn = {k: p.grad.double().norm().item() for k, p in model.named_parameters()}
    out = dict(case=case,
               preds=[p.detach().float() if p is not None else None for p in preds[:7]],
               mel_lens=preds[7],
               losses={k: float(v) for k, v in losses.items()},
               grad_norms=gn)
    return out


def main():
    torch.set_num_threads(4)
    g1 = run(docstring_case())
    torch.save(g1, os.path.join(HERE, "docstring_case.pt"))
    (batch, intensity), = data.synthetic_batches(4, 1, seed=11, min_tp=8, max_tp=24, max_frames=150, pool_factor=2)
    tokens, speakers, in_lens, mel, pitch, energy, dur, out_lens = batch[:8]
    case = dict(tokens=tokens, durations=dur, speakers=speakers, pitch=pitch, energy=energy, intensity=intensity,
                mel=mel, mel_len=out_lens, phon_len=in_lens)
    g2 = run(case)
    torch.save(g2, os.path.join(HERE, "synthetic_b4.pt"))
    for name, g in (("docstring_case", g1), ("synthetic_b4", g2)):
        print(name, g["mel_lens"].tolist(), {k: round(v, 6) for k, v in g["losses"].items()})


if __name__ == "__main__":
    main()
