"""Freezes outputs of the REAL reference glue (fastspeech2/model.py + loss.py executed from /root/reference, speechbrain
leaves supplied by the oracle's restatements: tests/reference_glue.py) on the seeded cases of make_golden.py, so that
boxes without /root/reference can still check the oracle's FastSpeech2 / Loss glue against the reference's own code.
Run in the build container:  python tests/golden/make_glue_golden.py   ->  tests/golden/reference_glue.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fs2_oracle as O  # noqa: E402
import reference_glue as RG  # noqa: E402


def run_reference(case, mode):
    """mode 'train': teacher-forced forward + Loss + backward in fp32 (the reference's Loss casts its duration targets to
    fp32, loss.py:104-125, so it cannot run in fp64); 'fwd64': the same forward in fp64, predictions only;
    'infer<pace>': predicted durations, no targets, fp64."""
    FS2, Loss = RG.load()
    torch.manual_seed(0)                                   # same seed and construction order as O.build(seed=0)
    model = FS2(**O.DEFAULT_MODEL_CONFIG, n_speakers=4).eval()
    c = case
    if mode == "train":
        preds = model(c["tokens"], c["speakers"], c["durations"], c["pitch"], c["energy"], intensity=c["intensity"])
        losses = Loss(**O.DEFAULT_LOSS_CONFIG)(preds, (c["mel"], c["durations"], c["pitch"], c["energy"], c["mel_len"],
                                                         c["phon_len"]), 0)
        losses["total_loss"].backward()
        return dict(preds=[p.detach() if p is not None else None for p in preds[:7]], mel_lens=preds[7],
                    losses={k: float(v) for k, v in losses.items()},
                    grad_norms={k: p.grad.double().norm().item() for k, p in model.named_parameters()})
    model, d = model.double(), torch.float64
    with torch.no_grad():
        if mode == "fwd64":
            preds = model(c["tokens"], c["speakers"], c["durations"], c["pitch"].to(d), c["energy"].to(d),
                          intensity=c["intensity"].to(d))
        else:
            model.durPred.linear.w.bias.data.fill_(DUR_BIAS)      # random init predicts ~0 frames per phoneme
            preds = model(c["tokens"], c["speakers"], pace=float(mode[5:]), intensity=c["intensity"].to(d))
    return dict(preds=[p.detach() if p is not None else None for p in preds[:7]], mel_lens=preds[7])


DUR_BIAS = 1.5
MODES = ("train", "fwd64", "infer0.8", "infer1.0", "infer1.2")


def main():
    torch.set_num_threads(4)
    cases = {"docstring_case": torch.load(os.path.join(HERE, "docstring_case.pt"))["case"],
             "synthetic_b4": torch.load(os.path.join(HERE, "synthetic_b4.pt"))["case"]}
    out = {}
    for name, case in cases.items():
        out[name] = {m: run_reference(case, m) for m in MODES}
    # train.py:16-51 run from its source with a stand-in extractor (tests/reference_glue.py)
    batch, frames = RG.intensity_case()
    out["intensity_rep"] = RG.load_get_intensity_representation()(lambda x, l, e: frames, batch, torch.device("cpu"))
    torch.save(out, os.path.join(HERE, "reference_glue.pt"))
    print({k: {m: list(v[m]["mel_lens"].tolist()) for m in v} for k, v in out.items() if k != "intensity_rep"})


if __name__ == "__main__":
    main()
