"""Freeze outputs of the REAL reference collate (fastspeech2/dataset.py:60-133) on seeded ragged samples.
The class is executed from its own source text (the module itself imports speechbrain-dependent helpers at the top):
    python tests/golden/make_collate_golden.py"""
import ast
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
def reference_collate():
    """The reference class, executed from its source text (only available in the build container)."""
    src = open("/root/reference/emo_rank_tts/fastspeech2/dataset.py").read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TextMelCollateWithAlignment")
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[cls], type_ignores=[]), "ref_collate", "exec"), ns)
    return ns["TextMelCollateWithAlignment"]()


def samples(seed=3, n=7):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        tp = int(torch.randint(3, 20, (1,), generator=g))
        dur = torch.randint(0, 7, (tp,), generator=g)
        dur[0] += 1
        tm = int(dur.sum())
        out.append(dict(phoneme=torch.randint(1, 95, (tp,), generator=g), duration=dur,
                        mel=torch.randn(80, tm, generator=g), pitch=torch.randn(tm, generator=g),
                        energy=torch.randn(tm, generator=g), speaker=int(torch.randint(0, 4, (1,), generator=g)),
                        emotion=int(torch.randint(0, 5, (1,), generator=g)), text=f"utt{i}", audio_path=f"/x/{i}.wav"))
    return out


if __name__ == "__main__":
    batch = samples()
    out = reference_collate()(batch)
    torch.save({"out": out}, os.path.join(HERE, "collate_ref.pt"))
    print([tuple(o.shape) if torch.is_tensor(o) else o for o in out])
