"""Deterministic IntensityExtractor weights for the golden fixture: every tensor is drawn from its own generator seeded
by its state_dict key, so the 64 M parameters never have to be stored -- the fixture keeps inputs and the REAL
reference's outputs only, and both the generator script and the tests rebuild the same state_dict from this file."""
import zlib

import torch

CFG = dict(n_mels=80, n_heads=2, n_emotions=5, n_encoder_layers=6, hidden_dim=384, kernel_size=9, dropout=0.1)
# rank_model/parameter.yaml:52-58


def shapes(cfg=CFG):
    D, k, n_in, ne = cfg["hidden_dim"], cfg["kernel_size"], cfg["n_mels"] + 2, cfg["n_emotions"]
    out = {"input_proj.weight": (D, n_in), "input_proj.bias": (D,)}
    for i in range(cfg["n_encoder_layers"]):
        p = f"fft_block.layers.{i}."
        out.update({p + "self_attn.in_proj_weight": (3 * D, D), p + "self_attn.in_proj_bias": (3 * D,),
                    p + "self_attn.out_proj.weight": (D, D), p + "self_attn.out_proj.bias": (D,),
                    p + "conv1.weight": (4 * D, D, k), p + "conv1.bias": (4 * D,),
                    p + "conv2.weight": (D, 4 * D, k), p + "conv2.bias": (D,),
                    p + "norm1.weight": (D,), p + "norm1.bias": (D,), p + "norm2.weight": (D,), p + "norm2.bias": (D,)})
    out.update({"emotion_embedding.weight": (ne, D), "classifier.weight": (ne, D), "classifier.bias": (ne,)})
    return out


def state_dict(cfg=CFG, dtype=torch.float32):
    sd = {}
    for key, shape in shapes(cfg).items():
        g = torch.Generator().manual_seed(zlib.crc32(key.encode()))
        t = torch.randn(*shape, generator=g)
        if key.endswith("norm1.weight") or key.endswith("norm2.weight"):
            t = 1.0 + 0.1 * t
        elif len(shape) == 1:
            t = 0.05 * t
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = t / fan_in ** 0.5
        sd[key] = t.to(dtype)
    return sd
