"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

CPU restatement of the reference's IntensityExtractor forward
(`/root/reference/emo_rank_tts/rank_model/model.py:8-109`) with explicit tensor algebra instead of
nn.MultiheadAttention / nn.TransformerEncoder, so that the GPU parity tests do not depend on torch's fused
inference fast paths.  Parity PINNED: the reference class itself is pure torch and importable in the build
container; `tests/golden/make_rank_golden.py` runs the REAL reference on seeded inputs and freezes its
state_dict, inputs and outputs in `tests/golden/rank_extractor.pt`; `tests/test_rank_oracle.py` checks this
restatement against that fixture (and against the live reference when /root/reference is present).

State-dict keys are the reference's (input_proj.*, fft_block.layers.{i}.{self_attn.in_proj_weight, ...,
conv1, conv2, norm1, norm2}.*, emotion_embedding.weight, classifier.*)."""
import math

import torch
import torch.nn.functional as F


def intensity_extractor_forward(sd, x, length, emotions, n_heads=2, eps=1e-5):
    """sd: reference state_dict (any float dtype); x (B, T, n_mels+2); length (B,); emotions (B,) -> (B, T, n_emotions)."""
    dt = sd["input_proj.weight"].dtype
    x = x.to(dt)
    B, T, _ = x.shape
    pad = torch.arange(T)[None, :] >= length[:, None]                       # model.py:84-92  True = padded frame
    h = x @ sd["input_proj.weight"].t() + sd["input_proj.bias"]             # model.py:99
    D = h.shape[-1]
    hd = D // n_heads
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("fft_block.layers."))
    for i in range(n_layers):
        p = f"fft_block.layers.{i}."
        # --- self attention, post-norm (model.py:34-36); dropout inactive in eval
        qkv = h @ sd[p + "self_attn.in_proj_weight"].t() + sd[p + "self_attn.in_proj_bias"]
        q, k, v = (t.reshape(B, T, n_heads, hd).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        s = s.masked_fill(pad[:, None, None, :], float("-inf"))            # key_padding_mask: every head, keys only
        a = torch.softmax(s, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, T, D)
        a = a @ sd[p + "self_attn.out_proj.weight"].t() + sd[p + "self_attn.out_proj.bias"]
        h = F.layer_norm(h + a, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
        # --- conv feed-forward (model.py:39-49): Conv1d(k, zero padding) -> GELU(erf) -> Conv1d(k, zero padding)
        w1, w2 = sd[p + "conv1.weight"], sd[p + "conv2.weight"]
        y = F.conv1d(h.transpose(1, 2), w1, sd[p + "conv1.bias"], padding=w1.shape[-1] // 2)
        y = F.gelu(y)
        y = F.conv1d(y, w2, sd[p + "conv2.bias"], padding=w2.shape[-1] // 2).transpose(1, 2)
        h = F.layer_norm(h + y, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps)
    i_ = h + sd["emotion_embedding.weight"][emotions][:, None, :]          # model.py:104-105
    i_ = i_.masked_fill(pad[:, :, None], 0.0)                               # model.py:107
    return i_ @ sd["classifier.weight"].t() + sd["classifier.bias"]         # model.py:108
