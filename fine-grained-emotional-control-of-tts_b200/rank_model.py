"""B200-native forward of the frozen IntensityExtractor (`/root/reference/emo_rank_tts/rank_model/model.py:8-109`,
SURVEY 8f row 2): the network the reference runs in front of EVERY FastSpeech2 training step (train.py:27) to turn
(mel, pitch, energy) frames into per-frame emotion intensities.  Same constructor arguments, same state_dict keys and
shapes as the reference class (parameters are held by ordinary torch modules -- plumbing only, their forward is never
called), same `forward(x, length, emotions) -> (B, T, n_emotions)`.

Everything numeric runs in the C-ABI library: the 82 -> 384 input projection, six post-norm FFT blocks (fused tcgen05
attention with a plain key-padding mask; Conv1d k=9 384 -> 1536, GELU, Conv1d k=9 1536 -> 384 as zero-padded
implicit GEMMs; residual + LayerNorm), the emotion-embedding shift, mask and 384 -> 5 classifier.  Inference only
(`torch.no_grad()`, dropout inactive), bf16 operands with fp32 accumulation; no CPU fallback.

Input layout: the module's contract is (B, T, n_mels + 2) (model.py:97-101), but the FastSpeech2 collate hands over
`rank_X` as (B, n_mels + 2, T) (fastspeech2/dataset.py:94, 116) -- both are accepted, told apart by where the
n_mels + 2 axis sits."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib as L

PAD = L.PAD


def _rup(x, m):
    return (x + m - 1) // m * m


class _Layer(nn.Module):
    """Parameter holder with the reference layer's attribute names (rank_model/model.py:19-31)."""

    def __init__(self, n_heads, hidden_dim, kernel_size, dropout):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(embed_dim=hidden_dim, num_heads=n_heads, dropout=dropout, batch_first=True)
        self.conv1 = nn.Conv1d(hidden_dim, hidden_dim * 4, kernel_size=kernel_size, padding=kernel_size // 2)
        self.conv2 = nn.Conv1d(hidden_dim * 4, hidden_dim, kernel_size=kernel_size, padding=kernel_size // 2)
        self.norm1 = nn.LayerNorm(hidden_dim)
        self.norm2 = nn.LayerNorm(hidden_dim)


class _Stack(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)        # nn.TransformerEncoder's attribute name -> "fft_block.layers.{i}.*"


class IntensityExtractor(nn.Module):
    def __init__(self, n_mels, n_heads, n_emotions, n_encoder_layers, hidden_dim, kernel_size, dropout):
        super().__init__()
        if hidden_dim // n_heads != 192 or hidden_dim % n_heads:
            raise NotImplementedError("fs2_b200: the fused attention kernel covers head_dim 192 (parameter.yaml: 384 / 2)")
        if kernel_size % 2 == 0 or kernel_size // 2 > PAD:
            raise NotImplementedError("fs2_b200: odd conv kernel sizes up to 9")
        self.n_in, self.H, self.D, self.k, self.n_out = n_mels + 2, n_heads, hidden_dim, kernel_size, n_emotions
        self.input_proj = nn.Linear(n_mels + 2, hidden_dim)
        self.fft_block = _Stack([_Layer(n_heads, hidden_dim, kernel_size, dropout) for _ in range(n_encoder_layers)])
        self.emotion_embedding = nn.Embedding(n_emotions, hidden_dim)
        self.classifier = nn.Linear(hidden_dim, n_emotions)
        self._packed = None
        self._packed_key = None
        self._ws = {}

    # ------------------------------------------------------------------ operand copies (weights are frozen)
    def _pack(self):
        dev = self.input_proj.weight.device
        key = (dev, tuple(p._version for p in self.parameters()), tuple(p.data_ptr() for p in self.parameters()))
        if self._packed is not None and self._packed_key == key:
            return self._packed
        bf = torch.bfloat16
        kin = _rup(self.n_in, 16)                                    # K of the input projection padded for TMA
        w_in = torch.zeros(self.D, kin, device=dev, dtype=bf)
        w_in[:, : self.n_in] = self.input_proj.weight.detach().to(bf)
        layers = []
        for ly in self.fft_block.layers:
            tap_major = lambda w: w.detach().permute(0, 2, 1).contiguous().reshape(w.shape[0], -1).to(bf)   # [Cout][tap][Cin]
            layers.append(dict(
                wqkv=ly.self_attn.in_proj_weight.detach().to(bf).contiguous(), bqkv=ly.self_attn.in_proj_bias.detach().float(),
                wo=ly.self_attn.out_proj.weight.detach().to(bf).contiguous(), bo=ly.self_attn.out_proj.bias.detach().float(),
                w1=tap_major(ly.conv1.weight), b1=ly.conv1.bias.detach().float(),
                w2=tap_major(ly.conv2.weight), b2=ly.conv2.bias.detach().float(),
                g1=ly.norm1.weight.detach().float(), be1=ly.norm1.bias.detach().float(), eps1=ly.norm1.eps,
                g2=ly.norm2.weight.detach().float(), be2=ly.norm2.bias.detach().float(), eps2=ly.norm2.eps))
        self._packed = dict(kin=kin, w_in=w_in, b_in=self.input_proj.bias.detach().float(), layers=layers,
                            emb=self.emotion_embedding.weight.detach().float().contiguous(),
                            wc=self.classifier.weight.detach().float().contiguous(),
                            bc=self.classifier.bias.detach().float().contiguous())
        self._packed_key = key
        return self._packed

    def _workspace(self, B, T, dev, kin):
        key = (B, T, dev)
        ws = self._ws.get(key)
        if ws is None:
            rows = B * (T + 2 * PAD)
            D, bf = self.D, torch.bfloat16
            z = lambda *s, dt=torch.float32: torch.zeros(*s, device=dev, dtype=dt)
            ws = dict(x_in=z(rows, kin, dt=bf), xa_f32=z(rows, D), xa_act=z(rows, D, dt=bf), xb_f32=z(rows, D),
                      xb_act=z(rows, D, dt=bf), qkv=z(rows, 3 * D, dt=bf), o=z(rows, D, dt=bf), proj=z(rows, D),
                      hid=z(rows, 4 * D, dt=bf), ffn=z(rows, D))
            self._ws = {key: ws}                       # one shape resident at a time
        return ws

    # ------------------------------------------------------------------ helpers
    @staticmethod
    def _gemm(x, w, out, *, rows, cin, cout, k, T, bias, c_bf16, act=0):
        """y[r] = sum_j x[r + j - k//2] . W_j + bias over the padded row space; halo rows of y are written as zeros
        (zero padding of the next conv), rows of the rectangle beyond an utterance's length stay live (as in torch)."""
        L.gemm(mode=0, M=rows, N=cout, K=cin, taps=k, A=x, lda=cin, a_rows=rows, a_inner=cin, a_row_off=-(k // 2),
               a_tap_step=1, B=w, ldb=k * cin, b_rows=cout, b_inner=k * cin, b_tap_step=cin, Cout=out, ldc=cout,
               c_bf16=c_bf16, ab_bf16=True, bias=bias, relu=act, rs_T=T, rs_Tp=T + 2 * PAD, lens=None, halo=0)

    @staticmethod
    def _ln(B, T, C, x, branch, gamma, beta, eps, out_f32, out_act):
        p = L.Fs2LnFwd()
        p.B, p.T, p.C = B, T, C
        p.x, p.branch = x.data_ptr(), branch.data_ptr()
        p.drop_b_p, p.drop_b_seed, p.drop_a_p, p.drop_a_seed = 0.0, 0, 0.0, 0
        p.gamma, p.beta, p.eps, p.tanh_act = gamma.data_ptr(), beta.data_ptr(), float(eps), 0
        p.lens, p.post_add = None, None
        p.out_f32, p.out_act, p.act_bf16, p.halo = out_f32.data_ptr(), out_act.data_ptr(), 1, 0
        p.mean, p.rstd, p.seed_dev = None, None, None
        L.call("fs2_ln_fwd", L.C.addressof(p))

    # ------------------------------------------------------------------ forward (rank_model/model.py:97-109)
    @torch.no_grad()
    def forward(self, x, length, emotions, channels_first=None):
        """channels_first: True for the collate's rank_X layout (B, n_mels + 2, T) (dataset.py:116-117), False for
        (B, T, n_mels + 2) as rank_model/model.py:97 documents it; None infers it from the shape and refuses the one
        ambiguous case (both axes equal n_mels + 2, i.e. a batch padded to exactly 82 frames) instead of guessing."""
        if not x.is_cuda:
            raise RuntimeError("fs2_b200: IntensityExtractor inputs must be CUDA tensors (there is no CPU fallback)")
        if x.dim() != 3:
            raise ValueError("x must be (B, T, n_mels + 2) or (B, n_mels + 2, T)")
        if channels_first is None:
            if x.shape[1] == self.n_in and x.shape[2] == self.n_in:
                raise ValueError(f"x is (B, {self.n_in}, {self.n_in}): the layout cannot be inferred, pass channels_first=")
            channels_first = x.shape[-1] != self.n_in
        channels_first = bool(channels_first)
        if x.shape[1 if channels_first else 2] != self.n_in:
            raise ValueError(f"the feature axis of x must have size n_mels + 2 = {self.n_in}")
        B = x.shape[0]
        T = x.shape[2] if channels_first else x.shape[1]
        if T <= PAD:
            raise ValueError("fs2_b200: need more than 4 frames")
        dev = x.device
        pk = self._pack()
        ws = self._workspace(B, T, dev, pk["kin"])
        D, H, k = self.D, self.H, self.k
        rows = B * (T + 2 * PAD)
        lens = length.to(device=dev, dtype=torch.int32).contiguous()
        emotions = emotions.to(device=dev, dtype=torch.int64).contiguous()
        x = x.contiguous().float()
        L.call("fs2_frames_to_rows", x, int(channels_first), B, self.n_in, T, pk["kin"], ws["x_in"], 1)
        # H = input_proj(x): fp32 stream + bf16 operand copy (a LayerNorm-free "residual" start: written by two GEMM passes
        # would double the work, so the fp32 stream comes from the GEMM and the operand copy from a cast)
        self._gemm(ws["x_in"], pk["w_in"], ws["xa_f32"], rows=rows, cin=pk["kin"], cout=D, k=1, T=T, bias=pk["b_in"], c_bf16=False)
        L.call("fs2_cast_bf16", ws["xa_f32"], ws["xa_act"], rows * D)
        cur_f32, cur_act, nxt_f32, nxt_act = ws["xa_f32"], ws["xa_act"], ws["xb_f32"], ws["xb_act"]
        ldk = _rup(T, 8)
        scale = 1.0 / math.sqrt(D // H)
        for ly in pk["layers"]:
            # self-attention (model.py:34-36): key-padding mask only, dropout inactive
            self._gemm(cur_act, ly["wqkv"], ws["qkv"], rows=rows, cin=D, cout=3 * D, k=1, T=T, bias=ly["bqkv"], c_bf16=True)
            L.call("fs2_flash_attn_fwd", ws["qkv"], lens, B, H, T, D, scale, 0.0, 0, None, None, ws["o"], 1)
            self._gemm(ws["o"], ly["wo"], ws["proj"], rows=rows, cin=D, cout=D, k=1, T=T, bias=ly["bo"], c_bf16=False)
            self._ln(B, T, D, cur_f32, ws["proj"], ly["g1"], ly["be1"], ly["eps1"], nxt_f32, nxt_act)
            cur_f32, cur_act, nxt_f32, nxt_act = nxt_f32, nxt_act, cur_f32, cur_act
            # convolutional feed-forward (model.py:39-46): conv k -> GELU -> conv k, zero padding
            self._gemm(cur_act, ly["w1"], ws["hid"], rows=rows, cin=D, cout=4 * D, k=k, T=T, bias=ly["b1"], c_bf16=True)
            L.call("fs2_gelu", ws["hid"], rows * 4 * D, 1)          # halo rows hold zeros: GELU(0) = 0 keeps the zero padding
            self._gemm(ws["hid"], ly["w2"], ws["ffn"], rows=rows, cin=4 * D, cout=D, k=k, T=T, bias=ly["b2"], c_bf16=False)
            self._ln(B, T, D, cur_f32, ws["ffn"], ly["g2"], ly["be2"], ly["eps2"], nxt_f32, nxt_act)
            cur_f32, cur_act, nxt_f32, nxt_act = nxt_f32, nxt_act, cur_f32, cur_act
        out = torch.empty(B, T, self.n_out, device=dev, dtype=torch.float32)
        L.call("fs2_intensity_head", cur_f32, pk["emb"], emotions, lens, pk["wc"], pk["bc"], B, T, D, self.n_out, out)
        return out
