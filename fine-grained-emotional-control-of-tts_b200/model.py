"""Drop-in `FastSpeech2` for `/root/reference/emo_rank_tts/fastspeech2/model.py:149-441`.

Same constructor kwargs (parameter.yaml:62-90 + n_speakers), same forward signature and 8-tuple, same
state_dict keys -- but every numeric op of the forward AND the backward is a hand-written sm_100a kernel
reached through the C ABI of include/fs2_b200.h (libfs2_b200.so).  torch is used for device memory,
streams and autograd wiring only.  There is no CPU path.

Layout: activations live in a "padded row space" (B items x (T+8) rows, 4 halo rows either side of every
item) so that every reflect-"same"-padded Conv1d of the reference becomes a shifted-row implicit GEMM, and
its dgrad / wgrad are the same GEMM kernel in its other two operand modes.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn

from . import _lib as L
from .params import ParamStore

PAD = L.PAD
_M64 = (1 << 64) - 1


def _rup(x, m):
    return (x + m - 1) // m * m


class Arena:
    """Bump allocator over a zero-initialised device buffer of ONE dtype (stale rows that a sweep-style GEMM
    may touch are then always finite values of the right type, never reinterpreted bits)."""

    def __init__(self, dtype):
        self.dtype = dtype
        self.buf = None
        self.off = 0
        self.extra = []
        self.high = 0

    def reset(self, device):
        need = self.high
        if self.buf is None or self.buf.device != device or self.extra:
            size = max(int(need * 1.25), 1 << 20)
            if self.buf is None or self.buf.device != device or size > self.buf.numel():
                self.buf = torch.zeros(size, dtype=self.dtype, device=device)
        self.extra = []
        self.off = 0
        self.high = 0

    def alloc(self, *shape):
        n = int(math.prod(shape))
        es = self.buf.element_size()
        n_al = _rup(n * es, 256) // es
        if self.off + n_al <= self.buf.numel():
            t = self.buf[self.off:self.off + n]
            self.off += n_al
        else:
            chunk = torch.zeros(n_al, dtype=self.dtype, device=self.buf.device)
            self.extra.append(chunk)
            t = chunk[:n]
        self.high += n_al
        return t.view(*shape)


class _Saved:
    pass


class FastSpeech2(nn.Module):
    """B200-native FastSpeech2 with speaker + emotion-intensity conditioning (reference model.py:30-441).

    Extra keyword (not in the reference): `precision` = "bf16" (tcgen05 tensor-core GEMMs, bf16 operands,
    fp32 accumulation / residual stream / LayerNorm statistics) or "fp32" (exact SIMT GEMMs).
    """

    def __init__(
        self, enc_num_layers, enc_num_head, enc_d_model, enc_ffn_dim, enc_k_dim, enc_v_dim, enc_dropout,
        dec_num_layers, dec_num_head, dec_d_model, dec_ffn_dim, dec_k_dim, dec_v_dim, dec_dropout,
        normalize_before, ffn_type, ffn_cnn_kernel_size_list, n_char, n_mels, postnet_embedding_dim,
        postnet_kernel_size, postnet_n_convolutions, postnet_dropout, padding_idx, dur_pred_kernel_size,
        pitch_pred_kernel_size, energy_pred_kernel_size, variance_predictor_dropout, n_speakers,
        precision="bf16",
    ):
        super().__init__()
        if ffn_type != "1dcnn":
            raise NotImplementedError("fs2_b200: only ffn_type='1dcnn' (the reference's parameter.yaml:78)")
        if normalize_before:
            raise NotImplementedError("fs2_b200: only normalize_before=False (the reference's parameter.yaml:77)")
        if enc_d_model != dec_d_model:
            raise ValueError("enc_d_model must equal dec_d_model (the length regulator feeds one into the other)")
        if enc_k_dim != enc_d_model or enc_v_dim != enc_d_model or dec_k_dim != dec_d_model or dec_v_dim != dec_d_model:
            raise NotImplementedError("fs2_b200: kdim/vdim must equal d_model (packed in_proj_weight, as the reference)")
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        D = enc_d_model
        self.D, self.n_mels, self.n_char, self.n_speakers = D, n_mels, n_char, n_speakers
        self.padding_idx = padding_idx
        self.precision = precision
        self.enc = dict(name="encoder", nl=enc_num_layers, H=enc_num_head, F=enc_ffn_dim, p=float(enc_dropout))
        self.dec = dict(name="decoder", nl=dec_num_layers, H=dec_num_head, F=dec_ffn_dim, p=float(dec_dropout))
        self.k0, self.k1 = int(ffn_cnn_kernel_size_list[0]), int(ffn_cnn_kernel_size_list[1])
        self.kd = int(dur_pred_kernel_size)          # all three predictors use it (model.py:211,217,223; quirk Q4)
        self.kp, self.ke = int(pitch_pred_kernel_size), int(energy_pred_kernel_size)
        self.var_p = float(variance_predictor_dropout)
        self.E, self.kpn, self.npn, self.pn_p = postnet_embedding_dim, int(postnet_kernel_size), int(postnet_n_convolutions), float(postnet_dropout)
        for k in (self.k0, self.k1, self.kd, self.kp, self.ke, self.kpn):
            if k % 2 != 1 or k > 2 * PAD + 1:
                raise ValueError("kernel sizes must be odd and <= 9")
        if D % 8 or enc_ffn_dim % 8 or dec_ffn_dim % 8 or n_mels % 8 or self.E % 8 or D > 512 or self.E > 512:
            raise ValueError("fs2_b200: channel sizes must be multiples of 8; d_model and postnet dim <= 512")
        if D % enc_num_head or D % dec_num_head:
            raise ValueError("d_model must be divisible by the number of heads")
        if self.npn < 2:
            raise ValueError("postnet_n_convolutions must be >= 2")

        st = ParamStore(self)
        self.store = st
        st.add("speaker_emb.Embedding.weight", (n_speakers, D), ("normal",))
        st.add("concat_proj.w.weight", (D, 2 * D + 5), ("kaiming_uniform", 2 * D + 5))
        st.add_packed("concat_proj.w.weight", D, D, 1, src_ld=2 * D + 5, src_col0=0, name="concat_proj.tok")
        st.add("encPreNet.token_embedding.Embedding.weight", (n_char, D), ("normal",))
        for pred in ("durPred", "pitchPred", "energyPred"):
            st.conv(f"{pred}.conv1.conv", D, D, self.kd)
            st.conv(f"{pred}.conv2.conv", D, D, self.kd)
            st.add(f"{pred}.linear.w.weight", (1, D), ("kaiming_uniform", D))
            st.add(f"{pred}.linear.w.bias", (1,), ("uniform_fan", D))
            st.layernorm(f"{pred}.ln1.norm", D)
            st.layernorm(f"{pred}.ln2.norm", D)
        st.conv("pitchEmbed.conv", D, 1, self.kp, pack=False)
        st.conv("energyEmbed.conv", D, 1, self.ke, pack=False)
        for cfg in (self.enc, self.dec):
            for l in range(cfg["nl"]):
                pre = f"{cfg['name']}.layers.{l}"
                st.add(f"{pre}.self_att.att.in_proj_weight", (3 * D, D), ("xavier_uniform",))
                st.add(f"{pre}.self_att.att.in_proj_bias", (3 * D,), ("zeros",))
                st.add_packed(f"{pre}.self_att.att.in_proj_weight", 3 * D, D, 1)
                st.add(f"{pre}.self_att.att.out_proj.weight", (D, D), ("kaiming_uniform", D))
                st.add(f"{pre}.self_att.att.out_proj.bias", (D,), ("zeros",))
                st.add_packed(f"{pre}.self_att.att.out_proj.weight", D, D, 1)
                st.conv(f"{pre}.pos_ffn.0.conv", cfg["F"], D, self.k0)
                st.conv(f"{pre}.pos_ffn.2.conv", D, cfg["F"], self.k1)
                st.layernorm(f"{pre}.norm1.norm", D)
                st.layernorm(f"{pre}.norm2.norm", D)
            st.layernorm(f"{cfg['name']}.norm.norm", D)
        st.linear("linear.w", n_mels, D)
        st.conv("postnet.conv_pre.conv", self.E, n_mels, self.kpn)
        for i in range(self.npn - 2):
            st.conv(f"postnet.convs_intermedite.{i}.conv", self.E, self.E, self.kpn)
        st.conv("postnet.conv_post.conv", n_mels, self.E, self.kpn)
        st.layernorm("postnet.ln1", self.E)
        st.layernorm("postnet.ln2", self.E)
        st.layernorm("postnet.ln3", n_mels)
        st.build()

        # persistent sinusoid buffers, same keys as speechbrain PositionalEncoding (model.py:187-192)
        pe = torch.zeros(2500, D)
        pos = torch.arange(0, 2500).unsqueeze(1).float()
        den = torch.exp(torch.arange(0, D, 2).float() * -(math.log(10000.0) / D))
        pe[:, 0::2] = torch.sin(pos * den)
        pe[:, 1::2] = torch.cos(pos * den)
        for nm in ("sinusoidal_positional_embed_encoder", "sinusoidal_positional_embed_decoder"):
            m = nn.Module()
            m.register_buffer("pe", pe.unsqueeze(0).clone())
            self.add_module(nm, m)

        self._anchor = torch.zeros(1, requires_grad=True)   # lets autograd reach our backward; never a parameter
        self._arenas = None
        self._seed_base = 0x1234
        # encoder layers at which the backward is cut into separately launched parts (descending): after each part the
        # gradients of the layers it covered are final and are reported through grad_ready_hook (parallel.py)
        self.enc_grad_splits = (4, 2, 1)
        self._generation = 0
        self._ctr = None              # device-side dropout step counter (uint64 in an int64 tensor)
        self.use_cuda_graphs = False  # opt-in: replay the captured step per (B, Tp, Tm) -- see _graph_forward
        self._graphs = {}
        self._seen = set()
        self._pre = None
        self.fused_attention = True   # bf16, head_dim 192: fs2_flash_attn_fwd / _bwd instead of GEMM + softmax + GEMM
        # bf16, model width 384, k = 1 second FFN conv: out-projection -> norm1 and FFN conv 2 -> norm2 each run as ONE kernel
        # (fs2_gemm_ln_tc: the fp32 branch never reaches HBM); the backward takes x_hat from the saved LayerNorm output
        # bf16 path: bias gradients ride inside the weight-gradient GEMMs (Fs2Gemm.a_colsum) instead of separate column-sum launches
        self.fold_bias_grad = os.environ.get("FS2_FOLD_BIAS_GRAD", "1") != "0"
        self.fused_ln = os.environ.get("FS2_FUSED_LN", "1") != "0"
        self.fused_ln_min_rows_ffn = int(os.environ.get("FS2_FUSED_LN_MIN_ROWS_FFN", "8192"))
        self.async_mel_lens = False   # opt-in: no host sync; Tm is taken from pitch.shape[1], mel_lens arrives in pinned memory
        # opt-in, inference with predicted durations in precision="bf16": frame counts are trunc(pace * expm1(pred)), which is
        # discontinuous -- a bf16 encoder moves ~1 % of the phonemes across an integer boundary (tests/
        # test_parity_bench_configs_gpu.py).  With exact_durations the phoneme encoder, the conditioning and the duration
        # predictor are evaluated a second time on the exact fp32 path and THAT prediction drives the LengthRegulator (and is
        # returned as predict_durations); everything at mel-frame length stays on the tensor cores.
        self.exact_durations = False
        self._arenas_exact = None
        self._async_bufs = None
        self.replayed_launches = 0    # kernels launched through graph replays (fs2_launch_count only sees captures)
        self.trace = None        # set to a dict to collect unpadded intermediates (debug / parity tests)
        self._pin_lens = None
        # weight-gradient GEMMs (+ bias column sums) are off the backward's critical path: with overlap_wgrad they are
        # queued on a second stream (fork/join through events, also inside graph captures) and fill the SMs the
        # dgrad chain leaves idle (short grids, wave tails, memory-bound LN kernels)
        self.overlap_wgrad = True
        self._split_lo = None
        self._split_ok = None
        self.grad_ready_hook = None   # callable(lo, hi): flat-gradient range [lo, hi) is final (set by DataParallelStep)
        self._side = None
        self._side_stream = None
        self._side_reads = {}
        self._side_last = None

    # ------------------------------------------------------------------ nn.Module plumbing
    def _apply(self, fn, *a, **kw):
        super()._apply(fn, *a, **kw)
        self.store.reflatten()
        self._anchor = torch.zeros(1, device=self.store.flat.device, requires_grad=True)
        self._ctr = None
        self._graphs.clear()
        self._pre = None
        return self

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = dict(state_dict)
        for i in (1, 2, 3):     # accept speechbrain-LayerNorm style keys for the PostNet norms
            for leaf in ("weight", "bias"):
                alt = f"postnet.ln{i}.norm.{leaf}"
                if alt in sd:
                    sd[f"postnet.ln{i}.{leaf}"] = sd.pop(alt)
        return super().load_state_dict(sd, strict=strict, **kw)

    def manual_seed(self, seed):
        """Seed of the counter-based dropout masks (also rewinds the per-step counter)."""
        self._seed_base = int(seed) & _M64
        if self._ctr is not None:
            self._ctr.zero_()
        self._graphs.clear()       # seeds are baked into captured launches

    # --------------------------------------------------------------------------- helpers
    @property
    def _bf16(self):
        return self.precision == "bf16"

    def _sync_operands(self):
        st = self.store
        if st.flat.is_cuda:
            if not st.views_intact():
                st.reflatten()
            st.sync_operands(self._bf16)

    def _P(self, key):
        return self.store.params[key]

    def _G(self, key):
        return self.store.grad(key)

    def _f32(self, *shape):
        return self._arenas[0].alloc(*shape)

    def _act(self, *shape):
        return self._arenas[1].alloc(*shape)

    def _i32(self, *shape):
        return self._arenas[2].alloc(*shape)

    def _tr(self, name, t, B, T, C):
        if self.trace is not None:
            self.trace[name] = t.view(B, T + 2 * PAD, C)[:, PAD:PAD + T].float().clone()

    @staticmethod
    def _site_seed(base, site):
        return (base * 0x9E3779B97F4A7C15 + site * 0xD1B54A32D192ED03 + 0x2545F4914F6CDD1D) & _M64

    def _fused_attn(self, H):
        """The fused tcgen05 attention kernel covers the reference geometry (2 heads x 192) in the bf16 path."""
        return self.fused_attention and self._bf16 and self.D // H == 192 and self.D % H == 0

    def _conv(self, x, B, T, wname, out, *, c_bf16, bias=None, relu=0, lens=None, halo=0):
        """y[r] = sum_j x[r + j - p] . W_j (+bias, ReLU, row mask, reflect-halo mirror): model.py Conv1d/Linear sites."""
        w = self.store.pw(wname)
        wbuf, woff, wld = self.store.operand(wname, self._bf16)
        rows = B * (T + 2 * PAD)
        p = (w.k - 1) // 2
        L.gemm(mode=0, M=rows, N=w.cout, K=w.cin, taps=w.k, A=x, lda=w.cin, a_rows=rows, a_inner=w.cin,
               a_row_off=-p, a_tap_step=1, B=wbuf, B_off=woff, ldb=wld, b_rows=w.cout,
               b_inner=w.k * w.cin, b_tap_step=w.cin, Cout=out, ldc=w.cout, c_bf16=c_bf16, ab_bf16=self._bf16,
               bias=bias, relu=relu, rs_T=T, rs_Tp=T + 2 * PAD, lens=lens, halo=halo)

    def _conv_dgrad(self, dy, B, T, wname, out, *, c_bf16=False, relu_aux=None, split=1):
        """dx[r] = sum_j dy[r + p - j] . W_j  (gradient wrt the padded input, halo rows included).
        split = 2: `out` is (2, rows, Cin); each half of the reduction is STORED to its own slice (deterministic split-K,
        the consumer adds the two) -- used where one wave of tiles leaves SMs idle, see _dgrad_split."""
        w = self.store.pw(wname)
        wbuf, woff, wld = self.store.operand(wname, self._bf16)
        rows = B * (T + 2 * PAD)
        p = (w.k - 1) // 2
        self._wait_side(out)
        L.gemm(mode=1, M=rows, N=w.cin, K=w.cout, taps=w.k, A=dy, lda=w.cout, a_rows=rows, a_inner=w.cout,
               a_row_off=p, a_tap_step=-1, B=wbuf, B_off=woff, ldb=wld, b_rows=w.cout,
               b_inner=w.k * w.cin, b_tap_step=w.cin, Cout=out, ldc=w.cin, c_bf16=c_bf16, ab_bf16=self._bf16,
               relu_aux=relu_aux, aux_bf16=int(self._bf16), split_k=split,
               c_split_stride=rows * w.cin if split > 1 else 0)
        # (split_k=0 would let fs2_gemm_tc split short-grid dgrads WITH ATOMICS; measured: no step-time gain once the weight gradients
        #  fill the idle SMs from the second stream, and the atomics' summation order would leak into the bf16 roundings
        #  of the activation gradients -- run-to-run differences of 5e-4 instead of 1e-6 -- so it stays off)

    def _dgrad_split(self, rows, cin):
        """2 when splitting the k = 9 FFN dgrad's reduction in two shortens its critical path: tiles * 2 work items fill
        the persistent grid in fewer (half-length) rounds than the tiles alone (the kernel choice of fs2_gemm_tc is mirrored
        here: CTA-pair 256 x 384 tiles from 8192 rows, else 128 x 192 tiles)."""
        if not self._bf16 or cin != 384 or os.environ.get("FS2_DGRAD_SPLIT", "1") == "0":
            return 1
        if rows >= 8192:
            tiles, units = -(-rows // 256), 74
        else:
            tiles, units = -(-rows // 128) * 2, 148
        one, two = -(-tiles // units), -(-2 * tiles // units) / 2.0
        return 2 if two <= 0.85 * one else 1

    def _side_begin(self, dev):
        if not self.overlap_wgrad:
            return
        if self._side_stream is None or self._side_stream.device != dev:
            self._side_stream = torch.cuda.Stream(device=dev)
        self._side = self._side_stream
        self._side_reads = {}
        self._side_last = None

    def _side_join(self):
        """Main stream waits for everything queued on the side stream (end of backward)."""
        if self._side is not None and self._side_last is not None:
            torch.cuda.current_stream().wait_event(self._side_last)
        self._side = None
        self._side_reads = {}
        self._side_last = None

    def _wait_side(self, buf):
        """Call before the main stream overwrites a scratch buffer a side-stream GEMM may still be reading."""
        if self._side is not None and buf is not None:
            ev = self._side_reads.pop(buf.data_ptr(), None)
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

    def _conv_wgrad(self, dy, x, B, T, wname, wkey, bkey=None):
        side = self._side
        if side is None:
            return self._conv_wgrad_launch(dy, x, B, T, wname, wkey, bkey)
        ready = torch.cuda.Event()
        ready.record()                       # everything the main stream has queued so far (dy's producer included)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            self._conv_wgrad_launch(dy, x, B, T, wname, wkey, bkey)
            done = torch.cuda.Event()
            done.record()
        self._side_reads[dy.data_ptr()] = done
        self._side_last = done

    def _conv_wgrad_launch(self, dy, x, B, T, wname, wkey, bkey=None):
        """dW[co, ci, j] += sum_r dy[r, co] * x[r + j - p, ci];  db[co] += sum_r dy[r, co].  The gradient buffer has the
        master's layout (params.py): tap-major [Cout][k][Cin] for k > 1, so every tap's (Cout, Cin) block is contiguous
        in n and the split-K partial sums go out as 16-byte vector atomics."""
        w = self.store.pw(wname)
        rows = B * (T + 2 * PAD)
        p = (w.k - 1) // 2
        split = 0            # auto: chosen by fs2_gemm_tc from the tile count and the SM count
        fold = self._bf16 and self.fold_bias_grad
        o, _ = self.store.offsets[wkey]
        if w.gather:         # a column block of a wider (Cout, src_ld) matrix
            c_off, ldc, tap_stride = o + w.src_col0, w.src_ld, 1
        else:
            c_off, ldc, tap_stride = o, w.k * w.cin, w.cin
        L.gemm(mode=2, M=w.cout, N=w.cin, K=rows, taps=w.k, A=dy, lda=w.cout, a_rows=rows, a_inner=w.cout,
               B=x, ldb=w.cin, b_rows=rows, b_inner=w.cin, b_row_off=-p, b_tap_step=1,
               Cout=self.store.flat_grad, C_off=c_off, ldc=ldc, c_tap_stride=tap_stride, c_col_stride=1,
               c_bf16=False, ab_bf16=self._bf16, accumulate=1, split_k=split,
               a_colsum=self._G(bkey) if (bkey is not None and fold) else None)
        # bf16 path: the bias gradient (column sums of dy) rides along with the weight-gradient GEMM (Fs2Gemm.a_colsum)
        if bkey is not None and not fold:
            L.call("fs2_colsum", dy, int(self._bf16), rows, w.cout, w.cout, self._G(bkey))

    def _ln_fwd(self, B, T, C, x, gamma, beta, eps, *, branch=None, drop_b=(0.0, 0), tanh=0, drop_a=(0.0, 0),
                lens=None, post_add=None, out_f32=None, out_act=None, halo=0, mean=None, rstd=None, head=None):
        p = L.Fs2LnFwd()
        p.B, p.T, p.C = B, T, C
        p.x = x.data_ptr()
        p.branch = branch.data_ptr() if branch is not None else None
        p.drop_b_p, p.drop_b_seed = drop_b
        p.gamma, p.beta, p.eps = gamma.data_ptr(), beta.data_ptr(), eps
        p.tanh_act = tanh
        p.drop_a_p, p.drop_a_seed = drop_a
        p.lens = lens.data_ptr() if lens is not None else None
        p.post_add = post_add.data_ptr() if post_add is not None else None
        p.out_f32 = out_f32.data_ptr() if out_f32 is not None else None
        p.out_act = out_act.data_ptr() if out_act is not None else None
        p.act_bf16, p.halo = int(self._bf16), halo
        p.mean = mean.data_ptr() if mean is not None else None
        p.rstd = rstd.data_ptr() if rstd is not None else None
        if head is not None:
            hw, hb, hout, hs = head
            p.head_w, p.head_b, p.head_out, p.head_scale = hw.data_ptr(), hb.data_ptr(), hout.data_ptr(), hs
        p.seed_dev = self._ctr.data_ptr()
        L.call("fs2_ln_fwd", L.C.addressof(p))

    def _fused_ln(self):
        """fs2_gemm_ln_tc covers the reference geometry (width 384, second FFN conv k = 1) on the bf16 path."""
        return self.fused_ln and self._bf16 and self.D == 384 and self.k1 == 1

    def _gemm_ln(self, a, B, T, wname, bias, x, gamma, beta, eps, *, drop=(0.0, 0), out_f32, out_act, halo, mean, rstd):
        """out = LN(x + dropout(a . W^T + bias)) in one kernel (speechbrain's post-norm residual step)."""
        w = self.store.pw(wname)
        wbuf, woff, wld = self.store.operand(wname, True)
        q = L.Fs2GemmLn()
        q.B, q.T, q.K, q.lda, q.ldw = B, T, w.cin, w.cin, wld
        q.A, q.W = a.data_ptr(), wbuf.data_ptr() + 2 * woff
        q.bias, q.x = bias.data_ptr(), x.data_ptr()
        q.drop_p, q.drop_seed = drop
        q.seed_dev = self._ctr.data_ptr()
        q.gamma, q.beta, q.eps = gamma.data_ptr(), beta.data_ptr(), eps
        q.out_f32, q.out_act, q.halo = out_f32.data_ptr(), out_act.data_ptr(), halo
        q.mean, q.rstd = mean.data_ptr(), rstd.data_ptr()
        L.call("fs2_gemm_ln_tc", L.C.addressof(q))

    def _ln_bwd(self, B, T, C, x, gamma, beta, eps, mean, rstd, *, dy=None, dy2=None, dy3=None, dy2_fold=0, dhead=None,
                head_w=None, head_scale=1.0, branch=None, drop_b=(0.0, 0), tanh=0, drop_a=(0.0, 0), lens=None,
                relu_x=0, dx_f32=None, dact=None, dgamma=None, dbeta=None, dhead_w=None, dhead_b=None, dact_colsum=None,
                y=None):
        self._wait_side(dact)
        p = L.Fs2LnBwd()
        p.B, p.T, p.C = B, T, C
        p.dy = dy.data_ptr() if dy is not None else None
        p.dy2 = dy2.data_ptr() if dy2 is not None else None
        p.dy3 = dy3.data_ptr() if dy3 is not None else None
        p.dy2_fold = dy2_fold
        p.dhead = dhead.data_ptr() if dhead is not None else None
        p.head_w = head_w.data_ptr() if head_w is not None else None
        p.head_scale = head_scale
        p.x = x.data_ptr() if x is not None else None
        p.y = y.data_ptr() if y is not None else None
        p.branch = branch.data_ptr() if branch is not None else None
        p.drop_b_p, p.drop_b_seed = drop_b
        p.gamma, p.beta, p.eps, p.tanh_act = gamma.data_ptr(), beta.data_ptr(), eps, tanh
        p.drop_a_p, p.drop_a_seed = drop_a
        p.lens = lens.data_ptr() if lens is not None else None
        p.mean, p.rstd = mean.data_ptr() if mean is not None else None, rstd.data_ptr()
        p.relu_x = relu_x
        p.dx_f32 = dx_f32.data_ptr() if dx_f32 is not None else None
        p.dact = dact.data_ptr() if dact is not None else None
        p.act_bf16 = int(self._bf16)
        p.dgamma = dgamma.data_ptr() if dgamma is not None else None
        p.dbeta = dbeta.data_ptr() if dbeta is not None else None
        p.dhead_w = dhead_w.data_ptr() if dhead_w is not None else None
        p.dhead_b = dhead_b.data_ptr() if dhead_b is not None else None
        p.seed_dev = self._ctr.data_ptr()
        p.dact_colsum = dact_colsum.data_ptr() if dact_colsum is not None else None
        L.call("fs2_ln_bwd", L.C.addressof(p))

    # ------------------------------------------------------------------- FFT block stack
    def _attn_gemms_fwd(self, qkv, B, T, H, S, lens, P, Pd, O, p_drop, seed):
        D = self.D
        hd = D // H
        TP = T + 2 * PAD
        ld = 3 * D
        ldk = _rup(T, 8)
        bf = self._bf16
        if S is None:
            # flash-style tcgen05 attention: scores, softmax, dropout and PV in one kernel, nothing T x T in HBM
            # (flash_attention.cu); P is the per-row log-sum-exp statistic here (None: inference)
            L.call("fs2_flash_attn_fwd", qkv, lens, B, H, T, D, 1.0 / math.sqrt(hd), p_drop, seed, self._ctr, P, O, 0)
            return
        # S[b,h] = Q K^T
        L.gemm(mode=0, M=T, N=T, K=hd, A=qkv, A_off=PAD * ld, lda=ld, a_rows=T, a_inner=hd, a_s1=hd, a_s2=TP * ld,
               B=qkv, B_off=PAD * ld + D, ldb=ld, b_rows=T, b_inner=hd, b_s1=hd, b_s2=TP * ld, batch1=H, batch2=B,
               Cout=S, ldc=ldk, c_s1=T * ldk, c_s2=H * T * ldk, c_bf16=False, ab_bf16=bf)
        L.call("fs2_softmax_fwd", S, lens, B, H, T, ldk, 1.0 / math.sqrt(hd), p_drop, seed, self._ctr, P,
               Pd if p_drop > 0 else None, int(bf))
        # O[b,:,h] = Pd V
        L.gemm(mode=1, M=T, N=hd, K=T, A=Pd, lda=ldk, a_rows=T, a_inner=T, a_s1=T * ldk, a_s2=H * T * ldk,
               B=qkv, B_off=PAD * ld + 2 * D, ldb=ld, b_rows=T, b_inner=hd, b_s1=hd, b_s2=TP * ld, batch1=H, batch2=B,
               Cout=O, C_off=PAD * D, ldc=D, c_s1=hd, c_s2=TP * D, c_bf16=bf, ab_bf16=bf)

    def _stack_fwd(self, cfg, x_f32, x_act, B, T, lens, final_lens, final_halo, base_seed, site0, training, keep_p=True):
        D, F, H, nl = self.D, cfg["F"], cfg["H"], cfg["nl"]
        name = cfg["name"]
        rows = B * (T + 2 * PAD)
        ldk = _rup(T, 8)
        bf = self._bf16
        p = cfg["p"] if training else 0.0
        fused = self._fused_attn(H)
        fuse_ln = self._fused_ln()
        S = None if fused else self._f32(B * H, T, ldk)          # scratch, shared by all layers of this stack
        saves = []
        h1, h2 = (self.k0 - 1) // 2, (self.k1 - 1) // 2
        for l in range(nl):
            pre = f"{name}.layers.{l}"
            sv = _Saved()
            sv.x_f32, sv.x_act = x_f32, x_act
            sv.seeds = [self._site_seed(base_seed, site0 + 3 * l + i) for i in range(3)]
            sv.qkv = self._act(rows, 3 * D)
            self._conv(x_act, B, T, f"{pre}.self_att.att.in_proj_weight", sv.qkv, c_bf16=bf,
                       bias=self._P(f"{pre}.self_att.att.in_proj_bias"))
            if fused:
                # the probabilities stay on chip; the backward recomputes them from one fp32 statistic per query row
                sv.P = self._f32(B * H, _rup(T, 128)) if keep_p else None
                sv.Pd = None
            else:
                sv.P = self._act(B * H, T, ldk)
                sv.Pd = self._act(B * H, T, ldk) if p > 0 else sv.P
            sv.O = self._act(rows, D)
            self._attn_gemms_fwd(sv.qkv, B, T, H, S, lens, sv.P, sv.Pd, sv.O, p, sv.seeds[0])
            sv.x1_f32, sv.x1_act = self._f32(rows, D), self._act(rows, D)
            sv.mean1, sv.rstd1 = self._f32(rows), self._f32(rows)
            if fuse_ln:
                # out-projection + bias + dropout + residual + norm1 in one kernel: the fp32 branch stays in tensor memory
                sv.proj = None
                self._gemm_ln(sv.O, B, T, f"{pre}.self_att.att.out_proj.weight", self._P(f"{pre}.self_att.att.out_proj.bias"),
                              x_f32, self._P(f"{pre}.norm1.norm.weight"), self._P(f"{pre}.norm1.norm.bias"), 1e-6,
                              drop=(p, sv.seeds[1]), out_f32=sv.x1_f32, out_act=sv.x1_act, halo=h1, mean=sv.mean1,
                              rstd=sv.rstd1)
            else:
                sv.proj = self._f32(rows, D)
                self._conv(sv.O, B, T, f"{pre}.self_att.att.out_proj.weight", sv.proj, c_bf16=False,
                           bias=self._P(f"{pre}.self_att.att.out_proj.bias"))
                self._ln_fwd(B, T, D, x_f32, self._P(f"{pre}.norm1.norm.weight"), self._P(f"{pre}.norm1.norm.bias"), 1e-6,
                             branch=sv.proj, drop_b=(p, sv.seeds[1]), out_f32=sv.x1_f32, out_act=sv.x1_act, halo=h1,
                             mean=sv.mean1, rstd=sv.rstd1)
            sv.Hh = self._act(rows, F)
            self._conv(sv.x1_act, B, T, f"{pre}.pos_ffn.0.conv.weight", sv.Hh, c_bf16=bf,
                       bias=self._P(f"{pre}.pos_ffn.0.conv.bias"), relu=1, halo=h2)
            y_f32, y_act = self._f32(rows, D), self._act(rows, D)
            sv.mean2, sv.rstd2 = self._f32(rows), self._f32(rows)
            # (at phoneme-length row counts the K = 1536 form is on par at best: 17 CTA pairs walk 24 k-blocks each where the
            #  plain GEMM spreads 128 x 192 tiles over 68 SMs -- 28.0 vs 26.9 us at 4352 rows, profiles/r02_gemm_ln_bench_final.txt;
            #  step A/B on one box: 8.52 / 8.55 ms with the default fused_ln_min_rows_ffn = 8192, 8.58 / 8.58 ms with 0)
            if fuse_ln and rows >= self.fused_ln_min_rows_ffn:
                sv.Fo = None
                self._gemm_ln(sv.Hh, B, T, f"{pre}.pos_ffn.2.conv.weight", self._P(f"{pre}.pos_ffn.2.conv.bias"),
                              sv.x1_f32, self._P(f"{pre}.norm2.norm.weight"), self._P(f"{pre}.norm2.norm.bias"), 1e-6,
                              drop=(p, sv.seeds[2]), out_f32=y_f32, out_act=y_act, halo=0, mean=sv.mean2, rstd=sv.rstd2)
            else:
                sv.Fo = self._f32(rows, D)
                self._conv(sv.Hh, B, T, f"{pre}.pos_ffn.2.conv.weight", sv.Fo, c_bf16=False,
                           bias=self._P(f"{pre}.pos_ffn.2.conv.bias"))
                self._ln_fwd(B, T, D, sv.x1_f32, self._P(f"{pre}.norm2.norm.weight"), self._P(f"{pre}.norm2.norm.bias"), 1e-6,
                             branch=sv.Fo, drop_b=(p, sv.seeds[2]), out_f32=y_f32, out_act=y_act, mean=sv.mean2,
                             rstd=sv.rstd2)
            sv.y_f32 = y_f32
            saves.append(sv)
            x_f32, x_act = y_f32, y_act
            self._tr(f"{name}.layer{l}", y_f32, B, T, D)
        fin = _Saved()
        fin.x_f32 = x_f32
        fin.mean, fin.rstd = self._f32(rows), self._f32(rows)
        out_f32, out_act = self._f32(rows, D), self._act(rows, D)
        self._ln_fwd(B, T, D, x_f32, self._P(f"{name}.norm.norm.weight"), self._P(f"{name}.norm.norm.bias"), 1e-6,
                     lens=final_lens, out_f32=out_f32, out_act=out_act, halo=final_halo, mean=fin.mean, rstd=fin.rstd)
        fin.p = p
        return out_f32, out_act, saves, fin

    def _stack_bwd(self, cfg, saves, fin, dout, dout2, B, T, lens, final_lens, upto=0, st=None):
        """dout (+dout2): fp32 padded-row gradient wrt the stack output.  Returns (dx_a, dx_b): the gradient wrt
        the stack input is their sum.  `upto` > 0 stops after layer `upto` and returns the resumable state instead
        (pass it back as `st` to continue): the data-parallel step reduces the finished layers' gradients meanwhile."""
        D, F, H, nl = self.D, cfg["F"], cfg["H"], cfg["nl"]
        name = cfg["name"]
        hd = D // H
        rows = B * (T + 2 * PAD)
        TP = T + 2 * PAD
        ldk = _rup(T, 8)
        bf = self._bf16
        p = fin.p
        ld = 3 * D
        h1, h2 = (self.k0 - 1) // 2, (self.k1 - 1) // 2
        scale = 1.0 / math.sqrt(hd)
        fused = self._fused_attn(H)
        fuse_bias = D <= 384           # bias gradients of the out-proj / FFN-2 convs come out of the LN backward kernels
        if st is None:
            # scratch shared by all layers
            st = _Saved()
            st.dPd = None if fused else self._f32(B * H, T, ldk)
            st.dS = self._f32(B * H, _rup(T, 128)) if fused else self._act(B * H, T, ldk)    # fused: rowsum(dO * O) scratch
            st.dqkv = self._act(rows, ld)
            st.dF_act, st.dH_act = self._act(rows, D), self._act(rows, F)
            st.dHc = self._f32(rows, F) if h2 > 0 else None
            st.dz2, st.dz1, st.dXa = self._f32(rows, D), self._f32(rows, D), self._f32(rows, D)
            st.ksplit = self._dgrad_split(rows, D) if h1 > 0 else 1
            st.dX1c = self._f32(st.ksplit, rows, D)
            st.dProj_act, st.dO_act = self._act(rows, D), self._act(rows, D)
            st.dy_a, st.dy_b = self._f32(rows, D), None
            st.next = nl - 1
            self._ln_bwd(B, T, D, fin.x_f32, self._P(f"{name}.norm.norm.weight"), self._P(f"{name}.norm.norm.bias"), 1e-6,
                         fin.mean, fin.rstd, dy=dout, dy2=dout2, lens=final_lens, dx_f32=st.dy_a,
                         dgamma=self._G(f"{name}.norm.norm.weight"), dbeta=self._G(f"{name}.norm.norm.bias"))
        dPd, dS, dqkv, dF_act, dH_act, dHc = st.dPd, st.dS, st.dqkv, st.dF_act, st.dH_act, st.dHc
        dz2, dz1, dXa, ksplit, dX1c, dProj_act, dO_act = st.dz2, st.dz1, st.dXa, st.ksplit, st.dX1c, st.dProj_act, st.dO_act
        dy_a, dy_b = st.dy_a, st.dy_b
        for l in reversed(range(upto, st.next + 1)):
            pre = f"{name}.layers.{l}"
            sv = saves[l]
            # ---- LN2 + FFN
            # (a forward through fs2_gemm_ln_tc kept no branch: x_hat comes from the saved LayerNorm output instead)
            self._ln_bwd(B, T, D, sv.x1_f32 if sv.Fo is not None else None, self._P(f"{pre}.norm2.norm.weight"),
                         self._P(f"{pre}.norm2.norm.bias"), 1e-6,
                         sv.mean2, sv.rstd2, dy=dy_a, dy2=dy_b, branch=sv.Fo, y=sv.y_f32 if sv.Fo is None else None,
                         drop_b=(p, sv.seeds[2]),
                         dx_f32=dz2, dact=dF_act, dgamma=self._G(f"{pre}.norm2.norm.weight"),
                         dbeta=self._G(f"{pre}.norm2.norm.bias"),
                         dact_colsum=self._G(f"{pre}.pos_ffn.2.conv.bias") if fuse_bias else None)
            self._conv_wgrad(dF_act, sv.Hh, B, T, f"{pre}.pos_ffn.2.conv.weight", f"{pre}.pos_ffn.2.conv.weight",
                             None if fuse_bias else f"{pre}.pos_ffn.2.conv.bias")
            if h2 == 0:
                self._conv_dgrad(dF_act, B, T, f"{pre}.pos_ffn.2.conv.weight", dH_act, c_bf16=bf, relu_aux=sv.Hh)
            else:
                self._conv_dgrad(dF_act, B, T, f"{pre}.pos_ffn.2.conv.weight", dHc)
                raise NotImplementedError("fs2_b200: second FFN conv with kernel > 1 is not supported in backward")
            self._conv_wgrad(dH_act, sv.x1_act, B, T, f"{pre}.pos_ffn.0.conv.weight", f"{pre}.pos_ffn.0.conv.weight",
                             f"{pre}.pos_ffn.0.conv.bias")
            self._conv_dgrad(dH_act, B, T, f"{pre}.pos_ffn.0.conv.weight", dX1c, split=ksplit)
            # ---- LN1 + attention
            self._ln_bwd(B, T, D, sv.x_f32 if sv.proj is not None else None, self._P(f"{pre}.norm1.norm.weight"),
                         self._P(f"{pre}.norm1.norm.bias"), 1e-6,
                         sv.mean1, sv.rstd1, dy=dz2, dy2=dX1c[0], dy3=dX1c[1] if ksplit > 1 else None, dy2_fold=h1,
                         branch=sv.proj, y=sv.x1_f32 if sv.proj is None else None, drop_b=(p, sv.seeds[1]),
                         dx_f32=dz1, dact=dProj_act, dgamma=self._G(f"{pre}.norm1.norm.weight"),
                         dbeta=self._G(f"{pre}.norm1.norm.bias"),
                         dact_colsum=self._G(f"{pre}.self_att.att.out_proj.bias") if fuse_bias else None)
            self._conv_wgrad(dProj_act, sv.O, B, T, f"{pre}.self_att.att.out_proj.weight",
                             f"{pre}.self_att.att.out_proj.weight", None if fuse_bias else f"{pre}.self_att.att.out_proj.bias")
            self._conv_dgrad(dProj_act, B, T, f"{pre}.self_att.att.out_proj.weight", dO_act, c_bf16=bf)
            if not fused:
                # dPd = dO V^T
                L.gemm(mode=0, M=T, N=T, K=hd, A=dO_act, A_off=PAD * D, lda=D, a_rows=T, a_inner=hd, a_s1=hd, a_s2=TP * D,
                       B=sv.qkv, B_off=PAD * ld + 2 * D, ldb=ld, b_rows=T, b_inner=hd, b_s1=hd, b_s2=TP * ld,
                       batch1=H, batch2=B, Cout=dPd, ldc=ldk, c_s1=T * ldk, c_s2=H * T * ldk, c_bf16=False, ab_bf16=bf)
            self._wait_side(dqkv)
            if fused:
                # dQ kernel (recomputes P from the saved row statistic, dS stays in tensor memory) + dK/dV kernel
                L.call("fs2_flash_attn_bwd", dO_act, sv.O, sv.qkv, sv.P, lens, B, H, T, D, scale, p, sv.seeds[0], self._ctr,
                       dS, dqkv, 0)
            else:
                # dV = Pd^T dO
                L.gemm(mode=2, M=T, N=hd, K=T, A=sv.Pd, lda=ldk, a_rows=T, a_inner=T, a_s1=T * ldk, a_s2=H * T * ldk,
                       B=dO_act, B_off=PAD * D, ldb=D, b_rows=T, b_inner=hd, b_s1=hd, b_s2=TP * D, batch1=H, batch2=B,
                       Cout=dqkv, C_off=PAD * ld + 2 * D, ldc=ld, c_s1=hd, c_s2=TP * ld, c_bf16=bf, ab_bf16=bf)
                L.call("fs2_softmax_bwd", sv.P, dPd, lens, B, H, T, ldk, scale, p, sv.seeds[0], self._ctr, dS, int(bf))
                # dQ = dS K
                L.gemm(mode=1, M=T, N=hd, K=T, A=dS, lda=ldk, a_rows=T, a_inner=T, a_s1=T * ldk, a_s2=H * T * ldk,
                       B=sv.qkv, B_off=PAD * ld + D, ldb=ld, b_rows=T, b_inner=hd, b_s1=hd, b_s2=TP * ld, batch1=H, batch2=B,
                       Cout=dqkv, C_off=PAD * ld, ldc=ld, c_s1=hd, c_s2=TP * ld, c_bf16=bf, ab_bf16=bf)
                # dK = dS^T Q
                L.gemm(mode=2, M=T, N=hd, K=T, A=dS, lda=ldk, a_rows=T, a_inner=T, a_s1=T * ldk, a_s2=H * T * ldk,
                       B=sv.qkv, B_off=PAD * ld, ldb=ld, b_rows=T, b_inner=hd, b_s1=hd, b_s2=TP * ld, batch1=H, batch2=B,
                       Cout=dqkv, C_off=PAD * ld + D, ldc=ld, c_s1=hd, c_s2=TP * ld, c_bf16=bf, ab_bf16=bf)
            self._conv_wgrad(dqkv, sv.x_act, B, T, f"{pre}.self_att.att.in_proj_weight",
                             f"{pre}.self_att.att.in_proj_weight", f"{pre}.self_att.att.in_proj_bias")
            self._conv_dgrad(dqkv, B, T, f"{pre}.self_att.att.in_proj_weight", dXa)
            # gradient wrt this layer's input = dz1 + dXa ; ping-pong the buffers
            dy_a, dz1 = dz1, dy_a
            if dy_b is None:
                dy_b = self._f32(rows, D)
            dy_b, dXa = dXa, dy_b
        if upto > 0:
            st.dz1, st.dXa, st.dy_a, st.dy_b, st.next = dz1, dXa, dy_a, dy_b, upto - 1
            return st
        return dy_a, dy_b

    # ---------------------------------------------------------------- variance predictor
    def _pred_fwd(self, pname, x_act, B, T, lens, scale, base_seed, site, training, out):
        D = self.D
        rows = B * (T + 2 * PAD)
        h = (self.kd - 1) // 2
        p = self.var_p if training else 0.0
        sv = _Saved()
        sv.x_act = x_act
        sv.p = p
        sv.scale = float(scale)
        sv.seeds = [self._site_seed(base_seed, site), self._site_seed(base_seed, site + 1)]
        sv.h1 = self._f32(rows, D)
        self._conv(x_act, B, T, f"{pname}.conv1.conv.weight", sv.h1, c_bf16=False,
                   bias=self._P(f"{pname}.conv1.conv.bias"), relu=1)
        sv.a1 = self._act(rows, D)
        sv.mean1, sv.rstd1 = self._f32(rows), self._f32(rows)
        self._ln_fwd(B, T, D, sv.h1, self._P(f"{pname}.ln1.norm.weight"), self._P(f"{pname}.ln1.norm.bias"), 1e-5,
                     drop_a=(p, sv.seeds[0]), lens=lens, out_act=sv.a1, halo=h, mean=sv.mean1, rstd=sv.rstd1)
        sv.h2 = self._f32(rows, D)
        self._conv(sv.a1, B, T, f"{pname}.conv2.conv.weight", sv.h2, c_bf16=False,
                   bias=self._P(f"{pname}.conv2.conv.bias"), relu=1)
        sv.mean2, sv.rstd2 = self._f32(rows), self._f32(rows)
        self._ln_fwd(B, T, D, sv.h2, self._P(f"{pname}.ln2.norm.weight"), self._P(f"{pname}.ln2.norm.bias"), 1e-5,
                     drop_a=(p, sv.seeds[1]), lens=lens, mean=sv.mean2, rstd=sv.rstd2,
                     head=(self._P(f"{pname}.linear.w.weight"), self._P(f"{pname}.linear.w.bias"), out, sv.scale))
        return out, sv

    def _pred_bwd(self, pname, sv, dpred, B, T, lens, dx_out):
        """dpred (B,T) plain -> dx_out (fp32 padded rows incl. halo, to be folded with width kd//2 and masked)."""
        D = self.D
        rows = B * (T + 2 * PAD)
        h = (self.kd - 1) // 2
        p = sv.p
        d2_act, d1_act = self._act(rows, D), self._act(rows, D)
        dA1c = self._f32(rows, D)
        fuse_bias = D <= 384           # the conv bias gradients (column sums of dact) come out of the LN backward kernels
        self._ln_bwd(B, T, D, sv.h2, self._P(f"{pname}.ln2.norm.weight"), self._P(f"{pname}.ln2.norm.bias"), 1e-5,
                     sv.mean2, sv.rstd2, dhead=dpred, head_w=self._P(f"{pname}.linear.w.weight"), head_scale=sv.scale,
                     drop_a=(p, sv.seeds[1]), lens=lens, relu_x=1, dact=d2_act,
                     dgamma=self._G(f"{pname}.ln2.norm.weight"), dbeta=self._G(f"{pname}.ln2.norm.bias"),
                     dhead_w=self._G(f"{pname}.linear.w.weight"), dhead_b=self._G(f"{pname}.linear.w.bias"),
                     dact_colsum=self._G(f"{pname}.conv2.conv.bias") if fuse_bias else None)
        self._conv_wgrad(d2_act, sv.a1, B, T, f"{pname}.conv2.conv.weight", f"{pname}.conv2.conv.weight",
                         None if fuse_bias else f"{pname}.conv2.conv.bias")
        self._conv_dgrad(d2_act, B, T, f"{pname}.conv2.conv.weight", dA1c)
        self._ln_bwd(B, T, D, sv.h1, self._P(f"{pname}.ln1.norm.weight"), self._P(f"{pname}.ln1.norm.bias"), 1e-5,
                     sv.mean1, sv.rstd1, dy2=dA1c, dy2_fold=h, drop_a=(p, sv.seeds[0]), lens=lens, relu_x=1,
                     dact=d1_act, dgamma=self._G(f"{pname}.ln1.norm.weight"), dbeta=self._G(f"{pname}.ln1.norm.bias"),
                     dact_colsum=self._G(f"{pname}.conv1.conv.bias") if fuse_bias else None)
        self._conv_wgrad(d1_act, sv.x_act, B, T, f"{pname}.conv1.conv.weight", f"{pname}.conv1.conv.weight",
                         None if fuse_bias else f"{pname}.conv1.conv.bias")
        self._conv_dgrad(d1_act, B, T, f"{pname}.conv1.conv.weight", dx_out)

    def _exact_log_durations(self, tokens, speakers, intensity, out):
        """Encoder + conditioning + duration predictor on the fp32 path (model.py:331-372), into `out` (B, Tp)."""
        saved = (self.precision, self._arenas)
        if self._arenas_exact is None:
            self._arenas_exact = (Arena(torch.float32), Arena(torch.float32), Arena(torch.int32))
        try:
            self.precision = "fp32"
            self._arenas = self._arenas_exact
            for a in self._arenas:
                a.reset(tokens.device)
            self.store.sync_operands(False)
            B, Tp = tokens.shape
            D = self.D
            rowsP = B * (Tp + 2 * PAD)
            src_lens = self._i32(B)
            x_f32, x_act = self._f32(rowsP, D), self._act(rowsP, D)
            L.call("fs2_embed_posenc", tokens, self._P("encPreNet.token_embedding.Embedding.weight"),
                   self.sinusoidal_positional_embed_encoder.pe, B, Tp, D, self.padding_idx, x_f32, x_act, 0, src_lens)
            _, enc_act, _, _ = self._stack_fwd(self.enc, x_f32, x_act, B, Tp, src_lens, src_lens, 0, self._seed_base, 100,
                                               False, keep_p=False)
            G = self._f32(rowsP, D)
            self._conv(enc_act, B, Tp, "concat_proj.tok", G, c_bf16=False)
            c_f32, c_act = self._f32(rowsP, D), self._act(rowsP, D)
            L.call("fs2_cond_finish", G, self._P("concat_proj.w.weight"), self._P("speaker_emb.Embedding.weight"), speakers,
                   intensity, src_lens, B, Tp, D, self._f32(B, D), c_f32, c_act, 0, (self.kd - 1) // 2)
            self._pred_fwd("durPred", c_act, B, Tp, src_lens, 1.0, self._seed_base, 10, False, out)
        finally:
            self.precision, self._arenas = saved

    # --------------------------------------------------------------------------- forward
    def forward(self, tokens, speakers, durations=None, pitch=None, energy=None, pace=1.0, pitch_rate=1.0,
                energy_rate=1.0, intensity=None):
        """Same contract as reference model.py:279-441.  Returns (mel_post, postnet_output, predict_durations,
        predict_pitch, avg_pitch, predict_energy, avg_energy, mel_lens[CPU int64])."""
        if intensity is None:
            raise ValueError("intensity (B, Tp, 5) is required (reference model.py:356-358 concatenates it)")
        outs = _FS2Function.apply(self._anchor, self, tokens, speakers, durations, pitch, energy, float(pace),
                                  float(pitch_rate), float(energy_rate), intensity)
        outs = outs[:7]
        mel, post, pd, pp, ap, pe, ae = outs
        has_p, has_e = pitch is not None, energy is not None
        return (mel, post, pd, pp, ap if has_p else None, pe, ae if has_e else None, self._last_mel_lens_cpu)

    def forward_batch(self, texts, src_lens, mels, durations, pitches, energies, intensity, speakers=None, **control):
        """The call as BASELINE.json's north_star spells it -- forward(texts, src_lens, mels, durations, pitches, energies,
        intensity) -- mapped onto the reference's real signature: `texts` are the padded token ids (positions
        >= src_lens[b] are forced to the padding id, which is what the reference's collate guarantees, dataset.py:78-81),
        `mels` only fixes the frame count the targets were padded to (the model never reads target mels, model.py:279-290),
        `speakers` defaults to speaker 0.  Same return tuple as forward()."""
        tokens = texts.long()
        if src_lens is not None:
            pos = torch.arange(tokens.shape[1], device=tokens.device)[None]
            tokens = torch.where(pos < src_lens.to(tokens.device)[:, None], tokens, torch.full_like(tokens, self.padding_idx))
        if speakers is None:
            speakers = torch.zeros(tokens.shape[0], dtype=torch.long, device=tokens.device)
        if mels is not None and pitches is not None and mels.shape[1] != pitches.shape[1]:
            raise ValueError("forward_batch: mels and pitches are padded to different frame counts")
        return self.forward(tokens, speakers, durations, pitches, energies, intensity=intensity, **control)

    def _forward_impl(self, tokens, speakers, durations, pitch, energy, pace, pitch_rate, energy_rate, intensity,
                      Tm_known=None):
        st = self.store
        if not tokens.is_cuda:
            raise RuntimeError("fs2_b200: inputs must be CUDA tensors (there is no CPU fallback)")
        dev = tokens.device
        if st.flat.device != dev:
            raise RuntimeError("fs2_b200: model and inputs are on different devices")
        if not st.views_intact():
            st.reflatten()
        if self._arenas is None:
            self._arenas = (Arena(torch.float32), Arena(torch.bfloat16 if self._bf16 else torch.float32), Arena(torch.int32))
        for a in self._arenas:
            a.reset(dev)
        self._generation += 1
        if self._ctr is None or self._ctr.device != dev:
            self._ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        training = self.training
        if training:
            L.call("fs2_counter_add", self._ctr, 1)
        bf = self._bf16
        D, n_mels = self.D, self.n_mels
        tokens = tokens.contiguous().long()
        speakers = speakers.contiguous().long()
        intensity = intensity.contiguous().float()
        B, Tp = tokens.shape
        if Tp <= PAD:
            raise ValueError("fs2_b200: need more than 4 phoneme positions (reflect padding of the k=9 conv, as torch)")
        rowsP = B * (Tp + 2 * PAD)
        base_seed = self._seed_base
        ctx = _Saved()
        ctx.generation = self._generation
        ctx.B, ctx.Tp = B, Tp
        ctx.tokens, ctx.speakers, ctx.intensity = tokens, speakers, intensity
        ctx.teacher = durations is not None
        can_bwd = durations is not None and pitch is not None and energy is not None    # the training call (train.py:72)
        pe_enc = self.sinusoidal_positional_embed_encoder.pe
        pe_dec = self.sinusoidal_positional_embed_decoder.pe

        # ---- LengthRegulator scan first (teacher-forced): its only host read-back overlaps the encoder
        ends = self._i32(B, Tp)
        mel_lens = self._i32(B)
        ctx.ends, ctx.mel_lens = ends, mel_lens
        if self._pin_lens is None or self._pin_lens.numel() < B:
            self._pin_lens = torch.empty(max(B, 64), dtype=torch.int32).pin_memory()
        ev = None
        if durations is not None:
            durations = durations.contiguous().long()
            L.call("fs2_lr_prepare", durations, None, pace, B, Tp, ends, mel_lens)
            if Tm_known is None:
                self._pin_lens[:B].copy_(mel_lens, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            elif self.async_mel_lens:
                lens64, flag, pin64, pinflag = self._async_buffers(dev)
                L.call("fs2_lr_finalize", mel_lens, B, Tm_known, lens64, flag)
                pin64.copy_(lens64, non_blocking=True)
                pinflag.copy_(flag, non_blocking=True)
                self._last_mel_lens_cpu = pin64[:B]

        # ---- encoder (model.py:331-347)
        src_lens = self._i32(B)
        ctx.src_lens = src_lens
        x_f32, x_act = self._f32(rowsP, D), self._act(rowsP, D)
        L.call("fs2_embed_posenc", tokens, self._P("encPreNet.token_embedding.Embedding.weight"), pe_enc, B, Tp, D,
               self.padding_idx, x_f32, x_act, int(bf), src_lens)
        self._tr("enc_in", x_f32, B, Tp, D)
        enc_f32, enc_act, ctx.enc_saves, ctx.enc_fin = self._stack_fwd(
            self.enc, x_f32, x_act, B, Tp, src_lens, src_lens, 0, base_seed, 100, training, keep_p=can_bwd)
        ctx.enc_act = enc_act
        self._tr("enc_out", enc_f32, B, Tp, D)

        # ---- speaker + intensity conditioning (model.py:352-360)
        hv = (self.kd - 1) // 2
        G = self._f32(rowsP, D)
        self._conv(enc_act, B, Tp, "concat_proj.tok", G, c_bf16=False)
        c_f32, c_act = self._f32(rowsP, D), self._act(rowsP, D)
        L.call("fs2_cond_finish", G, self._P("concat_proj.w.weight"), self._P("speaker_emb.Embedding.weight"), speakers,
               intensity, src_lens, B, Tp, D, self._f32(B, D), c_f32, c_act, int(bf), hv)
        self._tr("cond", c_f32, B, Tp, D)

        # ---- variance adaptor (model.py:365-403)
        # the five (B, Tp) outputs share one buffer (one clone hands them out under graph replay, see _FS2Function)
        out5 = torch.zeros(5, B, Tp, device=dev, dtype=torch.float32)
        ctx.out5 = out5
        pred_dur, ctx.sv_dur = self._pred_fwd("durPred", c_act, B, Tp, src_lens, 1.0, base_seed, 10, training, out5[0])
        if durations is None and self.exact_durations and bf and not training:
            self._exact_log_durations(tokens, speakers, intensity, out5[0])       # overwrites the bf16 prediction
        pred_pitch, ctx.sv_pitch = self._pred_fwd("pitchPred", c_act, B, Tp, src_lens, pitch_rate, base_seed, 12, training,
                                                  out5[1])
        avg_pitch = out5[2]
        if pitch is not None:
            if durations is None:
                raise ValueError("pitch targets need durations (average_over_durations, model.py:383)")
            pitch = pitch.contiguous().float()
            L.call("fs2_avg_over_durations", pitch, durations, B, Tp, pitch.shape[1], avg_pitch, None, None, None)
            contour_p = avg_pitch
        else:
            contour_p = pred_pitch
        ctx.contour_p = contour_p
        ap_f32, ap_act = self._f32(rowsP, D), self._act(rowsP, D)
        L.call("fs2_embed_add", c_f32, contour_p, self._P("pitchEmbed.conv.weight"), self._P("pitchEmbed.conv.bias"),
               self.kp, src_lens, B, Tp, D, ap_f32, ap_act, int(bf), hv)
        self._tr("after_pitch", ap_f32, B, Tp, D)
        pred_energy, ctx.sv_energy = self._pred_fwd("energyPred", ap_act, B, Tp, src_lens, energy_rate, base_seed, 14, training,
                                                    out5[3])
        avg_energy = out5[4]
        if energy is not None:
            if durations is None:
                raise ValueError("energy targets need durations (average_over_durations, model.py:397)")
            energy = energy.contiguous().float()
            L.call("fs2_avg_over_durations", energy, durations, B, Tp, energy.shape[1], avg_energy, None, None, None)
            contour_e = avg_energy
        else:
            contour_e = pred_energy
        ctx.contour_e = contour_e
        ae_f32 = self._f32(rowsP, D)
        L.call("fs2_embed_add", ap_f32, contour_e, self._P("energyEmbed.conv.weight"), self._P("energyEmbed.conv.bias"),
               self.ke, src_lens, B, Tp, D, ae_f32, None, int(bf), 0)
        self._tr("after_energy", ae_f32, B, Tp, D)

        # ---- LengthRegulator (model.py:406-423)
        if durations is None:
            fdur = self._f32(B, Tp)
            L.call("fs2_dur_decode", pred_dur, B * Tp, fdur)
            L.call("fs2_lr_prepare", None, fdur, pace, B, Tp, ends, mel_lens)
            self._pin_lens[:B].copy_(mel_lens, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        if Tm_known is None:
            ev.synchronize()
            mel_lens_cpu = self._pin_lens[:B].to(torch.int64).clone()
            Tm = int(mel_lens_cpu.max())
            self._last_mel_lens_cpu = mel_lens_cpu
        else:
            Tm = Tm_known
        if Tm > 2500:
            raise ValueError("fs2_b200: more than 2500 frames (the positional table of the reference ends there)")
        if Tm <= PAD:
            raise ValueError("fs2_b200: need more than 4 mel frames (reflect padding of the k=9 conv, as torch)")
        ctx.Tm = Tm
        rowsM = B * (Tm + 2 * PAD)
        d0_f32, d0_act = self._f32(rowsM, D), self._act(rowsM, D)
        L.call("fs2_lr_expand", ae_f32, Tp + 2 * PAD, PAD, ends, mel_lens, pe_dec, B, Tp, Tm, D, d0_f32, d0_act, int(bf),
               Tm + 2 * PAD, PAD, None)
        self._tr("dec_in", d0_f32, B, Tm, D)

        # ---- decoder + mel projection + PostNet (model.py:425-431)
        dec_f32, dec_act, ctx.dec_saves, ctx.dec_fin = self._stack_fwd(
            self.dec, d0_f32, d0_act, B, Tm, mel_lens, None, 0, base_seed, 200, training, keep_p=can_bwd)
        ctx.dec_act = dec_act
        self._tr("dec_out", dec_f32, B, Tm, D)
        hp = (self.kpn - 1) // 2
        mel_raw = self._f32(rowsM, n_mels)
        self._conv(dec_act, B, Tm, "linear.w.weight", mel_raw, c_bf16=False, bias=self._P("linear.w.bias"))
        mel2 = torch.empty(2, B, Tm, n_mels, device=dev, dtype=torch.float32)      # mel_post, postnet_output
        ctx.mel2 = mel2
        mel_post = mel2[0]
        mel_f32, mel_act = self._f32(rowsM, n_mels), self._act(rowsM, n_mels)
        L.call("fs2_unpad_mask", mel_raw, mel_lens, B, Tm, n_mels, mel_post, mel_act, int(bf), hp)
        L.call("fs2_pad_rows", mel_post, None, B, Tm, n_mels, 1.0, mel_f32, None, int(bf))
        ctx.mel_act = mel_act
        pn = _Saved()
        ctx.pn = pn
        pn.p = self.pn_p if training else 0.0
        pn.seeds = [self._site_seed(base_seed, 20 + i) for i in range(3)]
        E = self.E
        pn.c1 = self._f32(rowsM, E)
        self._conv(mel_act, B, Tm, "postnet.conv_pre.conv.weight", pn.c1, c_bf16=False, bias=self._P("postnet.conv_pre.conv.bias"))
        pn.a1 = self._act(rowsM, E)
        pn.mean1, pn.rstd1 = self._f32(rowsM), self._f32(rowsM)
        self._ln_fwd(B, Tm, E, pn.c1, self._P("postnet.ln1.weight"), self._P("postnet.ln1.bias"), 1e-5, tanh=1,
                     drop_a=(pn.p, pn.seeds[0]), out_act=pn.a1, halo=hp, mean=pn.mean1, rstd=pn.rstd1)
        pn.mid = [pn.a1]
        cur = pn.a1
        n_mid = self.npn - 2
        pn.c_last = None
        for i in range(n_mid):
            last = i == n_mid - 1
            wname = f"postnet.convs_intermedite.{i}.conv"
            if last:
                pn.c_last = self._f32(rowsM, E)
                self._conv(cur, B, Tm, wname + ".weight", pn.c_last, c_bf16=False, bias=self._P(wname + ".bias"))
            else:
                nxt = self._act(rowsM, E)
                self._conv(cur, B, Tm, wname + ".weight", nxt, c_bf16=bf, bias=self._P(wname + ".bias"), halo=hp)
                pn.mid.append(nxt)
                cur = nxt
        if n_mid == 0:
            raise NotImplementedError("fs2_b200: postnet_n_convolutions must be >= 3")
        pn.a2 = self._act(rowsM, E)
        pn.mean2, pn.rstd2 = self._f32(rowsM), self._f32(rowsM)
        self._ln_fwd(B, Tm, E, pn.c_last, self._P("postnet.ln2.weight"), self._P("postnet.ln2.bias"), 1e-5, tanh=1,
                     drop_a=(pn.p, pn.seeds[1]), out_act=pn.a2, halo=hp, mean=pn.mean2, rstd=pn.rstd2)
        pn.c5 = self._f32(rowsM, n_mels)
        self._conv(pn.a2, B, Tm, "postnet.conv_post.conv.weight", pn.c5, c_bf16=False, bias=self._P("postnet.conv_post.conv.bias"))
        pn.mean3, pn.rstd3 = self._f32(rowsM), self._f32(rowsM)
        post_pad = self._f32(rowsM, n_mels)
        self._ln_fwd(B, Tm, n_mels, pn.c5, self._P("postnet.ln3.weight"), self._P("postnet.ln3.bias"), 1e-5,
                     drop_a=(pn.p, pn.seeds[2]), post_add=mel_f32, out_f32=post_pad, mean=pn.mean3, rstd=pn.rstd3)
        postnet_out = mel2[1]
        L.call("fs2_unpad_mask", post_pad, None, B, Tm, n_mels, postnet_out, None, int(bf), 0)
        ctx.pitch_rate, ctx.energy_rate = pitch_rate, energy_rate
        ctx.has_targets = (pitch is not None) and (energy is not None)
        outs = (mel_post, postnet_out, pred_dur, pred_pitch.unsqueeze(-1), avg_pitch.unsqueeze(-1),
                pred_energy.unsqueeze(-1), avg_energy.unsqueeze(-1))
        return outs, ctx

    # -------------------------------------------------------------------------- backward
    def _backward_impl(self, ctx, dmel, dpost, dpd, dpp, dpe, capturing=False):
        """Whole backward = part A (PostNet, mel linear, decoder) then part B (LengthRegulator, variance adaptor,
        conditioning, encoder).  After A the gradients of `decoder.* / linear.* / postnet.*` -- the tail
        [grad_split_lo, numel) of the flat gradient buffer -- are final; `grad_ready_hook(lo, hi)` lets the data-parallel
        step start their all-reduce while B still runs."""
        mid = self._backward_a(ctx, dmel, dpost, capturing)
        if not capturing:
            self._grad_ready()
        mid2 = self._backward_b(ctx, mid, dpd, dpp, dpe)
        for i in range(len(self._enc_cuts())):
            if not capturing:
                self._grad_ready_part(i)
            mid2 = self._backward_c(ctx, mid2, i)

    @property
    def grad_split_lo(self):
        """First element of the flat buffers that belongs to decoder.* / linear.* / postnet.* (they are the tail of
        the reference's state_dict order)."""
        if self._split_lo is None:
            tail = ("decoder.", "linear.", "postnet.")
            lo = min(o for k, (o, _) in self.store.offsets.items() if k.startswith(tail))
            if any(o >= lo and not k.startswith(tail) for k, (o, _) in self.store.offsets.items()):
                lo = self.store.flat.numel()         # unexpected layout: nothing is declared ready early
            self._split_lo = lo
        return self._split_lo

    def _grad_ready(self):
        if self.grad_ready_hook is not None:
            self.grad_ready_hook(self.grad_split_lo, self.store.flat.numel())

    def _enc_cuts(self):
        """Validated cut layers, descending, inside (0, n_layers): part B covers layers >= cuts[0], part C_i the layers
        [cuts[i+1], cuts[i]) and the last one [0, cuts[-1]) + the token embedding.  () = one part C with every layer."""
        nl = self.enc["nl"]
        cuts = sorted({int(c) for c in self.enc_grad_splits if 0 < int(c) < nl}, reverse=True)
        return cuts if cuts else [nl]

    def _layer_offset(self, l):
        """First flat-buffer element of encoder.layers.{l} (l == n_layers: of encoder.norm)."""
        pre = f"encoder.layers.{l}." if l < self.enc["nl"] else "encoder.norm."
        return min(o for key, (o, _) in self.store.offsets.items() if key.startswith(pre))

    def _grad_ready_part(self, i):
        """Called before part C_i starts: the layers >= cuts[i] (down to the previous cut) have final gradients."""
        if self.grad_ready_hook is None:
            return
        cuts = self._enc_cuts()
        if cuts[i] >= self.enc["nl"]:
            return
        hi = self.grad_split_lo if i == 0 else self._layer_offset(cuts[i - 1])
        lo = self._layer_offset(cuts[i])
        # the flat order is the state_dict order: encoder.layers.0 ... 5, encoder.norm, then decoder.* (checked once)
        if self._split_ok is None:
            offs = [self._layer_offset(l) for l in range(self.enc["nl"] + 1)]
            self._split_ok = offs == sorted(offs) and offs[-1] < self.grad_split_lo
        if self._split_ok and lo < hi:
            self.grad_ready_hook(lo, hi)

    def _backward_a(self, ctx, dmel, dpost, capturing=False):
        if not capturing and ctx.generation != self._generation:
            raise RuntimeError("fs2_b200: backward() of a forward whose workspace has been reused by a later forward; "
                               "call backward before the next forward of the same model")
        if not (ctx.teacher and ctx.has_targets):
            raise NotImplementedError("fs2_b200: backward needs teacher-forced durations, pitch and energy "
                                      "(the reference's training call, train.py:72)")
        st = self.store
        if not capturing:
            st.ensure_grads()
        bf = self._bf16
        D, n_mels, E = self.D, self.n_mels, self.E
        B, Tp, Tm = ctx.B, ctx.Tp, ctx.Tm
        rowsP, rowsM = B * (Tp + 2 * PAD), B * (Tm + 2 * PAD)
        hp = (self.kpn - 1) // 2
        hv = (self.kd - 1) // 2
        pn = ctx.pn
        dev = dmel.device
        dmel = dmel.contiguous().float()
        dpost = dpost.contiguous().float()
        self._side_begin(dev)

        # ---- PostNet (postnet_output = LN3-drop(conv_post(...)) + mel_post)
        dpost_pad = self._f32(rowsM, n_mels)
        L.call("fs2_pad_rows", dpost, None, B, Tm, n_mels, 1.0, dpost_pad, None, int(bf))
        d5_act = self._act(rowsM, n_mels)
        self._ln_bwd(B, Tm, n_mels, pn.c5, self._P("postnet.ln3.weight"), self._P("postnet.ln3.bias"), 1e-5, pn.mean3,
                     pn.rstd3, dy=dpost_pad, drop_a=(pn.p, pn.seeds[2]), dact=d5_act,
                     dgamma=self._G("postnet.ln3.weight"), dbeta=self._G("postnet.ln3.bias"))
        self._conv_wgrad(d5_act, pn.a2, B, Tm, "postnet.conv_post.conv.weight", "postnet.conv_post.conv.weight",
                         "postnet.conv_post.conv.bias")
        dE_c = self._f32(rowsM, E)
        self._conv_dgrad(d5_act, B, Tm, "postnet.conv_post.conv.weight", dE_c)
        dE_act = self._act(rowsM, E)
        self._ln_bwd(B, Tm, E, pn.c_last, self._P("postnet.ln2.weight"), self._P("postnet.ln2.bias"), 1e-5, pn.mean2,
                     pn.rstd2, dy2=dE_c, dy2_fold=hp, tanh=1, drop_a=(pn.p, pn.seeds[1]), dact=dE_act,
                     dgamma=self._G("postnet.ln2.weight"), dbeta=self._G("postnet.ln2.bias"))
        n_mid = self.npn - 2
        for i in reversed(range(n_mid)):
            wname = f"postnet.convs_intermedite.{i}.conv"
            self._conv_wgrad(dE_act, pn.mid[i], B, Tm, wname + ".weight", wname + ".weight", wname + ".bias")
            self._conv_dgrad(dE_act, B, Tm, wname + ".weight", dE_c)
            if i > 0:
                self._wait_side(dE_act)
                L.call("fs2_fold_halo", dE_c, B, Tm, E, hp, None, None, None, None, dE_act, int(bf))
        d1_act = self._act(rowsM, E)
        self._ln_bwd(B, Tm, E, pn.c1, self._P("postnet.ln1.weight"), self._P("postnet.ln1.bias"), 1e-5, pn.mean1,
                     pn.rstd1, dy2=dE_c, dy2_fold=hp, tanh=1, drop_a=(pn.p, pn.seeds[0]), dact=d1_act,
                     dgamma=self._G("postnet.ln1.weight"), dbeta=self._G("postnet.ln1.bias"))
        self._conv_wgrad(d1_act, ctx.mel_act, B, Tm, "postnet.conv_pre.conv.weight", "postnet.conv_pre.conv.weight",
                         "postnet.conv_pre.conv.bias")
        dmel_c = self._f32(rowsM, n_mels)
        self._conv_dgrad(d1_act, B, Tm, "postnet.conv_pre.conv.weight", dmel_c)
        # total gradient wrt mel_post (= linear(dec) * mask): conv_pre path + direct + postnet residual, masked
        dmel_pad = self._f32(rowsM, n_mels)
        L.call("fs2_pad_rows", dmel, dpost, B, Tm, n_mels, 1.0, dmel_pad, None, int(bf))
        dmelm_act = self._act(rowsM, n_mels)
        L.call("fs2_fold_halo", dmel_c, B, Tm, n_mels, hp, dmel_pad, None, ctx.mel_lens, None, dmelm_act, int(bf))
        self._conv_wgrad(dmelm_act, ctx.dec_act, B, Tm, "linear.w.weight", "linear.w.weight", "linear.w.bias")
        ddec = self._f32(rowsM, D)
        self._conv_dgrad(dmelm_act, B, Tm, "linear.w.weight", ddec)

        # ---- decoder
        da, db_ = self._stack_bwd(self.dec, ctx.dec_saves, ctx.dec_fin, ddec, None, B, Tm, ctx.mel_lens, None)
        self._side_join()        # decoder / linear / postnet gradients are complete on the main stream from here on
        return da, db_

    def _backward_b(self, ctx, mid, dpd, dpp, dpe):
        da, db_ = mid
        bf = self._bf16
        D = self.D
        B, Tp, Tm = ctx.B, ctx.Tp, ctx.Tm
        rowsP = B * (Tp + 2 * PAD)
        hv = (self.kd - 1) // 2
        self._side_begin(da.device)
        # ---- LengthRegulator: segment sums (rows f >= mel_len are never inside a segment -> mask is implicit)
        dAE = self._f32(rowsP, D)
        L.call("fs2_lr_bwd", da, db_, Tm + 2 * PAD, PAD, ctx.ends, ctx.mel_lens, B, Tp, Tm, D, dAE, Tp + 2 * PAD, PAD)
        # ---- energy embed + predictor
        L.call("fs2_embed_add_bwd", dAE, ctx.contour_e, self.ke, B, Tp, D, self._G("energyEmbed.conv.weight"),
               self._G("energyEmbed.conv.bias"))
        dxe = self._f32(rowsP, D)
        self._pred_bwd("energyPred", ctx.sv_energy, dpe.contiguous().float().view(B, Tp), B, Tp, ctx.src_lens, dxe)
        dAP = self._f32(rowsP, D)
        L.call("fs2_fold_halo", dxe, B, Tp, D, hv, None, None, ctx.src_lens, dAP, None, int(bf))
        L.call("fs2_add_", dAP, dAE, rowsP * D)
        # ---- pitch embed + pitch / duration predictors
        L.call("fs2_embed_add_bwd", dAP, ctx.contour_p, self.kp, B, Tp, D, self._G("pitchEmbed.conv.weight"),
               self._G("pitchEmbed.conv.bias"))
        dxp, dxd = self._f32(rowsP, D), self._f32(rowsP, D)
        self._pred_bwd("pitchPred", ctx.sv_pitch, dpp.contiguous().float().view(B, Tp), B, Tp, ctx.src_lens, dxp)
        self._pred_bwd("durPred", ctx.sv_dur, dpd.contiguous().float().view(B, Tp), B, Tp, ctx.src_lens, dxd)
        L.call("fs2_add_", dxp, dxd, rowsP * D)
        # dC = (dAP + fold(dxp + dxd)) * mask   (conditioning output is masked, model.py:360)
        dC_f32, dC_act = self._f32(rowsP, D), self._act(rowsP, D)
        L.call("fs2_fold_halo", dxp, B, Tp, D, hv, None, None, ctx.src_lens, dC_f32, None, int(bf))
        L.call("fs2_fold_halo", None, B, Tp, D, 0, dC_f32, dAP, ctx.src_lens, dC_f32, dC_act, int(bf))
        # ---- conditioning
        L.call("fs2_cond_bwd", dC_f32, self._P("concat_proj.w.weight"), self._P("speaker_emb.Embedding.weight"),
               ctx.speakers, ctx.intensity, B, Tp, D, self._f32(B, D), self._G("concat_proj.w.weight"),
               self._G("speaker_emb.Embedding.weight"))
        self._conv_wgrad(dC_act, ctx.enc_act, B, Tp, "concat_proj.tok", "concat_proj.w.weight", None)
        denc = self._f32(rowsP, D)
        self._conv_dgrad(dC_act, B, Tp, "concat_proj.tok", denc)
        # ---- encoder, upper layers: their gradients (and encoder.norm's) are final when this part ends
        k = self._enc_cuts()[0]
        est = self._stack_bwd(self.enc, ctx.enc_saves, ctx.enc_fin, denc, None, B, Tp, ctx.src_lens, ctx.src_lens, upto=k) \
            if k < self.enc["nl"] else None
        self._side_join()
        return est, denc

    def _backward_c(self, ctx, mid, i):
        """Part C_i of the backward: the next group of encoder layers; the last one ends with the token embedding."""
        est, denc = mid
        cuts = self._enc_cuts()
        B, Tp, D = ctx.B, ctx.Tp, self.D
        rowsP = B * (Tp + 2 * PAD)
        last = i == len(cuts) - 1
        upto = 0 if last else cuts[i + 1]
        self._side_begin(denc.device)
        out = self._stack_bwd(self.enc, ctx.enc_saves, ctx.enc_fin, denc, None, B, Tp, ctx.src_lens, ctx.src_lens, upto=upto, st=est)
        if last:
            ea, eb = out
            if eb is not None:
                L.call("fs2_add_", ea, eb, rowsP * D)
            L.call("fs2_embedding_bwd", ea, ctx.tokens, B, Tp, D, self.padding_idx,
                   self._G("encPreNet.token_embedding.Embedding.weight"))
        self._side_join()
        return (out, denc) if not last else None

    # ---------------------------------------------------------------------- CUDA graphs
    def _graph_eligible(self, durations, pitch, energy):
        return self.use_cuda_graphs and durations is not None and pitch is not None and energy is not None \
            and self.trace is None

    def _async_buffers(self, dev):
        if self._async_bufs is None or self._async_bufs[0].device != dev:
            self._async_bufs = (torch.zeros(256, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int32, device=dev),
                                torch.zeros(256, dtype=torch.int64).pin_memory(), torch.zeros(1, dtype=torch.int32).pin_memory())
        return self._async_bufs

    def _check_async_flag(self):
        if self._async_bufs is not None and int(self._async_bufs[3][0]) != 0:
            got = int(self._async_bufs[3][0])
            self._async_bufs[1].zero_()
            self._async_bufs[3].zero_()
            raise RuntimeError(f"fs2_b200: async_mel_lens: a previous batch had max(sum(durations)) = {got} frames, which is "
                               "not pitch.shape[1]; its outputs were computed on the wrong rectangle")

    def _probe_Tm(self, durations, pace, B, Tp):
        """Eager duration scan + the step's single host read-back (mel_lens)."""
        dev = durations.device
        if self._pre is None or self._pre[0].numel() < B * Tp or self._pre[0].device != dev:
            self._pre = (torch.zeros(max(B * Tp, 64 * 128), dtype=torch.int32, device=dev),
                         torch.zeros(max(B, 64), dtype=torch.int32, device=dev))
        if self._pin_lens is None or self._pin_lens.numel() < B:
            self._pin_lens = torch.empty(max(B, 64), dtype=torch.int32).pin_memory()
        L.call("fs2_lr_prepare", durations, None, pace, B, Tp, self._pre[0], self._pre[1])
        self._pin_lens[:B].copy_(self._pre[1][:B], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        mel_lens_cpu = self._pin_lens[:B].to(torch.int64).clone()
        self._last_mel_lens_cpu = mel_lens_cpu
        return int(mel_lens_cpu.max())

    def _graph_forward(self, tokens, speakers, durations, pitch, energy, pace, pitch_rate, energy_rate, intensity):
        """Returns (outs, ctx, entry).  First sight of a shape runs eagerly (sizes the arenas, configures the
        kernels); the second captures forward and backward graphs; later steps copy the inputs into the static
        slots and replay."""
        tokens = tokens.contiguous().long()
        durations = durations.contiguous().long()
        B, Tp = tokens.shape
        self._sync_operands()               # eager, outside the captured graphs: usually no launch at all (params.py)
        if self.async_mel_lens and B <= 256:
            self._check_async_flag()
            Tm = int(pitch.shape[1])
        else:
            Tm = self._probe_Tm(durations, pace, B, Tp)
        key = (B, Tp, Tm, pitch.shape[1], energy.shape[1], self.training, self.precision, pace, pitch_rate, energy_rate)
        ins = (tokens, speakers.contiguous().long(), durations, pitch.contiguous().float(), energy.contiguous().float(),
               intensity.contiguous().float())
        entry = self._graphs.get(key)
        if entry is None:
            if key not in self._seen:
                self._seen.add(key)
                outs, ctx = self._forward_impl(tokens, speakers, durations, pitch, energy, pace, pitch_rate, energy_rate,
                                               intensity, Tm_known=Tm)
                return outs, ctx, None
            entry = self._capture(key, ins, Tm, pace, pitch_rate, energy_rate)
        else:
            for dst, src in zip(entry.static_in, ins):
                dst.copy_(src, non_blocking=True)
        self._generation += 1
        entry.ctx.generation = self._generation
        entry.fwd.replay()
        self.replayed_launches += entry.n_fwd
        return entry.outs, entry.ctx, entry

    def _capture(self, key, ins, Tm, pace, pitch_rate, energy_rate):
        st = self.store
        st.ensure_grads()
        assert self._ctr is not None and self._arenas is not None      # an eager step of this shape came first
        for a in self._arenas:
            a.reset(ins[0].device)          # any regrowth happens here, outside the capture
        entry = _Saved()
        entry.static_in = [t.clone() for t in ins]
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(g):
            outs, ctx = self._forward_impl(*entry.static_in[:3], entry.static_in[3], entry.static_in[4], pace, pitch_rate,
                                           energy_rate, entry.static_in[5], Tm_known=Tm)
        entry.fwd, entry.outs, entry.ctx = g, outs, ctx
        entry.n_fwd = L.launch_count() - n0
        n0 = L.launch_count()
        entry.arena_bufs = [a.buf for a in self._arenas]
        entry.grad_in = [torch.zeros_like(outs[i]) for i in (0, 1, 2, 3, 5)]
        ga = torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga):
            mid = self._backward_a(ctx, entry.grad_in[0], entry.grad_in[1], capturing=True)
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb):
            mid2 = self._backward_b(ctx, mid, *entry.grad_in[2:])
        entry.bwd_c = []
        for i in range(len(self._enc_cuts())):
            gc = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gc):
                mid2 = self._backward_c(ctx, mid2, i)
            entry.bwd_c.append(gc)
        entry.bwd_a, entry.bwd_b = ga, gb
        entry.n_bwd = L.launch_count() - n0
        self._graphs[key] = entry
        return entry


class _FS2Function(torch.autograd.Function):
    @staticmethod
    def forward(fctx, anchor, model, tokens, speakers, durations, pitch, energy, pace, pitch_rate, energy_rate, intensity):
        entry = None
        if model._graph_eligible(durations, pitch, energy):
            outs, ctx, entry = model._graph_forward(tokens, speakers, durations, pitch, energy, pace, pitch_rate,
                                                    energy_rate, intensity)
        else:
            tm = None
            model._sync_operands()
            if model.async_mel_lens and durations is not None and pitch is not None and tokens.shape[0] <= 256:
                model._check_async_flag()
                tm = int(pitch.shape[1])
            outs, ctx = model._forward_impl(tokens, speakers, durations, pitch, energy, pace, pitch_rate, energy_rate,
                                            intensity, Tm_known=tm)
        fctx.model, fctx.ctx, fctx.entry = model, ctx, entry
        if entry is not None:
            # graph replays write into static buffers; the reference hands out fresh tensors every call and train.py:87-90
            # still reads `predictions` after later steps, so the outputs are copied out (two launches, 16 MB at B = 32)
            m2, o5 = ctx.mel2.clone(), ctx.out5.clone()
            outs = (m2[0], m2[1], o5[0], o5[1].unsqueeze(-1), o5[2].unsqueeze(-1), o5[3].unsqueeze(-1), o5[4].unsqueeze(-1))
        fctx.mark_non_differentiable(outs[4], outs[6])
        return outs

    @staticmethod
    def backward(fctx, dmel, dpost, dpd, dpp, dap, dpe, dae):
        model, entry = fctx.model, fctx.entry
        if entry is None:
            model._backward_impl(fctx.ctx, dmel, dpost, dpd, dpp, dpe)
        else:
            if fctx.ctx.generation != model._generation:
                raise RuntimeError("fs2_b200: backward() of a forward whose workspace has been reused by a later forward")
            model.store.ensure_grads()
            for dst, src in zip(entry.grad_in, (dmel, dpost, dpd, dpp, dpe)):
                dst.copy_(src.reshape(dst.shape), non_blocking=True)
            entry.bwd_a.replay()
            model._grad_ready()
            entry.bwd_b.replay()
            for i, gc in enumerate(entry.bwd_c):
                model._grad_ready_part(i)
                gc.replay()
            model.replayed_launches += entry.n_bwd
        return (torch.zeros(1, device=dmel.device),) + (None,) * 10
