/* fs2_b200 -- C ABI of the B200-native FastSpeech2 hot path.
 *
 * The reference (Orca0917/fine-grained-emotional-control-of-tts) has no FFI: its
 * "operator API" is the Python class contract of emo_rank_tts/fastspeech2/model.py
 * (FastSpeech2.forward, :279-441) and loss.py (Loss.forward, :62-186), whose
 * arithmetic is stock torch ops reached through speechbrain.  Each entry point below
 * replaces one group of those torch-op call sites (cited per function) and is what a
 * ctypes binding on the reference side would bind.
 *
 * Conventions: all pointers are DEVICE pointers unless named h_*; `stream` is a
 * cudaStream_t passed as void*; every function returns 0 on success or a non-zero
 * FS2_ERR_* code (fs2_last_error() gives the message).  Nothing here allocates
 * persistent device memory; callers own every buffer.  No torch types.
 *
 * Activation layout ("padded row space"): a (B, T, C) activation is stored as
 * [B*(T+8), C]; row (b,t) is at b*(T+8)+4+t.  The 4 rows either side of each item hold
 * the reflect halo (conv inputs) or zeros (gradients).  `act_bf16` selects the storage
 * type of GEMM-operand activations: 1 = bf16 (tensor-core path), 0 = fp32 (exact path).
 */
#ifndef FS2_B200_H
#define FS2_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FS2_ABI_VERSION 1

const char* fs2_last_error(void);
int fs2_abi_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long fs2_launch_count(void);

/* ---------------------------------------------------------------- GEMM family --
 * One descriptor drives both the tcgen05/TMA tensor-core kernel (bf16 operands) and the
 * exact SIMT kernel (fp32 or bf16 operands).  Replaces: nn.Linear / nn.Conv1d /
 * baddbmm / bmm call sites of speechbrain's Conv1d, Linear and nn.MultiheadAttention
 * (reference model.py:199-276 constructors; forward model.py:344-346, 359, 425-431) and
 * their autograd backward (train.py:80).
 * mode 0: C[m,n] = sum_{j<taps} sum_{k<K} A[a_row_off+m+j*a_tap_step, k] * B[n, j*b_tap_step+k]
 * mode 1: C[m,n] = sum_j sum_k A[a_row_off+m+j*a_tap_step, k] * B[b_row_off+k, n+j*b_tap_step]
 * mode 2: C[m, j*c_tap_stride+n] = sum_{k<K} A[a_row_off+k, m] * B[b_row_off+k+j*b_tap_step, n]
 * Out-of-range operand coordinates read as zero.
 */
typedef struct {
  int mode, M, N, K, taps, batch1, batch2;
  const void* A; long long lda, a_s1, a_s2; int a_rows, a_inner, a_row_off, a_tap_step;
  const void* B; long long ldb, b_s1, b_s2; int b_rows, b_inner, b_row_off, b_tap_step;
  long long c_tap_stride;
  long long c_col_stride;  /* mode 2 only: element stride between consecutive n (0/1 = contiguous) */
  void* C; int c_bf16; long long ldc, c_s1, c_s2; int c_row_off, c_col_off;
  int accumulate;          /* 1: atomically add into fp32 C */
  int split_k;             /* mode 2 only; >1 implies accumulate */
  const float* bias; float alpha; int relu;
  const void* relu_aux;    /* optional, addressed like C: out *= (aux > 0) */
  int aux_bf16;
  int rs_T, rs_Tp;         /* row space of the output rows (rs_Tp = 0: every row valid) */
  const int* lens;         /* optional per-item valid length: rows t >= lens[b] -> 0 */
  int halo;                /* >0: mirror-write the reflect halo of this width */
  int ab_bf16;             /* operand storage: 1 = bf16, 0 = fp32 (SIMT only) */
  long long c_split_stride; /* modes 0/1, fs2_gemm_tc: with split_k = s > 1, split i of the reduction is STORED (no atomics) to
                              C + i * c_split_stride -- s partial results the consumer adds up (deterministic split-K for
                              short grids; fp32 C, plain epilogue).  0: not used. */
  float* a_colsum;          /* fs2_gemm_tc, mode 2 only, optional: a_colsum[m] += alpha * sum_k A[a_row_off+k, m] -- the bias gradient
                              that belongs to a weight gradient (db = column sums of dy).  Folded into the GEMM where the kernel
                              has room for it (128 x 192 tiles: extra N = 16 MMAs against a shared-memory tile of ones, the
                              k-blocks shared out over the column tiles of a row block); otherwise the library launches its
                              column-sum kernel. */
} Fs2Gemm;

int fs2_gemm_simt(const Fs2Gemm* g, void* stream);
int fs2_gemm_tc(const Fs2Gemm* g, void* stream);     /* tcgen05 + TMA, bf16 operands */
/* device-side error word of the tensor-core kernel (non-zero: an mbarrier wait timed out) */
int fs2_gemm_tc_error_flag(void);
/* kernel selection override for measurements: pair = 0 single-CTA tiles, 1 CTA-pair (cta_group::2) tiles, 2 built-in
 * heuristic (default); cfg = -1 automatic tile width, or 0|1|2 = 256|384(192)|128 columns */
int fs2_gemm_tc_tune(int pair, int cfg);
/* measurement hook: when non-NULL, every CTA-pair GEMM launch writes {SM cycles, nanoseconds} of its first CTA's
 * lifetime to dev_buf[0..1] (the SM clock the kernel really ran at = cycles / ns) */
int fs2_gemm_tc_set_debug(long long* dev_buf);

/* ------------------------------------------------------------ row-wise kernels -- */
/* Rule for every producer of a padded-row tensor: rect rows get values; halo rows get the reflect mirror
 * (width `halo`, needs T > halo) and zeros otherwise. */

/* model.py:331-337  key-padding mask + encPreNet embedding + sinusoidal pos-enc + mask.
 * tokens (B,Tp) i64; emb (V,D) f32; pe (>=Tp, D) f32.  src_lens[b] = number of non-pad tokens (the later
 * masks assume padding is a suffix, as the reference collate produces: dataset.py:76-80). */
int fs2_embed_posenc(const int64_t* tokens, const float* emb, const float* pe, int B, int Tp, int D,
                     int pad_idx, float* out_f32, void* out_act, int act_bf16, int* src_lens, void* stream);
int fs2_embedding_bwd(const float* dx /*padded rows*/, const int64_t* tokens, int B, int Tp, int D, int pad_idx,
                      float* demb /* += */, void* stream);

/* LayerNorm family (speechbrain LayerNorm / nn.LayerNorm call sites: TransformerEncoderLayer norm1/norm2,
 * TransformerEncoder.norm, DurationPredictor ln1/ln2 + linear head, PostNet ln1-3).
 *   z = x + drop_b(branch);  u = LN(z)*gamma+beta;  v = tanh?(u);  w = drop_a(v);
 *   out = w * rowmask? (+ post_add);   optional head: head_out[b,t] = (dot(out, head_w)+head_b)*head_scale */
typedef struct {
  int B, T, C;
  const float* x; const float* branch;     /* padded row space fp32; branch may be NULL */
  float drop_b_p; unsigned long long drop_b_seed;
  const float* gamma; const float* beta; float eps;
  int tanh_act;
  float drop_a_p; unsigned long long drop_a_seed;
  const int* lens;                          /* NULL: no row mask */
  const float* post_add;                    /* NULL or fp32 padded-row tensor added after everything */
  float* out_f32; void* out_act; int act_bf16; int halo;
  float* mean; float* rstd;                 /* [rows] saved statistics (may be NULL) */
  const float* head_w; const float* head_b; float* head_out; float head_scale; /* head_out is (B,T) plain */
  const unsigned long long* seed_dev;       /* optional device counter mixed into both dropout seeds (CUDA-graph replay) */
} Fs2LnFwd;
int fs2_ln_fwd(const Fs2LnFwd* p, void* stream);

/* Backward of the above.  Gradient wrt `out` of row (b,t) = dy + reflect-fold_p(dy2) + dhead*head_scale*head_w.
 * Outputs: dx_f32 = dz (grad wrt z; multiplied by (x>0) when relu_x, for x = relu(conv));
 *          dact   = drop_b-backward(dz) when branch != NULL else dz, in operand storage, zero halo rows. */
typedef struct {
  int B, T, C;
  const float* dy; const float* dy2; int dy2_fold;   /* both optional; dy2_fold=p>0: dy2 is a conv-dgrad output whose halo rows fold back */
  const float* dhead; const float* head_w; float head_scale;
  const float* x; const float* branch;
  float drop_b_p; unsigned long long drop_b_seed;
  const float* gamma; const float* beta; float eps; int tanh_act;
  float drop_a_p; unsigned long long drop_a_seed;
  const int* lens;
  const float* mean; const float* rstd;
  int relu_x;
  float* dx_f32; void* dact; int act_bf16;
  float* dgamma; float* dbeta; float* dhead_w; float* dhead_b;   /* accumulated (+=) */
  const unsigned long long* seed_dev;
  float* dact_colsum;   /* optional [C], accumulated (+=): column sums of `dact` over all rows = the bias gradient of the
                           GEMM that produced `branch` (saves the separate fs2_colsum pass over dact); C <= 384 only */
  const float* dy3;     /* optional second half of dy2 (same layout, same fold): the other partial result of a split-K dgrad */
  const float* y;       /* optional: the forward's fp32 OUTPUT (padded rows).  When set, the normalised value is taken from it,
                           x_hat = (y - beta) / gamma (0 where gamma == 0), and x / branch / mean are not read -- for forwards
                           that never materialise the branch (fs2_gemm_ln_tc).  Plain LayerNorm only: no tanh, dropout-after,
                           row mask, post_add or head.  drop_b_p still applies to `dact` (the branch gradient). */
} Fs2LnBwd;
int fs2_ln_bwd(const Fs2LnBwd* p, void* stream);
/* measurement hook: the LayerNorm kernels ask L2 for a warp's next row while the current one is reduced; `prefetch` =
 * distance in rows of a warp, 0 (off) .. 4, default 1 (measured best on B200; FS2_LN_PREFETCH sets it at load time) */
int fs2_ln_tune(int prefetch);

/* GEMM + bias + branch dropout + residual + LayerNorm in ONE tcgen05 kernel (gemm_ln.cu): the two places per FFT block
 * where speechbrain's TransformerEncoderLayer runs  LN(x + dropout(Linear/Conv1d-k1(a)))  -- out-projection -> norm1 and
 * FFN conv 2 -> norm2 (reference model.py:344-347, 425-428 through speechbrain's post-norm encoder layer).  Equivalent to
 * fs2_gemm_tc (mode 0, taps 1, fp32 C = branch) followed by fs2_ln_fwd (x, branch, drop_b, out_f32, out_act bf16, halo,
 * mean, rstd), but the fp32 branch never reaches HBM: a 128 x 384 tile holds full rows in tensor memory, the epilogue warps
 * add bias / dropout / residual, take the row statistics in two more passes over tensor memory and store fp32 + bf16.
 *   A (M, K) bf16, M = B*(T+8) padded rows;  W (384, K) bf16 K-major;  x, out_f32 fp32 (M, 384);  out_act bf16 (M, 384).
 * N is fixed to 384 (the model width).  Rows outside the rectangle follow the producer rule of the row-wise kernels. */
typedef struct {
  int B, T;                 /* row space: M = B * (T + 8) */
  int K; long long lda, ldw;
  const void* A; const void* W; const float* bias;
  const float* x;           /* residual */
  float drop_p; unsigned long long drop_seed; const unsigned long long* seed_dev;
  const float* gamma; const float* beta; float eps;
  float* out_f32; void* out_act; int halo;
  float* mean; float* rstd; /* optional [M] */
} Fs2GemmLn;
int fs2_gemm_ln_tc(const Fs2GemmLn* g, void* stream);
/* measurement hook: when non-NULL, CTA 0 of every fs2_gemm_ln_tc launch writes the cycle breakdown of its first tile's
 * epilogue (library built with -DFS2_TC_PROBE only; tools/gemm_ln_bench.py GEMM_LN_DBG=1) */
int fs2_gemm_ln_set_debug(long long* dev_buf);

/* softmax over keys with the reference's attn_mask quirk (model.py:338-343, 414-419):
 * keys valid for (b,h) are [0, min(len[b], len[(b*H+h) % B])).  S (B*H, T, ldk) fp32 -> P (and Pd = dropout(P)
 * when drop_p > 0).  P = softmax(scale*S); columns >= kv are written as 0. */
int fs2_softmax_fwd(const float* S, const int* lens, int B, int H, int T, int ldk, float scale,
                    float drop_p, unsigned long long seed, const unsigned long long* seed_dev, void* P, void* Pd,
                    int act_bf16, void* stream);
int fs2_softmax_bwd(const void* P, const float* dPd, const int* lens, int B, int H, int T, int ldk, float scale,
                    float drop_p, unsigned long long seed, const unsigned long long* seed_dev, void* dS, int act_bf16,
                    void* stream);
/* Flash-style attention (flash_attention.cu; reference model.py:344-346, 425-427 through nn.MultiheadAttention's math
 * path): nothing of size T x T reaches HBM.  qkv (B*(T+8), 3D) bf16 padded rows [Q | K | V], head_dim 192.
 * forward: O (B*(T+8), D) bf16, rows t < T written; lse (B*H, fs2_flash_attn_lse_len(T)) fp32 = per query row
 * max + log2(sum 2^(s - max)) of the scaled scores in the log2 domain (NULL: inference, not kept).
 * Dropout on the probabilities: one 32-bit counter hash per (item, head, query, key), regenerated by the backward
 * (fs2_flash_attn_mask returns the keep mask as bytes (B*H, T, T) for tests).  plain_mask = 1: an ordinary
 * key-padding mask (keys [0, lens[b]) for every head) as nn.MultiheadAttention(key_padding_mask=...) in the intensity
 * extractor (rank_model/model.py:34, 103); 0: FastSpeech2's attn_mask quirk, keys [0, min(len[b], len[(b*H+h) % B])).
 * backward: dQ, dK, dV into columns [0,D), [D,2D), [2D,3D) of dqkv (B*(T+8), 3D) bf16, rows t < T written;
 * a row-dot pre-pass (dvec = scale * rowsum(dO*O), same shape as lse, scratch) and ONE launch whose grid holds the dK/dV
 * tiles followed by the dQ tiles (two separate launches would each end in a nearly empty last wave). */
int fs2_flash_attn_lse_len(int T);
int fs2_flash_attn_fwd(const void* qkv, const int* lens, int B, int H, int T, int D, float scale, float drop_p,
                       unsigned long long seed, const unsigned long long* seed_dev, float* lse, void* O, int plain_mask,
                       void* stream);
int fs2_flash_attn_bwd(const void* dO, const void* O, const void* qkv, const float* lse, const int* lens, int B, int H,
                       int T, int D, float scale, float drop_p, unsigned long long seed,
                       const unsigned long long* seed_dev, float* dvec, void* dqkv, int plain_mask, void* stream);
int fs2_flash_attn_mask(int BH, int T, float drop_p, unsigned long long seed, const unsigned long long* seed_dev,
                        unsigned char* keep, void* stream);
/* measurement hook: 1 (default) = probabilities go to the second contraction through tensor memory, 0 = through shared memory */
int fs2_flash_attn_tune(int p_in_tmem);
/* measurement hook: when non-NULL, CTA 0 of every flash-attention launch writes a 64-slot cycle breakdown (library built
 * with -DFS2_TC_PROBE only; tools/flash_bench.py FLASH_DBG=1) */
int fs2_flash_attn_set_debug(long long* dev_buf);
/* *ctr += inc (the device-side dropout step counter; first node of a captured forward graph) */
int fs2_counter_add(unsigned long long* ctr, unsigned long long inc, void* stream);

/* model.py:352-360 speaker/intensity conditioning, split-weight form (no cat):
 * y = (G + Ws.spk_emb[spk[b]] + Wi.intensity[b,t]) * mask ; G = token_feats.Wt^T (a GEMM).
 * Wcat is concat_proj.w.weight (D, 2D+5).  sp_ws: B*D floats scratch. */
int fs2_cond_finish(const float* G, const float* Wcat, const float* spk_emb, const int64_t* speakers,
                    const float* intensity /*(B,Tp,5)*/, const int* lens, int B, int Tp, int D, float* sp_ws,
                    float* y_f32, void* y_act, int act_bf16, int halo, void* stream);
/* dy: padded rows fp32, already masked.  Accumulates dWcat[:, D:2D+5] and dspk_emb.  dsum_ws: B*D floats. */
int fs2_cond_bwd(const float* dy, const float* Wcat, const float* spk_emb, const int64_t* speakers,
                 const float* intensity, int B, int Tp, int D, float* dsum_ws, float* dWcat, float* dspk_emb,
                 void* stream);

/* speechbrain average_over_durations (model.py:383, 397): mean over non-zero frames per phoneme, computed as
 * differences of fp32 prefix sums (torch CPU cumsum semantics: double accumulator, fp32 prefixes).
 * values (B,Tm) f32; durs (B,Tp) i64 -> avg (B,Tp) f32; optional starts/ends (B,Tp) i32 and nz counts. */
int fs2_avg_over_durations(const float* values, const int64_t* durs, int B, int Tp, int Tm,
                           float* avg, int* starts, int* ends, int* nz, void* stream);

/* pitchEmbed / energyEmbed: Conv1d(1->D,k,reflect) over the phoneme axis + add (model.py:384-389, 398-403).
 * contour (B,Tp) f32 plain.  y_f32 is unmasked (as the reference); y_act = masked copy (next predictor's input). */
int fs2_embed_add(const float* x, const float* contour, const float* w /*(D,1,k)*/, const float* bias, int ksize,
                  const int* lens, int B, int Tp, int D, float* y_f32, void* y_act, int act_bf16, int halo, void* stream);
int fs2_embed_add_bwd(const float* dy /*padded rows*/, const float* contour, int ksize, int B, int Tp, int D,
                      float* dw, float* dbias, void* stream);

/* LengthRegulator = speechbrain upsample (model.py:406-410) + get_mask_from_lengths + decoder pos-enc + mask
 * (model.py:411-423).
 * fs2_dur_decode: fdur = clamp(expm1(log_dur), 0)                                    (model.py:372-375)
 * fs2_lr_prepare: frames[b,p] = (long)(pace * dur[b,p]) (fdur != NULL: float durations, inference),
 *                 ends = inclusive cumsum (i32), mel_lens[b].
 * fs2_lr_expand:  out[b,f,:] = in[b, idx(f), :] + pe[f,:] for f < mel_lens[b], else 0; frame2ph[b,f] = idx(f)
 *                 or -1.  Row pitch + offset let one kernel serve plain (B,T,D) tensors (pitch=T, off=0) and
 *                 the padded row space (pitch=T+8, off=4). */
int fs2_dur_decode(const float* log_dur, long long n, float* fdur, void* stream);
int fs2_lr_prepare(const int64_t* dur, const float* fdur, float pace, int B, int Tp, int* ends, int* mel_lens, void* stream);
/* sync-free variant: the caller states Tm (= the padded frame axis of the batch); out_i64 = mel_lens as int64 for an
 * asynchronous host copy, *flag is set (to the real maximum) when max(mel_lens) != Tm_expected */
int fs2_lr_finalize(const int* mel_lens, int B, int Tm_expected, long long* out_i64, int* flag, void* stream);
int fs2_lr_expand(const float* in, int in_pitch, int in_off, const int* ends, const int* mel_lens, const float* pe,
                  int B, int Tp, int Tm, int D, float* out_f32, void* out_act, int act_bf16,
                  int out_pitch, int out_off, int* frame2ph, void* stream);
/* backward: dphon[b,p,:] = sum over the phoneme's frames of (dframes + dframes2).  Default: frame-parallel segment sum
 * (the whole dphon buffer, B * p_pitch rows, is zeroed by the call; runs of frames are flushed with 16-byte vector atomics,
 * so the order of the partial sums is not fixed).  fs2_lr_tune_bwd(1) selects the reproducible form: one CTA per 8 output
 * rows, every row written exactly once with a plain store (no memset, no atomics, fixed summation order). */
int fs2_lr_bwd(const float* dframes, const float* dframes2, int f_pitch, int f_off, const int* ends,
               const int* mel_lens, int B, int Tp, int Tm, int D, float* dphon, int p_pitch, int p_off, void* stream);
/* measurement hook: rows in flight per warp (2, 4 or 8) in fs2_lr_expand / fs2_lr_bwd (default 4, chosen on B200) */
int fs2_lr_tune(int rows_in_flight);
/* With pe == NULL and out_act == NULL (plain fp32 expansion = speechbrain upsample itself, BASELINE configs[1])
 * fs2_lr_expand runs on the bulk-copy engine: cp.async.bulk global -> shared once per phoneme run, shared -> global
 * once per frame, no register staging.  Measurement hook: rows per CTA of that form (8 ... 128, default 8; 0 = never use it). */
int fs2_lr_bulk_rows(int rows);
/* 1 = reproducible segment-sum form of fs2_lr_bwd, 0 (default) = frame-parallel form (faster on B200: 23.5 vs 31.5 us at
 * BASELINE configs[1]'s size) */
int fs2_lr_tune_bwd(int segment_sum);

/* misc row-space utilities */
/* out = (src + reflect-fold_p(src) + add + add2) * rowmask, zero halo */
int fs2_fold_halo(const float* src, int B, int T, int C, int p, const float* add, const float* add2, const int* lens,
                  float* out_f32, void* out_act, int act_bf16, void* stream);
int fs2_colsum(const void* x, int x_bf16, long long rows, int C, long long ld, float* out /* += */, void* stream);
int fs2_unpad_mask(const float* src, const int* lens, int B, int T, int C, float* out_plain, void* out_act,
                   int act_bf16, int halo, void* stream);
int fs2_pad_rows(const float* src_plain, const float* src2_plain, int B, int T, int C, float scale, float* out_f32,
                 void* out_act, int act_bf16, void* stream);
/* weight packing, one launch for the whole parameter set: dst[co, j, ci] = src[co*src_ld + ci*k + j]
 * (torch Conv1d (Cout,Cin,k) -> tap-major K rows that the implicit GEMM reads) */
typedef struct { long long src_off, dst_off, src_ld; int cout, cin, k, pad_; } Fs2PackItem;
int fs2_pack_weights(const Fs2PackItem* items_dev, int n_items, const float* src_base, void* dst_base, int dst_bf16,
                     void* stream);
int fs2_cast_bf16(const float* src, void* dst, long long n, void* stream);
int fs2_add_(float* dst, const float* src, long long n, void* stream);
int fs2_memset(void* dst, int value, long long nbytes, void* stream);

/* ---------------------------------------------------------------------- losses -- */
/* loss.py:101-160 per-sample sliced MSE x5, mean over B; fused forward + gradient.
 * out[0..4] = mel, postnet, dur, pitch, energy (un-weighted).  Gradients (plain layouts, same shapes as the
 * predictions) are d(sum_i w[i]*loss_i). Pitch/energy slice the PHONEME axis with the mel length (quirk Q5). */
int fs2_mse_losses(const float* mel_out, const float* post_out, const float* mel_tgt, const float* log_dur_pred,
                   const int64_t* dur_tgt, const float* pitch_pred, const float* pitch_tgt,
                   const float* energy_pred, const float* energy_tgt, const int64_t* mel_len, const int64_t* phon_len,
                   int B, int Tp, int Tm, int n_mels, const float* w /*5 host floats*/, float* sums_ws /*5*B floats*/,
                   float* out, float* dmel, float* dpost, float* ddur, float* dpitch, float* denergy, void* stream);
/* speechbrain SSIMLoss (loss.py:155): masked per-sample min-max norm + 11x11 gaussian SSIM (valid conv).
 * out[0] = loss clamped as the reference does (>1 -> 1, <0 -> 0, both with zero gradient);
 * dmel_out += weight * dL/dmel_out when not NULL.  ws: fs2_ssim_ws_floats() floats of scratch. */
long long fs2_ssim_ws_floats(int B, int Tm, int n_mels);
int fs2_ssim_loss(const float* mel_out, const float* mel_tgt, const int64_t* mel_len, int B, int Tm, int n_mels,
                  float weight, float* out, float* dmel_out, float* ws, void* stream);

/* The whole Loss.forward of loss.py:62-186 (+ the gradients wrt the five predictions) in three launches: one pass for the
 * MSE sums / d(postnet) / SSIM min-max statistics, one fused SSIM map + gradient pass that also writes d(mel) (MSE + SSIM
 * parts; the map never leaves shared memory), one finalize (7 values, arg-min/max corrections, SSIMLoss's out-of-range
 * clamp).  w6 (HOST pointer) = weights of mel, postnet, dur, pitch, energy, ssim.  out8 = the six weighted components in
 * that order, total_loss, un-clamped ssim value.  ws = fs2_loss_ws_floats(B) floats that are ZERO on entry (allocate once
 * with zeros); every call leaves them zero again.  dmel..denergy: all NULL (values only) or all non-NULL. */
long long fs2_loss_ws_floats(int B);
int fs2_loss_fused(const float* mel_out, const float* post_out, const float* mel_tgt, const float* log_dur_pred,
                   const int64_t* dur_tgt, const float* pitch_pred, const float* pitch_tgt,
                   const float* energy_pred, const float* energy_tgt, const int64_t* mel_len, const int64_t* phon_len,
                   int B, int Tp, int Tm, int n_mels, const float* w6, float* ws, float* out8, float* dmel, float* dpost,
                   float* ddur, float* dpitch, float* denergy, void* stream);
/* multiplies the five gradients by the device scalar *g_dev unless it is exactly 1 (autograd's upstream gradient) */
int fs2_loss_scale_grads(const float* g_dev, float* dmel, float* dpost, long long n_mel, float* ddur, float* dpitch,
                         float* denergy, long long n_ph, void* stream);

/* train.py:81 AdamW (torch defaults: decoupled weight decay, bias correction) over one flat buffer */
int fs2_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
              float beta2, float eps, float wd, int step, float grad_scale, void* stream);
/* The same update over the slice [base, base+n) of the flat parameter buffer, which in the same pass also refreshes the
 * bf16 operand mirror of that slice (mirror_bf16 = mirror of the WHOLE flat buffer, or NULL) and one gathered operand:
 * columns [0, g_cols) of the g_rows x g_src_ld matrix at flat offset g_src_off, copied densely to mirror[g_dst_off..]
 * (g_rows = 0: none).  guard_tc_error != 0: the update is skipped while the tcgen05 kernels' error word is set
 * (fs2_gemm_tc_error_flag), so a step whose GEMMs timed out never reaches the parameters. */
int fs2_adamw_fused(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, float wd, int step, float grad_scale, void* mirror_bf16, long long base,
                    long long g_src_off, long long g_src_ld, int g_rows, int g_cols, long long g_dst_off,
                    int guard_tc_error, void* stream);

/* Intensity extractor glue (rank_model/model.py:56-109).
 * fs2_frames_to_rows: frame features -> padded row space [B*(T+8), Cpad] in operand storage, columns >= C and the
 *   halo rows zero.  channels_first = 1: x is (B, C, T) (the collate's rank_X, dataset.py:106-110); 0: (B, T, C).
 * fs2_intensity_head: I[b,t,:] = Wc . mask(h[b,t,:] + emb[emotions[b],:]) + bc with mask = (t < lens[b])
 *   (model.py:104-107); h is the fp32 padded-row output of the last FFT block. */
/* x = GELU(x) in place, exact erf form (nn.GELU(), rank_model/model.py:31), n elements of operand storage, n % 4 == 0 */
int fs2_gelu(void* x, long long n, int act_bf16, void* stream);
int fs2_frames_to_rows(const float* x, int channels_first, int B, int C, int T, int Cpad, void* out_act, int act_bf16,
                       void* stream);
int fs2_intensity_head(const float* h, const float* emb, const int64_t* emotions, const int* lens, const float* Wc,
                       const float* bc, int B, int T, int D, int n_out, float* out /*(B,T,n_out)*/, void* stream);

/* Device-side collate (fastspeech2/dataset.py:60-133): the ragged per-utterance arrays arrive in ONE
 * staging buffer (one host->device copy); this call writes every padded batch tensor, zero padding included.
 * Output row i takes utterance i of the (already sorted) descriptor arrays: ph_start/ph_len index phon_cat / dur_cat,
 * fr_start/fr_len index pitch_cat / energy_cat and, times n_mels, mel_cat whose utterance block is (n_mels, fr_len)
 * row-major (the dataset's mel layout).  Outputs: phoneme, duration (B,Tp) i64; mel (B,Tm,n_mels); pitch, energy
 * (B,Tm); rank_X (B,n_mels+2,Tm) = [mel; pitch; energy] channels-first (dataset.py:116-117). */
int fs2_collate(const int64_t* phon_cat, const int64_t* dur_cat, const float* mel_cat, const float* pitch_cat,
                const float* energy_cat, const int* ph_start, const int* ph_len, const int* fr_start, const int* fr_len,
                int B, int Tp, int Tm, int n_mels, int64_t* phoneme, int64_t* duration, float* mel, float* pitch,
                float* energy, float* rank_X, void* stream);

/* Intensity prototype buckets (rank_model/inference.py:88-114): utterances arrive sorted by
 * (group = speaker * n_emo + emotion, relevance score); frame t of utterance i is element frame_off[i] + t of its group's
 * concatenated frame list, which np.array_split cuts into `n_buckets` contiguous runs (the first n % k runs one longer).
 * sums (n_groups, n_buckets, D) must be zero on entry; the call accumulates the per-bucket sums and then divides by the
 * bucket sizes (empty bucket -> NaN, as numpy's mean of an empty slice; groups without frames stay 0 like the reference's
 * zero-initialised array). */
int fs2_prototype_buckets(const float* I /*(N,Tmax,D)*/, const int* lens, const int* group, const long long* frame_off,
                          const long long* group_total /*(n_groups)*/, int N, int Tmax, int D, int n_groups, int n_buckets,
                          float* sums /*(n_groups,n_buckets,D)*/, void* stream);

/* train.py:16-51 duration-segment mean of frame intensities ("next" row f-1) */
int fs2_intensity_segment_mean(const float* I /*(B,Tm,D)*/, const int64_t* dur, const int64_t* phon_len,
                               int B, int Tp, int Tm, int D, float* out /*(B,Tp,D)*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif
