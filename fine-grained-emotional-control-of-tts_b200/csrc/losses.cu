// Fused masked losses of fastspeech2/loss.py: the five per-sample sliced MSE terms (forward + gradient in one
// pass over the mel tensors) and speechbrain's SSIMLoss (masked per-sample min-max normalisation, separable
// 11-tap Gaussian SSIM over the valid (Tm-10) x (n_mels-10) map, analytic backward, device-side clamp).
#include <math.h>
#include "common.cuh"
#include "../../include/fs2_b200.h"

namespace {

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = (lane < (int)((blockDim.x + 31) >> 5)) ? sh[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;   // valid in warp 0
}

// ----------------------------------------------------------------------------- MSE x5 --
struct MseArgs {
  const float *mel_out, *post_out, *mel_tgt, *log_dur, *pitch_pred, *pitch_tgt, *energy_pred, *energy_tgt;
  const int64_t *dur_tgt, *mel_len, *phon_len;
  int B, Tp, Tm, n_mels;
  float w[5];
  float* sums;   // [5][B]
  float *dmel, *dpost, *ddur, *dpitch, *denergy;
};

__global__ void __launch_bounds__(256) mse_kernel(MseArgs a) {
  pdl_wait();
  __shared__ float sh[8];
  const int b = blockIdx.y;
  const int ml = (int)a.mel_len[b];
  const int mlc = min(ml, a.Tm);
  const long long per = (long long)a.Tm * a.n_mels;
  const long long base = (long long)b * per;
  const long long valid = (long long)mlc * a.n_mels;
  const float invB = 1.0f / (float)a.B;
  const float gm = 2.0f * invB / (float)((long long)ml * a.n_mels);   // torch slices [:ml]; ml <= Tm in practice
  float s1 = 0.f, s2 = 0.f;
  for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < per;
       e += (long long)gridDim.x * blockDim.x * 4) {
    float4 g1 = make_float4(0.f, 0.f, 0.f, 0.f), g2 = g1;
    if (e < valid) {   // n_mels % 4 == 0 so a float4 never straddles the valid boundary
      float4 t = ld4(a.mel_tgt + base + e), m = ld4(a.mel_out + base + e), p = ld4(a.post_out + base + e);
      float4 d1 = make_float4(m.x - t.x, m.y - t.y, m.z - t.z, m.w - t.w);
      float4 d2 = make_float4(p.x - t.x, p.y - t.y, p.z - t.z, p.w - t.w);
      s1 += d1.x * d1.x + d1.y * d1.y + d1.z * d1.z + d1.w * d1.w;
      s2 += d2.x * d2.x + d2.y * d2.y + d2.z * d2.z + d2.w * d2.w;
      float k1 = gm * a.w[0], k2 = gm * a.w[1];
      g1 = make_float4(d1.x * k1, d1.y * k1, d1.z * k1, d1.w * k1);
      g2 = make_float4(d2.x * k2, d2.y * k2, d2.z * k2, d2.w * k2);
    }
    if (a.dmel) st4(a.dmel + base + e, g1);
    if (a.dpost) st4(a.dpost + base + e, g2);
  }
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  if (threadIdx.x == 0) {
    atomicAdd(a.sums + 0 * a.B + b, s1);
    atomicAdd(a.sums + 1 * a.B + b, s2);
  }
  if (blockIdx.x == 0) {
    const int pl = min((int)a.phon_len[b], a.Tp);
    const int vl = min(ml, a.Tp);            // loss.py:126-133 slices the phoneme axis with the mel length
    float s3 = 0.f, s4 = 0.f, s5 = 0.f;
    for (int p = threadIdx.x; p < a.Tp; p += blockDim.x) {
      long long o = (long long)b * a.Tp + p;
      float gd = 0.f, gp = 0.f, ge = 0.f;
      if (p < pl) {
        float d = a.log_dur[o] - log1pf((float)a.dur_tgt[o]);
        s3 += d * d;
        gd = a.w[2] * 2.0f * d * invB / (float)pl;
      }
      if (p < vl) {
        float d = a.pitch_pred[o] - a.pitch_tgt[o];
        s4 += d * d;
        gp = a.w[3] * 2.0f * d * invB / (float)vl;
        d = a.energy_pred[o] - a.energy_tgt[o];
        s5 += d * d;
        ge = a.w[4] * 2.0f * d * invB / (float)vl;
      }
      if (a.ddur) a.ddur[o] = gd;
      if (a.dpitch) a.dpitch[o] = gp;
      if (a.denergy) a.denergy[o] = ge;
    }
    s3 = block_sum(s3, sh);
    s4 = block_sum(s4, sh);
    s5 = block_sum(s5, sh);
    if (threadIdx.x == 0) {
      a.sums[2 * a.B + b] = s3;
      a.sums[3 * a.B + b] = s4;
      a.sums[4 * a.B + b] = s5;
    }
  }
}

__global__ void mse_finalize_kernel(const float* sums, const int64_t* mel_len, const int64_t* phon_len, int B, int Tp,
                                    int n_mels, float* out) {
  pdl_wait();
  const int k = threadIdx.x;
  if (k >= 5) return;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    float cnt;
    int ml = (int)mel_len[b];
    if (k < 2) cnt = (float)((long long)ml * n_mels);
    else if (k == 2) cnt = (float)min((int)phon_len[b], Tp);
    else cnt = (float)min(ml, Tp);
    acc += sums[k * B + b] / cnt;        // per-sample mean, as nn.MSELoss on the slice
  }
  out[k] = acc / (float)B;
}

// -------------------------------------------------------------------------------- SSIM --
constexpr int KS = 11;
constexpr int TR = 8;                    // output rows per block
constexpr int MAXW = 80;                 // n_mels upper bound for the static tiles
__constant__ float c_gauss[KS];

struct SsimStats { float mn_p, mx_p, mn_t, mx_t; int argmin_p, argmax_p; float pad0, pad1; };

__device__ __forceinline__ unsigned orderable(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorderable(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void __launch_bounds__(512) ssim_minmax_kernel(const float* pred, const float* tgt, const int64_t* mel_len,
                                                          int Tm, int W, SsimStats* stats) {
  pdl_wait();
  __shared__ unsigned long long s_min[16], s_max[16];
  __shared__ float s_tmn[16], s_tmx[16];
  const int b = blockIdx.x;
  const int len = min((int)mel_len[b], Tm);
  const long long n = (long long)len * W;
  const float* p = pred + (long long)b * Tm * W;
  const float* t = tgt + (long long)b * Tm * W;
  unsigned long long kmin = ~0ull, kmax = 0ull;
  float tmn = INFINITY, tmx = -INFINITY;
  auto take = [&](float v, float tv, unsigned i) {
    const unsigned o = orderable(v);
    const unsigned long long k1 = ((unsigned long long)o << 32) | i;
    const unsigned long long k2 = ((unsigned long long)o << 32) | (unsigned)(~i);
    kmin = k1 < kmin ? k1 : kmin;
    kmax = k2 > kmax ? k2 : kmax;
    tmn = fminf(tmn, tv);
    tmx = fmaxf(tmx, tv);
  };
  long long done = 0;
  if ((((uintptr_t)p | (uintptr_t)t) & 15) == 0) {
    // 16-byte loads, two per operand in flight (one CTA per sample: the scalar loop was bound by load latency);
    // the (value, index) keys make the result independent of the visiting order
    const long long n4 = n >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(p);
    const float4* t4 = reinterpret_cast<const float4*>(t);
    long long i = threadIdx.x;
    for (; i + blockDim.x < n4; i += 2 * blockDim.x) {
      const float4 a = p4[i], b2 = p4[i + blockDim.x], c = t4[i], d = t4[i + blockDim.x];
      const unsigned e = (unsigned)(4 * i), g = (unsigned)(4 * (i + blockDim.x));
      take(a.x, c.x, e); take(a.y, c.y, e + 1); take(a.z, c.z, e + 2); take(a.w, c.w, e + 3);
      take(b2.x, d.x, g); take(b2.y, d.y, g + 1); take(b2.z, d.z, g + 2); take(b2.w, d.w, g + 3);
    }
    if (i < n4) {
      const float4 a = p4[i], c = t4[i];
      const unsigned e = (unsigned)(4 * i);
      take(a.x, c.x, e); take(a.y, c.y, e + 1); take(a.z, c.z, e + 2); take(a.w, c.w, e + 3);
    }
    done = n4 << 2;
  }
  for (long long i = done + threadIdx.x; i < n; i += blockDim.x) take(p[i], t[i], (unsigned)i);
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long a = __shfl_xor_sync(0xffffffffu, kmin, o), c = __shfl_xor_sync(0xffffffffu, kmax, o);
    kmin = a < kmin ? a : kmin;
    kmax = c > kmax ? c : kmax;
  }
  tmn = warp_min(tmn);
  tmx = warp_max(tmx);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_min[warp] = kmin; s_max[warp] = kmax; s_tmn[warp] = tmn; s_tmx[warp] = tmx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      kmin = s_min[w] < kmin ? s_min[w] : kmin;
      kmax = s_max[w] > kmax ? s_max[w] : kmax;
      tmn = fminf(tmn, s_tmn[w]);
      tmx = fmaxf(tmx, s_tmx[w]);
    }
    SsimStats st;
    st.mn_p = unorderable((unsigned)(kmin >> 32));
    st.argmin_p = (int)(unsigned)(kmin & 0xFFFFFFFFu);
    float mxp = unorderable((unsigned)(kmax >> 32));
    int amax = (int)(~(unsigned)(kmax & 0xFFFFFFFFu));
    // amax over x.masked_fill(~mask, 0): padded rows contribute a 0 candidate
    if (len < Tm) {
      if (mxp < 0.f) { mxp = 0.f; amax = -1; }
      if (tmx < 0.f) tmx = 0.f;
    }
    st.mx_p = mxp;
    st.argmax_p = amax;
    st.mn_t = tmn;
    st.mx_t = tmx;
    st.pad0 = st.pad1 = 0.f;
    stats[b] = st;
  }
}

// one block: TR output rows x (W-10) columns of sample b
__global__ void __launch_bounds__(256) ssim_map_kernel(const float* pred, const float* tgt, const int64_t* mel_len,
                                                       int Tm, int W, const SsimStats* stats, float* fa, float* fb,
                                                       float* fc, double* total) {
  pdl_wait();
  __shared__ float q[TR + KS - 1][MAXW], tn[TR + KS - 1][MAXW];
  __shared__ float hz[5][TR + KS - 1][MAXW - KS + 1];
  __shared__ float sh[8];
  const int b = blockIdx.y;
  const int Hm = Tm - (KS - 1), Wm = W - (KS - 1);
  const int r0 = blockIdx.x * TR;
  const int nr = min(TR, Hm - r0);
  const int len = min((int)mel_len[b], Tm);
  const SsimStats st = stats[b];
  const float ip = 1.0f / (st.mx_p - st.mn_p + 1e-8f), it = 1.0f / (st.mx_t - st.mn_t + 1e-8f);
  const int nin = nr + KS - 1;
  for (int i = threadIdx.x; i < nin * W; i += blockDim.x) {
    int r = i / W, c = i - r * W;
    int t = r0 + r;
    float pv = 0.f, tv = 0.f;
    if (t < len) {
      long long o = ((long long)b * Tm + t) * W + c;
      pv = (pred[o] - st.mn_p) * ip;
      tv = (tgt[o] - st.mn_t) * it;
    }
    q[r][c] = pv;
    tn[r][c] = tv;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nin * Wm; i += blockDim.x) {
    int r = i / Wm, c = i - r * Wm;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      float g = c_gauss[k], x = tn[r][c + k], y = q[r][c + k];
      a0 += g * x; a1 += g * y; a2 += g * x * x; a3 += g * y * y; a4 += g * x * y;
    }
    hz[0][r][c] = a0; hz[1][r][c] = a1; hz[2][r][c] = a2; hz[3][r][c] = a3; hz[4][r][c] = a4;
  }
  __syncthreads();
  const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
  float acc = 0.f;
  for (int i = threadIdx.x; i < nr * Wm; i += blockDim.x) {
    int r = i / Wm, c = i - r * Wm;
    float mx = 0.f, my = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      float g = c_gauss[k];
      mx += g * hz[0][r + k][c]; my += g * hz[1][r + k][c];
      exx += g * hz[2][r + k][c]; eyy += g * hz[3][r + k][c]; exy += g * hz[4][r + k][c];
    }
    // x = target, y = prediction (gradient wrt y)
    float sxx = exx - mx * mx, syy = eyy - my * my, sxy = exy - mx * my;
    float denL = mx * mx + my * my + c1, denC = sxx + syy + c2;
    float L = (2.f * mx * my + c1) / denL, CS = (2.f * sxy + c2) / denC;
    acc += L * CS;
    float dL_dmy = (2.f * mx - 2.f * my * L) / denL;
    float dEyy = -L * CS / denC;            // d ss / d E[y^2]
    float dExy = 2.f * L / denC;            // d ss / d E[xy]
    float dmy = CS * dL_dmy + L * ((2.f / denC) * (-mx) + (-CS / denC) * (-2.f * my));
    long long o = ((long long)b * Hm + r0 + r) * Wm + c;
    fa[o] = dmy; fb[o] = dEyy; fc[o] = dExy;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) atomicAdd(total, (double)acc);
}

// gradient wrt the normalised prediction, gathered from the three map fields (transposed Gaussian), then through
// the mask and the affine part of the min-max normalisation; S1/S2 feed the argmin/argmax corrections.
__global__ void __launch_bounds__(256) ssim_grad_kernel(const float* pred, const float* tgt, const int64_t* mel_len,
                                                        int B, int Tm, int W, const SsimStats* stats, const float* fa,
                                                        const float* fb, const float* fc, const double* total,
                                                        float weight, float* S, float* dmel) {
  pdl_wait();
  __shared__ float f[3][TR + KS - 1][MAXW - KS + 1];
  __shared__ float vz[3][TR][MAXW - KS + 1];
  __shared__ float sh[8];
  const int b = blockIdx.y;
  const int Hm = Tm - (KS - 1), Wm = W - (KS - 1);
  const int t0 = blockIdx.x * TR;
  const int nr = min(TR, Tm - t0);
  const int len = min((int)mel_len[b], Tm);
  const double N = (double)B * Hm * Wm;
  const float loss = (float)(1.0 - *total / N);
  const float gs = (loss > 1.0f || loss < 0.0f) ? 0.f : (float)(-(double)weight / N);
  if (t0 >= len) return;                    // masked rows get no gradient (uniform per block)
  const SsimStats st = stats[b];
  const float ip = 1.0f / (st.mx_p - st.mn_p + 1e-8f), it = 1.0f / (st.mx_t - st.mn_t + 1e-8f);
  // map rows t0-10 .. t0+nr-1
  for (int i = threadIdx.x; i < (nr + KS - 1) * Wm; i += blockDim.x) {
    int r = i / Wm, c = i - r * Wm;
    int mr = t0 - (KS - 1) + r;
    float a = 0.f, bb = 0.f, cc = 0.f;
    if (mr >= 0 && mr < Hm) {
      long long o = ((long long)b * Hm + mr) * Wm + c;
      a = fa[o]; bb = fb[o]; cc = fc[o];
    }
    f[0][r][c] = a; f[1][r][c] = bb; f[2][r][c] = cc;
  }
  __syncthreads();
  // vertical: input row t0+r gathers map rows (t0+r-k), k=0..10  -> tile rows r+10-k
  for (int i = threadIdx.x; i < nr * Wm; i += blockDim.x) {
    int r = i / Wm, c = i - r * Wm;
    float a = 0.f, bb = 0.f, cc = 0.f;
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      float g = c_gauss[k];
      a += g * f[0][r + KS - 1 - k][c]; bb += g * f[1][r + KS - 1 - k][c]; cc += g * f[2][r + KS - 1 - k][c];
    }
    vz[0][r][c] = a; vz[1][r][c] = bb; vz[2][r][c] = cc;
  }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < nr * W; i += blockDim.x) {
    int r = i / W, m = i - r * W;
    int t = t0 + r;
    if (t >= len) continue;
    float a = 0.f, bb = 0.f, cc = 0.f;
#pragma unroll
    for (int l = 0; l < KS; ++l) {
      int c = m - l;
      if (c >= 0 && c < Wm) { float g = c_gauss[l]; a += g * vz[0][r][c]; bb += g * vz[1][r][c]; cc += g * vz[2][r][c]; }
    }
    long long o = ((long long)b * Tm + t) * W + m;
    float pv = pred[o];
    float qv = (pv - st.mn_p) * ip, tv = (tgt[o] - st.mn_t) * it;
    float dq = gs * (a + 2.f * qv * bb + tv * cc);
    s1 += dq;
    s2 += dq * (pv - st.mn_p);
    dmel[o] += dq * ip;
  }
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  if (threadIdx.x == 0) { atomicAdd(S + 2 * b, s1); atomicAdd(S + 2 * b + 1, s2); }
}

__global__ void ssim_finalize_kernel(const double* total, const SsimStats* stats, const float* S, int B, int Tm, int W,
                                     float* out, float* dmel) {
  pdl_wait();
  const int Hm = Tm - (KS - 1), Wm = W - (KS - 1);
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) {
    float loss = (float)(1.0 - *total / ((double)B * Hm * Wm));
    if (loss > 1.0f) loss = 1.0f;       // loss.py:155 -> SSIMLoss: out-of-range values become constants
    if (loss < 0.0f) loss = 0.0f;
    out[0] = loss;
  }
  if (b < B && dmel) {
    const SsimStats st = stats[b];
    const float r = st.mx_p - st.mn_p + 1e-8f;
    const float s1 = S[2 * b], s2 = S[2 * b + 1];
    const long long base = (long long)b * Tm * W;
    dmel[base + st.argmin_p] += -s1 / r + s2 / (r * r);
    if (st.argmax_p >= 0) dmel[base + st.argmax_p] += -s2 / (r * r);
  }
}


// =====================================================================================================================
// Round 2: the whole Loss.forward (+ gradients) in THREE launches, no memsets, no torch glue
//   loss_pass1    one pass over mel_out / postnet_out / target: both mel MSE sums, d(postnet), per-sample min / max
//                 (+ arg) of prediction and target for SSIM's min-max normalisation (64-bit atomicMax on (value, index)
//                 keys), the three phoneme-level MSE terms and their gradients;
//   ssim_fused    per tile of TRI mel rows: normalise -> separable 11-tap Gaussian moments -> SSIM map -> analytic
//                 gradient gathered back through the transposed Gaussian -> d(mel) = MSE part + SSIM part.  The map and
//                 the three gradient fields never leave shared memory (the tile recomputes a 10-row halo of the map
//                 instead of exchanging it through HBM); every thread owns a short run of outputs of each 1-D pass, so
//                 one shared-memory load feeds ~10 FMAs;
//   loss_finalize the 7 loss values (weights applied, total), argmin / argmax corrections of the normalisation, the
//                 out-of-range clamp of speechbrain's SSIMLoss (value replaced by a constant => the SSIM gradient is
//                 removed again: d(mel) is rewritten MSE-only), and the workspace is handed back zeroed (last CTA).
// The workspace must be zero on entry (allocate it with zeros once); every call leaves it zero.
struct LossArgs {
  const float *mel_out, *post_out, *mel_tgt, *log_dur, *pitch_pred, *pitch_tgt, *energy_pred, *energy_tgt;
  const int64_t *dur_tgt, *mel_len, *phon_len;
  int B, Tp, Tm, W;
  float w[6];                  // mel, postnet, dur, pitch, energy, ssim
  float* ws;
  float* out;                  // [8]: six weighted components (order of w), total, un-clamped ssim value
  float *dmel, *dpost, *ddur, *dpitch, *denergy;       // dmel == nullptr: values only
};

__host__ __device__ inline int lw_sums() { return 4; }
__host__ __device__ inline int lw_S(int B) { return 4 + 5 * B; }
__host__ __device__ inline int lw_k64(int B) { return (4 + 7 * B + 1) & ~1; }
__host__ __device__ inline int lw_t32(int B) { return lw_k64(B) + 4 * B; }
__host__ __device__ inline int lw_words(int B) { return lw_t32(B) + 2 * B; }

struct LossNorm { float mn_p, ip, r, mn_t, it; int argmin, argmax; };

__device__ __forceinline__ LossNorm loss_norm(const float* ws, int B, int b, int len, int Tm) {
  const unsigned long long* k64 = reinterpret_cast<const unsigned long long*>(ws + lw_k64(B));
  const unsigned* t32 = reinterpret_cast<const unsigned*>(ws + lw_t32(B));
  const unsigned long long kmin = ~k64[b], kmax = k64[B + b];
  LossNorm n;
  n.mn_p = unorderable((unsigned)(kmin >> 32));
  n.argmin = (int)(unsigned)(kmin & 0xFFFFFFFFu);
  float mxp = unorderable((unsigned)(kmax >> 32));
  n.argmax = (int)(~(unsigned)(kmax & 0xFFFFFFFFu));
  n.mn_t = unorderable(~t32[b]);
  float tmx = unorderable(t32[B + b]);
  if (len < Tm) {                     // amax over x.masked_fill(~mask, 0): padded rows contribute a 0 candidate
    if (mxp < 0.f) { mxp = 0.f; n.argmax = -1; }
    if (tmx < 0.f) tmx = 0.f;
  }
  n.r = mxp - n.mn_p + 1e-8f;
  n.ip = 1.0f / n.r;
  n.it = 1.0f / (tmx - n.mn_t + 1e-8f);
  return n;
}

__global__ void __launch_bounds__(256) loss_pass1_kernel(LossArgs a) {
  pdl_wait();
  __shared__ float sh[8];
  __shared__ unsigned long long s_min[8], s_max[8];
  __shared__ float s_tmn[8], s_tmx[8];
  const int b = blockIdx.y;
  const int ml = (int)a.mel_len[b];
  const int mlc = min(ml, a.Tm);
  const long long per = (long long)a.Tm * a.W;
  const long long base = (long long)b * per;
  const long long valid = (long long)mlc * a.W;
  const float invB = 1.0f / (float)a.B;
  const float k2 = a.w[1] * 2.0f * invB / (float)((long long)ml * a.W);
  float s1 = 0.f, s2 = 0.f;
  unsigned long long kmin = ~0ull, kmax = 0ull;
  float tmn = INFINITY, tmx = -INFINITY;
  auto take = [&](float v, float tv, unsigned i) {
    const unsigned o = orderable(v);
    const unsigned long long k1 = ((unsigned long long)o << 32) | i;
    const unsigned long long k2_ = ((unsigned long long)o << 32) | (unsigned)(~i);
    kmin = k1 < kmin ? k1 : kmin;
    kmax = k2_ > kmax ? k2_ : kmax;
    tmn = fminf(tmn, tv);
    tmx = fmaxf(tmx, tv);
  };
  for (long long e = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; e < per;
       e += (long long)gridDim.x * blockDim.x * 4) {
    float4 g2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < valid) {   // W % 4 == 0 so a float4 never straddles the valid boundary
      const float4 t = ld4(a.mel_tgt + base + e), m = ld4(a.mel_out + base + e), p = ld4(a.post_out + base + e);
      const float4 d1 = make_float4(m.x - t.x, m.y - t.y, m.z - t.z, m.w - t.w);
      const float4 d2 = make_float4(p.x - t.x, p.y - t.y, p.z - t.z, p.w - t.w);
      s1 += d1.x * d1.x + d1.y * d1.y + d1.z * d1.z + d1.w * d1.w;
      s2 += d2.x * d2.x + d2.y * d2.y + d2.z * d2.z + d2.w * d2.w;
      g2 = make_float4(d2.x * k2, d2.y * k2, d2.z * k2, d2.w * k2);
      const unsigned i = (unsigned)e;
      take(m.x, t.x, i); take(m.y, t.y, i + 1); take(m.z, t.z, i + 2); take(m.w, t.w, i + 3);
    }
    if (a.dpost) st4(a.dpost + base + e, g2);
  }
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long x = __shfl_xor_sync(0xffffffffu, kmin, o), y = __shfl_xor_sync(0xffffffffu, kmax, o);
    kmin = x < kmin ? x : kmin;
    kmax = y > kmax ? y : kmax;
  }
  tmn = warp_min(tmn);
  tmx = warp_max(tmx);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_min[warp] = kmin; s_max[warp] = kmax; s_tmn[warp] = tmn; s_tmx[warp] = tmx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      kmin = s_min[w] < kmin ? s_min[w] : kmin;
      kmax = s_max[w] > kmax ? s_max[w] : kmax;
      tmn = fminf(tmn, s_tmn[w]);
      tmx = fmaxf(tmx, s_tmx[w]);
    }
    atomicAdd(a.ws + lw_sums() + 0 * a.B + b, s1);
    atomicAdd(a.ws + lw_sums() + 1 * a.B + b, s2);
    if (kmax != 0ull) {                          // this CTA saw at least one valid element
      unsigned long long* k64 = reinterpret_cast<unsigned long long*>(a.ws + lw_k64(a.B));
      unsigned* t32 = reinterpret_cast<unsigned*>(a.ws + lw_t32(a.B));
      atomicMax(k64 + b, ~kmin);                 // min kept as the max of the complement: zero is the identity of both
      atomicMax(k64 + a.B + b, kmax);
      atomicMax(t32 + b, ~orderable(tmn));
      atomicMax(t32 + a.B + b, orderable(tmx));
    }
  }
  if (blockIdx.x == 0) {
    const int pl = min((int)a.phon_len[b], a.Tp);
    const int vl = min(ml, a.Tp);            // loss.py:126-133 slices the phoneme axis with the mel length
    float s3 = 0.f, s4 = 0.f, s5 = 0.f;
    for (int p = threadIdx.x; p < a.Tp; p += blockDim.x) {
      const long long o = (long long)b * a.Tp + p;
      float gd = 0.f, gp = 0.f, ge = 0.f;
      if (p < pl) {
        const float d = a.log_dur[o] - log1pf((float)a.dur_tgt[o]);
        s3 += d * d;
        gd = a.w[2] * 2.0f * d * invB / (float)pl;
      }
      if (p < vl) {
        float d = a.pitch_pred[o] - a.pitch_tgt[o];
        s4 += d * d;
        gp = a.w[3] * 2.0f * d * invB / (float)vl;
        d = a.energy_pred[o] - a.energy_tgt[o];
        s5 += d * d;
        ge = a.w[4] * 2.0f * d * invB / (float)vl;
      }
      if (a.ddur) a.ddur[o] = gd;
      if (a.dpitch) a.dpitch[o] = gp;
      if (a.denergy) a.denergy[o] = ge;
    }
    s3 = block_sum(s3, sh);
    s4 = block_sum(s4, sh);
    s5 = block_sum(s5, sh);
    if (threadIdx.x == 0) {
      a.ws[lw_sums() + 2 * a.B + b] = s3;
      a.ws[lw_sums() + 3 * a.B + b] = s4;
      a.ws[lw_sums() + 4 * a.B + b] = s5;
    }
  }
}

constexpr int LT = 32;                    // mel rows whose gradient one CTA produces
constexpr int LTHREADS = 512;
constexpr int LNR = LT + 2 * (KS - 1);    // input rows staged (10-row halo either side)
constexpr int LNF = LT + (KS - 1);        // map rows computed (10 recomputed + LT owned)
constexpr int PW = MAXW;                  // row pitch of input tiles
constexpr int PM = MAXW - (KS - 1);       // row pitch of map tiles
constexpr int HB = 10;                    // outputs per thread along a horizontal pass
constexpr int VB = 7;                     // map rows per thread in the vertical pass
constexpr int GB = 8;                     // rows per thread in the transposed vertical pass
constexpr int SSIM_SMEM_FLOATS = 2 * LNR * PW + 5 * LNR * PM + 3 * LNF * PM + 16;
static_assert(3 * LT * PM + LT * PW <= 5 * LNR * PM, "vz + dq alias the moment tiles");

__global__ void __launch_bounds__(LTHREADS, 1) ssim_fused_kernel(LossArgs a) {
  pdl_wait();
  extern __shared__ float sm[];
  float* q = sm;                          // [LNR][PW] normalised prediction
  float* tn = q + LNR * PW;               // [LNR][PW] normalised target
  float* hz = tn + LNR * PW;              // [5][LNR][PM] horizontal moments
  float* f = hz + 5 * LNR * PM;           // [3][LNF][PM] gradient fields of the map
  float* sh = f + 3 * LNF * PM;           // [16]
  float* vz = hz;                         // [3][LT][PM]   (after the vertical pass the moments are dead)
  float* dqs = hz + 3 * LT * PM;          // [LT][PW]
  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int W = a.W, Wm = W - (KS - 1), Tm = a.Tm, Hm = Tm - (KS - 1);
  const int t0 = blockIdx.x * LT;
  const int nout = min(LT, Tm - t0);
  const int len = min((int)a.mel_len[b], Tm);
  const int n_owned = max(0, min(t0 + LT, Hm) - t0);
  double* total = reinterpret_cast<double*>(a.ws);
  const long long sbase = (long long)b * Tm * W;
  if (t0 - (KS - 1) >= len) {
    // the tile sees only padding: both images are 0 there, every map point is exactly 1, no gradient
    if (tid == 0 && n_owned > 0) atomicAdd(total, (double)n_owned * (double)Wm);
    if (a.dmel)
      for (int i = tid * 4; i < nout * W; i += blockDim.x * 4)
        st4(a.dmel + sbase + (long long)t0 * W + i, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  const LossNorm nm = loss_norm(a.ws, a.B, b, len, Tm);
  // ---- stage the normalised rows t0-10 .. t0+LT+9 (zero outside [0, len))
  for (int i = tid * 4; i < LNR * W; i += blockDim.x * 4) {
    const int r = i / W, c = i - r * W;
    const int t = t0 - (KS - 1) + r;
    float4 pv = make_float4(0.f, 0.f, 0.f, 0.f), tv = pv;
    if (t >= 0 && t < len) {
      const long long o = sbase + (long long)t * W + c;
      const float4 p4 = ld4(a.mel_out + o), t4 = ld4(a.mel_tgt + o);
      pv = make_float4((p4.x - nm.mn_p) * nm.ip, (p4.y - nm.mn_p) * nm.ip, (p4.z - nm.mn_p) * nm.ip, (p4.w - nm.mn_p) * nm.ip);
      tv = make_float4((t4.x - nm.mn_t) * nm.it, (t4.y - nm.mn_t) * nm.it, (t4.z - nm.mn_t) * nm.it, (t4.w - nm.mn_t) * nm.it);
    }
    st4(q + r * PW + c, pv);
    st4(tn + r * PW + c, tv);
  }
  __syncthreads();
  float g[KS];
#pragma unroll
  for (int k = 0; k < KS; ++k) g[k] = c_gauss[k];
  // ---- horizontal moments: a thread owns HB consecutive columns of one row
  {
    const int ncb = (Wm + HB - 1) / HB;
    for (int task = tid; task < LNR * ncb; task += blockDim.x) {
      const int r = task / ncb, c0 = (task - r * ncb) * HB;
      float x[HB + KS - 1], y[HB + KS - 1], xx[HB + KS - 1], yy[HB + KS - 1], xy[HB + KS - 1];
#pragma unroll
      for (int i = 0; i < HB + KS - 1; ++i) {
        const bool in = c0 + i < W;
        x[i] = in ? tn[r * PW + c0 + i] : 0.f;
        y[i] = in ? q[r * PW + c0 + i] : 0.f;
        xx[i] = x[i] * x[i]; yy[i] = y[i] * y[i]; xy[i] = x[i] * y[i];
      }
#pragma unroll
      for (int o = 0; o < HB; ++o) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k < KS; ++k) {
          a0 = fmaf(g[k], x[o + k], a0); a1 = fmaf(g[k], y[o + k], a1); a2 = fmaf(g[k], xx[o + k], a2);
          a3 = fmaf(g[k], yy[o + k], a3); a4 = fmaf(g[k], xy[o + k], a4);
        }
        if (c0 + o < Wm) {
          float* h = hz + r * PM + c0 + o;
          h[0] = a0; h[LNR * PM] = a1; h[2 * LNR * PM] = a2; h[3 * LNR * PM] = a3; h[4 * LNR * PM] = a4;
        }
      }
    }
  }
  __syncthreads();
  // ---- vertical moments + SSIM point + gradient fields: a thread owns VB consecutive map rows of one column
  const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
  float acc_total = 0.f;
  {
    const int nrb = (LNF + VB - 1) / VB;
    for (int task = tid; task < nrb * Wm; task += blockDim.x) {
      const int rb = task / Wm, c = task - rb * Wm;
      const int r0 = rb * VB;
      float m[VB][5];
#pragma unroll
      for (int o = 0; o < VB; ++o)
#pragma unroll
        for (int j = 0; j < 5; ++j) m[o][j] = 0.f;
#pragma unroll
      for (int ir = 0; ir < VB + KS - 1; ++ir) {
        const int row = r0 + ir;
        float h[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) h[j] = row < LNR ? hz[j * LNR * PM + row * PM + c] : 0.f;
#pragma unroll
        for (int o = 0; o < VB; ++o) {
          const int k = ir - o;
          if (k >= 0 && k < KS) {
#pragma unroll
            for (int j = 0; j < 5; ++j) m[o][j] += g[k] * h[j];
          }
        }
      }
#pragma unroll
      for (int o = 0; o < VB; ++o) {
        const int mlr = r0 + o;
        if (mlr >= LNF) continue;
        const int mr = t0 - (KS - 1) + mlr;
        float fa = 0.f, fb = 0.f, fc = 0.f;
        if (mr >= 0 && mr < Hm) {
          // x = target, y = prediction (gradient wrt y)
          const float mx = m[o][0], my = m[o][1];
          const float sxx = m[o][2] - mx * mx, syy = m[o][3] - my * my, sxy = m[o][4] - mx * my;
          // two reciprocals instead of seven IEEE divisions (rcp.approx: <= 1 ulp, far inside the 2e-5 loss tolerance)
          const float rL = __fdividef(1.0f, mx * mx + my * my + c1), rC = __fdividef(1.0f, sxx + syy + c2);
          const float Lm = (2.f * mx * my + c1) * rL, CS = (2.f * sxy + c2) * rC;
          if (mlr >= KS - 1) acc_total += Lm * CS;                    // owned rows: mr >= t0
          const float dL_dmy = (2.f * mx - 2.f * my * Lm) * rL;
          fb = -Lm * CS * rC;                                         // d ss / d E[y^2]
          fc = 2.f * Lm * rC;                                         // d ss / d E[xy]
          fa = CS * dL_dmy + 2.f * Lm * rC * (CS * my - mx);
        }
        f[mlr * PM + c] = fa; f[LNF * PM + mlr * PM + c] = fb; f[2 * LNF * PM + mlr * PM + c] = fc;
      }
    }
  }
  acc_total = block_sum(acc_total, sh);       // (its barriers also order the moment reads before vz overwrites them)
  if (tid == 0) atomicAdd(total, (double)acc_total);
  if (a.dmel == nullptr) return;
  __syncthreads();
  // ---- transposed vertical pass: input row t0+tl gathers map rows t0+tl-k  (field row tl + 10 - k)
  {
    const int nrb = (LT + GB - 1) / GB;
    for (int task = tid; task < nrb * Wm; task += blockDim.x) {
      const int rb = task / Wm, c = task - rb * Wm;
      const int r0 = rb * GB;
      float v[GB][3];
#pragma unroll
      for (int o = 0; o < GB; ++o) { v[o][0] = 0.f; v[o][1] = 0.f; v[o][2] = 0.f; }
#pragma unroll
      for (int ir = 0; ir < GB + KS - 1; ++ir) {
        const int row = r0 + ir;
        float h[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) h[j] = row < LNF ? f[j * LNF * PM + row * PM + c] : 0.f;
#pragma unroll
        for (int o = 0; o < GB; ++o) {
          const int k = o + (KS - 1) - ir;
          if (k >= 0 && k < KS) { v[o][0] += g[k] * h[0]; v[o][1] += g[k] * h[1]; v[o][2] += g[k] * h[2]; }
        }
      }
#pragma unroll
      for (int o = 0; o < GB; ++o) {
        const int tl = r0 + o;
        if (tl < LT) { vz[tl * PM + c] = v[o][0]; vz[LT * PM + tl * PM + c] = v[o][1]; vz[2 * LT * PM + tl * PM + c] = v[o][2]; }
      }
    }
  }
  __syncthreads();
  // ---- transposed horizontal pass + chain through the products: dq = A + 2 q B + t C  (not yet scaled)
  {
    const int nmb = (W + HB - 1) / HB;
    for (int task = tid; task < LT * nmb; task += blockDim.x) {
      const int tl = task / nmb, m0 = (task - tl * nmb) * HB;
      float wa[HB + KS - 1], wb[HB + KS - 1], wc[HB + KS - 1];
#pragma unroll
      for (int j = 0; j < HB + KS - 1; ++j) {
        const int c = m0 - (KS - 1) + j;
        const bool in = c >= 0 && c < Wm;
        wa[j] = in ? vz[tl * PM + c] : 0.f;
        wb[j] = in ? vz[LT * PM + tl * PM + c] : 0.f;
        wc[j] = in ? vz[2 * LT * PM + tl * PM + c] : 0.f;
      }
#pragma unroll
      for (int o = 0; o < HB; ++o) {
        if (m0 + o >= W) continue;
        float A = 0.f, Bv = 0.f, Cv = 0.f;
#pragma unroll
        for (int l = 0; l < KS; ++l) {
          A += g[l] * wa[o + (KS - 1) - l]; Bv += g[l] * wb[o + (KS - 1) - l]; Cv += g[l] * wc[o + (KS - 1) - l];
        }
        const float qv = q[(tl + KS - 1) * PW + m0 + o], tv = tn[(tl + KS - 1) * PW + m0 + o];
        dqs[tl * PW + m0 + o] = A + 2.f * qv * Bv + tv * Cv;
      }
    }
  }
  __syncthreads();
  // ---- d(mel_out) = MSE part + SSIM part (through the mask and the affine part of the normalisation), coalesced
  const double N = (double)a.B * Hm * Wm;
  const float gs = (float)(-(double)a.w[5] / N);
  const int ml = (int)a.mel_len[b];
  const float k1 = a.w[0] * 2.0f / ((float)a.B * (float)((long long)ml * W));
  float s1 = 0.f, s2 = 0.f;
  for (int i = tid * 4; i < nout * W; i += blockDim.x * 4) {
    const int tl = i / W, m = i - tl * W;
    const int t = t0 + tl;
    const long long o = sbase + (long long)t * W + m;
    float4 dm = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < len) {
      const float4 pv = ld4(a.mel_out + o), tv = ld4(a.mel_tgt + o), dq4 = ld4(dqs + tl * PW + m);
      const float dq[4] = {gs * dq4.x, gs * dq4.y, gs * dq4.z, gs * dq4.w};
      const float pp[4] = {pv.x, pv.y, pv.z, pv.w}, tt[4] = {tv.x, tv.y, tv.z, tv.w};
      float r4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s1 += dq[j];
        s2 += dq[j] * (pp[j] - nm.mn_p);
        r4[j] = k1 * (pp[j] - tt[j]) + dq[j] * nm.ip;
      }
      dm = make_float4(r4[0], r4[1], r4[2], r4[3]);
    }
    st4(a.dmel + o, dm);
  }
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  if (tid == 0) { atomicAdd(a.ws + lw_S(a.B) + 2 * b, s1); atomicAdd(a.ws + lw_S(a.B) + 2 * b + 1, s2); }
}

constexpr int FIN_X = 8;
__global__ void __launch_bounds__(256) loss_finalize_kernel(LossArgs a) {
  pdl_wait();
  __shared__ int s_last;
  const int b = blockIdx.y, tid = threadIdx.x;
  const int W = a.W, Tm = a.Tm, Hm = Tm - (KS - 1), Wm = W - (KS - 1);
  const double N = (double)a.B * Hm * Wm;
  const double total = *reinterpret_cast<const double*>(a.ws);
  const float raw = (float)(1.0 - total / N);
  const bool clamped = raw > 1.0f || raw < 0.0f;   // SSIMLoss returns a constant then: no SSIM gradient (loss.py:155)
  const int ml = (int)a.mel_len[b];
  const int len = min(ml, Tm);
  if (a.dmel) {
    const long long sbase = (long long)b * Tm * W;
    if (clamped) {
      const float k1 = a.w[0] * 2.0f / ((float)a.B * (float)((long long)ml * W));
      const long long per = (long long)Tm * W, valid = (long long)len * W;
      for (long long e = ((long long)blockIdx.x * blockDim.x + tid) * 4; e < per; e += (long long)gridDim.x * blockDim.x * 4) {
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < valid) {
          const float4 m = ld4(a.mel_out + sbase + e), t = ld4(a.mel_tgt + sbase + e);
          d = make_float4(k1 * (m.x - t.x), k1 * (m.y - t.y), k1 * (m.z - t.z), k1 * (m.w - t.w));
        }
        st4(a.dmel + sbase + e, d);
      }
    } else if (blockIdx.x == 0 && tid == 0) {
      // the normalisation's min / max depend on one element each: their share of the gradient
      const LossNorm nm = loss_norm(a.ws, a.B, b, len, Tm);
      const float s1 = a.ws[lw_S(a.B) + 2 * b], s2 = a.ws[lw_S(a.B) + 2 * b + 1];
      a.dmel[sbase + nm.argmin] += -s1 / nm.r + s2 / (nm.r * nm.r);
      if (nm.argmax >= 0) a.dmel[sbase + nm.argmax] += -s2 / (nm.r * nm.r);
    }
  }
  if (blockIdx.x == 0 && b == 0 && tid < 32) {
    float comp[5];
    for (int k = 0; k < 5; ++k) {
      float acc = 0.f;
      for (int i = tid; i < a.B; i += 32) {
        const int mli = (int)a.mel_len[i];
        float cnt;
        if (k < 2) cnt = (float)((long long)mli * W);
        else if (k == 2) cnt = (float)min((int)a.phon_len[i], a.Tp);
        else cnt = (float)min(mli, a.Tp);
        acc += a.ws[lw_sums() + k * a.B + i] / cnt;          // per-sample mean, as nn.MSELoss on the slice
      }
      comp[k] = a.w[k] * warp_sum(acc) / (float)a.B;
    }
    if (tid == 0) {
      const float ssim = raw > 1.0f ? 1.0f : (raw < 0.0f ? 0.0f : raw);
      const float cs = a.w[5] * ssim;
      a.out[0] = comp[0]; a.out[1] = comp[1]; a.out[2] = comp[2]; a.out[3] = comp[3]; a.out[4] = comp[4];
      a.out[5] = cs;
      a.out[6] = cs + comp[0] + comp[1] + comp[2] + comp[3] + comp[4];     // loss.py:161-168 order
      a.out[7] = raw;
    }
  }
  // last CTA to get here hands the workspace back zeroed
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned done = atomicAdd(reinterpret_cast<unsigned*>(a.ws) + 2, 1u);
    s_last = (done == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int i = tid; i < lw_words(a.B); i += blockDim.x) a.ws[i] = 0.f;
  }
}

// autograd hands the loss its upstream gradient as a device scalar; it is 1 for `total_loss.backward()`: nothing to do
__global__ void __launch_bounds__(256) loss_scale_kernel(const float* gptr, float* d0, float* d1, long long n_mel, float* d2,
                                                         float* d3, float* d4, long long n_ph) {
  pdl_wait();
  const float gv = *gptr;
  if (gv == 1.0f) return;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_mel; i += stride) { d0[i] *= gv; d1[i] *= gv; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_ph; i += stride) { d2[i] *= gv; d3[i] *= gv; d4[i] *= gv; }
}

bool g_gauss_ready = false;
int upload_gauss() {
  if (g_gauss_ready) return FS2_OK;
  float g[KS];
  double sum = 0.0;
  for (int i = 0; i < KS; ++i) {
    float c = (float)i - (KS - 1) / 2.0f;
    g[i] = expf(-(c * c) / (2.f * 1.5f * 1.5f));
    sum += g[i];
  }
  for (int i = 0; i < KS; ++i) g[i] = (float)(g[i] / sum);
  CUDA_CHECK_RET(cudaMemcpyToSymbol(c_gauss, g, sizeof g));
  g_gauss_ready = true;
  return FS2_OK;
}

}  // namespace

#define ST ((cudaStream_t)stream)
#define REQUIRE(cond, msg) do { if (!(cond)) { fs2_set_error(msg); return FS2_ERR_ARG; } } while (0)

extern "C" int fs2_mse_losses(const float* mel_out, const float* post_out, const float* mel_tgt,
                              const float* log_dur_pred, const int64_t* dur_tgt, const float* pitch_pred,
                              const float* pitch_tgt, const float* energy_pred, const float* energy_tgt,
                              const int64_t* mel_len, const int64_t* phon_len, int B, int Tp, int Tm, int n_mels,
                              const float* w, float* sums_ws, float* out, float* dmel, float* dpost, float* ddur,
                              float* dpitch, float* denergy, void* stream) {
  REQUIRE(mel_out && post_out && mel_tgt && log_dur_pred && dur_tgt && pitch_pred && pitch_tgt && energy_pred &&
              energy_tgt && mel_len && phon_len && w && sums_ws && out,
          "fs2_mse_losses: null pointer");
  REQUIRE(n_mels % 4 == 0, "fs2_mse_losses: n_mels must be a multiple of 4");
  CUDA_CHECK_RET(cudaMemsetAsync(sums_ws, 0, sizeof(float) * 5 * B, ST));
  MseArgs a;
  a.mel_out = mel_out; a.post_out = post_out; a.mel_tgt = mel_tgt; a.log_dur = log_dur_pred;
  a.pitch_pred = pitch_pred; a.pitch_tgt = pitch_tgt; a.energy_pred = energy_pred; a.energy_tgt = energy_tgt;
  a.dur_tgt = dur_tgt; a.mel_len = mel_len; a.phon_len = phon_len;
  a.B = B; a.Tp = Tp; a.Tm = Tm; a.n_mels = n_mels;
  for (int i = 0; i < 5; ++i) a.w[i] = w[i];
  a.sums = sums_ws;
  a.dmel = dmel; a.dpost = dpost; a.ddur = ddur; a.dpitch = dpitch; a.denergy = denergy;
  long long per = (long long)Tm * n_mels;
  int gx = (int)((per / 4 + 255) / 256);
  if (gx > 64) gx = 64;
  if (gx < 1) gx = 1;
  FS2_LAUNCH((mse_kernel), dim3(gx, B), 256, 0, ST, a);
  int rc = fs2_check_launch();
  if (rc) return rc;
  FS2_LAUNCH((mse_finalize_kernel), 1, 32, 0, ST, sums_ws, mel_len, phon_len, B, Tp, n_mels, out);
  return fs2_check_launch();
}

extern "C" long long fs2_ssim_ws_floats(int B, int Tm, int n_mels) {
  long long Hm = Tm - (KS - 1), Wm = n_mels - (KS - 1);
  if (Hm < 1 || Wm < 1) return 0;
  return 16 + 2LL * B + 8LL * B + 3LL * B * Hm * Wm;
}

extern "C" int fs2_ssim_loss(const float* mel_out, const float* mel_tgt, const int64_t* mel_len, int B, int Tm,
                             int n_mels, float weight, float* out, float* dmel_out, float* ws, void* stream) {
  REQUIRE(mel_out && mel_tgt && mel_len && out && ws, "fs2_ssim_loss: null pointer");
  REQUIRE(Tm >= KS && n_mels >= KS && n_mels <= MAXW, "fs2_ssim_loss: needs Tm >= 11 and 11 <= n_mels <= 80");
  int rc = upload_gauss();
  if (rc) return rc;
  const long long Hm = Tm - (KS - 1), Wm = n_mels - (KS - 1);
  double* total = (double*)ws;                       // ws[0..1]
  float* S = ws + 16;                                // 2B
  SsimStats* stats = (SsimStats*)(ws + 16 + 2LL * B);   // 8B floats
  float* fa = ws + 16 + 10LL * B;
  float* fb = fa + (long long)B * Hm * Wm;
  float* fc = fb + (long long)B * Hm * Wm;
  CUDA_CHECK_RET(cudaMemsetAsync(ws, 0, sizeof(float) * (16 + 2LL * B), ST));
  FS2_LAUNCH((ssim_minmax_kernel), B, 512, 0, ST, mel_out, mel_tgt, mel_len, Tm, n_mels, stats);
  if ((rc = fs2_check_launch())) return rc;
  FS2_LAUNCH((ssim_map_kernel), dim3((unsigned)((Hm + TR - 1) / TR), B), 256, 0, ST, mel_out, mel_tgt, mel_len, Tm, n_mels, stats,
                                                                           fa, fb, fc, total);
  if ((rc = fs2_check_launch())) return rc;
  if (dmel_out) {
    FS2_LAUNCH((ssim_grad_kernel), dim3((unsigned)((Tm + TR - 1) / TR), B), 256, 0, ST, mel_out, mel_tgt, mel_len, B, Tm, n_mels,
                                                                              stats, fa, fb, fc, total, weight, S,
                                                                              dmel_out);
    if ((rc = fs2_check_launch())) return rc;
  }
  FS2_LAUNCH((ssim_finalize_kernel), (B + 63) / 64, 64, 0, ST, total, stats, S, B, Tm, n_mels, out, dmel_out);
  return fs2_check_launch();
}

extern "C" long long fs2_loss_ws_floats(int B) { return lw_words(B) + 8; }

extern "C" int fs2_loss_fused(const float* mel_out, const float* post_out, const float* mel_tgt, const float* log_dur_pred,
                              const int64_t* dur_tgt, const float* pitch_pred, const float* pitch_tgt,
                              const float* energy_pred, const float* energy_tgt, const int64_t* mel_len,
                              const int64_t* phon_len, int B, int Tp, int Tm, int n_mels, const float* w6, float* ws,
                              float* out8, float* dmel, float* dpost, float* ddur, float* dpitch, float* denergy,
                              void* stream) {
  REQUIRE(mel_out && post_out && mel_tgt && log_dur_pred && dur_tgt && pitch_pred && pitch_tgt && energy_pred &&
              energy_tgt && mel_len && phon_len && w6 && ws && out8,
          "fs2_loss_fused: null pointer");
  REQUIRE(n_mels % 4 == 0 && n_mels >= KS && n_mels <= MAXW && Tm >= KS,
          "fs2_loss_fused: needs Tm >= 11 and 11 <= n_mels <= 80, n_mels % 4 == 0");
  REQUIRE((dmel == nullptr) == (dpost == nullptr), "fs2_loss_fused: dmel and dpost come together");
  REQUIRE((long long)Tm * n_mels < (1LL << 31), "fs2_loss_fused: sample too large for 32-bit arg indices");
  int rc = upload_gauss();
  if (rc) return rc;
  static bool attr_set = false;
  const size_t smem = sizeof(float) * SSIM_SMEM_FLOATS;
  if (!attr_set) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(ssim_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  LossArgs a;
  a.mel_out = mel_out; a.post_out = post_out; a.mel_tgt = mel_tgt; a.log_dur = log_dur_pred;
  a.pitch_pred = pitch_pred; a.pitch_tgt = pitch_tgt; a.energy_pred = energy_pred; a.energy_tgt = energy_tgt;
  a.dur_tgt = dur_tgt; a.mel_len = mel_len; a.phon_len = phon_len;
  a.B = B; a.Tp = Tp; a.Tm = Tm; a.W = n_mels;
  for (int i = 0; i < 6; ++i) a.w[i] = w6[i];
  a.ws = ws; a.out = out8;
  a.dmel = dmel; a.dpost = dpost; a.ddur = ddur; a.dpitch = dpitch; a.denergy = denergy;
  const long long per4 = (long long)Tm * n_mels / 4;
  int gx = (int)((per4 + 256 * 4 - 1) / (256 * 4));
  gx = gx < 1 ? 1 : (gx > 32 ? 32 : gx);
  FS2_LAUNCH((loss_pass1_kernel), dim3(gx, B), 256, 0, ST, a);
  if ((rc = fs2_check_launch())) return rc;
  FS2_LAUNCH((ssim_fused_kernel), dim3((unsigned)((Tm + LT - 1) / LT), B), LTHREADS, smem, ST, a);
  if ((rc = fs2_check_launch())) return rc;
  FS2_LAUNCH((loss_finalize_kernel), dim3(FIN_X, B), 256, 0, ST, a);
  return fs2_check_launch();
}

extern "C" int fs2_loss_scale_grads(const float* g_dev, float* dmel, float* dpost, long long n_mel, float* ddur,
                                    float* dpitch, float* denergy, long long n_ph, void* stream) {
  REQUIRE(g_dev && dmel && dpost && ddur && dpitch && denergy, "fs2_loss_scale_grads: null pointer");
  FS2_LAUNCH((loss_scale_kernel), 296, 256, 0, ST, g_dev, dmel, dpost, n_mel, ddur, dpitch, denergy, n_ph);
  return fs2_check_launch();
}
