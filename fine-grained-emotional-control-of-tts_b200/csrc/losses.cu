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
