// Memory-bound row kernels of the FastSpeech2 path (everything that is not a GEMM or a loss):
// embedding + positional encoding, the LayerNorm family (residual / dropout / tanh / mask / scalar head fused,
// forward and backward), masked softmax with the reference's attn_mask quirk, speaker/intensity conditioning,
// average_over_durations, pitch/energy embed-add, the LengthRegulator (scan, expand, segment-sum backward),
// halo folding, column sums, weight packing and the fused AdamW.
//
// All activations live in the padded row space of common.cuh.  Rule for every producer: rect rows get values,
// halo rows get the reflect mirror (when `halo` > 0, width `halo`) and zeros otherwise, so GEMMs that sweep
// whole row ranges never meet stale data.
#include <math.h>
#include <stdlib.h>
#include "common.cuh"
#include "../../include/fs2_b200.h"

namespace {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int MAXV = 4;  // float4 per lane: C <= 512

inline int grid_for_rows(long long rows) {
  long long b = (rows + WARPS - 1) / WARPS;
  const long long cap = 148 * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T>
__device__ __forceinline__ void store_act4(void* base, long long off, float4 v) {
  st4(reinterpret_cast<T*>(base) + off, v);
}

// write one row (and its reflect mirrors) to the fp32 and act copies
template <typename TA>
__device__ __forceinline__ void store_row4(float* of, TA* oa, long long rowoff, int c, float4 v, int t, int T, int C,
                                           int halo) {
  if (of) st4(of + rowoff + c, v);
  if (oa) st4(oa + rowoff + c, v);
  if (halo > 0) {
    if (t >= 1 && t <= halo) {
      long long m = rowoff - 2LL * t * C;
      if (of) st4(of + m + c, v);
      if (oa) st4(oa + m + c, v);
    }
    if (t >= T - 1 - halo && t <= T - 2) {
      long long m = rowoff + 2LL * (T - 1 - t) * C;
      if (of) st4(of + m + c, v);
      if (oa) st4(oa + m + c, v);
    }
  }
}

// halo rows that no mirror write reaches are zeroed (callers guarantee T > halo)
__device__ __forceinline__ bool halo_row_needs_zero(int t, int T, int halo) {
  if (t >= 0 && t < T) return false;
  const int d = (t < 0) ? -t : (t - (T - 1));
  return !(halo > 0 && d <= halo && d < T);
}

// ------------------------------------------------------------------ embedding + posenc --
__global__ void count_nonpad_kernel(const int64_t* tokens, int B, int Tp, int pad_idx, int* lens) {
  pdl_wait();
  int b = blockIdx.x;
  int cnt = 0;
  for (int t = threadIdx.x; t < Tp; t += 32) cnt += (tokens[(long long)b * Tp + t] != pad_idx) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (threadIdx.x == 0) lens[b] = cnt;
}

template <typename TA>
__global__ void __launch_bounds__(THREADS) embed_posenc_kernel(const int64_t* tokens, const float* emb, const float* pe,
                                                               int B, int Tp, int D, int pad_idx, float* of, TA* oa) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int TP = Tp + 2 * FS2_PAD;
  const long long rows = (long long)B * TP;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * WARPS) {
    int b = (int)(r / TP), t = (int)(r - (long long)b * TP) - FS2_PAD;
    const bool in = t >= 0 && t < Tp;
    int64_t tok = in ? tokens[(long long)b * Tp + t] : 0;
    const bool live = in && tok != pad_idx;
    for (int c = lane * 4; c < D; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        float4 e = ld4(emb + tok * D + c), p = ld4(pe + (long long)t * D + c);
        v = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
      }
      if (of) st4(of + r * D + c, v);
      if (oa) st4(oa + r * D + c, v);
    }
  }
}

__global__ void embedding_bwd_kernel(const float* dx, const int64_t* tokens, int B, int Tp, int D, int pad_idx,
                                     float* demb) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * Tp;
  const int TP = Tp + 2 * FS2_PAD;
  for (long long i = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); i < rows; i += (long long)gridDim.x * WARPS) {
    int b = (int)(i / Tp), t = (int)(i - (long long)b * Tp);
    int64_t tok = tokens[i];
    if (tok == pad_idx) continue;
    const float* src = dx + ((long long)b * TP + FS2_PAD + t) * D;
    for (int c = lane; c < D; c += 32) atomicAdd(demb + tok * D + c, src[c]);
  }
}

// ----------------------------------------------------------------------- LayerNorm fwd --
// L2 prefetch of one row (C floats) of a row-major operand, one 128-byte line per lane.  The LayerNorm kernels run one
// row per warp per iteration with two warp reductions between the loads and the stores, so a warp has nothing in flight
// for most of an iteration; asking L2 for the warp's NEXT row while the current one is being reduced turns that
// row's DRAM latency into an L2 hit without spending registers (fs2_ln_tune switches it off for A/B timing).
__device__ __forceinline__ void prefetch_row_l2(const float* base, long long off, int C, int lane) {
  if (base != nullptr && lane * 32 < C) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off + lane * 32));
}

template <typename TA>
__global__ void __launch_bounds__(THREADS) ln_fwd_kernel(Fs2LnFwd p, int pf) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int T = p.T, C = p.C, TP = T + 2 * FS2_PAD;
  const long long rows = (long long)p.B * TP;
  const float invC = 1.0f / (float)C;
  TA* oa = (TA*)p.out_act;
  const unsigned long long bump = p.seed_dev ? mix64(*p.seed_dev) : 0ull;
  const DropKey db = drop_key(p.drop_b_p, p.drop_b_seed ^ bump), da = drop_key(p.drop_a_p, p.drop_a_seed ^ bump);
  int pf_first = 1;                                           // the first live row also asks for the rows in between
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * WARPS) {
    int b = (int)(r / TP), t = (int)(r - (long long)b * TP) - FS2_PAD;
    const long long ro = r * C;
    if (t < 0 || t >= T) {
      if (halo_row_needs_zero(t, T, p.halo)) {
        for (int c = lane * 4; c < C; c += 128) {
          if (p.out_f32) st4(p.out_f32 + ro + c, make_float4(0.f, 0.f, 0.f, 0.f));
          if (oa) st4(oa + ro + c, make_float4(0.f, 0.f, 0.f, 0.f));
        }
      }
      continue;
    }
    float4 z[MAXV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = lane * 4 + i * 128;
      z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C) {
        z[i] = ld4(p.x + ro + c);
        if (p.branch) {
          float4 br = ld4(p.branch + ro + c);
          float4 k = drop_scale4(db, (uint64_t)(ro + c) >> 2);
          z[i].x += br.x * k.x; z[i].y += br.y * k.y; z[i].z += br.z * k.z; z[i].w += br.w * k.w;
        }
        s += z[i].x + z[i].y + z[i].z + z[i].w;
      }
    }
    for (int k = pf_first; k <= pf; ++k) {                    // pf = prefetch distance in rows of this warp (0 = off)
      const long long rn = r + (long long)k * gridDim.x * WARPS;
      if (rn < rows) {
        prefetch_row_l2(p.x, rn * C, C, lane);
        prefetch_row_l2(p.branch, rn * C, C, lane);
        prefetch_row_l2(p.post_add, rn * C, C, lane);
      }
    }
    pf_first = pf;
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = lane * 4 + i * 128;
      if (c < C) {
        float a = z[i].x - mean, bq = z[i].y - mean, cq = z[i].z - mean, d = z[i].w - mean;
        q += a * a + bq * bq + cq * cq + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invC + p.eps);
    if (lane == 0) {
      if (p.mean) p.mean[r] = mean;
      if (p.rstd) p.rstd[r] = rstd;
    }
    const bool live = (p.lens == nullptr) || (t < p.lens[b]);
    float hd = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = lane * 4 + i * 128;
      if (c < C) {
        float4 g = ld4(p.gamma + c), be = ld4(p.beta + c);
        float4 u;
        u.x = (z[i].x - mean) * rstd * g.x + be.x;
        u.y = (z[i].y - mean) * rstd * g.y + be.y;
        u.z = (z[i].z - mean) * rstd * g.z + be.z;
        u.w = (z[i].w - mean) * rstd * g.w + be.w;
        if (p.tanh_act) { u.x = tanhf(u.x); u.y = tanhf(u.y); u.z = tanhf(u.z); u.w = tanhf(u.w); }
        float4 k = drop_scale4(da, (uint64_t)(ro + c) >> 2);
        u.x *= k.x; u.y *= k.y; u.z *= k.z; u.w *= k.w;
        if (!live) u = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.post_add) {
          float4 a = ld4(p.post_add + ro + c);
          u.x += a.x; u.y += a.y; u.z += a.z; u.w += a.w;
        }
        if (p.head_w) {
          float4 hw = ld4(p.head_w + c);
          hd += u.x * hw.x + u.y * hw.y + u.z * hw.z + u.w * hw.w;
        }
        store_row4<TA>(p.out_f32, oa, ro, c, u, t, T, C, p.halo);
      }
    }
    if (p.head_w) {
      hd = warp_sum(hd);
      if (lane == 0) p.head_out[(long long)b * T + t] = (hd + p.head_b[0]) * p.head_scale;
    }
  }
}

// ----------------------------------------------------------------------- LayerNorm bwd --
template <int NV>
__device__ __forceinline__ void block_reduce_cols(float4 (&a)[NV], float* out, int C, float (*red)[128], int warp,
                                                  int lane) {
  if (!out) return;   // uniform across the block
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    __syncthreads();
    red[warp][lane * 4 + 0] = a[i].x;
    red[warp][lane * 4 + 1] = a[i].y;
    red[warp][lane * 4 + 2] = a[i].z;
    red[warp][lane * 4 + 3] = a[i].w;
    __syncthreads();
    if (threadIdx.x < 128) {
      int c = i * 128 + threadIdx.x;
      if (c < C) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) v += red[w][threadIdx.x];
        atomicAdd(out + c, v);
      }
    }
  }
}

// grad wrt the kernel's `out` for row (b,t): dy + fold(dy2) + dhead*head_w*scale, then back through
// mask, drop_a, tanh, affine, normalisation, (ReLU of x), and the branch dropout.
// NV = float4 per lane (ceil(C / 128)): 3 for the 384-wide model rows, 4 for the PostNet (512), 1 for n_mels (80);
// HEAD = the variance predictors' 384 -> 1 output layer is folded in.  Both only size the register arrays.
// HEAD doubles as the "rare features" switch: only the variance-predictor / PostNet calls use the 384 -> 1 head, tanh,
// ReLU-of-x or the dropout AFTER the norm; the FFT-block LayerNorms (24 of the ~30 calls per step) run the HEAD = false
// instantiation in which all of that is compiled out (3.2 k -> ~1.6 k instructions per kernel).
template <typename TA, int NV, bool HEAD>
__global__ void __launch_bounds__(THREADS, (NV <= 3 && !HEAD) ? 3 : 2) ln_bwd_kernel(Fs2LnBwd p, int pf) {
  pdl_wait();
  __shared__ float red[WARPS][128];
  // dgamma / dbeta partial sums live in shared memory, one private slab per warp (a lane only ever touches its own
  // columns, so no synchronisation is needed inside the row loop): 24 registers fewer than register accumulators,
  // which is the difference between 2 and 3 resident CTAs per SM for the 384-wide rows.
  __shared__ __align__(16) float acc_g[WARPS][NV * 128];
  __shared__ __align__(16) float acc_b[WARPS][NV * 128];
  constexpr bool CS_OK = NV <= 3;                                    // third slab only where static smem allows it
  __shared__ __align__(16) float acc_d[CS_OK ? WARPS : 1][CS_OK ? NV * 128 : 4];   // column sums of dact (bias gradient)
  const bool want_cs = CS_OK && p.dact_colsum != nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int T = p.T, C = p.C, TP = T + 2 * FS2_PAD;
  const long long rows = (long long)p.B * TP;
  const float invC = 1.0f / (float)C;
  TA* da_out = (TA*)p.dact;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    *reinterpret_cast<float4*>(&acc_g[warp][lane * 4 + i * 128]) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(&acc_b[warp][lane * 4 + i * 128]) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (CS_OK) *reinterpret_cast<float4*>(&acc_d[CS_OK ? warp : 0][lane * 4 + i * 128]) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const unsigned long long bump = p.seed_dev ? mix64(*p.seed_dev) : 0ull;
  const DropKey db = drop_key(p.drop_b_p, p.drop_b_seed ^ bump), da = drop_key(p.drop_a_p, p.drop_a_seed ^ bump);
  float4 dhw[HEAD ? NV : 1];
#pragma unroll
  for (int i = 0; i < (HEAD ? NV : 1); ++i) dhw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float dhb = 0.f;
  int pf_first = 1;
  for (long long r = (long long)blockIdx.x * WARPS + warp; r < rows; r += (long long)gridDim.x * WARPS) {
    int b = (int)(r / TP), t = (int)(r - (long long)b * TP) - FS2_PAD;
    const long long ro = r * C;
    if (t < 0 || t >= T) {
      for (int c = lane * 4; c < C; c += 128) {
        if (p.dx_f32) st4(p.dx_f32 + ro + c, make_float4(0.f, 0.f, 0.f, 0.f));
        if (da_out) st4(da_out + ro + c, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      continue;
    }
    const bool live = (p.lens == nullptr) || (t < p.lens[b]);
    const float mean = p.mean ? p.mean[r] : 0.f, rstd = p.rstd[r];
    const float dh = p.dhead ? p.dhead[(long long)b * T + t] * p.head_scale : 0.f;
    const int f = p.dy2_fold;
    const long long m1 = (f > 0 && t >= 1 && t <= f) ? -2LL * t * C : 0;
    const long long m2 = (f > 0 && t >= T - 1 - f && t <= T - 2) ? 2LL * (T - 1 - t) * C : 0;
    float4 xh[NV], gx[NV];
    float s1 = 0.f, s2 = 0.f;
    unsigned keep_b = 0u;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = lane * 4 + i * 128;
      xh[i] = gx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C) {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!HEAD && p.y) {
          // the forward never materialised x + branch (fs2_gemm_ln_tc): x_hat comes from its output below; only the
          // branch-dropout mask is regenerated here
          z = ld4(p.y + ro + c);
          float4 k = drop_scale4(db, (uint64_t)(ro + c) >> 2);
          keep_b |= ((k.x != 0.f ? 1u : 0u) | (k.y != 0.f ? 2u : 0u) | (k.z != 0.f ? 4u : 0u) | (k.w != 0.f ? 8u : 0u)) << (4 * i);
        } else {
          z = ld4(p.x + ro + c);
          if (p.branch) {
            float4 br = ld4(p.branch + ro + c);
            float4 k = drop_scale4(db, (uint64_t)(ro + c) >> 2);
            z.x += br.x * k.x; z.y += br.y * k.y; z.z += br.z * k.z; z.w += br.w * k.w;
            // remember the keep bits: the gradient of the branch needs the same mask again (one RNG evaluation, not two)
            keep_b |= ((k.x != 0.f ? 1u : 0u) | (k.y != 0.f ? 2u : 0u) | (k.z != 0.f ? 4u : 0u) | (k.w != 0.f ? 8u : 0u)) << (4 * i);
          }
        }
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.dy) g = ld4(p.dy + ro + c);
        if (p.dy2) {
          float4 a = ld4(p.dy2 + ro + c);
          g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
          if (m1) { a = ld4(p.dy2 + ro + m1 + c); g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w; }
          if (m2) { a = ld4(p.dy2 + ro + m2 + c); g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w; }
          if (p.dy3) {                    // second partial result of a split-K dgrad: same rows, same halo fold
            a = ld4(p.dy3 + ro + c);
            g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
            if (m1) { a = ld4(p.dy3 + ro + m1 + c); g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w; }
            if (m2) { a = ld4(p.dy3 + ro + m2 + c); g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w; }
          }
        }
        float4 gam = ld4(p.gamma + c), bet = ld4(p.beta + c);
        float4 h;
        if (!HEAD && p.y) {
          h.x = gam.x != 0.f ? __fdividef(z.x - bet.x, gam.x) : 0.f;
          h.y = gam.y != 0.f ? __fdividef(z.y - bet.y, gam.y) : 0.f;
          h.z = gam.z != 0.f ? __fdividef(z.z - bet.z, gam.z) : 0.f;
          h.w = gam.w != 0.f ? __fdividef(z.w - bet.w, gam.w) : 0.f;
        } else {
          h.x = (z.x - mean) * rstd; h.y = (z.y - mean) * rstd; h.z = (z.z - mean) * rstd; h.w = (z.w - mean) * rstd;
        }
        float4 u = make_float4(h.x * gam.x + bet.x, h.y * gam.y + bet.y, h.z * gam.z + bet.z, h.w * gam.w + bet.w);
        if (HEAD && p.tanh_act) { u.x = tanhf(u.x); u.y = tanhf(u.y); u.z = tanhf(u.z); u.w = tanhf(u.w); }
        float4 k = HEAD ? drop_scale4(da, (uint64_t)(ro + c) >> 2) : make_float4(1.f, 1.f, 1.f, 1.f);
        if (HEAD && p.head_w) {
          float4 hw = ld4(p.head_w + c);
          g.x += dh * hw.x; g.y += dh * hw.y; g.z += dh * hw.z; g.w += dh * hw.w;
          if (live) {   // d head_w = dhead * out, out = u * drop_a (masked rows contribute 0)
            dhw[HEAD ? i : 0].x += dh * u.x * k.x; dhw[HEAD ? i : 0].y += dh * u.y * k.y; dhw[HEAD ? i : 0].z += dh * u.z * k.z; dhw[HEAD ? i : 0].w += dh * u.w * k.w;
          }
        }
        if (!live) g = make_float4(0.f, 0.f, 0.f, 0.f);
        g.x *= k.x; g.y *= k.y; g.z *= k.z; g.w *= k.w;
        if (HEAD && p.tanh_act) {
          g.x *= (1.f - u.x * u.x); g.y *= (1.f - u.y * u.y); g.z *= (1.f - u.z * u.z); g.w *= (1.f - u.w * u.w);
        }
        {
          float4* ag = reinterpret_cast<float4*>(&acc_g[warp][c]);
          float4* ab = reinterpret_cast<float4*>(&acc_b[warp][c]);
          float4 sg = *ag, sb = *ab;
          sg.x += g.x * h.x; sg.y += g.y * h.y; sg.z += g.z * h.z; sg.w += g.w * h.w;
          sb.x += g.x; sb.y += g.y; sb.z += g.z; sb.w += g.w;
          *ag = sg;
          *ab = sb;
        }
        g.x *= gam.x; g.y *= gam.y; g.z *= gam.z; g.w *= gam.w;
        s1 += g.x + g.y + g.z + g.w;
        s2 += g.x * h.x + g.y * h.y + g.z * h.z + g.w * h.w;
        xh[i] = h;
        gx[i] = g;
      }
    }
    for (int k = pf_first; k <= pf; ++k) {                    // pf = prefetch distance in rows of this warp (0 = off)
      const long long rn = r + (long long)k * gridDim.x * WARPS;
      if (rn < rows) {
        prefetch_row_l2(p.x, rn * C, C, lane);
        prefetch_row_l2(p.branch, rn * C, C, lane);
        prefetch_row_l2(p.dy, rn * C, C, lane);
        prefetch_row_l2(p.dy2, rn * C, C, lane);
        prefetch_row_l2(p.dy3, rn * C, C, lane);
        prefetch_row_l2(p.y, rn * C, C, lane);
      }
    }
    pf_first = pf;
    if (HEAD && p.head_w && lane == 0) dhb += dh;
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int c = lane * 4 + i * 128;
      if (c < C) {
        float4 dz;
        dz.x = rstd * (gx[i].x - s1 - xh[i].x * s2);
        dz.y = rstd * (gx[i].y - s1 - xh[i].y * s2);
        dz.z = rstd * (gx[i].z - s1 - xh[i].z * s2);
        dz.w = rstd * (gx[i].w - s1 - xh[i].w * s2);
        if (HEAD && p.relu_x) {
          float4 x = ld4(p.x + ro + c);
          if (!(x.x > 0.f)) dz.x = 0.f;
          if (!(x.y > 0.f)) dz.y = 0.f;
          if (!(x.z > 0.f)) dz.z = 0.f;
          if (!(x.w > 0.f)) dz.w = 0.f;
        }
        if (p.dx_f32) st4(p.dx_f32 + ro + c, dz);
        if (da_out) {
          if (p.branch || (!HEAD && p.y)) {
            const float sk = db.p > 0.f ? 1.0f / (1.0f - db.p) : 1.0f;
            const unsigned kb = keep_b >> (4 * i);
            dz.x *= (kb & 1u) ? sk : 0.f; dz.y *= (kb & 2u) ? sk : 0.f; dz.z *= (kb & 4u) ? sk : 0.f; dz.w *= (kb & 8u) ? sk : 0.f;
          }
          st4(da_out + ro + c, dz);
          if (want_cs) {
            float4* ad = reinterpret_cast<float4*>(&acc_d[CS_OK ? warp : 0][CS_OK ? c : 0]);
            float4 sd = *ad;
            sd.x += dz.x; sd.y += dz.y; sd.z += dz.z; sd.w += dz.w;
            *ad = sd;
          }
        }
      }
    }
  }
  // block reduction of the parameter gradients, one atomicAdd per column per block
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += THREADS) {
    float vg = 0.f, vb = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) { vg += acc_g[w][c]; vb += acc_b[w][c]; }
    if (p.dgamma) atomicAdd(p.dgamma + c, vg);
    if (p.dbeta) atomicAdd(p.dbeta + c, vb);
    if (want_cs) {
      float vd = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) vd += acc_d[CS_OK ? w : 0][CS_OK ? c : 0];
      atomicAdd(p.dact_colsum + c, vd);
    }
  }
  if (HEAD) block_reduce_cols(dhw, p.dhead_w, C, red, warp, lane);
  if (HEAD && p.dhead_b) {
    __syncthreads();
    if (lane == 0) red[warp][0] = dhb;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = 0.f;
      for (int w = 0; w < WARPS; ++w) v += red[w][0];
      atomicAdd(p.dhead_b, v);
    }
  }
}

// ------------------------------------------------------------------------------ softmax --
__device__ __forceinline__ int quirk_kv(const int* lens, int B, int H, int bh) {
  // model.py:338-343 / 414-419: the (h*B+b)-ordered attn_mask is read as (b*H+h) by nn.MultiheadAttention,
  // so (b,h) masks pad[b] U pad[(b*H+h) % B]; with tail padding that is a min of two lengths.
  int b = bh / H;
  int o = bh % B;
  return min(lens[b], lens[o]);
}

template <typename TA>
__global__ void __launch_bounds__(THREADS) softmax_fwd_kernel(const float* S, const int* lens, int B, int H, int T,
                                                              int ldk, float scale, DropCfg dc, const unsigned long long* seed_dev, TA* P, TA* Pd) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * H * T;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * WARPS) {
    const int bh = (int)(r / T);
    const int kv = quirk_kv(lens, B, H, bh);
    // the same per-element dropout hash as the flash-style kernels (flash_attention.cu), so both paths draw one mask
    const ADrop dr = adrop_make(dc.p, dc.seed, seed_dev, bh);
    const uint32_t tpart = adrop_row(dr, (int)(r - (long long)bh * T));
    const float* s = S + r * ldk;
    float mx = -INFINITY;
    for (int c = lane * 4; c < kv; c += 128) {
      float4 v = ld4(s + c);
      mx = fmaxf(mx, v.x);
      if (c + 1 < kv) mx = fmaxf(mx, v.y);
      if (c + 2 < kv) mx = fmaxf(mx, v.z);
      if (c + 3 < kv) mx = fmaxf(mx, v.w);
    }
    mx = warp_max(mx) * scale;
    float sum = 0.f;
    for (int c = lane * 4; c < kv; c += 128) {
      float4 v = ld4(s + c);
      sum += __expf(v.x * scale - mx);
      if (c + 1 < kv) sum += __expf(v.y * scale - mx);
      if (c + 2 < kv) sum += __expf(v.z * scale - mx);
      if (c + 3 < kv) sum += __expf(v.w * scale - mx);
    }
    const float inv = 1.0f / warp_sum(sum);
    for (int c = lane * 4; c < ldk; c += 128) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < kv) {
        float4 v = ld4(s + c);
        o.x = __expf(v.x * scale - mx) * inv;
        if (c + 1 < kv) o.y = __expf(v.y * scale - mx) * inv;
        if (c + 2 < kv) o.z = __expf(v.z * scale - mx) * inv;
        if (c + 3 < kv) o.w = __expf(v.w * scale - mx) * inv;
      }
      st4(P + r * ldk + c, o);
      if (Pd) {
        const float4 k = adrop_scale4(dr, tpart, c);
        st4(Pd + r * ldk + c, make_float4(o.x * k.x, o.y * k.y, o.z * k.z, o.w * k.w));
      }
    }
  }
}

template <typename TA>
__global__ void __launch_bounds__(THREADS) softmax_bwd_kernel(const TA* P, const float* dPd, const int* lens, int B,
                                                              int H, int T, int ldk, float scale, DropCfg dc,
                                                              const unsigned long long* seed_dev, TA* dS) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long rows = (long long)B * H * T;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * WARPS) {
    const int bh = (int)(r / T);
    const int kv = quirk_kv(lens, B, H, bh);
    const ADrop dr = adrop_make(dc.p, dc.seed, seed_dev, bh);
    const uint32_t tpart = adrop_row(dr, (int)(r - (long long)bh * T));
    const long long ro = r * ldk;
    float dot = 0.f;
    for (int c = lane * 4; c < kv; c += 128) {
      float4 pv = ld4(P + ro + c), g = ld4(dPd + ro + c);
      const float4 k = adrop_scale4(dr, tpart, c);
      dot += pv.x * g.x * k.x + pv.y * g.y * k.y + pv.z * g.z * k.z + pv.w * g.w * k.w;   // P is 0 beyond kv
    }
    dot = warp_sum(dot);
    for (int c = lane * 4; c < ldk; c += 128) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < kv) {
        float4 pv = ld4(P + ro + c), g = ld4(dPd + ro + c);
        const float4 k = adrop_scale4(dr, tpart, c);
        o.x = scale * pv.x * (g.x * k.x - dot);
        o.y = scale * pv.y * (g.y * k.y - dot);
        o.z = scale * pv.z * (g.z * k.z - dot);
        o.w = scale * pv.w * (g.w * k.w - dot);
      }
      st4(dS + ro + c, o);
    }
  }
}

// ------------------------------------------------------------- speaker / intensity cond --
__global__ void spk_proj_kernel(const float* Wcat, const float* spk_emb, const int64_t* speakers, int B, int D,
                                float* sp) {
  pdl_wait();
  // sp[b, c] = sum_e Wcat[c, D + e] * spk_emb[speakers[b], e]; one warp per (b, c)
  const int lane = threadIdx.x & 31;
  const long long n = (long long)B * D;
  const int ldw = 2 * D + 5;
  for (long long i = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); i < n; i += (long long)gridDim.x * WARPS) {
    int b = (int)(i / D), c = (int)(i - (long long)b * D);
    const float* w = Wcat + (long long)c * ldw + D;
    const float* e = spk_emb + speakers[b] * D;
    float s = 0.f;
    for (int k = lane; k < D; k += 32) s += w[k] * e[k];
    s = warp_sum(s);
    if (lane == 0) sp[i] = s;
  }
}

template <typename TA>
__global__ void __launch_bounds__(THREADS) cond_finish_kernel(const float* G, const float* Wcat, const float* sp,
                                                              const float* intensity, const int* lens, int B, int Tp,
                                                              int D, float* yf, TA* ya, int halo) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int TP = Tp + 2 * FS2_PAD;
  const long long rows = (long long)B * TP;
  const int ldw = 2 * D + 5;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * WARPS) {
    int b = (int)(r / TP), t = (int)(r - (long long)b * TP) - FS2_PAD;
    const long long ro = r * D;
    if (t < 0 || t >= Tp) {
      if (halo_row_needs_zero(t, Tp, halo))
        for (int c = lane * 4; c < D; c += 128) {
          if (yf) st4(yf + ro + c, make_float4(0.f, 0.f, 0.f, 0.f));
          if (ya) st4(ya + ro + c, make_float4(0.f, 0.f, 0.f, 0.f));
        }
      continue;
    }
    const bool live = t < lens[b];
    float iv[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) iv[i] = intensity[((long long)b * Tp + t) * 5 + i];
    for (int c = lane * 4; c < D; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        float4 g = ld4(G + ro + c), s = ld4(sp + (long long)b * D + c);
        float o[4] = {g.x + s.x, g.y + s.y, g.z + s.z, g.w + s.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float* wi = Wcat + (long long)(c + q) * ldw + 2 * D;
#pragma unroll
          for (int i = 0; i < 5; ++i) o[q] += wi[i] * iv[i];
        }
        v = make_float4(o[0], o[1], o[2], o[3]);
      }
      store_row4<TA>(yf, ya, ro, c, v, t, Tp, D, halo);
    }
  }
}

// dsum[b,c] = sum_t dy[b,t,c];  dWi[c,i] += sum_{b,t} dy[b,t,c] * int[b,t,i]     (dy already masked)
constexpr int COND_ROWS = 16;     // phoneme rows per CTA of cond_bwd_rows_kernel
__global__ void cond_bwd_rows_kernel(const float* __restrict__ dy, const float* __restrict__ intensity, int B, int Tp,
                                     int D, float* dsum, float* dWcat) {
  pdl_wait();
  // grid (row chunks, B): every thread owns columns c, c+128, ... and COND_ROWS rows; dsum must be zero on entry
  __shared__ float iv[COND_ROWS][5];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * COND_ROWS;
  const int nt = min(COND_ROWS, Tp - t0);
  const int TP = Tp + 2 * FS2_PAD;
  const int ldw = 2 * D + 5;
  for (int i = threadIdx.x; i < nt * 5; i += blockDim.x) iv[i / 5][i % 5] = intensity[((long long)b * Tp + t0) * 5 + i];
  __syncthreads();
  const float* src = dy + ((long long)b * TP + FS2_PAD + t0) * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float g[COND_ROWS];
#pragma unroll
    for (int t = 0; t < COND_ROWS; ++t) g[t] = (t < nt) ? src[(long long)t * D + c] : 0.f;     // independent loads
    float s = 0.f, wi[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < COND_ROWS; ++t) {
      s += g[t];
      if (t < nt) {
#pragma unroll
        for (int i = 0; i < 5; ++i) wi[i] += g[t] * iv[t][i];
      }
    }
    atomicAdd(dsum + (long long)b * D + c, s);
#pragma unroll
    for (int i = 0; i < 5; ++i) atomicAdd(dWcat + (long long)c * ldw + 2 * D + i, wi[i]);
  }
}
// dWs[c,e] += sum_b dsum[b,c]*emb[spk[b],e];   dspk_emb[spk[b],e] += sum_c Ws[c,e]*dsum[b,c]
__global__ void cond_bwd_spk_kernel(const float* __restrict__ dsum, const float* __restrict__ Wcat,
                                    const float* __restrict__ spk_emb, const int64_t* __restrict__ speakers, int B, int D,
                                    float* dWcat, float* dspk_emb) {
  pdl_wait();
  const int ldw = 2 * D + 5;
  // dWcat[c, D + e] += sum_b dsum[b, c] * spk_emb[spk[b], e]: one item per (c, e, chunk of 8 utterances)
  const int bch = (B + 7) / 8;
  const long long n = (long long)D * D * bch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % D);
    const long long q = i / D;
    const int c = (int)(q % D), b0 = (int)(q / D) * 8;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = b0 + j;
      if (b < B) s += dsum[(long long)b * D + c] * spk_emb[speakers[b] * D + e];
    }
    atomicAdd(dWcat + (long long)c * ldw + D + e, s);
  }
  // dspk_emb[spk[b], e] += sum_c Wcat[c, D + e] * dsum[b, c]: one item per (b, e, chunk of 32 input channels)
  const int cch = (D + 31) / 32;
  const long long m = (long long)B * D * cch;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i % D);
    const long long q = i / D;
    const int b = (int)(q % B), c0 = (int)(q / B) * 32;
    float s = 0.f;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const int c = c0 + j;
      if (c < D) s += Wcat[(long long)c * ldw + D + e] * dsum[(long long)b * D + c];
    }
    atomicAdd(dspk_emb + speakers[b] * D + e, s);
  }
}

// ------------------------------------------------------------- average_over_durations --
// Restates speechbrain's prefix-sum formulation on torch's CPU semantics: cumsum accumulates in double and
// rounds every prefix to fp32; segment sum = difference of two fp32 prefixes; mean over non-zero frames.
__global__ void avg_over_durations_kernel(const float* values, const int64_t* durs, int B, int Tp, int Tm, float* avg,
                                          int* starts, int* ends, int* nz) {
  pdl_wait();
  extern __shared__ unsigned char smraw[];
  float* vc = (float*)smraw;            // Tm+1 prefix sums
  int* nc = (int*)(vc + Tm + 1);        // Tm+1 non-zero counts
  int* de = nc + Tm + 1;                // Tp inclusive duration ends
  const int b = blockIdx.x;
  const float* v = values + (long long)b * Tm;
  for (int i = threadIdx.x; i < Tm; i += blockDim.x) vc[i + 1] = v[i];
  for (int i = threadIdx.x; i < Tp; i += blockDim.x) de[i] = (int)durs[(long long)b * Tp + i];
  __syncthreads();
  // prefix sums with a double accumulator (torch CPU cumsum semantics, fp32 prefixes): warp 0 scans the values, warp 1
  // the durations.  Each lane sums a contiguous chunk, the lane totals are scanned with shuffles, then the chunk is
  // replayed with its offset -- ~Tm/32 dependent adds instead of Tm.
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (wid == 0) {
    const int per = (Tm + 31) / 32;
    const int i0 = 1 + lane * per, i1 = min(Tm + 1, i0 + per);
    double acc = 0.0;
    int cnt = 0;
    for (int i = i0; i < i1; ++i) {
      const float x = vc[i];
      acc += (double)x;
      cnt += (x != 0.0f) ? 1 : 0;
    }
    double off = acc;
    int coff = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double a = __shfl_up_sync(0xffffffffu, off, o);
      const int c = __shfl_up_sync(0xffffffffu, coff, o);
      if (lane >= o) { off += a; coff += c; }
    }
    off -= acc;                      // exclusive offsets of this lane's chunk
    coff -= cnt;
    if (lane == 0) { vc[0] = 0.f; nc[0] = 0; }
    for (int i = i0; i < i1; ++i) {
      const float x = vc[i];
      off += (double)x;
      coff += (x != 0.0f) ? 1 : 0;
      vc[i] = (float)off;
      nc[i] = coff;
    }
  } else if (wid == 1) {
    const int per = (Tp + 31) / 32;
    const int i0 = lane * per, i1 = min(Tp, i0 + per);
    int acc = 0;
    for (int i = i0; i < i1; ++i) acc += de[i];
    int off = acc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, off, o);
      if (lane >= o) off += a;
    }
    off -= acc;
    for (int i = i0; i < i1; ++i) { off += de[i]; de[i] = off; }
  }
  __syncthreads();
  for (int p = threadIdx.x; p < Tp; p += blockDim.x) {
    int e = de[p], s = p ? de[p - 1] : 0;
    int ec = min(max(e, 0), Tm), sc = min(max(s, 0), Tm);
    float sum = vc[ec] - vc[sc];
    int n = nc[ec] - nc[sc];
    long long o = (long long)b * Tp + p;
    avg[o] = (n == 0) ? 0.f : sum / (float)n;
    if (starts) starts[o] = s;
    if (ends) ends[o] = e;
    if (nz) nz[o] = n;
  }
}

// ----------------------------------------------------------------- pitch/energy embed-add --
__device__ __forceinline__ int reflect_idx(int i, int T) {
  if (i < 0) i = -i;
  if (i >= T) i = 2 * (T - 1) - i;
  return i;
}

template <typename TA>
__global__ void __launch_bounds__(THREADS) embed_add_kernel(const float* x, const float* contour, const float* w,
                                                            const float* bias, int ksize, const int* lens, int B, int Tp,
                                                            int D, float* yf, TA* ya, int halo) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int TP = Tp + 2 * FS2_PAD;
  const long long rows = (long long)B * TP;
  const int pad = (ksize - 1) / 2;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * WARPS) {
    int b = (int)(r / TP), t = (int)(r - (long long)b * TP) - FS2_PAD;
    const long long ro = r * D;
    if (t < 0 || t >= Tp) {
      for (int c = lane * 4; c < D; c += 128) {
        if (yf) st4(yf + ro + c, make_float4(0.f, 0.f, 0.f, 0.f));
        if (ya && halo_row_needs_zero(t, Tp, halo)) st4(ya + ro + c, make_float4(0.f, 0.f, 0.f, 0.f));
      }
      continue;
    }
    float cv[9];
    for (int j = 0; j < ksize; ++j) cv[j] = contour[(long long)b * Tp + reflect_idx(t + j - pad, Tp)];
    const bool live = t < lens[b];
    for (int c = lane * 4; c < D; c += 128) {
      float4 xv = ld4(x + ro + c), bv = ld4(bias + c);
      float o[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        for (int j = 0; j < ksize; ++j) o[q] += w[(long long)(c + q) * ksize + j] * cv[j];
      float4 y = make_float4(xv.x + o[0], xv.y + o[1], xv.z + o[2], xv.w + o[3]);
      if (yf) st4(yf + ro + c, y);
      if (ya) store_row4<TA>(nullptr, ya, ro, c, live ? y : make_float4(0.f, 0.f, 0.f, 0.f), t, Tp, D, halo);
    }
  }
}

// dw[c,j] += sum_{b,t} dy[b,t,c]*contour_r[b,t+j-p]; dbias[c] += sum dy   (all rect rows, unmasked)
__global__ void embed_add_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ contour, int ksize, int B,
                                     int Tp, int D, float* dw, float* dbias) {
  pdl_wait();
  const int TP = Tp + 2 * FS2_PAD;
  const int pad = (ksize - 1) / 2;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float sb = 0.f;
  const long long total = (long long)B * Tp;
  for (long long i0 = blockIdx.y; i0 < total; i0 += 4LL * gridDim.y) {
    float g[4];
    int bb[4], tt[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {                      // four independent row loads in flight
      const long long i = i0 + (long long)u * gridDim.y;
      g[u] = 0.f;
      bb[u] = 0;
      tt[u] = 0;
      if (i < total) {
        bb[u] = (int)(i / Tp);
        tt[u] = (int)(i - (long long)bb[u] * Tp);
        g[u] = dy[((long long)bb[u] * TP + FS2_PAD + tt[u]) * D + c];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      sb += g[u];
      for (int j = 0; j < ksize; ++j) acc[j] += g[u] * contour[(long long)bb[u] * Tp + reflect_idx(tt[u] + j - pad, Tp)];
    }
  }
  for (int j = 0; j < ksize; ++j) atomicAdd(dw + (long long)c * ksize + j, acc[j]);
  atomicAdd(dbias + c, sb);
}

// ---------------------------------------------------------------------- LengthRegulator --
__global__ void dur_decode_kernel(const float* log_dur, long long n, float* fdur) {
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fdur[i] = fmaxf(expm1f(log_dur[i]), 0.f);   // model.py:372-375
}

__global__ void lr_prepare_kernel(const int64_t* dur, const float* fdur, float pace, int B, int Tp, int* ends,
                                  int* mel_lens) {
  pdl_wait();
  const int b = blockIdx.x, lane = threadIdx.x;
  int carry = 0;
  for (int base = 0; base < Tp; base += 32) {
    int p = base + lane;
    int fr = 0;
    if (p < Tp) {
      // speechbrain upsample: (pace * durs).long() -- fp32 product, truncation toward zero
      float d = fdur ? fdur[(long long)b * Tp + p] : (float)dur[(long long)b * Tp + p];
      fr = (int)(long long)__fmul_rn(pace, d);
      if (fr < 0) fr = 0;   // repeat_interleave rejects negatives; durations are >= 0 by construction
    }
    int s = fr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    if (p < Tp) ends[(long long)b * Tp + p] = carry + s;
    carry += __shfl_sync(0xffffffffu, s, 31);
  }
  if (lane == 0) mel_lens[b] = carry;
}

// mel_lens -> int64 copy for the host + a check that the caller-supplied frame count equals max(mel_lens)
__global__ void lr_finalize_kernel(const int* mel_lens, int B, int Tm_expected, long long* out_i64, int* flag) {
  pdl_wait();
  int mx = 0;
  for (int b = threadIdx.x; b < B; b += 32) {
    int v = mel_lens[b];
    out_i64[b] = v;
    mx = max(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (threadIdx.x == 0 && mx != Tm_expected) atomicExch(flag, mx == 0 ? -1 : mx);
}

// LengthRegulator expand / segment-sum kernels.  One CTA = LR_ROWS consecutive output rows of ONE item: the item's
// duration prefix sums (`ends`, Tp ints) are staged in shared memory once, eight lanes of a warp binary-search eight
// rows at the same time (so the dependent-load chain is paid once per 8 rows, out of smem), and the row copies are
// issued four rows at a time so that every lane keeps >= 4 independent 16-byte loads in flight.
constexpr int LR_RPW = 8;                    // rows per warp
constexpr int LR_ROWS = WARPS * LR_RPW;      // rows per CTA

// first p with e[p] > f  (e = inclusive prefix sums of the frame counts), -1 when f is outside [0, total)
__device__ __forceinline__ int lr_search(const int* e, int Tp, int f, int total) {
  if (f < 0 || f >= total) return -1;
  int lo = 0, hi = Tp - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (e[mid] > f) hi = mid; else lo = mid + 1;
  }
  return lo;
}

template <typename TA, int RB>
__global__ void __launch_bounds__(THREADS) lr_expand_kernel(const float* __restrict__ in, int in_pitch, int in_off,
                                                            const int* __restrict__ ends, const int* __restrict__ mel_lens,
                                                            const float* __restrict__ pe, int B, int Tp, int Tm, int D,
                                                            float* __restrict__ of, TA* __restrict__ oa, int out_pitch,
                                                            int out_off, int* __restrict__ frame2ph) {
  pdl_wait();
  extern __shared__ int s_ends[];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < Tp; i += THREADS) s_ends[i] = ends[(long long)b * Tp + i];
  __syncthreads();
  const int total = min(min(mel_lens[b], Tm), s_ends[Tp - 1]);
  const int r0 = blockIdx.x * LR_ROWS + warp * LR_RPW;          // first row (inside the item's out_pitch rows)
  int my_idx = -1;
  if (lane < LR_RPW) {
    const int f = r0 + lane - out_off;
    my_idx = lr_search(s_ends, Tp, f, total);
    if (frame2ph && r0 + lane < out_pitch && f >= 0 && f < Tm) frame2ph[(long long)b * Tm + f] = my_idx;
  }
  const float* in_b = in + ((long long)b * in_pitch + in_off) * D;
  // RB rows in flight per warp (RB = 4: 78 registers -> 3 CTAs / SM; selected by measurement, see fs2_lr_tune)
#pragma unroll
  for (int j0 = 0; j0 < LR_RPW; j0 += RB) {
    int idx[RB];
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      idx[j] = __shfl_sync(0xffffffffu, my_idx, j0 + j);
      if (r0 + j0 + j >= out_pitch) idx[j] = -2;                // row does not exist
    }
    if (idx[0] == -2) break;
    for (int c = lane * 4; c < D; c += 128) {
      float4 v[RB], pv[RB];
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        pv[j] = v[j];
        if (idx[j] >= 0) {
          v[j] = ld4(in_b + (long long)idx[j] * D + c);
          if (pe) pv[j] = ld4(pe + (long long)(r0 + j0 + j - out_off) * D + c);
        }
      }
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        if (idx[j] == -2) continue;
        v[j].x += pv[j].x; v[j].y += pv[j].y; v[j].z += pv[j].z; v[j].w += pv[j].w;
        const long long ro = ((long long)b * out_pitch + r0 + j0 + j) * D + c;
        if (of) st4(of + ro, v[j]);
        if (oa) st4(oa + ro, v[j]);
      }
    }
  }
}

// Plain fp32 expansion (no pos-enc, no bf16 copy: BASELINE configs[1], the standalone LengthRegulator) on the bulk-copy
// engine.  One thread per output row, ROWS rows of one item per CTA.  The first row of every phoneme run inside the CTA
// ("leader") pulls the phoneme's D floats into its shared-memory slot with one cp.async.bulk (global -> shared,
// completion on an mbarrier), then every row pushes its run's slot -- or a zeroed slot for rows past the item's length
// -- to global memory with one cp.async.bulk (shared -> global).  Source rows are read once per CTA instead of once per
// frame, no data passes through registers, and a CTA keeps ROWS * D * 4 bytes of stores in flight from ~10 instructions
// per row.  Same bytes out as lr_expand_kernel (bit-exact copy), same frame2ph map.
__device__ __forceinline__ uint32_t lrb_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROWS>
__global__ void __launch_bounds__(ROWS) lr_expand_bulk_kernel(const float* __restrict__ in, int in_pitch, int in_off,
                                                              const int* __restrict__ ends, const int* __restrict__ mel_lens,
                                                              int B, int Tp, int Tm, int D, float* __restrict__ of,
                                                              int out_pitch, int out_off, int* __restrict__ frame2ph) {
  pdl_wait();
  extern __shared__ __align__(128) unsigned char lrb_raw[];
  const uint32_t row_bytes = (uint32_t)D * 4u;
  uint64_t* bar = reinterpret_cast<uint64_t*>(lrb_raw);                 // 16 bytes reserved
  float* zero_row = reinterpret_cast<float*>(lrb_raw + 16);
  unsigned char* slots = lrb_raw + 16 + row_bytes;
  int* s_ends = reinterpret_cast<int*>(slots + (size_t)ROWS * row_bytes);
  const int b = blockIdx.y, tid = threadIdx.x;
  const uint32_t bar_a = lrb_smem(bar);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(ROWS) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < Tp; i += ROWS) s_ends[i] = ends[(long long)b * Tp + i];
  for (int i = tid; i < D; i += ROWS) zero_row[i] = 0.f;
  const int mel_len = mel_lens[b];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // zero_row (generic stores) -> bulk-copy engine
  __syncthreads();
  const int total = min(min(mel_len, Tm), s_ends[Tp - 1]);
  const int r_cta = blockIdx.x * ROWS;
  const int r = r_cta + tid;
  const int f = r - out_off;
  const bool exists = r < out_pitch;
  const int idx = exists ? lr_search(s_ends, Tp, f, total) : -1;
  if (frame2ph && exists && f >= 0 && f < Tm) frame2ph[(long long)b * Tm + f] = idx;
  int slot = -1;
  if (idx >= 0) {
    const int start = idx > 0 ? s_ends[idx - 1] : 0;                    // first frame of phoneme idx
    slot = max(start + out_off, r_cta) - r_cta;                         // its first row inside this CTA
  }
  if (slot == tid) {                                                    // leader: fetch the phoneme row
    const float* src = in + ((long long)b * in_pitch + in_off + idx) * D;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(row_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(lrb_smem(slots + (size_t)tid * row_bytes)), "l"(src), "r"(row_bytes), "r"(bar_a)
                 : "memory");
  } else {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a) : "memory");
  }
  if (!exists) return;                                                  // (arrived above; issues no store)
  {
    uint32_t ok = 0, spins = 0;
    while (true) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok) : "r"(bar_a), "r"(0u) : "memory");
      if (ok) break;
      if (++spins > (1u << 22)) __trap();                               // a protocol bug must not hang the GPU
    }
  }
  const uint32_t src_s = idx >= 0 ? lrb_smem(slots + (size_t)slot * row_bytes) : lrb_smem(zero_row);
  float* dst = of + ((long long)b * out_pitch + r) * D;
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_s), "r"(row_bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");        // shared memory must outlive the engine's reads
}

// (The fused form the model uses -- pos-enc add, fp32 + bf16 outputs -- was also tried on this engine: bulk loads of the
// phoneme rows and the pos-enc slice, the add in shared memory, one bulk store per output tile.  Bit-exact, but 34.6 us
// at 8 rows per CTA and 52.1 us at 16 against 27.6 us for lr_expand_kernel at batch 64: the load -> add -> fence ->
// store chain inside one short-lived CTA costs more than the engine saves.  It stays on lr_expand_kernel.)

// backward of the expansion = per-phoneme sums over its frames.  Frame-parallel (a phoneme-parallel kernel waits on
// its longest segment): every warp reads LR_RPW consecutive frame rows, merges neighbours that belong to the same
// phoneme in registers and flushes each run with one 16-byte vector atomic.  dphon must be zero on entry.
template <int RB>
__global__ void __launch_bounds__(THREADS) lr_bwd_kernel(const float* __restrict__ df, const float* __restrict__ df2,
                                                         int f_pitch, int f_off, const int* __restrict__ ends, int B, int Tp,
                                                         int Tm, int D, float* __restrict__ dphon, int p_pitch, int p_off) {
  pdl_wait();
  extern __shared__ int s_ends[];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < Tp; i += THREADS) s_ends[i] = ends[(long long)b * Tp + i];
  __syncthreads();
  const int total = min(s_ends[Tp - 1], Tm);
  const int fw = blockIdx.x * LR_ROWS + warp * LR_RPW;
  if (fw >= total) return;
  int my_idx = -1;
  if (lane < LR_RPW) my_idx = lr_search(s_ends, Tp, fw + lane, total);
  float* dst = dphon + ((long long)b * p_pitch + p_off) * D;
#pragma unroll
  for (int j0 = 0; j0 < LR_RPW; j0 += RB) {
    int idx[RB];
#pragma unroll
    for (int j = 0; j < RB; ++j) idx[j] = __shfl_sync(0xffffffffu, my_idx, j0 + j);
    if (idx[0] < 0) break;                                       // frames are valid up to `total`, then never again
    const float* src = df + ((long long)b * f_pitch + f_off + fw + j0) * D;
    const float* src2 = df2 ? df2 + ((long long)b * f_pitch + f_off + fw + j0) * D : nullptr;
    for (int c = lane * 4; c < D; c += 128) {
      float4 v[RB];
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx[j] >= 0) {
          v[j] = ld4(src + (long long)j * D + c);
          if (src2) { const float4 w = ld4(src2 + (long long)j * D + c); v[j].x += w.x; v[j].y += w.y; v[j].z += w.z; v[j].w += w.w; }
        }
      }
      float4 acc = v[0];
      int cur = idx[0];
#pragma unroll
      for (int j = 1; j < RB; ++j) {
        if (idx[j] != cur) {
          if (cur >= 0) atomicAdd(reinterpret_cast<float4*>(dst + (long long)cur * D + c), acc);
          cur = idx[j];
          acc = v[j];
        } else {
          acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w;
        }
      }
      if (cur >= 0) atomicAdd(reinterpret_cast<float4*>(dst + (long long)cur * D + c), acc);
    }
  }
}

// Segment-sum form: one CTA owns 8 consecutive rows of the padded output (= 8 phonemes) and therefore one contiguous
// run of frames.  A warp owns a 128-column block of the rows and every P-th frame of the run (D = 384: 3 column blocks
// x 2 frame phases = 6 warps), so a long phoneme -- a 200-frame silence -- is shared by all warps instead of stalling
// one, every (phoneme, column) sum has exactly one writer per phase, and the phases are combined in a fixed order:
// no atomics, no memset of the output (halo rows and zero-duration phonemes come out as zeros), bit-reproducible.
constexpr int LRB_MAXD = 512;
__global__ void __launch_bounds__(THREADS) lr_bwd_seg_kernel(const float* __restrict__ df, const float* __restrict__ df2,
                                                             int f_pitch, int f_off, const int* __restrict__ ends, int B,
                                                             int Tp, int Tm, int D, float* __restrict__ dphon, int p_pitch,
                                                             int p_off) {
  pdl_wait();
  __shared__ float part[2 * WARPS * LRB_MAXD];     // [phase][local row][D], phases <= 2 * (512 / D) fit for D >= 256 ...
  __shared__ int s_end[WARPS + 1];                 // s_end[i] = first frame of local row i (s_end[8] = end of the run)
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = blockIdx.x * WARPS;             // first row of the padded (p_pitch-row) output of item b
  const int ncb = D / 128;                         // column blocks (host guarantees D % 128 == 0, D <= 512)
  const int P = min(WARPS / ncb, (2 * WARPS * LRB_MAXD) / (WARPS * D));      // frame phases
  if (threadIdx.x <= WARPS) {
    const int p = row0 - p_off + threadIdx.x;      // phoneme that STARTS at this boundary; rows outside [0, Tp) are empty
    const int* e = ends + (long long)b * Tp;
    s_end[threadIdx.x] = p <= 0 ? 0 : min(e[min(p, Tp) - 1], Tm);
  }
  __syncthreads();
  const int fa = s_end[0], fb = s_end[WARPS];
  const int cb = warp % ncb, phase = warp / ncb;
  const int c = cb * 128 + lane * 4;
  if (phase < P) {
    const float* src = df + ((long long)b * f_pitch + f_off) * D + c;
    const float* src2 = df2 ? df2 + ((long long)b * f_pitch + f_off) * D + c : nullptr;
    float* mine = part + (long long)phase * WARPS * D + c;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = 0;                                   // local row of the running sum
    for (int q = 0; q < WARPS; ++q) st4(mine + q * D, acc);
    for (int f = fa + phase; f < fb; f += 4 * P) {
      float4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ff = f + j * P;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ff < fb) {
          v[j] = ld4(src + (long long)ff * D);
          if (src2) { const float4 w = ld4(src2 + (long long)ff * D); v[j].x += w.x; v[j].y += w.y; v[j].z += w.z; v[j].w += w.w; }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ff = f + j * P;
        if (ff >= fb) break;
        int q = cur;
        while (ff >= s_end[q + 1]) ++q;            // frames are visited in increasing order: q only grows
        if (q != cur) { st4(mine + cur * D, acc); acc = make_float4(0.f, 0.f, 0.f, 0.f); cur = q; }
        acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w;
      }
    }
    st4(mine + cur * D, acc);
  }
  __syncthreads();
  // row `warp` of the tile: phases added in order, stored once
  const int row = row0 + warp;
  if (row < p_pitch) {
    float* dst = dphon + ((long long)b * p_pitch + row) * D;
    for (int cc = lane * 4; cc < D; cc += 128) {
      float4 t = ld4(part + warp * D + cc);
      for (int ph = 1; ph < P; ++ph) {
        const float4 u = ld4(part + ((long long)ph * WARPS + warp) * D + cc);
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      st4(dst + cc, t);
    }
  }
}

int g_lr_bwd_seg = 0; // 0: frame-parallel form with vector atomics (default: 23.5 us at B = 64), 1: segment-sum form (no memset, no
                      // atomics, bit-reproducible; 31.5 us) -- fs2_lr_tune_bwd
int g_lr_rb = 4;      // rows in flight per warp in lr_expand / lr_bwd: 2, 4 or 8 (fs2_lr_tune)
int g_lr_bulk = 8;    // rows per CTA of lr_expand_bulk_kernel: 8 ... 128 (8 measured best on B200: 82 % of the HBM peak); 0 = SIMT kernel (fs2_lr_bulk_rows)

// -------------------------------------------------------------------- row-space utilities --
template <typename TA>
__global__ void __launch_bounds__(THREADS) fold_halo_kernel(const float* src, int B, int T, int C, int pw,
                                                            const float* add, const float* add2, const int* lens,
                                                            float* of, TA* oa) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int TP = T + 2 * FS2_PAD;
  const long long rows = (long long)B * TP;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * WARPS) {
    int b = (int)(r / TP), t = (int)(r - (long long)b * TP) - FS2_PAD;
    const long long ro = r * C;
    const bool in = t >= 0 && t < T;
    const bool live = in && (lens == nullptr || t < lens[b]);
    const long long m1 = (in && pw > 0 && t >= 1 && t <= pw) ? -2LL * t * C : 0;
    const long long m2 = (in && pw > 0 && t >= T - 1 - pw && t <= T - 2) ? 2LL * (T - 1 - t) * C : 0;
    for (int c = lane * 4; c < C; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        if (src) {
          v = ld4(src + ro + c);
          if (m1) { float4 a = ld4(src + ro + m1 + c); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
          if (m2) { float4 a = ld4(src + ro + m2 + c); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
        }
        if (add) { float4 a = ld4(add + ro + c); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
        if (add2) { float4 a = ld4(add2 + ro + c); v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
      }
      if (of) st4(of + ro + c, v);
      if (oa) st4(oa + ro + c, v);
    }
  }
}

template <typename TX>
__global__ void colsum_kernel(const TX* x, long long rows, int C, long long ld, float* out) {
  pdl_wait();
  __shared__ float4 red[8][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    // eight independent row loads in flight per thread (the plain loop was latency-bound: 18 % of HBM peak at C = 384).
    // (A bf16 variant with 16-byte loads -- 8 columns per thread, half as many column blocks -- was measured SLOWER:
    //  33 vs 21.7 us at C = 1536, 27 vs 15.7 us at C = 384; rejected.)
    const long long step = (long long)gridDim.y * 8;
    long long r = (long long)blockIdx.y * 8 + threadIdx.y;
    for (; r + 7 * step < rows; r += 8 * step) {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ld4(x + (r + i * step) * ld + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) { a.x += v[i].x; a.y += v[i].y; a.z += v[i].z; a.w += v[i].w; }
    }
#pragma unroll 4
    for (; r < rows; r += step) {
      float4 v = ld4(x + r * ld + c);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    for (int i = 1; i < 8; ++i) { float4 v = red[i][threadIdx.x]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
    atomicAdd(out + c, a.x);
    atomicAdd(out + c + 1, a.y);
    atomicAdd(out + c + 2, a.z);
    atomicAdd(out + c + 3, a.w);
  }
}

template <typename TA>
__global__ void __launch_bounds__(THREADS) unpad_mask_kernel(const float* src, const int* lens, int B, int T, int C,
                                                             float* plain, TA* oa, int halo) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int TP = T + 2 * FS2_PAD;
  const long long rows = (long long)B * TP;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * WARPS) {
    int b = (int)(r / TP), t = (int)(r - (long long)b * TP) - FS2_PAD;
    const long long ro = r * C;
    if (t < 0 || t >= T) {
      if (oa && halo_row_needs_zero(t, T, halo))
        for (int c = lane * 4; c < C; c += 128) st4(oa + ro + c, make_float4(0.f, 0.f, 0.f, 0.f));
      continue;
    }
    const bool live = (lens == nullptr) || (t < lens[b]);
    for (int c = lane * 4; c < C; c += 128) {
      float4 v = live ? ld4(src + ro + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (plain) st4(plain + ((long long)b * T + t) * C + c, v);
      if (oa) store_row4<TA>(nullptr, oa, ro, c, v, t, T, C, halo);
    }
  }
}

template <typename TA>
__global__ void __launch_bounds__(THREADS) pad_rows_kernel(const float* a, const float* a2, int B, int T, int C,
                                                           float scale, float* of, TA* oa) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int TP = T + 2 * FS2_PAD;
  const long long rows = (long long)B * TP;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * WARPS) {
    int b = (int)(r / TP), t = (int)(r - (long long)b * TP) - FS2_PAD;
    const bool in = t >= 0 && t < T;
    for (int c = lane * 4; c < C; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in) {
        long long o = ((long long)b * T + t) * C + c;
        v = ld4(a + o);
        if (a2) { float4 w = ld4(a2 + o); v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
      }
      if (of) st4(of + r * C + c, v);
      if (oa) st4(oa + r * C + c, v);
    }
  }
}

// Weight packing: dst[co, j, ci] = src[co*src_ld + ci*k + j]  (torch Conv1d (Cout,Cin,k) -> tap-major K),
// one launch for the whole parameter set driven by a device-side table.
// Blocks of grid row y walk the output channels of item y; a channel's (Cin, k) slab is read contiguously into
// shared memory and written back tap-major, so both the fp32 reads and the operand-dtype writes are coalesced.
constexpr int PACK_MAX = 4608;      // floats of one output channel's slab staged in smem (18 KB)
template <typename TD>
__global__ void __launch_bounds__(256) pack_weights_kernel(const Fs2PackItem* items, const float* __restrict__ src_base,
                                                           TD* __restrict__ dst_base) {
  pdl_wait();
  __shared__ float slab[PACK_MAX];
  const Fs2PackItem it = items[blockIdx.y];
  const int rk = it.cin * it.k;
  const float* src = src_base + it.src_off;
  TD* dst = dst_base + it.dst_off;
  if (it.k == 1) {
    for (int co = blockIdx.x; co < it.cout; co += gridDim.x) {
      const float* s = src + (long long)co * it.src_ld;
      TD* d = dst + (long long)co * it.cin;
      for (int ci = threadIdx.x; ci < it.cin; ci += 256) ActT<TD>::st(d + ci, s[ci]);
    }
  } else if (rk <= PACK_MAX) {
    for (int co = blockIdx.x; co < it.cout; co += gridDim.x) {
      const float* s = src + (long long)co * it.src_ld;
      TD* d = dst + (long long)co * rk;
      for (int i = threadIdx.x; i < rk; i += 256) slab[i] = s[i];
      __syncthreads();
      for (int i = threadIdx.x; i < rk; i += 256) {
        const int j = i / it.cin, ci = i - j * it.cin;
        ActT<TD>::st(d + i, slab[ci * it.k + j]);
      }
      __syncthreads();
    }
  } else {
    const long long n = (long long)it.cout * rk;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
      int ci = (int)(i % it.cin);
      long long q = i / it.cin;
      int j = (int)(q % it.k);
      long long co = q / it.k;
      ActT<TD>::st(dst + i, src[co * it.src_ld + (long long)ci * it.k + j]);
    }
  }
}

__global__ void counter_add_kernel(unsigned long long* ctr, unsigned long long inc) {
  pdl_wait(); *ctr += inc; }

__global__ void cast_bf16_kernel(const float* s, bf16* d, long long n) {
  pdl_wait();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) st4(d + i, ld4(s + i));
  else for (; i < n; ++i) d[i] = __float2bfloat16_rn(s[i]);
}
__global__ void add_kernel(float* d, const float* s, long long n) {
  pdl_wait();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    float4 a = ld4(d + i), b = ld4(s + i);
    st4(d + i, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
  } else for (; i < n; ++i) d[i] += s[i];
}

// AdamW over a slice [base, base + n) of the flat parameter buffer.  In the same pass (the updated value is in
// registers anyway) it refreshes the bf16 operand MIRROR of the flat buffer (params.py: the tensor-core GEMMs read the
// mirror, so no separate weight-packing pass runs per step) and the one gathered operand (a column block of a wider
// matrix, copied to the mirror's tail).  `err`: the tcgen05 kernels' mbarrier-timeout word -- when it is set the step's
// gradients are not trustworthy and the update is skipped (the host raises at its next check, optim.py).
struct AdamMirror {
  bf16* mirror;           // nullptr: no mirror
  long long base;         // flat index of p[0]
  long long g_src_off, g_src_ld, g_dst_off;
  int g_rows, g_cols;     // g_rows == 0: no gathered operand
};
__global__ void adamw_kernel(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                             float eps, float wd, float bc1, float bc2_sqrt, float gscale, AdamMirror mr, const int* err) {
  pdl_wait();
  if (err != nullptr && *err != 0) return;
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float out[4];
  int cnt;
  if (i + 3 < n) {
    float4 P = ld4(p + i), G = ld4(g + i), M = ld4(m + i), V = ld4(v + i);
    float* pp = &P.x; float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float gr = gg[q] * gscale;
      pp[q] *= (1.f - lr * wd);
      mm[q] = b1 * mm[q] + (1.f - b1) * gr;
      vv[q] = b2 * vv[q] + (1.f - b2) * gr * gr;
      float denom = sqrtf(vv[q]) / bc2_sqrt + eps;
      pp[q] -= (lr / bc1) * (mm[q] / denom);
      out[q] = pp[q];
    }
    st4(p + i, P); st4(m + i, M); st4(v + i, V);
    if (mr.mirror) st4(mr.mirror + mr.base + i, P);
    cnt = 4;
  } else {
    cnt = (int)(n - i);
    for (int q = 0; q < cnt; ++q) {
      float gr = g[i + q] * gscale;
      float P = p[i + q] * (1.f - lr * wd);
      float M = b1 * m[i + q] + (1.f - b1) * gr;
      float V = b2 * v[i + q] + (1.f - b2) * gr * gr;
      P -= (lr / bc1) * (M / (sqrtf(V) / bc2_sqrt + eps));
      p[i + q] = P; m[i + q] = M; v[i + q] = V;
      out[q] = P;
      if (mr.mirror) mr.mirror[mr.base + i + q] = __float2bfloat16_rn(P);
    }
  }
  if (mr.mirror && mr.g_rows > 0) {
    const long long f0 = mr.base + i - mr.g_src_off;            // position inside the source matrix
    if (f0 + cnt > 0 && f0 < (long long)mr.g_rows * mr.g_src_ld) {
      for (int q = 0; q < cnt; ++q) {
        const long long f = f0 + q;
        if (f < 0 || f >= (long long)mr.g_rows * mr.g_src_ld) continue;
        const long long r = f / mr.g_src_ld;
        const int c = (int)(f - r * mr.g_src_ld);
        if (c < mr.g_cols) mr.mirror[mr.g_dst_off + r * mr.g_cols + c] = __float2bfloat16_rn(out[q]);
      }
    }
  }
}

// train.py:16-51: per-phoneme mean of frame intensities (plain mean, zero-duration phonemes -> 0)
__global__ void intensity_segment_mean_kernel(const float* I, const int64_t* dur, const int64_t* phon_len, int B, int Tp,
                                              int Tm, int D, float* out) {
  pdl_wait();
  extern __shared__ int ends_s[];
  const int b = blockIdx.x;
  const int pl = (int)phon_len[b];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int p = 0; p < Tp; ++p) { if (p < pl) acc += (int)dur[(long long)b * Tp + p]; ends_s[p] = acc; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Tp * D; i += blockDim.x) {
    int p = i / D, d = i - p * D;
    float r = 0.f;
    if (p < pl) {
      int e = min(ends_s[p], Tm), s = p ? min(ends_s[p - 1], Tm) : 0;
      float acc = 0.f;
      for (int f = s; f < e; ++f) acc += I[((long long)b * Tm + f) * D + d];
      float den = fmaxf((float)dur[(long long)b * Tp + p], 1.0f);
      r = acc / den;
    }
    out[((long long)b * Tp + p) * D + d] = r;
  }
}

}  // namespace

// ------------------------------------------------------------ intensity extractor glue --
template <typename TA>
__global__ void __launch_bounds__(THREADS) frames_to_rows_kernel(const float* __restrict__ x, int channels_first, int B,
                                                                 int C, int T, int Cpad, TA* __restrict__ out) {
  pdl_wait();
  const int TP = T + 2 * FS2_PAD;
  const long long n = (long long)B * TP * Cpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long long r = i / Cpad;
    const int b = (int)(r / TP), t = (int)(r - (long long)b * TP) - FS2_PAD;
    float v = 0.f;
    if (t >= 0 && t < T && c < C)
      v = channels_first ? x[((long long)b * C + c) * T + t] : x[((long long)b * T + t) * C + c];
    ActT<TA>::st(out + i, v);
  }
}

// one warp per (b, t): n_out (<= 8) dot products over D of the masked, emotion-shifted hidden row
__global__ void __launch_bounds__(THREADS) intensity_head_kernel(const float* __restrict__ h, const float* __restrict__ emb,
                                                                 const int64_t* __restrict__ emotions,
                                                                 const int* __restrict__ lens, const float* __restrict__ Wc,
                                                                 const float* __restrict__ bc, int B, int T, int D,
                                                                 int n_out, float* __restrict__ out) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int TP = T + 2 * FS2_PAD;
  const long long rows = (long long)B * T;
  for (long long r = (long long)blockIdx.x * WARPS + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * WARPS) {
    const int b = (int)(r / T), t = (int)(r - (long long)b * T);
    const bool live = t < lens[b];
    const float* hr = h + ((long long)b * TP + FS2_PAD + t) * D;
    const float* er = emb + emotions[b] * D;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (live) {
      for (int c = lane; c < D; c += 32) {
        const float v = hr[c] + er[c];
#pragma unroll
        for (int o = 0; o < 8; ++o)
          if (o < n_out) acc[o] += v * Wc[(long long)o * D + c];
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      if (o < n_out) {
        const float s = warp_sum(acc[o]);
        if (lane == 0) out[r * n_out + o] = s + bc[o];
      }
    }
  }
}

// exact (erf) GELU in place over an operand buffer: nn.GELU() of rank_model/model.py:31, 42
template <typename TA>
__global__ void __launch_bounds__(256) gelu_kernel(TA* x, long long n) {
  pdl_wait();
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (; i + 3 < n; i += stride) {
    float4 v = ld4(x + i);
    v.x = 0.5f * v.x * (1.0f + erff(v.x * 0.70710678118654752f));
    v.y = 0.5f * v.y * (1.0f + erff(v.y * 0.70710678118654752f));
    v.z = 0.5f * v.z * (1.0f + erff(v.z * 0.70710678118654752f));
    v.w = 0.5f * v.w * (1.0f + erff(v.w * 0.70710678118654752f));
    st4(x + i, v);
  }
}

// ---------------------------------------------------------------------- device collate --
constexpr int COL_T = 32;      // frames per CTA
// grid (ceil(Tm / 32) + 1, B): x < ceil(Tm/32) handles a 32-frame slab of utterance row i (all n_mels + 2 channels go
// through shared memory so that the channels-first reads, the channels-first rank_X writes and the frames-first mel
// writes are all coalesced); the last x handles the phoneme / duration row.
__global__ void __launch_bounds__(256) collate_kernel(const int64_t* __restrict__ phon_cat, const int64_t* __restrict__ dur_cat,
                                                      const float* __restrict__ mel_cat, const float* __restrict__ pitch_cat,
                                                      const float* __restrict__ energy_cat, const int* __restrict__ ph_start,
                                                      const int* __restrict__ ph_len, const int* __restrict__ fr_start,
                                                      const int* __restrict__ fr_len, int B, int Tp, int Tm, int n_mels,
                                                      int64_t* __restrict__ phoneme, int64_t* __restrict__ duration,
                                                      float* __restrict__ mel, float* __restrict__ pitch,
                                                      float* __restrict__ energy, float* __restrict__ rank_X) {
  pdl_wait();
  extern __shared__ float tile[];                 // [(n_mels + 2)][COL_T + 1]
  const int i = blockIdx.y;
  const int nslab = (Tm + COL_T - 1) / COL_T;
  if ((int)blockIdx.x == nslab) {
    const int n = ph_len[i];
    const long long s0 = ph_start[i];
    for (int p = threadIdx.x; p < Tp; p += blockDim.x) {
      phoneme[(long long)i * Tp + p] = p < n ? phon_cat[s0 + p] : 0;
      duration[(long long)i * Tp + p] = p < n ? dur_cat[s0 + p] : 0;
    }
    return;
  }
  const int C = n_mels + 2;
  const int t0 = blockIdx.x * COL_T;
  const int L = fr_len[i];
  const long long f0 = fr_start[i];
  for (int e = threadIdx.x; e < C * COL_T; e += blockDim.x) {
    const int c = e / COL_T, tt = e - c * COL_T;
    const int t = t0 + tt;
    float v = 0.f;
    if (t < L) {
      if (c < n_mels) v = mel_cat[f0 * n_mels + (long long)c * L + t];
      else if (c == n_mels) v = pitch_cat[f0 + t];
      else v = energy_cat[f0 + t];
    }
    tile[c * (COL_T + 1) + tt] = v;
    if (t < Tm) {
      rank_X[((long long)i * C + c) * Tm + t] = v;
      if (c == n_mels) pitch[(long long)i * Tm + t] = v;
      if (c == n_mels + 1) energy[(long long)i * Tm + t] = v;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < COL_T * n_mels; e += blockDim.x) {
    const int tt = e / n_mels, c = e - tt * n_mels;
    const int t = t0 + tt;
    if (t < Tm) mel[((long long)i * Tm + t) * n_mels + c] = tile[c * (COL_T + 1) + tt];
  }
}

// ------------------------------------------------------------------ prototype buckets --
__device__ __forceinline__ long long bucket_of(long long g, long long n, int k) {
  const long long q = n / k, rem = n % k;          // np.array_split: the first `rem` buckets hold q + 1 elements
  const long long edge = rem * (q + 1);
  return g < edge ? g / (q + 1) : rem + (g - edge) / q;
}
__global__ void __launch_bounds__(256) prototype_accum_kernel(const float* __restrict__ I, const int* __restrict__ lens,
                                                              const int* __restrict__ group,
                                                              const long long* __restrict__ frame_off,
                                                              const long long* __restrict__ group_total, int Tmax, int D,
                                                              int n_buckets, float* __restrict__ sums) {
  pdl_wait();
  const int i = blockIdx.x;
  const int L = lens[i], grp = group[i];
  const long long n = group_total[grp], f0 = frame_off[i];
  for (int e = threadIdx.x; e < L * D; e += blockDim.x) {
    const int t = e / D, d = e - t * D;
    const long long b = bucket_of(f0 + t, n, n_buckets);
    atomicAdd(sums + ((long long)grp * n_buckets + b) * D + d, I[((long long)i * Tmax + t) * D + d]);
  }
}
__global__ void prototype_finalize_kernel(const long long* __restrict__ group_total, int D, int n_groups, int n_buckets,
                                          float* __restrict__ sums) {
  pdl_wait();
  const long long total = (long long)n_groups * n_buckets * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long gb = e / D;
    const int grp = (int)(gb / n_buckets), b = (int)(gb - (long long)grp * n_buckets);
    const long long n = group_total[grp];
    if (n == 0) continue;                                   // the reference leaves such groups at zero
    const long long size = n / n_buckets + (b < n % n_buckets ? 1 : 0);
    sums[e] = sums[e] / (float)size;                        // size 0 -> 0/0 = NaN, numpy's mean of an empty slice
  }
}

#define ST ((cudaStream_t)stream)
#define REQUIRE(cond, msg) do { if (!(cond)) { fs2_set_error(msg); return FS2_ERR_ARG; } } while (0)

extern "C" int fs2_embed_posenc(const int64_t* tokens, const float* emb, const float* pe, int B, int Tp, int D,
                                int pad_idx, float* out_f32, void* out_act, int act_bf16, int* src_lens, void* stream) {
  REQUIRE(tokens && emb && pe && src_lens && D % 4 == 0, "fs2_embed_posenc: bad arguments");
  FS2_LAUNCH((count_nonpad_kernel), B, 32, 0, ST, tokens, B, Tp, pad_idx, src_lens);
  int rc = fs2_check_launch();
  if (rc) return rc;
  const long long rows = (long long)B * (Tp + 2 * FS2_PAD);
  if (act_bf16) FS2_LAUNCH((embed_posenc_kernel<bf16>), grid_for_rows(rows), THREADS, 0, ST, tokens, emb, pe, B, Tp, D, pad_idx, out_f32, (bf16*)out_act);
  else FS2_LAUNCH((embed_posenc_kernel<float>), grid_for_rows(rows), THREADS, 0, ST, tokens, emb, pe, B, Tp, D, pad_idx, out_f32, (float*)out_act);
  return fs2_check_launch();
}

extern "C" int fs2_embedding_bwd(const float* dx, const int64_t* tokens, int B, int Tp, int D, int pad_idx, float* demb,
                                 void* stream) {
  REQUIRE(dx && tokens && demb, "fs2_embedding_bwd: null pointer");
  FS2_LAUNCH((embedding_bwd_kernel), grid_for_rows((long long)B * Tp), THREADS, 0, ST, dx, tokens, B, Tp, D, pad_idx, demb);
  return fs2_check_launch();
}

// L2 prefetch of a warp's next row in ln_fwd / ln_bwd: on unless FS2_LN_PREFETCH=0 (whole-step A/B) or fs2_ln_tune(0)
int g_ln_prefetch = [] { const char* e = getenv("FS2_LN_PREFETCH"); return (e && e[0] >= '0' && e[0] <= '4') ? e[0] - '0' : 1; }();

// resident waves of CTAs the LayerNorm grids are capped at.  Measured on B200 (profiles/r01_summary.md section 5): the
// backward is fastest as ONE wave of long-lived CTAs (a warp prefetches 6 of its 7 rows, a third of the partial
// parameter-gradient flushes), the forward as two.  FS2_LN_WAVES=1|2 forces both for A/B runs.
int g_ln_waves_env = [] { const char* e = getenv("FS2_LN_WAVES"); return (e && (e[0] == '1' || e[0] == '2')) ? e[0] - '0' : 0; }();
int g_ln_fwd_waves = g_ln_waves_env ? g_ln_waves_env : 2;
int g_ln_bwd_waves = g_ln_waves_env ? g_ln_waves_env : 1;

/* measurement hook: next-row L2 prefetch in the LayerNorm kernels on (1, default) / off (0) */
extern "C" int fs2_ln_tune(int prefetch) {
  g_ln_prefetch = prefetch < 0 ? 0 : (prefetch > 4 ? 4 : prefetch);   // distance in rows of a warp
  return FS2_OK;
}

extern "C" int fs2_ln_fwd(const Fs2LnFwd* p, void* stream) {
  REQUIRE(p && p->x && p->gamma && p->beta, "fs2_ln_fwd: null pointer");
  REQUIRE(p->C % 4 == 0 && p->C <= 128 * MAXV, "fs2_ln_fwd: C must be a multiple of 4 and <= 512");
  REQUIRE(p->halo <= FS2_PAD && (p->halo == 0 || p->T > p->halo), "fs2_ln_fwd: halo too wide for T");
  const long long rows = (long long)p->B * (p->T + 2 * FS2_PAD);
  int grid = grid_for_rows(rows);
  if (grid > 148 * 4 * g_ln_fwd_waves) grid = 148 * 4 * g_ln_fwd_waves;   // four resident CTAs per SM (<= 64 registers)
  if (p->act_bf16) FS2_LAUNCH((ln_fwd_kernel<bf16>), grid, THREADS, 0, ST, *p, g_ln_prefetch);
  else FS2_LAUNCH((ln_fwd_kernel<float>), grid, THREADS, 0, ST, *p, g_ln_prefetch);
  return fs2_check_launch();
}

extern "C" int fs2_ln_bwd(const Fs2LnBwd* p, void* stream) {
  REQUIRE(p && (p->x || p->y) && p->gamma && p->beta && (p->mean || p->y) && p->rstd, "fs2_ln_bwd: null pointer");
  REQUIRE(p->y == nullptr || (p->head_w == nullptr && p->dhead_w == nullptr && p->dhead_b == nullptr && !p->tanh_act && !p->relu_x &&
                              p->drop_a_p <= 0.f && p->lens == nullptr),
          "fs2_ln_bwd: y (x_hat from the forward output) needs a plain LayerNorm");
  REQUIRE(p->C % 4 == 0 && p->C <= 128 * MAXV, "fs2_ln_bwd: C must be a multiple of 4 and <= 512");
  REQUIRE(p->dact_colsum == nullptr || (p->C <= 384 && p->dact != nullptr), "fs2_ln_bwd: dact_colsum needs dact and C <= 384");
  const long long rows = (long long)p->B * (p->T + 2 * FS2_PAD);
  const int nv = (p->C + 127) / 128;
  // "head" selects the general instantiation: the 384 -> 1 head, tanh, ReLU-of-x or dropout after the norm
  const bool head = p->head_w != nullptr || p->dhead_w != nullptr || p->dhead_b != nullptr || p->tanh_act || p->relu_x ||
                    p->drop_a_p > 0.f;
  int grid = grid_for_rows(rows);
  // FFT-block instantiation: waves of three resident CTAs per SM; the general one (two per SM) keeps its 148 * 6 cap
  const int cap = (nv <= 3 && !head) ? 148 * 3 * g_ln_bwd_waves : 148 * 6;
  if (grid > cap) grid = cap;
#define LN_BWD_LAUNCH(TA, NV, HEAD) FS2_LAUNCH((ln_bwd_kernel<TA, NV, HEAD>), grid, THREADS, 0, ST, *p, g_ln_prefetch)
#define LN_BWD_NV(TA, NV) do { if (head) LN_BWD_LAUNCH(TA, NV, true); else LN_BWD_LAUNCH(TA, NV, false); } while (0)
#define LN_BWD_TA(TA) do { if (nv == 1) LN_BWD_NV(TA, 1); else if (nv == 2) LN_BWD_NV(TA, 2); else if (nv == 3) LN_BWD_NV(TA, 3); else LN_BWD_NV(TA, 4); } while (0)
  if (p->act_bf16) LN_BWD_TA(bf16);
  else LN_BWD_TA(float);
#undef LN_BWD_TA
#undef LN_BWD_NV
#undef LN_BWD_LAUNCH
  return fs2_check_launch();
}

extern "C" int fs2_softmax_fwd(const float* S, const int* lens, int B, int H, int T, int ldk, float scale, float drop_p,
                               unsigned long long seed, const unsigned long long* seed_dev, void* P, void* Pd,
                               int act_bf16, void* stream) {
  REQUIRE(S && lens && P && ldk % 4 == 0 && ldk >= T, "fs2_softmax_fwd: bad arguments");
  DropCfg dc{drop_p, seed};
  if (drop_p <= 0.f) Pd = nullptr;
  const long long rows = (long long)B * H * T;
  if (act_bf16) FS2_LAUNCH((softmax_fwd_kernel<bf16>), grid_for_rows(rows), THREADS, 0, ST, S, lens, B, H, T, ldk, scale, dc, seed_dev, (bf16*)P, (bf16*)Pd);
  else FS2_LAUNCH((softmax_fwd_kernel<float>), grid_for_rows(rows), THREADS, 0, ST, S, lens, B, H, T, ldk, scale, dc, seed_dev, (float*)P, (float*)Pd);
  return fs2_check_launch();
}

extern "C" int fs2_softmax_bwd(const void* P, const float* dPd, const int* lens, int B, int H, int T, int ldk,
                               float scale, float drop_p, unsigned long long seed, const unsigned long long* seed_dev,
                               void* dS, int act_bf16, void* stream) {
  REQUIRE(P && dPd && lens && dS && ldk % 4 == 0, "fs2_softmax_bwd: bad arguments");
  DropCfg dc{drop_p, seed};
  const long long rows = (long long)B * H * T;
  if (act_bf16) FS2_LAUNCH((softmax_bwd_kernel<bf16>), grid_for_rows(rows), THREADS, 0, ST, (const bf16*)P, dPd, lens, B, H, T, ldk, scale, dc, seed_dev, (bf16*)dS);
  else FS2_LAUNCH((softmax_bwd_kernel<float>), grid_for_rows(rows), THREADS, 0, ST, (const float*)P, dPd, lens, B, H, T, ldk, scale, dc, seed_dev, (float*)dS);
  return fs2_check_launch();
}

extern "C" int fs2_cond_finish(const float* G, const float* Wcat, const float* spk_emb, const int64_t* speakers,
                               const float* intensity, const int* lens, int B, int Tp, int D, float* sp_ws, float* y_f32,
                               void* y_act, int act_bf16, int halo, void* stream) {
  REQUIRE(G && Wcat && spk_emb && speakers && intensity && lens && sp_ws && D % 4 == 0, "fs2_cond_finish: bad arguments");
  FS2_LAUNCH((spk_proj_kernel), grid_for_rows((long long)B * D), THREADS, 0, ST, Wcat, spk_emb, speakers, B, D, sp_ws);
  int rc = fs2_check_launch();
  if (rc) return rc;
  const long long rows = (long long)B * (Tp + 2 * FS2_PAD);
  if (act_bf16) FS2_LAUNCH((cond_finish_kernel<bf16>), grid_for_rows(rows), THREADS, 0, ST, G, Wcat, sp_ws, intensity, lens, B, Tp, D, y_f32, (bf16*)y_act, halo);
  else FS2_LAUNCH((cond_finish_kernel<float>), grid_for_rows(rows), THREADS, 0, ST, G, Wcat, sp_ws, intensity, lens, B, Tp, D, y_f32, (float*)y_act, halo);
  return fs2_check_launch();
}

extern "C" int fs2_cond_bwd(const float* dy, const float* Wcat, const float* spk_emb, const int64_t* speakers,
                            const float* intensity, int B, int Tp, int D, float* dsum_ws, float* dWcat, float* dspk_emb,
                            void* stream) {
  REQUIRE(dy && Wcat && spk_emb && speakers && intensity && dsum_ws && dWcat && dspk_emb, "fs2_cond_bwd: null pointer");
  REQUIRE(B > 0 && B <= 65535, "fs2_cond_bwd: B <= 65535");
  CUDA_CHECK_RET(cudaMemsetAsync(dsum_ws, 0, (size_t)B * D * sizeof(float), ST));
  FS2_LAUNCH((cond_bwd_rows_kernel), dim3((Tp + COND_ROWS - 1) / COND_ROWS, B), 128, 0, ST, dy, intensity, B, Tp, D, dsum_ws, dWcat);
  int rc = fs2_check_launch();
  if (rc) return rc;
  FS2_LAUNCH((cond_bwd_spk_kernel), 148 * 4, 256, 0, ST, dsum_ws, Wcat, spk_emb, speakers, B, D, dWcat, dspk_emb);
  return fs2_check_launch();
}

extern "C" int fs2_avg_over_durations(const float* values, const int64_t* durs, int B, int Tp, int Tm, float* avg,
                                      int* starts, int* ends, int* nz, void* stream) {
  REQUIRE(values && durs && avg, "fs2_avg_over_durations: null pointer");
  size_t smem = (size_t)(Tm + 1) * 8 + (size_t)Tp * 4;
  REQUIRE(smem <= 48 * 1024, "fs2_avg_over_durations: Tm too large for the shared-memory scan");
  FS2_LAUNCH((avg_over_durations_kernel), B, 128, smem, ST, values, durs, B, Tp, Tm, avg, starts, ends, nz);
  return fs2_check_launch();
}

extern "C" int fs2_embed_add(const float* x, const float* contour, const float* w, const float* bias, int ksize,
                             const int* lens, int B, int Tp, int D, float* y_f32, void* y_act, int act_bf16, int halo,
                             void* stream) {
  REQUIRE(x && contour && w && bias && lens && ksize >= 1 && ksize <= 9 && (ksize & 1) && Tp > (ksize - 1) / 2,
          "fs2_embed_add: bad arguments");
  const long long rows = (long long)B * (Tp + 2 * FS2_PAD);
  if (act_bf16) FS2_LAUNCH((embed_add_kernel<bf16>), grid_for_rows(rows), THREADS, 0, ST, x, contour, w, bias, ksize, lens, B, Tp, D, y_f32, (bf16*)y_act, halo);
  else FS2_LAUNCH((embed_add_kernel<float>), grid_for_rows(rows), THREADS, 0, ST, x, contour, w, bias, ksize, lens, B, Tp, D, y_f32, (float*)y_act, halo);
  return fs2_check_launch();
}

extern "C" int fs2_embed_add_bwd(const float* dy, const float* contour, int ksize, int B, int Tp, int D, float* dw,
                                 float* dbias, void* stream) {
  REQUIRE(dy && contour && dw && dbias && ksize <= 9, "fs2_embed_add_bwd: bad arguments");
  dim3 grid((D + 127) / 128, 256);
  FS2_LAUNCH((embed_add_bwd_kernel), grid, 128, 0, ST, dy, contour, ksize, B, Tp, D, dw, dbias);
  return fs2_check_launch();
}

extern "C" int fs2_dur_decode(const float* log_dur, long long n, float* fdur, void* stream) {
  REQUIRE(log_dur && fdur, "fs2_dur_decode: null pointer");
  FS2_LAUNCH((dur_decode_kernel), (unsigned)((n + 255) / 256), 256, 0, ST, log_dur, n, fdur);
  return fs2_check_launch();
}

extern "C" int fs2_lr_prepare(const int64_t* dur, const float* fdur, float pace, int B, int Tp, int* ends, int* mel_lens,
                              void* stream) {
  REQUIRE((dur || fdur) && ends && mel_lens, "fs2_lr_prepare: null pointer");
  FS2_LAUNCH((lr_prepare_kernel), B, 32, 0, ST, dur, fdur, pace, B, Tp, ends, mel_lens);
  return fs2_check_launch();
}

extern "C" int fs2_lr_finalize(const int* mel_lens, int B, int Tm_expected, long long* out_i64, int* flag, void* stream) {
  REQUIRE(mel_lens && out_i64 && flag, "fs2_lr_finalize: null pointer");
  FS2_LAUNCH((lr_finalize_kernel), 1, 32, 0, ST, mel_lens, B, Tm_expected, out_i64, flag);
  return fs2_check_launch();
}

extern "C" int fs2_lr_expand(const float* in, int in_pitch, int in_off, const int* ends, const int* mel_lens,
                             const float* pe, int B, int Tp, int Tm, int D, float* out_f32, void* out_act, int act_bf16,
                             int out_pitch, int out_off, int* frame2ph, void* stream) {
  REQUIRE(in && ends && mel_lens && D % 4 == 0, "fs2_lr_expand: bad arguments");
  REQUIRE(B > 0 && B <= 65535 && Tp > 0 && Tp <= 12000 && out_pitch > 0, "fs2_lr_expand: B <= 65535 and 0 < Tp <= 12000 (prefix sums are staged in shared memory)");
  // plain fp32 copy form -> bulk-copy engine kernel (rows and bases must be 16-byte aligned, slots must fit in smem)
  if (g_lr_bulk && !pe && !out_act && out_f32 && (((uintptr_t)in | (uintptr_t)out_f32) & 15) == 0) {
    const size_t smb = 16 + (size_t)(g_lr_bulk + 1) * D * 4 + (size_t)Tp * sizeof(int);
    if (smb <= 200 * 1024) {
      const dim3 gridb((out_pitch + g_lr_bulk - 1) / g_lr_bulk, B);
#define LR_BULK(R)                                                                                                       \
  do {                                                                                                                   \
    static bool smem_opt_in = false;                                                                                     \
    if (!smem_opt_in) {                                                                                                  \
      CUDA_CHECK_RET(cudaFuncSetAttribute(lr_expand_bulk_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      smem_opt_in = true;                                                                                                \
    }                                                                                                                    \
    FS2_LAUNCH((lr_expand_bulk_kernel<R>), gridb, R, smb, ST, in, in_pitch, in_off, ends, mel_lens, B, Tp, Tm, D, out_f32, \
               out_pitch, out_off, frame2ph);                                                                            \
  } while (0)
      if (g_lr_bulk == 8) LR_BULK(8); else if (g_lr_bulk == 16) LR_BULK(16); else if (g_lr_bulk == 32) LR_BULK(32); else if (g_lr_bulk == 128) LR_BULK(128); else if (g_lr_bulk == 64) LR_BULK(64); else LR_BULK(8);
#undef LR_BULK
      return fs2_check_launch();
    }
  }
  const dim3 grid((out_pitch + LR_ROWS - 1) / LR_ROWS, B);
  const size_t sm = (size_t)Tp * sizeof(int);
#define LR_EXP(TA, RB) FS2_LAUNCH((lr_expand_kernel<TA, RB>), grid, THREADS, sm, ST, in, in_pitch, in_off, ends, mel_lens, pe, B, Tp, Tm, D, out_f32, (TA*)out_act, out_pitch, out_off, frame2ph)
  if (act_bf16) { if (g_lr_rb == 2) LR_EXP(bf16, 2); else if (g_lr_rb == 8) LR_EXP(bf16, 8); else LR_EXP(bf16, 4); }
  else { if (g_lr_rb == 2) LR_EXP(float, 2); else if (g_lr_rb == 8) LR_EXP(float, 8); else LR_EXP(float, 4); }
#undef LR_EXP
  return fs2_check_launch();
}

extern "C" int fs2_lr_bwd(const float* dframes, const float* dframes2, int f_pitch, int f_off, const int* ends,
                          const int* mel_lens, int B, int Tp, int Tm, int D, float* dphon, int p_pitch, int p_off,
                          void* stream) {
  REQUIRE(dframes && ends && dphon && D % 4 == 0, "fs2_lr_bwd: bad arguments");
  REQUIRE(B > 0 && B <= 65535 && Tp > 0 && Tp <= 12000 && Tm > 0, "fs2_lr_bwd: B <= 65535 and 0 < Tp <= 12000 (prefix sums are staged in shared memory)");
  (void)mel_lens;
  if (g_lr_bwd_seg && D % 128 == 0 && D <= LRB_MAXD) {
    REQUIRE(p_off >= 0 && p_pitch >= p_off + Tp, "fs2_lr_bwd: bad output pitch");
    FS2_LAUNCH((lr_bwd_seg_kernel), dim3((p_pitch + WARPS - 1) / WARPS, B), THREADS, 0, ST, dframes, dframes2, f_pitch, f_off, ends,
               B, Tp, Tm, D, dphon, p_pitch, p_off);
    return fs2_check_launch();
  }
  CUDA_CHECK_RET(cudaMemsetAsync(dphon, 0, (size_t)B * p_pitch * D * sizeof(float), ST));
  const dim3 grid((Tm + LR_ROWS - 1) / LR_ROWS, B);
#define LR_BWD(RB) FS2_LAUNCH((lr_bwd_kernel<RB>), grid, THREADS, (size_t)Tp * sizeof(int), ST, dframes, dframes2, f_pitch, f_off, ends, B, Tp, Tm, D, dphon, p_pitch, p_off)
  if (g_lr_rb == 2) LR_BWD(2); else if (g_lr_rb == 8) LR_BWD(8); else LR_BWD(4);
#undef LR_BWD
  return fs2_check_launch();
}

/* measurement hook: rows in flight per warp (2, 4 or 8) of the LengthRegulator kernels */
extern "C" int fs2_lr_tune(int rows_in_flight) {
  REQUIRE(rows_in_flight == 2 || rows_in_flight == 4 || rows_in_flight == 8, "fs2_lr_tune: 2, 4 or 8");
  g_lr_rb = rows_in_flight;
  return FS2_OK;
}

/* 1 = segment-sum form of fs2_lr_bwd (plain stores, no memset, bit-reproducible), 0 (default) = frame-parallel form
 * (memset + 16-byte vector atomics; faster on B200: 23.5 vs 31.5 us at B = 64) */
extern "C" int fs2_lr_tune_bwd(int segment_sum) {
  g_lr_bwd_seg = segment_sum ? 1 : 0;
  return FS2_OK;
}

/* measurement hook: rows per CTA (8 ... 128, default 8) of the bulk-copy-engine form of fs2_lr_expand, 0 = always the SIMT kernel */
extern "C" int fs2_lr_bulk_rows(int rows) {
  REQUIRE(rows == 0 || rows == 8 || rows == 16 || rows == 32 || rows == 64 || rows == 128, "fs2_lr_bulk_rows: 0, 8, 16, 32, 64 or 128");
  g_lr_bulk = rows;
  return FS2_OK;
}

extern "C" int fs2_fold_halo(const float* src, int B, int T, int C, int p, const float* add, const float* add2,
                             const int* lens, float* out_f32, void* out_act, int act_bf16, void* stream) {
  REQUIRE(C % 4 == 0 && p <= FS2_PAD, "fs2_fold_halo: bad arguments");
  const long long rows = (long long)B * (T + 2 * FS2_PAD);
  if (act_bf16) FS2_LAUNCH((fold_halo_kernel<bf16>), grid_for_rows(rows), THREADS, 0, ST, src, B, T, C, p, add, add2, lens, out_f32, (bf16*)out_act);
  else FS2_LAUNCH((fold_halo_kernel<float>), grid_for_rows(rows), THREADS, 0, ST, src, B, T, C, p, add, add2, lens, out_f32, (float*)out_act);
  return fs2_check_launch();
}

extern "C" int fs2_colsum(const void* x, int x_bf16, long long rows, int C, long long ld, float* out, void* stream) {
  REQUIRE(x && out && C % 4 == 0 && ld % 4 == 0, "fs2_colsum: bad arguments");
  long long gy = (rows + 63) / 64;
  if (gy > 296) gy = 296;
  if (gy < 1) gy = 1;
  dim3 grid((C + 127) / 128, (unsigned)gy), block(32, 8);
  if (x_bf16) FS2_LAUNCH((colsum_kernel<bf16>), grid, block, 0, ST, (const bf16*)x, rows, C, ld, out);
  else FS2_LAUNCH((colsum_kernel<float>), grid, block, 0, ST, (const float*)x, rows, C, ld, out);
  return fs2_check_launch();
}

extern "C" int fs2_unpad_mask(const float* src, const int* lens, int B, int T, int C, float* out_plain, void* out_act,
                              int act_bf16, int halo, void* stream) {
  REQUIRE(src && C % 4 == 0, "fs2_unpad_mask: bad arguments");
  const long long rows = (long long)B * (T + 2 * FS2_PAD);
  if (act_bf16) FS2_LAUNCH((unpad_mask_kernel<bf16>), grid_for_rows(rows), THREADS, 0, ST, src, lens, B, T, C, out_plain, (bf16*)out_act, halo);
  else FS2_LAUNCH((unpad_mask_kernel<float>), grid_for_rows(rows), THREADS, 0, ST, src, lens, B, T, C, out_plain, (float*)out_act, halo);
  return fs2_check_launch();
}

extern "C" int fs2_pad_rows(const float* src_plain, const float* src2_plain, int B, int T, int C, float scale,
                            float* out_f32, void* out_act, int act_bf16, void* stream) {
  REQUIRE(src_plain && C % 4 == 0, "fs2_pad_rows: bad arguments");
  const long long rows = (long long)B * (T + 2 * FS2_PAD);
  if (act_bf16) FS2_LAUNCH((pad_rows_kernel<bf16>), grid_for_rows(rows), THREADS, 0, ST, src_plain, src2_plain, B, T, C, scale, out_f32, (bf16*)out_act);
  else FS2_LAUNCH((pad_rows_kernel<float>), grid_for_rows(rows), THREADS, 0, ST, src_plain, src2_plain, B, T, C, scale, out_f32, (float*)out_act);
  return fs2_check_launch();
}

extern "C" int fs2_pack_weights(const Fs2PackItem* items_dev, int n_items, const float* src_base, void* dst_base,
                                int dst_bf16, void* stream) {
  REQUIRE(items_dev && src_base && dst_base && n_items > 0, "fs2_pack_weights: bad arguments");
  dim3 grid(128, n_items);
  if (dst_bf16) FS2_LAUNCH((pack_weights_kernel<bf16>), grid, 256, 0, ST, items_dev, src_base, (bf16*)dst_base);
  else FS2_LAUNCH((pack_weights_kernel<float>), grid, 256, 0, ST, items_dev, src_base, (float*)dst_base);
  return fs2_check_launch();
}

extern "C" int fs2_cast_bf16(const float* src, void* dst, long long n, void* stream) {
  REQUIRE(src && dst, "fs2_cast_bf16: null pointer");
  FS2_LAUNCH((cast_bf16_kernel), (unsigned)((n / 4 + 256) / 256), 256, 0, ST, src, (bf16*)dst, n);
  return fs2_check_launch();
}

extern "C" int fs2_add_(float* dst, const float* src, long long n, void* stream) {
  REQUIRE(src && dst, "fs2_add_: null pointer");
  FS2_LAUNCH((add_kernel), (unsigned)((n / 4 + 256) / 256), 256, 0, ST, dst, src, n);
  return fs2_check_launch();
}

extern "C" int fs2_counter_add(unsigned long long* ctr, unsigned long long inc, void* stream) {
  REQUIRE(ctr, "fs2_counter_add: null pointer");
  FS2_LAUNCH((counter_add_kernel), 1, 1, 0, ST, ctr, inc);
  return fs2_check_launch();
}

extern "C" int fs2_memset(void* dst, int value, long long nbytes, void* stream) {
  CUDA_CHECK_RET(cudaMemsetAsync(dst, value, (size_t)nbytes, ST));
  return FS2_OK;
}

int fs2_tc_error_ptr(int** out);      // gemm_tc.cu

extern "C" int fs2_adamw_fused(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                               float beta2, float eps, float wd, int step, float grad_scale, void* mirror_bf16,
                               long long base, long long g_src_off, long long g_src_ld, int g_rows, int g_cols,
                               long long g_dst_off, int guard_tc_error, void* stream) {
  REQUIRE(p && g && m && v && step >= 1, "fs2_adamw: bad arguments");
  REQUIRE(mirror_bf16 == nullptr || base % 4 == 0, "fs2_adamw_fused: slice start must be a multiple of 4 elements");
  float bc1 = 1.0f - powf(beta1, (float)step);
  float bc2 = sqrtf(1.0f - powf(beta2, (float)step));
  AdamMirror mr;
  mr.mirror = (bf16*)mirror_bf16;
  mr.base = base;
  mr.g_src_off = g_src_off; mr.g_src_ld = g_src_ld > 0 ? g_src_ld : 1; mr.g_dst_off = g_dst_off;
  mr.g_rows = g_rows; mr.g_cols = g_cols;
  int* err = nullptr;
  if (guard_tc_error) {
    int rc = fs2_tc_error_ptr(&err);
    if (rc) return rc;
  }
  FS2_LAUNCH((adamw_kernel), (unsigned)((n / 4 + 256) / 256), 256, 0, ST, p, g, m, v, n, lr, beta1, beta2, eps, wd, bc1, bc2,
             grad_scale, mr, (const int*)err);
  return fs2_check_launch();
}

extern "C" int fs2_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                         float eps, float wd, int step, float grad_scale, void* stream) {
  return fs2_adamw_fused(p, g, m, v, n, lr, beta1, beta2, eps, wd, step, grad_scale, nullptr, 0, 0, 1, 0, 0, 0, 0, stream);
}

extern "C" int fs2_intensity_segment_mean(const float* I, const int64_t* dur, const int64_t* phon_len, int B, int Tp,
                                          int Tm, int D, float* out, void* stream) {
  REQUIRE(I && dur && phon_len && out, "fs2_intensity_segment_mean: null pointer");
  FS2_LAUNCH((intensity_segment_mean_kernel), B, 256, (size_t)Tp * 4, ST, I, dur, phon_len, B, Tp, Tm, D, out);
  return fs2_check_launch();
}

extern "C" int fs2_frames_to_rows(const float* x, int channels_first, int B, int C, int T, int Cpad, void* out_act,
                                  int act_bf16, void* stream) {
  REQUIRE(x && out_act && B > 0 && C > 0 && T > 0 && Cpad >= C, "fs2_frames_to_rows: bad arguments");
  const long long n = (long long)B * (T + 2 * FS2_PAD) * Cpad;
  const int grid = (int)((n + THREADS - 1) / THREADS < 148 * 16 ? (n + THREADS - 1) / THREADS : 148 * 16);
  if (act_bf16) FS2_LAUNCH((frames_to_rows_kernel<bf16>), grid, THREADS, 0, ST, x, channels_first, B, C, T, Cpad, (bf16*)out_act);
  else FS2_LAUNCH((frames_to_rows_kernel<float>), grid, THREADS, 0, ST, x, channels_first, B, C, T, Cpad, (float*)out_act);
  return fs2_check_launch();
}

extern "C" int fs2_intensity_head(const float* h, const float* emb, const int64_t* emotions, const int* lens,
                                  const float* Wc, const float* bc, int B, int T, int D, int n_out, float* out,
                                  void* stream) {
  REQUIRE(h && emb && emotions && lens && Wc && bc && out && n_out >= 1 && n_out <= 8, "fs2_intensity_head: bad arguments");
  FS2_LAUNCH((intensity_head_kernel), grid_for_rows((long long)B * T), THREADS, 0, ST, h, emb, emotions, lens, Wc, bc, B, T, D,
             n_out, out);
  return fs2_check_launch();
}

extern "C" int fs2_collate(const int64_t* phon_cat, const int64_t* dur_cat, const float* mel_cat, const float* pitch_cat,
                           const float* energy_cat, const int* ph_start, const int* ph_len, const int* fr_start,
                           const int* fr_len, int B, int Tp, int Tm, int n_mels, int64_t* phoneme, int64_t* duration,
                           float* mel, float* pitch, float* energy, float* rank_X, void* stream) {
  REQUIRE(phon_cat && dur_cat && mel_cat && pitch_cat && energy_cat && ph_start && ph_len && fr_start && fr_len && phoneme &&
              duration && mel && pitch && energy && rank_X,
          "fs2_collate: null pointer");
  REQUIRE(B > 0 && B <= 65535 && Tp > 0 && Tm > 0 && n_mels > 0 && n_mels <= 254, "fs2_collate: bad sizes");
  const dim3 grid((Tm + COL_T - 1) / COL_T + 1, B);
  const size_t sm = (size_t)(n_mels + 2) * (COL_T + 1) * sizeof(float);
  FS2_LAUNCH((collate_kernel), grid, 256, sm, ST, phon_cat, dur_cat, mel_cat, pitch_cat, energy_cat, ph_start, ph_len, fr_start,
             fr_len, B, Tp, Tm, n_mels, phoneme, duration, mel, pitch, energy, rank_X);
  return fs2_check_launch();
}

extern "C" int fs2_gelu(void* x, long long n, int act_bf16, void* stream) {
  REQUIRE(x && n >= 0 && n % 4 == 0, "fs2_gelu: n must be a multiple of 4");
  if (n == 0) return FS2_OK;
  const long long blocks = (n / 4 + 255) / 256;
  const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
  if (act_bf16) FS2_LAUNCH((gelu_kernel<bf16>), grid, 256, 0, ST, (bf16*)x, n);
  else FS2_LAUNCH((gelu_kernel<float>), grid, 256, 0, ST, (float*)x, n);
  return fs2_check_launch();
}

extern "C" int fs2_prototype_buckets(const float* I, const int* lens, const int* group, const long long* frame_off,
                                     const long long* group_total, int N, int Tmax, int D, int n_groups, int n_buckets,
                                     float* sums, void* stream) {
  REQUIRE(I && lens && group && frame_off && group_total && sums && N > 0 && D > 0 && n_groups > 0 && n_buckets > 0,
          "fs2_prototype_buckets: bad arguments");
  FS2_LAUNCH((prototype_accum_kernel), N, 256, 0, ST, I, lens, group, frame_off, group_total, Tmax, D, n_buckets, sums);
  int rc = fs2_check_launch();
  if (rc) return rc;
  FS2_LAUNCH((prototype_finalize_kernel), 64, 256, 0, ST, group_total, D, n_groups, n_buckets, sums);
  return fs2_check_launch();
}
