// Fused multi-head self-attention for the FFT blocks (reference: speechbrain MultiheadAttention -> nn.MultiheadAttention,
// model.py:344-346, 425-427; mask quirk Q1 of SURVEY Appendix C), head_dim 192, bf16 operands, fp32 accumulation.
//
// One CTA owns 128 query rows of one (item, head) and walks the keys in blocks of 64:
//   warp 0   : TMA producer  (Q / dO tile once; K and V blocks through smem rings)
//   warp 1   : tcgen05.mma issuer: S = A.B1^T (128 x 64, K = 192) into a double-buffered TMEM tile, and the second
//              contraction  O += P.B2  (128 x 192, K = 64) whose A operand the softmax warps write to shared memory
//   warps 2-5: one thread per query row: tcgen05.ld of the score tile, mask / softmax / dropout, P -> smem (+ global for
//              the backward), final O: TMEM -> bf16 -> global
// forward : pass 1 computes the row max / sum (scores only), pass 2 recomputes the scores, forms P = softmax, applies
//           dropout and accumulates O = Pd.V; the T x T score matrix never exists in HBM in fp32 and no separate softmax
//           kernel runs.  P and Pd are still written (bf16) because the backward consumes them.
// backward: the same pipeline with (dO, V, K) in place of (Q, K, V): dPd = dO.V^T in TMEM, dS = scale * P * (dPd*keep -
//           rowsum(dO*O)) in registers, dS -> global (for dK = dS^T Q) and -> smem, dQ += dS.K accumulated in TMEM.
#include <cuda.h>
#include <cstring>
#include "common.cuh"
#include "../../include/fs2_b200.h"
#include "tc_ptx.cuh"

namespace {

long long* g_attn_dbg = nullptr;
// the cycle probe of tools/attn_bench.py (ATTN_DBG=1) is compiled in with `make EXTRA=-DFS2_TC_PROBE` only
#ifdef FS2_TC_PROBE
constexpr bool kAttnProbe = true;
#else
constexpr bool kAttnProbe = false;
#endif

constexpr int AQ = 128;        // query rows per CTA
constexpr int AK = 64;         // keys per block (one 128-byte swizzle row of P)
constexpr int HD = 192;        // head dimension: three 64-wide swizzle atoms
constexpr int KSTAGES = 3;
constexpr int VSTAGES = 2;
// NCH = threads per query row (column slices of a key block): 2 in the forward (8 softmax warps), 4 in the backward
// (16 warps: the dS math has more loads in flight per column) -- measured on B200, see profiles/r01_summary.md section 4
constexpr int att_threads(int nch) { return 32 * (2 + 4 * nch); }    // TMA, MMA, 4*NCH softmax warps
constexpr int Q_BYTES = 3 * AQ * 128;          // 49152
constexpr int KV_BYTES = 3 * AK * 128;         // 24576
constexpr int P_BYTES = AQ * 128;              // 16384
constexpr int ATT_SMEM = Q_BYTES + KSTAGES * KV_BYTES + VSTAGES * KV_BYTES + 2 * P_BYTES + 1024 + 512 + 4 * AQ * 8;
constexpr int TMEM_S = 0;      // two 64-column score buffers
constexpr int TMEM_O = 128;    // 192-column output accumulator

struct AttnParams {
  int B, H, T, TP, ldk, D;
  int b1_col, b2_col;          // column of the first / second B operand inside the qkv matrix (plus h*HD)
  const int* lens;
  float scale;
  DropCfg dc;
  const unsigned long long* seed_dev;
  bf16* P;                     // fwd: out (softmax), bwd: in
  bf16* Pd;                    // fwd: out (dropout(P)), may be NULL when drop_p == 0
  bf16* dS;                    // bwd: out
  bf16* out;                   // fwd: O [B*TP, D]; bwd: dQ inside dqkv [B*TP, 3D]
  long long out_ld;
  const bf16* O;               // bwd: forward output (for rowsum(dO*O))
  const bf16* dO;              // bwd
  int* err;
  long long* dbg;              // optional cycle breakdown of CTA 0 (measurements only)
  int plain_mask;              // 0: FastSpeech2's attn_mask quirk, 1: plain key-padding mask
};

__device__ __forceinline__ int attn_kv(const int* lens, int B, int H, int bh, int plain) {
  // FastSpeech2 (quirk Q1): keys valid for (b,h) are [0, min(len[b], len[(b*H+h) % B]));
  // plain = 1: ordinary key-padding mask, keys [0, len[b])  (rank_model/model.py:34, 103)
  return plain ? lens[bh / H] : min(lens[bh / H], lens[bh % B]);
}

__device__ __forceinline__ float ex2f_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <bool BWD, int NCH>
__global__ void __launch_bounds__(att_threads(NCH), 1) attn_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmKV,
                                                              const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Q_BYTES;
  uint8_t* sV = sK + KSTAGES * KV_BYTES;
  uint8_t* sP = sV + VSTAGES * KV_BYTES;
  uint64_t* bars = (uint64_t*)(sP + 2 * P_BYTES);
  uint64_t* qfull = bars;                 // [1]
  uint64_t* kfull = qfull + 1;            // [KSTAGES]
  uint64_t* kempty = kfull + KSTAGES;
  uint64_t* vfull = kempty + KSTAGES;     // [VSTAGES]
  uint64_t* vempty = vfull + VSTAGES;
  uint64_t* sfull = vempty + VSTAGES;     // [2]
  uint64_t* sempty = sfull + 2;
  uint64_t* pfull = sempty + 2;           // [2]
  uint64_t* pempty = pfull + 2;
  uint64_t* ofull = pempty + 2;           // [1]
  uint32_t* tmem_slot = (uint32_t*)(ofull + 1);
  float2* xch = (float2*)(tmem_slot + 2);      // [NCH][AQ] exchange between the threads of a row

  constexpr int CW = AK / NCH;                 // score columns per thread and key block
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* err = p.err;
  const int nqb = (p.T + AQ - 1) / AQ;
  const int qb = blockIdx.x % nqb;
  const int bh = blockIdx.x / nqb;
  const int b = bh / p.H, h = bh % p.H;
  const int kv = attn_kv(p.lens, p.B, p.H, bh, p.plain_mask);
  const int nkb = (kv + AK - 1) / AK;
  const int njobs = BWD ? nkb : 2 * nkb;          // fwd: a statistics pass, then the P / PV pass
  const int first_main = BWD ? 0 : nkb;
  const int q0 = qb * AQ;
  const int row_base = b * p.TP + FS2_PAD;        // matrix row of (b, t = 0)

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(qfull), 1);
    for (int s = 0; s < KSTAGES; ++s) { mbar_init(smem_u32(&kfull[s]), 1); mbar_init(smem_u32(&kempty[s]), 1); }
    for (int s = 0; s < VSTAGES; ++s) { mbar_init(smem_u32(&vfull[s]), 1); mbar_init(smem_u32(&vempty[s]), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&sfull[s]), 1);
      mbar_init(smem_u32(&sempty[s]), 4 * NCH);
      mbar_init(smem_u32(&pfull[s]), 128 * NCH);
      mbar_init(smem_u32(&pempty[s]), 1);
    }
    mbar_init(smem_u32(ofull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // prologue above overlaps the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ------------------------------------------------------------------------------------ TMA producer
    if (lane == 0 && nkb > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
      mbar_expect_tx(smem_u32(qfull), Q_BYTES);
#pragma unroll
      for (int kc = 0; kc < 3; ++kc)
        tma_load_4d(smem_u32(sQ + kc * (AQ * 128)), &tmA, smem_u32(qfull), h * HD + kc * 64, row_base + q0, 0, 0);
      bool ok = true;
      int s = 0, sv = 0;
      uint32_t ph = 0, phv = 0;
      for (int job = 0; job < njobs && ok; ++job) {
        const int j = (job >= nkb) ? job - nkb : job;
        if (!mbar_wait(smem_u32(&kempty[s]), ph ^ 1, err)) { ok = false; break; }
        mbar_expect_tx(smem_u32(&kfull[s]), KV_BYTES);
#pragma unroll
        for (int kc = 0; kc < 3; ++kc)
          tma_load_4d(smem_u32(sK + s * KV_BYTES + kc * (AK * 128)), &tmKV, smem_u32(&kfull[s]), p.b1_col + h * HD + kc * 64,
                      row_base + j * AK, 0, 0);
        if (++s == KSTAGES) { s = 0; ph ^= 1; }
        if (job >= first_main) {
          if (!mbar_wait(smem_u32(&vempty[sv]), phv ^ 1, err)) { ok = false; break; }
          mbar_expect_tx(smem_u32(&vfull[sv]), KV_BYTES);
#pragma unroll
          for (int nc = 0; nc < 3; ++nc)
            tma_load_4d(smem_u32(sV + sv * KV_BYTES + nc * (AK * 128)), &tmKV, smem_u32(&vfull[sv]),
                        p.b2_col + h * HD + nc * 64, row_base + j * AK, 0, 0);
          if (++sv == VSTAGES) { sv = 0; phv ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0 && nkb > 0) {
      // D = f32, A = B = bf16; S: both K-major, N = 64; O: A K-major (P in smem), B MN-major (V / K rows), N = 192
      const uint32_t idescS = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AK >> 3) << 17) | ((uint32_t)(AQ >> 4) << 24);
      const uint32_t idescO = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) |
                              ((uint32_t)(AQ >> 4) << 24);
      bool ok = mbar_wait(smem_u32(qfull), 0, err);
      int s = 0, sv = 0;
      uint32_t ph = 0, phv = 0;
      const bool prof = kAttnProbe && p.dbg != nullptr && blockIdx.x == 0;
      long long w_k = 0, w_se = 0, w_pf = 0, w_v = 0, t_all = clock64(), tt = 0;
      for (int job = 0; job <= njobs && ok; ++job) {
        if (job < njobs) {
          const int sb = job & 1;
          if (prof) tt = clock64();
          if (!mbar_wait(smem_u32(&kfull[s]), ph, err)) { ok = false; break; }
          if (prof) { long long n = clock64(); w_k += n - tt; tt = n; }
          if (!mbar_wait(smem_u32(&sempty[sb]), ((job >> 1) & 1) ^ 1, err)) { ok = false; break; }
          if (prof) w_se += clock64() - tt;
          tc_fence_after();
          const uint32_t sk = smem_u32(sK + s * KV_BYTES);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint64_t ad = smem_desc(smem_u32(sQ) + (k >> 2) * (AQ * 128) + (k & 3) * 32, 16, 1024);
            const uint64_t bd = smem_desc(sk + (k >> 2) * (AK * 128) + (k & 3) * 32, 16, 1024);
            umma_bf16(tmem_base + TMEM_S + sb * AK, ad, bd, idescS, k > 0 ? 1u : 0u);
          }
          umma_commit(smem_u32(&kempty[s]));
          umma_commit(smem_u32(&sfull[sb]));
          if (++s == KSTAGES) { s = 0; ph ^= 1; }
        }
        const int jv = job - 1 - first_main;      // second contraction of the previous main-pass block
        if (jv >= 0) {
          const int pb = jv & 1;
          if (prof) tt = clock64();
          if (!mbar_wait(smem_u32(&pfull[pb]), (jv >> 1) & 1, err)) { ok = false; break; }
          if (prof) { long long n = clock64(); w_pf += n - tt; tt = n; }
          if (!mbar_wait(smem_u32(&vfull[sv]), phv, err)) { ok = false; break; }
          if (prof) w_v += clock64() - tt;
          tc_fence_after();
          const uint32_t sp = smem_u32(sP + pb * P_BYTES);
          const uint32_t svb = smem_u32(sV + sv * KV_BYTES);
#pragma unroll
          for (int k = 0; k < AK / 16; ++k) {
            const uint64_t ad = smem_desc(sp + k * 32, 16, 1024);
            const uint64_t bd = smem_desc(svb + k * 2048, AK * 128, 1024);
            umma_bf16(tmem_base + TMEM_O, ad, bd, idescO, (jv > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&pempty[pb]));
          umma_commit(smem_u32(&vempty[sv]));
          if (++sv == VSTAGES) { sv = 0; phv ^= 1; }
          if (jv == nkb - 1) umma_commit(smem_u32(ofull));
        }
      }
      if (prof) {
        p.dbg[0] = clock64() - t_all; p.dbg[1] = w_k; p.dbg[2] = w_se; p.dbg[3] = w_pf; p.dbg[4] = w_v; p.dbg[5] = njobs;
      }
    }
  } else {
    // ------------------------------------------------------------------------------------ softmax / dS warps
    // 16 warps: TMEM lane quadrant q = warp % 4 (32 query rows), column quarter ch = (warp - 2) / 4 (16 of the block's 64
    // keys); four warps per scheduler hide each other's latency.  The four threads of a row combine their statistics once.
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int t = q0 + row;
    const bool row_valid = t < p.T;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    DropCfg dc = p.dc;
    if (p.seed_dev) dc.seed ^= mix64(*p.seed_dev);
    const long long ro = ((long long)bh * p.T + t) * p.ldk;       // element offset of this row in P / Pd / dS
    const float sc2 = p.scale * 1.4426950408889634f;              // scores in the log2 domain: exp(x) = 2^(x*log2 e)
    float mx = -INFINITY, sum = 0.f, inv = 0.f, dsum = 0.f;
    bool ok = true;
    if (BWD && nkb > 0) {
      // D_i = sum_c dO[t,c] * O[t,c]  (== sum_k Pd*dPd, the softmax-backward row term); each thread takes 48 of the 192 dims
      if (row_valid) {
        const uint4* a = reinterpret_cast<const uint4*>(p.dO + (long long)(row_base + t) * p.D + h * HD + ch * (HD / NCH));
        const uint4* o = reinterpret_cast<const uint4*>(p.O + (long long)(row_base + t) * p.D + h * HD + ch * (HD / NCH));
#pragma unroll
        for (int i = 0; i < HD / NCH / 8; ++i) {
          const uint4 x = a[i], y = o[i];
          const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw[k]));
            const float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yw[k]));
            dsum += fx.x * fy.x + fx.y * fy.y;
          }
        }
      }
      xch[ch * AQ + row] = make_float2(dsum, 0.f);
      asm volatile("bar.sync 1, %0;" ::"n"(128 * NCH) : "memory");
      dsum = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) dsum += xch[c * AQ + row].x;
    }
    const bool prof = kAttnProbe && p.dbg != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0;
    long long w_sf = 0, w_ld = 0, w_cmp = 0, w_pe = 0, w_st = 0, t_all = clock64(), tt = 0;
    uint4 pk[CW / 8];                                             // bwd: this job's P chunk (packed bf16), prefetched
    auto load_p = [&](int j) {
      const int c0 = j * AK + ch * CW;
#pragma unroll
      for (int i = 0; i < CW / 8; ++i) pk[i] = make_uint4(0, 0, 0, 0);
      if (row_valid && j < nkb) {
        const uint4* src = reinterpret_cast<const uint4*>(p.P + ro + c0);
#pragma unroll
        for (int i = 0; i < CW / 8; ++i)
          if (c0 + 8 * i < p.ldk) pk[i] = src[i];
      }
    };
    if (BWD) load_p(0);
    for (int job = 0; job < njobs && ok; ++job) {
      const int sb = job & 1;
      const bool main_pass = job >= first_main;
      const int j = main_pass ? job - first_main : job;
      const int c0 = j * AK + ch * CW;                            // first key of this thread's 16 columns
      const int pb = j & 1;
      const bool full = (c0 + CW <= kv);
      const bool in_buf = row_valid && c0 < p.ldk;
      if (!BWD && job == first_main) {
        // combine the column quarters' (max, sum) of pass 1
        xch[ch * AQ + row] = make_float2(mx, sum);
        asm volatile("bar.sync 1, %0;" ::"n"(128 * NCH) : "memory");
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < NCH; ++c) m = fmaxf(m, xch[c * AQ + row].x);
        float l = 0.f;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          const float2 o = xch[c * AQ + row];
          l += o.y * ex2f_(o.x - m);                              // empty quarter: 0 * 2^-inf = 0
        }
        mx = m;
        sum = l;
        inv = 1.0f / sum;
      }
      if (prof) tt = clock64();
      if (!mbar_wait(smem_u32(&sfull[sb]), (job >> 1) & 1, err)) { ok = false; break; }
      if (prof) { long long n = clock64(); w_sf += n - tt; tt = n; }
      tc_fence_after();
      uint32_t r[CW];
      if constexpr (CW == 32) tmem_ld32(lane_addr + (uint32_t)(TMEM_S + sb * AK + ch * CW), r);
      else tmem_ld16(lane_addr + (uint32_t)(TMEM_S + sb * AK + ch * CW), r);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sempty[sb]));          // score buffer drained by this warp
      if (prof) { long long n = clock64(); w_ld += n - tt; tt = n; }
      if (!main_pass) {
        // ---- pass 1 (forward only): running max / sum over this thread's columns
        float bm = -INFINITY;
        if (full) {
#pragma unroll
          for (int i = 0; i < CW; ++i) bm = fmaxf(bm, __uint_as_float(r[i]));
        } else {
#pragma unroll
          for (int i = 0; i < CW; ++i) bm = fmaxf(bm, (c0 + i < kv) ? __uint_as_float(r[i]) : -INFINITY);
        }
        bm *= sc2;
        if (bm > mx) { sum *= ex2f_(mx - bm); mx = bm; }          // mx == -inf: sum is 0 and 2^-inf = 0
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (full) {
#pragma unroll
          for (int i = 0; i < CW; i += 4) {
            a0 += ex2f_(fmaf(__uint_as_float(r[i]), sc2, -mx));
            a1 += ex2f_(fmaf(__uint_as_float(r[i + 1]), sc2, -mx));
            a2 += ex2f_(fmaf(__uint_as_float(r[i + 2]), sc2, -mx));
            a3 += ex2f_(fmaf(__uint_as_float(r[i + 3]), sc2, -mx));
          }
        } else if (c0 < kv) {
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const float e = ex2f_(fmaf(__uint_as_float(r[i]), sc2, -mx));
            a0 += (c0 + i < kv) ? e : 0.f;
          }
        }
        sum += (a0 + a1) + (a2 + a3);
        continue;
      }
      // ---- main pass
      float v[CW];
      if (!BWD) {
        if (full) {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = ex2f_(fmaf(__uint_as_float(r[i]), sc2, -mx)) * inv;
        } else {
#pragma unroll
          for (int i = 0; i < CW; ++i) {
            const float e = ex2f_(fmaf(__uint_as_float(r[i]), sc2, -mx)) * inv;
            v[i] = (c0 + i < kv) ? e : 0.f;
          }
        }
        if (in_buf && p.P) {          // P is kept for the backward only (NULL in inference)
          uint4* dst = reinterpret_cast<uint4*>(p.P + ro + c0);
#pragma unroll
          for (int i = 0; i < CW / 8; ++i)
            if (c0 + 8 * i < p.ldk)
              dst[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                  pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
        }
        if (dc.p > 0.f) {
#pragma unroll
          for (int i = 0; i < CW / 4; ++i) {
            const float4 k = drop_scale4(dc, (uint64_t)(ro + c0 + 4 * i) >> 2);
            v[4 * i] *= k.x; v[4 * i + 1] *= k.y; v[4 * i + 2] *= k.z; v[4 * i + 3] *= k.w;
          }
          if (in_buf && p.Pd) {
            uint4* dst = reinterpret_cast<uint4*>(p.Pd + ro + c0);
#pragma unroll
            for (int i = 0; i < CW / 8; ++i)
              if (c0 + 8 * i < p.ldk)
                dst[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                    pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          }
        }
      } else {
        // dS = scale * P * (dPd * keep - D_i); P is 0 beyond kv, so no explicit mask is needed
        float pr[CW];
#pragma unroll
        for (int i = 0; i < CW / 8; ++i) {
          const uint32_t xw[4] = {pk[i].x, pk[i].y, pk[i].z, pk[i].w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw[k]));
            pr[8 * i + 2 * k] = f.x;
            pr[8 * i + 2 * k + 1] = f.y;
          }
        }
        load_p(j + 1);                                            // next block's P: a whole job of latency slack
        const float scale = p.scale;
        if (dc.p > 0.f) {
#pragma unroll
          for (int i = 0; i < CW / 4; ++i) {
            const float4 k = drop_scale4(dc, (uint64_t)(ro + c0 + 4 * i) >> 2);
            v[4 * i] = scale * pr[4 * i] * fmaf(__uint_as_float(r[4 * i]), k.x, -dsum);
            v[4 * i + 1] = scale * pr[4 * i + 1] * fmaf(__uint_as_float(r[4 * i + 1]), k.y, -dsum);
            v[4 * i + 2] = scale * pr[4 * i + 2] * fmaf(__uint_as_float(r[4 * i + 2]), k.z, -dsum);
            v[4 * i + 3] = scale * pr[4 * i + 3] * fmaf(__uint_as_float(r[4 * i + 3]), k.w, -dsum);
          }
        } else {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = scale * pr[i] * (__uint_as_float(r[i]) - dsum);
        }
        if (in_buf) {
          uint4* dst = reinterpret_cast<uint4*>(p.dS + ro + c0);
#pragma unroll
          for (int i = 0; i < CW / 8; ++i)
            if (c0 + 8 * i < p.ldk)
              dst[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                  pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
        }
      }
      // the P tile in smem is free once the second contraction of block j-2 has completed
      if (prof) { long long n = clock64(); w_cmp += n - tt; tt = n; }
      if (!mbar_wait(smem_u32(&pempty[pb]), ((j >> 1) & 1) ^ 1, err)) { ok = false; break; }
      if (prof) { long long n = clock64(); w_pe += n - tt; tt = n; }
      // A operand of the second contraction: row `row`, 16-byte chunks XOR-swizzled inside the 128-byte row
      uint8_t* prow = sP + pb * P_BYTES + row * 128;
#pragma unroll
      for (int i = 0; i < CW / 8; ++i) {
        const int chunk = ch * (CW / 8) + i;
        *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) =
            make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                       pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMA
      mbar_arrive(smem_u32(&pfull[pb]));
      if (prof) w_st += clock64() - tt;
    }
    if (prof) {
      p.dbg[8] = clock64() - t_all; p.dbg[9] = w_sf; p.dbg[10] = w_ld; p.dbg[11] = w_cmp; p.dbg[12] = w_pe; p.dbg[13] = w_st;
    }
    // columns of P / Pd / dS beyond the last processed key block are zero (the backward GEMMs read the full rows)
    if (row_valid) {
      for (int c = nkb * AK + ch * 8; c < p.ldk; c += 8 * NCH) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        if (!BWD) {
          if (p.P) *reinterpret_cast<uint4*>(p.P + ro + c) = z;
          if (p.Pd && dc.p > 0.f) *reinterpret_cast<uint4*>(p.Pd + ro + c) = z;
        } else {
          *reinterpret_cast<uint4*>(p.dS + ro + c) = z;
        }
      }
    }
    // final accumulator: TMEM -> bf16 -> global; each thread writes 48 of the row's 192 columns
    bf16* orow = p.out + (long long)(row_base + t) * p.out_ld + h * HD + ch * (HD / NCH);
    if (nkb > 0) {
      if (ok && mbar_wait(smem_u32(ofull), 0, err)) {
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HD / NCH / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(lane_addr + (uint32_t)(TMEM_O + ch * (HD / NCH) + c * 16), r);
          if (row_valid) {
            uint4* dst = reinterpret_cast<uint4*>(orow + c * 16);
#pragma unroll
            for (int i = 0; i < 2; ++i)
              dst[i] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1])),
                                  pack_bf16x2(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3])),
                                  pack_bf16x2(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5])),
                                  pack_bf16x2(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7])));
          }
        }
      }
    } else if (row_valid) {
      for (int c = 0; c < HD / NCH; c += 8) *reinterpret_cast<uint4*>(orow + c) = make_uint4(0, 0, 0, 0);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <bool BWD, int NCH>
int launch_attn(const CUtensorMap& ta, const CUtensorMap& tkv, const AttnParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(attn_kernel<BWD, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    configured = true;
  }
  const int nqb = (p.T + AQ - 1) / AQ;
  FS2_LAUNCH((attn_kernel<BWD, NCH>), nqb * p.B * p.H, att_threads(NCH), ATT_SMEM, st, ta, tkv, p);
  return fs2_check_launch();
}

int attn_common(const void* qkv, const int* lens, int B, int H, int T, int D, int ldk, AttnParams& p, CUtensorMap* tkv) {
  if (!qkv || !lens || B <= 0 || H <= 0 || T <= 0 || D != H * HD || ldk % 8 != 0 || ldk < T) {
    fs2_set_error("fs2_attn: bad arguments (needs head_dim 192, bf16, ldk a multiple of 8)");
    return FS2_ERR_ARG;
  }
  memset(&p, 0, sizeof p);
  p.B = B; p.H = H; p.T = T; p.TP = T + 2 * FS2_PAD; p.ldk = ldk; p.D = D;
  p.lens = lens;
  int rc = fs2_tc_error_ptr(&p.err);
  if (rc) return rc;
  p.dbg = g_attn_dbg;
  return fs2_tc_make_map_2d(qkv, 3LL * D, (long long)B * p.TP, 3LL * D, 64, AK, tkv);
}

}  // namespace

/* measurement hook: cycle breakdown of CTA 0 (MMA thread: [0] total, [1..4] waits on K tile / score buffer / P tile /
 * V tile, [5] jobs; first softmax thread: [8] total, [9] wait scores, [10] tcgen05.ld, [11] math + global stores,
 * [12] wait P buffer, [13] smem store + fence + arrive) */
extern "C" int fs2_attn_set_debug(long long* dev_buf) {
  g_attn_dbg = dev_buf;
  return FS2_OK;
}

extern "C" int fs2_attn_fwd(const void* qkv, const int* lens, int B, int H, int T, int D, int ldk, float scale, float drop_p,
                            unsigned long long seed, const unsigned long long* seed_dev, void* P, void* Pd, void* O,
                            void* stream) {
  return fs2_attn_fwd_ex(qkv, lens, B, H, T, D, ldk, scale, drop_p, seed, seed_dev, P, Pd, O, 0, stream);
}

extern "C" int fs2_attn_fwd_ex(const void* qkv, const int* lens, int B, int H, int T, int D, int ldk, float scale,
                               float drop_p, unsigned long long seed, const unsigned long long* seed_dev, void* P, void* Pd,
                               void* O, int plain_mask, void* stream) {
  AttnParams p;
  CUtensorMap tq, tkv;
  int rc = attn_common(qkv, lens, B, H, T, D, ldk, p, &tkv);
  if (rc) return rc;
  if (!O) { fs2_set_error("fs2_attn_fwd: null pointer"); return FS2_ERR_ARG; }
  rc = fs2_tc_make_map_2d(qkv, 3LL * D, (long long)B * p.TP, 3LL * D, 64, AQ, &tq);
  if (rc) return rc;
  p.b1_col = D;          // K
  p.b2_col = 2 * D;      // V
  p.scale = scale;
  p.dc = DropCfg{drop_p, seed};
  p.seed_dev = seed_dev;
  p.P = (bf16*)P;
  p.Pd = drop_p > 0.f ? (bf16*)Pd : nullptr;
  p.out = (bf16*)O;
  p.out_ld = D;
  p.plain_mask = plain_mask;
  return launch_attn<false, 2>(tq, tkv, p, (cudaStream_t)stream);
}

extern "C" int fs2_attn_bwd(const void* dO, const void* O, const void* qkv, const void* P, const int* lens, int B, int H,
                            int T, int D, int ldk, float scale, float drop_p, unsigned long long seed,
                            const unsigned long long* seed_dev, void* dS, void* dqkv, void* stream) {
  AttnParams p;
  CUtensorMap ta, tkv;
  int rc = attn_common(qkv, lens, B, H, T, D, ldk, p, &tkv);
  if (rc) return rc;
  if (!dO || !O || !P || !dS || !dqkv) { fs2_set_error("fs2_attn_bwd: null pointer"); return FS2_ERR_ARG; }
  rc = fs2_tc_make_map_2d(dO, D, (long long)B * p.TP, D, 64, AQ, &ta);
  if (rc) return rc;
  p.b1_col = 2 * D;      // dPd = dO . V^T
  p.b2_col = D;          // dQ  = dS . K
  p.scale = scale;
  p.dc = DropCfg{drop_p, seed};
  p.seed_dev = seed_dev;
  p.P = (bf16*)const_cast<void*>(P);
  p.dS = (bf16*)dS;
  p.out = (bf16*)dqkv;   // dQ occupies columns [h*192, (h+1)*192) of the (rows, 3D) gradient
  p.out_ld = 3LL * D;
  p.O = (const bf16*)O;
  p.dO = (const bf16*)dO;
  return launch_attn<true, 4>(ta, tkv, p, (cudaStream_t)stream);
}
