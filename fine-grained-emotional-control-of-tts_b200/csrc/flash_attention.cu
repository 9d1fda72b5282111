// Flash-style multi-head self-attention for the FFT blocks (reference: speechbrain MultiheadAttention ->
// nn.MultiheadAttention, model.py:344-346, 425-427; mask quirk Q1 of SURVEY Appendix C), head_dim 192, bf16 operands,
// fp32 accumulation.  Nothing of size T x T ever reaches HBM: the forward keeps O and one fp32 statistic per query row
// (log2-domain log-sum-exp), the backward recomputes the probabilities from it.
//
//   fa_fwd_kernel     one CTA = 128 query rows of one (item, head), keys in blocks of 64
//       warp 0   TMA producer (Q tile once, K / V blocks through rings)
//       warp 1   tcgen05.mma issuer: S = Q.K^T (128 x 64) into a double-buffered TMEM tile, O += P.V (128 x 192)
//       warps 2+ softmax: tcgen05.ld of the score tile, pass 1 = row maximum only (no exponentials), pass 2 = recompute
//                the scores, p = 2^(s - max) (un-normalised), dropout, P -> TENSOR MEMORY (bf16 pairs, the A operand
//                of the second contraction, tcgen05.st) ; O is scaled by 1 / sum in the epilogue
//   fa_bwd_kernel     ONE launch for the backward, two kinds of CTA (after fa_rowdot_kernel: D = scale * rowsum(dO * O)):
//     dQ tiles   same tiling as the forward: S = Q.K^T and dPd = dO.V^T per key block (TMEM), P = 2^(s - lse),
//                dS = scale * P * (dPd * keep - rowsum(dO * O)) -> tensor memory -> dQ += dS.K (TMEM accumulator)
//     dK/dV tiles one CTA = 128 KEY rows of one (item, head), queries in blocks of 32: S^T = K.Q^T and
//                dPd^T = V.dO^T (TMEM, two buffers worked by two warp groups in ping-pong), P^T / dS^T overwrite their own
//                score columns in tensor memory, dV += Pd^T.dO and dK += dS^T.Q (two 192-column TMEM accumulators)
#include <cuda.h>
#include <cstring>
#include "common.cuh"
#include "../../include/fs2_b200.h"
#include "tc_ptx.cuh"

namespace {

constexpr int AQ = 128;        // rows (queries, or keys in the dK/dV kernel) per CTA
constexpr int AK = 64;         // keys per block (forward, dQ kernel)
constexpr int BQ = 32;         // queries per block (dK/dV kernel)
constexpr int HD = 192;        // head dimension: three 64-wide swizzle atoms
constexpr int T128 = 3 * AQ * 128;     // 49152: [128 rows x 192] bf16 as three [rows x 64] 128-byte-swizzled atoms
constexpr int T64 = 3 * AK * 128;      // 24576
constexpr int T32 = 3 * BQ * 128;      // 12288
constexpr float LOG2E = 1.4426950408889634f;
constexpr int FA_MAXT = 1536;          // longest sequence the dropout word table in shared memory covers (training only)
constexpr int FA_TAB = FA_MAXT * 4;

// cycle probe of tools/flash_bench.py (FLASH_DBG=1): compiled in with `make EXTRA=-DFS2_TC_PROBE` only
long long* g_fa_dbg = nullptr;
#ifdef FS2_TC_PROBE
constexpr bool kProbe = true;
#else
constexpr bool kProbe = false;
#endif
#define PROBE_T(acc) do { if (prof) { const long long n__ = clock64(); (acc) += n__ - tt; tt = n__; } } while (0)

struct FaParams {
  int B, H, T, TP, D, Tl;
  const int* lens;
  float scale, drop_p;
  unsigned long long seed;
  const unsigned long long* seed_dev;
  float* lse;                  // [B*H][Tl] log2-domain row statistic (fwd: out, may be NULL; bwd: in)
  float* dvec;                 // [B*H][Tl] rowsum(dO * O) (dq kernel: out; dkv kernel: in)
  bf16* out;                   // fwd: O [B*TP, D]; bwd: dqkv [B*TP, 3D]
  const bf16* O;               // bwd: forward output
  const bf16* dO;              // bwd
  int* err;
  int plain_mask;              // 0: FastSpeech2's attn_mask quirk, 1: plain key-padding mask
  long long* dbg;              // optional cycle breakdown of CTA 0 (measurements only)
};

__device__ __forceinline__ int fa_kv(const int* lens, int B, int H, int bh, int plain) {
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
__device__ __forceinline__ void tmem_alloc512(uint32_t* slot) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(512u) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free512(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512u) : "memory");
}
// instruction descriptor: D = f32, A = B = bf16, M = 128, N = n; b_mn = 1: B is MN-major (rows of V / K / Q / dO used as
// the K dimension)
__device__ __forceinline__ uint32_t fa_idesc(int n, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(AQ >> 4) << 24);
}
// Shared-memory descriptors: everything but the 14-bit start-address field is constant, and that field only ever moves by
// (bytes >> 4) -- one descriptor per operand tile, then plain 64-bit adds of compile-time constants (the MMA-issuing
// thread shares its scheduler with softmax warps: every instruction between two tcgen05.mma counts).
// K-major operand tile [rows x 192]: K step k (16 columns) of the 64-column atom k / 4
__device__ __forceinline__ uint64_t fa_base_k(uint32_t base) { return smem_desc(base, 16, 1024); }
__host__ __device__ constexpr uint64_t fa_off_k(int rows, int k) {
  return (uint64_t)(((k >> 2) * (rows * 128) + (k & 3) * 32) >> 4);
}
// the same tile as an MN-major B operand (N = the 192 columns, K = its rows): K step k = rows [16k, 16k + 16)
__device__ __forceinline__ uint64_t fa_base_mn(uint32_t base, int rows) { return smem_desc(base, rows * 128, 1024); }
__host__ __device__ constexpr uint64_t fa_off_mn(int k) { return (uint64_t)((k * 2048) >> 4); }

// ---- per-job math of the softmax threads, specialised on (block fully inside the valid range, dropout on): the generic
// form costs two compares and a select more per element, and these threads are instruction-issue bound
template <int CW, bool FULL, bool DROP>
__device__ __forceinline__ float fwd_probs(const uint32_t* r, float* v, float sc2, float mx, int c0, int kv, uint32_t rw,
                                           const uint32_t* colw, uint32_t thr) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int i = 0; i < CW; i += 4) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float x = ex2f_(fmaf(__uint_as_float(r[i + e]), sc2, -mx));
      if (!FULL) x = (c0 + i + e < kv) ? x : 0.f;
      v[i + e] = x;
    }
    a0 += v[i]; a1 += v[i + 1]; a2 += v[i + 2]; a3 += v[i + 3];
  }
  if (DROP) {
#pragma unroll
    for (int i = 0; i < CW; i += 4) {      // keep or zero; the 1 / (1 - p) factor is folded into the final normalisation
      const uint4 cw = *reinterpret_cast<const uint4*>(colw + c0 + i);
      v[i] = adrop_keep(rw, cw.x, thr) ? v[i] : 0.f;
      v[i + 1] = adrop_keep(rw, cw.y, thr) ? v[i + 1] : 0.f;
      v[i + 2] = adrop_keep(rw, cw.z, thr) ? v[i + 2] : 0.f;
      v[i + 3] = adrop_keep(rw, cw.w, thr) ? v[i + 3] : 0.f;
    }
  }
  return (a0 + a1) + (a2 + a3);
}
// dS = scale * P * (dPd * keep / (1 - p) - D) = P * fma(dPd, keep ? scale / (1 - p) : 0, -scale * D)
template <int CW, bool FULL, bool DROP>
__device__ __forceinline__ void dq_ds(const uint32_t* r, const uint32_t* g, float* v, float sc2, float lse, float ks_s, float ds_s,
                                      int c0, int kv, uint32_t rw, const uint32_t* colw, uint32_t thr) {
#pragma unroll
  for (int i4 = 0; i4 < CW; i4 += 4) {
    uint4 cw = make_uint4(0, 0, 0, 0);
    if (DROP) cw = *reinterpret_cast<const uint4*>(colw + c0 + i4);
    const uint32_t cws[4] = {cw.x, cw.y, cw.z, cw.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = i4 + e;
      float pr = ex2f_(fmaf(__uint_as_float(r[i]), sc2, -lse));
      if (!FULL) pr = (c0 + i < kv) ? pr : 0.f;
      float mul = ks_s;
      if (DROP) mul = adrop_keep(rw, cws[e], thr) ? ks_s : 0.f;
      v[i] = pr * fmaf(__uint_as_float(g[i]), mul, -ds_s);
    }
  }
}
// the dK/dV kernel's 16 queries of one key row: P^T (keep or zero) and dS^T, packed bf16 pairs
template <bool FULL, bool DROP>
__device__ __forceinline__ void dkv_pt(const uint32_t* r, const uint32_t* d, const float* stat_l, const float* stat_d, uint32_t* pp,
                                       uint32_t* ps, float sc2, float ks_s, int t0, int T, uint32_t cw, const uint32_t* roww,
                                       uint32_t thr) {
#pragma unroll
  for (int e4 = 0; e4 < 4; ++e4) {
    const float4 l4 = reinterpret_cast<const float4*>(stat_l)[e4], d4 = reinterpret_cast<const float4*>(stat_d)[e4];
    uint4 rw4 = make_uint4(0, 0, 0, 0);
    if (DROP) rw4 = *reinterpret_cast<const uint4*>(roww + t0 + 4 * e4);
    const uint32_t rws[4] = {rw4.x, rw4.y, rw4.z, rw4.w};
    const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};     // dd = scale * rowsum(dO * O)
    float pd[4], ds[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * e4 + e;
      float pr = ex2f_(fmaf(__uint_as_float(r[j]), sc2, -ls[e]));
      float ddv = dd[e];
      if (!FULL) { const bool ok_t = t0 + j < T; pr = ok_t ? pr : 0.f; ddv = ok_t ? ddv : 0.f; }
      float mul = ks_s;
      pd[e] = pr;                              // keep or zero; 1 / (1 - p) is applied to dV in the epilogue
      if (DROP) {
        const bool keep = adrop_keep(rws[e], cw, thr);
        mul = keep ? ks_s : 0.f;
        pd[e] = keep ? pr : 0.f;
      }
      ds[e] = pr * fmaf(__uint_as_float(d[j]), mul, -ddv);
    }
    pp[2 * e4] = pack_bf16x2(pd[0], pd[1]);
    pp[2 * e4 + 1] = pack_bf16x2(pd[2], pd[3]);
    ps[2 * e4] = pack_bf16x2(ds[0], ds[1]);
    ps[2 * e4 + 1] = pack_bf16x2(ds[2], ds[3]);
  }
}

// ===================================================================================================== forward
#ifndef FS2_FWD_NCH
#define FS2_FWD_NCH 2
#endif
constexpr int FWD_NCH = FS2_FWD_NCH;             // threads per query row (column slices of a key block)
constexpr int FWD_THREADS = 32 * (2 + 4 * FWD_NCH);
template <bool PT> struct FwdCfg {
  static constexpr int KST = PT ? 4 : 3, VST = PT ? 3 : 2;
  static constexpr int SMEM = T128 + (KST + VST) * T64 + (PT ? 0 : 2 * AQ * 128) + FA_TAB + 1024 + 512 + FWD_NCH * AQ * 4;
};
constexpr int F_TMEM_S = 0, F_TMEM_O = 128, F_TMEM_P = 320;      // S: 2 x 64 columns, O: 192, P: 2 x 32 (bf16 pairs)

template <bool PT>
__global__ void __launch_bounds__(FWD_THREADS, 1) fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                const __grid_constant__ CUtensorMap tmKV,
                                                                const FaParams p) {
  constexpr int NCH = FWD_NCH, KST = FwdCfg<PT>::KST, VST = FwdCfg<PT>::VST, CW = AK / NCH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + T128;
  uint8_t* sV = sK + KST * T64;
  uint8_t* sP = sV + VST * T64;                              // !PT only: two [128 x 64] bf16 tiles
  uint32_t* sTab = (uint32_t*)(sP + (PT ? 0 : 2 * AQ * 128));     // dropout: one word per key of this (item, head)
  uint64_t* bars = (uint64_t*)(sTab + FA_MAXT);
  uint64_t* qfull = bars;
  uint64_t* kfull = qfull + 1;
  uint64_t* kempty = kfull + KST;
  uint64_t* vfull = kempty + KST;
  uint64_t* vempty = vfull + VST;
  uint64_t* sfull = vempty + VST;
  uint64_t* sempty = sfull + 2;
  uint64_t* pfull = sempty + 2;
  uint64_t* pempty = pfull + 2;
  uint64_t* ofull = pempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(ofull + 1);
  float* xch = (float*)(tmem_slot + 2);      // [NCH][AQ] exchange between the threads of a row

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* err = p.err;
  const int nqb = (p.T + AQ - 1) / AQ;
  const int qb = blockIdx.x % nqb;
  const int bh = blockIdx.x / nqb;
  const int b = bh / p.H, h = bh % p.H;
  const int kv = fa_kv(p.lens, p.B, p.H, bh, p.plain_mask);
  const int nkb = (kv + AK - 1) / AK;
  const int njobs = 2 * nkb;                      // the maximum pass, then the P / PV pass
  const int q0 = qb * AQ;
  const int row_base = b * p.TP + FS2_PAD;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(qfull), 1);
    for (int s = 0; s < KST; ++s) { mbar_init(smem_u32(&kfull[s]), 1); mbar_init(smem_u32(&kempty[s]), 1); }
    for (int s = 0; s < VST; ++s) { mbar_init(smem_u32(&vfull[s]), 1); mbar_init(smem_u32(&vempty[s]), 1); }
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
  if (warp == 1) tmem_alloc512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // warp-uniform for the compiler: no per-MMA election loop
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------------------------ TMA producer
    if (nkb > 0 && elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
      mbar_expect_tx(smem_u32(qfull), T128);
#pragma unroll
      for (int kc = 0; kc < 3; ++kc)
        tma_load_4d(smem_u32(sQ + kc * (AQ * 128)), &tmQ, smem_u32(qfull), h * HD + kc * 64, row_base + q0, 0, 0);
      int s = 0, sv = 0;
      uint32_t ph = 0, phv = 0;
      for (int job = 0; job < njobs; ++job) {
        const int j = (job >= nkb) ? job - nkb : job;
        if (!mbar_wait(smem_u32(&kempty[s]), ph ^ 1, err)) break;
        mbar_expect_tx(smem_u32(&kfull[s]), T64);
#pragma unroll
        for (int kc = 0; kc < 3; ++kc)
          tma_load_4d(smem_u32(sK + s * T64 + kc * (AK * 128)), &tmKV, smem_u32(&kfull[s]), p.D + h * HD + kc * 64,
                      row_base + j * AK, 0, 0);
        if (++s == KST) { s = 0; ph ^= 1; }
        if (job >= nkb) {
          if (!mbar_wait(smem_u32(&vempty[sv]), phv ^ 1, err)) break;
          mbar_expect_tx(smem_u32(&vfull[sv]), T64);
#pragma unroll
          for (int nc = 0; nc < 3; ++nc)
            tma_load_4d(smem_u32(sV + sv * T64 + nc * (AK * 128)), &tmKV, smem_u32(&vfull[sv]),
                        2 * p.D + h * HD + nc * 64, row_base + j * AK, 0, 0);
          if (++sv == VST) { sv = 0; phv ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------ MMA issuer
    if (nkb > 0 && elect_one()) {
      const uint32_t idescS = fa_idesc(AK, 0), idescO = fa_idesc(HD, 1);
      const uint64_t dQ0 = fa_base_k(smem_u32(sQ)), dK0 = fa_base_k(smem_u32(sK)), dV0 = fa_base_mn(smem_u32(sV), AK);
      const uint64_t dP0 = fa_base_k(smem_u32(sP));
      bool ok = mbar_wait(smem_u32(qfull), 0, err);
      int s = 0, sv = 0;
      uint32_t ph = 0, phv = 0;
      const bool prof = kProbe && p.dbg != nullptr && blockIdx.x == 0;
      long long w_k = 0, w_se = 0, w_pf = 0, w_v = 0, t_all = clock64(), tt = t_all;
      for (int job = 0; job <= njobs && ok; ++job) {
        if (job < njobs) {
          const int sb = job & 1;
          if (prof) tt = clock64();
          if (!mbar_wait(smem_u32(&kfull[s]), ph, err)) break;
          PROBE_T(w_k);
          if (!mbar_wait(smem_u32(&sempty[sb]), ((job >> 1) & 1) ^ 1, err)) break;
          PROBE_T(w_se);
          tc_fence_after();
          const uint64_t dk = dK0 + (uint64_t)s * (T64 >> 4);
          const uint32_t ts = tmem_base + F_TMEM_S + sb * AK;
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16(ts, dQ0 + fa_off_k(AQ, k), dk + fa_off_k(AK, k), idescS, k > 0 ? 1u : 0u);
          umma_commit(smem_u32(&kempty[s]));
          umma_commit(smem_u32(&sfull[sb]));
          if (++s == KST) { s = 0; ph ^= 1; }
        }
        const int jv = job - 1 - nkb;             // second contraction of the previous main-pass block
        if (jv >= 0) {
          const int pb = jv & 1;
          if (prof) tt = clock64();
          if (!mbar_wait(smem_u32(&pfull[pb]), (jv >> 1) & 1, err)) break;
          PROBE_T(w_pf);
          if (!mbar_wait(smem_u32(&vfull[sv]), phv, err)) break;
          PROBE_T(w_v);
          tc_fence_after();
          const uint64_t dv = dV0 + (uint64_t)sv * (T64 >> 4);
          const uint32_t tp = tmem_base + F_TMEM_P + pb * 32;
          const uint64_t dp = dP0 + (uint64_t)pb * ((AQ * 128) >> 4);
#pragma unroll
          for (int k = 0; k < AK / 16; ++k) {
            const uint32_t acc = (jv > 0 || k > 0) ? 1u : 0u;
            if constexpr (PT) umma_bf16_ts(tmem_base + F_TMEM_O, tp + k * 8, dv + fa_off_mn(k), idescO, acc);
            else umma_bf16(tmem_base + F_TMEM_O, dp + (uint64_t)(k * 2), dv + fa_off_mn(k), idescO, acc);
          }
          umma_commit(smem_u32(&pempty[pb]));
          umma_commit(smem_u32(&vempty[sv]));
          if (++sv == VST) { sv = 0; phv ^= 1; }
          if (jv == nkb - 1) umma_commit(smem_u32(ofull));
        }
      }
      if (prof) { p.dbg[16] = clock64() - t_all; p.dbg[17] = w_k; p.dbg[18] = w_se; p.dbg[19] = w_pf; p.dbg[20] = w_v; p.dbg[21] = njobs; }
    }
  } else {
    // ------------------------------------------------------------------------------------ softmax warps
    const int q = warp & 3;                        // TMEM lane quadrant of this warp
    const int ch = (warp - 2) >> 2;                // column half of the key block
    const int row = q * 32 + lane;
    const int t = q0 + row;
    const bool row_valid = t < p.T;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const ADrop dr = adrop_make(p.drop_p, p.seed, p.seed_dev, bh);
    const uint32_t tpart = adrop_row(dr, t);
    if (dr.thr != 0u) {
      for (int c = (warp - 2) * 32 + lane; c < nkb * AK; c += 128 * NCH) sTab[c] = adrop_col(dr, c);
      asm volatile("bar.sync 1, %0;" ::"n"(128 * NCH) : "memory");
    }
    const float sc2 = p.scale * LOG2E;             // scores in the log2 domain: exp(x) = 2^(x * log2 e)
    float mx = -INFINITY, sum = 0.f;
    const bool prof = kProbe && p.dbg != nullptr && blockIdx.x == 0 && warp == 2 && lane == 0;
    long long w1_sf = 0, w1_ld = 0, w1_m = 0, w2_sf = 0, w2_ld = 0, w2_m = 0, w2_pe = 0, w2_st = 0, t_all = clock64(), tt = t_all;
    // a failed (timed-out) wait sets the error word and the thread carries on: the named barriers below stay in step
    for (int job = 0; job < njobs; ++job) {
      const int sb = job & 1;
      const bool main_pass = job >= nkb;
      const int j = main_pass ? job - nkb : job;
      const int c0 = j * AK + ch * CW;
      const int pb = j & 1;
      const bool full = (c0 + CW <= kv);
      if (job == nkb) {
        // row maximum over both column halves, in the log2 domain
        xch[ch * AQ + row] = mx;
        asm volatile("bar.sync 1, %0;" ::"n"(128 * NCH) : "memory");
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < NCH; ++c) m = fmaxf(m, xch[c * AQ + row]);
        mx = m * sc2;
        asm volatile("bar.sync 1, %0;" ::"n"(128 * NCH) : "memory");     // xch is reused for the sums
      }
      if (prof) tt = clock64();
      mbar_wait(smem_u32(&sfull[sb]), (job >> 1) & 1, err);
      if (main_pass) PROBE_T(w2_sf); else PROBE_T(w1_sf);
      tc_fence_after();
      uint32_t r[CW];
      if constexpr (CW == 32) tmem_ld32(lane_addr + (uint32_t)(F_TMEM_S + sb * AK + ch * CW), r);
      else tmem_ld16(lane_addr + (uint32_t)(F_TMEM_S + sb * AK + ch * CW), r);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sempty[sb]));
      if (main_pass) PROBE_T(w2_ld); else PROBE_T(w1_ld);
      if (!main_pass) {
        float b0 = -INFINITY, b1 = -INFINITY, b2 = -INFINITY, b3 = -INFINITY;
        if (full) {
#pragma unroll
          for (int i = 0; i < CW; i += 4) {
            b0 = fmaxf(b0, __uint_as_float(r[i]));
            b1 = fmaxf(b1, __uint_as_float(r[i + 1]));
            b2 = fmaxf(b2, __uint_as_float(r[i + 2]));
            b3 = fmaxf(b3, __uint_as_float(r[i + 3]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < CW; ++i) b0 = fmaxf(b0, (c0 + i < kv) ? __uint_as_float(r[i]) : -INFINITY);
        }
        mx = fmaxf(mx, fmaxf(fmaxf(b0, b1), fmaxf(b2, b3)));
        PROBE_T(w1_m);
        continue;
      }
      // ---- main pass: un-normalised probabilities, their sum, dropout
      float v[CW];
      const bool drop = dr.thr != 0u;
      if (full) sum += drop ? fwd_probs<CW, true, true>(r, v, sc2, mx, c0, kv, tpart, sTab, dr.thr)
                            : fwd_probs<CW, true, false>(r, v, sc2, mx, c0, kv, tpart, sTab, dr.thr);
      else sum += drop ? fwd_probs<CW, false, true>(r, v, sc2, mx, c0, kv, tpart, sTab, dr.thr)
                       : fwd_probs<CW, false, false>(r, v, sc2, mx, c0, kv, tpart, sTab, dr.thr);
      uint32_t pk[CW / 2];
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      // the P buffer is free once the second contraction of block j-2 has completed
      PROBE_T(w2_m);
      mbar_wait(smem_u32(&pempty[pb]), ((j >> 1) & 1) ^ 1, err);
      PROBE_T(w2_pe);
      if constexpr (PT) {
        tc_fence_after();
        if constexpr (CW == 32) tmem_st16(lane_addr + (uint32_t)(F_TMEM_P + pb * 32 + ch * (CW / 2)), pk);
        else tmem_st8(lane_addr + (uint32_t)(F_TMEM_P + pb * 32 + ch * (CW / 2)), pk);
        tmem_wait_st();
        tc_fence_before();
      } else {
        uint8_t* prow = sP + pb * (AQ * 128) + row * 128;
#pragma unroll
        for (int i = 0; i < CW / 8; ++i) {
          const int chunk = ch * (CW / 8) + i;
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      mbar_arrive(smem_u32(&pfull[pb]));
      PROBE_T(w2_st);
    }
    if (prof) {
      p.dbg[0] = clock64() - t_all; p.dbg[1] = w1_sf; p.dbg[2] = w1_ld; p.dbg[3] = w1_m; p.dbg[4] = w2_sf; p.dbg[5] = w2_ld;
      p.dbg[6] = w2_m; p.dbg[7] = w2_pe; p.dbg[8] = w2_st;
    }
    // row sum over both column halves; O = acc / sum; lse = max + log2(sum)
    bf16* orow = p.out + (long long)(row_base + t) * p.D + h * HD + ch * (HD / NCH);
    if (nkb > 0) {
      xch[ch * AQ + row] = sum;
      asm volatile("bar.sync 1, %0;" ::"n"(128 * NCH) : "memory");
      float tot = 0.f;
#pragma unroll
      for (int c = 0; c < NCH; ++c) tot += xch[c * AQ + row];
      const float inv = (dr.thr != 0u ? dr.ks : 1.0f) / tot;
      if (p.lse && ch == 0 && row_valid) p.lse[(long long)bh * p.Tl + t] = mx + log2f(tot);
      if (mbar_wait(smem_u32(ofull), 0, err)) {
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HD / NCH / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(lane_addr + (uint32_t)(F_TMEM_O + ch * (HD / NCH) + c * 16), r);
          if (row_valid) {
            uint4* dst = reinterpret_cast<uint4*>(orow + c * 16);
#pragma unroll
            for (int i = 0; i < 2; ++i)
              dst[i] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * i]) * inv, __uint_as_float(r[8 * i + 1]) * inv),
                                  pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv),
                                  pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv),
                                  pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv));
          }
        }
      }
    } else if (row_valid) {
      for (int c = 0; c < HD / NCH; c += 8) *reinterpret_cast<uint4*>(orow + c) = make_uint4(0, 0, 0, 0);
      if (p.lse && ch == 0) p.lse[(long long)bh * p.Tl + t] = 0.f;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_free512(tmem_base);
}

// ===================================================================================================== backward: dQ
constexpr int BWD_NCH = 4;
constexpr int BWD_THREADS = 32 * (2 + 4 * BWD_NCH);             // 576
constexpr int DQ_KST = 3, DQ_VST = 2;
constexpr int DQ_SMEM = 2 * T128 + (DQ_KST + DQ_VST) * T64 + FA_TAB + 1024 + 512 + BWD_NCH * AQ * 4;
constexpr int A_TMEM_S = 0, A_TMEM_DP = 128, A_TMEM_DQ = 256, A_TMEM_DS = 448;   // 2 x 64, 2 x 64, 192, 2 x 32 columns

__device__ __forceinline__ void fa_dq_body(const CUtensorMap& tmQ, const CUtensorMap& tmdO, const CUtensorMap& tmKV,
                                           const FaParams& p, const int tile) {
  constexpr int NCH = BWD_NCH, KST = DQ_KST, VST = DQ_VST, CW = AK / NCH;      // 16 score columns per thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sdO = sQ + T128;
  uint8_t* sK = sdO + T128;
  uint8_t* sV = sK + KST * T64;
  uint32_t* sTab = (uint32_t*)(sV + VST * T64);      // dropout: one word per key of this (item, head)
  uint64_t* bars = (uint64_t*)(sTab + FA_MAXT);
  uint64_t* qfull = bars;
  uint64_t* kfull = qfull + 1;
  uint64_t* kempty = kfull + KST;
  uint64_t* vfull = kempty + KST;
  uint64_t* vempty = vfull + VST;
  uint64_t* sfull = vempty + VST;
  uint64_t* sempty = sfull + 2;
  uint64_t* pfull = sempty + 2;
  uint64_t* pempty = pfull + 2;
  uint64_t* ofull = pempty + 2;
  uint32_t* tmem_slot = (uint32_t*)(ofull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* err = p.err;
  const int nqb = (p.T + AQ - 1) / AQ;
  const int qb = tile % nqb;
  const int bh = tile / nqb;
  const int b = bh / p.H, h = bh % p.H;
  const int kv = fa_kv(p.lens, p.B, p.H, bh, p.plain_mask);
  const int nkb = (kv + AK - 1) / AK;
  const int q0 = qb * AQ;
  const int row_base = b * p.TP + FS2_PAD;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(qfull), 1);
    for (int s = 0; s < KST; ++s) { mbar_init(smem_u32(&kfull[s]), 1); mbar_init(smem_u32(&kempty[s]), 1); }
    for (int s = 0; s < VST; ++s) { mbar_init(smem_u32(&vfull[s]), 1); mbar_init(smem_u32(&vempty[s]), 1); }
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
  if (warp == 1) tmem_alloc512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // warp-uniform for the compiler: no per-MMA election loop
  pdl_wait();

  if (warp == 0) {
    if (nkb > 0 && elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmdO) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
      mbar_expect_tx(smem_u32(qfull), 2 * T128);
#pragma unroll
      for (int kc = 0; kc < 3; ++kc) {
        tma_load_4d(smem_u32(sQ + kc * (AQ * 128)), &tmQ, smem_u32(qfull), h * HD + kc * 64, row_base + q0, 0, 0);
        tma_load_4d(smem_u32(sdO + kc * (AQ * 128)), &tmdO, smem_u32(qfull), h * HD + kc * 64, row_base + q0, 0, 0);
      }
      int s = 0, sv = 0;
      uint32_t ph = 0, phv = 0;
      for (int j = 0; j < nkb; ++j) {
        if (!mbar_wait(smem_u32(&kempty[s]), ph ^ 1, err)) break;
        mbar_expect_tx(smem_u32(&kfull[s]), T64);
#pragma unroll
        for (int kc = 0; kc < 3; ++kc)
          tma_load_4d(smem_u32(sK + s * T64 + kc * (AK * 128)), &tmKV, smem_u32(&kfull[s]), p.D + h * HD + kc * 64,
                      row_base + j * AK, 0, 0);
        if (++s == KST) { s = 0; ph ^= 1; }
        if (!mbar_wait(smem_u32(&vempty[sv]), phv ^ 1, err)) break;
        mbar_expect_tx(smem_u32(&vfull[sv]), T64);
#pragma unroll
        for (int kc = 0; kc < 3; ++kc)
          tma_load_4d(smem_u32(sV + sv * T64 + kc * (AK * 128)), &tmKV, smem_u32(&vfull[sv]), 2 * p.D + h * HD + kc * 64,
                      row_base + j * AK, 0, 0);
        if (++sv == VST) { sv = 0; phv ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (nkb > 0 && elect_one()) {
      const uint32_t idescS = fa_idesc(AK, 0), idescO = fa_idesc(HD, 1);
      const uint64_t dQ0 = fa_base_k(smem_u32(sQ)), dO0 = fa_base_k(smem_u32(sdO)), dK0 = fa_base_k(smem_u32(sK));
      const uint64_t dV0 = fa_base_k(smem_u32(sV)), dKm0 = fa_base_mn(smem_u32(sK), AK);
      bool ok = mbar_wait(smem_u32(qfull), 0, err);
      int s = 0, sv = 0, sd = 0;                   // K stage of the score MMA, V stage, K stage of the dQ MMA
      uint32_t ph = 0, phv = 0;
      const bool prof = kProbe && p.dbg != nullptr && tile == 0;
      long long w_k = 0, w_v = 0, w_se = 0, w_pf = 0, t_all = clock64(), tt = t_all;
      for (int job = 0; job <= nkb && ok; ++job) {
        if (job < nkb) {
          const int sb = job & 1;
          if (prof) tt = clock64();
          if (!mbar_wait(smem_u32(&kfull[s]), ph, err)) break;
          PROBE_T(w_k);
          if (!mbar_wait(smem_u32(&vfull[sv]), phv, err)) break;
          PROBE_T(w_v);
          if (!mbar_wait(smem_u32(&sempty[sb]), ((job >> 1) & 1) ^ 1, err)) break;
          PROBE_T(w_se);
          tc_fence_after();
          const uint64_t dk = dK0 + (uint64_t)s * (T64 >> 4), dv = dV0 + (uint64_t)sv * (T64 >> 4);
          const uint32_t ts = tmem_base + A_TMEM_S + sb * AK, tdp = tmem_base + A_TMEM_DP + sb * AK;
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(ts, dQ0 + fa_off_k(AQ, k), dk + fa_off_k(AK, k), idescS, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(tdp, dO0 + fa_off_k(AQ, k), dv + fa_off_k(AK, k), idescS, k > 0 ? 1u : 0u);
          umma_commit(smem_u32(&vempty[sv]));
          umma_commit(smem_u32(&sfull[sb]));
          if (++s == KST) { s = 0; ph ^= 1; }
          if (++sv == VST) { sv = 0; phv ^= 1; }
        }
        const int jv = job - 1;
        if (jv >= 0) {
          const int pb = jv & 1;
          if (prof) tt = clock64();
          if (!mbar_wait(smem_u32(&pfull[pb]), (jv >> 1) & 1, err)) break;
          PROBE_T(w_pf);
          tc_fence_after();
          const uint64_t dkm = dKm0 + (uint64_t)sd * (T64 >> 4);
          const uint32_t tds = tmem_base + A_TMEM_DS + pb * 32;
#pragma unroll
          for (int k = 0; k < AK / 16; ++k)
            umma_bf16_ts(tmem_base + A_TMEM_DQ, tds + k * 8, dkm + fa_off_mn(k), idescO, (jv > 0 || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(&kempty[sd]));
          umma_commit(smem_u32(&pempty[pb]));
          if (++sd == KST) sd = 0;
          if (jv == nkb - 1) umma_commit(smem_u32(ofull));
        }
      }
      if (prof) { p.dbg[16] = clock64() - t_all; p.dbg[17] = w_k; p.dbg[18] = w_se; p.dbg[19] = w_pf; p.dbg[20] = w_v; p.dbg[21] = nkb; }
    }
  } else {
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int t = q0 + row;
    const bool row_valid = t < p.T;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const ADrop dr = adrop_make(p.drop_p, p.seed, p.seed_dev, bh);
    const uint32_t tpart = adrop_row(dr, t);
    if (dr.thr != 0u)          // visible to every softmax thread after the bar.sync of the row-term exchange below
      for (int c = (warp - 2) * 32 + lane; c < nkb * AK; c += 128 * NCH) sTab[c] = adrop_col(dr, c);
    const float sc2 = p.scale * LOG2E;
    const float scale = p.scale;
    // scale * rowsum(dO * O) of this row (the softmax-backward row term) comes from fa_rowdot_kernel
    const float ds_s = row_valid ? p.dvec[(long long)bh * p.Tl + t] : 0.f;
    if (dr.thr != 0u) asm volatile("bar.sync 1, %0;" ::"n"(128 * NCH) : "memory");       // the dropout word table is complete
    const float ks_s = (dr.thr != 0u ? dr.ks : 1.0f) * scale;
    // rows past T: lse = +inf makes every probability 2^-inf = 0 without a per-element test
    const float lse = row_valid ? p.lse[(long long)bh * p.Tl + t] : INFINITY;
    const bool prof = kProbe && p.dbg != nullptr && tile == 0 && warp == 2 && lane == 0;
    long long w_sf = 0, w_ld = 0, w_m = 0, w_pe = 0, w_st = 0, t_all = clock64(), tt = t_all;
    for (int j = 0; j < nkb; ++j) {
      const int sb = j & 1, pb = j & 1;
      const int c0 = j * AK + ch * CW;
      if (prof) tt = clock64();
      mbar_wait(smem_u32(&sfull[sb]), (j >> 1) & 1, err);
      PROBE_T(w_sf);
      tc_fence_after();
      uint32_t r[CW], g[CW];
      tmem_ld16_nw(lane_addr + (uint32_t)(A_TMEM_S + sb * AK + ch * CW), r);
      tmem_ld16_nw(lane_addr + (uint32_t)(A_TMEM_DP + sb * AK + ch * CW), g);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sempty[sb]));
      PROBE_T(w_ld);
      float v[CW];
      const bool full = (c0 + CW <= kv), drop = dr.thr != 0u;
      if (full) {
        if (drop) dq_ds<CW, true, true>(r, g, v, sc2, lse, ks_s, ds_s, c0, kv, tpart, sTab, dr.thr);
        else dq_ds<CW, true, false>(r, g, v, sc2, lse, ks_s, ds_s, c0, kv, tpart, sTab, dr.thr);
      } else {
        if (drop) dq_ds<CW, false, true>(r, g, v, sc2, lse, ks_s, ds_s, c0, kv, tpart, sTab, dr.thr);
        else dq_ds<CW, false, false>(r, g, v, sc2, lse, ks_s, ds_s, c0, kv, tpart, sTab, dr.thr);
      }
      uint32_t pk[CW / 2];
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
      PROBE_T(w_m);
      mbar_wait(smem_u32(&pempty[pb]), ((j >> 1) & 1) ^ 1, err);
      PROBE_T(w_pe);
      tc_fence_after();
      tmem_st8(lane_addr + (uint32_t)(A_TMEM_DS + pb * 32 + ch * (CW / 2)), pk);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(smem_u32(&pfull[pb]));
      PROBE_T(w_st);
    }
    if (prof) { p.dbg[0] = clock64() - t_all; p.dbg[1] = w_sf; p.dbg[2] = w_ld; p.dbg[3] = w_m; p.dbg[4] = w_pe; p.dbg[5] = w_st; }
    // dQ: TMEM -> bf16 -> columns [h*192, (h+1)*192) of the (rows, 3D) gradient; each thread writes 48 columns
    bf16* orow = p.out + (long long)(row_base + t) * (3LL * p.D) + h * HD + ch * (HD / NCH);
    if (nkb > 0) {
      if (mbar_wait(smem_u32(ofull), 0, err)) {
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HD / NCH / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(lane_addr + (uint32_t)(A_TMEM_DQ + ch * (HD / NCH) + c * 16), r);
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
  if (warp == 1) tmem_free512(tmem_base);
}

// ===================================================================================================== backward: dK, dV
constexpr int KV_QST = 4;
constexpr int KV_STAGE = 2 * T32 + 256;                          // Q block, dO block, 32 x (lse, D)
constexpr int KV_SMEM = 2 * T128 + KV_QST * 2 * T32 + KV_QST * 256 + FA_TAB + 1024 + 512;
constexpr int B_TMEM_ST = 0, B_TMEM_DPT = 64, B_TMEM_DV = 128, B_TMEM_DK = 320;   // 2 x 32, 2 x 32, 192, 192 columns

__device__ __forceinline__ void fa_dkv_body(const CUtensorMap& tmKV, const CUtensorMap& tmQ, const CUtensorMap& tmdO,
                                            const FaParams& p, const int tile) {
  constexpr int QST = KV_QST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;
  uint8_t* sV = sK + T128;
  uint8_t* sQ = sV + T128;                         // QST x [32 x 192]
  uint8_t* sdO = sQ + QST * T32;
  float* sStat = (float*)(sdO + QST * T32);        // QST x (32 lse, 32 D)
  uint32_t* sTab = (uint32_t*)(sStat + QST * 64);  // dropout: one word per query of this (item, head)
  uint64_t* bars = (uint64_t*)(sTab + FA_MAXT);
  uint64_t* kvfull = bars;
  uint64_t* qfull = kvfull + 1;                    // [QST]
  uint64_t* qempty = qfull + QST;
  uint64_t* sfull = qempty + QST;                  // [2]
  uint64_t* pfull = sfull + 2;
  uint64_t* bfree = pfull + 2;
  uint64_t* ofull = bfree + 2;
  uint32_t* tmem_slot = (uint32_t*)(ofull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* err = p.err;
  const int nkt = (p.T + AQ - 1) / AQ;
  const int kt = tile % nkt;
  const int bh = tile / nkt;
  const int b = bh / p.H, h = bh % p.H;
  const int kv = fa_kv(p.lens, p.B, p.H, bh, p.plain_mask);
  const int k0 = kt * AQ;
  const int nq = (k0 < kv) ? (p.T + BQ - 1) / BQ : 0;      // no valid key in this tile: dK = dV = 0
  const int row_base = b * p.TP + FS2_PAD;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(kvfull), 1);
    for (int s = 0; s < QST; ++s) { mbar_init(smem_u32(&qfull[s]), 1); mbar_init(smem_u32(&qempty[s]), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&sfull[s]), 1);
      mbar_init(smem_u32(&pfull[s]), 256);
      mbar_init(smem_u32(&bfree[s]), 1);
    }
    mbar_init(smem_u32(ofull), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) tmem_alloc512(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // warp-uniform for the compiler: no per-MMA election loop
  pdl_wait();

  if (warp == 0) {
    if (nq > 0 && elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmdO) : "memory");
      mbar_expect_tx(smem_u32(kvfull), 2 * T128);
#pragma unroll
      for (int kc = 0; kc < 3; ++kc) {
        tma_load_4d(smem_u32(sK + kc * (AQ * 128)), &tmKV, smem_u32(kvfull), p.D + h * HD + kc * 64, row_base + k0, 0, 0);
        tma_load_4d(smem_u32(sV + kc * (AQ * 128)), &tmKV, smem_u32(kvfull), 2 * p.D + h * HD + kc * 64, row_base + k0, 0, 0);
      }
      int s = 0;
      uint32_t ph = 0;
      const float* lse = p.lse + (long long)bh * p.Tl;
      const float* dvec = p.dvec + (long long)bh * p.Tl;
      for (int i = 0; i < nq; ++i) {
        if (!mbar_wait(smem_u32(&qempty[s]), ph ^ 1, err)) break;
        mbar_expect_tx(smem_u32(&qfull[s]), KV_STAGE);
#pragma unroll
        for (int kc = 0; kc < 3; ++kc) {
          tma_load_4d(smem_u32(sQ + s * T32 + kc * (BQ * 128)), &tmQ, smem_u32(&qfull[s]), h * HD + kc * 64, row_base + i * BQ, 0, 0);
          tma_load_4d(smem_u32(sdO + s * T32 + kc * (BQ * 128)), &tmdO, smem_u32(&qfull[s]), h * HD + kc * 64, row_base + i * BQ, 0, 0);
        }
        bulk_g2s(smem_u32(sStat + s * 64), lse + i * BQ, 128, smem_u32(&qfull[s]));
        bulk_g2s(smem_u32(sStat + s * 64 + 32), dvec + i * BQ, 128, smem_u32(&qfull[s]));
        if (++s == QST) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (nq > 0 && elect_one()) {
      const uint32_t idescS = fa_idesc(BQ, 0), idescO = fa_idesc(HD, 1);
      const uint64_t dK0 = fa_base_k(smem_u32(sK)), dV0 = fa_base_k(smem_u32(sV)), dQ0 = fa_base_k(smem_u32(sQ));
      const uint64_t dO0 = fa_base_k(smem_u32(sdO)), dQm0 = fa_base_mn(smem_u32(sQ), BQ), dOm0 = fa_base_mn(smem_u32(sdO), BQ);
      bool ok = mbar_wait(smem_u32(kvfull), 0, err);
      int s = 0, sd = 0;
      uint32_t ph = 0;
      const bool prof = kProbe && p.dbg != nullptr && tile == 0;
      long long w_q = 0, w_bf = 0, w_pf = 0, t_all = clock64(), tt = t_all;
      for (int job = 0; job <= nq && ok; ++job) {
        if (job < nq) {
          const int g = job & 1;
          if (prof) tt = clock64();
          if (!mbar_wait(smem_u32(&qfull[s]), ph, err)) break;
          PROBE_T(w_q);
          if (!mbar_wait(smem_u32(&bfree[g]), ((job >> 1) & 1) ^ 1, err)) break;
          PROBE_T(w_bf);
          tc_fence_after();
          const uint64_t dq = dQ0 + (uint64_t)s * (T32 >> 4), ddo = dO0 + (uint64_t)s * (T32 >> 4);
          const uint32_t tst = tmem_base + B_TMEM_ST + g * BQ, tdp = tmem_base + B_TMEM_DPT + g * BQ;
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(tst, dK0 + fa_off_k(AQ, k), dq + fa_off_k(BQ, k), idescS, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16(tdp, dV0 + fa_off_k(AQ, k), ddo + fa_off_k(BQ, k), idescS, k > 0 ? 1u : 0u);
          umma_commit(smem_u32(&sfull[g]));
          if (++s == QST) { s = 0; ph ^= 1; }
        }
        const int jv = job - 1;
        if (jv >= 0) {
          const int g = jv & 1;
          if (prof) tt = clock64();
          if (!mbar_wait(smem_u32(&pfull[g]), (jv >> 1) & 1, err)) break;
          PROBE_T(w_pf);
          tc_fence_after();
          const uint64_t dqm = dQm0 + (uint64_t)sd * (T32 >> 4), ddom = dOm0 + (uint64_t)sd * (T32 >> 4);
          const uint32_t tst = tmem_base + B_TMEM_ST + g * BQ, tdp = tmem_base + B_TMEM_DPT + g * BQ;
#pragma unroll
          for (int k = 0; k < BQ / 16; ++k)
            umma_bf16_ts(tmem_base + B_TMEM_DV, tst + k * 16, ddom + fa_off_mn(k), idescO, (jv > 0 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < BQ / 16; ++k)
            umma_bf16_ts(tmem_base + B_TMEM_DK, tdp + k * 16, dqm + fa_off_mn(k), idescO, (jv > 0 || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(&qempty[sd]));
          umma_commit(smem_u32(&bfree[g]));
          if (++sd == QST) sd = 0;
          if (jv == nq - 1) umma_commit(smem_u32(ofull));
        }
      }
      if (prof) { p.dbg[48] = clock64() - t_all; p.dbg[49] = w_q; p.dbg[50] = w_bf; p.dbg[51] = w_pf; p.dbg[52] = nq; }
    }
  } else {
    // two groups of 8 warps work the two score buffers in ping-pong; inside a group: lane quadrant x column half
    const int g = (warp - 2) >> 3;
    const int ch = ((warp - 2) >> 2) & 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int c = k0 + row;                        // key index of this thread's TMEM lane
    const bool key_valid = c < kv;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const ADrop dr = adrop_make(p.drop_p, p.seed, p.seed_dev, bh);
    const uint32_t cpart = adrop_col(dr, c);
    if (dr.thr != 0u && nq > 0) {
      for (int t = (warp - 2) * 32 + lane; t < nq * BQ; t += 512) sTab[t] = adrop_row(dr, t);
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }
    const float sc2 = p.scale * LOG2E;
    const float ks_s = (dr.thr != 0u ? dr.ks : 1.0f) * p.scale;
    const bool prof = kProbe && p.dbg != nullptr && tile == 0 && warp == 2 && lane == 0;
    long long w_sf = 0, w_ld = 0, w_m = 0, w_st = 0, t_all = clock64(), tt = t_all;
    for (int i = g; i < nq; i += 2) {
      const int s = i % QST;
      if (prof) tt = clock64();
      mbar_wait(smem_u32(&qfull[s]), (i / QST) & 1, err);      // the block's (lse, D)
      mbar_wait(smem_u32(&sfull[g]), (i >> 1) & 1, err);
      PROBE_T(w_sf);
      tc_fence_after();
      uint32_t r[16], d[16];
      tmem_ld16_nw(lane_addr + (uint32_t)(B_TMEM_ST + g * BQ + ch * 16), r);
      tmem_ld16_nw(lane_addr + (uint32_t)(B_TMEM_DPT + g * BQ + ch * 16), d);
      tmem_wait_ld();
      PROBE_T(w_ld);
      const float* st_l = sStat + s * 64 + ch * 16;
      const float* st_d = sStat + s * 64 + 32 + ch * 16;
      const int t0 = i * BQ + ch * 16;
      const bool full = (i * BQ + BQ <= p.T), drop = dr.thr != 0u;        // full: every query of the block exists
      uint32_t pp[8], ps[8];
      if (key_valid) {
        if (full) {
          if (drop) dkv_pt<true, true>(r, d, st_l, st_d, pp, ps, sc2, ks_s, t0, p.T, cpart, sTab, dr.thr);
          else dkv_pt<true, false>(r, d, st_l, st_d, pp, ps, sc2, ks_s, t0, p.T, cpart, sTab, dr.thr);
        } else {
          if (drop) dkv_pt<false, true>(r, d, st_l, st_d, pp, ps, sc2, ks_s, t0, p.T, cpart, sTab, dr.thr);
          else dkv_pt<false, false>(r, d, st_l, st_d, pp, ps, sc2, ks_s, t0, p.T, cpart, sTab, dr.thr);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) pp[e] = ps[e] = 0u;
      }
      // P^T / dS^T overwrite the first half of this thread's own score columns (already in registers)
      PROBE_T(w_m);
      tmem_st8(lane_addr + (uint32_t)(B_TMEM_ST + g * BQ + ch * 16), pp);
      tmem_st8(lane_addr + (uint32_t)(B_TMEM_DPT + g * BQ + ch * 16), ps);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(smem_u32(&pfull[g]));
      PROBE_T(w_st);
    }
    if (prof) { p.dbg[32] = clock64() - t_all; p.dbg[33] = w_sf; p.dbg[34] = w_ld; p.dbg[35] = w_m; p.dbg[36] = w_st; }
    // [dV | dK] = 384 accumulator columns per key row, 96 per thread -> columns [2D + h*192, ..) and [D + h*192, ..) of dqkv
    const int w4 = g * 2 + ch;
    const bool is_dv = w4 < 2;
    const int coff = (w4 & 1) * 96;
    bf16* orow = p.out + (long long)(row_base + c) * (3LL * p.D) + (is_dv ? 2 * p.D : p.D) + h * HD + coff;
    const bool wr = c < p.T;
    const float osc = (is_dv && dr.thr != 0u) ? dr.ks : 1.0f;      // dV = (keep-or-zero P)^T dO / (1 - p)
    if (nq > 0) {
      if (mbar_wait(smem_u32(ofull), 0, err)) {
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 6; ++cc) {
          uint32_t r[16];
          tmem_ld16(lane_addr + (uint32_t)((is_dv ? B_TMEM_DV : B_TMEM_DK) + coff + cc * 16), r);
          if (wr) {
            uint4* dst = reinterpret_cast<uint4*>(orow + cc * 16);
#pragma unroll
            for (int i = 0; i < 2; ++i)
              dst[i] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * i]) * osc, __uint_as_float(r[8 * i + 1]) * osc),
                                  pack_bf16x2(__uint_as_float(r[8 * i + 2]) * osc, __uint_as_float(r[8 * i + 3]) * osc),
                                  pack_bf16x2(__uint_as_float(r[8 * i + 4]) * osc, __uint_as_float(r[8 * i + 5]) * osc),
                                  pack_bf16x2(__uint_as_float(r[8 * i + 6]) * osc, __uint_as_float(r[8 * i + 7]) * osc));
          }
        }
      }
    } else if (wr) {
      for (int cc = 0; cc < 96; cc += 8) *reinterpret_cast<uint4*>(orow + cc) = make_uint4(0, 0, 0, 0);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_free512(tmem_base);
}

// One launch for the whole attention backward: the first half of the grid are dK/dV tiles (the longer ones first), the second
// half dQ tiles.  Two separate launches of 448 CTAs each end in two nearly empty fourth waves on 148 SMs (3.03 waves at
// B = 32, T = 800); 896 CTAs of one launch fill the machine until the very end.  The row term D = scale * rowsum(dO * O) both
// kinds of tile need comes from a small pre-pass (fa_rowdot_kernel), so the tiles are independent of each other.
__global__ void __launch_bounds__(BWD_THREADS, 1) fa_bwd_kernel(const __grid_constant__ CUtensorMap tmQ128,
                                                                const __grid_constant__ CUtensorMap tmdO128,
                                                                const __grid_constant__ CUtensorMap tmKV64,
                                                                const __grid_constant__ CUtensorMap tmQ32,
                                                                const __grid_constant__ CUtensorMap tmdO32, const FaParams p) {
  const int n_tiles = (p.T + AQ - 1) / AQ * p.B * p.H;
  if ((int)blockIdx.x < n_tiles) fa_dkv_body(tmQ128, tmQ32, tmdO32, p, (int)blockIdx.x);
  else fa_dq_body(tmQ128, tmdO128, tmKV64, p, (int)blockIdx.x - n_tiles);
}

// dvec[bh][t] = scale * sum_c dO[t, c] * O[t, c] over the head's 192 columns: one warp per (row, head)
__global__ void __launch_bounds__(256) fa_rowdot_kernel(const FaParams p) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long long n = (long long)p.B * p.H * p.T;
  for (long long i = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += (long long)gridDim.x * 8) {
    const int t = (int)(i % p.T);
    const int bh = (int)(i / p.T);
    const int b = bh / p.H, h = bh % p.H;
    const long long o = ((long long)b * p.TP + FS2_PAD + t) * p.D + h * HD + lane * 2;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < HD / 64; ++k) {
      const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.dO + o + 64 * k));
      const float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p.O + o + 64 * k));
      acc += x.x * y.x + x.y * y.y;
    }
    acc = warp_sum(acc);
    if (lane == 0) p.dvec[(long long)bh * p.Tl + t] = acc * p.scale;
  }
}

// the keep mask of the attention dropout, for tests (uint8 [B*H, T, T])
__global__ void fa_mask_kernel(int BH, int T, float drop_p, unsigned long long seed, const unsigned long long* seed_dev,
                               unsigned char* keep) {
  const long long n = (long long)BH * T * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % T), t = (int)((i / T) % T), bh = (int)(i / ((long long)T * T));
    const ADrop dr = adrop_make(drop_p, seed, seed_dev, bh);
    keep[i] = (dr.thr == 0u || adrop_keep(adrop_row(dr, t), adrop_col(dr, c), dr.thr)) ? 1 : 0;
  }
}

int g_fa_p_in_tmem = 1;

int fa_common(const void* qkv, const int* lens, int B, int H, int T, int D, FaParams& p) {
  if (!qkv || !lens || B <= 0 || H <= 0 || T <= 0 || D != H * HD) {
    fs2_set_error("fs2_flash_attn: bad arguments (needs head_dim 192, bf16)");
    return FS2_ERR_ARG;
  }
  memset(&p, 0, sizeof p);
  p.B = B; p.H = H; p.T = T; p.TP = T + 2 * FS2_PAD; p.D = D;
  p.Tl = (T + AQ - 1) / AQ * AQ;
  p.lens = lens;
  p.dbg = g_fa_dbg;
  return fs2_tc_error_ptr(&p.err);
}

template <typename K>
int fa_smem_attr(K kern, int bytes) {
  CUDA_CHECK_RET(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return FS2_OK;
}

}  // namespace

/* measurement hook: cycle breakdown of CTA 0 of every flash-attention launch (32 slots; library built with -DFS2_TC_PROBE) */
extern "C" int fs2_flash_attn_set_debug(long long* dev_buf) {
  g_fa_dbg = dev_buf;
  return FS2_OK;
}

extern "C" int fs2_flash_attn_tune(int p_in_tmem) {
  g_fa_p_in_tmem = p_in_tmem;
  return FS2_OK;
}

extern "C" int fs2_flash_attn_lse_len(int T) { return (T + AQ - 1) / AQ * AQ; }

extern "C" int fs2_flash_attn_fwd(const void* qkv, const int* lens, int B, int H, int T, int D, float scale, float drop_p,
                                  unsigned long long seed, const unsigned long long* seed_dev, float* lse, void* O,
                                  int plain_mask, void* stream) {
  FaParams p;
  int rc = fa_common(qkv, lens, B, H, T, D, p);
  if (rc) return rc;
  if (!O) { fs2_set_error("fs2_flash_attn_fwd: null pointer"); return FS2_ERR_ARG; }
  CUtensorMap tq, tkv;
  rc = fs2_tc_make_map_2d(qkv, 3LL * D, (long long)B * p.TP, 3LL * D, 64, AQ, &tq);
  if (rc) return rc;
  rc = fs2_tc_make_map_2d(qkv, 3LL * D, (long long)B * p.TP, 3LL * D, 64, AK, &tkv);
  if (rc) return rc;
  if (drop_p > 0.f && T > FA_MAXT - AQ) {
    fs2_set_error("fs2_flash_attn: attention dropout is supported up to 1408 rows per item");
    return FS2_ERR_UNSUPPORTED;
  }
  p.scale = scale; p.drop_p = drop_p; p.seed = seed; p.seed_dev = seed_dev;
  p.lse = lse;
  p.out = (bf16*)O;
  p.plain_mask = plain_mask;
  const int grid = (T + AQ - 1) / AQ * B * H;
  if (g_fa_p_in_tmem) {
    static bool cfg = false;
    if (!cfg) { rc = fa_smem_attr(fa_fwd_kernel<true>, FwdCfg<true>::SMEM); if (rc) return rc; cfg = true; }
    FS2_LAUNCH(fa_fwd_kernel<true>, grid, FWD_THREADS, FwdCfg<true>::SMEM, (cudaStream_t)stream, tq, tkv, p);
  } else {
    static bool cfg = false;
    if (!cfg) { rc = fa_smem_attr(fa_fwd_kernel<false>, FwdCfg<false>::SMEM); if (rc) return rc; cfg = true; }
    FS2_LAUNCH(fa_fwd_kernel<false>, grid, FWD_THREADS, FwdCfg<false>::SMEM, (cudaStream_t)stream, tq, tkv, p);
  }
  return fs2_check_launch();
}

extern "C" int fs2_flash_attn_bwd(const void* dO, const void* O, const void* qkv, const float* lse, const int* lens, int B,
                                  int H, int T, int D, float scale, float drop_p, unsigned long long seed,
                                  const unsigned long long* seed_dev, float* dvec, void* dqkv, int plain_mask, void* stream) {
  FaParams p;
  int rc = fa_common(qkv, lens, B, H, T, D, p);
  if (rc) return rc;
  if (!dO || !O || !lse || !dvec || !dqkv) { fs2_set_error("fs2_flash_attn_bwd: null pointer"); return FS2_ERR_ARG; }
  const long long prow = (long long)B * p.TP;
  CUtensorMap tq128, tdo128, tkv64, tq32, tdo32;
  if ((rc = fs2_tc_make_map_2d(qkv, 3LL * D, prow, 3LL * D, 64, AQ, &tq128))) return rc;
  if ((rc = fs2_tc_make_map_2d(dO, D, prow, D, 64, AQ, &tdo128))) return rc;
  if ((rc = fs2_tc_make_map_2d(qkv, 3LL * D, prow, 3LL * D, 64, AK, &tkv64))) return rc;
  if ((rc = fs2_tc_make_map_2d(qkv, 3LL * D, prow, 3LL * D, 64, BQ, &tq32))) return rc;
  if ((rc = fs2_tc_make_map_2d(dO, D, prow, D, 64, BQ, &tdo32))) return rc;
  if (drop_p > 0.f && T > FA_MAXT - AQ) {
    fs2_set_error("fs2_flash_attn: attention dropout is supported up to 1408 rows per item");
    return FS2_ERR_UNSUPPORTED;
  }
  p.scale = scale; p.drop_p = drop_p; p.seed = seed; p.seed_dev = seed_dev;
  p.lse = const_cast<float*>(lse);
  p.dvec = dvec;
  p.out = (bf16*)dqkv;
  p.O = (const bf16*)O;
  p.dO = (const bf16*)dO;
  p.plain_mask = plain_mask;
  constexpr int BWD_SMEM = DQ_SMEM > KV_SMEM ? DQ_SMEM : KV_SMEM;
  static bool cfg = false;
  if (!cfg) {
    if ((rc = fa_smem_attr(fa_bwd_kernel, BWD_SMEM))) return rc;
    cfg = true;
  }
  const int tiles = (T + AQ - 1) / AQ * B * H;
  const long long rows = (long long)B * H * T;
  FS2_LAUNCH(fa_rowdot_kernel, (unsigned)((rows + 7) / 8 < 4096 ? (rows + 7) / 8 : 4096), 256, 0, (cudaStream_t)stream, p);
  rc = fs2_check_launch();
  if (rc) return rc;
  FS2_LAUNCH(fa_bwd_kernel, 2 * tiles, BWD_THREADS, BWD_SMEM, (cudaStream_t)stream, tq128, tdo128, tkv64, tq32, tdo32, p);
  return fs2_check_launch();
}

extern "C" int fs2_flash_attn_mask(int BH, int T, float drop_p, unsigned long long seed, const unsigned long long* seed_dev,
                                   unsigned char* keep, void* stream) {
  if (!keep || BH <= 0 || T <= 0) { fs2_set_error("fs2_flash_attn_mask: bad arguments"); return FS2_ERR_ARG; }
  fa_mask_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>(BH, T, drop_p, seed, seed_dev, keep);
  return fs2_check_launch();
}
