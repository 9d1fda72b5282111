// fs2_gemm_ln_tc: GEMM + bias + branch dropout + residual + LayerNorm in one tcgen05 kernel.
//
// The post-norm FFT block of the reference (speechbrain TransformerEncoderLayer, reached from model.py:344-347 and
// 425-428) runs  y = LN(x + dropout(a . W^T + b))  twice per layer: out-projection -> norm1 and FFN conv 2 -> norm2.  As
// two launches the fp32 branch (a . W^T + b) makes a round trip through HBM: 8 of the 18 bytes per element the pair moves.
// Here one CTA owns a 128 x 384 tile, i.e. 128 COMPLETE LayerNorm rows, in tensor memory (384 of the 512 columns, fp32):
//   warp 0   : TMA producer   (A 128 x 64 and W 384 x 64 bf16 boxes, 128B swizzle, 3-stage ring of 64 KB)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (two N = 192 MMAs per K = 16 step)
//   warps 2-9: epilogue.  Each warp owns 16 rows (its TMEM lane quadrant, upper or lower half) and reads them in the
//              16x256b register layout: a quad of lanes holds 32 contiguous bytes of a row, so residual loads and output
//              stores are full-sector accesses straight from / to registers, and row statistics are quad shuffles.
//              pass 1: z = x + keep * (acc + bias) written back to tensor memory, row sums;  pass 2: centred squares;
//              pass 3: normalise, scale / shift, store fp32 + bf16 (+ reflect-halo mirror rows).
// While the MMAs of a tile run, the epilogue warps ask L2 for the tile's residual rows; the accumulator is handed back to
// the MMA thread after the last tensor-memory read of pass 3, so the next tile's MMAs overlap the final stores.
// The accumulator cannot be double-buffered (2 x 384 columns > 512): what overlaps a tile's epilogue is the TMA ring
// (the next tile's first 3 k-blocks) and the other SMs' main loops.
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include "common.cuh"
#include "../../include/fs2_b200.h"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int LN_N = 384;
constexpr int BNS = 192;                      // UMMA N (<= 256): two MMAs cover the row
constexpr int STAGES = 3;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NTHREADS = 32 * (2 + NUM_EPI_WARPS);
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int BSUB_BYTES = BNS * BK * 2;      // 24 KB
constexpr int STAGE_BYTES = A_BYTES + 2 * BSUB_BYTES;     // 64 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
constexpr int NCHUNK = LN_N / 32;             // 32-column chunks per row
constexpr int CH = 2;                         // chunks per epilogue step
static_assert(NCHUNK % CH == 0, "chunks per step");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct LnGemmParams {
  Fs2GemmLn g;
  int M, m_tiles, kb;
  int* err;
};

// what the epilogue needs to know about one output row
struct RowInfo {
  long long off;     // element offset of the row in the (M, 384) tensors
  long long m1, m2;  // element deltas of the reflect-halo mirror rows (0 = none)
  int r;             // padded row index
  bool ok;           // a real (b, t) row: gets LayerNorm output
  bool zero;         // a halo row no mirror reaches: gets zeros
};

__device__ __forceinline__ RowInfo row_info(int r, int M, int T, int halo) {
  RowInfo ri;
  ri.r = r;
  ri.off = (long long)r * LN_N;
  ri.m1 = ri.m2 = 0;
  ri.ok = ri.zero = false;
  if (r < M) {
    const int TP = T + 2 * FS2_PAD;
    const int b = r / TP, t = r - b * TP - FS2_PAD;
    if (t >= 0 && t < T) {
      ri.ok = true;
      if (halo > 0) {
        if (t >= 1 && t <= halo) ri.m1 = -2LL * t * LN_N;
        if (t >= T - 1 - halo && t <= T - 2) ri.m2 = 2LL * (T - 1 - t) * LN_N;
      }
    } else {
      const int d = (t < 0) ? -t : (t - (T - 1));
      ri.zero = !(halo > 0 && d <= halo && d < T);
    }
  }
  return ri;
}

__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

__global__ void __launch_bounds__(NTHREADS, 1) gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const LnGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;      // accumulator ready
  uint64_t* tempty_bar = tfull_bar + 1;          // accumulator drained
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Fs2GemmLn& g = p.g;
  int* err = p.err;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tfull_bar), 1);
    mbar_init(smem_u32(tempty_bar), NUM_EPI_WARPS);
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      const uint32_t smem0 = smem_u32(smem);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t s = 0, ph = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x) {
        const int m0 = t * BM;
        for (int kb = 0; kb < p.kb; ++kb) {
          if (!mbar_wait(empty0 + 8 * s, ph ^ 1, err)) { ok = false; break; }
          const uint32_t fb = full0 + 8 * s;
          mbar_expect_tx(fb, STAGE_BYTES);
          const uint32_t sa = smem0 + s * STAGE_BYTES;
          tma_load_4d(sa, &tmA, fb, kb * BK, m0, 0, 0);
          tma_load_4d(sa + A_BYTES, &tmB, fb, kb * BK, 0, 0, 0);
          tma_load_4d(sa + A_BYTES + BSUB_BYTES, &tmB, fb, kb * BK, BNS, 0, 0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (elect_one()) {
      // instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNS >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t adesc0 = smem_desc(smem0, 16, 1024);
      const uint64_t bdesc0 = smem_desc(smem0 + A_BYTES, 16, 1024);
      constexpr uint64_t KSTEP = 32 >> 4;                 // 16 bf16 = 32 B inside the 128 B swizzle row
      constexpr uint64_t STAGE_STEP = STAGE_BYTES >> 4;
      constexpr uint64_t SUB_STEP = BSUB_BYTES >> 4;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t s = 0, ph = 0, tc = 0;
      bool ok = true;
      for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x, ++tc) {
        if (!mbar_wait(smem_u32(tempty_bar), (tc & 1) ^ 1, err)) { ok = false; break; }
        tc_fence_after();
        uint32_t acc = 0;
        for (int kb = 0; kb < p.kb; ++kb) {
          if (!mbar_wait(full0 + 8 * s, ph, err)) { ok = false; break; }
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)s * STAGE_STEP, bd = bdesc0 + (uint64_t)s * STAGE_STEP;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16(tmem_base, ad + k * KSTEP, bd + k * KSTEP, idesc, acc);
            umma_bf16(tmem_base + BNS, ad + k * KSTEP, bd + SUB_STEP + k * KSTEP, idesc, acc);
            acc = 1;
          }
          umma_commit(empty0 + 8 * s);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (ok) umma_commit(smem_u32(tfull_bar));
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue: 16 full rows per warp
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    const int hf = (warp - 2) >> 2;              // lower / upper 16 lanes of the quadrant
    const int tr = lane >> 2, tq = lane & 3;
    const uint32_t tw = tmem_base + ((uint32_t)(q * 32 + hf * 16) << 16);
    const float invC = 1.0f / (float)LN_N;
    bf16* oa = (bf16*)g.out_act;
    // branch dropout: the mask of fs2_ln_fwd / fs2_ln_bwd (common.cuh:drop_scale4) -- one 64-bit mix per group of four
    // consecutive elements.  A lane holds two consecutive columns of rows tr and tr + 8; the even lane of a pair mixes the
    // group of row tr, the odd lane the group of row tr + 8, and one shuffle hands each the half it lacks.
    const bool drop = g.drop_p > 0.f;
    const uint64_t dseed = g.drop_seed ^ (g.seed_dev ? mix64(*g.seed_dev) : 0ull);
    const uint32_t dthr = (uint32_t)(g.drop_p * 65536.0f);
    const float dks = drop ? 1.0f / (1.0f - g.drop_p) : 1.0f;
    uint32_t tc = 0;
    bool ok = true;
    for (int t = blockIdx.x; t < p.m_tiles && ok; t += gridDim.x, ++tc) {
      const int row0 = t * BM + q * 32 + hf * 16;
      const RowInfo ra = row_info(row0 + tr, p.M, g.T, g.halo), rb = row_info(row0 + tr + 8, p.M, g.T, g.halo);
      // residual rows of this warp (16 x 1536 B, contiguous): into L2 while the tile's MMAs run
      {
        const char* xb = reinterpret_cast<const char*>(g.x + (long long)row0 * LN_N);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int line = i * 32 + lane;
          if (row0 + line / 12 < p.M) asm volatile("prefetch.global.L2 [%0];" ::"l"(xb + line * 128));
        }
      }
      const float* xa = g.x + ra.off + 2 * tq;
      const float* xb_ = g.x + rb.off + 2 * tq;
      // CH chunks (32 columns each) per step: CH tensor-memory loads in flight before one wait, residual loads one step ahead
      float2 xc[8 * CH], xn[8 * CH];
#pragma unroll
      for (int i = 0; i < 4 * CH; ++i) {
        xc[2 * i] = ra.ok ? ldg2(xa + 8 * i) : make_float2(0.f, 0.f);
        xc[2 * i + 1] = rb.ok ? ldg2(xb_ + 8 * i) : make_float2(0.f, 0.f);
      }
      if (!mbar_wait(smem_u32(tfull_bar), tc & 1, err)) { ok = false; break; }
      tc_fence_after();
      // ---- pass 1: z = x + keep * (acc + bias) -> tensor memory; row sums
      float sa = 0.f, sb = 0.f;
      const uint64_t ga0 = (uint64_t)((tq & 1) ? rb.r : ra.r) * (LN_N / 4) + (uint64_t)(tq >> 1);
#pragma unroll 1
      for (int c = 0; c < NCHUNK; c += CH) {
        uint32_t r[CH][16];
#pragma unroll
        for (int j = 0; j < CH; ++j) tmem_ld_16x256b_x4(tw + (uint32_t)((c + j) * 32), r[j]);
        if (c + CH < NCHUNK) {
#pragma unroll
          for (int i = 0; i < 4 * CH; ++i) {
            xn[2 * i] = ra.ok ? ldg2(xa + (c + CH) * 32 + 8 * i) : make_float2(0.f, 0.f);
            xn[2 * i + 1] = rb.ok ? ldg2(xb_ + (c + CH) * 32 + 8 * i) : make_float2(0.f, 0.f);
          }
        }
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < CH; ++j) {
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const float2 b2 = ldg2(g.bias + (c + j) * 32 + 8 * n + 2 * tq);
            float ka0 = 1.f, ka1 = 1.f, kb0 = 1.f, kb1 = 1.f;
            if (drop) {
              const uint64_t grp = ga0 + (uint64_t)((c + j) * 8 + 2 * n);
              const uint64_t rnd = mix64(dseed ^ (grp * 0xD6E8FEB86659FD93ull));
              const uint32_t mine_lo = (uint32_t)rnd, mine_hi = (uint32_t)(rnd >> 32);
              const uint32_t got = __shfl_xor_sync(0xffffffffu, (tq & 1) ? mine_lo : mine_hi, 1);
              const uint32_t bits_a = (tq & 1) ? got : mine_lo;       // row tr:     the even lane owns the group's mix
              const uint32_t bits_b = (tq & 1) ? mine_hi : got;       // row tr + 8: the odd lane owns it
              ka0 = ((bits_a & 0xFFFFu) >= dthr) ? dks : 0.f;
              ka1 = ((bits_a >> 16) >= dthr) ? dks : 0.f;
              kb0 = ((bits_b & 0xFFFFu) >= dthr) ? dks : 0.f;
              kb1 = ((bits_b >> 16) >= dthr) ? dks : 0.f;
            }
            const float2 xa2 = xc[2 * (4 * j + n)], xb2 = xc[2 * (4 * j + n) + 1];
            const float za0 = xa2.x + (__uint_as_float(r[j][4 * n]) + b2.x) * ka0;
            const float za1 = xa2.y + (__uint_as_float(r[j][4 * n + 1]) + b2.y) * ka1;
            const float zb0 = xb2.x + (__uint_as_float(r[j][4 * n + 2]) + b2.x) * kb0;
            const float zb1 = xb2.y + (__uint_as_float(r[j][4 * n + 3]) + b2.y) * kb1;
            sa += za0 + za1;
            sb += zb0 + zb1;
            r[j][4 * n] = __float_as_uint(za0);
            r[j][4 * n + 1] = __float_as_uint(za1);
            r[j][4 * n + 2] = __float_as_uint(zb0);
            r[j][4 * n + 3] = __float_as_uint(zb1);
          }
          tmem_st_16x256b_x4(tw + (uint32_t)((c + j) * 32), r[j]);
        }
#pragma unroll
        for (int i = 0; i < 8 * CH; ++i) xc[i] = xn[i];
      }
      tmem_wait_st();
      sa += __shfl_xor_sync(0xffffffffu, sa, 1);
      sa += __shfl_xor_sync(0xffffffffu, sa, 2);
      sb += __shfl_xor_sync(0xffffffffu, sb, 1);
      sb += __shfl_xor_sync(0xffffffffu, sb, 2);
      const float mean_a = sa * invC, mean_b = sb * invC;
      // ---- pass 2: centred sum of squares
      float qa = 0.f, qb = 0.f;
#pragma unroll 1
      for (int c = 0; c < NCHUNK; c += CH) {
        uint32_t r[CH][16];
#pragma unroll
        for (int j = 0; j < CH; ++j) tmem_ld_16x256b_x4(tw + (uint32_t)((c + j) * 32), r[j]);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < CH; ++j) {
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const float a0 = __uint_as_float(r[j][4 * n]) - mean_a, a1 = __uint_as_float(r[j][4 * n + 1]) - mean_a;
            const float b0 = __uint_as_float(r[j][4 * n + 2]) - mean_b, b1 = __uint_as_float(r[j][4 * n + 3]) - mean_b;
            qa += a0 * a0 + a1 * a1;
            qb += b0 * b0 + b1 * b1;
          }
        }
      }
      qa += __shfl_xor_sync(0xffffffffu, qa, 1);
      qa += __shfl_xor_sync(0xffffffffu, qa, 2);
      qb += __shfl_xor_sync(0xffffffffu, qb, 1);
      qb += __shfl_xor_sync(0xffffffffu, qb, 2);
      const float rstd_a = rsqrtf(qa * invC + g.eps), rstd_b = rsqrtf(qb * invC + g.eps);
      if (tq == 0) {
        if (ra.ok) {
          if (g.mean) g.mean[ra.r] = mean_a;
          if (g.rstd) g.rstd[ra.r] = rstd_a;
        }
        if (rb.ok) {
          if (g.mean) g.mean[rb.r] = mean_b;
          if (g.rstd) g.rstd[rb.r] = rstd_b;
        }
      }
      // ---- pass 3: normalise, scale / shift, store
      const bool wa = ra.ok || ra.zero, wb = rb.ok || rb.zero;
      const bool mirrors = __any_sync(0xffffffffu, (ra.m1 | ra.m2 | rb.m1 | rb.m2) != 0);
      float* fa = g.out_f32 ? g.out_f32 + ra.off + 2 * tq : nullptr;
      float* fb = g.out_f32 ? g.out_f32 + rb.off + 2 * tq : nullptr;
      bf16* ha = oa ? oa + ra.off + 2 * tq : nullptr;
      bf16* hb = oa ? oa + rb.off + 2 * tq : nullptr;
#pragma unroll 1
      for (int c = 0; c < NCHUNK; c += CH) {
        uint32_t r[CH][16];
#pragma unroll
        for (int j = 0; j < CH; ++j) tmem_ld_16x256b_x4(tw + (uint32_t)((c + j) * 32), r[j]);
        tmem_wait_ld();
        if (c + CH >= NCHUNK) {
          // last tensor-memory read of this warp: hand the accumulator back to the MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(tempty_bar));
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int col = (c + j) * 32 + 8 * n;
            const float2 g2 = ldg2(g.gamma + col + 2 * tq), be2 = ldg2(g.beta + col + 2 * tq);
            float2 ya, yb;
            ya.x = (__uint_as_float(r[j][4 * n]) - mean_a) * rstd_a * g2.x + be2.x;
            ya.y = (__uint_as_float(r[j][4 * n + 1]) - mean_a) * rstd_a * g2.y + be2.y;
            yb.x = (__uint_as_float(r[j][4 * n + 2]) - mean_b) * rstd_b * g2.x + be2.x;
            yb.y = (__uint_as_float(r[j][4 * n + 3]) - mean_b) * rstd_b * g2.y + be2.y;
            if (!ra.ok) ya = make_float2(0.f, 0.f);
            if (!rb.ok) yb = make_float2(0.f, 0.f);
            const __nv_bfloat162 pa = __floats2bfloat162_rn(ya.x, ya.y), pb = __floats2bfloat162_rn(yb.x, yb.y);
            if (wa) {
              if (fa) *reinterpret_cast<float2*>(fa + col) = ya;
              if (ha) *reinterpret_cast<__nv_bfloat162*>(ha + col) = pa;
            }
            if (wb) {
              if (fb) *reinterpret_cast<float2*>(fb + col) = yb;
              if (hb) *reinterpret_cast<__nv_bfloat162*>(hb + col) = pb;
            }
            if (mirrors) {                       // warp-uniform: only the first / last rows of an item (halo > 0)
              if (ra.m1) {
                if (fa) *reinterpret_cast<float2*>(fa + ra.m1 + col) = ya;
                if (ha) *reinterpret_cast<__nv_bfloat162*>(ha + ra.m1 + col) = pa;
              }
              if (ra.m2) {
                if (fa) *reinterpret_cast<float2*>(fa + ra.m2 + col) = ya;
                if (ha) *reinterpret_cast<__nv_bfloat162*>(ha + ra.m2 + col) = pa;
              }
              if (rb.m1) {
                if (fb) *reinterpret_cast<float2*>(fb + rb.m1 + col) = yb;
                if (hb) *reinterpret_cast<__nv_bfloat162*>(hb + rb.m1 + col) = pb;
              }
              if (rb.m2) {
                if (fb) *reinterpret_cast<float2*>(fb + rb.m2 + col) = yb;
                if (hb) *reinterpret_cast<__nv_bfloat162*>(hb + rb.m2 + col) = pb;
              }
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int g_ln_gemm_sms = 0;

}  // namespace

#define REQUIRE(cond, msg) do { if (!(cond)) { fs2_set_error(msg); return FS2_ERR_ARG; } } while (0)

extern "C" int fs2_gemm_ln_tc(const Fs2GemmLn* gp, void* stream) {
  REQUIRE(gp && gp->A && gp->W && gp->bias && gp->x && gp->gamma && gp->beta, "fs2_gemm_ln_tc: null pointer");
  REQUIRE(gp->out_f32 || gp->out_act, "fs2_gemm_ln_tc: no output");
  const Fs2GemmLn& g = *gp;
  REQUIRE(g.B > 0 && g.T > 0 && g.K > 0 && g.K % 8 == 0, "fs2_gemm_ln_tc: bad shape (K must be a multiple of 8)");
  REQUIRE(g.halo >= 0 && g.halo <= FS2_PAD && (g.halo == 0 || g.T > g.halo), "fs2_gemm_ln_tc: halo too wide for T");
  REQUIRE(g.drop_p >= 0.f && g.drop_p < 1.f, "fs2_gemm_ln_tc: bad dropout probability");
  REQUIRE((((uintptr_t)g.x | (uintptr_t)g.out_f32 | (uintptr_t)g.bias | (uintptr_t)g.gamma | (uintptr_t)g.beta) % 8) == 0 &&
              ((uintptr_t)g.out_act % 4) == 0,
          "fs2_gemm_ln_tc: fp32 operands must be 8-byte aligned");
  const long long Ml = (long long)g.B * (g.T + 2 * FS2_PAD);
  REQUIRE(Ml * LN_N < (1LL << 40) && Ml < 0x7FFFFFFF, "fs2_gemm_ln_tc: too many rows");
  LnGemmParams p;
  memset(&p, 0, sizeof p);
  p.g = g;
  p.M = (int)Ml;
  p.m_tiles = (p.M + BM - 1) / BM;
  p.kb = (g.K + BK - 1) / BK;
  int rc = fs2_tc_error_ptr(&p.err);
  if (rc) return rc;
  CUtensorMap ta, tb;
  if ((rc = fs2_tc_make_map_2d(g.A, g.K, p.M, g.lda, BK, BM, &ta))) return rc;
  if ((rc = fs2_tc_make_map_2d(g.W, g.K, LN_N, g.ldw, BK, BNS, &tb))) return rc;
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(gemm_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int dev = 0;
    CUDA_CHECK_RET(cudaGetDevice(&dev));
    CUDA_CHECK_RET(cudaDeviceGetAttribute(&g_ln_gemm_sms, cudaDevAttrMultiProcessorCount, dev));
    configured = true;
  }
  const int grid = p.m_tiles < g_ln_gemm_sms ? p.m_tiles : g_ln_gemm_sms;
  FS2_LAUNCH(gemm_ln_kernel, grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream, ta, tb, p);
  return fs2_check_launch();
}
