// tcgen05 / TMEM / TMA GEMM for the Fs2Gemm descriptor (bf16 operands, fp32 accumulate).
//
// One CTA computes one 128 x BN output tile:
//   warp 0   : TMA producer   (cp.async.bulk.tensor 4-D tiles, 128B swizzle, mbarrier tx)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16)
//   warps 2-5: epilogue       (tcgen05.ld TMEM -> registers -> bias/ReLU/mask -> global)
// A STAGES-deep smem ring decouples TMA from the tensor pipe; two CTAs are co-resident per
// SM so one tile's epilogue overlaps the neighbour's main loop.
//
// Operand majorness (mode 0: K/K, mode 1: K/MN, mode 2: MN/MN) is expressed purely in the
// shared-memory matrix descriptors + the instruction descriptor, so forward convs, dgrads and
// wgrads all read activations and weights in their natural layouts (no transposes in HBM).
// Conv taps are extra K-blocks whose TMA row coordinate is shifted (implicit GEMM over the
// reflect-padded row space, see common.cuh).
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <string>
#include "common.cuh"
#include "../../include/fs2_b200.h"
#include "gemm_epilogue.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int NUM_EPI_WARPS = 8;
constexpr int NTHREADS = 32 * (2 + NUM_EPI_WARPS);

// Measurement-only code (role-level cycle probes of tools/gemm_rounds.py, the FS2_TC_EPIT=3 direct-store epilogue) is
// compiled in with `make EXTRA=-DFS2_TC_PROBE`; in the product build it is dead code and is removed by the compiler --
// the GEMM kernels are sensitive to their instruction footprint (1-3 % of the training step).
#ifdef FS2_TC_PROBE
constexpr bool kTcProbe = true;
#else
constexpr bool kTcProbe = false;
#endif

struct TcParams {
  Fs2Gemm g;
  int kb_per_tap;     // ceil(K / 64)
  int total_kb;       // K-blocks in the whole reduction
  int kb_per_split;
  int nsplit;
  int ntiles_per_tap; // mode 2
  int m_tiles, n_tiles_total, total_tiles;
  int pa[4], pb[4];   // tensor-map dim slot of (inner,row,i1,i2) for A and B
  int* err;
  long long* dbg;     // optional: {SM cycles, nanoseconds} of unit 0's lifetime (clock probe for measurements)
  int dbg_mode;       // measurements only (results are garbage): 1 = MMA issue without TMA, 2 = TMA without MMA, 3 = no epilogue stores
  int epi_transpose;  // 1: coalesced epilogue stores through the per-warp smem staging tile (epi_chunk_t)
};

__device__ int g_tc_error = 0;

__device__ __forceinline__ void tma_issue(uint32_t dst, const CUtensorMap* tm, uint32_t bar, const int* perm, int inner,
                                          int row, int i1, int i2) {
  const int p1 = perm[1], p2 = perm[2];
  const int c1 = (p1 == 1) ? row : (p2 == 1) ? i1 : i2;
  const int c2 = (p1 == 2) ? row : (p2 == 2) ? i1 : i2;
  const int c3 = (p1 == 3) ? row : (p2 == 3) ? i1 : i2;
  tma_load_4d(dst, tm, bar, inner, c1, c2, c3);
}
// ------------------------------------------------------------------ CTA-pair (cta_group::2) helpers --
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_count_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t cluster_bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
template <int NCTA>
__device__ __forceinline__ void tma_issue_x(uint32_t dst, const CUtensorMap* tm, uint32_t bar, const int* perm, int inner,
                                            int row, int i1, int i2) {
  // perm[0] is always 0 (the contiguous dimension); select the rest without a dynamically indexed (local-memory) array
  const int p1 = perm[1], p2 = perm[2];
  const int c1 = (p1 == 1) ? row : (p2 == 1) ? i1 : i2;
  const int c2 = (p1 == 2) ? row : (p2 == 2) ? i1 : i2;
  const int c3 = (p1 == 3) ? row : (p2 == 3) ? i1 : i2;
  if (NCTA == 2) tma_load_4d_pair(dst, tm, bar, inner, c1, c2, c3);
  else tma_load_4d(dst, tm, bar, inner, c1, c2, c3);
}
template <int NCTA>
__device__ __forceinline__ void umma_bf16_x(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  if (NCTA == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
  }
}
// MMA-completion arrive; in pair mode the same barrier offset is signalled in both CTAs
template <int NCTA>
__device__ __forceinline__ void umma_commit_x(uint32_t bar) {
  if (NCTA == 2) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
  } else {
    umma_commit(bar);
  }
}

// One 32-column accumulator chunk of one output row: scale/bias/ReLU/masks, then vector or scalar stores
// (halo mirror rows included).  r[] holds the fp32 accumulators of columns [nb0, nb0+32).
// Row-per-thread scalar stores: the complete epilogue semantics (bias, ReLU, ReLU-mask, row mask, halo mirrors, bf16 /
// fp32 / atomics, strided columns).  Used for partial chunks, strided wgrad outputs (k > 1) and the few warp tiles that
// contain halo-mirror rows; everything else goes through epi_chunk_t.  (A row-per-thread VECTOR path used to live here
// as well; it is gone because its code, inlined next to the other two paths, cost more in instruction footprint than it
// saved on the rare tiles that still took it.)
template <int MODE>
__device__ __forceinline__ void epi_chunk(const Fs2Gemm& g, const EpiRow& er, const uint32_t* r, int nb0, long long col0,
                                          long long cstr, bool atomic, uint8_t* stage) {
  // Forward / dgrad kernels reach this path only for partial chunks and halo-mirror tiles: the 32 values of the row go
  // through the warp's staging tile (word i of lane l at (i*32 + l)*4: conflict-free) so that ONE epi_store body in a
  // rolled loop serves them all -- 32 inlined copies of it were the largest single piece of code in those kernels.
  if constexpr (MODE == 2) {
    // weight gradients with k > 1 live on this path: fp32 atomics into strided columns and nothing else (no bias, ReLU,
    // mask, mirror) -- 32 bare atomics from registers instead of 32 copies of the general store
    if (atomic && !g.c_bf16 && !g.bias && !g.relu && !g.relu_aux && g.rs_Tp == 0) {
      float* dst = (float*)g.C + er.base + col0;
      const float alpha = g.alpha;
      if (nb0 + 32 <= g.N) {
#pragma unroll
        for (int i = 0; i < 32; ++i) atomicAdd(dst + i * cstr, __uint_as_float(r[i]) * alpha);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (nb0 + i < g.N) atomicAdd(dst + i * cstr, __uint_as_float(r[i]) * alpha);
      }
      return;
    }
  }
  const uint32_t sb = smem_u32(stage) + (uint32_t)(threadIdx.x & 31) * 4u;
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(sb + (uint32_t)i * 128u), "r"(half ? r[16 + i] : r[i]) : "memory");
#pragma unroll 1
    for (int i = 0; i < 16; ++i) {
      const int n = nb0 + half * 16 + i;
      if (n >= g.N) break;
      uint32_t v;
      asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(sb + (uint32_t)i * 128u) : "memory");
      epi_store(g, er, col0 + (long long)(half * 16 + i) * cstr, n, __uint_as_float(v), atomic);
    }
  }
}

// Coalesced variant of the vector store path.  tcgen05.ld hands every lane one accumulator ROW, so a direct st.v4 per
// lane touches 32 different 128-byte lines per instruction; for the short-K GEMMs (projections, 1x1 convs) the L1
// tag stage then bounds the kernel (4-6 us of epilogue per 128 x N tile against ~1.5 us of MMA).  Here each warp bounces
// its 32 x 64-byte slab through a private, XOR-swizzled 2 KB shared-memory tile and writes it back transposed: four
// lanes cover 64 contiguous bytes of one row, one instruction covers 8 rows -- 4x fewer line visits, and the ReLU-mask
// operand (relu_aux, bf16) is fetched with the same coalesced pattern.  Preconditions (warp-uniform, checked by the
// caller): full 32-column chunk, unit column stride, no atomics, no halo mirrors in this warp's rows.
constexpr int EPI_STAGE_BYTES = 2048;                      // per epilogue warp
constexpr int EPI_STAGE_TOTAL = NUM_EPI_WARPS * EPI_STAGE_BYTES;
// Weight-gradient kernels (mode 2) with 128 x 192 tiles can fold the bias gradient in (Fs2Gemm.a_colsum): a 2 KB tile of
// bf16 ones serves as a second B operand -- one extra N = 16 MMA per K step gives the column sums of the A tile (dy) in 16
// spare tensor-memory columns per accumulator buffer.  All elements are equal, so layout and swizzle of the tile do not matter.
constexpr int ONES_OFF = 256 + EPI_STAGE_TOTAL + 768;      // after the barriers and the staging tiles, 1024-aligned
constexpr int ONES_BYTES = 2048;
constexpr int CS_COLS = 16;

__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}

// per-tile data of the transposed store path: computed once per tile, used by every 32-column chunk of the tile
struct EpiT {
  long long tb[4];      // C offsets of rows tr, tr+8, tr+16, tr+24 of this warp's 32 rows (tr = lane / 4)
  unsigned okmask;      // rows that exist and are written
  unsigned livemask;    // ... and are not masked (t < lens[b])
  bool any_dead;        // some written row is masked (t >= lens[b]): the value path has to select zeros
  bool mirrors;         // some row of this warp also writes a reflect-halo mirror copy (m1 / m2 != 0)
  long long m1[4], m2[4];   // mirror offsets of rows tr, tr+8, tr+16, tr+24 (valid when `mirrors`)
};

__device__ __forceinline__ void epi_t_setup(const EpiRow& er, bool row_ok, EpiT& et) {
  const int lane = threadIdx.x & 31;
  const bool mine = row_ok && !er.skip;
  et.okmask = __ballot_sync(0xffffffffu, mine);
  et.livemask = __ballot_sync(0xffffffffu, mine && er.live);
  et.any_dead = et.livemask != et.okmask;
  et.mirrors = __ballot_sync(0xffffffffu, row_ok && (er.mirror != 0 || er.mirror2 != 0)) != 0u;
  const long long mybase = mine ? er.base : 0;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) et.tb[jj] = __shfl_sync(0xffffffffu, mybase, jj * 8 + (lane >> 2));
  if (et.mirrors) {                       // warp-uniform; only GEMMs whose output feeds a k > 1 conv (halo > 0)
    const long long a = mine ? er.mirror : 0, b = mine ? er.mirror2 : 0;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      et.m1[jj] = __shfl_sync(0xffffffffu, a, jj * 8 + (lane >> 2));
      et.m2[jj] = __shfl_sync(0xffffffffu, b, jj * 8 + (lane >> 2));
    }
  }
}

// fp32 output straight from the 16x256b TMEM register layout (tmem_ld_16x256b_x4, two loads = 32 rows x 32 columns):
// lane (tr = lane/4, tq = lane%4) holds rows tr, tr+8, tr+16, tr+24 and, per 8-column group n, columns 8n + 2tq + {0,1}.
// One 8-byte store per (row, group): four lanes fill one 32-byte sector, one instruction covers 8 rows -- the same
// sector efficiency as the smem-transposed path without its st.shared / ld.shared / __syncwarp chain.
template <int MODE>
__device__ __forceinline__ void epi_chunk_q(const Fs2Gemm& g, const EpiT& et, const uint32_t* rlo, const uint32_t* rhi,
                                            int nb0, long long col0, bool atomic, bool no_stg) {
  const int lane = threadIdx.x & 31;
  const int tr = lane >> 2, tq = lane & 3;
  const unsigned okmask = no_stg ? 0u : et.okmask;
#pragma unroll
  for (int n = 0; n < 4; ++n) {
    const int col = 8 * n + 2 * tq;
    float2 b2 = make_float2(0.f, 0.f);
    if (g.bias) b2 = __ldg(reinterpret_cast<const float2*>(g.bias + nb0 + col));
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const uint32_t* r = (jj & 2) ? rhi : rlo;
      float x0 = __uint_as_float(r[4 * n + 2 * (jj & 1)]), x1 = __uint_as_float(r[4 * n + 2 * (jj & 1) + 1]);
      if (g.alpha != 1.f) { x0 *= g.alpha; x1 *= g.alpha; }
      x0 += b2.x;
      x1 += b2.y;
      if (g.relu) { x0 = epi_act(x0, g.relu); x1 = epi_act(x1, g.relu); }
      const int row = jj * 8 + tr;
      if (et.any_dead && !((et.livemask >> row) & 1u)) { x0 = 0.f; x1 = 0.f; }
      if ((okmask >> row) & 1u) {
        float2* dst = reinterpret_cast<float2*>((float*)g.C + et.tb[jj] + col0 + col);
        if (atomic) atomicAdd(dst, make_float2(x0, x1));
        else *dst = make_float2(x0, x1);
      }
    }
  }
}

template <int MODE>
__device__ __forceinline__ void epi_chunk_t(const Fs2Gemm& g, const EpiRow& er, bool row_ok, const EpiT& et, const uint32_t* r,
                                            int nb0, long long col0, bool out_bf16, uint8_t* stage, bool no_stg = false,
                                            bool atomic = false) {
  const int lane = threadIdx.x & 31;
  const unsigned okmask = no_stg ? 0u : et.okmask;
  const int tr = lane >> 2, tcq = lane & 3;                // transposed role: row tr (+8 per step), 16-byte quarter tcq
  const long long* tb = et.tb;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  if (g.alpha != 1.f) {                                    // every branch below is warp-uniform
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= g.alpha;
  }
  if (g.bias) {
    const float4* bp = reinterpret_cast<const float4*>(g.bias + nb0);     // nb0 % 32 == 0, cudaMalloc-aligned base
    if ((reinterpret_cast<uintptr_t>(bp) & 15) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 b4 = __ldg(bp + i);
        v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += g.bias[nb0 + i];
    }
  }
  if (g.relu) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (et.any_dead) {
    const bool live = row_ok && !er.skip && er.live;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = live ? v[i] : 0.f;
  }
  const uint32_t sbase = smem_u32(stage);
  const uint32_t wr = sbase + (uint32_t)lane * 64u, wsw = (uint32_t)((lane >> 1) & 3);
  if (out_bf16) {
    uint4 aux[4];
    const bool has_aux = g.relu_aux != nullptr;
    if (has_aux) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        aux[jj] = make_uint4(0, 0, 0, 0);
        if ((okmask >> (jj * 8 + tr)) & 1u)
          aux[jj] = *reinterpret_cast<const uint4*>((const bf16*)g.relu_aux + tb[jj] + col0 + tcq * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * k], v[8 * k + 1]);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * k + 2], v[8 * k + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * k + 4], v[8 * k + 5]);
      __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * k + 6], v[8 * k + 7]);
      uint4 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2);
      pk.w = *reinterpret_cast<uint32_t*>(&h3);
      sts128(wr + (((uint32_t)k ^ wsw) << 4), pk.x, pk.y, pk.z, pk.w);
    }
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int row = jj * 8 + tr;
      uint4 x = lds128(sbase + row * 64 + (((uint32_t)tcq ^ (uint32_t)((row >> 1) & 3)) << 4));
      if ((okmask >> row) & 1u) {
        if (has_aux) {
          // keep a value where the forward activation was > 0: bf16 sign clear and magnitude non-zero
          const uint32_t a[4] = {aux[jj].x, aux[jj].y, aux[jj].z, aux[jj].w};
          uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t lo = ((a[i] & 0x8000u) == 0 && (a[i] & 0x7FFFu) != 0) ? 0x0000FFFFu : 0u;
            const uint32_t hi = ((a[i] & 0x80000000u) == 0 && (a[i] & 0x7FFF0000u) != 0) ? 0xFFFF0000u : 0u;
            w[i] &= (lo | hi);
          }
          x = make_uint4(w[0], w[1], w[2], w[3]);
        }
        bf16* dst = (bf16*)g.C + tb[jj] + col0 + tcq * 8;
        *reinterpret_cast<uint4*>(dst) = x;
        if (et.mirrors) {
          if (et.m1[jj]) *reinterpret_cast<uint4*>(dst + et.m1[jj]) = x;
          if (et.m2[jj]) *reinterpret_cast<uint4*>(dst + et.m2[jj]) = x;
        }
      }
    }
    __syncwarp();
  } else {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        sts128(wr + (((uint32_t)k ^ wsw) << 4), __float_as_uint(v[half * 16 + 4 * k]), __float_as_uint(v[half * 16 + 4 * k + 1]),
               __float_as_uint(v[half * 16 + 4 * k + 2]), __float_as_uint(v[half * 16 + 4 * k + 3]));
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int row = jj * 8 + tr;
        const uint4 x = lds128(sbase + row * 64 + (((uint32_t)tcq ^ (uint32_t)((row >> 1) & 3)) << 4));
        if ((okmask >> row) & 1u) {
          float* dst = (float*)g.C + tb[jj] + col0 + half * 16 + tcq * 4;
          const float4 xv = make_float4(__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w));
          // split-K / accumulate: one 16-byte vector reduction per lane, 8 rows x 64 B per instruction -- 8x fewer L2
          // atomic operations than the row-per-thread scalar atomics (the k = 1 weight gradients were bound by them)
          if (atomic) atomicAdd(reinterpret_cast<float4*>(dst), xv);
          else {
            st4(dst, xv);
            if (et.mirrors) {
              if (et.m1[jj]) st4(dst + et.m1[jj], xv);
              if (et.m2[jj]) st4(dst + et.m2[jj], xv);
            }
          }
        }
      }
      __syncwarp();
    }
  }
}

template <int MODE, int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1) tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const TcParams p) {
  constexpr int A_BYTES = BM * BK * 2;          // 16 KB
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr bool A_MN = (MODE == 2);
  constexpr bool B_MN = (MODE != 0);
  constexpr int TMEM_COLS = (2 * BN <= 256) ? 256 : 512;     // two accumulator buffers
  constexpr int CHUNKS_PER_HALF = BN / 64;                   // 32-column chunks per epilogue column-half

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Fs2Gemm& g = p.g;
  int* err = p.err;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), NUM_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // bias-gradient fold (see ONES_OFF): possible where two accumulator buffers leave 2 x 16 tensor-memory columns free
  constexpr bool CS_OK = (MODE == 2) && (2 * BN + 2 * CS_COLS <= TMEM_COLS);
  const bool cs_on = CS_OK && g.a_colsum != nullptr;
  if (cs_on) {
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + ONES_OFF);
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NTHREADS) ones[i] = 0x3F803F80u;      // bf16 1.0 pairs
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);     // warp-uniform for the compiler
  pdl_wait();      // prologue above overlaps the previous kernel's tail; global memory is touched only below

  // Persistent tile loop: tile t -> (n-tile fastest so CTAs running together share the A rows in L2).
  // Every role walks the same sequence; smem stage / phase counters run across tiles.
  if (warp == 0) {
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      // The issue loops of this thread and of the MMA thread are single-thread instruction streams: every k-block must
      // cost them fewer cycles than its MMAs take (4 x 48 cycles at BN = 192), so stage / phase / tap counters are
      // incremental (no division, no modulo) and the tensor-map coordinate permutation is resolved once per kernel.
      bool ok = true;
      const bool prof = kTcProbe && p.dbg != nullptr && blockIdx.x == 0;
      long long w_empty = 0, t_all = prof ? clock64() : 0;
      const int ars = p.pa[1], ai1 = p.pa[2];        // slot (1..3) of the row coordinate / of i1 in A's tensor map
      const int brs = p.pb[1], bi1 = p.pb[2];
      const uint32_t smem0 = smem_u32(smem);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t s = 0, ph = 0;
      for (int t = blockIdx.x; t < p.total_tiles && ok; t += gridDim.x) {
        const int nt = t % p.n_tiles_total;
        const int rest = t / p.n_tiles_total;
        const int m0 = (rest % p.m_tiles) * BM;
        const int z = rest / p.m_tiles;
        const int zb = z / p.nsplit, zs = z % p.nsplit;
        const int i1 = zb % g.batch1, i2 = zb / g.batch1;
        int tapN = 0, n0 = nt * BN;
        if (MODE == 2) { tapN = nt / p.ntiles_per_tap; n0 = (nt % p.ntiles_per_tap) * BN; }
        const int kb_begin = zs * p.kb_per_split;
        const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
        // coordinates of the batch dims in their slots; the row slot is patched per load with three selects
        const int a1 = (ai1 == 1) ? i1 : i2, a2 = (ai1 == 2) ? i1 : i2, a3 = (ai1 == 3) ? i1 : i2;
        const int b1 = (bi1 == 1) ? i1 : i2, b2 = (bi1 == 2) ? i1 : i2, b3 = (bi1 == 3) ? i1 : i2;
        int j = (MODE == 2) ? 0 : kb_begin / p.kb_per_tap;            // tap of the current k-block (modes 0/1)
        int kk = (MODE == 2) ? kb_begin * BK : (kb_begin - j * p.kb_per_tap) * BK;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          const long long tq = prof ? clock64() : 0;
          if (!mbar_wait(empty0 + 8 * s, ph ^ 1, err)) { ok = false; break; }
          if (prof) w_empty += clock64() - tq;
          const uint32_t fb = full0 + 8 * s;
          mbar_expect_tx(fb, STAGE_BYTES);
          const uint32_t sa = smem0 + s * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
          if (MODE == 0 || MODE == 1) {
            const int arow = g.a_row_off + m0 + j * g.a_tap_step;
            tma_load_4d(sa, &tmA, fb, kk, ars == 1 ? arow : a1, ars == 2 ? arow : a2, ars == 3 ? arow : a3);
            if (MODE == 0) {
              tma_load_4d(sb, &tmB, fb, j * g.b_tap_step + kk, brs == 1 ? n0 : b1, brs == 2 ? n0 : b2, brs == 3 ? n0 : b3);
            } else {
              const int brow = g.b_row_off + kk;
#pragma unroll
              for (int i = 0; i < BN / 64; ++i)
                tma_load_4d(sb + i * (BK * 128), &tmB, fb, n0 + 64 * i + j * g.b_tap_step, brs == 1 ? brow : b1,
                            brs == 2 ? brow : b2, brs == 3 ? brow : b3);
            }
            kk += BK;
            if (kk >= p.kb_per_tap * BK) { kk = 0; ++j; }
          } else {
            const int arow = g.a_row_off + kk, brow = g.b_row_off + kk + tapN * g.b_tap_step;
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_4d(sa + i * (BK * 128), &tmA, fb, m0 + 64 * i, ars == 1 ? arow : a1, ars == 2 ? arow : a2,
                          ars == 3 ? arow : a3);
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_4d(sb + i * (BK * 128), &tmB, fb, n0 + 64 * i, brs == 1 ? brow : b1, brs == 2 ? brow : b2,
                          brs == 3 ? brow : b3);
            kk += BK;
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
      if (prof) { p.dbg[2] = clock64() - t_all; p.dbg[3] = w_empty; }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // instruction descriptor: D=f32, A=B=bf16, majorness, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t tc = 0;
      bool ok = true;
      const bool prof = kTcProbe && p.dbg != nullptr && blockIdx.x == 0;
      long long w_acc = 0, w_full = 0, t_all = prof ? clock64() : 0;
      // shared-memory descriptors: everything but the 14-bit start-address field is constant, and that field only ever
      // moves by (bytes >> 4) -- one descriptor per operand for stage 0, then plain 64-bit adds
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t adesc0 = A_MN ? smem_desc(smem0, BK * 128, 1024) : smem_desc(smem0, 16, 1024);
      const uint64_t bdesc0 = B_MN ? smem_desc(smem0 + A_BYTES, BK * 128, 1024) : smem_desc(smem0 + A_BYTES, 16, 1024);
      constexpr uint64_t A_STEP = A_MN ? (2048 >> 4) : (32 >> 4);      // one K = 16 step inside the stage
      constexpr uint64_t B_STEP = B_MN ? (2048 >> 4) : (32 >> 4);
      constexpr uint64_t STAGE_STEP = STAGE_BYTES >> 4;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t s = 0, ph = 0;
      // bias-gradient fold: the tile of ones as a second (MN-major) B operand, N = 16
      const uint64_t ones_desc = smem_desc(smem0 + STAGES * STAGE_BYTES + ONES_OFF, BK * 128, 1024);
      const uint32_t idesc_cs = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(CS_COLS >> 3) << 17);
      for (int t = blockIdx.x; t < p.total_tiles && ok; t += gridDim.x, ++tc) {
        const int z = (t / p.n_tiles_total) / p.m_tiles;
        const int zs = z % p.nsplit;
        const int kb_begin = zs * p.kb_per_split;
        const int nkb = min(p.total_kb, kb_begin + p.kb_per_split) - kb_begin;
        const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
        const long long tq0 = prof ? clock64() : 0;
        if (!mbar_wait(smem_u32(&tempty_bar[as]), aph ^ 1, err)) { ok = false; break; }
        if (prof) w_acc += clock64() - tq0;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        // bias-gradient fold: the column tiles of a row block (n_tiles_total of them walk the same A tiles) share the work --
        // tile nt sums the k-blocks with (kb + nt) % n_tiles_total == 0, every tile adds its partial sums in the epilogue.
        // (All of it in the first column tile made that tile twice as long as its neighbours: 182 -> 208 us for the k = 9
        // FFN weight gradient; a countdown instead of a modulo keeps the issue loop short.)
        const uint32_t tmem_cs = tmem_base + 2 * BN + as * CS_COLS;
        int cs_ctr = 0x7FFFFFFF;
        if (CS_OK && cs_on) {
            const int ntt = p.n_tiles_total;
            cs_ctr = (ntt - (kb_begin + t % ntt) % ntt) % ntt;
        }
        uint32_t acc_cs = 0;
        uint32_t acc = 0;
        for (int i = 0; i < nkb; ++i) {
          const long long tq1 = prof ? clock64() : 0;
          if (!mbar_wait(full0 + 8 * s, ph, err)) { ok = false; break; }
          if (prof) w_full += clock64() - tq1;
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)s * STAGE_STEP, bd = bdesc0 + (uint64_t)s * STAGE_STEP;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 elements = 32 B inside the 128 B swizzle row; 8-row groups 1024 B apart.
            // MN-major: 16 k-rows = 2 groups of 8 rows (1024 B each); 64-wide MN blocks 8192 B apart.
            umma_bf16(tmem_d, ad + k * A_STEP, bd + k * B_STEP, idesc, acc);
            acc = 1;
          }
          if (CS_OK && cs_ctr-- == 0) {
            // column sums of the same A tile (one branch per k-block, outside the back-to-back MMA sequence above: the
            // issue loop of this thread is the kernel's critical path)
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_cs, ad + k * A_STEP, ones_desc, idesc_cs, (k != 0) ? 1u : acc_cs);
            acc_cs = 1;
            cs_ctr = p.n_tiles_total - 1;
          }
          umma_commit(empty0 + 8 * s);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (ok) umma_commit(smem_u32(&tfull_bar[as]));
      }
      if (prof) { p.dbg[4] = clock64() - t_all; p.dbg[5] = w_acc; p.dbg[6] = w_full; p.dbg[7] = tc; }
    }
  } else {
    // ---------------- epilogue: 8 warps; TMEM lane quadrant = warp % 4, column half = (warp-2)/4 ----------------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* stage = smem + STAGES * STAGE_BYTES + 256 + (warp - 2) * EPI_STAGE_BYTES;
    const bool atomic = (p.nsplit > 1 && g.c_split_stride == 0) || g.accumulate;
    const long long cstr = (MODE == 2 && g.c_col_stride > 1) ? g.c_col_stride : 1;
    const bool al4 = ((g.ldc | g.c_col_off | g.c_s1 | g.c_s2) % 4 == 0) && (((uintptr_t)g.C) % 16 == 0);
    const bool al8 = ((g.ldc | g.c_col_off | g.c_s1 | g.c_s2) % 8 == 0) && (((uintptr_t)g.C) % 16 == 0);
    const bool aux_al = g.relu_aux == nullptr || (((uintptr_t)g.relu_aux) % 16 == 0);
    uint32_t tc = 0;
    bool ok = true;
    const bool prof = kTcProbe && p.dbg != nullptr && blockIdx.x == 0 && warp == 2;
    long long w_tfull = 0, t_chunks = 0, t_ld = 0, t_all = prof ? clock64() : 0;
    for (int t = blockIdx.x; t < p.total_tiles && ok; t += gridDim.x, ++tc) {
      const int nt = t % p.n_tiles_total;
      const int rest = t / p.n_tiles_total;
      const int m0 = (rest % p.m_tiles) * BM;
      const int z = rest / p.m_tiles;
      const int zb = z / p.nsplit;
      const int i1 = zb % g.batch1, i2 = zb / g.batch1;
      int tapN = 0, n0 = nt * BN;
      if (MODE == 2) { tapN = nt / p.ntiles_per_tap; n0 = (nt % p.ntiles_per_tap) * BN; }
      const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
      const int m = m0 + q * 32 + lane;
      EpiRow er;
      const bool row_ok = (m < g.M);
      if (row_ok) {
        epi_row_setup(g, i1, i2, m, er);
        er.base += (long long)(z % p.nsplit) * g.c_split_stride;       // 0 unless deterministic split-K
      }
      const long long colbase = (MODE == 2) ? (long long)tapN * g.c_tap_stride + n0 * cstr : n0;
      const bool vec_f32 = cstr == 1 && !g.c_bf16 && !atomic && al4 && (colbase % 4 == 0) && aux_al;
      const bool vec_bf16 = cstr == 1 && g.c_bf16 && al8 && (colbase % 8 == 0) && aux_al;
      // coalesced (smem-transposed) store path: decided and set up once per tile
      EpiT et;
      const bool vec_f32_at = cstr == 1 && !g.c_bf16 && atomic && al4 && (colbase % 4 == 0) && g.relu_aux == nullptr &&
                              g.bias == nullptr && !g.relu;
      bool t_path = p.epi_transpose && (vec_f32_at || (vec_f32 && g.relu_aux == nullptr) ||
                                        (vec_bf16 && (g.relu_aux == nullptr || g.aux_bf16)));
      if (t_path) {
        epi_t_setup(er, row_ok, et);
        t_path = !(et.mirrors && atomic);
      }
      // experiment switch FS2_TC_EPIT=3: fp32 outputs skip the smem transpose (direct 8-byte stores from the 16x256b TMEM
      // layout).  Measured equal-to-slower than the transpose (profiles/r01_summary.md section 3: in steady state these
      // GEMMs are bound by the HBM traffic of their fp32 output, not by the store instruction pattern), so it is off.
      const bool q_path = kTcProbe && t_path && p.epi_transpose == 3 && (vec_f32_at || (vec_f32 && g.relu_aux == nullptr)) &&
                          (g.bias == nullptr || (((uintptr_t)g.bias) % 8 == 0));
      if constexpr (MODE == 1) {
        // ReLU-mask operand of this tile (the forward activation, written long ago): the epilogue warps get here while the
        // tile's MMAs are still running, so pull this lane's row segment into L2 now instead of paying the HBM latency
        // inside the chunk loop
        if (g.relu_aux && g.aux_bf16 && row_ok && !er.skip) {
          const char* ap = reinterpret_cast<const char*>((const bf16*)g.relu_aux + er.base + colbase + half * (CHUNKS_PER_HALF * 32));
#pragma unroll
          for (int i = 0; i < CHUNKS_PER_HALF * 64; i += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(ap + i));
        }
      }
      const long long tq2 = prof ? clock64() : 0;
      if (!mbar_wait(smem_u32(&tfull_bar[as]), aph, err)) { ok = false; break; }
      const long long tq3 = prof ? clock64() : 0;
      if (prof) w_tfull += tq3 - tq2;
      tc_fence_after();
      bool cs_have = false;
      if (CS_OK && cs_on && half == 0) {
        // did this tile's share of the k-blocks ((kb + nt) % n_tiles_total == 0) contain any?
        const int zs_ = z % p.nsplit, kb0_ = zs_ * p.kb_per_split, kb1_ = min(p.total_kb, kb0_ + p.kb_per_split);
        const int ntt = p.n_tiles_total;
        cs_have = kb0_ + (ntt - (kb0_ + nt) % ntt) % ntt < kb1_;
      }
      if (CS_OK && cs_have) {
        // partial column sums of this row block's A tiles: every thread holds the sum of its row m
        uint32_t cs[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * BN + as * CS_COLS), cs);
        if (row_ok) atomicAdd(g.a_colsum + m, __uint_as_float(cs[0]) * g.alpha);
      }
#pragma unroll 1
      for (int ci = 0; ci < CHUNKS_PER_HALF; ++ci) {
        const int c = half * CHUNKS_PER_HALF + ci;
        uint32_t r[32];
        const int nb0 = n0 + c * 32;
        const bool q_use = q_path && nb0 + 32 <= g.N && !(kTcProbe && p.dbg_mode == 3);
        const long long tq4 = prof ? clock64() : 0;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32);
        if (q_use) {
          tmem_ld_16x256b_x4(taddr, r);
          tmem_ld_16x256b_x4(taddr + (16u << 16), r + 16);
          tmem_wait_ld();
        } else {
          tmem_ld32(taddr, r);
        }
        if (prof) t_ld += clock64() - tq4;
        if (ci == CHUNKS_PER_HALF - 1) {
          // all of this warp's TMEM reads are complete: hand the accumulator buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[as]));
        }
        if (q_use) {
          epi_chunk_q<MODE>(g, et, r, r + 16, nb0, colbase + (long long)c * 32, atomic, kTcProbe && p.dbg_mode == 4);
          continue;
        }
        if (nb0 >= g.N || (kTcProbe && p.dbg_mode == 3)) continue;            // warp-uniform (dbg_mode 3: measurement without stores)
        if (t_path && nb0 + 32 <= g.N) {
          epi_chunk_t<MODE>(g, er, row_ok, et, r, nb0, colbase + (long long)c * 32, vec_bf16, stage, kTcProbe && p.dbg_mode == 4, atomic);
          continue;
        }
        if (!row_ok || er.skip) continue;
        epi_chunk<MODE>(g, er, r, nb0, colbase + (long long)c * 32 * cstr, cstr, atomic, stage);
      }
      if (prof) t_chunks += clock64() - tq3;
    }
    if (prof && lane == 0) { p.dbg[8] = clock64() - t_all; p.dbg[9] = w_tfull; p.dbg[10] = t_chunks; p.dbg[11] = t_ld; }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
  }
}


// =====================================================================================================
// CTA-pair kernel: two SMs of one TPC (a 2-CTA cluster) compute one 256 x TILE_N tile with
// tcgen05.mma.cta_group::2.  CTA r of the pair stages its own 128 A rows and ITS HALF of every B sub-tile
// (BNS/2 columns), so per output element the L2 -> SM traffic is 1.5-2x lower than in the single-CTA kernel
// (the single-CTA kernel is L2-fabric bound, see profiles/).  TILE_N = NSUB * BNS: up to three MMAs of N = BNS
// share one A tile (N = 384 = 2 x 192 K-major or 3 x 128 MN-major); the accumulator is double-buffered in TMEM when
// 2 * TILE_N <= 512 columns.  The leader CTA (rank 0) issues every MMA; completion is multicast to both CTAs'
// barriers; both CTAs run their own TMA producer and epilogue warps.  NCTA = 1 gives the same tiling on one SM.
template <int MODE, int BNS, int NSUB, int STAGES, int NCTA>
__global__ void __launch_bounds__(NTHREADS, 1) tcx_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB,
                                                                const TcParams p) {
  constexpr int TILE_N = BNS * NSUB;
  constexpr int BSUB_ROWS = BNS / NCTA;               // B rows (output columns) this CTA stages per sub-tile
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int BSUB_BYTES = BSUB_ROWS * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + NSUB * BSUB_BYTES;
  constexpr bool A_MN = (MODE == 2);
  constexpr bool B_MN = (MODE != 0);
  constexpr int ACC_BUFS = (2 * TILE_N <= 512) ? 2 : 1;
  constexpr int TMEM_COLS = (ACC_BUFS * TILE_N <= 128) ? 128 : (ACC_BUFS * TILE_N <= 256) ? 256 : 512;
  constexpr int CHUNKS_PER_HALF = TILE_N / 64;
  static_assert(TILE_N % 64 == 0 && TILE_N <= 512, "tile width");
  static_assert(BNS % 16 == 0 && BNS <= 256 && BSUB_ROWS % 8 == 0, "UMMA N");
  static_assert(MODE == 0 || BSUB_ROWS % 64 == 0, "MN-major B is staged in 64-column boxes");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Fs2Gemm& g = p.g;
  int* err = p.err;
  const uint32_t rank = (NCTA == 2) ? cluster_ctarank() : 0u;
  const int unit = (NCTA == 2) ? (int)cluster_id_x() : (int)blockIdx.x;
  const int nunits = (NCTA == 2) ? (int)cluster_count_x() : (int)gridDim.x;
  long long dbg_c0 = 0;
  uint64_t dbg_t0 = 0;
  if (kTcProbe && p.dbg && blockIdx.x == 0 && threadIdx.x == 0) { dbg_c0 = clock64(); dbg_t0 = globaltimer_ns(); }

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);      // the leader's producer arrives once (+ the pair's transaction bytes)
      mbar_init(smem_u32(&empty_bar[s]), 1);     // one MMA commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), NUM_EPI_WARPS * NCTA);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (NCTA == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);     // warp-uniform for the compiler
  pdl_wait();      // prologue above overlaps the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (both CTAs of the pair)
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      int s = 0;
      uint32_t ph = 0;
      bool ok = !(kTcProbe && p.dbg_mode == 1);
      const int p1a = p.pa[1], p2a = p.pa[2], p1b = p.pb[1], p2b = p.pb[2];
      const int pa_[4] = {0, p1a, p2a, 0}, pb_[4] = {0, p1b, p2b, 0};
      for (int t = unit; t < p.total_tiles && ok; t += nunits) {
        const int nt = t % p.n_tiles_total;
        const int rest = t / p.n_tiles_total;
        const int m0 = (rest % p.m_tiles) * (BM * NCTA) + (int)rank * BM;
        const int z = rest / p.m_tiles;
        const int zb = z / p.nsplit, zs = z % p.nsplit;
        const int i1 = zb % g.batch1, i2 = zb / g.batch1;
        int tapN = 0, n0 = nt * TILE_N;
        if (MODE == 2) { tapN = nt / p.ntiles_per_tap; n0 = (nt % p.ntiles_per_tap) * TILE_N; }
        const int kb_begin = zs * p.kb_per_split;
        const int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
        // running (tap, k offset) of the current k-block: no division inside the loop
        int j = (MODE == 2) ? 0 : kb_begin / p.kb_per_tap;
        int kk = (MODE == 2) ? kb_begin * BK : (kb_begin - j * p.kb_per_tap) * BK;
        const int kk_end = p.kb_per_tap * BK;
        const int nbase = n0 + (int)rank * BSUB_ROWS;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          if (!mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1, err)) { ok = false; break; }
          const uint32_t fb_local = smem_u32(&full_bar[s]);
          const uint32_t fb = (NCTA == 2) ? mapa_rank(fb_local, 0) : fb_local;   // bytes are credited to the leader
          if (rank == 0) mbar_expect_tx(fb_local, STAGE_BYTES * NCTA);
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          if (MODE == 0 || MODE == 1) {
            tma_issue_x<NCTA>(sa, &tmA, fb, pa_, kk, g.a_row_off + m0 + j * g.a_tap_step, i1, i2);
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
              const int nb = nbase + sub * BNS;
              if (MODE == 0) {
                tma_issue_x<NCTA>(sb + sub * BSUB_BYTES, &tmB, fb, pb_, j * g.b_tap_step + kk, nb, i1, i2);
              } else {
#pragma unroll
                for (int i = 0; i < BSUB_ROWS / 64; ++i)
                  tma_issue_x<NCTA>(sb + sub * BSUB_BYTES + i * (BK * 128), &tmB, fb, pb_, nb + 64 * i + j * g.b_tap_step,
                                    g.b_row_off + kk, i1, i2);
              }
            }
            kk += BK;
            if (kk >= kk_end) { kk = 0; ++j; }
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_issue_x<NCTA>(sa + i * (BK * 128), &tmA, fb, pa_, m0 + 64 * i, g.a_row_off + kk, i1, i2);
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
              const int nb = nbase + sub * BNS;
#pragma unroll
              for (int i = 0; i < BSUB_ROWS / 64; ++i)
                tma_issue_x<NCTA>(sb + sub * BSUB_BYTES + i * (BK * 128), &tmB, fb, pb_, nb + 64 * i,
                                  g.b_row_off + kk + tapN * g.b_tap_step, i1, i2);
            }
            kk += BK;
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA only, one thread)
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BNS >> 3) << 17) | ((uint32_t)((BM * NCTA) >> 4) << 24);
      uint32_t tc = 0, ph = 0;
      int s = 0;
      bool ok = true;
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t adesc0 = A_MN ? smem_desc(smem0, BK * 128, 1024) : smem_desc(smem0, 16, 1024);
      const uint64_t bdesc0 = B_MN ? smem_desc(smem0 + A_BYTES, BK * 128, 1024) : smem_desc(smem0 + A_BYTES, 16, 1024);
      constexpr uint64_t A_STEP = A_MN ? (2048 >> 4) : (32 >> 4);
      constexpr uint64_t B_STEP = B_MN ? (2048 >> 4) : (32 >> 4);
      constexpr uint64_t STAGE_STEP = STAGE_BYTES >> 4;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      for (int t = unit; t < p.total_tiles && ok; t += nunits, ++tc) {
        const int z = (t / p.n_tiles_total) / p.m_tiles;
        const int zs = z % p.nsplit;
        const int kb_begin = zs * p.kb_per_split;
        const int nkb = min(p.total_kb, kb_begin + p.kb_per_split) - kb_begin;
        const uint32_t as = tc % ACC_BUFS, aph = (tc / ACC_BUFS) & 1;
        if (!mbar_wait(smem_u32(&tempty_bar[as]), aph ^ 1, err)) { ok = false; break; }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * TILE_N;
        uint32_t acc = 0;
        for (int i = 0; i < nkb; ++i) {
          if (!(kTcProbe && p.dbg_mode == 1) && !mbar_wait(full0 + 8 * s, ph, err)) { ok = false; break; }
          tc_fence_after();
          // descriptors: stage-0 descriptor + (byte offset >> 4), see tc_gemm_kernel
          const uint64_t ad0 = adesc0 + (uint64_t)s * STAGE_STEP, bd0 = bdesc0 + (uint64_t)s * STAGE_STEP;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (kTcProbe && p.dbg_mode == 2) break;
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub)
              umma_bf16_x<NCTA>(tmem_d + sub * BNS, ad0 + k * A_STEP, bd0 + sub * (uint64_t)(BSUB_BYTES >> 4) + k * B_STEP, idesc,
                                acc);
            acc = 1;
          }
          umma_commit_x<NCTA>(empty0 + 8 * s);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (ok) umma_commit_x<NCTA>(smem_u32(&tfull_bar[as]));
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (both CTAs; own 128 rows of the tile)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* stage = smem + STAGES * STAGE_BYTES + 256 + (warp - 2) * EPI_STAGE_BYTES;
    const bool atomic = (p.nsplit > 1 && g.c_split_stride == 0) || g.accumulate;
    const long long cstr = (MODE == 2 && g.c_col_stride > 1) ? g.c_col_stride : 1;
    const bool al4 = ((g.ldc | g.c_col_off | g.c_s1 | g.c_s2) % 4 == 0) && (((uintptr_t)g.C) % 16 == 0);
    const bool al8 = ((g.ldc | g.c_col_off | g.c_s1 | g.c_s2) % 8 == 0) && (((uintptr_t)g.C) % 16 == 0);
    const bool aux_al = g.relu_aux == nullptr || (((uintptr_t)g.relu_aux) % 16 == 0);
    uint32_t tc = 0;
    bool ok = true;
    for (int t = unit; t < p.total_tiles && ok; t += nunits, ++tc) {
      const int nt = t % p.n_tiles_total;
      const int rest = t / p.n_tiles_total;
      const int m0 = (rest % p.m_tiles) * (BM * NCTA) + (int)rank * BM;
      const int z = rest / p.m_tiles;
      const int zb = z / p.nsplit;
      const int i1 = zb % g.batch1, i2 = zb / g.batch1;
      int tapN = 0, n0 = nt * TILE_N;
      if (MODE == 2) { tapN = nt / p.ntiles_per_tap; n0 = (nt % p.ntiles_per_tap) * TILE_N; }
      const uint32_t as = tc % ACC_BUFS, aph = (tc / ACC_BUFS) & 1;
      const int m = m0 + q * 32 + lane;
      EpiRow er;
      const bool row_ok = (m < g.M);
      if (row_ok) {
        epi_row_setup(g, i1, i2, m, er);
        er.base += (long long)(z % p.nsplit) * g.c_split_stride;       // 0 unless deterministic split-K
      }
      const long long colbase = (MODE == 2) ? (long long)tapN * g.c_tap_stride + n0 * cstr : n0;
      const bool vec_f32 = cstr == 1 && !g.c_bf16 && !atomic && al4 && (colbase % 4 == 0) && aux_al;
      const bool vec_bf16 = cstr == 1 && g.c_bf16 && al8 && (colbase % 8 == 0) && aux_al;
      // coalesced (smem-transposed) store path: decided and set up once per tile
      EpiT et;
      const bool vec_f32_at = cstr == 1 && !g.c_bf16 && atomic && al4 && (colbase % 4 == 0) && g.relu_aux == nullptr &&
                              g.bias == nullptr && !g.relu;
      bool t_path = p.epi_transpose && (vec_f32_at || (vec_f32 && g.relu_aux == nullptr) ||
                                        (vec_bf16 && (g.relu_aux == nullptr || g.aux_bf16)));
      if (t_path) {
        epi_t_setup(er, row_ok, et);
        t_path = !(et.mirrors && atomic);
      }
      // experiment switch FS2_TC_EPIT=3: fp32 outputs skip the smem transpose (direct 8-byte stores from the 16x256b TMEM
      // layout).  Measured equal-to-slower than the transpose (profiles/r01_summary.md section 3: in steady state these
      // GEMMs are bound by the HBM traffic of their fp32 output, not by the store instruction pattern), so it is off.
      const bool q_path = kTcProbe && t_path && p.epi_transpose == 3 && (vec_f32_at || (vec_f32 && g.relu_aux == nullptr)) &&
                          (g.bias == nullptr || (((uintptr_t)g.bias) % 8 == 0));
      if (!mbar_wait(smem_u32(&tfull_bar[as]), aph, err)) { ok = false; break; }
      tc_fence_after();
#pragma unroll 1
      for (int ci = 0; ci < CHUNKS_PER_HALF; ++ci) {
        const int c = half * CHUNKS_PER_HALF + ci;
        uint32_t r[32];
        const int nb0 = n0 + c * 32;
        const bool q_use = q_path && nb0 + 32 <= g.N && !(kTcProbe && p.dbg_mode == 3);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * TILE_N + c * 32);
        if (q_use) {
          tmem_ld_16x256b_x4(taddr, r);
          tmem_ld_16x256b_x4(taddr + (16u << 16), r + 16);
          tmem_wait_ld();
        } else {
          tmem_ld32(taddr, r);
        }
        if (ci == CHUNKS_PER_HALF - 1) {
          // this warp's TMEM reads are done: hand the accumulator buffer back to the (leader's) MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            const uint32_t tb = smem_u32(&tempty_bar[as]);
            if (NCTA == 2 && rank != 0) mbar_arrive_cluster(mapa_rank(tb, 0));
            else mbar_arrive(tb);
          }
        }
        if (q_use) {
          epi_chunk_q<MODE>(g, et, r, r + 16, nb0, colbase + (long long)c * 32, atomic, kTcProbe && p.dbg_mode == 4);
          continue;
        }
        if (nb0 >= g.N || (kTcProbe && p.dbg_mode == 3)) continue;            // warp-uniform (dbg_mode 3: measurement without stores)
        if (t_path && nb0 + 32 <= g.N) {
          epi_chunk_t<MODE>(g, er, row_ok, et, r, nb0, colbase + (long long)c * 32, vec_bf16, stage, kTcProbe && p.dbg_mode == 4, atomic);
          continue;
        }
        if (!row_ok || er.skip) continue;
        epi_chunk<MODE>(g, er, r, nb0, colbase + (long long)c * 32 * cstr, cstr, atomic, stage);
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  if (NCTA == 2) cluster_sync_all(); else __syncthreads();     // nobody leaves while the peer can still signal its barriers
  if (kTcProbe && p.dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    p.dbg[0] = clock64() - dbg_c0;
    p.dbg[1] = (long long)(globaltimer_ns() - dbg_t0);
  }
  if (warp == 1) {
    if (NCTA == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------ host side --
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::mutex g_mu;
std::unordered_map<std::string, CUtensorMap> g_map_cache;

int get_encode() {
  if (g_encode) return FS2_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    fs2_set_error("cuTensorMapEncodeTiled not available from the driver");
    return FS2_ERR_CUDA;
  }
  g_encode = (EncodeTiledFn)fn;
  return FS2_OK;
}

// Build (or fetch) a 4-D bf16 tensor map for an operand described as
// (inner, rows, batch1, batch2) with element strides (1, ld, s1, s2).  Dimensions are ordered
// by ascending stride; perm[] returns the map slot of (inner,row,i1,i2).
int make_map(const void* base, long long inner, long long rows, int b1, int b2, long long ld, long long s1,
             long long s2, int box_inner, int box_rows, CUtensorMap* out, int perm[4]) {
  struct D { long long size, stride; int box, id; };
  D d[3] = {{rows, ld, box_rows, 1}, {b1, s1, 1, 2}, {b2, s2, 1, 3}};
  long long mx = ld * rows;
  for (int i = 1; i < 3; ++i)
    if (d[i].size > 1 && d[i].stride * d[i].size > mx) mx = d[i].stride * d[i].size;
  for (int i = 1; i < 3; ++i)
    if (d[i].size <= 1) { d[i].size = 1; d[i].stride = mx; mx *= 1; }
  // sort by stride (stable, 3 elements)
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (d[j].stride < d[i].stride) { D t = d[i]; d[i] = d[j]; d[j] = t; }
  perm[0] = 0;
  for (int i = 0; i < 3; ++i) perm[d[i].id] = i + 1;
  char key[256];
  snprintf(key, sizeof key, "%p|%lld|%lld|%lld|%lld|%lld|%lld|%lld|%lld|%d|%d|%d%d%d", base, inner, d[0].size, d[1].size,
           d[2].size, d[0].stride, d[1].stride, d[2].stride, 0LL, box_inner, box_rows, perm[1], perm[2], perm[3]);
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_map_cache.find(key);
  if (it != g_map_cache.end()) { *out = it->second; return FS2_OK; }
  cuuint64_t gdim[4] = {(cuuint64_t)inner, (cuuint64_t)d[0].size, (cuuint64_t)d[1].size, (cuuint64_t)d[2].size};
  cuuint64_t gstr[3] = {(cuuint64_t)d[0].stride * 2, (cuuint64_t)d[1].stride * 2, (cuuint64_t)d[2].stride * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_inner, (cuuint32_t)d[0].box, (cuuint32_t)d[1].box, (cuuint32_t)d[2].box};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (((uintptr_t)base % 16) != 0 || gstr[0] % 16 || gstr[1] % 16 || gstr[2] % 16) {
    fs2_set_error("fs2_gemm_tc: operand base/strides must be 16-byte aligned for TMA");
    return FS2_ERR_ARG;
  }
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[512];
    snprintf(msg, sizeof msg,
             "cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu strides %llu %llu %llu box %u %u %u %u", (int)r,
             (unsigned long long)gdim[0], (unsigned long long)gdim[1], (unsigned long long)gdim[2],
             (unsigned long long)gdim[3], (unsigned long long)gstr[0], (unsigned long long)gstr[1],
             (unsigned long long)gstr[2], box[0], box[1], box[2], box[3]);
    fs2_set_error(msg);
    return FS2_ERR_CUDA;
  }
  if (g_map_cache.size() > 8192) g_map_cache.clear();   // shapes vary per batch: keep the cache bounded
  g_map_cache[key] = *out;
  return FS2_OK;
}

int g_num_sms = 0;

template <int MODE, int BN, int STAGES>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, TcParams& p, cudaStream_t st) {
  constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 + 256 + EPI_STAGE_TOTAL + (MODE == 2 ? 768 + ONES_BYTES : 0);
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(tc_gemm_kernel<MODE, BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  if (g_num_sms == 0) {
    int dev = 0;
    CUDA_CHECK_RET(cudaGetDevice(&dev));
    CUDA_CHECK_RET(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  p.ntiles_per_tap = (p.g.N + BN - 1) / BN;
  p.n_tiles_total = (p.g.mode == 2) ? p.ntiles_per_tap * p.g.taps : p.ntiles_per_tap;
  p.m_tiles = (p.g.M + BM - 1) / BM;
  const long long total = (long long)p.n_tiles_total * p.m_tiles * p.g.batch1 * p.g.batch2 * p.nsplit;
  if (total > 0x7FFFFFFF) { fs2_set_error("fs2_gemm_tc: too many tiles"); return FS2_ERR_ARG; }
  p.total_tiles = (int)total;
  const int grid = p.total_tiles < g_num_sms ? p.total_tiles : g_num_sms;
  FS2_LAUNCH((tc_gemm_kernel<MODE, BN, STAGES>), grid, NTHREADS, SMEM, st, ta, tb, p);
  return fs2_check_launch();
}

template <int MODE>
int dispatch_bn(int bn, const CUtensorMap& ta, const CUtensorMap& tb, TcParams& p, cudaStream_t st) {
  if (bn == 256) return launch<MODE, 256, 4>(ta, tb, p, st);
  if (bn == 192) return launch<MODE, 192, 5>(ta, tb, p, st);
  return launch<MODE, 128, 6>(ta, tb, p, st);
}

// tile width: least padded columns, ties -> wider tile; when that leaves SMs without a tile (short M, e.g. the
// phoneme-side GEMMs) fall back to 128-wide tiles
int pick_bn(int N, long long m_tiles_x_batch) {
  int best = 128;
  long long best_pad = ((N + 127) / 128) * 128LL;
  const int cands[2] = {192, 256};
  for (int i = 0; i < 2; ++i) {
    long long pad = ((N + cands[i] - 1) / cands[i]) * (long long)cands[i];
    if (pad <= best_pad) { best_pad = pad; best = cands[i]; }
  }
  if (best != 128 && g_num_sms > 0 && m_tiles_x_batch * ((N + best - 1) / best) < g_num_sms &&
      ((N + 127) / 128) * 128LL <= best_pad + 64)
    best = 128;
  return best;
}


template <int MODE, int BNS, int NSUB, int STAGES, int NCTA>
int launch_x(const CUtensorMap& ta, const CUtensorMap& tb, TcParams& p, cudaStream_t st) {
  constexpr int TILE_N = BNS * NSUB;
  constexpr int SMEM = STAGES * (BM * BK * 2 + NSUB * (BNS / NCTA) * BK * 2) + 1024 + 256 + EPI_STAGE_TOTAL;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  static int units = 0;       // CTA pairs (or CTAs) that can be resident at once
  auto kern = tcx_gemm_kernel<MODE, BNS, NSUB, STAGES, NCTA>;
  if (g_num_sms == 0) {
    int dev = 0;
    CUDA_CHECK_RET(cudaGetDevice(&dev));
    CUDA_CHECK_RET(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = g_fs2_pdl ? 2 : 1;
  if (units == 0) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    if (NCTA == 2) {
      cfg.gridDim = dim3(g_num_sms / NCTA * NCTA, 1, 1);
      int nc = 0;
      CUDA_CHECK_RET(cudaOccupancyMaxActiveClusters(&nc, kern, &cfg));
      if (nc <= 0) { fs2_set_error("fs2_gemm_tc: no resident CTA pair possible"); return FS2_ERR_CUDA; }
      units = nc < g_num_sms / NCTA ? nc : g_num_sms / NCTA;
    } else {
      units = g_num_sms;
    }
  }
  p.ntiles_per_tap = (p.g.N + TILE_N - 1) / TILE_N;
  p.n_tiles_total = (p.g.mode == 2) ? p.ntiles_per_tap * p.g.taps : p.ntiles_per_tap;
  p.m_tiles = (p.g.M + BM * NCTA - 1) / (BM * NCTA);
  const long long tiles = (long long)p.n_tiles_total * p.m_tiles * p.g.batch1 * p.g.batch2;
  if (p.g.mode == 2 && p.g.split_k <= 0) {
    // auto split-K: the persistent grid walks tiles*split work items in waves of one item per unit; pick the split
    // that minimises waves * (k-blocks per item + the atomic epilogue's cost in k-block units)
    const int epi_kb = 8;
    long long best_cost = -1;
    int best = 1;
    for (int s = 1; s <= 64 && s <= p.total_kb; ++s) {
      const long long waves = (tiles * s + units - 1) / units;
      const long long cost = waves * ((p.total_kb + s - 1) / s + (s > 1 || p.g.accumulate ? epi_kb : epi_kb / 2));
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
    }
    p.nsplit = best;
    p.kb_per_split = (p.total_kb + p.nsplit - 1) / p.nsplit;
    p.nsplit = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;
  }
  const long long total = tiles * p.nsplit;
  if (total > 0x7FFFFFFF) { fs2_set_error("fs2_gemm_tc: too many tiles"); return FS2_ERR_ARG; }
  p.total_tiles = (int)total;
  const int nunits = p.total_tiles < units ? p.total_tiles : units;
  cfg.gridDim = dim3(nunits * NCTA, 1, 1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
  if (e != cudaSuccess) { fs2_set_error(cudaGetErrorString(e)); return FS2_ERR_CUDA; }
  return fs2_check_launch();
}

// Tile configurations of the pair kernel: 0 = 256 x 256 (double-buffered accumulator), 1 = 256 x 384 (2 x 192
// K-major B, or 3 x 128 MN-major B), 2 = 256 x 128.
template <int MODE>
int dispatch_x(int cfg, const CUtensorMap& ta, const CUtensorMap& tb, TcParams& p, cudaStream_t st) {
  if (cfg == 0) return launch_x<MODE, 256, 1, 6, 2>(ta, tb, p, st);
  if (cfg == 1) {
    if constexpr (MODE == 0) return launch_x<MODE, 192, 2, 5, 2>(ta, tb, p, st);
    else return launch_x<MODE, 128, 3, 5, 2>(ta, tb, p, st);
  }
  return launch_x<MODE, 128, 1, 8, 2>(ta, tb, p, st);
}

// the same tilings on one SM (measurement only: isolates what the CTA pair buys)
template <int MODE>
int dispatch_x1(int cfg, const CUtensorMap& ta, const CUtensorMap& tb, TcParams& p, cudaStream_t st) {
  if (cfg == 0) return launch_x<MODE, 256, 1, 4, 1>(ta, tb, p, st);
  if (cfg == 1) {
    if constexpr (MODE == 0) return launch_x<MODE, 192, 2, 3, 1>(ta, tb, p, st);
    else return launch_x<MODE, 128, 3, 3, 1>(ta, tb, p, st);
  }
  return launch_x<MODE, 128, 1, 6, 1>(ta, tb, p, st);
}

// per-CTA B box rows (mode 0) of a configuration
int cfg_tile_n(int cfg) { return cfg == 0 ? 256 : cfg == 1 ? 384 : 128; }
int cfg_bsub_rows(int mode, int cfg) { return cfg == 0 ? 128 : cfg == 1 ? (mode == 0 ? 96 : 64) : 64; }

// least padded columns, ties -> 256-wide (double-buffered accumulator, full-rate N = 256 MMAs); between the two widths
// that divide N equally well (N = 1536: 256 or 384) the one with fewer persistent rounds x tile width wins -- at
// phoneme-length row counts 17 x 4 tiles of 384 fit one round of the 74 CTA pairs where 17 x 6 tiles of 256 need two
int pick_cfg(int N, long long m_tiles, int units) {
  int best = 2;
  long long best_pad = ((N + 127) / 128) * 128LL;
  const int cands[2] = {1, 0};
  for (int i = 0; i < 2; ++i) {
    const int tn = cfg_tile_n(cands[i]);
    long long pad = ((N + tn - 1) / tn) * (long long)tn;
    if (pad <= best_pad) { best_pad = pad; best = cands[i]; }
  }
  if (best == 0 && N % 384 == 0 && units > 0 && m_tiles > 0) {
    const long long c256 = ((m_tiles * (N / 256) + units - 1) / units) * 256;
    const long long c384 = ((m_tiles * (N / 384) + units - 1) / units) * 384;
    if (c384 * 100 < c256 * 85) best = 1;      // the 384-wide tile is slower per FLOP (single accumulator buffer, N = 128 MMAs)
  }
  return best;
}

int g_dbg_mode = 0;
int g_use_pair = -1;    // 0: single-CTA kernel only, 1: CTA-pair kernel wherever it applies, 2: heuristic (FS2_TC_PAIR)
int g_force_cfg = -2;   // FS2_TC_CFG=0|1|2 forces a tile configuration
int g_epi_transpose = 1;   // FS2_TC_EPIT=0 switches the coalesced (smem-transposed) epilogue off (A/B measurements)

// heuristic kernel choice (measured on B200, tools/gemm_sweep.py): filled in from the sweep
bool prefer_pair(const Fs2Gemm& g) {
  // the pair kernel wins (1-8 %) where its 256-wide tiles divide N and the reduction is long; elsewhere the single-CTA
  // kernel's exact-fit 192 / 128 tiles are faster (gpurun_out/gemm_sweep4.log)
  if (g.N % 256 == 0 && (long long)g.K * g.taps >= 1024 && g.M >= 2048) return true;
  // the k = 9 FFN dgrad (N = 384, reduction 1536 x 9) on mel-length row counts: pair/384 151 us vs 164 us single-CTA/192 at
  // T = 488; at phoneme-length row counts (4352 rows) the pair kernel has too few tiles and loses 1.7x
  return g.mode == 1 && g.N == 384 && (long long)g.K * g.taps >= 8192 && g.M >= 8192;
}

}  // namespace

long long* g_dbg = nullptr;
extern "C" int fs2_gemm_tc_set_debug(long long* dev_buf) {
  g_dbg = dev_buf;
  return FS2_OK;
}

// shared with attention.cu: a (inner, rows) bf16 tensor map with a (box_inner, box_rows) box, 128B swizzle, addressed
// with 4-D coordinates (col, row, 0, 0); and the device error word of the bounded mbarrier waits
int fs2_tc_make_map_2d(const void* base, long long inner, long long rows, long long ld, int box_inner, int box_rows,
                       CUtensorMap* out) {
  int rc = get_encode();
  if (rc) return rc;
  int perm[4];
  return make_map(base, inner, rows, 1, 1, ld, 0, 0, box_inner, box_rows, out, perm);
}
int fs2_tc_error_ptr(int** out) {
  static int* errp = nullptr;
  if (!errp) CUDA_CHECK_RET(cudaGetSymbolAddress((void**)&errp, g_tc_error));
  *out = errp;
  return FS2_OK;
}

// tuning hook (tools/gemm_sweep.py): pair = 0 single-CTA, 1 CTA pair, 2 heuristic; cfg = -1 auto or 0|1|2
extern "C" int fs2_gemm_tc_tune(int pair, int cfg) {
  g_use_pair = pair & 3;
  g_dbg_mode = (pair >> 4) & 7;     // bits 4-5: pipeline-isolation experiments (tools/gemm_sweep.py), 0 in production
  g_force_cfg = cfg;
  return FS2_OK;
}

// read-and-clear
extern "C" int fs2_gemm_tc_error_flag(void) {
  int v = 0, z = 0;
  cudaMemcpyFromSymbol(&v, g_tc_error, sizeof(int));
  if (v) cudaMemcpyToSymbol(g_tc_error, &z, sizeof(int));
  return v;
}

// Fs2Gemm.a_colsum where the chosen GEMM kernel cannot fold it in: the stand-alone column-sum kernel on the same stream
static int colsum_fallback(const Fs2Gemm& g, cudaStream_t st) {
  if (g.mode != 2 || g.batch1 * g.batch2 != 1) { fs2_set_error("fs2_gemm_tc: a_colsum needs mode 2 without batching"); return FS2_ERR_ARG; }
  if (g.alpha != 1.f) { fs2_set_error("fs2_gemm_tc: a_colsum with alpha != 1 needs the 128 x 192-tile kernel"); return FS2_ERR_UNSUPPORTED; }
  const bf16* a = (const bf16*)g.A + (long long)g.a_row_off * g.lda;
  return fs2_colsum(a, 1, (long long)g.K, g.M, g.lda, g.a_colsum, (void*)st);
}

extern "C" int fs2_gemm_tc(const Fs2Gemm* gp, void* stream) {
  if (!gp || !gp->A || !gp->B || !gp->C) { fs2_set_error("fs2_gemm_tc: null pointer"); return FS2_ERR_ARG; }
  const Fs2Gemm& g = *gp;
  if (!g.ab_bf16) { fs2_set_error("fs2_gemm_tc: operands must be bf16"); return FS2_ERR_UNSUPPORTED; }
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return FS2_OK;
  int rc = get_encode();
  if (rc) return rc;
  const int BN = (g_force_cfg >= 0 && g_force_cfg <= 2 && g_use_pair == 0) ? (g_force_cfg == 0 ? 256 : g_force_cfg == 1 ? 192 : 128)
                                                                            : pick_bn(g.N, g.mode == 2 ? (1LL << 40) : (long long)((g.M + BM - 1) / BM) * g.batch1 * g.batch2);
  TcParams p;
  memset(&p, 0, sizeof p);
  p.g = g;
  p.kb_per_tap = (g.K + BK - 1) / BK;
  p.total_kb = (g.mode == 2) ? p.kb_per_tap : p.kb_per_tap * g.taps;
  p.nsplit = (g.mode == 2 && g.split_k > 1) ? g.split_k : 1;
  if (g.mode != 2 && g.split_k > 1 && g.c_split_stride != 0) {
    // deterministic split-K: every split stores its partial result to its own copy of C
    if (g.c_bf16 || g.accumulate || g.bias || g.relu || g.relu_aux) {
      fs2_set_error("fs2_gemm_tc: c_split_stride needs fp32 C and a plain epilogue");
      return FS2_ERR_ARG;
    }
    p.nsplit = g.split_k;
  }
  if (g.mode == 2 && g.split_k <= 0) {
    // auto split-K: the persistent grid walks tiles*split work items in waves of one item per SM; pick the split
    // that minimises waves * (k-blocks per item + the atomic epilogue's cost in k-block units)
    if (g_num_sms == 0) {
      int dev = 0;
      CUDA_CHECK_RET(cudaGetDevice(&dev));
      CUDA_CHECK_RET(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long tiles = (long long)((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN) * g.taps * g.batch1 * g.batch2;
    const int epi_kb = 8;
    long long best_cost = -1;
    int best = 1;
    for (int s = 1; s <= 64 && s <= p.total_kb; ++s) {
      const long long waves = (tiles * s + g_num_sms - 1) / g_num_sms;
      const long long cost = waves * ((p.total_kb + s - 1) / s + (s > 1 || g.accumulate ? epi_kb : epi_kb / 2));
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
    }
    p.nsplit = best;
  }
  if (p.nsplit > p.total_kb) p.nsplit = p.total_kb;
  p.kb_per_split = (p.total_kb + p.nsplit - 1) / p.nsplit;
  p.nsplit = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;
  if (p.nsplit > 1 && g.c_bf16) { fs2_set_error("fs2_gemm_tc: split_k needs fp32 C"); return FS2_ERR_ARG; }
  static int* errp = nullptr;     // resolved once (also keeps the call out of CUDA-graph captures)
  if (!errp) CUDA_CHECK_RET(cudaGetSymbolAddress((void**)&errp, g_tc_error));
  p.err = errp;
  p.dbg = g_dbg;
  p.dbg_mode = g_dbg_mode;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_use_pair < 0) {
    const char* e = getenv("FS2_TC_PAIR");
    g_use_pair = (e && e[0] == '0') ? 0 : (e && e[0] == '1') ? 1 : 2;
    const char* c = getenv("FS2_TC_CFG");
    g_force_cfg = c ? atoi(c) : -1;
    const char* t = getenv("FS2_TC_EPIT");
    if (t) g_epi_transpose = atoi(t);          // 0 off, 1 default (smem-transposed stores), 3 direct quad-layout stores for fp32 outputs (experiment)
  }
  p.epi_transpose = g_epi_transpose;
  const bool pair_ok = g.batch1 * g.batch2 == 1 && g.M > BM;
  const bool use_pair = pair_ok && (g_use_pair == 1 || g_use_pair == 3 || (g_use_pair == 2 && prefer_pair(g)));
  // dgrad with a plain fp32 epilogue and fewer tiles than ~2 waves (phoneme-length row counts: 56 tiles of a k = 9
  // reduction on 148 SMs): split the reduction like the weight gradients do; the partial sums meet in C through the
  // 16-byte vector atomics of the transposed epilogue.  C is zeroed here.
  if (!use_pair && g.mode == 1 && g.split_k == 0 && !g.c_bf16 && !g.accumulate && !g.bias && !g.relu && !g.relu_aux &&
      g.rs_Tp == 0 && g.lens == nullptr && g.halo == 0 && g.batch1 * g.batch2 == 1 && g.ldc == g.N && g.c_col_off == 0 &&
      g_epi_transpose) {
    if (g_num_sms == 0) {
      int dev = 0;
      CUDA_CHECK_RET(cudaGetDevice(&dev));
      CUDA_CHECK_RET(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long tiles = (long long)((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
    const int epi_kb = 4;
    long long best_cost = -1;
    int best = 1;
    for (int s = 1; s <= 16 && s <= p.total_kb; ++s) {
      const long long waves = (tiles * s + g_num_sms - 1) / g_num_sms;
      const long long cost = waves * ((p.total_kb + s - 1) / s + (s > 1 ? 2 * epi_kb : epi_kb));
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
    }
    if (best > 1) {
      p.nsplit = best;
      p.kb_per_split = (p.total_kb + p.nsplit - 1) / p.nsplit;
      p.nsplit = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;
      if (p.nsplit > 1)
        CUDA_CHECK_RET(cudaMemsetAsync((float*)g.C + (long long)g.c_row_off * g.ldc, 0, (size_t)g.M * g.N * sizeof(float), st));
    }
  }
  if (use_pair) {
    if (g.a_colsum) {                      // the pair kernel does not fold the bias gradient
      rc = colsum_fallback(g, st);
      if (rc) return rc;
      p.g.a_colsum = nullptr;
    }
    const int ncta = g_use_pair == 3 ? 1 : 2;
    if (g_num_sms == 0) {
      int dev = 0;
      CUDA_CHECK_RET(cudaGetDevice(&dev));
      CUDA_CHECK_RET(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const int cfg = (g_force_cfg >= 0 && g_force_cfg <= 2)
                        ? g_force_cfg
                        : pick_cfg(g.N, g.mode == 2 ? 0 : (g.M + BM * ncta - 1) / (BM * ncta), g_num_sms / ncta);
    CUtensorMap ta, tb;
    if (g.mode == 2) rc = make_map(g.A, g.a_inner, g.a_rows, g.batch1, g.batch2, g.lda, g.a_s1, g.a_s2, 64, BK, &ta, p.pa);
    else rc = make_map(g.A, g.a_inner, g.a_rows, g.batch1, g.batch2, g.lda, g.a_s1, g.a_s2, BK, BM, &ta, p.pa);
    if (rc) return rc;
    if (g.mode == 0)
      rc = make_map(g.B, g.b_inner, g.b_rows, g.batch1, g.batch2, g.ldb, g.b_s1, g.b_s2, BK, cfg_bsub_rows(0, cfg) * (3 - ncta),
                    &tb, p.pb);
    else rc = make_map(g.B, g.b_inner, g.b_rows, g.batch1, g.batch2, g.ldb, g.b_s1, g.b_s2, 64, BK, &tb, p.pb);
    if (rc) return rc;
    if (ncta == 1) {
      if (g.mode == 0) return dispatch_x1<0>(cfg, ta, tb, p, st);
      if (g.mode == 1) return dispatch_x1<1>(cfg, ta, tb, p, st);
      return dispatch_x1<2>(cfg, ta, tb, p, st);
    }
    if (g.mode == 0) return dispatch_x<0>(cfg, ta, tb, p, st);
    if (g.mode == 1) return dispatch_x<1>(cfg, ta, tb, p, st);
    if (g.mode == 2) return dispatch_x<2>(cfg, ta, tb, p, st);
    fs2_set_error("fs2_gemm_tc: bad mode");
    return FS2_ERR_ARG;
  }
  // bias-gradient fold: only the 128 x 192-tile weight-gradient kernel has the spare tensor-memory columns
  if (g.a_colsum && !(g.mode == 2 && BN == 192 && g.batch1 * g.batch2 == 1)) {
    rc = colsum_fallback(g, st);
    if (rc) return rc;
    p.g.a_colsum = nullptr;
  }
  CUtensorMap ta, tb;
  // A: mode 0/1 K-major box (64 k, 128 rows); mode 2 MN-major box (64 m, 64 k-rows)
  if (g.mode == 2) rc = make_map(g.A, g.a_inner, g.a_rows, g.batch1, g.batch2, g.lda, g.a_s1, g.a_s2, 64, BK, &ta, p.pa);
  else rc = make_map(g.A, g.a_inner, g.a_rows, g.batch1, g.batch2, g.lda, g.a_s1, g.a_s2, BK, BM, &ta, p.pa);
  if (rc) return rc;
  if (g.mode == 0) rc = make_map(g.B, g.b_inner, g.b_rows, g.batch1, g.batch2, g.ldb, g.b_s1, g.b_s2, BK, BN, &tb, p.pb);
  else rc = make_map(g.B, g.b_inner, g.b_rows, g.batch1, g.batch2, g.ldb, g.b_s1, g.b_s2, 64, BK, &tb, p.pb);
  if (rc) return rc;
  if (g.mode == 0) return dispatch_bn<0>(BN, ta, tb, p, st);
  if (g.mode == 1) return dispatch_bn<1>(BN, ta, tb, p, st);
  if (g.mode == 2) return dispatch_bn<2>(BN, ta, tb, p, st);
  fs2_set_error("fs2_gemm_tc: bad mode");
  return FS2_ERR_ARG;
}
