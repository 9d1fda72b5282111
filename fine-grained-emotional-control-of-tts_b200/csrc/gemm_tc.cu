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
#include <mutex>
#include <unordered_map>
#include <string>
#include "common.cuh"
#include "../../include/fs2_b200.h"
#include "gemm_epilogue.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int NTHREADS = 192;

struct TcParams {
  Fs2Gemm g;
  int kb_per_tap;     // ceil(K / 64)
  int total_kb;       // K-blocks in the whole reduction
  int kb_per_split;
  int nsplit;
  int ntiles_per_tap; // mode 2
  int pa[4], pb[4];   // tensor-map dim slot of (inner,row,i1,i2) for A and B
  int* err;
};

__device__ int g_tc_error = 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must never hang the GPU.  On timeout the error word is set
// and every role falls through to the teardown.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err) {
  if (mbar_try_wait(bar, parity)) return true;
  uint64_t t0 = globaltimer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3FF) == 0) {
      if (globaltimer_ns() - t0 > 400000000ull || *(volatile int*)err != 0) {
        atomicExch(err, 1);
        return false;
      }
    }
  }
  return true;
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_issue(uint32_t dst, const CUtensorMap* tm, uint32_t bar, const int* perm, int inner,
                                          int row, int i1, int i2) {
  int c[4];
  c[perm[0]] = inner;
  c[perm[1]] = row;
  c[perm[2]] = i1;
  c[perm[3]] = i2;
  tma_load_4d(dst, tm, bar, c[0], c[1], c[2], c[3]);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 "version 1".
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // LayoutType::SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int MODE, int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS) tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB,
                                                            const TcParams p) {
  constexpr int A_BYTES = BM * BK * 2;          // 16 KB
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr bool A_MN = (MODE == 2);
  constexpr bool B_MN = (MODE != 0);
  constexpr int TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : BN <= 256 ? 256 : 512;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = (uint32_t*)(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Fs2Gemm& g = p.g;
  int* err = p.err;

  const int nt = blockIdx.x;
  const int m0 = blockIdx.y * BM;
  const int zb = blockIdx.z / p.nsplit, zs = blockIdx.z % p.nsplit;
  const int i1 = zb % g.batch1, i2 = zb / g.batch1;
  int tapN = 0, n0 = nt * BN;
  if (MODE == 2) { tapN = nt / p.ntiles_per_tap; n0 = (nt % p.ntiles_per_tap) * BN; }
  int kb_begin = zs * p.kb_per_split;
  int kb_end = min(p.total_kb, kb_begin + p.kb_per_split);
  const int nkb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(accum_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && nkb > 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      for (int it = 0; it < nkb; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        if (!mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1, err)) break;
        const uint32_t fb = smem_u32(&full_bar[s]);
        mbar_expect_tx(fb, STAGE_BYTES);
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
        const int kb = kb_begin + it;
        if (MODE == 0 || MODE == 1) {
          const int j = kb / p.kb_per_tap;
          const int kk = (kb - j * p.kb_per_tap) * BK;
          tma_issue(sa, &tmA, fb, p.pa, kk, g.a_row_off + m0 + j * g.a_tap_step, i1, i2);
          if (MODE == 0) {
            tma_issue(sb, &tmB, fb, p.pb, j * g.b_tap_step + kk, n0, i1, i2);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_issue(sb + i * (BK * 128), &tmB, fb, p.pb, n0 + 64 * i + j * g.b_tap_step, g.b_row_off + kk, i1, i2);
          }
        } else {
          const int k0 = kb * BK;
#pragma unroll
          for (int i = 0; i < BM / 64; ++i)
            tma_issue(sa + i * (BK * 128), &tmA, fb, p.pa, m0 + 64 * i, g.a_row_off + k0, i1, i2);
#pragma unroll
          for (int i = 0; i < BN / 64; ++i)
            tma_issue(sb + i * (BK * 128), &tmB, fb, p.pb, n0 + 64 * i, g.b_row_off + k0 + tapN * g.b_tap_step, i1, i2);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nkb > 0) {
      // instruction descriptor: D=f32, A=B=bf16, majorness, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      bool ok = true;
      for (int it = 0; it < nkb && ok; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        if (!mbar_wait(smem_u32(&full_bar[s]), ph, err)) { ok = false; break; }
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // K-major: 16 elements = 32 B inside the 128 B swizzle row; 8-row groups 1024 B apart.
          // MN-major: 16 k-rows = 2 groups of 8 rows (1024 B each); 64-wide MN blocks 8192 B apart.
          const uint64_t ad = A_MN ? smem_desc(sa + k * 2048, BK * 128, 1024) : smem_desc(sa + k * 32, 16, 1024);
          const uint64_t bd = B_MN ? smem_desc(sb + k * 2048, BK * 128, 1024) : smem_desc(sb + k * 32, 16, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[s]));
      }
      umma_commit(smem_u32(accum_bar));
    }
  } else {
    // ---------------- epilogue: 4 warps, TMEM lane quadrant = warp % 4 ----------------
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    bool ok = true;
    if (nkb > 0) ok = mbar_wait(smem_u32(accum_bar), 0, err);
    tc_fence_after();
    EpiRow er;
    const bool row_ok = ok && (m < g.M);
    if (row_ok) epi_row_setup(g, i1, i2, m, er);
    const bool atomic = p.nsplit > 1 || g.accumulate;
    const long long cstr = (MODE == 2 && g.c_col_stride > 1) ? g.c_col_stride : 1;
    const long long colbase = (MODE == 2) ? (long long)tapN * g.c_tap_stride + n0 * cstr : n0;
    const bool vec_f32 = cstr == 1 && !g.c_bf16 && !atomic && ((g.ldc | g.c_col_off | g.c_s1 | g.c_s2 | colbase) % 4 == 0) &&
                         (((uintptr_t)g.C) % 16 == 0);
    const bool vec_bf16 = cstr == 1 && g.c_bf16 && ((g.ldc | g.c_col_off | g.c_s1 | g.c_s2 | colbase) % 8 == 0) &&
                          (((uintptr_t)g.C) % 16 == 0);
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      if (n0 + c * 32 >= g.N) break;       // warp-uniform
      uint32_t r[32];
      if (nkb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = 0;
      }
      if (!row_ok || er.skip) continue;
      const int nb0 = n0 + c * 32;
      const long long col0 = colbase + c * 32 * cstr;
      const bool full = (nb0 + 32 <= g.N);
      if (full && (vec_f32 || vec_bf16)) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = epi_value(g, er, col0 + i, nb0 + i, __uint_as_float(r[i]));
        if (vec_f32) {
          float* dst = (float*)g.C + er.base + col0;
#pragma unroll
          for (int i = 0; i < 8; ++i) st4(dst + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const long long mo = mi ? er.mirror2 : er.mirror;
            if (mo) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                st4(dst + mo + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            }
          }
        } else {
          bf16* dst = (bf16*)g.C + er.base + col0;
#pragma unroll
          for (int i = 0; i < 8; ++i) st4(dst + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
#pragma unroll
          for (int mi = 0; mi < 2; ++mi) {
            const long long mo = mi ? er.mirror2 : er.mirror;
            if (mo) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                st4(dst + mo + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (nb0 + i < g.N) epi_store(g, er, col0 + i * cstr, nb0 + i, __uint_as_float(r[i]), atomic);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS)
                 : "memory");
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
  g_map_cache[key] = *out;
  return FS2_OK;
}

template <int MODE, int BN, int STAGES>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const TcParams& p, dim3 grid, cudaStream_t st) {
  constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(tc_gemm_kernel<MODE, BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  tc_gemm_kernel<MODE, BN, STAGES><<<grid, NTHREADS, SMEM, st>>>(ta, tb, p);
  return fs2_check_launch();
}

}  // namespace

// read-and-clear
extern "C" int fs2_gemm_tc_error_flag(void) {
  int v = 0, z = 0;
  cudaMemcpyFromSymbol(&v, g_tc_error, sizeof(int));
  if (v) cudaMemcpyToSymbol(g_tc_error, &z, sizeof(int));
  return v;
}

extern "C" int fs2_gemm_tc(const Fs2Gemm* gp, void* stream) {
  if (!gp || !gp->A || !gp->B || !gp->C) { fs2_set_error("fs2_gemm_tc: null pointer"); return FS2_ERR_ARG; }
  const Fs2Gemm& g = *gp;
  if (!g.ab_bf16) { fs2_set_error("fs2_gemm_tc: operands must be bf16"); return FS2_ERR_UNSUPPORTED; }
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return FS2_OK;
  int rc = get_encode();
  if (rc) return rc;
  constexpr int BN = 128;
  TcParams p;
  memset(&p, 0, sizeof p);
  p.g = g;
  p.kb_per_tap = (g.K + BK - 1) / BK;
  p.total_kb = (g.mode == 2) ? p.kb_per_tap : p.kb_per_tap * g.taps;
  p.nsplit = (g.mode == 2 && g.split_k > 1) ? g.split_k : 1;
  if (p.nsplit > p.total_kb) p.nsplit = p.total_kb;
  p.kb_per_split = (p.total_kb + p.nsplit - 1) / p.nsplit;
  p.nsplit = (p.total_kb + p.kb_per_split - 1) / p.kb_per_split;
  if (p.nsplit > 1 && g.c_bf16) { fs2_set_error("fs2_gemm_tc: split_k needs fp32 C"); return FS2_ERR_ARG; }
  p.ntiles_per_tap = (g.N + BN - 1) / BN;
  int* errp = nullptr;
  CUDA_CHECK_RET(cudaGetSymbolAddress((void**)&errp, g_tc_error));
  p.err = errp;
  CUtensorMap ta, tb;
  // A: mode 0/1 K-major box (64 k, 128 rows); mode 2 MN-major box (64 m, 64 k-rows)
  if (g.mode == 2) rc = make_map(g.A, g.a_inner, g.a_rows, g.batch1, g.batch2, g.lda, g.a_s1, g.a_s2, 64, BK, &ta, p.pa);
  else rc = make_map(g.A, g.a_inner, g.a_rows, g.batch1, g.batch2, g.lda, g.a_s1, g.a_s2, BK, BM, &ta, p.pa);
  if (rc) return rc;
  if (g.mode == 0) rc = make_map(g.B, g.b_inner, g.b_rows, g.batch1, g.batch2, g.ldb, g.b_s1, g.b_s2, BK, BN, &tb, p.pb);
  else rc = make_map(g.B, g.b_inner, g.b_rows, g.batch1, g.batch2, g.ldb, g.b_s1, g.b_s2, 64, BK, &tb, p.pb);
  if (rc) return rc;
  int ntiles = (g.mode == 2) ? p.ntiles_per_tap * g.taps : p.ntiles_per_tap;
  dim3 grid(ntiles, (g.M + BM - 1) / BM, g.batch1 * g.batch2 * p.nsplit);
  cudaStream_t st = (cudaStream_t)stream;
  if (g.mode == 0) return launch<0, BN, 3>(ta, tb, p, grid, st);
  if (g.mode == 1) return launch<1, BN, 3>(ta, tb, p, grid, st);
  if (g.mode == 2) return launch<2, BN, 3>(ta, tb, p, grid, st);
  fs2_set_error("fs2_gemm_tc: bad mode");
  return FS2_ERR_ARG;
}
