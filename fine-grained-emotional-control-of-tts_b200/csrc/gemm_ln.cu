// fs2_gemm_ln_tc: GEMM + bias + branch dropout + residual + LayerNorm in one tcgen05 kernel.
//
// The post-norm FFT block of the reference (speechbrain TransformerEncoderLayer, reached from model.py:344-347 and
// 425-428) runs  y = LN(x + dropout(a . W^T + b))  twice per layer: out-projection -> norm1 and FFN conv 2 -> norm2.  As
// two launches the fp32 branch (a . W^T + b) makes a round trip through HBM: 8 of the 18 bytes per element the pair moves.
// Here a CTA PAIR (2-CTA cluster, tcgen05 cta_group::2) owns a 256 x 384 tile; each CTA holds 128 COMPLETE LayerNorm rows
// in its tensor memory (384 of the 512 columns, fp32):
//   warp 0   : TMA producer   (own A 128 x 64 box + its HALF of W per k-block: 40 KB per CTA instead of 64 -- the
//              single-CTA form was bound by the L2 -> SM traffic of re-reading W, 1.9 k cycles per k-block for 0.77 k of MMA)
//   warp 1   : TMEM allocator; in the leader CTA the single-thread tcgen05.mma issuer (two M = 256, N = 192 MMAs per
//              K = 16 step, completion multicast to both CTAs)
//   warps 2-9: epilogue.  Each warp owns 16 rows (its TMEM lane quadrant, upper or lower half) and reads them in the
//              16x256b register layout: a quad of lanes holds 8 consecutive columns of a row (row statistics = quad shuffles,
//              a lane's column pairs suit the packed fp32x2 instructions).  Nothing goes through the load/store unit to
//              global memory on the fast path -- in this layout an instruction touches 8 lines x 32 bytes, which bounded
//              the first versions: the residual arrives through a per-warp bulk-tensor ring (3 slots x 4 KB, 128B swizzle)
//              and is read back with conflict-free ld.shared; the outputs go through per-warp swizzled staging tiles and
//              bulk-tensor stores (warps that hold halo / mirror rows store from registers).
//              While the tile's MMAs run: ring loads + the dropout keep bits of the lane's 2 x 96 elements (six words).
//              pass 1: z = x + keep * (acc + bias) written back to tensor memory, row sums;  pass 2: centred squares;
//              pass 3: normalise, scale / shift, store fp32 + bf16 (+ reflect-halo mirror rows).
// The accumulator is handed back to the MMA thread after the last tensor-memory read of pass 3, so the next tile's MMAs
// overlap the final stores.  It cannot be double-buffered (2 x 384 columns > 512): what overlaps a tile's epilogue is the
// TMA ring (the next tile's first 3 k-blocks) and the other SMs' main loops.  History and cycle breakdown:
// profiles/r02_summary.md section 6.
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "../../include/fs2_b200.h"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int LN_N = 384;
constexpr int BNS = 192;                      // UMMA N (<= 256): two MMAs cover the row
constexpr int STAGES = 3;
constexpr int NCTA = 2;                       // CTA pair
constexpr int NUM_EPI_WARPS = 8;
constexpr int NTHREADS = 32 * (2 + NUM_EPI_WARPS);
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int BSUB_ROWS = BNS / NCTA;         // W rows (output columns) of a 192-wide sub-tile this CTA stages
constexpr int BSUB_BYTES = BSUB_ROWS * BK * 2;            // 12 KB
constexpr int STAGE_BYTES = A_BYTES + 2 * BSUB_BYTES;     // 40 KB per CTA
// per epilogue warp: a staging tile the bulk-tensor engine stores from -- one step (64 columns) of the warp's 16 rows as
// two fp32 boxes (16 x 128 B each) and one bf16 box (16 x 128 B), 128B-swizzled
constexpr int OUT_BOX_BYTES = 16 * 128;
constexpr int OUT_STAGE_BYTES = 3 * OUT_BOX_BYTES;                 // 6 KB, double-buffered
// the same 12 KB per warp serve pass 1 as a ring of XSLOTS residual slots (one step = two fp32 boxes = 4 KB each) filled by
// bulk-tensor loads; pass 1 is over before pass 3 writes the first staging tile
constexpr int XSLOTS = 3;
constexpr int XSLOT_BYTES = 2 * OUT_BOX_BYTES;
static_assert(XSLOTS * XSLOT_BYTES == 2 * OUT_STAGE_BYTES, "residual ring and output staging share one buffer");
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NUM_EPI_WARPS * 2 * OUT_STAGE_BYTES + 1024 + 512;
constexpr int NCHUNK = LN_N / 32;             // 32-column chunks per row
constexpr int CH = 2;                         // chunks per epilogue step
static_assert(NCHUNK % CH == 0, "chunks per step");
static_assert(CH == 2, "the staging tile holds one bf16 box of 64 columns and two fp32 boxes of 32");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// cycle probes of CTA 0 (tools/gemm_ln_bench.py GEMM_LN_DBG=1) are compiled in with `make probe` only
#ifdef FS2_TC_PROBE
constexpr bool kProbe = true;
#else
constexpr bool kProbe = false;
#endif

struct LnGemmParams {
  Fs2GemmLn g;
  int M, m_tiles, kb;
  int* err;
  long long* dbg;
  int dbg_mode;      // probe build only (results incomplete): pass-3 experiments, see fs2_gemm_ln_set_debug
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

// ---- CTA-pair plumbing (same protocol as tcx_gemm_kernel in gemm_tc.cu)
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
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// TMA load whose completion bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_pair(uint32_t dst, const CUtensorMap* tm, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %5}], [%2];"
      ::"r"(dst), "l"(tm), "r"(cluster_bar), "r"(c0), "r"(c1), "r"(0)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// MMA-completion arrive on the same barrier offset in both CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

// two accumulator columns of one row as a float2 (the 16x256b layout hands a lane pairs of consecutive columns), for the
// packed fp32x2 arithmetic of sm_100 (FADD2 / FMUL2 / FFMA2: one instruction per column pair, IEEE-identical per element)
__device__ __forceinline__ float2 f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

__device__ __forceinline__ float2 ldg2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }

// bulk-tensor store shared -> global of one box (coordinates: column in the map's element type, row)
__device__ __forceinline__ void tma_store_box(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %4}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(0)
               : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

template <bool DROP>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_ln_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmF,
                                                               const __grid_constant__ CUtensorMap tmH,
                                                               const __grid_constant__ CUtensorMap tmX,
                                                               const LnGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* out_stage = smem + STAGES * STAGE_BYTES;                       // 1024-aligned, 6 KB per epilogue warp
  uint64_t* full_bar = (uint64_t*)(out_stage + NUM_EPI_WARPS * 2 * OUT_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;      // accumulator ready
  uint64_t* tempty_bar = tfull_bar + 1;          // accumulator drained
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);
  uint64_t* xbar = (uint64_t*)(tmem_slot + 2);   // [NUM_EPI_WARPS][XSLOTS] residual slot filled

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Fs2GemmLn& g = p.g;
  int* err = p.err;
  const uint32_t rank = cluster_ctarank();
  const int unit = (int)cluster_id_x(), nunits = (int)cluster_count_x();

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);      // the leader's producer arrives once (+ the pair's transaction bytes)
      mbar_init(smem_u32(&empty_bar[s]), 1);     // one MMA commit (multicast to both CTAs)
    }
    mbar_init(smem_u32(tfull_bar), 1);
    mbar_init(smem_u32(tempty_bar), NUM_EPI_WARPS * NCTA);
    for (int i = 0; i < NUM_EPI_WARPS * XSLOTS; ++i) mbar_init(smem_u32(&xbar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (both CTAs of the pair)
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      const uint32_t smem0 = smem_u32(smem);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t s = 0, ph = 0;
      bool ok = true;
      for (int t = unit; t < p.m_tiles && ok; t += nunits) {
        const int m0 = t * (BM * NCTA) + (int)rank * BM;
        for (int kb = 0; kb < p.kb; ++kb) {
          if (!mbar_wait(empty0 + 8 * s, ph ^ 1, err)) { ok = false; break; }
          const uint32_t fb_local = full0 + 8 * s;
          const uint32_t fb = mapa_rank(fb_local, 0);               // bytes of both CTAs are credited to the leader
          if (rank == 0) mbar_expect_tx(fb_local, STAGE_BYTES * NCTA);
          const uint32_t sa = smem0 + s * STAGE_BYTES;
          tma_load_pair(sa, &tmA, fb, kb * BK, m0);
          tma_load_pair(sa + A_BYTES, &tmB, fb, kb * BK, (int)rank * BSUB_ROWS);
          tma_load_pair(sa + A_BYTES + BSUB_BYTES, &tmB, fb, kb * BK, BNS + (int)rank * BSUB_ROWS);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA, one thread)
    if (rank == 0 && elect_one()) {
      // instruction descriptor: D = f32, A = B = bf16, both K-major, N >> 3, M >> 4 (M = 256 across the pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BNS >> 3) << 17) | ((uint32_t)((BM * NCTA) >> 4) << 24);
      const uint32_t smem0 = smem_u32(smem);
      const uint64_t adesc0 = smem_desc(smem0, 16, 1024);
      const uint64_t bdesc0 = smem_desc(smem0 + A_BYTES, 16, 1024);
      constexpr uint64_t KSTEP = 32 >> 4;                 // 16 bf16 = 32 B inside the 128 B swizzle row
      constexpr uint64_t STAGE_STEP = STAGE_BYTES >> 4;
      constexpr uint64_t SUB_STEP = BSUB_BYTES >> 4;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t s = 0, ph = 0, tc = 0;
      bool ok = true;
      for (int t = unit; t < p.m_tiles && ok; t += nunits, ++tc) {
        if (!mbar_wait(smem_u32(tempty_bar), (tc & 1) ^ 1, err)) { ok = false; break; }
        tc_fence_after();
        uint32_t acc = 0;
        for (int kb = 0; kb < p.kb; ++kb) {
          if (!mbar_wait(full0 + 8 * s, ph, err)) { ok = false; break; }
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)s * STAGE_STEP, bd = bdesc0 + (uint64_t)s * STAGE_STEP;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            umma_bf16_pair(tmem_base, ad + k * KSTEP, bd + k * KSTEP, idesc, acc);
            umma_bf16_pair(tmem_base + BNS, ad + k * KSTEP, bd + SUB_STEP + k * KSTEP, idesc, acc);
            acc = 1;
          }
          umma_commit_pair(empty0 + 8 * s);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (ok) umma_commit_pair(smem_u32(tfull_bar));
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
    // branch dropout: the mask of fs2_ln_fwd / fs2_ln_bwd (common.cuh:drop_bits2) -- one 32-bit hash per pair of
    // consecutive elements, which is exactly what a lane holds per row and 8-column group
    const DropKey dk = drop_key(g.drop_p, g.drop_seed ^ (g.seed_dev ? mix64(*g.seed_dev) : 0ull));
    const uint32_t dthr = dk.thr;
    const float dks = dk.ks;
    uint32_t tc = 0;
    bool ok = true;
    const bool prof = kProbe && p.dbg != nullptr && blockIdx.x == 0 && warp == 5 && lane == 0;
    long long pt[6] = {0, 0, 0, 0, 0, 0};
    const uint32_t tempty_leader = mapa_rank(smem_u32(tempty_bar), 0);
    for (int t = unit; t < p.m_tiles && ok; t += nunits, ++tc) {
      if (prof) pt[0] = clock64();
      const int row0 = t * (BM * NCTA) + (int)rank * BM + q * 32 + hf * 16;
      const RowInfo ra = row_info(row0 + tr, p.M, g.T, g.halo), rb = row_info(row0 + tr + 8, p.M, g.T, g.halo);
      // Residual rows of this warp: bulk-tensor loads (16 rows x 32 fp32 columns per box, 128B-swizzled) into the warp's
      // ring while the tile's MMAs run; the lanes read them back in the accumulator layout with conflict-free ld.shared.
      // As plain global loads in that layout they cost eight load/store-unit wavefronts per instruction (32 bytes of eight
      // different lines) -- ~1 k cycles per step for the SM's eight warps, the bound of pass 1 -- and had to live in
      // registers a step ahead.
      const uint32_t wbuf = smem_u32(out_stage) + (uint32_t)(warp - 2) * (2 * OUT_STAGE_BYTES);
      const uint32_t xb0 = smem_u32(&xbar[(warp - 2) * XSLOTS]);
      if (elect_one()) {
        // the previous tile's bulk stores read the bytes the ring occupies
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
        for (int k = 0; k < XSLOTS; ++k) {
          mbar_expect_tx(xb0 + 8 * k, XSLOT_BYTES);
          tma_load_4d(wbuf + k * XSLOT_BYTES, &tmX, xb0 + 8 * k, (2 * k) * 64, row0, 0, 0);
          tma_load_4d(wbuf + k * XSLOT_BYTES + OUT_BOX_BYTES, &tmX, xb0 + 8 * k, (2 * k + 1) * 64, row0, 0, 0);
        }
      }
      __syncwarp();
      // CH chunks (32 columns each) per step.  Everything a step needs from memory (bias or gamma / beta) and the step's
      // dropout words are requested / computed as one unrolled batch BEFORE the wait on the tensor-memory load: with two
      // epilogue warps per scheduler there is no other latency hiding.
      const uint32_t xlane = (uint32_t)tr * 128u + 8u * (uint32_t)(tq & 1);
      // Dropout keep bits of this lane's 2 x 96 elements, computed NOW -- while the tile's MMAs run and the epilogue warps
      // have nothing else to do -- and packed into six words: the hashes are ~210 integer instructions per lane and
      // 64-column step, and the integer pipe (half rate) made them ~1 k cycles per step when they sat inside pass 1.
      // Word w of a row covers steps 2w and 2w + 1; bit 16 (step & 1) + 2 i + e is column 64 step + 8 i + 2 tq + e.
      uint32_t km_a0 = 0u, km_a1 = 0u, km_a2 = 0u, km_b0 = 0u, km_b1 = 0u, km_b2 = 0u;
      if (DROP) {
        const uint32_t pa0 = (uint32_t)ra.r * (LN_N / 2) + (uint32_t)tq, pb0 = (uint32_t)rb.r * (LN_N / 2) + (uint32_t)tq;
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          uint32_t ma = 0u, mb = 0u;
#pragma unroll
          for (int k = 0; k < 16; ++k) {           // k = 8 (step & 1) + i: element pair index 32 step + 4 i + tq of the row
            const uint32_t ha = drop_bits2(dk, pa0 + (uint32_t)(64 * w + 4 * k));
            const uint32_t hb = drop_bits2(dk, pb0 + (uint32_t)(64 * w + 4 * k));
            ma |= ((ha & 0xFFFFu) >= dthr ? 1u : 0u) << (2 * k);
            ma |= ((ha >> 16) >= dthr ? 1u : 0u) << (2 * k + 1);
            mb |= ((hb & 0xFFFFu) >= dthr ? 1u : 0u) << (2 * k);
            mb |= ((hb >> 16) >= dthr ? 1u : 0u) << (2 * k + 1);
          }
          if (w == 0) { km_a0 = ma; km_b0 = mb; }
          if (w == 1) { km_a1 = ma; km_b1 = mb; }
          if (w == 2) { km_a2 = ma; km_b2 = mb; }
        }
        // pin the computation here: without a use before the wait below the compiler sinks it behind the wait
        asm volatile("" ::"r"(km_a0), "r"(km_a1), "r"(km_a2), "r"(km_b0), "r"(km_b1), "r"(km_b2));
      }
      if (prof) pt[1] = clock64();
      if (!mbar_wait(smem_u32(tfull_bar), tc & 1, err)) { ok = false; break; }
      tc_fence_after();
      if (prof) pt[2] = clock64();
      // ---- pass 1: z = x + keep * (acc + bias) -> tensor memory; row sums
      float2 sa2 = make_float2(0.f, 0.f), sb2 = make_float2(0.f, 0.f);     // even / odd column partial sums
#pragma unroll 1
      for (int c = 0; c < NCHUNK; c += CH) {
        uint32_t r[CH][16];
#pragma unroll
        for (int j = 0; j < CH; ++j) tmem_ld_16x256b_x4(tw + (uint32_t)((c + j) * 32), r[j]);
        float2 b2[4 * CH];
#pragma unroll
        for (int i = 0; i < 4 * CH; ++i) b2[i] = (kProbe && (p.dbg_mode & 8)) ? make_float2(0.f, 0.f) : ldg2(g.bias + c * 32 + 8 * i + 2 * tq);
        const bool pp = prof && tc == 0 && c == 4;
        if (prof && tc == 0) p.dbg[22 + c / CH] = clock64();       // start of every pass-1 step
        if (pp) p.dbg[8] = clock64();
        // this step's keep bits: 16 per row
        const int stp = c / CH;
        const uint32_t wa_ = (stp < 2 ? km_a0 : stp < 4 ? km_a1 : km_a2) >> ((stp & 1) * 16);
        const uint32_t wb_ = (stp < 2 ? km_b0 : stp < 4 ? km_b1 : km_b2) >> ((stp & 1) * 16);
        if (pp) p.dbg[9] = clock64();
        // this step's residual: slot (step % 3), filled for the (step / 3)-th time in this tile
        const int step = c / CH;
        const uint32_t xslot = wbuf + (uint32_t)(step % XSLOTS) * XSLOT_BYTES;
        if (!(kProbe && (p.dbg_mode & 32) && step >= XSLOTS) &&
            !mbar_wait(xb0 + 8 * (step % XSLOTS), (uint32_t)(step / XSLOTS) & 1u, err)) { ok = false; break; }
        float2 xc[8 * CH];
#pragma unroll
        for (int i = 0; i < 4 * CH; ++i) {
          const uint32_t a = xslot + (uint32_t)(i >> 2) * OUT_BOX_BYTES + xlane + ((((uint32_t)(2 * (i & 3)) + (uint32_t)(tq >> 1)) ^ (uint32_t)tr) << 4);
          xc[2 * i] = lds64(a);
          xc[2 * i + 1] = lds64(a + 8u * 128u);
        }
        tmem_wait_ld();
        if (pp) p.dbg[10] = clock64();
#pragma unroll
        for (int j = 0; j < CH; ++j) {
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int i = 4 * j + n;
            float ka0 = 1.f, ka1 = 1.f, kb0 = 1.f, kb1 = 1.f;
            if (DROP) {
              ka0 = (wa_ & (1u << (2 * i))) ? dks : 0.f;
              ka1 = (wa_ & (2u << (2 * i))) ? dks : 0.f;
              kb0 = (wb_ & (1u << (2 * i))) ? dks : 0.f;
              kb1 = (wb_ & (2u << (2 * i))) ? dks : 0.f;
            }
            const float2 za = __ffma2_rn(__fadd2_rn(f2(r[j][4 * n], r[j][4 * n + 1]), b2[i]), make_float2(ka0, ka1), xc[2 * i]);
            const float2 zb = __ffma2_rn(__fadd2_rn(f2(r[j][4 * n + 2], r[j][4 * n + 3]), b2[i]), make_float2(kb0, kb1), xc[2 * i + 1]);
            sa2 = __fadd2_rn(sa2, za);
            sb2 = __fadd2_rn(sb2, zb);
            r[j][4 * n] = __float_as_uint(za.x);
            r[j][4 * n + 1] = __float_as_uint(za.y);
            r[j][4 * n + 2] = __float_as_uint(zb.x);
            r[j][4 * n + 3] = __float_as_uint(zb.y);
          }
          tmem_st_16x256b_x4(tw + (uint32_t)((c + j) * 32), r[j]);
        }
        if (pp) p.dbg[11] = clock64();
        // every lane has consumed the slot (its values went into the tensor-memory store above): refill it
        __syncwarp();
        if (step + XSLOTS < NCHUNK / CH && !(kProbe && (p.dbg_mode & 32)) && elect_one()) {
          const int k = step % XSLOTS, s2 = step + XSLOTS;
          mbar_expect_tx(xb0 + 8 * k, XSLOT_BYTES);
          tma_load_4d(xslot, &tmX, xb0 + 8 * k, (2 * s2) * 64, row0, 0, 0);
          tma_load_4d(xslot + OUT_BOX_BYTES, &tmX, xb0 + 8 * k, (2 * s2 + 1) * 64, row0, 0, 0);
        }
        if (pp) p.dbg[13] = clock64();
      }
      if (!ok) break;
      if (prof && tc == 0) p.dbg[28] = clock64();
      tmem_wait_st();
      if (prof) pt[3] = clock64();
      if (prof && tc == 0) p.dbg[29] = pt[3];
      float sa = sa2.x + sa2.y, sb = sb2.x + sb2.y;
      sa += __shfl_xor_sync(0xffffffffu, sa, 1);
      sa += __shfl_xor_sync(0xffffffffu, sa, 2);
      sb += __shfl_xor_sync(0xffffffffu, sb, 1);
      sb += __shfl_xor_sync(0xffffffffu, sb, 2);
      const float mean_a = sa * invC, mean_b = sb * invC;
      // ---- pass 2: centred sum of squares
      float2 qa2 = make_float2(0.f, 0.f), qb2 = make_float2(0.f, 0.f);
      const float2 nma = make_float2(-mean_a, -mean_a), nmb = make_float2(-mean_b, -mean_b);
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
            const float2 da = __fadd2_rn(f2(r[j][4 * n], r[j][4 * n + 1]), nma), db = __fadd2_rn(f2(r[j][4 * n + 2], r[j][4 * n + 3]), nmb);
            qa2 = __ffma2_rn(da, da, qa2);
            qb2 = __ffma2_rn(db, db, qb2);
          }
        }
      }
      float qa = qa2.x + qa2.y, qb = qb2.x + qb2.y;
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
      if (prof) pt[4] = clock64();
      // ---- pass 3: normalise, scale / shift, store
      // A warp whose 16 rows are all rectangle rows without halo mirrors (all but ~5 % at mel-frame lengths) writes each
      // 64-column step into its swizzled staging tile and lets the bulk-tensor engine store the three boxes: the
      // load/store unit sees two shared-memory wavefronts per instruction where the same registers stored to global
      // memory cost eight (one per row: 32 or 16 bytes of a 128-byte line each) -- that was the epilogue's bound.
      const float2 rsa = make_float2(rstd_a, rstd_a), rsb = make_float2(rstd_b, rstd_b);
      const bool fast = __all_sync(0xffffffffu, ra.ok && rb.ok && (ra.m1 | ra.m2 | rb.m1 | rb.m2) == 0);
      if (fast) {
        const uint32_t stw = smem_u32(out_stage) + (uint32_t)(warp - 2) * (2 * OUT_STAGE_BYTES);
        // 128B swizzle: 16-byte chunk k of row r sits at r * 128 + ((k ^ (r & 7)) << 4); rows tr and tr + 8 share r & 7
        const uint32_t tsw = (uint32_t)tr;
        const int row0w = row0;
#pragma unroll 1
        for (int c = 0; c < NCHUNK; c += CH) {
          uint32_t r[CH][16];
#pragma unroll
          for (int j = 0; j < CH; ++j) tmem_ld_16x256b_x4(tw + (uint32_t)((c + j) * 32), r[j]);
          float2 g2[4 * CH], be2[4 * CH];
#pragma unroll
          for (int i = 0; i < 4 * CH; ++i) {
            g2[i] = (kProbe && (p.dbg_mode & 8)) ? make_float2(1.f, 1.f) : ldg2(g.gamma + c * 32 + 8 * i + 2 * tq);
            be2[i] = (kProbe && (p.dbg_mode & 8)) ? make_float2(0.f, 0.f) : ldg2(g.beta + c * 32 + 8 * i + 2 * tq);
          }
          const bool pp = prof && tc == 0 && c == 4;
          if (pp) p.dbg[16] = clock64();
          tmem_wait_ld();
          if (pp) p.dbg[17] = clock64();
          if (c + CH >= NCHUNK) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader);
          }
          // two staging tiles in turn: the engine must have read the step before the previous one (bulk groups belong to the elected lane)
          const uint32_t st0 = stw + (uint32_t)((c / CH) & 1) * OUT_STAGE_BYTES;
          const uint32_t rowa = st0 + (uint32_t)tr * 128u;
          // (nothing is pending in the first two steps: the tile began with wait_group.read 0)
          if (c >= 2 * CH) {
            if (!(kProbe && (p.dbg_mode & 4)) && elect_one()) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
          }
          if (pp) p.dbg[18] = clock64();
#pragma unroll
          for (int j = 0; j < CH; ++j) {
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              const int i = 4 * j + n;
              // (z - mean) * rstd * gamma + beta, the expression order of fs2_ln_fwd
              const float2 ya = __ffma2_rn(__fmul2_rn(__fadd2_rn(f2(r[j][4 * n], r[j][4 * n + 1]), nma), rsa), g2[i], be2[i]);
              const float2 yb = __ffma2_rn(__fmul2_rn(__fadd2_rn(f2(r[j][4 * n + 2], r[j][4 * n + 3]), nmb), rsb), g2[i], be2[i]);
              const float ya0 = ya.x, ya1 = ya.y, yb0 = yb.x, yb1 = yb.y;
              const __nv_bfloat162 pa = __floats2bfloat162_rn(ya0, ya1), pb = __floats2bfloat162_rn(yb0, yb1);
              // fp32 box j: byte 32 n + 8 tq of the row -> chunk 2n + tq/2
              const uint32_t fo = (uint32_t)j * OUT_BOX_BYTES + ((((uint32_t)(2 * n) + (uint32_t)(tq >> 1)) ^ tsw) << 4) + 8u * (uint32_t)(tq & 1);
              sts64(rowa + fo, ya0, ya1);
              sts64(rowa + fo + 8u * 128u, yb0, yb1);
              // bf16 box: byte 64 j + 16 n + 4 tq of the row -> chunk 4j + n
              const uint32_t ho = 2u * OUT_BOX_BYTES + (((uint32_t)(4 * j + n) ^ tsw) << 4) + 4u * (uint32_t)tq;
              sts32(rowa + ho, *reinterpret_cast<const uint32_t*>(&pa));
              sts32(rowa + ho + 8u * 128u, *reinterpret_cast<const uint32_t*>(&pb));
            }
          }
          if (pp) p.dbg[19] = clock64();
          if (!(kProbe && (p.dbg_mode & 16))) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (pp) p.dbg[20] = clock64();
          if (!(kProbe && (p.dbg_mode & 1)) && elect_one()) {   // the same lane every time (lowest of the full mask): bulk groups are per thread
            tma_store_box(&tmF, st0, c * 64, row0w);                       // fp32 tensor addressed as 768 16-bit columns
            tma_store_box(&tmF, st0 + OUT_BOX_BYTES, (c + 1) * 64, row0w);
            tma_store_box(&tmH, st0 + 2 * OUT_BOX_BYTES, c * 32, row0w);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (pp) p.dbg[21] = clock64();
        }
      } else {
      // rows outside the rectangle: zeros or nothing; reflect-halo mirror rows: plain stores from the registers
      const bool wa = ra.ok || ra.zero, wb = rb.ok || rb.zero;
      const bool mirrors = __any_sync(0xffffffffu, (ra.m1 | ra.m2 | rb.m1 | rb.m2) != 0);
      float* fa = g.out_f32 + ra.off + 2 * tq;
      float* fb = g.out_f32 + rb.off + 2 * tq;
      bf16* ha = oa + ra.off + 2 * tq;
      bf16* hb = oa + rb.off + 2 * tq;
#pragma unroll 1
      for (int c = 0; c < NCHUNK; c += CH) {
        uint32_t r[CH][16];
#pragma unroll
        for (int j = 0; j < CH; ++j) tmem_ld_16x256b_x4(tw + (uint32_t)((c + j) * 32), r[j]);
        float2 g2[4 * CH], be2[4 * CH];
#pragma unroll
        for (int i = 0; i < 4 * CH; ++i) {
          g2[i] = ldg2(g.gamma + c * 32 + 8 * i + 2 * tq);
          be2[i] = ldg2(g.beta + c * 32 + 8 * i + 2 * tq);
        }
        tmem_wait_ld();
        if (c + CH >= NCHUNK) {
          // last tensor-memory read of this warp: hand the accumulator back to the MMA thread
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_leader);
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) {
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const int i = 4 * j + n;
            const int col = (c + j) * 32 + 8 * n;
            float2 ya, yb;
            ya = __ffma2_rn(__fmul2_rn(__fadd2_rn(f2(r[j][4 * n], r[j][4 * n + 1]), nma), rsa), g2[i], be2[i]);
            yb = __ffma2_rn(__fmul2_rn(__fadd2_rn(f2(r[j][4 * n + 2], r[j][4 * n + 3]), nmb), rsb), g2[i], be2[i]);
            ya.x = ra.ok ? ya.x : 0.f;           // halo rows no mirror reaches get zeros
            ya.y = ra.ok ? ya.y : 0.f;
            yb.x = rb.ok ? yb.x : 0.f;
            yb.y = rb.ok ? yb.y : 0.f;
            const __nv_bfloat162 pa = __floats2bfloat162_rn(ya.x, ya.y), pb = __floats2bfloat162_rn(yb.x, yb.y);
            if (wa) {
              *reinterpret_cast<float2*>(fa + col) = ya;
              *reinterpret_cast<__nv_bfloat162*>(ha + col) = pa;
            }
            if (wb) {
              *reinterpret_cast<float2*>(fb + col) = yb;
              *reinterpret_cast<__nv_bfloat162*>(hb + col) = pb;
            }
            if (mirrors) {                       // warp-uniform: only the first / last rows of an item (halo > 0)
              if (ra.m1) {
                *reinterpret_cast<float2*>(fa + ra.m1 + col) = ya;
                *reinterpret_cast<__nv_bfloat162*>(ha + ra.m1 + col) = pa;
              }
              if (ra.m2) {
                *reinterpret_cast<float2*>(fa + ra.m2 + col) = ya;
                *reinterpret_cast<__nv_bfloat162*>(ha + ra.m2 + col) = pa;
              }
              if (rb.m1) {
                *reinterpret_cast<float2*>(fb + rb.m1 + col) = yb;
                *reinterpret_cast<__nv_bfloat162*>(hb + rb.m1 + col) = pb;
              }
              if (rb.m2) {
                *reinterpret_cast<float2*>(fb + rb.m2 + col) = yb;
                *reinterpret_cast<__nv_bfloat162*>(hb + rb.m2 + col) = pb;
              }
            }
          }
        }
      }
      }
      if (prof && tc == 0) {
        pt[5] = clock64();
        p.dbg[0] = pt[1] - pt[0];      // row setup, L2 prefetch, first residual loads issued
        p.dbg[1] = pt[2] - pt[1];      // wait for the tile's MMAs
        p.dbg[2] = pt[3] - pt[2];      // pass 1
        p.dbg[3] = pt[4] - pt[3];      // pass 2
        p.dbg[4] = pt[5] - pt[4];      // pass 3
      }
    }
    // the staging tile must outlive the engine's reads; the stores themselves complete before the grid does
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();          // nobody leaves while the peer can still signal its barriers
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

long long* g_ln_gemm_dbg = nullptr;
int g_ln_gemm_dbg_mode = 0;

}  // namespace

extern "C" int fs2_gemm_ln_set_debug(long long* dev_buf) {
  g_ln_gemm_dbg = dev_buf;
  const char* e = getenv("GEMM_LN_DBG_MODE");     // bit 0: no bulk stores, bit 1: no staging writes, bit 2: no staging-reuse wait
  g_ln_gemm_dbg_mode = e ? atoi(e) : 0;
  return FS2_OK;
}

#define REQUIRE(cond, msg) do { if (!(cond)) { fs2_set_error(msg); return FS2_ERR_ARG; } } while (0)

extern "C" int fs2_gemm_ln_tc(const Fs2GemmLn* gp, void* stream) {
  REQUIRE(gp && gp->A && gp->W && gp->bias && gp->x && gp->gamma && gp->beta, "fs2_gemm_ln_tc: null pointer");
  REQUIRE(gp->out_f32 && gp->out_act, "fs2_gemm_ln_tc: both outputs (fp32 and bf16) are required");
  const Fs2GemmLn& g = *gp;
  REQUIRE(g.B > 0 && g.T > 0 && g.K > 0 && g.K % 8 == 0, "fs2_gemm_ln_tc: bad shape (K must be a multiple of 8)");
  REQUIRE(g.halo >= 0 && g.halo <= FS2_PAD && (g.halo == 0 || g.T > g.halo), "fs2_gemm_ln_tc: halo too wide for T");
  REQUIRE(g.drop_p >= 0.f && g.drop_p < 1.f, "fs2_gemm_ln_tc: bad dropout probability");
  REQUIRE((((uintptr_t)g.bias | (uintptr_t)g.gamma | (uintptr_t)g.beta) % 8) == 0 &&
              (((uintptr_t)g.x | (uintptr_t)g.out_f32 | (uintptr_t)g.out_act) % 16) == 0,
          "fs2_gemm_ln_tc: parameter vectors must be 8-byte aligned, residual and outputs 16-byte aligned");
  const long long Ml = (long long)g.B * (g.T + 2 * FS2_PAD);
  REQUIRE(Ml * LN_N < (1LL << 40) && Ml < 0x7FFFFFFF, "fs2_gemm_ln_tc: too many rows");
  LnGemmParams p;
  memset(&p, 0, sizeof p);
  p.g = g;
  p.M = (int)Ml;
  p.m_tiles = (p.M + BM * NCTA - 1) / (BM * NCTA);
  p.kb = (g.K + BK - 1) / BK;
  p.dbg = g_ln_gemm_dbg;
  p.dbg_mode = g_ln_gemm_dbg_mode;
  int rc = fs2_tc_error_ptr(&p.err);
  if (rc) return rc;
  CUtensorMap ta, tb;
  if ((rc = fs2_tc_make_map_2d(g.A, g.K, p.M, g.lda, BK, BM, &ta))) return rc;
  if ((rc = fs2_tc_make_map_2d(g.W, g.K, LN_N, g.ldw, BK, BSUB_ROWS, &tb))) return rc;
  // output maps for the epilogue's bulk-tensor stores: 16-row boxes of 128 bytes; the fp32 tensor is described as
  // (M, 768) 16-bit elements (the engine moves bytes)
  CUtensorMap tf, th;
  if ((rc = fs2_tc_make_map_2d(g.out_f32, 2 * LN_N, p.M, 2 * LN_N, 64, 16, &tf))) return rc;
  if ((rc = fs2_tc_make_map_2d(g.out_act, LN_N, p.M, LN_N, 64, 16, &th))) return rc;
  CUtensorMap tx;              // the residual, read the same way
  if ((rc = fs2_tc_make_map_2d(g.x, 2 * LN_N, p.M, 2 * LN_N, 64, 16, &tx))) return rc;
  auto kern = g.drop_p > 0.f ? gemm_ln_kernel<true> : gemm_ln_kernel<false>;
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
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = (cudaStream_t)stream;
  cfg.attrs = attr;
  cfg.numAttrs = g_fs2_pdl ? 2 : 1;
  static int units = 0;       // CTA pairs that can be resident at once
  if (units == 0) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(gemm_ln_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CUDA_CHECK_RET(cudaFuncSetAttribute(gemm_ln_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int dev = 0, sms = 0, nc = 0;
    CUDA_CHECK_RET(cudaGetDevice(&dev));
    CUDA_CHECK_RET(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cfg.gridDim = dim3(sms / NCTA * NCTA, 1, 1);
    CUDA_CHECK_RET(cudaOccupancyMaxActiveClusters(&nc, gemm_ln_kernel<true>, &cfg));
    if (nc <= 0) { fs2_set_error("fs2_gemm_ln_tc: no resident CTA pair possible"); return FS2_ERR_CUDA; }
    units = nc < sms / NCTA ? nc : sms / NCTA;
  }
  const int nunits = p.m_tiles < units ? p.m_tiles : units;
  cfg.gridDim = dim3(nunits * NCTA, 1, 1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tf, th, tx, p);
  if (e != cudaSuccess) { fs2_set_error(cudaGetErrorString(e)); return FS2_ERR_CUDA; }
  return fs2_check_launch();
}
