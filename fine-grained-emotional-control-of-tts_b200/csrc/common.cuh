// Shared device helpers for the fs2_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>

#define FS2_OK 0
#define FS2_ERR_ARG 1
#define FS2_ERR_CUDA 2
#define FS2_ERR_UNSUPPORTED 3

// Activation rows live in one flat "padded row space": every (B, T, C) activation is
// stored as [B * (T + 2*FS2_PAD), C]; row (b, t) sits at b*(T+2*FS2_PAD) + FS2_PAD + t.
// The FS2_PAD rows either side hold the reflect halo (conv inputs) or zeros (gradients),
// which turns every "same"-padded Conv1d into a plain shifted-row GEMM.
#define FS2_PAD 4

typedef __nv_bfloat16 bf16;

#define CUDA_CHECK_RET(x)                                   \
  do {                                                      \
    cudaError_t e__ = (x);                                  \
    if (e__ != cudaSuccess) { fs2_set_error(cudaGetErrorString(e__)); return FS2_ERR_CUDA; } \
  } while (0)

void fs2_set_error(const char* msg);
int fs2_check_launch();

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------------------
// Every kernel of the library starts with pdl_wait(): it blocks until the preceding kernel of the stream has completed
// and its writes are visible, then lets the NEXT kernel begin launching.  Together with the launch attribute set by
// fs2_launch() this overlaps a kernel's launch latency, block scheduling and prologue (barrier init, TMEM allocation,
// tensor-map prefetch) with the tail of its predecessor -- the step is ~400 short kernels, 3-6 us of ramp/tail each
// (profiles/r01_summary.md).  Works inside CUDA-graph captures (programmatic edges).  FS2_PDL=0 switches it off.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

extern int g_fs2_pdl;      // core.cu

template <typename... KArgs, typename... Args>
inline void fs2_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_fs2_pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);      // errors surface in fs2_check_launch()
}
#define FS2_LAUNCH(kern, grid, block, smem, stream, ...) fs2_launch(kern, dim3(grid), dim3(block), (size_t)(smem), stream, __VA_ARGS__)

template <typename T> struct ActT;
template <> struct ActT<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct ActT<bf16> {
  static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// 4-wide vector access helpers (T = float -> 16 B, T = bf16 -> 8 B)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y);
  __nv_bfloat162 b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Counter-based dropout RNG.  keep(i) is reproducible from (seed, element index), so the backward regenerates the mask
// instead of storing it.  One 32-bit hash per PAIR of consecutive elements gives two 16-bit uniforms:
//     h = mix32(pair * 0x9E3779B1 + s0, s1),   element 2*pair keeps iff (h & 0xFFFF) >= p * 65536, element 2*pair + 1 iff
//     (h >> 16) >= p * 65536,
// with two key words (s0, s1) derived once per kernel from the 64-bit seed.  mix32 is a two-round multiply / xor-shift
// finaliser with the second key word added between the rounds (9 integer instructions per pair; the 64-bit mix per four
// elements this replaces cost ~45 and made the fused GEMM + LayerNorm epilogue issue-bound).
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct DropCfg {
  float p;             // drop probability (0 => disabled)
  unsigned long long seed;  // already mixed with a per-call stream id
};
struct DropKey {
  uint32_t s0, s1;     // key words
  uint32_t thr;        // drop iff the 16-bit uniform < thr = p * 65536
  float ks;            // 1 / (1 - p)
  float p;
};
__device__ __forceinline__ DropKey drop_key(float p, unsigned long long seed) {
  const uint64_t m = mix64(seed);
  DropKey k;
  k.s0 = (uint32_t)m;
  k.s1 = (uint32_t)(m >> 32);
  k.thr = (uint32_t)(p * 65536.0f);
  k.ks = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  k.p = p;
  return k;
}
// 32 random bits of element pair `pair` (= element index / 2)
__device__ __forceinline__ uint32_t drop_bits2(const DropKey& k, uint32_t pair) {
  uint32_t x = pair * 0x9E3779B1u + k.s0;
  x ^= x >> 16;
  x = x * 0x21F0AAADu + k.s1;
  x ^= x >> 15;
  x *= 0x735A2D97u;
  x ^= x >> 15;
  return x;
}
// returns 4 keep-scales (0 or 1/(1-p)) for elements [4*g, 4*g+4)
__device__ __forceinline__ float4 drop_scale4(const DropKey& d, uint64_t group) {
  if (d.p <= 0.f) return make_float4(1.f, 1.f, 1.f, 1.f);
  const uint32_t a = drop_bits2(d, (uint32_t)group * 2u), b = drop_bits2(d, (uint32_t)group * 2u + 1u);
  float4 o;
  o.x = ((a & 0xFFFFu) >= d.thr) ? d.ks : 0.f;
  o.y = ((a >> 16) >= d.thr) ? d.ks : 0.f;
  o.z = ((b & 0xFFFFu) >= d.thr) ? d.ks : 0.f;
  o.w = ((b >> 16) >= d.thr) ? d.ks : 0.f;
  return o;
}

// Attention-probability dropout (nn.MultiheadAttention's dropout on the softmax output).  keep(t, c) of query t and key
// c is decided by the top bits of a 32-bit PRODUCT of one well-mixed odd word per query and one per key:
//     keep = (row(t) * col(c) mod 2^32) >= p * 2^32 ,   row(t) = mix64(key ^ 2t) | 1 ,  col(c) = mix64(key ^ (2c + 1)) | 1
// -- one IMAD and one compare per element once the words are at hand.  The forward and the dQ kernel hold row(t) in a
// register and read col(.) from a table in shared memory, the dK/dV kernel the other way round, so all kernels (and the
// unfused softmax kernels) regenerate the same mask.  Measured against numpy's generator on 800 x 800 masks: same drop
// rate, same maximum row-pair / column-pair correlation (0.18-0.20), flat 2-D spectrum (tests/test_flash_attention_gpu.py).
struct ADrop {
  uint64_t key;   // of this (launch, item, head)
  uint32_t thr;   // drop iff product < thr = p * 2^32
  float ks;       // 1 / (1 - p)
};
__device__ __forceinline__ ADrop adrop_make(float p, unsigned long long seed, const unsigned long long* seed_dev, int bh) {
  ADrop d;
  if (seed_dev) seed ^= mix64(*seed_dev);
  d.key = mix64(seed ^ (0xA24BAED4963EE407ull * (uint64_t)(bh + 1)));
  d.thr = p > 0.f ? (uint32_t)fminf(p * 4294967296.0f, 4294967040.0f) : 0u;
  d.ks = 1.0f / (1.0f - p);
  return d;
}
__device__ __forceinline__ uint32_t adrop_row(const ADrop& d, int t) { return (uint32_t)mix64(d.key ^ (uint64_t)(2 * t)) | 1u; }
__device__ __forceinline__ uint32_t adrop_col(const ADrop& d, int c) { return (uint32_t)mix64(d.key ^ (uint64_t)(2 * c + 1)) | 1u; }
__device__ __forceinline__ bool adrop_keep(uint32_t row, uint32_t col, uint32_t thr) { return row * col >= thr; }

// keep-scales (0 or 1/(1-p)) of keys c .. c+3 of one query row (slow form: the column words are computed on the spot)
__device__ __forceinline__ float4 adrop_scale4(const ADrop& d, uint32_t row, int c) {
  if (d.thr == 0u) return make_float4(1.f, 1.f, 1.f, 1.f);
  float4 o;
  o.x = adrop_keep(row, adrop_col(d, c), d.thr) ? d.ks : 0.f;
  o.y = adrop_keep(row, adrop_col(d, c + 1), d.thr) ? d.ks : 0.f;
  o.z = adrop_keep(row, adrop_col(d, c + 2), d.thr) ? d.ks : 0.f;
  o.w = adrop_keep(row, adrop_col(d, c + 3), d.thr) ? d.ks : 0.f;
  return o;
}

// (b, t) <-> flat padded row helpers
struct RowSpace {
  int T;      // valid rows per batch item
  int Tp;     // T + 2*FS2_PAD
};
__device__ __forceinline__ bool row_decode(int r, int Tp, int T, int& b, int& t) {
  b = r / Tp;
  t = r - b * Tp - FS2_PAD;
  return t >= 0 && t < T;
}
