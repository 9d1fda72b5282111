// Output-side semantics shared by the SIMT and the tcgen05 GEMM kernels.
#pragma once
#include "common.cuh"
#include "../../include/fs2_b200.h"

// Fs2Gemm.relu: 0 = none, 1 = ReLU.  (A GELU variant lived here for the intensity extractor; its erff code alone made
// every GEMM kernel ~2 % slower on the FastSpeech2 step -- instruction footprint -- so GELU is a separate elementwise
// kernel, fs2_gelu.)
__device__ __forceinline__ float epi_act(float v, int kind) { return kind ? fmaxf(v, 0.f) : v; }

struct EpiRow {
  long long base;      // element offset of (row, c_col_off) in C
  long long mirror;    // element offset delta of the reflect-halo mirror row at the item's start (0 = none)
  long long mirror2;   // ... and at the item's end (a short item can need both)
  bool in_rect;        // row is a real (b,t) row of the rectangle
  bool live;           // in_rect and t < lens[b]
  bool skip;           // do not write this row at all
};

__device__ __forceinline__ void epi_row_setup(const Fs2Gemm& g, int i1, int i2, int m, EpiRow& er) {
  long long r = (long long)g.c_row_off + m;
  er.base = (long long)i1 * g.c_s1 + (long long)i2 * g.c_s2 + r * g.ldc + g.c_col_off;
  er.mirror = 0;
  er.mirror2 = 0;
  er.in_rect = true;
  er.live = true;
  er.skip = false;
  if (g.rs_Tp > 0) {
    int b = (int)(r / g.rs_Tp);
    int t = (int)(r - (long long)b * g.rs_Tp) - FS2_PAD;
    er.in_rect = (t >= 0 && t < g.rs_T);
    er.live = er.in_rect && (g.lens == nullptr || t < g.lens[b]);
    if (g.halo > 0) {
      if (!er.in_rect) er.skip = true;
      else {
        if (t >= 1 && t <= g.halo) er.mirror = -2LL * t * g.ldc;
        if (t >= g.rs_T - 1 - g.halo && t <= g.rs_T - 2) er.mirror2 = 2LL * (g.rs_T - 1 - t) * g.ldc;
      }
    }
  }
}

__device__ __forceinline__ float epi_value(const Fs2Gemm& g, const EpiRow& er, long long col, int nb, float acc) {
  float v = acc * g.alpha;
  if (g.bias) v += g.bias[nb];
  if (g.relu) v = epi_act(v, g.relu);
  if (g.relu_aux) {
    float a = g.aux_bf16 ? __bfloat162float(((const bf16*)g.relu_aux)[er.base + col])
                         : ((const float*)g.relu_aux)[er.base + col];
    if (!(a > 0.f)) v = 0.f;
  }
  if (!er.live) v = 0.f;
  return v;
}

__device__ __forceinline__ void epi_store(const Fs2Gemm& g, const EpiRow& er, long long col, int nb, float acc,
                                          bool atomic) {
  if (er.skip) return;
  float v = epi_value(g, er, col, nb, acc);
  long long o = er.base + col;
  if (g.c_bf16) {
    bf16 h = __float2bfloat16_rn(v);
    ((bf16*)g.C)[o] = h;
    if (er.mirror) ((bf16*)g.C)[o + er.mirror] = h;
    if (er.mirror2) ((bf16*)g.C)[o + er.mirror2] = h;
  } else if (atomic || g.accumulate) {
    atomicAdd(((float*)g.C) + o, v);
  } else {
    ((float*)g.C)[o] = v;
    if (er.mirror) ((float*)g.C)[o + er.mirror] = v;
    if (er.mirror2) ((float*)g.C)[o + er.mirror2] = v;
  }
}
