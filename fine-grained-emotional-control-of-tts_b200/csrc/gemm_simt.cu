// Exact SIMT GEMM for the Fs2Gemm descriptor (fp32 accumulate, fp32 or bf16 operands).
// This is the fp32 "exact" precision path and the small/odd-shape path (K or N not
// TMA-friendly).  The bf16 hot GEMMs go through gemm_tc.cu (tcgen05).
#include "common.cuh"
#include "../../include/fs2_b200.h"
#include "gemm_epilogue.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename TA>
__device__ __forceinline__ float fetch(const TA* base, long long ld, int row, int col, int rows, int inner) {
  if (row < 0 || row >= rows || col < 0 || col >= inner) return 0.f;
  return ActT<TA>::ld(base + (long long)row * ld + col);
}

template <typename TA>
__global__ void __launch_bounds__(256) gemm_simt_kernel(Fs2Gemm g) {
  pdl_wait();
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  const int z = blockIdx.z;
  const int nsplit = g.split_k > 1 ? g.split_k : 1;
  const int zb = z / nsplit, zs = z % nsplit;
  const int i1 = zb % g.batch1, i2 = zb / g.batch1;
  const TA* A = (const TA*)g.A + i1 * g.a_s1 + i2 * g.a_s2;
  const TA* Bm = (const TA*)g.B + i1 * g.b_s1 + i2 * g.b_s2;
  const int m0 = blockIdx.y * TM;
  const int Ntot = (g.mode == 2) ? g.N * g.taps : g.N;
  const int n0 = blockIdx.x * TN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  const int Kred = (g.mode == 2) ? g.K : g.K * g.taps;
  int kbeg = 0, kend = Kred;
  if (nsplit > 1) {
    int chunk = ((Kred + nsplit - 1) / nsplit + TK - 1) / TK * TK;
    kbeg = zs * chunk;
    kend = min(Kred, kbeg + chunk);
  }
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    // load A tile: TM x TK
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      int mm, kk;
      if (g.mode == 2) { mm = e % TM; kk = e / TM; } else { kk = e % TK; mm = e / TK; }
      int m = m0 + mm, kidx = k0 + kk;
      float v = 0.f;
      if (m < g.M && kidx < kend) {
        if (g.mode == 2) {
          v = fetch<TA>(A, g.lda, g.a_row_off + kidx, m, g.a_rows, g.a_inner);
        } else {
          int j = kidx / g.K, k = kidx - j * g.K;
          v = fetch<TA>(A, g.lda, g.a_row_off + m + j * g.a_tap_step, k, g.a_rows, g.a_inner);
        }
      }
      As[kk][mm] = v;
    }
    for (int e = threadIdx.x; e < TN * TK; e += 256) {
      int nn, kk;
      if (g.mode == 0) { kk = e % TK; nn = e / TK; } else { nn = e % TN; kk = e / TN; }
      int n = n0 + nn, kidx = k0 + kk;
      float v = 0.f;
      if (n < Ntot && kidx < kend) {
        if (g.mode == 0) {
          int j = kidx / g.K, k = kidx - j * g.K;
          v = fetch<TA>(Bm, g.ldb, n, j * g.b_tap_step + k, g.b_rows, g.b_inner);
        } else if (g.mode == 1) {
          int j = kidx / g.K, k = kidx - j * g.K;
          v = fetch<TA>(Bm, g.ldb, g.b_row_off + k, n + j * g.b_tap_step, g.b_rows, g.b_inner);
        } else {
          int j = n / g.N, c = n - j * g.N;
          v = fetch<TA>(Bm, g.ldb, g.b_row_off + kidx + j * g.b_tap_step, c, g.b_rows, g.b_inner);
        }
      }
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i) b[i] = Bs[kk][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[i][jj] = fmaf(a[i], b[jj], acc[i][jj]);
    }
    __syncthreads();
  }
  EpiRow er;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
    epi_row_setup(g, i1, i2, m, er);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      int n = n0 + tx * 4 + jj;
      if (n >= Ntot) continue;
      long long col;
      int nb;
      if (g.mode == 2) { int j = n / g.N; nb = n - j * g.N; col = (long long)j * g.c_tap_stride + (long long)nb * (g.c_col_stride > 1 ? g.c_col_stride : 1); }
      else { nb = n; col = n; }
      epi_store(g, er, col, nb, acc[i][jj], nsplit > 1);
    }
  }
}

}  // namespace

extern "C" int fs2_gemm_simt(const Fs2Gemm* g, void* stream) {
  if (!g || !g->A || !g->B || !g->C) { fs2_set_error("fs2_gemm_simt: null pointer"); return FS2_ERR_ARG; }
  if (g->M <= 0 || g->N <= 0 || g->K <= 0) return FS2_OK;
  int Ntot = (g->mode == 2) ? g->N * g->taps : g->N;
  int nsplit = g->split_k > 1 ? g->split_k : 1;
  if (nsplit > 1 && g->c_bf16) { fs2_set_error("fs2_gemm_simt: split_k needs fp32 C"); return FS2_ERR_ARG; }
  dim3 grid((Ntot + TN - 1) / TN, (g->M + TM - 1) / TM, g->batch1 * g->batch2 * nsplit);
  if (g->ab_bf16) FS2_LAUNCH((gemm_simt_kernel<bf16>), grid, 256, 0, (cudaStream_t)stream, *g);
  else FS2_LAUNCH((gemm_simt_kernel<float>), grid, 256, 0, (cudaStream_t)stream, *g);
  return fs2_check_launch();
}
