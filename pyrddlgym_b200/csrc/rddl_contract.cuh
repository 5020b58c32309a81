// FP64 tensor-core contraction for RDDL aggregations that are genuine two-operand contractions
// over a shared object axis with a non-fluent matrix operand:
//     out(e, n) = sum_k x(e, k) * W(k, n)          e = env, k = ?a, n = ?b
// i.e. `y'(?b) = sum_{?a} W(?a, ?b) * x(?a)` (the reference materialises the
// [?b, ?a] product and reduces it per env, simulator.py:584-637,766-779).
//
// RDDL reals are float64 and parity is judged at 1e-5 against a float64 reference, so the
// contraction runs on the FP64 tensor pipe (DMMA, mma.sync.m8n8k4.f64 -- tcgen05 has no f64
// kind): exact products, float64 accumulation, only the summation order differs from NumPy.
//
// CTA tile 64 (envs) x 64 (outputs), K step 16, 4 warps (2 x 2), each warp 32 x 32 = 4 x 4 mma
// tiles. Operands are staged in shared memory (x tile row-major, W tile transposed to n-major,
// rows padded by 4 doubles against bank conflicts); loads are coalesced along the contiguous
// axis of each operand; tiles past the matrix edge are zero-filled.
#pragma once
#include <stdint.h>

#define CT_BM 64
#define CT_BN 64
#define CT_BK 16
#define CT_PAD 4

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128) rddl_contract_f64(const double* __restrict__ X, const double* __restrict__ W,
                                                         double* __restrict__ Out, int64_t M, int N, int K,
                                                         int64_t w_stride_k, int64_t w_stride_n) {
  __shared__ double As[CT_BM][CT_BK + CT_PAD];   // x tile: [env][k]
  __shared__ double Bs[CT_BN][CT_BK + CT_PAD];   // W tile, n-major: [n][k]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;    // warp origin inside the CTA tile
  const int64_t m0 = (int64_t)blockIdx.y * CT_BM;
  const int n0 = blockIdx.x * CT_BN;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int k0 = 0; k0 < K; k0 += CT_BK) {
    // x tile: 64 rows x 16 doubles, k contiguous in memory
    for (int i = tid; i < CT_BM * CT_BK; i += 128) {
      const int r = i / CT_BK, c = i - r * CT_BK;
      const int64_t m = m0 + r;
      const int k = k0 + c;
      As[r][c] = (m < M && k < K) ? X[m * K + k] : 0.0;
    }
    // W tile: element (k, n) at k * stride_k + n * stride_n; iterate with the contiguous axis fastest
    if (w_stride_n == 1) {
      for (int i = tid; i < CT_BK * CT_BN; i += 128) {
        const int r = i / CT_BN, c = i - r * CT_BN;        // r = k, c = n
        const int k = k0 + r, n = n0 + c;
        Bs[c][r] = (k < K && n < N) ? W[(int64_t)k * w_stride_k + n] : 0.0;
      }
    } else {
      for (int i = tid; i < CT_BK * CT_BN; i += 128) {
        const int c = i / CT_BK, r = i - c * CT_BK;        // c = n, r = k
        const int k = k0 + r, n = n0 + c;
        Bs[c][r] = (k < K && n < N) ? W[(int64_t)k * w_stride_k + (int64_t)n * w_stride_n] : 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < CT_BK; kk += 4) {
      double a[4], b[4];
      // fragments (PTX ISA m8n8k4): A[row = lane/4][col = lane%4], B[row = lane%4][col = lane/4]
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[wm + 8 * i + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[wn + 8 * j + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();
  }
  // C fragment: row = lane/4, cols = (lane%4)*2 + {0, 1}
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + wm + 8 * i + (lane >> 2);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + wn + 8 * j + (lane & 3) * 2;
      if (n < N) Out[m * N + n] = acc[i][j][0];
      if (n + 1 < N) Out[m * N + n + 1] = acc[i][j][1];
    }
  }
}
