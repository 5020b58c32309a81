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
// axis of each operand; tiles past the matrix edge are zero-filled. The K loop is software
// pipelined: the global loads of step k + 1 are issued into registers before the DMMAs of step k
// and stored into the other shared buffer after them (smaller CTA tiles for more CTAs measured slower).
#pragma once
#include <stdint.h>

#define CT_BM 64
#define CT_BK 16
#define CT_PAD 4

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Global -> register part of one K step's operand tiles (8 + 8 doubles per thread), and the
// register -> shared part. Splitting the two lets the loads of step k + 1 fly while the DMMAs of
// step k run (register double buffering; the shared tiles are double buffered too, one barrier per step).
template <int CT_BN>
__device__ __forceinline__ void contract_fetch(const double* __restrict__ X, const double* __restrict__ W, int64_t M, int N,
                                               int K, int64_t w_stride_k, int64_t w_stride_n, int64_t m0, int n0, int k0,
                                               int tid, double* xa, double* wb) {
#pragma unroll
  for (int q = 0; q < CT_BM * CT_BK / 128; ++q) {
    const int i = tid + 128 * q, r = i / CT_BK, c = i - r * CT_BK;
    const int64_t m = m0 + r;
    const int k = k0 + c;
    xa[q] = (m < M && k < K) ? X[m * K + k] : 0.0;
  }
#pragma unroll
  for (int q = 0; q < CT_BK * CT_BN / 128; ++q) {
    const int i = tid + 128 * q;
    int r, c;                                      // r = k, c = n; the contiguous axis of W varies fastest
    if (w_stride_n == 1) { r = i / CT_BN; c = i - r * CT_BN; }
    else { c = i / CT_BK; r = i - c * CT_BK; }
    const int k = k0 + r, n = n0 + c;
    wb[q] = (k < K && n < N) ? W[(int64_t)k * w_stride_k + (int64_t)n * w_stride_n] : 0.0;
  }
}

template <int CT_BN>
__device__ __forceinline__ void contract_stage(double (*As)[CT_BK + CT_PAD], double (*Bs)[CT_BK + CT_PAD],
                                               int64_t w_stride_n, int tid, const double* xa, const double* wb) {
#pragma unroll
  for (int q = 0; q < CT_BM * CT_BK / 128; ++q) {
    const int i = tid + 128 * q, r = i / CT_BK, c = i - r * CT_BK;
    As[r][c] = xa[q];
  }
#pragma unroll
  for (int q = 0; q < CT_BK * CT_BN / 128; ++q) {
    const int i = tid + 128 * q;
    int r, c;
    if (w_stride_n == 1) { r = i / CT_BN; c = i - r * CT_BN; }
    else { c = i / CT_BK; r = i - c * CT_BK; }
    Bs[c][r] = wb[q];
  }
}

template <int CT_BN>
__global__ void __launch_bounds__(128) rddl_contract_f64(const double* __restrict__ X, const double* __restrict__ W,
                                                         double* __restrict__ Out, int64_t M, int N, int K,
                                                         int64_t w_stride_k, int64_t w_stride_n) {
  __shared__ double As[2][CT_BM][CT_BK + CT_PAD];   // x tile: [env][k], double buffered
  __shared__ double Bs[2][CT_BN][CT_BK + CT_PAD];   // W tile, n-major: [n][k]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int TN = CT_BN / 16;                             // mma tiles per warp along n
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * (CT_BN / 2);    // warp origin inside the CTA tile
  const int64_t m0 = (int64_t)blockIdx.y * CT_BM;
  const int n0 = blockIdx.x * CT_BN;
  double acc[4][TN][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  double xa[CT_BM * CT_BK / 128], wb[CT_BK * CT_BN / 128];
  contract_fetch<CT_BN>(X, W, M, N, K, w_stride_k, w_stride_n, m0, n0, 0, tid, xa, wb);
  contract_stage<CT_BN>(As[0], Bs[0], w_stride_n, tid, xa, wb);
  __syncthreads();
  int cur = 0;
  for (int k0 = 0; k0 < K; k0 += CT_BK) {
    const bool more = k0 + CT_BK < K;
    if (more) contract_fetch<CT_BN>(X, W, M, N, K, w_stride_k, w_stride_n, m0, n0, k0 + CT_BK, tid, xa, wb);   // in flight during the DMMAs
#pragma unroll
    for (int kk = 0; kk < CT_BK; kk += 4) {
      double a[4], b[TN];
      // fragments (PTX ISA m8n8k4): A[row = lane/4][col = lane%4], B[row = lane%4][col = lane/4]
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[cur][wm + 8 * i + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[cur][wn + 8 * j + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (more) contract_stage<CT_BN>(As[cur ^ 1], Bs[cur ^ 1], w_stride_n, tid, xa, wb);
    __syncthreads();
    cur ^= 1;
  }
  // C fragment: row = lane/4, cols = (lane%4)*2 + {0, 1}
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + wm + 8 * i + (lane >> 2);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + wn + 8 * j + (lane & 3) * 2;
      if (n < N) Out[m * N + n] = acc[i][j][0];
      if (n + 1 < N) Out[m * N + n + 1] = acc[i][j][1];
    }
  }
}
