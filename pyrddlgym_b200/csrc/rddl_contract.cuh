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
//
// ASYNC variant (the default where the operands allow 16-byte copies: W n-contiguous, K and N even, 16-byte
// aligned bases): the operand tiles go from global to shared memory with cp.async (16 bytes = two doubles,
// zero-filled past the edges) through a CT_STAGES-deep ring, W k-major as it lies in memory ([k][n], rows
// padded by 4 doubles: the B fragment's four k rows x four n columns per half warp fall into 16 distinct
// 8-byte banks, like the A fragment's). Three K steps of loads are in flight behind the step being multiplied
// instead of one: the register-staged loop above spends ~1.1 us per K step of 16 at 512 x 512 (one CTA per SM,
// the loads' L2 latency only half hidden) where the DMMAs of the step need 0.5.
//
// Split K (gridDim.z = S > 1): at 512 x 512 x 1024 envs the 64 x 64 tiling gives 128 CTAs of four warps --
// one CTA per SM, 6 % warp occupancy, the DMMA pipe half idle behind its own latency. S CTAs per output
// tile each take a K range (multiples of 16), write their partial tile to a scratch buffer, and the CTA
// that arrives last at the tile's counter adds the S partials in range order and stores the result: the
// sum has a fixed order whatever the arrival order (deterministic), and the counter resets itself.
#pragma once
#include <stdint.h>

#define CT_BM 64
#define CT_BK 16
#define CT_PAD 4
#define CT_STAGES 4
#define CT_ASYNC_SMEM(BN) (CT_STAGES * (CT_BM * (CT_BK + CT_PAD) + CT_BK * ((BN) + CT_PAD)) * 8)

__device__ __forceinline__ void ct_cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int n = valid ? 16 : 0;            // (src-size 0: the 16 bytes are zero-filled)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

// One K step's operand tiles, global -> shared by cp.async: x tile [64 envs][16 k], W tile [16 k][BN n]
template <int CT_BN>
__device__ __forceinline__ void contract_issue(const double* __restrict__ X, const double* __restrict__ W, int64_t M, int N,
                                               int K, int kend, int64_t w_stride_k, int64_t m0, int n0, int k0, int tid,
                                               double (*As)[CT_BK + CT_PAD], double (*Bs)[CT_BN + CT_PAD]) {
#pragma unroll
  for (int q = 0; q < CT_BM * CT_BK / 2 / 128; ++q) {
    const int i = tid + 128 * q, r = i / (CT_BK / 2), c = (i - r * (CT_BK / 2)) * 2;
    const int64_t m = m0 + r;
    const int k = k0 + c;
    const bool ok = m < M && k < kend;
    ct_cp_async16(&As[r][c], ok ? X + m * K + k : X, ok);
  }
#pragma unroll
  for (int q = 0; q < CT_BK * CT_BN / 2 / 128; ++q) {
    const int i = tid + 128 * q, r = i / (CT_BN / 2), c = (i - r * (CT_BN / 2)) * 2;
    const int k = k0 + r, n = n0 + c;
    const bool ok = k < kend && n < N;
    ct_cp_async16(&Bs[r][c], ok ? W + (int64_t)k * w_stride_k + n : W, ok);
  }
}

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
                                               int K, int kend, int64_t w_stride_k, int64_t w_stride_n, int64_t m0, int n0,
                                               int k0, int tid, double* xa, double* wb) {
#pragma unroll
  for (int q = 0; q < CT_BM * CT_BK / 128; ++q) {
    const int i = tid + 128 * q, r = i / CT_BK, c = i - r * CT_BK;
    const int64_t m = m0 + r;
    const int k = k0 + c;
    xa[q] = (m < M && k < kend) ? X[m * K + k] : 0.0;      // (kend: end of this CTA's K range)
  }
#pragma unroll
  for (int q = 0; q < CT_BK * CT_BN / 128; ++q) {
    const int i = tid + 128 * q;
    int r, c;                                      // r = k, c = n; the contiguous axis of W varies fastest
    if (w_stride_n == 1) { r = i / CT_BN; c = i - r * CT_BN; }
    else { c = i / CT_BK; r = i - c * CT_BK; }
    const int k = k0 + r, n = n0 + c;
    wb[q] = (k < kend && n < N) ? W[(int64_t)k * w_stride_k + (int64_t)n * w_stride_n] : 0.0;
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

template <int CT_BN, bool ASYNC = false>
__global__ void __launch_bounds__(128) rddl_contract_f64(const double* __restrict__ X, const double* __restrict__ W,
                                                         double* __restrict__ Out, int64_t M, int N, int K,
                                                         int64_t w_stride_k, int64_t w_stride_n,
                                                         double* __restrict__ partial = nullptr,
                                                         unsigned int* __restrict__ arrived = nullptr, int k_per_split = 0) {
  // register-staged loop: x tile [env][k] and W tile n-major [n][k], double buffered (static);
  // cp.async loop: CT_STAGES x { x tile [env][k], W tile k-major [k][n] } in dynamic shared memory
  __shared__ double As[ASYNC ? 1 : 2][ASYNC ? 1 : CT_BM][CT_BK + CT_PAD];
  __shared__ double Bs[ASYNC ? 1 : 2][ASYNC ? 1 : CT_BN][CT_BK + CT_PAD];
  extern __shared__ __align__(16) double ct_dyn[];
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

  // this CTA's K range (everything without split K)
  const int S = (int)gridDim.z;
  const int kb = S > 1 ? (int)blockIdx.z * k_per_split : 0;
  const int ke = S > 1 ? min(K, kb + k_per_split) : K;
  if constexpr (ASYNC) {
    typedef double (*ATile)[CT_BK + CT_PAD];
    typedef double (*BTile)[CT_BN + CT_PAD];
    constexpr int kA = CT_BM * (CT_BK + CT_PAD), kB = CT_BK * (CT_BN + CT_PAD);
    auto a_of = [&](int st) { return reinterpret_cast<ATile>(ct_dyn + st * (kA + kB)); };
    auto b_of = [&](int st) { return reinterpret_cast<BTile>(ct_dyn + st * (kA + kB) + kA); };
    const int nsteps = (ke - kb + CT_BK - 1) / CT_BK;
#pragma unroll
    for (int p = 0; p < CT_STAGES - 1; ++p) {
      if (p < nsteps) contract_issue<CT_BN>(X, W, M, N, K, ke, w_stride_k, m0, n0, kb + p * CT_BK, tid, a_of(p), b_of(p));
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int it = 0; it < nsteps; ++it) {
      asm volatile("cp.async.wait_group %0;" ::"n"(CT_STAGES - 2) : "memory");     // step `it` has landed (this thread's part)
      __syncthreads();      // ... everybody's; and the stage refilled below was read by all in iteration it - 1
      const int nx = it + CT_STAGES - 1;
      if (nx < nsteps)
        contract_issue<CT_BN>(X, W, M, N, K, ke, w_stride_k, m0, n0, kb + nx * CT_BK, tid, a_of(nx % CT_STAGES), b_of(nx % CT_STAGES));
      asm volatile("cp.async.commit_group;" ::: "memory");
      const ATile Ac = a_of(it % CT_STAGES);
      const BTile Bc = b_of(it % CT_STAGES);
#pragma unroll
      for (int kk = 0; kk < CT_BK; kk += 4) {
        double a[4], b[TN];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = Ac[wm + 8 * i + (lane >> 2)][kk + (lane & 3)];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bc[kk + (lane & 3)][wn + 8 * j + (lane >> 2)];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else {
  double xa[CT_BM * CT_BK / 128], wb[CT_BK * CT_BN / 128];
  contract_fetch<CT_BN>(X, W, M, N, K, ke, w_stride_k, w_stride_n, m0, n0, kb, tid, xa, wb);
  contract_stage<CT_BN>(As[0], Bs[0], w_stride_n, tid, xa, wb);
  __syncthreads();
  int cur = 0;
  for (int k0 = kb; k0 < ke; k0 += CT_BK) {
    const bool more = k0 + CT_BK < ke;
    if (more) contract_fetch<CT_BN>(X, W, M, N, K, ke, w_stride_k, w_stride_n, m0, n0, k0 + CT_BK, tid, xa, wb);   // in flight during the DMMAs
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
  }
  if (S > 1) {
    // partial tiles: [tile][split][fragment slot][thread] -- a thread's slots are coalesced across the CTA
    __shared__ int last;
    const int64_t tile = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    double* mine = partial + (tile * S + blockIdx.z) * (CT_BM * CT_BN);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        mine[((i * TN + j) * 2 + 0) * 128 + tid] = acc[i][j][0];
        mine[((i * TN + j) * 2 + 1) * 128 + tid] = acc[i][j][1];
      }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
      const unsigned int t = atomicAdd(arrived + tile, 1u);
      last = t == (unsigned int)(S - 1);
      if (last) arrived[tile] = 0u;       // (every split has arrived: ready for the next launch)
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    const double* all = partial + tile * S * (CT_BM * CT_BN);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          double v = 0.0;
          for (int sp = 0; sp < S; ++sp) v += __ldcg(all + (int64_t)sp * (CT_BM * CT_BN) + ((i * TN + j) * 2 + h) * 128 + tid);
          acc[i][j][h] = v;
        }
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
