// Blackwell tensor-core path of the sum-product contraction  out(e, n) = sum_k x(e, k) * W(k, n)
// (the two-operand aggregation sum_{?a} W(?a, ?b) * x(?a) of pyRDDLGym/core/simulator.py:584-637,
// 766-779, for a non-fluent W shared by all envs): tcgen05.mma with TMA operand loads and the
// accumulator in tensor memory.
//
// RDDL reals are float64 and tcgen05 has no f64 kind, so this is the *tolerance* variant of the
// contraction (rddl_program_option(prog, 2, 1); the default stays the exact FP64 DMMA kernel of
// rddl_contract.cuh): every operand is split into three bfloat16 terms  v = hi + mid + lo
// (hi = bf16(v), mid = bf16(v - hi), lo = bf16(v - hi - mid): 24 significant bits), and the six
// products of combined order <= 2 -- hi*hi, hi*mid, mid*hi, mid*mid, hi*lo, lo*hi -- are accumulated
// in float32 in TMEM (the leading product and the five corrections in accumulators of their own,
// added in float64 at the end); what is dropped is O(2^-24) of |x||W| per term and the float32
// accumulation adds ~sqrt(K / 16) * 2^-24 relative to the size of the sum: measured < 1e-6 of
// max |out| at K = 512, i.e. inside the 1e-5 tolerance of the north star for reals wherever the result
// is not a near-total cancellation (tests/test_cuda_parity.py::test_tcgen05_split_bf16_contraction_vs_fp64).
//
// Two launches per contraction:
//   rddl_tc_split      f64 [rows, K] (strided) -> three bf16 planes in the tensor cores' K-major
//                      "core matrix" order [term][K/8][rows][8]   (x: every step; W: once per program)
//   rddl_contract_tc   CTA tile 128 (envs) x 64 (N), K in chunks of 64 through a 3-stage TMA ->
//                      mbarrier -> tcgen05.mma pipeline: lane 0 of warp 0 issues the TMA loads
//                      (cp.async.bulk.tensor.4d), lane 0 of warp 1 issues the MMAs (24 per chunk) and
//                      commits them to the stage's "empty" barrier; the four warps then read the
//                      128 x 64 float32 accumulator from TMEM (tcgen05.ld 32x32b), widen to float64
//                      and store.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_BM 128
#define TC_BN 64
#define TC_BK 64
#define TC_STAGES 3
#define TC_A_TERM (TC_BM * TC_BK * 2)      // bytes of one term's A tile: [8 k-columns][128 rows][16 B]
#define TC_B_TERM (TC_BN * TC_BK * 2)      //                     B tile: [8 k-columns][ 64 rows][16 B]
#define TC_STAGE_BYTES (3 * TC_A_TERM + 3 * TC_B_TERM)
#define TC_SMEM_BYTES (TC_STAGES * TC_STAGE_BYTES + 1024)

// f64 -> (hi, mid, lo) bfloat16, value = hi + mid + lo up to 2^-24 relative
__device__ __forceinline__ void tc_split3(double v, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(__double2float_rn(v));
  const double r1 = v - (double)__bfloat162float(h);
  m = __float2bfloat16_rn(__double2float_rn(r1));
  const double r2 = r1 - (double)__bfloat162float(m);
  l = __float2bfloat16_rn(__double2float_rn(r2));
}

// src(r, k) = src[r * stride_r + k * stride_k] for r < rows, k < K (zero outside) ->
// dst[term][k / 8][r][k % 8], r < rows_p, k < Kp. One thread per (k / 8, r), r fastest: 16-byte
// coalesced stores, full-sector reads.
__global__ void __launch_bounds__(256) rddl_tc_split(const double* __restrict__ src, int64_t stride_r, int64_t stride_k,
                                                     int64_t rows, int64_t K, int64_t rows_p, int64_t Kp,
                                                     __nv_bfloat16* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_p * (Kp >> 3)) return;
  const int64_t r = i % rows_p, ko = i / rows_p;
  __align__(16) __nv_bfloat16 h[8], m[8], l[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t k = ko * 8 + j;
    const double v = (r < rows && k < K) ? src[r * stride_r + k * stride_k] : 0.0;
    tc_split3(v, h[j], m[j], l[j]);
  }
  const int64_t plane = rows_p * Kp;
  uint4* d = reinterpret_cast<uint4*>(dst + i * 8);
  d[0] = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(dst + plane + i * 8) = *reinterpret_cast<const uint4*>(m);
  *reinterpret_cast<uint4*>(dst + 2 * plane + i * 8) = *reinterpret_cast<const uint4*>(l);
}

__device__ __forceinline__ uint32_t tc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem(bar)), "r"(bytes) : "memory");
}
// (bounded: a pipeline that never completes -- a rejected tensor map, say -- traps instead of hanging the GPU)
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = tc_smem(bar);
  for (uint32_t spins = 0;; ++spins) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (spins > (1u << 24)) __trap();
  }
}
// one 4-D TMA box -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tc_tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(tc_smem(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tc_smem(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: 8 x 16-byte core matrices; LBO = byte distance
// between the two core matrices of one MMA along K, SBO = between 8-row groups (cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem(bar)) : "memory");
}

__global__ void __launch_bounds__(128, 1)
rddl_contract_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 double* __restrict__ out, int64_t n_envs, int N, int nkb) {
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[TC_STAGES], empty[TC_STAGES], fin;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * TC_BN;

  if (tid == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { tc_mbar_init(&full[s], 1); tc_mbar_init(&empty[s], 1); }
    tc_mbar_init(&fin, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {      // two accumulators of 64 float32 columns x 128 lanes of tensor memory (see the MMA issuer)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&tmem_slot)), "n"(2 * TC_BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---- TMA producer: per K chunk, the three terms of the x tile and of the W tile
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_STAGES, it = kb / TC_STAGES;
      if (it > 0) tc_mbar_wait(&empty[s], (uint32_t)((it - 1) & 1));
      tc_mbar_expect_tx(&full[s], TC_STAGE_BYTES);
      uint8_t* st = tiles + (size_t)s * TC_STAGE_BYTES;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        tc_tma_load_4d(st + t * TC_A_TERM, &map_x, &full[s], 0, m0, kb * (TC_BK / 8), t);
        tc_tma_load_4d(st + 3 * TC_A_TERM + t * TC_B_TERM, &map_w, &full[s], 0, n0, kb * (TC_BK / 8), t);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer: D[128 x 64] (+)= A[128 x 16] * B[64 x 16]^T, bf16 x bf16 -> f32
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp): C = F32, A = B = BF16, K-major, N / 8, M / 16
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
    // The leading product hi*hi goes to accumulator 0, the five correction products (2^-8 and less of
    // it) to accumulator 1: the float32 rounding of a running sum is relative to its magnitude, so the
    // small terms are not swamped and the big accumulator takes K / 16 additions instead of 6 K / 16.
    const int pa[6] = {2, 0, 1, 1, 0, 0}, pb[6] = {0, 2, 1, 0, 1, 0};     // lo*hi, hi*lo, mid*mid, mid*hi, hi*mid | hi*hi
    uint32_t acc0 = 0, acc1 = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_STAGES, it = kb / TC_STAGES;
      tc_mbar_wait(&full[s], (uint32_t)(it & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = tc_smem(tiles + (size_t)s * TC_STAGE_BYTES), b0 = a0 + 3 * TC_A_TERM;
#pragma unroll
      for (int k16 = 0; k16 < TC_BK / 16; ++k16) {
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const uint64_t da = tc_desc(a0 + pa[q] * TC_A_TERM + k16 * 2 * (TC_BM * 16), TC_BM * 16, 128);
          const uint64_t db = tc_desc(b0 + pb[q] * TC_B_TERM + k16 * 2 * (TC_BN * 16), TC_BN * 16, 128);
          if (q == 5) { tc_mma(tmem, da, db, idesc, acc0); acc0 = 1; }
          else { tc_mma(tmem + TC_BN, da, db, idesc, acc1); acc1 = 1; }
        }
      }
      tc_commit(&empty[s]);          // arrives when the MMAs above have read this stage's tiles
    }
    tc_commit(&fin);                 // ... and when the accumulator is complete
  }
  __syncwarp();
  // ---- epilogue: thread t owns accumulator row t (TMEM lane t); 8 columns per tcgen05.ld
  tc_mbar_wait(&fin, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int64_t row = (int64_t)m0 + tid;
#pragma unroll 1
  for (int c = 0; c < TC_BN; c += 8) {
    uint32_t v[8], w[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(TC_BN + c)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row < n_envs) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (n0 + c + j < N) out[row * N + n0 + c + j] = (double)__uint_as_float(v[j]) + (double)__uint_as_float(w[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * TC_BN) : "memory");
}
