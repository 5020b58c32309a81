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
// in float32 in TMEM (every product in an accumulator of its own, added in float64 at the end); what is dropped is O(2^-24) of |x||W| per term and the float32
// accumulation adds ~sqrt(K / 16) * 2^-24 relative to the size of the sum: measured < 1e-6 of
// max |out| at K = 512, i.e. inside the 1e-5 tolerance of the north star for reals wherever the result
// is not a near-total cancellation (tests/test_cuda_parity.py::test_tcgen05_split_bf16_contraction_vs_fp64).
//
// Two launches per contraction:
//   rddl_tc_split      f64 [rows, K] (strided) -> three bf16 planes, row-major K-contiguous
//                      [term][rows][K]   (x: every step; W: once per program)
//   rddl_contract_tc   CTA tile 128 (envs) x BN (N; 32 when 64 would leave SMs without a CTA), K in
//                      chunks of 64 through a 3-stage TMA -> mbarrier -> tcgen05.mma pipeline: lane 0
//                      of warp 0 issues the TMA loads (cp.async.bulk.tensor.3d, one box of
//                      [rows x 64 bf16] = rows x 128 B per term, written to shared memory in the
//                      128-byte-swizzled K-major layout the MMA descriptors name: 8x fewer TMA rows than
//                      the unswizzled 16-byte core-matrix layout of the first version, which was
//                      bound by the TMA row rate -- 31 us at 512 x 512 x 1024 envs), lane 0 of warp 1
//                      issues the MMAs (12 per chunk: N = 3 BN, 2 BN, BN against the x terms hi, mid, lo) and
//                      commits them to the stage's "empty" barrier; the four warps then read the six
//                      128 x BN float32 accumulators from TMEM (tcgen05.ld 32x32b.x16), add them in float64
//                      and store.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_BM 128
#define TC_BN_MAX 64
#define TC_BK 64
#define TC_STAGES 3
#define TC_A_TERM (TC_BM * TC_BK * 2)          // bytes of one term's A tile: [128 rows][128 B], 8-row groups of 1 KiB
#define TC_B_TERM(BN) ((BN) * TC_BK * 2)       //                     B tile: [BN rows][128 B]
#define TC_STAGE_BYTES(BN) (3 * TC_A_TERM + 3 * TC_B_TERM(BN))
#define TC_SMEM_BYTES(BN) (TC_STAGES * TC_STAGE_BYTES(BN) + 1024)

// f64 -> (hi, mid, lo) bfloat16, value = hi + mid + lo up to 2^-24 relative
__device__ __forceinline__ void tc_split3(double v, __nv_bfloat16& h, __nv_bfloat16& m, __nv_bfloat16& l) {
  h = __float2bfloat16_rn(__double2float_rn(v));
  const double r1 = v - (double)__bfloat162float(h);
  m = __float2bfloat16_rn(__double2float_rn(r1));
  const double r2 = r1 - (double)__bfloat162float(m);
  l = __float2bfloat16_rn(__double2float_rn(r2));
}

// src(r, k) = src[r * stride_r + k * stride_k] for r < rows, k < K (zero outside) ->
// dst[term][r][k], r < rows_p, k < Kp. One thread per (r, k / 8), k fastest: 64-byte reads and
// 16-byte stores, both contiguous across the warp when stride_k == 1 (the per-step split of x).
__global__ void __launch_bounds__(256) rddl_tc_split(const double* __restrict__ src, int64_t stride_r, int64_t stride_k,
                                                     int64_t rows, int64_t K, int64_t rows_p, int64_t Kp,
                                                     __nv_bfloat16* __restrict__ dst) {
  // programmatic dependent launch: let the contraction kernel behind this one set up (barriers, tensor
  // memory, tensor-map prefetch) while the split runs; it waits (griddepcontrol.wait) before its first TMA load
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows_p * (Kp >> 3)) return;
  const int64_t kw = Kp >> 3, r = i / kw, ko = i - r * kw;
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
// one 3-D TMA box (k, row, term) -> shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tc_tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(tc_smem(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tc_smem(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// shared-memory matrix descriptor, K-major, 128-byte swizzle (cute/arch/mma_sm100_desc.hpp: layout type 2,
// descriptor version 1): a tile row is 64 bf16 = 128 B, 8 rows form one 1 KiB swizzle atom (the tile base is
// 1 KiB aligned), SBO = 1024 B between 8-row groups, LBO unused. A K step of 16 elements inside the atom is
// +32 B on the start address (the hardware applies the XOR swizzle to the address it computes).
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
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

template <int BN>
__global__ void __launch_bounds__(128, 1)
rddl_contract_tc(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 double* __restrict__ out, int64_t n_envs, int N, int nkb) {
  extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full[TC_STAGES], empty[TC_STAGES], fin;
  __shared__ uint32_t tmem_slot;
  constexpr uint32_t kStage = TC_STAGE_BYTES(BN), kBTerm = TC_B_TERM(BN);
  // six float32 accumulators of BN columns, one per product (192 or 384 columns; allocations are powers of two)
  constexpr uint32_t kCols = BN == 32 ? 256 : 512;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.y * TC_BM, n0 = blockIdx.x * BN;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) { tc_mbar_init(&full[s], 1); tc_mbar_init(&empty[s], 1); }
    tc_mbar_init(&fin, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
  if (warp == 0) {      // tensor memory for the accumulators (see the MMA issuer)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&tmem_slot)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;

  if (warp == 0 && lane == 0) {
    // ---- TMA producer: per K chunk, the three terms of the x tile and of the W tile
    asm volatile("griddepcontrol.wait;" ::: "memory");      // the split of x (the preceding launch) is complete and visible
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_STAGES, it = kb / TC_STAGES;
      if (it > 0) tc_mbar_wait(&empty[s], (uint32_t)((it - 1) & 1));
      tc_mbar_expect_tx(&full[s], kStage);
      uint8_t* st = tiles + (size_t)s * kStage;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        tc_tma_load_3d(st + t * TC_A_TERM, &map_x, &full[s], kb * TC_BK, m0, t);
        tc_tma_load_3d(st + 3 * TC_A_TERM + t * kBTerm, &map_w, &full[s], kb * TC_BK, n0, t);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer. The three terms of the W tile lie back to back in shared memory ([term][BN rows][128 B]), so
    // "the first j terms" is itself a K-major operand of j * BN rows: per 16-wide K step
    //   x_hi  [128 x 16] * (W_hi | W_mid | W_lo)^T -> columns [0, 3 BN)        hi*hi, hi*mid, hi*lo
    //   x_mid [128 x 16] * (W_hi | W_mid)^T        -> columns [3 BN, 5 BN)     mid*hi, mid*mid
    //   x_lo  [128 x 16] * (W_hi)^T                -> columns [5 BN, 6 BN)     lo*hi
    // i.e. the six products of combined order <= 2 in three MMAs instead of six: an MMA reads its whole A
    // tile from shared memory (4 KB = 32 cycles) however narrow N is, which is what bounded six N = BN
    // instructions (ncu: 55 cycles each at N = 32 against a 16-cycle tensor floor). Every product has a
    // float32 accumulator of its own: the float32 rounding of a running sum is relative to its magnitude, so
    // the small correction terms are not swamped by the leading one; they are added in float64 at the end.
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp): C = F32, A = B = BF16, K-major, N / 8, M / 16
    const uint32_t ibase = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    const uint32_t idesc3 = ibase | ((uint32_t)((3 * BN) >> 3) << 17), idesc2 = ibase | ((uint32_t)((2 * BN) >> 3) << 17),
                   idesc1 = ibase | ((uint32_t)(BN >> 3) << 17);
    uint32_t acc = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % TC_STAGES, it = kb / TC_STAGES;
      tc_mbar_wait(&full[s], (uint32_t)(it & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = tc_smem(tiles + (size_t)s * kStage), b0 = a0 + 3 * TC_A_TERM;
#pragma unroll
      for (int k16 = 0; k16 < TC_BK / 16; ++k16) {
        const uint64_t db = tc_desc(b0 + k16 * 32);
        tc_mma(tmem, tc_desc(a0 + k16 * 32), db, idesc3, acc);
        tc_mma(tmem + 3 * BN, tc_desc(a0 + TC_A_TERM + k16 * 32), db, idesc2, acc);
        tc_mma(tmem + 5 * BN, tc_desc(a0 + 2 * TC_A_TERM + k16 * 32), db, idesc1, acc);
        acc = 1;
      }
      tc_commit(&empty[s]);          // arrives when the MMAs above have read this stage's tiles
    }
    tc_commit(&fin);                 // ... and when the accumulators are complete
  }
  __syncwarp();
  // ---- epilogue: thread t owns accumulator row t (TMEM lane t); 16 columns of the six accumulators per round
  tc_mbar_wait(&fin, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int64_t row = (int64_t)m0 + tid;
#pragma unroll 1
  for (int c = 0; c < BN; c += 16) {
    uint32_t v[6][16];
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
#pragma unroll
    for (int q = 0; q < 6; ++q)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(v[q][0]), "=r"(v[q][1]), "=r"(v[q][2]), "=r"(v[q][3]), "=r"(v[q][4]), "=r"(v[q][5]), "=r"(v[q][6]),
                     "=r"(v[q][7]), "=r"(v[q][8]), "=r"(v[q][9]), "=r"(v[q][10]), "=r"(v[q][11]), "=r"(v[q][12]), "=r"(v[q][13]),
                     "=r"(v[q][14]), "=r"(v[q][15])
                   : "r"(ta + (uint32_t)(q * BN)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    // hi*hi + ((hi*mid + mid*hi) + (hi*lo + mid*mid + lo*hi)), in float64
    auto total = [&](int j) {
      const double first = (double)__uint_as_float(v[1][j]) + (double)__uint_as_float(v[3][j]);
      const double second = ((double)__uint_as_float(v[2][j]) + (double)__uint_as_float(v[4][j])) + (double)__uint_as_float(v[5][j]);
      return (double)__uint_as_float(v[0][j]) + (first + second);
    };
    if (row < n_envs) {
      double* o = out + row * N + n0 + c;
      if (n0 + c + 16 <= N && (N & 1) == 0) {      // whole 16-column run inside the tensor, 16-byte aligned rows
#pragma unroll
        for (int j = 0; j < 16; j += 2) *reinterpret_cast<double2*>(o + j) = make_double2(total(j), total(j + 1));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + c + j < N) o[j] = total(j);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kCols) : "memory");
}
