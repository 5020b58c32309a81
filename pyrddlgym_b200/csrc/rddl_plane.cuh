// Bit-plane interpreter: the fast path for launches that pyrddlgym_b200/plane_lowering.py
// proved to be boolean-grid work (every value of the launch is a function of a few bits per
// cell). Bool planes live in HBM as one byte per cell (the torch.bool / NumPy bool layout of
// the reference boundary, pyRDDLGym/core/compiler/initializer.py:16-23); inside the kernel a
// plane register is a packed 32-cell word, so one LOP3 evaluates a logical node for 32 cells.
//
// Tile: 256 threads x WPT words x 32 cells of consecutive (env, cell) index space (WPT = 4:
// 32768 cells; WPT = 1 for small batches so that the grid still covers the 148 SMs).
// Thread t owns words w = j*256 + t: every global access of a warp covers 1 KiB of contiguous
// bytes (32 B per thread, two 128-bit vector accesses), i.e. fully coalesced.
// The op table of the launch is pre-decoded on the host (PlaneProg, passed in the kernel
// parameter constant bank): truth tables arrive expanded to 32-bit masks, so a table lookup
// over m index planes is a (2^m - 1)-LOP3 multiplexer tree whose leaf masks are read
// straight from the constant bank.
//   PLD/PST   bytes <-> bits          PLDC    pre-packed non-fluent plane
//   PLUT      m-input truth table     PSHARE  stage a plane (+ halo rows) in shared memory
//   PCOUNT    3x3 stencil count as 4 bit-sliced planes (carry-save adders)
//   PBERN     Bernoulli(table[idx planes]) by bit-serial comparison against Philox words
//   PRED      per-env popcount reductions (reward sums)
#pragma once
#include "philox.cuh"
#include "rddl_plan.h"

#define PLANE_MAX_REGS 32
#define PLANE_MAX_OPS 40
#define PLANE_WORDS 1536
#define PLANE_BERN_LEVELS 8
#define PLANE_QUEUE 96

struct PlaneOp {        // 32 bytes, pre-decoded from RddlInsn + its Fin descriptor
  int32_t op, dst, a, b;
  int32_t m;            // index planes; bit 8: upper half of the table is constant
  uint32_t idx;         // index plane registers, 6 bits each
  int32_t off;          // offset of this op's tables in PlaneProg::words
  uint32_t imm;
};

struct PlaneProg {
  int32_t nops, nshared, w32_shift, wpe_shift;   // shifts: log2 of words per row / per env, or -1
  int32_t nfused, fused_nregs, rollout_ok, pad;  // scalar launches run in the epilogue (kind 2)
  int32_t pol_k, ntma;                           // device-side policy with a max-nondef rule: groundings kept per (env, step);
                                                 // ntma: pre-packed non-fluent planes staged in shared memory by TMA (PLDC slots)
  int32_t npol, pad2;                            // PPOL ops (device-side policy draws) in the program
  // thresholds / Philox node of the policy's Bernoulli entries (copied from the policy descriptor when the
  // program is rewritten for it: compile-time constants of the specialised build, like every other table)
  uint64_t pol_thr[RDDL_MAX_POLICY][2];          // [entry]: thr, thr_lo -- the draw is [thr_lo <= U < thr]
  uint32_t pol_node[RDDL_MAX_POLICY];
  int32_t pol_always[RDDL_MAX_POLICY];           // p >= 1
  RddlLaunch fused[2];
  int32_t red_off[RDDL_MAX_RED], red_m[RDDL_MAX_RED];   // PRED op of reduction j: table offset, #planes
  PlaneOp ops[PLANE_MAX_OPS];
  uint32_t words[PLANE_WORDS];
};

// Two build modes of this file. Ahead of time (nvcc, rddl_vm.cu): the program arrives in the kernel
// parameter bank and the op loop below is an interpreter. Specialised (PLANE_STATIC, compiled at
// program-create time with NVRTC, see rddl_jit.h): "plane_static_data.inc" defines the PlaneProg,
// the launch geometry and the tensor dtypes of ONE launch as constants (kG, kL, kDT, PS_*), the op
// loop unrolls, every switch folds, table masks become LOP3 immediates and the plane registers
// become machine registers.
#ifdef PLANE_STATIC
#include "plane_static_data.inc"
#define PLANE_OUTLINE __forceinline__
#define PLANE_UNROLL_DYN _Pragma("unroll")
#define PLANE_LAUNCH(A) kL
#if PS_FUSED_STATIC
// The fused scalar programs as a compile-time instruction sequence: every operand field is a
// constant, the register file r[] becomes machine registers, loads of the pools fold.
template <int PC, int END>
__device__ __forceinline__ void vm_exec_static(const VmArgs& A, const VmPools& P, const RddlLaunch& L, const int64_t env,
                                               const uint64_t step, int32_t* idx, uint64_t* redval, uint64_t* r,
                                               uint64_t* cap, const int cap_tensor, const uint64_t* red_local) {
  if constexpr (PC < END) {
    constexpr uint4 raw = {kInsnWords[PC - PS_INSN_BASE][0], kInsnWords[PC - PS_INSN_BASE][1],
                           kInsnWords[PC - PS_INSN_BASE][2], kInsnWords[PC - PS_INSN_BASE][3]};
    int pc = 0;
    vm_exec_one<40, (int)(raw.x & 0xff)>(A, P, L, env, step, idx, redval, r, nullptr, nullptr, nullptr, pc, raw, cap, cap_tensor,
                                         uint4{0u, 0u, 0u, 0u}, red_local, &kL);
    vm_exec_static<PC + 1, END>(A, P, L, env, step, idx, redval, r, cap, cap_tensor, red_local);
  }
}
#endif
#else
#define PLANE_OUTLINE __noinline__
#define PLANE_UNROLL_DYN _Pragma("unroll 1")
#define PLANE_LAUNCH(A) (A).L
#endif

// 32 bool bytes (each 0/1) -> 32-bit word, cell i -> bit i
__device__ __forceinline__ uint32_t plane_pack(const uint4 lo, const uint4 hi) {
  const uint32_t M = 0x01020408u;
  const uint32_t a = (lo.x + (lo.y << 4)) * M;   // top byte = cells 0..7
  const uint32_t b = (lo.z + (lo.w << 4)) * M;   // cells 8..15
  const uint32_t c = (hi.x + (hi.y << 4)) * M;   // cells 16..23
  const uint32_t d = (hi.z + (hi.w << 4)) * M;   // cells 24..31
  const uint32_t ab = __byte_perm(a, b, 0x0073);  // byte0 = a.byte3, byte1 = b.byte3
  const uint32_t cd = __byte_perm(c, d, 0x0073);
  return __byte_perm(ab, cd, 0x5410);
}

__device__ __forceinline__ uint32_t plane_nib(uint32_t n) { return (n * 0x00204081u) & 0x01010101u; }

__device__ __forceinline__ void plane_unpack(uint32_t w, uint4& lo, uint4& hi) {
  lo.x = plane_nib(w & 15u);          lo.y = plane_nib((w >> 4) & 15u);
  lo.z = plane_nib((w >> 8) & 15u);   lo.w = plane_nib((w >> 12) & 15u);
  hi.x = plane_nib((w >> 16) & 15u);  hi.y = plane_nib((w >> 20) & 15u);
  hi.z = plane_nib((w >> 24) & 15u);  hi.w = plane_nib(w >> 28);
}

// multiplexer tree: value of the table (2^M masks at mk[]) at the cell-wise index x[0..M)
template <int M>
__device__ __forceinline__ uint32_t plane_mux(const uint32_t* mk, const uint32_t* x) {
  if (M == 0) return mk[0];
  uint32_t v[M > 0 ? (1 << (M > 0 ? M - 1 : 0)) : 1];
#pragma unroll
  for (int i = 0; i < (1 << (M > 0 ? M - 1 : 0)); ++i) v[i] = (x[0] & mk[2 * i + 1]) | (~x[0] & mk[2 * i]);
#pragma unroll
  for (int k = 1; k < M; ++k) {
#pragma unroll
    for (int i = 0; i < (1 << (M - 1 - k)); ++i) v[i] = (x[k] & v[2 * i + 1]) | (~x[k] & v[2 * i]);
  }
  return v[0];
}

// same, for tables whose upper half is the constant mk[1 << (M-1)] (HC) or general tables
template <int M, bool HC>
__device__ __forceinline__ uint32_t plane_mux_hc(const uint32_t* mk, const uint32_t* x) {
  if (!HC || M == 0) return plane_mux<M>(mk, x);
  const uint32_t lo = plane_mux<(M > 0 ? M - 1 : 0)>(mk, x);
  return (x[M > 0 ? M - 1 : 0] & mk[1 << (M > 0 ? M - 1 : 0)]) | (~x[M > 0 ? M - 1 : 0] & lo);
}

// batched over NW words, leaf-major: every table mask is fetched once and applied to all NW words
template <int M, bool HC, int NW>
__device__ __forceinline__ void plane_mux_n(const uint32_t* mk, const uint32_t (*x)[M > 0 ? M : 1], uint32_t* out) {
  constexpr int ME = (HC && M > 0) ? M - 1 : M;      // depth of the non-constant part of the table
  if (ME == 0) {
#pragma unroll
    for (int jj = 0; jj < NW; ++jj) out[jj] = mk[0];
  } else {
    constexpr int NL = 1 << (ME > 0 ? ME - 1 : 0);
    uint32_t v[NW][NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
      const uint32_t m1 = mk[2 * i + 1], m0 = mk[2 * i];
#pragma unroll
      for (int jj = 0; jj < NW; ++jj) v[jj][i] = (x[jj][0] & m1) | (~x[jj][0] & m0);
    }
#pragma unroll
    for (int k = 1; k < ME; ++k) {
#pragma unroll
      for (int i = 0; i < (NL >> k); ++i) {
#pragma unroll
        for (int jj = 0; jj < NW; ++jj) v[jj][i] = (x[jj][k] & v[jj][2 * i + 1]) | (~x[jj][k] & v[jj][2 * i]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < NW; ++jj) out[jj] = v[jj][0];
  }
  if (HC && M > 0) {
    const uint32_t c = mk[1 << (M > 0 ? M - 1 : 0)];
#pragma unroll
    for (int jj = 0; jj < NW; ++jj)
      out[jj] = (x[jj][M > 0 ? M - 1 : 0] & c) | (~x[jj][M > 0 ? M - 1 : 0] & out[jj]);
  }
}

template <int WPT, int BLOCK>
struct PlaneCtx {
  // virtual plane registers live in shared memory, [reg][word][thread]: conflict-free, never
  // evicted (a local-memory register file of this size thrashes L1 at 3 CTAs per SM)
#ifdef PLANE_STATIC
  uint32_t R[PS_NREGS][WPT];   // every index is a compile-time constant after unrolling: registers
  __device__ __forceinline__ uint32_t& at(int r, int j) { return R[r][j]; }
#else
  __device__ __forceinline__ uint32_t& at(int r, int j) {
    // the register file starts at offset 0 of the dynamic shared memory; naming the extern array
    // here keeps the accesses in the shared address space (LDS/STS instead of generic LD/ST)
    extern __shared__ __align__(128) uint32_t plane_smem_base[];
    return plane_smem_base[(r * WPT + j) * BLOCK + threadIdx.x];
  }
#endif
  int env[WPT];       // local env index (relative to the launch's first env)
  int wi[WPT];        // word index inside the env
  uint32_t valid;     // bit j: word j is inside the batch
};

__device__ __forceinline__ int plane_reg(uint32_t idx, int i) { return (idx >> (6 * i)) & 63; }

// Bernoulli(table[idx]) for the thread's WPT words; M index planes
template <int M, bool HC, int WPT, int BLOCK>
__device__ PLANE_OUTLINE void plane_bernoulli(PlaneCtx<WPT, BLOCK>& C, const VmArgs& A, const PlaneProg& G,
                                             const PlaneOp& o, uint32_t* queue, int* qn, uint32_t* fix,
                                             int64_t env0, uint32_t step) {
  const int TT = 1 << M;
  const uint32_t* W = G.words + o.off;
  const uint32_t* bad = W;
  const uint32_t* always = W + TT;
  const uint32_t* lev = W + 2 * TT;                         // [levels][TT]
  const uint32_t* thr = W + (2 + PLANE_BERN_LEVELS) * TT;   // [TT][2] lo, hi
  const uint32_t* pvl = thr + 2 * TT;                       // [TT][2] float64 probability
  const double* inj = A.inj[o.a];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool has_bad = (o.b & 1) != 0, has_always = (o.b & 2) != 0;
  if (inj) {
    // parity mode: the reference's U <= p on the same float64 values (simulator.py:1053)
    PLANE_UNROLL_DYN
    for (int j = 0; j < WPT; ++j) {
      uint32_t res = 0;
      if ((C.valid >> j) & 1u) {
        uint32_t x[M > 0 ? M : 1];
#pragma unroll
        for (int i = 0; i < M; ++i) x[i] = C.at(plane_reg(o.idx, i), j);
        if (has_bad && plane_mux<M>(bad, x)) atomicOr(A.status + env0 + C.env[j], 1u);
        const double* u = inj + (env0 + C.env[j]) * PLANE_LAUNCH(A).E + ((int64_t)C.wi[j] << 5);
        for (int b = 0; b < 32; ++b) {
          int v = 0;
#pragma unroll
          for (int i = 0; i < M; ++i) v |= ((x[i] >> b) & 1u) << i;
          const double p = __hiloint2double((int)pvl[2 * v + 1], (int)pvl[2 * v]);
          res |= (uint32_t)(u[b] <= p) << b;
        }
      }
      C.at(o.dst, j) = res;
    }
    return;
  }
  // (the round keys precomputed on the host and read from the constant bank -- VmArgs::rk, what the scalar kernels
  // use -- measured 1.6 % slower here than the key schedule the compiler keeps in uniform registers)
  const uint2 key = make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32));
  const uint32_t node = ((uint32_t)o.b >> 8) & 0xffffu;   // Philox node id, placed here by the host decoder
  const uint32_t e0 = (uint32_t)(A.env_offset + env0);
  if (lane == 0) qn[warp] = 0;
  __syncwarp();
  // two words at a time: bounds the live registers (index planes, random words, res/und)
  PLANE_UNROLL_DYN
  for (int j0 = 0; j0 < WPT; j0 += 2) {
    constexpr int NW = WPT >= 2 ? 2 : 1;
    uint32_t x[NW][M > 0 ? M : 1], res[NW], und[NW];
#pragma unroll
    for (int jj = 0; jj < NW; ++jj) {
      const int j = j0 + jj;
#pragma unroll
      for (int i = 0; i < M; ++i) x[jj][i] = C.at(plane_reg(o.idx, i), j);
      const bool ok = (C.valid >> j) & 1u;
      if (has_bad && ok && plane_mux<M>(bad, x[jj])) atomicOr(A.status + env0 + C.env[j], 1u);
      const uint32_t al = has_always ? plane_mux<M>(always, x[jj]) : 0u;
      res[jj] = ok ? al : 0u;
      und[jj] = ok ? ~al : 0u;
    }
#pragma unroll
    for (int b = 0; b < PLANE_BERN_LEVELS / 4; ++b) {
      uint32_t rw[NW][4];
#pragma unroll
      for (int jj = 0; jj < NW; ++jj) {
        const uint4 r4 = rddl_philox4x32_10(
            make_uint4((uint32_t)C.wi[j0 + jj], e0 + (uint32_t)C.env[j0 + jj], step, node | ((0x8000u + b) << 16)), key);
        rw[jj][0] = r4.x; rw[jj][1] = r4.y; rw[jj][2] = r4.z; rw[jj][3] = r4.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t pj[NW];
        plane_mux_n<M, HC, NW>(lev + (4 * b + q) * TT, x, pj);
#pragma unroll
        for (int jj = 0; jj < NW; ++jj) {
          res[jj] |= und[jj] & ~rw[jj][q] & pj[jj];
          und[jj] &= ~(rw[jj][q] ^ pj[jj]);
        }
      }
    }
    // cells still tied after the first levels (2^-8 of them) are queued per warp and finished
    // cooperatively below, one 64-bit Philox draw per cell
#pragma unroll
    for (int jj = 0; jj < NW; ++jj) {
      const int j = j0 + jj;
      C.at(o.dst, j) = res[jj];
      fix[tid * WPT + j] = 0;
      uint32_t u = und[jj];
      while (u) {
        const int b = __ffs(u) - 1;
        u &= u - 1;
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < M; ++i) v |= ((x[jj][i] >> b) & 1u) << i;
        const int slot = atomicAdd(&qn[warp], 1);
        if (slot < PLANE_QUEUE) {
          queue[warp * PLANE_QUEUE + slot] = (uint32_t)lane | ((uint32_t)j << 5) | ((uint32_t)b << 8) | (v << 13);
        } else {   // queue full (practically never): finish inline
          const uint4 r4 = rddl_philox4x32_10(
              make_uint4(((uint32_t)C.wi[j] << 5) | b, e0 + (uint32_t)C.env[j], step, node | (0xC000u << 16)), key);
          const uint64_t t64 = ((uint64_t)thr[2 * v + 1] << 32) | thr[2 * v];
          if ((((uint64_t)r4.x << 32) | r4.y) < (t64 << PLANE_BERN_LEVELS)) C.at(o.dst, j) |= 1u << b;
        }
      }
    }
  }
  __syncwarp();
  const int n = min(qn[warp], PLANE_QUEUE);
  if (n > 0) {
    for (int i = lane; i < n; i += 32) {
      const uint32_t e = queue[warp * PLANE_QUEUE + i];
      const int ol = e & 31, oj = (e >> 5) & 7, b = (e >> 8) & 31;
      const uint32_t v = e >> 13;
      // the owner's word coordinates, recomputed from the tile geometry
      const int ow = oj * BLOCK + warp * 32 + ol;
      int env_l, wi_l;
      if (PLANE_LAUNCH(A).chunks == 1) {
        const int wpe = (int)(PLANE_LAUNCH(A).E >> 5);
        env_l = G.wpe_shift >= 0 ? (ow >> G.wpe_shift) : ow / wpe;
        wi_l = ow - env_l * wpe;
      } else {
        env_l = 0;
        wi_l = (blockIdx.x % PLANE_LAUNCH(A).chunks) * PLANE_LAUNCH(A).ipt + ow;
      }
      const uint4 r4 = rddl_philox4x32_10(
          make_uint4(((uint32_t)wi_l << 5) | b, e0 + (uint32_t)env_l, step, node | (0xC000u << 16)), key);
      const uint64_t t64 = ((uint64_t)thr[2 * v + 1] << 32) | thr[2 * v];
      if ((((uint64_t)r4.x << 32) | r4.y) < (t64 << PLANE_BERN_LEVELS))
        atomicOr(&fix[(warp * 32 + ol) * WPT + oj], 1u << b);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < WPT; ++j) C.at(o.dst, j) |= fix[tid * WPT + j];
  }
}

// Device-side policy draw of one or two bool action planes (OP_PPOL; rddl_device.cuh: rddl_policy_value is
// the per-cell form, oracle/philox_np.py: policy_actions the restatement): plane of entry e is
// [e.thr_lo <= U < e.thr] for the cell's bit-sliced uniform U under node e.node -- eight levels of
// Philox words compared MSB first against every threshold at once, ties finished with one 64-bit draw
// per cell by the warp together (same scheme as plane_bernoulli). Two entries of a pair share U.
//   o.dst / o.imm: register and policy entry of the first plane; o.a / o.b: of the second (o.a < 0: none)
//   cmp: [4][BLOCK * WPT] words of shared memory (results per comparator, patched by the tail)
template <int WPT, int BLOCK>
__device__ PLANE_OUTLINE void plane_policy(PlaneCtx<WPT, BLOCK>& C, const VmArgs& A, const PlaneProg& G, const PlaneOp& o,
                                          uint32_t* queue, int* qn, uint32_t* cmp, int64_t env0, uint32_t step) {
  const bool two = o.a >= 0;
  const int i0 = (int)o.imm, i1 = two ? o.b : (int)o.imm;
  const bool always0 = G.pol_always[i0] != 0, always1 = G.pol_always[i1] != 0;
  // comparators [U < Tk]: 0 = first.thr, 1 = first.thr_lo, 2 = second.thr, 3 = second.thr_lo (0 = unused: never
  // true; in the specialised build the thresholds are constants: unused comparators and all masks fold away)
  const uint64_t Tk[4] = {always0 ? 0ull : G.pol_thr[i0][0], G.pol_thr[i0][1], (two && !always1) ? G.pol_thr[i1][0] : 0ull,
                          two ? G.pol_thr[i1][1] : 0ull};
  const uint2 key = make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32));
  const uint32_t node = G.pol_node[i0] & 0xffffu;
  const uint32_t e0g = (uint32_t)(A.env_offset + env0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (lane == 0) qn[warp] = 0;
  __syncwarp();
  PLANE_UNROLL_DYN
  for (int j0 = 0; j0 < WPT; j0 += 2) {
    constexpr int NW = WPT >= 2 ? 2 : 1;
    uint32_t res[NW][4], und[NW][4];
#pragma unroll
    for (int jj = 0; jj < NW; ++jj) {
      const bool ok = (C.valid >> (j0 + jj)) & 1u;
#pragma unroll
      for (int k = 0; k < 4; ++k) { res[jj][k] = 0u; und[jj][k] = (ok && Tk[k] != 0ull) ? 0xffffffffu : 0u; }
    }
#pragma unroll
    for (int b = 0; b < PLANE_BERN_LEVELS / 4; ++b) {
      uint32_t rw[NW][4];
#pragma unroll
      for (int jj = 0; jj < NW; ++jj) {
        const uint4 r4 = rddl_philox4x32_10(
            make_uint4((uint32_t)C.wi[j0 + jj], e0g + (uint32_t)C.env[j0 + jj], step, node | ((0x8000u + b) << 16)), key);
        rw[jj][0] = r4.x; rw[jj][1] = r4.y; rw[jj][2] = r4.z; rw[jj][3] = r4.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t tb = ((Tk[k] >> (63 - (4 * b + q))) & 1ull) ? 0xffffffffu : 0u;     // (warp-uniform)
#pragma unroll
          for (int jj = 0; jj < NW; ++jj) {
            res[jj][k] |= und[jj][k] & ~rw[jj][q] & tb;
            und[jj][k] &= ~(rw[jj][q] ^ tb);
          }
        }
      }
    }
#pragma unroll
    for (int jj = 0; jj < NW; ++jj) {
      const int j = j0 + jj;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (Tk[k] != 0ull) cmp[k * (BLOCK * WPT) + tid * WPT + j] = res[jj][k];
      uint32_t u = und[jj][0] | und[jj][1] | und[jj][2] | und[jj][3];
      while (u) {      // cells still tied on some threshold (4 x 2^-8 of them): queued per warp
        const int b = __ffs(u) - 1;
        u &= u - 1;
        const uint32_t km = ((und[jj][0] >> b) & 1u) | (((und[jj][1] >> b) & 1u) << 1) | (((und[jj][2] >> b) & 1u) << 2) |
                            (((und[jj][3] >> b) & 1u) << 3);
        const int slot = atomicAdd(&qn[warp], 1);
        if (slot < PLANE_QUEUE) {
          queue[warp * PLANE_QUEUE + slot] = (uint32_t)lane | ((uint32_t)j << 5) | ((uint32_t)b << 8) | (km << 13);
        } else {     // queue full (practically never): finish inline
          const uint4 r4 = rddl_philox4x32_10(
              make_uint4(((uint32_t)C.wi[j] << 5) | b, e0g + (uint32_t)C.env[j], step, node | (0xC000u << 16)), key);
          const uint64_t u64 = ((uint64_t)r4.x << 32) | r4.y;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (((km >> k) & 1u) && u64 < (Tk[k] << PLANE_BERN_LEVELS)) cmp[k * (BLOCK * WPT) + tid * WPT + j] |= 1u << b;
        }
      }
    }
  }
  __syncwarp();
  const int n = min(qn[warp], PLANE_QUEUE);
  for (int i = lane; i < n; i += 32) {
    const uint32_t en = queue[warp * PLANE_QUEUE + i];
    const int ol = en & 31, oj = (en >> 5) & 7, b = (en >> 8) & 31;
    const uint32_t km = en >> 13;
    const int ow = oj * BLOCK + warp * 32 + ol;       // the owner's word coordinates, from the tile geometry
    int env_l, wi_l;
    if (PLANE_LAUNCH(A).chunks == 1) {
      const int wpe = (int)(PLANE_LAUNCH(A).E >> 5);
      env_l = ow / wpe;
      wi_l = ow - env_l * wpe;
    } else {
      env_l = 0;
      wi_l = (blockIdx.x % PLANE_LAUNCH(A).chunks) * PLANE_LAUNCH(A).ipt + ow;
    }
    const uint4 r4 = rddl_philox4x32_10(make_uint4(((uint32_t)wi_l << 5) | b, e0g + (uint32_t)env_l, step, node | (0xC000u << 16)), key);
    const uint64_t u64 = ((uint64_t)r4.x << 32) | r4.y;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (((km >> k) & 1u) && u64 < (Tk[k] << PLANE_BERN_LEVELS))
        atomicOr(&cmp[k * (BLOCK * WPT) + (warp * 32 + ol) * WPT + oj], 1u << b);
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < WPT; ++j) {
    const bool ok = (C.valid >> j) & 1u;
    const uint32_t c0 = Tk[0] != 0ull ? cmp[tid * WPT + j] : 0u, c1 = Tk[1] != 0ull ? cmp[BLOCK * WPT + tid * WPT + j] : 0u;
    C.at(o.dst, j) = always0 ? (ok ? 0xffffffffu : 0u) : (c0 & ~c1);
    if (two) {
      const uint32_t c2 = Tk[2] != 0ull ? cmp[2 * BLOCK * WPT + tid * WPT + j] : 0u;
      const uint32_t c3 = Tk[3] != 0ull ? cmp[3 * BLOCK * WPT + tid * WPT + j] : 0u;
      C.at(o.a, j) = always1 ? (ok ? 0xffffffffu : 0u) : (c2 & ~c3);
    }
  }
  __syncwarp();
}

template <int M, int WPT, int BLOCK>
__device__ __forceinline__ void plane_lut(PlaneCtx<WPT, BLOCK>& C, const PlaneProg& G, const PlaneOp& o) {
  constexpr int CH = WPT < 4 ? WPT : 4;
#pragma unroll
  for (int j0 = 0; j0 < WPT; j0 += CH) {
    uint32_t x[CH][M > 0 ? M : 1], out[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
#pragma unroll
      for (int i = 0; i < M; ++i) x[j][i] = C.at(plane_reg(o.idx, i), j0 + j);
    }
    plane_mux_n<M, false, CH>(G.words + o.off, x, out);
#pragma unroll
    for (int j = 0; j < CH; ++j) C.at(o.dst, j0 + j) = out[j];
  }
}

// Partial tiles (last CTA of the batch / of an env): per-word validity and 64-bit addressing.
// Kept out of line so that their address arithmetic is not hoisted into the dispatch loop.
template <int WPT, int BLOCK>
__device__ PLANE_OUTLINE void plane_ld_partial(PlaneCtx<WPT, BLOCK>& C, const uint8_t* base, int64_t env0, int wpe,
                                              int dst) {
  PLANE_UNROLL_DYN
  for (int j = 0; j < WPT; ++j) {
    uint32_t word = 0;
    if ((C.valid >> j) & 1u) {
      const uint4* p = reinterpret_cast<const uint4*>(base + (((env0 + C.env[j]) * wpe + C.wi[j]) << 5));
      word = plane_pack(__ldcs(p), __ldcs(p + 1));
    }
    C.at(dst, j) = word;
  }
}

template <int WPT, int BLOCK>
__device__ PLANE_OUTLINE void plane_st_partial(PlaneCtx<WPT, BLOCK>& C, uint8_t* base, int64_t env0, int wpe, int src) {
  PLANE_UNROLL_DYN
  for (int j = 0; j < WPT; ++j) {
    if ((C.valid >> j) & 1u) {
      uint4 lo, hi;
      plane_unpack(C.at(src, j), lo, hi);
      uint4* p = reinterpret_cast<uint4*>(base + (((env0 + C.env[j]) * wpe + C.wi[j]) << 5));
      __stcs(p, lo);
      __stcs(p + 1, hi);
    }
  }
}

#ifndef PLANE_MIN_BLOCKS
#define PLANE_MIN_BLOCKS 2
#endif

// Shared-memory layout of a staged stencil plane: rows are padded with one zero word on each
// side and separated (per env) by a zero row, so that PCOUNT needs no boundary conditionals:
//   index(row r of the tile, word c) = (r + 1 + r / H_rows_per_env) * (W32 + 2) + c + 1   (chunks == 1)
//   index(r, c) = (r + 1) * (W32 + 2) + c + 1, halo rows at r = -1 and r = rows             (chunks > 1)
__device__ __forceinline__ int plane_spos(int row, int c, int env_l, int pitch) {
  return (row + 1 + env_l) * pitch + c + 1;
}

// Per-tile state shared by the op handlers (all of it lives in registers after inlining).
template <bool B>
struct PlaneBool { static constexpr bool value = B; };

template <int WPT, int BLOCK>
struct PlaneTile {
  PlaneCtx<WPT, BLOCK> C;
  int spos[WPT];        // shared-memory index of the word inside a staged plane
  uint32_t* S;          // staged stencil planes
  int32_t* cnt;         // [32][MAX_RED] per-env popcounts
  uint32_t* queue;      // [8][PLANE_QUEUE]
  int* qn;              // [8]
  uint32_t* fix;        // [256][WPT]
  int32_t* sel;         // [32][pol_k] groundings that keep their policy draw this step (PPOLSEL)
  uint32_t* cmp;        // [4][BLOCK * WPT] comparator results of the policy draws (PPOL), when the program has any
  int tma_off;          // word offset (from S) of the TMA region: mbarrier (128 B) | [ntma][nf_words] staged non-fluent planes
  int64_t env0, tile_byte0;
  int64_t gw_limit;     // last valid global word index this tile may touch (ragged tiles clamp loads to it)
  int W32, wpe, envs_per_tile, tile_words, chunks, chunk, pitch, tile_rows, plane_stride;
  bool full, single_env;
};

#ifdef PLANE_STATIC
// Consecutive byte-plane loads (the action planes) are issued as one batch: all their 128-bit loads
// are in flight before the first is consumed, so the tile pays one memory latency instead of one
// per plane. plane_group_len(pc) = planes in the batch starting at op pc (0: pc is inside a batch).
__host__ __device__ constexpr bool plane_is_byte_load(int pc) {
  return pc >= 0 && pc < PS_NOPS && kG.ops[pc].op == OP_PLD && kG.ops[pc].a == 0 && kDT[kG.ops[pc].imm] != DT_PACKED;
}
__host__ __device__ constexpr int plane_group_len(int pc) {
  if (!plane_is_byte_load(pc) || plane_is_byte_load(pc - 1)) return 0;
  int n = 1;
  while (n < 4 && plane_is_byte_load(pc + n)) ++n;
  return n;
}
#endif

// What plane_exec does with an op: run it, or only start the DRAM reads of the planes it loads
// (all PLD planes of this step / the action planes of the next step of a fused rollout).
enum { PLANE_RUN = 0, PLANE_PREFETCH = 1, PLANE_PREFETCH_NEXT = 2 };

// One op of the program for this thread's WPT words. Interpreter: `pc` is a run-time index into the
// parameter-bank program. Specialised build: PC is a template constant and the op a constant
// expression, so only the selected case survives.
template <int WPT, int BLOCK, int MODE, int PC>
__device__ __forceinline__ void plane_exec(PlaneTile<WPT, BLOCK>& T, const VmArgs& A, const PlaneProg& G,
                                           const RddlLaunch& L, const uint8_t* DT, int pc, int t, bool last_step) {
#ifdef PLANE_STATIC
  constexpr PlaneOp o = kG.ops[PC];
#else
  const PlaneOp& o = G.ops[pc];
#endif
  PlaneCtx<WPT, BLOCK>& C = T.C;
  const int tid = threadIdx.x;
  const int64_t env0 = T.env0, tile_byte0 = T.tile_byte0;
  const bool full = T.full;
  if (MODE != PLANE_RUN) {
    if (o.op != OP_PLD || DT[o.imm] == DT_PACKED) return;
    if (MODE == PLANE_PREFETCH_NEXT && (o.a > 0 || A.tstride[o.imm] == 0)) return;
    const uint8_t* p = reinterpret_cast<const uint8_t*>(A.tensors[o.imm]) +
                       (MODE == PLANE_PREFETCH_NEXT ? (t + 1) * A.tstride[o.imm] : 0);
    // a warp's words of one plane are WPT runs of 1 KiB = WPT * 8 lines of 128 B: one line per lane,
    // one prefetch instruction per plane
    const int lane = tid & 31;
    if (lane < WPT * 8) {
      int64_t gw = (tile_byte0 >> 5) - lane + (lane >> 3) * BLOCK + (lane & 7) * 4;
      if (!full) gw = gw < T.gw_limit ? gw : T.gw_limit;      // ragged tile: stay inside the tensor
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (gw << 5)));
    }
    return;
  }
  const int W32 = T.W32, wpe = T.wpe, tile_words = T.tile_words, chunks = T.chunks, chunk = T.chunk;
  const int pitch = T.pitch, tile_rows = T.tile_rows, plane_stride = T.plane_stride;
  const bool single_env = T.single_env;
  uint32_t* S = T.S;
  int32_t* cnt = T.cnt;
  uint32_t* queue = T.queue;
  int* qn = T.qn;
  uint32_t* fix = T.fix;
  const int* spos = T.spos;
    switch (o.op) {
      case OP_PLD: {
        if (t > 0 && o.a > 0) {   // state fluent: carried over from the value stored last step
#pragma unroll
          for (int j = 0; j < WPT; ++j) C.at(o.dst, j) = C.at(o.a - 1, j);
          break;
        }
        const uint8_t* base = reinterpret_cast<const uint8_t*>(A.tensors[o.imm]) + t * A.tstride[o.imm];
        if (DT[o.imm] == DT_PACKED) {   // state kept bit-packed in HBM: one word per 32 cells
          const uint32_t* wp = reinterpret_cast<const uint32_t*>(base) + (tile_byte0 >> 5);
#pragma unroll
          for (int j = 0; j < WPT; ++j)
            C.at(o.dst, j) = (full || ((C.valid >> j) & 1u)) ? __ldcs(wp + j * BLOCK) : 0u;
          break;
        }
#ifdef PLANE_STATIC
        if constexpr (plane_is_byte_load(PC)) {
          constexpr int GL = plane_group_len(PC);
          if constexpr (GL >= 1) {
            // the whole batch of consecutive byte planes: loads first, then the packing. Ragged tiles
            // (RG: the last tile of the batch / of an env) clamp the word address to the last valid
            // word and zero the words outside the batch instead of taking a separate slow path.
            auto load_batch = [&](auto rg) {
              constexpr bool RG = decltype(rg)::value;
              constexpr int CHG = (4 / GL < 1 ? 1 : 4 / GL) < WPT ? (4 / GL < 1 ? 1 : 4 / GL) : WPT;   // words per batch
#pragma unroll
              for (int j0 = 0; j0 < WPT; j0 += CHG) {
                uint4 lo[GL][CHG], hi[GL][CHG];
#pragma unroll
                for (int q = 0; q < GL; ++q) {
                  const uint32_t ten = kG.ops[PC + q].imm;
                  const uint8_t* bq = reinterpret_cast<const uint8_t*>(A.tensors[ten]) + t * A.tstride[ten];
#pragma unroll
                  for (int j = 0; j < CHG; ++j) {
                    if (j0 + j < WPT) {
                      const uint4* pq;
                      if constexpr (RG) {
                        int64_t gw = (tile_byte0 >> 5) + (j0 + j) * BLOCK;
                        gw = gw < T.gw_limit ? gw : T.gw_limit;
                        pq = reinterpret_cast<const uint4*>(bq + (gw << 5));
                      } else {
                        pq = reinterpret_cast<const uint4*>(bq + tile_byte0) + (j0 + j) * (BLOCK * 2);
                      }
                      lo[q][j] = __ldcs(pq);
                      hi[q][j] = __ldcs(pq + 1);
                    }
                  }
                }
#pragma unroll
                for (int q = 0; q < GL; ++q) {
#pragma unroll
                  for (int j = 0; j < CHG; ++j) {
                    if (j0 + j < WPT) {
                      const uint32_t w = plane_pack(lo[q][j], hi[q][j]);
                      C.at(kG.ops[PC + q].dst, j0 + j) = (!RG || ((C.valid >> (j0 + j)) & 1u)) ? w : 0u;
                    }
                  }
                }
              }
            };
            if (full) load_batch(PlaneBool<false>{});
            else load_batch(PlaneBool<true>{});
          }
          break;   // GL == 0: this plane was loaded with the batch of an earlier op
        }
#endif
        if (!full) {
          plane_ld_partial<WPT, BLOCK>(C, base, env0, wpe, o.dst);
          break;
        }
        const uint4* p = reinterpret_cast<const uint4*>(base + tile_byte0);
        constexpr int CH = WPT < 4 ? WPT : 4;     // loads in flight per thread: 2 * CH x 16 B
#pragma unroll
        for (int j0 = 0; j0 < WPT; j0 += CH) {
          uint4 lo[CH], hi[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            lo[j] = __ldcs(p + (j0 + j) * (BLOCK * 2));
            hi[j] = __ldcs(p + (j0 + j) * (BLOCK * 2) + 1);
          }
#pragma unroll
          for (int j = 0; j < CH; ++j) C.at(o.dst, j0 + j) = plane_pack(lo[j], hi[j]);
        }
        break;
      }
      case OP_PLDC: {
        if (G.ntma > 0 && A.pool_map_ok) {
          // the tile's slice of the pre-packed non-fluent plane was copied to shared memory by TMA
          // (plane_tile prologue): make sure the copies have landed (a completed mbarrier phase answers
          // at once), then read it from there
          uint32_t* tma = T.S + T.tma_off;
          uint32_t landed = 0;
          if (t == 0) landed = rddl_mbar_wait(reinterpret_cast<uint64_t*>(tma), 0);       // 0 (later steps: already there)
          const int nf_words = ((chunks == 1 ? wpe : tile_words) + RDDL_TMA_BOX - 1) / RDDL_TMA_BOX * RDDL_TMA_BOX;
          const uint32_t* sp = tma + 32 + o.b * nf_words - (chunks == 1 ? 0 : chunk * tile_words) + landed;
#pragma unroll
          for (int j = 0; j < WPT; ++j) C.at(o.dst, j) = (full || ((C.valid >> j) & 1u)) ? sp[C.wi[j]] : 0u;
          break;
        }
#pragma unroll
        for (int j = 0; j < WPT; ++j)
          C.at(o.dst, j) = (full || ((C.valid >> j) & 1u)) ? (uint32_t)__ldg(A.pool + o.imm + C.wi[j]) : 0u;
        break;
      }
      case OP_PST: {
        if (!last_step && o.b) break;   // carried state: only the final step reaches HBM
        uint8_t* base = reinterpret_cast<uint8_t*>(A.tensors[o.imm]);
        if (DT[o.imm] == DT_PACKED) {
          uint32_t* wp = reinterpret_cast<uint32_t*>(base) + (tile_byte0 >> 5);
#pragma unroll
          for (int j = 0; j < WPT; ++j)
            if (full || ((C.valid >> j) & 1u)) __stcs(wp + j * BLOCK, C.at(o.a, j));
          break;
        }
        if (!full) {
          plane_st_partial<WPT, BLOCK>(C, base, env0, wpe, o.a);
          break;
        }
        uint4* p = reinterpret_cast<uint4*>(base + tile_byte0);
#pragma unroll
        for (int j = 0; j < WPT; ++j) {
          uint4 lo, hi;
          plane_unpack(C.at(o.a, j), lo, hi);
          __stcs(p + j * (BLOCK * 2), lo);
          __stcs(p + j * (BLOCK * 2) + 1, hi);
        }
        break;
      }
      case OP_PCONST: {
#pragma unroll
        for (int j = 0; j < WPT; ++j) C.at(o.dst, j) = o.imm ? 0xffffffffu : 0u;
        break;
      }
      case OP_PLUT: {
#if defined(PLANE_ABLATE) && (PLANE_ABLATE & 2)
        for (int j = 0; j < WPT; ++j) C.at(o.dst, j) = C.at(plane_reg(o.idx, 0), j);
        break;
#endif
        switch (o.m & 0xff) {
          case 1: plane_lut<1>(C, G, o); break;
          case 2: plane_lut<2>(C, G, o); break;
          case 3: plane_lut<3>(C, G, o); break;
          case 4: plane_lut<4>(C, G, o); break;
          default: plane_lut<5>(C, G, o); break;
        }
        break;
      }
      case OP_PSHARE: {
        uint32_t* s = S + o.b * plane_stride;
#pragma unroll
        for (int j = 0; j < WPT; ++j) {
          const int w = j * BLOCK + tid;
          if (w < tile_words) s[spos[j]] = (full || ((C.valid >> j) & 1u)) ? C.at(o.a, j) : 0u;
        }
        if (chunks > 1) {
          // halo rows above / below the tile, packed straight from the source tensor
          const uint8_t* base = reinterpret_cast<const uint8_t*>(A.tensors[o.imm]) + t * A.tstride[o.imm];
          for (int i = tid; i < 2 * W32; i += BLOCK) {
            const bool top = i < W32;
            const int c = top ? i : i - W32;
            const int wsrc = top ? chunk * tile_words - W32 + c : (chunk + 1) * tile_words + c;
            uint32_t word = 0;
            if (wsrc >= 0 && wsrc < wpe && env0 < A.n_envs) {
              if (DT[o.imm] == DT_PACKED) {
                word = __ldg(reinterpret_cast<const uint32_t*>(base) + env0 * wpe + wsrc);
              } else {
                const uint4* p = reinterpret_cast<const uint4*>(base + ((env0 * wpe + wsrc) << 5));
                word = plane_pack(__ldg(p), __ldg(p + 1));
              }
            }
            s[plane_spos(top ? -1 : tile_rows, c, 0, pitch)] = word;
          }
        }
        __syncthreads();
        break;
      }
      case OP_PCOUNT: {
#if defined(PLANE_ABLATE) && (PLANE_ABLATE & 8)
        for (int j = 0; j < WPT; ++j) { C.at(o.dst, j) = 0; C.at(o.dst + 1, j) = 0; C.at(o.dst + 2, j) = 0; C.at(o.dst + 3, j) = 0; }
        break;
#endif
        const uint32_t* s = S + o.b * plane_stride;
        const uint32_t m9 = o.imm;
        // uniform masks of the stencil offsets (all-ones for the common full neighbourhoods)
        uint32_t mk[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) mk[i] = ((m9 >> i) & 1u) ? 0xffffffffu : 0u;
#pragma unroll
        for (int j = 0; j < WPT; ++j) {
          const int q0 = spos[j];
          uint32_t sr[3], cr[3];   // per row: sum and carry of the (west, centre, east) planes
#pragma unroll
          for (int d = 0; d < 3; ++d) {
            const int q = q0 + (d - 1) * pitch;
            const uint32_t Cw = s[q], Lw = s[q - 1], Rw = s[q + 1];
            const uint32_t west = __funnelshift_l(Lw, Cw, 1) & mk[3 * d];
            const uint32_t mid = Cw & mk[3 * d + 1];
            const uint32_t east = __funnelshift_r(Cw, Rw, 1) & mk[3 * d + 2];
            sr[d] = west ^ mid ^ east;
            cr[d] = (west & mid) | (east & (west ^ mid));
          }
          const uint32_t k0 = sr[0] ^ sr[1] ^ sr[2];
          const uint32_t t1 = (sr[0] & sr[1]) | (sr[2] & (sr[0] ^ sr[1]));   // weight 2
          const uint32_t u1 = cr[0] ^ cr[1] ^ cr[2];                         // weight 2
          const uint32_t u2 = (cr[0] & cr[1]) | (cr[2] & (cr[0] ^ cr[1]));   // weight 4
          const uint32_t v = t1 & u1;                                        // weight 4
          C.at(o.dst, j) = k0;
          C.at(o.dst + 1, j) = t1 ^ u1;
          C.at(o.dst + 2, j) = u2 ^ v;
          C.at(o.dst + 3, j) = u2 & v;
        }
        break;
      }
      case OP_PBERN: {
#if defined(PLANE_ABLATE) && (PLANE_ABLATE & 1)
        for (int j = 0; j < WPT; ++j) C.at(o.dst, j) = C.at(plane_reg(o.idx, 0), j);
        break;
#endif
        switch (o.m) {
          case 0: plane_bernoulli<0, false>(C, A, G, o, queue, qn, fix, env0, (uint32_t)(A.step + t)); break;
          case 1: plane_bernoulli<1, false>(C, A, G, o, queue, qn, fix, env0, (uint32_t)(A.step + t)); break;
          case 2: plane_bernoulli<2, false>(C, A, G, o, queue, qn, fix, env0, (uint32_t)(A.step + t)); break;
          case 3: plane_bernoulli<3, false>(C, A, G, o, queue, qn, fix, env0, (uint32_t)(A.step + t)); break;
          case 4: plane_bernoulli<4, false>(C, A, G, o, queue, qn, fix, env0, (uint32_t)(A.step + t)); break;
          case 0x102: plane_bernoulli<2, true>(C, A, G, o, queue, qn, fix, env0, (uint32_t)(A.step + t)); break;
          case 0x103: plane_bernoulli<3, true>(C, A, G, o, queue, qn, fix, env0, (uint32_t)(A.step + t)); break;
          default: plane_bernoulli<4, true>(C, A, G, o, queue, qn, fix, env0, (uint32_t)(A.step + t)); break;
        }
        break;
      }
      case OP_PPOL: {
        plane_policy<WPT, BLOCK>(C, A, G, o, queue, qn, T.cmp, env0, (uint32_t)(A.step + t));
        break;
      }
      case OP_PPOLSEL: {
        // max-nondef rule of a device-side policy: only the groundings selected for this (env, step)
        // keep the draw already in the register, the others fall back to the fluent's default
        const VmPolicyEntry& e = A.pol.e[o.imm];
        const uint32_t dm = e.dflt ? 0xffffffffu : 0u;
        const int k = G.pol_k;
#pragma unroll
        for (int j = 0; j < WPT; ++j) {
          const int32_t* s = T.sel + C.env[j] * k;
          uint32_t mask = 0;
          for (int i = 0; i < k; ++i) {
            const int64_t gi = (int64_t)s[i] - e.base;
            if (gi >= 0 && gi < e.nelem && (gi >> 5) == C.wi[j]) mask |= 1u << (gi & 31);
          }
          C.at(o.dst, j) = (C.at(o.dst, j) & mask) | (dm & ~mask);
        }
        break;
      }
      case OP_PRED: {
#if defined(PLANE_ABLATE) && (PLANE_ABLATE & 4)
        break;
#endif
        // per-env popcount of a plane (m == 1) or of the valid cells (m == 0)
        const int r0 = plane_reg(o.idx, 0);
        if (single_env) {
          int c = 0;
#pragma unroll
          for (int j = 0; j < WPT; ++j)
            if (full || ((C.valid >> j) & 1u)) c += o.m ? __popc(C.at(r0, j)) : 32;
          c = __reduce_add_sync(0xffffffffu, c);
          if ((tid & 31) == 0 && c) atomicAdd(&cnt[o.b], c);
        } else {
          const bool warp_uniform_env = (wpe & 31) == 0;
#pragma unroll
          for (int j = 0; j < WPT; ++j) {
            int c = (full || ((C.valid >> j) & 1u)) ? (o.m ? __popc(C.at(r0, j)) : 32) : 0;
            if (warp_uniform_env) {
              c = __reduce_add_sync(0xffffffffu, c);
              if ((tid & 31) == 0 && c) atomicAdd(&cnt[C.env[j] * RDDL_MAX_RED + o.b], c);
            } else if (c) {
              atomicAdd(&cnt[C.env[j] * RDDL_MAX_RED + o.b], c);
            }
          }
        }
        break;
      }
      default: break;
    }
}

// All ops in program order: a loop in the interpreter, a compile-time sequence when specialised.
template <int WPT, int BLOCK, int MODE, int PC>
__device__ __forceinline__ void plane_run_ops(PlaneTile<WPT, BLOCK>& T, const VmArgs& A, const PlaneProg& G,
                                              const RddlLaunch& L, const uint8_t* DT, int t, bool last_step) {
#ifdef PLANE_STATIC
  if constexpr (PC < PS_NOPS) {
    plane_exec<WPT, BLOCK, MODE, PC>(T, A, G, L, DT, PC, t, last_step);
    plane_run_ops<WPT, BLOCK, MODE, PC + 1>(T, A, G, L, DT, t, last_step);
  }
#else
#pragma unroll 1
  for (int pc = 0; pc < G.nops; ++pc) plane_exec<WPT, BLOCK, MODE, 0>(T, A, G, L, DT, pc, t, last_step);
#endif
}

// One tile of one launch: all ops of the program (for all steps of a fused rollout) by one CTA.
// G / L / DT are kernel parameters (interpreter) or the constants of the specialised build.
template <int WPT, int BLOCK>
__device__ __forceinline__ void plane_tile(const VmArgs& A, const PlaneProg& G, const RddlLaunch& L,
                                           const uint8_t* DT) {
  const int tid = threadIdx.x;
  const int W32 = (int)((L.rank == 2 ? L.extent[1] : L.extent[0]) >> 5);
  const int wpe = (int)(L.E >> 5);              // words per env
  const int envs_per_tile = L.G, tile_words = L.ipt, chunks = L.chunks;
  const int chunk = blockIdx.x % chunks;
  const int64_t env0 = (int64_t)(blockIdx.x / chunks) * envs_per_tile;   // first env of this tile
  const int pitch = W32 + 2;
  const int tile_rows = tile_words / W32;

  // dynamic shared memory: (interpreter only: plane registers [nregs][WPT][256]) |
  // [nshared] padded stencil planes | cnt | queue | qn | fix
  extern __shared__ __align__(128) uint32_t smem[];
  const int plane_stride = (tile_rows + envs_per_tile + 2) * pitch;
  PlaneTile<WPT, BLOCK> T;
#ifdef PLANE_STATIC
  T.S = smem;                               // plane registers are machine registers
#else
  T.S = smem + L.nregs * WPT * BLOCK;
#endif
  uint32_t* S = T.S;
  int32_t* cnt = T.cnt = reinterpret_cast<int32_t*>(S + G.nshared * plane_stride);
  T.queue = reinterpret_cast<uint32_t*>(cnt + 32 * RDDL_MAX_RED);
  T.qn = reinterpret_cast<int*>(T.queue + 8 * PLANE_QUEUE);
  T.fix = reinterpret_cast<uint32_t*>(T.qn + 8);
  T.sel = reinterpret_cast<int32_t*>(T.fix + BLOCK * WPT);
  // TMA staging of the shared non-fluent planes: the words [nf_first, nf_first + nf_words) of every
  // PLDC plane -- the whole plane when a tile holds whole envs, the tile's rows otherwise -- arrive in
  // shared memory as boxes of RDDL_TMA_BOX words (cp.async.bulk.tensor.1d on the pool's tensor map)
  // while the tile sets up and issues its fluent loads; the first PLDC op waits on the mbarrier.
  // (region offset: everything before it, rounded up to 128 bytes; the dynamic shared memory is 128-byte aligned)
  T.cmp = reinterpret_cast<uint32_t*>(T.sel + 32 * G.pol_k);
  T.tma_off = (int)(((T.cmp + (G.npol > 0 ? 4 * BLOCK * WPT : 0) - S) + 31) & ~31);
  if (G.ntma > 0 && A.pool_map_ok) {
    uint32_t* tma = S + T.tma_off;
    const int nf_words = ((chunks == 1 ? wpe : tile_words) + RDDL_TMA_BOX - 1) / RDDL_TMA_BOX * RDDL_TMA_BOX;
    const int nf_first = chunks == 1 ? 0 : chunk * tile_words;
    if (tid == 0) {
      uint64_t* bar = reinterpret_cast<uint64_t*>(tma);
      rddl_mbar_init(bar, 1);
      rddl_mbar_expect_tx(bar, (uint32_t)(G.ntma * nf_words * 4));
      for (int a = 0; a < G.nops; ++a) {
        if (G.ops[a].op != OP_PLDC) continue;
        for (int b = 0; b < nf_words; b += RDDL_TMA_BOX)
          rddl_tma_load_1d(tma + 32 + G.ops[a].b * nf_words + b, &A.pool_map, bar, (int32_t)G.ops[a].imm + nf_first + b);
      }
    }
    __syncthreads();      // the barrier is initialised before anybody polls it
  }
  if (L.nred > 0)
    for (int i = tid; i < 32 * RDDL_MAX_RED; i += BLOCK) cnt[i] = 0;
  for (int i = tid; i < G.nshared * plane_stride; i += BLOCK) S[i] = 0u;           // zero padding, once

  PlaneCtx<WPT, BLOCK>& C = T.C;
  C.valid = 0;
#pragma unroll
  for (int j = 0; j < WPT; ++j) {
    const int w = j * BLOCK + tid;
    bool ok = w < tile_words;
    if (chunks == 1) {
      C.env[j] = G.wpe_shift >= 0 ? (w >> G.wpe_shift) : w / wpe;
      C.wi[j] = w - C.env[j] * wpe;
    } else {
      C.env[j] = 0;
      C.wi[j] = chunk * tile_words + w;
      ok = ok && C.wi[j] < wpe;
    }
    if (ok && env0 + C.env[j] < A.n_envs) C.valid |= 1u << j;
    const int row = G.w32_shift >= 0 ? (w >> G.w32_shift) : w / W32;        // row inside the tile
    T.spos[j] = plane_spos(row, w - row * W32, C.env[j], pitch);
  }
  // whole tile inside the batch (the common case): no per-word validity handling below
  const bool full = __syncthreads_and(C.valid == ((1u << WPT) - 1u)) != 0;
  T.full = full;
  T.env0 = env0;
  // byte offset of this thread's first word (word j is 256 words further), valid when `full`
  T.tile_byte0 = ((env0 * wpe + (int64_t)chunk * tile_words + tid) << 5);
  T.gw_limit = (chunks == 1 ? A.n_envs * (int64_t)wpe : (env0 + 1) * (int64_t)wpe) - 1;
  T.W32 = W32; T.wpe = wpe; T.envs_per_tile = envs_per_tile; T.tile_words = tile_words; T.chunks = chunks;
  T.chunk = chunk; T.pitch = pitch; T.tile_rows = tile_rows; T.plane_stride = plane_stride;
  T.single_env = envs_per_tile == 1;
  // issue the DRAM requests of every plane this tile reads up front (L2 prefetch): the loads of
  // the later PLD ops then hit L2 and the compute of this CTA overlaps its own memory traffic
#ifdef PLANE_STATIC
  plane_run_ops<WPT, BLOCK, PLANE_PREFETCH, 0>(T, A, G, L, DT, 0, false);
#else
  if (full) plane_run_ops<WPT, BLOCK, PLANE_PREFETCH, 0>(T, A, G, L, DT, 0, false);
#endif

  // per-env rollout state of the thread that runs env (env0 + tid)'s scalar epilogue
  double ret_acc = 0.0, disc = 1.0;
  bool alive = true;
  const int NT = A.n_steps > 1 ? A.n_steps : 1;
  const bool rollout = A.returns != nullptr;     // rddl_rollout (any n_steps >= 1): returns / per-step rewards are accumulated
#pragma unroll 1
  for (int t = 0; t < NT; ++t) {
  const bool last_step = t == NT - 1;
  // next step's action planes: start their DRAM reads now (L2 prefetch)
#ifdef PLANE_STATIC
  if (!last_step) plane_run_ops<WPT, BLOCK, PLANE_PREFETCH_NEXT, 0>(T, A, G, L, DT, t, last_step);
#else
  if (full && !last_step) plane_run_ops<WPT, BLOCK, PLANE_PREFETCH_NEXT, 0>(T, A, G, L, DT, t, last_step);
#endif
  if (G.pol_k > 0) {
    // device-side policy with a max-nondef rule: the pol_k groundings of each env of the tile that keep
    // their draw this step -- candidates in parallel, then Floyd's resolution by one thread per env
    // (the same functions as the scalar kernels' rddl_policy_select_kernel)
    const int k = G.pol_k;
    for (int i = tid; i < envs_per_tile * k; i += BLOCK) {
      const int e = i / k, q = i - e * k;
      T.sel[i] = (int32_t)rddl_policy_pick(A.seed, (uint32_t)(A.env_offset + env0 + e), (uint32_t)(A.step + t), (uint32_t)q,
                                           A.pol.total - k + q);
    }
    __syncthreads();
    if (tid < envs_per_tile) rddl_policy_floyd(T.sel + tid * k, T.sel + tid * k, k, A.pol.total);
    __syncthreads();
  }
  plane_run_ops<WPT, BLOCK, PLANE_RUN, 0>(T, A, G, L, DT, t, last_step);

  // Specialised single-tile launches with a fused scalar program: the thread that runs env e's
  // epilogue finishes e's reductions itself and hands them over in registers (one barrier, no
  // round trip through the reduction scratch in L2).
#if defined(PLANE_STATIC) && PS_FUSED_STATIC
  constexpr bool kRedLocal = kL.chunks == 1 && kL.nred > 0 && kG.nfused > 0;
#else
  constexpr bool kRedLocal = false;
#endif
  const int64_t cells = chunks == 1 ? L.E : (int64_t)min(tile_words, wpe - chunk * tile_words) * 32;
  // value of reduction jr of local env e from its popcount (the PRED op carries the per-cell values
  // at 0 and at 1 as raw 64-bit words); clears the counter for the next step of a fused rollout
  auto finish_red = [&](int e, int jr) -> uint64_t {
    const int off = G.red_off[jr], m = G.red_m[jr];
    const uint64_t v0 = ((uint64_t)G.words[off + 1] << 32) | G.words[off];
    const uint64_t v1 = ((uint64_t)G.words[off + 3] << 32) | G.words[off + 2];
    const int64_t ones = cnt[e * RDDL_MAX_RED + jr];
    cnt[e * RDDL_MAX_RED + jr] = 0;
    const int64_t n1 = m ? ones : 0, n0 = m ? cells - ones : cells;
    if (L.red_op[jr] == RED_SUM_F) {
      double acc = __dmul_rn(__longlong_as_double((long long)v0), (double)n0);
      acc = __dadd_rn(acc, __dmul_rn(__longlong_as_double((long long)v1), (double)n1));
      return (uint64_t)__double_as_longlong(acc);
    }
    return (uint64_t)((int64_t)v0 * n0 + (int64_t)v1 * n1);
  };
  if (L.nred > 0 && !kRedLocal) {
    __syncthreads();
    for (int i = tid; i < envs_per_tile * L.nred; i += BLOCK) {
      const int e = i / L.nred, jr = i - e * L.nred;
      const int64_t ge = env0 + e;
      if (ge >= A.n_envs) continue;
      A.red[A.red_off[L.red_slot[jr]] * A.n_envs + ge * chunks + chunk] = finish_red(e, jr);
    }
  }
  if (G.nfused > 0) {
    // scope-free programs fused into this launch (reward / termination from the reductions):
    // one thread per env of the tile runs them through the scalar interpreter
    if (!kRedLocal) __threadfence_block();
    __syncthreads();
    if (tid < envs_per_tile && env0 + tid < A.n_envs) {
      const int64_t ge = env0 + tid;
      int32_t idx0[RDDL_MAX_IDX];
      uint64_t rv[RDDL_MAX_RED];
#if defined(PLANE_STATIC) && PS_FUSED_STATIC
      // specialised epilogue: straight-line code, reward / done taken from the stores themselves
      uint64_t cap[2] = {0, 0};
      const VmPools P = {kConsts, kAddrs, kRedSlots, kRedOff, nullptr, kDT};
      uint64_t redv[RDDL_MAX_RED];
      const uint64_t* red_local = nullptr;
      if constexpr (kRedLocal) {
#pragma unroll
        for (int jr = 0; jr < RDDL_MAX_RED; ++jr) {
          if (jr < kL.nred) {
            redv[jr] = finish_red(tid, jr);
            A.red[A.red_off[L.red_slot[jr]] * A.n_envs + ge] = redv[jr];     // (kept for launches outside this kernel)
          }
        }
        red_local = redv;
      }
      {
        uint64_t r0[40];
        vm_exec_static<kG.fused[0].pc_begin, kG.fused[0].pc_end>(A, P, kG.fused[0], ge, A.step + t, idx0, rv, r0, cap, G.pad, red_local);
      }
      if constexpr (kG.nfused > 1) {
        uint64_t r1[40];
        vm_exec_static<kG.fused[1].pc_begin, kG.fused[1].pc_end>(A, P, kG.fused[1], ge, A.step + t, idx0, rv, r1, cap, G.pad, red_local);
      }
      if (rollout) {
        const double rwd = as_f(cap[0]);
        const bool done_now = cap[1] != 0;
#else
      for (int f = 0; f < G.nfused; ++f) vm_exec_element<40>(A, G.fused[f], ge, A.step + t, idx0, rv);
      if (rollout) {
        // BaseAgent.evaluate (pyRDDLGym/core/policy.py:100-117): total += reward * gamma^t until done
        const double rwd = __ldcg(reinterpret_cast<const double*>(A.tensors[G.pad]) + ge);
        const bool done_now = __ldcg(reinterpret_cast<const uint8_t*>(A.tensors[G.pad + 1]) + ge) != 0;
#endif
        if (A.rewards) A.rewards[(int64_t)t * A.n_envs + ge] = rwd;
        if (alive) ret_acc += rwd * disc;
        disc *= A.gamma;
        if (done_now) alive = false;
      }
    }
  }
  }   // step loop
  if (rollout && tid < envs_per_tile && env0 + tid < A.n_envs) A.returns[env0 + tid] += ret_acc;
}

#ifdef PLANE_STATIC
extern "C" __global__ void __launch_bounds__(PS_BLOCK, PS_MINB) rddl_plane_static(const __grid_constant__ VmArgs A) {
  plane_tile<PS_WPT, PS_BLOCK>(A, kG, kL, kDT);
}
#else
template <int WPT, int BLOCK>
__global__ void __launch_bounds__(BLOCK, WPT == 1 ? 4 : PLANE_MIN_BLOCKS) rddl_plane_interp(const __grid_constant__ VmArgs A,
                                                                          const __grid_constant__ PlaneProg G) {
  plane_tile<WPT, BLOCK>(A, G, A.L, A.dtype);
}
#endif
