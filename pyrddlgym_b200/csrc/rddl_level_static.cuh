// Run-time specialised scalar kernels: the level interpreter of rddl_device.cuh with the program of
// each scalar launch unrolled at compile time. "level_static_data.inc" (generated per program by
// rddl_jit.h) holds every instruction word, the const / address / CSR / reduction-slot pools and the
// launches as constants; rddl_level_static_<i> is the kernel of launch i. Each instruction runs
// through the same vm_exec_one as the interpreter, with constant operand fields: the dispatch
// switch folds, the virtual registers r[] become machine registers, address descriptors fold into
// immediate offsets, and dense / CSR loops become real loops around their (unrolled) bodies.
#pragma once
#include "rddl_device.cuh"
#include "level_static_data.inc"

struct VmStaticCtx {
  const VmArgs& A;
  const RddlLaunch& L;
  int64_t env;
  uint64_t step;
  int32_t* idx;
  uint64_t* redval;
  uint64_t* r;
  const VmStage* stage;
};

// Shared-memory staging of a gathered vector: when a launch's CTA works on one env (G = 256) and its
// program sums real values of a per-env tensor through a CSR list (SUMLOOP), the CTA first copies that
// env's slice of the tensor (<= 32 KB) into shared memory; the scattered 8-byte gathers then hit
// shared memory banks instead of occupying the L1 wavefront pipe with 32 sectors per request.
// level_stage_tensor<LI>() = tensor id to stage for launch LI, or -1.
constexpr int kStageMaxElems = 4096;
__host__ __device__ constexpr int level_stage_tensor(int li) {
  if (kLaunches[li].G != 256) return -1;
  for (int pc = kLaunches[li].pc_begin; pc < kLaunches[li].pc_end; ++pc) {
    const uint32_t op = kInsnWords[pc][0] & 0xff, rb = kInsnWords[pc][1] >> 16, imm = kInsnWords[pc][3];
    if (op != OP_SUMLOOP) continue;
    if ((rb & 1u) && (rb & 4u)) {     // CSR list, floating-point sum
      const RddlAddr& ad = kAddrs[imm];
      if (ad.tensor >= 0 && ad.env_stride > 0 && ad.env_stride <= kStageMaxElems && kDT[ad.tensor] == DT_REAL && ad.nreg == 0)
        return ad.tensor;
    }
    ++pc;   // skip the data word
  }
  return -1;
}
__host__ __device__ constexpr int64_t level_stage_elems(int li) {
  for (int pc = kLaunches[li].pc_begin; pc < kLaunches[li].pc_end; ++pc) {
    const uint32_t op = kInsnWords[pc][0] & 0xff, imm = kInsnWords[pc][3];
    if (op == OP_SUMLOOP && kAddrs[imm].tensor == level_stage_tensor(li)) return kAddrs[imm].env_stride;
  }
  return 0;
}

// Instructions [PC, END) of the launch whose program starts at BASE (jump targets are launch-relative).
template <int BASE, int PC, int END>
__device__ __forceinline__ void vm_seq(VmStaticCtx& c) {
  if constexpr (PC < END) {
    constexpr uint4 raw = {kInsnWords[PC][0], kInsnWords[PC][1], kInsnWords[PC][2], kInsnWords[PC][3]};
    constexpr uint32_t op = raw.x & 0xff, ra = raw.y & 0xffff, rb = raw.y >> 16, imm = raw.w;
    const VmPools P = {kConsts, kAddrs, kRedSlots, kRedOff, kCsrs, kDT};
    if constexpr (op == OP_LOOP) {
      // dense loop over axis slot ra, extent imm; rb = pc after the matching ENDLOOP
      constexpr int AFTER = BASE + (int)rb;
      for (c.idx[ra] = 0; c.idx[ra] < (int32_t)imm; ++c.idx[ra]) vm_seq<BASE, PC + 1, AFTER - 1>(c);
      vm_seq<BASE, AFTER, END>(c);
    } else if constexpr (op == OP_CSRLOOP) {
      // sparse loop over the non-zeros of CSR list imm in the row selected by the scope coordinates
      constexpr int AFTER = BASE + (int)rb;
      constexpr RddlCsr cs = kCsrs[imm];
      int64_t row = cs.row_c0;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < cs.row_naxis) row += cs.row_coef[i] * (int64_t)c.idx[cs.row_axis[i]];
      const int32_t j0 = __ldg(c.A.pool + cs.rowptr_off + row), j1 = __ldg(c.A.pool + cs.rowptr_off + row + 1);
      for (int32_t j = j0; j < j1; ++j) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < cs.ncols) c.idx[cs.slot[q]] = __ldg(c.A.pool + cs.col_off[q] + j);
        vm_seq<BASE, PC + 1, AFTER - 1>(c);
      }
      vm_seq<BASE, AFTER, END>(c);
    } else if constexpr (op == OP_SUMLOOP) {
      constexpr uint4 raw2 = {kInsnWords[PC + 1][0], kInsnWords[PC + 1][1], kInsnWords[PC + 1][2], kInsnWords[PC + 1][3]};
      int pc = 0;
      vm_exec_one<RDDL_MAX_REGS, (int)op>(c.A, P, c.L, c.env, c.step, c.idx, c.redval, c.r, nullptr, nullptr, nullptr, pc, raw,
                                 nullptr, 0, raw2, nullptr, nullptr, c.stage);
      vm_seq<BASE, PC + 2, END>(c);
    } else {
      int pc = 0;
      vm_exec_one<RDDL_MAX_REGS, (int)op>(c.A, P, c.L, c.env, c.step, c.idx, c.redval, c.r, nullptr, nullptr, nullptr, pc, raw,
                                 nullptr, 0);
      vm_seq<BASE, PC + 1, END>(c);
    }
  }
}

template <int LI>
struct LevelStaticExec {
  const VmArgs& A;
  const VmStage* stage;
  __device__ __forceinline__ void operator()(int64_t env, int32_t* idx, uint64_t* redval) const {
    uint64_t r[kLaunches[LI].nregs > 0 ? kLaunches[LI].nregs : 1];
    VmStage st = {nullptr, -1, 0};
    if (stage) { st = *stage; st.base = env * level_stage_elems(LI); }
    VmStaticCtx c = {A, kLaunches[LI], env, A.step, idx, redval, r, stage ? &st : nullptr};
    vm_seq<kLaunches[LI].pc_begin, kLaunches[LI].pc_begin, kLaunches[LI].pc_end>(c);
  }
};

// Copies the CTA's env slice of the staged tensor into shared memory (all 256 threads, coalesced).
template <int LI>
struct LevelStagePrologue {
  const VmArgs& A;
  double* smem;
  __device__ __forceinline__ void operator()(int64_t env, bool env_ok) const {
    constexpr int64_t n = level_stage_elems(LI);
    const double* src = reinterpret_cast<const double*>(A.tensors[level_stage_tensor(LI)]) + env * n;
    for (int i = threadIdx.x; i < (int)n; i += 256) smem[i] = env_ok ? src[i] : 0.0;
    __syncthreads();
  }
};

template <int LI>
__device__ __forceinline__ void level_static_kernel(const VmArgs& A) {
  if constexpr (level_stage_tensor(LI) >= 0) {
    __shared__ double staged[level_stage_elems(LI) > 0 ? level_stage_elems(LI) : 1];
    const VmStage st = {staged, level_stage_tensor(LI), 0};
    level_tile(A, kLaunches[LI], kRedOff, LevelStaticExec<LI>{A, &st}, LevelStagePrologue<LI>{A, staged});
  } else {
    level_tile(A, kLaunches[LI], kRedOff, LevelStaticExec<LI>{A, nullptr});
  }
}

#define RDDL_LEVEL_KERNEL(LI)                                                                            \
  extern "C" __global__ void __launch_bounds__(256) rddl_level_static_##LI(const __grid_constant__ VmArgs A) { \
    level_static_kernel<LI>(A);                                                                            \
  }
PS_LEVELS(RDDL_LEVEL_KERNEL)
