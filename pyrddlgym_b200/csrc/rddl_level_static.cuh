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
};

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
                                 nullptr, 0, raw2);
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
  __device__ __forceinline__ void operator()(int64_t env, int32_t* idx, uint64_t* redval) const {
    uint64_t r[kLaunches[LI].nregs > 0 ? kLaunches[LI].nregs : 1];
    VmStaticCtx c = {A, kLaunches[LI], env, A.step, idx, redval, r};
    vm_seq<kLaunches[LI].pc_begin, kLaunches[LI].pc_begin, kLaunches[LI].pc_end>(c);
  }
};

#define RDDL_LEVEL_KERNEL(LI)                                                                            \
  extern "C" __global__ void __launch_bounds__(256) rddl_level_static_##LI(const __grid_constant__ VmArgs A) { \
    level_tile(A, kLaunches[LI], kRedOff, LevelStaticExec<LI>{A});                                         \
  }
PS_LEVELS(RDDL_LEVEL_KERNEL)
