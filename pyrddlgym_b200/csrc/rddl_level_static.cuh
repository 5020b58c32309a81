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

// envs a CTA of a fused env-tile kernel owns (the fused kernels further down; "level_static_data.inc" sets it)
#ifndef FUSE_TE
#define FUSE_TE 4
#endif
static_assert((256 / FUSE_TE) % 8 == 0, "pair steps walk blocks of 16 elements with 8 threads per env");

struct VmStaticCtx {
  const VmArgs& A;
  const RddlLaunch& L;
  int64_t env;
  uint64_t step;
  int32_t* idx;
  uint64_t* redval;
  uint64_t* r;
  const VmStage* stage;
  uint64_t* pre;        // values of the launch's preloaded LOADs (vm_preload), or null
  double* zpair;        // pair mode (fused env-tile kernels): second normals of the element pair's RANDNs, or null
  int half;             // pair mode: 0 = even element of the pair (draws both normals), 1 = odd element (takes the second)
  double* fwd;          // fused kernels: staged copy the NEXT phase gathers from, filled by this phase's STOREs of fwd_tensor
  int fwd_tensor;       //   (fuse_forward_tensor), or null / -1; element i of env slot fwd_e lives at fwd[i * FUSE_TE + fwd_e]
  int fwd_e;
};

// Preloaded LOADs. A launch's program reads its fluents one LOAD at a time, each immediately before its
// first use (and, in the generic instruction body, behind the branch that separates policy-drawn action
// fluents from stored ones), so an element pays one full memory latency per fluent, back to back -- ncu on
// Reservoir R = 1000: 9 long-scoreboard stalls per issued instruction, one per LOAD. The LOADs whose address
// depends on nothing but the element's scope coordinates -- outside every loop, no index registers, tensor
// not stored earlier in the same launch -- are therefore all issued up front (vm_preload: one batch of
// independent loads, one latency), and the program picks the values up from registers where it would have
// loaded them. Same loads, same values; only their place in the instruction stream changes.
constexpr int kMaxPreload = 12;
constexpr int kNumInsn = (int)(sizeof(kInsnWords) / sizeof(kInsnWords[0]));
constexpr int kNumLaunch = (int)(sizeof(kLaunches) / sizeof(kLaunches[0]));
__host__ __device__ constexpr int level_insn_len(int pc) { return (kInsnWords[pc][0] & 0xff) == OP_SUMLOOP ? 2 : 1; }
struct LevelPreTable { signed char slot[kNumInsn]; };      // per instruction: preload slot of an eligible LOAD, or -1
__host__ __device__ constexpr LevelPreTable level_make_pre() {
  LevelPreTable t{};
  for (int i = 0; i < kNumInsn; ++i) t.slot[i] = -1;
  for (int li = 0; li < kNumLaunch; ++li) {
    bool stored[RDDL_MAX_TENSORS] = {};
    int depth = 0, n = 0;
    for (int q = kLaunches[li].pc_begin; q < kLaunches[li].pc_end && q < kNumInsn; q += level_insn_len(q)) {
      const uint32_t o = kInsnWords[q][0] & 0xff;
      if (o == OP_LOOP || o == OP_CSRLOOP) ++depth;
      if (o == OP_ENDLOOP || o == OP_CSREND) --depth;
      if (o != OP_LOAD && o != OP_STORE) continue;
      const RddlAddr& ad = kAddrs[kInsnWords[q][3]];
      if (ad.tensor < 0 || ad.tensor >= RDDL_MAX_TENSORS) continue;
      if (o == OP_STORE) { stored[ad.tensor] = true; continue; }
      if (depth == 0 && ad.nreg == 0 && !stored[ad.tensor] && n < kMaxPreload) t.slot[q] = (signed char)n++;
    }
  }
  return t;
}
__device__ constexpr LevelPreTable kPre = level_make_pre();
__host__ __device__ constexpr bool level_preload_ok(int base, int pc) { return kPre.slot[pc] >= 0; }
__host__ __device__ constexpr int level_preload_slot(int base, int pc) { return kPre.slot[pc]; }

// Box-Muller pairs. A RANDN outside every loop whose variate offset is the launch's flat element index draws, for
// elements e and e + 8 (bit 3 of e clear), the two branches of ONE Box-Muller transform (rddl_randn_pair). A fused env-tile kernel
// gives both elements of a pair to the same thread, which computes the transform at the first element and keeps
// the sine branch in a register for its partner: half of the Philox blocks, logarithms and square roots.
constexpr int kMaxZPair = 4;
__host__ __device__ constexpr int level_launch_of(int pc) {
  for (int li = 0; li < kNumLaunch; ++li)
    if (pc >= kLaunches[li].pc_begin && pc < kLaunches[li].pc_end) return li;
  return 0;
}
__host__ __device__ constexpr bool level_addr_is_flat_index(const RddlAddr& ad, const RddlLaunch& L) {
  if (ad.nreg != 0 || ad.c0 != 0 || ad.naxis != L.rank) return false;
  for (int k = 0; k < L.rank; ++k) {
    int64_t want = 1;
    for (int j = k + 1; j < L.rank; ++j) want *= L.extent[j];
    bool found = false;
    for (int i = 0; i < ad.naxis; ++i) found = found || (ad.axis[i] == k && ad.coef[i] == want);
    if (!found) return false;
  }
  return true;
}
struct LevelPairTable { signed char slot[kNumInsn]; };   // per instruction: pair slot of an eligible RANDN, or -1
__host__ __device__ constexpr LevelPairTable level_make_pairs() {
  LevelPairTable t{};
  for (int i = 0; i < kNumInsn; ++i) t.slot[i] = -1;
  for (int li = 0; li < kNumLaunch; ++li) {
    if (kLaunches[li].kind != 0) continue;
    int depth = 0, n = 0;
    for (int q = kLaunches[li].pc_begin; q < kLaunches[li].pc_end && q < kNumInsn; q += level_insn_len(q)) {
      const uint32_t o = kInsnWords[q][0] & 0xff;
      if (o == OP_LOOP || o == OP_CSRLOOP) ++depth;
      if (o == OP_ENDLOOP || o == OP_CSREND) --depth;
      if (o == OP_RANDN && depth == 0 && n < kMaxZPair && level_addr_is_flat_index(kAddrs[kInsnWords[q][3]], kLaunches[li]))
        t.slot[q] = (signed char)n++;
    }
  }
  return t;
}
__device__ constexpr LevelPairTable kPairs = level_make_pairs();
__host__ __device__ constexpr bool level_launch_has_pair(int li) {
  for (int q = kLaunches[li].pc_begin; q < kLaunches[li].pc_end && q < kNumInsn; ++q)
    if (kPairs.slot[q] >= 0) return true;
  return false;
}

template <int BASE, int PC, int END>
__device__ __forceinline__ void vm_preload(VmStaticCtx& c) {
  if constexpr (PC < END) {
    if constexpr (level_preload_ok(BASE, PC)) {
      constexpr uint4 raw = {kInsnWords[PC][0], kInsnWords[PC][1], kInsnWords[PC][2], kInsnWords[PC][3]};
      const VmPools P = {kConsts, kAddrs, kRedSlots, kRedOff, kCsrs, kDT, kMayPolicy};
      constexpr RddlAddr ad = kAddrs[raw.w];
      if (c.stage && c.stage->pad && ad.tensor == c.stage->tensor && ad.naxis == 1 && ad.coef[0] == 1 && ad.c0 == 0 &&
          ad.nreg == 0) {
        // the element's own entry of the vector the CTA has staged whole: from shared memory, not from L1 / L2
        c.pre[level_preload_slot(BASE, PC)] = from_f(c.stage->ptr[c.idx[ad.axis[0]] * c.stage->mul + c.stage->add]);
      } else {
        int pc = 0;
        // (the LOAD's own instruction body, into its destination register -- dead at this point of the program --
        // and from there into the preload slot)
        vm_exec_one<RDDL_MAX_REGS, (int)OP_LOAD>(c.A, P, c.L, c.env, c.step, c.idx, c.redval, c.r, nullptr, nullptr, nullptr, pc,
                                                 raw, nullptr, 0);
        c.pre[level_preload_slot(BASE, PC)] = c.r[raw.x >> 16];
      }
    }
    vm_preload<BASE, PC + level_insn_len(PC), END>(c);
  }
}

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
    const VmPools P = {kConsts, kAddrs, kRedSlots, kRedOff, kCsrs, kDT, kMayPolicy};
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
    } else if constexpr (op == OP_LOAD && level_preload_ok(BASE, PC)) {
      if (c.pre) {
        c.r[raw.x >> 16] = c.pre[level_preload_slot(BASE, PC)];
      } else {
        int pc = 0;
        vm_exec_one<RDDL_MAX_REGS, (int)op>(c.A, P, c.L, c.env, c.step, c.idx, c.redval, c.r, nullptr, nullptr, nullptr, pc, raw,
                                   nullptr, 0);
      }
      vm_seq<BASE, PC + 1, END>(c);
    } else if constexpr (op == OP_RANDN && kPairs.slot[PC] >= 0) {
      if (c.zpair && !c.A.inj[ra]) {
        if (c.half == 0) {
          constexpr RddlAddr ad = kAddrs[imm];
          int64_t off = 0;                     // (= the flat element index, bit 3 clear here)
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < ad.naxis) off += ad.coef[i] * (int64_t)c.idx[ad.axis[i]];
          double z0, z1;
          rddl_randn_pair(c.A.rk, (uint64_t)(c.A.env_offset + c.env), c.step, (uint32_t)c.A.variates[ra].node,
                          rddl_randn_counter((uint64_t)off), &z0, &z1);
          c.r[raw.x >> 16] = from_f(z0);
          c.zpair[kPairs.slot[PC]] = z1;
        } else {
          c.r[raw.x >> 16] = from_f(c.zpair[kPairs.slot[PC]]);
        }
      } else {
        int pc = 0;
        vm_exec_one<RDDL_MAX_REGS, (int)op>(c.A, P, c.L, c.env, c.step, c.idx, c.redval, c.r, nullptr, nullptr, nullptr, pc, raw,
                                   nullptr, 0);
      }
      vm_seq<BASE, PC + 1, END>(c);
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
      if constexpr (op == OP_STORE) {
        // store forwarding: the vector the next phase of a fused kernel gathers from goes straight into its
        // staged copy (same value as the global store above; that phase then skips re-reading it)
        if (c.fwd && kAddrs[imm].tensor == c.fwd_tensor) c.fwd[c.idx[0] * FUSE_TE + c.fwd_e] = as_f(c.r[ra]);
      }
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
    VmStage st = {nullptr, -1, 0, 1, 0, 0};
    if (stage) { st = *stage; st.base = env * level_stage_elems(LI); }
    uint64_t pre[kMaxPreload];
    VmStaticCtx c = {A, kLaunches[LI], env, A.step, idx, redval, r, stage ? &st : nullptr, pre, nullptr, 0, nullptr, -1, 0};
    vm_preload<kLaunches[LI].pc_begin, kLaunches[LI].pc_begin, kLaunches[LI].pc_end>(c);
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
    if (threadIdx.x == 0) smem[n] = 0.0;      // the zero the padded CSR rows point at
    __syncthreads();
  }
};

template <int LI>
__device__ __forceinline__ void level_static_kernel(const VmArgs& A) {
  if constexpr (level_stage_tensor(LI) >= 0) {
    __shared__ double staged[(level_stage_elems(LI) > 0 ? level_stage_elems(LI) : 1) + 1];
    const VmStage st = {staged, level_stage_tensor(LI), 0, 1, 0, 1};
    level_tile(A, kLaunches[LI], kRedOff, LevelStaticExec<LI>{A, &st}, LevelStagePrologue<LI>{A, staged});
  } else {
    level_tile(A, kLaunches[LI], kRedOff, LevelStaticExec<LI>{A, nullptr});
  }
}

// ---------------------------------------------------------------------------------------
// Fused env-tile kernels: a run of consecutive scalar launches of the step plan (PS_FUSED lists
// them: first launch, count) executed by ONE kernel in which a CTA owns FUSE_TE whole envs for every
// launch of the run, with a CTA barrier where the unfused plan has a launch boundary. What this buys
// for gather-heavy levels (Reservoir: inflow(?r) = sum_{?u} CONNECT(?u, ?r) * released(?u)):
//   * thread t works on env t % FUSE_TE, elements t / FUSE_TE, + 256 / FUSE_TE, ...: the lanes of a warp
//     cover FUSE_TE envs x 32 / FUSE_TE consecutive elements -- every global access still touches only
//     full 32-byte sectors, but the CSR row pointers and column indices of a sparse sum are shared by
//     the FUSE_TE lanes of an element (FUSE_TE x fewer index sectors per request, shorter divergence);
//   * the gathered per-env vector is staged once per CTA in shared memory, transposed to
//     [element][env]: the FUSE_TE lanes of an element read consecutive words (SpMM over the env axis);
//   * intermediate fluents written by one launch are re-read by the next from L1 / L2, not DRAM, and
//     the per-env reductions and the scalar reward / termination program need no extra launches.
// Same instruction bodies and element order as the unfused kernels; only the order in which a
// per-env reduction combines its partial results differs (float sums agree to rounding).
// ---------------------------------------------------------------------------------------
#ifndef FUSE_MINB
#define FUSE_MINB 4
#endif
constexpr int kFuseMaxStage = 5632;      // doubles of shared memory for the staged gather operand (44 KB)

// tensor a CSR float SUMLOOP of launch li gathers from (per-env, real, small enough to stage), or -1
__host__ __device__ constexpr int fuse_stage_tensor(int li) {
  for (int pc = kLaunches[li].pc_begin; pc < kLaunches[li].pc_end; ++pc) {
    const uint32_t op = kInsnWords[pc][0] & 0xff, rb = kInsnWords[pc][1] >> 16, imm = kInsnWords[pc][3];
    if (op != OP_SUMLOOP) continue;
    if ((rb & 1u) && (rb & 4u)) {
      const RddlAddr& ad = kAddrs[imm];
      if (ad.tensor >= 0 && ad.env_stride > 0 && (ad.env_stride + 1) * FUSE_TE <= kFuseMaxStage && kDT[ad.tensor] == DT_REAL && ad.nreg == 0)
        return ad.tensor;
    }
    ++pc;
  }
  return -1;
}
__host__ __device__ constexpr int64_t fuse_stage_elems(int li) {
  for (int pc = kLaunches[li].pc_begin; pc < kLaunches[li].pc_end; ++pc) {
    const uint32_t op = kInsnWords[pc][0] & 0xff, imm = kInsnWords[pc][3];
    if (op == OP_SUMLOOP && kAddrs[imm].tensor == fuse_stage_tensor(li)) return kAddrs[imm].env_stride;
  }
  return 0;
}

// Store forwarding between the phases of a fused kernel: the tensor launch li + 1 stages, when launch li (which must
// not use the staging buffer itself) writes every element of it with exactly one STORE outside loops at the element's
// own index -- that STORE then also fills the staged copy, and launch li + 1 skips its copy loop. -1: none.
__host__ __device__ constexpr int fuse_forward_tensor(int li, int last) {
  if (li + 1 > last || li + 1 >= kNumLaunch) return -1;
  const int t = fuse_stage_tensor(li + 1);
  if (t < 0 || fuse_stage_tensor(li) >= 0) return -1;
  if (kLaunches[li].rank != 1 || kLaunches[li].E != fuse_stage_elems(li + 1)) return -1;
  int n = 0, depth = 0;
  for (int q = kLaunches[li].pc_begin; q < kLaunches[li].pc_end && q < kNumInsn; q += level_insn_len(q)) {
    const uint32_t o = kInsnWords[q][0] & 0xff;
    if (o == OP_LOOP || o == OP_CSRLOOP) ++depth;
    if (o == OP_ENDLOOP || o == OP_CSREND) --depth;
    if (o == OP_STORE && kAddrs[kInsnWords[q][3]].tensor == t) {
      if (depth != 0 || !level_addr_is_flat_index(kAddrs[kInsnWords[q][3]], kLaunches[li])) return -1;
      ++n;
    }
  }
  return n == 1 ? t : -1;
}

template <int LI, int FIRST, int LAST>
__device__ __forceinline__ void fused_phase(const VmArgs& A, double* staged, uint64_t (*sm)[8][FUSE_TE]) {
  constexpr RddlLaunch L = kLaunches[LI];
  constexpr int FWD = fuse_forward_tensor(LI, LAST);                                  // filled for the next phase
  constexpr bool FILLED = LI > FIRST && fuse_forward_tensor(LI - 1, LAST) >= 0 &&     // filled by the previous one
                          fuse_forward_tensor(LI - 1, LAST) == fuse_stage_tensor(LI);
  const int tid = threadIdx.x;
  const int e = tid % FUSE_TE;
  const int64_t env0 = (int64_t)blockIdx.x * FUSE_TE, env = env0 + e;
  const bool env_ok = env < A.n_envs;
  VmStage st = {nullptr, -1, 0, 1, 0, 0};
  if constexpr (fuse_stage_tensor(LI) >= 0) {
    // the FUSE_TE envs' slices of the gathered tensor, transposed to [element][env]
    constexpr int64_t n = fuse_stage_elems(LI);
    const double* src = reinterpret_cast<const double*>(A.tensors[fuse_stage_tensor(LI)]);
    // (thread t always copies env t % FUSE_TE: 256 is a multiple of FUSE_TE; 32-bit indices, four loads in flight)
    const double* srce = src + (env_ok ? env * n : 0);
    if constexpr (!FILLED) {
#pragma unroll 4
      for (int i = tid; i < (int)(n * FUSE_TE); i += 256) staged[i] = env_ok ? srce[(uint32_t)i / FUSE_TE] : 0.0;
    }
    if (tid < FUSE_TE) staged[n * FUSE_TE + tid] = 0.0;      // the zeros the padded CSR rows point at
    __syncthreads();
    st = VmStage{staged, fuse_stage_tensor(LI), env * n, FUSE_TE, e, 1};
  }
  int32_t idx[RDDL_MAX_IDX];
  uint64_t redval[RDDL_MAX_RED];
#pragma unroll
  for (int j = 0; j < RDDL_MAX_RED; ++j) redval[j] = j < L.nred ? red_identity(L.red_op[j]) : 0;
  // elements per thread step: pairs (2p, 2p + 1) when the launch draws Box-Muller pairs, else single elements
  constexpr int EPS = level_launch_has_pair(LI) ? 2 : 1;
  // (the trip count is the same for every thread of the CTA)
  constexpr int64_t kSteps = (L.E + (256 / FUSE_TE) * EPS - 1) / ((256 / FUSE_TE) * EPS);
  {
    for (int64_t step_i = 0; step_i < kSteps; ++step_i) {
      // single elements: thread j = tid / FUSE_TE of the 256 / FUSE_TE takes element j of the step's block; pairs:
      // of each 16 consecutive elements, eight threads take elements 0..7 and then their partners 8..15
      const int j = tid / FUSE_TE;
      const int64_t e0 = EPS == 1 ? j + step_i * (256 / FUSE_TE) : step_i * (2 * (256 / FUSE_TE)) + (j >> 3) * 16 + (j & 7);
      double zpair[kMaxZPair];
#pragma unroll
      for (int half = 0; half < EPS; ++half) {
        const int64_t elem = e0 + half * 8;
        const bool active = env_ok && elem < L.E;
        if (active) {
          int64_t rem = elem;
#pragma unroll
          for (int k = RDDL_MAX_IDX - 1; k >= 0; --k) {
            if (k < L.rank) {
              idx[k] = (int32_t)(rem % L.extent[k]);
              rem /= L.extent[k];
            }
          }
          uint64_t r[L.nregs > 0 ? L.nregs : 1];
          uint64_t pre[kMaxPreload];
          VmStaticCtx c = {A, kLaunches[LI], env, A.step, idx, redval, r, st.tensor >= 0 ? &st : nullptr, pre,
                           EPS == 2 ? zpair : nullptr, half, FWD >= 0 ? staged : nullptr, FWD, e};
          vm_preload<L.pc_begin, L.pc_begin, L.pc_end>(c);
          vm_seq<L.pc_begin, L.pc_begin, L.pc_end>(c);
        }
      }
    }
  }
  if constexpr (L.nred > 0) {
    // per-env reductions: lanes of the same env (stride FUSE_TE), then the 8 warps in order
#pragma unroll
    for (int j = 0; j < L.nred; ++j) {
      uint64_t v = redval[j];
#pragma unroll
      for (int o = 16; o >= FUSE_TE; o >>= 1) v = red_combine(L.red_op[j], v, __shfl_xor_sync(0xffffffffu, v, o, 32));
      if ((tid & 31) < FUSE_TE) sm[j][tid >> 5][e] = v;
    }
    __syncthreads();
    if (tid < FUSE_TE && env_ok) {
#pragma unroll
      for (int j = 0; j < L.nred; ++j) {
        uint64_t v = sm[j][0][e];
        for (int w = 1; w < 8; ++w) v = red_combine(L.red_op[j], v, sm[j][w][e]);
        // (the slot keeps the partial-result layout of the unfused plan: total in the first, identities after)
        uint64_t* dst = A.red + kRedOff[L.red_slot[j]] * A.n_envs + env * kRedSlots[L.red_slot[j]].chunks;
        dst[0] = v;
        for (int c2 = 1; c2 < kRedSlots[L.red_slot[j]].chunks; ++c2) dst[c2] = red_identity(L.red_op[j]);
      }
    }
  }
  __syncthreads();      // launch boundary of the unfused plan: this CTA's stores are visible to its next phase
}

template <int LI, int FIRST, int LAST>
__device__ __forceinline__ void fused_phases(const VmArgs& A, double* staged, uint64_t (*sm)[8][FUSE_TE]) {
  if constexpr (LI <= LAST) {
    fused_phase<LI, FIRST, LAST>(A, staged, sm);
    fused_phases<LI + 1, FIRST, LAST>(A, staged, sm);
  }
}

// doubles of staging the run [first, first + count) needs
__host__ __device__ constexpr int64_t fuse_group_stage(int first, int count) {
  int64_t m = 1;
  for (int li = first; li < first + count; ++li)
    if (fuse_stage_tensor(li) >= 0 && (fuse_stage_elems(li) + 1) * FUSE_TE > m) m = (fuse_stage_elems(li) + 1) * FUSE_TE;
  return m;
}

#define RDDL_FUSED_KERNEL(FIRST, COUNT)                                                                          \
  extern "C" __global__ void __launch_bounds__(256, FUSE_MINB) rddl_level_fused_##FIRST(const __grid_constant__ VmArgs A) { \
    __shared__ double staged[fuse_group_stage(FIRST, COUNT)];                                                    \
    __shared__ uint64_t sm[RDDL_MAX_RED][8][FUSE_TE];                                                             \
    fused_phases<FIRST, FIRST, FIRST + COUNT - 1>(A, staged, sm);                                                \
  }
#ifdef PS_FUSED
PS_FUSED(RDDL_FUSED_KERNEL)
#endif

#define RDDL_LEVEL_KERNEL(LI)                                                                            \
  extern "C" __global__ void __launch_bounds__(256) rddl_level_static_##LI(const __grid_constant__ VmArgs A) { \
    level_static_kernel<LI>(A);                                                                            \
  }
PS_LEVELS(RDDL_LEVEL_KERNEL)
