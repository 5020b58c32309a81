// Device side of the scalar level interpreter: launch arguments and the per-element bytecode
// executor. Shared by the ahead-of-time build (rddl_vm.cu) and by the run-time specialised
// plane kernels (rddl_jit.cuh compiles rddl_plane.cuh with NVRTC, which includes this file).
#pragma once
#include "philox.cuh"
#include "rddl_plan.h"
#include "rddl_samplers.cuh"

#if defined(__CUDACC_RTC__) && !defined(INFINITY)   // NVRTC has no <math.h>; same bit patterns as its macros
#define INFINITY (__int_as_float(0x7f800000))
#define NAN (__int_as_float(0x7fc00000))
#endif

// A CUtensorMap (128 opaque bytes, 64-byte aligned; cuda.h is not available to the NVRTC build). The
// one of VmArgs views the op table's int32 pool as a 1-D tensor with a box of RDDL_TMA_BOX words: the
// plane kernels stage the pre-packed non-fluent planes of a tile in shared memory with it (TMA).
#define RDDL_TMA_BOX 256
struct alignas(64) RddlTensorMap { unsigned long long opaque[16]; };

struct VmArgs {
  const RddlInsn* insns;
  const uint64_t* consts;
  const RddlAddr* addrs;
  const RddlCsr* csrs;
  const int32_t* pool;
  const RddlVariate* variates;
  const RddlRedSlot* redslots;
  const int64_t* red_off;      // per slot: element offset (in units of n_envs) into `red`
  uint64_t* red;
  uint32_t* status;
  int64_t n_envs, env_offset;
  uint64_t seed, step;
  RddlPhiloxKeys rk;           // Philox round keys of `seed` (rddl_philox_round_keys; filled by the host)
  RddlLaunch L;
  void* tensors[RDDL_MAX_TENSORS];
  uint8_t dtype[RDDL_MAX_TENSORS];
  const double* inj[RDDL_MAX_VARIATES];
  // multi-step rollouts fused into one plane launch (n_steps > 1): per-tensor byte stride from
  // one step's inputs to the next (action sequences [T, n_envs, ...]), discounted return
  // accumulators and optional per-step rewards [T, n_envs]
  int32_t n_steps, pad0;
  double gamma;
  double* returns;
  double* rewards;
  int64_t tstride[RDDL_MAX_TENSORS];
  VmPolicy pol;                // pol.n > 0: action fluents listed there are drawn in-kernel (device-side policy)
  RddlTensorMap pool_map;      // tensor map over `pool` (valid when pool_map_ok)
  int32_t pool_map_ok, pad1;
};

// mbarrier / TMA primitives of the plane kernels' non-fluent staging
__device__ __forceinline__ uint32_t rddl_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rddl_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rddl_smem_addr(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void rddl_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rddl_smem_addr(bar)), "r"(bytes) : "memory");
}
// Returns 0 once the phase has completed. No "memory" clobber (it would pin every load of the kernel
// on one side of the wait): the caller orders its reads of the copied data after the wait by making
// their address depend on the returned value.
__device__ __forceinline__ uint32_t rddl_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = rddl_smem_addr(bar);
  for (uint32_t spins = 0;; ++spins) {      // (bounded: a copy that never lands traps instead of hanging the GPU)
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(a), "r"(parity));
    if (ok) return ok - 1u;
    if (spins > (1u << 24)) __trap();
  }
}
// one box of RDDL_TMA_BOX words of the 1-D tensor behind `map`, starting at element `c0`, into shared memory
__device__ __forceinline__ void rddl_tma_load_1d(void* dst, const RddlTensorMap* map, uint64_t* bar, int32_t c0) {
  asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];"
               ::"r"(rddl_smem_addr(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(rddl_smem_addr(bar)), "r"(c0) : "memory");
}

__device__ __forceinline__ double as_f(uint64_t v) { return __longlong_as_double((long long)v); }
__device__ __forceinline__ uint64_t from_f(double v) { return (uint64_t)__double_as_longlong(v); }

// simulator.py:1425-1441 (Stirling series after shifting the argument to >= 7)
__device__ double rddl_lngamma(double x) {
  double corr = 0.0;
  for (int i = 0; i < 4; ++i) {
    if (x < 7.0) {
      corr += log(x) + log(x + 1.0);
      x += 2.0;
    }
  }
  double y2 = x * x;
  double res = (x - 0.5) * log(x) - x + 0.5 * log(2.0 * 3.141592653589793) +
               1.0 / (12.0 * x) *
                   (1.0 + 1.0 / (30.0 * y2) *
                              (-1.0 + 1.0 / (7.0 * y2 / 2.0) *
                                          (1.0 + 1.0 / (4.0 * y2 / 3.0) *
                                                     (-1.0 + 1.0 / (99.0 * y2 / 140.0) *
                                                                 (1.0 + 1.0 / (910.0 * y2 / 3.0))))));
  return res - corr;
}

__device__ __forceinline__ int64_t ipow64(int64_t a, int64_t b) {
  if (b < 0) return a == 1 ? 1 : (a == -1 ? ((b & 1) ? -1 : 1) : 0);
  int64_t r = 1;
  while (b) {
    if (b & 1) r *= a;
    a *= a;
    b >>= 1;
  }
  return r;
}

__device__ __forceinline__ uint64_t red_identity(int op) {
  switch (op) {
    case RED_SUM_I: return 0;
    case RED_SUM_F: return from_f(0.0);
    case RED_PROD_I: return 1;
    case RED_PROD_F: return from_f(1.0);
    case RED_MIN_I: return (uint64_t)INT64_MAX;
    case RED_MIN_F: return from_f(INFINITY);
    case RED_MAX_I: return (uint64_t)INT64_MIN;
    case RED_MAX_F: return from_f(-INFINITY);
    case RED_AND: return 1;
    default: return 0;
  }
}

// np.minimum / np.maximum: a NaN in either operand propagates. (a < b || a != a) ? a : b is the expression of
// NumPy's own loops: a NaN `a` is taken by the second test, a NaN `b` because the comparison is false. Two DSETP
// + four FSEL, against 13 instructions for the fmin() + isnan() form (no f64 min / max instruction on sm_100).
__device__ __forceinline__ double np_fmin(double a, double b) { return (a < b || a != a) ? a : b; }
__device__ __forceinline__ double np_fmax(double a, double b) { return (a > b || a != a) ? a : b; }

__device__ __forceinline__ uint64_t red_combine(int op, uint64_t a, uint64_t b) {
  switch (op) {
    case RED_SUM_I: return (uint64_t)((int64_t)a + (int64_t)b);
    case RED_SUM_F: return from_f(as_f(a) + as_f(b));
    case RED_PROD_I: return (uint64_t)((int64_t)a * (int64_t)b);
    case RED_PROD_F: return from_f(as_f(a) * as_f(b));
    case RED_MIN_I: return (uint64_t)min((long long)a, (long long)b);
    case RED_MIN_F: return from_f(np_fmin(as_f(a), as_f(b)));
    case RED_MAX_I: return (uint64_t)max((long long)a, (long long)b);
    case RED_MAX_F: return from_f(np_fmax(as_f(a), as_f(b)));
    case RED_AND: return a & b;
    default: return a | b;
  }
}

// ---------------------------------------------------------------------------------------
// Device-side random policy: the batched counterpart of RandomAgent.sample_action
// (pyRDDLGym/core/policy.py:166-178) evaluated inside the kernels. Every draw is a pure function of
// (seed, global env id, step, action fluent, grounding), restated in oracle/philox_np.py:policy_actions.
// ---------------------------------------------------------------------------------------

// Bernoulli(p) of one cell under the bit-sliced mapping of the plane kernels (rddl_plane.cuh: PBERN;
// oracle/philox_np.py:bernoulli_bitsliced): the cell compares the top 8 bits of thr = floor(p * 2^64),
// MSB first, with bit (cell % 32) of the Philox words of (cell / 32, level); ties take one 64-bit draw.
__device__ __forceinline__ bool rddl_bern_bitsliced_cell(uint64_t seed, uint32_t env, uint32_t step, uint32_t node,
                                                         uint64_t thr, bool always, uint64_t cell) {
  if (always) return true;
  if (thr == 0) return false;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t w = (uint32_t)(cell >> 5), bit = (uint32_t)cell & 31u;
#pragma unroll 1
  for (uint32_t b = 0; b < 2; ++b) {
    const uint4 r4 = rddl_philox4x32_10(make_uint4(w, env, step, (node & 0xffffu) | ((0x8000u + b) << 16)), key);
    const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t rb = (rw[q] >> bit) & 1u;
      const uint32_t tb = (uint32_t)(thr >> (63 - (4 * b + q))) & 1u;
      if (rb != tb) return tb != 0;
    }
  }
  const uint4 r4 = rddl_philox4x32_10(make_uint4((uint32_t)cell, env, step, (node & 0xffffu) | (0xC000u << 16)), key);
  return (((uint64_t)r4.x << 32) | r4.y) < (thr << 8);
}

// i-th candidate of the max-nondef selection: uniform in [0, j]
__device__ __forceinline__ int64_t rddl_policy_pick(uint64_t seed, uint32_t env, uint32_t step, uint32_t i, int64_t j) {
  const uint4 r4 = rddl_philox4x32_10(make_uint4(i, env, step, RDDL_POLICY_SELECT_NODE),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return (int64_t)__umul64hi(((uint64_t)r4.x << 32) | r4.y, (uint64_t)(j + 1));
}

// The k groundings of [0, total) that keep their draw this step (Floyd's sampling without replacement
// over the candidates pick[i] ~ U[0, total - k + i]): sel[i] = pick[i] unless already chosen, then
// total - k + i. pick[] may alias sel[].
__device__ __forceinline__ void rddl_policy_floyd(const int32_t* pick, int32_t* sel, int k, int64_t total) {
  for (int i = 0; i < k; ++i) {
    int32_t t = pick[i];
    for (int q = 0; q < i; ++q)
      if (sel[q] == t) { t = (int32_t)(total - k + i); break; }
    sel[i] = t;
  }
}

// Value (raw 64 bits) of grounding `elem` of policy entry e for (env, step); sel = the env's selected
// groundings when the policy has a max-nondef rule (k > 0), else unused.
__device__ __forceinline__ uint64_t rddl_policy_value(const VmPolicy& pol, const VmPolicyEntry& e, uint64_t seed,
                                                      uint64_t env_global, uint64_t step, int64_t elem,
                                                      const int32_t* sel) {
  if (e.kind == POL_DEFAULT) return e.dflt;
  if (pol.k > 0) {
    bool chosen = false;
    for (int i = 0; i < pol.k; ++i) chosen |= (int64_t)sel[i] == e.base + elem;
    if (!chosen) return e.dflt;
  }
  const uint32_t node = e.node;
  if (e.kind == POL_BERNOULLI)     // [thr_lo <= U < thr] on the cell's bit-sliced uniform (both compares see the same bits)
    return rddl_bern_bitsliced_cell(seed, (uint32_t)env_global, (uint32_t)step, node, e.thr, e.always != 0, (uint64_t)elem) &&
           !rddl_bern_bitsliced_cell(seed, (uint32_t)env_global, (uint32_t)step, node, e.thr_lo, false, (uint64_t)elem);
  const uint4 x = rddl_philox_draw(seed, env_global, step, node, (uint64_t)elem);
  const double u = rddl_u53(x.x, x.y);
  if (e.kind == POL_UNIFORM_REAL) return (uint64_t)__double_as_longlong(__dadd_rn(e.lo, __dmul_rn(__dsub_rn(e.hi, e.lo), u)));
  // POL_UNIFORM_INT: lo + floor(u * (hi - lo + 1)), never above hi
  const int64_t lo = (int64_t)e.lo, hi = (int64_t)e.hi;
  int64_t v = lo + (int64_t)floor(__dmul_rn(u, (double)(hi - lo + 1)));
  return (uint64_t)(v > hi ? hi : v);
}

// Standard normals come in Box-Muller pairs: elements e and e + 8 of a variate (e with bit 3 clear) are the two
// branches of one transform of the Philox block of counter q(e) = e with bit 3 removed -- z0 = R cos(2 pi u2) for
// e, z1 = R sin(2 pi u2) for e + 8, R = sqrt(-2 ln(1 - u1)). A kernel whose thread owns both elements (fused
// env-tile kernels, rddl_level_static.cuh) pays for one block, one logarithm and one square root per pair;
// thread-per-element kernels compute the pair and keep their branch (same bits). Partners are 8 apart, not
// adjacent, so that the lanes of such a kernel still walk consecutive elements (full sectors) in each half.
// sincospi(2 u): no argument reduction of 2 pi u (no Payne-Hanek path, no stack slot for its quadrant).
__device__ __forceinline__ uint64_t rddl_randn_counter(uint64_t e) { return ((e >> 4) << 3) | (e & 7u); }
__device__ __forceinline__ bool rddl_randn_branch(uint64_t e) { return ((e >> 3) & 1u) != 0; }
__device__ __forceinline__ void rddl_randn_pair(const RddlPhiloxKeys& rk, uint64_t env, uint64_t step, uint32_t node, uint64_t q,
                                                double* z0, double* z1) {
  const uint4 x = rddl_philox_draw(rk, env, step, node, q);
  const double rad = sqrt(-2.0 * log(1.0 - rddl_u53(x.x, x.y)));
  double sn, cs;
  sincospi(2.0 * rddl_u53(x.z, x.w), &sn, &cs);
  *z0 = rad * cs;
  *z1 = rad * sn;
}

// Entry j of a column-index list: int32, or uint16 (padded lists whose indices fit)
__device__ __forceinline__ int32_t rddl_csr_col(const int32_t* cols, int32_t j, bool u16) {
  return u16 ? (int32_t)__ldg(reinterpret_cast<const uint16_t*>(cols) + j) : __ldg(cols + j);
}

// The constant / address-descriptor / reduction-slot pools a program executes from: device copies
// of the op table (interpreter) or the constants of a specialised build.
struct VmPools {
  const uint64_t* consts;
  const RddlAddr* addrs;
  const RddlRedSlot* redslots;
  const int64_t* red_off;
  const RddlCsr* csrs;
  const uint8_t* dtype;      // per tensor id
  const uint8_t* may_policy; // per tensor id: may a device-side policy drive it (null: not known here -- ask the descriptor)
};

// A per-env tensor slice staged in shared memory by the CTA (specialised scalar kernels whose CTA
// works on a single env): SUMLOOP gathers from `tensor` read ptr[offset - base] instead of global memory.
struct VmStage {
  const double* ptr;
  int tensor;
  int64_t base;
  int mul, add;     // staged index of element offset o (relative to base) = o * mul + add  (env-tiled staging: mul = envs per tile)
  int pad;          // 1: the staged slice is followed by one zero per env (element index = slice length): padded CSR rows
};

// One instruction (raw = its four 32-bit words) of the program of launch L for one (env, element):
// idx[] holds the element's scope coordinates, r[] the virtual registers. `pc` is advanced by
// control-flow ops (prog = the launch's instruction stream; null for straight-line specialised
// programs, which contain none). cap (optional): values stored to tensors cap_tensor, cap_tensor + 1.
// OPC >= 0: the opcode as a template constant (specialised builds: only that case is compiled in).
template <int NREG, int OPC = -1>
__device__ __forceinline__ void vm_exec_one(const VmArgs& A, const VmPools& P, const RddlLaunch& L, const int64_t env,
                                            const uint64_t step, int32_t* idx, uint64_t* redval, uint64_t* r,
                                            int32_t* jcur, int32_t* jend, const RddlInsn* prog, int& pc,
                                            const uint4 raw, uint64_t* cap, const int cap_tensor,
                                            const uint4 raw2_static = uint4{0u, 0u, 0u, 0u},
                                            const uint64_t* red_local = nullptr, const RddlLaunch* red_src = nullptr,
                                            const VmStage* stage = nullptr) {
  const uint32_t op = OPC >= 0 ? (uint32_t)OPC : (raw.x & 0xff);
  const uint32_t dst = raw.x >> 16;
  const uint32_t ra = raw.y & 0xffff, rb = raw.y >> 16;
  const uint32_t rc = raw.z & 0xffff;
  const uint32_t imm = raw.w;
#define RA r[ra]
#define RB r[rb]
#define FA as_f(r[ra])
#define FB as_f(r[rb])
#define IA ((int64_t)r[ra])
#define IB ((int64_t)r[rb])
#define SETF(v) r[dst] = from_f(v)
#define SETI(v) r[dst] = (uint64_t)(int64_t)(v)
  switch (op) {
    case OP_NOP: break;
    case OP_CONST: r[dst] = P.consts[imm]; break;
    case OP_AXIS: SETI(idx[ra]); break;
    case OP_MOV: r[dst] = RA; break;
    case OP_I2F: SETF((double)IA); break;
    case OP_F2I: SETI(__double2ll_rz(FA)); break;
    case OP_LOAD:
    case OP_STORE:
    case OP_RANDU:
    case OP_RANDN:
    case OP_RANDE:
    case OP_RANDC: {
      const RddlAddr* ad = P.addrs + imm;
      int64_t off = ad->c0;
      const int na = ad->naxis;
      for (int i = 0; i < na; ++i) off += ad->coef[i] * (int64_t)idx[ad->axis[i]];
      const int nr = ad->nreg;
      for (int i = 0; i < nr; ++i) {
        int64_t v = (int64_t)r[ad->reg[i]];
        const int64_t ext = ad->rextent[i];
        if (v < 0 || v >= ext) {
          atomicOr(A.status + env, 8u);
          v = v < 0 ? 0 : ext - 1;
        }
        off += v * ad->rstride[i];
      }
      const int t = ad->tensor;
      if (op == OP_LOAD) {
        if (A.pol.n > 0 && t >= 0 && (!P.may_policy || P.may_policy[t]) && A.pol.of_tensor[t]) {
          // action fluent under a device-side policy: drawn here instead of read from HBM
          r[dst] = rddl_policy_value(A.pol, A.pol.e[A.pol.of_tensor[t] - 1], A.seed, (uint64_t)(A.env_offset + env), step,
                                     off, A.pol.sel + env * A.pol.k);
          break;
        }
        off += env * ad->env_stride;
        const void* base = A.tensors[t];
        const int dt = P.dtype[t];
        if (dt == DT_BOOL) r[dst] = reinterpret_cast<const uint8_t*>(base)[off] != 0;
        else if (dt == DT_PACKED) r[dst] = (reinterpret_cast<const uint32_t*>(base)[off >> 5] >> (off & 31)) & 1u;
        else r[dst] = reinterpret_cast<const uint64_t*>(base)[off];
      } else if (op == OP_STORE) {
        off += env * ad->env_stride;
        void* base = A.tensors[t];
        const int dt = P.dtype[t];
        if (dt == DT_BOOL) reinterpret_cast<uint8_t*>(base)[off] = RA != 0;
        else reinterpret_cast<uint64_t*>(base)[off] = RA;
        if (cap && (t == cap_tensor || t == cap_tensor + 1)) cap[t - cap_tensor] = RA;
      } else {
        const int slot = (int)ra;
        const double* inj = A.inj[slot];
        double v;
        if (inj) {
          v = inj[env * ad->env_stride + off];
        } else {
          const uint32_t node = (uint32_t)A.variates[slot].node;
          if (op == OP_RANDN) {
            // Box-Muller pair of elements 2q, 2q + 1 (one Philox block): this element's branch
            double z0, z1;
            rddl_randn_pair(A.rk, (uint64_t)(A.env_offset + env), step, node, rddl_randn_counter((uint64_t)off), &z0, &z1);
            v = rddl_randn_branch((uint64_t)off) ? z1 : z0;
          } else {
            uint4 x = rddl_philox_draw(A.rk, (uint64_t)(A.env_offset + env), step, node, (uint64_t)off);
            const double u = rddl_u53(x.x, x.y);
            if (op == OP_RANDU) v = u;
            else if (op == OP_RANDE) v = -log1p(-u);
            else v = tan(3.141592653589793 * (u - 0.5));
          }
        }
        SETF(v);
      }
      break;
    }
    case OP_FADD: SETF(FA + FB); break;
    case OP_FSUB: SETF(FA - FB); break;
    case OP_FMUL: SETF(FA * FB); break;
    case OP_FDIV: SETF(FA / FB); break;
    case OP_FNEG: SETF(-FA); break;
    case OP_FABS: SETF(fabs(FA)); break;
    case OP_FMIN: SETF(np_fmin(FA, FB)); break;
    case OP_FMAX: SETF(np_fmax(FA, FB)); break;
    case OP_FPOW: SETF(pow(FA, FB)); break;
    case OP_FMOD: {  // np.mod (sign of the divisor), simulator.py:117-118
      const double a = FA, b = FB;
      double m = fmod(a, b);
      if (b != 0.0) {
        if (m != 0.0) { if ((b < 0) != (m < 0)) m += b; }
        else m = copysign(0.0, b);
      }
      SETF(m);
      break;
    }
    case OP_FLOGB: SETF(log(FA) / log(FB)); break;
    case OP_FHYPOT: SETF(hypot(FA, FB)); break;
    case OP_COS: SETF(cos(FA)); break;
    case OP_SIN: SETF(sin(FA)); break;
    case OP_TAN: SETF(tan(FA)); break;
    case OP_ACOS: SETF(acos(FA)); break;
    case OP_ASIN: SETF(asin(FA)); break;
    case OP_ATAN: SETF(atan(FA)); break;
    case OP_COSH: SETF(cosh(FA)); break;
    case OP_SINH: SETF(sinh(FA)); break;
    case OP_TANH: SETF(tanh(FA)); break;
    case OP_EXP: SETF(exp(FA)); break;
    case OP_LN: SETF(log(FA)); break;
    case OP_SQRT: SETF(sqrt(FA)); break;
    case OP_EXPM1: SETF(expm1(FA)); break;
    case OP_LOG1P: SETF(log1p(FA)); break;
    case OP_LNGAMMA: SETF(rddl_lngamma(FA)); break;
    case OP_GAMMA: SETF(exp(rddl_lngamma(FA))); break;
    case OP_FSGN: { const double a = FA; SETI((a > 0) - (a < 0)); break; }
    case OP_FROUND: SETI(__double2ll_rn(FA)); break;
    case OP_FFLOOR: SETI(__double2ll_rd(FA)); break;
    case OP_FCEIL: SETI(__double2ll_ru(FA)); break;
    case OP_IADD: SETI(IA + IB); break;
    case OP_ISUB: SETI(IA - IB); break;
    case OP_IMUL: SETI(IA * IB); break;
    case OP_INEG: SETI(-IA); break;
    case OP_IABS: SETI(llabs(IA)); break;
    case OP_IMIN: SETI(min((long long)IA, (long long)IB)); break;
    case OP_IMAX: SETI(max((long long)IA, (long long)IB)); break;
    case OP_IPOW: SETI(ipow64(IA, IB)); break;
    case OP_IFLOORDIV: {
      const int64_t a = IA, b = IB;
      int64_t q = 0;
      if (b != 0) { q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; }
      SETI(q);
      break;
    }
    case OP_IMOD: {
      const int64_t a = IA, b = IB;
      int64_t m = 0;
      if (b != 0) { m = a % b; if (m != 0 && ((m < 0) != (b < 0))) m += b; }
      SETI(m);
      break;
    }
    case OP_ISGN: { const int64_t a = IA; SETI((a > 0) - (a < 0)); break; }
    case OP_FLT: SETI(FA < FB); break;
    case OP_FLE: SETI(FA <= FB); break;
    case OP_FGT: SETI(FA > FB); break;
    case OP_FGE: SETI(FA >= FB); break;
    case OP_FEQ: SETI(FA == FB); break;
    case OP_FNE: SETI(FA != FB); break;
    case OP_ILT: SETI(IA < IB); break;
    case OP_ILE: SETI(IA <= IB); break;
    case OP_IGT: SETI(IA > IB); break;
    case OP_IGE: SETI(IA >= IB); break;
    case OP_IEQ: SETI(IA == IB); break;
    case OP_INE: SETI(IA != IB); break;
    case OP_AND: r[dst] = (RA & RB) & 1; break;
    case OP_OR: r[dst] = (RA | RB) & 1; break;
    case OP_XOR: r[dst] = (RA ^ RB) & 1; break;
    case OP_NOT: r[dst] = RA == 0; break;
    case OP_IMPLY: r[dst] = ((RA == 0) | RB) & 1; break;
    case OP_SEL: r[dst] = RA ? RB : r[rc]; break;
    case OP_LOOP:
      idx[ra] = 0;
      if (imm == 0) pc = (int)rb;
      break;
    case OP_ENDLOOP:
      if ((uint32_t)(++idx[ra]) < imm) pc = (int)rb;
      break;
    case OP_CSRLOOP:
    case OP_CSREND: {
      const RddlCsr* cs = P.csrs + imm;
      int32_t j;
      if (op == OP_CSRLOOP) {
        int64_t row = cs->row_c0;
        for (int i = 0; i < cs->row_naxis; ++i) row += cs->row_coef[i] * (int64_t)idx[cs->row_axis[i]];
        j = __ldg(A.pool + cs->rowptr_off + row);
        jend[ra] = __ldg(A.pool + cs->rowptr_off + row + 1);
        jcur[ra] = j;
        if (j >= jend[ra]) { pc = (int)rb; break; }
      } else {
        j = ++jcur[ra];
        if (j >= jend[ra]) break;
        pc = (int)rb;
      }
      for (int c = 0; c < cs->ncols; ++c) idx[cs->slot[c]] = __ldg(A.pool + cs->col_off[c] + j);
      break;
    }
    case OP_REDOUT: redval[rb] = red_combine(L.red_op[rb], redval[rb], RA); break;
    case OP_REDIN: {
      if (red_local) {
        // reductions of the producing launch (red_src) handed over in registers: slot imm -> its index there
#pragma unroll
        for (int j = 0; j < RDDL_MAX_RED; ++j)
          if (j < red_src->nred && (uint32_t)red_src->red_slot[j] == imm) r[dst] = red_local[j];
        break;
      }
      const RddlRedSlot s = P.redslots[imm];
      const uint64_t* base = A.red + (P.red_off[imm] * A.n_envs + env * s.chunks);
      uint64_t acc = __ldcg(base);
      for (int c = 1; c < s.chunks; ++c) acc = red_combine(s.op, acc, __ldcg(base + c));
      r[dst] = acc;
      break;
    }
    case OP_STATUS:
      if (RA) atomicOr(A.status + env, imm);
      break;
    case OP_SAMPLE: {
      // rejection / search samplers: one private Philox stream per (env, step, node, element)
      const RddlAddr* ad = P.addrs + imm;
      int64_t off = ad->c0;
      for (int i = 0; i < ad->naxis; ++i) off += ad->coef[i] * (int64_t)idx[ad->axis[i]];
      const uint32_t node = (uint32_t)A.variates[rc].node;
      r[dst] = rddl_sample((int)((raw.x >> 8) & 0xff), RA, RB, A.seed, (uint64_t)(A.env_offset + env), step,
                           node, (uint64_t)off);
      break;
    }
    case OP_SUMLOOP: {
      // sum over one loop variable of t1 [* t2], t = plain tensor reads: native MAC loop.
      // data word (next instruction): dst | a << 16 = extent or CSR id, imm = address of t2
      // (specialised programs pass the data word; the interpreter fetches it and skips over it)
      const uint4 raw2 = prog ? __ldg(reinterpret_cast<const uint4*>(prog + pc)) : raw2_static;
      if (prog) ++pc;
      const uint32_t arg = (raw2.x >> 16) | ((raw2.y & 0xffff) << 16);
      const bool is_csr = rb & 1u, two = rb & 2u, flt = rb & 4u;
      const int slot = (int)ra;
      int64_t base[2], stride[2];
      const void* ptr[2];
      int dts[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const RddlAddr* ad = P.addrs + (k == 0 ? imm : raw2.w);
        int64_t off = ad->c0 + env * ad->env_stride, st = 0;
        for (int i = 0; i < ad->naxis; ++i) {
          if (ad->axis[i] == slot) st = ad->coef[i];
          else off += ad->coef[i] * (int64_t)idx[ad->axis[i]];
        }
        base[k] = off; stride[k] = st;
        ptr[k] = A.tensors[ad->tensor];
        dts[k] = P.dtype[ad->tensor];
        if (stage && ad->tensor == stage->tensor && dts[k] == DT_REAL) {   // gather from the staged slice
          ptr[k] = stage->ptr + stage->add;
          base[k] = (off - stage->base) * stage->mul;
          stride[k] = st * stage->mul;
        }
      }
      int32_t j0 = 0, j1 = (int32_t)arg;
      const int32_t* cols = nullptr;
      int64_t sentinel = -2;                   // (filler index of a padded list; never a real index)
      bool u16 = false;                        // the padded list holds uint16 entries
      if (is_csr) {
        const RddlCsr* cs = P.csrs + arg;
        int64_t row = cs->row_c0;
        for (int i = 0; i < cs->row_naxis; ++i) row += cs->row_coef[i] * (int64_t)idx[cs->row_axis[i]];
        if (cs->rowptr4_off != 0) {
          // single-column lists are summed from their padded copy (rows of 4 k entries in the bank-scheduled order
          // of lowering._bank_schedule, filler = sentinel4): ONE order of the terms for every kernel and path
          j0 = __ldg(A.pool + cs->rowptr4_off + row);
          j1 = __ldg(A.pool + cs->rowptr4_off + row + 1);
          cols = A.pool + cs->col4_off;
          sentinel = cs->sentinel4;
          u16 = cs->col4_u16 != 0;
        } else {
          j0 = __ldg(A.pool + cs->rowptr_off + row);
          j1 = __ldg(A.pool + cs->rowptr_off + row + 1);
          cols = A.pool + cs->col_off[0];
        }
      }
      if (stage && is_csr && !two && flt && (P.addrs + imm)->tensor == stage->tensor && dts[0] == DT_REAL) {
        // sparse float sum of a per-env vector staged in shared memory (fused env-tile kernels): 32-bit
        // shared addresses, one LDG (column index, shared by the lanes of the element) + one LDS per
        // term, accumulation in index order like the general loop below
        const RddlCsr* cs4 = P.csrs + arg;
        const RddlAddr* ad4 = P.addrs + imm;
        if (stage->pad && cs4->rowptr4_off != 0 && ad4->naxis == 1 && ad4->coef[0] == 1 && ad4->c0 == 0 && ad4->nreg == 0 &&
            ad4->env_stride == cs4->sentinel4) {
          // the whole per-env vector is staged and followed by a zero: the padded copy of the list (rows of 4 k
          // entries on 16-byte boundaries, filler = the index of that zero) needs no single-index loops before
          // and after the 128-bit index loads -- they were a third of this loop's instructions on Reservoir
          // R = 1000 (rows of 21 +- 4 entries). Adding the trailing zeros leaves the sum unchanged (acc is never -0).
          int64_t row4 = cs4->row_c0;
          for (int i = 0; i < cs4->row_naxis; ++i) row4 += cs4->row_coef[i] * (int64_t)idx[cs4->row_axis[i]];
          const int2 q = make_int2(__ldg(A.pool + cs4->rowptr4_off + row4), __ldg(A.pool + cs4->rowptr4_off + row4 + 1));
          const int4* c4 = reinterpret_cast<const int4*>(A.pool + cs4->col4_off);
          const uint32_t s4 = (uint32_t)__cvta_generic_to_shared(stage->ptr) + 8u * (uint32_t)stage->add;
          const uint32_t k4 = 8u * (uint32_t)stage->mul;
          double acc4 = 0.0;
          // 16-byte groups: four int32 or eight uint16 entries
          const int sh = cs4->col4_u16 ? 3 : 2;
          int32_t g4 = q.x >> sh;
          const int32_t g4l = (q.y >> sh) - 1;         // last group of the row
#define RDDL_SUM4(X, Y, Z, W)                                                                \
  {                                                                                           \
    double v0, v1, v2, v3;                                                                    \
    asm("ld.shared.f64 %0, [%1];" : "=d"(v0) : "r"(s4 + k4 * (uint32_t)(X)));                \
    asm("ld.shared.f64 %0, [%1];" : "=d"(v1) : "r"(s4 + k4 * (uint32_t)(Y)));                \
    asm("ld.shared.f64 %0, [%1];" : "=d"(v2) : "r"(s4 + k4 * (uint32_t)(Z)));                \
    asm("ld.shared.f64 %0, [%1];" : "=d"(v3) : "r"(s4 + k4 * (uint32_t)(W)));                \
    acc4 = __dadd_rn(acc4, __dadd_rn(__dadd_rn(v0, v1), __dadd_rn(v2, v3)));                  \
  }
#define RDDL_SUMG(c)                                                                         \
  if (cs4->col4_u16) {                                                                        \
    RDDL_SUM4((c).x & 0xffffu, (uint32_t)(c).x >> 16, (c).y & 0xffffu, (uint32_t)(c).y >> 16) \
    RDDL_SUM4((c).z & 0xffffu, (uint32_t)(c).z >> 16, (c).w & 0xffffu, (uint32_t)(c).w >> 16) \
  } else {                                                                                    \
    RDDL_SUM4((c).x, (c).y, (c).z, (c).w)                                                     \
  }
          // one 16-byte group per half trip, the next group's indices in flight while this one is summed; two index
          // register sets used alternately (no copies), the look-ahead index clamped to the row's last group instead
          // of predicated (no default values to move). Measured on B200 and not kept (Reservoir R = 1000 x 4096 envs,
          // all within 124 +- 4 us): L1 prefetch of the thread's next row and next element, the warp's index
          // groups copied once into shared memory with coalesced loads -- the kernel is bound by the L1 data pipe
          // (ncu: l1tex__data_pipe_lsu_wavefronts 75 % of peak), and each of these only moves wavefronts around.
          if (g4 <= g4l) {
            int4 ca = __ldg(c4 + g4), cb;
#pragma unroll 1
            for (;;) {
              cb = __ldg(c4 + min(g4 + 1, g4l));
              RDDL_SUMG(ca)
              if (++g4 > g4l) break;
              ca = __ldg(c4 + min(g4 + 1, g4l));
              RDDL_SUMG(cb)
              if (++g4 > g4l) break;
            }
          }
#undef RDDL_SUMG
#undef RDDL_SUM4
          r[dst] = from_f(acc4);
          break;
        }
        const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(stage->ptr) + 8u * (uint32_t)(stage->add + (int32_t)base[0]);
        const int32_t sk = 8 * (int32_t)stride[0];
        // (lists without a padded copy, or a staged slice that is not the whole vector: single index loads, the
        // same groups of four from the start of the row as every other float SUMLOOP)
        double acc = 0.0;
        for (int32_t j = j0; j < j1; j += 4) {
          double v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            v[u] = 0.0;
            const int32_t cu = j + u < j1 ? rddl_csr_col(cols, j + u, u16) : (int32_t)sentinel;
            if (cu != (int32_t)sentinel) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v[u]) : "r"(s0 + (uint32_t)(sk * cu)));
          }
          acc = __dadd_rn(acc, __dadd_rn(__dadd_rn(v[0], v[1]), __dadd_rn(v[2], v[3])));
        }
        r[dst] = from_f(acc);
        break;
      }
      double facc = 0.0;
      int64_t iacc = 0;
      // four terms at a time: their (dependent) index and value loads are issued together. Float sums add each
      // group of four consecutive terms (from the start of the row; missing terms = +0) as (t0 + t1) + (t2 + t3) and
      // then to the running sum: ONE dependent addition per group on the loop-carried chain instead of four -- the
      // staged sums of the fused kernels were bound by that chain's latency (FP64 adds, ILP 1, 8 warps per
      // scheduler). Every float SUMLOOP path uses this order, so all kernels agree bit for bit; against the
      // reference (NumPy's own pairwise order) sums agree to rounding, as before.
      for (int32_t j = j0; j < j1; j += 4) {
        int64_t k[4];
        double fv[4][2];
        int64_t iv[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          k[u] = j + u < j1 ? (is_csr ? (int64_t)rddl_csr_col(cols, j + u, u16) : (int64_t)(j + u)) : -1;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (k[u] == sentinel) k[u] = -1;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          fv[u][0] = fv[u][1] = 1.0;
          iv[u][0] = iv[u][1] = 1;
          if (k[u] < 0) continue;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q == 1 && !two) break;
            const int64_t o = base[q] + stride[q] * k[u];
            if (dts[q] == DT_REAL) fv[u][q] = reinterpret_cast<const double*>(ptr[q])[o];
            else {
              iv[u][q] = dts[q] == DT_BOOL ? (int64_t)(reinterpret_cast<const uint8_t*>(ptr[q])[o] != 0)
                                           : reinterpret_cast<const int64_t*>(ptr[q])[o];
              fv[u][q] = (double)iv[u][q];
            }
          }
        }
        if (flt) {
          double t[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) t[u] = k[u] < 0 ? 0.0 : (two ? __dmul_rn(fv[u][0], fv[u][1]) : fv[u][0]);
          facc = __dadd_rn(facc, __dadd_rn(__dadd_rn(t[0], t[1]), __dadd_rn(t[2], t[3])));
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (k[u] < 0) continue;
            iacc += two ? iv[u][0] * iv[u][1] : iv[u][0];
          }
        }
      }
      r[dst] = flt ? from_f(facc) : (uint64_t)iacc;
      break;
    }
    default: break;
  }
}

// Interprets the program of launch L for one (env, element); idx[] holds the element's scope
// coordinates. Shared by the scalar kernel below and by the plane kernel's fused epilogue.
template <int NREG>
__device__ __noinline__ void vm_exec_element(const VmArgs& A, const RddlLaunch& L, const int64_t env,
                                             const uint64_t step, int32_t* idx, uint64_t* redval) {
  uint64_t r[NREG];
  int32_t jcur[4], jend[4];
  const int npc = L.pc_end - L.pc_begin;
  const RddlInsn* prog = A.insns + L.pc_begin;
  const VmPools P = {A.consts, A.addrs, A.redslots, A.red_off, A.csrs, A.dtype};
  int pc = 0;
  while (pc < npc) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(prog + pc));
    ++pc;
    vm_exec_one<NREG>(A, P, L, env, step, idx, redval, r, jcur, jend, prog, pc, raw, nullptr, 0);
  }
}

// One CTA of a scalar launch: a group of G consecutive threads works on one env, each thread visits
// `ipt` elements strided by G, then the per-env reductions are combined (shuffle within the group,
// shared memory across warps when G = 256). `exec(env, idx, redval)` runs the launch's program for
// one element: the bytecode interpreter (rddl_level_interp) or a specialised instruction sequence
// (rddl_level_static, rddl_level_static.cuh).
struct LevelNoPrologue {
  __device__ __forceinline__ void operator()(int64_t, bool) const {}
};

template <typename Exec, typename Pre = LevelNoPrologue>
__device__ __forceinline__ void level_tile(const VmArgs& A, const RddlLaunch& L, const int64_t* red_off, Exec exec,
                                           Pre prologue = Pre()) {
  const int G = L.G;
  const int tid = threadIdx.x;
  const int lane = tid & (G - 1);
  const int envs_per_cta = 256 / G;
  const int64_t chunk = blockIdx.x % L.chunks;
  const int64_t env = (int64_t)(blockIdx.x / L.chunks) * envs_per_cta + tid / G;
  const bool env_ok = env < A.n_envs;

  int32_t idx[RDDL_MAX_IDX];
  uint64_t redval[RDDL_MAX_RED];
#pragma unroll
  for (int j = 0; j < RDDL_MAX_RED; ++j) redval[j] = j < L.nred ? red_identity(L.red_op[j]) : 0;
  prologue(env, env_ok);      // (all threads of the CTA: may stage data in shared memory and synchronise)

  for (int it = 0; it < L.ipt; ++it) {
    const int64_t elem = chunk * ((int64_t)G * L.ipt) + (int64_t)it * G + lane;
    if (!env_ok || elem >= L.E) continue;
    {  // decode the element index into scope coordinates (C order)
      int64_t rem = elem;
#pragma unroll
      for (int k = RDDL_MAX_IDX - 1; k >= 0; --k) {
        if (k < L.rank) {
          int64_t e = L.extent[k];
          idx[k] = (int32_t)(rem % e);
          rem /= e;
        }
      }
    }
    exec(env, idx, redval);
  }

  // per-env reductions: shuffle within the group, shared memory across warps when G = 256
  if (L.nred > 0) {
    __shared__ uint64_t sm[RDDL_MAX_RED][8];
    for (int j = 0; j < L.nred; ++j) {
      const int op = L.red_op[j];
      uint64_t v = redval[j];
      const int w = G < 32 ? G : 32;
      for (int o = w >> 1; o > 0; o >>= 1) v = red_combine(op, v, __shfl_xor_sync(0xffffffffu, v, o, 32));
      if (G > 32) {
        if ((tid & 31) == 0) sm[j][tid >> 5] = v;
      } else if (lane == 0 && env_ok) {
        A.red[red_off[L.red_slot[j]] * A.n_envs + env * L.chunks + chunk] = v;
      }
    }
    if (G > 32) {
      __syncthreads();
      const int nw = G / 32;  // warps per group
      if (lane == 0 && env_ok) {
        const int w0 = (tid / G) * nw;
        for (int j = 0; j < L.nred; ++j) {
          const int op = L.red_op[j];
          uint64_t v = sm[j][w0];
          for (int w = 1; w < nw; ++w) v = red_combine(op, v, sm[j][w0 + w]);
          A.red[red_off[L.red_slot[j]] * A.n_envs + env * L.chunks + chunk] = v;
        }
      }
    }
  }
}
