// Fused level interpreter for batched RDDL stepping on sm_100a, and its C ABI.
//
// One thread evaluates the op-table program of one launch for one (env, scope element)
// pair; the launches of a plan are the fused CPF levels produced by
// pyrddlgym_b200/lowering.py. This replaces the recursive NumPy tree walk of
// RDDLSimulator._sample (pyRDDLGym/core/simulator.py:508-537) and everything it dispatches to
// (operators :584-952, distributions :958-1268).
//
// Thread mapping: a group of G consecutive threads works on one env (G = 1 for scope-free
// programs, 256 for large scopes), a CTA holds 256/G envs; large scopes are split into
// `chunks` CTAs per env, each thread visiting `ipt` elements strided by G so that
// consecutive threads touch consecutive elements (coalesced loads of state / action planes).
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rddl_b200.h"
#include "philox.cuh"
#include "rddl_plan.h"
#include "rddl_contract.cuh"
#include "rddl_samplers.cuh"

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
};

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

__device__ __forceinline__ double np_fmin(double a, double b) {
  return (isnan(a) || isnan(b)) ? NAN : fmin(a, b);
}
__device__ __forceinline__ double np_fmax(double a, double b) {
  return (isnan(a) || isnan(b)) ? NAN : fmax(a, b);
}

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

// Interprets the program of launch L for one (env, element); idx[] holds the element's scope
// coordinates. Shared by the scalar kernel below and by the plane kernel's fused epilogue.
template <int NREG>
__device__ __noinline__ void vm_exec_element(const VmArgs& A, const RddlLaunch& L, const int64_t env,
                                             const uint64_t step, int32_t* idx, uint64_t* redval) {
  uint64_t r[NREG];
  int32_t jcur[4], jend[4];
  const int npc = L.pc_end - L.pc_begin;
  const RddlInsn* prog = A.insns + L.pc_begin;
  {
    int pc = 0;
    while (pc < npc) {
      const uint4 raw = __ldg(reinterpret_cast<const uint4*>(prog + pc));
      const uint32_t op = raw.x & 0xff;
      const uint32_t dst = raw.x >> 16;
      const uint32_t ra = raw.y & 0xffff, rb = raw.y >> 16;
      const uint32_t rc = raw.z & 0xffff;
      const uint32_t imm = raw.w;
      ++pc;
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
        case OP_CONST: r[dst] = __ldg(A.consts + imm); break;
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
          const RddlAddr* ad = A.addrs + imm;
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
            off += env * ad->env_stride;
            const void* base = A.tensors[t];
            const int dt = A.dtype[t];
            if (dt == DT_BOOL) r[dst] = reinterpret_cast<const uint8_t*>(base)[off] != 0;
            else if (dt == DT_PACKED) r[dst] = (reinterpret_cast<const uint32_t*>(base)[off >> 5] >> (off & 31)) & 1u;
            else r[dst] = reinterpret_cast<const uint64_t*>(base)[off];
          } else if (op == OP_STORE) {
            off += env * ad->env_stride;
            void* base = A.tensors[t];
            const int dt = A.dtype[t];
            if (dt == DT_BOOL) reinterpret_cast<uint8_t*>(base)[off] = RA != 0;
            else reinterpret_cast<uint64_t*>(base)[off] = RA;
          } else {
            const int slot = (int)ra;
            const double* inj = A.inj[slot];
            double v;
            if (inj) {
              v = inj[env * ad->env_stride + off];
            } else {
              const uint32_t node = (uint32_t)A.variates[slot].node;
              uint4 x = rddl_philox_draw(A.seed, (uint64_t)(A.env_offset + env), step, node, (uint64_t)off);
              const double u = rddl_u53(x.x, x.y);
              if (op == OP_RANDU) v = u;
              else if (op == OP_RANDN) v = sqrt(-2.0 * log(1.0 - u)) * cos(2.0 * 3.141592653589793 * rddl_u53(x.z, x.w));
              else if (op == OP_RANDE) v = -log1p(-u);
              else v = tan(3.141592653589793 * (u - 0.5));
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
          const RddlCsr* cs = A.csrs + imm;
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
          const RddlRedSlot s = A.redslots[imm];
          const uint64_t* base = A.red + (A.red_off[imm] * A.n_envs + env * s.chunks);
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
          const RddlAddr* ad = A.addrs + imm;
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
          const uint4 raw2 = __ldg(reinterpret_cast<const uint4*>(prog + pc));
          ++pc;
          const uint32_t arg = (raw2.x >> 16) | ((raw2.y & 0xffff) << 16);
          const bool is_csr = rb & 1u, two = rb & 2u, flt = rb & 4u;
          const int slot = (int)ra;
          int64_t base[2], stride[2];
          const void* ptr[2];
          int dts[2];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const RddlAddr* ad = A.addrs + (k == 0 ? imm : raw2.w);
            int64_t off = ad->c0 + env * ad->env_stride, st = 0;
            for (int i = 0; i < ad->naxis; ++i) {
              if (ad->axis[i] == slot) st = ad->coef[i];
              else off += ad->coef[i] * (int64_t)idx[ad->axis[i]];
            }
            base[k] = off; stride[k] = st;
            ptr[k] = A.tensors[ad->tensor];
            dts[k] = A.dtype[ad->tensor];
          }
          int32_t j0 = 0, j1 = (int32_t)arg;
          const int32_t* cols = nullptr;
          if (is_csr) {
            const RddlCsr* cs = A.csrs + arg;
            int64_t row = cs->row_c0;
            for (int i = 0; i < cs->row_naxis; ++i) row += cs->row_coef[i] * (int64_t)idx[cs->row_axis[i]];
            j0 = __ldg(A.pool + cs->rowptr_off + row);
            j1 = __ldg(A.pool + cs->rowptr_off + row + 1);
            cols = A.pool + cs->col_off[0];
          }
          double facc = 0.0;
          int64_t iacc = 0;
          for (int32_t j = j0; j < j1; ++j) {
            const int64_t k = is_csr ? (int64_t)__ldg(cols + j) : (int64_t)j;
            double fv[2] = {1.0, 1.0};
            int64_t iv[2] = {1, 1};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              if (q == 1 && !two) break;
              const int64_t o = base[q] + stride[q] * k;
              if (dts[q] == DT_REAL) fv[q] = reinterpret_cast<const double*>(ptr[q])[o];
              else {
                iv[q] = dts[q] == DT_BOOL ? (int64_t)(reinterpret_cast<const uint8_t*>(ptr[q])[o] != 0)
                                          : reinterpret_cast<const int64_t*>(ptr[q])[o];
                fv[q] = (double)iv[q];
              }
            }
            if (flt) facc = __dadd_rn(facc, two ? __dmul_rn(fv[0], fv[1]) : fv[0]);
            else iacc += two ? iv[0] * iv[1] : iv[0];
          }
          r[dst] = flt ? from_f(facc) : (uint64_t)iacc;
          break;
        }
        default: break;
      }
    }
  }
}

#include "rddl_plane.cuh"

template <int NREG>
__global__ void __launch_bounds__(256) rddl_level_interp(const __grid_constant__ VmArgs A) {
  const RddlLaunch& L = A.L;
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

  for (int it = 0; it < L.ipt; ++it) {
    const int64_t elem = chunk * ((int64_t)G * L.ipt) + (int64_t)it * G + lane;
    if (!env_ok || elem >= L.E) continue;
    {  // decode the element index into scope coordinates (C order)
      int64_t rem = elem;
      for (int k = L.rank - 1; k >= 0; --k) {
        int64_t e = L.extent[k];
        idx[k] = (int32_t)(rem % e);
        rem /= e;
      }
    }
    vm_exec_element<NREG>(A, L, env, A.step, idx, redval);
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
        A.red[A.red_off[L.red_slot[j]] * A.n_envs + env * L.chunks + chunk] = v;
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
          A.red[A.red_off[L.red_slot[j]] * A.n_envs + env * L.chunks + chunk] = v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// host side: op-table parsing and the C ABI
// ---------------------------------------------------------------------------------------

static thread_local std::string g_create_error;

struct rddl_program {
  int device = 0;
  std::string err;
  std::vector<RddlTensor> tensors;
  std::vector<RddlLaunch> launches;
  std::vector<RddlRedSlot> redslots;
  std::vector<RddlVariate> variates;
  std::vector<int64_t> red_off;
  int64_t red_total = 0;  // sum of chunks over slots
  int64_t n_insns = 0;
  int64_t launches_issued = 0;
  // device copies
  RddlInsn* d_insns = nullptr;
  uint64_t* d_consts = nullptr;
  RddlAddr* d_addrs = nullptr;
  RddlCsr* d_csrs = nullptr;
  int32_t* d_pool = nullptr;
  RddlVariate* d_variates = nullptr;
  RddlRedSlot* d_redslots = nullptr;
  int64_t* d_red_off = nullptr;
  uint64_t* d_red = nullptr;
  int64_t red_capacity_envs = 0;
  // optional per-launch timing (rddl_profile): CUDA events on the launch stream
  bool profiling = false;
  struct Timed { int launch; cudaEvent_t e0, e1; };
  std::vector<Timed> timed;
  std::vector<PlaneProg> plane_progs;   // per launch index; pre-decoded programs of plane launches
  uint8_t* d_alive = nullptr;           // rollout (general path): env still running
  int64_t alive_capacity = 0;
};

// Pre-decodes the instructions of a plane launch (and the Fin descriptors they reference in the
// const pool: [m, regs, truth, bad, always, thr[2^m], vals[2^m], lev[levels]]) into a PlaneProg.
static int log2_or_neg(int64_t v) {
  if (v <= 0 || (v & (v - 1))) return -1;
  int s = 0;
  while ((1ll << s) < v) ++s;
  return s;
}

static bool decode_plane(const RddlLaunch& L, const RddlInsn* insns, const uint64_t* consts, int64_t n_consts,
                         PlaneProg* G, std::string* err) {
  memset(G, 0, sizeof *G);
  G->w32_shift = log2_or_neg((L.rank == 2 ? L.extent[1] : L.extent[0]) >> 5);
  G->wpe_shift = log2_or_neg(L.E >> 5);
  int nw = 0;
  auto need = [&](int n) { return nw + n <= PLANE_WORDS; };
  for (int pc = L.pc_begin; pc < L.pc_end; ++pc) {
    const RddlInsn& in = insns[pc];
    if (G->nops >= PLANE_MAX_OPS) { *err = "plane program too long"; return false; }
    PlaneOp& o = G->ops[G->nops++];
    o.op = in.op; o.dst = in.dst; o.a = in.a; o.b = in.b; o.imm = in.imm;
    o.off = nw;
    if (in.op == OP_PSHARE && (int)in.b + 1 > G->nshared) G->nshared = in.b + 1;
    if (in.op != OP_PLUT && in.op != OP_PBERN && in.op != OP_PRED) continue;
    if ((int64_t)in.imm + 5 > n_consts) { *err = "plane op: bad Fin index"; return false; }
    const uint64_t* c = consts + in.imm;
    const int m = (int)(c[0] & 0xff);
    const bool hi_const = ((c[0] >> 8) & 1) != 0;
    const int64_t fin_words = 5 + (in.op == OP_PLUT ? 0 : 2 * (1 << (m & 7))) + (in.op == OP_PBERN ? PLANE_BERN_LEVELS : 0);
    if (m < 0 || m > 5 || (int64_t)in.imm + fin_words > n_consts) { *err = "plane op: bad Fin"; return false; }
    const int TT = 1 << m;
    o.m = m;
    o.idx = 0;
    for (int i = 0; i < m; ++i) {
      const uint32_t r = (uint32_t)((c[1] >> (8 * i)) & 0xff);
      if (r >= PLANE_MAX_REGS) { *err = "plane op: register out of range"; return false; }
      o.idx |= r << (6 * i);
    }
    const uint64_t* thr = c + 5;
    const uint64_t* vals = c + 5 + TT;
    if (in.op == OP_PLUT) {
      if (m < 1 || !need(TT)) { *err = "plane tables too large"; return false; }
      for (int v = 0; v < TT; ++v) G->words[nw++] = ((c[2] >> v) & 1) ? 0xffffffffu : 0u;
    } else if (in.op == OP_PBERN) {
      if (m > 4 || !need((2 + PLANE_BERN_LEVELS) * TT + 4 * TT)) { *err = "plane tables too large"; return false; }
      const uint64_t* lev = c + 5 + 2 * TT;
      o.b = (c[3] != 0 ? 1 : 0) | (c[4] != 0 ? 2 : 0);
      if (hi_const && m >= 2) o.m |= 0x100;
      for (int v = 0; v < TT; ++v) G->words[nw++] = ((c[3] >> v) & 1) ? 0xffffffffu : 0u;
      for (int v = 0; v < TT; ++v) G->words[nw++] = ((c[4] >> v) & 1) ? 0xffffffffu : 0u;
      for (int j = 0; j < PLANE_BERN_LEVELS; ++j)
        for (int v = 0; v < TT; ++v) G->words[nw++] = ((lev[j] >> v) & 1) ? 0xffffffffu : 0u;
      for (int v = 0; v < TT; ++v) { G->words[nw++] = (uint32_t)thr[v]; G->words[nw++] = (uint32_t)(thr[v] >> 32); }
      for (int v = 0; v < TT; ++v) { G->words[nw++] = (uint32_t)vals[v]; G->words[nw++] = (uint32_t)(vals[v] >> 32); }
    } else {
      if (m > 1 || !need(4)) { *err = "plane reduction over more than one plane"; return false; }
      G->words[nw++] = (uint32_t)vals[0]; G->words[nw++] = (uint32_t)(vals[0] >> 32);
      const uint64_t v1 = m ? vals[1] : 0;
      G->words[nw++] = (uint32_t)v1; G->words[nw++] = (uint32_t)(v1 >> 32);
      if (in.b < RDDL_MAX_RED) { G->red_off[in.b] = o.off; G->red_m[in.b] = m; }
    }
  }
  return true;
}

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t e_ = (call);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      char buf_[512];                                                                 \
      snprintf(buf_, sizeof buf_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      *errp = buf_;                                                                   \
      return RDDL_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)

template <typename T>
static int upload(const void* src, int64_t n, T** dst, std::string* errp) {
  *dst = nullptr;
  CK(cudaMalloc((void**)dst, (size_t)(n > 0 ? n : 1) * sizeof(T)));
  if (n > 0) CK(cudaMemcpy(*dst, src, (size_t)n * sizeof(T), cudaMemcpyHostToDevice));
  return RDDL_OK;
}

extern "C" int rddl_program_create(const void* optable, size_t nbytes, int device, rddl_program_t** out) {
  std::string* errp = &g_create_error;
  if (!optable || !out || nbytes < 16 * sizeof(int64_t)) { *errp = "bad arguments"; return RDDL_ERR_ARG; }
  const uint8_t* p = (const uint8_t*)optable;
  int64_t hdr[16];
  memcpy(hdr, p, sizeof hdr);
  if (hdr[0] != RDDL_MAGIC || hdr[1] != RDDL_VERSION) { *errp = "op table: bad magic/version"; return RDDL_ERR_FORMAT; }
  const size_t sizes[9] = {sizeof(RddlTensor), sizeof(RddlLaunch), sizeof(RddlInsn), 8, sizeof(RddlAddr),
                           sizeof(RddlCsr), 4, sizeof(RddlRedSlot), sizeof(RddlVariate)};
  const uint8_t* sec[9];
  size_t off = sizeof hdr;
  for (int i = 0; i < 9; ++i) {
    sec[i] = p + off;
    size_t b = (size_t)hdr[2 + i] * sizes[i];
    off += (b + 7) & ~(size_t)7;
  }
  if (off > nbytes) { *errp = "op table: truncated"; return RDDL_ERR_FORMAT; }
  if (hdr[2] > RDDL_MAX_TENSORS) { *errp = "op table: too many tensors"; return RDDL_ERR_FORMAT; }
  if (hdr[10] > RDDL_MAX_VARIATES) { *errp = "op table: too many random nodes"; return RDDL_ERR_FORMAT; }
  rddl_program* g = new rddl_program();
  errp = &g->err;
  g->device = device;
  g->tensors.assign((const RddlTensor*)sec[0], (const RddlTensor*)sec[0] + hdr[2]);
  g->launches.assign((const RddlLaunch*)sec[1], (const RddlLaunch*)sec[1] + hdr[3]);
  g->redslots.assign((const RddlRedSlot*)sec[7], (const RddlRedSlot*)sec[7] + hdr[9]);
  g->variates.assign((const RddlVariate*)sec[8], (const RddlVariate*)sec[8] + hdr[10]);
  g->n_insns = hdr[4];
  g->plane_progs.resize(g->launches.size());
  const int reward_tensor = (int)g->tensors.size() - 3;   // specials: __reward, __done, __chk
  for (size_t i = 0; i < g->launches.size(); ++i) {
    if (g->launches[i].kind != 1) continue;
    PlaneProg* G = &g->plane_progs[i];
    if (!decode_plane(g->launches[i], (const RddlInsn*)sec[2], (const uint64_t*)sec[3], hdr[5], G,
                      &g_create_error)) {
      delete g;
      return RDDL_ERR_FORMAT;
    }
    G->pad = reward_tensor;
    // scalar programs fused into this launch's epilogue
    for (size_t k = i + 1; k < g->launches.size() && g->launches[k].kind == 2; ++k) {
      if (G->nfused >= 2 || g->launches[k].plan != g->launches[i].plan || g->launches[k].nregs > 40) {
        g_create_error = "op table: too many fused scalar launches";
        delete g;
        return RDDL_ERR_FORMAT;
      }
      G->fused[G->nfused++] = g->launches[k];
    }
    // state carry for multi-step rollouts: PLD of a state fluent whose primed partner is stored
    // by this launch takes last step's stored register instead of reloading from HBM
    bool ok = g->launches[i].chunks == 1 && G->nfused > 0;
    for (int a = 0; a < G->nops; ++a) {
      PlaneOp& o = G->ops[a];
      if (o.op != OP_PLD) continue;
      const int pair = g->tensors[o.imm].pair;
      if (pair < 0) continue;
      int src = -1;
      for (int b = 0; b < G->nops; ++b)
        if (G->ops[b].op == OP_PST && (int)G->ops[b].imm == pair) { src = G->ops[b].a; G->ops[b].b = 1; }
      if (src < 0) ok = false;
      else o.a = src + 1;
    }
    G->rollout_ok = ok;
  }
  for (auto& L : g->launches) {
    const bool bad_scalar = L.kind != 1 && L.kind != 3 && (L.nregs > RDDL_MAX_REGS || L.G > 256 || (L.G & (L.G - 1)));
    const bool bad_plane = L.kind == 1 && (L.nregs > PLANE_MAX_REGS || L.G > 32 || L.ipt > 1024 || (L.E & 31));
    if (bad_scalar || bad_plane || L.kind > 3 || L.G < 1 || L.nred > RDDL_MAX_RED) {
      g_create_error = "op table: launch out of limits";
      delete g;
      return RDDL_ERR_FORMAT;
    }
  }
  // the plane kernels keep their virtual register file in dynamic shared memory (> 48 KB)
  {
    const int max_smem = 200 * 1024;
    const void* fns[3] = {(const void*)rddl_plane_interp<1, 256>, (const void*)rddl_plane_interp<2, 256>,
                          (const void*)rddl_plane_interp<4, 256>};
    for (const void* f : fns) {
      cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
      if (getenv("RDDL_PLANE_CARVEOUT"))   // experiment knob: shared-memory share of the unified L1 array, percent
        cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(getenv("RDDL_PLANE_CARVEOUT")));
    }
    cudaGetLastError();   // no device in CPU-only builds: ignored here, create fails at cudaSetDevice below
  }
  int64_t acc = 0;
  for (auto& s : g->redslots) { g->red_off.push_back(acc); acc += s.chunks; }
  g->red_total = acc;
  int rc = RDDL_OK;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e); delete g; return RDDL_ERR_CUDA; }
  if ((rc = upload<RddlInsn>(sec[2], hdr[4], &g->d_insns, errp)) ||
      (rc = upload<uint64_t>(sec[3], hdr[5], &g->d_consts, errp)) ||
      (rc = upload<RddlAddr>(sec[4], hdr[6], &g->d_addrs, errp)) ||
      (rc = upload<RddlCsr>(sec[5], hdr[7], &g->d_csrs, errp)) ||
      (rc = upload<int32_t>(sec[6], hdr[8], &g->d_pool, errp)) ||
      (rc = upload<RddlVariate>(sec[8], hdr[10], &g->d_variates, errp)) ||
      (rc = upload<RddlRedSlot>(sec[7], hdr[9], &g->d_redslots, errp)) ||
      (rc = upload<int64_t>(g->red_off.data(), (int64_t)g->red_off.size(), &g->d_red_off, errp))) {
    g_create_error = g->err;
    rddl_program_destroy(g);
    return rc;
  }
  *out = g;
  return RDDL_OK;
}

extern "C" void rddl_program_destroy(rddl_program_t* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaFree(g->d_insns); cudaFree(g->d_consts); cudaFree(g->d_addrs); cudaFree(g->d_csrs);
  cudaFree(g->d_pool); cudaFree(g->d_variates); cudaFree(g->d_redslots); cudaFree(g->d_red_off);
  cudaFree(g->d_red);
  cudaFree(g->d_alive);
  delete g;
}

extern "C" const char* rddl_last_error(const rddl_program_t* g) {
  return g ? g->err.c_str() : g_create_error.c_str();
}

extern "C" int64_t rddl_program_info(const rddl_program_t* g, int what) {
  if (!g) return -1;
  switch (what) {
    case 0: return (int64_t)g->tensors.size();
    case 1: return (int64_t)g->variates.size();
    case 2: { int64_t n = 0; for (auto& L : g->launches) n += L.plan == RDDL_PLAN_STEP; return n; }
    case 3: return g->n_insns;
    case 4: return g->launches_issued;
    default: return -1;
  }
}

struct RolloutArgs {
  int n_steps;
  double gamma;
  double* returns;
  double* rewards;
  const int64_t* stride_bytes;
};

static int run_plan(rddl_program* g, int plan, void* const* tensors, const double* const* injected,
                    uint32_t* status, int64_t n_envs, int64_t env_offset, uint64_t seed, uint64_t step,
                    cudaStream_t stream, const RolloutArgs* ro = nullptr) {
  std::string* errp = &g->err;
  if (!tensors || n_envs <= 0 || !status) { g->err = "bad arguments"; return RDDL_ERR_ARG; }
  CK(cudaSetDevice(g->device));
  if (g->red_total > 0 && g->red_capacity_envs < n_envs) {
    // grow the per-env reduction scratch (not capturable: size the first call like the largest)
    CK(cudaStreamSynchronize(stream));
    cudaFree(g->d_red);
    g->d_red = nullptr;
    CK(cudaMalloc((void**)&g->d_red, (size_t)(g->red_total * n_envs) * sizeof(uint64_t)));
    g->red_capacity_envs = n_envs;
  }
  VmArgs A;
  memset(&A, 0, sizeof A);
  A.insns = g->d_insns; A.consts = g->d_consts; A.addrs = g->d_addrs; A.csrs = g->d_csrs;
  A.pool = g->d_pool; A.variates = g->d_variates; A.redslots = g->d_redslots; A.red_off = g->d_red_off;
  A.red = g->d_red; A.status = status; A.n_envs = n_envs; A.env_offset = env_offset;
  A.seed = seed; A.step = step;
  A.n_steps = 1;
  if (ro) {
    A.n_steps = ro->n_steps; A.gamma = ro->gamma; A.returns = ro->returns; A.rewards = ro->rewards;
    if (ro->stride_bytes) for (size_t i = 0; i < g->tensors.size(); ++i) A.tstride[i] = ro->stride_bytes[i];
  }
  for (size_t i = 0; i < g->tensors.size(); ++i) { A.tensors[i] = tensors[i]; A.dtype[i] = (uint8_t)g->tensors[i].dtype; }
  if (injected) for (size_t i = 0; i < g->variates.size(); ++i) A.inj[i] = injected[i];
  for (const RddlLaunch& L : g->launches) {
    if (L.plan != plan || L.kind == 2) continue;   // kind 2 runs inside the preceding plane launch
    A.L = L;
    // words per thread: the tile geometry of the op table (4 words, 32768 cells per CTA: big tiles
    // amortise the per-thread interpretation overhead); for very small batches of small grids
    // one word per thread, so that the grid still spans the SMs. RDDL_PLANE_FORCE_WPT overrides.
    int wpt = L.ipt <= 256 ? 1 : (L.ipt <= 512 ? 2 : 4);
    const int block = 256;
    if (L.kind == 1 && L.chunks == 1 && wpt > 1) {
      const char* force = getenv("RDDL_PLANE_FORCE_WPT");
      int want = force ? atoi(force) : ((n_envs + L.G - 1) / L.G < 74 ? 1 : wpt);
      if (want != 1 && want != 2 && want != 4) want = wpt;
      if (want < wpt && L.E <= 8192 * want) {
        int ept = (int)((8192 * want) / L.E);
        if (ept > 32) ept = 32;
        wpt = want;
        A.L.G = ept;
        A.L.ipt = (int)(ept * (L.E >> 5));
      }
    }
    const int envs_per_cta = L.kind == 1 ? A.L.G : 256 / L.G;
    const int64_t groups = (n_envs + envs_per_cta - 1) / envs_per_cta;
    const int64_t blocks = groups * L.chunks;
    if (blocks > 0x7fffffffll) { g->err = "grid too large"; return RDDL_ERR_ARG; }
    rddl_program::Timed tm;
    if (g->profiling) {
      tm.launch = (int)(&L - g->launches.data());
      CK(cudaEventCreate(&tm.e0));
      CK(cudaEventCreate(&tm.e1));
      CK(cudaEventRecord(tm.e0, stream));
    }
    if (L.kind == 3) {
      const int N = (int)L.extent[0], K = (int)L.extent[1];
      dim3 grid((unsigned)((N + CT_BN - 1) / CT_BN), (unsigned)((n_envs + CT_BM - 1) / CT_BM));
      rddl_contract_f64<<<grid, 128, 0, stream>>>((const double*)tensors[L.red_slot[1]],
                                                  (const double*)tensors[L.red_slot[0]],
                                                  (double*)tensors[L.red_slot[2]], n_envs, N, K,
                                                  (int64_t)L.red_op[0], (int64_t)L.red_op[1]);
    } else if (L.kind == 1) {
      const PlaneProg& G = g->plane_progs[(size_t)(&L - g->launches.data())];
      const int W32 = (int)((L.rank == 2 ? L.extent[1] : L.extent[0]) >> 5);
      const size_t smem = sizeof(uint32_t) * ((size_t)L.nregs * wpt * block +
                                              (size_t)G.nshared * ((A.L.ipt / W32 + A.L.G + 2) * (W32 + 2)) +
                                              32 * RDDL_MAX_RED +
                                              8 * PLANE_QUEUE + 8 + block * wpt);
      if (wpt == 1) rddl_plane_interp<1, 256><<<(unsigned)blocks, 256, smem, stream>>>(A, G);
      else if (wpt == 2) rddl_plane_interp<2, 256><<<(unsigned)blocks, 256, smem, stream>>>(A, G);
      else rddl_plane_interp<4, 256><<<(unsigned)blocks, 256, smem, stream>>>(A, G);
    } else if (L.nregs <= 16) rddl_level_interp<16><<<(unsigned)blocks, 256, 0, stream>>>(A);
    else if (L.nregs <= 40) rddl_level_interp<40><<<(unsigned)blocks, 256, 0, stream>>>(A);
    else rddl_level_interp<RDDL_MAX_REGS><<<(unsigned)blocks, 256, 0, stream>>>(A);
    ++g->launches_issued;
    CK(cudaGetLastError());
    if (g->profiling) {
      CK(cudaEventRecord(tm.e1, stream));
      g->timed.push_back(tm);
    }
  }
  return RDDL_OK;
}

extern "C" int rddl_step(rddl_program_t* g, void* const* tensors, const double* const* injected,
                         uint32_t* status, int64_t n_envs, int64_t env_offset, uint64_t seed,
                         uint64_t step, void* cuda_stream) {
  if (!g) return RDDL_ERR_ARG;
  return run_plan(g, RDDL_PLAN_STEP, tensors, injected, status, n_envs, env_offset, seed, step,
                  (cudaStream_t)cuda_stream);
}

__global__ void rddl_accumulate_returns(double* returns, double* rewards_t, const double* reward,
                                        const uint8_t* done, uint8_t* alive, double disc, int64_t n, int first) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool live = first ? true : alive[i] != 0;
  const double r = reward[i];
  if (rewards_t) rewards_t[i] = r;
  if (live) returns[i] += r * disc;
  alive[i] = live && !done[i];
}

extern "C" int rddl_rollout(rddl_program_t* g, void* const* tensors, uint32_t* status, int64_t n_envs,
                            int64_t env_offset, uint64_t seed, uint64_t step0, int32_t n_steps,
                            const int64_t* step_stride_bytes, double gamma, double* returns, double* rewards,
                            int32_t* state_in_next, void* cuda_stream) {
  if (!g) return RDDL_ERR_ARG;
  if (n_steps < 1 || !returns || !state_in_next) { g->err = "rddl_rollout: bad arguments"; return RDDL_ERR_ARG; }
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  std::string* errp = &g->err;
  // fused path: the step plan is one single-tile plane launch (+ its fused scalar epilogue)
  int n_step_launches = 0;
  const PlaneProg* G = nullptr;
  for (size_t i = 0; i < g->launches.size(); ++i) {
    if (g->launches[i].plan != RDDL_PLAN_STEP || g->launches[i].kind == 2) continue;
    ++n_step_launches;
    if (g->launches[i].kind == 1) G = &g->plane_progs[i];
  }
  if (n_step_launches == 1 && G && G->rollout_ok) {
    RolloutArgs ro{n_steps, gamma, returns, rewards, step_stride_bytes};
    *state_in_next = 1;
    return run_plan(g, RDDL_PLAN_STEP, tensors, nullptr, status, n_envs, env_offset, seed, step0, stream, &ro);
  }
  // general path: one step plan per step, state / next-state pointers swapped on the host
  CK(cudaSetDevice(g->device));
  if (g->alive_capacity < n_envs) {
    CK(cudaStreamSynchronize(stream));
    cudaFree(g->d_alive);
    g->d_alive = nullptr;
    CK(cudaMalloc((void**)&g->d_alive, (size_t)n_envs));
    g->alive_capacity = n_envs;
  }
  const size_t nt = g->tensors.size();
  std::vector<void*> cur(tensors, tensors + nt);
  const int reward_tensor = (int)nt - 3;
  double disc = 1.0;
  for (int t = 0; t < n_steps; ++t) {
    std::vector<void*> at(cur);
    if (step_stride_bytes)
      for (size_t i = 0; i < nt; ++i) at[i] = (uint8_t*)cur[i] + (int64_t)t * step_stride_bytes[i];
    int rc = run_plan(g, RDDL_PLAN_STEP, at.data(), nullptr, status, n_envs, env_offset, seed, step0 + t, stream);
    if (rc) return rc;
    rddl_accumulate_returns<<<(unsigned)((n_envs + 255) / 256), 256, 0, stream>>>(
        returns, rewards ? rewards + (int64_t)t * n_envs : nullptr, (const double*)cur[reward_tensor],
        (const uint8_t*)cur[reward_tensor + 1], g->d_alive, disc, n_envs, t == 0);
    CK(cudaGetLastError());
    ++g->launches_issued;
    disc *= gamma;
    for (size_t i = 0; i < nt; ++i) {
      const int pair = g->tensors[i].pair;
      if (pair >= 0) std::swap(cur[i], cur[pair]);
    }
  }
  *state_in_next = n_steps & 1;
  return RDDL_OK;
}

extern "C" int rddl_check(rddl_program_t* g, int which, void* const* tensors, uint32_t* status,
                          int64_t n_envs, int64_t env_offset, uint64_t seed, uint64_t step, void* cuda_stream) {
  if (!g) return RDDL_ERR_ARG;
  if (which != RDDL_PLAN_TERMINAL && which != RDDL_PLAN_INVARIANT && which != RDDL_PLAN_PRECONDITION) {
    g->err = "rddl_check: bad plan";
    return RDDL_ERR_ARG;
  }
  return run_plan(g, which, tensors, nullptr, status, n_envs, env_offset, seed, step, (cudaStream_t)cuda_stream);
}

extern "C" int rddl_profile(rddl_program_t* g, int enable) {
  if (!g) return RDDL_ERR_ARG;
  g->profiling = enable != 0;
  return RDDL_OK;
}

extern "C" int rddl_profile_read(rddl_program_t* g, double* ms, int64_t* count, int n) {
  if (!g || !ms || !count) return RDDL_ERR_ARG;
  std::string* errp = &g->err;
  for (int i = 0; i < n; ++i) { ms[i] = 0.0; count[i] = 0; }
  for (auto& t : g->timed) {
    CK(cudaEventSynchronize(t.e1));
    float f = 0.f;
    CK(cudaEventElapsedTime(&f, t.e0, t.e1));
    if (t.launch < n) { ms[t.launch] += f; count[t.launch] += 1; }
    cudaEventDestroy(t.e0);
    cudaEventDestroy(t.e1);
  }
  g->timed.clear();
  return RDDL_OK;
}
