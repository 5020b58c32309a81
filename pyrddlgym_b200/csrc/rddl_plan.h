// Op-table structs shared by host (C-ABI) and device (level interpreter).
// Field-for-field mirror of pyrddlgym_b200/optable.py -- keep the two in sync.
#pragma once
#include "rddl_stdint.h"

#define RDDL_MAGIC 0x31304C444452ll
#define RDDL_VERSION 11
#define RDDL_MAX_IDX 8
#define RDDL_MAX_RED 8
#define RDDL_MAX_REGS 96
#define RDDL_MAX_TENSORS 120
#define RDDL_MAX_VARIATES 64

enum RddlOp : uint8_t {
  OP_NOP, OP_CONST, OP_AXIS, OP_LOAD, OP_STORE, OP_MOV, OP_I2F, OP_F2I,
  OP_FADD, OP_FSUB, OP_FMUL, OP_FDIV, OP_FNEG, OP_FABS, OP_FMIN, OP_FMAX, OP_FPOW, OP_FMOD, OP_FLOGB,
  OP_FHYPOT, OP_COS, OP_SIN, OP_TAN, OP_ACOS, OP_ASIN, OP_ATAN, OP_COSH, OP_SINH, OP_TANH, OP_EXP, OP_LN,
  OP_SQRT, OP_LNGAMMA, OP_GAMMA, OP_FSGN, OP_FROUND, OP_FFLOOR, OP_FCEIL,
  OP_IADD, OP_ISUB, OP_IMUL, OP_INEG, OP_IABS, OP_IMIN, OP_IMAX, OP_IPOW, OP_IFLOORDIV, OP_IMOD, OP_ISGN,
  OP_FLT, OP_FLE, OP_FGT, OP_FGE, OP_FEQ, OP_FNE, OP_ILT, OP_ILE, OP_IGT, OP_IGE, OP_IEQ, OP_INE,
  OP_AND, OP_OR, OP_XOR, OP_NOT, OP_IMPLY,
  OP_SEL,
  OP_LOOP, OP_ENDLOOP, OP_CSRLOOP, OP_CSREND,
  OP_RANDU, OP_RANDN, OP_RANDE, OP_RANDC,
  OP_REDOUT, OP_REDIN, OP_STATUS,
  OP_EXPM1, OP_LOG1P,
  OP_PLD, OP_PLDC, OP_PST, OP_PCONST, OP_PLUT, OP_PSHARE, OP_PCOUNT, OP_PBERN, OP_PRED,
  OP_SUMLOOP,
  OP_SAMPLE,
  OP_PPOLSEL,
  OP_PPOL,
  OP_COUNT_
};

// (OP_PPOL / OP_PPOLSEL: plane ops the host inserts for device-side policies -- the Bernoulli draw of one or
// two action planes, the max-nondef selection; never in an op table)
enum RddlRed { RED_SUM_I, RED_SUM_F, RED_PROD_I, RED_PROD_F, RED_MIN_I, RED_MIN_F, RED_MAX_I, RED_MAX_F,
               RED_AND, RED_OR };
enum RddlDtype { DT_BOOL = 0, DT_INT = 1, DT_REAL = 2,
                 DT_PACKED = 3 };   // bool fluent stored bit-packed: cell i -> bit (i % 32) of uint32 word i / 32
enum RddlKind { KIND_SHARED = 0, KIND_ENV = 1 };

struct RddlInsn {      // 16 bytes
  uint8_t op, flags;
  uint16_t dst, a, b, c, pad;
  uint32_t imm;
};

struct RddlAddr {      // 184 bytes: element offset = env*env_stride + c0 + sum coef*idx + sum clamp(reg)*rstride
  int32_t tensor;      // >= 0 tensor id; < 0: variate slot -(1+slot)
  int32_t naxis;
  int64_t env_stride;
  int64_t c0;
  uint8_t axis[8];
  int64_t coef[8];
  int32_t nreg, pad;
  int32_t reg[4];
  int64_t rstride[4];
  int64_t rextent[4];
};

struct RddlCsr {       // 176 bytes
  int64_t rowptr_off;
  int32_t ncols, col4_u16;
  int64_t col_off[4];
  int32_t slot[4];
  int64_t row_c0;
  int32_t row_naxis, pad2;
  uint8_t row_axis[8];
  int64_t row_coef[8];
  // single-column lists: a copy whose rows start on 16 bytes and hold a multiple of four int32 entries -- or, when
  // col4_u16 is set, of eight uint16 entries -- padded with sentinel4 (= the extent of the reduced axis: one past
  // the last element), in the order lowering._bank_schedule chose; rowptr4 counts entries; rowptr4_off == 0: absent
  int64_t rowptr4_off, col4_off, sentinel4;
};

struct RddlTensor { int16_t dtype, pair; int32_t kind; int64_t nelem; };   // pair: next-state tensor id of a state fluent, else -1

struct RddlLaunch {    // 176 bytes
  int32_t plan, rank;
  int64_t extent[8];
  int64_t E;
  int32_t pc_begin, pc_end;
  int32_t nred, G, ipt, chunks;
  int32_t red_op[8];
  int32_t red_slot[8];
  int32_t nregs, kind;   // kind: 0 = scalar level interpreter, 1 = bit-plane interpreter,
                         // 2 = scalar program fused into the epilogue of the preceding plane launch,
                         // 3 = FP64 contraction out[e,n] = sum_k x[e,k] W(k,n): extent = {N, K},
                         //     red_slot = {W, x, out} tensor ids, red_op = {W stride of k, of n}
};

struct RddlRedSlot { int32_t op, chunks; };

// Device-side random policy (rddl_rollout_policy / rddl_policy_sample): how each action fluent is drawn
// inside the kernels instead of being read from a caller buffer. Philox node ids of the draws:
// RDDL_POLICY_NODE + tensor id for the values, RDDL_POLICY_SELECT_NODE for the max-nondef selection.
#define RDDL_MAX_POLICY 8
#define RDDL_POLICY_MAX_SELECT 32
#define RDDL_POLICY_NODE 0xF000u
#define RDDL_POLICY_SELECT_NODE 0xFF00u
enum RddlPolicyKind { POL_NONE = 0, POL_BERNOULLI = 1, POL_UNIFORM_REAL = 2, POL_UNIFORM_INT = 3, POL_DEFAULT = 4 };
struct VmPolicyEntry {   // 80 bytes
  int32_t tensor, kind;
  double lo, hi;         // uniform bounds (POL_BERNOULLI: lo = p)
  uint64_t thr;          // POL_BERNOULLI: the draw is [thr_lo <= U < thr], U the cell's bit-sliced uniform under `node`;
  uint64_t dflt;         //   unpaired: thr_lo = 0, thr = floor(p * 2^64).   dflt: raw 64-bit default (noop) value
  int64_t base, nelem;   // first flat grounding index of this fluent in the policy's grounding order, #groundings
  int32_t always, pair;  // POL_BERNOULLI: p >= 1; pair: 0 = own uniform, 1 / 2 = first / second of two Bernoulli
  uint64_t thr_lo;       //   fluents drawn from one uniform per cell (second: thr_lo = T(p1) - T(p1 p2), thr = thr_lo + T(p2))
  uint32_t node, pad;    // Philox node id of the draw (the first fluent's for a pair)
};
struct VmPolicy {
  int32_t n, k;          // entries; k > 0: only k randomly chosen groundings per (env, step) keep their draw
  int64_t total;         // groundings over all entries
  int32_t* sel;          // scratch int32[n_envs, k]: the selected groundings of the current step (scalar kernels)
  uint8_t of_tensor[RDDL_MAX_TENSORS];   // tensor id -> entry index + 1 (0: not policy-driven)
  VmPolicyEntry e[RDDL_MAX_POLICY];
};
struct RddlVariate { int32_t node, kind; int64_t nelem; };

static_assert(sizeof(RddlInsn) == 16, "insn");
static_assert(sizeof(RddlAddr) == 184, "addr");
static_assert(sizeof(RddlCsr) == 176, "csr");
static_assert(sizeof(RddlLaunch) == 176, "launch");
