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
#include "rddl_contract.cuh"
#include "rddl_contract_tc.cuh"
#include "rddl_device.cuh"

#include "rddl_plane.cuh"

// Functor form of the interpreter (a named type instead of a lambda: this kernel is built ahead of time).
template <int NREG>
struct LevelInterpExec {
  const VmArgs& A;
  __device__ __forceinline__ void operator()(int64_t env, int32_t* idx, uint64_t* redval) const {
    vm_exec_element<NREG>(A, A.L, env, A.step, idx, redval);
  }
};

template <int NREG>
__global__ void __launch_bounds__(256) rddl_level_interp(const __grid_constant__ VmArgs A) {
  level_tile(A, A.L, A.red_off, LevelInterpExec<NREG>{A});
}

// ---------------------------------------------------------------------------------------
// host side: op-table parsing and the C ABI
// ---------------------------------------------------------------------------------------

static thread_local std::string g_create_error;

#include "rddl_jit.h"

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda):
// tensor maps of the TMA copies (non-fluent plane staging, tcgen05 contraction operands)
typedef CUresult (*TcEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TcEncodeFn tc_encode_fn() {
  static TcEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaDriverEntryPointQueryResult q;
    void* p = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (TcEncodeFn)p;
    cudaGetLastError();
  }
  return fn;
}


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
  // split-K scratch of the FP64 contraction (rddl_contract.cuh): partial tiles and per-tile arrival counters
  double* d_ct_partial = nullptr;
  unsigned int* d_ct_arrived = nullptr;
  size_t ct_partial_bytes = 0;
  int64_t ct_tiles = 0;
  uint8_t* d_alive = nullptr;           // rollout (general path): env still running
  int64_t alive_capacity = 0;
  // device-side policies (rddl_rollout_policy): plane programs rewritten per policy (the action planes
  // are drawn in registers), scratch for the max-nondef selection, and which tensors some kernel reads
  // through a raw pointer (SUMLOOP terms, contraction operands, stencil halos) -- those are
  // materialised in HBM each step instead
  struct PolicyProg { int launch; uint64_t polsig, sig; PlaneProg G; };
  std::vector<PolicyProg*> policy_progs;
  std::vector<uint8_t> direct_read;
  int32_t* d_polsel = nullptr;
  int64_t polsel_capacity = 0;
  // tcgen05 path of the contraction (rddl_contract_tc.cuh; option 2): bf16 splits of the shared W
  // matrices (made once per W tensor) and of the per-step x operand
  RddlTensorMap pool_map;
  int pool_map_ok = 0;
  int64_t pool_len = 0;
  int contract_tc = 0;
  struct TcWeights { const void* W; int N, K; int64_t sk, sn; __nv_bfloat16* Wb; int Np, Kp, bn; CUtensorMap map; };
  std::vector<TcWeights> tc_w;
  __nv_bfloat16* d_xb = nullptr;
  int64_t xb_capacity = 0;
  CUtensorMap xb_map;                     // tensor map of d_xb for the padded shape (xb_map_rows, xb_map_k)
  int64_t xb_map_rows = 0, xb_map_k = 0;
  // run-time specialised plane kernels (rddl_jit.h), built on first use per (launch, tile geometry)
  bool specialize = true;
  std::vector<RddlInsn> h_insns;          // host copies for the specialised epilogue (rddl_jit.h)
  std::vector<uint64_t> h_consts;
  std::vector<RddlAddr> h_addrs;
  std::vector<RddlCsr> h_csrs;
  std::vector<uint8_t> dtypes() const {
    std::vector<uint8_t> d(RDDL_MAX_TENSORS, 0);
    for (size_t i = 0; i < tensors.size(); ++i) d[i] = (uint8_t)tensors[i].dtype;
    return d;
  }
  // tensors a device-side policy may drive (= action fluents as far as the op table shows: per-env, not
  // one half of a state pair, never written by the program); the specialised kernels compile the policy
  // draw into the LOADs of these tensors only
  std::vector<uint8_t> may_policy;
  JitPools pools() const {
    return JitPools{h_insns.data(), (int64_t)h_insns.size(), h_consts.data(), (int64_t)h_consts.size(),
                    h_addrs.data(), (int64_t)h_addrs.size(), redslots.data(), (int64_t)redslots.size(), red_off.data(),
                    h_csrs.data(), (int64_t)h_csrs.size(), may_policy.data()};
  }
  // specialised scalar kernels (rddl_level_static.cuh): one module per program, built once the
  // program has issued `level_hot` scalar launches (tiered: short-lived programs never pay for it)
  int64_t scalar_launches = 0, level_hot = 24;
  bool level_tried = false;
  CUmodule level_mod = nullptr;
  std::vector<CUfunction> level_fn;       // per launch index, nullptr = interpreter kernel
  std::vector<CUfunction> fused_fn;       // per launch index: fused env-tile kernel of the run starting there
  std::vector<int> fused_count;           // ... and the number of launches it covers
  std::vector<JitKernel> jit;
  int64_t jit_built = 0, jit_cache_hits = 0, jit_failed = 0;
  std::string jit_err;
};

// Resident CTAs per SM the specialised kernel is compiled for (its register budget): the plane
// registers are machine registers there. Launches of many waves (big grids / big batches) run best
// at 3 CTAs x 256 threads (<= 85 registers: Wildfire-sized programs still do not spill) -- more
// warps to hide the DRAM latency of each new tile; single-wave launches (fused rollouts of a few
// thousand envs) keep 2 CTAs with the full register budget, one-word-per-thread tiles 4.
// RDDL_PLANE_MINB overrides (tuning knob).
static int plane_static_min_blocks(int wpt, const RddlLaunch& adj, int64_t n_envs) {
  if (const char* e = getenv("RDDL_PLANE_MINB")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 8) return v;
  }
  const int64_t tiles = (n_envs + adj.G - 1) / adj.G * adj.chunks;
  if (tiles >= 148 * 2 * 3) return 3;
  return wpt == 1 ? 4 : PLANE_MIN_BLOCKS;
}

// The specialised kernel of plane launch `li` for the tile geometry `adj` / wpt, or nullptr (then the
// interpreter kernel runs): compiled with NVRTC on first use, cubins cached on disk.
static JitKernel* jit_get(rddl_program* g, int li, int wpt, int blk, const RddlLaunch& adj, int64_t n_envs,
                          const PlaneProg* prog = nullptr, uint64_t sig = 0) {
  if (!g->specialize) return nullptr;
  const int minb = plane_static_min_blocks(wpt, adj, n_envs);
  for (JitKernel& k : g->jit)
    if (k.launch == li && k.wpt == wpt && k.G == adj.G && k.ipt == adj.ipt && k.minb == minb && k.sig == sig)
      return k.failed ? nullptr : &k;
  JitKernel k;
  k.launch = li; k.wpt = wpt; k.G = adj.G; k.ipt = adj.ipt; k.minb = minb; k.sig = sig;
  uint8_t dt[RDDL_MAX_TENSORS] = {0};
  for (size_t i = 0; i < g->tensors.size(); ++i) dt[i] = (uint8_t)g->tensors[i].dtype;
  const std::string data = plane_static_data(prog ? *prog : g->plane_progs[li], adj, dt, wpt, minb, blk, g->pools());
  std::string cubin, err;
  bool hit = false;
  JitDriver& drv = jit_driver();
  if (!drv.ok) err = "driver entry points unavailable";
  else if (jit_compile_plane(data, &cubin, &err, &hit)) {
    if (drv.ModuleLoadData(&k.mod, cubin.data()) != CUDA_SUCCESS) err = "cuModuleLoadData failed";
    else if (drv.ModuleGetFunction(&k.fn, k.mod, "rddl_plane_static") != CUDA_SUCCESS) err = "cuModuleGetFunction failed";
    else drv.FuncSetAttribute(k.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, 200 * 1024);
  }
  if (!err.empty()) {
    k.failed = true;
    k.fn = nullptr;
    ++g->jit_failed;
    g->jit_err = err;
  } else {
    ++g->jit_built;
    g->jit_cache_hits += hit;
  }
  g->jit.push_back(k);
  return g->jit.back().failed ? nullptr : &g->jit.back();
}

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
    if (in.op == OP_PLDC) o.b = G->ntma++;        // slot of the plane in the TMA-staged shared-memory region
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

// Builds (or fetches from the cubin cache) the specialised scalar kernels of the whole program.
static void jit_levels(rddl_program* g) {
  g->level_tried = true;
  g->level_fn.assign(g->launches.size(), nullptr);
  g->fused_fn.assign(g->launches.size(), nullptr);
  g->fused_count.assign(g->launches.size(), 0);
  const std::string data = level_static_data(g->pools(), g->launches.data(), (int)g->launches.size(), g->dtypes().data());
  if (data.empty()) return;
  std::string cubin, err;
  bool hit = false;
  JitDriver& drv = jit_driver();
  if (!drv.ok) err = "driver entry points unavailable";
  else if (jit_compile_level(data, &cubin, &err, &hit)) {
    if (drv.ModuleLoadData(&g->level_mod, cubin.data()) != CUDA_SUCCESS) err = "cuModuleLoadData failed";
    else {
      for (size_t i = 0; i < g->launches.size(); ++i) {
        if (g->launches[i].kind != 0) continue;
        char name[64];
        snprintf(name, sizeof name, "rddl_level_static_%d", (int)i);
        // (launches too long to unroll are not in the module: they keep the interpreter kernel)
        if (drv.ModuleGetFunction(&g->level_fn[i], g->level_mod, name) != CUDA_SUCCESS) g->level_fn[i] = nullptr;
        else ++g->jit_built;
      }
      // fused env-tile kernels (PS_FUSED in the generated constants): "X(first,count)"
      size_t pos = data.find("#define PS_FUSED(X)");
      if (pos != std::string::npos) {
        const size_t eol = data.find('\n', pos);
        std::string line = data.substr(pos, eol - pos);
        size_t q = 0;
        while ((q = line.find("X(", q + 1)) != std::string::npos && q > 15) {
          int first = 0, count = 0;
          if (sscanf(line.c_str() + q, "X(%d,%d)", &first, &count) == 2 && first >= 0 && (size_t)(first + count) <= g->launches.size()) {
            char name[64];
            snprintf(name, sizeof name, "rddl_level_fused_%d", first);
            CUfunction fn = nullptr;
            if (drv.ModuleGetFunction(&fn, g->level_mod, name) == CUDA_SUCCESS) {
              g->fused_fn[(size_t)first] = fn;
              g->fused_count[(size_t)first] = count;
              ++g->jit_built;
            }
          }
        }
      }
    }
  }
  if (!err.empty()) {
    ++g->jit_failed;
    g->jit_err = err;
  } else {
    g->jit_cache_hits += hit;
  }
}

// Words per thread and tile geometry of a plane launch for a batch of n_envs. The op table's tile is
// 4 words per thread x 256 threads (32768 cells per CTA: big tiles amortise the per-tile overhead),
// which is what large batches and multi-tile grids use. When the whole batch fits one wave of
// resident CTAs (small grids, a few thousand envs -- the fused-rollout case, where a CTA keeps
// its envs for every step), the envs per tile are chosen so that the busiest SM holds as few envs
// as possible, preferring more, smaller CTAs per SM (more warps to hide latency behind).
// RDDL_PLANE_FORCE_WPT overrides. Returns wpt and the adjusted launch (G = envs per tile,
// ipt = words per tile); *block = threads of the specialised kernel (the tile's words / wpt,
// rounded up to a warp: the interpreter kernels always run 256 threads and mask the rest).
static int plane_geometry(const RddlLaunch& L, int64_t n_envs, RddlLaunch* adj, int* block, bool balanced = true) {
  *adj = L;
  *block = 256;
  int wpt = L.ipt <= 256 ? 1 : (L.ipt <= 512 ? 2 : 4);
  if (L.chunks != 1 || wpt == 1) return wpt;
  const int64_t wpe = L.E >> 5;
  auto set = [&](int w, int ept) {
    adj->G = ept;
    adj->ipt = (int)(ept * wpe);
    *block = (int)(((adj->ipt + w - 1) / w + 31) / 32 * 32);
    if (*block > 256) *block = 256;
    return w;
  };
  if (const char* force = getenv("RDDL_PLANE_FORCE_WPT")) {
    const int want = atoi(force);
    if ((want == 1 || want == 2 || want == 4) && want <= wpt && wpe <= 256 * want) {
      int ept = (int)((256 * want) / wpe);
      if (const char* fe = getenv("RDDL_PLANE_FORCE_EPT")) { const int v = atoi(fe); if (v >= 1 && v <= ept) ept = v; }
      return set(want, ept > 32 ? 32 : ept);
    }
    return wpt;
  }
  const int sms = 148;
  const int64_t tiles_full = (n_envs + L.G - 1) / L.G;
  if (!balanced) {
    // interpreter kernels (fixed 256 threads, masked partial tiles are slow): full tiles only; one
    // word per thread when the batch would otherwise leave half of the SMs idle
    if (tiles_full < 74 && wpe <= 256) {
      int ept = (int)(256 / wpe);
      return set(1, ept > 32 ? 32 : ept);
    }
    return wpt;
  }
  if (tiles_full >= (int64_t)sms * 2 * 3) return wpt;     // several waves anyway: keep the big tiles
  int best_w = 0, best_ept = 0;
  int64_t best_score = -1;
  const int order[3] = {2, 1, 4};                          // tie-break: preference order
  for (int k = 0; k < 3; ++k) {
    const int w = order[k];
    if (w > wpt || wpe > 256 * w) continue;
    int cap = (int)((256 * w) / wpe);
    if (cap > 32) cap = 32;
    const int resident = w == 1 ? 4 : 2;
    for (int ept = cap; ept >= 1 && ept * 2 > cap; --ept) {
      const int64_t tiles = (n_envs + ept - 1) / ept;
      const int64_t per_sm = (tiles + sms - 1) / sms;
      if (per_sm > resident) continue;
      const int64_t score = per_sm * ept * 8 + k;          // envs on the busiest SM, then preference
      if (best_score < 0 || score < best_score) { best_score = score; best_w = w; best_ept = ept; }
    }
  }
  if (best_w == 0) return wpt;
  return set(best_w, best_ept);
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

// Host-only part of create: parses the op table, pre-decodes the plane launches.
static int parse_program(const void* optable, size_t nbytes, int device, rddl_program** out, const uint8_t** secs,
                         int64_t* hdr_out) {
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
  if (const char* e = getenv("RDDL_B200_JIT")) g->specialize = atoi(e) != 0;
  g->tensors.assign((const RddlTensor*)sec[0], (const RddlTensor*)sec[0] + hdr[2]);
  g->launches.assign((const RddlLaunch*)sec[1], (const RddlLaunch*)sec[1] + hdr[3]);
  g->redslots.assign((const RddlRedSlot*)sec[7], (const RddlRedSlot*)sec[7] + hdr[9]);
  g->variates.assign((const RddlVariate*)sec[8], (const RddlVariate*)sec[8] + hdr[10]);
  g->n_insns = hdr[4];
  g->h_insns.assign((const RddlInsn*)sec[2], (const RddlInsn*)sec[2] + hdr[4]);
  g->h_consts.assign((const uint64_t*)sec[3], (const uint64_t*)sec[3] + hdr[5]);
  g->h_addrs.assign((const RddlAddr*)sec[4], (const RddlAddr*)sec[4] + hdr[6]);
  g->h_csrs.assign((const RddlCsr*)sec[5], (const RddlCsr*)sec[5] + hdr[7]);
  if (const char* e = getenv("RDDL_B200_JIT_HOT")) g->level_hot = atoll(e);
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
    // the Philox node id of every Bernoulli op, so that the kernel need not fetch the variate table
    for (int a = 0; a < G->nops; ++a) {
      PlaneOp& o = G->ops[a];
      if (o.op != OP_PBERN) continue;
      if (o.a < 0 || (size_t)o.a >= g->variates.size()) { g_create_error = "plane op: bad variate slot"; delete g; return RDDL_ERR_FORMAT; }
      o.b |= (g->variates[o.a].node & 0xffff) << 8;
    }
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
    // The fused scalar epilogue runs through the scalar LOAD / STORE / SUMLOOP, which know neither the
    // per-step strides of an action sequence nor the state carried on chip: a fused program that
    // touches any fluent other than the outputs (__reward, __done, __chk) keeps the per-step path.
    for (int f = 0; f < G->nfused && ok; ++f) {
      for (int pc = G->fused[f].pc_begin; pc < G->fused[f].pc_end && ok; ++pc) {
        const RddlInsn& in = g->h_insns[pc];
        auto fluent = [&](uint32_t ai) {
          if (ai >= g->h_addrs.size()) return true;
          const int t = g->h_addrs[ai].tensor;
          return t >= 0 && t < reward_tensor && g->tensors[t].kind == KIND_ENV;
        };
        if ((in.op == OP_LOAD || in.op == OP_STORE) && fluent(in.imm)) ok = false;
        if (in.op == OP_SUMLOOP) {
          if (fluent(in.imm)) ok = false;
          if ((in.b & 2u) && pc + 1 < G->fused[f].pc_end && fluent(g->h_insns[pc + 1].imm)) ok = false;
          ++pc;   // data word
        }
      }
    }
    G->rollout_ok = ok;
    // TMA staging of the pre-packed non-fluent planes (rddl_plane.cuh: cp.async.bulk.tensor.1d into shared
    // memory + mbarrier): implemented and parity-tested, but measured SLOWER than reading the (L1 / L2
    // resident, 4 KB per tile) planes directly on both Wildfire configurations -- 1024 x 1024 x 4096 envs:
    // 2429 -> 2752 us per step, 32 x 32 x 4096 envs fused rollout: 344 -> 373 us --
    // so it is off unless the program is created under RDDL_PLANE_TMA=1.
    {
      bool tma = false;
      if (const char* e = getenv("RDDL_PLANE_TMA")) tma = atoi(e) != 0;
      if (!tma) G->ntma = 0;
    }
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
  int64_t acc = 0;
  for (auto& s : g->redslots) { g->red_off.push_back(acc); acc += s.chunks; }
  g->red_total = acc;
  // tensors some kernel reads through a raw pointer instead of LOAD / PLD (see rddl_program::direct_read)
  g->direct_read.assign(g->tensors.size(), 0);
  for (size_t li = 0; li < g->launches.size(); ++li) {
    const RddlLaunch& L = g->launches[li];
    auto mark = [&](int64_t t) { if (t >= 0 && (size_t)t < g->direct_read.size()) g->direct_read[(size_t)t] = 1; };
    if (L.kind == 3) { mark(L.red_slot[0]); mark(L.red_slot[1]); mark(L.red_slot[2]); continue; }
    for (int pc = L.pc_begin; pc < L.pc_end; ++pc) {
      const RddlInsn& in = g->h_insns[pc];
      if (L.kind == 1) {
        if (in.op == OP_PSHARE && in.imm != 0xffffffffu && L.chunks > 1) mark(in.imm);
      } else if (in.op == OP_SUMLOOP) {
        if (in.imm < g->h_addrs.size()) mark(g->h_addrs[in.imm].tensor);
        if ((in.b & 2u) && pc + 1 < L.pc_end && g->h_insns[pc + 1].imm < g->h_addrs.size()) mark(g->h_addrs[g->h_insns[pc + 1].imm].tensor);
        ++pc;
      }
    }
  }
  g->may_policy.assign(RDDL_MAX_TENSORS, 0);
  for (int t = 0; t < reward_tensor && t < RDDL_MAX_TENSORS; ++t)
    g->may_policy[(size_t)t] = g->tensors[(size_t)t].kind == KIND_ENV && g->tensors[(size_t)t].pair < 0;
  for (auto& tt : g->tensors)        // (primed partners of state fluents)
    if (tt.pair >= 0 && tt.pair < RDDL_MAX_TENSORS) g->may_policy[(size_t)tt.pair] = 0;
  for (size_t li = 0; li < g->launches.size(); ++li) {
    const RddlLaunch& L = g->launches[li];
    if (L.kind == 3) { if (L.red_slot[2] >= 0 && L.red_slot[2] < RDDL_MAX_TENSORS) g->may_policy[(size_t)L.red_slot[2]] = 0; continue; }
    if (L.kind == 1) {
      const PlaneProg& G = g->plane_progs[li];
      for (int a = 0; a < G.nops; ++a)
        if (G.ops[a].op == OP_PST && G.ops[a].imm < RDDL_MAX_TENSORS) g->may_policy[G.ops[a].imm] = 0;
      continue;
    }
    for (int pc = L.pc_begin; pc < L.pc_end; ++pc) {
      const RddlInsn& in = g->h_insns[pc];
      if (in.op == OP_SUMLOOP) { ++pc; continue; }
      if (in.op == OP_STORE && in.imm < g->h_addrs.size()) {
        const int t = g->h_addrs[in.imm].tensor;
        if (t >= 0 && t < RDDL_MAX_TENSORS) g->may_policy[(size_t)t] = 0;
      }
    }
  }
  for (int i = 0; i < 9; ++i) secs[i] = sec[i];
  memcpy(hdr_out, hdr, sizeof hdr);
  *out = g;
  return RDDL_OK;
}

extern "C" int rddl_program_create(const void* optable, size_t nbytes, int device, rddl_program_t** out) {
  rddl_program* g = nullptr;
  const uint8_t* sec[9];
  int64_t hdr[16];
  int rc = parse_program(optable, nbytes, device, &g, sec, hdr);
  if (rc) return rc;
  std::string* errp = &g->err;
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
  // tensor map over the pool (1-D, int32, box of RDDL_TMA_BOX words) for the plane kernels' TMA staging of
  // non-fluent planes; absent driver entry point or an empty pool: the kernels read the pool directly
  g->pool_len = hdr[8];
  if (hdr[8] > 0) {
    if (TcEncodeFn enc = tc_encode_fn()) {
      const cuuint64_t dims[1] = {(cuuint64_t)hdr[8]};
      const cuuint64_t strides[1] = {4};
      const cuuint32_t box[1] = {RDDL_TMA_BOX};
      const cuuint32_t estr[1] = {1};
      static_assert(sizeof(CUtensorMap) == sizeof(RddlTensorMap), "tensor map size");
      g->pool_map_ok = enc(reinterpret_cast<CUtensorMap*>(&g->pool_map), CU_TENSOR_MAP_DATA_TYPE_INT32, 1, g->d_pool, dims, strides, box,
                           estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
  }
  *out = g;
  return RDDL_OK;
}

// Inspection aid: the constants ("plane_static_data.inc") the specialised build of plane launch
// `launch` gets for a batch of n_envs. Host only. Returns the text length (without the NUL), or < 0.
extern "C" int64_t rddl_plane_source(const void* optable, size_t nbytes, int launch, int64_t n_envs, char* out,
                                     size_t cap) {
  rddl_program* g = nullptr;
  const uint8_t* sec[9];
  int64_t hdr[16];
  int rc = parse_program(optable, nbytes, 0, &g, sec, hdr);
  if (rc) return rc;
  if (launch < 0 || (size_t)launch >= g->launches.size() || g->launches[launch].kind != 1) {
    g_create_error = "rddl_plane_source: not a plane launch";
    delete g;
    return RDDL_ERR_ARG;
  }
  RddlLaunch adj;
  int blk = 256;
  const int wpt = plane_geometry(g->launches[launch], n_envs, &adj, &blk);
  uint8_t dt[RDDL_MAX_TENSORS] = {0};
  for (size_t i = 0; i < g->tensors.size(); ++i) dt[i] = (uint8_t)g->tensors[i].dtype;
  const std::string s = plane_static_data(g->plane_progs[launch], adj, dt, wpt, plane_static_min_blocks(wpt, adj, n_envs), blk, g->pools());
  delete g;
  if (out && cap > 0) {
    const size_t n = s.size() < cap - 1 ? s.size() : cap - 1;
    memcpy(out, s.data(), n);
    out[n] = 0;
  }
  return (int64_t)s.size();
}

// Host only (no device needed): compiles the specialised kernels of every plane launch for a batch
// of n_envs into the disk cache. Returns the number of kernels compiled or found cached, < 0 on error.
extern "C" int rddl_plane_precompile(const void* optable, size_t nbytes, int64_t n_envs) {
  rddl_program* g = nullptr;
  const uint8_t* sec[9];
  int64_t hdr[16];
  int rc = parse_program(optable, nbytes, 0, &g, sec, hdr);
  if (rc) return rc;
  uint8_t dt[RDDL_MAX_TENSORS] = {0};
  for (size_t i = 0; i < g->tensors.size(); ++i) dt[i] = (uint8_t)g->tensors[i].dtype;
  int n = 0;
  for (size_t i = 0; i < g->launches.size(); ++i) {
    if (g->launches[i].kind != 1) continue;
    RddlLaunch adj;
    int blk = 256;
    const int wpt = plane_geometry(g->launches[i], n_envs, &adj, &blk);
    std::string cubin, err;
    if (!jit_compile_plane(plane_static_data(g->plane_progs[i], adj, dt, wpt, plane_static_min_blocks(wpt, adj, n_envs), blk, g->pools()), &cubin,
                           &err, nullptr)) {
      g_create_error = err;
      delete g;
      return RDDL_ERR_FORMAT;
    }
    ++n;
  }
  {
    const std::string data = level_static_data(g->pools(), g->launches.data(), (int)g->launches.size(), g->dtypes().data());
    std::string cubin, err;
    if (!data.empty()) {
      if (!jit_compile_level(data, &cubin, &err, nullptr)) {
        g_create_error = err;
        delete g;
        return RDDL_ERR_FORMAT;
      }
      ++n;
    }
  }
  delete g;
  return n;
}

extern "C" void rddl_program_destroy(rddl_program_t* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaFree(g->d_insns); cudaFree(g->d_consts); cudaFree(g->d_addrs); cudaFree(g->d_csrs);
  cudaFree(g->d_pool); cudaFree(g->d_variates); cudaFree(g->d_redslots); cudaFree(g->d_red_off);
  cudaFree(g->d_red); cudaFree(g->d_ct_partial); cudaFree(g->d_ct_arrived);
  cudaFree(g->d_alive);
  cudaFree(g->d_polsel);
  cudaFree(g->d_xb);
  for (auto& w : g->tc_w) cudaFree(w.Wb);
  for (auto* pp : g->policy_progs) delete pp;
  for (JitKernel& k : g->jit)
    if (k.mod) jit_driver().ModuleUnload(k.mod);
  if (g->level_mod) jit_driver().ModuleUnload(g->level_mod);
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
    case 5: return g->jit_built;        // specialised plane kernels in use
    case 6: return g->jit_failed;       // plane launches that fell back to the interpreter kernel (rddl_jit_error)
    case 7: return g->jit_cache_hits;
    case 8: { JitApi& a = jit_api(); return a.ok ? a.major * 100 + a.minor : 0; }   // NVRTC version in use
    case 9: { int64_t n = 0; for (auto f : g->fused_fn) n += f != nullptr; return n; }   // fused env-tile kernels in use
    default:
      // 1000 + tensor id: 1 if some kernel reads the tensor through a raw pointer (a device-side policy
      // then materialises it in the caller's buffer every step instead of drawing it in registers only)
      if (what >= 1000 && (size_t)(what - 1000) < g->direct_read.size()) return g->direct_read[(size_t)(what - 1000)];
      return -1;
  }
}

struct RolloutArgs {
  int n_steps;
  double gamma;
  double* returns;
  double* rewards;
  const int64_t* stride_bytes;
};

// ---------------------------------------------------------------------------------------
// device-side policies
// ---------------------------------------------------------------------------------------

// The groundings that keep their draw at `step`, one thread per env: sel[env, 0..k)
__global__ void rddl_policy_select_kernel(const __grid_constant__ VmPolicy pol, int32_t* sel, int64_t n_envs,
                                          int64_t env_offset, uint64_t seed, uint64_t step) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n_envs) return;
  int32_t* s = sel + env * pol.k;
  for (int i = 0; i < pol.k; ++i)
    s[i] = (int32_t)rddl_policy_pick(seed, (uint32_t)(env_offset + env), (uint32_t)step, (uint32_t)i, pol.total - pol.k + i);
  rddl_policy_floyd(s, s, pol.k, pol.total);
}

// RandomAgent.sample_action for every env (pyRDDLGym/core/policy.py:166-178): writes the step's
// actions of one policy entry into its action tensor, one thread per (env, grounding).
__global__ void rddl_policy_sample_kernel(const __grid_constant__ VmPolicy pol, int entry, void* out, int dtype,
                                          int64_t n_envs, int64_t env_offset, uint64_t seed, uint64_t step) {
  const VmPolicyEntry& e = pol.e[entry];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_envs * e.nelem) return;
  const int64_t env = i / e.nelem, elem = i - env * e.nelem;
  const uint64_t v = rddl_policy_value(pol, e, seed, (uint64_t)(env_offset + env), step, elem, pol.sel + env * pol.k);
  if (dtype == DT_BOOL) reinterpret_cast<uint8_t*>(out)[i] = v != 0;
  else reinterpret_cast<uint64_t*>(out)[i] = v;
}

// check_default_action_count (pyRDDLGym/core/simulator.py:329-360) as a device reduction: one warp per
// env counts the groundings of the listed action tensors that differ from their default value.
struct ActionCountArgs {
  int32_t n;
  const void* ptr[16];
  int64_t nelem[16];
  uint64_t dflt[16];
  uint8_t dtype[16];
};
__global__ void rddl_action_count_kernel(const __grid_constant__ ActionCountArgs a, int32_t* counts, int64_t n_envs) {
  const int64_t env = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (env >= n_envs) return;
  int c = 0;
  for (int t = 0; t < a.n; ++t) {
    const int64_t n = a.nelem[t];
    if (a.dtype[t] == DT_BOOL) {
      const uint8_t* p = reinterpret_cast<const uint8_t*>(a.ptr[t]) + env * n;
      const bool d = a.dflt[t] != 0;
      for (int64_t i = lane; i < n; i += 32) c += (p[i] != 0) != d;
    } else if (a.dtype[t] == DT_PACKED) {     // bit-packed bool fluent (n is a multiple of 32): popcount of the differing bits
      const uint32_t* p = reinterpret_cast<const uint32_t*>(a.ptr[t]) + env * (n >> 5);
      const uint32_t d = a.dflt[t] != 0 ? 0xffffffffu : 0u;
      for (int64_t i = lane; i < (n >> 5); i += 32) c += __popc(p[i] ^ d);
    } else if (a.dtype[t] == DT_REAL) {
      const double* p = reinterpret_cast<const double*>(a.ptr[t]) + env * n;
      const double d = __longlong_as_double((long long)a.dflt[t]);
      for (int64_t i = lane; i < n; i += 32) c += p[i] != d;
    } else {
      const uint64_t* p = reinterpret_cast<const uint64_t*>(a.ptr[t]) + env * n;
      for (int64_t i = lane; i < n; i += 32) c += p[i] != a.dflt[t];
    }
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if (lane == 0) counts[env] = c;
}

static uint64_t fnv_bytes(uint64_t h, const void* p, size_t n) {
  const unsigned char* c = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 0x100000001b3ull; }
  return h;
}

// Validates the public descriptor and converts it to the kernels' form.
static int make_policy(rddl_program* g, const rddl_policy_t* P, VmPolicy* out) {
  memset(out, 0, sizeof *out);
  if (!P || P->n_entries < 0 || P->n_entries > RDDL_MAX_POLICY) { g->err = "policy: bad entry count"; return RDDL_ERR_ARG; }
  const int n_user = (int)g->tensors.size() - 3;
  int64_t total = 0;
  for (int i = 0; i < P->n_entries; ++i) {
    const rddl_policy_entry_t& pe = P->entries[i];
    VmPolicyEntry& e = out->e[i];
    if (pe.tensor < 0 || pe.tensor >= n_user || pe.tensor >= RDDL_MAX_TENSORS || !g->may_policy[(size_t)pe.tensor] ||
        out->of_tensor[pe.tensor]) {
      g->err = "policy: entry does not name an action tensor (or names one twice)";
      return RDDL_ERR_ARG;
    }
    const int dt = g->tensors[pe.tensor].dtype;
    const bool kind_ok = (pe.kind == POL_BERNOULLI && dt == DT_BOOL && pe.lo >= 0.0 && pe.lo <= 1.0) ||
                         (pe.kind == POL_UNIFORM_REAL && dt == DT_REAL && pe.lo <= pe.hi) ||
                         (pe.kind == POL_UNIFORM_INT && dt == DT_INT && pe.lo <= pe.hi) || pe.kind == POL_DEFAULT;
    if (!kind_ok || dt == DT_PACKED) { g->err = "policy: kind / range does not fit the fluent's type"; return RDDL_ERR_ARG; }
    e.tensor = pe.tensor; e.kind = pe.kind; e.lo = pe.lo; e.hi = pe.hi;
    e.dflt = dt == DT_REAL ? (uint64_t)0 : (dt == DT_BOOL ? (uint64_t)(pe.dflt != 0.0) : (uint64_t)(int64_t)pe.dflt);
    if (dt == DT_REAL) memcpy(&e.dflt, &pe.dflt, 8);
    e.node = RDDL_POLICY_NODE + (uint32_t)pe.tensor;
    if (pe.kind == POL_BERNOULLI) {
      e.always = pe.lo >= 1.0;
      e.thr = e.always ? ~0ull : (uint64_t)(pe.lo * 18446744073709551616.0);
    }
    e.base = total;
    e.nelem = g->tensors[pe.tensor].nelem;
    total += e.nelem;
    out->of_tensor[pe.tensor] = (uint8_t)(i + 1);
  }
  out->n = P->n_entries;
  out->total = total;
  // Two consecutive Bernoulli entries of the same size share ONE uniform per cell (half the Philox work):
  // a = [U < T(p)], b = [T(p) - T(pq) <= U < T(p) - T(pq) + T(q)] -- P(a) = p, P(b) = q, P(a and b) = pq.
  // (oracle/philox_np.py: policy_actions states the same rule.)
  auto T = [](double x) { return x >= 1.0 ? ~0ull : (uint64_t)(x * 18446744073709551616.0); };
  for (int i = 0; i + 1 < out->n;) {
    VmPolicyEntry &a = out->e[i], &b = out->e[i + 1];
    if (a.kind == POL_BERNOULLI && b.kind == POL_BERNOULLI && a.nelem == b.nelem && a.lo < 1.0 && b.lo < 1.0 &&
        (1.0 - a.lo) * (1.0 - b.lo) >= 2.3283064365386963e-10 /* 2^-32 */) {
      a.pair = 1;
      b.pair = 2;
      b.node = a.node;
      b.thr_lo = T(a.lo) - T(a.lo * b.lo);
      b.thr = b.thr_lo + T(b.lo);
      i += 2;
    } else {
      ++i;
    }
  }
  if (total >= (1ll << 31)) { g->err = "policy: more than 2^31 action groundings"; return RDDL_ERR_ARG; }
  if (P->max_nondef >= 0 && P->max_nondef < total) {
    if (P->max_nondef == 0) {
      for (int i = 0; i < out->n; ++i) out->e[i].kind = POL_DEFAULT;
    } else if (P->max_nondef > RDDL_POLICY_MAX_SELECT) {
      g->err = "policy: max_nondef between 33 and the number of groundings is not supported on the device";
      return RDDL_ERR_ARG;
    } else {
      out->k = P->max_nondef;
    }
  }
  return RDDL_OK;
}

// The plane program of launch li with the loads of policy-driven action planes replaced by in-register
// draws: PLD -> PBERN over a constant table (same bit-sliced sampler and Philox counters as a
// Bernoulli CPF, node id RDDL_POLICY_NODE + tensor) [+ PPOLSEL for the max-nondef rule] or PCONST.
static const rddl_program::PolicyProg* policy_prog(rddl_program* g, int li, const VmPolicy& pol) {
  uint64_t polsig = 0xcbf29ce484222325ull;
  polsig = fnv_bytes(polsig, &pol.n, sizeof pol.n);
  polsig = fnv_bytes(polsig, &pol.k, sizeof pol.k);
  for (int i = 0; i < pol.n; ++i) polsig = fnv_bytes(polsig, &pol.e[i], sizeof pol.e[i]);
  for (auto* pp : g->policy_progs)
    if (pp->launch == li && pp->polsig == polsig) return pp;
  auto* pp = new rddl_program::PolicyProg();
  pp->launch = li;
  pp->polsig = polsig;
  const PlaneProg& S = g->plane_progs[li];
  PlaneProg& G = pp->G;
  G = S;
  G.nops = 0;
  G.pol_k = pol.k;
  for (int i = 0; i < pol.n && i < RDDL_MAX_POLICY; ++i) {
    G.pol_thr[i][0] = pol.e[i].thr; G.pol_thr[i][1] = pol.e[i].thr_lo;
    G.pol_node[i] = pol.e[i].node; G.pol_always[i] = pol.e[i].always;
  }
  bool ok = true;
  int partner_done = -1;       // op index of a paired plane's load that the first plane's PPOL already covered
  G.npol = 0;
  for (int a = 0; a < S.nops && ok; ++a) {
    PlaneOp o = S.ops[a];
    if (a == partner_done) continue;
    const int pe = (o.op == OP_PLD && o.a == 0 && o.imm < RDDL_MAX_TENSORS) ? pol.of_tensor[o.imm] : 0;
    if (!pe) {
      if (G.nops >= PLANE_MAX_OPS) { ok = false; break; }
      G.ops[G.nops++] = o;
      continue;
    }
    const VmPolicyEntry& e = pol.e[pe - 1];
    if (e.kind == POL_DEFAULT) {
      o.op = OP_PCONST; o.imm = e.dflt ? 1u : 0u;
      if (G.nops >= PLANE_MAX_OPS) { ok = false; break; }
      G.ops[G.nops++] = o;
      continue;
    }
    if (e.kind != POL_BERNOULLI || G.nops + 3 > PLANE_MAX_OPS) { ok = false; break; }
    // the Bernoulli draw of this plane -- and, for the first of a pair whose partner's load follows
    // with nothing but loads in between, of both planes from the one shared uniform (OP_PPOL)
    o.op = OP_PPOL; o.m = 0; o.idx = 0; o.off = 0; o.imm = (uint32_t)(pe - 1); o.a = -1; o.b = -1;
    if (e.pair == 1) {
      for (int c = a + 1; c < S.nops; ++c) {
        const PlaneOp& oc = S.ops[c];
        if (oc.op == OP_PLD && oc.a == 0 && oc.imm < RDDL_MAX_TENSORS && pol.of_tensor[oc.imm] == pe + 1) {
          // (the partner's register is written here, before its own load: nothing in between may write it)
          bool clash = false;
          for (int d = a + 1; d < c; ++d) clash |= S.ops[d].dst == oc.dst;
          if (!clash && oc.dst != o.dst) { o.a = oc.dst; o.b = pe; partner_done = c; }
          break;
        }
        if (oc.op != OP_PLD && oc.op != OP_PLDC) break;
      }
    }
    ++G.npol;
    G.ops[G.nops++] = o;
    if (o.a >= 0 && pol.k > 0) {       // (the partner's selection mask: emitted here, its own load is skipped below)
      PlaneOp q; memset(&q, 0, sizeof q);
      q.op = OP_PPOLSEL; q.dst = o.a; q.imm = (uint32_t)pe;
      G.ops[G.nops++] = q;
    }
    if (pol.k > 0) {
      PlaneOp q; memset(&q, 0, sizeof q);
      q.op = OP_PPOLSEL; q.dst = o.dst; q.imm = (uint32_t)(pe - 1);
      G.ops[G.nops++] = q;
    }
  }
  if (!ok) { delete pp; return nullptr; }
  pp->sig = fnv_bytes(fnv_bytes(0xcbf29ce484222325ull, &G, sizeof G), "policy", 6) | 1ull;
  g->policy_progs.push_back(pp);
  return pp;
}

// ---------------------------------------------------------------------------------------
// tcgen05 contraction (rddl_contract_tc.cuh): tensor maps + launches
// ---------------------------------------------------------------------------------------
// Tensor map over the split layout [term][rows][K] (bf16): dims (fastest first) K, rows, 3; boxes of
// 64 k x box_rows rows x 1 term land in shared memory in the 128-byte-swizzled K-major layout.
static bool tc_make_map(CUtensorMap* map, void* base, int rows_p, int Kp, int box_rows) {
  TcEncodeFn enc = tc_encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)Kp, (cuuint64_t)rows_p, 3};
  const cuuint64_t strides[2] = {(cuuint64_t)Kp * 2, (cuuint64_t)rows_p * (cuuint64_t)Kp * 2};
  const cuuint32_t box[3] = {TC_BK, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static int run_contract_tc(rddl_program* g, const double* W, const double* x, double* out, int64_t n_envs, int N, int K,
                           int64_t sk, int64_t sn, cudaStream_t stream) {
  std::string* errp = &g->err;
  const int Kp = (K + TC_BK - 1) / TC_BK * TC_BK;
  const int64_t Mp = (n_envs + TC_BM - 1) / TC_BM * TC_BM;
  // N tile: 64 columns, or 32 when the 64-column grid would leave SMs without a CTA (RDDL_CONTRACT_TC_BN overrides)
  int bn = ((int64_t)((N + 63) / 64) * (Mp / TC_BM) < 148) ? 32 : 64;
  if (const char* e = getenv("RDDL_CONTRACT_TC_BN")) bn = atoi(e) == 32 ? 32 : 64;
  rddl_program::TcWeights* tw = nullptr;
  for (auto& w : g->tc_w)
    if (w.W == W && w.N == N && w.K == K && w.sk == sk && w.sn == sn && w.bn == bn) tw = &w;
  if (!tw) {
    rddl_program::TcWeights w;
    memset(&w, 0, sizeof w);
    w.W = W; w.N = N; w.K = K; w.sk = sk; w.sn = sn; w.bn = bn;
    w.Np = (N + TC_BN_MAX - 1) / TC_BN_MAX * TC_BN_MAX; w.Kp = Kp;
    CK(cudaMalloc((void**)&w.Wb, (size_t)3 * w.Np * w.Kp * sizeof(__nv_bfloat16)));
    const int64_t n = (int64_t)w.Np * (w.Kp / 8);
    rddl_tc_split<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(W, sn, sk, N, K, w.Np, w.Kp, w.Wb);   // B(n, k) = W(k, n)
    CK(cudaGetLastError());
    if (!tc_make_map(&w.map, w.Wb, w.Np, w.Kp, bn)) { cudaFree(w.Wb); g->err = "cuTensorMapEncodeTiled failed (W)"; return RDDL_ERR_CUDA; }
    g->tc_w.push_back(w);
    tw = &g->tc_w.back();
    CK(cudaFuncSetAttribute((const void*)rddl_contract_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES(32)));
    CK(cudaFuncSetAttribute((const void*)rddl_contract_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES(64)));
  }
  if (g->xb_capacity < 3 * Mp * Kp) {
    CK(cudaStreamSynchronize(stream));
    cudaFree(g->d_xb);
    g->d_xb = nullptr;
    CK(cudaMalloc((void**)&g->d_xb, (size_t)(3 * Mp * Kp) * sizeof(__nv_bfloat16)));
    g->xb_capacity = 3 * Mp * Kp;
    g->xb_map_rows = 0;
  }
  if (g->xb_map_rows != Mp || g->xb_map_k != Kp) {      // the x tensor map only depends on the buffer and its padded shape
    if (!tc_make_map(&g->xb_map, g->d_xb, (int)Mp, Kp, TC_BM)) { g->err = "cuTensorMapEncodeTiled failed (x)"; return RDDL_ERR_CUDA; }
    g->xb_map_rows = Mp; g->xb_map_k = Kp;
  }
  const int64_t n = Mp * (Kp / 8);
  rddl_tc_split<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(x, K, 1, n_envs, K, Mp, Kp, g->d_xb);
  CK(cudaGetLastError());
  // launched as a programmatic dependent of the split: its prologue overlaps the split's execution
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)((N + bn - 1) / bn), (unsigned)(Mp / TC_BM));
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = bn == 32 ? TC_SMEM_BYTES(32) : TC_SMEM_BYTES(64);
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const int nkb = Kp / TC_BK;
  if (bn == 32) CK(cudaLaunchKernelEx(&cfg, rddl_contract_tc<32>, g->xb_map, tw->map, out, n_envs, N, nkb));
  else CK(cudaLaunchKernelEx(&cfg, rddl_contract_tc<64>, g->xb_map, tw->map, out, n_envs, N, nkb));
  CK(cudaGetLastError());
  g->launches_issued += 1;     // (the caller counts the contraction launch itself)
  return RDDL_OK;
}

static void rddl_program_destroy_host(rddl_program* g) {   // programs parsed without a device (precompile)
  for (auto* pp : g->policy_progs) delete pp;
  delete g;
}

static int run_plan(rddl_program* g, int plan, void* const* tensors, const double* const* injected,
                    uint32_t* status, int64_t n_envs, int64_t env_offset, uint64_t seed, uint64_t step,
                    cudaStream_t stream, const RolloutArgs* ro = nullptr, const VmPolicy* pol = nullptr) {
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
  rddl_philox_round_keys(seed, &A.rk);
  A.n_steps = 1;
  if (ro) {
    A.n_steps = ro->n_steps; A.gamma = ro->gamma; A.returns = ro->returns; A.rewards = ro->rewards;
    if (ro->stride_bytes) for (size_t i = 0; i < g->tensors.size(); ++i) A.tstride[i] = ro->stride_bytes[i];
  }
  for (size_t i = 0; i < g->tensors.size(); ++i) { A.tensors[i] = tensors[i]; A.dtype[i] = (uint8_t)g->tensors[i].dtype; }
  if (injected) for (size_t i = 0; i < g->variates.size(); ++i) A.inj[i] = injected[i];
  A.pool_map = g->pool_map;
  A.pool_map_ok = g->pool_map_ok;
  if (pol && pol->n > 0) {
    A.pol = *pol;
    A.pol.sel = g->d_polsel;     // (filled for this step by policy_prepare)
  }
  for (size_t launch_i = 0; launch_i < g->launches.size(); ++launch_i) {
    const RddlLaunch& L = g->launches[launch_i];
    if (L.plan != plan || L.kind == 2) continue;   // kind 2 runs inside the preceding plane launch
    A.L = L;
    const int block = 256;
    int jit_block = 256;
    int wpt = 1;
    JitKernel* jk = nullptr;
    const rddl_program::PolicyProg* pp = nullptr;
    if (L.kind == 1) {
      const int li = (int)(&L - g->launches.data());
      if (A.pol.n > 0) {
        pp = policy_prog(g, li, A.pol);
        if (!pp) { g->err = "policy: the plane program of this launch cannot take the policy (table space)"; return RDDL_ERR_ARG; }
      }
      wpt = plane_geometry(L, n_envs, &A.L, &jit_block, g->specialize);
      jk = jit_get(g, li, wpt, jit_block, A.L, n_envs, pp ? &pp->G : nullptr, pp ? pp->sig : 0);
      if (!jk && g->specialize) wpt = plane_geometry(L, n_envs, &A.L, &jit_block, false);   // interpreter kernel
    }
    const int envs_per_cta = L.kind == 1 ? A.L.G : 256 / L.G;
    const int64_t groups = (n_envs + envs_per_cta - 1) / envs_per_cta;
    const int64_t blocks = groups * L.chunks;
    if (blocks > 0x7fffffffll) { g->err = "grid too large"; return RDDL_ERR_ARG; }
    rddl_program::Timed tm;
    // (bounded: a caller that switches profiling on and never reads it stops accumulating events after 64 Ki launches)
    const bool timing = g->profiling && g->timed.size() < (size_t)(1 << 16);
    if (timing) {
      tm.launch = (int)(&L - g->launches.data());
      CK(cudaEventCreate(&tm.e0));
      CK(cudaEventCreate(&tm.e1));
      CK(cudaEventRecord(tm.e0, stream));
    }
    if (L.kind == 3 && g->contract_tc) {
      int rc = run_contract_tc(g, (const double*)tensors[L.red_slot[0]], (const double*)tensors[L.red_slot[1]],
                               (double*)tensors[L.red_slot[2]], n_envs, (int)L.extent[0], (int)L.extent[1],
                               (int64_t)L.red_op[0], (int64_t)L.red_op[1], stream);
      if (rc) return rc;
    } else if (L.kind == 3) {
      const int N = (int)L.extent[0], K = (int)L.extent[1];
      // 64 x 64 CTA tiles; 64 x 32 when that grid would not give every SM two CTAs (RDDL_CONTRACT_BN overrides)
      int bn = ((int64_t)((N + 63) / 64) * ((n_envs + CT_BM - 1) / CT_BM) < 2 * 148) ? 32 : 64;
      if (const char* e = getenv("RDDL_CONTRACT_BN")) bn = atoi(e) == 32 ? 32 : 64;
      dim3 grid((unsigned)((N + bn - 1) / bn), (unsigned)((n_envs + CT_BM - 1) / CT_BM));
      // split K when the output tiles cover less than half of the SMs (small batches: 128 envs per GPU is
      // 16 tiles): up to 8 K ranges per tile. Measured on B200 at 512 x 512: 128 envs 28.5 -> 21.7 us with 4
      // ranges; 1024 envs (128 tiles) 36.7 us without, 42 / 43 / 39 us with 2 / 3 / 4 -- a full grid gains nothing
      // from more resident CTAs and pays for the partial tiles. RDDL_CONTRACT_SPLITK overrides (1 = off).
      const int64_t tiles = (int64_t)grid.x * grid.y;
      int S = tiles * 2 <= 148 ? (int)std::min<int64_t>(8, 148 / std::max<int64_t>(1, tiles)) : 1;
      if (const char* e = getenv("RDDL_CONTRACT_SPLITK")) S = atoi(e);
      S = std::max(1, std::min(S, (K + CT_BK - 1) / CT_BK));
      int kps = ((K + S - 1) / S + CT_BK - 1) / CT_BK * CT_BK;
      S = (K + kps - 1) / kps;
      if (S > 1) {
        const size_t need = (size_t)tiles * S * CT_BM * bn * sizeof(double);
        if (g->ct_partial_bytes < need || g->ct_tiles < tiles) {
          CK(cudaStreamSynchronize(stream));
          cudaFree(g->d_ct_partial); cudaFree(g->d_ct_arrived);
          g->d_ct_partial = nullptr; g->d_ct_arrived = nullptr;
          CK(cudaMalloc((void**)&g->d_ct_partial, need));
          CK(cudaMalloc((void**)&g->d_ct_arrived, (size_t)tiles * sizeof(unsigned int)));
          CK(cudaMemsetAsync(g->d_ct_arrived, 0, (size_t)tiles * sizeof(unsigned int), stream));
          g->ct_partial_bytes = need;
          g->ct_tiles = tiles;
        }
        grid.z = (unsigned)S;
      }
      const double* Xp = (const double*)tensors[L.red_slot[1]];
      const double* Wp = (const double*)tensors[L.red_slot[0]];
      double* Op = (double*)tensors[L.red_slot[2]];
      const int64_t wsk = (int64_t)L.red_op[0], wsn = (int64_t)L.red_op[1];
      // cp.async pipeline where 16-byte copies are possible (RDDL_CONTRACT_ASYNC=0: the register-staged loop)
      bool async = wsn == 1 && (K & 1) == 0 && (N & 1) == 0 && (wsk & 1) == 0 && ((uintptr_t)Xp & 15) == 0 && ((uintptr_t)Wp & 15) == 0;
      if (const char* e = getenv("RDDL_CONTRACT_ASYNC")) async = async && atoi(e) != 0;
      if (async) {
        static uint64_t attr_set = 0;          // per device (function attributes belong to the device's context)
        if (!((attr_set >> (g->device & 63)) & 1ull)) {
          CK(cudaFuncSetAttribute(rddl_contract_f64<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_ASYNC_SMEM(64)));
          CK(cudaFuncSetAttribute(rddl_contract_f64<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_ASYNC_SMEM(32)));
          attr_set |= 1ull << (g->device & 63);
        }
        if (bn == 64)
          rddl_contract_f64<64, true><<<grid, 128, CT_ASYNC_SMEM(64), stream>>>(Xp, Wp, Op, n_envs, N, K, wsk, wsn, g->d_ct_partial,
                                                                                g->d_ct_arrived, kps);
        else
          rddl_contract_f64<32, true><<<grid, 128, CT_ASYNC_SMEM(32), stream>>>(Xp, Wp, Op, n_envs, N, K, wsk, wsn, g->d_ct_partial,
                                                                                g->d_ct_arrived, kps);
      } else if (bn == 64)
        rddl_contract_f64<64><<<grid, 128, 0, stream>>>(Xp, Wp, Op, n_envs, N, K, wsk, wsn, g->d_ct_partial, g->d_ct_arrived, kps);
      else
        rddl_contract_f64<32><<<grid, 128, 0, stream>>>(Xp, Wp, Op, n_envs, N, K, wsk, wsn, g->d_ct_partial, g->d_ct_arrived, kps);
    } else if (L.kind == 1) {
      const PlaneProg& G = pp ? pp->G : g->plane_progs[(size_t)(&L - g->launches.data())];
      const int W32 = (int)((L.rank == 2 ? L.extent[1] : L.extent[0]) >> 5);
      const size_t smem = sizeof(uint32_t) * ((size_t)L.nregs * wpt * block +
                                              (size_t)G.nshared * ((A.L.ipt / W32 + A.L.G + 2) * (W32 + 2)) +
                                              32 * RDDL_MAX_RED +
                                              8 * PLANE_QUEUE + 8 + block * wpt + 32 * G.pol_k + (G.npol > 0 ? 4 * block * wpt : 0)) +
                          (G.ntma > 0 ? 256 + 128 + sizeof(uint32_t) * (size_t)G.ntma *
                               (size_t)(((L.chunks == 1 ? (int)(L.E >> 5) : A.L.ipt) + RDDL_TMA_BOX - 1) / RDDL_TMA_BOX * RDDL_TMA_BOX) : 0);
      const size_t smem_static = smem - sizeof(uint32_t) * ((size_t)L.nregs * wpt * block + (size_t)(block - jit_block) * wpt);
      if (jk) {
        void* params[] = {(void*)&A};
        if (jit_driver().LaunchKernel(jk->fn, (unsigned)blocks, 1, 1, (unsigned)jit_block, 1, 1, (unsigned)smem_static, (CUstream)stream,
                                      params, nullptr) != CUDA_SUCCESS) {
          g->err = "cuLaunchKernel failed (specialised plane kernel)";
          return RDDL_ERR_CUDA;
        }
      } else if (wpt == 1) rddl_plane_interp<1, 256><<<(unsigned)blocks, 256, smem, stream>>>(A, G);
      else if (wpt == 2) rddl_plane_interp<2, 256><<<(unsigned)blocks, 256, smem, stream>>>(A, G);
      else rddl_plane_interp<4, 256><<<(unsigned)blocks, 256, smem, stream>>>(A, G);
    } else if (L.kind == 0 && g->specialize && (g->level_tried || ++g->scalar_launches > g->level_hot) &&
               (g->level_tried || (jit_levels(g), true)) && g->fused_fn[launch_i]) {
      // this launch and the fused_count - 1 after it: one env-tile kernel (rddl_level_static.cuh)
      void* params[] = {(void*)&A};
      const int64_t tiles = (n_envs + RDDL_FUSE_TE - 1) / RDDL_FUSE_TE;
      if (jit_driver().LaunchKernel(g->fused_fn[launch_i], (unsigned)tiles, 1, 1, 256, 1, 1, 0, (CUstream)stream, params, nullptr) !=
          CUDA_SUCCESS) {
        g->err = "cuLaunchKernel failed (fused scalar kernel)";
        return RDDL_ERR_CUDA;
      }
      launch_i += (size_t)g->fused_count[launch_i] - 1;
    } else if (L.kind == 0 && g->specialize && g->level_tried && g->level_fn[(size_t)(&L - g->launches.data())]) {
      void* params[] = {(void*)&A};
      if (jit_driver().LaunchKernel(g->level_fn[(size_t)(&L - g->launches.data())], (unsigned)blocks, 1, 1, 256, 1, 1, 0,
                                    (CUstream)stream, params, nullptr) != CUDA_SUCCESS) {
        g->err = "cuLaunchKernel failed (specialised scalar kernel)";
        return RDDL_ERR_CUDA;
      }
    } else if (L.nregs <= 16) rddl_level_interp<16><<<(unsigned)blocks, 256, 0, stream>>>(A);
    else if (L.nregs <= 40) rddl_level_interp<40><<<(unsigned)blocks, 256, 0, stream>>>(A);
    else rddl_level_interp<RDDL_MAX_REGS><<<(unsigned)blocks, 256, 0, stream>>>(A);
    ++g->launches_issued;
    CK(cudaGetLastError());
    if (timing) {
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

// Per step of a policy-driven plan: the max-nondef selection of the step (scalar kernels read it from
// d_polsel; the plane kernels recompute it on chip) and the action tensors that some kernel reads
// through a raw pointer (`force`: all of them -- rddl_policy_sample).
static int policy_prepare(rddl_program* g, VmPolicy* pol, void* const* tensors, int64_t n_envs, int64_t env_offset,
                          uint64_t seed, uint64_t step, cudaStream_t stream, bool force) {
  std::string* errp = &g->err;
  if (pol->k > 0) {
    if (g->polsel_capacity < n_envs * pol->k) {
      CK(cudaStreamSynchronize(stream));
      cudaFree(g->d_polsel);
      g->d_polsel = nullptr;
      CK(cudaMalloc((void**)&g->d_polsel, (size_t)(n_envs * RDDL_POLICY_MAX_SELECT) * sizeof(int32_t)));
      g->polsel_capacity = n_envs * RDDL_POLICY_MAX_SELECT;
    }
    rddl_policy_select_kernel<<<(unsigned)((n_envs + 127) / 128), 128, 0, stream>>>(*pol, g->d_polsel, n_envs, env_offset, seed, step);
    CK(cudaGetLastError());
    ++g->launches_issued;
  }
  pol->sel = g->d_polsel;
  for (int i = 0; i < pol->n; ++i) {
    const VmPolicyEntry& e = pol->e[i];
    if (!force && !g->direct_read[(size_t)e.tensor]) continue;
    if (!tensors[e.tensor]) { g->err = "policy: action tensor pointer is NULL"; return RDDL_ERR_ARG; }
    const int64_t n = n_envs * e.nelem;
    rddl_policy_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(*pol, i, tensors[e.tensor], g->tensors[e.tensor].dtype,
                                                                                n_envs, env_offset, seed, step);
    CK(cudaGetLastError());
    ++g->launches_issued;
  }
  return RDDL_OK;
}

// Host only: the specialised plane kernels of the op table under `policy` for batches of n_envs, into
// the disk cache (see rddl_plane_precompile). Returns the number of kernels, < 0 on error.
extern "C" int rddl_policy_precompile(const void* optable, size_t nbytes, int64_t n_envs, const rddl_policy_t* policy) {
  rddl_program* g = nullptr;
  const uint8_t* sec[9];
  int64_t hdr[16];
  int rc = parse_program(optable, nbytes, 0, &g, sec, hdr);
  if (rc) return rc;
  VmPolicy pol;
  if ((rc = make_policy(g, policy, &pol))) { g_create_error = g->err; rddl_program_destroy_host(g); return rc; }
  uint8_t dt[RDDL_MAX_TENSORS] = {0};
  for (size_t i = 0; i < g->tensors.size(); ++i) dt[i] = (uint8_t)g->tensors[i].dtype;
  int n = 0;
  for (size_t i = 0; i < g->launches.size(); ++i) {
    if (g->launches[i].kind != 1 || g->launches[i].plan != RDDL_PLAN_STEP) continue;
    const rddl_program::PolicyProg* pp = policy_prog(g, (int)i, pol);
    if (!pp) { g_create_error = "policy: plane program cannot take the policy"; rddl_program_destroy_host(g); return RDDL_ERR_ARG; }
    RddlLaunch adj;
    int blk = 256;
    const int wpt = plane_geometry(g->launches[i], n_envs, &adj, &blk);
    std::string cubin, err;
    if (!jit_compile_plane(plane_static_data(pp->G, adj, dt, wpt, plane_static_min_blocks(wpt, adj, n_envs), blk, g->pools()), &cubin,
                           &err, nullptr)) {
      g_create_error = err;
      rddl_program_destroy_host(g);
      return RDDL_ERR_FORMAT;
    }
    ++n;
  }
  rddl_program_destroy_host(g);
  return n;
}

extern "C" int rddl_policy_sample(rddl_program_t* g, const rddl_policy_t* policy, void* const* tensors, int64_t n_envs,
                                  int64_t env_offset, uint64_t seed, uint64_t step, void* cuda_stream) {
  if (!g) return RDDL_ERR_ARG;
  if (!tensors || n_envs <= 0) { g->err = "rddl_policy_sample: bad arguments"; return RDDL_ERR_ARG; }
  std::string* errp = &g->err;
  CK(cudaSetDevice(g->device));
  VmPolicy pol;
  int rc = make_policy(g, policy, &pol);
  if (rc) return rc;
  return policy_prepare(g, &pol, tensors, n_envs, env_offset, seed, step, (cudaStream_t)cuda_stream, true);
}

extern "C" int rddl_rollout_policy(rddl_program_t* g, const rddl_policy_t* policy, void* const* tensors, uint32_t* status,
                                   int64_t n_envs, int64_t env_offset, uint64_t seed, uint64_t step0, int32_t n_steps,
                                   double gamma, double* returns, double* rewards, int32_t* state_in_next,
                                   void* cuda_stream) {
  if (!g) return RDDL_ERR_ARG;
  if (n_steps < 1 || !returns || !state_in_next || !tensors || n_envs <= 0) { g->err = "rddl_rollout_policy: bad arguments"; return RDDL_ERR_ARG; }
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  std::string* errp = &g->err;
  CK(cudaSetDevice(g->device));
  VmPolicy pol;
  int rc = make_policy(g, policy, &pol);
  if (rc) return rc;
  bool direct = false;
  for (int i = 0; i < pol.n; ++i) direct |= g->direct_read[(size_t)pol.e[i].tensor] != 0;
  // fused path: one single-tile plane launch runs every step with the state on chip and the action
  // planes drawn in registers -- nothing but the final state and the returns touches HBM
  int n_step_launches = 0;
  const PlaneProg* G = nullptr;
  for (size_t i = 0; i < g->launches.size(); ++i) {
    if (g->launches[i].plan != RDDL_PLAN_STEP || g->launches[i].kind == 2) continue;
    ++n_step_launches;
    if (g->launches[i].kind == 1) G = &g->plane_progs[i];
  }
  if (n_step_launches == 1 && G && G->rollout_ok && !direct) {
    RolloutArgs ro{n_steps, gamma, returns, rewards, nullptr};
    *state_in_next = 1;
    return run_plan(g, RDDL_PLAN_STEP, tensors, nullptr, status, n_envs, env_offset, seed, step0, stream, &ro, &pol);
  }
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
    if ((rc = policy_prepare(g, &pol, cur.data(), n_envs, env_offset, seed, step0 + t, stream, false))) return rc;
    if ((rc = run_plan(g, RDDL_PLAN_STEP, cur.data(), nullptr, status, n_envs, env_offset, seed, step0 + t, stream, nullptr, &pol)))
      return rc;
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

extern "C" int rddl_action_count(rddl_program_t* g, void* const* tensors, const int32_t* action_tensors,
                                 const double* default_values, int32_t n_actions, int32_t* counts, int64_t n_envs,
                                 void* cuda_stream) {
  if (!g) return RDDL_ERR_ARG;
  if (!tensors || !counts || n_envs <= 0 || n_actions < 0 || n_actions > 16 || (n_actions && (!action_tensors || !default_values))) {
    g->err = "rddl_action_count: bad arguments (at most 16 action tensors)";
    return RDDL_ERR_ARG;
  }
  std::string* errp = &g->err;
  CK(cudaSetDevice(g->device));
  ActionCountArgs a;
  memset(&a, 0, sizeof a);
  a.n = n_actions;
  for (int i = 0; i < n_actions; ++i) {
    const int t = action_tensors[i];
    if (t < 0 || (size_t)t >= g->tensors.size() || !tensors[t]) {
      g->err = "rddl_action_count: bad action tensor";
      return RDDL_ERR_ARG;
    }
    a.ptr[i] = tensors[t];
    a.nelem[i] = g->tensors[t].nelem;
    a.dtype[i] = (uint8_t)g->tensors[t].dtype;
    const double d = default_values[i];
    if (a.dtype[i] == DT_REAL) memcpy(&a.dflt[i], &d, 8);
    else a.dflt[i] = (a.dtype[i] == DT_BOOL || a.dtype[i] == DT_PACKED) ? (uint64_t)(d != 0.0) : (uint64_t)(int64_t)d;
  }
  rddl_action_count_kernel<<<(unsigned)((n_envs * 32 + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(a, counts, n_envs);
  CK(cudaGetLastError());
  ++g->launches_issued;
  return RDDL_OK;
}

// Local part of the return statistics of BaseAgent.evaluate (pyRDDLGym/core/policy.py:120-126): one
// CTA, fixed summation order (deterministic): out = {sum, sum of squares, count, min, max}.
__global__ void __launch_bounds__(1024) rddl_return_stats_kernel(const double* __restrict__ r, int64_t n, double* out) {
  double s = 0.0, q = 0.0, mn = INFINITY, mx = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const double v = r[i];
    s += v; q += v * v; mn = fmin(mn, v); mx = fmax(mx, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_down_sync(0xffffffffu, s, o);
    q += __shfl_down_sync(0xffffffffu, q, o);
    mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
  }
  __shared__ double sh[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = q; sh[2][warp] = mn; sh[3][warp] = mx; }
  __syncthreads();
  if (warp == 0) {
    s = sh[0][lane]; q = sh[1][lane]; mn = sh[2][lane]; mx = sh[3][lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_down_sync(0xffffffffu, s, o);
      q += __shfl_down_sync(0xffffffffu, q, o);
      mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, o));
      mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) { out[0] = s; out[1] = q; out[2] = (double)n; out[3] = mn; out[4] = mx; }
  }
}

extern "C" int rddl_return_stats(const double* returns, int64_t n, double* out5, void* cuda_stream) {
  if (!returns || !out5 || n < 0) return RDDL_ERR_ARG;
  rddl_return_stats_kernel<<<1, 1024, 0, (cudaStream_t)cuda_stream>>>(returns, n, out5);
  return cudaGetLastError() == cudaSuccess ? RDDL_OK : RDDL_ERR_CUDA;
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

extern "C" int rddl_program_option(rddl_program_t* g, int what, int64_t value) {
  if (!g) return RDDL_ERR_ARG;
  if (what == 0) { g->specialize = value != 0; return RDDL_OK; }
  if (what == 1) { g->level_hot = value; return RDDL_OK; }
  if (what == 2) { g->contract_tc = value != 0; return RDDL_OK; }
  g->err = "rddl_program_option: unknown option";
  return RDDL_ERR_ARG;
}

extern "C" const char* rddl_jit_error(const rddl_program_t* g) { return g ? g->jit_err.c_str() : ""; }

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
