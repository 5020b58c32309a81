// Run-time specialisation of the kernels (host side).
//
// The kernel sources are written once and built twice: ahead of time (nvcc, rddl_vm.cu) as
// interpreters whose program arrives at run time, and -- here -- per created program with the op
// table emitted as compile-time constants and the same files recompiled for sm_100a by NVRTC:
//   * rddl_plane.cuh + "plane_static_data.inc" (the pre-decoded PlaneProg of one plane launch, its
//     tile geometry, the tensor dtypes, the fused scalar programs): the op sequence unrolls, the
//     dispatch switches fold, truth-table masks turn into LOP3 immediates, the plane registers
//     into machine registers  -> rddl_plane_static, one module per (launch, geometry);
//   * rddl_level_static.cuh + "level_static_data.inc" (every instruction word, the pools, the
//     launches): each scalar launch becomes straight-line code over machine registers with real
//     loops for its dense / CSR loops  -> rddl_level_static_<launch>, one module per program.
// Same source, same Philox counters, no FMA contraction in either build: results are bit-identical
// to the interpreters' (tests/test_cuda_parity.py::test_specialised_*).
//
// NVRTC and the driver entry points are bound lazily (dlopen / cudaGetDriverEntryPoint): the
// library keeps loading on machines without a driver, and when NVRTC is missing the interpreter
// kernels run instead (rddl_program_info(prog, 5 | 6) report specialised / fallen-back launches).
// Compiled cubins are cached on disk, keyed by a hash of every source byte, the generated
// constants, the options and the NVRTC version.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#define RDDL_FUSE_TE 4      // envs per CTA of the fused env-tile scalar kernels (rddl_level_static.cuh)

static std::string jit_fmt(const char* f, long long v) {
  char b[64];
  snprintf(b, sizeof b, f, v);
  return b;
}

static void jit_launch_init(std::string& s, const RddlLaunch& L) {
  s += "{" + jit_fmt("%lld,", L.plan) + jit_fmt("%lld,{", L.rank);
  for (int i = 0; i < 8; ++i) s += jit_fmt("%lldll,", (long long)L.extent[i]);
  s += "}," + jit_fmt("%lldll,", (long long)L.E) + jit_fmt("%lld,", L.pc_begin) + jit_fmt("%lld,", L.pc_end) +
       jit_fmt("%lld,", L.nred) + jit_fmt("%lld,", L.G) + jit_fmt("%lld,", L.ipt) + jit_fmt("%lld,{", L.chunks);
  for (int i = 0; i < 8; ++i) s += jit_fmt("%lld,", L.red_op[i]);
  s += "},{";
  for (int i = 0; i < 8; ++i) s += jit_fmt("%lld,", L.red_slot[i]);
  s += "}," + jit_fmt("%lld,", L.nregs) + jit_fmt("%lld}", L.kind);
}

// The constants of one (plane launch, words per thread) pair, as C++ source.
// Host copies of the op-table pools the fused scalar epilogue of a plane launch executes from.
struct JitPools {
  const RddlInsn* insns; int64_t n_insns;
  const uint64_t* consts; int64_t n_consts;
  const RddlAddr* addrs; int64_t n_addrs;
  const RddlRedSlot* redslots; int64_t n_redslots;
  const int64_t* red_off;
  const RddlCsr* csrs; int64_t n_csrs;
  const uint8_t* may_policy;      // [RDDL_MAX_TENSORS]: tensors a device-side policy may drive (action fluents)
};

// The scalar programs fused into a plane launch's epilogue (reward / termination from the
// reductions) as constants: instruction words, const pool, address descriptors, reduction slots.
// Only straight-line programs qualify (no loop ops); otherwise the epilogue calls the interpreter.
static bool jit_fused_static(const PlaneProg& G, const JitPools& P, std::string& s) {
  if (G.nfused <= 0 || P.n_consts > 4096 || P.n_addrs > 512 || P.n_redslots > 64) return false;
  int lo = 1 << 30, hi = 0;
  for (int f = 0; f < G.nfused; ++f) {
    const RddlLaunch& F = G.fused[f];
    if (F.rank != 0 || F.nred != 0 || F.pc_end - F.pc_begin > 256) return false;
    for (int pc = F.pc_begin; pc < F.pc_end; ++pc) {
      const int op = P.insns[pc].op;
      if (op == OP_LOOP || op == OP_ENDLOOP || op == OP_CSRLOOP || op == OP_CSREND || op == OP_SUMLOOP ||
          op == OP_REDOUT || op == OP_AXIS)
        return false;
    }
    if (F.pc_begin < lo) lo = F.pc_begin;
    if (F.pc_end > hi) hi = F.pc_end;
  }
  if (hi - lo > 512) return false;
  s += jit_fmt("#define PS_INSN_BASE %lld\n", lo);
  s += "__device__ constexpr uint32_t kInsnWords[][4] = {";
  for (int pc = lo; pc < hi; ++pc) {
    uint32_t w[4];
    memcpy(w, &P.insns[pc], 16);
    s += "{" + jit_fmt("0x%llxu,", w[0]) + jit_fmt("0x%llxu,", w[1]) + jit_fmt("0x%llxu,", w[2]) + jit_fmt("0x%llxu},", w[3]);
  }
  s += "};\n__device__ constexpr uint64_t kConsts[] = {";
  for (int64_t i = 0; i < P.n_consts; ++i) s += jit_fmt("0x%llxull,", (long long)P.consts[i]);
  s += "0};\n__device__ constexpr RddlAddr kAddrs[] = {";
  for (int64_t i = 0; i < P.n_addrs; ++i) {
    const RddlAddr& a = P.addrs[i];
    s += "{" + jit_fmt("%lld,", a.tensor) + jit_fmt("%lld,", a.naxis) + jit_fmt("%lldll,", (long long)a.env_stride) +
         jit_fmt("%lldll,{", (long long)a.c0);
    for (int k = 0; k < 8; ++k) s += jit_fmt("%lld,", a.axis[k]);
    s += "},{";
    for (int k = 0; k < 8; ++k) s += jit_fmt("%lldll,", (long long)a.coef[k]);
    s += "}," + jit_fmt("%lld,", a.nreg) + jit_fmt("%lld,{", a.pad);
    for (int k = 0; k < 4; ++k) s += jit_fmt("%lld,", a.reg[k]);
    s += "},{";
    for (int k = 0; k < 4; ++k) s += jit_fmt("%lldll,", (long long)a.rstride[k]);
    s += "},{";
    for (int k = 0; k < 4; ++k) s += jit_fmt("%lldll,", (long long)a.rextent[k]);
    s += "}},\n";
  }
  s += "{}};\n__device__ constexpr RddlRedSlot kRedSlots[] = {";
  for (int64_t i = 0; i < P.n_redslots; ++i) s += "{" + jit_fmt("%lld,", P.redslots[i].op) + jit_fmt("%lld},", P.redslots[i].chunks);
  s += "{}};\n__device__ constexpr int64_t kRedOff[] = {";
  for (int64_t i = 0; i < P.n_redslots; ++i) s += jit_fmt("%lldll,", (long long)P.red_off[i]);
  s += "0};\n";
  return true;
}

// The constants of the specialised scalar kernels of a whole program ("level_static_data.inc"): every
// instruction word, the pools, the launches, and the list of scalar launches (PS_LEVELS) that get a
// kernel each (rddl_level_static_<launch index>, rddl_level_static.cuh). Empty string: program too
// large to embed (the interpreter kernels run).
static std::string level_static_data(const JitPools& P, const RddlLaunch* launches, int n_launches,
                                     const uint8_t* dtype) {
  if (P.n_insns > 4096 || P.n_consts > 8192 || P.n_addrs > 1024 || P.n_csrs > 64 || P.n_redslots > 256) return "";
  std::string s, list;
  // a launch's program is one template recursion per instruction: very long programs stay interpreted
  int64_t total = 0;
  for (int i = 0; i < n_launches; ++i) {
    const int len = launches[i].pc_end - launches[i].pc_begin;
    if (launches[i].kind != 0 || len > 200 || total + len > 2000) continue;
    total += len;
    list += jit_fmt(" X(%lld)", i);
  }
  if (list.empty()) return "";
  s += "#define PS_LEVELS(X)" + list + "\n";
  // runs of consecutive scalar launches of the step plan over small scopes: one fused env-tile kernel
  // each (rddl_level_static.cuh: RDDL_FUSED_KERNEL); RDDL_B200_FUSE=0 disables
  {
    std::string fused;
    const char* fe = getenv("RDDL_B200_FUSE");
    const bool enabled = !fe || atoi(fe) != 0;
    for (int i = 0; enabled && i < n_launches;) {
      auto ok = [&](int k) {
        const RddlLaunch& L = launches[k];
        return L.kind == 0 && L.plan == 0 && L.rank <= 1 && L.E <= 1408 && L.pc_end - L.pc_begin <= 200;
      };
      if (!ok(i)) { ++i; continue; }
      int j = i;
      while (j + 1 < n_launches && ok(j + 1)) ++j;
      // (worth a kernel of its own only where a launch of the run gathers a per-env vector through a
      // sparse sum -- the staged SpMM; plain runs keep their per-launch kernels and compile times)
      bool gathers = false;
      for (int k = i; k <= j; ++k) {
        for (int pc = launches[k].pc_begin; pc < launches[k].pc_end; ++pc) {
          const RddlInsn& in = P.insns[pc];
          if (in.op != OP_SUMLOOP) continue;
          if ((in.b & 1u) && (in.b & 4u) && in.imm < (uint32_t)P.n_addrs) {
            const RddlAddr& ad = P.addrs[in.imm];
            if (ad.tensor >= 0 && ad.env_stride > 0 && ad.env_stride * RDDL_FUSE_TE <= 5632 && dtype[ad.tensor] == DT_REAL && ad.nreg == 0)
              gathers = true;
          }
          ++pc;
        }
      }
      if (j > i && gathers) fused += jit_fmt(" X(%lld,", i) + jit_fmt("%lld)", j - i + 1);
      i = j + 1;
    }
    if (!fused.empty()) {
      s += "#define PS_FUSED(X)" + fused + "\n" + jit_fmt("#define FUSE_TE %lld\n", RDDL_FUSE_TE);
      // resident CTAs per SM the fused kernels are compiled for (4 = 64 registers; experiments: RDDL_FUSE_MINB)
      const char* mb = getenv("RDDL_FUSE_MINB");
      s += jit_fmt("#define FUSE_MINB %lld\n", mb && atoi(mb) > 0 ? atoi(mb) : 4);
    }
  }
  s += "__device__ constexpr uint32_t kInsnWords[][4] = {";
  for (int64_t pc = 0; pc < P.n_insns; ++pc) {
    uint32_t w[4];
    memcpy(w, &P.insns[pc], 16);
    s += "{" + jit_fmt("0x%llxu,", w[0]) + jit_fmt("0x%llxu,", w[1]) + jit_fmt("0x%llxu,", w[2]) + jit_fmt("0x%llxu},", w[3]);
    if ((pc & 7) == 7) s += "\n";
  }
  s += "{0,0,0,0}};\n__device__ constexpr uint64_t kConsts[] = {";
  for (int64_t i = 0; i < P.n_consts; ++i) s += jit_fmt("0x%llxull,", (long long)P.consts[i]) + ((i & 7) == 7 ? "\n" : "");
  s += "0};\n__device__ constexpr RddlAddr kAddrs[] = {";
  for (int64_t i = 0; i < P.n_addrs; ++i) {
    const RddlAddr& a = P.addrs[i];
    s += "{" + jit_fmt("%lld,", a.tensor) + jit_fmt("%lld,", a.naxis) + jit_fmt("%lldll,", (long long)a.env_stride) +
         jit_fmt("%lldll,{", (long long)a.c0);
    for (int k = 0; k < 8; ++k) s += jit_fmt("%lld,", a.axis[k]);
    s += "},{";
    for (int k = 0; k < 8; ++k) s += jit_fmt("%lldll,", (long long)a.coef[k]);
    s += "}," + jit_fmt("%lld,", a.nreg) + jit_fmt("%lld,{", a.pad);
    for (int k = 0; k < 4; ++k) s += jit_fmt("%lld,", a.reg[k]);
    s += "},{";
    for (int k = 0; k < 4; ++k) s += jit_fmt("%lldll,", (long long)a.rstride[k]);
    s += "},{";
    for (int k = 0; k < 4; ++k) s += jit_fmt("%lldll,", (long long)a.rextent[k]);
    s += "}},\n";
  }
  s += "{}};\n__device__ constexpr RddlRedSlot kRedSlots[] = {";
  for (int64_t i = 0; i < P.n_redslots; ++i) s += "{" + jit_fmt("%lld,", P.redslots[i].op) + jit_fmt("%lld},", P.redslots[i].chunks);
  s += "{}};\n__device__ constexpr int64_t kRedOff[] = {";
  for (int64_t i = 0; i < P.n_redslots; ++i) s += jit_fmt("%lldll,", (long long)P.red_off[i]);
  s += "0};\n__device__ constexpr RddlCsr kCsrs[] = {";
  for (int64_t i = 0; i < P.n_csrs; ++i) {
    const RddlCsr& c = P.csrs[i];
    s += "{" + jit_fmt("%lldll,", (long long)c.rowptr_off) + jit_fmt("%lld,", c.ncols) + jit_fmt("%lld,{", c.col4_u16);
    for (int k = 0; k < 4; ++k) s += jit_fmt("%lldll,", (long long)c.col_off[k]);
    s += "},{";
    for (int k = 0; k < 4; ++k) s += jit_fmt("%lld,", c.slot[k]);
    s += "}," + jit_fmt("%lldll,", (long long)c.row_c0) + jit_fmt("%lld,", c.row_naxis) + jit_fmt("%lld,{", c.pad2);
    for (int k = 0; k < 8; ++k) s += jit_fmt("%lld,", c.row_axis[k]);
    s += "},{";
    for (int k = 0; k < 8; ++k) s += jit_fmt("%lldll,", (long long)c.row_coef[k]);
    s += "}," + jit_fmt("%lldll,", (long long)c.rowptr4_off) + jit_fmt("%lldll,", (long long)c.col4_off) +
         jit_fmt("%lldll", (long long)c.sentinel4) + "},\n";
  }
  s += "{}};\n__device__ constexpr RddlLaunch kLaunches[] = {";
  for (int i = 0; i < n_launches; ++i) { jit_launch_init(s, launches[i]); s += ",\n"; }
  s += "};\n__device__ constexpr uint8_t kDT[RDDL_MAX_TENSORS] = {";
  for (int i = 0; i < RDDL_MAX_TENSORS; ++i) s += jit_fmt("%lld,", dtype[i]);
  s += "};\n__device__ constexpr uint8_t kMayPolicy[RDDL_MAX_TENSORS] = {";
  for (int i = 0; i < RDDL_MAX_TENSORS; ++i) s += jit_fmt("%lld,", P.may_policy ? P.may_policy[i] : 1);
  s += "};\n";
  return s;
}

static std::string plane_static_data(const PlaneProg& G, const RddlLaunch& L, const uint8_t* dtype, int wpt,
                                     int min_blocks, int block, const JitPools& P) {
  std::string s;
  {
    std::string f;
    if (jit_fused_static(G, P, f)) s += "#define PS_FUSED_STATIC 1\n" + f;
    else s += "#define PS_FUSED_STATIC 0\n";
  }
  s += jit_fmt("#define PS_WPT %lld\n", wpt) + jit_fmt("#define PS_NREGS %lld\n", L.nregs > 0 ? L.nregs : 1) +
       jit_fmt("#define PS_NOPS %lld\n", G.nops) + jit_fmt("#define PS_MINB %lld\n", min_blocks) +
       jit_fmt("#define PS_BLOCK %lld\n", block);
  s += "__device__ constexpr PlaneProg kG = {" + jit_fmt("%lld,", G.nops) + jit_fmt("%lld,", G.nshared) +
       jit_fmt("%lld,", G.w32_shift) + jit_fmt("%lld,", G.wpe_shift) + jit_fmt("%lld,", G.nfused) +
       jit_fmt("%lld,", G.fused_nregs) + jit_fmt("%lld,", G.rollout_ok) + jit_fmt("%lld,", G.pad) +
       jit_fmt("%lld,", G.pol_k) + jit_fmt("%lld,", G.ntma) + jit_fmt("%lld,", G.npol) + jit_fmt("%lld,\n{", G.pad2);
  for (int i = 0; i < RDDL_MAX_POLICY; ++i) s += "{" + jit_fmt("0x%llxull,", (long long)G.pol_thr[i][0]) + jit_fmt("0x%llxull},", (long long)G.pol_thr[i][1]);
  s += "},{";
  for (int i = 0; i < RDDL_MAX_POLICY; ++i) s += jit_fmt("%lluu,", G.pol_node[i]);
  s += "},{";
  for (int i = 0; i < RDDL_MAX_POLICY; ++i) s += jit_fmt("%lld,", G.pol_always[i]);
  s += "},\n{";
  jit_launch_init(s, G.fused[0]);
  s += ",\n";
  jit_launch_init(s, G.fused[1]);
  s += "},\n{";
  for (int i = 0; i < RDDL_MAX_RED; ++i) s += jit_fmt("%lld,", G.red_off[i]);
  s += "},{";
  for (int i = 0; i < RDDL_MAX_RED; ++i) s += jit_fmt("%lld,", G.red_m[i]);
  s += "},\n{";
  for (int i = 0; i < G.nops; ++i) {
    const PlaneOp& o = G.ops[i];
    s += "{" + jit_fmt("%lld,", o.op) + jit_fmt("%lld,", o.dst) + jit_fmt("%lld,", o.a) + jit_fmt("%lld,", o.b) +
         jit_fmt("%lld,", o.m) + jit_fmt("%lluu,", o.idx) + jit_fmt("%lld,", o.off) + jit_fmt("%lluu},\n", o.imm);
  }
  s += "},\n{";
  int nw = 0;   // words in use: everything up to the last table
  for (int i = 0; i < PLANE_WORDS; ++i) if (G.words[i]) nw = i + 1;
  for (int i = 0; i < nw; ++i) s += jit_fmt("0x%llxu,", G.words[i]) + ((i & 15) == 15 ? "\n" : "");
  s += "}};\n__device__ constexpr RddlLaunch kL = ";
  jit_launch_init(s, L);
  s += ";\n__device__ constexpr uint8_t kDT[RDDL_MAX_TENSORS] = {";
  for (int i = 0; i < RDDL_MAX_TENSORS; ++i) s += jit_fmt("%lld,", dtype[i]);
  s += "};\n";
  return s;
}

// ---------------------------------------------------------------------------------------
// NVRTC + driver entry points, bound lazily
// ---------------------------------------------------------------------------------------
#include <nvrtc.h>

struct JitApi {
  bool tried = false, ok = false;
  std::string why;
  void* h = nullptr;
  nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
  nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*);
  nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetCUBIN)(nvrtcProgram, char*);
  nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*);
  nvrtcResult (*GetProgramLog)(nvrtcProgram, char*);
  nvrtcResult (*DestroyProgram)(nvrtcProgram*);
  nvrtcResult (*Version)(int*, int*);
  int major = 0, minor = 0;
};

static JitApi& jit_api() {
  static JitApi api;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (api.tried) return api;
  api.tried = true;
  const char* want = getenv("RDDL_B200_NVRTC");
  if (want && !strcmp(want, "none")) { api.why = "disabled (RDDL_B200_NVRTC=none)"; return api; }   // fallback tests
  const char* names[] = {want, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    api.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (api.h) break;
  }
  if (!api.h) { api.why = "libnvrtc.so.12 not found"; return api; }
#define JIT_SYM(field, name)                                               \
  *(void**)(&api.field) = dlsym(api.h, name);                               \
  if (!api.field) { api.why = std::string("missing symbol ") + name; return api; }
  JIT_SYM(CreateProgram, "nvrtcCreateProgram")
  JIT_SYM(CompileProgram, "nvrtcCompileProgram")
  JIT_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
  JIT_SYM(GetCUBIN, "nvrtcGetCUBIN")
  JIT_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
  JIT_SYM(GetProgramLog, "nvrtcGetProgramLog")
  JIT_SYM(DestroyProgram, "nvrtcDestroyProgram")
  JIT_SYM(Version, "nvrtcVersion")
#undef JIT_SYM
  api.Version(&api.major, &api.minor);
  api.ok = true;
  return api;
}

// directory of this shared library: the kernel sources (*.cuh) and the cubin cache live next to it
static std::string jit_lib_dir() {
  Dl_info info;
  if (dladdr((const void*)&jit_lib_dir, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    const size_t k = p.rfind('/');
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
  }
  return ".";
}

static bool jit_read_file(const std::string& path, std::string* out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char buf[1 << 16];
  size_t n;
  out->clear();
  while ((n = fread(buf, 1, sizeof buf, f)) > 0) out->append(buf, n);
  fclose(f);
  return true;
}

static uint64_t jit_hash(uint64_t h, const std::string& s) {
  for (unsigned char c : s) { h ^= c; h *= 0x100000001b3ull; }
  return h;
}

static const char* const kJitHeaders[] = {"rddl_stdint.h", "philox.cuh", "rddl_plan.h", "rddl_samplers.cuh",
                                          "rddl_device.cuh", "rddl_plane.cuh", "rddl_level_static.cuh"};
static const char* const kJitMain = "#define PLANE_STATIC 1\n#include \"rddl_device.cuh\"\n#include \"rddl_plane.cuh\"\n";
static const char* const kJitMainLevel = "#include \"rddl_level_static.cuh\"\n";

// Compiles (or fetches from the disk cache) the specialised kernel for `data`; the cubin comes back in *cubin.
static bool jit_compile(const std::string& data, const char* data_name, const char* main_src, const char* prefix,
                        std::string* cubin, std::string* err, bool* cache_hit);

static bool jit_compile_plane(const std::string& data, std::string* cubin, std::string* err, bool* cache_hit) {
  return jit_compile(data, "plane_static_data.inc", kJitMain, "plane", cubin, err, cache_hit);
}

static bool jit_compile_level(const std::string& data, std::string* cubin, std::string* err, bool* cache_hit) {
  return jit_compile(data, "level_static_data.inc", kJitMainLevel, "level", cubin, err, cache_hit);
}

static bool jit_compile(const std::string& data, const char* data_name, const char* main_src, const char* prefix,
                        std::string* cubin, std::string* err, bool* cache_hit) {
  const std::string dir = jit_lib_dir();
  std::vector<std::string> texts;
  uint64_t h = 0xcbf29ce484222325ull;
  for (const char* name : kJitHeaders) {
    std::string t;
    if (!jit_read_file(dir + "/" + name, &t)) { *err = std::string("kernel source not found: ") + dir + "/" + name; return false; }
    h = jit_hash(h, t);
    texts.push_back(std::move(t));
  }
  h = jit_hash(h, data);
  h = jit_hash(h, main_src);
  JitApi& api = jit_api();
  h = jit_hash(h, jit_fmt("nvrtc %lld.", api.major) + jit_fmt("%lld sm_100a O3 lineinfo default-device fmad=false", api.minor));
  const char* cdir = getenv("RDDL_B200_JIT_CACHE");
  const std::string cache_dir = cdir && *cdir ? std::string(cdir) : dir + "/jit_cache";
  char name[64];
  snprintf(name, sizeof name, "/%s_%016llx.cubin", prefix, (unsigned long long)h);
  const std::string path = cache_dir + name;
  if (cache_hit) *cache_hit = false;
  if (const char* dump = getenv("RDDL_B200_JIT_DUMP")) {
    // offline inspection (nvcc -cubin / cuobjdump -sass of the specialised source): <dir>/<prefix>_<hash>.inc
    snprintf(name, sizeof name, "/%s_%016llx.inc", prefix, (unsigned long long)h);
    if (FILE* f = fopen((std::string(dump) + name).c_str(), "wb")) { fwrite(data.data(), 1, data.size(), f); fclose(f); }
    snprintf(name, sizeof name, "/%s_%016llx.cubin", prefix, (unsigned long long)h);
  }
  if (jit_read_file(path, cubin) && cubin->size() > 64) {
    if (cache_hit) *cache_hit = true;
    return true;
  }
  if (!api.ok) { *err = "NVRTC unavailable: " + api.why; return false; }
  std::vector<const char*> hdr_src, hdr_name;
  for (size_t i = 0; i < texts.size(); ++i) { hdr_src.push_back(texts[i].c_str()); hdr_name.push_back(kJitHeaders[i]); }
  hdr_src.push_back(data.c_str());
  hdr_name.push_back(data_name);
  nvrtcProgram prog;
  if (api.CreateProgram(&prog, main_src, "rddl_static.cu", (int)hdr_src.size(), hdr_src.data(), hdr_name.data()) !=
      NVRTC_SUCCESS) {
    *err = "nvrtcCreateProgram failed";
    return false;
  }
  // --fmad=false: no contraction across what the interpreter executes as separate multiply / add
  // instructions (the ahead-of-time build uses the same flag), so both builds round identically
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--generate-line-info",
                        "--device-as-default-execution-space", "--fmad=false"};
  const nvrtcResult rc = api.CompileProgram(prog, 5, opts);
  if (rc != NVRTC_SUCCESS) {
    size_t n = 0;
    api.GetProgramLogSize(prog, &n);
    std::string log(n, 0);
    if (n) api.GetProgramLog(prog, &log[0]);
    *err = "NVRTC compile failed:\n" + log.substr(0, 4000);
    api.DestroyProgram(&prog);
    return false;
  }
  size_t n = 0;
  api.GetCUBINSize(prog, &n);
  cubin->assign(n, 0);
  if (n) api.GetCUBIN(prog, &(*cubin)[0]);
  api.DestroyProgram(&prog);
  if (n < 64) { *err = "NVRTC produced no cubin"; return false; }
  // best-effort cache write (atomic rename; a read-only tree just recompiles next time)
  mkdir(cache_dir.c_str(), 0777);
  const std::string tmp = path + jit_fmt(".tmp%lld", (long long)getpid());
  if (FILE* f = fopen(tmp.c_str(), "wb")) {
    const bool okw = fwrite(cubin->data(), 1, cubin->size(), f) == cubin->size();
    fclose(f);
    if (!okw || rename(tmp.c_str(), path.c_str()) != 0) unlink(tmp.c_str());
  }
  return true;
}

// Driver entry points through the runtime (no link-time dependency on libcuda).
struct JitDriver {
  bool ok = false;
  CUresult (*ModuleLoadData)(CUmodule*, const void*);
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*);
  CUresult (*ModuleUnload)(CUmodule);
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int);
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                           void**, void**);
};

static JitDriver& jit_driver() {
  static JitDriver d;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (d.ok) return d;
  bool ok = true;
  auto get = [&](const char* name, void** fn) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn)
      ok = false;
  };
  get("cuModuleLoadData", (void**)&d.ModuleLoadData);
  get("cuModuleGetFunction", (void**)&d.ModuleGetFunction);
  get("cuModuleUnload", (void**)&d.ModuleUnload);
  get("cuFuncSetAttribute", (void**)&d.FuncSetAttribute);
  get("cuLaunchKernel", (void**)&d.LaunchKernel);
  cudaGetLastError();
  d.ok = ok;
  return d;
}

// One specialised kernel of a program: (launch index, words per thread, tile geometry).
struct JitKernel {
  int launch = -1, wpt = 0, G = 0, ipt = 0, minb = 0;
  uint64_t sig = 0;     // 0: the program as lowered; else hash of its policy-rewritten PlaneProg
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  bool failed = false;
};
