// Bit-plane interpreter: the fast path for launches that pyrddlgym_b200/plane_lowering.py
// proved to be boolean-grid work (every value of the launch is a function of a few bits per
// cell). Bool planes live in HBM as one byte per cell (the torch.bool / NumPy bool layout of
// the reference boundary, pyRDDLGym/core/compiler/initializer.py:16-23); inside the kernel a
// plane register is a packed 32-cell word, so one LOP3 evaluates a logical node for 32 cells.
//
// Tile: 256 threads x 4 words x 32 cells = 32768 cells of consecutive (env, cell) index space.
// Thread t owns words w = j*256 + t (j = 0..3): each global access of a warp covers 1 KiB of
// contiguous bytes (32 B per thread), i.e. fully coalesced 128-bit vector loads/stores.
// Ops (one per op-table instruction, decoded once per thread for its 128 cells):
//   PLD/PST   bytes <-> bits          PLDC    pre-packed non-fluent plane
//   PLUT      m-input truth table     PSHARE  stage a plane (+ halo rows) in shared memory
//   PCOUNT    3x3 stencil count as 4 bit-sliced planes (carry-save adders)
//   PBERN     Bernoulli(table[idx planes]) by bit-serial comparison against Philox words
//   PRED      per-env popcount reductions (reward sums)
#pragma once
#include "philox.cuh"
#include "rddl_plan.h"

#define PLANE_MAX_REGS 32
#define PLANE_SMEM_WORDS 3072   // 1024 tile words + up to 1024 halo words on each side
#define PLANE_BERN_LEVELS 8

struct PlaneFin {
  int m;
  uint64_t regs, truth, bad, always;
  const uint64_t* thr;
  const uint64_t* vals;
  const uint64_t* lev;
};

__device__ __forceinline__ PlaneFin plane_fin(const uint64_t* c) {
  PlaneFin f;
  f.m = (int)c[0];
  f.regs = c[1]; f.truth = c[2]; f.bad = c[3]; f.always = c[4];
  f.thr = c + 5;
  f.vals = c + 5 + (1 << f.m);
  f.lev = c + 5 + 2 * (1 << f.m);
  return f;
}

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

__device__ __forceinline__ uint32_t bitmask(uint64_t tt, int v) { return 0u - (uint32_t)((tt >> v) & 1ull); }

// value of the truth table tt (2^M bits) at the cell-wise index formed by planes x[0..M)
template <int M>
__device__ __forceinline__ uint32_t plane_tt(uint64_t tt, const uint32_t* x) {
  if (M == 0) return bitmask(tt, 0);
  uint32_t v[M > 0 ? (1 << (M - 1)) : 1];
#pragma unroll
  for (int i = 0; i < (1 << (M - 1)); ++i)
    v[i] = (x[0] & bitmask(tt, 2 * i + 1)) | (~x[0] & bitmask(tt, 2 * i));
#pragma unroll
  for (int k = 1; k < M; ++k) {
#pragma unroll
    for (int i = 0; i < (1 << (M - 1 - k)); ++i) v[i] = (x[k] & v[2 * i + 1]) | (~x[k] & v[2 * i]);
  }
  return v[0];
}

__device__ __forceinline__ uint32_t plane_tt_dyn(int m, uint64_t tt, const uint32_t* x) {
  switch (m) {
    case 0: return plane_tt<0>(tt, x);
    case 1: return plane_tt<1>(tt, x);
    case 2: return plane_tt<2>(tt, x);
    case 3: return plane_tt<3>(tt, x);
    case 4: return plane_tt<4>(tt, x);
    default: return plane_tt<5>(tt, x);
  }
}

template <int NPR>
__global__ void __launch_bounds__(256) rddl_plane_interp(const __grid_constant__ VmArgs A) {
  const RddlLaunch& L = A.L;
  const int tid = threadIdx.x;
  const int H = L.rank == 2 ? (int)L.extent[0] : 1;
  const int W32 = (int)((L.rank == 2 ? L.extent[1] : L.extent[0]) >> 5);
  const int wpe = (int)(L.E >> 5);              // words per env
  const int envs_per_tile = L.G, tile_words = L.ipt, chunks = L.chunks;
  const int chunk = blockIdx.x % chunks;
  const int64_t envgrp = blockIdx.x / chunks;

  __shared__ uint32_t S[2][PLANE_SMEM_WORDS];
  __shared__ int32_t cnt[32][RDDL_MAX_RED][8];
  for (int i = tid; i < 32 * RDDL_MAX_RED * 8; i += 256) (&cnt[0][0][0])[i] = 0;

  uint32_t P[NPR][4];
  int64_t env[4];
  int wi[4], el[4];
  bool valid[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int w = j * 256 + tid;
    bool ok = w < tile_words;
    if (chunks == 1) {
      el[j] = w / wpe;
      wi[j] = w - el[j] * wpe;
      env[j] = envgrp * envs_per_tile + el[j];
    } else {
      el[j] = 0;
      wi[j] = chunk * tile_words + w;
      env[j] = envgrp;
      ok = ok && wi[j] < wpe;
    }
    valid[j] = ok && env[j] < A.n_envs;
  }
  __syncthreads();

  const RddlInsn* prog = A.insns + L.pc_begin;
  const int npc = L.pc_end - L.pc_begin;
  for (int pc = 0; pc < npc; ++pc) {
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(prog + pc));
    const uint32_t op = raw.x & 0xff;
    const uint32_t dst = raw.x >> 16;
    const uint32_t ra = raw.y & 0xffff, rb = raw.y >> 16;
    const uint32_t imm = raw.w;
    switch (op) {
      case OP_PLD: {
        const uint8_t* base = reinterpret_cast<const uint8_t*>(A.tensors[imm]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t word = 0;
          if (valid[j]) {
            const uint4* p = reinterpret_cast<const uint4*>(base + ((env[j] * wpe + wi[j]) << 5));
            word = plane_pack(__ldcs(p), __ldcs(p + 1));
          }
          P[dst][j] = word;
        }
        break;
      }
      case OP_PLDC: {
#pragma unroll
        for (int j = 0; j < 4; ++j) P[dst][j] = valid[j] ? (uint32_t)__ldg(A.pool + imm + wi[j]) : 0u;
        break;
      }
      case OP_PST: {
        uint8_t* base = reinterpret_cast<uint8_t*>(A.tensors[imm]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (valid[j]) {
            uint4 lo, hi;
            plane_unpack(P[ra][j], lo, hi);
            uint4* p = reinterpret_cast<uint4*>(base + ((env[j] * wpe + wi[j]) << 5));
            __stcs(p, lo);
            __stcs(p + 1, hi);
          }
        }
        break;
      }
      case OP_PCONST: {
#pragma unroll
        for (int j = 0; j < 4; ++j) P[dst][j] = imm ? 0xffffffffu : 0u;
        break;
      }
      case OP_PLUT: {
        const PlaneFin f = plane_fin(A.consts + imm);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t x[5];
#pragma unroll
          for (int i = 0; i < 5; ++i) x[i] = i < f.m ? P[(f.regs >> (8 * i)) & 0xff][j] : 0u;
          P[dst][j] = plane_tt_dyn(f.m, f.truth, x);
        }
        break;
      }
      case OP_PSHARE: {
        uint32_t* s = S[rb];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int w = j * 256 + tid;
          if (w < tile_words) s[W32 + w] = valid[j] ? P[ra][j] : 0u;
        }
        if (chunks > 1) {
          // halo rows above / below the tile, packed straight from the source tensor
          const uint8_t* base = reinterpret_cast<const uint8_t*>(A.tensors[imm]);
          const int64_t e = envgrp;
          for (int i = tid; i < 2 * W32; i += 256) {
            const bool top = i < W32;
            const int wsrc = top ? chunk * tile_words - W32 + i : (chunk + 1) * tile_words + (i - W32);
            uint32_t word = 0;
            if (wsrc >= 0 && wsrc < wpe && e < A.n_envs) {
              const uint4* p = reinterpret_cast<const uint4*>(base + ((e * wpe + wsrc) << 5));
              word = plane_pack(__ldg(p), __ldg(p + 1));
            }
            s[top ? i : W32 + tile_words + (i - W32)] = word;
          }
        }
        __syncthreads();
        break;
      }
      case OP_PCOUNT: {
        const uint32_t* s = S[rb];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t k0 = 0, k1 = 0, k2 = 0, k3 = 0;
          if (valid[j]) {
            const int x = wi[j] / W32, c = wi[j] - x * W32;
            const int pos = W32 + j * 256 + tid;
            uint32_t sr[3], cr[3];   // per row: sum and carry of the (west, centre, east) planes
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const bool rowok = (d == 0) ? x > 0 : (d == 2 ? x < H - 1 : true);
              uint32_t Cw = 0, Lw = 0, Rw = 0;
              if (rowok) {
                const int q = pos + (d - 1) * W32;
                Cw = s[q];
                if (c > 0) Lw = s[q - 1];
                if (c < W32 - 1) Rw = s[q + 1];
              }
              const uint32_t m3 = imm >> (d * 3);
              const uint32_t west = (m3 & 1u) ? ((Cw << 1) | (Lw >> 31)) : 0u;
              const uint32_t mid = (m3 & 2u) ? Cw : 0u;
              const uint32_t east = (m3 & 4u) ? ((Cw >> 1) | (Rw << 31)) : 0u;
              sr[d] = west ^ mid ^ east;
              cr[d] = (west & mid) | (east & (west ^ mid));
            }
            k0 = sr[0] ^ sr[1] ^ sr[2];
            const uint32_t t1 = (sr[0] & sr[1]) | (sr[2] & (sr[0] ^ sr[1]));   // weight 2
            const uint32_t u1 = cr[0] ^ cr[1] ^ cr[2];                         // weight 2
            const uint32_t u2 = (cr[0] & cr[1]) | (cr[2] & (cr[0] ^ cr[1]));   // weight 4
            k1 = t1 ^ u1;
            const uint32_t v = t1 & u1;                                        // weight 4
            k2 = u2 ^ v;
            k3 = u2 & v;
          }
          P[dst][j] = k0; P[dst + 1][j] = k1; P[dst + 2][j] = k2; P[dst + 3][j] = k3;
        }
        break;
      }
      case OP_PBERN: {
        const PlaneFin f = plane_fin(A.consts + imm);
        const double* inj = A.inj[ra];
        const uint32_t node = (uint32_t)A.variates[ra].node & 0xffffu;
        const uint2 key = make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32));
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          uint32_t res = 0;
          if (valid[j]) {
            uint32_t x[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) x[i] = i < f.m ? P[(f.regs >> (8 * i)) & 0xff][j] : 0u;
            if (f.bad) {
              if (plane_tt_dyn(f.m, f.bad, x)) atomicOr(A.status + env[j], 1u);
            }
            if (inj) {
              // parity mode: the reference's U <= p on the same float64 values
              const double* u = inj + env[j] * L.E + ((int64_t)wi[j] << 5);
              for (int b = 0; b < 32; ++b) {
                int v = 0;
                for (int i = 0; i < f.m; ++i) v |= ((x[i] >> b) & 1u) << i;
                res |= (uint32_t)(u[b] <= __longlong_as_double((long long)f.vals[v])) << b;
              }
            } else {
              const uint32_t g = (uint32_t)wi[j];
              const uint32_t e32 = (uint32_t)(A.env_offset + env[j]);
              const uint32_t always = plane_tt_dyn(f.m, f.always, x);
              res = always;
              uint32_t und = ~always;
#pragma unroll
              for (int b = 0; b < PLANE_BERN_LEVELS / 4; ++b) {
                const uint4 r4 = rddl_philox4x32_10(
                    make_uint4(g, e32, (uint32_t)A.step, node | ((0x8000u + b) << 16)), key);
                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint32_t pj = plane_tt_dyn(f.m, f.lev[4 * b + q], x);
                  res |= und & ~rw[q] & pj;
                  und &= ~(rw[q] ^ pj);
                }
              }
              // cells still tied after the first levels: finish each with one 64-bit draw
              while (und) {
                const int b = __ffs(und) - 1;
                und &= und - 1;
                int v = 0;
                for (int i = 0; i < f.m; ++i) v |= ((x[i] >> b) & 1u) << i;
                const uint4 r4 = rddl_philox4x32_10(
                    make_uint4((g << 5) | b, e32, (uint32_t)A.step, node | ((0xC000u + (g >> 27)) << 16)), key);
                const uint64_t u64 = ((uint64_t)r4.x << 32) | r4.y;
                if (u64 < (f.thr[v] << PLANE_BERN_LEVELS)) res |= 1u << b;
              }
            }
          }
          P[dst][j] = res;
        }
        break;
      }
      case OP_PRED: {
        const PlaneFin f = plane_fin(A.consts + imm);
        const bool warp_uniform_env = (wpe & 31) == 0;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          uint32_t x[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) x[i] = (valid[j] && i < f.m) ? P[(f.regs >> (8 * i)) & 0xff][j] : 0u;
          for (int v = 0; v < (1 << f.m); ++v) {
            if (f.vals[v] == 0ull || f.vals[v] == 0x8000000000000000ull) continue;   // +-0 contributes nothing
            uint32_t mt = valid[j] ? 0xffffffffu : 0u;
            for (int i = 0; i < f.m; ++i) mt &= ((v >> i) & 1) ? x[i] : ~x[i];
            int c = __popc(mt);
            if (warp_uniform_env) {
              c = __reduce_add_sync(0xffffffffu, c);
              if ((tid & 31) == 0 && c) atomicAdd(&cnt[el[j]][rb][v], c);
            } else if (c) {
              atomicAdd(&cnt[el[j]][rb][v], c);
            }
          }
        }
        break;
      }
      default: break;
    }
  }

  if (L.nred > 0) {
    __syncthreads();
    for (int i = tid; i < envs_per_tile * L.nred; i += 256) {
      const int e = i / L.nred, jr = i - e * L.nred;
      const int64_t ge = envgrp * envs_per_tile + e;
      if (ge >= A.n_envs) continue;
      // locate the PRED instruction of reduction jr to read its value table
      const uint64_t* fc = nullptr;
      for (int pc = 0; pc < npc; ++pc) {
        const RddlInsn in = prog[pc];
        if (in.op == OP_PRED && in.b == jr) fc = A.consts + in.imm;
      }
      const PlaneFin f = plane_fin(fc);
      uint64_t out;
      if (L.red_op[jr] == RED_SUM_F) {
        double acc = 0.0;
        for (int v = 0; v < (1 << f.m); ++v)
          acc = __dadd_rn(acc, __dmul_rn(__longlong_as_double((long long)f.vals[v]), (double)cnt[e][jr][v]));
        out = (uint64_t)__double_as_longlong(acc);
      } else {
        int64_t acc = 0;
        for (int v = 0; v < (1 << f.m); ++v) acc += (int64_t)f.vals[v] * (int64_t)cnt[e][jr][v];
        out = (uint64_t)acc;
      }
      A.red[A.red_off[L.red_slot[jr]] * A.n_envs + ge * chunks + chunk] = out;
    }
  }
}
