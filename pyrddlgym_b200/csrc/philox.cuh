// Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11)
// counter-based generator for the RDDL distribution samplers. The reference draws from NumPy's
// PCG64 stream in a state-dependent order (pyRDDLGym/core/simulator.py:1043-1256), which cannot
// be reproduced by construction; what is pinned (oracle/philox_np.py restates it) is
//   key     = (seed lo, seed hi)
//   counter = (elem lo, global env id, step, node id | (elem hi << 16))
// so a draw depends only on (seed, env, step, AST node, element): results are invariant to how
// envs are sharded over GPUs.
#pragma once
#include "rddl_stdint.h"

__device__ __forceinline__ uint4 rddl_philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// The same generator with the ten rounds' keys precomputed on the host (key + r * (W0, W1), r = 0..9): read
// from the kernel-parameter constant bank they are operands of the rounds' LOP3s -- no key-schedule
// additions in the instruction stream (18 uniform-datapath adds per block otherwise, ~12 % of a block).
struct RddlPhiloxKeys { uint32_t k[20]; };
__host__ __device__ inline void rddl_philox_round_keys(uint64_t seed, RddlPhiloxKeys* rk) {
  uint32_t kx = (uint32_t)seed, ky = (uint32_t)(seed >> 32);
  for (int i = 0; i < 10; ++i) {
    rk->k[2 * i] = kx;
    rk->k[2 * i + 1] = ky;
    kx += 0x9E3779B9u;
    ky += 0xBB67AE85u;
  }
}
__device__ __forceinline__ uint4 rddl_philox4x32_10(uint4 c, const RddlPhiloxKeys& rk) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ rk.k[2 * i], lo1, hi0 ^ c.w ^ rk.k[2 * i + 1], lo0);
  }
  return c;
}

__device__ __forceinline__ uint4 rddl_philox_draw(uint64_t seed, uint64_t env, uint64_t step, uint32_t node,
                                                  uint64_t elem) {
  uint4 c = make_uint4((uint32_t)elem, (uint32_t)env, (uint32_t)step,
                       (node & 0xffffu) | ((uint32_t)(elem >> 32) << 16));
  return rddl_philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}

__device__ __forceinline__ uint4 rddl_philox_draw(const RddlPhiloxKeys& rk, uint64_t env, uint64_t step, uint32_t node,
                                                  uint64_t elem) {
  uint4 c = make_uint4((uint32_t)elem, (uint32_t)env, (uint32_t)step,
                       (node & 0xffffu) | ((uint32_t)(elem >> 32) << 16));
  return rddl_philox4x32_10(c, rk);
}

// 53-bit uniform in [0, 1)
__device__ __forceinline__ double rddl_u53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
