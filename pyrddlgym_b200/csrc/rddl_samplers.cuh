// Native samplers for the RDDL distributions that NumPy draws by rejection / search
// (pyRDDLGym/core/simulator.py:1066-1073 Poisson, :1095-1104 Gamma, :1106-1116 Binomial,
// :1118-1127 NegativeBinomial, :1129-1138 Beta, :1140-1147 Geometric, :1160-1167 Student,
// :1214-1221 ChiSquare). They cannot share base variates with NumPy's generators, so parity is
// distributional (moment / KS tests); each (seed, env, step, node, element) owns a private
// Philox4x32-10 stream: counter = (element, env, step, node | sub << 16), sub = 1, 2, ...
#pragma once
#include "philox.cuh"

enum RddlDist { DIST_POISSON = 0, DIST_GAMMA, DIST_BINOMIAL, DIST_NEGBINOMIAL, DIST_BETA, DIST_GEOMETRIC,
                DIST_STUDENT, DIST_CHISQUARE };

struct RddlStream {
  uint2 key;
  uint32_t c0, c1, c2, node, sub;
  uint4 buf;
  int have;
  __device__ __forceinline__ void init(uint64_t seed, uint64_t env, uint64_t step, uint32_t nd, uint64_t elem) {
    key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    c0 = (uint32_t)elem; c1 = (uint32_t)env; c2 = (uint32_t)step; node = nd & 0xffffu; sub = 0; have = 0;
  }
  __device__ __forceinline__ uint32_t next32() {
    if (have == 0) {
      ++sub;
      buf = rddl_philox4x32_10(make_uint4(c0, c1, c2, node | ((sub & 0x7fffu) << 16)), key);
      have = 4;
    }
    --have;
    const uint32_t v = buf.x;
    buf = make_uint4(buf.y, buf.z, buf.w, 0u);
    return v;
  }
  __device__ __forceinline__ double uniform() {   // [0, 1)
    const uint32_t a = next32(), b = next32();
    return rddl_u53(a, b);
  }
  __device__ __forceinline__ double normal() {
    const double u1 = uniform(), u2 = uniform();
    return sqrt(-2.0 * log(1.0 - u1)) * cos(2.0 * 3.141592653589793 * u2);
  }
};

// Marsaglia & Tsang (2000), shape > 0, unit scale
__device__ double rddl_gamma(RddlStream& s, double shape) {
  double boost = 1.0;
  if (shape < 1.0) {
    boost = pow(1.0 - s.uniform(), 1.0 / shape);
    shape += 1.0;
  }
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int it = 0; it < 1000; ++it) {
    const double x = s.normal();
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = 1.0 - s.uniform();
    if (u < 1.0 - 0.0331 * x * x * x * x) return boost * d * v;
    if (log(u) < 0.5 * x * x + d * (1.0 - v + log(v))) return boost * d * v;
  }
  return boost * d;
}

// Knuth multiplication for small rates, Hoermann's PTRS (1993) above 10
__device__ int64_t rddl_poisson(RddlStream& s, double lam) {
  if (!(lam > 0.0)) return 0;
  if (lam < 10.0) {
    const double L = exp(-lam);
    int64_t k = 0;
    double p = 1.0;
    do {
      ++k;
      p *= s.uniform();
    } while (p > L && k < 100000);
    return k - 1;
  }
  const double slam = sqrt(lam), loglam = log(lam);
  const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
  const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
  for (int it = 0; it < 1000; ++it) {
    const double U = s.uniform() - 0.5, V = 1.0 - s.uniform();
    const double us = 0.5 - fabs(U);
    const double kf = floor((2.0 * a / us + b) * U + lam + 0.43);
    if (us >= 0.07 && V <= vr) return (int64_t)kf;
    if (kf < 0.0 || (us < 0.013 && V > us)) continue;
    if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + kf * loglam - lgamma(kf + 1.0)) return (int64_t)kf;
  }
  return (int64_t)lam;
}

// inversion for n*min(p,q) < 30, Hoermann's BTRS (1993) above
__device__ int64_t rddl_binomial(RddlStream& s, int64_t n, double p) {
  if (n <= 0 || !(p > 0.0)) return 0;
  if (p >= 1.0) return n;
  const bool flip = p > 0.5;
  if (flip) p = 1.0 - p;
  const double q = 1.0 - p, nd = (double)n;
  int64_t k;
  if (nd * p < 30.0) {
    const double sr = p / q, a = (nd + 1.0) * sr;
    double r = exp(nd * log1p(-p)), u = s.uniform();
    k = 0;
    while (u > r && k < n) {
      u -= r;
      ++k;
      r *= (a / (double)k - sr);
    }
  } else {
    const double spq = sqrt(nd * p * q);
    const double b = 1.15 + 2.53 * spq, a = -0.0873 + 0.0248 * b + 0.01 * p, c = nd * p + 0.5;
    const double vr = 0.92 - 4.2 / b, alpha = (2.83 + 5.1 / b) * spq;
    const double m = floor((nd + 1.0) * p), lpq = log(p / q);
    const double h = lgamma(m + 1.0) + lgamma(nd - m + 1.0);
    k = (int64_t)m;
    for (int it = 0; it < 1000; ++it) {
      const double u = s.uniform() - 0.5;
      double v = 1.0 - s.uniform();
      const double us = 0.5 - fabs(u);
      const double kf = floor((2.0 * a / us + b) * u + c);
      if (kf < 0.0 || kf > nd) continue;
      if (us >= 0.07 && v <= vr) { k = (int64_t)kf; break; }
      v = log(v * alpha / (a / (us * us) + b));
      if (v <= h - lgamma(kf + 1.0) - lgamma(nd - kf + 1.0) + (kf - m) * lpq) { k = (int64_t)kf; break; }
    }
  }
  return flip ? n - k : k;
}

// returns the raw 64-bit register value (int64 for the count distributions, double otherwise)
__device__ __noinline__ uint64_t rddl_sample(int dist, uint64_t pa, uint64_t pb, uint64_t seed, uint64_t env,
                                             uint64_t step, uint32_t node, uint64_t elem) {
  RddlStream s;
  s.init(seed, env, step, node, elem);
  const double a = __longlong_as_double((long long)pa), b = __longlong_as_double((long long)pb);
  switch (dist) {
    case DIST_POISSON: return (uint64_t)rddl_poisson(s, a);
    case DIST_GAMMA: return (uint64_t)__double_as_longlong(b * rddl_gamma(s, a));
    case DIST_BINOMIAL: return (uint64_t)rddl_binomial(s, (int64_t)pa, b);        // count arrives as int64
    case DIST_NEGBINOMIAL: {
      const double y = rddl_gamma(s, a) * ((1.0 - b) / b);
      return (uint64_t)rddl_poisson(s, y);
    }
    case DIST_BETA: {
      const double x = rddl_gamma(s, a), y = rddl_gamma(s, b);
      return (uint64_t)__double_as_longlong(x / (x + y));
    }
    case DIST_GEOMETRIC: {
      if (a >= 1.0) return 1;
      const double k = ceil(log(1.0 - s.uniform()) / log1p(-a));
      return (uint64_t)(int64_t)(k < 1.0 ? 1.0 : k);
    }
    case DIST_STUDENT: {
      const double z = s.normal(), g = rddl_gamma(s, 0.5 * a);
      return (uint64_t)__double_as_longlong(sqrt(0.5 * a) * z / sqrt(g));
    }
    default: return (uint64_t)__double_as_longlong(2.0 * rddl_gamma(s, 0.5 * a));   // ChiSquare
  }
}
