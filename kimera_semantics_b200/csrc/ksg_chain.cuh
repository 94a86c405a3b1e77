// ksg_chain.cuh - exact evaluation of a long same-sign float32 addition chain  s <- fl(s + a_k)  (s < 0, a_k <= 0, round to nearest
// even) as an associative scan.  This is the per-voxel, per-class log-probability recurrence of `merged` (base.cpp:306-307): the
// voxels next to the camera receive ~92 000 such updates per frame in strict bundle order, which is the critical path of the tile
// kernel today (DESIGN.md sections 7 and 9).  Model, proof sketch and Python reference: tools/exact_float_chain.py.
//
// Inside one binade [2^e', 2^(e'+1)) the running value is M * u with u = 2^g, M an integer in [2^23, 2^24), and
//     fl(s + a) = -(M + q) * u,   q = floor(x) + (frac(x) > 1/2),  x = |a| / u,
// except for exact ties frac(x) = 1/2, where q makes M + q even.  One record is therefore a function  parity(M) -> (q, new parity),
// two entries, and composition of such functions is associative.  A prefix leaves the binade when M + sum(q) >= 2^24; that one
// record is then added as an ordinary float and the grid coarsens.
//
// Everything here is plain integer code marked __host__ __device__: the warp kernel (next round) wraps it in shuffles, and
// csrc/test/chain_host_test.cpp checks it on the CPU against the sequential float loop.  NOT yet included by ksg_kernels.cuh.
#pragma once
#include <stdint.h>
#include <string.h>

#if !defined(__CUDACC__) && !defined(__host__)
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace ksg {

// q for both parities, saturated at kChainSat (>= 2^24 means "leaves the binade" whatever M is), and the resulting parities.
struct ChainTable {
  uint32_t inc[2];
  uint32_t par;   // bit p = parity of M after the record when it was p before
};
static constexpr uint32_t kChainSat = 1u << 25;

__host__ __device__ __forceinline__ uint32_t chain_bits(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  return b;
}
__host__ __device__ __forceinline__ float chain_float(uint32_t b) {
  float x;
  memcpy(&x, &b, 4);
  return x;
}

// |x| = m * 2^e with m < 2^24 (m = 0 for +-0); subnormals keep e = -149.
__host__ __device__ __forceinline__ void chain_decompose(float x, uint32_t& m, int& e) {
  const uint32_t b = chain_bits(x) & 0x7FFFFFFFu;
  const uint32_t ex = b >> 23, fr = b & 0x7FFFFFu;
  if (ex == 0) { m = fr; e = -149; }
  else { m = fr | 0x800000u; e = (int)ex - 150; }
}

// -(m * 2^e) for a normal result: m in [2^23, 2^24), e in [-149, 104].
__host__ __device__ __forceinline__ float chain_make_negative(uint32_t m, int e) {
  return chain_float(0x80000000u | ((uint32_t)(e + 150) << 23) | (m & 0x7FFFFFu));
}

__host__ __device__ __forceinline__ uint32_t chain_sat_add(uint32_t a, uint32_t b) {
  const uint32_t s = a + b;   // both <= 2^25: no wrap
  return s > kChainSat ? kChainSat : s;
}

// The record a (a <= 0 or zero) on the grid u = 2^g.
__host__ __device__ __forceinline__ ChainTable chain_record_table(float a, int g) {
  ChainTable t;
  uint32_t m;
  int e;
  chain_decompose(a, m, e);
  if (m == 0) { t.inc[0] = t.inc[1] = 0; t.par = 2u; return t; }          // identity: parity p stays p  (bit1 = 1, bit0 = 0)
  const int shift = g - e;                                                // |a| / u = m * 2^(-shift)
  if (shift <= 0) {                                                       // |a| is a multiple of u
    const uint32_t q = (-shift >= 2) ? kChainSat : (m << (-shift));       // m >= 2^23 for normals: two shifts already leave the binade
    const uint32_t qs = q > kChainSat ? kChainSat : q;
    t.inc[0] = t.inc[1] = qs;
    t.par = (qs & 1u) ? 1u : 2u;                                          // odd q flips the parity, even q keeps it
    return t;
  }
  if (shift > 24) { t.inc[0] = t.inc[1] = 0; t.par = 2u; return t; }      // |a| < u / 2: rounds away (a tie needs shift <= 24)
  const uint32_t n = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
  if (rem != half) {
    const uint32_t q = n + (rem > half ? 1u : 0u);
    t.inc[0] = t.inc[1] = q;
    t.par = (q & 1u) ? 1u : 2u;
    return t;
  }
  // exact tie: M + n + 1/2 -> the even neighbour
  t.inc[0] = n + ((n & 1u) ? 1u : 0u);        // M even: M + n is odd iff n is odd -> round up
  t.inc[1] = n + ((n & 1u) ? 0u : 1u);        // M odd
  t.par = 0u;                                 // result even for both
  return t;
}

// g after f
__host__ __device__ __forceinline__ ChainTable chain_compose(const ChainTable& f, const ChainTable& g) {
  ChainTable r;
  const uint32_t p0 = f.par & 1u, p1 = (f.par >> 1) & 1u;
  r.inc[0] = chain_sat_add(f.inc[0], g.inc[p0]);
  r.inc[1] = chain_sat_add(f.inc[1], g.inc[p1]);
  r.par = ((g.par >> p0) & 1u) | (((g.par >> p1) & 1u) << 1);
  return r;
}

// Reference driver with the structure of the future warp loop: `width` records per step (32 on the device), an inclusive scan of
// their tables, the first prefix that leaves the binade handled by one ordinary addition.  Returns the same float as the
// sequential loop.  s must be negative, normal and finite; terms <= 0.
__host__ __device__ inline float chain_sum_reference(float s, const float* terms, long long n, int width) {
  long long i = 0;
  ChainTable prefix[64];
  while (i < n) {
    uint32_t m;
    int g;
    chain_decompose(s, m, g);
    if (m < 0x800000u) { s = s + terms[i]; ++i; continue; }               // subnormal running value: plain addition
    const int cnt = (int)((n - i) < width ? (n - i) : width);
    for (int k = 0; k < cnt; ++k) {                                        // a Hillis-Steele scan over shuffles on the device
      const ChainTable t = chain_record_table(terms[i + k], g);
      prefix[k] = k ? chain_compose(prefix[k - 1], t) : t;
    }
    const uint32_t p = m & 1u;
    int leave = -1;
    for (int k = 0; k < cnt; ++k)                                          // a ballot + ffs on the device
      if (m + prefix[k].inc[p] >= (1u << 24)) { leave = k; break; }
    if (leave < 0) {
      s = chain_make_negative(m + prefix[cnt - 1].inc[p], g);
      i += cnt;
    } else {
      if (leave > 0) s = chain_make_negative(m + prefix[leave - 1].inc[p], g);
      s = s + terms[i + leave];                                            // the crossing record: one ordinary float addition
      i += leave + 1;
    }
  }
  return s;
}

#if defined(__CUDACC__)
// One warp evaluates one chain with lanes = records: the device form of chain_sum_reference (width 32).  Every lane ends with the
// same running value.  Not used by the tile kernel yet; exercised through ksg_debug_chain_sum.
__device__ __forceinline__ float chain_sum_warp(float s, const float* __restrict__ terms, long long n) {
  const int lane = threadIdx.x & 31;
  long long i = 0;
  while (i < n) {
    uint32_t m;
    int g;
    chain_decompose(s, m, g);
    if (m < 0x800000u) { s = s + terms[i]; ++i; continue; }               // uniform: every lane holds the same s
    const long long rest = n - i;
    const int cnt = rest < 32 ? (int)rest : 32;
    ChainTable t;
    if (lane < cnt) t = chain_record_table(terms[i + lane], g);
    else { t.inc[0] = t.inc[1] = 0u; t.par = 2u; }                         // identity
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {                              // inclusive scan: earlier records compose first
      ChainTable o;
      o.inc[0] = __shfl_up_sync(0xffffffffu, t.inc[0], off);
      o.inc[1] = __shfl_up_sync(0xffffffffu, t.inc[1], off);
      o.par = __shfl_up_sync(0xffffffffu, t.par, off);
      if (lane >= off) t = chain_compose(o, t);
    }
    const uint32_t mine = (m & 1u) ? t.inc[1] : t.inc[0];
    const unsigned leaves = __ballot_sync(0xffffffffu, lane < cnt && m + mine >= (1u << 24));
    if (!leaves) {
      s = chain_make_negative(m + __shfl_sync(0xffffffffu, mine, cnt - 1), g);
      i += cnt;
    } else {
      const int leave = __ffs(leaves) - 1;
      const uint32_t before = __shfl_sync(0xffffffffu, mine, leave > 0 ? leave - 1 : 0);
      if (leave > 0) s = chain_make_negative(m + before, g);
      s = s + terms[i + leave];                                            // the crossing record: one ordinary addition
      i += leave + 1;
    }
  }
  return s;
}
#endif

}  // namespace ksg
