// Counter-based RNG for dropout masks (host + device).
//
// A dropout decision is a pure function of (seed, site, element index): the forward kernels and the backward kernels
// regenerate the same mask instead of storing it.  `site` identifies the dropout call site within one model forward; `idx`
// is the row-major element index of the tensor the dropout is applied to.
//
// Generator: a keyed 32-bit integer hash of the PAIR index (idx >> 1); its 32 output bits give the two elements of the pair
// 16 random bits each (low half: even element, high half: odd element).  The hash is the three-round multiply-xorshift
// permutation "triple32" (C. Wellons, hash-prospector: bias 0.0208 over all input bits) with the (seed, site) key mixed in
// before the first and after the second round.  Cost: ~14 integer instructions per two elements.
//
// History: the first version drew 32 bits per element from Philox4x32-7 (~50 instructions per four elements, more where a
// kernel needed only part of a counter's output).  ncu on the image-encoder attention kernels showed the RNG to be about a
// third of all executed instructions and the integer-multiply pipe throttling the issue rate
// (profiles/r01k_ncu_vit_attn_summary.txt).  Dropout needs Bernoulli(1-p) decisions that are reproducible, uniform and
// uncorrelated across elements, sites and seeds -- not a cryptographic-strength stream; tests/test_host_logic.py checks
// keep-rate, bucket uniformity and row/column/seed/site correlations of the masks.
// A 16-bit threshold quantises p to 1/65536 (p = 0.1 -> 6554/65536 = 0.100006).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VC_HD __host__ __device__ __forceinline__
#else
#define VC_HD inline
#endif

namespace vck {

struct DropKey {
  uint32_t k0, k1;
};

// 4 decisions' worth of random bits, one word per element: the 16 random bits sit in the HIGH half of each word (low half
// zero), so that `word >= threshold` with the 32-bit threshold round(p * 2^32) is the keep test
struct Rand4 {
  uint32_t v[4];
};

VC_HD uint32_t hash32(uint32_t x) {
  x ^= x >> 17; x *= 0xed5ad4bbu;
  x ^= x >> 11; x *= 0xac4c1b51u;
  x ^= x >> 15; x *= 0x31848babu;
  x ^= x >> 14;
  return x;
}

VC_HD DropKey drop_key(uint64_t seed, uint32_t site) {
  DropKey k;
  k.k0 = hash32((uint32_t)seed ^ (site * 0x9E3779B9u));
  k.k1 = hash32((uint32_t)(seed >> 32) + 0x85EBCA6Bu + k.k0);
  return k;
}

// 32 random bits for elements 2*pair (low 16 bits) and 2*pair + 1 (high 16 bits)
VC_HD uint32_t dropout_word(const DropKey& k, uint64_t pair) {
  uint32_t x = (uint32_t)pair ^ k.k0 ^ ((uint32_t)(pair >> 32) * 0x9E3779B9u);
  x ^= x >> 17; x *= 0xed5ad4bbu;
  x ^= x >> 11; x = x * 0xac4c1b51u + k.k1;
  x ^= x >> 15; x *= 0x31848babu;
  x ^= x >> 14;
  return x;
}

// elements [4*q, 4*q+3]
VC_HD Rand4 dropout_words(const DropKey& k, uint64_t q) {
  const uint32_t a = dropout_word(k, 2 * q), b = dropout_word(k, 2 * q + 1);
  Rand4 out;
  out.v[0] = a << 16; out.v[1] = a & 0xffff0000u; out.v[2] = b << 16; out.v[3] = b & 0xffff0000u;
  return out;
}
VC_HD Rand4 dropout_words(uint64_t seed, uint32_t site, uint64_t q) { return dropout_words(drop_key(seed, site), q); }

// the word of ONE element (same bits as dropout_words(...).v[idx & 3])
VC_HD uint32_t dropout_element(const DropKey& k, uint64_t idx) {
  const uint32_t w = dropout_word(k, idx >> 1);
  return (idx & 1ull) ? (w & 0xffff0000u) : (w << 16);
}

// seed of a dropout site: the device-resident value if a pointer was given (CUDA-graph replay), else the by-value seed
template <class DropT>
VC_HD uint64_t drop_seed(const DropT& d) {
  return d.seed_ptr != nullptr ? *d.seed_ptr : d.seed;
}
// key of a dropout site; a kernel computes it ONCE per thread (two hashes and possibly a global load)
template <class DropT>
VC_HD DropKey drop_key_of(const DropT& d) {
  if (!(d.p > 0.f)) { DropKey k; k.k0 = k.k1 = 0u; return k; }
  return drop_key(drop_seed(d), d.site);
}

// keep-probability threshold: keep iff word >= thresh, thresh = round(p * 2^32)
VC_HD uint32_t dropout_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t < 0.0) t = 0.0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}

}  // namespace vck
