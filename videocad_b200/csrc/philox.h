// Philox4x32-7 counter-based RNG for dropout masks (host + device).  Seven rounds is the smallest round count of
// Philox4x32 that passes BigCrush (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11, Table 2); the
// three rounds saved relative to the customary 10 matter here because the masks are regenerated inside GEMM epilogues and
// attention kernels (about a third of the attention kernels' instructions were RNG at 10 rounds).
//
// A dropout decision is a pure function of (seed, site, element index): the forward kernels and the
// backward kernels regenerate the same mask instead of storing it.  `site` identifies the dropout
// call site within one model forward (see dropout_sites.h); `idx` is the row-major element index
// of the tensor the dropout is applied to.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VC_HD __host__ __device__ __forceinline__
#else
#define VC_HD inline
#endif

namespace vck {

struct Philox4 {
  uint32_t v[4];
};

VC_HD uint32_t philox_mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

VC_HD Philox4 philox4x32_r7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = philox_mulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = philox_mulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  Philox4 out;
  out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
  return out;
}

// 4 random words for elements [4*q, 4*q+3] of dropout site `site`
VC_HD Philox4 dropout_words(uint64_t seed, uint32_t site, uint64_t q) {
  return philox4x32_r7((uint32_t)q, (uint32_t)(q >> 32), site, 0x5eedu, (uint32_t)seed, (uint32_t)(seed >> 32));
}

// seed of a dropout site: the device-resident value if a pointer was given (CUDA-graph replay), else the by-value seed
template <class DropT>
VC_HD uint64_t drop_seed(const DropT& d) {
  return d.seed_ptr != nullptr ? *d.seed_ptr : d.seed;
}

// keep-probability threshold: keep iff word >= thresh, thresh = round(p * 2^32)
VC_HD uint32_t dropout_threshold(float p) {
  double t = (double)p * 4294967296.0;
  if (t < 0.0) t = 0.0;
  if (t > 4294967295.0) t = 4294967295.0;
  return (uint32_t)t;
}

}  // namespace vck
