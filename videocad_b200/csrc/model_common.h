// Helpers shared by the host orchestration (model_vit.cpp, model_seq.cpp).  Pure host C++: compiled by nvcc for
// the product library and by g++ for the CPU emulation used in host-logic tests; talks to the kernels only
// through kernels.h.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "kernels.h"
#include "host_util.h"

namespace vck {

#define VC_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc) return _rc;        \
  } while (0)

struct Split {  // split-bf16 matrix view
  bf16_t* hi;
  bf16_t* lo;
  int64_t ld;
};
static inline Split mk_split(bf16_t* hi, bf16_t* lo, int64_t ld) { Split s; s.hi = hi; s.lo = lo; s.ld = ld; return s; }
static inline Split wsplit(const vc_linear& l, int64_t in_features, int64_t row_offset = 0) {
  Split s;
  s.hi = const_cast<bf16_t*>(l.w_hi) + row_offset * in_features;
  s.lo = const_cast<bf16_t*>(l.w_lo) + row_offset * in_features;
  s.ld = in_features;
  return s;
}

// bump allocator over a caller-provided workspace; with base == nullptr it only measures
class Arena {
 public:
  Arena(void* base, size_t cap) : base_(reinterpret_cast<uint8_t*>(base)), cap_(cap), off_(0) {}
  template <class T>
  T* alloc(size_t count) {
    const size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
    T* p = base_ ? reinterpret_cast<T*>(base_ + off_) : nullptr;
    off_ += bytes;
    return p;
  }
  Split alloc_split(int64_t rows, int64_t cols) {
    Split s;
    s.hi = alloc<bf16_t>((size_t)rows * cols);
    s.lo = alloc<bf16_t>((size_t)rows * cols);
    s.ld = cols;
    return s;
  }
  size_t used() const { return off_; }
  bool ok() const { return base_ == nullptr || off_ <= cap_; }

 private:
  uint8_t* base_;
  size_t cap_, off_;
};

// ---- the three GEMM orientations of a Linear layer (W stored [out, in]) ---------------------------------------
// forward: y[M,out] = x[M,in] W^T  (+ fused epilogue configured by the caller in `d`)
static inline void gemm_linear_fwd(GemmDesc& d, const Split& x, const Split& w, int M, int out_f, int in_f, int passes) {
  gemm_desc_init(&d);
  d.a_hi = x.hi; d.a_lo = x.lo; d.lda = x.ld; d.a_mn_major = 0;
  d.b_hi = w.hi; d.b_lo = w.lo; d.ldb = w.ld; d.b_mn_major = 0;
  d.M = M; d.N = out_f; d.K = in_f; d.passes = passes;
}
// dgrad: dx[M,in] = g[M,out] W      (B operand = W read MN-major)
static inline void gemm_linear_dgrad(GemmDesc& d, const Split& g, const Split& w, int M, int out_f, int in_f, int passes) {
  gemm_desc_init(&d);
  d.a_hi = g.hi; d.a_lo = g.lo; d.lda = g.ld; d.a_mn_major = 0;
  d.b_hi = w.hi; d.b_lo = w.lo; d.ldb = w.ld; d.b_mn_major = 1;
  d.M = M; d.N = in_f; d.K = out_f; d.passes = passes;
}
// wgrad: dW[out,in] += g^T[out,M] x[M,in]  (both operands MN-major, contraction over the M rows, split-K atomics
// into the caller-zeroed dW)
static inline int wgrad_splitk(int M, int out_f, int in_f) {
  const long long tiles = (long long)((out_f + 127) / 128) * ((in_f + 127) / 128);
  const int num_kb = (M + 63) / 64;
  long long s = tiles >= 148 ? 1 : (148 + tiles - 1) / tiles;
  if (s > num_kb) s = num_kb;
  // keep at least 4 k-blocks per split so the pipeline has something to overlap
  if (num_kb / s < 4) s = num_kb / 4 > 0 ? num_kb / 4 : 1;
  return (int)s;
}
static inline int linear_wgrad(const Split& g, const Split& x, int M, int out_f, int in_f, float* dW, int passes, stream_t st) {
  GemmDesc d;
  gemm_desc_init(&d);
  d.a_hi = g.hi; d.a_lo = g.lo; d.lda = g.ld; d.a_mn_major = 1;
  d.b_hi = x.hi; d.b_lo = x.lo; d.ldb = x.ld; d.b_mn_major = 1;
  d.M = out_f; d.N = in_f; d.K = M; d.passes = passes;
  d.splitk = wgrad_splitk(M, out_f, in_f);  // > 1: atomic accumulation into the zeroed dW; == 1: plain store
  d.out_f32 = dW; d.ldo = in_f;
  return gemm(d, st);
}

static inline Drop site_drop(float p, int training, uint64_t seed, uint32_t site, const uint64_t* seed_dev = nullptr) {
  return make_drop(training ? p : 0.f, site, seed, seed_dev);
}

}  // namespace vck
