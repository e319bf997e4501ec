// TEST INFRASTRUCTURE ONLY -- plain C++ restatement of every kernel declared in videocad_b200/csrc/kernels.h,
// operating on HOST pointers.  Linked with the product's host orchestration (model_vit.cpp, model_seq.cpp) into
// oracle/_build/libvc_emu.so so that the orchestration (workspace layout, GEMM orientations, backward chain,
// dropout-site bookkeeping) can be checked against the torch oracle on a machine without a GPU.
// Never linked into, loaded by, or shipped with the product library.
//
// Each function follows the contract documented in kernels.h; arithmetic mirrors the CUDA kernels (same split-bf16
// operand model, same dropout masks) but accumulates in double.
#include <math.h>
#include <string.h>
#include <vector>
#include "kernels.h"
#include "host_util.h"
#include "dropout_rng.h"

namespace vck {

namespace {

inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);
  const uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)((u + r) >> 16);
}
inline float bf2f(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline void split1(float x, bf16_t& hi, bf16_t& lo) {
  hi = f2bf(x);
  lo = f2bf(x - bf2f(hi));
}
inline float keep_scale(const Drop& d, uint64_t idx) {
  if (d.p <= 0.f) return 1.0f;
  const Rand4 w = dropout_words(drop_seed(d), d.site, idx >> 2);
  return (w.v[idx & 3u] >= dropout_threshold(d.p)) ? 1.0f / (1.0f - d.p) : 0.0f;
}
inline double gelu_d(double x) { return 0.5 * x * (1.0 + erf(x * 0.70710678118654752440)); }
inline double gelu_grad_d(double x) {
  return 0.5 * (1.0 + erf(x * 0.70710678118654752440)) + x * 0.39894228040143267794 * exp(-0.5 * x * x);
}
inline double act_d(double x, int act) {
  switch (act) {
    case VC_ACT_GELU: return gelu_d(x);
    case VC_ACT_GELU_DSTORE: return gelu_d(x);
    case VC_ACT_RELU: return x > 0 ? x : 0;
    case VC_ACT_TANH: return tanh(x);
    default: return x;
  }
}
inline bool is_masked(int mask, int window, int i, int j) {
  if (mask == VC_MASK_CAUSAL) return j > i;
  if (mask == VC_MASK_WINDOW) return (j > i) || (j <= i - window);
  return false;
}

}  // namespace

void gemm_desc_init(GemmDesc* d) {
  memset(d, 0, sizeof *d);
  d->passes = 3;
  d->splitk = 1;
  d->rowadd_div = 1;
  d->rowadd_mod = 1;
}

int gemm(const GemmDesc& d, stream_t) {
  if (d.M <= 0 || d.N <= 0 || d.K <= 0) return set_error("gemm: empty problem");
  if (d.N % 8 != 0) return set_error("gemm: N must be a multiple of 8");
  if (d.lda % 8 != 0 || d.ldb % 8 != 0) return set_error("gemm: lda/ldb must be multiples of 8 elements");
  if (d.passes != 1 && d.passes != 3) return set_error("gemm: passes must be 1 or 3");
  if (!d.a_hi || !d.b_hi) return set_error("gemm: null operand");
  if (d.passes == 3 && (!d.a_lo || !d.b_lo)) return set_error("gemm: passes=3 needs lo operands");
  if (d.act_backward && ((d.act == VC_ACT_GELU || d.act == VC_ACT_TANH || d.act == VC_ACT_MUL_AUX) && !d.act_aux)) return set_error("gemm: act_aux required");
  if (d.act_backward && d.act == VC_ACT_RELU && !d.act_aux_hi) return set_error("gemm: act_aux_hi required");
  if (d.act == VC_ACT_GELU_DSTORE && (d.act_backward || !d.preact)) return set_error("gemm: VC_ACT_GELU_DSTORE is a forward activation and needs preact");
  if (d.act == VC_ACT_MUL_AUX && !d.act_backward) return set_error("gemm: VC_ACT_MUL_AUX is a backward-activation mode");
  if (d.splitk > 1 && (d.act != VC_ACT_NONE || d.drop.p > 0.f || d.residual || d.out_hi || d.preact || d.colsum))
    return set_error("gemm: split-K supports only the bias epilogue with an fp32 (atomic) output");
  if (!d.out_f32 && !d.out_hi) return set_error("gemm: no output");
  std::vector<double> colsum_acc(d.colsum ? d.N : 0, 0.0);
  std::vector<float> vals(d.colsum ? (size_t)d.M * d.N : 0);
  const int num_kb = (d.K + 63) / 64;
  int splitk = d.splitk < 1 ? 1 : d.splitk;
  if (splitk > num_kb) splitk = num_kb;
  const int kbps = (num_kb + splitk - 1) / splitk;
  splitk = (num_kb + kbps - 1) / kbps;
  // operands to float
  std::vector<float> Ah((size_t)d.M * d.K), Al, Bh((size_t)d.N * d.K), Bl;
  if (d.passes == 3) { Al.resize(Ah.size()); Bl.resize(Bh.size()); }
  for (int m = 0; m < d.M; ++m)
    for (int k = 0; k < d.K; ++k) {
      const size_t src = d.a_mn_major ? (size_t)k * d.lda + m : (size_t)m * d.lda + k;
      Ah[(size_t)m * d.K + k] = bf2f(d.a_hi[src]);
      if (d.passes == 3) Al[(size_t)m * d.K + k] = bf2f(d.a_lo[src]);
    }
  for (int n = 0; n < d.N; ++n)
    for (int k = 0; k < d.K; ++k) {
      const size_t src = d.b_mn_major ? (size_t)k * d.ldb + n : (size_t)n * d.ldb + k;
      Bh[(size_t)n * d.K + k] = bf2f(d.b_hi[src]);
      if (d.passes == 3) Bl[(size_t)n * d.K + k] = bf2f(d.b_lo[src]);
    }
  const int rdiv = d.rowadd_div > 0 ? d.rowadd_div : 1, rmod = d.rowadd_mod > 0 ? d.rowadd_mod : 1;
#pragma omp parallel for schedule(static)
  for (int m = 0; m < d.M; ++m) {
    for (int n = 0; n < d.N; ++n) {
      const float* ah = &Ah[(size_t)m * d.K];
      const float* bh = &Bh[(size_t)n * d.K];
      double acc = 0.0;
      if (d.passes == 3) {
        const float* al = &Al[(size_t)m * d.K];
        const float* bl = &Bl[(size_t)n * d.K];
        for (int k = 0; k < d.K; ++k) acc += (double)ah[k] * bh[k] + (double)al[k] * bh[k] + (double)ah[k] * bl[k];
      } else {
        for (int k = 0; k < d.K; ++k) acc += (double)ah[k] * bh[k];
      }
      double v = acc;
      if (d.bias) v += d.bias[n];
      if (d.rowadd) v += d.rowadd[(size_t)((m / rdiv) % rmod) * d.ld_rowadd + n];
      const bool dstore = d.act == VC_ACT_GELU_DSTORE;  // preact receives gelu'(v) * mask * scale instead of v
      if (d.preact && !dstore) d.preact[(size_t)m * d.ld_preact + n] = (float)v;
      double dv = dstore ? gelu_grad_d((double)(float)v) : 0.0;
      if (!d.act_backward) v = act_d(v, d.act);
      const double ks = keep_scale(d.drop, (uint64_t)m * d.N + n);
      v *= ks;
      if (dstore) d.preact[(size_t)m * d.ld_preact + n] = (float)(dv * ks);
      if (d.act_backward) {
        if (d.act == VC_ACT_GELU) v *= gelu_grad_d(d.act_aux[(size_t)m * d.ld_act_aux + n]);
        else if (d.act == VC_ACT_MUL_AUX) v *= d.act_aux[(size_t)m * d.ld_act_aux + n];
        else if (d.act == VC_ACT_TANH) { const double t = d.act_aux[(size_t)m * d.ld_act_aux + n]; v *= 1.0 - t * t; }
        else if (d.act == VC_ACT_RELU) { if ((d.act_aux_hi[(size_t)m * d.ld_act_aux_hi + n] & 0x7fffu) == 0) v = 0; }
      }
      if (d.residual) v += d.residual[(size_t)m * d.ld_res + n];
      const float vf = (float)v;
      if (d.colsum) vals[(size_t)m * d.N + n] = vf;
      if (d.out_f32) {
        float* o = d.out_f32 + (size_t)m * d.ldo + n;
        if (splitk > 1) *o += vf; else *o = vf;
      }
      if (d.out_hi) {
        bf16_t hi, lo;
        split1(vf, hi, lo);
        d.out_hi[(size_t)m * d.ldo_split + n] = hi;
        if (d.out_lo) d.out_lo[(size_t)m * d.ldo_split + n] = lo;
      }
    }
  }
  if (d.colsum) {
    for (int m = 0; m < d.M; ++m)
      for (int n = 0; n < d.N; ++n) colsum_acc[n] += vals[(size_t)m * d.N + n];
    for (int n = 0; n < d.N; ++n) d.colsum[n] += (float)colsum_acc[n];
  }
  return 0;
}

int split_f32(const float* x, int64_t ldx, int64_t rows, int64_t cols, bf16_t* hi, bf16_t* lo, int64_t ldo, stream_t) {
  if (cols % 4 != 0 || ldx % 4 != 0 || ldo % 4 != 0) return set_error("split_f32: cols/ld must be multiples of 4");
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < cols; ++c) split1(x[r * ldx + c], hi[r * ldo + c], lo[r * ldo + c]);
  return 0;
}

int split_many(const vc_split_item* items, int n_items, int64_t, stream_t) {
  for (int k = 0; k < n_items; ++k)
    for (int64_t i = 0; i < items[k].n4 * 4; ++i) split1(items[k].src[i], items[k].hi[i], items[k].lo[i]);
  return 0;
}

namespace {
template <class Load>
void ln_fwd_rows(Load load, int64_t rows, int C, const float* gamma, const float* beta, float eps, float* y, int64_t ldy,
                 bf16_t* y_hi, bf16_t* y_lo, int64_t ldys, float* mean, float* rstd) {
  for (int64_t r = 0; r < rows; ++r) {
    double s = 0;
    for (int c = 0; c < C; ++c) s += load(r, c);
    const double mu = s / C;
    double q = 0;
    for (int c = 0; c < C; ++c) { const double t = load(r, c) - mu; q += t * t; }
    const double rs = 1.0 / sqrt(q / C + eps);
    if (mean) mean[r] = (float)mu;
    if (rstd) rstd[r] = (float)rs;
    for (int c = 0; c < C; ++c) {
      const float o = (float)((load(r, c) - mu) * rs * gamma[c] + beta[c]);
      if (y) y[r * ldy + c] = o;
      if (y_hi) {
        bf16_t h, l;
        split1(o, h, l);
        y_hi[r * ldys + c] = h;
        if (y_lo) y_lo[r * ldys + c] = l;
      }
    }
  }
}
template <class Load>
void ln_bwd_rows(Load load, const float* dy, int64_t lddy, const float* mean, const float* rstd, const float* gamma,
                 int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx, float* dgamma,
                 float* dbeta) {
  std::vector<double> dg(C, 0.0), db(C, 0.0);
  for (int64_t r = 0; r < rows; ++r) {
    const double mu = mean[r], rs = rstd[r];
    double s1 = 0, s2 = 0;
    for (int c = 0; c < C; ++c) {
      const double xh = (load(r, c) - mu) * rs, d = dy[r * lddy + c];
      dg[c] += d * xh;
      db[c] += d;
      if (dx) { const double g = d * gamma[c]; s1 += g; s2 += g * xh; }
    }
    if (dx) {
      s1 /= C; s2 /= C;
      for (int c = 0; c < C; ++c) {
        const double xh = (load(r, c) - mu) * rs, g = (double)dy[r * lddy + c] * gamma[c];
        double o = rs * (g - s1 - xh * s2);
        if (dres) o += dres[r * lddres + c];
        dx[r * lddx + c] = (float)o;
      }
    }
  }
  for (int c = 0; c < C; ++c) { dgamma[c] += (float)dg[c]; dbeta[c] += (float)db[c]; }
}
struct PatchLoad {
  const float* img; int S, wp, N;
  double operator()(int64_t row, int c) const {
    const int64_t f = row / N;
    const int pidx = (int)(row % N), ph = pidx / wp, pw = pidx % wp, p1 = c >> 5, p2 = c & 31;
    return img[(f * S + (ph * 32 + p1)) * (int64_t)S + pw * 32 + p2];
  }
};
}  // namespace

int layernorm_fwd(const float* x, int64_t ldx, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                  float* y, int64_t ldy, bf16_t* y_hi, bf16_t* y_lo, int64_t ldy_split, float* mean, float* rstd, stream_t) {
  if (C % 128 != 0 || C > 1024) return set_error("layernorm_fwd: C must be a multiple of 128 and <= 1024");
  ln_fwd_rows([&](int64_t r, int c) { return (double)x[r * ldx + c]; }, rows, C, gamma, beta, eps, y, ldy, y_hi, y_lo, ldy_split,
              mean, rstd);
  return 0;
}
int layernorm_bwd(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                  const float* gamma, int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                  float* dgamma, float* dbeta, stream_t) {
  if (C % 128 != 0 || C > 1024) return set_error("layernorm_bwd: C must be a multiple of 128 and <= 1024");
  ln_bwd_rows([&](int64_t r, int c) { return (double)x[r * ldx + c]; }, dy, lddy, mean, rstd, gamma, rows, C, dres, lddres, dx, lddx,
              dgamma, dbeta);
  return 0;
}
int layernorm_bwd_fused(const float* dy, int64_t lddy, const float* x, int64_t ldx, const float* mean, const float* rstd,
                        const float* gamma, int64_t rows, int C, const float* dres, int64_t lddres, float* dx, int64_t lddx,
                        float* dgamma, float* dbeta, Drop gdrop, bf16_t* g_hi, bf16_t* g_lo, int64_t ldg, float* g_colsum,
                        stream_t st) {
  if (int rc = layernorm_bwd(dy, lddy, x, ldx, mean, rstd, gamma, rows, C, dres, lddres, dx, lddx, dgamma, dbeta, st)) return rc;
  if (g_hi) {
    for (int64_t r = 0; r < rows; ++r)
      for (int c = 0; c < C; ++c) {
        const float v = dx[r * lddx + c] * keep_scale(gdrop, (uint64_t)r * C + c);
        split1(v, g_hi[r * ldg + c], g_lo[r * ldg + c]);
        if (g_colsum) g_colsum[c] += v;
      }
  }
  return 0;
}

int patch_layernorm_fwd(const float* img, int F, int S, const float* gamma, const float* beta, float eps, bf16_t* y_hi,
                        bf16_t* y_lo, float* mean, float* rstd, stream_t) {
  if (S % 32 != 0) return set_error("patch_layernorm_fwd: image size must be a multiple of 32");
  const int wp = S / 32, N = wp * wp;
  PatchLoad ld{img, S, wp, N};
  ln_fwd_rows(ld, (int64_t)F * N, 1024, gamma, beta, eps, nullptr, 0, y_hi, y_lo, 1024, mean, rstd);
  return 0;
}
int patch_layernorm_bwd_params(const float* img, int F, int S, const float* mean, const float* rstd, const float* dy,
                               float* dgamma, float* dbeta, stream_t) {
  if (S % 32 != 0) return set_error("patch_layernorm_bwd_params: image size must be a multiple of 32");
  const int wp = S / 32, N = wp * wp;
  PatchLoad ld{img, S, wp, N};
  ln_bwd_rows(ld, dy, 1024, mean, rstd, nullptr, (int64_t)F * N, 1024, nullptr, 0, nullptr, 0, dgamma, dbeta);
  return 0;
}

int vit_assemble_fwd(const float* e, int F, int N, int C, const float* cls, const float* pos, Drop drop, float* x, stream_t) {
  const int n = N + 1;
  for (int64_t f = 0; f < F; ++f)
    for (int t = 0; t < n; ++t)
      for (int c = 0; c < C; ++c) {
        const int64_t idx = (f * n + t) * C + c;
        const float v = (t == 0 ? cls[c] : e[(f * N + t - 1) * C + c]) + pos[(int64_t)t * C + c];
        x[idx] = v * keep_scale(drop, (uint64_t)idx);
      }
  return 0;
}
int vit_assemble_bwd(const float* dx, int F, int N, int C, Drop drop, float* de, float* dcls, float* dpos, stream_t) {
  const int n = N + 1;
  for (int64_t f = 0; f < F; ++f)
    for (int t = 0; t < n; ++t)
      for (int c = 0; c < C; ++c) {
        const int64_t idx = (f * n + t) * C + c;
        const float g = dx[idx] * keep_scale(drop, (uint64_t)idx);
        if (t > 0) de[(f * N + t - 1) * C + c] = g;
        dpos[(int64_t)t * C + c] += g;
        if (t == 0) dcls[c] += g;
      }
  return 0;
}

namespace {
int attn_validate(const AttnDesc& a) {
  if (a.d % 4 != 0 || a.d > 256 || a.d <= 0) return set_error("attention: head dim must be a multiple of 4 and <= 256");
  if (a.mask != VC_MASK_NONE && a.Tq != a.Tk) return set_error("attention: masked attention needs Tq == Tk");
  if (a.mask == VC_MASK_WINDOW && a.window < 1) return set_error("attention: window must be >= 1");
  return 0;
}
}  // namespace

namespace {
// split-bf16 q/k/v -> temporary fp32 copies (the emulation always computes from fp32)
struct JoinedQKV {
  std::vector<float> q, k, v;
  AttnDesc d;
};
void join_inputs(const AttnDesc& a, JoinedQKV& j) {
  j.d = a;
  if (a.q_hi == nullptr) return;
  const int64_t W = (int64_t)a.nh * a.d, Rq = (int64_t)a.B * a.Tq, Rk = (int64_t)a.B * a.Tk;
  j.q.resize(Rq * W); j.k.resize(Rk * W); j.v.resize(Rk * W);
  for (int64_t r = 0; r < Rq; ++r)
    for (int64_t c = 0; c < W; ++c) j.q[r * W + c] = bf2f(a.q_hi[r * a.ldq + c]) + bf2f(a.q_lo[r * a.ldq + c]);
  for (int64_t r = 0; r < Rk; ++r)
    for (int64_t c = 0; c < W; ++c) {
      j.k[r * W + c] = bf2f(a.k_hi[r * a.ldk + c]) + bf2f(a.k_lo[r * a.ldk + c]);
      j.v[r * W + c] = bf2f(a.v_hi[r * a.ldv + c]) + bf2f(a.v_lo[r * a.ldv + c]);
    }
  j.d.q = j.q.data(); j.d.k = j.k.data(); j.d.v = j.v.data();
  j.d.ldq = j.d.ldk = j.d.ldv = W;
  j.d.q_hi = j.d.q_lo = j.d.k_hi = j.d.k_lo = j.d.v_hi = j.d.v_lo = nullptr;
}
}  // namespace

int attention_fwd(const AttnDesc& a_in, bf16_t* o_hi, bf16_t* o_lo, int64_t ldo, float* lse, stream_t) {
  JoinedQKV joined;
  join_inputs(a_in, joined);
  const AttnDesc& a = joined.d;
  if (int rc = attn_validate(a)) return rc;
  const int d = a.d;
  if ((a_in.bsq || a_in.bsk || a_in.bsv) && a_in.q_hi != nullptr) return set_error("attention_fwd: batch strides need fp32 inputs");
  const int64_t bsq = a_in.bsq ? a_in.bsq : (int64_t)a.Tq * a.ldq, bsk = a_in.bsk ? a_in.bsk : (int64_t)a.Tk * a.ldk,
                bsv = a_in.bsv ? a_in.bsv : (int64_t)a.Tk * a.ldv;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < a.B; ++b)
    for (int h = 0; h < a.nh; ++h) {
      std::vector<double> s(a.Tk), acc(d);
      for (int i = 0; i < a.Tq; ++i) {
        const float* q = a.q + (int64_t)b * bsq + (int64_t)i * a.ldq + (int64_t)h * d;
        double m = -INFINITY;
        for (int j = 0; j < a.Tk; ++j) {
          if (is_masked(a.mask, a.window, i, j)) { s[j] = -INFINITY; continue; }
          const float* k = a.k + (int64_t)b * bsk + (int64_t)j * a.ldk + (int64_t)h * d;
          double t = 0;
          for (int c = 0; c < d; ++c) t += (double)q[c] * k[c];
          s[j] = t * a.scale;
          if (s[j] > m) m = s[j];
        }
        double sum = 0;
        for (int j = 0; j < a.Tk; ++j) sum += exp(s[j] - m);
        const double l = m + log(sum);
        lse[((int64_t)b * a.nh + h) * a.Tq + i] = (float)l;
        for (int c = 0; c < d; ++c) acc[c] = 0;
        for (int j = 0; j < a.Tk; ++j) {
          if (s[j] == -INFINITY) continue;
          const uint64_t idx = (((uint64_t)b * a.nh + h) * a.Tq + i) * (uint64_t)a.Tk + j;
          const double pj = exp(s[j] - l) * keep_scale(a.drop, idx);
          const float* v = a.v + (int64_t)b * bsv + (int64_t)j * a.ldv + (int64_t)h * d;
          for (int c = 0; c < d; ++c) acc[c] += pj * v[c];
        }
        for (int c = 0; c < d; ++c) {
          bf16_t hi, lo;
          split1((float)acc[c], hi, lo);
          const int64_t off = ((int64_t)b * a.Tq + i) * ldo + (int64_t)h * d + c;
          o_hi[off] = hi;
          if (o_lo) o_lo[off] = lo;
        }
      }
    }
  return 0;
}

int attention_bwd(const AttnDesc& a_in, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse, const float* dout,
                  int64_t lddo, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv, int64_t lddv, stream_t) {
  JoinedQKV joined;
  join_inputs(a_in, joined);
  const AttnDesc& a = joined.d;
  if (int rc = attn_validate(a)) return rc;
  const int d = a.d;
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < a.B; ++b)
    for (int h = 0; h < a.nh; ++h) {
      std::vector<double> dK((size_t)a.Tk * d, 0.0), dV((size_t)a.Tk * d, 0.0), dQ(d);
      for (int i = 0; i < a.Tq; ++i) {
        const int64_t ro = (int64_t)b * a.Tq + i;
        const float* q = a.q + ro * a.ldq + (int64_t)h * d;
        const float* dO = dout + ro * lddo + (int64_t)h * d;
        double delta = 0;
        for (int c = 0; c < d; ++c) {
          const int64_t oo = ro * ldo + (int64_t)h * d + c;
          delta += (double)dO[c] * (bf2f(o_hi[oo]) + (o_lo ? bf2f(o_lo[oo]) : 0.f));
        }
        const double l = lse[((int64_t)b * a.nh + h) * a.Tq + i];
        for (int c = 0; c < d; ++c) dQ[c] = 0;
        for (int j = 0; j < a.Tk; ++j) {
          if (is_masked(a.mask, a.window, i, j)) continue;
          const float* k = a.k + ((int64_t)b * a.Tk + j) * a.ldk + (int64_t)h * d;
          const float* v = a.v + ((int64_t)b * a.Tk + j) * a.ldv + (int64_t)h * d;
          double s = 0, dp = 0;
          for (int c = 0; c < d; ++c) { s += (double)q[c] * k[c]; dp += (double)dO[c] * v[c]; }
          const double pr = exp(s * a.scale - l);
          const uint64_t idx = (((uint64_t)b * a.nh + h) * a.Tq + i) * (uint64_t)a.Tk + j;
          const double m = keep_scale(a.drop, idx);
          const double pt = pr * m, ds = pr * (dp * m - delta) * a.scale;
          for (int c = 0; c < d; ++c) {
            dV[(size_t)j * d + c] += pt * dO[c];
            dK[(size_t)j * d + c] += ds * q[c];
            dQ[c] += ds * k[c];
          }
        }
        for (int c = 0; c < d; ++c) dq[ro * lddq + (int64_t)h * d + c] = (float)dQ[c];
      }
      for (int j = 0; j < a.Tk; ++j)
        for (int c = 0; c < d; ++c) {
          dk[((int64_t)b * a.Tk + j) * lddk + (int64_t)h * d + c] = (float)dK[(size_t)j * d + c];
          dv[((int64_t)b * a.Tk + j) * lddv + (int64_t)h * d + c] = (float)dV[(size_t)j * d + c];
        }
    }
  return 0;
}

int attention_bwd_split(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                        const float* dout, const bf16_t* dout_hi, const bf16_t* dout_lo, int64_t lddo, float* scratch, bf16_t* dq_hi,
                        bf16_t* dq_lo, bf16_t* dk_hi, bf16_t* dk_lo, bf16_t* dv_hi, bf16_t* dv_lo, int64_t ld_split, stream_t st) {
  if (!scratch) return set_error("attention_bwd_split: scratch required");
  const int64_t W = (int64_t)a.nh * a.d, Rq = (int64_t)a.B * a.Tq, Rk = (int64_t)a.B * a.Tk;
  std::vector<float> dout_joined;
  if (dout == nullptr) {
    if (!dout_hi || !dout_lo) return set_error("attention_bwd_split: no upstream gradient");
    dout_joined.resize(Rq * W);
    for (int64_t r = 0; r < Rq; ++r)
      for (int64_t c = 0; c < W; ++c) dout_joined[r * W + c] = bf2f(dout_hi[r * lddo + c]) + bf2f(dout_lo[r * lddo + c]);
    dout = dout_joined.data();
    lddo = W;
  }
  float* dq = scratch;
  float* dk = dq + Rq * W;
  float* dv = dk + Rk * W;
  if (int rc = attention_bwd(a, o_hi, o_lo, ldo, lse, dout, lddo, dq, W, dk, W, dv, W, st)) return rc;
  if (int rc = split_f32(dq, W, Rq, W, dq_hi, dq_lo, ld_split, st)) return rc;
  if (int rc = split_f32(dk, W, Rk, W, dk_hi, dk_lo, ld_split, st)) return rc;
  return split_f32(dv, W, Rk, W, dv_hi, dv_lo, ld_split, st);
}

int act_dropout_bwd(const float* dy, int64_t lddy, int64_t M, int N, int act, const float* aux, int64_t ldaux,
                    const bf16_t* aux_hi, int64_t ldaux_hi, Drop drop, float* g, int64_t ldg, bf16_t* g_hi, bf16_t* g_lo,
                    int64_t ldg_split, float* colsum, stream_t) {
  if (N % 4 != 0) return set_error("act_dropout_bwd: N % 4 != 0");
  if ((act == VC_ACT_GELU || act == VC_ACT_TANH) && !aux) return set_error("act_dropout_bwd: aux required");
  if (act == VC_ACT_RELU && !aux_hi) return set_error("act_dropout_bwd: aux_hi required for relu");
  std::vector<double> cs(N, 0.0);
  for (int64_t r = 0; r < M; ++r)
    for (int c = 0; c < N; ++c) {
      double v = (double)dy[r * lddy + c] * keep_scale(drop, (uint64_t)r * N + c);
      if (act == VC_ACT_GELU) v *= gelu_grad_d(aux[r * ldaux + c]);
      else if (act == VC_ACT_TANH) { const double t = aux[r * ldaux + c]; v *= 1.0 - t * t; }
      else if (act == VC_ACT_RELU) { if ((aux_hi[r * ldaux_hi + c] & 0x7fffu) == 0) v = 0; }
      const float vf = (float)v;
      if (g) g[r * ldg + c] = vf;
      if (g_hi) {
        bf16_t hi, lo;
        split1(vf, hi, lo);
        g_hi[r * ldg_split + c] = hi;
        if (g_lo) g_lo[r * ldg_split + c] = lo;
      }
      cs[c] += vf;
    }
  if (colsum) for (int c = 0; c < N; ++c) colsum[c] += (float)cs[c];
  return 0;
}

int attention_bwd_split_bias(const AttnDesc& a, const bf16_t* o_hi, const bf16_t* o_lo, int64_t ldo, const float* lse,
                             const float* dout, int64_t lddo, float* scratch, bf16_t* dq_hi, bf16_t* dq_lo, bf16_t* dk_hi,
                             bf16_t* dk_lo, bf16_t* dv_hi, bf16_t* dv_lo, int64_t ld_split, float* dbq, float* dbk, float* dbv,
                             stream_t st) {
  if (!scratch) return set_error("attention_bwd_split_bias: scratch required");
  if (!dout) return set_error("attention_bwd_split_bias: an fp32 upstream gradient is required");
  const int64_t W = (int64_t)a.nh * a.d, Rq = (int64_t)a.B * a.Tq, Rk = (int64_t)a.B * a.Tk;
  float* dq = scratch;
  float* dk = dq + Rq * W;
  float* dv = dk + Rk * W;
  if (int rc = attention_bwd(a, o_hi, o_lo, ldo, lse, dout, lddo, dq, W, dk, W, dv, W, st)) return rc;
  if (int rc = act_dropout_bwd(dq, W, Rq, (int)W, VC_ACT_NONE, nullptr, 0, nullptr, 0, no_drop(), nullptr, 0, dq_hi, dq_lo, ld_split, dbq, st)) return rc;
  if (int rc = act_dropout_bwd(dk, W, Rk, (int)W, VC_ACT_NONE, nullptr, 0, nullptr, 0, no_drop(), nullptr, 0, dk_hi, dk_lo, ld_split, dbk, st)) return rc;
  return act_dropout_bwd(dv, W, Rk, (int)W, VC_ACT_NONE, nullptr, 0, nullptr, 0, no_drop(), nullptr, 0, dv_hi, dv_lo, ld_split, dbv, st);
}

int linear_rows_fwd(const float* x, const bf16_t* x_hi, const bf16_t* x_lo, int64_t ldx, int M, const float* W, const float* bias, int N,
                    int K, int act, const float* residual, int64_t ld_res, float* out_f32, int64_t ldo, bf16_t* out_hi, bf16_t* out_lo,
                    int64_t ldo_split, stream_t) {
  if (M > 16) return set_error("linear_rows_fwd: at most 16 rows");
  if ((x == nullptr) == (x_hi == nullptr)) return set_error("linear_rows_fwd: give x either as fp32 or as split-bf16");
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) {
        const float xv = x ? x[m * ldx + k] : bf2f(x_hi[m * ldx + k]) + (x_lo ? bf2f(x_lo[m * ldx + k]) : 0.f);
        acc += (double)xv * W[(int64_t)n * K + k];
      }
      double v = acc + (bias ? bias[n] : 0.f);
      if (act == VC_ACT_RELU) v = v > 0 ? v : 0;
      else if (act == VC_ACT_TANH) v = tanh(v);
      else if (act == VC_ACT_GELU) v = gelu_d(v);
      if (residual) v += residual[m * ld_res + n];
      const float vf = (float)v;
      if (out_f32) out_f32[m * ldo + n] = vf;
      if (out_hi) {
        bf16_t hi, lo;
        split1(vf, hi, lo);
        out_hi[m * ldo_split + n] = hi;
        if (out_lo) out_lo[m * ldo_split + n] = lo;
      }
    }
  return 0;
}

int frames_rgb_u8_ingest(const uint8_t* src, int64_t n, int Hin, int Win, int Hout, int Wout, const int* kk_h, const int* bounds_h, int ks_h,
                         const int* kk_v, const int* bounds_v, int ks_v, uint8_t* tmp, float mean, float std, float* dst, stream_t) {
  if (!src || !dst || n <= 0 || Hin <= 0 || Win <= 0 || Hout <= 0 || Wout <= 0) return set_error("frames_rgb_u8_ingest: bad arguments");
  if (!(std != 0.f)) return set_error("frames_rgb_u8_ingest: std must be non-zero");
  auto clip8 = [](int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); };
  const uint8_t* cur = src;
  if (Win != Wout) {
    if (!kk_h || !bounds_h || !tmp) return set_error("frames_rgb_u8_ingest: horizontal pass needs coefficients and a temporary");
    for (int64_t row = 0; row < n * Hin; ++row)
      for (int xx = 0; xx < Wout; ++xx)
        for (int c = 0; c < 3; ++c) {
          int acc = 1 << 21;
          for (int x = 0; x < bounds_h[2 * xx + 1]; ++x) acc += (int)src[(row * Win + bounds_h[2 * xx] + x) * 3 + c] * kk_h[(int64_t)xx * ks_h + x];
          tmp[(row * Wout + xx) * 3 + c] = (uint8_t)clip8(acc >> 22);
        }
    cur = tmp;
  }
  for (int64_t f = 0; f < n; ++f)
    for (int yy = 0; yy < Hout; ++yy)
      for (int x = 0; x < Wout; ++x) {
        int rgb[3];
        for (int c = 0; c < 3; ++c) {
          if (Hin != Hout) {
            if (!kk_v || !bounds_v) return set_error("frames_rgb_u8_ingest: vertical pass needs coefficients");
            int acc = 1 << 21;
            for (int y = 0; y < bounds_v[2 * yy + 1]; ++y) acc += (int)cur[((f * Hin + bounds_v[2 * yy] + y) * Wout + x) * 3 + c] * kk_v[(int64_t)yy * ks_v + y];
            rgb[c] = clip8(acc >> 22);
          } else {
            rgb[c] = cur[((f * Hin + yy) * Wout + x) * 3 + c];
          }
        }
        const unsigned l = ((unsigned)rgb[0] * 19595u + (unsigned)rgb[1] * 38470u + (unsigned)rgb[2] * 7471u + 0x8000u) >> 16;
        volatile float v = (float)l / 255.0f;
        volatile float w = v - mean;
        dst[(f * Hout + yy) * Wout + x] = w / std;
      }
  return 0;
}

int frames_u8_normalize(const uint8_t* src, int64_t n, float mean, float std, float* dst, stream_t) {
  if (n > 0 && (!src || !dst)) return set_error("frames_u8_normalize: null pointer");
  if (!(std != 0.f)) return set_error("frames_u8_normalize: std must be non-zero");
  for (int64_t i = 0; i < n; ++i) {
    volatile float v = (float)src[i] / 255.0f;  // one correctly-rounded fp32 operation per step, as torchvision does
    volatile float w = v - mean;
    dst[i] = w / std;
  }
  return 0;
}

void attention_small_enable(int) {}
void gemm_pair_force_tile(int) {}

int row_reduce_mod(const float* x, int64_t ldx, int64_t M, int N, int div, int mod, float* out, stream_t) {
  if (div < 1 || mod < 1) return set_error("row_reduce_mod: div/mod must be >= 1");
  for (int64_t m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) out[((m / div) % mod) * N + n] += x[m * ldx + n];
  return 0;
}

int broadcast_rows(const float* src, int64_t lds, int64_t M, int N, int div, float* dst, int64_t ldd, bf16_t* d_hi, bf16_t* d_lo,
                   int64_t ldd_split, stream_t) {
  for (int64_t m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      const float v = src[(m / div) * lds + n];
      if (dst) dst[m * ldd + n] = v;
      if (d_hi) {
        bf16_t hi, lo;
        split1(v, hi, lo);
        d_hi[m * ldd_split + n] = hi;
        if (d_lo) d_lo[m * ldd_split + n] = lo;
      }
    }
  return 0;
}

int embed_action_fwd(const float* actions, int64_t R, int A, int H, const float* W, const float* b, const float* E, int T,
                     float* y, bf16_t* y_hi, bf16_t* y_lo, stream_t) {
  for (int64_t r = 0; r < R; ++r)
    for (int h = 0; h < H; ++h) {
      double acc = b[h];
      for (int a = 0; a < A; ++a) acc += (double)actions[r * A + a] * W[(int64_t)h * A + a];
      if (E) acc += E[(r % T) * H + h];
      const float v = (float)tanh(acc);
      if (y) y[r * H + h] = v;
      if (y_hi) {
        bf16_t hi, lo;
        split1(v, hi, lo);
        y_hi[r * H + h] = hi;
        if (y_lo) y_lo[r * H + h] = lo;
      }
    }
  return 0;
}
int embed_action_bwd(const float* dy, const float* y, const float* actions, int64_t R, int A, int H, int T, float* dW, float* db,
                     float* dE, stream_t) {
  if (A > 8) return set_error("embed_action_bwd: act_dim > 8 unsupported");
  for (int64_t r = 0; r < R; ++r)
    for (int h = 0; h < H; ++h) {
      const float yv = y[r * H + h];
      const float d = dy[r * H + h] * (1.f - yv * yv);
      db[h] += d;
      for (int a = 0; a < A; ++a) dW[(int64_t)h * A + a] += d * actions[r * A + a];
      if (dE) dE[(r % T) * H + h] += d;
    }
  return 0;
}

// ---- fused clip_grad_norm_ + Adam (CPU restatement of videocad_b200/csrc/optim.cu)
size_t clip_adam_scratch_floats() { return (size_t)VC_ADAM_MAX_TENSORS * 1024 + 8; }
int clip_adam_step(const vc_adam_tensor* t, int nt, double beta1, double beta2, double eps, double max_norm, int64_t step, float* scratch,
                   float* total_norm_out, stream_t) {
  if (!t || nt <= 0 || nt > VC_ADAM_MAX_TENSORS || step < 1 || !scratch) return set_error("clip_adam_step: bad arguments");
  double ss = 0;
  for (int i = 0; i < nt; ++i)
    for (int64_t k = 0; k < t[i].n; ++k) ss += (double)t[i].g[k] * t[i].g[k];
  const float total = (float)sqrt(ss);
  float coef = 1.f;
  if (max_norm > 0) coef = fminf((float)max_norm / (total + 1e-6f), 1.f);
  if (total_norm_out) total_norm_out[0] = total;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  const float b2 = (float)beta2, omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2), e = (float)eps, bc2s = (float)sqrt(bc2);
  for (int i = 0; i < nt; ++i) {
    const float step_size = (float)((double)t[i].lr / bc1);
    for (int64_t k = 0; k < t[i].n; ++k) {
      float g = t[i].g[k] * coef, m = t[i].m[k], v = t[i].v[k];
      m = m + omb1 * (g - m);
      v = v * b2 + omb2 * g * g;
      const float denom = sqrtf(v) / bc2s + e;
      t[i].p[k] = t[i].p[k] - step_size * (m / denom);
      t[i].g[k] = g; t[i].m[k] = m; t[i].v[k] = v;
    }
  }
  return 0;
}

// ---- fused training loss (CPU restatement of videocad_b200/csrc/loss.cu; same workspace layout)
size_t loss_workspace_floats(int R, int NP) { return (size_t)5 * R * NP + (size_t)4 * R + 16; }

namespace {
struct LossWsC {
  float *lse, *sel, *rowloss, *nwc, *cnum, *cden, *clse, *scal, *pred, *cpred;
};
LossWsC loss_carve(float* ws, int R, int NP) {
  LossWsC w;
  const size_t rp = (size_t)R * NP;
  w.lse = ws; w.sel = ws + rp; w.rowloss = ws + 2 * rp; w.nwc = ws + 3 * rp;
  w.cnum = ws + 4 * rp; w.cden = w.cnum + R; w.clse = w.cden + R; w.scal = w.clse + R;
  w.pred = w.scal + 16; w.cpred = w.pred + rp;
  return w;
}
void loss_window(const LossCfg& c, int i, float target, bool& valid, long long& t, long long& hi, float& count) {
  const long long tg = (long long)target;
  valid = tg != -1;
  // allowed classes = { clamp(target + o, 0, NV - 1) : 0 <= o < tolerance } (trainer.py:880-905): a contiguous window whose two
  // ends are clamped separately, so an out-of-range target still has a one-class window (never empty, count >= 1)
  const long long top = c.NV - 1;
  t = tg < 0 ? 0 : (tg > top ? top : tg);
  hi = tg + (c.tolerance[i] - 1);
  hi = hi < 0 ? 0 : (hi > top ? top : hi);
  count = (float)(hi - t + 1);
}
}  // namespace

int loss_forward(const LossCfg& c, const float* cmds, const float* params, const float* targets, float* ws, float* loss_out, stream_t) {
  if (c.NP > VC_LOSS_MAX_PARAMS || c.NC > VC_LOSS_MAX_CLASSES || c.R <= 0) return set_error("loss: unsupported configuration");
  LossWsC w = loss_carve(ws, c.R, c.NP);
  const int ldt = 1 + c.NP;
  for (int r = 0; r < c.R; ++r) {
    const float* z = cmds + (size_t)r * c.NC;
    const long long tg = (long long)targets[(size_t)r * ldt];
    const bool valid = tg >= 0 && tg < c.NC;
    float m = -INFINITY, s = 0.f;
    int am = 0;
    for (int k = 0; k < c.NC; ++k) if (z[k] > m) { m = z[k]; am = k; }
    w.cpred[r] = (float)am;
    for (int k = 0; k < c.NC; ++k) s += expf(z[k] - m);
    const float lse = m + logf(s), wt = valid ? c.cmd_w[tg] : 0.f;
    w.clse[r] = lse; w.cden[r] = wt; w.cnum[r] = valid ? wt * (lse - z[tg]) : 0.f;
    for (int i = 0; i < c.NP; ++i) {
      const float* zp = params + ((size_t)r * c.NP + i) * c.NV;
      bool v; long long t, hi; float count;
      loss_window(c, i, targets[(size_t)r * ldt + 1 + i], v, t, hi, count);
      float best = -INFINITY; int bi = 0;
      for (int k = 0; k < c.NV; ++k) if (zp[k] > best) { best = zp[k]; bi = k; }
      double se = 0, sw = 0; float nw = 0.f;
      for (int k = 0; k < c.NV; ++k) {
        se += exp((double)zp[k] - best);
        if (k >= t && k <= hi) { sw += zp[k]; nw += 1.f; }
      }
      const float l = best + (float)log(se);
      const float sel = (v && !(bi >= t && bi <= hi)) ? 1.f : 0.f;
      const size_t o = (size_t)r * c.NP + i;
      w.lse[o] = l; w.sel[o] = sel;
      w.rowloss[o] = (-((float)sw - nw * l) / count) * sel;
      w.nwc[o] = nw / count;
      w.pred[o] = (float)bi;
    }
  }
  float loss = 0.f;
  for (int i = 0; i < c.NP; ++i) {
    double S = 0, n = 0;
    for (int r = 0; r < c.R; ++r) { S += w.rowloss[(size_t)r * c.NP + i]; n += w.sel[(size_t)r * c.NP + i]; }
    const float denom = n > 1.0 ? (float)n : 1.f;
    float term = (float)S / denom, coef = c.cmd_w[c.param_to_label[i]] / denom;
    if (term != term) { term = 0.f; coef = 0.f; }
    w.scal[i] = coef;
    loss += term * c.cmd_w[c.param_to_label[i]];
  }
  double num = 0, den = 0;
  for (int r = 0; r < c.R; ++r) { num += w.cnum[r]; den += w.cden[r]; }
  loss += 2.f * (float)(num / den);
  w.scal[c.NP] = 2.f / (float)den;
  loss_out[0] = loss;
  return 0;
}

int loss_metrics(const LossCfg& c, const vc_metrics_cfg& mc, const float* targets, const float* ws, int T, int64_t* counts, stream_t) {
  if (!targets || !ws || !counts) return set_error("loss_metrics: null argument");
  if (T <= 0 || c.R % T != 0) return set_error("loss_metrics: R must be a multiple of T");
  LossWsC w = loss_carve(const_cast<float*>(ws), c.R, c.NP);
  for (int k = 0; k < VC_METRIC_COUNT; ++k) counts[k] = 0;
  const int ldt = 1 + c.NP;
  for (int r = 0; r < c.R; ++r) {
    const long long tc = (long long)targets[(size_t)r * ldt], pc = (long long)w.cpred[r];
    const bool cmd_mask = tc != -1, cmd_ok = cmd_mask && pc == tc, in_topk = (r % T) < mc.topk;
    counts[VC_METRIC_TOTAL] += cmd_mask;
    counts[VC_METRIC_CORRECT] += cmd_ok;
    if (tc >= 0 && tc < c.NC) { counts[VC_METRIC_CMD_COUNTS + tc] += 1; counts[VC_METRIC_CMD_CORRECTS + tc] += (pc == tc); }
    counts[VC_METRIC_CMD_COUNTS_TOPK] += (in_topk && cmd_mask);
    counts[VC_METRIC_CMD_CORRECT_TOPK] += (in_topk && cmd_ok);
    for (int i = 0; i < c.NP; ++i) {
      const long long tp = (long long)targets[(size_t)r * ldt + 1 + i];
      if (!cmd_mask || tp == -1) continue;
      counts[VC_METRIC_PARAM_COUNTS + i] += 1;
      counts[VC_METRIC_TOTAL] += 1;
      counts[VC_METRIC_PARAM_COUNTS_TOPK] += in_topk;
      if (!cmd_ok) continue;
      const long long diff = (long long)w.pred[(size_t)r * c.NP + i] - tp;
      const bool ok = mc.above[i] ? (diff >= 0 && diff < mc.tolerance[i]) : ((diff < 0 ? -diff : diff) < mc.abs_tolerance);
      if (ok) { counts[VC_METRIC_PARAM_CORRECTS + i] += 1; counts[VC_METRIC_CORRECT] += 1; counts[VC_METRIC_PARAM_CORRECT_TOPK] += in_topk; }
    }
  }
  return 0;
}

int loss_backward(const LossCfg& c, const float* cmds, const float* params, const float* targets, const float* ws, const float* upstream,
                  float* dcmds, float* dparams, stream_t) {
  LossWsC w = loss_carve(const_cast<float*>(ws), c.R, c.NP);
  const int ldt = 1 + c.NP;
  const float up = upstream[0];
  for (int r = 0; r < c.R; ++r) {
    const float* z = cmds + (size_t)r * c.NC;
    const long long tg = (long long)targets[(size_t)r * ldt];
    const float wt = w.cden[r];
    for (int k = 0; k < c.NC; ++k)
      dcmds[(size_t)r * c.NC + k] = wt != 0.f ? wt * w.scal[c.NP] * up * (expf(z[k] - w.clse[r]) - (k == tg ? 1.f : 0.f)) : 0.f;
    for (int i = 0; i < c.NP; ++i) {
      const size_t o = (size_t)r * c.NP + i;
      const float* zp = params + o * c.NV;
      float* dz = dparams + o * c.NV;
      const float coef = w.scal[i] * w.sel[o] * up;
      if (coef == 0.f) { for (int k = 0; k < c.NV; ++k) dz[k] = 0.f; continue; }
      bool v; long long t, hi; float count;
      loss_window(c, i, targets[(size_t)r * ldt + 1 + i], v, t, hi, count);
      for (int k = 0; k < c.NV; ++k)
        dz[k] = coef * (w.nwc[o] * expf(zp[k] - w.lse[o]) - ((k >= t && k <= hi) ? 1.f / count : 0.f));
    }
  }
  return 0;
}

int head_small_fwd(const float* x, int64_t R, int H, const float* W, const float* b, int C, float* out, stream_t) {
  if (C > 8 || H % 4 != 0) return set_error("head_small_fwd: C <= 8 and H % 4 == 0 required");
  for (int64_t r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) {
      double acc = b[c];
      for (int h = 0; h < H; ++h) acc += (double)x[r * H + h] * W[(int64_t)c * H + h];
      out[r * C + c] = (float)acc;
    }
  return 0;
}
int head_small_bwd(const float* dout, const float* x, int64_t R, int H, const float* W, int C, float* dx, int accumulate_dx,
                   float* dW, float* db, stream_t) {
  if (C > 8 || H % 4 != 0) return set_error("head_small_bwd: C <= 8 and H % 4 == 0 required");
  for (int64_t r = 0; r < R; ++r) {
    for (int h = 0; h < H; ++h) {
      double o = accumulate_dx ? dx[r * H + h] : 0.0;
      for (int c = 0; c < C; ++c) o += (double)dout[r * C + c] * W[(int64_t)c * H + h];
      dx[r * H + h] = (float)o;
    }
    for (int c = 0; c < C; ++c) {
      db[c] += dout[r * C + c];
      for (int h = 0; h < H; ++h) dW[(int64_t)c * H + h] += dout[r * C + c] * x[r * H + h];
    }
  }
  return 0;
}

// ---- decode-step kernels (CPU twins of videocad_b200/csrc/decode.cu; contracts in kernels.h) ----
int dec_gemv(const DecGemv& g, stream_t) {
  if (g.M <= 0 || g.M > 16 || g.K <= 0 || g.N <= 0 || !g.W || !g.out) return set_error("dec_gemv: bad arguments");
  const int M = g.M, K = g.K, t = g.t_ptr ? *g.t_ptr : 0;
  std::vector<double> x((size_t)M * K, 0.0);
  for (int m = 0; m < M; ++m) {
    double* xr = x.data() + (size_t)m * K;
    if (g.in_mode == VC_DEC_IN_PLAIN) {
      for (int k = 0; k < K; ++k) xr[k] = g.x[(size_t)m * g.ldx + k];
    } else if (g.in_mode == VC_DEC_IN_LN) {
      double mean = 0.0, var = 0.0;
      for (int k = 0; k < K; ++k) mean += g.x[(size_t)m * g.ldx + k];
      mean /= K;
      for (int k = 0; k < K; ++k) { const double d = g.x[(size_t)m * g.ldx + k] - mean; var += d * d; }
      const double rstd = 1.0 / sqrt(var / K + 1e-5);
      for (int k = 0; k < K; ++k) xr[k] = (g.x[(size_t)m * g.ldx + k] - mean) * rstd * g.gamma[k] + g.beta[k];
    } else if (g.in_mode == VC_DEC_IN_EMBED) {
      for (int k = 0; k < K; ++k) {
        double acc = g.emb_b[k];
        for (int i = 0; i < g.act_dim; ++i) acc += (double)g.actions[(size_t)m * g.act_dim + i] * g.emb_W[(size_t)k * g.act_dim + i];
        if (g.emb_E) acc += g.emb_E[(size_t)t * K + k];
        xr[k] = tanh(acc);
      }
    }
    if (g.x_out) for (int k = 0; k < K; ++k) g.x_out[(size_t)m * K + k] = (float)xr[k];
  }
#pragma omp parallel for
  for (int n = 0; n < g.N; ++n) {
    for (int m = 0; m < M; ++m) {
      double acc = 0.0;
      const float* wr = g.W + (size_t)n * K;
      const double* xr = x.data() + (size_t)m * K;
      for (int k = 0; k < K; ++k) acc += xr[k] * wr[k];
      float v = (float)(acc + (g.bias ? g.bias[n] : 0.f));
      if (g.act == VC_ACT_RELU) v = v > 0.f ? v : 0.f;
      else if (g.act == VC_ACT_TANH) v = tanhf(v);
      else if (g.act == VC_ACT_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
      if (g.residual) v += g.residual[(size_t)m * g.ld_res + n];
      g.out[(size_t)m * g.out_row_stride + (size_t)t * g.out_t_stride + n] = v;
    }
  }
  return 0;
}

int dec_attn(const DecAttn& a, int B, stream_t) {
  if (B <= 0 || !a.q || !a.k || !a.v || !a.out || !a.t_ptr || a.nsplit < 1) return set_error("dec_attn: bad arguments");
  const int t = *a.t_ptr;
  const int j_lo = a.window > 0 ? (t - a.window + 1 > 0 ? t - a.window + 1 : 0) : 0;
  const int nkeys = t - j_lo + 1;
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < a.nh; ++h) {
      const float* q = a.q + (size_t)b * a.q_bstride + (size_t)t * a.q_tstride + h * a.dh;
      std::vector<double> sc(nkeys);
      double mx = -INFINITY, l = 0.0;
      for (int i = 0; i < nkeys; ++i) {
        const float* kr = a.k + (size_t)b * a.kv_bstride + (size_t)(j_lo + i) * a.kv_rstride + h * a.dh;
        double d = 0.0;
        for (int e = 0; e < a.dh; ++e) d += (double)q[e] * a.scale * kr[e];
        sc[i] = d;
        mx = fmax(mx, d);
      }
      std::vector<double> o(a.dh, 0.0);
      for (int i = 0; i < nkeys; ++i) {
        const double p = exp(sc[i] - mx);
        l += p;
        const float* vr = a.v + (size_t)b * a.kv_bstride + (size_t)(j_lo + i) * a.kv_rstride + h * a.dh;
        for (int e = 0; e < a.dh; ++e) o[e] += p * vr[e];
      }
      for (int e = 0; e < a.dh; ++e) a.out[(size_t)b * a.ld_out + h * a.dh + e] = (float)(o[e] / l);
    }
  return 0;
}

int dec_select(const DecSelect& a, stream_t) {
  static const int kMask[5][6] = {{1, 1, 0, 0, 0, 0}, {0, 0, 1, 1, 0, 0}, {0, 0, 0, 0, 1, 0}, {0, 0, 0, 0, 0, 1}, {0, 0, 0, 0, 0, 0}};
  if (a.B <= 0 || a.NC < 1 || a.NC > 8 || a.NPAR > 8 || !a.y || !a.t_ptr) return set_error("dec_select: bad arguments");
  const int t = *a.t_ptr, H = a.H;
  for (int b = 0; b < a.B; ++b) {
    const float* y = a.y + (size_t)b * H;
    double mean = 0.0, var = 0.0;
    for (int k = 0; k < H; ++k) mean += y[k];
    mean /= H;
    for (int k = 0; k < H; ++k) { const double d = y[k] - mean; var += d * d; }
    const double rstd = 1.0 / sqrt(var / H + 1e-5);
    int cmd = 0;
    float best = -INFINITY;
    for (int c = 0; c < a.NC; ++c) {
      double acc = a.bc[c];
      for (int k = 0; k < H; ++k) acc += ((y[k] - mean) * rstd * a.gamma[k] + a.beta[k]) * a.Wc[(size_t)c * H + k];
      const float v = (float)acc;
      a.cmds_all[((size_t)b * a.T + t) * a.NC + c] = v;
      if (v > best) { best = v; cmd = c; }
    }
    if (!a.action_next) continue;
    float par[8];
    for (int i = 0; i < a.NPAR; ++i) {
      const float* z = a.params_all + ((size_t)b * a.T + t) * ((size_t)a.NPAR * a.NV) + (size_t)i * a.NV;
      int bi = 0;
      for (int k = 1; k < a.NV; ++k) if (z[k] > z[bi]) bi = k;
      par[i] = (cmd < 5 && i < 6 && kMask[cmd][i]) ? (float)bi : -1.0f;
    }
    if (a.NPAR > 3 && !(par[2] >= 200.f && par[2] < 250.f)) par[3] = -1.0f;
    float* out = a.action_next + (size_t)b * (1 + a.NPAR);
    out[0] = (float)cmd / 4.0f;
    for (int i = 0; i < a.NPAR; ++i) out[1 + i] = par[i] / 1000.0f;
  }
  *a.t_ptr = t + 1;
  return 0;
}

int stream_fork(stream_t main, int, stream_t* side) { *side = main; return 0; }
int stream_join(stream_t, int) { return 0; }
void side_streams_enable(int) {}

int add_f32(const float* a, const float* b, float* out, int64_t n, stream_t) {
  for (int64_t i = 0; i < n; ++i) out[i] = a[i] + b[i];
  return 0;
}
int zero_f32(float* x, int64_t n, stream_t) {
  if (n > 0) memset(x, 0, (size_t)n * sizeof(float));
  return 0;
}
int dropout_mask_debug(Drop drop, int64_t n, float* out, stream_t) {
  for (int64_t i = 0; i < n; ++i) out[i] = (drop.p > 0.f) ? keep_scale(drop, (uint64_t)i) : 1.0f;
  return 0;
}

}  // namespace vck
