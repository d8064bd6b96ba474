// Covariance-tile generator: device functions that evaluate the composite kernels of
// loadest-gp / rating-gp (reference: src/loadest_gp/models/gpytorch.py:61-128,
// src/rating_gp/models/gpytorch.py:205-372, src/rating_gp/models/kernels.py:242-382) and their
// derivatives w.r.t. the natural hyper-parameters one matrix entry at a time,
// from per-point feature rows held in shared memory.  Nothing here touches global memory except
// cov_compile(), which folds theta into a shared-memory parameter block once per CTA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dgp.h"

#define DGP_XS 8  // doubles per feature-table row (64 B)
// internal feature-table column kinds appended by libdgp (never set by callers): sinpi / cospi of (feature column `src`)
// / theta[`theta`] for a periodic factor, so that sin(pi (xa - xb) / p) = sa cb - ca sb costs two FMAs per entry
// instead of a sincospi.  dgp_factor::pad_ of the internal spec copy holds 1 + the sin column (0: none).
#define DGP_COL_SINP 100
#define DGP_COL_COSP 101

namespace dgp {

constexpr int DGP_BATCH_MAX = 32;  // sites per batched launch (dgp_batch_*)

struct FactorC {
  int kind, ndims;
  int col[DGP_MAX_FDIMS];
  int ls_idx[DGP_MAX_FDIMS];
  int period_idx, sc_col;  // sc_col: feature column of sinpi(x / p) (cospi in the next one), -1: evaluate directly
  double inv_ls[DGP_MAX_FDIMS];
  double w;        // periodic: pi / period
  double inv_lam;  // periodic: 1 / lengthscale
  double inv_p;    // periodic: 1 / period
};

struct TermC {
  double scale;
  double gate_a;
  int scale_idx, gate, gate_col, gate_theta, nf, shape;  // shape: SH_* code of (factor kinds, dimensions), 0 = generic
  FactorC f[DGP_MAX_FACTORS];
};

struct CovC {
  int nterms, ntheta, noise_idx, pad_;
  double extra_noise;  // theta[noise_theta] (0 when absent) + jitter
  TermC t[DGP_MAX_TERMS];
};

// slots of the per-term derivative accumulator
constexpr int SLOT_SCALE = 0;
constexpr int SLOT_GATE = 1;
constexpr int SLOT_F0 = 2;                       // + f * SLOT_PER_F + d  (d < 4: lengthscale d; d == 4: period)
constexpr int SLOT_PER_F = DGP_MAX_FDIMS + 1;
constexpr int NSLOT = SLOT_F0 + DGP_MAX_FACTORS * SLOT_PER_F;  // 17

// Term shapes with straight-line evaluators.  The V-wide evaluators below are written once, against a shape class:
// ShapeRT reads the factor kinds / dimension counts from the term at run time (an interpreter: ~70 % of its
// instructions are loop and branch bookkeeping), Shape<...> fixes them at compile time so that the same code folds to
// the arithmetic of that one product of factors.  cov_compile() tags every term; unknown products stay generic.
enum { SH_GENERIC = 0, SH_PER_M52 = 1, SH_RBF1 = 2, SH_M32_2 = 3, SH_M52_1 = 4, SH_M52_M32 = 5, SH_M52_M52 = 6, SH_RBF2 = 7,
       SH_M32_1 = 8, SH_M52_2 = 9 };

struct ShapeRT {
  static constexpr bool fixed = false;
  static constexpr int nf = 0, k0 = 0, n0 = 0, k1 = 0, n1 = 0;
};
template <int NF, int K0, int N0, int K1 = 0, int N1 = 0>
struct Shape {
  static constexpr bool fixed = true;
  static constexpr int nf = NF, k0 = K0, n0 = N0, k1 = K1, n1 = N1;
};
template <class S> __device__ __forceinline__ int sh_nf(const TermC& tc) { if constexpr (S::fixed) return S::nf; else return tc.nf; }
template <class S> __device__ __forceinline__ int sh_kind(const TermC& tc, int f) {
  if constexpr (S::fixed) return f == 0 ? S::k0 : S::k1; else return tc.f[f].kind;
}
template <class S> __device__ __forceinline__ int sh_nd(const TermC& tc, int f) {
  if constexpr (S::fixed) return f == 0 ? S::n0 : (f == 1 ? S::n1 : 0); else return tc.f[f].ndims;
}

__device__ __forceinline__ int term_shape(const dgp_term& st) {
  const int nf = st.nfactors;
  const dgp_factor& a = st.factor[0];
  const dgp_factor& b = st.factor[1];
  if (nf == 1) {
    if (a.kind == DGP_RBF && a.ndims == 1) return SH_RBF1;
    if (a.kind == DGP_RBF && a.ndims == 2) return SH_RBF2;
    if (a.kind == DGP_MATERN32 && a.ndims == 1) return SH_M32_1;
    if (a.kind == DGP_MATERN32 && a.ndims == 2) return SH_M32_2;
    if (a.kind == DGP_MATERN52 && a.ndims == 1) return SH_M52_1;
    if (a.kind == DGP_MATERN52 && a.ndims == 2) return SH_M52_2;
  } else if (nf == 2) {
    if (a.kind == DGP_PERIODIC && b.kind == DGP_MATERN52 && b.ndims == 1) return SH_PER_M52;
    if (a.kind == DGP_MATERN52 && a.ndims == 1 && b.kind == DGP_MATERN32 && b.ndims == 1) return SH_M52_M32;
    if (a.kind == DGP_MATERN52 && a.ndims == 1 && b.kind == DGP_MATERN52 && b.ndims == 1) return SH_M52_M52;
  }
  return SH_GENERIC;
}

// Fold (spec, theta) into shared memory.  Call with all threads of the CTA; ends with a barrier
// only if `sync` is set (callers that have their own barrier pass false).
__device__ __forceinline__ void cov_compile(CovC* cc, const dgp_spec& sp, const double* __restrict__ theta,
                                            double jitter, int tid, int nthreads) {
  for (int t = tid; t < DGP_MAX_TERMS; t += nthreads) {
    TermC& tc = cc->t[t];
    if (t < sp.nterms) {
      const dgp_term& st = sp.term[t];
      tc.scale_idx = st.scale;
      tc.scale = st.scale >= 0 ? theta[st.scale] : 1.0;
      tc.gate = st.gate;
      tc.gate_col = st.gate_col;
      tc.gate_theta = -1;
      tc.gate_a = 0.0;
      if (st.gate != DGP_GATE_NONE) {
        tc.gate_theta = sp.col[st.gate_col].theta;
        tc.gate_a = sp.col[st.gate_col].aux;
      }
      tc.nf = st.nfactors;
      tc.shape = term_shape(st);
      for (int f = 0; f < DGP_MAX_FACTORS; f++) {
        FactorC& fc = tc.f[f];
        const dgp_factor& sf = st.factor[f];
        fc.kind = sf.kind;
        fc.ndims = f < st.nfactors ? sf.ndims : 0;
        fc.period_idx = -1;
        fc.sc_col = -1;
        fc.w = 0.0; fc.inv_lam = 0.0; fc.inv_p = 0.0;
        for (int d = 0; d < DGP_MAX_FDIMS; d++) {
          fc.col[d] = 0; fc.ls_idx[d] = -1; fc.inv_ls[d] = 0.0;
          if (d < fc.ndims) {
            fc.col[d] = sf.col[d];
            fc.ls_idx[d] = sf.ls[d];
            fc.inv_ls[d] = 1.0 / theta[sf.ls[d]];
          }
        }
        if (f < st.nfactors && sf.kind == DGP_PERIODIC) {
          fc.period_idx = sf.period;
          fc.sc_col = sf.pad_ > 0 ? sf.pad_ - 1 : -1;
          fc.inv_p = 1.0 / theta[sf.period];
          fc.w = 3.14159265358979323846 * fc.inv_p;
          fc.inv_lam = fc.inv_ls[0];
        }
      }
    } else {
      tc.nf = 0; tc.scale = 0.0; tc.scale_idx = -1; tc.gate = 0; tc.shape = SH_GENERIC;
    }
  }
  if (tid == 0) {
    cc->nterms = sp.nterms;
    cc->ntheta = sp.ntheta;
    cc->noise_idx = sp.noise_theta;
    cc->extra_noise = (sp.noise_theta >= 0 ? theta[sp.noise_theta] : 0.0) + jitter;
  }
}

// value of one stationary factor
__device__ __forceinline__ double factor_val(const FactorC& f, const double* xi, const double* xj) {
  if (f.kind == DGP_PERIODIC) {
    const double u = (xi[f.col[0]] - xj[f.col[0]]) * f.w;
    const double s = sin(u);
    return exp(-2.0 * s * s * f.inv_lam);
  }
  double d2 = 0.0;
  for (int d = 0; d < f.ndims; d++) {
    const double z = (xi[f.col[d]] - xj[f.col[d]]) * f.inv_ls[d];
    d2 = fma(z, z, d2);
  }
  if (f.kind == DGP_RBF) return exp(-0.5 * d2);
  const double r = sqrt(d2);
  if (f.kind == DGP_MATERN32) {
    const double a = 1.7320508075688772 * r;
    return (1.0 + a) * exp(-a);
  }
  const double a = 2.23606797749979 * r;  // MATERN52
  return (1.0 + a + a * a * (1.0 / 3.0)) * exp(-a);
}

// value of one covariance entry (no noise)
__device__ __forceinline__ double cov_entry(const CovC* cc, const double* xi, const double* xj) {
  double k = 0.0;
  for (int t = 0; t < cc->nterms; t++) {
    const TermC& tc = cc->t[t];
    double v = tc.scale;
    if (tc.gate != DGP_GATE_NONE) {
      double gi = xi[tc.gate_col], gj = xj[tc.gate_col];
      if (tc.gate == DGP_GATE_INV_SIGMOID) { gi = 1.0 - gi; gj = 1.0 - gj; }
      v *= gi * gj;
    }
    for (int f = 0; f < tc.nf; f++) v *= factor_val(tc.f[f], xi, xj);
    k += v;
  }
  return k;
}

// value of factor f and d(value)/d(param) for its parameters: dls[d] (lengthscales), dper (period)
__device__ __forceinline__ double factor_val_grad(const FactorC& f, const double* xi, const double* xj,
                                                  double (&dls)[DGP_MAX_FDIMS], double& dper) {
  dper = 0.0;
#pragma unroll
  for (int d = 0; d < DGP_MAX_FDIMS; d++) dls[d] = 0.0;
  if (f.kind == DGP_PERIODIC) {
    const double u = (xi[f.col[0]] - xj[f.col[0]]) * f.w;
    double s, c;
    sincos(u, &s, &c);
    const double val = exp(-2.0 * s * s * f.inv_lam);
    dls[0] = val * 2.0 * s * s * f.inv_lam * f.inv_lam;            // d/d lam
    dper = val * (4.0 * f.inv_lam) * s * c * u * f.inv_p;         // d/d period = val*(2/lam)*sin(2u)*u/p
    return val;
  }
  double d2 = 0.0;
  double z2[DGP_MAX_FDIMS];
#pragma unroll
  for (int d = 0; d < DGP_MAX_FDIMS; d++) {
    z2[d] = 0.0;
    if (d < f.ndims) {
      const double z = (xi[f.col[d]] - xj[f.col[d]]) * f.inv_ls[d];
      z2[d] = z * z;
      d2 += z2[d];
    }
  }
  double val, common;  // d val / d ls_d = common * z_d^2 / ls_d
  if (f.kind == DGP_RBF) {
    val = exp(-0.5 * d2);
    common = val;
  } else {
    const double r = sqrt(d2);
    if (f.kind == DGP_MATERN32) {
      const double a = 1.7320508075688772 * r;
      const double e = exp(-a);
      val = (1.0 + a) * e;
      common = 3.0 * e;
    } else {
      const double a = 2.23606797749979 * r;
      const double e = exp(-a);
      val = (1.0 + a + a * a * (1.0 / 3.0)) * e;
      common = (5.0 / 3.0) * (1.0 + a) * e;
    }
  }
#pragma unroll
  for (int d = 0; d < DGP_MAX_FDIMS; d++)
    if (d < f.ndims) dls[d] = common * z2[d] * f.inv_ls[d];
  return val;
}

// Accumulate w * d(term)/d(param) into acc[NSLOT] for one entry.  Returns nothing; the caller
// reduces acc over the tile and scatters slots to theta indices with term_scatter().
__device__ __forceinline__ void term_grad_accum(const TermC& tc, const double* xi, const double* xj, double w,
                                                double (&acc)[NSLOT]) {
  double G = 1.0, dG = 0.0;
  if (tc.gate != DGP_GATE_NONE) {
    double gi = xi[tc.gate_col], gj = xj[tc.gate_col];
    double sgn = tc.gate_a;
    if (tc.gate == DGP_GATE_INV_SIGMOID) { gi = 1.0 - gi; gj = 1.0 - gj; sgn = -sgn; }
    G = gi * gj;
    dG = sgn * G * (2.0 - gi - gj);
  }
  double fv[DGP_MAX_FACTORS];
  double dls[DGP_MAX_FACTORS][DGP_MAX_FDIMS];
  double dper[DGP_MAX_FACTORS];
  double F = 1.0;
#pragma unroll
  for (int f = 0; f < DGP_MAX_FACTORS; f++) {
    fv[f] = 1.0; dper[f] = 0.0;
#pragma unroll
    for (int d = 0; d < DGP_MAX_FDIMS; d++) dls[f][d] = 0.0;
    if (f < tc.nf) {
      fv[f] = factor_val_grad(tc.f[f], xi, xj, dls[f], dper[f]);
      F *= fv[f];
    }
  }
  acc[SLOT_SCALE] = fma(w, G * F, acc[SLOT_SCALE]);
  const double ws = w * tc.scale;
  acc[SLOT_GATE] = fma(ws, dG * F, acc[SLOT_GATE]);
  const double wsg = ws * G;
#pragma unroll
  for (int f = 0; f < DGP_MAX_FACTORS; f++) {
    if (f < tc.nf) {
      double others = 1.0;
#pragma unroll
      for (int g = 0; g < DGP_MAX_FACTORS; g++)
        if (g != f) others *= fv[g];
      const double wo = wsg * others;
#pragma unroll
      for (int d = 0; d < DGP_MAX_FDIMS; d++)
        acc[SLOT_F0 + f * SLOT_PER_F + d] = fma(wo, dls[f][d], acc[SLOT_F0 + f * SLOT_PER_F + d]);
      acc[SLOT_F0 + f * SLOT_PER_F + DGP_MAX_FDIMS] = fma(wo, dper[f], acc[SLOT_F0 + f * SLOT_PER_F + DGP_MAX_FDIMS]);
    }
  }
}

// ------------------------------------------------------------------ V-wide evaluators (the fast path)
// One "own" point A (per-thread, features read from a column-major shared array xaT[col * a_stride + a_idx]) against
// V "other" points B (the same for every thread of the warp: broadcast reads of row-major rows xb + v * DGP_XS).
// All the exponentials of a term are merged, term = scale * gate * prod_f poly_f * exp(-sum_f expo_f):
//   RBF: poly 1, expo d2/2 | Matern32: poly 1 + a, expo a = sqrt(3 d2) | Matern52: poly 1 + a + a^2/3, expo a = sqrt(5 d2)
//   Periodic: poly 1, expo 2 sin^2(pi dx / p) / lam   (sinpi: exact range reduction, no slow path)
// and the V entries are evaluated side by side so that the FP64 pipe sees V independent dependency chains.
__device__ __forceinline__ double fast_sqrt_nonneg(double d2) {
  return d2 > 0.0 ? d2 * rsqrt(d2) : d2;   // d2 = 0 -> 0; a NaN (poisoned length scale) stays a NaN instead of turning into 0
}

// out[v] += value of term tc for the V entries
template <int V, class S>
__device__ __forceinline__ void term_vals_v(const TermC& tc, const double* __restrict__ xaT, int a_stride, int a_idx,
                                            const double* __restrict__ xb, double (&out)[V]) {
  double P[V], EX[V];
#pragma unroll
  for (int v = 0; v < V; v++) { P[v] = tc.scale; EX[v] = 0.0; }
  if (tc.gate != DGP_GATE_NONE) {
    double ga = xaT[tc.gate_col * a_stride + a_idx];
    if (tc.gate == DGP_GATE_INV_SIGMOID) ga = 1.0 - ga;
#pragma unroll
    for (int v = 0; v < V; v++) {
      double gb = xb[v * DGP_XS + tc.gate_col];
      if (tc.gate == DGP_GATE_INV_SIGMOID) gb = 1.0 - gb;
      P[v] *= ga * gb;
    }
  }
#pragma unroll
  for (int f = 0; f < DGP_MAX_FACTORS; f++) {
    if (f < sh_nf<S>(tc)) {
      const FactorC& fc = tc.f[f];
      const int kind = sh_kind<S>(tc, f);
      if (kind == DGP_PERIODIC) {
        if (fc.sc_col >= 0) {  // sin(pi (xa - xb) / p) from the per-point sin / cos columns
          const double sa = xaT[fc.sc_col * a_stride + a_idx], ca = xaT[(fc.sc_col + 1) * a_stride + a_idx];
          const double c2 = 2.0 * fc.inv_lam;
#pragma unroll
          for (int v = 0; v < V; v++) {
            const double sn = fma(sa, xb[v * DGP_XS + fc.sc_col + 1], -ca * xb[v * DGP_XS + fc.sc_col]);
            EX[v] = fma(c2 * sn, sn, EX[v]);
          }
        } else {
          const double xa = xaT[fc.col[0] * a_stride + a_idx];
#pragma unroll
          for (int v = 0; v < V; v++) {
            const double sn = sinpi((xa - xb[v * DGP_XS + fc.col[0]]) * fc.inv_p);
            EX[v] = fma(2.0 * fc.inv_lam * sn, sn, EX[v]);
          }
        }
      } else {
        double d2[V];
#pragma unroll
        for (int v = 0; v < V; v++) d2[v] = 0.0;
#pragma unroll
        for (int d = 0; d < DGP_MAX_FDIMS; d++) {
          if (d < sh_nd<S>(tc, f)) {
            const double xa = xaT[fc.col[d] * a_stride + a_idx];
#pragma unroll
            for (int v = 0; v < V; v++) {
              const double z = (xa - xb[v * DGP_XS + fc.col[d]]) * fc.inv_ls[d];
              d2[v] = fma(z, z, d2[v]);
            }
          }
        }
        if (kind == DGP_RBF) {
#pragma unroll
          for (int v = 0; v < V; v++) EX[v] = fma(0.5, d2[v], EX[v]);
        } else if (kind == DGP_MATERN32) {
#pragma unroll
          for (int v = 0; v < V; v++) {
            const double a = 1.7320508075688772 * fast_sqrt_nonneg(d2[v]);
            P[v] *= 1.0 + a;
            EX[v] += a;
          }
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) {
            const double a = 2.23606797749979 * fast_sqrt_nonneg(d2[v]);
            P[v] *= fma(a, fma(a, 1.0 / 3.0, 1.0), 1.0);
            EX[v] += a;
          }
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < V; v++) out[v] = fma(P[v], exp(-EX[v]), out[v]);
}

// shape dispatch (uniform over the CTA): one switch per term and group of V entries
#define DGP_SHAPE_SWITCH(shape, CALL)                                             \
  switch (shape) {                                                                \
    case SH_PER_M52: { using S = Shape<2, DGP_PERIODIC, 1, DGP_MATERN52, 1>; CALL; } break; \
    case SH_RBF1: { using S = Shape<1, DGP_RBF, 1>; CALL; } break;                \
    case SH_RBF2: { using S = Shape<1, DGP_RBF, 2>; CALL; } break;                \
    case SH_M32_1: { using S = Shape<1, DGP_MATERN32, 1>; CALL; } break;          \
    case SH_M32_2: { using S = Shape<1, DGP_MATERN32, 2>; CALL; } break;          \
    case SH_M52_1: { using S = Shape<1, DGP_MATERN52, 1>; CALL; } break;          \
    case SH_M52_2: { using S = Shape<1, DGP_MATERN52, 2>; CALL; } break;          \
    case SH_M52_M32: { using S = Shape<2, DGP_MATERN52, 1, DGP_MATERN32, 1>; CALL; } break; \
    case SH_M52_M52: { using S = Shape<2, DGP_MATERN52, 1, DGP_MATERN52, 1>; CALL; } break; \
    default: { using S = ShapeRT; CALL; } break;                                  \
  }

template <int V>
__device__ __forceinline__ void cov_vals(const CovC* cc, const double* __restrict__ xaT, int a_stride, int a_idx,
                                         const double* __restrict__ xb, double (&out)[V]) {
#pragma unroll
  for (int v = 0; v < V; v++) out[v] = 0.0;
  for (int t = 0; t < cc->nterms; t++) {
    const TermC& tc = cc->t[t];
    DGP_SHAPE_SWITCH(tc.shape, (term_vals_v<V, S>(tc, xaT, a_stride, a_idx, xb, out)))
  }
}

// Accumulate w[v] * d(term)/d(param) over V entries into acc[NSLOT] (same slot layout as term_grad_accum).
template <int V, class S>
__device__ __forceinline__ void term_grad_accum_s(const TermC& tc, const double* __restrict__ xaT, int a_stride, int a_idx,
                                                  const double* __restrict__ xb, const double (&w)[V], double (&acc)[NSLOT]) {
  double G[V], dG[V], EX[V];
  double poly[DGP_MAX_FACTORS][V], aux[DGP_MAX_FACTORS][V];  // aux: Matern a | periodic sin*cos*u
#pragma unroll
  for (int v = 0; v < V; v++) { G[v] = 1.0; dG[v] = 0.0; EX[v] = 0.0; }
  if (tc.gate != DGP_GATE_NONE) {
    double ga = xaT[tc.gate_col * a_stride + a_idx];
    double sgn = tc.gate_a;
    if (tc.gate == DGP_GATE_INV_SIGMOID) { ga = 1.0 - ga; sgn = -sgn; }
#pragma unroll
    for (int v = 0; v < V; v++) {
      double gb = xb[v * DGP_XS + tc.gate_col];
      if (tc.gate == DGP_GATE_INV_SIGMOID) gb = 1.0 - gb;
      G[v] = ga * gb;
      dG[v] = sgn * G[v] * (2.0 - ga - gb);
    }
  }
  // pass 1: polynomial parts and the merged exponent
#pragma unroll
  for (int f = 0; f < DGP_MAX_FACTORS; f++) {
#pragma unroll
    for (int v = 0; v < V; v++) { poly[f][v] = 1.0; aux[f][v] = 0.0; }
    if (f < sh_nf<S>(tc)) {
      const FactorC& fc = tc.f[f];
      const int kind = sh_kind<S>(tc, f);
      if (kind == DGP_PERIODIC) {
        const double xa = xaT[fc.col[0] * a_stride + a_idx];
        double sa = 0.0, ca = 0.0;
        if (fc.sc_col >= 0) { sa = xaT[fc.sc_col * a_stride + a_idx]; ca = xaT[(fc.sc_col + 1) * a_stride + a_idx]; }
#pragma unroll
        for (int v = 0; v < V; v++) {
          const double x = (xa - xb[v * DGP_XS + fc.col[0]]) * fc.inv_p;
          double sn, cs;
          if (fc.sc_col >= 0) {
            const double sb = xb[v * DGP_XS + fc.sc_col], cb = xb[v * DGP_XS + fc.sc_col + 1];
            sn = fma(sa, cb, -ca * sb);
            cs = fma(ca, cb, sa * sb);
          } else {
            sincospi(x, &sn, &cs);
          }
          EX[v] = fma(2.0 * fc.inv_lam * sn, sn, EX[v]);
          aux[f][v] = sn * cs * x;  // u / pi, folded into the period slot below
          poly[f][v] = sn * sn;     // reused for the lam slot (the factor's polynomial part is 1)
        }
      } else {
        double d2[V];
#pragma unroll
        for (int v = 0; v < V; v++) d2[v] = 0.0;
#pragma unroll
        for (int d = 0; d < DGP_MAX_FDIMS; d++) {
          if (d < sh_nd<S>(tc, f)) {
            const double xa = xaT[fc.col[d] * a_stride + a_idx];
#pragma unroll
            for (int v = 0; v < V; v++) {
              const double z = (xa - xb[v * DGP_XS + fc.col[d]]) * fc.inv_ls[d];
              d2[v] = fma(z, z, d2[v]);
            }
          }
        }
        if (kind == DGP_RBF) {
#pragma unroll
          for (int v = 0; v < V; v++) EX[v] = fma(0.5, d2[v], EX[v]);
        } else if (kind == DGP_MATERN32) {
#pragma unroll
          for (int v = 0; v < V; v++) {
            const double a = 1.7320508075688772 * fast_sqrt_nonneg(d2[v]);
            poly[f][v] = 1.0 + a; aux[f][v] = a; EX[v] += a;
          }
        } else {
#pragma unroll
          for (int v = 0; v < V; v++) {
            const double a = 2.23606797749979 * fast_sqrt_nonneg(d2[v]);
            poly[f][v] = fma(a, fma(a, 1.0 / 3.0, 1.0), 1.0); aux[f][v] = a; EX[v] += a;
          }
        }
      }
    }
  }
  double E[V], F[V];
#pragma unroll
  for (int v = 0; v < V; v++) {
    E[v] = exp(-EX[v]);
    F[v] = E[v];
#pragma unroll
    for (int f = 0; f < DGP_MAX_FACTORS; f++)
      if (f < sh_nf<S>(tc) && sh_kind<S>(tc, f) != DGP_PERIODIC) F[v] *= poly[f][v];
  }
#pragma unroll
  for (int v = 0; v < V; v++) {
    acc[SLOT_SCALE] = fma(w[v], G[v] * F[v], acc[SLOT_SCALE]);
    acc[SLOT_GATE] = fma(w[v] * tc.scale, dG[v] * F[v], acc[SLOT_GATE]);
  }
  // pass 2: per-parameter derivative factors (the factor's own exponential is already inside E)
#pragma unroll
  for (int f = 0; f < DGP_MAX_FACTORS; f++) {
    if (f < sh_nf<S>(tc)) {
      const FactorC& fc = tc.f[f];
      const int kind = sh_kind<S>(tc, f);
      double wo[V];  // w * scale * gate * E * prod_{g != f} poly_g
#pragma unroll
      for (int v = 0; v < V; v++) {
        double o = w[v] * tc.scale * G[v] * E[v];
#pragma unroll
        for (int g = 0; g < DGP_MAX_FACTORS; g++)
          if (g != f && g < sh_nf<S>(tc) && sh_kind<S>(tc, g) != DGP_PERIODIC) o *= poly[g][v];
        wo[v] = o;
      }
      if (kind == DGP_PERIODIC) {
        const double cl = 2.0 * fc.inv_lam * fc.inv_lam;                        // d/d lam   : val * 2 s^2 / lam^2
        const double cp = 4.0 * 3.14159265358979323846 * fc.inv_lam * fc.inv_p;  // d/d period: val * (4/lam) s c u / p
#pragma unroll
        for (int v = 0; v < V; v++) {
          acc[SLOT_F0 + f * SLOT_PER_F + 0] = fma(wo[v] * cl, poly[f][v], acc[SLOT_F0 + f * SLOT_PER_F + 0]);
          acc[SLOT_F0 + f * SLOT_PER_F + DGP_MAX_FDIMS] = fma(wo[v] * cp, aux[f][v], acc[SLOT_F0 + f * SLOT_PER_F + DGP_MAX_FDIMS]);
        }
      } else {
        double cm[V];  // d factor / d ls_d = cm * z_d^2 / ls_d  (exponential excluded)
#pragma unroll
        for (int v = 0; v < V; v++)
          cm[v] = wo[v] * (kind == DGP_RBF ? 1.0 : (kind == DGP_MATERN32 ? 3.0 : (5.0 / 3.0) * (1.0 + aux[f][v])));
#pragma unroll
        for (int d = 0; d < DGP_MAX_FDIMS; d++) {
          if (d < sh_nd<S>(tc, f)) {
            const double xa = xaT[fc.col[d] * a_stride + a_idx];
            double sd = 0.0;
#pragma unroll
            for (int v = 0; v < V; v++) {
              const double z = (xa - xb[v * DGP_XS + fc.col[d]]) * fc.inv_ls[d];
              sd = fma(cm[v] * z, z, sd);
            }
            acc[SLOT_F0 + f * SLOT_PER_F + d] = fma(sd, fc.inv_ls[d], acc[SLOT_F0 + f * SLOT_PER_F + d]);
          }
        }
      }
    }
  }
}

template <int V>
__device__ __forceinline__ void term_grad_accum_v(const TermC& tc, const double* __restrict__ xaT, int a_stride, int a_idx,
                                                  const double* __restrict__ xb, const double (&w)[V], double (&acc)[NSLOT]) {
  DGP_SHAPE_SWITCH(tc.shape, (term_grad_accum_s<V, S>(tc, xaT, a_stride, a_idx, xb, w, acc)))
}

// theta index of a slot of term tc (-1: unused)
__device__ __forceinline__ int term_slot_theta(const TermC& tc, int slot) {
  if (slot == SLOT_SCALE) return tc.scale_idx;
  if (slot == SLOT_GATE) return tc.gate != DGP_GATE_NONE ? tc.gate_theta : -1;
  const int f = (slot - SLOT_F0) / SLOT_PER_F, d = (slot - SLOT_F0) % SLOT_PER_F;
  if (f >= tc.nf) return -1;
  if (d == DGP_MAX_FDIMS) return tc.f[f].period_idx;
  return d < tc.f[f].ndims ? tc.f[f].ls_idx[d] : -1;
}

}  // namespace dgp
