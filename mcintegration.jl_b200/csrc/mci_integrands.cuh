/*
 * mci_integrands.cuh — built-in device integrands (enum mci_integrand_id, include/mcib200.h).
 *
 * A Julia closure f(x, c) (src/main.jl:25-33) cannot run on the device, so each integrand of the
 * reference's tests / examples and of the BASELINE configs is a device function selected by id.
 * `X(leaf, pos)` is the value of pool slot `pos` (1-based, like the reference's x[pos]) of leaf
 * variable `leaf` (0-based, in top-level order then leaf order).
 *
 * MAXD > 0: the variable is one Continuous pool and the loops are fully unrolled over MAXD slots
 * (predicated on the run-time dof) so that x[] stays in registers; MAXD == 0: run-time loops.
 */
#pragma once
#include "mci_device.cuh"

namespace mci {

#define MCI_PI 3.14159265358979323846

/* value accessors */
template <int MAXD>
struct PoolX {
  double v[MAXD];
  __device__ __forceinline__ double operator()(int /*leaf*/, int pos) const { return v[pos - 1]; }
};
struct FlatX {
  const double *v;
  const Problem *P;
  __device__ __forceinline__ double operator()(int leaf, int pos) const {
    return v[P->leaf[leaf].base + (pos - 1) * P->leaf[leaf].stride];
  }
};

#define MCI_FOR_D(i, d, MAXD) _Pragma("unroll") for (int i = 0; i < ((MAXD) ? (MAXD) : (d)); ++i) if (i < (d))

/* dof of the pool as the integrand loops see it: the compile-time MAXD in an EXACT instantiation (no predication at
 * all), else the run-time value */
template <int MAXD, bool EXACT>
__device__ __forceinline__ int pool_dof(const Problem &P) { return EXACT ? MAXD : P.dof[0]; }

/* exp(x) for x <= 0 with the polynomial coefficients as constant-bank operands of the DFMAs (the libdevice exp
 * materialises its 13 coefficients through 26 UMOVs per call, which cost issue slots in a kernel that is issue-bound).
 * Cody-Waite reduction x = n ln2 + r, |r| <= ln2/2, degree-13 Taylor/Horner polynomial (truncation < 2e-18), result
 * scaled by 2^n through the exponent field; n < -1020 (subnormal results) and x < -746 (zero) take a scaled path.
 * Max observed error 1 ulp against glibc on [-745, 0] (tests/test_gpu_parity.py::test_device_exp). */
__constant__ double MCI_EXP_C[14] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                     1.0 / 40320.0,      1.0 / 5040.0,      1.0 / 720.0,      1.0 / 120.0,     1.0 / 24.0,
                                     1.0 / 6.0,          0.5,               1.0,              1.0};
__constant__ double MCI_EXP_K[4] = {1.4426950408889634074 /* log2(e) */, 6.93147180369123816490e-01 /* ln2 hi */,
                                    1.90821492927058770002e-10 /* ln2 lo */, 6755399441055744.0 /* 1.5 * 2^52 */};
__device__ __forceinline__ double exp_nonpos(double x) {
  /* n = rint(x * log2e) by the 1.5*2^52 trick: the low word of the biased sum is n as an int32 */
  const double t = fma(x, MCI_EXP_K[0], MCI_EXP_K[3]);
  const int n = __double2loint(t);
  const double nd = t - MCI_EXP_K[3];
  double r = fma(nd, -MCI_EXP_K[1], x);
  r = fma(nd, -MCI_EXP_K[2], r);
  double p = MCI_EXP_C[0];
#pragma unroll
  for (int i = 1; i < 14; ++i) p = fma(p, r, MCI_EXP_C[i]);
  if (n >= -1020) return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
  if (x < -746.0) return 0.0;
  /* subnormal range: scale in two exact steps */
  const double p1 = __hiloint2double(__double2hiint(p) + ((n + 1000) << 20), __double2loint(p));
  return p1 * __hiloint2double((1023 - 1000) << 20, 0);
}

/* s^n for a small positive integer n by repeated squaring */
__device__ __forceinline__ double ipow(double s, int n) {
  double r = 1.0;
  while (n) {
    if (n & 1) r *= s;
    s *= s;
    n >>= 1;
  }
  return r;
}

/* example/bubble.jl:40-50 */
/* The four cases (sign of tau, sign of w) of the reference differ only in the exponent of the numerator and in the
 * overall sign; the denominator is 1 + exp(-|w| beta) in all of them.  Selecting the exponent first leaves two exp and
 * one division per call whatever the mix of cases inside a warp — the same values, bit for bit, as the four-way branch:
 *   tau >= 0, w > 0 :  exp(-w tau)          tau >= 0, w <= 0 :  exp(w (beta - tau))
 *   tau <  0, w > 0 : -exp(-w (tau + beta)) tau <  0, w <= 0 : -exp(-w tau) */
__device__ __forceinline__ double green(double tau, double w, double beta) {
  const bool tp = tau >= 0.0, wp = w > 0.0;
  const double den = 1 + exp(-fabs(w) * beta);
  const double t2 = tp ? beta - tau : tau + beta;
  const double arg = tp == wp ? -w * tau : (tp ? w : -w) * t2;
  const double g = exp(arg) / den;
  return tp ? g : -g;
}

template <int FID>
struct Integrand;

/* doubles per integrand value: 1 for a real integrand, 2 (re, im) for a complex one (Configuration(type=ComplexF64),
 * configuration.jl:108).  A JIT program answers for the user integrand from its compile-time problem tables. */
template <int FID>
struct IntegrandWD {
#ifdef MCI_JIT
  static constexpr int value = FID == MCI_F_USER ? JitProblem::wdim : 1;
#else
  static constexpr int value = 1;
#endif
};
template <> struct IntegrandWD<MCI_F_COMPLEX1> { static constexpr int value = 2; };
template <> struct IntegrandWD<MCI_F_COMPLEX2> { static constexpr int value = 2; };

/* BASELINE C2: (2pi)^(-d/2) exp(-|x|^2/2); fconst[0] = (2pi)^(-d/2) */
template <>
struct Integrand<MCI_F_GAUSS_ND> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    const int d = pool_dof<MAXD, EXACT>(P);
    double s = 0.0;
    MCI_FOR_D(i, d, MAXD) { double x = X(0, i + 1); s += x * x; }
    w[0] = P.fconst[0] * exp_nonpos(-0.5 * s);
  }
};

/* test/montecarlo.jl:129-134 */
template <>
struct Integrand<MCI_F_LOG_OVER_SQRT> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    double x = X(0, 1);
    w[0] = log(x) / sqrt(x);
  }
};

/* Genz corner peak (1 + sum c_i x_i)^-(d+1), pool layout, c = userdata */
template <>
struct Integrand<MCI_F_GENZ_CORNER_POOL> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    const int d = pool_dof<MAXD, EXACT>(P);
    double s = 1.0;
    MCI_FOR_D(i, d, MAXD) s += P.userdata[i] * X(0, i + 1);
    w[0] = 1.0 / ipow(s, d + 1);
  }
};

/* same on a d-leaf CompositeVar, one slot per leaf */
template <>
struct Integrand<MCI_F_GENZ_CORNER_COMP> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    const int d = P.nleaf;
    double s = 1.0;
    for (int i = 0; i < d; ++i) s += P.userdata[i] * X(i, 1);
    w[0] = 1.0 / ipow(s, d + 1);
  }
};

/* example/bubble.jl:50-75; userdata = [kF, beta, me, spin, dim, Qsize, extQ_x...]; fconst[0] = (2pi)^-dim */
template <>
struct Integrand<MCI_F_BUBBLE> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    const double *ud = P.userdata;
    double kF = ud[0], beta = ud[1], me = ud[2], spin = ud[3];
    double R = X(0, 1), th = X(1, 1), ph = X(2, 1), T = X(3, 1);
    int ext = (int)X(4, 1);
    double r = R / (1 - R);
    double sth, cth, sph, cph;
    sincos(th, &sth, &cth);
    sincos(ph, &sph, &cph);
    double k0 = r * sth * cph, k1 = r * sth * sph, k2 = r * cth;
    double factor = P.fconst[0];
    factor *= r * r / ((1 - R) * (1 - R)) * sth;
    double q0 = ud[6 + ext - 1];
    double kq0 = k0 + q0, kq1 = k1, kq2 = k2;
    double tau = T;
    double w1 = ((k0 * k0 + k1 * k1 + k2 * k2) - kF * kF) / (2 * me);
    double g1 = green(tau, w1, beta);
    double w2 = ((kq0 * kq0 + kq1 * kq1 + kq2 * kq2) - kF * kF) / (2 * me);
    double g2 = green(-tau, w2, beta);
    w[0] = g1 * g2 * spin * factor; /* * cos(2 pi 0 tau / beta) == 1 */
  }
};

/* test/bubble_FermiK.jl:54-75: the same diagram on vars (T, K::FermiK{dim}, Ext) — the loop momentum is the FermiK
 * variable itself (leaves 1..dim = its Cartesian components), so there is no spherical Jacobian; userdata as above;
 * fconst[0] = 1/(2pi)^dim (the test fixes dim = 3) */
template <>
struct Integrand<MCI_F_BUBBLE_FERMIK> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    const double *ud = P.userdata;
    const int D = P.nleaf - 2;
    double kF = ud[0], beta = ud[1], me = ud[2], spin = ud[3];
    double tau = X(0, 1), k0 = X(1, 1), k1 = X(2, 1), k2 = D == 3 ? X(3, 1) : 0.0;
    int ext = (int)X(D + 1, 1);
    double kq0 = k0 + ud[6 + ext - 1];
    double w1 = ((k0 * k0 + k1 * k1 + k2 * k2) - kF * kF) / (2 * me);
    double g1 = green(tau, w1, beta);
    double w2 = ((kq0 * kq0 + k1 * k1 + k2 * k2) - kF * kF) / (2 * me);
    double g2 = green(-tau, w2, beta);
    w[0] = g1 * g2 * spin * P.fconst[0]; /* * cos(2 pi 0 tau / beta) == 1 */
  }
};

/* f_k = prod_i (1 + a_k (x_i - 1/2)), a = userdata[0..3] */
template <>
struct Integrand<MCI_F_PRODUCT_4> {
  static constexpr int NI = 4;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    const int d = pool_dof<MAXD, EXACT>(P);
    double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
    MCI_FOR_D(i, d, MAXD) {
      double xc = X(0, i + 1) - 0.5;
      p0 *= (1 + P.userdata[0] * xc);
      p1 *= (1 + P.userdata[1] * xc);
      p2 *= (1 + P.userdata[2] * xc);
      p3 *= (1 + P.userdata[3] * xc);
    }
    w[0] = p0; w[1] = p1; w[2] = p2; w[3] = p3;
  }
};

/* test/montecarlo.jl:4-9 */
template <>
struct Integrand<MCI_F_SPHERE1> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    double a = X(0, 1), b = X(0, 2);
    w[0] = (__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)) < 1.0) ? 1.0 : 0.0;
  }
};

/* test/montecarlo.jl:19-24 */
template <>
struct Integrand<MCI_F_SPHERE2> {
  static constexpr int NI = 2;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    double a = X(0, 1), b = X(0, 2), z = X(0, 3);
    double r2 = __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
    w[0] = (r2 < 1.0) ? 1.0 : 0.0;
    w[1] = (__dadd_rn(r2, __dmul_rn(z, z)) < 1.0) ? 1.0 : 0.0;
  }
};

/* test/montecarlo.jl:104-111 */
template <>
struct Integrand<MCI_F_DISCRETE_X> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) { w[0] = X(0, 1); }
};

/* test/montecarlo.jl:113-120 */
template <>
struct Integrand<MCI_F_DISCRETE2_CONST> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) { w[0] = 1.0; }
};

/* test/montecarlo.jl:136-147 */
template <>
struct Integrand<MCI_F_LATTICE_POOL> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    w[0] = 1.0 / (1.0 - cos(X(0, 1)) * cos(X(0, 2)) * cos(X(0, 3))) / (MCI_PI * MCI_PI * MCI_PI);
  }
};

/* test/montecarlo.jl:149-164 */
template <>
struct Integrand<MCI_F_LATTICE_COMP> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    w[0] = 1.0 / (1.0 - cos(X(0, 1)) * cos(X(1, 1)) * cos(X(2, 1))) / (MCI_PI * MCI_PI * MCI_PI);
  }
};

/* test/montecarlo.jl:200-216 with N=3; fconst[i-1] = volume_inverse(i+1) */
template <>
struct Integrand<MCI_F_HYPERSPHERE3> {
  static constexpr int NI = 3;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    double x1 = X(0, 1);
    double s = __dmul_rn(x1, x1);
#pragma unroll
    for (int i = 1; i <= 3; ++i) {
      double xi = X(0, i + 1);
      s = __dadd_rn(s, __dmul_rn(xi, xi));
      w[i - 1] = s < 1.0 ? P.fconst[i - 1] : 0.0;
    }
  }
};

/* test/thread.jl: (x, x^2, x^3) */
template <>
struct Integrand<MCI_F_POLY3> {
  static constexpr int NI = 3;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    double x = X(0, 1);
    w[0] = x; w[1] = x * x; w[2] = x * x * x;
  }
};

/* vegas/montecarlo.jl:69 */
template <>
struct Integrand<MCI_F_X2_PLUS_Y2> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    double a = X(0, 1), b = X(0, 2);
    w[0] = a * a + b * b;
  }
};

/* test/montecarlo.jl:166-170 (TestComplex1): f = x + i x^2 */
template <>
struct Integrand<MCI_F_COMPLEX1> {
  static constexpr int NI = 1;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    const double x = X(0, 1);
    w[0] = x; w[1] = __dmul_rn(x, x);
  }
};

/* test/montecarlo.jl:172-198 (TestComplex2): (x + 0i, i x^2) */
template <>
struct Integrand<MCI_F_COMPLEX2> {
  static constexpr int NI = 2;
  template <int MAXD, class XA, bool EXACT = false>
  __device__ __forceinline__ static void eval(const XA &X, const Problem &P, double *w) {
    const double x = X(0, 1);
    w[0] = x; w[1] = 0.0; w[2] = 0.0; w[3] = __dmul_rn(x, x);
  }
};

/* run-time dispatch for the generic (non-unrolled) kernels */
#define MCI_FOR_EACH_FID(M)                                                                                         \
  M(MCI_F_GAUSS_ND) M(MCI_F_LOG_OVER_SQRT) M(MCI_F_GENZ_CORNER_POOL) M(MCI_F_GENZ_CORNER_COMP) M(MCI_F_BUBBLE)      \
  M(MCI_F_PRODUCT_4) M(MCI_F_SPHERE1) M(MCI_F_SPHERE2) M(MCI_F_DISCRETE_X) M(MCI_F_DISCRETE2_CONST)                 \
  M(MCI_F_LATTICE_POOL) M(MCI_F_LATTICE_COMP) M(MCI_F_HYPERSPHERE3) M(MCI_F_POLY3) M(MCI_F_X2_PLUS_Y2)                 \
  M(MCI_F_COMPLEX1) M(MCI_F_COMPLEX2)
/* integrands on a FermiK variable: mcmc only (test/bubble_FermiK.jl:139-140) */
#define MCI_FOR_EACH_MCMC_ONLY_FID(M) M(MCI_F_BUBBLE_FERMIK)

}  // namespace mci
