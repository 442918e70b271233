// bfe2_dd.cuh -- double-double arithmetic (two FP64 numbers, ~106-bit significand, unit roundoff ~1.2e-32)
// for the single-large-structure path.
//
// Why: a cantilever of n Hermitian elements has cond(K) ~ 16 n^4 = 1.6e21 at n = 1e5.  Once K is rounded to
// FP64 the information about the low modes and the tip deflection is gone (every FP64 solver, scipy included,
// is ~100 % off).  The inputs (coordinates, E, A, I, rho) ARE exact FP64 numbers, so evaluating the element
// matrices, the block factorisation and the substitutions in double-double keeps eps * cond at ~2e-11.
//
// Error-free transforms (Dekker / Knuth / Shewchuk; QD library formulation).  Every multiplication that feeds
// an error term uses the explicitly rounded intrinsic on the device so nvcc's FMA contraction cannot fuse it
// with a following addition; host builds (tools/beam_host_check) compile with -ffp-contract=off.
#pragma once

#include <cuda_runtime.h>
#include <math.h>

#ifdef __CUDA_ARCH__
#define BFE2_MUL(a, b) __dmul_rn((a), (b))
#define BFE2_ADD(a, b) __dadd_rn((a), (b))
#define BFE2_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define BFE2_MUL(a, b) ((a) * (b))
#define BFE2_ADD(a, b) ((a) + (b))
#define BFE2_FMA(a, b, c) fma((a), (b), (c))
#endif
#define BFE2_HD __host__ __device__ __forceinline__

namespace bfe2 {

struct dd {
    double hi, lo;
};

BFE2_HD void two_sum(double a, double b, double& s, double& e) {
    s = BFE2_ADD(a, b);
    const double bb = BFE2_ADD(s, -a);
    e = BFE2_ADD(BFE2_ADD(a, -BFE2_ADD(s, -bb)), BFE2_ADD(b, -bb));
}
BFE2_HD void quick_two_sum(double a, double b, double& s, double& e) {     // |a| >= |b|
    s = BFE2_ADD(a, b);
    e = BFE2_ADD(b, -BFE2_ADD(s, -a));
}
BFE2_HD void two_prod(double a, double b, double& p, double& e) {
    p = BFE2_MUL(a, b);
    e = BFE2_FMA(a, b, -p);
}

BFE2_HD dd dd_make(double a) { return dd{a, 0.0}; }
BFE2_HD dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
// exact a + b, a - b, a * b of two FP64 numbers
BFE2_HD dd dd_sum2(double a, double b) { dd r; two_sum(a, b, r.hi, r.lo); return r; }
BFE2_HD dd dd_diff2(double a, double b) { dd r; two_sum(a, -b, r.hi, r.lo); return r; }
BFE2_HD dd dd_prod2(double a, double b) { dd r; two_prod(a, b, r.hi, r.lo); return r; }

BFE2_HD dd dd_add(dd a, dd b) {                        // "IEEE" addition of the QD library: error <= 2 u^2 |a + b|
    double s1, s2, t1, t2;
    two_sum(a.hi, b.hi, s1, s2);
    two_sum(a.lo, b.lo, t1, t2);
    s2 = BFE2_ADD(s2, t1);
    quick_two_sum(s1, s2, s1, s2);
    s2 = BFE2_ADD(s2, t2);
    quick_two_sum(s1, s2, s1, s2);
    return dd{s1, s2};
}
BFE2_HD dd dd_sub(dd a, dd b) { return dd_add(a, dd_neg(b)); }
BFE2_HD dd dd_mul(dd a, dd b) {
    double p1, p2;
    two_prod(a.hi, b.hi, p1, p2);
    p2 = BFE2_ADD(p2, BFE2_ADD(BFE2_MUL(a.hi, b.lo), BFE2_MUL(a.lo, b.hi)));
    quick_two_sum(p1, p2, p1, p2);
    return dd{p1, p2};
}
BFE2_HD dd dd_mul_d(dd a, double b) {
    double p1, p2;
    two_prod(a.hi, b, p1, p2);
    p2 = BFE2_ADD(p2, BFE2_MUL(a.lo, b));
    quick_two_sum(p1, p2, p1, p2);
    return dd{p1, p2};
}
BFE2_HD dd dd_div(dd a, dd b) {                        // three quotient digits (QD "accurate" division)
    const double q1 = a.hi / b.hi;
    dd r = dd_sub(a, dd_mul_d(b, q1));
    const double q2 = r.hi / b.hi;
    r = dd_sub(r, dd_mul_d(b, q2));
    const double q3 = r.hi / b.hi;
    double s, e;
    quick_two_sum(q1, q2, s, e);
    return dd_add(dd{s, e}, dd{q3, 0.0});
}
BFE2_HD dd dd_sqrt(dd a) {                             // Karp: sqrt(a) = a x + (a - (a x)^2) x / 2, x = 1/sqrt(a.hi)
    if (!(a.hi > 0.0)) return dd{a.hi == 0.0 ? 0.0 : NAN, 0.0};
    const double x = 1.0 / sqrt(a.hi);
    const double ax = BFE2_MUL(a.hi, x);
    const dd d = dd_sub(a, dd_prod2(ax, ax));
    double s, e;
    two_sum(ax, BFE2_MUL(d.hi, BFE2_MUL(x, 0.5)), s, e);
    return dd{s, e};
}

BFE2_HD dd operator+(dd a, dd b) { return dd_add(a, b); }
BFE2_HD dd operator-(dd a, dd b) { return dd_sub(a, b); }
BFE2_HD dd operator*(dd a, dd b) { return dd_mul(a, b); }
BFE2_HD dd operator/(dd a, dd b) { return dd_div(a, b); }
BFE2_HD dd operator-(dd a) { return dd_neg(a); }
BFE2_HD dd operator*(double a, dd b) { return dd_mul_d(b, a); }
BFE2_HD dd& operator+=(dd& a, dd b) { a = dd_add(a, b); return a; }
BFE2_HD dd& operator-=(dd& a, dd b) { a = dd_sub(a, b); return a; }

// ---- running sum of products  f +- a1 b1 +- a2 b2 ...  with ONE renormalisation at the end ---------------------
//   s carries the leading part through error-free two_sums; every error term and every low-order part of a product
//   is collected in e.  Those parts are 2^-53 of their terms, so adding them in FP64 costs 2^-106 relative to the
//   largest term: the accuracy of the same sum written as a chain of dd_mul / dd_add, at a quarter of the operations
//   and a fifth of the dependent depth (the chain kernels are latency bound: profiles/r2_config4_partitions.txt).
template <class S> struct DotAcc;
template <> struct DotAcc<double> {
    double v;
    BFE2_HD void init(double f) { v = f; }
    BFE2_HD void add(double a, double b) { v = BFE2_FMA(a, b, v); }
    BFE2_HD void sub(double a, double b) { v = BFE2_FMA(-a, b, v); }
    BFE2_HD double get() const { return v; }
};
template <> struct DotAcc<dd> {
    double s, e;
    BFE2_HD void init(dd f) { s = f.hi; e = f.lo; }
    BFE2_HD void add(dd a, dd b) {
        double p, q, t, r;
        two_prod(a.hi, b.hi, p, q);
        q = BFE2_ADD(q, BFE2_FMA(a.hi, b.lo, BFE2_MUL(a.lo, b.hi)));
        two_sum(s, p, t, r);
        s = t;
        e = BFE2_ADD(e, BFE2_ADD(r, q));
    }
    BFE2_HD void sub(dd a, dd b) { add(dd{-a.hi, -a.lo}, b); }
    BFE2_HD dd get() const { dd r; two_sum(s, e, r.hi, r.lo); return r; }
};

// ---- scalar traits: the chain kernels are templates over S = double (plain FP64) or S = dd ------------------
template <class S> struct sc;
template <> struct sc<double> {
    static BFE2_HD double make(double a) { return a; }
    static BFE2_HD double diff(double a, double b) { return a - b; }
    static BFE2_HD double prod(double a, double b) { return a * b; }
    static BFE2_HD double root(double a) { return sqrt(a); }
    static BFE2_HD double hi(double a) { return a; }
    static BFE2_HD double lo(double) { return 0.0; }
    static BFE2_HD double join(double h, double) { return h; }
};
template <> struct sc<dd> {
    static BFE2_HD dd make(double a) { return dd{a, 0.0}; }
    static BFE2_HD dd diff(double a, double b) { return dd_diff2(a, b); }
    static BFE2_HD dd prod(double a, double b) { return dd_prod2(a, b); }
    static BFE2_HD dd root(dd a) { return dd_sqrt(a); }
    static BFE2_HD double hi(dd a) { return a.hi; }
    static BFE2_HD double lo(dd a) { return a.lo; }
    static BFE2_HD dd join(double h, double l) { return dd_sum2(h, l); }
};

}  // namespace bfe2
