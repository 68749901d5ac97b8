// Built-in integrands, evaluated inline in the fused event kernel.
// symgauss and product follow the reference's operation order bit for bit; the translation
// units are compiled with -fmad=false so each multiply/add is rounded separately, as
// TensorFlow's unfused elementwise ops are.  The two LO matrix elements (parity bar 1e-12 on
// w*f, not bit-exactness) evaluate the reference's spinor chain but regroup what is exactly or
// harmlessly equivalent -- exact-zero parts of the spinor components, square roots instead of
// acos/sincos round trips, one collected quotient -- and keep the reference's own operations
// wherever its rounding is amplified (see the comments at each place).
// Reference citations are file:line relative to /root/reference.
#pragma once
#include "vf_common.cuh"

// block sizes of the matrix-element kernels (tunable at build time; measured defaults)
#ifndef VF_DY_THREADS
#define VF_DY_THREADS 1024
#endif
#ifndef VF_ST_THREADS
#define VF_ST_THREADS 896
#endif

namespace vf {

// ---------------------------------------------------------------------------
// symgauss, examples/simgauss_tf.py:22-32 (same body tests/test_algs.py:27-36)
// consts.p[0] = pref = (1/a/sqrt(pi))^d, consts.p[1] = C = sum_{i<=100d} i.
// The literal "+C ... -C" is kept: it quantises coef to ulp(C).
// ---------------------------------------------------------------------------
// y / 0.1 with IEEE round-to-nearest in TWO fp64 operations instead of the ~14 of the generic
// division.  a = fl(0.1) = (1 + 2^-54)/10 exactly, so y/a = 10y/(1 + 2^-54) =
// 10y - 10y*2^-54 + 10y*2^-108 - ...; t = rn(y * (-10*2^-54)) carries the second term with a
// rounding error <= 2^-51 ulp-units and fma(y, 10, t) adds it to the EXACT 10y with one rounding.
// The value before that rounding is within 2^-50.4 (in units of ulp(y)) of the true quotient,
// and no quotient n*10*2^54/(2^54+1) (n a 53-bit integer) lies that close to a rounding boundary
// 4M: 10n*2^54 - 4M(2^54+1) = r with 0 < |r| <= 12 forces 10n = j*2^55 + 2j - r (j = 1, 2), which
// is never divisible by 10 for r in {+-4, +-8, +-12}.  Hence the result equals `y / 0.1` for every
// normal y.  Checked against `/` on 6e8 host samples and by the parity tests.
__device__ __forceinline__ double div_by_tenth(double y) {
    const double t = __dmul_rn(y, -5.5511151231257827e-16);  // -10 * 2^-54, exact constant
    return __fma_rn(y, 10.0, t);
}

// exp(x) for x <= 0 (symgauss always calls exp(-coef) with coef >= 0).  Cody-Waite reduction
// x = k*ln2 + r, |r| <= ln2/2, degree-11 near-minimax polynomial (Chebyshev interpolant of
// (e^r-1-r)/r^2, approximation error 0.14 ulp; derivation in DESIGN.md), scaling by an exponent
// add.  Coefficients sit in constant memory so each Horner step is one DFMA with a constant-bank
// operand.  Total error ~1 ulp, i.e. the same class as libm/libdevice exp.
static __constant__ double kExpCoef[12] = {
    1.0, 1.0, 0.5000000000000001, 0.16666666666666669, 0.041666666666624164,
    0.008333333333330065, 0.0013888888917196719, 0.00019841269863040545,
    2.4801521322368692e-05, 2.7557268480310024e-06, 2.7620075879983367e-07,
    2.5100375832561234e-08};

__device__ __forceinline__ double exp_nonpositive(double x) {
    const double kMagic = 6755399441055744.0;  // 1.5 * 2^52: round-to-nearest-integer add
    const double t = fma(x, 1.4426950408889634, kMagic);
    const int k = __double2loint(t);
    const double kf = t - kMagic;
    double r = fma(kf, -6.93147180559945286e-01, x);
    r = fma(kf, -2.31904681384629956e-17, r);
    double p = kExpCoef[11];
#pragma unroll
    for (int i = 10; i >= 0; --i) p = fma(p, r, kExpCoef[i]);
    // x < -700 tested on the high word (x <= 0: the unsigned order of the words is the order of
    // |x|; 0xC085E000 = hi(-700.0)) -- an integer compare instead of a DSETP on the fp64 pipe.
    // For -700 - 2^-43 < x < -700 this takes the direct path, which is exact down to 2^-1022.
    if ((uint32_t)__double2hiint(x) > 0xC085E000u) {  // two-step scaling through the subnormals
        if (x < -800.0) return 0.0;
        const double q = __hiloint2double(__double2hiint(p) + ((k + 256) << 20), __double2loint(p));
        return q * 8.636168555094445e-78;  // 2^-256
    }
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

struct SymGauss {
    static constexpr int kFixedDim = 0;
    static constexpr bool kHeavy = false;
    template <int NDIM>
    static __device__ __forceinline__ double eval(const double (&x)[NDIM],
                                                  const IntegrandConsts& c) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < NDIM; ++j) {
            const double t = div_by_tenth(__dsub_rn(x[j], 0.5));  // :30 (x - 1/2)/a, a = 0.1
            const double q = __dmul_rn(t, t);
            s = (j == 0) ? q : __dadd_rn(s, q);  // reduce_sum axis=1, left to right
        }
        double coef = __dadd_rn(c.p[1], s);  // :29-30
        coef = __dsub_rn(coef, c.p[1]);      // :31
        return __dmul_rn(c.p[0], exp_nonpositive(-coef));  // :32
    }
};

// product, README.md:63-68 / tests/test_misc.py:24-26
struct Product {
    static constexpr int kFixedDim = 0;
    static constexpr bool kHeavy = false;
    template <int NDIM>
    static __device__ __forceinline__ double eval(const double (&x)[NDIM], const IntegrandConsts&) {
        double p = x[0];
#pragma unroll
        for (int j = 1; j < NDIM; ++j) p = __dmul_rn(p, x[j]);
        return p;
    }
};

// ---------------------------------------------------------------------------
// Spinor helpers shared by the two LO matrix elements
// (examples/drellyan_lo_tf.py:88-204, examples/singletop_lo_tf.py:105-218).
//
// The reference works with complex128 four-spinors.  Two of the four components are exact
// complex zeros for every helicity; of the other two, many are purely real (prefactor
// sqrt(2 E), E > 0, times a cosine) or purely imaginary (beam momenta enter as -p, E' < 0, so
// the prefactor is i sqrt(2|E'|)), i.e. complex numbers whose other part is an exact zero.
// Products and sums with exact zeros change nothing in the remaining terms (x*0 = 0, y - 0 = y
// for finite operands), so components carry their kind in the TYPE -- Zr, Rl, Ig or cplx -- and
// mul()/add() form only the terms that are not exact zeros: the same values, bit for bit, as
// complex arithmetic on the padded numbers (signed zeros aside), at a fraction of the
// multiplications.
// ---------------------------------------------------------------------------
struct cplx {
    double re, im;
};
struct Zr {};      // exact 0
struct Rl {        // v + 0i
    double v;
};
struct Ig {        // 0 + i v
    double v;
};
// Spinor products and their two-term sums use explicit fused multiply-adds (VF_ME_FMA, default
// on): the reference rounds every product, the fused forms skip that rounding -- a deviation of
// an ulp, far inside the 1e-12 bar of the matrix elements, for a third fewer instructions.  The
// translation unit is still compiled with -fmad=false: nothing else is contracted.
#ifndef VF_ME_FMA
#define VF_ME_FMA 1
#endif
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
#if VF_ME_FMA
    return {fma(a.re, b.re, -(a.im * b.im)), fma(a.re, b.im, a.im * b.re)};
#else
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
#endif
}
// products
__device__ __forceinline__ Zr mul(Zr, Zr) { return {}; }
template <class X> __device__ __forceinline__ Zr mul(Zr, X) { return {}; }
template <class X> __device__ __forceinline__ Zr mul(X, Zr) { return {}; }
__device__ __forceinline__ Rl mul(Rl a, Rl b) { return {a.v * b.v}; }
__device__ __forceinline__ Ig mul(Rl a, Ig b) { return {a.v * b.v}; }
__device__ __forceinline__ Ig mul(Ig a, Rl b) { return {a.v * b.v}; }
__device__ __forceinline__ Rl mul(Ig a, Ig b) { return {-(a.v * b.v)}; }
__device__ __forceinline__ cplx mul(Rl a, cplx b) { return {a.v * b.re, a.v * b.im}; }
__device__ __forceinline__ cplx mul(cplx a, Rl b) { return {a.re * b.v, a.im * b.v}; }
__device__ __forceinline__ cplx mul(Ig a, cplx b) { return {-(a.v * b.im), a.v * b.re}; }
__device__ __forceinline__ cplx mul(cplx a, Ig b) { return {-(a.im * b.v), a.re * b.v}; }
__device__ __forceinline__ cplx mul(cplx a, cplx b) { return cmul(a, b); }
// sums
__device__ __forceinline__ Zr add(Zr, Zr) { return {}; }
template <class X> __device__ __forceinline__ X add(Zr, X x) { return x; }
template <class X> __device__ __forceinline__ X add(X x, Zr) { return x; }
__device__ __forceinline__ Rl add(Rl a, Rl b) { return {a.v + b.v}; }
__device__ __forceinline__ Ig add(Ig a, Ig b) { return {a.v + b.v}; }
__device__ __forceinline__ cplx add(Rl a, Ig b) { return {a.v, b.v}; }
__device__ __forceinline__ cplx add(Ig a, Rl b) { return {b.v, a.v}; }
__device__ __forceinline__ cplx add(Rl a, cplx b) { return {a.v + b.re, b.im}; }
__device__ __forceinline__ cplx add(cplx a, Rl b) { return {a.re + b.v, a.im}; }
__device__ __forceinline__ cplx add(Ig a, cplx b) { return {b.re, a.v + b.im}; }
__device__ __forceinline__ cplx add(cplx a, Ig b) { return {a.re, a.im + b.v}; }
__device__ __forceinline__ cplx add(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
// x*y + acc
template <class X, class Y, class A>
__device__ __forceinline__ auto madd(X x, Y y, A acc) {
    return add(mul(x, y), acc);
}
#if VF_ME_FMA
__device__ __forceinline__ cplx madd(Rl x, cplx y, cplx acc) {
    return {fma(x.v, y.re, acc.re), fma(x.v, y.im, acc.im)};
}
__device__ __forceinline__ cplx madd(cplx x, Rl y, cplx acc) {
    return {fma(x.re, y.v, acc.re), fma(x.im, y.v, acc.im)};
}
__device__ __forceinline__ cplx madd(Rl x, Ig y, cplx acc) { return {acc.re, fma(x.v, y.v, acc.im)}; }
__device__ __forceinline__ cplx madd(Ig x, Rl y, cplx acc) { return {acc.re, fma(x.v, y.v, acc.im)}; }
__device__ __forceinline__ cplx madd(Ig x, cplx y, Ig acc) {
    return {-(x.v * y.im), fma(x.v, y.re, acc.v)};
}
__device__ __forceinline__ cplx madd(cplx x, Ig y, Ig acc) {
    return {-(x.im * y.v), fma(x.re, y.v, acc.v)};
}
__device__ __forceinline__ cplx madd(cplx x, cplx y, cplx acc) {
    return {fma(x.re, y.re, fma(-x.im, y.im, acc.re)), fma(x.re, y.im, fma(x.im, y.re, acc.im))};
}
#endif
// real part of a product
template <class X, class Y>
__device__ __forceinline__ double re_mul(X x, Y y) {
    return re_of(mul(x, y));
}
// real part, squared modulus
__device__ __forceinline__ double re_of(Zr) { return 0.0; }
__device__ __forceinline__ double re_of(Rl a) { return a.v; }
__device__ __forceinline__ double re_of(Ig) { return 0.0; }
__device__ __forceinline__ double re_of(cplx a) { return a.re; }
__device__ __forceinline__ double norm2(cplx a) {
#if VF_ME_FMA
    return fma(a.re, a.re, a.im * a.im);
#else
    return a.re * a.re + a.im * a.im;
#endif
}

struct Mom {
    double e, x, y, z;
};

// the two non-zero components of a spinor, each with the kind of its value
template <class A, class B>
struct Sp {
    A a;
    B b;
};
template <class A1, class B1, class A2, class B2>
__device__ __forceinline__ auto sdot(const Sp<A1, B1>& bra, const Sp<A2, B2>& ket) {
    return madd(bra.b, ket.b, mul(bra.a, ket.a));
}

__device__ __forceinline__ double clip1(double v) {
    // tf.where(v < -1, -1, v); tf.where(v > 1, 1, .)  (NaN passes through)
    double r = v;
    if (v < -1.0) r = -1.0;
    if (v > 1.0) r = 1.0;
    return r;
}

// What the spinors of a final-state momentum (E > 0: the prefactor pref = sqrt(2)*sqrt(E) is real)
// are built from.
struct HalfSpin {
    double pc, ps;  // pref*cos(theta/2), pref*sin(theta/2)
    double cp, sp;  // cos(phi), sin(phi) of ubar0 (and of Drell-Yan's u0)
};

// The reference takes theta = acos(pz/E), phi = +-acos(px/E/sin(theta)) and then needs only
// cos(theta/2), sin(theta/2), cos(phi), sin(phi) (drellyan :93-162, singletop :105-178).  With
// c = clip(pz/E) and cx = clip(px/E/sin(theta)) these are, for the SAME c and cx,
//     cos(theta/2) = sqrt((1+c)/2)   sin(theta/2) = sqrt((1-c)/2)   sin(theta) = 2 sin cos (theta/2)
//     cos(phi) = cx                  sin(phi) = +-sqrt((1-cx)(1+cx))
// i.e. square roots instead of two acos, one sin and two sincos per momentum -- accurate
// evaluations of the same functions of the same arguments (1+c, 1-c, 1-cx are exact near the ends
// where the acos route is accurate as well), inside the 1e-12 parity bar
// (tests/test_device_source_on_host.py, tests/test_parity_gpu.py).  With the prefactor folded in,
//     pc = pref cos(theta/2) = sqrt(E (1+c))      ps = pref sin(theta/2) = sqrt(E (1-c))
//     E sin(theta) = pc ps, hence  cx = clip(px / (pc ps)):
// three square roots and two divisions per momentum.  c itself stays the reference's own
// quotient: where theta -> pi amplifies its rounding, 1 + c reproduces it.
// The p.x == 0 branch keeps the reference's numerically evaluated constants:
// cos(pi/2) = 6.123e-17, sin(pi) = 1.2246e-16.
constexpr double kCosHalfPi = 6.123233995736766e-17;   // np.cos(np.pi / 2)
constexpr double kSinPi = 1.2246467991473532e-16;      // np.sin(np.pi)
constexpr double kSqrt2 = 1.4142135623730951;          // np.sqrt(2)

// (cos, sin)(theta/2) for the beam-axis case px == 0: theta in {0, pi, rz (0 or NaN)}
__device__ __forceinline__ void half_angle_axis(double rz, double& ch, double& sh) {
    ch = 1.0;
    sh = 0.0;
    if (rz < 0.0) {
        ch = kCosHalfPi;
        sh = 1.0;
    }
    if (rz != rz) ch = sh = rz;  // NaN passes through like sincos(NaN)
}

// cos(phi), sin(phi) of phi = +-acos(cx) (sign of py/E; E > 0), the reference's constants at the ends
__device__ __forceinline__ void phi_from_cx(double cx, double py, double& cp, double& sp) {
    cp = cx;
    sp = sqrt((1.0 - cx) * (1.0 + cx));
    if (cx == -1.0) sp = kSinPi;  // phi == pi: the reference's sin(np.pi)
    if (py < 0.0) sp = -sp;       // sign of py/E, E > 0
    if (cx == 1.0) sp = 0.0;      // phi == 0 (also -0): cos 0 = 1, sin 0 = 0
}

// theta/phi of drellyan u0 (:93-115) and ubar0 (:140-162) through the half-angle forms.
__device__ __forceinline__ HalfSpin angles_half(const Mom& p) {
    const double rz = p.z / p.e;
    HalfSpin a;
    if (p.x == 0.0) {
        double ch, sh;
        half_angle_axis(rz, ch, sh);
        const double pref = kSqrt2 * sqrt(p.e);
        a.pc = pref * ch;
        a.ps = pref * sh;
        a.cp = 1.0;  // phi = 0
        a.sp = 0.0;
    } else {
        const double c = clip1(rz);
        a.pc = sqrt(p.e * (1.0 + c));
        a.ps = sqrt(p.e * (1.0 - c));
        phi_from_cx(clip1(p.x / (a.pc * a.ps)), p.y, a.cp, a.sp);
    }
    return a;
}

// Single-top's ubar0 (singletop :156-178) takes theta = acos(c), c = clip(pz/E), and then needs
// cos(theta/2), sin(theta/2) and sin(theta) (for cos(phi) = px/E/sin(theta)).  Away from theta =
// pi the half-angle forms above are accurate evaluations of the same functions of the same c.
// Near threshold the projected top momentum is anti-parallel to the beam, theta -> pi, and the
// reference's cos(theta/2) and sin(theta) carry the ROUNDING of theta = rn(acos(c)) as a relative
// error 2e-16/(pi - theta) that reaches 1e-8: amplitudes proportional to them agree with the
// reference to 1e-12 only if theta is rounded the same way.  For 1 + c < 1e-5 the rounded theta
// is therefore reproduced:  acos(c) = pi - 2 asin(s), s = sqrt((1+c)/2) <= 2.3e-3 (1 + c exact;
// asin by its series, next term 15 s^7/336 < 2e-20), theta_r = rn(pi_hi + pi_lo - delta) by a
// compensated subtraction, and with e = pi - theta_r = (pi_hi - theta_r) + pi_lo (the difference
// is exact) cos(theta_r/2) = sin(e/2), sin(theta_r/2) = cos(e/2), sin(theta_r) = sin(e), again by
// their series (e <= 4.5e-3).  At c = -1 this gives e = pi_lo: cos(theta/2) = 6.12e-17 and
// sin(theta) = 1.22e-16, the reference's np.cos(np.pi/2) and np.sin(np.pi).
// phi = +-acos(cx), cx = clip(px/E/sin(theta)): cos(phi) = cx and sin(phi) = +-sqrt((1-cx)(1+cx))
// are accurate for both ends (py == 0 in single-top, so |cx| is 1 up to the rounding noise
// above; sin(phi) enters the squared amplitudes in second order).  Deep in the threshold region
// (x0 < 1e-4) sin(theta) itself is rounding noise, cx comes out anywhere in [-1, 1], and for
// |cx| < 1e-3 the reference's cos(phi) carries the rounding of phi = rn(acos(cx)) ~ pi/2 as a
// relative error 1e-16/|cx|: reproduced the same way, acos(cx) = pi/2 - asin(cx).
// Every operation here is an IEEE add, multiply, divide or square root: host and device agree
// bit for bit.  E > 0 (both final-state momenta of single-top, singletop :71-92, :236-239).
constexpr double kNearPi = 1e-5;
constexpr double kNearHalfPi = 1e-3;
__device__ __forceinline__ HalfSpin angles_acos(const Mom& p) {
    const double rz = p.z / p.e;
    HalfSpin a;
    a.cp = 1.0;  // phi == 0 (beam-axis momenta, and phi2 == 0): cos 0 = 1, sin 0 = 0
    a.sp = 0.0;
    if (p.x == 0.0) {
        double ch, sh;
        half_angle_axis(rz, ch, sh);  // theta1 in {0, pi}: sincos(theta/2) are constants
        const double pref = kSqrt2 * sqrt(p.e);
        a.pc = pref * ch;
        a.ps = pref * sh;
    } else {
        const double c = clip1(rz);
        const double opc = 1.0 + c;
        double e_sin_theta;  // E sin(theta)
        if (opc < kNearPi) {  // theta within 4.5e-3 of pi: reproduce rn(acos(c))
            const double pi_hi = 3.141592653589793, pi_lo = 1.2246467991473532e-16;
            const double sn = sqrt(0.5 * opc), s2 = sn * sn;
            const double delta = 2.0 * (sn + sn * s2 * (1.0 / 6.0 + s2 * (3.0 / 40.0)));
            const double t = pi_hi - delta;
            const double terr = (pi_hi - t) - delta;      // exact residual of the subtraction
            const double theta = t + (terr + pi_lo);      // rn(pi - delta)
            const double e = (pi_hi - theta) + pi_lo;     // pi - theta_r
            const double h = 0.5 * e, h2 = h * h, e2 = e * e;
            const double ch = h - h * h2 * (1.0 / 6.0 - h2 * (1.0 / 120.0));
            const double sh = 1.0 - h2 * (0.5 - h2 * (1.0 / 24.0 - h2 * (1.0 / 720.0)));
            const double pref = kSqrt2 * sqrt(p.e);
            a.pc = pref * ch;
            a.ps = pref * sh;
            e_sin_theta = p.e * (e - e * e2 * (1.0 / 6.0 - e2 * (1.0 / 120.0)));
        } else {
            a.pc = sqrt(p.e * opc);
            a.ps = sqrt(p.e * (1.0 - c));
            e_sin_theta = a.pc * a.ps;
        }
        const double cx = clip1(p.x / e_sin_theta);  // px/E/sin(theta)
        phi_from_cx(cx, p.y, a.cp, a.sp);
        if (fabs(cx) < kNearHalfPi) {  // phi within 1e-3 of pi/2: reproduce rn(acos(cx))
            const double po2_hi = 1.5707963267948966, po2_lo = 6.123233995736766e-17;
            const double x2 = cx * cx;
            const double as = cx + cx * x2 * (1.0 / 6.0 + x2 * (3.0 / 40.0));  // asin(cx)
            const double t = po2_hi - as;
            const double terr = (po2_hi - t) - as;
            const double phi = t + (terr + po2_lo);     // rn(pi/2 - asin(cx))
            const double g = (po2_hi - phi) + po2_lo;   // pi/2 - phi_r
            const double g2 = g * g;
            a.cp = g - g * g2 * (1.0 / 6.0 - g2 * (1.0 / 120.0));
            a.sp = 1.0 - g2 * (0.5 - g2 * (1.0 / 24.0));
            if (p.y < 0.0) a.sp = -a.sp;
        }
    }
    return a;
}

// u0(p, +1) = (pref cos, pref sin e^{+i phi}, 0, 0)      u0(p, -1) = (0, 0, pref sin e^{-i phi}, -pref cos)
// ubar0(p, +1) = (0, 0, pref cos, pref sin e^{-i phi})   ubar0(p, -1) = (pref sin e^{+i phi}, -pref |cos|, 0, 0)
// (drellyan :88-183, singletop :105-197) for a real prefactor
__device__ __forceinline__ Sp<Rl, cplx> u0_plus(const HalfSpin& g, double cp, double sp) {
    return {Rl{g.pc}, cplx{g.ps * cp, g.ps * sp}};
}
__device__ __forceinline__ Sp<cplx, Rl> u0_minus(const HalfSpin& g, double cp, double sp) {
    return {cplx{g.ps * cp, -(g.ps * sp)}, Rl{-g.pc}};
}
__device__ __forceinline__ Sp<Rl, cplx> ubar0_plus(const HalfSpin& g) {
    return {Rl{g.pc}, cplx{g.ps * g.cp, -(g.ps * g.sp)}};
}
__device__ __forceinline__ Sp<cplx, Rl> ubar0_minus(const HalfSpin& g) {
    return {cplx{g.ps * g.cp, g.ps * g.sp}, Rl{-fabs(g.pc)}};
}

// Beam-axis momenta.  Both examples build the spinors of -p1 = (-E, 0, 0, +E) and
// -p0 = (-E, 0, 0, -E) with E = ecmo2 > 0: the prefactor sqrt(2)*sqrt(complex(-E, 0)) = i q,
// q = sqrt(2) sqrt(E), is purely imaginary; pz/E' is exactly -1 / +1, so theta is pi / 0, phi is
// 0, and the generic construction multiplies by (cos phi, sin phi) = (1, 0), by sin(theta/2) = 1
// or 0 and by cos(theta/2) = 6.12e-17 or 1 at run time.  Written out, the theta = 0 spinors have
// ONE non-zero component (the other is pref * 0, an exact zero).  Same values, bit for bit, as
// the generic route (checked by the parity tests and the full-size checksums).
struct BeamSpin0 {   // theta = 0: u0(+1), u0(-1), ubar0(+1), ubar0(-1) of (E', 0, 0, E'), E' < 0
    Sp<Ig, Zr> up;
    Sp<Zr, Ig> um;
    Sp<Ig, Zr> bp;
    Sp<Zr, Ig> bm;
};
struct BeamSpinPi {  // theta = pi: of (E', 0, 0, -E')
    Sp<Ig, Ig> up, um, bp, bm;
};
__device__ __forceinline__ BeamSpin0 beam_spinors_theta0(double q) {
    return {{Ig{q}, Zr{}}, {Zr{}, Ig{-q}}, {Ig{q}, Zr{}}, {Zr{}, Ig{-q}}};
}
__device__ __forceinline__ BeamSpinPi beam_spinors_thetapi(double q) {
    const double small = q * kCosHalfPi, nsmall = (-q) * kCosHalfPi;
    return {{Ig{small}, Ig{q}}, {Ig{q}, Ig{nsmall}}, {Ig{small}, Ig{q}}, {Ig{q}, Ig{nsmall}}};
}

// ---------------------------------------------------------------------------
// Drell-Yan LO, examples/drellyan_lo_tf.py:27-249 (n_dim = 4)
// ---------------------------------------------------------------------------
struct DrellYanLO {
    static constexpr int kFixedDim = 4;
    static constexpr bool kHeavy = true;
    // measured (profiles/r2_me_threads.txt): 1024 threads x 64 registers beat 768 x 68 by 3 %
    static constexpr int kBlockThreads = VF_DY_THREADS;
    template <int NDIM>
    static __device__ double eval(const double (&xa)[NDIM], const IntegrandConsts&) {
        static_assert(NDIM == 4, "drellyan_lo is 4-dimensional");
        const double s = 14000.0 * 14000.0;      // :18-22
        const double conv = 0.3893793e9;         // :24
        // get_x1x2 :27-41
        const double kappa = xa[0], y = xa[1];
        const double logkappa = log(kappa);
        const double sqrtkappa = sqrt(kappa);
        const double Ycm = exp(logkappa * (y - 0.5));
        const double shat = s;
        const double x1 = sqrtkappa * Ycm;
        const double x2 = sqrtkappa / Ycm;
        const double jac = fabs(logkappa);
        // make_event :44-75
        // mV = sqrt(shat x1 x2) (:46).  Everything below is homogeneous in mV, which spans hundreds
        // of binades (kappa runs down to wherever the grid has zoomed in on the 1/kappa peak), so it
        // is evaluated on mV * 2^-2k in [1, 4) -- an EXACT rescaling: products, quotients, sums of
        // equally scaled terms and square roots of evenly scaled numbers commute with it bit for
        // bit -- and the result, which scales like 1/mV^2, gets its 2^-4k back at the end.  This
        // keeps the collected numerator and denominator below in range wherever the reference's
        // staged quotients are (checked down to kappa = 1e-290).
        const double mV_true = sqrt(shat * x1 * x2);
        const int mv_e = (((__double2hiint(mV_true) >> 20) & 0x7ff) - 1023) & ~1;  // even, <= exponent
        const double unscale = __hiloint2double((1023 - mv_e) << 20, 0);           // 2^-mv_e
        const double mV = mV_true * unscale;
        const double mV2 = mV * mV;
        const double ecmo2 = mV / 2;
        const Mom p0{ecmo2, 0.0, 0.0, ecmo2};
        const Mom p1{ecmo2, 0.0, 0.0, -ecmo2};
        // pV = p0 + p1 = (2 ecmo2, 0, 0, 0) with EXACT zeros, so YV = log(|1|)/2 = 0, pVt2 = 0 and
        // pV.x cos + pV.y sin = 0 exactly (:53-59, dropped like the exact-zero spinor components).
        const Mom pV{p0.e + p1.e, 0.0, 0.0, 0.0};
        const double phi = (2.0 * M_PI) * xa[3];  // :60, 2*np.pi*x3
        double sphi, cphi;
        sincos(phi, &sphi, &cphi);
        // root = sqrt(mV2) with mV2 = rn(mV*mV): in binary IEEE arithmetic sqrt(rn(x*x)) == |x|
        // exactly, so the reference's root IS mV (:61).  ptmax, pta and Delta keep the reference's
        // operations bit for bit: Delta - 1 below is ill-conditioned for x3 -> 1.
        const double root = mV;
        const double ptmax = 0.5 * mV2 / root;
        const double pta = ptmax * xa[2];
        const double ptx = pta * cphi, pty = pta * sphi;
        const double Delta = mV2 / 2.0 / pta / root;
        // yy = YV - acosh(Delta) = -acosh(Delta) (:64): cosh(yy) = Delta and
        // sinh(yy) = -sqrt((Delta-1)(Delta+1)), with Delta - 1 exact (Delta = 1/x2 > 1)
        const double shy = sqrt((Delta - 1.0) * (Delta + 1.0));
        const Mom p2{pta * Delta, ptx, pty, -(pta * shy)};
        const Mom p3{pV.e - p2.e, pV.x - p2.x, pV.y - p2.y, pV.z - p2.z};
        // The scalar factors of the result -- kallenF = 2 ptmax/root/shy (:65), psw =
        // kallenF/(8 pi) * jac (:71-72), flux = 1/(2 mV2) (:73), the 1/lsprod^2 of the squared
        // amplitudes and the 1/x1/x2 of the luminosity -- are collected into one numerator and one
        // denominator and divided ONCE at the end (the reference's seven quotients, regrouped: a
        // few ulp, the parity bar of the matrix elements is 1e-12).
        const double psw_num = ((1.0 / (8.0 * M_PI)) * 2.0) * ptmax * jac;  // psw * flux = psw_num/psw_den
        const double psw_den = root * shy * (2.0 * mV2);
        // qqxllx(-p1, -p0, p2, p3) :207-224
        // q0 = -p1 = (-ecmo2, 0, 0, +ecmo2): theta = pi; q1 = -p0 = (-ecmo2, 0, 0, -ecmo2): theta = 0
        const double bq = kSqrt2 * sqrt(ecmo2);  // prefactor of the beam spinors: i*bq
        const BeamSpinPi s0 = beam_spinors_thetapi(bq);
        const BeamSpin0 s1 = beam_spinors_theta0(bq);
        const HalfSpin a2 = angles_half(p2), a3 = angles_half(p3);
        // za(a,b) = ubar0(a,-1).u0(b,+1); zb(a,b) = ubar0(a,+1).u0(b,-1)
        const auto za01 = sdot(s0.bm, s1.up);
        const auto zb10 = sdot(s1.bp, s0.um);
        const double lsprod = re_mul(za01, zb10);  // sprod(p0,p1) :200-204
        const cplx za02 = sdot(s0.bm, u0_plus(a2, a2.cp, a2.sp));
        const cplx za03 = sdot(s0.bm, u0_plus(a3, a3.cp, a3.sp));
        const cplx zb31 = sdot(ubar0_plus(a3), s1.um);
        const cplx zb21 = sdot(ubar0_plus(a2), s1.um);
        // a = 2|za02 zb31|/lsprod, b = 2|za03 zb21|/lsprod, wgts = 6(a^2 + b^2)/36 (:210-213): the
        // squared moduli are formed without the square roots in between
        const cplx ma = cmul(za02, zb31), mb = cmul(za03, zb21);
        const double na = norm2(ma), nb = norm2(mb);
        const double wgts_num = (na + nb) * (24.0 / 36.0);  // wgts = wgts_num / lsprod^2
        // build_luminosity :232-239 with the toy pdf x1*x2 :226-229: lumis = 4 pdf / x1 / x2
        const double pdf = x1 * x2;
        const double lumis_num = pdf + pdf + pdf + pdf;
        // 2 * lumis * wgts * psw * flux * conv (:246-247)
        const double num = 2 * lumis_num * wgts_num * psw_num * conv;
        const double den = (x1 * x2) * (lsprod * lsprod) * psw_den;
        return num / den * unscale * unscale;
    }
};

// ---------------------------------------------------------------------------
// single-top LO (t-channel), examples/singletop_lo_tf.py:45-270 (n_dim = 3)
// ---------------------------------------------------------------------------
struct SingleTopLO {
    static constexpr int kFixedDim = 3;
    static constexpr bool kHeavy = true;
    // 896 threads x 70 registers: +6 % over 640 x 86, +2.4 % over 768 x 80 and over 1024 x 64
    // (which spills 36 B) since the acos-free angles shortened the live ranges
    // (profiles/r2_me_threads.txt)
    static constexpr int kBlockThreads = VF_ST_THREADS;

    struct AllSpin {  // u0(+1), u0(-1), ubar0(+1), ubar0(-1) of a final-state momentum
        Sp<Rl, cplx> up;
        Sp<cplx, Rl> um;
        Sp<Rl, cplx> bp;
        Sp<cplx, Rl> bm;
    };
    // u0 and ubar0 of one momentum share theta (singletop :110-129 and :156-178 both start from
    // acos(clip(pz/E)) and need sincos(theta/2)): evaluated once, the same operations
    static __device__ __forceinline__ AllSpin spinors(const Mom& p) {
        const HalfSpin g = angles_acos(p);
        double cpu = 1.0, spu = 0.0;  // phi of u0 is 0 or pi from the sign of px/E (:124-126)
        if (p.x < 0.0) {              // E > 0
            cpu = -1.0;               // the reference's np.cos(np.pi), np.sin(np.pi)
            spu = kSinPi;
        }
        return {u0_plus(g, cpu, spu), u0_minus(g, cpu, spu), ubar0_plus(g), ubar0_minus(g)};
    }
    // sprod(p1,p2) = Re(za(p1,p2)*zb(p2,p1)) :213-218 (any mix of generic and beam spinor sets)
    template <class S1, class S2>
    static __device__ __forceinline__ double sprod(const S1& s1, const S2& s2) {
        return re_mul(sdot(s1.bm, s2.up), sdot(s2.bp, s1.um));
    }
    // qqxtbx :221-230 = |(a + mt2 b/c) d| * 9/wprop * gw4/36, returned as the numerator
    // |(a c + mt2 b) d| and the propagator wprop; c = sprod(p2, p3) is the same in both channels
    // and the quotients are taken once, at the end of eval()
    template <class S0, class S1, class S2, class S3>
    static __device__ __forceinline__ void qqxtbx(const S0& p0, const S1& p1, const S2& p2,
                                                  const S3& p3, double c, double mt2, double mw2,
                                                  double gaw2, double& num, double& wprop) {
        const double pw2 = sprod(p0, p1);
        const double d0 = pw2 - mw2;
        wprop = d0 * d0 + mw2 * gaw2;
        const double a = sprod(p0, p2);
        const double b = sprod(p0, p3);
        const double d = sprod(p3, p1);
        num = fabs((a * c + mt2 * b) * d);
    }

    template <int NDIM>
    static __device__ double eval(const double (&xa)[NDIM], const IntegrandConsts&) {
        static_assert(NDIM == 3, "singletop_lo is 3-dimensional");
        // constants :19-42
        const double mt = 173.2, sqrts = 8000.0, sqrtsmin = 173.2, mw = 80.419, gaw = 2.1054,
                     gf = 1.16639e-5;
        const double mt2 = mt * mt;
        const double s = sqrts * sqrts;
        const double smin = sqrtsmin * sqrtsmin;
        const double bmax = sqrt(1 - smin / s);
        const double conv = 0.3893793e9;
        const double gaw2 = gaw * gaw;
        const double mw2 = mw * mw;
        const double g = 4 * 1.4142135623730951 * mw2 * gf;
        const double gw4 = g * g;
        // get_x1x2 :45-68
        const double b = bmax * xa[0];
        const double onemb2 = 1 - b * b;
        const double shat = smin / onemb2;
        const double tau = shat / s;
        const double ymax = -0.5 * log(tau);
        const double y = ymax * (2 * xa[1] - 1);
        // jac = 2 tau b bmax / onemb2 * (2 ymax) (:57-58): numerator here, onemb2 in the final quotient
        const double jac_num = 2 * tau * b * bmax * (2 * ymax);
        const double sqrttau = sqrt(tau);
        const double expy = exp(y);
        const double x1 = sqrttau * expy;
        const double x2 = sqrttau / expy;
        // make_event :71-92
        const double ecmo2 = sqrt(shat) / 2;
        const double one_m = 1 - mt2 / shat;  // used twice (:78, :88)
        const double cc = ecmo2 * one_m;
        const double cosv = 1 - 2 * xa[2];
        const double sinxi = cc * sqrt(1 - cosv * cosv);
        const double cosxi = cc * cosv;
        const Mom p0{ecmo2, 0.0, 0.0, ecmo2};
        const Mom p1{ecmo2, 0.0, 0.0, -ecmo2};
        const Mom p2{cc, sinxi, 0.0, cosxi};
        Mom p3{sqrt(cc * cc + mt2), -sinxi, 0.0, -cosxi};
        // massless projection :236-239; dot :95-102
        const double dot30 = p3.e * p0.e - p3.x * p0.x - p3.y * p0.y - p3.z * p0.z;
        const double k = mt2 / dot30 / 2;
        p3 = Mom{p3.e - p0.e * k, p3.x - p0.x * k, p3.y - p0.y * k, p3.z - p0.z * k};
        // channels :242-245
        // -p1 = (-ecmo2, 0, 0, +ecmo2): theta = pi; -p0 = (-ecmo2, 0, 0, -ecmo2): theta = 0
        const double bq = kSqrt2 * sqrt(ecmo2);  // prefactor of the beam spinors: i*bq
        const AllSpin A = spinors(p2), Cc = spinors(p3);
        const BeamSpinPi B = beam_spinors_thetapi(bq);
        const BeamSpin0 D = beam_spinors_theta0(bq);
        const double c = sprod(Cc, D);
        double n1, w1, n2, w2;
        qqxtbx(A, B, Cc, D, c, mt2, mw2, gaw2, n1, w1);
        qqxtbx(B, A, Cc, D, c, mt2, mw2, gaw2, n2, w2);
        // channel sum c1 + c2 = K (n1/w1 + n2/w2)/|c|, K = 9 gw4/36; luminosities :254-260
        // lumi1 = lumi2 = 2 pdf/x1/x2; psw = one_m/(8 pi) * jac (:88-89); flux = 1/(2 shat) (:90);
        // result = (2 lumi1 c1 + 2 lumi2 c2) psw flux conv (:267-268).  All of these quotients are
        // regrouped into one numerator and one denominator (a few ulp; parity bar 1e-12).
        const double pdf = x1 * x2;
        const double num = 2 * (pdf + pdf) * (9.0 * gw4 / 36) * (n1 * w2 + n2 * w1) *
                           (one_m * (1.0 / (8.0 * M_PI))) * jac_num * conv;
        const double den = (x1 * x2) * (fabs(c) * (w1 * w2)) * onemb2 * (2 * shat);
        return num / den;
    }
};

}  // namespace vf
