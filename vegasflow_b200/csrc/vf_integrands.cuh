// Built-in integrands, evaluated inline in the fused event kernel.
// Every function follows the reference's operation order; the translation
// units are compiled with -fmad=false so each multiply/add is rounded
// separately, as TensorFlow's unfused elementwise ops are.
// Reference citations are file:line relative to /root/reference.
#pragma once
#include "vf_common.cuh"

// block sizes of the matrix-element kernels (tunable at build time; measured defaults)
#ifndef VF_DY_THREADS
#define VF_DY_THREADS 1024
#endif
#ifndef VF_ST_THREADS
#define VF_ST_THREADS 640
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
// Components that the reference sets to exact complex zero are dropped: adding
// or multiplying by exact zero does not change the remaining terms.
// ---------------------------------------------------------------------------
struct cplx {
    double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx cscale(cplx a, double r) { return {a.re * r, a.im * r}; }  // a*complex(r,0)
__device__ __forceinline__ cplx cneg(cplx a) { return {-a.re, -a.im}; }
// |a| without hypot's scaling: the amplitudes are far from the overflow/underflow thresholds
__device__ __forceinline__ double cabs(cplx a) { return sqrt(a.re * a.re + a.im * a.im); }

struct Mom {
    double e, x, y, z;
};
__device__ __forceinline__ Mom mneg(Mom p) { return {-p.e, -p.x, -p.y, -p.z}; }

struct Angles {
    double ch, sh;   // cos(theta/2), sin(theta/2)
    double cp, sp;   // cos(phi), sin(phi)
    cplx pref;       // sqrt(2)*sqrt(complex(p0,0))
};

// sqrt(2)*sqrt(complex(p0, 0)): principal branch, +0 imaginary part.
__device__ __forceinline__ cplx spinor_prefact(double p0) {
    const double r2 = 1.4142135623730951;  // np.sqrt(2)
    if (p0 >= 0.0) return {r2 * sqrt(p0), 0.0};
    return {0.0, r2 * sqrt(-p0)};
}

__device__ __forceinline__ double clip1(double v) {
    // tf.where(v < -1, -1, v); tf.where(v > 1, 1, .)  (NaN passes through)
    double r = v;
    if (v < -1.0) r = -1.0;
    if (v > 1.0) r = 1.0;
    return r;
}

// The reference takes theta = acos(pz/E), phi = +-acos(px/E/sin(theta)) and then needs only
// cos(theta/2), sin(theta/2), cos(phi), sin(phi) (drellyan :93-162, singletop :105-178).  With
// c = clip(pz/E) and cx = clip(px/E/sin(theta)) these are, for the SAME c and cx,
//     cos(theta/2) = sqrt((1+c)/2)   sin(theta/2) = sqrt((1-c)/2)   sin(theta) = 2 sin cos (theta/2)
//     cos(phi) = cx                  sin(phi) = +-sqrt((1-cx)(1+cx))
// i.e. four square roots instead of two acos, one sin and two sincos per momentum -- accurate
// evaluations of the same functions of the same arguments (1-c, 1-cx are exact near the ends
// where the acos route is accurate as well), so both stay inside the 1e-12 parity bar
// (tests/test_device_source_on_host.py, tests/test_parity_gpu.py).  The p.x == 0 branch keeps
// the reference's numerically evaluated constants: cos(pi/2) = 6.123e-17, sin(pi) = 1.2246e-16.
constexpr double kCosHalfPi = 6.123233995736766e-17;   // np.cos(np.pi / 2)
constexpr double kSinPi = 1.2246467991473532e-16;      // np.sin(np.pi)

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

// theta/phi of drellyan u0 (:93-115) and ubar0 (:140-162) through the half-angle forms.
__device__ __forceinline__ Angles angles_half(const Mom& p) {
    const double rz = p.z / p.e;
    Angles a;
    if (p.x == 0.0) {
        half_angle_axis(rz, a.ch, a.sh);
        a.cp = 1.0;  // phi = 0
        a.sp = 0.0;
    } else {
        const double c = clip1(rz);
        a.ch = sqrt(0.5 * (1.0 + c));
        a.sh = sqrt(0.5 * (1.0 - c));
        const double sin_theta = 2.0 * (a.sh * a.ch);
        const double cx = clip1(p.x / p.e / sin_theta);
        a.cp = cx;
        a.sp = sqrt((1.0 - cx) * (1.0 + cx));
        if (cx == -1.0) a.sp = kSinPi;  // phi == pi: the reference's sin(np.pi)
        if (p.e > 0.0 ? (p.y < 0.0) : (p.y / p.e < 0.0)) a.sp = -a.sp;  // sign of py/E
        if (cx == 1.0) a.sp = 0.0;  // phi == 0 (also -0): cos 0 = 1, sin 0 = 0
    }
    a.pref = spinor_prefact(p.e);
    return a;
}

// The literal route (acos, sin, sincos) for single-top's ubar0 (singletop :156-178).  Single-top
// keeps it: near threshold the projected top momentum is anti-parallel to the beam, theta -> pi,
// and the reference's cos(theta/2) carries the rounding of theta = acos(.) as a RELATIVE error of
// 2e-16 / cos(theta/2) that reaches 1e-8 -- amplitudes proportional to it agree with the
// reference to 1e-12 only if theta is rounded the same way (0.4 % of uniformly drawn events would
// miss the bar with the more accurate half-angle forms).  Drell-Yan has no such region.
// `pxe` returns px/E (0 on the beam axis) for the caller that needs its sign again.
__device__ __forceinline__ Angles angles_acos(const Mom& p, double* pxe = nullptr) {
    const double rz = p.z / p.e;
    Angles a;
    a.cp = 1.0;  // phi == 0 (beam-axis momenta, and phi2 == 0): cos 0 = 1, sin 0 = 0
    a.sp = 0.0;
    double rxe = 0.0;
    if (p.x == 0.0) {
        half_angle_axis(rz, a.ch, a.sh);  // theta1 in {0, pi}: sincos(theta/2) are constants
    } else {
        const double theta = acos(clip1(rz));
        rxe = p.x / p.e;
        const double rx = rxe / sin(theta);
        double phi = acos(clip1(rx));
        // py/E < 0: with E > 0 (every physical momentum) that is py < 0, no division needed;
        // the quotient is formed only for E <= 0 or NaN, where its sign rules apply
        const bool neg = p.e > 0.0 ? (p.y < 0.0) : (p.y / p.e < 0.0);
        if (neg) phi = -phi;
        sincos(theta / 2, &a.sh, &a.ch);  // one argument reduction for both
        if (phi != 0.0) sincos(phi, &a.sp, &a.cp);
    }
    if (pxe) *pxe = rxe;
    a.pref = spinor_prefact(p.e);
    return a;
}

struct Spin2 {
    cplx a, b;  // the two non-zero components
};
// u0(p, +1): (pref*cos, pref*sin*e^{+i phi}, 0, 0)
__device__ __forceinline__ Spin2 u0_plus(const Angles& g) {
    return {cscale(g.pref, g.ch), cmul(cscale(g.pref, g.sh), cplx{g.cp, g.sp})};
}
// u0(p, -1): (0, 0, pref*sin*e^{-i phi}, -pref*cos)
__device__ __forceinline__ Spin2 u0_minus(const Angles& g) {
    return {cmul(cscale(g.pref, g.sh), cplx{g.cp, -g.sp}), cscale(cneg(g.pref), g.ch)};
}
// ubar0(p, -1): (pref*sin*e^{+i phi}, -pref*|cos|, 0, 0)
__device__ __forceinline__ Spin2 ubar0_minus(const Angles& g) {
    return {cmul(cscale(g.pref, g.sh), cplx{g.cp, g.sp}), cscale(cneg(g.pref), fabs(g.ch))};
}
// ubar0(p, +1): (0, 0, pref*cos, pref*sin*e^{-i phi})
__device__ __forceinline__ Spin2 ubar0_plus(const Angles& g) {
    return {cscale(g.pref, g.ch), cmul(cscale(g.pref, g.sh), cplx{g.cp, -g.sp})};
}
__device__ __forceinline__ cplx sdot(const Spin2& bra, const Spin2& ket) {
    return cadd(cmul(bra.a, ket.a), cmul(bra.b, ket.b));
}

// Beam-axis momenta.  Both examples build the spinors of -p1 = (-E, 0, 0, +E) and
// -p0 = (-E, 0, 0, -E) with E = ecmo2 > 0: pz/E' is exactly -1 / +1, so theta is pi / 0, phi is
// 0, and the generic construction multiplies by (cos phi, sin phi) = (1, 0), by sin(theta/2) = 1
// or 0 and by cos(theta/2) = 6.12e-17 or 1 at run time.  Written out, the theta = 0 spinors have
// ONE non-zero component (the other is pref * 0: an exact zero whose products and sums change
// nothing), so every sdot() with them is one complex product instead of two.  Same values, bit
// for bit, as the generic route (checked by the parity tests and the full-size checksums).
struct SpinA {
    cplx v;  // (v, 0)
};
struct SpinB {
    cplx v;  // (0, v)
};
__device__ __forceinline__ cplx sdot(const Spin2& bra, const SpinA& ket) { return cmul(bra.a, ket.v); }
__device__ __forceinline__ cplx sdot(const Spin2& bra, const SpinB& ket) { return cmul(bra.b, ket.v); }
__device__ __forceinline__ cplx sdot(const SpinA& bra, const Spin2& ket) { return cmul(bra.v, ket.a); }
__device__ __forceinline__ cplx sdot(const SpinB& bra, const Spin2& ket) { return cmul(bra.v, ket.b); }
struct BeamSpin0 {   // theta = 0: u0(+1), u0(-1), ubar0(+1), ubar0(-1) of (E', 0, 0, E'), E' < 0
    SpinA up;
    SpinB um;
    SpinA bp;
    SpinB bm;
};
struct BeamSpinPi {  // theta = pi: of (E', 0, 0, -E')
    Spin2 up, um, bp, bm;
};
__device__ __forceinline__ BeamSpin0 beam_spinors_theta0(cplx pref) {
    return {SpinA{pref}, SpinB{cneg(pref)}, SpinA{pref}, SpinB{cneg(pref)}};
}
__device__ __forceinline__ BeamSpinPi beam_spinors_thetapi(cplx pref) {
    const cplx small = cscale(pref, kCosHalfPi), nsmall = cscale(cneg(pref), kCosHalfPi);
    return {Spin2{small, pref}, Spin2{pref, nsmall}, Spin2{small, pref}, Spin2{pref, nsmall}};
}

// ---------------------------------------------------------------------------
// Drell-Yan LO, examples/drellyan_lo_tf.py:27-249 (n_dim = 4)
// ---------------------------------------------------------------------------
struct DrellYanLO {
    static constexpr int kFixedDim = 4;
    static constexpr bool kHeavy = true;
    // measured (profiles/r2_me_threads.txt): 1024 threads x 64 registers beat 512 x 116 by 14 %
    // since the half-angle rewrite shortened the live ranges
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
        const double mV = sqrt(shat * x1 * x2);
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
        const double root = sqrt(mV2);
        const double ptmax = 0.5 * mV2 / root;
        const double pta = ptmax * xa[2];
        const double ptx = pta * cphi, pty = pta * sphi;
        const double Delta = mV2 / 2.0 / pta / root;
        // yy = YV - acosh(Delta) = -acosh(Delta) (:64): cosh(yy) = Delta and
        // sinh(yy) = -sqrt((Delta-1)(Delta+1)), with Delta - 1 exact (Delta = 1/x2 > 1)
        const double shy = sqrt((Delta - 1.0) * (Delta + 1.0));
        const double kallenF = 2.0 * ptmax / root / shy;
        const Mom p2{pta * Delta, ptx, pty, -(pta * shy)};
        const Mom p3{pV.e - p2.e, pV.x - p2.x, pV.y - p2.y, pV.z - p2.z};
        double psw = (1.0 / (8.0 * M_PI)) * kallenF;  // :71, 1/(8*np.pi) folded in IEEE
        psw = psw * jac;
        const double flux = 1 / (2 * mV2);
        // qqxllx(-p1, -p0, p2, p3) :207-224
        // q0 = -p1 = (-ecmo2, 0, 0, +ecmo2): theta = pi; q1 = -p0 = (-ecmo2, 0, 0, -ecmo2): theta = 0
        const cplx bpref = spinor_prefact(-ecmo2);
        const BeamSpinPi s0 = beam_spinors_thetapi(bpref);
        const BeamSpin0 s1 = beam_spinors_theta0(bpref);
        const Angles a2 = angles_half(p2), a3 = angles_half(p3);
        // za(a,b) = ubar0(a,-1).u0(b,+1); zb(a,b) = ubar0(a,+1).u0(b,-1)
        const Spin2 ubm0 = s0.bm;
        const cplx za01 = sdot(ubm0, s1.up);
        const cplx zb10 = sdot(s1.bp, s0.um);
        const cplx sp = cmul(za01, zb10);
        const double lsprod = sp.re;  // sprod(p0,p1) :200-204
        const cplx za02 = sdot(ubm0, u0_plus(a2));
        const cplx za03 = sdot(ubm0, u0_plus(a3));
        const SpinB u1m = s1.um;
        const cplx zb31 = sdot(ubar0_plus(a3), u1m);
        const cplx zb21 = sdot(ubar0_plus(a2), u1m);
        const double a = 2 * cabs(cmul(za02, zb31)) / lsprod;
        const double b = 2 * cabs(cmul(za03, zb21)) / lsprod;
        const double wgts = 6.0 * (a * a + b * b) / 36.0;
        // build_luminosity :232-239 with the toy pdf x1*x2 :226-229
        const double pdf = x1 * x2;
        const double lumis = (pdf + pdf + pdf + pdf) / x1 / x2;
        const double lumi_me2 = 2 * lumis * wgts;  // :246
        return lumi_me2 * psw * flux * conv;       // :247
    }
};

// ---------------------------------------------------------------------------
// single-top LO (t-channel), examples/singletop_lo_tf.py:45-270 (n_dim = 3)
// ---------------------------------------------------------------------------
struct SingleTopLO {
    static constexpr int kFixedDim = 3;
    static constexpr bool kHeavy = true;
    static constexpr int kBlockThreads = VF_ST_THREADS;  // 640 x 96 registers: +1.4 % over 512 x 128

    struct AllSpin {
        Spin2 up, um, bp, bm;  // u0(+1), u0(-1), ubar0(+1), ubar0(-1)
    };
    // u0 and ubar0 of one momentum share theta (singletop :110-129 and :156-178 both start from
    // acos(clip(pz/E)) and need sincos(theta/2)): evaluated once, the same operations
    static __device__ __forceinline__ AllSpin spinors(const Mom& p) {
        double pxe;
        Angles gb = angles_acos(p, &pxe);
        Angles gu = gb;      // same theta -> same (ch, sh, pref)
        gu.cp = 1.0;         // phi of u0 is 0 or pi from the sign of px/E (:124-126)
        gu.sp = 0.0;
        if (pxe < 0.0) {     // (px/E is 0 on the beam axis)
            gu.cp = -1.0;    // the reference's np.cos(np.pi), np.sin(np.pi)
            gu.sp = kSinPi;
        }
        return {u0_plus(gu), u0_minus(gu), ubar0_plus(gb), ubar0_minus(gb)};
    }
    // sprod(p1,p2) = Re(za(p1,p2)*zb(p2,p1)) :213-218 (any mix of generic and beam spinor sets)
    template <class S1, class S2>
    static __device__ __forceinline__ double sprod(const S1& s1, const S2& s2) {
        const cplx za = sdot(s1.bm, s2.up);
        const cplx zb = sdot(s2.bp, s1.um);
        return cmul(za, zb).re;
    }
    // qqxtbx :221-230
    template <class S0, class S1, class S2, class S3>
    static __device__ __forceinline__ double qqxtbx(const S0& p0, const S1& p1, const S2& p2,
                                                    const S3& p3, double mt2, double mw2,
                                                    double gaw2, double gw4) {
        const double pw2 = sprod(p0, p1);
        const double d0 = pw2 - mw2;
        const double wprop = d0 * d0 + mw2 * gaw2;
        const double a = sprod(p0, p2);
        const double b = sprod(p0, p3);
        const double c = sprod(p2, p3);
        const double d = sprod(p3, p1);
        return fabs((a + mt2 * b / c) * d) * 9.0 / wprop * gw4 / 36;
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
        double jac = 2 * tau * b * bmax / onemb2;
        jac = jac * (2 * ymax);
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
        double psw = one_m / (8.0 * M_PI);  // :88
        psw = psw * jac;
        const double flux = 1 / (2 * shat);
        // massless projection :236-239; dot :95-102
        const double dot30 = p3.e * p0.e - p3.x * p0.x - p3.y * p0.y - p3.z * p0.z;
        const double k = mt2 / dot30 / 2;
        p3 = Mom{p3.e - p0.e * k, p3.x - p0.x * k, p3.y - p0.y * k, p3.z - p0.z * k};
        // channels :242-245
        // -p1 = (-ecmo2, 0, 0, +ecmo2): theta = pi; -p0 = (-ecmo2, 0, 0, -ecmo2): theta = 0
        const cplx bpref = spinor_prefact(-ecmo2);
        const AllSpin A = spinors(p2), Cc = spinors(p3);
        const BeamSpinPi B = beam_spinors_thetapi(bpref);
        const BeamSpin0 D = beam_spinors_theta0(bpref);
        const double c1 = qqxtbx(A, B, Cc, D, mt2, mw2, gaw2, gw4);
        const double c2 = qqxtbx(B, A, Cc, D, mt2, mw2, gaw2, gw4);
        // luminosities :254-260
        const double pdf = x1 * x2;
        const double lumi1 = (pdf + pdf) / x1 / x2;
        const double lumi2 = lumi1;  // the same expression in the reference (:257-260)
        const double lumi_me2 = 2 * lumi1 * c1 + 2 * lumi2 * c2;  // :267
        return lumi_me2 * psw * flux * conv;                       // :268
    }
};

}  // namespace vf
