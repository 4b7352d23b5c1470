// vfk_math.cuh -- per-thread math building blocks of the fused control-cycle kernel.
//
// Everything here works on values held in registers by ONE thread for ONE instance
// (fully unrolled, compile-time indexed).  Precision-specific primitives live in
// Prec<float> / Prec<double>: the FP32 path uses single MUFU approximations (rsqrt / rcp /
// sqrt / lg2 / ex2) whose error budget is stated in DESIGN.md (1e-4 relative on qdot); the
// FP64 path builds reciprocal, rsqrt, sqrt and division from the hardware seed plus one
// third-order step (branch-free, a few ulp), decay orders from multiplication chains and
// sin / cos from its own reduction + polynomials -- only atan2 and non-integer pow are
// libdevice (1e-9 relative on qdot).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>

namespace vfk {

template <typename T> struct Prec;

// FP32: one MUFU instruction per transcendental (rsqrt / rcp / sqrt / lg2 / ex2 .approx.ftz,
// max relative error ~2^-22), no IEEE division or sqrt subroutines in the hot path.
template <> struct Prec<float> {
    static __device__ __forceinline__ float rsqrt_pos(float x) {
        float r;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float sqrt_(float x) {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float rcp(float x) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
        return r;
    }
    static __device__ __forceinline__ float div(float a, float b) { return a * rcp(b); }
    // x^y for x >= 0, y > 0 through ex2(y * lg2(x)); x = 0 -> lg2 = -inf -> 0.
    static __device__ __forceinline__ float pow_pos(float x, float y) {
        float l, r;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
        l *= y;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l));
        return r;
    }
    // sin/cos of a joint angle.  Default: Cody-Waite reduction to [-pi/4, pi/4] (k = rint(2x/pi) taken from
    // the mantissa of a magic-number add) + Cephes minimax polynomials, ~1 ulp, branch-free, valid for
    // |x| < 1e4 rad.  -DVFK_SINCOS=1: libdevice sincosf.  -DVFK_SINCOS=2: MUFU.SIN/COS (2^-21.4 absolute; measured to
    // push 5e-5 of random LWR instances past the 1e-4 qdot tolerance, so it is NOT the default).
    static __device__ __forceinline__ void sincos_(float x, float* s, float* c) {
#if defined(VFK_SINCOS) && VFK_SINCOS == 1
        sincosf(x, s, c);
#elif defined(VFK_SINCOS) && VFK_SINCOS == 2
        __sincosf(x, s, c);
#else
        const float t = fmaf(x, 0.636619772367581343f, 12582912.0f);
        const int ki = __float_as_int(t);
        const float kf = t - 12582912.0f;
        float r = fmaf(kf, -1.57079637050628662109375f, x);
        r = fmaf(kf, 4.37113900018624283e-8f, r);
        const float z = r * r;
        const float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f), z, -1.6666654611e-1f), z * r, r);
        const float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f), z, 4.166664568298827e-2f),
                              z * z, fmaf(-0.5f, z, 1.0f));
        const float ss = (ki & 1) ? cp : sp;
        const float cs = (ki & 1) ? sp : cp;
        *s = __int_as_float(__float_as_int(ss) ^ ((ki & 2) << 30));
        *c = __int_as_float(__float_as_int(cs) ^ (((ki + 1) & 2) << 30));
#endif
    }
    static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
    static __device__ __forceinline__ float tiny() { return 1e-30f; }
    static __device__ __forceinline__ float fmin_(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float fmax_(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float fabs_(float a) { return fabsf(a); }
    static __device__ __forceinline__ float big() { return 1e18f; }
    static constexpr bool kSeriesAsin = true;
};

// FP64 reciprocal / reciprocal square root / square root / division without libdevice's special-case branches: the
// ~2^-20 hardware seed (MUFU.RCP64H / RSQ64H, PTX rcp/rsqrt.approx.ftz.f64) refined by one third-order step (error
// ~e0^3 < 2^-58) and, for the compound results, one residual correction -- a few ulp at most, straight-line code.  IEEE
// `/`, sqrt() and rsqrt() each bring a fast path of the same length PLUS a guarded slow-path call, and the kernel has ~40
// such sites: 63 CALLs, 119 BSSY/BSYNC pairs and ~500 argument moves in the FP64 cycle kernel before this.
// Domain: rcp / div / sqrt clamp the magnitude into [1e-300, inf) first, so 0 behaves like 1e-300 (1/0 -> 1e300,
// sqrt(0) -> 0 exactly); rsqrt_pos wants x > 0 and normal -- its call sites carry a 1e-300 bias or guard zero themselves.
__device__ __forceinline__ double rcp_seed(double x) { double y; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double rsqrt_seed(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
__device__ __forceinline__ double rcp_fast(double x) {
    x = copysign(fabs(x) > 1e-300 ? fabs(x) : 1e-300, x);
    const double y = rcp_seed(x);
    const double e = fma(-x, y, 1.0);                 // 1 - x y
    return fma(y, fma(e, e, e), y);                   // y (1 + e + e^2)
}
__device__ __forceinline__ double rsqrt_fast(double x) {          // x > 0 and normal (call sites bias or guard)
    const double y = rsqrt_seed(x);
    const double e = fma(-(x * y), y, 1.0);           // 1 - x y^2
    return fma(y * e, fma(0.375, e, 0.5), y);         // y (1 + e/2 + 3 e^2/8)
}

template <> struct Prec<double> {
    static __device__ __forceinline__ double rsqrt_pos(double x) { return rsqrt_fast(x); }
    static __device__ __forceinline__ double sqrt_(double x) {
        const double y = rsqrt_fast(x > 1e-300 ? x : 1e-300);
        const double s = x * y;
        return fma(fma(-s, s, x), 0.5 * y, s);        // one Newton correction on the product
    }
    static __device__ __forceinline__ double rcp(double x) { return rcp_fast(x); }
    static __device__ __forceinline__ double div(double a, double b) {
        const double y = rcp_fast(b);
        const double q = a * y;
        return fma(fma(-b, q, a), y, q);      // q + (a - b q) y  (b = 0: a - 0 q = a, q + a * 1e300: still huge)
    }
    // x^y for x >= 0, y > 0.  Decay orders are small integers in practice (20, 5, 3, 2: scripts/object_feeder:274-333):
    // square-and-multiply costs ~2 log2(y) DMULs instead of libdevice pow's ~100 instructions and is exact to a few ulp.
    static __device__ __forceinline__ double pow_pos(double x, double y) {
        const int n = (int)y;
        if ((double)n == y && n <= 64) {
            double r = 1.0, b = x;
            for (int k = n; k > 0; k >>= 1) {
                if (k & 1) r *= b;
                b *= b;
            }
            return r;
        }
        return pow(x, y);
    }
    static __device__ __forceinline__ void sincos_(double x, double* s, double* c) { sincos(x, s, c); }
    static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
    // compare-and-select: fmin() / fmax() add NaN-propagation fix-ups (DSETP + 2 SEL + LOP3) the call sites do not need
    static __device__ __forceinline__ double fmin_(double a, double b) { return a < b ? a : b; }
    static __device__ __forceinline__ double fmax_(double a, double b) { return a > b ? a : b; }
    static __device__ __forceinline__ double fabs_(double a) { return fabs(a); }
    static __device__ __forceinline__ double big() { return 1e150; }
    static __device__ __forceinline__ double tiny() { return 1e-300; }
    static constexpr bool kSeriesAsin = false;
};

// Working types of the two ill-conditioned stages.  Measured on B200 over 1,048,576 random LWR instances with 32
// obstacles (scripts/fp32_error.py, FP32 mode, max relative qdot error / launch time):
//     pure FP32                         1.2e-4   106 us     (-DVFK_PURE_FP32)  -- 2 instances in 10^6 beyond 1e-4
//     FK in double, normal eq. FP32     6.3e-5   111 us     (default)
//     FK in FP32, normal eq. double     1.2e-4   115 us     (-DVFK_WIDE_NE_ONLY)
//     both in double                    1.6e-5   127 us     (-DVFK_WIDE_NE)
// An FP32 kinematic chain leaves ~1e-7 m in the tool position, which the order-20 repeller decay turns into up to 1e-4
// of relative field error next to an obstacle -- the dominant tail.  Forming J J^T + lambda^2 I in FP32 perturbs its
// smallest eigenvalues (~lambda^2 = 0.01) by ~1e-6, the second, smaller tail.  Everything that touches HBM stays T.
#if defined(VFK_PURE_FP32)
template <typename T> struct WideOf { using type = T; };
template <typename T> struct WideNE { using type = T; };
#elif defined(VFK_WIDE_NE_ONLY)
template <typename T> struct WideOf { using type = T; };
template <typename T> struct WideNE { using type = double; };
#elif defined(VFK_WIDE_NE)
template <typename T> struct WideOf { using type = double; };
template <typename T> struct WideNE { using type = double; };
#else
template <typename T> struct WideOf { using type = double; };      // forward kinematics
template <typename T> struct WideNE { using type = T; };           // normal equations
#endif

// sin / cos in double for joint angles: Cody-Waite reduction to [-pi/4, pi/4] with k taken from the mantissa of a
// magic-number add, Taylor polynomials, branch-free quadrant selection.  The FP32 mode's wide kinematic chain stops at
// x^11 / x^12 (truncation < 7e-12 and < 4e-13 at pi/4: four decades below an FP32 ulp, which is all its consumer keeps).  The coefficients sit in constant memory so every DFMA takes its coefficient as a
// constant-bank operand; as immediates they cost two UMOV per use in the 10- and 17-joint kernels.
struct SinCosTab {
    double two_over_pi, magic, pio2_hi, pio2_lo;
    double s[7];     // -1/3!, 1/5!, -1/7!, 1/9!, -1/11!, 1/13!, -1/15!
    double c[7];     // 1/4!, -1/6!, 1/8!, -1/10!, 1/12!, -1/14!, 1/16!
    // sincos_quarter: 1 / 2 pi, the 53-bit 2 pi, and the coefficients of sin(r / 4) / r - 1/4 and (cos(r / 4) - 1 + r^2 / 32) / r^4
    // as polynomials in r^2: sq[i] = s[i] / (4 * 16^(i + 1)), cq[i] = c[i] / 16^(i + 2) (exact scalings, filled by sincos_tab())
    double one_over_two_pi, two_pi_hi;
    double sq[7], cq[7];
    double tab_inv_h, tab_h;    // sincos_table: 128 / 2 pi and the 53-bit 2 pi / 128
};
// The table travels inside the kernel's parameter block (KConst::sincos, constant bank 0) like every other constant: as a
// __constant__ variable (bank 3) its first use in every tile was a constant-cache miss -- 14 % of the FP64 kernel's stall
// samples sat on the one DFMA behind that LDC (ncu source page, r02m config 2).
#define VFK_SINCOS_TAB_INIT                                                                                                      \
    {0.63661977236758134308, 6755399441055744.0, 1.57079632679489655800e+00, 6.12323399573676603587e-17,                         \
     {-1.66666666666666666667e-01, 8.33333333333333333333e-03, -1.98412698412698412698e-04, 2.75573192239858906526e-06,          \
      -2.50521083854417187751e-08, 1.60590438368216145994e-10, -7.64716373181981647590e-13},                                     \
     {4.16666666666666666667e-02, -1.38888888888888888889e-03, 2.48015873015873015873e-05, -2.75573192239858906526e-07,          \
      2.08767569878680989792e-09, -1.14707455977297247139e-11, 4.77947733238738529744e-14},                                      \
     0.15915494309189533577, 4.0 * 1.57079632679489655800e+00, {0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0},           \
     128.0 * 0.15915494309189533577, 4.0 * 1.57079632679489655800e+00 / 128.0}
inline SinCosTab sincos_tab() {
    SinCosTab t = VFK_SINCOS_TAB_INIT;
    double p = 1.0 / 16.0;
    for (int i = 0; i < 7; ++i) {
        t.sq[i] = t.s[i] * p * 0.25;
        p *= 1.0 / 16.0;
        t.cq[i] = t.c[i] * p;
    }
    return t;
}

// TERMS = 5: x^11 / x^12 (the FP32 mode's wide chain); TERMS = 7: x^15 / x^16, truncation < 5e-17 (FP64 mode).
template <int TERMS>
__device__ __forceinline__ void sincos_wide(const SinCosTab& k, double x, double* s, double* c) {
    const double t = fma(x, k.two_over_pi, k.magic);
    const int ki = __double2loint(t);
    const double kf = t - k.magic;
    double r = fma(kf, -k.pio2_hi, x);
    r = fma(kf, -k.pio2_lo, r);
    const double z = r * r;
    double sp = k.s[TERMS - 1], cp = k.c[TERMS - 1];
#pragma unroll
    for (int i = TERMS - 2; i >= 0; --i) {
        sp = fma(sp, z, k.s[i]);
        cp = fma(cp, z, k.c[i]);
    }
    sp = fma(sp * z, r, r);
    cp = fma(cp, z, -0.5);
    cp = fma(cp, z, 1.0);
    const double ss = (ki & 1) ? cp : sp;
    const double cs = (ki & 1) ? sp : cp;
    *s = (ki & 2) ? -ss : ss;
    *c = ((ki + 1) & 2) ? -cs : cs;
}
// The same without quadrant logic: take whole turns off x (r = x - 2 pi k, k = rint(x / 2 pi), |r| <= pi), evaluate the
// polynomials at the QUARTER angle u = r / 4 (|u| <= pi / 4) and double the angle twice (sin 2a = 2 sin a cos a,
// cos 2a = 1 - 2 sin^2 a).  The division by four lives in the coefficients (sq, cq: the Taylor coefficients times exact
// powers of 1/16, polynomials in r^2), so it costs nothing.  Five more FP64 operations than sincos_wide but none of its ~15
// integer / select instructions per joint (the 64-bit swaps and sign flips of the quadrant selection); the two doublings
// multiply the absolute error by < 10: 4e-11 for TERMS = 5 (checked against long double over [-12, 12]; 2 pi k is taken off
// with the 53-bit 2 pi only, k * 2.4e-16 more), which is why only the FP32 mode's wide chain -- whose consumer keeps 6e-8 -- uses it.
template <int TERMS>
__device__ __forceinline__ void sincos_quarter(const SinCosTab& k, double x, double* s, double* c) {
    const double t = fma(x, k.one_over_two_pi, k.magic);
    const double kf = t - k.magic;
    const double r = fma(kf, -k.two_pi_hi, x);
    const double z = r * r;
    double sp = k.sq[TERMS - 1], cp = k.cq[TERMS - 1];
#pragma unroll
    for (int i = TERMS - 2; i >= 0; --i) {
        sp = fma(sp, z, k.sq[i]);
        cp = fma(cp, z, k.cq[i]);
    }
    double sn = r * fma(sp, z, 0.25);                         // sin(r / 4)
    double cn = fma(fma(cp, z, -0.03125), z, 1.0);            // cos(r / 4)
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        const double u = sn + sn;
        const double c2 = fma(-u, sn, 1.0);
        sn = u * cn;
        cn = c2;
    }
    *s = sn;
    *c = cn;
}
// Table form (the FP32 mode's wide chain; vfk_kernels.cuh: TAB): x = k h + r with h = 2 pi / 128, |r| <= h / 2 = 0.0245;
// {sin(k h), cos(k h)} comes from a 128-entry table in shared memory (2 KB per CTA, filled once per CTA by
// sincos_table_fill), sin r and cos r from r - r^3/6 + r^5/120 and 1 - r^2/2 + r^4/24 (truncation < 1e-15 and < 3e-13),
// and the angle sum puts them together: 13 FP64 operations, one LOP3, one address and one LDS.128 per joint against the
// quarter-angle form's 23 FP64 operations, no quadrant logic either, and a dependent chain half as long.
constexpr int kSinCosTabEntries = 128;
__device__ __forceinline__ void sincos_table(const SinCosTab& k, const double2* tab, double x, double* s, double* c) {
    const double t = fma(x, k.tab_inv_h, k.magic);
    const int ki = __double2loint(t);
    const double kf = t - k.magic;
    const double r = fma(kf, -k.tab_h, x);
    const double2 sc = tab[ki & (kSinCosTabEntries - 1)];
    const double z = r * r;
    const double sr = fma(r * z, fma(z, 1.0 / 120.0, -1.0 / 6.0), r);
    const double cr = fma(z, fma(z, 1.0 / 24.0, -0.5), 1.0);
    *s = fma(sc.x, cr, sc.y * sr);
    *c = fma(sc.y, cr, -(sc.x * sr));
}
// entry k = {sin, cos}(2 pi k / 128) to FP64 accuracy (the x^15 / x^16 polynomials); the caller synchronises the CTA afterwards
__device__ __forceinline__ void sincos_table_fill(const SinCosTab& k, double2* tab, int tid, int nthreads) {
    for (int e = tid; e < kSinCosTabEntries; e += nthreads) {
        double s, c;
        sincos_wide<7>(k, (double)e * k.tab_h, &s, &c);
        tab[e] = make_double2(s, c);
    }
}
template <int TERMS>
__device__ __forceinline__ void sincos_wide(const SinCosTab&, float x, float* s, float* c) { Prec<float>::sincos_(x, s, c); }

// x^ORDER by a fixed multiplication chain (FP64 repeller fast path; ORDER = 0 means "use Prec<T>::pow_pos").
template <int ORDER, typename T>
__device__ __forceinline__ T pow_fixed(T x, T y) {
    if constexpr (ORDER == 2) { return x * x; }
    else if constexpr (ORDER == 3) { return x * x * x; }
    else if constexpr (ORDER == 5) { const T x2 = x * x; return x2 * x2 * x; }
    else if constexpr (ORDER == 10) { const T x2 = x * x, x5 = x2 * x2 * x; return x5 * x5; }
    else if constexpr (ORDER == 20) { const T x2 = x * x, x5 = x2 * x2 * x, x10 = x5 * x5; return x10 * x10; }
    else { return Prec<T>::pow_pos(x, y); }
}

// {x, y, z, radius} of one obstacle of one instance: one 16-byte (FP32) or 32-byte (FP64) vector.
template <typename T> struct Vec4;
template <> struct __align__(16) Vec4<float> { float x, y, z, w; };
template <> struct __align__(32) Vec4<double> { double x, y, z, w; };
template <typename T> struct Vec2;
template <> struct __align__(8) Vec2<float> { float x, y; };
template <> struct __align__(16) Vec2<double> { double x, y; };

// ---- symmetric 6x6: packed lower triangle, index (i,j), i >= j -> i*(i+1)/2 + j
__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// Compile-time loops: `#pragma unroll` left one of the nested Cholesky loops rolled, which
// demoted the packed matrix to local memory; template recursion cannot be declined.
template <int I, int END, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < END) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, END>(static_cast<F&&>(f));
    }
}

// ---- rank-1 update, axpy and dot product on 6-vectors.  The float overloads issue on sm_100's packed FP32 instructions
// (fma.rn.f32x2 -> SASS FFMA2, which takes a scalar operand broadcast to both halves): adjacent entries of a row of the
// packed triangle, and adjacent components of a Jacobian column, travel as register pairs -- 12 instructions instead of 21
// for A += c c^T, 3 instead of 6 for y += c x, 4 instead of 6 for c . y.  Each half is the IEEE fma, so values are
// unchanged except for the association of the dot product (pairs first).
template <typename T>
__device__ __forceinline__ void syr6(T (&a)[21], const T (&c)[6]) {
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int s = 0; s <= r; ++s) a[tri(r, s)] = fma(c[r], c[s], a[tri(r, s)]);
}
__device__ __forceinline__ void syr6(float (&a)[21], const float (&c)[6]) {
    static_for<0, 6>([&](auto rc) {
        constexpr int r = decltype(rc)::value;
        static_for<0, (r + 2) / 2>([&](auto hc) {
            constexpr int s = 2 * decltype(hc)::value;
            if constexpr (s + 1 <= r) {
                const float2 v = __ffma2_rn(make_float2(c[r], c[r]), make_float2(c[s], c[s + 1]),
                                            make_float2(a[tri(r, s)], a[tri(r, s + 1)]));
                a[tri(r, s)] = v.x; a[tri(r, s + 1)] = v.y;
            } else {
                a[tri(r, s)] = fmaf(c[r], c[s], a[tri(r, s)]);
            }
        });
    });
}

template <typename T>
__device__ __forceinline__ void axpy6(T (&y)[6], const T (&c)[6], T x) {
#pragma unroll
    for (int r = 0; r < 6; ++r) y[r] = fma(c[r], x, y[r]);
}
__device__ __forceinline__ void axpy6(float (&y)[6], const float (&c)[6], float x) {
#pragma unroll
    for (int r = 0; r < 6; r += 2) {
        const float2 v = __ffma2_rn(make_float2(c[r], c[r + 1]), make_float2(x, x), make_float2(y[r], y[r + 1]));
        y[r] = v.x; y[r + 1] = v.y;
    }
}

// init + sign * (c . y): one fma chain starting from init (both precisions).  dot6x2 below runs two such chains in the halves of
// packed instructions, so the general instantiation (dot6) and the lean one (dot6x2) produce the same bits; dot6_packed is the
// FP32 pairs-first form (5 instructions instead of 6, another association) kept for the long-chain lean shape.
template <typename T>
__device__ __forceinline__ T dot6(const T (&c)[6], const T (&y)[6], T init, bool negate) {
    T acc = init;
#pragma unroll
    for (int r = 0; r < 6; ++r) acc = fma(negate ? -c[r] : c[r], y[r], acc);
    return acc;
}
template <typename T>
__device__ __forceinline__ T dot6_packed(const T (&c)[6], const T (&y)[6], T init, bool negate) { return dot6(c, y, init, negate); }
__device__ __forceinline__ float dot6_packed(const float (&c)[6], const float (&y)[6], float init, bool negate) {
    float2 v = __fmul2_rn(make_float2(c[0], c[1]), make_float2(y[0], y[1]));
    v = __ffma2_rn(make_float2(c[2], c[3]), make_float2(y[2], y[3]), v);
    v = __ffma2_rn(make_float2(c[4], c[5]), make_float2(y[4], y[5]), v);
    const float d = v.x + v.y;
    return negate ? init - d : init + d;
}

// Two dot products of one 6-vector at once: *ra = ia + c . ya, *rb = ib + c . yb.  FP32: the two right-hand sides travel as
// register pairs and every step is one packed fma with c[r] as the broadcast scalar operand -- 6 instructions for both
// products where two dot6 calls take 10.
template <typename T>
__device__ __forceinline__ void dot6x2(const T (&c)[6], const T (&ya)[6], const T (&yb)[6], T ia, T ib, T* ra, T* rb) {
    T a = ia, b = ib;
#pragma unroll
    for (int r = 0; r < 6; ++r) { a = fma(c[r], ya[r], a); b = fma(c[r], yb[r], b); }
    *ra = a; *rb = b;
}
__device__ __forceinline__ void dot6x2(const float (&c)[6], const float (&ya)[6], const float (&yb)[6], float ia, float ib,
                                       float* ra, float* rb) {
    float2 v = make_float2(ia, ib);
#pragma unroll
    for (int r = 0; r < 6; ++r) v = __ffma2_rn(make_float2(c[r], c[r]), make_float2(ya[r], yb[r]), v);
    *ra = v.x; *rb = v.y;
}

// In-place Cholesky A = L L^T of a packed SPD 6x6.  On return a[] holds L's strict
// lower part and inv_d[j] = 1 / L[j][j] (the diagonal is only ever needed inverted).
// `floor` (double only): pivots are not allowed below it -- see the call site; the float instantiation ignores it.
template <typename T>
__device__ __forceinline__ void chol6(T (&a)[21], T (&inv_d)[6], T floor = T(0)) {
    static_for<0, 6>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        T s = a[tri(j, j)];
        static_for<0, j>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            s = fma(-a[tri(j, k)], a[tri(j, k)], s);
        });
        if constexpr (sizeof(T) == 8) s = s > floor ? s : floor;
        const T id = Prec<T>::rsqrt_pos(s);
        inv_d[j] = id;
        static_for<j + 1, 6>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            T t = a[tri(i, j)];
            static_for<0, j>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                t = fma(-a[tri(i, k)], a[tri(j, k)], t);
            });
            a[tri(i, j)] = t * id;
        });
    });
}

// Solve L z = b (forward) in place.
template <typename T>
__device__ __forceinline__ void chol6_fwd(const T (&l)[21], const T (&inv_d)[6], T (&b)[6]) {
    static_for<0, 6>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        T t = b[i];
        static_for<0, i>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            t = fma(-l[tri(i, k)], b[k], t);
        });
        b[i] = t * inv_d[i];
    });
}

// Solve L^T y = z (backward) in place.
template <typename T>
__device__ __forceinline__ void chol6_bwd(const T (&l)[21], const T (&inv_d)[6], T (&b)[6]) {
    static_for<0, 6>([&](auto ic) {
        constexpr int i = 5 - decltype(ic)::value;
        T t = b[i];
        static_for<i + 1, 6>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            t = fma(-l[tri(k, i)], b[k], t);
        });
        b[i] = t * inv_d[i];
    });
}

// Both triangular solves for TWO right-hand sides at once.  FP32: the pair {a[i], b[i]} travels in one register pair and every
// step is one packed instruction with the matrix entry as the broadcast operand (21 instead of 42 per direction); each half
// is the scalar solve's own operation sequence, so the results equal chol6_fwd / chol6_bwd applied to a and b separately.
template <typename T>
__device__ __forceinline__ void chol6_solve2(const T (&l)[21], const T (&inv_d)[6], T (&a)[6], T (&b)[6]) {
    chol6_fwd<T>(l, inv_d, a);
    chol6_bwd<T>(l, inv_d, a);
    chol6_fwd<T>(l, inv_d, b);
    chol6_bwd<T>(l, inv_d, b);
}
__device__ __forceinline__ void chol6_solve2(const float (&l)[21], const float (&inv_d)[6], float (&a)[6], float (&b)[6]) {
    float2 v[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) v[r] = make_float2(a[r], b[r]);
    static_for<0, 6>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        float2 t = v[i];
        static_for<0, i>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            t = __ffma2_rn(make_float2(-l[tri(i, k)], -l[tri(i, k)]), v[k], t);
        });
        v[i] = __fmul2_rn(t, make_float2(inv_d[i], inv_d[i]));
    });
    static_for<0, 6>([&](auto ic) {
        constexpr int i = 5 - decltype(ic)::value;
        float2 t = v[i];
        static_for<i + 1, 6>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            t = __ffma2_rn(make_float2(-l[tri(k, i)], -l[tri(k, i)]), v[k], t);
        });
        v[i] = __fmul2_rn(t, make_float2(inv_d[i], inv_d[i]));
    });
#pragma unroll
    for (int r = 0; r < 6; ++r) { a[r] = v[r].x; b[r] = v[r].y; }
}

// Unit quaternion (w >= 0) of a rotation matrix given row-major m[9]; Shepperd's
// branch selection keeps it well conditioned near angle = pi.  Matches
// oracle/batch.py:rot_axis_angle.
template <typename T>
__device__ __forceinline__ void rot_to_quat(const T (&m)[9], T& w, T& x, T& y, T& z) {
    const T t = m[0] + m[4] + m[8];
    if (t > T(0)) {
        const T s = Prec<T>::sqrt_(T(1) + t) * T(2);
        const T is = Prec<T>::rcp(s);
        w = T(0.25) * s;
        x = (m[7] - m[5]) * is;
        y = (m[2] - m[6]) * is;
        z = (m[3] - m[1]) * is;
    } else if (m[0] >= m[4] && m[0] >= m[8]) {          // argmax picks the first maximum
        const T s = Prec<T>::sqrt_(Prec<T>::fmax_(T(1) + T(2) * m[0] - t, T(0))) * T(2);
        const T is = Prec<T>::rcp(s);
        x = T(0.25) * s;
        y = (m[3] + m[1]) * is;
        z = (m[6] + m[2]) * is;
        w = (m[7] - m[5]) * is;
    } else if (m[4] >= m[8]) {
        const T s = Prec<T>::sqrt_(Prec<T>::fmax_(T(1) + T(2) * m[4] - t, T(0))) * T(2);
        const T is = Prec<T>::rcp(s);
        y = T(0.25) * s;
        z = (m[7] + m[5]) * is;
        x = (m[1] + m[3]) * is;
        w = (m[2] - m[6]) * is;
    } else {
        const T s = Prec<T>::sqrt_(Prec<T>::fmax_(T(1) + T(2) * m[8] - t, T(0))) * T(2);
        const T is = Prec<T>::rcp(s);
        z = T(0.25) * s;
        x = (m[2] + m[6]) * is;
        y = (m[5] + m[7]) * is;
        w = (m[3] - m[1]) * is;
    }
    if (w < T(0)) { w = -w; x = -x; y = -y; z = -z; }
}

}  // namespace vfk
