// tdg_math.cuh — scalar maths of the voice / envelope kernels, host + device.
//
// Every function keeps the reference's f32 operation ORDER (one IEEE rounding per operation); the
// library is compiled with -fmad=false (device) and -ffp-contract=off (host) so that no product is
// contracted into an FMA.  The host planner uses the same functions for the handful of values the
// reference computes at event time (e.g. the Adsr vertex's stored release value).
#pragma once
#include <math.h>
#include <string.h>
#include "tdg_types.h"

#ifdef __CUDACC__
#define TDG_HD __host__ __device__ __forceinline__
#else
#define TDG_HD inline
#endif

namespace tdg {

#define TDG_PI_F 3.14159265358979323846f

TDG_HD float m_lerp(float a, float b, float t) { return a + t * (b - a); }                 // adsr.rs:42

TDG_HD float m_ads_internal(const AdsrConfD& c, float t) {                                  // adsr.rs:46-60
    if (t <= c.attack_sec) return m_lerp(c.std_vel, c.attack_vel, t / c.attack_sec);
    if (t <= c.attack_sec + c.decay_sec)
        return m_lerp(c.attack_vel, c.decay_vel, (t - c.attack_sec) / c.decay_sec);
    if (t <= c.attack_sec + c.decay_sec + c.sustain_sec)
        return m_lerp(c.decay_vel, c.sustain_vel, (t - c.attack_sec - c.decay_sec) / c.sustain_sec);
    return -1000.0f;
}
TDG_HD float m_apply_ads(const AdsrConfD& c, float t) {                                     // adsr.rs:62-69
    float res = m_ads_internal(c, t);
    return res <= -1.0f ? c.sustain_vel : res;
}
TDG_HD float m_apply_r(const AdsrConfD& c, float t, float old_val) {                        // adsr.rs:71-73
    return m_lerp(old_val, c.release_vel, fminf(t / c.release_sec, 1.0f));
}
TDG_HD float m_apply_adsr(const AdsrConfD& c, float t) {                                    // adsr.rs:75-86
    float res = m_ads_internal(c, t);
    if (res <= -1.0f)
        return m_lerp(c.sustain_vel, c.release_vel,
                      fminf((t - c.attack_sec - c.decay_sec - c.sustain_sec) / c.release_sec, 1.0f));
    return res;
}
TDG_HD float m_apply_r_rt(const AdsrConfD& c, float t, float rt) {                          // adsr.rs:89-92
    return m_apply_r(c, t, m_apply_ads(c, rt));
}
TDG_HD float m_max_vel(const AdsrConfD& c) {                                                // adsr.rs:32-38
    return fmaxf(fmaxf(fmaxf(fmaxf(c.std_vel, c.attack_vel), c.decay_vel), c.sustain_vel), c.release_vel);
}

// synth.rs:22-34 with sin(t*hz*2*PI) hoisted (square and topflat evaluate the same expression)
TDG_HD float m_osc_arg(float t, float hz) { return t * hz * 2.0f * TDG_PI_F; }
TDG_HD float m_square_from_sin(float s, float z) { return fminf(fmaxf(s, -z), z) * (1.0f / z); }
TDG_HD float m_topflat_from_sin(float s, float z) {
    return (fminf(s, z) + ((1.0f - z) / 2.0f)) * (2.0f / (1.0f + z));
}
TDG_HD float m_triangle(float t, float hz) {
    return 4.0f * fabsf((t * hz) - floorf((t * hz) + 0.5f)) - 1.0f;
}


// ---- sinf: glibc's algorithm, operation for operation ------------------------------------------------------
// The reference's oscillators call Rust's f32::sin, i.e. the platform libm's sinf (extensions.rs:452, synth.rs:23,28).
// glibc >= 2.28 computes sinf in double precision (sysdeps/ieee754/flt-32/s_sinf.c + sincosf.h, from ARM's
// optimized routines): three argument ranges (|y| < pi/4, < 120, else), the reduction of large arguments by a
// 96-bit window of the bits of 2/pi in integer arithmetic, one of two degree-7/8 polynomials in double, one
// rounding to float at the end.  Restated here with the contraction pattern of the x86-64 FMA build (the ifunc
// variant every FMA-capable host selects): tests/test_sinf_restatement.py checks this very function against the
// host's sinf over ALL 2^32 arguments (0 mismatches on glibc 2.39), so on the device the oscillators are bit-exact
// against the oracle instead of "within 2 ulp".  Constants: __sincosf_table / __inv_pio4 of glibc 2.39.
#define TDG_INV_PIO4_WORDS {0xa2u, 0xa2f9u, 0xa2f983u, 0xa2f9836eu, 0xf9836e4eu, 0x836e4e44u, 0x6e4e4415u, 0x4e441529u, \
                            0x441529fcu, 0x1529fc27u, 0x29fc2757u, 0xfc2757d1u, 0x2757d1f5u, 0x57d1f534u, 0xd1f534ddu, 0xf534ddc0u, \
                            0x34ddc0dbu, 0xddc0db62u, 0xc0db6295u, 0xdb629599u, 0x6295993cu, 0x95993c43u, 0x993c4390u, 0x3c439041u}
static const uint32_t kInvPio4Host[24] = TDG_INV_PIO4_WORDS;
#ifdef __NVCC__
static __constant__ uint32_t kInvPio4Dev[24] = TDG_INV_PIO4_WORDS;     // device copy: constant bank, near-uniform index
#endif
#ifdef __CUDA_ARCH__
#define TDG_INV_PIO4 kInvPio4Dev
#else
#define TDG_INV_PIO4 kInvPio4Host
#endif
// the polynomial / reduction constants of __sincosf_table: the three coefficient pairs are stored (sine, cosine) side
// by side so that the quadrant parity indexes them (one LDC.64 with a register offset each on the device)
#define TDG_SINF_CONSTS {-0x1.555545995a603p-3, 0x1.55553e1068f19p-5,  /* S1 | C2 */ \
                         0x1.1107605230bc4p-7, -0x1.6c087e89a359dp-10, /* S2 | C3 */ \
                         -0x1.994eb3774cf24p-13, 0x1.99343027bf8c3p-16, /* S3 | C4 */ \
                         -0x1.ffffffd0c621cp-2,                         /* C1 */ \
                         0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1.921FB54442D18p-62}
static const double kSinfConstHost[10] = TDG_SINF_CONSTS;
#ifdef __NVCC__
static __constant__ double kSinfConstDev[10] = TDG_SINF_CONSTS;
#endif
#ifdef __CUDA_ARCH__
#define TDG_SC(i) kSinfConstDev[i]
#else
#define TDG_SC(i) kSinfConstHost[i]
#endif
enum { SC_K1 = 0, SC_K2 = 2, SC_K3 = 4, SC_C1 = 6, SC_HPI_INV = 7, SC_HPI = 8, SC_PI63 = 9 };
TDG_HD uint32_t m_f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
TDG_HD float m_bits_f32(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
// sinf_poly (sincosf.h): n even -> sine polynomial of x, n odd -> cosine polynomial of x2.  The two have the same
// shape, so ONE chain evaluates whichever the quadrant asks for, with operands selected by the parity:
//     sine:    x3 = x*x2;   x7 = x3*x2;  s1 = S2 + x2*S3;  s = x  + x3*S1;  return s + x7*s1
//     cosine:  x4 = x2*x2;  x6 = x4*x2;  c2 = C3 + x2*C4;  c = c1 + x4*C2;  return c + x6*c2     (c1 = 1 + x2*C1)
// i.e. p = u*x2 with u = x | x2, q = p*x2, a = K2 + x2*K3, r = base + p*K1 with base = x | c1, result r + q*a: seven
// double operations instead of the twelve of both polynomials side by side (neighbouring frames of a warp sit in
// different quadrants, so a branch would execute both sides anyway); the operations of the side that is kept are
// exactly the reference's.  `negmask` (0 or 0x80000000) folds in the quadrant sign: the reference multiplies x by
// +-1.0 and picks the table with negated cosine coefficients, and both polynomials are exactly odd in that sign under
// round-to-nearest, so flipping the sign bit of the result is the same value.
TDG_HD float m_sinf_poly(double x, uint32_t oddbit, uint32_t odd, uint32_t negmask) {   // oddbit != 0 <=> odd == 1
    const double x2 = x * x;
    const double u = oddbit ? x2 : x;
    const double p = u * x2;
    const double q = p * x2;
    const double a = fma(x2, TDG_SC(SC_K3 + odd), TDG_SC(SC_K2 + odd));
    const double c1 = fma(x2, TDG_SC(SC_C1), 1.0);
    const double base = oddbit ? c1 : x;
    const double r = fma(p, TDG_SC(SC_K1 + odd), base);
    return m_bits_f32(m_f32_bits((float)fma(q, a, r)) ^ negmask);
}
// Quadrant bookkeeping: the reference's sign table {1,-1,-1,1}[(n + sign) & 3] together with its table switch on
// (n + sign) & 2 reduces to "negate iff bit 1 of n differs from the sign of y" for both parities of n (n even:
// sin(x * s) with s from the sign table; n odd: +-cos(x) by the table switch; (n + sign) carries into bit 1 exactly
// when n is odd and y negative).  reduce_fast works on the signed x, so there it is bit 1 of n alone.
TDG_HD float m_sinf_glibc(float y) {
    const uint32_t bits = m_f32_bits(y);
    const uint32_t top = bits & 0x7ff00000u;                         // abstop12, in place
    if (top >= 0x42f00000u) {                                        // |y| >= 120 (every late-timeline argument)
        if (top >= 0x7f800000u) return y - y;                        // inf / NaN -> NaN (__math_invalidf)
        // reduce_large: y's 24-bit significand times a 96-bit window of the bits of 4/pi, kept to 64 bits
        const uint32_t* arr = &TDG_INV_PIO4[(bits >> 26) & 15];
        const uint32_t xi = ((bits & 0x7fffffu) | 0x800000u) << ((bits >> 23) & 7);
        const uint32_t p0 = xi * arr[0];
        const uint64_t r2 = (uint64_t)xi * arr[8];
        const uint64_t res = (uint64_t)xi * arr[4] + (((uint64_t)p0 << 32) | (r2 >> 32));
        uint32_t hi = (uint32_t)(res >> 32);
        const uint32_t hq = hi + 0x20000000u;                        // n = hq >> 30 (quadrant, rounded to nearest)
        hi -= hq & 0xc0000000u;                                      // res -= n << 62
        const double x = (double)(int64_t)(((uint64_t)hi << 32) | (uint32_t)res) * TDG_SC(SC_PI63);
        return m_sinf_poly(x, hq & 0x40000000u, (hq >> 30) & 1u, (hq ^ bits) & 0x80000000u);
    }
    double x = (double)y;
    if (top < 0x3f400000u) {                                         // |y| < pi/4 (abstop12(0x1.921FB6p-1f))
        if (top < 0x39800000u) return y;                             // |y| < 2^-12
        return m_sinf_poly(x, 0u, 0u, 0u);
    }
    const double r = x * TDG_SC(SC_HPI_INV);                         // |y| < 120: reduce_fast
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = fma(-(double)n, TDG_SC(SC_HPI), x);
    return m_sinf_poly(x, (uint32_t)n & 1u, (uint32_t)n & 1u, ((uint32_t)n << 30) & 0x80000000u);
}

// ---- AdsrK: host builder (plain C++, same f32 operations as the reference's expressions) ---------
inline bool adsr_duration_exotic(float d) {
    if (d == 0.0f) return false;
    return !(d >= 1e-6f && d <= 1e6f);          // negative, NaN, inf, denormal-ish or huge
}
inline AdsrK make_adsr_k(const AdsrConfD& c) {
    AdsrK k;
    k.a = c.attack_sec; k.ad = c.attack_sec + c.decay_sec; k.ads = c.attack_sec + c.decay_sec + c.sustain_sec;
    k.dsec = c.decay_sec; k.ssec = c.sustain_sec; k.rsec = c.release_sec;
    k.ra = 1.0f / c.attack_sec; k.rd = 1.0f / c.decay_sec; k.rs = 1.0f / c.sustain_sec; k.rr = 1.0f / c.release_sec;
    k.v0 = c.std_vel; k.v1 = c.attack_vel; k.v2 = c.decay_vel; k.v3 = c.sustain_vel; k.v4 = c.release_vel;
    k.d01 = c.attack_vel - c.std_vel; k.d12 = c.decay_vel - c.attack_vel; k.d23 = c.sustain_vel - c.decay_vel;
    k.d34 = c.release_vel - c.sustain_vel;
    k.exotic = (adsr_duration_exotic(c.attack_sec) || adsr_duration_exotic(c.decay_sec) ||
                adsr_duration_exotic(c.sustain_sec) || adsr_duration_exotic(c.release_sec)) ? 1u : 0u;
    const float lo = fminf(fminf(fminf(fminf(c.std_vel, c.attack_vel), c.decay_vel), c.sustain_vel), c.release_vel);
    k.simple = (!k.exotic && c.attack_sec != 0.0f && c.decay_sec != 0.0f && c.sustain_sec != 0.0f && c.release_sec != 0.0f &&
                lo > -0.5f) ? 1u : 0u;
    return k;
}

// ---- device-only fast forms --------------------------------------------------------------------
#ifdef __CUDACC__
// t / d for a per-vertex constant d with r = RN(1/d) from the host: one Newton step on the exact FMA
// residual gives the correctly rounded quotient (Markstein; checked against IEEE division on 1.2e9
// operand pairs, DESIGN.md "arithmetic") without MUFU.RCP or the division slow path.  d == 0 makes
// r infinite and q = t*inf is already t/0 (inf, -inf or NaN), which the |q| < inf test passes on.
// EXACT = plain IEEE division, used for launches holding an exotic duration (AdsrK::exotic).
template <bool EXACT>
__device__ __forceinline__ float k_div(float t, float d, float r) {
    if (EXACT) return t / d;
    const float q = t * r;
    const float q1 = fmaf(fmaf(-q, d, t), r, q);
    return fabsf(q) < __int_as_float(0x7f800000) ? q1 : q;
}
// the same quotient without the overflow / zero-divisor test: for AdsrK::simple envelopes (d in [1e-6, 1e6],
// finite env time) t*r is finite, so the test can never fire
__device__ __forceinline__ float k_div_fast(float t, float d, float r) {
    const float q = t * r;
    return fmaf(fmaf(-q, d, t), r, q);
}
template <bool EXACT>
__device__ __forceinline__ float k_ads_internal(const AdsrK& c, float t) {                   // adsr.rs:46-60
    if (t <= c.a) return c.v0 + k_div<EXACT>(t, c.a, c.ra) * c.d01;
    if (t <= c.ad) return c.v1 + k_div<EXACT>(t - c.a, c.dsec, c.rd) * c.d12;
    if (t <= c.ads) return c.v2 + k_div<EXACT>(t - c.a - c.dsec, c.ssec, c.rs) * c.d23;
    return -1000.0f;
}
template <bool EXACT>
__device__ __forceinline__ float k_apply_ads(const AdsrK& c, float t) {                      // adsr.rs:62-69
    const float res = k_ads_internal<EXACT>(c, t);
    return res <= -1.0f ? c.v3 : res;
}
template <bool EXACT>
__device__ __forceinline__ float k_apply_r(const AdsrK& c, float t, float old_val) {         // adsr.rs:71-73
    return old_val + fminf(k_div<EXACT>(t, c.rsec, c.rr), 1.0f) * (c.v4 - old_val);
}
// The same two functions over the F consecutive frames of one thread.  env_time is non-decreasing over
// those frames and (for non-exotic envelopes: durations >= 0) so are the thresholds, hence when the
// first and the last frame sit in the same segment of the envelope all F do: the segment's constants
// are selected once and the F quotients/lerps are straight-line code that the scheduler can overlap.
// Return value: true when all F results are one and the same value (sustain plateau, finished release):
// callers then do the per-frame work that depends on the envelope once instead of F times.
template <bool EXACT, int F>
__device__ __forceinline__ bool k_apply_ads_n(const AdsrK& c, const float (&t)[F], float (&e)[F]) {
    if (!EXACT) {
        const float tf = t[0], tl = t[F - 1];
        const int s0 = (tf <= c.a) ? 0 : (tf <= c.ad) ? 1 : (tf <= c.ads) ? 2 : 3;
        const int s1 = (tl <= c.a) ? 0 : (tl <= c.ad) ? 1 : (tl <= c.ads) ? 2 : 3;
        if (s0 == s1) {
            if (s0 == 3) {
#pragma unroll
                for (int j = 0; j < F; ++j) e[j] = c.v3;
                return true;
            }
            const float sub1 = s0 >= 1 ? c.a : 0.0f, sub2 = s0 == 2 ? c.dsec : 0.0f;    // x - 0.0f == x
            const float d = s0 == 0 ? c.a : s0 == 1 ? c.dsec : c.ssec;
            const float r = s0 == 0 ? c.ra : s0 == 1 ? c.rd : c.rs;
            const float base = s0 == 0 ? c.v0 : s0 == 1 ? c.v1 : c.v2;
            const float dl = s0 == 0 ? c.d01 : s0 == 1 ? c.d12 : c.d23;
            if (c.simple) {
#pragma unroll
                for (int j = 0; j < F; ++j) e[j] = base + k_div_fast(t[j] - sub1 - sub2, d, r) * dl;
                return false;
            }
#pragma unroll
            for (int j = 0; j < F; ++j) {
                const float res = base + k_div<false>(t[j] - sub1 - sub2, d, r) * dl;
                e[j] = res <= -1.0f ? c.v3 : res;
            }
            return false;
        }
    }
#pragma unroll
    for (int j = 0; j < F; ++j) e[j] = k_apply_ads<EXACT>(c, t[j]);
    return false;
}
// t is non-decreasing over the F frames and (non-exotic) release_sec >= 0, so the quotient is
// non-decreasing too: once the first frame's quotient has reached 1 the min() clamps all of them and the
// result is old + 1.0*(v4 - old) for every frame (1.0f * x is x exactly).
template <bool EXACT, int F>
__device__ __forceinline__ bool k_apply_r_n(const AdsrK& c, const float (&t)[F], float old_val, float (&e)[F]) {
    const float dl = c.v4 - old_val;
    if (!EXACT) {
        // release towards the value it starts from (e.g. a percussive envelope that has already decayed to
        // its end value): old + q*0 is old for every finite q (old == -0.0 excluded: -0 + +0 is +0)
        if (dl == 0.0f && c.rsec > 0.0f && fabsf(t[F - 1]) < __int_as_float(0x7f800000) && __float_as_uint(old_val) != 0x80000000u) {
#pragma unroll
            for (int j = 0; j < F; ++j) e[j] = old_val;
            return true;
        }
        const float q0 = k_div<false>(t[0], c.rsec, c.rr);
        if (q0 >= 1.0f) {
            const float v = old_val + dl;
#pragma unroll
            for (int j = 0; j < F; ++j) e[j] = v;
            return true;
        }
        e[0] = old_val + fminf(q0, 1.0f) * dl;
        if (c.simple) {
#pragma unroll
            for (int j = 1; j < F; ++j) e[j] = old_val + fminf(k_div_fast(t[j], c.rsec, c.rr), 1.0f) * dl;
            return false;
        }
#pragma unroll
        for (int j = 1; j < F; ++j) e[j] = old_val + fminf(k_div<false>(t[j], c.rsec, c.rr), 1.0f) * dl;
        return false;
    }
#pragma unroll
    for (int j = 0; j < F; ++j) e[j] = old_val + fminf(k_div<EXACT>(t[j], c.rsec, c.rr), 1.0f) * dl;
    return false;
}
// Is the envelope one constant from env time t0 on (t only grows inside a run of frames)?  True on the
// sustain plateau of apply_ads and once a release (apply_r / the release leg of apply_adsr) has reached
// its end value; `val` is then exactly what the per-frame expressions produce there (see k_apply_r_n).
// mode: 0 apply_adsr, 1 apply_ads, 2 apply_r(old_val).  Never true for exotic envelopes (EXACT).
template <bool EXACT>
__device__ __forceinline__ bool k_env_plateau(const AdsrK& c, int mode, float old_val, float t0, float* val) {
    if (EXACT) return false;
    if (mode == 1) { *val = c.v3; return !(t0 <= c.ads) && t0 == t0; }
    if (mode == 2) { *val = old_val + (c.v4 - old_val); return k_div<false>(t0, c.rsec, c.rr) >= 1.0f; }
    *val = c.v3 + c.d34;
    return !(t0 <= c.ads) && t0 == t0 && k_div<false>(t0 - c.a - c.dsec - c.ssec, c.rsec, c.rr) >= 1.0f;
}
template <bool EXACT>
__device__ __forceinline__ float k_apply_adsr(const AdsrK& c, float t) {                     // adsr.rs:75-86
    const float res = k_ads_internal<EXACT>(c, t);
    if (res <= -1.0f)
        return c.v3 + fminf(k_div<EXACT>(t - c.a - c.dsec - c.ssec, c.rsec, c.rr), 1.0f) * c.d34;
    return res;
}
template <bool EXACT, int F>
__device__ __forceinline__ bool k_apply_adsr_n(const AdsrK& c, const float (&t)[F], float (&e)[F]) {
    if (!EXACT) {
        const float tf = t[0], tl = t[F - 1];
        const int s0 = (tf <= c.a) ? 0 : (tf <= c.ad) ? 1 : (tf <= c.ads) ? 2 : 3;
        const int s1 = (tl <= c.a) ? 0 : (tl <= c.ad) ? 1 : (tl <= c.ads) ? 2 : 3;
        if (s0 == s1) {
            if (s0 == 3) {
                // same monotonicity argument as k_apply_r_n: a finished release is one constant
                const float q0 = k_div<false>(t[0] - c.a - c.dsec - c.ssec, c.rsec, c.rr);
                if (q0 >= 1.0f) {
                    const float v = c.v3 + c.d34;
#pragma unroll
                    for (int j = 0; j < F; ++j) e[j] = v;
                    return true;
                }
                e[0] = c.v3 + fminf(q0, 1.0f) * c.d34;
#pragma unroll
                for (int j = 1; j < F; ++j)
                    e[j] = c.v3 + fminf(k_div<false>(t[j] - c.a - c.dsec - c.ssec, c.rsec, c.rr), 1.0f) * c.d34;
                return false;
            }
            const float sub1 = s0 >= 1 ? c.a : 0.0f, sub2 = s0 == 2 ? c.dsec : 0.0f;
            const float d = s0 == 0 ? c.a : s0 == 1 ? c.dsec : c.ssec;
            const float r = s0 == 0 ? c.ra : s0 == 1 ? c.rd : c.rs;
            const float base = s0 == 0 ? c.v0 : s0 == 1 ? c.v1 : c.v2;
            const float dl = s0 == 0 ? c.d01 : s0 == 1 ? c.d12 : c.d23;
            if (c.simple) {
#pragma unroll
                for (int j = 0; j < F; ++j) e[j] = base + k_div_fast(t[j] - sub1 - sub2, d, r) * dl;
                return false;
            }
#pragma unroll
            for (int j = 0; j < F; ++j) {
                const float res = base + k_div<false>(t[j] - sub1 - sub2, d, r) * dl;
                e[j] = res;
                if (res <= -1.0f)                        // adsr.rs:77: falls through to the release lerp
                    e[j] = c.v3 + fminf(k_div<false>(t[j] - c.a - c.dsec - c.ssec, c.rsec, c.rr), 1.0f) * c.d34;
            }
            return false;
        }
    }
#pragma unroll
    for (int j = 0; j < F; ++j) e[j] = k_apply_adsr<EXACT>(c, t[j]);
    return false;
}
#endif

#ifdef __CUDACC__
// floorf without the conversion pipe: exact for every finite input
__device__ __forceinline__ float d_floor(float y) {
    if (fabsf(y) < 4194304.0f) {
        float r = (y + 12582912.0f) - 12582912.0f;      // round to nearest integer (|y| < 2^22)
        return r > y ? r - 1.0f : r;
    }
    return floorf(y);
}
__device__ __forceinline__ float d_triangle(float t, float hz) {                             // synth.rs:32-34
    const float y = t * hz;
    return 4.0f * fabsf(y - d_floor(y + 0.5f)) - 1.0f;
}
// sin(x) of an f32 argument to ~7e-8 absolute (|x| < 2e6: 3-term Cody-Waite in FMAs, magic-number
// rounding, the usual degree 7/8 minimax pair) — the reference calls libm sinf on the same f32
// argument, so only the evaluation differs, not the (chaotic) argument.  Larger arguments reduce in
// f64; nothing here touches the conversion/transcendental pipe on the common path.
__device__ __forceinline__ float d_sin_poly(float r, int q) {
    const float s = r * r;
    float v;
    if (q & 1) {
        v = fmaf(s, 2.44331571e-5f, -1.38873163e-3f); v = fmaf(v, s, 4.16666457e-2f);
        v = fmaf(v, s, -0.5f); v = fmaf(v, s, 1.0f);
    } else {
        v = fmaf(s, -1.95152959e-4f, 8.33216087e-3f); v = fmaf(v, s, -1.66666546e-1f);
        v = v * s; v = fmaf(v, r, r);
    }
    return (q & 2) ? -v : v;
}
static __device__ __noinline__ float d_sin_big(float x) {           // |x| >= 2e6: reduce in f64 (rare on short renders)
    if (fabsf(x) < 1.0e15f) {
        const double xd = (double)x;
        const double kd = rint(xd * 0.63661977236758134308);
        double rd = fma(-kd, 1.57079632679489655800, xd);
        rd = fma(-kd, 6.12323399573676603587e-17, rd);
        return d_sin_poly((float)rd, (int)((long long)kd & 3));
    }
    return sinf(x);
}
__device__ __forceinline__ float d_sin(float x) {
    if (fabsf(x) < 2.0e6f) {
        const float kf = fmaf(x, 0.636619772367581343f, 12582912.0f);
        const int q = __float_as_int(kf);
        const float k = kf - 12582912.0f;
        float r = fmaf(k, -1.57079601e+00f, x);
        r = fmaf(k, -3.13916473e-07f, r);
        r = fmaf(k, -5.39030253e-15f, r);
        return d_sin_poly(r, q);
    }
    return d_sin_big(x);
}
#endif

}  // namespace tdg
