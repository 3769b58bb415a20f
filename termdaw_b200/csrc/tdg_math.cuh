// tdg_math.cuh — scalar maths of the voice / envelope kernels, host + device.
//
// Every function keeps the reference's f32 operation ORDER (one IEEE rounding per operation); the
// library is compiled with -fmad=false (device) and -ffp-contract=off (host) so that no product is
// contracted into an FMA.  The host planner uses the same functions for the handful of values the
// reference computes at event time (e.g. the Adsr vertex's stored release value).
#pragma once
#include <math.h>
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


// ---- device-only fast forms --------------------------------------------------------------------
#ifdef __CUDACC__
// t / d for a per-vertex constant d with rcp = 1/d precomputed on the host: one Newton step on the
// exact FMA residual gives the correctly rounded quotient (Markstein) without MUFU.RCP; anything
// unusual (d == 0, inf/nan, zero or extreme quotients) takes the IEEE division.
__device__ __forceinline__ float d_div(float t, float d, float rcp) {
    const float q = t * rcp;
    const float e = fmaf(-q, d, t);
    const float q1 = fmaf(e, rcp, q);
    const float aq = fabsf(q);
    return (aq > 1e-30f && aq < 1e30f && fabsf(rcp) < 1e30f) ? q1 : t / d;
}
__device__ __forceinline__ float d_ads_internal(const AdsrConfD& c, const AdsrRcp& r, float t) {   // adsr.rs:46-60
    if (t <= c.attack_sec) return m_lerp(c.std_vel, c.attack_vel, d_div(t, c.attack_sec, r.attack));
    if (t <= c.attack_sec + c.decay_sec)
        return m_lerp(c.attack_vel, c.decay_vel, d_div(t - c.attack_sec, c.decay_sec, r.decay));
    if (t <= c.attack_sec + c.decay_sec + c.sustain_sec)
        return m_lerp(c.decay_vel, c.sustain_vel, d_div(t - c.attack_sec - c.decay_sec, c.sustain_sec, r.sustain));
    return -1000.0f;
}
__device__ __forceinline__ float d_apply_ads(const AdsrConfD& c, const AdsrRcp& r, float t) {
    const float res = d_ads_internal(c, r, t);
    return res <= -1.0f ? c.sustain_vel : res;
}
__device__ __forceinline__ float d_apply_r(const AdsrConfD& c, const AdsrRcp& r, float t, float old_val) {
    return m_lerp(old_val, c.release_vel, fminf(d_div(t, c.release_sec, r.release), 1.0f));
}
__device__ __forceinline__ float d_apply_adsr(const AdsrConfD& c, const AdsrRcp& r, float t) {
    const float res = d_ads_internal(c, r, t);
    if (res <= -1.0f)
        return m_lerp(c.sustain_vel, c.release_vel,
                      fminf(d_div(t - c.attack_sec - c.decay_sec - c.sustain_sec, c.release_sec, r.release), 1.0f));
    return res;
}
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
