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

}  // namespace tdg
