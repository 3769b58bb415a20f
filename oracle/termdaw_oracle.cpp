// termdaw_oracle.cpp — CPU restatement of termdaw's offline render path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under termdaw_b200/ may include, link or call this file; it is
// used by tests/, by __graft_entry__.smoke() as the checker and by bench.py's cpu_baseline /
// `--impl reference` leg.  It follows the reference chunk for chunk, single threaded, f32 with one
// rounding per operation (build: g++ -O2 -ffp-contract=off, no fast-math), quirks included
// (SURVEY.md Appendix A).  File:line citations are relative to the reference repository.
//
// PARITY PINNING
//   * adsr.rs:121-204 (5 known-answer tests) are the only golden vectors the reference has for
//     this path; tests/test_oracle_adsr_kat.py checks this file against every one of them.
//   * Everything else is pinned by reading the source only (no Rust toolchain here, so the
//     reference cannot be run): PARITY UNPINNED beyond the ADSR KATs.
//   * rubato 0.15 (SincFixedIn), sampsyn 0.1 and hound 3.5 are not vendored: the resampler follows
//     the recalled published algorithm, the wavetable synth is a semantic
//     DEFINED here (Appendix B.4), PCM packing follows the WAV standard (Appendix B.2).
//     These three are "parity unpinned vs upstream".
//
// The exported C API mirrors include/termdaw_b200.h with the prefix tdo_.

#include <cmath>
#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <vector>
#include <algorithm>
#include <memory>

namespace {

constexpr float PI_F = 3.14159265358979323846f;        // core::f32::consts::PI
constexpr float E_F = 2.71828182845904523536f;         // std::f32::consts::E
constexpr float FRAC_1_SQRT_2_F = 0.70710678118654752440f;

enum { OK = 0, ERR_INVALID = -1, ERR_UNSUPPORTED_VERTEX = -3, ERR_ADSR_CONF = -4, ERR_SAMPLE = -5,
       ERR_BITDEPTH = -6, ERR_NO_OUTPUT = -7, ERR_TOO_WIDE = -8, ERR_CAPACITY = -9 };

// Rust `f32 as usize`: saturating, NaN -> 0.
inline uint64_t f32_as_usize(float x) {
    if (!(x == x)) return 0;
    if (x <= 0.0f) return 0;
    if (x >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)x;
}

// ---------------------------------------------------------------------------------- adsr.rs
struct AdsrConf {                       // adsr.rs:2-12
    float std_vel = 0, attack_sec = 0, attack_vel = 0, decay_sec = 0, decay_vel = 0,
          sustain_sec = 0, sustain_vel = 0, release_sec = 0, release_vel = 0;
};
AdsrConf hit_conf(float a_s, float d_s, float d_v, float s_s, float s_v, float r_s) {  // adsr.rs:15-30
    AdsrConf c; c.std_vel = 0.0f; c.attack_sec = a_s; c.attack_vel = 1.0f; c.decay_sec = d_s;
    c.decay_vel = d_v; c.sustain_sec = s_s; c.sustain_vel = s_v; c.release_sec = r_s;
    c.release_vel = 0.0f; return c;
}
float max_vel(const AdsrConf& c) {      // adsr.rs:32-38
    return fmaxf(fmaxf(fmaxf(fmaxf(c.std_vel, c.attack_vel), c.decay_vel), c.sustain_vel), c.release_vel);
}
inline float lerp(float a, float b, float t) { return a + t * (b - a); }   // adsr.rs:42-44
float apply_ads_internal(const AdsrConf& c, float t) {                      // adsr.rs:46-60
    if (t <= c.attack_sec) return lerp(c.std_vel, c.attack_vel, t / c.attack_sec);
    else if (t <= c.attack_sec + c.decay_sec)
        return lerp(c.attack_vel, c.decay_vel, (t - c.attack_sec) / c.decay_sec);
    else if (t <= c.attack_sec + c.decay_sec + c.sustain_sec)
        return lerp(c.decay_vel, c.sustain_vel, (t - c.attack_sec - c.decay_sec) / c.sustain_sec);
    else return -1000.0f;
}
float apply_ads(const AdsrConf& c, float t) {                               // adsr.rs:62-69
    float res = apply_ads_internal(c, t);
    return res <= -1.0f ? c.sustain_vel : res;
}
float apply_r(const AdsrConf& c, float t, float old_val) {                  // adsr.rs:71-73
    return lerp(old_val, c.release_vel, fminf(t / c.release_sec, 1.0f));
}
float apply_adsr(const AdsrConf& c, float t) {                              // adsr.rs:75-86
    float res = apply_ads_internal(c, t);
    if (res <= -1.0f)
        return lerp(c.sustain_vel, c.release_vel,
                    fminf((t - c.attack_sec - c.decay_sec - c.sustain_sec) / c.release_sec, 1.0f));
    return res;
}
float apply_r_rt(const AdsrConf& c, float t, float rt) {                    // adsr.rs:89-92
    return apply_r(c, t, apply_ads(c, rt));
}
bool build_adsr_conf(const float* arr, int n, AdsrConf* out) {              // adsr.rs:94-114
    if (n == 0) { *out = AdsrConf(); return true; }
    if (n == 6) { *out = hit_conf(arr[0], arr[1], arr[2], arr[3], arr[4], arr[5]); return true; }
    if (n == 9) {
        AdsrConf c; c.std_vel = arr[0]; c.attack_sec = arr[1]; c.attack_vel = arr[2]; c.decay_sec = arr[3];
        c.decay_vel = arr[4]; c.sustain_sec = arr[5]; c.sustain_vel = arr[6]; c.release_sec = arr[7];
        c.release_vel = arr[8]; *out = c; return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------- synth.rs
struct OscConf { float volume, param; AdsrConf adsr; };                     // synth.rs:5-9
inline float square_sine_sample(float t, float hz, float z) {               // synth.rs:22-24
    return fminf(fmaxf(sinf(t * hz * 2.0f * PI_F), -z), z) * (1.0f / z);
}
inline float topflat_sine_sample(float t, float hz, float z) {              // synth.rs:27-29
    return (fminf(sinf(t * hz * 2.0f * PI_F), z) + ((1.0f - z) / 2.0f)) * (2.0f / (1.0f + z));
}
inline float triangle_sample(float t, float hz) {                           // synth.rs:32-34
    return 4.0f * fabsf((t * hz) - floorf((t * hz) + 0.5f)) - 1.0f;
}

// ---------------------------------------------------------------------------------- sample.rs
float absmaxlen(const std::vector<float>& s, size_t len) {                  // sample.rs:12-14
    float max = 0.0f;
    size_t n = std::min(len, s.size());
    for (size_t i = 0; i < n; ++i) { float a = fabsf(s[i]); if (a > max) max = a; }
    return max;
}
float absmax(const std::vector<float>& s) { return absmaxlen(s, SIZE_MAX); }
float mean_energy(const std::vector<float>& s) {                            // sample.rs:16-22
    if (s.empty()) return 0.0f;
    float sum = 0.0f;
    for (float v : s) sum += fabsf(v);
    return sum / (float)s.size();
}

struct Resampler;   // rubato restatement, below

struct Sample {                                                             // sample.rs:25-28
    std::vector<float> l, r;
    explicit Sample(size_t bl = 0) : l(bl, 0.0f), r(bl, 0.0f) {}
    size_t len() const { return l.size(); }
    void zero() { std::fill(l.begin(), l.end(), 0.0f); std::fill(r.begin(), r.end(), 0.0f); }  // :87-95
    void apply_angle(float angle, size_t len) {                             // sample.rs:97-106
        if (fabsf(angle) < 0.001f) return;
        float angle_rad = angle * 0.5f * 0.01745329f;
        float l_amp = FRAC_1_SQRT_2_F * (cosf(angle_rad) + sinf(angle_rad));
        float r_amp = FRAC_1_SQRT_2_F * (cosf(angle_rad) - sinf(angle_rad));
        for (size_t i = 0; i < len; ++i) { l[i] *= l_amp; r[i] *= r_amp; }
    }
    void apply_gain(float gain, size_t len) {                               // sample.rs:108-114
        if (fabsf(gain - 1.0f) < 0.001f) return;
        size_t n = std::min(len, this->len());
        for (size_t i = 0; i < n; ++i) { l[i] *= gain; r[i] *= gain; }
    }
    float scan_max(size_t len) const { return fmaxf(absmaxlen(l, len), absmaxlen(r, len)); }   // :116-118
    void scale(size_t len, float scalar) {                                  // sample.rs:120-123
        size_t nl = std::min(len, l.size()), nr = std::min(len, r.size());
        for (size_t i = 0; i < nl; ++i) l[i] *= scalar;
        for (size_t i = 0; i < nr; ++i) r[i] *= scalar;
    }
    void normalize(size_t len) {                                            // sample.rs:125-130
        len = std::min(len, this->len());
        float max = scan_max(len);
        float scalar = 1.0f / max;
        scale(len, scalar);
    }
    void normalize_seperate() {                                             // sample.rs:132-137
        float scalel = 1.0f / absmax(l), scaler = 1.0f / absmax(r);
        for (auto& s : l) s *= scalel;
        for (auto& s : r) s *= scaler;
    }
    void mix_down() {                                                       // sample.rs:139-147
        size_t n = std::min(l.size(), r.size());
        std::vector<float> mix(n);
        for (size_t i = 0; i < n; ++i) mix[i] = l[i] + r[i];
        float scale = 1.0f / absmax(mix);
        for (auto& s : mix) s *= scale;
        l = mix; r = mix;
    }
};

// ------------------------------------------------------------------ rubato 0.15 SincFixedIn<f32>
// [UNVERIFIED-RECALL] restatement of the parameters termdaw uses
// (sample.rs:152-169, state.rs:534-544): sinc_len 256, f_cutoff 0.95, Linear interpolation,
// oversampling 256, BlackmanHarris2, scalar (8-accumulator) inner product.
struct Resampler {
    size_t sinc_len = 256, factor = 256, chunk = 0;
    double ratio = 1.0, last_index = 0.0;
    std::vector<std::vector<float>> sincs;          // [factor][sinc_len]
    std::vector<float> buf[2];                      // chunk + 2*sinc_len frames per channel

    static float sincf(float v) { if (v == 0.0f) return 1.0f; float x = v * PI_F; return sinf(x) / x; }

    Resampler(double resample_ratio, size_t chunk_size) {
        ratio = resample_ratio; chunk = chunk_size;
        float f_cutoff = 0.95f;
        float cutoff = resample_ratio >= 1.0 ? f_cutoff : f_cutoff * (float)resample_ratio;
        size_t npoints = sinc_len, totpoints = npoints * factor;
        // blackman_harris squared
        std::vector<float> window(totpoints);
        float pi2 = 2.0f * PI_F, pi4 = 4.0f * PI_F, pi6 = 6.0f * PI_F, np_f = (float)totpoints;
        float a0 = 0.35875f, a1 = 0.48829f, a2 = 0.14128f, a3 = 0.01168f;
        for (size_t x = 0; x < totpoints; ++x) {
            float xf = (float)x;
            float w = a0 - a1 * cosf(pi2 * xf / np_f) + a2 * cosf(pi4 * xf / np_f) - a3 * cosf(pi6 * xf / np_f);
            window[x] = w * w;
        }
        std::vector<float> y(totpoints);
        float sum = 0.0f;
        for (size_t x = 0; x < totpoints; ++x) {
            float val = window[x] * sincf(((float)x - (float)(totpoints / 2)) * cutoff / (float)factor);
            sum += val; y[x] = val;
        }
        sum /= (float)factor;
        sincs.assign(factor, std::vector<float>(npoints, 0.0f));
        for (size_t p = 0; p < npoints; ++p)
            for (size_t n = 0; n < factor; ++n) sincs[factor - n - 1][p] = y[factor * p + n] / sum;
        buf[0].assign(chunk + 2 * sinc_len, 0.0f);
        buf[1].assign(chunk + 2 * sinc_len, 0.0f);
        last_index = -(double)(sinc_len / 2);
    }
    float get_sinc_interpolated(const std::vector<float>& wave, size_t index, size_t subindex) const {
        const float* w = wave.data() + index; const float* s = sincs[subindex].data();
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (size_t i = 0; i < sinc_len; i += 8)
            for (int j = 0; j < 8; ++j) acc[j] += w[i + j] * s[i + j];
        return acc[0] + acc[1] + acc[2] + acc[3] + acc[4] + acc[5] + acc[6] + acc[7];
    }
    // one `process` call over exactly `chunk` input frames; appends to out_l/out_r
    void process(const float* in_l, const float* in_r, std::vector<float>& out_l, std::vector<float>& out_r) {
        double t_ratio = 1.0 / ratio;
        int64_t end_idx = (int64_t)chunk - ((int64_t)sinc_len + 1) - (int64_t)std::ceil(t_ratio);
        for (int c = 0; c < 2; ++c) {
            std::memmove(buf[c].data(), buf[c].data() + chunk, 2 * sinc_len * sizeof(float));
            std::memcpy(buf[c].data() + 2 * sinc_len, c == 0 ? in_l : in_r, chunk * sizeof(float));
        }
        double idx = last_index;
        while (idx < (double)end_idx) {
            idx += t_ratio;
            int64_t index = (int64_t)std::floor(idx);
            int64_t subindex = (int64_t)std::floor((idx - std::floor(idx)) * (double)factor);
            int64_t i0 = index, s0 = subindex, i1 = index, s1 = subindex + 1;
            if (s1 >= (int64_t)factor) { s1 -= (int64_t)factor; i1 += 1; }
            double fr = idx * (double)factor - std::floor(idx * (double)factor);
            float frac = (float)fr;
            for (int c = 0; c < 2; ++c) {
                float p0 = get_sinc_interpolated(buf[c], (size_t)(i0 + 2 * (int64_t)sinc_len), (size_t)s0);
                float p1 = get_sinc_interpolated(buf[c], (size_t)(i1 + 2 * (int64_t)sinc_len), (size_t)s1);
                float v = (1.0f - frac) * p0 + frac * p1;
                (c == 0 ? out_l : out_r).push_back(v);
            }
        }
        last_index = idx - (double)chunk;
    }
};

Sample resample_whole(const Sample& s, size_t from, size_t to) {           // sample.rs:150-175
    Resampler rs((double)to / (double)from, s.len());
    Sample out(0);
    rs.process(s.l.data(), s.r.data(), out.l, out.r);
    return out;
}

// ---------------------------------------------------------------------------------- floww.rs
struct FlowwEvent { float t, note, vel; };
struct FlowwBank {                                                          // floww.rs:6-16
    size_t sr = 0, bl = 0, frame = 0, block_index = 0;
    std::vector<std::vector<FlowwEvent>> flowws;
    std::vector<size_t> start_indices;
    void reset() { frame = 0; block_index = 0; flowws.clear(); start_indices.clear(); }   // :21-28
    size_t ev_frame(const FlowwEvent& e) const { return f32_as_usize(e.t * (float)sr); }
    void set_start_indices_to_frame(size_t t_frame, bool do_skip) {         // floww.rs:70-81
        for (size_t i = 0; i < flowws.size(); ++i) {
            size_t skip = do_skip ? start_indices[i] : 0;
            for (size_t j = skip; j < flowws[i].size(); ++j)
                if (ev_frame(flowws[i][j]) >= t_frame) { start_indices[i] = j; break; }
        }
    }
    void set_time(size_t t) { set_start_indices_to_frame(t, false); frame = t; }          // :83-86
    void set_time_to_next_block() { frame += bl; set_start_indices_to_frame(frame, true); }   // :88-91
    void start_block(size_t index) {                                        // floww.rs:93-96
        if (index >= flowws.size()) return;
        block_index = start_indices[index];
    }
    bool get_block_drum(size_t index, size_t offset_frame, float* note, float* vel) {     // :99-121
        if (index >= flowws.size()) return false;
        for (;;) {
            if (block_index >= flowws[index].size()) return false;
            FlowwEvent ne = flowws[index][block_index];
            if (ev_frame(ne) < frame + offset_frame) { block_index += 1; continue; }
            if (ev_frame(ne) == frame + offset_frame) {
                block_index += 1;
                if (ne.vel > 0.001f) { *note = ne.note; *vel = ne.vel; return true; }
            } else return false;
        }
    }
    struct Simple { bool on; float note, vel; };
    void get_block_simple(size_t index, size_t offset_frame, std::vector<Simple>& res) {  // :124-141
        res.clear();
        if (index >= flowws.size()) return;
        for (;;) {
            if (block_index >= flowws[index].size()) break;
            FlowwEvent ne = flowws[index][block_index];
            if (ev_frame(ne) == frame + offset_frame) {
                block_index += 1;
                res.push_back({ne.vel > 0.001f, ne.note, ne.vel});
            } else break;
        }
    }
};

// ------------------------------------------------------------- sampsyn 0.1 (DEFINED semantic, B.4)
// blob: "WTB1", u32 n_frames, u32 n_phase, f32 frames_per_sec, f32 data[n_frames][n_phase] (LE).
struct WaveTable {
    uint32_t n_frames = 1, n_phase = 1024; float fps = 1.0f; std::vector<float> data;
};
WaveTable default_wavetable() {          // stands in for WaveTable::default() (state.rs:421)
    WaveTable t; t.n_frames = 1; t.n_phase = 1024; t.fps = 1.0f; t.data.resize(1024);
    for (uint32_t p = 0; p < 1024; ++p) t.data[p] = sinf(6.2831855f * ((float)p / 1024.0f));
    return t;
}
bool parse_wavetable_from_buffer(const std::vector<uint8_t>& b, WaveTable* out) {   // state.rs:415
    if (b.size() < 16 || std::memcmp(b.data(), "WTB1", 4) != 0) return false;
    uint32_t nf, np; float fps;
    std::memcpy(&nf, b.data() + 4, 4); std::memcpy(&np, b.data() + 8, 4); std::memcpy(&fps, b.data() + 12, 4);
    if (nf < 1 || np < 2) return false;
    if ((uint64_t)nf * np * 4 + 16 != b.size()) return false;
    out->n_frames = nf; out->n_phase = np; out->fps = fps; out->data.resize((size_t)nf * np);
    std::memcpy(out->data.data(), b.data() + 16, (size_t)nf * np * 4);
    return true;
}
struct WaveTableState { float unused = 0.0f; };
WaveTableState initial_state(const WaveTable&, float) { return WaveTableState(); }  // extensions.rs:542
// extensions.rs:569 — `t` is seconds since note-on; `sr` is unused by this definition (stateless).
float wavetable_act_state(const WaveTable& wt, WaveTableState&, float hz, float t, float /*sr*/) {
    float ph = t * hz;
    ph = ph - floorf(ph);
    float x = ph * (float)wt.n_phase;
    uint32_t i0 = (uint32_t)x;
    if (i0 >= wt.n_phase) i0 = wt.n_phase - 1;
    float fx = x - (float)i0;
    uint32_t i1 = i0 + 1; if (i1 >= wt.n_phase) i1 = 0;
    float y = t * wt.fps;
    float ymax = (float)(wt.n_frames - 1);
    if (!(y > 0.0f)) y = 0.0f;
    if (y > ymax) y = ymax;
    uint32_t j0 = (uint32_t)y;
    uint32_t j1 = j0 + 1; if (j1 >= wt.n_frames) j1 = wt.n_frames - 1;
    float fy = y - (float)j0;
    const float* r0 = wt.data.data() + (size_t)j0 * wt.n_phase;
    const float* r1 = wt.data.data() + (size_t)j1 * wt.n_phase;
    float a = r0[i0] + fx * (r0[i1] - r0[i0]);
    float b = r1[i0] + fx * (r1[i1] - r1[i0]);
    return a + fy * (b - a);
}

// ---------------------------------------------------------------------------------- extensions.rs
enum Kind { SUM, NORMALIZE, SAMPLE_LOOP, SAMPLE_MULTI, SAMPLE_LERP, DEBUG_SINE, SYNTH, SAMPSYN, ADSR, BANDPASS };

struct SynthNote { float note, vel, env_t, rel_t; WaveTableState st; };
struct P2 { int64_t t; float vel; };
struct P3 { float t, vel, rel; };

struct VertexExt {                                                          // extensions.rs:15-80
    Kind kind = SUM;
    float max = 0.0f, scan_max = 0.0f;                                      // Normalize
    size_t sample_index = 0, floww_index = 0, t = 0;                        // SampleLoop etc.
    bool has_note = false; size_t note = 0;
    std::deque<P2> ts;                                                      // SampleMulti
    size_t lerp_len = 0, countdown = 0; P2 primary{0, 0.0f}, ghost{0, 0.0f};   // SampleLerp
    std::vector<std::pair<float, float>> dnotes;                            // DebugSine
    OscConf square{}, topflat{}, triangle{};                                // Synth
    std::vector<SynthNote> notes;                                           // Synth / SampSyn
    AdsrConf adsr{}; WaveTable wave_table;                                  // SampSyn / Adsr conf
    bool use_off = false, use_max = false; P3 aprimary{0, 0, 0}, aghost{0, 0, 0};   // Adsr
    float lgamma = 0, hgamma = 0, lprevl = 0, lprevr = 0, hprevl = 0, hprevr = 0;   // BandPass
    bool first = true, pass = true;

    bool has_input() const {                                                // extensions.rs:266-281
        return kind == SUM || kind == NORMALIZE || kind == ADSR || kind == BANDPASS;
    }
    void set_time(size_t time) {                                            // extensions.rs:196-204
        switch (kind) {
            case SAMPLE_LOOP: t = time; break;
            case DEBUG_SINE: dnotes.clear(); break;
            case SYNTH: notes.clear(); break;
            case BANDPASS: first = true; break;
            default: break;
        }
    }
};

void sum_inputs(Sample& buf, size_t len, const std::vector<const Sample*>& res) {   // :310-319
    buf.zero();
    for (const Sample* r : res) {
        size_t l = std::min(r->len(), len);
        for (size_t i = 0; i < l; ++i) { buf.l[i] += r->l[i]; buf.r[i] += r->r[i]; }
    }
}
void normalize_gen(Sample& buf, size_t len, float& max, float& scan_max, bool is_scan) {  // :321-329
    float buf_max = buf.scan_max(len);
    if (is_scan) scan_max = fmaxf(buf_max, scan_max);
    else max = fmaxf(buf_max, max);
    buf.scale(len, 1.0f / max);
}
void sample_loop_gen(Sample& buf, const std::vector<Sample>& sb, size_t len, size_t& t, size_t si) {  // :331-341
    const Sample& sample = sb[si];
    size_t l = sample.len();
    for (size_t i = 0; i < len; ++i) { buf.l[i] = sample.l[(t + i) % l]; buf.r[i] = sample.r[(t + i) % l]; }
    t += len;
}
void sample_multi_gen(Sample& buf, const std::vector<Sample>& sb, FlowwBank& fb, size_t len,
                      VertexExt& v) {                                       // extensions.rs:344-381
    const Sample& sample = sb[v.sample_index];
    fb.start_block(v.floww_index);
    for (size_t i = 0; i < len; ++i) {
        float note, vel;
        if (fb.get_block_drum(v.floww_index, i, &note, &vel)) {
            bool ok = v.has_note ? fabsf(note - (float)v.note) < 0.01f : true;
            if (ok) v.ts.push_back({-(int64_t)i, vel});
        }
        buf.l[i] = 0.0f; buf.r[i] = 0.0f;
        size_t pops = 0;
        for (const P2& h : v.ts) {
            size_t pos = (size_t)std::max<int64_t>(h.t + (int64_t)i, 0);
            if (pos >= sample.len()) pops += 1;
            else { buf.l[i] += sample.l[pos] * h.vel; buf.r[i] += sample.r[pos] * h.vel; }
        }
        for (size_t k = 0; k < pops; ++k) if (!v.ts.empty()) v.ts.pop_front();
    }
    for (P2& h : v.ts) h.t += (int64_t)len;
}
void sample_lerp_gen(Sample& buf, const std::vector<Sample>& sb, FlowwBank& fb, size_t len,
                     VertexExt& v) {                                        // extensions.rs:384-421
    const Sample& sample = sb[v.sample_index];
    fb.start_block(v.floww_index);
    for (size_t i = 0; i < len; ++i) {
        float note, vel;
        if (fb.get_block_drum(v.floww_index, i, &note, &vel)) {
            bool ok = v.has_note ? fabsf(note - (float)v.note) < 0.01f : true;
            if (ok) { v.ghost = v.primary; v.primary = {-(int64_t)i, vel}; v.countdown = v.lerp_len; }
        }
        size_t ppos = std::min((size_t)std::max<int64_t>(v.primary.t + (int64_t)i, 0), sample.len() - 1);
        float l = sample.l[ppos] * v.primary.vel;
        float r = sample.r[ppos] * v.primary.vel;
        if (v.countdown > 0) {
            v.countdown -= 1;
            float t = (float)v.countdown / (float)v.lerp_len;
            size_t gpos = std::min((size_t)std::max<int64_t>(v.ghost.t + (int64_t)i, 0), sample.len() - 1);
            float gl = sample.l[gpos] * v.ghost.vel;
            float gr = sample.r[gpos] * v.ghost.vel;
            l = gl * t + l * (1.0f - t);
            r = gr * t + r * (1.0f - t);
        }
        buf.l[i] = l; buf.r[i] = r;
    }
    v.primary.t += (int64_t)len;
    v.ghost.t += (int64_t)len;
}
void debug_sine_gen(Sample& buf, FlowwBank& fb, size_t len, VertexExt& v, size_t t, size_t sr) {  // :423-457
    fb.start_block(v.floww_index);
    std::vector<FlowwBank::Simple> evs;
    for (size_t i = 0; i < len; ++i) {
        fb.get_block_simple(v.floww_index, i, evs);
        for (auto& e : evs) {
            if (e.on) {
                bool has = false;
                for (auto& nv : v.dnotes)
                    if (fabsf(nv.first - e.note) < 0.001f) { nv.second = e.vel; has = true; break; }
                if (!has) v.dnotes.push_back({e.note, e.vel});
            } else {
                auto& d = v.dnotes;
                d.erase(std::remove_if(d.begin(), d.end(), [&](const std::pair<float, float>& x) {
                            return !(fabsf(x.first - e.note) > 0.001f); }), d.end());
            }
        }
        buf.l[i] = 0.0f; buf.r[i] = 0.0f;
        for (auto& nv : v.dnotes) {
            float time = (float)(t + i) / (float)sr;
            float hz = 440.0f * powf(2.0f, (nv.first - 69.0f) / 12.0f);
            float s = sinf(time * hz * 2.0f * PI_F) * nv.second;
            buf.l[i] += s; buf.r[i] += s;
        }
    }
}
// note-off handling shared by synth_gen / sampsyn_gen (extensions.rs:485-494, 545-554)
void synth_note_off(std::vector<SynthNote>& notes, float note, size_t i, size_t sr) {
    notes.erase(std::remove_if(notes.begin(), notes.end(), [&](const SynthNote& x) {
                    return !(fabsf(x.note - note) > 0.001f || x.rel_t == 0.0f); }), notes.end());
    for (auto& n : notes) {
        if (fabsf(n.note - note) > 0.001f) continue;
        if (n.rel_t == 0.0f) {
            n.rel_t = n.env_t + ((float)i / (float)sr);
            n.env_t = -((float)i / (float)sr);
        }   // else: unreachable in the reference (panic), all releasing same-note voices were removed
    }
}
void synth_gen(Sample& buf, FlowwBank& fb, size_t len, VertexExt& v, size_t t, size_t sr) {   // :460-529
    const OscConf &square = v.square, &topflat = v.topflat, &triangle = v.triangle;
    float osc_amp_multiplier = 1.0f / (square.volume * max_vel(square.adsr) +
                                       topflat.volume * max_vel(topflat.adsr) +
                                       triangle.volume * max_vel(triangle.adsr));
    float release_sec = 0.0f;
    if (square.volume > 0.0f) release_sec = square.adsr.release_sec;
    if (topflat.volume > 0.0f) release_sec = fmaxf(release_sec, topflat.adsr.release_sec);
    if (triangle.volume > 0.0f) release_sec = fmaxf(release_sec, triangle.adsr.release_sec);
    fb.start_block(v.floww_index);
    std::vector<FlowwBank::Simple> evs;
    for (size_t i = 0; i < len; ++i) {
        fb.get_block_simple(v.floww_index, i, evs);
        for (auto& e : evs) {
            if (e.on) v.notes.push_back({e.note, e.vel, -((float)i / (float)sr), 0.0f, WaveTableState()});
            else synth_note_off(v.notes, e.note, i, sr);
        }
        buf.l[i] = 0.0f; buf.r[i] = 0.0f;
        for (auto& n : v.notes) {
            float time = (float)(t + i) / (float)sr;
            float env_time = n.env_t + ((float)i / (float)sr);
            float hz = 440.0f * powf(2.0f, (n.note - 69.0f) / 12.0f);
            auto env_vel = [&](const AdsrConf& c) {
                return n.rel_t == 0.0f ? apply_ads(c, env_time) : apply_r_rt(c, env_time, n.rel_t);
            };
            float s = 0.0f;
            if (square.volume > 0.0f)
                s += square_sine_sample(time, hz, square.param) * n.vel * env_vel(square.adsr) * square.volume;
            if (topflat.volume > 0.0f)
                s += topflat_sine_sample(time, hz, topflat.param) * n.vel * env_vel(topflat.adsr) * topflat.volume;
            if (triangle.volume > 0.0f)
                s += triangle_sample(time, hz) * n.vel * env_vel(triangle.adsr) * triangle.volume;
            s *= osc_amp_multiplier;
            buf.l[i] += s; buf.r[i] += s;
        }
    }
    for (auto& n : v.notes) n.env_t += (float)len / (float)sr;
    v.notes.erase(std::remove_if(v.notes.begin(), v.notes.end(), [&](const SynthNote& x) {
                      return !(x.rel_t == 0.0f || x.env_t <= release_sec); }), v.notes.end());
}
void sampsyn_gen(Sample& buf, FlowwBank& fb, size_t len, VertexExt& v, size_t sr) {   // :532-578
    const AdsrConf& adsr = v.adsr;
    float amp_multiplier = 1.0f / max_vel(adsr);
    fb.start_block(v.floww_index);
    std::vector<FlowwBank::Simple> evs;
    for (size_t i = 0; i < len; ++i) {
        fb.get_block_simple(v.floww_index, i, evs);
        for (auto& e : evs) {
            if (e.on) {
                WaveTableState st = initial_state(v.wave_table, 0.0f);
                v.notes.push_back({e.note, e.vel, -((float)i / (float)sr), 0.0f, st});
            } else synth_note_off(v.notes, e.note, i, sr);
        }
        buf.l[i] = 0.0f; buf.r[i] = 0.0f;
        for (auto& n : v.notes) {
            float env_time = n.env_t + ((float)i / (float)sr);
            float hz = 440.0f * powf(2.0f, (n.note - 69.0f) / 12.0f);
            float ev = n.rel_t == 0.0f ? apply_ads(adsr, env_time) : apply_r_rt(adsr, env_time, n.rel_t);
            float s = 0.0f;
            float vel = n.vel * ev * amp_multiplier;
            s += wavetable_act_state(v.wave_table, n.st, hz, env_time + n.rel_t, (float)sr) * vel;
            buf.l[i] += s; buf.r[i] += s;
        }
    }
    for (auto& n : v.notes) n.env_t += (float)len / (float)sr;
    v.notes.erase(std::remove_if(v.notes.begin(), v.notes.end(), [&](const SynthNote& x) {
                      return !(x.rel_t == 0.0f || x.env_t <= adsr.release_sec); }), v.notes.end());
}
void adsr_gen(Sample& buf, size_t len, FlowwBank& fb, float wet, VertexExt& v, size_t sr) {   // :593-651
    if (wet < 0.0001f) return;
    const AdsrConf& conf = v.adsr;
    float maxmul = v.use_max ? 1.0f : 0.0f;
    float minmul = 1.0f - maxmul;
    P3 &primary = v.aprimary, &ghost = v.aghost;
    fb.start_block(v.floww_index);
    if (v.use_off) {
        std::vector<FlowwBank::Simple> evs;
        for (size_t i = 0; i < len; ++i) {
            float offset = (float)i / (float)sr;
            fb.get_block_simple(v.floww_index, i, evs);
            for (auto& e : evs) {
                if (v.has_note) { if (fabsf((float)v.note - e.note) > 0.01f) continue; }
                if (e.on) {
                    ghost = primary;
                    primary = {-((float)i / (float)sr), e.vel, 0.0f};
                } else if (ghost.rel == 0.0f) {
                    ghost.t = -((float)i / (float)sr);
                    ghost.rel = apply_ads(conf, ghost.t + offset) * ghost.vel;
                } else {
                    primary.t = -((float)i / (float)sr);
                    primary.rel = apply_ads(conf, primary.t + offset) * primary.vel;
                }
            }
            float pvel = primary.rel == 0.0f ? apply_ads(conf, primary.t + offset) * primary.vel
                                             : apply_r(conf, primary.t + offset, primary.rel) * primary.vel;
            float gvel = ghost.rel == 0.0f ? apply_ads(conf, ghost.t + offset) * ghost.vel
                                           : apply_r(conf, ghost.t + offset, ghost.rel) * ghost.vel;
            float adsr_vel = fmaxf(pvel, gvel) * maxmul + fminf(pvel, gvel) * minmul;
            float vel = lerp(1.0f, adsr_vel, wet);
            buf.l[i] *= vel; buf.r[i] *= vel;
        }
    } else {
        for (size_t i = 0; i < len; ++i) {
            float n, vv;
            if (fb.get_block_drum(v.floww_index, i, &n, &vv)) {
                if (v.has_note) { if (fabsf((float)v.note - n) > 0.01f) continue; }   // skips the FRAME (:633-635)
                ghost = primary;
                primary = {-((float)i / (float)sr), vv, 0.0f};
            }
            float offset = (float)i / (float)sr;
            float pvel = apply_adsr(conf, primary.t + offset) * primary.vel;
            float gvel = apply_adsr(conf, ghost.t + offset) * ghost.vel;
            float adsr_vel = fmaxf(pvel, gvel) * maxmul + fminf(pvel, gvel) * minmul;
            float vel = lerp(1.0f, adsr_vel, wet);
            buf.l[i] *= vel; buf.r[i] *= vel;
        }
    }
    primary.t += (float)len / (float)sr;
    ghost.t += (float)len / (float)sr;
}
void band_pass_gen(Sample& buf, size_t len, float wet, VertexExt& v) {      // extensions.rs:654-689
    if (wet < 0.0001f) return;
    float lgamma = v.lgamma, hgamma = v.hgamma;
    if (lgamma == 0.0f && hgamma == 0.0f) return;
    float lmul = lgamma == 0.0f ? 0.0f : 1.0f;
    float hmul = hgamma == 0.0f ? 0.0f : 1.0f;
    float pass_mul = v.pass ? 1.0f : 0.0f;
    float cut_mul = 1.0f - pass_mul;
    if (v.first) {
        v.lprevl = buf.l[0]; v.lprevr = buf.r[0]; v.hprevl = buf.l[0]; v.hprevr = buf.r[0];
        v.first = false;
    }
    for (size_t i = 0; i < len; ++i) {
        float l = buf.l[i], r = buf.r[i];
        float ll = v.lprevl + lgamma * (l - v.lprevl);
        float lr = v.lprevr + lgamma * (r - v.lprevr);
        float hl = v.hprevl + hgamma * (l - v.hprevl);
        float hr = v.hprevr + hgamma * (r - v.hprevr);
        v.lprevl = ll; v.lprevr = lr; v.hprevl = hl; v.hprevr = hr;
        float cutl = (lmul * ll + hmul * (l - hl)) * 0.5f;
        float cutr = (lmul * lr + hmul * (r - hr)) * 0.5f;
        float passl = l - cutl;
        float passr = r - cutl;           // sic: the reference subtracts the LEFT cut (:685)
        buf.l[i] = cutl * cut_mul + passl * pass_mul;
        buf.r[i] = cutr * cut_mul + passr * pass_mul;
    }
}

// ---------------------------------------------------------------------------------- graph.rs
struct Vertex {                                                             // graph.rs:242-248
    Sample buf; float gain, angle, wet; VertexExt ext;
    Vertex(size_t bl, float g, float a, float w, VertexExt e)               // graph.rs:251-259
        : buf(bl), gain(g), angle(fmaxf(fminf(a, 90.0f), -90.0f)), wet(fmaxf(fminf(w, 1.0f), 0.0f)),
          ext(std::move(e)) {}
};

struct Ctx {
    size_t sr, bl;
    std::string err;
    // banks
    std::vector<Sample> samples;
    std::vector<std::vector<uint8_t>> buffers;
    FlowwBank fb;
    // graph (graph.rs:12-22)
    std::vector<Vertex> vertices;
    std::vector<std::vector<size_t>> edges;
    std::vector<char> ran_status;
    bool has_output = false; size_t output_vertex = 0;
    size_t t = 0;
    // taps
    bool taps = false;
    std::vector<Sample> tap;     // per vertex, whole last pass

    int fail(int code, const std::string& m) { err = m; return code; }

    void generate(size_t index, size_t t_arg, bool is_scan, const std::vector<const Sample*>& ins) {
        Vertex& vx = vertices[index];
        size_t len = std::min(vx.buf.len(), bl);                            // graph.rs:270
        VertexExt& e = vx.ext;
        if (e.has_input()) sum_inputs(vx.buf, len, ins);                    // extensions.rs:212-214
        switch (e.kind) {                                                   // extensions.rs:215-261
            case SUM: break;
            case NORMALIZE: normalize_gen(vx.buf, len, e.max, e.scan_max, is_scan); break;
            case SAMPLE_LOOP: sample_loop_gen(vx.buf, samples, len, e.t, e.sample_index); break;
            case SAMPLE_MULTI: sample_multi_gen(vx.buf, samples, fb, len, e); break;
            case SAMPLE_LERP: sample_lerp_gen(vx.buf, samples, fb, len, e); break;
            case DEBUG_SINE: debug_sine_gen(vx.buf, fb, len, e, t_arg, sr); break;
            case SYNTH: synth_gen(vx.buf, fb, len, e, t_arg, sr); break;
            case SAMPSYN: sampsyn_gen(vx.buf, fb, len, e, sr); break;
            case ADSR: adsr_gen(vx.buf, len, fb, vx.wet, e, sr); break;
            case BANDPASS: band_pass_gen(vx.buf, len, vx.wet, e); break;
        }
        vx.buf.apply_angle(vx.angle, len);                                  // extensions.rs:262
        vx.buf.apply_gain(vx.gain, len);                                    // extensions.rs:263
        if (taps) {
            tap[index].l.insert(tap[index].l.end(), vx.buf.l.begin(), vx.buf.l.begin() + len);
            tap[index].r.insert(tap[index].r.end(), vx.buf.r.begin(), vx.buf.r.begin() + len);
        }
    }
    void run_vertex(size_t t_arg, size_t index, bool is_scan) {             // graph.rs:98-121
        if (index >= vertices.size()) return;
        if (ran_status[index]) return;
        ran_status[index] = 1;
        std::vector<size_t> e = edges[index];
        for (size_t incoming : e) run_vertex(t_arg, incoming, is_scan);
        std::vector<const Sample*> ins;
        for (size_t incoming : e) ins.push_back(&vertices[incoming].buf);
        generate(index, t_arg, is_scan, ins);
    }
    void set_time(size_t time) {                                            // graph.rs:123-128
        t = time;
        for (auto& v : vertices) v.ext.set_time(time);
    }
    void reset_ran() { std::fill(ran_status.begin(), ran_status.end(), 0); }
    const Sample* render_chunk() {                                          // graph.rs:182-193
        reset_ran();
        if (!has_output) return nullptr;
        run_vertex(t, output_vertex, false);
        t += bl;
        return &vertices[output_vertex].buf;
    }
    void clear_taps() { tap.assign(vertices.size(), Sample(0)); }
    void true_normalize_scan(size_t chunks) {                               // graph.rs:222-237
        if (!has_output) return;
        for (auto& v : vertices) if (v.ext.kind == NORMALIZE) v.ext.scan_max = 0.0f;   // :195-199
        fb.set_time(0);
        for (size_t j = 0; j < chunks; ++j) {
            reset_ran();
            run_vertex(j * bl, output_vertex, true);
            fb.set_time_to_next_block();
        }
        for (auto& v : vertices) if (v.ext.kind == NORMALIZE) v.ext.max = v.ext.scan_max;   // :201-205
        set_time(0);
        fb.set_time(0);
    }
    bool has_loop(size_t x, size_t b) const {                               // graph.rs:65-71
        if (x == b) return true;
        for (size_t y : edges[x]) if (has_loop(y, b)) return true;
        return false;
    }
    int add_vertex(Vertex v) {                                              // graph.rs:49-56
        vertices.push_back(std::move(v));
        ran_status.push_back(0);
        edges.emplace_back();
        return (int)vertices.size() - 1;
    }
};

// Rust `f32 as i16` / `as i32`: toward zero, saturating, NaN -> 0.
inline int32_t f32_as_i32(float x) {
    if (!(x == x)) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}
inline int16_t f32_as_i16(float x) {
    if (!(x == x)) return 0;
    if (x >= 32767.0f) return INT16_MAX;
    if (x <= -32768.0f) return INT16_MIN;
    return (int16_t)x;
}

// hound 3.5 `write_sample` narrowing for an Int spec. returns false on TooWide.
bool pack_sample(std::vector<uint8_t>& out, int32_t v, uint32_t bd) {
    switch (bd) {
        case 8: if (v < -128 || v > 127) return false; out.push_back((uint8_t)(v + 128)); return true;
        case 16: if (v < -32768 || v > 32767) return false;
                 out.push_back((uint8_t)(v & 0xff)); out.push_back((uint8_t)((v >> 8) & 0xff)); return true;
        case 24: if (v < -8388608 || v > 8388607) return false;
                 out.push_back((uint8_t)(v & 0xff)); out.push_back((uint8_t)((v >> 8) & 0xff));
                 out.push_back((uint8_t)((v >> 16) & 0xff)); return true;
        default: out.push_back((uint8_t)(v & 0xff)); out.push_back((uint8_t)((v >> 8) & 0xff));
                 out.push_back((uint8_t)((v >> 16) & 0xff)); out.push_back((uint8_t)((v >> 24) & 0xff)); return true;
    }
}

}  // namespace

struct tdo_ctx { Ctx c; };

extern "C" {

int tdo_sample_mode_from_string(const char* s) {                            // sample.rs:202-211
    std::string m = s ? s : "";
    if (m == "left") return 1; if (m == "right") return 2; if (m == "loudest") return 3;
    if (m == "normalize-seperate") return 4; if (m == "mix-down") return 5; return 0;
}
int tdo_ctx_create(uint32_t project_sr, uint32_t buffer_len, int /*device*/, tdo_ctx** out) {
    if (!out || project_sr == 0 || buffer_len == 0) return ERR_INVALID;
    tdo_ctx* x = new tdo_ctx();
    x->c.sr = project_sr; x->c.bl = buffer_len; x->c.fb.sr = project_sr; x->c.fb.bl = buffer_len;
    *out = x; return OK;
}
void tdo_ctx_destroy(tdo_ctx* x) { delete x; }
const char* tdo_last_error(tdo_ctx* x) { return x ? x->c.err.c_str() : "null ctx"; }
uint64_t tdo_kernel_launches(tdo_ctx*) { return 0; }

// numeric part of SampleBank::add, sample.rs:262-310
int tdo_sample_add(tdo_ctx* x, const float* linear, uint64_t n_values, uint32_t channels,
                   uint32_t src_sr, int mode) {
    Ctx& c = x->c;
    if (mode == 0 && channels != 2) return c.fail(ERR_SAMPLE, "only 2 channel samples are supported for stereo samples");
    if (mode != 0 && (channels > 2 || channels < 1)) return c.fail(ERR_SAMPLE, "only 1,2 channel samples are supported");
    std::vector<float> l, r;
    if (channels == 1) {                                                    // sample.rs:277-282
        if (mode == 1) l.assign(linear, linear + n_values); else r.assign(linear, linear + n_values);
    } else {                                                                // sample.rs:283-292
        uint64_t half = n_values / 2;
        for (uint64_t i = 0; i < half; ++i) { l.push_back(linear[i * 2]); r.push_back(linear[i * 2 + 1]); }
        if (n_values > half * 2) l.push_back(linear[n_values - 1]);
    }
    Sample s(0);
    switch (mode) {                                                         // Sample::from, sample.rs:38-77
        case 1: if (l.empty()) return c.fail(ERR_SAMPLE, "Sample::from: l has length 0"); s.l = l; s.r = l; break;
        case 2: if (r.empty()) return c.fail(ERR_SAMPLE, "Sample::from: r has length 0"); s.l = r; s.r = r; break;
        case 3: { float lm = mean_energy(l), rm = mean_energy(r);
                  if (lm > rm) { s.l = l; s.r = l; } else { s.l = r; s.r = r; } break; }
        default:
            if (l.size() != r.size()) return c.fail(ERR_SAMPLE, "Sample::from: l and r do not have the same length");
            if (l.empty()) return c.fail(ERR_SAMPLE, "Sample::from: l and r have length 0");
            s.l = l; s.r = r;
    }
    if (mode == 4) s.normalize_seperate();                                  // sample.rs:297-303
    else if (mode == 5) s.mix_down();
    else s.normalize(SIZE_MAX);
    if (src_sr != c.sr) s = resample_whole(s, src_sr, c.sr);                // sample.rs:305-310
    c.samples.push_back(std::move(s));
    return (int)c.samples.size() - 1;
}
int64_t tdo_sample_frames(tdo_ctx* x, int sample) {
    if (sample < 0 || (size_t)sample >= x->c.samples.size()) return ERR_INVALID;
    return (int64_t)x->c.samples[sample].len();
}
int tdo_sample_read(tdo_ctx* x, int sample, float* l, float* r, uint64_t cap) {
    if (sample < 0 || (size_t)sample >= x->c.samples.size()) return ERR_INVALID;
    const Sample& s = x->c.samples[sample];
    if (cap < s.l.size() || cap < s.r.size()) return ERR_CAPACITY;
    std::memcpy(l, s.l.data(), s.l.size() * 4); std::memcpy(r, s.r.data(), s.r.size() * 4);
    return OK;
}
// SampleBank::mark_dead + refresh (sample.rs:316-336) / the resource sweep (state.rs:222-237): slots are emptied,
// the other indices stay valid (same contract as tdg_sample_remove / tdg_resource_remove)
int tdo_sample_remove(tdo_ctx* x, int sample) {
    if (sample < 0 || (size_t)sample >= x->c.samples.size()) return ERR_INVALID;
    x->c.samples[sample] = Sample();
    return OK;
}
int tdo_resource_remove(tdo_ctx* x, int resource) {
    if (resource < 0 || (size_t)resource >= x->c.buffers.size()) return ERR_INVALID;
    std::vector<uint8_t>().swap(x->c.buffers[resource]);
    return OK;
}
int tdo_resource_add(tdo_ctx* x, const uint8_t* bytes, uint64_t n) {
    x->c.buffers.emplace_back(bytes, bytes + n);
    return (int)x->c.buffers.size() - 1;
}
int tdo_floww_add(tdo_ctx* x, const float* t, const float* note, const float* vel, uint64_t n) {
    std::vector<FlowwEvent> f(n);
    for (uint64_t i = 0; i < n; ++i) f[i] = {t[i], note[i], vel[i]};
    x->c.fb.flowws.push_back(std::move(f));
    x->c.fb.start_indices.push_back(0);
    return (int)x->c.fb.flowws.size() - 1;
}
int tdo_floww_append(tdo_ctx* x, int floww, const float* t, const float* note, const float* vel, uint64_t n) {   // floww.rs:55-57
    if (floww < 0 || (size_t)floww >= x->c.fb.flowws.size()) return x->c.fail(ERR_INVALID, "unknown floww");
    for (uint64_t i = 0; i < n; ++i) x->c.fb.flowws[(size_t)floww].push_back({t[i], note[i], vel[i]});
    return OK;
}
int tdo_floww_reset(tdo_ctx* x) { x->c.fb.reset(); return OK; }
int tdo_graph_reset(tdo_ctx* x) {                                           // graph.rs:39-47
    Ctx& c = x->c; c.vertices.clear(); c.edges.clear(); c.ran_status.clear(); c.has_output = false; c.t = 0;
    return OK;
}

static int chk_sample(Ctx& c, int s) { return (s < 0 || (size_t)s >= c.samples.size() || c.samples[(size_t)s].l.empty()) ? c.fail(ERR_INVALID, "unknown sample") : OK; }   // emptied slot: removed sample
static int chk_floww(Ctx& c, int f) { return (f < 0 || (size_t)f >= c.fb.flowws.size()) ? c.fail(ERR_INVALID, "unknown floww") : OK; }

int tdo_add_sum(tdo_ctx* x, float gain, float angle) {
    VertexExt e; e.kind = SUM; return x->c.add_vertex(Vertex(x->c.bl, gain, angle, 0.0f, e));
}
int tdo_add_normalize(tdo_ctx* x, float gain, float angle) {
    VertexExt e; e.kind = NORMALIZE; e.max = 0.0f; e.scan_max = 0.0f;
    return x->c.add_vertex(Vertex(x->c.bl, gain, angle, 0.0f, e));
}
int tdo_add_sample_loop(tdo_ctx* x, float gain, float angle, int sample) {
    if (int r = chk_sample(x->c, sample)) return r;
    VertexExt e; e.kind = SAMPLE_LOOP; e.sample_index = sample; e.t = 0;
    return x->c.add_vertex(Vertex(x->c.bl, gain, angle, 0.0f, e));
}
int tdo_add_sample_multi(tdo_ctx* x, float gain, float angle, int sample, int floww, int note) {
    if (int r = chk_sample(x->c, sample)) return r;
    if (int r = chk_floww(x->c, floww)) return r;
    VertexExt e; e.kind = SAMPLE_MULTI; e.sample_index = sample; e.floww_index = floww;
    e.has_note = note >= 0; e.note = note >= 0 ? (size_t)note : 0;
    return x->c.add_vertex(Vertex(x->c.bl, gain, angle, 0.0f, e));
}
int tdo_add_sample_lerp(tdo_ctx* x, float gain, float angle, int sample, int floww, int note, int lerp_len) {
    if (int r = chk_sample(x->c, sample)) return r;
    if (int r = chk_floww(x->c, floww)) return r;
    VertexExt e; e.kind = SAMPLE_LERP; e.sample_index = sample; e.floww_index = floww;
    e.has_note = note >= 0; e.note = note >= 0 ? (size_t)note : 0;
    e.lerp_len = (size_t)std::max(lerp_len, 0);                             // state.rs:370
    return x->c.add_vertex(Vertex(x->c.bl, gain, angle, 0.0f, e));
}
int tdo_add_debug_sine(tdo_ctx* x, float gain, float angle, int floww) {
    if (int r = chk_floww(x->c, floww)) return r;
    VertexExt e; e.kind = DEBUG_SINE; e.floww_index = floww;
    return x->c.add_vertex(Vertex(x->c.bl, gain, angle, 0.0f, e));
}
int tdo_add_synth(tdo_ctx* x, float gain, float angle, int floww, float sq_vol, float sq_z,
                  const float* sq_adsr, int sq_n, float tf_vol, float tf_z, const float* tf_adsr,
                  int tf_n, float tr_vol, const float* tr_adsr, int tr_n) {
    if (int r = chk_floww(x->c, floww)) return r;
    VertexExt e; e.kind = SYNTH; e.floww_index = floww;
    AdsrConf a, b, c3;
    if (!build_adsr_conf(sq_adsr, sq_n, &a) || !build_adsr_conf(tf_adsr, tf_n, &b) ||
        !build_adsr_conf(tr_adsr, tr_n, &c3))
        return x->c.fail(ERR_ADSR_CONF, "ADSR config must have 6 or 9 elements");
    e.square = {sq_vol, fmaxf(sq_z, 0.0001f), a};                           // state.rs:400
    e.topflat = {tf_vol, tf_z, b};
    e.triangle = {tr_vol, 0.0f, c3};
    return x->c.add_vertex(Vertex(x->c.bl, gain, angle, 0.0f, e));
}
int tdo_add_sampsyn(tdo_ctx* x, float gain, float angle, int floww, const float* adsr, int n, int resource) {
    if (int r = chk_floww(x->c, floww)) return r;
    if (resource < 0 || (size_t)resource >= x->c.buffers.size()) return x->c.fail(ERR_INVALID, "unknown resource");
    VertexExt e; e.kind = SAMPSYN; e.floww_index = floww;
    if (!build_adsr_conf(adsr, n, &e.adsr)) return x->c.fail(ERR_ADSR_CONF, "ADSR config must have 6 or 9 elements");
    if (!parse_wavetable_from_buffer(x->c.buffers[resource], &e.wave_table)) e.wave_table = default_wavetable();
    return x->c.add_vertex(Vertex(x->c.bl, gain, angle, 0.0f, e));
}
int tdo_add_lv2fx(tdo_ctx* x, float, float, float, int) {
    return x->c.fail(ERR_UNSUPPORTED_VERTEX, "lv2fx vertices are outside the restated path");
}
int tdo_add_adsr(tdo_ctx* x, float gain, float angle, float wet, int floww, int use_off, int use_max,
                 int note, const float* conf, int n) {
    if (int r = chk_floww(x->c, floww)) return r;
    VertexExt e; e.kind = ADSR; e.floww_index = floww; e.use_off = use_off != 0; e.use_max = use_max != 0;
    e.has_note = note >= 0; e.note = note >= 0 ? (size_t)note : 0;
    if (!build_adsr_conf(conf, n, &e.adsr)) return x->c.fail(ERR_ADSR_CONF, "ADSR config must have 6 or 9 elements");
    return x->c.add_vertex(Vertex(x->c.bl, gain, angle, wet, e));
}
int tdo_add_bandpass(tdo_ctx* x, float gain, float angle, float wet, float lo, float hi, int pass) {
    VertexExt e; e.kind = BANDPASS;                                         // extensions.rs:173-194
    float lco = fmaxf(fminf(lo, 20000.0f), 0.0f);
    float hco = fmaxf(fminf(hi, 20000.0f), 0.0f);
    e.lgamma = 1.0f - powf(E_F, -2.0f * PI_F * lco / (float)x->c.sr);
    e.hgamma = 1.0f - powf(E_F, -2.0f * PI_F * hco / (float)x->c.sr);
    e.first = true; e.pass = pass != 0;
    return x->c.add_vertex(Vertex(x->c.bl, gain, angle, wet, e));
}
int tdo_connect(tdo_ctx* x, int a, int b) {                                 // graph.rs:58-78
    Ctx& c = x->c;
    if (a < 0 || b < 0) return 0;
    if (a == b) return 0;
    size_t len = c.vertices.size();
    if ((size_t)a >= len || (size_t)b >= len) return 0;
    if (!c.vertices[b].ext.has_input()) return 0;
    if (c.has_loop((size_t)a, (size_t)b)) return 0;
    c.edges[b].push_back((size_t)a);
    return 1;
}
int tdo_set_output(tdo_ctx* x, int v) {
    if (v < 0 || (size_t)v >= x->c.vertices.size()) return ERR_INVALID;
    x->c.has_output = true; x->c.output_vertex = (size_t)v; return OK;
}
int tdo_check_graph(tdo_ctx* x, int* unreachable, int cap, int* n_unreachable) {   // graph.rs:150-174
    Ctx& c = x->c;
    if (n_unreachable) *n_unreachable = 0;
    if (!c.has_output) { c.err = "output vertex not found"; return 0; }
    if (c.edges[c.output_vertex].empty() && c.vertices[c.output_vertex].ext.has_input()) {
        c.err = "output receives no inputs"; return 0;
    }
    std::vector<char> set(c.vertices.size(), 0);
    std::vector<size_t> stack{c.output_vertex};
    while (!stack.empty()) { size_t v = stack.back(); stack.pop_back(); if (set[v]) continue; set[v] = 1;
                             for (size_t y : c.edges[v]) stack.push_back(y); }
    int n = 0;
    for (size_t i = 0; i < set.size(); ++i) if (!set[i]) { if (unreachable && n < cap) unreachable[n] = (int)i; ++n; }
    if (n_unreachable) *n_unreachable = n;
    return 1;
}
int tdo_reset_normalize(tdo_ctx* x) {                                       // extensions.rs:295-299
    for (auto& v : x->c.vertices) if (v.ext.kind == NORMALIZE) v.ext.max = 0.000001f;
    return OK;
}
float tdo_get_normalize_value(tdo_ctx* x, int v) {                          // extensions.rs:301-307
    if (v < 0 || (size_t)v >= x->c.vertices.size()) return -1.0f;
    return x->c.vertices[v].ext.kind == NORMALIZE ? x->c.vertices[v].ext.max : -1.0f;
}
int tdo_set_time(tdo_ctx* x, uint64_t frame) { x->c.set_time(frame); return OK; }
uint64_t tdo_change_time(tdo_ctx* x, uint64_t delta, int plus) {            // graph.rs:130-135
    Ctx& c = x->c;
    size_t nt = plus ? c.t + delta : c.t - std::min<size_t>(delta, c.t);
    c.set_time(nt); return nt;
}
uint64_t tdo_get_time(tdo_ctx* x) { return x->c.t; }
int tdo_floww_set_time(tdo_ctx* x, uint64_t frame) { x->c.fb.set_time(frame); return OK; }

int tdo_enable_taps(tdo_ctx* x, int on) { x->c.taps = on != 0; return OK; }
int tdo_read_vertex_buffer(tdo_ctx* x, int vertex, uint64_t frame0, uint64_t n, float* l, float* r) {
    Ctx& c = x->c;
    if (vertex < 0 || (size_t)vertex >= c.tap.size()) return c.fail(ERR_INVALID, "no tap for vertex");
    const Sample& s = c.tap[vertex];
    if (frame0 + n > s.l.size()) return c.fail(ERR_CAPACITY, "tap shorter than request");
    std::memcpy(l, s.l.data() + frame0, n * 4); std::memcpy(r, s.r.data() + frame0, n * 4);
    return OK;
}

int tdo_normalize_scan(tdo_ctx* x, uint64_t chunks) {
    if (x->c.taps) x->c.clear_taps();
    x->c.true_normalize_scan(chunks);
    return OK;
}

uint64_t tdo_render_bound_bytes(tdo_ctx* x, uint64_t chunks, uint32_t render_sr, uint32_t bd) {
    (void)render_sr;
    return chunks * x->c.bl * 2 * (bd / 8);   // down-sampling only ever shrinks
}

// State::render, state.rs:477-577 (the WAV container itself is written by the caller).
int tdo_render(tdo_ctx* x, uint64_t chunks, uint32_t render_sr, uint32_t bd, uint8_t* out_pcm,
               uint64_t cap_bytes, uint64_t* frames_written) {
    Ctx& c = x->c;
    if (!(bd == 8 || bd == 16 || bd == 24 || bd == 32)) return c.fail(ERR_BITDEPTH, "bitdepth not in {8,16,24,32}");
    if (c.taps) c.clear_taps();
    float amplitude = bd < 32 ? (float)((1 << (bd - 1)) - 1) : (float)INT32_MAX;   // state.rs:515-516
    std::vector<uint8_t> out;
    uint64_t frames = 0;
    bool too_wide = false;
    auto write = [&](const float* l, const float* r, size_t len) {          // state.rs:517-532
        for (size_t i = 0; i < len; ++i) {
            int32_t a, b;
            if (bd > 16) { a = f32_as_i32(l[i] * amplitude); b = f32_as_i32(r[i] * amplitude); }
            else { a = f32_as_i16(l[i] * amplitude); b = f32_as_i16(r[i] * amplitude); }
            if (!pack_sample(out, a, bd) || !pack_sample(out, b, bd)) { too_wide = true; return; }
        }
        frames += len;
    };
    if (c.sr > render_sr) {                                                 // state.rs:533-561
        Resampler rs((double)render_sr / (double)c.sr, c.bl);
        std::vector<float> ol, orr;
        for (uint64_t k = 0; k < chunks && !too_wide; ++k) {
            const Sample* chunk = c.render_chunk();
            if (!chunk) continue;
            ol.clear(); orr.clear();
            rs.process(chunk->l.data(), chunk->r.data(), ol, orr);
            write(ol.data(), orr.data(), ol.size());
            c.fb.set_time_to_next_block();
        }
    } else {                                                                // state.rs:562-574
        for (uint64_t k = 0; k < chunks && !too_wide; ++k) {
            const Sample* chunk = c.render_chunk();
            if (!chunk) continue;
            write(chunk->l.data(), chunk->r.data(), chunk->len());
            c.fb.set_time_to_next_block();
        }
    }
    c.set_time(0);                                                          // state.rs:575
    if (frames_written) *frames_written = frames;
    if (too_wide) return c.fail(ERR_TOO_WIDE, "sample too wide for bit depth (hound TooWide; reference panics)");
    if (out_pcm) {
        if (cap_bytes < out.size()) return c.fail(ERR_CAPACITY, "output buffer too small");
        std::memcpy(out_pcm, out.data(), out.size());
    }
    return OK;
}

// Playback feed (ui_workflow.rs:195-210, stream_workflow.rs:86-100): `chunks` x { Graph::render;
// Sample::interleave (sample.rs:177-184); fb.set_time_to_next_block() }.  Nothing is rewound: the next
// call continues where this one stopped, with every vertex state as the last chunk left it.
int tdo_play(tdo_ctx* x, uint64_t chunks, float* out_interleaved, uint64_t cap_frames, uint64_t* frames_written) {
    Ctx& c = x->c;
    if (c.taps) c.clear_taps();
    uint64_t frames = 0;
    for (uint64_t k = 0; k < chunks; ++k) {
        const Sample* chunk = c.render_chunk();
        if (!chunk) return c.fail(ERR_NO_OUTPUT, "no output vertex");       // the reference unwraps the Option (ui_workflow.rs:203)
        const size_t len = chunk->len();
        if (out_interleaved) {
            if (frames + len > cap_frames) return c.fail(ERR_CAPACITY, "output buffer too small");
            for (size_t i = 0; i < len; ++i) { out_interleaved[2 * (frames + i)] = chunk->l[i]; out_interleaved[2 * (frames + i) + 1] = chunk->r[i]; }
        }
        frames += len;
        c.fb.set_time_to_next_block();
    }
    if (frames_written) *frames_written = frames;
    return OK;
}

int tdo_normalize_scan_many(tdo_ctx** xs, int n, const uint64_t* chunks) {
    for (int i = 0; i < n; ++i) tdo_normalize_scan(xs[i], chunks[i]);
    return OK;
}
int tdo_render_many(tdo_ctx** xs, int n, const uint64_t* chunks, const uint32_t* render_sr,
                    const uint32_t* bd, uint8_t** out, const uint64_t* cap, uint64_t* frames_written) {
    for (int i = 0; i < n; ++i) {
        int r = tdo_render(xs[i], chunks[i], render_sr[i], bd[i], out ? out[i] : nullptr,
                           cap ? cap[i] : 0, frames_written ? frames_written + i : nullptr);
        if (r < 0) return r;
    }
    return OK;
}

int tdo_get_normalize_state(tdo_ctx* x, float* mx, float* smx, int cap) {
    Ctx& c = x->c;
    int n = (int)c.vertices.size();
    for (int i = 0; i < n && i < cap; ++i) {
        bool nz = c.vertices[i].ext.kind == NORMALIZE;
        mx[i] = nz ? c.vertices[i].ext.max : -1.0f; smx[i] = nz ? c.vertices[i].ext.scan_max : -1.0f;
    }
    return n;
}

// ---- direct access to the scalar maths for the known-answer tests (adsr.rs:116-205) -------------
float tdo_adsr_eval(const float* conf, int n, int fn, float t, float aux) {
    AdsrConf c; if (!build_adsr_conf(conf, n, &c)) return NAN;
    switch (fn) {
        case 0: return apply_adsr(c, t);
        case 1: return apply_ads(c, t);
        case 2: return apply_r(c, t, aux);
        case 3: return apply_r_rt(c, t, aux);
        case 4: return max_vel(c);
        default: return NAN;
    }
}
float tdo_osc_eval(int which, float t, float hz, float z) {                 // synth.rs:22-34
    switch (which) { case 0: return square_sine_sample(t, hz, z); case 1: return topflat_sine_sample(t, hz, z);
                     default: return triangle_sample(t, hz); }
}
uint64_t tdo_set_length_chunks(uint32_t psr, float seconds, uint32_t bl) {  // state.rs:103-106
    return f32_as_usize(ceilf((float)psr * seconds / (float)bl));
}
// One-shot rubato restatement on raw arrays (Sample::resample, sample.rs:150-175).
int64_t tdo_resample_whole(const float* l, const float* r, uint64_t n, uint32_t from, uint32_t to,
                           float* out_l, float* out_r, uint64_t cap) {
    Sample s(0); s.l.assign(l, l + n); s.r.assign(r, r + n);
    Sample o = resample_whole(s, from, to);
    if (o.len() > cap) return ERR_CAPACITY;
    std::memcpy(out_l, o.l.data(), o.len() * 4); std::memcpy(out_r, o.r.data(), o.len() * 4);
    return (int64_t)o.len();
}

}  // extern "C"
