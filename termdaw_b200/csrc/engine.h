// engine.h — host-side structures of the device graph executor (internal to the library).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <deque>
#include <string>
#include <vector>
#include "tdg_types.h"

namespace tdg {

enum Kind { K_SUM, K_NORMALIZE, K_SAMPLE_LOOP, K_SAMPLE_MULTI, K_SAMPLE_LERP, K_DEBUG_SINE, K_SYNTH, K_SAMPSYN,
            K_ADSR, K_BANDPASS };

struct DevSample { float* l = nullptr; float* r = nullptr; uint64_t frames = 0; };
struct DevTable { float* data = nullptr; uint32_t n_frames = 1, n_phase = 1024; float fps = 1.0f; };

struct Floww {
    std::vector<float> t, note, vel;
    std::vector<uint64_t> fe;        // (t * sr as f32) as usize, floww.rs:75
};

struct MultiEntry { int64_t t; float vel; };                       // extensions.rs:29
// rec / xclk are planner bookkeeping of the pass being planned: the voice's VoiceRec, its explicit clock (-1: none)
struct DNote { float note, vel, hz; int rec; };                     // extensions.rs:42
// env_t is the voice's clock at the START of the pass being planned (carried from the pass before); inside a pass the
// clock is (clk_chunk, clk_i): -(clk_i/sr) + one addition of bl/sr per chunk since clk_chunk, or the explicit clock xclk
struct SNote { float note, vel, env_t, rel_t, hz; int rec; int xclk; float relv[3] = {0, 0, 0};
               uint32_t clk_chunk = 0, clk_i = 0, drop_chunk = 0xffffffffu; };   // extensions.rs:49,55; relv: apply_ads(conf_o, rel_t)
struct A3 { float t, vel, rel; };                                  // extensions.rs:67-68

struct OscConfH { float volume = 0, param = 0; AdsrConfD adsr{}; };

struct HVertex {
    int kind = K_SUM;
    float gain = 1, angle = 0, wet = 0;
    int sample = -1, floww = -1, table = -1;
    bool has_note = false; uint32_t note = 0; uint32_t lerp_len = 0;
    OscConfH sq, tf, tr;
    AdsrConfD adsr{};
    bool use_off = false, use_max = false;
    float lgamma = 0, hgamma = 0; bool pass = true;
    // ---- state carried between passes (mirrors VertexExt's mutable fields)
    uint64_t loop_t = 0;
    std::deque<MultiEntry> ts;
    uint64_t countdown = 0; MultiEntry primary{0, 0.0f}, ghost{0, 0.0f};
    std::vector<DNote> dnotes;
    std::vector<SNote> notes;
    uint64_t seq = 0;
    A3 aprimary{0, 0, 0}, aghost{0, 0, 0};
    bool bp_first = true;
    std::vector<int> edges;          // inputs, in connect() order (graph.rs:76)
    bool has_input() const { return kind == K_SUM || kind == K_NORMALIZE || kind == K_ADSR || kind == K_BANDPASS; }
};

enum KClass { KC_LOOP, KC_MULTI, KC_LERP, KC_DSINE, KC_SYNTH, KC_SAMPSYN, KC_FX_SUM, KC_FX_ADSR, KC_FX_NORM, KC_NORM_PREFIX, KC_NORM_SCALE,
              KC_BP_AGG, KC_BP_CARRY, KC_BP_APPLY, KC_RESAMPLE, KC_QUANT, KC_XCHG, KC_RANGES, KC_BP_FUSED, KC_COUNT };
struct KStat { uint64_t launches = 0, bytes = 0; double ms = 0; };

struct PassStats {
    double device_ms = 0, plan_ms = 0, h2d_ms = 0, d2h_ms = 0;
    uint64_t h2d_bytes = 0, d2h_bytes = 0, kernel_launches = 0, algorithmic_bytes = 0, frames = 0, voice_frames = 0;
    uint64_t reused_vertices = 0, reused_bytes = 0;     // render pass: vertices (and their algorithmic bytes) taken from the scan pass
    uint64_t inlined_vertices = 0, inlined_bytes = 0;   // Sum vertices evaluated inside their consumer's kernel, and the intermediate-buffer bytes not moved
};

// What a normalize scan leaves behind for the render that follows it (DESIGN.md "pass-invariant cones"):
// the buffers of vertices whose whole input cone evaluates identically in both passes stay resident in
// the lane's keep arena, and the render pass reads them instead of evaluating the cone again.
struct KeepRecord {
    bool valid = false;
    int lane = -1;
    uint64_t epoch = 0;                      // the lane's keep-arena generation the buffers were written in
    uint64_t mut_gen = 0;                    // Ctx::mut_gen at the scan
    uint64_t total = 0, n_pad = 0;           // chunks / padded frames of the scan pass
    std::vector<char> inv;                   // per vertex: pass-invariant
    std::vector<int64_t> off;                // per vertex: keep-arena offset of a kept (frontier) buffer, else -1
    // per vertex: chunks at the head of the render that still have to be evaluated (UINT64_MAX: all of them).
    // sample_multi / sample_lerp differ between the two passes only while state leaked from the pass before is
    // audible: until the leaked hits have played out / until the first hit of the pass has faded in.  Their
    // scan buffer is kept whole and the render re-evaluates just that prefix into it.
    std::vector<uint64_t> hz;
};

struct Engine;

struct Ctx {
    Engine* eng = nullptr;
    uint32_t sr = 0, bl = 0;
    std::string err;
    std::vector<DevSample> samples;
    std::vector<std::vector<uint8_t>> resources;
    std::vector<DevTable> tables;            // one per SampSyn vertex built from a resource
    std::vector<Floww> flowws;
    std::vector<HVertex> verts;
    int output = -1;
    uint64_t t = 0;                          // Graph.t
    uint64_t fb_frame = 0;                   // FlowwBank.frame
    // per-vertex device state: [v*8 + {0:max, 1:scan_max, 2..5: lprevl,lprevr,hprevl,hprevr}]
    std::vector<float> vstate_host;
    float* vstate_dev = nullptr; size_t vstate_cap = 0;
    bool vstate_host_newer = true;
    bool taps = false;
    // layout of the last pass (for taps)
    std::vector<int64_t> last_buf_off;       // per vertex, -1 if not evaluated
    uint64_t last_n_frames = 0, last_n_pad = 0;
    uint64_t last_epoch = 0;
    int last_lane = 0;
    uint64_t batch_frames = 0, batch_voice_frames = 0;   // whole-call totals of the last batched pass
    PassStats stats;
    KStat kstats[KC_COUNT];                  // per kernel class, last pass (profiling on)
    uint64_t launches = 0, collectives = 0;
    void* nccl_comm = nullptr; int x_ranks = 1, x_rank = 0;      // time sharding over NCCL (tdg_comm_init_rank)
    uint64_t mut_gen = 0;                    // bumped by every call that edits the project or its transport state
    KeepRecord keep;
    int plays_in_flight = 0;                 // tdg_play_begin tickets of this context not yet collected by tdg_play_wait
    int fail(int code, const std::string& m) { err = m; return code; }
};

}  // namespace tdg
