// tdh.h — native host side above the C ABI (include/termdaw_b200.h).
//
// The reference's host is compiled Rust; cargo/rustc do not exist in this image, so the host half of
// `State` (state.rs:27-577) that a termdaw maintainer would keep in Rust is mirrored here in C++ with the
// reference's names, argument meaning and error behaviour (messages included):
//
//   tdh::Project      the Lua-visible globals of State::refresh (state.rs:103-157): every call only appends
//                     a tuple to a seed vector (state.rs:85-93)
//   tdh::run_script   the literal-call subset of Lua 5.3 the project scripts use (mlua, state.rs:83-164)
//   tdh::Assets       what the reference reads from disk: WAV samples (hound, sample.rs:231-274), MIDI
//                     flowws (floww crate, floww.rs:40-48), wavetable blobs (bufferbank.rs:26-40)
//   tdh::State        refresh / scan_exact / render / render_to_file through the C ABI, vertices added in
//                     the reference's FIXED kind order (state.rs:341-457), edges in script order (:459)
//   tdh::Config       project.toml: [settings] main / buffer_length / project_samplerate (config.rs:10-77)
//
// Nothing here computes audio: decoding/encoding files and driving the ABI only.  The ABI entry points
// are bound through TDH_FN so that the SAME host code can be linked against the CPU oracle (prefix tdo_,
// test infrastructure: tests/test_host_native.py) to check the host logic where no GPU exists.
#pragma once
#include <stdint.h>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#ifndef TDH_PREFIX
#define TDH_PREFIX tdg_
#endif
#define TDH_CAT_(a, b) a##b
#define TDH_CAT(a, b) TDH_CAT_(a, b)
#define TDH_FN(name) TDH_CAT(TDH_PREFIX, name)

extern "C" {
struct tdg_ctx;
int TDH_FN(ctx_create)(uint32_t project_sr, uint32_t buffer_len, int device, tdg_ctx** out);
void TDH_FN(ctx_destroy)(tdg_ctx* ctx);
const char* TDH_FN(last_error)(tdg_ctx* ctx);
int TDH_FN(sample_mode_from_string)(const char* s);
int TDH_FN(sample_add)(tdg_ctx* ctx, const float* linear, uint64_t n_values, uint32_t channels, uint32_t src_sr, int mode);
int TDH_FN(resource_add)(tdg_ctx* ctx, const uint8_t* bytes, uint64_t n);
int TDH_FN(sample_remove)(tdg_ctx* ctx, int sample);
int TDH_FN(resource_remove)(tdg_ctx* ctx, int resource);
int TDH_FN(floww_add)(tdg_ctx* ctx, const float* t_sec, const float* note, const float* vel, uint64_t n);
int TDH_FN(floww_reset)(tdg_ctx* ctx);
int TDH_FN(graph_reset)(tdg_ctx* ctx);
int TDH_FN(add_sum)(tdg_ctx* ctx, float gain, float angle);
int TDH_FN(add_normalize)(tdg_ctx* ctx, float gain, float angle);
int TDH_FN(add_sample_loop)(tdg_ctx* ctx, float gain, float angle, int sample);
int TDH_FN(add_sample_multi)(tdg_ctx* ctx, float gain, float angle, int sample, int floww, int note);
int TDH_FN(add_sample_lerp)(tdg_ctx* ctx, float gain, float angle, int sample, int floww, int note, int lerp_len);
int TDH_FN(add_debug_sine)(tdg_ctx* ctx, float gain, float angle, int floww);
int TDH_FN(add_synth)(tdg_ctx* ctx, float gain, float angle, int floww, float sq_vol, float sq_z, const float* sq_adsr, int sq_n,
                      float tf_vol, float tf_z, const float* tf_adsr, int tf_n, float tr_vol, const float* tr_adsr, int tr_n);
int TDH_FN(add_sampsyn)(tdg_ctx* ctx, float gain, float angle, int floww, const float* adsr, int n, int resource);
int TDH_FN(add_lv2fx)(tdg_ctx* ctx, float gain, float angle, float wet, int plugin);
int TDH_FN(add_adsr)(tdg_ctx* ctx, float gain, float angle, float wet, int floww, int use_off, int use_max, int note,
                     const float* conf, int n);
int TDH_FN(add_bandpass)(tdg_ctx* ctx, float gain, float angle, float wet, float lo, float hi, int pass);
int TDH_FN(connect)(tdg_ctx* ctx, int a, int b);
int TDH_FN(set_output)(tdg_ctx* ctx, int vertex);
int TDH_FN(check_graph)(tdg_ctx* ctx, int* unreachable, int cap, int* n_unreachable);
int TDH_FN(reset_normalize)(tdg_ctx* ctx);
float TDH_FN(get_normalize_value)(tdg_ctx* ctx, int vertex);
int TDH_FN(normalize_scan)(tdg_ctx* ctx, uint64_t chunks);
int TDH_FN(render)(tdg_ctx* ctx, uint64_t chunks, uint32_t render_sr, uint32_t bitdepth, uint8_t* out_pcm, uint64_t cap_bytes,
                   uint64_t* frames_written);
uint64_t TDH_FN(render_bound_bytes)(tdg_ctx* ctx, uint64_t chunks, uint32_t render_sr, uint32_t bitdepth);
int TDH_FN(play)(tdg_ctx* ctx, uint64_t chunks, float* out_interleaved, uint64_t cap_frames, uint64_t* frames_written);
int TDH_FN(floww_append)(tdg_ctx* ctx, int floww, const float* t_sec, const float* note, const float* vel, uint64_t n);
int TDH_FN(floww_set_time)(tdg_ctx* ctx, uint64_t frame);
uint64_t TDH_FN(get_time)(tdg_ctx* ctx);
int TDH_FN(normalize_scan_many)(tdg_ctx** ctxs, int n, const uint64_t* chunks);
int TDH_FN(render_many)(tdg_ctx** ctxs, int n, const uint64_t* chunks, const uint32_t* render_sr, const uint32_t* bitdepth,
                        uint8_t** out_pcm, const uint64_t* cap_bytes, uint64_t* frames_written);
#ifndef TDH_ORACLE          // CUDA library only (the oracle-linked test build defines TDH_ORACLE)
int tdg_normalize_render_many(tdg_ctx** ctxs, int n, const uint64_t* chunks, const uint32_t* render_sr, const uint32_t* bitdepth,
                              uint8_t** out_pcm, const uint64_t* cap_bytes, uint64_t* frames_written);
#endif
}

namespace tdh {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

// ---- seed vectors (state.rs:85-93) ---------------------------------------------------------------
struct SampleSeed { std::string name, path, mode; bool operator==(const SampleSeed& o) const { return name == o.name && path == o.path && mode == o.mode; } };
struct NamePath { std::string name, path; };
struct Lv2Param { std::string plugin, name; double value; };
struct BaseSeed { std::string name; double gain, angle; };
struct LoopSeed { std::string name; double gain, angle; std::string sample; };
struct MultiSeed { std::string name; double gain, angle; std::string sample, floww; int64_t note; };
struct LerpSeed { std::string name; double gain, angle; std::string sample, floww; int64_t note, lerp_len; };
struct SineSeed { std::string name; double gain, angle; std::string floww; };
struct SynthSeed { std::string name; double gain, angle; std::string floww; double sq_vel, sq_z; std::vector<double> sq_arr;
                   double tf_vel, tf_z; std::vector<double> tf_arr; double tr_vel; std::vector<double> tr_arr; };
struct SampsynSeed { std::string name; double gain, angle; std::string floww; std::vector<double> adsr; std::string resource; };
struct Lv2fxSeed { std::string name; double gain, angle, wet; std::string plugin; };
struct AdsrSeed { std::string name; double gain, angle, wet; std::string floww; bool use_off, use_max; int64_t note; std::vector<double> conf; };
struct BandpassSeed { std::string name; double gain, angle, wet, lo, hi; bool pass; };

// The Lua-visible API: method names and argument order are the registered globals (state.rs:103-157).
struct Project {
    bool has_length = false, has_render_sr = false, has_render_bd = false, has_output_file = false;
    double length_seconds = 0; int64_t render_samplerate = 0, render_bitdepth = 0;
    std::string output_file, output_vertex;
    std::vector<SampleSeed> samples; std::vector<NamePath> resources, midis, lv2plugins; std::vector<std::string> streams;
    std::vector<Lv2Param> lv2params;
    std::vector<BaseSeed> sums, norms; std::vector<LoopSeed> sampleloops; std::vector<MultiSeed> samplemultis;
    std::vector<LerpSeed> samplelerps; std::vector<SineSeed> debugsines; std::vector<SynthSeed> synths;
    std::vector<SampsynSeed> sampsyns; std::vector<Lv2fxSeed> lv2fxs; std::vector<AdsrSeed> adsrs;
    std::vector<BandpassSeed> bandpasses; std::vector<std::pair<std::string, std::string>> edges;

    void set_length(double seconds) { has_length = true; length_seconds = seconds; }                 // state.rs:103-106
    void set_render_samplerate(int64_t sr) { has_render_sr = true; render_samplerate = sr; }
    void set_render_bitdepth(int64_t bd) { has_render_bd = true; render_bitdepth = bd; }
    void set_output_file(const std::string& f) { has_output_file = true; output_file = f; }
    void load_sample(const std::string& n, const std::string& p, const std::string& mode = "") { samples.push_back({n, p, mode}); }
    void load_resource(const std::string& n, const std::string& p) { resources.push_back({n, p}); }
    void load_midi_floww(const std::string& n, const std::string& p) { midis.push_back({n, p}); }
    void declare_stream(const std::string& n) { streams.push_back(n); }
    void load_lv2(const std::string& n, const std::string& uri) { lv2plugins.push_back({n, uri}); }
    void parameter(const std::string& plugin, const std::string& n, double v) { lv2params.push_back({plugin, n, v}); }
    void add_sum(const std::string& n, double g, double a) { sums.push_back({n, g, a}); }
    void add_normalize(const std::string& n, double g, double a) { norms.push_back({n, g, a}); }
    void add_sampleloop(const std::string& n, double g, double a, const std::string& s) { sampleloops.push_back({n, g, a, s}); }
    void add_sample_multi(const std::string& n, double g, double a, const std::string& s, const std::string& f, int64_t note) {
        samplemultis.push_back({n, g, a, s, f, note}); }
    void add_sample_lerp(const std::string& n, double g, double a, const std::string& s, const std::string& f, int64_t note, int64_t ll) {
        samplelerps.push_back({n, g, a, s, f, note, ll}); }
    void add_debug_sine(const std::string& n, double g, double a, const std::string& f) { debugsines.push_back({n, g, a, f}); }
    void add_synth(const std::string& n, double g, double a, const std::string& f, double sqv, double sqz, std::vector<double> sqa,
                   double tfv, double tfz, std::vector<double> tfa, double trv, std::vector<double> tra) {
        synths.push_back({n, g, a, f, sqv, sqz, std::move(sqa), tfv, tfz, std::move(tfa), trv, std::move(tra)}); }
    void add_sampsyn(const std::string& n, double g, double a, const std::string& f, std::vector<double> adsr, const std::string& res) {
        sampsyns.push_back({n, g, a, f, std::move(adsr), res}); }
    void add_lv2fx(const std::string& n, double g, double a, double wet, const std::string& plugin) { lv2fxs.push_back({n, g, a, wet, plugin}); }
    void add_adsr(const std::string& n, double g, double a, double wet, const std::string& f, bool use_off, bool use_max, int64_t note,
                  std::vector<double> conf) { adsrs.push_back({n, g, a, wet, f, use_off, use_max, note, std::move(conf)}); }
    void add_bandpass(const std::string& n, double g, double a, double wet, double lo, double hi, bool pass) {
        bandpasses.push_back({n, g, a, wet, lo, hi, pass}); }
    void connect(const std::string& a, const std::string& b) { edges.push_back({a, b}); }
    void set_output(const std::string& o) { output_vertex = o; }
    std::string to_json() const;             // the seed vectors, for cross-checks against the Python mirror
};

// ---- Lua subset (tdh_lua.cpp) --------------------------------------------------------------------
struct LuaSubsetError : Error { using Error::Error; };
void run_script(const std::string& text, Project& p);
void run_file(const std::string& path, Project& p);

// ---- assets (tdh_assets.cpp) ---------------------------------------------------------------------
struct DecodedWav { std::vector<float> linear; uint32_t channels = 0, sample_rate = 0, bits = 16; };
struct FlowwEvents { std::vector<float> t, note, vel; };
struct WavError : Error { using Error::Error; };
struct MidiError : Error { using Error::Error; };
std::vector<uint8_t> read_file(const std::string& path);                         // throws Error
DecodedWav decode_wav(const std::vector<uint8_t>& buf);                           // hound semantics, sample.rs:231-274
std::vector<uint8_t> wav_header(uint64_t data_bytes, uint32_t sample_rate, uint32_t bitdepth, uint32_t channels = 2);
void write_wav(const std::string& path, const std::vector<uint8_t>& pcm, uint32_t sample_rate, uint32_t bitdepth);
FlowwEvents parse_midi(const std::vector<uint8_t>& data);                         // DEFINED in termdaw_b200/midi.py (parity unpinned)
struct Assets {                                                                   // disk, with a "<file>.floww.f32" side door
    virtual ~Assets() {}
    virtual DecodedWav wav(const std::string& path);
    virtual FlowwEvents floww(const std::string& path);
    virtual std::vector<uint8_t> blob(const std::string& path);
};

// ---- project.toml (config.rs:10-77) ---------------------------------------------------------------
struct Config { std::string main; uint32_t buffer_length = 1024, project_samplerate = 44100; std::string workflow; };
Config read_config(const std::string& path);

// ---- State (tdh_state.cpp) -------------------------------------------------------------------------
uint64_t set_length_chunks(uint32_t psr, double seconds, uint32_t bl);           // state.rs:104, in f32
struct State {
    uint32_t psr, bl;
    tdg_ctx* ctx = nullptr;
    Assets* assets;
    bool loaded = false;
    uint64_t cs = 0;
    uint32_t render_sr = 48000, bd = 16;                                          // main.rs:75-98
    std::string output_file = "outp.wav", output_vertex;
    std::vector<SampleSeed> cur_samples;
    std::vector<NamePath> cur_resources;
    std::map<std::string, int> sample_index, resource_index, floww_index, name_map;
    std::vector<std::string> names, messages;

    State(uint32_t project_samplerate, uint32_t buffer_length, int device, Assets* assets);   // throws Error without a device
    ~State();
    State(const State&) = delete;
    State& operator=(const State&) = delete;
    bool refresh(const Project& p);                                               // State::refresh, state.rs:50-471
    void scan_exact();                                                            // state.rs:473-475
    std::vector<uint8_t> render();                                                // state.rs:477-577: the data-chunk bytes
    std::string render_to_file(const std::string& path = "");
    // playback / stream feed (ui_workflow.rs:195-210, stream_workflow.rs:63-100): `chunks` more chunks from the
    // current transport position, interleaved (L, R) f32, nothing rewound; feed() appends events to a declared
    // stream and re-seats the floww cursor at the graph time, as the reference's Feed handler does
    std::vector<float> play(uint64_t chunks);
    bool feed(const std::string& stream, const FlowwEvents& events);
    std::vector<std::pair<std::string, float>> normalization_values();            // graph.rs:213-220
    // The batch of independent projects (BASELINE configs[4]): "normalize" then "render" of every state in one
    // call — tdg_normalize_render_many, pipelined over project groups — or render only (`normalize` false).
    // Returns the data-chunk bytes per state, identical to scan_exact() + render() on each.
    static std::vector<std::vector<uint8_t>> render_batch(const std::vector<State*>& states, bool normalize);
    void say(const std::string& m) { messages.push_back(m); }
    int chk(int rc);                                                              // rc < 0 -> throws Error(last_error)
};

}  // namespace tdh
