// tdg_types.h — plain structs shared by the host planner and the sm_100a kernels.
//
// A "pass" is one whole-timeline evaluation of a project's graph (a normalize scan or a render):
// `n_chunks` chunks of `bl` frames.  Every reachable vertex owns one planar stereo fp32 buffer
// [L: n_pad][R: n_pad] resident in HBM for the pass; a "task" is one vertex evaluated over the whole
// pass by one kernel class, and a launch covers all tasks of one class at one DAG level for every
// project in the batch (grid.y = task, grid.x = chunk / tile).
//
// Offsets named *_off are byte offsets into the per-pass "plan blob" (tables the host planner
// derives from floww events), which the kernels receive as a base pointer.
#pragma once
#include <stdint.h>

namespace tdg {

struct PanGain {            // Sample::apply_angle / apply_gain (sample.rs:97-114), constants from the host
    float l_amp, r_amp, gain;
    uint32_t flags;         // bit0: apply angle (|angle| >= 0.001), bit1: apply gain (|gain-1| >= 0.001)
};

struct AdsrConfD {          // adsr.rs:2-12
    float std_vel, attack_sec, attack_vel, decay_sec, decay_vel, sustain_sec, sustain_vel, release_sec,
        release_vel;
};

struct AdsrRcp { float attack, decay, sustain, release; };   // 1/x of the four durations (host division)

// Device form of an envelope: every loop-invariant f32 value of adsr.rs:46-92 evaluated ONCE on the
// host with the reference's own operations (thresholds a, a+d, (a+d)+s; lerp deltas b-a; IEEE
// reciprocals for the exact-division shortcut of tdg_math.cuh).  `exotic` marks a duration that is
// neither 0 nor in [1e-6, 1e6]: launches containing one use the plain-division kernel variant.
struct AdsrK {
    float a, ad, ads, dsec, ssec, rsec;
    float ra, rd, rs, rr;
    float v0, v1, v2, v3, v4;
    float d01, d12, d23, d34;
    uint32_t exotic;
    uint32_t simple;        // every duration in [1e-6, 1e6] and every level > -0.5: quotients are finite and the
                            // -1000 'past the sustain' sentinel of adsr.rs:59 cannot be hit by a lerp result, so the
                            // per-frame inf / sentinel tests are dropped (same values, fewer instructions)
};

struct TaskHdr {
    float* out_l;
    float* out_r;
    uint64_t n_frames;      // n_chunks * bl
    uint32_t n_chunks, bl;
    PanGain pg;
};

// ---- event records: the plan the host uploads is O(events), the per-chunk tables are built on the device ----
// Every voice / envelope clock of the reference is a run of f32 additions of d = bl as f32 / sr as f32, one per
// chunk, starting from -(i as f32 / sr as f32) at the frame offset i of the event that (re)started it
// (extensions.rs:481-494,525-527 voices; :610-618,640-650 adsr slots).  Its value n chunks later depends on
// (i, n) only: clock[n][i], ONE table per (sr, bl) for every vertex of every project on the device, built there
// by k_env_table with the same sequential additions.  A clock source word names where a record's clock comes
// from: bit 31 clear -> column i of the table; bit 31 set -> index into the task's explicit per-chunk array in
// the plan blob (state carried in from an earlier pass: the scan before a render, tdg_play).
static const uint32_t kEnvExplicit = 0x80000000u;
static const uint32_t kNever = 0xffffffffu;
struct VoiceRec {           // one voice of a Synth / SampSyn / DebugSine vertex over one pass, in Vec (push) order
    float hz, vel;
    uint32_t on_chunk, on_i;        // first sounding frame: pass chunk, frame offset
    uint32_t off_chunk, off_i;      // first frame of its release phase; off_chunk == kNever: not released in this pass
    uint32_t end_chunk, end_i;      // one past its last frame: chunk, offset in [0, bl]
    float rel_t, relv0, relv1, relv2;   // release phase: rel_t and apply_ads(conf_o, rel_t) per oscillator (adsr.rs:89-92)
    uint32_t envh, envr;            // clock source of the held / the release phase
    uint32_t pm_end, sm_on;         // prefix max of end_chunk over records [0, v], suffix min of on_chunk over [v, n):
                                    // both monotone, so the candidates of a chunk are one binary-searched index range
};
struct AdsrRec {            // Adsr vertex state from pass frame `start` on (extensions.rs:61-69)
    uint64_t start;                 // chunk * bl + frame offset
    uint32_t flags, pad;            // bit0: this single frame is passed through un-enveloped (extensions.rs:633-635)
    uint32_t p_src, p_chunk;        // primary: clock source and the pass chunk its clock (re)started in
    float pv, pr;                   //          vel, release_val
    uint32_t g_src, g_chunk;        // ghost
    float gv, gr;
};
struct MultiHit { int64_t start; float vel, pad; };                       // pass-relative first frame
struct LerpHit { int64_t frame, pos_origin, cd_origin; float vel, pad; }; // see plan_sample_lerp

// per-chunk index ranges into the records above, written by k_chunk_ranges (one thread per chunk, binary searches)
struct alignas(8) Range2 { uint32_t x, y; };
struct alignas(16) F4 { float x, y, z, w; };        // float4 without the CUDA headers (this file is plain C++ too)
enum RangeKind : uint32_t { RK_VOICE = 0, RK_ADSR = 1, RK_MULTI = 2, RK_LERP = 3 };
struct RangeTask {
    uint32_t kind, n_chunks, chunk0, bl;
    uint64_t recs_off;              // blob: VoiceRec[] / AdsrRec[] / MultiHit[] / LerpHit[]
    uint32_t n_recs, pad;
    int64_t len;                    // RK_MULTI: sample length
    Range2* out;                    // [n_chunks] x = lo, y = hi (RK_ADSR: first, count; RK_LERP: x = base)
    // RK_VOICE / RK_ADSR: the per-chunk segment table k_voice / k_fx<adsr> walk, built on the device from the records
    // (k_chunk_ranges counts, k_seg_scan turns counts into offsets, k_seg_fill writes the segments)
    uint32_t* seg_ptr;              // [n_chunks + 1]
    F4* segs;                       // RK_VOICE [capacity][2]: (hz, vel, env_t, rel_t), (ia, ib, relv0, relv1); RK_ADSR: AdsrSeg[capacity]
    float* relv2;                   // [capacity]
    const float* clock;             // clock table of this (sr, bl)
    uint64_t xenv_off;              // blob: explicit clocks
    uint32_t debug_sine, pad2;
};

// ---- per-chunk forms (device-built) ---------------------------------------------------------------------
struct AdsrSeg {            // Adsr vertex state valid from frame offset `ia` of its chunk (extensions.rs:61-69)
    uint32_t ia, flags;     // flags bit0: this single frame is passed through un-enveloped (extensions.rs:633-635)
    float pt, pv, pr, gt, gv, gr;   // primary (t, vel, release_val), ghost (t, vel, release_val)
};
struct VoiceSeg {           // one voice over frames [ia, ib) of one chunk, in voice (push) order
    float hz, vel, env_t, rel_t;
    uint32_t ia, ib;
    float relv0, relv1;     // released voices: apply_ads(conf, rel_t) of adsr.rs:89-92 for oscillator 0 / 1 (constant per
                            // voice, evaluated once by the host planner with the reference's own operations)
};

// ---- tasks ------------------------------------------------------------------------------------
static const int kInlineInputs = 8;     // the reference's graphs rarely fan in wider; wider vertices read the blob table
enum FxMode : uint32_t { FX_SUM = 0, FX_ADSR = 1, FX_NORM_PEAK = 2 };

// One input EDGE of a k_fx vertex when some of its inputs are inlined Sum vertices (engine.cu "inlined sums"): `count`
// pointer pairs starting at pair `first`; nested = 1: they are the inputs of a Sum vertex whose output is that edge's
// value — summed in order from +0.0, then that vertex's pan and gain (the operations of k_fx<FX_SUM>, in registers).
struct FxGroup { uint32_t first, count, nested, pad; PanGain pg; };
static const int kFxPairsStaged = 32;       // ... and that many input pointer pairs
static const int kFxGroupsStaged = 16;      // group records of a task staged in shared memory by the grouped kernel (more: read from the plan blob)

struct FxTask {             // sum_inputs (+ adsr_gen | + per-chunk peak) + pan + gain, one kernel
    TaskHdr h;
    uint32_t mode, n_in;
    uint64_t in_off;        // blob: n_in pairs of (const float* l, const float* r)
    const float* ins[2 * kInlineInputs];   // the same pairs inline when n_in <= kInlineInputs: one dependent load less per block
    AdsrK conf;
    float wet, fsr, rsr;    // fsr = sr as f32, rsr = 1/fsr
    uint32_t use_off, use_max, sr;
    // FX_ADSR: per-chunk envelope states in device scratch, built there from the AdsrRec records (RangeTask RK_ADSR:
    // k_chunk_ranges counts, k_seg_scan offsets, k_seg_fill writes them with their clocks looked up)
    const uint32_t* seg_ptr; // [n_chunks + 1]
    const AdsrSeg* segs;
    float* peaks;           // FX_NORM_PEAK: [n_chunks] max |x| over L and R of the summed chunk
    uint32_t no_store, n_grp; // no_store: FX_NORM_PEAK of the output vertex in a scan pass: only the peaks are needed
    uint64_t grp_off;       // blob: n_grp FxGroup records (n_grp = number of edges) when some inputs are inlined Sum vertices; n_grp = 0: plain pairs
};

// Time sharding: a pass split into `x_ranks` consecutive segments.  Every DAG level owns
// an exchange buffer of x_ranks slots of `x_stride` floats; a task's own values live at
// x_buf[x_rank * x_stride + x_index ...].  The slots of the other ranks arrive by ncclAllGather over
// NVLink (one process per GPU) or are written by sibling segments of the same launch ("virtual
// ranks", used to test the hand-off arithmetic on one GPU).
struct Xchg { float* buf; uint32_t stride, index, ranks, rank; };

struct NormPrefixTask {     // normalize_gen's max bookkeeping (extensions.rs:321-329), one block per task
    const float* peaks;
    float* runmax;          // [n_chunks] the `max` each chunk is scaled by
    float* state;           // device-resident (max, scan_max) of this vertex
    uint32_t n_chunks, is_scan;
    Xchg x;                 // x.ranks > 1: slot = { segment peak }
    const float* reduced;   // scan pass over NCCL: global peak from ncclAllReduce(max), else null
};
struct NormScaleTask {      // buf *= 1/max, then pan + gain, in place
    TaskHdr h;
    const float* runmax;
};

struct LoopTask { TaskHdr h; const float* sl; const float* sr; uint64_t len, t0; };
struct MultiTask {
    TaskHdr h; const float* sl; const float* sr; uint64_t len; uint64_t hits_off; uint32_t n_hits, pad;
    const Range2* ranges;    // device scratch [n_chunks] = [lo, hi) of the hits that can sound in the chunk (k_chunk_ranges:
                            // one thread per chunk does the two binary searches once, not every block at its head)
};
struct LerpTask {
    TaskHdr h; const float* sl; const float* sr; uint64_t len; uint64_t hits_off;
    uint32_t n_hits, lerp_len;
    const Range2* ranges;    // device scratch [n_chunks]: x = index of the last hit at or before the chunk's first frame
};

enum VoiceKind : uint32_t { VOICE_DEBUG_SINE = 0, VOICE_SYNTH = 1, VOICE_SAMPSYN = 2 };
struct VoiceTask {
    TaskHdr h;
    uint32_t kind, sr;
    float fsr, rsr;         // sr as f32 and its IEEE reciprocal
    uint64_t t0;            // graph time of pass frame 0 (time = (t0 + p) as f32 / sr as f32)
    // per-chunk segment table in device scratch, built there from the VoiceRec records (RangeTask)
    const uint32_t* seg_ptr; // [n_chunks + 1]
    const F4* segs;          // [][2]: (hz, vel, env_t, rel_t), (ia, ib, relv0, relv1) in push order
    const float* relv2;      // parallel: the release start value of oscillator 2 (Synth)
    // Synth (extensions.rs:44-50); shape constants of synth.rs:22-29 evaluated on the host:
    // sq_iz = 1/z, tf_add = (1-z)/2, tf_mul = 2/(1+z)
    float vol[3];           // square, topflat, triangle volumes
    uint32_t env12_same;    // topflat and triangle share one ADSR table (as in the reference's project.lua): evaluate once
    float sq_z, sq_iz, tf_z, tf_add, tf_mul, amp_mul;
    AdsrK env[3];           // square, topflat, triangle; SampSyn (extensions.rs:51-56) uses env[0]
    // SampSyn wavetable [n_frames][n_phase]
    const float* table;
    uint32_t n_frames, n_phase;
    float fps;
    uint32_t table_floats;
};

struct BpTask {             // band_pass_gen (extensions.rs:654-689) as a chunked linear-recurrence scan
    TaskHdr h;
    uint32_t n_in, pass;
    uint64_t in_off;
    const float* ins[2 * kInlineInputs];   // inline copy of the input pairs when n_in <= kInlineInputs
    float lgamma, hgamma;
    uint32_t first;         // seed the four states with frame 0 of this pass (extensions.rs:664-670)
    uint32_t n_tiles;
    float* state;           // device-resident lprevl, lprevr, hprevl, hprevr (carried between passes)
    float* agg;             // [n_tiles][8] end state of each tile started from zero + decays
    float* carry;           // [n_tiles][4] true state entering each tile (fused kernel: state LEAVING a group's last tile)
    Xchg x;                 // x.ranks > 1: slot = { al, ah, b0..b3 (segment map), init0..3 (rank 0 only) }
    uint32_t* sync;         // fused single-pass kernel: [0] tile ticket, [1 + t] publication flag of tile t (zeroed per pass)
};
static const int kBpXchgFloats = 10;

struct QuantTask {          // write_16s / write_32s (state.rs:517-532) + hound's byte layout
    const float* in_l;
    const float* in_r;
    uint8_t* out;
    uint64_t n;
    uint32_t bd;
    float amp;
    uint32_t* err;          // set non-zero when an 8/24-bit value is too wide
    uint32_t bl;            // frames per chunk (fused normalize only)
    // Output vertex = Normalize (the usual termdaw project): its scale pass is fused in here instead of
    // a separate read-modify-write of the whole buffer: x * (1/max_of_chunk), pan, gain, then quantise
    // (extensions.rs:328,262-263 then state.rs:517-532 — same operations, same order, in registers).
    const float* runmax;    // null: input is already final
    uint64_t frame0;        // pass frame of input frame 0: always a whole number of chunks
    PanGain pg;
    uint32_t chunk0, span;  // frame0 / bl; frames per block: bl when fused (a block = one chunk = one max), else 1024
};

struct ResampleTask {       // rubato SincFixedIn restated: one output frame per thread
    const float* in_l;      // stream; positions < 0 read as 0 (initial zero buffer)
    const float* in_r;
    float* out_l;
    float* out_r;
    const float* sincs;     // [factor][sinc_len]
    uint64_t n_in;
    uint64_t chunk_off;     // blob: ResampleChunk[n_chunks]
    uint32_t n_chunks, chunk;
    double t_ratio;
    // time-sharded render: frames in front of in_*[0].  `n_before` of them sit in the same buffer (warm-up
    // chunks of this segment); otherwise the previous segment's last kResampleHalo frames arrive through
    // the exchange buffer (halo_l/halo_r, frame -1 is halo[kResampleHalo - 1]); null => zeros (pass start)
    const float* halo_l;
    const float* halo_r;
    uint64_t n_before;
    double* idx;            // [total outputs] SincFixedIn's f64 position of every output: k_resample_index accumulates it
                            // chunk by chunk with the reference's sequential `idx += t_ratio` (bit-exact positions)
};
static const int kResampleHalo = 512;   // 2 * sinc_len: the history SincFixedIn keeps in front of a chunk
struct ResampleChunk { double last_index; uint64_t out_off; uint32_t n_out, pad; };

static const int kBpFramesPerThread = 8;
static const int kBpThreads = 256;
static const int kBpTile = kBpFramesPerThread * kBpThreads;     // 2048 frames
#ifndef TDG_BP_FUSED_THREADS
#define TDG_BP_FUSED_THREADS 128     // measured per bench step: 256 threads (3 blocks of 72 KB per SM) 6.68 ms, 128 threads (6 blocks of 36 KB) 6.18 ms
#endif
static const int kBpFusedThreads = TDG_BP_FUSED_THREADS;        // fused kernel: threads per block,
static const int kBpFpt = 32;                                   //   32 consecutive frames per thread,
static const int kBpStrip = kBpFusedThreads * kBpFpt;           //   a block owns a strip of 4096 frames
static const int kBpStripTiles = kBpStrip / kBpTile;            //   (2 tiles of the three-launch scan),
static const int kBpRow = kBpFpt + 4;                           //   staged in shared-memory rows padded to 36 words (conflict-free float4 access both ways)
static const int kBpStripSmem = 2 * kBpFusedThreads * kBpRow * 4;   //   36 KB per block
static const int kBpGroup = 32;                                 // fused kernel: strips per look-back group (one warp lane each)
static const uint32_t kBpFusedMaxTiles = 65535u * kBpStripTiles; // grid limit; the serial group chain (n_strips / 32 steps of ~2.5 us) still beats the second read of the
                                                                // three-launch scan on an hour-long pass (configs[3]: 15.5 vs 17.5 ms per step)

}  // namespace tdg
