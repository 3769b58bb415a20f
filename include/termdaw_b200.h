/*
 * termdaw_b200.h — C ABI of the B200-native offline renderer for termdaw's audio graph.
 *
 * This is the drop-in boundary.  termdaw (Rust) has no FFI seam of its own for
 * this path; the seam is cut exactly where `State` touches `Graph` / the banks, so every entry
 * point below names the reference call it replaces (paths relative to the reference repo).  A
 * Rust `extern "C"` block binding these symbols is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; the context owns all device memory; the caller owns every host array and
 *     no pointer is retained past a call (mirrors `graph.rs:189`, where `Graph` lends `&Sample`).
 *   - functions returning `int` return >= 0 on success (an index where one is produced) and a
 *     negative `TDG_ERR_*` on failure; `tdg_last_error(ctx)` gives the message.  Nothing aborts.
 *   - one context == one termdaw project (`State`: sample bank + floww bank + buffer bank + graph),
 *     single-threaded like the reference (all graph access is from the main thread,
 *     `ui_workflow.rs:17-21`), bound to one CUDA device.
 *   - there is NO CPU fallback: every call that renders fails with TDG_ERR_CUDA when no sm_100
 *     device is usable.
 *
 * The CPU oracle under oracle/ exports the same shape with the prefix `tdo_` (test infrastructure).
 */
#ifndef TERMDAW_B200_H
#define TERMDAW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tdg_ctx tdg_ctx;

enum {
    TDG_OK = 0,
    TDG_ERR_INVALID = -1,            /* bad argument / unknown index                                  */
    TDG_ERR_CUDA = -2,               /* CUDA runtime failure or no usable device                      */
    TDG_ERR_UNSUPPORTED_VERTEX = -3, /* add_lv2fx: arbitrary native plugin code is outside the path   */
    TDG_ERR_ADSR_CONF = -4,          /* adsr array is not 0, 6 or 9 floats (reference panics, state.rs:393) */
    TDG_ERR_SAMPLE = -5,             /* Sample::from / SampleBank::add error paths (sample.rs:38-77,241-254) */
    TDG_ERR_BITDEPTH = -6,           /* bitdepth not in {8,16,24,32} (state.rs:495-501)               */
    TDG_ERR_NO_OUTPUT = -7,          /* no output vertex set (graph.rs:186-192 returns None)          */
    TDG_ERR_TOO_WIDE = -8,           /* 8/24-bit sample out of range: hound TooWide -> reference panics (state.rs:521) */
    TDG_ERR_CAPACITY = -9,           /* caller's output array too small                               */
    TDG_ERR_UNSORTED = -10           /* floww events not sorted by frame (GPU path requires sorted)   */
};

/* SampleLoadMethod (sample.rs:200-215).  tdg_sample_mode_from_string mirrors `SampleLoadMethod::from`. */
enum {
    TDG_MODE_STEREO = 0, TDG_MODE_LEFT = 1, TDG_MODE_RIGHT = 2, TDG_MODE_LOUDEST = 3,
    TDG_MODE_NORM = 4 /* "normalize-seperate" */, TDG_MODE_MIX = 5 /* "mix-down" */
};
int tdg_sample_mode_from_string(const char* s);

/* ---- context ------------------------------------------------------------------------------------
 * Replaces Graph::new(max_buffer_len, sr) (graph.rs:25), SampleBank::new(sr) (sample.rs:213),
 * FlowwBank::new(sr, bl) (floww.rs:17), BufferBank::new() (bufferbank.rs:18) as built in main.rs:77-83. */
int tdg_ctx_create(uint32_t project_sr, uint32_t buffer_len, int device, tdg_ctx** out);
void tdg_ctx_destroy(tdg_ctx* ctx);
const char* tdg_last_error(tdg_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches claim). */
uint64_t tdg_kernel_launches(tdg_ctx* ctx);

/* ---- banks --------------------------------------------------------------------------------------
 * Numeric part of SampleBank::add (sample.rs:262-310): `linear` is the decoded WAV sample stream
 * (integer PCM cast `as f32` UNSCALED, or float samples), channel-interleaved, `n_values` values.
 * The device de-interleaves, applies the load mode, peak-normalises and sinc-resamples to the
 * project rate.  Returns the sample index (SampleBank::get_index, sample.rs:338). */
int tdg_sample_add(tdg_ctx* ctx, const float* linear, uint64_t n_values, uint32_t channels,
                   uint32_t src_sr, int mode);
/* SampleBank::mark_dead + refresh (sample.rs:316-336): drop a sample that left the script or changed; its
 * device memory is freed, the other indices stay valid.  tdg_resource_remove: the same for BufferBank blobs
 * (state.rs:222-237). */
int tdg_sample_remove(tdg_ctx* ctx, int sample);
int tdg_resource_remove(tdg_ctx* ctx, int resource);
/* Frames of a loaded sample at project rate (Sample::len, sample.rs:79). */
int64_t tdg_sample_frames(tdg_ctx* ctx, int sample);
/* Debug tap: copy a loaded sample (after normalise + resample) back to the host. */
int tdg_sample_read(tdg_ctx* ctx, int sample, float* l, float* r, uint64_t cap_frames);

/* BufferBank::add (bufferbank.rs:26) + parse_wavetable_from_buffer at graph build (state.rs:415). */
int tdg_resource_add(tdg_ctx* ctx, const uint8_t* bytes, uint64_t n);

/* FlowwBank::declare_floww / add_floww / declare_stream (floww.rs:32-53): one floww = time-sorted
 * events (t seconds, note, vel); vel <= 0.001 is a note-off (floww.rs:114,134). n may be 0.
 * PRECONDITION: event frames (t * sr) must be non-decreasing, else TDG_ERR_UNSORTED.  What the reference's
 * crate hands FlowwBank (a MIDI file read in tick order, stream packets in arrival order) is sorted; its
 * cursor (floww.rs:70-96) walks forward only, so an out-of-order event would merely be skipped or replayed
 * there — a degradation this ABI refuses to reproduce.  Callers with late packets sort before the call. */
int tdg_floww_add(tdg_ctx* ctx, const float* t_sec, const float* note, const float* vel, uint64_t n);
/* FlowwBank::append_streams (floww.rs:55-57; the Feed handler of stream_workflow.rs:63-71): more events for
 * a declared stream while the transport runs.  Events must not start before the stream's last event
 * (TDG_ERR_UNSORTED).  Follow with tdg_floww_set_time(ctx, tdg_get_time(ctx)) as the reference does. */
int tdg_floww_append(tdg_ctx* ctx, int floww, const float* t_sec, const float* note, const float* vel,
                     uint64_t n);
/* FlowwBank::reset (floww.rs:21) — State::refresh reloads every floww (state.rs:240). */
int tdg_floww_reset(tdg_ctx* ctx);

/* ---- graph construction -------------------------------------------------------------------------
 * Graph::reset (graph.rs:39).  Vertices get indices in call order (graph.rs:49-56). */
int tdg_graph_reset(tdg_ctx* ctx);
/* State::refresh on an UNCHANGED project: the reference drops the graph and re-adds every vertex
 * (state.rs:328-457), so all per-vertex state (loop counters, hit deques, voices, envelope slots,
 * filter state, normalize maxima) returns to its constructor value (extensions.rs:83-194) and the
 * transport to 0.  Same result as tdg_graph_reset + re-adding the same vertices and edges, without
 * the calls; device-resident samples, resources and flowws are kept (state.rs:201-219). */
int tdg_reset_vertex_state(tdg_ctx* ctx);
/* Graph::add(Vertex::new(bl, gain, angle, wet, VertexExt::…)) — graph.rs:49, ctors extensions.rs:83-194,
 * argument clamps of state.rs:341-457 are applied by the CALLER's front-end except where noted. */
int tdg_add_sum(tdg_ctx* ctx, float gain, float angle);                                   /* state.rs:341 */
int tdg_add_normalize(tdg_ctx* ctx, float gain, float angle);                             /* state.rs:344 */
int tdg_add_sample_loop(tdg_ctx* ctx, float gain, float angle, int sample);               /* state.rs:349 */
/* note < 0 => None (state.rs:358) */
int tdg_add_sample_multi(tdg_ctx* ctx, float gain, float angle, int sample, int floww, int note);
/* lerp_len is clamped to >= 0 (state.rs:370) */
int tdg_add_sample_lerp(tdg_ctx* ctx, float gain, float angle, int sample, int floww, int note,
                        int lerp_len);
int tdg_add_debug_sine(tdg_ctx* ctx, float gain, float angle, int floww);                 /* state.rs:378 */
/* square z is clamped to >= 0.0001 (state.rs:400); adsr arrays: 0, 6 or 9 floats (adsr.rs:94-114) */
int tdg_add_synth(tdg_ctx* ctx, float gain, float angle, int floww,
                  float sq_vol, float sq_z, const float* sq_adsr, int sq_n,
                  float tf_vol, float tf_z, const float* tf_adsr, int tf_n,
                  float tr_vol, const float* tr_adsr, int tr_n);
int tdg_add_sampsyn(tdg_ctx* ctx, float gain, float angle, int floww, const float* adsr, int n,
                    int resource);                                                        /* state.rs:406 */
/* Always TDG_ERR_UNSUPPORTED_VERTEX: the GPU renderer rejects graphs with LV2 fx vertices. */
int tdg_add_lv2fx(tdg_ctx* ctx, float gain, float angle, float wet, int plugin);         /* state.rs:429 */
int tdg_add_adsr(tdg_ctx* ctx, float gain, float angle, float wet, int floww, int use_off,
                 int use_max, int note, const float* conf, int n);                        /* state.rs:437 */
int tdg_add_bandpass(tdg_ctx* ctx, float gain, float angle, float wet, float cut_off_hz_low,
                     float cut_off_hz_high, int pass);                                    /* state.rs:453 */
/* Graph::connect_internal (graph.rs:58-78): 1 = connected, 0 = refused (self edge, unknown index,
 * target takes no input, would close a cycle).  Duplicate edges are legal. */
int tdg_connect(tdg_ctx* ctx, int a, int b);
int tdg_set_output(tdg_ctx* ctx, int vertex);                                             /* graph.rs:141 */
/* Graph::check_graph (graph.rs:150-174): 1 ok / 0 failed; lists vertices that do not reach output. */
int tdg_check_graph(tdg_ctx* ctx, int* unreachable, int cap, int* n_unreachable);
int tdg_reset_normalize(tdg_ctx* ctx);                                                    /* graph.rs:207 */
float tdg_get_normalize_value(tdg_ctx* ctx, int vertex);            /* extensions.rs:301; -1 if not one */

/* 'stop' + 'refresh' of a batch of UNCHANGED projects in one call (ui_workflow.rs:153-160; state.rs:240-250,
 * 328-467): per context, optionally reload its flowws (t != NULL; flattened arrays, context i owns
 * flowws_per_ctx[i] consecutive entries, registered in order so that floww ids are unchanged), then
 * tdg_reset_vertex_state + tdg_set_time(0) + tdg_floww_set_time(0) + tdg_reset_normalize. */
int tdg_rewind_many(tdg_ctx** ctxs, int n, const float* const* t_sec, const float* const* note,
                    const float* const* vel, const uint64_t* counts, const uint32_t* flowws_per_ctx);
/* tdg_rewind_many + tdg_normalize_render_many in one pipelined call: the refresh of a group of projects runs on the host
 * while the GPU renders the groups in front of it.  Same results as the two calls in sequence (the projects of a batch
 * are independent).  t == NULL: no floww reload, as in tdg_rewind_many. */
int tdg_refresh_normalize_render_many(tdg_ctx** ctxs, int n, const float* const* t, const float* const* note,
                                      const float* const* vel, const uint64_t* counts, const uint32_t* flowws_per_ctx,
                                      const uint64_t* chunks, const uint32_t* render_sr, const uint32_t* bitdepth,
                                      uint8_t** out_pcm, const uint64_t* cap_bytes, uint64_t* frames_written);

/* ---- transport ----------------------------------------------------------------------------------*/
int tdg_set_time(tdg_ctx* ctx, uint64_t frame);                      /* Graph::set_time, graph.rs:123 */
uint64_t tdg_change_time(tdg_ctx* ctx, uint64_t delta, int plus);    /* Graph::change_time, graph.rs:130 */
uint64_t tdg_get_time(tdg_ctx* ctx);                                 /* Graph::get_time, graph.rs:137 */
int tdg_floww_set_time(tdg_ctx* ctx, uint64_t frame);                /* FlowwBank::set_time, floww.rs:83 */

/* ---- the hot path -------------------------------------------------------------------------------
 * Graph::true_normalize_scan(sb, fb, host, chunks) (graph.rs:222-237), called by State::scan_exact. */
int tdg_normalize_scan(tdg_ctx* ctx, uint64_t chunks);
/* The chunk loop of State::render (state.rs:477-577): `chunks` × Graph::render, optional sinc
 * down-sampling when project_sr > render_sr, quantisation to `bitdepth` and interleaving (L,R) into
 * `out_pcm` exactly as hound would lay the data chunk out (little endian; 8-bit is offset binary).
 * `out_pcm` may be NULL to render without a device->host copy (frames_written is still set).
 * Ends like the reference: graph time rewound to 0, floww cursor NOT rewound (state.rs:575). */
int tdg_render(tdg_ctx* ctx, uint64_t chunks, uint32_t render_sr, uint32_t bitdepth,
               uint8_t* out_pcm, uint64_t cap_bytes, uint64_t* frames_written);
/* Pinned host memory for PCM sinks (cudaHostAlloc / cudaFreeHost): the device->host copy of a render into such a
 * buffer is asynchronous and runs at PCIe speed; the staging half of the WAV output stage (state.rs:508-531 hands
 * every sample to hound one by one; here a file is one header + one pinned block, termdaw_b200/batch.py). */
void* tdg_host_alloc(uint64_t bytes);
void tdg_host_free(void* p);
/* Upper bound of the PCM bytes tdg_render will write for these arguments. */
uint64_t tdg_render_bound_bytes(tdg_ctx* ctx, uint64_t chunks, uint32_t render_sr, uint32_t bitdepth);

/* Playback / stream feed (ui_workflow.rs:195-210, stream_workflow.rs:86-100):
 * `chunks` x { Graph::render; Sample::interleave (sample.rs:177-184); FlowwBank::set_time_to_next_block }
 * as one device pass from the current transport position.  Unlike tdg_render nothing is rewound: time,
 * floww cursor and all vertex state stay where the last chunk left them, so consecutive calls continue the
 * stream (a player keeps its half-second render-ahead ring full by asking for a few chunks at a time).
 * `out_interleaved` receives chunks*buffer_len frames of (L, R) f32 pairs; may be NULL to only advance. */
int tdg_play(tdg_ctx* ctx, uint64_t chunks, float* out_interleaved, uint64_t cap_frames,
             uint64_t* frames_written);
/* The same feed without blocking the player (the render-ahead ring of ui_workflow.rs:195-210 kept full from a
 * pinned host ring): tdg_play_begin queues `chunks` more chunks — one device pass on the playback lane,
 * Sample::interleave on the device, an asynchronous copy into a pinned ring slot — and returns a ticket at once;
 * transport and vertex state have advanced on return, so the next block can be begun immediately.
 * tdg_play_wait blocks until that block has landed and hands out a pointer into the ring (chunks*buffer_len
 * (L, R) f32 pairs; valid until 6 further begins).  At most 6 uncollected tickets per device
 * (TDG_ERR_CAPACITY).  Any other pass of the context first waits for its queued playback work. */
int tdg_play_begin(tdg_ctx* ctx, uint64_t chunks);
int tdg_play_wait(tdg_ctx* ctx, int ticket, const float** interleaved, uint64_t* frames);

/* Batched forms (independent projects): the same two calls over `n` contexts on
 * the same device, merged into one device schedule (one launch per DAG level and kernel class for
 * ALL projects).  Results are identical to calling the single forms in a loop. */
int tdg_normalize_scan_many(tdg_ctx** ctxs, int n, const uint64_t* chunks);
int tdg_render_many(tdg_ctx** ctxs, int n, const uint64_t* chunks, const uint32_t* render_sr,
                    const uint32_t* bitdepth, uint8_t** out_pcm, const uint64_t* cap_bytes,
                    uint64_t* frames_written);

/* "normalize" then "render" (ui_workflow.rs:120-135) for n projects in ONE pipelined call: each group of
 * projects runs its scan pass and its render pass back to back on one stream while the other stream
 * works on the next group, so host planning, kernels and the PCM device->host copies of different
 * groups overlap (give pinned `out_pcm` buffers for the copies to be asynchronous).  Results are
 * identical to tdg_normalize_scan_many followed by tdg_render_many. */
int tdg_normalize_render_many(tdg_ctx** ctxs, int n, const uint64_t* chunks, const uint32_t* render_sr,
                              const uint32_t* bitdepth, uint8_t** out_pcm, const uint64_t* cap_bytes,
                              uint64_t* frames_written);

/* ---- parity taps (pre-quantisation fp32 buffers) -------------------------------------------------
 * The reference keeps one chunk per vertex (graph.rs:242); the device keeps the whole pass.  With
 * taps enabled the buffers of the LAST pass (scan or render) stay readable. */
int tdg_enable_taps(tdg_ctx* ctx, int on);
int tdg_read_vertex_buffer(tdg_ctx* ctx, int vertex, uint64_t frame0, uint64_t n_frames,
                           float* l, float* r);

/* ---- time sharding -----------------------------------------------------------
 * Evaluate chunks [chunk0, chunk0+chunks) of a pass of `total_chunks` chunks that starts at the
 * current transport position.  The event-driven vertex state at chunk0 (voices, envelope clocks, hit
 * lists: extensions.rs:23-69) is rebuilt by replaying the floww events and the per-chunk f32 clock
 * additions from the pass start without rendering audio, and after the call the context's state is
 * what a FULL pass would have left (so that a render after a sharded scan starts from the same leaked
 * state as the reference, graph.rs:228-236).  The two true data dependencies cross the segment
 * boundaries through a communicator:
 *   - Normalize (extensions.rs:321-329): segment peaks — ncclAllReduce(max) in the scan pass,
 *     ncclAllGather + exclusive prefix max in the render pass (running max);
 *   - BandPass (extensions.rs:654-689): each segment publishes its affine map ((1-g)^n, y_end from
 *     zero) per filter vertex, ncclAllGather, and every rank composes the maps of the earlier segments
 *     onto rank 0's initial state — one exchange per DAG level that holds such vertices.
 * One process per GPU: rank 0 calls tdg_comm_unique_id, the launcher broadcasts the 128 bytes
 * (torch.distributed / MPI / a file), every rank calls tdg_comm_init_rank on its context.
 * Without a communicator tdg_*_segment falls back to re-rendering `warmup_chunks` chunks before
 * chunk0 so that filter state converges (approximate; normalize state is then the caller's job via
 * tdg_get/set_normalize_state).
 *   - Render-time resampler (state.rs:533-550): rubato keeps 2*sinc_len = 512 frames of history in
 *     front of each chunk and an f64 index carried from call to call.  `render_sr` below the project
 *     rate (0 = project rate) resamples the segment exactly as the full pass would: the index is
 *     replayed on the host from chunk 0, and the previous segment's last 512 output-vertex frames
 *     arrive as a halo (one more ncclAllGather, 4 KiB per rank).  Segments must then be >= 512 frames.
 *     tdg_segment_output_frames tells the caller where the segment's frames land in the file. */
int tdg_comm_unique_id(uint8_t* out128);
int tdg_comm_init_rank(tdg_ctx* ctx, int nranks, int rank, const uint8_t* id128);
int tdg_comm_destroy(tdg_ctx* ctx);
uint64_t tdg_collectives(tdg_ctx* ctx);        /* NCCL exchange rounds issued so far */
int tdg_scan_segment(tdg_ctx* ctx, uint64_t total_chunks, uint64_t chunk0, uint64_t chunks,
                     uint64_t warmup_chunks);
int tdg_render_segment(tdg_ctx* ctx, uint64_t total_chunks, uint64_t chunk0, uint64_t chunks,
                       uint64_t warmup_chunks, uint32_t render_sr, uint32_t bitdepth,
                       uint8_t* out_pcm, uint64_t cap_bytes, uint64_t* frames_written);
int tdg_segment_output_frames(uint32_t project_sr, uint32_t buffer_len, uint32_t render_sr,
                              uint64_t chunk0, uint64_t chunks, uint64_t* first_frame,
                              uint64_t* n_frames);   /* host only: no device needed */
/* "Virtual ranks": n contexts holding the SAME project on ONE device evaluate n consecutive segments
 * (chunk0[0] == 0, contiguous, covering total_chunks) in one launch sequence and exchange through
 * device memory exactly what NCCL carries between GPUs.  Same kernels, same hand-off arithmetic. */
int tdg_scan_segments_local(tdg_ctx** ctxs, int n, uint64_t total_chunks, const uint64_t* chunk0,
                            const uint64_t* chunks);
int tdg_render_segments_local(tdg_ctx** ctxs, int n, uint64_t total_chunks, const uint64_t* chunk0,
                              const uint64_t* chunks, uint32_t render_sr, uint32_t bitdepth,
                              uint8_t** out_pcm, const uint64_t* cap_bytes, uint64_t* frames_written);
/* Normalize state as (max, scan_max) pairs, one per vertex (non-normalize vertices: -1, -1). */
int tdg_get_normalize_state(tdg_ctx* ctx, float* max_out, float* scan_max_out, int cap);
int tdg_set_normalize_state(tdg_ctx* ctx, const float* max_in, const float* scan_max_in, int n);
/* Device pointer to the per-vertex (max, scan_max) float pairs, for an in-place NCCL all-reduce. */
void* tdg_normalize_state_device_ptr(tdg_ctx* ctx, int* n_floats);

/* ---- measurement ---------------------------------------------------------------------------------
 * Device time (ms, CUDA events on the context's stream) and kernel count of the last scan/render
 * call, and the algorithmic HBM bytes it moved (DESIGN.md "algorithmic bytes"). */
typedef struct tdg_pass_stats {
    double device_ms;        /* first kernel -> last kernel of the pass, CUDA events                */
    double plan_ms;          /* host planning (event -> voice tables)                               */
    double h2d_ms, d2h_ms;   /* host-side wall time spent in the copies                             */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t kernel_launches;
    uint64_t algorithmic_bytes;
    uint64_t frames;         /* project-rate frames rendered                                        */
    uint64_t voice_frames;   /* voice-frames evaluated by the synth kernels                         */
    uint64_t reused_vertices;/* render after a scan: vertices read from the scan's resident buffers  */
    uint64_t reused_bytes;   /* ... and the algorithmic bytes their evaluation would have moved      */
    uint64_t inlined_vertices;/* Sum vertices evaluated inside their consumer's kernel (no buffer)    */
    uint64_t inlined_bytes;  /* ... and the write + read of the intermediate buffer not moved        */
} tdg_pass_stats;
int tdg_last_pass_stats(tdg_ctx* ctx, tdg_pass_stats* out);
/* Per kernel class figures of the last pass (CUDA events around every launch on the context's
 * stream; enable with tdg_set_profiling).  For a batched pass they cover the whole batch. */
typedef struct tdg_kernel_stat {
    const char* name;        /* kernel (class) name as it appears in ncu                            */
    uint64_t launches;
    double device_ms;        /* summed over the launches                                            */
    uint64_t algorithmic_bytes;
} tdg_kernel_stat;
int tdg_set_profiling(tdg_ctx* ctx, int on);
/* Pass-invariant cones: a render that directly follows a normalize scan of the same,
 * unedited context reads the buffers of vertices whose whole input cone carries no state from pass to
 * pass (sample loops, synth / debug-sine voices, sums and band passes over such inputs: everything
 * Graph::set_time resets, extensions.rs:196-204) from the scan instead of evaluating them again; the
 * values are the scan's own, so the result is bit-identical.  On by default; 0 evaluates every vertex in
 * both passes as the reference does (graph.rs:222-237 then state.rs:545-574).  Device-wide setting. */
int tdg_set_pass_reuse(tdg_ctx* ctx, int on);
/* Oscillator sine.  1 (default): glibc's sinf restated operation for operation in double precision on the
 * device (tdg_math.cuh m_sinf_glibc, checked against the host libm over all 2^32 arguments): synth and
 * debug_sine buffers are bit-exact against what the reference's f32::sin produces (extensions.rs:452,
 * synth.rs:23,28) and IIR-free graphs reach <= 1 LSB at 24 bit.  0: an f32 Cody-Waite + minimax sine
 * (<= 2 ulp of libm, a few LSB at 24 bit), slightly faster.  Device-wide setting. */
int tdg_set_oscillator_mode(tdg_ctx* ctx, int libm_exact);
/* Batched calls (tdg_*_many) over more than `projects_per_group` projects are software-pipelined in
 * groups of that size over two streams: host planning of group g+1 and the PCM device->host copy of
 * group g-1 overlap the kernels of group g.  0 = one launch sequence for the whole batch (what the
 * per-kernel roofline numbers are measured on).  Default 32.  Device-wide setting. */
int tdg_set_pipeline_group(tdg_ctx* ctx, uint32_t projects_per_group);
/* Launch shaping (a tuning knob, no effect on results): the chunk kernels let one block walk up to 8 consecutive
 * chunks of a vertex as long as the grid keeps at least `waves` waves of blocks (default 16; 0 = always the maximum,
 * a huge value = always one chunk per block).  Process-wide.  The parity tests run under both extremes. */
int tdg_set_launch_waves(uint32_t waves);
/* Inlined sums (default on): a Sum vertex (add_sum: sum_inputs + pan + gain, extensions.rs:310-319) whose only consumer
 * is a Sum, Normalize or Adsr vertex is evaluated inside that consumer's kernel — same operations, same order, same
 * bits — instead of through a buffer of its own (16 bytes per frame less HBM traffic per inlined vertex).  Never with
 * taps enabled.  0 restores one buffer and one launch per vertex.  Device-wide setting. */
int tdg_set_inline_sums(tdg_ctx* ctx, int on);
int tdg_last_pass_kernel_stats(tdg_ctx* ctx, tdg_kernel_stat* out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* TERMDAW_B200_H */
