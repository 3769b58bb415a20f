// engine.cu — device graph executor, host planner and the C ABI of include/termdaw_b200.h.
//
// Replaces graph.rs's per-buffer_length vertex evaluation (graph.rs:98-121,182-237) with a
// whole-pass schedule: reachable vertices are grouped by DAG depth, every vertex buffer lives in HBM
// for the pass, and each (depth, kernel class) pair is ONE launch covering all projects of the
// batch.  Host work is limited to control data: floww events -> integer frames and per-chunk voice /
// envelope tables (the sequential f32 clock additions of extensions.rs:525-527,649-650 are replayed
// here, exactly), and glibc powf/cosf/sinf for per-voice / per-vertex constants.  No audio sample is
// ever computed on the host; without a CUDA device every entry point fails with TDG_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/termdaw_b200.h"
#include "engine.h"
#include "kernels.h"
#include "tdg_math.cuh"

namespace tdg {

static const float kE = 2.71828182845904523536f;
static const float kFrac1Sqrt2 = 0.70710678118654752440f;

static inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }
static bool trace_on() { static int v = -1; if (v < 0) { const char* e = getenv("TDG_TRACE"); v = (e && *e && *e != '0') ? 1 : 0; } return v == 1; }
// Rust `f32 as usize`
static inline uint64_t f32_as_usize(float x) {
    if (!(x == x) || x <= 0.0f) return 0;
    if (x >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)x;
}

// ------------------------------------------------------------------------------------ engine
struct DevBuf {
    uint8_t* p = nullptr; uint64_t cap = 0;
    // `hint`: the engine's high-water mark for this kind of buffer.  Growing means cudaFree (a device-wide
    // sync) + cudaMalloc in the middle of a pipelined batch, so every buffer of a kind is sized from the
    // largest request any of them has seen, with 25 % headroom: after the first batch nothing regrows.
    cudaError_t ensure(uint64_t bytes, uint64_t hint = 0) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const uint64_t base = hint <= 2 * bytes ? std::max(bytes, hint) : bytes;   // a much larger mark belongs to another batch shape
        uint64_t want = align_up(base + base / 4, 1 << 20);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        if (trace_on()) fprintf(stderr, "[tdg] grow device buffer to %.1f MB (request %.1f MB)\n", want / 1e6, bytes / 1e6);
        return e;
    }
};
struct PinnedBuf {                 // growable pinned staging buffer for the plan blob
    uint8_t* p = nullptr; uint64_t cap = 0, size = 0;
    bool reserve(uint64_t bytes, uint64_t hint = 0) {
        if (bytes <= cap) return true;
        const uint64_t base = hint <= 2 * bytes ? std::max(bytes, hint) : bytes;
        uint64_t want = std::max<uint64_t>(align_up(base + base / 4, 1 << 20), cap * 2);
        uint8_t* q = nullptr;
        if (cudaHostAlloc(&q, want, cudaHostAllocDefault) != cudaSuccess) return false;
        if (p) { memcpy(q, p, size); cudaFreeHost(p); }
        if (trace_on()) fprintf(stderr, "[tdg] grow pinned buffer %.1f -> %.1f MB (request %.1f MB)\n", cap / 1e6, want / 1e6, bytes / 1e6);
        p = q; cap = want;
        return true;
    }
    // returns the byte offset of `bytes` fresh bytes aligned to 16
    uint64_t alloc(uint64_t bytes) {
        uint64_t off = align_up(size, 16);
        if (!reserve(off + bytes)) return UINT64_MAX;
        if (off > size) memset(p + size, 0, off - size);
        size = off + bytes;
        return off;
    }
    template <typename T> T* at(uint64_t off) { return reinterpret_cast<T*>(p + off); }
};

// A lane = one stream + one work arena; a slot = the per-pass plan blob (pinned + device), events and
// pinned error flags.  kLanes lanes x two slots let a batched call keep eight passes in flight
// (run_jobs_pipelined): host planning, H2D, kernels and the PCM D2H of consecutive project groups
// overlap, and a group's render pass can be planned while its scan pass still runs.  With only two
// lanes both finish their kernels together and then both wait on PCIe (the D2H sits in the lane's
// stream because the PCM lives in the lane's arena): the measured timeline (profiles/) showed kernels
// and D2H alternating instead of overlapping; with four, two lanes compute while two drain.
// Single passes use lane 0 / slot 0.
struct Slot {
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_done = nullptr, ev_h2d = nullptr;
    DevBuf blob_dev;
    DevBuf pcm;                    // quantised output + TooWide flags of the pass: NOT in the lane's arena, so the
                                   // next pass of the lane may start while the copy stream still drains this one
    PinnedBuf blob;
    std::vector<cudaEvent_t> evpool;
    uint32_t* err_host = nullptr; size_t err_cap = 0;      // pinned: TooWide flags come back without a pageable copy
    bool ensure_err(size_t n) {
        if (n <= err_cap) return true;
        if (err_host) cudaFreeHost(err_host);
        err_host = nullptr; err_cap = 0;
        const size_t want = std::max<size_t>(n, 1024);
        if (cudaHostAlloc(&err_host, want * 4, cudaHostAllocDefault) != cudaSuccess) return false;
        err_cap = want;
        return true;
    }
};
static const int kLanes = 4, kSlots = 6;
struct Lane {
    cudaStream_t stream = nullptr;
    DevBuf work;
    DevBuf keep;                       // buffers a normalize scan leaves for the render behind it (KeepRecord)
    uint64_t keep_epoch = 0;           // bumped whenever a scan job lays its kept buffers out in `keep`
    Slot slots[kSlots];
};
// Persistent host workers for the planner: projects of a batch are planned independently, and a job of a
// pipelined batch is short (a few ms), so threads are created once per engine, not once per job.
struct WorkerPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::function<void(size_t, unsigned)> fn;    // (item, worker index)
    std::atomic<size_t> next{0};
    size_t n = 0;
    uint64_t gen = 0;
    unsigned busy = 0;
    bool stop = false;
    unsigned workers() const { return (unsigned)th.size() + 1; }     // + the calling thread
    void start(unsigned k) {
        for (unsigned i = 0; i < k; ++i)
            th.emplace_back([this, i]() {
                uint64_t seen = 0;
                for (;;) {
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [&]() { return stop || gen != seen; });
                        if (stop) return;
                        seen = gen;
                    }
                    for (;;) { const size_t j = next.fetch_add(1); if (j >= n) break; fn(j, i + 1); }
                    { std::lock_guard<std::mutex> lk(mu); if (--busy == 0) cv_done.notify_one(); }
                }
            });
    }
    // fn(item, worker) for item in [0, count); returns when all are done.  One caller at a time.
    void parallel_for(size_t count, unsigned max_workers, std::function<void(size_t, unsigned)> f) {
        if (count == 0) return;
        const unsigned use = (unsigned)std::min<size_t>(std::min<size_t>(max_workers, workers()), count);
        if (use <= 1 || th.empty()) { for (size_t j = 0; j < count; ++j) f(j, 0); return; }
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = std::move(f); n = count; next.store(0); busy = (unsigned)th.size(); ++gen;
        }
        cv.notify_all();
        for (;;) { const size_t j = next.fetch_add(1); if (j >= n) break; fn(j, 0); }
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&]() { return busy == 0; });
    }
    ~WorkerPool() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
};

struct Frag;
struct PlanScratch;
struct VPlan;
// clock table of one (sr, bl) (tdg_types.h): rows are only ever appended while passes are in flight; growing past
// the allocation waits for the device first
struct ClockTab { float* p = nullptr; uint32_t depth = 0, cap = 0; cudaEvent_t ev = nullptr; bool ev_set = false; };
struct Engine {
    int device = 0;
    Lane lanes[kLanes];
    bool taper = true;                           // TDG_TAPER=0: equal groups (A/B measurements)
    int n_lanes = 2;                             // lanes a pipelined batch rotates over (TDG_LANES=1..4; measured best: 2)
    cudaStream_t stream = nullptr;               // == lanes[0].stream (bank uploads, state sync)
    cudaStream_t copy_stream = nullptr;          // PCM device->host copies of every lane (one PCIe direction, one queue)
    cudaStream_t h2d_stream = nullptr;           // plan-blob uploads: run ahead of the lane's kernels instead of between them
    std::map<uint32_t, float*> sinc_tables;     // keyed by the f32 bits of the cutoff
    uint64_t epoch = 0;                          // bumped by every pass (taps of older passes are stale)
    int refs = 0;
    bool profiling = false;                      // CUDA events around every launch (per-class roofline)
    uint32_t pipeline_group = 32;                // projects per pipelined group (0 = never pipeline); measured per 512-project call: 16 -> 68, 32 -> 65, 64 -> 69, 128 -> 79 ms
    uint64_t hwm_work = 0, hwm_blob = 0, hwm_pcm = 0;         // largest work arena / plan blob any lane or slot has needed
    uint64_t hwm_keep = 0;
    bool reuse = true;                           // render reads pass-invariant cones from the scan (TDG_PASS_REUSE=0 / tdg_set_pass_reuse)
    // band-pass launch groups: inputs + outputs per group.  Measured on B200 (profiles/r02): with 48 MB groups the 512-project
    // step has 735 launch triples whose one-block-per-filter carry kernels and short grids cost far more (28 ms) than the
    // L2 hits save, so the default is one launch triple per level (0); the knob stays for single long renders (TDG_BP_L2_MB)
    uint64_t bp_l2_bytes = 0;
    bool inline_sums = true;                     // Sum vertices with one k_fx consumer are evaluated inside that consumer (TDG_INLINE_SUMS=0: every vertex owns a buffer)
    bool bp_fused = true;                        // single-pass band pass (k_bp_fused) where no exchange step is needed and the pass is <= kBpFusedMaxTiles long; TDG_BP_FUSED=0: three launches
    uint32_t bp_fused_max_tiles = kBpFusedMaxTiles;   // longer passes keep the three-launch scan (TDG_BP_FUSED_MAX_TILES)
    bool libm_sin = true;                        // oscillators evaluate glibc's sinf exactly (TDG_FAST_SIN=1 / tdg_set_oscillator_mode(ctx, 0): f32 polynomial, <= 2 ulp)
    WorkerPool* pool = nullptr;                  // planner threads (created on first batched pass)
    std::vector<Frag>* frags = nullptr;          // pooled per-project plan fragments (capacity is kept between jobs)
    std::vector<PlanScratch>* scratch = nullptr; // per worker: event-table scratch vectors
    std::vector<std::vector<VPlan>>* vplans = nullptr;   // per request of a job x vertex: event plans (pooled)
    std::map<uint64_t, ClockTab> clocks;         // key: sr << 32 | bl
    // asynchronous playback feed (tdg_play_begin / tdg_play_wait): its own lane, a ring of plan slots, one pinned
    // host buffer per slot that the interleaved stream lands in
    struct PlayJob* play_slot[kSlots] = {};
    float* play_ring[kSlots] = {}; uint64_t play_ring_cap[kSlots] = {};
    float* play_dev[kSlots] = {}; uint64_t play_dev_cap[kSlots] = {};
    cudaEvent_t play_ev[kSlots] = {};
    int play_next = 0;
    cudaEvent_t ev_batch0 = nullptr, ev_batch1[kLanes] = {}, ev_copy_end = nullptr;   // span of a pipelined call
};
static std::mutex g_mu;
static std::map<int, Engine*> g_engines;

static Engine* engine_get(int device, std::string* err) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_engines.find(device);
    if (it != g_engines.end()) { it->second->refs++; return it->second; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { *err = std::string("no CUDA device: ") + cudaGetErrorString(e); return nullptr; }
    if (device < 0 || device >= n) { *err = "device index out of range"; return nullptr; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { *err = cudaGetErrorString(e); return nullptr; }
    if (prop.major != 10) {
        *err = "termdaw_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        return nullptr;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { *err = cudaGetErrorString(e); return nullptr; }
    Engine* g = new Engine();
    g->device = device;
    for (Lane& ln : g->lanes) {
        bool ok = cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking) == cudaSuccess;
        for (Slot& sl : ln.slots)
            ok = ok && cudaEventCreate(&sl.ev0) == cudaSuccess && cudaEventCreate(&sl.ev1) == cudaSuccess &&
                 cudaEventCreate(&sl.ev_done) == cudaSuccess && cudaEventCreateWithFlags(&sl.ev_h2d, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { *err = "could not create stream/events"; delete g; return nullptr; }
    }
    g->stream = g->lanes[0].stream;
    if (cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&g->h2d_stream, cudaStreamNonBlocking) != cudaSuccess) { *err = "could not create the copy stream"; delete g; return nullptr; }
    cudaEventCreate(&g->ev_batch0); cudaEventCreate(&g->ev_copy_end);
    for (int i = 0; i < kLanes; ++i) cudaEventCreate(&g->ev_batch1[i]);
    if (const char* e = getenv("TDG_LANES")) g->n_lanes = std::max(1, std::min(kLanes, atoi(e)));
    if (const char* e = getenv("TDG_TAPER")) g->taper = atoi(e) != 0;
    if (const char* e = getenv("TDG_PASS_REUSE")) g->reuse = atoi(e) != 0;
    if (const char* e = getenv("TDG_FAST_SIN")) g->libm_sin = atoi(e) == 0;
    if (const char* e = getenv("TDG_BP_L2_MB")) g->bp_l2_bytes = (uint64_t)std::max(0, atoi(e)) << 20;
    if (const char* e = getenv("TDG_BP_FUSED")) g->bp_fused = atoi(e) != 0;
    if (const char* e = getenv("TDG_INLINE_SUMS")) g->inline_sums = atoi(e) != 0;
    if (const char* e = getenv("TDG_BP_FUSED_MAX_TILES")) g->bp_fused_max_tiles = (uint32_t)std::max(1, std::min(65535 * kBpStripTiles, atoi(e)));
    g->refs = 1;
    g_engines[device] = g;
    return g;
}

// ------------------------------------------------------------------------------------ NCCL (time sharding)
// Bound at run time (dlopen) so that the library loads on boxes without NCCL and shares whichever
// libnccl.so.2 the process already holds (torch's).  Only the handful of entry points used.
struct Id128 { char b[128]; };          // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128), passed by value
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string err;
    bool ok() const { return lib && GetUniqueId && CommInitRank && CommDestroy && AllGather && AllReduce; }
};
static const int kNcclFloat32 = 7, kNcclMax = 2;       // ncclFloat32 / ncclMax of nccl.h (stable ABI values)
static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, []() {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.lib) break; }
        if (!api.lib) { api.err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return; }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
        if (!api.ok()) api.err = "libnccl is missing an expected symbol";
    });
    return &api;
}
#define NC(call) do { int _e = (call); if (_e != 0) return c->fail(TDG_ERR_CUDA, std::string(#call) + ": " + (nccl_api()->GetErrorString ? nccl_api()->GetErrorString(_e) : "nccl error")); } while (0)

#define CU(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return c->fail(TDG_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); } while (0)

// ------------------------------------------------------------------------------------ sinc table
// rubato make_sincs (BlackmanHarris2, f32) — see DESIGN.md "resampler"; [UNVERIFIED-RECALL] upstream.
static void build_sinc_table(float cutoff, std::vector<float>& out) {
    const size_t npoints = 256, factor = 256, tot = npoints * factor;
    std::vector<float> y(tot);
    const float pi = TDG_PI_F, pi2 = 2.0f * pi, pi4 = 4.0f * pi, pi6 = 6.0f * pi, np_f = (float)tot;
    float sum = 0.0f;
    for (size_t x = 0; x < tot; ++x) {
        const float xf = (float)x;
        float w = 0.35875f - 0.48829f * cosf(pi2 * xf / np_f) + 0.14128f * cosf(pi4 * xf / np_f) -
                  0.01168f * cosf(pi6 * xf / np_f);
        w = w * w;
        const float v = (xf - (float)(tot / 2)) * cutoff / (float)factor;
        float s = 1.0f;
        if (v != 0.0f) { const float a = v * pi; s = sinf(a) / a; }
        const float val = w * s;
        sum += val;
        y[x] = val;
    }
    sum /= (float)factor;
    out.assign(tot, 0.0f);
    for (size_t p = 0; p < npoints; ++p)
        for (size_t n = 0; n < factor; ++n) out[(factor - n - 1) * npoints + p] = y[factor * p + n] / sum;
}
static float* sinc_table_dev(Ctx* c, double ratio) {
    const float cutoff = ratio >= 1.0 ? 0.95f : 0.95f * (float)ratio;
    uint32_t key; memcpy(&key, &cutoff, 4);
    Engine* g = c->eng;
    auto it = g->sinc_tables.find(key);
    if (it != g->sinc_tables.end()) return it->second;
    std::vector<float> tab;
    build_sinc_table(cutoff, tab);
    float* d = nullptr;
    if (cudaMalloc(&d, tab.size() * 4) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
    g->sinc_tables[key] = d;
    return d;
}
// exact f64 replay of SincFixedIn's index loop: per call, how many frames come out and where it starts
static void resample_replay(double ratio, uint64_t chunk, uint64_t n_chunks, std::vector<ResampleChunk>& out, uint64_t* total) {
    const double t_ratio = 1.0 / ratio;
    const int64_t end_idx = (int64_t)chunk - 257 - (int64_t)ceil(t_ratio);
    double last = -128.0;
    uint64_t off = 0;
    out.resize(n_chunks);
    for (uint64_t k = 0; k < n_chunks; ++k) {
        double idx = last; uint32_t n = 0;
        while (idx < (double)end_idx) { idx += t_ratio; ++n; }
        out[k].last_index = last; out[k].out_off = off; out[k].n_out = n; out[k].pad = 0;
        off += n;
        last = idx - (double)chunk;
    }
    *total = off;
}

// chunks [chunk0, chunk0 + n) of a pass that started at chunk 0: SincFixedIn's index is carried from call
// to call (f64, sequential), so a segment replays the calls in front of it; out_off is rebased to the
// segment and *first_out is the output frame the segment starts at in the file
static void resample_replay_slice(double ratio, uint64_t chunk, uint64_t chunk0, uint64_t n, std::vector<ResampleChunk>& out,
                                  uint64_t* first_out, uint64_t* total) {
    std::vector<ResampleChunk> all;
    uint64_t tot = 0;
    resample_replay(ratio, chunk, chunk0 + n, all, &tot);
    const uint64_t first = n ? all[chunk0].out_off : tot;
    out.assign(all.begin() + (ptrdiff_t)chunk0, all.end());
    for (ResampleChunk& c : out) c.out_off -= first;
    *first_out = first; *total = tot - first;
}

// ------------------------------------------------------------------------------------ vertex state sync
static int vstate_to_host(Ctx* c) {
    if (c->vstate_host_newer || !c->vstate_dev) return TDG_OK;
    cudaSetDevice(c->eng->device);
    size_t n = std::min(c->vstate_cap, c->vstate_host.size());
    CU(cudaMemcpyAsync(c->vstate_host.data(), c->vstate_dev, n * 4, cudaMemcpyDeviceToHost, c->eng->stream));
    CU(cudaStreamSynchronize(c->eng->stream));
    return TDG_OK;
}
static int vstate_to_dev(Ctx* c, cudaStream_t stream = nullptr) {
    if (!stream) stream = c->eng->stream;
    const size_t need = c->verts.size() * 8;
    if (c->vstate_host.size() != need) c->vstate_host.resize(need, 0.0f);
    if (c->vstate_cap < need || !c->vstate_dev) {
        if (int r = vstate_to_host(c)) return r;
        if (c->vstate_dev) cudaFree(c->vstate_dev);
        c->vstate_dev = nullptr;
        CU(cudaMalloc(&c->vstate_dev, std::max<size_t>(need, 8) * 4));
        c->vstate_cap = std::max<size_t>(need, 8);
        c->vstate_host_newer = true;
    }
    if (c->vstate_host_newer && need) {
        // pageable source: the runtime stages the bytes before returning, so vstate_host may be touched
        // right after; stream order makes the pass's kernels see them
        CU(cudaMemcpyAsync(c->vstate_dev, c->vstate_host.data(), need * 4, cudaMemcpyHostToDevice, stream));
    }
    c->vstate_host_newer = false;
    return TDG_OK;
}

// ------------------------------------------------------------------------------------ planner
struct PassParams {
    uint64_t F0;        // FlowwBank.frame at pass chunk 0
    uint64_t t0;        // graph time handed to generate() at pass chunk 0
    uint64_t total;     // chunks of the whole pass (state always advances over all of them)
    uint64_t emit0, emit_n;   // chunks whose audio is evaluated on this device
    uint32_t bl, sr;
};

struct EvCursor {       // events of one floww in pass order (sorted => equals the reference's cursor, floww.rs:70-141)
    const Floww* f; size_t j;
    EvCursor(const Floww* fl, uint64_t F0) : f(fl) { j = std::lower_bound(fl->fe.begin(), fl->fe.end(), F0) - fl->fe.begin(); }
};

static PanGain make_pg(float gain, float angle) {
    PanGain pg{1.0f, 1.0f, gain, 0};
    if (!(fabsf(angle) < 0.001f)) {                                 // sample.rs:98-101
        const float ar = angle * 0.5f * 0.01745329f;
        pg.l_amp = kFrac1Sqrt2 * (cosf(ar) + sinf(ar));
        pg.r_amp = kFrac1Sqrt2 * (cosf(ar) - sinf(ar));
        pg.flags |= 1u;
    }
    if (!(fabsf(gain - 1.0f) < 0.001f)) pg.flags |= 2u;             // sample.rs:109
    return pg;
}
static inline float note_hz(float note) { return 440.0f * powf(2.0f, (note - 69.0f) / 12.0f); }   // extensions.rs:451

// get_block_drum over one chunk (floww.rs:99-121): first event of a frame with vel > 0.001 is the hit
template <typename F>
static void drum_chunk(EvCursor& cur, uint64_t Fk, uint32_t bl, F&& on_hit) {
    const Floww& f = *cur.f;
    uint64_t done = UINT64_MAX;
    while (cur.j < f.fe.size() && f.fe[cur.j] < Fk + bl) {
        const size_t j = cur.j++;
        if (f.fe[j] < Fk || f.fe[j] == done) continue;
        if (f.vel[j] > 0.001f) { done = f.fe[j]; on_hit((uint32_t)(f.fe[j] - Fk), f.note[j], f.vel[j]); }
    }
}
// get_block_simple over one chunk (floww.rs:124-141): every event, in order
template <typename F>
static void simple_chunk(EvCursor& cur, uint64_t Fk, uint32_t bl, F&& on_ev) {
    const Floww& f = *cur.f;
    while (cur.j < f.fe.size() && f.fe[cur.j] < Fk + bl) {
        const size_t j = cur.j++;
        if (f.fe[j] < Fk) continue;
        on_ev((uint32_t)(f.fe[j] - Fk), f.vel[j] > 0.001f, f.note[j], f.vel[j]);
    }
}
static inline bool note_ok(const HVertex& v, float note) {          // extensions.rs:353-356
    return v.has_note ? fabsf(note - (float)v.note) < 0.01f : true;
}

// sample_multi_gen: hit list for the pass + end-of-pass deque (extensions.rs:344-381)
static void plan_multi(HVertex& v, const Floww& fl, uint64_t L, const PassParams& pp, std::vector<MultiHit>& hits) {
    hits.clear();
    for (const MultiEntry& e : v.ts) hits.push_back({-e.t, e.vel, 0.0f});
    EvCursor cur(&fl, pp.F0);
    for (uint64_t k = 0; k < pp.total; ++k)
        drum_chunk(cur, pp.F0 + k * pp.bl, pp.bl, [&](uint32_t i, float note, float vel) {
            if (note_ok(v, note)) hits.push_back({(int64_t)(k * pp.bl + i), vel, 0.0f});
        });
    const int64_t n = (int64_t)(pp.total * pp.bl);
    v.ts.clear();
    for (const MultiHit& h : hits) if (h.start + (int64_t)L >= n) v.ts.push_back({n - h.start, h.vel});
    const int64_t shift = (int64_t)(pp.emit0 * pp.bl);
    if (shift) for (MultiHit& h : hits) h.start -= shift;
}
// sample_lerp_gen (extensions.rs:384-421): [carried ghost, carried primary, hits...]
static void plan_lerp(HVertex& v, const Floww& fl, const PassParams& pp, std::vector<LerpHit>& hits) {
    hits.clear();
    const int64_t far = INT64_MIN / 4;
    hits.push_back({far, -v.ghost.t, far, v.ghost.vel, 0.0f});
    hits.push_back({far + 1, -v.primary.t, (int64_t)v.countdown - (int64_t)v.lerp_len, v.primary.vel, 0.0f});
    EvCursor cur(&fl, pp.F0);
    for (uint64_t k = 0; k < pp.total; ++k)
        drum_chunk(cur, pp.F0 + k * pp.bl, pp.bl, [&](uint32_t i, float note, float vel) {
            if (note_ok(v, note)) { const int64_t p = (int64_t)(k * pp.bl + i); hits.push_back({p, p, p, vel, 0.0f}); }
        });
    const int64_t n = (int64_t)(pp.total * pp.bl);
    const LerpHit &P = hits[hits.size() - 1], &G = hits[hits.size() - 2];
    v.primary = {n - P.pos_origin, P.vel};
    v.ghost = {n - G.pos_origin, G.vel};
    const int64_t cd = (int64_t)v.lerp_len - (n - P.cd_origin);
    v.countdown = cd > 0 ? (uint64_t)cd : 0;
    const int64_t shift = (int64_t)(pp.emit0 * pp.bl);
    if (shift) for (size_t h = 0; h < hits.size(); ++h) {
        if (h >= 2) hits[h].frame -= shift;
        hits[h].pos_origin -= shift; hits[h].cd_origin -= shift;
    }
}

// Event plan of one vertex for one pass: O(events) records; the per-chunk tables are made on the device
// (k_chunk_ranges, the in-kernel staging of k_voice / k_fx<adsr>).  Pooled per (request, vertex).
struct VPlan {
    std::vector<VoiceRec> vrecs;
    std::vector<AdsrRec> arecs;
    std::vector<MultiHit> mhits;
    std::vector<LerpHit> lhits;
    std::vector<float> xenv;            // explicit clocks, concatenated (state carried in from an earlier pass)
    uint64_t voice_frames = 0;
    uint64_t seg_cap = 0;               // voices: upper bound of the per-chunk segments inside the emitted range
    uint32_t depth = 0;                 // rows of the clock table this vertex reads
    bool planned = false;
};
struct PlanScratch {                    // per planner thread
    std::vector<std::vector<float>> xc; std::vector<size_t> xoff;       // explicit clocks under construction
    // host copy of the clock table (tdg_types.h) of one (sr, bl), rows appended on demand: a voice's clock at its
    // note-off in O(1) instead of one addition per chunk it was held
    std::vector<float> hclk; uint32_t hclk_sr = 0, hclk_bl = 0, hclk_rows = 0;
    std::vector<float> nv; std::vector<uint32_t> hit;   // plan_voices: the live voices' note values side by side (the per-event scan), the matches of one event
    std::vector<std::pair<uint32_t, uint32_t>> drops;   // plan_voices: pending drops of released voices (chunk, index), a min-heap
    static const uint32_t kHostClockRows = 16384;
    float clock(uint32_t sr, uint32_t bl, uint32_t i, uint32_t n) {
        const float fsr = (float)sr, d = (float)bl / fsr;
        if (n >= kHostClockRows || i >= bl) { float e = -((float)i / fsr); for (uint32_t q = 0; q < n; ++q) e += d; return e; }
        if (hclk_sr != sr || hclk_bl != bl) { hclk_sr = sr; hclk_bl = bl; hclk_rows = 0; hclk.clear(); }
        if (n >= hclk_rows) {
            const uint32_t want = std::min(kHostClockRows, std::max(n + 1, hclk_rows * 2));
            hclk.resize((size_t)want * bl);
            for (uint32_t r = hclk_rows; r < want; ++r) {
                float* row = hclk.data() + (size_t)r * bl;
                if (r == 0) for (uint32_t c = 0; c < bl; ++c) row[c] = -((float)c / fsr);
                else { const float* prev = row - bl; for (uint32_t c = 0; c < bl; ++c) row[c] = prev[c] + d; }
            }
            hclk_rows = want;
        }
        return hclk[(size_t)n * bl + i];
    }
};

// explicit clocks -> out.xenv; *src words holding (kEnvExplicit | clock id) become (kEnvExplicit | float index)
static void pack_explicit(PlanScratch& S, size_t n_x, std::vector<float>& xenv) {
    S.xoff.assign(n_x + 1, 0);
    xenv.clear();
    for (size_t c = 0; c < n_x; ++c) { S.xoff[c] = xenv.size(); xenv.insert(xenv.end(), S.xc[c].begin(), S.xc[c].end()); }
}
static inline uint32_t patch_src(const PlanScratch& S, uint32_t src) {
    return (src & kEnvExplicit) ? (kEnvExplicit | (uint32_t)S.xoff[src & ~kEnvExplicit]) : src;
}

// debug_sine_gen / synth_gen / sampsyn_gen voice lifecycle (extensions.rs:429-445, 481-496, 525-528): one
// VoiceRec per voice, in Vec order.  Event driven: nothing is done per chunk.  A voice's clock is needed on the
// host at its note-off only (rel_t = env_t + fi) and for the drop rule after it (env_t > release_sec at the end of
// a chunk), and both are replayed there with the reference's own sequence of f32 additions.
static void plan_voices(HVertex& v, const Floww& fl, const PassParams& pp, VPlan& out, PlanScratch& S) {
    std::vector<VoiceRec>& R = out.vrecs;
    R.clear(); out.xenv.clear(); out.voice_frames = 0; out.depth = 0; out.seg_cap = 0;
    size_t n_x = 0;
    auto new_xclk = [&]() -> int { if (S.xc.size() <= n_x) S.xc.emplace_back(); S.xc[n_x].clear(); return (int)n_x++; };
    const float fsr = (float)pp.sr;
    const float d = (float)pp.bl / fsr;                              // len as f32 / sr as f32
    float release_sec = 0.0f;
    if (v.kind == K_SYNTH) {                                         // extensions.rs:469-478
        if (v.sq.volume > 0.0f) release_sec = v.sq.adsr.release_sec;
        if (v.tf.volume > 0.0f) release_sec = fmaxf(release_sec, v.tf.adsr.release_sec);
        if (v.tr.volume > 0.0f) release_sec = fmaxf(release_sec, v.tr.adsr.release_sec);
    } else if (v.kind == K_SAMPSYN) release_sec = v.adsr.release_sec;
    if (pp.total == 0) return;
    const uint32_t total = (uint32_t)pp.total, last = total - 1;
    auto open_rec = [&](float hz, float vel, uint32_t k, uint32_t i, uint32_t envh) -> int {
        VoiceRec r; memset(&r, 0, sizeof r);
        r.hz = hz; r.vel = vel; r.on_chunk = k; r.on_i = i; r.off_chunk = kNever; r.end_chunk = kNever; r.envh = envh;
        R.push_back(r);
        return (int)R.size() - 1;
    };
    // a record right behind `at`, i.e. in the same place of the reference's Vec (in-place changes of a voice)
    auto insert_after = [&](int at, const VoiceRec& r) -> int {
        R.insert(R.begin() + at + 1, r);
        for (DNote& n : v.dnotes) if (n.rec > at) ++n.rec;
        for (SNote& n : v.notes) if (n.rec > at) ++n.rec;
        return at + 1;
    };
    auto close_rec = [&](int rec, uint32_t k, uint32_t i) { R[rec].end_chunk = k; R[rec].end_i = i; };
    // the clock of voice n at the start of pass chunk k (k >= its clk_chunk): explicit clocks grow on demand
    auto clock_of = [&](const SNote& n, uint32_t k) -> float {
        if (n.xclk >= 0) {
            std::vector<float>& x = S.xc[n.xclk];
            if (x.size() <= k) {                                     // one growth, then the additions in a tight loop
                size_t q = x.size();
                x.resize((size_t)k + 1);
                float e = x[q - 1];
                for (; q <= k; ++q) { e += d; x[q] = e; }
            }
            return x[k];
        }
        return S.clock(pp.sr, pp.bl, n.clk_i, k - n.clk_chunk);
    };
    // extensions.rs:528 runs at the end of every chunk: a released voice goes once env_t > release_sec.  From the
    // clock value e at the start of chunk k: the chunk at whose end it is dropped (kNever: outlives the pass)
    auto drop_after = [&](float e, uint32_t k) -> uint32_t {
        for (uint32_t q = k; q < total; ++q) { e += d; if (!(e <= release_sec)) return q; }
        return kNever;
    };
    // voices alive from an earlier pass: their clocks are not of the table's form -> explicit per-chunk values
    for (DNote& n : v.dnotes) n.rec = open_rec(n.hz, n.vel, 0, 0, 0);
    for (SNote& n : v.notes) {
        n.xclk = new_xclk();
        S.xc[n.xclk].push_back(n.env_t);
        n.clk_chunk = 0; n.clk_i = 0;
        n.rec = open_rec(n.hz, n.vel, 0, 0, 0);
        VoiceRec& r = R[n.rec];
        if (n.rel_t == 0.0f) { r.envh = kEnvExplicit | (uint32_t)n.xclk; n.drop_chunk = kNever; }
        else {
            r.off_chunk = 0; r.off_i = 0; r.envr = kEnvExplicit | (uint32_t)n.xclk; r.rel_t = n.rel_t; r.relv0 = n.relv[0]; r.relv1 = n.relv[1]; r.relv2 = n.relv[2];
            n.drop_chunk = drop_after(n.env_t, 0);
        }
    }
    // Bookkeeping of the live voices during the replay.  The reference erases a voice from its Vec the moment it goes
    // (`retain`, extensions.rs:485,528); here a voice that goes is only marked (its note value becomes +inf, which no
    // note-off matches) and the vector is compacted — stably: the order of the survivors is the reference's — once
    // half of it is dead and at the end of the pass.  A note-off looks for its voices in `nv`, the note values side by
    // side (one vectorised scan over a few KB instead of a walk over 64-byte records), and the drops scheduled by
    // released voices sit in a heap keyed by their chunk: per event O(live / SIMD width + log live) where erasing in
    // place was O(live) per dropped voice — configs[2] re-strikes 512 held voices 15 000 times per pass.
    std::vector<float>& nv = S.nv;
    nv.resize(v.notes.size());
    for (size_t q = 0; q < v.notes.size(); ++q) nv[q] = v.notes[q].note;
    const float kDead = __builtin_inff();
    size_t n_dead = 0;
    std::vector<std::pair<uint32_t, uint32_t>>& heap = S.drops;      // (drop chunk, index in v.notes), min-heap
    heap.clear();
    auto heap_cmp = [](const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& b) { return a > b; };
    auto schedule_drop = [&](uint32_t idx) {
        if (v.notes[idx].drop_chunk == kNever) return;
        heap.push_back({v.notes[idx].drop_chunk, idx}); std::push_heap(heap.begin(), heap.end(), heap_cmp);
    };
    for (size_t q = 0; q < v.notes.size(); ++q) schedule_drop((uint32_t)q);
    auto kill = [&](uint32_t idx) { nv[idx] = kDead; ++n_dead; };
    auto compact = [&]() {
        if (n_dead == 0) return;
        size_t w = 0;
        for (size_t q = 0; q < v.notes.size(); ++q) {
            if (nv[q] == kDead) continue;
            if (w != q) { v.notes[w] = v.notes[q]; nv[w] = nv[q]; }
            ++w;
        }
        v.notes.resize(w); nv.resize(w); n_dead = 0;
        heap.clear();
        for (size_t q = 0; q < w; ++q) if (v.notes[q].drop_chunk != kNever) heap.push_back({v.notes[q].drop_chunk, (uint32_t)q});
        std::make_heap(heap.begin(), heap.end(), heap_cmp);
    };
    auto purge = [&](uint32_t k) {                                   // everything dropped at the end of a chunk < k
        while (!heap.empty() && heap.front().first < k) {
            const uint32_t idx = heap.front().second, at = heap.front().first;
            std::pop_heap(heap.begin(), heap.end(), heap_cmp); heap.pop_back();
            if (nv[idx] == kDead || v.notes[idx].drop_chunk != at) continue;           // gone already / stale entry
            close_rec(v.notes[idx].rec, at, pp.bl);
            kill(idx);
        }
        if (n_dead > 64 && 2 * n_dead > v.notes.size()) compact();
    };
    EvCursor cur(&fl, pp.F0);
    for (uint64_t k64 = 0; k64 < pp.total; ++k64) {
        const uint32_t k = (uint32_t)k64;
        simple_chunk(cur, pp.F0 + k64 * pp.bl, pp.bl, [&](uint32_t i, bool on, float note, float vel) {
            const float fi = (float)i / fsr;
            if (v.kind == K_DEBUG_SINE) {
                if (on) {
                    bool has = false;
                    for (DNote& n : v.dnotes)
                        if (fabsf(n.note - note) < 0.001f) {                 // the velocity is replaced in place (extensions.rs:433-438)
                            close_rec(n.rec, k, i);
                            VoiceRec c = R[n.rec]; c.vel = vel; c.on_chunk = k; c.on_i = i; c.end_chunk = kNever; c.end_i = 0;
                            n.rec = insert_after(n.rec, c); n.vel = vel; has = true; break;
                        }
                    if (!has) { const float hz = note_hz(note); v.dnotes.push_back({note, vel, hz, open_rec(hz, vel, k, i, 0)}); }
                } else {
                    size_t w = 0;
                    for (size_t q = 0; q < v.dnotes.size(); ++q) {
                        if (fabsf(v.dnotes[q].note - note) > 0.001f) v.dnotes[w++] = v.dnotes[q];
                        else close_rec(v.dnotes[q].rec, k, i);
                    }
                    v.dnotes.resize(w);
                }
                return;
            }
            purge(k);
            if (on) {
                const float hz = note_hz(note);
                SNote n{note, vel, 0.0f, 0.0f, hz, open_rec(hz, vel, k, i, i), -1};
                n.clk_chunk = k; n.clk_i = i; n.drop_chunk = kNever;
                v.notes.push_back(n); nv.push_back(note);
                return;
            }
            // retain (extensions.rs:485) drops the voices of this note that are already releasing; the others of this
            // note start their release (extensions.rs:486-494).  One scan over the note values finds the note's
            // voices (usually one), in list order.
            bool any_rel = false, any_held = false;
            std::vector<uint32_t>& hit = S.hit;
            hit.clear();
            { const float* pv = nv.data(); const size_t nn = nv.size();
              for (size_t b0 = 0; b0 < nn; b0 += 32) {                      // branch-free per block of 32 (vectorises), indices only where a block matched
                  const size_t b1 = std::min(nn, b0 + 32);
                  unsigned any = 0;
                  for (size_t q = b0; q < b1; ++q) any |= (unsigned)!(fabsf(pv[q] - note) > 0.001f);
                  if (any) for (size_t q = b0; q < b1; ++q) if (!(fabsf(pv[q] - note) > 0.001f)) hit.push_back((uint32_t)q);
              } }
            for (uint32_t q : hit) { if (v.notes[q].rel_t == 0.0f) any_held = true; else any_rel = true; }
            if (any_rel) for (uint32_t q : hit) if (v.notes[q].rel_t != 0.0f) { close_rec(v.notes[q].rec, k, i); kill(q); }
            if (any_held) for (uint32_t hq : hit) {
                SNote& n = v.notes[hq];
                if (nv[hq] == kDead || n.rel_t != 0.0f) continue;
                n.rel_t = clock_of(n, k) + fi;                              // env_t + i/sr
                n.xclk = -1; n.clk_chunk = k; n.clk_i = i;                  // env_t = -(i/sr): the clock restarts, column i
                // apply_r_rt's start value apply_ads(conf, rel_t) (adsr.rs:89-92): constant from here on
                if (v.kind == K_SYNTH) {
                    n.relv[0] = m_apply_ads(v.sq.adsr, n.rel_t); n.relv[1] = m_apply_ads(v.tf.adsr, n.rel_t);
                    n.relv[2] = m_apply_ads(v.tr.adsr, n.rel_t);
                } else n.relv[0] = m_apply_ads(v.adsr, n.rel_t);
                if (n.rel_t != 0.0f) {
                    VoiceRec& r = R[n.rec];
                    r.off_chunk = k; r.off_i = i; r.envr = i; r.rel_t = n.rel_t;
                    r.relv0 = n.relv[0]; r.relv1 = n.relv[1]; r.relv2 = n.relv[2];
                    n.drop_chunk = drop_after(-fi, k);
                    schedule_drop(hq);
                } else {
                    // env_t + fi == 0.0 exactly (a note-off in the frame of its note-on): the voice still reads as
                    // held (rel_t == 0.0) but its clock restarted: a continuation record in its place
                    close_rec(n.rec, k, i);
                    VoiceRec c = R[n.rec]; c.on_chunk = k; c.on_i = i; c.envh = i; c.off_chunk = kNever; c.end_chunk = kNever; c.end_i = 0;
                    n.rec = insert_after(n.rec, c);
                }
            }
        });
    }
    purge(total);
    compact();
    for (DNote& n : v.dnotes) close_rec(n.rec, last, pp.bl);
    for (SNote& n : v.notes) {
        close_rec(n.rec, last, pp.bl);
        if (n.xclk >= 0) (void)clock_of(n, last);                           // the device reads the explicit clock up to the last chunk
        n.env_t = clock_of(n, total);                                        // carried into the next pass
    }
    // explicit clocks of voices that ended inside the pass: values up to their last chunk
    for (const VoiceRec& r : R) {
        const uint32_t srcs[2] = {r.envh, r.off_chunk != kNever ? r.envr : 0u};
        for (uint32_t src : srcs) if (src & kEnvExplicit) {
            std::vector<float>& x = S.xc[src & ~kEnvExplicit];
            const uint32_t upto = std::min(r.end_chunk, last);
            while (x.size() <= upto) x.push_back(x.back() + d);
        }
    }
    pack_explicit(S, n_x, out.xenv);
    const uint64_t e0 = pp.emit0 * pp.bl, e1 = (pp.emit0 + pp.emit_n) * pp.bl;
    uint32_t pm = 0;
    for (VoiceRec& r : R) {
        r.envh = patch_src(S, r.envh); r.envr = patch_src(S, r.envr);
        pm = std::max(pm, r.end_chunk); r.pm_end = pm;
        const uint32_t held_end = std::min(r.off_chunk, r.end_chunk);
        if (!(r.envh & kEnvExplicit) && held_end >= r.on_chunk) out.depth = std::max(out.depth, held_end - r.on_chunk + 1);
        if (r.off_chunk != kNever && !(r.envr & kEnvExplicit) && r.end_chunk >= r.off_chunk) out.depth = std::max(out.depth, r.end_chunk - r.off_chunk + 1);
        const uint64_t a = std::max<uint64_t>((uint64_t)r.on_chunk * pp.bl + r.on_i, e0), b = std::min<uint64_t>((uint64_t)r.end_chunk * pp.bl + r.end_i, e1);
        if (b > a) out.voice_frames += b - a;
        // chunks of the emitted range the voice touches, + 1 where its release starts inside one of them
        const uint64_t c_lo = std::max<uint64_t>(r.on_chunk, pp.emit0), c_hi = std::min<uint64_t>(r.end_chunk, pp.emit0 + pp.emit_n - 1);
        if (pp.emit_n && c_hi >= c_lo) out.seg_cap += c_hi - c_lo + 1 + ((r.off_chunk != kNever && r.off_chunk >= c_lo && r.off_chunk <= c_hi) ? 1 : 0);
    }
    uint32_t sm = kNever;
    for (size_t q = R.size(); q-- > 0;) { sm = std::min(sm, R[q].on_chunk); R[q].sm_on = sm; }
}

// adsr_gen state machine (extensions.rs:593-651): one AdsrRec per state change
static void plan_adsr(HVertex& v, const Floww& fl, const PassParams& pp, VPlan& out, PlanScratch& S) {
    std::vector<AdsrRec>& R = out.arecs;
    R.clear(); out.xenv.clear(); out.depth = 0;
    size_t n_x = 0;
    auto new_xclk = [&]() -> int { if (S.xc.size() <= n_x) S.xc.emplace_back(); S.xc[n_x].clear(); return (int)n_x++; };
    const float fsr = (float)pp.sr;
    const float d = (float)pp.bl / fsr;
    A3 &P = v.aprimary, &G = v.aghost;
    struct Clk { uint32_t src, chunk; int x; };
    // both slots arrive with whatever the pass before left in them: explicit clocks until an event restarts them
    Clk Pc, Gc;
    { const int a = new_xclk(), b = new_xclk(); Pc = Clk{kEnvExplicit | (uint32_t)a, 0, a}; Gc = Clk{kEnvExplicit | (uint32_t)b, 0, b}; }
    auto push = [&](uint64_t start, uint32_t flags) {
        const AdsrRec r{start, flags, 0, Pc.src, Pc.chunk, P.vel, P.rel, Gc.src, Gc.chunk, G.vel, G.rel};
        if (!R.empty() && R.back().start == start) R.back() = r; else R.push_back(r);
    };
    push(0, 0);
    EvCursor cur(&fl, pp.F0);
    for (uint64_t k64 = 0; k64 < pp.total; ++k64) {
        const uint32_t k = (uint32_t)k64;
        const uint64_t c0 = k64 * pp.bl;
        if (Pc.x >= 0) S.xc[Pc.x].push_back(P.t);
        if (Gc.x >= 0) S.xc[Gc.x].push_back(G.t);
        if (v.use_off) {
            simple_chunk(cur, pp.F0 + c0, pp.bl, [&](uint32_t i, bool on, float note, float vel) {
                if (v.has_note && fabsf((float)v.note - note) > 0.01f) return;
                const float fi = (float)i / fsr;
                const float offset = fi;
                if (on) { G = P; Gc = Pc; P = {-fi, vel, 0.0f}; Pc = Clk{i, k, -1}; }
                else if (G.rel == 0.0f) { G.t = -fi; Gc = Clk{i, k, -1}; G.rel = m_apply_ads(v.adsr, G.t + offset) * G.vel; }
                else { P.t = -fi; Pc = Clk{i, k, -1}; P.rel = m_apply_ads(v.adsr, P.t + offset) * P.vel; }
                push(c0 + i, 0);
            });
        } else {
            drum_chunk(cur, pp.F0 + c0, pp.bl, [&](uint32_t i, float note, float vel) {
                if (v.has_note && fabsf((float)v.note - note) > 0.01f) {    // `continue` on the frame loop
                    push(c0 + i, 1); push(c0 + i + 1, 0);
                    return;
                }
                const float fi = (float)i / fsr;
                G = P; Gc = Pc; P = {-fi, vel, 0.0f}; Pc = Clk{i, k, -1};
                push(c0 + i, 0);
            });
        }
        P.t += d; G.t += d;                                                 // extensions.rs:649-650
    }
    pack_explicit(S, n_x, out.xenv);
    const uint64_t n_frames = pp.total * pp.bl;
    for (size_t q = 0; q < R.size(); ++q) {
        AdsrRec& r = R[q];
        r.p_src = patch_src(S, r.p_src); r.g_src = patch_src(S, r.g_src);
        const uint64_t until = q + 1 < R.size() ? R[q + 1].start : n_frames;     // the record is read for frames [start, until)
        if (until == 0 || until <= r.start) continue;
        const uint32_t lastc = (uint32_t)((until - 1) / pp.bl);
        if (!(r.p_src & kEnvExplicit) && lastc >= r.p_chunk) out.depth = std::max(out.depth, lastc - r.p_chunk + 1);
        if (!(r.g_src & kEnvExplicit) && lastc >= r.g_chunk) out.depth = std::max(out.depth, lastc - r.g_chunk + 1);
    }
}

// ------------------------------------------------------------------------------------ pass executor
struct PassReq {
    Ctx* c = nullptr;
    bool is_scan = false;
    uint64_t total = 0, emit0 = 0, emit_n = 0, keep0 = 0;   // keep0: first emitted chunk that is output
    uint32_t render_sr = 0, bd = 16;
    uint8_t* out = nullptr; uint64_t cap = 0; uint64_t* frames_written = nullptr;
    bool segment = false;
    int x_mode = 0;                    // 0: no exchange, 1: virtual ranks in this launch, 2: NCCL
    int x_ranks = 1, x_rank = 0;
    // derived
    std::vector<uint32_t> x_stride, x_norms;       // per level: floats per rank slot, normalize vertices
    std::vector<int64_t> x_off, x_red_off;         // per level: exchange buffer / all-reduce buffer (arena offsets)
    std::vector<uint32_t> x_index;                 // per vertex: offset inside the rank slot
    std::vector<int> level;            // per vertex, -1 unreachable
    std::vector<int> last_use;
    std::vector<int64_t> buf_off, aux_off;   // work-arena byte offsets
    std::vector<int64_t> rng_off;            // per event-driven vertex: uint2[emit_n] written by k_chunk_ranges
    std::vector<int64_t> bps_off;            // per band-pass vertex: ticket + per-tile flags of the fused kernel, inside the job's zeroed sync region
    std::vector<int64_t> seg_off;            // per voice vertex: seg_ptr[emit_n + 1], segs[cap][2] (float4), relv2[cap]
    std::vector<uint64_t> seg_cap;
    uint64_t n_frames = 0, n_pad = 0;
    uint64_t pcm_off = 0, pcm_bytes = 0, err_off = 0, rs_off = 0, rs_total = 0, rs_idx_off = 0;
    bool resample = false;
    const float* sincs = nullptr;
    std::vector<ResampleChunk> rs_chunks;
    int64_t halo_off = -1;             // time-sharded resampling: [x_ranks][2][kResampleHalo] floats (arena offset)
    // pass-invariant cones (KeepRecord): vertices the render pass does not evaluate, and buffers living in the
    // lane's keep arena (written by the scan pass / read by the render pass)
    std::vector<char> skip;
    std::vector<int> inl;              // per vertex: >= 0: a Sum vertex evaluated INSIDE that consumer's kernel (no buffer, no launch)
    std::vector<int64_t> keep_rel;     // per vertex: keep-arena offset, -1 = work arena
    std::vector<uint64_t> prefix;      // per vertex: chunks to evaluate (UINT64_MAX = all); < emit_n only for kept multi / lerp in a render
    bool keep_write = false, keep_read = false;
    bool play = false;                 // playback feed: no quantiser, and the pass does not rewind at its end (tdg_play)
    uint64_t out_frames = 0;
    uint32_t err_flag = 0;
};

template <typename T> struct TaskList { std::vector<T> v; uint32_t max_chunks = 0; uint64_t bytes = 0; size_t n_plain = SIZE_MAX;   // fx lists: tasks [n_plain, size) have inlined inputs (grouped kernel variant)
                                        void clear() { v.clear(); max_chunks = 0; bytes = 0; n_plain = SIZE_MAX; } };
struct LevelTasks {
    TaskList<FxTask> fx, fx_adsr, fx_norm; TaskList<NormPrefixTask> nprefix; TaskList<NormScaleTask> nscale;
    TaskList<LoopTask> loop; TaskList<MultiTask> multi; TaskList<LerpTask> lerp;
    TaskList<VoiceTask> dsine, synth, sampsyn; TaskList<BpTask> bp;
    bool fx_adsr_exact = false, dsine_exact = false, synth_exact = false, sampsyn_exact = false;
    uint64_t bp_agg_bytes = 0;
    void clear() {                       // keeps the vectors' capacity (pooled fragments)
        fx.clear(); fx_adsr.clear(); fx_norm.clear(); nprefix.clear(); nscale.clear(); loop.clear(); multi.clear(); lerp.clear();
        dsine.clear(); synth.clear(); sampsyn.clear(); bp.clear();
        fx_adsr_exact = dsine_exact = synth_exact = sampsyn_exact = false; bp_agg_bytes = 0;
    }
};

static void compute_levels(PassReq& r) {
    Ctx* c = r.c;
    const int V = (int)c->verts.size();
    r.level.assign(V, -1); r.last_use.assign(V, -1);
    if (c->output < 0) return;
    // iterative post-order DFS from the output (graph.rs:98-121 evaluates exactly this set)
    std::vector<std::pair<int, size_t>> st;
    std::vector<char> seen(V, 0);
    st.push_back({c->output, 0}); seen[c->output] = 1;
    while (!st.empty()) {
        const int v = st.back().first;
        const size_t ei = st.back().second;
        const auto& e = c->verts[v].edges;
        if (ei < e.size()) { st.back().second++; const int u = e[ei]; if (!seen[u]) { seen[u] = 1; st.push_back({u, 0}); } }
        else {
            int lv = 0;
            for (int u : e) lv = std::max(lv, r.level[u] + 1);
            r.level[v] = lv;
            for (int u : e) r.last_use[u] = std::max(r.last_use[u], lv);
            st.pop_back();
        }
    }
}

struct HeapBlob {                   // per-project plan fragment (same alloc/at interface as PinnedBuf)
    std::vector<uint8_t> v;
    uint64_t alloc(uint64_t bytes) { const uint64_t off = align_up(v.size(), 16); v.resize(off + bytes, 0); return off; }
    template <typename T> T* at(uint64_t off) { return reinterpret_cast<T*>(v.data() + off); }
};
struct Frag {
    HeapBlob blob; uint64_t base = 0;
    std::vector<LevelTasks> levels;
    std::vector<QuantTask> quants; std::vector<ResampleTask> resamps;
    std::vector<RangeTask> ranges; uint32_t range_max = 0, resamp_chunks = 0; bool any_voice = false;
    uint64_t quant_max = 0, quant_bytes = 0, resamp_max = 0, resamp_bytes = 0, algo = 0, reused_v = 0, reused_b = 0, inlined_v = 0, inlined_b = 0;
    void reset(size_t n_levels) {        // pooled on the engine: capacities survive from job to job
        blob.v.clear(); base = 0; ranges.clear(); range_max = 0; resamp_chunks = 0; any_voice = false;
        if (levels.size() > n_levels) levels.resize(n_levels);
        for (LevelTasks& L : levels) L.clear();
        levels.resize(n_levels);
        quants.clear(); resamps.clear();
        quant_max = quant_bytes = resamp_max = resamp_bytes = algo = reused_v = reused_b = inlined_v = inlined_b = 0;
    }
};
template <typename T> static void append_list(TaskList<T>& d, TaskList<T>& s) {
    d.v.insert(d.v.end(), s.v.begin(), s.v.end());
    d.max_chunks = std::max(d.max_chunks, s.max_chunks); d.bytes += s.bytes;
}

struct Bump { uint64_t off = 0; uint64_t alloc(uint64_t bytes) { off = align_up(off, 256); uint64_t r = off; off += bytes; return r; } };

struct Off { uint64_t fx, fx_adsr, fx_norm, nprefix, nscale, loop, multi, lerp, dsine, synth, sampsyn, bp; };
struct Timed { int cls; uint64_t bytes; };
// One pass over a batch of requests, in three phases so that consecutive groups can overlap:
// plan (host: layout, event tables, task lists), submit (async: H2D, launches, D2H), finish (sync,
// statistics, end-of-pass transport state).
struct PassJob {
    std::vector<PassReq> reqs;
    Lane* lane = nullptr;
    int slot = 0;
    bool submitted = false, waited = false, in_batch = false;
    KStat ks[KC_COUNT];                 // per kernel class, filled by pass_wait
    int max_level = -1;
    bool all_vec = true;
    uint8_t* W = nullptr;
    uint8_t* P = nullptr;               // the slot's PCM buffer
    uint64_t err_base = 0, err_bytes = 0; // TooWide flags: one u32 per request, contiguous inside P
    std::vector<LevelTasks> levels;
    std::vector<Off> offs;
    std::vector<QuantTask> quants; std::vector<ResampleTask> resamps;
    std::vector<RangeTask> ranges; uint32_t range_max = 0, resamp_chunks = 0; uint64_t range_off = 0; bool any_voice = false;
    std::vector<cudaEvent_t> clock_events;   // clock tables this job reads that were extended on another stream
    uint64_t sync_off = 0, sync_bytes = 0;    // tickets + flags of the fused band-pass kernels of the job (work arena, zeroed at submit)
    uint64_t quant_max = 0, quant_bytes = 0, resamp_max = 0, resamp_bytes = 0, quant_off = 0, resamp_off = 0;
    uint64_t algo_bytes_total = 0, reused_v = 0, reused_b = 0, inlined_v = 0, inlined_b = 0;
    std::function<int()> pre;           // host work that must precede this job's planning (the group's refresh, tdg_refresh_normalize_render_many)
    double t_start = 0, t_planned = 0, t_h2d_issued = 0, t_d2h0 = 0, t_submitted = 0;
    float dev_t0 = 0, dev_t1 = 0, dev_tdone = 0;      // device timeline relative to the batch start (trace only)
    uint64_t launches = 0, collectives = 0, d2h_bytes = 0, h2d_bytes = 0;
    std::vector<Timed> timed;
};

static int pass_plan(PassJob& J) {
    std::vector<PassReq>& reqs = J.reqs;
    Lane& ln = *J.lane;
    if (reqs.empty()) return TDG_OK;
    Ctx* c = reqs[0].c;                       // error sink for CU()
    Engine* g = c->eng;
    for (auto& r : reqs) if (r.c->eng != g) return r.c->fail(TDG_ERR_INVALID, "all contexts of a batch must be on one device");
    cudaSetDevice(g->device);
    J.t_start = now_ms();
    g->epoch++;

    // ---- 1. graph analysis + work-arena layout -------------------------------------------------
    Bump bump, pbump, kbump, fbump;
    const int lane_index = (int)(J.lane - g->lanes);
    int& max_level = J.max_level;
    max_level = -1;
    for (auto& r : reqs) {
        Ctx* x = r.c;
        compute_levels(r);
        const int V = (int)x->verts.size();
        r.n_frames = r.emit_n * x->bl;
        r.n_pad = align_up(std::max<uint64_t>(r.n_frames, 1), kBpTile);
        r.buf_off.assign(V, -1); r.aux_off.assign(V, -1); r.rng_off.assign(V, -1); r.seg_off.assign(V, -1); r.seg_cap.assign(V, 0); r.bps_off.assign(V, -1);
        for (int v = 0; v < V; ++v) max_level = std::max(max_level, r.level[v]);
    }
    J.err_bytes = reqs.size() * 4;
    J.err_base = pbump.alloc(J.err_bytes);
    uint64_t keep_limit = 0; bool keep_limit_known = false;
    auto keep_fits = [&](uint64_t bytes) {
        if (bytes + 256 <= ln.keep.cap) return true;                       // no growth: nothing to ask the driver
        if (!keep_limit_known) {
            size_t free_b = 0, total_b = 0;
            keep_limit = cudaMemGetInfo(&free_b, &total_b) == cudaSuccess ? ln.keep.cap + free_b / 2 : 0;
            keep_limit_known = true;
        }
        return bytes + bytes / 4 + (2u << 20) <= keep_limit;               // DevBuf::ensure grows with 25 % headroom
    };
    for (size_t ri = 0; ri < reqs.size(); ++ri) {
        PassReq& r = reqs[ri];
        r.err_off = J.err_base + ri * 4;
        Ctx* x = r.c;
        const int V = (int)x->verts.size();
        const uint64_t slot_bytes = r.n_pad * 2 * 4;
        std::vector<uint64_t> free_slots;
        std::vector<std::vector<int>> by_level(max_level + 1);
        for (int v = 0; v < V; ++v) if (r.level[v] >= 0) by_level[r.level[v]].push_back(v);
        std::vector<char> freed(V, 0);
        // ---- pass-invariant cones.  A vertex is pass-invariant when the render pass that follows a normalize
        // scan evaluates it to the very values the scan produced: its kind carries no state across set_time
        // (sample loop restarts at 0, synth / debug-sine voices are cleared, band pass re-seeds from its first
        // frame: extensions.rs:196-204, 664-670), it starts the scan from that reset state, and every input is
        // pass-invariant.  The scan keeps the buffers such cones hand to the rest of the graph; the render reads
        // them (same kernels, same bits: they ARE the scan's output) instead of evaluating the cone again.
        r.skip.assign(V, 0); r.keep_rel.assign(V, -1); r.prefix.assign(V, UINT64_MAX); r.keep_write = r.keep_read = false;
        const bool plain = g->reuse && !r.segment && !r.play && r.x_ranks == 1 && !x->taps && r.emit0 == 0 && r.emit_n == r.total &&
                           x->output >= 0 && r.emit_n > 0;
        KeepRecord& K = x->keep;
        if (!r.is_scan) {
            r.keep_read = plain && K.valid && K.lane == lane_index && K.epoch == ln.keep_epoch && ln.keep.p != nullptr &&
                          K.mut_gen == x->mut_gen && K.total == r.total && K.n_pad == r.n_pad && (int)K.inv.size() == V &&
                          x->t == 0 && x->fb_frame == 0;
            K.valid = false;                                   // consumed, or stale: the next scan writes a new record
        } else {
            K.valid = false;
            if (plain) {
                K.inv.assign(V, 0); K.off.assign(V, -1); K.hz.assign(V, UINT64_MAX);
                for (int L = 0; L <= max_level; ++L) for (int v : by_level[L]) {
                    const HVertex& hv = x->verts[v];
                    bool inv = false;
                    switch (hv.kind) {
                        case K_SAMPLE_LOOP: inv = hv.loop_t == 0; break;
                        case K_DEBUG_SINE: inv = hv.dnotes.empty(); break;
                        case K_SYNTH: inv = hv.notes.empty(); break;
                        case K_SUM: inv = true; break;
                        case K_BANDPASS: inv = hv.bp_first || hv.wet < 0.0001f || (hv.lgamma == 0.0f && hv.hgamma == 0.0f); break;
                        default: break;
                    }
                    if (inv) for (int u : hv.edges) if (!K.inv[u]) { inv = false; break; }
                    K.inv[v] = inv ? 1 : 0;
                }
                // frontier: invariant vertices consumed by a vertex that is not (or the output itself)
                std::vector<char> fr(V, 0);
                for (int v = 0; v < V; ++v) if (r.level[v] >= 0 && !K.inv[v]) for (int u : x->verts[v].edges) if (K.inv[u]) fr[u] = 1;
                if (K.inv[x->output]) fr[x->output] = 1;
                // prefix candidates: kept whole, the planner below tells how many leading chunks the render must redo
                for (int v = 0; v < V; ++v)
                    if (r.level[v] >= 0 && (x->verts[v].kind == K_SAMPLE_MULTI || x->verts[v].kind == K_SAMPLE_LERP)) fr[v] = 1;
                // kept buffers are extra device memory: when the keep arena would have to grow past half of what is
                // free (an hours-long single render) this project is simply evaluated twice, as before
                uint64_t n_fr = 0;
                for (int v = 0; v < V; ++v) n_fr += fr[v] ? 1 : 0;
                if (n_fr && keep_fits(kbump.off + n_fr * (slot_bytes + 256)))
                    for (int v = 0; v < V; ++v) if (fr[v]) { r.keep_rel[v] = (int64_t)kbump.alloc(slot_bytes); K.off[v] = r.keep_rel[v]; r.keep_write = true; }
                K.lane = lane_index; K.mut_gen = x->mut_gen; K.total = r.total; K.n_pad = r.n_pad;
            }
        }
        // ---- inlined sums.  A Sum vertex (extensions.rs:310-319 + pan + gain) with exactly one consumer that is itself a
        // k_fx vertex (Sum, Normalize, Adsr) is evaluated inside that consumer's kernel: the consumer reads the sum's
        // inputs, adds them in the sum's edge order from +0.0, applies the sum's pan and gain, and adds the result to its
        // own running sum at the position of the edge — the reference's operations in the reference's order, without
        // the 8 B/frame write and the 8 B/frame read of the intermediate buffer and without its launch.  One level deep
        // only (an inlined sum's own inputs are real buffers); never for a buffer somebody else needs (taps, the kept
        // frontier of a pass-invariant cone, the output).
        r.inl.assign(V, -1);
        if (g->inline_sums && !x->taps && x->output >= 0) {
            std::vector<int> n_cons(V, 0), cons(V, -1);
            for (int v = 0; v < V; ++v) if (r.level[v] >= 0) for (int u : x->verts[v].edges) { ++n_cons[u]; cons[u] = v; }
            for (int L = 0; L <= max_level; ++L) for (int v : by_level[L]) {
                const HVertex& hv = x->verts[v];
                if (hv.kind != K_SUM || v == x->output || n_cons[v] != 1 || hv.edges.empty() || hv.edges.size() > 8) continue;
                if (r.keep_rel[v] >= 0) continue;                                               // scan: kept for the render
                if (r.keep_read && (K.inv[v] || K.off[v] >= 0)) continue;                       // render: read from the scan
                const int cv = cons[v];
                const HVertex& hc = x->verts[cv];
                if (hc.kind != K_SUM && hc.kind != K_NORMALIZE && hc.kind != K_ADSR) continue;
                if (r.keep_read && K.inv[cv]) continue;                                         // the consumer is not evaluated at all
                bool nested = false;
                for (int u : hv.edges) if (r.inl[u] >= 0) { nested = true; break; }
                if (nested) continue;
                r.inl[v] = cv;
                for (int u : hv.edges) r.last_use[u] = std::max(r.last_use[u], r.level[cv]);   // its inputs live until the consumer has run
            }
        }
        for (int L = 0; L <= max_level; ++L) {
            for (int v : by_level[L]) {
                if (r.inl[v] >= 0) continue;                                                            // no buffer of its own
                if (r.keep_read && K.inv[v]) { r.skip[v] = 1; r.keep_rel[v] = K.off[v]; continue; }    // not evaluated; kept buffer if frontier
                if (r.keep_read && K.off[v] >= 0 && (int)K.hz.size() == V) { r.keep_rel[v] = K.off[v]; r.prefix[v] = K.hz[v]; }   // only its head is redone
                if (r.keep_rel[v] >= 0) { /* scan pass: this buffer lives in the keep arena */ }
                else if (!free_slots.empty()) { r.buf_off[v] = (int64_t)free_slots.back(); free_slots.pop_back(); }
                else r.buf_off[v] = (int64_t)bump.alloc(slot_bytes);
                const HVertex& hv = x->verts[v];
                if (hv.kind == K_NORMALIZE) r.aux_off[v] = (int64_t)bump.alloc(r.emit_n * 2 * 4 + 16);     // peaks, runmax
                if (hv.kind == K_BANDPASS) {
                    r.aux_off[v] = (int64_t)bump.alloc((r.n_pad / kBpTile) * 12 * 4 + 16); // agg[8], carry[4]
                    r.bps_off[v] = (int64_t)fbump.alloc((r.n_pad / kBpTile + 1) * 4);      // ticket, flags: one memset per job clears them all
                }
                if (hv.kind == K_ADSR || hv.kind == K_SAMPLE_MULTI || hv.kind == K_SAMPLE_LERP || hv.kind == K_DEBUG_SINE ||
                    hv.kind == K_SYNTH || hv.kind == K_SAMPSYN) r.rng_off[v] = (int64_t)bump.alloc(std::max<uint64_t>(r.emit_n, 1) * 8 + 16);
            }
            if (!x->taps) {
                auto release = [&](int u) {
                    if (!freed[u] && r.last_use[u] == L && u != x->output && r.buf_off[u] >= 0) { freed[u] = 1; free_slots.push_back((uint64_t)r.buf_off[u]); }
                };
                for (int v : by_level[L]) {
                    if (r.inl[v] >= 0) continue;                 // its inputs are released by the consumer's level
                    for (int u : x->verts[v].edges) {
                        if (r.inl[u] == v) { for (int w : x->verts[u].edges) release(w); }
                        else release(u);
                    }
                }
            }
        }
        if (r.x_ranks > 1) {
            // exchange slots: per level [normalize peaks][band-pass maps]; identical on every rank
            // because every rank holds the same graph
            r.x_stride.assign(max_level + 1, 0); r.x_norms.assign(max_level + 1, 0);
            r.x_off.assign(max_level + 1, -1); r.x_red_off.assign(max_level + 1, -1);
            r.x_index.assign(V, 0);
            auto is_bp = [&](const HVertex& hv) { return hv.kind == K_BANDPASS && !(hv.wet < 0.0001f || (hv.lgamma == 0.0f && hv.hgamma == 0.0f)); };
            for (int v = 0; v < V; ++v) if (r.level[v] >= 0 && x->verts[v].kind == K_NORMALIZE) r.x_index[v] = r.x_norms[r.level[v]]++;
            for (int L = 0; L <= max_level; ++L) r.x_stride[L] = r.x_norms[L];
            for (int v = 0; v < V; ++v) if (r.level[v] >= 0 && is_bp(x->verts[v])) { r.x_index[v] = r.x_stride[r.level[v]]; r.x_stride[r.level[v]] += kBpXchgFloats; }
            PassReq* leader = &r;
            if (r.x_mode == 1) for (auto& q : reqs) if (q.x_mode == 1) { leader = &q; break; }
            if (leader == &r) {
                for (int L = 0; L <= max_level; ++L) if (r.x_stride[L]) {
                    r.x_off[L] = (int64_t)bump.alloc((uint64_t)r.x_ranks * r.x_stride[L] * 4);
                    r.x_red_off[L] = (int64_t)bump.alloc(std::max<uint32_t>(r.x_norms[L], 1) * 4);
                }
            } else { r.x_off = leader->x_off; r.x_red_off = leader->x_red_off; }
        }
        if (!r.is_scan && x->output >= 0) {
            r.resample = !r.play && x->sr > r.render_sr;                       // state.rs:533 (playback is at the project rate)
            r.out_frames = (r.emit0 + r.emit_n - r.keep0) * x->bl;
            if (r.resample) {
                const uint64_t keep_n = r.emit0 + r.emit_n - r.keep0;
                uint64_t first_out = 0;
                if (r.segment) resample_replay_slice((double)r.render_sr / (double)x->sr, x->bl, r.keep0, keep_n, r.rs_chunks, &first_out, &r.rs_total);
                else resample_replay((double)r.render_sr / (double)x->sr, x->bl, r.emit_n, r.rs_chunks, &r.rs_total);
                r.rs_off = bump.alloc(align_up(r.rs_total, 4) * 2 * 4 + 64);
                r.rs_idx_off = bump.alloc(std::max<uint64_t>(r.rs_total, 1) * 8);
                r.out_frames = r.rs_total;
                // the resampler's history in front of a segment (512-frame halo): warm-up chunks of
                // this segment if there are enough of them, else the previous rank's tail through the exchange
                const uint64_t before = (r.keep0 - r.emit0) * x->bl;
                if (r.keep0 > 0 && before < (uint64_t)kResampleHalo) {
                    if (r.x_ranks <= 1)
                        return x->fail(TDG_ERR_INVALID, "a resampled segment needs a communicator, virtual ranks or >= 512 warm-up frames");
                }
                if (r.x_ranks > 1) {
                    if (keep_n * x->bl < (uint64_t)kResampleHalo)
                        return x->fail(TDG_ERR_INVALID, "resampled segments must be at least 512 frames long");
                    PassReq* leader = &r;
                    if (r.x_mode == 1) for (auto& q : reqs) if (q.x_mode == 1) { leader = &q; break; }
                    if (leader == &r) r.halo_off = (int64_t)bump.alloc((uint64_t)r.x_ranks * 2 * kResampleHalo * 4);
                    else r.halo_off = leader->halo_off;
                }
            }
            r.pcm_bytes = r.play ? 0 : r.out_frames * 2 * (r.bd / 8);
            r.pcm_off = pbump.alloc(align_up(r.pcm_bytes, 64) + 64);
        }
    }
    const double tr_layout0 = now_ms();
    if (!g->pool) {
        g->pool = new WorkerPool();
        // planner threads: the host's cores shared fairly between the ranks of this node (one process per GPU):
        // 8 ranks x 31 threads on a 32-core host was what bounded the 8-GPU step
        unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned local_world = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) local_world = (unsigned)std::max(1, atoi(e));
        unsigned want = std::max(1u, std::min(hw / local_world, 32u));
        if (const char* e = getenv("TDG_PLAN_THREADS")) want = (unsigned)std::max(1, atoi(e));
        g->pool->start(want - 1);
        g->frags = new std::vector<Frag>();
        g->scratch = new std::vector<PlanScratch>(g->pool->workers());
        g->vplans = new std::vector<std::vector<VPlan>>();
    }
    if (g->frags->size() < reqs.size()) g->frags->resize(reqs.size());
    // ---- 2a. event plans: every event-driven vertex of every project is independent -> one work item each
    // (a batch of small projects and one long 64-vertex project both spread over the planner threads)
    if (g->vplans->size() < reqs.size()) g->vplans->resize(reqs.size());
    std::vector<std::vector<VPlan>>& vplans_a = *g->vplans;
    struct Item { uint32_t ri, v; };
    std::vector<Item> items;
    for (size_t ri = 0; ri < reqs.size(); ++ri) {
        PassReq& r = reqs[ri];
        Ctx* x = r.c;
        const size_t V = x->verts.size();
        if (vplans_a[ri].size() < V) vplans_a[ri].resize(V);
        for (size_t v = 0; v < V; ++v) {
            vplans_a[ri][v].planned = false;
            if (r.level[v] < 0 || r.skip[v]) continue;
            const HVertex& hv = x->verts[v];
            const bool ev = (hv.kind == K_ADSR && !(hv.wet < 0.0001f)) || hv.kind == K_SAMPLE_MULTI || hv.kind == K_SAMPLE_LERP ||
                            hv.kind == K_DEBUG_SINE || hv.kind == K_SYNTH || hv.kind == K_SAMPSYN;
            if (ev) items.push_back({(uint32_t)ri, (uint32_t)v});
        }
    }
    g->pool->parallel_for(items.size(), 64, [&](size_t q, unsigned w) {
        PassReq& r = reqs[items[q].ri];
        Ctx* x = r.c;
        const uint32_t v = items[q].v;
        HVertex& hv = x->verts[v];
        VPlan& VP = vplans_a[items[q].ri][v];
        PlanScratch& PS = (*g->scratch)[w];
        const PassParams pp{r.is_scan ? 0 : x->fb_frame, r.is_scan ? 0 : x->t, r.total, r.emit0, r.emit_n, x->bl, x->sr};
        // (true_normalize_scan: fb.set_time(0) and t = j*bl, graph.rs:228-231)
        VP.depth = 0; VP.voice_frames = 0;
        const double t_item = trace_on() ? now_ms() : 0.0;
        switch (hv.kind) {
            case K_ADSR: plan_adsr(hv, x->flowws[hv.floww], pp, VP, PS); break;
            case K_SAMPLE_MULTI: plan_multi(hv, x->flowws[hv.floww], x->samples[hv.sample].frames, pp, VP.mhits); break;
            case K_SAMPLE_LERP: plan_lerp(hv, x->flowws[hv.floww], pp, VP.lhits); break;
            default: plan_voices(hv, x->flowws[hv.floww], pp, VP, PS); break;
        }
        VP.planned = true;
        if (trace_on() && now_ms() - t_item > 5.0) fprintf(stderr, "[tdg] slow event plan: vertex %u kind %d %.2f ms (worker %u)\n", v, hv.kind, now_ms() - t_item, w);
    });
    // segment tables of the voice vertices (built on the device): their capacity is known once the events are planned
    for (const Item& it : items) {
        PassReq& r = reqs[it.ri];
        const HVertex& hv = r.c->verts[it.v];
        if (hv.kind != K_DEBUG_SINE && hv.kind != K_SYNTH && hv.kind != K_SAMPSYN && hv.kind != K_ADSR) continue;
        // adsr: one state per chunk plus one per state change inside a chunk (AdsrSeg is 32 B like a voice segment)
        const uint64_t cap = hv.kind == K_ADSR ? r.emit_n + vplans_a[it.ri][it.v].arecs.size() + 1
                                                : std::max<uint64_t>(vplans_a[it.ri][it.v].seg_cap, 1);
        r.seg_cap[it.v] = cap;
        r.seg_off[it.v] = (int64_t)bump.alloc(align_up((r.emit_n + 1) * 4, 16) + cap * 32 + cap * 4 + 16);
    }
    J.sync_bytes = fbump.off;
    J.sync_off = fbump.off ? bump.alloc(fbump.off + 256) : 0;
    const double tr_layout = now_ms();
    g->hwm_work = std::max<uint64_t>(g->hwm_work, bump.off + 256);
    CU(ln.work.ensure(bump.off + 256, g->hwm_work));
    uint8_t* W = J.W = ln.work.p;
    if (kbump.off) {
        // a scan job lays its kept buffers out from the start of the lane's keep arena: whatever an earlier
        // scan on this lane left there is gone (epoch), and growing it waits for the device (cudaFree)
        g->hwm_keep = std::max<uint64_t>(g->hwm_keep, kbump.off + 256);
        CU(ln.keep.ensure(kbump.off + 256, g->hwm_keep));
        ++ln.keep_epoch;
    }
    for (auto& r : reqs) {
        if (r.keep_write) r.c->keep.epoch = ln.keep_epoch;
        // kept buffers are addressed like every other buffer, as an offset from the work arena's base
        for (size_t v = 0; v < r.keep_rel.size(); ++v)
            if (r.keep_rel[v] >= 0) r.buf_off[v] = (int64_t)((ln.keep.p + r.keep_rel[v]) - W);
    }
    g->hwm_pcm = std::max<uint64_t>(g->hwm_pcm, pbump.off + 256);
    CU(ln.slots[J.slot].pcm.ensure(pbump.off + 256, g->hwm_pcm));
    uint8_t* P = J.P = ln.slots[J.slot].pcm.p;
    const double tr_arena = now_ms();

    // ---- 2. plan: tasks + tables into the pinned blob ------------------------------------------
    // Each project is planned independently (host threads) into its own fragment: a private blob and
    // private task lists whose *_off fields are relative to that blob; fragments are then copied into
    // the pinned blob and the offsets rebased.
    bool& all_vec = J.all_vec;
    all_vec = true;
    for (auto& r : reqs) {
        Ctx* x = r.c;
        if (int e = vstate_to_dev(x, ln.stream)) return e;
        x->stats = PassStats();
        if (x->bl % 4 != 0) all_vec = false;
        r.sincs = nullptr;
        if (r.resample) {
            r.sincs = sinc_table_dev(x, (double)r.render_sr / (double)x->sr);
            if (!r.sincs) return x->fail(TDG_ERR_CUDA, "could not build the sinc table");
        }
    }
    const double tr_vstate = now_ms();
    Frag* frags = g->frags->data();
    const size_t n_frags = reqs.size();
    std::vector<std::vector<VPlan>>& vplans = *g->vplans;
    // ---- 2b. clock tables: one per (sr, bl), deep enough for the longest-lived clock of this job
    J.clock_events.clear();
    {
        std::map<uint64_t, uint32_t> need;
        for (const Item& it : items) {
            const Ctx* x = reqs[it.ri].c;
            uint32_t& d = need[((uint64_t)x->sr << 32) | x->bl];
            d = std::max(d, std::max(vplans[it.ri][it.v].depth, 1u));
        }
        for (auto& kv : need) {
            ClockTab& tb = g->clocks[kv.first];
            const uint32_t sr = (uint32_t)(kv.first >> 32), bl = (uint32_t)kv.first;
            if (kv.second > tb.cap) {
                // growing past the allocation: earlier passes may still read the old table -> wait for the device
                // (once per engine and batch shape: the allocation carries 50 % headroom and at least 2048 rows)
                CU(cudaDeviceSynchronize());
                if (tb.p) cudaFree(tb.p);
                tb.p = nullptr; tb.depth = 0;
                tb.cap = std::max<uint32_t>(kv.second + kv.second / 2, 2048u);
                CU(cudaMalloc(&tb.p, (size_t)tb.cap * bl * 4));
                if (trace_on()) fprintf(stderr, "[tdg] clock table sr %u bl %u: %u rows (%.1f MB)\n", sr, bl, tb.cap, (double)tb.cap * bl * 4 / 1e6);
            }
            if (kv.second > tb.depth) {
                if (!tb.ev) CU(cudaEventCreateWithFlags(&tb.ev, cudaEventDisableTiming));
                const uint32_t upto = std::min(tb.cap, std::max(kv.second, tb.depth + 256u));          // rows are appended, never rewritten
                CU(launch_env_table(tb.p, bl, tb.depth, upto, (float)sr, ln.stream));
                CU(cudaEventRecord(tb.ev, ln.stream));
                tb.ev_set = true; tb.depth = upto;
            }
            if (tb.ev_set) J.clock_events.push_back(tb.ev);
        }
    }
    auto plan_one = [&](PassReq& r, Frag& F, std::vector<VPlan>& VPs) -> int {
        Ctx* x = r.c;
        F.reset((size_t)max_level + 1);
        HeapBlob& B = F.blob;
        std::vector<LevelTasks>& levels = F.levels;
        std::vector<QuantTask>& quants = F.quants; uint64_t& quant_max = F.quant_max; uint64_t& quant_bytes = F.quant_bytes;
        std::vector<ResampleTask>& resamps = F.resamps; uint64_t& resamp_max = F.resamp_max; uint64_t& resamp_bytes = F.resamp_bytes;
        const int V = (int)x->verts.size();
        const uint64_t pass_t0 = r.is_scan ? 0 : x->t;       // graph time of pass chunk 0 (true_normalize_scan: t = j*bl, graph.rs:228-231)
        const float* clock = nullptr;
        { auto it = g->clocks.find(((uint64_t)x->sr << 32) | x->bl); if (it != g->clocks.end()) clock = it->second.p; }
        auto range_task = [&](uint32_t kind, int v, uint64_t recs_off, size_t n_recs, int64_t len) -> const Range2* {
            RangeTask t; memset(&t, 0, sizeof t);
            t.kind = kind; t.n_chunks = (uint32_t)r.emit_n; t.chunk0 = (uint32_t)r.emit0; t.bl = x->bl;
            t.recs_off = recs_off; t.n_recs = (uint32_t)n_recs; t.len = len;
            t.out = reinterpret_cast<Range2*>(W + r.rng_off[v]);
            F.ranges.push_back(t); F.range_max = std::max(F.range_max, t.n_chunks);
            return t.out;
        };
        uint64_t algo = 0;
        const float* fused_runmax = nullptr; PanGain fused_pg{1.0f, 1.0f, 1.0f, 0};
        for (int v = 0; v < V; ++v) {
            if (r.level[v] < 0) continue;
            HVertex& hv = x->verts[v];
            if (r.skip[v]) { ++F.reused_v; F.reused_b += 8ull * (hv.edges.size() + 1) * r.n_frames; continue; }   // the scan's buffer stands
            if (r.inl[v] >= 0) { ++F.inlined_v; F.inlined_b += 16ull * r.n_frames; continue; }                  // evaluated inside its consumer: no write, no read back
            const uint64_t chunks_v = std::min<uint64_t>(r.prefix[v], r.emit_n);    // < emit_n: kept multi / lerp, head only
            const uint64_t frames_v = chunks_v * x->bl;
            LevelTasks& LT = levels[r.level[v]];
            TaskHdr h;
            h.out_l = reinterpret_cast<float*>(W + r.buf_off[v]);
            h.out_r = h.out_l + r.n_pad;
            h.n_frames = r.n_frames; h.n_chunks = (uint32_t)chunks_v; h.bl = x->bl;
            h.pg = make_pg(hv.gain, hv.angle);
            const uint32_t nch = (uint32_t)chunks_v;
            uint64_t in_off = 0, grp_off = 0;
            size_t n_pairs = hv.edges.size(), n_grp = 0;            // pointer pairs the kernel reads; groups (only with inlined inputs)
            if (hv.has_input()) {
                bool any_inl = false;
                for (int u : hv.edges) if (r.inl[u] == v) { any_inl = true; break; }
                if (any_inl) {
                    n_pairs = 0; n_grp = hv.edges.size();
                    for (int u : hv.edges) n_pairs += r.inl[u] == v ? x->verts[u].edges.size() : 1;
                }
                in_off = B.alloc(std::max<size_t>(n_pairs, 1) * 2 * sizeof(float*));
                if (any_inl) grp_off = B.alloc(n_grp * sizeof(FxGroup));
                if (in_off == UINT64_MAX || grp_off == UINT64_MAX) return x->fail(TDG_ERR_CUDA, "pinned plan buffer allocation failed");
                const float** ip = B.at<const float*>(in_off);
                FxGroup* gp = any_inl ? B.at<FxGroup>(grp_off) : nullptr;
                size_t at = 0;
                auto put = [&](int u) { const float* il = reinterpret_cast<const float*>(W + r.buf_off[u]); ip[2 * at] = il; ip[2 * at + 1] = il + r.n_pad; ++at; };
                for (size_t a = 0; a < hv.edges.size(); ++a) {
                    const int u = hv.edges[a];
                    if (gp) {
                        FxGroup& G = gp[a];
                        G.first = (uint32_t)at; G.nested = r.inl[u] == v ? 1u : 0u; G.pad = 0;
                        G.count = G.nested ? (uint32_t)x->verts[u].edges.size() : 1u;
                        G.pg = G.nested ? make_pg(x->verts[u].gain, x->verts[u].angle) : PanGain{1.0f, 1.0f, 1.0f, 0};
                    }
                    if (r.inl[u] == v) for (int w : x->verts[u].edges) put(w); else put(u);
                }
            }
            const uint64_t io_bytes = 8ull * (n_pairs + 1) * r.n_frames;
            auto fx_task = [&](uint32_t mode) {
                FxTask t; memset(&t, 0, sizeof t);
                t.h = h; t.mode = mode; t.n_in = (uint32_t)n_pairs; t.in_off = in_off; t.sr = x->sr;
                t.n_grp = (uint32_t)n_grp; t.grp_off = grp_off;
                if (n_pairs <= (size_t)kInlineInputs) memcpy(t.ins, B.at<const float*>(in_off), n_pairs * 2 * sizeof(float*));
                t.fsr = (float)x->sr; t.rsr = 1.0f / t.fsr;
                return t;
            };
            switch (hv.kind) {
                case K_SUM: { LT.fx.v.push_back(fx_task(FX_SUM)); LT.fx.max_chunks = std::max(LT.fx.max_chunks, nch); algo += io_bytes; LT.fx.bytes += io_bytes; break; }
                case K_NORMALIZE: {
                    FxTask t = fx_task(FX_NORM_PEAK);
                    float* peaks = reinterpret_cast<float*>(W + r.aux_off[v]);
                    float* runmax = peaks + r.emit_n;
                    t.peaks = peaks;
                    // Output Normalize without taps: the scan pass needs only its peaks (no buffer is consumed),
                    // the render pass scales inside the quantiser; with a render-time resampler the scaled
                    // buffer is still needed
                    const bool fuse_out = v == x->output && !x->taps && !r.resample && !r.play;
                    if (fuse_out && r.is_scan) t.no_store = 1;
                    LT.fx_norm.v.push_back(t); LT.fx_norm.max_chunks = std::max(LT.fx_norm.max_chunks, nch);
                    NormPrefixTask np{peaks, runmax, x->vstate_dev + (size_t)v * 8, nch, r.is_scan ? 1u : 0u, Xchg{nullptr, 0, 0, 1, 0}, nullptr};
                    if (r.x_ranks > 1) {
                        const int L = r.level[v];
                        np.x = Xchg{reinterpret_cast<float*>(W + r.x_off[L]), r.x_stride[L], r.x_index[v], (uint32_t)r.x_ranks, (uint32_t)r.x_rank};
                        if (r.x_mode == 2 && r.is_scan) np.reduced = reinterpret_cast<const float*>(W + r.x_red_off[L]);
                    }
                    LT.nprefix.v.push_back(np);
                    if (!fuse_out) {
                        NormScaleTask ns; ns.h = h; ns.runmax = runmax;
                        LT.nscale.v.push_back(ns); LT.nscale.max_chunks = std::max(LT.nscale.max_chunks, nch);
                        LT.nscale.bytes += 16ull * r.n_frames;
                    } else { fused_runmax = runmax; fused_pg = h.pg; }
                    // the pass's algorithmic bytes follow SURVEY.md §8(d): every vertex reads each input once and
                    // writes its output once.  The in-place scale of a non-output Normalize (k_norm_scale) is this
                    // design's own extra traffic: it is counted for that kernel's own roofline line, not here
                    algo += io_bytes;
                    LT.fx_norm.bytes += io_bytes - (t.no_store ? 8ull * r.n_frames : 0); LT.nprefix.bytes += 8ull * r.emit_n;
                    break;
                }
                case K_ADSR: {
                    if (hv.wet < 0.0001f) {                                              // extensions.rs:598
                        LT.fx.v.push_back(fx_task(FX_SUM)); LT.fx.max_chunks = std::max(LT.fx.max_chunks, nch); LT.fx.bytes += io_bytes;
                    } else {
                        VPlan& VP = VPs[v];
                        FxTask t = fx_task(FX_ADSR);
                        t.conf = make_adsr_k(hv.adsr); t.wet = hv.wet; t.use_off = hv.use_off; t.use_max = hv.use_max;
                        if (t.conf.exotic || adsr_duration_exotic(t.fsr)) LT.fx_adsr_exact = true;
                        const uint64_t recs_off = B.alloc(VP.arecs.size() * sizeof(AdsrRec));
                        memcpy(B.at<uint8_t>(recs_off), VP.arecs.data(), VP.arecs.size() * sizeof(AdsrRec));
                        const uint64_t xenv_off = B.alloc(std::max<size_t>(VP.xenv.size(), 1) * 4);
                        memcpy(B.at<uint8_t>(xenv_off), VP.xenv.data(), VP.xenv.size() * 4);
                        {
                            uint8_t* sb = W + r.seg_off[v];
                            uint32_t* seg_ptr = reinterpret_cast<uint32_t*>(sb);
                            F4* segs = reinterpret_cast<F4*>(sb + align_up((r.emit_n + 1) * 4, 16));
                            (void)range_task(RK_ADSR, v, recs_off, VP.arecs.size(), 0);
                            RangeTask& rt = F.ranges.back();
                            rt.seg_ptr = seg_ptr; rt.segs = segs; rt.relv2 = nullptr; rt.clock = clock; rt.xenv_off = xenv_off;
                            F.any_voice = true;
                            t.seg_ptr = seg_ptr; t.segs = reinterpret_cast<const AdsrSeg*>(segs);
                        }
                        LT.fx_adsr.v.push_back(t); LT.fx_adsr.max_chunks = std::max(LT.fx_adsr.max_chunks, nch);
                        LT.fx_adsr.bytes += io_bytes + VPs[v].arecs.size() * sizeof(AdsrRec);
                    }
                    algo += io_bytes;
                    break;
                }
                case K_BANDPASS: {
                    if (hv.wet < 0.0001f || (hv.lgamma == 0.0f && hv.hgamma == 0.0f)) {   // extensions.rs:657-658
                        LT.fx.v.push_back(fx_task(FX_SUM)); LT.fx.max_chunks = std::max(LT.fx.max_chunks, nch);
                        algo += io_bytes; LT.fx.bytes += io_bytes;
                        break;
                    }
                    BpTask t; memset(&t, 0, sizeof t);
                    t.h = h; t.n_in = (uint32_t)hv.edges.size(); t.in_off = in_off; t.pass = hv.pass;
                    if (hv.edges.size() <= (size_t)kInlineInputs) memcpy(t.ins, B.at<const float*>(in_off), hv.edges.size() * 2 * sizeof(float*));
                    t.lgamma = hv.lgamma; t.hgamma = hv.hgamma;
                    // a segment that does not start at the pass start has no true state: seed from its first frame
                    t.first = (hv.bp_first || r.emit0 > 0) ? 1u : 0u;
                    t.n_tiles = (uint32_t)((r.n_frames + kBpTile - 1) / kBpTile);
                    t.state = x->vstate_dev + (size_t)v * 8 + 2;
                    t.agg = reinterpret_cast<float*>(W + r.aux_off[v]);
                    t.carry = t.agg + (size_t)(r.n_pad / kBpTile) * 8;
                    t.sync = reinterpret_cast<uint32_t*>(W + J.sync_off + r.bps_off[v]);
                    t.x = Xchg{nullptr, 0, 0, 1, 0};
                    if (r.x_ranks > 1) {
                        const int L = r.level[v];
                        t.x = Xchg{reinterpret_cast<float*>(W + r.x_off[L]), r.x_stride[L], r.x_index[v], (uint32_t)r.x_ranks, (uint32_t)r.x_rank};
                    }
                    LT.bp.v.push_back(t); LT.bp.max_chunks = std::max(LT.bp.max_chunks, t.n_tiles);
                    hv.bp_first = false;
                    algo += io_bytes;      // §8(d) model; the aggregate pass's second read of the inputs is design overhead
                    LT.bp.bytes += io_bytes; LT.bp_agg_bytes += 8ull * hv.edges.size() * r.n_frames;
                    break;
                }
                case K_SAMPLE_LOOP: {
                    const DevSample& s = x->samples[hv.sample];
                    LoopTask t; t.h = h; t.sl = s.l; t.sr = s.r; t.len = s.frames; t.t0 = hv.loop_t + r.emit0 * x->bl;
                    LT.loop.v.push_back(t); LT.loop.max_chunks = std::max(LT.loop.max_chunks, nch);
                    hv.loop_t += r.total * x->bl;                                 // extensions.rs:340
                    algo += 16ull * r.n_frames; LT.loop.bytes += 16ull * r.n_frames;
                    break;
                }
                case K_SAMPLE_MULTI: {
                    const DevSample& s = x->samples[hv.sample];
                    std::vector<MultiHit>& mhits = VPs[v].mhits;
                    if (r.is_scan && r.keep_write && x->keep.off[v] >= 0) {
                        // the render differs from this scan only while hits carried INTO either pass still play:
                        // carried entries come first in `mhits` (start < 0), the ones carried out are in hv.ts now
                        int64_t last = 0;
                        for (const MultiHit& mh : mhits) { if (mh.start >= 0) break; last = std::max<int64_t>(last, mh.start + (int64_t)s.frames); }
                        for (const MultiEntry& e : hv.ts) last = std::max<int64_t>(last, (int64_t)s.frames - e.t);
                        x->keep.hz[v] = ((uint64_t)last + x->bl - 1) / x->bl;
                    }
                    if (chunks_v == 0) { ++F.reused_v; F.reused_b += 8ull * r.n_frames; break; }       // nothing leaked: the scan's buffer stands
                    if (chunks_v < r.emit_n) { ++F.reused_v; F.reused_b += 8ull * (r.n_frames - frames_v); }
                    MultiTask t; memset(&t, 0, sizeof t);
                    t.h = h; t.sl = s.l; t.sr = s.r; t.len = s.frames; t.n_hits = (uint32_t)mhits.size();
                    t.hits_off = B.alloc(std::max<size_t>(mhits.size(), 1) * sizeof(MultiHit));
                    memcpy(B.at<uint8_t>(t.hits_off), mhits.data(), mhits.size() * sizeof(MultiHit));
                    // per chunk: hits with start in (p0 - L, p0 + bl - 1], found on the device
                    t.ranges = range_task(RK_MULTI, v, t.hits_off, mhits.size(), (int64_t)s.frames);
                    LT.multi.v.push_back(t); LT.multi.max_chunks = std::max(LT.multi.max_chunks, nch);
                    uint64_t gathered = 0;                                        // live hit-frames inside the emitted range
                    for (const MultiHit& mh : mhits) {
                        int64_t a = std::max<int64_t>(mh.start, 0), b = std::min<int64_t>(mh.start + (int64_t)s.frames, (int64_t)frames_v);
                        if (b > a) gathered += (uint64_t)(b - a);
                    }
                    algo += 8ull * frames_v + 8ull * gathered; LT.multi.bytes += 8ull * frames_v + 8ull * gathered;
                    break;
                }
                case K_SAMPLE_LERP: {
                    const DevSample& s = x->samples[hv.sample];
                    std::vector<LerpHit>& lhits = VPs[v].lhits;
                    if (r.is_scan && r.keep_write && x->keep.off[v] >= 0)
                        // both passes play the same hits; only the pair carried in differs, and it is audible until the
                        // first hit of the pass has faded in (entries 0 and 1 are the carried pair)
                        x->keep.hz[v] = lhits.size() > 2 ? ((uint64_t)lhits[2].frame + hv.lerp_len + x->bl - 1) / x->bl : UINT64_MAX;
                    if (chunks_v == 0) { ++F.reused_v; F.reused_b += 16ull * r.n_frames; break; }
                    if (chunks_v < r.emit_n) { ++F.reused_v; F.reused_b += 16ull * (r.n_frames - frames_v); }
                    LerpTask t; memset(&t, 0, sizeof t);
                    t.h = h; t.sl = s.l; t.sr = s.r; t.len = s.frames; t.n_hits = (uint32_t)lhits.size(); t.lerp_len = hv.lerp_len;
                    t.hits_off = B.alloc(lhits.size() * sizeof(LerpHit));
                    memcpy(B.at<uint8_t>(t.hits_off), lhits.data(), lhits.size() * sizeof(LerpHit));
                    // per chunk: last hit with frame <= first frame of the chunk, found on the device
                    t.ranges = range_task(RK_LERP, v, t.hits_off, lhits.size(), 0);
                    LT.lerp.v.push_back(t); LT.lerp.max_chunks = std::max(LT.lerp.max_chunks, nch);
                    algo += 16ull * frames_v; LT.lerp.bytes += 16ull * frames_v;
                    break;
                }
                case K_DEBUG_SINE: case K_SYNTH: case K_SAMPSYN: {
                    VPlan& VP = VPs[v];
                    VoiceTask t; memset(&t, 0, sizeof t);
                    t.h = h; t.sr = x->sr; t.t0 = pass_t0 + r.emit0 * x->bl;
                    t.fsr = (float)x->sr; t.rsr = 1.0f / t.fsr;
                    bool exotic = adsr_duration_exotic(t.fsr);
                    // the records go up; the per-chunk segment table is built from them in device scratch
                    const uint64_t recs_off = B.alloc(std::max<size_t>(VP.vrecs.size(), 1) * sizeof(VoiceRec));
                    memcpy(B.at<uint8_t>(recs_off), VP.vrecs.data(), VP.vrecs.size() * sizeof(VoiceRec));
                    const uint64_t xenv_off = B.alloc(std::max<size_t>(VP.xenv.size(), 1) * 4);
                    memcpy(B.at<uint8_t>(xenv_off), VP.xenv.data(), VP.xenv.size() * 4);
                    {
                        uint8_t* sb = W + r.seg_off[v];
                        uint32_t* seg_ptr = reinterpret_cast<uint32_t*>(sb);
                        F4* segs = reinterpret_cast<F4*>(sb + align_up((r.emit_n + 1) * 4, 16));
                        float* relv2 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(segs) + r.seg_cap[v] * 32);
                        (void)range_task(RK_VOICE, v, recs_off, VP.vrecs.size(), 0);
                        RangeTask& rt = F.ranges.back();
                        rt.seg_ptr = seg_ptr; rt.segs = segs; rt.relv2 = relv2; rt.clock = clock; rt.xenv_off = xenv_off;
                        rt.debug_sine = hv.kind == K_DEBUG_SINE ? 1u : 0u;
                        F.any_voice = true;
                        t.seg_ptr = seg_ptr; t.segs = segs; t.relv2 = relv2;
                    }
                    x->stats.voice_frames += VP.voice_frames;
                    const uint64_t vbytes = 8ull * r.n_frames + VP.vrecs.size() * sizeof(VoiceRec);
                    const uint32_t nblk = nch;
                    if (hv.kind == K_DEBUG_SINE) { t.kind = VOICE_DEBUG_SINE; if (exotic) LT.dsine_exact = true; LT.dsine.v.push_back(t); LT.dsine.max_chunks = std::max(LT.dsine.max_chunks, nblk); LT.dsine.bytes += vbytes; }
                    else if (hv.kind == K_SYNTH) {
                        t.kind = VOICE_SYNTH;
                        t.vol[0] = hv.sq.volume; t.sq_z = hv.sq.param; t.vol[1] = hv.tf.volume; t.tf_z = hv.tf.param;
                        t.vol[2] = hv.tr.volume;
                        t.sq_iz = 1.0f / t.sq_z;                                             // synth.rs:23
                        t.tf_add = (1.0f - t.tf_z) / 2.0f; t.tf_mul = 2.0f / (1.0f + t.tf_z);  // synth.rs:28
                        t.env[0] = make_adsr_k(hv.sq.adsr); t.env[1] = make_adsr_k(hv.tf.adsr); t.env[2] = make_adsr_k(hv.tr.adsr);
                        t.env12_same = memcmp(&hv.tf.adsr, &hv.tr.adsr, sizeof(AdsrConfD)) == 0 ? 1u : 0u;
                        if (exotic || (t.vol[0] > 0.0f && t.env[0].exotic) || (t.vol[1] > 0.0f && t.env[1].exotic) ||
                            (t.vol[2] > 0.0f && t.env[2].exotic)) LT.synth_exact = true;
                        t.amp_mul = 1.0f / (hv.sq.volume * m_max_vel(hv.sq.adsr) + hv.tf.volume * m_max_vel(hv.tf.adsr) +
                                            hv.tr.volume * m_max_vel(hv.tr.adsr));           // extensions.rs:465-468
                        LT.synth.v.push_back(t); LT.synth.max_chunks = std::max(LT.synth.max_chunks, nblk); LT.synth.bytes += vbytes;
                    } else {
                        const DevTable& tb = x->tables[hv.table];
                        t.kind = VOICE_SAMPSYN; t.env[0] = make_adsr_k(hv.adsr); t.amp_mul = 1.0f / m_max_vel(hv.adsr);   // extensions.rs:537
                        if (exotic || t.env[0].exotic) LT.sampsyn_exact = true;
                        t.table = tb.data; t.n_frames = tb.n_frames; t.n_phase = tb.n_phase; t.fps = tb.fps;
                        t.table_floats = tb.n_frames * tb.n_phase;
                        LT.sampsyn.v.push_back(t); LT.sampsyn.max_chunks = std::max(LT.sampsyn.max_chunks, nblk); LT.sampsyn.bytes += vbytes;
                    }
                    algo += 8ull * r.n_frames;
                    break;
                }
            }
        }
        // output stage (state.rs:515-574); the playback feed hands out the output vertex's f32 buffer instead
        if (!r.is_scan && x->output >= 0 && !r.play) {
            const float* ol = reinterpret_cast<const float*>(W + r.buf_off[x->output]);
            const float* orr = ol + r.n_pad;
            QuantTask q; memset(&q, 0, sizeof q);
            q.bd = r.bd;
            q.amp = r.bd < 32 ? (float)((1 << (r.bd - 1)) - 1) : (float)INT32_MAX;
            q.out = P + r.pcm_off; q.err = reinterpret_cast<uint32_t*>(P + r.err_off);
            if (r.resample) {
                ResampleTask rt; memset(&rt, 0, sizeof rt);
                rt.in_l = ol; rt.in_r = orr;
                rt.out_l = reinterpret_cast<float*>(W + r.rs_off); rt.out_r = rt.out_l + align_up(r.rs_total, 4);
                rt.sincs = r.sincs; rt.idx = reinterpret_cast<double*>(W + r.rs_idx_off);
                const uint64_t before = (r.keep0 - r.emit0) * x->bl;
                rt.in_l += before; rt.in_r += before; rt.n_before = before;
                rt.n_in = r.n_frames - before; rt.n_chunks = (uint32_t)r.rs_chunks.size(); rt.chunk = x->bl;
                if (r.halo_off >= 0 && r.x_rank > 0 && before < (uint64_t)kResampleHalo) {
                    rt.halo_l = reinterpret_cast<const float*>(W + r.halo_off) + (size_t)(r.x_rank - 1) * 2 * kResampleHalo;
                    rt.halo_r = rt.halo_l + kResampleHalo;
                }
                rt.t_ratio = 1.0 / ((double)r.render_sr / (double)x->sr);
                rt.chunk_off = B.alloc(std::max<size_t>(r.rs_chunks.size(), 1) * sizeof(ResampleChunk));
                memcpy(B.at<uint8_t>(rt.chunk_off), r.rs_chunks.data(), r.rs_chunks.size() * sizeof(ResampleChunk));
                resamps.push_back(rt); resamp_max = std::max(resamp_max, r.rs_total); F.resamp_chunks = std::max<uint32_t>(F.resamp_chunks, rt.n_chunks);
                q.in_l = rt.out_l; q.in_r = rt.out_r; q.n = r.rs_total;
                algo += 8ull * r.n_frames + 8ull * r.rs_total; resamp_bytes += 8ull * r.n_frames + 8ull * r.rs_total;
            } else {
                const uint64_t skip = (r.keep0 - r.emit0) * x->bl;
                q.in_l = ol + skip; q.in_r = orr + skip; q.n = r.out_frames;
                q.runmax = fused_runmax; q.frame0 = skip; q.bl = x->bl; q.pg = fused_pg;
            }
            q.span = q.runmax ? x->bl : 1024u; q.chunk0 = q.runmax ? (uint32_t)(q.frame0 / x->bl) : 0u;
            quants.push_back(q); quant_max = std::max<uint64_t>(quant_max, (q.n + q.span - 1) / q.span);   // blocks
            algo += 8ull * q.n + r.pcm_bytes; quant_bytes += 8ull * q.n + r.pcm_bytes;
        }
        x->stats.algorithmic_bytes = algo; x->stats.frames = r.n_frames;
        F.algo = algo;
        x->last_buf_off = r.buf_off; x->last_n_frames = r.n_frames; x->last_n_pad = r.n_pad; x->last_epoch = g->epoch; x->last_lane = (int)(J.lane - g->lanes);
        return TDG_OK;
    };
    {
        std::atomic<int> first_err{TDG_OK};
        g->pool->parallel_for(reqs.size(), 32, [&](size_t i, unsigned w) {
            (void)w;
            const int e = plan_one(reqs[i], frags[i], vplans[i]);
            if (e != TDG_OK) { int z = TDG_OK; first_err.compare_exchange_strong(z, e); }
        });
        if (first_err.load() != TDG_OK) return first_err.load();
    }
    const double tr_planned = now_ms();
    // merge: reserve the pinned blob, rebase offsets, concatenate the task lists (project order)
    Slot& sl = ln.slots[J.slot];
    PinnedBuf& B = sl.blob;
    B.size = 0;
    std::vector<LevelTasks>& levels = J.levels;
    levels.assign(max_level + 1, LevelTasks());
    std::vector<QuantTask>& quants = J.quants; uint64_t &quant_max = J.quant_max, &quant_bytes = J.quant_bytes;
    std::vector<ResampleTask>& resamps = J.resamps; uint64_t &resamp_max = J.resamp_max, &resamp_bytes = J.resamp_bytes;
    uint64_t& algo_bytes_total = J.algo_bytes_total;
    quants.clear(); resamps.clear(); quant_max = quant_bytes = resamp_max = resamp_bytes = algo_bytes_total = 0;
    J.reused_v = J.reused_b = J.inlined_v = J.inlined_b = 0; J.ranges.clear(); J.range_max = 0; J.resamp_chunks = 0; J.any_voice = false;
    {
        uint64_t tot = 0;
        for (size_t i = 0; i < n_frags; ++i) { Frag& F = frags[i]; F.base = align_up(tot, 256); tot = F.base + F.blob.v.size(); }
        g->hwm_blob = std::max<uint64_t>(g->hwm_blob, tot + 256);
        if (!B.reserve(tot + 256, g->hwm_blob)) return c->fail(TDG_ERR_CUDA, "pinned plan buffer allocation failed");
        B.size = tot;
        g->pool->parallel_for(n_frags, tot < (8u << 20) ? 1 : 16, [&](size_t i, unsigned) {
            if (!frags[i].blob.v.empty()) memcpy(B.p + frags[i].base, frags[i].blob.v.data(), frags[i].blob.v.size());
        });
        for (size_t fi = 0; fi < n_frags; ++fi) {
            Frag& F = frags[fi];
            const uint64_t b = F.base;
            for (size_t L = 0; L < F.levels.size(); ++L) {
                LevelTasks& S = F.levels[L]; LevelTasks& Dst = levels[L];
                for (FxTask& t : S.fx.v) { t.in_off += b; t.grp_off += b; }
                for (FxTask& t : S.fx_adsr.v) { t.in_off += b; t.grp_off += b; }
                for (FxTask& t : S.fx_norm.v) { t.in_off += b; t.grp_off += b; }
                for (MultiTask& t : S.multi.v) t.hits_off += b;
                for (LerpTask& t : S.lerp.v) t.hits_off += b;
                for (BpTask& t : S.bp.v) t.in_off += b;
                append_list(Dst.fx, S.fx); append_list(Dst.fx_adsr, S.fx_adsr); append_list(Dst.fx_norm, S.fx_norm);
                append_list(Dst.nprefix, S.nprefix); append_list(Dst.nscale, S.nscale);
                append_list(Dst.loop, S.loop); append_list(Dst.multi, S.multi); append_list(Dst.lerp, S.lerp);
                append_list(Dst.dsine, S.dsine); append_list(Dst.synth, S.synth); append_list(Dst.sampsyn, S.sampsyn);
                append_list(Dst.bp, S.bp);
                Dst.fx_adsr_exact |= S.fx_adsr_exact; Dst.dsine_exact |= S.dsine_exact;
                Dst.synth_exact |= S.synth_exact; Dst.sampsyn_exact |= S.sampsyn_exact;
                Dst.bp_agg_bytes += S.bp_agg_bytes;
            }
            for (RangeTask& t : F.ranges) { t.recs_off += b; t.xenv_off += b; J.ranges.push_back(t); }
            J.any_voice = J.any_voice || F.any_voice;
            J.range_max = std::max(J.range_max, F.range_max); J.resamp_chunks = std::max(J.resamp_chunks, F.resamp_chunks);
            for (ResampleTask& t : F.resamps) { t.chunk_off += b; resamps.push_back(t); }
            for (QuantTask& t : F.quants) quants.push_back(t);
            quant_max = std::max(quant_max, F.quant_max); quant_bytes += F.quant_bytes;
            resamp_max = std::max(resamp_max, F.resamp_max); resamp_bytes += F.resamp_bytes;
            algo_bytes_total += F.algo; J.reused_v += F.reused_v; J.reused_b += F.reused_b; J.inlined_v += F.inlined_v; J.inlined_b += F.inlined_b;
        }
    }
    // task arrays go into the blob too: one H2D for the whole pass
    std::vector<Off>& offs = J.offs;
    offs.assign(levels.size(), Off());
    auto put = [&](const void* src, size_t bytes) -> uint64_t {
        uint64_t o = B.alloc(std::max<size_t>(bytes, 16));
        if (o != UINT64_MAX && bytes) memcpy(B.p + o, src, bytes);
        return o;
    };
    for (size_t L = 0; L < levels.size(); ++L) {
        LevelTasks& T = levels[L];
        // tasks with inlined inputs go to the end of their list: they are launched with the grouped kernel variant
        for (TaskList<FxTask>* tl : {&T.fx, &T.fx_adsr, &T.fx_norm})
            tl->n_plain = (size_t)(std::stable_partition(tl->v.begin(), tl->v.end(), [](const FxTask& t) { return t.n_grp == 0; }) - tl->v.begin());
        offs[L].fx = put(T.fx.v.data(), T.fx.v.size() * sizeof(FxTask));
        offs[L].fx_adsr = put(T.fx_adsr.v.data(), T.fx_adsr.v.size() * sizeof(FxTask));
        offs[L].fx_norm = put(T.fx_norm.v.data(), T.fx_norm.v.size() * sizeof(FxTask));
        offs[L].nprefix = put(T.nprefix.v.data(), T.nprefix.v.size() * sizeof(NormPrefixTask));
        offs[L].nscale = put(T.nscale.v.data(), T.nscale.v.size() * sizeof(NormScaleTask));
        offs[L].loop = put(T.loop.v.data(), T.loop.v.size() * sizeof(LoopTask));
        offs[L].multi = put(T.multi.v.data(), T.multi.v.size() * sizeof(MultiTask));
        offs[L].lerp = put(T.lerp.v.data(), T.lerp.v.size() * sizeof(LerpTask));
        offs[L].dsine = put(T.dsine.v.data(), T.dsine.v.size() * sizeof(VoiceTask));
        offs[L].synth = put(T.synth.v.data(), T.synth.v.size() * sizeof(VoiceTask));
        offs[L].sampsyn = put(T.sampsyn.v.data(), T.sampsyn.v.size() * sizeof(VoiceTask));
        offs[L].bp = put(T.bp.v.data(), T.bp.v.size() * sizeof(BpTask));
    }
    J.quant_off = put(quants.data(), quants.size() * sizeof(QuantTask));
    J.resamp_off = put(resamps.data(), resamps.size() * sizeof(ResampleTask));
    J.range_off = put(J.ranges.data(), J.ranges.size() * sizeof(RangeTask));
    if (J.quant_off == UINT64_MAX || J.resamp_off == UINT64_MAX || J.range_off == UINT64_MAX) return c->fail(TDG_ERR_CUDA, "pinned plan buffer allocation failed");
    const double tr_merged = now_ms();
    g->hwm_blob = std::max<uint64_t>(g->hwm_blob, B.size + 256);
    CU(sl.blob_dev.ensure(B.size + 256, g->hwm_blob));
    if (!sl.ensure_err(reqs.size())) return c->fail(TDG_ERR_CUDA, "pinned flag buffer allocation failed");
    J.t_planned = now_ms();
    if (trace_on())
        fprintf(stderr, "[tdg] plan lane %d reqs %zu: layout %.2f arena %.2f vstate+clocks %.2f events %.2f (%zu vertices) tasks %.2f merge %.2f blobdev %.2f ms (blob %.1f MB)\n",
                (int)(J.lane - g->lanes), reqs.size(), tr_layout0 - J.t_start, tr_arena - tr_layout, tr_vstate - tr_arena,
                tr_layout - tr_layout0, items.size(), tr_planned - tr_vstate, tr_merged - tr_planned, J.t_planned - tr_merged, B.size / 1e6);
    return TDG_OK;
}

static int pass_submit(PassJob& J) {
    std::vector<PassReq>& reqs = J.reqs;
    if (reqs.empty()) return TDG_OK;
    Lane& ln = *J.lane;
    Ctx* c = reqs[0].c;
    Engine* g = c->eng;
    cudaSetDevice(g->device);
    Slot& sl = ln.slots[J.slot];
    PinnedBuf& B = sl.blob;
    uint8_t* W = J.W;
    std::vector<LevelTasks>& levels = J.levels;
    std::vector<Off>& offs = J.offs;
    std::vector<QuantTask>& quants = J.quants; std::vector<ResampleTask>& resamps = J.resamps;
    const uint64_t quant_max = J.quant_max, quant_bytes = J.quant_bytes, resamp_max = J.resamp_max, resamp_bytes = J.resamp_bytes;
    const uint64_t quant_off = J.quant_off, resamp_off = J.resamp_off;
    const bool all_vec = J.all_vec;
    // ---- 3. upload + launches ------------------------------------------------------------------
    const uint8_t* D = sl.blob_dev.p;
    cudaStream_t s = ln.stream;
    // the blob goes up on the upload stream (it is ready long before the lane reaches this pass); the
    // slot's previous pass is complete (host-synchronised in run_jobs_pipelined), so the device blob is free
    CU(cudaMemcpyAsync(sl.blob_dev.p, B.p, B.size, cudaMemcpyHostToDevice, g->h2d_stream));
    CU(cudaEventRecord(sl.ev_h2d, g->h2d_stream));
    CU(cudaStreamWaitEvent(s, sl.ev_h2d, 0));
    J.h2d_bytes = B.size;
    if (J.err_bytes) CU(cudaMemsetAsync(J.P + J.err_base, 0, J.err_bytes, s));      // TooWide flags of every request, contiguous
    for (cudaEvent_t e : J.clock_events) CU(cudaStreamWaitEvent(s, e, 0));          // clock rows appended on another lane's stream
    if (J.sync_bytes) CU(cudaMemsetAsync(W + J.sync_off, 0, J.sync_bytes, s));       // tickets and flags of the fused band-pass kernels
    J.t_h2d_issued = now_ms();
    CU(cudaEventRecord(sl.ev0, s));
    uint64_t &launches = J.launches, &collectives = J.collectives;
    launches = collectives = 0;
    const int x_mode = reqs[0].x_mode;
    for (auto& r : reqs) if (r.x_mode != x_mode || (x_mode == 2 && reqs.size() != 1))
        return r.c->fail(TDG_ERR_INVALID, "time-sharded and plain passes cannot share a launch sequence");
    const uint32_t kMaxY = 65535;          // gridDim.y limit: tasks per launch
    std::vector<Timed>& timed = J.timed;
    timed.clear();
    auto ev_at = [&](size_t i) -> cudaEvent_t {
        while (sl.evpool.size() <= i) { cudaEvent_t e; cudaEventCreate(&e); sl.evpool.push_back(e); }
        return sl.evpool[i];
    };
    // per-chunk index ranges of every event-driven vertex of the pass: one launch, before the first level
    if (!J.ranges.empty()) {
        const size_t kMaxTasks = 65535;
        for (size_t b0 = 0; b0 < J.ranges.size(); b0 += kMaxTasks) {
            const uint32_t nb = (uint32_t)std::min(kMaxTasks, J.ranges.size() - b0);
            if (g->profiling) CU(cudaEventRecord(ev_at(2 * timed.size()), s));
            CU(launch_chunk_ranges(reinterpret_cast<const RangeTask*>(D + J.range_off) + b0, nb, J.range_max, J.any_voice, D, s));
            launches += J.any_voice ? 3 : 1;
            if (g->profiling) { CU(cudaEventRecord(ev_at(2 * timed.size() + 1), s)); timed.push_back({KC_RANGES, 0}); }
        }
    }
#define LAUNCH_SPLIT(cls, list, T, off, bytes_, expr) \
    for (size_t b0 = 0; b0 < (list).v.size(); b0 += kMaxY) { \
        const uint32_t nb = (uint32_t)std::min<size_t>(kMaxY, (list).v.size() - b0); \
        const T* tp = reinterpret_cast<const T*>(D + (off)) + b0; \
        if (g->profiling) CU(cudaEventRecord(ev_at(2 * timed.size()), s)); \
        CU(expr); ++launches; \
        if (g->profiling) { CU(cudaEventRecord(ev_at(2 * timed.size() + 1), s)); \
                            timed.push_back({cls, (bytes_) * nb / std::max<size_t>((list).v.size(), 1)}); } }
    for (size_t L = 0; L < levels.size(); ++L) {
        LevelTasks& T = levels[L];
        const bool vec = all_vec;
        LAUNCH_SPLIT(KC_LOOP, T.loop, LoopTask, offs[L].loop, T.loop.bytes, launch_sample_loop(tp, nb, T.loop.max_chunks, vec, s));
        LAUNCH_SPLIT(KC_MULTI, T.multi, MultiTask, offs[L].multi, T.multi.bytes, launch_sample_multi(tp, nb, T.multi.max_chunks, vec, D, s));
        LAUNCH_SPLIT(KC_LERP, T.lerp, LerpTask, offs[L].lerp, T.lerp.bytes, launch_sample_lerp(tp, nb, T.lerp.max_chunks, vec, D, s));
        LAUNCH_SPLIT(KC_DSINE, T.dsine, VoiceTask, offs[L].dsine, T.dsine.bytes, launch_voice(VOICE_DEBUG_SINE, tp, nb, T.dsine.max_chunks, T.dsine_exact, g->libm_sin, D, s));
        LAUNCH_SPLIT(KC_SYNTH, T.synth, VoiceTask, offs[L].synth, T.synth.bytes, launch_voice(VOICE_SYNTH, tp, nb, T.synth.max_chunks, T.synth_exact, g->libm_sin, D, s));
        LAUNCH_SPLIT(KC_SAMPSYN, T.sampsyn, VoiceTask, offs[L].sampsyn, T.sampsyn.bytes,
                     launch_voice(VOICE_SAMPSYN, tp, nb, T.sampsyn.max_chunks, T.sampsyn_exact, false, D, s));
        // k_fx lists: plain tasks first, then (other kernel instantiation) the tasks that evaluate inlined Sum inputs
#define LAUNCH_FX(cls, list, off, mode, exact) \
        for (int grouped = 0; grouped < 2; ++grouped) { \
            const size_t q0 = grouped ? std::min((list).n_plain, (list).v.size()) : 0, q1 = grouped ? (list).v.size() : std::min((list).n_plain, (list).v.size()); \
            for (size_t b0 = q0; b0 < q1; b0 += kMaxY) { \
                const uint32_t nb = (uint32_t)std::min<size_t>(kMaxY, q1 - b0); \
                const FxTask* tp = reinterpret_cast<const FxTask*>(D + (off)) + b0; \
                if (g->profiling) CU(cudaEventRecord(ev_at(2 * timed.size()), s)); \
                CU(launch_fx(mode, tp, nb, (list).max_chunks, vec, exact, grouped != 0, D, s)); ++launches; \
                if (g->profiling) { CU(cudaEventRecord(ev_at(2 * timed.size() + 1), s)); \
                                    timed.push_back({cls, (list).bytes * nb / std::max<size_t>((list).v.size(), 1)}); } } }
        LAUNCH_FX(KC_FX_SUM, T.fx, offs[L].fx, FX_SUM, false);
        LAUNCH_FX(KC_FX_ADSR, T.fx_adsr, offs[L].fx_adsr, FX_ADSR, T.fx_adsr_exact);
        LAUNCH_FX(KC_FX_NORM, T.fx_norm, offs[L].fx_norm, FX_NORM_PEAK, false);
        // band pass = aggregate, carry, apply over the same inputs.  Without an exchange step in between the three
        // launches run over groups of filters whose inputs + outputs fit the L2 (bp_l2_bytes): the apply pass then
        // finds the inputs the aggregate pass just streamed in cache instead of fetching them from HBM again
        // no exchange step between aggregate and apply, pass short enough for the serial group chain: ONE launch
        const bool bp_fused = !x_mode && g->bp_fused && !T.bp.v.empty() && T.bp.max_chunks <= g->bp_fused_max_tiles;
        const bool bp_grouped = !bp_fused && !x_mode && g->bp_l2_bytes > 0 && T.bp.v.size() > 1;
        if (bp_fused) {
            LAUNCH_SPLIT(KC_BP_FUSED, T.bp, BpTask, offs[L].bp, T.bp.bytes, launch_bp_fused(tp, nb, T.bp.max_chunks, vec, D, s));
        } else if (bp_grouped) {
            const size_t n = T.bp.v.size();
            auto io_of = [&](size_t q) { return 8ull * (T.bp.v[q].n_in + 1) * T.bp.v[q].h.n_frames; };
            for (size_t b0 = 0; b0 < n;) {
                size_t b1 = b0; uint64_t acc = 0, in_b = 0;
                do { acc += io_of(b1); in_b += 8ull * T.bp.v[b1].n_in * T.bp.v[b1].h.n_frames; ++b1; }
                while (b1 < n && b1 - b0 < kMaxY && acc + io_of(b1) <= g->bp_l2_bytes);
                const uint32_t nb = (uint32_t)(b1 - b0);
                const BpTask* tp = reinterpret_cast<const BpTask*>(D + offs[L].bp) + b0;
                auto timed_launch = [&](int cls, uint64_t bytes, cudaError_t (*fn)(const BpTask*, uint32_t, uint32_t, bool, const uint8_t*, cudaStream_t)) -> cudaError_t {
                    if (g->profiling) cudaEventRecord(ev_at(2 * timed.size()), s);
                    const cudaError_t e = fn(tp, nb, T.bp.max_chunks, vec, D, s); ++launches;
                    if (g->profiling) { cudaEventRecord(ev_at(2 * timed.size() + 1), s); timed.push_back({cls, bytes}); }
                    return e;
                };
                CU(timed_launch(KC_BP_AGG, in_b, launch_bp_aggregate));
                if (g->profiling) CU(cudaEventRecord(ev_at(2 * timed.size()), s));
                CU(launch_bp_carry(tp, nb, D, s)); ++launches;
                if (g->profiling) { CU(cudaEventRecord(ev_at(2 * timed.size() + 1), s)); timed.push_back({KC_BP_CARRY, 0}); }
                CU(timed_launch(KC_BP_APPLY, acc, launch_bp_apply));
                b0 = b1;
            }
        } else {
        LAUNCH_SPLIT(KC_BP_AGG, T.bp, BpTask, offs[L].bp, T.bp_agg_bytes, launch_bp_aggregate(tp, nb, T.bp.max_chunks, vec, D, s));
        }
        if (x_mode && (!T.nprefix.v.empty() || !T.bp.v.empty())) {
            // time sharding: publish this segment's peaks / filter maps, then fetch the other ranks'
            const PassReq& r0 = reqs[0];
            float* red = (x_mode == 2 && r0.is_scan && r0.x_red_off[L] >= 0) ? reinterpret_cast<float*>(W + r0.x_red_off[L]) : nullptr;
            LAUNCH_SPLIT(KC_NORM_PREFIX, T.nprefix, NormPrefixTask, offs[L].nprefix, (uint64_t)0, launch_norm_total(tp, nb, red, s));
            LAUNCH_SPLIT(KC_BP_CARRY, T.bp, BpTask, offs[L].bp, (uint64_t)0, launch_bp_total(tp, nb, D, s));
            if (x_mode == 2 && r0.x_stride[L]) {
                NcclApi* na = nccl_api();
                float* xb = reinterpret_cast<float*>(W + r0.x_off[L]);
                // one exchange round of this level (timed as its own class: the time includes waiting for the slowest rank)
                if (g->profiling) CU(cudaEventRecord(ev_at(2 * timed.size()), s));
                NC(na->AllGather(xb + (size_t)r0.x_rank * r0.x_stride[L], xb, r0.x_stride[L], kNcclFloat32, r0.c->nccl_comm, s));
                if (red && r0.x_norms[L]) NC(na->AllReduce(red, red, r0.x_norms[L], kNcclFloat32, kNcclMax, r0.c->nccl_comm, s));
                if (g->profiling) { CU(cudaEventRecord(ev_at(2 * timed.size() + 1), s)); timed.push_back({KC_XCHG, (uint64_t)r0.x_ranks * r0.x_stride[L] * 4}); }
                ++collectives;
            }
        }
        LAUNCH_SPLIT(KC_NORM_PREFIX, T.nprefix, NormPrefixTask, offs[L].nprefix, T.nprefix.bytes, launch_norm_prefix(tp, nb, s));
        LAUNCH_SPLIT(KC_NORM_SCALE, T.nscale, NormScaleTask, offs[L].nscale, T.nscale.bytes, launch_norm_scale(tp, nb, T.nscale.max_chunks, vec, s));
        if (!bp_grouped && !bp_fused) {
        LAUNCH_SPLIT(KC_BP_CARRY, T.bp, BpTask, offs[L].bp, (uint64_t)0, launch_bp_carry(tp, nb, D, s));
        LAUNCH_SPLIT(KC_BP_APPLY, T.bp, BpTask, offs[L].bp, T.bp.bytes, launch_bp_apply(tp, nb, T.bp.max_chunks, vec, D, s));
        }
    }
    if (x_mode) {
        // time-sharded resampling: every segment publishes the last 512 frames of the output vertex; the next
        // segment's resampler reads them as its history (state.rs:533-550 keeps that history inside rubato)
        bool any = false;
        for (auto& r : reqs) {
            if (r.halo_off < 0 || !r.resample) continue;
            const float* ol = reinterpret_cast<const float*>(W + r.buf_off[r.c->output]);
            float* slot = reinterpret_cast<float*>(W + r.halo_off) + (size_t)r.x_rank * 2 * kResampleHalo;
            CU(cudaMemcpyAsync(slot, ol + r.n_frames - kResampleHalo, kResampleHalo * 4, cudaMemcpyDeviceToDevice, s));
            CU(cudaMemcpyAsync(slot + kResampleHalo, ol + r.n_pad + r.n_frames - kResampleHalo, kResampleHalo * 4, cudaMemcpyDeviceToDevice, s));
            any = true;
        }
        if (any && x_mode == 2) {
            const PassReq& r0 = reqs[0];
            NcclApi* na = nccl_api();
            float* hb = reinterpret_cast<float*>(W + r0.halo_off);
            if (g->profiling) CU(cudaEventRecord(ev_at(2 * timed.size()), s));
            NC(na->AllGather(hb + (size_t)r0.x_rank * 2 * kResampleHalo, hb, 2 * kResampleHalo, kNcclFloat32, r0.c->nccl_comm, s));
            if (g->profiling) { CU(cudaEventRecord(ev_at(2 * timed.size() + 1), s)); timed.push_back({KC_XCHG, (uint64_t)r0.x_ranks * 2 * kResampleHalo * 4}); }
            ++collectives;
        }
    }
    { TaskList<ResampleTask> rl; rl.v.swap(resamps);
      LAUNCH_SPLIT(KC_RESAMPLE, rl, ResampleTask, resamp_off, resamp_bytes, launch_resample(tp, nb, resamp_max, J.resamp_chunks, D, s)); }
    { TaskList<QuantTask> ql; ql.v.swap(quants);
      LAUNCH_SPLIT(KC_QUANT, ql, QuantTask, quant_off, quant_bytes, launch_quantise(tp, nb, quant_max, s)); }
    CU(cudaEventRecord(sl.ev1, s));

    // ---- 4. results back -------------------------------------------------------------------------
    // PCM and TooWide flags leave on the engine's copy stream once the kernels are done (ev1): the lane's
    // stream is free for its next pass immediately, so kernels of group g+1 run under the D2H of group g.
    J.t_d2h0 = now_ms();
    uint64_t& d2h_bytes = J.d2h_bytes;
    d2h_bytes = 0;
    bool any_out = false;
    for (auto& r : reqs) if (!r.is_scan && r.c->output >= 0) any_out = true;
    cudaStream_t cs = any_out ? g->copy_stream : s;
    if (any_out) CU(cudaStreamWaitEvent(cs, sl.ev1, 0));
    for (size_t i = 0; i < reqs.size(); ++i) {
        PassReq& r = reqs[i];
        sl.err_host[i] = 0;
        if (r.is_scan || r.c->output < 0) continue;
        if (r.out) {
            if (r.cap < r.pcm_bytes) return r.c->fail(TDG_ERR_CAPACITY, "output buffer too small");
            // asynchronous when the caller's buffer is pinned (cudaHostAlloc / cudaHostRegister)
            CU(cudaMemcpyAsync(r.out, J.P + r.pcm_off, r.pcm_bytes, cudaMemcpyDeviceToHost, cs));
            d2h_bytes += r.pcm_bytes;
        }
    }
    if (any_out) CU(cudaMemcpyAsync(sl.err_host, J.P + J.err_base, J.err_bytes, cudaMemcpyDeviceToHost, cs));
    CU(cudaEventRecord(sl.ev_done, cs));
    for (auto& r : reqs) if (r.keep_write) r.c->keep.valid = true;      // the scan's kernels are queued: the record stands
    J.submitted = true;
    J.t_submitted = now_ms();
    if (trace_on()) fprintf(stderr, "[tdg] submit lane %d: enqueue %.2f ms, %llu launches\n", (int)(J.lane - g->lanes), now_ms() - J.t_planned, (unsigned long long)launches);
    return TDG_OK;
}

// end-of-pass transport state (graph.rs:235-236 / state.rs:575): host bookkeeping that does not depend
// on device results, applied right after submission so that the next pass of the same contexts can
// be planned while this one still runs
static void pass_end_state(PassJob& J) {
    for (auto& r : J.reqs) {
        Ctx* x = r.c;
        x->vstate_host_newer = false;
        if (r.is_scan) {
            if (x->output >= 0) { x->fb_frame = 0; }
        } else if (x->output >= 0) x->fb_frame += r.total * x->bl;      // set_time_to_next_block per chunk, never rewound
        if (r.play) { x->t += r.total * x->bl; continue; }              // Graph::render advanced t; nothing is rewound
        if (x->output >= 0 || !r.is_scan) {
            x->t = 0;                                                   // Graph::set_time(0)
            for (HVertex& hv : x->verts) {                              // VertexExt::set_time, extensions.rs:196-204
                if (hv.kind == K_SAMPLE_LOOP) hv.loop_t = 0;
                else if (hv.kind == K_DEBUG_SINE) hv.dnotes.clear();
                else if (hv.kind == K_SYNTH) hv.notes.clear();
                else if (hv.kind == K_BANDPASS) hv.bp_first = true;
            }
        }
    }
}

// wait for the pass's device work (its own completion event: later passes may already be queued on
// the same stream), then statistics and error flags
static int pass_wait(PassJob& J) {
    std::vector<PassReq>& reqs = J.reqs;
    if (reqs.empty() || !J.submitted || J.waited) return TDG_OK;
    J.waited = true;
    Lane& ln = *J.lane;
    Slot& sl = ln.slots[J.slot];
    Ctx* c = reqs[0].c;
    Engine* g = c->eng;
    cudaSetDevice(g->device);
    PinnedBuf& B = sl.blob;
    std::vector<Timed>& timed = J.timed;
    const uint64_t launches = J.launches, collectives = J.collectives, d2h_bytes = J.d2h_bytes, algo_bytes_total = J.algo_bytes_total;
    const double t_start = J.t_start, t_planned = J.t_planned, t_h2d_issued = J.t_h2d_issued, t_d2h0 = J.t_d2h0;
    const double t_sync0 = now_ms();
    CU(cudaEventSynchronize(sl.ev_done));
    const double t_end = now_ms();
    float dev_ms = 0.0f;
    cudaEventElapsedTime(&dev_ms, sl.ev0, sl.ev1);
    if (trace_on() && J.in_batch) {
        cudaEventElapsedTime(&J.dev_t0, g->ev_batch0, sl.ev0); cudaEventElapsedTime(&J.dev_t1, g->ev_batch0, sl.ev1);
        cudaEventElapsedTime(&J.dev_tdone, g->ev_batch0, sl.ev_done);
    }
    if (trace_on()) fprintf(stderr, "[tdg] wait lane %d slot %d: waited %.2f ms, device %.2f ms\n", (int)(J.lane - g->lanes), J.slot, t_end - t_sync0, dev_ms);
    KStat* ks = J.ks;
    for (int k = 0; k < KC_COUNT; ++k) ks[k] = KStat();
    for (size_t i = 0; i < timed.size(); ++i) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, sl.evpool[2 * i], sl.evpool[2 * i + 1]);
        ks[timed[i].cls].launches++; ks[timed[i].cls].ms += ms; ks[timed[i].cls].bytes += timed[i].bytes;
    }
    int rc = TDG_OK;
    for (size_t i = 0; i < reqs.size(); ++i) {
        PassReq& r = reqs[i];
        Ctx* x = r.c;
        r.err_flag = sl.err_host[i];
        if (r.frames_written) *r.frames_written = (r.is_scan || x->output < 0) ? 0 : r.out_frames;
        x->stats.device_ms = dev_ms;
        x->stats.plan_ms = t_planned - t_start;
        x->stats.h2d_ms = t_h2d_issued - t_planned;
        x->stats.d2h_ms = t_end - t_d2h0;
        x->stats.h2d_bytes = B.size; x->stats.d2h_bytes = d2h_bytes;
        x->stats.kernel_launches = launches; x->collectives += collectives;
        x->launches += launches;
        for (int k = 0; k < KC_COUNT; ++k) x->kstats[k] = ks[k];
        if (reqs.size() > 1) x->stats.algorithmic_bytes = algo_bytes_total;
        x->stats.reused_vertices = J.reused_v; x->stats.reused_bytes = J.reused_b; x->stats.inlined_vertices = J.inlined_v; x->stats.inlined_bytes = J.inlined_b;
        if (r.err_flag) rc = x->fail(TDG_ERR_TOO_WIDE, "sample too wide for bit depth (hound TooWide; the reference panics)");
    }
    return rc;
}

struct PlayJob { PassJob J; Ctx* c = nullptr; uint64_t frames = 0; bool waited = false; };
static int drain_plays(Ctx* c) {
    if (!c || c->plays_in_flight == 0) return TDG_OK;
    Engine* g = c->eng;
    cudaSetDevice(g->device);
    for (int sl = 0; sl < kSlots; ++sl) {
        PlayJob* pj = g->play_slot[sl];
        if (!pj || pj->c != c || pj->waited) continue;
        if (g->play_ev[sl]) cudaEventSynchronize(g->play_ev[sl]);
    }
    return TDG_OK;
}
static int run_pass(std::vector<PassReq>& reqs) {
    if (reqs.empty()) return TDG_OK;
    for (auto& r : reqs) drain_plays(r.c);
    PassJob J;
    J.reqs.swap(reqs);
    J.lane = &J.reqs[0].c->eng->lanes[0];
    J.slot = 0;
    int e = pass_plan(J);
    if (e == TDG_OK) e = pass_submit(J);
    if (e == TDG_OK) { pass_end_state(J); e = pass_wait(J); }
    else { cudaStreamSynchronize(J.lane->stream); cudaStreamSynchronize(J.lane == nullptr ? nullptr : J.reqs[0].c->eng->copy_stream); }
    reqs.swap(J.reqs);
    return e;
}

// Batches of independent projects, software-pipelined.  `jobs` are passes in host order; job j runs on
// lane/slot (jobs[j].lane, jobs[j].slot).  Before a job is planned, the previous job that used the
// same slot must have completed (its pinned and device blobs are reused); everything else overlaps:
// the host plans job j+1 while the SMs run job j and the copy engine drains job j-1's PCM.
// Results are identical to unpipelined passes (projects are independent; passes of one project stay
// in order on one stream).
static int run_jobs_pipelined(Engine* g, std::vector<PassJob>& jobs, PassStats* tot_out, KStat* ks_out) {
    int rc = TDG_OK;
    PassJob* last_in_slot[kLanes][kSlots] = {};
    bool first = true, used[kLanes] = {};
    for (size_t k = 0; k < jobs.size() && rc == TDG_OK; ++k) {
        PassJob& J = jobs[k];
        const int li = (int)(J.lane - g->lanes);
        if (PassJob* prev = last_in_slot[li][J.slot]) { rc = pass_wait(*prev); if (rc != TDG_OK) break; }
        if (J.pre) { rc = J.pre(); if (rc != TDG_OK) break; }
        J.in_batch = true;
        rc = pass_plan(J);
        if (rc != TDG_OK) break;
        if (first) { cudaEventRecord(g->ev_batch0, J.lane->stream); first = false; }
        rc = pass_submit(J);
        if (rc != TDG_OK) break;
        pass_end_state(J);
        cudaEventRecord(g->ev_batch1[li], J.lane->stream);          // after this job's kernels and copies
        used[li] = true;
        last_in_slot[li][J.slot] = &J;
    }
    if (!first) cudaEventRecord(g->ev_copy_end, g->copy_stream);     // after the last PCM copy
    for (PassJob& J : jobs) { const int e = pass_wait(J); if (rc == TDG_OK) rc = e; }
    if (rc != TDG_OK) { for (Lane& ln : g->lanes) cudaStreamSynchronize(ln.stream); cudaStreamSynchronize(g->copy_stream); return rc; }
    if (trace_on()) {
        const double h0 = jobs[0].t_start;
        for (size_t k = 0; k < jobs.size(); ++k) {
            const PassJob& J = jobs[k];
            fprintf(stderr, "[tdg] timeline job %2zu lane %d slot %d %s: host plan %.1f..%.1f submit ..%.1f | device kernels %.1f..%.1f d2h ..%.1f\n",
                    k, (int)(J.lane - g->lanes), J.slot, J.reqs[0].is_scan ? "scan  " : "render", J.t_start - h0, J.t_planned - h0,
                    J.t_submitted - h0, J.dev_t0, J.dev_t1, J.dev_tdone);
        }
    }
    // whole-call figures: device time from the first job's start to the last job's end (all lanes)
    float span = 0.0f;
    for (int li = 0; li < kLanes; ++li) if (used[li]) { float v = 0.0f; cudaEventElapsedTime(&v, g->ev_batch0, g->ev_batch1[li]); span = std::max(span, v); }
    { cudaStreamSynchronize(g->copy_stream); float v = 0.0f; if (cudaEventElapsedTime(&v, g->ev_batch0, g->ev_copy_end) == cudaSuccess) span = std::max(span, v); else cudaGetLastError(); }
    PassStats tot;
    for (PassJob& J : jobs) {
        const Ctx* x = J.reqs[0].c;          // every context of a job carries the job's figures
        // a context may appear in two jobs (scan, render): its stats hold the later one; per-job numbers were
        // accumulated below from the job records instead
        (void)x;
        tot.plan_ms += J.t_planned - J.t_start; tot.h2d_ms += J.t_h2d_issued - J.t_planned;
        tot.h2d_bytes += J.h2d_bytes; tot.d2h_bytes += J.d2h_bytes; tot.kernel_launches += J.launches;
        tot.algorithmic_bytes += J.algo_bytes_total; tot.reused_vertices += J.reused_v; tot.reused_bytes += J.reused_b;
        tot.inlined_vertices += J.inlined_v; tot.inlined_bytes += J.inlined_b;
        for (PassReq& r : J.reqs) { tot.frames += r.n_frames; }
        for (int q = 0; q < KC_COUNT; ++q) { ks_out[q].launches += J.ks[q].launches; ks_out[q].ms += J.ks[q].ms; ks_out[q].bytes += J.ks[q].bytes; }
    }
    tot.device_ms = span;
    *tot_out = tot;
    return TDG_OK;
}
static void publish_batch_stats(std::vector<PassJob>& jobs, const PassStats& tot, const KStat* ks) {
    for (PassJob& J : jobs) for (PassReq& r : J.reqs) {
        const uint64_t fr = r.c->stats.frames, vf = r.c->stats.voice_frames;
        r.c->stats = tot; r.c->stats.frames = fr; r.c->stats.voice_frames = vf;
        r.c->batch_frames = tot.frames;
        for (int q = 0; q < KC_COUNT; ++q) r.c->kstats[q] = ks[q];
    }
}

static int run_pass_pipelined(std::vector<PassReq>& reqs) {
    if (reqs.empty()) return TDG_OK;
    for (auto& r : reqs) drain_plays(r.c);
    Engine* g = reqs[0].c->eng;
    const size_t group = g->pipeline_group;
    if (group == 0 || reqs.size() <= group || reqs[0].x_mode != 0) return run_pass(reqs);
    for (auto& r : reqs) if (r.c->eng != g) return r.c->fail(TDG_ERR_INVALID, "all contexts of a batch must be on one device");
    cudaSetDevice(g->device);
    const size_t n_groups = (reqs.size() + group - 1) / group;
    std::vector<PassJob> jobs(n_groups);
    for (size_t i = 0; i < reqs.size(); ++i) jobs[i / group].reqs.push_back(std::move(reqs[i]));
    for (size_t k = 0; k < n_groups; ++k) { jobs[k].lane = &g->lanes[k % g->n_lanes]; jobs[k].slot = (int)((k / g->n_lanes) % kSlots); }
    PassStats tot; KStat ks[KC_COUNT];
    const int rc = run_jobs_pipelined(g, jobs, &tot, ks);
    if (rc == TDG_OK) publish_batch_stats(jobs, tot, ks);
    size_t i = 0;
    for (PassJob& J : jobs) for (PassReq& r : J.reqs) reqs[i++] = std::move(r);
    return rc;
}

// scan + render of every project in one pipelined call (the manual workflow's "normalize" then
// "render", ui_workflow.rs:120-135): group g's two passes run back to back on one lane while the other
// lane works on group g+1, so the PCM of group g-1 crosses PCIe under the kernels of both passes.
static int run_scan_render_pipelined(std::vector<PassReq>& scans, std::vector<PassReq>& renders,
                                     const std::function<int(size_t, size_t)>& refresh_range = nullptr) {
    Engine* g = renders[0].c->eng;
    for (auto& r : renders) drain_plays(r.c);
    for (auto& r : renders) if (r.c->eng != g) return r.c->fail(TDG_ERR_INVALID, "all contexts of a batch must be on one device");
    cudaSetDevice(g->device);
    const size_t group = g->pipeline_group ? g->pipeline_group : renders.size();
    // group boundaries: a geometric ramp (8, 16, ... < G projects), full groups, then a geometric taper (G/2,
    // G/4, ... >= 8 projects).  The PCM copies form one queue on the copy stream and are the longest chain of
    // the call once the kernels are fast enough: the ramp lets the first copy start after a few hundred
    // microseconds of kernels instead of a full group's scan + render (measured: first D2H at 15.8 ms of a
    // 77 ms call with equal groups), and what is left after the last kernel is only the last small group's PCM
    std::vector<size_t> start;                 // first project of every group, plus the end
    {
        const size_t n = renders.size(), floor_sz = std::min<size_t>(8, group);
        double tail_q = 0.8;                             // measured on B200 + PCIe 5: 0.5 -> 80.5, 0.65 -> 79.2, 0.8 -> 78.2 ms per 512-project call
        if (const char* e = getenv("TDG_TAIL_RATIO")) tail_q = std::min(0.9, std::max(0.1, atof(e)));
        size_t at = 0;
        if (g->taper && floor_sz > 0)
            for (size_t sz = floor_sz; sz < group && at + sz + 2 * group <= n; sz *= 2) { start.push_back(at); at += sz; }
        while (at < n) {
            start.push_back(at);
            const size_t rem = n - at;
            size_t sz = group;
            // geometric tail: group sizes G*q, G*q^2, ... (a tail of total length R starts with R*(1-q)).  Every
            // burst of PCM has to drain over PCIe while the next, smaller group renders; q is the measured
            // kernel/copy balance of the box (TDG_TAIL_RATIO, default 0.8)
            if (g->taper && (double)rem * (1.0 - tail_q) <= (double)group)
                sz = std::max(floor_sz, std::min(group, (size_t)ceil((double)rem * (1.0 - tail_q))));
            if (g->taper && rem > sz && rem - sz < floor_sz) sz = rem;          // no crumbs: a 1-project group costs two full launch sequences
            at += std::min(sz, rem);
        }
        start.push_back(n);
    }
    const size_t n_groups = start.size() - 1;
    std::vector<PassJob> jobs(n_groups * 2);
    for (size_t gi = 0; gi < n_groups; ++gi) for (size_t i = start[gi]; i < start[gi + 1]; ++i) {
        if (scans[i].c->output >= 0) jobs[2 * gi].reqs.push_back(std::move(scans[i]));     // graph.rs:225: no output, no scan
        jobs[2 * gi + 1].reqs.push_back(std::move(renders[i]));
    }
    // the refresh of a group's projects (floww reload, vertex state, transport, normalize maxima) runs on the host right
    // before the group's first pass is planned, i.e. while the GPU works on the groups in front of it
    if (refresh_range)
        for (size_t gi = 0; gi < n_groups; ++gi) {
            const size_t i0 = start[gi], i1 = start[gi + 1];
            PassJob& first_job = jobs[2 * gi].reqs.empty() ? jobs[2 * gi + 1] : jobs[2 * gi];
            first_job.pre = [&refresh_range, i0, i1]() { return refresh_range(i0, i1); };
        }
    std::vector<PassJob> live;
    live.reserve(jobs.size());
    int next_slot[kLanes] = {};
    for (size_t k = 0; k < jobs.size(); ++k) {
        if (jobs[k].reqs.empty()) continue;
        const int li = (int)((k >> 1) % g->n_lanes);
        jobs[k].lane = &g->lanes[li]; jobs[k].slot = next_slot[li]++ % kSlots;       // ring: a slot is reused kSlots passes later
        live.push_back(std::move(jobs[k]));
    }
    PassStats tot; KStat ks[KC_COUNT];
    const int rc = run_jobs_pipelined(g, live, &tot, ks);
    if (rc == TDG_OK) publish_batch_stats(live, tot, ks);
    return rc;
}

}  // namespace tdg

// ====================================================================================== C ABI
using namespace tdg;
struct tdg_ctx { Ctx c; };
#define C_ (&h->c)

extern "C" {

int tdg_sample_mode_from_string(const char* s) {
    std::string m = s ? s : "";
    if (m == "left") return TDG_MODE_LEFT;
    if (m == "right") return TDG_MODE_RIGHT;
    if (m == "loudest") return TDG_MODE_LOUDEST;
    if (m == "normalize-seperate") return TDG_MODE_NORM;
    if (m == "mix-down") return TDG_MODE_MIX;
    return TDG_MODE_STEREO;
}

int tdg_ctx_create(uint32_t project_sr, uint32_t buffer_len, int device, tdg_ctx** out) {
    if (!out || project_sr == 0 || buffer_len == 0) return TDG_ERR_INVALID;
    *out = nullptr;
    std::string err;
    Engine* g = engine_get(device, &err);
    if (!g) { fprintf(stderr, "termdaw_b200: %s\n", err.c_str()); return TDG_ERR_CUDA; }
    tdg_ctx* h = new tdg_ctx();
    h->c.eng = g; h->c.sr = project_sr; h->c.bl = buffer_len;
    *out = h;
    return TDG_OK;
}
void tdg_ctx_destroy(tdg_ctx* h) {
    if (!h) return;
    Ctx* c = C_;
    cudaSetDevice(c->eng->device);
    cudaStreamSynchronize(c->eng->stream);
    {                                      // playback tickets of this context die with it
        Engine* g = c->eng;
        cudaStreamSynchronize(g->lanes[kLanes - 1].stream);
        for (int sl = 0; sl < kSlots; ++sl) if (g->play_slot[sl] && g->play_slot[sl]->c == c) { delete g->play_slot[sl]; g->play_slot[sl] = nullptr; }
    }
    for (auto& s : c->samples) { if (s.l) cudaFree(s.l); }
    for (auto& t : c->tables) if (t.data) cudaFree(t.data);
    if (c->vstate_dev) cudaFree(c->vstate_dev);
    if (c->nccl_comm) { nccl_api()->CommDestroy(c->nccl_comm); c->nccl_comm = nullptr; }
    { std::lock_guard<std::mutex> lk(g_mu); c->eng->refs--; }
    delete h;
}
const char* tdg_last_error(tdg_ctx* h) { return h ? h->c.err.c_str() : "null context"; }
uint64_t tdg_kernel_launches(tdg_ctx* h) { return h ? h->c.launches : 0; }

// SampleBank::add numeric part (sample.rs:262-310) on the device
int tdg_sample_add(tdg_ctx* h, const float* linear, uint64_t n_values, uint32_t channels, uint32_t src_sr, int mode) {
    Ctx* c = C_;
    Engine* g = c->eng;
    cudaSetDevice(g->device);
    if (mode == TDG_MODE_STEREO && channels != 2)
        return c->fail(TDG_ERR_SAMPLE, "only 2 channel samples are supported for stereo samples");
    if (mode != TDG_MODE_STEREO && (channels > 2 || channels < 1))
        return c->fail(TDG_ERR_SAMPLE, "only 1,2 channel samples are supported for left or right samples");
    // channel lengths after de-interleaving (sample.rs:275-292)
    uint64_t nl, nr;
    if (channels == 1) { nl = mode == TDG_MODE_LEFT ? n_values : 0; nr = mode == TDG_MODE_LEFT ? 0 : n_values; }
    else { nl = n_values / 2 + (n_values & 1); nr = n_values / 2; }
    uint32_t stride = channels, off_l = 0, off_r = channels == 2 ? 1 : 0, mix = 0;
    uint64_t frames = 0;
    switch (mode) {                                                        // Sample::from, sample.rs:38-77
        case TDG_MODE_LEFT: if (nl == 0) return c->fail(TDG_ERR_SAMPLE, "Sample::from: l has length 0");
            frames = nl; off_r = off_l = 0; break;
        case TDG_MODE_RIGHT: if (nr == 0) return c->fail(TDG_ERR_SAMPLE, "Sample::from: r has length 0");
            frames = nr; off_l = off_r = channels == 2 ? 1 : 0; break;
        case TDG_MODE_LOUDEST: {
            // mean_energy is a sequential f32 sum (sample.rs:16-22): a decision, taken on the host
            float sl = 0.0f, sr = 0.0f;
            if (channels == 2) { for (uint64_t i = 0; i < nl; ++i) sl += fabsf(linear[i * 2]);
                                 for (uint64_t i = 0; i < nr; ++i) sr += fabsf(linear[i * 2 + 1]); }
            else { for (uint64_t i = 0; i < nr; ++i) sr += fabsf(linear[i]); }
            const float lm = nl ? sl / (float)nl : 0.0f, rm = nr ? sr / (float)nr : 0.0f;
            if (lm > rm) { frames = nl; off_l = off_r = 0; } else { frames = nr; off_l = off_r = channels == 2 ? 1 : 0; }
            break;
        }
        default:
            if (nl != nr) return c->fail(TDG_ERR_SAMPLE, "Sample::from: l and r do not have the same length");
            if (nl == 0) return c->fail(TDG_ERR_SAMPLE, "Sample::from: l and r have length 0");
            frames = nl; mix = mode == TDG_MODE_MIX ? 1 : 0;
    }
    if (frames == 0) return c->fail(TDG_ERR_SAMPLE, "empty sample");
    cudaStream_t s = g->stream;
    float* lin_d = nullptr; float* buf = nullptr; int* mx = nullptr;
    float* ob = nullptr; uint8_t* mini = nullptr; double* idxb = nullptr;
    struct Scratch {                      // every early return (CU) frees what has been allocated so far
        cudaStream_t s; void** p[6]; int n = 0;
        ~Scratch() { cudaStreamSynchronize(s); for (int i = 0; i < n; ++i) if (*p[i]) { cudaFree(*p[i]); *p[i] = nullptr; } }
    } scratch{s, {(void**)&lin_d, (void**)&buf, (void**)&mx, (void**)&ob, (void**)&mini, (void**)&idxb}, 6};
    const uint64_t fpad = align_up(frames, 4);
    CU(cudaMalloc(&lin_d, std::max<uint64_t>(n_values, 1) * 4));
    CU(cudaMalloc(&buf, fpad * 2 * 4));
    CU(cudaMalloc(&mx, 8));
    CU(cudaMemcpyAsync(lin_d, linear, n_values * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(mx, 0, 8, s));
    CU(cudaMemsetAsync(buf, 0, fpad * 2 * 4, s));
    float* dl = buf; float* dr = buf + fpad;
    CU(launch_deinterleave(lin_d, frames, stride, off_l, off_r, mix, dl, dr, s));
    CU(launch_absmax(dl, frames, mx, s));
    CU(launch_absmax(dr, frames, mx + 1, s));
    const float* ma = reinterpret_cast<const float*>(mx);
    if (mode == TDG_MODE_NORM) {                                           // normalize_seperate, sample.rs:132-137
        CU(launch_scale_by_inv(dl, frames, ma, nullptr, s));
        CU(launch_scale_by_inv(dr, frames, ma + 1, nullptr, s));
    } else {                                                               // normalize / mix_down (l == r there)
        CU(launch_scale_by_inv(dl, frames, ma, ma + 1, s));
        CU(launch_scale_by_inv(dr, frames, ma, ma + 1, s));
    }
    c->launches += 5;
    DevSample out;
    if (src_sr != c->sr) {                                                 // Sample::resample, sample.rs:150-175
        const double ratio = (double)c->sr / (double)src_sr;
        std::vector<ResampleChunk> ch; uint64_t total = 0;
        resample_replay(ratio, frames, 1, ch, &total);
        if (total == 0) return c->fail(TDG_ERR_SAMPLE, "sample too short to resample");
        const uint64_t opad = align_up(total, 4);
        CU(cudaMalloc(&ob, opad * 2 * 4));
        CU(cudaMalloc(&idxb, total * 8));
        CU(cudaMemsetAsync(ob, 0, opad * 2 * 4, s));
        CU(cudaMalloc(&mini, 256));
        ResampleTask rt; memset(&rt, 0, sizeof rt);
        rt.in_l = dl; rt.in_r = dr; rt.out_l = ob; rt.out_r = ob + opad;
        rt.sincs = sinc_table_dev(c, ratio);
        if (!rt.sincs) return c->fail(TDG_ERR_CUDA, "could not build the sinc table");
        rt.n_in = frames; rt.chunk_off = 128; rt.n_chunks = 1; rt.chunk = (uint32_t)frames; rt.t_ratio = 1.0 / ratio; rt.idx = idxb;
        uint8_t host[256]; memset(host, 0, sizeof host);
        memcpy(host, &rt, sizeof rt); memcpy(host + 128, &ch[0], sizeof(ResampleChunk));
        static_assert(sizeof(ResampleTask) <= 128, "mini blob layout");
        CU(cudaMemcpyAsync(mini, host, 256, cudaMemcpyHostToDevice, s));
        CU(launch_resample(reinterpret_cast<const ResampleTask*>(mini), 1, total, 1, mini, s));
        c->launches += 2;
        CU(cudaStreamSynchronize(s));
        out.l = ob; out.r = ob + opad; out.frames = total;
        ob = nullptr;                                                      // kept: the sample bank owns it now
    } else {
        CU(cudaStreamSynchronize(s));
        out.l = dl; out.r = dr; out.frames = frames;
        buf = nullptr;
    }
    c->samples.push_back(out);
    return (int)c->samples.size() - 1;
}
// SampleBank::mark_dead + refresh (sample.rs:316-336): a sample that left the script (or changed) is dropped from the
// bank.  Indices of the other samples stay valid (the reference re-indexes; here the slot is simply emptied), the
// device memory is returned.  Vertices are rebuilt by the refresh that follows; a vertex still naming the slot
// makes the next pass fail instead of reading freed memory.
int tdg_sample_remove(tdg_ctx* h, int sample) {
    Ctx* c = C_;
    if (sample < 0 || (size_t)sample >= c->samples.size()) return c->fail(TDG_ERR_INVALID, "unknown sample");
    DevSample& s = c->samples[(size_t)sample];
    if (s.l) { cudaSetDevice(c->eng->device); cudaFree(s.l); }
    s = DevSample();
    ++c->mut_gen; c->keep.valid = false;
    return TDG_OK;
}
// BufferBank's dead-resource sweep (state.rs:222-237): the blob of a removed / changed resource is released
int tdg_resource_remove(tdg_ctx* h, int resource) {
    Ctx* c = C_;
    if (resource < 0 || (size_t)resource >= c->resources.size()) return c->fail(TDG_ERR_INVALID, "unknown resource");
    std::vector<uint8_t>().swap(c->resources[(size_t)resource]);
    return TDG_OK;
}
int64_t tdg_sample_frames(tdg_ctx* h, int sample) {
    if (sample < 0 || (size_t)sample >= h->c.samples.size()) return TDG_ERR_INVALID;
    return (int64_t)h->c.samples[sample].frames;
}
int tdg_sample_read(tdg_ctx* h, int sample, float* l, float* r, uint64_t cap) {
    Ctx* c = C_;
    if (sample < 0 || (size_t)sample >= c->samples.size()) return c->fail(TDG_ERR_INVALID, "unknown sample");
    const DevSample& s = c->samples[sample];
    if (cap < s.frames) return c->fail(TDG_ERR_CAPACITY, "sample_read: capacity too small");
    cudaSetDevice(c->eng->device);
    CU(cudaMemcpy(l, s.l, s.frames * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(r, s.r, s.frames * 4, cudaMemcpyDeviceToHost));
    return TDG_OK;
}
int tdg_resource_add(tdg_ctx* h, const uint8_t* bytes, uint64_t n) {
    h->c.resources.emplace_back(bytes, bytes + n);
    return (int)h->c.resources.size() - 1;
}
int tdg_floww_add(tdg_ctx* h, const float* t, const float* note, const float* vel, uint64_t n) {
    Ctx* c = C_;
    Floww f;
    f.t.assign(t, t + n); f.note.assign(note, note + n); f.vel.assign(vel, vel + n);
    f.fe.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        f.fe[i] = f32_as_usize(t[i] * (float)c->sr);                       // floww.rs:75
        if (i && f.fe[i] < f.fe[i - 1]) return c->fail(TDG_ERR_UNSORTED, "floww events must be sorted by time");
    }
    c->flowws.push_back(std::move(f));
    ++c->mut_gen;
    return (int)c->flowws.size() - 1;
}
// FlowwBank::append_streams (floww.rs:55-57, stream_workflow.rs:63-71): more events for a declared stream while
// the transport runs.  The device planner walks sorted events, so the appended events must not start before
// the stream's last event (TDG_ERR_UNSORTED otherwise, as for tdg_floww_add).  The caller then re-seats the
// floww cursor like the reference's Feed handler does: tdg_floww_set_time(ctx, tdg_get_time(ctx)).
int tdg_floww_append(tdg_ctx* h, int floww, const float* t, const float* note, const float* vel, uint64_t n) {
    Ctx* c = C_;
    if (floww < 0 || (size_t)floww >= c->flowws.size()) return c->fail(TDG_ERR_INVALID, "unknown floww");
    Floww& f = c->flowws[(size_t)floww];
    uint64_t last = f.fe.empty() ? 0 : f.fe.back();
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t fe = f32_as_usize(t[i] * (float)c->sr);                  // floww.rs:75
        if (fe < last) return c->fail(TDG_ERR_UNSORTED, "appended floww events must be sorted by time");
        last = fe;
    }
    for (uint64_t i = 0; i < n; ++i) {
        f.t.push_back(t[i]); f.note.push_back(note[i]); f.vel.push_back(vel[i]);
        f.fe.push_back(f32_as_usize(t[i] * (float)c->sr));
    }
    ++c->mut_gen;
    return TDG_OK;
}
int tdg_floww_reset(tdg_ctx* h) { h->c.flowws.clear(); h->c.fb_frame = 0; ++h->c.mut_gen; return TDG_OK; }     // floww.rs:21-28

int tdg_graph_reset(tdg_ctx* h) {                                           // graph.rs:39-47
    Ctx* c = C_;
    // the graph is rebuilt from scratch on every refresh (state.rs:328): the wavetables its SampSyn vertices
    // uploaded go with it (no pass of this context is in flight: every entry point returns after its pass_wait)
    if (!c->tables.empty()) {
        cudaSetDevice(c->eng->device);
        for (auto& t : c->tables) if (t.data) cudaFree(t.data);
        c->tables.clear();
    }
    c->verts.clear(); c->output = -1; c->t = 0;
    ++c->mut_gen; c->keep = KeepRecord();
    c->vstate_host.clear(); c->vstate_host_newer = true;
    c->last_buf_off.clear();
    return TDG_OK;
}

static void vstate_init(float* st, int kind) {
    for (int i = 0; i < 8; ++i) st[i] = 0.0f;
    if (kind != K_NORMALIZE) { st[0] = -1.0f; st[1] = -1.0f; }
}
int tdg_reset_vertex_state(tdg_ctx* h) {                                    // state.rs:328-457 on an unchanged project
    Ctx* c = C_;
    c->vstate_host.assign(c->verts.size() * 8, 0.0f);
    for (size_t i = 0; i < c->verts.size(); ++i) {
        HVertex& v = c->verts[i];
        v.loop_t = 0; v.ts.clear(); v.countdown = 0; v.primary = {0, 0.0f}; v.ghost = {0, 0.0f};
        v.dnotes.clear(); v.notes.clear(); v.seq = 0;
        v.aprimary = {0, 0, 0}; v.aghost = {0, 0, 0};
        v.bp_first = true;
        vstate_init(&c->vstate_host[i * 8], v.kind);
    }
    c->vstate_host_newer = true;         // the device copy is stale; no read-back needed
    c->t = 0;
    ++c->mut_gen;
    return TDG_OK;
}

static int add_vertex(Ctx* c, HVertex v, float gain, float angle, float wet) {
    v.gain = gain;
    v.angle = fmaxf(fminf(angle, 90.0f), -90.0f);                           // graph.rs:255
    v.wet = fmaxf(fminf(wet, 1.0f), 0.0f);                                  // graph.rs:256
    if (vstate_to_host(c) != TDG_OK) return TDG_ERR_CUDA;
    ++c->mut_gen;
    c->verts.push_back(std::move(v));
    c->vstate_host.resize(c->verts.size() * 8, 0.0f);
    vstate_init(&c->vstate_host[(c->verts.size() - 1) * 8], c->verts.back().kind);
    c->vstate_host_newer = true;
    return (int)c->verts.size() - 1;
}
static bool parse_adsr(const float* a, int n, AdsrConfD* o) {               // adsr.rs:94-114
    memset(o, 0, sizeof *o);
    if (n == 0) return true;
    if (n == 6) { o->std_vel = 0.0f; o->attack_sec = a[0]; o->attack_vel = 1.0f; o->decay_sec = a[1]; o->decay_vel = a[2];
                  o->sustain_sec = a[3]; o->sustain_vel = a[4]; o->release_sec = a[5]; o->release_vel = 0.0f; return true; }
    if (n == 9) { o->std_vel = a[0]; o->attack_sec = a[1]; o->attack_vel = a[2]; o->decay_sec = a[3]; o->decay_vel = a[4];
                  o->sustain_sec = a[5]; o->sustain_vel = a[6]; o->release_sec = a[7]; o->release_vel = a[8]; return true; }
    return false;
}
#define NEED_SAMPLE(s) if ((s) < 0 || (size_t)(s) >= c->samples.size() || !c->samples[(size_t)(s)].l) return c->fail(TDG_ERR_INVALID, "unknown sample index")
#define NEED_FLOWW(f) if ((f) < 0 || (size_t)(f) >= c->flowws.size()) return c->fail(TDG_ERR_INVALID, "unknown floww index")

int tdg_add_sum(tdg_ctx* h, float gain, float angle) { HVertex v; v.kind = K_SUM; return add_vertex(C_, v, gain, angle, 0.0f); }
int tdg_add_normalize(tdg_ctx* h, float gain, float angle) { HVertex v; v.kind = K_NORMALIZE; return add_vertex(C_, v, gain, angle, 0.0f); }
int tdg_add_sample_loop(tdg_ctx* h, float gain, float angle, int sample) {
    Ctx* c = C_; NEED_SAMPLE(sample);
    HVertex v; v.kind = K_SAMPLE_LOOP; v.sample = sample; return add_vertex(c, v, gain, angle, 0.0f);
}
int tdg_add_sample_multi(tdg_ctx* h, float gain, float angle, int sample, int floww, int note) {
    Ctx* c = C_; NEED_SAMPLE(sample); NEED_FLOWW(floww);
    HVertex v; v.kind = K_SAMPLE_MULTI; v.sample = sample; v.floww = floww; v.has_note = note >= 0; v.note = note >= 0 ? note : 0;
    return add_vertex(c, v, gain, angle, 0.0f);
}
int tdg_add_sample_lerp(tdg_ctx* h, float gain, float angle, int sample, int floww, int note, int lerp_len) {
    Ctx* c = C_; NEED_SAMPLE(sample); NEED_FLOWW(floww);
    HVertex v; v.kind = K_SAMPLE_LERP; v.sample = sample; v.floww = floww; v.has_note = note >= 0; v.note = note >= 0 ? note : 0;
    v.lerp_len = (uint32_t)std::max(lerp_len, 0);                           // state.rs:370
    return add_vertex(c, v, gain, angle, 0.0f);
}
int tdg_add_debug_sine(tdg_ctx* h, float gain, float angle, int floww) {
    Ctx* c = C_; NEED_FLOWW(floww);
    HVertex v; v.kind = K_DEBUG_SINE; v.floww = floww; return add_vertex(c, v, gain, angle, 0.0f);
}
int tdg_add_synth(tdg_ctx* h, float gain, float angle, int floww, float sq_vol, float sq_z, const float* sq_adsr, int sq_n,
                  float tf_vol, float tf_z, const float* tf_adsr, int tf_n, float tr_vol, const float* tr_adsr, int tr_n) {
    Ctx* c = C_; NEED_FLOWW(floww);
    HVertex v; v.kind = K_SYNTH; v.floww = floww;
    if (!parse_adsr(sq_adsr, sq_n, &v.sq.adsr) || !parse_adsr(tf_adsr, tf_n, &v.tf.adsr) || !parse_adsr(tr_adsr, tr_n, &v.tr.adsr))
        return c->fail(TDG_ERR_ADSR_CONF, "ADSR config must have 6 or 9 elements");
    v.sq.volume = sq_vol; v.sq.param = fmaxf(sq_z, 0.0001f);                // state.rs:400
    v.tf.volume = tf_vol; v.tf.param = tf_z;
    v.tr.volume = tr_vol; v.tr.param = 0.0f;
    return add_vertex(c, v, gain, angle, 0.0f);
}
int tdg_add_sampsyn(tdg_ctx* h, float gain, float angle, int floww, const float* adsr, int n, int resource) {
    Ctx* c = C_; NEED_FLOWW(floww);
    if (resource < 0 || (size_t)resource >= c->resources.size()) return c->fail(TDG_ERR_INVALID, "unknown resource index");
    HVertex v; v.kind = K_SAMPSYN; v.floww = floww;
    if (!parse_adsr(adsr, n, &v.adsr)) return c->fail(TDG_ERR_ADSR_CONF, "ADSR config must have 6 or 9 elements");
    // parse_wavetable_from_buffer (state.rs:415): "WTB1", u32 n_frames, u32 n_phase, f32 fps, f32 data[][]
    const std::vector<uint8_t>& b = c->resources[resource];
    DevTable tb; std::vector<float> data;
    bool ok = b.size() >= 16 && memcmp(b.data(), "WTB1", 4) == 0;
    if (ok) {
        uint32_t nf, np; float fps;
        memcpy(&nf, b.data() + 4, 4); memcpy(&np, b.data() + 8, 4); memcpy(&fps, b.data() + 12, 4);
        ok = nf >= 1 && np >= 2 && (uint64_t)nf * np * 4 + 16 == b.size();
        if (ok) { tb.n_frames = nf; tb.n_phase = np; tb.fps = fps; data.resize((size_t)nf * np); memcpy(data.data(), b.data() + 16, data.size() * 4); }
    }
    if (!ok) {                                                              // WaveTable::default() stand-in (state.rs:421)
        tb.n_frames = 1; tb.n_phase = 1024; tb.fps = 1.0f; data.resize(1024);
        for (uint32_t p = 0; p < 1024; ++p) data[p] = sinf(6.2831855f * ((float)p / 1024.0f));
    }
    cudaSetDevice(c->eng->device);
    // device layout (kernels.cu wavetable_eval): pairs[j][i] = (tab[j][i], tab[min(j + 1, last)][i]), i in [0, n_phase]
    // with column n_phase repeating column 0: the four corners of a bilinear cell are two adjacent 8-byte loads
    {
        const size_t nf = tb.n_frames, np = tb.n_phase, stride = np + 1;
        std::vector<float> pairs(nf * stride * 2);
        for (size_t j = 0; j < nf; ++j) {
            const size_t jn = std::min(j + 1, nf - 1);
            for (size_t i = 0; i <= np; ++i) {
                const size_t ii = i == np ? 0 : i;
                pairs[(j * stride + i) * 2] = data[j * np + ii];
                pairs[(j * stride + i) * 2 + 1] = data[jn * np + ii];
            }
        }
        CU(cudaMalloc(&tb.data, pairs.size() * 4));
        CU(cudaMemcpy(tb.data, pairs.data(), pairs.size() * 4, cudaMemcpyHostToDevice));
    }
    c->tables.push_back(tb);
    v.table = (int)c->tables.size() - 1;
    return add_vertex(c, v, gain, angle, 0.0f);
}
int tdg_add_lv2fx(tdg_ctx* h, float, float, float, int) {
    return h->c.fail(TDG_ERR_UNSUPPORTED_VERTEX, "lv2fx vertices run arbitrary native plugin code and are rejected by the GPU renderer");
}
int tdg_add_adsr(tdg_ctx* h, float gain, float angle, float wet, int floww, int use_off, int use_max, int note, const float* conf, int n) {
    Ctx* c = C_; NEED_FLOWW(floww);
    HVertex v; v.kind = K_ADSR; v.floww = floww; v.use_off = use_off != 0; v.use_max = use_max != 0;
    v.has_note = note >= 0; v.note = note >= 0 ? note : 0;
    if (!parse_adsr(conf, n, &v.adsr)) return c->fail(TDG_ERR_ADSR_CONF, "ADSR config must have 6 or 9 elements");
    return add_vertex(c, v, gain, angle, wet);
}
int tdg_add_bandpass(tdg_ctx* h, float gain, float angle, float wet, float lo, float hi, int pass) {
    Ctx* c = C_;
    HVertex v; v.kind = K_BANDPASS;                                         // extensions.rs:173-194
    const float lco = fmaxf(fminf(lo, 20000.0f), 0.0f), hco = fmaxf(fminf(hi, 20000.0f), 0.0f);
    v.lgamma = 1.0f - powf(kE, -2.0f * TDG_PI_F * lco / (float)c->sr);
    v.hgamma = 1.0f - powf(kE, -2.0f * TDG_PI_F * hco / (float)c->sr);
    v.pass = pass != 0; v.bp_first = true;
    return add_vertex(c, v, gain, angle, wet);
}
static bool has_loop(const Ctx* c, int x, int b) {                          // graph.rs:65-71
    if (x == b) return true;
    for (int y : c->verts[x].edges) if (has_loop(c, y, b)) return true;
    return false;
}
int tdg_connect(tdg_ctx* h, int a, int b) {                                 // graph.rs:58-78
    Ctx* c = C_;
    const int n = (int)c->verts.size();
    if (a < 0 || b < 0 || a == b || a >= n || b >= n) return 0;
    if (!c->verts[b].has_input()) return 0;
    if (has_loop(c, a, b)) return 0;
    c->verts[b].edges.push_back(a);
    ++c->mut_gen;
    return 1;
}
int tdg_set_output(tdg_ctx* h, int v) {
    if (v < 0 || (size_t)v >= h->c.verts.size()) return h->c.fail(TDG_ERR_INVALID, "unknown vertex");
    h->c.output = v; ++h->c.mut_gen; return TDG_OK;
}
int tdg_check_graph(tdg_ctx* h, int* unreachable, int cap, int* n_unreachable) {   // graph.rs:150-174
    Ctx* c = C_;
    if (n_unreachable) *n_unreachable = 0;
    if (c->output < 0) { c->err = "output vertex not found"; return 0; }
    if (c->verts[c->output].edges.empty() && c->verts[c->output].has_input()) { c->err = "output receives no inputs"; return 0; }
    std::vector<char> set(c->verts.size(), 0);
    std::vector<int> st{c->output};
    while (!st.empty()) { int v = st.back(); st.pop_back(); if (set[v]) continue; set[v] = 1; for (int y : c->verts[v].edges) st.push_back(y); }
    int n = 0;
    for (size_t i = 0; i < set.size(); ++i) if (!set[i]) { if (unreachable && n < cap) unreachable[n] = (int)i; ++n; }
    if (n_unreachable) *n_unreachable = n;
    return 1;
}
int tdg_reset_normalize(tdg_ctx* h) {                                       // extensions.rs:295-299
    Ctx* c = C_;
    if (int r = vstate_to_host(c)) return r;
    c->vstate_host.resize(c->verts.size() * 8, 0.0f);
    for (size_t v = 0; v < c->verts.size(); ++v) if (c->verts[v].kind == K_NORMALIZE) c->vstate_host[v * 8] = 0.000001f;
    c->vstate_host_newer = true;
    return TDG_OK;
}
float tdg_get_normalize_value(tdg_ctx* h, int v) {                          // extensions.rs:301-307
    Ctx* c = C_;
    if (v < 0 || (size_t)v >= c->verts.size() || c->verts[v].kind != K_NORMALIZE) return -1.0f;
    if (vstate_to_host(c) != TDG_OK) return -1.0f;
    return c->vstate_host[(size_t)v * 8];
}
int tdg_get_normalize_state(tdg_ctx* h, float* mx, float* smx, int cap) {
    Ctx* c = C_;
    if (int r = vstate_to_host(c)) return r;
    const int n = (int)c->verts.size();
    for (int i = 0; i < n && i < cap; ++i) { mx[i] = c->vstate_host[(size_t)i * 8]; smx[i] = c->vstate_host[(size_t)i * 8 + 1]; }
    return n;
}
int tdg_set_normalize_state(tdg_ctx* h, const float* mx, const float* smx, int n) {
    Ctx* c = C_;
    if (int r = vstate_to_host(c)) return r;
    for (int i = 0; i < n && (size_t)i < c->verts.size(); ++i)
        if (c->verts[i].kind == K_NORMALIZE) { c->vstate_host[(size_t)i * 8] = mx[i]; c->vstate_host[(size_t)i * 8 + 1] = smx[i]; }
    c->vstate_host_newer = true;
    return TDG_OK;
}
void* tdg_normalize_state_device_ptr(tdg_ctx* h, int* n_floats) {
    Ctx* c = C_;
    cudaSetDevice(c->eng->device);
    if (vstate_to_dev(c) != TDG_OK) return nullptr;
    cudaStreamSynchronize(c->eng->stream);
    if (n_floats) *n_floats = (int)c->verts.size() * 8;
    // the caller may modify the device copy (NCCL all-reduce): it is the authoritative one from now on
    return c->vstate_dev;
}

static void graph_set_time(Ctx* c, uint64_t time) {                         // graph.rs:123-128 + extensions.rs:196-204
    c->t = time;
    ++c->mut_gen;
    for (HVertex& hv : c->verts) {
        if (hv.kind == K_SAMPLE_LOOP) hv.loop_t = time;
        else if (hv.kind == K_DEBUG_SINE) hv.dnotes.clear();
        else if (hv.kind == K_SYNTH) hv.notes.clear();
        else if (hv.kind == K_BANDPASS) hv.bp_first = true;
    }
}
int tdg_set_time(tdg_ctx* h, uint64_t frame) { graph_set_time(C_, frame); return TDG_OK; }
// 'stop' + 'refresh' of a whole batch of unchanged projects in one call (ui_workflow.rs:153-160,
// state.rs:240-250,328-467): per context the flowws are reloaded (when given), every vertex returns to its
// constructor state, the transport and the floww cursor rewind to 0 and the normalize maxima reset —
// exactly tdg_floww_reset + tdg_floww_add... + tdg_reset_vertex_state + tdg_set_time(0) +
// tdg_floww_set_time(0) + tdg_reset_normalize on each, without 9 calls per project across the FFI.
// Flowws are flattened: context i owns entries [sum(flowws_per_ctx[0..i)), +flowws_per_ctx[i]) of t/note/vel/counts.
// 'stop' + 'refresh' of one unchanged project (ui_workflow.rs:153-160, state.rs:240-250,328-467); `at` = index of its first
// floww in the flattened arrays
static int rewind_one(tdg_ctx* h, const float* const* t, const float* const* note, const float* const* vel,
                      const uint64_t* counts, uint32_t n_flowws, size_t at) {
    if (t) {
        if (int e = tdg_floww_reset(h)) return e;
        for (uint32_t k = 0; k < n_flowws; ++k, ++at) {
            const int id = tdg_floww_add(h, t[at], note[at], vel[at], counts[at]);
            if (id < 0) return id;
        }
    }
    if (int e = tdg_reset_vertex_state(h)) return e;
    if (int e = tdg_set_time(h, 0)) return e;
    if (int e = tdg_floww_set_time(h, 0)) return e;
    return tdg_reset_normalize(h);
}
int tdg_rewind_many(tdg_ctx** hs, int n, const float* const* t, const float* const* note, const float* const* vel,
                    const uint64_t* counts, const uint32_t* flowws_per_ctx) {
    if (n < 0 || (n > 0 && !hs)) return TDG_ERR_INVALID;
    size_t at = 0;
    for (int i = 0; i < n; ++i) {
        const bool reload = t && flowws_per_ctx;
        if (int e = rewind_one(hs[i], reload ? t : nullptr, note, vel, counts, reload ? flowws_per_ctx[i] : 0, at)) return e;
        if (reload) at += flowws_per_ctx[i];
    }
    return TDG_OK;
}
uint64_t tdg_change_time(tdg_ctx* h, uint64_t delta, int plus) {            // graph.rs:130-135
    Ctx* c = C_;
    const uint64_t nt = plus ? c->t + delta : c->t - std::min(delta, c->t);
    graph_set_time(c, nt);
    return nt;
}
uint64_t tdg_get_time(tdg_ctx* h) { return h->c.t; }
int tdg_floww_set_time(tdg_ctx* h, uint64_t frame) { h->c.fb_frame = frame; ++h->c.mut_gen; return TDG_OK; }   // floww.rs:83-86

int tdg_enable_taps(tdg_ctx* h, int on) { h->c.taps = on != 0; ++h->c.mut_gen; return TDG_OK; }
int tdg_read_vertex_buffer(tdg_ctx* h, int vertex, uint64_t frame0, uint64_t n, float* l, float* r) {
    Ctx* c = C_;
    if (vertex < 0 || (size_t)vertex >= c->last_buf_off.size() || c->last_buf_off[vertex] < 0)
        return c->fail(TDG_ERR_INVALID, "vertex was not evaluated by the last pass");
    if (!c->taps || c->last_epoch != c->eng->epoch) return c->fail(TDG_ERR_INVALID, "taps are only valid right after a pass with taps enabled");
    if (frame0 + n > c->last_n_frames) return c->fail(TDG_ERR_CAPACITY, "tap shorter than request");
    cudaSetDevice(c->eng->device);
    const float* bl_ = reinterpret_cast<const float*>(c->eng->lanes[c->last_lane].work.p + c->last_buf_off[vertex]);
    CU(cudaMemcpy(l, bl_ + frame0, n * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(r, bl_ + c->last_n_pad + frame0, n * 4, cudaMemcpyDeviceToHost));
    return TDG_OK;
}

uint64_t tdg_render_bound_bytes(tdg_ctx* h, uint64_t chunks, uint32_t, uint32_t bd) { return chunks * h->c.bl * 2 * (bd / 8); }

static int valid_bd(Ctx* c, uint32_t bd) {
    if (!(bd == 8 || bd == 16 || bd == 24 || bd == 32)) return c->fail(TDG_ERR_BITDEPTH, "bitdepth not supported: choose 8, 16, 24 or 32");
    return TDG_OK;
}

int tdg_normalize_scan_many(tdg_ctx** hs, int n, const uint64_t* chunks) {
    std::vector<PassReq> reqs;
    for (int i = 0; i < n; ++i) {
        if (hs[i]->c.output < 0) continue;                                  // graph.rs:225-226: no output, no scan
        PassReq r; r.c = &hs[i]->c; r.is_scan = true; r.total = chunks[i]; r.emit0 = 0; r.emit_n = chunks[i]; r.keep0 = 0;
        reqs.push_back(std::move(r));
    }
    return run_pass_pipelined(reqs);
}
int tdg_normalize_scan(tdg_ctx* h, uint64_t chunks) { return tdg_normalize_scan_many(&h, 1, &chunks); }

int tdg_render_many(tdg_ctx** hs, int n, const uint64_t* chunks, const uint32_t* render_sr, const uint32_t* bitdepth,
                    uint8_t** out_pcm, const uint64_t* cap_bytes, uint64_t* frames_written) {
    std::vector<PassReq> reqs;
    for (int i = 0; i < n; ++i) {
        Ctx* c = &hs[i]->c;
        if (int e = valid_bd(c, bitdepth[i])) return e;
        if (frames_written) frames_written[i] = 0;
        PassReq r; r.c = c; r.is_scan = false; r.total = chunks[i]; r.emit0 = 0; r.emit_n = chunks[i]; r.keep0 = 0;
        r.render_sr = render_sr[i]; r.bd = bitdepth[i];
        r.out = out_pcm ? out_pcm[i] : nullptr; r.cap = cap_bytes ? cap_bytes[i] : 0;
        r.frames_written = frames_written ? frames_written + i : nullptr;
        reqs.push_back(std::move(r));
    }
    return run_pass_pipelined(reqs);
}
int tdg_normalize_render_many(tdg_ctx** hs, int n, const uint64_t* chunks, const uint32_t* render_sr, const uint32_t* bitdepth,
                              uint8_t** out_pcm, const uint64_t* cap_bytes, uint64_t* frames_written) {
    if (n <= 0) return TDG_OK;
    std::vector<PassReq> scans(n), renders(n);
    for (int i = 0; i < n; ++i) {
        Ctx* c = &hs[i]->c;
        if (int e = valid_bd(c, bitdepth[i])) return e;
        if (frames_written) frames_written[i] = 0;
        PassReq& a = scans[i];
        a.c = c; a.is_scan = true; a.total = chunks[i]; a.emit0 = 0; a.emit_n = chunks[i]; a.keep0 = 0;
        PassReq& r = renders[i];
        r.c = c; r.is_scan = false; r.total = chunks[i]; r.emit0 = 0; r.emit_n = chunks[i]; r.keep0 = 0;
        r.render_sr = render_sr[i]; r.bd = bitdepth[i];
        r.out = out_pcm ? out_pcm[i] : nullptr; r.cap = cap_bytes ? cap_bytes[i] : 0;
        r.frames_written = frames_written ? frames_written + i : nullptr;
    }
    return run_scan_render_pipelined(scans, renders);
}
// refresh + normalize + render of a batch of unchanged projects in ONE pipelined call: tdg_rewind_many's work is done
// group by group on the host while the GPU renders the groups in front (same results as the two calls in sequence)
int tdg_refresh_normalize_render_many(tdg_ctx** hs, int n, const float* const* t, const float* const* note, const float* const* vel,
                                      const uint64_t* counts, const uint32_t* flowws_per_ctx,
                                      const uint64_t* chunks, const uint32_t* render_sr, const uint32_t* bitdepth,
                                      uint8_t** out_pcm, const uint64_t* cap_bytes, uint64_t* frames_written) {
    if (n <= 0) return TDG_OK;
    if (!hs) return TDG_ERR_INVALID;
    const bool reload = t && flowws_per_ctx;
    std::vector<size_t> at(n + 1, 0);
    if (reload) for (int i = 0; i < n; ++i) at[i + 1] = at[i] + flowws_per_ctx[i];
    std::vector<PassReq> scans(n), renders(n);
    for (int i = 0; i < n; ++i) {
        Ctx* c = &hs[i]->c;
        if (int e = valid_bd(c, bitdepth[i])) return e;
        if (frames_written) frames_written[i] = 0;
        PassReq& a = scans[i];
        a.c = c; a.is_scan = true; a.total = chunks[i]; a.emit0 = 0; a.emit_n = chunks[i]; a.keep0 = 0;
        PassReq& r = renders[i];
        r.c = c; r.is_scan = false; r.total = chunks[i]; r.emit0 = 0; r.emit_n = chunks[i]; r.keep0 = 0;
        r.render_sr = render_sr[i]; r.bd = bitdepth[i];
        r.out = out_pcm ? out_pcm[i] : nullptr; r.cap = cap_bytes ? cap_bytes[i] : 0;
        r.frames_written = frames_written ? frames_written + i : nullptr;
    }
    const std::function<int(size_t, size_t)> refresh = [&](size_t i0, size_t i1) -> int {
        for (size_t i = i0; i < i1; ++i)
            if (int e = rewind_one(hs[i], reload ? t : nullptr, note, vel, counts, reload ? flowws_per_ctx[i] : 0, at[i])) return e;
        return TDG_OK;
    };
    return run_scan_render_pipelined(scans, renders, refresh);
}
int tdg_render(tdg_ctx* h, uint64_t chunks, uint32_t render_sr, uint32_t bitdepth, uint8_t* out_pcm, uint64_t cap_bytes,
               uint64_t* frames_written) {
    return tdg_render_many(&h, 1, &chunks, &render_sr, &bitdepth, out_pcm ? &out_pcm : nullptr, &cap_bytes, frames_written);
}

// Playback / stream feed (ui_workflow.rs:195-210, stream_workflow.rs:86-100): `chunks` x { Graph::render,
// Sample::interleave, FlowwBank::set_time_to_next_block } evaluated as ONE pass from the current transport
// position.  Nothing is rewound at the end: graph time, floww cursor and every vertex state (loop counters,
// hit deques, voices, envelope slots, filter state, running normalize max) stay where the last chunk left
// them, so consecutive calls produce the stream the reference's per-chunk loop produces.  The output vertex's
// planar f32 buffer is copied back and interleaved (L, R) on the host: the render-ahead ring of a player is
// filled `chunks` at a time.
// ---- asynchronous playback feed: the render-ahead of ui_workflow.rs:195-210 without blocking the player ---------------
// tdg_play_begin queues `chunks` more chunks (one device pass on the playback lane, then Sample::interleave on the
// device and a copy into a pinned ring slot, all stream ordered) and returns a ticket at once; the host-side transport
// state has already advanced, so the next begin can be issued immediately.  tdg_play_wait blocks until that block
// is in the ring and hands out a pointer to it (valid until kSlots further begins).  Any other pass of the context
// first collects its outstanding tickets' device work (drain_plays).
int tdg_play_begin(tdg_ctx* h, uint64_t chunks) {
    Ctx* c = C_;
    Engine* g = c->eng;
    if (c->output < 0) return c->fail(TDG_ERR_NO_OUTPUT, "no output vertex");
    if (chunks == 0) return c->fail(TDG_ERR_INVALID, "tdg_play_begin: zero chunks");
    cudaSetDevice(g->device);
    const int slot = g->play_next;
    if (PlayJob* old = g->play_slot[slot]) {
        if (!old->waited) return c->fail(TDG_ERR_CAPACITY, "playback ring full: collect older tickets with tdg_play_wait first");
        delete old; g->play_slot[slot] = nullptr;
    }
    g->play_next = (slot + 1) % kSlots;
    ++c->mut_gen;
    PlayJob* pj = new PlayJob();
    pj->c = c; pj->frames = chunks * c->bl;
    pj->J.reqs.resize(1);
    PassReq& r = pj->J.reqs[0];
    r.c = c; r.is_scan = false; r.play = true; r.total = chunks; r.emit0 = 0; r.emit_n = chunks; r.keep0 = 0;
    r.render_sr = c->sr; r.bd = 32;
    pj->J.lane = &g->lanes[kLanes - 1]; pj->J.slot = slot;
    int e = pass_plan(pj->J);
    if (e == TDG_OK) e = pass_submit(pj->J);
    if (e != TDG_OK) { cudaStreamSynchronize(pj->J.lane->stream); delete pj; return e; }
    pass_end_state(pj->J);
    // interleave on the device, then into the pinned ring slot; everything on the lane's stream, in order
    cudaStream_t s = pj->J.lane->stream;
    const uint64_t n = pj->frames;
    if (g->play_dev_cap[slot] < n) {
        if (g->play_dev[slot]) cudaFree(g->play_dev[slot]);
        g->play_dev[slot] = nullptr; g->play_dev_cap[slot] = 0;
        CU(cudaMalloc(&g->play_dev[slot], n * 8)); g->play_dev_cap[slot] = n;
    }
    if (g->play_ring_cap[slot] < n) {
        if (g->play_ring[slot]) cudaFreeHost(g->play_ring[slot]);
        g->play_ring[slot] = nullptr; g->play_ring_cap[slot] = 0;
        CU(cudaHostAlloc(&g->play_ring[slot], n * 8, cudaHostAllocDefault)); g->play_ring_cap[slot] = n;
    }
    if (!g->play_ev[slot]) CU(cudaEventCreateWithFlags(&g->play_ev[slot], cudaEventDisableTiming));
    const float* dl = reinterpret_cast<const float*>(pj->J.W + r.buf_off[c->output]);
    CU(launch_interleave(dl, dl + r.n_pad, g->play_dev[slot], n, s));
    CU(cudaMemcpyAsync(g->play_ring[slot], g->play_dev[slot], n * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(g->play_ev[slot], s));
    g->play_slot[slot] = pj;
    ++c->plays_in_flight; ++c->launches;
    return slot;
}
int tdg_play_wait(tdg_ctx* h, int ticket, const float** interleaved, uint64_t* frames) {
    Ctx* c = C_;
    Engine* g = c->eng;
    if (ticket < 0 || ticket >= kSlots || !g->play_slot[ticket] || g->play_slot[ticket]->c != c || g->play_slot[ticket]->waited)
        return c->fail(TDG_ERR_INVALID, "unknown playback ticket");
    PlayJob* pj = g->play_slot[ticket];
    cudaSetDevice(g->device);
    CU(cudaEventSynchronize(g->play_ev[ticket]));
    const int e = pass_wait(pj->J);
    pj->waited = true; --c->plays_in_flight;
    if (interleaved) *interleaved = g->play_ring[ticket];
    if (frames) *frames = pj->frames;
    return e;
}

int tdg_play(tdg_ctx* h, uint64_t chunks, float* out_interleaved, uint64_t cap_frames, uint64_t* frames_written) {
    Ctx* c = C_;
    if (frames_written) *frames_written = 0;
    if (c->output < 0) return c->fail(TDG_ERR_NO_OUTPUT, "no output vertex");     // the reference unwraps the Option (ui_workflow.rs:203)
    if (chunks == 0) return TDG_OK;
    const uint64_t n = chunks * c->bl;
    if (out_interleaved && cap_frames < n) return c->fail(TDG_ERR_CAPACITY, "output buffer too small");
    ++c->mut_gen;                                   // the transport moves: a kept scan no longer describes the next render
    std::vector<PassReq> reqs(1);
    PassReq& r = reqs[0];
    r.c = c; r.is_scan = false; r.play = true; r.total = chunks; r.emit0 = 0; r.emit_n = chunks; r.keep0 = 0;
    r.render_sr = c->sr; r.bd = 32;
    if (int e = run_pass(reqs)) return e;
    if (out_interleaved) {
        if (c->output >= (int)c->last_buf_off.size() || c->last_buf_off[c->output] < 0) return c->fail(TDG_ERR_INVALID, "output vertex was not evaluated");
        cudaSetDevice(c->eng->device);
        const float* dl = reinterpret_cast<const float*>(c->eng->lanes[c->last_lane].work.p + c->last_buf_off[c->output]);
        std::vector<float> planar(2 * n);
        CU(cudaMemcpy(planar.data(), dl, n * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(planar.data() + n, dl + c->last_n_pad, n * 4, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < n; ++i) { out_interleaved[2 * i] = planar[i]; out_interleaved[2 * i + 1] = planar[n + i]; }   // sample.rs:177-184
    }
    if (frames_written) *frames_written = n;
    return TDG_OK;
}

// time sharding: audio is evaluated for [chunk0 - warmup, chunk0 + chunks) only; the event-driven
// state advances over the whole pass of `total_chunks` chunks, exactly as a full pass would leave it.
// With a communicator (tdg_comm_init_rank) or virtual ranks (…_segments_local) normalize peaks and
// band-pass state cross the segment boundaries exactly; without one, `warmup_chunks` re-rendered
// chunks let the filters converge (legacy, approximate).
static int make_segment_req(Ctx* c, PassReq& r, bool is_scan, uint64_t total_chunks, uint64_t chunk0, uint64_t chunks,
                            uint64_t warmup_chunks) {
    if (chunk0 + chunks > total_chunks) return c->fail(TDG_ERR_INVALID, "segment exceeds the pass");
    const uint64_t w = std::min(warmup_chunks, chunk0);
    r.c = c; r.is_scan = is_scan; r.segment = true; r.emit0 = chunk0 - w; r.emit_n = chunks + w; r.keep0 = chunk0;
    r.total = total_chunks; r.render_sr = c->sr;
    return TDG_OK;
}
int tdg_scan_segment(tdg_ctx* h, uint64_t total_chunks, uint64_t chunk0, uint64_t chunks, uint64_t warmup_chunks) {
    Ctx* c = C_;
    if (c->output < 0) return TDG_OK;
    std::vector<PassReq> reqs(1);
    if (c->nccl_comm) warmup_chunks = 0;
    if (int e = make_segment_req(c, reqs[0], true, total_chunks, chunk0, chunks, warmup_chunks)) return e;
    if (c->nccl_comm) { reqs[0].x_mode = 2; reqs[0].x_ranks = c->x_ranks; reqs[0].x_rank = c->x_rank; }
    return run_pass(reqs);
}
int tdg_render_segment(tdg_ctx* h, uint64_t total_chunks, uint64_t chunk0, uint64_t chunks, uint64_t warmup_chunks,
                       uint32_t render_sr, uint32_t bitdepth, uint8_t* out_pcm, uint64_t cap_bytes, uint64_t* frames_written) {
    Ctx* c = C_;
    if (int e = valid_bd(c, bitdepth)) return e;
    std::vector<PassReq> reqs(1);
    if (c->nccl_comm) warmup_chunks = 0;
    if (int e = make_segment_req(c, reqs[0], false, total_chunks, chunk0, chunks, warmup_chunks)) return e;
    PassReq& r = reqs[0];
    if (render_sr) r.render_sr = render_sr;
    r.bd = bitdepth; r.out = out_pcm; r.cap = cap_bytes; r.frames_written = frames_written;
    if (c->nccl_comm) { r.x_mode = 2; r.x_ranks = c->x_ranks; r.x_rank = c->x_rank; }
    return run_pass(reqs);
}
// virtual ranks: n contexts holding the SAME project on one device evaluate n consecutive segments
// in one launch sequence, exchanging through device memory what NCCL would carry between GPUs
static int segments_local(tdg_ctx** hs, int n, bool is_scan, uint64_t total_chunks, const uint64_t* chunk0, const uint64_t* chunks,
                          uint32_t render_sr, uint32_t bitdepth, uint8_t** out_pcm, const uint64_t* cap_bytes, uint64_t* frames_written) {
    if (n <= 0) return TDG_ERR_INVALID;
    std::vector<PassReq> reqs(n);
    uint64_t expect = 0;
    for (int i = 0; i < n; ++i) {
        Ctx* c = &hs[i]->c;
        if (is_scan && c->output < 0) return TDG_OK;
        if (!is_scan) if (int e = valid_bd(c, bitdepth)) return e;
        if (chunk0[i] != expect) return c->fail(TDG_ERR_INVALID, "segments must be consecutive and start at chunk 0");
        expect += chunks[i];
        if (int e = make_segment_req(c, reqs[i], is_scan, total_chunks, chunk0[i], chunks[i], 0)) return e;
        reqs[i].x_mode = n > 1 ? 1 : 0; reqs[i].x_ranks = n; reqs[i].x_rank = i; reqs[i].bd = bitdepth;
        if (render_sr) reqs[i].render_sr = render_sr;
        if (!is_scan) {
            reqs[i].out = out_pcm ? out_pcm[i] : nullptr; reqs[i].cap = cap_bytes ? cap_bytes[i] : 0;
            reqs[i].frames_written = frames_written ? frames_written + i : nullptr;
        }
    }
    if (expect != total_chunks) return hs[0]->c.fail(TDG_ERR_INVALID, "segments must cover the whole pass");
    return run_pass(reqs);
}
int tdg_scan_segments_local(tdg_ctx** hs, int n, uint64_t total_chunks, const uint64_t* chunk0, const uint64_t* chunks) {
    return segments_local(hs, n, true, total_chunks, chunk0, chunks, 0, 16, nullptr, nullptr, nullptr);
}
int tdg_render_segments_local(tdg_ctx** hs, int n, uint64_t total_chunks, const uint64_t* chunk0, const uint64_t* chunks,
                              uint32_t render_sr, uint32_t bitdepth, uint8_t** out_pcm, const uint64_t* cap_bytes,
                              uint64_t* frames_written) {
    return segments_local(hs, n, false, total_chunks, chunk0, chunks, render_sr, bitdepth, out_pcm, cap_bytes, frames_written);
}
// where a segment's frames land in the rendered file: the host-side f64 replay of the resampler's index
// (no device work); render_sr >= project_sr (or 0) means no resampling (state.rs:533)
int tdg_segment_output_frames(uint32_t project_sr, uint32_t buffer_len, uint32_t render_sr, uint64_t chunk0, uint64_t chunks,
                              uint64_t* first_frame, uint64_t* n_frames) {
    if (!project_sr || !buffer_len) return TDG_ERR_INVALID;
    uint64_t first = chunk0 * buffer_len, n = chunks * buffer_len;
    if (render_sr && project_sr > render_sr) {
        std::vector<ResampleChunk> ch;
        resample_replay_slice((double)render_sr / (double)project_sr, buffer_len, chunk0, chunks, ch, &first, &n);
    }
    if (first_frame) *first_frame = first;
    if (n_frames) *n_frames = n;
    return TDG_OK;
}

// NCCL communicator of a time-sharded render: one rank per GPU / process
int tdg_comm_unique_id(uint8_t* out128) {
    NcclApi* na = nccl_api();
    if (!na->ok() || !out128) return TDG_ERR_CUDA;
    Id128 id; memset(&id, 0, sizeof id);
    if (na->GetUniqueId(&id) != 0) return TDG_ERR_CUDA;
    memcpy(out128, id.b, 128);
    return TDG_OK;
}
int tdg_comm_init_rank(tdg_ctx* h, int nranks, int rank, const uint8_t* id128) {
    Ctx* c = C_;
    NcclApi* na = nccl_api();
    if (!na->ok()) return c->fail(TDG_ERR_CUDA, "NCCL unavailable: " + na->err);
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) return c->fail(TDG_ERR_INVALID, "bad communicator arguments");
    if (c->nccl_comm) return c->fail(TDG_ERR_INVALID, "context already has a communicator");
    cudaSetDevice(c->eng->device);
    Id128 id; memcpy(id.b, id128, 128);
    void* comm = nullptr;
    NC(na->CommInitRank(&comm, nranks, id, rank));
    c->nccl_comm = comm; c->x_ranks = nranks; c->x_rank = rank;
    return TDG_OK;
}
int tdg_comm_destroy(tdg_ctx* h) {
    Ctx* c = C_;
    if (!c->nccl_comm) return TDG_OK;
    cudaSetDevice(c->eng->device);
    cudaStreamSynchronize(c->eng->stream);
    nccl_api()->CommDestroy(c->nccl_comm);
    c->nccl_comm = nullptr; c->x_ranks = 1; c->x_rank = 0;
    return TDG_OK;
}
uint64_t tdg_collectives(tdg_ctx* h) { return h ? h->c.collectives : 0; }

// pinned host memory for the caller's PCM sinks (cudaHostAlloc): copies into it are asynchronous and run at PCIe speed
void* tdg_host_alloc(uint64_t bytes) {
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void tdg_host_free(void* p) { if (p) cudaFreeHost(p); }
int tdg_set_pass_reuse(tdg_ctx* h, int on) { h->c.eng->reuse = on != 0; h->c.keep.valid = false; return TDG_OK; }
int tdg_set_oscillator_mode(tdg_ctx* h, int libm_exact) { h->c.eng->libm_sin = libm_exact != 0; h->c.keep.valid = false; ++h->c.mut_gen; return TDG_OK; }
int tdg_set_profiling(tdg_ctx* h, int on) { h->c.eng->profiling = on != 0; return TDG_OK; }
int tdg_set_inline_sums(tdg_ctx* h, int on) { h->c.eng->inline_sums = on != 0; h->c.keep.valid = false; return TDG_OK; }
int tdg_set_launch_waves(uint32_t waves) { set_launch_waves(waves); return TDG_OK; }
int tdg_set_pipeline_group(tdg_ctx* h, uint32_t projects_per_group) { h->c.eng->pipeline_group = projects_per_group; return TDG_OK; }
int tdg_last_pass_kernel_stats(tdg_ctx* h, tdg_kernel_stat* out, int cap) {
    static const char* names[KC_COUNT] = {"k_sample_loop", "k_sample_multi", "k_sample_lerp", "k_voice<debug_sine>", "k_voice<synth>",
        "k_voice<sampsyn>", "k_fx<sum>", "k_fx<adsr>", "k_fx<norm_peak>", "k_norm_prefix", "k_norm_scale", "k_bp_aggregate", "k_bp_carry", "k_bp_apply", "k_resample", "k_quantise", "nccl_exchange", "k_chunk_ranges", "k_bp_fused"};
    int n = 0;
    for (int k = 0; k < KC_COUNT && n < cap; ++k) {
        const KStat& s = h->c.kstats[k];
        if (!s.launches) continue;
        out[n].name = names[k]; out[n].launches = s.launches; out[n].device_ms = s.ms; out[n].algorithmic_bytes = s.bytes;
        ++n;
    }
    return n;
}
int tdg_last_pass_stats(tdg_ctx* h, tdg_pass_stats* out) {
    if (!h || !out) return TDG_ERR_INVALID;
    const PassStats& s = h->c.stats;
    out->device_ms = s.device_ms; out->plan_ms = s.plan_ms; out->h2d_ms = s.h2d_ms; out->d2h_ms = s.d2h_ms;
    out->h2d_bytes = s.h2d_bytes; out->d2h_bytes = s.d2h_bytes; out->kernel_launches = s.kernel_launches;
    out->algorithmic_bytes = s.algorithmic_bytes; out->frames = s.frames; out->voice_frames = s.voice_frames;
    out->reused_vertices = s.reused_vertices; out->reused_bytes = s.reused_bytes;
    out->inlined_vertices = s.inlined_vertices; out->inlined_bytes = s.inlined_bytes;
    return TDG_OK;
}

}  // extern "C"
