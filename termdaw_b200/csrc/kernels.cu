// kernels.cu — the sm_100a kernels of the termdaw render path and their launchers.
//
// Layout: every vertex buffer is planar stereo fp32 in HBM for the whole pass; a block evaluates one
// chunk (buffer_length frames) of one task, a thread four consecutive frames (float4 loads/stores),
// grid = (chunks, tasks) so that one launch covers a whole DAG level of a batch of projects.
// All kernels are HBM-streaming (no dense contraction anywhere => no tensor cores); the voice
// kernels are FP32/SFU bound and say so in DESIGN.md.
//
// Compiled with -fmad=false: the reference rounds after every f32 operation.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "tdg_types.h"
#include "tdg_math.cuh"
#include "kernels.h"

namespace tdg {

#define TDG_THREADS 256
#ifndef TDG_PDL_DEFAULT
#define TDG_PDL_DEFAULT 1
#endif

// ------------------------------------------------------------------------------------ helpers
template <typename T>
__device__ __forceinline__ void load_task(T* dst, const T* src) {
    static_assert(sizeof(T) % 4 == 0, "task structs are word multiples");
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    for (int i = threadIdx.x; i < (int)(sizeof(T) / 4); i += blockDim.x) d[i] = s[i];
    __syncthreads();
}

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization attribute may become
// resident while the kernel in front of it in the stream is still draining.  Every pass kernel lets its dependents
// launch at once (they only fill SM slots the running grid no longer needs) and waits for the grid in front of it —
// complete, memory flushed — right after it has fetched its task and BEFORE it touches anything a kernel produces.
// Without the attribute both instructions do nothing.
__device__ __forceinline__ void pdl_launch_dependents() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// task fetch of a pass kernel: the task arrays are uploaded before the pass's first kernel, no kernel writes them
template <typename T>
__device__ __forceinline__ void load_task_pdl(T* dst, const T* src) {
    pdl_launch_dependents();
    load_task(dst, src);
    pdl_wait();
}

__device__ __forceinline__ void apply_pg(float& l, float& r, const PanGain& pg) {
    if (pg.flags & 1u) { l *= pg.l_amp; r *= pg.r_amp; }       // apply_angle, sample.rs:97-106
    if (pg.flags & 2u) { l *= pg.gain; r *= pg.gain; }         // apply_gain,  sample.rs:108-114
}

template <bool VEC>
__device__ __forceinline__ void ld4(const float* __restrict__ p, uint64_t idx, int nv, float (&v)[4]) {
    if (VEC) {
        float4 a = *reinterpret_cast<const float4*>(p + idx);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = j < nv ? p[idx + j] : 0.0f;
    }
}
template <bool VEC>
__device__ __forceinline__ void st4(float* __restrict__ p, uint64_t idx, int nv, const float (&v)[4]) {
    if (VEC) {
        *reinterpret_cast<float4*>(p + idx) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < nv) p[idx + j] = v[j];
    }
}

// absmaxlen's fold (sample.rs:12-14): NaN never wins because `a > max` is false for NaN
__device__ __forceinline__ float absmax_fold(float m, float x) { float a = fabsf(x); return a > m ? a : m; }

__device__ __forceinline__ float block_max(float m) {
    __shared__ float wm[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
    if (ln == 0) wm[w] = m;
    __syncthreads();
    if (w == 0) {
        m = ln < (int)(blockDim.x >> 5) ? wm[ln] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    return m;   // valid in thread 0
}

// ------------------------------------------------------------------------------------ K0: clocks and per-chunk ranges
// clock[n][i] = -(i as f32 / sr as f32) after n additions of (bl as f32 / sr as f32): the env_t of every voice and
// the t of every Adsr slot n chunks after the event that (re)started it at frame offset i (tdg_types.h).
// One thread per column, the additions in the reference's order; rows [row0, depth) are (re)built from row0 - 1.
__global__ void __launch_bounds__(256) k_env_table(float* __restrict__ clock, uint32_t bl, uint32_t row0, uint32_t depth, float fsr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bl) return;
    const float d = (float)bl / fsr;
    float v = row0 == 0 ? -((float)i / fsr) : clock[(size_t)(row0 - 1) * bl + i];
    if (row0 == 0) { clock[i] = v; }
    for (uint32_t n = row0 == 0 ? 1 : row0; n < depth; ++n) { v += d; clock[(size_t)n * bl + i] = v; }
}
__device__ __forceinline__ float clock_at(uint32_t src, uint32_t n, const float* __restrict__ clock, uint32_t bl, const float* __restrict__ xenv) {
    return (src & kEnvExplicit) ? xenv[(src & ~kEnvExplicit) + n] : __ldg(clock + (size_t)n * bl + src);
}
// first index in [0, n) for which pred(index) holds (pred monotone: false... true), n if none
template <typename P> __device__ __forceinline__ uint32_t first_true(uint32_t n, P pred) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (pred(m)) hi = m; else lo = m + 1; }
    return lo;
}
// The (up to two: held part, release part) segments a voice has inside pass chunk k (extensions.rs:481-496): returns
// their number and, with `seg` given, writes them (VoiceSeg layout + the third oscillator's release start value).
__device__ __forceinline__ uint32_t voice_segs_in_chunk(const VoiceRec& R, uint32_t k, uint32_t bl, const float* __restrict__ clock,
                                                        const float* __restrict__ xenv, float4* __restrict__ seg, int debug_sine,
                                                        float* __restrict__ relv2, uint32_t* written) {
    (void)written;
    if (!(R.on_chunk <= k && k <= R.end_chunk)) return 0;
    uint32_t cnt = 0;
    const uint32_t end_here = k == R.end_chunk ? R.end_i : bl;
    if (k <= R.off_chunk) {                                          // held part (off_chunk == kNever: never released)
        const uint32_t ia = k == R.on_chunk ? R.on_i : 0u;
        const uint32_t ib = k == R.off_chunk ? R.off_i : end_here;
        if (ib > ia) {
            if (seg) {
                const float env_t = debug_sine ? 0.0f : clock_at(R.envh, k - R.on_chunk, clock, bl, xenv);
                seg[0] = make_float4(R.hz, R.vel, env_t, 0.0f);
                seg[1] = make_float4(__uint_as_float(ia), __uint_as_float(ib), 0.0f, 0.0f);
                relv2[0] = 0.0f;
            }
            cnt = 1;
        }
    }
    if (R.off_chunk <= k) {                                          // release part
        const uint32_t ia = k == R.off_chunk ? R.off_i : 0u;
        if (end_here > ia) {
            if (seg) {
                const float env_t = clock_at(R.envr, k - R.off_chunk, clock, bl, xenv);
                seg[2 * cnt] = make_float4(R.hz, R.vel, env_t, R.rel_t);
                seg[2 * cnt + 1] = make_float4(__uint_as_float(ia), __uint_as_float(end_here), R.relv0, R.relv1);
                relv2[cnt] = R.relv2;
            }
            ++cnt;
        }
    }
    return cnt;
}
// The per-chunk index ranges of every event-driven vertex of a pass (voices, adsr, sample_multi, sample_lerp), one
// thread per (task, chunk): the host uploads O(events) records, the O(chunks) tables are made here.
__global__ void __launch_bounds__(256) k_chunk_ranges(const RangeTask* __restrict__ tasks, const uint8_t* __restrict__ blob) {
    pdl_launch_dependents();
    const RangeTask T = tasks[blockIdx.y];
    pdl_wait();
    const uint32_t kk = blockIdx.x * blockDim.x + threadIdx.x;
    if (kk >= T.n_chunks) return;
    const uint32_t n = T.n_recs;
    Range2 out{0u, 0u};
    if (T.kind == RK_VOICE) {
        const VoiceRec* R = reinterpret_cast<const VoiceRec*>(blob + T.recs_off);
        const uint32_t k = T.chunk0 + kk;
        out.x = first_true(n, [&](uint32_t v) { return R[v].pm_end >= k; });
        out.y = first_true(n, [&](uint32_t v) { return R[v].sm_on > k; });
        if (out.y < out.x) out.y = out.x;
        uint32_t cnt = 0;                                   // segments the chunk will hold: k_seg_scan turns the counts into offsets
        for (uint32_t v = out.x; v < out.y; ++v) cnt += voice_segs_in_chunk(R[v], k, T.bl, nullptr, nullptr, nullptr, 0, nullptr, nullptr);
        T.seg_ptr[kk + 1] = cnt;
        if (kk == 0) T.seg_ptr[0] = 0;
    } else if (T.kind == RK_ADSR) {
        const AdsrRec* R = reinterpret_cast<const AdsrRec*>(blob + T.recs_off);
        const uint64_t c0 = (uint64_t)(T.chunk0 + kk) * T.bl;
        const uint32_t a = first_true(n, [&](uint32_t v) { return R[v].start > c0; });            // record 0 starts at frame 0
        const uint32_t b = first_true(n, [&](uint32_t v) { return R[v].start >= c0 + T.bl; });
        out.x = a - 1; out.y = b - (a - 1);
        T.seg_ptr[kk + 1] = out.y;                          // one envelope state per record active in the chunk
        if (kk == 0) T.seg_ptr[0] = 0;
    } else if (T.kind == RK_MULTI) {
        const MultiHit* H = reinterpret_cast<const MultiHit*>(blob + T.recs_off);
        const int64_t c0 = (int64_t)kk * T.bl;                                                      // hits are relative to the emitted range
        out.x = first_true(n, [&](uint32_t v) { return H[v].start > c0 - T.len; });
        out.y = first_true(n, [&](uint32_t v) { return H[v].start > c0 + (int64_t)T.bl - 1; });
        if (out.y < out.x) out.y = out.x;
    } else {
        const LerpHit* H = reinterpret_cast<const LerpHit*>(blob + T.recs_off);
        const int64_t c0 = (int64_t)kk * T.bl;
        out.x = first_true(n, [&](uint32_t v) { return H[v].frame > c0; }) - 1;                     // entries 0, 1: the carried pair
    }
    T.out[kk] = out;
}

// counts -> offsets: seg_ptr[0] = 0, seg_ptr[k + 1] = segments of chunks [0, k].  One block per voice task.
__global__ void __launch_bounds__(1024) k_seg_scan(const RangeTask* __restrict__ tasks) {
    pdl_launch_dependents();
    const RangeTask T = tasks[blockIdx.x];
    pdl_wait();
    if (T.kind != RK_VOICE && T.kind != RK_ADSR) return;
    __shared__ uint32_t part[1024];
    const uint32_t n = T.n_chunks, tid = threadIdx.x, nt = blockDim.x;
    const uint32_t per = (n + nt - 1) / nt;
    const uint32_t lo = min(n, tid * per), hi = min(n, lo + per);
    uint32_t sum = 0;
    for (uint32_t k = lo; k < hi; ++k) sum += T.seg_ptr[k + 1];
    part[tid] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < nt; o <<= 1) {               // inclusive Hillis-Steele over per-thread sums
        const uint32_t v = tid >= o ? part[tid - o] : 0u;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    uint32_t run = tid > 0 ? part[tid - 1] : 0u;
    for (uint32_t k = lo; k < hi; ++k) { run += T.seg_ptr[k + 1]; T.seg_ptr[k + 1] = run; }
}
// the segments themselves: one thread per chunk walks the chunk's candidate records in push order
__global__ void __launch_bounds__(256) k_seg_fill(const RangeTask* __restrict__ tasks, const uint8_t* __restrict__ blob) {
    pdl_launch_dependents();
    const RangeTask T = tasks[blockIdx.y];
    pdl_wait();
    if (T.kind != RK_VOICE && T.kind != RK_ADSR) return;
    const uint32_t kk = blockIdx.x * blockDim.x + threadIdx.x;
    if (kk >= T.n_chunks) return;
    const float* xenv = reinterpret_cast<const float*>(blob + T.xenv_off);
    const uint32_t k = T.chunk0 + kk;
    const Range2 rg = T.out[kk];
    uint32_t at = T.seg_ptr[kk];
    if (T.kind == RK_ADSR) {                            // Adsr vertex: the states the chunk passes through, clocks looked up
        const AdsrRec* A = reinterpret_cast<const AdsrRec*>(blob + T.recs_off) + rg.x;
        AdsrSeg* out = reinterpret_cast<AdsrSeg*>(T.segs) + at;
        const uint64_t c0 = (uint64_t)k * T.bl;
        for (uint32_t q = 0; q < rg.y; ++q) {
            const AdsrRec R = A[q];
            AdsrSeg sg;
            sg.ia = R.start > c0 ? (uint32_t)(R.start - c0) : 0u; sg.flags = R.flags;
            sg.pt = clock_at(R.p_src, k - R.p_chunk, T.clock, T.bl, xenv); sg.pv = R.pv; sg.pr = R.pr;
            sg.gt = clock_at(R.g_src, k - R.g_chunk, T.clock, T.bl, xenv); sg.gv = R.gv; sg.gr = R.gr;
            out[q] = sg;
        }
        return;
    }
    const VoiceRec* R = reinterpret_cast<const VoiceRec*>(blob + T.recs_off);
    for (uint32_t v = rg.x; v < rg.y; ++v)
        at += voice_segs_in_chunk(R[v], k, T.bl, T.clock, xenv, reinterpret_cast<float4*>(T.segs) + 2 * (size_t)at, (int)T.debug_sine, T.relv2 + at, nullptr);
}

// ------------------------------------------------------------------------------------ K4: fx
// sum_inputs (extensions.rs:310-319) + { nothing | adsr_gen (:593-651) | per-chunk peak for
// normalize_gen (:321-329) } + apply_angle + apply_gain (:262-263), one pass over HBM.
#define TDG_FX_SEGS 32
template <bool EXACT>
__device__ __forceinline__ float adsr_frame_vel(const FxTask& T, const AdsrSeg& sg, float offset, float maxmul, float minmul) {
    float pvel, gvel;
    if (T.use_off) {                                     // extensions.rs:620-623
        pvel = (sg.pr == 0.0f ? k_apply_ads<EXACT>(T.conf, sg.pt + offset) : k_apply_r<EXACT>(T.conf, sg.pt + offset, sg.pr)) * sg.pv;
        gvel = (sg.gr == 0.0f ? k_apply_ads<EXACT>(T.conf, sg.gt + offset) : k_apply_r<EXACT>(T.conf, sg.gt + offset, sg.gr)) * sg.gv;
    } else {                                             // extensions.rs:640-641
        pvel = k_apply_adsr<EXACT>(T.conf, sg.pt + offset) * sg.pv;
        gvel = k_apply_adsr<EXACT>(T.conf, sg.gt + offset) * sg.gv;
    }
    const float adsr_vel = fmaxf(pvel, gvel) * maxmul + fminf(pvel, gvel) * minmul;
    return 1.0f + T.wet * (adsr_vel - 1.0f);             // lerp(1.0, adsr_vel, wet)
}

// (Blocks of 128 threads owning two 4-frame runs per thread, both loaded ahead, were measured at 12.7 ms per bench step
// against 7.2 for this shape: the second pair of accumulators pushes the envelope code into local memory.)
#ifndef TDG_ADSR_MINB
#define TDG_ADSR_MINB 5      // measured: 4 -> 9.0 ms, 5 -> 8.3 ms, 6 -> 8.5 ms per bench step
#endif
#ifndef TDG_FXG_CPB
#define TDG_FXG_CPB 4        // chunks per block of the grouped sum / norm_peak instantiations (longer chain at the head of a block: group records, pointer list)
#endif
#ifndef TDG_FXG_MINB
#define TDG_FXG_MINB 5       // resident blocks per SM of the grouped sum / norm_peak instantiations
#endif
#ifndef TDG_FX_CPB
#define TDG_FX_CPB 1         // chunks per block of k_fx<sum> / k_fx<norm_peak> (both already run at the HBM roof)
#endif
#ifndef TDG_ADSR_CPB
#define TDG_ADSR_CPB 8       // chunks per block of k_fx<adsr>: measured 1 -> 7.19, 2 -> 6.65, 4 -> 6.24, 8 -> 6.03 ms per bench step
#endif
#ifndef TDG_BP_MINB
#define TDG_BP_MINB 5        // measured: 4 -> 6.66 ms, 5 -> 6.48 ms, 6 -> 6.95 ms
#endif
// GRP: the task list holds vertices some of whose inputs are inlined Sum vertices (FxGroup records; engine.cu "inlined
// sums").  A separate instantiation, so that the plain kernels keep their register budget.
template <int MODE, bool VEC, bool EXACT, bool GRP>
__global__ void __launch_bounds__(TDG_THREADS, MODE == FX_ADSR ? (GRP ? 4 : TDG_ADSR_MINB) : GRP ? TDG_FXG_MINB : 6) k_fx(const FxTask* __restrict__ tasks, const uint8_t* __restrict__ blob, uint32_t cpb) {
    __shared__ FxTask T;
    __shared__ __align__(16) FxGroup s_grp[GRP ? kFxGroupsStaged : 1];
    load_task_pdl(&T, tasks + blockIdx.y);
    if (T.mode != (uint32_t)MODE) return;                // one launch per mode over the same task list
    __shared__ const float* s_ins[GRP ? 2 * kFxPairsStaged : 1];
    uint32_t n_grp_s = 0;                                // group records staged in shared memory (one dependent global load less per edge)
    bool ins_s = false;                                  // ... and the expanded pointer list, which usually outgrows the task's inline copy
    if (GRP) {
        if (T.n_grp <= (uint32_t)kFxGroupsStaged) {
            n_grp_s = T.n_grp;
            const uint32_t* src = reinterpret_cast<const uint32_t*>(blob + T.grp_off);
            uint32_t* dst = reinterpret_cast<uint32_t*>(s_grp);
            for (uint32_t w = threadIdx.x; w < n_grp_s * (uint32_t)(sizeof(FxGroup) / 4); w += blockDim.x) dst[w] = src[w];
        }
        if (T.n_in > (uint32_t)kInlineInputs && T.n_in <= (uint32_t)kFxPairsStaged) {
            ins_s = true;
            const float* const* src = reinterpret_cast<const float* const*>(blob + T.in_off);
            for (uint32_t w = threadIdx.x; w < 2 * T.n_in; w += blockDim.x) s_ins[w] = src[w];
        }
        __syncthreads();
    }
    // the adsr variant walks TDG_ADSR_CPB consecutive chunks per block (see the gather kernels: the task load, its
    // barrier and the block launch are paid once per group of chunks)
    constexpr uint32_t CPB = MODE == FX_ADSR ? TDG_ADSR_CPB : GRP ? TDG_FXG_CPB : TDG_FX_CPB;      // compile-time maximum; cpb <= CPB is the launch's choice
    const uint32_t kc0 = blockIdx.x * (CPB == 1 ? 1u : cpb);
    if (kc0 >= T.h.n_chunks) return;
    const uint32_t kc1 = CPB == 1 ? kc0 + 1 : min(kc0 + cpb, T.h.n_chunks);
    const uint32_t bl = T.h.bl;
    uint64_t p0 = 0;
    const bool staged = T.n_in <= (uint32_t)kInlineInputs;     // input pointers arrived with the task (shared memory)
    const float* const* ins = staged ? T.ins : ins_s ? s_ins : reinterpret_cast<const float* const*>(blob + T.in_off);
    // sum_inputs for the four frames at chunk offset i (edge order: the f32 sum is order dependent)
    auto load_sum = [&](uint32_t i, float (&l)[4], float (&r)[4]) {
        const int nv = min(4u, bl - i);
        const uint64_t p = p0 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) { l[j] = 0.0f; r[j] = 0.0f; }
        if (GRP) {
            // some edges are inlined Sum vertices (FxGroup): an edge's value is either a buffer or the sum of a
            // vertex's inputs from +0.0 in its edge order with its pan and gain applied — k_fx<FX_SUM>'s operations on
            // that vertex, in registers — and the edges are added in this vertex's edge order as always
            const FxGroup* G = n_grp_s ? s_grp : reinterpret_cast<const FxGroup*>(blob + T.grp_off);
            for (uint32_t q = 0; q < T.n_grp; ++q) {
                const uint4 ga = *reinterpret_cast<const uint4*>(G + q);               // first, count, nested
                float x[4], y[4];
                if (!ga.z) {
                    ld4<VEC>(ins[2 * ga.x], p, nv, x);
                    ld4<VEC>(ins[2 * ga.x + 1], p, nv, y);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { x[j] = 0.0f; y[j] = 0.0f; }
#pragma unroll 2
                    for (uint32_t b = 0; b < ga.y; ++b) {
                        float u[4], w[4];
                        ld4<VEC>(ins[2 * (ga.x + b)], p, nv, u);
                        ld4<VEC>(ins[2 * (ga.x + b) + 1], p, nv, w);
#pragma unroll
                        for (int j = 0; j < 4; ++j) { x[j] += u[j]; y[j] += w[j]; }
                    }
                    const float4 pf = *(reinterpret_cast<const float4*>(G + q) + 1);
                    PanGain pg; pg.l_amp = pf.x; pg.r_amp = pf.y; pg.gain = pf.z; pg.flags = __float_as_uint(pf.w);
#pragma unroll
                    for (int j = 0; j < 4; ++j) apply_pg(x[j], y[j], pg);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) { l[j] += x[j]; r[j] += y[j]; }
            }
        } else if (staged) {
#pragma unroll 2
            for (uint32_t a = 0; a < T.n_in; ++a) {
                float x[4], y[4];
                ld4<VEC>(T.ins[2 * a], p, nv, x);
                ld4<VEC>(T.ins[2 * a + 1], p, nv, y);
#pragma unroll
                for (int j = 0; j < 4; ++j) { l[j] += x[j]; r[j] += y[j]; }
            }
        } else {
            for (uint32_t a = 0; a < T.n_in; ++a) {
                float x[4], y[4];
                ld4<VEC>(ins[2 * a], p, nv, x);
                ld4<VEC>(ins[2 * a + 1], p, nv, y);
#pragma unroll
                for (int j = 0; j < 4; ++j) { l[j] += x[j]; r[j] += y[j]; }
            }
        }
    };
    float l[4], r[4];
    const uint32_t i_first = threadIdx.x * 4;
    float maxmul = 0.0f, minmul = 0.0f;
    if (MODE == FX_ADSR) { maxmul = T.use_max ? 1.0f : 0.0f; minmul = 1.0f - maxmul; }
#pragma unroll 1
    for (uint32_t k = kc0; k < kc1; ++k) {
    p0 = (uint64_t)k * bl;
    bool have = false;
    uint32_t n_seg = 0;
    const AdsrSeg* __restrict__ segs = nullptr;
    // state q of the chunk, straight from the device-built table (the same address for the whole block: one
    // L1-resident broadcast per read, no staging and no barrier: the block's warps never wait for each other)
    auto seg_ia = [&](uint32_t q) -> uint32_t { return __ldg(&segs[q].ia); };
    auto seg_get = [&](uint32_t q) -> AdsrSeg {
        const float4 a = __ldg(reinterpret_cast<const float4*>(segs + q)), b = __ldg(reinterpret_cast<const float4*>(segs + q) + 1);
        AdsrSeg sg;
        sg.ia = __float_as_uint(a.x); sg.flags = __float_as_uint(a.y); sg.pt = a.z; sg.pv = a.w; sg.pr = b.x; sg.gt = b.y; sg.gv = b.z; sg.gr = b.w;
        return sg;
    };
    if (MODE == FX_ADSR) {
        const uint32_t s_lo = T.seg_ptr[k];
        n_seg = T.seg_ptr[k + 1] - s_lo;
        segs = T.segs + s_lo;
        if (i_first < bl) { load_sum(i_first, l, r); have = true; }
    }
    float peak = 0.0f;
    for (uint32_t i = i_first; i < bl; i += TDG_THREADS * 4) {
        const int nv = min(4u, bl - i);
        const uint64_t p = p0 + i;
        if (!have) load_sum(i, l, r);
        have = false;
        if (MODE == FX_ADSR) {
            // last segment starting at or before the thread's first frame; the common case is that the
            // same state covers all four frames
            uint32_t s = 0;
            while (s + 1 < n_seg && seg_ia(s + 1) <= i) ++s;
            const bool one = s + 1 >= n_seg || seg_ia(s + 1) >= i + 4;
            if (one) {
                const AdsrSeg sg = seg_get(s);
                if (!(sg.flags & 1u)) {
                    // plateau test on the run's first frame only: most runs sit between two hits with both
                    // envelopes finished (or held), and then the whole run has one gain
                    const float o0 = k_div<EXACT>((float)i, T.fsr, T.rsr);
                    const int pm = T.use_off ? (sg.pr == 0.0f ? 1 : 2) : 0, gm = T.use_off ? (sg.gr == 0.0f ? 1 : 2) : 0;
                    float pval, gval;
                    if (k_env_plateau<EXACT>(T.conf, pm, sg.pr, sg.pt + o0, &pval) &&
                        k_env_plateau<EXACT>(T.conf, gm, sg.gr, sg.gt + o0, &gval)) {
                        const float pvel = pval * sg.pv, gvel = gval * sg.gv;
                        const float adsr_vel = fmaxf(pvel, gvel) * maxmul + fminf(pvel, gvel) * minmul;
                        const float vel = 1.0f + T.wet * (adsr_vel - 1.0f);      // lerp(1.0, adsr_vel, wet)
#pragma unroll
                        for (int j = 0; j < 4; ++j) { l[j] *= vel; r[j] *= vel; }
                    } else {
                        float tp[4], tg[4], ep[4], eg[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float offset = k_div<EXACT>((float)(i + j), T.fsr, T.rsr);
                            tp[j] = sg.pt + offset; tg[j] = sg.gt + offset;
                        }
                        if (T.use_off) {                     // extensions.rs:620-623
                            if (sg.pr == 0.0f) k_apply_ads_n<EXACT, 4>(T.conf, tp, ep); else k_apply_r_n<EXACT, 4>(T.conf, tp, sg.pr, ep);
                            if (sg.gr == 0.0f) k_apply_ads_n<EXACT, 4>(T.conf, tg, eg); else k_apply_r_n<EXACT, 4>(T.conf, tg, sg.gr, eg);
                        } else {                             // extensions.rs:640-641
                            k_apply_adsr_n<EXACT, 4>(T.conf, tp, ep);
                            k_apply_adsr_n<EXACT, 4>(T.conf, tg, eg);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float pvel = ep[j] * sg.pv, gvel = eg[j] * sg.gv;
                            const float adsr_vel = fmaxf(pvel, gvel) * maxmul + fminf(pvel, gvel) * minmul;
                            const float vel = 1.0f + T.wet * (adsr_vel - 1.0f);      // lerp(1.0, adsr_vel, wet)
                            l[j] *= vel; r[j] *= vel;
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    while (s + 1 < n_seg && seg_ia(s + 1) <= i + j) ++s;
                    const AdsrSeg sg = seg_get(s);
                    if (sg.flags & 1u) continue;             // un-enveloped frame (extensions.rs:633-635)
                    const float vel = adsr_frame_vel<EXACT>(T, sg, k_div<EXACT>((float)(i + j), T.fsr, T.rsr), maxmul, minmul);
                    if (j == 0) { l[0] *= vel; r[0] *= vel; } else if (j == 1) { l[1] *= vel; r[1] *= vel; }
                    else if (j == 2) { l[2] *= vel; r[2] *= vel; } else { l[3] *= vel; r[3] *= vel; }
                }
            }
        }
        if (MODE == FX_NORM_PEAK) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < nv) { peak = absmax_fold(peak, l[j]); peak = absmax_fold(peak, r[j]); }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) apply_pg(l[j], r[j], T.h.pg);
        }
        if (MODE != FX_NORM_PEAK || !T.no_store) {
            st4<VEC>(T.h.out_l, p, nv, l);
            st4<VEC>(T.h.out_r, p, nv, r);
        }
    }
    if (MODE == FX_NORM_PEAK) {
        peak = block_max(peak);
        if (threadIdx.x == 0) T.peaks[k] = peak;
    }
    }
}

// ------------------------------------------------------------------------------------ K8: normalize max bookkeeping
// One block per normalize vertex.  Scan pass: every chunk is scaled by the OLD max while scan_max
// collects the peak (extensions.rs:323-324, graph.rs:227-234).  Render pass: running max
// (extensions.rs:326).  max() is exact, so the parallel prefix is bit-identical to the fold.
// segment peak into the exchange slot (time sharding only)
__global__ void __launch_bounds__(256) k_norm_total(const NormPrefixTask* __restrict__ tasks, float* __restrict__ reduce_buf) {
    const NormPrefixTask T = tasks[blockIdx.x];
    float m = 0.0f;
    for (uint32_t k = threadIdx.x; k < T.n_chunks; k += blockDim.x) m = fmaxf(T.peaks[k], m);
    m = block_max(m);
    if (threadIdx.x == 0) {
        T.x.buf[(size_t)T.x.rank * T.x.stride + T.x.index] = m;
        if (reduce_buf) reduce_buf[T.x.index] = m;
    }
}
__global__ void __launch_bounds__(1024) k_norm_prefix(const NormPrefixTask* __restrict__ tasks) {
    pdl_launch_dependents();
    const NormPrefixTask T = tasks[blockIdx.x];
    pdl_wait();
    __shared__ float part[1024];
    const uint32_t n = T.n_chunks, tid = threadIdx.x, nt = blockDim.x;
    const uint32_t per = (n + nt - 1) / nt;
    const uint32_t lo = min(n, tid * per), hi = min(n, lo + per);
    const float old_max = T.state[0];
    float m = 0.0f;
    for (uint32_t k = lo; k < hi; ++k) m = fmaxf(T.peaks[k], m);
    part[tid] = m;
    __syncthreads();
    for (uint32_t o = 1; o < nt; o <<= 1) {           // inclusive Hillis-Steele over per-thread maxima
        float v = tid >= o ? part[tid - o] : 0.0f;
        __syncthreads();
        part[tid] = fmaxf(part[tid], v);
        __syncthreads();
    }
    float total = part[nt - 1];                       // peak of the whole pass
    float before = old_max;                           // max entering this segment's first chunk (render)
    if (T.x.ranks > 1) {                              // segments of other ranks (max is exact: order is free)
        for (uint32_t q = 0; q < T.x.ranks; ++q) {
            const float v = T.x.buf[(size_t)q * T.x.stride + T.x.index];
            total = fmaxf(total, v);
            if (q < T.x.rank) before = fmaxf(before, v);
        }
        if (T.reduced) total = fmaxf(total, T.reduced[T.x.index]);
    }
    if (T.is_scan) {
        for (uint32_t k = lo; k < hi; ++k) T.runmax[k] = old_max;
        __syncthreads();
        if (tid == 0) { const float sm = fmaxf(total, 0.0f); T.state[1] = sm; T.state[0] = sm; }
    } else {
        float run = fmaxf(tid > 0 ? part[tid - 1] : 0.0f, before);   // buf_max.max(*max)
        for (uint32_t k = lo; k < hi; ++k) { run = fmaxf(T.peaks[k], run); T.runmax[k] = run; }
        __syncthreads();
        if (tid == 0) T.state[0] = fmaxf(total, old_max);
    }
}

// K9a: buf.scale(len, 1.0 / max) (extensions.rs:328) + pan + gain, in place
template <bool VEC>
__global__ void __launch_bounds__(TDG_THREADS) k_norm_scale(const NormScaleTask* __restrict__ tasks) {
    __shared__ NormScaleTask T;
    load_task_pdl(&T, tasks + blockIdx.y);
    const uint32_t k = blockIdx.x;
    if (k >= T.h.n_chunks) return;
    const uint32_t bl = T.h.bl;
    const uint64_t p0 = (uint64_t)k * bl;
    const float f = 1.0f / T.runmax[k];
    for (uint32_t i = threadIdx.x * 4; i < bl; i += TDG_THREADS * 4) {
        const int nv = min(4u, bl - i);
        float l[4], r[4];
        ld4<VEC>(T.h.out_l, p0 + i, nv, l);
        ld4<VEC>(T.h.out_r, p0 + i, nv, r);
#pragma unroll
        for (int j = 0; j < 4; ++j) { l[j] *= f; r[j] *= f; apply_pg(l[j], r[j], T.h.pg); }
        st4<VEC>(T.h.out_l, p0 + i, nv, l);
        st4<VEC>(T.h.out_r, p0 + i, nv, r);
    }
}

// ------------------------------------------------------------------------------------ K2: sample gathers
// A chunk is 1024 frames: 8 KB of output per block behind a chain of dependent loads (task -> [range -> hits ->]
// sample), and ncu shows these kernels waiting on exactly that chain (13-15 stall cycles of long_scoreboard per issue
// at 85 % occupancy, DRAM 40-47 % busy).  A block therefore walks TDG_GATHER_CPB consecutive chunks of its vertex: the
// task load, the barrier behind it and the block launch are paid once per group of chunks, and the range of the next
// chunk is fetched while the current one is evaluated.  Measured per bench step (loop / multi / lerp): 1 chunk per block
// 1.77 / 2.44 / 2.53 ms, 4: 1.52 / 1.70 / 1.60, 8: 1.51 / 1.62 / 1.44.
#ifndef TDG_GATHER_CPB
#define TDG_GATHER_CPB 8
#endif
// sample_loop_gen (extensions.rs:331-341): buf[i] = sample[(t + i) % L]
template <bool VEC>
__global__ void __launch_bounds__(TDG_THREADS) k_sample_loop(const LoopTask* __restrict__ tasks, uint32_t cpb) {
    __shared__ LoopTask T;
    load_task_pdl(&T, tasks + blockIdx.y);
    const uint32_t kc0 = blockIdx.x * cpb;
    if (kc0 >= T.h.n_chunks) return;
    const uint32_t kc1 = min(kc0 + cpb, T.h.n_chunks);
    const uint32_t bl = T.h.bl;
    const uint64_t len = T.len;
#pragma unroll 1
    for (uint32_t k = kc0; k < kc1; ++k) {
    const uint64_t p0 = (uint64_t)k * bl;
    for (uint32_t i = threadIdx.x * 4; i < bl; i += TDG_THREADS * 4) {
        const int nv = min(4u, bl - i);
        // (a per-chunk phase table from the host was measured slower than this modulo: the extra dependent
        // load at the head of every block costs more than the arithmetic it saves)
        const uint64_t at = T.t0 + p0 + i;
        // 64-bit division is a ~60-instruction subroutine on the SM; a timeline shorter than 2^32 frames (24 h at
        // 48 kHz) and a sample shorter than 2^32 frames take the 32-bit form
        uint64_t s = ((at | len) >> 32) == 0 ? (uint64_t)((uint32_t)at % (uint32_t)len) : at % len;
        float l[4], r[4];
        if (s + 3 < len) {                                        // no wrap inside the run
            const float* __restrict__ pl = T.sl + s;
            const float* __restrict__ pr = T.sr + s;
#pragma unroll
            for (int j = 0; j < 4; ++j) { l[j] = __ldg(pl + j); r[j] = __ldg(pr + j); apply_pg(l[j], r[j], T.h.pg); }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                l[j] = __ldg(T.sl + s); r[j] = __ldg(T.sr + s);
                apply_pg(l[j], r[j], T.h.pg);
                if (++s == len) s = 0;
            }
        }
        st4<VEC>(T.h.out_l, p0 + i, nv, l);
        st4<VEC>(T.h.out_r, p0 + i, nv, r);
    }
    }
}

// sample_multi_gen (extensions.rs:344-381): every live hit adds sample[pos]*vel in deque order.
// All hits share one sample length, so the live set at frame p is the contiguous range of hits with
// start in (p - L, p]; hits are sorted by start (push order).
template <bool VEC>
__global__ void __launch_bounds__(TDG_THREADS) k_sample_multi(const MultiTask* __restrict__ tasks, const uint8_t* __restrict__ blob, uint32_t cpb) {
    __shared__ MultiTask T;
    load_task_pdl(&T, tasks + blockIdx.y);
    const uint32_t kc0 = blockIdx.x * cpb;
    if (kc0 >= T.h.n_chunks) return;
    const uint32_t bl = T.h.bl;
    const MultiHit* hits = reinterpret_cast<const MultiHit*>(blob + T.hits_off);
    // candidates for a chunk: start > p0 - L  and  start <= p0 + bl - 1 (k_chunk_ranges)
    const uint32_t kc1 = min(kc0 + cpb, T.h.n_chunks);
    Range2 rg_next = T.ranges[kc0];
    const int64_t len = (int64_t)T.len;
#pragma unroll 1
    for (uint32_t k = kc0; k < kc1; ++k) {
    const Range2 rg = rg_next;
    if (k + 1 < kc1) rg_next = T.ranges[k + 1];              // in flight while this chunk is evaluated
    const int64_t p0 = (int64_t)k * bl;
    const uint32_t lo = rg.x, hi = rg.y;
    for (uint32_t i = threadIdx.x * 4; i < bl; i += TDG_THREADS * 4) {
        const int nv = min(4u, bl - i);
        float l[4] = {0.0f, 0.0f, 0.0f, 0.0f}, r[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        for (uint32_t h = lo; h < hi; ++h) {
            const int64_t base = p0 + i - hits[h].start;          // sample position of the thread's first frame
            const float vel = hits[h].vel;
            if (base >= 0 && base + 3 < len) {                    // the hit covers all four frames: no per-frame tests
                const float* __restrict__ pl = T.sl + base;
                const float* __restrict__ pr = T.sr + base;
#pragma unroll
                for (int j = 0; j < 4; ++j) { l[j] += __ldg(pl + j) * vel; r[j] += __ldg(pr + j) * vel; }
            } else if (base + 3 >= 0 && base < len) {             // starts or ends inside the run
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int64_t pos = base + j;
                    if (pos >= 0 && pos < len) {
                        l[j] += __ldg(T.sl + pos) * vel;
                        r[j] += __ldg(T.sr + pos) * vel;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) apply_pg(l[j], r[j], T.h.pg);
        st4<VEC>(T.h.out_l, (uint64_t)p0 + i, nv, l);
        st4<VEC>(T.h.out_r, (uint64_t)p0 + i, nv, r);
    }
    }
}

// sample_lerp_gen (extensions.rs:384-421): primary = last hit at or before the frame, ghost = the
// one before it, cross-faded while the countdown (reset by each hit) runs.
template <bool VEC>
__global__ void __launch_bounds__(TDG_THREADS) k_sample_lerp(const LerpTask* __restrict__ tasks, const uint8_t* __restrict__ blob, uint32_t cpb) {
    __shared__ LerpTask T;
    load_task_pdl(&T, tasks + blockIdx.y);
    const uint32_t kc0 = blockIdx.x * cpb;
    if (kc0 >= T.h.n_chunks) return;
    const uint32_t bl = T.h.bl;
    const LerpHit* hits = reinterpret_cast<const LerpHit*>(blob + T.hits_off);
    // index of the last hit with frame <= p0 (entries 0 and 1 are the carried ghost / primary); k_chunk_ranges
    const uint32_t kc1 = min(kc0 + cpb, T.h.n_chunks);
    uint32_t base_next = T.ranges[kc0].x;
    const int64_t lm1 = (int64_t)T.len - 1;
    const int64_t lerp_len = (int64_t)T.lerp_len;
#pragma unroll 1
    for (uint32_t k = kc0; k < kc1; ++k) {
    const uint32_t base = base_next;
    if (k + 1 < kc1) base_next = T.ranges[k + 1].x;          // in flight while this chunk is evaluated
    const int64_t p0 = (int64_t)k * bl;
    for (uint32_t i = threadIdx.x * 4; i < bl; i += TDG_THREADS * 4) {
        const int nv = min(4u, bl - i);
        float l[4], r[4];
        uint32_t idx = base;
        const int64_t pf = p0 + i;
        while (idx + 1 < T.n_hits && hits[idx + 1].frame <= pf) ++idx;
        const bool one = idx + 1 >= T.n_hits || hits[idx + 1].frame > pf + 3;     // one primary for the whole run
        const LerpHit P0 = hits[idx];
        const int64_t pp0 = pf - P0.pos_origin;
        const int64_t cb0 = lerp_len - (pf - P0.cd_origin);                       // countdown before frame 0 of the run
        if (one && cb0 <= 0 && (pp0 >= lm1 || (pp0 >= 0 && pp0 + 3 <= lm1))) {
            // the common run: no hit inside it, cross-fade over; the sample is either being played straight
            // through or has ended (clamped to its last frame)
            if (pp0 >= lm1) {
                const float lv = __ldg(T.sl + lm1) * P0.vel, rv = __ldg(T.sr + lm1) * P0.vel;
#pragma unroll
                for (int j = 0; j < 4; ++j) { l[j] = lv; r[j] = rv; }
            } else {
                const float* __restrict__ pl = T.sl + pp0;
                const float* __restrict__ pr = T.sr + pp0;
#pragma unroll
                for (int j = 0; j < 4; ++j) { l[j] = __ldg(pl + j) * P0.vel; r[j] = __ldg(pr + j) * P0.vel; }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) apply_pg(l[j], r[j], T.h.pg);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t p = pf + j;
                while (idx + 1 < T.n_hits && hits[idx + 1].frame <= p) ++idx;
                const LerpHit P = hits[idx];
                const int64_t ppos = min(max(p - P.pos_origin, (int64_t)0), lm1);
                float lv = __ldg(T.sl + ppos) * P.vel;
                float rv = __ldg(T.sr + ppos) * P.vel;
                const int64_t c_before = lerp_len - (p - P.cd_origin);
                if (c_before > 0) {
                    const LerpHit G = hits[idx - 1];
                    const float t = (float)(c_before - 1) / (float)T.lerp_len;
                    const int64_t gpos = min(max(p - G.pos_origin, (int64_t)0), lm1);
                    const float gl = __ldg(T.sl + gpos) * G.vel;
                    const float gr = __ldg(T.sr + gpos) * G.vel;
                    lv = gl * t + lv * (1.0f - t);
                    rv = gr * t + rv * (1.0f - t);
                }
                apply_pg(lv, rv, T.h.pg);
                l[j] = lv; r[j] = rv;
            }
        }
        st4<VEC>(T.h.out_l, (uint64_t)p0 + i, nv, l);
        st4<VEC>(T.h.out_r, (uint64_t)p0 + i, nv, r);
    }
    }
}

// ------------------------------------------------------------------------------------ K5: voices
// debug_sine_gen / synth_gen / sampsyn_gen (extensions.rs:423-578).  The host planner turns the
// floww events into per-chunk voice segments in push order; every frame sums its live voices
// sequentially in that order (f32 summation order is part of the contract).
//
// These kernels are FP32-issue bound, not HBM bound (8 B of output per frame against ~60-90
// instructions per voice-frame), so the design minimises instructions per voice-frame:
//   * a thread owns 4 consecutive frames and walks the chunk's segments once (segment-major), so the
//     segment fetch, the held/released decision and every per-voice constant are paid once per 4
//     frames and the four dependent chains overlap;
//   * every loop-invariant f32 value (envelope thresholds / deltas / reciprocals, shape constants,
//     1/sr) is computed once on the host with the reference's operations (AdsrK, VoiceTask);
//   * the release start value apply_ads(conf, rel_t) of adsr.rs:89-92 is constant per segment and is
//     evaluated once per (segment, oscillator) while the segments are staged in shared memory;
//   * divisions by those constants use the exact-quotient FMA form (k_div) instead of MUFU.RCP +
//     Newton + slow path; sin is glibc's sinf restated (m_sinf_glibc, bit-exact) or, on request, a Cody-Waite +
//     minimax evaluation on the FMA pipe (d_sin);
//   * a block walks up to TDG_VOICE_CPB consecutive chunks (task load, barrier and block launch once per group).
// Accumulators stay in registers across segment batches; one float4 store per channel at the end.
#ifndef TDG_SYNTH_F
#define TDG_SYNTH_F 4           // frames per thread of the synth kernel (tuning knob, see profiles/)
#endif
#ifndef TDG_SYNTH_MINB
#define TDG_SYNTH_MINB 5        // min resident blocks per SM (register cap); measured per 256-project step: 2: 9.9, 3: 9.9, 4: 8.2, 5: 7.7, 6: 7.7 ms
#endif
#ifndef TDG_VOICE_CPB
#define TDG_VOICE_CPB 8         // chunks per block of the voice kernels: measured (synth / sampsyn) 1 -> 9.34 / 9.45, 2 -> 9.09 / 8.65, 4 -> 8.87 / 8.17,
                                // 8 -> 8.76 / 7.91, 16 -> 8.72 / 7.82 ms per bench step
#endif
#ifndef TDG_VOICE_HOIST
#define TDG_VOICE_HOIST 1
#endif
#ifndef TDG_SAMPSYN_MINB
#define TDG_SAMPSYN_MINB 5      // 4: 5.77, 5: 5.58, 6: 5.63 ms
#endif

// wavetable_act_state as DEFINED by the oracle (sampsyn is not vendored): bilinear in (time, phase).
// Device layout of the table (tdg_add_sampsyn): pairs[j][i] = (tab[j][i], tab[min(j + 1, last)][i]) for
// i in [0, n_phase], column n_phase repeating column 0.  The four corners of a cell are then two 8-byte loads at
// consecutive addresses from ONE computed address (the phase wrap and the last-row clamp are in the layout), where
// the row-major table needed four address computations, a wrap select and a row clamp per voice-frame.
// floor / truncation use the conversion pipe (FRND / F2I / I2F); the table is read through the read-only path and
// stays L1 resident.  Same values, same f32 operations in the same order as the oracle's definition.
__device__ __forceinline__ float wavetable_eval(const float2* __restrict__ row, float n_phase_f, int n_phase, float fy, float hz, float t) {
    float ph = t * hz;
    ph = ph - floorf(ph);
    const float x = ph * n_phase_f;
    int i0 = __float2int_rz(x);
    if (i0 >= n_phase) i0 = n_phase - 1;
    const float fx = x - (float)i0;
    const float2 c0 = __ldg(row + i0), c1 = __ldg(row + i0 + 1);          // (a0, b0), (a1, b1)
    const float a = c0.x + fx * (c1.x - c0.x);
    const float b = c0.y + fx * (c1.y - c0.y);
    return a + fy * (b - a);
}
// the time axis of the table for one frame: row j0 = trunc(clamp(t * fps, 0, ymax)) and the weight towards row j0 + 1
__device__ __forceinline__ int wavetable_row(float fps, float ymax, float t, float* fy) {
    float y = t * fps;
    if (!(y > 0.0f)) y = 0.0f;
    if (y > ymax) y = ymax;
    const int j0 = __float2int_rz(y);
    *fy = y - (float)j0;
    return j0;
}

// F = consecutive frames per thread (register pressure: the three-oscillator synth body runs 2-wide)
//
// No block-wide staging and no barrier after the task load: every thread walks the chunk's segments
// straight from the device-built segment table (k_seg_fill: 32 B per segment, the same address for the
// whole block, so each read is one L1-resident broadcast); the per-voice release start values come with it.
// The barrier-per-batch version spent a quarter of its stall cycles waiting for the two warps that
// staged and pre-evaluated the segments (profiles/: stall_barrier).
// LIBM: sin is glibc's sinf restated in double precision (m_sinf_glibc: bit-exact against the reference's libm
// call); false = the f32 Cody-Waite + minimax evaluation d_sin (<= 2 ulp, tdg_set_oscillator_mode(ctx, 0)).
template <int KIND, bool EXACT, int F, bool LIBM>
__global__ void __launch_bounds__(TDG_THREADS, KIND == VOICE_SYNTH ? TDG_SYNTH_MINB : TDG_SAMPSYN_MINB) k_voice(const VoiceTask* __restrict__ tasks, const uint8_t* __restrict__ blob, uint32_t cpb) {
    __shared__ VoiceTask T;
    load_task_pdl(&T, tasks + blockIdx.y);
    const uint32_t kc0 = blockIdx.x * cpb;
    if (kc0 >= T.h.n_chunks) return;
    const uint32_t kc1 = min(kc0 + cpb, T.h.n_chunks);
    const uint32_t bl = T.h.bl;
    const float4* __restrict__ gsegs = reinterpret_cast<const float4*>(T.segs);
    const float* __restrict__ grelv2 = T.relv2;
    const float2* __restrict__ tab = reinterpret_cast<const float2*>(T.table);
    const float fsr = T.fsr, rsr = T.rsr;
    const bool vec = (bl % F) == 0;
    const float n_phase_f = (float)T.n_phase, ymax = (float)(T.n_frames - 1);
    const int row_stride = (int)T.n_phase + 1;
    uint32_t s_next = T.seg_ptr[kc0];
    // i as f32 / sr as f32 does not depend on the chunk: with one run per thread and chunk (bl <= 1024) it is computed
    // once per block instead of once per chunk
    const bool one_run = TDG_VOICE_HOIST && bl <= (uint32_t)(TDG_THREADS * F);
    float off[F];
    if (one_run) {
#pragma unroll
        for (int j = 0; j < F; ++j) off[j] = k_div<EXACT>((float)(threadIdx.x * F + j), fsr, rsr);
    }
#pragma unroll 1
    for (uint32_t k = kc0; k < kc1; ++k) {
    const uint64_t p0 = (uint64_t)k * bl;
    const uint32_t s_lo = s_next, s_hi = T.seg_ptr[k + 1];
    s_next = s_hi;
    for (uint32_t base = 0; base < bl; base += TDG_THREADS * F) {
        const uint32_t i0 = base + threadIdx.x * F;
        if (i0 >= bl) break;
        float acc[F], time[F];
#pragma unroll
        for (int j = 0; j < F; ++j) {
            acc[j] = 0.0f;
            time[j] = k_div<EXACT>((float)(T.t0 + p0 + i0 + j), fsr, rsr);      // (t + i) as f32 / sr as f32
            if (!one_run) off[j] = k_div<EXACT>((float)(i0 + j), fsr, rsr);     // i as f32 / sr as f32
        }
#pragma unroll 1
        for (uint32_t s = s_lo; s < s_hi; ++s) {
            const float4 sb = __ldg(gsegs + 2 * (size_t)s + 1);                   // ia, ib, relv0, relv1
            const uint32_t ia = __float_as_uint(sb.x), ib = __float_as_uint(sb.y);
            if (i0 + F <= ia || i0 >= ib) continue;
            const float4 sa = __ldg(gsegs + 2 * (size_t)s);                       // hz, vel, env_t, rel_t
            const float hz = sa.x, vel = sa.y, env_t = sa.z, rel_t = sa.w;
            const bool rel = rel_t != 0.0f;
            float smp[F];
            if (KIND == VOICE_DEBUG_SINE) {                                       // extensions.rs:450-454
#pragma unroll
                for (int j = 0; j < F; ++j) smp[j] = (LIBM ? m_sinf_glibc(m_osc_arg(time[j], hz)) : d_sin(m_osc_arg(time[j], hz))) * vel;
            } else {
                float et[F], e[F];
#pragma unroll
                for (int j = 0; j < F; ++j) et[j] = env_t + off[j];               // env_time
                if (KIND == VOICE_SYNTH) {                                        // extensions.rs:500-522
                    float sn[F];
                    if (T.vol[0] > 0.0f || T.vol[1] > 0.0f) {
#pragma unroll
                        for (int j = 0; j < F; ++j) sn[j] = LIBM ? m_sinf_glibc(m_osc_arg(time[j], hz)) : d_sin(m_osc_arg(time[j], hz));
                    }
#pragma unroll
                    for (int j = 0; j < F; ++j) smp[j] = 0.0f;
                    bool flat = false;
#pragma unroll 1
                    for (int o = 0; o < 3; ++o) {                                 // square, topflat, triangle in this order
                        const float vol = T.vol[o];
                        if (!(vol > 0.0f)) continue;
                        // topflat and triangle sharing one ADSR table (project.lua does): same env time, same
                        // release start value => the values just computed for oscillator 1 are oscillator 2's
                        const bool reuse = o == 2 && T.env12_same && T.vol[1] > 0.0f;
                        if (!reuse)
                            flat = rel ? k_apply_r_n<EXACT, F>(T.env[o], et, o == 0 ? sb.z : o == 1 ? sb.w : __ldg(grelv2 + s), e)
                                       : k_apply_ads_n<EXACT, F>(T.env[o], et, e);
                        // envelope sitting at exactly 0 (a percussive oscillator after its decay): every term is
                        // shape*vel*0*vol = +-0 for finite factors, and adding +-0 to the running sum changes nothing
                        if (flat && e[0] == 0.0f && fabsf(vel) < __int_as_float(0x7f800000) && vol < __int_as_float(0x7f800000)) continue;
                        if (o == 0) {
#pragma unroll
                            for (int j = 0; j < F; ++j) smp[j] += fminf(fmaxf(sn[j], -T.sq_z), T.sq_z) * T.sq_iz * vel * e[j] * vol;
                        } else if (o == 1) {
#pragma unroll
                            for (int j = 0; j < F; ++j) smp[j] += (fminf(sn[j], T.tf_z) + T.tf_add) * T.tf_mul * vel * e[j] * vol;
                        } else {
#pragma unroll
                            for (int j = 0; j < F; ++j) smp[j] += m_triangle(time[j], hz) * vel * e[j] * vol;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < F; ++j) smp[j] *= T.amp_mul;
                } else {                                                          // extensions.rs:560-571
                    if (rel) k_apply_r_n<EXACT, F>(T.env[0], et, sb.z, e); else k_apply_ads_n<EXACT, F>(T.env[0], et, e);
                    // (a shortcut for voices whose table time has run past the last row — j0 = last, fy = +0, b == a — was
                    // measured SLOWER: configs[2] 58.1 -> 61.6 ms, the second inlined body costs more than the ten
                    // instructions it saves)
#pragma unroll
                    for (int j = 0; j < F; ++j) {
                        float fy;
                        const int jr = wavetable_row(T.fps, ymax, et[j] + rel_t, &fy);
                        const float g = vel * e[j] * T.amp_mul;
                        smp[j] = wavetable_eval(tab + (size_t)jr * row_stride, n_phase_f, (int)T.n_phase, fy, hz, et[j] + rel_t) * g;
                    }
                }
            }
            if (ia <= i0 && i0 + F <= ib) {                                    // the segment covers the whole run (nearly always)
#pragma unroll
                for (int j = 0; j < F; ++j) acc[j] += smp[j];
            } else {
#pragma unroll
                for (int j = 0; j < F; ++j)
                    if (i0 + j >= ia && i0 + j < ib) acc[j] += smp[j];
            }
        }
        {
            float l[F], r[F];
#pragma unroll
            for (int j = 0; j < F; ++j) { l[j] = acc[j]; r[j] = acc[j]; apply_pg(l[j], r[j], T.h.pg); }
            float* ol = T.h.out_l + p0 + i0; float* orr = T.h.out_r + p0 + i0;
            if (vec && F == 4) {
                *reinterpret_cast<float4*>(ol) = make_float4(l[0], l[1], l[2], l[3]);
                *reinterpret_cast<float4*>(orr) = make_float4(r[0], r[1], r[2], r[3]);
            } else if (vec && F == 2) {
                *reinterpret_cast<float2*>(ol) = make_float2(l[0], l[1]);
                *reinterpret_cast<float2*>(orr) = make_float2(r[0], r[1]);
            } else {
#pragma unroll
                for (int j = 0; j < F; ++j) if (i0 + j < bl) { ol[j] = l[j]; orr[j] = r[j]; }
            }
        }
    }
    }
}

// ------------------------------------------------------------------------------------ K7: band pass scan
// y <- y + g*(x - y) is the affine map y -> (1-g)*y + g*x.  A thread owns a run of consecutive frames (8 in the
// three-launch scan, 32 in k_bp_fused): it folds them from a zero state (B) with decay A = (1-g)^n, a warp-shuffle
// scan composes the maps across the block, block aggregates are scanned across blocks (k_bp_carry, or the look-back
// of k_bp_fused), and the final step re-runs the reference's exact operation sequence from the carried-in state.
// Only the carried-in state is reassociated (tolerance in DESIGN.md); inside a run the arithmetic is the reference's.
// k_bp_fused (below) is what a pass uses; k_bp_aggregate / k_bp_carry / k_bp_apply remain for time-sharded passes,
// where the exchange step sits between aggregate and apply, and behind TDG_BP_FUSED=0.
struct Aff { float al, ah, b0, b1, b2, b3; };       // (lprevl, lprevr) decay al; (hprevl, hprevr) decay ah
__device__ __forceinline__ Aff aff_id() { return Aff{1.0f, 1.0f, 0.0f, 0.0f, 0.0f, 0.0f}; }
__device__ __forceinline__ Aff aff_then(const Aff& f, const Aff& g) {   // apply f first, then g
    return Aff{f.al * g.al, f.ah * g.ah, fmaf(g.al, f.b0, g.b0), fmaf(g.al, f.b1, g.b1),
               fmaf(g.ah, f.b2, g.b2), fmaf(g.ah, f.b3, g.b3)};
}
__device__ __forceinline__ Aff aff_shfl_up(const Aff& a, int d) {
    return Aff{__shfl_up_sync(0xffffffffu, a.al, d), __shfl_up_sync(0xffffffffu, a.ah, d),
               __shfl_up_sync(0xffffffffu, a.b0, d), __shfl_up_sync(0xffffffffu, a.b1, d),
               __shfl_up_sync(0xffffffffu, a.b2, d), __shfl_up_sync(0xffffffffu, a.b3, d)};
}
// exclusive prefix of per-thread maps over the block (thread order), plus the block total
__device__ __forceinline__ Aff block_aff_scan(const Aff& mine, Aff* total) {
    static_assert(kBpThreads <= 256 && kBpFusedThreads <= 256, "wtot holds one map per warp");
    __shared__ Aff wtot[8];
    const int ln = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    Aff inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        Aff up = aff_shfl_up(inc, d);
        if (ln >= d) inc = aff_then(up, inc);
    }
    if (ln == 31) wtot[w] = inc;
    __syncthreads();
    if (w == 0) {
        Aff t = ln < nw ? wtot[ln] : aff_id();
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            Aff up = aff_shfl_up(t, d);
            if (ln >= d) t = aff_then(up, t);
        }
        if (ln < nw) wtot[ln] = t;     // inclusive over warps
    }
    __syncthreads();
    Aff exc = aff_shfl_up(inc, 1);
    if (ln == 0) exc = aff_id();
    if (w > 0) exc = aff_then(wtot[w - 1], exc);
    if (total) *total = wtot[nw - 1];
    __syncthreads();
    return exc;
}

template <bool VEC>
__device__ __forceinline__ void bp_load_run(const BpTask& T, const float* const* ins, uint64_t p, int nv,
                                            float (&xl)[8], float (&xr)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { xl[j] = 0.0f; xr[j] = 0.0f; }
    for (uint32_t a = 0; a < T.n_in; ++a) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float x[4], y[4];
            const int nvh = max(0, min(4, nv - 4 * h));
            if (VEC || nvh > 0) {
                ld4<VEC>(ins[2 * a], p + 4 * h, nvh, x);
                ld4<VEC>(ins[2 * a + 1], p + 4 * h, nvh, y);
#pragma unroll
                for (int j = 0; j < 4; ++j) { xl[4 * h + j] += x[j]; xr[4 * h + j] += y[j]; }
            }
        }
    }
}
__device__ __forceinline__ Aff bp_fold_run(const BpTask& T, const float (&xl)[8], const float (&xr)[8], int nv) {
    float ll = 0.0f, lr = 0.0f, hl = 0.0f, hr = 0.0f, al = 1.0f, ah = 1.0f;
    const float dl = 1.0f - T.lgamma, dh = 1.0f - T.hgamma;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < nv) {
            ll = ll + T.lgamma * (xl[j] - ll); lr = lr + T.lgamma * (xr[j] - lr);
            hl = hl + T.hgamma * (xl[j] - hl); hr = hr + T.hgamma * (xr[j] - hr);
            al *= dl; ah *= dh;
        }
    }
    return Aff{al, ah, ll, lr, hl, hr};
}

// pass A: per-tile aggregate (end state from zero, decay)
template <bool VEC>
__global__ void __launch_bounds__(kBpThreads) k_bp_aggregate(const BpTask* __restrict__ tasks, const uint8_t* __restrict__ blob) {
    __shared__ BpTask T;
    load_task(&T, tasks + blockIdx.y);
    if (blockIdx.x >= T.n_tiles) return;
    const float* const* ins = T.n_in <= (uint32_t)kInlineInputs ? T.ins : reinterpret_cast<const float* const*>(blob + T.in_off);
    const uint64_t p = (uint64_t)blockIdx.x * kBpTile + (uint64_t)threadIdx.x * 8;
    const int nv = p >= T.h.n_frames ? 0 : (int)min((uint64_t)8, T.h.n_frames - p);
    float xl[8], xr[8];
    bp_load_run<VEC>(T, ins, p, nv, xl, xr);
    Aff mine = bp_fold_run(T, xl, xr, nv), total;
    block_aff_scan(mine, &total);
    if (threadIdx.x == 0) {
        float* a = T.agg + (size_t)blockIdx.x * 8;
        a[0] = total.b0; a[1] = total.b1; a[2] = total.b2; a[3] = total.b3; a[4] = total.al; a[5] = total.ah;
    }
}
// pass B: scan of tile aggregates -> state entering every tile; one block per filter vertex
__device__ __forceinline__ void bp_init_state(const BpTask& T, const uint8_t* blob, float (&init)[4]) {
    if (T.first) {                                              // extensions.rs:664-670
        const float* const* ins = reinterpret_cast<const float* const*>(blob + T.in_off);
        float l = 0.0f, r = 0.0f;
        for (uint32_t a = 0; a < T.n_in; ++a) { l += ins[2 * a][0]; r += ins[2 * a + 1][0]; }
        init[0] = l; init[1] = r; init[2] = l; init[3] = r;
    } else {
        init[0] = T.state[0]; init[1] = T.state[1]; init[2] = T.state[2]; init[3] = T.state[3];
    }
}
__device__ __forceinline__ Aff bp_tiles_fold(const BpTask& T, uint32_t lo, uint32_t hi) {
    Aff mine = aff_id();
    for (uint32_t t = lo; t < hi; ++t) {
        const float* a = T.agg + (size_t)t * 8;
        mine = aff_then(mine, Aff{a[4], a[5], a[0], a[1], a[2], a[3]});
    }
    return mine;
}
// time sharding: the segment's whole affine map (and, on rank 0, the pass's initial state) into the slot
__global__ void __launch_bounds__(kBpThreads) k_bp_total(const BpTask* __restrict__ tasks, const uint8_t* __restrict__ blob) {
    __shared__ BpTask T;
    load_task(&T, tasks + blockIdx.x);
    const uint32_t n = T.n_tiles, per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
    Aff total;
    block_aff_scan(bp_tiles_fold(T, lo, hi), &total);
    if (threadIdx.x == 0) {
        float* o = T.x.buf + (size_t)T.x.rank * T.x.stride + T.x.index;
        o[0] = total.al; o[1] = total.ah; o[2] = total.b0; o[3] = total.b1; o[4] = total.b2; o[5] = total.b3;
        float init[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (T.x.rank == 0) bp_init_state(T, blob, init);
        o[6] = init[0]; o[7] = init[1]; o[8] = init[2]; o[9] = init[3];
    }
}
__global__ void __launch_bounds__(kBpThreads) k_bp_carry(const BpTask* __restrict__ tasks, const uint8_t* __restrict__ blob) {
    __shared__ BpTask T;
    __shared__ float init[4];
    load_task(&T, tasks + blockIdx.x);
    if (threadIdx.x == 0) {
        float st[4];
        if (T.x.ranks > 1) {
            // state entering this segment: rank 0's initial state pushed through the maps of all earlier
            // segments; the end-of-pass state (all maps) is what every rank keeps for the next pass
            const float* g0 = T.x.buf + T.x.index;
            st[0] = g0[6]; st[1] = g0[7]; st[2] = g0[8]; st[3] = g0[9];
            float mine[4] = {st[0], st[1], st[2], st[3]};
            for (uint32_t q = 0; q < T.x.ranks; ++q) {
                if (q == T.x.rank) { mine[0] = st[0]; mine[1] = st[1]; mine[2] = st[2]; mine[3] = st[3]; }
                const float* g = T.x.buf + (size_t)q * T.x.stride + T.x.index;
                st[0] = fmaf(g[0], st[0], g[2]); st[1] = fmaf(g[0], st[1], g[3]);
                st[2] = fmaf(g[1], st[2], g[4]); st[3] = fmaf(g[1], st[3], g[5]);
            }
            T.state[0] = st[0]; T.state[1] = st[1]; T.state[2] = st[2]; T.state[3] = st[3];
            init[0] = mine[0]; init[1] = mine[1]; init[2] = mine[2]; init[3] = mine[3];
        } else {
            bp_init_state(T, blob, st);
            init[0] = st[0]; init[1] = st[1]; init[2] = st[2]; init[3] = st[3];
        }
    }
    __syncthreads();
    const uint32_t n = T.n_tiles, per = (n + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(n, threadIdx.x * per), hi = min(n, lo + per);
    Aff exc = block_aff_scan(bp_tiles_fold(T, lo, hi), nullptr);
    float c0 = fmaf(exc.al, init[0], exc.b0), c1 = fmaf(exc.al, init[1], exc.b1);
    float c2 = fmaf(exc.ah, init[2], exc.b2), c3 = fmaf(exc.ah, init[3], exc.b3);
    for (uint32_t t = lo; t < hi; ++t) {
        float* c = T.carry + (size_t)t * 4;
        c[0] = c0; c[1] = c1; c[2] = c2; c[3] = c3;
        const float* a = T.agg + (size_t)t * 8;
        c0 = fmaf(a[4], c0, a[0]); c1 = fmaf(a[4], c1, a[1]); c2 = fmaf(a[5], c2, a[2]); c3 = fmaf(a[5], c3, a[3]);
    }
}
// pass C: exact recurrence from the carried-in state + cut/pass mix + pan + gain
template <bool VEC>
__global__ void __launch_bounds__(kBpThreads, TDG_BP_MINB) k_bp_apply(const BpTask* __restrict__ tasks, const uint8_t* __restrict__ blob) {
    __shared__ BpTask T;
    load_task(&T, tasks + blockIdx.y);
    if (blockIdx.x >= T.n_tiles) return;
    const float* const* ins = T.n_in <= (uint32_t)kInlineInputs ? T.ins : reinterpret_cast<const float* const*>(blob + T.in_off);
    const uint64_t p = (uint64_t)blockIdx.x * kBpTile + (uint64_t)threadIdx.x * 8;
    const int nv = p >= T.h.n_frames ? 0 : (int)min((uint64_t)8, T.h.n_frames - p);
    float xl[8], xr[8];
    bp_load_run<VEC>(T, ins, p, nv, xl, xr);
    const Aff mine = bp_fold_run(T, xl, xr, nv);
    const Aff exc = block_aff_scan(mine, nullptr);
    const float* c = T.carry + (size_t)blockIdx.x * 4;
    float ll = fmaf(exc.al, c[0], exc.b0), lr = fmaf(exc.al, c[1], exc.b1);
    float hl = fmaf(exc.ah, c[2], exc.b2), hr = fmaf(exc.ah, c[3], exc.b3);
    const float lmul = T.lgamma == 0.0f ? 0.0f : 1.0f, hmul = T.hgamma == 0.0f ? 0.0f : 1.0f;
    const float pass_mul = T.pass ? 1.0f : 0.0f, cut_mul = 1.0f - pass_mul;
    float ol[8], orr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < nv) {                                           // extensions.rs:671-688
            const float l = xl[j], r = xr[j];
            ll = ll + T.lgamma * (l - ll); lr = lr + T.lgamma * (r - lr);
            hl = hl + T.hgamma * (l - hl); hr = hr + T.hgamma * (r - hr);
            const float cutl = (lmul * ll + hmul * (l - hl)) * 0.5f;
            const float cutr = (lmul * lr + hmul * (r - hr)) * 0.5f;
            const float passl = l - cutl;
            const float passr = r - cutl;                       // sic (extensions.rs:685)
            float a = cutl * cut_mul + passl * pass_mul;
            float b = cutr * cut_mul + passr * pass_mul;
            apply_pg(a, b, T.h.pg);
            ol[j] = a; orr[j] = b;
        } else { ol[j] = 0.0f; orr[j] = 0.0f; }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int nvh = max(0, min(4, nv - 4 * h));
        if (nvh == 0) continue;
        float a[4] = {ol[4 * h], ol[4 * h + 1], ol[4 * h + 2], ol[4 * h + 3]};
        float b[4] = {orr[4 * h], orr[4 * h + 1], orr[4 * h + 2], orr[4 * h + 3]};
        if (VEC && nvh == 4) { st4<true>(T.h.out_l, p + 4 * h, 4, a); st4<true>(T.h.out_r, p + 4 * h, 4, b); }
        else { st4<false>(T.h.out_l, p + 4 * h, nvh, a); st4<false>(T.h.out_r, p + 4 * h, nvh, b); }
    }
    // the run holding the last frame of the pass leaves the state for the next pass
    if (T.x.ranks <= 1 && nv > 0 && p + nv == T.h.n_frames) { T.state[0] = ll; T.state[1] = lr; T.state[2] = hl; T.state[3] = hr; }
}

// Single pass: aggregate, look back, apply in ONE kernel — every input is read from HBM once (16 + 8*(indeg-1) bytes
// per frame where the three-launch scan moves 24 + 16*(indeg-1)).
//
// A block owns a STRIP of kBpStrip = 8192 consecutive frames, 32 per thread:
//   1. coalesced float4 loads of every input, summed in edge order, into shared memory (rows of 32 frames padded to 36
//      words: the float4 stores of this phase and the float4 row reads of the next are both bank-conflict free);
//   2. every thread folds ITS 32 consecutive frames from a zero state (the reference's operations), ONE block scan
//      composes the 256 maps, thread 0 publishes the strip's map;
//   3. warp 0 finds the state entering the strip: (a) the state leaving the previous GROUP of kBpGroup strips,
//      published by that group's last strip, pushed through (b) the maps of the strips of its own group in front of
//      it.  The composition is a fixed tree over the warp's lanes — which map meets which is decided by strip numbers
//      alone, never by timing — so the result is the same bits on every run (a free-running decoupled look-back would
//      not be).  Strips are handed out by a per-filter ticket: everything a strip waits for belongs to a block that
//      is already running;
//   4. every thread re-runs the reference's exact recurrence over its 32 frames from the carried-in state, writes the
//      outputs back into its shared-memory row, and the block stores them with coalesced float4 writes.
// Against the first version (2048-frame tiles, 8 frames per thread, inputs re-read from registers): one scan, one
// ticket, one publication and one look-back per 8192 frames instead of four, and the serial cross-group chain is four
// times shorter (n_strips / 32 steps; hence kBpFusedMaxTiles).
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
// 16 bytes global -> shared without passing through registers (LDGSTS; L2 only: the inputs are streamed once)
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
#ifdef __NVCC__
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gsrc) : "memory");
#else
    memcpy(smem_dst, gsrc, 16);           // tests/emu
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
#ifdef __NVCC__
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}
template <bool VEC>
__global__ void __launch_bounds__(kBpFusedThreads, 768 / kBpFusedThreads) k_bp_fused(const BpTask* __restrict__ tasks, const uint8_t* __restrict__ blob) {
#ifdef __NVCC__
    extern __shared__ float4 s_dyn4[];
    float* const s_l = reinterpret_cast<float*>(s_dyn4);
#else
    alignas(16) static float s_emu[2 * kBpFusedThreads * kBpRow];     // tests/emu: one block in flight at a time
    float* const s_l = s_emu;
#endif
    float* const s_r = s_l + kBpFusedThreads * kBpRow;
    __shared__ BpTask T;
    __shared__ uint32_t s_strip;
    __shared__ float s_carry[4];
    // grid = (filters, strips): consecutive blocks belong to DIFFERENT filters, so the blocks resident at any time
    // advance many filters' group chains at once (with the strips of one filter side by side the whole GPU waits
    // on that filter's serial chain)
    load_task_pdl(&T, tasks + blockIdx.x);
    if (blockIdx.y >= (T.n_tiles + kBpStripTiles - 1) / kBpStripTiles) return;
    if (threadIdx.x == 0) s_strip = atomicAdd(T.sync, 1u);
    __syncthreads();
    const uint32_t strip = s_strip;
    const float* const* ins = T.n_in <= (uint32_t)kInlineInputs ? T.ins : reinterpret_cast<const float* const*>(blob + T.in_off);
    const uint64_t pb = (uint64_t)strip * kBpStrip;
    const uint32_t n_here = (uint32_t)min((uint64_t)kBpStrip, T.h.n_frames - pb);                    // frames of the pass in this strip
    const uint32_t n_cov = (uint32_t)min((uint64_t)kBpStrip, (uint64_t)T.n_tiles * kBpTile - pb);    // frames the (padded) buffers cover
    // ---- 1. inputs -> shared memory (float4 q of the strip lands in row q / 8, words 4 * (q % 8) ..).
    // The first input goes straight from global to shared memory (cp.async, 16 bytes per request, no registers): the
    // whole strip of a block — 64 KB — is in flight at once, which is what keeps HBM busy with three blocks per SM
    // (with register loads in two halves of 32 KB the kernel sat at 0.6 of the HBM peak, latency bound).  Further
    // inputs are added in edge order through registers.  The reference's sum starts from +0.0, so a -0.0 never
    // survives it: the readers below add +0.0 to what they fetch, which is that same normalisation.
    if (VEC) {
#pragma unroll
        for (int i = 0; i < kBpFpt / 4; ++i) {
            const uint32_t q = threadIdx.x + kBpFusedThreads * i, f = 4 * q, w = (q >> 3) * kBpRow + 4 * (q & 7);
            if (f < n_cov) { cp_async16(s_l + w, ins[0] + pb + f); cp_async16(s_r + w, ins[1] + pb + f); }
            else {
                *reinterpret_cast<float4*>(s_l + w) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                *reinterpret_cast<float4*>(s_r + w) = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
        }
        bool landed = false;
        for (uint32_t a = 1; a < T.n_in; ++a) {
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                float4 x[4], y[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t f = 4u * (threadIdx.x + kBpFusedThreads * (4 * half + u));
                    if (f < n_cov) {
                        x[u] = *reinterpret_cast<const float4*>(ins[2 * a] + pb + f);
                        y[u] = *reinterpret_cast<const float4*>(ins[2 * a + 1] + pb + f);
                    } else { x[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); y[u] = x[u]; }
                }
                if (!landed) { cp_async_wait_all(); landed = true; }        // this thread's own copies: no barrier needed
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t q = threadIdx.x + kBpFusedThreads * (4 * half + u), w = (q >> 3) * kBpRow + 4 * (q & 7);
                    float4 l = *reinterpret_cast<const float4*>(s_l + w), r = *reinterpret_cast<const float4*>(s_r + w);
                    if (a == 1) { l.x += 0.0f; l.y += 0.0f; l.z += 0.0f; l.w += 0.0f; r.x += 0.0f; r.y += 0.0f; r.z += 0.0f; r.w += 0.0f; }
                    l.x += x[u].x; l.y += x[u].y; l.z += x[u].z; l.w += x[u].w;
                    r.x += y[u].x; r.y += y[u].y; r.z += y[u].z; r.w += y[u].w;
                    *reinterpret_cast<float4*>(s_l + w) = l;
                    *reinterpret_cast<float4*>(s_r + w) = r;
                }
            }
        }
        if (!landed) cp_async_wait_all();
    } else {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        float l[4][4], r[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int j = 0; j < 4; ++j) { l[u][j] = 0.0f; r[u][j] = 0.0f; }
        for (uint32_t a = 0; a < T.n_in; ++a) {
            float x[4][4], y[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t f = 4u * (threadIdx.x + kBpFusedThreads * (4 * half + u));
                const int nv = f >= n_here ? 0 : (int)min(4u, n_here - f);
                ld4<false>(ins[2 * a], pb + f, nv, x[u]); ld4<false>(ins[2 * a + 1], pb + f, nv, y[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) { l[u][j] += x[u][j]; r[u][j] += y[u][j]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t q = threadIdx.x + kBpFusedThreads * (4 * half + u), w = (q >> 3) * kBpRow + 4 * (q & 7);
            *reinterpret_cast<float4*>(s_l + w) = make_float4(l[u][0], l[u][1], l[u][2], l[u][3]);
            *reinterpret_cast<float4*>(s_r + w) = make_float4(r[u][0], r[u][1], r[u][2], r[u][3]);
        }
    }
    }
    __syncthreads();
    // ---- 2. the thread's 32 frames from a zero state
    const uint32_t f0 = threadIdx.x * kBpFpt;
    const int nv = f0 >= n_here ? 0 : (int)min((uint32_t)kBpFpt, n_here - f0);
    float* const row_l = s_l + threadIdx.x * kBpRow;
    float* const row_r = s_r + threadIdx.x * kBpRow;
    Aff mine;
    {
        float ll = 0.0f, lr = 0.0f, hl = 0.0f, hr = 0.0f, al = 1.0f, ah = 1.0f;
        const float dl = 1.0f - T.lgamma, dh = 1.0f - T.hgamma;
#pragma unroll 2
        for (int k = 0; k < kBpFpt / 4; ++k) {
            if (4 * k >= nv) break;
            const float4 a = *reinterpret_cast<const float4*>(row_l + 4 * k), b = *reinterpret_cast<const float4*>(row_r + 4 * k);
            const float xl[4] = {a.x + 0.0f, a.y + 0.0f, a.z + 0.0f, a.w + 0.0f}, xr[4] = {b.x + 0.0f, b.y + 0.0f, b.z + 0.0f, b.w + 0.0f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (4 * k + j < nv) {
                    ll = ll + T.lgamma * (xl[j] - ll); lr = lr + T.lgamma * (xr[j] - lr);
                    hl = hl + T.hgamma * (xl[j] - hl); hr = hr + T.hgamma * (xr[j] - hr);
                    al *= dl; ah *= dh;
                }
            }
        }
        mine = Aff{al, ah, ll, lr, hl, hr};
    }
    Aff total;
    const Aff exc = block_aff_scan(mine, &total);
    // ---- 3. publish, look back
    uint32_t* flags = T.sync + 1;
    if (threadIdx.x < 32) {
        const uint32_t ln = threadIdx.x;
        if (ln == 0) {                                          // publish this strip's map, then the flag
            float* a = T.agg + (size_t)strip * 8;
            a[0] = total.b0; a[1] = total.b1; a[2] = total.b2; a[3] = total.b3; a[4] = total.al; a[5] = total.ah;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(flags + strip) = 1u;
        }
        const uint32_t base = (strip / kBpGroup) * kBpGroup, n_before = strip - base;
        // lane q < n_before: the map of strip base + q (wait for its publication); the other lanes: identity
        Aff m = aff_id();
        if (ln < n_before) {
            while (ld_flag(flags + base + ln) < 1u) {}
            __threadfence();
            const volatile float* a = T.agg + (size_t)(base + ln) * 8;
            m = Aff{a[4], a[5], a[0], a[1], a[2], a[3]};
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {                      // inclusive composition over the lanes, fixed shape
            const Aff up = aff_shfl_up(m, d);
            if (ln >= (uint32_t)d) m = aff_then(up, m);
        }
        // lane 31 holds the map of strips [base, strip) (identity lanes do not change it)
        if (ln == 31) {
            float st[4];
            if (base == 0) {
                if (T.first) {                                  // extensions.rs:664-670: seeded with frame 0 of the pass
                    const float* const* gi = reinterpret_cast<const float* const*>(blob + T.in_off);
                    float l0 = 0.0f, r0 = 0.0f;
                    for (uint32_t q = 0; q < T.n_in; ++q) { l0 += gi[2 * q][0]; r0 += gi[2 * q + 1][0]; }
                    st[0] = l0; st[1] = r0; st[2] = l0; st[3] = r0;
                } else { st[0] = T.state[0]; st[1] = T.state[1]; st[2] = T.state[2]; st[3] = T.state[3]; }
            } else {
                while (ld_flag(flags + base - 1) < 2u) {}
                __threadfence();
                const volatile float* c = T.carry + (size_t)(base - 1) * 4;
                st[0] = c[0]; st[1] = c[1]; st[2] = c[2]; st[3] = c[3];
            }
            const float c0 = fmaf(m.al, st[0], m.b0), c1 = fmaf(m.al, st[1], m.b1);
            const float c2 = fmaf(m.ah, st[2], m.b2), c3 = fmaf(m.ah, st[3], m.b3);
            s_carry[0] = c0; s_carry[1] = c1; s_carry[2] = c2; s_carry[3] = c3;
            if ((strip + 1) % kBpGroup == 0 && (uint64_t)(strip + 1) * kBpStrip < T.h.n_frames) {   // the state leaving this group, for the next one
                float* c = T.carry + (size_t)strip * 4;
                c[0] = fmaf(total.al, c0, total.b0); c[1] = fmaf(total.al, c1, total.b1);
                c[2] = fmaf(total.ah, c2, total.b2); c[3] = fmaf(total.ah, c3, total.b3);
                __threadfence();
                *reinterpret_cast<volatile uint32_t*>(flags + strip) = 2u;
            }
        }
    }
    __syncthreads();
    // ---- 4. the exact recurrence from the carried-in state; outputs replace the inputs in the thread's row
    {
        float ll = fmaf(exc.al, s_carry[0], exc.b0), lr = fmaf(exc.al, s_carry[1], exc.b1);
        float hl = fmaf(exc.ah, s_carry[2], exc.b2), hr = fmaf(exc.ah, s_carry[3], exc.b3);
        const float lmul = T.lgamma == 0.0f ? 0.0f : 1.0f, hmul = T.hgamma == 0.0f ? 0.0f : 1.0f;
        const float pass_mul = T.pass ? 1.0f : 0.0f, cut_mul = 1.0f - pass_mul;
#pragma unroll 2
        for (int k = 0; k < kBpFpt / 4; ++k) {
            if (4 * k >= nv) break;
            const float4 a4 = *reinterpret_cast<const float4*>(row_l + 4 * k), b4 = *reinterpret_cast<const float4*>(row_r + 4 * k);
            const float xl[4] = {a4.x + 0.0f, a4.y + 0.0f, a4.z + 0.0f, a4.w + 0.0f}, xr[4] = {b4.x + 0.0f, b4.y + 0.0f, b4.z + 0.0f, b4.w + 0.0f};
            float ol[4], orr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (4 * k + j < nv) {                                   // extensions.rs:671-688
                    const float l = xl[j], r = xr[j];
                    ll = ll + T.lgamma * (l - ll); lr = lr + T.lgamma * (r - lr);
                    hl = hl + T.hgamma * (l - hl); hr = hr + T.hgamma * (r - hr);
                    const float cutl = (lmul * ll + hmul * (l - hl)) * 0.5f;
                    const float cutr = (lmul * lr + hmul * (r - hr)) * 0.5f;
                    const float passl = l - cutl;
                    const float passr = r - cutl;                       // sic (extensions.rs:685)
                    float a = cutl * cut_mul + passl * pass_mul;
                    float b = cutr * cut_mul + passr * pass_mul;
                    apply_pg(a, b, T.h.pg);
                    ol[j] = a; orr[j] = b;
                } else { ol[j] = 0.0f; orr[j] = 0.0f; }
            }
            *reinterpret_cast<float4*>(row_l + 4 * k) = make_float4(ol[0], ol[1], ol[2], ol[3]);
            *reinterpret_cast<float4*>(row_r + 4 * k) = make_float4(orr[0], orr[1], orr[2], orr[3]);
        }
        // the run holding the last frame of the pass leaves the state for the next pass
        if (nv > 0 && pb + f0 + nv == T.h.n_frames) { T.state[0] = ll; T.state[1] = lr; T.state[2] = hl; T.state[3] = hr; }
    }
    __syncthreads();
    // ---- 5. shared memory -> output, coalesced
#pragma unroll 4
    for (int i = 0; i < kBpFpt / 4; ++i) {
        const uint32_t q = threadIdx.x + kBpFusedThreads * i, f = 4 * q;
        if (f >= n_here) break;
        const int nvq = (int)min(4u, n_here - f);
        const uint32_t w = (q >> 3) * kBpRow + 4 * (q & 7);
        const float4 a4 = *reinterpret_cast<const float4*>(s_l + w), b4 = *reinterpret_cast<const float4*>(s_r + w);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
        if (VEC && nvq == 4) { st4<true>(T.h.out_l, pb + f, 4, a); st4<true>(T.h.out_r, pb + f, 4, b); }
        else { st4<false>(T.h.out_l, pb + f, nvq, a); st4<false>(T.h.out_r, pb + f, nvq, b); }
    }
}

// ------------------------------------------------------------------------------------ K9b: quantise
// (x * amplitude) as i16 | i32 (state.rs:517-532): toward zero, saturating, NaN -> 0; bytes laid out
// as hound writes them (8-bit offset binary, 16/24/32-bit little endian, L then R).
__device__ __forceinline__ int quant16(float x, float amp) {
    int v = __float2int_rz(x * amp);
    return max(-32768, min(32767, v));
}
// A block = `span` consecutive output frames of one task: one chunk when the output Normalize's scale is
// fused in (its 1/max is then a per-block constant), 1024 frames otherwise.  The bit depth and the fused
// flag are block-uniform, so the body is compiled once per (depth, fused) pair and picked by one switch.
template <int BD, bool FUSED>
__device__ __forceinline__ void quantise_block(const QuantTask& T, uint64_t b0, uint32_t span, float f, bool& wide) {
    const bool vec_in = ((reinterpret_cast<uintptr_t>(T.in_l + b0) | reinterpret_cast<uintptr_t>(T.in_r + b0)) & 15) == 0;
    for (uint32_t i = threadIdx.x * 4; i < span; i += TDG_THREADS * 4) {
        const uint64_t f0 = b0 + i;
        if (f0 >= T.n) break;
        const int nv = (int)min((uint64_t)4, min((uint64_t)(span - i), T.n - f0));
        float l[4], r[4];
        if (nv == 4 && vec_in) {
            const float4 a = *reinterpret_cast<const float4*>(T.in_l + f0), b = *reinterpret_cast<const float4*>(T.in_r + f0);
            l[0] = a.x; l[1] = a.y; l[2] = a.z; l[3] = a.w; r[0] = b.x; r[1] = b.y; r[2] = b.z; r[3] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) { l[j] = j < nv ? T.in_l[f0 + j] : 0.0f; r[j] = j < nv ? T.in_r[f0 + j] : 0.0f; }
        }
        int v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (FUSED) { l[j] *= f; r[j] *= f; apply_pg(l[j], r[j], T.pg); }   // x * (1/max), pan, gain (extensions.rs:328,262-263)
            if (BD <= 16) { v[2 * j] = quant16(l[j], T.amp); v[2 * j + 1] = quant16(r[j], T.amp); }
            else { v[2 * j] = __float2int_rz(l[j] * T.amp); v[2 * j + 1] = __float2int_rz(r[j] * T.amp); }
        }
        if (BD == 8) {
            uint8_t* o = T.out + f0 * 2;
            uint32_t w[2] = {0, 0};
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                wide |= (v[q] < -128 || v[q] > 127);
                w[q >> 2] |= (uint32_t)((v[q] + 128) & 0xff) << (8 * (q & 3));
            }
            if (nv == 4 && (f0 & 3) == 0) *reinterpret_cast<uint2*>(o) = make_uint2(w[0], w[1]);
            else for (int q = 0; q < nv * 2; ++q) o[q] = (uint8_t)((v[q] + 128) & 0xff);
        } else if (BD == 16) {
            uint8_t* o = T.out + f0 * 4;
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) w[q] = (uint32_t)(v[2 * q] & 0xffff) | ((uint32_t)(v[2 * q + 1] & 0xffff) << 16);
            if (nv == 4 && (f0 & 3) == 0) *reinterpret_cast<uint4*>(o) = make_uint4(w[0], w[1], w[2], w[3]);
            else for (int q = 0; q < nv; ++q) reinterpret_cast<uint32_t*>(o)[q] = w[q];
        } else if (BD == 24) {
            uint8_t* o = T.out + f0 * 6;
#pragma unroll
            for (int q = 0; q < 8; ++q) wide |= (v[q] < -8388608 || v[q] > 8388607);
            // 8 values x 3 little-endian bytes = six 32-bit words, assembled in registers
            uint32_t u[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) u[q] = (uint32_t)v[q] & 0xffffffu;
            const uint32_t w0 = u[0] | (u[1] << 24), w1 = (u[1] >> 8) | (u[2] << 16), w2 = (u[2] >> 16) | (u[3] << 8);
            const uint32_t w3 = u[4] | (u[5] << 24), w4 = (u[5] >> 8) | (u[6] << 16), w5 = (u[6] >> 16) | (u[7] << 8);
            if (nv == 4 && (f0 & 3) == 0) {                   // 24-byte records at multiples of 4 frames are 8-byte aligned
                uint2* o2 = reinterpret_cast<uint2*>(o);
                o2[0] = make_uint2(w0, w1); o2[1] = make_uint2(w2, w3); o2[2] = make_uint2(w4, w5);
            } else {
                const uint32_t ww[6] = {w0, w1, w2, w3, w4, w5};
                for (int q = 0; q < nv * 6; ++q) o[q] = (uint8_t)(ww[q >> 2] >> (8 * (q & 3)));
            }
        } else {
            uint8_t* o = T.out + f0 * 8;
            if (nv == 4 && (f0 & 1) == 0) {
                reinterpret_cast<uint4*>(o)[0] = make_uint4((uint32_t)v[0], (uint32_t)v[1], (uint32_t)v[2], (uint32_t)v[3]);
                reinterpret_cast<uint4*>(o)[1] = make_uint4((uint32_t)v[4], (uint32_t)v[5], (uint32_t)v[6], (uint32_t)v[7]);
            } else for (int q = 0; q < nv * 2; ++q) reinterpret_cast<uint32_t*>(o)[q] = (uint32_t)v[q];
        }
    }
}
#ifndef TDG_QUANT_CPB
#define TDG_QUANT_CPB 4         // spans (chunks) per block: 1 -> 1.59, 4 -> 1.53 ms per bench step
#endif
__global__ void __launch_bounds__(TDG_THREADS) k_quantise(const QuantTask* __restrict__ tasks, uint32_t cpb) {
    __shared__ QuantTask T;
    load_task_pdl(&T, tasks + blockIdx.y);
    const uint32_t span = T.span;
    bool wide = false;
#pragma unroll 1
    for (uint32_t c = 0; c < cpb; ++c) {
    const uint64_t blk = (uint64_t)blockIdx.x * cpb + c;
    const uint64_t b0 = blk * span;
    if (b0 >= T.n) break;
    if (T.runmax) {
        const float f = 1.0f / T.runmax[T.chunk0 + blk];             // scale(len, 1.0 / max) of this chunk
        switch (T.bd) {
            case 8: quantise_block<8, true>(T, b0, span, f, wide); break;
            case 16: quantise_block<16, true>(T, b0, span, f, wide); break;
            case 24: quantise_block<24, true>(T, b0, span, f, wide); break;
            default: quantise_block<32, true>(T, b0, span, f, wide); break;
        }
    } else {
        switch (T.bd) {
            case 8: quantise_block<8, false>(T, b0, span, 1.0f, wide); break;
            case 16: quantise_block<16, false>(T, b0, span, 1.0f, wide); break;
            case 24: quantise_block<24, false>(T, b0, span, 1.0f, wide); break;
            default: quantise_block<32, false>(T, b0, span, 1.0f, wide); break;
        }
    }
    }
    if (wide) atomicOr(T.err, 1u);        // hound's TooWide (8/24 bit); padding lanes were zeroed, only real frames count
}

// ------------------------------------------------------------------------------------ K3: sinc resampler
// rubato SincFixedIn<f32>, Linear, as restated in DESIGN.md: output j of chunk c sits at
// idx = last_index_c + (j+1)*t_ratio (f64), interpolating two 256-tap windows of neighbouring
// sub-phases.  The per-chunk (last_index, n_out) come from the host's exact f64 replay.
// positions first: one thread per chunk repeats the reference's `idx += t_ratio` from the chunk's carried-in index
// (a closed form (j+1)*t_ratio + last differs in the last bits, and the sub-phase fraction is rounded to f32)
__global__ void __launch_bounds__(64) k_resample_index(const ResampleTask* __restrict__ tasks, const uint8_t* __restrict__ blob) {
    const ResampleTask T = tasks[blockIdx.y];
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= T.n_chunks) return;
    const ResampleChunk C = reinterpret_cast<const ResampleChunk*>(blob + T.chunk_off)[c];
    double idx = C.last_index;
    double* out = T.idx + C.out_off;
    for (uint32_t j = 0; j < C.n_out; ++j) { idx += T.t_ratio; out[j] = idx; }
}
// A block = 128 consecutive output frames.  Their 256-tap windows overlap almost entirely (the window start moves
// by t_ratio ~ 1.1-2.2 input frames per output), so the block stages the span of input frames it touches in shared
// memory once (both channels, the history / halo / leading-zero rules applied while staging) and every thread
// runs its 4 x 256 multiply-adds from there; the two sinc rows of a frame are read as float4 and shared by both
// channels.  Before: 2048 scalar global loads per output frame, the kernel sat on the load/store unit
// (ncu: lg_throttle 83 stall cycles per issue, profiles/r02).  Accumulation order is SincFixedIn's: eight partial
// sums over taps t, t+8, ..., added 0..7 at the end, per channel and per sub-phase.
static const int kResampleSpan = 1024;               // staged input frames per block (128 * t_ratio + 257 <= 1024 up to t_ratio 5.9)
__global__ void __launch_bounds__(128) k_resample(const ResampleTask* __restrict__ tasks, const uint8_t* __restrict__ blob, uint64_t n_out_total) {
    __shared__ float s_l[kResampleSpan], s_r[kResampleSpan];
    __shared__ long long s_min[4], s_max[4];
    const ResampleTask T = tasks[blockIdx.y];
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const ResampleChunk* ch = reinterpret_cast<const ResampleChunk*>(blob + T.chunk_off);
    if (T.n_chunks == 0) return;
    const uint64_t total = ch[T.n_chunks - 1].out_off + ch[T.n_chunks - 1].n_out;
    const uint64_t lim = total < n_out_total ? total : n_out_total;
    if ((uint64_t)blockIdx.x * blockDim.x >= lim) return;           // whole block past the end (uniform)
    const bool live = g < lim;
    const int sinc_len = 256, factor = 256;
    int64_t pos0 = 0, pos1 = 0, s0 = 0, s1 = 0;
    float frac = 0.0f;
    if (live) {
        uint32_t a = 0, b = T.n_chunks;                  // last chunk with out_off <= g
        while (b - a > 1) { uint32_t m = (a + b) >> 1; if (ch[m].out_off <= g) a = m; else b = m; }
        const double idx = T.idx[g];
        const double fl = floor(idx);
        const int64_t i0 = (int64_t)fl;
        s0 = (int64_t)floor((idx - fl) * (double)factor);
        int64_t i1 = i0; s1 = s0 + 1;
        if (s1 >= factor) { s1 -= factor; i1 += 1; }
        const double xf = idx * (double)factor;
        frac = (float)(xf - floor(xf));
        const int64_t base = (int64_t)a * (int64_t)T.chunk;
        pos0 = base + i0; pos1 = base + i1;
    }
    // span of stream positions the block's windows cover
    long long lo = live ? (long long)pos0 : (long long)0x7fffffffffffffffLL, hi = live ? (long long)pos1 + sinc_len : (long long)(-0x7fffffffffffffffLL - 1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = lo; s_max[threadIdx.x >> 5] = hi; }
    __syncthreads();
    lo = s_min[0]; hi = s_max[0];
#pragma unroll
    for (int w = 1; w < 4; ++w) { lo = s_min[w] < lo ? s_min[w] : lo; hi = s_max[w] > hi ? s_max[w] : hi; }
    // a stream position: inside this segment's buffer (warm-up chunks sit at negative offsets), in front of it the
    // previous segment's tail (time sharding) or the zeros SincFixedIn's buffer starts with; behind the end nothing
    auto fetch = [&](const float* in, const float* halo, int64_t ps) -> float {
        if (ps >= -(int64_t)T.n_before) return ps < (int64_t)T.n_in ? __ldg(in + ps) : 0.0f;
        if (halo && ps >= -(int64_t)kResampleHalo) return halo[kResampleHalo + ps];
        return 0.0f;
    };
    const bool staged = hi - lo <= (long long)kResampleSpan;
    if (staged) {
        for (int k = threadIdx.x; k < (int)(hi - lo); k += 128) {
            s_l[k] = fetch(T.in_l, T.halo_l, (int64_t)lo + k);
            s_r[k] = fetch(T.in_r, T.halo_r, (int64_t)lo + k);
        }
    }
    __syncthreads();
    if (!live) return;
    float pl[2], pr[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int64_t pos = q == 0 ? pos0 : pos1;
        const float4* sn = reinterpret_cast<const float4*>(T.sincs + (size_t)(q == 0 ? s0 : s1) * sinc_len);
        float al[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ar[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (staged) {
            const float* wl = s_l + (pos - lo);
            const float* wr = s_r + (pos - lo);
            for (int t = 0; t < sinc_len; t += 8) {
                const float4 c0 = __ldg(sn + (t >> 2)), c1 = __ldg(sn + (t >> 2) + 1);
                const float cf[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) { al[u] += wl[t + u] * cf[u]; ar[u] += wr[t + u] * cf[u]; }
            }
        } else {
            for (int t = 0; t < sinc_len; t += 8) {
                const float4 c0 = __ldg(sn + (t >> 2)), c1 = __ldg(sn + (t >> 2) + 1);
                const float cf[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    al[u] += fetch(T.in_l, T.halo_l, pos + t + u) * cf[u];
                    ar[u] += fetch(T.in_r, T.halo_r, pos + t + u) * cf[u];
                }
            }
        }
        pl[q] = al[0] + al[1] + al[2] + al[3] + al[4] + al[5] + al[6] + al[7];
        pr[q] = ar[0] + ar[1] + ar[2] + ar[3] + ar[4] + ar[5] + ar[6] + ar[7];
    }
    T.out_l[g] = (1.0f - frac) * pl[0] + frac * pl[1];
    T.out_r[g] = (1.0f - frac) * pr[0] + frac * pr[1];
}

// Sample::interleave (sample.rs:177-184) for the playback ring: planar L / R -> (L, R) pairs
__global__ void __launch_bounds__(256) k_interleave(const float* __restrict__ l, const float* __restrict__ r, float2* __restrict__ out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float2(l[i], r[i]);
}
// ------------------------------------------------------------------------------------ sample loading
// numeric part of SampleBank::add (sample.rs:275-303)
__global__ void k_deinterleave(const float* __restrict__ lin, uint64_t n_frames, uint32_t stride, uint32_t off_l,
                               uint32_t off_r, uint32_t mix, float* __restrict__ out_l, float* __restrict__ out_r) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_frames) return;
    if (mix) { const float m = lin[i * 2] + lin[i * 2 + 1]; out_l[i] = m; out_r[i] = m; }   // mix_down, sample.rs:139-147
    else { out_l[i] = lin[i * stride + off_l]; out_r[i] = lin[i * stride + off_r]; }
}
__global__ void k_absmax(const float* __restrict__ x, uint64_t n, int* __restrict__ out_bits) {
    float m = 0.0f;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        m = absmax_fold(m, x[i]);
    m = block_max(m);
    if (threadIdx.x == 0) atomicMax(out_bits, __float_as_int(m));    // m >= 0 and never NaN
}
__global__ void k_scale_by_inv(float* __restrict__ x, uint64_t n, const float* __restrict__ max_a, const float* __restrict__ max_b) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = max_b ? fmaxf(*max_a, *max_b) : *max_a;          // scan_max (sample.rs:116-118)
    x[i] *= (1.0f / m);                                              // scale(len, 1.0 / max)
}

// ------------------------------------------------------------------------------------ launchers
static inline dim3 grid2(uint64_t x, uint32_t y) { return dim3((unsigned)(x ? x : 1), y, 1); }   // zero-length pass: blocks find nothing to do

// Launch of a pass kernel (one that calls pdl_wait before it reads what kernels produce) with programmatic stream
// serialization: its blocks may become resident while the last wave of the kernel in front drains (TDG_PDL=0: plain launches).
static bool g_pdl = []() { const char* e = getenv("TDG_PDL"); return e ? atoi(e) != 0 : TDG_PDL_DEFAULT != 0; }();
template <typename... KArgs, typename... Args>
static inline void pdl_launch(void (*k)(KArgs...), dim3 g, dim3 b, size_t smem, cudaStream_t s, Args... args) {
#ifdef __NVCC__
    if (g_pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = g; cfg.blockDim = b; cfg.dynamicSmemBytes = smem; cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k, KArgs(args)...);
        return;
    }
#endif
    k<<<g, b, smem, s>>>(KArgs(args)...);
}

// Chunks per block, chosen per launch: the compile-time maximum while the grid still has >= g_cpb_waves waves of blocks
// (a block that walks 8 chunks is 8 times longer, and a grid of a few waves ends in a long tail: configs[2]'s one
// 512-voice vertex went from 61.1 to 63.9 ms with a fixed 8), halved until it does.
static uint64_t g_cpb_waves = []() -> uint64_t { const char* e = getenv("TDG_LAUNCH_WAVES"); return e ? (uint64_t)atoll(e) : 16; }();
void set_launch_waves(uint32_t waves) { g_cpb_waves = waves; }
static inline uint32_t pick_cpb(uint32_t cpb_max, uint32_t max_chunks, uint32_t n_tasks, uint32_t blocks_per_sm) {
    static int n_sm = 0;
    if (n_sm == 0) { int dev = 0; cudaDeviceProp pr; if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&pr, dev) == cudaSuccess) n_sm = pr.multiProcessorCount; else n_sm = 148; }
    uint32_t cpb = cpb_max;
    while (cpb > 1 && (uint64_t)((max_chunks + cpb - 1) / cpb) * n_tasks < g_cpb_waves * (uint64_t)n_sm * blocks_per_sm) cpb >>= 1;
    return cpb;
}

#define TDG_VEC_DISPATCH(vec, K, ...) do { if (vec) K<true> __VA_ARGS__; else K<false> __VA_ARGS__; } while (0)

template <bool GRP>
static void launch_fx_t(int mode, const FxTask* t, uint32_t n, uint32_t max_chunks, bool vec, bool exact, const uint8_t* blob, cudaStream_t s) {
    const uint32_t c1 = pick_cpb(GRP ? TDG_FXG_CPB : TDG_FX_CPB, max_chunks, n, 6), ca = pick_cpb(TDG_ADSR_CPB, max_chunks, n, TDG_ADSR_MINB);
    const dim3 g = grid2((max_chunks + c1 - 1) / c1, n);
    if (mode == FX_SUM) { if (vec) pdl_launch(k_fx<FX_SUM, true, false, GRP>, g, TDG_THREADS, 0, s, t, blob, c1); else pdl_launch(k_fx<FX_SUM, false, false, GRP>, g, TDG_THREADS, 0, s, t, blob, c1); }
    else if (mode == FX_ADSR) {
        const dim3 ga = grid2((max_chunks + ca - 1) / ca, n);
        if (exact) { if (vec) pdl_launch(k_fx<FX_ADSR, true, true, GRP>, ga, TDG_THREADS, 0, s, t, blob, ca); else pdl_launch(k_fx<FX_ADSR, false, true, GRP>, ga, TDG_THREADS, 0, s, t, blob, ca); }
        else { if (vec) pdl_launch(k_fx<FX_ADSR, true, false, GRP>, ga, TDG_THREADS, 0, s, t, blob, ca); else pdl_launch(k_fx<FX_ADSR, false, false, GRP>, ga, TDG_THREADS, 0, s, t, blob, ca); }
    }
    else { if (vec) pdl_launch(k_fx<FX_NORM_PEAK, true, false, GRP>, g, TDG_THREADS, 0, s, t, blob, c1); else pdl_launch(k_fx<FX_NORM_PEAK, false, false, GRP>, g, TDG_THREADS, 0, s, t, blob, c1); }
}
cudaError_t launch_fx(int mode, const FxTask* t, uint32_t n, uint32_t max_chunks, bool vec, bool exact, bool grouped, const uint8_t* blob, cudaStream_t s) {
    if (grouped) launch_fx_t<true>(mode, t, n, max_chunks, vec, exact, blob, s); else launch_fx_t<false>(mode, t, n, max_chunks, vec, exact, blob, s);
    return cudaGetLastError();
}
cudaError_t launch_env_table(float* clock, uint32_t bl, uint32_t row0, uint32_t depth, float fsr, cudaStream_t s) {
    if (depth <= row0 || bl == 0) return cudaSuccess;
    k_env_table<<<(bl + 255) / 256, 256, 0, s>>>(clock, bl, row0, depth, fsr);
    return cudaGetLastError();
}
cudaError_t launch_chunk_ranges(const RangeTask* t, uint32_t n, uint32_t max_chunks, bool any_segs, const uint8_t* blob, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    pdl_launch(k_chunk_ranges, grid2((max_chunks + 255) / 256, n), 256, 0, s, t, blob);
    if (any_segs) {                                      // voices / adsr: counts -> offsets -> segments
        pdl_launch(k_seg_scan, dim3(n), 1024, 0, s, t);
        pdl_launch(k_seg_fill, grid2((max_chunks + 255) / 256, n), 256, 0, s, t, blob);
    }
    return cudaGetLastError();
}
cudaError_t launch_norm_prefix(const NormPrefixTask* t, uint32_t n, cudaStream_t s) {
    pdl_launch(k_norm_prefix, dim3(n), 1024, 0, s, t);
    return cudaGetLastError();
}
cudaError_t launch_norm_total(const NormPrefixTask* t, uint32_t n, float* reduce_buf, cudaStream_t s) {
    k_norm_total<<<n, 256, 0, s>>>(t, reduce_buf);
    return cudaGetLastError();
}
cudaError_t launch_bp_total(const BpTask* t, uint32_t n, const uint8_t* blob, cudaStream_t s) {
    k_bp_total<<<n, kBpThreads, 0, s>>>(t, blob);
    return cudaGetLastError();
}
cudaError_t launch_norm_scale(const NormScaleTask* t, uint32_t n, uint32_t max_chunks, bool vec, cudaStream_t s) {
    if (vec) pdl_launch(k_norm_scale<true>, grid2(max_chunks, n), TDG_THREADS, 0, s, t); else pdl_launch(k_norm_scale<false>, grid2(max_chunks, n), TDG_THREADS, 0, s, t);
    return cudaGetLastError();
}
cudaError_t launch_sample_loop(const LoopTask* t, uint32_t n, uint32_t max_chunks, bool vec, cudaStream_t s) {
    const uint32_t cpb = pick_cpb(TDG_GATHER_CPB, max_chunks, n, 6);
    if (vec) pdl_launch(k_sample_loop<true>, grid2((max_chunks + cpb - 1) / cpb, n), TDG_THREADS, 0, s, t, cpb); else pdl_launch(k_sample_loop<false>, grid2((max_chunks + cpb - 1) / cpb, n), TDG_THREADS, 0, s, t, cpb);
    return cudaGetLastError();
}
cudaError_t launch_sample_multi(const MultiTask* t, uint32_t n, uint32_t max_chunks, bool vec, const uint8_t* blob, cudaStream_t s) {
    const uint32_t cpb = pick_cpb(TDG_GATHER_CPB, max_chunks, n, 6);
    if (vec) pdl_launch(k_sample_multi<true>, grid2((max_chunks + cpb - 1) / cpb, n), TDG_THREADS, 0, s, t, blob, cpb); else pdl_launch(k_sample_multi<false>, grid2((max_chunks + cpb - 1) / cpb, n), TDG_THREADS, 0, s, t, blob, cpb);
    return cudaGetLastError();
}
cudaError_t launch_sample_lerp(const LerpTask* t, uint32_t n, uint32_t max_chunks, bool vec, const uint8_t* blob, cudaStream_t s) {
    const uint32_t cpb = pick_cpb(TDG_GATHER_CPB, max_chunks, n, 6);
    if (vec) pdl_launch(k_sample_lerp<true>, grid2((max_chunks + cpb - 1) / cpb, n), TDG_THREADS, 0, s, t, blob, cpb); else pdl_launch(k_sample_lerp<false>, grid2((max_chunks + cpb - 1) / cpb, n), TDG_THREADS, 0, s, t, blob, cpb);
    return cudaGetLastError();
}

// Occupancy cap of the voice kernels: they are FP32-issue bound and need no bandwidth, the streaming
// kernels of the other lane are the opposite, so leaving part of every SM to the latter lets the two
// co-reside (DESIGN.md "co-scheduling").  The cap is expressed as reserved dynamic shared memory.
static int g_voice_smem = -1;
void set_voice_reserved_smem(int bytes) { g_voice_smem = bytes; }
template <int KIND, bool EXACT, bool LIBM>
static cudaError_t launch_voice_t(const VoiceTask* t, uint32_t n, uint32_t max_chunks, const uint8_t* blob, cudaStream_t s) {
    if (g_voice_smem < 0) { const char* e = getenv("TDG_VOICE_SMEM"); g_voice_smem = e ? atoi(e) : 0; }
    auto kfn = k_voice<KIND, EXACT, KIND == VOICE_SYNTH ? TDG_SYNTH_F : 4, LIBM>;
    static int attr_set = -1;
    if (g_voice_smem > 48 * 1024 - 8192 && attr_set != g_voice_smem) {
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, g_voice_smem);
        if (e != cudaSuccess) return e;
        attr_set = g_voice_smem;
    }
    const uint32_t cpb = pick_cpb(TDG_VOICE_CPB, max_chunks, n, 5);
    pdl_launch(kfn, grid2((max_chunks + cpb - 1) / cpb, n), TDG_THREADS, (size_t)g_voice_smem, s, t, blob, cpb);
    return cudaGetLastError();
}
cudaError_t launch_voice(int kind, const VoiceTask* t, uint32_t n, uint32_t max_chunks, bool exact, bool libm_sin, const uint8_t* blob, cudaStream_t s) {
#define TDG_VOICE_DISPATCH(K) \
    (exact ? (libm_sin ? launch_voice_t<K, true, true>(t, n, max_chunks, blob, s) : launch_voice_t<K, true, false>(t, n, max_chunks, blob, s)) \
           : (libm_sin ? launch_voice_t<K, false, true>(t, n, max_chunks, blob, s) : launch_voice_t<K, false, false>(t, n, max_chunks, blob, s)))
    if (kind == VOICE_DEBUG_SINE) return TDG_VOICE_DISPATCH(VOICE_DEBUG_SINE);
    if (kind == VOICE_SYNTH) return TDG_VOICE_DISPATCH(VOICE_SYNTH);
    // the wavetable voices evaluate no sine: one instantiation
    return exact ? launch_voice_t<VOICE_SAMPSYN, true, false>(t, n, max_chunks, blob, s) : launch_voice_t<VOICE_SAMPSYN, false, false>(t, n, max_chunks, blob, s);
#undef TDG_VOICE_DISPATCH
}

cudaError_t launch_bp_aggregate(const BpTask* t, uint32_t n, uint32_t max_tiles, bool vec, const uint8_t* blob, cudaStream_t s) {
    TDG_VEC_DISPATCH(vec, k_bp_aggregate, <<<grid2(max_tiles, n), kBpThreads, 0, s>>>(t, blob));
    return cudaGetLastError();
}
cudaError_t launch_bp_carry(const BpTask* t, uint32_t n, const uint8_t* blob, cudaStream_t s) {
    k_bp_carry<<<n, kBpThreads, 0, s>>>(t, blob);
    return cudaGetLastError();
}
cudaError_t launch_bp_apply(const BpTask* t, uint32_t n, uint32_t max_tiles, bool vec, const uint8_t* blob, cudaStream_t s) {
    TDG_VEC_DISPATCH(vec, k_bp_apply, <<<grid2(max_tiles, n), kBpThreads, 0, s>>>(t, blob));
    return cudaGetLastError();
}
cudaError_t launch_bp_fused(const BpTask* t, uint32_t n, uint32_t max_tiles, bool vec, const uint8_t* blob, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {          // 72 KB of dynamic shared memory per block: opt in once (three blocks per SM)
        cudaError_t e = cudaFuncSetAttribute(k_bp_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBpStripSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_bp_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBpStripSmem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const uint32_t max_strips = (max_tiles + kBpStripTiles - 1) / kBpStripTiles;
    if (vec) pdl_launch(k_bp_fused<true>, dim3(n, max_strips ? max_strips : 1, 1), kBpFusedThreads, kBpStripSmem, s, t, blob); else pdl_launch(k_bp_fused<false>, dim3(n, max_strips ? max_strips : 1, 1), kBpFusedThreads, kBpStripSmem, s, t, blob);
    return cudaGetLastError();
}
cudaError_t launch_quantise(const QuantTask* t, uint32_t n, uint64_t max_blocks, cudaStream_t s) {
    const uint32_t cpb = pick_cpb(TDG_QUANT_CPB, (uint32_t)std::min<uint64_t>(max_blocks, 0xffffffffu), n, 8);
    pdl_launch(k_quantise, grid2(max_blocks ? (max_blocks + cpb - 1) / cpb : 1, n), TDG_THREADS, 0, s, t, cpb);
    return cudaGetLastError();
}
cudaError_t launch_resample(const ResampleTask* t, uint32_t n, uint64_t max_out, uint32_t max_chunks, const uint8_t* blob, cudaStream_t s) {
    k_resample_index<<<grid2((max_chunks + 63) / 64, n), 64, 0, s>>>(t, blob);
    const uint64_t blocks = (max_out + 127) / 128;
    k_resample<<<grid2(blocks ? blocks : 1, n), 128, 0, s>>>(t, blob, max_out);
    return cudaGetLastError();
}
cudaError_t launch_interleave(const float* l, const float* r, float* out_pairs, uint64_t n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    k_interleave<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(l, r, reinterpret_cast<float2*>(out_pairs), n);
    return cudaGetLastError();
}
cudaError_t launch_deinterleave(const float* lin, uint64_t n_frames, uint32_t stride, uint32_t off_l, uint32_t off_r,
                                uint32_t mix, float* out_l, float* out_r, cudaStream_t s) {
    if (n_frames == 0) return cudaSuccess;
    k_deinterleave<<<(unsigned)((n_frames + 255) / 256), 256, 0, s>>>(lin, n_frames, stride, off_l, off_r, mix, out_l, out_r);
    return cudaGetLastError();
}
cudaError_t launch_absmax(const float* x, uint64_t n, int* out_bits, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    unsigned blocks = (unsigned)min((uint64_t)1184, (n + 255) / 256);
    k_absmax<<<blocks, 256, 0, s>>>(x, n, out_bits);
    return cudaGetLastError();
}
cudaError_t launch_scale_by_inv(float* x, uint64_t n, const float* max_a, const float* max_b, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    k_scale_by_inv<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, n, max_a, max_b);
    return cudaGetLastError();
}

}  // namespace tdg
