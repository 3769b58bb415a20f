// tests/emu/cuda_runtime.h — TEST INFRASTRUCTURE: a CPU stand-in for the CUDA runtime and the device
// execution model, so that the REAL kernels.cu / engine.cu can be compiled by g++ and run in the `not gpu`
// tests (tests/test_emulated_kernels.py).  It is never part of the product: the shipped library is built by
// nvcc for sm_100a and fails loudly without a device; this header only exists so that the host planner, the
// launch sequences and the kernels' arithmetic can be checked against the oracle where no GPU exists.
//
// Model: a launch runs its blocks one after the other; the threads of a block are fibers (ucontext) that
// run until they reach __syncthreads() or a warp shuffle, where they yield to the next thread; a barrier
// opens when every live thread of the block (or of the warp) has arrived.  `__shared__` becomes `static`
// (one block is in flight at a time).  Arithmetic is IEEE f32 with -ffp-contract=off, i.e. what nvcc
// generates with -fmad=false; the kernels call no device math library function, so results are the
// device's bit for bit.  Streams are synchronous; events hold host timestamps.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <functional>
#include <type_traits>

// ---- qualifiers -------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__            /* empty: libstdc++ spells __attribute__((__noinline__)), which then reads __attribute__(()) */
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__
#ifndef __restrict__
#define __restrict__ __restrict
#endif

// ---- vector types -------------------------------------------------------------------------------------
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

// ---- execution model ------------------------------------------------------------------------------------
namespace emu {
struct Fiber;
extern thread_local uint3 t_threadIdx, t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
void launch(dim3 grid, dim3 block, const std::function<void()>& body);
void syncthreads();
uint32_t shfl(uint32_t v, int src_lane_or_delta, int mode);     // mode 0: up, 1: xor, 2: idx, 3: down
template <typename T> inline T shfl_t(T v, int a, int mode) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "4- and 8-byte shuffles only");
    uint32_t u[2] = {0, 0}; memcpy(u, &v, sizeof(T));
    u[0] = shfl(u[0], a, mode);
    if (sizeof(T) == 8) u[1] = shfl(u[1], a, mode);       // a 64-bit shuffle is two 32-bit ones (as in SASS)
    T r; memcpy(&r, u, sizeof(T));
    return r;
}
}  // namespace emu
#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)
static inline void __syncthreads() { emu::syncthreads(); }
static inline void __threadfence() {}                       // blocks run one after the other: every earlier store is visible
template <typename T> inline T __shfl_up_sync(unsigned, T v, int d) { return emu::shfl_t(v, d, 0); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::shfl_t(v, m, 1); }
template <typename T> inline T __shfl_sync(unsigned, T v, int l) { return emu::shfl_t(v, l, 2); }
template <typename T> inline T __shfl_down_sync(unsigned, T v, int d) { return emu::shfl_t(v, d, 3); }

// ---- device intrinsics -------------------------------------------------------------------------------------
template <typename T> inline T __ldg(const T* p) { return *p; }
static inline int __float2int_rz(float x) {
    if (!(x == x)) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int)x;
}
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
// one block, one thread at a time: atomics are plain read-modify-writes
template <typename T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
// CUDA's global-namespace min / max over mixed integer types
template <typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> min(A a, B b) { using C = std::common_type_t<A, B>; return (C)a < (C)b ? (C)a : (C)b; }
template <typename A, typename B, typename = std::enable_if_t<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>>
inline std::common_type_t<A, B> max(A a, B b) { using C = std::common_type_t<A, B>; return (C)a > (C)b ? (C)a : (C)b; }

// ---- runtime API ---------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef struct emuStream* cudaStream_t;
struct emuEvent { double t_ms; };
typedef emuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostRegisterDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };

static inline double emu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p); snprintf(p->name, sizeof p->name, "emulated sm_100 (tests/emu)"); p->major = 10; p->minor = 0;
    p->multiProcessorCount = 148; p->totalGlobalMem = (size_t)64 << 30; return cudaSuccess;
}
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) {       // EMU_FREE_MB: pretend the device is nearly full
    const char* e = getenv("EMU_FREE_MB");
    *f = e ? (size_t)atoll(e) << 20 : (size_t)32 << 30; *t = (size_t)64 << 30; return cudaSuccess;
}
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) {
    void* q = nullptr;
    if (posix_memalign(&q, 256, n ? n : 1) != 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
    memset(q, 0xCD, n);                                       // device memory is not zeroed: make reliance on it visible
    *p = (T*)q; return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostRegister(void*, size_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void*) { return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { return cudaStreamCreateWithFlags(s, 0); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{emu_now_ms()}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t_ms = emu_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
