// tools/emu/cuda_runtime.h - DEVELOPMENT TOOL, never part of the product.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile bullet_js_b200/csrc/*.cu* (with -DBB_EMU)
// into tools/emu/_build/libbulletb200_emu.so: every CUDA thread of a CTA becomes a fiber on ONE
// host thread, CTAs run one after another, __syncthreads / warp collectives / mbarriers are
// rendezvous points between the fibers, "device memory" is host memory.  It exists so that the
// LOGIC of the hand-written kernels (indexing, protocols, barrier counts, mbarrier byte counts) can
// be debugged in the GPU-less build container before GPU minutes are spent; it says nothing about
// races between CTAs or about speed.  Nothing under bullet_js_b200/ knows this file exists:
// capi.load() only ever opens csrc/libbulletb200.so, and that library refuses to run without an
// sm_100 device.  Only tests/emu/ loads the emulated build.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <type_traits>

// ---------------------------------------------------------------- qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

// ---------------------------------------------------------------- vector types
struct uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct dim3 {
  uint32_t x, y, z;
  dim3(uint32_t x_ = 1, uint32_t y_ = 1, uint32_t z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace emu {
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
void yield();               // let the other fibers of the CTA run (inside spin loops)
void syncthreads();
int syncthreads_or(int pred);
void syncwarp();
uint32_t ballot(uint32_t mask, int pred);
uint64_t shfl(uint32_t mask, uint64_t v, int src_lane);          // value of lane src_lane
uint32_t match_any(uint32_t mask, uint64_t v);
// mbarrier (shared memory word, 64 bits): faithful arrive / expect_tx / complete_tx / parity wait
void mbar_init(uint64_t* bar, uint32_t count);
void mbar_arrive(uint64_t* bar, uint32_t tx_bytes);
void mbar_complete_tx(uint64_t* bar, uint32_t bytes);
void mbar_wait(uint64_t* bar, uint32_t parity);
}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

// ---------------------------------------------------------------- device intrinsics
static inline void __syncthreads() { emu::syncthreads(); }
static inline int __syncthreads_or(int p) { return emu::syncthreads_or(p); }
static inline void __syncwarp(uint32_t = 0xffffffffu) { emu::syncwarp(); }
static inline uint32_t __ballot_sync(uint32_t m, int p) { return emu::ballot(m, p); }
static inline int __any_sync(uint32_t m, int p) { return emu::ballot(m, p) != 0; }
static inline int __all_sync(uint32_t m, int p) { return emu::ballot(m, !p) == 0; }
static inline uint32_t __match_any_sync(uint32_t m, uint32_t v) { return emu::match_any(m, v); }
static inline uint32_t __match_any_sync(uint32_t m, uint64_t v) { return emu::match_any(m, v); }
template <class T>
static inline T __shfl_sync(uint32_t m, T v, int src, int width = 32) {
  static_assert(sizeof(T) <= 8, "shfl of <= 64-bit values");
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  const int lane = (int)(emu::g_threadIdx.x & 31u);
  const int base = lane & ~(width - 1);
  raw = emu::shfl(m, raw, base + (src & (width - 1)));
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
template <class T>
static inline T __shfl_up_sync(uint32_t m, T v, unsigned delta) {
  const int lane = (int)(emu::g_threadIdx.x & 31u);
  const int src = lane - (int)delta;
  T got = __shfl_sync(m, v, src < 0 ? lane : src);
  return src < 0 ? v : got;
}
template <class T>
static inline T __shfl_down_sync(uint32_t m, T v, unsigned delta) {
  const int lane = (int)(emu::g_threadIdx.x & 31u);
  const int src = lane + (int)delta;
  T got = __shfl_sync(m, v, src > 31 ? lane : src);
  return src > 31 ? v : got;
}
template <class T>
static inline T __shfl_xor_sync(uint32_t m, T v, int x) {
  const int lane = (int)(emu::g_threadIdx.x & 31u);
  return __shfl_sync(m, v, lane ^ x);
}
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __popcll(uint64_t v) { return __builtin_popcountll(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
static inline int __clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
static inline double __longlong_as_double(long long v) {
  double d;
  memcpy(&d, &v, 8);
  return d;
}
static inline long long __double_as_longlong(double d) {
  long long v;
  memcpy(&v, &d, 8);
  return v;
}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}
static inline void __nanosleep(unsigned) { emu::yield(); }
static inline void __trap() { abort(); }
static inline long long clock64() { return 0; }
template <class T>
static inline T __ldcg(const T* p) { return *p; }
template <class T>
static inline T __ldcs(const T* p) { return *p; }
template <class T>
static inline T __ldg(const T* p) { return *p; }

// fibers never pre-empt one another, so plain read-modify-write IS atomic
template <class T, class U>
static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U>
static inline T atomicSub(T* p, U v) { T o = *p; *p = (T)(o - (T)v); return o; }
template <class T, class U>
static inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U>
static inline T atomicAnd(T* p, U v) { T o = *p; *p = (T)(o & (T)v); return o; }
template <class T, class U>
static inline T atomicExch(T* p, U v) { T o = *p; *p = (T)v; return o; }
template <class T, class U>
static inline T atomicMin(T* p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T, class U>
static inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U, class V>
static inline T atomicCAS(T* p, U cmp, V v) { T o = *p; if (o == (T)cmp) *p = (T)v; return o; }

template <class A, class B>
static inline typename std::common_type<A, B>::type min(A a, B b) {
  using C = typename std::common_type<A, B>::type;
  return (C)a < (C)b ? (C)a : (C)b;
}
template <class A, class B>
static inline typename std::common_type<A, B>::type max(A a, B b) {
  using C = typename std::common_type<A, B>::type;
  return (C)a > (C)b ? (C)a : (C)b;
}
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

// ---------------------------------------------------------------- runtime API (synchronous, host memory)
typedef int cudaError_t;
enum : int { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct emuStream* cudaStream_t;
typedef struct emuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; size_t totalGlobalMem; char name[256]; };
struct cudaIpcMemHandle_t { char reserved[64]; };
#define cudaStreamLegacy ((cudaStream_t)0x1)

cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaMallocHost(void** p, size_t bytes);
cudaError_t cudaFreeHost(void* p);
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc((void**)p, bytes); }
template <class T>
static inline cudaError_t cudaMallocHost(T** p, size_t bytes) { return cudaMallocHost((void**)p, bytes); }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t st = 0);
cudaError_t cudaMemset(void* d, int v, size_t n);
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t st = 0);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned flags, int prio);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = 0);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaGetDevice(int* d);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int d);
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi);
cudaError_t cudaGetLastError();
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
