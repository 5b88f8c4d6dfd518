// bb_ptx.cuh - every line of inline PTX the kernels use, in one place: volatile / system-scope
// loads and stores, cp.async (LDGSTS), cp.async.bulk (UBLKCP, the TMA engine) with mbarrier
// completion, programmatic dependent launch, and the launch helper.
//
// With -DBB_EMU (tools/emu, a development tool that runs the kernels' logic as fibers on the host;
// never part of the shipped library) the same names map to plain C++.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bb {

#ifdef BB_EMU
#define BB_SPIN_YIELD() emu::yield()
#define BB_DYN_SMEM(name) unsigned char* name = emu::g_dyn_smem
#else
#define BB_SPIN_YIELD() ((void)0)
#define BB_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

__device__ __forceinline__ uint32_t lanemask_lt() {
#ifdef BB_EMU
  return (1u << (threadIdx.x & 31u)) - 1u;
#else
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
#endif
}
__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) {
#ifdef BB_EMU
  return *(const volatile uint32_t*)p;
#else
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) {
#ifdef BB_EMU
  *(volatile uint32_t*)p = v;
#else
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ uint64_t ld_sys(const uint64_t* p) {
#ifdef BB_EMU
  return *(const volatile uint64_t*)p;
#else
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void st_sys(uint64_t* p, uint64_t v) {
#ifdef BB_EMU
  *(volatile uint64_t*)p = v;
#else
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
// streaming 16-byte load that does not allocate in L1 (rows and gathered payloads are touched once)
__device__ __forceinline__ uint4 ld_stream16(const uint4* p) {
#ifdef BB_EMU
  return *p;
#else
  uint4 v;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
#endif
}

// bring a 128-byte line into L2 ahead of its use; a hint, never a correctness dependency
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef BB_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// read one word so that its 32-byte sector is in L2 when the atomic that follows a little later needs it (prefetch.L2
// would fetch the whole 128-byte line); the value is handed back so that the caller can keep the load alive
__device__ __forceinline__ uint32_t touch_l2(const void* p) {
#ifdef BB_EMU
  return 0;
#else
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
#endif
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
#ifdef BB_EMU
  return 0;
#else
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#endif
}

// ---- cp.async (LDGSTS): 16 bytes global -> shared, no registers held
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
#ifdef BB_EMU
  memcpy(smem_dst, gmem_src, 16);
#else
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef BB_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
#ifndef BB_EMU
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_wait_all() {
  cp_async_commit();
  cp_async_wait_group<0>();
}

// ---- mbarrier (shared memory, 64-bit) + cp.async.bulk (UBLKCP)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
#ifdef BB_EMU
  emu::mbar_init(bar, count);
#else
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
#endif
}
// make the initialised barrier visible to the async proxy (the copy engine) before the first bulk copy names it
__device__ __forceinline__ void mbar_fence_init() {
#ifndef BB_EMU
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
#ifdef BB_EMU
  emu::mbar_arrive(bar, 0);
#else
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
#ifdef BB_EMU
  emu::mbar_arrive(bar, bytes);
#else
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
               : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef BB_EMU
  emu::mbar_wait(bar, parity);
#else
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "BB_MBAR_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra BB_MBAR_DONE;\n\t"
      "bra BB_MBAR_WAIT;\n\t"
      "BB_MBAR_DONE:\n\t}" ::"r"(a),
      "r"(parity)
      : "memory");
#endif
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
#ifdef BB_EMU
  if (((uintptr_t)smem_dst | (uintptr_t)gmem_src | bytes) & 15u) abort();
  memcpy(smem_dst, gmem_src, bytes);
  emu::mbar_complete_tx(bar, bytes);
#else
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
#endif
}
// shared -> global (peer memory included), completion through the bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
#ifdef BB_EMU
  if (((uintptr_t)smem_src | (uintptr_t)gmem_dst | bytes) & 15u) abort();
  memcpy(gmem_dst, smem_src, bytes);
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"((uint32_t)__cvta_generic_to_shared(smem_src)), "r"(bytes)
               : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() {
#ifndef BB_EMU
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_wait_read_all() {  // the sources may be overwritten
#ifndef BB_EMU
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
__device__ __forceinline__ void bulk_wait_all() {  // the writes are complete
#ifndef BB_EMU
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}
// generic-proxy writes to shared memory -> visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() {
#ifndef BB_EMU
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// ---- programmatic dependent launch: the next kernel of the stream may start its prologue while this
// one drains; it must not touch anything this one writes before pdl_wait()
__device__ __forceinline__ void pdl_launch_dependents() {
#ifndef BB_EMU
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#ifndef BB_EMU
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

}  // namespace bb

// ---------------------------------------------------------------- launch helper
// bb_launch(kernel, grid, block, smem, stream, pdl, args...): pdl = the kernel may start before the
// previous kernel of the stream has finished (it calls pdl_wait() before touching that kernel's output).
#ifdef BB_EMU
template <class... KArgs, class... Args>
static inline cudaError_t bb_launch(void (*kernel)(KArgs...), uint32_t grid, uint32_t block, size_t smem, cudaStream_t,
                                    bool, Args&&... args) {
  emu::run_grid(dim3(grid), dim3(block), smem, [&] { kernel(KArgs(args)...); });
  return cudaSuccess;
}
#else
template <class... KArgs, class... Args>
static inline cudaError_t bb_launch(void (*kernel)(KArgs...), uint32_t grid, uint32_t block, size_t smem, cudaStream_t s,
                                    bool pdl, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
