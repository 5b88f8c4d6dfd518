// tools/emu/emu.cpp - the fiber scheduler and the fake runtime behind tools/emu/cuda_runtime.h.
// DEVELOPMENT TOOL (see the header): one host thread, one fiber per CUDA thread of the running CTA.
#include <stdio.h>
#include <time.h>

#include <vector>

#include "cuda_runtime.h"
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim

// minimal x86-64 context switch (callee-saved registers + stack pointer): ~10 ns instead of swapcontext's
// two sigprocmask system calls
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

namespace emu {

uint3 g_threadIdx, g_blockIdx;
dim3 g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;

namespace {

constexpr size_t STACK_BYTES = 256 << 10;
constexpr int MAX_THREADS = 1024;

struct Warp {
  uint32_t alive = 0;        // lanes that have not returned
  uint64_t gen[32] = {};     // collectives this lane has entered
  uint64_t val[2][32] = {};  // deposits, by generation parity
};

struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  bool done = true;
  uint64_t bar_gen = 0;  // __syncthreads generations this thread has entered
};

void* g_sched_sp = nullptr;
Fiber g_fib[MAX_THREADS];
Warp g_warp[MAX_THREADS / 32];
int g_nthreads = 0, g_cur = -1, g_live = 0;
uint64_t g_bar_gen = 0;       // completed __syncthreads generations
int g_bar_arrived = 0, g_bar_or = 0, g_bar_or_result = 0;
uint64_t g_progress = 0;      // bumped whenever any fiber gets past a wait (deadlock detection)
const std::function<void()>* g_body = nullptr;
bool g_in_grid = false;

[[noreturn]] void die(const char* what) {
  fprintf(stderr, "[emu] FATAL: %s (block %u,%u thread %d of %d)\n", what, g_blockIdx.x, g_blockIdx.y, g_cur, g_nthreads);
  abort();
}

void fiber_main() {
  (*g_body)();
  Fiber& f = g_fib[g_cur];
  f.done = true;
  --g_live;
  g_warp[g_cur >> 5].alive &= ~(1u << (g_cur & 31));
  ++g_progress;
  // a thread that exits counts as arrived at every later barrier
  emu_switch(&f.sp, g_sched_sp);
  die("resumed a finished fiber");
}

void release_barrier_if_complete() {
  if (g_live > 0 && g_bar_arrived == g_live) {
    g_bar_or_result = g_bar_or;
    g_bar_or = 0;
    g_bar_arrived = 0;
    ++g_bar_gen;
    ++g_progress;
  }
}

}  // namespace

void yield() {
  if (!g_in_grid) return;
  Fiber& f = g_fib[g_cur];
  emu_switch(&f.sp, g_sched_sp);
}

// wait (yielding) until cond() holds; the scheduler detects "nobody can move"
template <class C>
static inline void wait_until(C cond) {
  while (!cond()) yield();
  ++g_progress;
}

int syncthreads_or(int pred) {
  Fiber& f = g_fib[g_cur];
  const uint64_t my = ++f.bar_gen;
  if (my != g_bar_gen + 1) die("__syncthreads generation mismatch (a thread skipped a barrier)");
  ++g_bar_arrived;
  g_bar_or |= pred ? 1 : 0;
  release_barrier_if_complete();
  wait_until([&] { return g_bar_gen >= my; });
  return g_bar_or_result;
}
void syncthreads() { (void)syncthreads_or(0); }

// every lane of `mask` that is still alive deposits and waits for the others
static inline void warp_rendezvous(uint32_t mask, uint64_t v, uint64_t* out /*[32]*/, uint32_t* present) {
  const int lane = g_cur & 31;
  Warp& w = g_warp[g_cur >> 5];
  if (!((mask >> lane) & 1u)) die("warp collective: calling lane is not in the mask");
  const uint64_t my = ++w.gen[lane];
  w.val[my & 1][lane] = v;
  wait_until([&] {
    const uint32_t need = mask & w.alive;
    for (int j = 0; j < 32; ++j)
      if (((need >> j) & 1u) && w.gen[j] < my) return false;
    return true;
  });
  uint32_t pres = 0;
  for (int j = 0; j < 32; ++j) {
    const bool in = ((mask >> j) & 1u) && w.gen[j] >= my;  // exited lanes contribute nothing
    out[j] = in ? w.val[my & 1][j] : 0;
    pres |= in ? (1u << j) : 0u;
  }
  for (int j = 0; j < 32; ++j)
    if (((pres >> j) & 1u) && w.gen[j] > my + 1) die("warp collective: lanes out of step");
  *present = pres;
}

void syncwarp() {
  uint64_t v[32];
  uint32_t p;
  warp_rendezvous(0xffffffffu, 0, v, &p);
}
uint32_t ballot(uint32_t mask, int pred) {
  uint64_t v[32];
  uint32_t p, r = 0;
  warp_rendezvous(mask, pred ? 1 : 0, v, &p);
  for (int j = 0; j < 32; ++j) r |= (v[j] ? 1u : 0u) << j;
  return r;
}
uint64_t shfl(uint32_t mask, uint64_t val, int src_lane) {
  uint64_t v[32];
  uint32_t p;
  warp_rendezvous(mask, val, v, &p);
  return v[src_lane & 31];
}
uint32_t match_any(uint32_t mask, uint64_t val) {
  uint64_t v[32];
  uint32_t p, r = 0;
  warp_rendezvous(mask, val, v, &p);
  for (int j = 0; j < 32; ++j)
    if (((p >> j) & 1u) && v[j] == val) r |= 1u << j;
  return r;
}

// mbarrier word layout (ours): [63] phase parity, [62:32] signed tx bytes + 2^30, [31:16] expected arrivals, [15:0] pending
namespace {
struct MBar { uint32_t pending : 16, expected : 16; int32_t tx : 31; uint32_t phase : 1; };
static_assert(sizeof(MBar) == 8, "mbarrier emulation fits the 64-bit word");
inline void mbar_maybe_flip(MBar* b) {
  if (b->pending == 0 && b->tx == 0) {
    b->phase ^= 1u;
    b->pending = b->expected;
    ++g_progress;
  }
}
}  // namespace
void mbar_init(uint64_t* bar, uint32_t count) {
  MBar* b = reinterpret_cast<MBar*>(bar);
  b->pending = count;
  b->expected = count;
  b->tx = 0;
  b->phase = 0;
}
void mbar_arrive(uint64_t* bar, uint32_t tx_bytes) {
  MBar* b = reinterpret_cast<MBar*>(bar);
  if (b->pending == 0) die("mbarrier: more arrivals than the barrier was initialised for");
  b->tx += (int32_t)tx_bytes;
  b->pending -= 1;
  mbar_maybe_flip(b);
}
void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  MBar* b = reinterpret_cast<MBar*>(bar);
  b->tx -= (int32_t)bytes;
  mbar_maybe_flip(b);
}
void mbar_wait(uint64_t* bar, uint32_t parity) {
  MBar* b = reinterpret_cast<MBar*>(bar);
  wait_until([&] { return b->phase != (parity & 1u); });
}

void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  const int nt = (int)(block.x * block.y * block.z);
  if (nt <= 0 || nt > MAX_THREADS) die("bad block size");
  if (g_in_grid) die("nested launch");
  static std::vector<unsigned char> dyn;
  if (dyn.size() < smem + 128) dyn.resize(smem + 128);
  g_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 127) & ~(uintptr_t)127);
  g_body = &body;
  g_blockDim = block;
  g_gridDim = grid;
  g_nthreads = nt;
  g_in_grid = true;
  for (uint32_t bz = 0; bz < grid.z; ++bz)
    for (uint32_t by = 0; by < grid.y; ++by)
      for (uint32_t bx = 0; bx < grid.x; ++bx) {
        g_bar_gen = 0;
        g_bar_arrived = 0;
        g_bar_or = 0;
        g_live = nt;
        for (int w = 0; w < (nt + 31) / 32; ++w) {
          g_warp[w] = Warp();
          const int lanes = nt - w * 32 >= 32 ? 32 : nt - w * 32;
          g_warp[w].alive = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
        }
        for (int t = 0; t < nt; ++t) {
          Fiber& f = g_fib[t];
          if (!f.stack) f.stack = (char*)malloc(STACK_BYTES);
          {  // initial frame: six callee-saved registers, then fiber_main as the return target of emu_switch
            uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
            void** fp = (void**)top;
            fp[-1] = nullptr;
            fp[-2] = (void*)&fiber_main;
            for (int k = 3; k <= 8; ++k) fp[-k] = nullptr;
            f.sp = (void*)(fp - 8);
          }
          f.done = false;
          f.bar_gen = 0;
        }
        uint64_t seen = g_progress;
        int idle_passes = 0;
        while (g_live > 0) {
          for (int t = 0; t < nt; ++t) {
            Fiber& f = g_fib[t];
            if (f.done) continue;
            g_cur = t;
            g_threadIdx.x = (uint32_t)t % block.x;
            g_threadIdx.y = ((uint32_t)t / block.x) % block.y;
            g_threadIdx.z = (uint32_t)t / (block.x * block.y);
            g_blockIdx.x = bx;
            g_blockIdx.y = by;
            g_blockIdx.z = bz;
            emu_switch(&g_sched_sp, f.sp);
            if (f.done) release_barrier_if_complete();  // an exit can complete a barrier the others wait at
          }
          if (g_progress == seen) {
            if (++idle_passes > 4) die("deadlock: every live thread of the CTA is waiting (barrier / warp collective / mbarrier / spin)");
          } else {
            idle_passes = 0;
            seen = g_progress;
          }
        }
      }
  g_in_grid = false;
  g_cur = -1;
}

}  // namespace emu

// ---------------------------------------------------------------- runtime API
struct emuStream { int dummy; };
struct emuEvent { double t_ms; };

static double now_ms() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

cudaError_t cudaMalloc(void** p, size_t bytes) {
  void* q = nullptr;
  if (posix_memalign(&q, 256, bytes ? bytes : 256) != 0) return cudaErrorMemoryAllocation;
  memset(q, 0xA5, bytes);  // device memory is NOT zeroed: make reads of uninitialised memory visible
  *p = q;
  return cudaSuccess;
}
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emuStream(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = new emuStream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { if (s != cudaStreamLegacy) delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emuEvent{0}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t_ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t_ms - a->t_ms); return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  memset(p, 0, sizeof(*p));
  p->major = 10;
  p->minor = 0;
  p->multiProcessorCount = 4;  // a small "GPU": persistent kernels loop more per CTA
  p->totalGlobalMem = (size_t)8 << 30;
  strcpy(p->name, "emulated sm_100 (tools/emu)");
  return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
