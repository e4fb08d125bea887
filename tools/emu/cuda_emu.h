// cuda_emu.h — TEST INFRASTRUCTURE ONLY: runs the CUDA kernels of csrc/b2env.cu on the host CPU.
//
// The build container has no GPU (nvcc only cross-compiles), and GPU box time is budgeted.  To debug
// kernel LOGIC here — lane mappings, shuffle schedules, shared-memory layouts, collective placement —
// the same .cu source is compiled by g++ with this header force-included (tools/emu/build.sh) into
// libb2env_emu.so.  Every CUDA thread of a block becomes a fiber (own stack, cooperative switching);
// warp collectives (__shfl_sync, __ballot_sync, __syncwarp ...) and __syncthreads are rendezvous points:
// a lane runs until it reaches one, then the next lane runs.  Between two rendezvous points lanes
// execute one after the other, never in lockstep — an adversarial but legal schedule under the
// independent-thread-scheduling model, so a missing __syncwarp shows up as a wrong result here.
// A rendezvous that not every live lane of the warp reaches is reported as a deadlock.
//
// Nothing in the product loads this library: bindings open csrc/libb2env.so; only tests/ may pass the
// emulation library explicitly.  Numerics differ from the GPU in libm (sincosf, atan2f) and FMA
// contraction only; tolerances in the tests account for it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define B2E_EMU 1

// ------------------------------------------------------------------ language surface
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n)
#define __shared__ __thread

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

using std::max;
using std::min;

namespace emu {

struct Warp;
struct Fiber {
  void* sp = nullptr;
  unsigned char* stack = nullptr;
  uint3 tidx{0, 0, 0};
  int lane = 0, state = 0;  // 0 ready, 1 waiting on the warp, 2 waiting on the block, 3 done
  unsigned wait_gen = 0;
  Warp* warp = nullptr;
};
struct Warp {
  unsigned gen = 0, alive_mask = 0;
  int arrived = 0, alive = 0;
  uint32_t vals[2][32];
  int kind[2] = {0, 0};
};
struct Block {
  unsigned gen = 0;
  int arrived = 0, alive = 0;
  int orflag[3] = {0, 0, 0};
};

extern thread_local Fiber* g_cur;
extern thread_local void* g_sched_sp;
extern thread_local Block g_block;
extern thread_local uint3 g_blockIdx;
extern thread_local dim3 g_blockDim, g_gridDim;

extern "C" void emu_switch(void** save_sp, void* load_sp);

inline void yield_to_scheduler() { emu_switch(&g_cur->sp, g_sched_sp); }

[[noreturn]] inline void die(const char* msg) {
  fprintf(stderr, "cuda_emu: %s (block %u, thread %u)\n", msg, g_blockIdx.x, g_cur ? g_cur->tidx.x : 0u);
  abort();
}

// rendezvous of the live lanes of the current warp; `kind` tags the collective so that two lanes meeting at
// different collectives (divergent control flow around a collective) are reported
inline int warp_enter(int kind, uint32_t v) {
  Fiber* f = g_cur;
  Warp& w = *f->warp;
  const int par = w.gen & 1;
  if (w.arrived == 0) w.kind[par] = kind;
  else if (w.kind[par] != kind) die("lanes of one warp met at different collectives");
  w.vals[par][f->lane] = v;
  const unsigned g = w.gen;
  if (++w.arrived == w.alive) {
    w.arrived = 0;
    w.gen++;
  } else {
    f->state = 1;
    f->wait_gen = g;
    while (w.gen == g) yield_to_scheduler();
    f->state = 0;
  }
  return par;
}

inline void block_barrier() {
  Fiber* f = g_cur;
  const unsigned g = g_block.gen;
  if (++g_block.arrived == g_block.alive) {
    g_block.arrived = 0;
    g_block.gen++;
  } else {
    f->state = 2;
    f->wait_gen = g;
    while (g_block.gen == g) yield_to_scheduler();
    f->state = 0;
  }
}

template <class T>
inline uint32_t bits_of(T v) {
  static_assert(sizeof(T) == 4, "4-byte shuffles only");
  uint32_t b;
  memcpy(&b, &v, 4);
  return b;
}
template <class T>
inline T from_bits(uint32_t b) {
  T v;
  memcpy(&v, &b, 4);
  return v;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, void (*body)(void*), void* arg);

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F f) {
  launch(grid, block, smem_bytes, [](void* p) { (*static_cast<F*>(p))(); }, &f);
}

}  // namespace emu

#define threadIdx (emu::g_cur->tidx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)

// ------------------------------------------------------------------ warp / block collectives
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  (void)mask;
  const int lane = emu::g_cur->lane;
  emu::Warp& w = *emu::g_cur->warp;
  const int par = emu::warp_enter(1, emu::bits_of(v));
  const int s = (lane & ~(width - 1)) + (src & (width - 1));
  if (!((w.alive_mask >> s) & 1u)) return v;
  return emu::from_bits<T>(w.vals[par][s]);
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
  (void)mask;
  const int lane = emu::g_cur->lane;
  emu::Warp& w = *emu::g_cur->warp;
  const int par = emu::warp_enter(2, emu::bits_of(v));
  const int s = lane ^ lanemask;
  if ((s & ~(width - 1)) != (lane & ~(width - 1)) || !((w.alive_mask >> s) & 1u)) return v;
  return emu::from_bits<T>(w.vals[par][s]);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
  emu::Warp& w = *emu::g_cur->warp;
  const int par = emu::warp_enter(3, pred ? 1u : 0u);
  unsigned r = 0;
  for (int l = 0; l < 32; l++)
    if (((w.alive_mask >> l) & 1u) && w.vals[par][l]) r |= 1u << l;
  return r & mask;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __all_sync(unsigned mask, int pred) {
  emu::Warp& w = *emu::g_cur->warp;
  return __ballot_sync(mask, pred) == (mask & w.alive_mask);
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
  emu::Warp& w = *emu::g_cur->warp;
  const int par = emu::warp_enter(4, v);
  unsigned r = 0;
  for (int l = 0; l < 32; l++)
    if (((w.alive_mask & mask) >> l) & 1u) r = std::max(r, w.vals[par][l]);
  return r;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  emu::Warp& w = *emu::g_cur->warp;
  const int par = emu::warp_enter(5, v);
  unsigned r = 0;
  for (int l = 0; l < 32; l++)
    if (((w.alive_mask & mask) >> l) & 1u) r |= w.vals[par][l];
  return r;
}
inline void __syncwarp(unsigned mask = 0xffffffffu) {
  (void)mask;
  emu::warp_enter(6, 0u);
}
inline void __syncthreads() { emu::block_barrier(); }
inline int __syncthreads_or(int pred) {   // three rotating accumulators keyed by the barrier generation (see block_barrier)
  int* acc = emu::g_block.orflag;
  const unsigned g = emu::g_block.gen;
  if (pred) acc[g % 3] = 1;
  emu::block_barrier();
  const int r = acc[g % 3];
  acc[(g + 2) % 3] = 0;   // not written before every fiber has left this call and arrived at the next barrier
  return r;
}

// ------------------------------------------------------------------ scalar intrinsics
template <class T>
inline T __ldg(const T* p) { return *p; }
inline unsigned __float_as_uint(float f) { return emu::bits_of(f); }
inline float __uint_as_float(unsigned u) { return emu::from_bits<float>(u); }
inline int __float_as_int(float f) { return (int)emu::bits_of(f); }
inline float __int_as_float(int i) { return emu::from_bits<float>((uint32_t)i); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return x ? __builtin_ctz(x) + 1 : 0; }
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
inline int atomicCAS(int* addr, int compare, int val) {
  const int old = *addr;
  if (old == compare) *addr = val;
  return old;
}
inline int atomicAdd(int* addr, int val) { const int old = *addr; *addr += val; return old; }
inline int atomicExch(int* addr, int val) { const int old = *addr; *addr = val; return old; }
inline long long clock64() { return 0; }

// ------------------------------------------------------------------ runtime API (host memory stands in for HBM)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void* cudaStream_t;
typedef struct { int dummy; }* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T>
inline cudaError_t cudaMallocHost(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ------------------------------------------------------------------ scheduler (one definition: the .cu is a single TU)
#ifdef B2E_EMU_IMPL
namespace emu {
thread_local Fiber* g_cur = nullptr;
thread_local void* g_sched_sp = nullptr;
thread_local Block g_block;
thread_local uint3 g_blockIdx{0, 0, 0};
thread_local dim3 g_blockDim, g_gridDim;
static thread_local void (*g_body)(void*) = nullptr;
static thread_local void* g_arg = nullptr;
static thread_local std::vector<Fiber>* g_fibers = nullptr;

asm(R"(
.pushsection .text
.globl emu_switch
.hidden emu_switch
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
.popsection
)");

static void fiber_exit() {
  Fiber* f = g_cur;
  Warp& w = *f->warp;
  f->state = 3;
  w.alive--;
  w.alive_mask &= ~(1u << f->lane);
  g_block.alive--;
  // lanes already waiting at a rendezvous may now be complete
  if (w.alive > 0 && w.arrived == w.alive) { w.arrived = 0; w.gen++; }
  if (g_block.alive > 0 && g_block.arrived == g_block.alive) { g_block.arrived = 0; g_block.gen++; }
  yield_to_scheduler();
  die("resumed a finished thread");
}
static void fiber_entry() {
  if (getenv("EMU_TRACE")) fprintf(stderr, "fiber_entry t=%u body=%p arg=%p\n", g_cur->tidx.x, (void*)g_body, g_arg);
  g_body(g_arg);
  fiber_exit();
}

#ifndef EMU_STACK_BYTES
#define EMU_STACK_BYTES (256 * 1024)
#endif

void launch(dim3 grid, dim3 block, size_t smem_bytes, void (*body)(void*), void* arg) {
  (void)smem_bytes;
  const int nt = (int)block.x;
  const int nw = (nt + 31) / 32;
  if (!g_fibers) g_fibers = new std::vector<Fiber>();
  std::vector<Fiber>& fb = *g_fibers;
  if ((int)fb.size() < nt) {
    const size_t old = fb.size();
    fb.resize(nt);
    for (size_t i = old; i < fb.size(); i++) fb[i].stack = (unsigned char*)aligned_alloc(64, EMU_STACK_BYTES);
  }
  std::vector<Warp> warps(nw);
  g_blockDim = block;
  g_gridDim = grid;
  g_body = body;
  g_arg = arg;
  for (unsigned b = 0; b < grid.x; b++) {
    g_blockIdx = uint3{b, 0, 0};
    g_block = Block();
    g_block.alive = nt;
    for (int wi = 0; wi < nw; wi++) warps[wi] = Warp();
    for (int t = 0; t < nt; t++) {
      Fiber& f = fb[t];
      f.tidx = uint3{(unsigned)t, 0, 0};
      f.lane = t & 31;
      f.state = 0;
      f.warp = &warps[t >> 5];
      f.warp->alive++;
      f.warp->alive_mask |= 1u << f.lane;
      uintptr_t top = ((uintptr_t)f.stack + EMU_STACK_BYTES) & ~(uintptr_t)15;
      void** sp = (void**)top;
      *--sp = nullptr;                 // fake return address of fiber_entry (keeps the ABI stack alignment)
      *--sp = (void*)&fiber_entry;     // popped by `ret` in emu_switch
      for (int k = 0; k < 6; k++) *--sp = nullptr;
      f.sp = (void*)sp;
    }
    while (g_block.alive > 0) {
      bool progress = false;
      for (int wi = 0; wi < nw; wi++) {
        Warp& w = warps[wi];
        bool any = true;
        while (any) {
          any = false;
          const int t0 = wi * 32, t1 = std::min(nt, t0 + 32);
          for (int t = t0; t < t1; t++) {
            Fiber& f = fb[t];
            if (f.state == 3) continue;
            if (f.state == 1 && w.gen == f.wait_gen) continue;
            if (f.state == 2 && g_block.gen == f.wait_gen) continue;
            g_cur = &f;
            emu_switch(&g_sched_sp, f.sp);
            any = true;
            progress = true;
          }
        }
      }
      if (!progress) {
        g_cur = nullptr;
        die("deadlock: a collective or barrier was not reached by every live thread");
      }
    }
  }
  g_cur = nullptr;
}
}  // namespace emu
#endif  // B2E_EMU_IMPL

#if defined(B2E_EMU_IMPL) && defined(EMU_SEGV_TRACE)
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
namespace emu {
#include <ucontext.h>
static void segv_handler(int sig, siginfo_t* si, void* uc_) {
  void* bt[64];
  ucontext_t* uc = (ucontext_t*)uc_;
  void* rip = (void*)uc->uc_mcontext.gregs[REG_RIP];
  fprintf(stderr, "rip %p\n", rip);
  backtrace_symbols_fd(&rip, 1, 2);
  fprintf(stderr, "cuda_emu: signal %d at address %p (block %u thread %u)\n", sig, si->si_addr, g_blockIdx.x,
          g_cur ? g_cur->tidx.x : 0u);
  const int n = backtrace(bt, 64);
  backtrace_symbols_fd(bt, n, 2);
  _exit(139);
}
__attribute__((constructor)) static void install_segv() {
  static unsigned char alt[1 << 16];
  stack_t ss;
  ss.ss_sp = alt; ss.ss_size = sizeof(alt); ss.ss_flags = 0;
  sigaltstack(&ss, nullptr);
  struct sigaction sa;
  memset(&sa, 0, sizeof(sa));
  sa.sa_sigaction = segv_handler;
  sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
  sigaction(SIGSEGV, &sa, nullptr);
}
}  // namespace emu
#endif
