// TEST-ONLY: a tiny CUDA-on-CPU emulator (one fibre per CUDA thread, CTAs run one after the
// other) so that the logic of the kernels in cavif_rs_b200/csrc/*.cu can be exercised on the
// GPU-less development container, under `-m "not gpu"` tests, before spending GPU time.
// It is NOT a fallback: nothing in the product library includes this header, and the C-ABI
// fails loudly when no CUDA device is present.  Kernels are compiled as plain C++ with
// -DB200_EMU; only the subset of CUDA used by this repo is provided (full-mask warp
// collectives, __syncthreads, static __shared__, threadIdx/blockIdx/blockDim/gridDim.x).
#ifndef CUDA_EMU_H_
#define CUDA_EMU_H_
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <algorithm>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __constant__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct emu_dim3 { unsigned x, y, z; };
namespace emu {
struct Fiber {
  ucontext_t ctx;
  char *stack;
  bool done;
  unsigned tid;
};
extern thread_local emu_dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
extern thread_local ucontext_t g_sched;
extern thread_local Fiber *g_cur;
extern thread_local unsigned g_cta_arrived, g_cta_gen;
extern thread_local unsigned g_warp_arrived[64], g_warp_gen[64];
extern thread_local uint64_t g_warp_buf[64][2][32];
void yield();
void launch(void (*fn)(void *), void *arg, unsigned grid, unsigned block);
}  // namespace emu

#define threadIdx emu::g_threadIdx
#define blockIdx emu::g_blockIdx
#define blockDim emu::g_blockDim
#define gridDim emu::g_gridDim

static inline void __syncthreads() {
  const unsigned gen = emu::g_cta_gen;
  if (++emu::g_cta_arrived == emu::g_blockDim.x) { emu::g_cta_arrived = 0; emu::g_cta_gen++; }
  while (emu::g_cta_gen == gen) emu::yield();
}
static inline void emu_warp_barrier() {
  const unsigned w = emu::g_threadIdx.x >> 5;
  const unsigned lanes = std::min(32u, emu::g_blockDim.x - w * 32);
  const unsigned gen = emu::g_warp_gen[w];
  if (++emu::g_warp_arrived[w] == lanes) { emu::g_warp_arrived[w] = 0; emu::g_warp_gen[w]++; }
  while (emu::g_warp_gen[w] == gen) emu::yield();
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <typename T>
static inline T emu_exchange(T v, int src_lane) {
  const unsigned w = emu::g_threadIdx.x >> 5, lane = emu::g_threadIdx.x & 31;
  const unsigned slot = emu::g_warp_gen[w] & 1;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  emu::g_warp_buf[w][slot][lane] = raw;
  emu_warp_barrier();
  T out;
  memcpy(&out, &emu::g_warp_buf[w][slot][src_lane & 31], sizeof(T));
  return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (int)((emu::g_threadIdx.x & 31) ^ m)); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  const int lane = emu::g_threadIdx.x & 31;
  return emu_exchange(v, lane + d < 32 ? lane + d : lane);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned r = 0;
  for (int l = 0; l < 32; l++) r |= (emu_exchange<int>(pred ? 1 : 0, l) ? 1u : 0u) << l;
  return r;
}
static inline unsigned __reduce_min_sync(unsigned, unsigned v) {
  unsigned r = v;
  for (int m = 16; m > 0; m >>= 1) { const unsigned o = emu_exchange(r, (int)((emu::g_threadIdx.x & 31) ^ m)); r = o < r ? o : r; }
  return r;
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
  unsigned r = v;
  for (int m = 16; m > 0; m >>= 1) r += emu_exchange(r, (int)((emu::g_threadIdx.x & 31) ^ m));
  return r;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
using std::max;
using std::min;

#define EMU_LAUNCH(kernel_wrapper, argptr, grid, block) emu::launch(kernel_wrapper, argptr, grid, block)

#ifdef CUDA_EMU_IMPLEMENTATION
namespace emu {
thread_local emu_dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
thread_local ucontext_t g_sched;
thread_local Fiber *g_cur;
thread_local unsigned g_cta_arrived, g_cta_gen;
thread_local unsigned g_warp_arrived[64], g_warp_gen[64];
thread_local uint64_t g_warp_buf[64][2][32];
static thread_local void (*g_fn)(void *);
static thread_local void *g_arg;

void yield() { swapcontext(&g_cur->ctx, &g_sched); }
static void trampoline() {
  g_fn(g_arg);
  g_cur->done = true;
  swapcontext(&g_cur->ctx, &g_sched);
}
void launch(void (*fn)(void *), void *arg, unsigned grid, unsigned block) {
  const size_t stack_sz = 256 * 1024;
  std::vector<Fiber> fibers(block);
  for (auto &f : fibers) f.stack = (char *)malloc(stack_sz);
  g_fn = fn; g_arg = arg;
  g_blockDim = {block, 1, 1};
  g_gridDim = {grid, 1, 1};
  for (unsigned b = 0; b < grid; b++) {
    g_blockIdx = {b, 0, 0};
    g_cta_arrived = 0; g_cta_gen = 0;
    memset(g_warp_arrived, 0, sizeof(g_warp_arrived));
    memset(g_warp_gen, 0, sizeof(g_warp_gen));
    for (unsigned t = 0; t < block; t++) {
      Fiber &f = fibers[t];
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack;
      f.ctx.uc_stack.ss_size = stack_sz;
      f.ctx.uc_link = nullptr;
      f.done = false; f.tid = t;
      makecontext(&f.ctx, trampoline, 0);
    }
    unsigned remaining = block;
    while (remaining) {
      for (unsigned t = 0; t < block; t++) {
        Fiber &f = fibers[t];
        if (f.done) continue;
        g_cur = &f;
        g_threadIdx = {t, 0, 0};
        swapcontext(&g_sched, &f.ctx);
        if (f.done) remaining--;
      }
    }
  }
  for (auto &f : fibers) free(f.stack);
}
}  // namespace emu
#endif  // CUDA_EMU_IMPLEMENTATION
#endif  // CUDA_EMU_H_
