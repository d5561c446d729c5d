// cuda_emu.h -- a lockstep-warp CPU emulator of the handful of CUDA features the
// chromo_b200 kernels use.  TEST TOOL ONLY: it lets the kernel LOGIC (hash table,
// warp collectives, proposal code, the C ABI host side) be exercised against the
// oracle in a container without a GPU (tests/test_emu_*.py).  It is never built
// into, shipped with, or loaded by the chromo_b200 package; the product library
// is compiled by nvcc for sm_100a and has no CPU path.
//
// Model: each thread block runs as blockDim.x cooperative FIBERS on the calling OS
// thread (own stacks, a 10-instruction x86-64 context switch); warp collectives
// (__shfl*_sync, __any_sync, __syncwarp) rendezvous on a per-warp barrier,
// __syncthreads on a per-block barrier -- a fiber that has to wait hands the core to
// the next fiber of the block; blocks of a grid run one after another.  (The first
// version ran a block as OS threads on pthread barriers: the same semantics at ~10x
// the wall time, nearly all of it in futex calls.)
// All collectives in the kernels are called with warp-uniform control flow and a
// full mask, which is what this model requires.
#pragma once
#if !defined(__x86_64__)
#error "cuda_emu.h switches fibers with x86-64 assembly"
#endif
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(x) alignas(x)
#define __shared__ static

using std::max;
using std::min;

struct emu_dim3 {
    unsigned x = 1, y = 1, z = 1;
    emu_dim3() {}
    emu_dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
typedef emu_dim3 dim3;
inline thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim; // one instance across TUs

struct EmuBarrier {
    unsigned count = 0, need = 0, gen = 0;
};
struct EmuBlock {
    EmuBarrier block_bar;
    std::vector<EmuBarrier> warp_bar;
    std::vector<uint64_t> slots; // one per thread
    EmuBarrier named[16];        // named barriers (bar.sync id, count)
    unsigned char *dyn = nullptr;
    // fibers
    unsigned nt = 0, cur = 0, live = 0;
    std::vector<void *> sp;          // saved stack pointers
    std::vector<unsigned char> done;
    void *main_sp = nullptr;
    std::function<void()> body;
};
inline thread_local EmuBlock *emu_blk = nullptr;

// save the callee-saved registers on the current stack, park its pointer in *from, continue on `to`
static __attribute__((naked, noinline)) void emu_switch(void **from, void *to) {
    asm volatile(
        "pushq %rbp\n pushq %rbx\n pushq %r12\n pushq %r13\n pushq %r14\n pushq %r15\n"
        "movq %rsp, (%rdi)\n movq %rsi, %rsp\n"
        "popq %r15\n popq %r14\n popq %r13\n popq %r12\n popq %rbx\n popq %rbp\n ret\n");
}
// hand the core to the next unfinished fiber of the block (round robin)
static inline void emu_yield() {
    EmuBlock *b = emu_blk;
    const unsigned me = b->cur;
    unsigned n = me;
    do n = n + 1 == b->nt ? 0 : n + 1; while (b->done[n]);
    if (n == me) return;
    b->cur = n;
    emu_switch(&b->sp[me], b->sp[n]);
    threadIdx.x = me; // back on this fiber
}
static inline void emu_barrier(EmuBarrier &bar) {
    const unsigned gen = bar.gen;
    if (++bar.count == bar.need) {
        bar.count = 0;
        bar.gen = gen + 1;
    } else {
        while (*(volatile unsigned *)&bar.gen == gen) emu_yield();
    }
}
static __attribute__((noinline)) void emu_fiber_entry() {
    EmuBlock *b = emu_blk;
    const unsigned me = b->cur;
    threadIdx.x = me;
    b->body();
    b->done[me] = 1;
    void *dummy;
    if (--b->live == 0) emu_switch(&dummy, b->main_sp);
    unsigned n = me;
    do n = n + 1 == b->nt ? 0 : n + 1; while (b->done[n]);
    b->cur = n;
    emu_switch(&dummy, b->sp[n]);
    abort(); // a finished fiber is never resumed
}

static inline void emu_warp_wait() { emu_barrier(emu_blk->warp_bar[threadIdx.x >> 5]); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp_wait(); }
static inline void __syncthreads() { emu_barrier(emu_blk->block_bar); }
static inline void cb_bar_sync(int id, int count) {
    emu_blk->named[id].need = (unsigned)count;
    emu_barrier(emu_blk->named[id]);
}

template <class T>
static inline T emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "");
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    unsigned base = threadIdx.x & ~31u;
    emu_blk->slots[threadIdx.x] = bits;
    emu_warp_wait();
    uint64_t got = emu_blk->slots[base + (unsigned)(src_lane & 31)];
    emu_warp_wait();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (int)((threadIdx.x & 31) ^ (unsigned)m)); }
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int d) {
    int src = (int)(threadIdx.x & 31) + d;
    return emu_exchange(v, src > 31 ? (int)(threadIdx.x & 31) : src);
}
static inline int __any_sync(unsigned, int pred) {
    unsigned base = threadIdx.x & ~31u;
    emu_blk->slots[threadIdx.x] = pred ? 1 : 0;
    emu_warp_wait();
    int r = 0;
    for (int i = 0; i < 32 && base + i < blockDim.x; i++) r |= (int)emu_blk->slots[base + i];
    emu_warp_wait();
    return r;
}

static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned base = threadIdx.x & ~31u;
    emu_blk->slots[threadIdx.x] = pred ? 1 : 0;
    emu_warp_wait();
    unsigned r = 0;
    for (int i = 0; i < 32 && base + i < blockDim.x; i++) r |= (unsigned)(emu_blk->slots[base + i] & 1) << i;
    emu_warp_wait();
    return r;
}
// position of the `offset`-th set bit of `mask` at or above `base` (offset >= 1), 0xffffffff if there is none
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    for (unsigned b = base; b < 32; b++)
        if ((mask >> b) & 1u)
            if (--offset == 0) return b;
    return 0xffffffffu;
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }

static inline int atomicCAS(int *a, int cmp, int val) {
    __atomic_compare_exchange_n(a, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
static inline int atomicAdd(int *a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline double atomicAdd(double *a, double v) {
    uint64_t *p = (uint64_t *)a, old = __atomic_load_n(p, __ATOMIC_SEQ_CST), nw;
    double o;
    do {
        memcpy(&o, &old, 8);
        double n = o + v;
        memcpy(&nw, &n, 8);
    } while (!__atomic_compare_exchange_n(p, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST));
    return o;
}
static inline unsigned atomicAdd(unsigned *a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *a, unsigned long long v) {
    return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST);
}
static inline long long __double2ll_rn(double x) { return llrint(x); } // default rounding mode: to nearest even
static inline double __hiloint2double(int hi, int lo) {
    uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
// CUDA's sincospi (production-mode sphere points); the emulator only has to agree to rounding
static inline void sincospi(double x, double *s, double *c) {
    *s = sin(3.14159265358979323846 * x);
    *c = cos(3.14159265358979323846 * x);
}
struct double2 { double x, y; };
struct double3 { double x, y, z; };
struct int2 { int x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline double3 make_double3(double x, double y, double z) { return double3{x, y, z}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
#define CB_NOINLINE __attribute__((noinline))
#define CB_GRID_CONSTANT
static inline void cb_prefetch(const void *) {}
typedef uintptr_t cb_saddr;
static inline cb_saddr cb_shared_addr(const void *p) { return (cb_saddr)p; }
static inline void cb_red_add_u32(cb_saddr a, uint32_t v) { __atomic_fetch_add((uint32_t *)a, v, __ATOMIC_SEQ_CST); }
static inline void cb_backoff() { emu_yield(); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline double __longlong_as_double(long long v) {
    double d;
    memcpy(&d, &v, 8);
    return d;
}

#define CB_DYN_SMEM(name) unsigned char *name = emu_blk->dyn

// ---- minimal CUDA runtime ------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp {
    int multiProcessorCount;
    size_t sharedMemPerBlockOptin;
    size_t sharedMemPerMultiprocessor;
};
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    p->multiProcessorCount = 148;
    p->sharedMemPerBlockOptin = 227 * 1024;
    p->sharedMemPerMultiprocessor = 228 * 1024;
    return 0;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, int) { *s = (void *)1; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
typedef void *cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, int) { *e = (void *)1; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(1, n); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) {
    memcpy(d, s, n);
    return 0;
}
static inline cudaError_t cudaGetLastError() { return 0; }
enum { cudaHostRegisterDefault = 0, cudaErrorHostMemoryAlreadyRegistered = 712 };
static inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return 0; }
static inline cudaError_t cudaHostUnregister(void *) { return 0; }
template <class K>
static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return 0; }

// ---- kernel launch ---------------------------------------------------------
struct EmuStacks { // one region per OS thread, reused by every launch
    unsigned char *base = nullptr;
    size_t each = 0, n = 0;
    ~EmuStacks() {
        if (base) munmap(base, each * n);
    }
    unsigned char *top(unsigned t) { return base + each * (t + 1); }
    void reserve(size_t threads) {
        if (threads <= n) return;
        if (base) munmap(base, each * n);
        each = (size_t)1 << 20; // 1 MiB of address space per fiber, touched pages only are backed
        n = threads;
        void *p = mmap(nullptr, each * n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) {
            perror("cuda_emu: mmap of the fiber stacks");
            abort();
        }
        base = (unsigned char *)p;
    }
};
inline thread_local EmuStacks emu_stacks;

template <class K, class... A>
static void emu_launch(K kernel, dim3 grid, dim3 block, size_t smem, A... args) {
    const unsigned nt = block.x;
    const unsigned nw = (nt + 31) / 32;
    EmuBlock blk;
    EmuBlock *outer = emu_blk;
    const emu_dim3 o_t = threadIdx, o_b = blockIdx, o_bd = blockDim, o_gd = gridDim;
    blk.slots.assign(nw * 32, 0);
    blk.warp_bar.resize(nw);
    std::vector<unsigned char> dyn(smem + 64);
    blk.dyn = (unsigned char *)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
    blk.nt = nt;
    blk.sp.resize(nt);
    blk.done.resize(nt);
    blk.body = [&]() { kernel(args...); };
    emu_stacks.reserve(nt);
    emu_blk = &blk;
    blockDim = block;
    gridDim = grid;
    for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++) {
            blk.block_bar = EmuBarrier{0, nt, 0};
            for (unsigned w = 0; w < nw; w++) blk.warp_bar[w] = EmuBarrier{0, std::min(32u, nt - 32 * w), 0};
            for (auto &b : blk.named) b = EmuBarrier{};
            blockIdx = emu_dim3(bx, by);
            for (unsigned t = 0; t < nt; t++) { // a fresh stack whose first `ret` enters emu_fiber_entry
                void **top = (void **)((uintptr_t)emu_stacks.top(t) & ~(uintptr_t)15);
                top[-1] = nullptr;                  // return address of the entry function (never used)
                top[-2] = (void *)&emu_fiber_entry; // popped by emu_switch's ret
                for (int k = 3; k <= 8; k++) top[-k] = nullptr; // rbp rbx r12 r13 r14 r15
                blk.sp[t] = (void *)(top - 8);
                blk.done[t] = 0;
            }
            blk.live = nt;
            blk.cur = 0;
            emu_switch(&blk.main_sp, blk.sp[0]); // returns when the last fiber has finished
        }
    emu_blk = outer;
    threadIdx = o_t, blockIdx = o_b, blockDim = o_bd, gridDim = o_gd;
}
#define CB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu_launch(kernel, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)
