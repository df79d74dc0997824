// TEST / DEBUG INFRASTRUCTURE ONLY — never compiled into the product library.
//
// A tiny lock-step SIMT emulator: lets the device code in nvidia-texture-tools_b200/csrc/kernels/*.cuh be
// compiled by plain g++ and executed on the CPU (one ucontext fiber per CUDA thread, warps synchronise at
// __shfl/__ballot/__syncwarp/__syncthreads).  The development container has no GPU, so this is how kernel
// logic is debugged against the oracle before a gpurun call.  Because the kernels are written without FMA
// contraction and only use IEEE +,-,*,/,sqrt, emulated results are bit-identical to the GPU's.
// It is NOT a CPU fallback: nothing in the product (C-ABI, host library, bench) can reach it.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <ucontext.h>
#include <functional>
#include <vector>
#include <thread>
#include <atomic>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __constant__
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))
#define NVB_EMU 1

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int3 { int x, y, z; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct ushort4 { unsigned short x, y, z, w; };
struct ushort2 { unsigned short x, y; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float3 make_float3(float a, float b, float c) { return {a, b, c}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline int2 make_int2(int a, int b) { return {a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return {a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }
static inline uchar4 make_uchar4(unsigned char a, unsigned char b, unsigned char c, unsigned char d) { return {a, b, c, d}; }

namespace emu {
struct Warp;
struct Fiber {
    ucontext_t ctx;
    uint3 tid;
    int lane;
    Warp *warp;
    bool done;
    char *stack;
};
struct Bar { uint32_t mask; int arrived; uint32_t gen; };
struct Warp {
    Bar bars[8];
    uint64_t xchg[32];
};
struct Cta {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    ucontext_t sched;
    int nthreads;
    int sync_arrived;
    uint32_t sync_gen;
    int live;
    uint64_t progress;
};
extern thread_local Fiber *cur;
extern thread_local Cta *cta;
extern thread_local uint3 t_blockIdx;
extern thread_local dim3 t_blockDim, t_gridDim;
extern thread_local char *t_dyn_smem;

void yield();
void warp_barrier(uint32_t mask);
void cta_barrier();
// Runs body() once per CUDA thread of a grid x block launch (1-D or 2-D), CTAs spread over host threads.
void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body);
}  // namespace emu

#define threadIdx (emu::cur->tid)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)
#define warpSize 32

static inline void __syncthreads() { emu::cta_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_barrier(mask); }

template <class T> static inline T emu_xchg(unsigned mask, T v, int srcLane) {
    static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes");
    emu::Warp *w = emu::cur->warp;
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    w->xchg[emu::cur->lane] = raw;
    emu::warp_barrier(mask);
    if (!((mask >> srcLane) & 1)) srcLane = emu::cur->lane;  // undefined on HW; keep own value
    raw = w->xchg[srcLane];
    emu::warp_barrier(mask);
    T r;
    memcpy(&r, &raw, sizeof(T));
    return r;
}
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    int lane = emu::cur->lane;
    int base = lane & ~(width - 1);
    return emu_xchg(mask, v, base + (src & (width - 1)));
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int laneMask, int width = 32) {
    int lane = emu::cur->lane;
    int s = lane ^ laneMask;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
    return emu_xchg(mask, v, s);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int lane = emu::cur->lane;
    int s = lane + (int)delta;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
    return emu_xchg(mask, v, s);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    int lane = emu::cur->lane;
    int s = lane - (int)delta;
    if (s < 0 || (s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
    return emu_xchg(mask, v, s);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu::Warp *w = emu::cur->warp;
    w->xchg[emu::cur->lane] = pred ? 1 : 0;
    emu::warp_barrier(mask);
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (((mask >> i) & 1) && w->xchg[i]) r |= 1u << i;
    emu::warp_barrier(mask);
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline unsigned __activemask() { return 0xffffffffu; }

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1) << (31 - i);
    return r;
}
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmaf_rn(float a, float b, float c) { return __builtin_fmaf(a, b, c); }
static inline float __saturatef(float x) { return (x > 0.0f) ? ((x < 1.0f) ? x : 1.0f) : 0.0f; }  // NaN -> 0
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __uint2float_rn(unsigned a) { return (float)a; }
static inline int __float2int_rz(float a) { return (int)a; }
static inline unsigned __float2uint_rz(float a) { return (unsigned)a; }
static inline int __float2int_rn(float a) { return (int)lrintf(a); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned b = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) b = (b & 0x80) ? 0xff : 0;
        r |= b << (8 * i);
    }
    return r;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
template <class T> static inline T atomicAdd(T *p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
