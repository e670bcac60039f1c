// TEST INFRASTRUCTURE: just enough of the CUDA C++ surface to compile csrc/device/*.cuh with g++ for ONE
// thread, so that the device algorithms can be checked on the CPU against the golden vectors
// (tests/test_hostdev.py).  Nothing in the product includes this; the product path is nvcc + a GPU.
#pragma once
// every standard header the including files use comes FIRST: libstdc++ spells __attribute__((__noinline__)), which the
// __noinline__ macro below would mangle if those headers were read after it
#include <sched.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

// the host builds pin expression order and control flow bit for bit against the reference: IEEE division / square root
// in the ORCA arithmetic too (the GPU build uses the SFU approximations there, device/geom.cuh)
#define ECM_ORCA_IEEE 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

using std::max;
using std::min;

template <class T>
static inline T __ldg(const T* p) { return *p; }
template <class T>
static inline T __ldcg(const T* p) { return *p; }

#ifndef HD_SIMT
// one thread == one "warp": the warp-collective operations degenerate to identities
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
static inline int __reduce_max_sync(unsigned, int v) { return v; }
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
template <class T>
static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
#endif
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { sched_yield(); }  // a spinning lane lets the other ranks' threads run

template <class T, class U>
static inline T atomicAdd(T* p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }
static inline int atomicMax(int* p, int v) { int old = *p; if (v > old) *p = v; return old; }
template <class T>
static inline T atomicCAS(T* p, T cmp, T val) { T old = *p; if (old == cmp) *p = val; return old; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }

#ifdef HD_SIMT
#include "simt.h"  // real blocks of co-operating threads (fibers) with working collectives
#else
// launch geometry: one thread per block (kernels_emul.cpp walks blockIdx.x over the grid)
struct HdDim3 { int x, y, z; };
static HdDim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
#endif
#ifdef HD_SIMT
#define __shared__ static thread_local  // CTAs run one after another per OS thread; ranks on other threads have their own
#else
#define __shared__ static
#endif
#ifdef HD_MOCK_RUNTIME
#include "mock_runtime.h"  // tests/hostdev/mock_cuda: the host side of the runtime API, synchronous (needs HD_SIMT)
#endif

// RotateVector / the VO half angle use the C library's sinf / cosf (UtilityFunctions.cpp:233-242) and so does
// the oracle: route the device code's single sincosf call to the same two functions
static inline void hd_sincosf(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }
#define sincosf hd_sincosf
