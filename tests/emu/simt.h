// simt.h -- a small SIMT emulator so that the product's .cu sources can be compiled with g++ and
// EXECUTED ON THE CPU BY THE TEST SUITE ONLY (there is no GPU in the authoring container).
//
// TEST INFRASTRUCTURE ONLY.  This header is force-included (g++ -x c++ -include simt.h
// -DRTK_SIMT_EMU) when tests/emu/build_emu.py compiles rtk_b200/csrc/*.cu into
// tests/emu/librtk_emu.so.  The product library librtk_b200.so is compiled by nvcc for sm_100a
// and contains none of this; rtk_b200/api.py never loads the emulated library.
//
// Model: every CUDA thread of a block is a fiber (hand-rolled x86-64 context switch).  Blocks
// run one after another; inside a block, warps are scheduled round-robin and a fiber yields
// whenever it reaches a warp collective (__shfl_*_sync, __ballot_sync, __match_any_sync,
// __syncwarp ...) or __syncthreads().  Atomics are plain operations (one OS thread).  Floating
// point follows IEEE with -ffp-contract=off, fmaf() is a true fused multiply-add, so device
// arithmetic is reproduced bit-for-bit.
#pragma once
#ifndef RTK_SIMT_EMU
#define RTK_SIMT_EMU 1
#endif

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <functional>
#include <mutex>
#include <vector>

// ------------------------------------------------------------------------------------------
// qualifiers
// ------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

// ------------------------------------------------------------------------------------------
// vector types
// ------------------------------------------------------------------------------------------
struct uint3_ { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct float3 { float x, y, z; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct __attribute__((aligned(8))) int2 { int x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline float3 make_float3(float x, float y, float z) { float3 r = {x, y, z}; return r; }
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = {x, y, z, w}; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = {x, y}; return r; }
static inline int2 make_int2(int x, int y) { int2 r = {x, y}; return r; }
static inline int4 make_int4(int x, int y, int z, int w) { int4 r = {x, y, z, w}; return r; }

// ------------------------------------------------------------------------------------------
// fibers
// ------------------------------------------------------------------------------------------
namespace simt {

struct Warp;
struct Fiber {
	void *sp;
	char *stack;
	uint3_ tid;
	int lane, warp;
	bool done, at_block_barrier;
	Warp *w;
};
struct Warp {
	unsigned gen, count, live_mask;
	uint64_t slot[32];
	uint64_t slot2[32];
};
struct Block {
	unsigned gen, count, live;
};

extern Fiber *cur;
extern Block blk;
extern void *sched_sp;
extern uint3_ g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern std::function<void()> *g_body;
extern unsigned long long g_switches, g_progress;

extern "C" void simt_switch(void **save_sp, void *load_sp);

static inline void yield() { g_switches++; simt_switch(&cur->sp, sched_sp); }

void launch(dim3 grid, dim3 block, const std::function<void()> &body);

// all lanes named in mask rendezvous here
static inline void warp_barrier(unsigned mask)
{
	Warp *w = cur->w;
	unsigned expect = (unsigned)__builtin_popcount(mask & w->live_mask);
	unsigned gen = w->gen;
	g_progress++;
	if (++w->count >= expect) { w->count = 0; w->gen++; }
	else while (w->gen == gen) yield();
}

} // namespace simt

#define threadIdx (simt::cur->tid)
#define blockIdx (simt::g_blockIdx)
#define blockDim (simt::g_blockDim)
#define gridDim (simt::g_gridDim)
#define warpSize 32

static inline void __syncthreads()
{
	simt::Fiber *f = simt::cur;
	unsigned gen = simt::blk.gen;
	f->at_block_barrier = true;
	simt::blk.count++;
	simt::g_progress++;
	while (simt::blk.gen == gen) simt::yield();
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::warp_barrier(mask); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <typename T> static inline T simt_exchange(unsigned mask, T v, int src_lane)
{
	static_assert(sizeof(T) <= 8, "shuffle payload");
	simt::Warp *w = simt::cur->w;
	uint64_t raw = 0;
	memcpy(&raw, &v, sizeof(T));
	w->slot[simt::cur->lane] = raw;
	simt::warp_barrier(mask);
	uint64_t got = w->slot[src_lane & 31];
	simt::warp_barrier(mask);
	T r;
	memcpy(&r, &got, sizeof(T));
	return r;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
	int lane = simt::cur->lane;
	int base = lane & ~(width - 1);
	return simt_exchange(mask, v, base + (src & (width - 1)));
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32)
{
	int lane = simt::cur->lane;
	int src = lane ^ lanemask;
	if ((src & ~(width - 1)) != (lane & ~(width - 1))) src = lane;
	return simt_exchange(mask, v, src);
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
	int lane = simt::cur->lane;
	int src = lane - (int)delta;
	if (src < (lane & ~(width - 1))) src = lane;
	return simt_exchange(mask, v, src);
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
	int lane = simt::cur->lane;
	int src = lane + (int)delta;
	if (src > (lane | (width - 1))) src = lane;
	return simt_exchange(mask, v, src);
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
	simt::Warp *w = simt::cur->w;
	w->slot[simt::cur->lane] = pred ? 1 : 0;
	simt::warp_barrier(mask);
	unsigned r = 0;
	for (int i = 0; i < 32; i++) if (((mask & w->live_mask) >> i & 1) && w->slot[i]) r |= 1u << i;
	simt::warp_barrier(mask);
	return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __match_any_sync(unsigned mask, unsigned v)
{
	simt::Warp *w = simt::cur->w;
	w->slot[simt::cur->lane] = v;
	simt::warp_barrier(mask);
	unsigned r = 0;
	for (int i = 0; i < 32; i++) if (((mask & w->live_mask) >> i & 1) && w->slot[i] == (uint64_t)v) r |= 1u << i;
	simt::warp_barrier(mask);
	return r;
}
static inline unsigned simt_reduce(unsigned mask, unsigned v, int op)
{
	simt::Warp *w = simt::cur->w;
	w->slot[simt::cur->lane] = v;
	simt::warp_barrier(mask);
	unsigned r = op == 0 ? 0xffffffffu : 0u;
	for (int i = 0; i < 32; i++) if ((mask & w->live_mask) >> i & 1) {
		unsigned x = (unsigned)w->slot[i];
		r = op == 0 ? (x < r ? x : r) : (op == 1 ? (x > r ? x : r) : r + x);
	}
	simt::warp_barrier(mask);
	return r;
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return simt_reduce(mask, v, 0); }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return simt_reduce(mask, v, 1); }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return simt_reduce(mask, v, 2); }
static inline unsigned __activemask() { return simt::cur->w->live_mask; }

// ------------------------------------------------------------------------------------------
// intrinsics
// ------------------------------------------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline int __clzll(long long v) { return v == 0 ? 64 : __builtin_clzll((unsigned long long)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __brev(unsigned v)
{
	v = (v >> 16) | (v << 16);
	v = ((v & 0xff00ff00u) >> 8) | ((v & 0x00ff00ffu) << 8);
	v = ((v & 0xf0f0f0f0u) >> 4) | ((v & 0x0f0f0f0fu) << 4);
	v = ((v & 0xccccccccu) >> 2) | ((v & 0x33333333u) << 2);
	v = ((v & 0xaaaaaaaau) >> 1) | ((v & 0x55555555u) << 1);
	return v;
}
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __uint2float_rn(unsigned a) { return (float)a; }
static inline int __float2int_rd(float a) { return (int)floorf(a); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
template <typename T> static inline void __stcg(T *p, T v) { *p = v; }
template <typename T> static inline void __stcs(T *p, T v) { *p = v; }
static inline float __saturatef(float x) { return x < 0 ? 0 : x > 1 ? 1 : x; }
static inline unsigned umin(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned umax(unsigned a, unsigned b) { return a > b ? a : b; }
#ifndef __CUDACC__
template <typename T> static inline T min(T a, T b) { return a < b ? a : b; }
template <typename T> static inline T max(T a, T b) { return a > b ? a : b; }
#endif

template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

// ------------------------------------------------------------------------------------------
// a fake CUDA runtime: device memory is host memory
// ------------------------------------------------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef struct simt_event { double t; } *cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
struct cudaDeviceProp { char name[256]; int multiProcessorCount; int l2CacheSize; size_t totalGlobalMem; int major, minor; };
static inline const char *cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
// SIMT_DEVICES=n in the environment emulates n devices (the in-library multi-GPU layer): the current
// device is per host thread as in CUDA, "device memory" remembers the device it was allocated on
// (cudaPointerGetAttributes reports it), peer copies are plain copies.
static inline int simt_device_count() { static const int n = getenv("SIMT_DEVICES") && atoi(getenv("SIMT_DEVICES")) > 0 ? atoi(getenv("SIMT_DEVICES")) : 1; return n; }
extern __thread int simt_cur_device;
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = simt_device_count(); return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= simt_device_count()) return cudaErrorInvalidValue; simt_cur_device = d; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = simt_cur_device; return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
void simt_track_alloc(void *p, size_t n, int kind);      // kind: 2 device, 1 pinned host
void simt_track_free(void *p);
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int)
{
	memset(p, 0, sizeof(*p));
	strcpy(p->name, "SIMT-EMU (CPU, tests only)");
	p->multiProcessorCount = 2; p->l2CacheSize = 1 << 20; p->totalGlobalMem = (size_t)8 << 30; p->major = 10;
	return cudaSuccess;
}
// With SIMT_SHM_MALLOC=1 in the environment "device memory" comes from named POSIX shared memory, so
// that another emulated process can map it through the fake CUDA IPC calls below (the two-process
// dry run of the peer-memory gather).  Otherwise it is plain heap memory.
struct simt_shm_block { void *ptr; size_t size; char name[40]; };
simt_shm_block *simt_shm_find(const void *p);
void *simt_shm_alloc(size_t n);
bool simt_shm_release(void *p);
template <typename T> static inline cudaError_t cudaMalloc(T **p, size_t n)
{
	void *q = NULL;
	static const bool shm = getenv("SIMT_SHM_MALLOC") && atoi(getenv("SIMT_SHM_MALLOC")) != 0;
	if (shm) { q = simt_shm_alloc(n ? n : 256); if (!q) return cudaErrorMemoryAllocation; }
	else if (posix_memalign(&q, 256, n ? n : 256)) return cudaErrorMemoryAllocation;
	memset(q, 0xCD, n);   // poison: device memory is uninitialised
	simt_track_alloc(q, n ? n : 256, 2);
	*p = (T*)q;
	return cudaSuccess;
}
template <typename T> static inline cudaError_t cudaMallocHost(T **p, size_t n)
{
	void *q = NULL;
	if (posix_memalign(&q, 256, n ? n : 256)) return cudaErrorMemoryAllocation;
	memset(q, 0xCD, n);
	simt_track_alloc(q, n ? n : 256, 1);
	*p = (T*)q;
	return cudaSuccess;
}
#define cudaHostAllocDefault 0
#define cudaHostAllocPortable 1
#define cudaHostAllocMapped 2
template <typename T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
static inline cudaError_t cudaFree(void *p) { if (!p) return cudaSuccess; simt_track_free(p); if (!simt_shm_release(p)) free(p); return cudaSuccess; }
template <typename T> static inline cudaError_t cudaMallocAsync(T **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { return cudaFree(p); }
#define cudaHostRegisterDefault 0
#define cudaHostRegisterPortable 1
#define cudaHostRegisterMapped 2
static inline cudaError_t cudaHostRegister(void *p, size_t n, unsigned) { simt_track_alloc(p, n, 1); return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void *p) { simt_track_free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { if (!p) return cudaSuccess; simt_track_free(p); free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyPeerAsync(void *d, int, const void *s, int, size_t n, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = NULL; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = NULL; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
#define cudaStreamNonBlocking 1
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)malloc(sizeof(simt_event)); return cudaSuccess; }
#define cudaEventDisableTiming 2
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0)
{
	struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
	e->t = ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
	return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <typename F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 2; return cudaSuccess; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
#define cudaFuncAttributeMaxDynamicSharedMemorySize 8
#define cudaFuncAttributePreferredSharedMemoryCarveout 9
struct cudaPointerAttributes { int type; int device; void *devicePointer; void *hostPointer; };
#define cudaMemoryTypeUnregistered 0
#define cudaMemoryTypeHost 1
#define cudaMemoryTypeDevice 2
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p);
// L2 persistence controls: accepted and ignored
#define cudaLimitPersistingL2CacheSize 6
static inline cudaError_t cudaDeviceSetLimit(int, size_t) { return cudaSuccess; }
#define cudaDevAttrMaxPersistingL2CacheSize 108
#define cudaDevAttrMaxAccessPolicyWindowSize 109
static inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 0; return cudaSuccess; }

// CUDA IPC.  Heap-backed memory: the emulated "other process" is this process and a handle is the
// pointer itself.  Shared-memory-backed memory (SIMT_SHM_MALLOC): a handle carries the name and size
// of the block and another process maps it.
struct cudaIpcMemHandle_t { char reserved[64]; };
#define cudaIpcMemLazyEnablePeerAccess 1
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p);
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void *p);

// kernel launch: RTK_LAUNCH(kernel, grid, block, stream, args...)
#define RTK_LAUNCH(kernel, grid, block, stream, ...) \
	simt::launch(dim3(grid), dim3(block), [&]() { kernel(__VA_ARGS__); })

// ------------------------------------------------------------------------------------------
// implementation (one translation unit defines SIMT_IMPL)
// ------------------------------------------------------------------------------------------
#ifdef SIMT_IMPL
namespace simt {
Fiber *cur = NULL;
Block blk;
void *sched_sp = NULL;
uint3_ g_blockIdx;
dim3 g_blockDim, g_gridDim;
std::function<void()> *g_body = NULL;
unsigned long long g_switches = 0, g_progress = 0;

asm(R"(
.text
.globl simt_switch
.type simt_switch,@function
simt_switch:
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
.size simt_switch,.-simt_switch
)");

static void fiber_entry()
{
	(*g_body)();
	cur->done = true;
	for (;;) simt_switch(&cur->sp, sched_sp);
}

static const size_t STACK = 256 * 1024;
static std::vector<Fiber> fibers;
static std::vector<Warp> warps;

static void prepare(Fiber *f)
{
	if (!f->stack) f->stack = (char*)malloc(STACK);
	uintptr_t top = ((uintptr_t)f->stack + STACK) & ~(uintptr_t)15;
	void **sp = (void**)(top - 64);
	for (int i = 0; i < 6; i++) sp[i] = 0;
	sp[6] = (void*)&fiber_entry;
	sp[7] = 0;
	f->sp = sp;
	f->done = false;
	f->at_block_barrier = false;
}

// One kernel at a time: the scheduler state, the fiber pool and the `static` shared memory of the
// kernels are process-wide.  Host threads that launch concurrently (tests of the thread-safe entry
// points) are serialised here, which is a legal schedule for independent streams.
static std::mutex g_launch_lock;

void launch(dim3 grid, dim3 block, const std::function<void()> &body)
{
	std::lock_guard<std::mutex> guard(g_launch_lock);
	std::function<void()> b = body;
	std::function<void()> *saved_body = g_body;
	g_body = &b;
	g_blockDim = block; g_gridDim = grid;
	unsigned nthreads = block.x * block.y * block.z;
	unsigned nwarps = (nthreads + 31) / 32;
	if (fibers.size() < nthreads) { fibers.resize(nthreads); }
	if (warps.size() < nwarps) warps.resize(nwarps);
	for (unsigned bz = 0; bz < grid.z; bz++)
	for (unsigned by = 0; by < grid.y; by++)
	for (unsigned bx = 0; bx < grid.x; bx++) {
		g_blockIdx.x = bx; g_blockIdx.y = by; g_blockIdx.z = bz;
		blk.gen = 0; blk.count = 0; blk.live = nthreads;
		for (unsigned w = 0; w < nwarps; w++) {
			warps[w].gen = 0; warps[w].count = 0;
			unsigned n = nthreads - w * 32 < 32 ? nthreads - w * 32 : 32;
			warps[w].live_mask = n == 32 ? 0xffffffffu : ((1u << n) - 1);
		}
		for (unsigned t = 0; t < nthreads; t++) {
			Fiber *f = &fibers[t];
			prepare(f);
			f->tid.x = t % block.x; f->tid.y = (t / block.x) % block.y; f->tid.z = t / (block.x * block.y);
			f->lane = t & 31; f->warp = t >> 5; f->w = &warps[t >> 5];
		}
		unsigned done = 0;
		while (done < nthreads) {
			unsigned long long p0 = g_progress;
			for (unsigned w = 0; w < nwarps; w++) {
				// run this warp until no lane can move (all finished, parked at __syncthreads,
				// or polling a warp barrier that cannot complete yet)
				unsigned base = w * 32, end = base + 32 < nthreads ? base + 32 : nthreads;
				for (;;) {
					unsigned long long p1 = g_progress;
					for (unsigned t = base; t < end; t++) {
						Fiber *f = &fibers[t];
						if (f->done || f->at_block_barrier) continue;
						cur = f;
						simt_switch(&sched_sp, f->sp);
						if (f->done) {
							done++;
							blk.live--;
							g_progress++;
							f->w->live_mask &= ~(1u << f->lane);
							// lanes waiting on a warp barrier that named this lane are released
							unsigned expect = (unsigned)__builtin_popcount(f->w->live_mask);
							if (f->w->count && f->w->count >= expect) { f->w->count = 0; f->w->gen++; }
						}
					}
					if (g_progress == p1) break;
				}
			}
			if (blk.live && blk.count >= blk.live) {
				blk.count = 0; blk.gen++; g_progress++;
				for (unsigned t = 0; t < nthreads; t++) fibers[t].at_block_barrier = false;
			}
			if (g_progress == p0 && done < nthreads) {
				fprintf(stderr, "simt: deadlock in block (%u,%u,%u): %u of %u threads finished, %u at __syncthreads\n",
				        bx, by, bz, done, nthreads, blk.count);
				abort();
			}
		}
	}
	g_body = saved_body;
	cur = NULL;
}
} // namespace simt

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <map>
__thread int simt_cur_device = 0;
// allocation table behind cudaPointerGetAttributes: base -> (size, kind, device)
struct simt_alloc_rec { size_t size; int kind, device; };
static std::mutex g_alloc_lock;
static std::map<uintptr_t, simt_alloc_rec> g_allocs;
void simt_track_alloc(void *p, size_t n, int kind)
{
	std::lock_guard<std::mutex> guard(g_alloc_lock);
	simt_alloc_rec r = { n, kind, simt_cur_device };
	g_allocs[(uintptr_t)p] = r;
}
void simt_track_free(void *p)
{
	std::lock_guard<std::mutex> guard(g_alloc_lock);
	g_allocs.erase((uintptr_t)p);
}
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p)
{
	std::lock_guard<std::mutex> guard(g_alloc_lock);
	a->type = cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = 0; a->hostPointer = 0;
	auto it = g_allocs.upper_bound((uintptr_t)p);
	if (it == g_allocs.begin()) return cudaSuccess;
	--it;
	if ((uintptr_t)p >= it->first + it->second.size) return cudaSuccess;
	a->type = it->second.kind; a->device = it->second.device;
	a->devicePointer = (void*)p;
	a->hostPointer = it->second.kind == cudaMemoryTypeHost ? (void*)p : 0;
	return cudaSuccess;
}
static std::mutex g_shm_lock;
static std::vector<simt_shm_block> g_shm_blocks;     // blocks this process created
static std::vector<simt_shm_block> g_shm_mapped;     // blocks of other processes mapped here
// whatever the process still holds at exit (staging buffers, scenes nobody freed) must not stay in /dev/shm
static struct simt_shm_cleanup {
	~simt_shm_cleanup() { for (auto &b : g_shm_blocks) shm_unlink(b.name); }
} g_shm_cleanup;
simt_shm_block *simt_shm_find(const void *p)
{
	for (auto &b : g_shm_blocks) if (b.ptr == p) return &b;
	return NULL;
}
void *simt_shm_alloc(size_t n)
{
	std::lock_guard<std::mutex> guard(g_shm_lock);
	static unsigned counter = 0;
	simt_shm_block b;
	snprintf(b.name, sizeof(b.name), "/simt_%d_%u", (int)getpid(), counter++);
	b.size = (n + 4095) & ~(size_t)4095;
	int fd = shm_open(b.name, O_CREAT | O_EXCL | O_RDWR, 0600);
	if (fd < 0) return NULL;
	if (ftruncate(fd, (off_t)b.size) != 0) { close(fd); shm_unlink(b.name); return NULL; }
	b.ptr = mmap(NULL, b.size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (b.ptr == MAP_FAILED) { shm_unlink(b.name); return NULL; }
	g_shm_blocks.push_back(b);
	return b.ptr;
}
bool simt_shm_release(void *p)
{
	std::lock_guard<std::mutex> guard(g_shm_lock);
	for (size_t i = 0; i < g_shm_blocks.size(); i++) if (g_shm_blocks[i].ptr == p) {
		munmap(p, g_shm_blocks[i].size);
		shm_unlink(g_shm_blocks[i].name);
		g_shm_blocks.erase(g_shm_blocks.begin() + (long)i);
		return true;
	}
	return false;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p)
{
	std::lock_guard<std::mutex> guard(g_shm_lock);
	memset(h, 0, sizeof(*h));
	if (simt_shm_block *b = simt_shm_find(p)) {
		h->reserved[0] = 'S';
		memcpy(h->reserved + 8, &b->size, sizeof(size_t));
		memcpy(h->reserved + 16, b->name, sizeof(b->name));
	} else {
		h->reserved[0] = 'P';
		memcpy(h->reserved + 8, &p, sizeof(p));
	}
	return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned)
{
	std::lock_guard<std::mutex> guard(g_shm_lock);
	if (h.reserved[0] == 'P') { memcpy(p, h.reserved + 8, sizeof(*p)); return *p ? cudaSuccess : cudaErrorInvalidValue; }
	if (h.reserved[0] != 'S') return cudaErrorInvalidValue;
	simt_shm_block b;
	memcpy(&b.size, h.reserved + 8, sizeof(size_t));
	memcpy(b.name, h.reserved + 16, sizeof(b.name));
	b.name[sizeof(b.name) - 1] = 0;
	int fd = shm_open(b.name, O_RDWR, 0600);
	if (fd < 0) return cudaErrorInvalidValue;
	b.ptr = mmap(NULL, b.size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (b.ptr == MAP_FAILED) return cudaErrorInvalidValue;
	g_shm_mapped.push_back(b);
	*p = b.ptr;
	return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *p)
{
	std::lock_guard<std::mutex> guard(g_shm_lock);
	for (size_t i = 0; i < g_shm_mapped.size(); i++) if (g_shm_mapped[i].ptr == p) {
		munmap(p, g_shm_mapped[i].size);
		g_shm_mapped.erase(g_shm_mapped.begin() + (long)i);
		return cudaSuccess;
	}
	return cudaSuccess;      // heap-backed "mapping": nothing to undo
}
#endif
