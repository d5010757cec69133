// rtk_device.cu -- the CUDA translation unit: kernels (k_build.cuh, k_trace.cuh) plus the thin
// C-ABI layer (rtk_device.h) that the C host code calls.  Compiled with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
// There is no CPU implementation behind these entry points: without a CUDA device every call
// fails with RTK_CUDA_ERR_NO_DEVICE.
#include "rtk_device.h"
#include "k_build.cuh"
#include "k_sah.cuh"
#include "k_trace.cuh"
#include "k_wavefront.cuh"
#include <stdarg.h>
#include <pthread.h>

#define RTKD_OK 0
#define RTKD_ERR_NO_DEVICE (-1)
#define RTKD_ERR_CUDA (-2)
#define RTKD_ERR_ARGUMENT (-3)
#define RTKD_ERR_SCENE (-4)
#define RTKD_ERR_MEMORY (-5)
#define RTKD_ERR_OVERFLOW (-6)

static __thread char g_err[512];
static int g_device = -1;
static int g_sm_count = 0, g_trace_ctas = 0, g_trace_lanes = RTK_TRACE_LANES, g_trace_pd = 1;
static int g_reserved_sms = 0;      // SMs the persistent traversal grid leaves free (for concurrent NCCL kernels)
static size_t g_l2_bytes = 0;
static uint64_t g_next_id = 1;

extern "C" const char *rtkd_last_error(void) { return g_err; }
extern "C" void rtkd_set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}

#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { \
	rtkd_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); return RTKD_ERR_CUDA; } } while (0)
#define CKP(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { \
	rtkd_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); return NULL; } } while (0)
#define CK_LAUNCH() CK(cudaGetLastError())

// CK() returns from the middle of a function: these release what the function holds on every path
struct event_pair {
	cudaEvent_t e0 = NULL, e1 = NULL;
	~event_pair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};
struct async_free {
	void *&ptr; cudaStream_t st;
	async_free(void *&p, cudaStream_t s) : ptr(p), st(s) {}
	~async_free() { if (ptr) cudaFreeAsync(ptr, st); }
};

extern "C" int rtkd_init(int device)
{
	if (g_device == device && g_sm_count) { CK(cudaSetDevice(device)); return RTKD_OK; }   // another host thread binding itself
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0) {
		rtkd_set_error("rtk_b200: no usable CUDA device (%s); this library has no CPU fallback",
		               e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
		return RTKD_ERR_NO_DEVICE;
	}
	if (device < 0 || device >= count) { rtkd_set_error("device %d out of range (0..%d)", device, count - 1); return RTKD_ERR_ARGUMENT; }
	CK(cudaSetDevice(device));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, device));
	g_sm_count = prop.multiProcessorCount;
	g_l2_bytes = (size_t)prop.l2CacheSize;
	{
		const char *e = getenv("RTK_B200_LANES");          // experiment knob: lanes per ray
		if (e && (atoi(e) == 8 || atoi(e) == 4 || atoi(e) == 2)) g_trace_lanes = atoi(e);
	}
	{
		const char *e = getenv("RTK_B200_PD");             // experiment knob: 0 = every ray's own lanes walk its leaf
		if (e) g_trace_pd = atoi(e) != 0;
		if (g_trace_lanes == 8) g_trace_pd = 0;
	}
	int ctas = 0;
	if (g_trace_lanes == 8) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<8, 1, false>, RTK_TRACE_THREADS, 0));
	else if (g_trace_lanes == 4 && g_trace_pd) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<4, 1, false, false, true>, RTK_TRACE_THREADS, 0));
	else if (g_trace_lanes == 4) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<4, 1, false>, RTK_TRACE_THREADS, 0));
	else if (g_trace_pd) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<2, 1, false, false, true>, RTK_TRACE_THREADS, 0));
	else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<2, 1, false>, RTK_TRACE_THREADS, 0));
	g_trace_ctas = ctas > 0 ? ctas : 1;
#ifndef RTK_SIMT_EMU
	{
		// keep freed scratch in the stream-ordered pool: the build allocates its temporaries
		// with cudaMallocAsync on every (re)build
		cudaMemPool_t pool;
		if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
			unsigned long long thr = ~0ull;
			cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
		}
	}
#endif
	g_device = device;
	return RTKD_OK;
}

// Every entry point passes through here (directly or via rtkd_bind_thread): the library is bound to
// ONE device per process, but CUDA's current device is per host thread and starts at 0 -- a worker
// thread of the caller that traces against a scene on device 3 must be switched to device 3 first
// (the reference's rtk_trace_ray is called from many user threads, rtk.h:129).
static int ensure_init(void)
{
	if (!g_sm_count) return rtkd_init(0);
	int cur = -1;
	if (cudaGetDevice(&cur) != cudaSuccess || cur != g_device) CK(cudaSetDevice(g_device));
	return RTKD_OK;
}
extern "C" int rtkd_bind_thread(void) { return ensure_init(); }

static void stage_shutdown(void);          // host-batch staging (defined with the pipeline below)
extern "C" void rtkd_shutdown(void)
{
	// scenes are the caller's to free first; what the library itself holds on the device goes here
	if (g_sm_count) { ensure_init(); cudaDeviceSynchronize(); stage_shutdown(); }
	g_device = -1; g_sm_count = 0;
}

extern "C" int rtkd_reserve_sms(int sms)
{
	int r = ensure_init();
	if (r) return r;
	if (sms < 0 || sms >= g_sm_count) { rtkd_set_error("cannot reserve %d of %d SMs", sms, g_sm_count); return RTKD_ERR_ARGUMENT; }
	g_reserved_sms = sms;
	return RTKD_OK;
}

extern "C" int rtkd_device_info(int *sm_count, size_t *l2_bytes, int *ctas_per_sm, int *threads_per_cta)
{
	int r = ensure_init();
	if (r) return r;
	if (sm_count) *sm_count = g_sm_count;
	if (l2_bytes) *l2_bytes = g_l2_bytes;
	if (ctas_per_sm) *ctas_per_sm = g_trace_ctas;
	if (threads_per_cta) *threads_per_cta = RTK_TRACE_THREADS;
	return RTKD_OK;
}

// ---------------------------------------------------------------------------------------------
// read-bandwidth probe: the denominator of the L2 roofline ("L2 peak must be measured
// with a read microbenchmark on the same box").  A persistent grid streams a buffer `passes` times
// with 16-byte loads; a buffer smaller than the L2 measures L2, a larger one HBM.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) k_read_probe(const float4 *buf, size_t n16, int passes, float *sink)
{
	float acc = 0.0f;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	for (int p = 0; p < passes; p++) {
		size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
		for (; i + 3 * stride < n16; i += 4 * stride) {
			float4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
			acc += (a.x + b.y) + (c.z + d.w);
		}
		for (; i < n16; i += stride) acc += __ldcg(buf + i).x;
	}
	if (acc == 123.456f) *sink = acc;            // never true: keeps the loads alive
}

extern "C" int rtkd_read_bandwidth(size_t bytes, int passes, double *gbs)
{
	if (ensure_init()) return RTKD_ERR_NO_DEVICE;
	if (bytes < 4096 || passes < 1 || !gbs) { rtkd_set_error("bad probe arguments"); return RTKD_ERR_ARGUMENT; }
	float4 *buf = NULL;
	float *sink = NULL;
	const size_t n16 = bytes / 16;
	struct dev_free { void *p = NULL; ~dev_free() { if (p) cudaFree(p); } } buf_guard, sink_guard;
	CK(cudaMalloc(&buf, n16 * 16));
	buf_guard.p = buf;
	CK(cudaMalloc(&sink, 4));
	sink_guard.p = sink;
	CK(cudaMemset(buf, 0, n16 * 16));
	event_pair ev;
	CK(cudaEventCreate(&ev.e0)); CK(cudaEventCreate(&ev.e1));
	cudaEvent_t e0 = ev.e0, e1 = ev.e1;
	const unsigned grid = (unsigned)g_sm_count * 8;
	RTK_LAUNCH(k_read_probe, grid, 256, 0, (const float4*)buf, n16, 2, sink);       // warm the cache
	CK(cudaEventRecord(e0, 0));
	RTK_LAUNCH(k_read_probe, grid, 256, 0, (const float4*)buf, n16, passes, sink);
	CK(cudaEventRecord(e1, 0));
	CK(cudaEventSynchronize(e1));
	float ms = 0.0f;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	*gbs = ms > 0.0f ? (double)n16 * 16.0 * passes / (ms * 1e-3) / 1e9 : 0.0;
	return RTKD_OK;
}

// ---------------------------------------------------------------------------------------------
// scene lifetime
// ---------------------------------------------------------------------------------------------

extern "C" rtkd_scene *rtkd_scene_new(uint32_t num_tris, uint32_t num_meshes, const uint32_t *mesh_first)
{
	if (ensure_init()) return NULL;
	rtkd_scene *s = (rtkd_scene*)calloc(1, sizeof(rtkd_scene));
	if (!s) { rtkd_set_error("out of host memory"); return NULL; }
	s->id = ((uint64_t)time(NULL) << 20) ^ (__atomic_fetch_add(&g_next_id, 1, __ATOMIC_RELAXED) * 0x9E3779B97F4A7C15ull);   // scenes may be built from several host threads
	s->num_tris = num_tris; s->num_meshes = num_meshes;
	s->h_mesh_first = (uint32_t*)malloc(sizeof(uint32_t) * (num_meshes + 1));
	memcpy(s->h_mesh_first, mesh_first, sizeof(uint32_t) * (num_meshes + 1));
	cudaError_t e = cudaMalloc((float4**)&s->tri_orig, sizeof(float4) * 3 * (size_t)(num_tris ? num_tris : 1));
	if (e == cudaSuccess) e = cudaMalloc((uint32_t**)&s->mesh_first, sizeof(uint32_t) * (num_meshes + 1));
	if (e == cudaSuccess) e = cudaMemcpy(s->mesh_first, mesh_first, sizeof(uint32_t) * (num_meshes + 1), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) {
		rtkd_set_error("scene allocation failed: %s", cudaGetErrorString(e));
		rtkd_scene_free(s);
		return NULL;
	}
	return s;
}

extern "C" void rtkd_scene_free(rtkd_scene *s)
{
	if (!s) return;
	if (g_sm_count) ensure_init();           // the freeing thread may not be the one that built the scene
	cudaDeviceSynchronize();
	if (s->tri_orig) cudaFree(s->tri_orig);
	if (s->tv0) cudaFree(s->tv0);
	if (s->tv1) cudaFree(s->tv1);
	if (s->tv2) cudaFree(s->tv2);
	if (s->nodes) cudaFree(s->nodes);
	if (s->node_level) cudaFree(s->node_level);
	if (s->mesh_first) cudaFree(s->mesh_first);
	if (s->scratch) cudaFree(s->scratch);
	if (s->overflow) cudaFree(s->overflow);
	if (s->hit16) cudaFree(s->hit16);
	if (s->filter_bits) cudaFree(s->filter_bits);
	free(s->h_mesh_first);
	free(s);
}

extern "C" void *rtkd_upload(const void *host, size_t bytes, void *stream)
{
	if (ensure_init()) return NULL;
	unsigned char *d = NULL;
	CKP(cudaMallocAsync(&d, bytes ? bytes : 16, (cudaStream_t)stream));
	if (bytes) CKP(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
	return d;
}
extern "C" void rtkd_free_async(void *dev, void *stream) { if (dev) cudaFreeAsync(dev, (cudaStream_t)stream); }
extern "C" int rtkd_sync(void *stream) { CK(cudaStreamSynchronize((cudaStream_t)stream)); return RTKD_OK; }

extern "C" int rtkd_max_index(const void *d_idx, size_t stride, int idx_bytes, uint32_t ntris, uint32_t *out, void *stream)
{
	uint32_t *d = NULL;
	CK(cudaMallocAsync(&d, sizeof(uint32_t), (cudaStream_t)stream));
	CK(cudaMemsetAsync(d, 0, sizeof(uint32_t), (cudaStream_t)stream));
	if (ntris) {
		RTK_LAUNCH(k_max_index, (ntris + 255) / 256, 256, stream, (const unsigned char*)d_idx, (unsigned long long)stride, idx_bytes, ntris, d);
		CK_LAUNCH();
	}
	CK(cudaMemcpyAsync(out, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	CK(cudaStreamSynchronize((cudaStream_t)stream));
	cudaFreeAsync(d, (cudaStream_t)stream);
	return RTKD_OK;
}

extern "C" int rtkd_decode_mesh(rtkd_scene *s, uint32_t first_prim, uint32_t ntris,
                                const void *d_pos, size_t pos_stride, int pos_f64,
                                const void *d_idx, size_t idx_stride, int idx_bytes, int pregathered, void *stream)
{
	return rtkd_decode_mesh_xf(s, first_prim, ntris, d_pos, pos_stride, pos_f64, d_idx, idx_stride, idx_bytes, pregathered, NULL, stream);
}

extern "C" int rtkd_decode_mesh_xf(rtkd_scene *s, uint32_t first_prim, uint32_t ntris,
                                   const void *d_pos, size_t pos_stride, int pos_f64,
                                   const void *d_idx, size_t idx_stride, int idx_bytes, int pregathered,
                                   const float *xf12, void *stream)
{
	if (!ntris) return RTKD_OK;
	if ((size_t)first_prim + ntris > s->num_tris) { rtkd_set_error("decode range outside the scene"); return RTKD_ERR_ARGUMENT; }
	rtkd_decode_args a;
	a.pos = (const unsigned char*)d_pos; a.idx = (const unsigned char*)d_idx;
	a.pos_stride = pos_stride; a.idx_stride = idx_stride;
	a.pos_f64 = pos_f64; a.idx_bytes = d_idx ? idx_bytes : 0; a.pregathered = pregathered;
	a.ntris = ntris; a.first_prim = first_prim;
	a.has_xf = xf12 != NULL;
	for (int k = 0; k < 12; k++) a.xf[k] = xf12 ? xf12[k] : 0.0f;
	RTK_LAUNCH(k_decode_mesh, (ntris + 255) / 256, 256, stream, a, (float4*)s->tri_orig);
	CK_LAUNCH();
	return RTKD_OK;
}

// ---------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------

// All temporaries of one build come from ONE stream-ordered allocation (the pool keeps it cached
// between rebuilds); sub-buffers are carved out with 256-byte alignment.
struct build_arena {
	unsigned char *base; size_t size, used;
	template <typename T> T *take(size_t count)
	{
		size_t bytes = (sizeof(T) * (count ? count : 1) + 255) & ~(size_t)255;
		if (!base) { used += bytes; return NULL; }            // sizing pass
		T *p = (T*)(base + used);
		used += bytes;
		return p;
	}
};

struct build_bufs {
	uint32_t *d_bounds;
	unsigned long long *keys[2];
	uint32_t *vals[2], *counts, *totals;
	rtkd_bvh2 t;
	rtkd_sah h;
	uint32_t *order;
	float4 *wide;
	uint2 *work[2];
	uint32_t *ctr;
	uint2 *leaf_list;
	unsigned char *node_level;
	double *d_cost;
	uint32_t cap, nblocks;
	size_t act_cap, small_cap;
};

#define RTKD_COLLAPSE_LEVELS 80

static void carve(build_arena &A, build_bufs &B, uint32_t n, bool use_sah)
{
	memset(&B.t, 0, sizeof(B.t));
	memset(&B.h, 0, sizeof(B.h));
	B.nblocks = (n + RTK_SORT_TILE - 1) / RTK_SORT_TILE;
	B.cap = (uint32_t)(((unsigned long long)n * 4) / 7 + 16);
	B.d_bounds = A.take<uint32_t>(8);
	B.keys[0] = A.take<unsigned long long>(n); B.keys[1] = A.take<unsigned long long>(n);
	B.vals[0] = A.take<uint32_t>(n); B.vals[1] = A.take<uint32_t>(n);
	B.counts = A.take<uint32_t>(256 * (size_t)B.nblocks); B.totals = A.take<uint32_t>(256);
	if (use_sah) {
		const size_t cap2 = 2 * (size_t)n + 2;
		B.act_cap = n / RTK_SAH_SMALL + 4;
		B.small_cap = 4 * (size_t)(n / RTK_SAH_SMALL) + 8;
		B.h.pb = A.take<float4>(2 * (size_t)n);
		B.h.idx0 = A.take<uint32_t>(n); B.h.idx1 = A.take<uint32_t>(n); B.h.idx_final = A.take<uint32_t>(n);
		B.h.left = A.take<int>(cap2); B.h.right = A.take<int>(cap2); B.h.first = A.take<int>(cap2); B.h.last = A.take<int>(cap2);
		B.h.blo = A.take<float4>(cap2); B.h.bhi = A.take<float4>(cap2); B.h.ndepth = A.take<uint32_t>(cap2);
		B.h.counters = A.take<uint32_t>(8);
		B.h.act_in = A.take<uint32_t>(B.act_cap); B.h.act_out = A.take<uint32_t>(B.act_cap);
		B.h.small_list = A.take<uint32_t>(B.small_cap);
		B.h.chunk_base = A.take<uint32_t>(B.act_cap + 1);
		B.h.bins = A.take<uint32_t>(B.act_cap * RTK_SAH_NODEBINS);
		B.h.split = A.take<int4>(B.act_cap); B.h.cursor = A.take<uint32_t>(2 * B.act_cap);
		B.h.node_cap = (uint32_t)cap2;
		B.order = A.take<uint32_t>(n);
	} else if (n > 1) {
		B.t.left = A.take<int>(n - 1); B.t.right = A.take<int>(n - 1);
		B.t.parent = A.take<int>(2 * (size_t)n - 1);
		B.t.first = A.take<int>(n - 1); B.t.last = A.take<int>(n - 1);
		B.t.blo = A.take<float4>(2 * (size_t)n - 1); B.t.bhi = A.take<float4>(2 * (size_t)n - 1);
		B.t.flags = A.take<int>(n - 1);
	}
	B.wide = A.take<float4>(16 * (size_t)B.cap);
	B.work[0] = A.take<uint2>(B.cap); B.work[1] = A.take<uint2>(B.cap);
	B.ctr = A.take<uint32_t>(8 + RTKD_COLLAPSE_LEVELS);
	B.leaf_list = A.take<uint2>(n);
	B.node_level = A.take<unsigned char>(B.cap);
	B.d_cost = A.take<double>(1);
}

// binned-SAH binary tree (k_sah.cuh): one host synchronisation per level of large nodes
static int build_sah(cudaStream_t st, const float4 *tri, const uint32_t *svals, build_bufs &B, uint32_t n)
{
	rtkd_sah &h = B.h;
	RTK_LAUNCH(k_sah_prim_bounds, (n + 255) / 256, 256, st, tri, svals, n, (float4*)h.pb, h.idx0); CK_LAUNCH();
	RTK_LAUNCH(k_sah_root, 1, 32, st, h, (const uint32_t*)B.d_bounds, n); CK_LAUNCH();
	RTK_LAUNCH(k_sah_plan, 1, 1024, st, h, (const uint32_t*)h.act_in); CK_LAUNCH();
	uint32_t hc[8];
	CK(cudaMemcpyAsync(hc, h.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	uint32_t n_act = hc[1], chunks = hc[5], depth = 0;
	int src_buf = 0;
	while (n_act) {
		if (n_act > B.act_cap) { rtkd_set_error("SAH active list overflow"); return RTKD_ERR_MEMORY; }
		CK(cudaMemsetAsync(h.counters + 1, 0, sizeof(uint32_t), st));
		RTK_LAUNCH(k_sah_bins_clear, n_act, 128, st, h, n_act); CK_LAUNCH();
		RTK_LAUNCH(k_sah_bin_large, chunks, 256, st, h, n_act, src_buf); CK_LAUNCH();
		RTK_LAUNCH(k_sah_split_large, (n_act + 3) / 4, 128, st, h, n_act, depth, src_buf ^ 1); CK_LAUNCH();
		RTK_LAUNCH(k_sah_partition_large, chunks, 256, st, h, n_act, src_buf); CK_LAUNCH();
		uint32_t *tmp = h.act_in; h.act_in = h.act_out; h.act_out = tmp;
		RTK_LAUNCH(k_sah_plan, 1, 1024, st, h, (const uint32_t*)h.act_in); CK_LAUNCH();
		CK(cudaMemcpyAsync(hc, h.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		n_act = hc[1]; chunks = hc[5];
		src_buf ^= 1;
		depth++;
	}
	if (hc[2] > B.small_cap) { rtkd_set_error("SAH small-subtree list overflow"); return RTKD_ERR_MEMORY; }
	if (hc[2]) { RTK_LAUNCH(k_sah_small, hc[2], RTK_SAH_SMALL_THREADS, st, h, hc[2]); CK_LAUNCH(); }
	RTK_LAUNCH(k_sah_compose, (n + 255) / 256, 256, st, (const uint32_t*)h.idx_final, svals, n, B.order); CK_LAUNCH();
	return RTKD_OK;
}

// (re)writes corner 0 of every leaf-slot triangle according to the scene's filter bitset
static int apply_filter(rtkd_scene *s, cudaStream_t st)
{
	if (!s->num_tv) return RTKD_OK;
	RTK_LAUNCH(k_apply_filter, (s->num_tv + 255) / 256, 256, st, (const float4*)s->tri_orig, (const uint32_t*)s->filter_bits,
	           s->num_tv, (float4*)s->tv0);
	CK_LAUNCH();
	return RTKD_OK;
}

extern "C" int rtkd_set_filter(rtkd_scene *s, const void *bits, size_t num_words, int on_device, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	const size_t words = ((size_t)s->num_tris + 31) / 32;
	if (!bits) {
		if (!s->filter_bits) return RTKD_OK;
		// the slots go back to the decoded corners; the bitset is released once that pass has run
		void *old = s->filter_bits;
		s->filter_bits = NULL;
		int r = apply_filter(s, st);
		CK(cudaStreamSynchronize(st));
		cudaFree(old);
		return r;
	}
	if (num_words < words) { rtkd_set_error("triangle filter needs %zu words for %u triangles, got %zu", words, s->num_tris, num_words); return RTKD_ERR_ARGUMENT; }
	if (!words) return RTKD_OK;
	if (!s->filter_bits) CK(cudaMalloc((uint32_t**)&s->filter_bits, sizeof(uint32_t) * words));
	CK(cudaMemcpyAsync(s->filter_bits, bits, sizeof(uint32_t) * words, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
	int r = apply_filter(s, st);
	// filters change rarely: return with the pass done, whatever stream the next query uses
	CK(cudaStreamSynchronize(st));
	return r;
}

extern "C" int rtkd_build(rtkd_scene *s, int mode, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	const uint32_t n = s->num_tris;
	s->build_mode = (uint32_t)mode;
	event_pair ev;
	CK(cudaEventCreate(&ev.e0)); CK(cudaEventCreate(&ev.e1));
	cudaEvent_t e0 = ev.e0, e1 = ev.e1;
	CK(cudaEventRecord(e0, st));

	s->num_nodes = 0; s->num_leaves = 0; s->num_tv = 0; s->depth = 0; s->sah_cost = 0.0;
	for (int k = 0; k < 3; k++) { s->bounds_min[k] = 0.0f; s->bounds_max[k] = 0.0f; }
	s->abs_max = 0.0f;
	if (n == 0) {
		CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
		s->build_device_ms = 0.0;
		return RTKD_OK;
	}
	const bool use_sah = mode == 1 && n > RTK_LEAF_MAX;
	build_arena A = { NULL, 0, 0 };
	build_bufs B;
	carve(A, B, n, use_sah);                               // sizing pass
	A.size = A.used; A.used = 0;
	CK(cudaMallocAsync(&A.base, A.size, st));
	void *arena_mem = A.base;
	async_free arena_guard(arena_mem, st);                 // the temporaries go back to the pool on every path
	carve(A, B, n, use_sah);

	const float4 *tri = (const float4*)s->tri_orig;
	// scene bounds (ordered-uint encoded: min starts at 0xffffffff, max at 0)
	CK(cudaMemsetAsync(B.d_bounds, 0xff, 3 * sizeof(uint32_t), st));
	CK(cudaMemsetAsync(B.d_bounds + 3, 0x00, 3 * sizeof(uint32_t), st));
	RTK_LAUNCH(k_scene_bounds, (n + 255) / 256, 256, st, tri, n, B.d_bounds); CK_LAUNCH();
	RTK_LAUNCH(k_morton, (n + 255) / 256, 256, st, tri, n, (const uint32_t*)B.d_bounds, B.keys[0], B.vals[0]); CK_LAUNCH();

	// LSD radix sort: all 63 code bits for the radix tree of the LBVH mode (it needs them to tell
	// neighbours apart); the SAH builder only wants spatial locality in its input order, for which
	// the upper 31 bits (10 per axis) are plenty -- half the passes
	int src = 0;
	for (int shift = use_sah ? 32 : 0; shift < 64; shift += 8) {
		RTK_LAUNCH(k_radix_hist, B.nblocks, RTK_SORT_THREADS, st, (const unsigned long long*)B.keys[src], n, shift, B.counts, B.nblocks); CK_LAUNCH();
		RTK_LAUNCH(k_radix_scan, 256, 256, st, B.counts, B.nblocks, B.totals); CK_LAUNCH();
		RTK_LAUNCH(k_radix_scatter, B.nblocks, RTK_SORT_THREADS, st, (const unsigned long long*)B.keys[src], (const uint32_t*)B.vals[src],
		           B.keys[src ^ 1], B.vals[src ^ 1], n, shift, (const uint32_t*)B.counts, (const uint32_t*)B.totals, B.nblocks); CK_LAUNCH();
		src ^= 1;
	}
	const unsigned long long *skeys = B.keys[src];
	const uint32_t *svals = B.vals[src];

	uint32_t h_bounds[6];
	CK(cudaMemcpyAsync(h_bounds, B.d_bounds, sizeof(h_bounds), cudaMemcpyDeviceToHost, st));

	// binary tree: binned SAH over the Morton-ordered triangles, or the radix tree itself
	rtkd_bvh2 t = B.t;
	if (use_sah) {
		int r = build_sah(st, tri, svals, B, n);
		if (r) return r;
		t.left = B.h.left; t.right = B.h.right; t.first = B.h.first; t.last = B.h.last; t.blo = B.h.blo; t.bhi = B.h.bhi;
		svals = B.order;
	} else if (n > 1) {
		CK(cudaMemsetAsync(t.flags, 0, sizeof(int) * (size_t)(n - 1), st));
		RTK_LAUNCH(k_hierarchy, (n - 1 + 255) / 256, 256, st, skeys, (int)n, t); CK_LAUNCH();
		RTK_LAUNCH(k_refit, (n + 255) / 256, 256, st, tri, svals, (int)n, t); CK_LAUNCH();
	}
	uint32_t num_nodes = 0, num_leaves = 0, depth = 0;
	double h_cost = 0.0;
	CK(cudaMemsetAsync(B.node_level, 0, B.cap, st));       // the root is at depth 0
	if (n == 1) {
		RTK_LAUNCH(k_single_root, 1, 32, st, tri, svals, B.wide); CK_LAUNCH();
		const uint2 l0 = make_uint2(0u, 1u);
		CK(cudaMemcpyAsync(B.leaf_list, &l0, sizeof(l0), cudaMemcpyHostToDevice, st));
		num_nodes = 1; num_leaves = 1; depth = 1;
	} else {
		// ctr: [0] node_alloc [1] leaf_count [2] err [8 + L] items queued for level L
		uint32_t h_ctr[8 + RTKD_COLLAPSE_LEVELS];
		memset(h_ctr, 0, sizeof(h_ctr));
		h_ctr[0] = 1; h_ctr[8] = 1;
		uint2 w0 = make_uint2(0u, 0u);
		CK(cudaMemcpyAsync(B.ctr, h_ctr, sizeof(h_ctr), cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(B.work[0], &w0, sizeof(w0), cudaMemcpyHostToDevice, st));
		CK(cudaMemsetAsync(B.d_cost, 0, sizeof(double), st));
		// levels are launched with an upper bound on their width (8^L, capped) and read their
		// true item count on the device; the host looks at the counters every 4 levels
		unsigned long long bound = 1;
		int level = 0;
		bool done = false;
		while (!done) {
			for (int k = 0; k < 4 && level < RTKD_COLLAPSE_LEVELS - 1; k++, level++) {
				rtkd_collapse_args a;
				a.work_in = B.work[level & 1]; a.n_in = B.ctr + 8 + level;
				a.work_out = B.work[(level & 1) ^ 1]; a.n_out = B.ctr + 8 + level + 1;
				a.node_alloc = B.ctr; a.node_cap = B.cap; a.leaf_count = B.ctr + 1; a.leaf_list = B.leaf_list; a.sah_cost = B.d_cost;
				a.node_level = B.node_level; a.level = (uint32_t)level;
				a.nodes = B.wide; a.n = (int)n; a.err = B.ctr + 2;
				uint32_t width = (uint32_t)(bound < B.cap ? bound : B.cap);
				RTK_LAUNCH(k_collapse, (width + 127) / 128, 128, st, a, t); CK_LAUNCH();
				bound = bound * 8 < B.cap ? bound * 8 : B.cap;
			}
			CK(cudaMemcpyAsync(h_ctr, B.ctr, sizeof(h_ctr), cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
			done = h_ctr[8 + level] == 0 || level >= RTKD_COLLAPSE_LEVELS - 1;
		}
		for (depth = 0; depth < RTKD_COLLAPSE_LEVELS && h_ctr[8 + depth]; depth++) { }
		num_nodes = h_ctr[0]; num_leaves = h_ctr[1];
		if (h_ctr[2] || h_ctr[8 + RTKD_COLLAPSE_LEVELS - 1]) {
			rtkd_set_error("wide-node pool exhausted (cap %u), tree deeper than %d or more than %u leaves", B.cap, RTKD_COLLAPSE_LEVELS, RTK_MAX_LEAVES);
			return RTKD_ERR_MEMORY;
		}
		CK(cudaMemcpyAsync(&h_cost, B.d_cost, sizeof(double), cudaMemcpyDeviceToHost, st));
	}

	// node array: reuse the previous allocation when it is large enough
	if (!s->nodes || s->nodes_cap < num_nodes) {
		if (s->nodes) cudaFree(s->nodes);
		s->nodes = NULL;
		CK(cudaMalloc((float4**)&s->nodes, sizeof(float4) * 16 * (size_t)num_nodes));
		if (s->node_level) cudaFree(s->node_level);
		s->node_level = NULL;
		CK(cudaMalloc((unsigned char**)&s->node_level, (size_t)num_nodes));
		s->nodes_cap = num_nodes;
	}
	if (!s->node_level) CK(cudaMalloc((unsigned char**)&s->node_level, (size_t)s->nodes_cap));
	CK(cudaMemcpyAsync(s->node_level, B.node_level, (size_t)num_nodes, cudaMemcpyDeviceToDevice, st));
	CK(cudaMemcpyAsync(s->nodes, B.wide, sizeof(float4) * 16 * (size_t)num_nodes, cudaMemcpyDeviceToDevice, st));
	// traversal triangles: one 8-entry slot per leaf, sized now that the leaves are counted
	const uint32_t num_tv = num_leaves * RTK_LEAF_MAX;
	if (!s->tv0 || s->tv_cap < num_tv) {
		if (s->tv0) { cudaFree(s->tv0); cudaFree(s->tv1); cudaFree(s->tv2); }
		s->tv0 = s->tv1 = s->tv2 = NULL;
		const uint32_t cap = num_tv + num_tv / 16 + 64;        // rebuilds of a deforming mesh rarely need a new allocation
		CK(cudaMalloc((float4**)&s->tv0, sizeof(float4) * (size_t)cap));
		CK(cudaMalloc((float4**)&s->tv1, sizeof(float4) * (size_t)cap));
		CK(cudaMalloc((float4**)&s->tv2, sizeof(float4) * (size_t)cap));
		s->tv_cap = cap;
	}
	RTK_LAUNCH(k_emit_leaves, (num_tv + 255) / 256, 256, st, tri, svals, (const uint2*)B.leaf_list, num_leaves,
	           (float4*)s->tv0, (float4*)s->tv1, (float4*)s->tv2); CK_LAUNCH();
	s->num_tv = num_tv;
	if (s->filter_bits) { int fr = apply_filter(s, st); if (fr) return fr; }

	CK(cudaEventRecord(e1, st));
	CK(cudaEventSynchronize(e1));
	float ms = 0.0f;
	CK(cudaEventElapsedTime(&ms, e0, e1));

	s->num_nodes = num_nodes; s->num_leaves = num_leaves; s->depth = depth;
	s->build_device_ms = ms;
	float amax = 0.0f;
	for (int k = 0; k < 3; k++) {
		uint32_t u = h_bounds[k], v = h_bounds[3 + k];
		u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
		v = (v & 0x80000000u) ? (v & 0x7fffffffu) : ~v;
		memcpy(&s->bounds_min[k], &u, 4);
		memcpy(&s->bounds_max[k], &v, 4);
		amax = fmaxf(amax, fmaxf(fabsf(s->bounds_min[k]), fabsf(s->bounds_max[k])));
	}
	s->abs_max = amax;
	{
		float x = s->bounds_max[0] - s->bounds_min[0], y = s->bounds_max[1] - s->bounds_min[1], z = s->bounds_max[2] - s->bounds_min[2];
		double ra = (double)x * y + (double)y * z + (double)z * x;
		s->sah_cost = ra > 0.0 ? 1.0 + h_cost / ra : 0.0;
	}
	// the stack scratch depends on the depth: re-created on the next query if too small
	return RTKD_OK;
}

// Refit: same tree, new boxes (k_refit_tris + one k_refit_level launch per depth, deepest first),
// new scene bounds.  Valid after rtkd_decode_mesh rewrote the corners of a scene built here.
extern "C" int rtkd_refit(rtkd_scene *s, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	if (!s->num_tris || !s->num_nodes) return RTKD_OK;
	if (!s->node_level) { rtkd_set_error("this scene carries no level table (loaded from a blob): rebuild it instead"); return RTKD_ERR_SCENE; }
	event_pair ev;
	CK(cudaEventCreate(&ev.e0)); CK(cudaEventCreate(&ev.e1));
	cudaEvent_t e0 = ev.e0, e1 = ev.e1;
	CK(cudaEventRecord(e0, st));
	const float4 *tri = (const float4*)s->tri_orig;
	uint32_t *d_bounds = NULL;
	CK(cudaMallocAsync(&d_bounds, 8 * sizeof(uint32_t), st));
	void *bounds_mem = d_bounds;
	async_free bounds_guard(bounds_mem, st);
	CK(cudaMemsetAsync(d_bounds, 0xff, 3 * sizeof(uint32_t), st));
	CK(cudaMemsetAsync(d_bounds + 3, 0x00, 3 * sizeof(uint32_t), st));
	RTK_LAUNCH(k_scene_bounds, (s->num_tris + 255) / 256, 256, st, tri, s->num_tris, d_bounds); CK_LAUNCH();
	RTK_LAUNCH(k_refit_tris, (s->num_tv + 255) / 256, 256, st, tri, s->num_tv, (float4*)s->tv0, (float4*)s->tv1, (float4*)s->tv2); CK_LAUNCH();
	const unsigned blocks = (unsigned)(((size_t)s->num_nodes * RTK_WIDE + 255) / 256);
	if (s->num_tris == 1) {
		// the single-triangle scene is one root with one leaf child
		RTK_LAUNCH(k_refit_level, blocks, 256, st, (float4*)s->nodes, (const unsigned char*)s->node_level, s->num_nodes, 0u,
		           (const float4*)s->tv0, (const float4*)s->tv1, (const float4*)s->tv2); CK_LAUNCH();
	} else for (int level = (int)s->depth - 1; level >= 0; level--) {
		RTK_LAUNCH(k_refit_level, blocks, 256, st, (float4*)s->nodes, (const unsigned char*)s->node_level, s->num_nodes, (uint32_t)level,
		           (const float4*)s->tv0, (const float4*)s->tv1, (const float4*)s->tv2); CK_LAUNCH();
	}
	if (s->filter_bits) { int fr = apply_filter(s, st); if (fr) return fr; }
	uint32_t h_bounds[6];
	CK(cudaMemcpyAsync(h_bounds, d_bounds, sizeof(h_bounds), cudaMemcpyDeviceToHost, st));
	CK(cudaEventRecord(e1, st));
	CK(cudaEventSynchronize(e1));
	float ms = 0.0f;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	s->build_device_ms = ms;
	float amax = 0.0f;
	for (int k = 0; k < 3; k++) {
		uint32_t u = h_bounds[k], v = h_bounds[3 + k];
		u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
		v = (v & 0x80000000u) ? (v & 0x7fffffffu) : ~v;
		memcpy(&s->bounds_min[k], &u, 4);
		memcpy(&s->bounds_max[k], &v, 4);
		amax = fmaxf(amax, fmaxf(fabsf(s->bounds_min[k]), fabsf(s->bounds_max[k])));
	}
	s->abs_max = amax;
	return RTKD_OK;
}

// ---------------------------------------------------------------------------------------------
// queries
// ---------------------------------------------------------------------------------------------

static void fill_arrays(const rtkd_scene *s, rtkd_arrays &a)
{
	a.tri_orig = (const float4*)s->tri_orig;
	a.tv0 = (const float4*)s->tv0; a.tv1 = (const float4*)s->tv1; a.tv2 = (const float4*)s->tv2;
	a.nodes = (const float4*)s->nodes;
	a.mesh_first = (const uint32_t*)s->mesh_first;
	a.num_tris = s->num_tris; a.num_meshes = s->num_meshes; a.num_nodes = s->num_nodes;
	a.num_tv = s->num_tv;
	a.abs_max = s->abs_max;
}

static int ensure_scratch(rtkd_scene *s)
{
	// scratch layout: +0 ray cursor | +64 stats[6] | +128 hit counter | +192 sticky error flags
	if (!s->scratch) {
		CK(cudaMalloc((unsigned char**)&s->scratch, 256));
		CK(cudaMemset(s->scratch, 0, 256));
	}
	size_t groups = (size_t)g_sm_count * g_trace_ctas * RTK_TRACE_WARPS * (32 / g_trace_lanes);
	size_t need = (size_t)7 * s->depth + 8;
	size_t entries = need > RTK_STACK_SMEM ? need - RTK_STACK_SMEM : 0;
	if (entries < 8) entries = 8;
	if (!s->overflow || s->overflow_entries < entries || s->overflow_groups < groups) {
		if (s->overflow) cudaFree(s->overflow);
		s->overflow = NULL;
		CK(cudaMalloc((uint2**)&s->overflow, sizeof(uint2) * entries * groups));
		s->overflow_entries = entries; s->overflow_groups = groups;
	}
	return RTKD_OK;
}

// one instantiation per (lanes per ray, cull mode, statistics, any-hit, leaf-phase variant)
template <int L, bool PD>
static void launch_trace_l(int cull_mode, bool stats, unsigned grid, cudaStream_t st, const rtkd_trace_args &p)
{
	const bool c1 = (cull_mode & 1) != 0;
	if (cull_mode & 2) {
		// occlusion query (any hit): the output is a byte per ray
		if (c1) { RTK_LAUNCH((k_trace<L, 1, false, true, PD>), grid, RTK_TRACE_THREADS, st, p); }
		else { RTK_LAUNCH((k_trace<L, 0, false, true, PD>), grid, RTK_TRACE_THREADS, st, p); }
	} else if (stats) {
		if (c1) { RTK_LAUNCH((k_trace<L, 1, true, false, PD>), grid, RTK_TRACE_THREADS, st, p); }
		else { RTK_LAUNCH((k_trace<L, 0, true, false, PD>), grid, RTK_TRACE_THREADS, st, p); }
	} else {
		if (c1) { RTK_LAUNCH((k_trace<L, 1, false, false, PD>), grid, RTK_TRACE_THREADS, st, p); }
		else { RTK_LAUNCH((k_trace<L, 0, false, false, PD>), grid, RTK_TRACE_THREADS, st, p); }
	}
}
static void launch_trace(int lanes, int cull_mode, bool stats, bool pd, unsigned grid, cudaStream_t st, const rtkd_trace_args &p)
{
	if (lanes == 8) launch_trace_l<8, false>(cull_mode, stats, grid, st, p);
	else if (lanes == 4) { if (pd) launch_trace_l<4, true>(cull_mode, stats, grid, st, p); else launch_trace_l<4, false>(cull_mode, stats, grid, st, p); }
	else { if (pd) launch_trace_l<2, true>(cull_mode, stats, grid, st, p); else launch_trace_l<2, false>(cull_mode, stats, grid, st, p); }
}

extern "C" int rtkd_trace(rtkd_scene *s, const void *d_rays, void *d_hit16, size_t n, int cull_mode,
                          rtkd_trace_stats *stats, void *stream)
{
	if (!n) { if (stats) memset(stats, 0, sizeof(*stats)); return RTKD_OK; }
	if (n > 0xfffffff0ull) { rtkd_set_error("batch too large (%zu rays); split it", n); return RTKD_ERR_ARGUMENT; }
	if (((uintptr_t)d_rays & 15) || (!(cull_mode & 2) && ((uintptr_t)d_hit16 & 15))) { rtkd_set_error("device ray / hit buffers must be 16-byte aligned"); return RTKD_ERR_ARGUMENT; }
	if ((cull_mode & 2) && stats) { rtkd_set_error("no statistics variant of the occlusion query"); return RTKD_ERR_ARGUMENT; }
	int r = ensure_scratch(s);
	if (r) return r;
	cudaStream_t st = (cudaStream_t)stream;
	CK(cudaMemsetAsync(s->scratch, 0, 128, st));
	rtkd_trace_args p;
	fill_arrays(s, p.sc);
	p.rays = (const float4*)d_rays; p.out = (float4*)d_hit16; p.nrays = (uint32_t)n;
	p.counter = (uint32_t*)s->scratch; p.err = (uint32_t*)((unsigned char*)s->scratch + 192);
	p.stats = (unsigned long long*)((unsigned char*)s->scratch + 64);
	p.overflow = (uint2*)s->overflow; p.ovf_entries = (uint32_t)s->overflow_entries;
	// persistent grid: one wave of resident CTAs, never more CTAs than ray batches
	size_t batches = (n + RTK_RAY_BATCH - 1) / RTK_RAY_BATCH;
	size_t ctas = (size_t)(g_sm_count - g_reserved_sms) * g_trace_ctas;
	size_t want = (batches + RTK_TRACE_WARPS - 1) / RTK_TRACE_WARPS;
	unsigned grid = (unsigned)(want < ctas ? want : ctas);
	launch_trace(g_trace_lanes, cull_mode, stats != NULL, g_trace_pd != 0, grid, st, p);
	CK_LAUNCH();
	if (stats) {
		unsigned long long h[8];
		uint32_t herr = 0;
		CK(cudaMemcpyAsync(h, p.stats, sizeof(unsigned long long) * 6, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(&herr, p.err, sizeof(herr), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		stats->rays = n; stats->hits = h[1]; stats->node_visits = h[2]; stats->leaf_visits = h[3];
		stats->tri_tests = h[4]; stats->stack_max = h[5];
		if (herr & 2u) { rtkd_set_error("traversal stack exhausted"); return RTKD_ERR_OVERFLOW; }
	}
	return RTKD_OK;
}

extern "C" int rtkd_trace_brute(rtkd_scene *s, const void *d_rays, void *d_hit16, size_t n, void *stream)
{
	if (!n) return RTKD_OK;
	rtkd_arrays a;
	fill_arrays(s, a);
	RTK_LAUNCH(k_trace_brute, (unsigned)((n + 127) / 128), 128, stream, a, (const float4*)d_rays, (float4*)d_hit16, (uint32_t)n);
	CK_LAUNCH();
	return RTKD_OK;
}

extern "C" int rtkd_resolve(rtkd_scene *s, const void *d_hit16, void *d_hits, void *d_mask, size_t n, void *stream)
{
	if (!n) return RTKD_OK;
	int r = ensure_scratch(s);
	if (r) return r;
	rtkd_arrays a;
	fill_arrays(s, a);
	unsigned long long *cnt = (unsigned long long*)((unsigned char*)s->scratch + 128);
	CK(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long), (cudaStream_t)stream));
	RTK_LAUNCH(k_resolve<false>, (unsigned)((n + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS), RTK_RESOLVE_THREADS, stream,
	           a, (const float4*)d_hit16, (uint32_t*)d_hits, (unsigned char*)d_mask, (uint32_t)n, cnt, (uint32_t*)NULL);
	CK_LAUNCH();
	return RTKD_OK;
}

// ---------------------------------------------------------------------------------------------
// hit gather over peer memory: the one exchange step of the sharded path.  The
// gathering process owns a window in its HBM and exports it with CUDA IPC; every other process of
// the box maps it (NVLink peer access is enabled on first use) and pushes its compact hit records
// into its own slice with cudaMemcpyAsync, i.e. with the COPY ENGINES: no send/receive kernel has to
// find room beside the persistent traversal grid, which owns every SM's registers.
// ---------------------------------------------------------------------------------------------

extern "C" int rtkd_peer_create(size_t bytes, void **d_window, unsigned char *handle64)
{
	if (ensure_init()) return RTKD_ERR_NO_DEVICE;
	if (!bytes || !d_window || !handle64) { rtkd_set_error("bad peer window arguments"); return RTKD_ERR_ARGUMENT; }
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
	unsigned char *p = NULL;
	// plain cudaMalloc: stream-ordered pool memory cannot be exported with the legacy IPC calls
	CK(cudaMalloc(&p, bytes));
	cudaIpcMemHandle_t h;
	cudaError_t e = cudaIpcGetMemHandle(&h, p);
	if (e != cudaSuccess) { rtkd_set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); cudaFree(p); return RTKD_ERR_CUDA; }
	memcpy(handle64, &h, 64);
	*d_window = p;
	return RTKD_OK;
}

extern "C" int rtkd_peer_open(const unsigned char *handle64, void **d_window)
{
	if (ensure_init()) return RTKD_ERR_NO_DEVICE;
	if (!handle64 || !d_window) { rtkd_set_error("bad peer window arguments"); return RTKD_ERR_ARGUMENT; }
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, 64);
	void *p = NULL;
	cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
	if (e != cudaSuccess) {
		rtkd_set_error("cudaIpcOpenMemHandle failed: %s (a window cannot be opened by the process that created it, "
		               "and both GPUs must be peers on one box)", cudaGetErrorString(e));
		return RTKD_ERR_CUDA;
	}
	*d_window = p;
	return RTKD_OK;
}

extern "C" int rtkd_peer_close(void *d_window)
{
	if (!d_window) return RTKD_OK;
	CK(cudaIpcCloseMemHandle(d_window));
	return RTKD_OK;
}

extern "C" int rtkd_peer_destroy(void *d_window)
{
	if (!d_window) return RTKD_OK;
	CK(cudaDeviceSynchronize());
	CK(cudaFree(d_window));
	return RTKD_OK;
}

extern "C" int rtkd_peer_push(void *d_dst, const void *d_src, size_t bytes, void *stream)
{
	if (!bytes) return RTKD_OK;
	if (!d_dst || !d_src) { rtkd_set_error("bad peer push arguments"); return RTKD_ERR_ARGUMENT; }
	// unified addressing resolves the owning devices; across GPUs this is an NVLink copy-engine transfer
	CK(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
	return RTKD_OK;
}

// ---------------------------------------------------------------------------------------------
// wavefront ray generation (k_wavefront.cuh)
// ---------------------------------------------------------------------------------------------

extern "C" int rtkd_gen_primary(const float *cam20, uint32_t width, uint32_t height, unsigned long long seed, uint32_t sample,
                                unsigned long long first_pixel, size_t count, void *d_rays, void *stream)
{
	if (!count) return RTKD_OK;
	if (ensure_init()) return RTKD_ERR_NO_DEVICE;
	if (!width || !height || first_pixel + count > (unsigned long long)width * height || count > 0xfffffff0ull) {
		rtkd_set_error("pixel range outside the frame"); return RTKD_ERR_ARGUMENT;
	}
	rtkd_camera cam;
	memcpy(cam.eye, cam20, 12); memcpy(cam.forward, cam20 + 3, 12); memcpy(cam.right, cam20 + 6, 12); memcpy(cam.up, cam20 + 9, 12);
	cam.tan_half_fov = cam20[12]; cam.width = width; cam.height = height;
	RTK_LAUNCH(k_gen_primary, (unsigned)((count + 255) / 256), 256, stream, cam, seed, sample, first_pixel, (uint32_t)count, (float4*)d_rays);
	CK_LAUNCH();
	return RTKD_OK;
}

extern "C" int rtkd_gen_bounce(rtkd_scene *s, const void *d_rays_in, const void *d_hit16, void *d_rays_out, void *d_alive,
                               size_t n, unsigned long long seed, uint32_t bounce, unsigned long long first_ray, uint32_t flags, void *stream)
{
	if (!n) return RTKD_OK;
	if (n > 0xfffffff0ull) { rtkd_set_error("batch too large"); return RTKD_ERR_ARGUMENT; }
	rtkd_arrays a;
	fill_arrays(s, a);
	// push the new origin off the surface by 2^-13 of the scene's largest coordinate
	float push = s->abs_max * 1.220703125e-4f;
	RTK_LAUNCH(k_gen_bounce, (unsigned)((n + 255) / 256), 256, stream, a, (const float4*)d_rays_in, (const float4*)d_hit16,
	           (float4*)d_rays_out, (unsigned char*)d_alive, (uint32_t)n, seed, bounce, first_ray, push, flags);
	CK_LAUNCH();
	return RTKD_OK;
}

extern "C" void *rtkd_scene_hit16(rtkd_scene *s, size_t n)
{
	if (s->hit16_cap < n) {
		if (s->hit16) cudaFree(s->hit16);
		s->hit16 = NULL; s->hit16_cap = 0;
		size_t cap = n + n / 8 + 1024;
		CKP(cudaMalloc((float4**)&s->hit16, sizeof(float4) * cap));
		s->hit16_cap = cap;
	}
	return s->hit16;
}

// Host-buffer batch: a three-stage pipeline over chunks of RTKD_HOST_CHUNK rays, RTKD_HOST_BUFS chunks
// in flight on their own streams.
//   upload             the rays of the whole batch go to one device buffer, chunk by chunk on a
//                      copy stream of their own that runs RTKD_HOST_AHEAD chunks ahead of the
//                      kernels, so the H2D engine -- the bottleneck resource of the batch: 32 bytes
//                      per ray up against ~26 down -- never waits for the host
//   stage A (enqueue)  k_trace -> k_resolve<dense> -> D2H of the chunk's mask bytes, block bases
//                      and hit count into pinned staging
//   stage B            once the count is known: D2H of exactly the rows of the rays that hit
//   stage C            the host worker pool (rtk_place.c) copies each row to hits[i]
// Only hit rows cross PCIe (68 bytes per HIT plus 1 byte per ray instead of 69 bytes per ray), and
// rows of rays that missed are left untouched in the caller's array, the reference's miss rule
// (rtk.c:571-576).  The traversal scratch (ray cursor, overflow slab) belongs to the scene, so the
// TRAVERSALS of consecutive chunks are serialised with an event; everything else overlaps.
#define RTKD_HOST_CHUNK ((size_t)1 << 20)
#define RTKD_HOST_BUFS 4
#define RTKD_HOST_AHEAD 4
#define RTKD_HOST_RING 8             // upload events: more than RTKD_HOST_AHEAD + 1
#define RTKD_HOST_SMALL ((size_t)2048)  // batches up to this size take the one-stream, one-synchronisation path

struct host_buf {
	cudaStream_t st;
	cudaEvent_t traced, meta_done, rows_done;
	float4 *d_h16;
	uint32_t *d_rows, *d_base;          // d_base: [blocks] block bases, then the 64-bit hit count
	unsigned char *d_mask;
	unsigned char *h_meta;              // pinned: mask bytes | block bases | hit count
	unsigned char *h_rows;              // pinned: dense rows
	size_t off, cnt, hits;              // the chunk this buffer currently carries
	int ticket, state;                  // state: 0 free, 1 stage A queued, 2 stage B queued, 3 placing
};
struct host_stage {
	size_t chunk, blocks, meta_bytes;
	host_buf b[RTKD_HOST_BUFS];
	cudaStream_t up;                    // upload stream
	cudaEvent_t uploaded[RTKD_HOST_RING];
	float4 *d_rays; size_t rays_cap;    // the whole batch's rays (grow-only)
	bool ready;                         // buffers allocated for `chunk`
	bool streams;                       // streams and events exist (they outlive a change of chunk size)
};
static host_stage g_stage;
static pthread_mutex_t g_stage_lock = PTHREAD_MUTEX_INITIALIZER;

static void stage_release(void)
{
	for (int k = 0; k < RTKD_HOST_BUFS; k++) {
		host_buf &B = g_stage.b[k];
		cudaFree(B.d_h16); cudaFree(B.d_rows); cudaFree(B.d_base); cudaFree(B.d_mask);
		cudaFreeHost(B.h_meta); cudaFreeHost(B.h_rows);
		B.d_h16 = NULL; B.d_rows = B.d_base = NULL; B.d_mask = NULL; B.h_meta = B.h_rows = NULL;
	}
}

static void stage_shutdown(void)
{
	pthread_mutex_lock(&g_stage_lock);
	if (g_stage.ready || g_stage.streams) {
		stage_release();
		for (int k = 0; k < RTKD_HOST_BUFS; k++) {
			host_buf &B = g_stage.b[k];
			if (B.st) cudaStreamDestroy(B.st);
			if (B.traced) cudaEventDestroy(B.traced);
			if (B.meta_done) cudaEventDestroy(B.meta_done);
			if (B.rows_done) cudaEventDestroy(B.rows_done);
		}
		if (g_stage.up) cudaStreamDestroy(g_stage.up);
		for (int k = 0; k < RTKD_HOST_RING; k++) if (g_stage.uploaded[k]) cudaEventDestroy(g_stage.uploaded[k]);
	}
	if (g_stage.d_rays) cudaFree(g_stage.d_rays);
	memset(&g_stage, 0, sizeof(g_stage));
	pthread_mutex_unlock(&g_stage_lock);
}

static int stage_prepare(size_t want)
{
	size_t max_chunk = RTKD_HOST_CHUNK;
	bool forced = false;
	{
		// experiment / test knob: chunk size 2^k rays (the tests use it to run many small chunks)
		const char *e = getenv("RTK_B200_HOST_CHUNK_LOG2");
		if (e && atoi(e) >= 12 && atoi(e) <= 24) { max_chunk = (size_t)1 << atoi(e); forced = true; }
	}
	size_t chunk = want < max_chunk ? want : max_chunk;
	if (chunk < 4096) chunk = 4096;
	if (g_stage.ready && (forced ? g_stage.chunk == chunk : g_stage.chunk >= chunk)) return RTKD_OK;
	stage_release();                     // also what an earlier, failed attempt left behind
	if (!g_stage.streams) {
		for (int k = 0; k < RTKD_HOST_BUFS; k++) {
			host_buf &B = g_stage.b[k];
			CK(cudaStreamCreateWithFlags(&B.st, cudaStreamNonBlocking));
			CK(cudaEventCreateWithFlags(&B.traced, cudaEventDisableTiming));
			CK(cudaEventCreateWithFlags(&B.meta_done, cudaEventDisableTiming));
			CK(cudaEventCreateWithFlags(&B.rows_done, cudaEventDisableTiming));
		}
		CK(cudaStreamCreateWithFlags(&g_stage.up, cudaStreamNonBlocking));
		for (int k = 0; k < RTKD_HOST_RING; k++) CK(cudaEventCreateWithFlags(&g_stage.uploaded[k], cudaEventDisableTiming));
		g_stage.streams = true;
	}
	g_stage.ready = false;
	g_stage.chunk = chunk;
	g_stage.blocks = ((chunk + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS + 3) & ~(size_t)3;   // the 64-bit count follows: keep it aligned
	g_stage.meta_bytes = ((chunk + 15) & ~(size_t)15) + 4 * g_stage.blocks + 16;
	for (int k = 0; k < RTKD_HOST_BUFS; k++) {
		host_buf &B = g_stage.b[k];
		CK(cudaMalloc(&B.d_h16, 16 * chunk));
		CK(cudaMalloc(&B.d_rows, 68 * chunk));
		CK(cudaMalloc(&B.d_base, 4 * g_stage.blocks + 16));
		CK(cudaMalloc(&B.d_mask, (chunk + 15) & ~(size_t)15));
		CK(cudaMallocHost(&B.h_meta, g_stage.meta_bytes));
		CK(cudaMallocHost(&B.h_rows, 68 * chunk));
		B.state = 0; B.ticket = -1;
	}
	g_stage.ready = true;
	return RTKD_OK;
}

// stage B of one buffer: wait for the chunk's count, then fetch exactly its rows
static int host_stage_b(host_buf &B)
{
	CK(cudaEventSynchronize(B.meta_done));
	const size_t mask_bytes = (g_stage.chunk + 15) & ~(size_t)15;
	unsigned long long hc = 0;
	memcpy(&hc, B.h_meta + mask_bytes + 4 * g_stage.blocks, sizeof(hc));
	B.hits = (size_t)hc;
	if (B.hits) CK(cudaMemcpyAsync(B.h_rows, B.d_rows, 68 * B.hits, cudaMemcpyDeviceToHost, B.st));
	CK(cudaEventRecord(B.rows_done, B.st));
	B.state = 2;
	return RTKD_OK;
}

// stage C: hand the rows to the placement workers
static int host_stage_c(host_buf &B, void *hits, unsigned char *mask)
{
	CK(cudaEventSynchronize(B.rows_done));
	const size_t mask_bytes = (g_stage.chunk + 15) & ~(size_t)15;
	rtkd_place_desc d;
	d.hits = hits; d.mask_out = mask;
	d.rows = B.h_rows; d.mask = B.h_meta; d.block_base = (const uint32_t*)(B.h_meta + mask_bytes);
	d.first_ray = B.off; d.nrays = B.cnt;
	B.ticket = rtkd_place_submit(&d);
	B.state = 3;
	return RTKD_OK;
}

extern "C" long long rtkd_trace_host(rtkd_scene *s, const void *rays, void *hits, unsigned char *mask, size_t n)
{
	if (!n) return 0;
	pthread_mutex_lock(&g_stage_lock);
	long long total = 0;
	int rc = stage_prepare(n);
	if (rc == RTKD_OK) rc = ensure_scratch(s);
	host_stage &G = g_stage;
	if (rc == RTKD_OK && G.rays_cap < n) {
		if (G.d_rays) cudaFree(G.d_rays);
		G.d_rays = NULL; G.rays_cap = 0;
		if (cudaMalloc(&G.d_rays, 32 * n) != cudaSuccess) { rtkd_set_error("out of device memory for %zu rays", n); rc = RTKD_ERR_MEMORY; }
		else G.rays_cap = n;
	}
	const size_t chunk = G.chunk;
	const size_t nchunks = (n + chunk - 1) / chunk;
	const size_t mask_bytes = (chunk + 15) & ~(size_t)15;
	rtkd_arrays a;
	fill_arrays(s, a);
	if (rc == RTKD_OK && n <= RTKD_HOST_SMALL) {
		// Small batches -- rtk_trace_ray is a batch of one -- are bound by latency, not by bytes: one
		// stream, rows expanded in place (no dense packing, no worker threads), ONE synchronisation.
		host_buf &B = G.b[0];
		cudaError_t e = cudaMemcpyAsync(G.d_rays, rays, 32 * n, cudaMemcpyHostToDevice, B.st);
		if (e == cudaSuccess) rc = rtkd_trace(s, G.d_rays, B.d_h16, n, 1, NULL, B.st);
		if (e == cudaSuccess && rc == RTKD_OK) {
			RTK_LAUNCH(k_resolve<false>, (unsigned)((n + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS), RTK_RESOLVE_THREADS, B.st,
			           a, (const float4*)B.d_h16, B.d_rows, B.d_mask, (uint32_t)n, (unsigned long long*)NULL, (uint32_t*)NULL);
			e = cudaGetLastError();
			if (e == cudaSuccess) e = cudaMemcpyAsync(B.h_rows, B.d_rows, 68 * n, cudaMemcpyDeviceToHost, B.st);
			if (e == cudaSuccess) e = cudaMemcpyAsync(B.h_meta, B.d_mask, n, cudaMemcpyDeviceToHost, B.st);
			if (e == cudaSuccess) e = cudaMemcpyAsync(B.h_meta + mask_bytes, (unsigned char*)s->scratch + 192, 4, cudaMemcpyDeviceToHost, B.st);
			if (e == cudaSuccess) e = cudaStreamSynchronize(B.st);
		}
		if (e != cudaSuccess) { rtkd_set_error("rtk_trace_rays: %s", cudaGetErrorString(e)); rc = RTKD_ERR_CUDA; }
		if (rc == RTKD_OK) {
			uint32_t herr = 0;
			memcpy(&herr, B.h_meta + mask_bytes, 4);
			if (herr & 2u) { rtkd_set_error("traversal stack exhausted"); rc = RTKD_ERR_OVERFLOW; }
		}
		if (rc == RTKD_OK) {
			for (size_t i = 0; i < n; i++) {
				if (B.h_meta[i]) { memcpy((char*)hits + 68 * i, B.h_rows + 68 * i, 68); total++; }    // rows of misses stay untouched (rtk.c:571-576)
			}
			if (mask) memcpy(mask, B.h_meta, n);
		}
		pthread_mutex_unlock(&g_stage_lock);
		return rc == RTKD_OK ? total : -1;
	}
	cudaEvent_t prev_traced = NULL;
	size_t uploads = 0;                 // chunks whose upload has been enqueued
	// chunk ci enters stage A in iteration ci, stage B in iteration ci+1, stage C in iteration ci+2
	// and its buffer is reused in iteration ci + RTKD_HOST_BUFS
	for (size_t it = 0; it < nchunks + 2 && rc == RTKD_OK; it++) {
		// keep the upload stream RTKD_HOST_AHEAD chunks ahead of the kernels
		for (; uploads < nchunks && uploads <= it + RTKD_HOST_AHEAD; uploads++) {
			const size_t off = uploads * chunk, cnt = n - off < chunk ? n - off : chunk;
			cudaError_t e = cudaMemcpyAsync(G.d_rays + 2 * off, (const char*)rays + 32 * off, 32 * cnt, cudaMemcpyHostToDevice, G.up);
			if (e == cudaSuccess) e = cudaEventRecord(G.uploaded[uploads % RTKD_HOST_RING], G.up);
			if (e != cudaSuccess) { rtkd_set_error("rtk_trace_rays: %s", cudaGetErrorString(e)); rc = RTKD_ERR_CUDA; break; }
		}
		if (rc != RTKD_OK) break;
		if (it < nchunks) {
			host_buf &B = G.b[it % RTKD_HOST_BUFS];
			if (B.state == 3) { rtkd_place_wait(B.ticket); total += (long long)B.hits; B.state = 0; }
			B.off = it * chunk; B.cnt = n - B.off < chunk ? n - B.off : chunk;
			unsigned long long *d_count = (unsigned long long*)((unsigned char*)B.d_base + 4 * G.blocks);
			cudaError_t e = cudaStreamWaitEvent(B.st, G.uploaded[it % RTKD_HOST_RING], 0);
			if (e == cudaSuccess) e = cudaMemsetAsync(d_count, 0, 16, B.st);
			if (e == cudaSuccess && prev_traced) e = cudaStreamWaitEvent(B.st, prev_traced, 0);
			if (e != cudaSuccess) { rtkd_set_error("rtk_trace_rays: %s", cudaGetErrorString(e)); rc = RTKD_ERR_CUDA; break; }
			rc = rtkd_trace(s, G.d_rays + 2 * B.off, B.d_h16, B.cnt, 1, NULL, B.st);
			if (rc) break;
			cudaEventRecord(B.traced, B.st);
			prev_traced = B.traced;
			const unsigned blocks = (unsigned)((B.cnt + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS);
			RTK_LAUNCH(k_resolve<true>, blocks, RTK_RESOLVE_THREADS, B.st, a, (const float4*)B.d_h16, B.d_rows, B.d_mask, (uint32_t)B.cnt, d_count, B.d_base);
			e = cudaGetLastError();
			// mask bytes and bases+count land in one pinned block: [mask | bases | count]
			if (e == cudaSuccess) e = cudaMemcpyAsync(B.h_meta, B.d_mask, B.cnt, cudaMemcpyDeviceToHost, B.st);
			if (e == cudaSuccess) e = cudaMemcpyAsync(B.h_meta + mask_bytes, B.d_base, 4 * G.blocks + 16, cudaMemcpyDeviceToHost, B.st);
			if (e == cudaSuccess) e = cudaEventRecord(B.meta_done, B.st);
			if (e != cudaSuccess) { rtkd_set_error("rtk_trace_rays: %s", cudaGetErrorString(e)); rc = RTKD_ERR_CUDA; break; }
			B.state = 1;
		}
		if (it >= 1 && it - 1 < nchunks) rc = host_stage_b(G.b[(it - 1) % RTKD_HOST_BUFS]);
		if (rc == RTKD_OK && it >= 2 && it - 2 < nchunks) rc = host_stage_c(G.b[(it - 2) % RTKD_HOST_BUFS], hits, mask);
	}
	// drain: placements still running, and -- after an error -- whatever is still queued
	for (int k = 0; k < RTKD_HOST_BUFS; k++) {
		host_buf &B = G.b[k];
		if (B.state == 3) { rtkd_place_wait(B.ticket); total += (long long)B.hits; }
		else if (B.state) cudaStreamSynchronize(B.st);
		B.state = 0; B.ticket = -1;
	}
	if (rc != RTKD_OK) cudaStreamSynchronize(G.up);
	if (rc == RTKD_OK) {
		uint32_t herr = 0;
		if (cudaMemcpy(&herr, (unsigned char*)s->scratch + 192, sizeof(herr), cudaMemcpyDeviceToHost) != cudaSuccess) rc = RTKD_ERR_CUDA;
		if (herr & 2u) { rtkd_set_error("traversal stack exhausted"); rc = RTKD_ERR_OVERFLOW; }
	}
	if (rc == RTKD_ERR_CUDA && !g_err[0]) rtkd_set_error("CUDA failure in rtk_trace_rays");
	pthread_mutex_unlock(&g_stage_lock);
	return rc == RTKD_OK ? total : -1;
}

// Host-buffer batch with COMPACT results: rays up, one 16-byte record (t, u, v, global triangle
// number or RTKD miss) per ray down, straight into the caller's array -- no dense packing, no row
// placement by host threads.  32 bytes per ray go up and 16 come down, on the two directions of the
// link, so the batch is bound by the upload alone.  Same staging, same chunking, same upload stream
// running ahead as rtkd_trace_host; each chunk's stream carries k_trace and the copy of its records.
extern "C" int rtkd_trace_host_compact(rtkd_scene *s, const void *rays, void *hit16, size_t n)
{
	if (!n) return RTKD_OK;
	pthread_mutex_lock(&g_stage_lock);
	int rc = stage_prepare(n);
	if (rc == RTKD_OK) rc = ensure_scratch(s);
	host_stage &G = g_stage;
	if (rc == RTKD_OK && G.rays_cap < n) {
		if (G.d_rays) cudaFree(G.d_rays);
		G.d_rays = NULL; G.rays_cap = 0;
		if (cudaMalloc(&G.d_rays, 32 * n) != cudaSuccess) { rtkd_set_error("out of device memory for %zu rays", n); rc = RTKD_ERR_MEMORY; }
		else G.rays_cap = n;
	}
	if (rc != RTKD_OK) { pthread_mutex_unlock(&g_stage_lock); return rc; }
	const size_t chunk = G.chunk;
	const size_t nchunks = (n + chunk - 1) / chunk;
	cudaEvent_t prev_traced = NULL;
	size_t uploads = 0;
	for (size_t it = 0; it < nchunks && rc == RTKD_OK; it++) {
		for (; uploads < nchunks && uploads <= it + RTKD_HOST_AHEAD; uploads++) {
			const size_t off = uploads * chunk, cnt = n - off < chunk ? n - off : chunk;
			cudaError_t e = cudaMemcpyAsync(G.d_rays + 2 * off, (const char*)rays + 32 * off, 32 * cnt, cudaMemcpyHostToDevice, G.up);
			if (e == cudaSuccess) e = cudaEventRecord(G.uploaded[uploads % RTKD_HOST_RING], G.up);
			if (e != cudaSuccess) { rtkd_set_error("rtk_trace_rays_compact: %s", cudaGetErrorString(e)); rc = RTKD_ERR_CUDA; break; }
		}
		if (rc != RTKD_OK) break;
		// the chunk's stream still holds the copy of the chunk that used this buffer before: stream
		// order keeps the new traversal from overwriting records that have not left yet
		host_buf &B = G.b[it % RTKD_HOST_BUFS];
		const size_t off = it * chunk, cnt = n - off < chunk ? n - off : chunk;
		cudaError_t e = cudaStreamWaitEvent(B.st, G.uploaded[it % RTKD_HOST_RING], 0);
		if (e == cudaSuccess && prev_traced) e = cudaStreamWaitEvent(B.st, prev_traced, 0);   // the traversal scratch is the scene's
		if (e != cudaSuccess) { rtkd_set_error("rtk_trace_rays_compact: %s", cudaGetErrorString(e)); rc = RTKD_ERR_CUDA; break; }
		rc = rtkd_trace(s, G.d_rays + 2 * off, B.d_h16, cnt, 1, NULL, B.st);
		if (rc) break;
		e = cudaEventRecord(B.traced, B.st);
		prev_traced = B.traced;
		if (e == cudaSuccess) e = cudaMemcpyAsync((char*)hit16 + 16 * off, B.d_h16, 16 * cnt, cudaMemcpyDeviceToHost, B.st);
		if (e != cudaSuccess) { rtkd_set_error("rtk_trace_rays_compact: %s", cudaGetErrorString(e)); rc = RTKD_ERR_CUDA; break; }
	}
	for (int k = 0; k < RTKD_HOST_BUFS; k++) if (cudaStreamSynchronize(G.b[k].st) != cudaSuccess && rc == RTKD_OK) rc = RTKD_ERR_CUDA;
	if (cudaStreamSynchronize(G.up) != cudaSuccess && rc == RTKD_OK) rc = RTKD_ERR_CUDA;
	if (rc == RTKD_OK) {
		uint32_t herr = 0;
		if (cudaMemcpy(&herr, (unsigned char*)s->scratch + 192, sizeof(herr), cudaMemcpyDeviceToHost) != cudaSuccess) rc = RTKD_ERR_CUDA;
		if (herr & 2u) { rtkd_set_error("traversal stack exhausted"); rc = RTKD_ERR_OVERFLOW; }
	}
	if (rc == RTKD_ERR_CUDA && !g_err[0]) rtkd_set_error("CUDA failure in rtk_trace_rays_compact");
	pthread_mutex_unlock(&g_stage_lock);
	return rc;
}

// ---------------------------------------------------------------------------------------------
// serialisation
// ---------------------------------------------------------------------------------------------

struct rtkd_blob_sub {          // 128 bytes, first thing in the payload
	uint64_t magic2;            // "B200RTK3"
	uint64_t id;
	uint32_t num_tris, num_meshes, num_nodes, num_leaves, depth, build_mode;
	float bounds_min[3], bounds_max[3], abs_max;
	uint32_t num_tv;            // entries of each traversal triangle array (8 per leaf)
	double sah_cost;
	uint64_t off_nodes, off_tv0, off_tv1, off_tv2, off_orig, off_mesh;   // from payload start
};

static size_t a128(size_t v) { return (v + 127) & ~(size_t)127; }

static void blob_layout(const rtkd_scene *s, rtkd_blob_sub *b)
{
	size_t o = a128(sizeof(rtkd_blob_sub));
	b->off_nodes = o; o = a128(o + 256 * (size_t)s->num_nodes);
	b->off_tv0 = o;   o = a128(o + 16 * (size_t)s->num_tv);
	b->off_tv1 = o;   o = a128(o + 16 * (size_t)s->num_tv);
	b->off_tv2 = o;   o = a128(o + 16 * (size_t)s->num_tv);
	b->off_orig = o;  o = a128(o + 48 * (size_t)s->num_tris);
	b->off_mesh = o;
}

extern "C" size_t rtkd_blob_payload_size(const rtkd_scene *s)
{
	rtkd_blob_sub b;
	blob_layout(s, &b);
	return a128((size_t)b.off_mesh + 4 * ((size_t)s->num_meshes + 1));
}

extern "C" int rtkd_blob_write(const rtkd_scene *s, void *payload)
{
	rtkd_blob_sub b;
	memset(&b, 0, sizeof(b));
	memcpy(&b.magic2, "B200RTK3", 8);
	b.id = s->id;
	b.num_tris = s->num_tris; b.num_meshes = s->num_meshes; b.num_nodes = s->num_nodes;
	b.num_leaves = s->num_leaves; b.depth = s->depth; b.build_mode = s->build_mode;
	memcpy(b.bounds_min, s->bounds_min, 12); memcpy(b.bounds_max, s->bounds_max, 12);
	b.abs_max = s->abs_max; b.sah_cost = s->sah_cost; b.num_tv = s->num_tv;
	blob_layout(s, &b);
	char *p = (char*)payload;
	memset(p, 0, a128(sizeof(b)));
	memcpy(p, &b, sizeof(b));
	CK(cudaDeviceSynchronize());
	if (s->num_nodes) CK(cudaMemcpy(p + b.off_nodes, s->nodes, 256 * (size_t)s->num_nodes, cudaMemcpyDeviceToHost));
	if (s->num_tris) {
		CK(cudaMemcpy(p + b.off_tv0, s->tv0, 16 * (size_t)s->num_tv, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(p + b.off_tv1, s->tv1, 16 * (size_t)s->num_tv, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(p + b.off_tv2, s->tv2, 16 * (size_t)s->num_tv, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(p + b.off_orig, s->tri_orig, 48 * (size_t)s->num_tris, cudaMemcpyDeviceToHost));
	}
	memcpy(p + b.off_mesh, s->h_mesh_first, 4 * ((size_t)s->num_meshes + 1));
	return RTKD_OK;
}

extern "C" rtkd_scene *rtkd_blob_read(const void *payload, size_t payload_size)
{
	if (ensure_init()) return NULL;
	rtkd_blob_sub b;
	if (payload_size < sizeof(b)) { rtkd_set_error("scene blob truncated"); return NULL; }
	memcpy(&b, payload, sizeof(b));
	if (memcmp(&b.magic2, "B200RTK3", 8) != 0) { rtkd_set_error("blob was not written by rtk_b200 (device-layout magic missing)"); return NULL; }
	if (b.off_mesh + 4 * ((size_t)b.num_meshes + 1) > payload_size) { rtkd_set_error("scene blob truncated"); return NULL; }
	const char *p = (const char*)payload;
	rtkd_scene *s = rtkd_scene_new(b.num_tris, b.num_meshes, (const uint32_t*)(p + b.off_mesh));
	if (!s) return NULL;
	s->id = b.id;
	s->num_nodes = b.num_nodes; s->num_leaves = b.num_leaves; s->depth = b.depth; s->build_mode = b.build_mode;
	memcpy(s->bounds_min, b.bounds_min, 12); memcpy(s->bounds_max, b.bounds_max, 12);
	s->abs_max = b.abs_max; s->sah_cost = b.sah_cost; s->num_tv = b.num_tv; s->tv_cap = b.num_tv;
	cudaError_t e = cudaSuccess;
	if (b.num_tris) {
		e = cudaMemcpy(s->tri_orig, p + b.off_orig, 48 * (size_t)b.num_tris, cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMalloc((float4**)&s->tv0, 16 * (size_t)b.num_tv);
		if (e == cudaSuccess) e = cudaMalloc((float4**)&s->tv1, 16 * (size_t)b.num_tv);
		if (e == cudaSuccess) e = cudaMalloc((float4**)&s->tv2, 16 * (size_t)b.num_tv);
		if (e == cudaSuccess) e = cudaMemcpy(s->tv0, p + b.off_tv0, 16 * (size_t)b.num_tv, cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(s->tv1, p + b.off_tv1, 16 * (size_t)b.num_tv, cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(s->tv2, p + b.off_tv2, 16 * (size_t)b.num_tv, cudaMemcpyHostToDevice);
	}
	if (e == cudaSuccess && b.num_nodes) {
		e = cudaMalloc((float4**)&s->nodes, 256 * (size_t)b.num_nodes);
		s->nodes_cap = b.num_nodes;
		if (e == cudaSuccess) e = cudaMemcpy(s->nodes, p + b.off_nodes, 256 * (size_t)b.num_nodes, cudaMemcpyHostToDevice);
	}
	if (e != cudaSuccess) {
		rtkd_set_error("scene upload failed: %s", cudaGetErrorString(e));
		rtkd_scene_free(s);
		return NULL;
	}
	return s;
}
