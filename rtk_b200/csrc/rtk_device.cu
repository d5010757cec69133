// rtk_device.cu -- the CUDA translation unit: kernels (k_build.cuh, k_trace.cuh) plus the thin
// C-ABI layer (rtk_device.h) that the C host code calls.  Compiled with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
// There is no CPU implementation behind these entry points: without a CUDA device every call
// fails with RTK_CUDA_ERR_NO_DEVICE.
#include "rtk_device.h"
#include "k_build.cuh"
#include "k_sah.cuh"
#include "k_trace.cuh"
#include <stdarg.h>

#define RTKD_OK 0
#define RTKD_ERR_NO_DEVICE (-1)
#define RTKD_ERR_CUDA (-2)
#define RTKD_ERR_ARGUMENT (-3)
#define RTKD_ERR_SCENE (-4)
#define RTKD_ERR_MEMORY (-5)
#define RTKD_ERR_OVERFLOW (-6)

static __thread char g_err[512];
static int g_device = -1;
static int g_sm_count = 0, g_trace_ctas = 0;
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

extern "C" int rtkd_init(int device)
{
	if (g_device == device && g_sm_count) return RTKD_OK;
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
	int ctas = 0;
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<1, false>, RTK_TRACE_THREADS, 0));
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

static int ensure_init(void) { return g_sm_count ? RTKD_OK : rtkd_init(0); }

extern "C" void rtkd_shutdown(void) { g_device = -1; g_sm_count = 0; }

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
// scene lifetime
// ---------------------------------------------------------------------------------------------

extern "C" rtkd_scene *rtkd_scene_new(uint32_t num_tris, uint32_t num_meshes, const uint32_t *mesh_first)
{
	if (ensure_init()) return NULL;
	rtkd_scene *s = (rtkd_scene*)calloc(1, sizeof(rtkd_scene));
	if (!s) { rtkd_set_error("out of host memory"); return NULL; }
	s->id = ((uint64_t)time(NULL) << 20) ^ (g_next_id++ * 0x9E3779B97F4A7C15ull);
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
	cudaDeviceSynchronize();
	if (s->tri_orig) cudaFree(s->tri_orig);
	if (s->tv0) cudaFree(s->tv0);
	if (s->tv1) cudaFree(s->tv1);
	if (s->tv2) cudaFree(s->tv2);
	if (s->nodes) cudaFree(s->nodes);
	if (s->mesh_first) cudaFree(s->mesh_first);
	if (s->scratch) cudaFree(s->scratch);
	if (s->overflow) cudaFree(s->overflow);
	if (s->hit16) cudaFree(s->hit16);
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
	if (!ntris) return RTKD_OK;
	if ((size_t)first_prim + ntris > s->num_tris) { rtkd_set_error("decode range outside the scene"); return RTKD_ERR_ARGUMENT; }
	rtkd_decode_args a;
	a.pos = (const unsigned char*)d_pos; a.idx = (const unsigned char*)d_idx;
	a.pos_stride = pos_stride; a.idx_stride = idx_stride;
	a.pos_f64 = pos_f64; a.idx_bytes = d_idx ? idx_bytes : 0; a.pregathered = pregathered;
	a.ntris = ntris; a.first_prim = first_prim;
	RTK_LAUNCH(k_decode_mesh, (ntris + 255) / 256, 256, stream, a, (float4*)s->tri_orig);
	CK_LAUNCH();
	return RTKD_OK;
}

// ---------------------------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------------------------

template <typename T> static cudaError_t tmp_alloc(T **p, size_t count, cudaStream_t st)
{
	return cudaMallocAsync(p, sizeof(T) * (count ? count : 1), st);
}

// ---- binned-SAH binary tree (k_sah.cuh) -------------------------------------------------------

static void free_sah(rtkd_sah &h, uint32_t *order, cudaStream_t st)
{
	cudaFreeAsync((void*)h.pb, st); cudaFreeAsync(h.idx0, st); cudaFreeAsync(h.idx1, st); cudaFreeAsync(h.idx_final, st);
	cudaFreeAsync(h.left, st); cudaFreeAsync(h.right, st); cudaFreeAsync(h.first, st); cudaFreeAsync(h.last, st);
	cudaFreeAsync(h.blo, st); cudaFreeAsync(h.bhi, st); cudaFreeAsync(h.ndepth, st); cudaFreeAsync(h.counters, st);
	cudaFreeAsync(h.act_in, st); cudaFreeAsync(h.act_out, st); cudaFreeAsync(h.small_list, st);
	cudaFreeAsync(h.chunk_base, st); cudaFreeAsync(h.bins, st); cudaFreeAsync(h.split, st); cudaFreeAsync(h.cursor, st);
	cudaFreeAsync(order, st);
}

static int build_sah(rtkd_scene *s, cudaStream_t st, const float4 *tri, const uint32_t *svals, const uint32_t *d_bounds,
                     uint32_t n, rtkd_sah &h, uint32_t **order_out)
{
	(void)s;
	const size_t cap = 2 * (size_t)n + 2;
	const size_t act_cap = n / RTK_SAH_SMALL + 4;
	const size_t small_cap = 4 * (size_t)(n / RTK_SAH_SMALL) + 8;
	float4 *pb = NULL;
	uint32_t *order = NULL;
	CK(tmp_alloc(&pb, 2 * (size_t)n, st)); h.pb = pb;
	CK(tmp_alloc(&h.idx0, n, st)); CK(tmp_alloc(&h.idx1, n, st)); CK(tmp_alloc(&h.idx_final, n, st));
	CK(tmp_alloc(&h.left, cap, st)); CK(tmp_alloc(&h.right, cap, st)); CK(tmp_alloc(&h.first, cap, st)); CK(tmp_alloc(&h.last, cap, st));
	CK(tmp_alloc(&h.blo, cap, st)); CK(tmp_alloc(&h.bhi, cap, st)); CK(tmp_alloc(&h.ndepth, cap, st));
	CK(tmp_alloc(&h.counters, 8, st));
	CK(tmp_alloc(&h.act_in, act_cap, st)); CK(tmp_alloc(&h.act_out, act_cap, st)); CK(tmp_alloc(&h.small_list, small_cap, st));
	CK(tmp_alloc(&h.chunk_base, act_cap + 1, st));
	CK(tmp_alloc(&h.bins, act_cap * RTK_SAH_NODEBINS, st));
	CK(tmp_alloc(&h.split, act_cap, st)); CK(tmp_alloc(&h.cursor, 2 * act_cap, st));
	CK(tmp_alloc(&order, n, st));
	h.node_cap = (uint32_t)cap;

	RTK_LAUNCH(k_sah_prim_bounds, (n + 255) / 256, 256, st, tri, svals, n, pb, h.idx0); CK_LAUNCH();
	RTK_LAUNCH(k_sah_root, 1, 32, st, h, d_bounds, n); CK_LAUNCH();
	uint32_t hc[8];
	CK(cudaMemcpyAsync(hc, h.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	uint32_t n_act = hc[1], depth = 0;
	int src_buf = 0;
	while (n_act) {
		if (n_act > act_cap) { rtkd_set_error("SAH active list overflow"); return RTKD_ERR_MEMORY; }
		RTK_LAUNCH(k_sah_plan, 1, 1024, st, h, n_act); CK_LAUNCH();
		uint32_t chunks = 0;
		CK(cudaMemcpyAsync(&chunks, h.chunk_base + n_act, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
		CK(cudaMemsetAsync(h.counters + 1, 0, sizeof(uint32_t), st));
		RTK_LAUNCH(k_sah_bins_clear, n_act, 128, st, h, n_act); CK_LAUNCH();
		CK(cudaStreamSynchronize(st));
		RTK_LAUNCH(k_sah_bin_large, chunks, 256, st, h, n_act, src_buf); CK_LAUNCH();
		RTK_LAUNCH(k_sah_split_large, (n_act + 3) / 4, 128, st, h, n_act, depth, src_buf ^ 1); CK_LAUNCH();
		RTK_LAUNCH(k_sah_partition_large, chunks, 256, st, h, n_act, src_buf); CK_LAUNCH();
		CK(cudaMemcpyAsync(hc, h.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		uint32_t *tmp = h.act_in; h.act_in = h.act_out; h.act_out = tmp;
		n_act = hc[1];
		src_buf ^= 1;
		depth++;
	}
	CK(cudaMemcpyAsync(hc, h.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	if (hc[2] > small_cap) { rtkd_set_error("SAH small-subtree list overflow"); return RTKD_ERR_MEMORY; }
	if (hc[2]) { RTK_LAUNCH(k_sah_small, hc[2], RTK_SAH_SMALL_THREADS, st, h, hc[2]); CK_LAUNCH(); }
	RTK_LAUNCH(k_sah_compose, (n + 255) / 256, 256, st, (const uint32_t*)h.idx_final, svals, n, order); CK_LAUNCH();
	CK(cudaMemcpyAsync(hc, h.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	if (hc[3]) { rtkd_set_error("SAH node pool exhausted"); return RTKD_ERR_MEMORY; }
	*order_out = order;
	return RTKD_OK;
}

extern "C" int rtkd_build(rtkd_scene *s, int mode, void *stream)
{
	cudaStream_t st = (cudaStream_t)stream;
	const uint32_t n = s->num_tris;
	s->build_mode = (uint32_t)mode;
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	CK(cudaEventRecord(e0, st));

	// (re)build: drop the previous traversal arrays
	if (s->tv0) { cudaFree(s->tv0); s->tv0 = NULL; }
	if (s->tv1) { cudaFree(s->tv1); s->tv1 = NULL; }
	if (s->tv2) { cudaFree(s->tv2); s->tv2 = NULL; }
	if (s->nodes) { cudaFree(s->nodes); s->nodes = NULL; }
	s->num_nodes = 0; s->num_leaves = 0; s->depth = 0; s->sah_cost = 0.0;
	for (int k = 0; k < 3; k++) { s->bounds_min[k] = 0.0f; s->bounds_max[k] = 0.0f; }
	s->abs_max = 0.0f;
	if (n == 0) {
		CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
		s->build_device_ms = 0.0;
		cudaEventDestroy(e0); cudaEventDestroy(e1);
		return RTKD_OK;
	}

	const float4 *tri = (const float4*)s->tri_orig;
	uint32_t *d_bounds = NULL;
	unsigned long long *keys[2] = { NULL, NULL };
	uint32_t *vals[2] = { NULL, NULL };
	uint32_t *counts = NULL, *totals = NULL;
	const uint32_t nblocks = (n + RTK_SORT_TILE - 1) / RTK_SORT_TILE;
	CK(tmp_alloc(&d_bounds, 8, st));
	CK(tmp_alloc(&keys[0], n, st)); CK(tmp_alloc(&keys[1], n, st));
	CK(tmp_alloc(&vals[0], n, st)); CK(tmp_alloc(&vals[1], n, st));
	CK(tmp_alloc(&counts, 256 * (size_t)nblocks, st)); CK(tmp_alloc(&totals, 256, st));

	// scene bounds (ordered-uint encoded: min starts at 0xffffffff, max at 0)
	CK(cudaMemsetAsync(d_bounds, 0xff, 3 * sizeof(uint32_t), st));
	CK(cudaMemsetAsync(d_bounds + 3, 0x00, 3 * sizeof(uint32_t), st));
	RTK_LAUNCH(k_scene_bounds, (n + 255) / 256, 256, st, tri, n, d_bounds); CK_LAUNCH();
	RTK_LAUNCH(k_morton, (n + 255) / 256, 256, st, tri, n, (const uint32_t*)d_bounds, keys[0], vals[0]); CK_LAUNCH();

	// LSD radix sort over the 63 code bits
	int src = 0;
	for (int shift = 0; shift < 64; shift += 8) {
		RTK_LAUNCH(k_radix_hist, nblocks, RTK_SORT_THREADS, st, (const unsigned long long*)keys[src], n, shift, counts, nblocks); CK_LAUNCH();
		RTK_LAUNCH(k_radix_scan, 256, 256, st, counts, nblocks, totals); CK_LAUNCH();
		RTK_LAUNCH(k_radix_scatter, nblocks, RTK_SORT_THREADS, st, (const unsigned long long*)keys[src], (const uint32_t*)vals[src],
		           keys[src ^ 1], vals[src ^ 1], n, shift, (const uint32_t*)counts, (const uint32_t*)totals, nblocks); CK_LAUNCH();
		src ^= 1;
	}
	const unsigned long long *skeys = keys[src];
	const uint32_t *svals = vals[src];

	CK(cudaMalloc((float4**)&s->tv0, sizeof(float4) * (size_t)n));
	CK(cudaMalloc((float4**)&s->tv1, sizeof(float4) * (size_t)n));
	CK(cudaMalloc((float4**)&s->tv2, sizeof(float4) * (size_t)n));

	uint32_t h_bounds[6];
	CK(cudaMemcpyAsync(h_bounds, d_bounds, sizeof(h_bounds), cudaMemcpyDeviceToHost, st));

	// binary tree: binned SAH over the Morton-ordered triangles, or the radix tree itself
	const bool use_sah = mode == 1 && n > RTK_LEAF_MAX;
	rtkd_bvh2 t;
	memset(&t, 0, sizeof(t));
	rtkd_sah sah;
	memset(&sah, 0, sizeof(sah));
	uint32_t *order = NULL;          // leaf order as triangle numbers (SAH mode)
	if (use_sah) {
		int r = build_sah(s, st, tri, svals, d_bounds, n, sah, &order);
		if (r) return r;
		t.left = sah.left; t.right = sah.right; t.first = sah.first; t.last = sah.last; t.blo = sah.blo; t.bhi = sah.bhi;
		svals = order;
	}
	// traversal triangles in leaf order
	RTK_LAUNCH(k_emit_tris, (n + 255) / 256, 256, st, tri, svals, n, (float4*)s->tv0, (float4*)s->tv1, (float4*)s->tv2); CK_LAUNCH();

	float4 *wide = NULL;
	uint32_t num_nodes = 0, num_leaves = 0, depth = 0;
	double h_cost = 0.0;
	if (n == 1) {
		CK(tmp_alloc(&wide, 16, st));
		RTK_LAUNCH(k_single_root, 1, 32, st, tri, svals, wide); CK_LAUNCH();
		num_nodes = 1; num_leaves = 1; depth = 1;
	} else {
		if (!use_sah) {
			CK(tmp_alloc(&t.left, n - 1, st)); CK(tmp_alloc(&t.right, n - 1, st));
			CK(tmp_alloc(&t.parent, 2 * (size_t)n - 1, st));
			CK(tmp_alloc(&t.first, n - 1, st)); CK(tmp_alloc(&t.last, n - 1, st));
			CK(tmp_alloc(&t.blo, 2 * (size_t)n - 1, st)); CK(tmp_alloc(&t.bhi, 2 * (size_t)n - 1, st));
			CK(tmp_alloc(&t.flags, n - 1, st));
			CK(cudaMemsetAsync(t.flags, 0, sizeof(int) * (size_t)(n - 1), st));
			RTK_LAUNCH(k_hierarchy, (n - 1 + 255) / 256, 256, st, skeys, (int)n, t); CK_LAUNCH();
			RTK_LAUNCH(k_refit, (n + 255) / 256, 256, st, tri, svals, (int)n, t); CK_LAUNCH();
		}

		const uint32_t cap = (uint32_t)(((unsigned long long)n * 4) / 7 + 16);
		uint2 *work[2] = { NULL, NULL };
		uint32_t *ctr = NULL;          // [0] n_out, [1] node_alloc, [2] leaf_count, [3] err
		double *d_cost = NULL;
		CK(tmp_alloc(&wide, 16 * (size_t)cap, st));
		CK(tmp_alloc(&work[0], cap, st)); CK(tmp_alloc(&work[1], cap, st));
		CK(tmp_alloc(&ctr, 4, st)); CK(tmp_alloc(&d_cost, 1, st));
		uint32_t h_ctr[4] = { 0, 1, 0, 0 };
		uint2 w0 = make_uint2(0u, 0u);
		CK(cudaMemcpyAsync(ctr, h_ctr, sizeof(h_ctr), cudaMemcpyHostToDevice, st));
		CK(cudaMemcpyAsync(work[0], &w0, sizeof(w0), cudaMemcpyHostToDevice, st));
		CK(cudaMemsetAsync(d_cost, 0, sizeof(double), st));
		uint32_t n_in = 1;
		int wi = 0;
		while (n_in) {
			rtkd_collapse_args a;
			a.work_in = work[wi]; a.n_in = n_in; a.work_out = work[wi ^ 1]; a.n_out = ctr;
			a.node_alloc = ctr + 1; a.node_cap = cap; a.leaf_count = ctr + 2; a.sah_cost = d_cost;
			a.nodes = wide; a.n = (int)n; a.err = ctr + 3;
			RTK_LAUNCH(k_collapse, (n_in + 127) / 128, 128, st, a, t); CK_LAUNCH();
			CK(cudaMemcpyAsync(h_ctr, ctr, sizeof(h_ctr), cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
			n_in = h_ctr[0];
			CK(cudaMemsetAsync(ctr, 0, sizeof(uint32_t), st));
			wi ^= 1;
			depth++;
		}
		num_nodes = h_ctr[1]; num_leaves = h_ctr[2];
		if (h_ctr[3]) { rtkd_set_error("wide-node pool exhausted (cap %u)", cap); return RTKD_ERR_MEMORY; }
		CK(cudaMemcpyAsync(&h_cost, d_cost, sizeof(double), cudaMemcpyDeviceToHost, st));
		if (!use_sah) {
			cudaFreeAsync(t.left, st); cudaFreeAsync(t.right, st); cudaFreeAsync(t.parent, st);
			cudaFreeAsync(t.first, st); cudaFreeAsync(t.last, st); cudaFreeAsync(t.blo, st); cudaFreeAsync(t.bhi, st);
			cudaFreeAsync(t.flags, st);
		}
		cudaFreeAsync(work[0], st); cudaFreeAsync(work[1], st);
		cudaFreeAsync(ctr, st); cudaFreeAsync(d_cost, st);
	}

	if (use_sah) free_sah(sah, order, st);

	// exact-size node array
	CK(cudaMalloc((float4**)&s->nodes, sizeof(float4) * 16 * (size_t)num_nodes));
	CK(cudaMemcpyAsync(s->nodes, wide, sizeof(float4) * 16 * (size_t)num_nodes, cudaMemcpyDeviceToDevice, st));
	cudaFreeAsync(wide, st);
	cudaFreeAsync(d_bounds, st); cudaFreeAsync(keys[0], st); cudaFreeAsync(keys[1], st);
	cudaFreeAsync(vals[0], st); cudaFreeAsync(vals[1], st); cudaFreeAsync(counts, st); cudaFreeAsync(totals, st);

	CK(cudaEventRecord(e1, st));
	CK(cudaEventSynchronize(e1));
	float ms = 0.0f;
	CK(cudaEventElapsedTime(&ms, e0, e1));
	cudaEventDestroy(e0); cudaEventDestroy(e1);

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
	// stack scratch depends on the depth: force re-creation
	if (s->overflow) { cudaFree(s->overflow); s->overflow = NULL; s->overflow_entries = 0; }
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
	a.abs_max = s->abs_max;
}

static int ensure_scratch(rtkd_scene *s)
{
	// scratch layout: +0 ray cursor | +64 stats[6] | +128 hit counter | +192 sticky error flags
	if (!s->scratch) {
		CK(cudaMalloc((unsigned char**)&s->scratch, 256));
		CK(cudaMemset(s->scratch, 0, 256));
	}
	size_t groups = (size_t)g_sm_count * g_trace_ctas * RTK_GROUPS_PER_CTA;
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

extern "C" int rtkd_trace(rtkd_scene *s, const void *d_rays, void *d_hit16, size_t n, int cull_mode,
                          rtkd_trace_stats *stats, void *stream)
{
	if (!n) { if (stats) memset(stats, 0, sizeof(*stats)); return RTKD_OK; }
	if (n > 0xfffffff0ull) { rtkd_set_error("batch too large (%zu rays); split it", n); return RTKD_ERR_ARGUMENT; }
	if (((uintptr_t)d_rays & 15) || ((uintptr_t)d_hit16 & 15)) { rtkd_set_error("device ray / hit buffers must be 16-byte aligned"); return RTKD_ERR_ARGUMENT; }
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
	size_t ctas = (size_t)g_sm_count * g_trace_ctas;
	size_t want = (batches + RTK_TRACE_WARPS - 1) / RTK_TRACE_WARPS;
	unsigned grid = (unsigned)(want < ctas ? want : ctas);
	if (stats) {
		if (cull_mode) { RTK_LAUNCH((k_trace<1, true>), grid, RTK_TRACE_THREADS, st, p); }
		else { RTK_LAUNCH((k_trace<0, true>), grid, RTK_TRACE_THREADS, st, p); }
	} else {
		if (cull_mode) { RTK_LAUNCH((k_trace<1, false>), grid, RTK_TRACE_THREADS, st, p); }
		else { RTK_LAUNCH((k_trace<0, false>), grid, RTK_TRACE_THREADS, st, p); }
	}
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
	RTK_LAUNCH(k_resolve, (unsigned)((n + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS), RTK_RESOLVE_THREADS, stream,
	           a, (const float4*)d_hit16, (uint32_t*)d_hits, (unsigned char*)d_mask, (uint32_t)n, cnt);
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

// Host-buffer batch.  Rays go up and hits come back in chunks on two streams so that the
// H2D copy of chunk k+1, the kernels of chunk k and the D2H copy of chunk k-1 overlap.
extern "C" long long rtkd_trace_host(rtkd_scene *s, const void *rays, void *hits, unsigned char *mask, size_t n)
{
	if (!n) return 0;
	const size_t CH = (size_t)1 << 21;            // 2 Mi rays = 64 MiB up, 136 MiB down
	const size_t chunk = n < CH ? n : CH;
	cudaStream_t st[2];
	float4 *d_rays[2] = { NULL, NULL }, *d_h16[2] = { NULL, NULL };
	uint32_t *d_hits[2] = { NULL, NULL };
	unsigned char *d_mask[2] = { NULL, NULL };
	long long total = 0;
	int rc = RTKD_OK;
	for (int k = 0; k < 2; k++) {
		if (cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking) != cudaSuccess) { rtkd_set_error("stream creation failed"); return -1; }
	}
	for (int k = 0; k < 2 && rc == RTKD_OK; k++) {
		cudaError_t e = cudaMalloc(&d_rays[k], 32 * chunk);
		if (e == cudaSuccess) e = cudaMalloc(&d_h16[k], 16 * chunk);
		if (e == cudaSuccess) e = cudaMalloc(&d_hits[k], 68 * chunk);
		if (e == cudaSuccess) e = cudaMalloc(&d_mask[k], chunk);
		if (e != cudaSuccess) { rtkd_set_error("batch staging allocation failed: %s", cudaGetErrorString(e)); rc = RTKD_ERR_MEMORY; }
	}
	// The traversal scratch (ray cursor, overflow slab) belongs to the scene, so the kernels of
	// consecutive chunks are serialised on purpose: chunk k+1's kernels wait for chunk k's, while
	// its H2D copy and chunk k's D2H copy run beside them (pinned caller buffers overlap fully).
	cudaEvent_t done[2];
	cudaEventCreate(&done[0]); cudaEventCreate(&done[1]);
	unsigned long long *d_count = NULL;
	if (rc == RTKD_OK) rc = ensure_scratch(s);
	if (rc == RTKD_OK) {
		d_count = (unsigned long long*)((unsigned char*)s->scratch + 128);
		if (cudaMemset(d_count, 0, sizeof(unsigned long long)) != cudaSuccess) rc = RTKD_ERR_CUDA;
	}
	size_t nchunks = (n + chunk - 1) / chunk;
	for (size_t ci = 0; ci < nchunks && rc == RTKD_OK; ci++) {
		int k = (int)(ci & 1);
		size_t off = ci * chunk, cnt = n - off < chunk ? n - off : chunk;
		cudaStream_t q = st[k];
		if (cudaMemcpyAsync(d_rays[k], (const char*)rays + 32 * off, 32 * cnt, cudaMemcpyHostToDevice, q) != cudaSuccess) { rc = RTKD_ERR_CUDA; break; }
		// rows of rays that miss come back zero-filled (see rtk_cuda.h)
		if (cudaMemsetAsync(d_hits[k], 0, 68 * cnt, q) != cudaSuccess) { rc = RTKD_ERR_CUDA; break; }
		if (ci > 0) cudaStreamWaitEvent(q, done[k ^ 1], 0);
		rc = rtkd_trace(s, d_rays[k], d_h16[k], cnt, 1, NULL, q);
		if (rc) break;
		{
			rtkd_arrays a;
			fill_arrays(s, a);
			RTK_LAUNCH(k_resolve, (unsigned)((cnt + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS), RTK_RESOLVE_THREADS, q,
			           a, (const float4*)d_h16[k], d_hits[k], d_mask[k], (uint32_t)cnt, d_count);
		}
		cudaEventRecord(done[k], q);
		if (cudaMemcpyAsync((char*)hits + 68 * off, d_hits[k], 68 * cnt, cudaMemcpyDeviceToHost, q) != cudaSuccess) { rc = RTKD_ERR_CUDA; break; }
		if (mask && cudaMemcpyAsync(mask + off, d_mask[k], cnt, cudaMemcpyDeviceToHost, q) != cudaSuccess) { rc = RTKD_ERR_CUDA; break; }
	}
	for (int k = 0; k < 2; k++) {
		if (cudaStreamSynchronize(st[k]) != cudaSuccess && rc == RTKD_OK) {
			rtkd_set_error("batch failed: %s", cudaGetErrorString(cudaGetLastError())); rc = RTKD_ERR_CUDA;
		}
	}
	if (rc == RTKD_OK) {
		unsigned long long hc = 0;
		uint32_t herr = 0;
		if (cudaMemcpy(&hc, d_count, sizeof(hc), cudaMemcpyDeviceToHost) != cudaSuccess) rc = RTKD_ERR_CUDA;
		if (cudaMemcpy(&herr, (unsigned char*)s->scratch + 192, sizeof(herr), cudaMemcpyDeviceToHost) != cudaSuccess) rc = RTKD_ERR_CUDA;
		total = (long long)hc;
		if (herr & 2u) { rtkd_set_error("traversal stack exhausted"); rc = RTKD_ERR_OVERFLOW; }
	}
	cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
	for (int k = 0; k < 2; k++) {
		cudaFree(d_rays[k]); cudaFree(d_h16[k]); cudaFree(d_hits[k]); cudaFree(d_mask[k]);
		cudaStreamDestroy(st[k]);
	}
	if (rc == RTKD_ERR_CUDA && !g_err[0]) rtkd_set_error("CUDA failure in rtk_trace_rays");
	return rc == RTKD_OK ? total : -1;
}

// ---------------------------------------------------------------------------------------------
// serialisation
// ---------------------------------------------------------------------------------------------

struct rtkd_blob_sub {          // 128 bytes, first thing in the payload
	uint64_t magic2;            // "B200RTK1"
	uint64_t id;
	uint32_t num_tris, num_meshes, num_nodes, num_leaves, depth, build_mode;
	float bounds_min[3], bounds_max[3], abs_max;
	float pad0;
	double sah_cost;
	uint64_t off_nodes, off_tv0, off_tv1, off_tv2, off_orig, off_mesh;   // from payload start
};

static size_t a128(size_t v) { return (v + 127) & ~(size_t)127; }

static void blob_layout(const rtkd_scene *s, rtkd_blob_sub *b)
{
	size_t o = a128(sizeof(rtkd_blob_sub));
	b->off_nodes = o; o = a128(o + 256 * (size_t)s->num_nodes);
	b->off_tv0 = o;   o = a128(o + 16 * (size_t)s->num_tris);
	b->off_tv1 = o;   o = a128(o + 16 * (size_t)s->num_tris);
	b->off_tv2 = o;   o = a128(o + 16 * (size_t)s->num_tris);
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
	memcpy(&b.magic2, "B200RTK1", 8);
	b.id = s->id;
	b.num_tris = s->num_tris; b.num_meshes = s->num_meshes; b.num_nodes = s->num_nodes;
	b.num_leaves = s->num_leaves; b.depth = s->depth; b.build_mode = s->build_mode;
	memcpy(b.bounds_min, s->bounds_min, 12); memcpy(b.bounds_max, s->bounds_max, 12);
	b.abs_max = s->abs_max; b.sah_cost = s->sah_cost;
	blob_layout(s, &b);
	char *p = (char*)payload;
	memset(p, 0, a128(sizeof(b)));
	memcpy(p, &b, sizeof(b));
	CK(cudaDeviceSynchronize());
	if (s->num_nodes) CK(cudaMemcpy(p + b.off_nodes, s->nodes, 256 * (size_t)s->num_nodes, cudaMemcpyDeviceToHost));
	if (s->num_tris) {
		CK(cudaMemcpy(p + b.off_tv0, s->tv0, 16 * (size_t)s->num_tris, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(p + b.off_tv1, s->tv1, 16 * (size_t)s->num_tris, cudaMemcpyDeviceToHost));
		CK(cudaMemcpy(p + b.off_tv2, s->tv2, 16 * (size_t)s->num_tris, cudaMemcpyDeviceToHost));
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
	if (memcmp(&b.magic2, "B200RTK1", 8) != 0) { rtkd_set_error("blob was not written by rtk_b200 (device-layout magic missing)"); return NULL; }
	if (b.off_mesh + 4 * ((size_t)b.num_meshes + 1) > payload_size) { rtkd_set_error("scene blob truncated"); return NULL; }
	const char *p = (const char*)payload;
	rtkd_scene *s = rtkd_scene_new(b.num_tris, b.num_meshes, (const uint32_t*)(p + b.off_mesh));
	if (!s) return NULL;
	s->id = b.id;
	s->num_nodes = b.num_nodes; s->num_leaves = b.num_leaves; s->depth = b.depth; s->build_mode = b.build_mode;
	memcpy(s->bounds_min, b.bounds_min, 12); memcpy(s->bounds_max, b.bounds_max, 12);
	s->abs_max = b.abs_max; s->sah_cost = b.sah_cost;
	cudaError_t e = cudaSuccess;
	if (b.num_tris) {
		e = cudaMemcpy(s->tri_orig, p + b.off_orig, 48 * (size_t)b.num_tris, cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMalloc((float4**)&s->tv0, 16 * (size_t)b.num_tris);
		if (e == cudaSuccess) e = cudaMalloc((float4**)&s->tv1, 16 * (size_t)b.num_tris);
		if (e == cudaSuccess) e = cudaMalloc((float4**)&s->tv2, 16 * (size_t)b.num_tris);
		if (e == cudaSuccess) e = cudaMemcpy(s->tv0, p + b.off_tv0, 16 * (size_t)b.num_tris, cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(s->tv1, p + b.off_tv1, 16 * (size_t)b.num_tris, cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(s->tv2, p + b.off_tv2, 16 * (size_t)b.num_tris, cudaMemcpyHostToDevice);
	}
	if (e == cudaSuccess && b.num_nodes) {
		e = cudaMalloc((float4**)&s->nodes, 256 * (size_t)b.num_nodes);
		if (e == cudaSuccess) e = cudaMemcpy(s->nodes, p + b.off_nodes, 256 * (size_t)b.num_nodes, cudaMemcpyHostToDevice);
	}
	if (e != cudaSuccess) {
		rtkd_set_error("scene upload failed: %s", cudaGetErrorString(e));
		rtkd_scene_free(s);
		return NULL;
	}
	return s;
}
