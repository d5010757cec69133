/*
 * rtk_host.c -- the C host layer of librtk_b200.so: the rtk.h entry points (reference
 * rtk.c:543-582, 1625-1797) and the batched extension of rtk_cuda.h, implemented on top of the
 * thin C-ABI device layer in rtk_device.h.  No compute happens here: this file validates
 * arguments, walks mesh descriptions (strides, element types, callbacks), moves bytes to the
 * device and keeps the table that maps scene blobs to their device-resident copies.
 *
 * There is no CPU fallback.  When the CUDA layer cannot initialise, every entry point returns
 * NULL / false / a negative status and prints the reason on stderr once.
 */
#include "../../include/rtk.h"
#include "../../include/rtk_cuda.h"
#include "rtk_device.h"

#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define RTK_B200_VERSION 0x00B20001u
#define HEADER_BLOCK 128          /* rtk_scene (56 B) + library fields, payload follows */

/* library fields inside the 128-byte header block, after the 56-byte rtk_scene */
typedef struct rtk_block_ext {
	uint64_t scene_id;            /* @64 */
	uint64_t flags;               /* @72  bit 0: payload (device layout) follows the block */
} rtk_block_ext;
#define EXT_OF(scene) ((rtk_block_ext*)((char*)(scene) + 64))
#define FLAG_PAYLOAD 1u

static const char k_magic[8] = { 0, 'R', 'T', 'K', '\r', '\n', 0x1a, '\n' };   /* rtk.c:1737 */

/* ---------------------------------------------------------------------------------------- */
/* errors                                                                                     */
/* ---------------------------------------------------------------------------------------- */

static int g_warned;
static void warn_once(void)
{
	if (!g_warned) {
		g_warned = 1;
		fprintf(stderr, "rtk_b200: %s\n", rtkd_last_error());
	}
}

const char *rtk_cuda_last_error(void) { return rtkd_last_error(); }

static int g_build_mode = RTK_CUDA_BUILD_SAH;   /* binned SAH is the default: 2x fewer traversal steps than the plain radix tree */
int rtk_cuda_set_build_mode(int mode)
{
	if (mode != RTK_CUDA_BUILD_LBVH && mode != RTK_CUDA_BUILD_SAH) { rtkd_set_error("unknown build mode %d", mode); return RTK_CUDA_ERR_ARGUMENT; }
	g_build_mode = mode;
	return RTK_CUDA_OK;
}

static int g_cull_mode = 1;
int rtk_cuda_set_cull_mode(int mode)
{
	if (mode != 0 && mode != 1) { rtkd_set_error("unknown cull mode %d", mode); return RTK_CUDA_ERR_ARGUMENT; }
	g_cull_mode = mode;
	return RTK_CUDA_OK;
}

int rtk_cuda_init(int device)
{
	int r = rtkd_init(device);
	if (r) warn_once();
	return r;
}
int rtk_cuda_init_devices(const int *devices, int num_devices)
{
	int r = rtkd_init_devices(devices, num_devices);
	if (r) warn_once();
	return r;
}
int rtk_cuda_device_count(void) { return rtkd_device_count(); }
void *rtk_cuda_host_alloc(size_t bytes) { return rtkd_host_alloc(bytes); }
void rtk_cuda_host_free(void *p) { rtkd_host_free(p); }
void *rtk_cuda_host_alloc_batch(size_t element_bytes, size_t count) { return rtkd_host_alloc_batch(element_bytes, count); }
int rtk_cuda_host_register(void *p, size_t bytes) { return rtkd_host_register(p, bytes); }
int rtk_cuda_host_unregister(void *p) { return rtkd_host_unregister(p); }
void rtk_cuda_shutdown(void) { rtkd_shutdown(); }
int rtk_cuda_device_info(int *sm_count, size_t *l2_bytes, int *ctas, int *threads)
{
	return rtkd_device_info(sm_count, l2_bytes, ctas, threads);
}

int rtk_cuda_reserve_sms(int sms) { return rtkd_reserve_sms(sms); }

int rtk_cuda_measure_read_bandwidth(size_t bytes, int passes, double *gb_per_s)
{
	return rtkd_read_bandwidth(bytes, passes, gb_per_s);
}
int rtk_cuda_measure_gather_bandwidth(size_t bytes, size_t record_bytes, int passes, double *gb_per_s)
{
	return rtkd_gather_bandwidth(bytes, record_bytes, passes, gb_per_s);
}
int rtk_cuda_measure_host_link(int num_devices, size_t bytes_per_device, int directions, int passes, double *gb_per_s)
{
	return rtkd_link_bandwidth(num_devices, bytes_per_device, directions, passes, gb_per_s);
}

/* ---------------------------------------------------------------------------------------- */
/* scene table: blob address -> device scene                                                  */
/* ---------------------------------------------------------------------------------------- */

typedef struct scene_entry {
	const rtk_scene *addr;
	rtkd_scene *dev;
	int owned;                    /* the blob memory was allocated by this library */
} scene_entry;

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static scene_entry *g_table;
static size_t g_table_n, g_table_cap;

static scene_entry *table_find(const rtk_scene *addr)
{
	for (size_t i = 0; i < g_table_n; i++) if (g_table[i].addr == addr) return &g_table[i];
	return NULL;
}

static int table_add(const rtk_scene *addr, rtkd_scene *dev, int owned)
{
	if (g_table_n == g_table_cap) {
		size_t cap = g_table_cap ? g_table_cap * 2 : 16;
		scene_entry *t = (scene_entry*)realloc(g_table, cap * sizeof(scene_entry));
		if (!t) return -1;
		g_table = t; g_table_cap = cap;
	}
	g_table[g_table_n].addr = addr; g_table[g_table_n].dev = dev; g_table[g_table_n].owned = owned;
	g_table_n++;
	return 0;
}

static void table_remove(scene_entry *e)
{
	*e = g_table[g_table_n - 1];
	g_table_n--;
}

static int header_valid(const rtk_scene *sc)
{
	if (memcmp(sc->magic, k_magic, 8) != 0) { rtkd_set_error("not an rtk scene (bad magic)"); return 0; }
	if (sc->endian != 0xaabb) { rtkd_set_error("scene blob has foreign endianness"); return 0; }
	if (sc->sizeof_real != sizeof(rtk_real)) { rtkd_set_error("scene blob uses %u-byte reals", sc->sizeof_real); return 0; }
	if (sc->version != RTK_B200_VERSION) {
		rtkd_set_error("scene blob version %#x was not written by rtk_b200 (reference-format blobs cannot be traced on the GPU)", sc->version);
		return 0;
	}
	return 1;
}

/* device scene of a blob; uploads a relocated / reloaded blob on first use */
static rtkd_scene *scene_device(const rtk_scene *scene)
{
	if (!scene) { rtkd_set_error("scene is NULL"); return NULL; }
	if (rtkd_bind_thread() != RTK_CUDA_OK) { warn_once(); return NULL; }     /* callers' worker threads start on device 0 */
	pthread_mutex_lock(&g_lock);
	scene_entry *e = table_find(scene);
	if (e && e->dev->id == EXT_OF(scene)->scene_id) {
		rtkd_scene *d = e->dev;
		pthread_mutex_unlock(&g_lock);
		return d;
	}
	rtkd_scene *dev = NULL;
	if (header_valid(scene)) {
		if (!(EXT_OF(scene)->flags & FLAG_PAYLOAD)) {
			rtkd_set_error("scene handle is not resident on this device (handles from rtk_build_scene cannot be copied; use rtk_finish_build_to for a relocatable blob)");
		} else {
			if (e) { rtkd_scene_free(e->dev); table_remove(e); }       /* a different blob now lives at this address */
			if (scene->size_in_bytes < HEADER_BLOCK) rtkd_set_error("scene blob truncated");
			else dev = rtkd_blob_read((const char*)scene + HEADER_BLOCK, (size_t)scene->size_in_bytes - HEADER_BLOCK);
			if (dev && table_add(scene, dev, 0) != 0) { rtkd_scene_free(dev); dev = NULL; rtkd_set_error("out of host memory"); }
		}
	}
	pthread_mutex_unlock(&g_lock);
	if (!dev) warn_once();
	return dev;
}

/* the copy of the scene on the device that owns the caller's device buffer (several devices in use) */
static rtkd_scene *scene_on_device_of(const rtk_scene *scene, const void *d_ptr)
{
	rtkd_scene *dev = scene_device(scene);
	if (!dev) return NULL;
	rtkd_scene *r = rtkd_scene_for_pointer(dev, d_ptr);
	if (!r) warn_once();
	return r;
}

int rtk_cuda_scene_status(const rtk_scene *scene)
{
	rtkd_scene *dev = scene_device(scene);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	if (rtkd_scene_status(dev) & 2u) { rtkd_set_error("traversal stack exhausted"); return RTK_CUDA_ERR_OVERFLOW; }
	return RTK_CUDA_OK;
}

int rtk_cuda_debug_limit_stack(int entries) { return rtkd_debug_limit_stack(entries); }

int rtk_cuda_attach_scene(const rtk_scene *scene) { return scene_device(scene) ? RTK_CUDA_OK : RTK_CUDA_ERR_SCENE; }

int rtk_cuda_detach_scene(const rtk_scene *scene)
{
	pthread_mutex_lock(&g_lock);
	scene_entry *e = table_find(scene);
	int r = RTK_CUDA_ERR_SCENE;
	if (e && !e->owned) { rtkd_scene_free(e->dev); table_remove(e); r = RTK_CUDA_OK; }
	pthread_mutex_unlock(&g_lock);
	return r;
}

static void write_header(rtk_scene *sc, const rtkd_scene *dev, size_t total, int payload)
{
	/* rtk.c:1737-1753 */
	memset(sc, 0, HEADER_BLOCK);
	memcpy(sc->magic, k_magic, 8);
	sc->endian = 0xaabb;
	sc->sizeof_real = (uint8_t)sizeof(rtk_real);
	sc->version = RTK_B200_VERSION;
	sc->size_in_bytes = total;
	sc->node_offset = HEADER_BLOCK;       /* the payload starts with its own section table */
	sc->leaf_offset = HEADER_BLOCK;
	sc->vertex_offset = HEADER_BLOCK;
	EXT_OF(sc)->scene_id = dev->id;
	EXT_OF(sc)->flags = payload ? FLAG_PAYLOAD : 0;
}

/* ---------------------------------------------------------------------------------------- */
/* build                                                                                      */
/* ---------------------------------------------------------------------------------------- */

struct rtk_build {
	rtk_scene_desc desc;          /* copied, rtk.c:1661 */
	rtk_mesh *meshes;             /* copied array */
	size_t num_triangles;
	uint32_t *mesh_first;
	rtkd_scene *dev;              /* set once the build task ran */
	int failed;
};

struct rtk_task_ctx { rtk_build *build; rtk_task *queue; size_t queue_capacity, queue_num; };

static void build_log(rtk_build *b, const char *fmt, ...)
{
	/* rtk.c:686-696 */
	if (b->desc.log_fn) {
		char buf[256];
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(buf, sizeof(buf), fmt, ap);
		va_end(ap);
		b->desc.log_fn(b->desc.log_user, b, buf);
	}
}

static double now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* Host gather of one mesh that uses callbacks or has buffers the device cannot read directly
 * (misaligned): produces per-corner positions and the index triples, chunk by chunk like the
 * reference (rtk.c:1141-1175, chunks of <= 128 triangles). */
static int gather_mesh_host(const rtk_mesh *mesh, float *pos9, uint32_t *idx3)
{
	size_t nt = mesh->num_triangles;
	rtk_type ptype = mesh->position.type;
	if (ptype == RTK_TYPE_DEFAULT || ptype == RTK_TYPE_REAL) ptype = RTK_TYPE_F32;
	rtk_type itype = mesh->index.type == RTK_TYPE_DEFAULT ? RTK_TYPE_U32 : mesh->index.type;
	for (size_t base = 0; base < nt; base += 128) {
		size_t cnt = nt - base < 128 ? nt - base : 128;
		uint32_t *ix = idx3 + 3 * base;
		if (mesh->index_cb) {
			mesh->index_cb(mesh->index_cb_user, mesh, ix, base, cnt);                 /* rtk.c:1030-1033 */
		} else if (mesh->index.data) {
			if (itype == RTK_TYPE_U16) {
				size_t stride = mesh->index.stride ? mesh->index.stride : 6;
				for (size_t i = 0; i < cnt; i++) {
					uint16_t t[3];
					memcpy(t, (const char*)mesh->index.data + (base + i) * stride, 6);
					ix[3 * i] = t[0]; ix[3 * i + 1] = t[1]; ix[3 * i + 2] = t[2];
				}
			} else if (itype == RTK_TYPE_U32) {
				size_t stride = mesh->index.stride ? mesh->index.stride : 12;
				for (size_t i = 0; i < cnt; i++) memcpy(ix + 3 * i, (const char*)mesh->index.data + (base + i) * stride, 12);
			} else { rtkd_set_error("bad index type %d", (int)itype); return RTK_CUDA_ERR_ARGUMENT; }
		} else {
			for (size_t i = 0; i < 3 * cnt; i++) ix[i] = (uint32_t)(3 * base + i);  /* rtk.c:1062-1068 */
		}
		float *dst = pos9 + 9 * base;
		if (mesh->position_cb) {
			mesh->position_cb(mesh->position_cb_user, mesh, (rtk_vec3*)dst, ix, cnt);   /* rtk.c:1074-1077 */
		} else if (mesh->position.data) {
			if (ptype == RTK_TYPE_F32) {
				size_t stride = mesh->position.stride ? mesh->position.stride : 12;
				for (size_t i = 0; i < 3 * cnt; i++) memcpy(dst + 3 * i, (const char*)mesh->position.data + (size_t)ix[i] * stride, 12);
			} else if (ptype == RTK_TYPE_F64) {
				size_t stride = mesh->position.stride ? mesh->position.stride : 24;
				for (size_t i = 0; i < 3 * cnt; i++) {
					double d[3];
					memcpy(d, (const char*)mesh->position.data + (size_t)ix[i] * stride, 24);
					dst[3 * i] = (float)d[0]; dst[3 * i + 1] = (float)d[1]; dst[3 * i + 2] = (float)d[2];
				}
			} else { rtkd_set_error("bad position type %d", (int)ptype); return RTK_CUDA_ERR_ARGUMENT; }
		} else { rtkd_set_error("mesh has neither a position buffer nor a position callback"); return RTK_CUDA_ERR_ARGUMENT; }
	}
	return RTK_CUDA_OK;
}

static int ingest_mesh(rtk_build *b, size_t mi)
{
	const rtk_mesh *mesh = &b->meshes[mi];
	size_t nt = mesh->num_triangles;
	uint32_t first = b->mesh_first[mi];
	if (!nt) return RTK_CUDA_OK;
	rtk_type ptype = mesh->position.type;
	if (ptype == RTK_TYPE_DEFAULT || ptype == RTK_TYPE_REAL) ptype = RTK_TYPE_F32;     /* rtk.c:1081-1083 */
	rtk_type itype = mesh->index.type == RTK_TYPE_DEFAULT ? RTK_TYPE_U32 : mesh->index.type;
	size_t pelt = ptype == RTK_TYPE_F64 ? 8 : 4;
	size_t pstride = mesh->position.stride ? mesh->position.stride : 3 * pelt;
	size_t ielt = itype == RTK_TYPE_U16 ? 2 : 4;
	size_t istride = mesh->index.stride ? mesh->index.stride : 3 * ielt;

	int direct = !mesh->position_cb && !mesh->index_cb && mesh->position.data &&
	             (ptype == RTK_TYPE_F32 || ptype == RTK_TYPE_F64) &&
	             ((uintptr_t)mesh->position.data % pelt) == 0 && (pstride % pelt) == 0;
	if (direct && mesh->index.data)
		direct = (itype == RTK_TYPE_U16 || itype == RTK_TYPE_U32) &&
		         ((uintptr_t)mesh->index.data % ielt) == 0 && (istride % ielt) == 0;

	if (direct) {
		/* device decode straight from the caller's layout */
		void *d_idx = NULL, *d_pos = NULL;
		size_t nverts = 3 * nt;
		int r;
		if (mesh->index.data) {
			d_idx = rtkd_upload(mesh->index.data, (nt - 1) * istride + 3 * ielt, NULL);
			if (!d_idx) return RTK_CUDA_ERR_CUDA;
			uint32_t mx = 0;
			r = rtkd_max_index(d_idx, istride, (int)ielt, (uint32_t)nt, &mx, NULL);
			if (r) { rtkd_free_async(d_idx, NULL); return r; }
			nverts = (size_t)mx + 1;
		}
		d_pos = rtkd_upload(mesh->position.data, (nverts - 1) * pstride + 3 * pelt, NULL);
		if (!d_pos) { rtkd_free_async(d_idx, NULL); return RTK_CUDA_ERR_CUDA; }
		r = rtkd_decode_mesh(b->dev, first, (uint32_t)nt, d_pos, pstride, ptype == RTK_TYPE_F64,
		                     d_idx, istride, (int)ielt, 0, NULL);
		rtkd_free_async(d_pos, NULL); rtkd_free_async(d_idx, NULL);
		return r;
	}

	/* callbacks / odd layouts: gather on the host, decode the tight arrays on the device */
	float *pos9 = (float*)malloc(sizeof(float) * 9 * nt);
	uint32_t *idx3 = (uint32_t*)malloc(sizeof(uint32_t) * 3 * nt);
	int r = (pos9 && idx3) ? gather_mesh_host(mesh, pos9, idx3) : RTK_CUDA_ERR_MEMORY;
	if (r == RTK_CUDA_OK) {
		void *d_pos = rtkd_upload(pos9, sizeof(float) * 9 * nt, NULL);
		void *d_idx = rtkd_upload(idx3, sizeof(uint32_t) * 3 * nt, NULL);
		if (d_pos && d_idx) r = rtkd_decode_mesh(b->dev, first, (uint32_t)nt, d_pos, 12, 0, d_idx, 12, 4, 1, NULL);
		else r = RTK_CUDA_ERR_CUDA;
		if (r == RTK_CUDA_OK) r = rtkd_sync(NULL);     /* host arrays are freed below */
		rtkd_free_async(d_pos, NULL); rtkd_free_async(d_idx, NULL);
	} else if (r == RTK_CUDA_ERR_MEMORY) rtkd_set_error("out of host memory while gathering mesh %zu", mi);
	free(pos9); free(idx3);
	return r;
}

/* the single build task: ingest every mesh, then the whole GPU build (rtk.c:1362-1507 are
 * five task stages in the reference) */
static void task_build(const rtk_task *task, rtk_task_ctx *ctx)
{
	(void)ctx;
	rtk_build *b = task->build;
	if (b->dev || b->failed) return;
	double t0 = now_ms();
	build_log(b, "Starting build");                                            /* rtk.c:1365 */
	b->dev = rtkd_scene_new((uint32_t)b->num_triangles, (uint32_t)b->desc.num_meshes, b->mesh_first);
	if (!b->dev) { b->failed = 1; warn_once(); return; }
	for (size_t mi = 0; mi < b->desc.num_meshes; mi++) {
		build_log(b, "Gathering triangles [%zu, %zu] of %zu", (size_t)b->mesh_first[mi],
		          (size_t)b->mesh_first[mi + 1], b->num_triangles);                 /* rtk.c:1124 */
		if (ingest_mesh(b, mi) != RTK_CUDA_OK) { b->failed = 1; break; }
	}
	if (!b->failed) {
		build_log(b, "Starting to build nodes");                                /* rtk.c:1396 */
		if (rtkd_build(b->dev, g_build_mode, NULL) != RTK_CUDA_OK) b->failed = 1;
	}
	if (b->failed) {
		warn_once();
		rtkd_scene_free(b->dev);
		b->dev = NULL;
		return;
	}
	b->dev->build_total_ms = now_ms() - t0;
	build_log(b, "Build finished: %u wide nodes, %u leaves, depth %u, %.3f ms on device",
	          b->dev->num_nodes, b->dev->num_leaves, b->dev->depth, b->dev->build_device_ms);
}

static void build_free(rtk_build *b)
{
	if (!b) return;
	free(b->meshes);
	free(b->mesh_first);
	free(b);
}

rtk_build *rtk_start_build(const rtk_scene_desc *desc, rtk_task *first_task)
{
	if (!desc || (desc->num_meshes && !desc->meshes)) { rtkd_set_error("bad scene description"); return NULL; }
	rtk_build *b = (rtk_build*)calloc(1, sizeof(rtk_build));
	if (!b) return NULL;                                                         /* rtk.c:1648 */
	b->desc = *desc;
	b->meshes = (rtk_mesh*)malloc(sizeof(rtk_mesh) * (desc->num_meshes ? desc->num_meshes : 1));
	b->mesh_first = (uint32_t*)malloc(sizeof(uint32_t) * (desc->num_meshes + 1));
	if (!b->meshes || !b->mesh_first) { build_free(b); return NULL; }
	size_t total = 0;
	for (size_t i = 0; i < desc->num_meshes; i++) {
		b->meshes[i] = desc->meshes[i];
		b->mesh_first[i] = (uint32_t)total;
		total += desc->meshes[i].num_triangles;                                 /* rtk.c:1631-1635 */
	}
	b->mesh_first[desc->num_meshes] = (uint32_t)total;
	if (total > 0x0fffffffu) { rtkd_set_error("scene has %zu triangles; the leaf reference holds 28 bits", total); build_free(b); warn_once(); return NULL; }
	b->desc.meshes = b->meshes;
	b->num_triangles = total;
	rtk_task t;
	memset(&t, 0, sizeof(t));
	t.build = b; t.fn = &task_build; t.cost = (double)total * 10.0;
	if (first_task) *first_task = t;                                             /* rtk.c:1679-1681 */
	else rtk_run_task(&t, NULL, 0);                                              /* rtk.c:1682-1688 */
	return b;
}

size_t rtk_run_task(const rtk_task *task, rtk_task *queue, size_t queue_size)
{
	/* rtk.c:1692-1717.  The GPU build is one task; it never queues follow-ups. */
	if (!task || !task->build || !task->fn) return 0;
	rtk_task_ctx ctx;
	ctx.build = task->build; ctx.queue = queue; ctx.queue_capacity = queue_size; ctx.queue_num = 0;
	task->fn(task, &ctx);
	return ctx.queue_num;
}

size_t rtk_get_build_size(const rtk_build *build)
{
	/* rtk.c:1719-1730 */
	if (!build || !build->dev) return 0;
	return HEADER_BLOCK + rtkd_blob_payload_size(build->dev);
}

rtk_scene *rtk_finish_build_to(rtk_build *build, void *buffer, size_t size)
{
	if (!build || !buffer) return NULL;
	if (!build->dev) {
		/* the task was never pumped: run it now rather than fail */
		rtk_task t;
		memset(&t, 0, sizeof(t));
		t.build = build; t.fn = &task_build;
		rtk_run_task(&t, NULL, 0);
		if (!build->dev) return NULL;
	}
	size_t required = rtk_get_build_size(build);
	if (size < required) return NULL;                                            /* rtk.c:1734-1735: build stays alive */
	rtk_scene *sc = (rtk_scene*)buffer;
	write_header(sc, build->dev, required, 1);
	if (rtkd_blob_write(build->dev, (char*)buffer + HEADER_BLOCK) != RTK_CUDA_OK) { warn_once(); return NULL; }
	pthread_mutex_lock(&g_lock);
	scene_entry *old = table_find(sc);
	if (old) { rtkd_scene_free(old->dev); table_remove(old); }
	int r = table_add(sc, build->dev, 0);
	pthread_mutex_unlock(&g_lock);
	if (r) return NULL;
	build->dev = NULL;
	build_free(build);                                                           /* rtk.c:1771 */
	return sc;
}

rtk_scene *rtk_finish_build(rtk_build *build)
{
	/* rtk.c:1776-1786 */
	if (!build) return NULL;
	size_t size = rtk_get_build_size(build);
	void *buffer = NULL;
	if (!size || posix_memalign(&buffer, 128, size) != 0) {
		if (build->dev) rtkd_scene_free(build->dev);
		build_free(build);
		return NULL;
	}
	rtk_scene *sc = rtk_finish_build_to(build, buffer, size);
	if (!sc) {
		free(buffer);
		if (build->dev) rtkd_scene_free(build->dev);
		build_free(build);
		return NULL;
	}
	pthread_mutex_lock(&g_lock);
	scene_entry *e = table_find(sc);
	if (e) e->owned = 1;
	pthread_mutex_unlock(&g_lock);
	return sc;
}

/* wrap a device scene into a library-owned, device-resident handle (header block only) */
static rtk_scene *make_handle(rtkd_scene *dev)
{
	void *block = NULL;
	if (posix_memalign(&block, 128, HEADER_BLOCK) != 0) { rtkd_scene_free(dev); return NULL; }
	rtk_scene *sc = (rtk_scene*)block;
	write_header(sc, dev, HEADER_BLOCK, 0);
	pthread_mutex_lock(&g_lock);
	int r = table_add(sc, dev, 1);
	pthread_mutex_unlock(&g_lock);
	if (r) { free(block); rtkd_scene_free(dev); return NULL; }
	return sc;
}

rtk_scene *rtk_build_scene(const rtk_scene_desc *desc)
{
	/* rtk.c:1788-1792.  Returns a device-resident handle: the scene stays in HBM and is not
	 * serialised; rtk_start_build + rtk_finish_build[_to] yield a relocatable blob instead. */
	rtk_build *b = rtk_start_build(desc, NULL);
	if (!b) return NULL;
	rtkd_scene *dev = b->dev;
	b->dev = NULL;
	build_free(b);
	/* "upload + build + replicate": with several devices in use every one of them gets its copy now */
	if (dev && rtkd_sync_replicas(dev) != RTK_CUDA_OK) { warn_once(); rtkd_scene_free(dev); return NULL; }
	return dev ? make_handle(dev) : NULL;
}

void rtk_free_scene(rtk_scene *scene)
{
	/* rtk.c:1794-1797 */
	if (!scene) return;
	pthread_mutex_lock(&g_lock);
	scene_entry *e = table_find(scene);
	int owned = 0;
	if (e) { owned = e->owned; rtkd_scene_free(e->dev); table_remove(e); }
	pthread_mutex_unlock(&g_lock);
	if (owned) free(scene);
}

/* ---------------------------------------------------------------------------------------- */
/* device-resident meshes                                                                     */
/* ---------------------------------------------------------------------------------------- */

rtk_scene *rtk_cuda_build_scene(const rtk_cuda_mesh *meshes, size_t num_meshes, void *stream)
{
	if (num_meshes && !meshes) { rtkd_set_error("bad mesh array"); return NULL; }
	uint32_t *first = (uint32_t*)malloc(sizeof(uint32_t) * (num_meshes + 1));
	if (!first) return NULL;
	size_t total = 0;
	for (size_t i = 0; i < num_meshes; i++) { first[i] = (uint32_t)total; total += meshes[i].num_triangles; }
	first[num_meshes] = (uint32_t)total;
	if (total > 0x0fffffffu) { free(first); rtkd_set_error("too many triangles"); return NULL; }
	double t0 = now_ms();
	rtkd_scene *dev = rtkd_scene_new((uint32_t)total, (uint32_t)num_meshes, first);
	int r = dev ? RTK_CUDA_OK : RTK_CUDA_ERR_CUDA;
	for (size_t i = 0; i < num_meshes && r == RTK_CUDA_OK; i++)
		r = rtkd_decode_mesh(dev, first[i], (uint32_t)meshes[i].num_triangles, meshes[i].d_positions, 12, 0,
		                     meshes[i].d_indices, 12, 4, 0, stream);
	if (r == RTK_CUDA_OK) r = rtkd_build(dev, g_build_mode, stream);
	free(first);
	if (r != RTK_CUDA_OK) { warn_once(); if (dev) rtkd_scene_free(dev); return NULL; }
	dev->build_total_ms = now_ms() - t0;
	return make_handle(dev);
}

int rtk_cuda_rebuild_scene(const rtk_scene *scene, void *stream)
{
	rtkd_scene *dev = scene_device(scene);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	return rtkd_build(dev, g_build_mode, stream);
}

int rtk_cuda_update_scene(const rtk_scene *scene, const rtk_cuda_mesh *meshes, size_t num_meshes, int mode, void *stream)
{
	rtkd_scene *dev = scene_device(scene);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	if (mode != RTK_CUDA_UPDATE_REFIT && mode != RTK_CUDA_UPDATE_REBUILD) { rtkd_set_error("unknown update mode %d", mode); return RTK_CUDA_ERR_ARGUMENT; }
	if (num_meshes != dev->num_meshes || (num_meshes && !meshes)) { rtkd_set_error("update needs the scene's %u meshes", dev->num_meshes); return RTK_CUDA_ERR_ARGUMENT; }
	for (size_t i = 0; i < num_meshes; i++) {
		if (meshes[i].num_triangles != (size_t)(dev->h_mesh_first[i + 1] - dev->h_mesh_first[i])) {
			rtkd_set_error("mesh %zu changed its triangle count: build a new scene", i);
			return RTK_CUDA_ERR_ARGUMENT;
		}
	}
	int r = RTK_CUDA_OK;
	for (size_t i = 0; i < num_meshes && r == RTK_CUDA_OK; i++)
		r = rtkd_decode_mesh(dev, dev->h_mesh_first[i], (uint32_t)meshes[i].num_triangles, meshes[i].d_positions, 12, 0,
		                     meshes[i].d_indices, 12, 4, 0, stream);
	if (r != RTK_CUDA_OK) return r;
	return mode == RTK_CUDA_UPDATE_REFIT ? rtkd_refit(dev, stream) : rtkd_build(dev, g_build_mode, stream);
}

/* baked instancing: instance i = mesh instances[i].mesh pushed through instances[i].transform */
static int instances_valid(const rtk_cuda_mesh *meshes, size_t num_meshes, const rtk_cuda_instance *inst, size_t n)
{
	if ((num_meshes && !meshes) || (n && !inst)) { rtkd_set_error("bad mesh / instance array"); return 0; }
	for (size_t i = 0; i < n; i++) {
		if (inst[i].mesh >= num_meshes) { rtkd_set_error("instance %zu names mesh %u of %zu", i, inst[i].mesh, num_meshes); return 0; }
		for (int k = 0; k < 12; k++) if (!isfinite(inst[i].transform[k])) { rtkd_set_error("instance %zu has a non-finite transform", i); return 0; }
	}
	return 1;
}

static int decode_instances(rtkd_scene *dev, const uint32_t *first, const rtk_cuda_mesh *meshes,
                            const rtk_cuda_instance *inst, size_t n, void *stream)
{
	int r = RTK_CUDA_OK;
	for (size_t i = 0; i < n && r == RTK_CUDA_OK; i++) {
		const rtk_cuda_mesh *m = &meshes[inst[i].mesh];
		r = rtkd_decode_mesh_xf(dev, first[i], (uint32_t)m->num_triangles, m->d_positions, 12, 0, m->d_indices, 12, 4, 0,
		                        inst[i].transform, stream);
	}
	return r;
}

rtk_scene *rtk_cuda_build_instanced_scene(const rtk_cuda_mesh *meshes, size_t num_meshes,
                                          const rtk_cuda_instance *instances, size_t num_instances, void *stream)
{
	if (!instances_valid(meshes, num_meshes, instances, num_instances)) return NULL;
	uint32_t *first = (uint32_t*)malloc(sizeof(uint32_t) * (num_instances + 1));
	if (!first) return NULL;
	size_t total = 0;
	for (size_t i = 0; i < num_instances; i++) {
		first[i] = (uint32_t)total;
		total += meshes[instances[i].mesh].num_triangles;
		if (total > 0x0fffffffu) { free(first); rtkd_set_error("too many instanced triangles (baked instancing holds at most 2^28 - 1)"); return NULL; }
	}
	first[num_instances] = (uint32_t)total;
	double t0 = now_ms();
	rtkd_scene *dev = rtkd_scene_new((uint32_t)total, (uint32_t)num_instances, first);
	int r = dev ? decode_instances(dev, first, meshes, instances, num_instances, stream) : RTK_CUDA_ERR_CUDA;
	if (r == RTK_CUDA_OK) r = rtkd_build(dev, g_build_mode, stream);
	free(first);
	if (r != RTK_CUDA_OK) { warn_once(); if (dev) rtkd_scene_free(dev); return NULL; }
	dev->build_total_ms = now_ms() - t0;
	return make_handle(dev);
}

int rtk_cuda_update_instanced_scene(const rtk_scene *scene, const rtk_cuda_mesh *meshes, size_t num_meshes,
                                    const rtk_cuda_instance *instances, size_t num_instances, int mode, void *stream)
{
	rtkd_scene *dev = scene_device(scene);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	if (mode != RTK_CUDA_UPDATE_REFIT && mode != RTK_CUDA_UPDATE_REBUILD) { rtkd_set_error("unknown update mode %d", mode); return RTK_CUDA_ERR_ARGUMENT; }
	if (!instances_valid(meshes, num_meshes, instances, num_instances)) return RTK_CUDA_ERR_ARGUMENT;
	if (num_instances != dev->num_meshes) { rtkd_set_error("update needs the scene's %u instances", dev->num_meshes); return RTK_CUDA_ERR_ARGUMENT; }
	for (size_t i = 0; i < num_instances; i++) {
		if (meshes[instances[i].mesh].num_triangles != (size_t)(dev->h_mesh_first[i + 1] - dev->h_mesh_first[i])) {
			rtkd_set_error("instance %zu changed its triangle count: build a new scene", i);
			return RTK_CUDA_ERR_ARGUMENT;
		}
	}
	int r = decode_instances(dev, dev->h_mesh_first, meshes, instances, num_instances, stream);
	if (r != RTK_CUDA_OK) return r;
	return mode == RTK_CUDA_UPDATE_REFIT ? rtkd_refit(dev, stream) : rtkd_build(dev, g_build_mode, stream);
}

int rtk_cuda_get_scene_info(const rtk_scene *scene, rtk_cuda_scene_info *info)
{
	rtkd_scene *dev = scene_device(scene);
	if (!dev || !info) return RTK_CUDA_ERR_SCENE;
	memset(info, 0, sizeof(*info));
	info->num_triangles = dev->num_tris; info->num_meshes = dev->num_meshes;
	info->num_wide_nodes = dev->num_nodes; info->num_leaves = dev->num_leaves;
	info->wide_depth = dev->depth; info->build_mode = dev->build_mode;
	info->device_bytes = 256ull * dev->num_nodes + 48ull * dev->num_tris + 48ull * dev->num_tv + 4ull * (dev->num_meshes + 1);
	info->build_device_ms = dev->build_device_ms; info->build_total_ms = dev->build_total_ms;
	info->sah_cost = dev->sah_cost;
	memcpy(info->bounds_min, dev->bounds_min, 12); memcpy(info->bounds_max, dev->bounds_max, 12);
	return RTK_CUDA_OK;
}

/* ---------------------------------------------------------------------------------------- */
/* queries                                                                                    */
/* ---------------------------------------------------------------------------------------- */

size_t rtk_trace_rays(const rtk_scene *scene, const rtk_ray *rays, rtk_hit *hits, uint8_t *hit_mask, size_t n)
{
	rtkd_scene *dev = scene_device(scene);
	if (!dev) return (size_t)-1;
	if (n && (!rays || !hits)) { rtkd_set_error("rays / hits is NULL"); return (size_t)-1; }
	long long r = rtkd_trace_host(dev, rays, hits, hit_mask, n);
	if (r < 0) { warn_once(); return (size_t)-1; }
	return (size_t)r;
}

int rtk_trace_rays_compact(const rtk_scene *scene, const rtk_ray *rays, rtk_cuda_hit16 *hits, size_t n)
{
	rtkd_scene *dev = scene_device(scene);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	if (n && (!rays || !hits)) { rtkd_set_error("rays / hits is NULL"); return RTK_CUDA_ERR_ARGUMENT; }
	int r = rtkd_trace_host_compact(dev, rays, hits, n);
	if (r) warn_once();
	return r;
}

int rtk_trace_rays_compact_device(const rtk_scene *scene, const void *d_rays, void *d_hit16, size_t n, void *stream)
{
	rtkd_scene *dev = scene_on_device_of(scene, d_rays);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	return rtkd_trace(dev, d_rays, d_hit16, n, g_cull_mode, NULL, stream);
}

int rtk_resolve_hits_device(const rtk_scene *scene, const void *d_hit16, void *d_hits, void *d_hit_mask, size_t n, void *stream)
{
	rtkd_scene *dev = scene_on_device_of(scene, d_hit16);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	return rtkd_resolve(dev, d_hit16, d_hits, d_hit_mask, n, stream);
}

int rtk_trace_rays_device(const rtk_scene *scene, const void *d_rays, void *d_hits, void *d_hit_mask, size_t n, void *stream)
{
	rtkd_scene *dev = scene_on_device_of(scene, d_rays);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	void *h16 = rtkd_scene_hit16(dev, n);
	if (!h16 && n) return RTK_CUDA_ERR_MEMORY;
	int r = rtkd_trace(dev, d_rays, h16, n, g_cull_mode, NULL, stream);
	if (r) return r;
	return rtkd_resolve(dev, h16, d_hits, d_hit_mask, n, stream);
}

int rtk_occluded_rays_device(const rtk_scene *scene, const void *d_rays, void *d_occluded, size_t n, void *stream)
{
	rtkd_scene *dev = scene_on_device_of(scene, d_rays);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	return rtkd_trace(dev, d_rays, d_occluded, n, g_cull_mode | 2, NULL, stream);
}

int rtk_cuda_peer_window_create(size_t bytes, void **d_window, rtk_cuda_peer_handle *handle)
{
	return rtkd_peer_create(bytes, d_window, handle ? handle->bytes : NULL);
}
int rtk_cuda_peer_window_open(const rtk_cuda_peer_handle *handle, void **d_window)
{
	return rtkd_peer_open(handle ? handle->bytes : NULL, d_window);
}
int rtk_cuda_peer_window_close(void *d_window) { return rtkd_peer_close(d_window); }
int rtk_cuda_peer_window_destroy(void *d_window) { return rtkd_peer_destroy(d_window); }
int rtk_cuda_peer_push(void *d_dst, const void *d_src, size_t bytes, void *stream) { return rtkd_peer_push(d_dst, d_src, bytes, stream); }

int rtk_cuda_set_triangle_filter(const rtk_scene *scene, const uint32_t *bits, size_t num_words)
{
	rtkd_scene *dev = scene_device(scene);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	return rtkd_set_filter(dev, bits, num_words, 0, NULL);
}

int rtk_cuda_set_triangle_filter_device(const rtk_scene *scene, const void *d_bits, size_t num_words, void *stream)
{
	rtkd_scene *dev = scene_device(scene);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	return rtkd_set_filter(dev, d_bits, num_words, 1, stream);
}

int rtk_cuda_generate_primary_rays(const rtk_cuda_camera *camera, uint64_t seed, uint32_t sample,
                                   size_t first_pixel, size_t count, void *d_rays, void *stream)
{
	if (!camera || (count && !d_rays)) { rtkd_set_error("bad camera / ray buffer"); return RTK_CUDA_ERR_ARGUMENT; }
	if (sample >= 64) { rtkd_set_error("sample number must be below 64"); return RTK_CUDA_ERR_ARGUMENT; }
	float cam[13];
	memcpy(cam, camera->eye, 12); memcpy(cam + 3, camera->forward, 12);
	memcpy(cam + 6, camera->right, 12); memcpy(cam + 9, camera->up, 12);
	cam[12] = camera->tan_half_fov;
	return rtkd_gen_primary(cam, camera->width, camera->height, seed, sample, first_pixel, count, d_rays, stream);
}

int rtk_cuda_generate_bounce_rays(const rtk_scene *scene, const void *d_rays_in, const void *d_hit16, void *d_rays_out,
                                  void *d_alive, size_t n, uint64_t seed, uint32_t bounce, uint64_t first_ray, uint32_t flags, void *stream)
{
	rtkd_scene *dev = scene_on_device_of(scene, d_rays_in);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	if (bounce >= 16) { rtkd_set_error("bounce number must be below 16"); return RTK_CUDA_ERR_ARGUMENT; }
	return rtkd_gen_bounce(dev, d_rays_in, d_hit16, d_rays_out, d_alive, n, seed, bounce, first_ray, flags, stream);
}

int rtk_trace_rays_bruteforce_device(const rtk_scene *scene, const void *d_rays, void *d_hit16, size_t n, void *stream)
{
	rtkd_scene *dev = scene_on_device_of(scene, d_rays);
	if (!dev) return RTK_CUDA_ERR_SCENE;
	return rtkd_trace_brute(dev, d_rays, d_hit16, n, stream);
}

int rtk_trace_stats_device(const rtk_scene *scene, const void *d_rays, void *d_hit16, size_t n, rtk_cuda_trace_stats *stats, void *stream)
{
	rtkd_scene *dev = stats ? scene_on_device_of(scene, d_rays) : NULL;
	if (!dev) return RTK_CUDA_ERR_SCENE;
	rtkd_trace_stats st;
	int r = rtkd_trace(dev, d_rays, d_hit16, n, g_cull_mode, &st, stream);
	stats->rays = st.rays; stats->hits = st.hits; stats->node_visits = st.node_visits;
	stats->leaf_visits = st.leaf_visits; stats->tri_tests = st.tri_tests; stats->stack_max = st.stack_max;
	return r;
}

bool rtk_trace_ray(const rtk_scene *scene, const rtk_ray *ray, rtk_hit *hit)
{
	/* rtk.c:543-577 as a one-ray batch; *hit is written only on a hit (rtk.c:571-576) */
	rtk_hit tmp;
	uint8_t m = 0;
	size_t r = rtk_trace_rays(scene, ray, &tmp, &m, 1);
	if (r == (size_t)-1 || !m) return false;
	*hit = tmp;
	return true;
}

bool rtk_trace_ray_filter(const rtk_scene *scene, const rtk_ray *ray, rtk_hit *hit, rtk_filter_fn *filter, void *filter_user)
{
	/* the reference's version is a stub (rtk.c:579-582) */
	rtk_hit tmp;
	if (!rtk_trace_ray(scene, ray, &tmp)) return false;
	if (filter && !filter(filter_user, ray, &tmp)) return false;
	*hit = tmp;
	return true;
}
