/*
 * rtk_device.h -- the thin C-ABI layer between the C host code (rtk_host.c) and the CUDA
 * translation unit (rtk_device.cu).  Internal: not installed, not part of the public ABI.
 * Plain C types only.
 */
#ifndef RTK_DEVICE_H
#define RTK_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTKD_MAX_DEVICES 16
#define RTKD_TRACE_SLOTS 4

/* What one traversal launch needs for itself: ray cursor, statistics, the global-memory part of the
 * traversal stack.  A scene owns a few of them so that queries from different streams and host
 * threads never share mutable state (rtk_trace_ray is re-entrant on a const scene, rtk.h:129). */
typedef struct rtkd_trace_slot {
	void *scratch;               /* 256 bytes: +0 ray cursor | +64 stats[6] | +128 hit counter */
	void *overflow; size_t overflow_entries, overflow_groups;
	void *done;                  /* cudaEvent_t: recorded after the slot's last launch */
	void *last_stream;           /* stream of that launch */
	int   used;
} rtkd_trace_slot;

/* A scene resident on ONE device.  Pointers are device pointers.  With several devices in use
 * (rtk_cuda_init_devices) the scene built on the first device carries one replica per further
 * device: the same struct with that device's pointers. */
typedef struct rtkd_scene {
	uint64_t id;                 /* unique per process, also stored in serialised blobs */
	uint32_t num_tris, num_meshes, num_nodes, num_leaves, depth, build_mode;
	void *tri_orig;              /* float4[3*num_tris] */
	/* what the traversal kernel reads lives in ONE allocation (nodes | tv0 | tv1 | tv2), so that one
	 * L2 access-policy window covers it */
	void *arena; size_t arena_cap, arena_used;
	void *nodes;                 /* float4[16*num_nodes] */
	void *tv0, *tv1, *tv2;       /* float4[num_tv] each: one 8-entry slot per leaf */
	uint32_t num_tv;
	unsigned char *node_level;   /* uint8[num_nodes]: depth of each wide node (NULL for a scene loaded from a blob) */
	uint32_t node_level_cap;
	void *mesh_first;            /* uint32[num_meshes+1] */
	uint32_t *h_mesh_first;      /* host copy */
	float bounds_min[3], bounds_max[3], abs_max;
	double build_device_ms, build_total_ms, sah_cost;
	double upload_ms;            /* host-to-device part of build_total_ms */
	/* traversal scratch, created on first use */
	rtkd_trace_slot slot[RTKD_TRACE_SLOTS];
	void *slot_lock;             /* pthread_mutex_t* */
	uint32_t *h_status;          /* pinned + mapped: bit 1 = traversal stack exhausted (sticky until the next build) */
	uint32_t *d_status;          /* device view of h_status */
	void *hit16; size_t hit16_cap;   /* compact hits of rtk_trace_rays_device */
	void *filter_bits;           /* uint32[(num_tris+31)/32] or NULL: triangle filter baked into the leaf slots */
	/* multi-device */
	int dev_index;               /* index into the library's device list (0 = the device the scene was built on) */
	uint64_t epoch;              /* bumped by every change of the device arrays */
	struct rtkd_scene *replica[RTKD_MAX_DEVICES];   /* [k] = copy on device k (k >= 1), or NULL */
	uint64_t replica_epoch;      /* epoch the replicas were copied at */
} rtkd_scene;

typedef struct rtkd_trace_stats {
	uint64_t rays, hits, node_visits, leaf_visits, tri_tests, stack_max;
} rtkd_trace_stats;

int         rtkd_init(int device);           /* 0 or negative rtk_cuda_status */
int         rtkd_init_devices(const int *devices, int n);   /* devices[0] builds; host batches are split over all n */
int         rtkd_device_count(void);         /* devices in use (0 before initialisation) */
int         rtkd_bind_thread(void);          /* make the library's device current on the calling thread (initialises device 0 if needed) */
void        rtkd_shutdown(void);
const char *rtkd_last_error(void);
void        rtkd_set_error(const char *fmt, ...);
int         rtkd_reserve_sms(int sms);       /* the traversal grid leaves this many SMs free */
int         rtkd_device_info(int *sm_count, size_t *l2_bytes, int *ctas_per_sm, int *threads_per_cta);

int         rtkd_read_bandwidth(size_t bytes, int passes, double *gbs);   /* read probe: L2 (small buffer) or HBM */
int         rtkd_gather_bandwidth(size_t bytes, size_t record_bytes, int passes, double *gbs);   /* random record gathers */
/* host link probe: concurrent pinned-memory copies on the first `ndev` devices; dir 1 = up, 2 = down, 3 = both */
int         rtkd_link_bandwidth(int ndev, size_t bytes_per_device, int dir, int passes, double *gbs);

/* page-locked host memory the devices can read and WRITE directly (rows of rtk_trace_rays land in it without staging) */
void       *rtkd_host_alloc(size_t bytes);
void        rtkd_host_free(void *p);
void       *rtkd_host_alloc_batch(size_t elem_bytes, size_t count);   /* array for batches of `count` rays: each device's share on its NUMA node */
int         rtkd_host_register(void *p, size_t bytes);
int         rtkd_host_unregister(void *p);

rtkd_scene *rtkd_scene_new(uint32_t num_tris, uint32_t num_meshes, const uint32_t *mesh_first);
void        rtkd_scene_free(rtkd_scene *s);

/* staging helpers for the host ingest */
void *rtkd_upload(const void *host, size_t bytes, void *stream);
void  rtkd_free_async(void *dev, void *stream);
int   rtkd_sync(void *stream);
int   rtkd_max_index(const void *d_idx, size_t stride, int idx_bytes, uint32_t ntris, uint32_t *out, void *stream);

/* decode one mesh (device buffers) into the scene's corner records */
int rtkd_decode_mesh(rtkd_scene *s, uint32_t first_prim, uint32_t ntris,
                     const void *d_pos, size_t pos_stride, int pos_f64,
                     const void *d_idx, size_t idx_stride, int idx_bytes, int pregathered, void *stream);

/* the same with an instance transform baked in: xf12 is a row-major 3x4 matrix (NULL = none) */
int rtkd_decode_mesh_xf(rtkd_scene *s, uint32_t first_prim, uint32_t ntris,
                        const void *d_pos, size_t pos_stride, int pos_f64,
                        const void *d_idx, size_t idx_stride, int idx_bytes, int pregathered,
                        const float *xf12, void *stream);

/* build the BVH from the decoded triangles; synchronous with respect to `stream` on return */
int rtkd_build(rtkd_scene *s, int mode, void *stream);

/* the vertices of the decoded triangles moved, the topology did not: refit all boxes in place */
int rtkd_refit(rtkd_scene *s, void *stream);

/* triangle filter: bit i of the bitset = triangle i takes part in every query on
 * this scene.  bits == NULL removes it.  on_device: bits is a device pointer.  Survives refits
 * and rebuilds of the scene. */
int rtkd_set_filter(rtkd_scene *s, const void *bits, size_t num_words, int on_device, void *stream);

/* queries: device pointers, asynchronous on stream.  cull_mode bit 0: provable culling,
 * bit 1: occlusion query (d_hit16 is then one byte per ray) */
int rtkd_trace(rtkd_scene *s, const void *d_rays, void *d_hit16, size_t n, int cull_mode,
               rtkd_trace_stats *stats, void *stream);
int rtkd_trace_brute(rtkd_scene *s, const void *d_rays, void *d_hit16, size_t n, void *stream);
int rtkd_resolve(rtkd_scene *s, const void *d_hit16, void *d_hits, void *d_mask, size_t n, void *stream);
void *rtkd_scene_hit16(rtkd_scene *s, size_t n);   /* scene-owned compact hit buffer of >= n records */
/* the copy of the scene on the device that owns device pointer `p` (replicas are brought up to date);
 * that device becomes current on the calling thread.  NULL: p belongs to no device in use */
rtkd_scene *rtkd_scene_for_pointer(rtkd_scene *s, const void *p);
int   rtkd_sync_replicas(rtkd_scene *s);           /* copy the scene to every further device in use */
uint32_t rtkd_scene_status(rtkd_scene *s);         /* sticky error bits of the scene on all its devices */
int   rtkd_debug_limit_stack(int entries);         /* test hook: global-memory stack entries per ray (0 = sized from the tree) */

/* hit gather over NVLink peer memory (CUDA IPC): a window owned by the gathering process, opened
 * by the others, filled with copy-engine pushes.  handle64 is a cudaIpcMemHandle_t. */
int rtkd_peer_create(size_t bytes, void **d_window, unsigned char *handle64);
int rtkd_peer_open(const unsigned char *handle64, void **d_window);
int rtkd_peer_close(void *d_window);
int rtkd_peer_destroy(void *d_window);
int rtkd_peer_push(void *d_dst, const void *d_src, size_t bytes, void *stream);

/* wavefront ray generation: cam20 = eye, forward, right, up (3 floats each), tan(half vertical fov) */
int rtkd_gen_primary(const float *cam20, uint32_t width, uint32_t height, unsigned long long seed, uint32_t sample,
                     unsigned long long first_pixel, size_t count, void *d_rays, void *stream);
int rtkd_gen_bounce(rtkd_scene *s, const void *d_rays_in, const void *d_hit16, void *d_rays_out, void *d_alive,
                    size_t n, unsigned long long seed, uint32_t bounce, unsigned long long first_ray, uint32_t flags, void *stream);

/* host-buffer batch: H2D, trace, resolve, D2H; returns hits or -1 */
long long rtkd_trace_host(rtkd_scene *s, const void *rays, void *hits, unsigned char *mask, size_t n);

/* host-buffer batch with compact results: one 16-byte record per ray straight into the caller's array */
int rtkd_trace_host_compact(rtkd_scene *s, const void *rays, void *hit16, size_t n);

/* host placement of dense hit rows (rtk_place.c): for every ray i of the chunk with mask[i] != 0
 * the next 68-byte row of its 128-ray block -- block b's rows start at rows[block_base[b]] --
 * is copied to hits[first_ray + i]; mask_out (may be NULL) receives the mask bytes. */
typedef struct rtkd_place_desc {
	void *hits;                       /* caller's rtk_hit array (whole batch) */
	unsigned char *mask_out;          /* caller's mask array (whole batch) or NULL */
	const void *rows;                 /* dense rows of this chunk */
	const unsigned char *mask;        /* mask bytes of this chunk */
	const uint32_t *block_base;       /* first row of each 128-ray block */
	size_t first_ray, nrays;          /* the chunk's position in the batch */
} rtkd_place_desc;
int  rtkd_place_submit(const rtkd_place_desc *d);   /* returns a ticket (or -1: done synchronously) */
void rtkd_place_wait(int ticket);
int  rtkd_copy_submit(void *dst, const void *src, size_t bytes);   /* parallel memcpy on the placement pool; ticket for rtkd_place_wait, or -1 (done) */
int  rtkd_place_threads(void);

/* serialisation of the device layout into a relocatable blob (payload after the 128-byte
 * header block that rtk_host.c writes) */
size_t      rtkd_blob_payload_size(const rtkd_scene *s);
int         rtkd_blob_write(const rtkd_scene *s, void *payload);          /* device -> host */
rtkd_scene *rtkd_blob_read(const void *payload, size_t payload_size);     /* host -> device */

#ifdef __cplusplus
}
#endif
#endif
