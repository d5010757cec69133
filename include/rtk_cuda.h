/*
 * rtk_cuda.h -- batched, data-parallel extension of the rtk.h C ABI.
 *
 * The reference offers only a single-ray query (rtk.h:129, rtk.c:543-577) that
 * the user calls in a loop from their own threads.  On a GPU the unit of work
 * is a batch, so this header adds the entry points the benchmark and any
 * throughput-minded caller drive.  Everything is extern "C", plain pointers and
 * sizes; CUDA streams cross the boundary as void* (a cudaStream_t / CUstream).
 *
 * Each function names the reference code it replaces.  All of them return 0 on
 * success and a negative rtk_cuda_status on failure unless stated otherwise;
 * rtk_cuda_last_error() returns a thread-local, human-readable reason.
 */
#ifndef RTK_B200_RTK_CUDA_H
#define RTK_B200_RTK_CUDA_H

#include "rtk.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rtk_cuda_status {
	RTK_CUDA_OK            =  0,
	RTK_CUDA_ERR_NO_DEVICE = -1,  /* no CUDA device / driver: there is NO CPU fallback */
	RTK_CUDA_ERR_CUDA      = -2,  /* a CUDA runtime call failed */
	RTK_CUDA_ERR_ARGUMENT  = -3,
	RTK_CUDA_ERR_SCENE     = -4,  /* blob is not a scene written by this library */
	RTK_CUDA_ERR_MEMORY    = -5,
	RTK_CUDA_ERR_OVERFLOW  = -6,  /* traversal stack exhausted (never seen; reported, not ignored) */
} rtk_cuda_status;

/* 16-byte compact hit record produced by the traversal kernel and consumed by
 * the resolve kernel; prim is the global triangle number (meshes concatenated
 * in rtk_scene_desc order, the numbering of rtk.c:1131-1170) or RTK_CUDA_MISS. */
#define RTK_CUDA_MISS 0xffffffffu
typedef struct rtk_cuda_hit16 {
	float    t, u, v;
	uint32_t prim;
} rtk_cuda_hit16;

/* Builder selection for rtk_cuda_set_build_mode(). */
typedef enum rtk_cuda_build_mode {
	RTK_CUDA_BUILD_LBVH = 0,      /* Morton radix sort + Karras hierarchy */
	RTK_CUDA_BUILD_SAH  = 1,      /* binned SAH (32 bins x 3 axes, rtk.c:867-1019) on top of the Morton order */
} rtk_cuda_build_mode;

/* What a built scene looks like on the device (rtk_cuda_get_scene_info). */
typedef struct rtk_cuda_scene_info {
	uint64_t num_triangles;
	uint64_t num_meshes;
	uint64_t num_wide_nodes;      /* 256-byte 8-wide nodes */
	uint64_t num_leaves;
	uint32_t wide_depth;          /* levels of 8-wide nodes */
	uint32_t build_mode;
	uint64_t device_bytes;        /* resident bytes of the scene on one GPU */
	double   build_device_ms;     /* CUDA-event time of the device build (no H2D) */
	double   build_total_ms;      /* wall time of ingest + H2D + build */
	double   sah_cost;            /* SAH cost of the wide BVH (node cost 1, leaf cost ceil(n/8)) */
	float    bounds_min[3];
	float    bounds_max[3];
} rtk_cuda_scene_info;

/* Per-batch traversal counters (rtk_cuda_trace_stats_device). */
typedef struct rtk_cuda_trace_stats {
	uint64_t rays;
	uint64_t hits;
	uint64_t node_visits;         /* 256-byte wide-node fetches */
	uint64_t leaf_visits;         /* leaf records opened */
	uint64_t tri_tests;           /* triangles pushed through the watertight test */
	uint64_t stack_max;           /* deepest traversal stack seen */
} rtk_cuda_trace_stats;

/* ---- device selection ------------------------------------------------- */

/* Bind the calling process to one CUDA device (one process per GPU).  Called
 * implicitly with device 0 by the first entry point that needs a device.  CUDA's
 * current device is per host thread: every entry point makes the library's device
 * current on the thread that calls it, so worker threads need no set-up of their own. */
int  rtk_cuda_init(int device);
/* Drive SEVERAL devices from this one process, the way a reference user gets the whole machine from one
 * rtk_build_scene / rtk_trace_ray call site (rtk.h:126-129): devices[0] builds every scene, the scene is
 * copied to the other devices (device to device, NVLink between peers) by rtk_build_scene, host batches
 * (rtk_trace_rays, rtk_trace_rays_compact) are split into contiguous ranges over all of them -- one upload
 * stream, one chunk pipeline and one worker thread per device -- and every device entry point runs on
 * the device that owns the caller's ray buffer.  rtk_free_scene frees the copies on every device.  The
 * devices must be of one kind.  rtk_cuda_init(d) is the list {d}; a different list needs
 * rtk_cuda_shutdown() first. */
int  rtk_cuda_init_devices(const int *devices, int num_devices);
int  rtk_cuda_device_count(void);
/* Page-locked host memory every device of the list reads and WRITES in place.  When the hits / hit_mask
 * arrays given to rtk_trace_rays are page-locked (these calls, cudaHostAlloc, cudaHostRegister, ...) the
 * rows of the rays that hit travel straight from the resolve kernel into the caller's array: no staging
 * copy, no host threads.  Pageable arrays work too, through pinned staging and a pool of host threads. */
void *rtk_cuda_host_alloc(size_t bytes);
void  rtk_cuda_host_free(void *p);
/* The same for an array that rtk_trace_rays / rtk_trace_rays_compact will split over SEVERAL devices: an
 * array of `count` elements (rays: 32 bytes, rtk_hit: 68, mask: 1, rtk_cuda_hit16: 16) whose part for
 * device k -- the split a batch of `count` rays gets -- is placed on device k's NUMA node before it is
 * page-locked, so that each link reads and writes the memory of its own socket.  Free with rtk_cuda_host_free. */
void *rtk_cuda_host_alloc_batch(size_t element_bytes, size_t count);
int   rtk_cuda_host_register(void *p, size_t bytes);
int   rtk_cuda_host_unregister(void *p);
/* Releases what the library itself holds on the device (streams, events, device and pinned
 * staging of the host batches).  Scenes are the caller's: free them first.  A later call
 * initialises the library again. */
void rtk_cuda_shutdown(void);
const char *rtk_cuda_last_error(void);
int  rtk_cuda_set_build_mode(int mode);
/* Distance used to cull a box against the current best hit: 1 (default) = slab of the ray's
 * dominant axis only, provably consistent with the fp32 watertight test; 0 = full box entry,
 * tighter but not provable for triangles seen exactly edge-on. */
int  rtk_cuda_set_cull_mode(int mode);
/* The traversal kernel is persistent and fills every SM.  A process that runs other kernels beside
 * it -- NCCL's send/receive kernels of a hit gather that overlaps the next batch -- can keep `sms`
 * SMs out of the traversal grid so that those kernels are not queued behind it.  Default 0
 * (with a 268 MB gather per step at 2 GPUs reserving 4 SMs cost 2 % and bought nothing). */
int  rtk_cuda_reserve_sms(int sms);
/* SM count, L2 bytes, resident CTAs of the traversal kernel... for the bench. */
int  rtk_cuda_device_info(int *sm_count, size_t *l2_bytes, int *trace_ctas_per_sm, int *trace_threads_per_cta);

/* Read-bandwidth probe for roofline denominators: streams a `bytes`-sized device buffer `passes`
 * times with 16-byte loads.  A buffer well below the L2 size measures L2 read bandwidth, one well
 * above it HBM read bandwidth. */
int  rtk_cuda_measure_read_bandwidth(size_t bytes, int passes, double *gb_per_s);
/* The access pattern of the traversal rather than of a copy: warps gather `record_bytes`-sized records
 * (256 = a wide node, 128 = one line of a leaf slot) at hashed offsets of a `bytes`-sized buffer.  With
 * the buffer inside the L2 this is the denominator of the L2 roofline. */
int  rtk_cuda_measure_gather_bandwidth(size_t bytes, size_t record_bytes, int passes, double *gb_per_s);
/* Ceiling of the host-buffer path: aggregate GB/s of concurrent copies between pinned host memory and the
 * first num_devices devices of the list; directions: 1 host-to-device, 2 device-to-host, 3 both at once. */
int  rtk_cuda_measure_host_link(int num_devices, size_t bytes_per_device, int directions, int passes, double *gb_per_s);

/* ---- batched closest hit (replaces a user loop over rtk_trace_ray) ----- */

/* Host buffers.  rays[n] in, hits[n] / hit_mask[n] out.  hits[i] is written
 * only where hit_mask[i] != 0; rows of rays that missed are left untouched,
 * the miss rule of rtk.c:571-576.  hit_mask may be NULL.  The batch runs as a
 * pipeline over 1M-ray chunks: rays go up, only the rows of rays that hit
 * (plus one mask byte per ray) come back.  With page-locked hits / hit_mask arrays
 * (rtk_cuda_host_alloc[_batch], cudaHostAlloc, cudaHostRegister ...) the device writes the rows
 * straight into them; with pageable (malloc'ed) arrays rays and rows travel through pinned
 * staging, moved by a few library threads (RTK_B200_HOST_THREADS, default 3/4 of the CPUs, at
 * most 16) -- about 60 % of the speed.  With several devices in use (rtk_cuda_init_devices) the
 * batch is split over all of them.  Thread-safe; batches of different host threads queue per device.
 * Returns the number of hits, or (size_t)-1 on error. */
size_t rtk_trace_rays(const rtk_scene *scene, const rtk_ray *rays, rtk_hit *hits, uint8_t *hit_mask, size_t n);

/* Host buffers, compact results: hits[i] is a 16-byte record for EVERY ray --
 * t, u, v and the global triangle number, or prim == RTK_CUDA_MISS (t, u, v = 0)
 * -- for callers that look the vertices up themselves (or only need
 * distances / visibility).  Half the bytes of the rays come back and nothing is
 * repacked on the host, so the batch is bound by the upload of the rays alone.
 * Returns 0 or a negative rtk_cuda_status. */
int rtk_trace_rays_compact(const rtk_scene *scene, const rtk_ray *rays, rtk_cuda_hit16 *hits, size_t n);

/* Device buffers (16-byte aligned), asynchronous on `stream`.  d_hits[i] is
 * written only where d_hit_mask[i] != 0.
 *
 * Concurrency of every *_device query below: like rtk_trace_ray (rtk.h:129) they are re-entrant on a
 * scene -- any number of host threads and streams may query one scene at once; each launch takes its
 * own cursor / stack scratch from the scene.  With several devices in use (rtk_cuda_init_devices) the
 * call runs on the device that owns d_rays; `stream` must belong to that device.  What must NOT overlap
 * a query, exactly as freeing a reference scene under a running rtk_trace_ray must not: rtk_free_scene,
 * rtk_cuda_rebuild_scene / update / set_triangle_filter on the same scene.  rtk_trace_rays_device
 * additionally uses one scene-owned buffer of compact records: one such call per scene at a time (use
 * the two halves below with your own buffer otherwise).  A traversal that ran out of stack sets a sticky
 * flag instead of returning wrong hits silently: rtk_cuda_scene_status reads it. */
int rtk_trace_rays_device(const rtk_scene *scene, const void *d_rays, void *d_hits, void *d_hit_mask, size_t n, void *stream);

/* The two halves of the call above: traversal to compact records
 * (rtk.c:390-539 + 181-388), then expansion to rtk_hit (rtk.c:371-381). */
int rtk_trace_rays_compact_device(const rtk_scene *scene, const void *d_rays, void *d_hit16, size_t n, void *stream);
int rtk_resolve_hits_device(const rtk_scene *scene, const void *d_hit16, void *d_hits, void *d_hit_mask, size_t n, void *stream);

/* RTK_CUDA_OK, or RTK_CUDA_ERR_OVERFLOW when some traversal on this scene exhausted its stack since the
 * scene was last built (cannot happen for trees this library builds: the stack is sized from the tree's
 * depth; a defence for blobs).  Meaningful once the caller has synchronised the streams it queried on;
 * the host-buffer entry points check it themselves. */
int rtk_cuda_scene_status(const rtk_scene *scene);

/* Test hook: cap the global-memory part of the traversal stack at `entries` per ray (0 = sized from the
 * tree depth, the default) so that the overflow report can be exercised. */
int rtk_cuda_debug_limit_stack(int entries);

/* Occlusion (shadow-ray) query: d_occluded[i] = 1 iff some triangle is hit in
 * (min_t, max_t), i.e. exactly where the closest-hit query reports a hit; the
 * traversal stops at the first accepted triangle.  The any-hit half of what
 * rtk_trace_ray_filter (rtk.h:130, a stub upstream) is for. */
int rtk_occluded_rays_device(const rtk_scene *scene, const void *d_rays, void *d_occluded, size_t n, void *stream);

/* Device-side triangle predicate (the filtered half of what
 * rtk_trace_ray_filter, rtk.h:117,130, is for): a bitset over the scene's GLOBAL
 * triangle numbers (mesh_first[mesh_index] + triangle_index, meshes in
 * rtk_scene_desc order, rtk.c:1131-1170).  Bit (i & 31) of word (i >> 5) set =
 * triangle i takes part; a cleared bit removes the triangle from EVERY later
 * query on this scene (closest hit, occlusion, host and device entry points)
 * exactly as if it had not been in the mesh, while triangle numbering stays as
 * it was.  num_words >= ceil(num_triangles / 32); bits == NULL removes the
 * filter.  The predicate is baked into the leaf slots by one pass over them
 * (16 + 16 bytes per triangle), so traversal costs the same with or without a
 * filter; it survives rtk_cuda_update_scene / rtk_cuda_rebuild_scene.  Not
 * stored in blobs.  Must not run concurrently with a query on the scene. */
int rtk_cuda_set_triangle_filter(const rtk_scene *scene, const uint32_t *bits, size_t num_words);
int rtk_cuda_set_triangle_filter_device(const rtk_scene *scene, const void *d_bits, size_t num_words, void *stream);

/* Exhaustive ray x triangle kernel with the same arithmetic: the GPU-side
 * check used by the parity tests and by the bench's self-check. */
int rtk_trace_rays_bruteforce_device(const rtk_scene *scene, const void *d_rays, void *d_hit16, size_t n, void *stream);

/* Counter-instrumented traversal (slow; for the algorithmic-bytes figure). */
int rtk_trace_stats_device(const rtk_scene *scene, const void *d_rays, void *d_hit16, size_t n, rtk_cuda_trace_stats *stats, void *stream);

/* ---- multi-GPU hit gather over NVLink peer memory ---------- */

/* The sharded path (one process per GPU, scene replicated, rays partitioned)
 * has ONE exchange step: the compact hit records end up on the gathering rank.
 * Besides NCCL (torch.distributed in the bench) the library offers the gather
 * as plain peer-memory pushes: the gathering process creates a window in its
 * HBM and hands the 64-byte handle (a CUDA IPC handle) to the other processes
 * of the box by whatever channel it has; they open it once and then push their
 * records into their slice with the copy engines -- no send/receive kernels
 * competing with the persistent traversal grid for SMs.  Ordering is the
 * caller's: a push is asynchronous on `stream` of the pushing process; the
 * gathering process may read a slice once the pusher has synchronised that
 * stream and told it so (a barrier).  The reference has no counterpart
 * (single-ray, single-process API). */
typedef struct rtk_cuda_peer_handle { unsigned char bytes[64]; } rtk_cuda_peer_handle;
int rtk_cuda_peer_window_create(size_t bytes, void **d_window, rtk_cuda_peer_handle *handle);
int rtk_cuda_peer_window_open(const rtk_cuda_peer_handle *handle, void **d_window);   /* in ANOTHER process */
int rtk_cuda_peer_window_close(void *d_window);                                        /* what _open returned */
int rtk_cuda_peer_window_destroy(void *d_window);                                      /* what _create returned */
int rtk_cuda_peer_push(void *d_dst, const void *d_src, size_t bytes, void *stream);

/* ---- wavefront ray generation on the device ------------ */

/* Pinhole camera: right/up need not be unit length; rays are
 * forward + sx*right + sy*up with sx,sy in [-tan, tan] (x scaled by the aspect). */
typedef struct rtk_cuda_camera {
	float eye[3], forward[3], right[3], up[3];
	float tan_half_fov;           /* vertical */
	uint32_t width, height;
} rtk_cuda_camera;

/* Jittered primary rays of pixels [first_pixel, first_pixel+count), row-major,
 * sample number `sample` (< 64); counter-based jitter from `seed`. */
int rtk_cuda_generate_primary_rays(const rtk_cuda_camera *camera, uint64_t seed, uint32_t sample,
                                   size_t first_pixel, size_t count, void *d_rays, void *stream);

/* Next-bounce rays from the compact hits of the previous bounce: origin = hit
 * point pushed off the surface along the geometric normal (turned towards the
 * incoming ray), direction = cosine-weighted hemisphere sample.  A path that
 * missed gets a dead ray (max_t = 0) and d_alive[i] = 0, or -- with
 * RTK_CUDA_BOUNCE_RELAUNCH in flags -- restarts on a uniformly chosen triangle
 * (d_alive[i] = 2), which keeps the ray count of a wavefront fixed.  Paths that
 * hit have d_alive[i] = 1; d_alive may be NULL.  first_ray numbers the rays for
 * the random-number counter; bounce < 16. */
#define RTK_CUDA_BOUNCE_RELAUNCH 1u
int rtk_cuda_generate_bounce_rays(const rtk_scene *scene, const void *d_rays_in, const void *d_hit16, void *d_rays_out,
                                  void *d_alive, size_t n, uint64_t seed, uint32_t bounce, uint64_t first_ray,
                                  uint32_t flags, void *stream);

/* ---- build with inputs already resident in HBM ------------------------ */

/* A mesh whose buffers live on the device: float xyz positions (tightly
 * packed) and optional uint32 index triples (NULL => 3i,3i+1,3i+2).
 * num_vertices is needed because the device cannot infer it. */
typedef struct rtk_cuda_mesh {
	const void *d_positions;
	const void *d_indices;
	size_t      num_vertices;
	size_t      num_triangles;
} rtk_cuda_mesh;

/* Device-to-device counterpart of rtk_build_scene (rtk.c:1788-1792): builds on
 * `stream`, returns a library-owned scene (free with rtk_free_scene). */
rtk_scene *rtk_cuda_build_scene(const rtk_cuda_mesh *meshes, size_t num_meshes, void *stream);

/* Re-run the device build of an existing scene from its resident decoded
 * triangles (bench loop for build Mtris/s; also refits after nothing moved). */
int rtk_cuda_rebuild_scene(const rtk_scene *scene, void *stream);

/* Animated meshes: new vertex positions for the scene's meshes, same triangle
 * counts.  RTK_CUDA_UPDATE_REFIT keeps the tree and recomputes every box bottom-up (fast; the tree
 * quality decays as the mesh deforms), RTK_CUDA_UPDATE_REBUILD builds a new tree.  Results after
 * either are exactly those of a scene built from the new positions: the BVH is not part of the
 * contract.  build_device_ms of rtk_cuda_get_scene_info reports the update's device time. */
#define RTK_CUDA_UPDATE_REFIT 0
#define RTK_CUDA_UPDATE_REBUILD 1
int rtk_cuda_update_scene(const rtk_scene *scene, const rtk_cuda_mesh *meshes, size_t num_meshes, int mode, void *stream);

/* Multi-mesh instancing (the other half of SURVEY 8(f) N4), BAKED: every instance is its mesh pushed
 * through a 3x4 transform (row-major, world = M * (x, y, z, 1), each product and sum rounded on
 * its own in fp32) into the scene's own triangle storage, and ONE tree is built over all of them --
 * traversal stays single-level and as fast as for any other scene, at 112 bytes of HBM per
 * instanced triangle (180 GB per GPU is the budget this trades against; up to 2^28 - 1 triangles per scene).
 * In a hit, mesh_index is the INSTANCE number, triangle_index the triangle within its mesh and the
 * vertices are in world space.  rtk_cuda_update_instanced_scene moves the instances (same
 * instance -> mesh assignment, same triangle counts) and refits or rebuilds as
 * rtk_cuda_update_scene does. */
typedef struct rtk_cuda_instance {
	uint32_t mesh;                /* index into the mesh array */
	float    transform[12];       /* row-major 3x4 */
} rtk_cuda_instance;
rtk_scene *rtk_cuda_build_instanced_scene(const rtk_cuda_mesh *meshes, size_t num_meshes,
                                          const rtk_cuda_instance *instances, size_t num_instances, void *stream);
int rtk_cuda_update_instanced_scene(const rtk_scene *scene, const rtk_cuda_mesh *meshes, size_t num_meshes,
                                    const rtk_cuda_instance *instances, size_t num_instances, int mode, void *stream);

int rtk_cuda_get_scene_info(const rtk_scene *scene, rtk_cuda_scene_info *info);

/* Upload a relocated / reloaded blob (written by rtk_finish_build_to) so that
 * later queries on this address hit the device directly.  Queries do this
 * lazily themselves; the explicit call lets the caller pay the cost up front. */
int rtk_cuda_attach_scene(const rtk_scene *scene);
int rtk_cuda_detach_scene(const rtk_scene *scene);

#ifdef __cplusplus
}
#endif

#endif /* RTK_B200_RTK_CUDA_H */
