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
#ifndef RTK_SIMT_EMU
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>
#endif

#define RTKD_OK 0
#define RTKD_ERR_NO_DEVICE (-1)
#define RTKD_ERR_CUDA (-2)
#define RTKD_ERR_ARGUMENT (-3)
#define RTKD_ERR_SCENE (-4)
#define RTKD_ERR_MEMORY (-5)
#define RTKD_ERR_OVERFLOW (-6)

static __thread char g_err[512];

static int g_sm_count = 0, g_trace_ctas = 0, g_trace_lanes = RTK_TRACE_LANES, g_trace_pd = 1;
static int g_reserved_sms = 0;      // SMs the persistent traversal grid leaves free (for concurrent NCCL kernels)
static size_t g_l2_bytes = 0;
static size_t g_l2_window_max = 0;  // largest access-policy window the device accepts (0: no L2 persistence)
static size_t g_l2_setaside = 0;    // L2 bytes set aside for persisting lines
static int g_sah_sort_bits = 24;    // RTK_B200_SAH_SORT_BITS (8, 16, 24 or 32): Morton bits the SAH builder's input order is sorted by (8 per axis:
                                    // the order only buys locality, the tree does not depend on it; 32 -> 24 bits: 1.69 -> 1.65 ms at 1M triangles, 12.4 -> 12.1 at 10M)
static int g_l2_persist = 0;        // RTK_B200_L2_PERSIST=1: persisting L2 window over nodes + leaf slots.  Off by default --
                                    // measured on C3/C4: k_trace gains nothing (1705 vs 1704 Mrays/s), while k_resolve, whose
                                    // corner gathers lose the set-aside part of the L2, goes from 0.47 to 1.10 ms per batch
static int g_host_mix = -1;         // RTK_B200_HOST_MIX: every k-th chunk of a direct batch takes the staged route (-1: 2 on one device, none on several)
static int g_push_sms = 8;          // RTK_B200_PUSH_SMS: SMs the traversal grid leaves to the row-push kernel of the direct host path
static __thread int t_reserve_extra = 0;   // set by the direct host pipeline around its traversal launches
static int g_host_direct = 1;       // RTK_B200_HOST_DIRECT=0: rows always travel through pinned staging
static size_t g_min_share = (size_t)1 << 18;   // a device joins a host batch only for at least this many rays (RTK_B200_HOST_MIN_SHARE_LOG2)
static int g_stack_limit = 0;       // test hook (rtkd_debug_limit_stack): spill entries per ray, 0 = sized from the tree
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

// ---------------------------------------------------------------------------------------------
// devices.  The library drives a LIST of devices from one process (rtk_cuda_init_devices): scenes are
// built on the first and replicated to the others, host batches are split over all of them, device
// entry points run on whichever of them owns the caller's buffers.  rtk_cuda_init(d) is the list {d}.
// Everything a host batch needs on one device (streams, staging, its worker thread) is in dev_ctx.
// ---------------------------------------------------------------------------------------------

#define RTKD_HOST_BUFS 4
#define RTKD_HOST_RING 8             // upload events: more than RTKD_HOST_AHEAD + 1

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
	// pageable rays: chunks are copied into pinned bounce buffers by the host worker pool first (cudaMemcpyAsync from
	// pageable memory stages through ONE driver thread at ~10 GB/s and blocks the caller meanwhile)
	unsigned char *h_up[RTKD_HOST_RING];
	cudaEvent_t up_read[RTKD_HOST_RING];
	bool up_used[RTKD_HOST_RING];
	cudaStream_t up;                    // upload stream
	cudaStream_t push;                  // direct host path: the stream of the row-push kernel
	cudaEvent_t uploaded[RTKD_HOST_RING];
	float4 *d_rays; size_t rays_cap;    // the rays of this device's share of the batch (grow-only)
	unsigned long long *d_count;        // hit counter of a batch whose rows go straight to the caller's memory
	bool ready;                         // device buffers allocated for `chunk`
	bool staged;                        // ... and the pinned staging of the placement path as well
	bool streams;                       // streams and events exist (they outlive a change of chunk size)
	// the small-batch path (rtk_trace_ray) has its own few kilobytes
	cudaStream_t sm_st;
	float4 *sm_d_rays, *sm_d_h16; uint32_t *sm_d_rows; unsigned char *sm_d_mask;
	unsigned char *sm_h_rows, *sm_h_mask;
};

struct batch_job;
struct dev_ctx {
	int device;                         // CUDA ordinal
	pthread_mutex_t lock;               // one host batch at a time per device
	host_stage stage;
	// worker thread: runs this device's share of a multi-device batch (the caller's thread runs the first share)
	pthread_t thread; bool thread_up, quit;
	pthread_mutex_t jm; pthread_cond_t jc;
	batch_job *job; bool job_done;
};
static dev_ctx g_ctx[RTKD_MAX_DEVICES];
static int g_ndev = 0;
static pthread_mutex_t g_init_lock = PTHREAD_MUTEX_INITIALIZER;

static int bind_index(int k)
{
	int cur = -1;
	if (cudaGetDevice(&cur) != cudaSuccess || cur != g_ctx[k].device) CK(cudaSetDevice(g_ctx[k].device));
	return RTKD_OK;
}

static void *ctx_worker(void *arg);
static void stage_shutdown(dev_ctx &X);          // host-batch staging (defined with the pipeline below)

static int init_devices_locked(const int *devices, int n)
{
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count <= 0) {
		rtkd_set_error("rtk_b200: no usable CUDA device (%s); this library has no CPU fallback",
		               e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
		return RTKD_ERR_NO_DEVICE;
	}
	if (n < 1 || n > RTKD_MAX_DEVICES || !devices) { rtkd_set_error("device list must name 1..%d devices", RTKD_MAX_DEVICES); return RTKD_ERR_ARGUMENT; }
	for (int i = 0; i < n; i++) {
		if (devices[i] < 0 || devices[i] >= count) { rtkd_set_error("device %d out of range (0..%d)", devices[i], count - 1); return RTKD_ERR_ARGUMENT; }
		// (test hook: RTK_B200_TEST_DUP_DEVICES=1 lets one physical device appear several times, so that the
		// whole multi-device layer -- replicas, worker threads, range split -- runs on a one-GPU box)
		const char *dup = getenv("RTK_B200_TEST_DUP_DEVICES");
		for (int j = 0; j < i; j++) if (devices[j] == devices[i] && !(dup && atoi(dup))) { rtkd_set_error("device %d named twice", devices[i]); return RTKD_ERR_ARGUMENT; }
	}
	if (g_ndev) {
		// already up: the same list is a no-op (another host thread binding itself), a different one is refused
		bool same = n == g_ndev;
		for (int i = 0; same && i < n; i++) same = g_ctx[i].device == devices[i];
		if (same) return bind_index(0);
		rtkd_set_error("rtk_b200 is already bound to %d device(s) starting at device %d: call rtk_cuda_shutdown() first", g_ndev, g_ctx[0].device);
		return RTKD_ERR_ARGUMENT;
	}
	int sm = 0;
	size_t l2 = 0;
	for (int i = 0; i < n; i++) {
		cudaDeviceProp prop;
		CK(cudaSetDevice(devices[i]));
		CK(cudaGetDeviceProperties(&prop, devices[i]));
		if (i == 0) { sm = prop.multiProcessorCount; l2 = (size_t)prop.l2CacheSize; }
		else if (prop.multiProcessorCount != sm) { rtkd_set_error("devices %d and %d differ (%d vs %d SMs): the device list must be homogeneous", devices[0], devices[i], sm, prop.multiProcessorCount); return RTKD_ERR_ARGUMENT; }
#ifndef RTK_SIMT_EMU
		{
			// keep freed scratch in the stream-ordered pool: the build allocates its temporaries
			// with cudaMallocAsync on every (re)build
			cudaMemPool_t pool;
			if (cudaDeviceGetDefaultMemPool(&pool, devices[i]) == cudaSuccess) {
				unsigned long long thr = ~0ull;
				cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
			}
		}
#endif
		// L2 persistence: the traversal working set (nodes + leaf slots) gets a persisting access-policy
		// window so that the ray / hit streams do not evict it
		int maxp = 0, maxw = 0;
		{ const char *e = getenv("RTK_B200_L2_PERSIST"); if (e) g_l2_persist = atoi(e) != 0; }
		if (g_l2_persist && cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, devices[i]) == cudaSuccess && maxp > 0) {
			cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp);
			cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, devices[i]);
			if (i == 0) { g_l2_window_max = (size_t)(maxw > 0 ? maxw : 0); g_l2_setaside = (size_t)maxp; }
		}
		cudaGetLastError();
	}
	// peers: replicas are copied device to device (NVLink when the devices are peers)
	for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) if (i != j) {
		int can = 0;
		if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
			cudaSetDevice(devices[i]);
			cudaDeviceEnablePeerAccess(devices[j], 0);
			cudaGetLastError();                         // "already enabled" is fine
		}
	}
	CK(cudaSetDevice(devices[0]));
	g_sm_count = sm; g_l2_bytes = l2;
	{
		const char *e = getenv("RTK_B200_LANES");          // experiment knob: lanes per ray
		if (e && (atoi(e) == 8 || atoi(e) == 4 || atoi(e) == 2)) g_trace_lanes = atoi(e);
	}
	{
		const char *e = getenv("RTK_B200_PD");             // experiment knob: 0 = every ray's own lanes walk its leaf
		if (e) g_trace_pd = atoi(e) != 0;
		if (g_trace_lanes == 8) g_trace_pd = 0;
	}
	{ const char *e = getenv("RTK_B200_PUSH_SMS"); if (e && atoi(e) >= 0 && atoi(e) < sm) g_push_sms = atoi(e); }
	{ const char *e = getenv("RTK_B200_SAH_SORT_BITS"); if (e && atoi(e) >= 8 && atoi(e) <= 32 && atoi(e) % 8 == 0) g_sah_sort_bits = atoi(e); }
	{ const char *e = getenv("RTK_B200_HOST_MIX"); if (e && atoi(e) >= 0) g_host_mix = atoi(e); }
	{ const char *e = getenv("RTK_B200_HOST_DIRECT"); if (e) g_host_direct = atoi(e) != 0; }
	{ const char *e = getenv("RTK_B200_HOST_MIN_SHARE_LOG2"); if (e && atoi(e) >= 7 && atoi(e) <= 30) g_min_share = (size_t)1 << atoi(e); }
	int ctas = 0;
	if (g_trace_lanes == 8) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<8, 1, false>, RTK_TRACE_THREADS, 0));
	else if (g_trace_lanes == 4 && g_trace_pd) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<4, 1, false, false, true>, RTK_TRACE_THREADS, 0));
	else if (g_trace_lanes == 4) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<4, 1, false>, RTK_TRACE_THREADS, 0));
	else if (g_trace_pd) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<2, 1, false, false, true>, RTK_TRACE_THREADS, 0));
	else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, k_trace<2, 1, false>, RTK_TRACE_THREADS, 0));
	g_trace_ctas = ctas > 0 ? ctas : 1;
	for (int i = 0; i < n; i++) {
		dev_ctx &X = g_ctx[i];
		memset(&X, 0, sizeof(X));
		X.device = devices[i];
		pthread_mutex_init(&X.lock, NULL);
		pthread_mutex_init(&X.jm, NULL);
		pthread_cond_init(&X.jc, NULL);
	}

	__atomic_store_n(&g_ndev, n, __ATOMIC_RELEASE);
	// one worker per further device; they sleep on a condition variable between batches
	for (int i = 1; i < n; i++) {
		// a device without a worker still takes part: the caller's thread runs its share after its own
		if (pthread_create(&g_ctx[i].thread, NULL, ctx_worker, &g_ctx[i]) == 0) g_ctx[i].thread_up = true;
	}
	return RTKD_OK;
}

extern "C" int rtkd_init_devices(const int *devices, int n)
{
	pthread_mutex_lock(&g_init_lock);
	int r = init_devices_locked(devices, n);
	pthread_mutex_unlock(&g_init_lock);
	return r;
}

extern "C" int rtkd_init(int device) { return rtkd_init_devices(&device, 1); }
extern "C" int rtkd_device_count(void) { return g_ndev; }

// Every entry point passes through here (directly or via rtkd_bind_thread): CUDA's current device is
// per host thread and starts at 0 -- a worker thread of the caller that traces against a scene on
// device 3 must be switched to device 3 first (the reference's rtk_trace_ray is called from many user
// threads, rtk.h:129).
static int ensure_init(void)
{
	if (!__atomic_load_n(&g_ndev, __ATOMIC_ACQUIRE)) { int d = 0; return rtkd_init_devices(&d, 1); }
	return bind_index(0);
}
extern "C" int rtkd_bind_thread(void) { return ensure_init(); }

extern "C" void rtkd_shutdown(void)
{
	// scenes are the caller's to free first; what the library itself holds on the devices goes here
	pthread_mutex_lock(&g_init_lock);
	const int n = g_ndev;
	for (int i = 1; i < n; i++) {
		dev_ctx &X = g_ctx[i];
		if (!X.thread_up) continue;
		pthread_mutex_lock(&X.jm); X.quit = true; pthread_cond_broadcast(&X.jc); pthread_mutex_unlock(&X.jm);
		pthread_join(X.thread, NULL);
		X.thread_up = false;
	}
	for (int i = 0; i < n; i++) {
		dev_ctx &X = g_ctx[i];
		if (cudaSetDevice(X.device) == cudaSuccess) { cudaDeviceSynchronize(); stage_shutdown(X); }
		pthread_mutex_destroy(&X.lock); pthread_mutex_destroy(&X.jm); pthread_cond_destroy(&X.jc);
	}
	if (n) cudaSetDevice(g_ctx[0].device);
	g_sm_count = 0;
	__atomic_store_n(&g_ndev, 0, __ATOMIC_RELEASE);
	pthread_mutex_unlock(&g_init_lock);
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

extern "C" int rtkd_debug_limit_stack(int entries) { g_stack_limit = entries > 0 ? entries : 0; return RTKD_OK; }

// page-locked host memory that every device of the list can read and write in place
extern "C" void *rtkd_host_alloc(size_t bytes)
{
	if (ensure_init()) return NULL;
	void *p = NULL;
	CKP(cudaHostAlloc(&p, bytes ? bytes : 16, cudaHostAllocPortable | cudaHostAllocMapped));
	return p;
}
extern "C" void rtkd_host_free_any(void *p);
extern "C" void rtkd_host_free(void *p) { rtkd_host_free_any(p); }
extern "C" int rtkd_host_register(void *p, size_t bytes)
{
	if (ensure_init()) return RTKD_ERR_NO_DEVICE;
	if (!p || !bytes) { rtkd_set_error("nothing to register"); return RTKD_ERR_ARGUMENT; }
	CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
	return RTKD_OK;
}
extern "C" int rtkd_host_unregister(void *p)
{
	if (!p) return RTKD_OK;
	CK(cudaHostUnregister(p));
	return RTKD_OK;
}

// ---------------------------------------------------------------------------------------------
// Host arrays for batches that are split over several devices.  cudaHostAlloc puts every page on the NUMA
// node of the calling thread; on a two-socket box all eight links then read and write ONE socket's memory
// (measured on the 8-GPU box: 8 devices reach 96 GB/s device-to-host against 56 GB/s for one).  A batch
// array is therefore allocated share by share: the part of the array that device k will read or write
// (the very split run_batch makes) is first touched by a thread pinned to the CPUs of that device's NUMA
// node, then the whole array is page-locked.  Linux only; anything that cannot be found out (no NUMA
// information, a container without the sysfs files) degrades to a plain page-locked allocation.
// ---------------------------------------------------------------------------------------------

struct batch_alloc { void *p; size_t bytes; bool mapped; };
static batch_alloc g_ballocs[64];
static pthread_mutex_t g_balloc_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_numa_policy = 1;       // RTK_B200_NUMA: 0 plain allocations, 1 device-local shares (default), 2 round-robin over the nodes

static void batch_split(size_t n, int *use_out, size_t *per_out)
{
	int use = g_ndev > 0 ? g_ndev : 1;
	while (use > 1 && n / (size_t)use < g_min_share) use--;
	*use_out = use;
	*per_out = (n / (size_t)use + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS * RTK_RESOLVE_THREADS;
}

#ifndef RTK_SIMT_EMU
static int numa_node_count(void)
{
	int n = 0;
	char path[96];
	for (; n < 64; n++) { snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", n); if (access(path, R_OK) != 0) break; }
	return n;
}
static bool numa_node_cpus(int node, cpu_set_t *set)
{
	char path[96], buf[4096];
	snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
	FILE *f = fopen(path, "r");
	if (!f) return false;
	const bool ok = fgets(buf, sizeof(buf), f) != NULL;
	fclose(f);
	if (!ok) return false;
	CPU_ZERO(set);
	int any = 0;
	for (char *q = buf; *q; ) {                               // "0-15,32-47"
		char *e; long a = strtol(q, &e, 10); if (e == q) break;
		long b = a; if (*e == '-') { q = e + 1; b = strtol(q, &e, 10); }
		for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int)c, set); any = 1; }
		q = *e == ',' ? e + 1 : e; if (*e != ',' ) break;
	}
	return any != 0;
}
static int numa_node_of_device(int ordinal)
{
	char bus[32] = { 0 }, path[128];
	if (cudaDeviceGetPCIBusId(bus, sizeof(bus), ordinal) != cudaSuccess) { cudaGetLastError(); return -1; }
	for (char *q = bus; *q; q++) if (*q >= 'A' && *q <= 'Z') *q = (char)(*q - 'A' + 'a');
	snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
	FILE *f = fopen(path, "r");
	if (!f) return -1;
	int node = -1;
	if (fscanf(f, "%d", &node) != 1) node = -1;
	fclose(f);
	return node;
}
struct touch_job { char *p; size_t bytes; int node; };
static void *touch_worker(void *arg)
{
	touch_job *j = (touch_job*)arg;
	cpu_set_t set;
	if (j->node >= 0 && numa_node_cpus(j->node, &set)) {
		cpu_set_t allowed;                                    // stay inside the process's own cpuset
		if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) { CPU_AND(&set, &set, &allowed); }
		if (CPU_COUNT(&set) > 0) sched_setaffinity(0, sizeof(set), &set);
	}
	for (size_t off = 0; off < j->bytes; off += 4096) j->p[off] = 0;           // first touch: the page lands on this thread's node
	return NULL;
}
#endif

// page-locked memory of share_bytes[0] + ... + share_bytes[use-1] bytes, share k placed on device k's NUMA node
static void *numa_alloc_shares(const size_t *share_bytes, int use)
{
	size_t bytes = 0;
	for (int k = 0; k < use; k++) bytes += share_bytes[k];
	if (!bytes) bytes = 16;
#ifndef RTK_SIMT_EMU
	{ const char *e = getenv("RTK_B200_NUMA"); if (e) g_numa_policy = atoi(e); }
	const int nodes = numa_node_count();
	if (g_numa_policy > 0 && nodes > 1 && use > 1) {
		const size_t total = (bytes + 4095) & ~(size_t)4095;
		char *p = (char*)mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
		if (p != (char*)MAP_FAILED) {
			touch_job jobs[RTKD_MAX_DEVICES];
			pthread_t th[RTKD_MAX_DEVICES];
			bool up[RTKD_MAX_DEVICES];
			size_t begin = 0, acc = 0;
			for (int k = 0; k < use; k++) {
				acc += share_bytes[k];
				size_t end = k == use - 1 ? total : acc & ~(size_t)4095;
				if (end > total) end = total;
				int node = g_numa_policy == 1 ? numa_node_of_device(g_ctx[k].device) : -1;
				if (node < 0 || node >= nodes) node = (int)((size_t)k * (size_t)nodes / (size_t)use);
				jobs[k].p = p + begin; jobs[k].bytes = end > begin ? end - begin : 0; jobs[k].node = node;
				up[k] = jobs[k].bytes && pthread_create(&th[k], NULL, touch_worker, &jobs[k]) == 0;
				if (!up[k] && jobs[k].bytes) touch_worker(&jobs[k]);
				begin = end;
			}
			for (int k = 0; k < use; k++) if (up[k]) pthread_join(th[k], NULL);
			if (cudaHostRegister(p, total, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess) {
				pthread_mutex_lock(&g_balloc_lock);
				for (int i = 0; i < 64; i++) if (!g_ballocs[i].p) { g_ballocs[i].p = p; g_ballocs[i].bytes = total; g_ballocs[i].mapped = true; break; }
				pthread_mutex_unlock(&g_balloc_lock);
				return p;
			}
			cudaGetLastError();
			munmap(p, total);
		}
	}
#endif
	return rtkd_host_alloc(bytes);
}

// page-locked array of `count` elements of `elem_bytes` for batches of `count` rays: the elements device k
// will read or write (run_batch's split) are placed on device k's NUMA node
extern "C" void *rtkd_host_alloc_batch(size_t elem_bytes, size_t count)
{
	if (ensure_init()) return NULL;
	int use = 1; size_t per = count;
	batch_split(count, &use, &per);
	size_t shares[RTKD_MAX_DEVICES];
	size_t left = count;
	for (int k = 0; k < use; k++) { size_t c = k == use - 1 ? left : (left < per ? left : per); shares[k] = c * elem_bytes; left -= c; }
	return numa_alloc_shares(shares, use);
}

extern "C" void rtkd_host_free_any(void *p)
{
	if (!p) return;
#ifndef RTK_SIMT_EMU
	pthread_mutex_lock(&g_balloc_lock);
	for (int i = 0; i < 64; i++) if (g_ballocs[i].p == p) {
		const size_t bytes = g_ballocs[i].bytes;
		g_ballocs[i].p = NULL;
		pthread_mutex_unlock(&g_balloc_lock);
		cudaHostUnregister(p);
		munmap(p, bytes);
		return;
	}
	pthread_mutex_unlock(&g_balloc_lock);
#endif
	cudaFreeHost(p);
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

// Random-gather probe: the access pattern of the traversal, not of a copy.  Every warp reads
// `record_bytes`-sized records (256 = a wide node: 8 lanes x LDG.256; 128 = one leaf-slot line: 8 lanes
// x LDG.128) at hashed offsets of a `bytes`-sized buffer, four independent records in flight per lane
// group.  With the buffer inside the L2 this is the L2 bandwidth the traversal can hope for, which is
// what roofline.l2 divides by; a streaming probe flatters that denominator.
__global__ void __launch_bounds__(256) k_gather_probe(const float4 *buf, uint32_t nrec, uint32_t rec16, int iters, float *sink)
{
	float acc = 0.0f;
	const uint32_t lane = threadIdx.x & 31, sub = lane & 7, grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
	uint32_t x = grp * 0x9E3779B9u + 0x7F4A7C15u;
	for (int it = 0; it < iters; it++) {
		float4 a[4];
#pragma unroll
		for (int j = 0; j < 4; j++) {
			x ^= x << 13; x ^= x >> 17; x ^= x << 5;                       // xorshift32, one record per 8 lanes
			const float4 *r = buf + (size_t)(x % nrec) * rec16;
			a[j] = __ldg(r + sub);
			if (rec16 > 8) { float4 b = __ldg(r + 8 + sub); a[j].x += b.y; }
		}
		acc += (a[0].x + a[1].y) + (a[2].z + a[3].w);
	}
	if (acc == 123.456f) *sink = acc;
}

extern "C" int rtkd_gather_bandwidth(size_t bytes, size_t record_bytes, int passes, double *gbs)
{
	if (ensure_init()) return RTKD_ERR_NO_DEVICE;
	if (bytes < 65536 || (record_bytes != 128 && record_bytes != 256) || passes < 1 || !gbs) { rtkd_set_error("bad probe arguments"); return RTKD_ERR_ARGUMENT; }
	float4 *buf = NULL;
	float *sink = NULL;
	struct dev_free { void *p = NULL; ~dev_free() { if (p) cudaFree(p); } } buf_guard, sink_guard;
	CK(cudaMalloc(&buf, bytes));
	buf_guard.p = buf;
	CK(cudaMalloc(&sink, 4));
	sink_guard.p = sink;
	CK(cudaMemset(buf, 0, bytes));
	event_pair ev;
	CK(cudaEventCreate(&ev.e0)); CK(cudaEventCreate(&ev.e1));
	const uint32_t nrec = (uint32_t)(bytes / record_bytes), rec16 = (uint32_t)(record_bytes / 16);
	const unsigned grid = (unsigned)g_sm_count * 8;
	const int iters = 64;
	RTK_LAUNCH(k_gather_probe, grid, 256, 0, (const float4*)buf, nrec, rec16, iters, sink);      // warm the cache
	CK(cudaEventRecord(ev.e0, 0));
	for (int p = 0; p < passes; p++) RTK_LAUNCH(k_gather_probe, grid, 256, 0, (const float4*)buf, nrec, rec16, iters, sink);
	CK(cudaEventRecord(ev.e1, 0));
	CK(cudaEventSynchronize(ev.e1));
	float ms = 0.0f;
	CK(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
	const double total = (double)grid * 256 / 8 * iters * 4 * (double)record_bytes * passes;
	*gbs = ms > 0.0f ? total / (ms * 1e-3) / 1e9 : 0.0;
	return RTKD_OK;
}

// Host-link probe: the ceiling of the host-buffer path.  On each of the first `ndev` devices of the list
// one copy stream moves `bytes_per_device` from pinned host memory up (dir & 1) and another one the
// same amount down (dir & 2), all devices at once, `passes` times; the result is the aggregate GB/s of
// the directions asked for (PCIe links, root complexes and host DRAM all included).
extern "C" int rtkd_link_bandwidth(int ndev, size_t bytes_per_device, int dir, int passes, double *gbs)
{
	if (ensure_init()) return RTKD_ERR_NO_DEVICE;
	if (ndev < 1 || ndev > g_ndev || !bytes_per_device || !(dir & 3) || passes < 1 || !gbs) { rtkd_set_error("bad link probe arguments"); return RTKD_ERR_ARGUMENT; }
	struct leg { void *h_up = NULL, *h_dn = NULL, *d_up = NULL, *d_dn = NULL; cudaStream_t su = NULL, sd = NULL; };
	leg L[RTKD_MAX_DEVICES];
	size_t shares[RTKD_MAX_DEVICES];
	for (int i = 0; i < ndev; i++) shares[i] = bytes_per_device;
	char *h_up_all = (dir & 1) ? (char*)numa_alloc_shares(shares, ndev) : NULL;
	char *h_dn_all = (dir & 2) ? (char*)numa_alloc_shares(shares, ndev) : NULL;
	int rc = RTKD_OK;
	cudaError_t e = cudaSuccess;
	for (int i = 0; i < ndev && e == cudaSuccess; i++) {
		e = cudaSetDevice(g_ctx[i].device);
		// the host side of every leg is ONE array of ndev shares, placed like a batch array (share i near device i)
		if (e == cudaSuccess && (dir & 1)) { L[i].h_up = h_up_all ? h_up_all + (size_t)i * bytes_per_device : NULL; if (!L[i].h_up) e = cudaErrorMemoryAllocation; if (e == cudaSuccess) e = cudaMalloc(&L[i].d_up, bytes_per_device); }
		if (e == cudaSuccess && (dir & 2)) { L[i].h_dn = h_dn_all ? h_dn_all + (size_t)i * bytes_per_device : NULL; if (!L[i].h_dn) e = cudaErrorMemoryAllocation; if (e == cudaSuccess) e = cudaMalloc(&L[i].d_dn, bytes_per_device); }
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&L[i].su, cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&L[i].sd, cudaStreamNonBlocking);
		if (e == cudaSuccess && L[i].h_up) memset(L[i].h_up, 1, bytes_per_device);      // touch: the pages exist before the clock starts
		if (e == cudaSuccess && L[i].h_dn) memset(L[i].h_dn, 1, bytes_per_device);
	}
	double best = 0.0;
	for (int rep = 0; rep < 2 && e == cudaSuccess; rep++) {          // first repetition warms up
		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		for (int p = 0; p < passes && e == cudaSuccess; p++)
			for (int i = 0; i < ndev && e == cudaSuccess; i++) {
				e = cudaSetDevice(g_ctx[i].device);
				if (e == cudaSuccess && (dir & 1)) e = cudaMemcpyAsync(L[i].d_up, L[i].h_up, bytes_per_device, cudaMemcpyHostToDevice, L[i].su);
				if (e == cudaSuccess && (dir & 2)) e = cudaMemcpyAsync(L[i].h_dn, L[i].d_dn, bytes_per_device, cudaMemcpyDeviceToHost, L[i].sd);
			}
		for (int i = 0; i < ndev && e == cudaSuccess; i++) {
			e = cudaSetDevice(g_ctx[i].device);
			if (e == cudaSuccess) e = cudaStreamSynchronize(L[i].su);
			if (e == cudaSuccess) e = cudaStreamSynchronize(L[i].sd);
		}
		clock_gettime(CLOCK_MONOTONIC, &t1);
		const double sec = (t1.tv_sec - t0.tv_sec) + (t1.tv_nsec - t0.tv_nsec) * 1e-9;
		const double total = (double)bytes_per_device * passes * ndev * (((dir & 1) ? 1 : 0) + ((dir & 2) ? 1 : 0));
		if (sec > 0.0) best = total / sec / 1e9;
	}
	if (e != cudaSuccess) { rtkd_set_error("link probe: %s", cudaGetErrorString(e)); rc = RTKD_ERR_CUDA; }
	for (int i = 0; i < ndev; i++) {
		cudaSetDevice(g_ctx[i].device);
		if (L[i].su) cudaStreamDestroy(L[i].su);
		if (L[i].sd) cudaStreamDestroy(L[i].sd);
		cudaFree(L[i].d_up); cudaFree(L[i].d_dn);
	}
	rtkd_host_free(h_up_all); rtkd_host_free(h_dn_all);
	bind_index(0);
	*gbs = best;
	return rc;
}

// ---------------------------------------------------------------------------------------------
// scene lifetime
// ---------------------------------------------------------------------------------------------

static rtkd_scene *scene_alloc_on(int dev_index, uint32_t num_tris, uint32_t num_meshes, const uint32_t *mesh_first)
{
	if (bind_index(dev_index)) return NULL;
	rtkd_scene *s = (rtkd_scene*)calloc(1, sizeof(rtkd_scene));
	if (!s) { rtkd_set_error("out of host memory"); return NULL; }
	s->dev_index = dev_index;
	s->num_tris = num_tris; s->num_meshes = num_meshes;
	s->h_mesh_first = (uint32_t*)malloc(sizeof(uint32_t) * (num_meshes + 1));
	pthread_mutex_t *m = (pthread_mutex_t*)malloc(sizeof(pthread_mutex_t));
	if (!s->h_mesh_first || !m) { free(s->h_mesh_first); free(m); free(s); rtkd_set_error("out of host memory"); return NULL; }
	pthread_mutex_init(m, NULL);
	s->slot_lock = m;
	memcpy(s->h_mesh_first, mesh_first, sizeof(uint32_t) * (num_meshes + 1));
	cudaError_t e = cudaMalloc((float4**)&s->tri_orig, sizeof(float4) * 3 * (size_t)(num_tris ? num_tris : 1));
	if (e == cudaSuccess) e = cudaMalloc((uint32_t**)&s->mesh_first, sizeof(uint32_t) * (num_meshes + 1));
	if (e == cudaSuccess) e = cudaMemcpy(s->mesh_first, mesh_first, sizeof(uint32_t) * (num_meshes + 1), cudaMemcpyHostToDevice);
	// the status word lives in pinned host memory the device writes to directly: the host reads it
	// without a copy once it has synchronised with the query
	if (e == cudaSuccess) e = cudaHostAlloc((uint32_t**)&s->h_status, 64, cudaHostAllocPortable | cudaHostAllocMapped);
	if (e == cudaSuccess) {
		*s->h_status = 0;
		cudaPointerAttributes a;
		if (cudaPointerGetAttributes(&a, s->h_status) == cudaSuccess && a.devicePointer) s->d_status = (uint32_t*)a.devicePointer;
		else s->d_status = s->h_status;
	}
	if (e != cudaSuccess) {
		rtkd_set_error("scene allocation failed: %s", cudaGetErrorString(e));
		rtkd_scene_free(s);
		return NULL;
	}
	return s;
}

extern "C" rtkd_scene *rtkd_scene_new(uint32_t num_tris, uint32_t num_meshes, const uint32_t *mesh_first)
{
	if (ensure_init()) return NULL;
	rtkd_scene *s = scene_alloc_on(0, num_tris, num_meshes, mesh_first);
	if (s) s->id = ((uint64_t)time(NULL) << 20) ^ (__atomic_fetch_add(&g_next_id, 1, __ATOMIC_RELAXED) * 0x9E3779B97F4A7C15ull);   // scenes may be built from several host threads
	return s;
}

static void scene_free_slots(rtkd_scene *s)
{
	for (int k = 0; k < RTKD_TRACE_SLOTS; k++) {
		rtkd_trace_slot &T = s->slot[k];
		if (T.scratch) cudaFree(T.scratch);
		if (T.overflow) cudaFree(T.overflow);
		if (T.done) cudaEventDestroy((cudaEvent_t)T.done);
		memset(&T, 0, sizeof(T));
	}
}

extern "C" void rtkd_scene_free(rtkd_scene *s)
{
	if (!s) return;
	for (int k = 1; k < RTKD_MAX_DEVICES; k++) if (s->replica[k]) { rtkd_scene_free(s->replica[k]); s->replica[k] = NULL; }
	if (g_ndev && s->dev_index < g_ndev) bind_index(s->dev_index);      // the freeing thread may not be the one that built the scene
	cudaDeviceSynchronize();
	if (s->tri_orig) cudaFree(s->tri_orig);
	if (s->arena) cudaFree(s->arena);
	if (s->node_level) cudaFree(s->node_level);
	if (s->mesh_first) cudaFree(s->mesh_first);
	scene_free_slots(s);
	if (s->h_status) cudaFreeHost(s->h_status);
	if (s->hit16) cudaFree(s->hit16);
	if (s->filter_bits) cudaFree(s->filter_bits);
	if (s->slot_lock) { pthread_mutex_destroy((pthread_mutex_t*)s->slot_lock); free(s->slot_lock); }
	free(s->h_mesh_first);
	free(s);
	if (g_ndev) bind_index(0);
}

// nodes | tv0 | tv1 | tv2 in one allocation (256-byte aligned sections); grows with some slack so that
// rebuilds of a deforming mesh rarely need a new allocation
static size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }
static int scene_arena_layout(rtkd_scene *s, uint32_t num_nodes, uint32_t num_tv, bool slack)
{
	const size_t nb = a256(256 * (size_t)(num_nodes ? num_nodes : 1)), tb = a256(16 * (size_t)(num_tv ? num_tv : 8));
	const size_t need = nb + 3 * tb;
	if (!s->arena || s->arena_cap < need) {
		if (s->arena) cudaFree(s->arena);
		s->arena = NULL; s->arena_cap = 0;
		const size_t cap = slack ? need + need / 16 + 4096 : need;
		CK(cudaMalloc((unsigned char**)&s->arena, cap));
		s->arena_cap = cap;
	}
	unsigned char *b = (unsigned char*)s->arena;
	s->nodes = b; s->tv0 = b + nb; s->tv1 = b + nb + tb; s->tv2 = b + nb + 2 * tb;
	s->arena_used = need;
	return RTKD_OK;
}

// Replicas: the scene as built on the first device, copied array by array to every further device of
// the list (device-to-device; NVLink between peers).  A 10M-triangle scene is 1.1 GB, i.e. a couple of
// milliseconds per replica, all replicas in flight at once -- cheaper than building it again on every
// device (15 ms) or pushing the meshes through PCIe once more.
extern "C" int rtkd_sync_replicas(rtkd_scene *s)
{
	if (g_ndev <= 1 || s->dev_index != 0) return RTKD_OK;
	pthread_mutex_t *m = (pthread_mutex_t*)s->slot_lock;
	RTK_NVTX("rtk_b200 scene replication");
	pthread_mutex_lock(m);
	if (s->replica_epoch == s->epoch && s->replica[1]) { pthread_mutex_unlock(m); return RTKD_OK; }
	int rc = RTKD_OK;
	const int src = g_ctx[0].device;
	cudaStream_t st[RTKD_MAX_DEVICES] = { NULL };
	for (int k = 1; k < g_ndev && rc == RTKD_OK; k++) {
		rtkd_scene *r = s->replica[k];
		if (r && (r->num_tris != s->num_tris || r->num_meshes != s->num_meshes)) { rtkd_scene_free(r); r = s->replica[k] = NULL; }
		if (!r) {
			r = scene_alloc_on(k, s->num_tris, s->num_meshes, s->h_mesh_first);
			if (!r) { rc = RTKD_ERR_MEMORY; break; }
			s->replica[k] = r;
		}
		if (bind_index(k)) { rc = RTKD_ERR_CUDA; break; }
		r->id = s->id;
		r->num_nodes = s->num_nodes; r->num_leaves = s->num_leaves; r->depth = s->depth; r->build_mode = s->build_mode;
		r->num_tv = s->num_tv;
		memcpy(r->bounds_min, s->bounds_min, 12); memcpy(r->bounds_max, s->bounds_max, 12);
		r->abs_max = s->abs_max; r->sah_cost = s->sah_cost;
		r->build_device_ms = s->build_device_ms; r->build_total_ms = s->build_total_ms;
		*r->h_status = 0;
		rc = scene_arena_layout(r, s->num_nodes, s->num_tv, false);
		if (rc) break;
		const int dst = g_ctx[k].device;
		cudaError_t e = cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking);
		if (e == cudaSuccess && s->num_tris) e = cudaMemcpyPeerAsync(r->tri_orig, dst, s->tri_orig, src, 48 * (size_t)s->num_tris, st[k]);
		if (e == cudaSuccess && s->arena_used && s->arena) e = cudaMemcpyPeerAsync(r->arena, dst, s->arena, src, s->arena_used, st[k]);
		if (e != cudaSuccess) { rtkd_set_error("scene replication to device %d failed: %s", dst, cudaGetErrorString(e)); rc = RTKD_ERR_CUDA; }
	}
	for (int k = 1; k < g_ndev; k++) if (st[k]) {
		bind_index(k);
		if (cudaStreamSynchronize(st[k]) != cudaSuccess && rc == RTKD_OK) { rtkd_set_error("scene replication failed"); rc = RTKD_ERR_CUDA; }
		cudaStreamDestroy(st[k]);
	}
	bind_index(0);
	if (rc == RTKD_OK) s->replica_epoch = s->epoch;
	pthread_mutex_unlock(m);
	return rc;
}

extern "C" rtkd_scene *rtkd_scene_for_pointer(rtkd_scene *s, const void *p)
{
	if (g_ndev <= 1) { if (bind_index(0)) return NULL; return s; }
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); rtkd_set_error("cannot tell which device owns the buffer"); return NULL; }
	if (a.type != cudaMemoryTypeDevice) { if (bind_index(0)) return NULL; return s; }      // managed / mapped host memory: the first device serves it
	int k = -1;
	for (int i = 0; i < g_ndev; i++) if (g_ctx[i].device == a.device) k = i;
	if (k < 0) { rtkd_set_error("the buffer lives on device %d, which is not in the library's device list", a.device); return NULL; }
	if (k == 0) { if (bind_index(0)) return NULL; return s; }
	if (rtkd_sync_replicas(s)) return NULL;
	if (bind_index(k)) return NULL;
	return s->replica[k];
}

extern "C" uint32_t rtkd_scene_status(rtkd_scene *s)
{
	uint32_t v = s->h_status ? *(volatile uint32_t*)s->h_status : 0u;
	for (int k = 1; k < RTKD_MAX_DEVICES; k++) if (s->replica[k] && s->replica[k]->h_status) v |= *(volatile uint32_t*)s->replica[k]->h_status;
	return v;
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
	if (bind_index(s->dev_index)) return RTKD_ERR_CUDA;
	s->epoch++;
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
	unsigned char *nleaf;          // [binary node] leaf counts of its two children, 4 bits each (k_collapse_prep)
	int4 *rec;                     // [binary node] children and their areas
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
		// small subtrees are disjoint, non-empty triangle ranges: there can never be more than n of them,
		// however skewed the splits are (a peeling split emits one per level); active large nodes are
		// disjoint ranges of more than RTK_SAH_SMALL triangles each, so fewer than n / RTK_SAH_SMALL
		B.small_cap = (size_t)n + 8;
		B.h.pb = A.take<float4>(2 * (size_t)n);
		B.h.idx0 = A.take<uint32_t>(n); B.h.idx1 = A.take<uint32_t>(n); B.h.idx_final = A.take<uint32_t>(n);
		B.h.left = A.take<int>(cap2); B.h.right = A.take<int>(cap2); B.h.first = A.take<int>(cap2); B.h.last = A.take<int>(cap2);
		B.h.blo = A.take<float4>(cap2); B.h.bhi = A.take<float4>(cap2); B.h.ndepth = A.take<uint32_t>(cap2);
		B.h.counters = A.take<uint32_t>(16);
		B.h.act_in = A.take<uint32_t>(B.act_cap); B.h.act_out = A.take<uint32_t>(B.act_cap);
		B.h.small_list = A.take<uint32_t>(B.small_cap);
		B.h.chunk_cap = (uint32_t)((size_t)n / RTK_SAH_CHUNK + 1 + B.act_cap);
		B.h.chunk_base = A.take<uint32_t>(B.act_cap + 1); B.h.chunk_base_out = A.take<uint32_t>(B.act_cap + 1);
		B.h.chunk_node = A.take<uint32_t>(B.h.chunk_cap); B.h.chunk_node_out = A.take<uint32_t>(B.h.chunk_cap);
		B.h.bins = A.take<uint32_t>(B.act_cap * RTK_SAH_NODEBINS);
		B.h.binpack = A.take<uint32_t>(n);
		B.h.split = A.take<int4>(B.act_cap); B.h.cursor = A.take<uint32_t>(2 * B.act_cap);
		B.h.node_cap = (uint32_t)cap2;
		B.h.act_cap = (uint32_t)B.act_cap; B.h.small_cap = (uint32_t)B.small_cap;
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
	B.nleaf = A.take<unsigned char>(2 * (size_t)n + 2);
	B.rec = A.take<int4>(use_sah ? 2 * (size_t)n + 2 : (size_t)n);
	B.d_cost = A.take<double>(1);
}

// binned-SAH binary tree (k_sah.cuh).  The level loop runs on the device's word: the split kernel of a level
// leaves the next level's node list, chunk table and counts in device memory (no planning kernel in between), the
// level's kernels are launched with upper-bound grids (at most 2^level nodes, never more than n / RTK_SAH_SMALL;
// chunks <= n / RTK_SAH_CHUNK + nodes) and blocks beyond the counts return at once.  The host looks at the counters
// only after the first ceil(log2(n / RTK_SAH_SMALL)) levels -- no tree is done before that -- and then every second
// level (round 1 read them back after every level: 13 round trips per 1M-triangle build).
static int build_sah(cudaStream_t st, const float4 *tri, const uint32_t *svals, build_bufs &B, uint32_t n)
{
	RTK_NVTX("rtk_b200 build: binned SAH levels");
	rtkd_sah &h = B.h;
	RTK_LAUNCH(k_sah_prim_bounds, (n + 255) / 256, 256, st, tri, svals, n, (float4*)h.pb, h.idx0); CK_LAUNCH();
	RTK_LAUNCH(k_sah_root, 1, 32, st, h, (const uint32_t*)B.d_bounds, n); CK_LAUNCH();
	uint32_t hc[16];
	memset(hc, 0, sizeof(hc));
	uint32_t blind = 0;
	for (uint32_t m = n; m > RTK_SAH_SMALL; m = (m + 1) / 2) blind++;
	const uint32_t max_chunks = n / RTK_SAH_CHUNK + 1;
	uint32_t depth = 0;
	int src_buf = 0;
	bool done = n <= RTK_SAH_SMALL;
	while (!done) {
		const unsigned long long pow2 = depth < 40 ? 1ull << depth : ~0ull;
		const uint32_t act_bound = (uint32_t)(pow2 < B.act_cap ? pow2 : B.act_cap);
		const uint32_t chunk_bound = max_chunks + act_bound;
		const int par = (int)(depth & 1u);
		if (depth == 0) { RTK_LAUNCH(k_sah_bins_clear, act_bound, 128, st, h, par); CK_LAUNCH(); }      // later levels: cleared by the partition kernel of the level above
		RTK_LAUNCH(k_sah_bin_large, chunk_bound, 256, st, h, src_buf, par); CK_LAUNCH();
		RTK_LAUNCH(k_sah_split_large, act_bound, 96, st, h, depth, src_buf ^ 1, par); CK_LAUNCH();
		RTK_LAUNCH(k_sah_partition_large, chunk_bound, 256, st, h, src_buf, par); CK_LAUNCH();
		// the lists the split kernel has filled are the next level's
		{ uint32_t *tmp = h.act_in; h.act_in = h.act_out; h.act_out = tmp; }
		{ uint32_t *tmp = h.chunk_base; h.chunk_base = h.chunk_base_out; h.chunk_base_out = tmp; }
		{ uint32_t *tmp = h.chunk_node; h.chunk_node = h.chunk_node_out; h.chunk_node_out = tmp; }
		src_buf ^= 1;
		depth++;
		if (depth >= blind && ((depth - blind) & 1u) == 0) {
			CK(cudaMemcpyAsync(hc, h.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
			CK(cudaStreamSynchronize(st));
			done = hc[8 + 2 * (depth & 1u)] == 0;             // large nodes of the level that would run next
			if (depth > 4096) { rtkd_set_error("SAH builder does not terminate"); return RTKD_ERR_MEMORY; }
		}
	}
	if (n <= RTK_SAH_SMALL) {
		CK(cudaMemcpyAsync(hc, h.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
	}
	if (hc[2] > B.small_cap || hc[3]) { rtkd_set_error("SAH builder ran out of list space (flags %u)", hc[3]); return RTKD_ERR_MEMORY; }
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
	if (bind_index(s->dev_index)) return RTKD_ERR_CUDA;
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
		s->epoch++;
		return r;
	}
	if (num_words < words) { rtkd_set_error("triangle filter needs %zu words for %u triangles, got %zu", words, s->num_tris, num_words); return RTKD_ERR_ARGUMENT; }
	if (!words) return RTKD_OK;
	if (!s->filter_bits) CK(cudaMalloc((uint32_t**)&s->filter_bits, sizeof(uint32_t) * words));
	CK(cudaMemcpyAsync(s->filter_bits, bits, sizeof(uint32_t) * words, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
	int r = apply_filter(s, st);
	// filters change rarely: return with the pass done, whatever stream the next query uses
	CK(cudaStreamSynchronize(st));
	s->epoch++;
	return r;
}

extern "C" int rtkd_build(rtkd_scene *s, int mode, void *stream)
{
	if (bind_index(s->dev_index)) return RTKD_ERR_CUDA;
	cudaStream_t st = (cudaStream_t)stream;
	const uint32_t n = s->num_tris;
	RTK_NVTX("rtk_b200 build");
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
	for (int shift = use_sah ? 64 - g_sah_sort_bits : 0; shift < 64; shift += 8) {
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
		{
			// one record per binary node: its children, their areas and the leaves below them (k_collapse_prep)
			const uint32_t bound = use_sah ? (uint32_t)(2 * (size_t)n + 2) : n - 1;
			RTK_LAUNCH(k_collapse_prep, (bound + 255) / 256, 256, st, t, use_sah ? (const uint32_t*)B.h.counters : (const uint32_t*)NULL, n - 1, (int)n, B.rec, B.nleaf); CK_LAUNCH();
		}
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
				a.nodes = B.wide; a.n = (int)n; a.err = B.ctr + 2; a.rec = B.rec; a.nleaf2 = B.nleaf;
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

	// node array and traversal triangles (one 8-entry slot per leaf, sized now that the leaves are
	// counted): one arena, reused when it is large enough
	const uint32_t num_tv = num_leaves * RTK_LEAF_MAX;
	{
		int ar = scene_arena_layout(s, num_nodes, num_tv, true);
		if (ar) return ar;
	}
	if (!s->node_level || s->node_level_cap < num_nodes) {
		if (s->node_level) cudaFree(s->node_level);
		s->node_level = NULL;
		const uint32_t cap = num_nodes + num_nodes / 16 + 64;
		CK(cudaMalloc((unsigned char**)&s->node_level, (size_t)cap));
		s->node_level_cap = cap;
	}
	CK(cudaMemcpyAsync(s->node_level, B.node_level, (size_t)num_nodes, cudaMemcpyDeviceToDevice, st));
	CK(cudaMemcpyAsync(s->nodes, B.wide, sizeof(float4) * 16 * (size_t)num_nodes, cudaMemcpyDeviceToDevice, st));
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
	s->epoch++;
	if (s->h_status) *s->h_status = 0;          // a new tree: whatever the old one could not hold is history
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
	if (bind_index(s->dev_index)) return RTKD_ERR_CUDA;
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
	s->epoch++;
	if (s->h_status) *s->h_status = 0;
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

// A traversal launch takes one of the scene's slots (ray cursor, statistics, spill slab of the stack).
// slot_acquire returns with the scene's slot lock HELD; slot_release records the slot's event on the
// stream and drops the lock, so that nobody can pick the slot between the choice and the launch.
static int slot_acquire(rtkd_scene *s, cudaStream_t st, rtkd_trace_slot **out)
{
	pthread_mutex_t *m = (pthread_mutex_t*)s->slot_lock;
	const size_t groups = (size_t)g_sm_count * g_trace_ctas * RTK_TRACE_WARPS * (32 / g_trace_lanes);
	const size_t need = (size_t)7 * s->depth + 8;            // at most 7 pushes per level of wide nodes
	size_t entries = need > RTK_STACK_SMEM ? need - RTK_STACK_SMEM : 0;
	if (entries < 8) entries = 8;
	if (g_stack_limit > 0) entries = (size_t)g_stack_limit;
	pthread_mutex_lock(m);
	int pick = -1;
	// (1) the slot this stream used last: stream order protects it
	for (int k = 0; k < RTKD_TRACE_SLOTS && pick < 0; k++) if (s->slot[k].used && s->slot[k].last_stream == (void*)st) pick = k;
	// (2) a slot whose last launch has finished  (3) a slot nobody has used yet
	for (int k = 0; k < RTKD_TRACE_SLOTS && pick < 0; k++) if (s->slot[k].used && cudaEventQuery((cudaEvent_t)s->slot[k].done) == cudaSuccess) pick = k;
	for (int k = 0; k < RTKD_TRACE_SLOTS && pick < 0; k++) if (!s->slot[k].used) pick = k;
	cudaGetLastError();                                      // cudaErrorNotReady of the queries above
	cudaError_t e = cudaSuccess;
	if (pick < 0) {
		// (4) every slot is busy on another stream: queue behind one of them
		static unsigned rr = 0;
		pick = (int)(__atomic_fetch_add(&rr, 1u, __ATOMIC_RELAXED) % RTKD_TRACE_SLOTS);
		e = cudaStreamWaitEvent(st, (cudaEvent_t)s->slot[pick].done, 0);
	}
	rtkd_trace_slot &T = s->slot[pick];
	if (e == cudaSuccess && !T.scratch) {
		e = cudaMalloc((unsigned char**)&T.scratch, 256);
		if (e == cudaSuccess) e = cudaMemset(T.scratch, 0, 256);
		cudaEvent_t ev = NULL;
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
		T.done = ev;
	}
	if (e == cudaSuccess && (!T.overflow || (g_stack_limit > 0 ? T.overflow_entries != entries : T.overflow_entries < entries) || T.overflow_groups < groups)) {
		// the tree got deeper since the slab was sized (cudaFree waits for whatever still reads it)
		if (T.overflow) cudaFree(T.overflow);
		T.overflow = NULL;
		e = cudaMalloc((uint2**)&T.overflow, sizeof(uint2) * entries * groups);
		T.overflow_entries = entries; T.overflow_groups = groups;
	}
	if (e != cudaSuccess) {
		rtkd_set_error("traversal scratch: %s", cudaGetErrorString(e));
		pthread_mutex_unlock(m);
		return e == cudaErrorMemoryAllocation ? RTKD_ERR_MEMORY : RTKD_ERR_CUDA;
	}
	T.used = 1; T.last_stream = (void*)st;
	*out = &T;
	return RTKD_OK;
}
static void slot_release(rtkd_scene *s, rtkd_trace_slot *T, cudaStream_t st)
{
	cudaEventRecord((cudaEvent_t)T->done, st);
	pthread_mutex_unlock((pthread_mutex_t*)s->slot_lock);
}

// one instantiation per (lanes per ray, cull mode, statistics, any-hit, leaf-phase variant); every launch
// carries the scene's access-policy window (RTK_LAUNCH_WIN: nodes + leaf slots persist in L2)
template <int L, bool PD>
static void launch_trace_l(int cull_mode, bool stats, unsigned grid, cudaStream_t st, const rtkd_trace_args &p, void *wb, size_t wn, float wr)
{
	const bool c1 = (cull_mode & 1) != 0;
	if (cull_mode & 2) {
		// occlusion query (any hit): the output is a byte per ray
		if (c1) { RTK_LAUNCH_WIN((k_trace<L, 1, false, true, PD>), grid, RTK_TRACE_THREADS, st, wb, wn, wr, p); }
		else { RTK_LAUNCH_WIN((k_trace<L, 0, false, true, PD>), grid, RTK_TRACE_THREADS, st, wb, wn, wr, p); }
	} else if (stats) {
		if (c1) { RTK_LAUNCH_WIN((k_trace<L, 1, true, false, PD>), grid, RTK_TRACE_THREADS, st, wb, wn, wr, p); }
		else { RTK_LAUNCH_WIN((k_trace<L, 0, true, false, PD>), grid, RTK_TRACE_THREADS, st, wb, wn, wr, p); }
	} else {
		if (c1) { RTK_LAUNCH_WIN((k_trace<L, 1, false, false, PD>), grid, RTK_TRACE_THREADS, st, wb, wn, wr, p); }
		else { RTK_LAUNCH_WIN((k_trace<L, 0, false, false, PD>), grid, RTK_TRACE_THREADS, st, wb, wn, wr, p); }
	}
}
static void launch_trace(int lanes, int cull_mode, bool stats, bool pd, unsigned grid, cudaStream_t st, const rtkd_trace_args &p, void *wb, size_t wn, float wr)
{
	if (lanes == 8) launch_trace_l<8, false>(cull_mode, stats, grid, st, p, wb, wn, wr);
	else if (lanes == 4) { if (pd) launch_trace_l<4, true>(cull_mode, stats, grid, st, p, wb, wn, wr); else launch_trace_l<4, false>(cull_mode, stats, grid, st, p, wb, wn, wr); }
	else { if (pd) launch_trace_l<2, true>(cull_mode, stats, grid, st, p, wb, wn, wr); else launch_trace_l<2, false>(cull_mode, stats, grid, st, p, wb, wn, wr); }
}

extern "C" int rtkd_trace(rtkd_scene *s, const void *d_rays, void *d_hit16, size_t n, int cull_mode,
                          rtkd_trace_stats *stats, void *stream)
{
	if (!n) { if (stats) memset(stats, 0, sizeof(*stats)); return RTKD_OK; }
	if (n > 0xfffffff0ull) { rtkd_set_error("batch too large (%zu rays); split it", n); return RTKD_ERR_ARGUMENT; }
	if (((uintptr_t)d_rays & 15) || (!(cull_mode & 2) && ((uintptr_t)d_hit16 & 15))) { rtkd_set_error("device ray / hit buffers must be 16-byte aligned"); return RTKD_ERR_ARGUMENT; }
	if ((cull_mode & 2) && stats) { rtkd_set_error("no statistics variant of the occlusion query"); return RTKD_ERR_ARGUMENT; }
	if (bind_index(s->dev_index)) return RTKD_ERR_CUDA;
	cudaStream_t st = (cudaStream_t)stream;
	rtkd_trace_slot *T = NULL;
	int r = slot_acquire(s, st, &T);
	if (r) return r;
	cudaError_t e = cudaMemsetAsync(T->scratch, 0, 128, st);
	rtkd_trace_args p;
	fill_arrays(s, p.sc);
	p.rays = (const float4*)d_rays; p.out = (float4*)d_hit16; p.nrays = (uint32_t)n;
	p.counter = (uint32_t*)T->scratch; p.status = s->d_status;
	p.stats = (unsigned long long*)((unsigned char*)T->scratch + 64);
	p.overflow = (uint2*)T->overflow; p.ovf_entries = (uint32_t)T->overflow_entries;
	// persistent grid: one wave of resident CTAs, never more CTAs than ray batches
	size_t batches = (n + RTK_RAY_BATCH - 1) / RTK_RAY_BATCH;
	int reserved = g_reserved_sms + t_reserve_extra;
	if (reserved >= g_sm_count) reserved = g_sm_count - 1;
	size_t ctas = (size_t)(g_sm_count - reserved) * g_trace_ctas;
	size_t want = (batches + RTK_TRACE_WARPS - 1) / RTK_TRACE_WARPS;
	unsigned grid = (unsigned)(want < ctas ? want : ctas);
	// L2 window: the arena (nodes first, then the leaf slots) as far as the device allows
	void *wb = NULL; size_t wn = 0; float wr = 1.0f;
	if (g_l2_persist && g_l2_window_max && s->arena && s->arena_used) {
		wb = s->arena;
		wn = s->arena_used < g_l2_window_max ? s->arena_used : g_l2_window_max;
		const size_t setaside = g_l2_setaside;
		wr = wn <= setaside ? 1.0f : (float)((double)setaside / (double)wn);
	}
	if (e == cudaSuccess) {
		launch_trace(g_trace_lanes, cull_mode, stats != NULL, g_trace_pd != 0, grid, st, p, wb, wn, wr);
		e = cudaGetLastError();
	}
	if (e != cudaSuccess) {
		rtkd_set_error("k_trace launch: %s", cudaGetErrorString(e));
		slot_release(s, T, st);
		return RTKD_ERR_CUDA;
	}
	if (stats) {
		// slow path: the counters are read before the slot can go to anybody else
		unsigned long long h[8];
		e = cudaMemcpyAsync(h, p.stats, sizeof(unsigned long long) * 6, cudaMemcpyDeviceToHost, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		slot_release(s, T, st);
		if (e != cudaSuccess) { rtkd_set_error("trace statistics: %s", cudaGetErrorString(e)); return RTKD_ERR_CUDA; }
		stats->rays = n; stats->hits = h[1]; stats->node_visits = h[2]; stats->leaf_visits = h[3];
		stats->tri_tests = h[4]; stats->stack_max = h[5];
		if (*(volatile uint32_t*)s->h_status & 2u) { rtkd_set_error("traversal stack exhausted"); return RTKD_ERR_OVERFLOW; }
		return RTKD_OK;
	}
	slot_release(s, T, st);
	return RTKD_OK;
}

extern "C" int rtkd_trace_brute(rtkd_scene *s, const void *d_rays, void *d_hit16, size_t n, void *stream)
{
	if (bind_index(s->dev_index)) return RTKD_ERR_CUDA;
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
	if (bind_index(s->dev_index)) return RTKD_ERR_CUDA;
	rtkd_arrays a;
	fill_arrays(s, a);
	RTK_LAUNCH(k_resolve<false>, (unsigned)((n + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS), RTK_RESOLVE_THREADS, stream,
	           a, (const float4*)d_hit16, (uint32_t*)d_hits, (unsigned char*)d_mask, (uint32_t)n, (unsigned long long*)NULL, (uint32_t*)NULL);
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
	if (bind_index(s->dev_index)) return RTKD_ERR_CUDA;
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
	if (bind_index(s->dev_index)) return NULL;
	if (s->hit16_cap < n) {
		if (s->hit16) cudaFree(s->hit16);
		s->hit16 = NULL; s->hit16_cap = 0;
		size_t cap = n + n / 8 + 1024;
		CKP(cudaMalloc((float4**)&s->hit16, sizeof(float4) * cap));
		s->hit16_cap = cap;
	}
	return s->hit16;
}

// ---------------------------------------------------------------------------------------------
// Host-buffer batches (rtk_trace_rays / rtk_trace_rays_compact): the reference's contract is that ONE
// process calls the library and gets the machine (rtk.h:126-129), so a batch is split into contiguous
// ranges over every device of the list; each device runs its range as a pipeline over chunks of
// RTKD_HOST_CHUNK rays, RTKD_HOST_BUFS chunks in flight on their own streams, driven by the device's
// worker thread (the caller's thread drives the first device).
//
//   upload     the rays of the range go to one device buffer, chunk by chunk on a copy stream of
//              their own that runs RTKD_HOST_AHEAD chunks ahead of the kernels
//   rows, DIRECT (the caller's hits / mask arrays are page-locked: cudaHostAlloc, cudaHostRegister,
//              rtk_cuda_host_alloc, torch pin_memory ...): k_trace -> k_resolve writes the 68-byte rows of
//              the rays that hit, and the mask bytes, STRAIGHT INTO THE CALLER'S ARRAYS over PCIe.
//              No staging, no host threads, nothing for the CPUs to do: the links are the limit, and
//              there is one per GPU.  Rows of rays that missed are never touched (rtk.c:571-576).
//   rows, STAGED (pageable arrays): k_trace -> k_resolve<dense> -> D2H of the mask bytes, block bases and
//              hit count -> D2H of exactly the hit rows -> the host worker pool (rtk_place.c) copies
//              each row to hits[i].
//   compact    k_trace -> D2H of the chunk's 16-byte records into the caller's array.
// Every traversal launch takes its own slot of the scene, so the chunks of a batch -- and batches of
// different host threads -- overlap freely.
// ---------------------------------------------------------------------------------------------
#define RTKD_HOST_CHUNK ((size_t)1 << 20)
#define RTKD_HOST_AHEAD 4
#define RTKD_HOST_SMALL ((size_t)2048)          // batches up to this size take the one-stream, one-synchronisation path

struct batch_job {
	rtkd_scene *s;                      // the copy of the scene on this device
	const char *rays; char *hits; unsigned char *mask;     // the caller's WHOLE arrays
	size_t first, n;                    // this device's range of them
	int mode;                           // 0: rtk_hit rows + mask, 1: compact records (hits = rtk_cuda_hit16[])
	int ndev;                           // devices that share this batch
	bool rays_pinned;                   // the caller's ray array is page-locked
	long long found;                    // rays that hit (mode 0)
	int rc;
	char err[320];
};

static void stage_release(host_stage &G)
{
	for (int k = 0; k < RTKD_HOST_BUFS; k++) {
		host_buf &B = G.b[k];
		cudaFree(B.d_h16); cudaFree(B.d_rows); cudaFree(B.d_base); cudaFree(B.d_mask);
		cudaFreeHost(B.h_meta); cudaFreeHost(B.h_rows);
		B.d_h16 = NULL; B.d_rows = B.d_base = NULL; B.d_mask = NULL; B.h_meta = B.h_rows = NULL;
	}
	// the bounce buffers of pageable uploads are sized by the chunk as well
	for (int k = 0; k < RTKD_HOST_RING; k++) { if (G.h_up[k]) cudaFreeHost(G.h_up[k]); G.h_up[k] = NULL; G.up_used[k] = false; }
	G.ready = G.staged = false;
}

static void stage_shutdown(dev_ctx &X)
{
	host_stage &G = X.stage;
	stage_release(G);
	if (G.streams) {
		for (int k = 0; k < RTKD_HOST_BUFS; k++) {
			host_buf &B = G.b[k];
			if (B.st) cudaStreamDestroy(B.st);
			if (B.traced) cudaEventDestroy(B.traced);
			if (B.meta_done) cudaEventDestroy(B.meta_done);
			if (B.rows_done) cudaEventDestroy(B.rows_done);
		}
		if (G.up) cudaStreamDestroy(G.up);
		if (G.push) cudaStreamDestroy(G.push);
		for (int k = 0; k < RTKD_HOST_RING; k++) if (G.up_read[k]) cudaEventDestroy(G.up_read[k]);
		for (int k = 0; k < RTKD_HOST_RING; k++) if (G.uploaded[k]) cudaEventDestroy(G.uploaded[k]);
	}
	if (G.d_rays) cudaFree(G.d_rays);
	if (G.d_count) cudaFree(G.d_count);
	if (G.sm_st) cudaStreamDestroy(G.sm_st);
	cudaFree(G.sm_d_rays); cudaFree(G.sm_d_h16); cudaFree(G.sm_d_rows); cudaFree(G.sm_d_mask);
	cudaFreeHost(G.sm_h_rows); cudaFreeHost(G.sm_h_mask);
	memset(&G, 0, sizeof(G));
}

// buffers of one device for chunks of up to `want` rays; `staged` adds the dense-row buffers and the
// pinned staging that only the placement path needs
static int stage_prepare(host_stage &G, size_t want, bool staged)
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
	const bool fits = G.ready && (forced ? G.chunk == chunk : G.chunk >= chunk);
	if (fits && (!staged || G.staged)) return RTKD_OK;
	if (!fits) stage_release(G);             // also what an earlier, failed attempt left behind
	else chunk = G.chunk;                    // only the staging half is missing
	if (!G.streams) {
		for (int k = 0; k < RTKD_HOST_BUFS; k++) {
			host_buf &B = G.b[k];
			CK(cudaStreamCreateWithFlags(&B.st, cudaStreamNonBlocking));
			CK(cudaEventCreateWithFlags(&B.traced, cudaEventDisableTiming));
			CK(cudaEventCreateWithFlags(&B.meta_done, cudaEventDisableTiming));
			CK(cudaEventCreateWithFlags(&B.rows_done, cudaEventDisableTiming));
		}
		CK(cudaStreamCreateWithFlags(&G.up, cudaStreamNonBlocking));
		CK(cudaStreamCreateWithFlags(&G.push, cudaStreamNonBlocking));
		for (int k = 0; k < RTKD_HOST_RING; k++) CK(cudaEventCreateWithFlags(&G.uploaded[k], cudaEventDisableTiming));
		for (int k = 0; k < RTKD_HOST_RING; k++) CK(cudaEventCreateWithFlags(&G.up_read[k], cudaEventDisableTiming));
		CK(cudaMalloc(&G.d_count, 64));
		G.streams = true;
	}
	G.chunk = chunk;
	G.blocks = ((chunk + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS + 3) & ~(size_t)3;   // the 64-bit count follows: keep it aligned
	G.meta_bytes = ((chunk + 15) & ~(size_t)15) + 4 * G.blocks + 16;
	for (int k = 0; k < RTKD_HOST_BUFS; k++) {
		host_buf &B = G.b[k];
		if (!B.d_h16) CK(cudaMalloc(&B.d_h16, 16 * chunk));
		if (!B.d_mask) CK(cudaMalloc(&B.d_mask, (chunk + 15) & ~(size_t)15));
		if (!B.d_rows) CK(cudaMalloc(&B.d_rows, 68 * chunk));
		if (staged) {
			if (!B.d_base) CK(cudaMalloc(&B.d_base, 4 * G.blocks + 16));
			if (!B.h_meta) CK(cudaMallocHost(&B.h_meta, G.meta_bytes));
			if (!B.h_rows) CK(cudaMallocHost(&B.h_rows, 68 * chunk));
		}
		B.state = 0; B.ticket = -1;
	}
	G.ready = true;
	if (staged) G.staged = true;
	return RTKD_OK;
}

static int stage_rays(host_stage &G, size_t n)
{
	if (G.rays_cap >= n) return RTKD_OK;
	if (G.d_rays) cudaFree(G.d_rays);
	G.d_rays = NULL; G.rays_cap = 0;
	if (cudaMalloc(&G.d_rays, 32 * n) != cudaSuccess) { cudaGetLastError(); rtkd_set_error("out of device memory for %zu rays", n); return RTKD_ERR_MEMORY; }
	G.rays_cap = n;
	return RTKD_OK;
}

// device-visible address of page-locked host memory [p, p + bytes), or NULL when the range is pageable
static void *host_mapped(const void *p, size_t bytes)
{
	if (!p || !bytes) return NULL;
	cudaPointerAttributes a, b;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess || cudaPointerGetAttributes(&b, (const char*)p + bytes - 1) != cudaSuccess) { cudaGetLastError(); return NULL; }
	if (a.type != cudaMemoryTypeHost || b.type != cudaMemoryTypeHost || !a.devicePointer) return NULL;
	return a.devicePointer;
}

#define PIPE_CK(call, what) do { cudaError_t _e = (call); if (_e != cudaSuccess) { rtkd_set_error("%s: %s", what, cudaGetErrorString(_e)); rc = RTKD_ERR_CUDA; } } while (0)

// keep the upload stream RTKD_HOST_AHEAD chunks ahead of the kernels
static int pipe_uploads(host_stage &G, const batch_job &J, size_t it, size_t nchunks, size_t &uploads, const char *what)
{
	int rc = RTKD_OK;
	for (; uploads < nchunks && uploads <= it + RTKD_HOST_AHEAD && rc == RTKD_OK; uploads++) {
		const size_t off = uploads * G.chunk, cnt = J.n - off < G.chunk ? J.n - off : G.chunk;
		const char *src = J.rays + 32 * (J.first + off);
		const int k = (int)(uploads % RTKD_HOST_RING);
		if (!J.rays_pinned) {
			// pageable rays: the worker pool copies the chunk into a pinned bounce buffer, the copy engine takes it from there
			if (!G.h_up[k] && cudaMallocHost(&G.h_up[k], 32 * G.chunk) != cudaSuccess) { cudaGetLastError(); G.h_up[k] = NULL; }
			if (G.h_up[k]) {
				if (G.up_used[k]) PIPE_CK(cudaEventSynchronize(G.up_read[k]), what);       // the upload that last read this buffer
				if (rc) break;
				// one device: the worker pool copies the chunk in parallel; several devices: every device's own
				// driving thread copies its chunk (the pool has the rows to place, and there are already ndev copiers)
				if (J.ndev > 1) memcpy(G.h_up[k], src, 32 * cnt);
				else rtkd_place_wait(rtkd_copy_submit(G.h_up[k], src, 32 * cnt));
				src = (const char*)G.h_up[k];
			}
		}
		PIPE_CK(cudaMemcpyAsync(G.d_rays + 2 * off, src, 32 * cnt, cudaMemcpyHostToDevice, G.up), what);
		if (rc == RTKD_OK && !J.rays_pinned && G.h_up[k]) { PIPE_CK(cudaEventRecord(G.up_read[k], G.up), what); G.up_used[k] = true; }
		if (rc == RTKD_OK) PIPE_CK(cudaEventRecord(G.uploaded[k], G.up), what);
	}
	return rc;
}

static int pipe_finish(dev_ctx &X, const batch_job &J, int rc, const char *what)
{
	host_stage &G = X.stage;
	for (int k = 0; k < RTKD_HOST_BUFS; k++) if (cudaStreamSynchronize(G.b[k].st) != cudaSuccess && rc == RTKD_OK) { rtkd_set_error("%s: device failure", what); rc = RTKD_ERR_CUDA; }
	if (cudaStreamSynchronize(G.up) != cudaSuccess && rc == RTKD_OK) { rtkd_set_error("%s: device failure", what); rc = RTKD_ERR_CUDA; }
	if (rc == RTKD_OK && (*(volatile uint32_t*)J.s->h_status & 2u)) { rtkd_set_error("traversal stack exhausted"); rc = RTKD_ERR_OVERFLOW; }
	return rc;
}

// stage B of one buffer: wait for the chunk's count, then fetch exactly its rows
static int host_stage_b(host_stage &G, host_buf &B)
{
	CK(cudaEventSynchronize(B.meta_done));
	const size_t mask_bytes = (G.chunk + 15) & ~(size_t)15;
	unsigned long long hc = 0;
	memcpy(&hc, B.h_meta + mask_bytes + 4 * G.blocks, sizeof(hc));
	B.hits = (size_t)hc;
	if (B.hits) CK(cudaMemcpyAsync(B.h_rows, B.d_rows, 68 * B.hits, cudaMemcpyDeviceToHost, B.st));
	CK(cudaEventRecord(B.rows_done, B.st));
	B.state = 2;
	return RTKD_OK;
}

// stage C: hand the rows to the placement workers
static int host_stage_c(host_stage &G, host_buf &B, const batch_job &J)
{
	CK(cudaEventSynchronize(B.rows_done));
	const size_t mask_bytes = (G.chunk + 15) & ~(size_t)15;
	rtkd_place_desc d;
	d.hits = J.hits; d.mask_out = J.mask;
	d.rows = B.h_rows; d.mask = B.h_meta; d.block_base = (const uint32_t*)(B.h_meta + mask_bytes);
	d.first_ray = J.first + B.off; d.nrays = B.cnt;
	B.ticket = rtkd_place_submit(&d);
	B.state = 3;
	return RTKD_OK;
}

// The rows of a host batch reach the caller's rtk_hit array in one of two ways, chosen CHUNK BY CHUNK:
//
//   direct   (needs page-locked caller arrays; m_hits / m_mask are their device-visible addresses)
//            k_resolve expands the chunk's rows in DEVICE memory; k_push_rows -- a few CTAs on SMs that the
//            traversal grid leaves free, on a stream of its own -- writes the rows of the rays that hit, and the
//            mask bytes, straight into the caller's arrays while the next chunks are traced.  No staging, no
//            host threads; but scattered 68-byte pieces cross PCIe at about 45 % efficiency (~25 GB/s measured),
//            and a warp that writes to host memory waits for the link, which is why the push is not done by the
//            resolve kernel itself (that was measured: 26 ms per 16.7M rays instead of 18).
//   staged   k_resolve<dense> packs the rows, the copy engine brings exactly those back at full link speed
//            (55 GB/s), and the host worker pool (rtk_place.c) puts each row where it belongs -- bound by the
//            host's memory system, a resource all devices share.
//
// staged_every = 0: every chunk direct; 1: every chunk staged (pageable arrays); k > 1: every k-th chunk
// staged, so that link and host threads both work (one device alone: k = 2; several devices: all direct).
static int pipeline_rows(dev_ctx &X, batch_job &J, uint32_t *m_hits, unsigned char *m_mask, unsigned staged_every)
{
	static const char *what = "rtk_trace_rays";
	host_stage &G = X.stage;
	const size_t chunk = G.chunk, nchunks = (J.n + chunk - 1) / chunk;
	const size_t mask_bytes = (chunk + 15) & ~(size_t)15;
	rtkd_arrays a;
	fill_arrays(J.s, a);
	int rc = RTKD_OK;
	long long total = 0;
	size_t uploads = 0;
	const unsigned push_ctas = (unsigned)((g_push_sms > 0 ? g_push_sms : 1) * (2048 / RTK_PUSH_THREADS));
	bool staged_of[RTKD_HOST_BUFS] = { false, false, false, false };     // mode of the chunk each buffer carries
	bool used[RTKD_HOST_BUFS] = { false, false, false, false };
	PIPE_CK(cudaMemsetAsync(G.d_count, 0, 8, G.up), what);          // ahead of the first upload event
	if (staged_every != 1) t_reserve_extra = g_push_sms;
	// chunk ci is enqueued in iteration ci; a staged chunk enters stage B in iteration ci+1 and stage C in
	// iteration ci+2; its buffer is reused in iteration ci + RTKD_HOST_BUFS
	for (size_t it = 0; it < nchunks + 2 && rc == RTKD_OK; it++) {
		RTK_NVTX("rtk_b200 host batch: chunk");
		rc = pipe_uploads(G, J, it, nchunks, uploads, what);
		if (rc) break;
		if (it < nchunks) {
			const int bi = (int)(it % RTKD_HOST_BUFS);
			host_buf &B = G.b[bi];
			if (B.state == 3) { rtkd_place_wait(B.ticket); total += (long long)B.hits; B.state = 0; }
			const bool staged = staged_every == 1 || (staged_every > 1 && it % staged_every == staged_every - 1);
			B.off = it * chunk; B.cnt = J.n - B.off < chunk ? J.n - B.off : chunk;
			PIPE_CK(cudaStreamWaitEvent(B.st, G.uploaded[it % RTKD_HOST_RING], 0), what);
			// the buffer's previous chunk: a staged one left through B.st itself, a direct one through the push stream
			if (rc == RTKD_OK && used[bi] && !staged_of[bi]) PIPE_CK(cudaStreamWaitEvent(B.st, B.rows_done, 0), what);
			if (rc) break;
			rc = rtkd_trace(J.s, G.d_rays + 2 * B.off, B.d_h16, B.cnt, 1, NULL, B.st);
			if (rc) break;
			const unsigned blocks = (unsigned)((B.cnt + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS);
			if (staged) {
				unsigned long long *d_count = (unsigned long long*)((unsigned char*)B.d_base + 4 * G.blocks);
				PIPE_CK(cudaMemsetAsync(d_count, 0, 16, B.st), what);
				RTK_LAUNCH(k_resolve<true>, blocks, RTK_RESOLVE_THREADS, B.st, a, (const float4*)B.d_h16, B.d_rows, B.d_mask, (uint32_t)B.cnt, d_count, B.d_base);
				PIPE_CK(cudaGetLastError(), what);
				// mask bytes and bases+count land in one pinned block: [mask | bases | count]
				if (rc == RTKD_OK) PIPE_CK(cudaMemcpyAsync(B.h_meta, B.d_mask, B.cnt, cudaMemcpyDeviceToHost, B.st), what);
				if (rc == RTKD_OK) PIPE_CK(cudaMemcpyAsync(B.h_meta + mask_bytes, B.d_base, 4 * G.blocks + 16, cudaMemcpyDeviceToHost, B.st), what);
				if (rc == RTKD_OK) PIPE_CK(cudaEventRecord(B.meta_done, B.st), what);
				if (rc) break;
				B.state = 1;
			} else {
				RTK_LAUNCH(k_resolve<false>, blocks, RTK_RESOLVE_THREADS, B.st, a, (const float4*)B.d_h16, B.d_rows, B.d_mask, (uint32_t)B.cnt, G.d_count, (uint32_t*)NULL);
				PIPE_CK(cudaGetLastError(), what);
				if (rc == RTKD_OK) PIPE_CK(cudaEventRecord(B.traced, B.st), what);
				if (rc == RTKD_OK) PIPE_CK(cudaStreamWaitEvent(G.push, B.traced, 0), what);
				if (rc) break;
				RTK_LAUNCH(k_push_rows, push_ctas < blocks ? push_ctas : blocks, RTK_PUSH_THREADS, G.push, (const uint32_t*)B.d_rows, (const unsigned char*)B.d_mask,
				           m_hits + 17 * (J.first + B.off), m_mask ? m_mask + J.first + B.off : (unsigned char*)NULL, (uint32_t)B.cnt);
				PIPE_CK(cudaGetLastError(), what);
				if (rc == RTKD_OK) PIPE_CK(cudaEventRecord(B.rows_done, G.push), what);
				B.state = 0;
			}
			staged_of[bi] = staged; used[bi] = true;
		}
		if (it >= 1 && it - 1 < nchunks && G.b[(it - 1) % RTKD_HOST_BUFS].state == 1) rc = host_stage_b(G, G.b[(it - 1) % RTKD_HOST_BUFS]);
		if (rc == RTKD_OK && it >= 2 && it - 2 < nchunks && G.b[(it - 2) % RTKD_HOST_BUFS].state == 2) rc = host_stage_c(G, G.b[(it - 2) % RTKD_HOST_BUFS], J);
	}
	t_reserve_extra = 0;
	// drain: placements still running, and -- after an error -- whatever is still queued
	for (int k = 0; k < RTKD_HOST_BUFS; k++) {
		host_buf &B = G.b[k];
		if (B.state == 3) { rtkd_place_wait(B.ticket); total += (long long)B.hits; }
		B.state = 0; B.ticket = -1;
	}
	if (cudaStreamSynchronize(G.push) != cudaSuccess && rc == RTKD_OK) { rtkd_set_error("%s: device failure", what); rc = RTKD_ERR_CUDA; }
	rc = pipe_finish(X, J, rc, what);
	if (rc == RTKD_OK && staged_every != 1) {
		unsigned long long hc = 0;
		PIPE_CK(cudaMemcpy(&hc, G.d_count, sizeof(hc), cudaMemcpyDeviceToHost), what);
		total += (long long)hc;
	}
	J.found = total;
	return rc;
}

// compact records: the copy engine puts each chunk's records straight into the caller's array
static int pipeline_compact(dev_ctx &X, batch_job &J)
{
	static const char *what = "rtk_trace_rays_compact";
	host_stage &G = X.stage;
	const size_t chunk = G.chunk, nchunks = (J.n + chunk - 1) / chunk;
	int rc = RTKD_OK;
	size_t uploads = 0;
	for (size_t it = 0; it < nchunks && rc == RTKD_OK; it++) {
		rc = pipe_uploads(G, J, it, nchunks, uploads, what);
		if (rc) break;
		// the chunk's stream still holds the copy of the chunk that used this buffer before: stream
		// order keeps the new traversal from overwriting records that have not left yet
		host_buf &B = G.b[it % RTKD_HOST_BUFS];
		const size_t off = it * chunk, cnt = J.n - off < chunk ? J.n - off : chunk;
		PIPE_CK(cudaStreamWaitEvent(B.st, G.uploaded[it % RTKD_HOST_RING], 0), what);
		if (rc) break;
		rc = rtkd_trace(J.s, G.d_rays + 2 * off, B.d_h16, cnt, 1, NULL, B.st);
		if (rc) break;
		PIPE_CK(cudaMemcpyAsync(J.hits + 16 * (J.first + off), B.d_h16, 16 * cnt, cudaMemcpyDeviceToHost, B.st), what);
	}
	return pipe_finish(X, J, rc, what);
}

// one device's share of a batch, on the thread that drives that device
static void run_job(dev_ctx &X, batch_job &J)
{
	RTK_NVTX("rtk_b200 host batch: one device's share");
	int rc = bind_index((int)(&X - g_ctx));
	uint32_t *m_hits = NULL;
	unsigned char *m_mask = NULL;
	bool direct = false;
	if (rc == RTKD_OK && J.mode == 0 && g_host_direct) {
		m_hits = (uint32_t*)host_mapped(J.hits, 68 * (J.first + J.n));
		m_mask = J.mask ? (unsigned char*)host_mapped(J.mask, J.first + J.n) : NULL;
		direct = m_hits && (!J.mask || m_mask);
	}
	// direct and staged chunks are mixed when the arrays allow direct writes and ONE device runs the batch: every
	// second chunk is staged (measured on a 16-vCPU host: 18.1 ms all direct, 17.5 every 2nd, 17.1 every 3rd, 20.0
	// every 4th chunk staged per 16.7M rays).  With several devices the host threads are the scarcer resource
	// (8 devices: 80.6 ms all direct against 84.0 with every 16th chunk staged): all direct.  RTK_B200_HOST_MIX=k
	// forces every k-th chunk onto the staged route, 0 none.
	unsigned staged_every = 1;
	if (direct) staged_every = g_host_mix >= 0 ? (unsigned)g_host_mix : (J.ndev > 1 ? 0u : 2u);
	J.rays_pinned = rc == RTKD_OK && host_mapped(J.rays + 32 * J.first, 32 * J.n) != NULL;
	if (rc == RTKD_OK) rc = stage_prepare(X.stage, J.n, J.mode == 0 && staged_every != 0);
	if (rc == RTKD_OK) rc = stage_rays(X.stage, J.n);
	if (rc == RTKD_OK) {
		if (J.mode == 1) rc = pipeline_compact(X, J);
		else rc = pipeline_rows(X, J, m_hits, m_mask, staged_every);
	}
	if (rc == RTKD_ERR_CUDA && !g_err[0]) rtkd_set_error("CUDA failure in a host batch");
	J.rc = rc;
	if (rc) { strncpy(J.err, g_err, sizeof(J.err) - 1); J.err[sizeof(J.err) - 1] = 0; }
}

static void *ctx_worker(void *arg)
{
	dev_ctx &X = *(dev_ctx*)arg;
	cudaSetDevice(X.device);
	pthread_mutex_lock(&X.jm);
	for (;;) {
		while (!X.quit && !(X.job && !X.job_done)) pthread_cond_wait(&X.jc, &X.jm);
		if (X.quit) break;
		batch_job *J = X.job;
		pthread_mutex_unlock(&X.jm);
		run_job(X, *J);
		pthread_mutex_lock(&X.jm);
		X.job_done = true;
		pthread_cond_broadcast(&X.jc);
	}
	pthread_mutex_unlock(&X.jm);
	return NULL;
}

// split [0, n) over the devices, run the shares, merge
static long long run_batch(rtkd_scene *s, const void *rays, void *hits, unsigned char *mask, size_t n, int mode)
{
	int use = 1;
	size_t per = n;
	// contiguous ranges, multiples of the resolve block (128 rays = 68 full 128-byte lines of rows)
	batch_split(n, &use, &per);
	if (use > 1 && rtkd_sync_replicas(s) != RTKD_OK) return -1;
	batch_job J[RTKD_MAX_DEVICES];
	size_t first = 0;
	int used = 0;
	for (int k = 0; k < use && first < n; k++, used++) {
		batch_job &j = J[k];
		memset(&j, 0, sizeof(j));
		j.s = k == 0 ? s : s->replica[k];
		j.rays = (const char*)rays; j.hits = (char*)hits; j.mask = mask;
		j.first = first; j.n = (k == use - 1 || n - first < per) ? n - first : per;
		j.mode = mode; j.ndev = use;
		first += j.n;
	}
	for (int k = 0; k < used; k++) pthread_mutex_lock(&g_ctx[k].lock);      // ascending order: no deadlock between batches
	for (int k = 1; k < used; k++) {
		dev_ctx &X = g_ctx[k];
		if (!X.thread_up) continue;                          // no worker (thread creation failed at start-up): run it here, below
		pthread_mutex_lock(&X.jm);
		X.job = &J[k]; X.job_done = false;
		pthread_cond_broadcast(&X.jc);
		pthread_mutex_unlock(&X.jm);
	}
	run_job(g_ctx[0], J[0]);
	for (int k = 1; k < used; k++) if (!g_ctx[k].thread_up) run_job(g_ctx[k], J[k]);
	for (int k = 1; k < used; k++) {
		dev_ctx &X = g_ctx[k];
		if (!X.thread_up) continue;
		pthread_mutex_lock(&X.jm);
		while (!X.job_done) pthread_cond_wait(&X.jc, &X.jm);
		X.job = NULL;
		pthread_mutex_unlock(&X.jm);
	}
	for (int k = used - 1; k >= 0; k--) pthread_mutex_unlock(&g_ctx[k].lock);
	bind_index(0);
	long long total = 0;
	for (int k = 0; k < used; k++) {
		if (J[k].rc) { rtkd_set_error("%s (device %d)", J[k].err, g_ctx[k].device); return -1; }
		total += J[k].found;
	}
	return total;
}

// Small batches -- rtk_trace_ray is a batch of one -- are bound by latency, not by bytes: one stream,
// rows expanded in place (no dense packing, no worker threads), ONE synchronisation.  Concurrent small
// batches of different host threads go to different devices of the list when there are several.
static int small_prepare(host_stage &G)
{
	if (G.sm_st) return RTKD_OK;
	CK(cudaMalloc(&G.sm_d_rays, 32 * RTKD_HOST_SMALL));
	CK(cudaMalloc(&G.sm_d_h16, 16 * RTKD_HOST_SMALL));
	CK(cudaMalloc(&G.sm_d_rows, 68 * RTKD_HOST_SMALL));
	CK(cudaMalloc(&G.sm_d_mask, RTKD_HOST_SMALL));
	CK(cudaMallocHost(&G.sm_h_rows, 68 * RTKD_HOST_SMALL));
	CK(cudaMallocHost(&G.sm_h_mask, RTKD_HOST_SMALL));
	CK(cudaStreamCreateWithFlags(&G.sm_st, cudaStreamNonBlocking));
	return RTKD_OK;
}

static long long small_batch(rtkd_scene *s0, const void *rays, void *hits, unsigned char *mask, size_t n)
{
	int k = 0;
	bool locked = false;
	for (int i = 0; i < g_ndev && !locked; i++) if (pthread_mutex_trylock(&g_ctx[i].lock) == 0) { k = i; locked = true; }
	if (!locked) { k = 0; pthread_mutex_lock(&g_ctx[0].lock); }
	dev_ctx &X = g_ctx[k];
	host_stage &G = X.stage;
	rtkd_scene *s = s0;
	int rc = RTKD_OK;
	if (k > 0) { rc = rtkd_sync_replicas(s0); s = s0->replica[k]; }
	if (rc == RTKD_OK) rc = bind_index(k);
	if (rc == RTKD_OK) rc = small_prepare(G);
	long long total = 0;
	if (rc == RTKD_OK) {
		rtkd_arrays a;
		fill_arrays(s, a);
		cudaStream_t st = G.sm_st;
		cudaError_t e = cudaMemcpyAsync(G.sm_d_rays, rays, 32 * n, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess) rc = rtkd_trace(s, G.sm_d_rays, G.sm_d_h16, n, 1, NULL, st);
		if (e == cudaSuccess && rc == RTKD_OK) {
			RTK_LAUNCH(k_resolve<false>, (unsigned)((n + RTK_RESOLVE_THREADS - 1) / RTK_RESOLVE_THREADS), RTK_RESOLVE_THREADS, st,
			           a, (const float4*)G.sm_d_h16, G.sm_d_rows, G.sm_d_mask, (uint32_t)n, (unsigned long long*)NULL, (uint32_t*)NULL);
			e = cudaGetLastError();
			if (e == cudaSuccess) e = cudaMemcpyAsync(G.sm_h_rows, G.sm_d_rows, 68 * n, cudaMemcpyDeviceToHost, st);
			if (e == cudaSuccess) e = cudaMemcpyAsync(G.sm_h_mask, G.sm_d_mask, n, cudaMemcpyDeviceToHost, st);
			if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		}
		if (e != cudaSuccess) { rtkd_set_error("rtk_trace_rays: %s", cudaGetErrorString(e)); rc = RTKD_ERR_CUDA; }
		if (rc == RTKD_OK && (*(volatile uint32_t*)s->h_status & 2u)) { rtkd_set_error("traversal stack exhausted"); rc = RTKD_ERR_OVERFLOW; }
		if (rc == RTKD_OK) {
			for (size_t i = 0; i < n; i++) {
				if (G.sm_h_mask[i]) { memcpy((char*)hits + 68 * i, G.sm_h_rows + 68 * i, 68); total++; }    // rows of misses stay untouched (rtk.c:571-576)
			}
			if (mask) memcpy(mask, G.sm_h_mask, n);
		}
	}
	pthread_mutex_unlock(&X.lock);
	bind_index(0);
	return rc == RTKD_OK ? total : -1;
}

extern "C" long long rtkd_trace_host(rtkd_scene *s, const void *rays, void *hits, unsigned char *mask, size_t n)
{
	if (!n) return 0;
	if (ensure_init()) return -1;
	if (n <= RTKD_HOST_SMALL) return small_batch(s, rays, hits, mask, n);
	return run_batch(s, rays, hits, mask, n, 0);
}

// Host-buffer batch with COMPACT results: rays up, one 16-byte record (t, u, v, global triangle
// number or RTKD miss) per ray down, straight into the caller's array -- no dense packing, no row
// placement.  32 bytes per ray go up and 16 come down, on the two directions of the link.
extern "C" int rtkd_trace_host_compact(rtkd_scene *s, const void *rays, void *hit16, size_t n)
{
	if (!n) return RTKD_OK;
	if (ensure_init()) return RTKD_ERR_NO_DEVICE;
	return run_batch(s, rays, hit16, NULL, n, 1) < 0 ? RTKD_ERR_CUDA : RTKD_OK;
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
	if (bind_index(s->dev_index)) return RTKD_ERR_CUDA;
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

// Blobs come from disk: nothing in one is trusted.  The host checks the section table, this kernel
// checks every reference the traversal and the resolve kernel would follow: a child is a leaf inside
// the slot array, or a node with a HIGHER index than its parent (the builder allocates children after
// their parents, so a well-formed tree has no other kind -- and a blob with a cycle cannot hang the
// traversal), and every triangle number of a leaf slot is inside the corner array.
__global__ void __launch_bounds__(256) k_validate_blob(const float4 *nodes, uint32_t num_nodes, const float4 *tv0, uint32_t num_tv,
                                                       uint32_t num_tris, uint32_t *bad)
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < num_nodes * RTK_WIDE) {
		const uint32_t ref = __float_as_uint(nodes[2ull * i].w), parent = i / RTK_WIDE;
		if (ref == RTK_REF_EMPTY) {
			// the traversal never looks at the reference of an empty slot: it relies on the slot's box being inverted
			const float4 lo = nodes[2ull * i], hi = nodes[2ull * i + 1];
			if (!(lo.x > hi.x && lo.y > hi.y && lo.z > hi.z)) atomicOr(bad, 8u);
		} else {
			if (rtk_ref_is_leaf(ref)) {
				const uint32_t first = rtk_leaf_first(ref);
				if ((first & 7u) || first + RTK_LEAF_MAX > num_tv) atomicOr(bad, 1u);
			} else if (ref >= num_nodes || ref <= parent) atomicOr(bad, 2u);
		}
	}
	if (i < num_tv) {
		const uint32_t id = __float_as_uint(tv0[i].w);
		if (id != RTK_MISS && id >= num_tris) atomicOr(bad, 4u);
	}
}

extern "C" rtkd_scene *rtkd_blob_read(const void *payload, size_t payload_size)
{
	if (ensure_init()) return NULL;
	rtkd_blob_sub b;
	if (payload_size < a128(sizeof(b))) { rtkd_set_error("scene blob truncated"); return NULL; }
	memcpy(&b, payload, sizeof(b));
	if (memcmp(&b.magic2, "B200RTK3", 8) != 0) { rtkd_set_error("blob was not written by rtk_b200 (device-layout magic missing)"); return NULL; }
	// section table: ordered, 128-byte aligned, inside the payload, sized by the counts
	{
		const uint64_t off[7] = { b.off_nodes, b.off_tv0, b.off_tv1, b.off_tv2, b.off_orig, b.off_mesh, (uint64_t)payload_size };
		const uint64_t len[6] = { 256ull * b.num_nodes, 16ull * b.num_tv, 16ull * b.num_tv, 16ull * b.num_tv, 48ull * b.num_tris, 4ull * ((uint64_t)b.num_meshes + 1) };
		bool ok = off[0] >= a128(sizeof(b));
		for (int k = 0; k < 6 && ok; k++) ok = (off[k] & 127) == 0 && off[k] <= off[k + 1] && len[k] <= off[k + 1] - off[k];
		if (ok) ok = b.num_tv == (uint64_t)b.num_leaves * RTK_LEAF_MAX && b.num_leaves <= RTK_MAX_LEAVES && b.num_tris <= 0x0fffffffu &&
		             b.depth <= RTKD_COLLAPSE_LEVELS && b.num_nodes <= 0x7fffffffu && (b.num_tris == 0 || (b.num_nodes >= 1 && b.depth >= 1)) &&
		             b.num_leaves <= b.num_tris && (b.num_tris == 0 || b.num_leaves >= 1);
		if (!ok) { rtkd_set_error("scene blob is corrupt or truncated (section table)"); return NULL; }
		const uint32_t *mf = (const uint32_t*)((const char*)payload + b.off_mesh);
		bool mok = mf[0] == 0 && mf[b.num_meshes] == b.num_tris;
		for (uint32_t m = 0; m < b.num_meshes && mok; m++) mok = mf[m] <= mf[m + 1];
		if (!mok) { rtkd_set_error("scene blob is corrupt (mesh table)"); return NULL; }
		for (int k = 0; k < 3; k++) if (!(b.bounds_min[k] <= b.bounds_max[k]) && b.num_tris) { rtkd_set_error("scene blob is corrupt (bounds)"); return NULL; }
		if (!(b.abs_max >= 0.0f) || !isfinite(b.abs_max)) { rtkd_set_error("scene blob is corrupt (bounds)"); return NULL; }
	}
	const char *p = (const char*)payload;
	rtkd_scene *s = rtkd_scene_new(b.num_tris, b.num_meshes, (const uint32_t*)(p + b.off_mesh));
	if (!s) return NULL;
	s->id = b.id;
	s->num_nodes = b.num_nodes; s->num_leaves = b.num_leaves; s->depth = b.depth; s->build_mode = b.build_mode;
	memcpy(s->bounds_min, b.bounds_min, 12); memcpy(s->bounds_max, b.bounds_max, 12);
	s->abs_max = b.abs_max; s->sah_cost = b.sah_cost; s->num_tv = b.num_tv;
	cudaError_t e = cudaSuccess;
	if (scene_arena_layout(s, b.num_nodes, b.num_tv, false) != RTKD_OK) { rtkd_scene_free(s); return NULL; }
	if (b.num_tris) {
		e = cudaMemcpy(s->tri_orig, p + b.off_orig, 48 * (size_t)b.num_tris, cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(s->tv0, p + b.off_tv0, 16 * (size_t)b.num_tv, cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(s->tv1, p + b.off_tv1, 16 * (size_t)b.num_tv, cudaMemcpyHostToDevice);
		if (e == cudaSuccess) e = cudaMemcpy(s->tv2, p + b.off_tv2, 16 * (size_t)b.num_tv, cudaMemcpyHostToDevice);
	}
	if (e == cudaSuccess && b.num_nodes) e = cudaMemcpy(s->nodes, p + b.off_nodes, 256 * (size_t)b.num_nodes, cudaMemcpyHostToDevice);
	uint32_t bad = 0;
	if (e == cudaSuccess && (b.num_nodes || b.num_tv)) {
		uint32_t *d_bad = NULL;
		e = cudaMalloc(&d_bad, sizeof(uint32_t));
		if (e == cudaSuccess) e = cudaMemset(d_bad, 0, sizeof(uint32_t));
		if (e == cudaSuccess) {
			const size_t items = (size_t)b.num_nodes * RTK_WIDE > b.num_tv ? (size_t)b.num_nodes * RTK_WIDE : b.num_tv;
			RTK_LAUNCH(k_validate_blob, (unsigned)((items + 255) / 256), 256, 0, (const float4*)s->nodes, b.num_nodes, (const float4*)s->tv0, b.num_tv, b.num_tris, d_bad);
			e = cudaGetLastError();
		}
		if (e == cudaSuccess) e = cudaMemcpy(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost);
		cudaFree(d_bad);
	}
	if (e != cudaSuccess || bad) {
		if (bad) rtkd_set_error("scene blob is corrupt (%s)", (bad & 2u) ? "child reference outside the tree" : (bad & 1u) ? "leaf reference outside the triangle slots" :
		                        (bad & 8u) ? "empty child slot without an inverted box" : "triangle number outside the scene");
		else rtkd_set_error("scene upload failed: %s", cudaGetErrorString(e));
		rtkd_scene_free(s);
		return NULL;
	}
	s->epoch++;
	return s;
}
