/*
 * rtk_oracle.c -- CPU restatement of the reference's closest-hit path.
 *
 * TEST INFRASTRUCTURE ONLY (see rtk_oracle.h).  Plain C, scalar, no SIMD and
 * no FMA: compile with -ffp-contract=off so that every multiply and add rounds
 * separately exactly like the reference's _mm_mul_ps/_mm_add_ps sequences.
 *
 * Citations are into /root/reference/rtk.c.
 */
#define _GNU_SOURCE
#include "rtk_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#define ORC_INF 3.402823e+38f   /* RTK_INF, rtk.h:11 */

int orc_num_cores(void)
{
	long n = sysconf(_SC_NPROCESSORS_ONLN);
	return n > 0 ? (int)n : 1;
}

static double orc_now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---------------------------------------------------------------------- */
/* Ray setup: rtk.c:550-566                                                */
/* ---------------------------------------------------------------------- */

typedef struct {
	int   kx, ky, kz;
	float sx, sy, sz;     /* shear */
	float ox, oy, oz;     /* origin permuted to (kx,ky,kz) */
} orc_setup;

/* SSE max: a > b ? a : b (second operand on NaN), rtk.c:148-151 */
static float sse_max(float a, float b) { return a > b ? a : b; }
static float sse_min(float a, float b) { return a < b ? a : b; }

static void orc_ray_setup(const orc_ray *r, orc_setup *s)
{
	float ax = fabsf(r->d[0]), ay = fabsf(r->d[1]), az = fabsf(r->d[2]);
	float mx = sse_max(sse_max(ax, ay), az);                 /* rtk.c:552 */
	int kz = (ax == mx) ? 0 : (ay == mx) ? 1 : 2;            /* rtk.c:553 */
	int kx = (kz + 1) % 3, ky = (kz + 2) % 3;                /* rtk.c:554-555 */
	s->kx = kx; s->ky = ky; s->kz = kz;
	s->sx = -r->d[kx] / r->d[kz];                            /* rtk.c:561 */
	s->sy = -r->d[ky] / r->d[kz];                            /* rtk.c:562 */
	s->sz = 1.0f / r->d[kz];                                 /* rtk.c:563 */
	s->ox = r->o[kx]; s->oy = r->o[ky]; s->oz = r->o[kz];    /* rtk.c:564-566 */
}

/* ---------------------------------------------------------------------- */
/* One triangle: rtk.c:256-354 for one lane, own-lane fp64 promotion       */
/* ---------------------------------------------------------------------- */

static int orc_tri(const orc_setup *s, const float *tri, float min_t, float max_t,
                   float *t_out, float *u_out, float *v_out)
{
	/* translate, rtk.c:256-258 (per vertex) */
	float a0x = tri[0 + s->kx] - s->ox, a0y = tri[0 + s->ky] - s->oy, a0z = tri[0 + s->kz] - s->oz;
	float a1x = tri[3 + s->kx] - s->ox, a1y = tri[3 + s->ky] - s->oy, a1z = tri[3 + s->kz] - s->oz;
	float a2x = tri[6 + s->kx] - s->ox, a2y = tri[6 + s->ky] - s->oy, a2z = tri[6 + s->kz] - s->oz;

	/* shear, rtk.c:284-292: add(v, mul(shear, vz)), two roundings */
	float m;
	m = s->sx * a0z; float x0 = a0x + m;
	m = s->sy * a0z; float y0 = a0y + m;
	float z0 = s->sz * a0z;
	m = s->sx * a1z; float x1 = a1x + m;
	m = s->sy * a1z; float y1 = a1y + m;
	float z1 = s->sz * a1z;
	m = s->sx * a2z; float x2 = a2x + m;
	m = s->sy * a2z; float y2 = a2y + m;
	float z2 = s->sz * a2z;

	/* edge functions, rtk.c:298-300: sub(mul, mul) */
	float p, q;
	p = x1 * y2; q = y1 * x2; float u = p - q;
	p = x2 * y0; q = y2 * x0; float v = p - q;
	p = x0 * y1; q = y0 * x1; float w = p - q;

	/* rtk.c:301-336, own-lane rule: promote iff one of THIS triangle's
	 * edge values is exactly zero */
	if (u == 0.0f || v == 0.0f || w == 0.0f) {
		double ud = (double)x1 * (double)y2 - (double)y1 * (double)x2;
		double vd = (double)x2 * (double)y0 - (double)y2 * (double)x0;
		double wd = (double)x0 * (double)y1 - (double)y0 * (double)x1;
		u = (float)ud; v = (float)vd; w = (float)wd;
	}

	/* rtk.c:340-344 */
	int neg = sse_min(sse_min(u, v), w) < 0.0f;
	int pos = sse_max(sse_max(u, v), w) > 0.0f;
	if (neg && pos) return 0;

	/* rtk.c:346-353 */
	float det = (u + v) + w;
	float rcp = 1.0f / det;
	float z = u * z0;
	m = v * z1; z = z + m;
	m = w * z2; z = z + m;
	float t = z * rcp;

	/* rtk.c:354 */
	if (!(t > min_t && t < max_t)) return 0;
	*t_out = t;
	*u_out = u * rcp;   /* rtk.c:363 */
	*v_out = v * rcp;   /* rtk.c:364 */
	return 1;
}

int orc_ray_triangle(const orc_ray *ray, const float *tri9, float max_t, float *t, float *u, float *v)
{
	orc_setup s;
	orc_ray_setup(ray, &s);
	return orc_tri(&s, tri9, ray->min_t, max_t, t, u, v);
}

/* ---------------------------------------------------------------------- */
/* Brute force                                                             */
/* ---------------------------------------------------------------------- */

typedef struct {
	const float *tri9; size_t ntris;
	const orc_ray *rays; orc_hit *out;
	size_t begin, end;
} brute_job;

static void *brute_worker(void *arg)
{
	brute_job *j = (brute_job*)arg;
	for (size_t r = j->begin; r < j->end; r++) {
		const orc_ray *ray = &j->rays[r];
		orc_setup s;
		orc_ray_setup(ray, &s);
		float best = ray->max_t;                 /* rtk.c:548 */
		orc_hit h = { 0.0f, 0.0f, 0.0f, 0xffffffffu };
		for (size_t i = 0; i < j->ntris; i++) {
			float t, u, v;
			/* triangles are visited in id order and the commit is a strict
			 * '<' (rtk.c:371), so an exact tie keeps the lowest id */
			if (orc_tri(&s, j->tri9 + 9 * i, ray->min_t, best, &t, &u, &v)) {
				best = t;
				h.t = t; h.u = u; h.v = v; h.prim = (uint32_t)i;
			}
		}
		j->out[r] = h;
	}
	return NULL;
}

static void run_jobs(void *(*fn)(void*), void *jobs, size_t job_size, int n)
{
	pthread_t th[256];
	if (n > 256) n = 256;
	for (int i = 1; i < n; i++) pthread_create(&th[i], NULL, fn, (char*)jobs + job_size * (size_t)i);
	fn(jobs);
	for (int i = 1; i < n; i++) pthread_join(th[i], NULL);
}

static int pick_threads(int threads, size_t n)
{
	if (threads <= 0) threads = orc_num_cores();
	if (threads > 256) threads = 256;
	if ((size_t)threads > n) threads = n ? (int)n : 1;
	return threads;
}

void orc_trace_brute(const float *tri9, size_t ntris, const orc_ray *rays, size_t nrays,
                     orc_hit *out, int threads)
{
	threads = pick_threads(threads, nrays);
	brute_job jobs[256];
	for (int i = 0; i < threads; i++) {
		jobs[i].tri9 = tri9; jobs[i].ntris = ntris; jobs[i].rays = rays; jobs[i].out = out;
		jobs[i].begin = nrays * (size_t)i / (size_t)threads;
		jobs[i].end = nrays * (size_t)(i + 1) / (size_t)threads;
	}
	run_jobs(brute_worker, jobs, sizeof(brute_job), threads);
}

/* ---- the same brute force, blocked and pre-filtered --------------------- */
/* For the big scenes (1M and 10M triangles) the plain loop above is minutes of CPU per thousand
 * rays.  This variant gives the SAME answers: the scalar orc_tri() above stays the only arbiter of
 * every hit; a vector pre-filter merely skips triangles that orc_tri() would reject at its sign
 * test (rtk.c:340-344).  The filter evaluates the three fp32 edge values of 8 triangles at a time
 * with the very operations of orc_tri (separate multiplies and adds/subtracts, minps/maxps
 * semantics) and drops a triangle only when none of them is exactly zero, the smallest is < 0 and
 * the largest is > 0 -- the case in which orc_tri returns 0 before looking at t.  Everything else
 * (exact zeros, NaNs, same-sign values) goes to orc_tri in triangle order, so ties still resolve
 * to the lowest id.  Triangles are transposed to SoA once and visited block by block for a batch
 * of rays, which keeps a block in cache.  tests/test_oracle.py checks it against orc_trace_brute. */
#if defined(__x86_64__)
#include <immintrin.h>
#define ORC_TB 2048             /* triangles per block: 9 x 8 KB */
#define ORC_RB 32               /* rays per batch */

typedef struct {
	const float *tri9; const float *soa; size_t ntris, npad;
	const orc_ray *rays; orc_hit *out;
	size_t begin, end;
} fast_job;

__attribute__((target("avx2")))
static void fast_block(const fast_job *j, size_t i0, size_t i1, const orc_ray *ray, const orc_setup *s,
                       float *best, orc_hit *h)
{
	const size_t np = j->npad;
	const float *c0x = j->soa + (0 + s->kx) * np, *c0y = j->soa + (0 + s->ky) * np, *c0z = j->soa + (0 + s->kz) * np;
	const float *c1x = j->soa + (3 + s->kx) * np, *c1y = j->soa + (3 + s->ky) * np, *c1z = j->soa + (3 + s->kz) * np;
	const float *c2x = j->soa + (6 + s->kx) * np, *c2y = j->soa + (6 + s->ky) * np, *c2z = j->soa + (6 + s->kz) * np;
	const __m256 ox = _mm256_set1_ps(s->ox), oy = _mm256_set1_ps(s->oy), oz = _mm256_set1_ps(s->oz);
	const __m256 sx = _mm256_set1_ps(s->sx), sy = _mm256_set1_ps(s->sy), zero = _mm256_setzero_ps();
	for (size_t i = i0; i < i1; i += 8) {
		__m256 a0z = _mm256_sub_ps(_mm256_loadu_ps(c0z + i), oz), a1z = _mm256_sub_ps(_mm256_loadu_ps(c1z + i), oz), a2z = _mm256_sub_ps(_mm256_loadu_ps(c2z + i), oz);
		__m256 x0 = _mm256_add_ps(_mm256_sub_ps(_mm256_loadu_ps(c0x + i), ox), _mm256_mul_ps(sx, a0z));
		__m256 y0 = _mm256_add_ps(_mm256_sub_ps(_mm256_loadu_ps(c0y + i), oy), _mm256_mul_ps(sy, a0z));
		__m256 x1 = _mm256_add_ps(_mm256_sub_ps(_mm256_loadu_ps(c1x + i), ox), _mm256_mul_ps(sx, a1z));
		__m256 y1 = _mm256_add_ps(_mm256_sub_ps(_mm256_loadu_ps(c1y + i), oy), _mm256_mul_ps(sy, a1z));
		__m256 x2 = _mm256_add_ps(_mm256_sub_ps(_mm256_loadu_ps(c2x + i), ox), _mm256_mul_ps(sx, a2z));
		__m256 y2 = _mm256_add_ps(_mm256_sub_ps(_mm256_loadu_ps(c2y + i), oy), _mm256_mul_ps(sy, a2z));
		__m256 u = _mm256_sub_ps(_mm256_mul_ps(x1, y2), _mm256_mul_ps(y1, x2));
		__m256 v = _mm256_sub_ps(_mm256_mul_ps(x2, y0), _mm256_mul_ps(y2, x0));
		__m256 w = _mm256_sub_ps(_mm256_mul_ps(x0, y1), _mm256_mul_ps(y0, x1));
		__m256 anyz = _mm256_or_ps(_mm256_or_ps(_mm256_cmp_ps(u, zero, _CMP_EQ_OQ), _mm256_cmp_ps(v, zero, _CMP_EQ_OQ)), _mm256_cmp_ps(w, zero, _CMP_EQ_OQ));
		__m256 neg = _mm256_cmp_ps(_mm256_min_ps(_mm256_min_ps(u, v), w), zero, _CMP_LT_OQ);
		__m256 pos = _mm256_cmp_ps(_mm256_max_ps(_mm256_max_ps(u, v), w), zero, _CMP_GT_OQ);
		int reject = _mm256_movemask_ps(_mm256_andnot_ps(anyz, _mm256_and_ps(neg, pos)));
		int keep = ~reject & 0xff;
		while (keep) {
			int k = __builtin_ctz((unsigned)keep);
			keep &= keep - 1;
			size_t id = i + (size_t)k;
			if (id >= j->ntris) break;
			float t, uu, vv;
			if (orc_tri(s, j->tri9 + 9 * id, ray->min_t, *best, &t, &uu, &vv)) {
				*best = t;
				h->t = t; h->u = uu; h->v = vv; h->prim = (uint32_t)id;
			}
		}
	}
}

static void *fast_worker(void *arg)
{
	fast_job *j = (fast_job*)arg;
	for (size_t r0 = j->begin; r0 < j->end; r0 += ORC_RB) {
		size_t nr = j->end - r0 < ORC_RB ? j->end - r0 : ORC_RB;
		orc_setup st[ORC_RB];
		float best[ORC_RB];
		orc_hit h[ORC_RB];
		for (size_t q = 0; q < nr; q++) {
			orc_ray_setup(&j->rays[r0 + q], &st[q]);
			best[q] = j->rays[r0 + q].max_t;
			h[q].t = h[q].u = h[q].v = 0.0f; h[q].prim = 0xffffffffu;
		}
		for (size_t i0 = 0; i0 < j->ntris; i0 += ORC_TB) {
			size_t i1 = i0 + ORC_TB < j->npad ? i0 + ORC_TB : j->npad;
			for (size_t q = 0; q < nr; q++) fast_block(j, i0, i1, &j->rays[r0 + q], &st[q], &best[q], &h[q]);
		}
		for (size_t q = 0; q < nr; q++) j->out[r0 + q] = h[q];
	}
	return NULL;
}
#endif

int orc_trace_brute_blocked(const float *tri9, size_t ntris, const orc_ray *rays, size_t nrays,
                            orc_hit *out, int threads)
{
#if defined(__x86_64__)
	if (!__builtin_cpu_supports("avx2") || !ntris) { orc_trace_brute(tri9, ntris, rays, nrays, out, threads); return 0; }
	size_t npad = (ntris + 7) & ~(size_t)7;
	float *soa = NULL;
	if (posix_memalign((void**)&soa, 64, sizeof(float) * 9 * npad)) { orc_trace_brute(tri9, ntris, rays, nrays, out, threads); return 0; }
	for (size_t c = 0; c < 9; c++) {
		float *d = soa + c * npad;
		for (size_t i = 0; i < ntris; i++) d[i] = tri9[9 * i + c];
		for (size_t i = ntris; i < npad; i++) d[i] = 0.0f;          /* padding never reaches orc_tri (id >= ntris) */
	}
	threads = pick_threads(threads, (nrays + ORC_RB - 1) / ORC_RB);
	fast_job jobs[256];
	size_t batches = (nrays + ORC_RB - 1) / ORC_RB;
	for (int i = 0; i < threads; i++) {
		jobs[i].tri9 = tri9; jobs[i].soa = soa; jobs[i].ntris = ntris; jobs[i].npad = npad; jobs[i].rays = rays; jobs[i].out = out;
		jobs[i].begin = batches * (size_t)i / (size_t)threads * ORC_RB;
		jobs[i].end = batches * (size_t)(i + 1) / (size_t)threads * ORC_RB;
		if (jobs[i].end > nrays) jobs[i].end = nrays;
	}
	run_jobs(fast_worker, jobs, sizeof(fast_job), threads);
	free(soa);
	return 1;
#else
	orc_trace_brute(tri9, ntris, rays, nrays, out, threads);
	return 0;
#endif
}

/* ---------------------------------------------------------------------- */
/* Reference blob structures (layout facts: rtk.c:69-86, rtk.h:78-89)      */
/* ---------------------------------------------------------------------- */

typedef struct { float pos[3]; uint32_t index; } ref_vertex;            /* 16 B */
typedef struct { float t, u, v; ref_vertex vertex[3]; uint32_t mesh_index, triangle_index; } ref_hit; /* 68 B */
typedef struct {
	char magic[8]; uint16_t endian; uint8_t sizeof_real, pad_0; uint32_t version, pad_1;
	uint64_t size_in_bytes, node_offset, leaf_offset, vertex_offset;
} ref_scene;                                                             /* 56 B */
typedef struct { float bx[2][4], by[2][4], bz[2][4]; uint64_t ptr[4]; } ref_node4;   /* 128 B */
typedef struct { uint8_t v[3]; uint8_t local_mesh; uint32_t tri; } ref_leaf_tri;    /* 8 B */

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void write_header(ref_scene *sc, size_t size, size_t node_off, size_t leaf_off, size_t vert_off)
{
	/* rtk.c:1737-1753 */
	memcpy(sc->magic, "\x00RTK\r\n\x1a\x0a", 8);
	sc->endian = 0xaabb; sc->sizeof_real = 4; sc->pad_0 = 0; sc->version = 1; sc->pad_1 = 0;
	sc->size_in_bytes = size; sc->node_offset = node_off; sc->leaf_offset = leaf_off; sc->vertex_offset = vert_off;
}

void orc_free(void *p) { free(p); }

/* ---- flat blob --------------------------------------------------------- */
/* layout: [0,56) header | [128,256) root node | [256,320) null leaf |
 *         [320, ...) leaf record | 64-aligned vertex group               */

static size_t flat_leaf_bytes(size_t n) { return align_up(8 + 8 * ((n + 3) & ~(size_t)3) + 4, 64); }
size_t orc_flat_blob_size(size_t n) { return align_up(320 + flat_leaf_bytes(n) + 16 * 3 * (n ? n : 1), 128); }

void orc_write_flat_blob(void *blob, const float *tri9, size_t n, uint32_t first_prim)
{
	char *d = (char*)blob;
	size_t size = orc_flat_blob_size(n);
	memset(d, 0, size);
	size_t leaf_off = 320, vert_off = 320 + flat_leaf_bytes(n);
	write_header((ref_scene*)d, size, 128, 256, vert_off);

	ref_node4 *root = (ref_node4*)(d + 128);
	for (int i = 0; i < 4; i++) {
		/* empty slot convention, rtk.c:1613-1619 (with the leaf tag the reference forgets) */
		root->bx[0][i] = root->by[0][i] = root->bz[0][i] = +1.0f;
		root->bx[1][i] = root->by[1][i] = root->bz[1][i] = -1.0f;
		root->ptr[i] = 256 | 1;
	}
	root->bx[0][0] = root->by[0][0] = root->bz[0][0] = -1e30f;
	root->bx[1][0] = root->by[1][0] = root->bz[1][0] = +1e30f;
	root->ptr[0] = leaf_off | 1;

	/* leaf record, rtk.c:186-193 */
	uint64_t info = (uint64_t)n | (uint64_t)vert_off;
	memcpy(d + leaf_off, &info, 8);
	ref_leaf_tri *lt = (ref_leaf_tri*)(d + leaf_off + 8);
	ref_vertex *vt = (ref_vertex*)(d + vert_off);
	for (size_t i = 0; i < n; i++) {
		for (int k = 0; k < 3; k++) {
			lt[i].v[k] = (uint8_t)(3 * i + k);
			memcpy(vt[3 * i + k].pos, tri9 + 9 * i + 3 * k, 12);
			vt[3 * i + k].index = (uint32_t)(3 * (first_prim + i) + k);
		}
		lt[i].local_mesh = 0;
		lt[i].tri = first_prim + (uint32_t)i;
	}
	uint32_t *mesh_table = (uint32_t*)(lt + ((n + 3) & ~(size_t)3));
	mesh_table[0] = 0;
}

typedef struct {
	orc_ref_trace_fn fn; const float *tri9; size_t ntris;
	const orc_ray *rays; orc_hit *out; size_t begin, end;
} flat_job;

static void *flat_worker(void *arg)
{
	flat_job *j = (flat_job*)arg;
	const size_t CH = 60;
	size_t nchunks = (j->ntris + CH - 1) / CH;
	size_t bsz = orc_flat_blob_size(CH);
	/* all chunk blobs are prepared once per worker */
	char *blobs = NULL;
	if (posix_memalign((void**)&blobs, 64, bsz * (nchunks ? nchunks : 1))) return NULL;
	for (size_t c = 0; c < nchunks; c++) {
		size_t n = j->ntris - c * CH; if (n > CH) n = CH;
		orc_write_flat_blob(blobs + c * bsz, j->tri9 + 9 * c * CH, n, (uint32_t)(c * CH));
	}
	for (size_t r = j->begin; r < j->end; r++) {
		orc_ray ray = j->rays[r];
		orc_hit h = { 0.0f, 0.0f, 0.0f, 0xffffffffu };
		for (size_t c = 0; c < nchunks; c++) {
			ref_hit rh;
			if (j->fn(blobs + c * bsz, &ray, &rh)) {
				/* chain: later chunks must beat this hit strictly (rtk.c:371) */
				ray.max_t = rh.t;
				h.t = rh.t; h.u = rh.u; h.v = rh.v; h.prim = rh.triangle_index;
			}
		}
		j->out[r] = h;
	}
	free(blobs);
	return NULL;
}

void orc_trace_flat_reference(orc_ref_trace_fn fn, const float *tri9, size_t ntris,
                              const orc_ray *rays, size_t nrays, orc_hit *out, int threads)
{
	threads = pick_threads(threads, nrays);
	flat_job jobs[256];
	for (int i = 0; i < threads; i++) {
		jobs[i].fn = fn; jobs[i].tri9 = tri9; jobs[i].ntris = ntris; jobs[i].rays = rays; jobs[i].out = out;
		jobs[i].begin = nrays * (size_t)i / (size_t)threads;
		jobs[i].end = nrays * (size_t)(i + 1) / (size_t)threads;
	}
	run_jobs(flat_worker, jobs, sizeof(flat_job), threads);
}

/* ---------------------------------------------------------------------- */
/* Binned-SAH build restated: rtk.c:765-1019, 1421-1453                    */
/* ---------------------------------------------------------------------- */

#define ORC_MAX_DEPTH      64   /* RTK_BVH_MAX_DEPTH, rtk.c:5 */
#define ORC_LEAF_MIN_ITEMS 4    /* rtk.c:6 */
#define ORC_LEAF_MAX_ITEMS 63   /* rtk.c:7 says 64 but the leaf header has 6 bits (rtk.c:188) */
#define ORC_SPLITS         32   /* RTK_BUILD_SPLITS, rtk.c:587 */

typedef struct { float mn[3], mx[3]; uint32_t prim; } b_item;
typedef struct {
	float mn[3], mx[3];
	size_t begin, count;
	size_t child;          /* index of first of two children, SIZE_MAX = leaf */
	uint16_t depth;
} b_node;

typedef struct {
	b_item *items; b_node *nodes; size_t num_nodes, cap_nodes;
	float item_cost, split_cost;
	uint32_t max_depth;
} b_ctx;

static float b_area(const float *mn, const float *mx)
{
	/* rtk.c:729-733 */
	float x = mx[0] - mn[0], y = mx[1] - mn[1], z = mx[2] - mn[2];
	return 2.0f * (x * y + y * z + z * x);
}

static void b_reset(float *mn, float *mx)
{
	for (int a = 0; a < 3; a++) { mn[a] = +ORC_INF; mx[a] = -ORC_INF; }
}
static void b_add(float *mn, float *mx, const float *amn, const float *amx)
{
	for (int a = 0; a < 3; a++) { mn[a] = sse_min(mn[a], amn[a]); mx[a] = sse_max(mx[a], amx[a]); }
}

static size_t b_alloc2(b_ctx *c)
{
	if (c->num_nodes + 2 > c->cap_nodes) {
		c->cap_nodes = c->cap_nodes * 2 + 64;
		c->nodes = (b_node*)realloc(c->nodes, c->cap_nodes * sizeof(b_node));
	}
	size_t i = c->num_nodes;
	c->num_nodes += 2;
	return i;
}

static int g_sort_axis;
static int cmp_axis(const void *a, const void *b)
{
	/* rtk.c:739-755 */
	const b_item *ia = (const b_item*)a, *ib = (const b_item*)b;
	float ma = ia->mn[g_sort_axis] + ia->mx[g_sort_axis];
	float mb = ib->mn[g_sort_axis] + ib->mx[g_sort_axis];
	if (ma < mb) return -1;
	if (ma > mb) return +1;
	return ia->prim < ib->prim ? -1 : ia->prim > ib->prim;
}

static int largest_axis(const b_node *n)
{
	/* rtk.c:773-775 computes max-max (defect D10); the intent is the extent */
	float sx = n->mx[0] - n->mn[0], sy = n->mx[1] - n->mn[1], sz = n->mx[2] - n->mn[2];
	float m = sse_max(sse_max(sx, sy), sz);
	return sx == m ? 0 : sy == m ? 1 : 2;
}

static void b_leaf(b_ctx *c, size_t ni)
{
	/* rtk.c:765-811: sort the leaf's items along the largest axis */
	b_node *n = &c->nodes[ni];
	n->child = SIZE_MAX;
	g_sort_axis = largest_axis(n);
	qsort(c->items + n->begin, n->count, sizeof(b_item), cmp_axis);
	if (n->depth > c->max_depth) c->max_depth = n->depth;
}

static void b_build(b_ctx *c, size_t ni);

static void b_make_children(b_ctx *c, size_t ni, size_t num_left,
                            const float *lmn, const float *lmx, const float *rmn, const float *rmx)
{
	size_t ci = b_alloc2(c);
	b_node *n = &c->nodes[ni];
	n->child = ci;
	b_node *ch = &c->nodes[ci];
	memcpy(ch[0].mn, lmn, 12); memcpy(ch[0].mx, lmx, 12);
	memcpy(ch[1].mn, rmn, 12); memcpy(ch[1].mx, rmx, 12);
	ch[0].begin = n->begin;            ch[0].count = num_left;
	ch[1].begin = n->begin + num_left; ch[1].count = n->count - num_left;
	ch[0].child = ch[1].child = SIZE_MAX;
	ch[0].depth = ch[1].depth = (uint16_t)(n->depth + 1);
	b_build(c, ci);
	b_build(c, ci + 1);
}

static void b_equal(b_ctx *c, size_t ni)
{
	/* rtk.c:813-865 */
	b_node *n = &c->nodes[ni];
	g_sort_axis = largest_axis(n);
	b_item *items = c->items + n->begin;
	qsort(items, n->count, sizeof(b_item), cmp_axis);
	size_t nl = n->count / 2;
	float lmn[3], lmx[3], rmn[3], rmx[3];
	b_reset(lmn, lmx); b_reset(rmn, rmx);
	for (size_t i = 0; i < nl; i++) b_add(lmn, lmx, items[i].mn, items[i].mx);
	for (size_t i = nl; i < n->count; i++) b_add(rmn, rmx, items[i].mn, items[i].mx);
	b_make_children(c, ni, nl, lmn, lmx, rmn, rmx);
}

static int b_bucket(const b_item *it, int axis, float min_2x, float rcp_scale_2x)
{
	/* rtk.c:899-902.  (int) of a NaN / out-of-range float is what cvttss2si
	 * gives on x86 (INT_MIN), which the clamp turns into 0. */
	float mid_2x = it->mn[axis] + it->mx[axis];
	float f = (mid_2x - min_2x) * rcp_scale_2x;
	if (!(f >= 0.0f)) return 0;
	if (f >= (float)ORC_SPLITS) return ORC_SPLITS - 1;
	return (int)f;
}

static void b_sah(b_ctx *c, size_t ni)
{
	/* rtk.c:867-1019 */
	b_node *n = &c->nodes[ni];
	b_item *items = c->items + n->begin;
	struct { float mn[3], mx[3], rmn[3], rmx[3]; uint32_t num; } bk[ORC_SPLITS];

	float best_cost = ORC_INF;
	int best_axis = -1, best_bucket = 0;
	float best_l[2][3], best_r[2][3];
	float rcp_parent_area = 1.0f / b_area(n->mn, n->mx);          /* rtk.c:880 */

	for (int axis = 0; axis < 3; axis++) {
		for (int i = 0; i < ORC_SPLITS; i++) { b_reset(bk[i].mn, bk[i].mx); bk[i].num = 0; }
		float mn = n->mn[axis], mx = n->mx[axis];
		float min_2x = mn + mn;
		float rcp_scale_2x = (0.5f * (float)ORC_SPLITS) / (mx - mn);   /* rtk.c:893 */
		for (size_t i = 0; i < n->count; i++) {
			int b = b_bucket(&items[i], axis, min_2x, rcp_scale_2x);
			b_add(bk[b].mn, bk[b].mx, items[i].mn, items[i].mx);
			bk[b].num++;
		}
		/* suffix bounds, rtk.c:910-915 */
		memcpy(bk[ORC_SPLITS - 1].rmn, bk[ORC_SPLITS - 1].mn, 12);
		memcpy(bk[ORC_SPLITS - 1].rmx, bk[ORC_SPLITS - 1].mx, 12);
		for (int i = ORC_SPLITS - 1; i > 0; i--) {
			for (int a = 0; a < 3; a++) {
				bk[i - 1].rmn[a] = sse_min(bk[i - 1].mn[a], bk[i].rmn[a]);
				bk[i - 1].rmx[a] = sse_max(bk[i - 1].mx[a], bk[i].rmx[a]);
			}
		}
		/* forward sweep, rtk.c:918-945 */
		float lmn[3], lmx[3];
		b_reset(lmn, lmx);
		size_t num_left = 0;
		for (int i = 0; i < ORC_SPLITS - 1; i++) {
			b_add(lmn, lmx, bk[i].mn, bk[i].mx);
			num_left += bk[i].num;
			size_t num_right = n->count - num_left;
			if (num_left == 0 || num_right == 0) continue;
			float area_l = b_area(lmn, lmx);
			float area_r = b_area(bk[i + 1].rmn, bk[i + 1].rmx);
			float cost_l = (float)((num_left + 3) / 4) * c->item_cost;     /* rtk.c:934 */
			float cost_r = (float)((num_right + 3) / 4) * c->item_cost;    /* rtk.c:935 */
			float cost = c->split_cost + (area_l * cost_l + area_r * cost_r) * rcp_parent_area;
			if (cost < best_cost) {
				memcpy(best_l[0], lmn, 12); memcpy(best_l[1], lmx, 12);
				memcpy(best_r[0], bk[i + 1].rmn, 12); memcpy(best_r[1], bk[i + 1].rmx, 12);
				best_cost = cost; best_axis = axis; best_bucket = i;
			}
		}
	}

	float leaf_cost = (float)n->count * c->item_cost;                      /* rtk.c:948 */
	if (best_cost < leaf_cost || n->count > ORC_LEAF_MAX_ITEMS) {
		if (best_axis < 0) {
			/* rtk.c:952-959 has the two branches the wrong way round
			 * (it would assert); the stated intent is an equal split */
			if (n->count > ORC_LEAF_MAX_ITEMS) b_equal(c, ni); else b_leaf(c, ni);
			return;
		}
		float mn = n->mn[best_axis], mx = n->mx[best_axis];
		float min_2x = mn + mn;
		float rcp_scale_2x = (0.5f * (float)ORC_SPLITS) / (mx - mn);
		/* in-place partition, rtk.c:968-986 */
		b_item *first = items, *last = items + n->count;
		while (first != last) {
			if (b_bucket(first, best_axis, min_2x, rcp_scale_2x) <= best_bucket) {
				first++;
			} else {
				last--;
				b_item tmp = *first; *first = *last; *last = tmp;
			}
		}
		size_t num_left = (size_t)(first - items);
		b_make_children(c, ni, num_left, best_l[0], best_l[1], best_r[0], best_r[1]);
	} else {
		b_leaf(c, ni);
	}
}

static void b_build(b_ctx *c, size_t ni)
{
	/* rtk.c:1421-1453 */
	b_node *n = &c->nodes[ni];
	if (n->depth == ORC_MAX_DEPTH) { b_leaf(c, ni); return; }
	uint64_t splits_left = ORC_MAX_DEPTH - n->depth - 1;
	if (splits_left > 63) splits_left = 63;
	uint64_t split_items = (uint64_t)n->count >> splits_left;
	if (split_items > ORC_LEAF_MAX_ITEMS) { b_equal(c, ni); return; }
	if (n->count <= ORC_LEAF_MIN_ITEMS) { b_leaf(c, ni); return; }
	b_sah(c, ni);
}

/* ---- packing: rtk.c:1509-1622, 1732-1765 ------------------------------ */

typedef struct { char *p; size_t size, cap; } bytebuf;
static void *bb_grow(bytebuf *b, size_t n)
{
	if (b->size + n > b->cap) {
		b->cap = (b->size + n) * 2 + 4096;
		b->p = (char*)realloc(b->p, b->cap);
	}
	void *r = b->p + b->size;
	memset(r, 0, n);
	b->size += n;
	return r;
}

void *orc_build_reference_blob(const float *tri9, size_t ntris, const uint32_t *mesh_of,
                               const uint32_t *tri_in_mesh, const uint32_t *vidx3,
                               orc_blob_stats *stats)
{
	double t0 = orc_now();
	b_ctx c;
	memset(&c, 0, sizeof(c));
	c.item_cost = 1.0f; c.split_cost = 1.0f;          /* unset in the reference (defect D2) */
	c.items = (b_item*)malloc(sizeof(b_item) * (ntris ? ntris : 1));
	c.cap_nodes = ntris / 2 + 64;
	c.nodes = (b_node*)malloc(sizeof(b_node) * c.cap_nodes);
	c.num_nodes = 1;

	/* per-item bounds and scene bounds, rtk.c:1150-1171, 1398-1404 */
	b_node *root = &c.nodes[0];
	b_reset(root->mn, root->mx);
	for (size_t i = 0; i < ntris; i++) {
		const float *t = tri9 + 9 * i;
		b_item *it = &c.items[i];
		for (int a = 0; a < 3; a++) {
			it->mn[a] = sse_min(sse_min(t[a], t[3 + a]), t[6 + a]);
			it->mx[a] = sse_max(sse_max(t[a], t[3 + a]), t[6 + a]);
		}
		it->prim = (uint32_t)i;
		b_add(root->mn, root->mx, it->mn, it->mx);
	}
	root->begin = 0; root->count = ntris; root->child = SIZE_MAX; root->depth = 0;
	b_build(&c, 0);

	/* a leaf root gets a virtual parent, rtk.c:1460-1476 (defect D9 fixed) */
	if (c.nodes[0].child == SIZE_MAX) {
		size_t ci = b_alloc2(&c);
		c.nodes[ci] = c.nodes[0];
		c.nodes[ci].depth = 1;
		c.nodes[ci + 1] = c.nodes[0];
		c.nodes[ci + 1].count = 0;
		c.nodes[ci + 1].depth = 1;
		c.nodes[0].child = ci;
	}

	/* breadth-first 2->4 collapse, rtk.c:1570-1622 */
	bytebuf nodes = {0}, leaves = {0}, verts = {0};
	size_t *queue = (size_t*)malloc(sizeof(size_t) * (c.num_nodes + 4));
	size_t qh = 0, qt = 0;
	queue[qt++] = 0;
	bb_grow(&nodes, 128);
	bb_grow(&leaves, 64);                    /* null leaf: info 0, rtk.c:1763-1765 */
	size_t num_leaves = 0;

	/* pointers are stored relative for now: nodes as (index<<1), leaves as
	 * (rel<<1)|1, and fixed up once the section offsets are known */
	while (qh < qt) {
		size_t dst_i = qh;
		const b_node *src = &c.nodes[queue[qh++]];
		float bx[2][4], by[2][4], bz[2][4]; uint64_t ptr[4];
		for (unsigned i = 0; i < 4; i++) {
			const b_node *mid = &c.nodes[src->child + (i >> 1)];
			const b_node *ch;
			if (mid->child != SIZE_MAX) ch = &c.nodes[mid->child + (i & 1)];
			else ch = (i & 1) == 0 ? mid : NULL;
			if (ch && ch->count == 0) ch = NULL;
			if (!ch) {
				bx[0][i] = by[0][i] = bz[0][i] = +1.0f;
				bx[1][i] = by[1][i] = bz[1][i] = -1.0f;
				ptr[i] = 1;                  /* null leaf at rel 0 */
				continue;
			}
			bx[0][i] = ch->mn[0]; bx[1][i] = ch->mx[0];
			by[0][i] = ch->mn[1]; by[1][i] = ch->mx[1];
			bz[0][i] = ch->mn[2]; bz[1][i] = ch->mx[2];
			if (ch->child != SIZE_MAX) {
				ptr[i] = (uint64_t)qt << 1;
				queue[qt++] = (size_t)(ch - c.nodes);
				bb_grow(&nodes, 128);
			} else {
				/* leaf record, rtk.c:186-193, 1509-1568 */
				size_t n = ch->count;
				size_t n4 = (n + 3) & ~(size_t)3;
				uint32_t meshes[ORC_LEAF_MAX_ITEMS + 1]; uint32_t nm = 0;
				size_t rel = leaves.size;
				size_t vrel = verts.size;
				const b_item *its = c.items + ch->begin;
				ref_vertex *vt = (ref_vertex*)bb_grow(&verts, align_up(48 * n, 64));
				ref_leaf_tri lt[ORC_LEAF_MAX_ITEMS + 4];
				memset(lt, 0, sizeof(lt));
				for (size_t k = 0; k < n; k++) {
					uint32_t prim = its[k].prim;
					uint32_t mesh = mesh_of ? mesh_of[prim] : 0;
					uint32_t lm = 0;
					while (lm < nm && meshes[lm] != mesh) lm++;
					if (lm == nm) meshes[nm++] = mesh;
					for (int vv = 0; vv < 3; vv++) {
						lt[k].v[vv] = (uint8_t)(3 * k + vv);
						memcpy(vt[3 * k + vv].pos, tri9 + 9 * (size_t)prim + 3 * vv, 12);
						vt[3 * k + vv].index = vidx3 ? vidx3[3 * (size_t)prim + vv] : 3 * prim + vv;
					}
					lt[k].local_mesh = (uint8_t)lm;
					lt[k].tri = tri_in_mesh ? tri_in_mesh[prim] : prim;
				}
				size_t bytes = align_up(8 + 8 * n4 + 4 * nm, 64);
				char *lp = (char*)bb_grow(&leaves, bytes);
				uint64_t info = (uint64_t)n | ((uint64_t)vrel << 8);     /* vrel patched below */
				memcpy(lp, &info, 8);
				memcpy(lp + 8, lt, 8 * n4);
				memcpy(lp + 8 + 8 * n4, meshes, 4 * nm);
				ptr[i] = ((uint64_t)rel << 1) | 1;
				num_leaves++;
			}
		}
		ref_node4 *dst = (ref_node4*)(nodes.p + 128 * dst_i);
		memcpy(dst->bx, bx, sizeof(bx)); memcpy(dst->by, by, sizeof(by)); memcpy(dst->bz, bz, sizeof(bz));
		memcpy(dst->ptr, ptr, sizeof(ptr));
	}

	size_t node_off = 128;
	size_t leaf_off = align_up(node_off + nodes.size, 128);
	size_t vert_off = align_up(leaf_off + leaves.size, 128);
	size_t total = align_up(vert_off + verts.size + 64, 128);
	char *blob = NULL;
	if (posix_memalign((void**)&blob, 64, total)) blob = NULL;
	if (blob) {
		memset(blob, 0, total);
		write_header((ref_scene*)blob, total, node_off, leaf_off, vert_off);
		memcpy(blob + node_off, nodes.p, nodes.size);
		memcpy(blob + leaf_off, leaves.p, leaves.size);
		memcpy(blob + vert_off, verts.p, verts.size);
		size_t nn = nodes.size / 128;
		for (size_t i = 0; i < nn; i++) {
			ref_node4 *nd = (ref_node4*)(blob + node_off + 128 * i);
			for (int k = 0; k < 4; k++) {
				uint64_t p = nd->ptr[k];
				if (p & 1) {
					size_t rel = (size_t)(p >> 1);
					nd->ptr[k] = (uint64_t)(leaf_off + rel) | 1;
					if (rel != 0) {
						uint64_t info;
						memcpy(&info, blob + leaf_off + rel, 8);
						if ((info >> 8) != 0 || (info & 0x3f) != 0) {
							uint64_t n = info & 0x3f, vrel = info >> 8;
							info = n | (uint64_t)(vert_off + vrel);   /* rtk.c:193: 64-B aligned offset | count */
							memcpy(blob + leaf_off + rel, &info, 8);
						}
					}
				} else {
					nd->ptr[k] = (uint64_t)(node_off + 128 * (size_t)(p >> 1));
				}
			}
		}
	}
	if (stats) {
		stats->size_bytes = total; stats->num_nodes4 = nodes.size / 128; stats->num_leaves = num_leaves;
		stats->num_build_nodes = c.num_nodes; stats->max_depth = c.max_depth;
		stats->build_seconds = orc_now() - t0;
	}
	free(queue); free(nodes.p); free(leaves.p); free(verts.p); free(c.items); free(c.nodes);
	return blob;
}

/* ---------------------------------------------------------------------- */
/* Timed traversal through a reference build                               */
/* ---------------------------------------------------------------------- */

typedef struct {
	orc_ref_trace_fn fn; const void *blob; const orc_ray *rays; orc_hit *out;
	const uint32_t *mesh_first; size_t begin, end;
} ref_job;

static void *ref_worker(void *arg)
{
	ref_job *j = (ref_job*)arg;
	for (size_t r = j->begin; r < j->end; r++) {
		ref_hit rh;
		orc_hit h = { 0.0f, 0.0f, 0.0f, 0xffffffffu };
		if (j->fn(j->blob, &j->rays[r], &rh)) {
			h.t = rh.t; h.u = rh.u; h.v = rh.v;
			h.prim = (j->mesh_first ? j->mesh_first[rh.mesh_index] : 0) + rh.triangle_index;
		}
		if (j->out) j->out[r] = h;
	}
	return NULL;
}

double orc_trace_reference_blob(orc_ref_trace_fn fn, const void *blob, const orc_ray *rays,
                                size_t nrays, orc_hit *out, const uint32_t *mesh_first, int threads)
{
	threads = pick_threads(threads, nrays);
	ref_job jobs[256];
	for (int i = 0; i < threads; i++) {
		jobs[i].fn = fn; jobs[i].blob = blob; jobs[i].rays = rays; jobs[i].out = out;
		jobs[i].mesh_first = mesh_first;
		jobs[i].begin = nrays * (size_t)i / (size_t)threads;
		jobs[i].end = nrays * (size_t)(i + 1) / (size_t)threads;
	}
	double t0 = orc_now();
	run_jobs(ref_worker, jobs, sizeof(ref_job), threads);
	return orc_now() - t0;
}
