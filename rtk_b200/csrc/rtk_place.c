/*
 * rtk_place.c -- host side of the batched result path: a small pool of worker threads that puts
 * the dense hit rows a chunk brought back from the device where the caller wants them.
 *
 * Why it exists: rtk_hit is 68 bytes (rtk.h:36-43) and the reference leaves *hit untouched on a
 * miss (rtk.c:571-576).  Copying hits[n] back in bulk moves 68 bytes for every ray, hit or not;
 * on an incoherent batch two thirds of that PCIe traffic is rows nobody may look at.  The device
 * therefore packs the rows of the rays that hit (k_resolve, dense mode: 128-ray blocks, each block's
 * rows contiguous from block_base[b]) and only those cross the bus, together with one mask byte per
 * ray.  Nothing is computed here: a row is copied to hits[i] for every i with mask[i] != 0.
 */
#include "rtk_device.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define PLACE_MAX_THREADS 32
#define PLACE_MAX_JOBS 8
#define PLACE_BLOCK 128            /* rays per k_resolve block */
#define PLACE_ROW 68

typedef struct place_job {
	rtkd_place_desc d;
	/* kind 1: a plain parallel copy (pageable rays into the pinned bounce buffer of the upload stream) */
	int kind;
	char *cdst; const char *csrc; size_t cbytes;
	size_t next_slice, num_slices, slice_blocks;
	size_t done_slices;
	int active;
} place_job;
#define COPY_SLICE ((size_t)1 << 18)

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_work = PTHREAD_COND_INITIALIZER;
static pthread_cond_t g_done = PTHREAD_COND_INITIALIZER;
static pthread_t g_threads[PLACE_MAX_THREADS];
static int g_nthreads;
static place_job g_jobs[PLACE_MAX_JOBS];

static void place_slice(const rtkd_place_desc *d, size_t b0, size_t b1)
{
	const size_t nblocks = (d->nrays + PLACE_BLOCK - 1) / PLACE_BLOCK;
	if (b1 > nblocks) b1 = nblocks;
	for (size_t b = b0; b < b1; b++) {
		const unsigned char *row = (const unsigned char*)d->rows + (size_t)PLACE_ROW * d->block_base[b];
		const size_t i0 = b * PLACE_BLOCK, i1 = i0 + PLACE_BLOCK < d->nrays ? i0 + PLACE_BLOCK : d->nrays;
		const unsigned char *m = d->mask + i0;
		unsigned char *dst = (unsigned char*)d->hits + (size_t)PLACE_ROW * (d->first_ray + i0);
		size_t i = i0;
		/* eight mask bytes at a time (the chunk's mask buffer is 16-byte aligned and padded):
		 * words without a hit are skipped, set bytes are found with count-trailing-zeros */
		for (; i + 8 <= i1; i += 8, m += 8, dst += 8 * PLACE_ROW) {
			uint64_t w;
			memcpy(&w, m, 8);
			while (w) {
				const unsigned k = (unsigned)__builtin_ctzll(w) >> 3;
				memcpy(dst + (size_t)PLACE_ROW * k, row, PLACE_ROW);
				row += PLACE_ROW;
				w &= ~((uint64_t)0xff << (8 * k));
			}
		}
		for (; i < i1; i++, m++, dst += PLACE_ROW) {
			if (*m) { memcpy(dst, row, PLACE_ROW); row += PLACE_ROW; }
		}
		if (d->mask_out) memcpy(d->mask_out + d->first_ray + i0, d->mask + i0, i1 - i0);
	}
}

static void *place_worker(void *arg)
{
	(void)arg;
	pthread_mutex_lock(&g_mu);
	for (;;) {
		place_job *job = NULL;
		size_t slice = 0;
		for (int j = 0; j < PLACE_MAX_JOBS && !job; j++) {
			if (g_jobs[j].active && g_jobs[j].next_slice < g_jobs[j].num_slices) { job = &g_jobs[j]; slice = job->next_slice++; }
		}
		if (!job) { pthread_cond_wait(&g_work, &g_mu); continue; }
		pthread_mutex_unlock(&g_mu);
		if (job->kind == 1) {
			const size_t off = slice * COPY_SLICE, len = job->cbytes - off < COPY_SLICE ? job->cbytes - off : COPY_SLICE;
			memcpy(job->cdst + off, job->csrc + off, len);
		} else place_slice(&job->d, slice * job->slice_blocks, (slice + 1) * job->slice_blocks);
		pthread_mutex_lock(&g_mu);
		if (++job->done_slices == job->num_slices) pthread_cond_broadcast(&g_done);
	}
	return NULL;
}

static void pool_start(void)
{
	/* called with g_mu held */
	if (g_nthreads) return;
	int want = 0;
	const char *e = getenv("RTK_B200_HOST_THREADS");
	if (e) want = atoi(e);
	if (want <= 0) {
		long n = sysconf(_SC_NPROCESSORS_ONLN);
		/* measured on a 16-vCPU host: 4 threads 31.6 ms, 8 threads 23.9 ms, 16 threads 22.1 ms per
		 * 16.7M-ray batch (the copies are bound by host memory bandwidth, which the PCIe traffic
		 * shares) -- three quarters of the CPUs, at most 16 */
		want = n >= 4 ? (int)(n * 3 / 4) : 1;
		if (want > 16) want = 16;
		/* one process per GPU: the ranks of a node share its CPUs (torchrun exports LOCAL_WORLD_SIZE) */
		const char *lw = getenv("LOCAL_WORLD_SIZE");
		if (lw && atoi(lw) > 1) want = want / atoi(lw) > 0 ? want / atoi(lw) : 1;
	}
	if (want > PLACE_MAX_THREADS) want = PLACE_MAX_THREADS;
	for (int i = 0; i < want; i++) {
		if (pthread_create(&g_threads[g_nthreads], NULL, place_worker, NULL) == 0) {
			pthread_detach(g_threads[g_nthreads]);
			g_nthreads++;
		}
	}
}

int rtkd_place_threads(void)
{
	pthread_mutex_lock(&g_mu);
	pool_start();
	int n = g_nthreads;
	pthread_mutex_unlock(&g_mu);
	return n;
}

int rtkd_place_submit(const rtkd_place_desc *d)
{
	static int skip = -1;             /* experiment knob: measure the pipeline without the placement */
	if (skip < 0) { const char *e = getenv("RTK_B200_SKIP_PLACE"); skip = e && atoi(e) != 0; }
	if (skip) return -1;
	pthread_mutex_lock(&g_mu);
	pool_start();
	int ticket = -1;
	for (int j = 0; j < PLACE_MAX_JOBS; j++) if (!g_jobs[j].active) { ticket = j; break; }
	if (ticket < 0 || g_nthreads == 0) {
		/* no free slot or no worker could be started: do it here */
		pthread_mutex_unlock(&g_mu);
		place_slice(d, 0, (d->nrays + PLACE_BLOCK - 1) / PLACE_BLOCK);
		return -1;
	}
	place_job *job = &g_jobs[ticket];
	job->d = *d;
	job->kind = 0;
	size_t nblocks = (d->nrays + PLACE_BLOCK - 1) / PLACE_BLOCK;
	job->slice_blocks = 256;                               /* 32 Ki rays per slice */
	job->num_slices = (nblocks + job->slice_blocks - 1) / job->slice_blocks;
	job->next_slice = 0; job->done_slices = 0;
	job->active = 1;
	if (job->num_slices == 0) job->active = 0, ticket = -1;
	pthread_cond_broadcast(&g_work);
	pthread_mutex_unlock(&g_mu);
	return ticket;
}

/* parallel memcpy on the same worker pool; returns a ticket for rtkd_place_wait (or -1: done synchronously) */
int rtkd_copy_submit(void *dst, const void *src, size_t bytes)
{
	if (!bytes) return -1;
	pthread_mutex_lock(&g_mu);
	pool_start();
	int ticket = -1;
	for (int j = 0; j < PLACE_MAX_JOBS; j++) if (!g_jobs[j].active) { ticket = j; break; }
	if (ticket < 0 || g_nthreads == 0 || bytes < 2 * COPY_SLICE) {
		pthread_mutex_unlock(&g_mu);
		memcpy(dst, src, bytes);
		return -1;
	}
	place_job *job = &g_jobs[ticket];
	job->kind = 1;
	job->cdst = (char*)dst; job->csrc = (const char*)src; job->cbytes = bytes;
	job->num_slices = (bytes + COPY_SLICE - 1) / COPY_SLICE;
	job->next_slice = 0; job->done_slices = 0;
	job->active = 1;
	pthread_cond_broadcast(&g_work);
	pthread_mutex_unlock(&g_mu);
	return ticket;
}

void rtkd_place_wait(int ticket)
{
	if (ticket < 0 || ticket >= PLACE_MAX_JOBS) return;
	pthread_mutex_lock(&g_mu);
	place_job *job = &g_jobs[ticket];
	while (job->active && job->done_slices < job->num_slices) pthread_cond_wait(&g_done, &g_mu);
	job->active = 0;
	pthread_mutex_unlock(&g_mu);
}
