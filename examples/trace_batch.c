/*
 * trace_batch.c -- what a reference user's program looks like after the switch.
 *
 * Before (reference, rtk.h:119-129): rtk_build_scene(&desc), then a loop over rtk_trace_ray().
 * After: the same rtk_build_scene(&desc); ONE rtk_trace_rays() for the whole batch, split by the
 * library over every GPU named in rtk_cuda_init_devices().
 *
 *   cc examples/trace_batch.c -Iinclude -Lrtk_b200 -lrtk_b200 -Wl,-rpath,$PWD/rtk_b200 -o trace_batch
 *   ./trace_batch [number of GPUs]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rtk.h"
#include "rtk_cuda.h"

int main(int argc, char **argv)
{
	int ndev = argc > 1 ? atoi(argv[1]) : 1;
	int devices[16];
	if (ndev < 1 || ndev > 16) ndev = 1;
	for (int i = 0; i < ndev; i++) devices[i] = i;
	if (rtk_cuda_init_devices(devices, ndev) != RTK_CUDA_OK) {          /* no CUDA device: there is no CPU fallback */
		fprintf(stderr, "rtk_b200: %s\n", rtk_cuda_last_error());
		return 1;
	}

	/* a 64 x 64 grid of quads in the plane z = 1: 8192 triangles, indexed, U32 */
	enum { G = 64 };
	const size_t nv = (G + 1) * (G + 1), nt = 2 * G * G;
	float *pos = malloc(sizeof(float) * 3 * nv);
	uint32_t *idx = malloc(sizeof(uint32_t) * 3 * nt);
	for (int y = 0; y <= G; y++) for (int x = 0; x <= G; x++) {
		float *p = pos + 3 * (y * (G + 1) + x);
		p[0] = (float)x / G; p[1] = (float)y / G; p[2] = 1.0f;
	}
	for (int y = 0, t = 0; y < G; y++) for (int x = 0; x < G; x++) {
		uint32_t a = (uint32_t)(y * (G + 1) + x), b = a + 1, c = a + (G + 1), d = c + 1;
		idx[t++] = a; idx[t++] = b; idx[t++] = d;
		idx[t++] = a; idx[t++] = d; idx[t++] = c;
	}
	rtk_mesh mesh;
	memset(&mesh, 0, sizeof(mesh));
	mesh.num_triangles = nt;
	mesh.position.data = pos; mesh.position.type = RTK_TYPE_F32;
	mesh.index.data = idx; mesh.index.type = RTK_TYPE_U32;
	rtk_scene_desc desc;
	memset(&desc, 0, sizeof(desc));
	desc.meshes = &mesh; desc.num_meshes = 1;
	rtk_scene *scene = rtk_build_scene(&desc);                              /* upload + build + replicate to every GPU */
	if (!scene) { fprintf(stderr, "build failed: %s\n", rtk_cuda_last_error()); return 1; }

	/* one million rays straight down +z; page-locked arrays let the devices write the rows in place */
	const size_t n = 1u << 20;
	rtk_ray *rays = rtk_cuda_host_alloc_batch(sizeof(rtk_ray), n);
	rtk_hit *hits = rtk_cuda_host_alloc_batch(sizeof(rtk_hit), n);
	uint8_t *mask = rtk_cuda_host_alloc_batch(1, n);
	if (!rays || !hits || !mask) { fprintf(stderr, "allocation failed: %s\n", rtk_cuda_last_error()); return 1; }
	for (size_t i = 0; i < n; i++) {
		rays[i].origin.x = (float)(i % 1024) / 1024.0f * 1.2f - 0.1f;       /* a tenth of them pass the grid by */
		rays[i].origin.y = (float)(i / 1024) / 1024.0f * 1.2f - 0.1f;
		rays[i].origin.z = 0.0f;
		rays[i].direction.x = 0.0f; rays[i].direction.y = 0.0f; rays[i].direction.z = 1.0f;
		rays[i].min_t = 0.0f; rays[i].max_t = RTK_INF;
	}
	size_t found = rtk_trace_rays(scene, rays, hits, mask, n);             /* rows of rays that missed stay untouched */
	if (found == (size_t)-1) { fprintf(stderr, "trace failed: %s\n", rtk_cuda_last_error()); return 1; }
	printf("%zu of %zu rays hit on %d GPU(s); ray 524800: %s", found, n, rtk_cuda_device_count(), mask[524800] ? "hit" : "miss");
	if (mask[524800]) printf(" t=%g triangle %u", hits[524800].t, hits[524800].triangle_index);
	printf("\n");

	/* the single-ray entry point of the reference still works (a one-ray batch: correct, never fast) */
	rtk_hit one;
	if (rtk_trace_ray(scene, &rays[524800], &one)) printf("rtk_trace_ray agrees: t=%g\n", one.t);

	rtk_cuda_host_free(rays); rtk_cuda_host_free(hits); rtk_cuda_host_free(mask);
	rtk_free_scene(scene);
	rtk_cuda_shutdown();
	free(pos); free(idx);
	return 0;
}
