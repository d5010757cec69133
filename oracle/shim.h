/*
 * shim.h -- force-included ahead of the UNMODIFIED reference rtk.c so that it
 * compiles with gcc.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference supplies these pieces only under _MSC_VER (rtk.c:47-58 atomics,
 * rtk.c:170-175 bit-scan / popcount / alignment) and its GCC allocator branch
 * passes aligned_alloc its arguments swapped (rtk.c:38); nothing here changes
 * the arithmetic of the traced path.
 */
#ifndef ORC_SHIM_H
#define ORC_SHIM_H
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <smmintrin.h>   /* SSE4.1: _mm_blendv_ps used at rtk.c:165 */
#include <nmmintrin.h>

typedef size_t _rtk_atomic_size;
#define _rtk_atomic_size_add(a, v) ((size_t)__atomic_fetch_add((a), (size_t)(v), __ATOMIC_SEQ_CST))
#define RTK_FIRSTBIT4(index, mask) ((index) = (uint32_t)__builtin_ctz((unsigned)(mask)))
#define RTK_POPCOUNT4(mask) ((uint32_t)__builtin_popcount((unsigned)(mask)))
#define RTK_ALIGN16 __attribute__((aligned(16)))

static inline void *orc_shim_alloc(size_t size)
{
	void *p = NULL;
	return posix_memalign(&p, 64, size) == 0 ? p : NULL;
}
#define rtk_alloc 1                       /* closes the guard at rtk.c:32 */
#define rtk_mem_alloc(size) orc_shim_alloc(size)
#define rtk_mem_free(ptr, size) free(ptr)
#endif
