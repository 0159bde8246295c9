/* corona_b200.h -- C ABI of libcorona_b200.so: corona-13's path-tracing hot path on B200 (sm_100a).
 *
 * Plain pointers and sizes only.  Each entry point names the reference interface it replaces
 * (file:line, relative to the reference root).  All functions return 0 on success, a negative
 * cb200 error code otherwise (pointer-returning ones return NULL) and leave a message in
 * cb200_last_error().  There is NO CPU fallback: without a CUDA device every compute entry fails
 * with CB200_ERR_NO_DEVICE.
 *
 * Threading: a scene/accel may be traversed from many host threads concurrently (like
 * accel_intersect, accel.h:40); build/import/destroy are single-threaded (like accel_build).
 */
#ifndef CORONA_B200_H
#define CORONA_B200_H

#include <stddef.h>
#include "corona_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_OK              0
#define CB200_ERR_NO_DEVICE  -1
#define CB200_ERR_CUDA       -2
#define CB200_ERR_ARG        -3
#define CB200_ERR_NOMEM      -4
#define CB200_ERR_UNSUPPORTED -5

typedef struct cb200_scene cb200_scene_t;
typedef struct cb200_accel cb200_accel_t;

/* ---- runtime ---------------------------------------------------------------------------- */
const char *cb200_version(void);
const char *cb200_last_error(void);
int   cb200_device_count(void);
int   cb200_set_device(int device);           /* one process per GPU: call once with LOCAL_RANK */
int   cb200_sm_count(void);
/* device buffers for callers without a CUDA runtime of their own (the C host layer) */
void *cb200_malloc(size_t bytes);
int   cb200_free(void *dptr);
void *cb200_malloc_host(size_t bytes);        /* pinned host memory */
int   cb200_free_host(void *hptr);
int   cb200_memcpy_h2d(void *dptr, const void *hptr, size_t bytes, void *stream);
int   cb200_memcpy_d2h(void *hptr, const void *dptr, size_t bytes, void *stream);
int   cb200_stream_sync(void *stream);        /* stream == NULL: legacy default stream */

/* ---- geometry: replaces the prims_t assembly done by prims_allocate / prims_load /
 *      prims_allocate_index (src/prims.c:718-828).  The shape arrays are copied into HBM;
 *      the caller keeps ownership of its (mmapped) host copies.                              */
cb200_scene_t *cb200_scene_create(const cb_shape_t *shapes, int num_shapes);
void           cb200_scene_destroy(cb200_scene_t *s);
uint64_t       cb200_scene_num_prims(const cb200_scene_t *s);

/* ---- acceleration structure: replaces accel_init + accel_build (include/accel.h:31,37;
 *      src/accel.d/qbvhmp.c:285-323,1181-1186).  Built entirely on the GPU: prim boxes ->
 *      63-bit Morton codes -> radix sort -> binary radix tree -> refit (shutter open + close
 *      boxes) -> collapse to the reference's 4-wide layout with axis0/axis00/axis01.
 *      Like the reference the build permutes the global primid list; the permuted list is
 *      returned through primid_out (host, num_prims entries, may be NULL).
 *      ghost_aabb (6 floats or NULL) is merged into the scene box (qbvhmp.c:1141-1143).       */
cb200_accel_t *cb200_accel_build(cb200_scene_t *s, const float *ghost_aabb, uint64_t *primid_out);

/* adopt a tree that already is in the reference's node layout (qbvhmp.c:62-81) together with the
 * primid permutation its leaves index into -- parity mode A: the GPU traverses the CPU-built
 * tree in the reference's order and must agree bit for bit, ties included.                     */
cb200_accel_t *cb200_accel_import_qbvh(cb200_scene_t *s, const cb_qbvh_node_t *nodes, uint64_t num_nodes,
                                       const uint64_t *primid_permuted, const float aabb[6]);
void     cb200_accel_destroy(cb200_accel_t *a);
uint64_t cb200_accel_num_nodes(const cb200_accel_t *a);
int      cb200_accel_depth(const cb200_accel_t *a);
/* accel_aabb (accel.h:50) */
int      cb200_accel_aabb(const cb200_accel_t *a, float aabb[6]);
/* export the tree in the reference layout (for checktree-style validation, qbvhmp.c:211-257) */
int      cb200_accel_export_qbvh(const cb200_accel_t *a, cb_qbvh_node_t *nodes, uint64_t cap, uint64_t *primid_out);
/* bytes one node / one primitive record occupy in HBM (roofline accounting) */
int      cb200_accel_layout(const cb200_accel_t *a, uint32_t *node_bytes, uint32_t *prim_bytes);

/* Which tree the traversal entries below walk.
 *   CB200_TRAVERSAL_EXACT4: the 4-wide tree in the reference's node layout, visited in the reference's order
 *     (qbvhmp.c:1262-1490): results are bit-identical to the reference algorithm on that tree, ties included.  Always
 *     available; the only mode for motion-blurred scenes and imported reference trees.
 *   CB200_TRAVERSAL_WIDE8: the 8-wide compressed tree cb200_accel_build also makes for static scenes (same primitive order,
 *     quantised conservative child boxes, octant visiting order): a third of the node bytes per ray; opt-in, because on
 *     B200 the traversal is issue bound, not byte bound, and the 4-wide kernel is the faster one.
 *     Every primitive test is the reference's arithmetic, so prim / u / v / dist are the same bits except where two
 *     primitives lie at exactly the same distance along the ray (the last one TESTED wins in the reference,
 *     geo/triangle.h:296, and the order is the tree's) or a child box is grazed within its last ulp -- the tree-dependent
 *     cases in which the reference's own answer changes with its builder.
 * set_traversal fails with CB200_ERR_UNSUPPORTED when WIDE8 is asked of an accel without that tree; not thread-safe
 * against concurrent traversal calls.                                                                                   */
#define CB200_TRAVERSAL_EXACT4 0
#define CB200_TRAVERSAL_WIDE8  1
int      cb200_accel_set_traversal(cb200_accel_t *a, int mode);
int      cb200_accel_traversal(const cb200_accel_t *a);

/* ---- traversal: batched accel_intersect / accel_visible (accel.h:40,43; qbvhmp.c:1262-1490).
 *      rays: n cb_ray_t;  max_dist: n floats or NULL (= FLT_MAX; the reference presets
 *      hit->dist, pathspace.c:762);  out: n cb_hitrec_t {prim,u,v,dist}, prim == INVALID and
 *      dist == max_dist when nothing was hit.  visible_n writes 1 = unoccluded like accel_visible.
 *      *_n take HOST pointers and include the copies; *_dev take DEVICE pointers and only
 *      enqueue on `stream` (a cudaStream_t, NULL = default stream).
 *      intersect_n / visible_n stage through persistent device buffers in 512 Ki-ray chunks on four streams (upload,
 *      traversal and download of consecutive chunks overlap when the host buffers are pinned); such calls from several host
 *      threads are serialised.  Batches of up to 128 rays -- the batch of one behind accel_intersect / accel_visible of every
 *      pinned worker thread -- take another route: the calling thread's own stream and block of mapped pinned memory (made at its
 *      first call, kept for the life of the thread), one launch + one stream synchronisation, no lock, concurrent across
 *      threads; the calling thread's current device is set to the accel's and left there.                               */
int cb200_accel_intersect_n(const cb200_accel_t *a, const cb_ray_t *rays, const float *max_dist,
                            cb_hitrec_t *out, uint64_t n);
int cb200_accel_visible_n(const cb200_accel_t *a, const cb_ray_t *rays, const float *max_dist,
                          int32_t *out, uint64_t n);
/* batched accel_closest (accel.h:47; qbvhmp.c:1493-1600): the hit nearest to centre[i] along ray i.  rays[i].min_dist and
 * io[i] = {prim,u,v,dist} are in/out exactly like the reference's ray_t* / hit_t* arguments (io[i].dist = search limit in). */
int cb200_accel_closest_n(const cb200_accel_t *a, cb_ray_t *rays, cb_hitrec_t *io, const float *centre, uint64_t n);
int cb200_accel_intersect_dev(const cb200_accel_t *a, const void *d_rays, const void *d_max_dist,
                              void *d_out, uint64_t n, void *stream);
int cb200_accel_visible_dev(const cb200_accel_t *a, const void *d_rays, const void *d_max_dist,
                            void *d_out, uint64_t n, void *stream);
/* instrumented closest-hit pass (device pointers): counters[4] = {rays, node visits with >= 1 child
 * hit, child boxes hit, prim tests}, the reference's ACCEL_DEBUG definitions (qbvhmp.c:83-90) */
int cb200_accel_intersect_counted(const cb200_accel_t *a, const void *d_rays, const void *d_max_dist,
                                  void *d_out, uint64_t n, uint64_t counters[4]);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
uint64_t cb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
