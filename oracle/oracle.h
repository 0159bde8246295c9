/* oracle.h -- TEST INFRASTRUCTURE: CPU restatement of corona-13's qbvhmp/prims hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 * The product (libcorona_b200.so) never links or loads this.
 *
 * Parity pin: tests/test_oracle_vs_ref.py checks every function here bit-for-bit
 * against oracle/_ref/libcorona_ref.so (the unmodified reference compiled in place)
 * and against the committed fixtures in tests/golden/ that were produced by it.
 */
#ifndef CORONA_ORACLE_H
#define CORONA_ORACLE_H
#include "corona_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene_t orc_scene_t;
typedef struct orc_accel_t orc_accel_t;

/* scene: borrows the shape arrays, owns the global (build-permuted) primid list */
orc_scene_t *orc_scene_new(const cb_shape_t *shapes, int num_shapes);
void         orc_scene_free(orc_scene_t *s);
uint64_t     orc_scene_num_prims(const orc_scene_t *s);
uint64_t    *orc_scene_primid(orc_scene_t *s);

void orc_prim_bounds(const orc_scene_t *s, cb_primid_t pi, int shutter_close, float aabb[6]);
void orc_prim_intersect(const orc_scene_t *s, cb_primid_t pi, const cb_ray_t *ray, cb_hit_t *hit);
int  orc_prim_visible(const orc_scene_t *s, cb_primid_t pi, const cb_ray_t *ray, float max_dist);

/* accel: serial restatement of the binned-SAH 4-wide build (== reference with one thread) */
orc_accel_t *orc_accel_build(orc_scene_t *s);
/* wrap an existing tree (reference-built or downloaded from the GPU builder); nodes are copied */
orc_accel_t *orc_accel_import(orc_scene_t *s, const cb_qbvh_node_t *nodes, uint64_t num_nodes, const float aabb[6]);
void         orc_accel_free(orc_accel_t *a);
uint64_t     orc_accel_num_nodes(const orc_accel_t *a);
const cb_qbvh_node_t *orc_accel_nodes(const orc_accel_t *a);
const float *orc_accel_aabb(const orc_accel_t *a);

/* single ray, reference signature semantics (accel.h:40,43) */
void orc_intersect(const orc_accel_t *a, const cb_ray_t *ray, cb_hit_t *hit, uint64_t counters[4]);
int  orc_visible(const orc_accel_t *a, const cb_ray_t *ray, float max_dist);

/* accel_closest (accel.h:47): mutates ray->min_dist and *hit like the reference */
void orc_closest(const orc_accel_t *a, cb_ray_t *ray, cb_hit_t *hit, float centre);
void orc_closest_n(const orc_accel_t *a, cb_ray_t *rays, cb_hitrec_t *io, const float *centre, uint64_t n);

/* batches; max_dist may be NULL (=FLT_MAX); counters (optional) accumulate
 * {rays, node visits with >=1 child hit, child boxes hit, prim tests} like ACCEL_DEBUG (qbvhmp.c:83-90) */
void orc_intersect_n(const orc_accel_t *a, const cb_ray_t *rays, const float *max_dist, cb_hitrec_t *out,
                     uint64_t n, int nthreads, uint64_t counters[4]);
void orc_visible_n(const orc_accel_t *a, const cb_ray_t *rays, const float *max_dist, int32_t *out,
                   uint64_t n, int nthreads);

/* structural check of a 4-wide tree: every prim referenced exactly once, child boxes contain
 * their prims at t=0 and t=1, axes in range.  returns 0 if ok, else an error code. */
int orc_accel_check(const orc_accel_t *a, uint64_t stats[4]);

#ifdef __cplusplus
}
#endif
#endif
