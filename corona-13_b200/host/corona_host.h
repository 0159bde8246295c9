/* corona_host.h -- the reference's module API for the hot path, served by libcorona_b200.so.
 *
 * Same function names, argument meaning and error behaviour as include/accel.h:28-50 of the
 * reference (no error codes: failures print to stderr; NULL handles are tolerated nowhere, as in
 * the reference).  Two build modes:
 *
 *   in-tree  (-DCORONA_B200_IN_TREE, compiled as src/accel.d/b200.c inside the reference tree):
 *            the reference's own corona_common.h / prims.h / accel.h provide the types.
 *   standalone (this repository): layout-compatible restatements below, so that the parity tests
 *            can drive the very same C code through ctypes.
 */
#ifndef CORONA_HOST_H
#define CORONA_HOST_H

#include <stdint.h>
#include <stdio.h>
#include "corona_types.h"

#ifdef CORONA_B200_IN_TREE
#include "corona_common.h"
#include "prims.h"
#include "accel.h"
#else
/* restated with the reference's names so that host code reads like the reference's callers */
typedef cb_ray_t ray_t;      /* include/corona_common.h:113-121 */
typedef cb_hit_t hit_t;      /* include/corona_common.h:123-137 */

typedef struct prims_shape_t /* include/prims.h:49-65, same field order and sizes */
{
  int64_t  material;
  uint64_t num_prims;
  char     name[1024];
  char     tex[512];
  int      fd;
  void    *data;
  size_t   data_size;
  cb_vtxidx_t *vtxidx;
  cb_vtx_t    *vtx;
  uint64_t    *primid;       /* primid_t[] */
}
prims_shape_t;

typedef struct prims_t       /* include/prims.h:67-83 */
{
  uint32_t num_shapes;
  prims_shape_t *shape;
  uint64_t num_loaded_prims;
  uint32_t num_loaded_shapes;
  uint64_t num_prims;
  uint64_t *primid;          /* global index array, permuted by accel_build */
  float ghost_aabb[6];
}
prims_t;

struct accel_t;
typedef struct accel_t accel_t;

/* include/accel.h:28-50 */
void         accel_print_info(FILE *fd);
accel_t     *accel_init(prims_t *p);
void         accel_cleanup(accel_t *b);
void         accel_build(accel_t *b, const char *filename);
void         accel_intersect(const accel_t *b, const ray_t *ray, hit_t *hit);
int          accel_visible(const accel_t *b, const ray_t *ray, const float max_dist);
void         accel_closest(const accel_t *b, ray_t *ray, hit_t *hit, const float centre);
const float *accel_aabb(const accel_t *b);
#endif

/* batched extensions (one synchronous ray per call cannot feed a GPU).
 * hits[i].dist is the search limit on input like accel_intersect; only prim,u,v,dist (spheres: x)
 * are written, and only when a closer hit is found. */
void accel_intersect_n(const accel_t *b, const ray_t *rays, hit_t *hits, uint64_t n);
void accel_visible_n(const accel_t *b, const ray_t *rays, const float *max_dist, int *visible, uint64_t n);

/* the device handle behind an accel_t (cb200_accel_t*), for the render module */
void *accel_b200_handle(const accel_t *b);

/* ---- MOD_render=b200 (host/render_b200.c): include/render.h:13-32 + the batched progression entry ------------ */
#include "corona_b200_render.h"
struct render_t;
struct render_tls_t;
#ifndef CORONA_B200_IN_TREE
struct render_t *render_init();                    /* the reference's argument-less form needs rt.*: in-tree only */
void render_cleanup(struct render_t *r);
struct render_tls_t *render_tls_init();
void render_tls_cleanup(struct render_tls_t *r);
void render_print_info(FILE *fd);
void render_sample_path(uint64_t index);
void render_clear();
#endif
struct render_t *render_b200_init(const accel_t *accel, const cb_render_desc_t *desc);
/* == for(i in [first_index, first_index+count)) render_sample_path(i); fb: host W*H*3 floats or NULL */
int render_b200_pass(struct render_t *r, uint64_t first_index, uint64_t count, float *fb);
/* finish the paths still in flight and fetch the complete image (before fb_export / screenshots) */
int render_b200_finish(struct render_t *r, float *fb);
uint64_t render_b200_overlays(const struct render_t *r);
/* `--dbor n` of the reference's view (src/view.c:291,339-350,497-522): switch the outlier rejection cascade on (n > 1) before
 * the first progression, fetch level `level` (host W*H*3 floats, un-gained like the framebuffer) after render_b200_finish */
int render_b200_set_dbor(struct render_t *r, int levels);
int render_b200_dbor(struct render_t *r, int level, float *fb);
void *render_b200_handle(const struct render_t *r);
int render_b200_dbor_is_set(struct render_t *r, int set);   /* bookkeeping for the view: has set_dbor been called (and mark it) */

/* ---- scene ingestion (host/scene_b200.c): .nra2 shader + shape lists, .cam, rgb2spec coefficients, measured tables ---- */
struct scene_b200_t;
struct scene_b200_t *scene_b200_open(const char *nra2_file, const char *coeff_file, const char *table_file);   /* no GPU needed */
/* the sky line and the shader list only (no shapes are loaded): for hosts that own the geometry themselves, like the in-tree render module */
struct scene_b200_t *scene_b200_open_shaders(const char *nra2_file, const char *coeff_file, const char *table_file);
/* materials, measured tables, media and sky of s into d (pointers borrowed from s: keep s alive while d is in use) */
void scene_b200_fill_desc(const struct scene_b200_t *s, cb_render_desc_t *d);
void scene_b200_free(struct scene_b200_t *s);
const cb_material_t *scene_b200_materials(const struct scene_b200_t *s, int *num);
/* the homogeneous media the shader list defines (cb_material_t.medium / *exterior are 1 + index, 0 = vacuum) */
const cb_medium_t *scene_b200_media(const struct scene_b200_t *s, int *num, int *exterior);
const char *scene_b200_basename(const struct scene_b200_t *s);
uint64_t scene_b200_num_prims(const struct scene_b200_t *s);
const float *scene_b200_aabb(const struct scene_b200_t *s);          /* accel_aabb of the built scene */
int scene_b200_read_camera(const char *filename, uint32_t width, uint32_t height, cb_camera_t *out);
/* accel_init + accel_build + camera + render_b200_init: what main.c's init() does for the hot path */
int scene_b200_prepare(struct scene_b200_t *s, uint32_t width, uint32_t height, int sampler, int pointsampler, int colour, uint64_t frame,
                       const char *cam_file);
struct render_t *scene_b200_render(struct scene_b200_t *s);
const cb_render_desc_t *scene_b200_desc(const struct scene_b200_t *s);
int scene_b200_write_pfm(const char *filename, const float *fb, uint32_t width, uint32_t height, float gain);

/* prims helpers for standalone use: the slice of prims_init/allocate/load/allocate_index the path needs
 * (src/prims.c:703-828) */
#ifndef CORONA_B200_IN_TREE
void prims_init(prims_t *p);
void prims_allocate(prims_t *p, const uint32_t num_shapes);
int  prims_load(prims_t *p, const char *filename, const char *texture, const int shader);
int  prims_add_shape_mem(prims_t *p, const cb_shape_t *s);   /* in-memory shape instead of a .geo file */
void prims_allocate_index(prims_t *p);
void prims_cleanup(prims_t *p);
#endif

#endif
