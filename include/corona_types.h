/* corona_types.h -- plain-old-data records shared by the host layer, the C-ABI
 * and the CPU oracle.
 *
 * These restate the *binary layouts* of the reference's records so that buffers
 * can be handed across the boundary without conversion.  Every record cites
 * the reference definition it is layout-compatible with; static asserts pin the
 * sizes.  Nothing here is copied code: the reference uses C bitfields, we use
 * explicit 64-bit packing so that the same header compiles as C11, C++ and CUDA.
 *
 *   cb_primid_t   <->  primid_t         include/corona_common.h:45-53
 *   cb_ray_t      <->  ray_t            include/corona_common.h:113-121
 *   cb_hit_t      <->  hit_t            include/corona_common.h:123-137
 *   cb_vtx_t      <->  prims_vtx_t      include/prims.h:37-47
 *   cb_vtxidx_t   <->  prims_vtxidx_t   include/prims.h:20-24
 *   cb_geo_header_t <-> prims_header_t  include/prims.h:26-35
 *   cb_qbvh_node_t <-> qbvh_node_t      src/accel.d/qbvhmp.c:62-81 (motion-blur variant, 256 B)
 */
#ifndef CORONA_B200_TYPES_H
#define CORONA_B200_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- packed primitive handle (64 bit) -------------------------------------
 * lo 32 bit:  extra:3 (bits 0-2) | shapeid:29 (bits 3-31)
 * hi 32 bit:  vi:28   (bits 0-27) | mb:1 (bit 28) | vcnt:3 (bits 29-31)
 * vcnt: 1 sphere, 2 line (truncated cone), 3 triangle, 4 quad, 5 shell
 * (include/prims.h:9-18).  All ones == INVALID (corona_common.h:55).          */
typedef uint64_t cb_primid_t;
#define CB_INVALID_PRIMID (~(uint64_t)0)

#define CB_PRIM_SPHERE 1u
#define CB_PRIM_LINE   2u
#define CB_PRIM_TRI    3u
#define CB_PRIM_QUAD   4u
#define CB_PRIM_SHELL  5u

static inline uint32_t cb_primid_extra  (cb_primid_t p) { return (uint32_t)(p      ) & 7u; }
static inline uint32_t cb_primid_shapeid(cb_primid_t p) { return (uint32_t)(p >>  3) & 0x1fffffffu; }
static inline uint32_t cb_primid_vi     (cb_primid_t p) { return (uint32_t)(p >> 32) & 0x0fffffffu; }
static inline uint32_t cb_primid_mb     (cb_primid_t p) { return (uint32_t)(p >> 60) & 1u; }
static inline uint32_t cb_primid_vcnt   (cb_primid_t p) { return (uint32_t)(p >> 61) & 7u; }
static inline cb_primid_t cb_primid_make(uint32_t extra, uint32_t shapeid, uint32_t vi, uint32_t mb, uint32_t vcnt)
{
  return  (uint64_t)(extra & 7u) | ((uint64_t)(shapeid & 0x1fffffffu) << 3)
        | ((uint64_t)(vi & 0x0fffffffu) << 32) | ((uint64_t)(mb & 1u) << 60) | ((uint64_t)(vcnt & 7u) << 61);
}
static inline cb_primid_t cb_primid_with_shapeid(cb_primid_t p, uint32_t shapeid)
{
  return (p & ~((uint64_t)0x1fffffffu << 3)) | ((uint64_t)(shapeid & 0x1fffffffu) << 3);
}

/* ---- ray, 40 bytes -------------------------------------------------------- */
typedef struct cb_ray_t
{
  float pos[3];
  float dir[3];
  float time;          /* in [0,1]: 0 shutter open, 1 shutter close */
  float min_dist;      /* exclusive lower bound for primitive distances */
  uint32_t ignore[2];  /* cb_primid_t split in two words: the struct is 4-byte aligned like the reference's */
}
cb_ray_t;

/* ---- hit record, 100 bytes -------------------------------------------------
 * traversal reads dist (search limit) and writes prim,u,v,dist (spheres: also x)
 * only when a closer hit is found; the rest belongs to shading.               */
typedef struct cb_hit_t
{
  uint32_t prim[2];    /* cb_primid_t */
  float u, v, w;
  float r, s, t;
  float a[3], b[3];
  float n[3];
  float x[3];
  float gn[3];
  int32_t shader;
  float dist;
}
cb_hit_t;

/* compact traversal result used by the batched entry points (24 bytes) */
typedef struct cb_hitrec_t
{
  uint32_t prim[2];
  float u, v, dist;
  uint32_t pad;
}
cb_hitrec_t;

/* ---- geometry storage ------------------------------------------------------ */
typedef struct cb_vtx_t    { float v[3]; uint32_t n; } cb_vtx_t;      /* n: oct-encoded normal, or float radius for sphere/line */
typedef struct cb_vtxidx_t { uint32_t v, uv; }         cb_vtxidx_t;   /* uv: two IEEE halfs */

#define CB_GEO_MAGIC   0xc01337
#define CB_GEO_VERSION 2
typedef struct cb_geo_header_t
{
  int32_t  magic, version;
  uint64_t num_prims;
  uint64_t vtxidx_offset;
  uint64_t vertex_offset;
}
cb_geo_header_t;

/* one shape as seen by the accel: borrowed pointers into caller memory
 * (the reference keeps them pointing into the mmapped .geo, prims.h:49-65).    */
typedef struct cb_shape_t
{
  const cb_primid_t *primid;   /* per-shape list, shapeid field ignored */
  uint64_t           num_prims;
  const cb_vtxidx_t *vtxidx;
  uint64_t           num_vtxidx;
  const cb_vtx_t    *vtx;      /* interleaved open,close if the prims have mb=1 */
  uint64_t           num_vtx;  /* number of cb_vtx_t records (already x2 for mb) */
  int64_t            material;
}
cb_shape_t;

/* ---- the reference's 4-wide node, 256 bytes -------------------------------- */
typedef struct cb_qbvh_node_t
{
  float    aabb0[6][4];   /* shutter open : [xmin,ymin,zmin,xmax,ymax,zmax][child] */
  float    aabb1[6][4];   /* shutter close */
  uint64_t child[4];      /* bit63: leaf -> ((begin<<5)|count), else node index */
  uint64_t parent;
  int64_t  axis0, axis00, axis01;
}
cb_qbvh_node_t;

#define CB_LEAF_BIT ((uint64_t)1 << 63)

#if defined(__cplusplus)
static_assert(sizeof(cb_ray_t) == 40, "ray_t layout");
static_assert(sizeof(cb_hit_t) == 100, "hit_t layout");
static_assert(sizeof(cb_vtx_t) == 16, "prims_vtx_t layout");
static_assert(sizeof(cb_qbvh_node_t) == 256, "qbvh_node_t layout");
static_assert(sizeof(cb_geo_header_t) == 32, "prims_header_t layout");
#else
_Static_assert(sizeof(cb_ray_t) == 40, "ray_t layout");
_Static_assert(sizeof(cb_hit_t) == 100, "hit_t layout");
_Static_assert(sizeof(cb_vtx_t) == 16, "prims_vtx_t layout");
_Static_assert(sizeof(cb_qbvh_node_t) == 256, "qbvh_node_t layout");
_Static_assert(sizeof(cb_geo_header_t) == 32, "prims_header_t layout");
#endif

#ifdef __cplusplus
}
#endif
#endif
