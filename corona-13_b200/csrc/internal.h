// internal.h -- data layout in HBM and host-side handles of libcorona_b200.so (not part of the ABI)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "corona_b200.h"

// ---------------------------------------------------------------------------------------------
// HBM layout
//
// prim records ("leaf order"): one or two 64-byte units per primitive, stored in the order of the
// build-permuted primid list so that leaf (begin,count) indexes them directly -- the reference's
// two dependent loads per vertex (vtxidx -> vtx, include/geo.h:108-138) are resolved once at build
// time.  Unit 0 = shutter-open vertices, unit 1 (only when the accel has rec_units == 2) =
// shutter-close vertices.
//   float4 r0 = (v0.xyz, primid.lo)      float4 r1 = (v1.xyz, primid.hi)
//   float4 r2 = (v2.xyz, aux0)           float4 r3 = (v3.xyz, aux1)
// sphere: v0 = centre, aux0 = radius bits.   line: v0,v1, aux0 = r0, aux1 = r1 (shutter-open radii).
//
// nodes: Node256 is bit-identical to the reference's qbvh_node_t (qbvhmp.c:62-81): used for scenes
// with motion blur and for imported (mode A) trees.  Node128 drops the shutter-close boxes for
// static scenes: half the bytes per node visit.
// ---------------------------------------------------------------------------------------------
struct __align__(16) Node256
{
  float    aabb0[6][4];
  float    aabb1[6][4];
  uint64_t child[4];
  uint64_t parent;
  int64_t  axis0, axis00, axis01;
};
struct __align__(16) Node128
{
  float    aabb0[6][4];
  uint64_t child[4];       // child[0] bits 56..61 carry axis0 | axis00<<2 | axis01<<4
};
// Node8 (static scenes, the throughput path): 8 children in 96 bytes = three 32-byte sectors.  The children's boxes are
// quantised to 8 bits per plane on a per-node grid: plane = origin[k] + q * 2^e[k]; lower planes are rounded down, upper
// planes up, so a stored box always CONTAINS the box of the 4-wide tree's child -- boxes only cull, the primitive tests
// behind them are the reference's own arithmetic, so prim / u / v / dist are unchanged except where two primitives tie or a
// box test decided by its last ulp (the tree-dependent cases mode B already classifies).
//   planes[4*k+0..1] = lower planes of children 0..3 / 4..7 along axis k (one byte each), planes[4*k+2..3] = upper planes
//   exps = E_x | E_y << 8 | E_z << 16: biased IEEE exponent of 128 * 2^e[k] (the traversal decodes a byte as q/128)
//   child[c]: 0 = empty slot (its box is inverted: lower 255, upper 0), bit 31 = leaf (begin << 3 | count, count <= 7),
//             else index of the child node.  Slots are assigned by octant (child centre relative to the node centre), so
//             that slot ^ ray octant approximates the front-to-back order (Ylitie, Karras, Laine 2017).
struct __align__(32) Node8
{
  float    origin[3];
  uint32_t exps;
  uint32_t planes[12];
  uint32_t child[8];
};
static_assert(sizeof(Node256) == 256, "Node256");
static_assert(sizeof(Node128) == 128, "Node128");
static_assert(sizeof(Node8) == 96, "Node8");
#define CB8_LEAF 0x80000000u
#define CB8_STACK 32          // group entries; deeper trees fall back to the 4-wide kernels

#define CB_CHILD_MASK  0x80ffffffffffffffull   // strips the axis bits from Node128::child[0]
#define CB_AXIS_SHIFT  56

struct DevAccel
{
  const void   *nodes;       // Node256* (mb != 0) or Node128*
  const float4 *recs;        // prim records, rec_units*4 float4 per primitive
  uint32_t      rec_units;   // 1: static scene, 2: open+close vertices
  uint32_t      mb;          // nodes carry shutter-close boxes
  uint32_t      imported;    // tree adopted from the reference builder (empty leaves may have ordinary boxes)
  uint32_t      pad_;
  uint64_t      num_nodes;
  uint64_t      num_prims;
  const Node8  *nodes8;      // 8-wide compressed tree over the same primitive order (static scenes), or null
  uint32_t      num_nodes8;
  uint32_t      depth8;
};

struct ShapeDev   // offsets of one shape inside the concatenated device arrays
{
  uint64_t vtx_off;
  uint64_t vtxidx_off;
};

struct cb200_scene
{
  int        device;
  int        num_shapes;
  uint64_t   num_prims, num_vtx, num_vtxidx;
  int        any_mb;
  int        any_analytic;   // spheres / lines present (selects the traversal kernel variant)
  cb_vtx_t    *d_vtx;
  cb_vtxidx_t *d_vtxidx;
  ShapeDev    *d_shapes;
  uint64_t    *d_primid;   // global list in load order (shapeid patched in, prims.c:741-757)
  std::vector<int64_t> h_material;   // shape -> material (prims_shader)
};

struct cb200_accel
{
  cb200_scene *scene;
  DevAccel     dev;
  void        *d_nodes;
  float4      *d_recs;
  uint64_t    *d_primid;   // permuted list the leaves index into
  float        aabb[6];
  int          depth;      // levels of 4-wide nodes
  int          imported;
  Node8       *d_nodes8;
  int          traversal;  // CB200_TRAVERSAL_*: which tree the launchers use (set by the build, changed by cb200_accel_set_traversal)
};

// error plumbing ------------------------------------------------------------------------------
void cb200_set_error(const std::string &msg);
int  cb200_cuda_fail(cudaError_t e, const char *what, const char *file, int line);
#define CB_CUDA(call) do { cudaError_t e__ = (call); if(e__ != cudaSuccess) return cb200_cuda_fail(e__, #call, __FILE__, __LINE__); } while(0)
#define CB_CUDA_NULL(call) do { cudaError_t e__ = (call); if(e__ != cudaSuccess) { cb200_cuda_fail(e__, #call, __FILE__, __LINE__); return nullptr; } } while(0)
void cb200_count_launch(uint64_t n = 1);
int  cb200_sm_count_cached();

// build.cu
int cb200_build_lbvh(cb200_accel *a, const float *ghost_aabb);
int cb200_build_records(cb200_accel *a, cudaStream_t stream);
// traverse.cu
// d_counters (optional, 4 words): rays, node visits, child boxes hit, primitive tests (ACCEL_DEBUG definitions, qbvhmp.c:83-90).
// The tree is picked by a->traversal: the 8-wide compressed one when the accel has it, else the 4-wide tree the reference's order is defined on
int cb200_launch_intersect(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                           uint64_t n, cudaStream_t stream, unsigned long long *d_counters);
int cb200_launch_visible(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, int32_t *d_out,
                         uint64_t n, cudaStream_t stream);
int cb200_launch_closest(const cb200_accel *a, cb_ray_t *d_rays, cb_hitrec_t *d_io, const float *d_centre, uint64_t n, cudaStream_t stream);
// next-event visibility (path_visible semantics): any primitive other than d_light_prim[i] accepted by the closest-hit rules
int cb200_launch_shadow(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, const uint2 *d_light_prim, int32_t *d_out,
                        uint64_t n, cudaStream_t stream);
// traverse8.cu: the same three entries on the 8-wide compressed tree (a->dev.nodes8 != null)
int cb200_launch_intersect8(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                            uint64_t n, cudaStream_t stream, unsigned long long *d_counters);
int cb200_launch_visible8(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, const uint2 *d_light_prim, int32_t *d_out,
                          uint64_t n, cudaStream_t stream);
bool cb200_use_wide8(const cb200_accel *a);
// traverse2.cu: closest hit with two rays per lane (static scenes, 32-bit child references, small stack)
#ifndef CB200_DUAL_DEFAULT
#define CB200_DUAL_DEFAULT 0
#endif
bool cb200_dual_enabled();
int cb200_launch_intersect_dual(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                                uint64_t n, cudaStream_t stream);
// shared launch plumbing (traverse.cu)
int cb200_get_ticket(cudaStream_t stream, unsigned int **t);
int cb200_trace_grid(uint64_t n, const void *kernel);
int cb200_prim_threshold();
int cb200_refill_threshold();
