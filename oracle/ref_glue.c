/* ref_glue.c -- TEST INFRASTRUCTURE.  Thin glue that turns the *unmodified* reference
 * sources (compiled where they lie under $(REF), never copied) into a shared
 * library with batch entry points:
 *
 *     oracle/_ref/libcorona_ref.so      (and _dbg with -DACCEL_DEBUG counters)
 *
 * It #includes the reference's src/accel.d/qbvhmp.c so that the private
 * accel_t / qbvh_node_t (qbvhmp.c:62-81,173-191) can be exported for Mode-A
 * parity (the GPU traverses the CPU-built tree verbatim).  src/prims.c and
 * ext/pthread-pool are compiled as separate objects by oracle/Makefile.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this library.  The product never does.
 */
#include "src/accel.d/qbvhmp.c"   /* resolved through -I$(REF) */

#include <omp.h>
#include "corona_types.h"

/* globals every reference module reaches its state through (corona_common.h:107-108) */
rt_t rt;
__thread rt_tls_t rt_tls;

/* threads.h:33-55 calls these per worker; the accel path does not need a renderer */
struct render_tls_t *render_tls_init() { return 0; }
void render_tls_cleanup(struct render_tls_t *r) { (void)r; }

static int ref_inited = 0;

int ref_init(int num_threads)
{
  if(ref_inited) return rt.num_threads;
  memset(&rt, 0, sizeof(rt));
  rt.num_threads = num_threads > 0 ? num_threads : 1;
  rt.threads = threads_init();          /* threads.h:68 */
  threads_tls_init(rt.threads);         /* threads.h:119: pins worker k to cpu k, sets rt_tls.tid */
  ref_inited = 1;
  return rt.num_threads;
}

/* ---- primitives ------------------------------------------------------------ */
void *ref_prims_new(int num_shapes)
{
  prims_t *p = malloc(sizeof(prims_t));
  prims_init(p);
  prims_allocate(p, num_shapes);
  memset(p->shape, 0, sizeof(prims_shape_t)*num_shapes);
  rt.prims = p;
  return p;
}

/* in-memory shape (same pointers the loader would set up from the mmap, prims.c:817-823) */
int ref_prims_add_shape(void *vp, const cb_shape_t *s)
{
  prims_t *p = vp;
  const int shapeid = p->num_loaded_shapes;
  prims_shape_t *sh = p->shape + shapeid;
  sh->material  = s->material;
  sh->num_prims = s->num_prims;
  sh->primid    = (primid_t *)s->primid;
  sh->vtxidx    = (prims_vtxidx_t *)s->vtxidx;
  sh->vtx       = (prims_vtx_t *)s->vtx;
  sh->fd = -1;
  p->num_prims += s->num_prims;
  p->num_loaded_shapes++;
  return shapeid;
}

int ref_prims_load_geo(void *vp, const char *basename, int material)
{
  return prims_load_with_flags(vp, basename, "none", material, 'r', 0);
}

void ref_prims_finish(void *vp) { prims_allocate_index(vp); }
uint64_t ref_prims_num(void *vp) { return ((prims_t *)vp)->num_prims; }
const uint64_t *ref_prims_primid(void *vp) { return (const uint64_t *)((prims_t *)vp)->primid; }
int ref_prims_shape(void *vp, int shapeid, cb_shape_t *out)
{
  prims_t *p = vp;
  if(shapeid < 0 || shapeid >= (int)p->num_shapes) return 1;
  prims_shape_t *sh = p->shape + shapeid;
  out->primid = (const cb_primid_t *)sh->primid;
  out->num_prims = sh->num_prims;
  out->vtxidx = (const cb_vtxidx_t *)sh->vtxidx;
  out->vtx = (const cb_vtx_t *)sh->vtx;
  out->material = sh->material;
  /* sizes of the index/vertex arrays follow from the file layout (prims.c:804-823) */
  if(sh->data)
  {
    const prims_header_t *h = sh->data;
    out->num_vtxidx = (h->vertex_offset - h->vtxidx_offset)/sizeof(prims_vtxidx_t);
    out->num_vtx    = (sh->data_size - h->vertex_offset)/sizeof(prims_vtx_t);
  }
  return 0;
}
void ref_prims_free(void *vp)
{
  prims_t *p = vp;
  for(uint32_t k=0;k<p->num_shapes;k++) if(p->shape[k].data) munmap(p->shape[k].data, p->shape[k].data_size);
  free(p->shape); free(p->primid); free(p);
}

void ref_prim_bounds(void *vp, uint64_t primid, int close, float *aabb)
{
  primid_t pi; memcpy(&pi, &primid, 8);
  for(int d=0;d<3;d++)
  {
    if(close) prims_get_bounds_shutter_close(vp, pi, d, aabb+d, aabb+3+d);
    else      prims_get_bounds_shutter_open (vp, pi, d, aabb+d, aabb+3+d);
  }
}

/* single primitive test, for tie proofs (prims.c:638) */
void ref_prim_intersect(void *vp, uint64_t primid, const cb_ray_t *ray, cb_hit_t *hit)
{
  primid_t pi; memcpy(&pi, &primid, 8);
  prims_intersect(vp, pi, (const ray_t *)ray, (hit_t *)hit);
}
int ref_prim_visible(void *vp, uint64_t primid, const cb_ray_t *ray, float max_dist)
{
  primid_t pi; memcpy(&pi, &primid, 8);
  return prims_intersect_visible(vp, pi, (const ray_t *)ray, max_dist);
}

/* ---- accel ------------------------------------------------------------------ */
void *ref_accel_build(void *vp)
{
  accel_t *a = accel_init(vp);
  accel_build(a, 0);
  rt.accel = a;
  return a;
}
void ref_accel_free(void *a) { accel_cleanup(a); }
uint64_t ref_accel_num_nodes(void *a) { return ((accel_t *)a)->num_nodes; }
const void *ref_accel_nodes(void *a) { return ((accel_t *)a)->tree; }
const float *ref_accel_aabb(void *a) { return accel_aabb(a); }

static void hit_reset(hit_t *h, float dist)
{
  memset(h, 0, sizeof(*h));
  h->prim = INVALID_PRIMID;
  h->dist = dist;
}

/* closest hit for n rays; max_dist may be NULL (FLT_MAX, pathspace.c:762) */
void ref_intersect_n(void *a, const cb_ray_t *rays, const float *max_dist, cb_hitrec_t *out, uint64_t n, int nthreads)
{
  if(nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    rt_tls.tid = omp_get_thread_num() % rt.num_threads;
#pragma omp for schedule(dynamic, 4096)
    for(uint64_t i=0;i<n;i++)
    {
      hit_t h;
      hit_reset(&h, max_dist ? max_dist[i] : FLT_MAX);
      accel_intersect(a, (const ray_t *)(rays+i), &h);
      memcpy(out[i].prim, &h.prim, 8);
      out[i].u = h.u; out[i].v = h.v; out[i].dist = h.dist; out[i].pad = 0;
    }
  }
}

/* full hit_t in/out variant (keeps hit->x for spheres) */
void ref_intersect_hits(void *a, const cb_ray_t *rays, cb_hit_t *hits, uint64_t n)
{
  for(uint64_t i=0;i<n;i++) accel_intersect(a, (const ray_t *)(rays+i), (hit_t *)(hits+i));
}

void ref_visible_n(void *a, const cb_ray_t *rays, const float *max_dist, int32_t *out, uint64_t n, int nthreads)
{
  if(nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
  {
    rt_tls.tid = omp_get_thread_num() % rt.num_threads;
#pragma omp for schedule(dynamic, 4096)
    for(uint64_t i=0;i<n;i++)
      out[i] = accel_visible(a, (const ray_t *)(rays+i), max_dist[i]);
  }
}

/* accel_closest for n queries: ray.min_dist and {prim,u,v,dist} are in/out like the reference's ray_t / hit_t arguments */
void ref_closest_n(void *a, cb_ray_t *rays, cb_hitrec_t *io, const float *centre, uint64_t n)
{
  for(uint64_t i=0;i<n;i++)
  {
    hit_t h;
    memset(&h, 0, sizeof(h));
    memcpy(&h.prim, io[i].prim, 8);
    h.u = io[i].u; h.v = io[i].v; h.dist = io[i].dist;
    accel_closest(a, (ray_t *)(rays+i), &h, centre[i]);
    memcpy(io[i].prim, &h.prim, 8);
    io[i].u = h.u; io[i].v = h.v; io[i].dist = h.dist; io[i].pad = 0;
  }
}

/* ACCEL_DEBUG counters summed over threads: {accel_intersect, aabb_intersect, aabb_true, prims_intersect} */
int ref_counters(void *va, uint64_t *out4, int reset)
{
#ifdef ACCEL_DEBUG
  accel_t *a = va;
  out4[0] = out4[1] = out4[2] = out4[3] = 0;
  for(int t=0;t<rt.num_threads;t++)
  {
    out4[0] += a->debug[t].accel_intersect;
    out4[1] += a->debug[t].aabb_intersect;
    out4[2] += a->debug[t].aabb_true;
    out4[3] += a->debug[t].prims_intersect;
    if(reset) memset(a->debug+t, 0, sizeof(accel_debug_t));
  }
  return 1;
#else
  (void)va; (void)out4; (void)reset;
  return 0;
#endif
}
