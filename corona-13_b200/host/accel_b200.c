/* accel_b200.c -- MOD_accel=b200: the reference's accel.h module implemented on libcorona_b200.so.
 *
 * Drop-in for src/accel.d/qbvhmp.c: same eight symbols (include/accel.h:28-50), same in/out
 * conventions: the caller presets hit->dist / hit->prim, we overwrite prim,u,v,dist
 * (spheres: also x) only on a closer hit; accel_build permutes prims->primid; no error codes.
 * All traversal happens on the GPU; there is no CPU path in here.
 */
#include "corona_host.h"
#include "corona_b200.h"

#include <float.h>
#include <stdlib.h>
#include <string.h>

struct accel_t
{
  prims_t       *prims;
  cb200_scene_t *scene;
  cb200_accel_t *accel;
  float          aabb[6];
};

void accel_print_info(FILE *fd)
{
  fprintf(fd, "accel    : b200 4-wide bvh with motion-blurred boxes, built and traversed on the gpu (%s).\n", cb200_version());
}

/* number of vtxidx / vtx records of a shape: from the .geo header when the shape is a mapped file
 * (src/prims.c:804-823), else by scanning the index arrays */
static void shape_extent(const prims_shape_t *sh, uint64_t *num_vtxidx, uint64_t *num_vtx)
{
  if(sh->data && sh->data_size >= sizeof(cb_geo_header_t))
  {
    const cb_geo_header_t *h = (const cb_geo_header_t *)sh->data;
    *num_vtxidx = (h->vertex_offset - h->vtxidx_offset)/sizeof(cb_vtxidx_t);
    *num_vtx    = (sh->data_size - h->vertex_offset)/sizeof(cb_vtx_t);
    return;
  }
  uint64_t ni = 0, nv = 0;
  int mb = 0;
  const uint64_t *pid = (const uint64_t *)sh->primid;
  for(uint64_t k=0;k<sh->num_prims;k++)
  {
    const uint64_t e = (uint64_t)cb_primid_vi(pid[k]) + cb_primid_vcnt(pid[k]);
    if(e > ni) ni = e;
    if(cb_primid_mb(pid[k])) mb = 1;
  }
  for(uint64_t k=0;k<ni;k++) if((uint64_t)sh->vtxidx[k].v + 1 > nv) nv = (uint64_t)sh->vtxidx[k].v + 1;
  *num_vtxidx = ni;
  *num_vtx = nv*(mb + 1);
}

/* accel_init returns NULL without a CUDA device; the reference's callers do not check (src/main.c:344-347), so every other
 * entry point says what happened and stops instead of dereferencing it */
static void no_accel(const char *fn)
{
  fprintf(stderr, "[accel b200] %s: no accel (accel_init failed or was never called). there is no cpu fallback in this module.\n", fn);
  abort();
}

accel_t *accel_init(prims_t *p)
{
  if(!p) { fprintf(stderr, "[accel b200] accel_init: no primitives\n"); return 0; }
  if(cb200_device_count() < 1)
  {
    fprintf(stderr, "[accel b200] no CUDA device: %s. there is no cpu fallback in this module.\n", cb200_last_error());
    return 0;
  }
  accel_t *b = (accel_t *)calloc(1, sizeof(accel_t));
  b->prims = p;
  b->aabb[0] = b->aabb[1] = b->aabb[2] = FLT_MAX;
  b->aabb[3] = b->aabb[4] = b->aabb[5] = -FLT_MAX;
  return b;
}

void accel_cleanup(accel_t *b)
{
  if(!b) return;
  cb200_accel_destroy(b->accel);
  cb200_scene_destroy(b->scene);
  free(b);
}

void accel_build(accel_t *b, const char *filename)
{
  (void)filename; /* unused by qbvhmp as well */
  if(!b) no_accel("accel_build");
  prims_t *p = b->prims;
  cb_shape_t *sh = (cb_shape_t *)calloc(p->num_shapes ? p->num_shapes : 1, sizeof(cb_shape_t));
  for(uint32_t k=0;k<p->num_shapes;k++)
  {
    sh[k].primid = (const cb_primid_t *)p->shape[k].primid;
    sh[k].num_prims = p->shape[k].num_prims;
    sh[k].vtxidx = (const cb_vtxidx_t *)p->shape[k].vtxidx;
    sh[k].vtx = (const cb_vtx_t *)p->shape[k].vtx;
    sh[k].material = p->shape[k].material;
    shape_extent(p->shape + k, &sh[k].num_vtxidx, &sh[k].num_vtx);
  }
  cb200_accel_destroy(b->accel); b->accel = 0;
  cb200_scene_destroy(b->scene);
  b->scene = cb200_scene_create(sh, (int)p->num_shapes);
  free(sh);
  if(!b->scene) { fprintf(stderr, "[accel b200] scene upload failed: %s\n", cb200_last_error()); return; }
  /* side effect to preserve: the build permutes prims->primid; leaves index the permuted array */
  b->accel = cb200_accel_build(b->scene, p->ghost_aabb, (uint64_t *)p->primid);
  if(!b->accel) { fprintf(stderr, "[accel b200] build failed: %s\n", cb200_last_error()); return; }
  cb200_accel_aabb(b->accel, b->aabb);
}

const float *accel_aabb(const accel_t *b) { if(!b) no_accel("accel_aabb"); return b->aabb; }
void *accel_b200_handle(const accel_t *b) { return b ? b->accel : 0; }

void accel_intersect_n(const accel_t *b, const ray_t *rays, hit_t *hits, uint64_t n)
{
  if(!n) return;
  if(!b) no_accel("accel_intersect");
  float md1[16];
  cb_hitrec_t out1[16];   /* the single-ray callers (accel_intersect from every pinned worker) stay off the heap */
  float *md = n <= 16 ? md1 : (float *)malloc(sizeof(float)*n);
  cb_hitrec_t *out = n <= 16 ? out1 : (cb_hitrec_t *)malloc(sizeof(cb_hitrec_t)*n);
  if(!md || !out) { fprintf(stderr, "[accel b200] intersect: out of memory\n"); if(n > 16) { free(md); free(out); } return; }
  for(uint64_t i=0;i<n;i++) md[i] = hits[i].dist;
  if(cb200_accel_intersect_n(b->accel, (const cb_ray_t *)rays, md, out, n))
    fprintf(stderr, "[accel b200] intersect failed: %s\n", cb200_last_error());
  else for(uint64_t i=0;i<n;i++)
  {
    if((out[i].prim[0] & out[i].prim[1]) == 0xffffffffu) continue;   /* nothing closer than hit->dist */
    memcpy(&hits[i].prim, out[i].prim, 8);
    hits[i].u = out[i].u; hits[i].v = out[i].v; hits[i].dist = out[i].dist;
    if((out[i].prim[1] >> 29) == CB_PRIM_SPHERE)   /* include/geo/sphere.h:157 */
      for(int k=0;k<3;k++) hits[i].x[k] = rays[i].pos[k] + out[i].dist*rays[i].dir[k];
  }
  if(n > 16) { free(md); free(out); }
}

void accel_intersect(const accel_t *b, const ray_t *ray, hit_t *hit)
{
  accel_intersect_n(b, ray, hit, 1);
}

void accel_visible_n(const accel_t *b, const ray_t *rays, const float *max_dist, int *visible, uint64_t n)
{
  if(!n) return;
  if(!b) no_accel("accel_visible");
  if(cb200_accel_visible_n(b->accel, (const cb_ray_t *)rays, max_dist, (int32_t *)visible, n))
    fprintf(stderr, "[accel b200] visible failed: %s\n", cb200_last_error());
}

int accel_visible(const accel_t *b, const ray_t *ray, const float max_dist)
{
  int v = 0;
  accel_visible_n(b, ray, &max_dist, &v, 1);
  return v;
}

void accel_closest(const accel_t *b, ray_t *ray, hit_t *hit, const float centre)
{ /* half-vector MLT samplers only (include/pathspace/halfvec.h:718,912); a batch of one, in/out like qbvhmp.c:1493-1600 */
  if(!b) no_accel("accel_closest");
  cb_hitrec_t io;
  memcpy(io.prim, &hit->prim, 8);
  io.u = hit->u; io.v = hit->v; io.dist = hit->dist; io.pad = 0;
  if(cb200_accel_closest_n(b->accel, (cb_ray_t *)ray, &io, &centre, 1))
  { fprintf(stderr, "[accel b200] closest failed: %s\n", cb200_last_error()); return; }
  if(memcmp(io.prim, &hit->prim, 8) || io.dist != hit->dist)
  {
    memcpy(&hit->prim, io.prim, 8);
    hit->u = io.u; hit->v = io.v;
    if((io.prim[1] >> 29) == CB_PRIM_SPHERE && (io.prim[0] & io.prim[1]) != 0xffffffffu)
      for(int k=0;k<3;k++) hit->x[k] = ray->pos[k] + io.dist*ray->dir[k];
  }
  hit->dist = io.dist;
}

#ifndef CORONA_B200_IN_TREE
/* ---- the slice of src/prims.c the standalone host needs ------------------------------------------ */
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

void prims_init(prims_t *p)
{
  memset(p, 0, sizeof(prims_t));
  for(int k=0;k<3;k++) p->ghost_aabb[k] =  FLT_MAX;
  for(int k=3;k<6;k++) p->ghost_aabb[k] = -FLT_MAX;
}

void prims_allocate(prims_t *p, const uint32_t num_shapes)
{
  p->num_shapes = num_shapes;
  p->shape = (prims_shape_t *)calloc(num_shapes ? num_shapes : 1, sizeof(prims_shape_t));
}

/* .geo layout: header, num_prims primids, vertex indices at vtxidx_offset, vertices at vertex_offset up to
 * the end of the file.  Every section must lie inside the file and every primitive's indices inside their sections:
 * primitive -> vtxidx[vi .. vi+vcnt) -> vtx[(mb+1)*v (+mb)] (include/geo.h:108-138). */
static int geo_validate(const cb_geo_header_t *h, uint64_t size)
{
  const uint64_t np = h->num_prims;
  if(np > (size - sizeof(*h))/sizeof(uint64_t)) return 1;
  if(h->vtxidx_offset < sizeof(*h) + np*sizeof(uint64_t) || h->vtxidx_offset > size) return 1;
  if(h->vertex_offset < h->vtxidx_offset || h->vertex_offset > size) return 1;
  if((h->vtxidx_offset % 8) || (h->vertex_offset % 16)) return 1;
  const uint64_t num_idx = (h->vertex_offset - h->vtxidx_offset)/sizeof(cb_vtxidx_t);
  const uint64_t num_vtx = (size - h->vertex_offset)/sizeof(cb_vtx_t);
  const uint64_t *primid = (const uint64_t *)(h + 1);
  const cb_vtxidx_t *idx = (const cb_vtxidx_t *)((const uint8_t *)h + h->vtxidx_offset);
  for(uint64_t k=0;k<np;k++)
  {
    const uint32_t vcnt = cb_primid_vcnt(primid[k]), vi = cb_primid_vi(primid[k]), mb = cb_primid_mb(primid[k]);
    if(vcnt < 1 || vcnt > 4) return 1;   /* sphere, line, triangle, quad (prims.h:9-18); shells are not part of the hot path */
    if((uint64_t)vi + vcnt > num_idx) return 1;
    for(uint32_t c=0;c<vcnt;c++)
      if((uint64_t)(mb + 1)*idx[vi + c].v + mb >= num_vtx) return 1;
  }
  return 0;
}

/* maps <filename>.geo read-only; on failure drops the shape like the reference (prims.c:783-788) */
int prims_load(prims_t *p, const char *filename, const char *texture, const int shader)
{
  const int shapeid = p->num_loaded_shapes;
  prims_shape_t *s = p->shape + shapeid;
  s->material = shader;
  strncpy(s->tex, texture ? texture : "none", sizeof(s->tex)-1);
  char geoname[1100];
  snprintf(geoname, sizeof(geoname), "%s.geo", filename);
  const int fd = open(geoname, O_RDONLY);
  if(fd == -1)
  {
    p->num_shapes--;
    fprintf(stderr, "[prims_load] could not load geo `%s'! decreasing shape count to %d.\n", filename, p->num_shapes);
    return 1;
  }
  struct stat sb;
  fstat(fd, &sb);
  s->data_size = sb.st_size;
  s->data = mmap(0, s->data_size, PROT_READ, MAP_SHARED, fd, 0);
  close(fd);
  s->fd = -1;
  snprintf(s->name, sizeof(s->name), "%s", filename);
  if(s->data == MAP_FAILED) { perror("[prims_load] mmap"); s->data = 0; p->num_shapes--; return 1; }
  const cb_geo_header_t *h = (const cb_geo_header_t *)s->data;
  if(s->data_size < sizeof(*h) || h->magic != CB_GEO_MAGIC || h->version != CB_GEO_VERSION)
  {
    fprintf(stderr, "[prims_load] geo `%s' magic/version mismatch!\n", filename);
    munmap(s->data, s->data_size); s->data = 0;
    p->num_shapes--;
    return 1;
  }
  if(geo_validate(h, s->data_size))
  { /* the reference maps the file and trusts it (prims.c:790-806); a truncated or inconsistent file would make the device
     * kernels read outside their buffers here, so it is dropped like an unreadable one */
    fprintf(stderr, "[prims_load] geo `%s' is truncated or inconsistent (sections or indices outside the file)! decreasing shape count to %d.\n",
            filename, p->num_shapes - 1);
    munmap(s->data, s->data_size); s->data = 0;
    p->num_shapes--;
    return 1;
  }
  s->primid = (uint64_t *)(h + 1);
  s->num_prims = h->num_prims;
  s->vtxidx = (cb_vtxidx_t *)((uint8_t *)h + h->vtxidx_offset);
  s->vtx = (cb_vtx_t *)((uint8_t *)h + h->vertex_offset);
  p->num_prims += h->num_prims;
  p->num_loaded_shapes++;
  return 0;
}

int prims_add_shape_mem(prims_t *p, const cb_shape_t *sh)
{
  const int shapeid = p->num_loaded_shapes;
  prims_shape_t *s = p->shape + shapeid;
  s->material = sh->material;
  s->num_prims = sh->num_prims;
  s->primid = (uint64_t *)sh->primid;
  s->vtxidx = (cb_vtxidx_t *)sh->vtxidx;
  s->vtx = (cb_vtx_t *)sh->vtx;
  s->fd = -1;
  snprintf(s->name, sizeof(s->name), "mem%d", shapeid);
  p->num_prims += sh->num_prims;
  p->num_loaded_shapes++;
  return shapeid;
}

void prims_allocate_index(prims_t *p)
{ /* prims.c:741-757 */
  p->primid = (uint64_t *)malloc(sizeof(uint64_t)*(p->num_prims ? p->num_prims : 1));
  if(!p->primid) { fprintf(stderr, "[prims] out of memory for %lu primitive ids\n", (unsigned long)p->num_prims); p->num_prims = 0; return; }
  uint64_t n = 0;
  for(uint32_t shapeid=0;shapeid<p->num_shapes;shapeid++)
    for(uint64_t k=0;k<p->shape[shapeid].num_prims;k++)
      p->primid[n++] = cb_primid_with_shapeid(p->shape[shapeid].primid[k], shapeid);
}

void prims_cleanup(prims_t *p)
{
  for(uint32_t k=0;k<p->num_shapes;k++)
    if(p->shape[k].data) munmap(p->shape[k].data, p->shape[k].data_size);
  free(p->shape);
  free(p->primid);
  memset(p, 0, sizeof(prims_t));
}
#endif
