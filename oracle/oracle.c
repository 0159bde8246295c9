/* oracle.c -- TEST INFRASTRUCTURE (see oracle.h).  Plain-C restatement of the reference's
 * acceleration-structure hot path, written from the reference's behaviour with the same
 * floating point operation order (no FMA contraction: build with -ffp-contract=off).
 *
 * Each function cites the reference file:line it follows (paths relative to the reference
 * root).  Pinned bit-for-bit against the compiled reference by tests/test_oracle_vs_ref.py.
 */
#include "oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <omp.h>

#define ORC_MAX_TREE_DEPTH 100   /* qbvhmp.c:38 */
#define ORC_SAH_TESTS 7          /* qbvhmp.c:40 */
#define ORC_SAH_LOG_STEP 3       /* qbvhmp.c:42 */
#define ORC_PRIMS_PER_LEAF 6     /* qbvhmp.c:44 */

#define MINM(a, b) ((a) < (b) ? (a) : (b))   /* corona_common.h:169-171, same select semantics */
#define MAXM(a, b) ((a) > (b) ? (a) : (b))

struct orc_scene_t
{
  int num_shapes;
  cb_shape_t *shape;
  uint64_t num_prims;
  uint64_t *primid;      /* global list, shapeid patched in (prims.c:741-757), permuted by the build */
};

struct orc_accel_t
{
  orc_scene_t *scene;
  cb_qbvh_node_t *tree;
  uint64_t num_nodes, node_bufsize;
  float aabb[6];
  float *prim_aabb;      /* build scratch */
};

/* ---------------------------------------------------------------------------- small math */
static inline float dot3(const float *u, const float *v) { return (u[0]*v[0] + u[1]*v[1]) + u[2]*v[2]; } /* corona_common.h:166 */
static inline void cross3(const float *v1, const float *v2, float *res)                                 /* corona_common.h:161-164 */
{
  res[0] = v1[1]*v2[2] - v2[1]*v1[2];
  res[1] = v1[2]*v2[0] - v2[2]*v1[0];
  res[2] = v1[0]*v2[1] - v2[0]*v1[1];
}
static inline void normalise3(float *f)                                                                  /* corona_common.h:173-177 */
{
  const float len = 1.0f/sqrtf(dot3(f, f));
  for(int k=0;k<3;k++) f[k] *= len;
}
static inline void onb3(const float *n, float *u, float *v)                                              /* corona_common.h:179-200 */
{
  if(fabsf(n[1]) < 0.5) { const float up[3] = {0, 1, 0}; cross3(n, up, u); }
  else                  { const float rg[3] = {1, 0, 0}; cross3(n, rg, u); }
  normalise3(u);
  cross3(n, u, v);
}
static inline float u2f(uint32_t i) { union { uint32_t i; float f; } u = {.i=i}; return u.f; }

/* ---------------------------------------------------------------------------- scene */
orc_scene_t *orc_scene_new(const cb_shape_t *shapes, int num_shapes)
{
  orc_scene_t *s = calloc(1, sizeof(*s));
  s->num_shapes = num_shapes;
  s->shape = malloc(sizeof(cb_shape_t)*(num_shapes > 0 ? num_shapes : 1));
  for(int i=0;i<num_shapes;i++) { s->shape[i] = shapes[i]; s->num_prims += shapes[i].num_prims; }
  s->primid = malloc(sizeof(uint64_t)*(s->num_prims ? s->num_prims : 1));
  uint64_t n = 0;
  for(int i=0;i<num_shapes;i++)
    for(uint64_t k=0;k<shapes[i].num_prims;k++)
      s->primid[n++] = cb_primid_with_shapeid(shapes[i].primid[k], (uint32_t)i);
  return s;
}
void orc_scene_free(orc_scene_t *s) { if(!s) return; free(s->shape); free(s->primid); free(s); }
uint64_t orc_scene_num_prims(const orc_scene_t *s) { return s->num_prims; }
uint64_t *orc_scene_primid(orc_scene_t *s) { return s->primid; }

/* vertex fetches: geo.h:108-138.  two dependent loads: vtxidx -> vtx */
static inline const cb_vtx_t *vtx_open(const orc_scene_t *s, cb_primid_t pi, int v)
{
  const cb_shape_t *sh = s->shape + cb_primid_shapeid(pi);
  return sh->vtx + (uint64_t)(cb_primid_mb(pi)+1)*sh->vtxidx[cb_primid_vi(pi) + v].v;
}
static inline const cb_vtx_t *vtx_close(const orc_scene_t *s, cb_primid_t pi, int v)
{
  const cb_shape_t *sh = s->shape + cb_primid_shapeid(pi);
  return sh->vtx + (uint64_t)(cb_primid_mb(pi)+1)*sh->vtxidx[cb_primid_vi(pi) + v].v + cb_primid_mb(pi);
}
static inline void vtx_time(const orc_scene_t *s, cb_primid_t pi, int v, float time, float *out)
{
  const cb_shape_t *sh = s->shape + cb_primid_shapeid(pi);
  if(cb_primid_mb(pi))
  { /* geo.h:128-130: (1-time)*open + time*close, mul mul add */
    const cb_vtx_t *a = sh->vtx + 2*(uint64_t)sh->vtxidx[cb_primid_vi(pi) + v].v;
    const float t0 = 1.0f-time;
    for(int k=0;k<3;k++) out[k] = t0*a[0].v[k] + time*a[1].v[k];
  }
  else
  {
    const cb_vtx_t *a = sh->vtx + sh->vtxidx[cb_primid_vi(pi) + v].v;
    for(int k=0;k<3;k++) out[k] = a->v[k];
  }
}
/* radii come from the shutter-open vertex only: sphere.h:9-13, line.h:9-15 */
static inline float radius_of(const orc_scene_t *s, cb_primid_t pi, int v) { return u2f(vtx_open(s, pi, v)->n); }

/* ---------------------------------------------------------------------------- bounds */
static void line_bounds(const float *v0, const float *v1, float r0, float r1, int dim, float *min, float *max)
{ /* line.h:24-39 */
  float d[3], a[3], b[3];
  for(int k=0;k<3;k++) d[k] = v1[k] - v0[k];
  normalise3(d);
  onb3(d, a, b);
  const float theta = atan2f(a[dim], b[dim]);
  const float m = fabsf(sinf(theta)*a[dim]) + fabsf(cosf(theta)*b[dim]);
  *min = fminf(v0[dim] - r0*m, v1[dim] - r1*m);
  *max = fmaxf(v0[dim] + r0*m, v1[dim] + r1*m);
}

static void prim_bounds_dim(const orc_scene_t *s, cb_primid_t pi, int close, int dim, float *min, float *max)
{ /* prims.c:20-60 dispatch */
  const uint32_t vcnt = cb_primid_vcnt(pi);
  if(vcnt == CB_PRIM_SPHERE)
  { /* sphere.h:16-30 */
    const float *v = close ? vtx_close(s, pi, 0)->v : vtx_open(s, pi, 0)->v;
    const float radius = radius_of(s, pi, 0);
    *min = v[dim] - radius;
    *max = v[dim] + radius;
  }
  else if(vcnt == CB_PRIM_LINE)
  {
    const float *v0 = close ? vtx_close(s, pi, 0)->v : vtx_open(s, pi, 0)->v;
    const float *v1 = close ? vtx_close(s, pi, 1)->v : vtx_open(s, pi, 1)->v;
    line_bounds(v0, v1, radius_of(s, pi, 0), radius_of(s, pi, 1), dim, min, max);
  }
  else
  { /* triangle.h:7-33 (tri and quad; shells are out of scope) */
    const float *v = close ? vtx_close(s, pi, 0)->v : vtx_open(s, pi, 0)->v;
    *min = v[dim]; *max = v[dim];
    for(uint32_t k=1;k<vcnt;k++)
    {
      const float *w = close ? vtx_close(s, pi, k)->v : vtx_open(s, pi, k)->v;
      *min = fminf(w[dim], *min);
      *max = fmaxf(w[dim], *max);
    }
  }
}

void orc_prim_bounds(const orc_scene_t *s, cb_primid_t pi, int shutter_close, float aabb[6])
{
  for(int d=0;d<3;d++) prim_bounds_dim(s, pi, shutter_close, d, aabb+d, aabb+3+d);
}

/* ---------------------------------------------------------------------------- primitive tests */
static inline uint64_t ray_ignore(const cb_ray_t *r) { return (uint64_t)r->ignore[0] | ((uint64_t)r->ignore[1] << 32); }
static inline void hit_set_prim(cb_hit_t *h, cb_primid_t pi) { h->prim[0] = (uint32_t)pi; h->prim[1] = (uint32_t)(pi>>32); }

/* triangle.h:263-305: float moeller-trumbore, epsilon free; 'v' weights v1, 'u' weights v2 */
static int tri_intersect(const float *v0, const float *v1, const float *v2, cb_primid_t pi, const cb_ray_t *ray, cb_hit_t *hit)
{
  if(pi == ray_ignore(ray)) return 0;
  float edge1[3], edge2[3], tvec[3], pvec[3], qvec[3];
  for(int k=0;k<3;k++) { edge1[k] = v1[k] - v0[k]; edge2[k] = v2[k] - v0[k]; }
  cross3(ray->dir, edge2, pvec);
  const float det = dot3(edge1, pvec);
  const float inv_det = 1.0f / det;
  for(int k=0;k<3;k++) tvec[k] = ray->pos[k] - v0[k];
  const float v = dot3(tvec, pvec) * inv_det;
  if(v < 0.0f || v > 1.0f) return 0;
  cross3(tvec, edge1, qvec);
  const float u = dot3(ray->dir, qvec) * inv_det;
  if(u < 0.0f || u + v > 1.0f) return 0;
  const float dist = dot3(edge2, qvec) * inv_det;
  if(dist > ray->min_dist && dist <= hit->dist)
  {
    hit->dist = dist;
    hit_set_prim(hit, pi);
    hit->u = u;
    hit->v = v;
    return 1;
  }
  return 0;
}

/* triangle.h:308-343: no ignore test; 1.0/det in double then rounded == float division */
static int tri_visible(const float *v0, const float *v1, const float *v2, const cb_ray_t *ray, float max_dist)
{
  float edge1[3], edge2[3], tvec[3], pvec[3], qvec[3];
  for(int k=0;k<3;k++) { edge1[k] = v1[k] - v0[k]; edge2[k] = v2[k] - v0[k]; }
  cross3(ray->dir, edge2, pvec);
  const float det = dot3(edge1, pvec);
  const float inv_det = (float)(1.0 / (double)det);
  for(int k=0;k<3;k++) tvec[k] = ray->pos[k] - v0[k];
  const float v = dot3(tvec, pvec) * inv_det;
  if(v < 0.0 || v > 1.0) return 0;
  cross3(tvec, edge1, qvec);
  const float u = dot3(ray->dir, qvec) * inv_det;
  if(u < 0.0 || u + v > 1.0) return 0;
  const float dist = dot3(edge2, qvec) * inv_det;
  if(dist > 0.0 && dist <= max_dist) return 1;
  return 0;
}

/* sphere.h:112-144 */
static float sphere_t(const float *center, float radius, const cb_ray_t *ray)
{
  const float a = dot3(ray->dir, ray->dir);
  const float o[3] = {ray->pos[0]-center[0], ray->pos[1]-center[1], ray->pos[2]-center[2]};
  const float b = 2.0f*dot3(o, ray->dir);
  const float c = dot3(o, o) - radius*radius;
  if(a == 0)
  {
    if(b != 0) return -c / b;
    return -FLT_MAX;
  }
  const float discrim = b*b - 4.0f*a*c;
  if(discrim < 0) return -FLT_MAX;
  float temp;
  const float sqrt_discrim = sqrtf(discrim);
  if(b < 0) temp = -0.5f * (b - sqrt_discrim);
  else      temp = -0.5f * (b + sqrt_discrim);
  const float x0 = temp / a;
  const float x1 = c / temp;
  if(x0 <= 0.0f) return x1;
  else if(x1 <= 0.0f) return x0;
  else return fminf(x0, x1);
}

/* line.h:362-445 (the live float variant) */
static float cylinder_t(const float *v0, const float *v1, float r, const cb_ray_t *ray, float *out, float *len)
{
  float d[3], a[3], b[3], o[3] = {0.0f, 0.0f, 0.0f};
  for(int k=0;k<3;k++) d[k] = v1[k] - v0[k];
  if(r < 0.01)
  { /* hair: line strip */
    cross3(d, ray->dir, a);
    for(int k=0;k<3;k++) o[k] = v0[k] - ray->pos[k];
    const float dotp = dot3(o, a);
    const float ilen = (float)(1.0/(double)sqrtf(dot3(a, a)));
    const float dist = fabsf(dotp*ilen);
    if(dist > r) return -1.0f;
    const float dlen = sqrtf(dot3(d, d));
    if(len) *len = dlen;
    cross3(d, a, b);
    const float t = dot3(o, b)/dot3(b, ray->dir);
    for(int k=0;k<3;k++) out[k] = t * ray->dir[k] - o[k];
    out[0] = dot3(out, d)/dlen;
    out[1] = 0.0f;
    out[2] = 1.0f;
    if(out[0] >= 0.0 && out[0] <= dlen) return t;
    return -1.0f;
  }
  const float dlen = sqrtf(dot3(d, d));
  if(len) *len = dlen;
  for(int k=0;k<3;k++) d[k] *= 1.0f/dlen;
  onb3(d, a, b);
  float w[3] = {0.0f, 0.0f, 0.0f};
  for(int k=0;k<3;k++)
  {
    o[0] += (ray->pos[k] - v0[k])*d[k];
    o[1] += (ray->pos[k] - v0[k])*a[k];
    o[2] += (ray->pos[k] - v0[k])*b[k];
    w[0] += ray->dir[k]*d[k];
    w[1] += ray->dir[k]*a[k];
    w[2] += ray->dir[k]*b[k];
  }
  const float A = w[1]*w[1] + w[2]*w[2];
  const float B = 2.0f*(o[1]*w[1]+o[2]*w[2]);
  const float C = o[1]*o[1]+o[2]*o[2] - r*r;
  const float discr = (float)((double)(B*B) - 4.0*(double)A*(double)C);
  if(discr < 0.0) return -1.0f;
  float temp;
  const float sqrt_discrim = sqrtf(discr);
  if(B < 0) temp = -0.5f * (B - sqrt_discrim);
  else      temp = -0.5f * (B + sqrt_discrim);
  const float t0 = temp / A;
  const float t1 = C / temp;
  float t;
  if(t0 <= 0.0f) t = t1;
  else if(t1 <= 0.0f) t = t0;
  else
  {
    t = fminf(t0, t1);
    for(int i=0;i<2;i++)
    {
      for(int k=0;k<3;k++) out[k] = o[k] + t*w[k];
      if(out[0] >= 0.0 && out[0] <= dlen) return t;
      t = fmaxf(t0, t1);
    }
    return -1.0f;
  }
  for(int k=0;k<3;k++) out[k] = o[k] + t*w[k];
  if(out[0] >= 0.0 && out[0] <= dlen) return t;
  else return -1.0f;
}

/* line.h:447-516 */
static float cone_t(const float *v0, const float *v1, float r0, float r1, const cb_ray_t *ray, float dist, cb_hit_t *hit)
{
  float iraylen = 1.0f;
  if(!hit) iraylen = 1.0f/sqrtf(dot3(ray->dir, ray->dir));
  float d[3];
  for(int k=0;k<3;k++) d[k] = v1[k] - v0[k];
  const float d_len = sqrtf(dot3(d, d));
  for(int k=0;k<3;k++) d[k] = (float)((double)d[k] * (1.0/(double)d_len));
  const float cos_dr = dot3(d, ray->dir)*iraylen;
  const float cos_a2 = d_len*d_len/((r1-r0)*(r1-r0) + d_len*d_len);
  float tip[3], o[3];
  const float tt = -r0*d_len/(r1-r0);
  for(int k=0;k<3;k++) tip[k] = v0[k] + tt*d[k];
  for(int k=0;k<3;k++) o[k] = ray->pos[k] - tip[k];
  const float cos_do = dot3(d, o);
  const float cos_ro = dot3(ray->dir, o)*iraylen;
  const float cos_oo = dot3(o, o);
  const float c2 = cos_dr*cos_dr - cos_a2;
  const float c1 = cos_dr*cos_do - cos_a2*cos_ro;
  const float c0 = cos_do*cos_do - cos_a2*cos_oo;
  float tmin = -1.0f;
  if(fabsf(c2) > 0.0)
  {
    const float discr = c1*c1 - c0*c2;
    if(discr < 0.0f) return -1.0f;
    const float root = sqrtf(discr);
    float x[3];
    for(int i=-1;i<2;i+=2)
    {
      const float t = (-c1 + i*root)/c2;
      if(t > 0.0 && t < dist/iraylen)
      {
        for(int k=0;k<3;k++) x[k] = ray->pos[k] + t*ray->dir[k]*iraylen - v0[k];
        const float dt = dot3(x, d);
        if(dt >= 0.0f && dt <= d_len)
        {
          if(hit)
          {
            hit->u = dt/d_len;
            float a[3], b[3];
            onb3(d, a, b);
            hit->v = (float)((double)atan2f(dot3(a, x), dot3(b, x))/(2.0f*M_PI));
          }
          tmin = dist = t;
        }
      }
    }
  }
  return tmin;
}

/* prims.c:638-672 */
void orc_prim_intersect(const orc_scene_t *s, cb_primid_t pi, const cb_ray_t *ray, cb_hit_t *hit)
{
  const uint32_t vcnt = cb_primid_vcnt(pi);
  if(vcnt == CB_PRIM_TRI || vcnt == CB_PRIM_QUAD)
  {
    float v0[3], v1[3], v2[3], v3[3];
    vtx_time(s, pi, 0, ray->time, v0);
    vtx_time(s, pi, 1, ray->time, v1);
    vtx_time(s, pi, 2, ray->time, v2);
    if(vcnt == 3) tri_intersect(v0, v1, v2, pi, ray, hit);
    else
    { /* quad = (v0 v1 v2) then, only if that missed, (v0 v2 v3) */
      if(tri_intersect(v0, v1, v2, pi, ray, hit)) { hit->v += hit->u; return; }
      vtx_time(s, pi, 3, ray->time, v3);
      if(tri_intersect(v0, v2, v3, pi, ray, hit)) hit->u += hit->v;
    }
  }
  else if(vcnt == CB_PRIM_SPHERE)
  { /* sphere.h:146-166; note: strict '<' against hit->dist, ignore is not honoured */
    float center[3];
    vtx_time(s, pi, 0, ray->time, center);
    const float radius = radius_of(s, pi, 0);
    const float t = sphere_t(center, radius, ray);
    if(t > ray->min_dist && t < hit->dist)
    {
      hit->dist = t;
      hit_set_prim(hit, pi);
      for(int k=0;k<3;k++) hit->x[k] = ray->pos[k] + t*ray->dir[k];
      hit->u = (float)((double)atan2f((hit->x[1]-center[1])/radius, (hit->x[0]-center[0])/radius)/(2.0f*M_PI));
      const float cz = (hit->x[2]-center[2])/radius;
      hit->v = (float)((double)acosf(MINM(MAXM(cz, -1.0f), 1.0f))/M_PI);
    }
  }
  else if(vcnt == CB_PRIM_LINE)
  { /* line.h:518-560 */
    const float r0 = radius_of(s, pi, 0), r1 = radius_of(s, pi, 1);
    const int linestrip = MAXM(r0, r1) <= 1e-2f;
    if(linestrip && ray_ignore(ray) == pi) return;
    float v0[3], v1[3];
    vtx_time(s, pi, 0, ray->time, v0);
    vtx_time(s, pi, 1, ray->time, v1);
    float out[3], len;
    if(fabsf(r1-r0) < 1e-3)
    {
      const float t = cylinder_t(v0, v1, r0, ray, out, &len);
      if(t > ray->min_dist && t < hit->dist)
      {
        hit->dist = t;
        hit_set_prim(hit, pi);
        hit->u = out[0]/len;
        hit->v = (float)((double)atan2f(out[1], out[2])/(2.0f*M_PI));
      }
    }
    else
    {
      const float t = cone_t(v0, v1, r0, r1, ray, hit->dist, hit);
      if((linestrip && t > MAXM(ray->min_dist, 1e-3f)) || (!linestrip && t > ray->min_dist))
      {
        hit->dist = t;
        hit_set_prim(hit, pi);
      }
    }
  }
  /* shells (vcnt 5): out of scope, never hit */
}

/* prims.c:674-701 */
int orc_prim_visible(const orc_scene_t *s, cb_primid_t pi, const cb_ray_t *ray, float max_dist)
{
  const uint32_t vcnt = cb_primid_vcnt(pi);
  if(vcnt == CB_PRIM_TRI || vcnt == CB_PRIM_QUAD)
  {
    float v0[3], v1[3], v2[3], v3[3];
    vtx_time(s, pi, 0, ray->time, v0);
    vtx_time(s, pi, 1, ray->time, v1);
    vtx_time(s, pi, 2, ray->time, v2);
    if(vcnt == 3) return tri_visible(v0, v1, v2, ray, max_dist);
    if(tri_visible(v0, v1, v2, ray, max_dist)) return 1;
    vtx_time(s, pi, 3, ray->time, v3);
    if(tri_visible(v0, v2, v3, ray, max_dist)) return 1;
    return 0;
  }
  else if(vcnt == CB_PRIM_SPHERE)
  { /* sphere.h:168-180 */
    float center[3];
    vtx_time(s, pi, 0, ray->time, center);
    const float t = sphere_t(center, radius_of(s, pi, 0), ray);
    return (t > 0.0f && t <= max_dist) ? 1 : 0;
  }
  else if(vcnt == CB_PRIM_LINE)
  { /* line.h:563-592 */
    const float r0 = radius_of(s, pi, 0), r1 = radius_of(s, pi, 1);
    const int linestrip = MAXM(r0, r1) <= 1e-2f;
    if(linestrip && ray_ignore(ray) == pi) return 0;
    float v0[3], v1[3];
    vtx_time(s, pi, 0, ray->time, v0);
    vtx_time(s, pi, 1, ray->time, v1);
    if(fabsf(r1-r0) < 1e-3)
    {
      float out[3];
      const float t = cylinder_t(v0, v1, r0, ray, out, 0);
      if(((linestrip && t > MAXM(ray->min_dist, 1e-3f)) || (!linestrip && t > ray->min_dist)) && t <= max_dist) return 1;
    }
    else
    {
      const float t = cone_t(v0, v1, r0, r1, ray, max_dist, 0);
      if(t > ray->min_dist) return 1;
    }
    return 0;
  }
  return 0;
}

/* ---------------------------------------------------------------------------- traversal */
/* qbvhmp.c:1188-1246: four slab tests with time-interpolated boxes.
 * SSE min/max select semantics: min(a,b) = a<b?a:b, max(a,b) = a>b?a:b (second operand on NaN). */
static inline void node_boxes(const cb_qbvh_node_t *n, float t0, float t1, const float *pos, const float *invdir,
                              float *tmin, float *tmax)
{
  for(int k=0;k<3;k++)
    for(int c=0;c<4;c++)
    {
      const float lo = ((n->aabb0[k  ][c]*t0 + n->aabb1[k  ][c]*t1) - pos[k]) * invdir[k];
      const float hi = ((n->aabb0[k+3][c]*t0 + n->aabb1[k+3][c]*t1) - pos[k]) * invdir[k];
      const float mn = lo < hi ? lo : hi;
      const float mx = lo > hi ? lo : hi;
      tmin[c] = tmin[c] > mn ? tmin[c] : mn;
      tmax[c] = tmax[c] < mx ? tmax[c] : mx;
    }
}

static inline uint32_t signbit_u(float f) { union { float f; uint32_t i; } u = {.f=f}; return u.i >> 31; }

/* qbvhmp.c:1262-1390 */
void orc_intersect(const orc_accel_t *acc, const cb_ray_t *ray, cb_hit_t *hit, uint64_t cnt[4])
{
  if(cnt) cnt[0]++;
  uint32_t near[3], far[3];
  for(int k=0;k<3;k++) { near[k] = signbit_u(ray->dir[k]); far[k] = 1 ^ near[k]; }
  const cb_qbvh_node_t *node = acc->tree;
  uint64_t stack[3*ORC_MAX_TREE_DEPTH];
  float stack_dist[3*ORC_MAX_TREE_DEPTH];
  int stackpos = 0;
  uint64_t current;
  const float t0 = 1.0f - ray->time, t1 = ray->time;
  float invdir[3];
  for(int k=0;k<3;k++) invdir[k] = 1.0f/ray->dir[k];

  while(1)
  {
    float tmin[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float tmax[4] = {hit->dist, hit->dist, hit->dist, hit->dist};
    node_boxes(node, t0, t1, ray->pos, invdir, tmin, tmax);
    int i[4], any = 0;
    for(int c=0;c<4;c++) { i[c] = tmin[c] <= tmax[c]; any |= i[c]; }
    int need_pop = 0;
    if(!any) need_pop = 1;
    else
    {
      if(cnt) { cnt[1]++; for(int c=0;c<4;c++) cnt[2] += i[c]; }
      const int axis0  = (int)node->axis0;
      const int axis1n = near[axis0] ? (int)node->axis01 : (int)node->axis00;
      const int axis1f = near[axis0] ? (int)node->axis00 : (int)node->axis01;
      const uint32_t n11 = (far [axis0]<<1) | far [axis1f];
      const uint32_t n10 = (far [axis0]<<1) | near[axis1f];
      const uint32_t n01 = (near[axis0]<<1) | far [axis1n];
      const uint32_t n00 = (near[axis0]<<1) | near[axis1n];
#define PUSH(c) do { stack_dist[stackpos] = tmin[c]; stack[stackpos++] = node->child[c]; } while(0)
      if(i[n00])
      {
        current = node->child[n00];
        if(i[n11]) PUSH(n11);
        if(i[n10]) PUSH(n10);
        if(i[n01]) PUSH(n01);
      }
      else if(i[n01])
      {
        current = node->child[n01];
        if(i[n11]) PUSH(n11);
        if(i[n10]) PUSH(n10);
      }
      else if(i[n10])
      {
        current = node->child[n10];
        if(i[n11]) PUSH(n11);
      }
      else current = node->child[n11]; /* any => i[n11] */
#undef PUSH
    }
    if(need_pop)
    {
      do
      {
        if(stackpos == 0) return;
        stackpos--;
        current = stack[stackpos];
      }
      while(stack_dist[stackpos] > hit->dist);
    }
    while(current & CB_LEAF_BIT)
    {
      uint64_t idx = (current ^ CB_LEAF_BIT) >> 5;
      const uint64_t num = current & 31;
      for(uint64_t k=0;k<num;k++)
      {
        if(cnt) cnt[3]++;
        orc_prim_intersect(acc->scene, acc->scene->primid[idx], ray, hit);
        idx++;
      }
      do
      {
        if(stackpos == 0) return;
        --stackpos;
        current = stack[stackpos];
      }
      while(stack_dist[stackpos] > hit->dist);
    }
    node = acc->tree + current;
  }
}

/* qbvhmp.c:1392-1490.  The reference enters at a per-thread cached node and walks up through
 * ->parent until the root; the set of leaves whose boxes the ray touches is the same as a plain
 * top-down sweep from the root, and the result is a boolean, so the restatement sweeps from the
 * root (cache state cannot change the answer).  tmax = max_dist, boxes clipped at 0. */
int orc_visible(const orc_accel_t *acc, const cb_ray_t *ray, float max_dist)
{
  uint64_t stack[3*ORC_MAX_TREE_DEPTH + 4];
  int stackpos = 0;
  const float t0 = 1.0f - ray->time, t1 = ray->time;
  float invdir[3];
  for(int k=0;k<3;k++) invdir[k] = 1.0f/ray->dir[k];
  const cb_qbvh_node_t *node = acc->tree;
  while(1)
  {
    float tmin[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float tmax[4] = {max_dist, max_dist, max_dist, max_dist};
    node_boxes(node, t0, t1, ray->pos, invdir, tmin, tmax);
    for(int c=0;c<4;c++) if(tmin[c] <= tmax[c]) stack[stackpos++] = node->child[c];
    uint64_t current;
    while(1)
    {
      if(stackpos == 0) return 1;
      current = stack[--stackpos];
      if(!(current & CB_LEAF_BIT)) break;
      uint64_t idx = (current ^ CB_LEAF_BIT) >> 5;
      const uint64_t num = current & 31;
      for(uint64_t k=0;k<num;k++, idx++)
        if(orc_prim_visible(acc->scene, acc->scene->primid[idx], ray, max_dist)) return 0;
    }
    node = acc->tree + current;
  }
}

static void hit_reset(cb_hit_t *h, float dist)
{
  memset(h, 0, sizeof(*h));
  h->prim[0] = h->prim[1] = 0xffffffffu;
  h->dist = dist;
}

void orc_intersect_n(const orc_accel_t *a, const cb_ray_t *rays, const float *max_dist, cb_hitrec_t *out,
                     uint64_t n, int nthreads, uint64_t counters[4])
{
  if(nthreads < 1) nthreads = 1;
  uint64_t tot[4] = {0, 0, 0, 0};
#pragma omp parallel num_threads(nthreads)
  {
    uint64_t cnt[4] = {0, 0, 0, 0};
#pragma omp for schedule(dynamic, 4096)
    for(uint64_t i=0;i<n;i++)
    {
      cb_hit_t h;
      hit_reset(&h, max_dist ? max_dist[i] : FLT_MAX);
      orc_intersect(a, rays+i, &h, counters ? cnt : 0);
      out[i].prim[0] = h.prim[0]; out[i].prim[1] = h.prim[1];
      out[i].u = h.u; out[i].v = h.v; out[i].dist = h.dist; out[i].pad = 0;
    }
#pragma omp critical
    for(int k=0;k<4;k++) tot[k] += cnt[k];
  }
  if(counters) for(int k=0;k<4;k++) counters[k] += tot[k];
}

void orc_visible_n(const orc_accel_t *a, const cb_ray_t *rays, const float *max_dist, int32_t *out,
                   uint64_t n, int nthreads)
{
  if(nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4096) num_threads(nthreads)
  for(uint64_t i=0;i<n;i++) out[i] = orc_visible(a, rays+i, max_dist[i]);
}

/* qbvhmp.c:1493-1600: nearest hit to `centre` along the ray for the half-vector samplers.  Boxes are clipped to
 * [ray->min_dist, hit->dist] (the only traversal that uses min_dist for boxes), there is no entry-distance culling on pop,
 * and after EVERY primitive test the search interval is re-centred around `centre` by mutating ray->min_dist / hit->dist;
 * `tent` keeps the hit closest to centre. */
void orc_closest(const orc_accel_t *acc, cb_ray_t *ray, cb_hit_t *hit, float centre)
{
  uint32_t near[3], far[3];
  for(int k=0;k<3;k++) { near[k] = signbit_u(ray->dir[k]); far[k] = 1 ^ near[k]; }
  const cb_qbvh_node_t *node = acc->tree;
  uint64_t stack[3*ORC_MAX_TREE_DEPTH];
  int stackpos = 0;
  uint64_t current;
  cb_hit_t tent = *hit;
  const float t0 = 1.0f - ray->time, t1 = ray->time;
  float invdir[3];
  for(int k=0;k<3;k++) invdir[k] = 1.0f/ray->dir[k];
  while(1)
  {
    float tmin[4] = {ray->min_dist, ray->min_dist, ray->min_dist, ray->min_dist};
    float tmax[4] = {hit->dist, hit->dist, hit->dist, hit->dist};
    node_boxes(node, t0, t1, ray->pos, invdir, tmin, tmax);
    int i[4];
    for(int c=0;c<4;c++) i[c] = tmin[c] <= tmax[c];
    const int axis0  = (int)node->axis0;
    const int axis1n = near[axis0] ? (int)node->axis01 : (int)node->axis00;
    const int axis1f = near[axis0] ? (int)node->axis00 : (int)node->axis01;
    const uint32_t n11 = (far [axis0]<<1) | far [axis1f];
    const uint32_t n10 = (far [axis0]<<1) | near[axis1f];
    const uint32_t n01 = (near[axis0]<<1) | far [axis1n];
    const uint32_t n00 = (near[axis0]<<1) | near[axis1n];
    if(i[n00])
    {
      current = node->child[n00];
      if(i[n11]) stack[stackpos++] = node->child[n11];
      if(i[n10]) stack[stackpos++] = node->child[n10];
      if(i[n01]) stack[stackpos++] = node->child[n01];
    }
    else if(i[n01])
    {
      current = node->child[n01];
      if(i[n11]) stack[stackpos++] = node->child[n11];
      if(i[n10]) stack[stackpos++] = node->child[n10];
    }
    else if(i[n10])
    {
      current = node->child[n10];
      if(i[n11]) stack[stackpos++] = node->child[n11];
    }
    else if(i[n11]) current = node->child[n11];
    else
    {
      if(stackpos == 0) { *hit = tent; return; }
      stackpos--;
      current = stack[stackpos];
    }
    while(current & CB_LEAF_BIT)
    {
      uint64_t idx = (current ^ CB_LEAF_BIT) >> 5;
      const uint64_t num = current & 31;
      for(uint64_t k=0;k<num;k++)
      {
        orc_prim_intersect(acc->scene, acc->scene->primid[idx], ray, hit);
        if(fabsf(hit->dist - centre) < fabsf(tent.dist - centre)) tent = *hit;
        if(hit->dist > centre + 1e-6f) ray->min_dist = 2.0f*centre - hit->dist;
        else if(hit->dist < centre - 1e-6f) { ray->min_dist = hit->dist; hit->dist = 2.0f*centre - ray->min_dist; }
        else return;
        idx++;
      }
      if(stackpos == 0) { *hit = tent; return; }
      --stackpos;
      current = stack[stackpos];
    }
    node = acc->tree + current;
  }
}

/* batch form used by the tests: rays (min_dist is updated in place), hits {prim,u,v,dist} in/out */
void orc_closest_n(const orc_accel_t *a, cb_ray_t *rays, cb_hitrec_t *io, const float *centre, uint64_t n)
{
  for(uint64_t i=0;i<n;i++)
  {
    cb_hit_t h;
    memset(&h, 0, sizeof(h));
    memcpy(h.prim, io[i].prim, 8);
    h.u = io[i].u; h.v = io[i].v; h.dist = io[i].dist;
    orc_closest(a, rays + i, &h, centre[i]);
    memcpy(io[i].prim, h.prim, 8);
    io[i].u = h.u; io[i].v = h.v; io[i].dist = h.dist; io[i].pad = 0;
  }
}

/* ---------------------------------------------------------------------------- build */
/* qbvhmp.c:425-525 (+ split_job_work :325-357): binned "kd" SAH along one axis */
static float get_split_with_dim(orc_accel_t *b, int64_t left_in, int64_t right_in, const float aabb[6], int d, float *split)
{
  float best = (aabb[d] + aabb[3+d])*0.5f;
  float best_score = FLT_MAX;
  if(right_in == left_in) { *split = best; return best_score; }

  int binmin[ORC_SAH_TESTS+1] = {0}, binmax[ORC_SAH_TESTS+1] = {0};
  const int p = d == 2 ? 0 : d+1;
  const int q = d == 0 ? 2 : d-1;
  const int64_t step = (int)(log10f(ORC_SAH_LOG_STEP*(right_in - left_in) + 1.0f) + 1.0f);
  const float jaabb0 = aabb[d], jaabb1 = aabb[3+d];
  {
    const float width = jaabb1 - jaabb0;
    for(int64_t k=left_in;k<right_in;k+=step)
    {
      const float min = b->prim_aabb[6*k + d];
      const float max = b->prim_aabb[6*k + 3 + d];
      const float fmin_ = (ORC_SAH_TESTS+1)*(min - jaabb0)/width;
      const float fmax_ = (ORC_SAH_TESTS+1)*(max - jaabb0)/width;
      const int imin = (int)MINM(MAXM(fmin_, 0), ORC_SAH_TESTS);
      const int imax = (int)MINM(MAXM(fmax_, 0), ORC_SAH_TESTS);
      int cost = 8;
      const uint32_t vcnt = cb_primid_vcnt(b->scene->primid[k]);
      if(vcnt == 4) cost = 16;
      else if(vcnt == 2) cost = 1;
      binmin[imin] += cost;
      binmax[imax] += cost;
    }
  }
  int left = binmin[0];
  int right = 0;
  const float com = (aabb[3+p] - aabb[p])*(aabb[3+q] - aabb[q]);
  const float width = aabb[3+d] - aabb[d];
  for(int i=1;i<ORC_SAH_TESTS+1;i++) right += binmax[i];
  for(int k=0;k<ORC_SAH_TESTS;k++)
  {
    const float splitc = aabb[d] + (k+1)*width/(ORC_SAH_TESTS + 1.0f);
    const float splitwl = splitc - aabb[d];
    const float splitwr = aabb[3+d] - splitc;
    const float stepl = com + (aabb[3+p] - aabb[p])*splitwl + (aabb[3+q] - aabb[q])*splitwl;
    const float stepr = com + (aabb[3+p] - aabb[p])*splitwr + (aabb[3+q] - aabb[q])*splitwr;
    const float score = stepl*left + stepr*right;
    if(score < best_score) { best_score = score; best = splitc; }
    left  += binmin[k+1];
    right -= binmax[k+1];
  }
  *split = best;
  return best_score;
}

/* qbvhmp.c:528-570: in-place partition by centroid >= split; swaps primid and cached boxes */
static uint64_t sort_primids(orc_accel_t *b, int axis, float split, int64_t begin, int64_t back, float *aabbl, float *aabbr)
{
  int64_t right = back;
  uint64_t *primid = b->scene->primid;
  for(int k=0;k<3;k++) aabbl[k] = aabbr[k] =  FLT_MAX;
  for(int k=3;k<6;k++) aabbl[k] = aabbr[k] = -FLT_MAX;
  for(int64_t i=begin;i<right;)
  {
    const float pmin = b->prim_aabb[6*i + axis];
    const float pmax = b->prim_aabb[6*i + 3 + axis];
    if(.5f*(pmin + pmax) >= split)
    {
      for(int k=0;k<3;k++) aabbr[k] = MINM(aabbr[k], b->prim_aabb[6*i+k]);
      for(int k=3;k<6;k++) aabbr[k] = MAXM(aabbr[k], b->prim_aabb[6*i+k]);
      --right;
      const uint64_t tmp = primid[i]; primid[i] = primid[right]; primid[right] = tmp;
      for(int k=0;k<6;k++)
      {
        const float t = b->prim_aabb[6*i+k];
        b->prim_aabb[6*i+k] = b->prim_aabb[6*right+k];
        b->prim_aabb[6*right+k] = t;
      }
    }
    else
    {
      for(int k=0;k<3;k++) aabbl[k] = MINM(aabbl[k], b->prim_aabb[6*i+k]);
      for(int k=3;k<6;k++) aabbl[k] = MAXM(aabbl[k], b->prim_aabb[6*i+k]);
      i++;
    }
  }
  return right;
}

/* qbvhmp.c:854-873 */
static void bound_leaf_t1(orc_accel_t *b, cb_qbvh_node_t *node, int c)
{
  for(int k=0;k<3;k++) node->aabb1[k][c] =  FLT_MAX;
  for(int k=3;k<6;k++) node->aabb1[k][c] = -FLT_MAX;
  const uint64_t num_prims = node->child[c] & 31;
  const uint64_t prims = (node->child[c] ^ CB_LEAF_BIT) >> 5;
  for(uint64_t k=prims;k<prims+num_prims;k++)
    for(int d=0;d<3;d++)
    {
      float m, M;
      prim_bounds_dim(b->scene, b->scene->primid[k], 1, d, &m, &M);
      node->aabb1[d][c]   = MINM(node->aabb1[d][c], m);
      node->aabb1[3+d][c] = MAXM(node->aabb1[3+d][c], M);
    }
}

/* the reference's one-thread scheduling (qbvhmp.c:371-423, 998-1018): a node job is queued when the
 * worker's queue is empty, otherwise the child is built by direct recursion.  Reproducing that order
 * reproduces the reference's node numbering for `-t 1`. */
typedef struct { int64_t node, parent; int64_t left, right; float paabb[6]; int depth, child; } orc_job_t;
typedef struct { orc_job_t *jobs; int num, cap; } orc_queue_t;

static uint64_t node_work(orc_accel_t *b, orc_queue_t *q, int64_t node_i, int64_t left, int64_t right,
                          const float *paabb_in, int depth, int64_t parent_i, int child)
{ /* qbvhmp.c:875-1022 */
  cb_qbvh_node_t *node = b->tree + node_i;
  cb_qbvh_node_t *parent = parent_i >= 0 ? b->tree + parent_i : 0;
  float paabb[6];
  memcpy(paabb, paabb_in, sizeof(paabb));
  node->parent = (uint64_t)parent_i;
  int axis0 = 0, axis00, axis01;
  uint64_t part[5];
  float split0, split1l, split1r, score, split;
  float aabb[4][6];
  part[0] = left;
  part[4] = right;
  float best_score = get_split_with_dim(b, left, right, paabb, 0, &split0);
  if((score = get_split_with_dim(b, left, right, paabb, 1, &split)) < best_score) { best_score = score; split0 = split; axis0 = 1; }
  if((score = get_split_with_dim(b, left, right, paabb, 2, &split)) < best_score) { best_score = score; split0 = split; axis0 = 2; }

  if(best_score < 0.0f && (right - left) < 32)
  {
    if(parent)
    {
      parent->child[child] = CB_LEAF_BIT | ((uint64_t)left<<5) | ((uint64_t)(right - left) & 31);
      bound_leaf_t1(b, parent, child);
      return right - left;
    }
  }

  part[2] = sort_primids(b, axis0, split0, left, right, aabb[0], aabb[1]);

  axis00 = 0;
  best_score = get_split_with_dim(b, left, part[2], aabb[0], 0, &split1l);
  if((score = get_split_with_dim(b, left, part[2], aabb[0], 1, &split)) < best_score) { best_score = score; split1l = split; axis00 = 1; }
  if((score = get_split_with_dim(b, left, part[2], aabb[0], 2, &split)) < best_score) { split1l = split; axis00 = 2; }
  axis01 = 0;
  best_score = get_split_with_dim(b, part[2], right, aabb[1], 0, &split1r);
  if((score = get_split_with_dim(b, part[2], right, aabb[1], 1, &split)) < best_score) { best_score = score; split1r = split; axis01 = 1; }
  if((score = get_split_with_dim(b, part[2], right, aabb[1], 2, &split)) < best_score) { split1r = split; axis01 = 2; }

  part[1] = sort_primids(b, axis00, split1l, left, part[2], aabb[0], aabb[1]);
  part[3] = sort_primids(b, axis01, split1r, part[2], right, aabb[2], aabb[3]);

  const uint64_t all = (uint64_t)(right - left);
  if((right - part[3] == all) || (part[3] - part[2] == all) || (part[2] - part[1] == all) || (part[1] - left == all))
  {
    if((parent && (right - left > ORC_PRIMS_PER_LEAF)) || !parent)
    { /* median thirds fallback */
      part[2] = (right   + left)/2;
      part[3] = (right   + part[2])/2;
      part[1] = (part[2] + left)/2;
      for(int k=0;k<3;k++) for(int i=0;i<4;i++) { aabb[i][k] = FLT_MAX; aabb[i][k+3] = -FLT_MAX; }
      for(int p=0;p<4;p++)
        for(uint64_t k=part[p];k<part[p+1];k++)
          for(int i=0;i<3;i++)
          {
            const float min = b->prim_aabb[6*k+i];
            const float max = b->prim_aabb[6*k+3+i];
            if(min < aabb[p][i  ]) aabb[p][i  ] = min;
            if(max > aabb[p][i+3]) aabb[p][i+3] = max;
          }
    }
    else
    {
      parent->child[child] = CB_LEAF_BIT | ((uint64_t)left<<5) | ((uint64_t)(right - left) & 31);
      bound_leaf_t1(b, parent, child);
      return right - left;
    }
  }

  node->axis0  = axis0;
  node->axis00 = axis00;
  node->axis01 = axis01;
  for(int k=0;k<6;k++) for(int p=0;p<4;p++) node->aabb0[k][p] = aabb[p][k];

  uint64_t done = 0;
  int childcnt = 0;
  for(int p=0;p<4;p++)
    if(!(depth == ORC_MAX_TREE_DEPTH || part[p+1] - part[p] <= ORC_PRIMS_PER_LEAF || b->num_nodes >= b->node_bufsize - 1))
      childcnt++;
  uint64_t num_nodes = b->num_nodes;
  b->num_nodes += childcnt;
  for(int p=0;p<4;p++)
  {
    if(depth == ORC_MAX_TREE_DEPTH || part[p+1] - part[p] <= ORC_PRIMS_PER_LEAF || b->num_nodes >= b->node_bufsize - 1)
    {
      node->child[p] = CB_LEAF_BIT | (part[p]<<5) | ((part[p+1] - part[p]) & 31);
      bound_leaf_t1(b, node, p);
      done += part[p+1]-part[p];
    }
    else
    {
      node->child[p] = num_nodes++;
      if(q->num >= 1)
        done += node_work(b, q, (int64_t)node->child[p], part[p], part[p+1], aabb[p], depth + 1, node_i, p);
      else
      {
        orc_job_t *job = q->jobs + q->num++;
        job->node = (int64_t)node->child[p];
        job->left = part[p]; job->right = part[p+1];
        for(int k=0;k<6;k++) job->paabb[k] = aabb[p][k];
        job->depth = depth + 1;
        job->parent = node_i;
        job->child = p;
      }
    }
  }
  return done;
}

/* qbvhmp.c:259-283 */
static void refit(orc_accel_t *b, cb_qbvh_node_t *node)
{
  if(b->scene->num_prims == 0) return;
  for(int c=0;c<4;c++)
    if(!(node->child[c] & CB_LEAF_BIT))
    {
      cb_qbvh_node_t *child = b->tree + node->child[c];
      refit(b, child);
      for(int d=0;d<6;d++) node->aabb1[d][c] = child->aabb1[d][0];
      for(int k=1;k<4;k++)
        for(int d=0;d<3;d++)
        {
          node->aabb1[d][c]   = MINM(node->aabb1[d][c],   child->aabb1[d][k]);
          node->aabb1[3+d][c] = MAXM(node->aabb1[3+d][c], child->aabb1[3+d][k]);
        }
    }
}

static orc_accel_t *accel_alloc(orc_scene_t *s, uint64_t bufsize)
{
  orc_accel_t *b = calloc(1, sizeof(*b));
  b->scene = s;
  b->node_bufsize = bufsize;
  b->tree = aligned_alloc(128, bufsize*sizeof(cb_qbvh_node_t));
  memset(b->tree, 0, bufsize*sizeof(cb_qbvh_node_t));
  b->aabb[0] = b->aabb[1] = b->aabb[2] = FLT_MAX;
  b->aabb[3] = b->aabb[4] = b->aabb[5] = -FLT_MAX;
  return b;
}

/* qbvhmp.c:285-323, 1034-1186 */
orc_accel_t *orc_accel_build(orc_scene_t *s)
{
  orc_accel_t *b = accel_alloc(s, s->num_prims > 100 ? s->num_prims : 100);
  b->num_nodes = 1;
  const uint64_t n = s->num_prims;
  if(n == 0)
  { /* qbvhmp.c:1081-1099: root of four empty leaves */
    cb_qbvh_node_t *node = b->tree;
    for(int c=0;c<4;c++)
    {
      for(int k=0;k<3;k++) node->aabb0[k][c] = node->aabb1[k][c] =  FLT_MAX;
      for(int k=3;k<6;k++) node->aabb0[k][c] = node->aabb1[k][c] = -FLT_MAX;
      node->child[c] = CB_LEAF_BIT;
    }
    node->axis0 = 0; node->axis00 = 1; node->axis01 = 1;
    node->parent = (uint64_t)-1;
    return b;
  }
  b->prim_aabb = malloc(sizeof(float)*6*n);
  for(uint64_t i=0;i<n;i++)
    for(int k=0;k<3;k++)
    { /* compute_aabb, qbvhmp.c:1034-1065: shutter-open boxes drive the topology */
      float min, max;
      prim_bounds_dim(s, s->primid[i], 0, k, &min, &max);
      b->prim_aabb[6*i+k] = min;
      b->prim_aabb[6*i+3+k] = max;
      if(b->aabb[k]   > min) b->aabb[k]   = min;
      if(b->aabb[k+3] < max) b->aabb[k+3] = max;
    }
  orc_queue_t q;
  q.cap = 3*ORC_MAX_TREE_DEPTH + 8;
  q.jobs = malloc(sizeof(orc_job_t)*q.cap);
  q.num = 1;
  q.jobs[0] = (orc_job_t){ .node = 0, .parent = -1, .left = 0, .right = (int64_t)n, .depth = 0, .child = 0 };
  memcpy(q.jobs[0].paabb, b->aabb, sizeof(float)*6);
  while(q.num > 0)
  {
    const orc_job_t job = q.jobs[--q.num];
    node_work(b, &q, job.node, job.left, job.right, job.paabb, job.depth, job.parent, job.child);
  }
  free(q.jobs);
  b->tree[0].parent = (uint64_t)-1;
  refit(b, b->tree);
  free(b->prim_aabb);
  b->prim_aabb = 0;
  return b;
}

orc_accel_t *orc_accel_import(orc_scene_t *s, const cb_qbvh_node_t *nodes, uint64_t num_nodes, const float aabb[6])
{
  orc_accel_t *b = accel_alloc(s, num_nodes ? num_nodes : 1);
  memcpy(b->tree, nodes, num_nodes*sizeof(cb_qbvh_node_t));
  b->num_nodes = num_nodes;
  if(aabb) memcpy(b->aabb, aabb, sizeof(float)*6);
  return b;
}

void orc_accel_free(orc_accel_t *a) { if(!a) return; free(a->tree); free(a->prim_aabb); free(a); }
uint64_t orc_accel_num_nodes(const orc_accel_t *a) { return a->num_nodes; }
const cb_qbvh_node_t *orc_accel_nodes(const orc_accel_t *a) { return a->tree; }
const float *orc_accel_aabb(const orc_accel_t *a) { return a->aabb; }

/* ---------------------------------------------------------------------------- tree check */
/* in the spirit of the reference's disabled checktree() (qbvhmp.c:211-257) */
static int check_node(const orc_accel_t *a, uint64_t ni, int depth, uint8_t *seen, uint64_t stats[4])
{
  if(ni >= a->num_nodes) return 2;
  if(depth > 3*ORC_MAX_TREE_DEPTH) return 3;
  const cb_qbvh_node_t *n = a->tree + ni;
  stats[0]++;
  if((uint64_t)depth > stats[2]) stats[2] = depth;
  if(n->axis0 < 0 || n->axis0 > 2 || n->axis00 < 0 || n->axis00 > 2 || n->axis01 < 0 || n->axis01 > 2) return 4;
  for(int c=0;c<4;c++)
  {
    if(n->child[c] & CB_LEAF_BIT)
    {
      const uint64_t num = n->child[c] & 31, beg = (n->child[c] ^ CB_LEAF_BIT) >> 5;
      stats[1]++;
      if(num && beg + num > a->scene->num_prims) return 5;
      for(uint64_t k=beg;k<beg+num;k++)
      {
        if(seen[k]) return 6;
        seen[k] = 1;
        stats[3]++;
        float b0[6], b1[6];
        orc_prim_bounds(a->scene, a->scene->primid[k], 0, b0);
        orc_prim_bounds(a->scene, a->scene->primid[k], 1, b1);
        for(int d=0;d<3;d++)
        {
          if(!(n->aabb0[d][c] <= b0[d]) || !(n->aabb0[d+3][c] >= b0[d+3])) return 7;
          if(!(n->aabb1[d][c] <= b1[d]) || !(n->aabb1[d+3][c] >= b1[d+3])) return 8;
        }
      }
    }
    else
    {
      const uint64_t ci = n->child[c];
      if(ci >= a->num_nodes) return 2;
      const cb_qbvh_node_t *ch = a->tree + ci;
      /* parent box must contain the non-empty boxes of the child node */
      for(int k=0;k<4;k++)
      {
        const int empty = (ch->child[k] & CB_LEAF_BIT) && ((ch->child[k] & 31) == 0);
        if(empty) continue;
        for(int d=0;d<3;d++)
        {
          if(ch->aabb0[d][k] <= ch->aabb0[d+3][k])
            if(!(n->aabb0[d][c] <= ch->aabb0[d][k]) || !(n->aabb0[d+3][c] >= ch->aabb0[d+3][k])) return 9;
          if(ch->aabb1[d][k] <= ch->aabb1[d+3][k])
            if(!(n->aabb1[d][c] <= ch->aabb1[d][k]) || !(n->aabb1[d+3][c] >= ch->aabb1[d+3][k])) return 10;
        }
      }
      const int r = check_node(a, ci, depth+1, seen, stats);
      if(r) return r;
    }
  }
  return 0;
}

int orc_accel_check(const orc_accel_t *a, uint64_t stats[4])
{
  uint64_t st[4] = {0, 0, 0, 0};
  const uint64_t n = a->scene->num_prims;
  uint8_t *seen = calloc(n ? n : 1, 1);
  int r = check_node(a, 0, 0, seen, st);
  if(!r) for(uint64_t k=0;k<n;k++) if(!seen[k]) { r = 11; break; }
  free(seen);
  if(stats) memcpy(stats, st, sizeof(st));
  return r;
}
