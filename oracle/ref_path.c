/* ref_path.c -- TEST INFRASTRUCTURE.  Drives the UNMODIFIED reference renderer's own path construction one path at a time and
 * hands back what it computed, so that the GPU's camera sampling, first hit and ray offsets have known-answer vectors from the
 * reference itself (tests/golden/make_golden_paths.py -> tests/golden/paths.npz).  Linked by oracle/Makefile with the reference's
 * renderer sources in place (src/main.c compiled with -Dmain=corona_ref_main so that its init() can be reused) into
 * oracle/_ref/libref_path_<pointsampler>.so.  Nothing in the product uses this.
 *
 *   ref_path_open(nra2, argc, argv)   = main.c's init(): view_init (camera, frame size), shader_init, common_load_scene,
 *                                       pointsampler_init(--frame), accel_build -- the reference's own start-up
 *   ref_path_camera(first, n, out)    for path index i: path_init + path_extend (pathspace.c:167-260: lambda, time, camera_sample
 *                                       of the thin lens, view_cam_init_frame, path_propagate -> accel_intersect)
 *   ref_path_offset(x, dir, prim, n, out)   prims_offset_ray (src/prims.c:374-388) on given hit points / directions
 *   ref_path_bounce(first, n, scr, nee, out)   for path index i: path_init, tangent_frame_scrambling preset to `scr` (upstream draws it
 *                                     from the worker's own twister, pathspace.c:212-213), path_extend to the first hit, optionally what
 *                                     ptdl.c does there (nee_sample + path_pop, which folds four random dimensions into the vertex,
 *                                     pathspace.c:298), then path_extend AGAIN: the BSDF-sampled direction, the second hit, pdf, throughput
 *   ref_path_nee(first, n, out)       for path index i: path_init + path_extend, then -- like sampler.d/ptdl.c:139-147 does at every
 *                                     vertex -- nee_sample (include/pathspace/nee.h:87-243: lights_pdf_type, the light list's
 *                                     sample_cdf + prims_sample, shader_brdf, path_G, path_visible) and the weight of
 *                                     sampler_mis (ptdl.c:78-88) against path_pdf_extend
 */
#include "corona_common.h"
#include "pathspace.h"
#include "prims.h"
#include "view.h"
#include "pathspace/nee.h"
#include <stdlib.h>
#include <string.h>

extern int init(const char *filename, int argc, char *argv[]);   /* src/main.c:250 */

int ref_path_open(const char *nra2, int argc, char **argv)
{
  const int rc = init(nra2, argc, argv);
  if(!rc) lights_prepare_frame();   /* what view_render does before the first sample (src/view.c:640): light list cdf, p_sky / p_geo */
  return rc;
}

/* out: n rows of 20 floats:
 *  0 pixel_i, 1 pixel_j, 2 lambda, 3 time, 4..6 v[0].hit.x (point on the lens), 7..9 e[1].omega, 10 e[1].dist,
 *  11 v[0].throughput, 12 v[1].throughput (path_extend's return state), 13,14 v[1].hit.prim (bit pattern), 15 v[1].hit.u, 16 v[1].hit.v,
 *  17 path->length after the call, 18 path_extend's return value, 19 v[1].pdf */
void ref_path_camera(uint64_t first, uint64_t n, float *out)
{
  path_t *p = (path_t *)malloc(sizeof(path_t));
  for(uint64_t i=0;i<n;i++)
  {
    float *o = out + 20*i;
    path_init(p, first + i, 0);
    const int rc = path_extend(p);
    o[0] = p->sensor.pixel_i; o[1] = p->sensor.pixel_j;
    o[2] = mf(p->lambda, 0); o[3] = p->time;
    for(int k=0;k<3;k++) { o[4+k] = p->v[0].hit.x[k]; o[7+k] = p->e[1].omega[k]; }
    o[10] = p->e[1].dist;
    o[11] = mf(p->v[0].throughput, 0); o[12] = mf(p->v[1].throughput, 0);
    memcpy(o + 13, &p->v[1].hit.prim, 8);
    o[15] = p->v[1].hit.u; o[16] = p->v[1].hit.v;
    o[17] = (float)p->length; o[18] = (float)rc; o[19] = mf(p->v[1].pdf, 0);
  }
  free(p);
}

/* prims_offset_ray for n (hit point, direction) pairs: out = n rows of {pos[3], min_dist} */
void ref_path_offset(const float *x, const float *dir, uint64_t n, float *out)
{
  for(uint64_t i=0;i<n;i++)
  {
    hit_t hit;
    ray_t ray;
    memset(&hit, 0, sizeof(hit));
    memset(&ray, 0, sizeof(ray));
    for(int k=0;k<3;k++) { hit.x[k] = x[3*i+k]; ray.dir[k] = dir[3*i+k]; }
    ray.min_dist = 7.0f;
    prims_offset_ray(&hit, &ray);
    for(int k=0;k<3;k++) out[4*i+k] = ray.pos[k];
    out[4*i+3] = ray.min_dist;
  }
}

/* out: n rows of 20 floats:
 *  0 pixel_i, 1 pixel_j, 2 lambda, 3 path->length after path_extend (2 = camera + first hit), 4 nee_sample's return value (-1: not called),
 *  5 path->length after nee_sample, 6 path_throughput, 7 v[2].pdf, 8 path_pdf_extend(path, 2), 9 mis weight, 10,11 v[2].hit.prim (bit
 *  pattern), 12..14 v[2].hit.x, 15..17 e[2].omega, 18 e[2].dist, 19 v[2].mode */
void ref_path_nee(uint64_t first, uint64_t n, float *out)
{
  path_t *p = (path_t *)malloc(sizeof(path_t));
  for(uint64_t i=0;i<n;i++)
  {
    float *o = out + 20*i;
    memset(o, 0, 20*sizeof(float));
    path_init(p, first + i, 0);
    const int rc = path_extend(p);
    o[0] = p->sensor.pixel_i; o[1] = p->sensor.pixel_j; o[2] = mf(p->lambda, 0);
    o[3] = (float)p->length; o[4] = -1.0f;
    if(rc || p->length != 2) continue;
    const int rn = nee_sample(p);
    o[4] = (float)rn; o[5] = (float)p->length;
    if(rn || p->length != 3) continue;
    const int v2 = p->length - 1;
    const float thr = mf(path_throughput(p), 0);
    o[6] = thr; o[7] = mf(p->v[v2].pdf, 0); o[19] = (float)p->v[v2].mode;
    if(thr > 0.0f && (p->v[v2].mode & s_emit))
    { /* (a failed nee_sample leaves parts of v[2] / e[2] as the previous path left them: only contributing samples are recorded) */
      memcpy(o + 10, &p->v[v2].hit.prim, 8);
      for(int k=0;k<3;k++) { o[12+k] = p->v[v2].hit.x[k]; o[15+k] = p->e[v2].omega[k]; }
      o[18] = p->e[v2].dist;
      /* sampler_mis (static in ptdl.c:78-88), one wavelength: both pdfs times the product of v[1..length-2].pdf in double, back
       * to float, our / (other + our) */
      const float pe = mf(path_pdf_extend(p, v2), 0);
      double pdf_path = 1.0;
      for(int v=1;v<p->length-1;v++) pdf_path *= (double)mf(p->v[v].pdf, 0);
      const double our = (double)o[7]*pdf_path, other = (double)pe*pdf_path;
      o[8] = pe;
      o[9] = (float)our/(float)(other + our);
    }
  }
  free(p);
}

/* out: n rows of 20 floats:
 *  0 pixel_i, 1 pixel_j, 2 lambda, 3 path->length after the first path_extend, 4 return value of the second path_extend (-1: not
 *  called), 5 path->length after it, 6..8 e[2].omega, 9 e[2].dist, 10,11 v[2].hit.prim (bit pattern), 12..14 v[1].hit.x,
 *  15 v[2].throughput, 16 v[2].pdf, 17 v[1].mode after sampling, 18 v[2].flags, 19 v[1].throughput */
void ref_path_bounce(uint64_t first, uint64_t n, float scrambling, int with_nee, float *out)
{
  path_t *p = (path_t *)malloc(sizeof(path_t));
  for(uint64_t i=0;i<n;i++)
  {
    float *o = out + 20*i;
    memset(o, 0, 20*sizeof(float));
    path_init(p, first + i, 0);
    p->tangent_frame_scrambling = scrambling;
    const int rc = path_extend(p);
    o[0] = p->sensor.pixel_i; o[1] = p->sensor.pixel_j; o[2] = mf(p->lambda, 0);
    o[3] = (float)p->length; o[4] = -1.0f;
    if(rc || p->length != 2) continue;
    o[19] = mf(p->v[1].throughput, 0);
    if(with_nee)
    { /* sampler.d/ptdl.c:136-148 */
      if(nee_sample(p)) continue;
      path_pop(p);
    }
    const int r2 = path_extend(p);
    o[4] = (float)r2; o[5] = (float)p->length;
    for(int k=0;k<3;k++) o[12+k] = p->v[1].hit.x[k];
    o[17] = (float)p->v[1].mode;
    if(p->length < 3) continue;
    for(int k=0;k<3;k++) o[6+k] = p->e[2].omega[k];
    o[9] = p->e[2].dist;
    memcpy(o + 10, &p->v[2].hit.prim, 8);
    o[15] = mf(p->v[2].throughput, 0); o[16] = mf(p->v[2].pdf, 0); o[18] = (float)p->v[2].flags;
  }
  free(p);
}

/* sampler.d/ptdl.c's (or pt.c's) loop up to the second vertex, recording what it would splat for emission found by EXTENSION:
 * out: n rows of 8 floats: 0 pixel_i, 1 pixel_j, 2 lambda, 3 value at v[1] (throughput x weight; 0: not an emitter), 4 value at v[2],
 * 5 mis weight at v[2], 6 path->length at the end, 7 v[2].pdf */
static float ref_mis(const path_t *p, const float pdf, const float pdf2)
{ /* sampler_mis (static in ptdl.c:78-88), one wavelength */
  double pdf_path = 1.0;
  for(int v=1;v<p->length-1;v++) pdf_path *= (double)mf(p->v[v].pdf, 0);
  const double our = (double)pdf*pdf_path, other = (double)pdf2*pdf_path;
  return (float)our/(float)(other + our);
}
void ref_path_emission(uint64_t first, uint64_t n, float scrambling, int ptdl, float *out)
{
  path_t *p = (path_t *)malloc(sizeof(path_t));
  for(uint64_t i=0;i<n;i++)
  {
    float *o = out + 8*i;
    memset(o, 0, 8*sizeof(float));
    path_init(p, first + i, 0);
    p->tangent_frame_scrambling = scrambling;
    const int rc = path_extend(p);
    o[0] = p->sensor.pixel_i; o[1] = p->sensor.pixel_j; o[2] = mf(p->lambda, 0); o[6] = (float)p->length;
    if(rc) continue;
    int v = p->length - 1;
    if(p->v[v].mode & s_emit)
    { /* ptdl.c:124-127 / pt.c:44-47 (pt's sampler_mis is 1 for a single technique) */
      const float w = ptdl ? ref_mis(p, mf(p->v[v].pdf, 0), mf(nee_pdf(p, v), 0)) : 1.0f;
      o[3] = mf(path_throughput(p), 0)*w;
    }
    if(ptdl)
    {
      if(nee_sample(p)) continue;
      path_pop(p);
    }
    if(path_extend(p)) { o[6] = (float)p->length; continue; }
    v = p->length - 1;
    o[6] = (float)p->length; o[7] = mf(p->v[v].pdf, 0);
    if(p->v[v].mode & s_emit)
    {
      const float w = ptdl ? ref_mis(p, mf(p->v[v].pdf, 0), mf(nee_pdf(p, v), 0)) : 1.0f;
      o[4] = mf(path_throughput(p), 0)*w; o[5] = w;
    }
  }
  free(p);
}
