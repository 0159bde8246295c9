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
 */
#include "corona_common.h"
#include "pathspace.h"
#include "prims.h"
#include "view.h"
#include <stdlib.h>
#include <string.h>

extern int init(const char *filename, int argc, char *argv[]);   /* src/main.c:250 */

int ref_path_open(const char *nra2, int argc, char **argv)
{
  return init(nra2, argc, argv);
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
