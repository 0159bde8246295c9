/* ref_bsdf.c -- TEST INFRASTRUCTURE.  Calls the reference's own BSDF callbacks (the dlopen'ed shader modules
 * shaders/lib<name>.so and the built-in diffuse of src/shader.c:157-257) one query at a time, with the path set up exactly
 * like the reference's tools/battle-test.c:57-140 does (which no longer runs: SURVEY F5).  Linked by oracle/Makefile with the
 * unmodified reference sources in place into oracle/_ref/libref_bsdf.so.  Used by tests/golden/make_golden_bsdf.py to record
 * known-answer vectors for cb200_render_bsdf and by nothing in the product. */
#include "corona_common.h"
#include "pathspace.h"
#include "shader.h"
#include "points.h"
#include "spectrum.h"
#include "pathspace/manifold.h"
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

rt_t rt;
__thread rt_tls_t rt_tls;

/* the random dimensions sample() asks for come from here (battle-test relies on the same trick through the point
 * sampler's fake_randoms_t, src/pointsampler.d/halton.c:30-35,72-73) */
static float g_rand[64];
float pointsampler(path_t *p, int i) { (void)p; return g_rand[i & 63]; }

/* symbols only src/main.c defines */
void main_screenshot() {}

#include "cb_bsdf_abi.h"

extern mf_t sample_d(path_t *p, void *data);
extern mf_t brdf_d(path_t *p, int v, void *data);
extern mf_t pdf_d(const path_t *p, int e1, int v, int e2, void *data);
extern float prepare_d(path_t *p, int v, void *data);

#define MAX_SHADERS 16
static shader_so_t g_shader[MAX_SHADERS];
static int g_num = 0;

/* so_path == NULL or "diffuse": the built-in diffuse shader.  init_line: the rest of the shader's .nra2 line */
int ref_bsdf_open(const char *so_path, const char *init_line)
{
  if(g_num >= MAX_SHADERS) return -1;
  shader_so_t *s = g_shader + g_num;
  memset(s, 0, sizeof(*s));
  if(!rt.points) { rt.num_threads = 1; rt.points = points_init(1, 1); }
  if(!so_path || !strcmp(so_path, "diffuse"))
  {
    s->sample = sample_d; s->brdf = brdf_d; s->pdf = pdf_d; s->prepare = prepare_d;
    return g_num++;
  }
  void *h = dlopen(so_path, RTLD_LAZY | RTLD_LOCAL);
  if(!h) { fprintf(stderr, "[ref_bsdf] %s\n", dlerror()); return -1; }
  s->sample  = (sample_t)dlsym(h, "sample");
  s->brdf    = (brdf_t)dlsym(h, "brdf");
  s->pdf     = (pdf_t)dlsym(h, "pdf");
  s->prepare = (prepare_t)dlsym(h, "prepare");
  s->init    = (init_t)dlsym(h, "init");
  if((!s->sample || !s->brdf || !s->pdf) && !s->prepare) return -1;   /* prepare-only modules (color) are fine */
  if(s->init)
  {
    char line[1024];
    snprintf(line, sizeof(line), "%s # x\n\n", init_line ? init_line : "");
    FILE *f = fmemopen(line, strlen(line), "r");
    const int rc = s->init(f, &s->data);
    fclose(f);
    if(rc) return -1;
  }
  return g_num++;
}

void ref_bsdf_eval(int handle, const cb_bsdf_query_t *q, cb_bsdf_result_t *out, uint64_t n)
{
  shader_so_t *s = g_shader + handle;
  static path_t path;
  for(uint64_t i=0;i<n;i++)
  {
    const cb_bsdf_query_t *Q = q + i;
    path_init(&path, 0, 0);
    path.lambda = Q->lambda;
    path.length = 2;                       /* now sampling v[2] */
    hit_t *hit = &path.v[1].hit;
    memset(hit, 0, sizeof(*hit));
    hit->n[2] = hit->gn[2] = Q->flip ? -1.0f : 1.0f;
    hit->prim.extra = 0; hit->prim.shapeid = 0; hit->prim.vi = 0; hit->prim.mb = 0; hit->prim.vcnt = 2;
    path.v[0].hit.prim = INVALID_PRIMID;
    path.v[0].mode = s_sensor;
    path_volume_vacuum(&path.e[0].vol);
    path.v[0].interior = path.e[0].vol;
    path.e[1].vol = path.e[0].vol;
    path.v[1].interior = path.e[0].vol;
    vertex_shading_t *sh = &path.v[1].shading;
    sh->rs = Q->rs; sh->rd = Q->rd; sh->em = 0.0f; sh->rg = Q->rg; sh->roughness = Q->roughness;
    get_onb(hit->n, hit->a, hit->b);
    for(int k=0;k<3;k++) path.e[1].omega[k] = Q->wi[k];
    g_rand[s_dim_omega_x] = Q->rand[0]; g_rand[s_dim_omega_y] = Q->rand[1]; g_rand[s_dim_scatter_mode] = Q->rand[2];

    if(s->prepare) s->prepare(&path, 1, s->data);
    path.v[1].diffgeo.eta = path_eta_ratio(&path, 1);     /* shader_prepare caches it (shader.c:538) */
    path.e[2].vol = path.e[1].vol;
    path.v[1].mode = s_absorb;
    path.v[2].pdf = 1.0f;                                  /* path_extend's initialisation (pathspace.c:186-190) */
    const vertex_t keep = path.v[1];
    out[i].s_weight = s->sample(&path, s->data);
    for(int k=0;k<3;k++) out[i].s_wo[k] = path.e[2].omega[k];
    out[i].s_pdf = path.v[2].pdf;
    out[i].s_mode = path.v[1].mode;

    path.v[1] = keep;
    path.v[1].mode = s_absorb;
    for(int k=0;k<3;k++) path.e[2].omega[k] = Q->wo[k];
    out[i].f = s->brdf(&path, 1, s->data);
    out[i].f_mode = path.v[1].mode;
    out[i].pdf = s->pdf(&path, 1, 1, 2, s->data);
  }
}

/* ---- homogeneous media: the reference's medium_rgb module behind an optional `color v' step, its free-flight sampling and
 * edge terms (src/shader.c:46-131) and the phase function callbacks at a volume vertex (src/shaders/medium_rgb.c:62-103).
 * coeff_file: the rgb2spec model `color' / `medium_rgb' fetch their coefficients from at init (src/main.c:292-293). */
int ref_medium_setup(const char *coeff_file)
{
  if(!rt.rgb2spec) rt.rgb2spec = rgb2spec_init(coeff_file);
  if(!rt.shader)
  {
    rt.shader = calloc(1, sizeof(shader_t));
    rt.shader->shader = g_shader;
    rt.shader->exterior_medium_shader = -1;
  }
  return rt.rgb2spec ? 0 : -1;
}

/* h_medium: handle of a medium_rgb, h_albedo: handle of the `color v ...' that precedes it in its mult chain, or -1 */
void ref_medium_eval(int h_medium, int h_albedo, const cb_medium_query_t *q, cb_medium_result_t *out, uint64_t n)
{
  shader_so_t *s = g_shader + h_medium;
  static path_t path;
  rt.shader->num_shaders = g_num;
  for(uint64_t i=0;i<n;i++)
  {
    const cb_medium_query_t *Q = q + i;
    path_init(&path, 0, 0);
    path.lambda = Q->lambda;
    path.tangent_frame_scrambling = 0.5f;
    path.v[0].hit.prim = INVALID_PRIMID;
    path.v[0].mode = s_sensor;
    path.v[1].hit.prim = INVALID_PRIMID;
    /* what mult.prepare leaves in vertex.interior (mult.c:154-167): vacuum, pre steps, host */
    path_volume_vacuum(&path.v[1].interior);
    if(h_albedo >= 0) g_shader[h_albedo].prepare(&path, 1, g_shader[h_albedo].data);
    s->prepare(&path, 1, s->data);
    path.v[1].interior.shader = h_medium;
    path.e[1].vol = path.v[1].interior;
    out[i].mu_t = mf(path.e[1].vol.mu_t, 0);
    out[i].mu_s = mf(path.e[1].vol.mu_s, 0);
    for(int k=0;k<3;k++) path.e[1].omega[k] = Q->wi[k];
    /* distance sampling of edge 1 as path_propagate does it (pathspace.c:742-747) */
    path.length = 1;
    g_rand[s_dim_free_path] = Q->rand[2];
    path.e[1].dist = FLT_MAX;
    out[i].free_dist = shader_vol_sample(&path, 1);
    out[i].free_pdf = mf(path.e[1].pdf, 0);
    /* an edge of length dist that ends on this volume vertex: transmittance and distance pdf (shader.c:46-72,107-131) */
    path.e[1].dist = Q->dist;
    out[i].transmittance = mf(shader_vol_transmittance(&path, 1), 0);
    out[i].vol_pdf = mf(shader_vol_pdf(&path, 1), 0);
    /* phase function at the volume vertex */
    path.length = 2;
    manifold_init(&path, 1);
    path.v[1].material_modes = s_volume | s_glossy;
    path.v[1].mode = s_absorb;
    path.e[2].vol = path.e[1].vol;
    path.v[2].pdf = 1.0f;
    g_rand[s_dim_omega_x] = Q->rand[0]; g_rand[s_dim_omega_y] = Q->rand[1];
    const vertex_t keep = path.v[1];
    out[i].s_weight = mf(s->sample(&path, s->data), 0);
    for(int k=0;k<3;k++) out[i].s_wo[k] = path.e[2].omega[k];
    out[i].s_pdf = mf(path.v[2].pdf, 0);
    out[i].s_mode = path.v[1].mode;
    path.v[1] = keep;
    for(int k=0;k<3;k++) path.e[2].omega[k] = Q->wo[k];
    out[i].f = mf(s->brdf(&path, 1, s->data), 0);
    out[i].f_mode = path.v[1].mode;
    out[i].pdf = mf(s->pdf(&path, 1, 1, 2, s->data), 0);
  }
}
