/* render_b200.c -- MOD_render=b200: the reference's render.h module (include/render.h:13-32) on libcorona_b200.so.
 *
 * The reference's progression is view_render() fanning work_sample() out to the pinned pool, every worker pulling path
 * indices from one counter and calling render_sample_path(i) (src/view.c:618-645, src/render.d/gi.c:81-88).  One path per
 * synchronous call cannot feed a GPU, so this module adds the batched entry the new view_render branch calls instead
 * (SURVEY 8b):
 *
 *     render_b200_pass(r, first_index, count, fb)      == for(i in [first, first+count)) render_sample_path(i);
 *
 * with `fb` the host framebuffer of the view (W*H*3 floats, un-gained like fb->fb, include/framebuffer.h:19-36); it
 * receives the accumulated image of the pass asynchronously, while the next pass renders (NULL: keep it on the device, e.g.
 * between the passes of a --batch).
 * Progressions are streamed (paths that outlive their progression finish during the next one); render_b200_finish()
 * completes them -- call it where the reference saves a screenshot or exits (src/main.c: main_screenshot / cleanup).
 * render_sample_path() itself is exported for link compatibility and refuses loudly: there is no CPU path in here.
 */
#include "corona_host.h"
#include "corona_b200.h"
#include "corona_b200_render.h"

#include <stdlib.h>
#include <string.h>

#ifdef CORONA_B200_IN_TREE
/* inside the reference tree (src/render.d/b200.c): the scene comes from the global rt */
#include "render.h"
#include "view.h"
#include "camera.h"
#include "pathspace.h"
/* the modules the binary is built with, as the Makefile selects them (MOD_sampler, MOD_pointsampler, COL_camera); the recipe
 * passes them on: -DCB200_SAMPLER=CB_SAMPLER_PTDL -DCB200_POINTS=CB_POINTS_RAND -DCB200_COLOUR=CB_COLOUR_XYZ */
#ifndef CB200_SAMPLER
#define CB200_SAMPLER CB_SAMPLER_PTDL
#endif
#ifndef CB200_POINTS
#define CB200_POINTS CB_POINTS_RAND
#endif
#ifndef CB200_COLOUR
#define CB200_COLOUR CB_COLOUR_XYZ
#endif
#endif

struct render_t
{
  cb200_render_t *r;
  uint32_t width, height;
  uint64_t overlays;     /* progressions accumulated so far (view->overlays) */
  struct scene_b200_t *shaders;   /* in-tree: the flattened shader list the device object borrows from */
  int dbor_set;
};

static struct render_t *g_render = 0;   /* the reference reaches its modules through the global rt (corona_common.h:72-106) */

void render_print_info(FILE *fd)
{
  fprintf(fd, "render   : b200 wavefront path tracer, global illumination on the gpu (%s)\n", cb200_version());
}

struct render_t *render_b200_init(const accel_t *accel, const cb_render_desc_t *desc)
{
  if(!accel || !desc) { fprintf(stderr, "[render b200] init: no accel / description\n"); return 0; }
  cb200_accel_t *a = (cb200_accel_t *)accel_b200_handle(accel);
  if(!a) { fprintf(stderr, "[render b200] init: the accel has not been built\n"); return 0; }
  struct render_t *r = (struct render_t *)calloc(1, sizeof(*r));
  r->r = cb200_render_create(a, desc);
  if(!r->r)
  {
    fprintf(stderr, "[render b200] init failed: %s. there is no cpu fallback in this module.\n", cb200_last_error());
    free(r);
    return 0;
  }
  r->width = desc->width; r->height = desc->height;
  g_render = r;
  return r;
}

#ifdef CORONA_B200_IN_TREE
/* The reference's argument-less constructor runs BEFORE the shader list, the geometry and the accel exist (src/main.c:301-347:
 * view_init, render_init, lights_init, shader_init, ..., common_load_scene, accel_build), so it only makes the handle; the
 * device object is created by the first progression (intree_create), when rt.view, rt.accel and rt.anim_frame are final. */
struct render_t *render_init()
{
  struct render_t *r = (struct render_t *)calloc(1, sizeof(*r));
  g_render = r;
  return r;
}

/* cb_render_desc_t from the reference's state:
 *   width, height      view_width() / view_height() (already padded to 32, src/view.c:295-296)
 *   camera             view_get_camera() of camera 0: include/camera.h:13-35
 *   frame              rt.anim_frame (seeds halton_init_random / the counter generator)
 *   sampler, points, colour  the modules this binary was built with (CB200_* above)
 *   materials, tables, media, sky   the same `.nra2` text shader_init() read (rt.basename + ".nra2"), flattened by
 *                      host/scene_b200.c -- the shader modules keep their parsed arguments in private structs behind
 *                      dlopen'ed callbacks (src/shader.c:694-757), the text is the public form of the same data
 *   geometry           rt.accel (MOD_accel=b200: the prims_t the reference loaded, uploaded by accel_build)                   */
static int intree_create(struct render_t *r)
{
  if(!rt.accel || !accel_b200_handle(rt.accel))
  { fprintf(stderr, "[render b200] MOD_render=b200 needs MOD_accel=b200 and a built accel; there is no cpu fallback\n"); return 1; }
  char nra2[1100];
  snprintf(nra2, sizeof(nra2), "%s.nra2", rt.basename);
  r->shaders = scene_b200_open_shaders(nra2, "data/ergb2spec.coeff", getenv("CORONA_B200_TABLES"));
  if(!r->shaders) { fprintf(stderr, "[render b200] could not flatten the shader list of %s\n", nra2); return 1; }
  cb_render_desc_t d;
  memset(&d, 0, sizeof(d));
  d.width = (uint32_t)view_width(); d.height = (uint32_t)view_height();
  path_t *p = (path_t *)calloc(1, sizeof(path_t));      /* view_get_camera() wants a path for its camera id: camera 0 */
  const camera_t *c = view_get_camera(p);
  memcpy(d.camera.pos, c->pos, sizeof(float)*3); memcpy(d.camera.pos_t1, c->pos_t1, sizeof(float)*3);
  d.camera.orient[0] = c->orient.w;       memcpy(d.camera.orient + 1, c->orient.x, sizeof(float)*3);
  d.camera.orient_t1[0] = c->orient_t1.w; memcpy(d.camera.orient_t1 + 1, c->orient_t1.x, sizeof(float)*3);
  d.camera.focus = c->focus; d.camera.film_width = c->film_width; d.camera.film_height = c->film_height;
  d.camera.aperture_value = c->aperture_value; d.camera.exposure_value = c->exposure_value;
  d.camera.focal_length = c->focal_length; d.camera.iso = c->iso;
  free(p);
  scene_b200_fill_desc(r->shaders, &d);
  d.sampler = CB200_SAMPLER; d.pointsampler = CB200_POINTS; d.colour_camera = CB200_COLOUR;
  d.max_path_len = 32; d.frame = rt.anim_frame; d.rank = 0; d.world = 1;
  r->r = cb200_render_create((cb200_accel_t *)accel_b200_handle(rt.accel), &d);
  if(!r->r) { fprintf(stderr, "[render b200] init failed: %s. there is no cpu fallback in this module.\n", cb200_last_error()); return 1; }
  r->width = d.width; r->height = d.height;
  return 0;
}

/* pointsampler_splat() of the cpu point samplers ends here (src/pointsampler.d/rand.c:59, halton.c:88); nothing on the gpu path
 * calls them -- the splat happens on the device (view_splat's filter and colour conversion, csrc/render.cu) */
void render_splat(const struct path_t *p, const mf_t value)
{
  (void)p; (void)value;
  static int said = 0;
  if(!said++) fprintf(stderr, "[render b200] render_splat: a cpu sampler produced a path; MOD_render=b200 ignores it (paths are traced and splatted on the gpu)\n");
}
#else
/* the reference's argument-less constructor (main.c: rt.render = render_init()) reads the camera, the shader list and the
 * film size from rt.*: that is the in-tree build above.  Standalone there is no rt: refuse instead of guessing. */
struct render_t *render_init()
{
  if(g_render) return g_render;
  fprintf(stderr, "[render b200] render_init: no scene description; call render_b200_init(accel, desc)\n");
  return 0;
}
#endif

void render_cleanup(struct render_t *r)
{
  if(!r) return;
  if(g_render == r) g_render = 0;
  if(r->r) cb200_render_destroy(r->r);
  if(r->shaders) scene_b200_free(r->shaders);
  free(r);
}

/* per-thread state of the cpu renderers (threads.h:33-55 calls these for every worker): nothing to keep */
struct render_tls_t *render_tls_init() { return 0; }
void render_tls_cleanup(struct render_tls_t *r) { (void)r; }

void render_clear()
{
  if(!g_render || !g_render->r) return;
  if(cb200_render_clear(g_render->r, 0)) fprintf(stderr, "[render b200] clear failed: %s\n", cb200_last_error());
  g_render->overlays = 0;
}

void render_sample_path(uint64_t index)
{
  (void)index;
  fprintf(stderr, "[render b200] render_sample_path: single paths are not served by the gpu module, "
                  "view_render must call render_b200_pass (no cpu fallback)\n");
  abort();
}

int render_b200_pass(struct render_t *r, uint64_t first_index, uint64_t count, float *fb)
{
  if(!r) { fprintf(stderr, "[render b200] pass: not initialised\n"); return 1; }
#ifdef CORONA_B200_IN_TREE
  if(!r->r && intree_create(r)) { fprintf(stderr, "[render b200] cannot render: giving up\n"); abort(); }
#endif
  /* streamed: paths still bouncing when every index has been started ride along with the next progression; the
   * framebuffer handed back is the progressive image as it stands (like the reference's display reading fb mid-flight) */
  int rc = cb200_render_pass_stream(r->r, first_index, count, 0);
  /* the copy to the host runs behind this progression on its own stream and overlaps the next one: what `fb` shows lags by at
   * most one progression, and it must stay valid until render_b200_finish / render_cleanup (the view's framebuffer does) */
  if(!rc && fb) rc = cb200_render_snapshot_async(r->r, fb, 0);
  if(rc) { fprintf(stderr, "[render b200] pass failed: %s\n", cb200_last_error()); return 1; }
  r->overlays += count/((uint64_t)r->width*r->height);
  return 0;
}

/* trace the stragglers to the end and hand back the finished image: before fb_export / a screenshot / the end of a batch */
int render_b200_finish(struct render_t *r, float *fb)
{
  if(!r || !r->r) { fprintf(stderr, "[render b200] finish: not initialised\n"); return 1; }
  int rc = cb200_render_flush(r->r, 0);
  if(!rc && fb) rc = cb200_render_download(r->r, fb, 0);
  if(rc) { fprintf(stderr, "[render b200] finish failed: %s\n", cb200_last_error()); return 1; }
  return 0;
}

int render_b200_set_dbor(struct render_t *r, int levels)
{
  if(!r) { fprintf(stderr, "[render b200] dbor: not initialised\n"); return 1; }
#ifdef CORONA_B200_IN_TREE
  if(!r->r && intree_create(r)) return 1;
#endif
  if(cb200_render_set_dbor(r->r, levels)) { fprintf(stderr, "[render b200] dbor failed: %s\n", cb200_last_error()); return 1; }
  return 0;
}

int render_b200_dbor(struct render_t *r, int level, float *fb)
{
  if(!r || !r->r || !fb) { fprintf(stderr, "[render b200] dbor: bad arguments\n"); return 1; }
  if(cb200_render_download_dbor(r->r, level, fb, 0)) { fprintf(stderr, "[render b200] dbor failed: %s\n", cb200_last_error()); return 1; }
  return 0;
}

uint64_t render_b200_overlays(const struct render_t *r) { return r ? r->overlays : 0; }
void *render_b200_handle(const struct render_t *r) { return r ? r->r : 0; }
int render_b200_dbor_is_set(struct render_t *r, int set) { const int was = r ? r->dbor_set : 0; if(r && set) r->dbor_set = 1; return was; }
