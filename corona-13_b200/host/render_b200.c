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

struct render_t
{
  cb200_render_t *r;
  uint32_t width, height;
  uint64_t overlays;     /* progressions accumulated so far (view->overlays) */
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

/* the reference's argument-less constructor (main.c: rt.render = render_init()) reads the camera, the shader list and the
 * film size from rt.*; the in-tree build fills a cb_render_desc_t from those and calls render_b200_init.
 * Standalone there is no rt: refuse instead of guessing. */
struct render_t *render_init()
{
  if(g_render) return g_render;
  fprintf(stderr, "[render b200] render_init: no scene description; call render_b200_init(accel, desc)\n");
  return 0;
}

void render_cleanup(struct render_t *r)
{
  if(!r) return;
  if(g_render == r) g_render = 0;
  cb200_render_destroy(r->r);
  free(r);
}

/* per-thread state of the cpu renderers (threads.h:33-55 calls these for every worker): nothing to keep */
struct render_tls_t *render_tls_init() { return 0; }
void render_tls_cleanup(struct render_tls_t *r) { (void)r; }

void render_clear()
{
  if(!g_render) return;
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
  if(!r) { fprintf(stderr, "[render b200] finish: not initialised\n"); return 1; }
  int rc = cb200_render_flush(r->r, 0);
  if(!rc && fb) rc = cb200_render_download(r->r, fb, 0);
  if(rc) { fprintf(stderr, "[render b200] finish failed: %s\n", cb200_last_error()); return 1; }
  return 0;
}

int render_b200_set_dbor(struct render_t *r, int levels)
{
  if(!r) { fprintf(stderr, "[render b200] dbor: not initialised\n"); return 1; }
  if(cb200_render_set_dbor(r->r, levels)) { fprintf(stderr, "[render b200] dbor failed: %s\n", cb200_last_error()); return 1; }
  return 0;
}

int render_b200_dbor(struct render_t *r, int level, float *fb)
{
  if(!r || !fb) { fprintf(stderr, "[render b200] dbor: bad arguments\n"); return 1; }
  if(cb200_render_download_dbor(r->r, level, fb, 0)) { fprintf(stderr, "[render b200] dbor failed: %s\n", cb200_last_error()); return 1; }
  return 0;
}

uint64_t render_b200_overlays(const struct render_t *r) { return r ? r->overlays : 0; }
void *render_b200_handle(const struct render_t *r) { return r ? r->r : 0; }
