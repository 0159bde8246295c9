/* ref_halton.c -- TEST INFRASTRUCTURE.  The reference's own Halton point set (ext/halton/halton.h, used by
 * src/pointsampler.d/halton.c:69-84) behind a batch entry point, compiled in place by oracle/Makefile into
 * oracle/_ref/libref_halton.so.  Used by tests/golden/make_golden.py to record known-answer vectors for
 * pointsampler() and by tests/test_oracle_vs_ref.py to pin oracle/points.py.  Never loaded by the product. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "ext/halton/halton.h"   /* resolved through -I$(REF) */

static halton_t g_h;
static uint64_t g_frame = ~0ull;

void ref_halton_sample(uint64_t frame, const uint64_t *index, const int32_t *dim, float *out, uint64_t n)
{
  if(frame != g_frame) { halton_init_random(&g_h, frame); g_frame = frame; }
  /* pointsampler(): "note that this clips the bits in p->index to 32" (halton.c:82) */
  for(uint64_t i=0;i<n;i++) out[i] = halton_sample(&g_h, (unsigned)dim[i], (unsigned)index[i]);
}
int ref_halton_dims(void) { return halton_get_num_dimensions(); }
