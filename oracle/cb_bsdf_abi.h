/* the two records of cb200_render_bsdf (include/corona_b200_render.h), restated for oracle/ref_bsdf.c which cannot include the
 * product header next to the reference's own (both define MAX/MIN style macros and a `hit_t`) */
#pragma once
#include <stdint.h>
typedef struct cb_bsdf_query_t { float wi[3], wo[3], lambda, rand[3], rd, rs, rg, roughness; int32_t flip; } cb_bsdf_query_t;
typedef struct cb_bsdf_result_t { float s_wo[3], s_weight, s_pdf; uint32_t s_mode; float f; uint32_t f_mode; float pdf; } cb_bsdf_result_t;
/* the two records of cb200_render_medium */
typedef struct cb_medium_query_t { float wi[3], wo[3], lambda, rand[3], dist; } cb_medium_query_t;
typedef struct cb_medium_result_t { float mu_t, mu_s, free_dist, free_pdf, transmittance, vol_pdf; float s_wo[3], s_weight, s_pdf; uint32_t s_mode; float f; uint32_t f_mode; float pdf; } cb_medium_result_t;
