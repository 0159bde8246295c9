// traverse2.cu -- closest hit on the 4-wide tree with TWO rays per lane (sm_100a).
//
// Same per-ray semantics as k_intersect (traverse.cu; accel_intersect, src/accel.d/qbvhmp.c:1262-1490): every ray sees exactly the
// reference's sequence of node and primitive tests, so the answers are the same bits.  What changes is which lane runs them when.
//
// k_intersect votes per warp iteration between the NODE step and the PRIM step; lanes whose ray is in the other state sit the
// iteration out (ncu, profiles/r2m: slab test at 19 of 32 lanes, primitive loop at 7-10).  Here a lane owns two rays: one in
// registers, one parked in shared memory (21 words, 5 x STS.128 + 1).  After the vote a lane whose register ray is in the wrong
// state but whose parked ray is in the voted one exchanges them (12 shared-memory instructions against the ~140 of a node step),
// so nearly every lane that holds a ray at all takes part in every iteration.  The two rays have their own halves of the lane's
// local-memory stack.  Refill: a lane with a free slot parks its register ray and takes the new one into registers.
#include <cstdlib>
#include <type_traits>

#include "traverse_common.cuh"

#ifndef DUAL_MIN_BLOCKS
#define DUAL_MIN_BLOCKS 6   // <= 80 registers: 24 warps per SM, 48 rays per warp-slot pair
#endif

#define CSWAP2(cond, ka, ca, kb, cb) do { const float tk__ = ka; const uint32_t tc__ = ca; \
  ka = (cond) ? kb : ka; ca = (cond) ? cb : ca; kb = (cond) ? tk__ : kb; cb = (cond) ? tc__ : cb; } while(0)

// sign-selected slab test of a static node, identical to node_slabs_fast in traverse.cu
__device__ __forceinline__ void node_slabs_fast2(const Node128 *__restrict__ n, const uint32_t near_off[3], float px, float py, float pz,
                                                 float ix, float iy, float iz, float tmax_init, float key[4])
{
  const float4 *a0 = reinterpret_cast<const float4 *>(n->aabb0);
  const float4 nx = __ldg(a0 + near_off[0]),     ny = __ldg(a0 + 1 + near_off[1]),     nz = __ldg(a0 + 2 + near_off[2]);
  const float4 fx = __ldg(a0 + 3 - near_off[0]), fy = __ldg(a0 + 4 - near_off[1]),     fz = __ldg(a0 + 5 - near_off[2]);
  float a[4], b[4], c[4], d[4], e[4], f[4];
  slab2(nx.x, nx.y, px, ix, a[0], a[1]); slab2(nx.z, nx.w, px, ix, a[2], a[3]);
  slab2(ny.x, ny.y, py, iy, b[0], b[1]); slab2(ny.z, ny.w, py, iy, b[2], b[3]);
  slab2(nz.x, nz.y, pz, iz, c[0], c[1]); slab2(nz.z, nz.w, pz, iz, c[2], c[3]);
  slab2(fx.x, fx.y, px, ix, d[0], d[1]); slab2(fx.z, fx.w, px, ix, d[2], d[3]);
  slab2(fy.x, fy.y, py, iy, e[0], e[1]); slab2(fy.z, fy.w, py, iy, e[2], e[3]);
  slab2(fz.x, fz.y, pz, iz, f[0], f[1]); slab2(fz.z, fz.w, pz, iz, f[2], f[3]);
#pragma unroll
  for(int k=0;k<4;k++)
  {
    const float tmin = fmaxf(max3f(a[k], b[k], c[k]), 0.0f);
    const float tmax = fminf(min3f(d[k], e[k], f[k]), tmax_init);
    key[k] = tmin <= tmax ? tmin : KEY_MISS;
  }
}

// static scenes (Node128), < 2^26 primitives (32-bit child references)
template<int STACK, bool ANALYTIC>
__global__ void __launch_bounds__(TRACE_BLOCK, DUAL_MIN_BLOCKS)
k_intersect_dual(DevAccel A, const cb_ray_t *__restrict__ rays, const float *__restrict__ max_dist,
                 cb_hitrec_t *__restrict__ out, uint32_t n, unsigned int *ticket, int prim_threshold, int refill_threshold)
{
  __shared__ float4 park[5][TRACE_BLOCK];     // the parked ray of every lane
  __shared__ uint32_t park_w[TRACE_BLOCK];    // nearbits | exact << 3 | sp << 8
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint64_t stack[2*STACK];                    // (child << 32) | entry distance bits; the register ray's half starts at sbase
  uint32_t sbase = 0;
  int sp = 0;
  int state = ST_IDLE, state2 = ST_IDLE;      // register ray, parked ray
  bool exhausted = false;
  RayD r;
  HitD h;
  uint32_t ray_i = 0;
  uint32_t cur = 0;
  float ix = 0.0f, iy = 0.0f, iz = 0.0f;
  uint32_t nearbits = 0;
  uint32_t near_off[3] = {0, 0, 0};
  bool exact = false;
  const uint32_t rec_stride = A.rec_units*4;
  r.px = r.py = r.pz = r.dx = r.dy = r.dz = r.time = r.min_dist = 0.0f; r.ign_lo = r.ign_hi = 0;
  h.dist = 0.0f; h.u = h.v = 0.0f; h.prim_lo = h.prim_hi = 0;

#define PARK_STORE() do { \
    park[0][tid] = make_float4(r.px, r.py, r.pz, r.dx); \
    park[1][tid] = make_float4(r.dy, r.dz, ix, iy); \
    park[2][tid] = make_float4(iz, r.min_dist, __uint_as_float(r.ign_lo), __uint_as_float(r.ign_hi)); \
    park[3][tid] = make_float4(h.dist, h.u, h.v, r.time); \
    park[4][tid] = make_float4(__uint_as_float(h.prim_lo), __uint_as_float(h.prim_hi), __uint_as_float(ray_i), __uint_as_float(cur)); \
    park_w[tid] = nearbits | (exact ? 8u : 0u) | ((uint32_t)sp << 8); } while(0)
#define PARK_ASSIGN(q0, q1, q2, q3, q4, w) do { \
    r.px = q0.x; r.py = q0.y; r.pz = q0.z; r.dx = q0.w; r.dy = q1.x; r.dz = q1.y; ix = q1.z; iy = q1.w; \
    iz = q2.x; r.min_dist = q2.y; r.ign_lo = __float_as_uint(q2.z); r.ign_hi = __float_as_uint(q2.w); \
    h.dist = q3.x; h.u = q3.y; h.v = q3.z; r.time = q3.w; \
    h.prim_lo = __float_as_uint(q4.x); h.prim_hi = __float_as_uint(q4.y); ray_i = __float_as_uint(q4.z); cur = __float_as_uint(q4.w); \
    nearbits = w & 7u; exact = (w & 8u) != 0u; sp = (int)(w >> 8); \
    near_off[0] = 3u*(nearbits & 1u); near_off[1] = 3u*((nearbits >> 1) & 1u); near_off[2] = 3u*(nearbits >> 2); } while(0)

  while(true)
  {
    // ---- a lane whose register ray has finished goes on with its parked one
    if(state == ST_IDLE && state2 != ST_IDLE)
    {
      const float4 q0 = park[0][tid], q1 = park[1][tid], q2 = park[2][tid], q3 = park[3][tid], q4 = park[4][tid];
      const uint32_t w = park_w[tid];
      PARK_ASSIGN(q0, q1, q2, q3, q4, w);
      state = state2; state2 = ST_IDLE; sbase ^= (uint32_t)STACK;
    }
    // ---- refill: every lane with a free slot takes one ray (a lane with both slots free gets its second one at the next refill)
    const uint32_t emptyA = __ballot_sync(FULL, state == ST_IDLE);
    const uint32_t emptyB = __ballot_sync(FULL, state2 == ST_IDLE);
    if(emptyB && !exhausted && (__popc(emptyB) >= refill_threshold || emptyA == FULL))
    {
      const uint32_t want = __popc(emptyB);
      unsigned int base = 0;
      if(lane == 0) base = atomicAdd(ticket, want);
      base = __shfl_sync(FULL, base, 0);
      if(base + want >= n) exhausted = true;
      if(state2 == ST_IDLE)
      {
        const uint32_t i = base + __popc(emptyB & lt_mask);
        if(i < n)
        {
          if(state != ST_IDLE) { PARK_STORE(); state2 = state; sbase ^= (uint32_t)STACK; }
          load_ray(rays, i, r);
          ray_i = i;
          h.dist = max_dist ? __ldg(max_dist + i) : FLT_MAX;
          h.u = 0.0f; h.v = 0.0f; h.prim_lo = 0xffffffffu; h.prim_hi = 0xffffffffu;
          nearbits = (__float_as_uint(r.dx) >> 31) | ((__float_as_uint(r.dy) >> 31) << 1) | ((__float_as_uint(r.dz) >> 31) << 2);
          near_off[0] = 3u*(nearbits & 1u); near_off[1] = 3u*((nearbits >> 1) & 1u); near_off[2] = 3u*(nearbits >> 2);
          ix = 1.0f/r.dx; iy = 1.0f/r.dy; iz = 1.0f/r.dz;
          exact = !(finite_nonzero(ix) && finite_nonzero(iy) && finite_nonzero(iz) &&
                    finite(r.px) && finite(r.py) && finite(r.pz) && finite(r.time) && !(h.dist != h.dist));
          sp = 0; cur = 0; state = ST_NODE;
        }
      }
    }
    // ---- vote: which step runs; a lane that can only take part with its parked ray exchanges the two
    const uint32_t canN = __ballot_sync(FULL, state == ST_NODE || state2 == ST_NODE);
    const uint32_t canP = __ballot_sync(FULL, state == ST_PRIM || state2 == ST_PRIM);
    if(!(canN | canP)) break;   // nothing in flight: the refill above ran (all lanes idle) and the queue is empty
    const int live = __popc(__ballot_sync(FULL, state != ST_IDLE));
    const int thr = prim_threshold >= 0 ? prim_threshold : max(2, (live*(-prim_threshold) + 31) >> 5);
    const bool do_prims = (canN == 0u) || (__popc(canP) >= thr);
    const int want_state = do_prims ? ST_PRIM : ST_NODE;
    if(state != want_state && state2 == want_state)
    {
      const float4 q0 = park[0][tid], q1 = park[1][tid], q2 = park[2][tid], q3 = park[3][tid], q4 = park[4][tid];
      const uint32_t w = park_w[tid];
      PARK_STORE();
      PARK_ASSIGN(q0, q1, q2, q3, q4, w);
      const int t = state; state = state2; state2 = t;
      sbase ^= (uint32_t)STACK;
    }

    bool need_pop = false, new_cur = false;
    if(do_prims)
    {
      if(state == ST_PRIM)
      { // the whole leaf in primid[] order (qbvhmp.c:1371-1379)
        const float4 *rec = A.recs + (uint64_t)((cur ^ 0x80000000u) >> 5)*(uint64_t)rec_stride;
        uint32_t prims_left = cur & 31u;   // empty leaves are never pushed: >= 1
        do
        {
          prim_intersect<ANALYTIC>(rec, A.rec_units, r, h);
          rec += rec_stride;
        }
        while(--prims_left);
        need_pop = true;
      }
    }
    else if(state == ST_NODE)
    {
      float key[4];
      uint32_t child[4];
      int axis0, axis00, axis01;
      if(exact)
      { // rays with zero / infinite / NaN components: the reference's select semantics
        NodeOut o;
        node_slabs<false, true>(A.nodes, cur, r.px, r.py, r.pz, ix, iy, iz, 1.0f - r.time, r.time, h.dist, o);
#pragma unroll
        for(int c=0;c<4;c++)
        {
          key[c] = (o.hit[c] && !is_empty_leaf(o.child[c])) ? o.tmin[c] : KEY_MISS;
          child[c] = (uint32_t)o.child[c] | (uint32_t)(o.child[c] >> 32);
        }
        axis0 = o.axis0; axis00 = o.axis00; axis01 = o.axis01;
      }
      else
      {
        const Node128 *nd = reinterpret_cast<const Node128 *>(A.nodes) + cur;
        node_slabs_fast2(nd, near_off, r.px, r.py, r.pz, ix, iy, iz, h.dist, key);
        const uint4 *ch = reinterpret_cast<const uint4 *>(nd->child);
        const uint4 c01 = __ldg(ch), c23 = __ldg(ch + 1);   // {lo0, hi0, lo1, hi1}, {lo2, hi2, lo3, hi3}
        const uint32_t ax = (c01.y >> (CB_AXIS_SHIFT - 32)) & 63u;
        const uint32_t hi0 = c01.y & (uint32_t)(CB_CHILD_MASK >> 32);
        child[0] = c01.x | hi0; child[1] = c01.z | c01.w; child[2] = c23.x | c23.y; child[3] = c23.z | c23.w;
        axis0 = ax & 3; axis00 = (ax >> 2) & 3; axis01 = (ax >> 4) & 3;
      }
      // the reference's topological order (qbvhmp.c:1313-1320) as conditional swaps
      const bool s00 = (nearbits >> axis00) & 1u, s01 = (nearbits >> axis01) & 1u, s0 = (nearbits >> axis0) & 1u;
      CSWAP2(s00, key[0], child[0], key[1], child[1]);
      CSWAP2(s01, key[2], child[2], key[3], child[3]);
      CSWAP2(s0,  key[0], child[0], key[2], child[2]);
      CSWAP2(s0,  key[1], child[1], key[3], child[3]);
      need_pop = true;
      const int first = KEY_HIT(key[0]) ? 0 : KEY_HIT(key[1]) ? 1 : KEY_HIT(key[2]) ? 2 : KEY_HIT(key[3]) ? 3 : 4;
      if(first < 4)
      {
#define PUSH(k) do { stack[sbase + sp] = ((uint64_t)child[k] << 32) | __float_as_uint(key[k]); sp++; } while(0)
        if(KEY_HIT(key[3]) && first < 3) PUSH(3);
        if(KEY_HIT(key[2]) && first < 2) PUSH(2);
        if(KEY_HIT(key[1]) && first < 1) PUSH(1);
#undef PUSH
        cur = first == 0 ? child[0] : first == 1 ? child[1] : first == 2 ? child[2] : child[3];
        need_pop = false;
        new_cur = true;
      }
    }
    if(need_pop)
    {
      while(sp > 0)
      {
        --sp;
        const uint64_t e = stack[sbase + sp];
        if(__uint_as_float((uint32_t)e) > h.dist) continue;
        cur = (uint32_t)(e >> 32);
        new_cur = true;
        break;
      }
      if(!new_cur)
      { // ray finished: 24-byte record as three 8-byte stores
        uint2 *o2 = reinterpret_cast<uint2 *>(out + ray_i);
        o2[0] = make_uint2(h.prim_lo, h.prim_hi);
        o2[1] = make_uint2(__float_as_uint(h.u), __float_as_uint(h.v));
        o2[2] = make_uint2(__float_as_uint(h.dist), 0u);
        state = ST_IDLE;
      }
    }
    if(new_cur) state = (cur & 0x80000000u) ? ST_PRIM : ST_NODE;
  }
#undef PARK_STORE
#undef PARK_ASSIGN
}

// CB200_DUAL=0/1 in the environment overrides the default (A/B measurements)
static int g_dual = -1;
bool cb200_dual_enabled()
{
  if(g_dual < 0) { const char *e = getenv("CB200_DUAL"); g_dual = e ? (atoi(e) ? 1 : 0) : CB200_DUAL_DEFAULT; }
  return g_dual == 1;
}
static int env_int(const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; }

template<bool ANALYTIC>
static int launch_dual(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out, uint64_t n, cudaStream_t stream)
{
  static int prim_thr = 1000, refill_thr = -1;
  if(prim_thr == 1000) { prim_thr = env_int("CB200_DUAL_PRIM_THRESHOLD", -20); if(prim_thr == 0) prim_thr = 1; }
  if(refill_thr < 0) { refill_thr = env_int("CB200_DUAL_REFILL_THRESHOLD", 16); if(refill_thr < 1) refill_thr = 1; if(refill_thr > 32) refill_thr = 32; }
  auto k = k_intersect_dual<STACK_SMALL, ANALYTIC>;
  const uint64_t LAUNCH_MAX = 1ull << 30;
  for(uint64_t first=0; first<n; first+=LAUNCH_MAX)
  {
    const uint64_t m = n - first < LAUNCH_MAX ? n - first : LAUNCH_MAX;
    unsigned int *ticket;
    if(int rc = cb200_get_ticket(stream, &ticket)) return rc;
    // two rays per thread: half the blocks cover a small launch
    k<<<cb200_trace_grid((m + 1)/2, (const void *)k), TRACE_BLOCK, 0, stream>>>(a->dev, d_rays + first, d_max_dist ? d_max_dist + first : nullptr,
                                                                                 d_out + first, (uint32_t)m, ticket, prim_thr, refill_thr);
    cb200_count_launch();
    CB_CUDA(cudaGetLastError());
  }
  return 0;
}

// static scene, 32-bit child references, tree within the small stack: the caller (traverse.cu) checks
int cb200_launch_intersect_dual(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                                uint64_t n, cudaStream_t stream)
{
  return a->scene->any_analytic ? launch_dual<true>(a, d_rays, d_max_dist, d_out, n, stream)
                                : launch_dual<false>(a, d_rays, d_max_dist, d_out, n, stream);
}
