// traverse8.cu -- closest-hit and any-hit traversal of the 8-wide compressed BVH (Node8, internal.h): the throughput path of
// static scenes (sm_100a).
//
// What the reference computes stays what it is: every primitive test is prim_intersect / prim_visible of prims.cuh, the
// reference's own arithmetic in its own operation order (src/prims.c:638-701, include/geo/triangle.h:263-343), with the
// acceptance rules of accel_intersect / accel_visible (src/accel.d/qbvhmp.c:1262-1490).  What changes is the culling structure
// in front of it.  ncu on the 4-wide kernel (profiles/r1h_k_intersect_ncu.json) shows it co-limited by issue slots and by L1
// wavefronts: 3/4 of those are node and record gathers (one 128-byte node per lane and visit, 20.6 visits per ray), 1/5 the
// per-lane stack in local memory.  Node8 answers that:
//   * 8 children in 96 bytes (three sectors) instead of 4 in 128: quantised boxes, 8 bits per plane on a per-node grid, rounded
//     outwards -- a stored box contains the 4-wide tree's box, so no primitive the reference tests is lost; fewer, smaller gathers
//     per ray, and the whole node array of the 10 M-triangle bench (85 MB) stays L2-resident;
//   * the slab test works on the bytes directly: PRMT builds the float 256 + q/128 from a plane byte, one packed FADD2 removes
//     the 256 and one packed FFMA2 (f32x2, new with sm_100) evaluates q * (2^e / d) + (origin - p) / d for two children at
//     once, rounded down for entry planes and up for exit planes (FFMA2.RM / .RP); the three entry distances and the three exit
//     distances are combined with the three-input FMNMX3;
//   * the children of a node are visited in octant order (slot ^ ray octant, slots assigned by the builder from the child
//     centres; Ylitie, Karras, Laine: "Efficient incoherent ray traversal on GPUs through compressed wide BVHs", HPG 2017), so
//     no sorting network; the children still to visit are ONE stack word (node index << 8 | hit mask) instead of up to three
//     (child, distance) pairs per node: a ray pushes ~10 words instead of ~60;
//   * rays the byte arithmetic cannot represent conservatively (zero / non-finite direction components, |1/d| or |p| beyond
//     2^40, NaN limits) take the reference's exact select semantics on the 4-wide tree, one ray per lane, out of line.
// Lane scheduling is the 4-wide kernel's: persistent warps, batched refill from a ticket counter, NODE / PRIM phases by vote.
//
// Results: prim / u / v / dist are bit-identical to the 4-wide kernels' except where the visiting order decides (two primitives
// at exactly the same distance: the last one tested wins, triangle.h:296) or a box test is decided by its last ulp; both are
// tree-dependent in the reference as well and are the cases tests/helpers.py:classify_mismatches proves one by one.
#include "traverse_common.cuh"
#include <cstdlib>

#ifndef TRACE8_MIN_BLOCKS
#define TRACE8_MIN_BLOCKS 6   // 80 registers -> 24 resident warps per SM
#endif

// ---------------------------------------------------------------------------------------------
// packed helpers (inline PTX: f32x2 arithmetic and three-input min / max exist from sm_100 on)
// ---------------------------------------------------------------------------------------------
// bytes j and j+1 of w -> q/128 each -> q/128 * S + B, rounded down (entry planes) or up (exit planes)
template<int J, bool UP>
__device__ __forceinline__ void planes2(uint32_t w, float S, float B, float &t0, float &t1)
{
  const uint32_t a = __byte_perm(w, 0x43800000u, 0x7604 | (J << 4));         // 256 + q/128: q in mantissa bits 8..15
  const uint32_t b = __byte_perm(w, 0x43800000u, 0x7604 | ((J + 1) << 4));
  if(UP)
    asm("{ .reg .b64 v, m, s, c;\n"
        "mov.b64 v, {%2, %3};\n"
        "mov.b64 m, {%4, %4};\n"
        "add.rn.f32x2 v, v, m;\n"
        "mov.b64 s, {%5, %5};\n"
        "mov.b64 c, {%6, %6};\n"
        "fma.rp.f32x2 v, v, s, c;\n"
        "mov.b64 {%0, %1}, v; }\n" : "=f"(t0), "=f"(t1) : "r"(a), "r"(b), "f"(-256.0f), "f"(S), "f"(B));
  else
    asm("{ .reg .b64 v, m, s, c;\n"
        "mov.b64 v, {%2, %3};\n"
        "mov.b64 m, {%4, %4};\n"
        "add.rn.f32x2 v, v, m;\n"
        "mov.b64 s, {%5, %5};\n"
        "mov.b64 c, {%6, %6};\n"
        "fma.rm.f32x2 v, v, s, c;\n"
        "mov.b64 {%0, %1}, v; }\n" : "=f"(t0), "=f"(t1) : "r"(a), "r"(b), "f"(-256.0f), "f"(S), "f"(B));
}

struct Ray8   // per-ray constants of the byte slab test
{
  float ix, iy, iz;
  uint32_t octinv;   // 7 - octant: bit k set = the ray travels towards +k
};

// can the byte arithmetic serve this ray?  (finite non-zero direction, magnitudes that keep 2^e/d and (o-p)/d finite)
__device__ __forceinline__ bool ray8_ok(const RayD &r, float ix, float iy, float iz, float limit)
{
  const float big = 1099511627776.0f;   // 2^40
  const float ax = fabsf(ix), ay = fabsf(iy), az = fabsf(iz);
  return ax > 0.0f && ax < big && ay > 0.0f && ay < big && az > 0.0f && az < big &&
         fabsf(r.px) < big && fabsf(r.py) < big && fabsf(r.pz) < big && !(limit != limit);
}

// 8 slab tests of one node against [0, tmax]: bit c of the result = child slot c is hit
template<bool CNT>
__device__ __forceinline__ uint32_t node8_test(const Node8 *__restrict__ nd, const RayD &r, const Ray8 &q, float tmax)
{
  const uint4 *n4 = reinterpret_cast<const uint4 *>(nd);
  const uint4 hd = __ldg(n4), X = __ldg(n4 + 1), Y = __ldg(n4 + 2), Z = __ldg(n4 + 3);
  const float Sx = __uint_as_float((hd.w & 0x000000ffu) << 23)*q.ix;
  const float Sy = __uint_as_float((hd.w & 0x0000ff00u) << 15)*q.iy;
  const float Sz = __uint_as_float((hd.w & 0x00ff0000u) << 7)*q.iz;
  const float Bx = (__uint_as_float(hd.x) - r.px)*q.ix;
  const float By = (__uint_as_float(hd.y) - r.py)*q.iy;
  const float Bz = (__uint_as_float(hd.z) - r.pz)*q.iz;
  // entry / exit plane words by the ray's direction: (x,y) = lower planes of children 0..3 / 4..7, (z,w) = upper planes
  const bool px = q.octinv & 1u, py = q.octinv & 2u, pz = q.octinv & 4u;
  const uint32_t nx[2] = {px ? X.x : X.z, px ? X.y : X.w}, fx[2] = {px ? X.z : X.x, px ? X.w : X.y};
  const uint32_t ny[2] = {py ? Y.x : Y.z, py ? Y.y : Y.w}, fy[2] = {py ? Y.z : Y.x, py ? Y.w : Y.y};
  const uint32_t nz[2] = {pz ? Z.x : Z.z, pz ? Z.y : Z.w}, fz[2] = {pz ? Z.z : Z.x, pz ? Z.w : Z.y};
  uint32_t mask = 0;
#define PAIR(H, J, C) do { \
    float a0, a1, b0, b1, c0, c1, d0, d1, e0, e1, f0, f1; \
    planes2<J, false>(nx[H], Sx, Bx, a0, a1); planes2<J, false>(ny[H], Sy, By, b0, b1); planes2<J, false>(nz[H], Sz, Bz, c0, c1); \
    planes2<J, true >(fx[H], Sx, Bx, d0, d1); planes2<J, true >(fy[H], Sy, By, e0, e1); planes2<J, true >(fz[H], Sz, Bz, f0, f1); \
    const float lo0 = fmaxf(max3f(a0, b0, c0), 0.0f), hi0 = fminf(min3f(d0, e0, f0), tmax); \
    const float lo1 = fmaxf(max3f(a1, b1, c1), 0.0f), hi1 = fminf(min3f(d1, e1, f1), tmax); \
    if(lo0 <= hi0) mask |= 1u << (C); \
    if(lo1 <= hi1) mask |= 2u << (C); } while(0)
  PAIR(0, 0, 0); PAIR(0, 2, 2); PAIR(1, 0, 4); PAIR(1, 2, 6);
#undef PAIR
  return mask;
}

// slot-order hit mask -> visiting-priority order: slot c moves to bit c ^ octinv (the highest bit is visited first)
__device__ __forceinline__ uint32_t to_visit_order(uint32_t m, uint32_t octinv)
{
  if(octinv & 1u) m = ((m & 0x55u) << 1) | ((m >> 1) & 0x55u);
  if(octinv & 2u) m = ((m & 0x33u) << 2) | ((m >> 2) & 0x33u);
  if(octinv & 4u) m = ((m & 0x0fu) << 4) | (m >> 4);
  return m;
}

// ---------------------------------------------------------------------------------------------
// exact path for the rays ray8_ok() turns away: accel_intersect / the any-hit sweep on the 4-wide tree with the reference's
// select semantics, one ray per lane (the code of traverse.cu's `exact` branch without the lane scheduling)
// ---------------------------------------------------------------------------------------------
#define CSWAP64(cond, ka, ca, kb, cb) do { const float tk__ = ka; const uint64_t tc__ = ca; \
  ka = (cond) ? kb : ka; ca = (cond) ? cb : ca; kb = (cond) ? tk__ : kb; cb = (cond) ? tc__ : cb; } while(0)

template<bool ANALYTIC>
__device__ __noinline__ void trace_exact_closest(const DevAccel &A, const RayD &r, HitD &h)
{
  const uint32_t nearbits = (__float_as_uint(r.dx) >> 31) | ((__float_as_uint(r.dy) >> 31) << 1) | ((__float_as_uint(r.dz) >> 31) << 2);
  const float ix = 1.0f/r.dx, iy = 1.0f/r.dy, iz = 1.0f/r.dz;
  uint64_t stack[STACK_EXACT];
  float stack_dist[STACK_EXACT];
  int sp = 0;
  uint64_t cur = 0;
  const uint32_t rec_stride = A.rec_units*4;
  while(true)
  {
    bool have = false;
    if(cur & CB_LEAF_BIT)
    {
      const float4 *rec = A.recs + ((cur ^ CB_LEAF_BIT) >> 5)*(uint64_t)rec_stride;
      for(uint32_t k=(uint32_t)cur & 31u; k; k--, rec += rec_stride) prim_intersect<ANALYTIC>(rec, A.rec_units, r, h);
    }
    else
    {
      NodeOut o;
      node_slabs<false, true>(A.nodes, cur, r.px, r.py, r.pz, ix, iy, iz, 1.0f, 0.0f, h.dist, o);
      float key[4];
      uint64_t child[4];
#pragma unroll
      for(int c=0;c<4;c++) { key[c] = (o.hit[c] && !is_empty_leaf(o.child[c])) ? o.tmin[c] : KEY_MISS; child[c] = o.child[c]; }
      const bool s00 = (nearbits >> o.axis00) & 1u, s01 = (nearbits >> o.axis01) & 1u, s0 = (nearbits >> o.axis0) & 1u;
      CSWAP64(s00, key[0], child[0], key[1], child[1]);
      CSWAP64(s01, key[2], child[2], key[3], child[3]);
      CSWAP64(s0,  key[0], child[0], key[2], child[2]);
      CSWAP64(s0,  key[1], child[1], key[3], child[3]);
      const int first = KEY_HIT(key[0]) ? 0 : KEY_HIT(key[1]) ? 1 : KEY_HIT(key[2]) ? 2 : KEY_HIT(key[3]) ? 3 : 4;
      if(first < 4)
      {
        if(KEY_HIT(key[3]) && first < 3) { stack_dist[sp] = key[3]; stack[sp++] = child[3]; }
        if(KEY_HIT(key[2]) && first < 2) { stack_dist[sp] = key[2]; stack[sp++] = child[2]; }
        if(KEY_HIT(key[1]) && first < 1) { stack_dist[sp] = key[1]; stack[sp++] = child[1]; }
        cur = first == 0 ? child[0] : first == 1 ? child[1] : first == 2 ? child[2] : child[3];
        have = true;
      }
    }
    if(have) continue;
    while(sp > 0)
    {
      --sp;
      if(stack_dist[sp] > h.dist) continue;
      cur = stack[sp];
      have = true;
      break;
    }
    if(!have) return;
  }
}

// returns 1 = nothing blocks the ray (k_visible's semantics, traverse.cu)
template<bool ANALYTIC, bool SHADOW>
__device__ __noinline__ int trace_exact_any(const DevAccel &A, const RayD &r, float md, uint2 skip_id)
{
  const float ix = 1.0f/r.dx, iy = 1.0f/r.dy, iz = 1.0f/r.dz;
  uint64_t stack[STACK_EXACT + 4];
  int sp = 0;
  uint64_t cur = 0;
  const uint32_t rec_stride = A.rec_units*4;
  while(true)
  {
    if(cur & CB_LEAF_BIT)
    {
      const float4 *rec = A.recs + ((cur ^ CB_LEAF_BIT) >> 5)*(uint64_t)rec_stride;
      for(uint32_t k=(uint32_t)cur & 31u; k; k--, rec += rec_stride)
      {
        if(SHADOW)
        {
          HitD ht;
          ht.dist = md; ht.u = ht.v = 0.0f; ht.prim_lo = ht.prim_hi = 0xffffffffu;
          prim_intersect<ANALYTIC>(rec, A.rec_units, r, ht);
          if((ht.prim_lo & ht.prim_hi) != 0xffffffffu && !(ht.prim_lo == skip_id.x && ht.prim_hi == skip_id.y)) return 0;
        }
        else if(prim_visible<ANALYTIC>(rec, A.rec_units, r, md)) return 0;
      }
    }
    else
    {
      NodeOut o;
      node_slabs<false, true>(A.nodes, cur, r.px, r.py, r.pz, ix, iy, iz, 1.0f, 0.0f, md, o);
#pragma unroll
      for(int c=0;c<4;c++) if(o.hit[c] && !is_empty_leaf(o.child[c])) stack[sp++] = o.child[c];
    }
    if(sp == 0) return 1;
    cur = stack[--sp];
  }
}

// ---------------------------------------------------------------------------------------------
// closest hit
// ---------------------------------------------------------------------------------------------
template<bool CNT, bool ANALYTIC>
__global__ void __launch_bounds__(TRACE_BLOCK, (CNT || ANALYTIC) ? 1 : TRACE8_MIN_BLOCKS)
k_intersect8(DevAccel A, const cb_ray_t *__restrict__ rays, const float *__restrict__ max_dist,
             cb_hitrec_t *__restrict__ out, uint32_t n, unsigned int *ticket, unsigned long long *counters,
             int prim_threshold, int refill_threshold)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  unsigned long long cnt[4] = {0, 0, 0, 0};
  uint32_t stack[CB8_STACK];   // groups: node index << 8 | children still to visit (visiting-priority positions)
  int sp = 0;
  int state = ST_IDLE;
  bool exhausted = false;
  RayD r;
  HitD h;
  Ray8 q;
  uint32_t ray_i = 0;
  uint32_t cur = 0;            // ST_NODE: the node to test next
  uint32_t grp = 0;            // the group being worked on
  const float4 *rec = nullptr;
  uint32_t prims_left = 0;
  const uint32_t rec_stride = A.rec_units*4;
  r.px = r.py = r.pz = r.dx = r.dy = r.dz = r.time = r.min_dist = 0.0f; r.ign_lo = r.ign_hi = 0;
  h.dist = 0.0f; h.u = h.v = 0.0f; h.prim_lo = h.prim_hi = 0;
  q.ix = q.iy = q.iz = 0.0f; q.octinv = 0;

  while(true)
  {
    // ---- refill idle lanes from the global ray queue in batches (traverse.cu)
    const uint32_t idle = __ballot_sync(FULL, state == ST_IDLE);
    if(idle && !exhausted && (__popc(idle) >= refill_threshold || idle == FULL))
    {
      const uint32_t want = __popc(idle);
      unsigned int base = 0;
      if(lane == 0) base = atomicAdd(ticket, want);
      base = __shfl_sync(FULL, base, 0);
      if(base + want >= n) exhausted = true;
      if(state == ST_IDLE)
      {
        const uint32_t i = base + __popc(idle & lt_mask);
        if(i < n)
        {
          load_ray(rays, i, r);
          ray_i = i;
          h.dist = max_dist ? __ldg(max_dist + i) : FLT_MAX;
          h.u = 0.0f; h.v = 0.0f; h.prim_lo = 0xffffffffu; h.prim_hi = 0xffffffffu;
          q.ix = 1.0f/r.dx; q.iy = 1.0f/r.dy; q.iz = 1.0f/r.dz;
          q.octinv = 7u ^ ((__float_as_uint(r.dx) >> 31) | ((__float_as_uint(r.dy) >> 31) << 1) | ((__float_as_uint(r.dz) >> 31) << 2));
          if(CNT) cnt[0]++;
          if(ray8_ok(r, q.ix, q.iy, q.iz, h.dist)) { sp = 0; cur = 0; grp = 0; state = ST_NODE; }
          else
          { // the reference's exact semantics on the 4-wide tree; the lane stays idle and takes part in the next refill
            trace_exact_closest<ANALYTIC>(A, r, h);
            uint2 *o2 = reinterpret_cast<uint2 *>(out + ray_i);
            o2[0] = make_uint2(h.prim_lo, h.prim_hi);
            o2[1] = make_uint2(__float_as_uint(h.u), __float_as_uint(h.v));
            o2[2] = make_uint2(__float_as_uint(h.dist), 0u);
          }
        }
      }
    }
    const uint32_t mN = __ballot_sync(FULL, state == ST_NODE);
    const uint32_t mP = __ballot_sync(FULL, state == ST_PRIM);
    if(!(mN | mP))
    {
      if(exhausted) break;
      continue;   // every lane of the last batch took the exact path: fetch again
    }
    const int live = __popc(mN | mP);
    const int thr = prim_threshold >= 0 ? prim_threshold : max(2, (live*(-prim_threshold) + 31) >> 5);
    const bool do_prims = (mN == 0u) || (__popc(mP) >= thr);

    bool stepped = false;
    if(do_prims)
    {
      if(state == ST_PRIM)
      { // the whole leaf in primid[] order
        do
        {
          if(CNT) cnt[3]++;
          prim_intersect<ANALYTIC>(rec, A.rec_units, r, h);
          rec += rec_stride;
        }
        while(--prims_left);
        stepped = true;
      }
    }
    else if(state == ST_NODE)
    {
      const uint32_t m = node8_test<CNT>(A.nodes8 + cur, r, q, h.dist);
      if(CNT) { cnt[1]++; cnt[2] += __popc(m); }
      grp = (cur << 8) | to_visit_order(m, q.octinv);
      stepped = true;
    }
    if(stepped)
    { // next child of the current group, or of the most recent group on the stack
      if(!(grp & 0xffu))
      {
        if(sp > 0) grp = stack[--sp];
        else
        { // ray finished: 24-byte record as three 8-byte stores
          uint2 *o2 = reinterpret_cast<uint2 *>(out + ray_i);
          o2[0] = make_uint2(h.prim_lo, h.prim_hi);
          o2[1] = make_uint2(__float_as_uint(h.u), __float_as_uint(h.v));
          o2[2] = make_uint2(__float_as_uint(h.dist), 0u);
          state = ST_IDLE;
        }
      }
      if(grp & 0xffu)
      {
        const uint32_t pos = 31u - __clz(grp & 0xffu);
        grp ^= 1u << pos;
        const uint32_t ref = __ldg(&A.nodes8[grp >> 8].child[pos ^ q.octinv]);
        if(ref & CB8_LEAF)
        {
          rec = A.recs + (uint64_t)((ref ^ CB8_LEAF) >> 3)*(uint64_t)rec_stride;
          prims_left = ref & 7u;
          state = ST_PRIM;
        }
        else
        {
          if(grp & 0xffu) stack[sp++] = grp;
          cur = ref; grp = 0;
          state = ST_NODE;
        }
      }
    }
  }
  if(CNT)
    for(int k=0;k<4;k++) if(cnt[k]) atomicAdd(counters + k, cnt[k]);
}

// ---------------------------------------------------------------------------------------------
// any hit (accel_visible / path_visible semantics of k_visible, traverse.cu): visiting order is free
// ---------------------------------------------------------------------------------------------
template<bool ANALYTIC, bool SHADOW>
__global__ void __launch_bounds__(TRACE_BLOCK, ANALYTIC ? 1 : TRACE8_MIN_BLOCKS)
k_visible8(DevAccel A, const cb_ray_t *__restrict__ rays, const float *__restrict__ max_dist, const uint2 *__restrict__ skip,
           int32_t *__restrict__ out, uint32_t n, unsigned int *ticket, int prim_threshold, int refill_threshold)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint32_t stack[CB8_STACK];
  int sp = 0;
  int state = ST_IDLE;
  bool exhausted = false;
  RayD r;
  Ray8 q;
  uint32_t ray_i = 0;
  uint32_t cur = 0, grp = 0;
  float md = 0.0f;
  const float4 *rec = nullptr;
  uint32_t prims_left = 0;
  uint2 skip_id = make_uint2(0xffffffffu, 0xffffffffu);
  const uint32_t rec_stride = A.rec_units*4;
  r.px = r.py = r.pz = r.dx = r.dy = r.dz = r.time = r.min_dist = 0.0f; r.ign_lo = r.ign_hi = 0;
  q.ix = q.iy = q.iz = 0.0f; q.octinv = 0;

  while(true)
  {
    const uint32_t idle = __ballot_sync(FULL, state == ST_IDLE);
    if(idle && !exhausted && (__popc(idle) >= refill_threshold || idle == FULL))
    {
      const uint32_t want = __popc(idle);
      unsigned int base = 0;
      if(lane == 0) base = atomicAdd(ticket, want);
      base = __shfl_sync(FULL, base, 0);
      if(base + want >= n) exhausted = true;
      if(state == ST_IDLE)
      {
        const uint32_t i = base + __popc(idle & lt_mask);
        if(i < n)
        {
          load_ray(rays, i, r);
          ray_i = i;
          md = __ldg(max_dist + i);
          if(SHADOW) skip_id = __ldg(skip + i);
          q.ix = 1.0f/r.dx; q.iy = 1.0f/r.dy; q.iz = 1.0f/r.dz;
          q.octinv = 7u ^ ((__float_as_uint(r.dx) >> 31) | ((__float_as_uint(r.dy) >> 31) << 1) | ((__float_as_uint(r.dz) >> 31) << 2));
          if(ray8_ok(r, q.ix, q.iy, q.iz, md)) { sp = 0; cur = 0; grp = 0; state = ST_NODE; }
          else out[ray_i] = trace_exact_any<ANALYTIC, SHADOW>(A, r, md, skip_id);
        }
      }
    }
    const uint32_t mN = __ballot_sync(FULL, state == ST_NODE);
    const uint32_t mP = __ballot_sync(FULL, state == ST_PRIM);
    if(!(mN | mP))
    {
      if(exhausted) break;
      continue;
    }
    const int live = __popc(mN | mP);
    const int thr = prim_threshold >= 0 ? prim_threshold : max(2, (live*(-prim_threshold) + 31) >> 5);
    const bool do_prims = (mN == 0u) || (__popc(mP) >= thr);

    bool stepped = false;
    int result = -1;
    if(do_prims)
    {
      if(state == ST_PRIM)
      {
        do
        {
          if(SHADOW)
          {
            HitD ht;
            ht.dist = md; ht.u = ht.v = 0.0f; ht.prim_lo = ht.prim_hi = 0xffffffffu;
            prim_intersect<ANALYTIC>(rec, A.rec_units, r, ht);
            if((ht.prim_lo & ht.prim_hi) != 0xffffffffu && !(ht.prim_lo == skip_id.x && ht.prim_hi == skip_id.y)) { result = 0; break; }
          }
          else if(prim_visible<ANALYTIC>(rec, A.rec_units, r, md)) { result = 0; break; }
          rec += rec_stride;
        }
        while(--prims_left);
        stepped = true;
      }
    }
    else if(state == ST_NODE)
    {
      grp = (cur << 8) | node8_test<false>(A.nodes8 + cur, r, q, md);   // slot order: any order will do
      stepped = true;
    }
    if(stepped && result < 0)
    {
      if(!(grp & 0xffu))
      {
        if(sp > 0) grp = stack[--sp];
        else result = 1;
      }
      if(grp & 0xffu)
      {
        const uint32_t pos = 31u - __clz(grp & 0xffu);
        grp ^= 1u << pos;
        const uint32_t ref = __ldg(&A.nodes8[grp >> 8].child[pos]);
        if(ref & CB8_LEAF)
        {
          rec = A.recs + (uint64_t)((ref ^ CB8_LEAF) >> 3)*(uint64_t)rec_stride;
          prims_left = ref & 7u;
          state = ST_PRIM;
        }
        else
        {
          if(grp & 0xffu) stack[sp++] = grp;
          cur = ref; grp = 0;
          state = ST_NODE;
        }
      }
    }
    if(result >= 0) { out[ray_i] = result; state = ST_IDLE; }
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
#define LAUNCH8_MAX_RAYS (1ull << 30)

template<bool CNT, bool ANALYTIC>
static int launch_intersect8_k(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                               uint64_t n, cudaStream_t stream, unsigned long long *d_counters)
{
  auto k = k_intersect8<CNT, ANALYTIC>;
  for(uint64_t first=0; first<n; first+=LAUNCH8_MAX_RAYS)
  {
    const uint64_t m = n - first < LAUNCH8_MAX_RAYS ? n - first : LAUNCH8_MAX_RAYS;
    unsigned int *ticket;
    if(int rc = cb200_get_ticket(stream, &ticket)) return rc;
    k<<<cb200_trace_grid(m, (const void *)k), TRACE_BLOCK, 0, stream>>>(a->dev, d_rays + first, d_max_dist ? d_max_dist + first : nullptr, d_out + first,
                                                                       (uint32_t)m, ticket, d_counters, cb200_prim_threshold(), cb200_refill_threshold());
    cb200_count_launch();
    CB_CUDA(cudaGetLastError());
  }
  return 0;
}

int cb200_launch_intersect8(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                            uint64_t n, cudaStream_t stream, unsigned long long *d_counters)
{
  if(n == 0) return 0;
  const bool analytic = a->scene->any_analytic != 0;
  if(d_counters) return launch_intersect8_k<true, true>(a, d_rays, d_max_dist, d_out, n, stream, d_counters);
  return analytic ? launch_intersect8_k<false, true >(a, d_rays, d_max_dist, d_out, n, stream, nullptr)
                  : launch_intersect8_k<false, false>(a, d_rays, d_max_dist, d_out, n, stream, nullptr);
}

template<bool ANALYTIC, bool SHADOW>
static int launch_visible8_k(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, const uint2 *d_skip, int32_t *d_out,
                             uint64_t n, cudaStream_t stream)
{
  auto k = k_visible8<ANALYTIC, SHADOW>;
  for(uint64_t first=0; first<n; first+=LAUNCH8_MAX_RAYS)
  {
    const uint64_t m = n - first < LAUNCH8_MAX_RAYS ? n - first : LAUNCH8_MAX_RAYS;
    unsigned int *ticket;
    if(int rc = cb200_get_ticket(stream, &ticket)) return rc;
    k<<<cb200_trace_grid(m, (const void *)k), TRACE_BLOCK, 0, stream>>>(a->dev, d_rays + first, d_max_dist + first, d_skip ? d_skip + first : nullptr,
                                                                       d_out + first, (uint32_t)m, ticket, cb200_prim_threshold(), cb200_refill_threshold());
    cb200_count_launch();
    CB_CUDA(cudaGetLastError());
  }
  return 0;
}

int cb200_launch_visible8(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, const uint2 *d_light_prim, int32_t *d_out,
                          uint64_t n, cudaStream_t stream)
{
  if(n == 0) return 0;
  const bool analytic = a->scene->any_analytic != 0;
  if(d_light_prim)
    return analytic ? launch_visible8_k<true, true >(a, d_rays, d_max_dist, d_light_prim, d_out, n, stream)
                    : launch_visible8_k<false, true>(a, d_rays, d_max_dist, d_light_prim, d_out, n, stream);
  return analytic ? launch_visible8_k<true, false >(a, d_rays, d_max_dist, nullptr, d_out, n, stream)
                  : launch_visible8_k<false, false>(a, d_rays, d_max_dist, nullptr, d_out, n, stream);
}
