// traverse.cu -- closest-hit and any-hit traversal of the 4-wide BVH (sm_100a).
//
// Semantics follow accel_intersect / accel_visible (src/accel.d/qbvhmp.c:1262-1490):
//   * four slab tests per node on time-interpolated boxes, clipped to [0, hit.dist]
//     (tmin starts at 0, not ray.min_dist), SSE min/max select semantics (second operand wins on
//     NaN), qbvhmp.c:1188-1246;
//   * children visited in the reference's topological order from axis0/axis00/axis01 and the ray's
//     sign bits, the others pushed far->near together with their entry distance, qbvhmp.c:1313-1354;
//   * popped entries whose entry distance exceeds the current hit distance are skipped, :1357-1386;
//   * leaf primitives tested in primid[] order, "dist <= hit.dist" so the last tested wins ties.
//
// Mapping (each item follows an ncu finding, profiles/README.md; plain one-ray-per-thread while-while ran at 4.8 of 32 lanes):
//   * persistent warps; every lane owns one ray at a time and idle lanes are REFILLED from a global ticket counter in
//     batches (>= refill threshold idle lanes, or nothing else left to do), so short rays do not leave lanes idle and the
//     fetch path (an atomic round trip + cold ray loads) does not run for one or two lanes on every iteration.  (A per-warp
//     ray queue in shared memory -- 32 rays per ticket, lanes pop individually -- was measured in round 2: it keeps 30 of 32
//     lanes busy, and is 10 % SLOWER: the kernel is bound by L1 wavefronts per ray, which lane packing does not change, and
//     the 45 KB of shared memory per SM come out of the L1 that serves the node gathers;  profiles/README.md);
//   * each lane is in one of two states, NODE (next: a 4-box node test) or PRIM (next: all primitives of its current
//     leaf); per warp iteration a vote picks which of the two code paths runs, so both execute with many lanes active.
//     Every ray still sees exactly the reference's own sequence of node and primitive tests;
//   * static nodes + rays with finite non-zero direction: the near / far plane of each axis is selected when the row is
//     loaded, so min(lo,hi) / max(lo,hi) disappear (identical tmin / tmax / hit mask); motion-blur nodes the same after the
//     reference's time interpolation; all other rays take the exact SSE-select form;
//   * the visiting order is three conditional swaps on (entry distance, child) pairs;
//   * child references are 32 bits inside the kernel while the scene allows (< 2^26 primitives): one 8-byte stack word per
//     entry; kernel variants by scene content (motion blur, analytic primitives, instrumentation).
#include <atomic>
#include <mutex>
#include <cstdlib>
#include <type_traits>

#include "traverse_common.cuh"

#define SEL4(arr, i) ((i) == 0 ? arr[0] : (i) == 1 ? arr[1] : (i) == 2 ? arr[2] : arr[3])


// Fast slab test for static nodes and rays whose direction components are all finite and non-zero: the near / far plane
// of every axis is picked by the ray's sign when the row is LOADED (rows 0..2 = min, 3..5 = max of Node128::aabb0), so that
// lo <= hi holds by monotonicity of IEEE subtraction and multiplication and min(lo,hi) / max(lo,hi) of the reference
// (qbvhmp.c:1222-1223) need not be computed: identical tmin, tmax and hit mask for every non-empty box.  (An inverted, empty
// box "hits" in the reference and misses here; empty leaves are never visited either way.)
// key[c] = entry distance when child c is hit, NaN otherwise (a hit implies tmin <= tmax, so tmin is never NaN itself;
// it CAN be negative for rays with NaN slab products, where the reference's select semantics drop the clip at 0).
__device__ __forceinline__ void node_slabs_fast(const Node128 *__restrict__ n, const uint32_t near_off[3], float px, float py, float pz,
                                                float ix, float iy, float iz, float tmax_init, float key[4])
{
  const float4 *a0 = reinterpret_cast<const float4 *>(n->aabb0);
  const float4 nx = __ldg(a0 + near_off[0]),     ny = __ldg(a0 + 1 + near_off[1]),     nz = __ldg(a0 + 2 + near_off[2]);
  const float4 fx = __ldg(a0 + 3 - near_off[0]), fy = __ldg(a0 + 4 - near_off[1]),     fz = __ldg(a0 + 5 - near_off[2]);
#ifdef CB200_SCALAR_SLABS
  const float nxa[4] = {nx.x, nx.y, nx.z, nx.w}, nya[4] = {ny.x, ny.y, ny.z, ny.w}, nza[4] = {nz.x, nz.y, nz.z, nz.w};
  const float fxa[4] = {fx.x, fx.y, fx.z, fx.w}, fya[4] = {fy.x, fy.y, fy.z, fy.w}, fza[4] = {fz.x, fz.y, fz.z, fz.w};
#pragma unroll
  for(int c=0;c<4;c++)
  {
    const float tmin = fmaxf(fmaxf(0.0f, (nxa[c] - px)*ix), fmaxf((nya[c] - py)*iy, (nza[c] - pz)*iz));
    const float tmax = fminf(fminf(tmax_init, (fxa[c] - px)*ix), fminf((fya[c] - py)*iy, (fza[c] - pz)*iz));
    key[c] = tmin <= tmax ? tmin : KEY_MISS;
  }
#else
  // the 24 plane distances as 12 packed subtractions + 12 packed multiplications (two children per instruction), the entry /
  // exit distances with the three-input min / max: no operand can be NaN here (finite non-zero 1/d, finite origin, finite planes)
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
#endif
}

// the same for nodes with shutter-open and shutter-close boxes: every plane is first interpolated to the ray's time exactly
// like aabb_intersect does (box0*t0 + box1*t1, qbvhmp.c:1206-1215)
__device__ __forceinline__ void node_slabs_fast_mb(const Node256 *__restrict__ n, const uint32_t near_off[3], float px, float py, float pz,
                                                   float ix, float iy, float iz, float t0, float t1, float tmax_init, float key[4])
{
  const float4 *a0 = reinterpret_cast<const float4 *>(n->aabb0), *a1 = reinterpret_cast<const float4 *>(n->aabb1);
  float4 q[6], w[6];
  q[0] = __ldg(a0 + near_off[0]);     w[0] = __ldg(a1 + near_off[0]);
  q[1] = __ldg(a0 + 1 + near_off[1]); w[1] = __ldg(a1 + 1 + near_off[1]);
  q[2] = __ldg(a0 + 2 + near_off[2]); w[2] = __ldg(a1 + 2 + near_off[2]);
  q[3] = __ldg(a0 + 3 - near_off[0]); w[3] = __ldg(a1 + 3 - near_off[0]);
  q[4] = __ldg(a0 + 4 - near_off[1]); w[4] = __ldg(a1 + 4 - near_off[1]);
  q[5] = __ldg(a0 + 5 - near_off[2]); w[5] = __ldg(a1 + 5 - near_off[2]);
  float d[6][4];   // plane distances: rows 0..2 entry (x, y, z), 3..5 exit; two children per packed instruction
  const float pos[3] = {px, py, pz}, inv[3] = {ix, iy, iz};
#pragma unroll
  for(int k=0;k<6;k++)
  {
    lerp_slab2(q[k].x, q[k].y, w[k].x, w[k].y, t0, t1, pos[k % 3], inv[k % 3], d[k][0], d[k][1]);
    lerp_slab2(q[k].z, q[k].w, w[k].z, w[k].w, t0, t1, pos[k % 3], inv[k % 3], d[k][2], d[k][3]);
  }
#pragma unroll
  for(int c=0;c<4;c++)
  {
    const float tmin = fmaxf(max3f(d[0][c], d[1][c], d[2][c]), 0.0f);
    const float tmax = fminf(min3f(d[3][c], d[4][c], d[5][c]), tmax_init);
    key[c] = tmin <= tmax ? tmin : KEY_MISS;
  }
}

// 1 when the child was hit (its key is an entry distance), 0 when the key is KEY_MISS (NaN); a word select.  Inline PTX so that the
// optimiser sees neither a comparison chain nor a boolean it could branch on (k_intersect's node step, below).
__device__ __forceinline__ uint32_t hit_bit(float k)
{
  uint32_t r;
  asm("{ .reg .pred q;\n setp.eq.f32 q, %1, %1;\n selp.u32 %0, 1, 0, q; }" : "=r"(r) : "f"(k));
  return r;
}
__device__ __forceinline__ uint32_t sel32(uint32_t c, uint32_t a, uint32_t b)
{
  uint32_t r;
  asm("{ .reg .pred q;\n setp.ne.u32 q, %1, 0;\n selp.b32 %0, %2, %3, q; }" : "=r"(r) : "r"(c), "r"(a), "r"(b));
  return r;
}

#define CSWAP(cond, ka, ca, kb, cb) do { const float tk__ = ka; const ref_t tc__ = ca; \
  ka = (cond) ? kb : ka; ca = (cond) ? cb : ca; kb = (cond) ? tk__ : kb; cb = (cond) ? tc__ : cb; } while(0)


// ---------------------------------------------------------------------------------------------
// closest hit
//   ANALYTIC: the scene has spheres / cylinders / cones (their tests carry double precision and libm calls; scenes
//             without them get a kernel without that code and with fewer registers)
// ---------------------------------------------------------------------------------------------
//   C32     : child references are handled as 32 bits inside the kernel (leaf flag in bit 31, begin<<5|count or the node index
//             below; possible while the scene has < 2^26 primitives): the (entry distance, child) pair of a stack entry is
//             then ONE 8-byte word -- stack pushes/pops were 22 % of this kernel's instructions -- and the ordering
//             swaps move half the registers.  Larger scenes use the 64-bit form.
template<bool MB, bool CNT, int STACK, bool ANALYTIC, bool C32>
__global__ void __launch_bounds__(TRACE_BLOCK, (!MB && !CNT && !ANALYTIC) ? TRACE_MIN_BLOCKS : 1)
k_intersect(DevAccel A, const cb_ray_t *__restrict__ rays, const float *__restrict__ max_dist,
            cb_hitrec_t *__restrict__ out, uint32_t n, unsigned int *ticket, unsigned long long *counters,
            int prim_threshold, int refill_threshold)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  typedef typename std::conditional<C32, uint32_t, uint64_t>::type ref_t;
  unsigned long long cnt[4] = {0, 0, 0, 0};
  uint64_t stack[STACK];                      // C32: (child << 32) | entry distance bits
  float stack_dist[C32 ? 1 : STACK];
  int sp = 0;
  int state = ST_IDLE;
  bool exhausted = false;
  RayD r;
  HitD h;
  uint32_t ray_i = 0;
  ref_t cur = 0;
  float ix = 0.0f, iy = 0.0f, iz = 0.0f, t0 = 1.0f, t1 = 0.0f;
  uint32_t nearbits = 0;
  uint32_t near_off[3] = {0, 0, 0};
  bool exact = false;
  const uint32_t rec_stride = A.rec_units*4;
  r.px = r.py = r.pz = r.dx = r.dy = r.dz = r.time = r.min_dist = 0.0f; r.ign_lo = r.ign_hi = 0;
  h.dist = 0.0f; h.u = h.v = 0.0f; h.prim_lo = h.prim_hi = 0;

  while(true)
  {
    // ---- refill idle lanes from the global ray queue, but only once enough of them have gathered (or nothing else is
    //      left to do): the fetch code would otherwise run on nearly every iteration for one or two lanes
    uint32_t mN = __ballot_sync(FULL, state == ST_NODE);
    const uint32_t mP = __ballot_sync(FULL, state == ST_PRIM);
    const uint32_t idle = ~(mN | mP);
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
          nearbits = (__float_as_uint(r.dx) >> 31) | ((__float_as_uint(r.dy) >> 31) << 1) | ((__float_as_uint(r.dz) >> 31) << 2);
          near_off[0] = 3u*(nearbits & 1u); near_off[1] = 3u*((nearbits >> 1) & 1u); near_off[2] = 3u*(nearbits >> 2);
          ix = 1.0f/r.dx; iy = 1.0f/r.dy; iz = 1.0f/r.dz;
          t1 = r.time; t0 = 1.0f - r.time;
          exact = !(finite_nonzero(ix) && finite_nonzero(iy) && finite_nonzero(iz) &&
                    finite(r.px) && finite(r.py) && finite(r.pz) && finite(r.time) && !(h.dist != h.dist));
          sp = 0; cur = 0; state = ST_NODE;
          if(CNT) cnt[0]++;
        }
      }
      mN = __ballot_sync(FULL, state == ST_NODE);   // the new rays start at the root
    }
    if(!(mN | mP)) break;   // nothing in flight: the refill above ran (all lanes idle) and the queue is empty
    // prim_threshold < 0: relative to the lanes that hold a ray (-12 = 12/32 of them), so that a warp waiting for its next
    // refill with many finished lanes does not wait for nearly all remaining lanes to reach a leaf
    const int live = __popc(mN | mP);
    const int thr = prim_threshold >= 0 ? prim_threshold : max(2, (live*(-prim_threshold) + 31) >> 5);
    const bool do_prims = (mN == 0u) || (__popc(mP) >= thr);

    bool need_pop = false, new_cur = false;
    // (loading the stack entry a pop will look at first BEFORE the primitive tests / the node's loads, so that the local-memory
    // latency -- 14 % of this kernel's stall samples sit on the pop's comparison -- hides behind them: 86.9 -> 88.3 / 86.9 ms per 8
    // progressions, profiles/r3l: the other warps already cover that stall)
    if(do_prims)
    {
      if(state == ST_PRIM)
      { // the whole leaf in primid[] order (qbvhmp.c:1371-1379); empty leaves are never pushed, so there is at least one primitive
        const ref_t leaf_bit = C32 ? (ref_t)0x80000000u : (ref_t)CB_LEAF_BIT;
        const float4 *rec = A.recs + (uint64_t)((cur ^ leaf_bit) >> 5)*(uint64_t)rec_stride;
        uint32_t prims_left = (uint32_t)cur & 31u;
        do
        {
          if(CNT) cnt[3]++;
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
      ref_t child[4];
      int axis0, axis00, axis01;
      if(MB && !CNT && !exact)
      { // motion-blur nodes, ordinary ray
        const Node256 *nd = reinterpret_cast<const Node256 *>(A.nodes) + cur;
        node_slabs_fast_mb(nd, near_off, r.px, r.py, r.pz, ix, iy, iz, t0, t1, h.dist, key);
        const ulonglong2 *ch = reinterpret_cast<const ulonglong2 *>(nd->child);
        const ulonglong2 c01 = __ldg(ch), c23 = __ldg(ch + 1), pa = __ldg(ch + 2), aa = __ldg(ch + 3);
        const uint64_t c64[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
        for(int c=0;c<4;c++)
        { // imported reference trees may carry empty leaves with ordinary boxes; this library's builder inverts them
          if(A.imported && is_empty_leaf(c64[c])) key[c] = KEY_MISS;
          child[c] = C32 ? (ref_t)((uint32_t)c64[c] | (uint32_t)(c64[c] >> 32)) : (ref_t)c64[c];
        }
        axis0 = (int)pa.y; axis00 = (int)aa.x; axis01 = (int)aa.y;
      }
      else if(MB || CNT || exact)
      {
        NodeOut o;
        if(exact) node_slabs<MB, true >(A.nodes, cur, r.px, r.py, r.pz, ix, iy, iz, t0, t1, h.dist, o);
        else      node_slabs<MB, false>(A.nodes, cur, r.px, r.py, r.pz, ix, iy, iz, t0, t1, h.dist, o);
        if(CNT && (o.hit[0] | o.hit[1] | o.hit[2] | o.hit[3])) { cnt[1]++; for(int c=0;c<4;c++) cnt[2] += o.hit[c] ? 1 : 0; }
        // empty leaves (count 0) can only be popped and dropped again: never visit them
#pragma unroll
        for(int c=0;c<4;c++)
        {
          key[c] = (o.hit[c] && !is_empty_leaf(o.child[c])) ? o.tmin[c] : KEY_MISS;
          child[c] = C32 ? (ref_t)((uint32_t)o.child[c] | (uint32_t)(o.child[c] >> 32)) : (ref_t)o.child[c];
        }
        axis0 = o.axis0; axis00 = o.axis00; axis01 = o.axis01;
      }
      else
      { // no empty-leaf test needed here: this library's builder gives empty slots an inverted box, which the sign-selected
        // slab test misses by construction (imported reference trees are Node256 and take the other path)
        const Node128 *nd = reinterpret_cast<const Node128 *>(A.nodes) + cur;
        node_slabs_fast(nd, near_off, r.px, r.py, r.pz, ix, iy, iz, h.dist, key);
        const uint4 *ch = reinterpret_cast<const uint4 *>(nd->child);
        const uint4 c01 = __ldg(ch), c23 = __ldg(ch + 1);   // {lo0, hi0, lo1, hi1}, {lo2, hi2, lo3, hi3}
        const uint32_t ax = (c01.y >> (CB_AXIS_SHIFT - 32)) & 63u;
        const uint32_t hi0 = c01.y & (uint32_t)(CB_CHILD_MASK >> 32);
        if(C32) { child[0] = (ref_t)(c01.x | hi0); child[1] = (ref_t)(c01.z | c01.w); child[2] = (ref_t)(c23.x | c23.y); child[3] = (ref_t)(c23.z | c23.w); }
        else
        {
          child[0] = (ref_t)(((uint64_t)hi0 << 32) | c01.x);   child[1] = (ref_t)(((uint64_t)c01.w << 32) | c01.z);
          child[2] = (ref_t)(((uint64_t)c23.y << 32) | c23.x); child[3] = (ref_t)(((uint64_t)c23.w << 32) | c23.z);
        }
        axis0 = ax & 3; axis00 = (ax >> 2) & 3; axis01 = (ax >> 4) & 3;
      }
      // the reference's topological order (qbvhmp.c:1313-1320) as three conditional swaps: inside the lower pair by the
      // sign along axis00, inside the upper pair by the sign along axis01, the two pairs by the sign along axis0
      const bool s00 = (nearbits >> axis00) & 1u, s01 = (nearbits >> axis01) & 1u, s0 = (nearbits >> axis0) & 1u;
      CSWAP(s00, key[0], child[0], key[1], child[1]);
      CSWAP(s01, key[2], child[2], key[3], child[3]);
      CSWAP(s0,  key[0], child[0], key[2], child[2]);
      CSWAP(s0,  key[1], child[1], key[3], child[3]);
      // nearest hit child becomes current, the others are pushed far -> near with their entry distance.  One straight-line block
      // for all lanes of the node step: the source form "first = hit0 ? 0 : hit1 ? 1 : ...; if(first < 4) { pushes; cur = child[first]; }"
      // compiled into a chain of divergent branches -- one path per value of `first`, ~75 instructions at 5-9 lanes that cost as
      // many issue slots as the four slab tests in front of them (profiles/r3e, source view).  The hit flags are made opaque 0 / 1
      // words, pushes are predicated stores, the new current child is a chain of selects.
      const uint32_t h0 = hit_bit(key[0]), h1 = hit_bit(key[1]), h2 = hit_bit(key[2]), h3 = hit_bit(key[3]);
      const uint32_t h01 = h0 | h1, h012 = h01 | h2, any = h012 | h3;
      const uint32_t p3 = h3 & h012, p2 = h2 & h01, p1 = h1 & h0;
#define PUSH_IF(p, k) do { if(p) { if(C32) stack[sp] = ((uint64_t)child[k] << 32) | __float_as_uint(key[k]); \
                                  else { stack_dist[sp] = key[k]; stack[sp] = child[k]; } } sp += (int)(p); } while(0)
      PUSH_IF(p3, 3);
      PUSH_IF(p2, 2);
      PUSH_IF(p1, 1);
#undef PUSH_IF
      if(C32) cur = (ref_t)sel32(h0, (uint32_t)child[0], sel32(h1, (uint32_t)child[1], sel32(h2, (uint32_t)child[2], sel32(h3, (uint32_t)child[3], (uint32_t)cur))));
      else    cur = h0 ? child[0] : h1 ? child[1] : h2 ? child[2] : h3 ? child[3] : cur;
      need_pop = any == 0u;
      new_cur = any != 0u;
    }
    if(need_pop)
    {
      while(sp > 0)
      {
        --sp;
        if(C32)
        {
          const uint64_t e = stack[sp];
          if(__uint_as_float((uint32_t)e) > h.dist) continue;
          cur = (ref_t)(e >> 32);
        }
        else
        {
          if(stack_dist[sp] > h.dist) continue;
          cur = (ref_t)stack[sp];
        }
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
    // a leaf reference (top bit) sends the lane to the primitive step, which works out the record address itself
    if(new_cur) state = (int)(1u + (uint32_t)(cur >> (C32 ? 31 : 63)));   // ST_NODE = 1, ST_PRIM = 2
  }
  if(CNT)
    for(int k=0;k<4;k++) if(cnt[k]) atomicAdd(counters + k, cnt[k]);
}

// ---------------------------------------------------------------------------------------------
// any hit: returns 1 when nothing blocks the ray up to max_dist (accel_visible semantics; the result
// is a boolean, so visiting order is free)
// ---------------------------------------------------------------------------------------------
//   SHADOW: next-event visibility in path_visible's terms (src/pathspace.c:311-344), which asks accel_intersect -- not
//           accel_visible -- whether anything lies in front of the light: primitives are tested with the CLOSEST-hit rules
//           (ray.ignore honoured, dist > min_dist && dist <= limit, strict < for analytic prims) and the sampled light
//           primitive skip[i] never occludes.  The caller has already clipped max_dist to the first crossing of the light
//           primitive itself, so "any accepted primitive" == "the closest hit is not the light".
template<bool MB, int STACK, bool ANALYTIC, bool SHADOW, bool C32>
__global__ void __launch_bounds__(TRACE_BLOCK, (!MB && !ANALYTIC) ? TRACE_MIN_BLOCKS : 1)
k_visible(DevAccel A, const cb_ray_t *__restrict__ rays, const float *__restrict__ max_dist, const uint2 *__restrict__ skip,
          int32_t *__restrict__ out, uint32_t n, unsigned int *ticket, int prim_threshold, int refill_threshold)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  typedef typename std::conditional<C32, uint32_t, uint64_t>::type ref_t;   // see k_intersect
  ref_t stack[STACK];
  int sp = 0;
  int state = ST_IDLE;
  bool exhausted = false;
  RayD r;
  uint32_t ray_i = 0;
  ref_t cur = 0;
  float ix = 0.0f, iy = 0.0f, iz = 0.0f, t0 = 1.0f, t1 = 0.0f, md = 0.0f;
  uint32_t near_off[3] = {0, 0, 0};
  bool exact = false;
  uint2 skip_id = make_uint2(0xffffffffu, 0xffffffffu);
  const uint32_t rec_stride = A.rec_units*4;
  r.px = r.py = r.pz = r.dx = r.dy = r.dz = r.time = r.min_dist = 0.0f; r.ign_lo = r.ign_hi = 0;

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
          const uint32_t nearbits = (__float_as_uint(r.dx) >> 31) | ((__float_as_uint(r.dy) >> 31) << 1) | ((__float_as_uint(r.dz) >> 31) << 2);
          near_off[0] = 3u*(nearbits & 1u); near_off[1] = 3u*((nearbits >> 1) & 1u); near_off[2] = 3u*(nearbits >> 2);
          ix = 1.0f/r.dx; iy = 1.0f/r.dy; iz = 1.0f/r.dz;
          t1 = r.time; t0 = 1.0f - r.time;
          exact = !(finite_nonzero(ix) && finite_nonzero(iy) && finite_nonzero(iz) &&
                    finite(r.px) && finite(r.py) && finite(r.pz) && finite(r.time) && !(md != md));
          sp = 0; cur = 0; state = ST_NODE;
        }
      }
    }
    const uint32_t mN = __ballot_sync(FULL, state == ST_NODE);
    const uint32_t mP = __ballot_sync(FULL, state == ST_PRIM);
    if(!(mN | mP)) break;
    // prim_threshold < 0: relative to the lanes that hold a ray (-12 = 12/32 of them), so that a warp waiting for its next
    // refill with many finished lanes does not wait for nearly all remaining lanes to reach a leaf
    const int live = __popc(mN | mP);
    const int thr = prim_threshold >= 0 ? prim_threshold : max(2, (live*(-prim_threshold) + 31) >> 5);
    const bool do_prims = (mN == 0u) || (__popc(mP) >= thr);
    bool need_pop = false;
    int result = -1;
    if(do_prims)
    {
      if(state == ST_PRIM)
      {
        const ref_t leaf_bit = C32 ? (ref_t)0x80000000u : (ref_t)CB_LEAF_BIT;
        const float4 *rec = A.recs + (uint64_t)((cur ^ leaf_bit) >> 5)*(uint64_t)rec_stride;
        uint32_t prims_left = (uint32_t)cur & 31u;
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
        need_pop = true;
      }
    }
    else if(state == ST_NODE)
    {
      if(MB && !exact)
      {
        const Node256 *nd = reinterpret_cast<const Node256 *>(A.nodes) + cur;
        float key[4];
        node_slabs_fast_mb(nd, near_off, r.px, r.py, r.pz, ix, iy, iz, t0, t1, md, key);
        const ulonglong2 *ch = reinterpret_cast<const ulonglong2 *>(nd->child);
        const ulonglong2 c01 = __ldg(ch), c23 = __ldg(ch + 1);
        const uint64_t c64[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
        for(int c=0;c<4;c++) if(KEY_HIT(key[c]) && !(A.imported && is_empty_leaf(c64[c])))
          stack[sp++] = C32 ? (ref_t)((uint32_t)c64[c] | (uint32_t)(c64[c] >> 32)) : (ref_t)c64[c];
      }
      else if(MB || exact)
      {
        NodeOut o;
        if(exact) node_slabs<MB, true >(A.nodes, cur, r.px, r.py, r.pz, ix, iy, iz, t0, t1, md, o);
        else      node_slabs<MB, false>(A.nodes, cur, r.px, r.py, r.pz, ix, iy, iz, t0, t1, md, o);
#pragma unroll
        for(int c=0;c<4;c++) if(o.hit[c] && !is_empty_leaf(o.child[c]))
          stack[sp++] = C32 ? (ref_t)((uint32_t)o.child[c] | (uint32_t)(o.child[c] >> 32)) : (ref_t)o.child[c];
      }
      else
      {
        const Node128 *nd = reinterpret_cast<const Node128 *>(A.nodes) + cur;
        float key[4];
        node_slabs_fast(nd, near_off, r.px, r.py, r.pz, ix, iy, iz, md, key);
        const uint4 *ch = reinterpret_cast<const uint4 *>(nd->child);
        const uint4 c01 = __ldg(ch), c23 = __ldg(ch + 1);
        const uint32_t hi0 = c01.y & (uint32_t)(CB_CHILD_MASK >> 32);
        ref_t child[4];
        if(C32) { child[0] = (ref_t)(c01.x | hi0); child[1] = (ref_t)(c01.z | c01.w); child[2] = (ref_t)(c23.x | c23.y); child[3] = (ref_t)(c23.z | c23.w); }
        else
        {
          child[0] = (ref_t)(((uint64_t)hi0 << 32) | c01.x);   child[1] = (ref_t)(((uint64_t)c01.w << 32) | c01.z);
          child[2] = (ref_t)(((uint64_t)c23.y << 32) | c23.x); child[3] = (ref_t)(((uint64_t)c23.w << 32) | c23.z);
        }
#pragma unroll
        for(int c=0;c<4;c++) if(KEY_HIT(key[c])) stack[sp++] = child[c];   // empty slots carry inverted boxes: never hit here
      }
      need_pop = true;
    }
    if(result < 0 && need_pop)
    {
      if(sp > 0)
      {
        cur = stack[--sp];
        state = (int)(1u + (uint32_t)(cur >> (C32 ? 31 : 63)));   // ST_NODE, or ST_PRIM for a leaf reference (the primitive step works out the record address)
      }
      else result = 1;
    }
    if(result >= 0) { out[ray_i] = result; state = ST_IDLE; }
  }
}

// ---------------------------------------------------------------------------------------------
// accel_closest (qbvhmp.c:1493-1600): the hit nearest to `centre` along the ray, for the half-vector samplers.  One query
// per thread in the reference's exact order (boxes clipped to [min_dist, hit.dist], no entry-distance culling, the search
// interval re-centred after every primitive test); not a throughput path.
// ---------------------------------------------------------------------------------------------
template<bool MB, int STACK>
__global__ void __launch_bounds__(TRACE_BLOCK)
k_closest(DevAccel A, cb_ray_t *__restrict__ rays, cb_hitrec_t *__restrict__ io, const float *__restrict__ centre_in, uint64_t n)
{
  const uint64_t i = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x;
  if(i >= n) return;
  RayD r;
  load_ray(rays, i, r);
  HitD h, tent;
  h.prim_lo = io[i].prim[0]; h.prim_hi = io[i].prim[1]; h.u = io[i].u; h.v = io[i].v; h.dist = io[i].dist;
  tent = h;
  const float centre = centre_in[i];
  const uint32_t nearbits = (__float_as_uint(r.dx) >> 31) | ((__float_as_uint(r.dy) >> 31) << 1) | ((__float_as_uint(r.dz) >> 31) << 2);
  const float ix = 1.0f/r.dx, iy = 1.0f/r.dy, iz = 1.0f/r.dz, t1 = r.time, t0 = 1.0f - r.time;
  uint64_t stack[STACK];
  int sp = 0;
  uint64_t cur = 0;
  const uint32_t rec_stride = A.rec_units*4;
  bool done = false, straight = false;
  while(!done)
  {
    NodeOut o;
    node_slabs<MB, true>(A.nodes, cur, r.px, r.py, r.pz, ix, iy, iz, t0, t1, h.dist, o, r.min_dist);
    const uint32_t n0 = (nearbits >> o.axis0) & 1u;
    const int axis1n = n0 ? o.axis01 : o.axis00, axis1f = n0 ? o.axis00 : o.axis01;
    const uint32_t n1n = (nearbits >> axis1n) & 1u, n1f = (nearbits >> axis1f) & 1u, f0 = n0 ^ 1u;
    const uint32_t n11 = (f0 << 1) | (n1f ^ 1u), n10 = (f0 << 1) | n1f, n01 = (n0 << 1) | (n1n ^ 1u), n00 = (n0 << 1) | n1n;
    const bool h00 = SEL4(o.hit, n00), h01 = SEL4(o.hit, n01), h10 = SEL4(o.hit, n10), h11 = SEL4(o.hit, n11);
    const int first = h00 ? 0 : h01 ? 1 : h10 ? 2 : h11 ? 3 : 4;
    if(first < 4)
    {
      if(h11 && first < 3) stack[sp++] = SEL4(o.child, n11);
      if(h10 && first < 2) stack[sp++] = SEL4(o.child, n10);
      if(h01 && first < 1) stack[sp++] = SEL4(o.child, n01);
      const uint32_t nf = first == 0 ? n00 : first == 1 ? n01 : first == 2 ? n10 : n11;
      cur = SEL4(o.child, nf);
    }
    else
    {
      if(sp == 0) break;
      cur = stack[--sp];
    }
    while(cur & CB_LEAF_BIT)
    {
      const float4 *rec = A.recs + ((cur ^ CB_LEAF_BIT) >> 5)*(uint64_t)rec_stride;
      const uint32_t num = (uint32_t)cur & 31u;
      for(uint32_t k=0;k<num && !done;k++, rec += rec_stride)
      {
        prim_intersect<true>(rec, A.rec_units, r, h);
        if(fabsf(h.dist - centre) < fabsf(tent.dist - centre)) tent = h;
        if(h.dist > centre + 1e-6f) r.min_dist = 2.0f*centre - h.dist;
        else if(h.dist < centre - 1e-6f) { r.min_dist = h.dist; h.dist = 2.0f*centre - r.min_dist; }
        else { done = true; straight = true; }
      }
      if(done) break;
      if(sp == 0) { done = true; break; }
      cur = stack[--sp];
    }
  }
  if(!straight) h = tent;
  io[i].prim[0] = h.prim_lo; io[i].prim[1] = h.prim_hi; io[i].u = h.u; io[i].v = h.v; io[i].dist = h.dist; io[i].pad = 0;
  rays[i].min_dist = r.min_dist;
}

int cb200_launch_closest(const cb200_accel *a, cb_ray_t *d_rays, cb_hitrec_t *d_io, const float *d_centre, uint64_t n, cudaStream_t stream)
{
  if(n == 0) return 0;
  if(3*a->depth + 1 > 304) { cb200_set_error("tree deeper than the reference's MAX_TREE_DEPTH"); return CB200_ERR_UNSUPPORTED; }
  const unsigned grid = (unsigned)((n + TRACE_BLOCK - 1)/TRACE_BLOCK);
  if(a->dev.mb) k_closest<true,  304><<<grid, TRACE_BLOCK, 0, stream>>>(a->dev, d_rays, d_io, d_centre, n);
  else          k_closest<false, 304><<<grid, TRACE_BLOCK, 0, stream>>>(a->dev, d_rays, d_io, d_centre, n);
  cb200_count_launch();
  CB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
// ticket counters: one ring per device (a launch takes the next counter of the CURRENT device's ring and zeroes it on its own
// stream, so launches on different streams never share a live counter while fewer than NUM_TICKETS are in flight per device)
#define NUM_TICKETS 1024
#define MAX_DEVICES 64
static unsigned int *g_tickets[MAX_DEVICES] = {nullptr};
static std::atomic<unsigned> g_ticket_next[MAX_DEVICES];
static std::mutex g_ticket_mutex;
static int g_prim_threshold = 1000;

int cb200_prim_threshold()
{
  if(g_prim_threshold == 1000)
  {
    const char *e = getenv("CB200_PRIM_THRESHOLD");
    int v = e ? atoi(e) : -16;   // swept on the 10 M-triangle bench: absolute 12 -> 104.5 ms of closest-hit time per 8 progressions, relative 10/12/16/20/24/28 of 32 -> 103.8/102.4/101.4/102.5/105.9/115.0
    if(v == 0) v = 1;
    if(v > 32) v = 32;
    if(v < -32) v = -32;
    g_prim_threshold = v;
  }
  return g_prim_threshold;
}

int cb200_get_ticket(cudaStream_t stream, unsigned int **t)
{
  int dev = 0;
  CB_CUDA(cudaGetDevice(&dev));
  if(dev < 0 || dev >= MAX_DEVICES) { cb200_set_error("device index beyond the ticket table"); return CB200_ERR_UNSUPPORTED; }
  {
    std::lock_guard<std::mutex> lock(g_ticket_mutex);
    if(!g_tickets[dev]) CB_CUDA(cudaMalloc(&g_tickets[dev], sizeof(unsigned int)*NUM_TICKETS));
  }
  unsigned int *p = g_tickets[dev] + (g_ticket_next[dev]++ % NUM_TICKETS);
  CB_CUDA(cudaMemsetAsync(p, 0, sizeof(unsigned int), stream));
  *t = p;
  return 0;
}

int cb200_trace_grid(uint64_t n, const void *kernel)
{
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TRACE_BLOCK, 0);
  if(per_sm < 1) per_sm = 1;
  const uint64_t want = (n + TRACE_BLOCK - 1)/TRACE_BLOCK;
  const uint64_t full = (uint64_t)cb200_sm_count_cached()*per_sm;   // a multiple of the SM count: one resident wave
  return (int)(want < full ? (want ? want : 1) : full);
}

// the stack must hold 3 entries per tree level (qbvhmp.c:1277); pick the smallest variant that fits
// rays per launch: tickets are 32 bits and every warp draws one batch beyond the end
#define LAUNCH_MAX_RAYS (1ull << 30)

// CB200_FORCE_REF64=1 runs the 64-bit child-reference kernels that scenes with >= 2^26 primitives get (tests)
static bool force_ref64() { static int v = -1; if(v < 0) { const char *e = getenv("CB200_FORCE_REF64"); v = (e && atoi(e)) ? 1 : 0; } return v == 1; }

static int g_refill_threshold = -1;
int cb200_refill_threshold()
{
  if(g_refill_threshold < 0)
  {
    const char *e = getenv("CB200_REFILL_THRESHOLD");
    int v = e ? atoi(e) : 24;   // swept 4..32 on the 10 M-triangle bench: 16-20 was the flat optimum with pool-sized waves, 24 with the ~20 M-ray waves of the larger pool (profiles/r3m)
    if(v < 1) v = 1;
    if(v > 32) v = 32;
    g_refill_threshold = v;
  }
  return g_refill_threshold;
}

template<bool MB, bool CNT, int STACK, bool ANALYTIC, bool C32>
static int launch_intersect_k2(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                              uint64_t n, cudaStream_t stream, unsigned long long *d_counters)
{
  auto k = k_intersect<MB, CNT, STACK, ANALYTIC, C32>;
  for(uint64_t first=0; first<n; first+=LAUNCH_MAX_RAYS)
  {
    const uint64_t m = n - first < LAUNCH_MAX_RAYS ? n - first : LAUNCH_MAX_RAYS;
    unsigned int *ticket;
    if(int rc = cb200_get_ticket(stream, &ticket)) return rc;
    k<<<cb200_trace_grid(m, (const void *)k), TRACE_BLOCK, 0, stream>>>(a->dev, d_rays + first, d_max_dist ? d_max_dist + first : nullptr, d_out + first,
                                                                (uint32_t)m, ticket, d_counters, cb200_prim_threshold(), cb200_refill_threshold());
    cb200_count_launch();
    CB_CUDA(cudaGetLastError());
  }
  return 0;
}

template<bool MB, bool CNT, int STACK, bool ANALYTIC>
static int launch_intersect_k(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                              uint64_t n, cudaStream_t stream, unsigned long long *d_counters)
{
  // 32-bit child references inside the kernel while begin<<5|count and node indices fit 31 bits
  const bool c32 = !CNT && a->dev.num_prims < (1ull << 26) && a->dev.num_nodes < (1ull << 31) && !force_ref64();
  if(c32) return launch_intersect_k2<MB, CNT, STACK, ANALYTIC, !CNT>(a, d_rays, d_max_dist, d_out, n, stream, d_counters);
  return launch_intersect_k2<MB, CNT, STACK, ANALYTIC, false>(a, d_rays, d_max_dist, d_out, n, stream, d_counters);
}

template<bool MB>
static int launch_intersect_t(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                              uint64_t n, cudaStream_t stream, unsigned long long *d_counters)
{
  const int need = 3*a->depth + 1;
  if(need > STACK_BIG) { cb200_set_error("tree deeper than the reference's MAX_TREE_DEPTH"); return CB200_ERR_UNSUPPORTED; }
  if(d_counters) return launch_intersect_k<MB, true, STACK_BIG, true>(a, d_rays, d_max_dist, d_out, n, stream, d_counters);
  const bool analytic = a->scene->any_analytic != 0;
  // static scenes with 32-bit child references: two rays per lane (traverse2.cu)
  if(!MB && need <= STACK_SMALL && cb200_dual_enabled() && a->dev.num_prims < (1ull << 26) && a->dev.num_nodes < (1ull << 31) && !force_ref64())
    return cb200_launch_intersect_dual(a, d_rays, d_max_dist, d_out, n, stream);
  if(need <= STACK_SMALL)
    return analytic ? launch_intersect_k<MB, false, STACK_SMALL, true >(a, d_rays, d_max_dist, d_out, n, stream, nullptr)
                    : launch_intersect_k<MB, false, STACK_SMALL, false>(a, d_rays, d_max_dist, d_out, n, stream, nullptr);
  return analytic ? launch_intersect_k<MB, false, STACK_BIG, true >(a, d_rays, d_max_dist, d_out, n, stream, nullptr)
                  : launch_intersect_k<MB, false, STACK_BIG, false>(a, d_rays, d_max_dist, d_out, n, stream, nullptr);
}

// the 8-wide kernels serve an accel that has the compressed tree, selected it (cb200_accel_set_traversal, or CB200_WIDE8=1 in
// the environment at build time for A/B measurements) and fits their stacks
bool cb200_use_wide8(const cb200_accel *a)
{
  return a->traversal == CB200_TRAVERSAL_WIDE8 && a->dev.nodes8 && a->dev.depth8 < CB8_STACK && 3*a->depth + 4 <= STACK_SMALL;
}

int cb200_launch_intersect(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, cb_hitrec_t *d_out,
                           uint64_t n, cudaStream_t stream, unsigned long long *d_counters)
{
  if(n == 0) return 0;
  if(cb200_use_wide8(a)) return cb200_launch_intersect8(a, d_rays, d_max_dist, d_out, n, stream, d_counters);
  return a->dev.mb ? launch_intersect_t<true >(a, d_rays, d_max_dist, d_out, n, stream, d_counters)
                   : launch_intersect_t<false>(a, d_rays, d_max_dist, d_out, n, stream, d_counters);
}

template<bool MB, int STACK, bool ANALYTIC>
static int launch_visible_k(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, const uint2 *d_skip, int32_t *d_out,
                            uint64_t n, cudaStream_t stream)
{
  const bool c32 = a->dev.num_prims < (1ull << 26) && a->dev.num_nodes < (1ull << 31) && !force_ref64();
  for(uint64_t first=0; first<n; first+=LAUNCH_MAX_RAYS)
  {
    const uint64_t m = n - first < LAUNCH_MAX_RAYS ? n - first : LAUNCH_MAX_RAYS;
    unsigned int *ticket;
    if(int rc = cb200_get_ticket(stream, &ticket)) return rc;
#define VIS_LAUNCH(SH, C) do { auto k = k_visible<MB, STACK, ANALYTIC, SH, C>; \
    k<<<cb200_trace_grid(m, (const void *)k), TRACE_BLOCK, 0, stream>>>(a->dev, d_rays + first, d_max_dist + first, d_skip ? d_skip + first : nullptr, d_out + first, \
                                                                (uint32_t)m, ticket, cb200_prim_threshold(), cb200_refill_threshold()); } while(0)
    if(d_skip) { if(c32) VIS_LAUNCH(true, true); else VIS_LAUNCH(true, false); }
    else       { if(c32) VIS_LAUNCH(false, true); else VIS_LAUNCH(false, false); }
#undef VIS_LAUNCH
    cb200_count_launch();
    CB_CUDA(cudaGetLastError());
  }
  return 0;
}

template<bool MB>
static int launch_visible_t(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, const uint2 *d_skip, int32_t *d_out,
                            uint64_t n, cudaStream_t stream)
{
  const int need = 3*a->depth + 4;   // the any-hit sweep pushes up to 4 children and pops one per level
  if(need > STACK_BIG + 8) { cb200_set_error("tree deeper than the reference's MAX_TREE_DEPTH"); return CB200_ERR_UNSUPPORTED; }
  const bool analytic = a->scene->any_analytic != 0;
  if(need <= STACK_SMALL)
    return analytic ? launch_visible_k<MB, STACK_SMALL, true >(a, d_rays, d_max_dist, d_skip, d_out, n, stream)
                    : launch_visible_k<MB, STACK_SMALL, false>(a, d_rays, d_max_dist, d_skip, d_out, n, stream);
  return analytic ? launch_visible_k<MB, STACK_BIG + 8, true >(a, d_rays, d_max_dist, d_skip, d_out, n, stream)
                  : launch_visible_k<MB, STACK_BIG + 8, false>(a, d_rays, d_max_dist, d_skip, d_out, n, stream);
}

int cb200_launch_visible(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, int32_t *d_out,
                         uint64_t n, cudaStream_t stream)
{
  if(n == 0) return 0;
  if(!d_max_dist) { cb200_set_error("visible: max_dist is required"); return CB200_ERR_ARG; }
  if(cb200_use_wide8(a)) return cb200_launch_visible8(a, d_rays, d_max_dist, nullptr, d_out, n, stream);
  return a->dev.mb ? launch_visible_t<true>(a, d_rays, d_max_dist, nullptr, d_out, n, stream)
                   : launch_visible_t<false>(a, d_rays, d_max_dist, nullptr, d_out, n, stream);
}

int cb200_launch_shadow(const cb200_accel *a, const cb_ray_t *d_rays, const float *d_max_dist, const uint2 *d_light_prim, int32_t *d_out,
                        uint64_t n, cudaStream_t stream)
{
  if(n == 0) return 0;
  if(!d_max_dist || !d_light_prim) { cb200_set_error("shadow: max_dist and light prims are required"); return CB200_ERR_ARG; }
  if(cb200_use_wide8(a)) return cb200_launch_visible8(a, d_rays, d_max_dist, d_light_prim, d_out, n, stream);
  return a->dev.mb ? launch_visible_t<true>(a, d_rays, d_max_dist, d_light_prim, d_out, n, stream)
                   : launch_visible_t<false>(a, d_rays, d_max_dist, d_light_prim, d_out, n, stream);
}
