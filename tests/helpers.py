"""shared test helpers: golden fixtures -> scenes / record arrays, comparison utilities"""
import importlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
cb = importlib.import_module("corona-13_b200")
S, R = cb.scenes, cb.records

GOLDEN_NAMES = ["static_tris", "motion_quads_analytic", "c10_geometry"]


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        shapes = []
        for i in range(int(z["num_shapes"])):
            shapes.append(S.Shape(z[f"s{i}_primid"], np.ascontiguousarray(z[f"s{i}_vtxidx"]).view(R.VTXIDX).reshape(-1),
                                  np.ascontiguousarray(z[f"s{i}_vtx"]).view(R.VTX).reshape(-1), int(z[f"s{i}_material"]), f"s{i}"))
        self.scene = S.Scene(shapes, name)
        rec = lambda k, dt: np.ascontiguousarray(z[k]).view(dt).reshape(-1)
        self.rays = rec("rays", R.RAY)
        self.hits = rec("hits", R.HITREC)
        self.bounce = rec("bounce", R.RAY)
        self.hits_bounce = rec("hits_bounce", R.HITREC)
        self.shadow = rec("shadow", R.RAY)
        self.shadow_max_dist = z["shadow_max_dist"]
        self.vis = z["vis"]
        self.max_dist = z["max_dist"]
        self.hits_md = rec("hits_md", R.HITREC)
        self.nodes = rec("nodes", R.QBVH_NODE)
        self.primid = z["primid"]
        self.aabb = z["aabb"]
        self.bounds0 = z["bounds0"]
        self.bounds1 = z["bounds1"]


def analytic_mask(hits):
    """hits on sphere / line prims: their u,v go through atan2f/acosf (libm, not bit-reproducible)"""
    vc = R.primid_vcnt(R.hit_prim64(hits))
    return (vc == R.PRIM_SPHERE) | (vc == R.PRIM_LINE)


UV_TOL = 2e-6   # absolute tolerance on u,v of analytic prims (values in [-0.5, 1])


def assert_hits_equal(got, want, what=""):
    """prim id and dist bit-exact everywhere; u,v bit-exact on triangles/quads, within UV_TOL on analytic prims"""
    gp, wp = R.hit_prim64(got), R.hit_prim64(want)
    bad = np.nonzero(gp != wp)[0]
    assert len(bad) == 0, f"{what}: {len(bad)} prim id mismatches, first at ray {bad[:5]}"
    bad = np.nonzero(got["dist"].view("u4") != want["dist"].view("u4"))[0]
    assert len(bad) == 0, f"{what}: {len(bad)} dist mismatches, first at ray {bad[:5]}"
    an = analytic_mask(want)
    hit = wp != R.INVALID_PRIMID
    for f in ("u", "v"):
        tri = hit & ~an
        bad = np.nonzero(got[f].view("u4")[tri] != want[f].view("u4")[tri])[0]
        assert len(bad) == 0, f"{what}: {len(bad)} {f} mismatches on triangles/quads"
        if an.any():
            err = np.abs(got[f][an] - want[f][an]).max()
            assert err <= UV_TOL, f"{what}: analytic {f} off by {err}"


def reachable_nodes(nodes):
    """indices of the nodes reachable from the root (the reference leaves abandoned nodes in its buffer)"""
    seen = []
    st = [0]
    while st:
        i = st.pop()
        seen.append(i)
        for c in nodes["child"][i]:
            if not (int(c) >> 63):
                st.append(int(c))
    return np.array(sorted(seen))


def _slab_margin(orc, prim, ray, limit):
    """width of the ray's slab interval against the primitive's own (time-interpolated) bounding box, relative
    to the distance scale -- a flat or grazed box has a margin of a few ulps and may be culled by rounding"""
    t = np.float32(ray["time"])
    b0, b1 = orc.prim_bounds(prim, False), orc.prim_bounds(prim, True)
    box = (b0 * (np.float32(1) - t) + b1 * t).astype(np.float32)
    with np.errstate(all="ignore"):
        inv = np.float32(1) / ray["dir"].astype(np.float32)
        lo = (box[:3] - ray["pos"]) * inv
        hi = (box[3:] - ray["pos"]) * inv
        tmin = max(np.float32(0), np.nanmax(np.minimum(lo, hi)))
        tmax = min(np.float32(limit), np.nanmin(np.maximum(lo, hi)))
    return float(tmax - tmin) / max(1.0, abs(float(tmax)))


def classify_mismatches(orc, rays, got, want, max_dist=None):
    """mode-B bookkeeping.  The GPU result is asserted bit-identical to the reference ALGORITHM (the oracle) run
    on the same GPU-built tree; what is classified here are the rays where that differs from the answer the
    reference-built tree gives.  Two legitimate causes:
      order : "dist <= hit->dist" lets the last-tested primitive win ties, and a quad's second triangle is only
              tested when its first one is rejected (src/prims.c:654-663; matters for non-planar / moving quads).
              Reproduced by applying the oracle's single-primitive test to the two candidates in both orders.
      cull  : the slab test and the primitive test are different roundings of the same distance; a box that the
              ray grazes or that is flat (axis-aligned quads) can be culled by one ulp when hit->dist is preset
              to the hit distance itself, or shortened to it by a hit already found on the neighbouring triangle of a
              shared edge.  Which boxes enclose a primitive depends on the tree.  Accepted when the
              nearer answer's primitive, tested alone, reproduces that answer and the ray's slab interval against
              the primitive's own box is empty to within 1e-5.
    returns (num_mismatch, num_explained)"""
    gp, wp = R.hit_prim64(got), R.hit_prim64(want)
    idx = np.nonzero((gp != wp) | (got["dist"].view("u4") != want["dist"].view("u4")))[0]
    explained = 0

    def limit(i):
        return R.FLT_MAX if max_dist is None else max_dist[i]

    def run(order, i):
        h = np.zeros(1, R.HIT)
        h["prim"] = 0xFFFFFFFF
        h["dist"] = limit(i)
        for p in order:
            if p != R.INVALID_PRIMID:
                orc.prim_intersect(p, rays[i:i + 1], h)
        return R.hit_prim64(h)[0], h["dist"].view("u4")[0]

    for i in idx:
        a = run([gp[i], wp[i]], i)
        b = run([wp[i], gp[i]], i)
        g = (gp[i], got["dist"].view("u4")[i])
        w = (wp[i], want["dist"].view("u4")[i])
        if (a == g and b == w) or (a == w and b == g):
            explained += 1
            continue
        # culling: one side holds a nearer (or the only) hit that the other tree's boxes rejected
        near, far = (g, w) if got["dist"][i] < want["dist"][i] or wp[i] == R.INVALID_PRIMID else (w, g)
        if near[0] != R.INVALID_PRIMID and run([near[0]], i) == near:
            # the limit the culling side's slab test saw: the preset one, or -- when it had already found the other candidate, a hit
            # an ulp or two farther along a shared edge -- that candidate's distance (the pop test "entry > hit->dist", qbvhmp.c:1440)
            lim_far = limit(i) if far[0] == R.INVALID_PRIMID else min(limit(i), np.uint32(far[1]).view(np.float32))
            if _slab_margin(orc, near[0], rays[i], lim_far) <= 1e-5:
                explained += 1
                continue
        print(f"unexplained difference at ray {i}: ray={rays[i]} gpu={got[i]} ref={want[i]} "
              f"order(gpu,ref)->{a} order(ref,gpu)->{b}")
    return len(idx), explained


MODE_B_LOG = []   # (what, rays, differences) of the GPU-built tree against the reference-built tree's answers (mode B)
TIE_LOG = []   # (what, rays, differences) of every two-mode comparison of this session, printed by conftest at the end


def intersect_modes(acc, orc, rays, max_dist=None, what=""):
    """closest hits in BOTH traversal modes of a GPU-built accel (include/corona_b200.h, CB200_TRAVERSAL_*).
    Returns the EXACT4 answer -- the one the callers hold bit-exact against the oracle on the exported tree.  Where the
    accel also has the 8-wide compressed tree, its answer must be the same bits on every ray except the tree-dependent
    cases (equal-distance ties, grazed boxes), each of which is proven by classify_mismatches; the rate is logged."""
    acc.set_traversal(0)
    got4 = acc.intersect(rays, max_dist)
    if acc.try_traversal(1):
        got8 = acc.intersect(rays, max_dist)
        acc.set_traversal(0)
        nm, nt = classify_mismatches(orc, rays, got8, got4, max_dist)
        TIE_LOG.append((what, len(rays), nm))
        assert nm == nt, f"{what}: {nm - nt} of {nm} WIDE8-vs-EXACT4 differences are neither ties nor grazed boxes"
        same = (R.hit_prim64(got8) == R.hit_prim64(got4)) & (got8["dist"].view("u4") == got4["dist"].view("u4"))
        tri = same & ~analytic_mask(got4)
        assert np.array_equal(got8["u"].view("u4")[tri], got4["u"].view("u4")[tri]) and np.array_equal(got8["v"].view("u4")[tri], got4["v"].view("u4")[tri]), what
    return got4


def visible_modes(acc, rays, max_dist, what=""):
    """any-hit answers in both traversal modes; they must agree except on grazing rays, which are counted and bounded"""
    acc.set_traversal(0)
    v4 = acc.visible(rays, max_dist)
    if acc.try_traversal(1):
        v8 = acc.visible(rays, max_dist)
        acc.set_traversal(0)
        nd = int((v4 != v8).sum())
        TIE_LOG.append((what + " (any-hit)", len(rays), nd))
        assert nd <= max(1, len(rays) // 100000), f"{what}: {nd} of {len(rays)} any-hit answers differ between the two trees"
    return v4


class GoldenImage:
    """scene + reference renders (two seeds per integrator variant) written by tests/golden/make_golden_images.py"""

    def __init__(self, name):
        import ctypes as C
        IO = cb.scene_io
        z = np.load(os.path.join(GOLDEN, "img_" + name + ".npz"))
        self.z = z
        self.name = name
        shapes = []
        mats = z["shape_mats"]
        for i in range(int(z["num_shapes"])):
            shapes.append(S.Shape(z[f"s{i}_primid"], np.ascontiguousarray(z[f"s{i}_vtxidx"]).view(R.VTXIDX).reshape(-1),
                                  np.ascontiguousarray(z[f"s{i}_vtx"]).view(R.VTX).reshape(-1), int(mats[i]), f"s{i}"))
        self.scene = S.Scene(shapes, name)
        self.w, self.h, self.spp = int(z["w"]), int(z["h"]), int(z["spp"])
        self.variants = [str(v) for v in z["variants"]]
        tmp = os.path.join("/tmp", f"golden_cam_{os.getpid()}_{name}.cam")
        open(tmp, "wb").write(z["cam"].tobytes())
        self.camera = IO.read_cam(tmp)
        os.remove(tmp)
        raw = z["materials"].tobytes()
        n = len(raw) // C.sizeof(IO.CMaterial)
        self.materials = IO.MaterialSet()
        arr = (IO.CMaterial * n).from_buffer_copy(raw)
        self.materials.materials = [arr[i] for i in range(n)]
        for i in range(int(z["num_tables"])):
            lmin, step = z[f"tab{i}_meta"]
            self.materials.add_table(lmin, step, z[f"tab{i}_data"])
        if "media" in z.files:
            rawm = z["media"].tobytes()
            nm = len(rawm) // C.sizeof(IO.CMedium)
            arrm = (IO.CMedium * nm).from_buffer_copy(rawm)
            self.materials.media = [arrm[i] for i in range(nm)]
            self.materials.exterior_medium = int(z["exterior_medium"])
        self.sky = str(z["sky"]) if "sky" in z.files else "black"
        self.sky_args = dict(sky=cb.scene_io.sky_kind(self.sky.split()[0]))
        if "env_pixels" in z.files:
            self.sky_args.update(envmap=dict(pixels=z["env_pixels"], mul=float(z["env_mul"]), world=z["env_world"], world_inv=z["env_world_inv"]))
        if "sky_coeff" in z.files:
            self.sky_args.update(sky_coeff=[float(x) for x in z["sky_coeff"]], sky_scale=float(z["sky_scale"]))

    def ref(self, key, seed):
        return self.z[f"{key}_seed{seed}"]

    def write_files(self, directory):
        """the scene in the reference's own file formats: shape<i>.geo, test.nra2, test01.cam -> path of the .nra2"""
        IO = cb.scene_io
        shapes = []
        for i, sh in enumerate(self.scene.shapes):
            sh.write_geo(os.path.join(directory, f"shape{i}.geo"))
            shapes.append((int(self.z["shape_mats"][i]), f"shape{i}"))
        nra2 = os.path.join(directory, "test.nra2")
        if "env_pixels" in self.z.files:   # the sky line names the map relative to the scene
            IO.write_fb(os.path.join(directory, self.sky.split()[1]), self.z["env_pixels"])
        IO.write_nra2(nra2, [str(x) for x in self.z["shader_lines"]], shapes, sky=self.sky)
        open(os.path.join(directory, "test01.cam"), "wb").write(self.z["cam"].tobytes())
        return nra2

    @staticmethod
    def variant_args(key):
        """'ptdl_halton_rec709' -> keyword arguments of lib.Render"""
        IO = cb.scene_io
        f = key.split("_")
        return dict(sampler={"pt": IO.SAMPLER_PT, "ptdl": IO.SAMPLER_PTDL, "ptnee": IO.SAMPLER_PTNEE}[f[0]],
                    pointsampler=IO.POINTS_HALTON if f[1] == "halton" else IO.POINTS_RAND,
                    colour=IO.COLOUR_REC709 if "rec709" in f else IO.COLOUR_XYZ)

    def render(self, lib, acc, key, frame=1, spp=None, **kw):
        """the GPU image of one variant at the golden's resolution and sample count"""
        r = lib.Render(acc, self.camera, self.materials, self.w, self.h, frame=frame, **self.sky_args, **self.variant_args(key), **kw)
        for _ in range(spp or self.spp):
            r.render_pass()
        img, st = r.image(), r.stats()
        r.close()
        return img, st


def image_stats(a, b):
    """(relative RMSE of b against a, per-channel mean ratio b/a); relRMSE = sqrt(mean (a-b)^2) / mean(a)"""
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    rel = np.sqrt(((a64 - b64) ** 2).mean()) / max(a64.mean(), 1e-30)
    return float(rel), b64.mean(axis=(0, 1)) / np.maximum(a64.mean(axis=(0, 1)), 1e-30)
