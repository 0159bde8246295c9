"""Procedural scenes in the reference's geometry layout (.geo, SURVEY Appendix A;
reader/writer follow include/prims.h:26-35 and src/prims.c:804-823) plus ray-set generators.

The reference's regression geometry is partly unavailable offline, so the
benchmark and most tests run on seed-fixed procedural meshes: an fBm height-field terrain, an
icosphere soup on top for depth complexity, an emissive quad above, optional analytic
sphere / cylinder / cone primitives and motion-blurred variants.
"""
import os
import struct
import numpy as np

from .records import (VTX, VTXIDX, RAY, CShape, primid_make, make_rays, INVALID_PRIMID,
                      PRIM_SPHERE, PRIM_LINE, PRIM_TRI, PRIM_QUAD)

GEO_MAGIC = 0xC01337
GEO_VERSION = 2


# ----------------------------------------------------------------------------- normals / uv codecs
def encode_normal(n):
    """vectorised restatement of geo_encode_normal (include/geo.h:46-74): 2x(sign+15 bit) octahedral"""
    n = np.asarray(n, dtype=np.float32).reshape(-1, 3)
    inv = np.float32(1.0) / (np.abs(n[:, 0]) + np.abs(n[:, 1]) + np.abs(n[:, 2]))
    sx = np.where(n[:, 0] < 0, np.float32(-1), np.float32(1))
    sy = np.where(n[:, 1] < 0, np.float32(-1), np.float32(1))
    e0 = np.where(n[:, 2] < 0, (np.float32(1) - np.abs(n[:, 1] * inv)) * sx, n[:, 0] * inv).astype(np.float32)
    e1 = np.where(n[:, 2] < 0, (np.float32(1) - np.abs(n[:, 0] * inv)) * sy, n[:, 1] * inv).astype(np.float32)
    i0 = ((np.abs(e0) + np.float32(2)) / np.float32(2)).astype(np.float32).view(np.uint32)
    i1 = ((np.abs(e1) + np.float32(2)) / np.float32(2)).astype(np.float32).view(np.uint32)
    p0 = ((e0.view(np.uint32) & np.uint32(0x80000000)) >> np.uint32(16)) | ((i0 & np.uint32(0x7FFFFF)) >> np.uint32(8))
    p1 = ((e1.view(np.uint32) & np.uint32(0x80000000)) >> np.uint32(16)) | ((i1 & np.uint32(0x7FFFFF)) >> np.uint32(8))
    p0 = np.where((p0 & np.uint32(0x7FFF)) == 0, np.uint32(0), p0)
    p1 = np.where((p1 & np.uint32(0x7FFF)) == 0, np.uint32(0), p1)
    return (p0 & np.uint32(0xFFFF)) | (p1 << np.uint32(16))


def encode_uv(u, v):
    """two IEEE halfs in one word (include/geo.h:76-82); numpy rounds to nearest where the reference
    truncates, which only matters for data we generate ourselves"""
    h = np.stack([np.asarray(u, np.float32).astype(np.float16), np.asarray(v, np.float32).astype(np.float16)], -1)
    return np.ascontiguousarray(h).view(np.uint32).reshape(-1)


# ----------------------------------------------------------------------------- shapes
class Shape:
    """one .geo worth of geometry.  vtx is interleaved (open, close) when mb is set."""

    def __init__(self, primid, vtxidx, vtx, material=0, name="shape"):
        self.primid = np.ascontiguousarray(primid, dtype=np.uint64)
        self.vtxidx = np.ascontiguousarray(vtxidx, dtype=VTXIDX)
        self.vtx = np.ascontiguousarray(vtx, dtype=VTX)
        self.material = int(material)
        self.name = name

    @property
    def num_prims(self):
        return len(self.primid)

    def cshape(self):
        s = CShape()
        s.primid = self.primid.ctypes.data
        s.num_prims = len(self.primid)
        s.vtxidx = self.vtxidx.ctypes.data
        s.num_vtxidx = len(self.vtxidx)
        s.vtx = self.vtx.ctypes.data
        s.num_vtx = len(self.vtx)
        s.material = self.material
        return s

    def write_geo(self, path):
        hdr = 32
        vio = hdr + 8 * len(self.primid)
        vo = (vio + 8 * len(self.vtxidx) + 15) & ~15
        with open(path, "wb") as f:
            f.write(struct.pack("<iiQQQ", GEO_MAGIC, GEO_VERSION, len(self.primid), vio, vo))
            f.write(self.primid.tobytes())
            f.write(self.vtxidx.tobytes())
            f.write(b"\0" * (vo - vio - 8 * len(self.vtxidx)))
            f.write(self.vtx.tobytes())


def read_geo(path, material=0):
    d = np.fromfile(path, dtype=np.uint8)
    magic, ver, n, vio, vo = struct.unpack("<iiQQQ", d[:32].tobytes())
    if magic != GEO_MAGIC or ver != GEO_VERSION:
        raise ValueError(f"{path}: not a version-{GEO_VERSION} .geo file")
    primid = d[32:32 + 8 * n].view("<u8")
    vtxidx = d[vio:vo - ((vo - vio) % 8)].view(VTXIDX)
    nv = (len(d) - vo) // 16
    vtx = d[vo:vo + 16 * nv].view(VTX)
    return Shape(primid.copy(), vtxidx.copy(), vtx.copy(), material, os.path.basename(path))


class Scene:
    def __init__(self, shapes, name="scene"):
        self.shapes = list(shapes)
        self.name = name

    @property
    def num_prims(self):
        return sum(s.num_prims for s in self.shapes)

    def cshapes(self):
        arr = (CShape * len(self.shapes))()
        for i, s in enumerate(self.shapes):
            arr[i] = s.cshape()
        return arr

    def bounds(self):
        lo = np.min([s.vtx["v"].min(axis=0) for s in self.shapes], axis=0)
        hi = np.max([s.vtx["v"].max(axis=0) for s in self.shapes], axis=0)
        return lo, hi


def _mk_vtx(pos, normals=None, close=None):
    """build the VTX array; interleave open/close positions when close is given"""
    pos = np.asarray(pos, np.float32)
    nrm = encode_normal(normals) if normals is not None else np.zeros(len(pos), np.uint32)
    if close is None:
        v = np.zeros(len(pos), VTX)
        v["v"] = pos
        v["n"] = nrm
        return v
    v = np.zeros(2 * len(pos), VTX)
    v["v"][0::2] = pos
    v["v"][1::2] = np.asarray(close, np.float32)
    v["n"][0::2] = nrm
    v["n"][1::2] = nrm
    return v


def _vertex_normals(pos, tris):
    fn = np.cross(pos[tris[:, 1]] - pos[tris[:, 0]], pos[tris[:, 2]] - pos[tris[:, 0]])
    n = np.zeros((len(pos), 3), np.float64)
    for k in range(3):
        for c in range(3):
            n[:, c] += np.bincount(tris[:, k], weights=fn[:, c], minlength=len(pos))
    ln = np.linalg.norm(n, axis=1)
    n[ln == 0] = (0, 0, 1)
    ln[ln == 0] = 1
    return (n / ln[:, None]).astype(np.float32)


def mesh_shape(pos, faces, material=0, motion=None, name="mesh", uv=None):
    """indexed triangle (faces.shape[1]==3) or quad (==4) mesh -> Shape.  motion: translation
    vector or callable pos->pos giving the shutter-close positions (sets primid.mb)."""
    pos = np.asarray(pos, np.float32)
    faces = np.asarray(faces, np.int64)
    vc = faces.shape[1]
    n = len(faces)
    tri = faces[:, :3] if vc == 3 else np.concatenate([faces[:, [0, 1, 2]], faces[:, [0, 2, 3]]])
    nrm = _vertex_normals(pos.astype(np.float64), tri)
    close = None
    if motion is not None:
        close = motion(pos) if callable(motion) else pos + np.asarray(motion, np.float32)
    vtx = _mk_vtx(pos, nrm, close)
    vtxidx = np.zeros(n * vc, VTXIDX)
    vtxidx["v"] = faces.reshape(-1)
    if uv is not None:
        vtxidx["uv"] = encode_uv(uv[faces.reshape(-1), 0], uv[faces.reshape(-1), 1])
    primid = primid_make(0, 0, np.arange(n, dtype=np.uint64) * np.uint64(vc), 0 if motion is None else 1,
                         PRIM_TRI if vc == 3 else PRIM_QUAD)
    return Shape(primid, vtxidx, vtx, material, name)


# ----------------------------------------------------------------------------- generators
def _value_noise(res, freq, rng):
    lat = rng.random((freq + 2, freq + 2)).astype(np.float32)
    x = np.linspace(0, freq, res, dtype=np.float32)
    i = np.minimum(x.astype(np.int64), freq - 1)
    f = x - i
    f = f * f * (3 - 2 * f)
    a = lat[i][:, i]
    b = lat[i + 1][:, i]
    c = lat[i][:, i + 1]
    d = lat[i + 1][:, i + 1]
    fx = f[:, None]
    fy = f[None, :]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def terrain(num_tris, seed=1, size=20.0, height=5.0, quads=False, material=0, motion=None):
    """fBm height field on a g x g grid in a size x size x height box"""
    rng = np.random.default_rng(seed)
    g = max(1, int(round(np.sqrt(num_tris / (1 if quads else 2)))))
    res = g + 1
    h = np.zeros((res, res), np.float32)
    amp, freq = 1.0, 2
    tot = 0.0
    while freq <= max(2, min(res, 512)):
        h += amp * _value_noise(res, freq, rng)
        tot += amp
        amp *= 0.5
        freq *= 2
    h = h / tot * height
    xs = np.linspace(-size / 2, size / 2, res, dtype=np.float32)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    pos = np.stack([X, Y, h], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    v00 = (i * res + j).reshape(-1)
    v10 = v00 + res
    v01 = v00 + 1
    v11 = v10 + 1
    if quads:
        faces = np.stack([v00, v10, v11, v01], -1)
    else:
        faces = np.concatenate([np.stack([v00, v10, v11], -1), np.stack([v00, v11, v01], -1)])
    return mesh_shape(pos, faces, material, motion, "terrain")


_ICO_V = None
_ICO_F = None


def _icosahedron():
    global _ICO_V, _ICO_F
    if _ICO_V is None:
        t = (1 + 5 ** 0.5) / 2
        v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                      [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
        _ICO_V = (v / np.linalg.norm(v, axis=1)[:, None]).astype(np.float32)
        _ICO_F = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                           [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5],
                           [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], np.int64)
    return _ICO_V, _ICO_F


def soup(num_tris, seed=2, size=20.0, zmin=1.0, zmax=6.0, rmin=0.01, rmax=0.5, material=0, motion=None):
    """icosphere soup: uniform random centres, log-uniform radii"""
    rng = np.random.default_rng(seed)
    iv, itri = _icosahedron()
    m = max(1, num_tris // 20)
    c = np.stack([rng.uniform(-size / 2, size / 2, m), rng.uniform(-size / 2, size / 2, m),
                  rng.uniform(zmin, zmax, m)], -1).astype(np.float32)
    r = np.exp(rng.uniform(np.log(rmin), np.log(rmax), m)).astype(np.float32)
    pos = (c[:, None, :] + r[:, None, None] * iv[None, :, :]).reshape(-1, 3)
    faces = (itri[None, :, :] + 12 * np.arange(m)[:, None, None]).reshape(-1, 3)
    return mesh_shape(pos, faces, material, motion, "soup")


def quad_light(center=(0.0, 0.0, 9.0), half=2.0, material=1):
    cx, cy, cz = center
    pos = np.array([[cx - half, cy - half, cz], [cx - half, cy + half, cz],
                    [cx + half, cy + half, cz], [cx + half, cy - half, cz]], np.float32)
    return mesh_shape(pos, np.array([[0, 1, 2, 3]]), material, None, "light")  # normal points down (-z)


def analytic_shape(kind, p0, r0, p1=None, r1=None, material=0, motion=None):
    """one analytic primitive: 'sphere' (vcnt 1) or 'line' = cylinder / truncated cone (vcnt 2).
    the radius lives in the vertex' n word as a float (sphere.h:9-13, line.h:9-15)"""
    pts = [p0] if kind == "sphere" else [p0, p1]
    rad = [r0] if kind == "sphere" else [r0, r1]
    pos = np.asarray(pts, np.float32)
    close = None if motion is None else pos + np.asarray(motion, np.float32)
    vtx = _mk_vtx(pos, None, close)
    rbits = np.asarray(rad, np.float32).view(np.uint32)
    if motion is None:
        vtx["n"] = rbits
    else:
        vtx["n"][0::2] = rbits
        vtx["n"][1::2] = rbits
    vtxidx = np.zeros(len(pts), VTXIDX)
    vtxidx["v"] = np.arange(len(pts))
    primid = primid_make(0, 0, np.zeros(1, np.uint64), 0 if motion is None else 1,
                         PRIM_SPHERE if kind == "sphere" else PRIM_LINE)
    return Shape(primid, vtxidx, vtx, material, kind)


def synthetic_scene(num_tris, seed=1, motion=False, quads=False, analytic=False, soup_frac=0.5):
    """the benchmark family: terrain + soup + light (+ analytic prims), seed-fixed"""
    nt = int(num_tris * (1 - soup_frac))
    ns = num_tris - nt
    shapes = [terrain(max(2, nt), seed, quads=quads, material=0)]
    if ns >= 20:
        shapes.append(soup(ns, seed + 1, material=0, motion=(0.15, 0.05, -0.1) if motion else None))
    shapes.append(quad_light(material=1))
    if analytic:
        shapes.append(analytic_shape("sphere", (1.0, 2.0, 6.5), 0.8, material=2,
                                     motion=(0.2, 0.0, 0.1) if motion else None))
        shapes.append(analytic_shape("line", (-3.0, -2.0, 5.0), 0.5, (-3.0, -2.0, 7.5), 0.5, material=2))
        shapes.append(analytic_shape("line", (3.0, -3.0, 5.0), 0.9, (3.5, -3.0, 7.0), 0.2, material=2))
    return Scene(shapes, f"synthetic_{num_tris}")


def c10_like_scene(geo_dir=None):
    """the 0010_pt scene geometry (regression/0010_pt/test.nra2:17-23): the six in-tree .geo files when
    a copy is available (oracle/_ref/scenes, made by oracle/Makefile), else a procedural look-alike
    with the same primitive mix (4096+3+6 quads, 1 sphere, 1 cylinder, 1 cone)."""
    names = [("plane", 2), ("emitter", 5), ("cylinder_cap", 2), ("sphere", 10), ("cylinder", 10), ("cone", 10)]
    if geo_dir and all(os.path.exists(os.path.join(geo_dir, n + ".geo")) for n, _ in names):
        return Scene([read_geo(os.path.join(geo_dir, n + ".geo"), m) for n, m in names], "0010_pt")
    g = 64
    xs = np.linspace(-8, 8, g + 1, dtype=np.float32)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    pos = np.stack([X, Y, np.zeros_like(X)], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    v00 = (i * (g + 1) + j).reshape(-1)
    plane = mesh_shape(pos, np.stack([v00, v00 + g + 1, v00 + g + 2, v00 + 1], -1), 2, None, "plane")
    lights = []
    for k in range(3):
        lights.append(quad_light((-4.0 + 4.0 * k, 0.0, 6.0), 1.0, 5))
    emitter = merge_shapes(lights, 5, "emitter")
    caps = merge_shapes([quad_light((-2.0 + k, -2.0, 1.0 + 0.1 * k), 0.3, 2) for k in range(6)], 2, "cylinder_cap")
    return Scene([plane, emitter, caps,
                  analytic_shape("sphere", (0.0, 1.0, 1.0), 1.0, material=10),
                  analytic_shape("line", (2.5, 0.0, 0.0), 0.6, (2.5, 0.0, 2.0), 0.6, material=10),
                  analytic_shape("line", (-2.5, 0.0, 0.0), 0.8, (-2.5, 0.0, 2.0), 0.1, material=10)], "0010_pt_like")


def merge_shapes(shapes, material, name):
    """concatenate shapes with identical mb-ness into one"""
    primid, vtxidx, vtx = [], [], []
    vi_off = 0
    v_off = 0
    for s in shapes:
        p = s.primid.copy()
        vi = (p >> np.uint64(32)) & np.uint64(0x0FFFFFFF)
        p = (p & ~(np.uint64(0x0FFFFFFF) << np.uint64(32))) | ((vi + np.uint64(vi_off)) << np.uint64(32))
        mb = int((p[0] >> np.uint64(60)) & np.uint64(1)) if len(p) else 0
        x = s.vtxidx.copy()
        x["v"] += v_off
        primid.append(p)
        vtxidx.append(x)
        vtx.append(s.vtx)
        vi_off += len(s.vtxidx)
        v_off += len(s.vtx) // (mb + 1)
    return Shape(np.concatenate(primid), np.concatenate(vtxidx), np.concatenate(vtx), material, name)


# ----------------------------------------------------------------------------- ray sets
def camera_rays(n, scene, seed=3, eye=None, time_max=0.0, frame=None):
    """pinhole-ish primary rays looking at the scene (kernel-only ray set 1).
    frame=(w,h): one jittered ray per pixel in 8x4-pixel tile order (n must be w*h) -- the order a wavefront
    camera kernel emits them in; otherwise uniformly random film positions (the reference's per-path pixel
    sampling, thinlens.c:117-118)"""
    rng = np.random.default_rng(seed)
    lo, hi = scene.bounds()
    ctr = (lo + hi) / 2
    ext = float(np.max(hi - lo))
    eye = np.asarray(eye if eye is not None else ctr + np.array([0.9 * ext, 0.7 * ext, 0.6 * ext]), np.float32)
    fwd = ctr - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0, 0, 1.0])
    right /= np.linalg.norm(right)
    up = np.cross(right, fwd)
    if frame is not None:
        w, h = frame
        assert n == w * h and w % 8 == 0 and h % 4 == 0
        i = np.arange(n, dtype=np.int64)
        tile, inner = i // 32, i % 32
        tx, ty = tile % (w // 8), tile // (w // 8)
        px = (tx * 8 + inner % 8).astype(np.float32)
        py = (ty * 4 + inner // 8).astype(np.float32)
        jit = rng.random((n, 2)).astype(np.float32)
        uv = np.stack([(px + jit[:, 0]) / np.float32(w), (py + jit[:, 1]) / np.float32(h)], -1).astype(np.float32) - 0.5
    else:
        uv = rng.random((n, 2)).astype(np.float32) - 0.5
    d = fwd[None, :] + 0.9 * uv[:, :1] * right[None, :] + 0.6 * uv[:, 1:] * up[None, :]
    d = (d / np.linalg.norm(d, axis=1)[:, None]).astype(np.float32)
    t = (rng.random(n) * time_max).astype(np.float32)
    return make_rays(np.broadcast_to(eye, (n, 3)), d, t)


def random_rays(n, scene, seed=4, time_max=0.0):
    """incoherent rays: origins uniform in the (slightly grown) scene box, uniform directions"""
    rng = np.random.default_rng(seed)
    lo, hi = scene.bounds()
    pad = 0.05 * (hi - lo)
    o = (lo - pad + rng.random((n, 3)) * (hi - lo + 2 * pad)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d = (d / np.linalg.norm(d, axis=1)[:, None]).astype(np.float32)
    t = (rng.random(n) * time_max).astype(np.float32)
    return make_rays(o, d, t)


def bounce_rays(rays, hits, seed=5):
    """diffuse bounce set: from the hit points of `rays`, uniform-hemisphere-ish directions, origin
    offset and ignore prim as prims_offset_ray does (src/prims.c:374-388)"""
    rng = np.random.default_rng(seed)
    from .records import hit_prim64
    prim = hit_prim64(hits)
    ok = prim != INVALID_PRIMID
    r = rays[ok]
    h = hits[ok]
    x = r["pos"] + h["dist"][:, None] * r["dir"]
    d = rng.normal(size=(len(r), 3))
    d = (d / np.linalg.norm(d, axis=1)[:, None]).astype(np.float32)
    flip = np.sum(d * r["dir"], axis=1) > 0
    d[flip] = -d[flip]
    eps = (np.maximum(0.5, np.abs(x).max(axis=1)) * np.float32(1e-4)).astype(np.float32)
    out = make_rays((x + eps[:, None] * d).astype(np.float32), d, r["time"], 0.0, prim[ok])
    return out


def shadow_rays(rays, hits, light_pos, seed=6):
    """next-event set: from hit points toward random points on the light; returns (rays, max_dist)"""
    rng = np.random.default_rng(seed)
    from .records import hit_prim64
    prim = hit_prim64(hits)
    ok = prim != INVALID_PRIMID
    r = rays[ok]
    h = hits[ok]
    x = r["pos"] + h["dist"][:, None] * r["dir"]
    lp = np.asarray(light_pos, np.float32)[None, :] + (rng.random((len(r), 3)).astype(np.float32) - 0.5) * np.float32([4, 4, 0])
    d = lp - x
    dist = np.linalg.norm(d, axis=1).astype(np.float32)
    d = (d / dist[:, None]).astype(np.float32)
    eps = (np.maximum(0.5, np.abs(x).max(axis=1)) * np.float32(1e-4)).astype(np.float32)
    out = make_rays((x + eps[:, None] * d).astype(np.float32), d, r["time"], 0.0, prim[ok])
    return out, (dist - 2 * eps).astype(np.float32)
