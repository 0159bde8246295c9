"""Scene description I/O for tests and bench: the reference's text/binary formats and the
translation of its shader list into the flattened material table of include/corona_b200_render.h.

  .nra2   sky line, shader list, shape list            src/shader.c:605-788, src/corona_common.c:30-68
  .cam    camera_t (104 B, "CCAM" v1) / camera_v0_t     include/camera.h:13-35,77-99,153-196
  rgb -> spectrum coefficients                          include/rgb2spec.h:87-128, include/spectrum.h:29-38
  .pfm    framebuffer export                            include/framebuffer.h:142-175
"""
import ctypes as C
import os
import struct

import numpy as np

VIEW_FSTOP = [0.5, 0.7, 1.0, 1.4, 2, 2.8, 4, 5.6, 8, 11, 16, 22, 32, 45, 64, 90, 128]        # src/view.c:71-73
VIEW_EXPOSURE = [60.0, 30.0, 15.0, 8.0, 4.0, 2.0, 1.0, 0.5, 1 / 4, 1 / 8, 1 / 15, 1 / 30, 1 / 60, 1 / 125, 1 / 250, 1 / 500,
                 1 / 1000, 1 / 2000, 1 / 4000, 1 / 8000]                                        # src/view.c:75-79
FULL_FRAME_WIDTH = 0.35                                                                        # src/view.c:70

CB_MAX_MATOPS = 6
OP_COLOR, OP_CHECKERSG = 1, 2
SLOTS = {"d": 0, "s": 1, "e": 2, "v": 3, "g": 4, "r": 5, "t": 6}
BSDF_DIFFUSE, BSDF_DIELECTRIC, BSDF_METAL, BSDF_DIFFDIEL = 0, 1, 2, 3
SAMPLER_PT, SAMPLER_PTDL, SAMPLER_PTNEE = 0, 1, 2
POINTS_RAND, POINTS_HALTON = 0, 1
COLOUR_XYZ, COLOUR_REC709 = 0, 1
SKY_BLACK, SKY_CLOUDY, SKY_CONST, SKY_ENVMAP = 0, 1, 2, 3
SKIES = {"sky_const": SKY_CONST, "black": SKY_BLACK, "cloudy": SKY_CLOUDY, "cloudy_sky": SKY_CLOUDY, "clear_sky": SKY_CLOUDY}   # src/shader.c:626-641
SKY_MODULES = ("daylight",)    # real sky implementations the GPU path does not have: refused, never substituted


def sky_kind(name):
    """line 1 of a .nra2 -> SKY_*, following shader_init (src/shader.c:612-676): prefix matches for the built-ins, `sky_const`
    is the module of that name, and a name that is neither a built-in nor one of the reference's sky modules fails its
    dlopen() there and leaves the default in place, which is the cloudy sky (regression/0090_vstack's `const 1 1 1 2000`)."""
    if name.startswith("black"):
        return SKY_BLACK
    if name.startswith("cloudy") or name.startswith("clear_sky"):
        return SKY_CLOUDY
    if name == "sky_const":
        return SKY_CONST
    if name == "sky_envmap":
        return SKY_ENVMAP
    if name.startswith("daylight") or name in SKY_MODULES:
        raise ValueError(f"sky `{name}' is not supported by the gpu path (no cpu fallback)")
    return SKY_CLOUDY



class CCamera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("pos_t1", C.c_float * 3), ("orient", C.c_float * 4), ("orient_t1", C.c_float * 4),
                ("focus", C.c_float), ("film_width", C.c_float), ("film_height", C.c_float),
                ("aperture_value", C.c_int32), ("exposure_value", C.c_int32), ("focal_length", C.c_float), ("iso", C.c_float)]


class CMatOp(C.Structure):
    _fields_ = [("op", C.c_int32), ("slot", C.c_int32), ("coeff", C.c_float * 3), ("mul", C.c_float),
                ("roughness", C.c_float), ("table", C.c_int32)]


class CMaterial(C.Structure):
    _fields_ = [("num_ops", C.c_int32), ("bsdf", C.c_int32), ("param", C.c_float * 4), ("table", C.c_int32), ("medium", C.c_int32),
                ("ops", CMatOp * CB_MAX_MATOPS)]


class CMedium(C.Structure):
    _fields_ = [("mu_t_coeff", C.c_float * 3), ("mu_t_mul", C.c_float), ("g", C.c_float), ("has_albedo", C.c_int32),
                ("albedo_coeff", C.c_float * 3), ("albedo_mul", C.c_float)]


class CTable(C.Structure):
    _fields_ = [("lambda_min", C.c_float), ("lambda_step", C.c_float), ("num_lambda", C.c_int32), ("rows", C.c_int32),
                ("data", C.c_void_p)]


class CRenderDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("camera", CCamera),
                ("materials", C.c_void_p), ("num_materials", C.c_int32),
                ("tables", C.c_void_p), ("num_tables", C.c_int32),
                ("sampler", C.c_int32), ("pointsampler", C.c_int32), ("colour_camera", C.c_int32), ("max_path_len", C.c_int32),
                ("frame", C.c_uint64), ("rank", C.c_uint32), ("world", C.c_uint32), ("batch_paths", C.c_uint64),
                ("sky", C.c_int32), ("sky_coeff", C.c_float * 3), ("sky_scale", C.c_float), ("exterior_medium", C.c_int32),
                ("media", C.c_void_p), ("num_media", C.c_int32), ("pad", C.c_int32), ("envmap", C.c_void_p)]


class CEnvmap(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("pixels", C.c_void_p), ("mul", C.c_float),
                ("world", C.c_float * 9), ("world_inv", C.c_float * 9)]


class CRenderStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays_closest", C.c_uint64), ("rays_shadow", C.c_uint64), ("splats", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("ms", C.c_double * 5), ("trav_closest", C.c_uint64 * 4),
                ("trav_shadow", C.c_uint64 * 4)]


BSDF_QUERY = np.dtype([("wi", "<f4", 3), ("wo", "<f4", 3), ("lambda_", "<f4"), ("rand", "<f4", 3), ("rd", "<f4"), ("rs", "<f4"),
                       ("rg", "<f4"), ("roughness", "<f4"), ("flip", "<i4")])
BSDF_RESULT = np.dtype([("s_wo", "<f4", 3), ("s_weight", "<f4"), ("s_pdf", "<f4"), ("s_mode", "<u4"), ("f", "<f4"), ("f_mode", "<u4"),
                        ("pdf", "<f4")])
MEDIUM_QUERY = np.dtype([("wi", "<f4", 3), ("wo", "<f4", 3), ("lambda_", "<f4"), ("rand", "<f4", 3), ("dist", "<f4")])
MEDIUM_RESULT = np.dtype([("mu_t", "<f4"), ("mu_s", "<f4"), ("free_dist", "<f4"), ("free_pdf", "<f4"), ("transmittance", "<f4"),
                          ("vol_pdf", "<f4"), ("s_wo", "<f4", 3), ("s_weight", "<f4"), ("s_pdf", "<f4"), ("s_mode", "<u4"), ("f", "<f4"),
                          ("f_mode", "<u4"), ("pdf", "<f4")])


# ----------------------------------------------------------------------------------------------- camera
def quat_from_frame(a, b, n):
    """quaternion (w,x,y,z) rotating the unit axes onto the orthonormal right-handed frame (a,b,n)"""
    m = np.stack([a, b, n], axis=1).astype(np.float64)
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(m)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + m[i, i] - m[j, j] - m[k, k]) * 2
        q = [0.0, 0.0, 0.0, 0.0]
        q[0] = (m[k, j] - m[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (m[j, i] + m[i, j]) / s
        q[1 + k] = (m[k, i] + m[i, k]) / s
    return np.asarray(q, np.float32)


class Camera:
    def __init__(self, pos, lookat, up=(0, 0, 1), focus=None, aperture_value=6, exposure_value=13, focal_length=0.5,
                 iso=100.0, crop_factor=1.0, pos_t1=None):
        self.pos = np.asarray(pos, np.float32)
        self.pos_t1 = np.asarray(pos_t1 if pos_t1 is not None else pos, np.float32)
        n = np.asarray(lookat, np.float64) - np.asarray(pos, np.float64)
        dist = np.linalg.norm(n)
        n /= dist
        a = np.cross(np.asarray(up, np.float64), n)
        a /= np.linalg.norm(a)
        b = np.cross(n, a)
        self.orient = quat_from_frame(a, b, n)
        self.orient_t1 = self.orient.copy()
        self.focus = float(focus if focus is not None else dist)
        self.aperture_value, self.exposure_value = int(aperture_value), int(exposure_value)
        self.focal_length, self.iso, self.crop_factor = float(focal_length), float(iso), float(crop_factor)

    def write(self, path):
        """camera_t, 104 bytes"""
        with open(path, "wb") as f:
            f.write(b"CCAM" + struct.pack("<i", 1))
            f.write(self.pos.tobytes() + self.pos_t1.tobytes() + self.orient.tobytes() + self.orient_t1.tobytes())
            f.write(struct.pack("<fffff f ii ff", 1.0, 0.0, self.focus, 0.0, 0.0, self.crop_factor,
                                self.aperture_value, self.exposure_value, self.focal_length, self.iso))

    def film(self, width, height):
        """view_cam_read recomputes the film back from the aspect ratio (src/view.c:939-948)"""
        if width > height:
            fw = np.float32(FULL_FRAME_WIDTH) / np.float32(self.crop_factor)
            return float(fw), float(np.float32(height) / np.float32(width) * fw)
        fh = np.float32(FULL_FRAME_WIDTH) / np.float32(self.crop_factor)
        return float(np.float32(width) / np.float32(height) * fh), float(fh)

    def cstruct(self, width, height):
        c = CCamera()
        c.pos[:] = self.pos.tolist()
        c.pos_t1[:] = self.pos_t1.tolist()
        c.orient[:] = self.orient.tolist()
        c.orient_t1[:] = self.orient_t1.tolist()
        c.focus = self.focus
        c.film_width, c.film_height = self.film(width, height)
        c.aperture_value, c.exposure_value = self.aperture_value, self.exposure_value
        c.focal_length, c.iso = self.focal_length, self.iso
        return c


def read_cam(path):
    d = open(path, "rb").read()
    cam = Camera.__new__(Camera)
    if len(d) == 152:   # camera_v0_t; crop_factor is NOT taken over by camera_read (include/camera.h:168-181): stays 1.0
        f = struct.unpack("<i3f4ff7if4f3ff4ffffffiffi", d)
        cam.pos = np.float32(f[1:4])
        cam.orient = np.float32(f[4:8])
        cam.iso = f[16]
        cam.orient_t1 = np.float32(f[17:21])
        cam.pos_t1 = np.float32(f[21:24])
        cam.focus = f[29]
        cam.crop_factor = 1.0
        cam.aperture_value = f[34]
        cam.focal_length = f[35]
        cam.exposure_value = f[37]
    elif len(d) == 104:
        assert d[:4] == b"CCAM"
        f = struct.unpack("<i3f3f4f4fffffffiiff", d[4:])
        cam.pos, cam.pos_t1 = np.float32(f[1:4]), np.float32(f[4:7])
        cam.orient, cam.orient_t1 = np.float32(f[7:11]), np.float32(f[11:15])
        cam.focus = f[17]
        cam.crop_factor = f[20]
        cam.aperture_value, cam.exposure_value = f[21], f[22]
        cam.focal_length, cam.iso = f[23], f[24]
    else:
        raise ValueError(f"{path}: unknown camera file size {len(d)}")
    if cam.exposure_value < 0 or cam.exposure_value > len(VIEW_EXPOSURE):
        cam.exposure_value = 13
    if cam.iso < 1 or cam.iso > 409600:
        cam.iso = 100.0
    return cam


# ----------------------------------------------------------------------------------------------- rgb -> spectrum
class Rgb2Spec:
    """numpy restatement of rgb2spec_init / rgb2spec_fetch (include/rgb2spec.h:40-128) over the coefficient file the
    reference's own tool writes (data/ergb2spec.coeff)"""

    def __init__(self, path):
        d = open(path, "rb").read()
        assert d[:4] == b"SPEC"
        self.res = struct.unpack("<I", d[4:8])[0]
        r = self.res
        self.scale = np.frombuffer(d, "<f4", r, 8)
        self.data = np.frombuffer(d, "<f4", r * r * r * 3 * 3, 8 + 4 * r)

    def fetch(self, rgb):
        rgb = np.asarray(rgb, np.float32)
        res = self.res
        i = 0
        for j in (1, 2):
            if rgb[j] >= rgb[i]:
                i = j
        z = rgb[i]
        if not z > 0:   # black: 0*inf = NaN coefficients in the reference too; CLAMP(NaN, 0, 1) = 0 downstream (corona_common.h:168-170)
            return np.full(3, np.nan, np.float32)
        scale = np.float32(res - 1) / z
        x = rgb[(i + 1) % 3] * scale
        y = rgb[(i + 2) % 3] * scale
        xi, yi = min(int(x), res - 2), min(int(y), res - 2)
        left, last, size = 0, res - 2, res - 2   # rgb2spec_find_interval
        while size > 0:
            half = size >> 1
            mid = left + half + 1
            if self.scale[mid] < z:
                left = mid
                size -= half + 1
            else:
                size = half
        zi = min(left, last)
        off = (((i * res + zi) * res + yi) * res + xi) * 3
        dx, dy, dz = 3, 3 * res, 3 * res * res
        x1 = np.float32(x - xi)
        x0 = np.float32(1) - x1
        y1 = np.float32(y - yi)
        y0 = np.float32(1) - y1
        z1 = (z - self.scale[zi]) / (self.scale[zi + 1] - self.scale[zi])
        z0 = np.float32(1) - z1
        D = self.data
        out = np.zeros(3, np.float32)
        for j in range(3):
            o = off + j
            out[j] = ((D[o] * x0 + D[o + dx] * x1) * y0 + (D[o + dy] * x0 + D[o + dy + dx] * x1) * y1) * z0 + \
                     ((D[o + dz] * x0 + D[o + dz + dx] * x1) * y0 + (D[o + dz + dy] * x0 + D[o + dz + dy + dx] * x1) * y1) * z1
        return out

    def rgb_to_coeff(self, rgb):
        """spectrum_rgb_to_coeff (include/spectrum.h:29-38): returns (mul, coeff[3])"""
        rgb = np.asarray(rgb, np.float32)
        mul = np.float32(max(rgb))
        if mul == 0.0 or mul < 1.0:
            mul = np.float32(1.0)
        return float(mul), self.fetch(rgb / mul)


def coeff_path(root):
    return os.path.join(root, "oracle", "_ref", "data", "ergb2spec.coeff")


# ----------------------------------------------------------------------------------------------- materials
class MaterialSet:
    """flattened material table + the lookup tables it references; keeps the ctypes arrays alive"""

    def __init__(self):
        self.materials = []   # CMaterial
        self.tables = []      # (lambda_min, step, np.ndarray rows x num)
        self.text = []        # the .nra2 shader lines this set corresponds to
        self.media = []       # CMedium; CMaterial.medium / exterior_medium are 1 + index
        self.exterior_medium = 0

    def add_table(self, lambda_min, step, data):
        self.tables.append((float(lambda_min), float(step), np.ascontiguousarray(data, np.float32)))
        return len(self.tables) - 1

    def carrays(self):
        mats = (CMaterial * max(1, len(self.materials)))(*self.materials)
        tabs = (CTable * max(1, len(self.tables)))()
        for i, (lm, st, d) in enumerate(self.tables):
            tabs[i].lambda_min, tabs[i].lambda_step = lm, st
            tabs[i].rows, tabs[i].num_lambda = d.shape
            tabs[i].data = d.ctypes.data
        return mats, tabs

    def cmedia(self):
        return (CMedium * max(1, len(self.media)))(*self.media)


def parse_nra2(path, rgb2spec, checker_table=None, metal_tables=None):
    """shader list of a .nra2 -> (MaterialSet indexed by shader number, [(shader index, geo path)], sky).
    Mirrors shader_init (src/shader.c:605-788): every line is one shader; `mult n pre... host` is flattened
    (src/shaders/mult.c:90-128,154-167); a chain that ends in `medium_rgb` becomes a CMedium, `interior s m` attaches it to
    surface s (src/shaders/interior.c), `exterior id` makes it the camera's medium (src/shader.c:699-716).
    Shader kinds outside the hot path raise (no fallback)."""
    lines = [l.split("#")[0].strip() for l in open(path).read().split("\n")]
    sky = lines[0].split()[0]
    n = int(lines[1].split()[0])
    raw = [lines[2 + i].split() for i in range(n)]
    ms = MaterialSet()
    ms.text = [" ".join(r) for r in raw]
    chk = None
    medium_of = {}    # shader index -> 1 + index into ms.media

    def medium_index(i):
        """shader i as a homogeneous medium: `medium_rgb` itself or a mult of `color v` steps ending in one"""
        if i in medium_of:
            return medium_of[i]
        r = raw[i]
        albedo = None
        host = i
        if r[0] == "mult":
            k = int(r[1])
            pre = [int(x) for x in r[2:2 + k]]
            host = int(r[2 + k])
            pre = [i + q if q < 0 else q for q in pre]
            host = i + host if host < 0 else host
            for q in pre:
                if raw[q][0] != "color" or raw[q][1] != "v":
                    raise ValueError("only `color v' steps may precede a medium")
                albedo = rgb2spec.rgb_to_coeff([float(raw[q][2]), float(raw[q][3]), float(raw[q][4])])
        if raw[host][0] != "medium_rgb":
            raise ValueError(f"shader {i} is not a homogeneous medium")
        h = raw[host]
        with np.errstate(divide="ignore"):
            mu_t = (np.float32(1.0) / np.array([float(h[1]), float(h[2]), float(h[3])], np.float32)).astype(np.float32)   # medium_rgb.c:127-128
        mul, co = rgb2spec.rgb_to_coeff(mu_t)
        m = CMedium()
        m.mu_t_coeff[:] = [float(x) for x in co]
        m.mu_t_mul = float(mul)
        m.g = float(h[4])
        if albedo is not None:
            m.has_albedo = 1
            m.albedo_mul = float(albedo[0])
            m.albedo_coeff[:] = [float(x) for x in albedo[1]]
        ms.media.append(m)
        medium_of[i] = len(ms.media)
        return medium_of[i]

    def op_of(i):
        """prepare() steps shader i contributes, and its bsdf if it has one"""
        r = raw[i]
        kind = r[0]
        if kind == "color":
            op = CMatOp()
            op.op, op.slot = OP_COLOR, SLOTS[r[1]]
            mul, co = rgb2spec.rgb_to_coeff([float(r[2]), float(r[3]), float(r[4])])
            op.coeff[:] = co.tolist()
            op.mul = mul
            op.roughness = float(r[5]) if len(r) > 5 else 1.0
            return [op], None
        if kind == "colorcheckersg":
            nonlocal chk
            if checker_table is None:
                raise ValueError("colorcheckersg needs the reflectance table fixture")
            if chk is None:
                chk = ms.add_table(380.0, 10.0, checker_table)
            op = CMatOp()
            op.op, op.slot, op.table = OP_CHECKERSG, SLOTS[r[1]], chk
            op.roughness = float(r[2]) if len(r) > 2 else 1.0
            return [op], None
        if kind == "diffuse":
            return [], (BSDF_DIFFUSE, [0, 0, 0, 0], -1)
        if kind == "dielectric":
            return [], (BSDF_DIELECTRIC, [float(r[1]), float(r[2]) if len(r) > 2 else 50.0, 0, 0], -1)
        if kind == "diffdiel":
            return [], (BSDF_DIFFDIEL, [float(r[1]), float(r[2]) if len(r) > 2 else 50.0, 0, 0], -1)
        if kind == "metal":
            if metal_tables is None or r[1].lower() not in metal_tables:
                raise ValueError(f"metal `{r[1]}' needs its ior table fixture")
            t = ms.add_table(360.0, 5.0, metal_tables[r[1].lower()])
            return [], (BSDF_METAL, [0, 0, 0, 0], t)
        if kind == "mult":
            k = int(r[1])
            pre = [int(x) for x in r[2:2 + k]]
            host = int(r[2 + k])
            pre = [i + p if p < 0 else p for p in pre]
            host = i + host if host < 0 else host
            ops = []
            for p in pre:
                o, _ = op_of(p)
                ops += o
            ho, hb = op_of(host)
            return ops + ho, hb
        raise ValueError(f"shader `{kind}' is outside the pt/ptdl hot path (SURVEY 2.1)")

    for i in range(n):
        m = CMaterial()
        medium = 0
        try:
            if raw[i][0] == "exterior":
                eid = int(raw[i][1])
                if len(raw[i]) > 2 and int(raw[i][2]):
                    raise ValueError("volume lights are outside the hot path")
                ms.exterior_medium = medium_index(eid) if eid >= 0 else 0
                raise ValueError("not a surface")
            if raw[i][0] == "interior":
                surf, inner = int(raw[i][1]), int(raw[i][2])
                surf = i + surf if surf < 0 else surf
                inner = i + inner if inner < 0 else inner
                medium = medium_index(inner)
                ops, bsdf = op_of(surf)
            else:
                ops, bsdf = op_of(i)
        except ValueError:
            m.num_ops, m.bsdf = -1, -1     # not usable as a shape material; rejected by cb200_render_create if referenced
            ms.materials.append(m)
            continue
        if bsdf is None:
            bsdf = (BSDF_DIFFUSE, [0, 0, 0, 0], -1)   # a bare prepare-only shader on a shape falls back to the diffuse callbacks (shader.c:761-787)
        m.num_ops = len(ops)
        for k, o in enumerate(ops[:CB_MAX_MATOPS]):
            m.ops[k] = o
        m.bsdf = bsdf[0]
        m.param[:] = bsdf[1]
        m.table = bsdf[2]
        m.medium = medium
        ms.materials.append(m)
    ns = int(lines[2 + n].split()[0])
    shapes = []
    for i in range(ns):
        r = lines[3 + n + i].split()
        shapes.append((int(r[0]), r[1]))
    return ms, shapes, sky


FB_MAGIC = 1936686951


def write_fb(path, pixels):
    """framebuffer file as fb_map reads it (include/framebuffer.h:26-35,52-84): 32-byte header + float32 texels"""
    px = np.ascontiguousarray(pixels, np.float32)
    h, w, c = px.shape
    with open(path, "wb") as f:
        f.write(np.array([FB_MAGIC, w, h], "<u8").tobytes() + np.array([c, 0], "<u2").tobytes() + np.float32(1.0).tobytes())
        f.write(px.tobytes())


def read_fb(path):
    raw = open(path, "rb").read()
    magic, w, h = np.frombuffer(raw[:24], "<u8")
    c = int(np.frombuffer(raw[24:26], "<u2")[0])
    if magic != FB_MAGIC or len(raw) != 32 + int(w) * int(h) * c * 4:
        raise ValueError(f"{path}: not a framebuffer file")
    return np.frombuffer(raw[32:], "<f4").reshape(int(h), int(w), c).copy()


def _mat3_rotate(axis, angle_deg):
    """mat3_rotate (include/matrix3.inc:111-134), float arithmetic"""
    f = np.float32
    a = f(angle_deg) / f(180) * np.pi
    s, c = f(np.sin(f(a))), f(np.cos(f(a)))
    x, y, z = (f(v) for v in axis)
    one = f(1)
    return np.array([[x*x + (one - x*x)*c, x*y*(one - c) - z*s, x*z*(one - c) + y*s],
                     [y*x*(one - c) + z*s, y*y + (one - y*y)*c, y*z*(one - c) - x*s],
                     [z*x*(one - c) - y*s, z*y*(one - c) + x*s, z*z + (one - z*z)*c]], np.float32)


def envmap_params(args, searchpath="."):
    """sky_envmap's init (src/shaders/sky_envmap.c:272-311): '<file.fb> [brightness] [rot_x rot_y rot_z]' (degrees) ->
    dict(pixels, mul, world, world_inv)"""
    f = args.split()
    name = f[0]
    vals = [float(x) for x in f[1:5]] + [1.0, 0.0, 0.0, 0.0][len(f[1:5]):]
    path = name if os.path.exists(name) else os.path.join(searchpath, name)
    px = read_fb(path)
    if px.shape[2] != 4 or px.shape[1] != 2 * px.shape[0]:
        raise ValueError("environment map has to be w = 2 h with 4 channels (rgb2spec coefficients + scale)")
    rx, ry, rz = _mat3_rotate((1, 0, 0), vals[1]), _mat3_rotate((0, 1, 0), vals[2]), _mat3_rotate((0, 0, 1), vals[3])
    world = (rx.astype(np.float32) @ (ry @ rz).astype(np.float32)).astype(np.float32)     # mat3_mul(ry, rz, tmp); mat3_mul(rx, tmp, world)
    return dict(pixels=px, mul=vals[0], world=world, world_inv=np.linalg.inv(world.astype(np.float64)).astype(np.float32))


def sky_const_params(rgb2spec, args):
    """sky_const's init (src/shaders/sky_const.c:89-101): 'r g b [scale]' -> (coeff[3], scale*mul)"""
    f = [float(x) for x in args.split()[:4]]
    col = (f + [1.0, 1.0, 1.0])[:3] if len(f) < 3 else f[:3]
    scale = f[3] if len(f) > 3 else 1.0
    mul, co = rgb2spec.rgb_to_coeff(col)
    return [float(x) for x in co], float(np.float32(scale) * np.float32(mul))


def write_nra2(path, shader_lines, shapes, sky="black"):
    with open(path, "w") as f:
        f.write(sky + "\n%d\n" % len(shader_lines))
        for i, l in enumerate(shader_lines):
            f.write(l + " # %d\n" % i)   # the trailing comment matters: shader init()s end with fscanf("%*[^\n]\n"), which
                                          # swallows the NEXT line when nothing is left on the current one (color.c:47)
        f.write("%d\n" % len(shapes))
        for mat, geo in shapes:
            f.write("%d %s\n" % (mat, geo))


# ----------------------------------------------------------------------------------------------- pfm
def read_pfm(path):
    """PFM as fb_export writes it (include/framebuffer.h:142-175): 'PF', w h, scale, header padded with '0' to 16 bytes"""
    d = open(path, "rb").read()
    parts = d.split(b"\n", 3)
    assert parts[0] == b"PF"
    w, h = [int(x) for x in parts[1].split()]
    body = d[len(d) - w * h * 12:]
    return np.frombuffer(body, "<f4").reshape(h, w, 3).copy()


def pfmdiff_rmse(a, b):
    """tools/img/pfmdiff.c:75-86: sqrt(sum over pixels and channels of d^2 / (W*H))"""
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.sqrt((d * d).sum() / (a.shape[0] * a.shape[1])))
