"""ctypes binding of the product's C ABI (include/corona_b200.h) for tests and bench.py.

Loading fails loudly when libcorona_b200.so has not been built; every compute call raises
Cb200Error when the library reports an error (no CUDA device, bad arguments, ...).  There is no
CPU path behind any of this.
"""
import ctypes as C
import os

import numpy as np

from .records import RAY, HITREC, QBVH_NODE, CShape

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcorona_b200.so")

# every symbol include/corona_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "cb200_version", "cb200_last_error", "cb200_device_count", "cb200_set_device", "cb200_sm_count",
    "cb200_malloc", "cb200_free", "cb200_malloc_host", "cb200_free_host", "cb200_memcpy_h2d",
    "cb200_memcpy_d2h", "cb200_stream_sync",
    "cb200_scene_create", "cb200_scene_destroy", "cb200_scene_num_prims",
    "cb200_accel_build", "cb200_accel_import_qbvh", "cb200_accel_destroy", "cb200_accel_num_nodes",
    "cb200_accel_depth", "cb200_accel_aabb", "cb200_accel_export_qbvh", "cb200_accel_layout",
    "cb200_accel_intersect_n", "cb200_accel_visible_n", "cb200_accel_closest_n", "cb200_accel_intersect_dev",
    "cb200_accel_visible_dev", "cb200_accel_intersect_counted", "cb200_launch_count",
    "cb200_accel_set_traversal", "cb200_accel_traversal",
]


class Cb200Error(RuntimeError):
    pass


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Cb200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, u64, i32, f32p = C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_float)
    L.cb200_version.restype = C.c_char_p
    L.cb200_last_error.restype = C.c_char_p
    L.cb200_device_count.restype = i32
    L.cb200_set_device.argtypes = [i32]
    L.cb200_sm_count.restype = i32
    L.cb200_malloc.restype = vp
    L.cb200_malloc.argtypes = [C.c_size_t]
    L.cb200_free.argtypes = [vp]
    L.cb200_malloc_host.restype = vp
    L.cb200_malloc_host.argtypes = [C.c_size_t]
    L.cb200_free_host.argtypes = [vp]
    L.cb200_memcpy_h2d.argtypes = [vp, vp, C.c_size_t, vp]
    L.cb200_memcpy_d2h.argtypes = [vp, vp, C.c_size_t, vp]
    L.cb200_stream_sync.argtypes = [vp]
    L.cb200_scene_create.restype = vp
    L.cb200_scene_create.argtypes = [vp, i32]
    L.cb200_scene_destroy.argtypes = [vp]
    L.cb200_scene_num_prims.restype = u64
    L.cb200_scene_num_prims.argtypes = [vp]
    L.cb200_accel_build.restype = vp
    L.cb200_accel_build.argtypes = [vp, vp, vp]
    L.cb200_accel_import_qbvh.restype = vp
    L.cb200_accel_import_qbvh.argtypes = [vp, vp, u64, vp, vp]
    L.cb200_accel_destroy.argtypes = [vp]
    L.cb200_accel_num_nodes.restype = u64
    L.cb200_accel_num_nodes.argtypes = [vp]
    L.cb200_accel_depth.argtypes = [vp]
    L.cb200_accel_aabb.argtypes = [vp, vp]
    L.cb200_accel_export_qbvh.argtypes = [vp, vp, u64, vp]
    L.cb200_accel_layout.argtypes = [vp, vp, vp]
    L.cb200_accel_intersect_n.argtypes = [vp, vp, vp, vp, u64]
    L.cb200_accel_visible_n.argtypes = [vp, vp, vp, vp, u64]
    L.cb200_accel_closest_n.argtypes = [vp, vp, vp, vp, u64]
    L.cb200_accel_intersect_dev.argtypes = [vp, vp, vp, vp, u64, vp]
    L.cb200_accel_visible_dev.argtypes = [vp, vp, vp, vp, u64, vp]
    L.cb200_accel_intersect_counted.argtypes = [vp, vp, vp, vp, u64, vp]
    L.cb200_launch_count.restype = u64
    L.cb200_accel_set_traversal.argtypes = [vp, i32]
    L.cb200_accel_traversal.argtypes = [vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _check(rc, what):
    if rc != 0:
        raise Cb200Error(f"{what} failed ({rc}): {load().cb200_last_error().decode()}")


def _nonnull(p, what):
    if not p:
        raise Cb200Error(f"{what} failed: {load().cb200_last_error().decode()}")
    return p


def device_count():
    return load().cb200_device_count()


def set_device(i):
    _check(load().cb200_set_device(i), "cb200_set_device")


def launch_count():
    return int(load().cb200_launch_count())


class Accel:
    """scene upload + acceleration structure on the current device"""

    def __init__(self, scene):
        self.L = load()
        self.host_scene = scene
        self._cshapes = scene.cshapes()
        self.s = _nonnull(self.L.cb200_scene_create(C.cast(self._cshapes, C.c_void_p), len(scene.shapes)),
                          "cb200_scene_create")
        self.a = None
        self.primid = None

    def num_prims(self):
        return int(self.L.cb200_scene_num_prims(self.s))

    def build(self, ghost_aabb=None):
        self._drop_accel()
        n = self.num_prims()
        self.primid = np.zeros(n, np.uint64)
        g = np.ascontiguousarray(ghost_aabb, np.float32) if ghost_aabb is not None else None
        self.a = _nonnull(self.L.cb200_accel_build(self.s, _ptr(g), _ptr(self.primid) if n else None), "cb200_accel_build")
        return self

    def import_qbvh(self, nodes, primid, aabb=None):
        self._drop_accel()
        nodes = np.ascontiguousarray(nodes, dtype=QBVH_NODE)
        primid = np.ascontiguousarray(primid, np.uint64)
        ab = np.ascontiguousarray(aabb, np.float32) if aabb is not None else None
        self.a = _nonnull(self.L.cb200_accel_import_qbvh(self.s, _ptr(nodes), len(nodes), _ptr(primid), _ptr(ab)),
                          "cb200_accel_import_qbvh")
        self.primid = primid.copy()
        return self

    def export_qbvh(self):
        n = int(self.L.cb200_accel_num_nodes(self.a))
        nodes = np.zeros(n, QBVH_NODE)
        primid = np.zeros(self.num_prims(), np.uint64)
        _check(self.L.cb200_accel_export_qbvh(self.a, _ptr(nodes), n, _ptr(primid) if len(primid) else None),
               "cb200_accel_export_qbvh")
        return nodes, primid

    def num_nodes(self):
        return int(self.L.cb200_accel_num_nodes(self.a))

    def set_traversal(self, mode):
        """EXACT4 (0) = the reference's order on the 4-wide tree, WIDE8 (1) = the 8-wide compressed tree"""
        _check(self.L.cb200_accel_set_traversal(self.a, int(mode)), "cb200_accel_set_traversal")
        return self

    def try_traversal(self, mode):
        """select the mode if this accel supports it (WIDE8 needs the compressed tree: static scene, GPU-built)"""
        return self.L.cb200_accel_set_traversal(self.a, int(mode)) == 0

    def traversal(self):
        return int(self.L.cb200_accel_traversal(self.a))

    def depth(self):
        return int(self.L.cb200_accel_depth(self.a))

    def aabb(self):
        out = np.zeros(6, np.float32)
        _check(self.L.cb200_accel_aabb(self.a, _ptr(out)), "cb200_accel_aabb")
        return out

    def layout(self):
        nb, pb = C.c_uint32(0), C.c_uint32(0)
        _check(self.L.cb200_accel_layout(self.a, C.byref(nb), C.byref(pb)), "cb200_accel_layout")
        return nb.value, pb.value

    # host-buffer calls (copies inside)
    def intersect(self, rays, max_dist=None):
        rays = np.ascontiguousarray(rays, dtype=RAY)
        md = np.ascontiguousarray(max_dist, np.float32) if max_dist is not None else None
        out = np.zeros(len(rays), HITREC)
        _check(self.L.cb200_accel_intersect_n(self.a, _ptr(rays), _ptr(md), _ptr(out), len(rays)), "cb200_accel_intersect_n")
        return out

    def visible(self, rays, max_dist):
        rays = np.ascontiguousarray(rays, dtype=RAY)
        md = np.ascontiguousarray(max_dist, np.float32)
        out = np.zeros(len(rays), np.int32)
        _check(self.L.cb200_accel_visible_n(self.a, _ptr(rays), _ptr(md), _ptr(out), len(rays)), "cb200_accel_visible_n")
        return out

    def closest(self, rays, hits, centre):
        """accel_closest for a batch: returns (rays with updated min_dist, hits {prim,u,v,dist})"""
        rays = np.array(rays, dtype=RAY, copy=True)
        io = np.array(hits, dtype=HITREC, copy=True)
        c = np.ascontiguousarray(centre, np.float32)
        _check(self.L.cb200_accel_closest_n(self.a, _ptr(rays), _ptr(io), _ptr(c), len(rays)), "cb200_accel_closest_n")
        return rays, io

    # device-pointer calls (ints from torch .data_ptr(), stream from torch.cuda.current_stream().cuda_stream)
    def intersect_dev(self, d_rays, d_max_dist, d_out, n, stream=0):
        _check(self.L.cb200_accel_intersect_dev(self.a, d_rays, d_max_dist or None, d_out, n, stream or None),
               "cb200_accel_intersect_dev")

    def visible_dev(self, d_rays, d_max_dist, d_out, n, stream=0):
        _check(self.L.cb200_accel_visible_dev(self.a, d_rays, d_max_dist, d_out, n, stream or None),
               "cb200_accel_visible_dev")

    def intersect_counted(self, d_rays, d_max_dist, d_out, n):
        cnt = np.zeros(4, np.uint64)
        _check(self.L.cb200_accel_intersect_counted(self.a, d_rays, d_max_dist or None, d_out, n, _ptr(cnt)),
               "cb200_accel_intersect_counted")
        return cnt

    def _drop_accel(self):
        if self.a:
            self.L.cb200_accel_destroy(self.a)
            self.a = None

    def close(self):
        self._drop_accel()
        if self.s:
            self.L.cb200_scene_destroy(self.s)
            self.s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------ integrator
def _load_render():
    L = load()
    if getattr(L, "_render_ready", False):
        return L
    vp, u64 = C.c_void_p, C.c_uint64
    L.cb200_render_create.restype = vp
    L.cb200_render_create.argtypes = [vp, vp]
    L.cb200_render_destroy.argtypes = [vp]
    L.cb200_render_pass.argtypes = [vp, u64, u64, vp]
    L.cb200_render_pass_stream.argtypes = [vp, u64, u64, vp]
    L.cb200_render_flush.argtypes = [vp, vp]
    L.cb200_render_clear.argtypes = [vp, vp]
    L.cb200_render_instrument.argtypes = [vp, C.c_int, C.c_int]
    L.cb200_render_fb_device.restype = vp
    L.cb200_render_fb_device.argtypes = [vp]
    L.cb200_render_download.argtypes = [vp, vp, vp]
    L.cb200_render_set_framebuffer.argtypes = [vp, vp]
    L.cb200_render_set_dbor.argtypes = [vp, C.c_int32]
    L.cb200_render_num_dbors.argtypes = [vp]
    L.cb200_render_dbor_device.restype = vp
    L.cb200_render_dbor_device.argtypes = [vp, C.c_int32]
    L.cb200_render_download_dbor.argtypes = [vp, C.c_int32, vp, vp]
    L.cb200_render_snapshot.argtypes = [vp, vp, vp]
    L.cb200_render_snapshot_async.argtypes = [vp, vp, vp]
    L.cb200_render_snapshot_wait.argtypes = [vp]
    L.cb200_render_stats.argtypes = [vp, vp]
    L.cb200_render_path_stats.argtypes = [vp, C.c_int]
    L.cb200_render_set_accumulation.argtypes = [vp, C.c_int]
    L.cb200_render_accumulation.argtypes = [vp]
    L.cb200_render_get_path_stats.argtypes = [vp, vp, vp]
    L.cb200_render_point.argtypes = [vp, vp, vp, vp, u64]
    L.cb200_render_camera_rays.argtypes = [vp, u64, u64, vp, vp]
    L.cb200_render_offset_ray.argtypes = [vp, vp, vp, vp, u64]
    L.cb200_render_nee_records.argtypes = [vp, u64, u64, vp, vp]
    L.cb200_render_bounce_records.argtypes = [vp, u64, u64, C.c_float, vp, vp]
    L.cb200_render_emission_records.argtypes = [vp, u64, u64, C.c_float, C.c_int32, vp, vp]
    L.cb200_render_bsdf.argtypes = [vp, C.c_int32, vp, vp, u64]
    L.cb200_render_medium.argtypes = [vp, C.c_int32, vp, vp, u64]
    L._render_ready = True
    return L


RENDER_SYMBOLS = ["cb200_render_create", "cb200_render_destroy", "cb200_render_pass", "cb200_render_pass_stream", "cb200_render_flush", "cb200_render_clear", "cb200_render_instrument",
                  "cb200_render_fb_device", "cb200_render_set_framebuffer", "cb200_render_download", "cb200_render_snapshot", "cb200_render_snapshot_async", "cb200_render_snapshot_wait", "cb200_render_stats", "cb200_render_point",
                  "cb200_render_camera_rays", "cb200_render_offset_ray", "cb200_render_nee_records", "cb200_render_bounce_records", "cb200_render_emission_records", "cb200_render_bsdf", "cb200_render_medium",
                  "cb200_render_set_dbor", "cb200_render_num_dbors", "cb200_render_dbor_device", "cb200_render_download_dbor",
                  "cb200_render_path_stats", "cb200_render_get_path_stats", "cb200_render_set_accumulation", "cb200_render_accumulation",
                  "cb200_comm_unique_id", "cb200_reducer_create", "cb200_reducer_destroy", "cb200_reducer_begin", "cb200_reducer_end",
                  "cb200_reducer_finish", "cb200_reducer_clear"]


class Render:
    """wavefront pt/ptdl integrator on top of an Accel (include/corona_b200_render.h)"""

    def __init__(self, accel, camera, materials, width, height, sampler=1, pointsampler=0, colour=0, max_path_len=32,
                 frame=0, rank=0, world=1, batch_paths=0, sky=0, sky_coeff=(0.0, 0.0, 0.0), sky_scale=1.0, envmap=None):
        from . import scene_io as sio
        self.L = _load_render()
        self.accel = accel
        self.width, self.height = (width + 31) // 32 * 32, (height + 31) // 32 * 32   # view.c:295-296
        self.camera = camera
        self._mats, self._tabs = materials.carrays()
        self._materials = materials
        d = sio.CRenderDesc()
        d.width, d.height = self.width, self.height
        d.camera = camera.cstruct(self.width, self.height)
        d.materials = C.cast(self._mats, C.c_void_p)
        d.num_materials = len(materials.materials)
        d.tables = C.cast(self._tabs, C.c_void_p)
        d.num_tables = len(materials.tables)
        d.sampler, d.pointsampler, d.colour_camera, d.max_path_len = sampler, pointsampler, colour, max_path_len
        d.frame, d.rank, d.world, d.batch_paths = frame, rank, world, batch_paths
        d.sky = sky
        d.sky_coeff[:] = [float(x) for x in sky_coeff]
        d.sky_scale = float(sky_scale)
        if envmap is not None:   # dict(pixels (h, w, 4) float32, mul, world, world_inv) as scene_io.envmap_params returns it
            self._env_px = np.ascontiguousarray(envmap["pixels"], np.float32)
            e = sio.CEnvmap()
            e.height, e.width = self._env_px.shape[:2]
            e.pixels = self._env_px.ctypes.data
            e.mul = float(envmap["mul"])
            e.world[:] = [float(x) for x in np.asarray(envmap["world"], np.float32).reshape(-1)]
            e.world_inv[:] = [float(x) for x in np.asarray(envmap["world_inv"], np.float32).reshape(-1)]
            self._env = e
            d.envmap = C.cast(C.pointer(e), C.c_void_p)
        self._media = materials.cmedia()
        d.media = C.cast(self._media, C.c_void_p)
        d.num_media = len(materials.media)
        d.exterior_medium = materials.exterior_medium
        self.desc = d
        self.r = _nonnull(self.L.cb200_render_create(accel.a, C.byref(d)), "cb200_render_create")
        self.spp = 0
        self.next_index = 0

    def flush(self, stream=0):
        _check(self.L.cb200_render_flush(self.r, stream or None), "cb200_render_flush")

    def render_pass(self, first=None, count=None, stream=0, streaming=False):
        """one progression = width*height path indices (src/view.c:636-638); streaming=True leaves stragglers in the pool"""
        count = self.width * self.height if count is None else count
        first = self.next_index if first is None else first
        f = self.L.cb200_render_pass_stream if streaming else self.L.cb200_render_pass
        _check(f(self.r, first, count, stream or None), "cb200_render_pass")
        self.next_index = first + count
        self.spp += count / (self.width * self.height)

    def clear(self):
        _check(self.L.cb200_render_clear(self.r, None), "cb200_render_clear")
        self.spp = 0
        self.next_index = 0

    def fb_device(self):
        return self.L.cb200_render_fb_device(self.r)

    def set_framebuffer(self, device_ptr):
        _check(self.L.cb200_render_set_framebuffer(self.r, device_ptr or None), "cb200_render_set_framebuffer")

    def set_dbor(self, levels):
        """`--dbor n` (src/view.c:291): n > 1 switches the outlier rejection cascade of view_splat_col on"""
        _check(self.L.cb200_render_set_dbor(self.r, int(levels)), "cb200_render_set_dbor")

    def num_dbors(self):
        return int(self.L.cb200_render_num_dbors(self.r))

    def dbor_device(self, level):
        return self.L.cb200_render_dbor_device(self.r, int(level))

    def dbor_images(self, spp=None):
        """the cascade buffers scaled by the framebuffer's gain (view.c:659-664), (levels, H, W, 3)"""
        spp = self.spp if spp is None else spp
        out = np.zeros((self.num_dbors(), self.height, self.width, 3), np.float32)
        for l in range(out.shape[0]):
            _check(self.L.cb200_render_download_dbor(self.r, l, _ptr(out[l]), None), "cb200_render_download_dbor")
        return out * np.float32(self.camera.iso / (100.0 * max(spp, 1e-9)))

    def framebuffer(self):
        fb = np.zeros((self.height, self.width, 3), np.float32)
        _check(self.L.cb200_render_download(self.r, _ptr(fb), None), "cb200_render_download")
        return fb

    def snapshot(self):
        """the accumulation buffer as it stands (no flush): what a progressive display shows between streamed progressions"""
        fb = np.zeros((self.height, self.width, 3), np.float32)
        _check(self.L.cb200_render_snapshot(self.r, _ptr(fb), None), "cb200_render_snapshot")
        return fb

    def snapshot_async(self, fb, stream=0):
        """the same into the caller's (ideally pinned) float32 buffer without waiting; snapshot_wait() / the next synchronous
        download waits for it"""
        assert fb.dtype == np.float32 and fb.size == self.height * self.width * 3 and fb.flags["C_CONTIGUOUS"]
        _check(self.L.cb200_render_snapshot_async(self.r, _ptr(fb), stream or None), "cb200_render_snapshot_async")

    def snapshot_wait(self):
        _check(self.L.cb200_render_snapshot_wait(self.r), "cb200_render_snapshot_wait")

    def image(self, spp=None):
        """fb * gain, gain = iso / (100 * spp) (src/view.c:656)"""
        spp = self.spp if spp is None else spp
        return self.framebuffer() * np.float32(self.camera.iso / (100.0 * max(spp, 1e-9)))

    def stats(self):
        from . import scene_io as sio
        s = sio.CRenderStats()
        _check(self.L.cb200_render_stats(self.r, C.byref(s)), "cb200_render_stats")
        out = {}
        for k, _ in s._fields_:
            v = getattr(s, k)
            out[k] = list(v) if hasattr(v, "__len__") else int(v)
        return out

    def set_accumulation(self, mode):
        """0 = fp32 atomics per tap (the reference's scheme), 1 = atomic-free per-tile accumulation"""
        _check(self.L.cb200_render_set_accumulation(self.r, int(mode)), "cb200_render_set_accumulation")

    def accumulation(self):
        return int(self.L.cb200_render_accumulation(self.r))

    def path_stats(self, enable=True):
        _check(self.L.cb200_render_path_stats(self.r, int(enable)), "cb200_render_path_stats")

    def get_path_stats(self):
        """(energy[33], count[33]) per path length, the reference's view->stat_enery / stat_cnt"""
        e, c = np.zeros(33, np.float64), np.zeros(33, np.uint64)
        _check(self.L.cb200_render_get_path_stats(self.r, _ptr(e), _ptr(c)), "cb200_render_get_path_stats")
        return e, c

    def instrument(self, timing=True, counters=False):
        _check(self.L.cb200_render_instrument(self.r, int(timing), int(counters)), "cb200_render_instrument")

    def points(self, index, dim):
        index = np.ascontiguousarray(index, np.uint64)
        dim = np.ascontiguousarray(dim, np.int32)
        out = np.zeros(len(index), np.float32)
        _check(self.L.cb200_render_point(self.r, _ptr(index), _ptr(dim), _ptr(out), len(index)), "cb200_render_point")
        return out

    def bsdf(self, material, queries):
        """battle-test protocol (tools/battle-test.c:57-236) for one material: scene_io.BSDF_QUERY[] -> BSDF_RESULT[]"""
        from . import scene_io as sio
        q = np.ascontiguousarray(queries, sio.BSDF_QUERY)
        out = np.zeros(len(q), sio.BSDF_RESULT)
        _check(self.L.cb200_render_bsdf(self.r, material, _ptr(q), _ptr(out), len(q)), "cb200_render_bsdf")
        return out

    def medium(self, medium, queries):
        """free flight, edge terms and phase function of medium `medium` (index into the material set's media):
        scene_io.MEDIUM_QUERY[] -> MEDIUM_RESULT[]"""
        from . import scene_io as sio
        q = np.ascontiguousarray(queries, sio.MEDIUM_QUERY)
        out = np.zeros(len(q), sio.MEDIUM_RESULT)
        _check(self.L.cb200_render_medium(self.r, medium, _ptr(q), _ptr(out), len(q)), "cb200_render_medium")
        return out

    def camera_rays(self, first, n):
        rays = np.zeros(n, RAY)
        aux = np.zeros((n, 4), np.float32)
        _check(self.L.cb200_render_camera_rays(self.r, first, n, _ptr(rays), _ptr(aux)), "cb200_render_camera_rays")
        return rays, aux

    def nee_records(self, first, n):
        """next-event samples at the first hit vertex of path indices [first, first + n): (m, 16) float32 rows, see the header"""
        out = np.zeros((n, 16), np.float32)
        m = C.c_uint64(0)
        _check(self.L.cb200_render_nee_records(self.r, first, n, _ptr(out), C.byref(m)), "cb200_render_nee_records")
        return out[:m.value]

    def bounce_records(self, first, n, scrambling=0.5):
        """the paths that go on behind their first hit vertex: (m, 16) float32 rows, see the header"""
        out = np.zeros((n, 16), np.float32)
        m = C.c_uint64(0)
        _check(self.L.cb200_render_bounce_records(self.r, first, n, scrambling, _ptr(out), C.byref(m)), "cb200_render_bounce_records")
        return out[:m.value]

    def emission_records(self, first, n, wave, scrambling=0.5):
        """emission found by extension at the first (wave 1) / second (wave 2) vertex: (m, 8) float32 rows, see the header"""
        out = np.zeros((n, 8), np.float32)
        m = C.c_uint64(0)
        _check(self.L.cb200_render_emission_records(self.r, first, n, scrambling, wave, _ptr(out), C.byref(m)), "cb200_render_emission_records")
        return out[:m.value]

    def offset_ray(self, x, direction):
        """prims_offset_ray for arrays of surface points / directions, (n, 3) each"""
        x = np.ascontiguousarray(x, np.float32)
        d = np.ascontiguousarray(direction, np.float32)
        out = np.zeros_like(x)
        _check(self.L.cb200_render_offset_ray(self.r, _ptr(x), _ptr(d), _ptr(out), len(x)), "cb200_render_offset_ray")
        return out

    def close(self):
        if self.r:
            self.L.cb200_render_destroy(self.r)
            self.r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
