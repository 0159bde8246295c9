"""ctypes bindings for the CPU checkers -- TEST INFRASTRUCTURE.

  Oracle : oracle/liboracle.so            our plain-C restatement (oracle.c)
  Ref    : oracle/_ref/libcorona_ref.so   the unmodified reference compiled in place (ref_glue.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
_pkg = importlib.import_module("corona-13_b200.records")
RAY, HITREC, HIT, QBVH_NODE, CShape = _pkg.RAY, _pkg.HITREC, _pkg.HIT, _pkg.QBVH_NODE, _pkg.CShape

u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def build(ref=True):
    """compile liboracle.so and, when the reference tree is present, oracle/_ref/*.so"""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.exists("/root/reference/src/accel.d/qbvhmp.c"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


class Oracle:
    """scene + accel held by the restatement"""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(HERE, "liboracle.so")
            if not os.path.exists(path):
                build(ref=False)
            L = C.CDLL(path)
            L.orc_scene_new.restype = C.c_void_p
            L.orc_scene_new.argtypes = [C.c_void_p, C.c_int]
            L.orc_scene_free.argtypes = [C.c_void_p]
            L.orc_scene_num_prims.restype = C.c_uint64
            L.orc_scene_num_prims.argtypes = [C.c_void_p]
            L.orc_scene_primid.restype = u64p
            L.orc_scene_primid.argtypes = [C.c_void_p]
            L.orc_prim_bounds.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
            L.orc_prim_intersect.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
            L.orc_prim_visible.restype = C.c_int
            L.orc_prim_visible.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_float]
            L.orc_accel_build.restype = C.c_void_p
            L.orc_accel_build.argtypes = [C.c_void_p]
            L.orc_accel_import.restype = C.c_void_p
            L.orc_accel_import.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
            L.orc_accel_free.argtypes = [C.c_void_p]
            L.orc_accel_num_nodes.restype = C.c_uint64
            L.orc_accel_num_nodes.argtypes = [C.c_void_p]
            L.orc_accel_nodes.restype = C.c_void_p
            L.orc_accel_nodes.argtypes = [C.c_void_p]
            L.orc_accel_aabb.restype = f32p
            L.orc_accel_aabb.argtypes = [C.c_void_p]
            L.orc_intersect_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
            L.orc_visible_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
            L.orc_closest_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
            L.orc_accel_check.restype = C.c_int
            L.orc_accel_check.argtypes = [C.c_void_p, C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, scene):
        self.L = self.lib()
        self.scene = scene                      # keeps the numpy arrays alive
        self._cshapes = scene.cshapes()
        self.s = self.L.orc_scene_new(C.cast(self._cshapes, C.c_void_p), len(scene.shapes))
        self.a = None

    def num_prims(self):
        return int(self.L.orc_scene_num_prims(self.s))

    def primid(self):
        n = self.num_prims()
        return np.ctypeslib.as_array(self.L.orc_scene_primid(self.s), shape=(n,)).copy() if n else np.zeros(0, np.uint64)

    def set_primid(self, primid):
        n = self.num_prims()
        dst = np.ctypeslib.as_array(self.L.orc_scene_primid(self.s), shape=(n,))
        dst[:] = np.asarray(primid, np.uint64)

    def prim_bounds(self, primid, close=False):
        out = np.zeros(6, np.float32)
        self.L.orc_prim_bounds(self.s, int(primid), int(close), _ptr(out))
        return out

    def prim_intersect(self, primid, ray, hit):
        self.L.orc_prim_intersect(self.s, int(primid), _ptr(ray), _ptr(hit))

    def build(self):
        self.a = self.L.orc_accel_build(self.s)
        return self

    def import_tree(self, nodes, aabb=None, primid=None):
        if primid is not None:
            self.set_primid(primid)
        nodes = np.ascontiguousarray(nodes, dtype=QBVH_NODE)
        ab = np.ascontiguousarray(aabb, np.float32) if aabb is not None else None
        self.a = self.L.orc_accel_import(self.s, _ptr(nodes), len(nodes), _ptr(ab))
        return self

    def nodes(self):
        n = int(self.L.orc_accel_num_nodes(self.a))
        buf = (C.c_char * (n * 256)).from_address(self.L.orc_accel_nodes(self.a))
        return np.frombuffer(buf, dtype=QBVH_NODE, count=n).copy()

    def aabb(self):
        return np.ctypeslib.as_array(self.L.orc_accel_aabb(self.a), shape=(6,)).copy()

    def check(self):
        st = np.zeros(4, np.uint64)
        return int(self.L.orc_accel_check(self.a, _ptr(st))), st

    def intersect(self, rays, max_dist=None, nthreads=None, counters=False):
        rays = np.ascontiguousarray(rays, dtype=RAY)
        out = np.zeros(len(rays), HITREC)
        md = np.ascontiguousarray(max_dist, np.float32) if max_dist is not None else None
        cnt = np.zeros(4, np.uint64) if counters else None
        self.L.orc_intersect_n(self.a, _ptr(rays), _ptr(md), _ptr(out), len(rays), nthreads or os.cpu_count(), _ptr(cnt))
        return (out, cnt) if counters else out

    def closest(self, rays, hits, centre):
        """accel_closest: returns (rays with updated min_dist, hits {prim,u,v,dist})"""
        rays = np.array(rays, dtype=RAY, copy=True)
        io = np.array(hits, dtype=HITREC, copy=True)
        c = np.ascontiguousarray(centre, np.float32)
        self.L.orc_closest_n(self.a, _ptr(rays), _ptr(io), _ptr(c), len(rays))
        return rays, io

    def visible(self, rays, max_dist, nthreads=None):
        rays = np.ascontiguousarray(rays, dtype=RAY)
        md = np.ascontiguousarray(max_dist, np.float32)
        out = np.zeros(len(rays), np.int32)
        self.L.orc_visible_n(self.a, _ptr(rays), _ptr(md), _ptr(out), len(rays), nthreads or os.cpu_count())
        return out

    def close(self):
        if self.a:
            self.L.orc_accel_free(self.a)
        if self.s:
            self.L.orc_scene_free(self.s)
        self.a = self.s = None


def ref_available(dbg=False):
    return os.path.exists(os.path.join(HERE, "_ref", "libcorona_ref_dbg.so" if dbg else "libcorona_ref.so"))


class Ref:
    """the compiled reference (one process-wide rt; one live scene at a time)"""
    _libs = {}

    @classmethod
    def lib(cls, dbg=False):
        if dbg not in cls._libs:
            path = os.path.join(HERE, "_ref", "libcorona_ref_dbg.so" if dbg else "libcorona_ref.so")
            L = C.CDLL(path)
            L.ref_init.restype = C.c_int
            L.ref_init.argtypes = [C.c_int]
            L.ref_prims_new.restype = C.c_void_p
            L.ref_prims_new.argtypes = [C.c_int]
            L.ref_prims_add_shape.argtypes = [C.c_void_p, C.c_void_p]
            L.ref_prims_load_geo.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
            L.ref_prims_finish.argtypes = [C.c_void_p]
            L.ref_prims_num.restype = C.c_uint64
            L.ref_prims_num.argtypes = [C.c_void_p]
            L.ref_prims_primid.restype = u64p
            L.ref_prims_primid.argtypes = [C.c_void_p]
            L.ref_prims_free.argtypes = [C.c_void_p]
            L.ref_prim_bounds.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p]
            L.ref_prim_intersect.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
            L.ref_prim_visible.restype = C.c_int
            L.ref_prim_visible.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_float]
            L.ref_accel_build.restype = C.c_void_p
            L.ref_accel_build.argtypes = [C.c_void_p]
            L.ref_accel_free.argtypes = [C.c_void_p]
            L.ref_accel_num_nodes.restype = C.c_uint64
            L.ref_accel_num_nodes.argtypes = [C.c_void_p]
            L.ref_accel_nodes.restype = C.c_void_p
            L.ref_accel_nodes.argtypes = [C.c_void_p]
            L.ref_accel_aabb.restype = f32p
            L.ref_accel_aabb.argtypes = [C.c_void_p]
            L.ref_intersect_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
            L.ref_intersect_hits.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
            L.ref_visible_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
            L.ref_closest_n.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
            L.ref_counters.restype = C.c_int
            L.ref_counters.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            cls._libs[dbg] = L
        return cls._libs[dbg]

    def __init__(self, scene, threads=1, dbg=False):
        self.L = self.lib(dbg)
        self.threads = self.L.ref_init(threads)   # first call fixes the pool size for the process
        self.scene = scene
        self._cshapes = scene.cshapes()
        self.p = self.L.ref_prims_new(len(scene.shapes))
        for i in range(len(scene.shapes)):
            self.L.ref_prims_add_shape(self.p, C.byref(self._cshapes[i]))
        self.L.ref_prims_finish(self.p)
        self.a = None

    def num_prims(self):
        return int(self.L.ref_prims_num(self.p))

    def primid(self):
        n = self.num_prims()
        return np.ctypeslib.as_array(self.L.ref_prims_primid(self.p), shape=(n,)).copy() if n else np.zeros(0, np.uint64)

    def prim_bounds(self, primid, close=False):
        out = np.zeros(6, np.float32)
        self.L.ref_prim_bounds(self.p, int(primid), int(close), _ptr(out))
        return out

    def prim_intersect(self, primid, ray, hit):
        self.L.ref_prim_intersect(self.p, int(primid), _ptr(ray), _ptr(hit))

    def build(self):
        self.a = self.L.ref_accel_build(self.p)
        return self

    def nodes(self):
        n = int(self.L.ref_accel_num_nodes(self.a))
        buf = (C.c_char * (n * 256)).from_address(self.L.ref_accel_nodes(self.a))
        return np.frombuffer(buf, dtype=QBVH_NODE, count=n).copy()

    def aabb(self):
        return np.ctypeslib.as_array(self.L.ref_accel_aabb(self.a), shape=(6,)).copy()

    def intersect(self, rays, max_dist=None, nthreads=None):
        rays = np.ascontiguousarray(rays, dtype=RAY)
        out = np.zeros(len(rays), HITREC)
        md = np.ascontiguousarray(max_dist, np.float32) if max_dist is not None else None
        self.L.ref_intersect_n(self.a, _ptr(rays), _ptr(md), _ptr(out), len(rays), min(nthreads or self.threads, self.threads))
        return out

    def intersect_hits(self, rays, hits):
        rays = np.ascontiguousarray(rays, dtype=RAY)
        self.L.ref_intersect_hits(self.a, _ptr(rays), _ptr(hits), len(rays))
        return hits

    def closest(self, rays, hits, centre):
        """accel_closest: returns (rays with updated min_dist, hits {prim,u,v,dist})"""
        rays = np.array(rays, dtype=RAY, copy=True)
        io = np.array(hits, dtype=HITREC, copy=True)
        c = np.ascontiguousarray(centre, np.float32)
        self.L.ref_closest_n(self.a, _ptr(rays), _ptr(io), _ptr(c), len(rays))
        return rays, io

    def visible(self, rays, max_dist, nthreads=None):
        rays = np.ascontiguousarray(rays, dtype=RAY)
        md = np.ascontiguousarray(max_dist, np.float32)
        out = np.zeros(len(rays), np.int32)
        self.L.ref_visible_n(self.a, _ptr(rays), _ptr(md), _ptr(out), len(rays), min(nthreads or self.threads, self.threads))
        return out

    def counters(self, reset=True):
        c = np.zeros(4, np.uint64)
        ok = self.L.ref_counters(self.a, _ptr(c), int(reset))
        return c if ok else None

    def close(self):
        if self.a:
            self.L.ref_accel_free(self.a)
        self.a = None
