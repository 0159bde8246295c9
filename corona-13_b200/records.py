"""numpy views of the records in include/corona_types.h (which restate the reference's
ray_t / hit_t / primid_t / prims_vtx_t layouts, include/corona_common.h:45-137,
include/prims.h:20-47).  Used by tests, bench and the ctypes bindings only."""
import ctypes as C
import numpy as np

RAY = np.dtype([("pos", "<f4", 3), ("dir", "<f4", 3), ("time", "<f4"), ("min_dist", "<f4"),
                ("ignore", "<u4", 2)])
HITREC = np.dtype([("prim", "<u4", 2), ("u", "<f4"), ("v", "<f4"), ("dist", "<f4"), ("pad", "<u4")])
HIT = np.dtype([("prim", "<u4", 2), ("u", "<f4"), ("v", "<f4"), ("w", "<f4"),
                ("r", "<f4"), ("s", "<f4"), ("t", "<f4"),
                ("a", "<f4", 3), ("b", "<f4", 3), ("n", "<f4", 3), ("x", "<f4", 3), ("gn", "<f4", 3),
                ("shader", "<i4"), ("dist", "<f4")])
VTX = np.dtype([("v", "<f4", 3), ("n", "<u4")])
VTXIDX = np.dtype([("v", "<u4"), ("uv", "<u4")])
QBVH_NODE = np.dtype([("aabb0", "<f4", (6, 4)), ("aabb1", "<f4", (6, 4)), ("child", "<u8", 4),
                      ("parent", "<u8"), ("axis0", "<i8"), ("axis00", "<i8"), ("axis01", "<i8")])
assert RAY.itemsize == 40 and HITREC.itemsize == 24 and HIT.itemsize == 100
assert VTX.itemsize == 16 and VTXIDX.itemsize == 8 and QBVH_NODE.itemsize == 256

INVALID_PRIMID = np.uint64(0xFFFFFFFFFFFFFFFF)
LEAF_BIT = np.uint64(1 << 63)
FLT_MAX = np.float32(3.4028234663852886e38)

PRIM_SPHERE, PRIM_LINE, PRIM_TRI, PRIM_QUAD = 1, 2, 3, 4


def primid_make(extra, shapeid, vi, mb, vcnt):
    """pack primid fields (corona_common.h:45-53): lo = extra:3|shapeid:29, hi = vi:28|mb:1|vcnt:3"""
    vi = np.asarray(vi, dtype=np.uint64)
    return (np.uint64(extra) & np.uint64(7)) | (np.uint64(shapeid) << np.uint64(3)) \
        | (vi << np.uint64(32)) | (np.uint64(mb) << np.uint64(60)) \
        | (np.asarray(vcnt, dtype=np.uint64) << np.uint64(61))


def primid_vcnt(p):
    return (np.asarray(p, dtype=np.uint64) >> np.uint64(61)) & np.uint64(7)


def primid_mb(p):
    return (np.asarray(p, dtype=np.uint64) >> np.uint64(60)) & np.uint64(1)


def primid_vi(p):
    return (np.asarray(p, dtype=np.uint64) >> np.uint64(32)) & np.uint64(0x0FFFFFFF)


def primid_shapeid(p):
    return (np.asarray(p, dtype=np.uint64) >> np.uint64(3)) & np.uint64(0x1FFFFFFF)


def hit_prim64(h):
    """prim field (two u32 words) of a HITREC/HIT array as uint64"""
    p = np.ascontiguousarray(h["prim"])
    return p.view("<u8").reshape(p.shape[:-1])


class CShape(C.Structure):
    """cb_shape_t (include/corona_types.h)"""
    _fields_ = [("primid", C.c_void_p), ("num_prims", C.c_uint64),
                ("vtxidx", C.c_void_p), ("num_vtxidx", C.c_uint64),
                ("vtx", C.c_void_p), ("num_vtx", C.c_uint64),
                ("material", C.c_int64)]


def make_rays(pos, dir, time=0.0, min_dist=0.0, ignore=None):
    n = len(pos)
    r = np.zeros(n, dtype=RAY)
    r["pos"] = pos
    r["dir"] = dir
    r["time"] = time
    r["min_dist"] = min_dist
    ig = np.full(n, INVALID_PRIMID, dtype=np.uint64) if ignore is None else np.asarray(ignore, np.uint64)
    r["ignore"] = ig.view("<u4").reshape(n, 2)
    return r
