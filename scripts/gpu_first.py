"""scratch GPU bring-up: mode A / mode B parity and first timings"""
import sys, os, time, importlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
cb = importlib.import_module("corona-13_b200")
lib = importlib.import_module("corona-13_b200.lib")
from oracle.binding import Oracle, Ref, ref_available
S, R = cb.scenes, cb.records

def cmp(name, a, b):
    eq = np.array_equal(a.view('u1'), b.view('u1'))
    pa, pb = R.hit_prim64(a), R.hit_prim64(b)
    nd = int((pa != pb).sum())
    dd = int((a['dist'].view('u4') != b['dist'].view('u4')).sum())
    du = int(((a['u'].view('u4') != b['u'].view('u4')) | (a['v'].view('u4') != b['v'].view('u4'))).sum())
    print(f"  {name}: bitexact={eq} prim_diff={nd} dist_diff={dd} uv_diff={du} of {len(a)} hit_rate={(pa != R.INVALID_PRIMID).mean():.3f}", flush=True)
    return eq

print("devices", lib.device_count(), lib.load().cb200_version())
for cfg in [dict(num_tris=20000, analytic=True), dict(num_tris=20000, analytic=True, motion=True), dict(num_tris=30000, quads=True)]:
    sc = S.synthetic_scene(**cfg)
    tm = 1.0 if cfg.get('motion') else 0.0
    print("scene", cfg, sc.num_prims)
    orc = Oracle(sc).build()
    rays = np.concatenate([S.camera_rays(100000, sc, time_max=tm), S.random_rays(100000, sc, time_max=tm)])
    ho = orc.intersect(rays)
    br = S.bounce_rays(rays, ho)
    hb = orc.intersect(br)
    sr, smd = S.shadow_rays(rays, ho, (0, 0, 9.0))
    vo = orc.visible(sr, smd)
    acc = lib.Accel(sc)
    # mode A
    acc.import_qbvh(orc.nodes(), orc.primid(), orc.aabb())
    cmp("modeA primary", acc.intersect(rays), ho)
    cmp("modeA bounce ", acc.intersect(br), hb)
    v = acc.visible(sr, smd)
    print("  modeA visible equal", np.array_equal(v, vo), v.mean())
    # mode B
    t = time.time(); acc.build(); print("  gpu build s", time.time()-t, "nodes", acc.num_nodes(), "depth", acc.depth(), "layout", acc.layout())
    nodes, primid = acc.export_qbvh()
    chk = Oracle(sc).import_tree(nodes, acc.aabb(), primid)
    print("  tree check", chk.check(), "aabb eq", np.array_equal(acc.aabb(), orc.aabb()))
    hg = acc.intersect(rays)
    cmp("modeB primary vs oracle(ref tree)", hg, ho)
    cmp("modeB primary vs oracle(gpu tree)", hg, chk.intersect(rays))
    cmp("modeB bounce vs oracle(gpu tree) ", acc.intersect(br), chk.intersect(br))
    v = acc.visible(sr, smd)
    print("  modeB visible equal", np.array_equal(v, vo), int((v != vo).sum()))
    acc.close()

# timing on bigger scenes
import torch
for nt in [1000000, 10000000]:
    t = time.time(); sc = S.synthetic_scene(nt); print("gen", nt, time.time()-t, flush=True)
    acc = lib.Accel(sc)
    t = time.time(); acc.build(); torch.cuda.synchronize(); print("gpu build s", time.time()-t, "nodes", acc.num_nodes(), "depth", acc.depth(), flush=True)
    n = 1 << 22
    for nm, rays in [("camera", S.camera_rays(n, sc)), ("random", S.random_rays(n, sc))]:
        d_r = torch.from_numpy(rays.view('u1').reshape(-1)).cuda()
        d_o = torch.zeros(n*24, dtype=torch.uint8, device='cuda')
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3): acc.intersect_dev(d_r.data_ptr(), 0, d_o.data_ptr(), n, st)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5): acc.intersect_dev(d_r.data_ptr(), 0, d_o.data_ptr(), n, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/5
        cnt = acc.intersect_counted(d_r.data_ptr(), 0, d_o.data_ptr(), n)
        print(f"  {nm}: {n/ms/1e6:.3f} Grays/s  ms={ms:.3f} counters/ray={cnt[1:]/cnt[0]}", flush=True)
    acc.close()
