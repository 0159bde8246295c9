#!/usr/bin/env python
"""bench.py -- rays/s (closest-hit + shadow) of the B200 hot path on the synthetic 10 M-triangle config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one progressive pass worth of traversal over a 4K frame on the seed-fixed procedural
10 M-triangle mesh (BASELINE.json configs[4]): a primary wave (one camera ray per pixel of the
3840x2176 padded frame), the diffuse-bounce wave leaving its hit points (offset origins, ignore prim:
prims_offset_ray, prims.c:374) and the next-event shadow wave toward the quad light -- the three kinds of
traversal calls pt/ptdl issue (pathspace.c:763, 329).  Rays are counted as calls to the traversal kernels,
the reference's accel_intersect count.

value  : whole-job rays/s with the ray waves already resident in HBM (CUDA events, max over ranks).
e2e    : the same step through the host-buffer C ABI (cb200_accel_intersect_n / _visible_n): pinned host
         rays in, hit records out, copies inside the timed region.
roofline: dominant kernel = closest-hit traversal of the primary+bounce waves; algorithmic bytes per ray =
         N_node*S_node + N_prim*S_prim + 40 + 24 with N_* from the instrumented kernel.
cpu_baseline / --impl reference: the unmodified reference (oracle/_ref, compiled in place) on the host cores,
         on a bounded sample of the same waves.
Multi-GPU: weak scaling, every rank traces its own sample slice (different seeds), no data-path collective.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_TRIS = 10_000_000
WIDTH, HEIGHT = 3840, 2176          # 4K padded to multiples of 32 (view.c:295-296)
LIGHT = (0.0, 0.0, 9.0)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(self.rows)}


def make_waves(cb, scene, trace, n_primary, seed, frame=None):
    """primary / bounce / shadow ray sets; `trace(rays)` returns closest hits (GPU arm: the GPU; reference arm: the reference)"""
    S = cb.scenes
    prim = S.camera_rays(n_primary, scene, seed=100 + seed, frame=frame)
    hits = trace(prim)
    bounce = S.bounce_rays(prim, hits, seed=200 + seed)
    shadow, smd = S.shadow_rays(prim, hits, LIGHT, seed=300 + seed)
    return prim, bounce, shadow, smd


def run_reference(args, rank, world):
    """--impl reference: accel_build + accel_intersect of the unmodified reference on all host cores"""
    if rank != 0:
        return
    cb = importlib.import_module("corona-13_b200")
    from oracle.binding import Ref, ref_available
    if not ref_available():
        from oracle.binding import Oracle   # reference not compiled here: time the oracle port instead
    cores = os.cpu_count() or 1
    scene = cb.scenes.synthetic_scene(NUM_TRIS, seed=1)
    kind = "reference" if ref_available() else "port"
    if kind == "reference":
        impl = Ref(scene, threads=cores).build()
        trace = lambda r, md=None: impl.intersect(r, md, nthreads=cores)
        vis = lambda r, md: impl.visible(r, md, nthreads=cores)
    else:
        impl = Oracle(scene).build()
        trace = lambda r, md=None: impl.intersect(r, md, nthreads=cores)
        vis = lambda r, md: impl.visible(r, md, nthreads=cores)
    n_sample = 1 << 19
    prim, bounce, shadow, smd = make_waves(cb, scene, trace, n_sample, 0, (1024, 512))
    rays_per_step = len(prim) + len(bounce) + len(shadow)

    def step():
        trace(prim)
        trace(bounce)
        # ptdl's next-event visibility goes through accel_intersect as well (path_visible, pathspace.c:311-344)
        trace(shadow, smd)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = rays_per_step * args.steps / dt
    sample = f"{len(prim)} primary + {len(bounce)} bounce + {len(shadow)} shadow rays per step of the 4K wave, 10M-tri mesh, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "rays/s (closest-hit + shadow)", "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic 10M-triangle procedural mesh, 4K frame, primary+bounce+shadow waves", "num_tris": scene.num_prims},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--tris", type=int, default=NUM_TRIS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    cb = importlib.import_module("corona-13_b200")
    lib = importlib.import_module("corona-13_b200.lib")     # raises when the CUDA library is missing
    R = cb.records
    if lib.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (" + lib.load().cb200_last_error().decode() + "); there is no CPU fallback")
    torch.cuda.set_device(local)
    lib.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    scene = cb.scenes.synthetic_scene(args.tris, seed=1)
    acc = lib.Accel(scene)
    t0 = time.perf_counter()
    acc.build()
    build_s = time.perf_counter() - t0
    node_b, prim_b = acc.layout()
    n_primary = WIDTH * HEIGHT
    prim, bounce, shadow, smd = make_waves(cb, scene, lambda r: acc.intersect(r), n_primary, rank, (WIDTH, HEIGHT))
    waves = [("primary", prim, None), ("bounce", bounce, None), ("shadow", shadow, smd)]
    rays_per_step = sum(len(w[1]) for w in waves)

    st = torch.cuda.current_stream().cuda_stream
    dev = []
    for name, rays, md in waves:
        d_r = torch.from_numpy(rays.view("u1").reshape(-1)).cuda()
        d_md = torch.from_numpy(md).cuda() if md is not None else None
        d_o = torch.empty(len(rays) * (24 if md is None else 4), dtype=torch.uint8, device="cuda")
        dev.append((name, d_r, d_md, d_o, len(rays)))

    # per-ray work of the dominant kernel (outside the timed region), reference ACCEL_DEBUG definitions
    cnt = np.zeros(4, np.float64)
    for name, d_r, d_md, d_o, n in dev[:2]:
        cnt += acc.intersect_counted(d_r.data_ptr(), 0, d_o.data_ptr(), n).astype(np.float64)
    n_node, n_prim = cnt[1] / cnt[0], cnt[3] / cnt[0]
    bytes_per_ray = n_node * node_b + n_prim * prim_b + 40 + 24

    def step(events=None):
        for k, (name, d_r, d_md, d_o, n) in enumerate(dev):
            if events is not None:
                events[k][0].record()
            if d_md is None:
                acc.intersect_dev(d_r.data_ptr(), 0, d_o.data_ptr(), n, st)
            else:
                acc.visible_dev(d_r.data_ptr(), d_md.data_ptr(), d_o.data_ptr(), n, st)
            if events is not None:
                events[k][1].record()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in dev] for _ in range(args.steps)]
    l0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for s in range(args.steps):
        step(ev[s])
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = lib.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.summary() if rank == 0 else None
    k_ms = [np.mean([ev[s][k][0].elapsed_time(ev[s][k][1]) for s in range(args.steps)]) for k in range(len(dev))]

    # e2e: host buffers through the C ABI, copies inside the timed region
    pinned = []
    for name, rays, md in waves:
        pr = torch.from_numpy(rays.view("u1").reshape(-1).copy()).pin_memory()
        pm = torch.from_numpy(md.copy()).pin_memory() if md is not None else None
        po = torch.empty(len(rays) * (24 if md is None else 4), dtype=torch.uint8).pin_memory()
        pinned.append((pr, pm, po, len(rays)))
    L = lib.load()

    def e2e_step():
        for pr, pm, po, n in pinned:
            if pm is None:
                rc = L.cb200_accel_intersect_n(acc.a, pr.data_ptr(), None, po.data_ptr(), n)
            else:
                rc = L.cb200_accel_visible_n(acc.a, pr.data_ptr(), pm.data_ptr(), po.data_ptr(), n)
            if rc:
                raise SystemExit("e2e step failed: " + L.cb200_last_error().decode())
    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    h2d = sum(n * 40 + (n * 4 if pm is not None else 0) for pr, pm, po, n in pinned)
    d2h = sum(n * (24 if pm is None else 4) for pr, pm, po, n in pinned)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm, which = peaks()
    closest_rays = dev[0][4] + dev[1][4]
    closest_ms = k_ms[0] + k_ms[1]
    achieved = closest_rays * bytes_per_ray / (closest_ms * 1e-3) / 1e9
    out = {
        "metric": "rays/s (closest-hit + shadow)", "value": rays_per_step * world * args.steps / (ms_total * 1e-3), "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "synthetic 10M-triangle procedural mesh, 4K frame, primary+bounce+shadow waves",
                   "num_tris": scene.num_prims, "frame": [WIDTH, HEIGHT], "rays_per_step_per_gpu": rays_per_step,
                   "waves": {w[0]: len(w[1]) for w in waves}, "l2": "inputs larger than L2 (ray waves 335+ MB each, scene 0.8 GB)",
                   "bvh": {"nodes": acc.num_nodes(), "depth": acc.depth(), "node_bytes": node_b, "prim_bytes": prim_b,
                           "gpu_build_s": build_s}},
        "kernel_ms": {w[0]: float(k) for w, k in zip(waves, k_ms)},
        "grays_per_s": {w[0]: len(w[1]) / (k * 1e-3) / 1e9 for w, k in zip(waves, k_ms)},
        "roofline": {"bound": "hbm", "kernel": "k_intersect (primary+bounce waves)", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                     "frac": achieved / hbm, "peak_source": which, "traffic": None,
                     "bytes_per_ray": bytes_per_ray, "nodes_per_ray": n_node, "prims_per_ray": n_prim},
        "e2e": {"value": rays_per_step * world * e2e_steps / float(e2e_t.item()), "unit": "rays/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(cb, scene, waves)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(cb, scene, waves):
    """the reference (oracle/_ref) -- or the oracle port where it is not built -- on a bounded sample of the same waves"""
    from oracle.binding import Ref, Oracle, ref_available
    cores = os.cpu_count() or 1
    n = 1 << 19
    kind = "reference" if ref_available() else "port"
    impl = Ref(scene, threads=cores).build() if kind == "reference" else Oracle(scene).build()
    t0 = time.perf_counter()
    total = 0
    for name, rays, md in waves:
        r = rays[:n]
        impl.intersect(r, md[:n] if md is not None else None, nthreads=cores)
        total += len(r)
    dt = time.perf_counter() - t0
    impl.close()
    return {"value": total / dt, "unit": "rays/s", "cores": cores, "kind": kind,
            "sample": f"first {n} rays of each of the 3 waves of rank 0's step, own tree built by accel_build with {cores} threads"}


if __name__ == "__main__":
    main()
