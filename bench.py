#!/usr/bin/env python
"""bench.py -- rays/s (closest-hit + shadow) and spp/s of the B200 path-tracing hot path on the synthetic 10 M-triangle config.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[4]): seed-fixed procedural 10 M-triangle mesh (fBm terrain + icosphere soup, one
quad light), 3840x2176 frame (4K padded to multiples of 32, view.c:295-296), ptdl integrator (path tracing with next-event
estimation, the reference's 0011_ptdl sampler), `rand` point sampler, one wavelength per path.

One "step" = one progression of the reference's view_render(): W*H path indices = 1 sample per pixel per GPU, every path
traced to its end (camera ray, closest-hit traversal per vertex, material + BSDF, next-event shadow ray, splat).  Rays are
counted as calls to the traversal kernels, the reference's accel_intersect count (path_propagate + path_visible).

value    : whole-job rays/s over the K timed steps, scene + BVH resident in HBM, CUDA events on the launching stream, max over
           ranks.  Steps are issued back to back through the streaming entry (cb200_render_pass_stream); the stragglers
           of the last step are flushed INSIDE the timed region, so every path started in it is also finished in it.
           spp_per_s / paths_per_s ride along (the metric's second half).
e2e      : the same passes through the host-side render module the reference would link (MOD_render=b200,
           render_b200_pass per progression + render_b200_finish at the end): path-index range in, HOST framebuffer out after
           every progression -- a device->host copy of the W*H*3 float image per step inside the timed region.  The inputs of this path ARE an index range (render_sample_path(i),
           gi.c:81): h2d is the 16 bytes that name it.  `e2e_accel` gives the accel.h boundary (host ray batches in, hit
           records out) for the same scene.
roofline : dominant kernel = closest-hit traversal (k_intersect).  achieved = algorithmic bytes of all its launches in the
           timed region / their summed duration (CUDA event pairs recorded around each launch on the launching stream);
           algorithmic bytes per ray = N_node*S_node + N_prim*S_prim + 40 + 24 with N_* counted by the
           instrumented kernel on one untimed progression of the same workload (ACCEL_DEBUG definitions).
cpu_baseline / --impl reference: the UNMODIFIED reference renderer (oracle/_ref/corona_ptdl_rand, compiled in place from the
           reference sources by oracle/Makefile) on the same scene, camera, materials and frame size, all host cores, its own
           binned-SAH build; ray counts from its -DACCEL_DEBUG twin.
Multi-GPU: weak scaling by splitting samples per pixel: rank g renders progressions g, g+N, ... (path-index ranges),
           one NCCL reduce of the framebuffer to rank 0 per progression (torch.distributed on a side stream, overlapped with the
           next progression), all inside the timed region.
"""
import argparse
import ctypes as C
import importlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_TRIS = 10_000_000
WIDTH, HEIGHT = 3840, 2176          # 4K padded to multiples of 32 (view.c:295-296)
CAMERA = dict(pos=(18.0, 14.0, 14.5), lookat=(0.0, 0.0, 2.5), aperture_value=6, exposure_value=13, focal_length=0.4, iso=100.0)
WORKLOAD = "synthetic 10M-triangle procedural mesh, 4K frame (3840x2176), ptdl, 1 spp per GPU per step"
METRIC = "rays/s (closest-hit + shadow)"
REFDIR = os.path.join(ROOT, "oracle", "_ref")


NCU_SUMMARY = "profiles/r3u_k_intersect_ncu.json"
NCU_COMMAND = ("BENCH_NO_WARM=1 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 3 -c 1 -o gpurun_out/r3u_k_intersect "
               "python scripts/gpu_render_bench.py; python scripts/ncu_summary.py gpurun_out/r3u_k_intersect.ncu-rep profiles/r3u_k_intersect_ncu.json")
NCU_RAYS_PER_LAUNCH = 18.0e6   # the captured launch: one streamed wave (new paths + survivors); 431 MB of 24-byte hit records written


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel from the committed `ncu --set full` capture
    of the same workload (NCU_SUMMARY: one full streamed wave); None when the summary is not there"""
    p = os.path.join(ROOT, NCU_SUMMARY)
    try:
        rows = json.load(open(p))
        tot = []
        for r in rows:
            rd = [v for k, v in r.items() if k.startswith("dram__bytes_read.sum")]
            wr = [v for k, v in r.items() if k.startswith("dram__bytes_write.sum")]
            unit_r = [k for k in r if k.startswith("dram__bytes_read.sum")][0]
            unit_w = [k for k in r if k.startswith("dram__bytes_write.sum")][0]
            scale = lambda u: 1e9 if "Gbyte" in u else 1e6 if "Mbyte" in u else 1e3 if "Kbyte" in u else 1.0
            tot.append((wr[0] * scale(unit_w), rd[0] * scale(unit_r) + wr[0] * scale(unit_w)))
        full = [t for w, t in tot if w >= 0.9 * max(w for w, _ in tot)]   # hit records written ~ rays of the launch
        return sum(full) / len(full), os.path.relpath(p, ROOT)
    except Exception:
        return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    return 6650.0, "B200_PROFILING.md fallback"


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
            time.sleep(0.02)

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


def bench_scene(cb, tris):
    """scene, flattened materials, camera: identical for both arms"""
    IO, S = cb.scene_io, cb.scenes
    z = np.load(os.path.join(ROOT, "corona-13_b200", "data", "bench_materials.npz"))
    ms = IO.MaterialSet()
    raw = z["materials"].tobytes()
    ms.materials = list((IO.CMaterial * (len(raw) // C.sizeof(IO.CMaterial))).from_buffer_copy(raw))
    scene = S.synthetic_scene(tris, seed=1)
    for sh, m in zip(scene.shapes, z["shape_mats"]):
        sh.material = int(m)
    return scene, ms, IO.Camera(**CAMERA), [str(x) for x in z["shader_lines"]], [int(m) for m in z["shape_mats"]]


# ------------------------------------------------------------------------------------------------ reference arm
class ReferenceRenderer:
    """the unmodified reference binary on the bench scene written out in its own file formats"""

    def __init__(self, cb, scene, lines, shape_mats, camera):
        if not os.path.exists(os.path.join(REFDIR, "corona_ptdl_rand")):
            raise RuntimeError("oracle/_ref/corona_ptdl_rand is not built (python -c 'import __graft_entry__ as g; g.build()' "
                               "where /root/reference exists)")
        IO = cb.scene_io
        self.IO = IO
        self.tmp = tempfile.mkdtemp(prefix="corona_bench_")
        shapes = []
        for i, sh in enumerate(scene.shapes):
            sh.write_geo(os.path.join(self.tmp, f"shape{i}.geo"))
            shapes.append((shape_mats[i], f"shape{i}"))
        self.nra2 = os.path.join(self.tmp, "test.nra2")
        IO.write_nra2(self.nra2, lines, shapes)
        camera.write(os.path.join(self.tmp, "test01.cam"))
        self.cores = os.cpu_count() or 1

    def run(self, binary, w, h, spp, frame=1):
        """returns (seconds per progression [list], accel build seconds, total accel_intersect calls or None); the image the
        run leaves behind (<scene>render_fb00.pfm, fb_export at exit) is read by image()"""
        cmd = [os.path.join(REFDIR, binary), self.nra2, "-x", "-s", str(spp), "-w", str(w), "-h", str(h), "-b", "0",
               "-t", str(self.cores), "--frame", str(frame)]
        p = subprocess.run(cmd, cwd=REFDIR, capture_output=True, text=True)
        self.last_stderr = p.stderr
        if p.returncode != 0:
            raise RuntimeError(f"{binary} failed ({p.returncode}): {p.stderr[-400:]}")
        frames = [float(x) for x in re.findall(r"([0-9.]+) s/frame, \d+ spp", p.stdout)]
        build = re.findall(r"construction took ([0-9.]+) seconds", p.stdout)
        rays = [int(x) for x in re.findall(r"accel_intersect: (\d+)", p.stderr)]
        return frames, (float(build[0]) if build else None), (sum(rays) if rays else None)

    def image(self):
        return self.IO.read_pfm(os.path.splitext(self.nra2)[0] + "render_fb00.pfm")

    def rays_per_path(self, counters=False):
        """-DACCEL_DEBUG twin at a quarter of the frame in each direction (rays per path do not depend on resolution).
        counters=True also returns the reference's own per-ray traversal work on ITS binned-SAH tree of this scene
        (qbvhmp.c:1168-1171: calls, node tests with >= 1 child hit, child boxes hit, primitive tests -- summed over the
        closest-hit and the next-event calls, which both go through accel_intersect in the reference)"""
        w, h = WIDTH // 4, HEIGHT // 4
        _, _, rays = self.run("corona_ptdl_rand_dbg", w, h, 1)
        if not counters:
            return rays / float(w * h)
        m = re.findall(r"accel_intersect: (\d+) aabb_intersect (\d+) / (\d+) prims_intersect (\d+)", self.last_stderr)
        tot = np.array([[int(x) for x in row] for row in m], np.float64).sum(axis=0) if m else None
        dbg = None
        if tot is not None and tot[0] > 0:
            dbg = {"rays": tot[0], "nodes_per_ray": tot[2] / tot[0], "child_boxes_hit_per_ray": tot[1] / tot[0], "prims_per_ray": tot[3] / tot[0],
                   "note": "reference renderer, own binned-SAH tree, closest-hit + next-event calls together (qbvhmp.c ACCEL_DEBUG)"}
        return rays / float(w * h), dbg

    def close(self):
        shutil.rmtree(self.tmp, ignore_errors=True)


PARITY_SPP = 2            # progressions per image of the bench-scene parity check (reference: ~2 s each on 32 threads)
PARITY_BLOCK = 16         # compared after averaging 16x16 pixel blocks: 8.4 M pixels at 2 spp are mostly noise
REF_ARM_IMAGE = os.path.join(tempfile.gettempdir(), "corona_b200_reference_arm.npz")


def downsample(img, k=PARITY_BLOCK):
    h, w, c = img.shape
    return img[:h // k * k, :w // k * k].astype(np.float64).reshape(h // k, k, w // k, k, c).mean(axis=(1, 3))


def image_parity(gpu, ref_a, ref_b):
    """bench-scene image parity: the GPU image against the reference renderer's image of the same scene, frame size and sample count
    (rand point sampler: independent streams), with the reference's own seed-to-seed difference as the noise floor"""
    g, a, b = downsample(gpu), downsample(ref_a), downsample(ref_b)
    rel = lambda x, y: float(np.sqrt(((x - y) ** 2).mean()) / max(x.mean(), 1e-30))
    floor = rel(a, b)
    got = min(rel(a, g), rel(b, g))
    both = 0.5 * (ref_a.astype(np.float64).mean(axis=(0, 1)) + ref_b.astype(np.float64).mean(axis=(0, 1)))
    means = (gpu.astype(np.float64).mean(axis=(0, 1)) / both).tolist()
    ok = bool(got <= 1.15 * floor and all(abs(m - 1) < 0.01 for m in means))
    return {"ok": ok, "spp": PARITY_SPP, "block": PARITY_BLOCK, "rel_rmse_gpu_vs_reference": got, "rel_rmse_reference_seed_to_seed": floor,
            "limit": "rel_rmse <= 1.15 x floor, channel means within 1 %", "channel_mean_ratio_gpu_over_reference": means,
            "reference_channel_means": both.tolist(),
            "what": f"{PARITY_SPP}-spp 4K images of the bench scene (ptdl, rand), {PARITY_BLOCK}x{PARITY_BLOCK} block means; reference = oracle/_ref/corona_ptdl_rand, "
                    "--frame 1 and --frame 2"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, 1 spp at 4K per step"""
    if rank != 0:
        return
    cb = importlib.import_module("corona-13_b200")
    scene, ms, cam, lines, shape_mats = bench_scene(cb, args.tris)
    try:
        ref = ReferenceRenderer(cb, scene, lines, shape_mats, cam)
    except RuntimeError as e:     # oracle/_ref was not built where this snapshot was taken
        print(json.dumps({"impl": "reference", "unavailable": str(e)[:200]}))
        return
    try:
        rpp = ref.rays_per_path()
        frames, build_s, _ = ref.run("corona_ptdl_rand", WIDTH, HEIGHT, args.warmup + args.steps)
        img = ref.image()
        # kept for the own arm, which the driver runs next on the same box: the image of these progressions, block-averaged
        try:
            np.savez(REF_ARM_IMAGE, image=downsample(img).astype(np.float32), spp=np.int64(args.warmup + args.steps),
                     means=img.astype(np.float64).mean(axis=(0, 1)), tris=np.int64(scene.num_prims))
        except OSError:
            pass
    finally:
        ref.close()
    timed = frames[args.warmup:args.warmup + args.steps]
    dt = sum(timed)
    paths = WIDTH * HEIGHT * len(timed)
    value = paths * rpp / dt
    sample = (f"{len(timed)} progressions of the full 4K frame ({WIDTH * HEIGHT} paths each) after {args.warmup} warm-up progressions, "
              f"{ref.cores} pinned threads, own binned-SAH tree ({build_s} s build); rays = paths x {rpp:.3f} rays/path counted by the "
              f"-DACCEL_DEBUG build at {WIDTH // 4}x{HEIGHT // 4}")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world,
        "steps": len(timed), "warmup": args.warmup, "ms_per_step": dt / len(timed) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "num_tris": scene.num_prims, "frame": [WIDTH, HEIGHT], "sampler": "ptdl", "pointsampler": "rand"},
        "spp_per_s": len(timed) / dt, "paths_per_s": paths / dt, "rays_per_path": rpp,
        "image_mean_xyz": [float(x) for x in img.astype(np.float64).mean(axis=(0, 1))], "image_spp": args.warmup + args.steps,
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": ref.cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(cb, scene, lines, shape_mats, cam):
    """bounded sample for the own arm's line: one 4K progression of the reference renderer (plus one warm-up)"""
    ref = ReferenceRenderer(cb, scene, lines, shape_mats, cam)
    try:
        rpp, dbg = ref.rays_per_path(counters=True)
        frames, build_s, _ = ref.run("corona_ptdl_rand", WIDTH, HEIGHT, PARITY_SPP, frame=1)
        img_a = ref.image()
        ref.run("corona_ptdl_rand", WIDTH, HEIGHT, PARITY_SPP, frame=2)
        img_b = ref.image()
    finally:
        ref.close()
    dt = frames[-1]
    return {"value": WIDTH * HEIGHT * rpp / dt, "unit": "rays/s", "cores": ref.cores, "kind": "reference",
            "spp_per_s": 1.0 / dt, "rays_per_path": rpp, "accel_build_s": build_s, "accel_debug": dbg,
            "sample": f"the last of {PARITY_SPP} 4K progressions ({WIDTH * HEIGHT} paths each) of oracle/_ref/corona_ptdl_rand on the same scene files, "
                      f"{ref.cores} pinned threads; rays = paths x {rpp:.3f} (ACCEL_DEBUG build)"}, img_a, img_b


# ------------------------------------------------------------------------------------------------ own arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
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

    # stdout carries exactly ONE line, the JSON: whatever libraries print on file descriptor 1 while the job runs (NCCL's version
    # banner, for one) is sent to stderr; the descriptor is restored just before the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    cb = importlib.import_module("corona-13_b200")
    lib = importlib.import_module("corona-13_b200.lib")     # raises when the CUDA library is missing
    IO = cb.scene_io
    if lib.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (" + lib.load().cb200_last_error().decode() + "); there is no CPU fallback")
    torch.cuda.set_device(local)
    lib.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3)

    scene, ms, cam, lines, shape_mats = bench_scene(cb, args.tris)
    acc = lib.Accel(scene)
    t0 = time.perf_counter()
    acc.build()
    build_s = time.perf_counter() - t0
    node_b, prim_b = acc.layout()
    n_pass = WIDTH * HEIGHT
    st = torch.cuda.current_stream().cuda_stream
    r = lib.Render(acc, cam, ms, WIDTH, HEIGHT, sampler=IO.SAMPLER_PTDL, pointsampler=IO.POINTS_RAND, frame=1, rank=rank, world=world)

    # rank g renders progressions g, g+N, g+2N, ...: the same path indices a 1-GPU run would use for those progressions
    P = importlib.import_module("corona-13_b200.progressive")
    prog = [0]

    def next_first():
        first, _ = P.progression_range(prog[0], rank, world, n_pass)
        prog[0] += 1
        return first

    # per-ray work of the dominant kernel, reference ACCEL_DEBUG definitions: one untimed, counted progression
    r.instrument(False, True)
    r.render_pass(next_first(), n_pass, st)
    cs = r.stats()
    tc = np.array(cs["trav_closest"], np.float64)
    n_node, n_prim = tc[1] / tc[0], tc[3] / tc[0]
    bytes_per_ray = n_node * node_b + n_prim * prim_b + 40 + 24
    r.clear()
    r.instrument(True, False)

    # framebuffers: double-buffered so that the reduce of progression s overlaps the rendering of s+1 (N > 1)
    red = P.FramebufferReducer(HEIGHT, WIDTH, "cuda", rank, world, dist if world > 1 else None)

    def step(s, last=False):
        r.set_framebuffer(red.acquire(s).data_ptr())
        r.render_pass(next_first(), n_pass, st, streaming=not last)     # the last one flushes the stragglers
        red.submit(s)

    for s in range(warmup):
        step(s, last=(s == warmup - 1))
    red.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    r.clear()
    r.instrument(True, False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for s in range(args.steps):
        step(s, last=(s == args.steps - 1))
    fb_sum = red.finish()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = lib.launch_count() - l0
    stt = r.stats()
    ms_t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    cnt = torch.tensor([stt["rays_closest"], stt["rays_shadow"], stt["paths"], launches], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_total = float(ms_t.item())
    rays_closest, rays_shadow, paths, launches_all = [float(x) for x in cnt.tolist()]
    clocks = sampler.summary() if rank == 0 else None
    image_mean = None
    if rank == 0:
        image_mean = [float(x) for x in (fb_sum.mean(dim=(0, 1)) * (cam.iso / (100.0 * args.steps * world))).tolist()]

    # ---- e2e: host-side render module (MOD_render=b200), host framebuffer after every progression.  N > 1: every rank renders its
    #      progressions, the per-progression reduce runs as in the device-timed loop, and rank 0 copies the summed image to the host
    #      after every progression (what the reference's display / fb file sees)
    r.clear()
    red.clear()
    r.instrument(False, False)
    host_fb = torch.empty(HEIGHT, WIDTH, 3).pin_memory()
    e2e_steps = max(2, min(args.steps, 64))   # like the device-timed loop: the final flush of the stragglers is inside the timing and amortised over the same number of steps
    L = lib.load()
    if world > 1 and rank == 0:
        red.host_mirror = host_fb     # the root's running sum is copied to the host after every accumulate, on the reducer's side stream

    def e2e_step(s):
        if world == 1:           # == host/render_b200.c: render_b200_pass(r, first, count, fb)
            lib._check(L.cb200_render_pass_stream(r.r, next_first(), n_pass, None), "cb200_render_pass_stream")
            lib._check(L.cb200_render_snapshot_async(r.r, host_fb.data_ptr(), None), "cb200_render_snapshot_async")
        else:
            step(s)

    def e2e_finish(s):           # == render_b200_finish(r, fb)
        if world == 1:
            lib._check(L.cb200_render_flush(r.r, None), "cb200_render_flush")
            lib._check(L.cb200_render_download(r.r, host_fb.data_ptr(), None), "cb200_render_download")
        else:
            r.set_framebuffer(red.acquire(s + 1).data_ptr())     # the last step's buffer is being reduced: stragglers go to the next one
            lib._check(L.cb200_render_flush(r.r, None), "cb200_render_flush")
            red.submit(s + 1)
            red.finish()
    if world == 1:
        r.set_framebuffer(0)
    e2e_step(0)
    e2e_finish(0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    s1 = r.stats()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        e2e_step(s)
    e2e_finish(e2e_steps - 1)
    torch.cuda.synchronize()
    dt_t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    s2 = r.stats()
    rays_t = torch.tensor([(s2["rays_closest"] + s2["rays_shadow"]) - (s1["rays_closest"] + s1["rays_shadow"])], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays_t, op=dist.ReduceOp.SUM)
    dt, rays = float(dt_t.item()), float(rays_t.item())
    e2e = {"value": rays / dt, "unit": "rays/s", "h2d_bytes_per_step": 16,
           "d2h_bytes_per_step": HEIGHT * WIDTH * 3 * 4 * (e2e_steps + 1) // e2e_steps,
           "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3, "spp_per_s": e2e_steps * world / dt,
           "boundary": "render_b200_pass per progression: path-index range in, the progressive HOST framebuffer (W*H*3 f32, pinned) "
                       "out after every progression (N > 1: the reduced image on rank 0); render_b200_finish (flush + finished image) "
                       "once at the end, inside the timing"}
    if world == 1:
        e2e["accel_h"] = accel_boundary(cb, lib, acc, scene, torch)
    red.host_mirror = None

    # ---- image-level checks, outside every timed region
    multi_gpu_image = None
    gpu_parity_img = None
    if world > 1:
        # the N-GPU image of progressions 0..N-1 (one per rank, reduced) against rank 0 rendering the same N progressions alone:
        # every draw is a function of (frame, path index), so the two must agree up to fp32 summation order
        r.clear()
        red.clear()
        r.set_framebuffer(red.acquire(0).data_ptr())
        r.render_pass(rank * n_pass, n_pass, st)
        red.submit(0)
        total = red.finish()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            img_n = total.clone()
            r.clear()
            r.set_framebuffer(0)
            for g in range(world):
                r.render_pass(g * n_pass, n_pass, st)
            img_1 = torch.from_numpy(r.framebuffer()).cuda()
            d = (img_n - img_1).double()
            scale = float(img_1.double().mean())
            multi_gpu_image = {"progressions": world, "rel_rmse": float(d.pow(2).mean().sqrt()) / scale, "max_abs_diff_over_mean": float(d.abs().max()) / scale,
                               "ok": bool(float(d.pow(2).mean().sqrt()) / scale < 1e-5),
                               "what": f"sum over {world} ranks of progressions 0..{world - 1} vs the same progressions rendered by rank 0 alone"}
        dist.barrier()
    elif not args.no_cpu_baseline:
        r.clear()
        r.set_framebuffer(0)
        for _ in range(PARITY_SPP):
            r.render_pass(None, n_pass, st)
        gpu_parity_img = r.image(spp=PARITY_SPP)
    r.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm, which = peaks()
    ms_closest, ms_shadow = stt["ms"][1], stt["ms"][3]
    achieved = stt["rays_closest"] * bytes_per_ray / (ms_closest * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic()
    n_launch = max(1, stt["launches_closest"]) if "launches_closest" in stt else None
    total_rays = rays_closest + rays_shadow
    out = {
        "metric": METRIC, "value": total_rays / (ms_total * 1e-3), "unit": "rays/s",
        "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "num_tris": scene.num_prims, "frame": [WIDTH, HEIGHT], "sampler": "ptdl", "pointsampler": "rand",
                   "paths_per_step_per_gpu": n_pass, "rays_per_path": total_rays / paths,
                   "l2": "inputs larger than L2 (scene 0.8 GB, path pool 17 GB, ray/hit streams 1.2 GB per wave)",
                   "parallelism": f"spp split over {world} GPU(s), one framebuffer reduce per progression" if world > 1 else "1 GPU",
                   "bvh": {"nodes": acc.num_nodes(), "depth": acc.depth(), "node_bytes": node_b, "prim_bytes": prim_b, "gpu_build_s": build_s}},
        "spp_per_s": args.steps * world / (ms_total * 1e-3), "paths_per_s": paths / (ms_total * 1e-3),
        "kernel_ms_per_step_rank0": {k: v / args.steps for k, v in zip(["path_start", "closest_hit", "shade", "shadow", "nee_resolve"], stt["ms"])},
        "grays_per_s_rank0": {"closest_hit": stt["rays_closest"] / (ms_closest * 1e-3) / 1e9, "shadow": stt["rays_shadow"] / max(ms_shadow * 1e-3, 1e-12) / 1e9},
        "roofline": {"bound": "hbm", "kernel": "k_intersect (closest-hit traversal, all launches of the timed region, rank 0)",
                     "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "peak_source": which, "traffic": traffic,
                     "binding_limit": "L1 LSU wavefronts (79 %) + issue slots (66 %) at ~18 of 32 lanes per instruction (ncu: " + NCU_SUMMARY + "), not HBM: "
                                      "`bound`/`frac` follow SURVEY 8(d)'s algorithmic-byte definition (bytes a ray NEEDS from the tree / time), "
                                      "`traffic` shows that almost all of them are served by L1/L2 (which is also why `frac` can approach or pass 1)",
                     "traffic_source": {"kind": "committed ncu --set full capture, not measured in this run", "file": traffic_src,
                                        "command": NCU_COMMAND} if traffic else None,
                     "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one ~18 M-ray launch (a streamed wave): the ray and hit streams only" if traffic else None,
                     "algorithmic_bytes_per_launch": bytes_per_ray * NCU_RAYS_PER_LAUNCH,
                     "bytes_per_ray": bytes_per_ray, "bytes_per_ray_formula": f"{n_node:.3f} nodes x {node_b} B + {n_prim:.3f} prims x {prim_b} B + 40 B ray + 24 B hit",
                     "nodes_per_ray": n_node, "prims_per_ray": n_prim,
                     "counters_definition": "ACCEL_DEBUG (qbvhmp.c:83-90): node tests with >= 1 child hit, primitive tests, per closest-hit ray; "
                                            "counted by the instrumented kernel on one untimed progression of this workload",
                     "rays": stt["rays_closest"], "kernel_ms": ms_closest},
        "e2e": e2e,
        "gpu_launches": int(launches_all),
        "image_mean_xyz": image_mean,
        "clocks": clocks,
    }
    if multi_gpu_image is not None:
        out["multi_gpu_image_check"] = multi_gpu_image
    if world == 1 and not args.no_cpu_baseline:
        try:
            out["cpu_baseline"], ref_a, ref_b = cpu_baseline(cb, scene, lines, shape_mats, cam)
            out["parity_vs_reference"] = image_parity(gpu_parity_img, ref_a, ref_b)
            out["roofline"]["reference_tree"] = out["cpu_baseline"].pop("accel_debug")
            if os.path.exists(REF_ARM_IMAGE):      # left by `bench.py --impl reference` on this box: its image at its own sample count
                z = np.load(REF_ARM_IMAGE)
                if int(z["tris"]) == scene.num_prims:
                    out["parity_vs_reference"]["reference_arm"] = {
                        "spp": int(z["spp"]), "channel_mean_ratio_gpu_over_reference_arm": (np.array(image_mean) / z["means"]).tolist(),
                        "what": f"channel means of the {args.steps}-spp image of the timed region over those of the reference arm's {int(z['spp'])}-spp image"}
        except Exception as e:   # reference not built on this box: say so instead of inventing a number
            out["cpu_baseline"] = {"value": None, "unit": "rays/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(out), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.destroy_process_group()


def accel_boundary(cb, lib, acc, scene, torch):
    """the accel.h boundary for the same scene: pinned host ray batches in, hit records out (cb200_accel_intersect_n)"""
    S = cb.scenes
    n = 1 << 22
    rays = S.camera_rays(n, scene, seed=100)
    pr = torch.from_numpy(rays.view("u1").reshape(-1).copy()).pin_memory()
    po = torch.empty(n * 24, dtype=torch.uint8).pin_memory()
    L = lib.load()
    lib._check(L.cb200_accel_intersect_n(acc.a, pr.data_ptr(), None, po.data_ptr(), n), "cb200_accel_intersect_n")
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        lib._check(L.cb200_accel_intersect_n(acc.a, pr.data_ptr(), None, po.data_ptr(), n), "cb200_accel_intersect_n")
    dt = time.perf_counter() - t0
    return {"value": n * reps / dt, "unit": "rays/s", "h2d_bytes_per_step": n * 40, "d2h_bytes_per_step": n * 24,
            "boundary": "accel_intersect_n: 4 Mi random-pixel camera rays per call from pinned host memory"}


if __name__ == "__main__":
    main()
