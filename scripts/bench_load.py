"""scratch: the step BEFORE the path: 10 M-triangle bench scene from its .geo / .nra2 / .cam files to a built
accelerator, corona_b200 (mmap + upload + GPU LBVH) vs the unmodified reference (mmap + binned-SAH build on all host threads)."""
import importlib, os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
cb = importlib.import_module("corona-13_b200")
tris = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
scene, ms, cam, lines, shape_mats = bench.bench_scene(cb, tris)
t = time.time()
ref = bench.ReferenceRenderer(cb, scene, lines, shape_mats, cam)
print(f"wrote scene files in {time.time() - t:.2f} s:", {f: os.path.getsize(os.path.join(ref.tmp, f)) >> 20 for f in os.listdir(ref.tmp)}, "MiB")
REF = os.path.join(ROOT, "oracle", "_ref")
for rep in range(2):
    t = time.time()
    p = subprocess.run([os.path.join(ROOT, "corona-13_b200", "corona_b200"), ref.nra2, "-s", "1", "-w", "256", "-h", "128", "--coeff",
                        os.path.join(REF, "data", "ergb2spec.coeff")], capture_output=True, text=True, env=dict(os.environ, CB200_TIMING="1"))
    print(p.stderr.strip()[-600:])
    wall = time.time() - t
    print("corona_b200:", re.findall(r"\[main\].*took.*", p.stdout), re.findall(r".*build.*", p.stdout)[:3], f"process wall {wall:.2f} s (incl. CUDA context, 1 spp at 256x128)")
t = time.time()
q = subprocess.run([os.path.join(REF, "corona_ptdl_rand"), ref.nra2, "-x", "-s", "1", "-w", "256", "-h", "128", "-b", "0", "-t", str(os.cpu_count()), "--frame", "1"],
                   cwd=REF, capture_output=True, text=True)
print("reference:", re.findall(r".*construction took.*", q.stdout), f"process wall {time.time() - t:.2f} s")
ref.close()
