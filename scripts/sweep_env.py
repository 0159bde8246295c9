"""scratch: A/B sweeps of the traversal kernels' tuning knobs on the bench scene (GPU box).

    python scripts/sweep_env.py CB200_POP_THRESHOLD=1,2,4,8 [CB200_PRIM_THRESHOLD=-12,-16] ...

Runs scripts/gpu_render_bench.py (8 streamed 4K ptdl progressions on the 10 M-triangle scene, CUDA-event time per kernel class)
once per combination in a fresh process (the knobs are read once per process) and prints one line each."""
import itertools, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
axes = []
for a in sys.argv[1:]:
    k, v = a.split("=")
    axes.append([(k, x) for x in v.split(",")])
for combo in itertools.product(*axes):
    env = dict(os.environ)
    env.update(dict(combo))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "gpu_render_bench.py")], env=env, capture_output=True, text=True)
    line = [l for l in p.stdout.split("\n") if l.startswith("batch")]
    print(" ".join(f"{k}={v}" for k, v in combo), "->", line[0] if line else ("FAILED " + p.stderr[-300:]), flush=True)
