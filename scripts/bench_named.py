"""regression/0010_pt (pt) and 0011_ptdl (ptdl) -- and, named on the command line, the media fixture scenes -- at the reference's own arguments (-s 128 -w 1024 -h 576): the corona_b200 command
line against the unmodified reference binary on the box's host cores, same scene files (written from the golden fixture).
Prints one JSON line per sampler; numbers go to profiles/README.md."""
import json, os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import GoldenImage
REF = os.path.join(ROOT, "oracle", "_ref")
W, H, SPP = 1024, 576, int(sys.argv[1]) if len(sys.argv) > 1 else 128
NAMES = {"c10": "regression/0010_pt geometry + regenerated fill light", "fog": "fixture scene `fog` (exterior scattering medium)",
         "subsurf": "fixture scene `subsurf` (media behind dielectric interfaces)"}
RUNS = [(c, s) for c in (sys.argv[2:] or ["c10"]) for s in (("pt", "ptdl") if c == "c10" else ("ptdl",))]
for case, sampler in RUNS:
    g = GoldenImage(case)
    tmp = tempfile.mkdtemp()
    nra2 = g.write_files(tmp)
    cmd_warm = [os.path.join(ROOT, "corona-13_b200", "corona_b200"), nra2, "-s", "32", "-w", str(W), "-h", str(H), "--frame", "1", "--sampler", sampler, "--points", "rand",
                "--coeff", os.path.join(REF, "data", "ergb2spec.coeff"), "--tables", os.path.join(ROOT, "corona-13_b200", "data", "ref_tables.cbt"), "-q"]
    subprocess.run(cmd_warm, capture_output=True, text=True)     # first process on a fresh box: page cache, clocks, driver start-up (measured -25 %)
    t0 = time.time()
    p = subprocess.run([os.path.join(ROOT, "corona-13_b200", "corona_b200"), nra2, "-s", str(SPP), "-w", str(W), "-h", str(H), "--frame", "1", "--batch", "16",
                        "--sampler", sampler, "--points", "rand", "--coeff", os.path.join(REF, "data", "ergb2spec.coeff"),
                        "--tables", os.path.join(ROOT, "corona-13_b200", "data", "ref_tables.cbt")], capture_output=True, text=True)
    gpu_wall = time.time() - t0
    gpu_frame = float(re.findall(r"average of ([0-9.]+) s/frame", p.stdout)[0])
    cores = os.cpu_count()
    q = subprocess.run([os.path.join(REF, f"corona_{sampler}_rand"), nra2, "-x", "-s", str(max(SPP // 16, 2)), "-w", str(W), "-h", str(H), "-b", "0", "-t", str(cores),
                        "--frame", "1"], cwd=REF, capture_output=True, text=True)
    frames = [float(x) for x in re.findall(r"([0-9.]+) s/frame, \d+ spp", q.stdout)]
    ref_frame = sum(frames[1:]) / max(1, len(frames) - 1)
    print(json.dumps({"scene": NAMES.get(case, case), "sampler": sampler, "frame": [W, H], "spp": SPP,
                      "b200_s_per_spp": gpu_frame, "b200_spp_per_s": 1 / gpu_frame, "b200_paths_per_s": W * H / gpu_frame, "b200_wall_s_incl_load": gpu_wall,
                      "reference_s_per_spp": ref_frame, "reference_spp_per_s": 1 / ref_frame, "reference_cores": cores, "speedup": ref_frame / gpu_frame}))
