"""Golden IMAGES from the unmodified reference renderer (oracle/_ref/corona_*, built by oracle/Makefile from the
reference sources in place).  Run in the build container only:

    python tests/golden/make_golden_images.py [case ...]

For each case: writes the scene (.geo / .nra2 / .cam) to a scratch directory, runs the reference with two different
--frame seeds per integrator variant, and stores the images (fb*gain as the reference exports them) plus the complete
scene description (geometry, flattened materials, lookup tables, camera) in tests/golden/img_<case>.npz.  The second seed
gives the Monte Carlo noise floor the GPU image is judged against; with the Halton point sampler and the same --frame the
GPU integrator draws the SAME sample points as the reference, so those images agree far below the noise floor.

Variant keys name the reference binary: <sampler>_<pointsampler>[_rec709]  (oracle/Makefile `render` targets).
"""
import importlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
cb = importlib.import_module("corona-13_b200")
S, IO = cb.scenes, cb.scene_io
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def run_reference(binary, nra2, w, h, spp, frame, threads=None):
    """corona <scene> -x -s spp -w w -h h -b 0 --frame f, cwd = oracle/_ref (data/ergb2spec.coeff is cwd-relative, main.c:292)"""
    threads = threads or os.cpu_count()
    cmd = [os.path.join(REFDIR, binary), nra2, "-x", "-s", str(spp), "-w", str(w), "-h", str(h), "-b", "0",
           "-t", str(threads), "--frame", str(frame), "-q"]
    subprocess.run(cmd, cwd=REFDIR, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    base = os.path.splitext(nra2)[0]
    return IO.read_pfm(base + "render_fb00.pfm")


def load_tables():
    z = np.load(os.path.join(HERE, "ref_tables.npz"))
    return z["checker"], {k[6:]: z[k] for k in z.files if k.startswith("metal_")}


def golden_case(name, scene, shader_lines, shape_mats, cam, w, h, spp, variants, sky="black", env_pixels=None):
    tmp = tempfile.mkdtemp(prefix="corona_golden_")
    try:
        if env_pixels is not None:    # `sky_envmap <file> ...`: the map lives beside the scene (fb_map falls back to rt.searchpath)
            IO.write_fb(os.path.join(tmp, sky.split()[1]), env_pixels)
        shapes = []
        for i, sh in enumerate(scene.shapes):
            sh.write_geo(os.path.join(tmp, f"shape{i}.geo"))
            shapes.append((shape_mats[i], f"shape{i}"))
        nra2 = os.path.join(tmp, "test.nra2")
        IO.write_nra2(nra2, shader_lines, shapes, sky=sky)
        cam.write(os.path.join(tmp, "test01.cam"))
        out = {}
        for key in variants:
            for seed in (1, 2):
                img = run_reference("corona_" + key, nra2, w, h, spp, seed)
                out[f"{key}_seed{seed}"] = img
                print(name, key, seed, img.shape, img.mean(axis=(0, 1)), flush=True)
        checker, metals = load_tables()
        ms, _, _ = IO.parse_nra2(nra2, IO.Rgb2Spec(IO.coeff_path(ROOT)), checker, metals)
        mats, _ = ms.carrays()
        pack = {"materials": np.frombuffer(bytes(mats), np.uint8),   # flattened cb_material_t[] incl. the rgb2spec coefficients
                "num_shapes": np.int64(len(scene.shapes)), "shape_mats": np.int64(shape_mats), "w": np.int64(w), "h": np.int64(h),
                "spp": np.int64(spp), "shader_lines": np.array(shader_lines), "variants": np.array(list(variants)),
                "cam": np.frombuffer(open(os.path.join(tmp, "test01.cam"), "rb").read(), np.uint8),
                "num_tables": np.int64(len(ms.tables)), "sky": np.array(sky)}
        if ms.media:
            pack["media"] = np.frombuffer(bytes(ms.cmedia()), np.uint8)   # flattened cb_medium_t[]
            pack["exterior_medium"] = np.int64(ms.exterior_medium)
        if env_pixels is not None:
            e = IO.envmap_params(sky[len("sky_envmap"):], tmp)
            pack["env_pixels"], pack["env_mul"], pack["env_world"], pack["env_world_inv"] = e["pixels"], np.float32(e["mul"]), e["world"], e["world_inv"]
        if sky.startswith("sky_const"):
            co, sc = IO.sky_const_params(IO.Rgb2Spec(IO.coeff_path(ROOT)), sky[len("sky_const"):])
            pack["sky_coeff"], pack["sky_scale"] = np.float32(co), np.float32(sc)
        for i, (lmin, step, data) in enumerate(ms.tables):
            pack[f"tab{i}_meta"] = np.float32([lmin, step])
            pack[f"tab{i}_data"] = data
        for i, s in enumerate(scene.shapes):
            pack[f"s{i}_primid"] = s.primid
            pack[f"s{i}_vtxidx"] = s.vtxidx.view("<u4").reshape(-1, 2)
            pack[f"s{i}_vtx"] = s.vtx.view("<u4").reshape(-1, 4)
        for k, v in out.items():
            pack[k] = v.astype(np.float32)
        path = os.path.join(HERE, f"img_{name}.npz")
        np.savez_compressed(path, **pack)
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def reference_shader_lines(nra2):
    """the shader list of one of the reference's own regression scenes, comments stripped"""
    lines = [l.split("#")[0].strip() for l in open(nra2).read().split("\n")]
    n = int(lines[1].split()[0])
    return lines[2:2 + n]


def case_diffuse_static():
    """diffuse terrain + icosphere soup under a quad light, triangles only"""
    sc = S.synthetic_scene(3000, seed=7)
    lines = ["diffuse", "color d 0.6 0.5 0.4", "mult 1 1 0", "color d 0 0 0", "color e 30 30 30 1.", "mult 2 3 4 0"]
    cam = IO.Camera(pos=(14.0, 11.0, 9.0), lookat=(0.0, 0.0, 2.0), aperture_value=6, exposure_value=13, focal_length=0.35, iso=400.0)
    golden_case("diffuse_static", sc, lines, [2, 2, 5], cam, 160, 96, 256, ["pt_rand", "ptdl_rand", "ptdl_halton"])


def case_c10():
    """regression/0010_pt and 0011_ptdl as shipped offline: the six in-tree .geo files (4105 quads, analytic sphere /
    cylinder / cone with the rough dispersive dielectric, colour checker plane), the scene's own shader list and camera;
    the missing filllight.geo is regenerated as one emissive quad"""
    geo = os.path.join(REFDIR, "scenes", "geo")
    sc = S.c10_like_scene(geo)
    assert sc.name == "0010_pt", "needs the reference .geo files (make -C oracle ref)"
    order = [("plane", 2), ("emitter", 5), ("cone", 10), ("sphere", 10), ("cylinder", 10), ("cylinder_cap", 2)]   # test.nra2:17-22
    by_name = {s.name.replace(".geo", ""): s for s in sc.shapes}
    shapes = [by_name[n] for n, _ in order] + [S.quad_light((4.0, 0.0, 8.0), 1.5, 12)]
    mats = [m for _, m in order] + [12]
    lines = reference_shader_lines(os.path.join(REFDIR, "scenes", "0010_pt", "test.nra2"))
    cam = IO.read_cam(os.path.join(REFDIR, "scenes", "0010_pt", "test01.cam"))
    golden_case("c10", S.Scene(shapes, "0010_pt"), lines, mats, cam, 256, 160, 64, ["pt_rand", "ptdl_rand", "pt_halton", "ptdl_halton"])


def case_motion():
    """regression/0002_mb's ingredients (its geometry is not available offline): ptdl + halton + rec709 framebuffer, motion-blurred
    triangle soup, a moving analytic sphere and a translating camera, 1/30 s exposure so that time covers [0,1)"""
    terrain = S.terrain(1200, 3, quads=True, material=0)
    soup = S.soup(1500, 4, material=0, motion=(0.9, 0.3, -0.4))
    ball = S.analytic_shape("sphere", (2.0, 1.0, 4.5), 1.2, material=0, motion=(-1.0, 0.8, 0.3))
    light = S.quad_light((0.0, 0.0, 9.0), 2.5, 1)
    lines = ["diffuse", "color d 0.7 0.7 0.7", "mult 1 1 0", "color d 0 0 0", "color e 40 36 30 1.", "mult 2 3 4 0",
             "color d 0.8 0.2 0.1", "mult 1 6 0"]
    cam = IO.Camera(pos=(13.0, -12.0, 10.0), lookat=(0.0, 0.0, 3.0), aperture_value=5, exposure_value=11, focal_length=0.35,
                    iso=100.0, pos_t1=(13.4, -11.7, 10.1))
    golden_case("motion", S.Scene([terrain, soup, ball, light], "motion"), lines, [2, 7, 7, 5], cam, 160, 96, 128,
                ["ptdl_halton_rec709", "ptdl_rand", "pt_halton"])


def case_glass_metal():
    """smooth and rough dielectrics (mesh with shading normals + analytic sphere: nested media, specular chains, dispersion),
    rough gold, mirror-like silver, colour-checker floor with uvs"""
    g = 8
    xs = np.linspace(-8, 8, g + 1, dtype=np.float32)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    pos = np.stack([X, Y, np.zeros_like(X)], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    v00 = (i * (g + 1) + j).reshape(-1)
    uv = np.stack([(X / 16 + 0.5).reshape(-1), (Y / 16 + 0.5).reshape(-1)], -1) * 0.999 + 0.0005
    floor = S.mesh_shape(pos, np.stack([v00, v00 + g + 1, v00 + g + 2, v00 + 1], -1), 0, None, "floor", uv=uv)
    iv, itri = S._icosahedron()
    ctr = np.float32([[0.0, 0.0, 1.3], [-0.5, -4.5, 1.0], [3.0, -1.2, 1.5]])     # the third one intersects the analytic glass ball: nested media
    rad = np.float32([1.3, 1.0, 1.0])
    glass_mesh = S.mesh_shape((ctr[:, None, :] + rad[:, None, None] * iv[None]).reshape(-1, 3),
                              (itri[None] + 12 * np.arange(3)[:, None, None]).reshape(-1, 3), 0, None, "glass_mesh")
    glass_ball = S.analytic_shape("sphere", (3.0, -3.0, 1.5), 1.5, material=0)
    rough_ball = S.analytic_shape("sphere", (-3.5, 3.0, 1.2), 1.2, material=0)
    gold = S.analytic_shape("line", (-3.0, -3.5, 0.0), 0.9, (-3.0, -3.5, 2.5), 0.9, material=0)
    silver = S.analytic_shape("line", (3.5, 3.5, 0.0), 1.2, (3.5, 3.5, 2.8), 0.3, material=0)
    light = S.quad_light((0.0, 0.0, 8.0), 3.0, 1)
    lines = ["diffuse",                       # 0
             "colorcheckersg d",              # 1
             "mult 1 1 0",                    # 2 floor
             "color d 0 0 0",                 # 3
             "color e 25 25 25 1.",           # 4
             "mult 2 3 4 0",                  # 5 light
             "dielectric 1.5 40",             # 6
             "color g 1 1 1 0.0",             # 7
             "mult 1 7 6",                    # 8 smooth glass
             "color g 0.95 1 0.95 0.15",      # 9
             "dielectric 1.3 23",             # 10
             "mult 1 9 10",                   # 11 rough glass
             "metal Au",                      # 12
             "color g 1 1 1 0.2",             # 13
             "mult 1 13 12",                  # 14 rough gold
             "metal Ag",                      # 15
             "color g 1 1 1 0.0",             # 16
             "mult 1 16 15"]                  # 17 polished silver
    cam = IO.Camera(pos=(12.0, -9.0, 8.0), lookat=(0.0, 0.0, 1.2), aperture_value=6, exposure_value=13, focal_length=0.4, iso=400.0)
    golden_case("glass_metal", S.Scene([floor, glass_mesh, glass_ball, rough_ball, gold, silver, light], "glass_metal"), lines,
                [2, 8, 8, 11, 14, 17, 5], cam, 192, 128, 128, ["pt_halton", "ptdl_halton", "ptdl_rand"])


def synthetic_envmap(w=128, h=64):
    """a small HDR latitude-longitude map in sky_envmap's texel format (rgb2spec coefficients + scale): blue-white gradient above
    the horizon, a sun 500x brighter than the sky, brown ground; row 0 is the +z pole (theta = pi*y/h, sky_envmap.c:150-160)"""
    r2s = IO.Rgb2Spec(IO.coeff_path(ROOT))
    px = np.zeros((h, w, 4), np.float32)
    sun = np.float64([np.sin(1.0) * np.sin(0.6), np.sin(1.0) * np.cos(0.6), np.cos(1.0)])
    for j in range(h):
        th = np.pi * (j + 0.5) / h
        for i in range(w):
            ph = 2 * np.pi * (i + 0.5) / w - np.pi
            d = np.float64([np.sin(ph) * np.sin(th), np.cos(ph) * np.sin(th), np.cos(th)])
            if d[2] > 0:
                t = d[2]
                rgb = np.float64([0.35, 0.5, 0.9]) * t + np.float64([0.9, 0.9, 0.85]) * (1 - t)
                if d @ sun > 0.985:
                    rgb = rgb + np.float64([500.0, 450.0, 350.0])
            else:
                rgb = np.float64([0.25, 0.18, 0.1])
            mul, co = r2s.rgb_to_coeff(rgb.astype(np.float32))
            px[j, i, :3], px[j, i, 3] = co, mul
    return px


def case_envmap():
    """the environment-map sky module (src/shaders/sky_envmap.c): emission lookup, mip-hierarchy importance sampling for next event
    estimation and its pdf for MIS, brightness and a rotation about two axes; a geometric light beside it"""
    case_sky(True, "sky_envmap env.fb 40 0 25 30", "envmap", env_pixels=synthetic_envmap())


def case_sky(with_light, sky="cloudy", name=None, env_pixels=None):
    """the built-in `cloudy' sky (src/shader.c:268-334): environment vertices, sky next-event estimation and its MIS; with and
    without a geometric light beside it (lights_pdf_type splits the next-event budget 50:50 then, list.c:76-88)"""
    terrain = S.terrain(1800, 5, material=0)
    soup = S.soup(800, 6, rmin=0.2, rmax=0.9, material=0)
    ball = S.analytic_shape("sphere", (1.5, -1.0, 4.0), 1.3, material=0)
    rod = S.analytic_shape("line", (-3.0, 2.0, 2.0), 0.7, (-3.0, 2.0, 5.5), 0.7, material=0)
    shapes, mats = [terrain, soup, ball, rod], [2, 7, 10, 13]
    if with_light:
        shapes.append(S.quad_light((0.0, 0.0, 9.0), 1.5, 1))
        mats.append(5)
    lines = ["diffuse", "color d 0.5 0.55 0.4", "mult 1 1 0", "color d 0 0 0", "color e 2500 2300 2000 1.", "mult 2 3 4 0",
             "color d 0.75 0.3 0.2", "mult 1 6 0", "dielectric 1.5 40", "color g 1 1 1 0.0", "mult 1 9 8",
             "metal Cu", "color g 1 1 1 0.25", "mult 1 12 11"]
    cam = IO.Camera(pos=(13.0, 10.0, 8.0), lookat=(0.0, 0.0, 3.0), aperture_value=7, exposure_value=14, focal_length=0.35, iso=100.0)
    golden_case(name or ("sky_light" if with_light else "sky"), S.Scene(shapes, "sky"), lines, mats, cam, 160, 96, 128,
                ["ptdl_halton", "pt_halton", "ptdl_rand"], sky=sky, env_pixels=env_pixels)


def case_fog():
    """the whole scene inside a forward-scattering homogeneous medium (`exterior`, medium_rgb + color v): free-flight sampling on
    every edge, volume vertices with the Henyey-Greenstein phase function, transmittance on the next-event edges"""
    terrain = S.terrain(1800, 5, material=0)
    soup = S.soup(600, 6, rmin=0.3, rmax=1.0, material=0)
    ball = S.analytic_shape("sphere", (1.5, -1.0, 4.0), 1.3, material=0)
    light = S.quad_light((0.0, 0.0, 9.0), 1.2, 1)
    lines = ["diffuse", "color d 0.5 0.55 0.4", "mult 1 1 0", "color d 0 0 0", "color e 6000 5500 5000 1.", "mult 2 3 4 0",
             "color d 0.75 0.3 0.2", "mult 1 6 0",
             "medium_rgb 9 12 16 0.6",        # 8 mean free paths per colour, mean cosine
             "color v 0.95 0.95 0.9",         # 9 albedo
             "mult 1 9 8",                    # 10 fog
             "exterior 10 0"]                 # 11
    cam = IO.Camera(pos=(13.0, 10.0, 8.0), lookat=(0.0, 0.0, 3.0), aperture_value=7, exposure_value=14, focal_length=0.35, iso=100.0)
    golden_case("fog", S.Scene([terrain, soup, ball, light], "fog"), lines, [2, 7, 7, 5], cam, 160, 96, 128,
                ["ptdl_halton", "pt_halton", "ptdl_rand"])


def case_subsurf():
    """media behind dielectric interfaces (`interior`): a rough-glass ball and a smooth icosphere mesh filled with coloured
    scattering media, an overlapping pair (nested media: the smaller shape id wins, pathspace.c:107-113), an index-matched
    absorbing-only medium (no `color v`), all on a colour-checker floor"""
    g = 8
    xs = np.linspace(-8, 8, g + 1, dtype=np.float32)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    pos = np.stack([X, Y, np.zeros_like(X)], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    v00 = (i * (g + 1) + j).reshape(-1)
    uv = np.stack([(X / 16 + 0.5).reshape(-1), (Y / 16 + 0.5).reshape(-1)], -1) * 0.999 + 0.0005
    floor = S.mesh_shape(pos, np.stack([v00, v00 + g + 1, v00 + g + 2, v00 + 1], -1), 0, None, "floor", uv=uv)
    iv, itri = S._icosahedron()
    ctr = np.float32([[0.0, 0.0, 1.3], [3.0, -1.6, 1.5]])       # the second one intersects the analytic ball below
    rad = np.float32([1.3, 1.1])
    milk_mesh = S.mesh_shape((ctr[:, None, :] + rad[:, None, None] * iv[None]).reshape(-1, 3),
                             (itri[None] + 12 * np.arange(2)[:, None, None]).reshape(-1, 3), 0, None, "milk_mesh")
    skin_ball = S.analytic_shape("sphere", (3.0, -3.0, 1.5), 1.5, material=0)
    ink_ball = S.analytic_shape("sphere", (-3.5, 3.0, 1.2), 1.2, material=0)
    light = S.quad_light((0.0, 0.0, 8.0), 3.0, 1)
    lines = ["diffuse",                       # 0
             "colorcheckersg d",              # 1
             "mult 1 1 0",                    # 2 floor
             "color d 0 0 0",                 # 3
             "color e 40 40 40 1.",           # 4
             "mult 2 3 4 0",                  # 5 light
             "dielectric 1.33 40",            # 6
             "color g 1 1 1 0.0",             # 7
             "mult 1 7 6",                    # 8 smooth interface
             "color g 1 1 1 0.15",            # 9
             "mult 1 9 6",                    # 10 rough interface
             "medium_rgb 0.5 0.35 0.25 0.0",  # 11 milk: isotropic
             "color v 0.98 0.96 0.9",         # 12
             "mult 1 12 11",                  # 13
             "medium_rgb 0.3 0.12 0.08 0.7",  # 14 skin: forward scattering
             "color v 0.95 0.8 0.7",          # 15
             "mult 1 15 14",                  # 16
             "medium_rgb 4 1.5 0.8 0.0",      # 17 ink: absorbs only (no colour v)
             "dielectric 1.0 0",              # 18 index-matched boundary
             "mult 1 7 18",                   # 19
             "interior 8 13",                 # 20 milk mesh
             "interior 10 16",                # 21 skin ball
             "interior 19 17"]                # 22 ink ball
    cam = IO.Camera(pos=(12.0, -9.0, 8.0), lookat=(0.0, 0.0, 1.2), aperture_value=6, exposure_value=13, focal_length=0.4, iso=400.0)
    golden_case("subsurf", S.Scene([floor, milk_mesh, skin_ball, ink_ball, light], "subsurf"), lines,
                [2, 20, 21, 22, 5], cam, 192, 128, 128, ["pt_halton", "ptdl_halton", "ptdl_rand"])


def _box(lo, hi, name):
    """axis-aligned box as 6 quads with their own corners (flat shading normals), wound counter-clockwise seen from outside"""
    lo, hi = np.float32(lo), np.float32(hi)
    c = np.float32([[x, y, z] for z in (lo[2], hi[2]) for y in (lo[1], hi[1]) for x in (lo[0], hi[0])])   # index = x + 2y + 4z
    quads = [[0, 2, 3, 1], [4, 5, 7, 6], [0, 1, 5, 4], [2, 6, 7, 3], [0, 4, 6, 2], [1, 3, 7, 5]]      # -z +z -y +y -x +x
    pos = np.concatenate([c[q] for q in quads])
    return S.mesh_shape(pos, np.arange(24).reshape(6, 4), 0, None, name)


def case_vstack():
    """regression/0090_vstack ("volume stack priorities") as shipped: its shader list, its camera and its first line
    `const 1 1 1 2000` -- not a sky module of the reference, so dlopen fails there and the default cloudy sky stays
    (src/shader.c:612-614,643-675).  Two overlapping dielectric cubes filled with absorbing media (interior + medium_rgb behind
    a black `color v`), the smaller shape id wins inside the overlap (pathspace.c:107-113).  The scene's geometry is not
    available offline: plane, cubes and emitter are regenerated in the places the camera looks at."""
    ref = os.path.join(REFDIR, "scenes", "0090_vstack")
    lines = reference_shader_lines(os.path.join(ref, "test.nra2"))
    cam = IO.read_cam(os.path.join(ref, "test01.cam"))
    sky = open(os.path.join(ref, "test.nra2")).read().split("\n")[0].split("#")[0].strip()
    g = 6
    xs = np.linspace(-9, 9, g + 1, dtype=np.float32)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    pos = np.stack([X, Y, np.zeros_like(X)], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    v00 = (i * (g + 1) + j).reshape(-1)
    plane = S.mesh_shape(pos, np.stack([v00, v00 + g + 1, v00 + g + 2, v00 + 1], -1), 0, None, "plane")
    cube2 = _box((-0.2, -1.0, 0.02), (2.2, 1.4, 2.4), "cube2")
    cube = _box((-2.0, -1.4, 0.02), (0.6, 1.0, 2.0), "cube")
    emitter = S.quad_light((0.5, -1.0, 7.0), 0.6, 1)
    golden_case("vstack", S.Scene([plane, cube2, cube, emitter], "vstack"), lines, [2, 17, 16, 5], cam, 192, 128, 128,
                ["pt_halton", "ptdl_halton", "ptdl_rand"], sky=sky)


def case_skin():
    """regression/0030_subsurf as shipped: its shader list (`diffdiel 1.33 30` under a rough glossy layer, `interior` with a
    strongly scattering medium_rgb + `color v`), its camera and sky; the scene's geometry (skincube, sphere, plane, emitter) is not
    available offline and is regenerated where the camera looks: a box and an analytic sphere of the skin material on the plane"""
    ref = os.path.join(REFDIR, "scenes", "0030_subsurf")
    lines = reference_shader_lines(os.path.join(ref, "test.nra2"))
    cam = IO.read_cam(os.path.join(ref, "test01.cam"))
    c, look = np.float64(cam.pos), None
    print("0030_subsurf camera", cam.pos, cam.focus)
    g = 6
    xs = np.linspace(-9, 9, g + 1, dtype=np.float32)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    pos = np.stack([X, Y, np.zeros_like(X)], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    v00 = (i * (g + 1) + j).reshape(-1)
    plane = S.mesh_shape(pos, np.stack([v00, v00 + g + 1, v00 + g + 2, v00 + 1], -1), 0, None, "plane")
    cube = _box((-1.8, -0.9, 0.02), (-0.2, 0.7, 1.6), "skincube")
    ball = S.analytic_shape("sphere", (1.1, 0.2, 0.9), 0.88, material=0)
    emitter = S.quad_light((0.0, -0.5, 4.0), 0.5, 1)
    if surface_only == "dielectric":   # debugging aid: same media behind a plain dielectric
        lines = [l.replace("diffdiel", "dielectric") for l in lines]
        golden_case("skin_dielectric", S.Scene([emitter, plane, cube, ball], "skin"), lines, [5, 2, 12, 12], cam, 192, 128, 64, ["ptdl_halton"])
        return
    golden_case("skin", S.Scene([emitter, plane, cube, ball], "skin"), lines, [5, 2, 12, 12], cam, 192, 128, 128,
                ["ptdl_halton", "pt_halton", "ptdl_rand"])


def case_furnace():
    """white furnace: the skin scene's box and ball alone under a constant sky, index-matched interface, non-absorbing dense medium
    (mean free paths 0.003 - 0.014).  Physically every pixel would show the sky's radiance minus what the 32-vertex cap
    truncates (~12 %); the reference shows about HALF of it, because its sampler_mis() multiplies the path pdf up in double
    precision and converts to float: vertex pdfs of 1e7 and more overflow 3.4e38 after a handful of scattering events and
    those samples are dropped (pt.c:30-38, ptdl.c:78-88, view.c:457-458).  This fixture pins that behaviour."""
    ref = os.path.join(REFDIR, "scenes", "0030_subsurf")
    lines = reference_shader_lines(os.path.join(ref, "test.nra2"))
    lines = [l.replace("diffdiel 1.33 30", "dielectric 1.0 0").replace("color v 0.99 0.91 0.85", "color v 1 1 1") for l in lines]
    cam = IO.read_cam(os.path.join(ref, "test01.cam"))
    cube = _box((-1.8, -0.9, 0.02), (-0.2, 0.7, 1.6), "skincube")
    ball = S.analytic_shape("sphere", (1.1, 0.2, 0.9), 0.88, material=0)
    golden_case("furnace", S.Scene([cube, ball], "furnace"), lines, [12, 12], cam, 192, 128, 64, ["pt_halton", "ptdl_halton"],
                sky="sky_const 1 1 1 1")


def case_sphere_light():
    """emissive analytic spheres (one of them moving) beside a quad light: next-event estimation picks primitives by area x
    radiance over all three (lights/list.c:56-104), samples sphere points with prims_sample's sphere branch (prims.c:225-230,
    sphere.h:38-49), and the shadow ray ends on the sphere's own near side (path_visible, pathspace.c:311-344)"""
    terrain = S.terrain(1800, 11, material=0)
    soup = S.soup(900, 12, rmin=0.2, rmax=0.8, material=0)
    ball = S.analytic_shape("sphere", (-2.5, 1.5, 3.2), 1.0, material=0)
    lamp0 = S.analytic_shape("sphere", (2.0, -1.5, 5.5), 0.6, material=0)
    lamp1 = S.analytic_shape("sphere", (-4.0, -3.0, 4.0), 0.35, material=0, motion=(0.8, 0.4, 0.3))
    light = S.quad_light((3.0, 4.0, 9.0), 1.0, 1)
    lines = ["diffuse", "color d 0.6 0.6 0.55", "mult 1 1 0", "color d 0 0 0", "color e 60 50 40 1.", "mult 2 3 4 0",
             "color d 0.2 0.5 0.8", "mult 1 6 0", "color e 90 120 160 1.", "mult 2 3 8 0", "dielectric 1.5 40", "color g 1 1 1 0.0", "mult 1 11 10"]
    cam = IO.Camera(pos=(13.0, -11.0, 9.0), lookat=(0.0, 0.0, 3.0), aperture_value=6, exposure_value=11, focal_length=0.35, iso=100.0)
    golden_case("sphere_light", S.Scene([terrain, soup, ball, lamp0, lamp1, light], "sphere_light"), lines, [2, 7, 12, 5, 9, 5], cam, 160, 96, 128,
                ["ptdl_halton", "pt_halton", "ptdl_rand"])


CASES = {"sphere_light": case_sphere_light, "envmap": case_envmap, "furnace": case_furnace, "skin": case_skin, "vstack": case_vstack, "fog": case_fog, "subsurf": case_subsurf, "sky_const": lambda: case_sky(True, "sky_const 0.3 0.5 0.9 800", "sky_const"), "sky": lambda: case_sky(False), "sky_light": lambda: case_sky(True), "diffuse_static": case_diffuse_static, "c10": case_c10, "motion": case_motion, "glass_metal": case_glass_metal}

if __name__ == "__main__":
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()
