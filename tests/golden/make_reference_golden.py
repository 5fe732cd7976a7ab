"""Golden vectors produced by RUNNING THE REFERENCE'S OWN SOURCES (build container only).

    python tests/golden/make_reference_golden.py            # rewrites tests/golden/ref_*.json

The reference is Gauche Scheme and no Scheme runtime exists in the build image, so its .scm files
are executed, unmodified and from where they lie (/root/reference), by the minimal interpreter
oracle/minischeme.py (its header lists where Gauche behaviour had to be restated: numeric tower,
f64vector / array primitives, libm).  Everything written here is an OUTPUT OF REFERENCE CODE:

  ref_prims.json      closest hits of the reference's constructors (sphere, moving sphere, rects, flip, box,
                      translate / rotate-y instances, Bezier curve) called through `g:hit` on seeded rays
  ref_scenes.json     `g:hit` on the scenes main.scm itself defines (cornell-box, test-scene2, test-scene,
                      cornell-bezier), on rays made by the reference's own `cam:get-ray` and on seeded rays
  ref_textures.json   the Perlin tables perlin.scm builds at module load (from the scripted random-real) and
                      `noise`, `turb`, checker / noise / marble texture values
  ref_materials.json  `reflect`, `refract`, `schlick`, `(local uvw (random-cosine-direction))` (six draws: Q15), sky-color,
                      lambertian scatter / scattering-pdf, diffuse-light emitted
  ref_color.json      main.scm's `trace-all` (color, running sum, gamma, 8-bit) and `save-as-ppm` on cornell-box and test-scene2
                      with `random-real` scripted to return the oracle's Philox draws in the reference's call
                      order, so the oracle (and the CUDA path) must reproduce the radiance of every pixel

`random-real` is scripted by this file (srfi-27's MT19937 stream is not a parity target).  Rays and
primitive parameters are fp32-representable so that the CUDA path can be fed the very same inputs.
The .json files travel to the GPU box; this script and /root/reference do not need to.
"""
import json
import os
import sys
import threading
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REFERENCE = os.environ.get("SRT_REFERENCE", "/root/reference")

from oracle.minischeme import Interp, Sym, F64, Values, read_all   # noqa: E402
from tests import raybatch                                          # noqa: E402
from tests.refspec import build_host, host_scene                    # noqa: E402

MAXF = 999999999999


def v3(*a):
    return F64(float(x) for x in a)


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


class Ref:
    """The reference, loaded."""

    def __init__(self, random_real):
        self.it = Interp(REFERENCE, random_real)
        self.rng = random_real
        for mod in ["vec", "util", "ray", "onb", "perlin", "texture", "material", "geometry", "bezier", "camera", "pdf", "points", "constant"]:
            self.it.require(mod)
        self.lam = self.call("material", "make-lambertian", self.call("texture", "constant-texture", v3(0.5, 0.5, 0.5)))

    def call(self, module, name, *args):
        return self.it.call(module, name, *args)

    def build(self, spec):
        k, c = spec[0], self.call
        if k == "sphere":
            return c("geometry", "make-sphere", v3(*spec[1]), spec[2], self.lam)
        if k == "moving-sphere":
            return c("geometry", "make-moving-sphere", v3(*spec[1]), v3(*spec[2]), spec[3], spec[4], spec[5], self.lam)
        if k in ("xy-rect", "xz-rect", "yz-rect"):
            return c("geometry", "make-" + k, *spec[1:6], self.lam)
        if k == "flip":
            return c("geometry", "flip-normals", self.build(spec[1]))
        if k == "box":
            return c("geometry", "make-box", v3(*spec[1]), v3(*spec[2]), self.lam)
        if k == "translate":
            return c("geometry", "translate", self.build(spec[1]), v3(*spec[2]))
        if k == "rotate-y":
            return c("geometry", "rotate-y", self.build(spec[1]), spec[2])
        if k == "bezier":
            return c("bezier", "make-bezier", v3(*spec[1]), v3(*spec[2]), v3(*spec[3]), v3(*spec[4]), spec[5], self.lam)
        if k == "medium":
            return c("geometry", "make-constant-medium", self.build(spec[1]), spec[2], c("texture", "constant-texture", v3(1, 1, 1)))
        if k == "klein":
            return c("geometry", "make-klein", v3(*spec[1]), self.lam)
        raise ValueError(k)

    def hit(self, obj, ray7, t_min=0.001, t_max=MAXF):
        ray = self.call("ray", "make-ray-with-time", v3(*ray7[0:3]), v3(*ray7[3:6]), float(ray7[6]))
        r = self.call("geometry", "hit", obj, ray, t_min, t_max)
        if r[0] is False:
            return None
        rec = r[1]
        return dict(t=float(rec[0]), p=list(rec[1]), n=list(rec[2]), u=float(rec[4]), v=float(rec[5]), material=rec[3])

    def load_main(self, names):
        """Take the named top-level definitions (and the module header) out of main.scm."""
        forms = read_all(open(os.path.join(REFERENCE, "main.scm")).read())
        main = None
        for form in forms:
            if not isinstance(form, list) or not form:
                continue
            if form[0] == "define-module":
                self.it.eval(form, self.it.user)
                main = self.it.modules["main"]
            elif form[0] in ("define", "define-inline") and main is not None:
                name = form[1][0] if isinstance(form[1], list) else form[1]
                if name in names:
                    self.it.eval(form, main)
        return main


def tag_medium(ref, medium, leaf_id):
    """Instrumentation only: tells the scripted random-real WHICH constant medium is drawing its free-flight
    sample (geometry.scm:562), so that it can return the draw the oracle / the C-ABI's trace_batch assign to that
    leaf - Philox key (ray index, seed 0), counter (sample 0, bounce 1, block 16 + leaf id, 0), component 0 -
    independent of visit order; inside a path (make_color) the key is (pixel, seed) and the counter (sample, depth + 1,
    16 + leaf id, 0) of the path being traced.  The medium's own hit closure runs unchanged."""
    inner, it, rng = medium[0], ref.it, ref.rng
    if getattr(inner, "srt_tagged", False):
        return

    def hit_fn(ray, t_min, t_max):
        rng.medium_leaf = leaf_id
        try:
            return it.apply(inner, [ray, t_min, t_max])
        finally:
            rng.medium_leaf = None
    hit_fn.srt_tagged = True
    medium[0] = hit_fn


def hits_table(ref, obj, rays, leaf_materials=None):
    """`leaf_materials`: the material objects of the scene's leaves in depth-first list order, when every leaf has its OWN
    material object - the hit record's material (ray.scm:27) then identifies WHICH object the reference hit, and the table
    gets a `prim` column (the north star's "primitive ids must match the reference's intersection routines")."""
    hit, t, p, n, uv, prim = [], [], [], [], [], []
    for i, r in enumerate(rays):
        ref.rng.ray_index = i
        h = ref.hit(obj, r)
        if leaf_materials is not None:
            prim.append(-1 if h is None else [k for k, mat in enumerate(leaf_materials) if mat is h["material"]][0])
        hit.append(h is not None)
        t.append(h["t"] if h else 0.0)
        p.append(h["p"] if h else [0.0, 0.0, 0.0])
        n.append(h["n"] if h else [0.0, 0.0, 0.0])
        uv.append([h["u"], h["v"]] if h else [0.0, 0.0])
    out = dict(rays=[list(map(float, r)) for r in rays], hit=hit, t=t, p=p, n=n, uv=uv)
    if leaf_materials is not None:
        assert len({id(mat) for mat in leaf_materials}) == len(leaf_materials)
        out["prim"] = prim
    return out


def stability_mask(scene, rays):
    """Rays on which the CUDA path (fp32) may legitimately differ (used only by the GPU test), decided with
    the oracle: near-ties and fp32-flipped decisions exactly like tests/raybatch.compare, plus rays that are
    ill-conditioned for the 1e-4 bar in fp32 - the reference algorithm itself, evaluated in fp32 (f32 oracle),
    is already off by more than half the bar in t or normal (a small sphere far away: the rounding of t is
    divided by the radius in the normal)."""
    from oracle import oracle as O
    S = O.OracleScene(scene)
    r64 = np.asarray(rays, np.float64)
    o, o32 = S.trace_batch(r64), S.trace_batch(r64, precision=32)
    t2 = S.second_best_t(r64, o["prim"])
    near = (o["prim"] >= 0) & (np.abs(t2 - o["t"]) < 1e-5 * np.abs(o["t"])) & (t2 != o["t"])
    both = (o["prim"] >= 0) & (o32["prim"] == o["prim"])
    t_err = np.abs(o32["t"] - o["t"]) / np.maximum(np.abs(o["t"]), 1e-30)
    n_err = np.linalg.norm(o32["n"] - o["n"], axis=1) / np.maximum(np.linalg.norm(o["n"], axis=1), 1e-30)
    ptype = S.flat.prims["type"][S.flat.first_of_logical[np.maximum(o["prim"], 0)]]
    # (the Klein fractal's central-difference normal is held to 2e-2 on the GPU, see tests/test_gpu_parity.py: not an `ill` criterion)
    ill = both & ((t_err > 5e-5) | ((n_err > 5e-5) & (ptype != 8)))
    # edge rays: the f64 answer itself flips when one direction component moves by 2 fp32 ulps (a ray that runs
    # exactly into the floor / wall edge of the Cornell room hits at x = 0.0 in f64; any rounding puts it outside)
    edge = np.zeros(len(r64), bool)
    for axis in range(3):
        for sgn in (-1.0, 1.0):
            q = r64.copy()
            q[:, 3 + axis] *= 1.0 + sgn * 2.4e-7
            edge |= S.trace_batch(q)["prim"] != o["prim"]
    return [bool(x) for x in (near | (o32["prim"] != o["prim"]) | ill | edge)]


PRIM_CASES = [
    ("sphere", ["sphere", [0.5, -0.25, -1.0], 0.75]),
    ("sphere-negative-radius", ["sphere", [-1.0, 0.0, -1.0], -0.45]),
    ("sphere-huge", ["sphere", [0.0, -1000.0, 0.0], 1000.0]),
    ("moving-sphere", ["moving-sphere", [0.0, 0.25, 0.0], [0.0, 0.75, 0.5], 0.0, 1.0, 0.5]),
    ("xy-rect", ["xy-rect", 3.0, 5.0, 1.0, 3.0, -2.0]),
    ("xz-rect", ["xz-rect", 213.0, 343.0, 227.0, 332.0, 554.0]),
    ("yz-rect-flipped", ["flip", ["yz-rect", 0.0, 555.0, 0.0, 555.0, 555.0]]),
    ("box", ["box", [0.0, 0.0, 0.0], [165.0, 330.0, 165.0]]),
    ("box-rotated-translated", ["translate", ["rotate-y", ["box", [0.0, 0.0, 0.0], [165.0, 165.0, 165.0]], -18.0], [130.0, 0.0, 65.0]]),
    ("box-rotated-translated-2", ["translate", ["rotate-y", ["box", [0.0, 0.0, 0.0], [165.0, 330.0, 165.0]], 15.0], [265.0, 0.0, 295.0]]),
    ("bezier", ["bezier", [-1.0, 0.0, -1.0], [-0.8125, 1.0, 1.0], [0.8125, -1.0, 1.0], [1.0, 0.0, -1.0], 0.125]),
]
# second file (ref_prims2.json): the "next" rows of SURVEY 8(f).  The free-flight draw of a medium is scripted per
# (ray index, leaf), see tag_medium.
PRIM_CASES2 = [
    ("medium-sphere", ["medium", ["sphere", [0.0, 0.0, 0.0], 2.0], 0.75]),
    ("medium-box-instance", ["medium", ["translate", ["rotate-y", ["box", [0.0, 0.0, 0.0], [165.0, 165.0, 165.0]], -18.0], [130.0, 0.0, 65.0]], 0.015625]),
    ("klein", ["klein", [250.0, 200.0, 280.0]]),
]


INSTANCE_BOUNDS = {(130.0, 0.0, 65.0): ([78.0, -1.0, 64.0], [339.0, 166.0, 274.0]),
                   (265.0, 0.0, 295.0): ([264.0, -1.0, 251.0], [468.0, 331.0, 456.0])}


def curve_point(spec, s):
    a, b, c, d = (np.asarray(q, float) for q in spec[1:5])
    return a * (1 - s) ** 3 + 3 * b * (1 - s) ** 2 * s + 3 * c * (1 - s) * s ** 2 + d * s ** 3


def rays_for(spec, n, seed):
    """Half the batch: seeded rays through the object's extent (tests/raybatch.random_rays); the other half
    keeps those origins but aims at a point of the object's box (a Bezier curve: a point of the centre line
    displaced by up to 1.5 half-widths), so that thin objects are hit and grazed, not only missed."""
    from scheme_raytrace_b200.host import geometry as g, scenes
    from scheme_raytrace_b200.host.flatten import flatten_scene
    flat = flatten_scene(g.make_scene([build_host(spec)], scenes.default_camera(), scenes.sky_color))
    lo, hi = raybatch.interest_bounds(flat, clip=4.0) if spec[0] not in ("medium", "klein") else (None, None)
    if spec[0] == "medium" and spec[1][0] == "translate":
        lo, hi = (np.asarray(q, float) for q in INSTANCE_BOUNDS[tuple(spec[1][2])])
    if spec[0] == "medium" and spec[1][0] == "sphere":
        lo, hi = np.asarray(spec[1][1]) - spec[1][2] - 1.0, np.asarray(spec[1][1]) + spec[1][2] + 1.0
    if spec[0] == "klein":                           # the fractal lives inside radius ~ 125 + of its centre (geometry.scm:596)
        lo, hi = np.asarray(spec[1]) - 150.0, np.asarray(spec[1]) + 150.0
    if spec[0] == "translate":                       # instances: the extent of the transformed box, by hand (+-1 like interest_bounds)
        lo, hi = (np.asarray(q, float) for q in INSTANCE_BOUNDS[tuple(spec[2])])
    rays = raybatch.random_rays((lo, hi), n, seed).astype(np.float64)
    rs = np.random.RandomState(seed + 5000)
    for r in rays[n // 2:]:
        if spec[0] == "bezier":
            aim = curve_point(spec, rs.uniform(0.02, 0.98)) + rs.normal(size=3) * spec[5] * 0.5
        else:
            aim = (lo + 1.0) + rs.random_sample(3) * ((hi - 1.0) - (lo + 1.0))
        d = aim - r[0:3]
        r[3:6] = d / np.linalg.norm(d) * rs.uniform(0.5, 2.0)
    return f32(rays)


def make_prims(ref, case_list=None, seed0=100):
    from scheme_raytrace_b200.host import geometry as g, scenes
    cases = []
    for i, (name, spec) in enumerate(case_list or PRIM_CASES):
        n = 64 if spec[0] == "bezier" else (100 if spec[0] == "klein" else 200)
        rays = rays_for(spec, n, seed0 + i)
        obj = ref.build(spec)
        if spec[0] == "medium":
            tag_medium(ref, obj, 0)
        tab = hits_table(ref, obj, rays)
        tab["unstable"] = stability_mask(g.make_scene([build_host(spec)], scenes.default_camera(), scenes.sky_color), rays)
        cases.append(dict(name=name, spec=spec, **tab))
        print(f"prims/{name}: {sum(tab['hit'])}/{n} hits, {sum(tab['unstable'])} flagged")
    return dict(source="reference constructors of geometry.scm / bezier.scm through g:hit (oracle/minischeme.py)", t_min=0.001, t_max=MAXF, cases=cases)


MAIN_NAMES = {"random-scene", "save-as-ppm", "trace-line", "line-upped-spheres", "*spheres-list*", "*bvh-sah-node*", "*bvh-node*", "test-scene-non-bvh", "test-scene-bvh", "test-scene-bvh-sah",
              "test-bezier", "cornell-smoke", "klein-scene", "cornell-klein", "+max-depth+", "+black+", "+white+", "sky-color", "black", "color", "correct-gamma", "*size-x*", "*size-y*",
              "*cornell-camera*", "*camera*", "test-scene", "test-scene2", "cornell-box", "cornell-bezier", "trace-all",
              "*image*", "*raw-data*"}


SCENES1 = ["cornell-box", "test-scene2", "test-scene", "cornell-bezier"]
SCENES2 = ["cornell-smoke", "cornell-klein", "klein-scene", "test-bezier", "test-scene-non-bvh", "test-scene-bvh", "test-scene-bvh-sah", "random-scene"]
RANDOM_SCENE_SEED = 41


def reference_random_scene(ref, main):
    """main.scm:31-89 (random-scene), the generator behind cfg2 / cfg3.  At HEAD its last form calls (g:make-scene obj-list)
    with one argument although make-scene takes three (geometry.scm:52): executed as is, it raises - checked here.  For the
    golden the call is repeated with g:make-scene replaced, inside module main only, by a procedure that hands back the
    object list; random-real replays numpy's RandomState(RANDOM_SCENE_SEED) stream, which is what the host mirror
    `scenes.random_scene(seed, -5, 10, moving=True, checker_ground=True)` draws from."""
    from oracle.minischeme import SchemeError
    rs = np.random.RandomState(RANDOM_SCENE_SEED)
    ref.rng.gen = iter(lambda: float(rs.random_sample()), None)
    try:
        ref.it.call("main", "random-scene")
        raise AssertionError("random-scene was expected to fail on make-scene's arity")
    except SchemeError as e:
        assert "wrong number of arguments" in str(e), e
    rs = np.random.RandomState(RANDOM_SCENE_SEED)
    ref.rng.gen = iter(lambda: float(rs.random_sample()), None)
    main.vars[Sym("g:make-scene")] = lambda *a: a[0]
    try:
        obj_list = ref.it.call("main", "random-scene")
    finally:
        del main.vars[Sym("g:make-scene")]
        ref.rng.gen = None
    return ref.call("geometry", "make-scene", obj_list, main.lookup(Sym("*camera*")), main.lookup(Sym("sky-color")))


def material_row(ref, mat):
    """(kind, rgb of the texture at the origin or None, parameter) read out of a reference material vector:
    lambertian #(scatter spdf emitted albedo), metal #(scatter emitted albedo fuzz), dielectric #(scatter emitted ref-idx)."""
    if len(mat) == 3:
        return ["dielectric", None, float(mat[2])]
    if isinstance(mat[3], list):
        return ["lambertian", [float(x) for x in ref.call("texture", "value", mat[3], 0, 0, v3(0.05, 0.05, 0.05))], 0.0]
    return ["metal", [float(x) for x in ref.call("texture", "value", mat[2], 0, 0, v3(0.05, 0.05, 0.05))], float(mat[3])]


def make_scenes(ref, main, names=None, seed0=200):
    out = []
    for i, name in enumerate(names or SCENES1):
        scene = reference_random_scene(ref, main) if name == "random-scene" else main.lookup(Sym(name))
        if name == "cornell-smoke":                # list positions 6, 7 = leaves 6, 7 (the boundary boxes are not scene primitives)
            objs = list(ref.call("geometry", "scene-obj-list", scene))
            tag_medium(ref, objs[6], 6)
            tag_medium(ref, objs[7], 7)
        cam = ref.call("geometry", "scene-camera", scene)
        # (i) camera rays from the reference's own get-ray on a 12 x 12 grid of (s, t); random-real is scripted (0.5)
        # (the grid is offset off-centre: on a symmetric grid the diagonal rays of the Cornell camera run EXACTLY into
        # the edges of the room - hit at x = 0.0 in f64, a coin flip under any rounding)
        cam_rays, cam_st = [], []
        for a in range(12):
            for b in range(12):
                cam_st.append(((a + 0.37) / 12, (b + 0.61) / 12))
                ray = ref.call("camera", "get-ray", cam, *cam_st[-1])
                cam_rays.append(list(ray[0]) + list(ray[1]) + [float(ray[2])])
        cam_rays = f32(cam_rays)                       # rounded to fp32 AFTER the reference made them; stored as the inputs of g:hit
        # (ii) seeded rays through the scene's extent
        from scheme_raytrace_b200.host.flatten import flatten_scene
        hs = host_scene(name)
        nr = 60 if name in ("cornell-bezier", "test-bezier", "cornell-klein", "klein-scene") else 200
        rnd = raybatch.random_rays(raybatch.interest_bounds(flatten_scene(hs)), nr, seed0 + i).astype(np.float64)
        if name.startswith("test-scene-"):           # the 10 x 10 sphere grid: 150 more rays aimed at (or just past) individual spheres, so that many ids occur
            rs = np.random.RandomState(seed0 + 50 + i)
            o3 = np.stack([rs.uniform(-2, 11, 150), rs.uniform(0.2, 5, 150), rs.uniform(-2, 11, 150)], axis=1)
            aim = np.stack([rs.randint(0, 10, 150), np.zeros(150), rs.randint(0, 10, 150)], axis=1) + rs.normal(size=(150, 3)) * 0.3
            d3 = aim - o3
            d3 *= (rs.uniform(0.5, 2.0, 150) / np.linalg.norm(d3, axis=1))[:, None]
            rnd = np.concatenate([rnd, f32(np.concatenate([o3, d3, rs.random_sample((150, 1))], axis=1))])
        rays = np.concatenate([cam_rays[::3] if "klein" in name else cam_rays, rnd])
        cam_rays = cam_rays[::3] if "klein" in name else cam_rays
        cam_st = cam_st[::3] if "klein" in name else cam_st
        leaf_materials = None
        if name in ("test-scene", "test-scene2", "test-scene-non-bvh", "random-scene"):   # every object of the list has its own material object
            leaf_materials = [o[2] for o in ref.call("geometry", "scene-obj-list", scene)]
        elif name in ("test-scene-bvh", "test-scene-bvh-sah"):                      # ground + the BVH over *spheres-list* (depth-first list order)
            leaf_materials = [list(ref.call("geometry", "scene-obj-list", scene))[0][2]] + [o[2] for o in main.lookup(Sym("*spheres-list*"))]
        tab = hits_table(ref, scene, rays, leaf_materials)
        tab["unstable"] = stability_mask(hs, rays)
        tab["n_camera_rays"] = len(cam_rays)
        tab["camera_st"] = [list(x) for x in cam_st]
        if name == "random-scene":                 # the material table the reference built: kind, albedo, fuzz / ref-idx per object
            tab["materials"] = [material_row(ref, mat) for mat in leaf_materials]
        out.append(dict(name=name, **tab))
        print(f"scenes/{name}: {sum(tab['hit'])}/{len(rays)} hits, {sum(tab['unstable'])} flagged")
    return dict(source="scenes defined by main.scm, closest hit by (g:hit scene ray 0.001 +max-float+); the first n_camera_rays rays were made by cam:get-ray",
                scenes=out)


def make_textures(ref):
    it = ref.it
    per = it.modules["perlin"]
    ranvec = [list(x) for x in per.lookup(Sym("+ranvec+"))]
    perms = [list(per.lookup(Sym(n))) for n in ("+perm-x+", "+perm-y+", "+perm-z+")]
    rs = np.random.RandomState(7)
    pts = f32(np.concatenate([rs.uniform(-6, 6, (150, 3)), rs.uniform(-300, 300, (50, 3))]))
    noise = [float(ref.call("perlin", "noise", v3(*p))) for p in pts]
    turb = [float(ref.call("perlin", "turb", v3(*p))) for p in pts]
    tex = {"checker": ref.call("texture", "checker-texture", ref.call("texture", "constant-texture", v3(0.2, 0.3, 0.1)), ref.call("texture", "constant-texture", v3(0.9, 0.9, 0.9))),
           "noise4": ref.call("texture", "noise-texture", 4), "marble1": ref.call("texture", "marble-texture", 1), "marble0.25": ref.call("texture", "marble-texture", 0.25)}
    vals = {k: [list(ref.call("texture", "value", tx, 0, 0, v3(*p))) for p in pts] for k, tx in tex.items()}
    print(f"textures: {len(pts)} points, noise range [{min(noise):.3f}, {max(noise):.3f}]")
    return dict(source="perlin.scm tables as built at module load from the scripted random-real; noise / turb / t:value outputs",
                ranvec=ranvec, perm_x=perms[0], perm_y=perms[1], perm_z=perms[2], points=[list(p) for p in pts], noise=noise, turb=turb, textures=vals)


def make_materials(ref, main, rng):
    rs = np.random.RandomState(11)
    out = dict(source="material.scm / util.scm / onb.scm / main.scm procedures called directly")
    vs, ns = f32(rs.normal(size=(40, 3)) * rs.uniform(0.5, 3, (40, 1))), rs.normal(size=(40, 3))
    ns = f32(ns / np.linalg.norm(ns, axis=1, keepdims=True))
    out["reflect"] = [dict(v=list(a), n=list(b), out=list(ref.call("material", "reflect", v3(*a), v3(*b)))) for a, b in zip(vs, ns)]
    refr = []
    for a, b, ratio in zip(vs, ns, [1 / 1.5, 1.5] * 20):
        r = ref.call("material", "refract", v3(*a), v3(*b), ratio)
        refr.append(dict(v=list(a), n=list(b), ni_over_nt=ratio, ok=r[0] is not False, out=list(r[1]) if r[0] is not False else None))
    out["refract"] = refr
    out["schlick"] = [dict(cosine=float(c), ref_idx=1.5, out=float(ref.call("material", "schlick", float(c), 1.5))) for c in f32(rs.uniform(0, 1, 20))]
    cos = []
    for b in ns[:30]:
        r6 = [float(x) for x in f32(rs.uniform(0.001, 0.999, 6))]
        rng.script = list(r6)
        uvw = ref.call("onb", "make-onb-from-w", v3(*b))
        d = ref.it.eval([Sym("local"), Sym("uvw"), [Sym("random-cosine-direction")]], _env_with(ref.it.modules["material"], uvw=uvw))
        assert not rng.script, "(local uvw (random-cosine-direction)) is expected to consume six draws (Q15)"
        cos.append(dict(w=list(b), draws=r6, out=list(d)))
    out["onb_cosine"] = cos
    # the same macro with a VARIABLE operand (the single-evaluation case): onb.scm:31-36 itself
    loc = []
    for b, a in zip(ns[:10], vs[10:20]):
        uvw = ref.call("onb", "make-onb-from-w", v3(*b))
        d = ref.it.eval([Sym("local"), Sym("uvw"), Sym("a")], _env_with(ref.it.modules["material"], uvw=uvw, a=v3(*a)))
        loc.append(dict(w=list(b), a=list(a), out=list(d)))
    out["onb_local"] = loc
    sky = main.lookup(Sym("sky-color"))
    out["sky"] = [dict(d=list(a), out=list(ref.it.apply(sky, [ref.call("ray", "make-ray", v3(0, 0, 0), v3(*a))]))) for a in vs[:20]]
    # lambertian scatter + scattering-pdf and diffuse-light emitted on a hit record
    lam, scat = ref.lam, []
    for a, b in zip(vs[:20], ns[:20]):
        r6 = [float(x) for x in f32(rs.uniform(0.001, 0.999, 6))]
        rng.script = list(r6)
        rec = ref.call("ray", "make-hit-record", 1.0, v3(1, 2, 3), v3(*b), lam, 0, 0)
        ray = ref.call("ray", "make-ray", v3(0, 0, 0), v3(*a))
        valid, scattered, atten, pdf = ref.call("material", "scatter", lam, ray, rec)
        spdf = ref.call("material", "scattering-pdf", lam, ray, rec, scattered)
        assert not rng.script
        scat.append(dict(n=list(b), draws=r6, dir=list(scattered[1]), time=float(scattered[2]), atten=list(atten), pdf=float(pdf), spdf=float(spdf)))
    out["lambertian"] = scat
    light = ref.call("material", "make-diffuse-light", ref.call("texture", "constant-texture", v3(4, 4, 4)))
    em = []
    for a, b in zip(vs[:20], ns[:20]):
        rec = ref.call("ray", "make-hit-record", 1.0, v3(1, 2, 3), v3(*b), light, 0.25, 0.75)
        e = ref.call("material", "emitted", light, ref.call("ray", "make-ray", v3(0, 0, 0), v3(*a)), rec, 0.25, 0.75, v3(1, 2, 3))
        em.append(dict(d=list(a), n=list(b), out=list(e)))
    out["diffuse_light_emitted"] = em
    # What HEAD cannot run (SURVEY "five facts" 3), established by running it.  Each entry is the error the call raised.
    from oracle.minischeme import SchemeError, to_list
    errs = {}

    def failing(label, thunk):
        try:
            thunk()
            errs[label] = None
        except SchemeError as e:
            errs[label] = str(e)
    color = main.lookup(Sym("color"))
    for label, mat in (("color_on_metal", ref.call("material", "make-metal", ref.call("texture", "constant-texture", v3(0.8, 0.6, 0.2)), 0.3)),
                       ("color_on_dielectric", ref.call("material", "make-dielectric", 1.5))):
        scene = ref.call("geometry", "make-scene", to_list([ref.call("geometry", "make-sphere", v3(0, 0, -1), 0.5, mat)]),
                         main.lookup(Sym("*camera*")), sky)
        rng.script = [0.25] * 8
        failing(label, lambda: ref.it.apply(color, [ref.call("ray", "make-ray", v3(0, 0, 0), v3(0, 0, -1)), scene]))
        rng.script = []
    sph = ref.call("geometry", "make-sphere", v3(0, 0, -1), 0.5, lam)
    hp = ref.call("pdf", "make-hitable-pdf", sph, v3(0, 0, 0))
    failing("hitable_pdf_value", lambda: ref.call("pdf", "pdf-value", hp, v3(0, 0, -1)))
    failing("hitable_pdf_generate", lambda: ref.call("pdf", "generate", hp))
    out["head_errors"] = errs
    print("materials: done")
    return out


TIE_OBJECTS = [["box", [0, 0, 0], [1, 1, 1]], ["sphere", [0, 0, -1], 0.5], ["xy-rect", -1, 1, -1, 1, -0.5], ["sphere", [3, 0, 0], 1],
               ["xy-rect", 2, 4, -1, 1, 0], ["box", [1, 0, 0], [2, 1, 1]]]
TIE_RAYS = [
    [2, 0.5, 2, -1, 0, -1, 0],        # box edge shared by two faces
    [2, 2, 2, -1, -1, -1, 0],         # box corner (three faces)
    [0.5, 0.5, 5, 0, 0, -1, 0],       # straight through the box
    [0, 0, 0, 0, 0, -1, 0],           # sphere front == rect plane (strict vs inclusive tie)
    [0, 0, 5, 0, 0, -1, 0],
    [3, 0, 5, 0, 0, -1, 0],           # sphere then coplanar rect through its centre
    [1, 0.5, 5, 0, 0, -1, 0],         # shared face plane x = 1 of the two boxes, ray in the plane
    [1.5, 0.5, 5, 0, 0, -1, 0],
    [-5, 0.5, 0.5, 1, 0, 0, 0],       # through both boxes, coincident faces at x = 1
    [5, 0.5, 0.5, -1, 0, 0, 0],
    [0.5, 5, 0.5, 0, -1, 0, 0],
    [0.5, 0.5, 0.5, 0, 0, 1, 0],      # from inside the box
    [3, 0, 0, 0, 1, 0, 0],            # from the sphere centre, in the rect plane (parallel: t = NaN)
    [0, 0.5, 0.5, 1, 0, 0, 0],        # origin ON a face (t = 0 < t-min), exits through the opposite one
    [0.5, 0.5, -1, 0, 0, 1, 0],       # sphere centre towards the box: exits the sphere at t = 0.5, rect at -0.5 in between
    [2, 1, 0.5, -1, 0, 0, 0],         # along a box edge line (y = 1 is the inclusive rect bound)
]


# tie rule proper: sphere (strict <) against rect (inclusive range test) at exactly equal t, identical twins of
# each kind, in both list orders; no ray lies in a rect's plane here
TIE_OBJECTS_B = [["sphere", [0, 0, -1], 0.5], ["xy-rect", -1, 1, -1, 1, -0.5], ["sphere", [0, 0, -1], 0.5], ["xy-rect", -1, 1, -1, 1, -0.5],
                 ["sphere", [0, 0, 2], 0.5], ["xy-rect", -2, 2, -2, 2, 2.5]]
TIE_RAYS_B = [[0, 0, 5, 0, 0, -1, 0], [0, 0, 5, 0, 0, -2, 0.5], [0.125, 0.25, 5, 0, 0, -1, 0], [0, 0, -5, 0, 0, 1, 0], [0, 0, 1, 0, 0, -1, 0],
              [0, 0, 0.5, 0, 0, 1, 0], [0.75, 0.75, 5, 0, 0, -1, 0], [0, 0, -1, 0, 0, 1, 0], [0, 0, -1, 0, 0, -1, 0], [1, 1, 5, 0, 0, -1, 0], [1, -1, 5, 0, 0, -1, 0]]


def make_bezier_kats(ref):
    """KAT7-10 of SURVEY.md 8(c) were derived by the survey from a transliteration, "not from Gauche".  Here the reference's
    own bezier.scm answers the same rays (curve of main.scm:259-263, origin (0, 5, 5))."""
    from tests.test_golden import KAT, bezier_kat_ray
    bz = KAT["bezier"]
    obj = ref.build(["bezier"] + bz["cp"] + [bz["width"]])
    out = []
    for k in bz["cases"]:
        ray = bezier_kat_ray(k)
        h = ref.hit(obj, ray, KAT["t_min"], KAT["t_max"])
        out.append(dict(kat=k["kat"], ray=[float(x) for x in ray], hit=h is not None, t=h and h["t"], p=h and h["p"], n=h and h["n"]))
    print("bezier KATs:", [(o["kat"], o["hit"], o["t"]) for o in out])
    return out


def make_ties(ref):
    a = make_ties_scene(ref, TIE_OBJECTS, TIE_RAYS)
    b = make_ties_scene(ref, TIE_OBJECTS_B, TIE_RAYS_B)
    c = make_ties_scene(ref, TIE_OBJECTS_B[::-1], TIE_RAYS_B)
    return dict(source=a.pop("source"), t_min=0.001, scenes=[a, b, c], bezier_kats=make_bezier_kats(ref))


def make_ties_scene(ref, TIE_OBJECTS, TIE_RAYS):
    """ref_ties.json: the structured adversarial rays of tests/test_gpu_parity.test_structured_ties_cornell through the
    REFERENCE's hit-obj-list (geometry.scm:33-50): exact ties between objects (strict `<` of spheres vs inclusive range
    test of rects, list order), rays inside a rect's plane (NaN), origins on a surface, t-max just before / beyond a hit.
    Every object has its own material, so the record tells which object won."""
    mats = [ref.call("material", "make-lambertian", ref.call("texture", "constant-texture", v3(0.1 * (i + 1), 0.5, 0.5))) for i in range(len(TIE_OBJECTS))]
    objs = []
    for spec, mat in zip(TIE_OBJECTS, mats):
        saved, ref.lam = ref.lam, mat
        objs.append(ref.build(spec))
        ref.lam = saved
    from oracle.minischeme import to_list
    scene = ref.call("geometry", "make-scene", to_list(objs), None, None)
    runs = []
    for tmax in (MAXF, 4.0, 3.999, 4.001, 1.0, 0.5):
        rows = []
        for r in TIE_RAYS:
            ray = ref.call("ray", "make-ray-with-time", v3(*r[0:3]), v3(*r[3:6]), float(r[6]))
            res = ref.call("geometry", "hit", scene, ray, 0.001, tmax)
            if res[0] is False:
                rows.append(dict(obj=-1))
            else:
                rec = res[1]
                rows.append(dict(obj=[k for k, mm in enumerate(mats) if mm is rec[3]][0], t=float(rec[0]), n=list(rec[2]), p=list(rec[1])))
        runs.append(dict(t_max=tmax, hits=rows))
    print(f"ties: {len(TIE_RAYS)} rays x {len(runs)} t-max values, hits at t-max = inf: {sum(h['obj'] >= 0 for h in runs[0]['hits'])}")
    return dict(source="geometry.scm:33-50 hit-obj-list over boxes / spheres / rects, each object with its own material", objects=TIE_OBJECTS,
                rays=TIE_RAYS, runs=runs)


def make_scatter(ref, rng):
    """ref_scatter.json: the scatter closures of metal / dielectric called directly (under `color` they cannot run at
    HEAD: 3 values into a 4-value receive), the rejection samplers, random-to-sphere, the cosine pdf,
    get-ray with a lens and a shutter interval, and the AABB slab test (Q11)."""
    from oracle import oracle as O
    rs = np.random.RandomState(23)
    out = dict(source="material.scm:45-57, 76-101 scatter closures; util.scm:9-23, 46-54; pdf.scm:18-41; camera.scm:80-92; geometry.scm:73-105",
               rng="draws are this repo's Philox uniforms: key (pixel, seed), counter (sample, bounce, block, 0)")
    vs, ns = f32(rs.normal(size=(36, 3)) * rs.uniform(0.5, 3, (36, 1))), rs.normal(size=(36, 3))
    ns = f32(ns / np.linalg.norm(ns, axis=1, keepdims=True))
    tex = ref.call("texture", "constant-texture", v3(0.8, 0.6, 0.2))

    def sphere_stream(seed, pixel, sample, bounce, first=0):       # util.scm:9-15: iteration j = block first + j, components 0..2
        j = first
        while True:
            u = O.rng_block(seed, pixel, sample, bounce, j)
            yield float(u[0]); yield float(u[1]); yield float(u[2])
            j += 1

    def disk_stream(seed, pixel, sample, bounce, first, then):     # util.scm:17-23: two candidates per block; `then` follows the accepted one
        j = first
        while True:
            u = O.rng_block(seed, pixel, sample, bounce, j)
            for h in (0, 2):
                yield float(u[h]); yield float(u[h + 1])
                if (2 * u[h] - 1) ** 2 + (2 * u[h + 1] - 1) ** 2 < 1:
                    for x in then:
                        yield x
                    raise AssertionError("more draws than get-ray is known to make")
            j += 1
    metal = []
    for i, (a, b) in enumerate(zip(vs, ns)):
        fuzz = [0.0, 0.3, 1.0][i % 3]
        m_ = ref.call("material", "make-metal", tex, fuzz)
        rec = ref.call("ray", "make-hit-record", 1.0, v3(1, 2, 3), v3(*b), m_, 0, 0)
        addr = (5, i, 2, 3)
        rng.gen = sphere_stream(*addr)
        valid, scattered, atten = ref.call("material", "scatter", m_, ref.call("ray", "make-ray-with-time", v3(0, 0, 0), v3(*a), 0.75), rec)
        rng.gen = None
        metal.append(dict(d=list(a), n=list(b), fuzz=fuzz, addr=addr, valid=valid is not False, dir=list(scattered[1]), time=float(scattered[2]), atten=list(atten)))
    out["metal"] = metal
    die = []
    for i, (a, b) in enumerate(zip(vs, ns)):
        ref_idx = [1.5, 2.4][i % 2]
        xi = float(f32(rs.uniform(0.001, 0.999)))
        m_ = ref.call("material", "make-dielectric", ref_idx)
        rec = ref.call("ray", "make-hit-record", 1.0, v3(1, 2, 3), v3(*b), m_, 0, 0)
        rng.script = [xi]
        valid, scattered, atten = ref.call("material", "scatter", m_, ref.call("ray", "make-ray-with-time", v3(0, 0, 0), v3(*a), 0.75), rec)
        assert not rng.script
        die.append(dict(d=list(a), n=list(b), ref_idx=ref_idx, xi=xi, valid=valid is not False, dir=list(scattered[1]), time=float(scattered[2]), atten=list(atten)))
    out["dielectric"] = die
    sph, dsk = [], []
    for i in range(20):
        addr = (9, i, 1, 4)
        rng.gen = sphere_stream(*addr, first=3)
        sph.append(dict(addr=addr, first_block=3, out=list(ref.call("util", "random-in-unit-sphere"))))
        rng.gen = disk_stream(*addr, first=1, then=[])
        dsk.append(dict(addr=addr, first_block=1, out=list(ref.call("util", "random-in-unit-disk"))))
        rng.gen = None
    out["random_in_unit_sphere"], out["random_in_unit_disk"] = sph, dsk
    rts = []
    for i in range(20):
        radius, dist = float(f32(rs.uniform(0.2, 3))), float(f32(rs.uniform(3.5, 20)))
        r1, r2 = (float(x) for x in f32(rs.uniform(0.001, 0.999, 2)))
        rng.script = [r1, r2]
        rts.append(dict(radius=radius, distance_sq=dist * dist, r1=r1, r2=r2, out=list(ref.call("util", "random-to-sphere", radius, dist * dist))))
    out["random_to_sphere"] = rts
    cpdf = []
    for a, b in zip(vs[:30], ns[:30]):
        pdf = ref.call("pdf", "make-cosine-pdf", v3(*b))
        p2 = ref.call("pdf", "make-cosine-pdf", v3(*a))
        mix = ref.call("pdf", "make-mixture-pdf", pdf, p2)
        dirv = f32(rs.normal(size=3))
        cpdf.append(dict(w=list(b), w2=list(a), direction=list(dirv), value=float(ref.call("pdf", "pdf-value", pdf, v3(*dirv))),
                         mixture_value=float(ref.call("pdf", "pdf-value", mix, v3(*dirv)))))
    out["cosine_pdf"] = cpdf
    # camera.scm:63-92 with a lens and a shutter interval (the cfg2 / cfg3 cameras of this repo)
    cams = []
    for ci, args in enumerate([((13, 2, 3), (0, 0, 0), (0, 1, 0), 20, 1.5, 0.1, 10, 0, 1), ((278, 278, -800), (278, 278, 0), (0, 1, 0), 40, 1, 2.0, 800, 0.25, 0.75)]):
        cam = ref.call("camera", "make-camera", v3(*args[0]), v3(*args[1]), v3(*args[2]), *args[3:])
        rays = []
        for i in range(24):
            s_, t_ = (float(x) for x in f32(rs.uniform(0, 1, 2)))
            addr = (3, 100 * ci + i, i % 4)
            xi_time = float(O.rng_block(addr[0], addr[1], addr[2], 0, 0)[2])
            rng.gen = disk_stream(addr[0], addr[1], addr[2], 0, first=1, then=[xi_time])
            ray = ref.call("camera", "get-ray", cam, s_, t_)
            rng.gen = None
            rays.append(dict(s=s_, t=t_, addr=addr, xi_time=xi_time, ray=list(ray[0]) + list(ray[1]) + [float(ray[2])]))
        cams.append(dict(args=[list(a) if isinstance(a, tuple) else a for a in args], slots=[list(x) if isinstance(x, F64) else float(x) for x in cam], rays=rays))
    out["camera"] = cams
    # geometry.scm:73-105 make-aabb hit (Q11: per-axis tests, the interval is not carried across axes)
    boxes = []
    for i in range(60):
        lo = f32(rs.uniform(-3, 1, 3)); hi = f32(lo + rs.uniform(0.2, 3, 3))
        o = f32(rs.uniform(-6, 6, 3)); d = f32(rs.normal(size=3))
        if i % 2 == 0:
            d = f32((lo + (hi - lo) * rs.uniform(-0.2, 1.2, 3)) - o)  # aimed at the box and just past its faces
        if i % 5 == 0:
            d[i % 3] = 0.0                                          # 1 / 0.0 = +inf.0
        if i % 7 == 0:
            o = f32(lo + (hi - lo) * rs.uniform(0.1, 0.9, 3))       # origin inside
        tmin, tmax = 0.001, [MAXF, 2.0, 0.5][i % 3]
        box = ref.call("geometry", "make-aabb", v3(*lo), v3(*hi))
        r = ref.it.apply(box[0], [ref.call("ray", "make-ray", v3(*o), v3(*d)), tmin, tmax])
        boxes.append(dict(bmin=list(lo), bmax=list(hi), o=list(o), d=list(d), t_min=tmin, t_max=tmax, hit=r[0] is not False))
    out["aabb"] = boxes
    print(f"scatter: metal valid {sum(k['valid'] for k in metal)}/{len(metal)}, aabb hits {sum(k['hit'] for k in boxes)}/{len(boxes)}")
    return out


def _env_with(module, **vars):
    from oracle.minischeme import Env
    return Env(module, {Sym(k): val for k, val in vars.items()})


class ScriptedRng:
    """random-real for the reference.  Default: the constant 0.5 (lens disk -> zero offset, shutter time
    0.5).  `script`: a list consumed first.  `path`: (seed, pixel-of-call, sample) -> the oracle's Philox
    draws in the order main.scm / camera.scm / material.scm consume them along one path."""

    def __init__(self):
        self.script, self.path, self.k = [], None, 0
        self.medium_leaf, self.ray_index, self.gen = None, 0, None
        self.lens, self.pathgen, self.cur_depth = False, None, -1

    def __call__(self):
        if self.script:
            return self.script.pop(0)
        if self.gen is not None:                   # a Python generator that decides the next draw from what it handed out before
            return next(self.gen)
        if self.medium_leaf is not None:           # see tag_medium
            from oracle import oracle as O
            if self.path is not None:              # inside a path: the hit at depth d draws from bounce d + 1; cur_depth = depth of the last scatter drawn
                seed, pixel, sample = self.path
                return float(O.rng_block(seed, pixel, sample, self.cur_depth + 2, 16 + self.medium_leaf)[0])
            return float(O.rng_block(0, self.ray_index, 0, 1, 16 + self.medium_leaf)[0])
        if self.path is None:
            return 0.5
        if self.k == 0:
            self.pathgen, self.cur_depth = self.path_stream(*self.path), -1
        self.k += 1
        return float(next(self.pathgen))

    def path_stream(self, seed, pixel, sample):
        """The draws of one path in the reference's call order, each taken from the Philox slot the oracle / the CUDA
        path assign to it (key (pixel, seed), counter (sample, bounce, block, 0))."""
        from oracle import oracle as O
        b0 = O.rng_block(seed, pixel, sample, 0, 0)
        yield b0[0]                                                # u, v jitter (main.scm:476-477)
        yield b0[1]
        if not self.lens:
            yield 0.5                                              # lens disk (camera.scm:81): radius 0, the first candidate (0, 0) is accepted
            yield 0.5                                              # and multiplied by zero - the oracle draws nothing
        else:
            j, done = 1, False                                     # random-in-unit-disk (util.scm:17-23): two candidates per block from block 1 on
            while not done:
                u = O.rng_block(seed, pixel, sample, 0, j)
                for h in (0, 2):
                    yield u[h]
                    yield u[h + 1]
                    if (2 * u[h] - 1) ** 2 + (2 * u[h + 1] - 1) ** 2 < 1:
                        done = True
                        break
                j += 1
        yield b0[2]                                                # shutter time (camera.scm:84)
        depth = 0
        while True:
            # lambertian at this depth (material.scm:27): (local uvw (random-cosine-direction)) evaluates its operand THREE
            # times (onb.scm:27-36, Q15) = 6 draws per scatter: block 0 (x, y), then block 2 (x, y), (z, w) of bounce depth + 1
            a, b = O.rng_block(seed, pixel, sample, depth + 1, 0), O.rng_block(seed, pixel, sample, depth + 1, 2)
            self.cur_depth = depth
            for x in (a[0], a[1], b[0], b[1], b[2], b[3]):
                yield x
            depth += 1


NEXTWEEK_SCENE = """
(define ref-nextweek-scene
  (g:make-scene
   (list (g:make-sphere (v:vec3 0 -1000 0) 1000
                        (m:make-lambertian (t:checker-texture (t:constant-texture (v:vec3 0.2 0.3 0.1))
                                                              (t:constant-texture (v:vec3 0.9 0.9 0.9)))))
         (g:make-moving-sphere (v:vec3 0 1 0) (v:vec3 0 1.75 0) 0 1 0.75
                               (m:make-lambertian (t:constant-texture (v:vec3 0.7 0.3 0.1))))
         (g:make-sphere (v:vec3 -2.5 1 0) 1 (m:make-lambertian (t:noise-texture 4)))
         (g:make-sphere (v:vec3 2.5 1 0) 1 (m:make-lambertian (t:marble-texture 1)))
         (g:flip-normals (g:make-xz-rect -1 1 -1 1 3.5 (m:make-diffuse-light (t:constant-texture (v:vec3 4 4 4))))))
   (cam:make-camera (v:vec3 13 2 3) (v:vec3 0 0 0) (v:vec3 0 1 0) 20 1 0.5 10 0 1)
   sky-color))
"""


def make_color(ref, main, rng, size=10, spp=2, seed=7, max_depth=12):
    it = ref.it
    forms = {}
    for form in read_all(open(os.path.join(REFERENCE, "main.scm")).read()):
        if isinstance(form, list) and form and form[0] == "define" and not isinstance(form[1], list) and form[1] in ("*image*", "*raw-data*"):
            forms[form[1]] = form
    main.vars[Sym("*size-x*")] = size           # parameters of the run, like +max-depth+; the cameras keep aspect 1
    main.vars[Sym("*size-y*")] = size
    main.vars[Sym("+max-depth+")] = max_depth
    color = main.lookup(Sym("color"))
    state = dict(pixel=0)

    def color_hook(ray, scene):                  # instrumentation only: tells the scripted RNG where a path ends
        r = it.apply(color, [ray, scene])
        state["pixel"] += 1
        rng.path, rng.k = (seed, state["pixel"], rng.path[2]), 0
        return r
    main.vars[Sym("color")] = color_hook
    out = []
    # a third scene, not in main.scm but built with the reference's constructors and rendered by the reference's trace-all:
    # the Next-Week features in one path loop - moving sphere (shutter 0..1, Q6), thin-lens camera, checker / noise / marble
    # textures, a flipped light, sky-color
    for form in read_all(NEXTWEEK_SCENE):
        it.eval(form, main)
    # (test-bezier last: curve hits carry the un-normalised normal -d (Q9), which scales the lambertian weight by its length)
    for name in ["cornell-box", "test-scene2", "ref-nextweek-scene", "cornell-smoke", "test-bezier"]:
        scene = main.lookup(Sym(name))
        if name == "cornell-smoke":                # constant media inside the path loop: the free-flight draw comes from block 16 + leaf of the hit's bounce
            objs = list(ref.call("geometry", "scene-obj-list", scene))
            tag_medium(ref, objs[6], 6)
            tag_medium(ref, objs[7], 7)
        rng.lens = name == "ref-nextweek-scene"
        it.eval(forms["*image*"], main)
        it.eval(forms["*raw-data*"], main)
        for s in range(spp):
            state["pixel"] = 0
            rng.path, rng.k = (seed, 0, s), 0
            it.call("main", "trace-all", scene, s + 1)
        raw = [list(x) for x in main.lookup(Sym("*raw-data*"))]
        img = list(main.lookup(Sym("*image*")))
        # the progressive viewer's route (main.scm:452-469, 533-544): animate calls trace-line row by row, one new sample per
        # pass.  Same draws => it must leave the same *raw-data* / *image* behind as trace-all did.
        it.eval(forms["*image*"], main)
        it.eval(forms["*raw-data*"], main)
        passes = []                                            # *raw-data* / *image* as the viewer shows them after every pass
        for s in range(spp):
            state["pixel"] = 0
            rng.path, rng.k = (seed, 0, s), 0
            for y in range(size):
                it.call("main", "trace-line", scene, y, s + 1)
            praw = [list(x) for x in main.lookup(Sym("*raw-data*"))]
            pund = [any(c < 0 for c in px) for px in praw]
            passes.append(dict(raw_data=praw, image=[-1 if pund[i // 3] else q for i, q in enumerate(list(main.lookup(Sym("*image*"))))]))
        assert [list(x) for x in main.lookup(Sym("*raw-data*"))] == raw and list(main.lookup(Sym("*image*"))) == img, "trace-line != trace-all"
        import tempfile
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:           # (save-as-ppm nx ny) writes "test.ppm" into the current directory (main.scm:440)
            os.chdir(tmp)
            try:
                it.call("main", "save-as-ppm", size, size)
                ppm = open("test.ppm").read()
            finally:
                os.chdir(cwd)
        # L4 (main.scm:481-487): a negative radiance sum (the noise texture goes negative, texture.scm:25-28) makes correct-gamma
        # take the sqrt of a negative number - a complex in Gauche, which `min` then rejects with an error.  The interpreter does
        # not model that: such pixels are stored as -1 (undefined upstream) and the PPM text is dropped for the run.
        undefined = [any(c < 0 for c in px) for px in raw]
        if any(undefined):
            img = [-1 if undefined[i // 3] else q for i, q in enumerate(img)]
            ppm = None
        out.append(dict(scene=name, width=size, height=size, spp=spp, seed=seed, max_depth=max_depth, raw_data=raw, image=img, ppm=ppm,
                        trace_line_equals_trace_all=True, trace_line_passes=passes))
        print(f"color/{name}: mean radiance {np.mean(raw) / spp:.4f}")
    rng.path, rng.lens = None, False
    main.vars[Sym("color")] = color
    return dict(source="main.scm trace-all (color, running sum, correct-gamma, 8-bit) with random-real returning the oracle's Philox draws "
                       "(key (pixel, seed), counter (sample, bounce, block, 0)) in the reference's call order; y = 0 is the bottom row",
                runs=out)


def main():
    rng = ScriptedRng()
    rs = np.random.RandomState(3)
    rng.script = [float(x) for x in rs.random_sample(256 + 3 * 256 + 3 * 255)]     # consumed by perlin.scm at module load
    ref = Ref(rng)
    assert not rng.script, "perlin.scm consumed a different number of random-real calls than expected"
    main_mod = ref.load_main(MAIN_NAMES)

    def dump(name, obj):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(obj, f)
        print("wrote", name, os.path.getsize(os.path.join(HERE, name)) // 1024, "KB")
    if "--only-color" in sys.argv:                             # the other files do not depend on it
        dump("ref_color.json", make_color(ref, main_mod, rng))
        return
    dump("ref_prims.json", make_prims(ref))
    dump("ref_scenes.json", make_scenes(ref, main_mod))
    dump("ref_prims2.json", make_prims(ref, PRIM_CASES2, 300))
    dump("ref_scenes2.json", make_scenes(ref, main_mod, SCENES2, 400))
    dump("ref_textures.json", make_textures(ref))
    dump("ref_materials.json", make_materials(ref, main_mod, rng))
    dump("ref_scatter.json", make_scatter(ref, rng))
    dump("ref_ties.json", make_ties(ref))
    dump("ref_color.json", make_color(ref, main_mod, rng))


if __name__ == "__main__":
    sys.setrecursionlimit(200000)
    threading.stack_size(512 * 1024 * 1024)
    t = threading.Thread(target=main)
    t.start()
    t.join()
