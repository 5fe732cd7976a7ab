"""CPU: the repo's Gauche host (scheme_raytrace_b200/scheme/*.scm) EXECUTED.

No `gosh` exists in the build image, so the Scheme host modules are run by oracle/minischeme.py (the
interpreter that also executes the reference for the golden vectors).  Two checks:

1. a scene script written in the reference's style (tests/scheme_scripts/demo-scene.scm: every constructor
   of the drop-in API, instances, a medium, a curve, a patch, an image texture) is flattened and written by
   `srt:write-scene`; the file must equal, table by table, the one the Python host writes for the same scene;
2. the scene definitions of the REFERENCE's own main.scm (test-scene, test-scene2, cornell-box, cornell-bezier,
   cornell-smoke, test-bezier, klein-scene, cornell-klein, test-scene-bvh: their text is evaluated unmodified)
   load on top of the repo's modules - that is what "drop-in" means - and flatten to the same tables as the
   Python mirrors of those scenes (only when /root/reference is present).

Equality: integer columns exact; float columns equal after rounding to fp32 (the precision of the C-ABI tables;
the Python writer rounds, the Scheme writer prints f64) within 2 fp32 ulps for composed transforms.
"""
import os
import sys
import threading
import numpy as np
import pytest
import scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import geometry as g, material as m, texture as t, bezier as b, camera as cam, scenes, scenefile, vec as v
from scheme_raytrace_b200.host.perlin import perlin_generate
from tests.refspec import host_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "scheme_raytrace_b200", "scheme")
SCRIPTS = os.path.join(ROOT, "tests", "scheme_scripts")
REFERENCE = os.environ.get("SRT_REFERENCE", "/root/reference")
INT_COLS = {"textures": 3, "materials": 2, "xforms": 0, "patches": 0, "prims": 4}


def in_big_stack(fn):
    """The interpreter recurses on the Python stack."""
    box = {}

    def work():
        sys.setrecursionlimit(200000)
        try:
            box["out"] = fn()
        except BaseException as e:          # re-raised in the caller's thread
            box["err"] = e
    threading.stack_size(256 * 1024 * 1024)
    th = threading.Thread(target=work)
    th.start()
    th.join()
    threading.stack_size(0)
    if "err" in box:
        raise box["err"]
    return box["out"]


def interpreter(load_path):
    from oracle.minischeme import Interp
    it = Interp(load_path, lambda: 0.5)
    it.stub_modules |= {"srfi-1", "gauche.process"}
    return it


def write_scene(it, scene, path):
    """(srt:write-scene scene path) with random-real replaying the draws host/perlin.perlin_generate(3) consumes
    (write-scene builds the Perlin tables like perlin.scm:10-36 does at load time)."""
    stream = iter(np.random.RandomState(3).random_sample(2048))
    it.random_real = lambda: float(next(stream))
    it.call("srt-scene", "write-scene", scene, path)
    it.random_real = lambda: 0.5


def parse(path):
    toks = open(path).read().split()
    assert toks[:2] == ["srt-scene", "1"]
    pos = 2
    out = {}

    def take(n):
        nonlocal pos
        r = toks[pos:pos + n]
        pos += n
        return r
    assert take(1) == ["sky"]
    out["sky"] = int(take(1)[0])
    assert take(1) == ["camera"]
    out["camera"] = np.array(take(24), np.float64)
    assert take(1) == ["perlin-ranvec"]
    out["ranvec"] = np.array(take(768), np.float64)
    assert take(1) == ["perlin-perm"]
    out["perm"] = np.array(take(768), np.int64)
    for name, width in (("textures", 7), ("materials", 3), ("xforms", 5), ("patches", 48), ("prims", 20)):
        assert take(1) == [name], name
        n = int(take(1)[0])
        out[name] = np.array(take(n * width), np.float64).reshape(n, width)
    assert take(1) == ["lights"]
    out["lights"] = np.array(take(int(take(1)[0])), np.int64)
    assert take(1) == ["images"]
    out["images"] = []
    for _ in range(int(take(1)[0])):
        nx, ny = int(take(1)[0]), int(take(1)[0])
        out["images"].append((nx, ny, np.array(take(3 * nx * ny), np.int64)))
    assert pos == len(toks)
    return out


def assert_same_tables(a, c, label):
    """a: written by the Scheme host, c: written by the Python host."""
    assert a["sky"] == c["sky"], label
    f32 = lambda x: np.asarray(x, np.float64).astype(np.float32)
    assert np.array_equal(f32(a["camera"]), f32(c["camera"])), label
    assert np.array_equal(f32(a["ranvec"]), f32(c["ranvec"])) and np.array_equal(a["perm"], c["perm"]), label
    for name, nint in INT_COLS.items():
        assert a[name].shape == c[name].shape, (label, name, a[name].shape, c[name].shape)
        assert np.array_equal(a[name][:, :nint], c[name][:, :nint]), (label, name)
        assert np.allclose(f32(a[name][:, nint:]), f32(c[name][:, nint:]), rtol=2.4e-7, atol=1e-30), (label, name)
    assert np.array_equal(a["lights"], c["lights"])
    assert len(a["images"]) == len(c["images"])
    for (nx, ny, ta), (mx, my, tc) in zip(a["images"], c["images"]):
        assert (nx, ny) == (mx, my) and np.array_equal(ta, tc)


def python_file(tmp_path, scene, name):
    path = str(tmp_path / f"{name}.py.srt")
    scenefile.write_scene_file(path, srt.flatten_scene(scene), perlin_generate(3))
    return parse(path)


def demo_scene_python():
    checker = t.checker_texture(t.constant_texture(v.vec3(0.2, 0.3, 0.1)), t.constant_texture(v.vec3(0.9, 0.9, 0.9)))
    white = m.make_lambertian(t.constant_texture(v.vec3(0.73, 0.73, 0.73)))
    marble, noise = m.make_lambertian(t.marble_texture(0.25)), m.make_lambertian(t.noise_texture(4))
    light = m.make_diffuse_light(t.constant_texture(v.vec3(4, 4, 4)))
    image = m.make_lambertian(t.image_texture([255, 0, 0, 0, 255, 0, 0, 0, 255, 10, 20, 30, 40, 50, 60, 70, 80, 90], 3, 2))
    box = g.translate(g.rotate_y(g.make_box(v.vec3(0, 0, 0), v.vec3(1, 2, 1), white), 15), v.vec3(-3, 0, 1))
    net = [[v.vec3(i, 0.25 * i * j, j) for j in range(4)] for i in range(4)]
    objs = [
        g.make_sphere(v.vec3(0, -1000, 0), 1000, m.make_lambertian(checker)),
        g.make_sphere(v.vec3(0, 1, 0), 1, m.make_dielectric(1.5)),
        g.make_sphere(v.vec3(0, 1, 0), -0.95, m.make_dielectric(1.5)),
        g.make_sphere(v.vec3(4, 1, 0), 1, m.make_metal(t.constant_texture(v.vec3(0.7, 0.6, 0.5)), 0.25)),
        g.make_moving_sphere(v.vec3(-4, 1, 0), v.vec3(-4, 1.5, 0), 0, 1, 0.5, marble),
        g.flip_normals(g.make_xz_rect(-1, 1, -1, 1, 5, light)),
        g.make_xy_rect(3, 5, 1, 3, -2, noise),
        g.flip_normals(g.make_yz_rect(0, 2, 0, 2, 6, image)),
        box,
        g.make_constant_medium(g.translate(g.make_box(v.vec3(0, 0, 0), v.vec3(1, 1, 1), white), v.vec3(2, 0, 2)), 0.5, t.constant_texture(v.vec3(1, 1, 1))),
        g.make_bvh_node([b.make_bezier(v.vec3(-1, 0, -1), v.vec3(-0.8, 1, 1), v.vec3(0.8, -1, 1), v.vec3(1, 0, -1), 0.1, white),
                         g.make_sphere(v.vec3(2, 0.25, 2), 0.25, m.make_isotropic(t.constant_texture(v.vec3(0.5, 0.5, 0.5))))], 0, 1),
        b.make_bezier_patch(net, white),
        g.make_klein(v.vec3(250, 200, 280), white),
    ]
    c = cam.make_camera(v.vec3(13, 2, 3), v.vec3(0, 0, 0), v.vec3(0, 1, 0), 20, 3 / 2, 0.1, 10, 0, 1)
    return g.make_scene(objs, c, scenes.sky_color)


def test_scene_script_through_the_scheme_host(tmp_path):
    from oracle.minischeme import Sym
    out = str(tmp_path / "demo.scm.srt")

    def work():
        it = interpreter([HOST])
        it.load_file(os.path.join(SCRIPTS, "demo-scene.scm"))      # defines module `demo`
        scene = it.modules["demo"].lookup(Sym("demo-scene"))
        write_scene(it, scene, out)
    in_big_stack(work)
    assert_same_tables(parse(out), python_file(tmp_path, demo_scene_python(), "demo"), "demo-scene")


REF_SCENES = ["test-scene", "test-scene2", "cornell-box", "cornell-bezier", "cornell-smoke", "test-bezier", "klein-scene", "cornell-klein", "test-scene-bvh"]


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference sources only exist in the build container")
def test_reference_scene_scripts_are_drop_in(tmp_path):
    """main.scm's scene definitions, text unmodified, evaluated on top of the repo's modules."""
    from oracle.minischeme import Sym, read_all
    want = {"+black+", "+white+", "sky-color", "black", "*size-x*", "*size-y*", "*cornell-camera*", "*camera*", "line-upped-spheres",
            "*spheres-list*", "*bvh-node*"} | set(REF_SCENES)

    def work():
        it = interpreter([HOST, REFERENCE])    # -I order of scheme/README.md: the repo's modules shadow the reference's
        it.require("srt-scene")
        main = None
        for form in read_all(open(os.path.join(REFERENCE, "main.scm")).read()):
            if not isinstance(form, list) or not form:
                continue
            if form[0] == "define-module":     # main.scm's own (use ...) header: vec, geometry, material, texture, camera, bezier ... resolve to the repo's files
                it.eval(form, it.user)
                main = it.modules["main"]
            elif form[0] in ("define", "define-inline") and main is not None:
                name = form[1][0] if isinstance(form[1], list) else form[1]
                if name in want:
                    it.eval(form, main)
        # main.scm defines its own sky-color / black procedures; the host recognises the sky by identity with ITS exported pair
        sky = {id(main.lookup(Sym("sky-color"))): "sky-color", id(main.lookup(Sym("black"))): "black"}
        for name in REF_SCENES:
            scene = main.lookup(Sym(name))
            scene[4] = it.modules["srt-scene"].lookup(Sym(sky[id(scene[4])]))
            write_scene(it, scene, str(tmp_path / f"{name}.scm.srt"))
    in_big_stack(work)
    for name in REF_SCENES:
        got = parse(str(tmp_path / f"{name}.scm.srt"))
        want_tables = python_file(tmp_path, host_scene(name), name)
        if name == "test-scene-bvh":
            # line-upped-spheres draws its colours from random-real (main.scm:188-190): not part of the comparison
            got["textures"][:, 4:7] = want_tables["textures"][:, 4:7]
        assert_same_tables(got, want_tables, name)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference sources only exist in the build container")
def test_reference_points_module_loads_unchanged_on_the_scheme_host(tmp_path):
    """scheme/README.md: points.scm of the reference (CSV -> control points -> cubic Bezier chain, points.scm:10-50)
    needs no replacement - it only calls vec / bezier procedures the repo's modules export under the same names.
    Executed here: the reference's file, the repo's vec.scm / bezier.scm underneath, against host/points.py."""
    from scheme_raytrace_b200.host import points as pts
    csv = tmp_path / "pts.csv"
    rs = np.random.RandomState(5)
    P = rs.uniform(-2, 2, (7, 3)).round(3)
    csv.write_text("".join(",".join(repr(float(x)) for x in row) + "\n" for row in P))

    def work():
        it = interpreter([HOST, REFERENCE])
        it.stub_modules |= {"srfi-13", "gauche.collection"}
        it.require("points")
        it.require("texture")
        assert it.modules["vec"] is it.require("vec") and "fold-vec" in {str(k) for k in it.modules["vec"].vars}      # the repo's vec.scm, not the reference's
        mat = it.call("material", "make-lambertian", it.call("texture", "constant-texture", it.call("vec", "vec3", 0.5, 0.5, 0.5)))
        points = it.call("points", "load-points", str(csv), 2)
        objs = it.call("points", "bezier->objs", it.call("points", "points->bezier", points), 0.05, mat)
        return [(o[1], [float(x) for x in o[3]]) for o in objs]
    got = in_big_stack(work)
    mat = m.make_lambertian(t.constant_texture(v.vec3(0.5, 0.5, 0.5)))
    want = pts.bezier_to_objs(pts.points_to_bezier(pts.load_points(str(csv), 2)), 0.05, mat)
    assert len(got) == len(want) == 4
    for (kind, params), w in zip(got, want):
        assert kind == g.BEZIER == w.kind and np.allclose(params, w.params, rtol=1e-15, atol=1e-15)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference sources only exist in the build container")
def test_vec_modules_agree_with_the_references():
    """vec.scm of the reference, vec.scm of the repo's Scheme host and host/vec.py on the same operands: sum / diff / prod
    (variadic), scale, quot, dot, length, sq-len, unit, cross - equal to the last bit."""
    rs = np.random.RandomState(8)
    vs = [tuple(float(x) for x in rs.normal(size=3) * 3) for _ in range(12)]
    ops = [("sum", 3), ("diff", 3), ("prod", 3), ("sum", 2), ("diff", 2), ("quot", 2), ("dot", 2), ("cross", 2), ("length", 1), ("sq-len", 1), ("unit", 1)]

    def run(load_path):
        def work():
            it = interpreter(load_path)
            it.require("vec")
            out = []
            for name, n in ops:
                for k in range(0, 12 - n + 1, n):
                    r = it.call("vec", name, *[it.call("vec", "vec3", *vs[k + j]) for j in range(n)])
                    out.append([float(x) for x in r] if isinstance(r, list) else float(r))
            out.append([float(x) for x in it.call("vec", "scale", it.call("vec", "vec3", *vs[0]), 2.5)])
            return out
        return in_big_stack(work)
    ref, mine = run([REFERENCE]), run([HOST])
    assert ref == mine
    py = []
    fn = {"sum": v.sum, "diff": v.diff, "prod": v.prod, "dot": v.dot, "cross": v.cross, "length": v.length, "sq-len": v.sq_len, "unit": v.unit,
          "quot": lambda a, b: (a[0] / b[0], a[1] / b[1], a[2] / b[2])}
    for name, n in ops:
        for k in range(0, 12 - n + 1, n):
            r = fn[name](*[vs[k + j] for j in range(n)])
            py.append(list(r) if isinstance(r, tuple) else float(r))
    py.append(list(v.scale(vs[0], 2.5)))
    assert py == ref


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference sources only exist in the build container")
def test_chain_fixtures_are_what_the_reference_scene_text_produces(tmp_path):
    """tests/golden/chain_<scene>.srt (the committed first half of the drop-in chain, see make_chain_fixtures.py) must be,
    byte for byte, what main.scm's unmodified scene definitions write through the repo's Scheme host today; the GPU half
    (tests/test_gpu_chain.py) feeds these files to cli/srt_render and compares the PPM with the reference-written one."""
    from tests.golden.make_chain_fixtures import write_chain_files, CHAIN_SCENES
    write_chain_files(str(tmp_path))
    gold = os.path.join(ROOT, "tests", "golden")
    for name in CHAIN_SCENES:
        assert open(os.path.join(gold, f"chain_{name}.srt")).read() == open(tmp_path / f"chain_{name}.srt").read(), name
