"""CPU (-m "not gpu"): the oracle against OUTPUTS OF THE REFERENCE ITSELF.

`tests/golden/ref_*.json` were produced by executing /root/reference/*.scm, unmodified, with
oracle/minischeme.py (tests/golden/make_reference_golden.py, build container only).  They are what
pins the oracle: it is evaluated here WITHOUT the fp32 rounding of the scene tables
(`quantise=False`), i.e. as the f64 restatement of the Scheme code on the same f64 inputs, and must
reproduce the reference's numbers to the last few ulps:

    hit / miss of every ray                        exact
    t, p, normal, u, v of g:hit                    |err| <= 1e-12 * max(1, |ref|)
    Perlin noise / turb / texture values           <= 1e-12
    reflect / refract / schlick / cosine+onb / sky <= 1e-12
    trace-all radiance sums, 10 x 10 x 2 spp       <= 1e-9 relative (a sum of products over <= 12 bounces);
    8-bit image                                    exact

A final test re-runs a slice of the generator when /root/reference is present (this container),
so the committed files cannot drift from the interpreter + reference pair that made them.
"""
import json
import os
import numpy as np
import pytest
from scheme_raytrace_b200.host import geometry as g, material as m, texture as t, scenes
from tests.refspec import build_host, host_scene

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-12


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def _check_hits(O, scene, case, t_min=0.001, t_max=999999999999.0):
    S = O.OracleScene(scene, quantise=False)
    rays = np.asarray(case["rays"], np.float64)
    assert np.array_equal(rays.astype(np.float32).astype(np.float64), rays), "golden rays must be fp32-representable"
    o = S.trace_batch(rays, t_min, t_max)
    hit = np.asarray(case["hit"], bool)
    assert np.array_equal(o["prim"] >= 0, hit), np.nonzero((o["prim"] >= 0) != hit)[0][:10]
    if "prim" in case:        # the object the reference hit, identified by the material object in its hit record: ids must be EQUAL
        assert np.array_equal(o["prim"], np.asarray(case["prim"])), np.nonzero(o["prim"] != np.asarray(case["prim"]))[0][:10]
    worst = 0.0
    for key, mine in (("t", o["t"]), ("p", o["p"]), ("n", o["n"]), ("uv", o["uv"])):
        ref = np.asarray(case[key], np.float64)
        a, b = mine[hit], ref[hit]
        assert np.array_equal(np.isnan(a), np.isnan(b)), key          # Q5: sphere u, v are NaN where asin leaves its domain
        err = np.abs(a - b) / np.maximum(np.abs(b), 1.0)
        if err.size:
            worst = max(worst, float(np.nanmax(err)))
    assert worst <= TOL, worst
    return int(hit.sum()), worst


def test_reference_primitives(orc):
    P = load("ref_prims.json")
    kinds = set()
    for case in P["cases"]:
        scene = g.make_scene([build_host(case["spec"])], scenes.default_camera(), scenes.sky_color)
        nhit, worst = _check_hits(orc, scene, case, P["t_min"], P["t_max"])
        assert nhit >= 20, (case["name"], nhit)        # every constructor is actually exercised by hits
        kinds.add(case["spec"][0])
    assert kinds == {"sphere", "moving-sphere", "xy-rect", "xz-rect", "flip", "box", "translate", "bezier"}


def test_reference_scenes(orc):
    with_ids = 0
    for case in load("ref_scenes.json")["scenes"]:
        nhit, _ = _check_hits(orc, host_scene(case["name"]), case)
        assert nhit > 150, case["name"]
        with_ids += "prim" in case
    assert with_ids == 2          # test-scene, test-scene2 (+ the three 101-sphere scenes of ref_scenes2.json): primitive ids against the reference


def test_reference_medium_and_klein(orc):
    """make-constant-medium (geometry.scm:545-578; sphere and instanced-box boundaries; the
    free-flight draw is scripted per (ray, leaf), see the generator's tag_medium) and the sphere-traced Klein
    primitive (geometry.scm:596-664)."""
    P = load("ref_prims2.json")
    for case in P["cases"]:
        scene = g.make_scene([build_host(case["spec"])], scenes.default_camera(), scenes.sky_color)
        nhit, _ = _check_hits(orc, scene, case, P["t_min"], P["t_max"])
        assert nhit >= 40, (case["name"], nhit)


def test_reference_scenes_with_media_klein_curves_and_the_references_bvh(orc):
    """More scenes of main.scm: cornell-smoke, cornell-klein, klein-scene, test-bezier (curves inside make-bvh-node)
    and test-scene-non-bvh / -bvh / -bvh-sah.  The last three hold the same 100 spheres as a plain list, under the
    reference's own make-bvh-node and under make-bvh-with-sah: the reference's BVH traversals return what its list
    returns, and the oracle (which treats make-bvh-* as grouping) reproduces all three - the equivalence the LBVH
    replacement rests on."""
    names = set()
    for case in load("ref_scenes2.json")["scenes"]:
        nhit, _ = _check_hits(orc, host_scene(case["name"]), case)
        assert nhit > 80, case["name"]
        names.add(case["name"])
    assert {"cornell-smoke", "cornell-klein", "klein-scene", "test-bezier", "test-scene-non-bvh", "test-scene-bvh", "test-scene-bvh-sah"} <= names
    ids = {c["name"]: c["prim"] for c in load("ref_scenes2.json")["scenes"] if "prim" in c}
    assert set(ids) == {"test-scene-non-bvh", "test-scene-bvh", "test-scene-bvh-sah", "random-scene"}
    assert all(len(set(v)) >= 60 for k, v in ids.items() if k != "random-scene")


def test_reference_tie_rule_and_edge_rays(orc):
    """ref_ties.json: exact ties, edges, corners, coincident faces, origins on a surface and t-max on either side of a hit,
    through the reference's hit-obj-list.  The winner (which object, by material identity), t and normal must be the
    oracle's.  One documented divergence: a ray lying IN a rect's plane gives t = 0/0 = NaN upstream, every
    comparison is false, the NaN hit is accepted and poisons closest-so-far; the oracle and the CUDA path reject NaN t.
    Those rays (origin coordinate on a rect's plane constant with a zero direction component) are only counted - one of
    them shows the poisoning: at t-max 0.5 the reference returns a hit at t = 1.0."""
    F = load("ref_ties.json")
    compared = poisoned = ties_seen = 0
    for T in F["scenes"]:
        objs = [build_host(spec) for spec in T["objects"]]
        leaf_to_obj = []
        for k, spec in enumerate(T["objects"]):
            leaf_to_obj += [k] * (6 if spec[0] == "box" else 1)
        S = orc.OracleScene(g.make_scene(objs, scenes.default_camera(), scenes.sky_color), quantise=False)
        rays = np.asarray(T["rays"], np.float64)
        planes = []                                   # (axis, k) of every rect of the scene
        for spec in T["objects"]:
            if spec[0] == "box":
                planes += [(a, spec[1][a]) for a in range(3)] + [(a, spec[2][a]) for a in range(3)]
            elif spec[0].endswith("-rect"):
                planes.append(({"xy-rect": 2, "xz-rect": 1, "yz-rect": 0}[spec[0]], spec[5]))
        in_plane = [any(r[3 + a] == 0 and r[a] == k for a, k in planes) for r in T["rays"]]
        assert all(np.isnan(h["t"]) <= in_plane[i] for i, h in enumerate(T["runs"][0]["hits"]) if h["obj"] >= 0)    # every NaN record comes from such a ray
        for run in T["runs"]:
            o = S.trace_batch(rays, F["t_min"], run["t_max"])
            for i, h in enumerate(run["hits"]):
                if in_plane[i]:                       # NaN may have passed through closest-so-far even if the final record is finite
                    poisoned += 1
                    continue
                mine = leaf_to_obj[o["prim"][i]] if o["prim"][i] >= 0 else -1
                assert mine == h["obj"], (T["objects"], run["t_max"], i, mine, h)
                if mine >= 0:
                    assert o["t"][i] == h["t"] and np.array_equal(o["n"][i], np.asarray(h["n"], np.float64)), (run["t_max"], i)
                    # an exact tie: another object of the scene is hit at the very same t (the winner is then decided by the rule alone)
                    t2 = S.trace_batch(rays[i:i + 1], F["t_min"], run["t_max"], exclude_leaf=int(o["prim"][i]))
                    ties_seen += bool(t2["prim"][0] >= 0 and t2["t"][0] == o["t"][i])
                compared += 1
    assert compared >= 150 and 0 < poisoned < compared and ties_seen >= 30, (compared, poisoned, ties_seen)


def test_reference_answers_the_surveys_bezier_kats(orc):
    """KAT7-10 of SURVEY.md 8(c) came from the survey's transliteration of bezier.scm, "not from Gauche".  The reference's
    own bezier.scm (executed) gives the same answers - t = 6.731228242402701 for KAT7 / 8, a miss for KAT9 - and so does
    the oracle."""
    from tests.test_golden import KAT
    bz = KAT["bezier"]
    got = {k["kat"]: k for k in load("ref_ties.json")["bezier_kats"]}
    cps = [c for p in bz["cp"] for c in p]
    for k in bz["cases"]:
        r = got[k["kat"]]
        assert r["hit"] == k["hit"], k["kat"]
        mine = orc.bezier_hit(cps, bz["width"], r["ray"], KAT["t_min"], KAT["t_max"])
        assert mine["hit"] == r["hit"]
        if k["hit"]:
            assert abs(r["t"] - k["t"]) <= 1e-12 and np.allclose(r["p"], k["p"], rtol=0, atol=1e-10), k["kat"]          # reference vs survey
            assert mine["t"] == r["t"] and np.array_equal(mine["p"], r["p"]) and np.array_equal(mine["n"], r["n"]), k["kat"]   # oracle vs reference: bit for bit


def test_reference_random_scene_generator():
    """main.scm:31-89 random-scene, the generator behind cfg2 / cfg3, EXECUTED (its broken last form - make-scene with
    one argument - was confirmed to raise and then bypassed, see the generator): the host mirror
    `scenes.random_scene(seed, -5, 10, moving=True, checker_ground=True)` fed the same random stream builds the same
    objects in the same order - geometry and ids through the hit records above, and here the material table:
    kind, albedo (the reference's texture evaluated), fuzz / refractive index of all 226 objects."""
    from scheme_raytrace_b200.host.flatten import flatten_scene
    case = [c for c in load("ref_scenes2.json")["scenes"] if c["name"] == "random-scene"][0]
    flat = flatten_scene(host_scene("random-scene"))
    assert len(case["materials"]) == len(flat.prims) == 226
    kinds = {"lambertian": 0, "metal": 1, "dielectric": 2}
    seen = set()
    for i, (kind, rgb, param) in enumerate(case["materials"]):
        mrow = flat.materials[flat.prims[i]["material"]]
        assert int(mrow["kind"]) == kinds[kind], i
        seen.add(kind)
        assert np.float32(mrow["param"]) == np.float32(param), i
        if rgb is not None and i != len(case["materials"]) - 1:              # (the last object is the checker ground)
            trow = flat.textures[mrow["tex"]]
            assert int(trow["kind"]) == 0 and np.array_equal(trow["rgb"], np.asarray(rgb, np.float32)), i
    assert seen == set(kinds)
    assert sum(int(p["type"]) == 1 for p in flat.prims) > 100               # moving spheres (the lambertians)


def test_reference_camera_rays(orc):
    """The first n_camera_rays rays of every scene came out of the reference's cam:get-ray on a 12 x 12 (s, t)
    grid `camera_st` (lens radius 0, shutter draw 0.5); the oracle's get_ray must produce the same rays."""
    for case in load("ref_scenes.json")["scenes"]:
        S = orc.OracleScene(host_scene(case["name"]), quantise=False)
        assert len(case["camera_st"]) == case["n_camera_rays"] == 144
        for k, (s_, t_) in enumerate(case["camera_st"]):
            mine = S.get_ray(s_, t_, 0.5, 1, 0, 0)
            ref = np.asarray(case["rays"][k])
            # the stored rays were rounded to fp32 after the reference made them
            assert np.allclose(mine.astype(np.float32), ref.astype(np.float32), rtol=2e-7, atol=1e-30), (case["name"], k)


def _ref_perlin(T):
    return (np.asarray(T["ranvec"], np.float64), np.asarray(T["perm_x"], np.int32), np.asarray(T["perm_y"], np.int32), np.asarray(T["perm_z"], np.int32))


def test_reference_perlin_and_textures(orc):
    T = load("ref_textures.json")
    rv, px, py, pz = _ref_perlin(T)
    assert rv.shape == (256, 3) and sorted(px) == list(range(256)) and sorted(py) == list(range(256)) and sorted(pz) == list(range(256))
    assert np.allclose(np.linalg.norm(rv, axis=1), 1.0, atol=1e-12)                    # perlin.scm:16-23 unit gradients
    lam = m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5)))
    tex = {"checker": t.checker_texture(t.constant_texture((0.2, 0.3, 0.1)), t.constant_texture((0.9, 0.9, 0.9))),
           "noise4": t.noise_texture(4), "marble1": t.marble_texture(1), "marble0.25": t.marble_texture(0.25)}
    objs = [g.make_sphere((i, 0, 0), 0.25, m.make_lambertian(tx)) for i, tx in enumerate(tex.values())]
    scene = g.make_scene(objs + [g.make_sphere((9, 9, 9), 0.25, lam)], scenes.default_camera(), scenes.sky_color)
    S = orc.OracleScene(scene, perlin=(rv, px, py, pz), quantise=False)
    pts = np.asarray(T["points"], np.float64)
    assert np.abs(S.noise(pts) - np.asarray(T["noise"])).max() <= TOL
    assert np.abs(S.noise(pts, turb=True) - np.asarray(T["turb"])).max() <= TOL
    uvp = np.concatenate([np.zeros((len(pts), 2)), pts], axis=1)
    for name, tx in tex.items():
        tid = S._tex[id(tx)]
        mine = S.tex_value(tid, uvp)
        assert np.abs(mine - np.asarray(T["textures"][name])).max() <= TOL, name


def test_host_perlin_tables_are_the_references():
    """perlin.scm:10-30 builds its tables at module load from random-real; the generator scripted random-real
    with numpy's RandomState(3) stream, which is also what the host mirror `perlin_generate(3)` (the tables the
    product uploads by default) draws from: same draws => the reference's tables, bit for bit."""
    from scheme_raytrace_b200.host.perlin import perlin_generate
    T = load("ref_textures.json")
    rv, px, py, pz = perlin_generate(3)
    assert np.array_equal(rv, np.asarray(T["ranvec"], np.float64))
    assert np.array_equal(px, T["perm_x"]) and np.array_equal(py, T["perm_y"]) and np.array_equal(pz, T["perm_z"])


def test_reference_material_functions(orc):
    M = load("ref_materials.json")
    for k in M["reflect"]:
        assert np.abs(orc.reflect(k["v"], k["n"]) - k["out"]).max() <= TOL
    n_ok = 0
    for k in M["refract"]:
        ok, out = orc.refract(k["v"], k["n"], k["ni_over_nt"])
        assert ok == k["ok"]
        if ok:
            n_ok += 1
            assert np.abs(out - k["out"]).max() <= TOL * max(1.0, np.abs(k["out"]).max())
    assert 5 < n_ok < len(M["refract"])                # both branches (refracted / total internal reflection) present
    for k in M["schlick"]:
        assert abs(orc.schlick(k["cosine"], k["ref_idx"]) - k["out"]) <= TOL
    lib = orc.load()
    for k in M["onb_local"]:                           # onb.scm:27-36 with a variable operand: u*a.x + v*a.y + w*a.z
        w, a, out = np.asarray(k["w"], np.float64), np.asarray(k["a"], np.float64), np.zeros(3)
        lib.orc_onb_local(w.ctypes.data, a.ctypes.data, out.ctypes.data)
        assert np.abs(out - k["out"]).max() <= TOL
    for k in M["onb_cosine"]:
        # Q15: (local uvw (random-cosine-direction)) consumed SIX draws in the reference: the macro evaluates its
        # operand three times.  The oracle with the quirk reproduces the result; without it, it does not.
        w, r6, out = np.asarray(k["w"], np.float64), np.asarray(k["draws"], np.float64), np.zeros(3)
        lib.orc_onb_local_cosine(w.ctypes.data, r6.ctypes.data, 31, out.ctypes.data)
        assert np.abs(out - k["out"]).max() <= TOL
        lib.orc_onb_local_cosine(w.ctypes.data, r6.ctypes.data, 15, out.ctypes.data)
        assert np.abs(out - k["out"]).max() > 1e-3
    lam = m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5)))
    S = orc.OracleScene(g.make_scene([g.make_sphere((0, 0, 0), 1, lam)], scenes.default_camera(), scenes.sky_color), quantise=False)
    for k in M["sky"]:                                 # main.scm:91-95 sky-color
        assert np.abs(S.sky(k["d"]) - k["out"]).max() <= TOL
    for k in M["lambertian"]:
        # material.scm:24-39: scattered direction = unit(local uvw (random-cosine-direction)), time 0 (Q6),
        # pdf = dot(w, dir) / pi, scattering-pdf = max(0, cos) / pi, attenuation = albedo
        nrm, r6, out = np.asarray(k["n"], np.float64), np.asarray(k["draws"], np.float64), np.zeros(3)
        lib.orc_onb_local_cosine(nrm.ctypes.data, r6.ctypes.data, 31, out.ctypes.data)
        d = out / np.linalg.norm(out)
        assert np.abs(d - k["dir"]).max() <= TOL and k["time"] == 0.0
        assert abs(np.dot(nrm / np.linalg.norm(nrm), d) / np.pi - k["pdf"]) <= TOL
        assert abs(max(0.0, np.dot(nrm, d)) / np.pi - k["spdf"]) <= TOL
        assert k["atten"] == [0.5, 0.5, 0.5]
    seen = set()
    for k in M["diffuse_light_emitted"]:               # material.scm:103-111: emits only towards the side the normal points to
        facing = float(np.dot(k["n"], k["d"])) < 0.0
        seen.add(facing)
        assert k["out"] == ([4.0, 4.0, 4.0] if facing else [0.0, 0.0, 0.0])
    assert seen == {True, False}
    # what HEAD cannot run, as the reference itself reported it when executed (SURVEY "five facts" 3): `color` receives the
    # three values of the metal / dielectric scatter into four variables; make-hitable-pdf calls procedures that do not exist.
    # These are the parts the oracle defines from the book series and labels "parity unpinned".
    e = M["head_errors"]
    assert "3 values where 4" in e["color_on_metal"] and "3 values where 4" in e["color_on_dielectric"]
    assert "g:pdf-value" in e["hitable_pdf_value"] and "g:random" in e["hitable_pdf_generate"]


def test_reference_scatter_samplers_camera_aabb(orc):
    """ref_scatter.json: metal / dielectric scatter closures called directly, the rejection samplers, random-to-sphere,
    the cosine / mixture pdf values, get-ray with a lens and a shutter interval, the AABB slab test."""
    import ctypes as C
    R = load("ref_scatter.json")
    lib = orc.load()
    u32, dbl, vp = C.c_uint32, C.c_double, C.c_void_p
    lib.orc_metal_scatter.argtypes = [vp, vp, dbl, u32, u32, u32, u32, vp]
    lib.orc_dielectric_scatter.argtypes = [vp, vp, dbl, dbl, C.c_int32, vp]
    lib.orc_random_in_unit_sphere.argtypes = [u32, u32, u32, u32, u32, vp]
    lib.orc_random_in_unit_disk.argtypes = [u32, u32, u32, u32, u32, vp]
    lib.orc_random_to_sphere.argtypes = [dbl, dbl, dbl, dbl, vp]
    lib.orc_cosine_pdf_value.argtypes = [vp, vp]
    lib.orc_cosine_pdf_value.restype = dbl
    out = np.zeros(3)
    seen = set()
    for k in R["metal"]:                               # material.scm:45-53
        d, n = np.asarray(k["d"], np.float64), np.asarray(k["n"], np.float64)
        valid = lib.orc_metal_scatter(d.ctypes.data, n.ctypes.data, k["fuzz"], *k["addr"], out.ctypes.data)
        assert bool(valid) == k["valid"] and np.abs(out - k["dir"]).max() <= TOL, k
        assert k["time"] == 0.0 and k["atten"] == [0.8, 0.6, 0.2]          # Q6: make-ray drops the ray's time (0.75 here)
        seen.add(k["valid"])
    assert seen == {True, False}
    differs = 0
    for k in R["dielectric"]:                          # material.scm:76-98 (Q10: raw d in reflect / refract)
        d, n = np.asarray(k["d"], np.float64), np.asarray(k["n"], np.float64)
        lib.orc_dielectric_scatter(d.ctypes.data, n.ctypes.data, k["ref_idx"], k["xi"], 31, out.ctypes.data)
        assert k["valid"] and np.abs(out - k["dir"]).max() <= TOL * max(1.0, np.abs(k["dir"]).max()), k
        assert k["time"] == 0.0 and k["atten"] == [1.0, 1.0, 1.0]
        lib.orc_dielectric_scatter(d.ctypes.data, n.ctypes.data, k["ref_idx"], k["xi"], 31 & ~8, out.ctypes.data)
        differs += np.abs(out - k["dir"]).max() > 1e-3
    assert differs > len(R["dielectric"]) // 2         # the reference really is in Q10 mode (un-normalised d)
    for k in R["random_in_unit_sphere"]:               # util.scm:9-15
        lib.orc_random_in_unit_sphere(*k["addr"], k["first_block"], out.ctypes.data)
        assert np.abs(out - k["out"]).max() <= TOL and np.dot(out, out) < 1
    for k in R["random_in_unit_disk"]:                 # util.scm:17-23
        lib.orc_random_in_unit_disk(*k["addr"], k["first_block"], out.ctypes.data)
        assert np.abs(out - k["out"]).max() <= TOL and out[2] == 0.0
    for k in R["random_to_sphere"]:                    # util.scm:46-54
        lib.orc_random_to_sphere(k["radius"], k["distance_sq"], k["r1"], k["r2"], out.ctypes.data)
        assert np.abs(out - k["out"]).max() <= TOL
    for k in R["cosine_pdf"]:                          # pdf.scm:18-23, 34-37
        w, w2, dv = (np.asarray(k[x], np.float64) for x in ("w", "w2", "direction"))
        a = lib.orc_cosine_pdf_value(w.ctypes.data, dv.ctypes.data)
        b = lib.orc_cosine_pdf_value(w2.ctypes.data, dv.ctypes.data)
        assert abs(a - k["value"]) <= TOL and abs(0.5 * a + 0.5 * b - k["mixture_value"]) <= TOL
    for cam in R["camera"]:                            # camera.scm:63-92
        a = cam["args"]
        slots = orc.make_camera(a[0], a[1], a[2], *a[3:])
        ref_slots = np.concatenate([np.atleast_1d(np.asarray(x, np.float64)) for x in cam["slots"]])
        assert np.abs(slots - ref_slots).max() <= TOL * max(1.0, np.abs(ref_slots).max())
        from scheme_raytrace_b200.host import camera as hc
        c = hc.make_camera(a[0], a[1], a[2], *a[3:])
        S = orc.OracleScene(g.make_scene([g.make_sphere((0, 0, 0), 1, m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5))))], c, scenes.sky_color), quantise=False)
        for k in cam["rays"]:
            mine = S.get_ray(k["s"], k["t"], k["xi_time"], *k["addr"])
            assert np.abs(mine - k["ray"]).max() <= TOL * max(1.0, np.abs(k["ray"]).max()), k
        assert len({r["ray"][6] for r in cam["rays"]}) > 10            # the shutter draw varies
    hits = 0
    for k in R["aabb"]:                                # geometry.scm:73-105 (Q11)
        got = orc.aabb_hit(k["bmin"], k["bmax"], k["o"] + k["d"] + [0.0], k["t_min"], k["t_max"])
        assert got == k["hit"], k
        hits += got
    assert 10 < hits < len(R["aabb"]) - 10


@pytest.mark.parametrize("idx", [0, 1, 2, 3, 4])
def test_reference_trace_all(orc, idx, tmp_path):
    """main.scm's trace-all, run by the reference on 10 x 10 x 2 spp with random-real returning the
    oracle's Philox draws: the oracle's render must reproduce every pixel's radiance sum and 8-bit value."""
    run = load("ref_color.json")["runs"][idx]
    w, h, spp = run["width"], run["height"], run["spp"]
    assert run["trace_line_equals_trace_all"]                        # the generator also ran the viewer's row-by-row trace-line (main.scm:452-469): same frame
    scene = host_scene(run["scene"], w, h)
    S = orc.OracleScene(scene, quantise=False)
    img, nrays = S.render(w, h, spp, max_depth=run["max_depth"], seed=run["seed"], nthreads=1)
    raw = np.asarray(run["raw_data"], np.float64).reshape(h, w, 3)
    assert raw.max() > 0 and nrays > w * h * spp
    err = np.abs(img - raw) / np.maximum(np.abs(raw), 1.0)
    assert err.max() <= 1e-9, (err.max(), np.unravel_index(err.argmax(), err.shape))
    img8 = orc.resolve(img, spp)                                    # y = 0 bottom row, like *image* (main.scm:484-488)
    ref8 = np.asarray(run["image"], np.int64).reshape(h, w, -1)[..., :3]
    defined = ref8 >= 0                                             # -1: negative radiance sum, a Gauche error upstream; clamped to 0 here
    assert np.array_equal(img8.astype(np.int64)[defined], ref8[defined]) and defined.mean() > 0.9
    assert np.all(img8[~defined] == 0) and np.all(raw[~defined.all(axis=2)].min(axis=1) < 0)
    if run["ppm"] is None:
        assert idx == 2
        return
    path = str(tmp_path / "test.ppm")                               # main.scm:439-450 save-as-ppm: the file the reference wrote, byte for byte
    assert orc.save_ppm(path, img8) == 0
    assert open(path).read() == run["ppm"] and run["ppm"].startswith(f"P3\n {w} {h}\n255\n")
    import scheme_raytrace_b200 as srt                              # the product's writer (a host function of libsrt.so: no GPU involved)
    srt.save_as_ppm(str(tmp_path / "lib.ppm"), ref8.astype(np.uint8))
    assert open(str(tmp_path / "lib.ppm")).read() == run["ppm"]


@pytest.mark.skipif(not os.path.isdir(os.environ.get("SRT_REFERENCE", "/root/reference")), reason="the reference sources only exist in the build container")
def test_committed_files_match_a_fresh_run_of_the_reference():
    """Re-executes the reference (minischeme) for the cheap tables and compares with the committed JSON."""
    import sys
    import threading
    from tests.golden import make_reference_golden as mk
    out = {}

    def work():
        sys.setrecursionlimit(200000)
        rng = mk.ScriptedRng()
        rng.script = [float(x) for x in np.random.RandomState(3).random_sample(256 + 3 * 256 + 3 * 255)]
        ref = mk.Ref(rng)
        main_mod = ref.load_main(mk.MAIN_NAMES)
        out["textures"] = mk.make_textures(ref)
        out["materials"] = mk.make_materials(ref, main_mod, rng)
        name, spec = mk.PRIM_CASES[0]
        rays = mk.rays_for(spec, 200, 100)
        out["sphere"] = mk.hits_table(ref, ref.build(spec), rays)
    threading.stack_size(256 * 1024 * 1024)
    th = threading.Thread(target=work)
    th.start()
    th.join()
    threading.stack_size(0)
    assert json.loads(json.dumps(out["textures"])) == load("ref_textures.json")
    assert json.loads(json.dumps(out["materials"])) == load("ref_materials.json")
    gold = load("ref_prims.json")["cases"][0]
    for k in ("rays", "hit", "t", "p", "n", "uv"):
        assert json.loads(json.dumps(out["sphere"][k])) == gold[k], k


# --------------------------------------------------------------------------------------------------
# converged images rendered by the reference itself (tests/golden/make_reference_render.py)
def render_stats(gold, key, mine_mean, mine_spp):
    """Compares a converged render (`mine_mean`, linear radiance per sample) with the reference's own CPU render of
    the same scene (independent random numbers).  The reference render carries its per-pixel sum and sum of squares,
    so the difference is judged against ITS Monte-Carlo standard error: z = diff / se per channel value.  `bias_z` is the
    image-wide sum of the differences over its standard error, per colour channel (the three channels of a pixel share
    their paths, so they are not independent observations), the largest of the three."""
    w, h, n = (int(x) for x in gold[key + "_meta"])
    s1, s2 = gold[key + "_sum"], gold[key + "_sumsq"]
    mean = s1 / n
    var = np.maximum(s2 / n - mean ** 2, 0.0) * n / (n - 1)
    se = np.sqrt(var / n * (1.0 + n / mine_spp))                 # both renders are noisy; ours has mine_spp samples
    ok = se > 1e-12                                              # (a pixel that only ever saw the black part of a scene)
    diff = mine_mean - mean
    z = diff[ok] / se[ok]
    a, b = np.minimum(mine_mean, 1.0), np.minimum(mean, 1.0)
    rmse = float(np.sqrt(np.mean((a - b) ** 2)))
    a8, b8 = np.floor(255.99 * np.sqrt(a)), np.floor(255.99 * np.sqrt(b))       # main.scm:123-124, 481-487
    psnr8 = float(10 * np.log10(255.0 ** 2 / max(np.mean((a8 - b8) ** 2), 1e-12)))
    return dict(median_abs_z=float(np.median(np.abs(z))), frac_within_3=float(np.mean(np.abs(z) < 3)), frac_within_4=float(np.mean(np.abs(z) < 4)),
                bias_z=max((float((diff[..., c] * ok[..., c]).sum() / np.sqrt(((se[..., c] * ok[..., c]) ** 2).sum())) for c in range(3)), key=abs),
                rel_mean=float(mine_mean.mean() / mean.mean()),
                rmse=rmse, expected_rmse=float(np.sqrt(np.mean(np.minimum(se, 1.0) ** 2))), psnr8=psnr8, n=n, width=w, height=h)


RENDER_SCENES = {"cornell_box": "cornell-box", "test_bezier": "test-bezier", "cornell_smoke": "cornell-smoke", "test_scene2": "test-scene2"}
# (fraction within 3 se, within 4 se) required per scene.  The per-value z statistic leans on the reference render's SAMPLE variance;
# in a scene lit only by small emitters (test-scene2: black sky, two lights, 1792 spp) many pixels have seen few bright paths and
# under-estimate their standard error, so the tails are heavy (cornell-smoke needed 768 spp to meet the common bars).
# The image-wide figures (bias, mean radiance, RMSE against the predicted RMSE) do not suffer from that and keep the same bars.
RENDER_TAILS = {"cornell_box": (0.975, 0.995), "test_bezier": (0.975, 0.995), "cornell_smoke": (0.975, 0.995), "test_scene2": (0.92, 0.94)}
# the 32 x 32 renders of ref_render32.npz meet the same bars.  cornell-smoke needed 1152 spp for it (three batches of 384): the
# free-flight paths through the media are rare and bright, so at few samples a pixel's SAMPLE variance under-estimates its
# standard error (within 3 / 4 se - 384 spp: 97.6 % / 98.7 %; 768 spp: 98.2 % / 99.3 %; 1152 spp: 99.1 % / 99.7 %)
RENDER_TAILS32 = dict(RENDER_TAILS)


def render_tails(fname, key):
    return (RENDER_TAILS32 if fname == "ref_render32.npz" else RENDER_TAILS)[key]


def _render_cases():
    out = [("ref_render.npz", k) for k in RENDER_SCENES]
    big = os.path.join(GOLD, "ref_render32.npz")            # round 2: 32 x 32 thumbnails (make_reference_render.py, SRT_RENDER_SIZE=32)
    if os.path.exists(big):
        out += [("ref_render32.npz", k[:-5]) for k in np.load(big).files if k.endswith("_meta")]
    return out


@pytest.mark.parametrize("fname,key", _render_cases())
def test_converged_image_against_the_references_own_render(orc, fname, key):
    """North star, second criterion: the converged image must match the reference's CPU render.  The oracle renders the
    same scene at 4096 spp with its own Philox stream; tolerance, stated: per channel value the difference is within the
    reference render's Monte-Carlo standard error - median |z| in [0.45, 0.95] (0.674 for pure noise), >= 97.5 % of the
    values within 3 se and >= 99.5 % within 4 se (looser, stated in RENDER_TAILS, for the two scenes lit only by small emitters), no global bias beyond 4 se of the summed image (per channel), mean radiance within 2 %, RMSE of the clamped
    linear image <= 1.3 x the noise-predicted RMSE.  Without Q15 (quirks = 15: one cosine direction per scatter instead
    of the three the `local` macro evaluates) the Cornell box FAILS the same test - the quirk is visible in the image."""
    gold = np.load(os.path.join(GOLD, fname))
    w, h, n = (int(x) for x in gold[key + "_meta"])
    S = orc.OracleScene(host_scene(RENDER_SCENES[key], w, h), quantise=False)
    spp = 4096
    img, _ = S.render(w, h, spp, max_depth=100, seed=77)
    st = render_stats(gold, key, img / spp, spp)
    print(f"\n[reference render {fname}:{key} {w}x{h}@{n}] {st}")
    assert 0.45 <= st["median_abs_z"] <= 0.95 and st["frac_within_3"] >= render_tails(fname, key)[0] and st["frac_within_4"] >= render_tails(fname, key)[1]
    assert abs(st["bias_z"]) <= 4.0 and abs(st["rel_mean"] - 1.0) <= 0.02
    assert st["rmse"] <= 1.3 * st["expected_rmse"]
    if key == "cornell_box" and fname == "ref_render.npz":
        img15, _ = S.render(w, h, spp, max_depth=100, seed=77, quirks=15)
        st15 = render_stats(gold, key, img15 / spp, spp)
        print(f"[reference render {key}, quirks=15] {st15}")
        assert abs(st15["bias_z"]) > 4.0 or st15["frac_within_3"] < 0.975 or st15["median_abs_z"] > 0.95
