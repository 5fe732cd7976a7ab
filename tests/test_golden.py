"""CPU (-m "not gpu"): the oracle against the COMMITTED golden fixtures of tests/golden/.

`kat_survey.json` holds the known-answer vectors of SURVEY.md §8c (not oracle output: they pin the
oracle); `rays_*.npz` / `image_*.npz` are frozen oracle outputs on seeded inputs (they catch drift of
the oracle, of the scene generators and of the flattening)."""
import glob
import json
import os
import numpy as np
import pytest
from scheme_raytrace_b200.host import geometry as g, material as m, texture as t, camera as cam, scenes
from scheme_raytrace_b200.host.flatten import flatten_scene
from tests.golden import make_golden as mg

GOLD = os.path.dirname(os.path.abspath(mg.__file__))
LAMB = m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5)))
KAT = json.load(open(os.path.join(GOLD, "kat_survey.json")))


def _scene(objs):
    return g.make_scene(objs, scenes.default_camera(), scenes.sky_color)


def bezier_kat_ray(case):
    """The ray of a Bezier known-answer case (origin (0,5,5), aimed at a curve point or given raw)."""
    bz = KAT["bezier"]
    o = np.asarray(bz["origin"], float)
    if "raw_dir" in case:
        d = np.asarray(case["raw_dir"], float)
    else:
        if "aim" in case:
            aim = np.asarray(case["aim"], float)
        else:
            a, b, c, e = (np.asarray(q, float) for q in bz["cp"])
            s = case["aim_param"]
            aim = a * (1 - s) ** 3 + 3 * b * (1 - s) ** 2 * s + 3 * c * (1 - s) * s ** 2 + e * s ** 3
        d = aim - o
        if case["normalise"]:
            d = d / np.linalg.norm(d)
    return np.concatenate([o, d, [0.0]])


def test_fixture_files_present():
    names = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLD, "*.npz")) if not os.path.basename(f).startswith("ref_"))   # ref_*: outputs of the reference itself, tests/test_reference_golden.py
    assert names == sorted([f"rays_{n}.npz" for n in mg.RAY_SCENES] + [f"image_{n}.npz" for n in mg.IMAGE_SCENES])
    assert KAT == json.loads(json.dumps(mg.KAT_SURVEY))           # the committed file is what the generator holds


def test_oracle_against_survey_kats(orc):
    for k in KAT["sphere"]:
        S = orc.OracleScene(quantise=False, scene=_scene([g.make_sphere(k["center"], k["radius"], LAMB)]))
        r = S.trace_batch([k["ray"]], KAT["t_min"], KAT["t_max"])
        assert r["prim"][0] == 0 and abs(r["t"][0] - k["t"]) < 1e-15, k["kat"]
        assert np.allclose(r["p"][0], k["p"], atol=1e-15) and np.allclose(r["n"][0], k["n"], atol=1e-12), k["kat"]
    for k in KAT["camera"]:
        c = cam.make_camera(*k["args"])
        assert np.allclose(c[0], k["llc"], rtol=0, atol=1e-12) and np.allclose(c[1], k["horiz"], atol=1e-14) and np.allclose(c[2], k["vert"], atol=1e-14)
        if "w" in k:
            assert np.allclose(c[4], k["w"]) and np.allclose(c[5], k["u"]) and np.allclose(c[6], k["v"])
        S = orc.OracleScene(quantise=False, scene=g.make_scene([g.make_sphere((0, 0, 0), 1, LAMB)], c, scenes.sky_color))
        assert np.allclose(S.get_ray(0.5, 0.5, 0.0, 1, 0, 0)[3:6], k["centre_dir"], atol=1e-12)
    for k in KAT["xz_rect"]:
        S = orc.OracleScene(quantise=False, scene=_scene([g.make_xz_rect(*k["rect"], LAMB)]))
        r = S.trace_batch([k["ray"]])
        assert r["t"][0] == k["t"] and np.allclose(r["p"][0], k["p"]) and np.allclose(r["n"][0], k["n"]) and np.allclose(r["uv"][0], k["uv"])
    bz = KAT["bezier"]
    cps = [c for p in bz["cp"] for c in p]
    for k in bz["cases"]:
        r = orc.bezier_hit(cps, bz["width"], bezier_kat_ray(k))
        assert r["hit"] == k["hit"] and r["max_depth"] == bz["max_depth"], k["kat"]
        if "converge_calls" in k:
            assert r["converge_calls"] == k["converge_calls"], k["kat"]
        if k["hit"]:
            assert abs(r["t"] - k["t"]) < 1e-12 and np.allclose(r["p"], k["p"], atol=1e-10), k["kat"]
            if "n" in k:
                assert np.allclose(r["n"], k["n"], atol=1e-12), k["kat"]


@pytest.mark.parametrize("name", list(mg.RAY_SCENES))
def test_oracle_reproduces_golden_rays(orc, name):
    gold = np.load(os.path.join(GOLD, f"rays_{name}.npz"))
    scene, flat, rays = mg.golden_rays(name)
    assert mg.flat_digest(flat) == str(gold["digest"]), "scene generator / flattening changed: regenerate tests/golden"
    assert np.array_equal(rays, gold["rays"])
    o = orc.OracleScene(scene, flat=flat).trace_batch(rays.astype(np.float64))
    assert np.array_equal(o["prim"], gold["prim"])
    hit = gold["prim"] >= 0
    assert hit.mean() > 0.5
    for key in ("t", "p", "n", "uv"):
        assert np.allclose(o[key][hit], gold[key][hit], rtol=1e-12, atol=1e-12, equal_nan=True), key   # Q5: sphere uv is NaN for |p.y| > 1


@pytest.mark.parametrize("name", list(mg.IMAGE_SCENES))
def test_oracle_reproduces_golden_image(orc, name):
    gold = np.load(os.path.join(GOLD, f"image_{name}.npz"))
    fn = mg.IMAGE_SCENES[name][0]
    w, h, spp, seed = (int(gold[k]) for k in ("width", "height", "spp", "seed"))
    scene = fn(w, h)
    flat = flatten_scene(scene)
    assert mg.flat_digest(flat) == str(gold["digest"])
    img, nrays = orc.OracleScene(scene, flat=flat).render(w, h, spp, max_depth=int(gold["max_depth"]), seed=seed)
    assert nrays == int(gold["rays"])
    assert np.allclose(img, gold["rgb_sum"], rtol=1e-6, atol=1e-6)
