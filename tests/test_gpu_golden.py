"""GPU (-m gpu): the CUDA path, through the C-ABI, against the COMMITTED golden fixtures of
tests/golden/ — no oracle call at run time.  Bars: primitive id exact, t / normal within 1e-4
relative (rays the fixture marks as near-ties or fp32-unstable are excluded and counted); images
under the shared Philox stream within the tolerances of test_gpu_parity.test_image_same_stream."""
import json
import os
import numpy as np
import pytest
import scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import geometry as g, material as m, texture as t, bezier as bz, scenes
from tests.golden import make_golden as mg
from tests.test_golden import GOLD, KAT, bezier_kat_ray

pytestmark = pytest.mark.gpu
LAMB = m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5)))


def _one(objs, ray, **kw):
    r = srt.Renderer(g.make_scene(objs, scenes.default_camera(), scenes.sky_color), device=0)
    out = r.trace_batch(np.asarray([ray], dtype=np.float32), **kw)[0]
    r.close()
    return out


def test_survey_kats_through_c_abi():
    """KAT1-4, 6, 7-10 of SURVEY.md §8c on the CUDA path (fp32: t within 1e-6 relative)."""
    for k in KAT["sphere"]:
        h = _one([g.make_sphere(k["center"], k["radius"], LAMB)], k["ray"])
        assert h["prim"] == 0 and abs(h["t"] - k["t"]) <= 1e-6 * k["t"], k["kat"]
        assert np.allclose(h["p"], k["p"], atol=1e-6) and np.allclose(h["n"], k["n"], atol=1e-6), k["kat"]
    for k in KAT["xz_rect"]:
        h = _one([g.make_xz_rect(*k["rect"], LAMB)], k["ray"])
        assert h["prim"] == 0 and h["t"] == k["t"] and np.allclose(h["n"], k["n"]) and abs(h["u"] - 0.5) < 1e-6 and abs(h["v"] - 0.5) < 1e-6
    b = KAT["bezier"]
    curve = bz.make_bezier(*b["cp"], b["width"], LAMB)
    for k in b["cases"]:
        h = _one([curve], bezier_kat_ray(k))
        assert (h["prim"] == 0) == k["hit"], k["kat"]
        if k["hit"] and k["kat"] in (7, 8):
            # KAT7/8 aim exactly at the centre line (ray-space x = 0 in exact arithmetic), a discontinuity of
            # the subdivision's leaf choice (Q8): any rounding of the direction (KAT7: the normalised direction
            # is not representable in fp32) or of the ray-space projection (KAT8: direction (0,-5,-4.5) is
            # exact, but the fp32 rotation is not) lands on the neighbouring leaf.  The oracle shows the same:
            # f64 on the fp32-rounded KAT7 ray and the f32 oracle on KAT8 both return 6.7225183 instead of
            # 6.7312282 (the "fp32-unstable" class the ray batches filter).  Accepted: either leaf value.
            assert min(abs(h["t"] - k["t"]), abs(h["t"] - 6.722518295689991)) <= 1e-4 * k["t"], (k["kat"], h["t"])
            assert np.allclose(h["n"], k["n"], rtol=1e-5, atol=1e-5), k["kat"]
        elif k["hit"]:
            assert abs(h["t"] - k["t"]) <= 1e-4 * k["t"], (k["kat"], h["t"])
            assert np.allclose(h["p"], k["p"], rtol=1e-4, atol=2e-3), (k["kat"], h["p"])
            if "n" in k:
                assert np.allclose(h["n"], k["n"], rtol=1e-5, atol=1e-5), k["kat"]


@pytest.mark.parametrize("name", list(mg.RAY_SCENES))
def test_trace_batch_against_golden(name):
    gold = np.load(os.path.join(GOLD, f"rays_{name}.npz"))
    scene, flat, rays = mg.golden_rays(name)
    assert mg.flat_digest(flat) == str(gold["digest"])
    r = srt.Renderer(scene, device=0)
    gp = r.trace_batch(gold["rays"])
    r.close()
    keep = ~gold["filtered"]
    assert keep.mean() >= 0.99
    assert np.array_equal(gp["prim"][keep], gold["prim"][keep]), np.nonzero(keep & (gp["prim"] != gold["prim"]))[0][:10]
    hit = keep & (gold["prim"] >= 0)
    t_err = np.abs(gp["t"][hit] - gold["t"][hit]) / np.abs(gold["t"][hit])
    n_err = np.linalg.norm(gp["n"][hit] - gold["n"][hit], axis=1) / np.linalg.norm(gold["n"][hit], axis=1)
    print(f"\n[golden {name}] n={len(keep)} filtered={int((~keep).sum())} t_err_max={t_err.max():.2e} n_err_max={n_err.max():.2e}")
    assert t_err.max() <= 1e-4 and n_err.max() <= 1e-4


@pytest.mark.parametrize("name", list(mg.IMAGE_SCENES))
def test_image_against_golden(name):
    gold = np.load(os.path.join(GOLD, f"image_{name}.npz"))
    fn = mg.IMAGE_SCENES[name][0]
    w, h, spp, seed = (int(gold[k]) for k in ("width", "height", "spp", "seed"))
    r = srt.Renderer(fn(w, h), device=0)
    assert mg.flat_digest(r.flat) == str(gold["digest"])
    img, st = r.render(w, h, spp, max_depth=int(gold["max_depth"]), seed=seed)
    r.close()
    ref = gold["rgb_sum"].astype(np.float64)
    diff = np.abs(img.astype(np.float64) - ref) / spp
    a8 = srt.correct_gamma_quantise(img, spp).astype(np.float64)
    b8 = srt.correct_gamma_quantise(gold["rgb_sum"], spp).astype(np.float64)
    mse = np.mean((a8 - b8) ** 2)
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    print(f"\n[golden image {name}] rays gpu={st.rays} golden={int(gold['rays'])} median={np.median(diff):.2e} within1e-2={np.mean(diff < 1e-2):.4f} psnr8={psnr:.1f} dB")
    assert abs(st.rays - int(gold["rays"])) <= 0.02 * int(gold["rays"])
    assert np.median(diff) < 1e-4 and np.mean(diff < 1e-2) >= 0.97 and psnr >= 35.0
