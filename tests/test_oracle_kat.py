"""Pins the CPU oracle against the known-answer vectors of SURVEY.md §8c (KAT1-10, derived from the
reference sources; KAT1-6 are verifiable by hand) and the Random123 Philox vectors.  The reference
itself ships no tests or golden vectors, so these are the only anchors ("parity unpinned")."""
import numpy as np
import pytest
from scheme_raytrace_b200.host import geometry as g, material as m, texture as t, vec as v, camera as cam, scenes

MAXF = 999999999999.0
LAMB = m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5)))


def _scene(objs):
    return g.make_scene(objs, scenes.default_camera(), scenes.sky_color)


def test_philox_kat(orc):
    # Random123 kat_vectors: philox4x32 10 rounds
    assert list(orc.philox4x32_10([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(orc.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(orc.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = orc.rng_block(1, 2, 3, 4, 5)
    assert np.all((u > 0) & (u < 1))
    assert np.all(u * 2 ** 23 - 0.5 == np.round(u * 2 ** 23 - 0.5))  # 23-bit grid
    top = (2 ** 23 - 1 + 0.5) / 2 ** 23                               # largest uniform: exact in fp32 and < 1
    assert np.float32(top) == top and np.float32(top) < 1.0 and np.float32((2 ** 24 - 0.5) / 2 ** 24) == 1.0


@pytest.mark.parametrize("d,t_exp", [((0, 0, -1), 0.5), ((0, 0, -2), 0.25)])
def test_kat1_kat2_sphere(orc, d, t_exp):
    S = orc.OracleScene(quantise=False, scene=_scene([g.make_sphere((0, 0, -1), 0.5, LAMB)]))
    r = S.trace_batch([[0, 0, 0, *d, 0]])
    assert r["prim"][0] == 0 and r["t"][0] == t_exp
    assert np.allclose(r["p"][0], (0, 0, -0.5)) and np.allclose(r["n"][0], (0, 0, 1))


def test_kat3_from_centre(orc):
    S = orc.OracleScene(quantise=False, scene=_scene([g.make_sphere((0, 0, -1), 0.5, LAMB)]))
    r = S.trace_batch([[0, 0, -1, 0, 1, 0, 0]])
    assert r["t"][0] == 0.5 and np.allclose(r["p"][0], (0, 0.5, -1)) and np.allclose(r["n"][0], (0, 1, 0))


def test_kat4_negative_radius(orc):
    S = orc.OracleScene(quantise=False, scene=_scene([g.make_sphere((-1, 0, -1), -0.45, LAMB)]))
    r = S.trace_batch([[-1, 0, 0, 0, 0, -1, 0]])
    assert abs(r["t"][0] - 0.55) < 1e-15 and np.allclose(r["p"][0], (-1, 0, -0.55)) and np.allclose(r["n"][0], (0, 0, -1))


def test_kat5_cameras(orc):
    c = cam.make_camera((278, 278, -800), (278, 278, 0), (0, 1, 0), 40, 1, 0, 1, 0, 1)
    assert np.allclose(c[0], (278.3639702342662, 277.6360297657338, -799), rtol=0, atol=1e-12)
    assert np.allclose(c[1], (-0.7279404685324047, 0, 0), atol=1e-15) and np.allclose(c[2], (0, 0.7279404685324047, 0), atol=1e-15)
    assert np.allclose(c[4], (0, 0, -1)) and np.allclose(c[5], (-1, 0, 0)) and np.allclose(c[6], (0, 1, 0))
    # the oracle's C++ restatement of make-camera agrees with the host mirror to the last bit
    assert np.array_equal(orc.make_camera((278, 278, -800), (278, 278, 0), (0, 1, 0), 40, 1, 0, 1, 0, 1), np.asarray(cam.camera_to_floats(c)))
    c2 = cam.make_camera((0, 5, 5), (0, 0, 0), (0, 1, 0), 40, 1, 0, 1, 0, 1)
    assert np.allclose(c2[0], (-0.36397023426620234, 4.035527398013764, 4.55025903961314), atol=1e-14)
    assert np.allclose(c2[2], (0, 0.5147316415993759, -0.5147316415993759), atol=1e-14)
    assert np.array_equal(orc.make_camera((0, 5, 5), (0, 0, 0), (0, 1, 0), 40, 1, 0, 1, 0, 1), np.asarray(cam.camera_to_floats(c2)))
    # centre rays (s = t = 0.5, zero lens offset)
    S = orc.OracleScene(quantise=False, scene=g.make_scene([g.make_sphere((0, 0, 0), 1, LAMB)], c, scenes.sky_color))
    assert np.allclose(S.get_ray(0.5, 0.5, 0.0, 1, 0, 0)[3:6], (0, 0, 1), atol=1e-12)
    S2 = orc.OracleScene(quantise=False, scene=g.make_scene([g.make_sphere((0, 0, 0), 1, LAMB)], c2, scenes.sky_color))
    assert np.allclose(S2.get_ray(0.5, 0.5, 0.0, 1, 0, 0)[3:6], (0, -0.7071067811865479, -0.7071067811865479), atol=1e-12)
    # weekend camera of cfg1 == the dead <camera> class defaults (camera.scm:12-22)
    c3 = cam.make_camera((0, 0, 0), (0, 0, -1), (0, 1, 0), 90, 2, 0, 1, 0, 1)
    assert np.allclose(c3[0], (-2, -1, -1), atol=1e-12) and np.allclose(c3[1], (4, 0, 0), atol=1e-12) and np.allclose(c3[2], (0, 2, 0), atol=1e-12)


def test_kat6_xz_rect(orc):
    S = orc.OracleScene(quantise=False, scene=_scene([g.make_xz_rect(213, 343, 227, 332, 554, LAMB)]))
    r = S.trace_batch([[278, 0, 279.5, 0, 1, 0, 0]])
    assert r["t"][0] == 554 and np.allclose(r["p"][0], (278, 554, 279.5)) and np.allclose(r["n"][0], (0, 1, 0))
    assert np.allclose(r["uv"][0], (0.5, 0.5))


BEZ = [-1, 0, -1, -0.8, 1, 1, 0.8, -1, 1, 1, 0, -1]


def _curve(tt):
    a, b, c, d = [np.array(BEZ[3 * i:3 * i + 3], float) for i in range(4)]
    return a * (1 - tt) ** 3 + b * 3 * (1 - tt) ** 2 * tt + c * 3 * (1 - tt) * tt ** 2 + d * tt ** 3


def test_kat7_bezier(orc):
    o = np.array([0, 5, 5.0])
    d = _curve(0.5) - o
    d /= np.linalg.norm(d)
    r = orc.bezier_hit(BEZ, 0.1, [*o, *d, 0])
    assert r["hit"] and r["max_depth"] == 6 and r["converge_calls"] == 39
    assert abs(r["t"] - 6.731228242402701) < 1e-12
    assert np.allclose(r["p"], (0, -0.003282549631530, 0.497045705331623), atol=1e-12)
    assert np.allclose(r["n"], (0, 0.7432941462471664, 0.6689647316224497), atol=1e-12)


def test_kat8_bezier_unnormalised(orc):
    r = orc.bezier_hit(BEZ, 0.1, [0, 5, 5, 0, -5, -4.5, 0])
    assert r["hit"] and abs(r["t"] - 6.731228242402701) < 1e-12
    assert np.allclose(r["p"], (0, -28.65614121201351, -25.290527090812155), atol=1e-10)   # Q9
    assert np.allclose(r["n"], (0, 5, 4.5))


def test_kat9_bezier_miss(orc):
    r = orc.bezier_hit(BEZ, 0.1, [0, 5, 5, 0, -0.7071067811865479, -0.7071067811865479, 0])
    assert not r["hit"] and r["max_depth"] == 6 and r["converge_calls"] == 9


def test_kat10_bezier(orc):
    o = np.array([0, 5, 5.0])
    d = _curve(0.1) - o
    d /= np.linalg.norm(d)
    r = orc.bezier_hit(BEZ, 0.1, [*o, *d, 0])
    assert r["hit"] and abs(r["t"] - 7.340737709682441) < 1e-10
    assert np.allclose(r["p"], (-0.9039654664425595, 0.19918873061589348, -0.4791867748405565), atol=1e-10)


def test_tie_rule_box_edge(orc):
    """make-box faces are inclusive-type, later rect wins an equal-t tie."""
    S = orc.OracleScene(quantise=False, scene=_scene([g.make_box((0, 0, 0), (1, 1, 1), LAMB)]))
    # ray into the exact edge shared by face 0 (xy @ z=1) and face 4 (yz @ x=1): both t = 1
    r = S.trace_batch([[2, 0.5, 2, -1, 0, -1, 0]])
    assert r["t"][0] == 1.0 and r["prim"][0] == 4
    # sphere (strict) listed after a rect at the same t loses; listed before, the rect (inclusive) wins
    S2 = orc.OracleScene(quantise=False, scene=_scene([g.make_xy_rect(-1, 1, -1, 1, -0.5, LAMB), g.make_sphere((0, 0, -1), 0.5, LAMB)]))
    assert S2.trace_batch([[0, 0, 0, 0, 0, -1, 0]])["prim"][0] == 0
    S3 = orc.OracleScene(quantise=False, scene=_scene([g.make_sphere((0, 0, -1), 0.5, LAMB), g.make_xy_rect(-1, 1, -1, 1, -0.5, LAMB)]))
    assert S3.trace_batch([[0, 0, 0, 0, 0, -1, 0]])["prim"][0] == 1


def test_instances_cornell_block(orc):
    """translate(rotate-y(box)) (geometry.scm:465-543): a ray down onto the short block's top face."""
    sc = scenes.cfg4_cornell_box(64, 64)
    S = orc.OracleScene(quantise=False, scene=sc)
    # top face centre of the short block in world space
    import math
    s, c = math.sin(math.radians(-18)), math.cos(math.radians(-18))
    px, pz = 82.5, 82.5
    wx, wz = c * px + s * pz + 130, -s * px + c * pz + 65
    r = S.trace_batch([[wx, 500, wz, 0, -1, 0, 0]])
    assert r["prim"][0] == 8 and abs(r["t"][0] - 335) < 1e-9          # xz-rect @ y=165 is the 3rd box face -> id 6+2
    assert np.allclose(r["n"][0], (0, 1, 0), atol=1e-12) and np.allclose(r["uv"][0], (0.5, 0.5), atol=1e-9)


def test_aabb_quirk_q11(orc):
    # per-axis independent slabs: a ray that misses a true box can still pass (looser than standard)
    assert orc.aabb_hit((0, 0, 0), (1, 1, 1), [0.5, 0.5, -1, 0, 0, 1, 0], 0.001, MAXF)
    assert not orc.aabb_hit((0, 0, 0), (1, 1, 1), [2, 0.5, -1, 0, 0, 1, 0], 0.001, MAXF)
    assert orc.aabb_hit((0, 0, 0), (1, 1, 1), [-1, 2.5, 0.5, 1, -1, 0, 0], 0.001, MAXF)   # true slab test would miss


def test_perlin_q4_and_textures(orc):
    S = orc.OracleScene(quantise=False, scene=scenes.test_scene2(32, 32))
    p = np.random.RandomState(0).uniform(-5, 5, (64, 3))
    n_ref, n_fix = S.noise(p, quirks=15), S.noise(p, quirks=0)
    assert np.all(np.abs(n_ref) < 2) and not np.allclose(n_ref, n_fix)
    tb = S.noise(p, quirks=15, turb=True)
    assert np.all(tb >= 0)
    # marble = 0.5*(1+sin(sc*z + 10*turb)) (texture.scm:30-34); texture 0 of test-scene2 is marble(1)
    uvp = np.concatenate([np.zeros((64, 2)), p], axis=1)
    assert np.allclose(S.tex_value(0, uvp)[:, 0], 0.5 * (1 + np.sin(p[:, 2] + 10 * tb)))


def test_resolve_and_ppm(orc, tmp_path):
    rgb = np.zeros((2, 3, 3))
    rgb[0, 0] = (4.0, 1.0, 0.25)          # spp 4 -> (1, .25, .0625) -> sqrt -> (1, .5, .25)
    rgb[1, 2] = (100.0, 0.0, 4.0)
    img = orc.resolve(rgb, 4)
    assert tuple(img[0, 0]) == (255, 127, 63) and tuple(img[1, 2]) == (255, 0, 255)
    path = tmp_path / "test.ppm"
    orc.save_ppm(path, img)
    lines = path.read_text().split("\n")
    assert lines[0] == "P3" and lines[1] == " 3 2" and lines[2] == "255"   # note the leading space (main.scm:442)
    assert lines[3] == "0 0 0" and lines[5] == "255 0 255" and lines[6] == "255 127 63"   # top row first (y flipped)


def test_render_cfg1_smoke(orc):
    sc = scenes.cfg1_weekend(40, 20)
    S = orc.OracleScene(sc)
    a, nrays = S.render(40, 20, 4, max_depth=50, seed=1)
    b, _ = S.render(40, 20, 4, max_depth=50, seed=1, nthreads=1)
    assert np.array_equal(a, b)                     # counter-based RNG: thread-count independent
    assert nrays > 40 * 20 * 4 and np.all(np.isfinite(a)) and a.min() >= 0
    top = a[-1].mean(axis=0) / 4                    # sky gradient at the top rows is bluish
    assert top[2] > top[0]


def test_patch_flat_matches_plane(orc):
    """Bicubic patch (north-star extension): a flat control net must intersect like the plane."""
    from scheme_raytrace_b200.host import bezier as bz
    cp = [[(-1 + i, -1 + j, -2.0) for j in range(4)] for i in range(4)]
    S = orc.OracleScene(quantise=False, scene=_scene([bz.make_bezier_patch(cp, LAMB)]))
    r = S.trace_batch([[0, 0, 0, 0, 0, -1, 0], [0.3, 0.7, 1, 0.1, -0.2, -2, 0], [5, 5, 0, 0, 0, -1, 0], [0, 0, -5, 0, 0, 1, 0]])
    assert list(r["prim"]) == [0, 0, -1, 0] and np.allclose(r["t"][[0, 1, 3]], (2.0, 1.5, 3.0), atol=1e-9)
    assert np.allclose(r["uv"][0], (1 / 3, 1 / 3), atol=1e-9) and np.allclose(r["n"][0], (0, 0, 1), atol=1e-9) and np.allclose(r["n"][3], (0, 0, -1), atol=1e-9)


def test_patch_revolved_profile_is_round(orc):
    """A revolved straight profile (cylinder r = 2): hits lie on the cylinder within the cubic
    circle approximation error (2.7e-4 relative)."""
    pts = scenes.revolve_profile([(2.0, 0.0), (2.0, 1.0), (2.0, 2.0), (2.0, 3.0)], LAMB)
    S = orc.OracleScene(quantise=False, scene=_scene(pts))
    rs = np.random.RandomState(3)
    ang = rs.uniform(0, 2 * np.pi, 200)
    rays = np.stack([6 * np.cos(ang), rs.uniform(0.2, 2.8, 200), 6 * np.sin(ang), -np.cos(ang), 0 * ang, -np.sin(ang), 0 * ang], axis=1)
    r = S.trace_batch(rays)
    assert np.all(r["prim"] >= 0)
    rad = np.hypot(r["p"][:, 0], r["p"][:, 2])
    assert np.max(np.abs(rad - 2.0)) < 2.0 * 3e-4 and np.allclose(r["t"], 6 - rad, atol=1e-9)


def test_image_texture_lookup(orc):
    """texture.scm:36-50: i = u*nx, j = (1-v)*ny - 0.001, clamped to [0, n-1], texel = floor; /255.
    Hand-checked on a 4x2 image (row 0 = top)."""
    from scheme_raytrace_b200.host import geometry as g, material as m, texture as t, vec as v, scenes
    data = np.arange(4 * 2 * 3) * 10            # texel (i, j) = 10 * (3*i + 12*j + c)
    img = t.image_texture(data, 4, 2)
    scene = g.make_scene([g.make_sphere(v.vec3(0, 0, -1), 0.5, m.make_diffuse_light(img))], scenes.default_camera(), scenes.black)
    S = orc.OracleScene(scene, quantise=False)
    uvp = np.array([[0.0, 1.0, 0, 0, 0],        # i=0, j=-0.001 -> 0      : texel (0,0)
                    [0.3, 0.9, 0, 0, 0],        # i=1.2, j=0.199          : texel (1,0)
                    [0.99, 0.4, 0, 0, 0],       # i=3.96 -> clamp 3, j=1.199 -> clamp 1: texel (3,1)
                    [0.0, 0.0, 0, 0, 0],        # lambertian's u=v=0: i=0, j=1.999 -> clamp 1: texel (0,1)
                    [2.0, -1.0, 0, 0, 0]])      # beyond the edges: texel (3,1)
    got = S.tex_value(0, uvp, 15)
    exp = np.array([[0, 10, 20], [30, 40, 50], [210, 220, 230], [120, 130, 140], [210, 220, 230]]) / 255.0
    assert np.allclose(got, exp, atol=1e-15)
    with pytest.raises(ValueError):
        t.image_texture([0] * 5, 2, 1)


def test_material_helpers_hand_derived(orc):
    """material.scm:41-43 reflect, :59-67 refract (with and without Q10), :69-74 schlick - values
    worked out by hand from the source lines."""
    assert np.allclose(orc.reflect([1, -1, 0], [0, 1, 0]), [1, 1, 0], atol=1e-15)          # v - 2 (v.n) n
    assert np.allclose(orc.reflect([0, -3, 0], [0, 1, 0]), [0, 3, 0], atol=1e-15)          # length is kept (Q10: d unnormalised)
    # head-on, v = (0,-2,0) unnormalised, n = (0,1,0), ni/nt = 1/1.5: uv = (0,-1,0), dt = -1, disc = 1
    #   upstream (Q10): r (v - n dt) - n sqrt(disc) = (0,-2+1,0)/1.5 - (0,1,0) = (0,-5/3,0)
    #   book:           r (uv - n dt) - n           = (0,0,0)/1.5    - (0,1,0) = (0,-1,0)
    ok, r = orc.refract([0, -2, 0], [0, 1, 0], 1 / 1.5, quirks=15)
    assert ok and np.allclose(r, [0, -5.0 / 3.0, 0], atol=1e-15)
    ok, r = orc.refract([0, -2, 0], [0, 1, 0], 1 / 1.5, quirks=0)
    assert ok and np.allclose(r, [0, -1, 0], atol=1e-15)
    # 45 degrees into glass: uv = (1,-1,0)/sqrt2, dt = -1/sqrt2, disc = 1 - (1/2.25)(1/2) = 7/9;
    #   book: r (uv - n dt) - n sqrt(disc) = (1/sqrt2/1.5, 0, 0) - (0, sqrt7/3, 0): Snell sin(t) = sin(45)/1.5
    ok, r = orc.refract([1, -1, 0], [0, 1, 0], 1 / 1.5, quirks=0)
    assert ok and np.allclose(r, [1 / np.sqrt(2) / 1.5, -np.sqrt(7) / 3, 0], atol=1e-15) and abs(np.linalg.norm(r) - 1) < 1e-15
    # grazing from inside glass (ni/nt = 1.5): disc = 1 - 2.25 (1 - dt^2) < 0 -> total internal reflection
    ok, _ = orc.refract([1, -0.1, 0], [0, 1, 0], 1.5)
    assert not ok
    # ni/nt = 1, dt = 0 exactly (v perpendicular to n): disc = 1 - 1*(1 - 0) = 0 is NOT refracted (`(> discriminant 0)`)
    ok, _ = orc.refract([1, 0, 0], [0, 1, 0], 1.0)
    assert not ok
    assert abs(orc.schlick(1.0, 1.5) - 0.04) < 1e-15            # r0 = ((1-1.5)/(1+1.5))^2
    assert abs(orc.schlick(0.0, 1.5) - 1.0) < 1e-15
    assert abs(orc.schlick(0.5, 1.5) - (0.04 + 0.96 / 32)) < 1e-15


def test_sampling_helpers_hand_derived(orc):
    """util.scm:37-44 random-cosine-direction through onb.scm:8-36 (Q1: x, y scaled by 2) and the
    sky gradient main.scm:91-95, by hand."""
    lib = orc.load()
    out = np.zeros(3)
    n = np.array([0.0, 0.0, 1.0])                     # w = n, a = (1,0,0), v = unit(w x a) = (0,1,0), u = w x v = (-1,0,0)
    for quirks, k in ((15, 2.0), (0, 1.0)):
        lib.orc_onb_cosine(orc._p(n), 0.25, 0.36, quirks, orc._p(out))      # phi = pi/2: x = 0, y = k*0.6, z = sqrt(1-0.36) = 0.8
        assert np.allclose(out, [-0.0 * k, 0.6 * k, 0.8], atol=1e-15), (quirks, out)
        lib.orc_onb_cosine(orc._p(n), 0.0, 0.25, quirks, orc._p(out))       # phi = 0: x = k*0.5 -> along u = (-1,0,0)
        assert np.allclose(out, [-0.5 * k, 0.0, np.sqrt(0.75)], atol=1e-15), (quirks, out)
    # Q15 (onb.scm:27-36): (local uvw (random-cosine-direction)) evaluates its operand three times; draws in call order
    # (r1, r2), (r1', r2'), (r1'', r2''): x = cos(0)*2*sqrt(.25) = 1 of the 1st, y = sin(pi/2)*2*sqrt(.36) = 1.2 of the 2nd,
    # z = sqrt(1-.64) = .6 of the 3rd -> u*1 + v*1.2 + w*.6
    r6 = np.array([0.0, 0.25, 0.25, 0.36, 0.7, 0.64])
    lib.orc_onb_local_cosine(orc._p(n), orc._p(r6), 31, orc._p(out))
    assert np.allclose(out, [-1.0, 1.2, 0.6], atol=1e-15), out
    lib.orc_onb_local_cosine(orc._p(n), orc._p(r6), 15, orc._p(out))         # without the quirk: the first evaluation alone
    assert np.allclose(out, [-1.0, 0.0, np.sqrt(0.75)], atol=1e-15), out
    from scheme_raytrace_b200.host import geometry as g, scenes
    S = orc.OracleScene(g.make_scene([], scenes.default_camera(), scenes.sky_color), quantise=False)
    assert np.allclose(S.sky([0, 1, 0]), [0.5, 0.7, 1.0], atol=1e-15)       # t = 1
    assert np.allclose(S.sky([0, -2, 0]), [1.0, 1.0, 1.0], atol=1e-15)      # t = 0 (direction is normalised first)
    assert np.allclose(S.sky([3, 0, 0]), [0.75, 0.85, 1.0], atol=1e-15)     # t = 0.5


def _medium_scene(R, rho):
    from scheme_raytrace_b200.host import geometry as g, material as m, texture as t, vec as v, scenes
    white = m.make_lambertian(t.constant_texture(v.vec3(1, 1, 1)))
    med = g.make_constant_medium(g.make_sphere(v.vec3(0, 0, 0), R, white), rho, t.constant_texture(v.vec3(1, 1, 1)))
    return g.make_scene([med], scenes.default_camera(), scenes.black)


def test_constant_medium_free_flight_law(orc):
    """geometry.scm:545-578: hit distance inside the boundary = -log(xi)/density, accepted if it ends
    before the exit.  Along a diameter of length L (|d| = 2): P(hit) = 1 - exp(-rho L) and
    E[distance | hit] = 1/rho - L exp(-rho L)/(1 - exp(-rho L)), from the exponential law."""
    R, rho, n = 2.0, 0.3, 200000
    S = orc.OracleScene(_medium_scene(R, rho), quantise=False)
    rays = np.tile(np.array([0, 0, -10, 0, 0, 2.0, 0.0]), (n, 1))
    o = S.trace_batch(rays)
    hit = o["prim"] >= 0
    L, q = 2 * R, np.exp(-rho * 2 * R)
    assert abs(hit.mean() - (1 - q)) < 4 * np.sqrt(q * (1 - q) / n)
    dist = (o["t"][hit] - 4.0) * 2.0                      # enters at z = -2: t = 4; |d| = 2
    assert dist.min() > 0 and dist.max() < L
    assert abs(dist.mean() - (1 / rho - L * q / (1 - q))) < 0.01


FURNACE_LE = 0.75


def _furnace_cases():
    """A unit sphere of material X at the centre of an emitting enclosure (sphere of radius -100:
    the negative radius turns the normal inward, geometry.scm:146-175, so n.d < 0 and the
    diffuse-light emits, material.scm:103-111).  A convex sphere is never re-hit, so every path is
    camera -> [sphere ->] emitter and `color` (main.scm:100-121) gives exactly Le off the sphere and
    weight_X * Le on it: albedo for lambertian (atten * spdf / pdf = albedo for a unit normal, with or
    without Q1) and for fuzz-0 metal, 1 for the dielectric."""
    from scheme_raytrace_b200.host import geometry as g, material as m, texture as t, vec as v, scenes, camera as cam
    light = m.make_diffuse_light(t.constant_texture(v.vec3(FURNACE_LE, FURNACE_LE, FURNACE_LE)))
    c = cam.make_camera(v.vec3(0, 0, 5), v.vec3(0, 0, 0), v.vec3(0, 1, 0), 40, 1.0, 0.0, 5.0, 0.0, 1.0)
    cases = (("lambertian", m.make_lambertian(t.constant_texture(v.vec3(0.5, 0.25, 0.8))), (0.5, 0.25, 0.8)),
             ("metal", m.make_metal(t.constant_texture(v.vec3(0.9, 0.6, 0.3)), 0.0), (0.9, 0.6, 0.3)),
             ("dielectric", m.make_dielectric(1.5), (1.0, 1.0, 1.0)))
    for name, mat, w in cases:
        objs = [g.make_sphere(v.vec3(0, 0, 0), -100.0, light), g.make_sphere(v.vec3(0, 0, 0), 1.0, mat)]
        yield name, g.make_scene(objs, c, scenes.black), np.asarray(w) * FURNACE_LE


def check_furnace(img, expected, tol):
    on = img[12:20, 12:20].reshape(-1, 3)            # pixels whose footprint lies on the sphere
    off = np.concatenate([img[:4].reshape(-1, 3), img[-4:].reshape(-1, 3)])
    assert np.abs(on - expected).max() < tol, np.abs(on - expected).max()
    assert np.abs(off - FURNACE_LE).max() < tol


@pytest.mark.parametrize("quirks", [31, 15, 0])
def test_furnace_analytic_radiance(orc, quirks):
    for name, scene, expected in _furnace_cases():
        S = orc.OracleScene(scene, quantise=False)
        img, _ = S.render(32, 32, 4, max_depth=50, seed=9, quirks=quirks)
        check_furnace(img / 4, expected, 1e-12)


def test_utah_teapot_data_and_seams(orc):
    """The complete 32-patch Utah teapot (`scenes.utah_teapot`): 127 control points / 10 patches of
    the compact Newell data set; mirrored copies join on the symmetry planes; the known extent of
    the pot (spout tip x = 3.525 / handle x = -3 / knob z = 3.15, in teapot units); rays aimed at
    the handle and the spout hit handle and spout patches on the oracle."""
    from scheme_raytrace_b200.host import geometry as g
    assert len(scenes.TEAPOT_CP) == 127 and len(scenes.TEAPOT_PATCHES) == 10
    assert all(len(idx) == 16 and max(idx) < 127 for _, _, idx in scenes.TEAPOT_PATCHES)
    used = sorted({i for _, _, idx in scenes.TEAPOT_PATCHES for i in idx})
    assert used == list(range(127))                                  # every control point belongs to a patch
    cp = np.asarray(scenes.TEAPOT_CP)
    for part, copies, idx in scenes.TEAPOT_PATCHES:
        net = cp[list(idx)].reshape(4, 4, 3)
        e0, e1 = net[:, 0], net[:, 3]                                # the v = 0 and v = 1 edges lie in mirror planes
        if copies == 4:                                              # quadrant patch: one edge in y = 0, the other in x = 0
            assert (np.all(e0[:, 1] == 0) and np.all(e1[:, 0] == 0)) or (np.all(e0[:, 0] == 0) and np.all(e1[:, 1] == 0)), part
        else:                                                        # handle / spout halves: both edges in y = 0
            assert np.all(e0[:, 1] == 0) and np.all(e1[:, 1] == 0), part
    pot = scenes.utah_teapot(LAMB)
    assert len(pot) == 32 and all(o.kind == g.PATCH for o in pot)
    allcp = np.asarray([o.params for o in pot]).reshape(-1, 3)      # world frame: (x, z_teapot, -y_teapot)
    assert np.allclose(allcp.min(axis=0), (-3.0, 0.0, -2.0)) and np.allclose(allcp.max(axis=0), (3.525, 3.15, 2.0))
    parts = [part for part, copies, _ in scenes.TEAPOT_PATCHES for _ in range(copies)]
    S = orc.OracleScene(quantise=False, scene=_scene(pot))
    rays = [[-2.9, 1.8, 5, 0, 0, -1, 0],         # through the handle loop's outer arc
            [2.5, 1.6, 5, 0, 0, -1, 0],          # through the spout
            [0.0, 1.2, 5, 0, 0, -1, 0],          # body, front
            [0.0, 5.0, 0.1, 0, -1, 0, 0],        # straight down onto the lid knob (degenerate pole row)
            [0.3, -5.0, 0.2, 0, 1, 0, 0],        # straight up onto the bottom
            [-2.2, 1.6, 5, 0, 0, -1, 0]]         # through the hole of the handle: miss
    r = S.trace_batch(rays)
    assert [parts[i] if i >= 0 else None for i in r["prim"]] == ["handle", "spout", "body", "lid", "bottom", None]
    assert np.all(np.isfinite(r["n"])) and np.allclose(np.linalg.norm(r["n"][:5], axis=1), 1.0)
    assert abs(r["p"][2][2] - 2.0) < 0.04 and r["n"][2][2] > 0.95    # body radius just under 2 near the equator
    assert 0 <= r["p"][4][1] < 0.01 and np.allclose(r["n"][4], (0, -1, 0), atol=0.05)   # nearly flat bottom at height 0
