"""GPU parity tests proper (-m gpu): every check goes through the C-ABI (libsrt.so) and compares
against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest
import scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import geometry as g, material as m, texture as t, scenes
from tests import raybatch

pytestmark = pytest.mark.gpu

SMALL = {"cfg1": (scenes.cfg1_weekend, 200, 100), "cfg2": (scenes.cfg2_random_spheres, 120, 80),
         "cfg3": (scenes.cfg3_next_week, 80, 80), "cfg4": (scenes.cfg4_cornell_box, 64, 64),
         "bezier": (scenes.test_bezier, 64, 64), "cornell_bezier": (scenes.cornell_bezier, 64, 64),
         "scene2": (scenes.test_scene2, 64, 64), "bvh100": (scenes.test_scene_bvh, 64, 64),
         "smoke": (scenes.cornell_smoke, 64, 64), "patches": (scenes.cfg5_patches, 96, 54),
         "klein": (scenes.cornell_klein, 48, 48), "image": (scenes.image_scene, 64, 64),
         "teapot": (scenes.teapot_scene, 96, 54)}


@pytest.fixture(scope="module", params=list(SMALL))
def pair(request, orc):
    fn, w, h = SMALL[request.param]
    scene = fn(w, h)
    r = srt.Renderer(scene, device=0)
    S = orc.OracleScene(scene, flat=r.flat, perlin=r.perlin)
    yield request.param, scene, r, S, w, h
    r.close()


def _tree_areas(nodes, n_items):
    """Surface-area sums of a host-built tree, as k_tree_area computes them (csrc/lbvh.cu)."""
    le, re = nodes["le"].astype(np.float64), nodes["re"].astype(np.float64)
    al = 8 * (le[:, 0] * le[:, 1] + le[:, 1] * le[:, 2] + le[:, 0] * le[:, 2])
    ar = 8 * (re[:, 0] * re[:, 1] + re[:, 1] * re[:, 2] + re[:, 0] * re[:, 2])
    if n_items <= 1:
        ar = ar * 0
    d = np.maximum(nodes["lc"][0].astype(np.float64) + le[0], nodes["rc"][0].astype(np.float64) + re[0]) - \
        np.minimum(nodes["lc"][0].astype(np.float64) - le[0], nodes["rc"][0].astype(np.float64) - re[0])
    root = 2 * (d[0] * d[1] + d[1] * d[2] + d[0] * d[2])
    a_int = al[nodes["left"] >= 0].sum() + ar[nodes["right"] >= 0].sum() + root
    a_leaf = al[nodes["left"] < 0].sum() + (ar[nodes["right"] < 0].sum() if n_items > 1 else 0.0)
    return a_int, a_leaf, root


def _global_selection(orc, types, aabb):
    """The commit rule of csrc/srt_api.cu restated: forced (Klein) + huge (>= 0.5 E) primitives, then
    the cost-minimal prefix of the outliers (>= 1.5 x median extent) by tree surface area."""
    n = len(aabb)
    klein = types == 8
    ext = (aabb[:, 3:] - aabb[:, :3]).max(axis=1).astype(np.float32)
    ext[klein] = 0
    E = np.float32((aabb[~klein, 3:].max(axis=0) - aabb[~klein, :3].min(axis=0)).max()) if (~klein).any() else np.float32(0)
    forced = np.nonzero(klein)[0].tolist()
    huge = [i for i in range(n) if not klein[i] and n > 2 and ext[i] >= np.float32(0.5) * E]
    huge.sort(key=lambda i: -ext[i])                                   # stable: ties keep the lower id first
    huge = huge[:max(8 - len(forced), 0)]
    base = huge + forced
    rest = [i for i in range(n) if i not in set(base)]
    extra = []
    if len(rest) >= 16 and len(base) < 8:
        med = np.sort(ext[rest])[len(rest) // 2]
        extra = [i for i in rest if types[i] <= 4 and ext[i] >= np.float32(1.5) * med]
        extra.sort(key=lambda i: -ext[i])
        extra = extra[:8 - len(base)]
    if not extra:
        return base, extra, 0
    costs = []
    for k in range(len(extra) + 1):
        g = set(base + extra[:k])
        items = np.array([i for i in range(n) if i not in g])
        _, _, nodes = orc.lbvh_build(aabb[items])
        a_int, a_leaf, root = _tree_areas(nodes, len(items))
        if k == 0:
            ref = root
        costs.append((a_int + 2.0 * a_leaf) / ref + 0.25 * k)
    k = int(np.argmin(costs))
    second = np.partition(costs, 1)[1] if len(costs) > 1 else np.inf
    assert second - costs[k] > 1e-9 * costs[k], ("ambiguous cost minimum: pick another test scene", costs)
    return base, extra, k


def test_lbvh_bit_exact(pair, orc):
    """GPU Morton keys / sort order / Karras topology / node boxes == sequential host reference,
    byte for byte, over the same (GPU-computed) primitive AABBs."""
    name, scene, r, S, w, h = pair
    aabb = r.prim_bounds()
    items, glob = r.bvh_items()
    assert sorted(items.tolist() + glob.tolist()) == list(range(len(aabb)))     # every surface is in the tree or global
    base, extra, k = _global_selection(orc, r.flat.prims["type"][:len(aabb)], aabb)
    assert sorted(glob.tolist()) == sorted(base + extra[:k]), (glob, base, extra, k)
    keys_g, order_g = r.bvh_keys()
    if len(items) == 0:                                                # everything is global: no tree to compare
        assert len(keys_g) == 0
        return
    nodes_g = r.bvh_nodes()
    keys_h, order_h, nodes_h = orc.lbvh_build(aabb[items])            # host reference over the same AABBs
    for side in ("left", "right", "sibling"):                         # its leaf refs are item indices -> primitive ids
        leaf = nodes_h[side] < 0
        if side == "sibling":
            leaf[0] = False                                           # the root's -1 means "none", not leaf ~0
        nodes_h[side][leaf] = ~items[~nodes_h[side][leaf]]
    assert np.array_equal(keys_g, keys_h)
    assert np.array_equal(order_g, order_h)
    assert nodes_g.tobytes() == nodes_h.tobytes()
    assert orc.lbvh_depth(nodes_h) <= 64


def test_prim_bounds_contain_oracle_hits(pair):
    """Every oracle hit point lies inside the GPU-computed AABB of the primitive it hit."""
    name, scene, r, S, w, h = pair
    rays = raybatch.random_rays(raybatch.interest_bounds(r.flat), 20000, 11)
    o = S.trace_batch(rays.astype(np.float64))
    aabb = r.prim_bounds().astype(np.float64)
    hit = o["prim"] >= 0
    if name in ("bezier", "cornell_bezier", "patches", "teapot"):     # Q9: curve hit points are off the curve for |d| != 1
        hit &= r.flat.prims["type"][r.flat.first_of_logical[np.maximum(o["prim"], 0)]] != 5
    if name in ("patches", "teapot", "klein"):             # sub-patches: union of 16 leaf boxes; Klein: no box at all
        hit &= ~np.isin(r.flat.prims["type"][r.flat.first_of_logical[np.maximum(o["prim"], 0)]], (7, 8))
    b = aabb[r.flat.first_of_logical[o["prim"][hit]]]
    p = o["p"][hit]
    tol = 1e-4 * np.maximum(np.abs(p), 1.0)
    assert np.all(p >= b[:, :3] - tol) and np.all(p <= b[:, 3:] + tol)


@pytest.mark.parametrize("batch", ["camera", "random"])
def test_trace_batch_parity(pair, batch):
    name, scene, r, S, w, h = pair
    if batch == "camera":
        rays = raybatch.camera_grid(r, 64, 64)
    else:
        rays = raybatch.random_rays(raybatch.interest_bounds(r.flat), 100000 if name not in ("bezier", "patches", "teapot", "klein") else 30000, 5)
    rays64 = rays.astype(np.float64)
    gp = r.trace_batch(rays)
    o64 = S.trace_batch(rays64)
    o32 = S.trace_batch(rays64, precision=32)
    t2 = S.second_best_t(rays64, o64["prim"])
    c = raybatch.compare(gp, o64, o32, t2)
    print(f"\n[{name}/{batch}] n={c['n']} near_tie={c['filtered_near_tie']} unstable={c['filtered_unstable']} "
          f"id_mismatch={c['id_mismatch']} t_err_max={c['t_err_max']:.2e} t_bad={c['t_bad']} n_err_max={c['n_err_max']:.2e} p_err_max={c['p_err_max']:.2e}")
    assert c["filtered_near_tie"] + c["filtered_unstable"] <= 0.01 * c["n"]
    assert c["id_mismatch"] == 0, f"prim id mismatches at rays {c['id_mismatch_idx'][:10]}"
    assert c["t_bad"] == 0
    if name == "klein":
        # the fractal's central-difference normal (eps 0.01, geometry.scm:626-632) amplifies the fp32
        # rounding of t in the hit record (~1e-4 absolute at t ~ 800): stated tolerance 2e-2
        assert c["n_err_max"] <= 2e-2
    else:
        assert c["n_bad"] == 0
    # uv: rects everywhere; spheres only where |p.y| <= 1 (Q5: asin of the raw point)
    hm = c["hit_mask"]
    ptype = r.flat.prims["type"][r.flat.first_of_logical[np.maximum(o64["prim"], 0)]]
    uv_ok = hm & ((ptype >= 2) | (np.abs(o64["p"][:, 1]) <= 0.9))
    # bicubic patches (extension): (u, v) is the Newton root of a 2x2 system that is ill-conditioned towards
    # silhouettes and the degenerate pole rows of the teapot's lid / bottom: stated tolerance 1e-3 there
    for sel, tol in ((uv_ok & (ptype != 7), 2e-4), (uv_ok & (ptype == 7), 1e-3)):
        if sel.any():
            print(f"[{name}/{batch}] uv_err_max={max(np.max(np.abs(gp['u'][sel] - o64['uv'][sel, 0])), np.max(np.abs(gp['v'][sel] - o64['uv'][sel, 1]))):.2e} (tol {tol:g}, {int(sel.sum())} rays)")
            assert np.max(np.abs(gp["u"][sel] - o64["uv"][sel, 0])) <= tol
            assert np.max(np.abs(gp["v"][sel] - o64["uv"][sel, 1])) <= tol


def test_structured_ties_cornell(orc):
    """(iii) adversarial rays: box edges / corners, coplanar exact ties, parallel-to-plane rays,
    t-max just before / beyond a hit.  Coordinates are exactly representable, so the tie rule
     must give the oracle's primitive id exactly."""
    LAMB = m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5)))
    objs = [g.make_box((0, 0, 0), (1, 1, 1), LAMB), g.make_sphere((0, 0, -1), 0.5, LAMB),
            g.make_xy_rect(-1, 1, -1, 1, -0.5, LAMB), g.make_sphere((3, 0, 0), 1, LAMB), g.make_xy_rect(2, 4, -1, 1, 0, LAMB),
            g.make_box((1, 0, 0), (2, 1, 1), LAMB)]
    scene = g.make_scene(objs, scenes.default_camera(), scenes.sky_color)
    r = srt.Renderer(scene, device=0)
    S = orc.OracleScene(scene, flat=r.flat)
    rays = np.array([
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
        [3, 0, 0, 0, 1, 0, 0],            # from the sphere centre, in the rect plane (parallel, NaN t rejected)
    ], dtype=np.float32)
    for tmax in (999999999999.0, 4.0, 3.999, 4.001, 1.0):
        gp = r.trace_batch(rays, t_max=tmax)
        o = S.trace_batch(rays.astype(np.float64), t_max=tmax)
        assert np.array_equal(gp["prim"], o["prim"]), (tmax, gp["prim"], o["prim"])
        hit = o["prim"] >= 0
        assert np.allclose(gp["t"][hit], o["t"][hit], rtol=1e-6)
        assert np.allclose(gp["n"][hit], o["n"][hit], atol=1e-6)
    r.close()


def test_texture_parity(orc):
    scene = scenes.cfg3_next_week(32, 32)
    r = srt.Renderer(scene, device=0)
    S = orc.OracleScene(scene, flat=r.flat, perlin=r.perlin)
    rs = np.random.RandomState(4)
    uvp = np.concatenate([rs.random_sample((4000, 2)), rs.uniform(-6, 6, (4000, 3))], axis=1).astype(np.float32)
    kinds = r.flat.textures["kind"]
    for kind in (0, 1, 2, 3):
        ids = np.nonzero(kinds == kind)[0]
        assert len(ids), kind
        for quirks in (15, 0):
            a = r.eval_texture(int(ids[0]), uvp, quirks)
            b = S.tex_value(int(ids[0]), uvp.astype(np.float64), quirks)
            bad = np.abs(a - b).max(axis=1) > 2e-4
            # checker flips on sign(sin*sin*sin): fp32 may disagree only within rounding of a tile edge
            assert bad.mean() <= (0.002 if kind == 1 else 0.0), (kind, quirks, bad.sum(), np.abs(a - b).max())
    r.close()


def test_image_texture_parity(orc):
    """image-texture lookup (texture.scm:36-50): every texel, clamping beyond [0,1], NaN uv."""
    scene = scenes.image_scene(32, 32, nx=16, ny=8)
    r = srt.Renderer(scene, device=0)
    S = orc.OracleScene(scene, flat=r.flat, perlin=r.perlin)
    tid = int(np.nonzero(r.flat.textures["kind"] == 4)[0][0])
    rs = np.random.RandomState(5)
    # sample points away from texel edges (fp32 u*nx may round across an edge the f64 oracle does not)
    iu, iv = np.meshgrid(np.arange(16), np.arange(8), indexing="ij")
    fr = rs.uniform(0.05, 0.9, (2, iu.size))
    u = (iu.ravel() + fr[0]) / 16.0
    vv = 1.0 - (iv.ravel() + fr[1] + 0.001) / 8.0
    uvp = np.zeros((iu.size + 6, 5), dtype=np.float32)
    uvp[:iu.size, 0], uvp[:iu.size, 1] = u, vv
    uvp[iu.size:, :2] = [(-0.5, 0.5), (1.5, 0.5), (0.5, -0.5), (0.5, 1.5), (np.nan, 0.5), (0.5, np.nan)]
    a = r.eval_texture(tid, uvp, 15)
    b = S.tex_value(tid, uvp.astype(np.float64), 15)
    assert np.array_equal(a, b.astype(np.float32)), np.abs(a - b).max()
    img = r.flat.texture_objs[tid].image
    assert np.array_equal(a[:iu.size], (img[iv.ravel(), iu.ravel()].astype(np.float32) / np.float32(255)))
    r.close()


def test_raygen_parity(orc):
    scene = scenes.cfg2_random_spheres(120, 80)
    r = srt.Renderer(scene, device=0)
    S = orc.OracleScene(scene, flat=r.flat)
    p = r.params(120, 80, 0, 1, seed=9)
    pix = np.arange(0, 120 * 80, 7, dtype=np.int32)
    smp = (pix % 5).astype(np.int32)
    rays = r.eval_raygen(p, pix, smp)
    for k in range(0, len(pix), 97):
        px, s = int(pix[k]), int(smp[k])
        xi = orc.rng_block(9, px, s, 0, 0)
        x, y = px % 120, px // 120
        ref = S.get_ray((x + xi[0]) / 120, (y + xi[1]) / 80, xi[2], 9, px, s)
        assert np.allclose(rays[k], ref, rtol=2e-6, atol=2e-6)
    r.close()


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4", "bezier", "smoke", "patches", "klein", "teapot"])
def test_image_same_stream(name, orc):
    """Image parity under IDENTICAL Philox streams: GPU fp32 vs oracle f64 follow the same paths
    except where rounding flips a decision, so the per-pixel linear difference is tiny for almost
    all pixels; stated tolerance: median |diff| < 1e-4, >= 97 % of channel values within 1e-2,
    8-bit PSNR >= 35 dB."""
    fn, w, h = SMALL[name]
    spp = 8
    scene = fn(w, h)
    r = srt.Renderer(scene, device=0)
    S = orc.OracleScene(scene, flat=r.flat, perlin=r.perlin)
    img, st = r.render(w, h, spp, max_depth=50, seed=3)
    ref, nrays = S.render(w, h, spp, max_depth=50, seed=3)
    diff = np.abs(img.astype(np.float64) - ref) / spp
    a8 = srt.correct_gamma_quantise(img, spp).astype(np.float64)
    b8 = orc.resolve(ref, spp).astype(np.float64)
    mse = np.mean((a8 - b8) ** 2)
    psnr = 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)
    print(f"\n[{name}] rays gpu={st.rays} oracle={nrays} median={np.median(diff):.2e} within1e-2={np.mean(diff < 1e-2):.4f} psnr8={psnr:.1f} dB")
    assert np.all(np.isfinite(img))
    assert abs(st.rays - nrays) <= 0.02 * nrays
    # the Klein fractal scatters chaotically (see the normal tolerance above): looser gates
    assert np.median(diff) < 1e-4 and np.mean(diff < 1e-2) >= (0.85 if name == "klein" else 0.97)
    assert psnr >= (28.0 if name == "klein" else 35.0)
    r.close()


def test_image_converged_independent_seeds(orc):
    """Converged-image parity with INDEPENDENT seeds (cfg1): both estimators converge to the same
    image; tolerance: RMSE of the linear image <= 0.02, 8-bit PSNR >= 30 dB at 256 vs 256 spp."""
    w, h, spp = 100, 50, 256
    scene = scenes.cfg1_weekend(w, h)
    r = srt.Renderer(scene, device=0)
    S = orc.OracleScene(scene, flat=r.flat)
    img, _ = r.render(w, h, spp, max_depth=50, seed=101)
    ref, _ = S.render(w, h, spp, max_depth=50, seed=202)
    a, b = img.astype(np.float64) / spp, ref / spp
    rmse = np.sqrt(np.mean((np.minimum(a, 1) - np.minimum(b, 1)) ** 2))
    a8 = srt.correct_gamma_quantise(img, spp).astype(np.float64)
    b8 = orc.resolve(ref, spp).astype(np.float64)
    psnr = 10 * np.log10(255.0 ** 2 / np.mean((a8 - b8) ** 2))
    print(f"\n[converged cfg1] rmse={rmse:.4f} psnr8={psnr:.1f} dB")
    assert rmse <= 0.02 and psnr >= 30.0
    r.close()


def test_sample_range_additivity():
    """Sample-range sharding: rendering [0,8) equals [0,3) + [3,8) accumulated, and the
    result does not depend on the wave size (the Philox key is (pixel, sample, bounce))."""
    w, h = 96, 64
    r = srt.Renderer(scenes.cfg2_random_spheres(w, h), device=0)
    full, _ = r.render(w, h, 8, seed=5)
    part, _ = r.render(w, h, 3, seed=5, spp_begin=0)
    part, _ = r.render(w, h, 5, seed=5, spp_begin=3, rgb_sum=part)
    one, _ = r.render(w, h, 8, seed=5, wave_spp=1)
    assert np.allclose(full, part, rtol=1e-5, atol=1e-5)
    assert np.allclose(full, one, rtol=1e-5, atol=1e-5)
    r.close()


def test_full_size_cfg2_properties(orc):
    """BASELINE.json's headline configuration at its FULL size (1200x800 @ 500 spp, depth 50), checked
    through size-independent properties: (a) run-to-run determinism (integer accumulation), (b) the
    two-rank sample-range split sums to the single-GPU frame and its ray counts add exactly, (c) the
    frame box-filtered 10x10 equals the CPU oracle's render of the same camera at 120x80 (same
    image-plane footprint per pixel) within the oracle's Monte-Carlo noise."""
    cfg = scenes.CONFIGS["cfg2"]
    w, h, spp, depth = cfg["width"], cfg["height"], cfg["spp"], cfg["max_depth"]
    scene = cfg["scene"](w, h)
    r = srt.Renderer(scene, device=0)
    full, st = r.render(w, h, spp, max_depth=depth, seed=cfg["seed"])
    again, st2 = r.render(w, h, spp, max_depth=depth, seed=cfg["seed"])
    assert st.rays == st2.rays and np.array_equal(full, again)                                  # (a)
    lo, st_lo = r.render(w, h, spp // 2, max_depth=depth, seed=cfg["seed"], spp_begin=0)
    hi, st_hi = r.render(w, h, spp - spp // 2, max_depth=depth, seed=cfg["seed"], spp_begin=spp // 2)
    assert st_lo.rays + st_hi.rays == st.rays                                                   # (b)
    assert np.allclose(lo + hi, full, rtol=2e-6, atol=1e-4)
    assert np.all(np.isfinite(full)) and full.min() >= 0.0
    assert 3.5 < st.rays / (w * h * spp) < 5.0                  # mean path length of the Weekend scene
    small = cfg["scene"](w // 10, h // 10)
    S = orc.OracleScene(small)
    ospp = 96
    ref, _ = S.render(w // 10, h // 10, ospp, max_depth=depth, seed=77)
    ref /= ospp
    box = (full.astype(np.float64) / spp).reshape(h // 10, 10, w // 10, 10, 3).mean(axis=(1, 3))  # (c)
    rel = abs(box.mean() - ref.mean()) / ref.mean()
    rmse = np.sqrt(np.mean((box - ref) ** 2))
    print(f"\n[cfg2 full size] rays {st.rays} ({st.rays / (st.ms_total * 1e3):.0f} Mrays/s)  mean {box.mean():.4f} vs oracle {ref.mean():.4f} (rel {rel:.2e})  rmse {rmse:.4f}")
    assert rel < 0.01 and rmse < 0.06
    r.close()


def test_resolve_and_ppm(orc, tmp_path):
    rs = np.random.RandomState(0)
    rgb = (rs.random_sample((20, 30, 3)) * 40).astype(np.float32)
    a = srt.correct_gamma_quantise(rgb, 32)
    b = orc.resolve(rgb.astype(np.float64), 32)
    assert np.max(np.abs(a.astype(int) - b.astype(int))) <= 1 and np.mean(a != b) < 0.01
    srt.save_as_ppm(tmp_path / "a.ppm", a)
    orc.save_ppm(tmp_path / "b.ppm", a)
    assert (tmp_path / "a.ppm").read_text() == (tmp_path / "b.ppm").read_text()


def test_edge_cases():
    LAMB = m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5)))
    # empty scene: every ray misses, image = sky
    r = srt.Renderer(g.make_scene([], scenes.default_camera(16, 16), scenes.sky_color), device=0)
    assert np.all(r.trace_batch(np.array([[0, 0, 0, 0, 0, -1, 0]], np.float32))["prim"] == -1)
    img, st = r.render(16, 16, 2)
    assert st.rays == 16 * 16 * 2 and np.all(img > 0)
    r.close()
    # single primitive
    r = srt.Renderer(g.make_scene([g.make_sphere((0, 0, -1), 0.5, LAMB)], scenes.default_camera(16, 16), scenes.sky_color), device=0)
    assert r.trace_batch(np.array([[0, 0, 0, 0, 0, -1, 0]], np.float32))["prim"][0] == 0
    assert len(r.bvh_nodes()) == 1
    r.close()
    # zero-sample render is a no-op; bad parameters are reported, not crashed on
    r = srt.Renderer(scenes.cfg1_weekend(16, 8), device=0)
    img, st = r.render(16, 8, 0)
    assert np.all(img == 0)
    with pytest.raises(srt.host.ffi.SrtError):
        r.render(16, 8, 1, max_depth=-1)
    r.close()


@pytest.mark.parametrize("n", [5000, 70000])
def test_large_scene_global_memory_tree(orc, n):
    """Maximum sizes: trees that do not fit the shared-memory staging (5,000 spheres) and trees with
    >= 65536 nodes (70,000 spheres: no 16-bit far-child cache).  LBVH bit-exact, closest hits and a
    small image equal to the oracle's."""
    scene = scenes.sphere_cloud(n, 24, 24)
    r = srt.Renderer(scene, device=0)
    S = orc.OracleScene(scene, flat=r.flat, perlin=r.perlin)
    aabb = r.prim_bounds()
    items, glob = r.bvh_items()
    assert glob.tolist() == [0] and len(items) == n
    keys_h, order_h, nodes_h = orc.lbvh_build(aabb[items])
    for side in ("left", "right", "sibling"):
        leaf = nodes_h[side] < 0
        if side == "sibling":
            leaf[0] = False
        nodes_h[side][leaf] = ~items[~nodes_h[side][leaf]]
    assert r.bvh_nodes().tobytes() == nodes_h.tobytes()
    rays = np.concatenate([raybatch.camera_grid(r, 32, 32), raybatch.random_rays(([-20, 0, -20], [20, 41, 20]), 3000, 21)])
    gp = r.trace_batch(rays)
    rays64 = rays.astype(np.float64)
    o64 = S.trace_batch(rays64)
    c = raybatch.compare(gp, o64, S.trace_batch(rays64, precision=32), S.second_best_t(rays64, o64["prim"]))
    assert (o64["prim"] > 0).mean() > 0.15                     # the batch does exercise the cloud, not only the ground
    assert c["id_mismatch"] == 0 and c["t_bad"] == 0 and c["n_bad"] == 0
    assert c["filtered_near_tie"] + c["filtered_unstable"] <= 0.01 * c["n"]
    # short paths: every bounce off a small sphere amplifies fp32 rounding by ~distance/radius, so deep
    # paths through the cloud decorrelate from the f64 oracle's (same distribution, other samples)
    img, st = r.render(24, 24, 4, max_depth=3, seed=3)
    ref, nrays = S.render(24, 24, 4, max_depth=3, seed=3)
    diff = np.abs(img.astype(np.float64) - ref) / 4
    print(f"\n[cloud {n}] rays gpu={st.rays} oracle={nrays} median={np.median(diff):.2e} within1e-2={np.mean(diff < 1e-2):.4f}")
    assert abs(st.rays - nrays) <= 0.02 * nrays
    assert np.median(diff) < 1e-4 and np.mean(diff < 1e-2) >= 0.95
    r.close()


@pytest.mark.parametrize("quirks", [31, 15, 0])
def test_furnace_analytic_radiance_gpu(quirks):
    """The analytic furnace of tests/test_oracle_kat.py rendered by the CUDA path: every pixel on the
    sphere is weight_X * Le and every pixel off it is Le, to fp32 rounding."""
    from tests.test_oracle_kat import _furnace_cases, check_furnace
    for name, scene, expected in _furnace_cases():
        r = srt.Renderer(scene, device=0)
        img, st = r.render(32, 32, 4, max_depth=50, seed=9, quirks=quirks)
        check_furnace(img / 4, expected, 2e-6)
        r.close()


def test_constant_medium_free_flight_law_gpu(orc):
    """The analytic free-flight law of tests/test_oracle_kat.py on the CUDA path, and ray-by-ray
    agreement with the oracle (same Philox address: pixel = ray index)."""
    from tests.test_oracle_kat import _medium_scene
    R, rho, n = 2.0, 0.3, 200000
    scene = _medium_scene(R, rho)
    r = srt.Renderer(scene, device=0)
    rays = np.tile(np.array([0, 0, -10, 0, 0, 2.0, 0.0], dtype=np.float32), (n, 1))
    gp = r.trace_batch(rays)
    hit = gp["prim"] >= 0
    L, q = 2 * R, np.exp(-rho * 2 * R)
    assert abs(hit.mean() - (1 - q)) < 4 * np.sqrt(q * (1 - q) / n)
    dist = (gp["t"][hit] - 4.0) * 2.0
    assert abs(dist.mean() - (1 / rho - L * q / (1 - q))) < 0.01
    o = orc.OracleScene(scene, flat=r.flat).trace_batch(rays.astype(np.float64))
    same = (o["prim"] >= 0) == hit
    assert same.mean() > 0.9999                                     # xi within fp32 rounding of the exit decides a handful
    both = hit & (o["prim"] >= 0)
    assert np.max(np.abs(gp["t"][both] - o["t"][both]) / o["t"][both]) < 1e-4
    r.close()


def test_mixture_pdf_estimator(orc):
    """Rest-of-Life estimator mixture(hittable(light), cosine) (pdf.scm:18-41; the hittable part is
    absent upstream -> parity unpinned): GPU == oracle under identical streams, and with the
    book-correct cosine sampler (quirks = 0) it converges to the same image as the cosine-only
    estimator (both unbiased for the same integrand)."""
    from scheme_raytrace_b200.host import pdf
    w = h = 48
    scene = scenes.cfg4_cornell_box(w, h)
    flat = srt.flatten_scene(scene)
    light_obj = scene.obj_list[2]
    est, lights = pdf.estimator_of(pdf.make_mixture_pdf(pdf.make_hitable_pdf(light_obj), pdf.make_cosine_pdf()), flat)
    assert est == 1 and lights == [2]
    r = srt.Renderer(flat, device=0, lights=lights)
    S = orc.OracleScene(scene, flat=r.flat)
    S.set_lights(lights)
    img, st = r.render(w, h, 8, max_depth=50, seed=3, quirks=0, estimator=est)
    ref, nrays = S.render(w, h, 8, max_depth=50, seed=3, quirks=0, estimator=est)
    diff = np.abs(img.astype(np.float64) - ref) / 8
    print(f"\n[mixture] rays gpu={st.rays} oracle={nrays} median={np.median(diff):.2e} within1e-2={np.mean(diff < 1e-2):.4f}")
    assert np.all(np.isfinite(img)) and abs(st.rays - nrays) <= 0.02 * nrays
    assert np.median(diff) < 1e-4 and np.mean(diff < 1e-2) >= 0.97
    spp = 512
    a, _ = r.render(w, h, spp, max_depth=50, seed=11, quirks=0, estimator=1)
    b, _ = r.render(w, h, spp, max_depth=50, seed=12, quirks=0, estimator=0)
    ma, mb = a.mean(axis=(0, 1)) / spp, b.mean(axis=(0, 1)) / spp
    print(f"[mixture] image means mixture={ma} cosine-only={mb}")
    assert np.allclose(ma, mb, rtol=0.02)
    r.close()


def test_isotropic_material(orc):
    """isotropic (absent upstream, book semantics): GPU == oracle under identical streams."""
    iso = m.make_isotropic(t.constant_texture((0.6, 0.7, 0.8)))
    scene = g.make_scene([g.make_sphere((0, 0, -1), 0.5, iso), g.make_sphere((0, -100.5, -1), 100, m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5))))],
                         scenes.cfg1_weekend(64, 32).camera, scenes.sky_color)
    r = srt.Renderer(scene, device=0)
    S = orc.OracleScene(scene, flat=r.flat)
    img, st = r.render(64, 32, 8, seed=2)
    ref, nrays = S.render(64, 32, 8, seed=2)
    diff = np.abs(img.astype(np.float64) - ref) / 8
    assert abs(st.rays - nrays) <= 0.02 * nrays and np.median(diff) < 1e-4 and np.mean(diff < 1e-2) >= 0.97
    r.close()


def test_progressive_equals_batch():
    """Progressive passes (one sample per pass, main.scm:533-544) accumulate to exactly the batch
    render: the Philox key is (pixel, sample, bounce) and the accumulator is integer."""
    w, h = 64, 32
    scene = scenes.cfg1_weekend(w, h)
    pr = srt.ProgressiveRenderer(scene, w, h, max_depth=50, seed=4)
    for _ in range(5):
        img = pr.step()
    r = srt.Renderer(scene, device=0)
    full, _ = r.render(w, h, 5, max_depth=50, seed=4)
    assert pr.sample_count == 5 and np.allclose(pr.raw_data, full, rtol=1e-6, atol=1e-6)
    assert np.array_equal(img, srt.correct_gamma_quantise(pr.raw_data, 5))
    pr.close(); r.close()
