"""CPU-side checks (-m "not gpu"): host logic, LBVH host reference, C-ABI library loads and
exports every symbol of include/srt.h (no compute without a GPU)."""
import re
import os
import numpy as np
import pytest
import scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import ffi, flatten, scenes, geometry as g, material as m, texture as t

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    hdr = open(os.path.join(ROOT, "include", "srt.h")).read()
    declared = set(re.findall(r"\b(srt_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(ffi.SYMBOLS), declared ^ set(ffi.SYMBOLS)
    lib = ffi.load()
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_no_cpu_fallback_without_gpu():
    lib = ffi.load()
    if lib.srt_device_count() > 0:
        pytest.skip("GPU present")
    assert lib.srt_init(0) == -1 and b"no CPU fallback" in lib.srt_last_error()
    with pytest.raises(ffi.SrtError):
        srt.Renderer(scenes.cfg1_weekend(8, 8))


def test_struct_layouts_match_header(tmp_path):
    """The ctypes / numpy mirrors against include/srt.h as a C compiler lays it out (sizes and the
    offsets of the fields the host writes or reads)."""
    import ctypes as C
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "srt.h"\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(SrtRenderParams), sizeof(SrtStats), sizeof(SrtPrim),'
        ' sizeof(SrtCamera), sizeof(SrtXform), sizeof(SrtMaterial), sizeof(SrtTexture), sizeof(SrtBvhNode), sizeof(SrtRay), sizeof(SrtHit),'
        ' offsetof(SrtStats, nonfinite), offsetof(SrtStats, rays_per_bounce)); return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    from scheme_raytrace_b200.host import render
    want = [C.sizeof(ffi.RenderParams), C.sizeof(ffi.Stats), flatten.PRIM_DTYPE.itemsize, flatten.CAMERA_DTYPE.itemsize, flatten.XFORM_DTYPE.itemsize,
            flatten.MATERIAL_DTYPE.itemsize, flatten.TEXTURE_DTYPE.itemsize, render.BVH_NODE_DTYPE.itemsize, render.RAY_DTYPE.itemsize, render.HIT_DTYPE.itemsize,
            ffi.Stats.nonfinite.offset, ffi.Stats.rays_per_bounce.offset]
    assert got == want, (got, want)
    assert C.sizeof(ffi.RenderParams) == 64 and C.sizeof(ffi.Stats) == 136


def test_flatten_order_and_instances():
    f = srt.flatten_scene(scenes.cfg4_cornell_box(32, 32))
    assert len(f.prims) == 18 and len(f.xforms) == 2
    assert list(f.prims["type"][:6]) == [4, 4, 3, 3, 3, 2] and list(f.prims["flags"][:6]) == [1, 0, 1, 1, 0, 1]
    # make-box face order geometry.scm:446-457: xy@z1, flip xy@z0, xz@y1, flip xz@y0, yz@x1, flip yz@x0
    assert list(f.prims["type"][6:12]) == [2, 2, 3, 3, 4, 4] and list(f.prims["flags"][6:12]) == [0, 1, 0, 1, 0, 1]
    assert np.allclose(f.xforms[0]["off"], (130, 0, 65)) and np.isclose(f.xforms[0]["sin_t"], np.sin(np.radians(-18)))
    # nested chain composes: translate(rotate-y(translate(x)))
    LAMB = m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5)))
    o = g.translate(g.rotate_y(g.translate(g.make_xy_rect(0, 1, 0, 1, 0, LAMB), (1, 0, 0)), 90), (0, 0, 5))
    f2 = srt.flatten_scene(g.make_scene([o], scenes.default_camera(), scenes.black))
    # rotate-y(90): (x,y,z) -> (z, y, -x); inner offset (1,0,0) -> (0,0,-1); plus (0,0,5)
    assert np.allclose(f2.xforms[0]["off"], (0, 0, 4), atol=1e-6) and np.isclose(f2.xforms[0]["sin_t"], 1.0)
    assert f2.sky == flatten.SKY_BLACK


def test_random_scene_is_deterministic_and_reference_shaped():
    a = srt.flatten_scene(scenes.cfg2_random_spheres())
    b = srt.flatten_scene(scenes.cfg2_random_spheres())
    assert a.prims.tobytes() == b.prims.tobytes() and 450 <= len(a.prims) <= 500
    # push! prepends: the three big spheres come first, the ground last (main.scm:36-88)
    assert a.prims[0]["p"][3] == 1.0 and a.prims[-1]["p"][3] == 1000.0
    c3 = srt.flatten_scene(scenes.cfg3_next_week())
    assert set(c3.prims["type"]) == {0, 1, 2} and {0, 1, 2, 3} <= set(c3.textures["kind"])


def test_lbvh_host_reference_properties(orc):
    rs = np.random.RandomState(1)
    for n in (1, 2, 3, 17, 257, 1000):
        c = rs.uniform(-10, 10, (n, 3)).astype(np.float32)
        if n == 257:
            c[100:160] = c[100]                      # duplicate centroids -> equal Morton keys
        r = rs.uniform(0.01, 1, (n, 1)).astype(np.float32)
        aabb = np.concatenate([c - r, c + r], axis=1)
        keys, order, nodes = orc.lbvh_build(aabb)
        assert np.all(keys[:-1] <= keys[1:]) and sorted(order.tolist()) == list(range(n))
        eq = keys[:-1] == keys[1:]
        assert np.all(order[:-1][eq] < order[1:][eq])          # stable
        if n == 1:
            assert len(nodes) == 1 and nodes[0]["left"] == ~0
            continue
        assert len(nodes) == n - 1
        leaves = [~x for x in list(nodes["left"]) + list(nodes["right"]) if x < 0]
        assert sorted(leaves) == list(range(n))                # every primitive exactly once
        # each stored child box (centre / padded half extent) contains every primitive AABB below it
        import functools

        @functools.lru_cache(maxsize=None)
        def box_of(ch):
            if ch < 0:
                return tuple(aabb[~ch, :3].astype(np.float64)), tuple(aabb[~ch, 3:].astype(np.float64))
            nd = nodes[ch]
            (llo, lhi), (rlo, rhi) = box_of(int(nd["left"])), box_of(int(nd["right"]))
            return tuple(np.minimum(llo, rlo)), tuple(np.maximum(lhi, rhi))
        for i, nd in enumerate(nodes):
            for side, ch in (("l", int(nd["left"])), ("r", int(nd["right"]))):
                lo, hi = box_of(ch)
                c, e = nd[side + "c"].astype(np.float64), nd[side + "e"].astype(np.float64)
                assert np.all(c - e <= lo) and np.all(c + e >= hi)
                if ch >= 0:
                    assert nodes[ch]["parent"] == i and nodes[ch]["sibling"] == (nd["right"] if side == "l" else nd["left"])
        assert nodes[0]["parent"] == -1 and orc.lbvh_depth(nodes) <= 64


def test_morton_expand_bits(orc):
    # key of a 2-primitive set: the centroid at cmax gets all-ones on that axis
    aabb = np.array([[0, 0, 0, 0, 0, 0], [1, 2, 4, 1, 2, 4]], np.float32)
    keys, order, _ = orc.lbvh_build(aabb)
    assert keys[0] == 0 and keys[1] == (1 << 63) - 1 and list(order) == [0, 1]


def test_flatten_constant_medium():
    f = srt.flatten_scene(scenes.cornell_smoke(32, 32))
    assert len(f.prims) == 20 and list(f.prims["type"][6:8]) == [6, 6]
    assert np.all((f.prims["flags"][:8] & 2) == 0) and np.all((f.prims["flags"][8:] & 2) == 2)     # boundaries form a suffix
    assert tuple(f.prims[6]["p"][:3]) == (np.float32(0.01), 8.0, 6.0) and tuple(f.prims[7]["p"][1:3]) == (14.0, 6.0)
    assert f.materials[f.prims[6]["material"]]["kind"] == 0        # phase function = lambertian (geometry.scm:546)


def test_points_csv_to_bezier_chain(tmp_path, orc):
    """points.scm:10-50: CSV -> points (x magnitude) -> tightness-0.5 control points -> curve objects.
    Expected control points worked out by hand: [p1, p1 + (p2 - p0)/6, p2 - (p3 - p1)/6, p2]."""
    from scheme_raytrace_b200.host import points as pts, bezier as bz
    f = tmp_path / "pts.csv"
    f.write_text("0,0,0\n1,0,0\n2,1,0\n3,1,0\n4,0,0\n")
    P = pts.load_points(str(f), 2)
    assert [tuple(p) for p in P] == [(0, 0, 0), (2, 0, 0), (4, 2, 0), (6, 2, 0), (8, 0, 0)]
    segs = pts.points_to_bezier(P)                           # last = n - 2 = 3: i = 1, 2
    assert len(segs) == 2
    exp = [[(2, 0, 0), (2 + 4 / 6, 2 / 6, 0), (4 - 4 / 6, 2 - 2 / 6, 0), (4, 2, 0)],
           [(4, 2, 0), (4 + 4 / 6, 2 + 2 / 6, 0), (6 - 4 / 6, 2 + 2 / 6, 0), (6, 2, 0)]]
    assert np.allclose(np.asarray(segs, float), np.asarray(exp, float), atol=1e-15)
    assert np.allclose(np.subtract(segs[0][3], segs[0][2]), np.subtract(segs[1][1], segs[1][0]))   # C1 at the joint
    mat = m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5)))
    objs = pts.bezier_to_objs(segs, 0.2, mat)
    assert len(objs) == 2 and all(o.kind == g.BEZIER for o in objs)
    flat = srt.flatten_scene(g.make_scene(objs, scenes.default_camera(), scenes.sky_color))
    assert list(flat.prims["type"]) == [5, 5] and np.allclose(flat.prims["p"][1][:3], (4, 2, 0))
    # the oracle sees the chain: a ray aimed at the joint hits one of the two segments
    S = orc.OracleScene(g.make_scene(objs, scenes.default_camera(), scenes.sky_color), quantise=False)
    r = S.trace_batch([[4.03, 2.0, 5, 0, 0, -1, 0], [3.0, 5.0, 5, 0, 0, -1, 0]])
    assert r["prim"][0] in (0, 1) and r["prim"][1] == -1


def test_quirk_bits_match_header():
    """host/ffi.py's quirk bits are the #defines of include/srt.h; the reference mask carries Q15 (found by executing the reference)."""
    import re
    from scheme_raytrace_b200.host import ffi
    text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "srt.h")).read()
    hdr = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+SRT_(Q\w+)\s+(\d+)", text)}
    mine = {k: v for k, v in vars(ffi).items() if re.fullmatch(r"Q\d+_\w+|QUIRKS_REFERENCE", k)}
    assert hdr == mine and mine["QUIRKS_REFERENCE"] == sum(v for k, v in mine.items() if k != "QUIRKS_REFERENCE") == 31


def path_order(path_id, width, height, n_samples, tlw=4, tlh=4):
    """The path order k_regen uses (csrc/wavefront.cu, "Path order"): path id -> (pixel, sample).  Frames that
    divide into 2^tlw x 2^tlh tiles are walked tile-major (all samples of a tile, then the next tile; inside a sample the tile's
    pixels as 8 x 4 blocks, row-major); other frames sample-major in scanlines.  Restated here as the specification the
    kernel's index arithmetic has to meet; the frame itself cannot depend on it (tests/test_gpu_fullsize.py, additivity and
    determinism at full size; checksums of both orders in profiles/r2_sweep17_path_order.txt)."""
    npix = width * height
    tiled = (width % (1 << tlw) == 0) and (height % (1 << tlh) == 0)
    if not tiled:
        sample, pos = divmod(path_id, npix)
        return pos, sample
    span = n_samples << (tlw + tlh)
    tile, inner = divmod(path_id, span)
    sample, r = divmod(inner, 1 << (tlw + tlh))
    by, bx = divmod(tile, width >> tlw)
    blk, lane = divmod(r, 32)
    byi, bxi = divmod(blk, 1 << (tlw - 3))
    x = (bx << tlw) + bxi * 8 + (lane & 7)
    y = (by << tlh) + byi * 4 + (lane >> 3)
    return y * width + x, sample


@pytest.mark.parametrize("w,h,s", [(48, 32, 3), (64, 16, 1), (50, 30, 2), (16, 16, 5)])
def test_path_order_is_a_bijection_and_keeps_tiles_together(w, h, s):
    ids = np.arange(w * h * s)
    pix, smp = np.array([path_order(int(i), w, h, s) for i in ids]).T
    assert sorted(zip(pix.tolist(), smp.tolist())) == [(p, k) for p in range(w * h) for k in range(s)]      # every (pixel, sample) once
    if w % 16 == 0 and h % 16 == 0:
        x, y = pix % w, pix // w
        for k in range(0, len(ids), 32):                                   # a warp of fresh paths: one 8 x 4 block of one sample
            assert len(set(smp[k:k + 32])) == 1 and np.ptp(x[k:k + 32]) == 7 and np.ptp(y[k:k + 32]) == 3
        for k in range(0, len(ids), 256 * s):                              # all samples of one 16 x 16 tile are consecutive
            assert len(set(zip((x[k:k + 256 * s] // 16).tolist(), (y[k:k + 256 * s] // 16).tolist()))) == 1
    else:
        assert np.array_equal(pix, np.tile(np.arange(w * h), s))              # scanlines, sample-major
