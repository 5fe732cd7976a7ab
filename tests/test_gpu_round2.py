"""GPU (-m gpu): round-2 additions behind the C-ABI - the one-launch drain kernel, the cached iteration graph, the
device-resident progressive path against the reference's trace-line passes, multi-GPU from one process
(srt_render_multi), instanced spheres, argument validation, the non-finite guard and rays-per-bounce statistics."""
import ctypes as C
import json
import os
import numpy as np
import pytest
import scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import ffi, scenes, geometry as g, material as m, texture as t
from tests import raybatch
from tests.refspec import host_scene

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAMB = m.make_lambertian(t.constant_texture((0.5, 0.5, 0.5)))


def _render(r, w, h, spp, seed=3, depth=50, spp_begin=0, no_tail=False, no_graph=False, wave_spp=0, profile=False, two_pipes=False):
    p = r.params(w, h, spp_begin, spp_begin + spp, depth, seed, wave_spp=wave_spp)
    p.reserved[0], p.reserved[1], p.reserved[2], p.reserved[3], p.reserved[4] = int(profile), int(no_graph), 1, int(no_tail), 2 if two_pipes else 0
    out = np.zeros((h, w, 3), dtype=np.float32)
    st = ffi.Stats()
    ffi.check(r.lib.srt_render_host(r.h, C.byref(p), out.ctypes.data_as(C.c_void_p), C.byref(st)), "render_host")
    return out, st


@pytest.mark.parametrize("name", ["cfg1", "cfg3", "cfg4"])
def test_tail_kernel_is_bit_identical_to_the_wavefront_drain(name):
    """The drain kernel (every remaining path run to its end by one thread) must leave exactly the frame and the ray
    count of the launch-per-bounce drain it replaces: same Philox addresses, integer accumulation."""
    cfg = scenes.CONFIGS[name]
    w, h = (200, 100) if name == "cfg1" else (160, 160)
    r = srt.Renderer(cfg["scene"](w, h), device=0)
    for spp, wave in ((4, 0), (16, 3)):            # a frame that fits one wave / a queue refilled several times
        a, sa = _render(r, w, h, spp, wave_spp=wave)
        b, sb = _render(r, w, h, spp, wave_spp=wave, no_tail=True)
        c, sc = _render(r, w, h, spp, wave_spp=wave, no_graph=True)
        assert sa.tail_runs >= 1 and sb.tail_runs == 0
        assert sa.rays == sb.rays == sc.rays and np.array_equal(a, b) and np.array_equal(a, c)
        assert sa.kernel_launches < sb.kernel_launches
        print(f"\n[tail {name} spp={spp} wave={wave}] rays={sa.rays} launches {sa.kernel_launches} vs {sb.kernel_launches}  ms {sa.ms_total:.3f} vs {sb.ms_total:.3f}")
    r.close()


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg4"])
def test_two_concurrent_pipelines_equal_one(name):
    """Optional mode (params.reserved[4] == 2): the sample range split over two streaming pipelines that run concurrently
    on half-size grids and share the integer accumulator: frame and ray count must equal the single-pipeline render bit
    for bit (odd spp, queue refills)."""
    cfg = scenes.CONFIGS[name]
    w, h, spp = 640, 480, 57                     # 17.5 M paths: above the two-pipeline threshold
    r = srt.Renderer(cfg["scene"](w, h), device=0)
    a, sa = _render(r, w, h, spp, wave_spp=16, two_pipes=True)
    b, sb = _render(r, w, h, spp, wave_spp=16)
    c, sc = _render(r, w, h, spp, two_pipes=True)
    assert sa.pipes == 2 and sb.pipes == 1
    assert sa.rays == sb.rays == sc.rays and np.array_equal(a, b) and np.array_equal(a, c)
    print(f"\n[pipes {name}] rays={sa.rays}  two pipelines {sa.ms_total:.2f} ms  one {sb.ms_total:.2f} ms")
    r.close()


def test_cached_graph_is_reused_and_invalidated():
    """The executable iteration graph is cached on the scene: equal calls replay it, a change of seed / sample range /
    size / a re-commit must still give the right frame."""
    w, h = 96, 64
    r = srt.Renderer(scenes.cfg2_random_spheres(w, h), device=0)
    a, _ = _render(r, w, h, 8, seed=5, wave_spp=2)
    b, _ = _render(r, w, h, 8, seed=5, wave_spp=2)
    c, _ = _render(r, w, h, 8, seed=6, wave_spp=2)
    d, _ = _render(r, w, h, 8, seed=5, wave_spp=2, no_graph=True)
    assert np.array_equal(a, b) and np.array_equal(a, d) and not np.array_equal(a, c)
    lo, _ = _render(r, w, h, 4, seed=5, wave_spp=2)
    hi, _ = _render(r, w, h, 4, seed=5, wave_spp=2, spp_begin=4)        # same graph, other sample range
    assert np.allclose(lo + hi, a, rtol=2e-6, atol=1e-5)
    e, _ = _render(r, 48, 32, 8, seed=5, wave_spp=2)                      # other frame size
    r.commit()                                                           # re-commit: tables move, the key changes
    f, _ = _render(r, w, h, 8, seed=5, wave_spp=2)
    assert np.array_equal(a, f) and e.shape == (32, 48, 3) and np.all(np.isfinite(e))
    r.close()


@pytest.mark.parametrize("idx", [0, 1, 2, 3, 4])
def test_progressive_passes_against_the_references_trace_line(idx):
    """main.scm:452-469 trace-line, run row by row and pass by pass BY THE REFERENCE (tests/golden/ref_color.json,
    `trace_line_passes`: *raw-data* / *image* as the viewer shows them after every pass), against the device-resident
    progressive path: one srt_progressive_step per pass, only the 8-bit frame returned."""
    run = json.load(open(os.path.join(GOLD, "ref_color.json")))["runs"][idx]
    w, h = run["width"], run["height"]
    pr = srt.ProgressiveRenderer(host_scene(run["scene"], w, h), w, h, max_depth=run["max_depth"], seed=run["seed"])
    for k, ps in enumerate(run["trace_line_passes"]):
        img = pr.step(1).astype(np.int64)
        raw = np.asarray(ps["raw_data"], np.float64).reshape(h, w, 3)
        ref8 = np.asarray(ps["image"], np.int64).reshape(h, w, -1)[..., :3]
        mine = pr.raw_data.astype(np.float64)
        diff = np.abs(mine - raw) / (k + 1)
        lsb = np.abs(img - ref8)[ref8 >= 0]
        print(f"\n[trace-line {run['scene']} pass {k + 1}] median={np.median(diff):.2e} max={diff.max():.2e} 8-bit equal={np.mean(lsb == 0):.4f} within1={np.mean(lsb <= 1):.4f}")
        assert pr.sample_count == k + 1
        assert np.median(diff) < 1e-4 and np.mean(diff < 1e-2) >= 0.97 and np.mean(lsb <= 1) >= 0.97
    pr.close()


def test_progressive_step_rejects_a_gap():
    r = srt.Renderer(scenes.cfg1_weekend(32, 16), device=0)
    r.progressive_step(32, 16, 0, 2)
    with pytest.raises(ffi.SrtError):
        r.progressive_step(32, 16, 3, 4)            # the running sum holds 2 samples
    with pytest.raises(ffi.SrtError):
        r.progressive_step(16, 16, 2, 3)            # other size, not a fresh start
    r.close()


def test_instanced_spheres_against_the_oracle(orc):
    """translate / rotate-y above a sphere or moving-sphere leaf (legal upstream, geometry.scm:465-543): the ray goes to
    object space, p / normal come back; ids exact and t / p / normal within 1e-4 of the f64 oracle."""
    mov = g.make_moving_sphere((0.0, 0.5, 0.0), (0.0, 1.0, 0.0), 0.0, 1.0, 0.4, LAMB)
    objs = [g.make_sphere((0, -100.5, -1), 100, LAMB),
            g.translate(g.make_sphere((0, 0, 0), 0.5, LAMB), (1.0, 0.25, -2.0)),
            g.rotate_y(g.make_sphere((2.0, 0.0, 0.0), 0.5, LAMB), 40.0),
            g.translate(g.rotate_y(g.make_sphere((1.0, 0.5, 0.0), 0.3, LAMB), -65.0), (-1.0, 0.0, -1.5)),
            g.translate(g.rotate_y(mov, 30.0), (-2.0, 0.0, -2.0)),
            g.translate(g.rotate_y(g.make_box((0, 0, 0), (0.5, 0.8, 0.5), LAMB), 20.0), (0.5, -0.5, -3.0))]
    scene = g.make_scene(objs, scenes.default_camera(), scenes.sky_color)
    r = srt.Renderer(scene, device=0)
    S = orc.OracleScene(scene, flat=r.flat)
    # rays aimed at the instanced spheres' WORLD positions (rotate-y(40) of (2,0,0) etc.) from all around, plus a random batch
    rs = np.random.RandomState(12)
    centres = np.array([(1.0, 0.25, -2.0), (2 * np.cos(np.radians(40)), 0.0, -2 * np.sin(np.radians(40))),
                        (np.cos(np.radians(65)) - 1.0, 0.5, np.sin(np.radians(65)) - 1.5), (-2.0, 0.75, -2.0)])
    c = centres[rs.randint(0, 4, 20000)]
    o = c + rs.normal(size=(20000, 3)) * 2.0
    d = (c + rs.normal(size=(20000, 3)) * 0.35 - o) * rs.uniform(0.5, 2.0, (20000, 1))
    aimed = np.concatenate([o, d, rs.random_sample((20000, 1))], axis=1).astype(np.float32)
    rays = np.concatenate([raybatch.camera_grid(r, 48, 48), aimed, raybatch.random_rays(raybatch.interest_bounds(r.flat), 20000, 11)])
    rays64 = rays.astype(np.float64)
    got = r.trace_batch(rays)
    o64 = S.trace_batch(rays64)
    c = raybatch.compare(got, o64, S.trace_batch(rays64, precision=32), S.second_best_t(rays64, o64["prim"]))
    inst = np.isin(o64["prim"], [1, 2, 3, 4])
    print(f"\n[instanced spheres] rays={c['n']} hits on instanced spheres={int(inst.sum())} near_tie={c['filtered_near_tie']} unstable={c['filtered_unstable']} "
          f"id_mismatch={c['id_mismatch']} t={c['t_err_max']:.2e} n={c['n_err_max']:.2e} p={c['p_err_max']:.2e}")
    assert inst.sum() > 2000
    assert c["filtered_near_tie"] + c["filtered_unstable"] <= 0.01 * c["n"]
    assert c["id_mismatch"] == 0 and c["t_bad"] == 0 and c["n_bad"] == 0 and c["p_err_max"] <= 1e-4
    # and the picture agrees under shared random streams
    img, st = r.render(64, 48, 8, seed=9)
    oimg, nr = S.render(64, 48, 8, max_depth=50, seed=9)
    err = np.abs(img.astype(np.float64) - oimg) / 8
    assert np.median(err) < 1e-4 and np.mean(err < 1e-2) > 0.97
    r.close()


def test_arguments_are_validated():
    scene = scenes.cfg4_cornell_box(16, 16)
    r = srt.Renderer(scene, device=0)
    out = np.zeros((16, 16, 3), np.float32)
    for field, val in (("estimator", 7), ("sky", 9), ("spp_begin", -1), ("max_depth", 5000), ("width", 0)):
        p = r.params(16, 16, 0, 1)
        setattr(p, field, val)
        rc = r.lib.srt_render_host(r.h, C.byref(p), out.ctypes.data_as(C.c_void_p), None)
        assert rc == -3, (field, rc)
    r.close()
    # lights: only un-instanced spheres / rects (a box face under rotate-y is instanced)
    with pytest.raises(ffi.SrtError):
        srt.Renderer(scene, device=0, lights=[8])
    # an instanced curve would be intersected untransformed: refused at commit
    from scheme_raytrace_b200.host import bezier as b
    curve = g.translate(b.make_bezier((-1, 0, -1), (-0.8, 1, 1), (0.8, -1, 1), (1, 0, -1), 0.1, LAMB), (1, 0, 0))
    with pytest.raises(ffi.SrtError):
        srt.Renderer(g.make_scene([curve], scenes.default_camera(), scenes.sky_color), device=0)


def test_nonfinite_contributions_are_dropped_and_counted():
    inf_light = m.make_diffuse_light(t.constant_texture((float("inf"), 1.0, 1.0)))
    objs = [g.make_sphere((0, 0, -1), 0.5, inf_light), g.make_sphere((0, -100.5, -1), 100, LAMB)]
    r = srt.Renderer(g.make_scene(objs, scenes.default_camera(), scenes.sky_color), device=0)
    img, st = _render(r, 32, 16, 4)
    assert st.nonfinite > 0 and np.all(np.isfinite(img))
    r.close()
    r = srt.Renderer(scenes.cfg1_weekend(32, 16), device=0)
    img, st = _render(r, 32, 16, 4)
    assert st.nonfinite == 0
    r.close()


def test_rays_per_bounce_in_profile_mode():
    r = srt.Renderer(scenes.cfg1_weekend(64, 32), device=0)
    img, st = _render(r, 64, 32, 8, profile=True)
    per = list(st.rays_per_bounce)
    assert sum(per) == st.rays and per[0] == 64 * 32 * 8 and all(per[k] >= per[k + 1] for k in range(6))
    img2, st2 = _render(r, 64, 32, 8)
    assert np.array_equal(img, img2) and sum(st2.rays_per_bounce) == 0
    r.close()


def test_two_scenes_on_one_device_and_explicit_device_binding():
    """Device / kernel-variant state is per scene: two live handles, interleaved calls."""
    a = srt.Renderer(scenes.cfg1_weekend(32, 16), device=0)
    b = srt.Renderer(scenes.cfg4_cornell_box(32, 32), device=0)
    ia, _ = _render(a, 32, 16, 2)
    ib, _ = _render(b, 32, 32, 2)
    ia2, _ = _render(a, 32, 16, 2)
    ib2, _ = _render(b, 32, 32, 2)
    assert np.array_equal(ia, ia2) and np.array_equal(ib, ib2)
    a.close(); b.close()


needs2 = pytest.mark.skipif(ffi.load().srt_device_count() < 2, reason="needs >= 2 GPUs")


@needs2
@pytest.mark.parametrize("mode", ["nccl", "p2p"])
def test_multi_gpu_frame_is_bit_identical(mode, tmp_path):
    """srt_render_multi (one process, every GPU, sample-range sharding, one reduce of the integer accumulators) must
    return exactly the single-GPU frame.  Run in a child process so that SRT_MULTI_REDUCE selects the reduce."""
    import subprocess
    import sys
    code = f'''
import numpy as np, scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import scenes, ffi
import ctypes as C
w, h, spp = 160, 96, 13
scene = scenes.cfg3_next_week(w, h)
r1 = srt.Renderer(scene, device=0)
one, st1 = r1.render(w, h, spp, seed=4)
img1 = srt.correct_gamma_quantise(one, spp)
r1.close()
rm = srt.Renderer(scene, gpus=0)
n = rm.gpus
ver = C.c_int32(0)
mode = rm.lib.srt_multi_reduce_mode(C.byref(ver))
many, img, stm = rm.render_multi(w, h, spp, seed=4)
assert n >= 2 and stm.rays == st1.rays and stm.paths == st1.paths, (n, stm.rays, st1.rays)
assert np.array_equal(one, many) and np.array_equal(img, img1)
# running sum semantics: a second pass adds to the first
more, img2, _ = rm.render_multi(w, h, 3, seed=4, spp_begin=spp, rgb_sum=many.copy())
r1 = srt.Renderer(scene, device=0)
ref, _ = r1.render(w, h, 3, seed=4, spp_begin=spp, rgb_sum=one.copy())
assert np.array_equal(more, ref)
# fewer samples than GPUs: empty ranges contribute zero
tiny, _, stt = rm.render_multi(w, h, 1, seed=4)
t1, _ = r1.render(w, h, 1, seed=4)
assert np.array_equal(tiny, t1)
print("MULTI_OK", n, mode, ver.value, stm.rays)
'''
    env = dict(os.environ, SRT_MULTI_REDUCE=mode, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    print(out.stdout[-400:], out.stderr[-2000:])
    assert out.returncode == 0 and "MULTI_OK" in out.stdout
    got_mode = int(out.stdout.split("MULTI_OK")[1].split()[1])
    assert got_mode == (0 if mode == "nccl" else 1)


@needs2
def test_cli_on_all_gpus_matches_one_gpu(tmp_path):
    import subprocess
    from scheme_raytrace_b200.host import scenefile
    from scheme_raytrace_b200.host.perlin import perlin_generate
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scheme_raytrace_b200", "cli", "srt_render")
    path = str(tmp_path / "scene.srt")
    scenefile.write_scene_file(path, srt.flatten_scene(scenes.cfg4_cornell_box(48, 48)), perlin_generate(3))
    outs = []
    for gpus in ("1", "0"):
        out = str(tmp_path / f"g{gpus}.ppm")
        rr = subprocess.run([cli, path, "--width", "48", "--height", "48", "--spp", "9", "--depth", "50", "--seed", "3", "--gpus", gpus, "--out", out], capture_output=True, text=True)
        assert rr.returncode == 0, rr.stderr
        outs.append(open(out).read())
    assert outs[0] == outs[1]
