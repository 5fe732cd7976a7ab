"""cli/srt_render (native host over the C-ABI) + the flat scene file: the route Gauche scene
scripts take (scheme/srt-scene.scm writes the same format)."""
import os
import subprocess
import numpy as np
import pytest
import scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import scenefile, scenes, ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "scheme_raytrace_b200", "cli", "srt_render")


def _write(tmp_path, scene, lights=()):
    from scheme_raytrace_b200.host.perlin import perlin_generate
    flat = srt.flatten_scene(scene)
    path = str(tmp_path / "scene.srt")
    scenefile.write_scene_file(path, flat, perlin_generate(3), lights)
    return path, flat


def test_scene_file_roundtrip_fields(tmp_path):
    path, flat = _write(tmp_path, scenes.cornell_smoke(16, 16))
    toks = open(path).read().split()
    assert toks[:2] == ["srt-scene", "1"]
    i = toks.index("prims")
    n = int(toks[i + 1])
    assert n == len(flat.prims)
    first = toks[i + 2:i + 2 + 20]
    assert [int(first[0]), int(first[1]), int(first[2]), int(first[3])] == [int(flat.prims[0][k]) for k in ("type", "flags", "material", "xform")]
    assert np.allclose([float(x) for x in first[4:]], flat.prims[0]["p"])
    assert toks[toks.index("textures") + 1] == str(len(flat.textures)) and toks[toks.index("lights") + 1] == "0"


def test_cli_fails_loudly_without_gpu(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    if ffi.load().srt_device_count() > 0:
        pytest.skip("GPU present")
    path, _ = _write(tmp_path, scenes.cfg1_weekend(16, 8))
    r = subprocess.run([CLI, path, "--width", "16", "--height", "8", "--spp", "1", "--out", str(tmp_path / "o.ppm")], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and not (tmp_path / "o.ppm").exists()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1", "smoke", "patches", "image"])
def test_cli_matches_python_api(tmp_path, name):
    fn = {"cfg1": scenes.cfg1_weekend, "smoke": scenes.cornell_smoke, "patches": scenes.cfg5_patches, "image": scenes.image_scene}[name]
    w, h, spp = 64, 32, 4
    scene = fn(w, h)
    path, flat = _write(tmp_path, scene)
    out = str(tmp_path / "cli.ppm")
    r = subprocess.run([CLI, path, "--width", str(w), "--height", str(h), "--spp", str(spp), "--depth", "50", "--seed", "7", "--out", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ren = srt.Renderer(scene, device=0)
    rgb, _ = ren.render(w, h, spp, max_depth=50, seed=7)
    ref = str(tmp_path / "api.ppm")
    srt.save_as_ppm(ref, srt.correct_gamma_quantise(rgb, spp))
    assert open(out).read() == open(ref).read()
    ren.close()
