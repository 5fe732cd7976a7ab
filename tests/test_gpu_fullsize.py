"""GPU (-m gpu): BASELINE.json's configs 3, 4 and 5 at their FULL sizes, through size-independent properties (the
oracle cannot render 4e9 paths): modelled on test_full_size_cfg2_properties of tests/test_gpu_parity.py.

 (a) sample-range additivity: the frame rendered as two sample ranges sums to the frame rendered at once, and the ray
     counts add exactly (cfg4: 1024 x 1024 x 4096 = 2^32 paths, path ids beyond 32 bits; spp 4096 inside the 20-bit
     sample field of the packed path state);
 (b) run-to-run determinism (integer accumulation): a sample range rendered twice is bit-identical;
 (c) the frame box-filtered down to the oracle's resolution equals the CPU oracle's independent render of the same
     camera: mean radiance within 1 % (2 % for the small-emitter scene), RMSE within 1.5 x the oracle's own noise
     (estimated from two oracle renders with different seeds)."""
import numpy as np
import pytest
import scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import scenes

pytestmark = pytest.mark.gpu
# name -> (box-filter factor, oracle spp, mean tolerance, plausible mean path length)
CASES = {"cfg3": (10, 192, 0.03, (1.5, 6.0)), "cfg4": (16, 96, 0.01, (4.0, 12.0)), "cfg5": (40, 24, 0.01, (1.5, 6.0)), "cfg5_teapot": (40, 24, 0.01, (1.5, 6.0))}


@pytest.mark.parametrize("name", list(CASES))
def test_full_size_properties(orc, name):
    f, ospp, mean_tol, plen = CASES[name]
    cfg = scenes.CONFIGS[name]
    w, h, spp, depth, seed = cfg["width"], cfg["height"], cfg["spp"], cfg["max_depth"], cfg["seed"]
    assert (w * h * spp > 2 ** 32 - 1) == (name in ("cfg4", "cfg5", "cfg5_teapot"))
    r = srt.Renderer(cfg["scene"](w, h), device=0)
    full, st = r.render(w, h, spp, max_depth=depth, seed=seed)
    lo, st_lo = r.render(w, h, spp // 2, max_depth=depth, seed=seed, spp_begin=0)
    hi, st_hi = r.render(w, h, spp - spp // 2, max_depth=depth, seed=seed, spp_begin=spp // 2)
    hi2, st_hi2 = r.render(w, h, spp - spp // 2, max_depth=depth, seed=seed, spp_begin=spp // 2)
    assert st.paths == w * h * spp and st.nonfinite == 0
    assert st_lo.rays + st_hi.rays == st.rays                                                   # (a)
    assert np.allclose(lo + hi, full, rtol=4e-6, atol=1e-3)
    assert st_hi.rays == st_hi2.rays and np.array_equal(hi, hi2)                               # (b)
    assert np.all(np.isfinite(full)) and full.min() >= 0.0
    assert plen[0] < st.rays / (w * h * spp) < plen[1], st.rays / (w * h * spp)
    sw, sh = w // f, h // f                                                                     # (c)
    S = orc.OracleScene(cfg["scene"](sw, sh))
    a, _ = S.render(sw, sh, ospp, max_depth=depth, seed=77)
    b, _ = S.render(sw, sh, ospp, max_depth=depth, seed=78)
    ref = (a + b) / (2 * ospp)
    noise = np.sqrt(np.mean(((a - b) / ospp) ** 2)) / 2.0          # expected RMSE of (noise-free frame - ref)
    box = (full.astype(np.float64) / spp).reshape(sh, f, sw, f, 3).mean(axis=(1, 3))
    rel = abs(box.mean() - ref.mean()) / ref.mean()
    rmse = np.sqrt(np.mean((box - ref) ** 2))
    print(f"\n[{name} full size {w}x{h}@{spp}] rays {st.rays} ({st.rays / (st.ms_total * 1e3):.0f} Mrays/s, {st.ms_total:.0f} ms, {st.kernel_launches} launches, "
          f"tail {st.tail_runs})  mean {box.mean():.4f} vs oracle {ref.mean():.4f} (rel {rel:.2e})  rmse {rmse:.4f} (oracle noise {noise:.4f})")
    assert rel < mean_tol and rmse < 1.5 * noise + 1e-3
    r.close()
