"""Converged-image parity (SURVEY §8c "Image parity"): GPU render vs the f64 oracle's CPU render
with INDEPENDENT seeds at the config's own spp (resolution reduced so the oracle finishes in
seconds-minutes).  Reports RMSE / PSNR on the linear image (clamped to [0,1]) and on the 8-bit
image, and the Monte-Carlo noise floor (oracle vs oracle with another seed at a subset).
Usage (GPU box): python tools/converged_parity.py > gpurun_out/converged_parity.txt"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import scheme_raytrace_b200 as srt
from oracle import oracle as O

CASES = [("cfg1", 200, 100, 16), ("cfg1", 200, 100, 1024), ("cfg2", 240, 160, 500), ("cfg3", 160, 160, 1000), ("cfg4", 128, 128, 4096), ("cfg5", 192, 108, 1024)]


def psnr8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


for name, w, h, spp in CASES:
    cfg = srt.scenes.CONFIGS[name]
    scene = cfg["scene"](w, h)
    r = srt.Renderer(scene)
    S = O.OracleScene(scene, flat=r.flat, perlin=r.perlin)
    t0 = time.perf_counter(); g, st = r.render(w, h, spp, max_depth=cfg["max_depth"], seed=1001); tg = time.perf_counter() - t0
    t0 = time.perf_counter(); o, nr = S.render(w, h, spp, max_depth=cfg["max_depth"], seed=2002); to = time.perf_counter() - t0
    o2, _ = S.render(w, h, spp, max_depth=cfg["max_depth"], seed=3003) if spp * w * h <= 4e7 else (None, 0)
    a, b = np.minimum(g.astype(np.float64) / spp, 1), np.minimum(o / spp, 1)
    rmse = np.sqrt(np.mean((a - b) ** 2))
    g8, o8 = srt.correct_gamma_quantise(g, spp), O.resolve(o, spp)
    line = f"{name} {w}x{h} @ {spp} spp: linear RMSE {rmse:.4f}  8-bit PSNR {psnr8(g8, o8):.1f} dB  mean gpu {a.mean():.4f} oracle {b.mean():.4f}  rays gpu {st.rays} oracle {nr}  (gpu {tg:.2f} s, oracle {to:.1f} s on {os.cpu_count()} threads)"
    if o2 is not None:
        line += f"  | noise floor oracle-vs-oracle: RMSE {np.sqrt(np.mean((np.minimum(o2 / spp, 1) - b) ** 2)):.4f} PSNR {psnr8(O.resolve(o2, spp), o8):.1f} dB"
    print(line, flush=True)
    r.close()
