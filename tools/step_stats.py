"""Instrumented build (-DSRT_COUNT_STEPS): node steps and primitive tests per ray for the rays the
renderer actually traces (primary rays + oracle-free bounce rays reconstructed from the GPU hits).
Usage (GPU box): python tools/step_stats.py [cfg2]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "scheme_raytrace_b200", "csrc")
lib = os.path.join(CSRC, "libsrt_stats.so")
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                       "-DSRT_COUNT_STEPS", "-o", lib] + [os.path.join(CSRC, f) for f in ("srt_api.cu", "lbvh.cu", "wavefront.cu")])
from scheme_raytrace_b200.host import ffi
ffi.LIB_PATH = lib
import numpy as np
import scheme_raytrace_b200 as srt
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
cfg = srt.scenes.CONFIGS[name]
w, h = cfg["width"] // 4, cfg["height"] // 4
r = srt.Renderer(cfg["scene"](w, h))
p = r.params(w, h, 0, 1)
pix = np.arange(w * h, dtype=np.int32)
rays = r.eval_raygen(p, pix, np.zeros_like(pix))
rs = np.random.RandomState(0)
worst_sp = 0
for bounce in range(4):
    hit = r.trace_batch(rays)
    maxsp = (hit['v'] // 1000).astype(int); hit['v'] = hit['v'] % 1000     # instrumented build packs the deepest stack use
    worst_sp = max(worst_sp, int(maxsp.max()))
    print(f"{name} bounce {bounce}: rays {len(rays)}  hit {np.mean(hit['prim'] >= 0):.3f}  node steps/ray mean {hit['u'].mean():.1f} p50 {np.median(hit['u']):.0f} p95 {np.percentile(hit['u'], 95):.0f} max {hit['u'].max():.0f}"
          f"  prim tests/ray mean {hit['v'].mean():.2f} max {hit['v'].max():.0f}")
    m = hit["prim"] >= 0
    # diffuse bounce: cosine-ish direction around the normal (statistics only)
    n = hit["n"][m]; n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = n + rs.normal(size=n.shape) * 0.7
    rays = np.concatenate([hit["p"][m], d, np.zeros((m.sum(), 1))], axis=1).astype(np.float32)
depth = r.render(8, 8, 1)[1].bvh_depth
print("bvh depth", depth, "nodes", len(r.bvh_nodes()), "| deepest traversal-stack use", worst_sp, "of", depth + 1, "allocated entries")
assert worst_sp <= depth + 1
