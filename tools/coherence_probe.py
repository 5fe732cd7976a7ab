"""How much SIMT lane utilisation would ray re-ordering at compaction buy?  Instrumented build
(-DSRT_COUNT_STEPS, like tools/step_stats.py): node steps per ray for camera rays and synthetic
diffuse bounces in queue order; utilisation of a 32-ray warp = mean(steps) / max(steps).
Orderings: queue order | direction-octant buckets inside 256-ray tiles | global octant sort.
Usage (GPU box): python tools/coherence_probe.py [cfg2]"""
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
w, h = cfg["width"] // 2, cfg["height"] // 2
r = srt.Renderer(cfg["scene"](w, h))
p = r.params(w, h, 0, 1)
pix = np.arange(w * h, dtype=np.int32)
rays = r.eval_raygen(p, pix, np.zeros_like(pix))
rs = np.random.RandomState(0)


def util(steps):
    n = len(steps) // 32 * 32
    s = steps[:n].reshape(-1, 32)
    return s.sum() / (32.0 * s.max(axis=1).sum())


def octant(d):
    return (d[:, 0] > 0).astype(np.int64) | ((d[:, 1] > 0).astype(np.int64) << 1) | ((d[:, 2] > 0).astype(np.int64) << 2)


for bounce in range(5):
    hit = r.trace_batch(rays)
    steps = hit["u"].astype(np.float64)
    o8 = octant(rays[:, 3:6])
    tile = np.arange(len(rays)) // 256
    in_tile = np.lexsort((np.arange(len(rays)), o8, tile))          # stable: tile, then octant, then queue order
    glob = np.argsort(o8, kind="stable")
    print(f"{name} bounce {bounce}: {len(rays)} rays, mean steps {steps.mean():.1f}; warp utilisation: queue order {util(steps):.3f}"
          f" | octant buckets in 256-tiles {util(steps[in_tile]):.3f} | global octant sort {util(steps[glob]):.3f} | sorted by steps {util(np.sort(steps)):.3f}")
    m = hit["prim"] >= 0
    n = hit["n"][m]; n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = n + rs.normal(size=n.shape) * 0.7
    rays = np.concatenate([hit["p"][m], d, np.zeros((m.sum(), 1))], axis=1).astype(np.float32)
