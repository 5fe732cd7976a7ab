"""Is the traversal length of a bounce ray predictable from cheap ray features?  (instrumented build)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CSRC = os.path.join(ROOT, "scheme_raytrace_b200", "csrc")
lib = os.path.join(CSRC, "libsrt_stats.so")
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                       "-DSRT_COUNT_STEPS", "-o", lib] + [os.path.join(CSRC, f) for f in ("srt_api.cu", "lbvh.cu", "wavefront.cu")], stderr=subprocess.DEVNULL)
from scheme_raytrace_b200.host import ffi
ffi.LIB_PATH = lib
import numpy as np
import scheme_raytrace_b200 as srt
cfg = srt.scenes.CONFIGS["cfg2"]
w, h = 300, 200
r = srt.Renderer(cfg["scene"](w, h))
p = r.params(w, h, 0, 1)
pix = np.arange(w * h, dtype=np.int32)
rays = r.eval_raygen(p, pix, np.zeros_like(pix))
rs = np.random.RandomState(0)
allr, alls = [], []
for bounce in range(4):
    hit = r.trace_batch(rays)
    if bounce > 0:
        allr.append(rays.copy()); alls.append(hit["u"].copy())
    m = hit["prim"] >= 0
    n = hit["n"][m]; n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = n + rs.normal(size=n.shape) * 0.7
    rays = np.concatenate([hit["p"][m], d, np.zeros((m.sum(), 1))], axis=1).astype(np.float32)
R = np.concatenate(allr); S = np.concatenate(alls)
dy = R[:, 4] / np.linalg.norm(R[:, 3:6], axis=1)
print("bounce rays", len(S), "mean steps", S.mean(), "std", S.std())
for lo, hi in [(-1, -0.5), (-0.5, -0.2), (-0.2, 0), (0, 0.1), (0.1, 0.2), (0.2, 0.35), (0.35, 0.5), (0.5, 0.7), (0.7, 1.01)]:
    m = (dy >= lo) & (dy < hi)
    if m.any():
        print(f"dy in [{lo:5.2f},{hi:5.2f}): frac {m.mean():.3f} mean steps {S[m].mean():5.1f} p95 {np.percentile(S[m], 95):4.0f}")
def eff(order):
    s = S[order][: len(S) // 32 * 32].reshape(-1, 32)
    return s.sum() / (s.max(axis=1).sum())
print("lane efficiency: random order", eff(rs.permutation(len(S))), " sorted by dy", eff(np.argsort(dy)), " sorted by true steps", eff(np.argsort(S)))
