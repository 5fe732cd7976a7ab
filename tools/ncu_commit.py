#!/usr/bin/env python
"""Commit (GPU LBVH build) of a 70,000-sphere cloud + a small render through the global-memory tree: the workload the ncu
captures of the LBVH kernels and of the large-scene extend variant are taken on."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scheme_raytrace_b200 as srt                     # noqa: E402
from scheme_raytrace_b200.host import scenes           # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 70000
w, h, spp = 960, 540, int(sys.argv[2]) if len(sys.argv) > 2 else 16
r = srt.Renderer(scenes.sphere_cloud(n, w, h), device=0)
best = 1e9
for _ in range(3):
    img, st = r.render(w, h, spp, max_depth=8, seed=1)
    best = min(best, st.ms_total)
print(f"cloud {n}: commit {st.ms_commit:.3f} ms, {st.bvh_nodes} nodes depth {st.bvh_depth}, {st.rays} rays in {best:.2f} ms = {st.rays / best / 1e3:.0f} Mrays/s")
