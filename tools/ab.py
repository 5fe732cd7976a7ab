#!/usr/bin/env python
"""A/B timing of one workload through the C-ABI (device-resident accumulation, CUDA events inside the library).

    [SRT_LIB=exp/libsrt_X.so] [SRT_TAIL_MAX=..] python tools/ab.py cfg2 [--spp 63] [--reps 5] [--profile]

Prints the best and median frame time, Mrays/s and (with --profile) the summed extend / shade kernel times of one
profiling pass.  Used for the experiment tables of profiles/README.md."""
import argparse
import ctypes as C
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scheme_raytrace_b200 as srt                     # noqa: E402
from scheme_raytrace_b200.host import ffi, scenes      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--wave-spp", type=int, default=0)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    cfg = scenes.CONFIGS[a.workload]
    w, h, spp = cfg["width"], cfg["height"], a.spp or cfg["spp"]
    r = srt.Renderer(cfg["scene"](w, h), device=0)
    import torch
    acc = torch.zeros(h, w, 3, dtype=torch.float32, device="cuda")
    ms, rays, st = [], 0, None
    for k in range(a.reps + 1):
        acc.zero_()
        st = r.render_device(acc.data_ptr(), w, h, spp, max_depth=cfg["max_depth"], seed=cfg["seed"], wave_spp=a.wave_spp)
        if k or a.reps == 0:
            ms.append(st.ms_total)
        rays = st.rays
    ms.sort()
    line = f"{a.tag or os.environ.get('SRT_LIB', 'libsrt.so'):28s} {a.workload:12s} spp={spp:5d} best {ms[0]:9.3f} ms  median {ms[len(ms) // 2]:9.3f} ms  {rays / ms[0] / 1e3:8.0f} Mrays/s  launches {st.kernel_launches} tail_runs {st.tail_runs} pipes {st.pipes} checksum {float(acc.double().sum()):.6e}"
    if a.profile:
        p = r.params(w, h, 0, min(spp, max(1, (128 << 20) // (w * h))), cfg["max_depth"], cfg["seed"])
        p.reserved[0] = 1
        pst = ffi.Stats()
        acc.zero_()
        ffi.check(r.lib.srt_render_device(r.h, C.byref(p), C.c_void_p(acc.data_ptr()), C.byref(pst)), "profile")
        line += f"  | profile: extend {pst.ms_extend:8.3f} ms shade {pst.ms_shade:8.3f} ms over {pst.rays} rays ({pst.rays / max(pst.ms_extend, 1e-9) / 1e3:.0f} / {pst.rays / max(pst.ms_shade, 1e-9) / 1e3:.0f} Mrays/s)"
    print(line, flush=True)
    r.close()


if __name__ == "__main__":
    main()
