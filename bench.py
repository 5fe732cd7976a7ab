#!/usr/bin/env python
"""bench.py — headline benchmark of the per-sample radiance loop (BASELINE.json metric: Mrays/s;
sec per 1200x800@500spp frame vs CPU).

One "step" = one full frame of the workload: every sample of every pixel through ray generation,
LBVH closest-hit, shading/scatter and accumulation.  Default workload = BASELINE.json configs[1]
(Weekend final random-spheres scene, 1200x800, 500 spp, depth 50, GPU LBVH).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the frame is sharded by sample range and
the per-GPU float accumulation buffers are combined with one NCCL reduce (strong scaling: the
frame is fixed, so per-GPU work shrinks as N grows).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_RAY_LOOP = 160      # ray w+r 2x32 + ray re-read by shade 32 ... = 160 B per ray-bounce (whole loop)
B_PER_RAY_EXTEND = 48     # the extend kernel's part: ray read 32 B + hit record write 16 B
NCU_EXTEND_DRAM_B_PER_RAY = 47.6   # measured: ncu dram__bytes_read+write of a 61.44M-ray launch (profiles/r1_traffic.txt)
# FP32 work model: flops/ray = 27*N_box + 3 + sum F_type*N_type + F_shade, with the
# per-ray counts MEASURED by the instrumented build (tools/step_stats.py, profiles/r1_step_stats.txt):
# node steps per ray (2 box tests each), primitive tests per ray, dominant primitive cost, shade cost.
FP32_MODEL = {  # workload: (node steps/ray, prim tests/ray, flops per prim test, shade flops/ray)
    # (profiles/r1_step_stats.txt; prim tests include the primitives tested before traversal)
    "cfg1": (2.0, 2.4, 35, 110), "cfg2": (7.8, 4.6, 35, 100), "cfg3": (8.9, 3.8, 45, 160),
    "cfg4": (2.8, 8.3, 24, 100), "cfg5": (11.0, 2.3, 600, 100), "cfg5_teapot": (11.0, 2.3, 600, 100), "cfg5_curves": (3.0, 1.7, 1500, 100),
}


def read_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(name):
    import scheme_raytrace_b200 as srt
    cfg = dict(srt.scenes.CONFIGS[name])
    return cfg


def cpu_oracle_rate(cfg, name, budget_s=12.0, nthreads=0):
    """Times the oracle (CPU restatement, linear hit-obj-list like the reference) on a bounded
    sample of the same workload: the same scene at 1/8 x 1/8 resolution, spp scaled to ~budget_s."""
    from oracle import oracle as O
    O.build()
    w, h = max(cfg["width"] // 8, 8), max(cfg["height"] // 8, 8)
    scene = cfg["scene"](w, h)
    S = O.OracleScene(scene)
    cores = nthreads if nthreads > 0 else (os.cpu_count() or 1)
    t0 = time.perf_counter(); _, nr = S.render(w, h, 1, max_depth=cfg["max_depth"], seed=cfg["seed"], nthreads=cores); dt1 = time.perf_counter() - t0
    spp = int(max(1, min(1024, budget_s / max(dt1, 1e-3))))      # ~budget_s seconds of CPU work on all host threads
    t0 = time.perf_counter(); _, nrays = S.render(w, h, spp, max_depth=cfg["max_depth"], seed=cfg["seed"], nthreads=cores); dt = time.perf_counter() - t0
    mrays = nrays / dt / 1e6
    return dict(value=mrays, unit="Mrays/s", cores=cores, kind="port",
                sample=f"{name} scene at {w}x{h}, {spp} spp, depth {cfg['max_depth']} ({nrays} rays in {dt:.2f} s); "
                       "C++ f64 restatement of the .scm files (oracle/), linear hit-obj-list like the reference; not Gauche",
                seconds=dt, rays=nrays, rays_per_path=nrays / (w * h * spp))


def run_reference(args, cfg):
    """--impl reference: the reference's own CPU implementation of the path.  No Scheme runtime
    exists in this image, so this is the oracle port on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    w, h = max(cfg["width"] // 8, 8), max(cfg["height"] // 8, 8)
    scene = cfg["scene"](w, h)
    S = O.OracleScene(scene)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter(); S.render(w, h, 1, max_depth=cfg["max_depth"], seed=cfg["seed"], nthreads=cores); dt1 = time.perf_counter() - t0
    total = max(args.steps + args.warmup, 1)
    spp = int(max(1, min(256, (90.0 / total) / max(dt1, 1e-3))))     # a few seconds of CPU work per step
    for _ in range(args.warmup):
        S.render(w, h, spp, max_depth=cfg["max_depth"], seed=cfg["seed"], nthreads=cores)
    t0 = time.perf_counter(); nrays = 0
    for _ in range(args.steps):
        _, nr = S.render(w, h, spp, max_depth=cfg["max_depth"], seed=cfg["seed"], nthreads=cores); nrays += nr
    dt = time.perf_counter() - t0
    v = nrays / dt / 1e6
    sample = f"{args.workload} scene at {w}x{h}, {spp} spp per step (bounded sample of the {cfg['width']}x{cfg['height']}@{cfg['spp']}spp frame)"
    line = {"impl": "reference", "metric": "Mrays/s (closest-hit queries per second, primary + every bounce)", "value": v, "unit": "Mrays/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {cfg['width']}x{cfg['height']} @ {cfg['spp']} spp, depth {cfg['max_depth']}", "sample": sample},
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (invalidates the headline config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wave-spp", type=int, default=0, help="path-queue capacity in samples/pixel (0 = library default, ~8M paths)")
    args = ap.parse_args()
    cfg = workload(args.workload)
    if args.spp > 0:
        cfg["spp"] = args.spp
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import numpy as np
    import torch
    import scheme_raytrace_b200 as srt
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the radiance loop has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H, SPP, D, SEED = cfg["width"], cfg["height"], cfg["spp"], cfg["max_depth"], cfg["seed"]
    from scheme_raytrace_b200.host import sharding
    s_begin, s_end = sharding.sample_range(rank, world, SPP)            # sample-range sharding
    scene = cfg["scene"](W, H)
    flat = srt.flatten_scene(scene)
    r = srt.Renderer(flat, device=local)
    accum = torch.zeros(H, W, 3, dtype=torch.float32, device="cuda")
    host_img = torch.empty(H, W, 3, dtype=torch.float32).pin_memory()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        """value: scene resident in HBM, accumulate into the device buffer, one NCCL reduce."""
        accum.zero_()
        st = r.render_device(accum.data_ptr(), W, H, s_end - s_begin, max_depth=D, seed=SEED, spp_begin=s_begin, wave_spp=args.wave_spp)
        sharding.reduce_accumulators(accum, dist)
        return st

    def step_e2e():
        """e2e: host tables in -> (H2D + LBVH build) -> render -> reduce -> image back in host memory."""
        r.commit()
        accum.zero_()
        st = r.render_device(accum.data_ptr(), W, H, s_end - s_begin, max_depth=D, seed=SEED, spp_begin=s_begin)
        sharding.reduce_accumulators(accum, dist)
        if rank == 0:
            host_img.copy_(accum, non_blocking=True)
        torch.cuda.synchronize()
        return st

    for _ in range(max(args.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    rays = launches = 0
    for _ in range(args.steps):
        st = step_device()
        rays += st.rays; launches += st.kernel_launches
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    # end-to-end steps (1 warm-up + K timed, wall clock around the public API incl. H2D/D2H)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    rays_e2e = 0
    for _ in range(args.steps):
        rays_e2e += step_e2e().rays
    barrier()
    e2e_s = time.perf_counter() - t0
    tt = torch.tensor([ms, e2e_s, float(rays), float(rays_e2e), float(launches)], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_s = float(mx[0]), float(mx[1]); rays, rays_e2e, launches = float(sm[2]), float(sm[3]), float(sm[4])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    value = rays / (ms * 1e-3) / 1e6
    e2e_value = rays_e2e / e2e_s / 1e6
    # roofline of the dominant kernel (extend), measured live with CUDA events around every
    # extend launch of one extra profiling pass (events on the launching stream, stream 0)
    prof_spp = max(1, min(s_end - s_begin, 128))     # enough samples for steady-state iterations (full queue + regeneration)
    accum.zero_()
    p = r.params(W, H, s_begin, s_begin + prof_spp, D, SEED)
    p.reserved[0] = 1
    import ctypes as C
    from scheme_raytrace_b200.host import ffi
    pst = ffi.Stats()
    ffi.check(r.lib.srt_render_device(r.h, C.byref(p), C.c_void_p(accum.data_ptr()), C.byref(pst)), "profile pass")
    peak, peak_kind = read_peaks()
    ext_ms = pst.ms_extend
    ext_gbs = (B_PER_RAY_EXTEND * pst.rays) / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_extend (LBVH closest hit)", "achieved": ext_gbs, "peak": peak, "unit": "GB/s",
                "frac": ext_gbs / peak, "traffic": NCU_EXTEND_DRAM_B_PER_RAY * pst.rays / max(pst.extend_launches, 1), "peak_kind": peak_kind,
                "traffic_note": "bytes per launch = ncu-measured 47.6 B/ray (profiles/r1_traffic.txt) x rays per launch of this pass",
                "bytes_per_ray": B_PER_RAY_EXTEND, "rays_per_launch": pst.rays / max(pst.extend_launches, 1),
                "avg_launch_ms": ext_ms / max(pst.extend_launches, 1),
                "extend_share_of_loop": ext_ms / max(ext_ms + pst.ms_shade, 1e-9),
                "loop_hbm_gbs": B_PER_RAY_LOOP * value * 1e6 / 1e9,
                "note": "scenes fit in shared memory, so the HBM fraction is small by construction (SURVEY §8d); "
                        "the binding resource is fp32 issue + latency under divergence"}
    ns, nt, fp, fs = FP32_MODEL.get(args.workload, FP32_MODEL["cfg2"])
    flops_per_ray = 27 * 2 * ns + 3 + fp * nt + fs
    tfl = C.c_float(0.0)
    ffi.check(r.lib.srt_measure_fp32_peak(C.byref(tfl)), "fp32 peak")
    ext_rays_s = pst.rays / (ext_ms * 1e-3) if ext_ms > 0 else 0.0
    roofline["fp32"] = {"flops_per_ray": flops_per_ray, "model": {"node_steps": ns, "prim_tests": nt, "prim_flops": fp, "shade_flops": fs},
                        "achieved_tflops_loop": value * 1e6 * flops_per_ray / 1e12,
                        "achieved_tflops_extend": ext_rays_s * (flops_per_ray - fs) / 1e12,
                        "peak_tflops": float(tfl.value), "peak_kind": "measured FFMA microbenchmark (srt_measure_fp32_peak)",
                        "frac_loop": value * 1e6 * flops_per_ray / 1e12 / max(float(tfl.value), 1e-9)}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_oracle_rate(cfg, args.workload)
    frame_rays = rays / max(args.steps, 1)
    line = {"metric": "Mrays/s (closest-hit queries per second, primary + every bounce)", "value": value, "unit": "Mrays/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {W}x{H} @ {SPP} spp, depth {D}, {len(flat.prims)} primitives, GPU LBVH, quirks=REFERENCE",
                       "sharding": f"sample range, {world} rank(s)", "l2": "per-step working set (ray/hit queues ~1 GB per wave) exceeds the 126 MB L2",
                       "rays_per_frame": frame_rays, "sec_per_frame": ms / max(args.steps, 1) * 1e-3},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": flat.h2d_bytes() + 3072 + 3 * 1024,
                    "d2h_bytes_per_step": W * H * 3 * 4, "sec_per_frame": e2e_s / max(args.steps, 1)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if cpu is not None:
        cpu["sec_per_frame_extrapolated"] = frame_rays / (cpu["value"] * 1e6)
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
