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
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample, "all_threads_per_core": v / max(cores, 1)},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


PER_CONFIG = ["cfg1", "cfg3", "cfg4", "cfg5", "cfg5_teapot", "cfg5_curves"]     # every other BASELINE config, one full-size frame each


def _fp32_peak(r):
    import ctypes as C
    from scheme_raytrace_b200.host import ffi
    tfl = C.c_float(0.0)
    ffi.check(r.lib.srt_measure_fp32_peak(C.byref(tfl)), "fp32 peak")
    return float(tfl.value)


def roofline_of(r, name, W, H, D, SEED, s_begin, s_end, accum, value_mrays, peak, peak_kind, fp32_peak):
    """Roofline of the dominant kernel (extend), measured live with CUDA events around every extend / shade launch of
    one extra profiling pass (about two queue fills of the workload's own frame, so most launches are full-queue
    iterations)."""
    import ctypes as C
    from scheme_raytrace_b200.host import ffi
    prof_spp = max(1, min(s_end - s_begin, (128 << 20) // (W * H) or 1))
    accum.zero_()
    p = r.params(W, H, s_begin, s_begin + prof_spp, D, SEED)
    p.reserved[0] = 1
    pst = ffi.Stats()
    ffi.check(r.lib.srt_render_device(r.h, C.byref(p), C.c_void_p(accum.data_ptr()), C.byref(pst)), "profile pass")
    ext_ms = pst.ms_extend
    n_ext = max(pst.extend_launches, 1)
    ext_gbs = (B_PER_RAY_EXTEND * pst.rays) / (ext_ms * 1e-3) / 1e9 if ext_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": "k_extend (LBVH closest hit)", "achieved": ext_gbs, "peak": peak, "unit": "GB/s",
            "frac": ext_gbs / peak, "traffic": NCU_EXTEND_DRAM_B_PER_RAY * pst.rays / n_ext, "peak_kind": peak_kind,
            "traffic_note": "bytes per launch = ncu-measured 47.6 B/ray (profiles/r1_traffic.txt) x rays per launch of this pass",
            "bytes_per_ray": B_PER_RAY_EXTEND, "rays_per_launch": pst.rays / n_ext, "avg_launch_ms": ext_ms / n_ext,
            "extend_share_of_loop": ext_ms / max(ext_ms + pst.ms_shade, 1e-9),
            "loop_hbm_gbs": B_PER_RAY_LOOP * value_mrays * 1e6 / 1e9, "loop_hbm_frac": B_PER_RAY_LOOP * value_mrays * 1e6 / 1e9 / peak,
            "rays_per_bounce": [int(x) for x in pst.rays_per_bounce],
            "note": "scenes fit in shared memory, so the HBM fraction is small by construction (SURVEY 8d); "
                    "the binding resource is fp32 issue + latency under divergence"}
    ns, nt, fp, fs = FP32_MODEL.get(name, FP32_MODEL["cfg2"])
    flops_per_ray = 27 * 2 * ns + 3 + fp * nt + fs
    ext_rays_s = pst.rays / (ext_ms * 1e-3) if ext_ms > 0 else 0.0
    roof["fp32"] = {"flops_per_ray": flops_per_ray, "model": {"node_steps": ns, "prim_tests": nt, "prim_flops": fp, "shade_flops": fs},
                    "achieved_tflops_loop": value_mrays * 1e6 * flops_per_ray / 1e12,
                    "achieved_tflops_extend": ext_rays_s * (flops_per_ray - fs) / 1e12,
                    "peak_tflops": fp32_peak, "peak_kind": "measured FFMA microbenchmark (srt_measure_fp32_peak)",
                    "frac_loop": value_mrays * 1e6 * flops_per_ray / 1e12 / max(fp32_peak, 1e-9)}
    return roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (invalidates the headline config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-per-config", action="store_true", help="skip the one-frame-per-config block (cfg1/3/4/5...)")
    ap.add_argument("--wave-spp", type=int, default=0, help="path-queue capacity in samples/pixel (0 = library default, 64 Mi paths)")
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
    dist = cpu_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")          # host-side barrier: ranks wait here WITHOUT a kernel spinning on their GPU
    from scheme_raytrace_b200.host import sharding

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(name, c, steps, warm, timed_e2e_steps):
        """One workload on this rank's sample range: `steps` device-timed frames (scene resident, one NCCL reduce per
        frame), then `timed_e2e_steps` end-to-end frames (commit = H2D + LBVH build, render, reduce, D2H into pinned
        host memory).  Returns the max-over-ranks times and the summed counters."""
        W, H, SPP, D, SEED = c["width"], c["height"], c["spp"], c["max_depth"], c["seed"]
        s_begin, s_end = sharding.sample_range(rank, world, SPP)            # sample-range sharding
        flat = srt.flatten_scene(c["scene"](W, H))
        r = srt.Renderer(flat, device=local)
        accum = torch.zeros(H, W, 3, dtype=torch.float32, device="cuda")
        host_img = torch.empty(H, W, 3, dtype=torch.float32).pin_memory()

        def step_device(b=s_begin, e=s_end):
            accum.zero_()
            st = r.render_device(accum.data_ptr(), W, H, e - b, max_depth=D, seed=SEED, spp_begin=b, wave_spp=args.wave_spp)
            sharding.reduce_accumulators(accum, dist)
            return st

        def step_e2e():
            r.commit()
            st = step_device()
            if rank == 0:
                host_img.copy_(accum, non_blocking=True)
            torch.cuda.synchronize()
            return st
        # warm-up: full frames for the headline workload; for the one-frame configs enough samples to fill the 64 Mi-path
        # queue once (allocations, kernel variants, the cached iteration graph)
        warm_spp = (s_end - s_begin) if steps > 1 else max(4, -(-(64 << 20) // (W * H)))
        for _ in range(warm):
            step_device(s_begin, min(s_end, s_begin + warm_spp))
        sampler = ClockSampler(local) if name == args.workload else None
        barrier()
        if sampler and rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        rays = launches = 0
        for _ in range(steps):
            st = step_device()
            rays += st.rays; launches += st.kernel_launches
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        clocks = sampler.stop() if (sampler and rank == 0) else None
        if timed_e2e_steps and steps > 1:
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        rays_e2e = 0
        for _ in range(timed_e2e_steps):
            st = step_e2e()
            rays_e2e += st.rays
        barrier()
        e2e_s = time.perf_counter() - t0
        tt = torch.tensor([ms, e2e_s, float(rays), float(rays_e2e), float(launches), float(st.tail_runs)], dtype=torch.float64, device="cuda")
        if dist is not None:
            mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            tt = torch.stack([mx[0], mx[1], sm[2], sm[3], sm[4], mx[5]])
        ms, e2e_s, rays, rays_e2e, launches, tails = (float(x) for x in tt.cpu())
        return dict(r=r, flat=flat, accum=accum, ms=ms, e2e_s=e2e_s, rays=rays, rays_e2e=rays_e2e, launches=launches, clocks=clocks,
                    range=(s_begin, s_end), dims=(W, H, SPP, D, SEED), commit_ms=st.ms_commit, tail_runs=tails, host_img=host_img)

    peak, peak_kind = read_peaks()
    K = args.steps
    m = measure(args.workload, cfg, K, max(args.warmup, 3), K)
    W, H, SPP, D, SEED = m["dims"]
    r, flat = m["r"], m["flat"]
    value = m["rays"] / (m["ms"] * 1e-3) / 1e6
    e2e_value = m["rays_e2e"] / m["e2e_s"] / 1e6
    e2e = {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": flat.h2d_bytes() + 3072 + 3 * 1024, "d2h_bytes_per_step": W * H * 3 * 4,
           "sec_per_frame": m["e2e_s"] / max(K, 1), "path": "one process per GPU: Renderer.commit + srt_render_device + NCCL reduce + D2H (pinned)"}
    # multi-GPU image check: rank 0 renders the whole frame alone and compares it with the reduced frame of the N ranks
    image_check = None
    if world > 1:
        reduced = m["accum"].clone() if rank == 0 else None
        if rank == 0:
            dist.barrier(group=cpu_group)            # the other ranks wait on the host, their GPUs idle
            alone = torch.zeros_like(m["accum"])
            r.render_device(alone.data_ptr(), W, H, SPP, max_depth=D, seed=SEED)
            dmax = float((alone - reduced).abs().max())
            image_check = {"max_abs_diff": dmax, "bound_1e-4_x_spp": 1e-4 * SPP, "ok": dmax <= 1e-4 * SPP,
                           "what": f"{world}-rank reduced radiance sums (cfg frame, float reduce) vs the same frame rendered by rank 0 alone"}
            dist.barrier(group=cpu_group)
        else:
            dist.barrier(group=cpu_group); dist.barrier(group=cpu_group)
    # the drop-in call from ONE process: srt_render_multi over all N GPUs with host buffers (rank 0; the others wait on the host)
    if world > 1:
        one = None
        if rank == 0:
            dist.barrier(group=cpu_group)
            try:
                rm = srt.Renderer(flat, gpus=world)
                host_sum = torch.empty(H, W, 3, dtype=torch.float32).pin_memory().numpy()      # *raw-data* / *image* land in pinned host memory
                host_img8 = torch.empty(H, W, 3, dtype=torch.uint8).pin_memory().numpy()
                rm.render_multi(W, H, SPP, max_depth=D, seed=SEED, rgb_sum=host_sum, image=host_img8, write_only=True)      # warm-up: replicas, graphs, NCCL channels
                t0 = time.perf_counter(); nr = 0
                for _ in range(K):
                    rm.commit()
                    _, _, stm = rm.render_multi(W, H, SPP, max_depth=D, seed=SEED, image=host_img8, want_sum=False)     # the step's result: the 8-bit frame (*image*)
                    nr += stm.rays
                dt = time.perf_counter() - t0
                import ctypes as C
                ver = C.c_int32(0)
                mode = rm.lib.srt_multi_reduce_mode(C.byref(ver))
                full, _, _ = rm.render_multi(W, H, SPP, max_depth=D, seed=SEED)
                alone = np.zeros((H, W, 3), dtype=np.float32)
                r.render(W, H, SPP, max_depth=D, seed=SEED, rgb_sum=alone)
                one = {"value": nr / dt / 1e6, "unit": "Mrays/s", "sec_per_frame": dt / K, "gpus": rm.gpus,
                       "reduce": "ncclReduce(uint64 accumulators)" if mode == 0 else "peer-read reduce kernel over NVLink", "nccl_version": ver.value,
                       "bit_identical_to_one_gpu": bool(np.array_equal(full, alone)),
                       "h2d_bytes_per_step": (flat.h2d_bytes() + 3072 + 3 * 1024) * world, "d2h_bytes_per_step": W * H * 3}
                rm.close()
            except Exception as ex:                       # never lose the bench line to the optional leg
                one = {"error": str(ex)[:300]}
            dist.barrier(group=cpu_group)
        else:
            dist.barrier(group=cpu_group); dist.barrier(group=cpu_group)
        if rank == 0 and one and "value" in one:
            e2e["one_process"] = one
            if one["value"] > 0:
                e2e.update({"value": one["value"], "sec_per_frame": one["sec_per_frame"], "h2d_bytes_per_step": one["h2d_bytes_per_step"],
                            "d2h_bytes_per_step": one["d2h_bytes_per_step"], "torchrun_ranks_value": e2e_value,
                            "path": "ONE process, host buffers: srt_scene_commit (every GPU) + srt_render_multi (sample ranges, one reduce over NVLink, 8-bit + float frame D2H)"})
        elif rank == 0:
            e2e["one_process"] = one
    fp32_peak = _fp32_peak(r) if rank == 0 else 0.0
    roofline = roofline_of(r, args.workload, W, H, D, SEED, m["range"][0], m["range"][1], m["accum"], value / world, peak, peak_kind, fp32_peak) if rank == 0 else None
    if world > 1:
        dist.barrier(group=cpu_group)
    # ---- every other BASELINE config: one full-size frame each (sample-sharded over the ranks) ------------------------------
    per_config = {}
    if not args.no_per_config and args.spp == 0:
        del m["accum"]; m["r"].close(); torch.cuda.empty_cache()
        for name in PER_CONFIG:
            if name == args.workload:
                continue
            c = workload(name)
            # one full-size frame each; a frame of well under a millisecond (cfg1) is timed as the mean of 16 after 3 warm-up frames
            nf = 16 if c["width"] * c["height"] * c["spp"] < (1 << 22) else 1
            q = measure(name, c, nf, 3 if nf > 1 else 1, nf)
            w_, h_, spp_, d_, seed_ = q["dims"]
            v = q["rays"] / (q["ms"] * 1e-3) / 1e6
            entry = {"workload": f"{name}: {w_}x{h_} @ {spp_} spp, depth {d_}, {len(q['flat'].prims)} primitives", "Mrays/s": v, "ms_per_frame": q["ms"] / nf, "frames": nf,
                     "rays_per_frame": q["rays"] / nf, "e2e": {"Mrays/s": q["rays_e2e"] / q["e2e_s"] / 1e6, "sec_per_frame": q["e2e_s"] / nf, "commit_ms": q["commit_ms"]},
                     "gpu_launches": int(q["launches"] / nf), "tail_runs": int(q["tail_runs"])}
            if rank == 0:
                entry["roofline"] = roofline_of(q["r"], name, w_, h_, d_, seed_, q["range"][0], q["range"][1], q["accum"], v / world, peak, peak_kind, fp32_peak)
            if world > 1:
                dist.barrier(group=cpu_group)
            per_config[name] = entry
            q["r"].close(); del q; torch.cuda.empty_cache()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_oracle_rate(cfg, args.workload)
        one_core = cpu_oracle_rate(cfg, args.workload, budget_s=6.0, nthreads=1)
        cpu["per_core"] = {"value": one_core["value"], "unit": "Mrays/s", "cores": 1, "sample": one_core["sample"],
                           "all_threads_per_core": cpu["value"] / max(cpu["cores"], 1)}
    frame_rays = m["rays"] / max(K, 1)
    line = {"metric": "Mrays/s (closest-hit queries per second, primary + every bounce)", "value": value, "unit": "Mrays/s",
            "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": m["ms"] / max(K, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {W}x{H} @ {SPP} spp, depth {D}, {len(flat.prims)} primitives, GPU LBVH, quirks=REFERENCE",
                       "sharding": f"sample range, {world} rank(s)", "l2": "per-step working set (ray/hit queues ~1 GB per wave) exceeds the 126 MB L2",
                       "rays_per_frame": frame_rays, "sec_per_frame": m["ms"] / max(K, 1) * 1e-3, "commit_ms": m["commit_ms"], "tail_runs_last_frame": int(m["tail_runs"])},
            "e2e": e2e, "gpu_launches": int(m["launches"]), "clocks": m["clocks"], "roofline": roofline}
    if image_check is not None:
        line["multi_gpu_image_check"] = image_check
    if per_config:
        line["per_config"] = per_config
    if cpu is not None:
        cpu["sec_per_frame_extrapolated"] = frame_rays / (cpu["value"] * 1e6)
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
