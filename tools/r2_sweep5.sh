# round-2 sweep 5: cfg5 regression (instanced-sphere code out of line), 32-bit stack on the 70k cloud, commit latency, bounds-checked build
mkdir -p gpurun_out
O=gpurun_out/r2_sweep5.txt; : > $O
for w in cfg5 cfg5_teapot cfg5_curves; do
  SRT_LIB=$PWD/exp/libsrt_r1.so python tools/ab.py $w --spp 32 --reps 3 --tag "r1" >> $O 2>&1
  python tools/ab.py $w --spp 32 --reps 3 --tag "main" >> $O 2>&1
done
python tools/ncu_commit.py 70000 >> $O 2>&1
python tools/ncu_commit.py 5000 >> $O 2>&1
python - >> $O 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.')
import scheme_raytrace_b200 as srt
from scheme_raytrace_b200.host import scenes
for name in ("cfg1", "cfg2", "cfg3", "cfg4", "cfg5_teapot"):
    cfg = scenes.CONFIGS[name]
    r = srt.Renderer(cfg["scene"](cfg["width"], cfg["height"]), device=0)
    best_dev, best_wall = 1e9, 1e9
    for _ in range(20):
        t0 = time.perf_counter(); r.commit(); dt = (time.perf_counter() - t0) * 1e3
        img, st = r.render(64, 64, 1)
        best_dev, best_wall = min(best_dev, st.ms_commit), min(best_wall, dt)
    print(f"commit {name}: device {best_dev:.3f} ms, wall (set_* + commit) {best_wall:.3f} ms, {len(r.flat.prims)} prims")
    r.close()
PY
cat $O
bash tools/r2_bounds.sh
