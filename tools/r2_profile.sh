# round-2 evidence run (GPU box): compute-sanitizer summaries + ncu captures for every kernel variant
mkdir -p gpurun_out
# (compute-sanitizer is closed on this pool: see profiles/r2_sanitizer_unavailable.txt; the bounds-checked build is tools/r2_bounds.sh)
N="ncu --clock-control none"
# launch list of the default bench command (shares of the step)
python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu-baseline --no-per-config > gpurun_out/r2_launchlist_bench.json 2>gpurun_out/r2_launchlist_bench.err && \
$N --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu-baseline --no-per-config > gpurun_out/r2_ncu_launches.log 2>&1
# full captures: one steady-state iteration (extend, shade, regen) per scene
for w in cfg2 cfg3 cfg4; do
  python tools/ab.py $w --spp 192 --reps 0 > /dev/null 2>&1 && \
  $N --set full --import-source on -k regex:"k_extend|k_shade|k_regen" -s 12 -c 3 -f -o gpurun_out/r2_$w python tools/ab.py $w --spp 192 --reps 0 > gpurun_out/r2_ncu_$w.log 2>&1
done
python tools/ab.py cfg5_teapot --spp 16 --reps 0 > /dev/null 2>&1 && \
$N --set full --import-source on -k regex:"k_extend" -s 2 -c 1 -f -o gpurun_out/r2_cfg5_teapot python tools/ab.py cfg5_teapot --spp 16 --reps 0 > gpurun_out/r2_ncu_teapot.log 2>&1
python tools/ab.py cfg5_curves --spp 16 --reps 0 > /dev/null 2>&1 && \
$N --set full --import-source on -k regex:"k_extend" -s 2 -c 1 -f -o gpurun_out/r2_cfg5_curves python tools/ab.py cfg5_curves --spp 16 --reps 0 > gpurun_out/r2_ncu_curves.log 2>&1
# drain kernel (cfg1: the whole frame) and the LBVH build kernels + global-memory-tree extend (70k spheres)
$N --set full -k regex:"k_tail" -c 1 -f -o gpurun_out/r2_tail python tools/ab.py cfg1 --reps 0 > gpurun_out/r2_ncu_tail.log 2>&1
python tools/ncu_commit.py 70000 > gpurun_out/r2_cloud70k.txt 2>&1 && \
$N --set full -k regex:"k_prim_bounds|k_bounds_reduce|k_morton|k_rs_hist|k_rs_scan|k_rs_scatter|k_karras|k_refit|k_tree_area" -c 12 -f -o gpurun_out/r2_lbvh python tools/ncu_commit.py 70000 1 > gpurun_out/r2_ncu_lbvh.log 2>&1
$N --set full -k regex:"k_extend" -s 1 -c 1 -f -o gpurun_out/r2_cloud_extend python tools/ncu_commit.py 70000 4 > gpurun_out/r2_ncu_cloud.log 2>&1
# keep the evidence small: the raw counter page of every capture as CSV, the source page (per-instruction execution counts) for the
# captures taken with --import-source, then drop the binary reports (gpurun merges at most 64 MiB back)
for r in gpurun_out/r2_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
  case $b in *cfg2|*cfg3|*cfg4|*teapot|*curves) ncu -i $r --page source --csv > ${b}_source.csv 2>/dev/null;; esac
  rm -f $r
done
ls -la gpurun_out/ | tail -30
