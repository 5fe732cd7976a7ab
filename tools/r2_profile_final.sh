# final captures of the round (GPU box): launch list of the bench command + one steady-state iteration on cfg2 / cfg3
mkdir -p gpurun_out
N="ncu --clock-control none"
python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu-baseline --no-per-config > gpurun_out/r2_launchlist_bench.json 2>gpurun_out/r2_launchlist_bench.err && \
$N --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --spp 64 --no-cpu-baseline --no-per-config > gpurun_out/r2_ncu_launches.log 2>&1
for w in cfg2 cfg3; do
  python tools/ab.py $w --spp 192 --reps 0 > /dev/null 2>&1 && \
  $N --set full --import-source on -k regex:"k_extend|k_shade|k_regen" -s 12 -c 3 -f -o gpurun_out/r2_$w python tools/ab.py $w --spp 192 --reps 0 > gpurun_out/r2_ncu_$w.log 2>&1
done
for r in gpurun_out/r2_*.ncu-rep; do
  b=${r%.ncu-rep}
  ncu -i $r --page raw --csv > ${b}_raw.csv 2>/dev/null
  ncu -i $r --page source --csv > ${b}_source.csv 2>/dev/null
  rm -f $r
done
ls -la gpurun_out/ | tail -12
