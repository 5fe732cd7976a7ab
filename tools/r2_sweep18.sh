mkdir -p gpurun_out
O=gpurun_out/r2_sweep18.txt; : > $O
for ws in 9 17 35 0 105 140 280; do
  python tools/ab.py cfg2 --reps 3 --wave-spp $ws --tag "wave_spp=$ws" >> $O 2>&1
done
for ws in 26 52 0 157 210 420; do
  python tools/ab.py cfg3 --reps 3 --wave-spp $ws --tag "wave_spp=$ws" >> $O 2>&1
done
for ws in 16 32 0 128; do
  python tools/ab.py cfg4 --spp 512 --reps 2 --wave-spp $ws --tag "wave_spp=$ws" >> $O 2>&1
done
cut -c1-200 $O
