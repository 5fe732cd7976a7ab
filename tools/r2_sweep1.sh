# round-2 A/B sweep 1: staged planes + FFMA2 + fp32 leaf spheres, tail threshold (run on the GPU box)
mkdir -p gpurun_out
O=gpurun_out/r2_sweep1.txt; : > $O
for w in cfg2 cfg3 cfg4 cfg1; do
  python tools/ab.py $w --spp 64 --profile --tag "new" >> $O 2>&1
  SRT_NO_LEAF32=1 python tools/ab.py $w --spp 64 --profile --tag "new,leaf64" >> $O 2>&1
done
python tools/ab.py cfg2 --profile --tag "new full" >> $O 2>&1
for t in 0 600000 2000000 4000000 8000000 16000000 70000000; do
  SRT_TAIL_MAX=$t python tools/ab.py cfg2 --spp 63 --reps 7 --tag "tail_max=$t" >> $O 2>&1
done
for t in 0 4000000 70000000; do
  SRT_TAIL_MAX=$t python tools/ab.py cfg3 --spp 125 --reps 5 --tag "tail_max=$t" >> $O 2>&1
  SRT_TAIL_MAX=$t python tools/ab.py cfg4 --spp 64 --reps 5 --tag "tail_max=$t" >> $O 2>&1
  SRT_TAIL_MAX=$t python tools/ab.py cfg1 --reps 9 --tag "tail_max=$t" >> $O 2>&1
done
cat $O
