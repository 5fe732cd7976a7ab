# round-2 A/B sweep 2: node layout / FFMA2, shade tables + prefetch, lane advance, tail threshold (run on the GPU box)
mkdir -p gpurun_out
O=gpurun_out/r2_sweep2.txt; : > $O
for w in cfg2 cfg3 cfg4; do
  for v in head aos; do SRT_LIB=$PWD/exp/libsrt_$v.so python tools/ab.py $w --spp 128 --profile --tag "$v" >> $O 2>&1; done
  python tools/ab.py $w --spp 128 --profile --tag "main" >> $O 2>&1
  SRT_NO_SHADE_TABS=1 python tools/ab.py $w --spp 128 --profile --tag "main,no-tabs" >> $O 2>&1
  SRT_NO_LEAF32=1 python tools/ab.py $w --spp 128 --profile --tag "main,leaf64" >> $O 2>&1
  for v in pf3 pf4 adv8 adv16 adv24; do SRT_LIB=$PWD/exp/libsrt_$v.so python tools/ab.py $w --spp 128 --profile --tag "$v" >> $O 2>&1; done
done
for t in 0 150000 300000 600000 1000000 2000000; do
  SRT_TAIL_MAX=$t python tools/ab.py cfg2 --spp 63 --reps 7 --tag "tail_max=$t" >> $O 2>&1
  SRT_TAIL_MAX=$t python tools/ab.py cfg3 --spp 125 --reps 7 --tag "tail_max=$t" >> $O 2>&1
  SRT_TAIL_MAX=$t python tools/ab.py cfg4 --spp 64 --reps 7 --tag "tail_max=$t" >> $O 2>&1
done
cat $O
