# round-2 A/B sweep 4: one vs two concurrent pipelines
mkdir -p gpurun_out
O=gpurun_out/r2_sweep4.txt; : > $O
for w in cfg2 cfg3 cfg4; do
  for n in 1 2; do SRT_PIPES=$n python tools/ab.py $w --reps 3 --tag "pipes=$n" >> $O 2>&1; done
done
for n in 1 2; do SRT_PIPES=$n python tools/ab.py cfg2 --spp 63 --reps 5 --tag "pipes=$n" >> $O 2>&1; done
for w in cfg5_teapot; do SRT_LIB=$PWD/exp/libsrt_r1.so python tools/ab.py $w --spp 32 --reps 3 --tag "r1" >> $O 2>&1; done
for w in cfg5; do SRT_LIB=$PWD/exp/libsrt_r1.so python tools/ab.py $w --spp 32 --reps 3 --tag "r1" >> $O 2>&1; done
cat $O
