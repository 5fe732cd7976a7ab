# round-2 A/B sweep 3: heavy variants (slab pre-cull), Perlin alias fast path, regression hunt for cfg5 against the round-1 library
mkdir -p gpurun_out
O=gpurun_out/r2_sweep3.txt; : > $O
for w in cfg5 cfg5_teapot cfg5_curves; do
  for v in r1 head; do SRT_LIB=$PWD/exp/libsrt_$v.so python tools/ab.py $w --spp 32 --reps 3 --tag "$v" >> $O 2>&1; done
  python tools/ab.py $w --spp 32 --reps 3 --profile --tag "main" >> $O 2>&1
done
for w in cfg3 cfg2 cfg4 cfg1; do
  SRT_LIB=$PWD/exp/libsrt_r1.so python tools/ab.py $w --spp 128 --tag "r1" >> $O 2>&1
  python tools/ab.py $w --spp 128 --profile --tag "main" >> $O 2>&1
done
cat $O
