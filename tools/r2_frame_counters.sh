# whole-frame instruction counters of extend / shade (every launch of a 192-spp cfg2 / cfg3 frame), tile-major vs scanline path order
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
for w in cfg2 cfg3; do
for lib in "" exp/libsrt_untiled.so; do
  tag=${lib:-tiled}; tag=$(basename $tag .so); tag=${tag#libsrt_}
  SRT_LIB=$lib ncu --clock-control none --metrics $M -k regex:"k_extend|k_shade" -c 400 --csv --log-file gpurun_out/r2_frame_${w}_$tag.csv python tools/ab.py $w --spp 192 --reps 0 > gpurun_out/r2_frame_${w}_$tag.log 2>&1
done; done
ls -la gpurun_out | tail -6
