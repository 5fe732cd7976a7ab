mkdir -p gpurun_out
O=gpurun_out/r2_sweep6.txt; : > $O
SRT_LIB=$PWD/exp/libsrt_r1.so python tools/ab.py cfg5 --spp 32 --reps 2 --profile --tag "r1" >> $O 2>&1
python tools/ab.py cfg5 --spp 32 --reps 2 --profile --tag "main" >> $O 2>&1
SRT_PARK_VOTE=12 python tools/ab.py cfg5 --spp 32 --reps 2 --tag "main vote12" >> $O 2>&1
SRT_PARK_VOTE=20 python tools/ab.py cfg5 --spp 32 --reps 2 --tag "main vote20" >> $O 2>&1
SRT_PARK_VOTE=12 python tools/ab.py cfg5_teapot --spp 32 --reps 2 --tag "main vote12" >> $O 2>&1
SRT_PARK_VOTE=20 python tools/ab.py cfg5_teapot --spp 32 --reps 2 --tag "main vote20" >> $O 2>&1
cut -c1-330 $O
