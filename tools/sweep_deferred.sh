mkdir -p gpurun_out
run() { # workload vote refill
  SRT_PARK_VOTE=$2 SRT_REFILL_MIN=$3 timeout 100 python bench.py --workload $1 --spp 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s_$1_$2_$3.json 2>/dev/null
  python -c "import json; d=json.load(open('gpurun_out/s_$1_$2_$3.json')); print('$1 vote $2 refill $3:', round(d['value']), round(d['ms_per_step'],1))"
}
for w in cfg5 cfg5_curves; do run $w 12 32; run $w 12 8; run $w 16 16; run $w 16 12; done
run cfg5_teapot 16 16; run cfg5_teapot 16 12; run cfg5_teapot 20 16; run cfg5_teapot 16 24
