# A/B of experimental builds of libsrt.so (exp/libsrt_<name>.so, built with -D switches) on the GPU box:
#   bash tools/sweep_libs.sh <workload> <name> [<name> ...]        (name "base" = the committed build)
mkdir -p gpurun_out
cp scheme_raytrace_b200/csrc/libsrt.so /tmp/libsrt_base.so
w=$1; shift
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/libsrt_base.so scheme_raytrace_b200/csrc/libsrt.so; else cp exp/libsrt_$v.so scheme_raytrace_b200/csrc/libsrt.so; fi
  timeout 100 python bench.py --workload $w --spp 64 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/x_${w}_$v.json 2>gpurun_out/x_${w}_$v.err
  python -c "import json; d=json.load(open('gpurun_out/x_${w}_$v.json')); print('$w $v:', round(d['value']), round(d['ms_per_step'],1))" || tail -n 3 gpurun_out/x_${w}_$v.err
done
cp /tmp/libsrt_base.so scheme_raytrace_b200/csrc/libsrt.so
