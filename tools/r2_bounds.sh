# The sanitizer substitute (compute-sanitizer is closed on the B200 pool): the bounds-checked build of the library
# (-DSRT_BOUNDS_CHECK: every index the kernels trust is checked on the device, violations are counted and make the
# entry point fail) runs the sanitizer workload (every kernel variant, graph / no-graph / no-tail / profile paths,
# progressive passes, re-commit) and the GPU parity suite.
mkdir -p gpurun_out exp
# (exp/ is git-ignored: build the checked library here if it is not there or older than the sources)
C=scheme_raytrace_b200/csrc
if [ ! -e exp/libsrt_bounds.so ] || [ -n "$(find $C -newer exp/libsrt_bounds.so \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' \))" ]; then
  (cd $C && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DSRT_BOUNDS_CHECK -o ../../exp/libsrt_bounds.so srt_api.cu lbvh.cu wavefront.cu -ldl) || exit 1
fi
O=gpurun_out/r2_bounds_check.txt
echo "bounds-checked build (exp/libsrt_bounds.so, -DSRT_BOUNDS_CHECK) on $(nvidia-smi -L | head -1)" > $O
SRT_LIB=$PWD/exp/libsrt_bounds.so python tools/sanitize.py >> $O 2>&1; echo "sanitize.py rc=$?" >> $O
SRT_LIB=$PWD/exp/libsrt_bounds.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_reference.py -q -m gpu -x 2>&1 | tail -4 >> $O
cat $O
