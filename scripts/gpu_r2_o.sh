# round 2, session 2, call 12: new edge-case test, compute-sanitizer (memcheck + racecheck) on the
# kernels with the explicit shared-window addressing
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ragged" > $O/r2o_ragged.log 2>&1
tail -4 $O/r2o_ragged.log
export CUDA_VISIBLE_DEVICES=0
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -m gpu -q -x \
   -k "golden or refine or plus_kernel or run_integration_reproduces or vegasflowplus_reproduces or iteration_epilogue or accumulate or same_stream" > $O/r2_sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> $O/r2_sanitizer_memcheck.log
tail -6 $O/r2_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -m gpu -q -x \
   -k "refine or iteration_epilogue or plus_kernel_against_golden or c1-symgauss or plus3a" > $O/r2_sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> $O/r2_sanitizer_racecheck.log
tail -6 $O/r2_sanitizer_racecheck.log
