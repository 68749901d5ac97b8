# compute-sanitizer on the VEGAS+ paths with the cluster tail
mkdir -p gpurun_out
O=gpurun_out
export CUDA_VISIBLE_DEVICES=0
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x \
   -k "plus or Plus" > $O/r2u_memcheck.log 2>&1
echo "memcheck exit $?" >> $O/r2u_memcheck.log
tail -5 $O/r2u_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -m gpu -q -x \
   -k "vegasflowplus_reproduces or plus_fused" > $O/r2u_racecheck.log 2>&1
echo "racecheck exit $?" >> $O/r2u_racecheck.log
tail -5 $O/r2u_racecheck.log
timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -m gpu -q -x \
   -k "vegasflowplus_reproduces" > $O/r2u_synccheck.log 2>&1
echo "synccheck exit $?" >> $O/r2u_synccheck.log
tail -5 $O/r2u_synccheck.log
