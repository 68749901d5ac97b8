# round 2, session 2, call 2: parity after the stream re-definition (v = 2 - r) and the VEGAS+
# warp-contiguous ranges; default bench; pipe-overlap microbenchmark
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2e_tests.log 2>&1
echo "pytest exit $?" >> $O/r2e_tests.log
tail -5 $O/r2e_tests.log
timeout 600 python bench.py > $O/r2e_bench.json 2> $O/r2e_bench.err
tail -c 300 $O/r2e_bench.json
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv,noheader > $O/r2_pipes2.txt
timeout 300 scripts/exp/pipes2 >> $O/r2_pipes2.txt 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv,noheader >> $O/r2_pipes2.txt
cat $O/r2_pipes2.txt
