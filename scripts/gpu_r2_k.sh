# round 2, session 2, call 8: new matrix-element random-grid tests, tail-kernel phase record
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s -k "matrix_elements_on_a_large" > $O/r2k_me.log 2>&1
grep -E "median|passed|failed" $O/r2k_me.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2k_tests.log 2>&1
echo "pytest exit $?" >> $O/r2k_tests.log
tail -3 $O/r2k_tests.log
timeout 120 scripts/exp/epi_phases > $O/r2_epilogue_phases.txt 2>&1
cat $O/r2_epilogue_phases.txt
