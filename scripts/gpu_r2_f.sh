# round 2, session 2, call 3: product state after the integer-side cleanup + kernel-variant harness
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2f_tests.log 2>&1
echo "pytest exit $?" >> $O/r2f_tests.log
tail -5 $O/r2f_tests.log
timeout 600 python bench.py --no-cpu-baseline > $O/r2f_bench.json 2> $O/r2f_bench.err
tail -c 300 $O/r2f_bench.json
: > $O/r2_k1_r3.txt
for v in base f1 f2 lea exptab exptab_f2 prmt prmt_f2; do
  timeout 120 scripts/exp/k1_r3_$v >> $O/r2_k1_r3.txt 2>&1
done
cat $O/r2_k1_r3.txt
