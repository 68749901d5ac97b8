# round 2, session 3, call 1: matrix elements through typed spinor components, merged quotients
# and the acos-free single-top angles: GPU tests, block-size / FMA variants, bench rows, ncu
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > $O/r2x_tests.log 2>&1
echo "pytest exit $?" >> $O/r2x_tests.log
grep -a "singletop\|drellyan" $O/r2x_tests.log | head -12
tail -3 $O/r2x_tests.log
: > $O/r2x_me_variants.txt
for v in dy768f0 dy768f1 dy1024f0 dy1024f1 st512f1 st640f0 st640f1 st768f1 st1024f1 dy1024f1 dy768f1 st640f1 st768f1; do
  timeout 120 scripts/exp/k1_r3_$v 50000000 >> $O/r2x_me_variants.txt 2>&1
done
cat $O/r2x_me_variants.txt
timeout 600 python bench.py --workload c4st --no-cpu-baseline --no-table > $O/r2x_bench_c4st.json 2> $O/r2x_bench.err
timeout 600 python bench.py --workload c4dy --no-cpu-baseline --no-table > $O/r2x_bench_c4dy.json 2>> $O/r2x_bench.err
python - <<'PY'
import json
for w in ('c4st','c4dy'):
    d=json.loads(open(f'gpurun_out/r2x_bench_{w}.json').read().strip().splitlines()[-1])
    print(w, d['value'], d['ms_per_step'], d['roofline']['frac'])
PY
for v in dy1024f1 st640f1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 2 -c 1 -f \
      -o $O/r2x_prof_$v scripts/exp/k1_r3_$v 20000000 > $O/r2x_ncu_$v.log 2>&1
  ncu -i $O/r2x_prof_$v.ncu-rep --page details > $O/r2x_prof_${v}_details.txt 2>&1
  ncu -i $O/r2x_prof_$v.ncu-rep --page raw --csv > $O/r2x_prof_${v}_raw.csv 2>&1
  rm -f $O/r2x_prof_$v.ncu-rep
done
ls -la $O | tail -15
