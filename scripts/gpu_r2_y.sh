# round 2, session 3, call 2: Drell-Yan on the exactly rescaled mV, single-top block sizes, bench rows
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > $O/r2y_tests.log 2>&1
echo "pytest exit $?" >> $O/r2y_tests.log
grep -a "singletop\|drellyan" $O/r2y_tests.log | head -12
tail -3 $O/r2y_tests.log
: > $O/r2y_me_variants.txt
for v in dy896 dy1024 st768 st896 st1024 dy1024 dy896 st768 st896 st1024; do
  timeout 120 scripts/exp/k1_r3_$v 50000000 >> $O/r2y_me_variants.txt 2>&1
done
cat $O/r2y_me_variants.txt
timeout 600 python bench.py --workload c4st --no-cpu-baseline --no-table > $O/r2y_bench_c4st.json 2> $O/r2y_bench.err
timeout 600 python bench.py --workload c4dy --no-cpu-baseline --no-table > $O/r2y_bench_c4dy.json 2>> $O/r2y_bench.err
tail -5 $O/r2y_bench.err
python - <<'PY'
import json
for w in ('c4st','c4dy'):
    d=json.loads(open(f'gpurun_out/r2y_bench_{w}.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print(w, d['value'], d['ms_per_step'], r['frac'], r['kernel_ms'], r['kernel_share_of_step'], r.get('frac_implemented_chain'), d['e2e']['value'])
PY
