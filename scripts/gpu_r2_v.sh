mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2v_tests.log 2>&1
echo "pytest exit $?" >> $O/r2v_tests.log
tail -3 $O/r2v_tests.log
timeout 600 python bench.py --workload c4st --no-cpu-baseline --no-table > $O/r2v_bench_c4st.json 2> $O/r2v_bench.err
timeout 600 python bench.py --workload c4dy --no-cpu-baseline --no-table > $O/r2v_bench_c4dy.json 2>> $O/r2v_bench.err
python - <<'PY'
import json
for w in ('c4st','c4dy'):
    d=json.loads(open(f'gpurun_out/r2v_bench_{w}.json').read().strip().splitlines()[-1])
    print(w, d['value'], d['ms_per_step'], d['roofline']['frac'])
PY
