# round 2, session 2: VEGAS+ tail on a thread-block cluster -- parity + timing
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2s_tests.log 2>&1
echo "pytest exit $?" >> $O/r2s_tests.log
tail -12 $O/r2s_tests.log
timeout 600 python bench.py --workload c3 --no-cpu-baseline --no-table > $O/r2s_bench_c3.json 2> $O/r2s_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_bench_c3.json').read().strip().splitlines()[-1])
print('c3', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['epilogue_kernel_ms'], d['roofline']['frac'])
PY
