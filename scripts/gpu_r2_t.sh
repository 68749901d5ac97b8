# round 2, session 2 (2 GPUs): cluster tail in the multi-GPU VEGAS+ path -- tests + N=2 bench
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_api_gpu.py -m gpu -x -q -k "two_gpu" > $O/r2t_tests2.log 2>&1
echo "pytest exit $?" >> $O/r2t_tests2.log
tail -6 $O/r2t_tests2.log
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests -m gpu -x -q > $O/r2t_tests1.log 2>&1
echo "pytest exit $?" >> $O/r2t_tests1.log
tail -3 $O/r2t_tests1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 2 --workload c3 --no-table --no-cpu-baseline > $O/r2t_bench_c3_n2.json 2> $O/r2t_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2t_bench_c3_n2.json') if l.startswith('{"metric"')][-1])
print('c3 N=2', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['epilogue_kernel_ms'])
PY
