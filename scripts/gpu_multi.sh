# usage: bash scripts/gpu_multi.sh <ngpus> <workloads...>   (run under gpurun --gpus N)
N=${1:-2}; shift; WLS=${@:-c2 c5}
mkdir -p gpurun_out
python -m pytest tests/test_api_gpu.py -k two_gpu -q 2>&1 | tail -5
for w in $WLS; do
  for n in 1 $N; do
    if [ "$n" = "1" ]; then
      python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${w}_n$n.json 2> gpurun_out/scale_${w}_n$n.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --workload $w --steps 20 --warmup 3 > gpurun_out/scale_${w}_n$n.json 2> gpurun_out/scale_${w}_n$n.err
    fi
    python -c "import sys,json; d=json.loads(open('gpurun_out/scale_${w}_n$n.json').read().strip().splitlines()[-1]); print('$w', 'n=%d'%d['n_gpus'], '%.3e ev/s'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kern_ms %.4f'%d['roofline']['kernel_ms'], 'e2e %.3e'%d['e2e']['value'], 'launches', d['gpu_launches'])" || tail -5 gpurun_out/scale_${w}_n$n.err
  done
done
