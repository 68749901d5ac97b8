# Every multi-rank command runs under `timeout`: a hung collective must not hold N GPUs.
# usage: bash scripts/gpu_multi.sh <ngpus> <workloads...>   (run under gpurun --gpus N)
N=${1:-2}; shift; WLS=${@:-c2 c5}
mkdir -p gpurun_out
python -m pytest tests/test_api_gpu.py -k two_gpu -q -x 2>&1 | tail -15
show() { python -c "import sys,json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2', 'n=%d'%d['n_gpus'], '%.3e ev/s'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kern_ms %.4f'%d['roofline']['kernel_ms'], 'e2e %.3e'%d['e2e']['value'], 'launches', d['gpu_launches'], 'epi_ms', d['roofline'].get('epilogue_kernel_ms'), d['config']['collective'][:40])" || tail -5 ${1%.json}.err; }
for w in $WLS; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${w}_n1.json 2> gpurun_out/scale_${w}_n1.err
  show gpurun_out/scale_${w}_n1.json $w
  for n in $(seq 2 $N); do
    case " 2 4 8 " in *" $n "*) ;; *) continue;; esac
    for x in p2p nccl; do
      VEGASFLOW_B200_EXCHANGE=$x timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --workload $w --steps 20 --warmup 3 > gpurun_out/scale_${w}_n${n}_$x.json 2> gpurun_out/scale_${w}_n${n}_$x.err
      show gpurun_out/scale_${w}_n${n}_$x.json "$w/$x"
    done
  done
done
