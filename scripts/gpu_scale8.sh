# Every multi-rank command runs under `timeout`: a hung collective must not hold N GPUs.
# usage: bash scripts/gpu_scale8.sh   (run under gpurun --gpus 8): weak scaling 1/2/4/8 for c2 and c5
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(open('$1').read().strip().splitlines()[-1]); print('$2', 'n=%d'%d['n_gpus'], '%.3e ev/s'%d['value'], 'ms/step %.4f'%d['ms_per_step'], 'kern_ms %.4f'%d['roofline']['kernel_ms'], 'epi', d['roofline'].get('epilogue_kernel_ms'), 'e2e %.3e'%d['e2e']['value'], d['config']['collective'][:30])" || tail -5 ${1%.json}.err; }
for w in c2 c5; do
  python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s8_${w}_n1.json 2> gpurun_out/s8_${w}_n1.err
  show gpurun_out/s8_${w}_n1.json $w
  for n in 2 4 8; do
    timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $n --workload $w --steps 20 --warmup 3 > gpurun_out/s8_${w}_n${n}.json 2> gpurun_out/s8_${w}_n${n}.err
    show gpurun_out/s8_${w}_n${n}.json "$w"
  done
done
VEGASFLOW_B200_EXCHANGE=nccl timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --workload c2 --steps 20 --warmup 3 > gpurun_out/s8_c2_n8_nccl.json 2> gpurun_out/s8_c2_n8_nccl.err
show gpurun_out/s8_c2_n8_nccl.json "c2/nccl"
