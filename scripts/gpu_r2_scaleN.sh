# round 2, session 3: one point of the driver's SCALE protocol on the final kernels: bash scripts/gpu_r2_scaleN.sh N
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) bench.py --gpus $N > gpurun_out/r2s3_scale_n$N.json 2> gpurun_out/r2s3_scale_n$N.err
echo "N=$N exit $?"
python - $N <<'PY'
import json, sys
n=sys.argv[1]
for line in open(f"gpurun_out/r2s3_scale_n{n}.json"):
    if line.startswith('{"metric"'):
        d=json.loads(line)
        print("N=%s sg8 %.4e ms/step %.4f epi %.4f e2e %.4e"%(n,d["value"],d["ms_per_step"],d["roofline"]["epilogue_kernel_ms"],d["e2e"]["value"]))
        for k,v in d["workloads"].items(): print("     %-5s %.4e ms/step %.4f epi %.4f"%(k,v["value"],v["ms_per_step"],v["epilogue_kernel_ms"]))
PY
