# round 2, session 3: the 8-GPU point of the driver's SCALE protocol on the final kernels
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29808 bench.py --gpus 8 > gpurun_out/r2s3_scale_n8.json 2> gpurun_out/r2s3_scale_n8.err
echo "N=8 exit $?"
tail -3 gpurun_out/r2s3_scale_n8.err
python - <<'PY'
import json
for line in open("gpurun_out/r2s3_scale_n8.json"):
    if line.startswith('{"metric"'):
        d=json.loads(line)
        print("N=8 sg8 %.4e ms/step %.4f epi %.4f e2e %.4e"%(d["value"],d["ms_per_step"],d["roofline"]["epilogue_kernel_ms"],d["e2e"]["value"]))
        for k,v in d["workloads"].items(): print("     %-5s %.4e ms/step %.4f epi %.4f"%(k,v["value"],v["ms_per_step"],v["epilogue_kernel_ms"]))
PY
