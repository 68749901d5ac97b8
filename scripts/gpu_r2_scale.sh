# round 2 scaling pass on one 8-GPU box: the driver's SCALE protocol (N = 1, 2, 4, 8 back to back),
# NVLink counters around the 8-GPU run
mkdir -p gpurun_out
nvidia-smi nvlink -gt d -i 0 > gpurun_out/r2_nvlink_before.txt 2>&1
python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29800+N)) bench.py --gpus $N > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err
  echo "N=$N exit $?"
done
nvidia-smi nvlink -gt d -i 0 > gpurun_out/r2_nvlink_after.txt 2>&1
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    for line in open(f"gpurun_out/r2_scale_n{n}.json"):
        if line.startswith('{"metric"'):
            d=json.loads(line)
            if n==1: base=d
            print("N=%d sg8 %.4e (x%.3f) ms/step %.4f epi %.4f e2e %.4e"%(n,d["value"],d["value"]/base["value"],d["ms_per_step"],d["roofline"]["epilogue_kernel_ms"],d["e2e"]["value"]))
            for k,v in d["workloads"].items():
                b=base["workloads"].get(k)
                print("     %-5s %.4e%s ms/step %.4f epi %.4f"%(k,v["value"]," (x%.3f)"%(v["value"]/b["value"]) if b else "",v["ms_per_step"],v["epilogue_kernel_ms"]))
PY
