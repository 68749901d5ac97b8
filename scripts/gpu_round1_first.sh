mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/tests_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for w in sg8 c1 c5 c3 c4st; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
tail -5 gpurun_out/tests_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_*.json
