# usage: bash scripts/gpu_profile.sh <workload> <tag>
# ncu launch list of the default bench command + full capture of the event kernel
WL=${1:-sg8}; TAG=${2:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${WL}_${TAG}.csv \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_${WL}_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:event_kernel -s 3 -c 1 -f -o gpurun_out/prof_${WL}_${TAG} \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${WL}_${TAG}.log 2>&1
ls -la gpurun_out | tail -5
