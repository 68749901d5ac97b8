# round 2, session 2, call 14 (2 GPUs): ncu on rank 0 of a 2-rank run, single-pass metrics of the
# fused exchange kernel; NVLink metric availability
mkdir -p gpurun_out
O=gpurun_out
ncu --query-metrics 2>/dev/null | grep -i -E "^nvl|nvlink|^pcie|c2c" | head -60 > $O/r2_ncu_nvl_metrics.txt
wc -l $O/r2_ncu_nvl_metrics.txt
head -20 $O/r2_ncu_nvl_metrics.txt
export NCU_METRICS="gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum"
export NCU_OUT=$O/r2_ncu_exchange_rank0.csv
timeout 600 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 \
   scripts/ncu_rank0.sh --gpus 2 --workload c2 --steps 20 --warmup 3 --no-table --no-cpu-baseline > $O/r2q_run1.log 2>&1
echo "run1 exit $?"; tail -3 $O/r2q_run1.log | cut -c1-300
head -5 $O/r2_ncu_exchange_rank0.csv | cut -c1-300
NVL=$(grep -o -E "^nvl[rt]x__bytes[a-z_]*" $O/r2_ncu_nvl_metrics.txt | sort -u | head -4 | sed 's/$/.sum/' | paste -sd, -)
echo "nvlink metrics: $NVL"
if [ -n "$NVL" ]; then
  export NCU_METRICS="gpu__time_duration.sum,$NVL"
  export NCU_OUT=$O/r2_ncu_exchange_rank0_nvl.csv
  timeout 600 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29812 \
     scripts/ncu_rank0.sh --gpus 2 --workload c2 --steps 20 --warmup 3 --no-table --no-cpu-baseline > $O/r2q_run2.log 2>&1
  echo "run2 exit $?"; tail -3 $O/r2q_run2.log | cut -c1-300
  head -8 $O/r2_ncu_exchange_rank0_nvl.csv | cut -c1-300
fi
