# round 2, 2-GPU pass: sharding tests (VegasFlow p2p/nccl, PlainFlow, VegasFlowPlus cubes), bench at N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_api_gpu.py -m gpu -x -q -k "two_gpu" > gpurun_out/r2b_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2b_tests.log
tail -15 gpurun_out/r2b_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
echo "bench exit $?"
tail -c 2500 gpurun_out/r2b_bench_n2.json
tail -5 gpurun_out/r2b_bench_n2.err
