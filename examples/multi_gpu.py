"""Multi-GPU: one process per GPU, events sharded over ranks, one fused
reduce + NVLink exchange + refine kernel per iteration.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/multi_gpu.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # repo root

import torch
import torch.distributed as dist

import vegasflow_b200 as vf

if __name__ == "__main__":
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    inst = vf.VegasFlow(20, int(1e9), verbose=(local == 0))
    inst.compile(vf.integrands.symgauss)
    res, err = inst.run_integration(5)
    if local == 0:
        print(f"symgauss d=20, 1e9 events/iter: {res:.6f} +/- {err:.6f}")
    if dist.is_initialized():
        dist.destroy_process_group()
