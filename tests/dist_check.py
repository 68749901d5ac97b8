"""Run under torchrun (2+ ranks, NCCL): sharded VegasFlow integration; rank 0 writes the result.
Used by tests/test_api_gpu.py::test_two_gpu_sharding_matches_single_gpu."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vegasflow_b200 as vf  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    inst = vf.VegasFlow(4, 400000, verbose=False)
    inst.set_seed(2718)
    inst.compile(vf.integrands.symgauss)
    res, err = inst.run_integration(4)
    grid = inst.divisions.cpu().numpy()
    # every rank must hold the identical refined grid
    g = inst.divisions.clone()
    dist.broadcast(g, src=0)
    assert torch.equal(g, inst.divisions), "ranks diverged"
    if dist.get_rank() == 0:
        with open(sys.argv[1], "w") as f:
            json.dump({"res": res, "err": err, "grid": grid.tolist(),
                       "exchange": "p2p" if inst._exchange is not None else "nccl"}, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
