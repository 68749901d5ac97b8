"""Run under torchrun (2+ ranks, NCCL): sharded integration; rank 0 writes the result.

    dist_check.py <out.json> [vegas|plus|plain|timeout]

`timeout`: rank 1 runs one iteration less than rank 0 -- the peer exchange of the missing
iteration must end in NaN + RuntimeError on rank 0 after VEGASFLOW_B200_EXCHANGE_TIMEOUT_S,
not in a hung GPU and not in a silently wrong result.

Used by tests/test_api_gpu.py::test_two_gpu_sharding_matches_single_gpu."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vegasflow_b200 as vf  # noqa: E402


def make(alg):
    if alg == "plus":
        return vf.VegasFlowPlus(4, 400000, adaptive=True, verbose=False)
    if alg == "plain":
        return vf.PlainFlow(4, 400000, verbose=False)
    return vf.VegasFlow(4, 400000, verbose=False)


def main():
    alg = sys.argv[2] if len(sys.argv) > 2 else "vegas"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if alg == "timeout":
        return timeout_case()
    inst = make(alg)
    inst.set_seed(2718)
    inst.compile(vf.integrands.symgauss)
    res, err = inst.run_integration(4)
    out = {"res": res, "err": err, "exchange": "p2p" if inst._exchange is not None else "nccl",
           "history": [[h[0], h[1]] for h in inst.history]}
    if alg != "plain":
        # every rank must hold the identical refined grid
        g = inst.divisions.clone()
        dist.broadcast(g, src=0)
        assert torch.equal(g, inst.divisions), "ranks diverged (grid)"
        out["grid"] = inst.divisions.cpu().numpy().tolist()
    if alg == "plus":
        n_ev = inst.n_ev.clone()
        dist.broadcast(n_ev, src=0)
        assert torch.equal(n_ev, inst.n_ev), "ranks diverged (n_ev)"
        out["n_ev"] = inst.n_ev.cpu().numpy().tolist()
        out["n_events"] = inst.n_events
        out["events_log"] = inst.events_log
    if dist.get_rank() == 0:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f)
    dist.destroy_process_group()


def timeout_case():
    import time

    rank = dist.get_rank()
    inst = vf.VegasFlow(4, 200000, verbose=False)
    inst.set_seed(1)
    inst.compile(vf.integrands.symgauss)
    inst.run_integration(2)  # both ranks: healthy exchanges
    outcome = "none"
    t0 = time.time()
    if rank == 0:
        try:
            inst.run_integration(1)  # rank 1 never joins this exchange
            outcome = "returned"
        except RuntimeError as exc:
            outcome = "raised: " + str(exc)[:60]
    else:
        time.sleep(float(os.environ.get("VEGASFLOW_B200_EXCHANGE_TIMEOUT_S", "2")) + 3.0)
    dt = time.time() - t0
    torch.cuda.synchronize()
    if rank == 0:
        with open(sys.argv[1], "w") as f:
            json.dump({"outcome": outcome, "seconds": dt,
                       "poisoned": int(inst._exchange.buf[-1].item())}, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
