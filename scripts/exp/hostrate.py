import time, torch, sys
sys.path.insert(0, '.')
import __graft_entry__ as g; g.build()
import vegasflow_b200 as vf
for d, n in ((4, 10**6), (8, 10**7)):
    inst = vf.VegasFlow(d, n, verbose=False); inst.set_seed(1); inst.compile(vf.integrands.symgauss)
    inst._run_batched(20); torch.cuda.synchronize()
    for K in (60, 60, 200):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        inst._run_batched(K)
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"d={d} n={n:.0e} K={K}: host enqueue {1e6*(t1-t0)/K:.2f} us/step, total {1e6*(t2-t0)/K:.2f} us/step")
