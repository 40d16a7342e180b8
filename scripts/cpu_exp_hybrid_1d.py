"""CPU experiment (oracle emulation): 1D PG-SGD multi-rank schedules - all-reduce only, hybrid (all-reduce for the first third,
then one shared Hogwild) and peer only - final sampled path stress vs the reference band.  Results: DESIGN.md section 6."""
import sys, json, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc
from odgi_b200.arrays import read_arrays
bands=json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/golden/stress_reference.json')))
for name in ("DRB1-3123","LPA"):
    go = orc.Graph.from_arrays(read_arrays(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), f'tests/golden/{name}.graph.arr.gz')))
    co = orc.default_sort_config(go)
    n_iters = co.iter_max + 1
    U = co.min_term_updates
    b = bands[f"{name}.sort1d"]
    def stress(x): return orc.path_stress_1d(go, x, 1000000, 12345)
    x = orc.sort_init(go); st=np.zeros(4*64,dtype=np.uint64)
    orc.run_range(go, co, 64, co.seed, U, 0, n_iters, 2, X=x, rng_state=st)
    print(name, 'band mean', b['mean'], 'sd', b['sd'], 'single(64 streams)', stress(x), flush=True)
    for ranks in (2, 8):
        for sw in (0, co.iter_max//3, n_iters):   # 0 = peer only (one Hogwild), n_iters = allreduce only
            reps=[orc.sort_init(go) for _ in range(ranks)]; states=[np.zeros(4*32,dtype=np.uint64) for _ in range(ranks)]
            for it in range(min(sw, n_iters)):
                for r in range(ranks):
                    share=U//ranks+(1 if r<U%ranks else 0)
                    orc.run_range(go, co, 32, co.seed+r*32, share, it, it+1, 2, X=reps[r], rng_state=states[r])
                m = sum(reps)/ranks
                for r in range(ranks): reps[r][...]=m
            x=reps[0]
            if sw < n_iters:
                st=np.zeros(4*64,dtype=np.uint64)
                orc.run_range(go, co, 64, co.seed+1000, U, sw, n_iters, 2, X=x, rng_state=st)
            tag = 'peer-only' if sw==0 else ('allreduce-only' if sw>=n_iters else f'hybrid@{sw}')
            print(f'  ranks {ranks} {tag:15s} stress {stress(x):.4f}', flush=True)
