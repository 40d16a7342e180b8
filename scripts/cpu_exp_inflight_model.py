"""CPU experiment: the oracle's Hogwild staleness model (orc_run_inflight: waves of K terms that all read before any of them
writes; summed = red.add, last-writer-wins = exch) against the GPU sweeps in profiles/r01_stream_sweep.log, and on the 1D
cases behind the hub safeguard.  K = terms in flight; hub_terms = K * 2 * max node depth / S."""
import json
import os
import sys

import numpy as np  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from odgi_b200.arrays import read_arrays  # noqa: E402
from oracle import oracle as orc  # noqa: E402

bands = json.load(open(os.path.join(ROOT, "tests", "golden", "stress_reference.json")))


def load(name):
    return orc.Graph.from_arrays(read_arrays(os.path.join(ROOT, "tests", "golden", f"{name}.graph.arr.gz")))


# calibration against profiles/r01_stream_sweep.log: DRB1 2D
go=load('DRB1-3123'); co=orc.default_layout_config(go); X0,Y0=orc.layout_init(go,42)
ref=bands['DRB1-3123.layout2d']['mean']
for exch in (True, False):
    for K in (256, 1024, 4096, 16384, 151552):
        xy=orc.XY_to_xy(X0,Y0); orc.run_inflight(go,co,K,2,exch,xy=xy)
        X,Y=orc.xy_to_XY(xy); s=orc.path_stress_2d(go,X,Y,1000000,12345)
        print(f'DRB1 2D {"exch" if exch else "add "} K={K:6d} hub_terms={K*2*12/go.S:6.2f} stress={s:.5g} rel={100*(s/ref-1):+.2f}%', flush=True)
for name, depth in (('LPA',244),('DRB1-3123',12)):
    go=load(name); co=orc.default_sort_config(go); b=bands[f'{name}.sort1d']
    print(name,'band',b['mean'],'+-',b['sd'], 'tolerance', 0.03*b['mean']+2*b['sd'], flush=True)
    for exch in (False, True):
        for K in (256, 768, 1024, 1536, 1792, 2048, 4096):
            x=orc.sort_init(go); orc.run_inflight(go,co,K,1,exch,X=x)
            s=orc.path_stress_1d(go,x,1000000,12345)
            print(f'  1D {"exch" if exch else "add "} K={K:5d} hub_terms={K*2*depth/go.S:6.2f} stress={s:.5g} rel={100*(s/b["mean"]-1):+.2f}%', flush=True)
