"""Random small GFAs (reverse-strand steps, 1-step paths, revisited nodes): `pgsgd flatten` (odgi_b200/host/gfa_lite.hpp) against the
reference's own GFA ingest + XP tables (oracle/_ref/ref_driver dump).  Authoring container only."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200
from odgi_b200 import synth
from odgi_b200.arrays import read_arrays
REF = os.path.join(ROOT, 'oracle', '_ref', 'ref_driver')
CLI = os.path.join(ROOT, 'odgi_b200', 'host', 'pgsgd')
tmp=tempfile.mkdtemp(); rng=np.random.default_rng(7); bad=0
for case in range(40):
    N=int(rng.integers(1,60)); P=int(rng.integers(1,9))
    node_len=rng.integers(1,40,size=N).astype(np.uint32)
    counts=rng.integers(1,50,size=P); counts[rng.integers(0,P)]=1          # at least one 1-step path
    first=np.concatenate([[0],np.cumsum(counts)]).astype(np.uint64)
    S=int(first[-1])
    step_node=rng.integers(0,N,size=S).astype(np.uint32)
    # every node must be on some path? not required; but make sure ids compact (all nodes exist as S lines anyway)
    step_rev=(rng.random(S)<0.3).astype(np.uint8)
    g=odgi_b200.FlatGraph(node_len,first,step_node,step_rev)
    gfa=os.path.join(tmp,'g.gfa'); synth.write_gfa(g,gfa)
    a=os.path.join(tmp,'mine.arr'); d=os.path.join(tmp,'ref.arr')
    subprocess.run([CLI,'flatten','-i',gfa,'-o',a],check=True,capture_output=True)
    r=subprocess.run([REF,'dump',gfa,d],capture_output=True,text=True,cwd=tmp)
    if r.returncode!=0: print('ref failed',case,r.stderr[-200:]); bad+=1; continue
    m,rf=read_arrays(a),read_arrays(d)
    ok=all(np.array_equal(m[k],rf[k]) for k in ('node_len','path_first_step','step_node','step_rev','step_pos')) and np.array_equal(m['step_pos'],rf['xp_position_of_step'])
    if not ok: bad+=1; print('MISMATCH',case,N,P,S)
print('gfa fuzz done, mismatches:',bad)
