import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200
from odgi_b200 import capi, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "mid"
g = synth.preset(wl)
X0, Y0 = odgi_b200.layout_init(g, 42)
with odgi_b200.Engine(g) as e:
    e.set_coords_2d(X0, Y0)
    for flags, name in ((0, "st.cg"), (4, "st"), (8, "st.wt"), (12, "atomicExch64"), (1, "red.add.v2")):
        for batch in (1, 4):
            cd = capi.layout_defaults(g, n_streams=148 * 256 * 3 if batch == 1 else 148 * 256 * 2, batch=batch, flags=flags)
            e.run_range(cd, 2, 0, 1)
            st = e.run_range(cd, 2, 1, 4)
            print(f"{name:12s} batch={batch} {st['term_updates']/st['seconds_iterations']/1e9:6.2f} G/s", flush=True)
