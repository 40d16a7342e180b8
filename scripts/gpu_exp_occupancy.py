"""GPU experiment: updates/s of the 2D iteration kernel vs streams / batch / flags on a synthetic graph."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import odgi_b200
from odgi_b200 import capi, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "mid"
g = synth.preset(wl)
X0, Y0 = odgi_b200.layout_init(g, 42)
print(f"L2_FETCH={os.environ.get('PGSGD_L2_FETCH','default32')} workload={wl} N={g.N} S={g.S}", flush=True)
with odgi_b200.Engine(g) as e:
    e.set_coords_2d(X0, Y0)
    for flags in (0, 1):  # 0 = red.add (default), 1 = atom.exch
        for batch in (1, 2, 4):
            for mult in (2, 3, 4, 8):
                ns = 148 * 256 * mult
                cd = capi.layout_defaults(g, n_streams=ns, batch=batch, flags=flags)
                try:
                    e.run_range(cd, 2, 0, 1)
                    st = e.run_range(cd, 2, 1, 4)
                    st2 = e.run_range(cd, 2, 20, 23)
                except Exception as ex:
                    print("fail", flags, batch, mult, ex); continue
                print(f"flags={flags} batch={batch} streams=148*256*{mult}  early {st['term_updates']/st['seconds_iterations']/1e9:6.2f} G/s   cooling {st2['term_updates']/st2['seconds_iterations']/1e9:6.2f} G/s", flush=True)
