"""e2e breakdown of the drop-in path (pull / compute / push) on the 1.6M disc."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from gadget_leicester_b200 import ics, aos
from gadget_leicester_b200.gadget import ForcePath
par = ics.default_params()
ps = ics.disc(1600000, seed=42)
ps['hsml'] = ics.guess_hsml(ps, 40); ps['hsml'][ps['type'] == 5] = par['Softening'][5]
fp = ForcePath(par); arr = aos.GadgetArrays(ps, pinned=True); drop = aos.GadgetDropIn(fp, arr)
drop.compute_accelerations()
import ctypes as C
from gadget_leicester_b200 import lib as L
lib = fp._lib
for it in range(4):
    fp.set_TreeReconstructFlag(1); t0 = time.perf_counter(); drop.compute_accelerations(); print('pipelined %.1f ms' % ((time.perf_counter()-t0)*1e3))
for it in range(0):
    fp.set_TreeReconstructFlag(1)
    t0 = time.perf_counter(); fp._chk(lib.gl_gadget_pull(fp.ctx, L._ptr(arr.P), L._ptr(arr.SphP), arr.NumPart, arr.N_gas))
    t1 = time.perf_counter(); fp._chk(lib.gl_compute_accelerations(fp.ctx, 0))
    t2 = time.perf_counter(); fp._chk(lib.gl_gadget_push(fp.ctx, L._ptr(arr.P), L._ptr(arr.SphP), 7))
    t3 = time.perf_counter()
    st = fp.stats()
    print('pull %.1f ms (h2d %.1f) compute %.1f ms push %.1f ms (d2h %.1f)' % ((t1-t0)*1e3, st['ms_h2d'], (t2-t1)*1e3, (t3-t2)*1e3, st['ms_d2h']))
