import sys, torch
sys.path.insert(0, '.')
from gadget_leicester_b200 import ics
from gadget_leicester_b200.gadget import ForcePath
kind, N = sys.argv[1], int(sys.argv[2])
ps = getattr(ics, kind)(N); par = ics.default_params()
if kind == 'plummer': par['Softening'][1] = ics.plummer_softening(N)
ps['hsml'] = ics.guess_hsml(ps, 40) if (ps['type'] == 0).any() else ps['hsml']
fp = ForcePath(par); fp.load(ps)
for i in range(3):
    fp.gravity_tree_nocopy(); st = fp.stats(); print('walk ms %.2f  interactions %d  -> %.3g /s' % (st['ms_gravity_walk'], st['interactions'], st['interactions'] / st['ms_gravity_walk'] * 1e3))
