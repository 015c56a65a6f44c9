import sys, time, os
import numpy as np
sys.path.insert(0, '.')
from gadget_leicester_b200 import ics
from gadget_leicester_b200.gadget import ForcePath
kind, N = sys.argv[1], int(sys.argv[2])
ps = getattr(ics, kind)(N); par = ics.default_params()
ps['hsml'] = ics.guess_hsml(ps, par['DesNumNgb']); ps['hsml'][ps['type'] == 5] = par['Softening'][5]
fp = ForcePath(par); fp.load(ps)
fp.gravity_tree_nocopy(); fp.gravity_tree_nocopy()
d = fp.density(); print('cold', fp.stats()['ms_density'], fp.stats()['density_passes_max'])
for i in range(3):
    d2 = fp.density(); print('steady', fp.stats()['ms_density'], fp.stats()['density_passes_max'])
des = np.where(ps["type"] == 5, 160, 40)
out = np.abs(d2['NumNgb'] - des) > 0.5
sph = (ps['type'] == 0) | (ps['type'] == 5)
idx = np.where(out & sph)[0]
print('outside window:', len(idx))
for i in idx[:20]:
    print(i, ps['type'][i], ps['pos'][i], 'h', d2['Hsml'][i], 'N', d2['NumNgb'][i], 'pairs', d2['DensPairs'][i])
print('sink h', d2['Hsml'][ps['type'] == 5], d2['NumNgb'][ps['type'] == 5], d2['DensPairs'][ps['type'] == 5])
print('h quantiles', np.quantile(d2['Hsml'][ps['type'] == 0], [0.5, 0.99, 0.9999, 1.0]))
fp.force_update_hmax(); fp.hydro_force_nocopy(); print('hydro', fp.stats()['ms_hydro'])
