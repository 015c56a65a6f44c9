"""Ad-hoc GPU probe: compares the CUDA path against the compiled reference on one seeded set."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
from gadget_leicester_b200 import ics
from gadget_leicester_b200.gadget import ForcePath
import util

kind, N = sys.argv[1], int(sys.argv[2])
ps = getattr(ics, kind)(N)
n = ps['pos'].shape[0]
par = ics.default_params()
if kind == 'plummer':
    par['Softening'][1] = ics.plummer_softening(N)
t = time.time()
ref = util.run_oracle(util.REF_HARNESS, ps, par, "T S D R G G D U H N C")
print('ref wall %.2fs' % (time.time() - t))
ps['hsml'] = util.by_id(ref, '1_S_', 'Hsml', n)
fp = ForcePath(par)
fp.load(ps)
print('extent', fp.domain_findExtent(), 'ref', ref['9_N_domain'])
nn = fp.force_treebuild()
print('nodes', nn, 'ref numnodes', ref['0_T_numnodes'])
ngas = int((ps['type'] == 0).sum())
if ngas:
    d1 = fp.density()
    print('density1 stats', fp.stats())
    for k in ['Hsml', 'NumNgb']:
        print(util.summarize('D1 ' + k, util.rel_err_vec(d1[k], util.by_id(ref, '2_D_', k, n))))
    for k in ['Density', 'DhsmlDensityFactor', 'DivVel', 'CurlVel', 'Pressure']:
        print(util.summarize('D1 ' + k, util.rel_err_vec(d1[k][:ngas], util.by_id(ref, '2_D_', k, ngas))))
fp.set_TreeReconstructFlag(1)
g1 = fp.gravity_tree()
st = fp.stats(); print('grav1', st['interactions'], 'ref', ref['4_G_interactions'], 'ms', st['ms_gravity_walk'], st['ms_keys_sort'], st['ms_treebuild'], st['ms_moments'])
print('cost mismatches', int((g1['GravCost'] != util.by_id(ref, '4_G_', 'GravCost', n)).sum()))
print(util.summarize('G1 acc', util.rel_err_vec(g1['GravAccel'], util.by_id(ref, '4_G_', 'GravAccel', n))))
g2 = fp.gravity_tree()
st = fp.stats(); print('grav2', st['interactions'], 'ref', ref['5_G_interactions'], 'ms', st['ms_gravity_walk'])
print('cost mismatches', int((g2['GravCost'] != util.by_id(ref, '5_G_', 'GravCost', n)).sum()))
print(util.summarize('G2 acc', util.rel_err_vec(g2['GravAccel'], util.by_id(ref, '5_G_', 'GravAccel', n))))
print(util.summarize('G2 oldacc', util.rel_err_vec(g2['OldAcc'], util.by_id(ref, '5_G_', 'OldAcc', n))))
if ngas:
    d2 = fp.density()
    print('density2 passes', fp.stats()['density_passes_max'], 'ms', fp.stats()['ms_density'])
    for k in ['Hsml', 'NumNgb']:
        print(util.summarize('D2 ' + k, util.rel_err_vec(d2[k], util.by_id(ref, '6_D_', k, n))))
    for k in ['Density', 'DhsmlDensityFactor', 'DivVel', 'CurlVel', 'Pressure']:
        print(util.summarize('D2 ' + k, util.rel_err_vec(d2[k][:ngas], util.by_id(ref, '6_D_', k, ngas))))
    print('dens pair mismatches', int((d2['DensPairs'] != util.by_id(ref, '10_C_', 'DensPairs', n)).sum()))
    fp.force_update_hmax()
    h = fp.hydro_force()
    print('hydro ms', fp.stats()['ms_hydro'])
    for k in ['HydroAccel', 'DtEntropy', 'MaxSignalVel']:
        print(util.summarize('H ' + k, util.rel_err_vec(h[k][:ngas], util.by_id(ref, '8_H_', k, ngas))))
    print('hydro pair mismatches', int((h['HydroPairs'] != util.by_id(ref, '10_C_', 'HydroPairs', n)).sum()))
# node table
tab = fp.nodes()
rn, rf = ref['9_N_nodes'], ref['9_N_nodeflags']
multi = (rf & 128) != 0
rn = rn[multi]
key = lambda t: {tuple(r[0:4]): r for r in t}
mine, theirs = key(tab), key(rn)
print('nodes mine', len(mine), 'ref multi', len(theirs), 'common', len(set(mine) & set(theirs)))
bad = 0
for k in set(mine) & set(theirs):
    a, b = mine[k], theirs[k]
    if not (a[4] == b[4] and (a[5:8] == b[5:8]).all() and a[8] == b[8]):
        bad += 1
        if bad <= 5:
            print('  lvl', a[11], 'mass', a[4] - b[4], 's', a[5:8] - b[5:8], 'maxsoft', a[8], b[8], 'hmax', a[9], b[9])
print('node moment mismatches (mass,s,maxsoft bitwise)', bad)
print('ref times: dens %.3f grav1 %.3f grav2 %.3f dens2 %.3f hydro %.3f' % (ref.get('2_D_time', [0])[0], ref['4_G_time'][0], ref['5_G_time'][0], ref.get('6_D_time', [0])[0], ref.get('8_H_time', [0])[0]))
