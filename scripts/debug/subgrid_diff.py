import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import util, test_gpu_subgrid as T
from gadget_leicester_b200 import ics
OPS=T.OPS
par = ics.default_params(BlackHoleFeedbackFactor=2e-3, SinkBoundary=0.03, SMBHmass=10.0, MinEgySpec=1e-6,
                         Softening=[1.0e-5, 0.01, 5.0e-5, 1.0e-7, 0.01, 0.03])
ps = ics.disc_subgrid(8000, 1500, seed=11)
t = ps["type"]; s = np.nonzero(t == 5)[0]
ps["pos"][s[2]] = ps["pos"][s[1]] + [0.004, 0.0, 0.0]
ps["pos"][s[3]] = ps["pos"][s[1]] + [0.0, 0.005, 0.0]
ps["mass"][s[1:]] = [4.0e-3, 4.0e-3, 3.9e-3]
ps["vel"][s[2]] = ps["vel"][s[1]]; ps["vel"][s[3]] = ps["vel"][s[1]]
ps["timebin"][:] = 4; ps["timebin"][s] = [6, 3, 5, 2]
out = util.run_oracle(util.REF_HARNESS, ps, par, OPS)
pre, post, q = T._split(T._stage(out, OPS.split().index("B"), "B"))
n = pre["Pos"].shape[0]
fp = T._forcepath(pre, q, par)
S = dict(Vel=pre["Vel"].copy(), ID=pre["id"].copy(), SwallowID=pre["SwallowID"].copy(), BH_Mass=pre["BH_Mass"].copy(),
         BH_Mdot=pre["BH_Mdot"].copy(), AccDisc_Mass=pre["AccDisc_Mass"].copy(), Dust_Mass=pre["Dust_Mass"].copy(),
         Total_Mass=pre["Total_Mass"].copy(), Injected_BH_Energy=T._pad(pre["Injected_BH_Energy"], n), SinkDensity=pre["SinkDensity"].copy())
mass, counts = fp.blackhole_accretion(S, active=pre["active"])
t = pre["Type"]
bad = np.nonzero(S["SwallowID"] != post["SwallowID"])[0]
print("mismatches", bad.size, "by type", np.unique(t[bad], return_counts=True))
s5 = np.nonzero(t==5)[0]
for b in bad[:20]:
    d = np.linalg.norm(pre["Pos"][s5]-pre["Pos"][b],axis=1)
    print(b, "type", t[b], "mine", S["SwallowID"][b], "ref", post["SwallowID"][b], "dist to sinks", d, "ids", pre["id"][s5], "hsml", pre["Hsml"][s5])
