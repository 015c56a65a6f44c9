import glob
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(HERE, "golden")
OPS = "K T S D R G G D U H N C P"
STAGES = dict(keys="0_K_", tree0="1_T_", hsml0="2_S_", dens1="3_D_", grav1="5_G_", grav2="6_G_", dens2="7_D_",
              hydro="9_H_", nodes="10_N_", pairs="11_C_", pot="12_P_")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    ps = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    par = json.loads(bytes(z["params_json"]).decode())
    ref = {k: z[k] for k in z.files if "__" in k}
    return ps, par, ref
