"""Writes tests/golden/aux/snapshot_io_60.npz: P[] / SphP[] images of a 60-particle set with all six types, the
snapshot files the REFERENCE's savepositions() (io.c:33) writes from them in formats 1 and 2 (CoolingOn = StarformationOn
= 1), and what the REFERENCE's read_ic() (read_ic.c:28) reads back from those files with RestartFlag 0 and 2.
Needs oracle/_ref/ref_io (build container: `make -C oracle/refbuild io`).  Run from the repo root:
    python tests/golden/make_golden_io.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gadget_leicester_b200 import aos, snapshot      # noqa: E402
import io_util as U                                  # noqa: E402
from test_snapshot import _write_par                 # noqa: E402


def main():
    lay = aos.layout()
    a = U.particle_arrays(zero_u=3)
    out = dict(n=a.NumPart, ngas=a.N_gas, P=a.P[: a.NumPart * lay["p_stride"]].copy(), SphP=a.SphP[: a.N_gas * lay["sph_stride"]].copy())
    with tempfile.TemporaryDirectory() as td:
        for fmt in (1, 2):
            data = U.reference_write(a, _write_par(fmt, 1, 1), td)
            out["file_fmt%d" % fmt] = np.frombuffer(data, np.uint8)
            fn = os.path.join(td, "g%d" % fmt)
            with open(fn, "wb") as f:
                f.write(data)
            for restart in (0, 2):
                par = snapshot.io_params(format=fmt, CoolingOn=1, StarformationOn=1, RestartFlag=restart, MinEgySpec=0.35)
                fields, meta = U.reference_read(fn, par, td)
                pre = "read%d_fmt%d_" % (restart, fmt)
                for k, v in fields.items():
                    out[pre + k] = v
                out[pre + "MassTable"] = np.array([meta["MassTable%d" % t] for t in range(6)])
    dst = os.path.join(ROOT, "tests", "golden", "aux", "snapshot_io_60.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
