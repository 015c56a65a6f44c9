"""Reader/writer for the two tiny binary containers the oracle binaries speak.

TEST INFRASTRUCTURE (oracle/): imported only by tests/, bench.py's cpu_baseline / reference arm
and __graft_entry__.smoke() - never by the product package.

GLIN  (input):  "GLIN" i32 version, i64 N, i64 N_gas, f64 par[32], then
                pos[N,3] vel[N,3] mass[N] (f64), type[N] (i32), hsml[N] (f64), id[N] (i32),
                timebin[N] (i32), oldacc[N] (f64), entropy[N_gas] dtentropy[N_gas] (f64)
GLOU  (output): "GLOU" then records: 48-byte zero-padded tag, 1-byte kind ('d' f64, 'i' i32,
                'q' i64), i32 ncomp, i64 n, payload n*ncomp elements.
"""
import numpy as np

NPAR = 32
# indices into par[] - keep in sync with oracle/refbuild/ref_harness.c and oracle/gl_oracle.c
P_THETA, P_ERRTOLFORCEACC, P_OPENCRIT, P_DESNUMNGB, P_MAXNGBDEV, P_ARTVISC, P_G, P_SOFT0 = range(8)
P_MINGASHSMLFRAC = P_SOFT0 + 6
P_TIMEBASE_INTERVAL = P_MINGASHSMLFRAC + 1
P_TI_CURRENT = P_TIMEBASE_INTERVAL + 1
P_BHNGBFACTOR = P_TI_CURRENT + 1
P_PARTALLOC = P_BHNGBFACTOR + 1
P_MINEGYSPEC = P_PARTALLOC + 1
P_BHFEEDBACK, P_SMBHMASS, P_INNERBOUNDARY, P_SINKBOUNDARY, P_VISCOUSTIMESCALE, P_MEANWEIGHT = range(P_MINEGYSPEC + 1, P_MINEGYSPEC + 7)


def pack_params(p):
    """p: dict as produced by gadget_leicester_b200.ics.default_params()."""
    par = np.zeros(NPAR, dtype=np.float64)
    par[P_THETA] = p["ErrTolTheta"]
    par[P_ERRTOLFORCEACC] = p["ErrTolForceAcc"]
    par[P_OPENCRIT] = p["TypeOfOpeningCriterion"]
    par[P_DESNUMNGB] = p["DesNumNgb"]
    par[P_MAXNGBDEV] = p["MaxNumNgbDeviation"]
    par[P_ARTVISC] = p["ArtBulkViscConst"]
    par[P_G] = p["G"]
    par[P_SOFT0:P_SOFT0 + 6] = p["Softening"]
    par[P_MINGASHSMLFRAC] = p["MinGasHsmlFractional"]
    par[P_TIMEBASE_INTERVAL] = p["Timebase_interval"]
    par[P_TI_CURRENT] = p["Ti_Current"]
    par[P_BHNGBFACTOR] = p["BlackHoleNgbFactor"]
    par[P_PARTALLOC] = p.get("PartAllocFactor", 1.2)
    par[P_MINEGYSPEC] = p.get("MinEgySpec", 0.0)
    # sub-grid loops (ops B / Y / Z): 0 = the shipped run's value
    par[P_BHFEEDBACK] = p.get("BlackHoleFeedbackFactor", 0.0)
    par[P_SMBHMASS] = p.get("SMBHmass", 0.0)
    par[P_INNERBOUNDARY] = p.get("InnerBoundary", 0.0)
    par[P_SINKBOUNDARY] = p.get("SinkBoundary", 0.0)
    par[P_VISCOUSTIMESCALE] = p.get("ViscousTimescale", 0.0)
    par[P_MEANWEIGHT] = p.get("MeanWeight", 0.0)
    return par


def write_glin(path, ps, params):
    """ps: dict with pos, vel, mass, type, hsml, id, timebin, oldacc, entropy, dtentropy (gas first)."""
    n = ps["pos"].shape[0]
    ngas = int((ps["type"] == 0).sum())
    assert (ps["type"][:ngas] == 0).all(), "gas particles must come first (allvars.h P[]/SphP[] convention)"
    with open(path, "wb") as f:
        f.write(b"GLIN")
        f.write(np.int32(1).tobytes())
        f.write(np.int64(n).tobytes())
        f.write(np.int64(ngas).tobytes())
        f.write(pack_params(params).tobytes())
        f.write(np.ascontiguousarray(ps["pos"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(ps["vel"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(ps["mass"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(ps["type"], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(ps["hsml"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(ps["id"], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(ps["timebin"], dtype=np.int32).tobytes())
        f.write(np.ascontiguousarray(ps["oldacc"], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(ps["entropy"][:ngas], dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(ps["dtentropy"][:ngas], dtype=np.float64).tobytes())


_KINDS = {b"d": np.float64, b"i": np.int32, b"q": np.int64}


def read_glou(path):
    out = {}
    with open(path, "rb") as f:
        if f.read(4) != b"GLOU":
            raise ValueError("not a GLOU file: %s" % path)
        while True:
            tag = f.read(48)
            if len(tag) < 48:
                break
            name = tag.split(b"\0", 1)[0].decode()
            kind = f.read(1)
            ncomp = int(np.frombuffer(f.read(4), dtype=np.int32)[0])
            n = int(np.frombuffer(f.read(8), dtype=np.int64)[0])
            dt = np.dtype(_KINDS[kind])
            a = np.frombuffer(f.read(dt.itemsize * n * ncomp), dtype=dt)
            out[name] = a.reshape(n, ncomp) if ncomp > 1 else a.copy()
    return out
