"""CPU suite: the snapshot writer and the initial-conditions reader (gl_snapshot_write / gl_snapshot_read, SURVEY.md
8(f) #4) against the reference's own savepositions() / read_ic() - run here through oracle/_ref/ref_io where that binary
exists (the build container), and against the committed files it wrote (tests/golden/aux/snapshot_io_60.npz) anywhere.
Bar: files byte-identical; every field a reader sets bit-identical."""
import os

import numpy as np
import pytest

from gadget_leicester_b200 import aos, snapshot
import io_util as U

HAVE_REF = os.path.exists(U.REF_IO)
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/ref_io not built (needs /root/reference)")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux", "snapshot_io_60.npz")


def _write_par(fmt, cooling, sfr, masstable=(0,) * 6):
    return snapshot.io_params(format=fmt, CoolingOn=cooling, StarformationOn=sfr, Ti_Current=96, Timebase_interval=1.0 / 1024,
                              Time=0.09375, MassTable=masstable, MinEgySpec=0.3, BoxSize=7.0, Omega0=0.25, OmegaLambda=0.75,
                              HubbleParam=0.7)


def _ours(a, par, tmp_path, name="ours"):
    nh, sfr = U.derived_records(a)
    fn = tmp_path / name
    snapshot.write_snapshot(fn, a, par, nh=nh, sfr=sfr)
    return fn, fn.read_bytes()


@needs_ref
@pytest.mark.parametrize("fmt", [1, 2])
@pytest.mark.parametrize("cooling,sfr", [(0, 0), (1, 1), (1, 0)])
@pytest.mark.parametrize("masstable", [(0,) * 6, (0, 0.5, 0, 0, 0.25, 0)])
def test_writer_matches_savepositions_byte_for_byte(tmp_path, fmt, cooling, sfr, masstable):
    a = U.particle_arrays()
    par = _write_par(fmt, cooling, sfr, masstable)
    ref = U.reference_write(a, par, str(tmp_path))
    _, ours = _ours(a, par, tmp_path)
    assert len(ours) == len(ref)
    assert ours == ref


@needs_ref
def test_writer_without_gas_and_with_one_type(tmp_path):
    a = U.particle_arrays(n=17, ngas=0, seed=3)
    a.pview("p_Type", np.int16)[:] = 1                       # collisionless only: no U/RHO/HSML records at all
    par = _write_par(1, 1, 1)
    assert _ours(a, par, tmp_path)[1] == U.reference_write(a, par, str(tmp_path))


@needs_ref
@pytest.mark.parametrize("fmt", [1, 2])
def test_writer_and_reader_at_a_size_that_spans_many_work_items(tmp_path, fmt):
    # 140 001 particles: several index stretches (65 536) and well over a hundred fused work items (1 024) per type
    a = U.particle_arrays(n=140001, ngas=90001, seed=5, zero_u=100)
    par = _write_par(fmt, 1, 1, (0, 0.5, 0, 0, 0, 0))
    ref = U.reference_write(a, par, str(tmp_path))
    fn, ours = _ours(a, par, tmp_path)
    assert ours == ref
    _compare_read(fn, dict(format=fmt, CoolingOn=1, StarformationOn=1, RestartFlag=2, MinEgySpec=0.35), tmp_path, max_part=150000)
    _compare_read(fn, dict(format=fmt, CoolingOn=1, StarformationOn=1, RestartFlag=0, MinEgySpec=0.35, InitGasTemp=30.0), tmp_path,
                  max_part=150000)


def _compare_read(fn, par_kw, tmp_path, max_part=200):
    par_ref = snapshot.io_params(**par_kw)
    want, meta = U.reference_read(fn, par_ref, str(tmp_path))
    if par_kw.get("InitGasTemp", 0) > 0 and par_kw.get("InitGasU") is None:
        par_kw = dict(par_kw, InitGasU=meta["InitGasU"])          # the glue computes it like read_ic.c:181-192; here: taken over
    par_ours = snapshot.io_params(**par_kw)
    a, hdr = snapshot.read_snapshot(fn, par_ours, max_part)
    assert (a.NumPart, a.N_gas) == (int(meta["NumPart"]), int(meta["N_gas"]))
    got = U.read_fields(a, par_ours.RestartFlag, bool(par_ours.CoolingOn))
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    for t in range(6):
        assert par_ours.MassTable[t] == meta["MassTable%d" % t], t
        assert hdr.npart[t] == int(meta["npart%d" % t])
    assert hdr.time == meta["time"]
    return a


@needs_ref
@pytest.mark.parametrize("fmt", [1, 2])
@pytest.mark.parametrize("restart", [0, 2])
@pytest.mark.parametrize("masstable", [(0,) * 6, (0.125, 0.5, 0, 0, 0, 0)])
def test_reader_matches_read_ic(tmp_path, fmt, restart, masstable):
    a = U.particle_arrays(zero_u=4)
    fn, _ = _ours(a, _write_par(fmt, 1, 1, masstable), tmp_path, "snap")
    _compare_read(fn, dict(format=fmt, CoolingOn=1, StarformationOn=1, RestartFlag=restart, MinEgySpec=0.35), tmp_path)


@needs_ref
def test_reader_initial_gas_temperature_and_floor(tmp_path):
    # gas read with u == 0 gets All.InitGasU when InitGasTemp > 0 (read_ic.c:196-210); everything is floored at MinEgySpec
    a = U.particle_arrays(zero_u=6)
    par_w = _write_par(1, 0, 0)
    par_w.MinEgySpec = 0.0
    a.sview("s_Density", np.float64)[:6] = 1.0
    fn, _ = _ours(a, par_w, tmp_path, "ic")
    temp = 50.0
    par_ref = snapshot.io_params(format=1, RestartFlag=0, InitGasTemp=temp, MinEgySpec=1e-3)
    _, meta = U.reference_read(fn, par_ref, str(tmp_path))
    assert meta["InitGasU"] > 0
    b = _compare_read(fn, dict(format=1, RestartFlag=0, InitGasTemp=temp, InitGasU=meta["InitGasU"], MinEgySpec=1e-3), tmp_path)
    assert np.all(b.sview("s_Entropy", np.float64)[:6] == meta["InitGasU"])


# ---- the reference host with read_ic() / savepositions() taken from integration/gl_gadget_glue.c ----------------
needs_dropin = pytest.mark.skipif(not (HAVE_REF and os.path.exists(U.REF_IO_DROPIN)),
                                  reason="oracle/_ref/ref_io_dropin not built (needs /root/reference)")


@needs_dropin
@pytest.mark.parametrize("fmt", [1, 2])
@pytest.mark.parametrize("cooling,sfr", [(0, 0), (1, 1)])
def test_glue_savepositions_writes_the_reference_file(tmp_path, fmt, cooling, sfr):
    a = U.particle_arrays()
    par = _write_par(fmt, cooling, sfr, (0, 0.5, 0, 0, 0, 0))
    d1, d2 = tmp_path / "ref", tmp_path / "glue"
    d1.mkdir()
    d2.mkdir()
    assert U.reference_write(a, par, str(d2), binary=U.REF_IO_DROPIN) == U.reference_write(a, par, str(d1))


@needs_dropin
@pytest.mark.parametrize("fmt", [1, 2])
@pytest.mark.parametrize("restart,temp", [(0, 0.0), (0, 80.0), (2, 0.0)])
def test_glue_read_ic_fills_the_reference_arrays(tmp_path, fmt, restart, temp):
    a = U.particle_arrays(zero_u=5)
    fn, _ = _ours(a, _write_par(fmt, 1, 1, (0, 0.5, 0, 0, 0, 0.75)), tmp_path, "snap")
    par = snapshot.io_params(format=fmt, CoolingOn=1, StarformationOn=1, RestartFlag=restart, MinEgySpec=0.2, InitGasTemp=temp)
    d1, d2 = tmp_path / "ref", tmp_path / "glue"
    d1.mkdir()
    d2.mkdir()
    want, meta1 = U.reference_read(fn, par, str(d1))
    got, meta2 = U.reference_read(fn, par, str(d2), binary=U.REF_IO_DROPIN)
    assert meta1 == meta2                                     # NumPart, N_gas, totals, MassTable, InitGasU, header
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k


# ---- anywhere: the files the reference wrote, committed ----------------------------------------------------
def _golden_arrays(g):
    lay = aos.layout()
    n, ng = int(g["n"]), int(g["ngas"])
    a = aos.GadgetArrays.empty(n, ng)
    a.NumPart, a.N_gas = n, ng
    a.P[: n * lay["p_stride"]] = g["P"]
    a.SphP[: ng * lay["sph_stride"]] = g["SphP"]
    return a


@pytest.mark.parametrize("fmt", [1, 2])
def test_writer_reproduces_committed_reference_files(tmp_path, fmt):
    g = np.load(GOLD)
    a = _golden_arrays(g)
    _, ours = _ours(a, _write_par(fmt, 1, 1), tmp_path)
    assert ours == g["file_fmt%d" % fmt].tobytes()


@pytest.mark.parametrize("fmt", [1, 2])
@pytest.mark.parametrize("restart", [0, 2])
def test_reader_reproduces_committed_reference_readback(tmp_path, fmt, restart):
    g = np.load(GOLD)
    fn = tmp_path / "gold"
    fn.write_bytes(g["file_fmt%d" % fmt].tobytes())
    par = snapshot.io_params(format=fmt, CoolingOn=1, StarformationOn=1, RestartFlag=restart, MinEgySpec=0.35)
    a, hdr = snapshot.read_snapshot(fn, par, 100)
    got = U.read_fields(a, restart)
    pre = "read%d_fmt%d_" % (restart, fmt)
    names = [k[len(pre):] for k in g.files if k.startswith(pre) and not k.endswith("MassTable")]
    assert set(names) == set(got)
    for k in names:
        assert np.array_equal(got[k], g[pre + k]), k
    assert np.array_equal(np.array(par.MassTable[:]), g[pre + "MassTable"])


def test_round_trip_is_exact_in_file_precision(tmp_path):
    # write -> read (as a snapshot) -> write again: the second file equals the first except for the records the
    # writer derives (VEL is kicked, U is converted: the reader stores u where the entropy was)
    a = U.particle_arrays()
    par = _write_par(2, 1, 1)
    par.Ti_Current = 0
    a.pview("p_TimeBin", np.int16)[:] = 0
    a.pview("p_Ti_begstep", np.int32)[:] = 0
    fn, first = _ours(a, par, tmp_path, "one")
    rp = snapshot.io_params(format=2, CoolingOn=1, StarformationOn=1, RestartFlag=2)
    b, hdr = snapshot.read_snapshot(fn, rp, 64)
    assert (b.NumPart, b.N_gas) == (a.NumPart, a.N_gas) and list(hdr.npart) == [25, 7, 7, 7, 7, 7]
    order = np.argsort(a.pview("p_Type", np.int16), kind="stable")          # types come back in order (read_ic.c:953-966)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)
    assert np.array_equal(b.pview("p_Type", np.int16), a.pview("p_Type", np.int16)[order])
    assert np.array_equal(b.pview("p_Pos", np.float64, 3), f32(a.pview("p_Pos", np.float64, 3)[order]))
    assert np.array_equal(b.pview("p_Vel", np.float64, 3), f32(a.pview("p_Vel", np.float64, 3)[order]))    # no kick at Ti_Current = 0
    assert np.array_equal(b.pview("p_ID", np.uint32), a.pview("p_ID", np.uint32)[order])
    assert np.array_equal(b.pview("p_Mass", np.float64), f32(a.pview("p_Mass", np.float64)[order]))
    assert np.array_equal(b.pview("p_Hsml", np.float64)[:25], f32(a.pview("p_Hsml", np.float64)[:25]))
    # write what was read: same file, except U (the reader keeps u where the writer expects the entropy)
    par2 = _write_par(2, 1, 1)
    par2.Ti_Current = 0
    b.pview("p_TimeBin", np.int16)[:] = 0
    b.pview("p_Ti_begstep", np.int32)[:] = 0
    b.pview("p_GravAccel", np.float64, 3)[:] = 0
    b.sview("s_HydroAccel", np.float64, 3)[:] = 0
    nh, sfr = U.derived_records(a)                    # the same derived records as in the first file
    snapshot.write_snapshot(tmp_path / "two", b, par2, nh=nh, sfr=sfr)
    second = (tmp_path / "two").read_bytes()
    assert len(second) == len(first)
    k = first.index(b"U   ")
    u0, u1 = k - 4, k + 12 + 4 + 25 * 4 + 4
    assert second[:u0] == first[:u0] and second[u1:] == first[u1:]


def test_reader_fails_where_the_reference_would_endrun(tmp_path):
    a = U.particle_arrays()
    fn1, data1 = _ours(a, _write_par(1, 1, 1), tmp_path, "f1")
    fn2, data2 = _ours(a, _write_par(2, 1, 1), tmp_path, "f2")
    rd = lambda fn, n=100, **kw: snapshot.read_snapshot(fn, snapshot.io_params(**dict(dict(CoolingOn=1, StarformationOn=1, RestartFlag=2), **kw)), n)
    with pytest.raises(snapshot.SnapshotError, match="can't open"):
        rd(tmp_path / "missing", format=1)
    with pytest.raises(snapshot.SnapshotError, match="incorrect header format"):          # endrun(890)
        rd(fn2, format=1)
    with pytest.raises(snapshot.SnapshotError, match="Not enough space"):                 # endrun(173)
        rd(fn1, n=10, format=1)
    with pytest.raises(snapshot.SnapshotError, match="not 1, 2 or 3"):
        rd(fn1, format=4)
    with pytest.raises(snapshot.SnapshotError, match="not an HDF5 file"):
        rd(fn1, format=3)
    # a format-2 record with the wrong label: endrun(1890)
    bad = bytearray(data2)
    k = bad.index(b"VEL ")
    bad[k:k + 4] = b"XVEL"[:4]
    (tmp_path / "badlabel").write_bytes(bytes(bad))
    with pytest.raises(snapshot.SnapshotError, match="incorrect block-structure"):
        rd(tmp_path / "badlabel", format=2)
    # double-precision file: read_ic.c:744-750
    dbl = bytearray(data1)
    dbl[4 + 196: 4 + 200] = (1).to_bytes(4, "little")          # header.flag_doubleprecision
    (tmp_path / "dbl").write_bytes(bytes(dbl))
    with pytest.raises(snapshot.SnapshotError, match="double precision"):
        rd(tmp_path / "dbl", format=1)
    # truncated in the middle of a record; and a file read with the wrong switches (record lengths no longer fit)
    (tmp_path / "short").write_bytes(data1[: len(data1) // 2])
    with pytest.raises(snapshot.SnapshotError):
        rd(tmp_path / "short", format=1)
    with pytest.raises(snapshot.SnapshotError, match="record"):
        rd(fn1, format=1, CoolingOn=0)
