"""CPU suite: snapshot format 3 (HDF5) - SURVEY.md 8(f) #4, io.c:2619-2866 / read_ic.c:648-1060.

The image has no HDF5 library, so the container is written and parsed by the library itself (csrc/gl_hdf5.h).  What pins
it:
  * a file the HDF5 library itself wrote (scipy ships one, a MATLAB v7.3 file) is parsed by the library's reader and by
    an independent pure-Python reader (tests/h5mini.py), and both find the known contents;
  * files written by gl_snapshot_write(format 3) are parsed by the independent reader: group and dataset names, shapes,
    types and attribute values are what io.c asks libhdf5 for, and every dataset holds exactly the values of the
    format-1 file (which is byte-identical to the reference's, tests/test_snapshot.py);
  * gl_snapshot_read(format 3) of such a file fills P[] / SphP[] bit for bit like reading the format-1 file."""
import os
import struct

import numpy as np
import pytest

from gadget_leicester_b200 import aos, snapshot
import h5mini
import io_util as U
from test_snapshot import _write_par, _ours

try:
    import scipy.io.matlab
    LIBHDF5_FILE = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
except ImportError:                                                                     # pragma: no cover
    LIBHDF5_FILE = ""
needs_libfile = pytest.mark.skipif(not os.path.exists(LIBHDF5_FILE), reason="scipy's HDF5 test file is not installed")

NAMES = {"POS ": "Coordinates", "VEL ": "Velocities", "ID  ": "ParticleIDs", "MASS": "Masses", "U   ": "InternalEnergy", "RHO ": "Density",
         "NE  ": "ElectronAbundance", "NH  ": "NeutralHydrogenAbundance", "HSML": "SmoothingLength", "DUST": "DustRadius",
         "SFR ": "StarFormationRate", "AGE ": "StellarFormationTime", "BHMA": "BH_Mass", "BHMD": "BH_Mdot", "ACDM": "AccDisc_Mass",
         "BHMS": "BH_Mseed", "NEWD": "NewDensity", "DPhM": "DeltaPhotonMomentum", "OPhM": "OldPhotonMomentum", "BPhE": "BirthPhotonEnergy",
         "ACMA": "accreted_mass", "ACMO": "accreted_momentum"}          # get_Tab_IO_Label io.c:2038 -> get_dataset_name io.c:2270
GAS, STARS, SINKS, VIRT, ALL = 1, 16, 32, 8, 63
TYPES = {"POS ": ALL, "VEL ": ALL, "ID  ": ALL, "MASS": ALL, "U   ": GAS, "RHO ": GAS, "NE  ": GAS, "NH  ": GAS, "HSML": GAS, "DUST": ALL, "SFR ": GAS,
         "AGE ": STARS, "BHMA": SINKS, "BHMD": SINKS, "ACDM": SINKS, "BHMS": SINKS, "NEWD": VIRT, "DPhM": VIRT, "OPhM": VIRT, "BPhE": VIRT,
         "ACMA": GAS, "ACMO": GAS}                                       # get_particles_in_block io.c:1460


def _format2_records(data):
    """label -> raw bytes of the record, from a format-2 file"""
    out, p = {}, 0
    while p < len(data):
        assert struct.unpack_from("<i", data, p)[0] == 8
        label = data[p + 4:p + 8].decode()
        p += 16
        n = struct.unpack_from("<i", data, p)[0]
        out[label] = data[p + 4:p + 4 + n]
        p += 8 + n
    return out


@needs_libfile
def test_independent_reader_on_a_file_written_by_libhdf5():
    f = h5mini.File(LIBHDF5_FILE)
    assert f.keys() == ["testdouble"]
    d = f["testdouble"]
    assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
    assert np.array_equal(d.read().ravel(), np.linspace(0, 2 * np.pi, 9))
    assert d.attrs["MATLAB_class"] == b"double"


@needs_libfile
def test_library_reader_on_a_file_written_by_libhdf5():
    a = snapshot.hdf5_read(LIBHDF5_FILE, "/testdouble")
    assert a.shape == (9, 1)
    assert np.array_equal(a.ravel(), np.linspace(0, 2 * np.pi, 9))
    with pytest.raises(snapshot.SnapshotError, match="no such object"):
        snapshot.hdf5_read(LIBHDF5_FILE, "/nothing")
    with pytest.raises(snapshot.SnapshotError, match="only integer and floating-point"):
        snapshot.hdf5_read(LIBHDF5_FILE, "/testdouble@MATLAB_class")          # a string attribute


@pytest.mark.parametrize("cooling,sfr", [(1, 1), (0, 0)])
@pytest.mark.parametrize("masstable", [(0,) * 6, (0, 0.5, 0, 0, 0.25, 0)])
def test_written_file_holds_what_io_c_asks_libhdf5_for(tmp_path, cooling, sfr, masstable):
    a = U.particle_arrays()
    par2 = _write_par(2, cooling, sfr, masstable)
    _, f2 = _ours(a, par2, tmp_path, "f2")
    par3 = _write_par(3, cooling, sfr, masstable)
    snapshot.write_snapshot(tmp_path / "snap", a, par3, *U.derived_records(a))
    fn = tmp_path / "snap.hdf5"                                        # io.c:2623 appends the suffix
    assert fn.exists() and not (tmp_path / "snap").exists()
    f = h5mini.File(str(fn))
    rec = _format2_records(f2)
    npart = np.frombuffer(rec["HEAD"][:24], dtype=np.int32)
    assert f.keys() == ["Header"] + ["PartType%d" % t for t in range(6) if npart[t] > 0]
    A = f["Header"].attrs
    assert list(A) == ["NumPart_ThisFile", "NumPart_Total", "NumPart_Total_HighWord", "MassTable", "Time", "Redshift", "BoxSize",
                       "NumFilesPerSnapshot", "Omega0", "OmegaLambda", "HubbleParam", "Flag_Sfr", "Flag_Cooling", "Flag_StellarAge",
                       "Flag_Metals", "Flag_Feedback", "Flag_DoublePrecision"]                    # io.c:2866-2990, in that order
    assert A["NumPart_ThisFile"].dtype == np.dtype("<i4") and A["NumPart_Total"].dtype == np.dtype("<u4")
    assert np.array_equal(A["NumPart_ThisFile"], npart) and np.array_equal(A["NumPart_Total"], npart)
    assert np.array_equal(A["MassTable"], np.frombuffer(rec["HEAD"][24:72], dtype=np.float64))
    assert A["Time"] == 0.09375 and A["Redshift"] == 0 and A["BoxSize"] == 7.0 and A["NumFilesPerSnapshot"] == 1
    assert (A["Omega0"], A["OmegaLambda"], A["HubbleParam"]) == (0.25, 0.75, 0.7)
    assert A["Flag_Cooling"] == 1 and A["Flag_Sfr"] == 1 and A["Flag_Feedback"] == 1 and A["Flag_DoublePrecision"] == 0
    # every record of the binary file = the datasets of the types that carry it, in type order
    seen = set()
    for label, raw in rec.items():
        if label == "HEAD":
            continue
        ncomp = 3 if label in ("POS ", "VEL ") else 1
        parts = []
        for t in range(6):
            if npart[t] == 0 or not (TYPES[label] >> t) & 1 or (label == "MASS" and masstable[t] != 0):
                continue
            d = f["PartType%d" % t][NAMES[label]]
            seen.add(("PartType%d" % t, NAMES[label]))
            assert d.shape == ((npart[t], 3) if ncomp == 3 else (npart[t],)), label
            assert d.dtype == (np.dtype("<u4") if label == "ID  " else np.dtype("<f4")), label
            parts.append(d.read().tobytes())
        assert b"".join(parts) == raw, label
    for t in range(6):
        if npart[t] > 0:
            assert {("PartType%d" % t, k) for k in f["PartType%d" % t].keys()} == {q for q in seen if q[0] == "PartType%d" % t}
    # the library's own reader sees the same file
    assert np.array_equal(snapshot.hdf5_read(fn, "/Header@NumPart_ThisFile"), npart.astype(np.float64))
    pos0 = snapshot.hdf5_read(fn, "/PartType0/Coordinates")
    assert np.array_equal(pos0.astype(np.float32).tobytes(), f["PartType0"]["Coordinates"].read().tobytes())
    assert snapshot.read_header(fn, format=3).npart[:] == list(npart)


@pytest.mark.parametrize("restart", [0, 2])
@pytest.mark.parametrize("masstable", [(0,) * 6, (0, 0.5, 0, 0, 0, 0)])
def test_reading_format_3_equals_reading_format_1(tmp_path, restart, masstable):
    a = U.particle_arrays(n=3001, ngas=1500, seed=9, zero_u=40)
    fn1, _ = _ours(a, _write_par(1, 1, 1, masstable), tmp_path, "f1")
    snapshot.write_snapshot(tmp_path / "f3", a, _write_par(3, 1, 1, masstable), *U.derived_records(a))
    kw = dict(CoolingOn=1, StarformationOn=1, RestartFlag=restart, MinEgySpec=0.3, InitGasTemp=50.0, InitGasU=0.7)
    p1 = snapshot.io_params(format=1, **kw)
    p3 = snapshot.io_params(format=3, **kw)
    b1, h1 = snapshot.read_snapshot(fn1, p1, 4000)
    b3, h3 = snapshot.read_snapshot(tmp_path / "f3.hdf5", p3, 4000)
    assert (b1.NumPart, b1.N_gas) == (b3.NumPart, b3.N_gas) == (3001, 1500)
    assert bytes(h1)[:196] == bytes(h3)[:196]
    assert list(p1.MassTable) == list(p3.MassTable)
    n, ng, lay = b1.NumPart, b1.N_gas, aos.layout()
    assert np.array_equal(b1.P[: n * lay["p_stride"]], b3.P[: n * lay["p_stride"]])
    assert np.array_equal(b1.SphP[: ng * lay["sph_stride"]], b3.SphP[: ng * lay["sph_stride"]])


def test_reader_names_what_it_does_not_handle(tmp_path):
    a = U.particle_arrays()
    snapshot.write_snapshot(tmp_path / "s", a, _write_par(3, 1, 1), *U.derived_records(a))
    data = bytearray((tmp_path / "s.hdf5").read_bytes())
    rd = lambda fn, n=100: snapshot.read_snapshot(fn, snapshot.io_params(format=3, CoolingOn=1, StarformationOn=1, RestartFlag=2), n)
    rd(tmp_path / "s.hdf5")
    with pytest.raises(snapshot.SnapshotError, match="Not enough space"):
        rd(tmp_path / "s.hdf5", n=10)
    # a dataset the reference would fail to open
    bad = bytes(data).replace(b"Velocities\0", b"Velocitiez\0")
    (tmp_path / "noname.hdf5").write_bytes(bad)
    with pytest.raises(snapshot.SnapshotError, match="no dataset /PartType0/Velocities"):
        rd(tmp_path / "noname.hdf5")
    # cut in the middle of the raw data
    (tmp_path / "short.hdf5").write_bytes(bytes(data[: len(data) - 200]))
    with pytest.raises(snapshot.SnapshotError, match="outside the file"):
        rd(tmp_path / "short.hdf5")
    # superblock of a libver='latest' file
    v2 = bytearray(data)
    v2[8] = 2
    (tmp_path / "v2.hdf5").write_bytes(bytes(v2))
    with pytest.raises(snapshot.SnapshotError, match="superblock version"):
        rd(tmp_path / "v2.hdf5")
    with pytest.raises(snapshot.SnapshotError, match="not an HDF5 file"):
        (tmp_path / "junk.hdf5").write_bytes(b"x" * 4096)
        rd(tmp_path / "junk.hdf5")


# ---- through the reference-side glue: the reference's own main program with read_ic() / savepositions() from
# integration/gl_gadget_glue.c (-DGL_SNAPSHOT_IO), All.SnapFormat = All.ICFormat = 3 -------------------------------------
import subprocess

needs_dropin = pytest.mark.skipif(not (os.path.exists(U.REF_IO) and os.path.exists(U.REF_IO_DROPIN)),
                                  reason="oracle/_ref/ref_io_dropin not built (needs /root/reference)")


@needs_dropin
def test_glue_savepositions_format_3(tmp_path):
    a = U.particle_arrays()
    par = _write_par(3, 1, 1, (0, 0.5, 0, 0, 0, 0))
    lay = aos.layout()
    pf, sf = str(tmp_path / "P.img"), str(tmp_path / "SphP.img")
    a.P[: a.NumPart * lay["p_stride"]].tofile(pf)
    a.SphP[: max(a.N_gas, 1) * lay["sph_stride"]].tofile(sf)
    subprocess.run([U.REF_IO_DROPIN, "W", "p=" + pf, "sph=" + sf, "n=%d" % a.NumPart, "ngas=%d" % a.N_gas, "dir=" + str(tmp_path), "base=ref"]
                   + U.common_args(par), check=True, stdout=subprocess.DEVNULL)
    snapshot.write_snapshot(tmp_path / "ours", a, par, *U.derived_records(a))
    assert (tmp_path / "ref_000.hdf5").read_bytes() == (tmp_path / "ours.hdf5").read_bytes()


@needs_dropin
@pytest.mark.parametrize("restart,temp", [(0, 0.0), (0, 80.0), (2, 0.0)])
def test_glue_read_ic_format_3_equals_the_reference_reading_format_1(tmp_path, restart, temp):
    a = U.particle_arrays(zero_u=5)
    mt = (0, 0.5, 0, 0, 0, 0.75)
    fn1, _ = _ours(a, _write_par(1, 1, 1, mt), tmp_path, "snap1")
    snapshot.write_snapshot(tmp_path / "snap3", a, _write_par(3, 1, 1, mt), *U.derived_records(a))
    kw = dict(CoolingOn=1, StarformationOn=1, RestartFlag=restart, MinEgySpec=0.2, InitGasTemp=temp)
    d1, d2 = tmp_path / "ref", tmp_path / "glue"
    d1.mkdir()
    d2.mkdir()
    want, meta1 = U.reference_read(fn1, snapshot.io_params(format=1, **kw), str(d1))                 # the reference's own read_ic()
    got, meta2 = U.reference_read(tmp_path / "snap3", snapshot.io_params(format=3, **kw), str(d2), binary=U.REF_IO_DROPIN)
    assert meta1 == meta2
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k], want[k]), k
