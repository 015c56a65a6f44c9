"""Initial-conditions reader and snapshot writer on the reference's own P[] / SphP[] arrays.

Python stand-in for the two reference entry points `read_ic(fname)` (read_ic.c:28) and `savepositions(num)` (io.c:33):
it only owns the buffers (aos.GadgetArrays) and the parameter block; the files are read and written by the library
(gl_snapshot_read / gl_snapshot_write in csrc/gl_snapshot.cu, host code, no GPU needed).  Binary formats 1 and 2,
single file, 4-byte values - what the shipped -D list produces - and format 3 (HDF5; the container is written and parsed
by the library itself, csrc/gl_hdf5.h).
"""
import ctypes as C

import numpy as np

from . import aos
from . import lib as L

BUILD_COOLING, BUILD_SFR, BUILD_STELLARAGE, BUILD_METALS, BUILD_BH_SWALLOWGAS = 1, 2, 4, 8, 16
SHIPPED_BUILD = BUILD_COOLING | BUILD_SFR | BUILD_STELLARAGE | BUILD_BH_SWALLOWGAS      # reference Makefile:8-271
GAMMA_MINUS1 = 7. / 5. - 1                                                                # allvars.h:64-65


def io_layout_struct():
    lay = aos.layout()
    g = L.GLIOLayout()
    g.p_ID = int(lay["p_ID"])
    for name, _ in L.GLIOLayout._fields_:
        if name != "p_ID":
            setattr(g, name, int(lay.get("io_" + name, -1)))
    return g


def io_params(format=1, CoolingOn=0, StarformationOn=0, RestartFlag=0, Ti_Current=0, Timebase_interval=0.0, Time=0.0,
              MassTable=(0,) * 6, MinEgySpec=0.0, InitGasTemp=0.0, InitGasU=0.0, BoxSize=0.0, Omega0=0.0, OmegaLambda=0.0,
              HubbleParam=1.0, build=SHIPPED_BUILD):
    """The entries of the reference's parameter file / `All` that reach the files (names as in allvars.h)."""
    p = L.GLIOParams()
    p.format, p.build, p.CoolingOn, p.StarformationOn, p.RestartFlag = format, build, CoolingOn, StarformationOn, RestartFlag
    p.Ti_Current, p.Timebase_interval, p.Time = Ti_Current, Timebase_interval, Time
    for t in range(6):
        p.MassTable[t] = MassTable[t]
    p.MinEgySpec, p.GammaMinus1, p.InitGasTemp, p.InitGasU = MinEgySpec, GAMMA_MINUS1, InitGasTemp, InitGasU
    p.BoxSize, p.Omega0, p.OmegaLambda, p.HubbleParam = BoxSize, Omega0, OmegaLambda, HubbleParam
    return p


class SnapshotError(RuntimeError):
    """Raised where the reference would call endrun()."""


def _opt(a, n):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.shape[0] >= n
    return a, C.c_void_p(a.ctypes.data)


def write_snapshot(fname, arrays, par, nh=None, sfr=None):
    """savepositions(): arrays = aos.GadgetArrays; nh / sfr = the NH and SFR records per gas particle (they come out of
    the cooling / star-formation modules, which are not part of the path); None writes zeros."""
    lib = L.load_library()
    lay, io = aos.layout_struct(), io_layout_struct()
    err = C.create_string_buffer(512)
    _nh, pnh = _opt(nh, arrays.N_gas)
    _sf, psf = _opt(sfr, arrays.N_gas)
    rc = lib.gl_snapshot_write(str(fname).encode(), C.byref(lay), C.byref(io), C.byref(par), L._ptr(arrays.P), L._ptr(arrays.SphP),
                               arrays.NumPart, pnh, psf, err, len(err))
    if rc:
        raise SnapshotError(err.value.decode())


def read_header(fname, format=1):
    """The header record alone (what read_file() sizes the particle storage from, read_ic.c:754-790)."""
    lib = L.load_library()
    hdr = L.GLIOHeader()
    err = C.create_string_buffer(512)
    if lib.gl_snapshot_read_header(str(fname).encode(), format, C.byref(hdr), err, len(err)):
        raise SnapshotError(err.value.decode())
    return hdr


def read_snapshot(fname, par, max_part, max_part_sph=None, pinned=False):
    """read_ic(): returns (arrays, header); par.MassTable is updated like All.MassTable.  max_part / max_part_sph play
    All.MaxPart / All.MaxPartSph."""
    lib = L.load_library()
    lay, io = aos.layout_struct(), io_layout_struct()
    a = aos.GadgetArrays.empty(max_part, max_part_sph, pinned=pinned)
    hdr = L.GLIOHeader()
    n, ng = C.c_longlong(0), C.c_longlong(0)
    err = C.create_string_buffer(512)
    rc = lib.gl_snapshot_read(str(fname).encode(), C.byref(lay), C.byref(io), C.byref(par), L._ptr(a.P), L._ptr(a.SphP),
                              a.MaxPart, a.MaxPartSph, C.byref(n), C.byref(ng), C.byref(hdr), err, len(err))
    if rc:
        raise SnapshotError(err.value.decode())
    a.NumPart, a.N_gas = int(n.value), int(ng.value)
    return a, hdr


def hdf5_read(fname, path):
    """A numeric dataset ("/PartType0/Coordinates") or attribute ("/Header@Time") of an HDF5 file as a float64 array."""
    lib = L.load_library()
    lib.gl_hdf5_read_f64.argtypes = [C.c_char_p, C.c_char_p, C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.c_char_p, C.c_int]
    dims = (C.c_longlong * 4)()
    rank = C.c_int(0)
    err = C.create_string_buffer(512)
    if lib.gl_hdf5_read_f64(str(fname).encode(), path.encode(), None, 0, dims, C.byref(rank), err, len(err)):
        raise SnapshotError(err.value.decode())
    shape = tuple(int(dims[k]) for k in range(rank.value))
    out = np.zeros(int(np.prod(shape)) if shape else 1)
    if lib.gl_hdf5_read_f64(str(fname).encode(), path.encode(), out.ctypes.data, out.size, dims, C.byref(rank), err, len(err)):
        raise SnapshotError(err.value.decode())
    return out.reshape(shape) if shape else out[0]
