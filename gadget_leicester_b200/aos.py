"""The reference's array-of-structs particle storage, as host byte buffers.

`struct particle_data` (536 B) and `struct sph_particle_data` (264 B) are laid out as the reference's
allvars.h defines them under the shipped Makefile's -D list; the byte offsets in gadget_layout.json were
printed by offsetof() against that header (oracle/refbuild/layout_probe.c, `make layout`).  This module is
the Python stand-in for the reference host around the drop-in entry points: it owns P[]/SphP[] in host
memory and calls gl_gadget_pull / gl_compute_accelerations / gl_gadget_push exactly as
integration/gl_gadget_glue.c does from C.
"""
import ctypes as C
import json
import os

import numpy as np

from . import lib as L

_LAYOUT = None


def layout():
    global _LAYOUT
    if _LAYOUT is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gadget_layout.json")) as f:
            _LAYOUT = json.load(f)
    return _LAYOUT


def layout_struct():
    lay = layout()
    g = L.GLGadgetLayout()
    for name, _ in L.GLGadgetLayout._fields_:
        setattr(g, name, int(lay[name]))
    return g


def tree_layout_struct():
    """gl_tree_layout for the reference's struct NODE / struct extNODE (offsets probed by oracle/refbuild `make layout`)."""
    lay = layout()
    g = L.GLTreeLayout()
    g.node_stride, g.ext_stride = int(lay["sizeof_NODE"]), int(lay["sizeof_extNODE"])
    for name, _ in L.GLTreeLayout._fields_:
        if name in lay:
            setattr(g, name, int(lay[name]))
    return g


class ReferenceTree:
    """Nodes[] / Extnodes[] / Nextnode[] / Father[] as the reference stores them (allvars.h:1845-1916), filled from
    the GPU tree by gl_export_tree.  Index space: [0, MaxPart) particles, MaxPart + k = cell k, -1 = end."""

    def __init__(self, fp, max_part=None, tree_alloc_factor=2.0, ti_current=0):
        lay = layout()
        n = fp.N
        self.MaxPart = int(max_part or n)
        self.MaxNodes = int(tree_alloc_factor * self.MaxPart) + 16
        self.lay = lay
        self.nodes = np.zeros(self.MaxNodes * lay["sizeof_NODE"], np.uint8)
        self.ext = np.zeros(self.MaxNodes * lay["sizeof_extNODE"], np.uint8)
        self.Nextnode = np.full(self.MaxPart, -2, np.int32)
        self.Father = np.full(self.MaxPart, -2, np.int32)
        tl = tree_layout_struct()
        nn = C.c_longlong(0)
        fp._chk(fp._lib.gl_export_tree(fp.ctx, C.byref(tl), L._ptr(self.nodes), L._ptr(self.ext), L._ptr(self.Nextnode),
                                       L._ptr(self.Father), self.MaxPart, self.MaxNodes, int(ti_current), C.byref(nn)))
        self.Numnodestree = int(nn.value)

    def nfield(self, name, dtype, count=1):
        st, m = self.lay["sizeof_NODE"], self.Numnodestree
        item = np.dtype(dtype).itemsize
        return np.ndarray((m, count) if count > 1 else (m,), dtype=dtype, buffer=self.nodes.data, offset=self.lay[name],
                          strides=(st, item) if count > 1 else (st,))

    def efield(self, name, dtype, count=1):
        st, m = self.lay["sizeof_extNODE"], self.Numnodestree
        item = np.dtype(dtype).itemsize
        return np.ndarray((m, count) if count > 1 else (m,), dtype=dtype, buffer=self.ext.data, offset=self.lay[name],
                          strides=(st, item) if count > 1 else (st,))

    def ngb_treefind_variable(self, pos, types, center, hsml):
        """ngb.c:169-297 on the exported arrays (plain Python, for tests): gas particles within the box/sphere tests of
        the reference's walk, in the reference's visiting order."""
        ln, ctr = self.nfield("n_len", np.float64), self.nfield("n_center", np.float64, 3)
        sib, nxt = self.nfield("n_sibling", np.int32), self.nfield("n_nextnode", np.int32)
        out, no, MP = [], self.MaxPart, self.MaxPart
        while no >= 0:
            if no < MP:
                p = no
                no = int(self.Nextnode[no])
                if types[p] > 0:
                    continue
                d = pos[p] - center
                if (np.abs(d) > hsml).any() or (d * d).sum() > hsml * hsml:
                    continue
                out.append(p)
            else:
                k = no - MP
                no = int(sib[k])
                dist = hsml + 0.5 * ln[k]
                d = np.abs(ctr[k] - center)
                if (d > dist).any():
                    continue
                dist += 0.366025403785 * ln[k]      # FACT1, ngb.c:30
                if (d * d).sum() > dist * dist:
                    continue
                no = int(nxt[k])
        return out


def _host_buffer(nbytes, pinned):
    if pinned:
        import torch
        t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        t.zero_()
        return t.numpy(), t
    return np.zeros(nbytes, dtype=np.uint8), None


class GadgetArrays:
    """P[] and SphP[] byte buffers plus typed strided views of the fields on the force path."""

    def __init__(self, ps, pinned=True):
        lay = layout()
        n = ps["pos"].shape[0]
        ngas = int((ps["type"] == 0).sum())
        self.NumPart, self.N_gas = n, ngas
        self.P, self._pk = _host_buffer(n * lay["p_stride"], pinned)
        self.SphP, self._sk = _host_buffer(max(ngas, 1) * lay["sph_stride"], pinned)
        self.set_from(ps)

    @classmethod
    def empty(cls, max_part, max_part_sph=None, pinned=False):
        """Unfilled P[] / SphP[] with room for max_part / max_part_sph structs (allocate_memory, allocate.c:56-82);
        NumPart = N_gas = 0 until a reader fills them."""
        lay = layout()
        self = cls.__new__(cls)
        self.NumPart = self.N_gas = 0
        self.MaxPart, self.MaxPartSph = int(max_part), int(max_part if max_part_sph is None else max_part_sph)
        self.P, self._pk = _host_buffer(max(self.MaxPart, 1) * lay["p_stride"], pinned)
        self.SphP, self._sk = _host_buffer(max(self.MaxPartSph, 1) * lay["sph_stride"], pinned)
        return self

    def pview(self, field, dtype, count=1):
        lay = layout()
        n = self.NumPart
        item = np.dtype(dtype).itemsize
        return np.ndarray((n, count) if count > 1 else (n,), dtype=dtype, buffer=self.P.data, offset=lay[field],
                          strides=(lay["p_stride"], item) if count > 1 else (lay["p_stride"],))

    def sview(self, field, dtype, count=1):
        lay = layout()
        n = self.N_gas
        item = np.dtype(dtype).itemsize
        return np.ndarray((n, count) if count > 1 else (n,), dtype=dtype, buffer=self.SphP.data, offset=lay[field],
                          strides=(lay["sph_stride"], item) if count > 1 else (lay["sph_stride"],))

    def set_from(self, ps):
        ng = self.N_gas
        self.pview("p_Pos", np.float64, 3)[:] = ps["pos"]
        self.pview("p_Vel", np.float64, 3)[:] = ps["vel"]
        self.pview("p_Mass", np.float64)[:] = ps["mass"]
        self.pview("p_ID", np.uint32)[:] = ps["id"]
        self.pview("p_Hsml", np.float64)[:] = ps["hsml"]
        self.pview("p_OldAcc", np.float64)[:] = ps["oldacc"]
        self.pview("p_GravCost", np.float32)[:] = 1.0            # init.c:252
        self.pview("p_Type", np.int16)[:] = ps["type"]
        self.pview("p_TimeBin", np.int16)[:] = ps["timebin"]
        self.pview("p_Ti_begstep", np.int32)[:] = ps.get("ti_begstep", 0)
        if ng:
            self.sview("s_VelPred", np.float64, 3)[:] = ps["vel"][:ng]
            self.sview("s_Entropy", np.float64)[:] = ps["entropy"][:ng]
            self.sview("s_DtEntropy", np.float64)[:] = ps["dtentropy"][:ng]
            self.sview("s_Density", np.float64)[:] = -1.0

    def results(self):
        out = dict(GravAccel=self.pview("p_GravAccel", np.float64, 3).copy(), OldAcc=self.pview("p_OldAcc", np.float64).copy(),
                   GravCost=self.pview("p_GravCost", np.float32).copy(), Hsml=self.pview("p_Hsml", np.float64).copy(),
                   NumNgb=self.pview("p_NumNgb", np.float64).copy())
        if self.N_gas:
            for name, f, c in (("Density", "s_Density", 1), ("Pressure", "s_Pressure", 1), ("HydroAccel", "s_HydroAccel", 3),
                               ("DtEntropy", "s_DtEntropy", 1), ("MaxSignalVel", "s_MaxSignalVel", 1),
                               ("DhsmlDensityFactor", "s_DhsmlDensityFactor", 1), ("DivVel", "s_DivVel", 1),
                               ("CurlVel", "s_CurlVel", 1)):
                out[name] = self.sview(f, np.float64, c).copy()
        return out


class GadgetDropIn:
    """compute_accelerations() on host P[]/SphP[] through the drop-in C ABI: pack + H2D, force path, D2H + scatter."""

    def __init__(self, force_path, arrays):
        self.fp, self.a = force_path, arrays
        self._lay = layout_struct()
        fp = self.fp
        fp._chk(fp._lib.gl_gadget_bind(fp.ctx, C.byref(self._lay)))

    def bytes_per_step(self):
        n = self.a.NumPart
        # pack A (from P[]): pos, vel 24 + 24, mass, hsml, oldacc 8 each, type, timebin, ti_begstep 4 each;
        # pack B (from SphP[], gas only): velpred 24, entropy, dtentropy 8 each
        h2d = n * 84 + self.a.N_gas * 40
        other = 5 * 8 if n > self.a.N_gas else 0      # sink / dust columns only travel when such targets exist
        d2h = n * ((3 + 1) * 8 + 4) + n * (7 * 8 + other) + n * 5 * 8
        return h2d, d2h

    def compute_accelerations(self, mode=0):
        fp, a = self.fp, self.a
        lib = fp._lib
        fp.N = a.NumPart
        fp._chk(lib.gl_gadget_compute_accelerations(fp.ctx, L._ptr(a.P), L._ptr(a.SphP), a.NumPart, a.N_gas, int(mode)))
