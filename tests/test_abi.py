"""CPU suite: the C-ABI shared library loads and exports every symbol include/gl_b200.h declares;
host-only entry points (scalar Peano-Hilbert / Morton keys) are checked against the golden keys."""
import ctypes
import os
import re

import numpy as np
import pytest

import golden_util as G
import util
from gadget_leicester_b200 import lib as L


def _declared_functions():
    hdr = open(os.path.join(util.ROOT, "include", "gl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(gl_[a-z0-9_]+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_functions() == sorted(L.EXPORTS)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(L.LIB_PATH), "libgl_b200.so not built (run __graft_entry__.build())"
    dll = ctypes.CDLL(L.LIB_PATH)
    missing = [f for f in _declared_functions() if not hasattr(dll, f)]
    assert not missing, missing


def test_gl_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load_library()
    ctx = ctypes.c_void_p()
    assert lib.gl_create(ctypes.byref(ctx), 0, 1000) != 0  # no CPU fallback
    from gadget_leicester_b200.gadget import ForcePath
    from gadget_leicester_b200 import ics
    with pytest.raises(L.GLError):
        ForcePath(ics.default_params())


@pytest.mark.parametrize("name", ["cube_2k", "disc_3k_sinks", "plummer_2k"])
def test_scalar_keys_match_reference(name):
    """gl_peano_hilbert_key / gl_morton_key (host code, no GPU) vs peano_and_morton_key of the reference."""
    lib = L.load_library()
    ps, par, ref = G.load_golden(name)
    dom = ref["nodes__domain"]
    fac, corner = dom[1], dom[2:5]
    ijk = ((ps["pos"] - corner) * fac).astype(np.int64)  # forcetree.c:181-183
    ph, mo = ref["keys__phkey"], ref["keys__mortonkey"]
    for i in range(0, ps["pos"].shape[0], 7):
        x, y, z = (int(v) for v in ijk[i])
        assert lib.gl_peano_hilbert_key(x, y, z, 21) == int(ph[i])
        assert lib.gl_morton_key(x, y, z, 21) == int(mo[i])


def test_hilbert_key_small_cases():
    lib = L.load_library()
    # 1 bit: the visit order of the base cell
    order = sorted(range(8), key=lambda p: lib.gl_peano_hilbert_key((p >> 2) & 1, (p >> 1) & 1, p & 1, 1))
    assert order == [0b000, 0b010, 0b110, 0b100, 0b101, 0b111, 0b011, 0b001]
    # the curve is continuous: consecutive keys are face neighbours (3 bits per dimension)
    pts = {}
    for x in range(8):
        for y in range(8):
            for z in range(8):
                pts[lib.gl_peano_hilbert_key(x, y, z, 3)] = (x, y, z)
    assert sorted(pts) == list(range(512))
    for k in range(511):
        a, b = pts[k], pts[k + 1]
        assert sum(abs(a[i] - b[i]) for i in range(3)) == 1
