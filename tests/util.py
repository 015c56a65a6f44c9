"""Shared helpers of the parity tests.  The oracle binaries are the CHECKER, never the product."""
import os
import subprocess
import tempfile

import numpy as np

from oracle import glio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
PORT_BIN = os.path.join(ROOT, "oracle", "gl_oracle")


def have_ref():
    return os.path.exists(REF_HARNESS)


def run_oracle(binary, ps, params, ops, timeout=3600):
    """Run an oracle binary (the compiled reference or the C restatement) and return its GLOU dict."""
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.glin"), os.path.join(td, "out.glou")
        glio.write_glin(fin, ps, params)
        r = subprocess.run([binary, fin, fout, ops], capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError("oracle failed rc=%d\n%s\n%s" % (r.returncode, r.stdout[-2000:], r.stderr[-2000:]))
        return glio.read_glou(fout)


def by_id(out, prefix, name, n=None):
    """Oracle arrays come in the reference's Peano-Hilbert order with the ID alongside (ids are 1..N);
    return them in input order.  Gas-only arrays cover the leading N_gas entries of that order."""
    a = out[prefix + name]
    ids = out[prefix + "id"][: a.shape[0]]
    res = np.zeros((n if n is not None else ids.max(),) + a.shape[1:], dtype=a.dtype)
    res[ids - 1] = a
    return res


def rel_err_vec(a, b):
    """per-particle |a-b| / |b| on vectors (norm over the last axis)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.ndim == 1:
        num, den = np.abs(a - b), np.abs(b)
    else:
        num, den = np.linalg.norm(a - b, axis=-1), np.linalg.norm(b, axis=-1)
    den = np.where(den > 0, den, 1.0)
    return num / den


def summarize(name, err):
    err = np.asarray(err)
    return "%s: median %.3e  p99 %.3e  max %.3e" % (name, np.median(err), np.percentile(err, 99), err.max())
