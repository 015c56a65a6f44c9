"""CPU suite: the time-line arithmetic of the resident loop (gl_next_sync_point, host code of libgl_b200.so) against a
restatement of find_next_sync_point_and_drift (run.c:199-218, :276-290), on one rank and - the way gl_run_steps uses it
at N > 1 - on TimeBinCount[] rows all-gathered between two processes (gloo): both ranks must arrive at the sync point the
reference gets from MPI_Allreduce(MPI_MIN) of the ranks' own next kick times."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

from gadget_leicester_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TIMEBINS, TIMEBASE = 29, 1 << 29


def reference_next_kick(count, ti_current):
    """run.c:199-216 for ONE rank's TimeBinCount[]"""
    nxt = TIMEBASE
    for n in range(TIMEBINS):
        if count[n]:
            if n > 0:
                dt = 1 << n
                t = (ti_current // dt) * dt + dt
            else:
                t = ti_current
            nxt = min(nxt, t)
    return nxt


def reference_active(ti_next):
    """run.c:276-288"""
    return [1] + [1 if ti_next % (1 << n) == 0 else 0 for n in range(1, TIMEBINS)]


def call(rows, rank, ti_current):
    lib = L.load_library()
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    act = np.zeros(32, dtype=np.int32)
    own, glob, tot = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
    t = lib.gl_next_sync_point(L._ptr(rows), rows.shape[0], rank, int(ti_current), L._ptr(act), C.byref(own), C.byref(glob), C.byref(tot))
    return t, act, own.value, glob.value, tot.value


def random_rows(g, world):
    rows = np.zeros((world, 32), dtype=np.int32)
    for r in range(world):
        used = g.choice(np.arange(0, 12), size=g.integers(1, 5), replace=False)
        rows[r, used] = g.integers(1, 1000, size=used.size)
        if g.random() < 0.7:
            rows[r, 0] = 0                                    # bin 0 (particles on a zero step) is rare
        rows[r, 31] = rows[r, :TIMEBINS].sum()
    return rows


def test_one_rank_equals_the_reference_statements():
    g = np.random.Generator(np.random.Philox(key=3))
    for _ in range(300):
        rows = random_rows(g, 1)
        ti = int(g.integers(0, 4096))
        t, act, own, glob, tot = call(rows, 0, ti)
        assert t == reference_next_kick(rows[0], ti)
        assert list(act[:TIMEBINS]) == reference_active(t) and not act[TIMEBINS:].any()
        want = sum(int(rows[0, b]) for b in range(TIMEBINS) if act[b])
        assert own == glob == want and tot == rows[0, 31]


def test_ranks_get_the_allreduce_min_of_their_own_answers():
    g = np.random.Generator(np.random.Philox(key=4))
    for world in (2, 3, 8):
        for _ in range(100):
            rows = random_rows(g, world)
            ti = int(g.integers(0, 4096))
            want = min(reference_next_kick(rows[r], ti) for r in range(world))          # MPI_Allreduce(MPI_MIN), run.c:218
            act_ref = reference_active(want)
            tot_act = 0
            for r in range(world):
                t, act, own, glob, tot = call(rows, r, ti)
                assert t == want and list(act[:TIMEBINS]) == act_ref
                assert own == sum(int(rows[r, b]) for b in range(TIMEBINS) if act[b])
                assert tot == rows[:, 31].sum()
                tot_act += own
            assert tot_act == glob                                                       # sumup_large_ints, run.c:290


def test_end_of_the_time_line_and_empty_set():
    rows = np.zeros((1, 32), dtype=np.int32)
    t, act, own, glob, tot = call(rows, 0, 100)
    assert t == TIMEBASE and own == 0 and glob == 0                                      # nobody anywhere: nothing to kick
    rows[0, 28] = 5
    rows[0, 31] = 5
    t, act, own, glob, tot = call(rows, 0, 0)
    assert t == 1 << 28 and act[28] == 1 and own == 5


WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import test_sync_point as T
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = np.random.Generator(np.random.Philox(key=100 + rank))
ti = 0
bad = 0
for step in range(50):
    mine = torch.from_numpy(T.random_rows(g, 1)[0].copy())
    rows = [torch.zeros(32, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(rows, mine)                         # what gl_run_steps does with the device histograms over NCCL
    rows = torch.stack(rows).numpy()
    t, act, own, glob, tot = T.call(rows, rank, ti)
    # the reference: every rank's own next kick time, MPI_Allreduce(MPI_MIN)
    loc = torch.tensor([T.reference_next_kick(mine.numpy(), ti)], dtype=torch.int64)
    dist.all_reduce(loc, op=dist.ReduceOp.MIN)
    s = torch.tensor([own], dtype=torch.int64)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)            # sumup_large_ints(1, &NumForceUpdate, &GlobNumForceUpdate)
    bad += int(t != int(loc.item())) + int(int(s.item()) != glob)
    ti = t if t < (1 << 29) else 0
print("RANK %d BAD %d" % (rank, bad))
dist.destroy_process_group()
"""


def test_two_processes_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script), ROOT], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "RANK 0 BAD 0" in r.stdout and "RANK 1 BAD 0" in r.stdout, r.stdout[-2000:]
