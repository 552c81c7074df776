"""The interval schedule of the ensemble kernel (csrc/nb_ensemble.cu, Worker / Piece), checked on the CPU through
nb_ensemble_worker_plan: every step of every system is run exactly once and in order, a system is shared by at most two
neighbouring workers, the head of a shared system is its first worker's FIRST piece and the tail its second worker's
LAST piece, and the work of the workers differs by at most one step."""
import ctypes

import numpy as np
import pytest


def plan(lib, B, n_steps, workers):
    out = []
    buf = (ctypes.c_int * (5 * 4096))()
    for w in range(workers):
        n = ctypes.c_int(0)
        assert lib.nb_ensemble_worker_plan(B, n_steps, workers, w, buf, 4096, n) == 0, lib.nb_last_error()
        assert 1 <= n.value <= 4096
        out.append([tuple(buf[5 * i + j] for j in range(5)) for i in range(n.value)])
    return out


@pytest.mark.parametrize("B,n_steps,workers", [(300, 400, 296), (296, 400, 296), (297, 1, 296), (301, 2, 296), (445, 3, 296),
                                               (1200, 400, 296), (150, 400, 148), (149, 7, 148), (5, 0, 5), (300, 0, 296),
                                               (893, 40, 296), (2 * 148 + 7, 12, 296), (1000, 1000, 37)])
def test_interval_schedule_covers_every_step_once(B, n_steps, workers):
    from hpc import _cuda
    lib = _cuda.load_library()
    pieces = plan(lib, B, n_steps, workers)
    n = max(n_steps, 1)
    covered = np.zeros((B, n), dtype=np.int32)
    owners = [[] for _ in range(B)]
    work = []
    for w, plist in enumerate(pieces):
        steps = 0
        for i, (b, k0, k1, wait, publish) in enumerate(plist):
            assert 0 <= b < B and 0 <= k0 <= k1 <= n_steps
            if n_steps > 0:
                covered[b, k0:k1] += 1
                steps += k1 - k0
            else:
                covered[b, 0] += 1
                steps += 1
            owners[b].append((w, k0, k1))
            # a head (parks its state) is the worker's first piece and starts at step 0; a tail (waits) is its last
            # piece and ends at n_steps; whole systems lie in between and neither wait nor park
            if publish:
                assert i == 0 and k0 == 0 and k1 < n_steps and not wait
            if wait:
                assert i == len(plist) - 1 and k0 > 0 and k1 == n_steps and not publish
            if not wait and not publish:
                assert k0 == 0 and k1 == n_steps
        work.append(steps)
    assert (covered == 1).all()
    assert max(work) - min(work) <= 1
    for b, own in enumerate(owners):
        assert len(own) in (1, 2)
        if len(own) == 2:
            (w0, a0, a1), (w1, b0, b1) = sorted(own, key=lambda t: t[1])
            assert w1 == w0 + 1 and a0 == 0 and a1 == b0 and b1 == n_steps     # head on w, tail on w + 1, contiguous


def test_plan_rejects_more_workers_than_systems():
    from hpc import _cuda
    lib = _cuda.load_library()
    buf, n = (ctypes.c_int * 20)(), ctypes.c_int(0)
    assert lib.nb_ensemble_worker_plan(3, 10, 4, 0, buf, 4, n) == 1 and lib.nb_last_error()
