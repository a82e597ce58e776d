"""Host-side logic of the announced-batch pipeline (SolveBoxQP.prefetch / solve_ahead) that needs no GPU: identity keys of
host tensors, the staging policy for pageable tensors and the choice of staging threads."""
import os

import torch

from lqp_py_b200 import solve_box_qp_admm_torch as M


def _batch(n=4, B=2):
    Q = torch.eye(n).repeat(B, 1, 1)
    p = torch.zeros(B, n, 1)
    return [Q, p, torch.ones(B, 1, n), torch.ones(B, 1, 1), -torch.ones(B, n, 1), torch.ones(B, n, 1)]


def test_prefetch_key_follows_identity_and_version():
    ts = _batch()
    k0 = M._prefetch_key(ts)
    # detached views that autograd leaves are made of (bench / tests: t.detach().requires_grad_()) keep the key ...
    assert M._prefetch_key([t.detach().requires_grad_(j < 2) for j, t in enumerate(ts)]) == k0
    # ... a None entry is part of it, an in-place change moves it, a copy has another address
    assert M._prefetch_key([ts[0], ts[1], None, None, ts[4], ts[5]]) != k0
    ts[1].add_(1.0)
    assert M._prefetch_key(ts) != k0
    assert M._prefetch_key([t.clone() for t in ts]) != M._prefetch_key(ts)


def test_announcements_are_refused_without_a_cuda_device_or_for_cuda_free_paths():
    if torch.cuda.is_available():
        return
    assert M.prefetch_inputs(*_batch()) is False
    assert M.SolveBoxQP(control={}).solve_ahead(*_batch()) is False
    assert not M._PREFETCH


def test_staging_policy_and_thread_choice(monkeypatch):
    small, big = torch.zeros(16), torch.zeros(M._STAGE_MIN_BYTES // 4)
    assert not M._needs_staging(None) and not M._needs_staging(small)
    assert M._needs_staging(big)                       # ordinary (pageable) memory of 4 MB and more
    # torch's own intra-op pool when it is wide enough ...
    monkeypatch.setattr(torch, "get_num_threads", lambda: 16)
    assert M._stage_pool() == (None, 1)
    # ... under torchrun (OMP_NUM_THREADS=1) the usable CPUs divided by the ranks of this host, at most 8
    monkeypatch.setattr(torch, "get_num_threads", lambda: 1)
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(range(16)), raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "2")
    pool, k = M._stage_pool()
    assert k == 8 and pool is not None
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert M._stage_pool()[1] == 2
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "16")
    assert M._stage_pool() == (None, 1)
    # the pool copies what it is told to: one chunk cut over its threads
    import ctypes as C
    src, dst = torch.arange(1000, dtype=torch.float32), torch.zeros(1000)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "4")
    pool, k = M._stage_pool()
    step = -(-1000 // k)
    futs = [pool.submit(C.memmove, dst.data_ptr() + c * 4, src.data_ptr() + c * 4, (min(1000, c + step) - c) * 4)
            for c in range(0, 1000, step)]
    for f in futs:
        f.result()
    assert torch.equal(src, dst)
