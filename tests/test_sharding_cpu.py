"""Host logic of the multi-GPU path on CPU: shard arithmetic, and a world-size-2 ``gloo`` run of the
Experiment-2 learning loop (lqp_py_b200/sharding.py) checked against the single-process run.

The CUDA layer cannot run here, so the ``qp_layer`` passed to the loop is a stand-in built on the CPU
oracle (tests may use the oracle; the product never does) -- what is under test is the sharding, the
seeded mini-batch draw and the gradient all-reduce, i.e. everything above the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lqp_py_b200 import sharding
from oracle import box_qp_oracle as orc


class _OracleLayer(torch.autograd.Function):
    """SolveBoxQPLayer semantics (reference solve_box_qp_admm_torch.py:21-67) on the CPU oracle."""

    @staticmethod
    def forward(ctx, Q, p, A, b, lb, ub, control):
        sol = orc.solve(Q, p, A, b, lb, ub, control)
        ctx.save_for_backward(sol["x"], sol["u"], sol["lams"], sol["nus"], Q, A, lb, ub)
        ctx.rho = sol["rho"]
        return sol["x"]

    @staticmethod
    def backward(ctx, g):
        x, u, lams, nus, Q, A, lb, ub = ctx.saved_tensors
        return (*orc.grad(g, x, u, lams, nus, Q, A, lb, ub, ctx.rho), None)


def _layer(control):
    return lambda Q, p, A, b, lb, ub: _OracleLayer.apply(Q, p, A, b, lb, ub, control)


def _problem(n_x=12, n_batch=16, n_feat=3, seed=3):
    dt = torch.float64
    Q, p, A, b, lb, ub = orc.make_exp1_data(n_x, n_batch, seed=seed, dtype=dt)
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(n_batch, n_feat, generator=g, dtype=dt)
    beta = torch.randn(n_feat, n_x, generator=g, dtype=dt)
    p_true = torch.matmul(feats, beta).unsqueeze(2)        # experiment_2.py:52-54
    return Q, p_true, A, b, lb, ub, feats


def _train(n_epochs=6, mini=8):
    torch.set_default_dtype(torch.float64)
    control = orc.default_control(eps_abs=1e-10, eps_rel=1e-10)   # tight: shard-local stopping is then invisible
    Q, p_true, A, b, lb, ub, feats = _problem()
    model, hist = sharding.train_learn_p(_layer(control), Q, p_true, A, b, lb, ub, feats, n_epochs=n_epochs,
                                         n_mini_batch=mini, lr=5e-3, seed=1)
    return [p.detach().clone() for p in model.parameters()], hist


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        params, hist = _train()
        # every rank must hold identical weights after the all-reduced steps
        flat = torch.cat([p.reshape(-1) for p in params])
        gathered = [torch.empty_like(flat) for _ in range(world_size)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        if rank == 0:
            torch.save({"params": params, "hist": hist}, out)
    finally:
        dist.destroy_process_group()


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 32, 128, 1000):
        for ws in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def test_shard_batch_passes_none_and_slices():
    Q, p, A = torch.arange(24.).reshape(6, 2, 2), torch.arange(6.).reshape(6, 1, 1), None
    q1, p1, a1 = sharding.shard_batch([Q, p, A], rank=1, world_size=4)
    assert a1 is None and torch.equal(q1, Q[2:4]) and torch.equal(p1, p[2:4])


def test_allreduce_grads_single_process_is_identity():
    lin = torch.nn.Linear(3, 2)
    lin(torch.ones(1, 3)).sum().backward()
    before = [p.grad.clone() for p in lin.parameters()]
    assert sharding.allreduce_grads(lin.parameters()) == 8
    assert all(torch.equal(a, p.grad) for a, p in zip(before, lin.parameters()))


def test_experiment2_loop_world2_gloo_matches_single_process(tmp_path):
    prev = torch.get_default_dtype()
    try:
        ref_params, ref_hist = _train()
    finally:
        torch.set_default_dtype(prev)
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    # summed shard losses and all-reduced gradients reproduce the single-process run
    np.testing.assert_allclose(got["hist"], ref_hist, rtol=1e-7, atol=1e-9)
    for a, r in zip(got["params"], ref_params):
        np.testing.assert_allclose(a.numpy(), r.numpy(), rtol=1e-7, atol=1e-9)
