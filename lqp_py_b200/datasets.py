"""Synthetic QP generators of the reference's experiments (experiments/utils.py:35-61, 64-131).

They are data recipes, not solver code, but every BASELINE config is defined through them and the
RNG call order decides the bits, so they are restated here for drop-in use by benchmark and
training scripts: ``manual_seed -> randn(L) -> randn(p) -> rand(lb) -> rand(ub)`` on the CPU
generator, in the current default dtype (pass ``dtype=`` to override it for the call).
"""
import numpy as np
import torch


def torch_uniform(*size, lower=0, upper=1):
    return torch.rand(*size) * (upper - lower) + lower


def create_qp_data(n_x, n_batch, n_samples, seed=0, requires_grad=True, dtype=None):
    """Experiment-1/2 data: Q = L^T L / n_samples (SPD), p ~ N(0,1), one budget row A = 1, b = 1,
    box -U(1,2) <= x <= U(1,2).  Returns ``(Q, p, A, b, lb, ub, G, h)`` like the reference."""
    prev = torch.get_default_dtype()
    if dtype is not None:
        torch.set_default_dtype(dtype)
    try:
        torch.manual_seed(seed)
        L = torch.randn(n_batch, n_samples, n_x)
        Q = torch.matmul(torch.transpose(L, 1, 2), L) / n_samples
        Q.requires_grad = requires_grad
        p = torch.randn(n_batch, n_x, 1, requires_grad=requires_grad)
        A = torch.ones(n_batch, 1, n_x)
        b = torch.ones(n_batch, 1, 1)
        lb = -torch_uniform(n_batch, n_x, 1, lower=1, upper=2)
        ub = torch_uniform(n_batch, n_x, 1, lower=1, upper=2)
        eye = torch.eye(n_x)
        G = torch.cat((-eye, eye)).unsqueeze(0) * torch.ones(n_batch, 1, 1)
        h = torch.cat((-lb, ub), dim=1)
    finally:
        torch.set_default_dtype(prev)
    return Q, p, A, b, lb, ub, G, h


def generate_random_A(n_x, prob):
    m = round(n_x ** 0.5)
    A = np.zeros((m, n_x))
    for i in range(m):
        row = np.random.normal(size=(1, n_x))
        keep = np.zeros(1)
        while keep.sum() == 0:
            keep = np.random.binomial(1, prob, size=(1, n_x))
        A[i, :] = row * keep
    return A


def generate_hard_qp(n_x, prob, seed):
    np.random.seed(seed)
    M = np.random.normal(size=(n_x, n_x)) * np.random.binomial(1, prob, size=(n_x, n_x))
    Q = M.T @ M + 1e-2 * np.eye(n_x)
    p = np.random.normal(size=(n_x, 1))
    x0 = np.random.normal(size=(n_x, 1))
    s_lb = -np.random.uniform(size=(n_x, 1))
    s_ub = np.random.uniform(size=(n_x, 1))
    A = generate_random_A(n_x=n_x, prob=prob)
    return Q, p, A, A @ x0, x0 + s_lb, x0 + s_ub


def generate_hard_qp_torch(n_x, prob, seeds, dtype=torch.float64):
    parts = [generate_hard_qp(n_x, prob, s) for s in seeds]
    Q, p, A, b, lb, ub = (torch.tensor(np.stack([q[k] for q in parts]), dtype=dtype, requires_grad=True)
                          for k in range(6))
    eye = torch.eye(n_x, dtype=dtype)
    G = torch.cat((-eye, eye)).unsqueeze(0) * torch.ones(len(seeds), 1, 1, dtype=dtype)
    h = torch.cat((-lb, ub), dim=1)
    return Q, p, A, b, lb, ub, G, h
