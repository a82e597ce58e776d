"""Batch sharding across GPUs and the Experiment-2 learning loop.

Problems of a batch are independent, so the multi-GPU form of the layer is one process per GPU, each
solving a contiguous shard of the batch with **no collective inside setup, solve or backward**
(SURVEY 8e).  The only exchange step in the north-star configs is the gradient all-reduce of the
``Linear(n_features, n_x)`` predictor in the reference's learning experiment
(``experiments/experiment_2.py:52-99``): ``allreduce_grads`` does it with ONE collective over the
flattened gradients (NCCL on GPUs; gloo in the CPU tests).

Semantics note (reference ``solve_box_qp_admm_torch.py:312``): the ADMM stop test is global over the
batch a call sees, so a shard reproduces the reference run on *that shard*; sharded and unsharded runs
differ by at most the stopping tolerance.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_items, rank=None, world_size=None):
    """Contiguous, balanced split: the first ``n_items % world_size`` ranks get one extra item.
    Returns ``(start, stop)``; empty shards are legal (start == stop)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(int(n_items), world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank=None, world_size=None):
    """Slice every tensor of ``tensors`` (``None`` entries pass through) along dim 0 to this rank's shard."""
    first = next(t for t in tensors if t is not None)
    lo, hi = shard_range(first.shape[0], rank, world_size)
    return [None if t is None else t[lo:hi] for t in tensors]


def allreduce_grads(params, group=None, average=False):
    """Sum (or average) the ``.grad`` of ``params`` over all ranks with one all-reduce of the flattened
    gradients.  Parameters without a gradient contribute zeros (so every rank issues the same collective
    even when its shard was empty).  Returns the number of elements reduced."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    _, ws = world()
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    if ws > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= ws
    off = 0
    for p in params:
        k = p.numel()
        g = flat[off:off + k].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += k
    return int(flat.numel())


def qp_cost(z, Q, p):
    """Experiment-2 loss ``sum_b 1/2 z^T Q z + p^T z`` (experiment_2.py:83)."""
    return 0.5 * torch.matmul(torch.matmul(torch.transpose(z, 1, 2), Q), z).sum() + (p * z).sum()


def train_learn_p(qp_layer, Q, p_true, A, b, lb, ub, feats, n_epochs=100, n_mini_batch=32, lr=5e-4, seed=0,
                  device=None, log=None):
    """The reference's learning experiment (experiments/experiment_2.py:52-99), data-parallel.

    A ``Linear(n_features, n_x)`` predicts the linear cost ``p_hat`` from features; every epoch draws a
    mini-batch (with replacement) of the ``n_batch`` stored QPs, solves them with ``qp_layer``
    (``SolveBoxQP``), evaluates the true cost of the decisions, back-propagates through the layer and
    takes an SGD step.  With ``world_size > 1`` the *same* mini-batch indices are drawn on every rank
    (seeded NumPy generator -- the reference leaves NumPy unseeded), rank ``g`` solves the contiguous
    slice ``shard_range(n_mini_batch)`` of them, and the parameter gradients are summed with one
    all-reduce, which reproduces the single-process gradient of the summed loss.

    Returns ``(model, loss_history)`` where ``loss_history[e]`` is the global (all ranks) loss.
    """
    rank, ws = world()
    device = device if device is not None else Q.device
    n_batch, n_x = Q.shape[0], Q.shape[1]
    torch.manual_seed(seed)                        # identical initial weights on every rank
    model = torch.nn.Linear(feats.shape[1], n_x).to(device=device, dtype=Q.dtype)
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    rng = np.random.RandomState(seed)
    hist = []
    for epoch in range(n_epochs):
        idx_all = rng.randint(low=0, high=n_batch, size=n_mini_batch)
        lo, hi = shard_range(n_mini_batch, rank, ws)
        idx = torch.as_tensor(idx_all[lo:hi], dtype=torch.long, device=device)
        opt.zero_grad()
        loss = torch.zeros((), dtype=Q.dtype, device=device)
        if hi > lo:
            p_hat = model(feats[idx]).unsqueeze(2)
            Qi = Q[idx]                                # one gather of the mini-batch's matrices (32 MB at dz = 500) for both uses
            z = qp_layer(Qi, p_hat, A[idx], b[idx], lb[idx], ub[idx])
            loss = qp_cost(z, Qi, p_true[idx])
            loss.backward()
        allreduce_grads(model.parameters())
        opt.step()
        total = loss.detach().clone()
        if ws > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM)
        hist.append(float(total))
        if log is not None and rank == 0:
            log(f"epoch {epoch}, loss {hist[-1]}")
    return model, hist
