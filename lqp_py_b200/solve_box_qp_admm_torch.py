"""Batched, differentiable box-constrained QP layer (ADMM) -- B200-native drop-in.

    x* = argmin_x 1/2 x^T Q x + p^T x   s.t.  A x = b,  lb <= x <= ub

Same module / autograd surface as the reference file of the same name
(``lqp_py/solve_box_qp_admm_torch.py``):

* ``SolveBoxQP(control).forward(Q, p, A, b, lb, ub) -> x (B, n, 1)``            (:7-18)
* ``SolveBoxQPLayer.apply(Q, p, A, b, lb, ub, control)``                        (:21-67)
* ``torch_solve_box_qp(Q, p, A, b, lb, ub, control) -> dict(x,z,u,lams,nus,rho,iter)`` (:108-333)
* ``torch_solve_box_qp_grad(dl_dz, x, u, lams, nus, Q, A, lb, ub, rho) -> 7-tuple``     (:349-432)
* ``torch_solve_box_qp_grad_kkt(dl_dz, x, lams, nus, Q, A, lb, ub) -> 7-tuple``         (:435-462)

but none of the arithmetic happens in torch: the functions flatten the settings into a POD
struct and call the hand-written sm_100a kernels through the C ABI of ``include/lqpb.h``
(scaling + Gauss-Jordan operator setup, the persistent TMA-streamed ADMM kernel, the
masked-inverse backward).  PyTorch only owns device memory and the CUDA stream.  The one exception
is ``unroll=True``: its iterations run in the same kernels (recorded, then swept backwards) and the
O(n^2) part of the scaling adjoint is a kernel too, but the O(n) vector expressions around them (D from
the column norms, the scaled vectors, the rho-update ratio) are torch operators on the CUDA copies so
that autograd reproduces the reference's subgradient conventions (see ``_solve_unrolled``).

Tensors may live on the GPU (zero copy) or on the CPU like in the reference's experiments;
CPU tensors are staged to the current CUDA device and the results are returned on the CPU.
There is no CPU compute path: without a CUDA device (or without ``_lqpb.so``) the calls raise.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from . import _abi
from .utils import get_ncon

_INF = float("inf")
_UNROLL_TAPE_CAP = 256      # iterations a first-pass tape of the unrolled mode can hold (longer solves: solve, then record)


class SolveBoxQP(nn.Module):
    """``nn.Module`` front end (reference :7-18)."""

    def __init__(self, control):
        super().__init__()
        self.control = control

    def forward(self, Q, p, A, b, lb, ub):
        if self.control.get('unroll', False):
            return torch_solve_box_qp(Q=Q, p=p, A=A, b=b, lb=lb, ub=ub, control=self.control)
        return SolveBoxQPLayer.apply(Q, p, A, b, lb, ub, self.control)

    @staticmethod
    def prefetch(Q, p, A, b, lb, ub):
        """Addition to the reference's API (see :func:`prefetch_inputs`): start uploading the NEXT batch of host tensors
        while the GPU and the PCIe downlink are busy with the current one."""
        return prefetch_inputs(Q, p, A, b, lb, ub)

    def solve_ahead(self, Q, p, A, b, lb, ub, requires_grad=True):
        """Addition to the reference's API: announce a batch like :meth:`prefetch` AND let the layer solve it with this
        module's control as soon as its upload has landed (a worker thread and a second stream: the solve of step k + 1
        runs while step k's gradients still travel to the host).  The later ``forward`` on the same, unmodified tensors
        with an unchanged control returns the finished solution; anything else falls back to the prefetched / plain
        path.  ``requires_grad``: also queue the dl_dz-independent part of the backward (as ``forward`` does for leaves
        that require grad)."""
        if self.control.get('unroll', False):
            return prefetch_inputs(Q, p, A, b, lb, ub)
        return prefetch_inputs(Q, p, A, b, lb, ub, control=self.control, requires_grad=requires_grad)


class SolveBoxQPLayer(torch.autograd.Function):
    """ADMM forward solve + implicit (fixed-point) backward (reference :21-67)."""

    @staticmethod
    def forward(ctx, Q, p, A, b, lb, ub, control):
        out_device = p.device
        # gradient buffers and the backward workspace are taken from the allocator NOW, while the GPU still works on
        # whatever preceded this call: the backward then starts launching as soon as autograd reaches it
        ctx.pre = None
        prep = None
        ahead = _take_solved_ahead((Q, p, A, b, lb, ub), control, any(ctx.needs_input_grad[:6])) if _PREFETCH else None
        if ahead is not None:
            sol, ctx.pre = ahead
        elif not p.is_cuda and any(ctx.needs_input_grad[:6]) and _all_on_host((Q, p, A, b, lb, ub)) and torch.cuda.is_available():
            # host callers: only the workspace is needed here (their gradient buffers are pinned host memory)
            L = _abi.lib()
            dev = _cuda_device(p)
            nb_ws = _ws_bytes("backward", _abi.suffix(p.dtype), Q.shape[0], p.shape[1], get_ncon(A, dim=1))
            ctx.pre = dict(ws=torch.empty(nb_ws, dtype=torch.uint8, device=dev), key=None)
            prep = dict(ws=ctx.pre["ws"], kkt=control.get('backward', 'fixed_point') == 'kkt')
        if p.is_cuda and any(ctx.needs_input_grad[:6]):
            ctx.pre = _prealloc_backward(Q, p, A, ctx.needs_input_grad[:6])
            # ... and the part of the backward that does not depend on dl_dz (mask, assembly, block LDL^T of the adjoint
            # system) is queued right behind the solve by the same C call (lqpb_forward_prep_*): the GPU works through
            # it while Python travels from here to .backward()
            prep = dict(ws=ctx.pre["ws"], kkt=control.get('backward', 'fixed_point') == 'kkt')
        if ahead is None:
            sol = _solve_device(Q, p, A, b, lb, ub, control, host_keys=("x",), prep=prep, allow_async=True)
        ctx.prepared = bool(sol.get("_prepared", False))
        # reference :33-38 -- with no finite bound the caller's dict is switched to rho = 0
        if not (sol["_any_lb"] or sol["_any_ub"]):
            control['rho'] = 0
        ctx.rho = sol["rho"]
        ctx.backward_method = control.get('backward', 'fixed_point')
        ctx.any_bounds = (sol["_any_lb"], sol["_any_ub"])
        ctx.out_device = out_device
        ctx.input_devices = tuple(None if t is None else t.device for t in (Q, p, A, b, lb, ub))
        # saved tensors are the CUDA-resident copies: the backward never re-uploads Q
        dv = sol["_dev"]
        ctx.save_for_backward(sol["_x_dev"], sol["_u_dev"], sol["_lams_dev"], sol["_nus_dev"],
                              dv["Q"], dv["A"], dv["lb"], dv["ub"])
        return sol["x"]

    @staticmethod
    def backward(ctx, dl_dz):
        x, u, lams, nus, Q, A, lb, ub = ctx.saved_tensors
        need = ctx.needs_input_grad[:6]
        kkt = ctx.backward_method == 'kkt'       # reference :63-64
        if dl_dz.device.type == "cpu" and _all_on_host_devices(ctx.input_devices):
            # the caller lives on the host: gradients stream back chunk by chunk while the next chunk is differentiated
            pre, ctx.pre = ctx.pre, None
            ws = pre["ws"] if (ctx.prepared and pre is not None) else None
            ctx.prepared = False
            return (*_grad_host(dl_dz, x, u, lams, nus, Q, A, lb, ub, ctx.rho, need, kkt, ctx.any_bounds, prepared_ws=ws), None)
        pre, ctx.pre = ctx.pre, None                 # one use: a second backward through a retained graph allocates afresh
        finish = ctx.prepared and pre is not None    # the workspace holds the factorised adjoint system of THIS solve
        ctx.prepared = False
        if kkt:
            grads = _grad_kkt_device(dl_dz, x, lams, nus, Q, A, lb, ub, need, ctx.any_bounds, pre=pre, finish=finish)
        else:
            grads = _grad_device(dl_dz, x, u, lams, nus, Q, A, lb, ub, ctx.rho, need, pre=pre, finish=finish)
        return (*_to_devices(grads, ctx.input_devices), None)


class BoxQPTH:
    """Stateful holder of one batch of problems (reference :70-105): ``solve()`` runs the forward solver on the
    stored data, keeps the solution dict in ``self.sol`` and returns ``x``; ``update(...)`` replaces stored fields.
    Like the reference's, ``update`` stores ``None`` -- not the new tensor -- when ``lb`` / ``ub`` are passed
    (:99-102), so a caller that relied on that sees the same behaviour.  No autograd (the functional solver)."""

    def __init__(self, Q, p, A, b, lb, ub, control):
        self.Q, self.p, self.A, self.b, self.lb, self.ub = Q, p, A, b, lb, ub
        self.control = control
        self.sol = {}

    def solve(self, warm_start=False):
        """``warm_start=True`` (an addition: the reference always starts from zero, :221-223) starts the ADMM loop from
        the ``z`` / ``u`` of the previous ``solve()`` of this holder -- the natural mode when only ``p`` changed a
        little between two solves, as in a learning loop."""
        z0 = u0 = None
        if warm_start and self.sol.get('z') is not None and self.sol.get('u') is not None:
            z0, u0 = self.sol['z'], self.sol['u']
        sol = torch_solve_box_qp(Q=self.Q, p=self.p, A=self.A, b=self.b, lb=self.lb, ub=self.ub, control=self.control,
                                 z0=z0, u0=u0)
        self.sol = sol
        return sol.get('x')

    def update(self, Q=None, p=None, A=None, b=None, lb=None, ub=None, control=None):
        for name, val in (("Q", Q), ("p", p), ("A", A), ("b", b), ("control", control)):
            if val is not None:
                setattr(self, name, val)
        if lb is not None:
            self.lb = None                                           # :99-100
        if ub is not None:
            self.ub = None                                           # :101-102
        return None


# ------------------------------------------------------------------------------------------
# functional API
# ------------------------------------------------------------------------------------------
STATUS_CONVERGED, STATUS_MAX_ITERS, STATUS_BREAKDOWN = 1, 2, 4      # lqpb_info.status


def torch_solve_box_qp(Q, p, A, b, lb, ub, control, z0=None, u0=None):
    """Forward solve (reference :108-333).  Returns the reference's dict
    ``{"x","z","u","lams","nus","rho","iter"}``; ``rho`` is a ``(B,1,1)`` tensor when it was
    selected automatically or adapted and the caller's scalar otherwise, ``iter`` a Python int.

    Additions (the reference has neither): ``z0`` / ``u0`` -- the ``z`` / ``u`` of an earlier solve --
    warm-start the loop (the reference always starts from zero, :221-223); and the dict also says HOW the solve ended,
    which the reference drops (:235, :331): ``status`` (1 converged, 2 max_iters reached, 4 numerical breakdown: an
    iterate became NaN / inf) and, per problem, ``converged`` (B,) bool = its own stop test at the last check
    (:307-309), ``primal_residual`` / ``dual_residual`` / ``primal_tolerance`` / ``dual_tolerance`` (B,1,1)
    (:286-304).  ``control['validate'] = True`` checks the documented preconditions (finite inputs, symmetric Q) first
    and raises ``ValueError`` instead of returning a silently different answer."""
    if control.get('validate', False):
        _validate_inputs(Q, p, A, b, lb, ub)
    if control.get('unroll', False):
        return _solve_unrolled(Q, p, A, b, lb, ub, control)           # :328-329 -- a bare x, connected to autograd
    sol = _solve_device(Q, p, A, b, lb, ub, control, z0=z0, u0=u0, want_status=True)
    return {k: sol[k] for k in ("x", "z", "u", "lams", "nus", "rho", "iter", "status", "converged", "primal_residual",
                                "dual_residual", "primal_tolerance", "dual_tolerance")}


def _validate_inputs(Q, p, A, b, lb, ub):
    """Preconditions of the path (reference docstring :109-121: Q an SPD tensor): the factorisations read the lower
    triangle of Q and do not pivot, so a non-symmetric Q would give a different answer than the reference's pivoted LU."""
    for name, t in (("Q", Q), ("p", p), ("A", A), ("b", b)):
        if t is not None and not bool(torch.isfinite(t).all()):
            raise ValueError(f"{name} contains non-finite entries")
    if bool(torch.isnan(lb).any()) or bool(torch.isnan(ub).any()):
        raise ValueError("lb / ub contain NaN")
    asym = float((Q - Q.transpose(1, 2)).abs().max())
    scale = float(Q.abs().max())
    if asym > 1e-6 * max(scale, 1e-300) * (1.0 if Q.dtype == torch.float32 else 1e-6):
        raise ValueError(f"Q is not symmetric (max |Q - Q^T| = {asym:.3e}); the solver reads its lower triangle")


def torch_solve_box_qp_grad(dl_dz, x, u, lams, nus, Q, A, lb, ub, rho):
    """Fixed-point backward (reference :349-432).  Returns ``(dQ, dp, dA, db, dlb, dub, None)``."""
    devs = [t.device for t in (Q, x, A if A is not None else x, A if A is not None else x, lb, ub)]
    dv = _stage(dict(dl_dz=dl_dz, x=x, u=u, lams=lams, nus=nus, Q=Q, A=A, lb=lb, ub=ub))
    rho_d = rho.to(dv["x"].device) if torch.is_tensor(rho) else rho
    grads = _grad_device(dv["dl_dz"], dv["x"], dv["u"], dv["lams"], dv["nus"], dv["Q"], dv["A"], dv["lb"], dv["ub"],
                         rho_d, (True,) * 6)
    return (*_to_devices(grads, devs), None)


def torch_solve_box_qp_grad_kkt(dl_dz, x, lams, nus, Q, A, lb, ub):
    """KKT backward (reference :435-462 and helpers :465-584).  Returns ``(dQ, dp, dA, db, dlb, dub, None)``
    with ``dlb`` / ``dub`` = ``None`` when the batch has no finite lower / upper bound (:572-579).  With a
    one-sided or partly infinite box the reference's dense system contains ``-inf`` and every gradient comes out
    NaN; here an infinite bound contributes exactly zero."""
    devs = [t.device for t in (Q, x, A if A is not None else x, A if A is not None else x, lb, ub)]
    dv = _stage(dict(dl_dz=dl_dz, x=x, lams=lams, nus=nus, Q=Q, A=A, lb=lb, ub=ub))
    grads = _grad_kkt_device(dv["dl_dz"], dv["x"], dv["lams"], dv["nus"], dv["Q"], dv["A"], dv["lb"], dv["ub"],
                             (True,) * 6, None)
    return (*_to_devices(grads, devs), None)


# ------------------------------------------------------------------------------------------
# internals
# ------------------------------------------------------------------------------------------
def _cuda_device(ref):
    if ref.is_cuda:
        return ref.device
    if not torch.cuda.is_available():
        raise RuntimeError("lqp_py_b200 runs on a CUDA device (B200, sm_100a) only; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


_PREFETCH = {}            # inputs on their way to the device: key = identity of the caller's host tensors
_PREFETCH_STREAM = {}     # one copy stream per device


def _prefetch_key(tensors):
    return tuple(None if t is None else (t.data_ptr(), tuple(t.shape), t.dtype, t._version) for t in tensors)


def prefetch_inputs(Q, p, A, b, lb, ub, control=None, requires_grad=True):
    """Announce a batch of host tensors that a LATER ``SolveBoxQP.forward`` / ``torch_solve_box_qp`` call will be given
    (the same tensor objects, unmodified in between), so that its host -> device copy can overlap the work in flight.
    The reference's callers hold CPU tensors, so a training step moves ``Q`` up and ``dQ`` down over PCIe; the link is
    full duplex, and a data loader that announces its next batch between ``forward`` and ``backward`` lets the upload of
    step k + 1 run on a copy stream while step k's backward computes and streams its gradients down.

    The copies are STARTED by the next backward of a host caller, right after it has submitted its own ``dl_dz`` upload
    (the H2D copy engine serves its queue in order: a 128 MB prefetch submitted first would hold the backward up for the
    whole transfer -- measured), or by the forward that consumes the announcement if no backward came in between.  That
    forward finds the device copies by the identity of the host tensors (data pointer, shape, dtype, version counter),
    waits for the copy's event on its own stream and takes the device-pointer path; a call on tensors that were never
    announced, or were modified since, uploads as before.  Pinned host memory is what makes the copy asynchronous.
    Returns True when the batch was registered (CPU tensors and a CUDA device), False otherwise.

    With ``control`` (``SolveBoxQP.solve_ahead``) the batch is also SOLVED ahead of the call that asks for it: a worker
    thread runs the forward (and, with ``requires_grad``, the dl_dz-independent part of the backward) on a second
    stream behind the copy's event, so the kernels of step k + 1 run while step k's gradients are still on their way to
    the host.  Announced two batches ahead, the H2D engine, the GPU and the D2H engine each work on a different step
    and the loop runs at the pace of the slowest of the three.  The consuming forward checks the
    identity of the tensors, the control dict (compared by value with a snapshot taken here) and whether gradients
    are wanted; on any mismatch the solution is dropped and the call proceeds as if only the upload had been announced."""
    ts = (Q, p, A, b, lb, ub)
    if not _all_on_host(ts) or not torch.cuda.is_available():
        return False
    for k, t in zip(("Q", "p", "A", "b", "lb", "ub"), ts):
        if t is not None and t.dtype != p.dtype:
            raise TypeError(f"all tensors must share one dtype, got {p.dtype} and {t.dtype} ({k})")
    key = _prefetch_key(ts)
    if key in _PREFETCH:
        return True
    while len(_PREFETCH) >= 4:                       # announcements that were never used must not pile up
        _PREFETCH.pop(next(iter(_PREFETCH)))
    pf = _PREFETCH[key] = dict(dev=_cuda_device(p), tensors=None, event=None, hold=ts)   # `hold` keeps the host buffers alive
    if control is not None and not any(torch.is_tensor(v) for v in control.values()):
        pf["control"] = dict(control)                # snapshot: the consuming forward compares it with its own dict
        pf["requires_grad"] = bool(requires_grad)
    return True


_STAGE_MIN_BYTES = 4 << 20       # pageable host tensors at least this large are uploaded through page-locked staging memory
_STAGE_CHUNK_BYTES = 16 << 20


def _needs_staging(t):
    return t is not None and t.numel() * t.element_size() >= _STAGE_MIN_BYTES and not t.is_pinned()


def _copy_stream(dev):
    st = _PREFETCH_STREAM.get(dev)
    if st is None:
        st = _PREFETCH_STREAM[dev] = torch.cuda.Stream(device=dev)
    return st


_STAGE_POOL = {}


def _stage_pool():
    """Threads for the pageable -> page-locked staging copies, or None when torch's own intra-op pool is wide enough.
    Under torchrun every rank runs with OMP_NUM_THREADS=1: torch's host copy is then single-threaded (128 MB: 13 ms, slower
    than the driver's own pageable path), so the copy is cut over a few Python threads instead (ctypes ``memmove``, no GIL) --
    the usable CPUs divided by the ranks on this host, at most 8."""
    if torch.get_num_threads() >= 4:
        return None, 1
    try:
        cpus = len(os.sched_getaffinity(0))
    except AttributeError:
        cpus = os.cpu_count() or 1
    ranks = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
    k = max(1, min(8, cpus // ranks))
    if k < 2:
        return None, 1
    if k not in _STAGE_POOL:
        from concurrent.futures import ThreadPoolExecutor
        _STAGE_POOL[k] = ThreadPoolExecutor(max_workers=k, thread_name_prefix="lqpb-stage")
    return _STAGE_POOL[k], k


def _copy_up(dst, src, st):
    """``dst`` (device) <- ``src`` (host, contiguous) on stream ``st``.  The reference's callers hold ordinary (pageable) CPU
    tensors; cudaMemcpyAsync moves those through the driver's own bounce buffer at ~11 GB/s (128 MB of Q: 11.2 ms against
    2.4 ms from page-locked memory).  Large pageable sources therefore go through page-locked staging memory here, chunk
    by chunk: a multi-threaded host copy fills chunk c + 1 (~49 GB/s with 16 threads) while the copy engine moves chunk c."""
    if not _needs_staging(src):
        with torch.cuda.stream(st):
            dst.copy_(src, non_blocking=True)
        return
    flat_src, flat_dst = src.reshape(-1), dst.view(-1)
    total = flat_src.numel()
    chunk = max(1, _STAGE_CHUNK_BYTES // src.element_size())
    pin = torch.empty(total, dtype=src.dtype, pin_memory=True)         # (torch's caching host allocator keeps the block
    pool, k = _stage_pool()                                            # alive until the copies that read it are through)
    for a in range(0, total, chunk):
        b = min(total, a + chunk)
        if pool is None:
            pin[a:b].copy_(flat_src[a:b])
        else:
            es, sp, dp = src.element_size(), flat_src.data_ptr(), pin.data_ptr()
            step = -(-(b - a) // k)
            futs = [pool.submit(C.memmove, dp + c * es, sp + c * es, (min(b, c + step) - c) * es) for c in range(a, b, step)]
            for fu in futs:                          # (ctypes calls run without the GIL)
                fu.result()
        with torch.cuda.stream(st):
            flat_dst[a:b].copy_(pin[a:b], non_blocking=True)


def _launch_prefetches():
    """Start the copies of every announced batch that has not been started yet (copy stream, one event per batch).
    Batches that are to be solved ahead, and batches in pageable memory (their staging copies take host time), are
    handed to the worker thread."""
    for pf in _PREFETCH.values():
        if pf.get("started"):
            continue
        pf["started"] = True
        dev, ts = pf["dev"], pf["hold"]
        staged = any(_needs_staging(t) for t in ts)
        with _on_device(dev):
            st = _copy_stream(dev)
            cur = torch.cuda.current_stream(dev)
            names = ("Q", "p", "A", "b", "lb", "ub")
            dt = ts[1].dtype
            dv = {k: (None if t is None else torch.empty(t.shape, dtype=dt, device=dev)) for k, t in zip(names, ts)}
            st.wait_stream(cur)                      # the buffers were allocated in the current stream's order
            pf["tensors"] = dv
            if not staged:
                for k, t in zip(names, ts):
                    if t is not None:
                        _copy_up(dv[k], _host_view(t), st)
                ev = torch.cuda.Event()
                ev.record(st)
                pf["event"] = ev
        if staged or "control" in pf:
            pf["future"] = _ahead_pool().submit(_ahead_task, pf, staged)


def _ahead_task(pf, staged):
    """Worker thread: staged upload of an announced batch in pageable memory and / or its solve."""
    dev = pf["dev"]
    with torch.cuda.device(dev):
        if staged:
            st = _copy_stream(dev)
            for k, t in zip(("Q", "p", "A", "b", "lb", "ub"), pf["hold"]):
                if t is not None:
                    _copy_up(pf["tensors"][k], _host_view(t), st)
            ev = torch.cuda.Event()
            ev.record(st)
            pf["event"] = ev
    return _solve_ahead(pf) if "control" in pf else None


_AHEAD_POOL = []
_AHEAD_STREAM = {}


def _ahead_pool():
    if not _AHEAD_POOL:
        from concurrent.futures import ThreadPoolExecutor
        _AHEAD_POOL.append(ThreadPoolExecutor(max_workers=1, thread_name_prefix="lqpb-solve-ahead"))
    return _AHEAD_POOL[0]


def _solve_ahead(pf):
    """Worker thread: the forward solve of an announced batch (what SolveBoxQPLayer.forward does for a host caller whose
    tensors were prefetched) on the solve-ahead stream of the device, behind the event of the batch's upload."""
    dev, dv, control = pf["dev"], pf["tensors"], pf["control"]
    with torch.cuda.device(dev):
        st = _AHEAD_STREAM.get(dev)
        if st is None:
            st = _AHEAD_STREAM[dev] = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            st.wait_event(pf["event"])
            pre = prep = None
            if pf["requires_grad"]:
                pd = dv["p"]
                nb_ws = _ws_bytes("backward", _abi.suffix(pd.dtype), dv["Q"].shape[0], pd.shape[1], get_ncon(dv["A"], dim=1))
                pre = dict(ws=torch.empty(nb_ws, dtype=torch.uint8, device=dev), key=None)
                prep = dict(ws=pre["ws"], kkt=control.get('backward', 'fixed_point') == 'kkt')
            sol = _solve_device(dv["Q"], dv["p"], dv["A"], dv["b"], dv["lb"], dv["ub"], control, host_keys=(), prep=prep)
            # x goes to the host as SM stores into page-locked memory (lqpb_copy_mapped), not through the D2H copy engine:
            # that engine is in the middle of the previous batch's gradients and would hold this solve up to their end
            xd = sol["_x_dev"]
            hx = torch.empty(xd.shape, dtype=xd.dtype, pin_memory=True)
            _abi.check(_abi.lib().lqpb_copy_mapped(_abi.ptr(hx), _abi.ptr(xd), xd.numel() * xd.element_size(),
                                                   C.c_void_p(_raw_stream(dev))), "lqpb_copy_mapped")
            st.synchronize()
            sol["x"] = hx
    return sol, pre


def _take_solved_ahead(tensors, control, wants_grad):
    """The finished forward of a batch announced through ``SolveBoxQP.solve_ahead`` -- ``(sol, pre)`` as
    ``SolveBoxQPLayer.forward`` would have produced them -- or None (never announced, tensors modified since, another
    control, gradients wanted but not prepared)."""
    key = _prefetch_key(tensors)
    pf = _PREFETCH.get(key)
    if pf is None or "control" not in pf:
        return None
    if not pf.get("started"):
        _launch_prefetches()                         # no backward came in between: copy and solve start now
    sol, pre = pf.pop("future").result()             # (a failed solve raises here, in the caller's thread)
    same = (not any(torch.is_tensor(v) for v in control.values())) and dict(control) == pf.pop("control")
    if not same or (wants_grad and pre is None):
        return None                                  # the entry stays: its device copies serve the prefetched path
    _PREFETCH.pop(key)
    # the solve's tensors were allocated in the solve-ahead stream's pool and are used on the caller's stream from here on
    cur = torch.cuda.current_stream(pf["dev"])
    held = [v for v in sol.values() if torch.is_tensor(v) and v.is_cuda] + [v for v in sol["_dev"].values() if v is not None]
    if pre is not None:
        held.append(pre["ws"])
    for t in held:
        t.record_stream(cur)
    return sol, pre


def _take_prefetched(tensors):
    """The device copies announced for exactly these host tensors, ordered into the current stream; None otherwise."""
    if not _PREFETCH:
        return None
    key = _prefetch_key(tensors)
    if key not in _PREFETCH:
        return None
    if not _PREFETCH[key].get("started"):
        _launch_prefetches()                         # no backward came in between: the copy starts now
    pf = _PREFETCH.pop(key)
    fut = pf.pop("future", None)
    if fut is not None:                              # a staged upload by the worker thread, or a solve-ahead nobody took
        fut.result()                                 # (e.g. torch_solve_box_qp called directly)
    torch.cuda.current_stream(pf["dev"]).wait_event(pf["event"])
    return pf["tensors"]


def _stage(tensors):
    """Move a dict of (optional) tensors to one CUDA device, contiguous, same dtype."""
    ref = next(t for t in tensors.values() if t is not None)
    dev = _cuda_device(next((t for t in tensors.values() if t is not None and t.is_cuda), ref))
    out = {}
    for k, t in tensors.items():
        if t is None:
            out[k] = None
            continue
        if t.dtype != ref.dtype:
            raise TypeError(f"all tensors must share one dtype, got {ref.dtype} and {t.dtype} ({k})")
        t = t.detach()
        if t.device != dev:
            t = t.to(dev, non_blocking=True)
        out[k] = t.contiguous()
    return out


def _to_devices(tensors, devices):
    """Return every tensor of ``tensors`` on the matching device of ``devices`` (``None`` entries pass through).
    Device -> host copies go to pinned staging memory, are all enqueued first and waited for with ONE stream
    synchronisation."""
    out, pending = [], None
    for t, device in zip(tensors, devices):
        if t is None or device is None or t.device == device:
            out.append(t)
        elif device.type == "cpu":
            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host.copy_(t, non_blocking=True)
            pending = t.device
            out.append(host)
        else:
            out.append(t.to(device))
    if pending is not None:
        torch.cuda.current_stream(pending).synchronize()
    return out


def _to_device_of(t, device):
    return _to_devices([t], [device])[0]


_DTYPE_VALUE_CACHE = {}


def _default_dtype_value(v):
    """The reference builds some constants with ``torch.ones(1) * v`` in the *default* dtype
    (:150, :230); reproduce the rounding (cached per default dtype: the tensor round trip costs ~10 us)."""
    key = (torch.get_default_dtype(), v)
    r = _DTYPE_VALUE_CACHE.get(key)
    if r is None:
        r = _DTYPE_VALUE_CACHE[key] = float((torch.ones(1) * v).item())
    return r


def _raw_stream(dev):
    """cudaStream_t of torch's current stream on ``dev`` (the raw-handle accessor skips building a Stream object)."""
    try:
        return torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None else torch.cuda.current_device())
    except Exception:
        return _raw_stream(dev)


class _on_device:
    """``torch.cuda.device(dev)`` only when ``dev`` is not already the current device (the context manager costs ~5 us)."""

    def __init__(self, dev):
        self.ctx = None if (dev.index is None or dev.index == torch.cuda.current_device()) else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


_PINNED_POOL = []
_PINNED_NEXT = [0]


def _pinned_ctrl():
    """A page-locked buffer for the control block of an asynchronous forward (lqpb_forward_async_*), from a small ring:
    nobody reads a buffer unless its own call asks for ``iter`` / ``status``, so reuse 32 calls later is harmless."""
    if not _PINNED_POOL:
        nbytes = int(_abi.lib().lqpb_ctrl_bytes())
        _PINNED_POOL.extend(torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(32))
    k = _PINNED_NEXT[0]
    _PINNED_NEXT[0] = (k + 1) % len(_PINNED_POOL)
    return _PINNED_POOL[k]


_WS_BYTES_CACHE = {}


def _ws_bytes(kind, sfx, B, n, m):
    """lqpb_{forward,backward}_workspace_bytes_* (pure functions of the shape), cached."""
    key = (kind, sfx, B, n, m, os.environ.get("LQPB_FACTOR"))      # (the developer switch changes the layout)
    r = _WS_BYTES_CACHE.get(key)
    if r is None:
        r = _WS_BYTES_CACHE[key] = getattr(_abi.lib(), f"lqpb_{kind}_workspace_bytes_{sfx}")(B, n, m)
    return r


def _derive_config(control, n_x):
    """Flatten the control dict exactly as the reference unpacks it (:134-154)."""
    g = control.get
    cfg = _abi.Config()
    cfg.max_iters = int(g('max_iters', 10_000))
    cfg.eps_abs = max(g('eps_abs', 1e-3), 1e-12)
    cfg.eps_rel = max(g('eps_rel', 1e-3), 1e-12)
    check = g('check_solved', max(round((n_x ** 0.5) / 10) * 10, 1))
    cfg.check_solved = int(check)
    rho = g('rho', None)
    cfg.rho_auto = 1 if rho is None else 0
    # a tensor rho (the reference broadcasts a (B,1,1) tensor, e.g. sol['rho'] fed back) travels as a device array
    # (lqpb_forward_warm_*'s rho0); a one-element tensor is just a number
    if torch.is_tensor(rho) and rho.numel() == 1:
        rho = float(rho)
    cfg.rho = 0.0 if (rho is None or torch.is_tensor(rho)) else float(rho)
    cfg.rho_min = g('rho_min', 1e-6)
    cfg.rho_max = g('rho_max', 1e6)
    cfg.adaptive_rho = 1 if g('adaptive_rho', False) else 0
    cfg.adaptive_rho_tol = g('adaptive_rho_tol', 5)
    it = g('adaptive_rho_iter', 100)
    cfg.adaptive_rho_iter = int(max(round(it / check) * check, 1))
    cfg.adaptive_rho_max_iter = int(g('adaptive_max_iter', 1000))
    cfg.adaptive_rho_threshold = _default_dtype_value(g('adaptive_rho_threshold', 1e-5))
    cfg.verbose = 1 if g('verbose', False) else 0
    cfg.scale = 1 if g('scale', False) else 0
    beta = g('beta')
    cfg.beta_auto = 1 if beta is None else 0
    cfg.beta = 0.0 if beta is None else float(beta)
    cfg.zero_clamp = _default_dtype_value(1e-16)
    if cfg.max_iters < 1:
        raise ValueError("max_iters must be >= 1")
    if cfg.check_solved < 1:
        raise ValueError("check_solved must be >= 1")
    return cfg


def _all_on_host_devices(devices):
    ds = [d for d in devices if d is not None]
    return len(ds) > 0 and all(d.type == "cpu" for d in ds)


def _all_on_host(tensors):
    ts = [t for t in tensors if t is not None]
    return len(ts) > 0 and all(t.device.type == "cpu" for t in ts)


def _host_view(t):
    """Contiguous, detached host tensor (no copy for the usual contiguous inputs)."""
    return None if t is None else t.detach().contiguous()


def _solve_device(Q, p, A, b, lb, ub, control, host_keys=None, prep=None, tape_cap=None, z0=None, u0=None,
                  want_status=False, keep_operators=False, allow_async=False):
    L = _abi.lib()
    out_device = p.device
    host_mode = _all_on_host((Q, p, A, b, lb, ub))
    hx = None
    prepared = False
    deferred = False
    tape_info = None
    pre_dv = _take_prefetched((Q, p, A, b, lb, ub)) if host_mode else None
    if pre_dv is not None:
        # the batch was announced (prefetch_inputs): its device copies are on their way / there; from here on this is the
        # device-pointer path, only the results travel back to the caller's host tensors
        host_mode = False
        dv = pre_dv
    elif host_mode:
        # CPU tensors in (the reference's callers): lqpb_forward_host_* uploads Q in chunks on a copy stream and
        # overlaps the per-problem setup with the transfer; the device copies it fills are kept for the backward
        hv = {k: _host_view(t) for k, t in dict(Q=Q, p=p, A=A, b=b, lb=lb, ub=ub).items()}
        dt = hv["p"].dtype
        for k, t in hv.items():
            if t is not None and t.dtype != dt:
                raise TypeError(f"all tensors must share one dtype, got {dt} and {t.dtype} ({k})")
        dev = _cuda_device(hv["p"])
        dv = {k: (None if t is None else torch.empty(t.shape, dtype=dt, device=dev)) for k, t in hv.items()}
        if _needs_staging(hv["Q"]):
            # pageable memory (what the reference's callers hold): staged upload, then the device-pointer path
            with _on_device(dev):
                st, cur = _copy_stream(dev), torch.cuda.current_stream(dev)
                st.wait_stream(cur)
                for k, t in hv.items():
                    if t is not None:
                        _copy_up(dv[k], t, st)
                cur.wait_stream(st)
            host_mode = False
    else:
        dv = _stage(dict(Q=Q, p=p, A=A, b=b, lb=lb, ub=ub))
    Qd, pd = dv["Q"], dv["p"]
    dev, dt = pd.device, pd.dtype
    sfx = _abi.suffix(dt)
    B, n = Qd.shape[0], pd.shape[1]
    m = get_ncon(dv["A"], dim=1)
    cfg = _derive_config(control, n)
    cfg.keep_operators = 1 if keep_operators else 0
    with _on_device(dev):
        new = lambda *shape: torch.empty(shape, dtype=dt, device=dev)
        x, z, u = new(B, n, 1), new(B, n, 1), new(B, n, 1)
        lams = new(B, 2 * n, 1)
        nus = new(B, m, 1) if m > 0 else None
        rho_t = new(B, 1, 1)
        ws_bytes = _ws_bytes("forward", sfx, B, n, m)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        info = _abi.Info()
        stream = _raw_stream(dev)
        flag = C.c_int32(0)
        # the blocked-sweep sizes (n + m > 128, fp32 and fp64) queue the backward's factorisation behind the solve (prep); the
        # others can take the asynchronous forward
        rho_vec = None
        user_rho_t = control.get('rho', None)
        if torch.is_tensor(user_rho_t) and user_rho_t.numel() > 1:
            if user_rho_t.numel() != B:
                raise ValueError(f"control['rho'] has {user_rho_t.numel()} entries for a batch of {B}")
            if host_mode or tape_cap is not None:
                raise NotImplementedError("a per-problem rho tensor needs CUDA input tensors and unroll=False")
            prep = None                # (the per-problem rho travels through lqpb_forward_warm_* only)
            rho_vec = user_rho_t.detach().to(device=dev, dtype=dt).reshape(B).contiguous()
        use_async = (allow_async and not host_mode and tape_cap is None and z0 is None and u0 is None and not cfg.verbose
                     and rho_vec is None and not (n + m > 128))
        if host_mode:
            hx = torch.empty((B, n, 1), dtype=dt, pin_memory=True)
            rc = getattr(L, f"lqpb_forward_host_{sfx}")(
                C.byref(cfg), B, n, m, _abi.ptr(hv["Q"]), _abi.ptr(hv["p"]), _abi.ptr(hv["A"]), _abi.ptr(hv["b"]),
                _abi.ptr(hv["lb"]), _abi.ptr(hv["ub"]), _abi.ptr(Qd), _abi.ptr(pd), _abi.ptr(dv["A"]),
                _abi.ptr(dv["b"]), _abi.ptr(dv["lb"]), _abi.ptr(dv["ub"]), _abi.ptr(x), _abi.ptr(z), _abi.ptr(u),
                _abi.ptr(lams), _abi.ptr(nus), _abi.ptr(rho_t), _abi.ptr(hx), C.byref(info), _abi.ptr(ws), ws_bytes,
                C.c_void_p(stream), 0, _abi.ptr(prep["ws"]) if prep else None, prep["ws"].numel() if prep else 0,
                1 if (prep and prep["kkt"]) else 0, C.byref(flag))
            _abi.check(rc, "lqpb_forward_host")
            prepared = bool(flag.value)
        elif tape_cap is not None:
            # unrolled mode, first-pass recording: the solve itself writes the tape (capacity tape_cap iterations)
            cap = int(min(tape_cap, cfg.max_iters))
            tape = [torch.empty((B, cap, n), dtype=dt, device=dev) for _ in range(3)]
            tape.append(torch.empty((B, cap, m), dtype=dt, device=dev) if m > 0 else None)
            segs = (C.c_int32 * 2)()
            rc = getattr(L, f"lqpb_unroll_forward_{sfx}")(
                C.byref(cfg), B, n, m, cap, 1, _abi.ptr(Qd), _abi.ptr(pd), _abi.ptr(dv["A"]), _abi.ptr(dv["b"]),
                _abi.ptr(dv["lb"]), _abi.ptr(dv["ub"]), _abi.ptr(x), _abi.ptr(z), _abi.ptr(u), _abi.ptr(lams),
                _abi.ptr(nus), _abi.ptr(rho_t), *[_abi.ptr(t) for t in tape], None, 0, segs, None, C.byref(info),
                _abi.ptr(ws), ws_bytes, C.c_void_p(stream))
            if rc == _abi.E_TAPE:
                return None                       # does not fit (refactorisation / more iterations): caller falls back
            _abi.check(rc, "lqpb_unroll_forward")
            tape_info = (tuple(tape), cap)
        elif use_async:
            # the layer only hands x to autograd: small problems run as ONE launch whose adaptive-rho refactorisations
            # happen on the device, so the call returns once the bound flags are known and the GPU keeps working while
            # Python travels on to .backward(); iter / status stay in a pinned buffer until somebody asks
            pinned = _pinned_ctrl()
            rc = getattr(L, f"lqpb_forward_async_{sfx}")(
                C.byref(cfg), B, n, m, _abi.ptr(Qd), _abi.ptr(pd), _abi.ptr(dv["A"]), _abi.ptr(dv["b"]),
                _abi.ptr(dv["lb"]), _abi.ptr(dv["ub"]), None, None, _abi.ptr(x), _abi.ptr(z), _abi.ptr(u),
                _abi.ptr(lams), _abi.ptr(nus), _abi.ptr(rho_t), _abi.ptr(pinned), C.byref(info), _abi.ptr(ws), ws_bytes,
                C.c_void_p(stream), C.byref(flag))
            _abi.check(rc, "lqpb_forward_async")
            deferred = bool(flag.value)
        elif prep is not None:
            rc = getattr(L, f"lqpb_forward_prep_{sfx}")(
                C.byref(cfg), B, n, m, _abi.ptr(Qd), _abi.ptr(pd), _abi.ptr(dv["A"]), _abi.ptr(dv["b"]),
                _abi.ptr(dv["lb"]), _abi.ptr(dv["ub"]), _abi.ptr(x), _abi.ptr(z), _abi.ptr(u), _abi.ptr(lams),
                _abi.ptr(nus), _abi.ptr(rho_t), C.byref(info), _abi.ptr(ws), ws_bytes, _abi.ptr(prep["ws"]),
                prep["ws"].numel(), 1 if prep["kkt"] else 0, C.byref(flag), C.c_void_p(stream))
            _abi.check(rc, "lqpb_forward_prep")
            prepared = bool(flag.value)
        elif z0 is not None or u0 is not None or rho_vec is not None:
            if (z0 is None) != (u0 is None):
                raise ValueError("a warm start needs both z0 and u0")
            wz = wu = None
            if z0 is not None:
                wz, wu = (t.detach().to(device=dev, dtype=dt).reshape(B, n).contiguous() for t in (z0, u0))
            rc = getattr(L, f"lqpb_forward_warm_{sfx}")(
                C.byref(cfg), B, n, m, _abi.ptr(Qd), _abi.ptr(pd), _abi.ptr(dv["A"]), _abi.ptr(dv["b"]),
                _abi.ptr(dv["lb"]), _abi.ptr(dv["ub"]), _abi.ptr(wz), _abi.ptr(wu), _abi.ptr(rho_vec), _abi.ptr(x),
                _abi.ptr(z), _abi.ptr(u),
                _abi.ptr(lams), _abi.ptr(nus), _abi.ptr(rho_t), C.byref(info), _abi.ptr(ws), ws_bytes, C.c_void_p(stream))
            _abi.check(rc, "lqpb_forward_warm")
        else:
            rc = getattr(L, f"lqpb_forward_{sfx}")(
                C.byref(cfg), B, n, m, _abi.ptr(Qd), _abi.ptr(pd), _abi.ptr(dv["A"]), _abi.ptr(dv["b"]),
                _abi.ptr(dv["lb"]), _abi.ptr(dv["ub"]), _abi.ptr(x), _abi.ptr(z), _abi.ptr(u), _abi.ptr(lams),
                _abi.ptr(nus), _abi.ptr(rho_t), C.byref(info), _abi.ptr(ws), ws_bytes, C.c_void_p(stream))
            _abi.check(rc, "lqpb_forward")
        extra = {}
        if want_status:
            resid = torch.empty((B, 4), dtype=dt, device=dev)
            conv = torch.empty((B,), dtype=torch.int32, device=dev)
            rc = getattr(L, f"lqpb_solution_status_{sfx}")(C.byref(cfg), B, n, m, _abi.ptr(ws), ws_bytes, _abi.ptr(resid),
                                                          _abi.ptr(conv), C.c_void_p(stream))
            _abi.check(rc, "lqpb_solution_status")
            if out_device.type == "cpu":
                resid, conv = resid.cpu(), conv.cpu()
            extra = {"converged": conv != 0, "primal_residual": resid[:, 0].reshape(B, 1, 1),
                     "dual_residual": resid[:, 1].reshape(B, 1, 1), "primal_tolerance": resid[:, 2].reshape(B, 1, 1),
                     "dual_tolerance": resid[:, 3].reshape(B, 1, 1)}
    if cfg.verbose:
        for k in range(info.n_log):          # same text as the reference prints (:289-294)
            print(f'iteration = {info.log_iter[k]}')
            print(f'|| primal_error|| = {info.log_primal[k]:.10f}')
            print(f'|| dual_error|| = {info.log_dual[k]:.10f}')
    any_ineq = bool(info.any_lb or info.any_ub)
    user_rho = control.get('rho', None)
    if not any_ineq:
        rho = 0                                   # :157-158
    elif user_rho is None or info.n_factor > 1 or deferred:
        rho = rho_t                               # :200-203 / :248-250 (deferred: the per-problem tensor always --
                                                  # it holds the caller's scalar when rho was given and never adapted)
    else:
        rho = user_rho
    # results go back to where the caller's tensors live; the layer only hands x to autograd (host_keys)
    keys = ("x", "z", "u", "lams", "nus", "rho")
    vals = [x, z, u, lams, nus, rho if torch.is_tensor(rho) else None]
    want = [out_device if (host_keys is None or k in host_keys) else None for k in keys]
    if hx is not None:
        want[0] = None                            # x already came back inside lqpb_forward_host_*
    x_out, hz, hu, hlams, hnus, hrho = _to_devices(vals, want)
    hx = hx if hx is not None else x_out
    rho_out = hrho if torch.is_tensor(rho) else rho
    return {**extra, "x": hx, "z": hz, "u": hu, "lams": hlams, "nus": hnus, "rho": rho_out,
            "iter": int(info.iter), "status": int(info.status), "n_factor": int(info.n_factor), "_deferred": deferred,
            "_any_lb": bool(info.any_lb), "_any_ub": bool(info.any_ub), "_dev": dv,
            "_x_dev": x, "_u_dev": u, "_lams_dev": lams, "_nus_dev": nus, "_ws": ws, "_cfg": cfg, "_prepared": prepared, "_tape": tape_info,
            "rho_dev": rho if torch.is_tensor(rho) else None}


# ------------------------------------------------------------------------------------------
# unrolled mode (control['unroll'] = True)
# ------------------------------------------------------------------------------------------
def _solve_unrolled(Q, p, A, b, lb, ub, control):
    """``unroll=True`` (reference :13-15): ``x`` comes back attached to an autograd graph that differentiates the
    scaling (:161-197), the rho selection (:200-203), every ADMM iteration (:259-282), every adaptive-rho update
    (:237-256, where ``rho_new = rho * ratio`` stays in the graph) and the un-scaling (:316).

    The iterations -- all the O(T n^2) work -- are autograd nodes backed by CUDA kernels (``_UnrolledSegment``): the
    loop is recorded on a tape (``lqpb_unroll_record_*`` / ``lqpb_unroll_forward_*``) and a node's backward is the
    reverse-sweep kernel over its iteration range (``lqpb_unroll_backward_*``: one symmetric K11 GEMV per recorded
    iteration, then dQ~ / dA~ as rank-T products over the tape instead of the reference's dense ``dx xv^T`` per
    iteration, lu_layer.py:53).  Without an adaptive-rho update the whole loop is ONE node.  An update at iteration
    i reads the residual norms of the last check (iteration c = i - check_solved), so the loop is cut after c and
    before i and the T-independent pieces in between -- like the O(n^2) map between the caller's tensors and the
    scaled problem (D, E, Q~ = D Q D, rho = ||Q~||_F / sqrt(n)) -- are written with the torch operators the
    reference itself uses, on the CUDA copies, so that autograd applies exactly the reference's (sub)gradient
    conventions for ``norm(inf)``, ``quantile``, ``maximum`` and ``clamp``."""
    L = _abi.lib()
    out_device = p.device
    for t in (Q, p, A, b, lb, ub):
        if t is not None and t.dtype != p.dtype:
            raise TypeError(f"all tensors must share one dtype, got {p.dtype} and {t.dtype}")
    dev = _cuda_device(next((t for t in (Q, p, A, b, lb, ub) if t is not None and t.is_cuda), p))
    # device copies that stay connected to the caller's leaves (``.to`` is differentiable)
    Qd, pd, Ad, bd, lbd, ubd = (None if t is None else t.to(dev).contiguous() for t in (Q, p, A, b, lb, ub))
    plain = dict(control)
    plain['unroll'] = False
    # first try: the solve records itself (tape capacity _UNROLL_TAPE_CAP iterations); a solve that needs an adaptive-rho
    # refactorisation or more iterations falls back to a plain solve followed by a recording pass
    sol = _solve_device(Qd, pd, Ad, bd, lbd, ubd, plain, host_keys=(), tape_cap=_UNROLL_TAPE_CAP)
    first_pass = sol is not None
    if not first_pass:
        sol = _solve_device(Qd, pd, Ad, bd, lbd, ubd, plain, host_keys=(), keep_operators=True)
    any_lb, any_ub = sol["_any_lb"], sol["_any_ub"]
    B, n, dt = Qd.shape[0], pd.shape[1], pd.dtype
    m = get_ncon(Ad, dim=1)
    sfx = _abi.suffix(dt)
    K, S = sol["iter"] + 1, sol["n_factor"]
    cfg, ws = sol["_cfg"], sol["_ws"]
    state = dict(ws=ws, B=B, n=n, m=m, n_iter=K, snaps=None, snap_each=0)
    wants = None
    if first_pass:
        (*tape, tape_nu), state["n_iter"] = sol["_tape"]       # rows beyond K are unused; n_iter is the row stride
        seg_start = [0, K]
    with _on_device(dev):
        if first_pass:
            pass
        elif S == 1:
            tape = [torch.empty((B, K, n), dtype=dt, device=dev) for _ in range(3)]
            tape_nu = torch.empty((B, K, m), dtype=dt, device=dev) if m > 0 else None
            stream = _raw_stream(dev)
            rc = getattr(L, f"lqpb_unroll_record_{sfx}")(
                C.byref(cfg), B, n, m, K, _abi.ptr(ws), ws.numel(), _abi.ptr(tape[0]), _abi.ptr(tape[1]),
                _abi.ptr(tape[2]), _abi.ptr(tape_nu), C.c_void_p(stream))
            _abi.check(rc, "lqpb_unroll_record")
            seg_start = [0, K]
        else:
            # adaptive-rho updates: a full recording solve that also keeps the operators of every segment
            tape = [torch.empty((B, K, n), dtype=dt, device=dev) for _ in range(3)]
            tape_nu = torch.empty((B, K, m), dtype=dt, device=dev) if m > 0 else None
            stream = _raw_stream(dev)
            d = sol["_dev"]
            each = getattr(L, f"lqpb_unroll_snapshot_bytes_{sfx}")(B, n, m)
            snaps = torch.empty(S * each, dtype=torch.uint8, device=dev)
            wants = torch.empty((S - 1, B), dtype=torch.int32, device=dev)
            segs = (C.c_int32 * (S + 1))()
            info = _abi.Info()
            scratch = [sol["_x_dev"], torch.empty_like(sol["_x_dev"]), sol["_u_dev"], sol["_lams_dev"], sol["_nus_dev"],
                       torch.empty((B, 1, 1), dtype=dt, device=dev)]
            rc = getattr(L, f"lqpb_unroll_forward_{sfx}")(
                C.byref(cfg), B, n, m, K, S, _abi.ptr(d["Q"]), _abi.ptr(d["p"]), _abi.ptr(d["A"]), _abi.ptr(d["b"]),
                _abi.ptr(d["lb"]), _abi.ptr(d["ub"]), *[_abi.ptr(t) for t in scratch], _abi.ptr(tape[0]),
                _abi.ptr(tape[1]), _abi.ptr(tape[2]), _abi.ptr(tape_nu), _abi.ptr(snaps), snaps.numel(), segs,
                _abi.ptr(wants), C.byref(info), _abi.ptr(ws), ws.numel(), C.c_void_p(stream))
            _abi.check(rc, "lqpb_unroll_forward")
            seg_start = list(segs)
            state.update(snaps=snaps, snap_each=each)
    state["tape"] = (*tape, tape_nu)

    # without an adaptive-rho update nothing downstream reads the VALUES of Q~ (the kernels hold their own copy): the
    # two O(n^2) glue steps Q~ = D Q D and rho = ||Q~||_F / sqrt(n) then exist only as one fused adjoint kernel
    Qt, pt, At, bt, lbt, ubt, D, rho = _scaled_problem(Qd, pd, Ad, bd, lbd, ubd, control, any_lb, any_ub,
                                                       fused_rho=sol["rho_dev"] if S == 1 else None, fused=S == 1,
                                                       ws=ws if S == 1 else None)
    check = cfg.check_solved
    z_in = u_in = x_l = None
    at_check = rho_at_check = None
    for s_idx in range(S):
        lo, hi = seg_start[s_idx], seg_start[s_idx + 1] - 1
        # every segment is cut after the last check iteration it contains: an adaptive-rho update reads the residual
        # norms of the most recent check (:239-243), which may lie in THIS segment or -- when adaptive_rho_iter is
        # shorter than check_solved -- in an earlier one, whose state (and the rho then in force) is kept
        c = (hi // check) * check
        pieces = [(lo, hi)] if (s_idx == S - 1 or c < lo) else [(lo, c)] + ([(c + 1, hi)] if c < hi else [])
        for (k_lo, k_hi) in pieces:
            x_l, z_l, u_l, zp_l = _UnrolledSegment.apply(
                Qt, pt, At, bt, lbt if any_lb else None, ubt if any_ub else None, rho if torch.is_tensor(rho) else None,
                z_in, u_in, state, k_lo, k_hi, s_idx)
            z_in, u_in = z_l, u_l
            if s_idx < S - 1 and c >= lo and k_hi == c:
                at_check, rho_at_check = (x_l, z_l, u_l, zp_l), rho
        if s_idx < S - 1:
            rho = _adapted_rho(rho, rho_at_check, at_check, wants[s_idx].view(B, 1, 1) != 0, Qt, pd, D, cfg, control)
    x = D * x_l                                                          # :316
    return x if x.device == out_device else x.to(out_device)


def _adapted_rho(rho, rho_chk, at_check, wants, Qt, p, D, cfg, control):
    """rho after an adaptive update (reference :239-250) as a differentiable function of the state recorded at the
    last check (:286-304) and of ``rho_chk``, the rho in force at that check (it differs from ``rho`` when several
    updates follow one check).  ``wants`` is the kernel's do_rho_update mask (:310-311), a decision, not differentiated."""
    x, z, u, z_prev = at_check
    ninf = lambda t: torch.linalg.norm(t, ord=_INF, dim=1, keepdim=True)
    tiny = torch.full((1,), cfg.zero_clamp, dtype=x.dtype, device=x.device)                    # :229-230
    r, s = x - z, rho_chk * (z - z_prev)                                                      # :279-280
    primal, dual = ninf(D * r), ninf(D * s)                                                  # :286-287
    scale_p = torch.maximum(torch.maximum(ninf(D * x), ninf(D * z)), tiny)                   # :296-301
    scale_d = torch.maximum(torch.maximum(torch.maximum(ninf(rho_chk * D * u), ninf(torch.matmul(Qt, x) / D)), ninf(p)), tiny)
    num = torch.clamp(primal / scale_p, min=cfg.zero_clamp)                                  # :239-242
    den = torch.clamp(dual / scale_d, min=cfg.zero_clamp)
    ratio = (num / den) ** 0.5                                                               # :243
    rho = rho * torch.logical_not(wants) + (rho * ratio) * wants                             # :248-249
    return torch.clamp(rho, min=control.get('rho_min', 1e-6), max=control.get('rho_max', 1e6))


def _scaled_problem(Q, p, A, b, lb, ub, control, any_lb, any_ub, fused=False, fused_rho=None, ws=None):
    """Scaling and rho selection of the reference (:156-203) as differentiable torch expressions of the device
    copies.  Returns ``(Q~, p~, A~, b~, lb~, ub~, D, rho)``; ``D`` is ``(B,n,1)`` (or 1.0 without scaling) and ``rho``
    a ``(B,1,1)`` tensor when it is selected from ``||Q~||_F``, otherwise the caller's number."""
    n = p.shape[1]
    any_ineq = any_lb or any_ub
    rho = control.get('rho', None) if any_ineq else 0                   # :157-158
    D = 1.0
    Dv = None
    if control.get('scale', False) and fused and ws is not None:
        # the whole map from the caller's tensors to the scaled problem (:161-203) is ONE autograd node backed by
        # kernels: values out of the recording solve's workspace, adjoint = scale_grad (Q~ = D Q D, rho), scale_vec_grad
        # (D, p~, A~, b~, lb~, ub~ with torch's subgradient rules) and the column-max scatter, all on one dense buffer
        auto = rho is None
        Q, rho_t, Dv, p, At_, bt_, lbt_, ubt_ = _ScaledProblem.apply(
            Q, p, A, b, lb, ub, ws, fused_rho if auto else None, control.get('beta'), bool(any_lb), bool(any_ub),
            control.get('rho_min', 1e-6), control.get('rho_max', 1e6))
        if auto:
            rho = rho_t
        if A is not None:
            A, b = At_, bt_
        if any_ineq:
            lb, ub = lbt_, ubt_
        return Q, p, A, b, lb, ub, Dv.unsqueeze(2), rho
    elif control.get('scale', False):
        colmax = _ColumnMax.apply(Q) if (fused and Q.requires_grad) else torch.linalg.norm(Q, ord=_INF, dim=1)   # :163
        bad = colmax <= 0.0
        if bool(bad.any()):                                              # :164-168
            floor = colmax.mean(dim=1).clamp(min=1e-6).unsqueeze(1)
            colmax = torch.where(bad, torch.maximum(colmax, floor), colmax)
        D = torch.sqrt(1 / colmax)                                       # :170
        beta = control.get('beta')
        if beta is None:                                                 # :171-174
            q = torch.quantile(D, torch.tensor([0.10, 0.90], dtype=D.dtype, device=D.device), dim=1)
            beta = (1 - q[0] / q[1]).unsqueeze(1)
        D = (1 - beta) * D + beta * D.mean(dim=1, keepdim=True)          # :175
        Dv = D
        if not fused:
            Q = D.unsqueeze(2) * Q * D.unsqueeze(1)                      # :176
        p = D.unsqueeze(2) * p                                           # :177
        if A is not None:
            A = A * D.unsqueeze(1)                                       # :180
            rown = torch.linalg.norm(A, ord=_INF, dim=2)                 # :181
            bad = rown <= 0.0
            if bool(bad.any()):                                          # :182-186
                floor = rown.mean(dim=1).clamp(min=1e-6).unsqueeze(1)
                rown = torch.where(bad, torch.maximum(rown, floor), rown)
            E = (1 / rown).unsqueeze(2)                                  # :187-188
            A, b = E * A, E * b                                          # :189-190
        D = D.unsqueeze(2)
        if any_ineq:
            lb, ub = lb / D, ub / D                                      # :193-194
    if fused:
        auto = rho is None
        if Q.requires_grad and (Dv is not None or auto):
            Q, rho_t = _ScaledQAndRho.apply(Q, Dv, fused_rho if auto else None, control.get('rho_min', 1e-6),
                                            control.get('rho_max', 1e6))
            if auto:
                rho = rho_t
        elif auto:
            rho = fused_rho                                              # no gradient can reach Q: plain values
        return Q, p, A, b, lb, ub, D, rho
    if rho is None:                                                      # :200-203
        rho = torch.linalg.matrix_norm(Q, keepdim=True) / n ** 0.5
        rho = torch.clamp(rho, min=control.get('rho_min', 1e-6), max=control.get('rho_max', 1e6))
    return Q, p, A, b, lb, ub, D, rho


class _ColumnMax(torch.autograd.Function):
    """``norm(Q, inf, dim=1)`` (:163) with a sparse adjoint: one entry per column when the maximiser is unique.  torch's
    own backward of the inf-norm splits the gradient evenly among EXACT ties (constant blocks, equicorrelation
    matrices, |Q_ij| == Q_jj); columns with ties take that dense rule here too, so dQ matches the reference's."""

    @staticmethod
    def forward(ctx, Q):
        aQ = Q.abs()
        colmax, idx = aQ.max(dim=1)
        ties = (aQ == colmax.unsqueeze(1)).sum(dim=1)              # maximisers per column
        sign = torch.sign(torch.gather(Q, 1, idx.unsqueeze(1)).squeeze(1))
        ctx.tied = bool((ties > 1).any())
        if ctx.tied:
            ctx.save_for_backward(idx, sign, Q, colmax, ties)
        else:
            ctx.save_for_backward(idx, sign)
        ctx.shape = Q.shape
        return colmax

    @staticmethod
    def backward(ctx, g):
        if ctx.tied:
            idx, sign, Q, colmax, ties = ctx.saved_tensors
            mask = Q.abs() == colmax.unsqueeze(1)
            return mask * torch.sign(Q) * (g / ties).unsqueeze(1)
        idx, sign = ctx.saved_tensors
        out = torch.zeros(ctx.shape, dtype=g.dtype, device=g.device)
        out.scatter_(1, idx.unsqueeze(1), (sign * g).unsqueeze(1))
        return out


class _ScaledProblem(torch.autograd.Function):
    """(Q, p, A, b, lb, ub) -> (Q~, rho, D, p~, A~, b~, lb~, ub~): the scaling and rho selection of :161-203 as one
    autograd node of the unrolled mode (scale=True, no adaptive-rho update inside the loop).

    Forward: no arithmetic -- the recording solve computed all of it; D and the scaled vectors are copied out of its
    workspace (``lqpb_unroll_scaled_vectors_*``), ``rho_vals`` is its rho, and since nothing downstream reads the VALUES
    of Q~ (the kernels hold their own packed copy) Q~ is a zero-stride placeholder.  Backward: three kernels on one dense
    buffer -- ``lqpb_unroll_scale_grad_*`` (adjoint of Q~ = D Q D and of rho = clamp(||Q~||_F / sqrt(n)), in place),
    ``lqpb_unroll_scale_vec_grad_*`` (adjoint of D / p~ / A~ / b~ / lb~ / ub~ with torch's subgradient conventions for
    norm(inf), quantile, maximum / where) and ``lqpb_unroll_colmax_grad_*`` (adjoint of the column inf-norms of :163,
    exact ties split evenly like torch's, added in place)."""

    @staticmethod
    def forward(ctx, Q, p, A, b, lb, ub, ws, rho_vals, beta, any_lb, any_ub, rho_min, rho_max):
        L = _abi.lib()
        dev, dt = p.device, p.dtype
        sfx = _abi.suffix(dt)
        B, n = p.shape[0], p.shape[1]
        m = get_ncon(A, dim=1)
        with _on_device(dev):
            new = lambda *shape: torch.empty(shape, dtype=dt, device=dev)
            D, pt, lbt, ubt = new(B, n), new(B, n, 1), new(B, n, 1), new(B, n, 1)
            At = new(B, m, n) if m > 0 else None
            bt = new(B, m, 1) if m > 0 else None
            E = new(B, m) if m > 0 else None
            rc = getattr(L, f"lqpb_unroll_scaled_vectors_{sfx}")(
                B, n, m, _abi.ptr(ws), ws.numel(), _abi.ptr(D), _abi.ptr(pt), _abi.ptr(At), _abi.ptr(bt), _abi.ptr(lbt),
                _abi.ptr(ubt), _abi.ptr(E), C.c_void_p(_raw_stream(dev)))
            _abi.check(rc, "lqpb_unroll_scaled_vectors")
        ctx.save_for_backward(Q, p, A, b, lb, ub, D, E, rho_vals)
        ctx.meta = (B, n, m, beta, any_lb, any_ub, rho_min, rho_max)
        token = torch.zeros((), dtype=dt, device=dev).expand(Q.shape)
        rho = rho_vals.detach().clone() if rho_vals is not None else torch.zeros((), dtype=dt, device=dev)
        if rho_vals is None:
            ctx.mark_non_differentiable(rho)
        ctx.set_materialize_grads(False)
        return token, rho, D, pt, At, bt, lbt, ubt

    @staticmethod
    def backward(ctx, gQt, grho, gD, gpt, gAt, gbt, glbt, gubt):
        L = _abi.lib()
        Q, p, A, b, lb, ub, D, E, rho_vals = ctx.saved_tensors
        B, n, m, beta, any_lb, any_ub, rho_min, rho_max = ctx.meta
        dev, dt = p.device, p.dtype
        sfx = _abi.suffix(dt)
        need_Q = ctx.needs_input_grad[0]
        cont = lambda t: None if t is None else t.detach().to(device=dev, dtype=dt).contiguous()
        gD, gpt, gAt, gbt, glbt, gubt = (cont(t) for t in (gD, gpt, gAt, gbt, glbt, gubt))
        Qc = Q.detach().contiguous()
        with _on_device(dev):
            stream = C.c_void_p(_raw_stream(dev))
            new = lambda *shape: torch.empty(shape, dtype=dt, device=dev)
            G = gDq = None
            if need_Q:
                # the adjoint of Q~ has one producer (the loop node) and one consumer (this node): overwritten in place
                G = gQt.contiguous() if gQt is not None else torch.zeros((B, n, n), dtype=dt, device=dev)
                coef = None
                if grho is not None and rho_vals is not None:
                    r = rho_vals.reshape(B)
                    inside = (r > rho_min) & (r < rho_max)
                    coef = torch.where(inside, grho.reshape(B) / (n * r), torch.zeros_like(r)).contiguous()
                gDq = new(B, n)
                nscr = getattr(L, f"lqpb_unroll_scale_grad_scratch_elems_{sfx}")(B, n)
                scratch = torch.empty(nscr, dtype=dt, device=dev)
                rc = getattr(L, f"lqpb_unroll_scale_grad_{sfx}")(B, n, _abi.ptr(G), _abi.ptr(Qc), _abi.ptr(D), _abi.ptr(coef),
                                                                 _abi.ptr(gDq), _abi.ptr(scratch), stream)
                _abi.check(rc, "lqpb_unroll_scale_grad")
            colmax = new(B, n)
            _abi.check(getattr(L, f"lqpb_unroll_colmax_{sfx}")(B, n, _abi.ptr(Qc), _abi.ptr(colmax), stream), "lqpb_unroll_colmax")
            gc, gp, glb, gub = new(B, n), new(B, n, 1), new(B, n, 1), new(B, n, 1)
            gA = new(B, m, n) if m > 0 else None
            gb = new(B, m, 1) if m > 0 else None
            rc = getattr(L, f"lqpb_unroll_scale_vec_grad_{sfx}")(
                B, n, m, 1 if beta is None else 0, 0.0 if beta is None else float(beta), 1 if any_lb else 0,
                1 if any_ub else 0, _abi.ptr(colmax), _abi.ptr(p.contiguous()), _abi.ptr(cont(A)), _abi.ptr(cont(b)),
                _abi.ptr(lb.contiguous()), _abi.ptr(ub.contiguous()), _abi.ptr(D), _abi.ptr(E), _abi.ptr(gD), _abi.ptr(gDq),
                _abi.ptr(gpt), _abi.ptr(gAt), _abi.ptr(gbt), _abi.ptr(glbt), _abi.ptr(gubt), _abi.ptr(gc), _abi.ptr(gp),
                _abi.ptr(gA), _abi.ptr(gb), _abi.ptr(glb), _abi.ptr(gub), stream)
            _abi.check(rc, "lqpb_unroll_scale_vec_grad")
            if need_Q:
                rc = getattr(L, f"lqpb_unroll_colmax_grad_{sfx}")(B, n, _abi.ptr(Qc), _abi.ptr(colmax), _abi.ptr(gc), _abi.ptr(G),
                                                                  stream)
                _abi.check(rc, "lqpb_unroll_colmax_grad")
        any_ineq = any_lb or any_ub
        return (G, gp, gA, gb, glb if any_ineq else None, gub if any_ineq else None) + (None,) * 7


class _ScaledQAndRho(torch.autograd.Function):
    """``Q~ = D Q D`` (:176) and ``rho = clamp(||Q~||_F / sqrt(n))`` (:200-203) as one autograd node without a forward
    computation: the solver kernels scaled Q and selected rho themselves (``rho_vals`` is their rho), and no consumer
    reads the values of Q~, so the forward returns a zero-stride placeholder of the right shape.  The backward is one
    fused kernel (``lqpb_unroll_scale_grad_*``) over the adjoint of Q~ that the reverse sweep delivers."""

    @staticmethod
    def forward(ctx, Q, Dv, rho_vals, rho_min, rho_max):
        ctx.save_for_backward(Q, Dv, rho_vals)
        ctx.clamp = (rho_min, rho_max)
        token = torch.zeros((), dtype=Q.dtype, device=Q.device).expand(Q.shape)
        rho = rho_vals.detach().clone() if rho_vals is not None else torch.zeros((), dtype=Q.dtype, device=Q.device)
        if rho_vals is None:
            ctx.mark_non_differentiable(rho)
        ctx.set_materialize_grads(False)
        return token, rho

    @staticmethod
    def backward(ctx, gQt, grho):
        L = _abi.lib()
        Q, Dv, rho_vals = ctx.saved_tensors
        B, n = Q.shape[0], Q.shape[1]
        dev, dt = Q.device, Q.dtype
        sfx = _abi.suffix(dt)
        with _on_device(dev):
            # the adjoint of Q~ has one producer (the loop node) and one consumer (this node): overwritten in place
            G = gQt.contiguous() if gQt is not None else torch.zeros((B, n, n), dtype=dt, device=dev)
            coef = None
            if grho is not None and rho_vals is not None:
                r = rho_vals.reshape(B)
                inside = (r > ctx.clamp[0]) & (r < ctx.clamp[1])
                coef = torch.where(inside, grho.reshape(B) / (n * r), torch.zeros_like(r)).contiguous()
            gD = torch.empty((B, n), dtype=dt, device=dev) if Dv is not None else None
            nscr = getattr(L, f"lqpb_unroll_scale_grad_scratch_elems_{sfx}")(B, n)
            scratch = torch.empty(nscr, dtype=dt, device=dev)
            stream = _raw_stream(dev)
            Dc = Dv.detach().contiguous() if Dv is not None else None
            rc = getattr(L, f"lqpb_unroll_scale_grad_{sfx}")(
                B, n, _abi.ptr(G), _abi.ptr(Q.detach().contiguous()), _abi.ptr(Dc), _abi.ptr(coef), _abi.ptr(gD),
                _abi.ptr(scratch), C.c_void_p(stream))
            _abi.check(rc, "lqpb_unroll_scale_grad")
        return G, gD, None, None, None


class _UnrolledSegment(torch.autograd.Function):
    """Iterations k_lo .. k_hi of the recorded loop (one operator segment: rho and K11 fixed) as one autograd node:
    (scaled problem data, rho, z_{k_lo-1}, u_{k_lo-1}) -> (x~, z, u at k_hi, z at k_hi - 1), reference :259-282 with
    every KKT solve differentiated as in lu_layer.py:41-58.  The forward only reads the tape."""

    @staticmethod
    def forward(ctx, Qt, pt, At, bt, lbt, ubt, rho, z_in, u_in, state, k_lo, k_hi, seg):
        tx, tz, tu, _ = state["tape"]
        ctx.set_materialize_grads(False)
        ctx.state, ctx.range, ctx.seg = state, (k_lo, k_hi), seg
        ctx.shapes = tuple(None if t is None else t.shape for t in (Qt, pt, At, bt, lbt, ubt, rho, z_in, u_in))
        pick = lambda t, k: t[:, k, :].unsqueeze(2).clone()
        z_prev = pick(tz, k_hi - 1) if k_hi > 0 else torch.zeros_like(pick(tz, 0))
        return pick(tx, k_hi), pick(tz, k_hi), pick(tu, k_hi), z_prev

    @staticmethod
    def backward(ctx, gx, gz, gu, gzp):
        L = _abi.lib()
        st = ctx.state
        B, n, m, K = st["B"], st["n"], st["m"], st["n_iter"]
        k_lo, k_hi = ctx.range
        tx, tz, tu, tnu = st["tape"]
        dev, dt = tx.device, tx.dtype
        sfx = _abi.suffix(dt)
        ws = st["ws"]
        need = ctx.needs_input_grad
        prep = lambda t: None if t is None else t.detach().to(device=dev, dtype=dt).contiguous()
        gx, gz, gu, gzp = prep(gx), prep(gz), prep(gu), prep(gzp)
        snap = None
        if st["snaps"] is not None:
            snap = C.c_void_p(st["snaps"].data_ptr() + ctx.seg * st["snap_each"])
        with _on_device(dev):
            new = lambda *shape: torch.empty(shape, dtype=dt, device=dev)
            if "tw" not in st:                        # scratch for the adjoint solves, shared by all segments
                st["tw"] = new(B, K, n)
                st["twnu"] = new(B, K, m) if m > 0 else None
            gQ = new(B, n, n) if need[0] else None
            gA = new(B, m, n) if (m > 0 and need[2]) else None
            gp, glb, gub, grho = new(B, n, 1), new(B, n, 1), new(B, n, 1), new(B, 1, 1)
            gb = new(B, m, 1) if m > 0 else None
            gz_in = new(B, n, 1) if need[7] else None
            gu_in = new(B, n, 1) if need[8] else None
            stream = _raw_stream(dev)
            rc = getattr(L, f"lqpb_unroll_backward_{sfx}")(
                B, n, m, K, k_lo, k_hi, _abi.ptr(ws), ws.numel(), snap, _abi.ptr(gx), _abi.ptr(gz), _abi.ptr(gu),
                _abi.ptr(gzp), _abi.ptr(tx), _abi.ptr(tz), _abi.ptr(tu), _abi.ptr(tnu), _abi.ptr(st["tw"]),
                _abi.ptr(st["twnu"]), _abi.ptr(gQ), _abi.ptr(gp), _abi.ptr(gA), _abi.ptr(gb), _abi.ptr(glb), _abi.ptr(gub),
                _abi.ptr(grho), _abi.ptr(gz_in), _abi.ptr(gu_in), C.c_void_p(stream))
            _abi.check(rc, "lqpb_unroll_backward")
        outs = [gQ, gp, gA, gb, glb, gub, grho, gz_in, gu_in]
        outs = [o if (shape is not None and nd) else None for o, shape, nd in zip(outs, ctx.shapes, need[:9])]
        return (*outs, None, None, None, None)


def _prealloc_backward(Q, p, A, need):
    """Output buffers (only the gradients asked for) and the workspace of one backward call on ``p``'s device."""
    L = _abi.lib()
    dev, dt = p.device, p.dtype
    B, n = Q.shape[0], p.shape[1]
    m = get_ncon(A, dim=1)
    nQ, np_, nA, nb, nlb, nub = need
    with _on_device(dev):
        new = lambda *shape: torch.empty(shape, dtype=dt, device=dev)
        ws_bytes = _ws_bytes("backward", _abi.suffix(dt), B, n, m)
        return dict(dQ=new(B, n, n) if nQ else None, dp=new(B, n, 1) if np_ else None,
                    dA=new(B, m, n) if (nA and m > 0) else None, db=new(B, m, 1) if (nb and m > 0) else None,
                    dlb=new(B, n, 1) if nlb else None, dub=new(B, n, 1) if nub else None,
                    ws=torch.empty(ws_bytes, dtype=torch.uint8, device=dev), key=(dev, dt, B, n, m, tuple(need)))


def _grad_device(dl_dz, x, u, lams, nus, Q, A, lb, ub, rho, need, pre=None, finish=False):
    """All tensor arguments already on one CUDA device (except dl_dz, which is staged here)."""
    L = _abi.lib()
    dev, dt = x.device, x.dtype
    sfx = _abi.suffix(dt)
    B, n = Q.shape[0], Q.shape[1]
    m = get_ncon(A, dim=1)
    g = dl_dz.detach()
    if g.dtype != dt:
        raise TypeError(f"dl_dz has dtype {g.dtype}, expected {dt}")
    if g.device != dev:
        g = g.to(dev, non_blocking=True)
    g = g.contiguous()
    if rho is None:
        rho = 1.0                                  # :356-357
    rho_dev = None
    rho_scalar = 0.0
    if torch.is_tensor(rho):
        rho_dev = rho.detach().to(device=dev, dtype=dt).reshape(-1).contiguous()
        if rho_dev.numel() == 1 and B > 1:
            rho_dev = rho_dev.expand(B).contiguous()
    else:
        rho_scalar = float(rho)
    if pre is None or pre["key"] != (dev, dt, B, n, m, tuple(need)):
        pre = _prealloc_backward(Q, x, A, need)
    dQ, dp, dA, db, dlb, dub, ws = (pre[k] for k in ("dQ", "dp", "dA", "db", "dlb", "dub", "ws"))
    ws_bytes = ws.numel()
    with _on_device(dev):
        stream = _raw_stream(dev)
        if finish:                          # the forward call left the factorised adjoint system in pre["ws"]
            rc = getattr(L, f"lqpb_backward_finish_{sfx}")(
                B, n, m, 0, _abi.ptr(g), _abi.ptr(x), _abi.ptr(u), _abi.ptr(lams), _abi.ptr(nus), _abi.ptr(Q), _abi.ptr(A),
                _abi.ptr(lb), _abi.ptr(ub), _abi.ptr(rho_dev), rho_scalar, _abi.ptr(dQ), _abi.ptr(dp), _abi.ptr(dA),
                _abi.ptr(db), _abi.ptr(dlb), _abi.ptr(dub), _abi.ptr(ws), ws_bytes, C.c_void_p(stream))
            _abi.check(rc, "lqpb_backward_finish")
            return dQ, dp, dA, db, dlb, dub
        rc = getattr(L, f"lqpb_backward_{sfx}")(
            B, n, m, _abi.ptr(g), _abi.ptr(x), _abi.ptr(u), _abi.ptr(lams), _abi.ptr(nus), _abi.ptr(Q), _abi.ptr(A),
            _abi.ptr(lb), _abi.ptr(ub), _abi.ptr(rho_dev), rho_scalar, _abi.ptr(dQ), _abi.ptr(dp), _abi.ptr(dA),
            _abi.ptr(db), _abi.ptr(dlb), _abi.ptr(dub), _abi.ptr(ws), ws_bytes, C.c_void_p(stream))
        _abi.check(rc, "lqpb_backward")
    return dQ, dp, dA, db, dlb, dub


def _grad_kkt_device(dl_dz, x, lams, nus, Q, A, lb, ub, need, any_bounds, pre=None, finish=False):
    """KKT backward on one CUDA device.  ``any_bounds`` = ``(any_lb, any_ub)`` when the caller already knows the
    flags (the layer does, from the forward solve: the call stays asynchronous) or ``None`` to have them
    evaluated on the device (one stream synchronisation)."""
    L = _abi.lib()
    dev, dt = x.device, x.dtype
    sfx = _abi.suffix(dt)
    B, n = Q.shape[0], Q.shape[1]
    m = get_ncon(A, dim=1)
    g = dl_dz.detach()
    if g.dtype != dt:
        raise TypeError(f"dl_dz has dtype {g.dtype}, expected {dt}")
    if g.device != dev:
        g = g.to(dev, non_blocking=True)
    g = g.contiguous()
    nQ, np_, nA, nb, nlb, nub = need
    flags = (C.c_int32 * 2)()
    if pre is None or pre["key"] != (dev, dt, B, n, m, tuple(need)):
        pre = _prealloc_backward(Q, x, A, need)
    dQ, dp, dA, db, ws = (pre[k] for k in ("dQ", "dp", "dA", "db", "ws"))
    dlb = pre["dlb"] if (any_bounds is None or any_bounds[0]) else None
    dub = pre["dub"] if (any_bounds is None or any_bounds[1]) else None
    ws_bytes = ws.numel()
    with _on_device(dev):
        stream = _raw_stream(dev)
        if finish and any_bounds is not None:
            rc = getattr(L, f"lqpb_backward_finish_{sfx}")(
                B, n, m, 1, _abi.ptr(g), _abi.ptr(x), None, _abi.ptr(lams), _abi.ptr(nus), _abi.ptr(Q), _abi.ptr(A),
                _abi.ptr(lb), _abi.ptr(ub), None, 0.0, _abi.ptr(dQ), _abi.ptr(dp), _abi.ptr(dA), _abi.ptr(db),
                _abi.ptr(dlb), _abi.ptr(dub), _abi.ptr(ws), ws_bytes, C.c_void_p(stream))
            _abi.check(rc, "lqpb_backward_finish")
            return dQ, dp, dA, db, dlb, dub
        rc = getattr(L, f"lqpb_backward_kkt_{sfx}")(
            B, n, m, _abi.ptr(g), _abi.ptr(x), _abi.ptr(lams), _abi.ptr(nus), _abi.ptr(Q), _abi.ptr(A), _abi.ptr(lb),
            _abi.ptr(ub), _abi.ptr(dQ), _abi.ptr(dp), _abi.ptr(dA), _abi.ptr(db), _abi.ptr(dlb), _abi.ptr(dub),
            flags if any_bounds is None else None, _abi.ptr(ws), ws_bytes, C.c_void_p(stream))
        _abi.check(rc, "lqpb_backward_kkt")
    if any_bounds is None:                      # reference :572-579
        if not flags[0]:
            dlb = None
        if not flags[1]:
            dub = None
    return dQ, dp, dA, db, dlb, dub


def _grad_host(dl_dz, x, u, lams, nus, Q, A, lb, ub, rho, need, kkt, any_bounds, prepared_ws=None):
    """Backward for host callers: saved tensors are the device copies of the forward, ``dl_dz`` is a CPU tensor and
    the gradients are returned as (pinned) CPU tensors through ``lqpb_backward_host_*``.  ``any_bounds`` is the
    (any_lb, any_ub) pair of the forward solve (KKT mode: dlb / dub are None without such bounds, :572-579)."""
    L = _abi.lib()
    dev, dt = x.device, x.dtype
    sfx = _abi.suffix(dt)
    B, n = Q.shape[0], Q.shape[1]
    m = get_ncon(A, dim=1)
    g = dl_dz.detach()
    if g.dtype != dt:
        raise TypeError(f"dl_dz has dtype {g.dtype}, expected {dt}")
    g = g.contiguous()
    rho_dev, rho_scalar = None, 0.0
    if not kkt:
        if rho is None:
            rho = 1.0                              # :356-357
        if torch.is_tensor(rho):
            rho_dev = rho.detach().to(device=dev, dtype=dt).reshape(-1).contiguous()
            if rho_dev.numel() == 1 and B > 1:
                rho_dev = rho_dev.expand(B).contiguous()
        else:
            rho_scalar = float(rho)
    nQ, np_, nA, nb, nlb, nub = need
    if kkt:
        nlb = nlb and any_bounds[0]
        nub = nub and any_bounds[1]
    shapes = [(B, n, n) if nQ else None, (B, n, 1) if np_ else None, (B, m, n) if (nA and m > 0) else None,
              (B, m, 1) if (nb and m > 0) else None, (B, n, 1) if nlb else None, (B, n, 1) if nub else None]
    with _on_device(dev):
        dbuf = [None if s is None else torch.empty(s, dtype=dt, device=dev) for s in shapes]
        hbuf = [None if s is None else torch.empty(s, dtype=dt, pin_memory=True) for s in shapes]
        g_dev = torch.empty((B, n, 1), dtype=dt, device=dev)
        g_src = g
        if _PREFETCH:
            # a batch is announced: dl_dz goes up FIRST (the H2D engine serves its queue in order), then the announced
            # copies start, and the C call below is handed the device copy of dl_dz.  With batches solved ahead the H2D
            # engine may still be busy with the batch announced a step ago: dl_dz is then read by a kernel straight from
            # page-locked memory (lqpb_copy_mapped) and never meets the engine's queue
            if any("control" in pf for pf in _PREFETCH.values()):
                gp = g if g.is_pinned() else torch.empty(g.shape, dtype=dt, pin_memory=True).copy_(g)
                _abi.check(L.lqpb_copy_mapped(_abi.ptr(g_dev), _abi.ptr(gp), gp.numel() * gp.element_size(),
                                              C.c_void_p(_raw_stream(dev))), "lqpb_copy_mapped")
                g_src = g_dev                        # "already in place" (gp stays referenced until the call below has synchronised)
            else:
                g_src = torch.empty((B, n, 1), dtype=dt, device=dev)
                g_src.copy_(g.reshape(B, n, 1), non_blocking=True)
            _launch_prefetches()
        ws_bytes = _ws_bytes("backward", sfx, B, n, m)
        ws = prepared_ws if prepared_ws is not None else torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        stream = _raw_stream(dev)
        rc = getattr(L, f"lqpb_backward_host_{sfx}")(
            B, n, m, 1 if kkt else 0, _abi.ptr(g_src), _abi.ptr(g_dev), _abi.ptr(x), _abi.ptr(u), _abi.ptr(lams),
            _abi.ptr(nus), _abi.ptr(Q), _abi.ptr(A), _abi.ptr(lb), _abi.ptr(ub), _abi.ptr(rho_dev), rho_scalar,
            *[_abi.ptr(t) for t in dbuf], *[_abi.ptr(t) for t in hbuf], None, _abi.ptr(ws), ws_bytes,
            C.c_void_p(stream), 0, 1 if prepared_ws is not None else 0)
        _abi.check(rc, "lqpb_backward_host")
    return tuple(hbuf)
