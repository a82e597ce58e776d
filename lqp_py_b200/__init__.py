"""lqp_py_b200 -- B200-native (sm_100a) batched ADMM box-QP layer, drop-in for the ADMM path of
ipo-lab/lqp_py (``SolveBoxQP`` forward solve + implicit fixed-point backward + ``box_qp_control``).

Sub-modules keep the reference's names so imports translate one-to-one:

    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP
    from lqp_py_b200.control import box_qp_control
"""
from .control import box_qp_control
from .utils import get_ncon

__all__ = ["box_qp_control", "get_ncon", "SolveBoxQP", "SolveBoxQPLayer", "BoxQPTH", "torch_solve_box_qp",
           "torch_solve_box_qp_grad", "torch_solve_box_qp_grad_kkt", "TorchLU", "TorchLULayer"]


def __getattr__(name):
    # torch is imported lazily so that `import lqp_py_b200.build` works in a bare interpreter
    if name in ("SolveBoxQP", "SolveBoxQPLayer", "BoxQPTH", "torch_solve_box_qp", "torch_solve_box_qp_grad",
                "torch_solve_box_qp_grad_kkt"):
        from . import solve_box_qp_admm_torch as m
        return getattr(m, name)
    if name in ("TorchLU", "TorchLULayer"):
        from . import lu_layer as m
        return getattr(m, name)
    raise AttributeError(name)
