"""CPU checks of the torch-side glue of the unrolled mode (the parts that are plain torch and need no GPU)."""
import torch

from lqp_py_b200.solve_box_qp_admm_torch import _ColumnMax


def test_column_max_node_matches_torch_inf_norm():
    """_ColumnMax = norm(Q, inf, dim=1) (reference :163) with a sparse adjoint: same values, same gradient as torch's
    dense backward whenever the column maximiser is unique."""
    gen = torch.Generator().manual_seed(3)
    Q = torch.randn(4, 9, 9, generator=gen, dtype=torch.float64, requires_grad=True)
    g = torch.randn(4, 9, generator=gen, dtype=torch.float64)
    ref_val = torch.linalg.norm(Q, ord=float("inf"), dim=1)
    ref_grad = torch.autograd.grad(ref_val, Q, g)[0]
    val = _ColumnMax.apply(Q)
    grad = torch.autograd.grad(val, Q, g)[0]
    assert torch.equal(val, ref_val)
    assert torch.equal(grad, ref_grad)


def test_column_max_node_zero_column_carries_no_gradient():
    Q = torch.zeros(1, 3, 3, dtype=torch.float64)
    Q[0, :, 0] = torch.tensor([1.0, -2.0, 0.5])
    Q.requires_grad_(True)
    val = _ColumnMax.apply(Q)
    grad = torch.autograd.grad(val, Q, torch.ones(1, 3, dtype=torch.float64))[0]
    assert val.tolist() == [[2.0, 0.0, 0.0]]
    assert grad[0, 1, 0].item() == -1.0 and float(grad.abs().sum()) == 1.0


def test_column_max_node_splits_gradient_among_ties_like_torch():
    """Exact ties (equicorrelation / constant-block Q, |Q_ij| == Q_jj): torch's inf-norm backward splits the gradient
    evenly among the maximisers; the node does the same, so dQ in unroll + scale mode matches the reference's."""
    Q = torch.full((2, 4, 4), 0.5, dtype=torch.float64)
    Q[0] += 0.5 * torch.eye(4, dtype=torch.float64)            # unique maximiser (the diagonal) in problem 0
    Q[1, 2, :] = -0.5                                           # problem 1: every column all-tied, mixed signs
    Q[1, :, 3] = torch.tensor([0.1, -0.7, 0.7, 0.2])            # a two-way tie with opposite signs
    Q.requires_grad_(True)
    g = torch.tensor([[1.0, 2.0, 3.0, 4.0], [5.0, 6.0, 7.0, 8.0]], dtype=torch.float64)
    ref_val = torch.linalg.norm(Q, ord=float("inf"), dim=1)
    ref_grad = torch.autograd.grad(ref_val, Q, g)[0]
    val = _ColumnMax.apply(Q)
    grad = torch.autograd.grad(val, Q, g)[0]
    assert torch.equal(val, ref_val)
    assert torch.allclose(grad, ref_grad, rtol=0, atol=1e-15)
