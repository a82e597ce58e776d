"""Small helpers shared by the layer code."""


def get_ncon(x, dim=0):
    """Number of constraints: size of ``x`` along ``dim``, 0 for ``None``
    (reference lqp_py/utils.py:14-20)."""
    return 0 if x is None else x.shape[dim]
