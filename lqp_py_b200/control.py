"""Settings factory of the ADMM box-QP layer.

Mirror of the reference's ``box_qp_control`` (lqp_py/control.py:1-24): same keyword
arguments, same defaults, same resulting dict -- including its two historical quirks,
because downstream code indexes the dict by these exact names:

* ``check_solved`` is stored under the misspelt key ``'check_terimnation'`` (control.py:8),
  which the solver never reads (it reads ``'check_solved'``, solve_box_qp_admm_torch.py:139);
* ``adaptive_rho_max_iter`` is stored under that name (control.py:15) while the solver reads
  ``'adaptive_max_iter'`` (:148).

Unknown keyword arguments are merged into the dict untouched (control.py:23), e.g. the
``reduce='max'`` every experiment passes.
"""

_FIELDS = (
    # (keyword argument, key in the dict, default)
    ("max_iters", "max_iters", 10_000),
    ("eps_abs", "eps_abs", 1e-3),
    ("eps_rel", "eps_rel", 1e-3),
    ("check_solved", "check_terimnation", None),
    ("rho", "rho", None),
    ("rho_min", "rho_min", 1e-6),
    ("rho_max", "rho_max", 1e6),
    ("adaptive_rho", "adaptive_rho", True),
    ("adaptive_rho_tol", "adaptive_rho_tol", 10),
    ("adaptive_rho_iter", "adaptive_rho_iter", 100),
    ("adaptive_rho_max_iter", "adaptive_rho_max_iter", 1000),
    ("adaptive_rho_threshold", "adaptive_rho_threshold", 1e-5),
    ("verbose", "verbose", False),
    ("scale", "scale", True),
    ("unroll", "unroll", False),
    ("beta", "beta", None),
    ("backward", "backward", "fixed_point"),
)
_POSITIONAL = ("max_iters", "eps_abs", "eps_rel", "check_solved", "rho", "rho_min", "rho_max", "adaptive_rho",
               "adaptive_rho_tol", "adaptive_rho_iter", "adaptive_rho_max_iter", "adaptive_rho_threshold",
               "verbose", "scale", "beta", "unroll", "backward")


def box_qp_control(*args, **kwargs):
    """Build the control dict consumed by :class:`SolveBoxQP`.

    Accepts the reference's arguments positionally (in its order) or by keyword."""
    if len(args) > len(_POSITIONAL):
        raise TypeError(f"box_qp_control() takes at most {len(_POSITIONAL)} positional arguments")
    given = dict(zip(_POSITIONAL, args))
    for name in given:
        if name in kwargs:
            raise TypeError(f"box_qp_control() got multiple values for argument '{name}'")
    control = {}
    for name, key, default in _FIELDS:
        if name in given:
            control[key] = given[name]
        elif name in kwargs:
            control[key] = kwargs.pop(name)
        else:
            control[key] = default
    control.update(**kwargs)
    return control
