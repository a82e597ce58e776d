"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/lqpb.h declares, the ctypes structs match the header, the Python adapter derives the
settings exactly like the reference (via the oracle and the reference's own control dicts stored in
the golden fixtures), and no compute happens without a GPU (no CPU fallback)."""
import ctypes
import json
import os
import re

import pytest
import torch

import __graft_entry__ as entry
from oracle import box_qp_oracle as orc
from tests._golden import Case, case_names

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()          # nvcc cross-compiles without a GPU; no-op when _lqpb.so is current
    from lqp_py_b200 import _abi
    return _abi.lib()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "lqpb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lqpb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    from lqp_py_b200 import _abi
    declared = _header_functions()
    assert declared, "no functions parsed from include/lqpb.h"
    assert sorted(_abi.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), f"_lqpb.so does not export {name}"
    assert lib.lqpb_abi_version() == 1


def test_header_is_plain_c(tmp_path):
    """include/lqpb.h is the C ABI: it must compile as C (no C++ in the boundary) and the struct sizes the ctypes
    binding assumes must be the compiler's."""
    import subprocess
    src = tmp_path / "abi_check.c"
    src.write_text('#include "lqpb.h"\n#include <stdio.h>\n'
                   'int main(void) { printf("%zu %zu %zu\\n", sizeof(lqpb_config), sizeof(lqpb_info), sizeof(lqpb_profile));'
                   ' return 0; }\n')
    exe = tmp_path / "abi_check"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    from lqp_py_b200 import _abi
    assert sizes == [ctypes.sizeof(_abi.Config), ctypes.sizeof(_abi.Info), ctypes.sizeof(_abi.Profile)]


def test_struct_layouts_match_header(lib):
    from lqp_py_b200 import _abi
    assert ctypes.sizeof(_abi.Config) == 10 * 4 + 9 * 8
    assert ctypes.sizeof(_abi.Info) == 6 * 4 + 64 * 4 + 2 * 64 * 8
    assert ctypes.sizeof(_abi.Profile) == 7 * 4 + 3 * 4


def test_workspace_sizes(lib):
    for sfx, s in (("f32", 4), ("f64", 8)):
        f = getattr(lib, f"lqpb_forward_workspace_bytes_{sfx}")
        g = getattr(lib, f"lqpb_backward_workspace_bytes_{sfx}")
        small, big = f(1, 10, 0), f(128, 500, 1)
        assert 0 < small < big
        # at least the packed lower triangles of Q~ and K11 (136 resp. 272 tiles of 4 KB at n = 500)
        # and the Gauss-Jordan work matrix
        tiles = 136 if sfx == "f32" else 272
        assert big >= 128 * (2 * tiles * 4096 + s * 512 * 512)
        assert big < 128 * s * (2 * 500 * 500 + 512 * 512)        # ... and less than two full matrices
        # backward: the LDL^T work matrix (no explicit inverse is formed)
        assert g(128, 500, 1) >= 128 * s * 512 * 512


def test_null_and_bad_arguments_are_rejected_without_a_gpu(lib):
    from lqp_py_b200 import _abi
    cfg, info = _abi.Config(), _abi.Info()
    rc = lib.lqpb_forward_f32(ctypes.byref(cfg), 1, 4, 0, None, None, None, None, None, None, None, None, None,
                              None, None, None, ctypes.byref(info), None, 0, None)
    assert rc == 1 and b"null" in lib.lqpb_last_error()
    rc = lib.lqpb_lu_factor_f64(0, 4, None, None, None, None)
    assert rc == 1


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import SolveBoxQP, torch_solve_box_qp
    Q, p, A, b, lb, ub = orc.make_exp1_data(8, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SolveBoxQP(box_qp_control()).forward(Q, p, A, b, lb, ub)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        torch_solve_box_qp(Q, p, A, b, lb, ub, box_qp_control())


@pytest.mark.parametrize("name", case_names())
def test_control_factory_matches_reference_dicts(name):
    """The golden fixtures store the dict produced by the reference's own box_qp_control."""
    from lqp_py_b200.control import box_qp_control
    case = Case(name)
    kw = json.loads(str(case.z["control_kw"]))
    mine = box_qp_control(**kw)
    mine.update(json.loads(str(case.z["extra_control"])))
    ref = case.control_dict()
    if name.startswith("unbounded"):
        mine["rho"] = 0            # set by SolveBoxQPLayer.forward (:37-38) before the fixture was written
    assert mine == ref and list(mine.keys()) == list(ref.keys())


@pytest.mark.parametrize("n", [1, 10, 37, 50, 100, 250, 500, 1000, 2500])
def test_derived_settings_match_oracle(n):
    from lqp_py_b200.control import box_qp_control
    from lqp_py_b200.solve_box_qp_admm_torch import _derive_config
    variants = [box_qp_control(), box_qp_control(eps_abs=1e-5, eps_rel=1e-14, rho=2.5, beta=0.25, scale=False),
                box_qp_control(adaptive_rho=False, adaptive_rho_iter=35, check_solved=3, max_iters=17),
                {"check_solved": 7, "adaptive_max_iter": 300}, {}]
    for c in variants:
        st, cfg = orc.derive_settings(c, n), _derive_config(c, n)
        assert (cfg.max_iters, cfg.check_solved, cfg.adaptive_rho_iter, cfg.adaptive_rho_max_iter) == \
               (st.max_iters, st.check_every, st.adaptive_every, st.adaptive_until)
        assert (cfg.eps_abs, cfg.eps_rel, cfg.rho_min, cfg.rho_max) == (st.eps_abs, st.eps_rel, st.rho_min, st.rho_max)
        assert bool(cfg.adaptive_rho) == bool(st.adaptive) and bool(cfg.scale) == bool(st.scale)
        assert bool(cfg.rho_auto) == (st.rho is None) and bool(cfg.beta_auto) == (st.beta is None)
        assert cfg.adaptive_rho_tol == st.adaptive_tol
        if st.rho is not None:
            assert cfg.rho == st.rho


def test_datasets_match_oracle_generators():
    from lqp_py_b200.datasets import create_qp_data, generate_hard_qp_torch
    a = create_qp_data(12, 3, 24, seed=5)[:6]
    b = orc.make_exp1_data(12, 3, seed=5)
    assert all(torch.equal(x.detach(), y) for x, y in zip(a, b))
    a = generate_hard_qp_torch(16, 0.5, [3, 4])[:6]
    b = orc.make_hard_data(16, 0.5, [3, 4])
    assert all(torch.equal(x.detach(), y) for x, y in zip(a, b))
