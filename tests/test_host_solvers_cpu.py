"""CPU-side checks of the solver / GP host layer (`cg.py`, `low_rank.py`, `gp.py`, `pde.py`): factory
signatures, parameter dictionaries, the soft-plus constraint against the reference's golden values, error
conventions that fire before any device work, workspace queries.  No compute calls (no GPU here)."""

import numpy as np
import pytest
from conftest import golden

import experiments_lanczos_adjoints_b200 as bl
from experiments_lanczos_adjoints_b200 import _lib, cg, gp, low_rank, pde
from oracle import operators as oracle_ops


def test_constraint_matches_reference_softplus_golden():
    g = golden("gp_kernels_n40_d3_f64")  # gp_util.constraint_greater_than(0.0) on a fixed grid, from the reference
    c = gp.constraint_greater_than(0.0)
    assert np.allclose(c(g["softplus_x"]), g["softplus_y"], rtol=1e-12, atol=0)
    x = np.linspace(-25, 25, 101)
    shifted = gp.constraint_greater_than(1e-4)
    assert np.allclose(shifted(x), 1e-4 + oracle_ops.softplus(x), rtol=1e-9, atol=1e-15)
    h = 1e-5
    xs = np.concatenate([np.linspace(-10, 19, 30), np.linspace(21, 25, 5)])  # away from the switch at 20
    assert np.allclose(shifted.grad(xs), (shifted(xs + h) - shifted(xs - h)) / (2 * h), atol=1e-6)
    assert np.allclose(shifted.grad(x), oracle_ops.softplus_grad(x))


def test_gp_factories_mirror_the_reference_parameter_dictionaries():
    # gp_util.kernel_scaled_matern_32(shape_in=(d,), shape_out=()) -> (parametrize, {"raw_lengthscale", "raw_outputscale"})
    for make in (gp.kernel_scaled_matern_32, gp.kernel_scaled_matern_12, gp.kernel_scaled_rbf):
        k, p = make(shape_in=(9,), shape_out=())
        assert p["raw_lengthscale"].shape == (9,) and p["raw_outputscale"].shape == ()
        kern = k(raw_lengthscale=np.ones(9), raw_outputscale=0.5)
        assert kern.kind in ("matern32", "matern12", "rbf")
    m, p = gp.mean_constant(shape_out=())
    assert p["constant_value"].shape == ()
    prior = gp.model_gp(m, gp.kernel_scaled_rbf(shape_in=(2,))[0])
    mean, kernel = prior(params_mean={"constant_value": 0.3}, params_kernel={"raw_lengthscale": np.zeros(2), "raw_outputscale": 0.0})
    assert mean.constant_value == 0.3 and kernel.kind == "rbf"
    lik, p = gp.likelihood_pdf_p(gp.gram_matvec(), gp.logpdf_krylov_p(solve_p=None, logdet=None), precondition=None,
                                 constrain=gp.constraint_greater_than(1e-4))  # fmt: skip
    assert set(p) == {"raw_noise"} and p["raw_noise"].shape == ()
    assert callable(gp.target_logml(prior, lik)) and hasattr(gp.target_logml(prior, lik), "value_and_grad")
    assert gp.gram_matvec_partitioned(4, checkpoint=True) == gp.gram_matvec()


def test_solver_factories_and_error_conventions_before_device_work():
    # cg.py: factories take the reference's arguments
    assert cg.pcg_fixed_step(7).max_steps == 7 and cg.pcg_fixed_step(7).atol < 0
    s = cg.pcg_adaptive(atol=1e-2, rtol=0.0, maxiter=1000, miniter=10)
    assert (s.max_steps, s.min_steps, s.atol, s.rtol) == (1000, 10, 1e-2, 0.0)
    with pytest.raises(TypeError):
        cg.pcg_adaptive(atol=1e-2, rtol=0.0, maxiter=10)  # miniter is required, as in cg.py:75
    with pytest.raises(TypeError, match="operator object"):
        cg.cg_fixed_step(3)(lambda v: v, np.ones(4))  # closures cannot run on the device
    # low_rank.py:67-72 / 124-129: the rank checks and their messages come before anything else
    op = bl.operators.bound(bl.operators.DenseOperator(5), np.eye(5))
    with pytest.raises(ValueError, match="Rank exceeds n: 6 >= 5."):
        low_rank.cholesky_partial_pivot(rank=6)(op, 5)
    with pytest.raises(ValueError, match="Rank must be positive, but 0 < 1."):
        low_rank.cholesky_partial(rank=0)(op, 5)
    with pytest.raises(TypeError):
        low_rank.cholesky_partial(rank=2)(lambda i, j: 1.0, 5)


def test_pde_factories_and_workspace_queries():
    st = pde.stencil_laplacian(0.5)
    assert st.shape == (3, 3) and st[1, 1] == -2.0 / 0.25 and st[0, 1] == 1.0 / 0.25  # pde_util.py:18-20
    with pytest.raises(NotImplementedError):
        pde.pde_wave_anisotropic(np.ones((4, 4)), st, constrain="exp", boundary="neumann")
    with pytest.raises(ValueError):
        pde.pde_wave_anisotropic(np.ones((4, 5)), st)
    e = pde.expm_arnoldi(6, max_squarings=8, reortho="full", custom_vjp=True)
    assert e.K == 6 and e.kwargs == {"reortho": "full", "custom_vjp": True}
    lib = _lib.load()
    assert lib.bl_pcg_workspace_bytes(1000, _lib.BL_F32) < lib.bl_pcg_workspace_bytes(1_000_000, _lib.BL_F64)
    assert lib.bl_cholesky_workspace_bytes(1000, 10, _lib.BL_F32) < lib.bl_cholesky_workspace_bytes(100_000, 100, _lib.BL_F64)
