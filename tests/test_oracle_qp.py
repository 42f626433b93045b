"""The oracle's Goldfarb-Idnani solver against KKT conditions and SciPy (independent implementation)."""
import numpy as np
import pytest
from scipy.optimize import minimize

import oracle_lib as ol


def random_qp(rng, n, me, mi, feasible=True):
    A = rng.normal(size=(n + 3, n))
    H = A.T @ A + 1e-3 * np.eye(n)
    g = rng.normal(size=n)
    x0 = rng.normal(size=n)
    CE = rng.normal(size=(me, n))
    ce0 = -CE @ x0
    CI = rng.normal(size=(mi, n))
    slack = rng.uniform(-0.5, 1.0, mi) if feasible else rng.uniform(-3, -1, mi)
    ci0 = -CI @ x0 + np.where(slack > 0, slack, 0.0)   # x0 is feasible
    return H, g, CE, ce0, CI, ci0


def kkt_residuals(H, g, CE, ce0, CI, ci0, x, lam):
    me = len(ce0)
    le, li = lam[:me], lam[me:]
    stat = H @ x + g - (CE.T @ le if me else 0) - CI.T @ li
    prim_e = CE @ x + ce0 if me else np.zeros(1)
    s = CI @ x + ci0
    return np.abs(stat).max(), np.abs(prim_e).max(), min(s.min(), 0.0), min(li.min(), 0.0), np.abs(li * s).max()


@pytest.mark.parametrize("n,me,mi", [(9, 0, 0), (12, 0, 24), (20, 5, 30), (40, 12, 80), (74, 18, 244)])
def test_kkt_and_scipy(n, me, mi):
    rng = np.random.default_rng(n * 1000 + mi)
    for trial in range(3):
        H, g, CE, ce0, CI, ci0 = random_qp(rng, n, me, max(mi, 1))
        if mi == 0:
            CI, ci0 = np.zeros((1, n)), np.ones(1)
        r = ol.qp_solve(H, g, CE if me else None, ce0 if me else None, CI, ci0)
        assert r["status"] == 0
        scale = max(1.0, np.abs(H).max() * np.abs(r["x"]).max())
        stat, pe, pi, dneg, comp = kkt_residuals(H, g, CE, ce0, CI, ci0, r["x"], r["lam"])
        assert stat < 1e-9 * scale and pe < 1e-9 * scale and pi > -1e-9 * scale and dneg >= -1e-12 and comp < 1e-8 * scale
        if n <= 40:
            cons = [{"type": "ineq", "fun": lambda x: CI @ x + ci0, "jac": lambda x: CI}]
            if me:
                cons.append({"type": "eq", "fun": lambda x: CE @ x + ce0, "jac": lambda x: CE})
            s = minimize(lambda x: 0.5 * x @ H @ x + g @ x, np.zeros(n), jac=lambda x: H @ x + g, constraints=cons, method="SLSQP",
                         options={"ftol": 1e-14, "maxiter": 500})
            f_or = 0.5 * r["x"] @ H @ r["x"] + g @ r["x"]
            assert f_or <= s.fun + 1e-7 * max(1, abs(s.fun))
            np.testing.assert_allclose(r["x"], s.x, atol=2e-5 * max(1, np.abs(s.x).max()))


def test_unconstrained_is_cholesky_solve():
    rng = np.random.default_rng(0)
    H, g, *_ = random_qp(rng, 9, 0, 1)
    r = ol.qp_solve(H, g, None, None, None, None)
    np.testing.assert_allclose(r["x"], -np.linalg.solve(H, g), rtol=1e-12)
    assert r["status"] == 0 and r["active"] == []


def test_infeasible_and_redundant_statuses():
    n = 6
    H = np.eye(n)
    g = np.zeros(n)
    # x0 >= 1 and x0 <= -1 : infeasible -> eiquadprog UNBOUNDED(dual) -> HQP infeasible (1)
    CI = np.zeros((2, n))
    CI[0, 0], CI[1, 0] = 1, -1
    r = ol.qp_solve(H, g, None, None, CI, np.array([-1.0, -1.0]))
    assert r["status"] == 1
    # two identical equalities -> REDUNDANT_EQUALITIES -> HQP error (4)
    CE = np.zeros((2, n))
    CE[:, 0] = 1.0
    r = ol.qp_solve(H, g, CE, np.array([1.0, 1.0]), None, None)
    assert r["status"] == 4


def test_first_index_wins_on_ties():
    # two identical violated constraints: the active set must hold the first one only
    n = 4
    H = np.eye(n)
    g = np.zeros(n)
    CI = np.zeros((3, n))
    CI[0, 1] = CI[1, 1] = 1.0
    CI[2, 2] = 1.0
    r = ol.qp_solve(H, g, None, None, CI, np.array([-1.0, -1.0, 0.5]))
    assert r["status"] == 0 and r["active"] == [0]
    np.testing.assert_allclose(r["x"], [0, 1, 0, 0], atol=1e-14)
