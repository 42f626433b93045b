"""Algorithm-independent identities that pin the oracle's rigid-body dynamics (SURVEY.md hard part H1)."""
import numpy as np
import pytest

import oracle_lib as ol
from conftest import random_states


def _integrate(P, q, dv, h):
    """q (+) h*dv with the oracle's own free-flyer convention (body-frame twist), first order in h."""
    q2 = q.copy()
    na = P.na
    if P.nv > na:
        x, y, z, w = q[3:7]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        q2[:3] = q[:3] + R @ (h * dv[:3])
        wv = h * dv[3:6]
        dq = 0.5 * np.array([w * wv[0] + y * wv[2] - z * wv[1], w * wv[1] + z * wv[0] - x * wv[2], w * wv[2] + x * wv[1] - y * wv[0],
                             -x * wv[0] - y * wv[1] - z * wv[2]])
        qq = q[3:7] + dq
        q2[3:7] = qq / np.linalg.norm(qq)
    q2[P.nq - na:] = q[P.nq - na:] + h * dv[P.nv - na:]
    return q2


@pytest.mark.parametrize("name", ["franka", "talos", "tiago"])
def test_mass_matrix_spd_and_rnea_consistency(name, request):
    P = request.getfixturevalue(name)
    q, v = random_states(P, 4, 11, qn=0.3, vn=0.5, base=True)
    rng = np.random.default_rng(5)
    for i in range(4):
        d = ol.cycle_debug(P, q[i], v[i])
        M = d["M"]
        assert np.abs(M - M.T).max() == 0.0
        assert np.linalg.eigvalsh(M).min() > 0
        a = rng.normal(size=P.nv)
        tau = ol.rnea(P, q[i], v[i], a)           # independent recursion (no CRBA)
        np.testing.assert_allclose(M @ a + d["nle"], tau, rtol=0, atol=1e-10 * max(1, np.abs(tau).max()))


def test_gravity_term_matches_potential_energy_gradient(talos):
    P = talos
    q, _ = random_states(P, 1, 3, qn=0.2, vn=0.0)
    q = q[0]
    z = np.zeros(P.nv)
    g = ol.cycle_debug(P, q, z)["nle"]
    m = P.struct().total_mass
    h = 1e-6
    for k in (0, 2, 4, 9, 20, 33, 49):
        e = np.zeros(P.nv)
        e[k] = 1
        cp = ol.cycle_debug(P, _integrate(P, q, e, h), z)["com"][2]
        cm = ol.cycle_debug(P, _integrate(P, q, e, -h), z)["com"][2]
        dU = m * 9.81 * (cp - cm) / (2 * h)
        assert abs(g[k] - dU) < 1e-5 * max(1.0, abs(dU)), (k, g[k], dU)


@pytest.mark.parametrize("name", ["franka", "talos"])
def test_jacobians_by_finite_differences(name, request):
    P = request.getfixturevalue(name)
    q, v = random_states(P, 1, 21, qn=0.2, vn=0.3, base=True)
    q, v = q[0], v[0]
    d = ol.cycle_debug(P, q, v)
    h = 1e-6
    # CoM Jacobian and CoM velocity
    Jfd = np.zeros((3, P.nv))
    for k in range(P.nv):
        e = np.zeros(P.nv)
        e[k] = 1
        Jfd[:, k] = (ol.cycle_debug(P, _integrate(P, q, e, h), v)["com"][:3] - ol.cycle_debug(P, _integrate(P, q, e, -h), v)["com"][:3]) / (2 * h)
    np.testing.assert_allclose(d["Jcom"], Jfd, atol=2e-8)
    np.testing.assert_allclose(d["Jcom"] @ v, d["com"][3:6], atol=1e-12)
    np.testing.assert_allclose(d["Ag"][:3] / P.struct().total_mass, d["Jcom"], atol=1e-12)
    # frame 0 (first task frame): local Jacobian vs finite differences of the placement, local velocity = J v
    f0 = ol.frame(P, q, v, 0)
    R = f0["oMf"][:9].reshape(3, 3)
    p = f0["oMf"][9:]
    for k in range(0, P.nv, max(1, P.nv // 9)):
        e = np.zeros(P.nv)
        e[k] = 1
        fp = ol.frame(P, _integrate(P, q, e, h), v, 0)["oMf"]
        fm = ol.frame(P, _integrate(P, q, e, -h), v, 0)["oMf"]
        lin = R.T @ ((fp[9:] - fm[9:]) / (2 * h))
        dR = (fp[:9].reshape(3, 3) - fm[:9].reshape(3, 3)) / (2 * h)
        W = R.T @ dR
        ang = np.array([W[2, 1], W[0, 2], W[1, 0]])
        np.testing.assert_allclose(f0["J"][:3, k], lin, atol=2e-8)
        np.testing.assert_allclose(f0["J"][3:, k], ang, atol=2e-8)
    np.testing.assert_allclose(f0["J"] @ v, f0["vel"], atol=1e-12)
    assert np.linalg.norm(p) > 0


def test_centroidal_momentum_and_its_drift(talos):
    P = talos
    q, v = random_states(P, 1, 8, qn=0.1, vn=0.4, base=True)
    q, v = q[0], v[0]
    d = ol.cycle_debug(P, q, v)
    np.testing.assert_allclose(d["Ag"] @ v, d["hg"][:6], atol=1e-11)
    # drift = d/dt (Ag v) at zero joint acceleration: finite difference along the flow of v
    h = 1e-6
    hp = ol.cycle_debug(P, _integrate(P, q, v, h), v)
    hm = ol.cycle_debug(P, _integrate(P, q, v, -h), v)
    fd = (hp["hg"][:6] - hm["hg"][:6]) / (2 * h)
    np.testing.assert_allclose(d["hg"][6:], fd, atol=5e-6 * max(1, np.abs(fd).max()))
    acom = (hp["com"][3:6] - hm["com"][3:6]) / (2 * h)
    np.testing.assert_allclose(d["com"][6:9], acom, atol=5e-6)


def test_frame_classic_acceleration_by_finite_differences(talos):
    P = talos
    q, v = random_states(P, 1, 9, qn=0.1, vn=0.5, base=True)
    q, v = q[0], v[0]
    f = ol.frame(P, q, v, 3)
    R = f["oMf"][:9].reshape(3, 3)
    h = 1e-6
    fp = ol.frame(P, _integrate(P, q, v, h), v, 3)
    fm = ol.frame(P, _integrate(P, q, v, -h), v, 3)
    # classic acceleration (local axes) = R^T d/dt (R v_lin_local), angular: R^T d/dt (R w_local)
    vw = lambda ff: (ff["oMf"][:9].reshape(3, 3) @ ff["vel"][:3], ff["oMf"][:9].reshape(3, 3) @ ff["vel"][3:])
    lp, ap = vw(fp)
    lm, am = vw(fm)
    np.testing.assert_allclose(f["acc"][:3], R.T @ ((lp - lm) / (2 * h)), atol=1e-5)
    np.testing.assert_allclose(f["acc"][3:], R.T @ ((ap - am) / (2 * h)), atol=1e-5)
