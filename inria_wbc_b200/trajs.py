"""Minimum-jerk reference generators (host side, closed form).

Restates include/inria_wbc/trajs/trajectory_generator.hpp:29-147 of the reference (the polynomials
6t^5 - 15t^4 + 10t^3 and their derivatives; SE3 version interpolates the angle of the relative
rotation about a fixed axis) with numpy, producing whole tables at once.  Used by the re-hosted
behaviors (behaviors.py) and the Franka known-answer test.
"""
import numpy as np


def _poly(td, order, duration):
    if order == 0:
        return 6 * td ** 5 - 15 * td ** 4 + 10 * td ** 3
    if order == 1:
        return (30 * td ** 4 - 60 * td ** 3 + 30 * td ** 2) / duration
    return (120 * td ** 3 - 180 * td ** 2 + 60 * td) / duration ** 2


def n_steps(dt, duration):
    return int(np.floor(duration / dt))


def min_jerk_vec(x0, xf, dt, duration, order=0):
    """[n_steps, dim] table of minimum_jerk_polynom<order>(x0, xf, dt*i, duration)."""
    x0 = np.asarray(x0, dtype=np.float64)
    xf = np.asarray(xf, dtype=np.float64)
    td = (dt * np.arange(n_steps(dt, duration))) / duration
    s = _poly(td, order, duration)[:, None]
    return (x0[None, :] if order == 0 else 0.0) + (xf - x0)[None, :] * s


def _angle_axis(R):
    """Eigen::AngleAxisd(R): through the quaternion, angle in [0, pi]."""
    tr = np.trace(R)
    if tr > 0:
        t = np.sqrt(tr + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        x, y, z = (R[2, 1] - R[1, 2]) * t, (R[0, 2] - R[2, 0]) * t, (R[1, 0] - R[0, 1]) * t
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        qv = np.zeros(3)
        qv[i] = 0.5 * t
        t = 0.5 / t
        w = (R[k, j] - R[j, k]) * t
        qv[j] = (R[j, i] + R[i, j]) * t
        qv[k] = (R[k, i] + R[i, k]) * t
        x, y, z = qv
    n = np.sqrt(x * x + y * y + z * z)
    if n == 0:
        return 0.0, np.array([1.0, 0.0, 0.0])
    angle = 2 * np.arctan2(n, abs(w))
    if w < 0:
        n = -n
    return angle, np.array([x, y, z]) / n


def _rot(axis, angle):
    a = np.asarray(axis, dtype=np.float64)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


def min_jerk_se3(p0, R0, pf, Rf, dt, duration):
    """Returns (pos[n,3], rot[n,3,3], vel[n,6], acc[n,6]) as the reference's three SE3 overloads do."""
    n = n_steps(dt, duration)
    angle, axis = _angle_axis(R0.T @ Rf)
    pos = min_jerk_vec(p0, pf, dt, duration, 0)
    ang = min_jerk_vec([0.0], [angle], dt, duration, 0)[:, 0]
    rot = np.stack([R0 @ _rot(axis, a) for a in ang])
    out = [pos, rot]
    for order in (1, 2):
        d = np.zeros((n, 6))
        d[:, :3] = min_jerk_vec(p0, pf, dt, duration, order)
        a = min_jerk_vec([0.0], [angle], dt, duration, order)[:, 0]
        d[:, 3:] = (R0 @ axis)[None, :] * a[:, None]
        out.append(d)
    return tuple(out)


def se3_sample(p, R, vel=None, acc=None):
    """TrajectorySample(12,6) layout of the reference block: p(3), R column-major(9), vel(6), acc(6)."""
    s = np.zeros(24)
    s[:3] = p
    s[3:12] = np.asarray(R).T.reshape(9)
    if vel is not None:
        s[12:18] = vel
    if acc is not None:
        s[18:24] = acc
    return s


def sample_se3(s):
    return np.array(s[:3]), np.array(s[3:12]).reshape(3, 3).T


def rpy_matrix(rpy):
    """AngleAxis(rz, Z) * AngleAxis(ry, Y) * AngleAxis(rx, X) (src/behaviors/generic/cartesian.cpp:36-38)"""
    return _rot([0, 0, 1], rpy[2]) @ _rot([0, 1, 0], rpy[1]) @ _rot([1, 0, 0], rpy[0])
