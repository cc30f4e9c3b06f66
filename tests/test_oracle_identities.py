"""A second, independent pin of the oracle (the first is the reference itself: tests/test_ref_pin.py against the
reference's own translation units, frozen in tests/golden/): mathematical identities that do not depend on the
restatement:
  * M symmetric, PSD; dead dofs (root slot 6, Q1) exactly zero
  * RNEA(q, qd, qdd) = M qdd + C           (CRBA and RNEA agree)
  * 1/2 qd^T M qd = kinetic energy from finite-differenced FK (independent numpy evaluation)
  * C(q, 0) = dV/dq from finite-differenced potential energy
  * pivoted LDLT solves A x = b and zeroes null components
"""
import numpy as np
import pytest

from conftest import CONFIG_NAMES
from terrainrlsim_b200 import synth


def _inputs(name, E=6):
    return synth.make_inputs(name, E)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_mass_matrix_symmetric_psd_and_dead_dofs(oracle, name):
    m = oracle.Model(name)
    x = _inputs(name)
    for e in range(x["pose"].shape[0]):
        M, C = m.rbd_update(x["pose"][e], x["vel"][e])
        assert np.allclose(M, M.T, rtol=0, atol=1e-12 * np.abs(M).max())
        assert np.all(M[6, :] == 0) and np.all(M[:, 6] == 0)
        dead = [6] + [int(m.offset[j]) + 3 for j in range(1, m.N) if m.jtype[j] == 4]
        live = [i for i in range(m.n) if i not in dead]
        for i in dead:
            assert np.all(M[i, :] == 0)
            assert C[i] == 0
        w = np.linalg.eigvalsh(M[np.ix_(live, live)])
        assert w.min() > 0


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_rnea_equals_mass_times_acc_plus_bias(oracle, name):
    m = oracle.Model(name)
    x = _inputs(name)
    rng = np.random.default_rng(1)
    for e in range(x["pose"].shape[0]):
        pose, vel = x["pose"][e], x["vel"][e]
        M, C = m.rbd_update(pose, vel)
        acc = rng.normal(size=m.n)
        tau = m.inv_dyna(pose, vel, acc)
        ref = M @ acc + C
        assert np.allclose(tau, ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max())


def _integrate(m, pose, qd, eps):
    """pose (+) eps*qd using the reference's own quaternion-rate map (VelToPoseDiff) + renormalise."""
    d = m.vel_to_pose_diff(pose, qd)
    return m.post_process_pose(pose + eps * d)


def _body_frames(m, pose):
    out = []
    for j in range(m.N):
        T = m.joint_world_trans(pose, j)
        out.append(T)
    return out


def _energies(m, pose, qd, eps=1e-6):
    """Kinetic energy via central finite differences of body frames; potential energy m g y."""
    p1, p0 = _integrate(m, pose, qd, eps), _integrate(m, pose, qd, -eps)
    F1, F0, Fc = _body_frames(m, p1), _body_frames(m, p0), _body_frames(m, pose)
    T = 0.0
    V = 0.0
    for j in range(m.N):
        I = m.inertia_spatial(j)  # about the joint origin, joint-frame axes
        R = Fc[j][:3, :3]
        v_origin = (F1[j][:3, 3] - F0[j][:3, 3]) / (2 * eps)
        Rd = (F1[j][:3, :3] - F0[j][:3, :3]) / (2 * eps)
        W = R.T @ Rd  # [w]x in the joint frame
        w = np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]) * 0.5
        sv = np.concatenate([w, R.T @ v_origin])
        T += 0.5 * sv @ I @ sv
        mass = I[3, 3]
        # com in joint frame from the spatial inertia: upper-right block = m [c]x
        mc = np.array([I[2, 4], I[0, 5], I[1, 3]])
        com_world = Fc[j][:3, 3] + R @ (mc / mass)
        V += mass * (-m.gravity) @ com_world
    return T, V


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_kinetic_energy_matches_finite_difference_fk(oracle, name):
    m = oracle.Model(name)
    x = _inputs(name, E=3)
    for e in range(3):
        pose, vel = x["pose"][e], x["vel"][e]
        M, _ = m.rbd_update(pose, vel)
        T_fd, _ = _energies(m, pose, vel)
        T_m = 0.5 * vel @ M @ vel
        assert T_m == pytest.approx(T_fd, rel=2e-7)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_gravity_bias_matches_potential_gradient(oracle, name):
    m = oracle.Model(name)
    x = _inputs(name, E=2)
    for e in range(2):
        pose = x["pose"][e]
        _, C0 = m.rbd_update(pose, np.zeros(m.n))
        dead = [6] + [int(m.offset[j]) + 3 for j in range(1, m.N) if m.jtype[j] == 4]
        eps = 1e-6
        for i in range(m.n):
            if i in dead:
                continue
            d = np.zeros(m.n)
            d[i] = 1.0
            _, V1 = _energies(m, _integrate(m, pose, d, eps), np.zeros(m.n))
            _, V0 = _energies(m, _integrate(m, pose, d, -eps), np.zeros(m.n))
            dV = (V1 - V0) / (2 * eps)
            assert C0[i] == pytest.approx(dV, rel=1e-6, abs=1e-6)


def test_ldlt_restatement(oracle):
    rng = np.random.default_rng(0)
    for n in (1, 2, 9, 27):
        A = rng.normal(size=(n, n))
        A = A @ A.T + n * np.eye(n)
        b = rng.normal(size=n)
        x, zp = oracle.ldlt_solve(A, b)
        assert zp == 0
        assert np.allclose(A @ x, b, rtol=1e-11, atol=1e-11)
    # structurally-null row/column -> that component is 0 and the rest solves the reduced system (Q1)
    n = 8
    A = rng.normal(size=(n, n))
    A = A @ A.T + n * np.eye(n)
    A[3, :] = 0
    A[:, 3] = 0
    b = rng.normal(size=n)
    x, zp = oracle.ldlt_solve(A, b)
    keep = [i for i in range(n) if i != 3]
    assert zp == 1 and x[3] == 0
    assert np.allclose(A[np.ix_(keep, keep)] @ x[keep], b[keep], rtol=1e-11, atol=1e-11)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_stable_pd_definition(oracle, name):
    """tau = Kp e_p + Kd (e_v - dt qdd), (M + dt Kd) qdd = Kp e_p + Kd e_v - C, evaluated independently in numpy
    from the oracle's own M, C and error vectors; root torque entries are 0; inactive joints use masked gains on
    the right-hand side but the unmasked Kd on the diagonal (Q2). Targets here are joint-local (UseWorldCoord cleared;
    the world-target conversion is pinned against the reference in test_ref_pin.py / test_golden.py)."""
    m0 = oracle.Model(name)
    pd = m0.pd_params.copy()
    pd[:, 17] = 0
    m = oracle.Model(joint_mat=m0.joint_mat, body_defs=m0.body_defs, pd_params=pd)
    x = _inputs(name, E=4)
    dt = 1.0 / 600
    for e in range(4):
        pose, vel, tp, tv = x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e]
        act = x["joint_active"][e].copy()
        if e == 1 and m.N > 2:
            act[2] = 0
        tau, M, C, zp = m.imp_pd(dt, pose, vel, tp, tv, act)
        kp = np.zeros(m.n)
        kd = np.zeros(m.n)
        kd_diag = np.zeros(m.n)
        for j in range(1, m.N):
            o, s = int(m.offset[j]), int(m.size[j])
            kd_diag[o:o + s] = m.pd_params[j, 2]
            if act[j]:
                kp[o:o + s] = m.pd_params[j, 1]
                kd[o:o + s] = m.pd_params[j, 2]
        pose_inc = m.post_process_pose(pose + dt * m.vel_to_pose_diff(pose, vel))
        tpe = tp.copy()
        tpe[:7] = 0
        e_p = m.calc_vel(pose_inc, tpe, 1.0)
        e_v = tv - vel
        e_v[:7] = -vel[:7]
        A = M + np.diag(dt * kd_diag)
        rhs = kp * e_p + kd * e_v - C
        live = [i for i in range(m.n) if A[i, i] != 0]
        qdd = np.zeros(m.n)
        qdd[live] = np.linalg.solve(A[np.ix_(live, live)], rhs[live])
        ref = kp * e_p + kd * (e_v - dt * qdd)
        assert np.all(tau[:7] == 0)
        assert np.allclose(tau, ref, rtol=1e-9, atol=1e-9 * max(1.0, np.abs(ref).max()))
        assert zp >= 1  # root slot 6


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_com_velocity_is_time_derivative_of_com(oracle, name):
    """cRBDUtil::CalcCoM's velocity (end-effector Jacobians) equals the finite-difference derivative of its position."""
    m = oracle.Model(name)
    x = _inputs(name, E=3)
    eps = 1e-6
    for e in range(3):
        pose, vel = x["pose"][e], x["vel"][e]
        _, cv = oracle.calc_com(m, pose, vel)
        c1, _ = oracle.calc_com(m, _integrate(m, pose, vel, eps), vel)
        c0, _ = oracle.calc_com(m, _integrate(m, pose, vel, -eps), vel)
        assert np.allclose(cv, (c1 - c0) / (2 * eps), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["biped3d", "dog"])
def test_imitation_reward_properties(oracle, name):
    """Reward of a pose against itself is the maximum of every term that depends only on the poses; a fallen
    character gets 0; the weights sum to 1."""
    m = oracle.Model(name)
    x = _inputs(name, E=3)
    for e in range(3):
        kp, kv = m.kin_char_pose(x["kin_time"][e])
        r, t = oracle.imitate_reward(m, None, kp, kv, kp, kv, np.zeros((m.N, 3)))
        assert t[0] < 1e-13 and t[1] == 0 and t[3] == 0 and t[2] < 1e-20  # 2*acos(w~1) amplifies rounding
        r2, _ = oracle.imitate_reward(m, None, x["pose"][e], x["vel"][e], kp, kv, x["body_vel"][e])
        assert 0 < r2 < r <= 1.0
        assert oracle.imitate_reward(m, None, x["pose"][e], x["vel"][e], kp, kv, x["body_vel"][e], fallen=1)[0] == 0


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_rbd_extras_identities(oracle, name):
    """SURVEY 8f rank 2 (BuildJacobian / BuildCOMJacobian / CalcGravityForce / BuildEndEffectorJacobian), pinned by
    identities that do not depend on the restatement: tau_g = -C(q, 0); the linear rows of Jcom times qdot equal the
    CoM velocity of trlo_calc_com (computed through the end-effector-Jacobian path); J restricted to a joint's chain
    equals that joint's end-effector Jacobian; the chain Jacobian times qdot is the finite-difference velocity of the
    joint origin."""
    m = oracle.Model(name)
    x = synth.make_inputs(name, 6, seed=11)
    for e in range(6):
        q, qd = x["pose"][e], x["vel"][e]
        J, Jc, g = m.rbd_extras(q)
        _, C0 = m.rbd_update(q, np.zeros_like(qd))
        assert np.allclose(g, -C0, rtol=1e-10, atol=1e-9 * max(1.0, np.abs(C0).max()))
        com, com_vel = oracle.calc_com(m, q, qd)
        assert np.allclose(Jc[3:] @ qd, com_vel, rtol=1e-9, atol=1e-9)
        for j in range(m.N):
            Je = m.end_eff_jacobian(q, j)
            mask = np.zeros(m.n, dtype=bool)
            k = j
            while k != -1:
                o, sz = int(m.offset[k]), int(m.size[k])
                mask[o:o + sz] = True
                k = int(m.parent[k])
            assert np.allclose(Je[:, mask], J[:, mask], rtol=1e-10, atol=1e-11)
            assert np.all(Je[:, ~mask] == 0)
        # finite-difference velocity of the last joint's origin vs J qdot (linear part at the world origin:
        # v_origin + omega x (-p) ... compare the point velocity v + omega x p)
        j = m.N - 1
        h = 1e-6
        q1 = _integrate(m, q, qd, h)
        p0 = m.joint_world_trans(q, j)[:3, 3]
        p1 = m.joint_world_trans(q1, j)[:3, 3]
        sv = m.end_eff_jacobian(q, j) @ qd
        v_point = sv[3:] + np.cross(sv[:3], p0)
        assert np.allclose((p1 - p0) / h, v_point, rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_apply_control_forces_identities(oracle, name):
    """SURVEY 8f rank 3, pinned by: action = reaction (body wrenches sum to zero), clamping (norm <= limit, direction
    kept, small torques untouched), and power balance for the rotational joints: sum_b T_b . omega_b = tau . qdot with
    the body angular velocities taken from the end-effector Jacobians."""
    m = oracle.Model(name)
    x = synth.make_inputs(name, 5, seed=3)
    rng = np.random.default_rng(9)
    lim = m.joint_mat[:, 14]
    for e in range(5):
        q, qd = x["pose"][e], x["vel"][e]
        tau = rng.normal(size=m.n) * 20.0
        tau[:7] = 0
        ct, wr = m.apply_control_forces(q, tau)
        assert np.allclose(wr.sum(axis=0), 0, atol=1e-9)
        small = tau * 1e-3
        ct_s, wr_s = m.apply_control_forces(q, small)
        assert np.array_equal(ct_s, small)
        big = tau * 1e3
        ct_b, _ = m.apply_control_forces(q, big)
        for j in range(1, m.N):
            o, sz, ty = int(m.offset[j]), int(m.size[j]), int(m.jtype[j])
            if ty in (0, 4):  # revolute / spherical: pure torque
                k = 1 if ty == 0 else 3
                nb = np.linalg.norm(ct_b[o:o + k])
                assert nb <= lim[j] * (1 + 1e-12)
                if nb > 0:
                    assert np.allclose(ct_b[o:o + k] / nb, big[o:o + k] / np.linalg.norm(big[o:o + k]), atol=1e-12)
        if all(int(t) in (0, 3, 4) for t in m.jtype[1:]):
            power = 0.0
            for j in range(m.N):
                om = (m.end_eff_jacobian(q, j) @ qd)[:3]
                power += wr_s[j, :3] @ om
            ref = 0.0
            for j in range(1, m.N):
                o, ty = int(m.offset[j]), int(m.jtype[j])
                k = 1 if ty == 0 else (3 if ty == 4 else 0)
                ref += small[o:o + k] @ qd[o:o + k]
            assert np.isclose(power, ref, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("name", CONFIG_NAMES)
def test_stable_pd_solve_against_50_digit_arithmetic(oracle, name):
    """an independent high-precision evaluation bounds the oracle's own rounding. Given the oracle's M, C and
    error vectors (doubles), the solve (M + dt Kd) qdd = Kp e_p + Kd e_v - C and tau = Kp e_p + Kd (e_v - dt qdd) are
    redone with mpmath at 50 digits for three poses per character; the oracle's pivoted LDLT in double must agree to 1e-11
    relative (condition numbers of the bundled characters are ~1e3..1e5)."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    m0 = oracle.Model(name)
    pd = m0.pd_params.copy()
    pd[:, 17] = 0
    m = oracle.Model(joint_mat=m0.joint_mat, body_defs=m0.body_defs, pd_params=pd)
    x = _inputs(name, E=3)
    dt = 1.0 / 600
    for e in range(3):
        pose, vel, tp, tv = x["pose"][e], x["vel"][e], x["tar_pose"][e], x["tar_vel"][e]
        tau, M, C, _ = m.imp_pd(dt, pose, vel, tp, tv, x["joint_active"][e])
        kp, kd, kd_diag = np.zeros(m.n), np.zeros(m.n), np.zeros(m.n)
        for j in range(1, m.N):
            o, s = int(m.offset[j]), int(m.size[j])
            kd_diag[o:o + s] = m.pd_params[j, 2]
            if x["joint_active"][e][j]:
                kp[o:o + s], kd[o:o + s] = m.pd_params[j, 1], m.pd_params[j, 2]
        pose_inc = m.post_process_pose(pose + dt * m.vel_to_pose_diff(pose, vel))
        tpe = tp.copy()
        tpe[:7] = 0
        e_p = m.calc_vel(pose_inc, tpe, 1.0)
        e_v = tv - vel
        e_v[:7] = -vel[:7]
        live = [i for i in range(m.n) if M[i, i] + dt * kd_diag[i] != 0]
        A = mp.matrix(len(live), len(live))
        b = mp.matrix(len(live), 1)
        for r, i in enumerate(live):
            b[r] = mp.mpf(kp[i]) * mp.mpf(e_p[i]) + mp.mpf(kd[i]) * mp.mpf(e_v[i]) - mp.mpf(C[i])
            for c, k in enumerate(live):
                A[r, c] = mp.mpf(M[i, k]) + (mp.mpf(dt) * mp.mpf(kd_diag[i]) if i == k else 0)
        qdd = mp.lu_solve(A, b)
        ref = np.zeros(m.n)
        for r, i in enumerate(live):
            ref[i] = float(mp.mpf(kp[i]) * mp.mpf(e_p[i]) + mp.mpf(kd[i]) * (mp.mpf(e_v[i]) - mp.mpf(dt) * qdd[r]))
        for i in range(m.n):
            if i not in live:
                ref[i] = kp[i] * e_p[i] + kd[i] * e_v[i]
        scale = max(1.0, np.abs(ref).max())
        assert np.abs(tau - ref).max() <= 1e-11 * scale, (name, e, np.abs(tau - ref).max() / scale)
