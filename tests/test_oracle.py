"""CPU tests of the oracle (oracle/dq_oracle.c), runnable without a GPU.

The reference ships no expected values, so the oracle is pinned by
  * analytic properties: closed forms, KKT residuals, finite differences (test_script.py:23-43 recipe),
    the Tikhonov closed form of the backward;
  * the deterministic input fixtures recoverable from the reference (Solver.cpp:708-712, :901-923);
  * the reference's own Solver.cpp compiled against the stand-in linear-algebra header (oracle/_ref),
    (test_oracle_equals_live_reference_build) and outputs of that build committed as golden vectors
    (test_oracle_reproduces_reference_build_outputs);
  * committed golden vectors (tests/golden/, made by scripts/make_golden.py).
"""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def rng(seed):
    return np.random.default_rng(seed)


def spd(r, n, shift=0.1):
    S = 2 * r.random((n, n)) - 1
    return S @ S.T / n + shift * np.eye(n)


# ------------------------------------------------------------------ helpers (Solver.cpp:15-59)
def test_power_iteration_matches_eigenvalue(oracle):
    r = rng(0)
    for n in (2, 5, 8, 16, 32):
        A = spd(r, n)
        lam = np.linalg.eigvalsh(A).max()
        assert abs(oracle.power_iteration(A, 100) - lam) <= 1e-3 * lam
        # the fixed 10-step estimate (QP, Solver.cpp:71) is a lower bound within a few percent here
        L10 = oracle.power_iteration(A, 10)
        assert 0.5 * lam <= L10 <= lam * (1 + 1e-12)


def test_power_iteration_diagonal_uniform(oracle):
    # v0 is the uniform vector: for c*I every step returns it and L == c
    assert oracle.power_iteration(3.5 * np.eye(7), 10) == pytest.approx(3.5, rel=1e-15)


def test_iterative_refinement_is_tikhonov_solve(oracle):
    r = rng(1)
    for m in (1, 3, 8, 20, 48):
        A = spd(r, m, 0.5)
        b = r.standard_normal(m)
        x, it = oracle.iterative_refinement(A, b)
        ref = np.linalg.solve(A.T @ A + 1e-7 * np.eye(m), A.T @ b)
        assert np.abs(x - ref).max() <= 1e-9 * max(1, np.abs(ref).max())
        assert 1 <= it <= 10


# ------------------------------------------------------------------ QP forward (Solver.cpp:61-123)
def test_qp_fixture_solver_cpp_708(oracle):
    # Solver.cpp:708-712: P = diag(5e-4, 3, 0, 0), q = (-8000, 0, 0, 0) -> x* = (1.6e7, 0, 0, 0)
    P = np.diag([5e-4, 3.0, 0.0, 0.0])
    q = np.array([-8000.0, 0, 0, 0])
    x, it = oracle.solveQP(P, q, np.zeros(4), 1e-10, 1e-7, 1000, return_iters=True)
    assert x[0] == pytest.approx(1.6e7, rel=1e-9)
    assert np.all(x[1:] == 0.0)
    assert 1 < it < 1000


def test_qp_readme_example_is_degenerate(oracle):
    # README.md:35-39: q = rand >= 0, diagonal P >= 0 -> x* = 0 after one iteration
    r = rng(2)
    P = np.stack([np.diag(r.random(8)) for _ in range(10)])
    q = r.random((10, 8, 1))
    x, it = oracle.qp_forward(P, q, None, 1e-7, 1000, return_iters=True)
    assert np.all(x == 0.0) and np.all(it == 1)


def test_qp_diagonal_closed_form(oracle):
    r = rng(3)
    B, N = 512, 8
    p = r.random((B, N)) + 0.1
    q = 2 * r.random((B, N, 1)) - 1
    P = np.einsum("bi,ij->bij", p, np.eye(N))
    x = oracle.qp_forward(P, q, None, 1e-12, 10000)
    # the prox term mu_prox only regularises the iteration, the fixed point is the QP optimum
    assert np.abs(x[:, :, 0] - np.maximum(-q[:, :, 0] / p, 0)).max() <= 1e-6  # dual-residual stop only (F3)


def test_qp_dense_kkt(oracle):
    # The QP stops on the dual residual alone (Solver.cpp:88), so a few problems end before a clipped
    # coordinate has left the bound: KKT holds for the large majority, x >= 0 for all.
    r = rng(4)
    ok = tot = 0
    for n in (3, 8, 16, 32):
        for _ in range(50):
            P = spd(r, n)
            q = 2 * r.random(n) - 1
            x = oracle.solveQP(P, q, np.zeros(n), 1e-12, 1e-7, 10000)
            g = P @ x + q
            assert np.all(x >= 0)
            stat = np.abs(g[x > 0]).max(initial=0) <= 1e-6     # stationarity on the free set
            dual = g[x == 0].min(initial=0) >= -1e-6           # dual feasibility on the active set
            ok += bool(stat and dual)
            tot += 1
    assert ok >= 0.9 * tot


def test_qp_warm_start_is_dead(oracle):
    # Solver.cpp:70 -> :80: l = warm_start is overwritten before it is read
    r = rng(5)
    P, q = spd(r, 8), 2 * r.random(8) - 1
    a = oracle.solveQP(P, q, np.zeros(8), 1e-7, 1e-7, 1000)
    b = oracle.solveQP(P, q, 100 * r.standard_normal(8), 1e-7, 1e-7, 1000)
    assert np.array_equal(a, b)


def test_qp_max_iter_and_iters(oracle):
    r = rng(6)
    P, q = spd(r, 8), 2 * r.random(8) - 1
    x1, it1 = oracle.solveQP(P, q, np.zeros(8), 1e-12, 1e-7, 3, return_iters=True)
    assert it1 == 3
    x0, it0 = oracle.solveQP(P, q, np.zeros(8), 1e-12, 1e-7, 0, return_iters=True)
    assert it0 == 0 and np.all(x0 == 0)


def test_qp_early_stop_differs_from_optimum(oracle):
    # the QP stops on the dual residual alone, so x_ref is a trajectory point, not the optimum
    r = rng(7)
    B, N = 4096, 8
    p = r.random((B, N))
    q = 2 * r.random((B, N, 1)) - 1
    P = np.einsum("bi,ij->bij", p, np.eye(N))
    x = oracle.qp_forward(P, q, None, 1e-7, 1000)
    d = np.abs(x[:, :, 0] - np.maximum(-q[:, :, 0] / p, 0)).max(1)
    assert (d > 1e-6).mean() > 0.02


# ------------------------------------------------------------------ QP backward (Solver.cpp:125-196)
def test_qp_backward_tikhonov_closed_form(oracle):
    # for diagonal P the free entries are p g / (p^2 + 1e-7), active entries 0
    r = rng(8)
    B, N = 256, 8
    p = r.random((B, N)) + 1e-3
    q = 2 * r.random((B, N, 1)) - 1
    g = 2 * r.random((B, N, 1)) - 1
    P = np.einsum("bi,ij->bij", p, np.eye(N))
    x = oracle.qp_forward(P, q, None, 1e-7, 1000)
    gP, gq = oracle.qp_backward(P, q, x, g)
    gamma = -(p * x[:, :, 0] + q[:, :, 0])
    gamma[x[:, :, 0] > 1e-10] = 0
    free = ~(gamma < -1e-10)
    dl = np.where(free, p * g[:, :, 0] / (p * p + 1e-7), 0.0)
    assert np.abs(-gq[:, :, 0] - dl).max() <= 1e-9 * np.abs(dl).max()
    assert np.abs(gP + dl[:, :, None] * x[:, None, :, 0]).max() <= 1e-9 * max(1, np.abs(gP).max())


def test_qp_backward_finite_differences(oracle):
    # recipe of test_script.py:23-43: n=2, P = S S^T, S = rand + 0.01, q = -rand - 0.1, eps=1e-12
    r = rng(5)
    S = r.random((2, 2)) + 0.01
    P, q = S @ S.T, -r.random(2) - 0.1
    solve = lambda P_, q_: oracle.solveQP(P_, q_, np.zeros(2), 1e-13, 1e-7, 100000)
    x = solve(P, q)
    for j in range(2):
        e = np.zeros(2); e[j] = 1.0
        dl = oracle.solveDerivativesQP(P, q, x, e)          # d x_j / d(-q) and -(d x_j / dP) / x
        gq, gP = -dl, -np.outer(dl, x)
        d = 1e-6
        for k in range(2):
            dq = np.zeros(2); dq[k] = d
            num = (solve(P, q + dq)[j] - solve(P, q - dq)[j]) / (2 * d)
            assert num == pytest.approx(gq[k], abs=2e-5)
            for m in range(2):
                dP = np.zeros((2, 2)); dP[k, m] = d
                num = (solve(P + dP, q)[j] - solve(P - dP, q)[j]) / (2 * d)
                assert num == pytest.approx(gP[k, m], abs=2e-5)


# ------------------------------------------------------------------ QCQP (Solver.cpp:505-691)
def qcqp_case(r, n, shift=0.1):
    P = spd(r, n, shift)
    q = 2 * r.random(n) - 1
    l_n = 2 * r.random(n // 2)
    mu = r.random(n // 2)
    return P, q, l_n, mu


def test_qcqp_feasible_and_stationary(oracle):
    r = rng(9)
    for n in (2, 6, 8, 16, 24, 32):
        P, q, l_n, mu = qcqp_case(r, n)
        x = oracle.solveQCQP(P, q, l_n, mu, np.zeros(n), 1e-12, 1e-7, 100000)
        rad = l_n * mu
        nrm = np.hypot(x[0::2], x[1::2])
        assert np.all(nrm <= rad * (1 + 1e-9) + 1e-12)
        g = P @ x + q
        for c in range(n // 2):
            xc, gc = x[2 * c:2 * c + 2], g[2 * c:2 * c + 2]
            if nrm[c] < rad[c] * (1 - 1e-6):
                assert np.abs(gc).max() <= 1e-5                    # interior: gradient vanishes
            else:
                gam = -(xc @ gc) / (2 * xc @ xc)                   # boundary: g = -2 gamma x, gamma >= 0
                assert gam >= -1e-7
                assert np.abs(gc + 2 * gam * xc).max() <= 1e-4 * max(1, np.abs(gc).max())


def test_qcqp_fixture_solver_cpp_901(oracle):
    # Solver.cpp:901-923: the 8x8 G4 (singular, rank 4), g4, radii l_ng4 * 0.15; 10 iterations as at :936.
    G4 = np.array([[2.8750, -0.3750, 2.1250, -0.3750, 2.8750, 0.3750, 2.1250, 0.3750],
                   [-0.3750, 2.8750, 0.3750, 2.8750, -0.3750, 2.1250, 0.3750, 2.1250],
                   [2.1250, 0.3750, 2.8750, 0.3750, 2.1250, -0.3750, 2.8750, -0.3750],
                   [-0.3750, 2.8750, 0.3750, 2.8750, -0.3750, 2.1250, 0.3750, 2.1250],
                   [2.8750, -0.3750, 2.1250, -0.3750, 2.8750, 0.3750, 2.1250, 0.3750],
                   [0.3750, 2.1250, -0.3750, 2.1250, 0.3750, 2.8750, -0.3750, 2.8750],
                   [2.1250, 0.3750, 2.8750, 0.3750, 2.1250, -0.3750, 2.8750, -0.3750],
                   [0.3750, 2.1250, -0.3750, 2.1250, 0.3750, 2.8750, -0.3750, 2.8750]])
    g4 = np.array([3.9650e-01, 1.3222e-16, 3.9650e-01, 1.3222e-16, 3.9650e-01, 2.9742e-16, 3.9650e-01, 2.9742e-16])
    rad = np.array([0.0159, 0.0159, 0.0086, 0.0086]) * 0.15
    # the C++ solver takes mul_n directly; the binding multiplies l_n and mu (pybindings.cpp:57)
    x, it = oracle.solveQCQP(G4, g4, rad, np.ones(4), np.zeros(8), 1e-10, 1e-7, 10, return_iters=True)
    assert it <= 10 and np.all(np.isfinite(x))
    assert np.all(np.hypot(x[0::2], x[1::2]) <= rad * (1 + 1e-12))
    x, it = oracle.solveQCQP(G4, g4, rad, np.ones(4), np.zeros(8), 1e-10, 1e-7, 100000, return_iters=True)
    # the linear term pushes every contact to its friction limit along -e_x
    assert np.allclose(x[0::2], -rad, rtol=1e-6) and np.abs(x[1::2]).max() <= 1e-6


def test_qcqp_backward_finite_differences(oracle):
    r = rng(10)
    n = 6
    P, q, l_n, mu = qcqp_case(r, n, shift=1.0)
    l_n = l_n * 0.2 + 0.05                        # keep most contacts on the boundary
    solve = lambda q_, l_, m_: oracle.solveQCQP(P, q_, l_, m_, np.zeros(n), 1e-13, 1e-7, 200000)
    x = solve(q, l_n, mu)
    w = 2 * r.random(n) - 1
    gP, gq, gl, gm = oracle.qcqp_backward(P[None], q[None, :, None], l_n[None, :, None], mu[None, :, None],
                                          x[None, :, None], w[None, :, None])
    d = 1e-6
    for k in range(n):
        e = np.zeros(n); e[k] = d
        num = w @ (solve(q + e, l_n, mu) - solve(q - e, l_n, mu)) / (2 * d)
        # the QCQP stop test is relative (eps_rel = 1e-4, Solver.cpp:524,548): x* carries ~1e-5 noise
        assert num == pytest.approx(gq[0, k, 0], abs=5e-4)
    for c in range(n // 2):
        e = np.zeros(n // 2); e[c] = d
        num = w @ (solve(q, l_n + e, mu) - solve(q, l_n - e, mu)) / (2 * d)
        assert num == pytest.approx(gl[0, c, 0], abs=2e-3, rel=5e-3)
        num = w @ (solve(q, l_n, mu + e) - solve(q, l_n, mu - e)) / (2 * d)
        assert num == pytest.approx(gm[0, c, 0], abs=2e-3, rel=5e-3)


def test_qcqp_zero_radius_contact(oracle):
    # l_n = 0 ("takes into account indefinite case when l_n is null", Solver.cpp:598,:639)
    r = rng(11)
    P, q, l_n, mu = qcqp_case(r, 8)
    l_n[1] = 0.0
    x = oracle.solveQCQP(P, q, l_n, mu, np.zeros(8), 1e-10, 1e-7, 10000)
    assert np.all(x[2:4] == 0.0)
    E1, E2, blg = oracle.solveDerivativesQCQP(P, q, l_n, mu, x, np.ones(8))
    assert np.all(np.isfinite(blg)) and blg[1] == 0.0 and E1[1, 1] == 0.0 and E2[1, 1] == 0.0


# ------------------------------------------------------------------ batched entry points (qcqp.py loops)
def test_batched_equals_per_problem_and_threads(oracle):
    r = rng(12)
    B, N = 33, 8
    P = np.stack([spd(r, N) for _ in range(B)])
    q = 2 * r.random((B, N, 1)) - 1
    g = 2 * r.random((B, N, 1)) - 1
    x_all, it_all = oracle.qp_forward(P, q, None, 1e-7, 1000, return_iters=True)
    x_1 = oracle.qp_forward(P, q, None, 1e-7, 1000, threads=1)
    assert np.array_equal(x_all, x_1)
    for i in (0, 7, 32):
        xi, iti = oracle.solveQP(P[i], q[i], np.zeros(N), 1e-7, 1e-7, 1000, return_iters=True)
        assert np.array_equal(xi, x_all[i, :, 0]) and iti == it_all[i]
    gP, gq = oracle.qp_backward(P, q, x_all, g)
    for i in (0, 7, 32):
        dl = oracle.solveDerivativesQP(P[i], q[i], x_all[i], g[i])
        assert np.array_equal(gq[i, :, 0], -dl)                                     # qcqp.py:51
        assert np.array_equal(gP[i], -(dl[:, None] * x_all[i, :, 0][None, :]))      # qcqp.py:49


def test_qcqp_batched_grad_assembly(oracle):
    r = rng(13)
    B, N = 9, 8
    nc = N // 2
    cases = [qcqp_case(r, N) for _ in range(B)]
    P = np.stack([c[0] for c in cases]); q = np.stack([c[1] for c in cases])[:, :, None]
    l_n = np.stack([c[2] for c in cases])[:, :, None]; mu = np.stack([c[3] for c in cases])[:, :, None]
    g = 2 * r.random((B, N, 1)) - 1
    x = oracle.qcqp_forward(P, q, l_n, mu, None, 1e-7, 1000)
    gP, gq, gl, gm = oracle.qcqp_backward(P, q, l_n, mu, x, g)
    for i in range(B):
        E1, E2, blg = oracle.solveDerivativesQCQP(P[i], q[i], l_n[i], mu[i], x[i], g[i])
        dgam, dl = blg[:nc], blg[nc:]                                                # qcqp.py:170-171
        assert np.array_equal(gq[i, :, 0], -dl)
        assert np.array_equal(gP[i], -(dl[:, None] * x[i, :, 0][None, :]))
        assert np.allclose(gl[i, :, 0], E2 @ dgam, rtol=1e-15, atol=0)
        assert np.allclose(gm[i, :, 0], E1 @ dgam, rtol=1e-15, atol=0)


# ------------------------------------------------------------------ committed golden vectors
def test_oracle_reproduces_golden(oracle):
    path = os.path.join(HERE, "golden", "golden_v1.npz")
    G = np.load(path)
    for tag in ("qp_diag8", "qp_dense8", "qp_dense5", "qp_dense32"):
        P, q, g = G[f"{tag}_P"], G[f"{tag}_q"], G[f"{tag}_g"]
        eps = float(G[f"{tag}_eps"])
        x, it = oracle.qp_forward(P, q, None, eps, 1000, return_iters=True)
        assert np.array_equal(it, G[f"{tag}_iters"])
        assert np.abs(x - G[f"{tag}_x"]).max() <= 1e-12 * max(1, np.abs(x).max())
        gP, gq = oracle.qp_backward(P, q, G[f"{tag}_x"], g)
        assert np.allclose(gq, G[f"{tag}_gq"], rtol=1e-9, atol=1e-12)
        assert np.allclose(gP, G[f"{tag}_gP"], rtol=1e-9, atol=1e-12)
    for tag in ("qcqp_dense8", "qcqp_dense16", "qcqp_dense24", "qcqp_diag32"):
        P, q, l_n, mu, g = (G[f"{tag}_{k}"] for k in ("P", "q", "l_n", "mu", "g"))
        eps = float(G[f"{tag}_eps"])
        x, it = oracle.qcqp_forward(P, q, l_n, mu, None, eps, 1000, return_iters=True)
        assert np.array_equal(it, G[f"{tag}_iters"])
        assert np.abs(x - G[f"{tag}_x"]).max() <= 1e-12


def test_oracle_reproduces_reference_build_outputs(oracle):
    """The pin.  golden_v1.npz carries outputs of the reference's OWN qcqplib/Solver.cpp (compiled unmodified
    against oracle/eigen_standin and run in the authoring container by scripts/make_golden.py).  The oracle
    restatement must reproduce them: bit for bit wherever the refinement loop of Solver.cpp:28-42 stops after
    its first step (QP forward/backward, QCQP forward), to a few ulp for the QCQP gradients."""
    G = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))
    assert bool(G["ref_checked"]), "golden file was generated without the reference build"
    for tag in ("qp_diag8", "qp_dense8", "qp_dense5", "qp_dense32", "qp_solver_cpp_708"):
        P, q, g = G[f"{tag}_P"], G[f"{tag}_q"], G[f"{tag}_g"]
        eps = float(G[f"{tag}_eps"])
        x = oracle.qp_forward(P, q, None, eps, 1000)
        assert np.array_equal(x, G[f"{tag}_xref"]), tag
        gP, gq = oracle.qp_backward(P, q, x, g)
        assert np.array_equal(gq, G[f"{tag}_gqref"]) and np.array_equal(gP, G[f"{tag}_gPref"]), tag
    for tag in ("qcqp_dense8", "qcqp_dense16", "qcqp_dense24", "qcqp_diag32"):
        P, q, l_n, mu, g = (G[f"{tag}_{k}"] for k in ("P", "q", "l_n", "mu", "g"))
        eps = float(G[f"{tag}_eps"])
        x = oracle.qcqp_forward(P, q, l_n, mu, None, eps, 1000)
        assert np.array_equal(x, G[f"{tag}_xref"]), tag
        got = oracle.qcqp_backward(P, q, l_n, mu, x, g)
        for a, nm in zip(got, ("gPref", "gqref", "glref", "gmref")):
            b = G[f"{tag}_{nm}"]
            # (mu_ir*AAinv)*x in the reference's expression vs mu_ir*(AAinv*x) here: last-bit differences only
            assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max()), (tag, nm)
    # Solver.cpp:708-712 has the analytic answer x* = (1.6e7, 0, 0, 0)
    assert abs(G["qp_solver_cpp_708_xref"][0, 0, 0] - 1.6e7) <= 1e-2


def test_oracle_equals_live_reference_build(oracle):
    """Same check against the reference build itself (oracle/_ref/libdq_ref.so) on fresh seeded batches,
    wherever that library exists (it is built from /root/reference by `make -C oracle ref`)."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/libdq_ref.so not built (needs /root/reference)")
    r = rng(7)
    for n, B, eps in ((1, 5, 1e-7), (3, 40, 1e-10), (8, 300, 1e-7), (13, 60, 1e-7), (32, 20, 1e-10)):
        P = np.stack([spd(r, n) for _ in range(B)])
        if n == 8:
            P[: B // 2] = np.stack([np.diag(r.random(n)) for _ in range(B // 2)])  # diagonal, ill-conditioned
        q = 2 * r.random((B, n, 1)) - 1
        g = 2 * r.random((B, n, 1)) - 1
        x = oracle.qp_forward(P, q, None, eps, 1000)
        assert np.array_equal(x, pyref.qp_forward(P, q, None, eps, 1000)), n
        for a, b in zip(oracle.qp_backward(P, q, x, g), pyref.qp_backward(P, q, x, g)):
            assert np.array_equal(a, b), n
    for n, B, eps in ((2, 30, 1e-7), (8, 200, 1e-7), (16, 60, 1e-10), (24, 30, 1e-7), (32, 16, 1e-7)):
        P = np.stack([spd(r, n) for _ in range(B)])
        q = 2 * r.random((B, n, 1)) - 1
        l_n, mu = 2 * r.random((B, n // 2, 1)), r.random((B, n // 2, 1))
        g = 2 * r.random((B, n, 1)) - 1
        x = oracle.qcqp_forward(P, q, l_n, mu, None, eps, 1000)
        assert np.array_equal(x, pyref.qcqp_forward(P, q, l_n, mu, None, eps, 1000)), n
        for a, b in zip(oracle.qcqp_backward(P, q, l_n, mu, x, g), pyref.qcqp_backward(P, q, l_n, mu, x, g)):
            assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(b).max()), n
    # per-problem entry points with the binding's default arguments (pybindings.cpp:76-82)
    P, q = spd(r, 6), 2 * r.random(6) - 1
    assert np.array_equal(oracle.solveQP(P, q, np.zeros(6)), pyref.solveQP(P, q, np.zeros(6)))
    assert np.array_equal(oracle.solveQP(P, q, np.ones(6), 1e-7, 1e-7, 50, False), pyref.solveQP(P, q, np.ones(6), 1e-7, 1e-7, 50, False))


def test_oracle_box_forward_equals_live_reference_build(oracle):
    """solveBoxQP / solveSignedBoxQP restatements against the reference build, bit for bit."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/libdq_ref.so not built (needs /root/reference)")
    r = rng(11)
    for n, B, eps in ((1, 9, 1e-7), (4, 60, 1e-10), (8, 200, 1e-7), (13, 40, 1e-7), (32, 12, 1e-10)):
        P = np.stack([spd(r, n) for _ in range(B)])
        if n == 8:
            P[: B // 2] = np.stack([np.diag(r.random(n)) for _ in range(B // 2)])
        q = 2 * r.random((B, n, 1)) - 1
        lo, hi, v = -r.random((B, n, 1)), r.random((B, n, 1)), 2 * r.random((B, n, 1)) - 1
        x = oracle.boxqp_forward(P, q, lo, hi, eps, 1000)
        assert np.array_equal(x, pyref.boxqp_forward(P, q, lo, hi, eps, 1000)), n
        assert np.all(x >= lo) and np.all(x <= hi)
        xs = oracle.boxqp_forward(P, q, lo, hi, eps, 1000, v=v)
        assert np.array_equal(xs, pyref.boxqp_forward(P, q, lo, hi, eps, 1000, v=v)), n
        assert np.all(np.sign(v) * xs <= 0)
    # with the box at [0, +inf) the box QP is the QP
    P = np.stack([spd(r, 6) for _ in range(20)])
    q = 2 * r.random((20, 6, 1)) - 1
    xb = oracle.boxqp_forward(P, q, np.zeros((20, 6, 1)), np.full((20, 6, 1), np.inf), 1e-7, 1000)
    assert np.array_equal(xb, oracle.qp_forward(P, q, None, 1e-7, 1000))


def test_oracle_box_backward(oracle):
    """solveDerivativesBoxQP restatement: bit-identical to the reference build away from degenerate boxes, and its
    gradients (with the documented sign for l_max) agree with central finite differences."""
    from oracle import pyref
    r = rng(12)
    if pyref.available():
        for n, B in ((1, 7), (4, 50), (8, 150), (13, 30), (32, 8)):
            P = np.stack([spd(r, n) for _ in range(B)])
            q, g = 2 * r.random((B, n, 1)) - 1, 2 * r.random((B, n, 1)) - 1
            lo, hi = -0.3 * r.random((B, n, 1)), 0.3 * r.random((B, n, 1))
            x = oracle.boxqp_forward(P, q, lo, hi, 1e-7, 1000)
            for a, b in zip(oracle.boxqp_backward(P, q, lo, hi, x, g), pyref.boxqp_backward(P, q, lo, hi, x, g)):
                assert np.array_equal(a, b), n
    n = 5
    P = spd(r, n)[None]
    q, g = 2 * r.random((1, n, 1)) - 1, 2 * r.random((1, n, 1)) - 1
    lo, hi = -0.2 * np.ones((1, n, 1)), 0.25 * np.ones((1, n, 1))
    f = lambda P, q, lo, hi: float((oracle.boxqp_forward(P, q, lo, hi, 1e-13, 50000) * g).sum())
    x = oracle.boxqp_forward(P, q, lo, hi, 1e-13, 50000)
    assert np.any(x <= lo + 1e-9) or np.any(x >= hi - 1e-9)
    gP, gq, glo, ghi = oracle.boxqp_backward(P, q, lo, hi, x, g)
    d = 1e-6
    for which, grad in ((1, gq), (2, glo), (3, ghi)):
        for i in range(n):
            args = [P.copy(), q.copy(), lo.copy(), hi.copy()]
            args[which][0, i, 0] += d
            up = f(*args)
            args[which][0, i, 0] -= 2 * d
            num = (up - f(*args)) / (2 * d)
            assert abs(num - grad[0, i, 0]) <= 2e-4 * max(1.0, abs(num)), (which, i, num, grad[0, i, 0])


# ------------------------------------------------------------------ the warm-start extension's CPU checker (NOT reference behaviour)
def test_warm_start_extension_checker(oracle):
    """SURVEY.md 8(f) row 2.  The reference never reads warm_start (F2); the CUDA library has an opt-in flag that
    starts the ADMM iteration there.  The oracle restates that extension behind a test hook so the kernels have a
    CPU checker.  With the hook at its default the batched wrappers behave exactly as before."""
    r = rng(21)
    B, N = 256, 8
    P = np.stack([np.diag(r.random(N) + 0.05) for _ in range(B)])
    q = 2 * r.random((B, N, 1)) - 1
    ws = r.random((B, N, 1))
    x_cold, it_cold = oracle.qp_forward(P, q, None, 1e-8, 1000, return_iters=True)
    x_dead, it_dead = oracle.qp_forward(P, q, ws, 1e-8, 1000, return_iters=True)          # default: dead (F2)
    assert np.array_equal(x_cold, x_dead) and np.array_equal(it_cold, it_dead)
    try:
        oracle.set_batch_flags(3)
        x1, it1 = oracle.qp_forward(P, q, x_cold, 1e-8, 1000, return_iters=True)             # start at the solution
        assert it1.mean() <= 3 and np.abs(x1 - x_cold).max() <= 1e-4
        q2 = q + 0.01 * r.standard_normal(q.shape)                                           # the next time step
        oracle.set_batch_flags(1)
        x2c, it2c = oracle.qp_forward(P, q2, None, 1e-8, 1000, return_iters=True)
        oracle.set_batch_flags(3)
        x2w, it2w = oracle.qp_forward(P, q2, x_cold, 1e-8, 1000, return_iters=True)
        assert it2w.mean() < 0.75 * it2c.mean()
        scale = np.maximum(1.0, np.abs(x2c).max(axis=(1, 2)))
        assert (np.abs(x2w - x2c).max(axis=(1, 2)) <= 1e-4 * scale).all()  # both stop on the dual residual only (F3)
        # QCQP: same hook
        nc = N // 2
        l_n, mu = 2 * r.random((B, nc, 1)), r.random((B, nc, 1))
        oracle.set_batch_flags(1)
        xq, itq = oracle.qcqp_forward(P, q, l_n, mu, None, 1e-8, 1000, return_iters=True)
        xq2c, itq2c = oracle.qcqp_forward(P, q2, l_n, mu, None, 1e-8, 1000, return_iters=True)
        oracle.set_batch_flags(3)
        xq1, itq1 = oracle.qcqp_forward(P, q, l_n, mu, xq, 1e-8, 1000, return_iters=True)
        assert itq1.mean() <= 3 and np.abs(xq1 - xq).max() <= 1e-5
        xq2w, itq2w = oracle.qcqp_forward(P, q2, l_n, mu, xq, 1e-8, 1000, return_iters=True)
        assert itq2w.mean() < 0.9 * itq2c.mean() and np.abs(xq2w - xq2c).max() <= 1e-5
    finally:
        oracle.set_batch_flags(1)
