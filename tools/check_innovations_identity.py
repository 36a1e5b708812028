#!/usr/bin/env python3
"""Numerical check (numpy, fp64) of the identity the constraint phase of the step kernel rests on (DESIGN.md section 3,
"whitened rows"):  for a floating-base tree, with the articulated-body quantities U_c = IA_c S_c, d_c = S_c.U_c and
the innovations u^f of a generalized impulse f (what the INWARD sweep of the articulated-body algorithm computes:
u_c = f_c - S_c.p_c,  p_parent += p_c + U_c u_c / d_c,  u_0 = f_0 - p_0),

      g^T M^-1 f  =  sum_c u_c^g u_c^f / d_c  +  u_0^g . IA_0^-1 u_0^f                  (M = joint-space inertia)

i.e. M^-1 = L^T D^-1 L with L the inward sweep: a constraint row J_i only needs its inward sweep z_i = D^-1/2 L J_i^T,
the Delassus matrix is Z Z^T, and projected Gauss-Seidel can run on z = Z^T lambda without any outward sweep per row.
M here comes from an independent composite-Jacobian construction."""
import numpy as np


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def main(seed=0):
    rng = np.random.default_rng(seed)
    parent = [-1, 0, 1, 2, 3, 4, 2, 6, 7, -1, 9, 10, -1, 12]   # humanoid-like: spine of 3, two legs off link 2, two arms off the base
    n = len(parent)

    def rand_inertia():
        m = rng.uniform(0.5, 5)
        c = rng.normal(size=3) * 0.3
        A = rng.normal(size=(3, 3))
        Ic = A @ A.T * 0.05 + np.eye(3) * 0.01
        C = skew(c)
        return np.block([[Ic + m * C @ C.T, m * C], [m * C.T, m * np.eye(3)]])

    Ib = [rand_inertia() for _ in range(n)]
    I0 = rand_inertia()
    S = []
    for _ in range(n):
        ax = rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        r = rng.normal(size=3)
        S.append(np.concatenate([ax, np.cross(r, ax)]))
    anc = []
    for c in range(n):
        a, k = [], c
        while k >= 0:
            a.append(k)
            k = parent[k]
        anc.append(a)
    # joint-space inertia from body Jacobians (generalized velocity = [base 6 | joints n])
    M = np.zeros((6 + n, 6 + n))
    Jb = np.zeros((6, 6 + n))
    Jb[:, :6] = np.eye(6)
    M += Jb.T @ I0 @ Jb
    for b in range(n):
        J = np.zeros((6, 6 + n))
        J[:, :6] = np.eye(6)
        for c in anc[b]:
            J[:, 6 + c] = S[c]
        M += J.T @ Ib[b] @ J
    # articulated inertias
    IA = [Ib[c].copy() for c in range(n)]
    IA0 = I0.copy()
    U, d = [None] * n, [0.0] * n
    for c in reversed(range(n)):
        U[c] = IA[c] @ S[c]
        d[c] = S[c] @ U[c]
        down = IA[c] - np.outer(U[c], U[c]) / d[c]
        if parent[c] >= 0:
            IA[parent[c]] += down
        else:
            IA0 += down

    def inward(f):
        p = [np.zeros(6) for _ in range(n)]
        p0 = np.zeros(6)
        u = np.zeros(n)
        for c in reversed(range(n)):
            u[c] = f[6 + c] - S[c] @ p[c]
            out = p[c] + U[c] * u[c] / d[c]
            if parent[c] >= 0:
                p[parent[c]] += out
            else:
                p0 += out
        return u, f[:6] - p0

    worst = 0.0
    Minv = np.linalg.inv(M)
    for _ in range(200):
        f, g = rng.normal(size=6 + n), rng.normal(size=6 + n)
        if rng.uniform() < 0.5:   # sparse impulses, as constraint rows are
            f[:] = 0
            f[6 + rng.integers(n)] = 1
        uf, bf = inward(f)
        ug, bg = inward(g)
        lhs = g @ Minv @ f
        rhs = sum(ug[c] * uf[c] / d[c] for c in range(n)) + bg @ np.linalg.solve(IA0, bf)
        worst = max(worst, abs(lhs - rhs) / max(1.0, abs(lhs)))
    print("max relative deviation over 200 random impulse pairs: %.2e" % worst)
    assert worst < 1e-9
    return worst


if __name__ == "__main__":
    main()
