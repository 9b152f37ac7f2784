"""Spectral-element reference-element quantities for npol-th order GLL / GLJ(0,1) bases.

Host-side restatement (numpy, float64) of what the reference MESHER computes once and
stores in the mesh database:

  * GLL nodes / weights          MESHER/splib.f90:162-203 (zelegl), :465-482 (get_welegl)
  * GLJ(0,1) nodes / weights     MESHER/splib.f90:253-292 (zemngl2), :495-520
                                 (get_welegl_axial, iflag=2), :534-577 (vamnpo)
  * derivative matrices          MESHER/gllmeshgen.f90:61-94
        G2(j,i) = l_j'(eta_i)      (GLL Lagrange interpolants)
        G1(j,i) = lbar_j'(xi_i)    (Lagrange interpolants through the GLJ nodes)
        G0      = G1(:,0)
    all *stored and used in single precision* (SOLVER/data_spec.f90:38-39).

Nothing here is on the device hot path; it only produces inputs.
"""
from __future__ import annotations

import numpy as np

__all__ = ["gll_points_weights", "glj_points_weights", "lagrange_deriv_matrix",
           "SpectralBasis"]


def _legendre(n: int, x: np.ndarray):
    """P_n(x) and P_n'(x) by the three-term recurrence."""
    x = np.asarray(x, dtype=np.float64)
    p0 = np.ones_like(x)
    if n == 0:
        return p0, np.zeros_like(x)
    p1 = x.copy()
    dp0 = np.zeros_like(x)
    dp1 = np.ones_like(x)
    for k in range(2, n + 1):
        p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
        dp2 = dp0 + (2 * k - 1) * p1
        p0, p1 = p1, p2
        dp0, dp1 = dp1, dp2
    return p1, dp1


def gll_points_weights(n: int):
    """Gauss-Lobatto-Legendre nodes (roots of (1-x^2) P_n') and weights 2/(n(n+1) P_n^2).

    splib.f90:162-203 finds them by Newton iteration; here the interior nodes are the
    eigenvalues of the Jacobi(1,1) Golub-Welsch matrix polished by Newton on P_n'.
    """
    if n < 1:
        raise ValueError("n >= 1")
    if n == 1:
        x = np.array([-1.0, 1.0])
    else:
        k = np.arange(1, n - 1, dtype=np.float64)
        # Jacobi (1,1) recurrence off-diagonals
        b = np.sqrt(k * (k + 2.0) / ((2.0 * k + 1.0) * (2.0 * k + 3.0)))
        J = np.diag(b, 1) + np.diag(b, -1)
        xi = np.sort(np.linalg.eigvalsh(J)) if n > 2 else np.array([0.0])
        # Newton polish on q(x) = P_n'(x)
        for _ in range(50):
            p, dp = _legendre(n, xi)
            # P_n'' from the Legendre ODE: (1-x^2) P'' = 2 x P' - n(n+1) P
            d2p = (2.0 * xi * dp - n * (n + 1) * p) / (1.0 - xi * xi)
            dx = dp / d2p
            xi = xi - dx
            if np.max(np.abs(dx)) < 1e-16:
                break
        x = np.concatenate(([-1.0], xi, [1.0]))
    # symmetrise
    x = 0.5 * (x - x[::-1])
    p, _ = _legendre(n, x)
    w = 2.0 / (n * (n + 1) * p * p)
    return x, w


def _vamnpo(n: int, x: float) -> float:
    """m_n(x) = (L_n + L_{n+1})/(1+x), recurrence of splib.f90:534-577 (value only)."""
    y = 1.0
    if n == 0:
        return y
    y = 1.5 * x - 0.5
    if n == 1:
        return y
    yp = 1.0
    for i in range(2, n + 1):
        c1 = float(i - 1)
        ym = y
        y = (x - 1.0 / ((2 * c1 + 1.0) * (2 * c1 + 3.0))) * y - (c1 / (2.0 * c1 + 1.0)) * yp
        y = (2.0 * c1 + 3.0) * y / (c1 + 2.0)
        yp = ym
    return y


def glj_points_weights(n: int):
    """Gauss-Lobatto-Jacobi(0,1) nodes on [-1,1] used along the axis (xi direction of
    axial elements) and their weights for the measure (1+xi) d xi.

    Nodes: eigenvalues of the (n-1)x(n-1) tridiagonal of splib.f90:275-287.
    Weights: 4/(n(n+2)) / m_n(xi)^2, doubled at xi=-1 (splib.f90:506-513).
    """
    if n < 2:
        raise ValueError("n >= 2")
    i = np.arange(1, n, dtype=np.float64)
    d = 3.0 / (4.0 * (i + 0.5) * (i + 1.5))
    k = np.arange(1, n - 1, dtype=np.float64)
    e = np.sqrt(k * (k + 3.0)) / (2.0 * (k + 1.5))
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    inner = np.sort(np.linalg.eigvalsh(T))
    x = np.concatenate(([-1.0], inner, [1.0]))
    fact = 4.0 / (n * (n + 2.0))
    w = np.array([fact / _vamnpo(n, float(xj)) ** 2 for xj in x])
    w[0] *= 2.0
    return x, w


def lagrange_deriv_matrix(x: np.ndarray) -> np.ndarray:
    """D[j, i] = l_j'(x_i) for the Lagrange interpolants l_j through nodes x.

    Barycentric form; equals hn_jprime (splib.f90:136-157) on GLL nodes and
    lag_interp_deriv_wgl (splib.f90:71-129) on GLJ nodes.
    """
    x = np.asarray(x, dtype=np.float64)
    n = x.size
    c = np.array([np.prod([x[j] - x[m] for m in range(n) if m != j]) for j in range(n)])
    D = np.zeros((n, n))
    for j in range(n):
        for i in range(n):
            if i != j:
                D[j, i] = (c[i] / c[j]) / (x[i] - x[j])
    for i in range(n):
        D[i, i] = -np.sum([D[j, i] for j in range(n) if j != i])
    return D


class SpectralBasis:
    """Everything `data_spec` holds (SOLVER/data_spec.f90:36-47) for one npol."""

    def __init__(self, npol: int = 4):
        self.npol = npol
        self.eta, self.wt = gll_points_weights(npol)
        self.xi_k, self.wt_axial_k = glj_points_weights(npol)
        D2 = lagrange_deriv_matrix(self.eta)     # D2[j,i] = l_j'(eta_i)
        D1 = lagrange_deriv_matrix(self.xi_k)
        # Fortran G2(j,i) (first index j) -> numpy array indexed [j, i]
        self.G2_dp = D2
        self.G1_dp = D1
        # single precision copies, as stored in the mesh database
        self.G2 = D2.astype(np.float32)
        self.G2T = np.ascontiguousarray(self.G2.T)
        self.G1 = D1.astype(np.float32)
        self.G1T = np.ascontiguousarray(self.G1.T)
        self.G0 = np.ascontiguousarray(self.G1[:, 0])
