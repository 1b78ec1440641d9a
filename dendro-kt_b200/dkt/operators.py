"""Host-side setup of reference-cell operator matrices for the dense device operator
(K_e = scale * h^alpha * K_ref on an axis-aligned cell of edge h = 2^-level).

These are the dim-generic tensor-product Laplacian / mass matrices on uniform Lagrange nodes -
what HeatMat/HeatVec (FEM/examples/src/heatMat.cpp:46-117, heatVec.cpp:29-76) compute by
sum-factorisation for 3 axes, here for any dim and assembled once on the host.
"""
import numpy as np


def _ref_1d(order):
    M = order + 1
    xs = np.linspace(0.0, 1.0, M)
    g, w = np.polynomial.legendre.leggauss(M + 1)
    g, w = 0.5 * (g + 1.0), 0.5 * w
    phi = np.zeros((M, len(g)))
    dphi = np.zeros((M, len(g)))
    for k in range(M):
        c = np.poly1d([1.0])
        for m in range(M):
            if m != k:
                c = c * np.poly1d([1.0, -xs[m]]) / (xs[k] - xs[m])
        phi[k], dphi[k] = c(g), c.deriv()(g)
    return (phi * w) @ phi.T, (dphi * w) @ dphi.T


def laplace_kref(dim, order):
    """Unit-cube stiffness matrix, axis 0 fastest; K_e = h^(dim-2) K_ref  (alpha = dim-2)."""
    Mm, Km = _ref_1d(order)
    N = (order + 1) ** dim
    K = np.zeros((N, N))
    for a in range(dim):
        T = np.ones((1, 1))
        for d in range(dim):
            T = np.kron(Km if d == a else Mm, T)
        K += T
    return K


def mass_kref(dim, order):
    """Unit-cube mass matrix; M_e = h^dim M_ref  (alpha = dim)."""
    Mm, _ = _ref_1d(order)
    T = np.ones((1, 1))
    for _ in range(dim):
        T = np.kron(Mm, T)
    return T


def laplace_terms(dim, order):
    """The unit-cube stiffness matrix as a sum of `dim` Kronecker terms for dkt.Operator.kron: term a has the 1-D stiffness
    matrix along axis a and the 1-D mass matrix along the others (what laplace_kref assembles densely).  Shape (dim, dim, M, M)."""
    Mm, Km = _ref_1d(order)
    return np.array([[Km if d == a else Mm for d in range(dim)] for a in range(dim)])


def mass_terms(dim, order):
    Mm, _ = _ref_1d(order)
    return np.array([[Mm for _ in range(dim)]])
