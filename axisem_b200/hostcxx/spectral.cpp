#include "spectral.hpp"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace axisem {
namespace {

// P_n and P_n' by the three-term recurrence
void legendre(int n, double x, double &p, double &dp) {
    double p0 = 1.0, p1 = x, d0 = 0.0, d1 = 1.0;
    if (n == 0) { p = 1.0; dp = 0.0; return; }
    for (int k = 2; k <= n; k++) {
        const double p2 = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
        const double d2 = d0 + (2 * k - 1) * p1;
        p0 = p1; p1 = p2; d0 = d1; d1 = d2;
    }
    p = p1; dp = d1;
}

// eigenvalues of a symmetric tridiagonal matrix (diagonal d, off-diagonal e), cyclic Jacobi on
// the full matrix: the matrices here are at most (npol-1) x (npol-1)
std::vector<double> sym_tridiag_eigenvalues(const std::vector<double> &d, const std::vector<double> &e) {
    const int n = (int)d.size();
    std::vector<double> a((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) a[(size_t)i * n + i] = d[i];
    for (int i = 0; i + 1 < n; i++) a[(size_t)i * n + i + 1] = a[(size_t)(i + 1) * n + i] = e[i];
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += a[(size_t)p * n + q] * a[(size_t)p * n + q];
        if (off < 1e-32) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = a[(size_t)p * n + q];
                if (std::abs(apq) < 1e-300) continue;
                const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
                    a[(size_t)k * n + p] = c * akp - s * akq;
                    a[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
                    a[(size_t)p * n + k] = c * apk - s * aqk;
                    a[(size_t)q * n + k] = s * apk + c * aqk;
                }
            }
    }
    std::vector<double> ev(n);
    for (int i = 0; i < n; i++) ev[i] = a[(size_t)i * n + i];
    std::sort(ev.begin(), ev.end());
    return ev;
}

// m_n(x) = (L_n + L_{n+1}) / (1 + x): the recurrence of the GLJ(0,1) quadrature
double vamnpo(int n, double x) {
    if (n == 0) return 1.0;
    double y = 1.5 * x - 0.5, yp = 1.0;
    for (int i = 2; i <= n; i++) {
        const double c1 = i - 1.0, ym = y;
        y = (x - 1.0 / ((2 * c1 + 1.0) * (2 * c1 + 3.0))) * y - (c1 / (2.0 * c1 + 1.0)) * yp;
        y = (2.0 * c1 + 3.0) * y / (c1 + 2.0);
        yp = ym;
    }
    return y;
}

// D[j + n*i] = l_j'(x_i) for the Lagrange interpolants through x
std::vector<double> lagrange_deriv_matrix(const std::vector<double> &x) {
    const int n = (int)x.size();
    std::vector<double> c(n, 1.0), D((size_t)n * n, 0.0);
    for (int j = 0; j < n; j++) for (int m = 0; m < n; m++) if (m != j) c[j] *= x[j] - x[m];
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) if (i != j) D[j + (size_t)n * i] = (c[i] / c[j]) / (x[i] - x[j]);
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int j = 0; j < n; j++) if (j != i) s += D[j + (size_t)n * i];
        D[i + (size_t)n * i] = -s;
    }
    return D;
}

}  // namespace

SpectralBasis spectral_basis(int npol) {
    if (npol < 2) throw std::invalid_argument("spectral_basis: npol >= 2");
    const int n = npol, n1 = npol + 1;
    SpectralBasis b;
    b.npol = npol;
    // GLL: interior nodes are the roots of P_n', Newton from the Chebyshev-Lobatto points
    b.eta.assign(n1, 0.0);
    b.eta[0] = -1.0; b.eta[n] = 1.0;
    for (int j = 1; j < n; j++) {
        double x = -std::cos(M_PI * j / n);
        for (int it = 0; it < 100; it++) {
            double p, dp;
            legendre(n, x, p, dp);
            const double d2p = (2.0 * x * dp - n * (n + 1.0) * p) / (1.0 - x * x);   // Legendre ODE
            const double dx = dp / d2p;
            x -= dx;
            if (std::abs(dx) < 1e-16) break;
        }
        b.eta[j] = x;
    }
    for (int j = 0; j <= n / 2; j++) {                       // symmetrise
        const double v = 0.5 * (b.eta[j] - b.eta[n - j]);
        b.eta[j] = v; b.eta[n - j] = -v;
    }
    b.wt.resize(n1);
    for (int j = 0; j < n1; j++) {
        double p, dp;
        legendre(n, b.eta[j], p, dp);
        b.wt[j] = 2.0 / (n * (n + 1.0) * p * p);
    }
    // GLJ(0,1): interior nodes = eigenvalues of the Jacobi matrix, weights 4/(n(n+2)) / m_n^2
    {
        std::vector<double> d(n - 1), e(std::max(n - 2, 0));
        for (int i = 1; i < n; i++) d[i - 1] = 3.0 / (4.0 * (i + 0.5) * (i + 1.5));
        for (int k = 1; k < n - 1; k++) e[k - 1] = std::sqrt(k * (k + 3.0)) / (2.0 * (k + 1.5));
        const std::vector<double> inner = sym_tridiag_eigenvalues(d, e);
        b.xi_k.assign(n1, 0.0);
        b.xi_k[0] = -1.0; b.xi_k[n] = 1.0;
        for (int i = 1; i < n; i++) b.xi_k[i] = inner[i - 1];
        b.wt_axial_k.resize(n1);
        const double fact = 4.0 / (n * (n + 2.0));
        for (int j = 0; j < n1; j++) {
            const double m = vamnpo(n, b.xi_k[j]);
            b.wt_axial_k[j] = fact / (m * m);
        }
        b.wt_axial_k[0] *= 2.0;
    }
    b.G2_dp = lagrange_deriv_matrix(b.eta);
    b.G1_dp = lagrange_deriv_matrix(b.xi_k);
    b.G1.resize((size_t)n1 * n1); b.G1T.resize((size_t)n1 * n1);
    b.G2.resize((size_t)n1 * n1); b.G2T.resize((size_t)n1 * n1);
    b.G0.resize(n1);
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n1; j++) {
            // Fortran G(j,i) at [j + n1*i]; the transposes are copies of the rounded values
            b.G2[j + (size_t)n1 * i] = (float)b.G2_dp[j + (size_t)n1 * i];
            b.G1[j + (size_t)n1 * i] = (float)b.G1_dp[j + (size_t)n1 * i];
        }
    for (int i = 0; i < n1; i++)
        for (int j = 0; j < n1; j++) {
            b.G2T[i + (size_t)n1 * j] = b.G2[j + (size_t)n1 * i];
            b.G1T[i + (size_t)n1 * j] = b.G1[j + (size_t)n1 * i];
        }
    for (int j = 0; j < n1; j++) b.G0[j] = b.G1[j + (size_t)n1 * 0];      // G0(j) = G1(j,0)
    return b;
}

}  // namespace axisem
