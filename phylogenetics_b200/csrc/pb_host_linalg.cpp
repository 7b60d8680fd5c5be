#include "pb_host_linalg.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace pbh {

void jacobi_eigh(int n, double* a, double* w, double* v)
{
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) v[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < n; i++) {
            diag += a[i * n + i] * a[i * n + i];
            for (int j = i + 1; j < n; j++) off += a[i * n + j] * a[i * n + j];
        }
        if (off <= 1e-34 * (diag + off) || off == 0.0) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = a[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < n; k++) {           // A <- A J (columns p, q)
                    const double akp = a[k * n + p], akq = a[k * n + q];
                    a[k * n + p] = cs * akp - sn * akq;
                    a[k * n + q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < n; k++) {           // A <- J^T A (rows p, q)
                    const double apk = a[p * n + k], aqk = a[q * n + k];
                    a[p * n + k] = cs * apk - sn * aqk;
                    a[q * n + k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < n; k++) {           // V <- V J
                    const double vkp = v[k * n + p], vkq = v[k * n + q];
                    v[k * n + p] = cs * vkp - sn * vkq;
                    v[k * n + q] = sn * vkp + cs * vkq;
                }
            }
    }
    for (int i = 0; i < n; i++) w[i] = a[i * n + i];
}

bool reversible_eigensystem(int n, const double* Q, const double* pi, double* U, double* lambda,
                            double* Uinv, std::string* why)
{
    auto q = [&](int i, int j) { return Q[(size_t)j * n + i]; };   // column-major in
    for (int i = 0; i < n; i++)
        if (!(pi[i] > 0.0)) { if (why) *why = "stationary distribution must be strictly positive"; return false; }
    double scale = 0.0;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) scale = std::max(scale, std::fabs(pi[i] * q(i, j)));
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            if (std::fabs(pi[i] * q(i, j) - pi[j] * q(j, i)) > 1e-9 * scale) {
                if (why) *why = "rate matrix is not reversible with respect to pi (pi_i q_ij != pi_j q_ji); "
                                "use pb_set_rate_matrix_general or pb_set_pmatrices";
                return false;
            }
    std::vector<double> s((size_t)n * n), v((size_t)n * n), sq(n), isq(n);
    for (int i = 0; i < n; i++) { sq[i] = std::sqrt(pi[i]); isq[i] = 1.0 / sq[i]; }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            // symmetrise explicitly so rounding in Q cannot leave S asymmetric
            const double a = sq[i] * q(i, j) * isq[j], b = sq[j] * q(j, i) * isq[i];
            s[(size_t)i * n + j] = 0.5 * (a + b);
        }
    jacobi_eigh(n, s.data(), lambda, v.data());
    for (int i = 0; i < n; i++)
        for (int k = 0; k < n; k++) {
            U[(size_t)k * n + i] = isq[i] * v[(size_t)i * n + k];        // U(i,k)
            Uinv[(size_t)i * n + k] = v[(size_t)i * n + k] * sq[i];      // Uinv(k,i)
        }
    return true;
}

}  // namespace pbh
