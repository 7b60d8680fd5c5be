// Host-side linear algebra the C-ABI needs once per model (never per site):
// the symmetric eigenproblem behind pb_set_rate_matrix.  One n x n problem,
// n <= 64: stays on the host like Matrix.diagonalize (lib/linear_algebra.ml:238-241).
#pragma once
#include <string>

namespace pbh {

// Cyclic Jacobi eigen-decomposition of a symmetric matrix a (row-major, n x n,
// destroyed): eigenvalues in w, eigenvectors as COLUMNS of v (row-major).
void jacobi_eigh(int n, double* a, double* w, double* v);

// Q reversible w.r.t. pi (pi_i q_ij = pi_j q_ji): Q = U diag(lambda) Uinv with
// U = D^-1/2 V, Uinv = V^T D^1/2, V from the symmetrised S = D^1/2 Q D^-1/2.
// All matrices column-major (the ABI convention).  false + *why on failure.
bool reversible_eigensystem(int n, const double* Q, const double* pi, double* U, double* lambda,
                            double* Uinv, std::string* why);

}  // namespace pbh
