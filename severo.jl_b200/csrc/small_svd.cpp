// small_svd.cpp — host SVD of the w x w projected matrix B of the Lanczos bidiagonalisation
// (libcell calls LAPACK dgesdd here; w = nu+7 <= ~110, so a one-sided Jacobi (Hestenes) sweep in
// plain C++ is both sufficient and free of any LAPACK dependency). B = P diag(s) Q', s descending.
//
// Cost matters once the cells are sharded over 8 GPUs (the SVD is serial host time between sweeps): the
// squared column norms are cached and updated by the rotation formulas (de Rijk), so a pair that needs no
// rotation costs one dot product; columns are pre-sorted by norm, which is the natural order of B after a
// restart (diag(sigma) | residual column | bidiagonal tail) and cuts the number of sweeps.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace svb {

void small_svd(int w, const double *A, double *P, double *s, double *Q) {
    // G = A * Vm (columns rotated from the right), Vm accumulates the right rotations
    std::vector<double> G((size_t)w * w), Vm((size_t)w * w, 0.0), n2(w);
    std::vector<int> perm(w);
    std::iota(perm.begin(), perm.end(), 0);
    for (int c = 0; c < w; ++c) {
        double a = 0.0;
        for (int i = 0; i < w; ++i) a += A[(size_t)c * w + i] * A[(size_t)c * w + i];
        n2[c] = a;
    }
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return n2[a] > n2[b]; });
    for (int c = 0; c < w; ++c) {
        std::copy(A + (size_t)perm[c] * w, A + (size_t)perm[c] * w + w, G.begin() + (size_t)c * w);
        Vm[(size_t)c * w + perm[c]] = 1.0;
    }
    const double eps = 2.220446049250313e-16;
    for (int sweep = 0; sweep < 60; ++sweep) {
        // exact norms once per sweep (the update formulas drift by a few ulps per rotation)
        for (int c = 0; c < w; ++c) {
            const double *g = &G[(size_t)c * w];
            double a = 0.0;
            for (int i = 0; i < w; ++i) a += g[i] * g[i];
            n2[c] = a;
        }
        bool rotated = false;
        for (int p = 0; p < w - 1; ++p) {
            double *gp = &G[(size_t)p * w];
            for (int q = p + 1; q < w; ++q) {
                double *gq = &G[(size_t)q * w];
                double gamma = 0.0;
                for (int i = 0; i < w; ++i) gamma += gp[i] * gq[i];
                const double alpha = n2[p], beta = n2[q];
                if (gamma == 0.0 || std::fabs(gamma) <= eps * std::sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
                double *vp = &Vm[(size_t)p * w], *vq = &Vm[(size_t)q * w];
                for (int i = 0; i < w; ++i) {
                    const double a = gp[i], b = gq[i];
                    gp[i] = c * a - sn * b;
                    gq[i] = sn * a + c * b;
                }
                for (int i = 0; i < w; ++i) {
                    const double va = vp[i], vb = vq[i];
                    vp[i] = c * va - sn * vb;
                    vq[i] = sn * va + c * vb;
                }
                n2[p] = std::max(0.0, alpha - t * gamma);
                n2[q] = std::max(0.0, beta + t * gamma);
            }
        }
        if (!rotated) break;
    }
    std::vector<double> nrm(w);
    for (int c = 0; c < w; ++c) {
        double a = 0.0;
        for (int i = 0; i < w; ++i) a += G[(size_t)c * w + i] * G[(size_t)c * w + i];
        nrm[c] = std::sqrt(a);
    }
    std::vector<int> order(w);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nrm[a] > nrm[b]; });
    for (int k = 0; k < w; ++k) {
        const int c = order[k];
        s[k] = nrm[c];
        const double inv = nrm[c] > 0.0 ? 1.0 / nrm[c] : 0.0;
        for (int i = 0; i < w; ++i) {
            P[(size_t)k * w + i] = G[(size_t)c * w + i] * inv;
            Q[(size_t)k * w + i] = Vm[(size_t)c * w + i];
        }
    }
}

}  // namespace svb
