// Spectral clustering of the scaffold components (HOST code; SURVEY.md §8f-2, first piece).
//
// Replaces spectral_clustering(connections, dims) (clustering/ReadClusteringEngine.cpp:653-697) and what it calls in
// lib/clustering: SpectralClustering::SpectralClustering (normalised affinity + eigen-decomposition, SpectralClustering.cpp:20-56),
// ClusterRotate::cluster (incremental alignment over 2 .. dims eigenvectors, ClusterRotate.cpp:22-76) and Evrot (gradient descent
// over Givens angles on the alignment cost of Zelnik-Manor & Perona, "Self-tuning spectral clustering", Evrot.cpp:40-246).
// The matrices here are S x S with S = number of scaffold components (tens to a few thousand): this is sequential host
// arithmetic in the reference and stays host arithmetic here; nothing in it is worth a kernel.
//
// Written against flat row-major arrays (the reference uses Eigen2 expression objects). The constants that decide the result
// are the reference's: exponent range 0.3 .. 20 for the affinities (:670-674), step 1.0, at most 200 sweeps, stop when the
// quality gained over two sweeps is below 1e-3 (Evrot.cpp:44,88,105-109), a dimension count is kept when its quality is within
// 1e-3 of the best so far (ClusterRotate.cpp:38-45), members ordered by their distance to the cluster centre (:62-74; element [0]
// becomes the surviving component id in merge_components). The eigen-solver is Householder tridiagonalisation + implicit QL -
// the same routine as in the Eigen2 stand-in the oracle build compiles the reference's sources against
// (oracle/shim/eigen2/Eigen/Core), checked on its own against numpy.linalg.eigh; against a real Eigen2 build eigenvectors agree up
// to sign and rotation inside degenerate eigenspaces.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "../../include/hga_b200.h"

void hga_set_error(const char *fmt, ...);

namespace {

struct Mat {
    int r = 0, c = 0;
    std::vector<double> v;
    Mat() = default;
    Mat(int rows, int cols) : r(rows), c(cols), v((size_t) rows * cols, 0.0) {}
    double &at(int i, int j) { return v[(size_t) i * c + j]; }
    double at(int i, int j) const { return v[(size_t) i * c + j]; }
};

Mat matmul(const Mat &a, const Mat &b) {
    Mat o(a.r, b.c);
    for (int i = 0; i < a.r; i++)
        for (int k = 0; k < a.c; k++) {
            const double x = a.at(i, k);
            if (x == 0.0) continue;
            for (int j = 0; j < b.c; j++) o.at(i, j) += x * b.at(k, j);
        }
    return o;
}

// Symmetric eigen-decomposition: Householder reduction to tridiagonal form, then implicit-shift QL (the classical EISPACK
// tred2 / tql2 pair). a (n x n, row major) holds the matrix on entry and the eigenvectors (columns) on exit; d the eigenvalues,
// in no particular order. Returns false when an eigenvalue does not converge in 60 iterations.
static bool sym_eigen_tridiagonal_ql(int n, std::vector<double> &a, std::vector<double> &d) {
    auto A = [&](int i, int j) -> double & { return a[(size_t) i * n + j]; };
    std::vector<double> e((size_t) n, 0.0);
    d.assign((size_t) n, 0.0);
    for (int i = n - 1; i >= 1; i--) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; k++) scale += std::fabs(A(i, k));
            if (scale == 0.0) {
                e[i] = A(i, l);
            } else {
                for (int k = 0; k <= l; k++) { A(i, k) /= scale; h += A(i, k) * A(i, k); }
                double f = A(i, l);
                double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                A(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j <= l; j++) {
                    A(j, i) = A(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; k++) g += A(j, k) * A(i, k);
                    for (int k = j + 1; k <= l; k++) g += A(k, j) * A(i, k);
                    e[j] = g / h;
                    f += e[j] * A(i, j);
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; j++) {
                    f = A(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; k++) A(j, k) -= f * e[k] + g * A(i, k);
                }
            }
        } else {
            e[i] = A(i, l);
        }
        d[i] = h;
    }
    if (n > 0) { d[0] = 0.0; e[0] = 0.0; }
    for (int i = 0; i < n; i++) {
        const int l = i - 1;
        if (d[i] != 0.0) {
            for (int j = 0; j <= l; j++) {
                double g = 0.0;
                for (int k = 0; k <= l; k++) g += A(i, k) * A(k, j);
                for (int k = 0; k <= l; k++) A(k, j) -= g * A(k, i);
            }
        }
        d[i] = A(i, i);
        A(i, i) = 1.0;
        for (int j = 0; j <= l; j++) A(j, i) = A(i, j) = 0.0;
    }
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    if (n > 0) e[n - 1] = 0.0;
    for (int l = 0; l < n; l++) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; m++) {
                const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 60) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; i--) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = r = std::hypot(f, g);
                    if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
                    s = f / r;
                    c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    for (int k = 0; k < n; k++) {
                        f = A(k, i + 1);
                        A(k, i + 1) = s * A(k, i) + c * f;
                        A(k, i) = c * A(k, i) - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p;
                e[l] = g;
                e[m] = 0.0;
            }
        } while (m != l);
    }
    return true;
}

// eigenvalues ascending (stable), eigenvectors in the columns of vec
bool sym_eigen(const Mat &A0, std::vector<double> &val, Mat &vec) {
    const int n = A0.r;
    std::vector<double> a = A0.v, d;
    const bool converged = sym_eigen_tridiagonal_ql(n, a, d);
    std::vector<int> order((size_t) n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return d[x] < d[y]; });
    val.assign((size_t) n, 0.0);
    vec = Mat(n, n);
    for (int j = 0; j < n; j++) { val[j] = d[order[j]]; for (int i = 0; i < n; i++) vec.at(i, j) = a[(size_t) i * n + order[j]]; }
    return converged;
}

// The alignment of one set of eigenvectors X (n x d): angles of the d (d - 1) / 2 Givens rotations that make every row of X R as
// axis-aligned as possible (Evrot, method 1 = true derivative).
struct Aligner {
    const Mat &X;
    const int n, d, n_angles;
    std::vector<int> ik, jk;
    Mat rotated;
    double quality = 0;
    std::vector<std::vector<int>> clusters;

    explicit Aligner(const Mat &x) : X(x), n(x.r), d(x.c), n_angles(x.c * (x.c - 1) / 2), clusters((size_t) x.c) {
        for (int i = 0; i < d - 1; i++) for (int j = i + 1; j <= d - 1; j++) { ik.push_back(i); jk.push_back(j); }   // upper triangle, row by row
        run();
    }

    // product of the Givens rotations a .. b (identity when b < a)
    Mat rotation(const std::vector<double> &theta, int a, int b) const {
        Mat U(d, d);
        for (int i = 0; i < d; i++) U.at(i, i) = 1.0;
        for (int k = a; k <= b; k++) {
            const double cs = std::cos(theta[k]), sn = std::sin(theta[k]);
            for (int i = 0; i < d; i++) {
                const double u = U.at(i, ik[k]) * cs - U.at(i, jk[k]) * sn;
                U.at(i, jk[k]) = U.at(i, ik[k]) * sn + U.at(i, jk[k]) * cs;
                U.at(i, ik[k]) = u;
            }
        }
        return U;
    }
    Mat rotate(const std::vector<double> &theta) const { return matmul(X, rotation(theta, 0, n_angles - 1)); }

    // 1 - (mean over rows of sum_j (x_ij / max_j |x_ij|)^2 - 1) / d
    double cost(const Mat &Y) const {
        double total = 0;
        for (int i = 0; i < n; i++) {
            double mx = Y.at(i, 0) * Y.at(i, 0);
            for (int j = 1; j < d; j++) mx = std::max(mx, Y.at(i, j) * Y.at(i, j));
            for (int j = 0; j < d; j++) total += (Y.at(i, j) * Y.at(i, j)) / mx;
        }
        return 1.0 - (total / n - 1.0) / d;
    }

    double gradient(const std::vector<double> &theta, int k) const {
        Mat V(d, d);
        V.at(ik[k], ik[k]) = -std::sin(theta[k]);
        V.at(ik[k], jk[k]) = std::cos(theta[k]);
        V.at(jk[k], ik[k]) = -std::cos(theta[k]);
        V.at(jk[k], jk[k]) = -std::sin(theta[k]);
        const Mat A = matmul(matmul(matmul(X, rotation(theta, 0, k - 1)), V), rotation(theta, k + 1, n_angles - 1));
        const Mat Y = rotate(theta);
        std::vector<double> mv(n);
        std::vector<int> mc(n);
        for (int i = 0; i < n; i++) {
            int best = 0;
            for (int j = 1; j < d; j++) if (std::fabs(Y.at(i, j)) > std::fabs(Y.at(i, best))) best = j;
            mv[i] = Y.at(i, best); mc[i] = best;
        }
        double dJ = 0;
        for (int j = 0; j < d; j++)
            for (int i = 0; i < n; i++) {
                const double t1 = A.at(i, j) * Y.at(i, j) / (mv[i] * mv[i]);
                const double t2 = A.at(i, mc[i]) * (Y.at(i, j) * Y.at(i, j)) / (mv[i] * mv[i] * mv[i]);
                dJ += t1 - t2;
            }
        return 2 * dJ / n / d;
    }

    void run() {
        std::vector<double> theta((size_t) n_angles, 0.0), trial((size_t) n_angles, 0.0);
        double Q = cost(X), q1 = Q, q2 = Q;
        for (int iter = 1; iter <= 200; iter++) {
            for (int k = 0; k < n_angles; k++) {
                trial[k] = theta[k] - 1.0 * gradient(theta, k);
                const double qn = cost(rotate(trial));
                if (qn > Q) { theta[k] = trial[k]; Q = qn; } else trial[k] = theta[k];
            }
            if (iter > 2 && Q - q2 < 1e-3) break;
            q2 = q1; q1 = Q;
        }
        rotated = rotate(trial);
        for (int i = 0; i < n; i++) {
            int best = 0;
            for (int j = 1; j < d; j++) if (std::fabs(rotated.at(i, j)) > std::fabs(rotated.at(i, best))) best = j;
            clusters[(size_t) best].push_back(i);
        }
        quality = Q;
    }
};

}  // namespace

extern "C" int hga_spectral_clustering(const uint32_t *conn_x, const uint32_t *conn_y, const uint64_t *conn_score, uint64_t n_conn, int dims,
                                       uint32_t *out_component, uint64_t *out_cluster_off, uint64_t *out_n_components, uint64_t *out_n_clusters) {
    if (!out_component || !out_cluster_off || !out_n_components || !out_n_clusters || (n_conn && (!conn_x || !conn_y || !conn_score))) {
        hga_set_error("hga_spectral_clustering: NULL argument");
        return HGA_E_ARG;
    }
    *out_n_components = 0; *out_n_clusters = 0; out_cluster_off[0] = 0;
    if (n_conn == 0 || dims < 2) return HGA_OK;
    // node numbering in order of first appearance, x before y (:657-663)
    std::unordered_map<uint32_t, int> id_of;
    std::vector<uint32_t> comp_of;
    auto node = [&](uint32_t c) { auto it = id_of.find(c); if (it != id_of.end()) return it->second; const int id = (int) comp_of.size(); id_of.emplace(c, id); comp_of.push_back(c); return id; };
    std::vector<std::pair<int, int>> edges(n_conn);
    uint64_t smax = conn_score[0], smin = conn_score[0];
    for (uint64_t i = 0; i < n_conn; i++) {
        const int a = node(conn_x[i]), b = node(conn_y[i]);
        edges[i] = {a, b};
        smax = std::max(smax, conn_score[i]); smin = std::min(smin, conn_score[i]);
    }
    const int S = (int) comp_of.size();
    // affinity exp(0.3 .. 20) (:670-683), normalised D^-1/2 W D^-1/2 (SpectralClustering.cpp:26-31)
    Mat W(S, S);
    for (uint64_t i = 0; i < n_conn; i++) {
        const double scaled = ((double) (20 - 0.3) * (double) (conn_score[i] - smin)) / (double) (smax - smin) + 0.3;
        W.at(edges[i].first, edges[i].second) = std::exp(scaled);
        W.at(edges[i].second, edges[i].first) = std::exp(scaled);
    }
    std::vector<double> deg(S);
    for (int i = 0; i < S; i++) { double s = 0; for (int j = 0; j < S; j++) s += W.at(i, j); deg[i] = 1 / std::sqrt(s); }
    Mat L(S, S);
    for (int i = 0; i < S; i++) for (int j = 0; j < S; j++) L.at(i, j) = deg[i] * W.at(i, j) * deg[j];
    std::vector<double> val;
    Mat vec;
    // A matrix of NaNs (all scores equal: 0 / 0 above) never converges; the reference carries on with what its solver left behind,
    // every quality comparison below then fails and no cluster is formed. Same here: the return value is not an error.
    (void) sym_eigen(L, val, vec);
    for (int i = 0; i < S - 1; i++) {                                  // largest eigenvalue first (selection, :38-46)
        int k = 0;
        for (int j = 1; j < S - i; j++) if (val[i + j] > val[i + k]) k = j;
        if (k > 0) { std::swap(val[i], val[i + k]); for (int r = 0; r < S; r++) std::swap(vec.at(r, i), vec.at(r, i + k)); }
    }
    const int D = std::min(S, dims);
    Mat X(S, D);
    for (int i = 0; i < S; i++) for (int j = 0; j < D; j++) X.at(i, j) = vec.at(i, j);

    // incremental alignment: 2, 3, ..., D eigenvectors, each step starting from the previous step's rotated vectors
    double best_quality = 0;
    std::vector<std::vector<int>> clusters;
    Mat best_rot, in(S, 2);
    for (int i = 0; i < S; i++) { in.at(i, 0) = X.at(i, 0); in.at(i, 1) = X.at(i, 1); }
    Mat prev_rot;
    for (int g = 2; g <= D; g++) {
        if (g > 2) {
            in = Mat(S, g);
            for (int i = 0; i < S; i++) { for (int j = 0; j < g - 1; j++) in.at(i, j) = prev_rot.at(i, j); in.at(i, g - 1) = X.at(i, g - 1); }
        }
        Aligner e(in);
        if (e.quality > best_quality) best_quality = e.quality;
        if (e.quality > best_quality || best_quality - e.quality <= 0.001) { clusters = e.clusters; best_rot = e.rotated; }   // prefer more clusters
        prev_rot = e.rotated;
    }
    // members ordered by their distance to the cluster centre, ties in point order (std::multimap, ClusterRotate.cpp:62-74)
    uint64_t at = 0, n_clusters = 0;
    for (auto &cl : clusters) {
        const int dcols = best_rot.c;
        std::vector<double> centre((size_t) dcols, 0.0);
        for (int p : cl) for (int j = 0; j < dcols; j++) centre[j] += best_rot.at(p, j);
        for (int j = 0; j < dcols; j++) centre[j] = centre[j] / (double) cl.size();
        std::vector<std::pair<double, int>> byd;
        for (int p : cl) {
            double d2 = 0;
            for (int j = 0; j < dcols; j++) { const double t = best_rot.at(p, j) - centre[j]; d2 += t * t; }
            byd.push_back({d2, p});
        }
        std::stable_sort(byd.begin(), byd.end(), [](const std::pair<double, int> &a, const std::pair<double, int> &b) { return a.first < b.first; });
        for (auto &e : byd) out_component[at++] = comp_of[(size_t) e.second];
        out_cluster_off[++n_clusters] = at;                              // empty clusters are kept (an unused rotated dimension)
    }
    *out_n_components = (uint64_t) S;
    *out_n_clusters = n_clusters;
    return HGA_OK;
}


// The eigen-solver on its own (exposed so that it can be checked against an independent implementation): a = symmetric n x n, row
// major; eigenvalues ascending in val[n], eigenvectors in the columns of vec (n x n, row major).
extern "C" int hga_host_sym_eigen(int n, const double *a, double *val, double *vec) {
    if (n < 0 || (n && (!a || !val || !vec))) { hga_set_error("hga_host_sym_eigen: bad argument"); return HGA_E_ARG; }
    Mat A(n, n), V;
    std::copy(a, a + (size_t) n * n, A.v.begin());
    std::vector<double> w;
    if (!sym_eigen(A, w, V)) { hga_set_error("hga_host_sym_eigen: no convergence"); return HGA_E_STATE; }
    std::copy(w.begin(), w.end(), val);
    std::copy(V.v.begin(), V.v.end(), vec);
    return HGA_OK;
}
