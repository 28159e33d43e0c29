#include "design.h"

#include <cmath>
#include <complex>

namespace mm {

typedef std::complex<double> cd;

static std::vector<cd> poly_from_roots(const std::vector<cd>& r) {
    std::vector<cd> c(1, cd(1.0, 0.0));
    for (size_t i = 0; i < r.size(); ++i) {
        std::vector<cd> n(c.size() + 1, cd(0.0, 0.0));
        for (size_t j = 0; j < c.size(); ++j) {
            n[j] += c[j];
            n[j + 1] -= c[j] * r[i];
        }
        c.swap(n);
    }
    return c;
}

bool butter(int order, BType bt, const double* wn, Ba* out) {
    if (order < 1 || order > 2) return false;
    const double pi = 3.14159265358979323846;
    // analogue prototype: poles on the unit circle, left half plane, gain 1
    std::vector<cd> z, p;
    for (int mm_ = -order + 1; mm_ < order; mm_ += 2) p.push_back(-std::exp(cd(0.0, pi * mm_ / (2.0 * order))));
    double k = 1.0;
    const double fs = 2.0;
    const int degree = (int)p.size();
    if (bt == kLow || bt == kHigh) {
        if (!(wn[0] > 0.0 && wn[0] < 1.0)) return false;
        const double wo = 2.0 * fs * std::tan(pi * wn[0] / fs);
        if (bt == kLow) {
            for (auto& q : p) q *= wo;
            k *= std::pow(wo, degree);
        } else {
            cd prod_p(1.0, 0.0);
            for (auto& q : p) prod_p *= -q;
            for (auto& q : p) q = wo / q;
            z.assign(degree, cd(0.0, 0.0));
            k *= (cd(1.0, 0.0) / prod_p).real();
        }
    } else {
        if (!(wn[0] > 0.0 && wn[1] < 1.0 && wn[0] < wn[1])) return false;
        const double w0 = 2.0 * fs * std::tan(pi * wn[0] / fs);
        const double w1 = 2.0 * fs * std::tan(pi * wn[1] / fs);
        const double bw = w1 - w0;
        const double wo = std::sqrt(w0 * w1);
        std::vector<cd> plus, minus;
        for (auto& q : p) {
            cd lp = q * (bw / 2.0);
            cd root = std::sqrt(lp * lp - wo * wo);
            plus.push_back(lp + root);
            minus.push_back(lp - root);
        }
        p = plus;
        p.insert(p.end(), minus.begin(), minus.end());
        z.assign(degree, cd(0.0, 0.0));
        k *= std::pow(bw, degree);
    }
    // bilinear transform, fs = 2
    const double fs2 = 2.0 * fs;
    cd num(1.0, 0.0), den(1.0, 0.0);
    for (auto& q : z) num *= (fs2 - q);
    for (auto& q : p) den *= (fs2 - q);
    for (auto& q : z) q = (fs2 + q) / (fs2 - q);
    for (auto& q : p) q = (fs2 + q) / (fs2 - q);
    const size_t nz = z.size();
    for (size_t i = nz; i < p.size(); ++i) z.push_back(cd(-1.0, 0.0));
    k *= (num / den).real();
    std::vector<cd> bb = poly_from_roots(z), aa = poly_from_roots(p);
    const int m = (int)p.size();
    if (m > kMaxOrder) return false;
    out->m = m;
    for (int i = 0; i <= m; ++i) {
        out->b[i] = k * bb[i].real();
        out->a[i] = aa[i].real();
    }
    return true;
}

static bool solve_small(int m, long double* Amat, long double* rhs) {  // Gaussian elimination, partial pivoting
    for (int c = 0; c < m; ++c) {
        int piv = c;
        for (int r = c + 1; r < m; ++r)
            if (fabsl(Amat[r * m + c]) > fabsl(Amat[piv * m + c])) piv = r;
        if (fabsl(Amat[piv * m + c]) < 1e-300L) return false;
        if (piv != c) {
            for (int j = 0; j < m; ++j) std::swap(Amat[c * m + j], Amat[piv * m + j]);
            std::swap(rhs[c], rhs[piv]);
        }
        for (int r = c + 1; r < m; ++r) {
            long double f = Amat[r * m + c] / Amat[c * m + c];
            for (int j = c; j < m; ++j) Amat[r * m + j] -= f * Amat[c * m + j];
            rhs[r] -= f * rhs[c];
        }
    }
    for (int r = m - 1; r >= 0; --r) {
        long double s = rhs[r];
        for (int j = r + 1; j < m; ++j) s -= Amat[r * m + j] * rhs[j];
        rhs[r] = s / Amat[r * m + r];
    }
    return true;
}

static void state_matrices(const Ba& f, long double* A, long double* B) {
    const int m = f.m;
    for (int i = 0; i < m * m; ++i) A[i] = 0.0L;
    for (int i = 0; i < m; ++i) {
        A[i * m + 0] = -(long double)f.a[i + 1];
        if (i + 1 < m) A[i * m + i + 1] = 1.0L;
        B[i] = (long double)f.b[i + 1] - (long double)f.a[i + 1] * (long double)f.b[0];
    }
}

bool lfilter_zi(const Ba& f, double* zi) {
    const int m = f.m;
    long double A[kMaxOrder * kMaxOrder], B[kMaxOrder], I_A[kMaxOrder * kMaxOrder];
    state_matrices(f, A, B);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) I_A[i * m + j] = (i == j ? 1.0L : 0.0L) - A[i * m + j];
    if (!solve_small(m, I_A, B)) return false;
    for (int i = 0; i < m; ++i) zi[i] = (double)B[i];
    return true;
}

Ba k_weighting_stage(int stage, double rate) {
    const double pi = 3.14159265358979323846;
    Ba f;
    f.m = 2;
    double b0, b1, b2, a0, a1, a2;
    if (stage == 0) {
        const double G = 4.0, Q = 1.0 / std::sqrt(2.0), fc = 1500.0;
        const double A = std::pow(10.0, G / 40.0);
        const double w0 = 2.0 * pi * (fc / rate);
        const double alpha = std::sin(w0) / (2.0 * Q);
        const double cw = std::cos(w0), sA = std::sqrt(A);
        b0 = A * ((A + 1) + (A - 1) * cw + 2 * sA * alpha);
        b1 = -2 * A * ((A - 1) + (A + 1) * cw);
        b2 = A * ((A + 1) + (A - 1) * cw - 2 * sA * alpha);
        a0 = (A + 1) - (A - 1) * cw + 2 * sA * alpha;
        a1 = 2 * ((A - 1) - (A + 1) * cw);
        a2 = (A + 1) - (A - 1) * cw - 2 * sA * alpha;
    } else {
        const double Q = 0.5, fc = 38.0;
        const double w0 = 2.0 * pi * (fc / rate);
        const double alpha = std::sin(w0) / (2.0 * Q);
        const double cw = std::cos(w0);
        b0 = (1 + cw) / 2;
        b1 = -(1 + cw);
        b2 = (1 + cw) / 2;
        a0 = 1 + alpha;
        a1 = -2 * cw;
        a2 = 1 - alpha;
    }
    f.b[0] = b0 / a0; f.b[1] = b1 / a0; f.b[2] = b2 / a0;
    f.a[0] = 1.0;     f.a[1] = a1 / a0; f.a[2] = a2 / a0;
    return f;
}

// ---- small long-double matrix helpers ---------------------------------------------------------
static void mat_mul(int m, const long double* X, const long double* Y, long double* Z) {
    long double t[kMaxOrder * kMaxOrder];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            long double s = 0.0L;
            for (int k = 0; k < m; ++k) s += X[i * m + k] * Y[k * m + j];
            t[i * m + j] = s;
        }
    for (int i = 0; i < m * m; ++i) Z[i] = t[i];
}
static void mat_eye(int m, long double* X) {
    for (int i = 0; i < m * m; ++i) X[i] = 0.0L;
    for (int i = 0; i < m; ++i) X[i * m + i] = 1.0L;
}
static void mat_pow(int m, const long double* X, int64_t e, long double* Z) {
    long double base[kMaxOrder * kMaxOrder], acc[kMaxOrder * kMaxOrder];
    for (int i = 0; i < m * m; ++i) base[i] = X[i];
    mat_eye(m, acc);
    while (e > 0) {
        if (e & 1) mat_mul(m, acc, base, acc);
        mat_mul(m, base, base, base);
        e >>= 1;
    }
    for (int i = 0; i < m * m; ++i) Z[i] = acc[i];
}
static void put(std::vector<double>& v, size_t at, int mm_, const long double* X) {
    for (int i = 0; i < mm_; ++i) v[at + i] = (double)X[i];
}

bool build_scan_tables(const Ba& f, int S, int T, ScanTables* out, int max_window) {
    const int m = f.m;
    if (m < 1 || m > kMaxOrder || T % 32 != 0) return false;
    long double A[kMaxOrder * kMaxOrder], B[kMaxOrder];
    state_matrices(f, A, B);
    out->m = m; out->S = S; out->T = T;
    const int mm2 = m * m;
    // g[j] = A^(S-1-j) B, built backwards from j = S-1
    out->g.assign((size_t)S * m, 0.0);
    {
        long double v[kMaxOrder];
        for (int i = 0; i < m; ++i) v[i] = B[i];
        for (int j = S - 1; j >= 0; --j) {
            for (int i = 0; i < m; ++i) out->g[(size_t)j * m + i] = (double)v[i];
            long double w[kMaxOrder];
            for (int i = 0; i < m; ++i) {
                long double s = 0.0L;
                for (int k = 0; k < m; ++k) s += A[i * m + k] * v[k];
                w[i] = s;
            }
            for (int i = 0; i < m; ++i) v[i] = w[i];
        }
    }
    long double P[kMaxOrder * kMaxOrder], X[kMaxOrder * kMaxOrder];
    mat_pow(m, A, S, P);
    out->Apow.assign((size_t)(S + 1) * mm2, 0.0);
    mat_eye(m, X);
    for (int j = 0; j <= S; ++j) { put(out->Apow, (size_t)j * mm2, mm2, X); mat_mul(m, X, A, X); }
    out->Pw.assign((size_t)5 * mm2, 0.0);
    for (int i = 0; i < mm2; ++i) X[i] = P[i];
    for (int d = 0; d < 5; ++d) { put(out->Pw, (size_t)d * mm2, mm2, X); mat_mul(m, X, X, X); }
    out->Plane.assign((size_t)32 * mm2, 0.0);
    mat_eye(m, X);
    for (int l = 0; l < 32; ++l) { put(out->Plane, (size_t)l * mm2, mm2, X); mat_mul(m, X, P, X); }
    long double Q[kMaxOrder * kMaxOrder];
    mat_pow(m, P, 32, Q);
    const int NW = T / 32;
    out->Qpow.assign((size_t)(NW + 1) * mm2, 0.0);
    mat_eye(m, X);
    for (int w = 0; w <= NW; ++w) { put(out->Qpow, (size_t)w * mm2, mm2, X); mat_mul(m, X, Q, X); }
    long double M[kMaxOrder * kMaxOrder];
    mat_pow(m, Q, NW, M);
    // look-back window: stop once every entry of M^j is below 1e-18 (fp64-negligible)
    out->Mpow.clear();
    mat_eye(m, X);
    int W = 0;
    out->Wh = 0;
    for (; W < max_window; ++W) {
        long double mx = 0.0L;
        for (int i = 0; i < mm2; ++i) mx = fmaxl(mx, fabsl(X[i]));
        if (W > 0 && out->Wh == 0 && mx < 1e-13L) out->Wh = W;
        if (W > 0 && mx < 1e-18L) break;
        out->Mpow.resize((size_t)(W + 1) * mm2);
        put(out->Mpow, (size_t)W * mm2, mm2, X);
        mat_mul(m, X, M, X);
    }
    if (W >= max_window) return false;   // pole too close to the unit circle for this tile size
    out->W = W;
    if (out->Wh == 0) out->Wh = W;
    if (!lfilter_zi(f, out->zi)) {
        for (int i = 0; i < m; ++i) out->zi[i] = 0.0;   // scipy would raise LinAlgError; callers decide
    }
    // spectral radius estimate from M's decay: r ~ (max|M|)^(1/L)
    {
        long double mx = 0.0L;
        for (int i = 0; i < mm2; ++i) mx = fmaxl(mx, fabsl(M[i]));
        out->pole_radius = mx > 0 ? (double)expl(logl(mx) / (long double)((int64_t)S * T)) : 0.0;
    }
    return true;
}

}  // namespace mm
