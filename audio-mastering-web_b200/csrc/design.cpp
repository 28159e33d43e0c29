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

bool iirpeak(double w0, double Q, Ba* out) {
    if (!(w0 > 0.0 && w0 < 1.0) || !(Q > 0.0)) return false;
    const double pi = 3.14159265358979323846;
    const double bw = (w0 / Q) * pi, w = w0 * pi;
    // scipy 1.18 (_design_notch_peak_filter): beta = math.tan(bw / 2) -- the -3 dB factor sqrt(1 - gb^2) / gb is taken as exactly 1
    // (older releases multiplied by its float64 value 1.0000000000000002); plain double arithmetic and libm tan / cos, so that
    // the degenerate cases st_dynamic_eq classifies (a summing to 0.0 in float64) fall where scipy's own do
    const double beta = std::tan(bw / 2.0);
    const double gain = 1.0 / (1.0 + beta);
    Ba f;
    f.m = 2;
    f.b[0] = 1.0 - gain; f.b[1] = 0.0; f.b[2] = -(1.0 - gain);
    f.a[0] = 1.0; f.a[1] = -2.0 * gain * std::cos(w); f.a[2] = 2.0 * gain - 1.0;
    *out = f;
    return true;
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

static std::complex<long double> freq_resp(const Ba& f, long double w) {
    std::complex<long double> num(0, 0), den(0, 0);
    for (int i = 0; i <= f.m; ++i) {
        const std::complex<long double> e = std::polar(1.0L, -w * (long double)i);
        num += (long double)f.b[i] * e;
        den += (long double)f.a[i] * e;
    }
    return num / den;
}

bool linear_phase_target_ir(int sr, int n_fft, float* ir) {
    if (sr <= 0 || n_fft < 8 || (n_fft & 1)) return false;
    const long double pi = 3.14159265358979323846264338327950288L;
    const double nyq = sr / 2.0;
    Ba hp, lp, pres, mud;
    double w1[2];
    w1[0] = std::min(40.0 / nyq, 0.99);
    if (!butter(2, kHigh, w1, &hp)) return false;
    w1[0] = std::min(18000.0 / nyq, 0.99);
    if (!butter(2, kLow, w1, &lp)) return false;
    const double fp = std::min(3000.0 / nyq, 0.99), fm = std::min(300.0 / nyq, 0.99);
    double wb[2] = {fp * 0.7, fp * 1.3};
    if (!butter(1, kBand, wb, &pres)) return false;
    wb[0] = fm * 0.7; wb[1] = fm * 1.3;
    if (!butter(1, kBand, wb, &mud)) return false;
    const long double gp = powl(10.0L, 0.35L / 20.0L), gm = powl(10.0L, -0.25L / 20.0L);
    const int N = n_fft, H = N / 2;
    std::vector<long double> mag(H + 1);
    for (int k = 0; k <= H; ++k) {
        const long double w = pi * (long double)k / (long double)H;
        const std::complex<long double> Hc = freq_resp(hp, w) * freq_resp(lp, w) *
            (std::complex<long double>(1, 0) + (gp - 1.0L) * freq_resp(pres, w) + (gm - 1.0L) * freq_resp(mud, w));
        long double m = std::abs(Hc);
        mag[k] = fminl(fmaxl(m, 1e-8L), 1e8L);
    }
    // H_full[k] = mag[k] e^{i phi_k}, phi_k = -2 pi k (N-1) / (2N); H_full[N-k] = conj; H_full[N/2] = its real part.
    // ir[n] = Re ifft = (1/N) [ mag0 + 2 sum_{k=1}^{H-1} mag_k cos(2 pi k n / N + phi_k) + mag_H cos(phi_H) cos(pi n) ]
    for (int n = 0; n < N; ++n) {
        long double acc = mag[0];
        for (int k = 1; k < H; ++k) {
            // 2 pi k n / N + phi_k = 2 pi k (n - (N-1)/2) / N = pi k (2n - N + 1) / N; reduce the integer product mod 2N
            const long long t = ((long long)k * (long long)(2 * n - N + 1)) % (2LL * N);
            acc += 2.0L * mag[k] * cosl(pi * (long double)t / (long double)N);
        }
        const long long tH = ((long long)H * (long long)(N - 1)) % (2LL * N);
        const long double re_nyq = mag[H] * cosl(-pi * (long double)tH / (long double)N);
        acc += re_nyq * ((n & 1) ? -1.0L : 1.0L);
        ir[n] = (float)(double)(acc / (long double)N);
    }
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

void df2t_realization(const Ba& f, StateSpace* out) {
    out->m = f.m;
    state_matrices(f, out->A, out->B);
    for (int i = 0; i < f.m; ++i) out->C[i] = (i == 0) ? 1.0L : 0.0L;
    out->D = (long double)f.b[0];
}

bool svf_highpass_realization(const Ba& ba, StateSpace* out, double* f_out, double* q_out, double* g_out) {
    if (ba.m != 2) return false;
    const long double g = ba.b[0];
    if (g == 0.0L || std::fabs((double)(ba.b[1] + 2.0L * g)) > 1e-12 * std::fabs((double)g) ||
        std::fabs((double)(ba.b[2] - g)) > 1e-12 * std::fabs((double)g))
        return false;
    const long double f2 = 1.0L + (long double)ba.a[1] + (long double)ba.a[2];
    if (!(f2 > 0.0L)) return false;
    const long double f = std::sqrt(f2), q = (1.0L - (long double)ba.a[2]) / f;
    StateSpace r;
    r.m = 2;
    // s[n] = A s[n-1] + B x[n], y[n] = C s[n-1] + D x[n] with s = (lp, bp)
    r.A[0] = 1.0L;  r.A[1] = f;
    r.A[2] = -f;    r.A[3] = 1.0L - f * (f + q);
    r.B[0] = 0.0L;  r.B[1] = f;
    r.C[0] = -g;    r.C[1] = -g * (f + q);
    r.D = g;
    *out = r;
    if (f_out) *f_out = (double)f;
    if (q_out) *q_out = (double)q;
    if (g_out) *g_out = (double)g;
    return true;
}

void cascade_realization(const StateSpace& s1, const StateSpace& s2, StateSpace* out) {
    const int m1 = s1.m, m2 = s2.m, m = m1 + m2;
    StateSpace r;
    r.m = m;
    for (int i = 0; i < m * m; ++i) r.A[i] = 0.0L;
    for (int i = 0; i < m1; ++i) {
        for (int j = 0; j < m1; ++j) r.A[i * m + j] = s1.A[i * m1 + j];
        r.B[i] = s1.B[i];
        r.C[i] = s2.D * s1.C[i];
    }
    for (int i = 0; i < m2; ++i) {
        for (int j = 0; j < m1; ++j) r.A[(m1 + i) * m + j] = s2.B[i] * s1.C[j];     // the second section sees y1 = C1 s1 + D1 x
        for (int j = 0; j < m2; ++j) r.A[(m1 + i) * m + m1 + j] = s2.A[i * m2 + j];
        r.B[m1 + i] = s2.B[i] * s1.D;
        r.C[m1 + i] = s2.C[i];
    }
    r.D = s2.D * s1.D;
    *out = r;
}

// X = A X A^T + Q  (discrete Lyapunov), m <= 4: (I - A (x) A) vec(X) = vec(Q)
static bool solve_dlyap(int m, const long double* A, const long double* Q, long double* X) {
    const int n = m * m;
    long double K[256], rhs[16];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) {
            const int r = i * m + j;
            for (int k = 0; k < m; ++k)
                for (int l = 0; l < m; ++l) K[r * n + k * m + l] = ((r == k * m + l) ? 1.0L : 0.0L) - A[i * m + k] * A[j * m + l];
            rhs[r] = Q[r];
        }
    if (!solve_small(n, K, rhs)) return false;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) X[i * m + j] = 0.5L * (rhs[i * m + j] + rhs[j * m + i]);
    return true;
}

// cyclic Jacobi eigen-decomposition of a symmetric m x m matrix: S = U diag(ev) U^T
static void jacobi_eig(int m, long double* S, long double* U, long double* ev) {
    mat_eye(m, U);
    for (int sweep = 0; sweep < 60; ++sweep) {
        long double off = 0.0L, diag = 0.0L;
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) (i == j ? diag : off) += S[i * m + j] * S[i * m + j];
        if (off <= 1e-40L * diag || off == 0.0L) break;
        for (int p = 0; p < m; ++p)
            for (int q = p + 1; q < m; ++q) {
                if (S[p * m + q] == 0.0L) continue;
                const long double theta = (S[q * m + q] - S[p * m + p]) / (2.0L * S[p * m + q]);
                const long double t = (theta >= 0 ? 1.0L : -1.0L) / (fabsl(theta) + sqrtl(theta * theta + 1.0L));
                const long double cs = 1.0L / sqrtl(t * t + 1.0L), sn = t * cs;
                for (int k = 0; k < m; ++k) {      // columns p, q
                    const long double a = S[k * m + p], b = S[k * m + q];
                    S[k * m + p] = cs * a - sn * b;
                    S[k * m + q] = sn * a + cs * b;
                }
                for (int k = 0; k < m; ++k) {      // rows p, q
                    const long double a = S[p * m + k], b = S[q * m + k];
                    S[p * m + k] = cs * a - sn * b;
                    S[q * m + k] = sn * a + cs * b;
                }
                for (int k = 0; k < m; ++k) {
                    const long double a = U[k * m + p], b = U[k * m + q];
                    U[k * m + p] = cs * a - sn * b;
                    U[k * m + q] = sn * a + cs * b;
                }
            }
    }
    for (int i = 0; i < m; ++i) ev[i] = S[i * m + i];
}

bool balanced_realization(const Ba& f, StateSpace* out, long double* Tmap) {
    StateSpace d;
    df2t_realization(f, &d);
    const int m = d.m;
    if (m < 1 || m > kMaxOrder) return false;
    long double Q[16], At[16], Wc[16], Wo[16];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) { Q[i * m + j] = d.B[i] * d.B[j]; At[i * m + j] = d.A[j * m + i]; }
    if (!solve_dlyap(m, d.A, Q, Wc)) return false;
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Q[i * m + j] = d.C[i] * d.C[j];
    if (!solve_dlyap(m, At, Q, Wo)) return false;
    // Cholesky Wc = L L^T
    long double L[16];
    for (int i = 0; i < m * m; ++i) L[i] = 0.0L;
    for (int j = 0; j < m; ++j) {
        long double s = Wc[j * m + j];
        for (int k = 0; k < j; ++k) s -= L[j * m + k] * L[j * m + k];
        if (!(s > 0.0L)) return false;
        L[j * m + j] = sqrtl(s);
        for (int i = j + 1; i < m; ++i) {
            long double t = Wc[i * m + j];
            for (int k = 0; k < j; ++k) t -= L[i * m + k] * L[j * m + k];
            L[i * m + j] = t / L[j * m + j];
        }
    }
    // S = L^T Wo L = U diag(hsv^2) U^T
    long double Lt[16], S[16], U[16], ev[4], tmp[16];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Lt[i * m + j] = L[j * m + i];
    mat_mul(m, Lt, Wo, tmp);
    mat_mul(m, tmp, L, S);
    for (int i = 0; i < m; ++i)
        for (int j = i + 1; j < m; ++j) S[i * m + j] = S[j * m + i] = 0.5L * (S[i * m + j] + S[j * m + i]);
    jacobi_eig(m, S, U, ev);
    for (int i = 0; i < m; ++i)
        if (!(ev[i] > 0.0L)) return false;
    // Tinv = L U diag(ev^-1/4),  T = diag(ev^1/4) U^T L^-1
    long double Tinv[16], T[16], Linv[16];
    for (int c = 0; c < m; ++c) {          // L^-1 by forward substitution on unit vectors
        for (int i = 0; i < m; ++i) {
            long double s = (i == c) ? 1.0L : 0.0L;
            for (int k = 0; k < i; ++k) s -= L[i * m + k] * Linv[k * m + c];
            Linv[i * m + c] = s / L[i * m + i];
        }
    }
    mat_mul(m, L, U, Tinv);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Tinv[i * m + j] *= powl(ev[j], -0.25L);
    long double Ut[16];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) Ut[i * m + j] = U[j * m + i];
    mat_mul(m, Ut, Linv, T);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) T[i * m + j] *= powl(ev[i], 0.25L);
    StateSpace r;
    r.m = m;
    mat_mul(m, T, d.A, tmp);
    mat_mul(m, tmp, Tinv, r.A);
    for (int i = 0; i < m; ++i) {
        long double sb = 0.0L, sc = 0.0L;
        for (int k = 0; k < m; ++k) { sb += T[i * m + k] * d.B[k]; sc += d.C[k] * Tinv[k * m + i]; }
        r.B[i] = sb;
        r.C[i] = sc;
    }
    r.D = d.D;
    *out = r;
    if (Tmap) for (int i = 0; i < m * m; ++i) Tmap[i] = T[i];
    return true;
}

static double spectral_norm(int m, const long double* A) {
    long double At[16], S[16], U[16], ev[4];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) At[i * m + j] = A[j * m + i];
    mat_mul(m, At, A, S);
    jacobi_eig(m, S, U, ev);
    long double mx = 0.0L;
    for (int i = 0; i < m; ++i) mx = fmaxl(mx, ev[i]);
    return (double)sqrtl(mx);
}

double balanced_norm(const Ba& f) {
    StateSpace s;
    if (!balanced_realization(f, &s, nullptr)) return 1.0;
    return spectral_norm(s.m, s.A);
}

bool build_scan_tables_ss(const StateSpace& ss, int S, int T, ScanTables* out, int max_window) {
    const int m = ss.m;
    if (m < 1 || m > kMaxOrder || T % 32 != 0) return false;
    const long double* A = ss.A;
    const long double* B = ss.B;
    out->m = m; out->S = S; out->T = T;
    const int mm2 = m * m;
    for (int i = 0; i < mm2; ++i) out->A[i] = (double)ss.A[i];
    for (int i = 0; i < m; ++i) { out->B[i] = (double)ss.B[i]; out->C[i] = (double)ss.C[i]; }
    out->D = (double)ss.D;
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
    // the same criterion at warp-tile granularity: first j with every entry of (A^(32 S))^j AND of the next power below 1e-18
    // (an entry of a rotating state-transfer matrix may pass through zero; two consecutive powers cannot both by accident)
    {
        long double Y[kMaxOrder * kMaxOrder];
        mat_eye(m, Y);
        int wq = 0, below = 0;
        for (int j = 0; j <= W * NW + 1; ++j) {
            long double mx = 0.0L;
            for (int i = 0; i < mm2; ++i) mx = fmaxl(mx, fabsl(Y[i]));
            below = (j > 0 && mx < 1e-18L) ? below + 1 : 0;
            if (below == 2) { wq = j - 1; break; }
            wq = j;
            mat_mul(m, Y, Q, Y);
        }
        out->Wq = std::max(1, std::min(wq, W * NW));
    }
    if (out->Wh == 0) out->Wh = W;
    for (int i = 0; i < m; ++i) out->zi[i] = 0.0;
    // spectral radius estimate from M's decay: r ~ (max|M|)^(1/L)
    {
        long double mx = 0.0L;
        for (int i = 0; i < mm2; ++i) mx = fmaxl(mx, fabsl(M[i]));
        out->pole_radius = mx > 0 ? (double)expl(logl(mx) / (long double)((int64_t)S * T)) : 0.0;
    }
    return true;
}

bool build_scan_tables(const Ba& f, int S, int T, ScanTables* out, int max_window) {
    StateSpace ss;
    if (f.m < 1 || f.m > kMaxOrder) return false;
    df2t_realization(f, &ss);
    if (!build_scan_tables_ss(ss, S, T, out, max_window)) return false;
    out->mode = kDf2tF64;
    if (!lfilter_zi(f, out->zi)) {
        for (int i = 0; i < f.m; ++i) out->zi[i] = 0.0;   // scipy would raise LinAlgError; callers decide
    }
    return true;
}

bool build_scan_tables_balanced(const Ba& f, int S, int T, ScanTables* out, int max_window) {
    StateSpace ss;
    long double Tm[kMaxOrder * kMaxOrder];
    if (f.m < 1 || f.m > kMaxOrder) return false;
    if (!balanced_realization(f, &ss, Tm)) return false;
    if (!build_scan_tables_ss(ss, S, T, out, max_window)) return false;
    // the float32 pass 2 runs the realization with its states rescaled to B = (1, ..., 1) (common.cuh ss32_step): every state
    // must be driven by the input
    {
        long double bmax = 0.0L;
        for (int i = 0; i < ss.m; ++i) bmax = std::max(bmax, fabsl(ss.B[i]));
        for (int i = 0; i < ss.m; ++i) if (!(fabsl(ss.B[i]) > 1e-9L * bmax)) return false;
    }
    out->mode = kBalancedF32;
    out->norm2 = spectral_norm(ss.m, ss.A);
    // steady state of the unit-step response in these coordinates: (I - A) s = B
    {
        const int m = ss.m;
        long double I_A[kMaxOrder * kMaxOrder], rhs[kMaxOrder];
        for (int i = 0; i < m; ++i) {
            for (int j = 0; j < m; ++j) I_A[i * m + j] = (i == j ? 1.0L : 0.0L) - ss.A[i * m + j];
            rhs[i] = ss.B[i];
        }
        if (solve_small(m, I_A, rhs))
            for (int i = 0; i < m; ++i) out->zi[i] = (double)rhs[i];
    }
    return true;
}

}  // namespace mm
