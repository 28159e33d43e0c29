// Host-side filter design and scan tables (float64 / long double), no CUDA dependency.
//
// Restates the closed-form maths behind scipy.signal.butter(N, Wn, btype, output="ba") and
// scipy.signal.lfilter_zi, which the reference calls at backend/app/pipeline.py:175-183,
// :345-353, :590-599, :1231, :1303, :1427, :1463, and the RBJ K-weighting biquads pyloudnorm
// evaluates per sample rate (call sites pipeline.py:646-648).
#pragma once
#include <cstdint>
#include <vector>

namespace mm {

constexpr int kMaxOrder = 4;            // state dimension of the widest section (de-esser BP4)

struct Ba {
    int m = 0;                          // order = state dimension; ncoef = m + 1
    double b[kMaxOrder + 1] = {0};
    double a[kMaxOrder + 1] = {0};      // a[0] == 1
};

enum BType { kLow = 0, kHigh = 1, kBand = 2 };

// scipy.signal.butter(order, wn, btype, analog=False, output="ba"); wn normalised to Nyquist.
bool butter(int order, BType bt, const double* wn, Ba* out);
// scipy.signal.iirpeak(w0, Q) (fs = 2): bandwidth w0 / Q, -3 dB gain; b = (1 - g) [1, 0, -1], a = [1, -2 g cos(pi w0), 2 g - 1]
// with g = 1 / (1 + tan(pi w0 / (2 Q))).  Returns false outside 0 < w0 < 1.
bool iirpeak(double w0, double Q, Ba* out);
// scipy.signal.lfilter_zi(b, a): steady-state DF2T state of the unit step response.
bool lfilter_zi(const Ba& f, double* zi);
// _build_linear_phase_ir (backend/app/pipeline.py:187-217): the magnitude of the target curve HP*LP*(1 + (gp-1) Hpres +
// (gm-1) Hmud) on the n_fft/2+1 grid, clipped to [1e-8, 1e8], with linear phase of (n_fft-1)/2 samples; real part of
// the inverse DFT, cast to float32.  ir receives n_fft values.
bool linear_phase_target_ir(int sr, int n_fft, float* ir);
// pyloudnorm IIRfilter coefficients: stage 0 = high shelf (+4 dB, Q 1/sqrt2, 1500 Hz),
// stage 1 = high pass (Q 0.5, 38 Hz).
Ba k_weighting_stage(int stage, double rate);

// DF2T as a state-space system  z[n] = A z[n-1] + B x[n],  y[n] = z[n-1][0] + b0 x[n]
// A = companion(a)^T (row-major m x m), B[i] = b[i+1] - a[i+1] b0.
// A section as a general state-space system  s[n] = A s[n-1] + B x[n],  y[n] = C s[n-1] + D x[n]
// (row-major A).  DF2T is the special case A = companion(a)^T, C = e_0, D = b0.
struct StateSpace {
    int m = 0;
    long double A[kMaxOrder * kMaxOrder] = {0};
    long double B[kMaxOrder] = {0};
    long double C[kMaxOrder] = {0};
    long double D = 0;
};
void df2t_realization(const Ba& f, StateSpace* out);
// Internally balanced realization (equal, diagonal controllability / observability Gramians): ||A||_2 <= 1 and
// the round-off noise gain of a float32 state update is O(1) instead of O(1/(1-r)^2) for DF2T.  T (m x m,
// row-major) maps DF2T states into the balanced coordinates, s_bal = T z.  False if the section is not minimal.
bool balanced_realization(const Ba& f, StateSpace* out, long double* T);
// A high-pass biquad b = g [1, -2, 1] as a Chamberlin state-variable filter, state (lp, bp):
//   lp' = lp + f bp;   hp = x - lp' - q bp;   bp' = bp + f hp;   y = g hp        (f^2 = 1 + a1 + a2,  f q = 1 - a2)
// -- four float32 operations per sample where the balanced realization needs seven, with states of the signal's own magnitude
// (the low-passed input and a band-pass of gain ~1/q), so a float32 update loses ~1e-7 of the SIGNAL per step instead of ~1e-7 of
// a state that is 1/(1-r) times larger as in DF2T.  False if the numerator is not a double zero at z = 1 or f^2 <= 0.
bool svf_highpass_realization(const Ba& f, StateSpace* out, double* f_out, double* q_out, double* g_out);
// y2(y1(x)): state [s1; s2]
void cascade_realization(const StateSpace& s1, const StateSpace& s2, StateSpace* out);

enum Realization { kDf2tF64 = 0, kBalancedF32 = 1 };

struct ScanTables {
    int mode = kDf2tF64;                // coordinates the tables are expressed in
    double A[kMaxOrder * kMaxOrder] = {0}, B[kMaxOrder] = {0}, C[kMaxOrder] = {0}, D = 0;   // the realization
    double norm2 = 0;                   // ||A||_2 of the balanced realization (0 if not computed)
    int m = 0;
    int S = 0, T = 0;                   // samples per thread, threads per tile
    int W = 0;                          // look-back window (tiles) after which A^(L*W) < 1e-18
    int Wh = 0;                         // halo length (tiles): every entry of A^(L*Wh) below 1e-13
    int Wq = 0;                         // the 1e-18 look-back window counted in WARP-tiles of 32 S samples (<= W * T / 32): what a
                                        // warp-autonomous kernel re-reads before a segment
    std::vector<double> g;              // [S][m]      g[j] = A^(S-1-j) B
    std::vector<double> Pw;             // [5][m*m]    (A^S)^(2^d)
    std::vector<double> Plane;          // [32][m*m]   (A^S)^l
    std::vector<double> Qpow;           // [T/32+1][m*m]  (A^(32 S))^w
    std::vector<double> Mpow;           // [W][m*m]    (A^(S*T))^j
    std::vector<double> Apow;           // [S+1][m*m]  A^j   (dead-head handling)
    double zi[kMaxOrder] = {0};
    double pole_radius = 0;
};
bool build_scan_tables(const Ba& f, int S, int T, ScanTables* out, int max_window = 4096);
// Tables of an arbitrary realization; zi (DF2T steady state of the unit step, or null) is mapped through T_map (or identity).
bool build_scan_tables_ss(const StateSpace& ss, int S, int T, ScanTables* out, int max_window = 4096);
// Same section in balanced coordinates for the float32 pass 2 (falls back to false if balancing fails).
bool build_scan_tables_balanced(const Ba& f, int S, int T, ScanTables* out, int max_window = 4096);
// ||A||_2 of the balanced realization of f (1.0 if balancing fails): the precision policy's "how long do state errors live"
double balanced_norm(const Ba& f);

}  // namespace mm
