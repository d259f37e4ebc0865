// racinglmpc_b200/csrc/ftocp_pdip.cuh
//
// One warp = one finite-time optimal control QP (the FTOCP that the reference assembles in
// src/fnc/controller/PredictiveControllers.py:166-257 (MPC) and :340-362 (LMPC terminal set)
// and hands to OSQP at :259-283).  Nothing is assembled: the solver works directly on the
// stage data (A_k,B_k,C_k, x(t), u(t-1), SS_sel, Qfun_sel) and solves
//
//   min  sum_k (x_k-xRef)'Q(x_k-xRef) + (x_N-xRef)'Qf(x_N-xRef) + u_k'R u_k
//        + sum_k (u_k-u_{k-1})'dR(u_k-u_{k-1}) + qs_quad|s|^2 + qs_lin 1's + Qfun'lam + xi'Qts xi
//   s.t. x_{k+1} = A_k x_k + B_k u_k + C_k,  Fx x_k - s_k <= bx (k<N),  Fu u_k <= bu,
//        s >= 0,  lam >= 0,  x_N - SS lam + xi = 0,  1'lam = 1
//
// with a Mehrotra predictor-corrector primal-dual interior-point method.  The Newton system is
// solved by (i) analytic elimination of the lane slacks and bound multipliers, (ii) a 6x6
// covariance-form elimination of the simplex/terminal block (centred at the D^-1-weighted
// centroid so the simplex multiplier decouples), (iii) a Riccati recursion over the horizon
// with the input-rate coupling carried as a 2-dim augmented state.  All in IEEE fp64.
// Executable specification + derivation: oracle/pdip_model.py (NumPy, test-only).
//
// Execution model: every function below is called by all 32 lanes of a warp with identical
// arguments.  Three idioms only, so that the same source also compiles as a 1-lane host
// emulation (tests/host_core.cpp; never part of the product path):
//   FOR_LANES(e, n) {..}  +  wsync()     work items spread over lanes, shared-memory results
//   lane-redundant scalar code           every lane computes the same value (free under SIMT)
//   FOR_SLOTS(r,row,CNT) + wsum/wmin     per-constraint state held in registers, row = lane + 32 r
// A "phase" never reads shared memory that another lane writes in the same phase.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LMPC_HD __host__ __device__ __forceinline__
#else
#define LMPC_HD inline
#endif

namespace lmpc {

// ------------------------------------------------------------------------------------------
// lane abstraction
// ------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define LMPC_NLANE 32
#define LMPC_LANE ((int)(threadIdx.x & 31))
__device__ __forceinline__ void wsync() { __syncwarp(); }
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double wmin(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
#else
#define LMPC_NLANE 1
#define LMPC_LANE 0
inline void wsync() {}
inline double wsum(double v) { return v; }
inline double wmin(double v) { return v; }
inline double wmax(double v) { return v; }
#endif

#if defined(__CUDA_ARCH__)
#define FOR_LANES(e, n) _Pragma("unroll") for (int e = LMPC_LANE; e < (n); e += LMPC_NLANE)
#else
#define FOR_LANES(e, n) for (int e = LMPC_LANE; e < (n); e += LMPC_NLANE)
#endif
#define NSLOT(CNT) (((CNT) + LMPC_NLANE - 1) / LMPC_NLANE)
// row = lane + 32*r ; the body runs only for valid rows.  No warp collectives inside.
#define FOR_SLOTS(r, row, CNT) \
    _Pragma("unroll") for (int r = 0, row = LMPC_LANE; r < NSLOT(CNT); ++r, row += LMPC_NLANE) if (row < (CNT))

// ------------------------------------------------------------------------------------------
// problem constants (one copy per controller configuration; __constant__ on the device)
// ------------------------------------------------------------------------------------------
constexpr int NX = 6;       // state dimension  (reference: n = 6, main.py:44)
constexpr int NU = 2;       // input dimension  (reference: d = 2)
constexpr int MAX_NCX = 4;  // rows of Fx supported
constexpr int MAX_NCU = 8;  // rows of Fu supported

struct FtocpConst {
    double Q2[36], Qf2[36], R2[4];  // 2Q, 2Qf, 2R           (buildCost: H = 2*blkdiag, PC.py:257)
    double qx[6], qxN[6];           // -2 Q xRef, -2 Qf xRef  (PC.py:245)
    double dR2[2];                  // 2 dR                   (PC.py:233-242)
    double qs2, ql;                 // 2*Qslack[0], Qslack[1] (PC.py:249-250)
    double Fx[MAX_NCX * 6], bx[MAX_NCX], Fu[MAX_NCU * 2], bu[MAX_NCU];  // PC.py:166-198
    double T[36], Tinv[36];         // 2*QterminalSlack and its inverse (PC.py:361)
    double eps_res, eps_gap, d4_min;
    int max_iter;
    int pad_;
};

enum Status : int {
    ST_SOLVED = 1,
    ST_MAX_ITER = 2,
    ST_NUMERICAL = 3,   // non-positive pivot / non-finite step
    ST_BAD_INPUT = 4
};

// ------------------------------------------------------------------------------------------
// per-instance workspace (shared memory on the device)
// ------------------------------------------------------------------------------------------
template <int N, int M, int NCX, int NCU>
struct Work {
    static constexpr int MM = (M > 0 ? M : 1);
    static constexpr int TM = (M > 0 ? 36 : 1);
    // --- model (filled by the loader; ABC via cp.async.bulk) ---
    // As loaded: per stage A (36, row major a*6+b) | B (12, a*2+q) | C (6).  prepare_model() transposes A and B in
    // place, after which ABC[k][j*6 + c] = [A_k B_k](c, j), j < 8: every dot product of the sweeps then runs over a
    // contiguous, 16 B-aligned 6-vector (LDS.128).
    alignas(16) double ABC[N][54];
    alignas(16) double SS[6 * MM];  // SS[a*M + l]      (PC.py:411 SS_PointSelectedTot, 6 x M)
    double Qfun[MM];                // Qfun_SelectedTot (PC.py:412)
    double uOld[2];                 // OldInput         (PC.py:136,247)
    // --- iterate ---
    alignas(16) double x[(N + 1) * 6];
    alignas(16) double dx[(N + 1) * 6];
    double u[N * 2], du[N * 2];
    // (dx also stages nu1 | nu2 for the stage-gradient pre-pass: it is dead between the update and the next forward sweep)
    static_assert(N * (NCX + NCU) <= (N + 1) * 6, "staging area too small");
    // --- per-row quantities shared between lanes ---
    double Dt[N * NCX], ex[N * NCX];  // condensed lane-constraint Hessian weights / rhs
    double d2[N * NCU], eu[N * NCU];  // input-bound Hessian weights / rhs
    double d4i[MM];                   // 1 / max(nu4/lam, d4_min)
    // --- Riccati factor, per stage ---
    alignas(16) double Zt[N][16];   // Z~ = L^-1 [ (B'Pxx+Pxv')A | -diag(dR2) ]  (2 x 8, row major)
    double Li[N][3];                // 1/L00, L10, 1/L11   (L = chol of the 2x2 input Hessian)
    double z0[N][2];                // L^-1 g0 for the current right-hand side
    double ru[N][2];                // input-stationarity residual
    alignas(16) double gst[N][8];   // stage gradient: (2Q x + qx + Fx'nu1 | 2R u + rate + Fu'nu2)
    // --- sweep scratch (augmented state (x, v = previous input), 8 x 8) ---
    static constexpr int PS = 10;   // padded row stride of Paug / Gt (80 B: the 8 rows hit distinct 16 B bank groups)
    alignas(16) double Paug[8 * PS];  // cost-to-go Hessian of stage k+1
    alignas(16) double Gt[8 * PS];    // Gt[j][a] = (Paug A~)(a, j),  A~ = [A B; 0 I]
    alignas(16) double S[64];       // A~' Paug A~ + stage Hessian
    alignas(16) double pb[2][8];    // cost-to-go gradient (px | pv), double buffered
    alignas(16) double pi[2][8];    // costate (6 used), double buffered
    double hv[8];                   // h~ = r~ + A~' p
    // --- terminal block ---
    double Wm[TM], Wi[TM];
    double sbar[6], yT[6];
    int flag;
};

// per-lane register state -------------------------------------------------------------------
template <int N, int M, int NCX, int NCU>
struct Regs {
    static constexpr int R1 = N * NCX, R2 = N * NCU, R4 = (M > 0 ? M : 1);
    // lane-constraint pair (k,i):  Fx_i x_k - s <= bx_i  and  s >= 0
    double s[NSLOT(R1)], nu1[NSLOT(R1)], nu3[NSLOT(R1)];
    double w1[NSLOT(R1)], rs[NSLOT(R1)], d1[NSLOT(R1)], p1[NSLOT(R1)], p3[NSLOT(R1)];
    double iw1[NSLOT(R1)], is_[NSLOT(R1)], ihs[NSLOT(R1)];   // reciprocals of w1, s, hs = qs2 + nu1/w1 + nu3/s (one division each per iteration)
    // input bound (k,j):  Fu_j u_k <= bu_j
    double nu2[NSLOT(R2)], w2[NSLOT(R2)], iw2[NSLOT(R2)], p2[NSLOT(R2)];
    // simplex multiplier l:  lam_l >= 0
    double lam[NSLOT(R4)], ilam[NSLOT(R4)], nu4[NSLOT(R4)], rl[NSLOT(R4)], p4[NSLOT(R4)], rho[NSLOT(R4)], dlam[NSLOT(R4)];
    double y1;
};

constexpr double CENTRALITY_GAMMA = 0.01;
constexpr int LATE_ACCEPT_IT = 20;
constexpr int RECENTRE_AFTER = 3;   // step reductions before the corrector is replaced by a centring step

struct SolveInfo {
    int status, iters;
    int late;            // 1 = reported solved by the late-acceptance safety net (1e-6 contract), not at eps_res / eps_gap
    double r_prim, r_dual, gap;
};

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
LMPC_HD double dot6(const double* a, const double* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
// dot product of two contiguous, 16 B-aligned 6-vectors in shared memory (three LDS.128 per operand)
LMPC_HD double dot6v(const double* a, const double* b) {
#if defined(__CUDA_ARCH__)
    const double2 a0 = *reinterpret_cast<const double2*>(a), a1 = *reinterpret_cast<const double2*>(a + 2),
                  a2 = *reinterpret_cast<const double2*>(a + 4);
    const double2 b0 = *reinterpret_cast<const double2*>(b), b1 = *reinterpret_cast<const double2*>(b + 2),
                  b2 = *reinterpret_cast<const double2*>(b + 4);
    return (a0.x * b0.x + a0.y * b0.y) + (a1.x * b1.x + a1.y * b1.y) + (a2.x * b2.x + a2.y * b2.y);
#else
    return (a[0] * b[0] + a[1] * b[1]) + (a[2] * b[2] + a[3] * b[3]) + (a[4] * b[4] + a[5] * b[5]);
#endif
}
LMPC_HD double rsqrt_f64(double v) {
#if defined(__CUDA_ARCH__)
    return rsqrt(v);
#else
    return 1.0 / sqrt(v);
#endif
}
// (i, j), i <= j, of the e-th entry of the upper triangle of an 8 x 8 matrix (row-wise)
LMPC_HD void tri8(int e, int& i, int& j) {
    i = 0;
    while (e >= 8 - i) { e -= 8 - i; ++i; }
    j = e + i;
}
// Ratio test without divisions: (rn, rd) is the largest -dv/v seen so far as a fraction (start 0/1; v, rd > 0).
// The step bound of the lane is rd/rn, one division per test instead of one per constraint.
LMPC_HD void ratio_update(double v, double dv, double& rn, double& rd) {
    const double nn = -dv;
    if (nn * rd > rn * v) { rn = nn; rd = v; }
}
LMPC_HD double ratio_bound(double rn, double rd, double a) { return (rn > 0.0) ? fmin(a, rd / rn) : a; }

template <int N, int M, int NCX, int NCU>
struct Pdip {
    using W = Work<N, M, NCX, NCU>;
    using RG = Regs<N, M, NCX, NCU>;
    static constexpr int R1 = N * NCX, R2 = N * NCU, R4 = (M > 0 ? M : 1);
    static constexpr bool LMPC = (M > 0);
    static constexpr int TMW = (M > 0 ? 36 : 1);
    static constexpr int PS = W::PS;

    // ---------------------------------------------------------------- initial point ------
    static LMPC_HD void init_point(W& w, RG& g, const FtocpConst& c, const double* x0) {
        // inputs: a strictly feasible multiple of the previous input, held over the horizon
        double tau = 1.0;
#pragma unroll
        for (int j = 0; j < NCU; ++j) {
            double v = c.Fu[j * 2] * w.uOld[0] + c.Fu[j * 2 + 1] * w.uOld[1];
            if (v > 0.9 * c.bu[j]) tau = fmin(tau, 0.9 * c.bu[j] / v);
        }
        FOR_LANES(e, N * 2) w.u[e] = tau * w.uOld[e & 1];
        FOR_LANES(e, 6) w.x[e] = x0[e];
        FOR_LANES(k, N) {               // transpose A and B in place: ABC[k][j*6+c] = [A B](c, j)
            double* A = &w.ABC[k][0];
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = a + 1; b < 6; ++b) { double t = A[a * 6 + b]; A[a * 6 + b] = A[b * 6 + a]; A[b * 6 + a] = t; }
            double bt[12];
#pragma unroll
            for (int e = 0; e < 12; ++e) bt[e] = A[36 + e];
#pragma unroll
            for (int cc = 0; cc < 6; ++cc) { A[36 + cc] = bt[cc * 2]; A[42 + cc] = bt[cc * 2 + 1]; }
        }
        wsync();
        for (int k = 0; k < N; ++k) {   // roll the model out (dynamics hold from the start)
            FOR_LANES(a, 6) {
                const double* T = &w.ABC[k][0];
                double v = T[48 + a] + T[36 + a] * w.u[k * 2] + T[42 + a] * w.u[k * 2 + 1];
#pragma unroll
                for (int b = 0; b < 6; ++b) v += T[b * 6 + a] * w.x[k * 6 + b];
                w.x[(k + 1) * 6 + a] = v;
            }
            wsync();
        }
        // Dual-feasible, centred start (oracle/pdip_model.py, mu0 = "auto"): the slack-stationarity row
        // nu1 + nu3 = 2 qs s + ql holds exactly with w1 nu1 = s nu3 = mu_row; every other constraint
        // family starts at the mean of those products.
        double mu_acc = 0.0;
        FOR_SLOTS(r, row, R1) {
            int k = row / NCX, i = row % NCX;
            double fx = dot6(&c.Fx[i * 6], &w.x[k * 6]) - c.bx[i];
            double s = fmax(fx, 0.0) + 0.3;
            double w1 = s - fx;
            double mur = fmax((c.qs2 * s + c.ql) * (w1 * s) / (w1 + s), 1e-3);
            g.s[r] = s;
            g.nu1[r] = mur / w1;
            g.nu3[r] = mur / s;
            mu_acc += mur;
        }
        const double mu0 = fmax(wsum(mu_acc) / (double)R1, 1e-3);
        FOR_SLOTS(r, row, R2) {
            int k = row / NCU, j = row % NCU;
            double w2 = c.bu[j] - (c.Fu[j * 2] * w.u[k * 2] + c.Fu[j * 2 + 1] * w.u[k * 2 + 1]);
            g.nu2[r] = mu0 / w2;
        }
        if (LMPC) {
            FOR_SLOTS(r, row, R4) { g.lam[r] = 1.0 / M; }
            terminal_state(w, g, c);           // xi, yT for lam = 1/M
            double red_min = 1e300;
            FOR_SLOTS(r, row, R4) {
                double red = w.Qfun[row];
#pragma unroll
                for (int a = 0; a < 6; ++a) red -= w.SS[a * M + row] * w.yT[a];
                g.nu4[r] = red;
                red_min = fmin(red_min, red);
            }
            red_min = wmin(red_min);
            g.y1 = -red_min + mu0 * (double)M;
            FOR_SLOTS(r, row, R4) { g.nu4[r] += g.y1; }
        }
    }

    // xi = SS lam - x_N, yT = -T xi (both derived every iteration: the terminal equality and the
    // xi-stationarity row then hold by construction).  Returns sum(lam) - 1.
    static LMPC_HD double terminal_state(W& w, RG& g, const FtocpConst& c) {
        double acc[6] = {0, 0, 0, 0, 0, 0}, sl = 0.0;
        FOR_SLOTS(r, row, R4) {
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[a] += w.SS[a * M + row] * g.lam[r];
            sl += g.lam[r];
        }
        double xi[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) xi[a] = wsum(acc[a]) - w.x[N * 6 + a];
        sl = wsum(sl);
        wsync();   // previous readers of yT are done
        FOR_LANES(a, 6) {
            double y = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) y -= c.T[a * 6 + b] * xi[b];
            w.yT[a] = y;
        }
        wsync();
        return sl - 1.0;
    }

    // ---------------------------------------------------------------- terminal factor ----
    // d4i, centroid, W = sum d4i s~ s~' + Tinv, Wi = W^-1 via Cholesky (PSD by construction).
    static LMPC_HD double terminal_factor(W& w, RG& g, const FtocpConst& c, double d4_floor) {
        double acc[6] = {0, 0, 0, 0, 0, 0}, dl = 0.0;
        FOR_SLOTS(r, row, R4) {
            double d4 = fmax(g.nu4[r] * g.ilam[r], d4_floor);
            double di = 1.0 / d4;
            w.d4i[row] = di;
            dl += di;
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[a] += w.SS[a * M + row] * di;
        }
        double delta = wsum(dl);
        double sb[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sb[a] = wsum(acc[a]) / delta;
#pragma unroll
        for (int a = 0; a < 6; ++a)
            if (LMPC_LANE == (a % LMPC_NLANE)) w.sbar[a] = sb[a];   // same value on every lane
        wsync();
        // W (21 unique entries), four partial accumulators each
        FOR_LANES(e, 21) {
            int a = 0, b = e;
            while (b >= 6 - a) { b -= 6 - a; ++a; }
            b += a;   // (a,b), a <= b
            double sa = w.sbar[a], sbb = w.sbar[b];
            double t0 = 0, t1 = 0, t2 = 0, t3 = 0;
            const double* Sa = &w.SS[a * M];
            const double* Sb = &w.SS[b * M];
            int l = 0;
            for (; l + 3 < M; l += 4) {
                t0 += (Sa[l] - sa) * (Sb[l] - sbb) * w.d4i[l];
                t1 += (Sa[l + 1] - sa) * (Sb[l + 1] - sbb) * w.d4i[l + 1];
                t2 += (Sa[l + 2] - sa) * (Sb[l + 2] - sbb) * w.d4i[l + 2];
                t3 += (Sa[l + 3] - sa) * (Sb[l + 3] - sbb) * w.d4i[l + 3];
            }
            for (; l < M; ++l) t0 += (Sa[l] - sa) * (Sb[l] - sbb) * w.d4i[l];
            double v = (t0 + t1) + (t2 + t3) + c.Tinv[a * 6 + b];
            w.Wm[a * 6 + b] = v;
            w.Wm[b * 6 + a] = v;
        }
        wsync();
        // lane-redundant 6x6 Cholesky, triangular inverse and Wi = Linv' Linv
        double L[21];   // packed lower, row i: i(i+1)/2 + j
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) {
                double v = w.Wm[i * 6 + j];
#pragma unroll
                for (int k = 0; k < j; ++k) v -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
                if (i == j) {
                    if (!(v > 0.0)) { ok = false; v = 1.0; }
                    L[i * (i + 1) / 2 + i] = rsqrt_f64(v);           // store the reciprocal pivot
                } else {
                    L[i * (i + 1) / 2 + j] = v * L[j * (j + 1) / 2 + j];
                }
            }
        }
        double X[21];   // Linv, packed lower
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            X[j * (j + 1) / 2 + j] = L[j * (j + 1) / 2 + j];
#pragma unroll
            for (int i = j + 1; i < 6; ++i) {
                double v = 0.0;
#pragma unroll
                for (int k = j; k < i; ++k) v -= L[i * (i + 1) / 2 + k] * X[k * (k + 1) / 2 + j];
                X[i * (i + 1) / 2 + j] = v * L[i * (i + 1) / 2 + i];
            }
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) {
#pragma unroll
            for (int b = 0; b <= a; ++b) {
                double v = 0.0;
#pragma unroll
                for (int k = a; k < 6; ++k) v += X[k * (k + 1) / 2 + a] * X[k * (k + 1) / 2 + b];
                if (LMPC_LANE == ((a * (a + 1) / 2 + b) % LMPC_NLANE)) {
                    w.Wi[a * 6 + b] = v;
                    w.Wi[b * 6 + a] = v;
                }
            }
        }
        if (!ok) w.flag = ST_NUMERICAL;
        wsync();
        return delta;
    }

    // right-hand side of the terminal block for one solve: c1 (all lanes) and beta.
    // b1 = rhs of the simplex row (= -(sum lam - 1)).
    static LMPC_HD void terminal_rhs(W& w, RG& g, double b1, double* c1, double& beta) {
        double acc[6] = {0, 0, 0, 0, 0, 0}, sb = 0.0;
        FOR_SLOTS(r, row, R4) {
            double t = w.d4i[row] * g.rho[r];
            sb += t;
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[a] += (w.SS[a * M + row] - w.sbar[a]) * t;
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) c1[a] = -wsum(acc[a]) - w.sbar[a] * b1;
        beta = b1 - wsum(sb);
    }

    // ---------------------------------------------------------------- backward sweeps ----
    // Start of a backward sweep: terminal cost-to-go into (Paug | pb[N&1]) and pi_N.
    template <bool FACTOR>
    static LMPC_HD void backward_start(W& w, const FtocpConst& c, const double* c1) {
        if (FACTOR) {
            FOR_LANES(e, 64) {
                const int i = e >> 3, j = e & 7;
                w.Paug[i * PS + j] = (i < 6 && j < 6) ? c.Qf2[i * 6 + j] + (LMPC ? w.Wi[(i * 6 + j) % TMW] : 0.0) : 0.0;
            }
            FOR_LANES(a, 8) {   // pi_N = -(Qf2 x_N + qxN + yT)
                double v = 0.0;
                if (a < 6) {
                    v = c.qxN[a] + (LMPC ? w.yT[a] : 0.0);
#pragma unroll
                    for (int b = 0; b < 6; ++b) v += c.Qf2[a * 6 + b] * w.x[N * 6 + b];
                }
                w.pi[N & 1][a] = -v;
            }
        }
        FOR_LANES(a, 8) {
            double v = 0.0;
            if (LMPC && a < 6) {
#pragma unroll
                for (int b = 0; b < 6; ++b) v += w.Wi[(a * 6 + b) % TMW] * c1[b];
            }
            w.pb[N & 1][a] = v;
        }
        wsync();
    }

    // Stage gradients for the costate / input-residual recursion (embarrassingly parallel over stages).
    static LMPC_HD void stage_gradients(W& w, const FtocpConst& c) {
        FOR_LANES(e, N * 8) {
            const int k = e >> 3, j = e & 7;
            double v;
            if (j < 6) {
                v = c.qx[j];
#pragma unroll
                for (int b = 0; b < 6; ++b) v += c.Q2[j * 6 + b] * w.x[k * 6 + b];
#pragma unroll
                for (int i = 0; i < NCX; ++i) v += c.Fx[i * 6 + j] * w.dx[k * NCX + i];
            } else {
                const int r = j - 6;
                const double uk = w.u[k * 2 + r];
                const double up = (k == 0) ? w.uOld[r] : w.u[(k - 1) * 2 + r];
                v = c.R2[r * 2] * w.u[k * 2] + c.R2[r * 2 + 1] * w.u[k * 2 + 1] + c.dR2[r] * (uk - up);
                if (k < N - 1) v += c.dR2[r] * (uk - w.u[(k + 1) * 2 + r]);
#pragma unroll
                for (int jj = 0; jj < NCU; ++jj) v += c.Fu[jj * 2 + r] * w.dx[N * NCX + k * NCU + jj];
            }
            w.gst[k][j] = v;
        }
    }

    // r~_k: right-hand side contribution of the eliminated inequality rows (state part j < 6, input part r < 2)
    static LMPC_HD double rhs_x(const W& w, const FtocpConst& c, int k, int j) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NCX; ++i) v += c.Fx[i * 6 + j] * w.ex[k * NCX + i];
        return v;
    }
    static LMPC_HD double rhs_u(const W& w, const FtocpConst& c, int k, int r) {
        double v = 0.0;
#pragma unroll
        for (int jj = 0; jj < NCU; ++jj) v += c.Fu[jj * 2 + r] * w.eu[k * NCU + jj];
        return v;
    }

    // work item e < 36 of phase b/e: entry (i, j), i <= j, of the 8 x 8 stage matrix and its constant data
    static constexpr int KF = (NCX > NCU ? NCX : NCU);
    struct SEnt {
        int i, j, cls;       // class 0 = state-state, 1 = input-state, 2 = input-input
        double kq, kf[KF];
    };
    static LMPC_HD void s_entry(const FtocpConst& c, int e, SEnt& q) {
        q.i = q.j = 0; q.cls = 1; q.kq = 0.0;
#pragma unroll
        for (int z = 0; z < KF; ++z) q.kf[z] = 0.0;
        if (e >= 36) return;
        tri8(e, q.i, q.j);
        if (q.j < 6) {
            q.cls = 0;
            q.kq = c.Q2[q.i * 6 + q.j];
#pragma unroll
            for (int z = 0; z < NCX; ++z) q.kf[z] = c.Fx[z * 6 + q.i] * c.Fx[z * 6 + q.j];
        } else if (q.i >= 6) {
            q.cls = 2;
            q.kq = c.R2[(q.i - 6) * 2 + (q.j - 6)];
#pragma unroll
            for (int z = 0; z < NCU; ++z) q.kf[z] = c.Fu[z * 2 + (q.i - 6)] * c.Fu[z * 2 + (q.j - 6)];
        }
    }

    // Factorising backward sweep on the augmented state (x, v): Riccati matrices, gradient recursion for the
    // predictor right-hand side and the costate / input-residual recursion.  Four uniform phases per stage.
    static LMPC_HD double backward_factor(W& w, const FtocpConst& c) {
        double ru_max = 0.0;
        // loop-invariant per-lane work assignment, cached in registers on the device (2 items per lane)
        constexpr bool CACHED = (LMPC_NLANE == 32);
        SEnt se[CACHED ? 2 : 1];
        if (CACHED) {
            s_entry(c, LMPC_LANE, se[0]);
            s_entry(c, LMPC_LANE + 32, se[CACHED ? 1 : 0]);
        }
        for (int k = N - 1; k >= 0; --k) {
            const double* T = &w.ABC[k][0];           // T[j*6 + c] = [A B](c, j)
            const double* pn = w.pb[(k + 1) & 1];     // px | pv of stage k+1
            const double* pin = w.pi[(k + 1) & 1];    // pi_{k+1}
            // ---- phase a: Gt[j][a] = (Paug A~)(a, j)
            FOR_LANES(e, 64) {
                const int j = e >> 3, a = e & 7;
                double v = dot6v(&T[j * 6], &w.Paug[a * PS]);
                if (j >= 6) v += w.Paug[a * PS + j];
                w.Gt[j * PS + a] = v;
            }
            wsync();
            // ---- phase b: S = A~' G~ + stage Hessian (36 entries), then the vector recursions (8 items)
            {
                int t = 0;
                FOR_LANES(e, 44) {
                    if (e < 36) {
                        SEnt qq;
                        if (!CACHED) s_entry(c, e, qq);
                        const SEnt& q = CACHED ? se[t] : qq;
                        const int i = q.i, j = q.j;
                        double v = dot6v(&T[i * 6], &w.Gt[j * PS]);
                        if (i >= 6) v += w.Gt[j * PS + i];
                        if (q.cls == 0) {
                            v += q.kq;
#pragma unroll
                            for (int z = 0; z < NCX; ++z) v += w.Dt[k * NCX + z] * q.kf[z];
                        } else if (q.cls == 2) {
                            v += q.kq;
                            if (i == j) v += ((k < N - 1) ? 2.0 : 1.0) * c.dR2[i - 6];
#pragma unroll
                            for (int z = 0; z < NCU; ++z) v += w.d2[k * NCU + z] * q.kf[z];
                        }
                        w.S[i * 8 + j] = v;
                        w.S[j * 8 + i] = v;
                    } else {
                        const int j = e - 36;                       // 0..7
                        const double hp = dot6v(&T[j * 6], pn);     // A~' p
                        const double tp = dot6v(&T[j * 6], pin);    // A~' pi
                        if (j < 6) {
                            w.hv[j] = rhs_x(w, c, k, j) + hp;
                            w.pi[k & 1][j] = tp - w.gst[k][j];
                        } else {
                            const double ru = w.gst[k][j] - tp;
                            w.ru[k][j - 6] = ru;
                            w.hv[j] = ru + rhs_u(w, c, k, j - 6) + hp + pn[j];
                        }
                    }
                    ++t;
                }
            }
            wsync();
            // ---- phase c/d: 2x2 Cholesky (lane redundant), Z~, z0
            {
                double l00s = w.S[6 * 8 + 6], l10 = w.S[7 * 8 + 6], l11s = w.S[7 * 8 + 7];
                bool bad = !(l00s > 0.0);
                if (bad) l00s = 1.0;
                const double i00 = rsqrt_f64(l00s);
                l10 *= i00;
                double tt = l11s - l10 * l10;
                if (!(tt > 0.0)) { bad = true; tt = 1.0; }
                const double i11 = rsqrt_f64(tt);
                ru_max = fmax(ru_max, fmax(fabs(w.ru[k][0]), fabs(w.ru[k][1])));
                FOR_LANES(j, 9) {
                    if (j < 8) {
                        const double y0 = (j < 6) ? w.S[6 * 8 + j] : (j == 6 ? -c.dR2[0] : 0.0);
                        const double y1 = (j < 6) ? w.S[7 * 8 + j] : (j == 7 ? -c.dR2[1] : 0.0);
                        const double z0c = y0 * i00;
                        w.Zt[k][j] = z0c;
                        w.Zt[k][8 + j] = (y1 - l10 * z0c) * i11;
                    } else {
                        const double a0 = w.hv[6] * i00;
                        w.z0[k][0] = a0;
                        w.z0[k][1] = (w.hv[7] - l10 * a0) * i11;
                        w.Li[k][0] = i00;
                        w.Li[k][1] = l10;
                        w.Li[k][2] = i11;
                        if (bad) w.flag = ST_NUMERICAL;
                    }
                }
            }
            wsync();
            // ---- phase e: cost-to-go of stage k: Paug = [Sxx 0; 0 0] - Z~'Z~ ; p = [hx; 0] - Z~' z0
            {
                const double* Z = w.Zt[k];
                const double a0 = w.z0[k][0], a1 = w.z0[k][1];
                double* po = w.pb[k & 1];
                int t = 0;
                FOR_LANES(e, 44) {
                    if (e < 36) {
                        SEnt qq;
                        if (!CACHED) s_entry(c, e, qq);
                        const SEnt& q = CACHED ? se[t] : qq;
                        const int i = q.i, j = q.j;
                        const double v = ((q.cls == 0) ? w.S[i * 8 + j] : 0.0) - Z[i] * Z[j] - Z[8 + i] * Z[8 + j];
                        w.Paug[i * PS + j] = v;
                        w.Paug[j * PS + i] = v;
                    } else {
                        const int a = e - 36;
                        po[a] = ((a < 6) ? w.hv[a] : 0.0) - Z[a] * a0 - Z[8 + a] * a1;
                    }
                    ++t;
                }
            }
            wsync();
        }
        return ru_max;
    }

    // Gradient-only backward sweep for a new right-hand side (corrector), one phase per stage.
    static LMPC_HD void backward_rhs(W& w, const FtocpConst& c) {
        for (int k = N - 1; k >= 0; --k) {
            const double* T = &w.ABC[k][0];
            const double* pn = w.pb[(k + 1) & 1];
            double* po = w.pb[k & 1];
            double g0[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) g0[r] = w.ru[k][r] + rhs_u(w, c, k, r) + pn[6 + r] + dot6v(&T[(6 + r) * 6], pn);
            const double a0 = g0[0] * w.Li[k][0];
            const double a1 = (g0[1] - w.Li[k][1] * a0) * w.Li[k][2];
            FOR_LANES(a, 8) {
                double v = 0.0;
                if (a < 6) v = rhs_x(w, c, k, a) + dot6v(&T[a * 6], pn);
                po[a] = v - w.Zt[k][a] * a0 - w.Zt[k][8 + a] * a1;
                if (a >= 6) w.z0[k][a - 6] = (a == 6) ? a0 : a1;
            }
            wsync();
        }
    }

    // ---------------------------------------------------------------- forward sweep ------
    static LMPC_HD void forward(W& w) {
        FOR_LANES(a, 6) w.dx[a] = 0.0;
        wsync();
        double dv0 = 0.0, dv1 = 0.0;
        for (int k = 0; k < N; ++k) {
            const double* T = &w.ABC[k][0];
            const double* Z = w.Zt[k];
            const double* d = &w.dx[k * 6];
            const double t0 = w.z0[k][0] + dot6v(Z, d) + Z[6] * dv0 + Z[7] * dv1;
            const double t1 = w.z0[k][1] + dot6v(Z + 8, d) + Z[14] * dv0 + Z[15] * dv1;
            const double du1 = -t1 * w.Li[k][2];
            const double du0 = (-t0 - w.Li[k][1] * du1) * w.Li[k][0];
            FOR_LANES(a, 8) {
                if (a < 6) {
                    double v = T[36 + a] * du0 + T[42 + a] * du1;
#pragma unroll
                    for (int b = 0; b < 6; ++b) v += T[b * 6 + a] * d[b];
                    w.dx[(k + 1) * 6 + a] = v;
                } else {
                    w.du[k * 2 + (a - 6)] = (a == 6) ? du0 : du1;
                }
            }
            dv0 = du0;
            dv1 = du1;
            wsync();
        }
    }

    // terminal recovery after a forward sweep: dlam (registers); returns dy1.
    // The 6x6 covariance-form solve loses ~cond(W)*eps in dyT, which the division by a small d4 amplifies in
    // the dlam of active safe-set points.  One step of iterative refinement with the residual of
    //     Tinv dyT + S~ dlam = dx_N - sbar*b1
    // evaluated THROUGH the recovered dlam removes it (oracle/pdip_model.py W_REFINE).
    static LMPC_HD double terminal_recover(W& w, RG& g, const FtocpConst& c, const double* c1, double beta,
                                           double delta, double b1) {
        double v[6], dyT[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) v[a] = w.dx[N * 6 + a] + c1[a];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double t = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) t += w.Wi[a * 6 + b] * v[b];
            dyT[a] = t;
        }
        const double dy1t = -beta / delta;
        double acc[6] = {0, 0, 0, 0, 0, 0};
        FOR_SLOTS(r, row, R4) {
            double t = g.rho[r] - dy1t;
#pragma unroll
            for (int a = 0; a < 6; ++a) t += (w.SS[a * M + row] - w.sbar[a]) * dyT[a];
            t *= w.d4i[row];
            g.dlam[r] = t;
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[a] += (w.SS[a * M + row] - w.sbar[a]) * t;
        }
        double e[6], ddy[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double t = w.dx[N * 6 + a] - w.sbar[a] * b1 - wsum(acc[a]);
#pragma unroll
            for (int b = 0; b < 6; ++b) t -= c.Tinv[a * 6 + b] * dyT[b];
            e[a] = t;
        }
        double sdy = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double t = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) t += w.Wi[a * 6 + b] * e[b];
            ddy[a] = t;
            sdy += w.sbar[a] * (dyT[a] + t);
        }
        FOR_SLOTS(r, row, R4) {
            double t = 0.0;
#pragma unroll
            for (int a = 0; a < 6; ++a) t += (w.SS[a * M + row] - w.sbar[a]) * ddy[a];
            g.dlam[r] += t * w.d4i[row];
        }
        return dy1t + sdy;
    }

    // ---------------------------------------------------------------- the solver ---------
    // Preconditions: w.ABC, w.SS, w.Qfun, w.uOld loaded and visible to the warp; x0[6].
    static LMPC_HD void solve(W& w, const FtocpConst& c, const double* x0, SolveInfo& info,
                              double* lam_out /* M or null */, double* slack_out /* R1 or null */) {
        RG g;
        if (LMPC_LANE == 0) w.flag = 0;
        wsync();
        init_point(w, g, c, x0);
        const double n_ineq = (double)(2 * R1 + R2 + (LMPC ? M : 0));
        int it = 0, status = ST_MAX_ITER, late = 0;
        double r_prim = 0.0, r_dual = 0.0, mu = 0.0, ru_prev = 0.0, al_prev = 0.0;
        const double d4_floor = c.d4_min;

        for (;; ++it) {
            // ---- lane-local residuals, barrier diagonals, predictor right-hand sides -------
            double comp = 0.0, rd_loc = 0.0;
            FOR_SLOTS(r, row, R1) {
                int k = row / NCX, i = row % NCX;
                double w1 = c.bx[i] - dot6(&c.Fx[i * 6], &w.x[k * 6]) + g.s[r];   // derived slack
                g.w1[r] = w1;
                double rs = c.qs2 * g.s[r] + c.ql - g.nu1[r] - g.nu3[r];
                g.rs[r] = rs;
                rd_loc = fmax(rd_loc, fabs(rs));
                comp += w1 * g.nu1[r] + g.s[r] * g.nu3[r];
                const double iw1 = 1.0 / w1, is = 1.0 / g.s[r];
                g.iw1[r] = iw1;
                g.is_[r] = is;
                double d1 = g.nu1[r] * iw1, d3 = g.nu3[r] * is;
                const double ihs = 1.0 / (c.qs2 + d1 + d3);
                g.d1[r] = d1;
                g.ihs[r] = ihs;
                w.Dt[row] = d1 * (c.qs2 + d3) * ihs;
                // predictor: rc1 = w1 nu1, rc3 = s nu3  ->  e1 = -nu1, rs + rc3/s = rs + nu3
                w.ex[row] = (-g.nu1[r] * (c.qs2 + d3) + d1 * (rs + g.nu3[r])) * ihs;
                w.dx[row] = g.nu1[r];
            }
            FOR_SLOTS(r, row, R2) {
                int k = row / NCU, j = row % NCU;
                double w2 = c.bu[j] - (c.Fu[j * 2] * w.u[k * 2] + c.Fu[j * 2 + 1] * w.u[k * 2 + 1]);
                g.w2[r] = w2;
                comp += w2 * g.nu2[r];
                const double iw2 = 1.0 / w2;
                g.iw2[r] = iw2;
                w.d2[row] = g.nu2[r] * iw2;
                w.eu[row] = -g.nu2[r];           // predictor: -rc2/w2
                w.dx[N * NCX + row] = g.nu2[r];
            }
            double rone = 0.0, delta = 1.0, beta = 0.0, c1[6] = {0, 0, 0, 0, 0, 0};
            if (LMPC) {
                rone = terminal_state(w, g, c);
                FOR_SLOTS(r, row, R4) {
                    double rl = w.Qfun[row] + g.y1 - g.nu4[r];
#pragma unroll
                    for (int a = 0; a < 6; ++a) rl -= w.SS[a * M + row] * w.yT[a];
                    g.rl[r] = rl;
                    rd_loc = fmax(rd_loc, fabs(rl));
                    comp += g.lam[r] * g.nu4[r];
                    g.ilam[r] = 1.0 / g.lam[r];
                    g.rho[r] = -rl - g.nu4[r];   // predictor: rc4/lam = nu4
                }
            } else {
                wsync();
            }
            comp = wsum(comp);
            mu = comp / n_ineq;
            rd_loc = wmax(rd_loc);
            r_prim = fabs(rone);
            if (it > 0) {
                // The input-stationarity residual is linear in the iterate and every variable moved by the same step
                // length, so after a step alpha it is exactly (1 - alpha) times its previous value: convergence can be
                // decided here, before paying for another factorisation.
                r_dual = fmax(rd_loc, (1.0 - al_prev) * ru_prev);
                if (r_prim <= c.eps_res && r_dual <= c.eps_res && mu <= c.eps_gap) { status = ST_SOLVED; break; }
                // Stragglers: on a few LMPC instances (LP-degenerate simplex block) the covariance-form recovery of
                // d(lambda) puts a noise floor of ~1e-7..1e-6 under the dual residual and the tail converges linearly.
                // Once the iterate meets the 1e-6 parity contract, stop after LATE_ACCEPT_IT iterations (and at max_iter).
                if ((it >= LATE_ACCEPT_IT || it >= c.max_iter) && r_prim <= 1e-6 && r_dual <= 1e-6 && mu <= 1e-6) {
                    status = ST_SOLVED;
                    late = 1;
                    break;
                }
                if (it >= c.max_iter) { status = ST_MAX_ITER; break; }
            }
            // ---- factorising backward sweep (also yields the input residual) ------------------
            if (LMPC) {
                delta = terminal_factor(w, g, c, d4_floor);
                terminal_rhs(w, g, -rone, c1, beta);
            }
            stage_gradients(w, c);
            backward_start<true>(w, c, c1);
            const double ru_max = backward_factor(w, c);
            ru_prev = ru_max;
            r_dual = fmax(rd_loc, ru_max);
            if (w.flag != 0) { status = w.flag; break; }
            if (r_prim <= c.eps_res && r_dual <= c.eps_res && mu <= c.eps_gap) { status = ST_SOLVED; break; }
            if (it >= c.max_iter) {
                if (r_prim <= 1e-6 && r_dual <= 1e-6 && mu <= 1e-6) { status = ST_SOLVED; late = 1; } else { status = ST_MAX_ITER; }
                break;
            }

            // ---- predictor -----------------------------------------------------------------------
            forward(w);
            double dy1 = 0.0;
            if (LMPC) dy1 = terminal_recover(w, g, c, c1, beta, delta, -rone);
            double rn = 0.0, rd = 1.0;
            FOR_SLOTS(r, row, R1) {
                int k = row / NCX, i = row % NCX;
                double fdx = dot6(&c.Fx[i * 6], &w.dx[k * 6]);
                double gs = -g.rs[r] - g.nu1[r] - g.nu3[r];
                double ds = (gs + g.d1[r] * fdx) * g.ihs[r];
                double dw1 = -fdx + ds;
                double dn1 = -g.nu1[r] - g.d1[r] * dw1;
                double dn3 = -g.nu3[r] - (g.nu3[r] * g.is_[r]) * ds;
                ratio_update(g.w1[r], dw1, rn, rd);
                ratio_update(g.s[r], ds, rn, rd);
                ratio_update(g.nu1[r], dn1, rn, rd);
                ratio_update(g.nu3[r], dn3, rn, rd);
                g.p1[r] = dw1 * dn1;
                g.p3[r] = ds * dn3;
            }
            FOR_SLOTS(r, row, R2) {
                int k = row / NCU, j = row % NCU;
                double dw2 = -(c.Fu[j * 2] * w.du[k * 2] + c.Fu[j * 2 + 1] * w.du[k * 2 + 1]);
                double dn2 = -g.nu2[r] - (g.nu2[r] * g.iw2[r]) * dw2;
                ratio_update(g.w2[r], dw2, rn, rd);
                ratio_update(g.nu2[r], dn2, rn, rd);
                g.p2[r] = dw2 * dn2;
            }
            if (LMPC) {
                FOR_SLOTS(r, row, R4) {
                    double dn4 = -g.nu4[r] - (g.nu4[r] * g.ilam[r]) * g.dlam[r];
                    ratio_update(g.lam[r], g.dlam[r], rn, rd);
                    ratio_update(g.nu4[r], dn4, rn, rd);
                    g.p4[r] = g.dlam[r] * dn4;
                }
            }
            const double a_aff = wmin(ratio_bound(rn, rd, 1.0));
            // complementarity after the affine step.  With p = dw*dnu and dw*nu + w*dnu = -w*nu:
            //   (w + a dw)(nu + a dnu) = w nu (1 - a) + a^2 p
            double comp_aff = 0.0;
            FOR_SLOTS(r, row, R1) {
                comp_aff += (g.w1[r] * g.nu1[r] + g.s[r] * g.nu3[r]) * (1.0 - a_aff) + a_aff * a_aff * (g.p1[r] + g.p3[r]);
            }
            FOR_SLOTS(r, row, R2) { comp_aff += g.w2[r] * g.nu2[r] * (1.0 - a_aff) + a_aff * a_aff * g.p2[r]; }
            if (LMPC) {
                FOR_SLOTS(r, row, R4) { comp_aff += g.lam[r] * g.nu4[r] * (1.0 - a_aff) + a_aff * a_aff * g.p4[r]; }
            }
            comp_aff = wsum(comp_aff);
            double sig = comp_aff / comp;
            sig = sig * sig * sig;
            const double sm = sig * mu;

            // ---- corrector (pass 0), recentring (pass 1, rare) -----------------------------------------
            // Pass 0 is Mehrotra's corrector.  If its step cannot be brought back into the central-path neighbourhood within
            // RECENTRE_AFTER reductions, the iterate sits on the neighbourhood boundary and an aggressive sigma would only allow
            // tiny steps from now on (0.002 % of closed-loop LMPC steps stalled like that): pass 1 re-uses the factorisation
            // for a pure centring direction (sigma = 1, no second-order term) instead.
            double al = 0.0;
            double ds_[NSLOT(R1)], dw1_[NSLOT(R1)], dn1_[NSLOT(R1)], dn3_[NSLOT(R1)], dw2_[NSLOT(R2)], dn2_[NSLOT(R2)], dn4_[NSLOT(R4)];
            bool bad_step = false;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const double tgt = pass ? mu : sm;      // complementarity target
                const double so = pass ? 0.0 : 1.0;     // second-order (dw*dnu) term on/off
                FOR_SLOTS(r, row, R1) {
                    double rc1 = g.w1[r] * g.nu1[r] + so * g.p1[r] - tgt;
                    double rc3 = g.s[r] * g.nu3[r] + so * g.p3[r] - tgt;
                    double d3 = g.nu3[r] * g.is_[r];
                    double e1 = -rc1 * g.iw1[r];
                    w.ex[row] = (e1 * (c.qs2 + d3) + g.d1[r] * (g.rs[r] + rc3 * g.is_[r])) * g.ihs[r];
                    g.p1[r] = rc1;   // keep rc for the final recovery
                    g.p3[r] = rc3;
                }
                FOR_SLOTS(r, row, R2) {
                    double rc2 = g.w2[r] * g.nu2[r] + so * g.p2[r] - tgt;
                    w.eu[row] = -rc2 * g.iw2[r];
                    g.p2[r] = rc2;
                }
                if (LMPC) {
                    FOR_SLOTS(r, row, R4) {
                        double rc4 = g.lam[r] * g.nu4[r] + so * g.p4[r] - tgt;
                        g.rho[r] = -g.rl[r] - rc4 * g.ilam[r];
                        g.p4[r] = rc4;
                    }
                    terminal_rhs(w, g, -rone, c1, beta);
                }
                wsync();
                backward_start<false>(w, c, c1);
                backward_rhs(w, c);
                forward(w);
                if (LMPC) dy1 = terminal_recover(w, g, c, c1, beta, delta, -rone);

                // ---- step length ---------------------------------------------------------------------
                rn = 0.0;
                rd = 1.0;
                FOR_SLOTS(r, row, R1) {
                    int k = row / NCX, i = row % NCX;
                    double fdx = dot6(&c.Fx[i * 6], &w.dx[k * 6]);
                    double rc1 = g.p1[r], rc3 = g.p3[r];
                    double gs = -g.rs[r] - rc1 * g.iw1[r] - rc3 * g.is_[r];
                    double ds = (gs + g.d1[r] * fdx) * g.ihs[r];
                    double dw1 = -fdx + ds;
                    double dn1 = (-rc1 - g.nu1[r] * dw1) * g.iw1[r];
                    double dn3 = (-rc3 - g.nu3[r] * ds) * g.is_[r];
                    ratio_update(g.w1[r], dw1, rn, rd);
                    ratio_update(g.s[r], ds, rn, rd);
                    ratio_update(g.nu1[r], dn1, rn, rd);
                    ratio_update(g.nu3[r], dn3, rn, rd);
                    ds_[r] = ds; dw1_[r] = dw1; dn1_[r] = dn1; dn3_[r] = dn3;
                }
                FOR_SLOTS(r, row, R2) {
                    int k = row / NCU, j = row % NCU;
                    double dw2 = -(c.Fu[j * 2] * w.du[k * 2] + c.Fu[j * 2 + 1] * w.du[k * 2 + 1]);
                    double dn2 = (-g.p2[r] - g.nu2[r] * dw2) * g.iw2[r];
                    ratio_update(g.w2[r], dw2, rn, rd);
                    ratio_update(g.nu2[r], dn2, rn, rd);
                    dw2_[r] = dw2; dn2_[r] = dn2;
                }
                if (LMPC) {
                    FOR_SLOTS(r, row, R4) {
                        double dn4 = (-g.p4[r] - g.nu4[r] * g.dlam[r]) * g.ilam[r];
                        ratio_update(g.lam[r], g.dlam[r], rn, rd);
                        ratio_update(g.nu4[r], dn4, rn, rd);
                        dn4_[r] = dn4;
                    }
                }
                const double amax = wmin(ratio_bound(rn, rd, 1e300));
                al = fmin(1.0, 0.995 * amax);
                if (!(al > 0.0) || !(al <= 1.0)) { bad_step = true; break; }
                // Stay in a wide neighbourhood of the central path, min_i w_i nu_i >= gamma * mean: without it
                // Mehrotra steps can 2-cycle against a blocking bound (seen on 3 of 4096 workload QPs).
                bool inside = false;
                for (int tries = 0; tries < 12; ++tries) {
                    double pmin = 1e300, psum = 0.0;
                    FOR_SLOTS(r, row, R1) {
                        double a1 = (g.w1[r] + al * dw1_[r]) * (g.nu1[r] + al * dn1_[r]);
                        double a3 = (g.s[r] + al * ds_[r]) * (g.nu3[r] + al * dn3_[r]);
                        pmin = fmin(pmin, fmin(a1, a3));
                        psum += a1 + a3;
                    }
                    FOR_SLOTS(r, row, R2) {
                        double a2 = (g.w2[r] + al * dw2_[r]) * (g.nu2[r] + al * dn2_[r]);
                        pmin = fmin(pmin, a2);
                        psum += a2;
                    }
                    if (LMPC) {
                        FOR_SLOTS(r, row, R4) {
                            double a4 = (g.lam[r] + al * g.dlam[r]) * (g.nu4[r] + al * dn4_[r]);
                            pmin = fmin(pmin, a4);
                            psum += a4;
                        }
                    }
                    pmin = wmin(pmin);
                    psum = wsum(psum);
                    if (pmin >= CENTRALITY_GAMMA * psum / n_ineq) { inside = true; break; }
                    if (pass == 0 && tries >= RECENTRE_AFTER) break;
                    al *= 0.8;
                }
                if (inside) break;
            }
            if (bad_step) { status = ST_NUMERICAL; break; }
#if defined(LMPC_HOST_TRACE) && !defined(__CUDA_ARCH__)
            printf("it %2d rp %.2e rd %.2e mu %.2e a_aff %.4f sig %.2e al %.4e\n", it, r_prim, r_dual, mu, a_aff, sig, al);
#endif
            FOR_SLOTS(r, row, R1) {
                g.s[r] += al * ds_[r];
                g.nu1[r] += al * dn1_[r];
                g.nu3[r] += al * dn3_[r];
            }
            FOR_SLOTS(r, row, R2) { g.nu2[r] += al * dn2_[r]; }
            if (LMPC) {
                FOR_SLOTS(r, row, R4) {
                    g.lam[r] += al * g.dlam[r];
                    g.nu4[r] += al * dn4_[r];
                }
                g.y1 += al * dy1;
            }
            al_prev = al;
            FOR_LANES(e, (N + 1) * 6) w.x[e] += al * w.dx[e];
            FOR_LANES(e, N * 2) w.u[e] += al * w.du[e];
            wsync();
        }

        // ---- final residual (adds the dynamics defect, which the iteration keeps at rounding level)
        double rdyn = 0.0;
        FOR_LANES(e, N * 6) {
            int k = e / 6, a = e % 6;
            const double* T = &w.ABC[k][0];
            double v = T[48 + a] + T[36 + a] * w.u[k * 2] + T[42 + a] * w.u[k * 2 + 1];
#pragma unroll
            for (int b = 0; b < 6; ++b) v += T[b * 6 + a] * w.x[k * 6 + b];
            rdyn = fmax(rdyn, fabs(w.x[(k + 1) * 6 + a] - v));
        }
        r_prim = fmax(r_prim, wmax(rdyn));
        info.status = status;
        info.iters = it;
        info.late = late;
        info.r_prim = r_prim;
        info.r_dual = r_dual;
        info.gap = mu;
        if (LMPC && lam_out) {
            FOR_SLOTS(r, row, R4) { lam_out[row] = g.lam[r]; }
        }
        if (slack_out) {
            FOR_SLOTS(r, row, R1) { slack_out[row] = g.s[r]; }
        }
        wsync();
    }
};

}  // namespace lmpc
