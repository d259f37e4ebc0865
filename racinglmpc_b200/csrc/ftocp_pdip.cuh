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

#define FOR_LANES(e, n) for (int e = LMPC_LANE; e < (n); e += LMPC_NLANE)
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
    // --- model (filled by the loader; ABC via cp.async.bulk) ---
    alignas(16) double ABC[N][54];  // per stage: A (36, row major a*6+b) | B (12, a*2+c) | C (6)
    alignas(16) double SS[6 * MM];  // SS[a*M + l]      (PC.py:411 SS_PointSelectedTot, 6 x M)
    double Qfun[MM];                // Qfun_SelectedTot (PC.py:412)
    double uOld[2];                 // OldInput         (PC.py:136,247)
    // --- iterate ---
    double x[(N + 1) * 6], u[N * 2];
    double dx[(N + 1) * 6], du[N * 2];
    // --- per-row quantities shared between lanes ---
    double Dt[N * NCX], ex[N * NCX];  // condensed lane-constraint Hessian weights / rhs
    double d2[N * NCU], eu[N * NCU];  // input-bound Hessian weights / rhs
    double nu1s[N * NCX], nu2s[N * NCU];  // multipliers staged for the costate sweep
    double d4i[MM];                   // 1 / max(nu4/lam, d4_min)
    // --- Riccati factor, per stage ---
    double Li[N][3];    // 1/L00, L10, 1/L11   (L = chol of the 2x2 input Hessian)
    double Z[N][12];    // L^-1 (B'Pxx + Pxv')A,  Z[r*6+c]
    double Zv[N][3];    // L^-1 diag(dR2): Zv00, Zv10, Zv11
    double z0[N][2];    // L^-1 g0 for the current right-hand side
    double ru[N][2];    // input-stationarity residual
    // --- sweep scratch ---
    double G[48];                     // Pxx [A B]   G[a*8+j]
    double Pxx[36], Pxv[12], Pvv[3];  // cost-to-go Hessian blocks
    double Sx[36], Y[12], Lam[3], g0[2], hx[6];
    double pb[2][8];                  // cost-to-go gradient (px | pv), double buffered
    double pi[2][6];                  // costate, double buffered
    // --- terminal block ---
    double Wm[36], Wi[36];
    double sbar[6], c1[6], yT[6], dyT[6];
    int flag;
};

// per-lane register state -------------------------------------------------------------------
template <int N, int M, int NCX, int NCU>
struct Regs {
    static constexpr int R1 = N * NCX, R2 = N * NCU, R4 = (M > 0 ? M : 1);
    // lane-constraint pair (k,i):  Fx_i x_k - s <= bx_i  and  s >= 0
    double s[NSLOT(R1)], nu1[NSLOT(R1)], nu3[NSLOT(R1)];
    double w1[NSLOT(R1)], rs[NSLOT(R1)], d1[NSLOT(R1)], hs[NSLOT(R1)], p1[NSLOT(R1)], p3[NSLOT(R1)];
    // input bound (k,j):  Fu_j u_k <= bu_j
    double nu2[NSLOT(R2)], w2[NSLOT(R2)], p2[NSLOT(R2)];
    // simplex multiplier l:  lam_l >= 0
    double lam[NSLOT(R4)], nu4[NSLOT(R4)], rl[NSLOT(R4)], p4[NSLOT(R4)], rho[NSLOT(R4)], dlam[NSLOT(R4)];
    double y1;
};

constexpr double CENTRALITY_GAMMA = 0.01;

struct SolveInfo {
    int status, iters;
    double r_prim, r_dual, gap;
};

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
LMPC_HD double dot6(const double* a, const double* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
LMPC_HD double step_bound(double v, double dv, double a) {
    // largest alpha keeping v + alpha dv >= 0
    return (dv < 0.0) ? fmin(a, -v / dv) : a;
}

template <int N, int M, int NCX, int NCU>
struct Pdip {
    using W = Work<N, M, NCX, NCU>;
    using RG = Regs<N, M, NCX, NCU>;
    static constexpr int R1 = N * NCX, R2 = N * NCU, R4 = (M > 0 ? M : 1);
    static constexpr bool LMPC = (M > 0);

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
        wsync();
        for (int k = 0; k < N; ++k) {   // roll the model out (dynamics hold from the start)
            FOR_LANES(a, 6) {
                const double* A = &w.ABC[k][a * 6];
                const double* B = &w.ABC[k][36 + a * 2];
                w.x[(k + 1) * 6 + a] = dot6(A, &w.x[k * 6]) + B[0] * w.u[k * 2] + B[1] * w.u[k * 2 + 1] + w.ABC[k][48 + a];
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
    static LMPC_HD double terminal_factor(W& w, RG& g, const FtocpConst& c) {
        double acc[6] = {0, 0, 0, 0, 0, 0}, dl = 0.0;
        FOR_SLOTS(r, row, R4) {
            double d4 = fmax(g.nu4[r] / g.lam[r], c.d4_min);
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
                    L[i * (i + 1) / 2 + i] = 1.0 / sqrt(v);          // store the reciprocal pivot
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
    // Start of a backward sweep: terminal cost-to-go into (Pxx,Pxv,Pvv | pb[0]) and pi_N.
    template <bool FACTOR>
    static LMPC_HD void backward_start(W& w, const FtocpConst& c, const double* c1) {
        if (FACTOR) {
            FOR_LANES(e, 36) w.Pxx[e] = c.Qf2[e] + (LMPC ? w.Wi[e] : 0.0);
            FOR_LANES(e, 12) w.Pxv[e] = 0.0;
            FOR_LANES(e, 3) w.Pvv[e] = 0.0;
            FOR_LANES(a, 6) {   // pi_N = -(Qf2 x_N + qxN + yT)
                double v = c.qxN[a] + (LMPC ? w.yT[a] : 0.0);
#pragma unroll
                for (int b = 0; b < 6; ++b) v += c.Qf2[a * 6 + b] * w.x[N * 6 + b];
                w.pi[N & 1][a] = -v;
            }
        }
        FOR_LANES(a, 8) {
            double v = 0.0;
            if (LMPC && a < 6) {
#pragma unroll
                for (int b = 0; b < 6; ++b) v += w.Wi[a * 6 + b] * c1[b];
            }
            w.pb[N & 1][a] = v;
        }
        wsync();
    }

    static LMPC_HD double rtx(const W& w, const FtocpConst& c, int k, int a) {
        double v = 0.0;
#pragma unroll
        for (int i = 0; i < NCX; ++i) v += c.Fx[i * 6 + a] * w.ex[k * NCX + i];
        return v;
    }
    static LMPC_HD double fu_t(const FtocpConst& c, const double* row, int r) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < NCU; ++j) v += c.Fu[j * 2 + r] * row[j];
        return v;
    }

    // Factorising backward sweep: Riccati matrices, gradient recursion for the predictor
    // right-hand side and the costate / input-residual recursion, four phases per stage.
    static LMPC_HD double backward_factor(W& w, const FtocpConst& c) {
        double ru_max = 0.0;
        for (int k = N - 1; k >= 0; --k) {
            const double* A = &w.ABC[k][0];
            const double* B = &w.ABC[k][36];
            const double* pn = w.pb[(k + 1) & 1];    // px | pv of stage k+1
            const double* pin = w.pi[(k + 1) & 1];   // pi_{k+1}
            // ---- phase a: G = Pxx [A B]
            FOR_LANES(e, 48) {
                int a = e >> 3, j = e & 7;
                double v = 0.0;
                if (j < 6) {
#pragma unroll
                    for (int b = 0; b < 6; ++b) v += w.Pxx[a * 6 + b] * A[b * 6 + j];
                } else {
#pragma unroll
                    for (int b = 0; b < 6; ++b) v += w.Pxx[a * 6 + b] * B[b * 2 + (j - 6)];
                }
                w.G[e] = v;
            }
            wsync();
            // ---- phase b: stage matrix blocks, gradient pieces, costate
            FOR_LANES(e, 21 + 12 + 3 + 2 + 6 + 6) {
                if (e < 21) {                      // Sx(a,b), a <= b
                    int a = 0, b = e;
                    while (b >= 6 - a) { b -= 6 - a; ++a; }
                    b += a;
                    double v = c.Q2[a * 6 + b];
#pragma unroll
                    for (int i = 0; i < NCX; ++i) v += w.Dt[k * NCX + i] * c.Fx[i * 6 + a] * c.Fx[i * 6 + b];
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc) v += A[cc * 6 + a] * w.G[cc * 8 + b];
                    w.Sx[a * 6 + b] = v;
                    w.Sx[b * 6 + a] = v;
                } else if (e < 33) {               // Y(r,b) = B'G_A + Pxv'A
                    int r = (e - 21) / 6, b = (e - 21) % 6;
                    double v = 0.0;
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc) v += B[cc * 2 + r] * w.G[cc * 8 + b] + w.Pxv[cc * 2 + r] * A[cc * 6 + b];
                    w.Y[r * 6 + b] = v;
                } else if (e < 36) {               // Lam: (0,0), (1,0), (1,1)
                    int idx = e - 33;
                    int r = (idx == 0) ? 0 : 1, q = (idx == 2) ? 1 : 0;
                    double v = c.R2[r * 2 + q] + w.Pvv[idx];
                    if (r == q) v += ((k < N - 1) ? 2.0 : 1.0) * c.dR2[r];
#pragma unroll
                    for (int j = 0; j < NCU; ++j) v += w.d2[k * NCU + j] * c.Fu[j * 2 + r] * c.Fu[j * 2 + q];
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc)
                        v += B[cc * 2 + r] * w.G[cc * 8 + 6 + q] + B[cc * 2 + r] * w.Pxv[cc * 2 + q] + w.Pxv[cc * 2 + r] * B[cc * 2 + q];
                    w.Lam[idx] = v;
                } else if (e < 38) {               // ru_k(r) and g0(r)
                    int r = e - 36;
                    double uk = w.u[k * 2 + r];
                    double up = (k == 0) ? w.uOld[r] : w.u[(k - 1) * 2 + r];
                    double v = c.R2[r * 2] * w.u[k * 2] + c.R2[r * 2 + 1] * w.u[k * 2 + 1] + c.dR2[r] * (uk - up);
                    if (k < N - 1) v += c.dR2[r] * (uk - w.u[(k + 1) * 2 + r]);
                    v += fu_t(c, &w.nu2s[k * NCU], r);
                    double bp = 0.0, bpi = 0.0;
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc) { bp += B[cc * 2 + r] * pn[cc]; bpi += B[cc * 2 + r] * pin[cc]; }
                    v -= bpi;
                    w.ru[k][r] = v;
                    w.g0[r] = v + fu_t(c, &w.eu[k * NCU], r) + bp + pn[6 + r];
                } else if (e < 44) {               // hx(a) = rtx + A'px
                    int a = e - 38;
                    double v = rtx(w, c, k, a);
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc) v += A[cc * 6 + a] * pn[cc];
                    w.hx[a] = v;
                } else {                           // pi_k(a)
                    int a = e - 44;
                    double v = c.qx[a];
#pragma unroll
                    for (int b = 0; b < 6; ++b) v += c.Q2[a * 6 + b] * w.x[k * 6 + b];
#pragma unroll
                    for (int i = 0; i < NCX; ++i) v += c.Fx[i * 6 + a] * w.nu1s[k * NCX + i];
                    double t = 0.0;
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc) t += A[cc * 6 + a] * pin[cc];
                    w.pi[k & 1][a] = t - v;
                }
            }
            wsync();
            // ---- phase c/d: 2x2 Cholesky (lane redundant), Z, Zv, z0
            {
                double l00s = w.Lam[0], l10 = w.Lam[1], l11s = w.Lam[2];
                bool bad = !(l00s > 0.0);
                if (bad) l00s = 1.0;
                double i00 = 1.0 / sqrt(l00s);
                l10 *= i00;
                double t = l11s - l10 * l10;
                if (!(t > 0.0)) { bad = true; t = 1.0; }
                double i11 = 1.0 / sqrt(t);
                ru_max = fmax(ru_max, fmax(fabs(w.ru[k][0]), fabs(w.ru[k][1])));
                FOR_LANES(e, 8) {
                    if (e < 6) {
                        double z0c = w.Y[e] * i00;
                        w.Z[k][e] = z0c;
                        w.Z[k][6 + e] = (w.Y[6 + e] - l10 * z0c) * i11;
                    } else if (e == 6) {
                        double a0 = w.g0[0] * i00;
                        w.z0[k][0] = a0;
                        w.z0[k][1] = (w.g0[1] - l10 * a0) * i11;
                    } else {
                        double zv00 = c.dR2[0] * i00;
                        w.Zv[k][0] = zv00;
                        w.Zv[k][1] = -l10 * zv00 * i11;
                        w.Zv[k][2] = c.dR2[1] * i11;
                        w.Li[k][0] = i00;
                        w.Li[k][1] = l10;
                        w.Li[k][2] = i11;
                        if (bad) w.flag = ST_NUMERICAL;
                    }
                }
            }
            wsync();
            // ---- phase e: cost-to-go of stage k
            {
                const double* Z = w.Z[k];
                const double zv00 = w.Zv[k][0], zv10 = w.Zv[k][1], zv11 = w.Zv[k][2];
                const double a0 = w.z0[k][0], a1 = w.z0[k][1];
                double* po = w.pb[k & 1];
                FOR_LANES(e, 36 + 12 + 3 + 8) {
                    if (e < 36) {
                        int a = e / 6, b = e % 6;
                        w.Pxx[e] = w.Sx[e] - Z[a] * Z[b] - Z[6 + a] * Z[6 + b];
                    } else if (e < 48) {
                        int a = (e - 36) >> 1, q = (e - 36) & 1;
                        w.Pxv[e - 36] = (q == 0) ? (Z[a] * zv00 + Z[6 + a] * zv10) : (Z[6 + a] * zv11);
                    } else if (e < 51) {
                        int idx = e - 48;
                        w.Pvv[idx] = (idx == 0) ? -(zv00 * zv00 + zv10 * zv10) : (idx == 1 ? -(zv10 * zv11) : -(zv11 * zv11));
                    } else {
                        int a = e - 51;
                        if (a < 6) po[a] = w.hx[a] - Z[a] * a0 - Z[6 + a] * a1;
                        else po[a] = (a == 6) ? (zv00 * a0 + zv10 * a1) : (zv11 * a1);
                    }
                }
            }
            wsync();
        }
        return ru_max;
    }

    // Gradient-only backward sweep for a new right-hand side (corrector), one phase per stage.
    static LMPC_HD void backward_rhs(W& w, const FtocpConst& c) {
        for (int k = N - 1; k >= 0; --k) {
            const double* A = &w.ABC[k][0];
            const double* B = &w.ABC[k][36];
            const double* pn = w.pb[(k + 1) & 1];
            double* po = w.pb[k & 1];
            double g0[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                double v = w.ru[k][r] + fu_t(c, &w.eu[k * NCU], r) + pn[6 + r];
#pragma unroll
                for (int cc = 0; cc < 6; ++cc) v += B[cc * 2 + r] * pn[cc];
                g0[r] = v;
            }
            const double a0 = g0[0] * w.Li[k][0];
            const double a1 = (g0[1] - w.Li[k][1] * a0) * w.Li[k][2];
            FOR_LANES(a, 8) {
                if (a < 6) {
                    double v = rtx(w, c, k, a);
#pragma unroll
                    for (int cc = 0; cc < 6; ++cc) v += A[cc * 6 + a] * pn[cc];
                    po[a] = v - w.Z[k][a] * a0 - w.Z[k][6 + a] * a1;
                } else if (a == 6) {
                    po[6] = w.Zv[k][0] * a0 + w.Zv[k][1] * a1;
                    w.z0[k][0] = a0;
                } else {
                    po[7] = w.Zv[k][2] * a1;
                    w.z0[k][1] = a1;
                }
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
            const double* A = &w.ABC[k][0];
            const double* B = &w.ABC[k][36];
            const double* Z = w.Z[k];
            const double* d = &w.dx[k * 6];
            double t0 = w.z0[k][0] - w.Zv[k][0] * dv0;
            double t1 = w.z0[k][1] - w.Zv[k][1] * dv0 - w.Zv[k][2] * dv1;
#pragma unroll
            for (int b = 0; b < 6; ++b) { t0 += Z[b] * d[b]; t1 += Z[6 + b] * d[b]; }
            const double du1 = -t1 * w.Li[k][2];
            const double du0 = (-t0 - w.Li[k][1] * du1) * w.Li[k][0];
            FOR_LANES(a, 8) {
                if (a < 6) w.dx[(k + 1) * 6 + a] = dot6(&A[a * 6], d) + B[a * 2] * du0 + B[a * 2 + 1] * du1;
                else w.du[k * 2 + (a - 6)] = (a == 6) ? du0 : du1;
            }
            dv0 = du0;
            dv1 = du1;
            wsync();
        }
    }

    // terminal recovery after a forward sweep: dyT (shared), dlam (registers); returns dy1
    static LMPC_HD double terminal_recover(W& w, RG& g, const double* c1, double beta, double delta) {
        double v[6], dyT[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) v[a] = w.dx[N * 6 + a] + c1[a];
        double sdy = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double t = 0.0;
#pragma unroll
            for (int b = 0; b < 6; ++b) t += w.Wi[a * 6 + b] * v[b];
            dyT[a] = t;
            sdy += w.sbar[a] * t;
        }
        const double dy1t = -beta / delta;
        FOR_SLOTS(r, row, R4) {
            double t = g.rho[r] - dy1t;
#pragma unroll
            for (int a = 0; a < 6; ++a) t += (w.SS[a * M + row] - w.sbar[a]) * dyT[a];
            g.dlam[r] = t * w.d4i[row];
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
        int it = 0, status = ST_MAX_ITER;
        double r_prim = 0.0, r_dual = 0.0, mu = 0.0;

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
                double d1 = g.nu1[r] / w1, d3 = g.nu3[r] / g.s[r];
                double hs = c.qs2 + d1 + d3;
                g.d1[r] = d1;
                g.hs[r] = hs;
                w.Dt[row] = d1 * (c.qs2 + d3) / hs;
                // predictor: rc1 = w1 nu1, rc3 = s nu3  ->  e1 = -nu1, rs + rc3/s = rs + nu3
                w.ex[row] = (-g.nu1[r] * (c.qs2 + d3) + d1 * (rs + g.nu3[r])) / hs;
                w.nu1s[row] = g.nu1[r];
            }
            FOR_SLOTS(r, row, R2) {
                int k = row / NCU, j = row % NCU;
                double w2 = c.bu[j] - (c.Fu[j * 2] * w.u[k * 2] + c.Fu[j * 2 + 1] * w.u[k * 2 + 1]);
                g.w2[r] = w2;
                comp += w2 * g.nu2[r];
                w.d2[row] = g.nu2[r] / w2;
                w.eu[row] = -g.nu2[r];           // predictor: -rc2/w2
                w.nu2s[row] = g.nu2[r];
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
                    g.rho[r] = -rl - g.nu4[r];   // predictor: rc4/lam = nu4
                }
                delta = terminal_factor(w, g, c);
                terminal_rhs(w, g, -rone, c1, beta);
            } else {
                wsync();
            }
            comp = wsum(comp);
            mu = comp / n_ineq;
            // ---- factorising backward sweep (also yields the input residual) ------------------
            backward_start<true>(w, c, c1);
            double ru_max = backward_factor(w, c);
            r_dual = fmax(wmax(rd_loc), ru_max);
            r_prim = fabs(rone);
            if (w.flag != 0) { status = w.flag; break; }
            if (r_prim <= c.eps_res && r_dual <= c.eps_res && mu <= c.eps_gap) { status = ST_SOLVED; break; }
            if (it >= c.max_iter) { status = ST_MAX_ITER; break; }

            // ---- predictor -----------------------------------------------------------------------
            forward(w);
            double dy1 = 0.0;
            if (LMPC) dy1 = terminal_recover(w, g, c1, beta, delta);
            double a_aff = 1.0;
            FOR_SLOTS(r, row, R1) {
                int k = row / NCX, i = row % NCX;
                double fdx = dot6(&c.Fx[i * 6], &w.dx[k * 6]);
                double gs = -g.rs[r] - g.nu1[r] - g.nu3[r];
                double ds = (gs + g.d1[r] * fdx) / g.hs[r];
                double dw1 = -fdx + ds;
                double dn1 = -g.nu1[r] - g.d1[r] * dw1;
                double dn3 = -g.nu3[r] - (g.nu3[r] / g.s[r]) * ds;
                a_aff = step_bound(g.w1[r], dw1, a_aff);
                a_aff = step_bound(g.s[r], ds, a_aff);
                a_aff = step_bound(g.nu1[r], dn1, a_aff);
                a_aff = step_bound(g.nu3[r], dn3, a_aff);
                g.p1[r] = dw1 * dn1;
                g.p3[r] = ds * dn3;
            }
            FOR_SLOTS(r, row, R2) {
                int k = row / NCU, j = row % NCU;
                double dw2 = -(c.Fu[j * 2] * w.du[k * 2] + c.Fu[j * 2 + 1] * w.du[k * 2 + 1]);
                double dn2 = -g.nu2[r] - (g.nu2[r] / g.w2[r]) * dw2;
                a_aff = step_bound(g.w2[r], dw2, a_aff);
                a_aff = step_bound(g.nu2[r], dn2, a_aff);
                g.p2[r] = dw2 * dn2;
            }
            if (LMPC) {
                FOR_SLOTS(r, row, R4) {
                    double dn4 = -g.nu4[r] - (g.nu4[r] / g.lam[r]) * g.dlam[r];
                    a_aff = step_bound(g.lam[r], g.dlam[r], a_aff);
                    a_aff = step_bound(g.nu4[r], dn4, a_aff);
                    g.p4[r] = g.dlam[r] * dn4;
                }
            }
            a_aff = wmin(a_aff);
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

            // ---- corrector right-hand sides --------------------------------------------------------
            FOR_SLOTS(r, row, R1) {
                double rc1 = g.w1[r] * g.nu1[r] + g.p1[r] - sm;
                double rc3 = g.s[r] * g.nu3[r] + g.p3[r] - sm;
                double d3 = g.nu3[r] / g.s[r];
                double e1 = -rc1 / g.w1[r];
                w.ex[row] = (e1 * (c.qs2 + d3) + g.d1[r] * (g.rs[r] + rc3 / g.s[r])) / g.hs[r];
                g.p1[r] = rc1;   // keep rc for the final recovery
                g.p3[r] = rc3;
            }
            FOR_SLOTS(r, row, R2) {
                double rc2 = g.w2[r] * g.nu2[r] + g.p2[r] - sm;
                w.eu[row] = -rc2 / g.w2[r];
                g.p2[r] = rc2;
            }
            if (LMPC) {
                FOR_SLOTS(r, row, R4) {
                    double rc4 = g.lam[r] * g.nu4[r] + g.p4[r] - sm;
                    g.rho[r] = -g.rl[r] - rc4 / g.lam[r];
                    g.p4[r] = rc4;
                }
                terminal_rhs(w, g, -rone, c1, beta);
            }
            wsync();
            backward_start<false>(w, c, c1);
            backward_rhs(w, c);
            forward(w);
            if (LMPC) dy1 = terminal_recover(w, g, c1, beta, delta);

            // ---- step length and update ----------------------------------------------------------
            double amax = 1e300;
            double ds_[NSLOT(R1)], dw1_[NSLOT(R1)], dn1_[NSLOT(R1)], dn3_[NSLOT(R1)], dw2_[NSLOT(R2)], dn2_[NSLOT(R2)], dn4_[NSLOT(R4)];
            FOR_SLOTS(r, row, R1) {
                int k = row / NCX, i = row % NCX;
                double fdx = dot6(&c.Fx[i * 6], &w.dx[k * 6]);
                double rc1 = g.p1[r], rc3 = g.p3[r];
                double gs = -g.rs[r] - rc1 / g.w1[r] - rc3 / g.s[r];
                double ds = (gs + g.d1[r] * fdx) / g.hs[r];
                double dw1 = -fdx + ds;
                double dn1 = (-rc1 - g.nu1[r] * dw1) / g.w1[r];
                double dn3 = (-rc3 - g.nu3[r] * ds) / g.s[r];
                amax = step_bound(g.w1[r], dw1, amax);
                amax = step_bound(g.s[r], ds, amax);
                amax = step_bound(g.nu1[r], dn1, amax);
                amax = step_bound(g.nu3[r], dn3, amax);
                ds_[r] = ds; dw1_[r] = dw1; dn1_[r] = dn1; dn3_[r] = dn3;
            }
            FOR_SLOTS(r, row, R2) {
                int k = row / NCU, j = row % NCU;
                double dw2 = -(c.Fu[j * 2] * w.du[k * 2] + c.Fu[j * 2 + 1] * w.du[k * 2 + 1]);
                double dn2 = (-g.p2[r] - g.nu2[r] * dw2) / g.w2[r];
                amax = step_bound(g.w2[r], dw2, amax);
                amax = step_bound(g.nu2[r], dn2, amax);
                dw2_[r] = dw2; dn2_[r] = dn2;
            }
            if (LMPC) {
                FOR_SLOTS(r, row, R4) {
                    double dn4 = (-g.p4[r] - g.nu4[r] * g.dlam[r]) / g.lam[r];
                    amax = step_bound(g.lam[r], g.dlam[r], amax);
                    amax = step_bound(g.nu4[r], dn4, amax);
                    dn4_[r] = dn4;
                }
            }
            amax = wmin(amax);
            double al = fmin(1.0, 0.995 * amax);
            if (!(al > 0.0) || !(al <= 1.0)) { status = ST_NUMERICAL; break; }
            // Stay in a wide neighbourhood of the central path, min_i w_i nu_i >= gamma * mean: without it
            // Mehrotra steps can 2-cycle against a blocking bound (seen on 3 of 4096 workload QPs).
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
                if (pmin >= CENTRALITY_GAMMA * psum / n_ineq) break;
                al *= 0.8;
            }
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
            FOR_LANES(e, (N + 1) * 6) w.x[e] += al * w.dx[e];
            FOR_LANES(e, N * 2) w.u[e] += al * w.du[e];
            wsync();
        }

        // ---- final residual (adds the dynamics defect, which the iteration keeps at rounding level)
        double rdyn = 0.0;
        FOR_LANES(e, N * 6) {
            int k = e / 6, a = e % 6;
            const double* A = &w.ABC[k][a * 6];
            const double* B = &w.ABC[k][36 + a * 2];
            double v = w.x[(k + 1) * 6 + a] - (dot6(A, &w.x[k * 6]) + B[0] * w.u[k * 2] + B[1] * w.u[k * 2 + 1] + w.ABC[k][48 + a]);
            rdyn = fmax(rdyn, fabs(v));
        }
        r_prim = fmax(r_prim, wmax(rdyn));
        info.status = status;
        info.iters = it;
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
