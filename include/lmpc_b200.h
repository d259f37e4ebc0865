/* include/lmpc_b200.h — C ABI of liblmpc_b200.so (B200 / sm_100a batched LMPC hot path).
 *
 * The reference (urosolia/RacingLMPC) is pure Python and has no FFI; its drop-in boundary is the
 * duck-typed controller protocol  solve(x0) / addPoint(x,u) / addTrajectory(x,u,x_glob)  with results
 * read back as attributes (src/fnc/simulator/SysModel.py:34-38, src/main.py:110,117).  This header is
 * the native surface a ctypes shim (racinglmpc_b200/PredictiveControllers.py) binds to reproduce that
 * protocol for a BATCH of independent controllers.  Each entry point cites the reference code it
 * replaces.  Conventions:
 *   - plain C types only; all arrays are contiguous, row-major, fp64 unless stated, int32 indices;
 *   - `_host` entry points take HOST pointers and perform the H2D/D2H copies themselves (pinned staging
 *     buffers inside the handle); `_dev` entry points take DEVICE pointers and only enqueue work on the
 *     handle's stream (call lmpc_sync() before reading results);
 *   - return value 0 = ok, negative = error code below; no exceptions, no global state; a handle is
 *     not thread-safe; the library never falls back to a CPU path.
 */
#ifndef LMPC_B200_H
#define LMPC_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define LMPC_OK 0
#define LMPC_E_INVALID -1      /* bad argument / unsupported size                     */
#define LMPC_E_CUDA -2         /* CUDA runtime error (lmpc_last_error() has the text) */
#define LMPC_E_NODEVICE -3     /* no sm_100-class device available                    */
#define LMPC_E_STATE -4        /* call sequence error (e.g. solve before addTrajectory) */

/* per-instance solver status written to status[] (racinglmpc_b200/csrc/ftocp_pdip.cuh) */
#define LMPC_ST_SOLVED 1       /* == OSQP status_val 1 -> reference `feasible = 1` (PC.py:279-282) */
#define LMPC_ST_MAX_ITER 2
#define LMPC_ST_NUMERICAL 3
#define LMPC_ST_BAD_INPUT 4

#define LMPC_MAX_NCX 4
#define LMPC_MAX_NCU 8
#define LMPC_MAX_SEG 16

/* Controller parameters.  Field names/meaning follow MPCParams (PredictiveControllers.py:24-51) and the
 * extra LMPC constructor arguments (PredictiveControllers.py:293).  n = 6, d = 2 are fixed (the
 * reference's vehicle model, PredictiveModel.py:28-30 is hard-wired to them as well). */
typedef struct lmpc_params {
    int N;                        /* horizon                                   MPCParams.N        */
    int ncx, ncu;                 /* rows of Fx (<=4) and Fu (<=8)                                */
    double Q[36], R[4], Qf[36];   /* stage / terminal weights                  MPCParams.Q,R,Qf   */
    double dR[2];                 /* input-rate weight                         MPCParams.dR       */
    double Qslack[2];             /* [quadratic, linear] lane-slack cost (order as used at PC.py:249-250) */
    double xRef[6];               /*                                           MPCParams.xRef     */
    double Fx[LMPC_MAX_NCX * 6], bx[LMPC_MAX_NCX];   /* Fx x <= bx (soft)      MPCParams.Fx,bx    */
    double Fu[LMPC_MAX_NCU * 2], bu[LMPC_MAX_NCU];   /* Fu u <= bu             MPCParams.Fu,bu    */
    int numSS_Points, numSS_it;   /* 0,0 for a plain MPC                       LMPC.__init__      */
    double QterminalSlack[36];    /*                                           LMPC.__init__      */
    /* interior-point settings (no reference counterpart; OSQP's eps are 1e-3 + polish, PC.py:275) */
    double eps_res, eps_gap;      /* <= 0 selects the defaults 1e-9 / 1e-11    */
    int max_iter;                 /* <= 0 selects the default 40               */
} lmpc_params;

/* Local-regression model parameters: PredictiveModel.__init__ (PredictiveModel.py:12-32) and the
 * track table consumed by Map.curvature (Track.py:292-310). */
typedef struct lmpc_model_params {
    int trToUse;                  /* number of stored laps used     PredictiveModel.usedIt     */
    int MaxNumPoint;              /* 7                              PredictiveModel.MaxNumPoint */
    double h, lamb, dt;           /* 5, 0.0, 0.1                                               */
    double scaling[5];            /* diag(0.1,1,1,1,1)              PredictiveModel.scaling     */
    int nseg;                     /* rows of the track table                                   */
    double seg[LMPC_MAX_SEG * 3]; /* [s_start, length, curvature] = Map.PointAndTangent[:,3:6]  */
    double TrackLength;
} lmpc_model_params;

typedef struct lmpc_handle lmpc_handle;

const char* lmpc_last_error(void);
int lmpc_device_count(void);

/* One handle = `batch` independent controllers with identical parameters on one device. */
int lmpc_create(const lmpc_params* p, int batch, int device, lmpc_handle** out);
int lmpc_destroy(lmpc_handle* h);
int lmpc_sync(lmpc_handle* h);
void* lmpc_stream(lmpc_handle* h);            /* cudaStream_t the handle enqueues on */
long long lmpc_kernel_launches(lmpc_handle* h); /* kernels launched by this handle so far */

/* ---- FTOCP solve with the model given by the caller ("fixed A/B/C") ---------------------------------
 * Replaces, per instance: buildCost + buildEqConstr + addTerminalComponents + osqp_solve_qp +
 * unpackSolution (PredictiveControllers.py:110-137,200-283) for an MPC-type problem (no safe set).
 *   x0[B,6], uOld[B,2]            current state, previously applied input (OldInput, PC.py:136)
 *   abc                           stage model, per stage [A(36) row-major | B(12) | C(6)] = 54 doubles
 *   abc_inst_stride / abc_stage_stride   in doubles; (N*54, 54) = per-instance LTV (PC.py:212-215),
 *                                 (0, 0) = one LTI model shared by all (PC.py:216-218, C = 0 is the caller's job)
 * Outputs (caller-allocated): xPred[B,N+1,6], uPred[B,N,2] (PC.py:163-164), slack[B,N*ncx] or NULL,
 *   status[B] int32, iters[B] int32, resid[B,3] = (r_prim, r_dual, gap) in the unscaled inf-norm.
 */
int lmpc_solve_mpc_host(lmpc_handle* h, const double* x0, const double* uOld, const double* abc,
                        long long abc_inst_stride, long long abc_stage_stride, double* xPred, double* uPred,
                        double* slack, int* status, int* iters, double* resid);
int lmpc_solve_mpc_dev(lmpc_handle* h, const double* x0, const double* uOld, const double* abc,
                       long long abc_inst_stride, long long abc_stage_stride, double* xPred, double* uPred,
                       double* slack, int* status, int* iters, double* resid);

/* ---- FTOCP solve with a sampled safe set given by the caller -------------------------------------------
 * Adds addSafeSetEqConstr/addSafeSetCost + LMPC.unpackSolution + feasibleStateInput
 * (PredictiveControllers.py:345-384).  M = numSS_Points.
 *   SS_sel[B,6,M], Qfun_sel[B,M]          SS_PointSelectedTot, Qfun_SelectedTot (PC.py:411-412)
 *   Succ_SS[B,6,M], Succ_uSS[B,2,M]       successor states/inputs (PC.py:409-410), may be NULL
 * Extra outputs: lambd[B,M] (PC.py:374), slackTerminal[B,6] (PC.py:375) or NULL,
 *   zt[B,6], zt_u[B,2] (PC.py:383-384) or NULL when Succ_* are NULL.
 */
int lmpc_solve_lmpc_host(lmpc_handle* h, const double* x0, const double* uOld, const double* abc,
                         long long abc_inst_stride, long long abc_stage_stride, const double* SS_sel,
                         const double* Qfun_sel, const double* Succ_SS, const double* Succ_uSS, double* xPred,
                         double* uPred, double* slack, double* lambd, double* slackTerminal, double* zt,
                         double* zt_u, int* status, int* iters, double* resid);
int lmpc_solve_lmpc_dev(lmpc_handle* h, const double* x0, const double* uOld, const double* abc,
                        long long abc_inst_stride, long long abc_stage_stride, const double* SS_sel,
                        const double* Qfun_sel, const double* Succ_SS, const double* Succ_uSS, double* xPred,
                        double* uPred, double* slack, double* lambd, double* slackTerminal, double* zt,
                        double* zt_u, int* status, int* iters, double* resid);

#ifdef __cplusplus
}
#endif
#endif /* LMPC_B200_H */
