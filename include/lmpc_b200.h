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
    double eps_res, eps_gap;      /* <= 0 selects the defaults 1e-9 / 1e-11 (unscaled inf-norms); an instance still
                                     running at iteration 20 is reported solved once all three are <= 1e-6 (safety
                                     net; counted by lmpc_late_accepts, 0 on every recorded workload) */
    int max_iter;                 /* <= 0 selects the default 40               */
    int warm_start;               /* != 0: the controllers of the device-resident step (lmpc_step_* / lmpc_rollout_step) start each
                                     solve from a mid-path interior-point iterate of their previous solve, shifted by one stage
                                     (SURVEY §8f rank 2).  The reference always starts cold (PC.py:124,276: `initvals` unused);
                                     the optimum returned is the same, only the iteration count changes.  0 = cold start. */
    double eps_step;              /* <= 0 selects the default 1e-7: besides the residuals, the last primal step
                                     |alpha (dx, du)|_inf must be below it -- on QPs without strict complementarity the
                                     iterate is O(sqrt(gap)) from the optimum while the residuals are already tiny */
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
/* Number of QPs (since lmpc_create) that were reported LMPC_ST_SOLVED by the late-acceptance safety net of the interior-point
 * loop -- residuals and gap <= 1e-6 at iteration >= 20 / at max_iter -- instead of at eps_res / eps_gap (see lmpc_params).
 * Synchronises the handle's streams; -1 on error.  0 on every recorded BASELINE workload (tests assert it). */
long long lmpc_late_accepts(lmpc_handle* h);

/* Pinned host memory for the arrays handed to the `_host` / `_host_async` entry points (no reference counterpart: NumPy arrays
 * in the reference never leave the host).  The pages are placed on the NUMA node the device's PCIe link hangs off
 * (/sys/bus/pci/devices/<bus id>/numa_node; an anonymous mapping bound with mbind(MPOL_PREFERRED), touched, then pinned with
 * cudaHostRegister), which is what a single DMA stream needs to run at link rate on a multi-socket host; zero-filled.
 * lmpc_host_numa_node: that node, -1 = unknown / switched off with LMPC_B200_NUMA=off (LMPC_B200_NUMA=<n> forces node n).
 * lmpc_host_page_nodes: node of n (<= 64) evenly sampled pages of a block (diagnostics).  Plain pinned or pageable arrays remain
 * valid arguments of every entry point. */
int lmpc_host_alloc(int device, size_t bytes, void** out);
int lmpc_host_free(void* p);
int lmpc_host_numa_node(int device);
int lmpc_host_page_nodes(const void* p, size_t bytes, int n, int* nodes_out);

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
/* Asynchronous forms of the two *_host entry points, for streaming batch after batch: the call enqueues the copies and the
 * solve on buffer set `slot` (0 .. 3; the further sets of device buffers are allocated on first use) and returns;
 * lmpc_host_wait(h, slot) returns once the results of that slot are in the caller's output arrays.  With two batches in
 * flight the H2D copy of one overlaps the solve of the other; a third keeps the copy engine busy while the host waits for the
 * oldest batch and enqueues the next (measured: 8.7 -> see DESIGN §4 M solves/s on configs[1]).  All host arrays must stay valid (and should be pinned) until
 * the wait; a slot must be waited for before it is reused.  lmpc_solve_*_host == *_async(slot 0) + lmpc_host_wait(0). */
int lmpc_solve_mpc_host_async(lmpc_handle* h, int slot, const double* x0, const double* uOld, const double* abc,
                              long long abc_inst_stride, long long abc_stage_stride, double* xPred, double* uPred,
                              double* slack, int* status, int* iters, double* resid);
int lmpc_solve_lmpc_host_async(lmpc_handle* h, int slot, const double* x0, const double* uOld, const double* abc,
                               long long abc_inst_stride, long long abc_stage_stride, const double* SS_sel,
                               const double* Qfun_sel, const double* Succ_SS, const double* Succ_uSS, double* xPred,
                               double* uPred, double* slack, double* lambd, double* slackTerminal, double* zt,
                               double* zt_u, int* status, int* iters, double* resid);
int lmpc_host_wait(lmpc_handle* h, int slot);
/* Instance ranges (1 .. 4, each on its own stream: H2D | solve | D2H overlap inside one batch) the most recent enqueue on `slot`
 * was cut into: 4 for a batch of >= 2048 QPs solved through the synchronous entry points; through the asynchronous ones 1 when
 * another buffer set is in flight (the overlap then comes from the other batches, and fewer, larger pieces cost the enqueueing
 * thread less), else 2.
 * Diagnostics; -1 on a bad argument. */
int lmpc_host_chunks(lmpc_handle* h, int slot);
int lmpc_solve_lmpc_dev(lmpc_handle* h, const double* x0, const double* uOld, const double* abc,
                        long long abc_inst_stride, long long abc_stage_stride, const double* SS_sel,
                        const double* Qfun_sel, const double* Succ_SS, const double* Succ_uSS, double* xPred,
                        double* uPred, double* slack, double* lambd, double* slackTerminal, double* zt,
                        double* zt_u, int* status, int* iters, double* resid);

/* ===== device-resident lap stores, k-NN regression, safe-set selection, fused controller step =========
 *
 * lmpc_store_create: allocates, per instance, `ss_cap` safe-set lap slots (x[T,6], u[T,2], Qfun[T]) and
 * `model_cap` regression lap slots (x, u) of at most Tmax rows, plus the controller state
 * (xLin, uLin, zt, OldInput, timeStep, previous xPred) that MPC/LMPC keep between solves (PC.py:88-93,129-137,330).
 * Which slots are "the numSS_it fastest laps" (PC.py:395-402) and "usedIt" (PredictiveModel.py:31,35-46) is
 * decided by the host shim at addTrajectory time (once per lap) and passed as small index arrays.
 */
int lmpc_store_create(lmpc_handle* h, const lmpc_model_params* mp, int ss_cap, int model_cap, int Tmax);

/* PredictiveModel.addTrajectory (PredictiveModel.py:35-46): copy one lap into a regression slot. */
int lmpc_model_put_lap(lmpc_handle* h, int inst, int slot, int T, const double* x, const double* u);
/* usedIt order: slots[B,trToUse]. */
int lmpc_model_set_used(lmpc_handle* h, const int* slots);

/* LMPC.addTrajectory (PC.py:418-445): copy one lap into a safe-set slot; qfun == NULL computes
 * LMPC.computeCost (PC.py:447-464) on the device. */
int lmpc_ss_put_lap(lmpc_handle* h, int inst, int slot, int T, const double* x, const double* u, const double* qfun);
/* Selection for the coming lap: slots[B,numSS_it] in argsort(LapTime) order (PC.py:395,402); is_prev[B,numSS_it] = 1
 * where that lap is iteration it-1 (PC.py:506-512); prev_slot[B] = slot receiving addPoint rows (-1: none). */
int lmpc_ss_set_selection(lmpc_handle* h, const int* slots, const int* is_prev, const int* prev_slot);
/* LMPC.addPoint (PC.py:466-476) for every instance: x[B,6], u[B,2] (host). */
int lmpc_ss_add_point(lmpc_handle* h, const double* x, const double* u);
/* Read a safe-set lap back (tests, plotting): x[Tmax,6], u[Tmax,2], qfun[Tmax] caller-allocated, *T rows valid. */
int lmpc_ss_get_lap(lmpc_handle* h, int inst, int slot, int* T, double* x, double* u, double* qfun);
/* Overwrite one stored state row (used to reproduce the aliased write of PC.py:394 on the very first solve). */
int lmpc_ss_patch_row(lmpc_handle* h, int inst, int slot, int row, const double* x6, int also_model_slot);

/* Controller state (host arrays; any pointer may be NULL = leave / skip):
 * xLin[B,N+1,6], uLin[B,N,2] (PC.py:89-90,131-133,432-433), zt[B,6] (PC.py:330,383), OldInput[B,2] (PC.py:93,136),
 * timeStep[B] (PC.py:107,137,445), has_pred[B] (`self.xPred == []`, PC.py:502), xPred[B,N+1,6]. */
int lmpc_state_set(lmpc_handle* h, const double* xLin, const double* uLin, const double* zt, const double* OldInput,
                   const int* timeStep, const int* has_pred, const double* xPred);
int lmpc_state_get(lmpc_handle* h, double* xLin, double* uLin, double* zt, double* OldInput, int* timeStep);

/* K1 alone — MPC.computeLTVdynamics (PC.py:140-145) from the stored xLin/uLin: abc_out[B,N,54] (may be NULL: result
 * stays on the device for the next solve), flags[B] (0 ok; 1 singular regression, 2 curvature lookup failed,
 * 4 single neighbour in a lap — cases where the reference raises). */
int lmpc_identify_host(lmpc_handle* h, double* abc_out, int* flags);
/* K2 alone — LMPC.addTerminalComponents (PC.py:386-416) for x0[B,6]: outputs as in lmpc_solve_lmpc_host plus
 * min_index[B,numSS_it] (argmin row per lap) and flags[B] (8 = window past the stored lap). */
int lmpc_select_host(lmpc_handle* h, const double* x0, double* SS_sel, double* Qfun_sel, double* Succ_SS,
                     double* Succ_uSS, int* min_index, int* flags);

/* One full controller step for the whole batch = MPC.solve / LMPC.solve (PC.py:110-137):
 *   mode 0: LTV-MPC   K1 -> QP -> shift          (timeVarying MPC, main.py:91-94)
 *   mode 1: LMPC      K1 -> K2 -> QP -> shift    (main.py:113-120)
 * x0[B,6] host.  Outputs (host, any may be NULL): xPred[B,N+1,6], uPred[B,N,2], lambd[B,M], zt[B,6], zt_u[B,2],
 * SS_sel[B,6,M] (for SSStoredPredTraj, PC.py:379), status[B], iters[B], resid[B,3], flags[B] (K1/K2 bits). */
int lmpc_step_host(lmpc_handle* h, int mode, const double* x0, double* xPred, double* uPred, double* lambd, double* zt,
                   double* zt_u, double* SS_sel, int* status, int* iters, double* resid, int* flags);
/* Same with x0 already on the device; results stay in the handle's device buffers (lmpc_device_buffer). */
int lmpc_step_dev(lmpc_handle* h, int mode, const double* x0_dev);
/* lmpc_step_dev with CUDA events between its kernels: ms4 = milliseconds of K1 (k-NN regression), K2 (safe-set selection, 0 for
 * mode 0), the QP kernel and the state shift.  Measurement support for per-kernel rooflines; synchronises. */
int lmpc_step_profile(lmpc_handle* h, int mode, const double* x0_dev, float* ms4);
/* QP status / iterations / residuals [B,3] / step flags of the most recent lmpc_step_dev or lmpc_rollout_step (any NULL). */
int lmpc_step_results(lmpc_handle* h, int* status, int* iters, double* resid, int* flags);
/* Inspection: copy `bytes` from the named device buffer (names as lmpc_device_buffer) at `offset_bytes` to host memory.
 * The caller is responsible for staying inside the buffer (sizes follow from batch, N and numSS_Points). */
int lmpc_read_buffer(lmpc_handle* h, const char* name, size_t offset_bytes, void* dst, size_t bytes);
/* Device address of an internal buffer by name: "xPred","uPred","lambd","zt","zt_u","abc","SS_sel","Qfun_sel",
 * "Succ_SS","Succ_uSS","status","iters","resid","flags","xLin","uLin","x0","OldInput","slack","slackT" (lane slacks [B,N*ncx] and
 * terminal slack [B,6] of the most recent step), "abc_lti". */
void* lmpc_device_buffer(lmpc_handle* h, const char* name);

/* ===== device-resident closed loop (the caller of the hot path) =======================================
 * Simulator.sim's loop body (SysModel.py:33-48) for every instance without host round trips:
 *   Controller.solve(x) -> u = uPred[0] -> Controller.addPoint(x,u) -> x+ = Simulator.dynModel(x, x_glob, u)
 * (100 explicit-Euler sub-steps of the dynamic bicycle model with Pacejka tyres, SysModel.py:56-147).
 * z_host[B,3] = standard-normal draws for the (vx, vy, wz) noise in the reference's order; NULL = Philox4x32-10 on the
 * device keyed by (seed, instance, step).  The visited (x, u) pairs are recorded per instance (Tcl rows). */
int lmpc_rollout_create(lmpc_handle* h, int Tcl);
int lmpc_rollout_set_state(lmpc_handle* h, const double* x, const double* xglob);              /* [B,6] host, may be NULL */
int lmpc_rollout_get_state(lmpc_handle* h, double* x, double* xglob, int* done, int* cl_len);  /* any may be NULL      */
int lmpc_rollout_step(lmpc_handle* h, int mode, const double* z_host, unsigned long long seed);   /* mode 0 LTV-MPC, 1 LMPC, 2 LTI-MPC */
/* Since lmpc_rollout_create: OR of the per-step flags (bits as in lmpc_step_host) and number of steps whose QP was not
 * reported solved (the reference would have driven on with feasible = 0, PC.py:279-283), per instance; either may be NULL. */
int lmpc_rollout_get_health(lmpc_handle* h, int* flags_or, int* unsolved_steps);                 /* [B] host             */
/* The rest of main.py's pipeline on the device.
 * lmpc_rollout_pid_step: one closed-loop step under the PID path follower (Utilities.py:42-68, target speed vt); z_pid_host[B,2]
 *   and z_sim_host[B,3] are the reference's standard-normal draws in its order, NULL = Philox.  1000 such steps from
 *   x = [0.5,0,0,0,0,0] are main.py:65-66's PID lap (the record must hold them: lmpc_rollout_create(h, Tcl >= 1000)).
 * lmpc_rollout_sysid: Regression(x, u, lamb) (Utilities.py:5-28) of every record -> device buffer "abc_lti" [B,54] = A | B | 0
 *   (and abc_host / flags_host when given; flag bit 1 = singular normal matrix).  lmpc_rollout_step(mode 2) then runs the LTI
 *   MPC of main.py:72-80 with it.
 * lmpc_rollout_seed_from_record: main.py:99-110 -- the record becomes `copies` identical laps in the safe set and in the
 *   regression model (the caller's lap books must say the same: lmpc_ss_set_selection / lmpc_model_set_used), xLin/uLin/zt/
 *   OldInput/timeStep are initialised as LMPC.__init__/addTrajectory do, the record restarts. */
int lmpc_rollout_pid_step(lmpc_handle* h, double vt, const double* z_pid_host, const double* z_sim_host, unsigned long long seed);
int lmpc_rollout_sysid(lmpc_handle* h, double lamb, double* abc_host, int* flags_host);
int lmpc_rollout_seed_from_record(lmpc_handle* h, int copies, int ss_slot0, int model_slot0);
int lmpc_rollout_get_lap(lmpc_handle* h, int inst, int* T, double* x, double* u);              /* closed-loop record   */
/* Lap hand-over on the device = LMPC.addTrajectory + PredictiveModel.addTrajectory of the lap just driven (either slot may
 * be -1), then s -= TrackLength (SysModel.py:50), record restarted, timeStep = 0 (PC.py:445). */
int lmpc_rollout_commit_lap(lmpc_handle* h, int inst, int ss_slot, int model_slot);
/* The same for every instance with fin_host[b] != 0 in one launch; slots per instance as above ([B] host arrays). */
int lmpc_rollout_commit_laps(lmpc_handle* h, const int* fin_host, const int* ss_slots_host, const int* model_slots_host);
/* Pack all closed-loop records into rows_dev[B,Tpad,8] = (x | u), lens_dev[B] (device buffers of the caller): the send
 * buffer of the once-per-lap all-gather of the pooled-safe-set mode. */
int lmpc_rollout_export_laps_dev(lmpc_handle* h, int Tpad, double* rows_dev, int* lens_dev);

/* ---- pooled-safe-set exchange (no reference counterpart: the reference runs one controller) ------------------
 * Send side: pack stored lap slots_host[b] (-1: none, lens = 0) of every instance, including the rows LMPC.addPoint
 * appended after the finish line (PC.py:466-476), into rows_dev[B,Tpad,9] = (x 6 | u 2 | Qfun 1), lens_dev[B] -- device
 * buffers of the caller, which all-gathers them (NCCL) across ranks.
 * Receive side: instance b stores gathered lap src_host[b] (index into rows_dev[n_src,Tpad,9]; -1: nothing) in safe-set slot
 * ss_slots_host[b] (-1: skip) with its Qfun, i.e. LMPC.addTrajectory (PC.py:418-445) of a lap driven by another
 * controller, and in regression-model slot model_slots_host[b] (NULL or -1: skip; rows up to the finish line only), i.e.
 * PredictiveModel.addTrajectory (PredictiveModel.py:35-46).  Which laps are "the numSS_it fastest" stays the caller's
 * bookkeeping (lmpc_ss_set_selection / lmpc_model_set_used), exactly as for laps driven locally. */
int lmpc_ss_export_laps_dev(lmpc_handle* h, const int* slots_host, int Tpad, double* rows_dev, int* lens_dev);
int lmpc_ss_import_laps_dev(lmpc_handle* h, const int* ss_slots_host, const int* model_slots_host, const int* src_host, int n_src,
                            int Tpad, const double* rows_dev, const int* lens_dev);

/* ---- lap bookkeeping on the device (racinglmpc_b200/csrc/lapbooks.cuh) -------------------------------------
 * The reference decides once per lap, from Python lists, which stored laps the next lap uses: np.argsort(LapTime)[:numSS_it]
 * (PC.py:395,402), lap it-1 (PC.py:466-476,506-512) and usedIt = the trToUse shortest laps (PredictiveModel.py:31,35-46).  The
 * host-driven entry points above take those decisions as index arrays (lmpc_ss_set_selection / lmpc_model_set_used); the
 * entry points below keep the books on the device instead, so that a Monte-Carlo batch hands laps over without the host.
 * Slot tables per controller: safe set ss_time[B,ss_cap], ss_lap[B,ss_cap] (lap number, -1 = free), it[B]; regression model
 * md_time[B,model_cap], md_seq[B,model_cap] (arrival number, -1 = free), md_cnt[B].  lmpc_books_set uploads them (e.g. after
 * seeding through the host-driven path); lmpc_books_get reads them back with the derived selection (any pointer may be NULL):
 * sel/is_prev[B,numSS_it], prev_slot[B], used[B,trToUse], lap_hist[B,16] = lengths of the laps driven since, lap_n[B]. */
int lmpc_books_set(lmpc_handle* h, const int* ss_time, const int* ss_lap, const int* it, const int* md_time, const int* md_seq,
                   const int* md_cnt);
int lmpc_books_get(lmpc_handle* h, int* ss_time, int* ss_lap, int* it, int* md_time, int* md_seq, int* md_cnt, int* sel, int* is_prev,
                   int* prev_slot, int* used, int* lap_hist, int* lap_n);
/* main.py:113-119 for every controller whose lap just ended (its done flag is set): LMPC.addTrajectory + PredictiveModel.addTrajectory
 * of the recorded lap with the bookkeeping above, s -= TrackLength, timeStep = 0.  Enqueues one kernel; no host data. */
int lmpc_rollout_commit_laps_dev(lmpc_handle* h);
/* main.py:99-110 with device books: the record becomes `copies` laps of both stores (books are reset first). */
int lmpc_rollout_seed_from_record_dev(lmpc_handle* h, int copies);
/* Progress of the batch: out4 = { min laps driven, max laps driven, min steps into the current lap among the controllers at the
 * minimum, controllers with a health flag }.  Synchronises. */
int lmpc_rollout_stats(lmpc_handle* h, int* out4);
/* Pooled safe-set exchange with device books.  Only the globally fastest laps are ever selected (PC.py:395), so a
 * rank ships its `kbest` fastest latest-own laps instead of all of them: rows_dev[kbest,Tpad,9] = (x | u | Qfun) incl. the rows
 * addPoint appended so far, meta_dev[kbest,4] = (rows, lap time, gid_base + controller, 0).  After the all-gather,
 * lmpc_pool_import_dev ranks the n_src <= 64 gathered laps by (lap time, global id) and every controller files the `share`
 * fastest laps it does not own before its own latest lap (addTrajectory of a lap another controller drove, into both stores).
 * took_host (may be NULL; non-NULL synchronises) receives the number of laps stored. */
int lmpc_pool_export_dev(lmpc_handle* h, int kbest, int Tpad, long long gid_base, double* rows_dev, int* meta_dev);
int lmpc_pool_import_dev(lmpc_handle* h, int n_src, int share, int Tpad, long long gid_base, const double* rows_dev, const int* meta_dev,
                         int* took_host);

/* ---- presentation support (the inputs of the reference's plot.py for device-resident batches) --------------
 * lmpc_track_global_position: Map.getGlobalPosition (Track.py:135-189) for n points; table6[nseg,6] = Map.PointAndTangent
 *   (x_end, y_end, psi_end, s_start, length, curvature), host arrays, xy[n,2] out, ok[n] (may be NULL) = 0 where the reference
 *   would raise (no segment holds s).
 * lmpc_rollout_trace_create: from now on every lmpc_rollout_step records, for the n chosen controllers, what
 *   LMPC.unpackSolution keeps per step for plotting (PC.py:377-379) -- the prediction xPred[N+1,6], the selected safe-set points
 *   SS_sel[6,M] -- plus the closed-loop state x[6], the global state x_glob[6] (SysModel.py), the applied input u[2] and the
 *   number of laps the controller had driven (to split the trace into laps), at most cap_steps rows each.
 * lmpc_rollout_trace_get: rows of trace `tr` (index into the chosen controllers); any array pointer may be NULL. */
int lmpc_track_global_position(int device, const double* table6, int nseg, double TrackLength, int n, const double* s, const double* ey,
                               double* xy, int* ok);
int lmpc_rollout_trace_create(lmpc_handle* h, int n, const int* inst, int cap_steps);
int lmpc_rollout_trace_get(lmpc_handle* h, int tr, int* steps, double* x, double* xglob, double* u, double* xPred, double* SS_sel, int* lap);

/* fp64 micro-benchmarks on `device` (no reference counterpart; measurement support for the roofline of the QP kernel, which is
 * bound by fp64 issue and dependent-chain latency rather than HBM): out8 = { DFMA TFLOP/s, DMMA m8n8k4 TFLOP/s,
 * latency in SM cycles of: DFMA, DMMA (accumulator chain), DMMA (result -> A operand), LDS (dependent), SHFL of a double,
 * rsqrt(double) }. */
int lmpc_probe_fp64(int device, double* out8);

int lmpc_sizeof_params(void);
int lmpc_sizeof_model_params(void);

#ifdef __cplusplus
}
#endif
#endif /* LMPC_B200_H */
