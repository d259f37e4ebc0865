// racinglmpc_b200/csrc/lmpc_b200.cu — kernels + C ABI (include/lmpc_b200.h) for sm_100a.
//
// K3/K4/K5 of SURVEY §2.3 fused in one kernel: one warp (= one 32-thread CTA) per FTOCP instance.
//   load   : stage model (A_k,B_k,C_k) and the selected safe set are staged into shared memory with
//            1-D bulk async copies (cp.async.bulk -> UBLKCP, completion on an mbarrier)
//   solve  : racinglmpc_b200/csrc/ftocp_pdip.cuh (Mehrotra PDIP + Riccati, fp64, all in smem/registers)
//   unpack : xPred/uPred/lambd/slack + zt = Succ_SS lam, zt_u = Succ_uSS lam
//            (PredictiveControllers.py:364-384), coalesced stores
// No tensor cores: the largest contraction is 6x6x8 in fp64 (tcgen05 has no f64 kind) — see DESIGN.md.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <string>
#include <new>

#include "../../include/lmpc_b200.h"
#include "ftocp_pdip.cuh"

using namespace lmpc;

// ------------------------------------------------------------------------------------------------
// device helpers: mbarrier + bulk async copy (PTX ISA: cp.async.bulk, sm_90+)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

struct FtocpArgs {
    int batch;
    const double* x0;     // [B,6]
    const double* uOld;   // [B,2]
    const double* abc;    // stage model
    long long abc_inst_stride, abc_stage_stride;   // doubles
    const double* SS;     // [B,6,M]
    const double* Qfun;   // [B,M]
    const double* SuccSS; // [B,6,M] or null
    const double* SuccU;  // [B,2,M] or null
    double* xPred;        // [B,N+1,6]
    double* uPred;        // [B,N,2]
    double* slack;        // [B,N*NCX] or null
    double* lambd;        // [B,M] or null
    double* slackT;       // [B,6] or null
    double* zt;           // [B,6] or null
    double* ztu;          // [B,2] or null
    int* status;
    int* iters;
    double* resid;        // [B,3]
};

template <int N, int M, int NCX, int NCU>
struct KernelSmem {
    Work<N, M, NCX, NCU> w;
    alignas(8) uint64_t bar;
};

template <int N, int M, int NCX, int NCU>
__global__ void __launch_bounds__(32) ftocp_kernel(const __grid_constant__ FtocpConst c, const FtocpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using KS = KernelSmem<N, M, NCX, NCU>;
    KS& ks = *reinterpret_cast<KS*>(smem_raw);
    auto& w = ks.w;
    const int b = blockIdx.x;
    if (b >= a.batch) return;
    const int lane = threadIdx.x;

    // ---- stage the instance's model into shared memory (TMA 1-D bulk copies) ----
    if (lane == 0) {
        mbar_init(&ks.bar, 1);
        uint32_t bytes = N * 54 * 8 + (M > 0 ? (6 * M + M) * 8 : 0);
        mbar_expect_tx(&ks.bar, bytes);
        const double* src = a.abc + (long long)b * a.abc_inst_stride;
        if (a.abc_stage_stride == 54) {
            bulk_g2s(&w.ABC[0][0], src, N * 54 * 8, &ks.bar);
        } else {
            for (int k = 0; k < N; ++k) bulk_g2s(&w.ABC[k][0], src + (long long)k * a.abc_stage_stride, 54 * 8, &ks.bar);
        }
        if (M > 0) {
            bulk_g2s(&w.SS[0], a.SS + (long long)b * 6 * M, 6 * M * 8, &ks.bar);
            bulk_g2s(&w.Qfun[0], a.Qfun + (long long)b * M, M * 8, &ks.bar);
        }
    }
    double x0[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) x0[i] = a.x0[(long long)b * 6 + i];
    if (lane < 2) w.uOld[lane] = a.uOld[(long long)b * 2 + lane];
    __syncwarp();
    mbar_wait(&ks.bar, 0);

    // ---- solve ----
    SolveInfo info;
    Pdip<N, M, NCX, NCU>::solve(w, c, x0, info, (M > 0) ? w.d4i : nullptr, a.slack ? a.slack + (long long)b * N * NCX : nullptr);

    // ---- unpack (PC.py:364-384) ----
    for (int e = lane; e < (N + 1) * 6; e += 32) a.xPred[(long long)b * (N + 1) * 6 + e] = w.x[e];
    for (int e = lane; e < N * 2; e += 32) a.uPred[(long long)b * N * 2 + e] = w.u[e];
    if (M > 0) {
        const double* lam = w.d4i;   // solve() left lambda here
        if (a.lambd)
            for (int e = lane; e < M; e += 32) a.lambd[(long long)b * M + e] = lam[e];
        for (int e = lane; e < 14; e += 32) {
            double v = 0.0;
            if (e < 6) {           // slackTerminal = SS lam - x_N
                if (a.slackT) {
                    for (int l = 0; l < M; ++l) v += w.SS[e * M + l] * lam[l];
                    a.slackT[(long long)b * 6 + e] = v - w.x[N * 6 + e];
                }
            } else if (e < 12) {   // zt = Succ_SS lam
                if (a.zt && a.SuccSS) {
                    const double* S = a.SuccSS + ((long long)b * 6 + (e - 6)) * M;
                    for (int l = 0; l < M; ++l) v += S[l] * lam[l];
                    a.zt[(long long)b * 6 + (e - 6)] = v;
                }
            } else {               // zt_u = Succ_uSS lam
                if (a.ztu && a.SuccU) {
                    const double* S = a.SuccU + ((long long)b * 2 + (e - 12)) * M;
                    for (int l = 0; l < M; ++l) v += S[l] * lam[l];
                    a.ztu[(long long)b * 2 + (e - 12)] = v;
                }
            }
        }
    }
    if (lane == 0) {
        a.status[b] = info.status;
        a.iters[b] = info.iters;
        a.resid[(long long)b * 3 + 0] = info.r_prim;
        a.resid[(long long)b * 3 + 1] = info.r_dual;
        a.resid[(long long)b * 3 + 2] = info.gap;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) return fail(LMPC_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct lmpc_handle {
    lmpc_params p;
    FtocpConst c;
    int batch, device, N, M;
    cudaStream_t stream;
    long long launches;
    // device buffers used by the *_host entry points
    double *d_x0, *d_uOld, *d_abc, *d_SS, *d_Qfun, *d_SuccSS, *d_SuccU;
    double *d_xPred, *d_uPred, *d_slack, *d_lambd, *d_slackT, *d_zt, *d_ztu, *d_resid;
    int *d_status, *d_iters;
};

static bool inv6(const double* A, double* Ai) {
    double m[6][12];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) { m[i][j] = A[i * 6 + j]; m[i][6 + j] = (i == j) ? 1.0 : 0.0; }
    for (int col = 0; col < 6; ++col) {
        int piv = col;
        for (int r = col + 1; r < 6; ++r) if (fabs(m[r][col]) > fabs(m[piv][col])) piv = r;
        if (fabs(m[piv][col]) < 1e-300) return false;
        if (piv != col) for (int j = 0; j < 12; ++j) { double t = m[col][j]; m[col][j] = m[piv][j]; m[piv][j] = t; }
        double d = 1.0 / m[col][col];
        for (int j = 0; j < 12; ++j) m[col][j] *= d;
        for (int r = 0; r < 6; ++r)
            if (r != col) { double f = m[r][col]; if (f != 0.0) for (int j = 0; j < 12; ++j) m[r][j] -= f * m[col][j]; }
    }
    for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Ai[i * 6 + j] = m[i][6 + j];
    return true;
}

static int build_const(const lmpc_params& p, FtocpConst& c) {
    memset(&c, 0, sizeof(c));
    for (int i = 0; i < 36; ++i) { c.Q2[i] = 2.0 * p.Q[i]; c.Qf2[i] = 2.0 * p.Qf[i]; }
    for (int i = 0; i < 4; ++i) c.R2[i] = 2.0 * p.R[i];
    for (int a = 0; a < 6; ++a) {
        double s = 0, sN = 0;
        for (int b = 0; b < 6; ++b) { s += p.Q[a * 6 + b] * p.xRef[b]; sN += p.Qf[a * 6 + b] * p.xRef[b]; }
        c.qx[a] = -2.0 * s;
        c.qxN[a] = -2.0 * sN;
    }
    c.dR2[0] = 2.0 * p.dR[0];
    c.dR2[1] = 2.0 * p.dR[1];
    c.qs2 = 2.0 * p.Qslack[0];
    c.ql = p.Qslack[1];
    for (int i = 0; i < p.ncx * 6; ++i) c.Fx[i] = p.Fx[i];
    for (int i = 0; i < p.ncx; ++i) c.bx[i] = p.bx[i];
    for (int i = 0; i < p.ncu * 2; ++i) c.Fu[i] = p.Fu[i];
    for (int i = 0; i < p.ncu; ++i) c.bu[i] = p.bu[i];
    if (p.numSS_Points > 0) {
        for (int i = 0; i < 36; ++i) c.T[i] = 2.0 * p.QterminalSlack[i];
        if (!inv6(c.T, c.Tinv)) return fail(LMPC_E_INVALID, "QterminalSlack is singular");
    } else {
        for (int i = 0; i < 6; ++i) c.T[i * 6 + i] = c.Tinv[i * 6 + i] = 1.0;
    }
    c.eps_res = p.eps_res > 0 ? p.eps_res : 1e-9;
    c.eps_gap = p.eps_gap > 0 ? p.eps_gap : 1e-11;
    c.d4_min = 1e-4;
    c.max_iter = p.max_iter > 0 ? p.max_iter : 40;
    return LMPC_OK;
}

template <int N, int M>
static int launch_t(lmpc_handle* h, const FtocpArgs& a) {
    using KS = KernelSmem<N, M, 2, 4>;
    auto kern = ftocp_kernel<N, M, 2, 4>;
    static thread_local int configured_dev = -1;
    if (configured_dev != h->device) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(KS)));
        configured_dev = h->device;
    }
    kern<<<a.batch, 32, sizeof(KS), h->stream>>>(h->c, a);
    CK(cudaGetLastError());
    h->launches += 1;
    return LMPC_OK;
}

static int launch(lmpc_handle* h, const FtocpArgs& a, bool lmpc_mode) {
    const int N = h->N, M = lmpc_mode ? h->M : 0;
#define LCASE(n, m) if (N == n && M == m) return launch_t<n, m>(h, a);
    LCASE(6, 0) LCASE(12, 0) LCASE(14, 0) LCASE(24, 0) LCASE(48, 0)
    LCASE(6, 48) LCASE(12, 48) LCASE(14, 48) LCASE(24, 48) LCASE(48, 48)
#undef LCASE
    return fail(LMPC_E_INVALID, "unsupported (N, numSS_Points): built for N in {6,12,14,24,48}, numSS_Points in {0,48}");
}

extern "C" {

const char* lmpc_last_error(void) { return g_err.c_str(); }

int lmpc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int lmpc_create(const lmpc_params* p, int batch, int device, lmpc_handle** out) {
    if (!p || !out || batch <= 0) return fail(LMPC_E_INVALID, "null argument or batch <= 0");
    if (p->ncx != 2 || p->ncu != 4)
        return fail(LMPC_E_INVALID, "this build supports ncx == 2 lane rows and ncu == 4 input rows (the reference's values)");
    for (int j = 0; j < p->ncu; ++j)
        if (!(p->bu[j] > 0.0)) return fail(LMPC_E_INVALID, "bu must be positive (u = 0 strictly feasible)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(LMPC_E_NODEVICE, "no CUDA device");
    if (device < 0 || device >= ndev) return fail(LMPC_E_INVALID, "bad device index");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(LMPC_E_NODEVICE, "liblmpc_b200 is built for sm_100a only");
    lmpc_handle* h = new (std::nothrow) lmpc_handle();
    if (!h) return fail(LMPC_E_INVALID, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->p = *p;
    h->batch = batch;
    h->device = device;
    h->N = p->N;
    h->M = p->numSS_Points;
    int rc = build_const(*p, h->c);
    if (rc != LMPC_OK) { delete h; return rc; }
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const size_t B = batch, N = p->N, M = p->numSS_Points > 0 ? p->numSS_Points : 1;
#define DALLOC(ptr, count) CK(cudaMalloc((void**)&h->ptr, sizeof(*h->ptr) * (count)))
    DALLOC(d_x0, B * 6); DALLOC(d_uOld, B * 2); DALLOC(d_abc, B * N * 54);
    DALLOC(d_SS, B * 6 * M); DALLOC(d_Qfun, B * M); DALLOC(d_SuccSS, B * 6 * M); DALLOC(d_SuccU, B * 2 * M);
    DALLOC(d_xPred, B * (N + 1) * 6); DALLOC(d_uPred, B * N * 2); DALLOC(d_slack, B * N * 2);
    DALLOC(d_lambd, B * M); DALLOC(d_slackT, B * 6); DALLOC(d_zt, B * 6); DALLOC(d_ztu, B * 2);
    DALLOC(d_resid, B * 3); DALLOC(d_status, B); DALLOC(d_iters, B);
#undef DALLOC
    *out = h;
    return LMPC_OK;
}

int lmpc_destroy(lmpc_handle* h) {
    if (!h) return LMPC_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    double* dbl[] = {h->d_x0, h->d_uOld, h->d_abc, h->d_SS, h->d_Qfun, h->d_SuccSS, h->d_SuccU, h->d_xPred, h->d_uPred,
                     h->d_slack, h->d_lambd, h->d_slackT, h->d_zt, h->d_ztu, h->d_resid};
    for (double* q : dbl) cudaFree(q);
    cudaFree(h->d_status);
    cudaFree(h->d_iters);
    cudaStreamDestroy(h->stream);
    delete h;
    return LMPC_OK;
}

int lmpc_sync(lmpc_handle* h) {
    if (!h) return fail(LMPC_E_INVALID, "null handle");
    CK(cudaStreamSynchronize(h->stream));
    return LMPC_OK;
}

void* lmpc_stream(lmpc_handle* h) { return h ? (void*)h->stream : nullptr; }
long long lmpc_kernel_launches(lmpc_handle* h) { return h ? h->launches : 0; }

static int check_align(const void* p, const char* what) {
    if (((uintptr_t)p) & 15) return fail(LMPC_E_INVALID, std::string(what) + " must be 16-byte aligned");
    return LMPC_OK;
}

int lmpc_solve_lmpc_dev(lmpc_handle* h, const double* x0, const double* uOld, const double* abc, long long abc_inst_stride,
                        long long abc_stage_stride, const double* SS_sel, const double* Qfun_sel, const double* Succ_SS,
                        const double* Succ_uSS, double* xPred, double* uPred, double* slack, double* lambd,
                        double* slackTerminal, double* zt, double* zt_u, int* status, int* iters, double* resid) {
    if (!h || !x0 || !uOld || !abc || !xPred || !uPred || !status || !iters || !resid) return fail(LMPC_E_INVALID, "null argument");
    const bool lm = (SS_sel != nullptr);
    if (lm && (h->M <= 0 || !Qfun_sel)) return fail(LMPC_E_INVALID, "handle was created without a safe set (numSS_Points == 0)");
    if ((abc_inst_stride % 2) || (abc_stage_stride % 2)) return fail(LMPC_E_INVALID, "abc strides must be even (16-byte rows)");
    int rc;
    if ((rc = check_align(abc, "abc")) != LMPC_OK) return rc;
    if (lm && ((rc = check_align(SS_sel, "SS_sel")) != LMPC_OK || (rc = check_align(Qfun_sel, "Qfun_sel")) != LMPC_OK)) return rc;
    CK(cudaSetDevice(h->device));
    FtocpArgs a;
    a.batch = h->batch;
    a.x0 = x0; a.uOld = uOld; a.abc = abc;
    a.abc_inst_stride = abc_inst_stride; a.abc_stage_stride = abc_stage_stride;
    a.SS = SS_sel; a.Qfun = Qfun_sel; a.SuccSS = Succ_SS; a.SuccU = Succ_uSS;
    a.xPred = xPred; a.uPred = uPred; a.slack = slack; a.lambd = lambd; a.slackT = slackTerminal;
    a.zt = zt; a.ztu = zt_u; a.status = status; a.iters = iters; a.resid = resid;
    return launch(h, a, lm);
}

int lmpc_solve_mpc_dev(lmpc_handle* h, const double* x0, const double* uOld, const double* abc, long long abc_inst_stride,
                       long long abc_stage_stride, double* xPred, double* uPred, double* slack, int* status, int* iters,
                       double* resid) {
    return lmpc_solve_lmpc_dev(h, x0, uOld, abc, abc_inst_stride, abc_stage_stride, nullptr, nullptr, nullptr, nullptr, xPred,
                               uPred, slack, nullptr, nullptr, nullptr, nullptr, status, iters, resid);
}

int lmpc_solve_lmpc_host(lmpc_handle* h, const double* x0, const double* uOld, const double* abc, long long abc_inst_stride,
                         long long abc_stage_stride, const double* SS_sel, const double* Qfun_sel, const double* Succ_SS,
                         const double* Succ_uSS, double* xPred, double* uPred, double* slack, double* lambd,
                         double* slackTerminal, double* zt, double* zt_u, int* status, int* iters, double* resid) {
    if (!h || !x0 || !uOld || !abc || !xPred || !uPred || !status || !iters || !resid) return fail(LMPC_E_INVALID, "null argument");
    const bool lm = (SS_sel != nullptr);
    if (lm && (h->M <= 0 || !Qfun_sel)) return fail(LMPC_E_INVALID, "handle was created without a safe set (numSS_Points == 0)");
    CK(cudaSetDevice(h->device));
    const size_t B = h->batch, N = h->N, M = h->M > 0 ? h->M : 1;
    cudaStream_t s = h->stream;
    const size_t D = sizeof(double);
    CK(cudaMemcpyAsync(h->d_x0, x0, B * 6 * D, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(h->d_uOld, uOld, B * 2 * D, cudaMemcpyHostToDevice, s));
    long long dis, dss;
    if (abc_inst_stride == 0 && abc_stage_stride == 0) {            // one shared LTI model
        CK(cudaMemcpyAsync(h->d_abc, abc, 54 * D, cudaMemcpyHostToDevice, s));
        dis = 0; dss = 0;
    } else if (abc_inst_stride == (long long)N * 54 && abc_stage_stride == 54) {
        CK(cudaMemcpyAsync(h->d_abc, abc, B * N * 54 * D, cudaMemcpyHostToDevice, s));
        dis = N * 54; dss = 54;
    } else if (abc_inst_stride == 0 && abc_stage_stride == 54) {     // one shared LTV model
        CK(cudaMemcpyAsync(h->d_abc, abc, N * 54 * D, cudaMemcpyHostToDevice, s));
        dis = 0; dss = 54;
    } else {
        return fail(LMPC_E_INVALID, "host entry supports abc strides (N*54,54), (0,54) or (0,0)");
    }
    if (lm) {
        CK(cudaMemcpyAsync(h->d_SS, SS_sel, B * 6 * M * D, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(h->d_Qfun, Qfun_sel, B * M * D, cudaMemcpyHostToDevice, s));
        if (Succ_SS) CK(cudaMemcpyAsync(h->d_SuccSS, Succ_SS, B * 6 * M * D, cudaMemcpyHostToDevice, s));
        if (Succ_uSS) CK(cudaMemcpyAsync(h->d_SuccU, Succ_uSS, B * 2 * M * D, cudaMemcpyHostToDevice, s));
    }
    int rc = lmpc_solve_lmpc_dev(h, h->d_x0, h->d_uOld, h->d_abc, dis, dss, lm ? h->d_SS : nullptr, lm ? h->d_Qfun : nullptr,
                                 (lm && Succ_SS) ? h->d_SuccSS : nullptr, (lm && Succ_uSS) ? h->d_SuccU : nullptr, h->d_xPred,
                                 h->d_uPred, slack ? h->d_slack : nullptr, (lm && lambd) ? h->d_lambd : nullptr,
                                 (lm && slackTerminal) ? h->d_slackT : nullptr, (lm && zt) ? h->d_zt : nullptr,
                                 (lm && zt_u) ? h->d_ztu : nullptr, h->d_status, h->d_iters, h->d_resid);
    if (rc != LMPC_OK) return rc;
    CK(cudaMemcpyAsync(xPred, h->d_xPred, B * (N + 1) * 6 * D, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(uPred, h->d_uPred, B * N * 2 * D, cudaMemcpyDeviceToHost, s));
    if (slack) CK(cudaMemcpyAsync(slack, h->d_slack, B * N * 2 * D, cudaMemcpyDeviceToHost, s));
    if (lm && lambd) CK(cudaMemcpyAsync(lambd, h->d_lambd, B * M * D, cudaMemcpyDeviceToHost, s));
    if (lm && slackTerminal) CK(cudaMemcpyAsync(slackTerminal, h->d_slackT, B * 6 * D, cudaMemcpyDeviceToHost, s));
    if (lm && zt && Succ_SS) CK(cudaMemcpyAsync(zt, h->d_zt, B * 6 * D, cudaMemcpyDeviceToHost, s));
    if (lm && zt_u && Succ_uSS) CK(cudaMemcpyAsync(zt_u, h->d_ztu, B * 2 * D, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(status, h->d_status, B * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(iters, h->d_iters, B * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(resid, h->d_resid, B * 3 * D, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return LMPC_OK;
}

int lmpc_solve_mpc_host(lmpc_handle* h, const double* x0, const double* uOld, const double* abc, long long abc_inst_stride,
                        long long abc_stage_stride, double* xPred, double* uPred, double* slack, int* status, int* iters,
                        double* resid) {
    return lmpc_solve_lmpc_host(h, x0, uOld, abc, abc_inst_stride, abc_stage_stride, nullptr, nullptr, nullptr, nullptr, xPred,
                                uPred, slack, nullptr, nullptr, nullptr, nullptr, status, iters, resid);
}

}  // extern "C"
