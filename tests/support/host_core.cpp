// tests/support/host_core.cpp — TEST INFRASTRUCTURE ONLY.
// Compiles racinglmpc_b200/csrc/ftocp_pdip.cuh as a 1-lane host emulation so that the solver's
// arithmetic can be checked against the oracle on a machine without a GPU.  Never loaded by the
// product package (racinglmpc_b200 fails loudly when its CUDA library is missing).
#include "../../racinglmpc_b200/csrc/ftocp_pdip.cuh"
#include <new>
#include <cstring>

using namespace lmpc;

template <int N, int M>
static int run(const FtocpConst* c, const double* abc, const double* ss, const double* qfun, const double* x0,
               const double* uold, double* xpred, double* upred, double* lam, double* slack, double* info_out, double* warm, int* warm_valid) {
    using P = Pdip<N, M, 2, 4>;
    typename P::W* w = new typename P::W();
    std::memcpy(w->ABC, abc, sizeof(double) * N * 54);
    if (M > 0) {
        std::memcpy(w->SS, ss, sizeof(double) * 6 * M);
        std::memcpy(w->Qfun, qfun, sizeof(double) * M);
    }
    w->uOld[0] = uold[0];
    w->uOld[1] = uold[1];
    SolveInfo info;
    P::solve(*w, *c, x0, info, lam, slack, warm, warm ? warm_valid : nullptr);
    std::memcpy(xpred, w->x, sizeof(double) * (N + 1) * 6);
    std::memcpy(upred, w->u, sizeof(double) * N * 2);
    info_out[0] = info.status; info_out[1] = info.iters; info_out[2] = info.r_prim; info_out[3] = info.r_dual; info_out[4] = info.gap;
    delete w;
    return info.status;
}

extern "C" int host_core_solve(int N, int M, const FtocpConst* c, const double* abc, const double* ss, const double* qfun,
                               const double* x0, const double* uold, double* xpred, double* upred, double* lam,
                               double* slack, double* info_out, double* warm, int* warm_valid) {
#define CASE(n, m) if (N == n && M == m) return run<n, m>(c, abc, ss, qfun, x0, uold, xpred, upred, lam, slack, info_out, warm, warm_valid);
    CASE(6, 0) CASE(12, 0) CASE(14, 0) CASE(24, 0) CASE(48, 0)
    CASE(6, 48) CASE(12, 48) CASE(14, 48) CASE(24, 48) CASE(48, 48)
#undef CASE
    return -1;
}
#ifdef LMPC_HOST_COUNT
extern "C" void host_core_counts(long* out3, int reset) {
    for (int i = 0; i < 3; ++i) { out3[i] = lmpc::g_host_count[i]; if (reset) lmpc::g_host_count[i] = 0; }
}
#endif
extern "C" int host_core_const_size() { return (int)sizeof(FtocpConst); }
