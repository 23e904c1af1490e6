// myr_fbsm_solve / myr_host_fbsm_solve: launch + dispatch of the forward-backward sweep (fbsm.cuh).  Its own translation
// unit: the sweep needs only the generated dynamics, not the NLP kernels of sys_unit.cu.
#pragma nv_diag_suppress 177  // generated system code declares symbols it may not use
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/myriad_b200.h"
#include "fbsm.cuh"

namespace myr {
int fail(int code, const char* fmt, const char* a = "", long long v = 0);  // api.cu
double system_default_T(int id);                                           // api.cu

// One warp per CTA: a warp is the unit that runs until its slowest instance has converged, registers (not the CTA count)
// bound the occupancy, and 32-thread CTAs spread a small batch over 4x as many SMs as 128-thread ones would.
constexpr int kFbsmThreads = 32;
#ifndef MYR_FBSM_MIN_CTAS
#define MYR_FBSM_MIN_CTAS 1  // resident CTAs per SM the register allocation must allow (A/B knob)
#endif

template <class Sys>
__global__ void __launch_bounds__(kFbsmThreads, MYR_FBSM_MIN_CTAS) fbsm_kernel(const __grid_constant__ FbsmParams P) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= P.B) return;
  FbsmInstance<Sys>(P, b).run();
}

template <class Sys>
static int fbsm_run(const MyrDesc* d, const MyrFbsmOpts* o, int B, const double* x0, const double* adj_T, const double* char_lb,
                    const double* char_ub, double* x, double* u, double* adj, int32_t* iters, int32_t* status, bool host, void* stream) {
  if (!Indirect<Sys>::available) return fail(MYR_E_UNSUPPORTED, "system %s has no adjoint ODE / optimality characterisation", Sys::name);
  if (B < 0 || d->intervals < 1) return fail(MYR_E_BADARG, "myr_fbsm_solve: B >= 0 and intervals >= 1 required%s", "");
  if (B == 0) return MYR_OK;
  if (!x0 || !x || !u || !adj || !iters || !status || !char_lb || !char_ub) return fail(MYR_E_BADARG, "myr_fbsm_solve: null array%s", "");
  FbsmParams P;
  P.B = B;
  P.N = d->intervals;
  P.T = d->T > 0 ? d->T : system_default_T(Sys::id);
  if (Indirect<Sys>::discrete && (double)P.N != P.T)  // forward_backward_sweep.py:33-35: N = int(T), h = 1
    return fail(MYR_E_BADARG, "discrete system %s: intervals must equal T (%lld given)", Sys::name, P.N);
  P.delta = (o && o->delta > 0) ? o->delta : 1e-3;             // stopping_criterion's default delta (base.py:129)
  P.secant_tol = (o && o->secant_tol > 0) ? o->secant_tol : 1e-10;  // forward_backward_sweep.py:137
  P.max_iter = (o && o->max_iter > 0) ? o->max_iter : 10000;
  P.max_secant = (o && o->max_secant > 0) ? o->max_secant : 100;
  P.term_state = o ? o->term_state : -1;
  if (P.term_state >= Sys::n) return fail(MYR_E_BADARG, "term_state out of range for %s (%lld)", Sys::name, P.term_state);
  P.term_value = o ? o->term_value : 0.0;
  P.guess_a = o ? o->guess_a : 0.0;
  P.guess_b = o ? o->guess_b : 0.0;
  Sys::default_params(P.p);
  if (d->n_params > 0) {
    if (d->n_params != Sys::np) return fail(MYR_E_BADARG, "n_params does not match system %s (%lld given)", Sys::name, d->n_params);
    for (int i = 0; i < Sys::np; ++i) P.p[i] = d->params[i];
  }
  for (int k = 0; k < 8; ++k) P.adj_T[k] = (adj_T && k < Sys::n) ? adj_T[k] : 0.0;
  for (int k = 0; k < 4; ++k) {
    P.lb[k] = k < Sys::m ? char_lb[k] : 0.0;
    P.ub[k] = k < Sys::m ? char_ub[k] : 0.0;
  }
  P.x0 = x0; P.x = x; P.u = u; P.adj = adj; P.iters = iters; P.status = status;
  if (host) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) FbsmInstance<Sys>(P, b).run();
    return MYR_OK;
  }
  const int threads = kFbsmThreads;
  fbsm_kernel<Sys><<<(B + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(P);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MYR_E_CUDA, "fbsm_kernel launch: %s", cudaGetErrorString(e));
  return MYR_OK;
}

static int fbsm_dispatch(const MyrDesc* d, const MyrFbsmOpts* o, int B, const double* x0, const double* adj_T, const double* char_lb,
                         const double* char_ub, double* x, double* u, double* adj, int32_t* iters, int32_t* status, bool host, void* stream) {
  if (!d) return fail(MYR_E_BADARG, "null descriptor%s", "");
#define MYR_FBSM_CASE(SYS) \
  case SYS::id: return fbsm_run<SYS>(d, o, B, x0, adj_T, char_lb, char_ub, x, u, adj, iters, status, host, stream);
  switch (d->system_id) {
    MYR_FBSM_CASE(SysSimplecase)
    MYR_FBSM_CASE(SysSimplecasewithbounds)
    MYR_FBSM_CASE(SysCancertreatment)
    MYR_FBSM_CASE(SysMouldfungicide)
    MYR_FBSM_CASE(SysBioreactor)
    MYR_FBSM_CASE(SysGlucose)
    MYR_FBSM_CASE(SysHarvest)
    MYR_FBSM_CASE(SysTimberharvest)
    MYR_FBSM_CASE(SysEpidemicseirn)
    MYR_FBSM_CASE(SysHivtreatment)
    MYR_FBSM_CASE(SysBacteria)
    MYR_FBSM_CASE(SysPredatorprey)
    MYR_FBSM_CASE(SysBearpopulations)
    MYR_FBSM_CASE(SysInvasiveplant)
    default:
      return fail(MYR_E_UNSUPPORTED, "FBSM needs a system with adj_ODE / optim_characterization (system_id %s%lld has none)", "",
                  d->system_id);
  }
#undef MYR_FBSM_CASE
}
}  // namespace myr

extern "C" int myr_fbsm_solve(const MyrDesc* desc, const MyrFbsmOpts* opts, int B, const double* x0, const double* adj_T,
                              const double* char_lb, const double* char_ub, double* x, double* u, double* adj, int32_t* iters,
                              int32_t* status, void* stream) {
  return myr::fbsm_dispatch(desc, opts, B, x0, adj_T, char_lb, char_ub, x, u, adj, iters, status, false, stream);
}
extern "C" int myr_host_fbsm_solve(const MyrDesc* desc, const MyrFbsmOpts* opts, int B, const double* x0, const double* adj_T,
                                   const double* char_lb, const double* char_ub, double* x, double* u, double* adj, int32_t* iters,
                                   int32_t* status) {
  return myr::fbsm_dispatch(desc, opts, B, x0, adj_T, char_lb, char_ub, x, u, adj, iters, status, true, nullptr);
}
