// C ABI (include/myriad_b200.h): error state and dispatch to the per-system tables (sys_unit.cu).
#include <stdio.h>
#include <string.h>

#include "../../include/myriad_b200.h"
#include "common.cuh"
#include "kernels_decl.h"

namespace myr {

static thread_local char g_err[512] = "";
int fail(int code, const char* fmt, const char* a, long long v) {
  snprintf(g_err, sizeof(g_err), fmt, a, v);
  return code;
}

double system_default_T(int id) {
  if (id >= MYR_SYS_NODE_BASE) id -= MYR_SYS_NODE_BASE;  // NodeSystem takes T from the true system
  switch (id) {  // T of each system's constructor
    case MYR_SYS_SIMPLECASE: return 1.0;
    case MYR_SYS_CARTPOLE: return 2.0;
    case MYR_SYS_VANDERPOL: return 10.0;
    case MYR_SYS_CANCERTREATMENT: return 20.0;
    case MYR_SYS_MOULDFUNGICIDE: return 5.0;
    case MYR_SYS_BIOREACTOR: return 2.0;
    case MYR_SYS_SIMPLECASEWITHBOUNDS: return 1.0;
    case MYR_SYS_GLUCOSE: return 0.2;
    case MYR_SYS_HARVEST: return 10.0;
    case MYR_SYS_TIMBERHARVEST: return 5.0;
    case MYR_SYS_SEIR: return 20.0;
    case MYR_SYS_EPIDEMICSEIRN: return 20.0;
    case MYR_SYS_HIVTREATMENT: return 20.0;
    case MYR_SYS_BACTERIA: return 1.0;
    case MYR_SYS_TUMOUR: return 1.2;
    case MYR_SYS_PREDATORPREY: return 10.0;
    case MYR_SYS_BEARPOPULATIONS: return 25.0;
    case MYR_SYS_ROCKETLANDING: return 16.0;
    case MYR_SYS_PENDULUM: return 15.0;
    case MYR_SYS_MOUNTAINCAR: return 300.0;
    case MYR_SYS_INVASIVEPLANT: return 10.0;
    default: return 1.0;
  }
}

}  // namespace myr

using namespace myr;

// Systems compiled into this build: MYR_BUILD_SYSTEMS(X) is passed by build.py, default = all generated.
#ifndef MYR_BUILD_SYSTEMS
#define MYR_BUILD_SYSTEMS(X) MYR_FOR_EACH_SYSTEM(X)
#endif

#define MYR_DECL(SYS) extern "C" const myr::SysVTable* myr_vtable_##SYS(void);
MYR_BUILD_SYSTEMS(MYR_DECL)
#undef MYR_DECL

// NodeSystem wrappers compiled into this build (build.py passes the list)
#ifndef MYR_BUILD_NODE_SYSTEMS
#define MYR_BUILD_NODE_SYSTEMS(X)
#endif
#define MYR_DECL(SYS) extern "C" const myr::SysVTable* myr_vtable_node_##SYS(void);
MYR_BUILD_NODE_SYSTEMS(MYR_DECL)
#undef MYR_DECL

// systems registered at run time (myr_register_system): one extra shared library per user-defined system, built by
// myriad_b200/plugin.py from the same generator and the same kernel templates as the built-in ones
static const SysVTable* g_user_systems[MYR_MAX_USER_SYSTEMS];
static int g_num_user_systems = 0;

extern "C" int myr_register_system(const void* vtable) {
  const SysVTable* vt = static_cast<const SysVTable*>(vtable);
  if (!vt || vt->id < MYR_SYS_USER_BASE) return fail(MYR_E_BADARG, "user systems need an id >= MYR_SYS_USER_BASE (got %s%lld)", "", vt ? vt->id : -1);
  for (int i = 0; i < g_num_user_systems; ++i)
    if (g_user_systems[i]->id == vt->id) { g_user_systems[i] = vt; return MYR_OK; }
  if (g_num_user_systems >= MYR_MAX_USER_SYSTEMS) return fail(MYR_E_UNSUPPORTED, "too many registered systems%s", "", 0);
  g_user_systems[g_num_user_systems++] = vt;
  return MYR_OK;
}

static const SysVTable* find_system(int id) {
  for (int i = 0; i < g_num_user_systems; ++i)
    if (g_user_systems[i]->id == id) return g_user_systems[i];
#define MYR_TRY(SYS) if (SYS::id == id) return myr_vtable_##SYS();
  MYR_BUILD_SYSTEMS(MYR_TRY)
#undef MYR_TRY
#define MYR_TRY(SYS) if (MYR_SYS_NODE_BASE + SYS::id == id) return myr_vtable_node_##SYS();
  MYR_BUILD_NODE_SYSTEMS(MYR_TRY)
#undef MYR_TRY
  fail(MYR_E_BADARG, "unknown or not-built system_id %s%lld", "", id);
  return nullptr;
}

#define MYR_GET(desc)                                                 \
  if (!(desc)) return fail(MYR_E_BADARG, "null descriptor%s", "", 0); \
  const SysVTable* vt = find_system((desc)->system_id);               \
  if (!vt) return MYR_E_BADARG;

extern "C" int myr_abi_version(void) { return MYR_ABI_VERSION; }
extern "C" const char* myr_last_error(void) { return g_err; }

extern "C" int myr_problem_sizes(const MyrDesc* desc, MyrSizes* out) {
  if (!out) return fail(MYR_E_BADARG, "null output%s", "", 0);
  memset(out, 0, sizeof(*out));
  MYR_GET(desc);
  return vt->sizes(desc, out);
}
extern "C" int myr_eval(const MyrDesc* desc, int B, const double* z, const double* lam, double* f, double* grad, double* c, double* Jblk,
                        double* Hblk, void* stream) {
  MYR_GET(desc);
  return vt->eval(desc, B, z, lam, f, grad, c, Jblk, Hblk, stream);
}
extern "C" int myr_host_eval(const MyrDesc* desc, int B, const double* z, const double* lam, double* f, double* grad, double* c,
                             double* Jblk, double* Hblk) {
  MYR_GET(desc);
  return vt->host_eval(desc, B, z, lam, f, grad, c, Jblk, Hblk);
}
extern "C" int myr_kkt_solve(const MyrDesc* desc, int B, const double* Hblk, const double* Jblk, const double* sigma, const double* rhs_z,
                             const double* rhs_c, double delta_w, double delta_c, double* dz, double* dlam, int32_t* inertia_ok, double* ws,
                             size_t ws_doubles, void* stream) {
  MYR_GET(desc);
  return vt->kkt(desc, B, Hblk, Jblk, sigma, rhs_z, rhs_c, delta_w, delta_c, dz, dlam, inertia_ok, ws, ws_doubles, stream);
}
extern "C" int myr_host_kkt_solve(const MyrDesc* desc, int B, const double* Hblk, const double* Jblk, const double* sigma,
                                  const double* rhs_z, const double* rhs_c, double delta_w, double delta_c, double* dz, double* dlam,
                                  int32_t* inertia_ok, double* ws, size_t ws_doubles) {
  MYR_GET(desc);
  return vt->host_kkt(desc, B, Hblk, Jblk, sigma, rhs_z, rhs_c, delta_w, delta_c, dz, dlam, inertia_ok, ws, ws_doubles);
}
extern "C" int myr_ipm_solve(const MyrDesc* desc, const MyrIpmOpts* opts, int B, const double* z0, const double* lb, const double* ub, double* z,
                             double* lam, double* zL, double* zU, double* obj, double* kkt_err, double* con_inf, int32_t* status, int32_t* iters,
                             double* ws, size_t ws_doubles, void* stream) {
  MYR_GET(desc);
  return vt->ipm(desc, opts, B, z0, lb, ub, z, lam, zL, zU, obj, kkt_err, con_inf, status, iters, ws, ws_doubles, stream);
}
extern "C" int myr_host_ipm_solve(const MyrDesc* desc, const MyrIpmOpts* opts, int B, const double* z0, const double* lb, const double* ub,
                                  double* z, double* lam, double* zL, double* zU, double* obj, double* kkt_err, double* con_inf,
                                  int32_t* status, int32_t* iters, double* ws, size_t ws_doubles) {
  MYR_GET(desc);
  return vt->host_ipm(desc, opts, B, z0, lb, ub, z, lam, zL, zU, obj, kkt_err, con_inf, status, iters, ws, ws_doubles);
}
extern "C" int myr_rollout_cost(const MyrDesc* desc, int B, int nu_rows, const double* u, const double* x0, double* xs, double* cost,
                                void* stream) {
  MYR_GET(desc);
  return vt->rollout(desc, B, nu_rows, u, x0, xs, cost, stream);
}
extern "C" int myr_host_rollout_cost(const MyrDesc* desc, int B, int nu_rows, const double* u, const double* x0, double* xs, double* cost) {
  MYR_GET(desc);
  return vt->host_rollout(desc, B, nu_rows, u, x0, xs, cost);
}

extern "C" int myr_dynamics(const MyrDesc* desc, int B, const double* x, const double* u, const double* t, double* f, double* g, void* stream) {
  MYR_GET(desc);
  return vt->dynamics(desc, B, x, u, t, f, g, stream);
}
extern "C" int myr_host_dynamics(const MyrDesc* desc, int B, const double* x, const double* u, const double* t, double* f, double* g) {
  MYR_GET(desc);
  return vt->host_dynamics(desc, B, x, u, t, f, g);
}

extern "C" int myr_jtvec(const MyrDesc* desc, int B, const double* Jblk, const double* lam, double* out, void* stream) {
  MYR_GET(desc);
  return vt->jtvec(desc, B, Jblk, lam, out, stream);
}
extern "C" int myr_host_jtvec(const MyrDesc* desc, int B, const double* Jblk, const double* lam, double* out) {
  MYR_GET(desc);
  return vt->host_jtvec(desc, B, Jblk, lam, out);
}

// ------------------------------------------------------------------ measurement helper
// fp64 FMA peak of the device: 8 independent dependent-chains per thread, 1024 threads x `blocks` CTAs, `iters` rounds of
// 8 FMAs.  2 * 8 * iters * threads flops.  Used by bench.py as the denominator of the KKT/IPM kernel's FLOP/s
// (MEASURED_PEAKS.json has no fp64 figure).  out[blockIdx*blockDim + tid] receives a value so nothing is optimised away.
__global__ void __launch_bounds__(1024) dfma_peak_kernel(int iters, double* out) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
extern "C" int myr_bench_dfma(int blocks, int iters, double* out, void* stream) {
  if (blocks < 1 || iters < 1 || !out) return fail(MYR_E_BADARG, "bad arguments to myr_bench_dfma%s", "", 0);
  dfma_peak_kernel<<<blocks, 1024, 0, (cudaStream_t)stream>>>(iters, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MYR_E_CUDA, "myr_bench_dfma: %s", cudaGetErrorString(e), 0);
  return MYR_OK;
}
