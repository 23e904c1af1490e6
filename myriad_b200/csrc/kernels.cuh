// Kernel templates and per-system launchers.  Compiled once per system (sys_unit.cu) so the build
// parallelises; api.cu dispatches through the per-system tables.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../../include/myriad_b200.h"
#include "engine.cuh"
#include "rollout.cuh"
#include "kernels_decl.h"

namespace myr {

int fail(int code, const char* fmt, const char* a = "", long long v = 0);  // api.cu
double system_default_T(int id);                                            // api.cu

template <class Sys>
static int make_problem(const MyrDesc* d, int B, Problem& P) {
  memset(&P, 0, sizeof(P));
  if (!d) return fail(MYR_E_BADARG, "null descriptor%s", "");
  if (d->intervals < 1) return fail(MYR_E_BADARG, "intervals must be >= 1 (got %s%lld)", "", d->intervals);
  if (B < 0) return fail(MYR_E_BADARG, "negative batch%s", "");
  P.B = B;
  P.N = d->intervals;
  P.cpi = d->optimizer == MYR_OPT_SHOOTING ? d->controls_per_interval : 1;
  if (P.cpi < 1) return fail(MYR_E_BADARG, "controls_per_interval must be >= 1%s", "");
  P.method = d->integration_method;
  if (P.method < 0 || P.method > 3) return fail(MYR_E_BADARG, "unknown integration_method %s%lld", "", P.method);
  P.terminal_cost = d->terminal_cost;
  P.T = d->T > 0 ? d->T : system_default_T(Sys::id);
  Sys::default_params(P.p);
  if (d->n_params > 0) {
    if (d->n_params != Sys::np) return fail(MYR_E_BADARG, "n_params does not match system %s (%lld given)", Sys::name, d->n_params);
    for (int i = 0; i < Sys::np; ++i) P.p[i] = d->params[i];
  }
  if (sys_is_node<Sys>::value) {
    const int k = d->node_num_hidden;
    if (k < 1 || k > kMaxMlpLayers - 1) return fail(MYR_E_BADARG, "node_num_hidden must be 1..4 (got %s%lld)", "", k);
    if (!d->theta) return fail(MYR_E_BADARG, "NODE system %s needs theta (MLP weights)", Sys::name);
    MlpDesc& M = P.mlp;
    M.L = k + 1;
    M.size[0] = Sys::n + Sys::m;
    M.hp = 8;
    for (int j = 0; j < k; ++j) {
      const int h = d->node_hidden[j];
      if (h < 1 || h > kMaxMlpWidth) return fail(MYR_E_BADARG, "hidden layer width must be 1..128 (got %s%lld)", "", h);
      M.size[j + 1] = h;
      if ((h + 7) / 8 * 8 > M.hp) M.hp = (h + 7) / 8 * 8;
    }
    M.size[k + 1] = Sys::n;
    long long off = 0;
    for (int j = 0; j <= k; ++j) {
      M.woff[j] = (int)off; off += (long long)M.size[j] * M.size[j + 1];
      M.boff[j] = (int)off; off += M.size[j + 1];
    }
    if (d->theta_doubles != off) return fail(MYR_E_BADARG, "theta_doubles does not match the layer sizes (%s%lld expected)", "", off);
    M.theta = d->theta;
  }
  return MYR_OK;
}

template <class Sys, class F>
static int dispatch_scheme(const MyrDesc* d, int B, F&& fn) {
  Problem P;
  int rc = make_problem<Sys>(d, B, P);
  if (rc) return rc;
  switch (d->optimizer) {
    case MYR_OPT_TRAPEZOIDAL: {
      using S = Trapezoid<Sys>;
      P.h = P.T / P.N; P.nvars = S::nvars(P); P.ncon = S::ncon(P);
      return fn(S{}, P);
    }
    case MYR_OPT_HERMITE_SIMPSON: {
      using S = HermiteSimpson<Sys>;
      P.h = P.T / P.N; P.nvars = S::nvars(P); P.ncon = S::ncon(P);
      return fn(S{}, P);
    }
    case MYR_OPT_SHOOTING:
      return fail(MYR_E_UNSUPPORTED, "this entry point has no SHOOTING block form (%s); use myr_eval / myr_ipm_solve", Sys::name);
    default:
      return fail(MYR_E_BADARG, "unknown optimizer %s%lld", "", d->optimizer);
  }
}

// shooting: lifted scheme selected by the number of control slots per step
template <class Sys, class F>
static int dispatch_shooting(const MyrDesc* d, int B, F&& fn) {
  Problem P;
  int rc = make_problem<Sys>(d, B, P);
  if (rc) return rc;
  if (d->optimizer != MYR_OPT_SHOOTING) return fail(MYR_E_BADARG, "not a shooting descriptor%s", "");
  P.h = P.T / (P.N * P.cpi);
  if (P.method == EULER) { using S = ShootingLifted<Sys, 1>; P.nvars = S::nvars(P); P.ncon = S::ncon(P); return fn(S{}, P); }
  if (P.method == RK4) { using S = ShootingLifted<Sys, 3>; P.nvars = S::nvars(P); P.ncon = S::ncon(P); return fn(S{}, P); }
  using S = ShootingLifted<Sys, 2>; P.nvars = S::nvars(P); P.ncon = S::ncon(P);
  return fn(S{}, P);
}

// any optimizer (entry points that work for all three transcriptions)
template <class Sys, class F>
static int dispatch_any(const MyrDesc* d, int B, F&& fn) {
  if (d->optimizer == MYR_OPT_SHOOTING) return dispatch_shooting<Sys>(d, B, fn);
  return dispatch_scheme<Sys>(d, B, fn);
}

static int cuda_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MYR_E_CUDA, "%s: CUDA error", (std::string(what) + ": " + cudaGetErrorString(e)).c_str());
  return MYR_OK;
}

static int threads_for(int Q, bool coop_mlp = false) {
  if (coop_mlp) return 256;  // 8 warps: one 8-row tile of a 64-wide layer each (node_mlp.cuh)
  int t = (Q + 31) / 32 * 32;
  if (t < 32) t = 32;
  if (t > 256) t = 256;
  return t;
}

template <class Sys>
int sys_problem_sizes(const MyrDesc* desc, MyrSizes* out) {
  return dispatch_any<Sys>(desc, 0, [&](auto s, const Problem& P) {
    using S = decltype(s);
    const Layout<S> L(P);
    out->n = S::n; out->m = S::m;
    out->nvars = P.nvars; out->ncon = P.ncon;
    out->nodes = L.Q; out->stages = L.St; out->nw = S::NW; out->nc = S::NC;
    {  // a workspace SLOT with every array in global memory (what the host twin uses; the device kernels place the
       // hot arrays in shared memory and use less): callers size the workspace with this upper bound
      int sm, gl;
      L.place(0, sm, gl);
      out->ipm_workspace_doubles = gl;
      out->ipm_workspace_slots = MYR_WS_MAX_SLOTS;
    }
    if (scheme_is_lifted<S>::value) {
      const int mc = P.method == RK4 ? 2 : 1;
      out->nx_nodes = P.N + 1; out->nu_nodes = mc * P.N * P.cpi + 1;
      out->stage_nodes = 1;
      // per interval: (n + 1) rows x (n + (mc*cpi+1) m) columns; row n is the gradient of the interval's cost
      out->jac_block_doubles = (int64_t)P.N * (S::n + 1) * (S::n + (mc * P.cpi + 1) * S::m);
      out->hess_block_doubles = 0;
    } else {
      out->nx_nodes = out->nu_nodes = L.Q;
      out->stage_nodes = S::kMaxStageNodes;
      out->jac_block_doubles = (int64_t)L.St * S::kMaxStageNodes * S::NC * S::NW;
      out->hess_block_doubles = (int64_t)L.Q * S::NWP;
    }
    return (int)MYR_OK;
  });
}

// Shooting K1 (reference-level): one thread per (instance, interval) rolls the interval out with forward-mode
// sensitivities.  Jblk[b][k] is (n+1) x ncol row-major: rows 0..n-1 = d px_k / d (xs[k], interval controls), row n =
// gradient of the interval's integrated cost; c = px - xs[k+1]; f and grad are accumulated with atomics.
template <class Sys, int NU>
MYR_HDI void shooting_eval_pair(const Problem& P, long long pair, const double* z_all, double* f, double* grad, double* c, double* J,
                                double* scratch_row) {
  using SI = ShootingInterval<Sys, NU>;
  constexpr int n = Sys::n, m = Sys::m;
  const int b = (int)(pair / P.N), k = (int)(pair % P.N);
  const int M = SI::mc * P.cpi, ncol = n + (M + 1) * m;
  const double* z = z_all + (long long)b * P.nvars;
  double px[n], cst;
  double* S = J ? J + ((long long)b * P.N + k) * (n + 1) * ncol : scratch_row;
  if (S) SI::template run<true>(P, k, z, px, cst, S, ncol);
  else SI::template run<false>(P, k, z, px, cst, nullptr, 0);
  if (c) {
#pragma unroll
    for (int i = 0; i < n; ++i) c[(long long)b * P.ncon + k * n + i] = px[i] - z[(k + 1) * n + i];
  }
#ifdef __CUDA_ARCH__
  if (f) atomicAdd(f + b, cst);
#else
  if (f) f[b] += cst;
#endif
  if (grad && S) {
    double* g = grad + (long long)b * P.nvars;
    const int ubase = (P.N + 1) * n;
    for (int col = 0; col < ncol; ++col) {
      const int dst = col < n ? k * n + col : ubase + k * M * m + (col - n);
#ifdef __CUDA_ARCH__
      atomicAdd(g + dst, S[n * ncol + col]);
#else
      g[dst] += S[n * ncol + col];
#endif
    }
  }
}

template <class Sys, int NU>
__global__ void shooting_eval_kernel(Problem P, const double* z, double* f, double* grad, double* c, double* J) {
  const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pair < (long long)P.B * P.N) shooting_eval_pair<Sys, NU>(P, pair, z, f, grad, c, J, nullptr);
}

// views of the reference's flat layouts for the cooperative MLP pass
template <class S>
struct RefZView {
  const Problem* P; const double* z;
  MYR_HDI double operator()(int q, int i) const { return z[S::zidx(*P, q, i)]; }
};
template <class S>
struct RefLamView {
  const Problem* P; const double* lam;
  MYR_HDI double operator()(int j, int r) const { return lam ? lam[S::cidx(*P, j, r)] : 0.0; }
};

// ------------------------------------------------------------------ K1 kernel
// One CTA per instance, one thread per node.  phi/psi meet in shared memory to form the stage constraints;
// each node writes its Jacobian blocks straight into the compact stage-row layout.
template <class S, bool kHost>
__host__ __device__ inline void eval_instance(const Problem& P, int b, const double* z_all, const double* lam_all, double* f_out,
                                              double* grad_out, double* c_out, double* J_out, double* H_out,
                                              double* sphi, double* spsi, double* red) {
  const Layout<S> L(P);
  const double* z = z_all + (long long)b * P.nvars;
  const double* lam = lam_all ? lam_all + (long long)b * P.ncon : nullptr;
  const bool want_h = lam && H_out;
  const long long jstride = (long long)L.St * S::kMaxStageNodes * S::NC * S::NW;
  double fsum = 0.0;
  for (int q = MYR_TID; q < L.Q; q += MYR_NT) {
    double v[S::NW], lp[S::NC], ls[S::NC];
#pragma unroll
    for (int i = 0; i < S::NW; ++i) v[i] = z[S::zidx(P, q, i)];
    const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
#pragma unroll
    for (int r = 0; r < S::NC; ++r) {
      lp[r] = (want_h && jp >= 0) ? lam[S::cidx(P, jp, r)] : 0.0;
      ls[r] = (want_h && js >= 0) ? lam[S::cidx(P, js, r)] : 0.0;
    }
    double ell, gl[S::NW], phi[S::NC], psi[S::NC], G[S::NC * S::NW], F[S::NC * S::NW], W[S::NWP];
    if (want_h) S::template eval_node<2>(P, q, v, lp, ls, ell, gl, phi, psi, G, F, W);
    else S::template eval_node<1>(P, q, v, lp, ls, ell, gl, phi, psi, G, F, W);
    fsum += ell;
#pragma unroll
    for (int r = 0; r < S::NC; ++r) { sphi[q * S::NC + r] = phi[r]; spsi[q * S::NC + r] = psi[r]; }
    if (grad_out) {
#pragma unroll
      for (int i = 0; i < S::NW; ++i) grad_out[(long long)b * P.nvars + S::zidx(P, q, i)] = gl[i];
    }
    if (J_out) {
      double* Jb = J_out + (long long)b * jstride;
      if (jp >= 0) {
        double* dst = Jb + ((long long)jp * S::kMaxStageNodes + S::phi_slot(P, q)) * S::NC * S::NW;
#pragma unroll
        for (int i = 0; i < S::NC * S::NW; ++i) dst[i] = G[i];
      }
      if (js >= 0) {
        double* dst = Jb + ((long long)js * S::kMaxStageNodes + S::psi_slot(P, q)) * S::NC * S::NW;
#pragma unroll
        for (int i = 0; i < S::NC * S::NW; ++i) dst[i] = F[i];
      }
    }
    if (want_h) {
      double* dst = H_out + ((long long)b * L.Q + q) * S::NWP;
#pragma unroll
      for (int i = 0; i < S::NWP; ++i) dst[i] = W[i];
    }
  }
  const double f = block_sum(fsum, red);
  MYR_SYNC();
  if (f_out && MYR_TID == 0) f_out[b] = f;
  if (c_out) {
    for (int j = MYR_TID; j < L.St; j += MYR_NT) {
      const int nk = S::stage_nodes(P, j);
#pragma unroll
      for (int r = 0; r < S::NC; ++r) {
        double a = 0.0;
        for (int k = 0; k < nk; ++k) {
          int role; const int q = S::stage_node(P, j, k, role);
          a += role ? spsi[q * S::NC + r] : sphi[q * S::NC + r];
        }
        c_out[(long long)b * P.ncon + S::cidx(P, j, r)] = a;
      }
    }
  }
}

// Device K1: one CTA per instance, one thread per node.  Node blocks of the Jacobian are staged in shared memory
// (rows padded by one double against bank conflicts) and written out with fully coalesced stores -- the Jacobian is
// 74% of the kernel's HBM traffic; z (read) and grad (write) are accessed as 32-byte-per-thread runs directly.
// Launch shape of K1: the kernel is bound by the latency of its own loads and store drain, so resident CTAs are what
// matters -- two-node-per-stage schemes run 128 threads with registers capped for MYR_K1_MINBLOCKS CTAs per SM.
#ifndef MYR_K1_MINBLOCKS
#define MYR_K1_MINBLOCKS 5
#endif
template <class S>
struct EvalLaunch {
  static constexpr bool kWide = Layout<S>::kCoopMlp || S::kMaxStageNodes >= 3;
  static constexpr int kThreads = kWide ? 256 : 128;
  static constexpr int kMinBlocks = kWide ? 1 : MYR_K1_MINBLOCKS;
  static int threads(int Q) { const int t = threads_for(Q, Layout<S>::kCoopMlp); return t < kThreads ? t : kThreads; }
};

template <class S>
__global__ void __launch_bounds__(EvalLaunch<S>::kThreads, EvalLaunch<S>::kMinBlocks) eval_kernel(Problem P, const double* __restrict__ z_all, const double* __restrict__ lam_all,
                                                   double* __restrict__ f_out, double* __restrict__ grad_out, double* __restrict__ c_out,
                                                   double* __restrict__ J_out, double* __restrict__ H_out) {
  extern __shared__ __align__(16) double smem[];
  constexpr int NW = S::NW, NC = S::NC, BLK = NC * NW, ROW = S::kMaxStageNodes * BLK, ROWP = ROW + 2;
  static_assert(ROW % 2 == 0, "stage rows must be a multiple of 16 bytes for the bulk stores");
  const Layout<S> L(P);
  const int Q = L.Q, St = L.St;
  double* red = smem;
  double* sphi = smem + 64;
  double* spsi = sphi + Q * NC;
  double* sJ = spsi + Q * NC;             // St * ROWP
  double* sH = sJ + (J_out ? (size_t)St * ROWP : 0);  // Q * NWP (only when the Hessian is requested); the Jacobian rows are
                                                       // staged only when asked for -- same rule as the launch's size
  const bool want_h = lam_all && H_out;
  // NODE systems: per-node MLP outputs + the tensor-core pass's scratch live behind the other regions
  double* sDyn = sH + ((lam_all && H_out) ? (size_t)Q * S::NWP : 0);
  double* sScr = sDyn + (size_t)Q * (S::n + S::n * NW + S::NWP);
  const long long jstride = (long long)St * ROW;
  for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
    const double* z = z_all + (long long)b * P.nvars;
    const double* lam = lam_all ? lam_all + (long long)b * P.ncon : nullptr;
    double fsum = 0.0;
    if constexpr (Layout<S>::kCoopMlp) {
      double* df = sDyn; double* dJ = df + Q * S::n; double* dH = dJ + Q * S::n * NW;
      const RefZView<S> zv{&P, z};
      const RefLamView<S> lv{&P, lam};
      if (want_h) mlp_nodes_pass<S, 2>(P, Q, zv, lv, df, dJ, dH, sScr);
      else mlp_nodes_pass<S, 1>(P, Q, zv, lv, df, dJ, dH, sScr);
    }
    for (int q = threadIdx.x; q < Q; q += blockDim.x) {
      double v[NW], lp[NC], ls[NC];
#pragma unroll
      for (int i = 0; i < NW; ++i) v[i] = z[S::zidx(P, q, i)];
      const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        lp[r] = (want_h && jp >= 0) ? lam[S::cidx(P, jp, r)] : 0.0;
        ls[r] = (want_h && js >= 0) ? lam[S::cidx(P, js, r)] : 0.0;
      }
      PreDyn pre;
      if constexpr (Layout<S>::kCoopMlp) { pre.f = sDyn + q * S::n; pre.J = sDyn + Q * S::n + q * S::n * NW; pre.H = sDyn + Q * (S::n + S::n * NW) + q * S::NWP; }
      double ell, gl[NW], phi[NC], psi[NC], G[BLK], F[BLK], W[S::NWP];
      if (want_h) S::template eval_node<2>(P, q, v, lp, ls, ell, gl, phi, psi, G, F, W, pre);
      else S::template eval_node<1>(P, q, v, lp, ls, ell, gl, phi, psi, G, F, W, pre);
      fsum += ell;
#pragma unroll
      for (int r = 0; r < NC; ++r) { sphi[q * NC + r] = phi[r]; spsi[q * NC + r] = psi[r]; }
      if (grad_out) {
#pragma unroll
        for (int i = 0; i < NW; ++i) grad_out[(long long)b * P.nvars + S::zidx(P, q, i)] = gl[i];
      }
      if (J_out) {
        if (jp >= 0) {
          double* dst = sJ + (size_t)jp * ROWP + S::phi_slot(P, q) * BLK;
#pragma unroll
          for (int i = 0; i < BLK; ++i) dst[i] = G[i];
        }
        if (js >= 0) {
          double* dst = sJ + (size_t)js * ROWP + S::psi_slot(P, q) * BLK;
#pragma unroll
          for (int i = 0; i < BLK; ++i) dst[i] = F[i];
        }
      }
      if (want_h) {
#pragma unroll
        for (int i = 0; i < S::NWP; ++i) sH[q * S::NWP + i] = W[i];
      }
    }
    const double f = block_sum(fsum, red);  // contains the barriers that publish sphi/spsi/sJ/sH
    __syncthreads();
    if (f_out && threadIdx.x == 0) f_out[b] = f;
    if (c_out) {
      for (int e = threadIdx.x; e < St * NC; e += blockDim.x) {
        const int j = e / NC, r = e - j * NC;
        const int nk = S::stage_nodes(P, j);
        double a = 0.0;
        for (int k = 0; k < nk; ++k) {
          int role; const int q = S::stage_node(P, j, k, role);
          a += role ? spsi[q * NC + r] : sphi[q * NC + r];
        }
        c_out[(long long)b * P.ncon + S::cidx(P, j, r)] = a;
      }
    }
    if (J_out) {
      // one TMA bulk store (cp.async.bulk shared -> global, UBLKCP) per stage row: the 32 KB Jacobian of the instance
      // leaves the SM without occupying load/store issue slots; the stores drain while c / H are written below
      double* Jb = J_out + (long long)b * jstride;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      for (int j = threadIdx.x; j < St; j += blockDim.x) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(sJ + (size_t)j * ROWP);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(Jb + (size_t)j * ROW), "r"(src), "r"((unsigned)(ROW * sizeof(double))) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (want_h) {
      double* Hb = H_out + (long long)b * Q * S::NWP;
      for (int o = threadIdx.x; o < Q * S::NWP; o += blockDim.x) Hb[o] = sH[o];
    }
    if (J_out) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory of the rows has been read
    __syncthreads();
  }
}

template <class Sys>
int sys_eval(const MyrDesc* desc, int B, const double* z, const double* lam, double* f, double* grad, double* c,
                        double* Jblk, double* Hblk, void* stream) {
  if (desc->optimizer == MYR_OPT_SHOOTING) {
    return dispatch_shooting<Sys>(desc, B, [&](auto s, const Problem& P) {
      using S = decltype(s);
      constexpr int NU = (S::NW - S::n) / S::m;
      if (B == 0) return (int)MYR_OK;
      if (!z) return fail(MYR_E_BADARG, "z is null%s", "");
      if (Hblk) return fail(MYR_E_UNSUPPORTED, "no reference-level Hessian blocks for SHOOTING (%s)", Sys::name);
      if (grad && !Jblk) return fail(MYR_E_BADARG, "SHOOTING grad needs Jblk (the cost row lives there)%s", "");
      cudaStream_t st = (cudaStream_t)stream;
      if (f) cudaMemsetAsync(f, 0, sizeof(double) * B, st);
      if (grad) cudaMemsetAsync(grad, 0, sizeof(double) * (size_t)B * P.nvars, st);
      const long long pairs = (long long)B * P.N;
      shooting_eval_kernel<Sys, NU><<<(unsigned)((pairs + 127) / 128), 128, 0, st>>>(P, z, f, grad, c, Jblk);
      return cuda_check("myr_eval(shooting)");
    });
  }
  return dispatch_scheme<Sys>(desc, B, [&](auto s, const Problem& P) {
    using S = decltype(s);
    if (B == 0) return (int)MYR_OK;
    if (!z) return fail(MYR_E_BADARG, "z is null%s", "");
    const Layout<S> L(P);
    const size_t row = (size_t)S::kMaxStageNodes * S::NC * S::NW + 2;
    size_t smd = 64 + 2 * (size_t)L.Q * S::NC + (Jblk ? (size_t)L.St * row : 0) + ((lam && Hblk) ? (size_t)L.Q * S::NWP : 0);
    if (Layout<S>::kCoopMlp) smd += (size_t)L.Q * (S::n + S::n * S::NW + S::NWP) + mlp_scratch_doubles<S>(P.mlp);
    const size_t sm = smd * sizeof(double);
    if (sm > 227 * 1024) return fail(MYR_E_UNSUPPORTED, "problem too large for the shared-memory staged K1 (%s%lld bytes)", "", (long long)sm);
    if (sm > 48 * 1024 && cudaFuncSetAttribute(eval_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
      return fail(MYR_E_CUDA, "myr_eval: cannot reserve %s%lld bytes of shared memory", "", (long long)sm);
    eval_kernel<S><<<B, EvalLaunch<S>::threads(L.Q), sm, (cudaStream_t)stream>>>(P, z, lam, f, grad, c, Jblk, Hblk);
    return cuda_check("myr_eval");
  });
}

template <class Sys>
int sys_host_eval(const MyrDesc* desc, int B, const double* z, const double* lam, double* f, double* grad, double* c,
                             double* Jblk, double* Hblk) {
  if (desc->optimizer == MYR_OPT_SHOOTING) {
    return dispatch_shooting<Sys>(desc, B, [&](auto s, const Problem& P) {
      using S = decltype(s);
      constexpr int NU = (S::NW - S::n) / S::m;
      if (Hblk) return fail(MYR_E_UNSUPPORTED, "no reference-level Hessian blocks for SHOOTING (%s)", Sys::name);
      if (grad && !Jblk) return fail(MYR_E_BADARG, "SHOOTING grad needs Jblk (the cost row lives there)%s", "");
      if (f) for (int b = 0; b < B; ++b) f[b] = 0.0;
      if (grad) for (long long i = 0; i < (long long)B * P.nvars; ++i) grad[i] = 0.0;
      for (long long pair = 0; pair < (long long)B * P.N; ++pair) shooting_eval_pair<Sys, NU>(P, pair, z, f, grad, c, Jblk, nullptr);
      return (int)MYR_OK;
    });
  }
  return dispatch_scheme<Sys>(desc, B, [&](auto s, const Problem& P) {
    using S = decltype(s);
    const int Q = S::num_nodes(P);
    std::vector<double> sh(64 + 2 * (size_t)Q * S::NC);
    for (int b = 0; b < B; ++b)
      eval_instance<S, true>(P, b, z, lam, f, grad, c, Jblk, Hblk, sh.data() + 64, sh.data() + 64 + Q * S::NC, sh.data());
    return (int)MYR_OK;
  });
}

// ------------------------------------------------------------------ workspace slots and shared-memory plan
// Launch shape of the per-instance kernels: persistent CTAs, one instance at a time each; 128 threads stride over
// nodes / stages (256 for the cooperative tensor-core MLP pass of NODE systems and for the 3-node stages of
// Hermite-Simpson).  kMinBlocks bounds registers so that several instances are resident per SM: the kernel is
// latency-bound, resident warps are what hides it.
#ifndef MYR_IPM_MINBLOCKS
#define MYR_IPM_MINBLOCKS 2
#endif
#ifndef MYR_IPM_THREADS
#define MYR_IPM_THREADS 128
#endif
template <class S>
struct IpmLaunch {
  static constexpr bool kWide = Layout<S>::kCoopMlp || S::kMaxStageNodes >= 3;
  static constexpr int kThreads = kWide ? 256 : MYR_IPM_THREADS;
  static constexpr int kMinBlocks = kWide ? 1 : MYR_IPM_MINBLOCKS;
  // one thread per variable (the vector phases are flat loops over the Q * NW variables), at least one per node
  static int threads(int Q) {
    if (Layout<S>::kCoopMlp) return kThreads;
    int t = (Q * S::NW + 31) / 32 * 32;
    const int tq = (Q + 31) / 32 * 32;
    if (t < tq) t = tq;
    return t < kThreads ? t : kThreads;
  }
};

// Which arrays of the slot live in shared memory (engine.cuh, "Memory plan") and how many CTAs per SM that allows.
// sm_100: 228 KB of shared memory per SM, 227 KB per CTA at most, 1 KB reserved per resident CTA.
template <class S>
struct SlotPlan {
  unsigned long long mask;
  int smem_doubles, glob_doubles, ctas_per_sm, theta_doubles;
  size_t fixed_bytes, smem_bytes;
  explicit SlotPlan(const Problem& P, int max_ctas) {
    const Layout<S> L(P);
    // NODE systems: the cooperative MLP pass's scratch AND a copy of the weights (70 KB for 3 x 64): every node group
    // re-reads every weight, and with shared memory carved out to the maximum the L1 that used to hold them is gone
    size_t mlp = 0;
    theta_doubles = 0;
    if (Layout<S>::kCoopMlp) {
      const MlpDesc& M = P.mlp;
      theta_doubles = (M.boff[M.L - 1] + M.size[M.L] + 1) & ~1;
      mlp = (size_t)(mlp_scratch_doubles<S>(P.mlp) + theta_doubles) * sizeof(double);
    }
    fixed_bytes = kFixedScratch * sizeof(double) + mlp;
    auto budget = [&](int k) -> long long {
      long long per = 233472 / k - 1024;
      if (per > 232448) per = 232448;
      if (const char* e = getenv("MYR_IPM_SMEM_KB")) { const long long cap = atoll(e) * 1024; if (cap < per) per = cap; }
      per -= (long long)fixed_bytes;
      return per > 0 ? per / (long long)sizeof(double) : 0;
    };
    int k = max_ctas < 1 ? 1 : max_ctas;
    if (const char* e = getenv("MYR_IPM_CTAS")) { if (atoi(e) >= 1) k = atoi(e); }   // tuning knob (debug)
    else {
      // the cyclic-reduction scratch is what shared memory is for: give up resident CTAs until it fits
      int kk = k;
      while (kk > 1 && budget(kk) < L.cr_doubles()) --kk;
      if (budget(kk) >= L.cr_doubles()) k = kk;
    }
    ctas_per_sm = k;
    mask = L.place(budget(k), smem_doubles, glob_doubles);
    smem_bytes = fixed_bytes + (size_t)smem_doubles * sizeof(double);
  }
};

static int device_sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < 1) sms = 1;
  }
  return sms;
}

// grid of a persistent per-instance kernel: resident CTAs of the device, at most one per instance / workspace slot
template <class K>
static int persistent_grid(K kernel, int threads, size_t smem, int B, int* err) {
  *err = 0;
  if (smem > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { *err = 1; return 0; }
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) { *err = 1; return 0; }
  long long g = (long long)occ * device_sm_count();
  if (g > MYR_WS_MAX_SLOTS) g = MYR_WS_MAX_SLOTS;
  if (g > B) g = B;
  return (int)g;
}

// ------------------------------------------------------------------ K2 kernel
template <class S>
__host__ __device__ inline void kkt_instance(const Problem& P, int b, const double* Hblk, const double* Jblk, const double* sigma,
                                             const double* rhs_z, const double* rhs_c, double dw, double dc, double* dz, double* dlam,
                                             int32_t* inertia_ok, const WS<S>& ws) {
  using D = Dims<S>;
  constexpr int NW = S::NW, NC = S::NC;
  const long long jstride = (long long)ws.St * S::kMaxStageNodes * NC * NW;
  const double* Jb = Jblk + (long long)b * jstride;
  for (int q = MYR_TID; q < ws.Q; q += MYR_NT) {
    const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
    double* Wq = ws.W + q * D::WSZ;
    if constexpr (D::kAff) {
      // the slot keeps the dynamics Jacobian J: recover it from the caller's role block  a J + b [I 0]  (Jblk as myr_eval
      // returns it; row group with the largest |a| of whichever role the node has)
      constexpr int n = S::n, NG = S::NG;
      double ap[NG], bp[NG], as[NG], bs[NG];
      S::role_coefs(P, q, ap, bp, as, bs);
      const bool use_phi = jp >= 0;
      const double* al = use_phi ? ap : as; const double* be = use_phi ? bp : bs;
      int g = 0;
      for (int k = 1; k < NG; ++k) if (fabs(al[k]) > fabs(al[g])) g = k;
      const double* blk = Jb + ((long long)(use_phi ? jp : js) * S::kMaxStageNodes + (use_phi ? S::phi_slot(P, q) : S::psi_slot(P, q))) * NC * NW;
      double* Jq = ws.G + q * D::GS;
      for (int r = 0; r < n; ++r)
        for (int i = 0; i < NW; ++i) Jq[r * NW + i] = (blk[(g * n + r) * NW + i] - (i == r ? be[g] : 0.0)) / al[g];
    } else {
      double* Gq = ws.G + q * D::GS; double* Fq = ws.F + q * D::GS;
      for (int i = 0; i < NC * NW; ++i) {
        Gq[i] = jp >= 0 ? Jb[((long long)jp * S::kMaxStageNodes + S::phi_slot(P, q)) * NC * NW + i] : 0.0;
        Fq[i] = js >= 0 ? Jb[((long long)js * S::kMaxStageNodes + S::psi_slot(P, q)) * NC * NW + i] : 0.0;
      }
    }
    for (int i = 0; i < S::NWP; ++i) Wq[i] = Hblk[((long long)b * ws.Q + q) * S::NWP + i];
    uint32_t fm = 0;
    for (int i = 0; i < NW; ++i) {
      const int id = S::zidx(P, q, i);
      const double sg = sigma[(long long)b * P.nvars + id];
      const bool fx = isinf(sg);
      if (fx) fm |= 1u << i;
      NQ(sig, q, i) = fx ? 0.0 : sg;
      NQ(rb, q, i) = fx ? 0.0 : rhs_z[(long long)b * P.nvars + id];
    }
    ws.fix()[q] = fm;
  }
  for (int j = MYR_TID; j < ws.St; j += MYR_NT)
    for (int r = 0; r < NC; ++r) NS(c, j, r) = rhs_c[(long long)b * P.ncon + S::cidx(P, j, r)];
  MYR_SYNC();
  int parity = 0;
  const bool ok = kkt_solve<S>(P, ws, dw, dc, 0.0, 0, parity);
  for (int q = MYR_TID; q < ws.Q; q += MYR_NT)
    for (int i = 0; i < NW; ++i) dz[(long long)b * P.nvars + S::zidx(P, q, i)] = ok ? NQ(dz, q, i) : 0.0;
  for (int j = MYR_TID; j < ws.St; j += MYR_NT)
    for (int r = 0; r < NC; ++r) dlam[(long long)b * P.ncon + S::cidx(P, j, r)] = NS(dlam, j, r);
  if (MYR_TID == 0 && inertia_ok) inertia_ok[b] = ok ? 1 : 0;
  MYR_SYNC();
}

template <class S>
__global__ void __launch_bounds__(256) kkt_kernel(Problem P, const double* Hblk, const double* Jblk, const double* sigma, const double* rhs_z,
                                                  const double* rhs_c, double dw, double dc, double* dz, double* dlam, int32_t* inertia_ok,
                                                  unsigned long long mask, double* work, long long slot_stride) {
  extern __shared__ __align__(16) double smem[];
  const Layout<S> L(P);
  WS<S> ws(L, mask, smem + kFixedScratch, work + (long long)blockIdx.x * slot_stride);
  ws.red = smem;
  for (int b = blockIdx.x; b < P.B; b += gridDim.x)
    kkt_instance<S>(P, b, Hblk, Jblk, sigma, rhs_z, rhs_c, dw, dc, dz, dlam, inertia_ok, ws);
}

template <class Sys>
int sys_kkt_solve(const MyrDesc* desc, int B, const double* Hblk, const double* Jblk, const double* sigma, const double* rhs_z,
                             const double* rhs_c, double delta_w, double delta_c, double* dz, double* dlam, int32_t* inertia_ok, double* ws,
                             size_t ws_doubles, void* stream) {
  return dispatch_scheme<Sys>(desc, B, [&](auto s, const Problem& P) {
    using S = decltype(s);
    if (B == 0) return (int)MYR_OK;
    const Layout<S> L(P);
    SlotPlan<S> sp(P, 1);
    sp.smem_bytes -= sp.fixed_bytes - kFixedScratch * sizeof(double);   // no MLP scratch in this kernel
    const int threads = threads_for(L.Q);
    int err;
    const int grid = persistent_grid(kkt_kernel<S>, threads, sp.smem_bytes, B, &err);
    if (err) return fail(MYR_E_CUDA, "myr_kkt_solve: cannot configure the kernel (%s%lld bytes of shared memory)", "", (long long)sp.smem_bytes);
    const size_t need = MYR_WS_HEADER + (size_t)grid * sp.glob_doubles;
    if (ws_doubles < need) return fail(MYR_E_WORKSPACE, "workspace too small: need %s%lld doubles", "", (long long)need);
    kkt_kernel<S><<<grid, threads, sp.smem_bytes, (cudaStream_t)stream>>>(P, Hblk, Jblk, sigma, rhs_z, rhs_c, delta_w, delta_c, dz, dlam,
                                                                          inertia_ok, sp.mask, ws + MYR_WS_HEADER, sp.glob_doubles);
    return cuda_check("myr_kkt_solve");
  });
}

// host twins: one slot (all arrays in "global" memory) per OpenMP thread, instances shared dynamically
template <class S>
static int host_slots(const Problem& P, int B, size_t ws_doubles, int& stride) {
  const Layout<S> L(P);
  int sm;
  L.place(0, sm, stride);
  int nt = 1;
#ifdef _OPENMP
  nt = omp_get_max_threads();
#endif
  if (nt > B) nt = B;
  if (nt > MYR_WS_MAX_SLOTS) nt = MYR_WS_MAX_SLOTS;
  if (nt < 1) nt = 1;
  while (nt > 1 && MYR_WS_HEADER + (size_t)nt * stride > ws_doubles) --nt;
  if (MYR_WS_HEADER + (size_t)nt * stride > ws_doubles) return 0;
  return nt;
}

template <class Sys>
int sys_host_kkt_solve(const MyrDesc* desc, int B, const double* Hblk, const double* Jblk, const double* sigma,
                                  const double* rhs_z, const double* rhs_c, double delta_w, double delta_c, double* dz, double* dlam,
                                  int32_t* inertia_ok, double* ws, size_t ws_doubles) {
  return dispatch_scheme<Sys>(desc, B, [&](auto s, const Problem& P) {
    using S = decltype(s);
    if (B == 0) return (int)MYR_OK;
    const Layout<S> L(P);
    int stride;
    const int nt = host_slots<S>(P, B, ws_doubles, stride);
    if (!nt) return fail(MYR_E_WORKSPACE, "workspace too small: need %s%lld doubles", "", (long long)(MYR_WS_HEADER + stride));
#pragma omp parallel for schedule(dynamic) num_threads(nt)
    for (int b = 0; b < B; ++b) {
      int t = 0;
#ifdef _OPENMP
      t = omp_get_thread_num();
#endif
      WS<S> w(L, 0ull, nullptr, ws + MYR_WS_HEADER + (size_t)t * stride);
      kkt_instance<S>(P, b, Hblk, Jblk, sigma, rhs_z, rhs_c, delta_w, delta_c, dz, dlam, inertia_ok, w);
    }
    return (int)MYR_OK;
  });
}

// ------------------------------------------------------------------ K3 kernel
inline IpmOpts make_opts(const MyrIpmOpts* o) {
  IpmOpts r;
  memset(&r, 0, sizeof(r));
  r.max_iter = (o && o->max_iter > 0) ? o->max_iter : 1000;
  r.max_ls = (o && o->max_ls > 0) ? o->max_ls : 40;
  r.acceptable_iter = (o && o->acceptable_iter > 0) ? o->acceptable_iter : 15;
  r.tol = (o && o->tol > 0) ? o->tol : 1e-8;
  r.acceptable_tol = (o && o->acceptable_tol > 0) ? o->acceptable_tol : 1e-6;
  r.mu_init = (o && o->mu_init > 0) ? o->mu_init : 0.1;
  r.mu_min = 1e-11; r.kappa_eps = 10.0; r.kappa_mu = 0.2; r.theta_mu = 1.5; r.tau_min = 0.99;
  r.bound_push = 1e-2; r.bound_frac = 1e-2; r.bound_relax = 1e-8; r.kappa_sigma = 1e10; r.s_max = 100.0;
  r.delta_min = 1e-20; r.delta_0 = 1e-4; r.delta_max = 1e40; r.delta_c = 0.0;
  r.kappa_w_minus = 1.0 / 3.0; r.kappa_w_plus = 8.0; r.kappa_w_plus_first = 100.0;
  r.eta = 1e-4; r.rho = 0.1;
  r.use_filter = 1;
  if (const char* e = getenv("MYR_FILTER")) r.use_filter = atoi(e) != 0;   // A/B knob: 0 = l1-merit acceptance only
  r.use_watchdog = 1;
  if (const char* e = getenv("MYR_WATCHDOG")) r.use_watchdog = atoi(e) != 0;
  r.delta_reg = 1e-8; r.max_refine = 1;
  r.max_soc = (o && o->max_soc != 0) ? (o->max_soc > 0 ? o->max_soc : 0) : 4;
  if (const char* e = getenv("MYR_MU_INIT")) r.mu_init = atof(e);          // tuning knobs (debug)
  if (const char* e = getenv("MYR_KAPPA_MU")) r.kappa_mu = atof(e);
  if (const char* e = getenv("MYR_THETA_MU")) r.theta_mu = atof(e);
  if (const char* e = getenv("MYR_KAPPA_EPS")) r.kappa_eps = atof(e);
  if (const char* e = getenv("MYR_TAU_MIN")) r.tau_min = atof(e);
  if (const char* e = getenv("MYR_DELTA_REG")) r.delta_reg = atof(e);
  if (const char* e = getenv("MYR_MAX_REFINE")) r.max_refine = atoi(e);
  if (const char* e = getenv("MYR_MAX_SOC")) r.max_soc = atoi(e) > 0 ? atoi(e) : 0;
  return r;
}

// Persistent CTAs: each pulls the next instance from a counter (instances differ in iteration count, so a static
// assignment would leave SMs idle at the tail) and solves it in its own workspace slot.
template <class S>
__global__ void __launch_bounds__(IpmLaunch<S>::kThreads, IpmLaunch<S>::kMinBlocks)
ipm_kernel(Problem P, IpmOpts O, IpmIO io, unsigned long long mask, int smem_doubles, int theta_doubles, double* work, long long slot_stride,
           int* counter) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_next;
  const Layout<S> L(P);
  WS<S> ws(L, mask, smem + kFixedScratch, work + (long long)blockIdx.x * slot_stride);
  ws.red = smem;
  ws.filt = smem + 2 * kRedStride;
  ws.mlp_scr = smem + kFixedScratch + smem_doubles + theta_doubles;
  {
    using LY = Layout<S>;
    const unsigned long long cr = (1ull << LY::kCrArrays) - 1ull;
    const unsigned long long nodem = (1ull << LY::A_Hinv) | (1ull << LY::A_G) | (1ull << LY::A_F);
    ws.sh = (mask & cr) == cr ? (((mask & nodem) == nodem) ? 2 : 1) : 0;
  }
  if (Layout<S>::kCoopMlp) {   // weights into shared memory once per (persistent) CTA
    double* th = smem + kFixedScratch + smem_doubles;
    for (int e = threadIdx.x; e < theta_doubles; e += blockDim.x) th[e] = e < P.mlp.boff[P.mlp.L - 1] + P.mlp.size[P.mlp.L] ? P.mlp.theta[e] : 0.0;
    ws.theta = th;
    __syncthreads();
  }
  while (true) {
    if (threadIdx.x == 0) s_next = atomicAdd(counter, 1);
    __syncthreads();
    const int b = s_next;
    if (b >= P.B) break;
    ipm_solve_entry<S>(P, O, io, b, ws);   // ends with a barrier: s_next is not overwritten before every thread has read it
  }
}

template <class Sys>
int sys_ipm_solve(const MyrDesc* desc, const MyrIpmOpts* opts, int B, const double* z0, const double* lb, const double* ub, double* z,
                             double* lam, double* zL, double* zU, double* obj, double* kkt_err, double* con_inf, int32_t* status, int32_t* iters,
                             double* ws, size_t ws_doubles, void* stream) {
  return dispatch_any<Sys>(desc, B, [&](auto s, const Problem& P) {
    using S = decltype(s);
    if (B == 0) return (int)MYR_OK;
    if (!z0 || !lb || !ub || !z || !lam || !zL || !zU || !obj || !kkt_err || !con_inf || !status || !iters || !ws)
      return fail(MYR_E_BADARG, "null buffer passed to myr_ipm_solve%s", "");
    const Layout<S> L(P);
    const SlotPlan<S> sp(P, IpmLaunch<S>::kMinBlocks);
    const int threads = IpmLaunch<S>::threads(L.Q);
    int err;
    const int grid = persistent_grid(ipm_kernel<S>, threads, sp.smem_bytes, B, &err);
    if (err) return fail(MYR_E_CUDA, "myr_ipm_solve: cannot configure the kernel (%s%lld bytes of shared memory)", "", (long long)sp.smem_bytes);
    const size_t need = MYR_WS_HEADER + (size_t)grid * sp.glob_doubles;
    if (ws_doubles < need) return fail(MYR_E_WORKSPACE, "workspace too small: need %s%lld doubles", "", (long long)need);
    IpmIO io{z0, lb, ub, z, lam, zL, zU, obj, kkt_err, con_inf, status, iters};
    const IpmOpts O = make_opts(opts);
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(ws, 0, MYR_WS_HEADER * sizeof(double), st) != cudaSuccess) return cuda_check("myr_ipm_solve(memset)");
    ipm_kernel<S><<<grid, threads, sp.smem_bytes, st>>>(P, O, io, sp.mask, sp.smem_doubles, sp.theta_doubles, ws + MYR_WS_HEADER,
                                                         sp.glob_doubles, reinterpret_cast<int*>(ws));
    return cuda_check("myr_ipm_solve");
  });
}

#ifdef MYR_PROFILE_PHASES
extern "C" int myr_debug_phase_cycles(double* out16, int reset) {
  unsigned long long h[16];
  cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof(h));
  for (int i = 0; i < 16; ++i) out16[i] = (double)h[i];
  if (reset) { memset(h, 0, sizeof(h)); cudaMemcpyToSymbol(g_phase_cycles, h, sizeof(h)); }
  return 0;
}
#endif

template <class Sys>
int sys_host_ipm_solve(const MyrDesc* desc, const MyrIpmOpts* opts, int B, const double* z0, const double* lb, const double* ub,
                                  double* z, double* lam, double* zL, double* zU, double* obj, double* kkt_err, double* con_inf,
                                  int32_t* status, int32_t* iters, double* ws, size_t ws_doubles) {
  return dispatch_any<Sys>(desc, B, [&](auto s, const Problem& P) {
    using S = decltype(s);
    if (B == 0) return (int)MYR_OK;
    const Layout<S> L(P);
    int stride;
    const int nt = host_slots<S>(P, B, ws_doubles, stride);
    if (!nt) return fail(MYR_E_WORKSPACE, "workspace too small: need %s%lld doubles", "", (long long)(MYR_WS_HEADER + stride));
    IpmIO io{z0, lb, ub, z, lam, zL, zU, obj, kkt_err, con_inf, status, iters};
    const IpmOpts O = make_opts(opts);
#pragma omp parallel for schedule(dynamic) num_threads(nt)
    for (int b = 0; b < B; ++b) {
      int t = 0;
#ifdef _OPENMP
      t = omp_get_thread_num();
#endif
      WS<S> w(L, 0ull, nullptr, ws + MYR_WS_HEADER + (size_t)t * stride);
      ipm_solve_entry<S>(P, O, io, b, w);
    }
    return (int)MYR_OK;
  });
}

// ------------------------------------------------------------------ verification rollout (utils.py:258-298)
template <class Sys>
__global__ void rollout_kernel(RolloutArgs A) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < A.B) rollout_instance<Sys>(A, b);
}

template <class Sys>
int sys_rollout_cost(const MyrDesc* desc, int B, int nu_rows, const double* u, const double* x0, double* xs, double* cost,
                     void* stream) {
  Problem P;
  int rc = make_problem<Sys>(desc, B, P);
  if (rc) return rc;
  if (B == 0) return (int)MYR_OK;
  RolloutArgs A = make_rollout_args<Sys>(P, desc->intervals * P.cpi, nu_rows, u, x0, xs, cost);
  rollout_kernel<Sys><<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(A);
  return cuda_check("myr_rollout_cost");
}

template <class Sys>
int sys_host_rollout_cost(const MyrDesc* desc, int B, int nu_rows, const double* u, const double* x0, double* xs, double* cost) {
  Problem P;
  int rc = make_problem<Sys>(desc, B, P);
  if (rc) return rc;
  RolloutArgs A = make_rollout_args<Sys>(P, desc->intervals * P.cpi, nu_rows, u, x0, xs, cost);
  for (int b = 0; b < B; ++b) rollout_instance<Sys>(A, b);
  return (int)MYR_OK;
}

// ------------------------------------------------------------------ VJP with the compact block Jacobian
// out = J^T lam in the reference's flat layout, from the Jblk that myr_eval returned.  This is the product the
// reference's extragradient solver takes through jax.grad of the Lagrangian (nlp_solvers/extra_gradient.py:21-33).
template <class S>
MYR_HDI void jtvec_node(const Problem& P, int b, int q, int Q, int St, const double* Jblk, const double* lam, double* out) {
  constexpr int NW = S::NW, NC = S::NC;
  const long long jstride = (long long)St * S::kMaxStageNodes * NC * NW;
  const double* Jb = Jblk + (long long)b * jstride;
  const double* lb = lam + (long long)b * P.ncon;
  double acc[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) acc[i] = 0.0;
  const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
  if (jp >= 0) {
    const double* G = Jb + ((long long)jp * S::kMaxStageNodes + S::phi_slot(P, q)) * NC * NW;
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      const double l = lb[S::cidx(P, jp, r)];
#pragma unroll
      for (int i = 0; i < NW; ++i) acc[i] += G[r * NW + i] * l;
    }
  }
  if (js >= 0) {
    const double* F = Jb + ((long long)js * S::kMaxStageNodes + S::psi_slot(P, q)) * NC * NW;
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      const double l = lb[S::cidx(P, js, r)];
#pragma unroll
      for (int i = 0; i < NW; ++i) acc[i] += F[r * NW + i] * l;
    }
  }
#pragma unroll
  for (int i = 0; i < NW; ++i) out[(long long)b * P.nvars + S::zidx(P, q, i)] = acc[i];
}

template <class S>
__global__ void jtvec_kernel(Problem P, int Q, int St, const double* Jblk, const double* lam, double* out) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id < (long long)P.B * Q) jtvec_node<S>(P, (int)(id / Q), (int)(id % Q), Q, St, Jblk, lam, out);
}

// shooting: Jblk[b][k] is (n+1) x ncol (rows 0..n-1: d px_k / d (xs[k], controls of interval k)); c_k = px_k - xs[k+1].
// One thread per OUTPUT variable gathers from the (at most two) intervals that touch it: no atomics, deterministic.
template <class Sys, int NU>
MYR_HDI void jtvec_shooting_var(const Problem& P, int b, int v, const double* Jblk, const double* lam, double* out) {
  constexpr int n = Sys::n, m = Sys::m;
  constexpr int mc = NU == 3 ? 2 : 1;
  const int M = mc * P.cpi, ncol = n + (M + 1) * m, K = P.N;
  const double* lb = lam + (long long)b * P.ncon;
  const double* Jb = Jblk + (long long)b * K * (n + 1) * ncol;
  double a = 0.0;
  const int nx = (K + 1) * n;
  if (v < nx) {
    const int k = v / n, i = v % n;
    if (k < K) for (int r = 0; r < n; ++r) a += Jb[((long long)k * (n + 1) + r) * ncol + i] * lb[k * n + r];
    if (k >= 1) a -= lb[(k - 1) * n + i];
  } else {
    const int uidx = (v - nx) / m, c = (v - nx) % m;   // control row uidx, component c
    for (int k = max(0, (uidx - 1) / M - 1); k < K && k * M <= uidx; ++k) {
      const int loc = uidx - k * M;
      if (loc < 0 || loc > M) continue;
      for (int r = 0; r < n; ++r) a += Jb[((long long)k * (n + 1) + r) * ncol + n + loc * m + c] * lb[k * n + r];
    }
  }
  out[(long long)b * P.nvars + v] = a;
}

template <class Sys, int NU>
__global__ void jtvec_shooting_kernel(Problem P, const double* Jblk, const double* lam, double* out) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id < (long long)P.B * P.nvars) jtvec_shooting_var<Sys, NU>(P, (int)(id / P.nvars), (int)(id % P.nvars), Jblk, lam, out);
}

template <class Sys>
int sys_jtvec(const MyrDesc* desc, int B, const double* Jblk, const double* lam, double* out, void* stream, int host) {
  if (B > 0 && (!Jblk || !lam || !out)) return fail(MYR_E_BADARG, "null buffer passed to myr_jtvec%s", "");
  if (desc->optimizer == MYR_OPT_SHOOTING) {
    return dispatch_shooting<Sys>(desc, B, [&](auto s, const Problem& P) {
      using S = decltype(s);
      constexpr int NU = (S::NW - S::n) / S::m;
      if (B == 0) return (int)MYR_OK;
      const long long tot = (long long)B * P.nvars;
      if (host) {
        for (long long id = 0; id < tot; ++id) jtvec_shooting_var<Sys, NU>(P, (int)(id / P.nvars), (int)(id % P.nvars), Jblk, lam, out);
        return (int)MYR_OK;
      }
      jtvec_shooting_kernel<Sys, NU><<<(unsigned)((tot + 127) / 128), 128, 0, (cudaStream_t)stream>>>(P, Jblk, lam, out);
      return cuda_check("myr_jtvec(shooting)");
    });
  }
  return dispatch_scheme<Sys>(desc, B, [&](auto s, const Problem& P) {
    using S = decltype(s);
    if (B == 0) return (int)MYR_OK;
    const int Q = S::num_nodes(P), St = S::num_stages(P);
    const long long tot = (long long)B * Q;
    if (host) {
      for (long long id = 0; id < tot; ++id) jtvec_node<S>(P, (int)(id / Q), (int)(id % Q), Q, St, Jblk, lam, out);
      return (int)MYR_OK;
    }
    jtvec_kernel<S><<<(unsigned)((tot + 127) / 128), 128, 0, (cudaStream_t)stream>>>(P, Q, St, Jblk, lam, out);
    return cuda_check("myr_jtvec");
  });
}
template <class Sys>
int sys_jtvec_dev(const MyrDesc* desc, int B, const double* Jblk, const double* lam, double* out, void* stream) {
  return sys_jtvec<Sys>(desc, B, Jblk, lam, out, stream, 0);
}
template <class Sys>
int sys_jtvec_host(const MyrDesc* desc, int B, const double* Jblk, const double* lam, double* out) {
  return sys_jtvec<Sys>(desc, B, Jblk, lam, out, nullptr, 1);
}

// ------------------------------------------------------------------ point evaluation of a system (systems/base.py:45-73)
template <class Sys>
MYR_HDI void dynamics_point(const Problem& P, int b, const double* x, const double* u, const double* t, double* f, double* g) {
  constexpr int n = Sys::n, m = Sys::m;
  double xv[n], uv[m], fv[n];
#pragma unroll
  for (int i = 0; i < n; ++i) xv[i] = x[(long long)b * n + i];
#pragma unroll
  for (int i = 0; i < m; ++i) uv[i] = u[(long long)b * m + i];
  if (f) {
    dyn_f<Sys>(P, xv, uv, fv);
#pragma unroll
    for (int i = 0; i < n; ++i) f[(long long)b * n + i] = fv[i];
  }
  if (g) g[b] = Sys::cost(xv, uv, t ? t[b] : 0.0, P.p);
}

template <class Sys>
__global__ void dynamics_kernel(Problem P, const double* x, const double* u, const double* t, double* f, double* g) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < P.B) dynamics_point<Sys>(P, b, x, u, t, f, g);
}

template <class Sys>
int sys_dynamics(const MyrDesc* desc, int B, const double* x, const double* u, const double* t, double* f, double* g, void* stream) {
  Problem P;
  int rc = make_problem<Sys>(desc, B, P);
  if (rc) return rc;
  if (B == 0) return (int)MYR_OK;
  if (!x || !u) return fail(MYR_E_BADARG, "x / u is null%s", "");
  dynamics_kernel<Sys><<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(P, x, u, t, f, g);
  return cuda_check("myr_dynamics");
}

template <class Sys>
int sys_host_dynamics(const MyrDesc* desc, int B, const double* x, const double* u, const double* t, double* f, double* g) {
  Problem P;
  int rc = make_problem<Sys>(desc, B, P);
  if (rc) return rc;
  if (B > 0 && (!x || !u)) return fail(MYR_E_BADARG, "x / u is null%s", "");
  for (int b = 0; b < B; ++b) dynamics_point<Sys>(P, b, x, u, t, f, g);
  return (int)MYR_OK;
}

// the verification rollout of a NodeSystem integrates the TRUE dynamics (node_system.py:32-33, useful_scripts.py:47-49)
template <class Sys, class = void>
struct rollout_system { using type = Sys; };
template <class Sys>
struct rollout_system<Sys, typename std::enable_if<Sys::kNode>::type> { using type = typename Sys::TrueSystem; };

template <class Sys>
SysVTable make_vtable() {
  using R = typename rollout_system<Sys>::type;
  return SysVTable{Sys::id, Sys::name, &sys_problem_sizes<Sys>, &sys_eval<Sys>, &sys_host_eval<Sys>, &sys_kkt_solve<Sys>,
                   &sys_host_kkt_solve<Sys>, &sys_ipm_solve<Sys>, &sys_host_ipm_solve<Sys>, &sys_rollout_cost<R>,
                   &sys_host_rollout_cost<R>, &sys_dynamics<Sys>, &sys_host_dynamics<Sys>, &sys_jtvec_dev<Sys>, &sys_jtvec_host<Sys>};
}

}  // namespace myr
