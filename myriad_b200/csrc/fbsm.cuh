// Forward-Backward Sweep Method (indirect optimizer), batched over instances.
//
// Replaces myriad/trajectory_optimizers/forward_backward_sweep.py:88-158 (solve / sequencesolver), its RK4 sweeps
// myriad/utils.py:138-197 (integrate_fbsm) and the stopping rule myriad/trajectory_optimizers/base.py:128-141.
//
// One THREAD per instance (instances differ in their start state).  The sweeps are recurrences in time, so the
// parallelism is across instances; the three trajectories (state, control, adjoint) live in global memory TIME-MAJOR with
// the instance index fastest -- x[(i * n + k) * B + b] -- so that the 32 lanes of a warp, which are always at the same
// time index i, touch 32 consecutive doubles (one 256-byte segment) on every load and store.  The output arrays are the
// working storage: nothing else is allocated.  Per sweep iteration an instance reads/writes each trajectory entry a fixed
// number of times (forward: read u, read+write x; backward: read x, read u, read+write adj, write u), i.e.
// (N + 1) * 8 * (4 n + 3 m) bytes, which is the kernel's algorithmic HBM traffic.
//
// The control update  u <- (u* (adj, x, t) + u) / 2  (forward_backward_sweep.py:105-107) is fused into the backward sweep:
// the step that produces adj[i-1] is the last reader of the OLD u[i], so u[i] is replaced right there from the final adj[i].
// The three |.|-sums of the stopping rule are accumulated while the entries stream through registers.
#pragma once
#include <math.h>
#include <stdint.h>

#include "systems_gen.cuh"

namespace myr {

// ------------------------------------------------------------------------------------------------------------------
// Adjoint ODEs and optimality characterisations of the Lenhart & Workman systems, restated from the reference's
// hand-derived formulas (NOT from -dH/dx: a few of them deliberately differ, e.g. BEARPOPULATIONS' third row).
// p[] follows each generated system's default_params order (systems_gen.cuh).
//   adj(a, x, u, t, p, out)            <->  system.adj_ODE(adj_t, x_t, u_t, t)
//   opt(a, x, t, p, lb, ub, out)       <->  system.optim_characterization(adj_t, x_t, t)  (one time row)
// lb / ub: the bounds row(s) the reference clamps with (the host passes the row each system indexes).
// ------------------------------------------------------------------------------------------------------------------
MYR_HD double fbsm_clamp(double v, double lo, double hi) { return fmin(hi, fmax(lo, v)); }
MYR_HD double fbsm_sign(double v) { return v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0); }

template <class Sys>
struct Indirect {
  static constexpr bool available = false;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double*, const double*, const double*, double, const double*, double*) {}
  MYR_HD static void opt(const double*, const double*, double, const double*, const double*, const double*, double*) {}
};

// simple_case.py:55-63
template <>
struct Indirect<SysSimplecase> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)u; (void)t;
    o[0] = -p[0] + x[0] * a[0];
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)x; (void)t;
    o[0] = fbsm_clamp((p[2] * a[0]) / (2.0 * p[1]), lb[0], ub[0]);
  }
};

// simple_case_with_bounds.py:58-66
template <>
struct Indirect<SysSimplecasewithbounds> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)u; (void)t;
    o[0] = -p[0] + x[0] * a[0];
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)x; (void)t;
    o[0] = fbsm_clamp((p[1] * a[0]) / 2.0, lb[0], ub[0]);
  }
};

// cancer_treatment.py:88-96  (p = r, a, delta)
template <>
struct Indirect<SysCancertreatment> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)t;
    o[0] = a[0] * (p[0] + p[2] * u[0] - p[0] * log(1.0 / x[0])) - 2.0 * p[1] * x[0];
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)t;
    o[0] = fbsm_clamp(0.5 * a[0] * p[2] * x[0], lb[0], ub[0]);
  }
};

// mould_fungicide.py:62-70  (p = r, M, A)
template <>
struct Indirect<SysMouldfungicide> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)t;
    o[0] = a[0] * (p[0] + u[0]) - 2.0 * p[2] * x[0];
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)t; (void)p;
    o[0] = fbsm_clamp(0.5 * a[0] * x[0], lb[0], ub[0]);
  }
};

// bioreactor.py:82-93  (p = K, G, D): bang-bang characterisation
template <>
struct Indirect<SysBioreactor> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)t;
    o[0] = -p[0] - p[1] * u[0] * a[0] + 2.0 * p[2] * x[0] * a[0];
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)t;
    const double temp = -1.0 + p[1] * a[0] * x[0];
    const double M = fmax(fabs(lb[0]), fabs(ub[0]));
    o[0] = fbsm_clamp(fbsm_sign(temp) * 2.0 * M + M, lb[0], ub[0]);
  }
};

// glucose.py:93-105  (p = a, b, c, A, l): the characterisation is not clamped
template <>
struct Indirect<SysGlucose> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)u; (void)t;
    o[0] = -2.0 * p[3] * (x[0] - p[4]) + a[0] * p[0];
    o[1] = a[0] * p[1] + a[1] * p[2];
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)x; (void)t; (void)p; (void)lb; (void)ub;
    o[0] = -a[1] / 2.0;
  }
};

// harvest.py:64-72  (p = A, k, m): time-dependent price
template <>
struct Indirect<SysHarvest> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)x;
    o[0] = a[0] * (p[2] + u[0]) - p[0] * (p[1] * t / (t + 1.0)) * u[0];
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    o[0] = fbsm_clamp(0.5 * x[0] * (p[0] * (p[1] * t / (t + 1.0)) - a[0]), lb[0], ub[0]);
  }
};

// timber_harvest.py:76-87  (p = r, k): bang-bang characterisation
template <>
struct Indirect<SysTimberharvest> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)x;
    const double e = exp(-p[0] * t);
    o[0] = u[0] * (e - p[1] * a[0]) - e;
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    const double temp = x[0] * (p[1] * a[0] - exp(-p[0] * t));
    const double M = fmax(fabs(lb[0]), fabs(ub[0]));
    o[0] = fbsm_clamp(fbsm_sign(temp) * 2.0 * M + M, lb[0], ub[0]);
  }
};

// epidemic_seirn.py:97-112  (p = A, b, d, c, e, g, a)
template <>
struct Indirect<SysEpidemicseirn> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)t;
    const double A = p[0], b = p[1], d = p[2], c = p[3], e = p[4], g = p[5], al = p[6];
    o[0] = a[0] * (d + c * x[2] + u[0]) - a[1] * c * x[2];
    o[1] = a[1] * (e + d) - a[2] * e;
    o[2] = -A + a[0] * c * x[0] - a[1] * c * x[0] + a[2] * (g + al + d) + a[3] * al;
    o[3] = -b * a[0] + a[3] * (d - d);
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)t; (void)p;
    o[0] = fbsm_clamp(a[0] * x[0] / 2.0, lb[0], ub[0]);
  }
};

// hiv_treatment.py:113-129  (p = s, m_1, m_2, m_3, r, T_max, k, N, A)
template <>
struct Indirect<SysHivtreatment> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)t;
    const double s = p[0], m1 = p[1], m2 = p[2], m3 = p[3], r = p[4], Tm = p[5], k = p[6], N = p[7], A = p[8];
    o[0] = -A + a[0] * (m1 - r * (1.0 - (x[0] + x[1]) / Tm) + r * x[0] / Tm + u[0] * k * x[2]) - a[1] * u[0] * k * x[2];
    o[1] = a[0] * r * x[0] / Tm + a[1] * m2 - a[2] * N * m2;
    o[2] = a[0] * (s / ((1.0 + x[2]) * (1.0 + x[2])) + u[0] * k * x[0]) - a[1] * u[0] * k * x[0] + a[2] * m3;
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)t;
    o[0] = fbsm_clamp(1.0 + 0.5 * p[6] * x[0] * x[2] * (a[1] - a[0]), lb[0], ub[0]);
  }
};

// bacteria.py:88-96  (p = r, A, B, C)
template <>
struct Indirect<SysBacteria> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)t;
    o[0] = -a[0] * (p[0] + p[1] * u[0] + p[2] * u[0] * u[0] * exp(-x[0]));
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)t;
    o[0] = fbsm_clamp(a[0] * p[1] * x[0] / (2.0 * (1.0 + p[2] * a[0] * exp(-x[0]))), lb[0], ub[0]);
  }
};

// predator_prey.py:124-137  (p = d_1, d_2, A)
template <>
struct Indirect<SysPredatorprey> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)t;
    o[0] = a[0] * (x[1] - 1.0 + p[0] * u[0]) - a[1] * x[1];
    o[1] = a[0] * x[0] + a[1] * (1.0 - x[0] + p[1] * u[0]);
    o[2] = 0.0;
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)t;
    o[0] = fbsm_clamp((a[0] * p[0] * x[0] + a[1] * p[1] * x[1] - a[2]) / p[2], lb[0], ub[0]);
  }
};

// bear_populations.py:112-139  (p = r, K, m_p, m_f, c_p, c_f): two controls
template <>
struct Indirect<SysBearpopulations> {
  static constexpr bool available = true;
  static constexpr bool discrete = false;
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)t;
    const double r = p[0], K = p[1], mp = p[2], mf = p[3];
    const double k = r / K, k2 = r / (K * K);
    o[0] = a[0] * (2.0 * k * x[0] + k2 * mf * x[1] * x[1] + u[0] - r) - a[1] * (2.0 * k * mp * (1.0 - x[1] / K) * x[0]) +
           a[2] * (2.0 * k * (mp - 1.0) * x[0] - k2 * mf * x[1] * x[1] - 2.0 * k2 * mp * x[0] * x[1]);
    o[1] = a[1] * (2.0 * k * x[1] + k2 * mp * x[0] * x[0] + u[1] - r) - a[0] * (2.0 * k * mf * (1.0 - x[0] / K) * x[1]) +
           a[2] * (2.0 * k * (mf - 1.0) * x[1] - 2.0 * k2 * mf * x[0] * x[1] - k2 * mp * x[0] * x[0]);
    o[2] = -1.0;
  }
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)t;
    o[0] = fbsm_clamp(a[0] * x[0] / (2.0 * p[4]), lb[0], ub[0]);
    o[1] = fbsm_clamp(a[1] * x[1] / (2.0 * p[5]), lb[1], ub[1]);
  }
};

// INVASIVEPLANT (invasive_plant.py:38-101): the one DISCRETE member of SystemType -- x_{t+1} = f(x_t, u_t), five foci with one
// control each.  The reference accepts it only in the FBSM (the direct optimizers raise, trajectory_optimizers/base.py:66-67),
// so it has no generated NLP code: its map lives here.  p = B, k, eps.
struct SysInvasiveplant {
  static constexpr int id = MYR_SYS_INVASIVEPLANT, n = 5, m = 5, nw = 10, np = 3;
  static constexpr const char* name = "INVASIVEPLANT";
  MYR_HD static void default_params(double* p) { p[0] = 1.0; p[1] = 1.0; p[2] = 0.01; }
  MYR_HD static void f(const double* x, const double* u, const double* p, double* o) {  // next state, invasive_plant.py:71-72
    for (int k = 0; k < n; ++k) o[k] = (x[k] + x[k] * p[1] / (p[2] + x[k])) * (1.0 - u[k]);
  }
};
template <>
struct Indirect<SysInvasiveplant> {
  static constexpr bool available = true;
  static constexpr bool discrete = true;
  // previous adjoint (invasive_plant.py:79-82)
  MYR_HD static void adj(const double* a, const double* x, const double* u, double t, const double* p, double* o) {
    (void)t;
    for (int k = 0; k < 5; ++k) o[k] = a[k] * (1.0 - u[k]) * (1.0 + p[2] * p[1] / ((p[2] + x[k]) * (p[2] + x[k])));
  }
  // one row of optim_characterization (invasive_plant.py:84-90): a = adj[i + 1], x = x[i]; every control clamps with the LAST row
  MYR_HD static void opt(const double* a, const double* x, double t, const double* p, const double* lb, const double* ub, double* o) {
    (void)t;
    for (int k = 0; k < 5; ++k) o[k] = fbsm_clamp(0.5 * a[k] / p[0] * (x[k] + x[k] * p[1] / (p[2] + x[k])), lb[0], ub[0]);
  }
};

// ------------------------------------------------------------------------------------------------------------------
struct FbsmParams {
  int B, N;
  double T, delta, secant_tol;
  int max_iter, max_secant;
  int term_state;  // index of the state with a terminal value, or -1
  double term_value, guess_a, guess_b;
  double p[MYR_MAX_PARAMS];
  double adj_T[8];
  double lb[4], ub[4];
  const double* x0;  // [B][n]
  double* x;         // [N+1][n][B]
  double* u;         // [N+1][m][B]  (discrete systems: [N][m][B], forward_backward_sweep.py:37-38)
  double* adj;       // [N+1][n][B]
  int32_t* iters;    // [B] sweeps performed (summed over the secant solves)
  int32_t* status;   // [B] MYR_ST_SOLVED / MYR_ST_MAXITER / MYR_ST_NAN
};

template <class Sys>
struct FbsmInstance {
  static constexpr int n = Sys::n, m = Sys::m;
  using Ind = Indirect<Sys>;
  const FbsmParams& P;
  const size_t B, b;
  const double h;

  MYR_HD FbsmInstance(const FbsmParams& P_, int b_) : P(P_), B((size_t)P_.B), b((size_t)b_), h(P_.T / P_.N) {}

  MYR_HD double time_at(int i) const { return i == P.N ? P.T : i * h; }  // linspace(0, T, N + 1)
  MYR_HD void load(const double* a, int i, int w, double* o) const {
    for (int k = 0; k < w; ++k) o[k] = a[((size_t)i * w + k) * B + b];
  }
  MYR_HD void store(double* a, int i, int w, const double* v) const {
    for (int k = 0; k < w; ++k) a[((size_t)i * w + k) * B + b] = v[k];
  }

  // x_guess = [x_0; 0], u_guess = 0, adj_guess = [0; adj_T] with adj_guess[-1, term_state] = a
  // (forward_backward_sweep.py:40-50 and reinitiate, :75-89)
  MYR_HD void init(bool with_a, double a) const {
    double z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double v[8];
    for (int k = 0; k < n; ++k) v[k] = P.x0[b * n + k];
    store(P.x, 0, n, v);
    for (int i = 1; i <= P.N; ++i) store(P.x, i, n, z);
    for (int i = 0; i < P.N + (Ind::discrete ? 0 : 1); ++i) store(P.u, i, m, z);
    for (int i = 0; i < P.N; ++i) store(P.adj, i, n, z);
    for (int k = 0; k < n; ++k) v[k] = P.adj_T[k];
    if (with_a) v[P.term_state] = a;
    store(P.adj, P.N, n, v);
  }

  // One sweep of a DISCRETE system (integrate_fbsm's ``discrete`` branches, utils.py:182-186): x[i+1] = f(x[i], u[i]);
  // adj[i-1] = adj_ODE(adj[i], x[i], u[i-1], t[i-1]); u[i] <- (u*(adj[i+1], x[i]) + u[i]) / 2 for i < N (u has N rows).
  MYR_HD void sweep_discrete(double* sx, double* dx, double* sa, double* da, double* su, double* du) const {
    const int N = P.N;
    double xc[n], xn[n], xo[n], uc[m], ue[m], ac[n], an[n], ao[n], xp[n];
    load(P.x, 0, n, xc);
    for (int k = 0; k < n; ++k) sx[k] += fabs(xc[k]);
    for (int i = 0; i < N; ++i) {
      load(P.u, i, m, uc);
      Sys::f(xc, uc, P.p, xn);
      load(P.x, i + 1, n, xo);
      for (int k = 0; k < n; ++k) {
        xc[k] = xn[k];
        sx[k] += fabs(xc[k]);
        dx[k] += fabs(xc[k] - xo[k]);
      }
      store(P.x, i + 1, n, xc);
    }
    load(P.adj, N, n, ac);
    for (int k = 0; k < n; ++k) sa[k] += fabs(ac[k]);
    for (int i = N; i >= 1; --i) {  // xc == x[i]
      load(P.u, i - 1, m, uc);
      load(P.x, i - 1, n, xp);
      Ind::adj(ac, xc, uc, time_at(i - 1), P.p, an);
      Ind::opt(ac, xp, time_at(i - 1), P.p, P.lb, P.ub, ue);  // adj[i] is final; this step is the last reader of the old u[i-1]
      for (int k = 0; k < m; ++k) {
        const double v = 0.5 * (ue[k] + uc[k]);
        su[k] += fabs(v);
        du[k] += fabs(v - uc[k]);
        ue[k] = v;
      }
      store(P.u, i - 1, m, ue);
      load(P.adj, i - 1, n, ao);
      for (int k = 0; k < n; ++k) {
        ac[k] = an[k];
        sa[k] += fabs(ac[k]);
        da[k] += fabs(ac[k] - ao[k]);
        xc[k] = xp[k];
      }
      store(P.adj, i - 1, n, ac);
    }
  }

  // the while loop of solve(), forward_backward_sweep.py:94-110; returns the number of sweeps, sets nan_seen
  MYR_HD int sweep_until_converged(bool& nan_seen, bool& capped) const {
    const int N = P.N;
    int it = 0;
    bool go = true;
    nan_seen = false;
    capped = false;
    while (go) {
      double sx[n], dx[n], sa[n], da[n], su[m], du[m];
      for (int k = 0; k < n; ++k) sx[k] = dx[k] = sa[k] = da[k] = 0.0;
      for (int k = 0; k < m; ++k) su[k] = du[k] = 0.0;
      if constexpr (Ind::discrete) {
        sweep_discrete(sx, dx, sa, da, su, du);
      } else {
      double xc[n], uc[m], un[m], um[m], k1[n], k2[n], k3[n], k4[n], w[n], xo[n];
      // ---- forward sweep of the states (utils.py:166-176 with h > 0)
      // Loads run ONE STEP AHEAD of their use (un2 / xo2 / xp2 ...): the addresses depend only on the step index, and the
      // compiler may not hoist them over the stores itself, so each thread keeps the next step's operands in flight while
      // it works through the current step's dependent RK4 chain.
      double un2[m] = {}, xo2[n] = {};
      load(P.x, 0, n, xc);
      load(P.u, 0, m, uc);
      load(P.u, 1, m, un);
      load(P.x, 1, n, xo);
      for (int k = 0; k < n; ++k) sx[k] += fabs(xc[k]);
      for (int i = 0; i < N; ++i) {
        if (i + 1 < N) {
          load(P.u, i + 2, m, un2);
          load(P.x, i + 2, n, xo2);
        }
        for (int k = 0; k < m; ++k) um[k] = (uc[k] + un[k]) / 2.0;
        Sys::f(xc, uc, P.p, k1);
        for (int k = 0; k < n; ++k) w[k] = xc[k] + h * k1[k] / 2.0;
        Sys::f(w, um, P.p, k2);
        for (int k = 0; k < n; ++k) w[k] = xc[k] + h * k2[k] / 2.0;
        Sys::f(w, um, P.p, k3);
        for (int k = 0; k < n; ++k) w[k] = xc[k] + h * k3[k];
        Sys::f(w, un, P.p, k4);
        for (int k = 0; k < n; ++k) {
          xc[k] = xc[k] + (h / 6.0) * (k1[k] + 2.0 * k2[k] + 2.0 * k3[k] + k4[k]);
          sx[k] += fabs(xc[k]);
          dx[k] += fabs(xc[k] - xo[k]);
          xo[k] = xo2[k];
        }
        store(P.x, i + 1, n, xc);
        for (int k = 0; k < m; ++k) { uc[k] = un[k]; un[k] = un2[k]; }
      }
      // ---- backward sweep of the adjoints (h < 0) with the control update fused in
      const double hb = -h;
      double ac[n], xp[n], xm[n], up[m], ue[m], ao[n];
      load(P.adj, N, n, ac);  // terminal condition: never changes
      for (int k = 0; k < n; ++k) sa[k] += fabs(ac[k]);
      // xc == x[N], uc == u[N] (old) at this point
      double xp2[n] = {}, up2[m] = {}, ao2[n] = {};
      load(P.x, N - 1, n, xp);
      load(P.u, N - 1, m, up);
      load(P.adj, N - 1, n, ao);
      for (int i = N; i >= 1; --i) {
        const double t = time_at(i);
        if (i >= 2) {
          load(P.x, i - 2, n, xp2);
          load(P.u, i - 2, m, up2);
          load(P.adj, i - 2, n, ao2);
        }
        for (int k = 0; k < n; ++k) xm[k] = (xc[k] + xp[k]) / 2.0;
        for (int k = 0; k < m; ++k) um[k] = (uc[k] + up[k]) / 2.0;
        Ind::adj(ac, xc, uc, t, P.p, k1);
        for (int k = 0; k < n; ++k) w[k] = ac[k] + hb * k1[k] / 2.0;
        Ind::adj(w, xm, um, t + hb / 2.0, P.p, k2);
        for (int k = 0; k < n; ++k) w[k] = ac[k] + hb * k2[k] / 2.0;
        Ind::adj(w, xm, um, t + hb / 2.0, P.p, k3);
        for (int k = 0; k < n; ++k) w[k] = ac[k] + hb * k3[k];
        Ind::adj(w, xp, up, t + hb, P.p, k4);
        // u[i] <- (u*(adj[i], x[i], t_i) + u[i]) / 2: adj[i] is final, and nobody reads the old u[i] after this step
        Ind::opt(ac, xc, t, P.p, P.lb, P.ub, ue);
        for (int k = 0; k < m; ++k) {
          const double v = 0.5 * (ue[k] + uc[k]);
          su[k] += fabs(v);
          du[k] += fabs(v - uc[k]);
          ue[k] = v;
        }
        store(P.u, i, m, ue);
        for (int k = 0; k < n; ++k) {
          ac[k] = ac[k] + (hb / 6.0) * (k1[k] + 2.0 * k2[k] + 2.0 * k3[k] + k4[k]);
          sa[k] += fabs(ac[k]);
          da[k] += fabs(ac[k] - ao[k]);
          xc[k] = xp[k];
          xp[k] = xp2[k];
          ao[k] = ao2[k];
        }
        store(P.adj, i - 1, n, ac);
        for (int k = 0; k < m; ++k) { uc[k] = up[k]; up[k] = up2[k]; }
      }
      Ind::opt(ac, xc, time_at(0), P.p, P.lb, P.ub, ue);
      for (int k = 0; k < m; ++k) {
        const double v = 0.5 * (ue[k] + uc[k]);
        su[k] += fabs(v);
        du[k] += fabs(v - uc[k]);
        ue[k] = v;
      }
      store(P.u, 0, m, ue);
      }  // continuous
      ++it;
      // ---- stopping rule (base.py:128-141): continue while min(|v| sum * delta - |v - old| sum) < 0
      double mn = INFINITY;
      bool bad = false;
      for (int k = 0; k < m; ++k) { const double v = su[k] * P.delta - du[k]; if (v != v) bad = true; mn = fmin(mn, v); }
      for (int k = 0; k < n; ++k) { const double v = sx[k] * P.delta - dx[k]; if (v != v) bad = true; mn = fmin(mn, v); }
      for (int k = 0; k < n; ++k) { const double v = sa[k] * P.delta - da[k]; if (v != v) bad = true; mn = fmin(mn, v); }
      if (bad) { nan_seen = true; go = false; }  // jnp.min propagates NaN and NaN < 0 is False: the reference stops too
      else go = mn < 0.0;
      if (go && it >= P.max_iter) { capped = true; go = false; }
    }
    return it;
  }

  MYR_HD double terminal_gap() const { return P.x[((size_t)P.N * n + P.term_state) * B + b] - P.term_value; }

  MYR_HD void run() const {
    bool nan_seen = false, capped = false, any_nan = false, any_cap = false;
    int total = 0;
    if (P.term_state < 0) {
      init(false, 0.0);
      total = sweep_until_converged(nan_seen, capped);
      any_nan = nan_seen; any_cap = capped;
    } else {
      // secant iteration on the free terminal adjoint (sequencesolver, forward_backward_sweep.py:118-158)
      double a = P.guess_a, c = P.guess_b;
      init(true, a);
      total += sweep_until_converged(nan_seen, capped); any_nan |= nan_seen; any_cap |= capped;
      double Va = terminal_gap();
      init(true, c);
      total += sweep_until_converged(nan_seen, capped); any_nan |= nan_seen; any_cap |= capped;
      double Vc = terminal_gap();
      int count = 0;
      while (fabs(Va) > P.secant_tol) {
        if (count >= P.max_secant) { any_cap = true; break; }
        if (fabs(Va) > fabs(Vc)) {
          double s = a; a = c; c = s;
          s = Va; Va = Vc; Vc = s;
        }
        const double d = Va * (c - a) / (Vc - Va);
        c = a;
        Vc = Va;
        a = a - d;
        init(true, a);
        total += sweep_until_converged(nan_seen, capped); any_nan |= nan_seen; any_cap |= capped;
        Va = terminal_gap();
        ++count;
      }
      if (Va != Va) any_nan = true;
    }
    P.iters[b] = total;
    P.status[b] = any_nan ? MYR_ST_NAN : (any_cap ? MYR_ST_MAXITER : MYR_ST_SOLVED);
  }
};

}  // namespace myr
