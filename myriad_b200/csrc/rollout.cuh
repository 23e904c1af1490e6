// Fixed-step rollouts of the cost-augmented dynamics: the reference's integrate()
// (myriad/utils.py:22-73) as used by get_state_trajectory_and_cost (myriad/utils.py:258-298).
// One thread integrates one instance; the four IntegrationMethod schemes follow utils.py:33-54,
// including the full-step "midpoint" state (utils.py:47-50) and JAX's clamp-to-last indexing of the
// control array.
#pragma once
#include "common.cuh"

namespace myr {

struct RolloutArgs {
  int B, num_steps, nu_rows, method, terminal_cost;
  double h, T;
  double p[kMaxParams];
  const double* u;   // [B][nu_rows][m]
  const double* x0;  // [B][n]
  double* xs;        // [B][num_steps+1][n] or null
  double* cost;      // [B]
};

template <class Sys>
static RolloutArgs make_rollout_args(const Problem& P, int num_steps, int nu_rows, const double* u, const double* x0, double* xs, double* cost) {
  RolloutArgs A;
  A.B = P.B; A.num_steps = num_steps; A.nu_rows = nu_rows; A.method = P.method; A.terminal_cost = P.terminal_cost;
  A.T = P.T; A.h = P.T / num_steps;
  for (int i = 0; i < kMaxParams; ++i) A.p[i] = P.p[i];
  A.u = u; A.x0 = x0; A.xs = xs; A.cost = cost;
  return A;
}

// augmented derivative [f(x,u); g(x,u,t)]
template <class Sys>
MYR_HDI void aug_dyn(const double* p, const double* xc, const double* u, double t, double* out) {
  Sys::f(xc, u, p, out);
  out[Sys::n] = Sys::cost(xc, u, t, p);
}

template <class Sys>
MYR_HDI void rollout_instance(const RolloutArgs& A, int b) {
  constexpr int n = Sys::n, m = Sys::m, na = n + 1;
  const double h = A.h;
  const double* ub = A.u + (long long)b * A.nu_rows * m;
  auto U = [&](int i, double* dst) {
    const int k = i < A.nu_rows ? i : A.nu_rows - 1;
#pragma unroll
    for (int c = 0; c < m; ++c) dst[c] = ub[k * m + c];
  };
  double x[na];
#pragma unroll
  for (int i = 0; i < n; ++i) x[i] = A.x0[(long long)b * n + i];
  x[n] = 0.0;
  double* xs = A.xs ? A.xs + (long long)b * (A.num_steps + 1) * n : nullptr;
  if (xs) {
#pragma unroll
    for (int i = 0; i < n; ++i) xs[i] = x[i];
  }
  for (int idx = 0; idx < A.num_steps; ++idx) {
    // times = linspace(0, T, num_steps + 1)
    const double t = (idx == A.num_steps) ? A.T : idx * (A.T / A.num_steps);
    double k1[na], k2[na], k3[na], k4[na], y[na], u1[m], u2[m], u3[m];
    if (A.method == EULER) {
      U(idx, u1);
      aug_dyn<Sys>(A.p, x, u1, t, k1);
#pragma unroll
      for (int i = 0; i < na; ++i) x[i] += h * k1[i];
    } else if (A.method == HEUN) {
      U(idx, u1); U(idx + 1, u2);
      aug_dyn<Sys>(A.p, x, u1, t, k1);
#pragma unroll
      for (int i = 0; i < na; ++i) y[i] = x[i] + h * k1[i];
      aug_dyn<Sys>(A.p, y, u2, t + h, k2);
#pragma unroll
      for (int i = 0; i < na; ++i) x[i] += h / 2 * (k1[i] + k2[i]);
    } else if (A.method == MIDPOINT) {
      U(idx, u1); U(idx + 1, u2);
      aug_dyn<Sys>(A.p, x, u1, t, k1);
#pragma unroll
      for (int i = 0; i < na; ++i) y[i] = x[i] + h * k1[i];
#pragma unroll
      for (int c = 0; c < m; ++c) u3[c] = (u1[c] + u2[c]) / 2;
      aug_dyn<Sys>(A.p, y, u3, t + h / 2, k2);
#pragma unroll
      for (int i = 0; i < na; ++i) x[i] += h * k2[i];
    } else {
      U(2 * idx, u1); U(2 * idx + 1, u2); U(2 * idx + 2, u3);
      aug_dyn<Sys>(A.p, x, u1, t, k1);
#pragma unroll
      for (int i = 0; i < na; ++i) y[i] = x[i] + h * k1[i] / 2;
      aug_dyn<Sys>(A.p, y, u2, t + h / 2, k2);
#pragma unroll
      for (int i = 0; i < na; ++i) y[i] = x[i] + h * k2[i] / 2;
      aug_dyn<Sys>(A.p, y, u2, t + h / 2, k3);
#pragma unroll
      for (int i = 0; i < na; ++i) y[i] = x[i] + h * k3[i];
      aug_dyn<Sys>(A.p, y, u3, t + h, k4);
#pragma unroll
      for (int i = 0; i < na; ++i) x[i] += h / 6 * (k1[i] + 2 * k2[i] + 2 * k3[i] + k4[i]);
    }
    if (xs) {
#pragma unroll
      for (int i = 0; i < n; ++i) xs[(idx + 1) * n + i] = x[i];
    }
  }
  double cst = x[n];
  if (Sys::has_terminal && A.terminal_cost) {  // utils.py:294-295: + terminal_cost_fn(x(T), us[-1]), linear in x(T)
    double tc[n];
    Sys::terminal_coef(A.p, tc);
#pragma unroll
    for (int i = 0; i < n; ++i) cst += tc[i] * x[i];
  }
  A.cost[b] = cst;
}

}  // namespace myr
