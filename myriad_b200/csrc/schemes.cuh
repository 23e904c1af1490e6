// Transcription schemes as node/stage-structured NLPs.
//
// Every direct transcription of the reference is the same shape of NLP:
//   nodes q = 0..Q-1 with NW variables each (block v_q), stages j = 0..S-1 with NC equality rows each,
//   objective  sum_q ell_q(v_q),
//   stage constraint  c_j = sum over the nodes of stage j of their role function:
//        role PHI: phi_q(v_q)   (node's contribution to the stage it "starts")
//        role PSI: psi_q(v_q)   (node's contribution to the previous stage it "ends")
// so the Lagrangian Hessian is block-diagonal per node and the constraint Jacobian is block
// bi-diagonal in stages -> Schur complement J H^-1 J^T is block tridiagonal (NC x NC blocks).
//
//   trapezoid (myriad/trajectory_optimizers/collocation/trapezoidal.py:151-163, 183-192)
//       c_j = h/2 (f_j + f_{j+1}) - (x_{j+1} - x_j)
//       phi_j = h/2 f(v_j) + x_j ,  psi_{j+1} = h/2 f(v_{j+1}) - x_{j+1}
//   Hermite-Simpson (collocation/hermite_simpson.py:110-170, 325-335), nodes = knots and mid points
//       defect  c_j = (x_b - x_a) - h/6 (f_a + 4 f_m + f_b) ;  interp d_j = x_m - (x_a + x_b)/2 - h/8 (f_a - f_b)
//
// A scheme supplies index maps between the reference's flat layouts (ravel_pytree((x,u)) and the
// constraint vector) and nodes/stages, plus eval_node<MODE>().
#pragma once
#include "common.cuh"
#include "node_mlp.cuh"

namespace myr {

// Collocation on a NODE system never evaluates the MLP per thread on the device: the cooperative tensor-core pass
// (mlp_nodes_pass) has filled PreDyn for every node.  Compiling the scalar fallback out of device code keeps the
// kernels' stack frames small.
#ifdef __CUDA_ARCH__
#define MYR_DIRECT_DYN(call) do { if constexpr (!sys_is_node<Sys>::value) { call; } } while (0)
#else
#define MYR_DIRECT_DYN(call) do { call; } while (0)
#endif

// MODE: 0 = values only (ell, phi, psi); 1 = + first derivatives (gl, G, F); 2 = + Hessian block W
// W (packed upper NW) = hess( ell + lam_phi . phi + lam_psi . psi ) w.r.t. v.

template <class Sys>
struct Trapezoid {
  using System = Sys;
  static constexpr int n = Sys::n, m = Sys::m;
  static constexpr int NW = n + m, NC = n, NWP = NW * (NW + 1) / 2;
  static constexpr int kMaxStageNodes = 2;
  // role Jacobians are affine in the dynamics Jacobian J (n x NW):  G = a_p J + b_p [I 0],  F = a_s J + b_s [I 0]  per
  // row group (NG groups of n rows).  The interior-point kernel stores J once instead of G and F (engine.cuh).
  static constexpr bool kAffineJ = true;
  static constexpr int NG = 1;
  MYR_HDI static void role_coefs(const Problem& P, int q, double* ap, double* bp, double* as, double* bs) {
    const double hh = 0.5 * P.h;
    const bool has_phi = q < P.N, has_psi = q >= 1;
    ap[0] = has_phi ? hh : 0.0; bp[0] = has_phi ? 1.0 : 0.0;
    as[0] = has_psi ? hh : 0.0; bs[0] = has_psi ? -1.0 : 0.0;
  }

  MYR_HDI static int num_nodes(const Problem& P) { return P.N + 1; }
  MYR_HDI static int num_stages(const Problem& P) { return P.N; }
  MYR_HDI static int nvars(const Problem& P) { return (P.N + 1) * NW; }
  MYR_HDI static int ncon(const Problem& P) { return P.N * n; }
  // reference flat index of variable i of node q: states time-major, then controls time-major
  MYR_HDI static int zidx(const Problem& P, int q, int i) { return i < n ? q * n + i : (P.N + 1) * n + q * m + (i - n); }
  MYR_HDI static int cidx(const Problem&, int j, int r) { return j * n + r; }
  // stages a node contributes to (-1: none)
  MYR_HDI static int phi_stage(const Problem& P, int q) { return q < P.N ? q : -1; }
  MYR_HDI static int psi_stage(const Problem&, int q) { return q >= 1 ? q - 1 : -1; }
  // nodes of a stage: k = 0..stage_nodes-1 -> node id, role (0 phi, 1 psi)
  MYR_HDI static int stage_nodes(const Problem&, int) { return 2; }
  MYR_HDI static int stage_node(const Problem&, int j, int k, int& role) { role = k; return j + k; }
  // the node shared with the next stage (psi role here, phi role there)
  MYR_HDI static int link_node(const Problem&, int j) { return j + 1; }
  // position of the node's block inside its stage row of Jblk
  MYR_HDI static int phi_slot(const Problem&, int) { return 0; }
  MYR_HDI static int psi_slot(const Problem&, int) { return 1; }

  // weights of the dynamics Hessians in the node's Lagrangian block: mu_i = d (lam . c) / d f_i(v_q);
  // lam = the instance's multipliers in the reference's constraint order
  // lam(j, r): multiplier of row r of stage j
  template <class LamView>
  MYR_HDI static void node_mu(const Problem& P, int q, const LamView& lam, double* mu) {
    const double hh = 0.5 * P.h;
#pragma unroll
    for (int i = 0; i < n; ++i) mu[i] = hh * ((q < P.N ? lam(q, i) : 0.0) + (q >= 1 ? lam(q - 1, i) : 0.0));
  }

  template <int MODE>
  MYR_HDI static void eval_node(const Problem& P, int q, const double* v, const double* lam_phi, const double* lam_psi,
                                double& ell, double* gl, double* phi, double* psi, double* G, double* F, double* W,
                                const PreDyn& pre = PreDyn()) {
    eval_node_impl<MODE, false>(P, q, v, lam_phi, lam_psi, ell, gl, phi, psi, G, F, W, nullptr, pre);
  }
  // same, but returns the raw dynamics Jacobian J (n x NW) instead of the role Jacobians G, F
  template <int MODE>
  MYR_HDI static void eval_node_j(const Problem& P, int q, const double* v, const double* lam_phi, const double* lam_psi,
                                  double& ell, double* gl, double* phi, double* psi, double* Jraw, double* W,
                                  const PreDyn& pre = PreDyn()) {
    eval_node_impl<MODE, true>(P, q, v, lam_phi, lam_psi, ell, gl, phi, psi, nullptr, nullptr, W, Jraw, pre);
  }
  template <int MODE, bool RAWJ>
  MYR_HDI static void eval_node_impl(const Problem& P, int q, const double* v, const double* lam_phi, const double* lam_psi,
                                     double& ell, double* gl, double* phi, double* psi, double* G, double* F, double* W,
                                     double* Jraw, const PreDyn& pre) {
    const double h = P.h;
    const double hh = 0.5 * h;
    const bool has_phi = q < P.N, has_psi = q >= 1;
    const double wq = (q == 0 || q == P.N) ? hh : h;  // trapezoid quadrature weights, trapezoidal.py:115-128
    const double t = (q == P.N) ? P.T : q * h;       // jnp.linspace(0, T, N+1)[q]
    double f[n];
    if (MODE == 0) {
      if (pre.f) {
#pragma unroll
        for (int i = 0; i < n; ++i) f[i] = pre.f[i];
      } else {
        MYR_DIRECT_DYN(dyn_f<Sys>(P, v, v + n, f));
      }
      ell = wq * Sys::cost(v, v + n, t, P.p);
    } else {
      double J[n * NW];
      if (pre.f) {
#pragma unroll
        for (int i = 0; i < n; ++i) f[i] = pre.f[i];
#pragma unroll
        for (int i = 0; i < n * NW; ++i) J[i] = pre.J[i];
      }
      if (MODE == 2) {
        if (pre.f) {
#pragma unroll
          for (int i = 0; i < NWP; ++i) W[i] = pre.H[i];
        } else {
          double mu[n];
#pragma unroll
          for (int i = 0; i < n; ++i) mu[i] = hh * ((has_phi ? lam_phi[i] : 0.0) + (has_psi ? lam_psi[i] : 0.0));
#pragma unroll
          for (int i = 0; i < NWP; ++i) W[i] = 0.0;
          MYR_DIRECT_DYN(dyn_fjac_hess<Sys>(P, v, v + n, mu, f, J, W));
        }
        ell = wq * Sys::cost_grad_hess(v, v + n, t, P.p, wq, gl, W);
      } else {
        if (!pre.f) { MYR_DIRECT_DYN(dyn_fjac<Sys>(P, v, v + n, f, J)); }
        ell = wq * Sys::cost_grad(v, v + n, t, P.p, gl);
      }
#pragma unroll
      for (int i = 0; i < NW; ++i) gl[i] *= wq;
      if (RAWJ) {
#pragma unroll
        for (int i = 0; i < n * NW; ++i) Jraw[i] = J[i];
      } else {
#pragma unroll
        for (int r = 0; r < n; ++r)
#pragma unroll
          for (int i = 0; i < NW; ++i) {
            const double a = hh * J[r * NW + i];
            const double e = (i == r) ? 1.0 : 0.0;
            G[r * NW + i] = has_phi ? a + e : 0.0;
            F[r * NW + i] = has_psi ? a - e : 0.0;
          }
      }
    }
#pragma unroll
    for (int r = 0; r < n; ++r) {
      phi[r] = has_phi ? hh * f[r] + v[r] : 0.0;
      psi[r] = has_psi ? hh * f[r] - v[r] : 0.0;
    }
    // terminal cost on the last node: trapezoidal.py:126-127 (linear in x_T for every reference system: no Hessian term)
    if (Sys::has_terminal && P.terminal_cost && q == P.N) {
      double tc[n];
      Sys::terminal_coef(P.p, tc);
#pragma unroll
      for (int i = 0; i < n; ++i) {
        ell += tc[i] * v[i];
        if (MODE >= 1) gl[i] += tc[i];
      }
    }
  }
};

template <class Sys>
struct HermiteSimpson {
  using System = Sys;
  static constexpr int n = Sys::n, m = Sys::m;
  static constexpr int NW = n + m, NC = 2 * n, NWP = NW * (NW + 1) / 2;
  static constexpr int kMaxStageNodes = 3;
  // row groups: 0 = defect rows, 1 = interpolation rows (see Trapezoid::kAffineJ)
  static constexpr bool kAffineJ = true;
  static constexpr int NG = 2;
  MYR_HDI static void role_coefs(const Problem& P, int q, double* ap, double* bp, double* as, double* bs) {
    const double h = P.h;
    const bool mid = (q & 1);
    const bool has_phi = q < 2 * P.N, has_psi = (!mid) && q >= 2;
    ap[0] = has_phi ? (mid ? -4.0 * h / 6.0 : -h / 6.0) : 0.0; bp[0] = has_phi ? (mid ? 0.0 : -1.0) : 0.0;
    ap[1] = has_phi ? (mid ? 0.0 : -h / 8.0) : 0.0;            bp[1] = has_phi ? (mid ? 1.0 : -0.5) : 0.0;
    as[0] = has_psi ? -h / 6.0 : 0.0; bs[0] = has_psi ? 1.0 : 0.0;
    as[1] = has_psi ? h / 8.0 : 0.0;  bs[1] = has_psi ? -0.5 : 0.0;
  }

  MYR_HDI static int num_nodes(const Problem& P) { return 2 * P.N + 1; }
  MYR_HDI static int num_stages(const Problem& P) { return P.N; }
  MYR_HDI static int nvars(const Problem& P) { return (2 * P.N + 1) * NW; }
  MYR_HDI static int ncon(const Problem& P) { return 2 * P.N * n; }
  MYR_HDI static int zidx(const Problem& P, int q, int i) { return i < n ? q * n + i : (2 * P.N + 1) * n + q * m + (i - n); }
  // hstack(all defects, all interpolations): hermite_simpson.py:325-335
  MYR_HDI static int cidx(const Problem& P, int j, int r) { return r < n ? j * n + r : P.N * n + j * n + (r - n); }
  MYR_HDI static int phi_stage(const Problem& P, int q) { return q < 2 * P.N ? q / 2 : -1; }          // knots and mids
  MYR_HDI static int psi_stage(const Problem&, int q) { return (q % 2 == 0 && q >= 2) ? q / 2 - 1 : -1; }  // end knots
  MYR_HDI static int stage_nodes(const Problem&, int) { return 3; }
  MYR_HDI static int stage_node(const Problem&, int j, int k, int& role) { role = (k == 2) ? 1 : 0; return 2 * j + k; }
  MYR_HDI static int link_node(const Problem&, int j) { return 2 * j + 2; }
  MYR_HDI static int phi_slot(const Problem&, int q) { return q & 1; }
  MYR_HDI static int psi_slot(const Problem&, int) { return 2; }

  template <class LamView>
  MYR_HDI static void node_mu(const Problem& P, int q, const LamView& lam, double* mu) {
    const double h = P.h;
    const bool mid = (q & 1);
    const bool has_phi = q < 2 * P.N, has_psi = (!mid) && q >= 2;
    const int jp = q / 2, js = q / 2 - 1;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      double a = 0.0;
      if (has_phi) a += (mid ? -4.0 * h / 6.0 : -h / 6.0) * lam(jp, i) + (mid ? 0.0 : -h / 8.0) * lam(jp, n + i);
      if (has_psi) a += (-h / 6.0) * lam(js, i) + (h / 8.0) * lam(js, n + i);
      mu[i] = a;
    }
  }

  template <int MODE>
  MYR_HDI static void eval_node(const Problem& P, int q, const double* v, const double* lam_phi, const double* lam_psi,
                                double& ell, double* gl, double* phi, double* psi, double* G, double* F, double* W,
                                const PreDyn& pre = PreDyn()) {
    eval_node_impl<MODE, false>(P, q, v, lam_phi, lam_psi, ell, gl, phi, psi, G, F, W, nullptr, pre);
  }
  // same, but returns the raw dynamics Jacobian J (n x NW) instead of the role Jacobians G, F
  template <int MODE>
  MYR_HDI static void eval_node_j(const Problem& P, int q, const double* v, const double* lam_phi, const double* lam_psi,
                                  double& ell, double* gl, double* phi, double* psi, double* Jraw, double* W,
                                  const PreDyn& pre = PreDyn()) {
    eval_node_impl<MODE, true>(P, q, v, lam_phi, lam_psi, ell, gl, phi, psi, nullptr, nullptr, W, Jraw, pre);
  }
  template <int MODE, bool RAWJ>
  MYR_HDI static void eval_node_impl(const Problem& P, int q, const double* v, const double* lam_phi, const double* lam_psi,
                                     double& ell, double* gl, double* phi, double* psi, double* G, double* F, double* W,
                                     double* Jraw, const PreDyn& pre) {
    const double h = P.h;
    const bool mid = (q & 1);
    const bool has_phi = q < 2 * P.N, has_psi = (!mid) && q >= 2;
    // objective sum_k h/6 (g_a + 4 g_m + g_b): hermite_simpson.py:243-257
    const double wq = mid ? 4.0 * h / 6.0 : ((q == 0 || q == 2 * P.N) ? h / 6.0 : 2.0 * h / 6.0);
    const double t = (q == 2 * P.N) ? P.T : q * (P.T / (2 * P.N));  // linspace(0, T, 2N+1)[q]
    // role coefficients: phi rows [defect; interp], psi rows [defect; interp]
    //  knot as start (a): defect -x_a - h/6 f_a ; interp -x_a/2 - h/8 f_a
    //  mid  (m)         : defect      - 4h/6 f_m ; interp  x_m
    //  knot as end (b)  : defect  x_b - h/6 f_b ; interp -x_b/2 + h/8 f_b
    const double cf_phi_d = mid ? -4.0 * h / 6.0 : -h / 6.0, cx_phi_d = mid ? 0.0 : -1.0;
    const double cf_phi_i = mid ? 0.0 : -h / 8.0, cx_phi_i = mid ? 1.0 : -0.5;
    const double cf_psi_d = -h / 6.0, cx_psi_d = 1.0, cf_psi_i = h / 8.0, cx_psi_i = -0.5;
    double f[n];
    if (MODE == 0) {
      if (pre.f) {
#pragma unroll
        for (int i = 0; i < n; ++i) f[i] = pre.f[i];
      } else {
        MYR_DIRECT_DYN(dyn_f<Sys>(P, v, v + n, f));
      }
      ell = wq * Sys::cost(v, v + n, t, P.p);
    } else {
      double J[n * NW];
      if (pre.f) {
#pragma unroll
        for (int i = 0; i < n; ++i) f[i] = pre.f[i];
#pragma unroll
        for (int i = 0; i < n * NW; ++i) J[i] = pre.J[i];
      }
      if (MODE == 2) {
        if (pre.f) {
#pragma unroll
          for (int i = 0; i < NWP; ++i) W[i] = pre.H[i];
        } else {
          double mu[n];
#pragma unroll
          for (int i = 0; i < n; ++i) {
            double a = 0.0;
            if (has_phi) a += cf_phi_d * lam_phi[i] + cf_phi_i * lam_phi[n + i];
            if (has_psi) a += cf_psi_d * lam_psi[i] + cf_psi_i * lam_psi[n + i];
            mu[i] = a;
          }
#pragma unroll
          for (int i = 0; i < NWP; ++i) W[i] = 0.0;
          MYR_DIRECT_DYN(dyn_fjac_hess<Sys>(P, v, v + n, mu, f, J, W));
        }
        ell = wq * Sys::cost_grad_hess(v, v + n, t, P.p, wq, gl, W);
      } else {
        if (!pre.f) { MYR_DIRECT_DYN(dyn_fjac<Sys>(P, v, v + n, f, J)); }
        ell = wq * Sys::cost_grad(v, v + n, t, P.p, gl);
      }
#pragma unroll
      for (int i = 0; i < NW; ++i) gl[i] *= wq;
      if (RAWJ) {
#pragma unroll
        for (int i = 0; i < n * NW; ++i) Jraw[i] = J[i];
      } else {
#pragma unroll
        for (int r = 0; r < n; ++r)
#pragma unroll
          for (int i = 0; i < NW; ++i) {
            const double a = J[r * NW + i];
            const double e = (i == r) ? 1.0 : 0.0;
            G[r * NW + i] = has_phi ? cf_phi_d * a + cx_phi_d * e : 0.0;
            G[(n + r) * NW + i] = has_phi ? cf_phi_i * a + cx_phi_i * e : 0.0;
            F[r * NW + i] = has_psi ? cf_psi_d * a + cx_psi_d * e : 0.0;
            F[(n + r) * NW + i] = has_psi ? cf_psi_i * a + cx_psi_i * e : 0.0;
          }
      }
    }
#pragma unroll
    for (int r = 0; r < n; ++r) {
      phi[r] = has_phi ? cf_phi_d * f[r] + cx_phi_d * v[r] : 0.0;
      phi[n + r] = has_phi ? cf_phi_i * f[r] + cx_phi_i * v[r] : 0.0;
      psi[r] = has_psi ? cf_psi_d * f[r] + cx_psi_d * v[r] : 0.0;
      psi[n + r] = has_psi ? cf_psi_i * f[r] + cx_psi_i * v[r] : 0.0;
    }
  }
};

}  // namespace myr
