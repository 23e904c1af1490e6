// Multiple / single shooting (myriad/trajectory_optimizers/shooting.py:15-278) as a LIFTED node/stage NLP.
//
// The reference's shooting NLP has the interval-start states xs[k] and all controls us[] as variables and
// one constraint block per interval, px_k(xs[k], us) - xs[k+1] = 0, where px_k is a cpi-step rollout
// (myriad/utils.py:80-131); its objective integrates the cost along the same rollouts (shooting.py:169-210).
// Eliminating nothing instead -- every integrator step g = k*cpi + j becomes a stage with its own state s_g --
// gives an NLP with the SAME feasible set and optimum (hidden states are determined by the step equations) whose
// Lagrangian Hessian is block-diagonal per step and whose Jacobian is block-bidiagonal, i.e. exactly the
// structure K2/K3 exploit for collocation.  Steps of HEUN / MIDPOINT / RK4 read the NEXT node's control
// (utils.py:41-50, 33-38); a per-node copy ubar_g with the linear consensus row  ubar_g - u_{g+1} = 0  keeps the
// Hessian block-diagonal.
//
//   node g (g = 0..G, G = K*cpi):  v_g = [ s_g (n) | u-slot 0 (m) | ... | u-slot NU-1 (m) ]
//       NU = 1 (EULER: u_g), 2 (HEUN, MIDPOINT: u_g, ubar_g), 3 (RK4: u_2g, u_2g+1, ubar_g)
//   stage g (g = 0..G-1):  [ Phi(v_g) - s_{g+1} ;  ubar_g - u_{g+1, slot 0} ]      NC = n (+ m if NU > 1)
//   cost of stage g: the step's quadrature of g(x,u,t) with the integrator's own stage points (cost-augmented
//   dynamics, shooting.py:80-92).
//
// At shooting nodes (g multiple of cpi) s_g IS the reference variable xs[g/cpi] (with its bounds / fixed values);
// other s_g are hidden free variables initialised by rolling out the guess.  The multiplier of the dynamics rows of
// stage (k+1)*cpi-1 is the reference's multiplier of constraint block k (same sign: px - x_next).
#pragma once
#include "common.cuh"
#include "node_mlp.cuh"

namespace myr {

template <class Sys, int NU>
struct ShootingLifted {
  using System = Sys;
  static constexpr int n = Sys::n, m = Sys::m;
  static constexpr int NW = n + NU * m, NC = n + (NU > 1 ? m : 0), NWP = NW * (NW + 1) / 2;
  static constexpr int kMaxStageNodes = 2;
  static constexpr bool kLifted = true;
  static constexpr int kCalls = NU == 1 ? 1 : (NU == 2 ? 2 : 4);
  static constexpr int mc = NU == 3 ? 2 : 1;  // midpoint controls are decision variables for RK4 (shooting.py:31)

  MYR_HDI static int G(const Problem& P) { return P.N * P.cpi; }
  MYR_HDI static int num_nodes(const Problem& P) { return G(P) + 1; }
  MYR_HDI static int num_stages(const Problem& P) { return G(P); }
  // sizes of the REFERENCE NLP
  MYR_HDI static int nvars(const Problem& P) { return (P.N + 1) * n + (mc * G(P) + 1) * m; }
  MYR_HDI static int ncon(const Problem& P) { return P.N * n; }
  // sizes of the lifted NLP (internal)
  MYR_HDI static int nvars_int(const Problem& P) { return num_nodes(P) * NW; }
  MYR_HDI static int ncon_int(const Problem& P) { return num_stages(P) * NC; }
  // internal flat layouts: node-major / stage-major
  MYR_HDI static int zidx(const Problem&, int q, int i) { return q * NW + i; }
  MYR_HDI static int cidx(const Problem&, int j, int r) { return j * NC + r; }
  MYR_HDI static int phi_stage(const Problem& P, int q) { return q < G(P) ? q : -1; }
  MYR_HDI static int psi_stage(const Problem&, int q) { return q >= 1 ? q - 1 : -1; }
  MYR_HDI static int stage_nodes(const Problem&, int) { return 2; }
  MYR_HDI static int stage_node(const Problem&, int j, int k, int& role) { role = k; return j + k; }
  MYR_HDI static int link_node(const Problem&, int j) { return j + 1; }
  MYR_HDI static int phi_slot(const Problem&, int) { return 0; }
  MYR_HDI static int psi_slot(const Problem&, int) { return 1; }

  // reference index (ravel_pytree((xs, us))) of internal variable i of node q, or -1 (hidden state / control copy)
  MYR_HDI static int ref_index(const Problem& P, int q, int i) {
    const int Gn = G(P);
    if (i < n) return (q % P.cpi == 0) ? (q / P.cpi) * n + i : -1;
    const int slot = (i - n) / m, c = (i - n) % m;
    const int ubase = (P.N + 1) * n;
    if (NU == 3) {
      if (slot == 0) return ubase + (2 * q) * m + c;
      if (slot == 1) return q < Gn ? ubase + (2 * q + 1) * m + c : -1;
      return -1;
    }
    return slot == 0 ? ubase + q * m + c : -1;
  }
  // variables that do not enter any function: padding slots of the last node, and with EULER the final control
  // (utils.py:115-116 never reads it: SURVEY.md section 9-15).  They are held at their initial value.
  MYR_HDI static bool is_dead(const Problem& P, int q, int i) {
    if (i < n) return false;
    const int slot = (i - n) / m;
    if (q == G(P)) return NU == 1 ? true : slot >= 1;
    return false;
  }

  // ---- one explicit Runge-Kutta step with first/second derivatives w.r.t. the node block v
  template <int MODE>
  MYR_HDI static void eval_node(const Problem& P, int q, const double* v, const double* lam_phi, const double* lam_psi,
                                double& ell, double* gl, double* phi, double* psi, double* Gm, double* Fm, double* W,
                                const PreDyn& = PreDyn()) {
    const int Gn = G(P);
    const bool has_phi = q < Gn, has_psi = q >= 1;
    const double h = P.h;
    // psi role: -s_q (dynamics rows of the previous stage), -u_{q,slot0} (its consensus rows)
#pragma unroll
    for (int r = 0; r < NC; ++r) psi[r] = has_psi ? -v[r] : 0.0;  // rows 0..n-1 <-> s, rows n..n+m-1 <-> slot 0 (contiguous in v)
    if (MODE >= 1) {
#pragma unroll
      for (int r = 0; r < NC; ++r)
#pragma unroll
        for (int i = 0; i < NW; ++i) Fm[r * NW + i] = (has_psi && i == r) ? -1.0 : 0.0;
#pragma unroll
      for (int i = 0; i < NW; ++i) gl[i] = 0.0;
#pragma unroll
      for (int i = 0; i < NC * NW; ++i) Gm[i] = 0.0;
    }
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < NWP; ++i) W[i] = 0.0;
    }
    ell = 0.0;
#pragma unroll
    for (int r = 0; r < NC; ++r) phi[r] = 0.0;
    if (!has_phi) {
      // last node: terminal cost on the end state of the last interval (shooting.py:205-208), linear in x_T
      if (Sys::has_terminal && P.terminal_cost) {
        double tc[n];
        Sys::terminal_coef(P.p, tc);
#pragma unroll
        for (int i = 0; i < n; ++i) {
          ell += tc[i] * v[i];
          if (MODE >= 1) gl[i] += tc[i];
        }
      }
      return;
    }

    // Butcher data per method.  a[c] = coefficient of the PREVIOUS call's k in x_c (all four schemes only chain one
    // call back), b[c] = weight of k_c in the update, tc[c] = time offset / h, control of call c = sum_s cu[c][s] u_slot_s
    double a[4] = {0, 0, 0, 0}, bw[4] = {0, 0, 0, 0}, tc[4] = {0, 0, 0, 0};
    double cu[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    int ncall = 1;
    if (NU == 1) {  // EULER, utils.py:53-54
      bw[0] = 1.0; cu[0][0] = 1.0;
    } else if (NU == 2) {
      ncall = 2;
      if (P.method == MIDPOINT) {  // utils.py:47-50: x_mid = x + h f(x,u1); x + h f(x_mid,(u1+u2)/2)
        a[1] = 1.0; bw[1] = 1.0; tc[1] = 0.5; cu[0][0] = 1.0; cu[1][0] = 0.5; cu[1][1] = 0.5;
      } else {                     // HEUN, utils.py:41-44
        a[1] = 1.0; bw[0] = 0.5; bw[1] = 0.5; tc[1] = 1.0; cu[0][0] = 1.0; cu[1][1] = 1.0;
      }
    } else {                       // RK4, utils.py:33-38
      ncall = 4;
      a[1] = 0.5; a[2] = 0.5; a[3] = 1.0;
      bw[0] = 1.0 / 6; bw[1] = 2.0 / 6; bw[2] = 2.0 / 6; bw[3] = 1.0 / 6;
      tc[1] = 0.5; tc[2] = 0.5; tc[3] = 1.0;
      cu[0][0] = 1.0; cu[1][1] = 1.0; cu[2][1] = 1.0; cu[3][2] = 1.0;
    }
    const double t0 = q * h;  // linspace(0, T, num_steps+1)[q]

    // forward sweep: stage points, k_c, and (MODE>=1) sensitivities
    double xk[kCalls][n], uk[kCalls][m], kk[kCalls][n];
    double Xs[kCalls][n * NW];       // d x_c / d v
    double Ks[kCalls][n * NW];       // d k_c / d v
    double Aj[kCalls][n * (n + m)];  // [A|B] of call c
    double gyc[kCalls][n + m];       // grad of running cost at call c (w.r.t. y=(x,u))
    double gval[kCalls];
    for (int c = 0; c < ncall; ++c) {
#pragma unroll
      for (int i = 0; i < n; ++i) xk[c][i] = v[i] + (c > 0 ? h * a[c] * kk[c - 1][i] : 0.0);
#pragma unroll
      for (int j = 0; j < m; ++j) {
        double u = 0.0;
#pragma unroll
        for (int s = 0; s < NU; ++s) u += cu[c][s] * v[n + s * m + j];
        uk[c][j] = u;
      }
      if (MODE == 0) {
        dyn_f<Sys>(P, xk[c], uk[c], kk[c]);
        gval[c] = Sys::cost(xk[c], uk[c], t0 + tc[c] * h, P.p);
      } else {
        dyn_fjac<Sys>(P, xk[c], uk[c], kk[c], Aj[c]);
        gval[c] = Sys::cost_grad(xk[c], uk[c], t0 + tc[c] * h, P.p, gyc[c]);
        // X_c = [I 0] + h a_c K_{c-1}
#pragma unroll
        for (int r = 0; r < n; ++r)
#pragma unroll
          for (int i = 0; i < NW; ++i) Xs[c][r * NW + i] = ((i == r) ? 1.0 : 0.0) + (c > 0 ? h * a[c] * Ks[c - 1][r * NW + i] : 0.0);
        // K_c = A X_c + B Uc,  Uc[j][n + s*m + j] = cu[c][s]
#pragma unroll
        for (int r = 0; r < n; ++r)
#pragma unroll
          for (int i = 0; i < NW; ++i) {
            double s_ = 0.0;
#pragma unroll
            for (int k2 = 0; k2 < n; ++k2) s_ += Aj[c][r * (n + m) + k2] * Xs[c][k2 * NW + i];
            if (i >= n) {
              const int sl = (i - n) / m, j = (i - n) % m;
              s_ += Aj[c][r * (n + m) + n + j] * cu[c][sl];
            }
            Ks[c][r * NW + i] = s_;
          }
      }
    }
    // values
    double cst = 0.0;
    for (int c = 0; c < ncall; ++c) cst += bw[c] * gval[c];
    ell = h * cst;
#pragma unroll
    for (int r = 0; r < n; ++r) {
      double s_ = v[r];
      for (int c = 0; c < ncall; ++c) s_ += h * bw[c] * kk[c][r];
      phi[r] = s_;
    }
    if (NU > 1) {
#pragma unroll
      for (int j = 0; j < m; ++j) phi[n + j] = v[n + (NU - 1) * m + j];  // + ubar
    }
    if (MODE == 0) return;
    // first derivatives
#pragma unroll
    for (int r = 0; r < n; ++r)
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        double s_ = (i == r) ? 1.0 : 0.0;
        for (int c = 0; c < ncall; ++c) s_ += h * bw[c] * Ks[c][r * NW + i];
        Gm[r * NW + i] = s_;
      }
    if (NU > 1) {
#pragma unroll
      for (int j = 0; j < m; ++j) Gm[(n + j) * NW + n + (NU - 1) * m + j] = 1.0;
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      double s_ = 0.0;
      for (int c = 0; c < ncall; ++c) {
        double t_ = 0.0;
#pragma unroll
        for (int k2 = 0; k2 < n; ++k2) t_ += gyc[c][k2] * Xs[c][k2 * NW + i];
        if (i >= n) {
          const int sl = (i - n) / m, j = (i - n) % m;
          t_ += gyc[c][n + j] * cu[c][sl];
        }
        s_ += h * bw[c] * t_;
      }
      gl[i] = s_;
    }
    if (MODE < 2) return;
    // second-order adjoint sweep: abar_c = dL/dk_c,  W = sum_c Y_c^T [ Hf(abar_c; y_c) + h b_c Hg(y_c) ] Y_c
    double ab[kCalls][n];
    for (int c = ncall - 1; c >= 0; --c) {
#pragma unroll
      for (int r = 0; r < n; ++r) {
        double s_ = h * bw[c] * lam_phi[r];
        if (c + 1 < ncall) {
          // x_{c+1} = s + h a_{c+1} k_c  ->  contribution h a_{c+1} (A_{c+1}^T abar_{c+1} + h b_{c+1} gx_{c+1})
          double t_ = h * bw[c + 1] * gyc[c + 1][r];
#pragma unroll
          for (int k2 = 0; k2 < n; ++k2) t_ += Aj[c + 1][k2 * (n + m) + r] * ab[c + 1][k2];
          s_ += h * a[c + 1] * t_;
        }
        ab[c][r] = s_;
      }
    }
    for (int c = 0; c < ncall; ++c) {
      double Hy[(n + m) * (n + m + 1) / 2];
#pragma unroll
      for (int i = 0; i < (n + m) * (n + m + 1) / 2; ++i) Hy[i] = 0.0;
      double fd[n], Jd[n * (n + m)], gd[n + m];
      dyn_fjac_hess<Sys>(P, xk[c], uk[c], ab[c], fd, Jd, Hy);
      Sys::cost_grad_hess(xk[c], uk[c], t0 + tc[c] * h, P.p, h * bw[c], gd, Hy);
      // Y_c rows: x rows = Xs[c], u rows j: e_{n+s*m+j} * cu[c][s]
      // T = Hy * Y  ((n+m) x NW), then W += Y^T T
      double T[(n + m) * NW];
#pragma unroll
      for (int a_ = 0; a_ < n + m; ++a_)
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          double s_ = 0.0;
#pragma unroll
          for (int k2 = 0; k2 < n; ++k2) s_ += Hy[pidx(a_, k2, n + m)] * Xs[c][k2 * NW + i];
          if (i >= n) {
            const int sl = (i - n) / m, j = (i - n) % m;
            s_ += Hy[pidx(a_, n + j, n + m)] * cu[c][sl];
          }
          T[a_ * NW + i] = s_;
        }
#pragma unroll
      for (int i = 0; i < NW; ++i)
#pragma unroll
        for (int j2 = i; j2 < NW; ++j2) {
          double s_ = 0.0;
#pragma unroll
          for (int k2 = 0; k2 < n; ++k2) s_ += Xs[c][k2 * NW + i] * T[k2 * NW + j2];
          if (i >= n) {
            const int sl = (i - n) / m, j = (i - n) % m;
            s_ += cu[c][sl] * T[(n + j) * NW + j2];
          }
          W[pidx(i, j2, NW)] += s_;
        }
    }
  }
};

// ---- reference-level evaluation of the shooting NLP for one interval (values and forward-mode Jacobian):
// px_k, cost_k and their sensitivities w.r.t. (xs[k], the interval's mc*cpi+1 controls), by chaining the per-step
// derivatives of the lifted scheme.  Used by myr_eval for OPT_SHOOTING.
template <class Sys, int NU>
struct ShootingInterval {
  using L = ShootingLifted<Sys, NU>;
  static constexpr int n = Sys::n, m = Sys::m, NW = L::NW, NC = L::NC;
  static constexpr int mc = L::mc;

  // z: reference decision vector of the instance.  Outputs (any may be null):
  //   px[n] end state, cost (scalar), Jx [n][n] = d px / d xs[k], Ju [n][(mc*cpi+1)*m] = d px / d interval controls,
  //   gx [n], gu [(mc*cpi+1)*m] = gradient of the interval cost.  Sx/Su are caller scratch of the same sizes as Jx/Ju (+1 row).
  template <bool DERIV>
  MYR_HDI static void run(const Problem& P, int k, const double* z, double* px, double& cost, double* S /* (n+1) x ncol */, int ncol) {
    const int M = mc * P.cpi;
    const int ubase = (P.N + 1) * n;
    double s[n];
#pragma unroll
    for (int i = 0; i < n; ++i) s[i] = z[k * n + i];
    cost = 0.0;
    if (DERIV) {
      for (int i = 0; i < (n + 1) * ncol; ++i) S[i] = 0.0;
#pragma unroll
      for (int i = 0; i < n; ++i) S[i * ncol + i] = 1.0;
    }
    for (int j = 0; j < P.cpi; ++j) {
      const int g = k * P.cpi + j;
      double v[NW];
#pragma unroll
      for (int i = 0; i < n; ++i) v[i] = s[i];
      // controls of this step in interval-local numbering: slot s_ <-> local control index lj
      int lidx[NU];
#pragma unroll
      for (int sl = 0; sl < NU; ++sl) {
        lidx[sl] = (NU == 3) ? 2 * j + sl : j + sl;
#pragma unroll
        for (int c = 0; c < m; ++c) v[n + sl * m + c] = z[ubase + (k * M + lidx[sl]) * m + c];
      }
      double ell, gl[NW], phi[NC], psi[NC], Gm[NC * NW], Fm[NC * NW], W[1];
      double lam0[NC];
#pragma unroll
      for (int r = 0; r < NC; ++r) lam0[r] = 0.0;
      Problem Pl = P;  // the lifted scheme uses q only for the time and for has_phi/has_psi
      if (DERIV) L::template eval_node<1>(Pl, g, v, lam0, lam0, ell, gl, phi, psi, Gm, Fm, W);
      else L::template eval_node<0>(Pl, g, v, lam0, lam0, ell, gl, phi, psi, Gm, Fm, W);
      if (DERIV) {
        // new sensitivities: rows 0..n-1: Phi_s S_x + Phi_u e ; row n (cost): old + gl_s S_x + gl_u e
        double Sn[n];
        for (int col = 0; col < ncol; ++col) {
          double cc = S[n * ncol + col];
#pragma unroll
          for (int r = 0; r < n; ++r) {
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < n; ++i) a += Gm[r * NW + i] * S[i * ncol + col];
            Sn[r] = a;
          }
#pragma unroll
          for (int i = 0; i < n; ++i) cc += gl[i] * S[i * ncol + col];
#pragma unroll
          for (int r = 0; r < n; ++r) S[r * ncol + col] = Sn[r];
          S[n * ncol + col] = cc;
        }
#pragma unroll
        for (int sl = 0; sl < NU; ++sl)
#pragma unroll
          for (int c = 0; c < m; ++c) {
            const int col = n + lidx[sl] * m + c;
#pragma unroll
            for (int r = 0; r < n; ++r) S[r * ncol + col] += Gm[r * NW + n + sl * m + c];
            S[n * ncol + col] += gl[n + sl * m + c];
          }
      }
      cost += ell;
#pragma unroll
      for (int i = 0; i < n; ++i) s[i] = phi[i];
    }
#pragma unroll
    for (int i = 0; i < n; ++i) px[i] = s[i];
    if (Sys::has_terminal && P.terminal_cost && k == P.N - 1) {  // shooting.py:205-208
      double tc[n];
      Sys::terminal_coef(P.p, tc);
#pragma unroll
      for (int i = 0; i < n; ++i) cost += tc[i] * s[i];
      if (DERIV) {
        for (int col = 0; col < ncol; ++col) {
          double a = 0.0;
#pragma unroll
          for (int i = 0; i < n; ++i) a += tc[i] * S[i * ncol + col];
          S[n * ncol + col] += a;
        }
      }
    }
  }
};

}  // namespace myr
