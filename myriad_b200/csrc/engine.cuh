// Per-instance engine: K1 (evaluation of the transcribed NLP with block Jacobian / Hessian),
// K2 (block-structured KKT solve: node-block inverses, Schur complement, block cyclic reduction with
// inertia) and K3 (primal-dual interior-point iteration).  One CTA works on one problem instance;
// threads stride over nodes / stages.  Written once for device and host (see common.cuh).
#pragma once
#include <type_traits>

#include "schemes.cuh"
#include "shooting.cuh"

namespace myr {

struct IpmOpts {
  int max_iter;
  int max_ls;
  int acceptable_iter;
  int reserved;
  double tol, acceptable_tol;
  double mu_init, mu_min, kappa_eps, kappa_mu, theta_mu, tau_min;
  double bound_push, bound_frac, bound_relax, kappa_sigma, s_max;
  double delta_min, delta_0, delta_max, delta_c, kappa_w_minus, kappa_w_plus, kappa_w_plus_first;
  double eta, rho;
  double delta_reg;   // tiny primal regularisation of the node blocks, removed again by iterative refinement
  int max_refine;
  int reserved2;
};

enum Status : int { ST_SOLVED = 0, ST_ACCEPTABLE = 1, ST_MAXITER = -1, ST_LINESEARCH = -2, ST_INERTIA = -3, ST_NAN = -13 };

// schemes whose internal NLP differs from the reference's (lifted shooting) declare kLifted
template <class S, class = void>
struct scheme_is_lifted { static constexpr bool value = false; };
template <class S>
struct scheme_is_lifted<S, typename std::enable_if<S::kLifted>::type> { static constexpr bool value = true; };

// ------------------------------------------------------------------ workspace layout (doubles / instance)
template <class S>
struct Layout {
  // collocation on a NODE system: the MLP is evaluated for all nodes of the instance cooperatively (node_mlp.cuh)
  static constexpr bool kCoopMlp = sys_is_node<typename S::System>::value && !scheme_is_lifted<S>::value;
  int Q, St;
  int G, F, W, Hinv, gl, phi, psi, rb, dz, dzL, dzU, c, dlam;
  int crD, crU, crVL, crVU, crb, crx;
  int dynf, dynJ, dynH;   // NODE systems: per-node MLP dynamics values / Jacobians / contracted Hessians
  int ext;
  int total;
  MYR_HDI explicit Layout(const Problem& P) {
    Q = S::num_nodes(P); St = S::num_stages(P);
    int o = 0;
    G = o; o += Q * S::NC * S::NW;
    F = o; o += Q * S::NC * S::NW;
    W = o; o += Q * S::NWP;
    Hinv = o; o += Q * S::NW * S::NW;
    gl = o; o += Q * S::NW;
    phi = o; o += Q * S::NC;
    psi = o; o += Q * S::NC;
    rb = o; o += Q * S::NW;
    dz = o; o += Q * S::NW;
    dzL = o; o += Q * S::NW;
    dzU = o; o += Q * S::NW;
    c = o; o += St * S::NC;
    dlam = o; o += St * S::NC;
    crD = o; o += St * S::NC * S::NC;
    crU = o; o += St * S::NC * S::NC;
    crVL = o; o += St * S::NC * S::NC;
    crVU = o; o += St * S::NC * S::NC;
    crb = o; o += St * S::NC;
    crx = dlam;  // CR writes its solution straight into dlam
    dynf = dynJ = dynH = o;
    if (kCoopMlp) {
      dynf = o; o += Q * S::n;
      dynJ = o; o += Q * S::n * S::NW;
      dynH = o; o += Q * S::NWP;
    }
    // lifted schemes keep their internal iterate / bounds / multipliers in the workspace too
    ext = o;
    if (scheme_is_lifted<S>::value) o += 6 * Q * S::NW + St * S::NC;
    total = (o + 15) & ~15;
  }
  // doubles of the CR scratch (D,U,VL,VU,b), contiguous from crD
  MYR_HDI int cr_doubles() const { return St * (4 * S::NC * S::NC + S::NC); }
};

// ------------------------------------------------------------------ bounds helpers
struct Bnd {
  bool fixed, hasL, hasU;
  double lbr, ubr;
};
MYR_HDI Bnd make_bnd(double lb, double ub, double relax) {
  Bnd b;
  b.fixed = (lb == ub);
  b.hasL = !b.fixed && isfinite(lb);
  b.hasU = !b.fixed && isfinite(ub);
  b.lbr = b.hasL ? lb - relax * fmax(1.0, fabs(lb)) : lb;
  b.ubr = b.hasU ? ub + relax * fmax(1.0, fabs(ub)) : ub;
  return b;
}

// ------------------------------------------------------------------ K1: node evaluation sweep
// Evaluates every node at point z (reference layout), stores node arrays in the workspace and returns the
// objective.  MODE as in schemes.cuh.  zsrc may be the iterate or a trial point.
template <class S, int MODE>
MYR_HDI double eval_nodes(const Problem& P, const Layout<S>& L, const double* z, const double* lam, double* w, double* red,
                          double* mlp_scr = nullptr) {
  const int Q = L.Q;
  double fsum = 0.0;
  bool have_pre = false;
#ifdef __CUDA_ARCH__
  if constexpr (Layout<S>::kCoopMlp) {
    mlp_nodes_pass<S, MODE>(P, Q, z, lam, w + L.dynf, w + L.dynJ, w + L.dynH, mlp_scr);
    have_pre = true;
  }
#endif
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
    PreDyn pre;
    if (have_pre) { pre.f = w + L.dynf + q * S::n; pre.J = w + L.dynJ + q * S::n * S::NW; pre.H = w + L.dynH + q * S::NWP; }
    double v[S::NW], lp[S::NC], ls[S::NC];
#pragma unroll
    for (int i = 0; i < S::NW; ++i) v[i] = z[S::zidx(P, q, i)];
    const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
    if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < S::NC; ++r) {
        lp[r] = jp >= 0 ? lam[S::cidx(P, jp, r)] : 0.0;
        ls[r] = js >= 0 ? lam[S::cidx(P, js, r)] : 0.0;
      }
    }
    double ell, gl[S::NW], phi[S::NC], psi[S::NC], G[S::NC * S::NW], F[S::NC * S::NW], W[S::NWP];
    S::template eval_node<MODE>(P, q, v, lp, ls, ell, gl, phi, psi, G, F, W, pre);
    fsum += ell;
#pragma unroll
    for (int r = 0; r < S::NC; ++r) { w[L.phi + q * S::NC + r] = phi[r]; w[L.psi + q * S::NC + r] = psi[r]; }
    if (MODE >= 1) {
#pragma unroll
      for (int i = 0; i < S::NW; ++i) w[L.gl + q * S::NW + i] = gl[i];
#pragma unroll
      for (int i = 0; i < S::NC * S::NW; ++i) { w[L.G + q * S::NC * S::NW + i] = G[i]; w[L.F + q * S::NC * S::NW + i] = F[i]; }
    }
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < S::NWP; ++i) w[L.W + q * S::NWP + i] = W[i];
    }
  }
  const double f = block_sum(fsum, red);
  MYR_SYNC();
  return f;
}

// stage constraints from node role values; writes c (stage-major) and returns (inf-norm, 1-norm)
template <class S>
MYR_HDI void stage_constraints(const Problem& P, const Layout<S>& L, double* w, double* cdst, double* red, double& cinf, double& c1) {
  double mx = 0.0, sm = 0.0;
  for (int j = MYR_TID; j < L.St; j += MYR_NT) {
    const int nk = S::stage_nodes(P, j);
#pragma unroll
    for (int r = 0; r < S::NC; ++r) {
      double a = 0.0;
      for (int k = 0; k < nk; ++k) {
        int role; const int q = S::stage_node(P, j, k, role);
        a += role ? w[L.psi + q * S::NC + r] : w[L.phi + q * S::NC + r];
      }
      cdst[j * S::NC + r] = a;
      const double aa = (a != a) ? INFINITY : fabs(a);
      mx = fmax(mx, aa); sm += aa;
    }
  }
  cinf = block_max(mx, red);
  c1 = block_sum(sm, red);
  MYR_SYNC();
}

// ------------------------------------------------------------------ K2 pieces
// small dense helpers on row-major blocks
template <int R, int K, int C>
MYR_HDI void mm(const double* A, const double* Bm, double* Cm) {  // C = A(RxK) B(KxC)
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) s += A[i * K + k] * Bm[k * C + j];
      Cm[i * C + j] = s;
    }
}
template <int R, int K, int C>
MYR_HDI void mm_nt_sub(const double* A, const double* Bm, double* Cm) {  // C -= A(RxK) B(CxK)^T
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < C; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) s += A[i * K + k] * Bm[j * K + k];
      Cm[i * C + j] -= s;
    }
}

// Block cyclic reduction for the symmetric block-tridiagonal system
//   U_{i-1}^T x_{i-1} + D_i x_i + U_i x_{i+1} = b_i,   i = 0..St-1,  blocks NC x NC (D symmetric, may be indefinite).
// Factor and solve are fused (single right-hand side).  Pivot-block inertias are accumulated: by Sylvester's
// law their sum is the inertia of the whole matrix.  D,U,b are destroyed; x receives the solution.
template <int NC>
MYR_HDI void block_cr_solve(int St, double* D, double* U, double* VL, double* VU, double* b, double* x,
                            double* red, int& npos, int& nneg, int& nzero) {
  constexpr int BB = NC * NC;
  int cp = 0, cn = 0, cz = 0;
  int s = 1;
  for (; s < St; s <<= 1) {
    // eliminate odd multiples of s
    for (int i = (2 * MYR_TID + 1) * s; i < St; i += 2 * MYR_NT * s) {
      double A[BB], Dinv[BB];
#pragma unroll
      for (int k = 0; k < BB; ++k) A[k] = D[i * BB + k];
      int p_, n_, z_;
      sym_inverse<NC>(A, 0u, Dinv, p_, n_, z_);
      cp += p_; cn += n_; cz += z_;
      // VL = Dinv * U_{i-s}^T
      const double* Ul = U + (i - s) * BB;
      double T[BB];
#pragma unroll
      for (int r = 0; r < NC; ++r)
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NC; ++k) a += Dinv[r * NC + k] * Ul[c * NC + k];
          T[r * NC + c] = a;
        }
#pragma unroll
      for (int k = 0; k < BB; ++k) VL[i * BB + k] = T[k];
      if (i + s < St) {
        mm<NC, NC, NC>(Dinv, U + i * BB, T);
#pragma unroll
        for (int k = 0; k < BB; ++k) VU[i * BB + k] = T[k];
      }
      double vb[NC];
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) a += Dinv[r * NC + k] * b[i * NC + k];
        vb[r] = a;
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) b[i * NC + r] = vb[r];
#pragma unroll
      for (int k = 0; k < BB; ++k) D[i * BB + k] = Dinv[k];  // keep the pivot inverse for later re-solves
    }
    MYR_SYNC();
    // update even multiples of s
    for (int i = 2 * MYR_TID * s; i < St; i += 2 * MYR_NT * s) {
      double Dn[BB], bn[NC], Un[BB];
#pragma unroll
      for (int k = 0; k < BB; ++k) { Dn[k] = D[i * BB + k]; Un[k] = 0.0; }
#pragma unroll
      for (int r = 0; r < NC; ++r) bn[r] = b[i * NC + r];
      const int er = i + s;
      if (er < St) {
        const double* Ui = U + i * BB;
        const double* vl = VL + er * BB;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < NC; ++k) a += Ui[r * NC + k] * vl[k * NC + c];
            Dn[r * NC + c] -= a;
          }
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NC; ++k) a += Ui[r * NC + k] * b[er * NC + k];
          bn[r] -= a;
        }
        if (er + s < St) {
          const double* vu = VU + er * BB;
#pragma unroll
          for (int r = 0; r < NC; ++r)
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              double a = 0.0;
#pragma unroll
              for (int k = 0; k < NC; ++k) a += Ui[r * NC + k] * vu[k * NC + c];
              Un[r * NC + c] = -a;
            }
        }
      }
      const int el = i - s;
      if (el >= 0) {
        const double* Ue = U + el * BB;     // coupling el -> i  (row el, col i); we need Ue^T
        const double* vu = VU + el * BB;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < NC; ++k) a += Ue[k * NC + r] * vu[k * NC + c];
            Dn[r * NC + c] -= a;
          }
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NC; ++k) a += Ue[k * NC + r] * b[el * NC + k];
          bn[r] -= a;
        }
      }
      // symmetrise D (rounding) and store
#pragma unroll
      for (int r = 0; r < NC; ++r)
#pragma unroll
        for (int c = 0; c < NC; ++c) D[i * BB + r * NC + c] = 0.5 * (Dn[r * NC + c] + Dn[c * NC + r]);
#pragma unroll
      for (int k = 0; k < BB; ++k) U[i * BB + k] = Un[k];
#pragma unroll
      for (int r = 0; r < NC; ++r) b[i * NC + r] = bn[r];
    }
    MYR_SYNC();
  }
  // root
  if (MYR_TID == 0) {
    double A[BB], Dinv[BB];
#pragma unroll
    for (int k = 0; k < BB; ++k) A[k] = D[k];
    int p_, n_, z_;
    sym_inverse<NC>(A, 0u, Dinv, p_, n_, z_);
    cp += p_; cn += n_; cz += z_;
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NC; ++k) a += Dinv[r * NC + k] * b[k];
      x[r] = a;
    }
#pragma unroll
    for (int k = 0; k < BB; ++k) D[k] = Dinv[k];
  }
  MYR_SYNC();
  // back substitution
  for (s >>= 1; s >= 1; s >>= 1) {
    for (int i = (2 * MYR_TID + 1) * s; i < St; i += 2 * MYR_NT * s) {
      double xi[NC];
#pragma unroll
      for (int r = 0; r < NC; ++r) xi[r] = b[i * NC + r];
      const double* vl = VL + i * BB;
      const double* xl = x + (i - s) * NC;
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) a += vl[r * NC + k] * xl[k];
        xi[r] -= a;
      }
      if (i + s < St) {
        const double* vu = VU + i * BB;
        const double* xr = x + (i + s) * NC;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NC; ++k) a += vu[r * NC + k] * xr[k];
          xi[r] -= a;
        }
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) x[i * NC + r] = xi[r];
    }
    MYR_SYNC();
  }
  npos = (int)(block_sum((double)cp, red) + 0.5);
  nneg = (int)(block_sum((double)cn, red) + 0.5);
  nzero = (int)(block_sum((double)cz, red) + 0.5);
  MYR_SYNC();
}

// Re-solve with the factors left by block_cr_solve (pivot inverses in D, VL, VU) for a new right-hand side b
// (destroyed); x receives the solution.  One barrier per level: since VL_e = Dinv_e U_{e-s}^T and VU_e = Dinv_e U_e,
// the elimination of e from its surviving neighbours is  b_i -= VL_e^T b_e  (right neighbour e = i+s) and
// b_i -= VU_e^T b_e (left neighbour e = i-s), which only needs data of already-final eliminated nodes.
template <int NC>
MYR_HDI void block_cr_resolve(int St, const double* D, const double* VL, const double* VU, double* b, double* x) {
  constexpr int BB = NC * NC;
  int s = 1;
  for (; s < St; s <<= 1) {
    for (int i = 2 * MYR_TID * s; i < St; i += 2 * MYR_NT * s) {
      double bn[NC];
#pragma unroll
      for (int r = 0; r < NC; ++r) bn[r] = b[i * NC + r];
      const int er = i + s, el = i - s;
      if (er < St) {
        const double* vl = VL + er * BB;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NC; ++k) a += vl[k * NC + r] * b[er * NC + k];
          bn[r] -= a;
        }
      }
      if (el >= 0) {
        const double* vu = VU + el * BB;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NC; ++k) a += vu[k * NC + r] * b[el * NC + k];
          bn[r] -= a;
        }
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) b[i * NC + r] = bn[r];
    }
    MYR_SYNC();
  }
  if (MYR_TID == 0) {
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NC; ++k) a += D[r * NC + k] * b[k];
      x[r] = a;
    }
  }
  MYR_SYNC();
  for (s >>= 1; s >= 1; s >>= 1) {
    for (int i = (2 * MYR_TID + 1) * s; i < St; i += 2 * MYR_NT * s) {
      double xi[NC];
      const double* di = D + i * BB;
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) a += di[r * NC + k] * b[i * NC + k];
        xi[r] = a;
      }
      const double* vl = VL + i * BB;
      const double* xl = x + (i - s) * NC;
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) a += vl[r * NC + k] * xl[k];
        xi[r] -= a;
      }
      if (i + s < St) {
        const double* vu = VU + i * BB;
        const double* xr = x + (i + s) * NC;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NC; ++k) a += vu[r * NC + k] * xr[k];
          xi[r] -= a;
        }
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) x[i * NC + r] = xi[r];
    }
    MYR_SYNC();
  }
}

// KKT solve for one instance:
//   [ H + Sigma + dw I   J^T ] [dz  ]     [ rb ]
//   [ J               -dc I ] [dlam] = - [ c  ]
// node data (G,F,W,rb) and c are in the workspace; sigma/fixed per variable via callback arrays.
// Returns inertia-ok flag; dz (node-major) and dlam (stage-major) in the workspace.
template <class S>
MYR_HDI bool kkt_solve(const Problem& P, const Layout<S>& L, double* w, const double* sigma /*node-major*/,
                       const uint32_t* fixmask /*per node*/, double delta_w, double delta_c, double* cr, double* red,
                       double delta_reg = 0.0, int max_refine = 0) {
  constexpr int NW = S::NW, NC = S::NC;
  const int Q = L.Q, St = L.St;
  // ---- node blocks: Hinv = (W + Sigma + dw)^-1 with fixed variables removed; inertia of H
  int hp = 0, hn = 0, hz = 0;
  double minpr = INFINITY;
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
    double A[NW * NW], inv[NW * NW];
    const double* Wq = w + L.W + q * S::NWP;
#pragma unroll
    for (int i = 0; i < NW; ++i)
#pragma unroll
      for (int j = 0; j < NW; ++j) A[i * NW + j] = Wq[pidx(i, j, NW)];
#pragma unroll
    for (int i = 0; i < NW; ++i) A[i * NW + i] += sigma[q * NW + i] + delta_w + delta_reg;
    int p_, n_, z_;
    double pr_;
    sym_inverse<NW>(A, fixmask[q], inv, p_, n_, z_, &pr_);
    minpr = fmin(minpr, pr_);
    hp += p_; hn += n_; hz += z_;
#pragma unroll
    for (int i = 0; i < NW * NW; ++i) w[L.Hinv + q * NW * NW + i] = inv[i];
  }
  const int Hneg = (int)(block_sum((double)hn, red) + 0.5);
  const int Hzero = (int)(block_sum((double)hz, red) + 0.5);
  minpr = block_min(minpr, red);
  // refinement is only worth its cost when some node block was close to singular
  if (!(minpr < 1e-4)) max_refine = 0;
  (void)hp;
  MYR_SYNC();
  // ---- stage blocks of the Schur complement S = J Hinv J^T + dc I and its right-hand side  c - J Hinv rb
  double* D = cr; double* U = D + St * NC * NC; double* VL = U + St * NC * NC; double* VU = VL + St * NC * NC;
  double* bb = VU + St * NC * NC;
  for (int j = MYR_TID; j < St; j += MYR_NT) {
    double Dj[NC * NC], bj[NC];
#pragma unroll
    for (int i = 0; i < NC * NC; ++i) Dj[i] = 0.0;
#pragma unroll
    for (int r = 0; r < NC; ++r) { Dj[r * NC + r] = delta_c; bj[r] = w[L.c + j * NC + r]; }
    const int nk = S::stage_nodes(P, j);
    for (int k = 0; k < nk; ++k) {
      int role; const int q = S::stage_node(P, j, k, role);
      const double* Jq = w + (role ? L.F : L.G) + q * NC * NW;
      const double* Hi = w + L.Hinv + q * NW * NW;
      double T[NC * NW];
      mm<NC, NW, NW>(Jq, Hi, T);
#pragma unroll
      for (int r = 0; r < NC; ++r) {
#pragma unroll
        for (int c2 = 0; c2 < NC; ++c2) {
          double a = 0.0;
#pragma unroll
          for (int i = 0; i < NW; ++i) a += T[r * NW + i] * Jq[c2 * NW + i];
          Dj[r * NC + c2] += a;
        }
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < NW; ++i) a += T[r * NW + i] * w[L.rb + q * NW + i];
        bj[r] -= a;
      }
      if (role == 1 && j + 1 < St) {  // link node: coupling to the next stage  U_j = F Hinv G^T
        const double* Gq = w + L.G + q * NC * NW;
#pragma unroll
        for (int r = 0; r < NC; ++r)
#pragma unroll
          for (int c2 = 0; c2 < NC; ++c2) {
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < NW; ++i) a += T[r * NW + i] * Gq[c2 * NW + i];
            U[j * NC * NC + r * NC + c2] = a;
          }
      }
    }
#pragma unroll
    for (int r = 0; r < NC; ++r)
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) D[j * NC * NC + r * NC + c2] = 0.5 * (Dj[r * NC + c2] + Dj[c2 * NC + r]);
#pragma unroll
    for (int r = 0; r < NC; ++r) bb[j * NC + r] = bj[r];
  }
  MYR_SYNC();
  int sp, sn, sz;
  block_cr_solve<NC>(St, D, U, VL, VU, bb, w + L.dlam, red, sp, sn, sz);
  // inertia(K) = inertia(H) + inertia(-S): correct iff  n-(S) == n-(H)  and nothing is singular
  const bool ok = (Hzero == 0) && (sz == 0) && (sn == Hneg);
  // ---- dz = -Hinv (rb + G^T dlam_phi + F^T dlam_psi)
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
    double v[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) v[i] = w[L.rb + q * NW + i];
    const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
    if (jp >= 0) {
      const double* Gq = w + L.G + q * NC * NW;
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        const double d = w[L.dlam + jp * NC + r];
#pragma unroll
        for (int i = 0; i < NW; ++i) v[i] += Gq[r * NW + i] * d;
      }
    }
    if (js >= 0) {
      const double* Fq = w + L.F + q * NC * NW;
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        const double d = w[L.dlam + js * NC + r];
#pragma unroll
        for (int i = 0; i < NW; ++i) v[i] += Fq[r * NW + i] * d;
      }
    }
    const double* Hi = w + L.Hinv + q * NW * NW;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NW; ++k) a += Hi[i * NW + k] * v[k];
      w[L.dz + q * NW + i] = -a;
    }
  }
  MYR_SYNC();
  // ---- iterative refinement against the matrix WITHOUT delta_reg: the node blocks are factorised with a tiny
  // regularisation (directions in which W + Sigma is singular, e.g. a state that enters neither cost nor dynamics and is
  // far from its bounds, would otherwise make the block elimination break down although the KKT matrix is regular);
  // the refinement removes its effect and recovers the digits the Schur complement loses.
  for (int itr = 0; ok && itr < max_refine; ++itr) {
    // node residual  rz = -rb - (H dz + G^T dlam_phi + F^T dlam_psi)   (stored in dzL), norms
    double rmax = 0.0, smax = 0.0;
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      double v[NW], d[NW];
#pragma unroll
      for (int i = 0; i < NW; ++i) { v[i] = -w[L.rb + q * NW + i]; d[i] = w[L.dz + q * NW + i]; smax = fmax(smax, fabs(v[i])); }
      const double* Wq = w + L.W + q * S::NWP;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        double a = (sigma[q * NW + i] + delta_w) * d[i];
#pragma unroll
        for (int k = 0; k < NW; ++k) a += Wq[pidx(i, k, NW)] * d[k];
        v[i] -= a;
      }
      const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
      if (jp >= 0) {
        const double* Gq = w + L.G + q * NC * NW;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          const double dl = w[L.dlam + jp * NC + r];
#pragma unroll
          for (int i = 0; i < NW; ++i) v[i] -= Gq[r * NW + i] * dl;
        }
      }
      if (js >= 0) {
        const double* Fq = w + L.F + q * NC * NW;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          const double dl = w[L.dlam + js * NC + r];
#pragma unroll
          for (int i = 0; i < NW; ++i) v[i] -= Fq[r * NW + i] * dl;
        }
      }
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const bool fx = (fixmask[q] >> i) & 1u;
        const double r_ = fx ? 0.0 : v[i];
        w[L.dzL + q * NW + i] = r_;
        rmax = fmax(rmax, (r_ != r_) ? INFINITY : fabs(r_));
      }
    }
    MYR_SYNC();
    // stage residual rc = -c - (J dz - dc dlam), and the Schur right-hand side  J Hinv rz - rc
    for (int j = MYR_TID; j < St; j += MYR_NT) {
      double rc[NC], bj[NC];
#pragma unroll
      for (int r = 0; r < NC; ++r) { rc[r] = -w[L.c + j * NC + r] + delta_c * w[L.dlam + j * NC + r]; bj[r] = 0.0; smax = fmax(smax, fabs(w[L.c + j * NC + r])); }
      const int nk = S::stage_nodes(P, j);
      for (int k = 0; k < nk; ++k) {
        int role; const int q = S::stage_node(P, j, k, role);
        const double* Jq = w + (role ? L.F : L.G) + q * NC * NW;
        const double* Hi = w + L.Hinv + q * NW * NW;
        double t[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          double a = 0.0;
#pragma unroll
          for (int k2 = 0; k2 < NW; ++k2) a += Hi[i * NW + k2] * w[L.dzL + q * NW + k2];
          t[i] = a;
        }
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0, b_ = 0.0;
#pragma unroll
          for (int i = 0; i < NW; ++i) { a += Jq[r * NW + i] * w[L.dz + q * NW + i]; b_ += Jq[r * NW + i] * t[i]; }
          rc[r] -= a; bj[r] += b_;
        }
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        rmax = fmax(rmax, (rc[r] != rc[r]) ? INFINITY : fabs(rc[r]));
        bb[j * NC + r] = bj[r] - rc[r];
      }
    }
    rmax = block_max(rmax, red);
    smax = block_max(smax, red);
    MYR_SYNC();
    if (!(rmax > 1e-13 * fmax(1.0, smax)) || !isfinite(rmax)) break;
    block_cr_resolve<NC>(St, D, VL, VU, bb, w + L.dzU /* ddlam: St*NC <= Q*NW */);
    for (int j = MYR_TID; j < St; j += MYR_NT)
#pragma unroll
      for (int r = 0; r < NC; ++r) w[L.dlam + j * NC + r] += w[L.dzU + j * NC + r];
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      double v[NW];
#pragma unroll
      for (int i = 0; i < NW; ++i) v[i] = w[L.dzL + q * NW + i];
      const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
      if (jp >= 0) {
        const double* Gq = w + L.G + q * NC * NW;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          const double d = w[L.dzU + jp * NC + r];
#pragma unroll
          for (int i = 0; i < NW; ++i) v[i] -= Gq[r * NW + i] * d;
        }
      }
      if (js >= 0) {
        const double* Fq = w + L.F + q * NC * NW;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          const double d = w[L.dzU + js * NC + r];
#pragma unroll
          for (int i = 0; i < NW; ++i) v[i] -= Fq[r * NW + i] * d;
        }
      }
      const double* Hi = w + L.Hinv + q * NW * NW;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NW; ++k) a += Hi[i * NW + k] * v[k];
        w[L.dz + q * NW + i] += a;
      }
    }
    MYR_SYNC();
  }
  return ok;
}

// ------------------------------------------------------------------ K3: interior-point solve of one instance
struct IpmIO {
  const double* z0;   // [B][nvars] initial guess (reference layout)
  const double* lb;   // [B][nvars]
  const double* ub;   // [B][nvars]
  double* z;          // [B][nvars] out
  double* lam;        // [B][ncon] out (reference constraint order and sign convention)
  double* zL;         // [B][nvars] out (bound multipliers)
  double* zU;         // [B][nvars] out
  double* obj;        // [B]
  double* kkt_err;    // [B]  scaled optimality error E_0 at exit
  double* con_inf;    // [B]  max |c|
  int32_t* status;    // [B]
  int32_t* iters;     // [B]
  double* work;       // [B][work_stride]
  long long work_stride;
};

// per-instance pointers of the NLP the IPM iterates on (the reference NLP itself for collocation, the lifted one for
// shooting); nv / ncn are that NLP's sizes
struct InstPtrs {
  const double* z0; const double* lb; const double* ub;
  double* z; double* lam; double* zL; double* zU;
  double* w;
  int nv, ncn;
};
struct InstResult { double f, E0, cinf; int status, iters; };

template <class S>
MYR_HDI InstResult ipm_solve_instance(const Problem& P, const IpmOpts& O, const InstPtrs& ip, double* cr, double* red,
                                      double* sig_sh /*Q*NW doubles*/, uint32_t* fix_sh /*Q*/, double* mlp_scr = nullptr) {
  constexpr int NW = S::NW, NC = S::NC;
  const Layout<S> L(P);
  const int Q = L.Q, St = L.St;
  const int ncn = ip.ncn;
  double* w = ip.w;
  double* z = ip.z;
  double* lam = ip.lam;
  double* zL = ip.zL;
  double* zU = ip.zU;
  const double* lb = ip.lb;
  const double* ub = ip.ub;
  const double* z0 = ip.z0;

  // ---- initial point: push into the (relaxed) box, unit bound multipliers, zero lambda
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
    uint32_t fm = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int id = S::zidx(P, q, i);
      const Bnd bd = make_bnd(lb[id], ub[id], O.bound_relax);
      double x = z0[id];
      if (bd.fixed) { x = lb[id]; fm |= (1u << i); }
      else {
        if (bd.hasL) {
          double pL = O.bound_push * fmax(1.0, fabs(bd.lbr));
          if (bd.hasU) pL = fmin(pL, O.bound_frac * (bd.ubr - bd.lbr));
          x = fmax(x, bd.lbr + pL);
        }
        if (bd.hasU) {
          double pU = O.bound_push * fmax(1.0, fabs(bd.ubr));
          if (bd.hasL) pU = fmin(pU, O.bound_frac * (bd.ubr - bd.lbr));
          x = fmin(x, bd.ubr - pU);
        }
      }
      z[id] = x;
      zL[id] = bd.hasL ? 1.0 : 0.0;
      zU[id] = bd.hasU ? 1.0 : 0.0;
    }
    fix_sh[q] = fm;
  }
  for (int k = MYR_TID; k < ncn; k += MYR_NT) lam[k] = 0.0;
  MYR_SYNC();

  double mu = O.mu_init, nu = 1.0, delta_last = 0.0;
  int it = 0, status = ST_MAXITER, n_acceptable = 0;
  double f = 0.0, E0 = INFINITY, cinf = INFINITY, c1 = 0.0;
  const double mu_floor = fmax(O.mu_min, O.tol / 10.0);

  while (true) {
    // ---------------- K1: evaluate with derivatives
    f = eval_nodes<S, 2>(P, L, z, lam, w, red, mlp_scr);
    stage_constraints<S>(P, L, w, w + L.c, red, cinf, c1);
    // ---------------- dual residual, complementarity, scaling sums
    double rdmax = 0.0, szmax = -INFINITY, szmin = INFINITY, sumz = 0.0, nbnd = 0.0, slog = 0.0;
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
      const double* Gq = w + L.G + q * NC * NW;
      const double* Fq = w + L.F + q * NC * NW;
      double r[NW];
#pragma unroll
      for (int i = 0; i < NW; ++i) r[i] = w[L.gl + q * NW + i];
      if (jp >= 0) {
#pragma unroll
        for (int rr = 0; rr < NC; ++rr) {
          const double l = lam[S::cidx(P, jp, rr)];
#pragma unroll
          for (int i = 0; i < NW; ++i) r[i] += Gq[rr * NW + i] * l;
        }
      }
      if (js >= 0) {
#pragma unroll
        for (int rr = 0; rr < NC; ++rr) {
          const double l = lam[S::cidx(P, js, rr)];
#pragma unroll
          for (int i = 0; i < NW; ++i) r[i] += Fq[rr * NW + i] * l;
        }
      }
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const int id = S::zidx(P, q, i);
        const Bnd bd = make_bnd(lb[id], ub[id], O.bound_relax);
        double rd = 0.0;
        if (!bd.fixed) {
          rd = r[i] - zL[id] + zU[id];
          if (bd.hasL) { const double sl = z[id] - bd.lbr; const double pz = sl * zL[id]; szmax = fmax(szmax, pz); szmin = fmin(szmin, pz); sumz += zL[id]; nbnd += 1.0; slog += log(sl); }
          if (bd.hasU) { const double su = bd.ubr - z[id]; const double pz = su * zU[id]; szmax = fmax(szmax, pz); szmin = fmin(szmin, pz); sumz += zU[id]; nbnd += 1.0; slog += log(su); }
        }
        w[L.rb + q * NW + i] = bd.fixed ? 0.0 : r[i];  // grad f + J^T lam (barrier terms added after mu is known)
        const double a = (rd != rd) ? INFINITY : fabs(rd);
        rdmax = fmax(rdmax, a);
      }
    }
    double suml = 0.0;
    for (int k = MYR_TID; k < ncn; k += MYR_NT) suml += fabs(lam[k]);
    rdmax = block_max(rdmax, red);
    szmax = block_max(szmax, red);
    szmin = block_min(szmin, red);
    sumz = block_sum(sumz, red);
    nbnd = block_sum(nbnd, red);
    slog = block_sum(slog, red);
    suml = block_sum(suml, red);
    const double sd = fmax(O.s_max, (suml + sumz) / fmax(1.0, (double)ncn + nbnd)) / O.s_max;
    const double sc = nbnd > 0 ? fmax(O.s_max, sumz / nbnd) / O.s_max : 1.0;
    // lifted shooting: step defects accumulate along an interval's rollout, so the per-step feasibility tolerance is
    // tightened by cpi to keep the REFERENCE constraint px_k - x_{k+1} within tol
    const double cscale = scheme_is_lifted<S>::value ? 10.0 * (double)P.cpi : 1.0;
    auto Emu = [&](double m_, double cs = 1.0) {
      const double comp = nbnd > 0 ? fmax(szmax - m_, m_ - szmin) / sc : 0.0;
      return fmax(fmax(rdmax / sd, cinf * cs), comp);
    };
    E0 = Emu(0.0, cscale);  // the tightening applies to termination only, not to the barrier-parameter schedule
    if (!(E0 == E0) || !isfinite(f)) { status = ST_NAN; break; }
    if (E0 <= O.tol) { status = ST_SOLVED; break; }
    if (E0 <= O.acceptable_tol) { if (++n_acceptable >= O.acceptable_iter) { status = ST_ACCEPTABLE; break; } } else n_acceptable = 0;
    if (it >= O.max_iter) { status = ST_MAXITER; break; }
    while (mu > mu_floor && Emu(mu) <= O.kappa_eps * mu) mu = fmax(mu_floor, fmin(O.kappa_mu * mu, pow(mu, O.theta_mu)));
    const double tau = fmax(O.tau_min, 1.0 - mu);

    // ---------------- barrier gradient rb and Sigma
    double dphi_lin = 0.0;
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const int id = S::zidx(P, q, i);
        const Bnd bd = make_bnd(lb[id], ub[id], O.bound_relax);
        double sg = 0.0, rbv = w[L.rb + q * NW + i];
        if (bd.hasL) { const double sl = z[id] - bd.lbr; sg += zL[id] / sl; rbv -= mu / sl; }
        if (bd.hasU) { const double su = bd.ubr - z[id]; sg += zU[id] / su; rbv += mu / su; }
        sig_sh[q * NW + i] = sg;
        w[L.rb + q * NW + i] = bd.fixed ? 0.0 : rbv;
      }
    }
    (void)dphi_lin;
    MYR_SYNC();

    // ---------------- K2: KKT solve with inertia correction (IPOPT Algorithm IC)
    double delta = 0.0;
    bool ok = false;
    for (int tries = 0; tries < 60; ++tries) {
      // accurate steps only matter near the solution: refine the linear solve in the end game only
      ok = kkt_solve<S>(P, L, w, sig_sh, fix_sh, delta, O.delta_c, cr, red, O.delta_reg, E0 < 1e-3 ? O.max_refine : 0);
      if (ok) break;
      if (delta == 0.0) delta = (delta_last == 0.0) ? O.delta_0 : fmax(O.delta_min, O.kappa_w_minus * delta_last);
      else delta *= (delta_last == 0.0) ? O.kappa_w_plus_first : O.kappa_w_plus;
      if (delta > O.delta_max) break;
    }
    if (!ok) { status = ST_INERTIA; break; }
    if (delta > 0.0) delta_last = delta;

    // ---------------- step sizes (fraction to the boundary), bound-multiplier steps, merit derivative
    double a_pr = 1.0, a_du = 1.0, dphi = 0.0, dHd = 0.0;
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
      const double* Gq = w + L.G + q * NC * NW;
      const double* Fq = w + L.F + q * NC * NW;
      // J^T (lam + dlam) part of  H dz = -(rb + J^T dlam): dz^T H dz = -dz.(rb + J^T dlam)
      double jt[NW], jl[NW];
#pragma unroll
      for (int i = 0; i < NW; ++i) { jt[i] = 0.0; jl[i] = 0.0; }
      if (jp >= 0) {
#pragma unroll
        for (int rr = 0; rr < NC; ++rr) {
          const double d = w[L.dlam + jp * NC + rr], l = lam[S::cidx(P, jp, rr)];
#pragma unroll
          for (int i = 0; i < NW; ++i) { jt[i] += Gq[rr * NW + i] * d; jl[i] += Gq[rr * NW + i] * l; }
        }
      }
      if (js >= 0) {
#pragma unroll
        for (int rr = 0; rr < NC; ++rr) {
          const double d = w[L.dlam + js * NC + rr], l = lam[S::cidx(P, js, rr)];
#pragma unroll
          for (int i = 0; i < NW; ++i) { jt[i] += Fq[rr * NW + i] * d; jl[i] += Fq[rr * NW + i] * l; }
        }
      }
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const int id = S::zidx(P, q, i);
        const Bnd bd = make_bnd(lb[id], ub[id], O.bound_relax);
        const double d = w[L.dz + q * NW + i];
        double dl = 0.0, du = 0.0;
        if (!bd.fixed) {
          const double rbv = w[L.rb + q * NW + i];
          dHd -= d * (rbv + jt[i]);
          dphi += d * (rbv - jl[i]);  // barrier-objective gradient = rb - J^T lam
          if (bd.hasL) {
            const double sl = z[id] - bd.lbr;
            dl = mu / sl - zL[id] - zL[id] / sl * d;
            if (d < 0.0) a_pr = fmin(a_pr, -tau * sl / d);
            if (dl < 0.0) a_du = fmin(a_du, -tau * zL[id] / dl);
          }
          if (bd.hasU) {
            const double su = bd.ubr - z[id];
            du = mu / su - zU[id] + zU[id] / su * d;
            if (d > 0.0) a_pr = fmin(a_pr, tau * su / d);
            if (du < 0.0) a_du = fmin(a_du, -tau * zU[id] / du);
          }
        }
        w[L.dzL + q * NW + i] = dl;
        w[L.dzU + q * NW + i] = du;
      }
    }
    a_pr = block_min(a_pr, red);
    a_du = block_min(a_du, red);
    dphi = block_sum(dphi, red);
    dHd = block_sum(dHd, red);
    if (!(dphi == dphi)) { status = ST_NAN; break; }

    // ---------------- l1-merit backtracking line search
    if (c1 > 0.0) {
      const double nu_trial = (dphi + 0.5 * fmax(dHd, 0.0)) / ((1.0 - O.rho) * c1);
      if (nu < nu_trial) nu = nu_trial + 1.0;
    }
    const double Dm = dphi - nu * c1;
    const double phi0 = f - mu * slog + nu * c1;
    double alpha = a_pr;
    bool accepted = false;
    double* zt = w + L.Hinv;  // Hinv is dead after kkt_solve: reuse as the trial point (reference layout needs nv <= Q*NW*NW)
    double f_t = 0.0;
    for (int ls = 0; ls < O.max_ls; ++ls) {
      double blog = 0.0;
      for (int q = MYR_TID; q < Q; q += MYR_NT) {
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const int id = S::zidx(P, q, i);
          const Bnd bd = make_bnd(lb[id], ub[id], O.bound_relax);
          const double x = z[id] + alpha * w[L.dz + q * NW + i];
          zt[id] = x;
          if (bd.hasL) blog += log(x - bd.lbr);
          if (bd.hasU) blog += log(bd.ubr - x);
        }
      }
      MYR_SYNC();
      f_t = eval_nodes<S, 0>(P, L, zt, lam, w, red, mlp_scr);
      double ci_t, c1_t;
      stage_constraints<S>(P, L, w, w + L.crb /*scratch*/, red, ci_t, c1_t);
      blog = block_sum(blog, red);
      const double phit = f_t - mu * blog + nu * c1_t;
      // Armijo with IPOPT's rounding-error relaxation (10 eps |phi|) so that converged iterates are not rejected by cancellation
      if (isfinite(phit) && phit <= phi0 + O.eta * alpha * Dm + 10.0 * 2.220446049250313e-16 * fabs(phi0)) { accepted = true; break; }
      alpha *= 0.5;
    }
    if (!accepted) { status = (E0 <= O.acceptable_tol) ? ST_ACCEPTABLE : ST_LINESEARCH; break; }

    // ---------------- accept: primal, equality multipliers, bound multipliers (with the kappa_sigma safeguard)
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const int id = S::zidx(P, q, i);
        const Bnd bd = make_bnd(lb[id], ub[id], O.bound_relax);
        if (bd.fixed) continue;
        const double x = zt[id];
        z[id] = x;
        if (bd.hasL) {
          const double sl = x - bd.lbr;
          double v = zL[id] + a_du * w[L.dzL + q * NW + i];
          v = fmax(fmin(v, O.kappa_sigma * mu / sl), mu / (O.kappa_sigma * sl));
          zL[id] = v;
        }
        if (bd.hasU) {
          const double su = bd.ubr - x;
          double v = zU[id] + a_du * w[L.dzU + q * NW + i];
          v = fmax(fmin(v, O.kappa_sigma * mu / su), mu / (O.kappa_sigma * su));
          zU[id] = v;
        }
      }
    }
    for (int j = MYR_TID; j < St; j += MYR_NT)
#pragma unroll
      for (int r = 0; r < NC; ++r) lam[S::cidx(P, j, r)] += alpha * w[L.dlam + j * NC + r];
    MYR_SYNC();
    ++it;
  }
  InstResult res;
  res.f = f; res.E0 = E0; res.cinf = cinf; res.status = status; res.iters = it;
  return res;
}

// Collocation: the IPM works directly on the caller's arrays.
template <class S>
MYR_HDI typename std::enable_if<!scheme_is_lifted<S>::value>::type
ipm_solve_entry(const Problem& P, const IpmOpts& O, const IpmIO& io, int b, double* cr, double* red, double* sig_sh, uint32_t* fix_sh,
                double* mlp_scr = nullptr) {
  const long long nv = P.nvars, nc = P.ncon;
  InstPtrs ip{io.z0 + b * nv, io.lb + b * nv, io.ub + b * nv, io.z + b * nv, io.lam + b * nc, io.zL + b * nv, io.zU + b * nv,
              io.work + (long long)b * io.work_stride, P.nvars, P.ncon};
  const InstResult r = ipm_solve_instance<S>(P, O, ip, cr, red, sig_sh, fix_sh, mlp_scr);
  if (MYR_TID == 0) {
    io.obj[b] = r.f; io.kkt_err[b] = r.E0; io.con_inf[b] = r.cinf; io.status[b] = r.status; io.iters[b] = r.iters;
  }
}

// Lifted shooting: expand the reference NLP data into the lifted NLP, solve, map the solution back and report the
// objective / constraint violation of the REFERENCE NLP (rollouts from the interval-start states).
template <class S>
MYR_HDI typename std::enable_if<scheme_is_lifted<S>::value>::type
ipm_solve_entry(const Problem& P, const IpmOpts& O, const IpmIO& io, int b, double* cr, double* red, double* sig_sh, uint32_t* fix_sh,
                double* = nullptr) {
  constexpr int n = S::n, m = S::m, NW = S::NW, NC = S::NC;
  const Layout<S> L(P);
  const int Q = L.Q, St = L.St;
  const int nvI = Q * NW, ncI = St * NC;
  const long long nv = P.nvars, nc = P.ncon;
  double* w = io.work + (long long)b * io.work_stride;
  double* zI0 = w + L.ext; double* lbI = zI0 + nvI; double* ubI = lbI + nvI; double* zI = ubI + nvI;
  double* zLI = zI + nvI; double* zUI = zLI + nvI; double* lamI = zUI + nvI;
  const double* z0 = io.z0 + b * nv; const double* lb = io.lb + b * nv; const double* ub = io.ub + b * nv;
  // ---- expansion: real variables copy guess and bounds; copies / hidden states are free
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int r = S::ref_index(P, q, i);
      double g0 = 0.0, l0 = -INFINITY, u0 = INFINITY;
      if (r >= 0) { g0 = z0[r]; l0 = lb[r]; u0 = ub[r]; }
      else if (i >= n && q + 1 < Q) {  // control copy: value of the next node's first control
        const int r2 = S::ref_index(P, q + 1, n + (i - n) % m);
        g0 = z0[r2];
      }
      if (S::is_dead(P, q, i)) { l0 = g0; u0 = g0; }
      zI0[q * NW + i] = g0; lbI[q * NW + i] = l0; ubI[q * NW + i] = u0;
    }
  }
  MYR_SYNC();
  // hidden states: roll the guess out inside every interval (what the reference's first constraint evaluation does)
  for (int k = MYR_TID; k < P.N; k += MYR_NT) {
    double s[n];
#pragma unroll
    for (int i = 0; i < n; ++i) s[i] = zI0[(k * P.cpi) * NW + i];
    for (int j = 0; j + 1 < P.cpi; ++j) {
      const int q = k * P.cpi + j;
      double v[NW], ell, gl[1], phi[NC], psi[NC], Gm[1], Fm[1], Wd[1], l0[NC];
#pragma unroll
      for (int i = 0; i < n; ++i) v[i] = s[i];
#pragma unroll
      for (int i = n; i < NW; ++i) v[i] = zI0[q * NW + i];
#pragma unroll
      for (int r = 0; r < NC; ++r) l0[r] = 0.0;
      S::template eval_node<0>(P, q, v, l0, l0, ell, gl, phi, psi, Gm, Fm, Wd);
#pragma unroll
      for (int i = 0; i < n; ++i) { s[i] = phi[i]; zI0[(q + 1) * NW + i] = phi[i]; }
    }
  }
  MYR_SYNC();
  Problem Pi = P;
  Pi.nvars = nvI; Pi.ncon = ncI;
  InstPtrs ip{zI0, lbI, ubI, zI, lamI, zLI, zUI, w, nvI, ncI};
  InstResult r = ipm_solve_instance<S>(Pi, O, ip, cr, red, sig_sh, fix_sh);
  MYR_SYNC();
  // ---- map back
  double* z = io.z + b * nv; double* lam = io.lam + b * nc; double* zL = io.zL + b * nv; double* zU = io.zU + b * nv;
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int rr = S::ref_index(P, q, i);
      if (rr >= 0) { z[rr] = zI[q * NW + i]; zL[rr] = zLI[q * NW + i]; zU[rr] = zUI[q * NW + i]; }
    }
  }
  for (int k = MYR_TID; k < P.N; k += MYR_NT)
#pragma unroll
    for (int i = 0; i < n; ++i) lam[k * n + i] = lamI[((k + 1) * P.cpi - 1) * NC + i];
  MYR_SYNC();
  // objective and constraint violation as the reference defines them (shooting.py:169-241)
  double fs = 0.0, cm = 0.0;
  for (int k = MYR_TID; k < P.N; k += MYR_NT) {
    double px[n], cst;
    ShootingInterval<typename S::System, (NW - n) / m>::template run<false>(P, k, z, px, cst, nullptr, 0);
    fs += cst;
#pragma unroll
    for (int i = 0; i < n; ++i) { const double d = px[i] - z[(k + 1) * n + i]; cm = fmax(cm, (d != d) ? INFINITY : fabs(d)); }
  }
  fs = block_sum(fs, red);
  cm = block_max(cm, red);
  if (MYR_TID == 0) {
    io.obj[b] = fs; io.kkt_err[b] = r.E0; io.con_inf[b] = cm; io.status[b] = r.status; io.iters[b] = r.iters;
  }
}

}  // namespace myr
