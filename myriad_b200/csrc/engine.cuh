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
  int max_soc;        // second-order correction attempts when the first trial step is rejected (IPOPT max_soc = 4)
};

enum Status : int { ST_SOLVED = 0, ST_ACCEPTABLE = 1, ST_MAXITER = -1, ST_LINESEARCH = -2, ST_INERTIA = -3, ST_NAN = -13 };

// schemes whose internal NLP differs from the reference's (lifted shooting) declare kLifted
template <class S, class = void>
struct scheme_is_lifted { static constexpr bool value = false; };
template <class S>
struct scheme_is_lifted<S, typename std::enable_if<S::kLifted>::type> { static constexpr bool value = true; };

// ------------------------------------------------------------------ optional per-phase cycle accounting (debug builds)
// -DMYR_PROFILE_PHASES: thread 0 of every CTA accumulates clock64() deltas per phase into g_phase_cycles (read back
// with myr_debug_phase_cycles); compiled out of the product build.
#if defined(MYR_PROFILE_PHASES) && defined(__CUDA_ARCH__)
#define MYR_PH_DECL long long ph_t0_ = clock64();
#define MYR_PH(idx) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_cycles[idx], (unsigned long long)(t_ - ph_t0_)); ph_t0_ = t_; } } while (0)
#else
#define MYR_PH_DECL
#define MYR_PH(idx) do { } while (0)
#endif

// ------------------------------------------------------------------ workspace layout (doubles / instance)
template <class S>
struct Layout {
  // collocation on a NODE system: the MLP is evaluated for all nodes of the instance cooperatively (node_mlp.cuh)
  static constexpr bool kCoopMlp = sys_is_node<typename S::System>::value && !scheme_is_lifted<S>::value;
  int Q, St;
  int ldq, lds;   // leading dimensions of the element-major node / stage arrays
  int G, F, W, Hinv, gl, phi, psi, rb, dz, dzL, dzU, c, dlam;
  int crD, crU, crVL, crVU, crb, crx;
  int lbr, ubr, rsl, rsu, zt, dz2, sch, ct, csoc, dl2;
  int dynf, dynJ, dynH;   // NODE systems: per-node MLP dynamics values / Jacobians / contracted Hessians
  int ext;
  int total;
  MYR_HDI explicit Layout(const Problem& P) {
    Q = S::num_nodes(P); St = S::num_stages(P);
    ldq = (Q + 3) & ~3;
    lds = ((St + 3) & ~3) + 1;  // = 1 mod 4: the NC rows of a block (stride NC * lds doubles) fall into distinct shared-memory banks
    int o = 0;
    G = o; o += ldq * S::NC * S::NW;
    F = o; o += ldq * S::NC * S::NW;
    W = o; o += ldq * S::NWP;
    Hinv = o; o += ldq * S::NW * S::NW;
    gl = o; o += ldq * S::NW;
    phi = o; o += ldq * S::NC;
    psi = o; o += ldq * S::NC;
    rb = o; o += ldq * S::NW;
    dz = o; o += ldq * S::NW;
    dzL = o; o += ldq * S::NW;
    dzU = o; o += ldq * S::NW;
    c = o; o += lds * S::NC;
    dlam = o; o += lds * S::NC;
    crD = o; o += lds * S::NC * S::NC;
    crU = o; o += lds * S::NC * S::NC;
    crVL = o; o += lds * S::NC * S::NC;
    crVU = o; o += lds * S::NC * S::NC;
    crb = o; o += lds * S::NC;
    crx = dlam;  // CR writes its solution straight into dlam
    lbr = o; o += ldq * S::NW;   // relaxed bounds per variable (-inf / +inf: none)
    ubr = o; o += ldq * S::NW;
    zt = o; o += ldq * S::NW;    // trial point of the line search (reference layout, nv <= Q * NW)
    dz2 = o; o += ldq * S::NW;   // second-order-correction step
    sch = o; o += lds * S::NC;   // -J Hinv rb per stage (kept from the last KKT solve for re-solves)
    ct = o; o += lds * S::NC;    // constraints at the trial point
    csoc = o; o += lds * S::NC;  // accumulated SOC right-hand side
    dl2 = o; o += lds * S::NC;   // multiplier step of the SOC solve
    rsl = o; o += ldq * S::NW;   // reciprocal slacks 1 / (x - lbr), 1 / (ubr - x) of the current iterate
    rsu = o; o += ldq * S::NW;
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
  MYR_HDI int cr_doubles() const { return lds * (4 * S::NC * S::NC + S::NC); }
};

// Workspace arrays are element-major ("structure of arrays"): element e of node q lives at  base + e * ldq + q  (stage
// arrays: base + e * lds + j), so the threads of a warp -- one node / stage each -- touch consecutive addresses.
#define NQ(arr, q, e) w[L.arr + (e) * L.ldq + (q)]
#define NS(arr, j, e) w[L.arr + (e) * L.lds + (j)]
template <class T>
struct Strided {
  T* p; int s;
  MYR_HDI T& operator[](int i) const { return p[(long long)i * s]; }
};
using SV = Strided<double>;
using CSV = Strided<const double>;

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

// sum of log(slack) accumulated as log(product of slacks): one log per node instead of one per bound.  The product
// is flushed whenever it leaves [1e-200, 1e200]; a non-positive slack poisons the result with NaN like log() would.
struct LogProd {
  double prod = 1.0, sum = 0.0;
  bool bad = false;
  MYR_HDI void mul(double s) {
    bad = bad || !(s > 0.0);
    prod *= s;
    if (!(prod > 1e-200 && prod < 1e200)) { sum += log(prod); prod = 1.0; }
  }
  MYR_HDI double value() const { return bad ? NAN : sum + log(prod); }
};

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
    for (int r = 0; r < S::NC; ++r) { NQ(phi, q, r) = phi[r]; NQ(psi, q, r) = psi[r]; }
    if (MODE >= 1) {
#pragma unroll
      for (int i = 0; i < S::NW; ++i) NQ(gl, q, i) = gl[i];
#pragma unroll
      for (int i = 0; i < S::NC * S::NW; ++i) { NQ(G, q, i) = G[i]; NQ(F, q, i) = F[i]; }
    }
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < S::NWP; ++i) NQ(W, q, i) = W[i];
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
        a += role ? NQ(psi, q, r) : NQ(phi, q, r);
      }
      cdst[r * L.lds + j] = a;
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
template <int R, int K, int C, class TA, class TB>
MYR_HDI void mm(const TA& A, const TB& Bm, double* Cm) {  // C = A(RxK) B(KxC)
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
// Block cyclic reduction for the symmetric block-tridiagonal system
//   U_{i-1}^T x_{i-1} + D_i x_i + U_i x_{i+1} = b_i,   i = 0..St-1,  blocks NC x NC (D symmetric, may be indefinite).
// Factor and solve are fused (single right-hand side).  Pivot-block inertias are accumulated: by Sylvester's
// law their sum is the inertia of the whole matrix.  D,U,b are destroyed; x receives the solution.
// All block arrays are element-major with leading dimension ld: entry k of block i is at  A[k * ld + i].
//
// Work decomposition: a GROUP of G = pow2 >= NC adjacent lanes owns one block; lane r of the group produces ROW r of
// every block product of that block (the pivot-block inverse is computed redundantly by the lanes of the group: it is
// a serial chain anyway).  With one thread per block the cost of a level does not shrink with the number of blocks
// left, and the deep levels (13, 6, 3, 2, 1 blocks) dominate; with row-parallel groups a level costs roughly the
// latency of one inverse plus a few row products.  On the host (one "thread") a group degenerates to a loop over rows.
template <int NC>
struct CrGroup { static constexpr int G = NC <= 1 ? 1 : (NC <= 2 ? 2 : (NC <= 4 ? 4 : (NC <= 8 ? 8 : 16))); };

#ifdef __CUDA_ARCH__
#define MYR_CR_ROWS(r) for (int r = int(threadIdx.x) % G; r < NC; r += G)
#else
#define MYR_CR_ROWS(r) for (int r = 0; r < NC; ++r)
#endif

template <int NC>
MYR_HDI void block_cr_solve(int St, int ld, double* D, double* U, double* VL, double* VU, double* b, double* x,
                            double* red, int& npos, int& nneg, int& nzero) {
  constexpr int BB = NC * NC;
  constexpr int G = CrGroup<NC>::G;
#ifdef __CUDA_ARCH__
  const int grp = int(threadIdx.x) / G, ngrp = int(blockDim.x) / G;
  const bool counter = (int(threadIdx.x) % G) == 0;
#else
  const int grp = 0, ngrp = 1;
  const bool counter = true;
#endif
  int cp = 0, cn = 0, cz = 0;
#if defined(MYR_PROFILE_PHASES) && defined(__CUDA_ARCH__)
  long long cph_t0_ = clock64();
#define MYR_CPH(idx) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_cycles[idx], (unsigned long long)(t_ - cph_t0_)); cph_t0_ = t_; } } while (0)
#else
#define MYR_CPH(idx) do { } while (0)
#endif
  // pivot inverse of block i (symmetrised on load), row r selected into dr without dynamic register indexing
  auto pivot_rows = [&](int i, double* Dinv) {
    double A[BB];
#pragma unroll
    for (int r = 0; r < NC; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c) A[r * NC + c] = 0.5 * (D[(r * NC + c) * ld + i] + D[(c * NC + r) * ld + i]);
    int p_, n_, z_;
    sym_inverse<NC>(A, 0u, Dinv, p_, n_, z_);
    if (counter) { cp += p_; cn += n_; cz += z_; }
  };
  int s = 1;
  for (; s < St; s <<= 1) {
    const int nodd = (St - s + 2 * s - 1) / (2 * s);   // blocks i = (2k+1) s < St
    const int neven = (St + 2 * s - 1) / (2 * s);      // blocks i = 2k s < St
    // ---- eliminate the odd blocks: Dinv_i (kept in D), VL_i = Dinv_i U_{i-s}^T, VU_i = Dinv_i U_i, x_i = Dinv_i b_i
    for (int k0 = 0; k0 < nodd; k0 += ngrp) {   // trip count uniform across the CTA (the loop contains a warp barrier)
      const int k = k0 + grp;
      const bool active = k < nodd;
      const int i = active ? (2 * k + 1) * s : s;
      double Dinv[BB], Ul[BB], Ui[BB], bi[NC];
      const bool hr = i + s < St;
      if (active) {
        pivot_rows(i, Dinv);
#pragma unroll
        for (int e = 0; e < BB; ++e) { Ul[e] = U[e * ld + (i - s)]; Ui[e] = hr ? U[e * ld + i] : 0.0; }
#pragma unroll
        for (int m = 0; m < NC; ++m) bi[m] = b[m * ld + i];
      }
#ifdef __CUDA_ARCH__
      __syncwarp();  // every lane of the group has read D_i before rows of it are overwritten
#endif
      if (active) {
        MYR_CR_ROWS(r) {
          double dr[NC];
#pragma unroll
          for (int m = 0; m < NC; ++m) {
            double v = 0.0;
#pragma unroll
            for (int rr = 0; rr < NC; ++rr) v = (rr == r) ? Dinv[rr * NC + m] : v;
            dr[m] = v;
          }
          double xr = 0.0;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            double vl = 0.0, vu = 0.0;
#pragma unroll
            for (int m = 0; m < NC; ++m) { vl += dr[m] * Ul[c * NC + m]; vu += dr[m] * Ui[m * NC + c]; }
            VL[(r * NC + c) * ld + i] = vl;
            if (hr) VU[(r * NC + c) * ld + i] = vu;
            D[(r * NC + c) * ld + i] = dr[c];
            xr += dr[c] * bi[c];
          }
          x[r * ld + i] = xr;
        }
      }
    }
    MYR_SYNC();
    MYR_CPH(s == 1 ? 10 : (s == 2 ? 12 : 14));
    // ---- update the even blocks i from their eliminated neighbours er = i + s, el = i - s (row r per lane, in place)
    for (int k = grp; k < neven; k += ngrp) {
      const int i = 2 * k * s, er = i + s, el = i - s;
      const bool hr = er < St, hl = el >= 0, hrr = hr && er + s < St;
      MYR_CR_ROWS(r) {
        double dn[NC], un[NC], ur[NC], uc[NC];
        double bn = b[r * ld + i];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          dn[c] = D[(r * NC + c) * ld + i]; un[c] = 0.0;
          ur[c] = hr ? U[(r * NC + c) * ld + i] : 0.0;      // row r of U_i
          uc[c] = hl ? U[(c * NC + r) * ld + el] : 0.0;     // column r of U_el = row r of U_el^T
        }
        if (hr) {
#pragma unroll
          for (int m = 0; m < NC; ++m) {
            const double u_ = ur[m];
#pragma unroll
            for (int c = 0; c < NC; ++c) dn[c] -= u_ * VL[(m * NC + c) * ld + er];
            bn -= u_ * x[m * ld + er];
          }
          if (hrr) {
#pragma unroll
            for (int m = 0; m < NC; ++m) {
              const double u_ = ur[m];
#pragma unroll
              for (int c = 0; c < NC; ++c) un[c] -= u_ * VU[(m * NC + c) * ld + er];
            }
          }
        }
        if (hl) {
#pragma unroll
          for (int m = 0; m < NC; ++m) {
            const double u_ = uc[m];
#pragma unroll
            for (int c = 0; c < NC; ++c) dn[c] -= u_ * VU[(m * NC + c) * ld + el];
            bn -= u_ * x[m * ld + el];
          }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) { D[(r * NC + c) * ld + i] = dn[c]; U[(r * NC + c) * ld + i] = un[c]; }
        b[r * ld + i] = bn;
      }
    }
    MYR_SYNC();
    MYR_CPH(s == 1 ? 11 : (s == 2 ? 13 : 15));
  }
  // ---- root (block 0): loads and inverse, CTA barrier, then the rows (the barrier separates reads of D_0 from writes)
  {
    double Dinv[BB], bi[NC];
    if (grp == 0) {
      pivot_rows(0, Dinv);
#pragma unroll
      for (int m = 0; m < NC; ++m) bi[m] = b[m * ld];
    }
    MYR_SYNC();
    if (grp == 0) {
      MYR_CR_ROWS(r) {
        double xr = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          double v = 0.0;
#pragma unroll
          for (int rr = 0; rr < NC; ++rr) v = (rr == r) ? Dinv[rr * NC + c] : v;
          D[(r * NC + c) * ld] = v;
          xr += v * bi[c];
        }
        x[r * ld] = xr;
      }
    }
  }
  MYR_SYNC();
  // ---- back substitution: x_i = (Dinv_i b_i) - VL_i x_{i-s} - VU_i x_{i+s}
  for (s >>= 1; s >= 1; s >>= 1) {
    const int nodd = (St - s + 2 * s - 1) / (2 * s);
    for (int k = grp; k < nodd; k += ngrp) {
      const int i = (2 * k + 1) * s;
      MYR_CR_ROWS(r) {
        double a = x[r * ld + i];
#pragma unroll
        for (int m = 0; m < NC; ++m) a -= VL[(r * NC + m) * ld + i] * x[m * ld + (i - s)];
        if (i + s < St) {
#pragma unroll
          for (int m = 0; m < NC; ++m) a -= VU[(r * NC + m) * ld + i] * x[m * ld + (i + s)];
        }
        x[r * ld + i] = a;
      }
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
MYR_HDN void block_cr_resolve(int St, int ld, const double* D, const double* VL, const double* VU, double* b, double* x) {
  int s = 1;
  for (; s < St; s <<= 1) {
    for (int i = 2 * MYR_TID * s; i < St; i += 2 * MYR_NT * s) {
      double bn[NC];
#pragma unroll
      for (int r = 0; r < NC; ++r) bn[r] = b[r * ld + i];
      const int er = i + s, el = i - s;
      if (er < St) {
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NC; ++k) a += VL[(k * NC + r) * ld + er] * b[k * ld + er];
          bn[r] -= a;
        }
      }
      if (el >= 0) {
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NC; ++k) a += VU[(k * NC + r) * ld + el] * b[k * ld + el];
          bn[r] -= a;
        }
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) b[r * ld + i] = bn[r];
    }
    MYR_SYNC();
  }
  if (MYR_TID == 0) {
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NC; ++k) a += D[(r * NC + k) * ld] * b[k * ld];
      x[r * ld] = a;
    }
  }
  MYR_SYNC();
  for (s >>= 1; s >= 1; s >>= 1) {
    for (int i = (2 * MYR_TID + 1) * s; i < St; i += 2 * MYR_NT * s) {
      double xi[NC];
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) a += D[(r * NC + k) * ld + i] * b[k * ld + i];
        xi[r] = a;
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NC; ++k) a += VL[(r * NC + k) * ld + i] * x[k * ld + (i - s)];
        xi[r] -= a;
      }
      if (i + s < St) {
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NC; ++k) a += VU[(r * NC + k) * ld + i] * x[k * ld + (i + s)];
          xi[r] -= a;
        }
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) x[r * ld + i] = xi[r];
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
#if defined(MYR_PROFILE_PHASES) && defined(__CUDA_ARCH__)
  long long kph_t0_ = clock64();
#define MYR_KPH(idx) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_cycles[idx], (unsigned long long)(t_ - kph_t0_)); kph_t0_ = t_; } } while (0)
#else
#define MYR_KPH(idx) do { } while (0)
#endif
  // ---- node blocks: Hinv = (W + Sigma + dw)^-1 with fixed variables removed; inertia of H
  int hp = 0, hn = 0, hz = 0;
  double minpr = INFINITY;
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
    double A[NW * NW], inv[NW * NW];
    const CSV Wq{w + L.W + q, L.ldq};
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
    for (int i = 0; i < NW * NW; ++i) NQ(Hinv, q, i) = inv[i];
  }
  const int Hneg = (int)(block_sum((double)hn, red) + 0.5);
  const int Hzero = (int)(block_sum((double)hz, red) + 0.5);
  minpr = block_min(minpr, red);
  // refinement is only worth its cost when some node block was close to singular
  if (!(minpr < 1e-4)) max_refine = 0;
  (void)hp;
  MYR_SYNC();
  MYR_KPH(6);
  // ---- stage blocks of the Schur complement S = J Hinv J^T + dc I and its right-hand side  c - J Hinv rb
  const int ld = L.lds;
  double* D = cr; double* U = D + ld * NC * NC; double* VL = U + ld * NC * NC; double* VU = VL + ld * NC * NC;
  double* bb = VU + ld * NC * NC;
  for (int j = MYR_TID; j < St; j += MYR_NT) {
    double Dj[NC * NC], bj[NC];
#pragma unroll
    for (int i = 0; i < NC * NC; ++i) Dj[i] = 0.0;
#pragma unroll
    for (int r = 0; r < NC; ++r) { Dj[r * NC + r] = delta_c; bj[r] = NS(c, j, r); }
    const int nk = S::stage_nodes(P, j);
    for (int k = 0; k < nk; ++k) {
      int role; const int q = S::stage_node(P, j, k, role);
      double Jq[NC * NW], Hi[NW * NW], T[NC * NW];
      {
        const CSV Jv{w + (role ? L.F : L.G) + q, L.ldq};
        const CSV Hv{w + L.Hinv + q, L.ldq};
#pragma unroll
        for (int i = 0; i < NC * NW; ++i) Jq[i] = Jv[i];
#pragma unroll
        for (int i = 0; i < NW * NW; ++i) Hi[i] = Hv[i];
      }
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
        for (int i = 0; i < NW; ++i) a += T[r * NW + i] * NQ(rb, q, i);
        bj[r] -= a;
      }
      if (role == 1 && j + 1 < St) {  // link node: coupling to the next stage  U_j = F Hinv G^T
        const CSV Gq{w + L.G + q, L.ldq};
#pragma unroll
        for (int r = 0; r < NC; ++r)
#pragma unroll
          for (int c2 = 0; c2 < NC; ++c2) {
            double a = 0.0;
#pragma unroll
            for (int i = 0; i < NW; ++i) a += T[r * NW + i] * Gq[c2 * NW + i];
            U[(r * NC + c2) * ld + j] = a;
          }
      }
    }
#pragma unroll
    for (int r = 0; r < NC; ++r)
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) D[(r * NC + c2) * ld + j] = 0.5 * (Dj[r * NC + c2] + Dj[c2 * NC + r]);
#pragma unroll
    for (int r = 0; r < NC; ++r) { bb[r * ld + j] = bj[r]; NS(sch, j, r) = bj[r] - NS(c, j, r); }
  }
  MYR_SYNC();
  MYR_KPH(7);
  int sp, sn, sz;
  block_cr_solve<NC>(St, ld, D, U, VL, VU, bb, w + L.dlam, red, sp, sn, sz);
  MYR_KPH(8);
  // inertia(K) = inertia(H) + inertia(-S): correct iff  n-(S) == n-(H)  and nothing is singular
  const bool ok = (Hzero == 0) && (sz == 0) && (sn == Hneg);
  // ---- dz = -Hinv (rb + G^T dlam_phi + F^T dlam_psi)
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
    double v[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) v[i] = NQ(rb, q, i);
    const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
    if (jp >= 0) {
      const CSV Gq{w + L.G + q, L.ldq};
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        const double d = NS(dlam, jp, r);
#pragma unroll
        for (int i = 0; i < NW; ++i) v[i] += Gq[r * NW + i] * d;
      }
    }
    if (js >= 0) {
      const CSV Fq{w + L.F + q, L.ldq};
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        const double d = NS(dlam, js, r);
#pragma unroll
        for (int i = 0; i < NW; ++i) v[i] += Fq[r * NW + i] * d;
      }
    }
    const CSV Hi{w + L.Hinv + q, L.ldq};
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NW; ++k) a += Hi[i * NW + k] * v[k];
      NQ(dz, q, i) = -a;
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
      for (int i = 0; i < NW; ++i) { v[i] = -NQ(rb, q, i); d[i] = NQ(dz, q, i); smax = fmax(smax, fabs(v[i])); }
      const CSV Wq{w + L.W + q, L.ldq};
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        double a = (sigma[q * NW + i] + delta_w) * d[i];
#pragma unroll
        for (int k = 0; k < NW; ++k) a += Wq[pidx(i, k, NW)] * d[k];
        v[i] -= a;
      }
      const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
      if (jp >= 0) {
        const CSV Gq{w + L.G + q, L.ldq};
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          const double dl = NS(dlam, jp, r);
#pragma unroll
          for (int i = 0; i < NW; ++i) v[i] -= Gq[r * NW + i] * dl;
        }
      }
      if (js >= 0) {
        const CSV Fq{w + L.F + q, L.ldq};
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          const double dl = NS(dlam, js, r);
#pragma unroll
          for (int i = 0; i < NW; ++i) v[i] -= Fq[r * NW + i] * dl;
        }
      }
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const bool fx = (fixmask[q] >> i) & 1u;
        const double r_ = fx ? 0.0 : v[i];
        NQ(dzL, q, i) = r_;
        rmax = fmax(rmax, (r_ != r_) ? INFINITY : fabs(r_));
      }
    }
    MYR_SYNC();
    // stage residual rc = -c - (J dz - dc dlam), and the Schur right-hand side  J Hinv rz - rc
    for (int j = MYR_TID; j < St; j += MYR_NT) {
      double rc[NC], bj[NC];
#pragma unroll
      for (int r = 0; r < NC; ++r) { rc[r] = -NS(c, j, r) + delta_c * NS(dlam, j, r); bj[r] = 0.0; smax = fmax(smax, fabs(NS(c, j, r))); }
      const int nk = S::stage_nodes(P, j);
      for (int k = 0; k < nk; ++k) {
        int role; const int q = S::stage_node(P, j, k, role);
        const CSV Jq{w + (role ? L.F : L.G) + q, L.ldq};
        const CSV Hi{w + L.Hinv + q, L.ldq};
        double t[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          double a = 0.0;
#pragma unroll
          for (int k2 = 0; k2 < NW; ++k2) a += Hi[i * NW + k2] * NQ(dzL, q, k2);
          t[i] = a;
        }
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0, b_ = 0.0;
#pragma unroll
          for (int i = 0; i < NW; ++i) { a += Jq[r * NW + i] * NQ(dz, q, i); b_ += Jq[r * NW + i] * t[i]; }
          rc[r] -= a; bj[r] += b_;
        }
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        rmax = fmax(rmax, (rc[r] != rc[r]) ? INFINITY : fabs(rc[r]));
        bb[r * ld + j] = bj[r] - rc[r];
      }
    }
    rmax = block_max(rmax, red);
    smax = block_max(smax, red);
    MYR_SYNC();
    if (!(rmax > 1e-13 * fmax(1.0, smax)) || !isfinite(rmax)) break;
    block_cr_resolve<NC>(St, ld, D, VL, VU, bb, w + L.dzU /* ddlam: lds*NC <= ldq*NW */);
    for (int j = MYR_TID; j < St; j += MYR_NT)
#pragma unroll
      for (int r = 0; r < NC; ++r) NS(dlam, j, r) += NS(dzU, j, r);
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      double v[NW];
#pragma unroll
      for (int i = 0; i < NW; ++i) v[i] = NQ(dzL, q, i);
      const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
      if (jp >= 0) {
        const CSV Gq{w + L.G + q, L.ldq};
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          const double d = NS(dzU, jp, r);
#pragma unroll
          for (int i = 0; i < NW; ++i) v[i] -= Gq[r * NW + i] * d;
        }
      }
      if (js >= 0) {
        const CSV Fq{w + L.F + q, L.ldq};
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          const double d = NS(dzU, js, r);
#pragma unroll
          for (int i = 0; i < NW; ++i) v[i] -= Fq[r * NW + i] * d;
        }
      }
      const CSV Hi{w + L.Hinv + q, L.ldq};
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NW; ++k) a += Hi[i * NW + k] * v[k];
        NQ(dz, q, i) += a;
      }
    }
    MYR_SYNC();
  }
  return ok;
}

// Second-order-correction solve (IPOPT A-5.5 ff.) with the factors the last kkt_solve left behind: same matrix,
// constraint right-hand side csoc instead of c.   S dl2 = csoc - J Hinv rb,   dz2 = -Hinv (rb + J^T dl2).
template <class S>
MYR_HDN void kkt_soc_solve(const Problem& P, const Layout<S>& L, double* w, double* cr) {
  constexpr int NW = S::NW, NC = S::NC;
  const int Q = L.Q, St = L.St, ld = L.lds;
  double* D = cr; double* VL = D + 2 * ld * NC * NC; double* VU = VL + ld * NC * NC;
  double* bb = VU + ld * NC * NC;
  for (int j = MYR_TID; j < St; j += MYR_NT)
#pragma unroll
    for (int r = 0; r < NC; ++r) bb[r * ld + j] = NS(csoc, j, r) + NS(sch, j, r);
  MYR_SYNC();
  block_cr_resolve<NC>(St, ld, D, VL, VU, bb, w + L.dl2);
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
    double v[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) v[i] = NQ(rb, q, i);
    const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
    if (jp >= 0) {
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        const double d = NS(dl2, jp, r);
#pragma unroll
        for (int i = 0; i < NW; ++i) v[i] += NQ(G, q, r * NW + i) * d;
      }
    }
    if (js >= 0) {
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        const double d = NS(dl2, js, r);
#pragma unroll
        for (int i = 0; i < NW; ++i) v[i] += NQ(F, q, r * NW + i) * d;
      }
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NW; ++k) a += NQ(Hinv, q, i * NW + k) * v[k];
      NQ(dz2, q, i) = -a;
    }
  }
  MYR_SYNC();
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
      NQ(lbr, q, i) = bd.fixed ? lb[id] : (bd.hasL ? bd.lbr : -INFINITY);
      NQ(ubr, q, i) = bd.fixed ? lb[id] : (bd.hasU ? bd.ubr : INFINITY);
    }
    fix_sh[q] = fm;
  }
  for (int k = MYR_TID; k < ncn; k += MYR_NT) lam[k] = 0.0;
  MYR_SYNC();

  MYR_PH_DECL
  double mu = O.mu_init, nu = 1.0, delta_last = 0.0;
  int it = 0, status = ST_MAXITER, n_acceptable = 0, hard_iters = 0;
  bool soc_armed = false;
  double f = 0.0, E0 = INFINITY, cinf = INFINITY, c1 = 0.0;
  const double mu_floor = fmax(O.mu_min, O.tol / 10.0);

  while (true) {
    // ---------------- K1: evaluate with derivatives
    MYR_PH(9);
    f = eval_nodes<S, 2>(P, L, z, lam, w, red, mlp_scr);
    stage_constraints<S>(P, L, w, w + L.c, red, cinf, c1);
    MYR_PH(0);
    // ---------------- dual residual, complementarity, scaling sums
    // (per-variable loops are deliberately NOT unrolled: the body is long and the kernel is instruction-fetch bound)
    double rdmax = 0.0, szmax = -INFINITY, szmin = INFINITY, sumz = 0.0, nbnd = 0.0, slog = 0.0;
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
      double lp[NC], ls[NC];
#pragma unroll
      for (int rr = 0; rr < NC; ++rr) {
        lp[rr] = jp >= 0 ? lam[S::cidx(P, jp, rr)] : 0.0;
        ls[rr] = js >= 0 ? lam[S::cidx(P, js, rr)] : 0.0;
      }
      const uint32_t fm = fix_sh[q];
      LogProd lpq;
#pragma unroll 1
      for (int i = 0; i < NW; ++i) {
        const int id = S::zidx(P, q, i);
        double r = NQ(gl, q, i);
#pragma unroll
        for (int rr = 0; rr < NC; ++rr) r += NQ(G, q, rr * NW + i) * lp[rr] + NQ(F, q, rr * NW + i) * ls[rr];
        const bool fixed = (fm >> i) & 1u;
        double rd = 0.0;
        if (!fixed) {
          const double lo = NQ(lbr, q, i), hi = NQ(ubr, q, i), x = z[id], zl = zL[id], zu = zU[id];
          rd = r - zl + zu;
          if (lo > -INFINITY) { const double sl = x - lo; const double pz = sl * zl; szmax = fmax(szmax, pz); szmin = fmin(szmin, pz); sumz += zl; nbnd += 1.0; lpq.mul(sl); }
          if (hi < INFINITY) { const double su = hi - x; const double pz = su * zu; szmax = fmax(szmax, pz); szmin = fmin(szmin, pz); sumz += zu; nbnd += 1.0; lpq.mul(su); }
        }
        NQ(rb, q, i) = fixed ? 0.0 : r;  // grad f + J^T lam (barrier terms added after mu is known)
        const double a = (rd != rd) ? INFINITY : fabs(rd);
        rdmax = fmax(rdmax, a);
      }
      slog += lpq.value();
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

    MYR_PH(1);
    // ---------------- barrier gradient rb and Sigma (reciprocal slacks are kept for the step-size phase)
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      const uint32_t fm = fix_sh[q];
#pragma unroll 1
      for (int i = 0; i < NW; ++i) {
        const int id = S::zidx(P, q, i);
        const bool fixed = (fm >> i) & 1u;
        double sg = 0.0, rbv = NQ(rb, q, i), r1 = 0.0, r2 = 0.0;
        if (!fixed) {
          const double lo = NQ(lbr, q, i), hi = NQ(ubr, q, i), x = z[id];
          if (lo > -INFINITY) { r1 = 1.0 / (x - lo); sg += zL[id] * r1; rbv -= mu * r1; }
          if (hi < INFINITY) { r2 = 1.0 / (hi - x); sg += zU[id] * r2; rbv += mu * r2; }
        }
        NQ(rsl, q, i) = r1; NQ(rsu, q, i) = r2;
        sig_sh[q * NW + i] = sg;
        NQ(rb, q, i) = fixed ? 0.0 : rbv;
      }
    }
    MYR_SYNC();

    MYR_PH(2);
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

    MYR_PH(3);
    // ---------------- step sizes (fraction to the boundary), bound-multiplier steps, merit derivative
    // alpha = min(1, tau / max_i(-d_i / slack_i)): one division at the end instead of one per variable
    double m_pr = 0.0, m_du = 0.0, dphi = 0.0, dHd = 0.0;
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
      double lp[NC], ls[NC], dp[NC], ds[NC];
#pragma unroll
      for (int rr = 0; rr < NC; ++rr) {
        lp[rr] = jp >= 0 ? lam[S::cidx(P, jp, rr)] : 0.0;
        ls[rr] = js >= 0 ? lam[S::cidx(P, js, rr)] : 0.0;
        dp[rr] = jp >= 0 ? NS(dlam, jp, rr) : 0.0;
        ds[rr] = js >= 0 ? NS(dlam, js, rr) : 0.0;
      }
      const uint32_t fm = fix_sh[q];
#pragma unroll 1
      for (int i = 0; i < NW; ++i) {
        const int id = S::zidx(P, q, i);
        const bool fixed = (fm >> i) & 1u;
        // J^T (lam + dlam) part of  H dz = -(rb + J^T dlam): dz^T H dz = -dz.(rb + J^T dlam)
        double jt = 0.0, jl = 0.0;
#pragma unroll
        for (int rr = 0; rr < NC; ++rr) {
          const double g = NQ(G, q, rr * NW + i), f_ = NQ(F, q, rr * NW + i);
          jt += g * dp[rr] + f_ * ds[rr];
          jl += g * lp[rr] + f_ * ls[rr];
        }
        const double d = NQ(dz, q, i);
        double dl = 0.0, du = 0.0;
        if (!fixed) {
          const double rbv = NQ(rb, q, i);
          dHd -= d * (rbv + jt);
          dphi += d * (rbv - jl);  // barrier-objective gradient = rb - J^T lam
          const double r1 = NQ(rsl, q, i), r2 = NQ(rsu, q, i);
          if (r1 != 0.0) {
            const double zl = zL[id];
            dl = mu * r1 - zl - zl * r1 * d;
            m_pr = fmax(m_pr, -d * r1);
            if (dl < 0.0) m_du = fmax(m_du, -dl / zl);
          }
          if (r2 != 0.0) {
            const double zu = zU[id];
            du = mu * r2 - zu + zu * r2 * d;
            m_pr = fmax(m_pr, d * r2);
            if (du < 0.0) m_du = fmax(m_du, -du / zu);
          }
        }
        NQ(dzL, q, i) = dl;
        NQ(dzU, q, i) = du;
      }
    }
    m_pr = block_max(m_pr, red);
    m_du = block_max(m_du, red);
    dphi = block_sum(dphi, red);
    dHd = block_sum(dHd, red);
    if (!(dphi == dphi)) { status = ST_NAN; break; }
    const double a_pr = m_pr > tau ? tau / m_pr : 1.0;
    const double a_du = m_du > tau ? tau / m_du : 1.0;

    MYR_PH(4);
    // ---------------- l1-merit backtracking line search
    if (c1 > 0.0) {
      const double nu_trial = (dphi + 0.5 * fmax(dHd, 0.0)) / ((1.0 - O.rho) * c1);
      if (nu < nu_trial) nu = nu_trial + 1.0;
    }
    const double Dm = dphi - nu * c1;
    const double phi0 = f - mu * slog + nu * c1;
    double alpha = a_pr;
    bool accepted = false;
    double* zt = w + L.zt;
    double f_t = 0.0;
    // Armijo with IPOPT's rounding-error relaxation (10 eps |phi|) so that converged iterates are not rejected by cancellation
    const double armijo_slack = 10.0 * 2.220446049250313e-16 * fabs(phi0);
    // One loop serves the backtracking trials (soc == false: step dz, length alpha) and the second-order-correction
    // trials (soc == true: step dz2 from kkt_soc_solve, length a2), so the trial evaluation exists once in the code.
    int ls = 0, ks = 0, ls_used = 0;
    bool soc = false;
    double a2 = 1.0, c1_prev = 0.0;
    while (ls < O.max_ls) {
      if (soc) {
        kkt_soc_solve<S>(P, L, w, cr);
        double m2 = 0.0;
        for (int q = MYR_TID; q < Q; q += MYR_NT) {
#pragma unroll 1
          for (int i = 0; i < NW; ++i) {
            const double d = NQ(dz2, q, i);
            m2 = fmax(m2, fmax(-d * NQ(rsl, q, i), d * NQ(rsu, q, i)));
          }
        }
        m2 = block_max(m2, red);
        a2 = m2 > tau ? tau / m2 : 1.0;
      }
      const double a = soc ? a2 : alpha;
      const int step_off = soc ? L.dz2 : L.dz;
      ls_used = ls;
      // ---- trial point z + a * step, its barrier log-sum, objective and constraints
      double blog = 0.0;
      for (int q = MYR_TID; q < Q; q += MYR_NT) {
        const uint32_t fm = fix_sh[q];
        LogProd lpq;
#pragma unroll 1
        for (int i = 0; i < NW; ++i) {
          const int id = S::zidx(P, q, i);
          const double x = z[id] + a * w[step_off + i * L.ldq + q];
          zt[id] = x;
          if (!((fm >> i) & 1u)) {
            const double lo = NQ(lbr, q, i), hi = NQ(ubr, q, i);
            if (lo > -INFINITY) lpq.mul(x - lo);
            if (hi < INFINITY) lpq.mul(hi - x);
          }
        }
        blog += lpq.value();
      }
      MYR_SYNC();
      f_t = eval_nodes<S, 0>(P, L, zt, lam, w, red, mlp_scr);
      double ci_t, c1_t;
      stage_constraints<S>(P, L, w, w + L.ct, red, ci_t, c1_t);
      blog = block_sum(blog, red);
      const double phit = f_t - mu * blog + nu * c1_t;
      if (isfinite(phit) && phit <= phi0 + O.eta * alpha * Dm + armijo_slack) { accepted = true; break; }
      if (!soc) {
        // second-order correction (IPOPT A-5.5 .. A-5.9): the full step was rejected without reducing the infeasibility,
        // typically because the constraint curvature along dz outweighs the predicted decrease (Maratos effect).
        // Re-solve with the same factors and the constraint right-hand side  alpha c + c(z + alpha dz).
        // It is armed per instance only after two consecutive iterations that needed >= 4 step halvings: on the well
        // behaved instances (the bulk of a batch) an unconditional SOC costs more re-solves than it saves iterations
        // (CARTPOLE N=100: -12 % iterations, +16 % time), on the hard ones it is what makes the method converge.
        if (ls == 0 && soc_armed && O.max_soc > 0 && c1_t >= c1) {
          for (int j = MYR_TID; j < St; j += MYR_NT)
#pragma unroll
            for (int r = 0; r < NC; ++r) NS(csoc, j, r) = alpha * NS(c, j, r) + NS(ct, j, r);
          MYR_SYNC();
          soc = true; ks = 0; c1_prev = c1_t;
          continue;
        }
        alpha *= 0.5; ++ls;
      } else {
        // kappa_soc: the correction must keep reducing the infeasibility, at most max_soc times
        if (!(c1_t <= 0.99 * c1_prev) || ++ks >= O.max_soc) { soc = false; alpha *= 0.5; ++ls; continue; }
        c1_prev = c1_t;
        for (int j = MYR_TID; j < St; j += MYR_NT)
#pragma unroll
          for (int r = 0; r < NC; ++r) NS(csoc, j, r) = a2 * NS(csoc, j, r) + NS(ct, j, r);
        MYR_SYNC();
      }
    }
    if (accepted && soc) {  // the multiplier step of the corrected system goes with the corrected primal step
      for (int j = MYR_TID; j < St; j += MYR_NT)
#pragma unroll
        for (int r = 0; r < NC; ++r) NS(dlam, j, r) = NS(dl2, j, r);
      alpha = a2;
    }
    hard_iters = ls_used >= 4 ? hard_iters + 1 : 0;
#ifndef MYR_SOC_ARM_AFTER
#define MYR_SOC_ARM_AFTER 2
#endif
    if (hard_iters >= MYR_SOC_ARM_AFTER) soc_armed = true;
    if (!accepted) { status = (E0 <= O.acceptable_tol) ? ST_ACCEPTABLE : ST_LINESEARCH; break; }

    MYR_PH(5);
    // ---------------- accept: primal, equality multipliers, bound multipliers (with the kappa_sigma safeguard)
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      const uint32_t fm = fix_sh[q];
#pragma unroll 1
      for (int i = 0; i < NW; ++i) {
        if ((fm >> i) & 1u) continue;
        const int id = S::zidx(P, q, i);
        const double x = zt[id];
        z[id] = x;
        const double lo = NQ(lbr, q, i), hi = NQ(ubr, q, i);
        if (lo > -INFINITY) {
          const double ms = mu / (x - lo);
          double v = zL[id] + a_du * NQ(dzL, q, i);
          v = fmax(fmin(v, O.kappa_sigma * ms), ms / O.kappa_sigma);
          zL[id] = v;
        }
        if (hi < INFINITY) {
          const double ms = mu / (hi - x);
          double v = zU[id] + a_du * NQ(dzU, q, i);
          v = fmax(fmin(v, O.kappa_sigma * ms), ms / O.kappa_sigma);
          zU[id] = v;
        }
      }
    }
    for (int j = MYR_TID; j < St; j += MYR_NT)
#pragma unroll
      for (int r = 0; r < NC; ++r) lam[S::cidx(P, j, r)] += alpha * NS(dlam, j, r);
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
  // the lifted NLP puts a terminal cost tc . s_G on the last shooting node itself, the reference's objective puts it on
  // the rollout end px_{K-1} (shooting.py:205-208): same optimum, but the multiplier of the last block differs by tc
  double tcoef[n];
#pragma unroll
  for (int i = 0; i < n; ++i) tcoef[i] = 0.0;
  if (S::System::has_terminal && P.terminal_cost) S::System::terminal_coef(P.p, tcoef);
  for (int k = MYR_TID; k < P.N; k += MYR_NT)
#pragma unroll
    for (int i = 0; i < n; ++i) lam[k * n + i] = lamI[((k + 1) * P.cpi - 1) * NC + i] - (k == P.N - 1 ? tcoef[i] : 0.0);
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
