// Per-instance engine: K1 (evaluation of the transcribed NLP with block Jacobian / Hessian),
// K2 (block-structured KKT solve: node-block inverses, Schur complement, block cyclic reduction with
// inertia) and K3 (primal-dual interior-point iteration).  One CTA works on one problem instance at a
// time (persistent CTAs pull instances from a counter); threads stride over nodes / stages.  Written
// once for device and host (see common.cuh).
//
// Memory plan (round 2).  All per-instance state lives in a WORKSPACE SLOT owned by the resident CTA -- not by
// the instance -- so that the footprint is (resident CTAs) x (slot) whatever the batch.  A slot is a set of named
// arrays; each array is placed either in the CTA's shared memory or in the slot's global-memory part (L2 resident)
// by a priority list filled greedily against the shared-memory budget (Layout::place): the block-cyclic-reduction
// scratch first, then the node-block inverses and the Jacobian blocks, then the iterate vectors.
//   * matrices are "array of blocks": block q at base + q * stride, stride odd (in doubles) so that consecutive
//     nodes' 8-byte accesses fall into distinct banks and every element has a compile-time offset;
//   * node vectors are element-major (base + i * ldq + q): coalesced in global memory, conflict-free in shared;
//   * stage vectors are stage-major (base + j * NC + r), the layout the cyclic reduction reads.
#pragma once
#include <type_traits>

#include "schemes.cuh"
#include "shooting.cuh"

namespace myr {

struct IpmOpts {
  int max_iter;
  int max_ls;
  int acceptable_iter;
  int use_filter;     // line search: a trial point is accepted by the l1-merit Armijo test OR by IPOPT's filter rules
  int use_watchdog;   // after 10 shortened steps in a row: up to 3 full steps judged against the stored iterate, else roll back
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

// collocation schemes whose role Jacobians are affine in the dynamics Jacobian (schemes.cuh, kAffineJ): the slot stores
// J (n x NW per node) instead of G and F, and every product with a role Jacobian is formed from J and the role's
// coefficients -- half the shared memory and less than half the flops of the per-node phase
template <class S, class = void>
struct scheme_is_affine { static constexpr bool value = false; };
template <class S>
struct scheme_is_affine<S, typename std::enable_if<S::kAffineJ>::type> { static constexpr bool value = true; };

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

// ------------------------------------------------------------------ workspace arrays
// smallest v >= n with v = 2 (mod 4)
MYR_HDI constexpr int pad2(int n) { return ((n + 1) & ~3) + 2; }
// smallest odd v >= n: stride of per-node blocks.  The compiler emits 64-bit accesses (generic pointers, no alignment
// facts), for which 16 consecutive nodes hit 16 distinct bank pairs exactly when the stride in doubles is odd.
MYR_HDI constexpr int oddpad(int n) { return n | 1; }

// Layout of the NC x NC blocks of the cyclic reduction (crD / crU / crVL / crVU): row stride RS, block stride BS, and
// the POSITION of block i (cr_pos).  A lane group of G lanes owns a block, lane r its row r; consecutive groups work on
// consecutive positions.  The strides are chosen so that the 16 lanes of a half-warp (4 groups x 4 rows, 8 x 2, ...) fall
// into 16 distinct bank pairs: 4 x 4 blocks keep dense rows and an odd block stride (17: row offsets {0,4,8,12} + 4
// consecutive positions), 2 x 2 blocks pad the rows (3 / 6); 8 x 8 blocks keep the dense layout (their scratch already
// fills the shared memory of an SM at N = 100).
template <int NC>
struct CrLay {
  static constexpr int RS = NC == 2 ? 3 : NC;
  static constexpr int BS = NC == 2 ? 6 : (NC <= 4 ? oddpad(NC * NC) : pad2(NC * NC));
};
// Blocks are stored in ELIMINATION ORDER: the blocks eliminated at level l (i = (2k+1) 2^l) occupy consecutive positions
// base_l + k, block 0 (the root) comes last -- so the groups of a warp always touch adjacent blocks, at every level (in
// natural order the active blocks of level l are 2^(l+1) apart and collide in the same banks).
MYR_HDI int cr_pos(int i, int St) {
  if (i == 0) return St - 1;
#ifdef __CUDA_ARCH__
  const int l = __ffs(i) - 1;
#else
  const int l = __builtin_ctz((unsigned)i);
#endif
  return (St - 1) - ((St - 1) >> l) + (i >> (l + 1));
}

template <class S>
struct Dims {
  static constexpr int NW = S::NW, NC = S::NC, NWP = S::NWP;
  static constexpr bool kAff = scheme_is_affine<S>::value;
  static constexpr int GS = kAff ? oddpad(S::n * NW) : oddpad(NC * NW);   // Jacobian block of a node role (affine schemes: J itself)
  static constexpr int HS = oddpad(NWP);     // node-block inverse (symmetric: packed upper triangle)
  static constexpr int WSZ = oddpad(NWP);    // packed Hessian block
  static constexpr int BS = CrLay<NC>::BS;   // Schur-complement block (rows RS = CrLay<NC>::RS apart)
};

// X(name, doubles): in shared-memory PRIORITY order
#define MYR_WS_ARRAYS(X)                                                                                              \
  X(crD, St * D::BS) X(crU, St * D::BS) X(crVL, St * D::BS) X(crVU, St * D::BS) X(crb, St * NC) X(dlam, St * NC)       \
  X(Hinv, Q * D::HS) X(G, Q * D::GS) X(F, D::kAff ? 0 : Q * D::GS)                                                                  \
  X(dynf, kCoopMlp ? Q * S::n : 0) X(dynJ, kCoopMlp ? Q * S::n * NW : 0) X(dynH, kCoopMlp ? Q * S::NWP : 0)           \
  X(lam, St * NC) X(z, ldq * NW) X(rb, ldq * NW) X(dz, ldq * NW) X(fixm, (Q + 1) / 2)                                  \
  X(zL, ldq * NW) X(zU, ldq * NW) X(lbr, ldq * NW) X(ubr, ldq * NW) X(gl, ldq * NW) X(W, Q * D::WSZ)                    \
  X(sig, ldq * NW) X(rsl, ldq * NW) X(rsu, ldq * NW) X(dzL, ldq * NW) X(dzU, ldq * NW)                  \
  X(phi, Q * NC) X(psi, Q * NC) X(c, St * NC) X(sch, St * NC) X(ct, St * NC)                                          \
  X(dz2, ldq * NW) X(csoc, St * NC) X(dl2, St * NC)                                                                   \
  X(wz, ldq * NW) X(wzL, ldq * NW) X(wzU, ldq * NW) X(wlam, St * NC)   /* watchdog: the stored iterate */               \
  X(ext, scheme_is_lifted<S>::value ? 6 * Q * NW + St * NC : 0)

template <class S>
struct Layout {
  using D = Dims<S>;
  static constexpr int NW = S::NW, NC = S::NC;
  // collocation on a NODE system: the MLP is evaluated for all nodes of the instance cooperatively (node_mlp.cuh)
  static constexpr bool kCoopMlp = sys_is_node<typename S::System>::value && !scheme_is_lifted<S>::value;
  enum : int {
#define X(name, sz) A_##name,
    MYR_WS_ARRAYS(X)
#undef X
    A_COUNT
  };
  static constexpr int kCrArrays = 6;  // the leading arrays that make up the cyclic-reduction scratch
  int Q, St, ldq;
  int size[A_COUNT];
  MYR_HDI explicit Layout(const Problem& P) {
    Q = S::num_nodes(P); St = S::num_stages(P);
    ldq = Q;   // node vectors are element-major with 8-byte accesses: any leading dimension is conflict-free
#define X(name, sz) size[A_##name] = ((sz) + 1) & ~1;
    MYR_WS_ARRAYS(X)
#undef X
    // The coupling blocks crU are dead from the last update phase of the cyclic reduction until the next node phase
    // writes them again -- exactly the window in which the role values phi / psi (node evaluation -> stage constraints,
    // main evaluation and line-search trials) and the trial constraints ct are alive: they live inside crU when it is
    // large enough (block sizes >= 4), i.e. in shared memory whenever the reduction scratch is.
    alias_u = size[A_crU] >= 2 * Q * NC + St * NC;
    if (alias_u) size[A_phi] = size[A_psi] = size[A_ct] = 0;
  }
  bool alias_u;
  // greedy placement against a shared-memory budget (doubles): bit a of the result <=> array a is in shared memory
  MYR_HDI unsigned long long place(long long budget, int& smem_doubles, int& glob_doubles) const {
    unsigned long long mask = 0;
    long long used = 0, glob = 0;
    for (int a = 0; a < A_COUNT; ++a) {
      if (size[a] > 0 && used + size[a] <= budget) { mask |= 1ull << a; used += size[a]; }
      else glob += size[a];
    }
    smem_doubles = (int)used;
    glob_doubles = (int)((glob + 15) & ~15ll);
    return mask;
  }
  MYR_HDI int cr_doubles() const { int s = 0; for (int a = 0; a < kCrArrays; ++a) s += size[a]; return s; }
};

// pointers to the arrays of the slot this CTA works in
template <class S>
struct WS {
  using L = Layout<S>;
  int Q, St, ldq;
#define X(name, sz) double* name;
  MYR_WS_ARRAYS(X)
#undef X
  double* red;       // reduction scratch: 2 * kRedStride doubles
  double* filt;      // line-search filter: 2 * kFilterMax doubles (shared memory; null on the host, which keeps it on the stack)
  double* mlp_scr;   // NODE systems: scratch of the cooperative MLP pass (shared memory), else null
  const double* theta;   // NODE systems: shared-memory copy of the MLP weights (null: read the caller's vector)
  int vq0, vi0, vqs, vis;   // VarIter: start (node, component) of this thread and its stride
  int sh;                   // 0: nothing assumed; 1: CR scratch in shared memory; 2: also Hinv, G, F (MYR_ASSUME_SHARED)
  MYR_HDI WS(const L& lay, unsigned long long mask, double* smem, double* glob) {
    Q = lay.Q; St = lay.St; ldq = lay.ldq;
    vi0 = MYR_TID / Q; vq0 = MYR_TID - vi0 * Q; vis = MYR_NT / Q; vqs = MYR_NT - vis * Q;
    double* sp = smem; double* gp = glob;
#define X(name, sz) if ((mask >> L::A_##name) & 1ull) { name = sp; sp += lay.size[L::A_##name]; } else { name = gp; gp += lay.size[L::A_##name]; }
    MYR_WS_ARRAYS(X)
#undef X
    if (lay.alias_u) { phi = crU; psi = crU + Q * L::NC; ct = psi + Q * L::NC; }
    red = nullptr; filt = nullptr; mlp_scr = nullptr; theta = nullptr; sh = 0;
  }
  MYR_HDI uint32_t* fix() const { return reinterpret_cast<uint32_t*>(fixm); }
};

// Flat loop over all variables (node q, component i) of the instance: thread t takes elements t, t + NT, ... of the
// Q * NW variables (i-major), advanced without integer divisions.  Vector phases use it so that the work spreads over
// every thread of the CTA (not just one thread per node) and each thread's loads are independent of one another.
template <class S>
struct VarIter {
  int q, i;
  MYR_HDI explicit VarIter(const WS<S>& ws) : q(ws.vq0), i(ws.vi0) {}
  MYR_HDI bool valid() const { return i < S::NW; }
  MYR_HDI void next(const WS<S>& ws) {
    q += ws.vqs; i += ws.vis;
    if (q >= ws.Q) { q -= ws.Q; ++i; }
  }
};
#define MYR_FOR_VARS(it) for (VarIter<S> it(ws); it.valid(); it.next(ws))


// Address-space hints.  The slot's arrays are reached through generic pointers (an array may live in shared OR global
// memory, Layout::place decides at run time); a generic load from shared memory goes through the global-load path --
// long-scoreboard latency, 64-bit addressing.  The hot routines are therefore compiled in variants that ASSUME their
// matrices are in shared memory (SH = 1: the cyclic-reduction scratch; SH = 2: also Hinv / G / F) and the kernel
// picks the variant that matches the placement mask; nvcc then emits LDS / STS for them.
#ifdef __CUDA_ARCH__
#define MYR_ASSUME_SHARED(p) __builtin_assume(__isShared(p))
#else
#define MYR_ASSUME_SHARED(p) ((void)0)
#endif

#define NQ(arr, q, e) ws.arr[(e) * ws.ldq + (q)]
#define NS(arr, j, r) ws.arr[(j) * NC + (r)]

// ------------------------------------------------------------------ fused block reductions
// K values are reduced in ONE pass (independent shuffle chains, one barrier): a per-quantity block_sum costs two
// barriers and a dependent shuffle chain each, and the interior-point iteration needs about twenty of them.
// red: 2 * kRedStride doubles of shared memory, used alternately (parity) so that no leading barrier is needed.
constexpr int kRedMaxK = 12;
constexpr int kRedStride = 8 * kRedMaxK;
constexpr int kFilterMax = 16;                                   // entries (theta, phi) of the line-search filter
constexpr int kFixedScratch = 2 * kRedStride + 2 * kFilterMax;   // per-CTA shared memory ahead of the slot arrays
enum RedOp : int { R_SUM = 0, R_MAX = 1, R_MIN = 2 };
MYR_HDI double red_apply(int op, double a, double b) { return op == R_SUM ? a + b : (op == R_MAX ? fmax(a, b) : fmin(a, b)); }

template <int... OPS>
MYR_HDI void block_reduce_multi(double* v, double* red, int& parity) {
#ifdef __CUDA_ARCH__
  constexpr int K = sizeof...(OPS);
  static_assert(K <= kRedMaxK, "too many fused reductions");
  constexpr int ops[K] = {OPS...};
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  // the loops over shuffle distances / warps stay rolled (K independent operations inside): the kernel is
  // instruction-fetch sensitive and this helper is expanded at every call site
#pragma unroll 1
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = red_apply(ops[k], v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
  }
  double* buf = red + parity * kRedStride;
  parity ^= 1;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) buf[w * K + k] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = buf[k];
#pragma unroll 1
  for (int ww = 1; ww < nw; ++ww) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = red_apply(ops[k], v[k], buf[ww * K + k]);
  }
#else
  (void)v; (void)red; (void)parity;
#endif
}

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
#ifdef __CUDA_ARCH__
__device__ __noinline__ double log_outlined(double x) { return log(x); }   // one copy of the (long) log sequence in the kernel
#else
inline double log_outlined(double x) { return log(x); }
#endif
struct LogProd {
  double prod = 1.0, sum = 0.0;
  bool bad = false;
  MYR_HDI void mul(double s) {
    bad = bad || !(s > 0.0);
    prod *= s;
    if (!(prod > 1e-200 && prod < 1e200)) { sum += log_outlined(prod); prod = 1.0; }
  }
  MYR_HDI double value() const { return bad ? NAN : sum + log_outlined(prod); }
};

// multipliers of the slot's lam array as the schemes' node_mu wants them
template <int NC>
struct WsLamView {
  const double* lam;
  MYR_HDI double operator()(int j, int r) const { return lam[j * NC + r]; }
};
// iterate of the slot (element-major)
struct WsZView {
  const double* z; int ldq;
  const double* step; double a;   // optional trial point z + a * step (step == nullptr: the iterate itself)
  MYR_HDI double operator()(int q, int i) const { return step ? z[i * ldq + q] + a * step[i * ldq + q] : z[i * ldq + q]; }
};

// ------------------------------------------------------------------ affine role Jacobians (scheme_is_affine)
// w = sum_g a_p[g] dp_g + a_s[g] ds_g,  e = sum_g b_p[g] dp_g + b_s[g] ds_g   (dp / ds: the node's phi / psi stage rows of
// a stage vector; null = role absent), so that  J_roles^T d = J^T w + [e; 0]
template <class S>
MYR_HDI void affine_combine(const Problem& P, int q, const double* dp, const double* ds, double* w, double* e) {
  constexpr int n = S::n, NG = S::NG;
  double ap[NG], bp[NG], as[NG], bs[NG];
  S::role_coefs(P, q, ap, bp, as, bs);
#pragma unroll
  for (int r = 0; r < n; ++r) { w[r] = 0.0; e[r] = 0.0; }
#pragma unroll
  for (int g = 0; g < NG; ++g)
#pragma unroll
    for (int r = 0; r < n; ++r) {
      const double p_ = dp ? dp[g * n + r] : 0.0, s_ = ds ? ds[g * n + r] : 0.0;
      w[r] += ap[g] * p_ + as[g] * s_;
      e[r] += bp[g] * p_ + bs[g] * s_;
    }
}

// ------------------------------------------------------------------ K1: node evaluation sweep
// Evaluates every node at the point zv (element-major node vector: the iterate or a trial point) and stores the node
// arrays in the slot.  MODE as in schemes.cuh; MODE 2 also leaves  grad f + J^T lam  (zero on fixed variables) in rb.
// Returns this thread's part of the objective (the caller reduces it together with its other sums).
// step / a / blog: line-search trials evaluate z + a * step directly (the trial point is never stored) and return the
// thread's part of the barrier log-sum of the trial slacks.
template <class S, int MODE>
MYR_HDI double eval_nodes(const Problem& P, const WS<S>& ws, const double* zv, const double* step = nullptr, double a_step = 0.0,
                          double* blog = nullptr) {
  using D = Dims<S>;
  constexpr int NW = S::NW, NC = S::NC;
  const int Q = ws.Q;
  double fsum = 0.0;
  bool have_pre = false;
#ifdef __CUDA_ARCH__
  if constexpr (Layout<S>::kCoopMlp) {
    mlp_nodes_pass<S, MODE>(P, Q, WsZView{zv, ws.ldq, step, a_step}, WsLamView<NC>{ws.lam}, ws.dynf, ws.dynJ, ws.dynH, ws.mlp_scr, ws.theta);
    have_pre = true;
  }
#endif
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
    PreDyn pre;
    if (have_pre) { pre.f = ws.dynf + q * S::n; pre.J = ws.dynJ + q * S::n * NW; pre.H = ws.dynH + q * S::NWP; }
    double v[NW], lp[NC], ls[NC];
#pragma unroll
    for (int i = 0; i < NW; ++i) v[i] = zv[i * ws.ldq + q];
    if (MODE == 0 && step) {
      const uint32_t fm = ws.fix()[q];
      LogProd lpq;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const double lo = NQ(lbr, q, i), hi = NQ(ubr, q, i);
        v[i] += a_step * step[i * ws.ldq + q];
        if (!((fm >> i) & 1u)) {
          if (lo > -INFINITY) lpq.mul(v[i] - lo);
          if (hi < INFINITY) lpq.mul(hi - v[i]);
        }
      }
      *blog += lpq.value();
    }
    const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
    if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        lp[r] = jp >= 0 ? NS(lam, jp, r) : 0.0;
        ls[r] = js >= 0 ? NS(lam, js, r) : 0.0;
      }
    }
    double ell, gl[NW], phi[NC], psi[NC], W[S::NWP];
    if constexpr (D::kAff) {
      double J[S::n * NW];
      S::template eval_node_j<MODE>(P, q, v, lp, ls, ell, gl, phi, psi, J, W, pre);
      if (MODE >= 1) {
        double* Jq = ws.G + q * D::GS;
#pragma unroll
        for (int i = 0; i < S::n * NW; ++i) Jq[i] = J[i];
      }
      if (MODE == 2) {
        double w_[S::n], e_[S::n];
        affine_combine<S>(P, q, jp >= 0 ? lp : nullptr, js >= 0 ? ls : nullptr, w_, e_);
        const uint32_t fm = ws.fix()[q];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          double r = gl[i] + (i < S::n ? e_[i < S::n ? i : 0] : 0.0);
#pragma unroll
          for (int rr = 0; rr < S::n; ++rr) r += J[rr * NW + i] * w_[rr];
          NQ(rb, q, i) = ((fm >> i) & 1u) ? 0.0 : r;
        }
      }
    } else {
      double G[NC * NW], F[NC * NW];
      S::template eval_node<MODE>(P, q, v, lp, ls, ell, gl, phi, psi, G, F, W, pre);
      if (MODE >= 1) {
        double* Gq = ws.G + q * D::GS; double* Fq = ws.F + q * D::GS;
#pragma unroll
        for (int i = 0; i < NC * NW; ++i) { Gq[i] = G[i]; Fq[i] = F[i]; }
      }
      if (MODE == 2) {
        const uint32_t fm = ws.fix()[q];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          double r = gl[i];
#pragma unroll
          for (int rr = 0; rr < NC; ++rr) r += G[rr * NW + i] * lp[rr] + F[rr * NW + i] * ls[rr];
          NQ(rb, q, i) = ((fm >> i) & 1u) ? 0.0 : r;
        }
      }
    }
    fsum += ell;
#pragma unroll
    for (int r = 0; r < NC; ++r) { ws.phi[q * NC + r] = phi[r]; ws.psi[q * NC + r] = psi[r]; }
    if (MODE >= 1) {
#pragma unroll
      for (int i = 0; i < NW; ++i) NQ(gl, q, i) = gl[i];
    }
    if (MODE == 2) {
      double* Wq = ws.W + q * D::WSZ;
#pragma unroll
      for (int i = 0; i < S::NWP; ++i) Wq[i] = W[i];
    }
  }
  return fsum;
}

// stage constraints from node role values; writes cdst (stage-major) and accumulates this thread's (max |c|, sum |c|).
// mag (optional) accumulates the sum of the MAGNITUDES of the role values that were added up: eps * mag is the rounding
// noise of sum |c|, which the line search must not mistake for an increase of the infeasibility.
template <class S>
MYR_HDI void stage_constraints(const Problem& P, const WS<S>& ws, double* cdst, double& mx, double& sm, double* mag = nullptr) {
  constexpr int NC = S::NC;
  double mg = 0.0;
  for (int j = MYR_TID; j < ws.St; j += MYR_NT) {
    const int nk = S::stage_nodes(P, j);
    double a[NC];
#pragma unroll
    for (int r = 0; r < NC; ++r) a[r] = 0.0;
    for (int k = 0; k < nk; ++k) {
      int role; const int q = S::stage_node(P, j, k, role);
      const double* src = (role ? ws.psi : ws.phi) + q * NC;
#pragma unroll
      for (int r = 0; r < NC; ++r) { a[r] += src[r]; mg += fabs(src[r]); }
    }
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      cdst[j * NC + r] = a[r];
      const double aa = (a[r] != a[r]) ? INFINITY : fabs(a[r]);
      mx = fmax(mx, aa); sm += aa;
    }
  }
  if (mag) *mag += mg;
}

// ------------------------------------------------------------------ K2 pieces
// Block cyclic reduction for the symmetric block-tridiagonal system
//   U_{i-1}^T x_{i-1} + D_i x_i + U_i x_{i+1} = b_i,   i = 0..St-1,  blocks NC x NC (D symmetric, may be indefinite).
// Factor and solve are fused (single right-hand side).  Pivot-block inertias are accumulated: by Sylvester's
// law their sum is the inertia of the whole matrix.  D,U,b are destroyed; x receives the solution.
// Blocks are stored one after the other with stride BS = pad2(NC*NC) (row-major inside); b and x are stage-major.
//
// Work decomposition: a GROUP of G = pow2 >= NC adjacent lanes owns one block; lane r of the group holds / produces
// ROW r.  The pivot-block inverse is computed COOPERATIVELY by the group (Gauss-Jordan without pivoting, the pivot row
// broadcast with warp shuffles; the pivots are those of the LDL^T factorisation, so their signs give the inertia and
// the same acceptance test applies), which needs NC registers per lane instead of NC^2 -- that is what lets the
// 8 x 8 blocks of Hermite-Simpson run without spills.  A block whose natural-order pivots are not acceptable (very rare)
// falls back to the pivoted Bunch-Parlett routine, executed redundantly by the lanes of the group.
// On the host (one "thread") a group degenerates to a loop over rows around a full Gauss-Jordan inverse.
template <int NC>
struct CrGroup { static constexpr int G = NC <= 1 ? 1 : (NC <= 2 ? 2 : (NC <= 4 ? 4 : (NC <= 8 ? 8 : 16))); };

#ifdef __CUDA_ARCH__
#define MYR_CR_ROWS(r) for (int r = int(threadIdx.x) % G; r < NC; r += G)
#else
#define MYR_CR_ROWS(r) for (int r = 0; r < NC; ++r)
#endif

// reciprocal for pivots: the hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps -- full double accuracy for the
// normal-range pivots that pass the acceptance tests, a third of the instructions of an IEEE division
MYR_HDI double pivot_rcp(double x) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x;
#endif
}

// Closed-form inverses with inertia for the block sizes that allow them: N = 1, N = 2 (adjugate) and N = 4 (2 x 2
// partition: A11 and its Schur complement are inverted as 2 x 2 blocks -- i.e. block pivots, which also handle a zero
// diagonal entry with a non-zero off-diagonal one, the typical shape of an indefinite block).  Dependent chain: two
// reciprocals and ~15 multiply-adds, no cross-lane traffic.  M: symmetric N x N row-major (both triangles read and
// averaged).  ok = false when a block pivot is too small relative to its entries (the caller falls back to the pivoted
// routine).
template <int N> struct has_closed_inverse { static constexpr bool value = (N == 1 || N == 2 || N == 4); };

MYR_HDI void inv2_sym(double a, double b, double c, double& x00, double& x01, double& x11, bool& ok, int& np, int& nn) {
  const double det = a * c - b * b;
  // the adjugate formula loses eps * (|ac| + b^2) / |det| digits to the cancellation in det: accept it only while that
  // stays below ~1e-9 (invariant under diagonal scaling, so merely badly SCALED blocks still take the fast path)
  ok = ok && (fabs(det) > 1e-7 * (fabs(a * c) + b * b)) && (fabs(det) > 1e-290);
  if (det < 0.0) { ++np; ++nn; } else if (a > 0.0) np += 2; else nn += 2;
  const double r = pivot_rcp(det);
  x00 = c * r; x01 = -b * r; x11 = a * r;
}

template <int N>
MYR_HDI void small_sym_inverse(const double* M, double* X, bool& ok, int& np, int& nn) {
  ok = true; np = 0; nn = 0;
  if (N == 1) {
    const double d = M[0];
    ok = fabs(d) > 1e-290;
    if (d > 0.0) ++np; else ++nn;
    X[0] = pivot_rcp(d);
  } else if (N == 2) {
    double x00, x01, x11;
    inv2_sym(M[0], 0.5 * (M[1] + M[2]), M[3], x00, x01, x11, ok, np, nn);
    X[0] = x00; X[1] = x01; X[2] = x01; X[3] = x11;
  } else {
    const double a00 = M[0], a11 = M[5], a22 = M[10], a33 = M[15];
    const double a01 = 0.5 * (M[1] + M[4]), a02 = 0.5 * (M[2] + M[8]), a03 = 0.5 * (M[3] + M[12]);
    const double a12 = 0.5 * (M[6] + M[9]), a13 = 0.5 * (M[7] + M[13]), a23 = 0.5 * (M[11] + M[14]);
    double p00, p01, p11;                        // A11^-1
    inv2_sym(a00, a01, a11, p00, p01, p11, ok, np, nn);
    const double t00 = p00 * a02 + p01 * a12, t01 = p00 * a03 + p01 * a13;   // T = A11^-1 A12
    const double t10 = p01 * a02 + p11 * a12, t11 = p01 * a03 + p11 * a13;
    // multiplier growth bound, the block analogue of |l_ik| <= 1e7 in the scalar LDL^T test
    ok = ok && (fmax(fmax(fabs(t00), fabs(t01)), fmax(fabs(t10), fabs(t11))) <= 1e7);
    const double s00 = a22 - (a02 * t00 + a12 * t10), s01 = a23 - (a02 * t01 + a12 * t11), s11 = a33 - (a03 * t01 + a13 * t11);
    double q00, q01, q11;                        // S^-1
    inv2_sym(s00, s01, s11, q00, q01, q11, ok, np, nn);
    const double x02 = -(t00 * q00 + t01 * q01), x03 = -(t00 * q01 + t01 * q11);   // X12 = -T S^-1
    const double x12 = -(t10 * q00 + t11 * q01), x13 = -(t10 * q01 + t11 * q11);
    const double x00 = p00 - (x02 * t00 + x03 * t01), x01 = p01 - (x02 * t10 + x03 * t11), x11 = p11 - (x12 * t10 + x13 * t11);
    X[0] = x00; X[1] = x01; X[2] = x02; X[3] = x03;
    X[4] = x01; X[5] = x11; X[6] = x12; X[7] = x13;
    X[8] = x02; X[9] = x12; X[10] = q00; X[11] = q01;
    X[12] = x03; X[13] = x13; X[14] = q01; X[15] = q11;
  }
}

// host: full Gauss-Jordan inverse without pivoting, same arithmetic as the cooperative device version
template <int N>
inline bool gj_inverse_full(double* a /* N x N in/out */, int& np, int& nn) {
  bool ok = true;
  np = nn = 0;
  for (int k = 0; k < N; ++k) {
    double rk[N];
    for (int j = 0; j < N; ++j) rk[j] = a[k * N + j];
    const double p = rk[k];
    double colmax = 0.0, rowmax = 0.0;
    for (int j = 0; j < N; ++j) { const double v = fabs(rk[j]); rowmax = fmax(rowmax, v); if (j > k) colmax = fmax(colmax, v); }
    ok = ok && (fabs(p) > 1e-7 * colmax) && (fabs(p) > 1e-14 * rowmax) && (fabs(p) > 1e-290);
    if (p > 0) ++np; else ++nn;
    const double ip = pivot_rcp(p);
    for (int r = 0; r < N; ++r) {
      const bool piv = (r == k);
      const double f = a[r * N + k];
      for (int j = 0; j < N; ++j) {
        if (j == k) continue;
        const double s = rk[j] * ip;
        a[r * N + j] = piv ? s : a[r * N + j] - f * s;
      }
      a[r * N + k] = piv ? ip : -f * ip;
    }
  }
  return ok;
}

#ifdef __CUDA_ARCH__
// device: lane r of a G-lane group holds row r of the block in a[]; on return a[] is row r of the inverse.
// Must be executed by all 32 lanes of the warp.  ok / np / nn are identical on all lanes of a group.
template <int N, int G>
__device__ __forceinline__ void coop_inverse(double (&a)[N], int r, bool& ok, int& np, int& nn) {
  ok = true; np = 0; nn = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double rk[N];
#pragma unroll
    for (int j = 0; j < N; ++j) rk[j] = __shfl_sync(0xffffffffu, a[j], k, G);
    const double p = rk[k];
    double colmax = 0.0, rowmax = 0.0;
#pragma unroll
    for (int j = 0; j < N; ++j) { const double v = fabs(rk[j]); rowmax = fmax(rowmax, v); if (j > k) colmax = fmax(colmax, v); }
    ok = ok && (fabs(p) > 1e-7 * colmax) && (fabs(p) > 1e-14 * rowmax) && (fabs(p) > 1e-290);
    if (p > 0) ++np; else ++nn;
    const double ip = pivot_rcp(p);
    const bool piv = (r == k);
    const double f = a[k];
#pragma unroll
    for (int j = 0; j < N; ++j) {
      if (j == k) continue;
      const double s = rk[j] * ip;
      a[j] = piv ? s : a[j] - f * s;
    }
    a[k] = piv ? ip : -f * ip;
  }
}
#endif

template <int NC, int SH = 0>
MYR_HDI void block_cr_solve(int St, double* D, double* U, double* VL, double* VU, double* b, double* x,
                            int& cp, int& cn, int& cz) {
  if (SH >= 1) { MYR_ASSUME_SHARED(D); MYR_ASSUME_SHARED(U); MYR_ASSUME_SHARED(VL); MYR_ASSUME_SHARED(VU); MYR_ASSUME_SHARED(b); MYR_ASSUME_SHARED(x); }
  constexpr int BB = NC * NC;
  constexpr int RS = CrLay<NC>::RS, BS = CrLay<NC>::BS;
  constexpr int G = CrGroup<NC>::G;
#ifdef __CUDA_ARCH__
  const int grp = int(threadIdx.x) / G, ngrp = int(blockDim.x) / G;
  const int rl = int(threadIdx.x) % G;
  const bool rvalid = rl < NC;
  const bool counter = rl == 0;
#else
  const int grp = 0, ngrp = 1;
  const bool counter = true;
#endif
  cp = cn = cz = 0;
#if defined(MYR_PROFILE_PHASES) && defined(__CUDA_ARCH__)
  long long cph_t0_ = clock64();
#define MYR_CPH(idx) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_cycles[idx], (unsigned long long)(t_ - cph_t0_)); cph_t0_ = t_; } } while (0)
#else
#define MYR_CPH(idx) do { } while (0)
#endif
  // ---- pivot-block inverse of the block at position pi.  Device: a[] = row rl of the inverse.  Host: Dinv = full inverse.
#ifdef __CUDA_ARCH__
  auto pivot_row = [&](int pi, bool active, double (&a)[NC]) {
    const double* Di = D + pi * BS;
    bool ok; int p_, n_;
    if constexpr (has_closed_inverse<NC>::value) {
      // every lane inverts the whole (small) block redundantly: no cross-lane traffic, and lane rl keeps row rl
      double M[BB], X[BB];
#pragma unroll
      for (int e = 0; e < BB; ++e) M[e] = active ? Di[(e / NC) * RS + (e % NC)] : ((e % (NC + 1)) == 0 ? 1.0 : 0.0);   // idle groups invert the identity
      small_sym_inverse<NC>(M, X, ok, p_, n_);
      __syncwarp();   // every lane of the group has read ALL rows of D_i before any lane overwrites its row below
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        double v = 0.0;
#pragma unroll
        for (int rr = 0; rr < NC; ++rr) v = (rr == rl) ? X[rr * NC + c] : v;
        a[c] = v;
      }
    } else {
#pragma unroll
      for (int c = 0; c < NC; ++c) a[c] = (rvalid && active) ? Di[rl * RS + c] : (c == rl ? 1.0 : 0.0);
      coop_inverse<NC, G>(a, rl, ok, p_, n_);
    }
    int z_ = 0;
    if (!__all_sync(0xffffffffu, ok)) {   // rare: natural-order pivots rejected somewhere in this warp
      if (!ok) {
        double A[BB], inv[BB];
#pragma unroll
        for (int r = 0; r < NC; ++r)
#pragma unroll
          for (int c = 0; c < NC; ++c) A[r * NC + c] = 0.5 * (Di[r * RS + c] + Di[c * RS + r]);
        int p2_, n2_, z2_;
        sym_inverse_inertia<NC>(A, 0u, inv, p2_, n2_, z2_);
        p_ = p2_; n_ = n2_; z_ = z2_;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          double v = 0.0;
#pragma unroll
          for (int rr = 0; rr < NC; ++rr) v = (rr == rl) ? inv[rr * NC + c] : v;
          a[c] = v;
        }
      }
      __syncwarp();   // every lane of the group has read D_i before rows of it are overwritten
    }
    if (counter && active) { cp += p_; cn += n_; cz += z_; }
  };
#else
  auto pivot_full = [&](int pi, double* Dinv) {
    const double* Di = D + pi * BS;
    double M[BB];
    for (int r = 0; r < NC; ++r)
      for (int c = 0; c < NC; ++c) M[r * NC + c] = Di[r * RS + c];
    for (int e = 0; e < BB; ++e) Dinv[e] = M[e];
    int p_, n_, z_ = 0;
    bool ok_;
    if constexpr (has_closed_inverse<NC>::value) {
      small_sym_inverse<NC>(M, Dinv, ok_, p_, n_);
    } else {
      ok_ = gj_inverse_full<NC>(Dinv, p_, n_);
    }
    if (!ok_) {
      double A[BB];
      for (int r = 0; r < NC; ++r)
        for (int c = 0; c < NC; ++c) A[r * NC + c] = 0.5 * (M[r * NC + c] + M[c * NC + r]);
      sym_inverse_inertia<NC>(A, 0u, Dinv, p_, n_, z_);
    }
    cp += p_; cn += n_; cz += z_;
  };
#endif
  int s = 1, ls = 0;   // s = 2^ls
  for (; s < St; s <<= 1, ++ls) {
    const int nodd = (St + s - 1) >> (ls + 1);          // blocks i = (2k+1) s < St
    const int neven = (St + 2 * s - 1) >> (ls + 1);     // blocks i = 2k s < St
    const int pb = (St - 1) - ((St - 1) >> ls);         // position of the first block eliminated at this level (cr_pos)
    // ---- eliminate the odd blocks: Dinv_i (kept in D), VL_i = Dinv_i U_{i-s}^T, VU_i = Dinv_i U_i, x_i = Dinv_i b_i
    for (int k0 = 0; k0 < nodd; k0 += ngrp) {   // trip count uniform across the CTA (the body contains warp shuffles)
      const int k = k0 + grp;
      const bool active = k < nodd;
      const int i = active ? (2 * k + 1) * s : s;
      const int pi = active ? pb + k : pb;
      const bool hr = i + s < St;
#ifdef __CUDA_ARCH__
      double dr[NC];
      pivot_row(pi, active, dr);
      if (active && rvalid) {
        const int r = rl;
#else
      double Dinv[BB];
      pivot_full(pi, Dinv);
      for (int r = 0; r < NC; ++r) {
        double dr[NC];
        for (int m = 0; m < NC; ++m) dr[m] = Dinv[r * NC + m];
#endif
        const double* Ul = U + cr_pos(i - s, St) * BS;
        const double* Ui = U + pi * BS;
        double* Dd = D + pi * BS + r * RS;
        double* VLd = VL + pi * BS + r * RS;
        double* VUd = VU + pi * BS + r * RS;
        double xr = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          double vl = 0.0;
#pragma unroll
          for (int m = 0; m < NC; ++m) vl += dr[m] * Ul[c * RS + m];
          VLd[c] = vl;
          xr += dr[c] * b[i * NC + c];
        }
        if (hr) {
          double vu[NC];
#pragma unroll
          for (int c = 0; c < NC; ++c) vu[c] = 0.0;
#pragma unroll
          for (int m = 0; m < NC; ++m)
#pragma unroll
            for (int c = 0; c < NC; ++c) vu[c] += dr[m] * Ui[m * RS + c];
#pragma unroll
          for (int c = 0; c < NC; ++c) VUd[c] = vu[c];
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) Dd[c] = dr[c];
        x[i * NC + r] = xr;
      }
    }
    MYR_SYNC();
    MYR_CPH(s == 1 ? 10 : (s == 2 ? 12 : 14));
    // ---- update the even blocks i from their eliminated neighbours er = i + s, el = i - s (row r per lane, in place)
    for (int k = grp; k < neven; k += ngrp) {
      const int i = 2 * k * s, er = i + s, el = i - s;
      const bool hr = er < St, hl = el >= 0, hrr = hr && er + s < St;
      const int pi = cr_pos(i, St), per = pb + k, pel = pb + k - 1;
      MYR_CR_ROWS(r) {
        double dn[NC], un[NC];
        double* Dd = D + pi * BS + r * RS;
        double* Ud = U + pi * BS + r * RS;
        double bn = b[i * NC + r];
#pragma unroll
        for (int c = 0; c < NC; ++c) { dn[c] = Dd[c]; un[c] = 0.0; }
        if (hr) {
          const double* VLe = VL + per * BS;
          const double* VUe = VU + per * BS;
          const double* xe = x + er * NC;
#pragma unroll
          for (int m = 0; m < NC; ++m) {
            const double u_ = Ud[m];   // row r of U_i
#pragma unroll
            for (int c = 0; c < NC; ++c) dn[c] -= u_ * VLe[m * RS + c];
            bn -= u_ * xe[m];
            if (hrr) {
#pragma unroll
              for (int c = 0; c < NC; ++c) un[c] -= u_ * VUe[m * RS + c];
            }
          }
        }
        if (hl) {
          const double* Ue = U + pel * BS;
          const double* VUe = VU + pel * BS;
          const double* xe = x + el * NC;
#pragma unroll
          for (int m = 0; m < NC; ++m) {
            const double u_ = Ue[m * RS + r];   // column r of U_el = row r of U_el^T
#pragma unroll
            for (int c = 0; c < NC; ++c) dn[c] -= u_ * VUe[m * RS + c];
            bn -= u_ * xe[m];
          }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) { Dd[c] = dn[c]; Ud[c] = un[c]; }
        b[i * NC + r] = bn;
      }
    }
    MYR_SYNC();
    MYR_CPH(s == 1 ? 11 : (s == 2 ? 13 : 15));
  }
  // ---- root (block 0, last position): every lane only reads its own row of D_0 before overwriting it
  {
    const int p0 = St - 1;
#ifdef __CUDA_ARCH__
    if (int(threadIdx.x) < 32) {   // warp 0 (group 0 lives there); the shuffles need the whole warp
      double dr[NC];
      pivot_row(p0, grp == 0, dr);
      if (grp == 0 && rvalid) {
        const int r = rl;
#else
    {
      double Dinv[BB];
      pivot_full(p0, Dinv);
      for (int r = 0; r < NC; ++r) {
        double dr[NC];
        for (int m = 0; m < NC; ++m) dr[m] = Dinv[r * NC + m];
#endif
        double xr = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) { D[p0 * BS + r * RS + c] = dr[c]; xr += dr[c] * b[c]; }
        x[r] = xr;
      }
    }
  }
  MYR_SYNC();
  // ---- back substitution: x_i = (Dinv_i b_i) - VL_i x_{i-s} - VU_i x_{i+s}
  for (s >>= 1, --ls; s >= 1; s >>= 1, --ls) {
    const int nodd = (St + s - 1) >> (ls + 1);
    const int pb = (St - 1) - ((St - 1) >> ls);
    for (int k = grp; k < nodd; k += ngrp) {
      const int i = (2 * k + 1) * s, pi = pb + k;
      MYR_CR_ROWS(r) {
        double a = x[i * NC + r];
        const double* VLd = VL + pi * BS + r * RS;
        const double* xl = x + (i - s) * NC;
#pragma unroll
        for (int m = 0; m < NC; ++m) a -= VLd[m] * xl[m];
        if (i + s < St) {
          const double* VUd = VU + pi * BS + r * RS;
          const double* xr_ = x + (i + s) * NC;
#pragma unroll
          for (int m = 0; m < NC; ++m) a -= VUd[m] * xr_[m];
        }
        x[i * NC + r] = a;
      }
    }
    MYR_SYNC();
  }
}

// Re-solve with the factors left by block_cr_solve (pivot inverses in D, VL, VU) for a new right-hand side b
// (destroyed); x receives the solution.  One barrier per level: since VL_e = Dinv_e U_{e-s}^T and VU_e = Dinv_e U_e,
// the elimination of e from its surviving neighbours is  b_i -= VL_e^T b_e  (right neighbour e = i+s) and
// b_i -= VU_e^T b_e (left neighbour e = i-s), which only needs data of already-final eliminated nodes.
template <int NC>
MYR_HDN void block_cr_resolve(int St, const double* D, const double* VL, const double* VU, double* b, double* x) {
  constexpr int RS = CrLay<NC>::RS, BS = CrLay<NC>::BS;
  int s = 1, ls = 0;
  for (; s < St; s <<= 1, ++ls) {
    const int pb = (St - 1) - ((St - 1) >> ls);
    for (int k = MYR_TID; 2 * k * s < St; k += MYR_NT) {
      const int i = 2 * k * s;
      double bn[NC];
#pragma unroll
      for (int r = 0; r < NC; ++r) bn[r] = b[i * NC + r];
      const int er = i + s, el = i - s;
      if (er < St) {
        const double* V = VL + (pb + k) * BS;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0;
#pragma unroll
          for (int kk = 0; kk < NC; ++kk) a += V[kk * RS + r] * b[er * NC + kk];
          bn[r] -= a;
        }
      }
      if (el >= 0) {
        const double* V = VU + (pb + k - 1) * BS;
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0;
#pragma unroll
          for (int kk = 0; kk < NC; ++kk) a += V[kk * RS + r] * b[el * NC + kk];
          bn[r] -= a;
        }
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) b[i * NC + r] = bn[r];
    }
    MYR_SYNC();
  }
  if (MYR_TID == 0) {
    const double* D0 = D + (St - 1) * BS;
#pragma unroll
    for (int r = 0; r < NC; ++r) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NC; ++k) a += D0[r * RS + k] * b[k];
      x[r] = a;
    }
  }
  MYR_SYNC();
  for (s >>= 1, --ls; s >= 1; s >>= 1, --ls) {
    const int pb = (St - 1) - ((St - 1) >> ls);
    for (int k = MYR_TID; (2 * k + 1) * s < St; k += MYR_NT) {
      const int i = (2 * k + 1) * s;
      double xi[NC];
      const double* Di = D + (pb + k) * BS;
      const double* VLi = VL + (pb + k) * BS;
      const double* VUi = VU + (pb + k) * BS;
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        double a = 0.0;
#pragma unroll
        for (int kk = 0; kk < NC; ++kk) a += Di[r * RS + kk] * b[i * NC + kk];
        xi[r] = a;
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        double a = 0.0;
#pragma unroll
        for (int kk = 0; kk < NC; ++kk) a += VLi[r * RS + kk] * x[(i - s) * NC + kk];
        xi[r] -= a;
      }
      if (i + s < St) {
#pragma unroll
        for (int r = 0; r < NC; ++r) {
          double a = 0.0;
#pragma unroll
          for (int kk = 0; kk < NC; ++kk) a += VUi[r * RS + kk] * x[(i + s) * NC + kk];
          xi[r] -= a;
        }
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) x[i * NC + r] = xi[r];
    }
    MYR_SYNC();
  }
}

// J_q^T d for the two roles of node q:  u[i] = sum_r G[r][i] dp[r] + F[r][i] ds[r]
template <class S, int SH = 0>
MYR_HDI void jt_times(const Problem& P, const WS<S>& ws, int q, const double* dvec /* stage-major */, double* u) {
  using D = Dims<S>;
  constexpr int NW = S::NW, NC = S::NC;
  const double* const Gb = ws.G; const double* const Fb = ws.F;
  if (SH >= 2) { MYR_ASSUME_SHARED(Gb); MYR_ASSUME_SHARED(Fb); }
  const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
  if constexpr (D::kAff) {
    double w_[S::n], e_[S::n];
    affine_combine<S>(P, q, jp >= 0 ? dvec + jp * NC : nullptr, js >= 0 ? dvec + js * NC : nullptr, w_, e_);
    const double* Jq = Gb + q * D::GS;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      double a = i < S::n ? e_[i < S::n ? i : 0] : 0.0;
#pragma unroll
      for (int r = 0; r < S::n; ++r) a += Jq[r * NW + i] * w_[r];
      u[i] = a;
    }
    (void)Fb;
  } else {
#pragma unroll
    for (int i = 0; i < NW; ++i) u[i] = 0.0;
    if (jp >= 0) {
      const double* Gq = Gb + q * D::GS;
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        const double d = dvec[jp * NC + r];
#pragma unroll
        for (int i = 0; i < NW; ++i) u[i] += Gq[r * NW + i] * d;
      }
    }
    if (js >= 0) {
      const double* Fq = Fb + q * D::GS;
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        const double d = dvec[js * NC + r];
#pragma unroll
        for (int i = 0; i < NW; ++i) u[i] += Fq[r * NW + i] * d;
      }
    }
  }
}

// KKT factorisation + multiplier solve for one instance:
//   [ H + Sigma + dw I   J^T ] [dz  ]     [ rb ]
//   [ J               -dc I ] [dlam] = - [ c  ]
// node data (G, F, W, sig, rb) and c are in the slot.  Phases:
//   (1) per node: Hinv = (W + Sigma + dw)^-1 with fixed variables removed (inertia of H), tv = Hinv rb, and the node's
//       ROLE PRODUCTS  G Hinv G^T, F Hinv F^T (diagonal-block parts), F Hinv G^T (coupling block), G tv, F tv -- every
//       node datum is read once and the products are formed while Hinv is still in registers;
//   (2) per stage: sum the role products of the stage's nodes into the Schur-complement block, right-hand side;
//   (3) block cyclic reduction -> dlam, inertia of S.
// Returns the inertia-ok flag; minpr = smallest relative pivot of the node blocks (refinement is only worth it when small).
template <class S, int SH = 0>
MYR_HDI bool kkt_factor(const Problem& P, const WS<S>& ws, double delta_w, double delta_c, double delta_reg, double& minpr_out,
                        int& parity, double mu_apply = -1.0) {
  using D = Dims<S>;
  constexpr int NW = S::NW, NC = S::NC, BS = D::BS, RS = CrLay<NC>::RS;
  const int Q = ws.Q, St = ws.St;
  double* const crD = ws.crD; double* const crU = ws.crU; double* const crVL = ws.crVL; double* const crVU = ws.crVU;
  double* const crb = ws.crb; double* const Hb = ws.Hinv; const double* const Gb = ws.G; const double* const Fb = ws.F;
  if (SH >= 1) { MYR_ASSUME_SHARED(crD); MYR_ASSUME_SHARED(crU); MYR_ASSUME_SHARED(crVL); MYR_ASSUME_SHARED(crVU); MYR_ASSUME_SHARED(crb); }
  if (SH >= 2) { MYR_ASSUME_SHARED(Hb); MYR_ASSUME_SHARED(Gb); MYR_ASSUME_SHARED(Fb); }
#if defined(MYR_PROFILE_PHASES) && defined(__CUDA_ARCH__)
  long long kph_t0_ = clock64();
#define MYR_KPH(idx) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_phase_cycles[idx], (unsigned long long)(t_ - kph_t0_)); kph_t0_ = t_; } } while (0)
#else
#define MYR_KPH(idx) do { } while (0)
#endif
  // role-product slots: the k-th node block of a stage (phi_slot / psi_slot) writes into its own scratch
  double* const slotM[3] = {crD, crVL, crVU};
  double* const slotV[3] = {crb, ws.dlam, ws.dl2};   // dlam: dead until the reduction writes the solution (not ct: it may live inside crU)
  int hn = 0, hz = 0;
  double minpr = INFINITY;
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
    double A[NW * NW], inv[NW * NW], t[NW];
    const double* Wq = ws.W + q * D::WSZ;
#pragma unroll
    for (int i = 0; i < NW; ++i)
#pragma unroll
      for (int j = 0; j < NW; ++j) A[i * NW + j] = Wq[pidx(i, j, NW)];
    double rbn[NW];
    if (mu_apply >= 0.0) {
      // interior-point caller, first factorisation of the iteration: Sigma and the barrier part of the right-hand side
      // from the reciprocal slacks (kept for inertia-correction retries and the refinement, which read sig / rb)
      const uint32_t fm = ws.fix()[q];
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const double r1 = NQ(rsl, q, i), r2 = NQ(rsu, q, i), zl = NQ(zL, q, i), zu = NQ(zU, q, i);
        double sg = 0.0, rbv = NQ(rb, q, i);
        sg += zl * r1; rbv -= mu_apply * r1;
        sg += zu * r2; rbv += mu_apply * r2;
        rbv = ((fm >> i) & 1u) ? 0.0 : rbv;
        NQ(sig, q, i) = sg; NQ(rb, q, i) = rbv;
        rbn[i] = rbv;
        A[i * NW + i] += sg + delta_w + delta_reg;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NW; ++i) { rbn[i] = NQ(rb, q, i); A[i * NW + i] += NQ(sig, q, i) + delta_w + delta_reg; }
    }
    int p_, n_, z_;
    double pr_;
    sym_inverse<NW>(A, ws.fix()[q], inv, p_, n_, z_, &pr_);
    minpr = fmin(minpr, pr_);
    hn += n_; hz += z_;
    double* Hq = Hb + q * D::HS;
#pragma unroll
    for (int i = 0; i < NW; ++i)
#pragma unroll
      for (int j = i; j < NW; ++j) Hq[pidx(i, j, NW)] = inv[i * NW + j];
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NW; ++k) a += inv[i * NW + k] * rbn[k];
      t[i] = a;
    }
    const int jp = S::phi_stage(P, q), js = S::psi_stage(P, q);
    if constexpr (D::kAff) {
      // role Jacobians  a J + b [I 0]:  with  M = J Hinv,  N = M J^T,  Y = M[:, :n],  X = Hinv[:n, :n]
      //   (a1 J + b1 E) Hinv (a2 J + b2 E)^T = a1 a2 N + a1 b2 Y + b1 a2 Y^T + b1 b2 X
      constexpr int n = S::n, NG = S::NG;
      const double* Jq = Gb + q * D::GS;
      double M[n * NW], Nn[n * n], Jt[n];
#pragma unroll
      for (int r = 0; r < n; ++r) {
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          double a = 0.0;
#pragma unroll
          for (int k = 0; k < NW; ++k) a += Jq[r * NW + k] * inv[k * NW + i];
          M[r * NW + i] = a;
        }
        double a = 0.0;
#pragma unroll
        for (int i = 0; i < NW; ++i) a += Jq[r * NW + i] * t[i];
        Jt[r] = a;
      }
#pragma unroll
      for (int r = 0; r < n; ++r)
#pragma unroll
        for (int c2 = 0; c2 < n; ++c2) {
          double a = 0.0;
#pragma unroll
          for (int i = 0; i < NW; ++i) a += M[r * NW + i] * Jq[c2 * NW + i];
          Nn[r * n + c2] = a;
        }
      double ap[NG], bp[NG], as[NG], bs[NG];
      S::role_coefs(P, q, ap, bp, as, bs);
      auto blk = [&](double a1, double b1, double a2, double b2, int r, int c2) {
        return a1 * a2 * Nn[r * n + c2] + a1 * b2 * M[r * NW + c2] + b1 * a2 * M[c2 * NW + r] + b1 * b2 * inv[r * NW + c2];
      };
      if (jp >= 0) {
        const int sl = S::phi_slot(P, q);
        double* Dst = slotM[sl] + cr_pos(jp, St) * BS;
        double* vst = slotV[sl] + jp * NC;
#pragma unroll
        for (int ga = 0; ga < NG; ++ga)
#pragma unroll
          for (int r = 0; r < n; ++r) {
#pragma unroll
            for (int gb = 0; gb < NG; ++gb)
#pragma unroll
              for (int c2 = 0; c2 < n; ++c2) Dst[(ga * n + r) * RS + gb * n + c2] = blk(ap[ga], bp[ga], ap[gb], bp[gb], r, c2);
            vst[ga * n + r] = ap[ga] * Jt[r] + bp[ga] * t[r];
          }
      }
      if (js >= 0) {
        const int sl = S::psi_slot(P, q);
        const int pjs = cr_pos(js, St);
        double* Dst = slotM[sl] + pjs * BS;
        double* vst = slotV[sl] + js * NC;
        double* Ust = crU + pjs * BS;
        const bool link = jp >= 0;   // the node also starts the next stage: coupling block U_js = F Hinv G^T
#pragma unroll
        for (int ga = 0; ga < NG; ++ga)
#pragma unroll
          for (int r = 0; r < n; ++r) {
#pragma unroll
            for (int gb = 0; gb < NG; ++gb)
#pragma unroll
              for (int c2 = 0; c2 < n; ++c2) {
                Dst[(ga * n + r) * RS + gb * n + c2] = blk(as[ga], bs[ga], as[gb], bs[gb], r, c2);
                if (link) Ust[(ga * n + r) * RS + gb * n + c2] = blk(as[ga], bs[ga], ap[gb], bp[gb], r, c2);
              }
            vst[ga * n + r] = as[ga] * Jt[r] + bs[ga] * t[r];
          }
      }
      (void)Fb;
    } else {
      const double* Gq = Gb + q * D::GS;
      const double* Fq = Fb + q * D::GS;
      if (jp >= 0) {
        const int sl = S::phi_slot(P, q);
        double* Dst = slotM[sl] + cr_pos(jp, St) * BS;
        double* vst = slotV[sl] + jp * NC;
  #pragma unroll
        for (int r = 0; r < NC; ++r) {
          double T[NW];
  #pragma unroll
          for (int i = 0; i < NW; ++i) {
            double a = 0.0;
  #pragma unroll
            for (int k = 0; k < NW; ++k) a += Gq[r * NW + k] * inv[k * NW + i];
            T[i] = a;
          }
  #pragma unroll
          for (int c2 = 0; c2 < NC; ++c2) {
            double a = 0.0;
  #pragma unroll
            for (int i = 0; i < NW; ++i) a += T[i] * Gq[c2 * NW + i];
            Dst[r * RS + c2] = a;
          }
          double a = 0.0;
  #pragma unroll
          for (int i = 0; i < NW; ++i) a += Gq[r * NW + i] * t[i];
          vst[r] = a;
        }
      }
      if (js >= 0) {
        const int sl = S::psi_slot(P, q);
        const int pjs = cr_pos(js, St);
        double* Dst = slotM[sl] + pjs * BS;
        double* vst = slotV[sl] + js * NC;
        double* Ust = crU + pjs * BS;
        const bool link = jp >= 0;   // the node also starts the next stage: coupling block U_js = F Hinv G^T
  #pragma unroll
        for (int r = 0; r < NC; ++r) {
          double T[NW];
  #pragma unroll
          for (int i = 0; i < NW; ++i) {
            double a = 0.0;
  #pragma unroll
            for (int k = 0; k < NW; ++k) a += Fq[r * NW + k] * inv[k * NW + i];
            T[i] = a;
          }
  #pragma unroll
          for (int c2 = 0; c2 < NC; ++c2) {
            double a = 0.0;
  #pragma unroll
            for (int i = 0; i < NW; ++i) a += T[i] * Fq[c2 * NW + i];
            Dst[r * RS + c2] = a;
          }
          if (link) {
  #pragma unroll
            for (int c2 = 0; c2 < NC; ++c2) {
              double a = 0.0;
  #pragma unroll
              for (int i = 0; i < NW; ++i) a += T[i] * Gq[c2 * NW + i];
              Ust[r * RS + c2] = a;
            }
          }
          double a = 0.0;
  #pragma unroll
          for (int i = 0; i < NW; ++i) a += Fq[r * NW + i] * t[i];
          vst[r] = a;
        }
      }
    }
  }
  MYR_SYNC();
  MYR_KPH(6);
  // ---- stage blocks of the Schur complement S = J Hinv J^T + dc I and its right-hand side  c - J Hinv rb
  for (int j = MYR_TID; j < St; j += MYR_NT) {
    double Dj[NC * NC], bj[NC];
    const int nk = S::stage_nodes(P, j);
    const int pj = cr_pos(j, St);
#pragma unroll
    for (int r = 0; r < NC; ++r)
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) Dj[r * NC + c2] = slotM[0][pj * BS + r * RS + c2];
#pragma unroll
    for (int r = 0; r < NC; ++r) bj[r] = NS(c, j, r) - slotV[0][j * NC + r];
    for (int k = 1; k < nk; ++k) {
      const double* Ms = slotM[k] + pj * BS;
      const double* vs = slotV[k] + j * NC;
#pragma unroll
      for (int r = 0; r < NC; ++r)
#pragma unroll
        for (int c2 = 0; c2 < NC; ++c2) Dj[r * NC + c2] += Ms[r * RS + c2];
#pragma unroll
      for (int r = 0; r < NC; ++r) bj[r] -= vs[r];
    }
    double* Dd = crD + pj * BS;
#pragma unroll
    for (int r = 0; r < NC; ++r)
#pragma unroll
      for (int c2 = 0; c2 < NC; ++c2) Dd[r * RS + c2] = 0.5 * (Dj[r * NC + c2] + Dj[c2 * NC + r]) + (r == c2 ? delta_c : 0.0);
#pragma unroll
    for (int r = 0; r < NC; ++r) { crb[j * NC + r] = bj[r]; NS(sch, j, r) = bj[r] - NS(c, j, r); }
  }
  MYR_SYNC();
  MYR_KPH(7);
  int sp, sn, sz;
  block_cr_solve<NC, SH>(St, crD, crU, crVL, crVU, crb, ws.dlam, sp, sn, sz);
  MYR_KPH(8);
  double rv[5] = {(double)hn, (double)hz, minpr, (double)sn, (double)sz};
  block_reduce_multi<R_SUM, R_SUM, R_MIN, R_SUM, R_SUM>(rv, ws.red, parity);
  const int Hneg = (int)(rv[0] + 0.5), Hzero = (int)(rv[1] + 0.5), Sneg = (int)(rv[3] + 0.5), Szero = (int)(rv[4] + 0.5);
  minpr_out = rv[2];
  // inertia(K) = inertia(H) + inertia(-S): correct iff  n-(S) == n-(H)  and nothing is singular
  return (Hzero == 0) && (Szero == 0) && (Sneg == Hneg);
}

// dz = -Hinv (rb + J^T dl) for the multiplier step dl (stage-major); written to the node vector dst.
// Returns this thread's part of  dz^T H dz = -dz . (rb + J^T dl)  (H dz = -(rb + J^T dl) on the free variables).
template <class S, int SH = 0>
MYR_HDI double kkt_backsub(const Problem& P, const WS<S>& ws, const double* dl, double* dst) {
  using D = Dims<S>;
  constexpr int NW = S::NW;
  const double* const Hb = ws.Hinv;
  if (SH >= 2) MYR_ASSUME_SHARED(Hb);
  if (SH >= 1) MYR_ASSUME_SHARED(dl);   // dlam / dl2... only dlam belongs to the CR group: callers pass SH accordingly
  double dHd = 0.0;
  for (int q = MYR_TID; q < ws.Q; q += MYR_NT) {
    double u[NW];
    jt_times<S, SH>(P, ws, q, dl, u);
    const double* Hq = Hb + q * D::HS;
#pragma unroll
    for (int i = 0; i < NW; ++i) u[i] += NQ(rb, q, i);
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < NW; ++k) a += Hq[pidx(i, k, NW)] * u[k];
      dst[i * ws.ldq + q] = -a;
      dHd += a * u[i];   // a = -dz_i (zero on fixed variables: their rows of Hinv vanish)
    }
  }
  return dHd;
}

// Iterative refinement against the matrix WITHOUT delta_reg: the node blocks are factorised with a tiny
// regularisation (directions in which W + Sigma is singular, e.g. a state that enters neither cost nor dynamics and is
// far from its bounds, would otherwise make the block elimination break down although the KKT matrix is regular);
// the refinement removes its effect and recovers the digits the Schur complement loses.  Scratch: dzL (node residual),
// dl2 (multiplier correction), crb.
template <class S>
MYR_HDI void kkt_refine(const Problem& P, const WS<S>& ws, double delta_w, double delta_c, int max_refine, int& parity) {
  using D = Dims<S>;
  constexpr int NW = S::NW, NC = S::NC;
  const int Q = ws.Q, St = ws.St;
  for (int itr = 0; itr < max_refine; ++itr) {
    // node residual  rz = -rb - (H dz + G^T dlam_phi + F^T dlam_psi)   (stored in dzL), norms
    double rmax = 0.0, smax = 0.0;
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      double v[NW], d[NW], u[NW];
#pragma unroll
      for (int i = 0; i < NW; ++i) { v[i] = -NQ(rb, q, i); d[i] = NQ(dz, q, i); smax = fmax(smax, fabs(v[i])); }
      const double* Wq = ws.W + q * D::WSZ;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        double a = (NQ(sig, q, i) + delta_w) * d[i];
#pragma unroll
        for (int k = 0; k < NW; ++k) a += Wq[pidx(i, k, NW)] * d[k];
        v[i] -= a;
      }
      jt_times<S>(P, ws, q, ws.dlam, u);
      const uint32_t fm = ws.fix()[q];
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        const bool fx = (fm >> i) & 1u;
        const double r_ = fx ? 0.0 : v[i] - u[i];
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
        const double* Hq = ws.Hinv + q * D::HS;
        double t[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          double a = 0.0;
#pragma unroll
          for (int k2 = 0; k2 < NW; ++k2) a += Hq[pidx(i, k2, NW)] * NQ(dzL, q, k2);
          t[i] = a;
        }
        if constexpr (D::kAff) {
          constexpr int n = S::n, NG = S::NG;
          const double* Jq = ws.G + q * D::GS;
          double ap[NG], bp[NG], as[NG], bs[NG], Jd[n], Jtt[n];
          S::role_coefs(P, q, ap, bp, as, bs);
#pragma unroll
          for (int r = 0; r < n; ++r) {
            double a = 0.0, b_ = 0.0;
#pragma unroll
            for (int i = 0; i < NW; ++i) { a += Jq[r * NW + i] * NQ(dz, q, i); b_ += Jq[r * NW + i] * t[i]; }
            Jd[r] = a; Jtt[r] = b_;
          }
#pragma unroll
          for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int r = 0; r < n; ++r) {
              const double al = role ? as[g] : ap[g], be = role ? bs[g] : bp[g];
              rc[g * n + r] -= al * Jd[r] + be * NQ(dz, q, r);
              bj[g * n + r] += al * Jtt[r] + be * t[r];
            }
        } else {
          const double* Jq = (role ? ws.F : ws.G) + q * D::GS;
#pragma unroll
          for (int r = 0; r < NC; ++r) {
            double a = 0.0, b_ = 0.0;
#pragma unroll
            for (int i = 0; i < NW; ++i) { a += Jq[r * NW + i] * NQ(dz, q, i); b_ += Jq[r * NW + i] * t[i]; }
            rc[r] -= a; bj[r] += b_;
          }
        }
      }
#pragma unroll
      for (int r = 0; r < NC; ++r) {
        rmax = fmax(rmax, (rc[r] != rc[r]) ? INFINITY : fabs(rc[r]));
        ws.crb[j * NC + r] = bj[r] - rc[r];
      }
    }
    double rv[2] = {rmax, smax};
    block_reduce_multi<R_MAX, R_MAX>(rv, ws.red, parity);   // its barrier also publishes crb
    rmax = rv[0]; smax = rv[1];
    if (!(rmax > 1e-13 * fmax(1.0, smax)) || !isfinite(rmax)) break;
    block_cr_resolve<NC>(St, ws.crD, ws.crVL, ws.crVU, ws.crb, ws.dl2);
    for (int k = MYR_TID; k < St * NC; k += MYR_NT) ws.dlam[k] += ws.dl2[k];
    for (int q = MYR_TID; q < Q; q += MYR_NT) {
      double v[NW], u[NW];
      jt_times<S>(P, ws, q, ws.dl2, u);
#pragma unroll
      for (int i = 0; i < NW; ++i) v[i] = NQ(dzL, q, i) - u[i];
      const double* Hq = ws.Hinv + q * D::HS;
#pragma unroll
      for (int i = 0; i < NW; ++i) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < NW; ++k) a += Hq[pidx(i, k, NW)] * v[k];
        NQ(dz, q, i) += a;
      }
    }
    MYR_SYNC();
  }
}

// complete KKT solve (factor, multipliers, primal step, optional refinement): what myr_kkt_solve exposes
template <class S, int SH = 0>
MYR_HDI bool kkt_solve(const Problem& P, const WS<S>& ws, double delta_w, double delta_c, double delta_reg, int max_refine, int& parity,
                       double* dHd_part = nullptr, double mu_apply = -1.0) {
  double minpr;
  const bool ok = kkt_factor<S, SH>(P, ws, delta_w, delta_c, delta_reg, minpr, parity, mu_apply);
  if (!ok) return false;
  const double dHd = kkt_backsub<S, SH>(P, ws, ws.dlam, ws.dz);
  if (dHd_part) *dHd_part = dHd;
  MYR_SYNC();
  // refinement is only worth its cost when some node block was close to singular
  if (max_refine > 0 && minpr < 1e-4) kkt_refine<S>(P, ws, delta_w, delta_c, max_refine, parity);
  return true;
}

// Second-order-correction solve (IPOPT A-5.5 ff.) with the factors the last kkt_solve left behind: same matrix,
// constraint right-hand side csoc instead of c.   S dl2 = csoc - J Hinv rb,   dz2 = -Hinv (rb + J^T dl2).
template <class S>
MYR_HDI void kkt_soc_solve(const Problem& P, const WS<S>& ws) {
  constexpr int NC = S::NC;
  for (int k = MYR_TID; k < ws.St * NC; k += MYR_NT) ws.crb[k] = ws.csoc[k] + ws.sch[k];
  MYR_SYNC();
  block_cr_resolve<NC>(ws.St, ws.crD, ws.crVL, ws.crVU, ws.crb, ws.dl2);
  kkt_backsub<S>(P, ws, ws.dl2, ws.dz2);
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
};

// per-instance pointers of the NLP the IPM iterates on (the reference NLP itself for collocation, the lifted one for
// shooting), in that NLP's flat layouts (S::zidx / S::cidx); nv / ncn are that NLP's sizes
struct InstPtrs {
  const double* z0; const double* lb; const double* ub;
  double* z; double* lam; double* zL; double* zU;
  int nv, ncn;
};
struct InstResult { double f, E0, cinf; int status, iters; };

// host twin: MYR_TRACE=1 prints one line per interior-point iteration (debugging aid for the tests); the device prints
// the same line only in -DMYR_TRACE_DEVICE builds (tools/build_variant.sh)
#ifndef __CUDA_ARCH__
inline bool trace_enabled() { static const bool on = getenv("MYR_TRACE") != nullptr; return on; }
#elif defined(MYR_TRACE_DEVICE)
__device__ __forceinline__ bool trace_enabled() { return true; }
#endif

// watchdog: store (save = true) or restore the iterate (z, lambda, z_L, z_U); out of line -- it runs once in thousands of
// iterations and must not cost the iteration loop registers or instruction-cache footprint
template <class S>
MYR_HDN void watchdog_copy(const WS<S>& ws, int ncn, bool save) {
  MYR_FOR_VARS(it) {
    const int q = it.q, i = it.i;
    if (save) { NQ(wz, q, i) = NQ(z, q, i); NQ(wzL, q, i) = NQ(zL, q, i); NQ(wzU, q, i) = NQ(zU, q, i); }
    else { NQ(z, q, i) = NQ(wz, q, i); NQ(zL, q, i) = NQ(wzL, q, i); NQ(zU, q, i) = NQ(wzU, q, i); }
  }
  for (int k = MYR_TID; k < ncn; k += MYR_NT) {
    if (save) ws.wlam[k] = ws.lam[k]; else ws.lam[k] = ws.wlam[k];
  }
}

template <class S>
MYR_HDI InstResult ipm_solve_instance(const Problem& P, const IpmOpts& O, const InstPtrs& ip, const WS<S>& ws) {
  constexpr int NW = S::NW, NC = S::NC;
  const int Q = ws.Q, St = ws.St;
  const int ncn = St * NC;
  int parity = 0;

  // ---- initial point: push into the (relaxed) box, unit bound multipliers, zero lambda
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
    uint32_t fm = 0;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int id = S::zidx(P, q, i);
      const double lb = ip.lb[id], ub = ip.ub[id];
      const Bnd bd = make_bnd(lb, ub, O.bound_relax);
      double x = ip.z0[id];
      if (bd.fixed) { x = lb; fm |= (1u << i); }
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
      NQ(z, q, i) = x;
      NQ(zL, q, i) = bd.hasL ? 1.0 : 0.0;
      NQ(zU, q, i) = bd.hasU ? 1.0 : 0.0;
      NQ(lbr, q, i) = bd.fixed ? lb : (bd.hasL ? bd.lbr : -INFINITY);
      NQ(ubr, q, i) = bd.fixed ? lb : (bd.hasU ? bd.ubr : INFINITY);
    }
    ws.fix()[q] = fm;
  }
  for (int k = MYR_TID; k < ncn; k += MYR_NT) ws.lam[k] = 0.0;
  MYR_SYNC();

  MYR_PH_DECL
  double mu = O.mu_init, nu = 1.0, delta_last = 0.0;
  int it = 0, status = ST_MAXITER, n_acceptable = 0, hard_iters = 0;
  bool soc_armed = false;
  double f = 0.0, E0 = INFINITY, cinf = INFINITY, c1 = 0.0;
  const double mu_floor = fmax(O.mu_min, O.tol / 10.0);
  const double inv_kappa_sigma = 1.0 / O.kappa_sigma;
  // Line-search filter (Waechter & Biegler 2006, section 2.3; IPOPT's defaults): pairs (theta, phi) = (sum |c|, barrier
  // objective) that a trial point must improve on.  A ring of kFilterMax entries, emptied whenever mu changes (phi
  // depends on mu).  All threads carry the same count; thread 0 writes the entries (accept phase, before its barrier).
  constexpr double kGammaTheta = 1e-5, kGammaPhi = 1e-8, kEtaPhi = 1e-8, kSTheta = 1.1, kSPhi = 2.3;
#ifdef __CUDA_ARCH__
  double* const filt = ws.filt;
#else
  double filt_host[2 * kFilterMax];
  double* const filt = filt_host;
#endif
  int nfilt = 0;
  double theta_max = -1.0, theta_min = -1.0, mu_filter = mu;
  // Watchdog (IPOPT: watchdog_shortened_iter_trigger = 10, watchdog_trial_iter_max = 3).  A run of shortened steps can
  // go on for hundreds of iterations when the full step is a good one that neither the merit function nor the filter's
  // objective-type test will take (one CARTPOLE start state in 65 536: 432 iterations of 1-3 % steps, then three full
  // steps to 1e-8).  After kWdTrigger shortened iterations in a row the iterate is stored and full steps are taken
  // without a test; a later trial point that is acceptable RELATIVE TO THE STORED ITERATE ends the procedure, kWdMaxIter
  // steps without one restore the stored iterate and the regular line search carries on (with a cool-down).
  // Both safeguards (watchdog, regularised retry after a failed line search) are compiled into the plain collocation
  // kernels only: the lifted-shooting and NODE kernels are the most register-starved ones and lost 6-12 % to the extra
  // live state, for instance classes on which neither event has been observed.
  constexpr bool kSafeguards = !scheme_is_lifted<S>::value && !Layout<S>::kCoopMlp;
  constexpr int kWdTrigger = 10, kWdMaxIter = 3, kWdCooldown = 10;
  int wd_short = 0, wd_iter = 0, wd_cool = 0;
  double delta_force = 0.0;   // > 0: the line search of this iterate failed, factorise with at least this regularisation
  int ls_retries = 0;
  double wd_c1 = 0.0, wd_bphi = 0.0, wd_mu = 0.0, wd_nu = 0.0;

  while (true) {
    // ---------------- K1: evaluate with derivatives; rb = grad f + J^T lam
    MYR_PH(9);
    double fpart = eval_nodes<S, 2>(P, ws, ws.z);
    MYR_SYNC();
    MYR_PH(0);
    // ---------------- constraints, dual residual, complementarity, scaling sums: one fused reduction
    double cmx = 0.0, csm = 0.0, suml = 0.0, cmag = 0.0;
    stage_constraints<S>(P, ws, ws.c, cmx, csm, &cmag);
    for (int k = MYR_TID; k < ncn; k += MYR_NT) suml += fabs(ws.lam[k]);
    double rdmax = 0.0, szmax = -INFINITY, szmin = INFINITY, sumz = 0.0, nbnd = 0.0, slog = 0.0;
    {
      LogProd lpq;
      MYR_FOR_VARS(it) {
        const int q = it.q, i = it.i;
        const bool fixed = (ws.fix()[q] >> i) & 1u;
        // all loads first, unconditionally: vectors that live in global memory cost one round trip, not one per branch
        const double lo = NQ(lbr, q, i), hi = NQ(ubr, q, i), x = NQ(z, q, i), zl = NQ(zL, q, i), zu = NQ(zU, q, i), r0 = NQ(rb, q, i);
        double rd = 0.0, r1 = 0.0, r2 = 0.0;
        if (!fixed) {
          rd = r0 - zl + zu;
          if (lo > -INFINITY) { const double sl = x - lo; const double pz = sl * zl; szmax = fmax(szmax, pz); szmin = fmin(szmin, pz); sumz += zl; nbnd += 1.0; lpq.mul(sl); r1 = pivot_rcp(sl); }
          if (hi < INFINITY) { const double su = hi - x; const double pz = su * zu; szmax = fmax(szmax, pz); szmin = fmin(szmin, pz); sumz += zu; nbnd += 1.0; lpq.mul(su); r2 = pivot_rcp(su); }
        }
        // reciprocal slacks for the barrier terms (node phase of the KKT solve) and the step-size phase: they do not depend
        // on mu, so this pass -- which has every operand in registers anyway -- produces them
        NQ(rsl, q, i) = r1; NQ(rsu, q, i) = r2;
        const double a = (rd != rd) ? INFINITY : fabs(rd);
        rdmax = fmax(rdmax, a);
      }
      slog = lpq.value();
    }
    {
      double rv[11] = {fpart, cmx, csm, rdmax, szmax, szmin, sumz, nbnd, slog, suml, cmag};
      block_reduce_multi<R_SUM, R_MAX, R_SUM, R_MAX, R_MAX, R_MIN, R_SUM, R_SUM, R_SUM, R_SUM, R_SUM>(rv, ws.red, parity);
      f = rv[0]; cinf = rv[1]; c1 = rv[2]; rdmax = rv[3]; szmax = rv[4]; szmin = rv[5]; sumz = rv[6]; nbnd = rv[7]; slog = rv[8]; suml = rv[9];
      cmag = rv[10];
    }
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
    if (theta_max < 0.0) { theta_max = 1e4 * fmax(1.0, c1); theta_min = 1e-4 * fmax(1.0, c1); }
    if (mu != mu_filter) { nfilt = 0; mu_filter = mu; }

    MYR_PH(1);
    // (the barrier gradient  rb -= mu / s_L - mu / s_U  and  Sigma = z_L / s_L + z_U / s_U  are formed per node inside the
    // KKT node phase, first factorisation attempt only: kkt_factor, mu_apply)
    MYR_PH(2);
    // ---------------- K2: KKT solve with inertia correction (IPOPT Algorithm IC)
    double delta = kSafeguards ? delta_force : 0.0, dHd = 0.0;   // dHd: this thread's part of dz^T H dz, from the back-substitution
    bool ok = false;
    for (int tries = 0; tries < 60; ++tries) {
      // accurate steps only matter near the solution: refine the linear solve in the end game only
      const int refine = E0 < 1e-3 ? O.max_refine : 0;
#if defined(__CUDA_ARCH__) && defined(MYR_FORCE_SH)
      ok = kkt_solve<S, MYR_FORCE_SH>(P, ws, delta, O.delta_c, O.delta_reg, refine, parity, &dHd, tries == 0 ? mu : -1.0);   // experiment: one variant only
#elif defined(__CUDA_ARCH__) && defined(MYR_DISPATCH_SH)
      if (ws.sh == 2) ok = kkt_solve<S, 2>(P, ws, delta, O.delta_c, O.delta_reg, refine, parity, &dHd, tries == 0 ? mu : -1.0);
      else if (ws.sh == 1) ok = kkt_solve<S, 1>(P, ws, delta, O.delta_c, O.delta_reg, refine, parity, &dHd, tries == 0 ? mu : -1.0);
      else ok = kkt_solve<S, 0>(P, ws, delta, O.delta_c, O.delta_reg, refine, parity, &dHd, tries == 0 ? mu : -1.0);
#else
      ok = kkt_solve<S, 0>(P, ws, delta, O.delta_c, O.delta_reg, refine, parity, &dHd, tries == 0 ? mu : -1.0);
#endif
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
    double m_pr = 0.0, m_du = 0.0, dphi = 0.0;
    MYR_FOR_VARS(it) {
      const int q = it.q, i = it.i;
      const bool fixed = (ws.fix()[q] >> i) & 1u;
      const double d = NQ(dz, q, i), r1 = NQ(rsl, q, i), r2 = NQ(rsu, q, i), zl = NQ(zL, q, i), zu = NQ(zU, q, i), g0 = NQ(gl, q, i);
      double dl = 0.0, du = 0.0;
      if (!fixed) {
        // barrier-objective gradient = grad f - mu / s_L + mu / s_U
        dphi += d * (g0 - mu * r1 + mu * r2);
        if (r1 != 0.0) {
          dl = mu * r1 - zl - zl * r1 * d;
          m_pr = fmax(m_pr, -d * r1);
          if (dl < 0.0) m_du = fmax(m_du, -dl * pivot_rcp(zl));
        }
        if (r2 != 0.0) {
          du = mu * r2 - zu + zu * r2 * d;
          m_pr = fmax(m_pr, d * r2);
          if (du < 0.0) m_du = fmax(m_du, -du * pivot_rcp(zu));
        }
      }
      NQ(dzL, q, i) = dl;
      NQ(dzU, q, i) = du;
    }
    {
      double rv[4] = {m_pr, m_du, dphi, dHd};
      block_reduce_multi<R_MAX, R_MAX, R_SUM, R_SUM>(rv, ws.red, parity);
      m_pr = rv[0]; m_du = rv[1]; dphi = rv[2]; dHd = rv[3];
    }
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
    const double bphi0 = f - mu * slog;   // barrier objective of the iterate (filter)
    const double sw_lhs = (O.use_filter && dphi < 0.0 && c1 <= theta_min) ? pow(-dphi, kSPhi) : 0.0;
    const double sw_rhs = (O.use_filter && dphi < 0.0 && c1 <= theta_min) ? pow(c1, kSTheta) : INFINITY;
    bool ftype = false;
    double alpha = a_pr;
    bool accepted = false;
    bool wd_rollback = false, wd_free = false;   // wd_free: the accepted step was a watchdog step (no filter entry for it)
    if (kSafeguards && wd_iter > 0 && mu != wd_mu) { wd_iter = 0; wd_short = 0; }   // a barrier update ends the watchdog
    if (kSafeguards && O.use_watchdog && wd_iter == 0 && wd_cool == 0 && wd_short >= kWdTrigger) {
      // store the iterate and its reference values; the trial loop below then runs in watchdog mode
      watchdog_copy<S>(ws, ncn, true);
      wd_c1 = c1; wd_bphi = bphi0; wd_mu = mu; wd_nu = nu;
      wd_iter = 1;
    }
    if (kSafeguards && wd_cool > 0) --wd_cool;
    double f_t = 0.0;
    // Armijo with IPOPT's rounding-error relaxation (10 eps |phi|) so that converged iterates are not rejected by
    // cancellation.  The infeasibility term gets its own allowance: sum |c| is a sum of differences of O(|x|) role values
    // and cannot be evaluated more accurately than eps * (sum of their magnitudes); with states of order 1e3 (SEIR) that
    // noise times nu is far above 10 eps |phi| and made converged iterates stall a hair above the tolerance.
    const double armijo_slack = 2.220446049250313e-16 * (10.0 * fabs(phi0) + 2.0 * nu * cmag);
    // One loop serves the backtracking trials (soc == false: step dz, length alpha) and the second-order-correction
    // trials (soc == true: step dz2 from kkt_soc_solve, length a2), so the trial evaluation exists once in the code.
    int ls = 0, ks = 0, ls_used = 0;
    bool soc = false;
    double a2 = 1.0, c1_prev = 0.0;
    while (ls < O.max_ls) {
      if (soc) {
        kkt_soc_solve<S>(P, ws);
        double m2 = 0.0;
        MYR_FOR_VARS(it) {
          const double d = NQ(dz2, it.q, it.i);
          m2 = fmax(m2, fmax(-d * NQ(rsl, it.q, it.i), d * NQ(rsu, it.q, it.i)));
        }
        double rv[1] = {m2};
        block_reduce_multi<R_MAX>(rv, ws.red, parity);
        m2 = rv[0];
        a2 = m2 > tau ? tau / m2 : 1.0;
      }
      const double a = soc ? a2 : alpha;
      const double* step = soc ? ws.dz2 : ws.dz;
      ls_used = ls;
      // ---- trial point z + a * step (formed per node inside the evaluation, never stored), its barrier log-sum,
      // objective and constraints
      double blog = 0.0;
      const double ft_part = eval_nodes<S, 0>(P, ws, ws.z, step, a, &blog);
      MYR_SYNC();
      double ci_t = 0.0, c1_t = 0.0;
      stage_constraints<S>(P, ws, ws.ct, ci_t, c1_t);
      {
        double rv[4] = {ft_part, ci_t, c1_t, blog};
        block_reduce_multi<R_SUM, R_MAX, R_SUM, R_SUM>(rv, ws.red, parity);
        f_t = rv[0]; ci_t = rv[1]; c1_t = rv[2]; blog = rv[3];
      }
      const double phit = f_t - mu * blog + nu * c1_t;
      if (isfinite(phit) && phit <= phi0 + O.eta * alpha * Dm + armijo_slack) { accepted = true; break; }
      // Not enough decrease of the l1 merit function.  IPOPT's filter rules accept the point all the same when it makes
      // sufficient progress in EITHER the infeasibility or the barrier objective (and is not dominated by an earlier
      // iterate): far from the solution the merit test rejects long steps that cut the infeasibility because the barrier
      // objective rises, and the iteration crawls with steps of a few percent (CARTPOLE: 27 of one instance's 41
      // iterations).  The merit test above stays as the alternative that needs no restoration phase.
      if (O.use_filter && isfinite(phit) && c1_t <= theta_max) {
        const double bphi_t = f_t - mu * blog;
        bool fok = true;
        const int nf = nfilt < kFilterMax ? nfilt : kFilterMax;
        for (int j = 0; j < nf; ++j)
          if (c1_t >= (1.0 - kGammaTheta) * filt[2 * j] && bphi_t >= filt[2 * j + 1] - kGammaPhi * filt[2 * j]) fok = false;
        if (fok) {
          if (c1 <= theta_min && sw_lhs * alpha > sw_rhs) {   // switching condition: an objective-type step needs Armijo on phi
            if (bphi_t <= bphi0 + kEtaPhi * alpha * dphi + armijo_slack) { accepted = true; ftype = true; break; }
          } else if (c1_t <= (1.0 - kGammaTheta) * c1 || bphi_t <= bphi0 - kGammaPhi * c1) { accepted = true; break; }
        }
      }
      if (kSafeguards && wd_iter > 0 && !soc && ls == 0 && isfinite(phit)) {
        // watchdog mode and the full step failed the regular tests against the CURRENT iterate: judge it against the stored
        // one (sufficient progress in infeasibility or barrier objective); failing that, take it anyway -- or give up
        const double bphi_t = f_t - mu * blog;
        if (wd_iter > 1 && (c1_t <= (1.0 - kGammaTheta) * wd_c1 || bphi_t <= wd_bphi - kGammaPhi * wd_c1)) { wd_iter = 0; wd_short = 0; accepted = true; break; }
        if (wd_iter > kWdMaxIter) { wd_rollback = true; break; }
        ++wd_iter; wd_free = true; accepted = true; break;
      }
      if (!soc) {
        // second-order correction (IPOPT A-5.5 .. A-5.9): the full step was rejected without reducing the infeasibility,
        // typically because the constraint curvature along dz outweighs the predicted decrease (Maratos effect).
        // Re-solve with the same factors and the constraint right-hand side  alpha c + c(z + alpha dz).
        // It is armed per instance only after two consecutive iterations that needed >= 4 step halvings: on the well
        // behaved instances (the bulk of a batch) an unconditional SOC costs more re-solves than it saves iterations
        // (CARTPOLE N=100: -12 % iterations, +16 % time), on the hard ones it is what makes the method converge.
        if (ls == 0 && soc_armed && O.max_soc > 0 && c1_t >= c1) {
          for (int k = MYR_TID; k < ncn; k += MYR_NT) ws.csoc[k] = alpha * ws.c[k] + ws.ct[k];
          MYR_SYNC();
          soc = true; ks = 0; c1_prev = c1_t;
          continue;
        }
        alpha *= 0.5; ++ls;
      } else {
        // kappa_soc: the correction must keep reducing the infeasibility, at most max_soc times
        if (!(c1_t <= 0.99 * c1_prev) || ++ks >= O.max_soc) { soc = false; alpha *= 0.5; ++ls; continue; }
        c1_prev = c1_t;
        for (int k = MYR_TID; k < ncn; k += MYR_NT) ws.csoc[k] = a2 * ws.csoc[k] + ws.ct[k];
        MYR_SYNC();
      }
    }
    if (kSafeguards && wd_rollback) {   // the watchdog steps led nowhere: back to the stored iterate, regular line search from there
      watchdog_copy<S>(ws, ncn, false);
      nu = wd_nu; wd_iter = 0; wd_short = 0; wd_cool = kWdCooldown;
      MYR_SYNC();
      continue;
    }
    const double* dl_acc = ws.dlam;
    const double* step_acc = ws.dz;
    if (accepted && soc) {  // the multiplier step of the corrected system goes with the corrected primal step
      dl_acc = ws.dl2; step_acc = ws.dz2;
      alpha = a2;
    }
#if !defined(__CUDA_ARCH__) || defined(MYR_TRACE_DEVICE)
    if (MYR_TID == 0 && trace_enabled())
      printf("[ipm] it %3d f %.12e E0 %.3e cinf %.2e rd %.2e mu %.1e delta %.1e a_pr %.3e a_du %.3e alpha %.3e ls %d soc %d acc %d\n",
              it, f, E0, cinf, rdmax / sd, mu, delta, a_pr, a_du, alpha, ls_used, (int)soc, (int)accepted);
#endif
    hard_iters = ls_used >= 4 ? hard_iters + 1 : 0;
#ifndef MYR_SOC_ARM_AFTER
#define MYR_SOC_ARM_AFTER 2
#endif
    if (hard_iters >= MYR_SOC_ARM_AFTER) soc_armed = true;
    if (!accepted) {
      // No step length was acceptable.  Far from the solution that almost always means the step is not a descent
      // direction because the inertia test passed on rounding noise (nearly singular pivot blocks; seen on 1 of 8192
      // Hermite-Simpson instances, and only in some builds): redo the iteration with a larger regularisation -- a few
      // times -- before giving up.
      if (kSafeguards && E0 > O.acceptable_tol && ls_retries < 3 && delta < O.delta_max) {
        delta_force = fmax(delta, O.delta_0) * O.kappa_w_plus_first;
        ++ls_retries;
        continue;
      }
      status = (E0 <= O.acceptable_tol) ? ST_ACCEPTABLE : ST_LINESEARCH; break;
    }
    if (kSafeguards) { delta_force = 0.0; ls_retries = 0; }

    MYR_PH(5);
    // ---------------- accept: primal, equality multipliers, bound multipliers (with the kappa_sigma safeguard)
    MYR_FOR_VARS(it) {
      const int q = it.q, i = it.i;
      if ((ws.fix()[q] >> i) & 1u) continue;
      const double lo = NQ(lbr, q, i), hi = NQ(ubr, q, i), x = NQ(z, q, i) + alpha * step_acc[i * ws.ldq + q];   // the accepted trial point
      const double zl = NQ(zL, q, i), zu = NQ(zU, q, i), dl = NQ(dzL, q, i), du = NQ(dzU, q, i);
      NQ(z, q, i) = x;
      if (lo > -INFINITY) {
        const double ms = mu * pivot_rcp(x - lo);
        NQ(zL, q, i) = fmax(fmin(zl + a_du * dl, O.kappa_sigma * ms), ms * inv_kappa_sigma);
      }
      if (hi < INFINITY) {
        const double ms = mu * pivot_rcp(hi - x);
        NQ(zU, q, i) = fmax(fmin(zu + a_du * du, O.kappa_sigma * ms), ms * inv_kappa_sigma);
      }
    }
    for (int k = MYR_TID; k < ncn; k += MYR_NT) ws.lam[k] += alpha * dl_acc[k];
    if (kSafeguards) {
      wd_short = (ls_used >= 1 && !wd_free) ? wd_short + 1 : (wd_iter > 0 ? wd_short : 0);
      if (wd_iter > 0 && !wd_free && ls_used == 0) { wd_iter = 0; wd_short = 0; }   // a regular full step: the watchdog is done
    }
    if (O.use_filter && !ftype && !(kSafeguards && wd_free)) {   // the iterate just left joins the filter (unless the step was an objective-type step)
      if (MYR_TID == 0) {
        const int pos = nfilt % kFilterMax;
        filt[2 * pos] = (1.0 - kGammaTheta) * c1; filt[2 * pos + 1] = bphi0 - kGammaPhi * c1;
      }
      ++nfilt;
    }
    MYR_SYNC();
    ++it;
  }
  // ---- results in the NLP's flat layouts
  for (int q = MYR_TID; q < Q; q += MYR_NT) {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int id = S::zidx(P, q, i);
      ip.z[id] = NQ(z, q, i); ip.zL[id] = NQ(zL, q, i); ip.zU[id] = NQ(zU, q, i);
    }
  }
  for (int j = MYR_TID; j < St; j += MYR_NT)
#pragma unroll
    for (int r = 0; r < NC; ++r) ip.lam[S::cidx(P, j, r)] = NS(lam, j, r);
  MYR_SYNC();
  InstResult res;
  res.f = f; res.E0 = E0; res.cinf = cinf; res.status = status; res.iters = it;
  return res;
}

// Collocation: the IPM works directly on the caller's arrays.
template <class S>
MYR_HDI typename std::enable_if<!scheme_is_lifted<S>::value>::type
ipm_solve_entry(const Problem& P, const IpmOpts& O, const IpmIO& io, int b, const WS<S>& ws) {
  const long long nv = P.nvars, nc = P.ncon;
  InstPtrs ip{io.z0 + b * nv, io.lb + b * nv, io.ub + b * nv, io.z + b * nv, io.lam + b * nc, io.zL + b * nv, io.zU + b * nv,
              P.nvars, P.ncon};
  const InstResult r = ipm_solve_instance<S>(P, O, ip, ws);
  if (MYR_TID == 0) {
    io.obj[b] = r.f; io.kkt_err[b] = r.E0; io.con_inf[b] = r.cinf; io.status[b] = r.status; io.iters[b] = r.iters;
  }
}

// Lifted shooting: expand the reference NLP data into the lifted NLP, solve, map the solution back and report the
// objective / constraint violation of the REFERENCE NLP (rollouts from the interval-start states).
template <class S>
MYR_HDI typename std::enable_if<scheme_is_lifted<S>::value>::type
ipm_solve_entry(const Problem& P, const IpmOpts& O, const IpmIO& io, int b, const WS<S>& ws) {
  constexpr int n = S::n, m = S::m, NW = S::NW, NC = S::NC;
  const int Q = ws.Q, St = ws.St;
  const int nvI = Q * NW, ncI = St * NC;
  const long long nv = P.nvars, nc = P.ncon;
  double* zI0 = ws.ext; double* lbI = zI0 + nvI; double* ubI = lbI + nvI; double* zI = ubI + nvI;
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
  InstPtrs ip{zI0, lbI, ubI, zI, lamI, zLI, zUI, nvI, ncI};
  InstResult r = ipm_solve_instance<S>(Pi, O, ip, ws);
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
  int parity = 0;
  double rv[2] = {fs, cm};
  block_reduce_multi<R_SUM, R_MAX>(rv, ws.red, parity);
  MYR_SYNC();
  if (MYR_TID == 0) {
    io.obj[b] = rv[0]; io.kkt_err[b] = r.E0; io.con_inf[b] = rv[1]; io.status[b] = r.status; io.iters[b] = r.iters;
  }
}

}  // namespace myr
