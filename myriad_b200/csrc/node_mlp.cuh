// NODE dynamics: the MLP of myriad/neural_ode/create_node.py:110-117 (hk.Linear = x @ w + b, sigmoid after every hidden
// layer, linear output) as used by NodeSystem.parametrized_dynamics (myriad/systems/neural_ode/node_system.py:36-38) and
// by plan_with_node_model (myriad/utils.py:230-242), with its input Jacobian and the multiplier-contracted Hessian
//     H = hess_v ( mu . y(v) ) = sum_j  Zdot_j^T diag( hbar_j * sigma''(z_j) ) Zdot_j
// (z_j pre-activations of hidden layer j, Zdot_j = d z_j / d v, hbar_j = d (mu . y) / d h_j), which needs one forward
// value pass, one reverse pass and NW forward tangent passes -- all matrix products with the layer weights.
//
// Two implementations of the same arithmetic:
//   * MlpScalar: plain loops, one thread per evaluation.  Host twin, shooting (where the stage points of a step depend
//     on each other) and the reference for the tensor-core path.
//   * mlp_nodes_pass (device only): the whole CTA evaluates ALL nodes of one instance, eight nodes at a time, on the
//     fp64 tensor cores (mma.sync m8n8k4 f64 = DMMA; tcgen05 has no fp64 kind, and the reference is fp64 end to end:
//     run.py:15).  Tiles are "one column per node": the value tile holds h_j of 8 nodes, tangent tile i holds
//     d h_j / d v_i of the same 8 nodes, so every elementwise combination (sigma', sigma'' scalings, Hessian terms) is
//     between identical fragment positions of different tiles and stays in registers.
#pragma once
#include "common.cuh"

namespace myr {

// sizes / offsets are filled by make_problem (kernels.cuh)
MYR_HDI double sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }

// ------------------------------------------------------------------ scalar reference implementation
template <int NIN, int NOUT>
struct MlpScalar {
  static constexpr int NWP = NIN * (NIN + 1) / 2;

  MYR_HDN static void f(const MlpDesc& M, const double* v, double* y) {
    double a[2][kMaxMlpWidth];
    int cur = 0;
    const int Lh = M.L - 1;
    for (int j = 0; j <= Lh; ++j) {
      const int in = M.size[j], out = M.size[j + 1];
      const double* w = M.theta + M.woff[j];
      const double* b = M.theta + M.boff[j];
      for (int o = 0; o < out; ++o) {
        double s = b[o];
        for (int i = 0; i < in; ++i) s += (j == 0 ? v[i] : a[cur][i]) * w[i * out + o];
        if (j < Lh) a[cur ^ 1][o] = sigmoid(s); else y[o] = s;
      }
      cur ^= 1;
    }
  }

  // MODE 1: y, J (NOUT x NIN row-major).  MODE 2: additionally H (packed upper NIN) += hess_v(mu . y)
  template <int MODE>
  MYR_HDN static void run(const MlpDesc& M, const double* v, const double* mu, double* y, double* J, double* H) {
    const int Lh = M.L - 1;
    double a[kMaxMlpLayers - 1][kMaxMlpWidth];   // hidden activations h_1..h_Lh
    double d[kMaxMlpLayers - 1][kMaxMlpWidth];   // hbar_j * sigma''(z_j)
    for (int j = 0; j < Lh; ++j) {
      const int in = M.size[j], out = M.size[j + 1];
      const double* w = M.theta + M.woff[j];
      const double* b = M.theta + M.boff[j];
      for (int o = 0; o < out; ++o) {
        double s = b[o];
        for (int i = 0; i < in; ++i) s += (j == 0 ? v[i] : a[j - 1][i]) * w[i * out + o];
        a[j][o] = sigmoid(s);
      }
    }
    {
      const int in = M.size[Lh];
      const double* w = M.theta + M.woff[Lh];
      const double* b = M.theta + M.boff[Lh];
      for (int o = 0; o < NOUT; ++o) {
        double s = b[o];
        for (int i = 0; i < in; ++i) s += a[Lh - 1][i] * w[i * NOUT + o];
        y[o] = s;
      }
    }
    if (MODE == 2) {
      double hb[2][kMaxMlpWidth];
      int cur = 0;
      for (int j = Lh; j >= 1; --j) {   // hbar_j = W_j (j == Lh ? mu : zbar_{j+1})
        const int rows = M.size[j], k = M.size[j + 1];
        const double* w = M.theta + M.woff[j];
        for (int i = 0; i < rows; ++i) {
          double s = 0.0;
          for (int o = 0; o < k; ++o) s += w[i * k + o] * (j == Lh ? mu[o] : hb[cur][o]);
          const double av = a[j - 1][i], s1 = av * (1.0 - av);
          d[j - 1][i] = s * s1 * (1.0 - 2.0 * av);
          hb[cur ^ 1][i] = s * s1;
        }
        cur ^= 1;
      }
    }
    // forward tangents, one input direction at a time would need the Hessian's cross terms: keep all NIN directions
    double t[2][NIN][kMaxMlpWidth];
    int cur = 0;
    for (int j = 0; j < Lh; ++j) {
      const int in = M.size[j], out = M.size[j + 1];
      const double* w = M.theta + M.woff[j];
      for (int o = 0; o < out; ++o) {
        double zt[NIN];
        for (int i = 0; i < NIN; ++i) {
          double s = 0.0;
          if (j == 0) s = w[i * out + o];
          else for (int k = 0; k < in; ++k) s += w[k * out + o] * t[cur][i][k];
          zt[i] = s;
        }
        if (MODE == 2) {
          const double dv = d[j][o];
          for (int i = 0; i < NIN; ++i)
            for (int k = i; k < NIN; ++k) H[pidx(i, k, NIN)] += dv * zt[i] * zt[k];
        }
        const double av = a[j][o], s1 = av * (1.0 - av);
        for (int i = 0; i < NIN; ++i) t[cur ^ 1][i][o] = s1 * zt[i];
      }
      cur ^= 1;
    }
    {
      const int in = M.size[Lh];
      const double* w = M.theta + M.woff[Lh];
      for (int o = 0; o < NOUT; ++o)
        for (int i = 0; i < NIN; ++i) {
          double s = 0.0;
          for (int k = 0; k < in; ++k) s += w[k * NOUT + o] * t[cur][i][k];
          J[o * NIN + i] = s;
        }
    }
  }
};

// ------------------------------------------------------------------ NodeSystem wrapper
template <class True>
struct SysNode {
  using TrueSystem = True;
  static constexpr int id = 100 + True::id, n = True::n, m = True::m, nw = True::nw, np = True::np;
  static constexpr bool time_dependent_cost = True::time_dependent_cost;
  static constexpr bool kNode = true;
  static constexpr bool has_terminal = True::has_terminal;
  MYR_HD static void terminal_coef(const double* p, double* tc) { True::terminal_coef(p, tc); }
  static constexpr const char* name = True::name;  // reported as NODE(<name>)
  MYR_HD static void default_params(double* p) { True::default_params(p); }
  // true cost (node_system.py:41-42)
  MYR_HD static double cost(const double* x, const double* u, double t, const double* p) { return True::cost(x, u, t, p); }
  MYR_HD static double cost_grad(const double* x, const double* u, double t, const double* p, double* g) { return True::cost_grad(x, u, t, p, g); }
  MYR_HD static double cost_grad_hess(const double* x, const double* u, double t, const double* p, double w, double* g, double* H) {
    return True::cost_grad_hess(x, u, t, p, w, g, H);
  }
};

template <class Sys, class = void>
struct sys_is_node { static constexpr bool value = false; };
template <class Sys>
struct sys_is_node<Sys, typename std::enable_if<Sys::kNode>::type> { static constexpr bool value = true; };

// Dynamics dispatch used by every scheme: analytic generated code, or the MLP (scalar path).
template <class Sys>
MYR_HDI void dyn_f(const Problem& P, const double* x, const double* u, double* f) {
  if constexpr (sys_is_node<Sys>::value) {
    double v[Sys::nw];
#pragma unroll
    for (int i = 0; i < Sys::n; ++i) v[i] = x[i];
#pragma unroll
    for (int i = 0; i < Sys::m; ++i) v[Sys::n + i] = u[i];
    MlpScalar<Sys::nw, Sys::n>::f(P.mlp, v, f);
  } else {
    Sys::f(x, u, P.p, f);
  }
}
template <class Sys>
MYR_HDI void dyn_fjac(const Problem& P, const double* x, const double* u, double* f, double* J) {
  if constexpr (sys_is_node<Sys>::value) {
    double v[Sys::nw];
#pragma unroll
    for (int i = 0; i < Sys::n; ++i) v[i] = x[i];
#pragma unroll
    for (int i = 0; i < Sys::m; ++i) v[Sys::n + i] = u[i];
    MlpScalar<Sys::nw, Sys::n>::template run<1>(P.mlp, v, nullptr, f, J, nullptr);
  } else {
    Sys::fjac(x, u, P.p, f, J);
  }
}
template <class Sys>
MYR_HDI void dyn_fjac_hess(const Problem& P, const double* x, const double* u, const double* mu, double* f, double* J, double* H) {
  if constexpr (sys_is_node<Sys>::value) {
    double v[Sys::nw];
#pragma unroll
    for (int i = 0; i < Sys::n; ++i) v[i] = x[i];
#pragma unroll
    for (int i = 0; i < Sys::m; ++i) v[Sys::n + i] = u[i];
    MlpScalar<Sys::nw, Sys::n>::template run<2>(P.mlp, v, mu, f, J, H);
  } else {
    Sys::fjac_hess(x, u, P.p, mu, f, J, H);
  }
}

// Dynamics values precomputed for a node by the cooperative tensor-core pass (f == nullptr: evaluate directly).
struct PreDyn {
  const double* f = nullptr;   // [n]
  const double* J = nullptr;   // [n][nw]
  const double* H = nullptr;   // [nw (nw + 1) / 2]
};

// ------------------------------------------------------------------ tensor-core path (device)
MYR_HDI int mlp_round_up(int v, int k) { return (v + k - 1) / k * k; }

// doubles of shared scratch mlp_nodes_pass needs
template <class S>
MYR_HDI int mlp_scratch_doubles(const MlpDesc& M) {
  if (M.L == 0) return 0;
  const int Lh = M.L - 1, Hp = M.hp;
  return 8 * Hp * (2 * Lh + 2 + S::NW) + 8 * mlp_round_up(S::NW, 4) + 8 * mlp_round_up(S::n, 4) + 8 * S::NWP;
}

#ifdef __CUDA_ARCH__
__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

// Evaluate the MLP dynamics at every node of one instance.  MODE 0: values f; 1: + Jacobian; 2: + contracted Hessian
// (mu per node from S::node_mu and the multipliers lam).  Outputs: dynf [Q][n], dynJ [Q][n][NW], dynH [Q][NWP].
// scr: mlp_scratch_doubles<S>() doubles of shared memory.  Must be called by all threads of the CTA
// (blockDim.x multiple of 32; hidden widths <= 16 * warps).
// z(q, i): variable i of node q; lam(j, r): multiplier of row r of stage j (views over the caller's layout)
template <class S, int MODE, class ZView, class LamView>
__device__ __noinline__ void mlp_nodes_pass(const Problem& P, int Q, const ZView z, const LamView lam, double* dynf, double* dynJ,
                                            double* dynH, double* scr, const double* theta_override = nullptr) {
#ifdef __CUDA_ARCH__
  constexpr int NW = S::NW, n = S::n, NWP = S::NWP;
  constexpr int kRounds = 2;
  const MlpDesc& M = P.mlp;
  const double* const theta = theta_override ? theta_override : M.theta;   // weights: the caller's vector or a shared-memory copy
  const int Lh = M.L - 1, Hp = M.hp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int Kin = mlp_round_up(NW, 4), n4 = mlp_round_up(n, 4);
  double* act = scr;                         // [Lh][Hp][8]   h_j of the 8 nodes of the group
  double* dd = act + Lh * Hp * 8;            // [Lh][Hp][8]   hbar_j * sigma''(z_j)
  double* zb = dd + Lh * Hp * 8;             // [2][Hp][8]    reverse operand zbar_j
  double* tb = zb + 2 * Hp * 8;              // [NW][Hp][8]   tangent operand d h_j / d v_i
  double* in0 = tb + NW * Hp * 8;            // [Kin][8]
  double* mub = in0 + Kin * 8;               // [n4][8]
  double* hbuf = mub + n4 * 8;               // [8][NWP]
  for (int q0 = 0; q0 < Q; q0 += 8) {
    // ---- inputs of the group
    for (int e = tid; e < Kin * 8; e += blockDim.x) {
      const int i = e >> 3, col = e & 7;
      const int q = min(q0 + col, Q - 1);
      in0[e] = i < NW ? z(q, i) : 0.0;
    }
    if (MODE == 2) {
      if (tid < 8) {
        const int q = min(q0 + tid, Q - 1);
        double mu[n];
        S::node_mu(P, q, lam, mu);
#pragma unroll
        for (int r = 0; r < n4; ++r) mub[r * 8 + tid] = r < n ? mu[r < n ? r : 0] : 0.0;
      }
      for (int e = tid; e < 8 * NWP; e += blockDim.x) hbuf[e] = 0.0;
    }
    __syncthreads();
    // ---- F: values
    for (int j = 0; j < Lh; ++j) {
      const int in = M.size[j], out = M.size[j + 1];
      const double* w = theta + M.woff[j];
      const double* b = theta + M.boff[j];
      const double* Bs = j == 0 ? in0 : act + (j - 1) * Hp * 8;
      const int KS = mlp_round_up(in, 4) >> 2, MT = mlp_round_up(out, 8) >> 3;
      for (int mt = warp; mt < MT; mt += nwarp) {
        const int o = 8 * mt + g;
        double c0 = o < out ? b[o] : 0.0, c1 = c0;
        _Pragma("unroll 4") for (int ks = 0; ks < KS; ++ks) {
          const int i = 4 * ks + t;
          const double a = (i < in && o < out) ? (*(w + i * out + o)) : 0.0;
          dmma_m8n8k4(c0, c1, a, Bs[i * 8 + g]);
        }
        double* dst = act + j * Hp * 8 + o * 8 + 2 * t;
        dst[0] = sigmoid(c0); dst[1] = sigmoid(c1);
      }
      __syncthreads();
    }
    {  // output layer: y = W_Lh^T h_Lh + b
      const int in = M.size[Lh];
      const double* w = theta + M.woff[Lh];
      const double* b = theta + M.boff[Lh];
      const double* Bs = act + (Lh - 1) * Hp * 8;
      const int KS = mlp_round_up(in, 4) >> 2, MT = mlp_round_up(n, 8) >> 3;
      for (int mt = warp; mt < MT; mt += nwarp) {
        const int o = 8 * mt + g;
        double c0 = o < n ? b[o] : 0.0, c1 = c0;
        _Pragma("unroll 4") for (int ks = 0; ks < KS; ++ks) {
          const int i = 4 * ks + t;
          const double a = (i < in && o < n) ? (*(w + i * n + o)) : 0.0;
          dmma_m8n8k4(c0, c1, a, Bs[i * 8 + g]);
        }
        if (o < n) {
          const int qa = q0 + 2 * t;
          if (qa < Q) dynf[qa * n + o] = c0;
          if (qa + 1 < Q) dynf[(qa + 1) * n + o] = c1;
        }
      }
    }
    if (MODE == 0) { __syncthreads(); continue; }
    // ---- R: adjoints hbar_j of mu . y, stored as dd_j = hbar_j sigma''(z_j) (and zbar_j = hbar_j sigma'(z_j) as operand)
    if (MODE == 2) {
      for (int j = Lh; j >= 1; --j) {
        const int rows = M.size[j], k = M.size[j + 1];
        const double* w = theta + M.woff[j];
        const double* Bs = j == Lh ? mub : zb + ((j + 1) & 1) * Hp * 8;
        const int KS = mlp_round_up(k, 4) >> 2, MT = mlp_round_up(rows, 8) >> 3;
        for (int mt = warp; mt < MT; mt += nwarp) {
          const int i = 8 * mt + g;
          double c0 = 0.0, c1 = 0.0;
          _Pragma("unroll 4") for (int ks = 0; ks < KS; ++ks) {
            const int o = 4 * ks + t;
            const double a = (i < rows && o < k) ? (*(w + i * k + o)) : 0.0;
            dmma_m8n8k4(c0, c1, a, Bs[o * 8 + g]);
          }
          const int pos = i * 8 + 2 * t;
          const double a0 = act[(j - 1) * Hp * 8 + pos], a1 = act[(j - 1) * Hp * 8 + pos + 1];
          const double s0 = a0 * (1.0 - a0), s1 = a1 * (1.0 - a1);
          dd[(j - 1) * Hp * 8 + pos] = c0 * s0 * (1.0 - 2.0 * a0);
          dd[(j - 1) * Hp * 8 + pos + 1] = c1 * s1 * (1.0 - 2.0 * a1);
          zb[(j & 1) * Hp * 8 + pos] = c0 * s0;
          zb[(j & 1) * Hp * 8 + pos + 1] = c1 * s1;
        }
        __syncthreads();
      }
    }
    // ---- T: tangents of all NW input directions, Hessian accumulation
    double hacc[MODE == 2 ? NWP : 1][2];
    if (MODE == 2) {
#pragma unroll
      for (int pp = 0; pp < NWP; ++pp) { hacc[pp][0] = 0.0; hacc[pp][1] = 0.0; }
    }
    for (int j = 0; j < Lh; ++j) {
      const int in = M.size[j], out = M.size[j + 1];
      const double* w = theta + M.woff[j];
      const int KS = mlp_round_up(in, 4) >> 2, MT = mlp_round_up(out, 8) >> 3;
      double c[kRounds][NW][2];
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int mt = warp + r * nwarp;
        if (mt < MT) {
          const int o = 8 * mt + g;
          if (j == 0) {
#pragma unroll
            for (int i = 0; i < NW; ++i) { const double v = o < out ? (*(w + i * out + o)) : 0.0; c[r][i][0] = v; c[r][i][1] = v; }
          } else {
#pragma unroll
            for (int i = 0; i < NW; ++i) { c[r][i][0] = 0.0; c[r][i][1] = 0.0; }
            _Pragma("unroll 4") for (int ks = 0; ks < KS; ++ks) {
              const int k = 4 * ks + t;
              const double a = (k < in && o < out) ? (*(w + k * out + o)) : 0.0;
#pragma unroll
              for (int i = 0; i < NW; ++i) dmma_m8n8k4(c[r][i][0], c[r][i][1], a, tb[i * Hp * 8 + k * 8 + g]);
            }
          }
          const int pos = o * 8 + 2 * t;
          const double a0 = act[j * Hp * 8 + pos], a1 = act[j * Hp * 8 + pos + 1];
          if (MODE == 2) {
            const double d0 = dd[j * Hp * 8 + pos], d1 = dd[j * Hp * 8 + pos + 1];
            int pp = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i)
#pragma unroll
              for (int k = i; k < NW; ++k, ++pp) {
                hacc[pp][0] += d0 * c[r][i][0] * c[r][k][0];
                hacc[pp][1] += d1 * c[r][i][1] * c[r][k][1];
              }
          }
          const double s0 = a0 * (1.0 - a0), s1 = a1 * (1.0 - a1);
#pragma unroll
          for (int i = 0; i < NW; ++i) { c[r][i][0] *= s0; c[r][i][1] *= s1; }
        }
      }
      __syncthreads();  // every warp has finished reading tb
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int mt = warp + r * nwarp;
        if (mt < MT) {
          const int pos = (8 * mt + g) * 8 + 2 * t;
#pragma unroll
          for (int i = 0; i < NW; ++i) { tb[i * Hp * 8 + pos] = c[r][i][0]; tb[i * Hp * 8 + pos + 1] = c[r][i][1]; }
        }
      }
      __syncthreads();
    }
    {  // Jacobian rows: ydot_i = W_Lh^T (d h_Lh / d v_i)
      const int in = M.size[Lh];
      const double* w = theta + M.woff[Lh];
      const int KS = mlp_round_up(in, 4) >> 2, MT = mlp_round_up(n, 8) >> 3;
      for (int mt = warp; mt < MT; mt += nwarp) {
        const int o = 8 * mt + g;
        double c[NW][2];
#pragma unroll
        for (int i = 0; i < NW; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
        _Pragma("unroll 4") for (int ks = 0; ks < KS; ++ks) {
          const int k = 4 * ks + t;
          const double a = (k < in && o < n) ? (*(w + k * n + o)) : 0.0;
#pragma unroll
          for (int i = 0; i < NW; ++i) dmma_m8n8k4(c[i][0], c[i][1], a, tb[i * Hp * 8 + k * 8 + g]);
        }
        if (o < n) {
          const int qa = q0 + 2 * t;
#pragma unroll
          for (int i = 0; i < NW; ++i) {
            if (qa < Q) dynJ[(qa * n + o) * NW + i] = c[i][0];
            if (qa + 1 < Q) dynJ[((qa + 1) * n + o) * NW + i] = c[i][1];
          }
        }
      }
    }
    if (MODE == 2) {
      // sum over the rows held by the 8 lane groups (xor 4, 8, 16), then over warps through shared memory
#pragma unroll
      for (int pp = 0; pp < NWP; ++pp)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          double v = hacc[pp][cc];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) atomicAdd(hbuf + (2 * t + cc) * NWP + pp, v);
        }
      __syncthreads();
      for (int e = tid; e < 8 * NWP; e += blockDim.x) {
        const int col = e / NWP, pp = e - col * NWP;
        if (q0 + col < Q) dynH[(q0 + col) * NWP + pp] = hbuf[e];
      }
    }
    __syncthreads();
  }
#endif
}

}  // namespace myr
