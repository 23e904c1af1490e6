// Common device helpers: problem descriptor, block reductions, small symmetric-indefinite
// factorisation with inertia (Bunch-Parlett complete pivoting) used by the KKT kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#pragma nv_diag_suppress 177   // generated system code declares symbols it may not use
#pragma nv_diag_suppress 550

#include "systems_gen.cuh"
#ifdef MYR_EXTRA_SYSTEM_HEADER   // a user-defined system (myriad_b200/plugin.py): generated the same way, compiled on its own
#include MYR_EXTRA_SYSTEM_HEADER
#endif

namespace myr {

#ifdef MYR_PROFILE_PHASES
__device__ unsigned long long g_phase_cycles[16];  // per-phase cycle / event counters (debug builds)
#endif

constexpr int kMaxParams = 16;
constexpr int kMaxMlpLayers = 5;    // == MYR_MAX_NODE_LAYERS: Linear layers of a NODE MLP (hidden + output)
constexpr int kMaxMlpWidth = 128;   // == MYR_MAX_NODE_WIDTH

// NODE dynamics (node_mlp.cuh): layer sizes and offsets into the caller's weight vector theta
struct MlpDesc {
  int L;                          // number of Linear layers; 0 = analytic dynamics
  int hp;                         // widest hidden layer rounded up to a multiple of 8
  int size[kMaxMlpLayers + 1];    // size[0] = n + m, size[1..L-1] hidden widths, size[L] = n
  int woff[kMaxMlpLayers], boff[kMaxMlpLayers];
  const double* theta;
};

enum Method : int { EULER = 0, HEUN = 1, MIDPOINT = 2, RK4 = 3 };
enum Optimizer : int { OPT_SHOOTING = 0, OPT_TRAPEZOIDAL = 1, OPT_HERMITE_SIMPSON = 2 };

// Kernel-side problem description (by value in kernel params -> constant bank).
struct Problem {
  int B;          // instances
  int N;          // collocation intervals, or shooting intervals K
  int cpi;        // controls per interval (shooting)
  int method;     // Method
  int nvars;      // reference decision-vector length
  int ncon;       // reference constraint-vector length
  int terminal_cost;
  double T;
  double h;       // T / N (collocation) or T / (N * cpi) (shooting step)
  double p[kMaxParams];
  MlpDesc mlp;
};

__host__ __device__ __forceinline__ constexpr int packed_size(int n) { return n * (n + 1) / 2; }
// packed upper triangle, row-major
__host__ __device__ __forceinline__ int pidx(int i, int j, int n) {
  if (i > j) { int t = i; i = j; j = t; }
  return i * n - (i * (i - 1)) / 2 + (j - i);
}

// ------------------------------------------------------------------ execution model
// Every per-instance routine is written once and compiled twice: for the device (one CTA per
// instance, MYR_NT threads striding over nodes/stages, __syncthreads between phases) and for the
// host (one "thread", barriers vanish).  The host build is a debugging twin used by CPU tests; the
// product path never dispatches to it.
#ifdef __CUDA_ARCH__
#define MYR_TID (int(threadIdx.x))
#define MYR_NT (int(blockDim.x))
#define MYR_SYNC() __syncthreads()
#else
#define MYR_TID 0
#define MYR_NT 1
#define MYR_SYNC() ((void)0)
#endif
#define MYR_HDI __host__ __device__ __forceinline__
#define MYR_HDN __host__ __device__ __noinline__

// ------------------------------------------------------------------ block reductions
// red: shared scratch of >= 33 doubles.  All threads of the CTA must call.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <int OP>  // 0 sum, 1 max, 2 min
__host__ __device__ __forceinline__ double block_reduce(double v, double* red) {
#ifndef __CUDA_ARCH__
  (void)red;
  return v;
#else
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = OP == 0 ? warp_sum(v) : (OP == 1 ? warp_max(v) : warp_min(v));
  __syncthreads();  // protect red from a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = OP == 0 ? 0.0 : (OP == 1 ? -INFINITY : INFINITY);
  if (lane < nw) r = red[lane];
  r = OP == 0 ? warp_sum(r) : (OP == 1 ? warp_max(r) : warp_min(r));
  return r;  // every thread has the result
#endif
}
MYR_HDI double block_sum(double v, double* red) { return block_reduce<0>(v, red); }
MYR_HDI double block_max(double v, double* red) { return block_reduce<1>(v, red); }
MYR_HDI double block_min(double v, double* red) { return block_reduce<2>(v, red); }

// NaN-propagating max for error norms: fmax() drops NaNs, which would hide a blown-up lane.
MYR_HDI double nanmax(double a, double b) { return (a != a || b != b) ? NAN : fmax(a, b); }

// ------------------------------------------------------------------ small symmetric indefinite inverse + inertia
// A: full symmetric N x N (row-major, both triangles valid) is destroyed.  inv: full N x N out.
// mask bit i set => variable i is eliminated (fixed): it is skipped, its row/col of inv are zero and it
// does not count in the inertia.  Bunch-Parlett: complete pivoting with 1x1 / 2x2 pivots.
template <int N>
__host__ __device__ __noinline__ void sym_inverse_inertia(double* A, uint32_t mask, double* inv, int& npos, int& nneg, int& nzero) {
  int perm[N];   // perm[k] = original index placed at position k (active ones first)
  int na = 0;
#pragma unroll
  for (int i = 0; i < N; ++i) if (!((mask >> i) & 1u)) perm[na++] = i;
  for (int i = 0, k = na; i < N; ++i) if ((mask >> i) & 1u) perm[k++] = i;
  // work on permuted copy M (na x na) held in local array
  double M[N * N];
  double scale = 0.0;
  for (int i = 0; i < na; ++i)
    for (int j = 0; j < na; ++j) { M[i * N + j] = A[perm[i] * N + perm[j]]; scale = fmax(scale, fabs(M[i * N + j])); }
  int p2[N];       // second-level permutation from pivoting (positions within active set)
  for (int i = 0; i < na; ++i) p2[i] = i;
  int piv2[N];     // piv2[k] = 1 if a 2x2 pivot starts at k
  double L[N * N]; // unit lower factors (column k / k,k+1)
  for (int i = 0; i < N * N; ++i) L[i] = 0.0;
  for (int i = 0; i < N; ++i) piv2[i] = 0;
  const double alpha = 0.6403882032022076;
  const double tiny = 1e-13 * fmax(scale, 1e-300);
  npos = nneg = nzero = 0;
  int k = 0;
  while (k < na) {
    // largest diagonal and largest off-diagonal in trailing block
    double mu0 = -1.0, mu1 = -1.0; int r = k, pp = k, qq = k;
    for (int i = k; i < na; ++i) {
      double d = fabs(M[i * N + i]);
      if (d > mu0) { mu0 = d; r = i; }
      for (int j = k; j < i; ++j) { double o = fabs(M[i * N + j]); if (o > mu1) { mu1 = o; pp = i; qq = j; } }
    }
    auto swap_sym = [&](int a, int b) {
      if (a == b) return;
      for (int j = 0; j < na; ++j) { double t = M[a * N + j]; M[a * N + j] = M[b * N + j]; M[b * N + j] = t; }
      for (int i = 0; i < na; ++i) { double t = M[i * N + a]; M[i * N + a] = M[i * N + b]; M[i * N + b] = t; }
      for (int j = 0; j < k; ++j) { double t = L[a * N + j]; L[a * N + j] = L[b * N + j]; L[b * N + j] = t; }
      int t = p2[a]; p2[a] = p2[b]; p2[b] = t;
    };
    if (mu0 <= tiny && mu1 <= tiny) {  // remaining block is numerically zero
      nzero += na - k;
      for (int i = k; i < na; ++i) { M[i * N + i] = 1.0; for (int j = k; j < na; ++j) if (j != i) M[i * N + j] = 0.0; }
      break;
    }
    if (mu0 >= alpha * mu1) {
      swap_sym(k, r);
      const double d = M[k * N + k];
      if (d > 0) ++npos; else ++nneg;
      for (int i = k + 1; i < na; ++i) L[i * N + k] = M[i * N + k] / d;
      for (int i = k + 1; i < na; ++i)
        for (int j = k + 1; j <= i; ++j) {
          M[i * N + j] -= L[i * N + k] * d * L[j * N + k];
          M[j * N + i] = M[i * N + j];
        }
      k += 1;
    } else {
      // 2x2 pivot on (qq, pp), qq < pp
      swap_sym(k, qq);
      swap_sym(k + 1, pp);
      const double a = M[k * N + k], b = M[(k + 1) * N + k], c = M[(k + 1) * N + k + 1];
      const double det = a * c - b * b;  // < 0 by the pivot test
      ++npos; ++nneg;
      piv2[k] = 1;
      for (int i = k + 2; i < na; ++i) {
        const double s0 = M[i * N + k], s1 = M[i * N + k + 1];
        L[i * N + k] = (s0 * c - s1 * b) / det;
        L[i * N + k + 1] = (-s0 * b + s1 * a) / det;
      }
      for (int i = k + 2; i < na; ++i)
        for (int j = k + 2; j <= i; ++j) {
          const double s0 = M[j * N + k], s1 = M[j * N + k + 1];
          M[i * N + j] -= L[i * N + k] * s0 + L[i * N + k + 1] * s1;
          M[j * N + i] = M[i * N + j];
        }
      k += 2;
    }
  }
  // M now holds D on its (block) diagonal.  Build inverse column by column.
  for (int i = 0; i < N * N; ++i) inv[i] = 0.0;
  for (int col = 0; col < na; ++col) {
    double y[N];
    for (int i = 0; i < na; ++i) y[i] = (p2[i] == col) ? 1.0 : 0.0;   // P^T e_col
    for (int i = 0; i < na; ++i) {                                  // L y = b
      double s = y[i];
      for (int j = 0; j < i; ++j) s -= L[i * N + j] * y[j];
      y[i] = s;
    }
    for (int i = 0; i < na; ++i) {                                  // D^-1
      if (piv2[i]) {
        const double a = M[i * N + i], b = M[(i + 1) * N + i], c = M[(i + 1) * N + i + 1];
        const double det = a * c - b * b;
        const double y0 = y[i], y1 = y[i + 1];
        y[i] = (c * y0 - b * y1) / det;
        y[i + 1] = (-b * y0 + a * y1) / det;
        ++i;
      } else {
        y[i] /= M[i * N + i];
      }
    }
    for (int i = na - 1; i >= 0; --i) {                             // L^T x = y
      double s = y[i];
      for (int j = i + 1; j < na; ++j) s -= L[j * N + i] * y[j];
      y[i] = s;
    }
    for (int i = 0; i < na; ++i) inv[perm[p2[i]] * N + perm[col]] = y[i];
  }
}


// Fast path: LDL^T in natural order (no pivoting), everything in registers (static indexing, fully unrolled).
// Valid whenever no pivot is small relative to the block's scale; otherwise returns false and the caller falls
// back to the pivoted Bunch-Parlett routine above.  Masked (fixed) variables are decoupled before factorising
// and their rows/cols of the inverse are zero.  A: full symmetric N x N (not modified).
template <int N>
MYR_HDI bool sym_inverse_inertia_fast(const double* A, uint32_t mask, double* inv, int& npos, int& nneg, double& piv_ratio) {
  double a[N][N];
  double scale = 0.0;
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const bool mi = (mask >> i) & 1u, mj = (mask >> j) & 1u;
      const double v = (mi || mj) ? ((i == j) ? 1.0 : 0.0) : A[i * N + j];
      a[i][j] = v;
      scale = fmax(scale, fabs(v));
    }
  // A 1x1 pivot is accepted when it bounds the multipliers of its own column (|d_k| > 1e-7 max_i |v_ik|, the
  // Bunch-Kaufman test with a loose alpha) and is not numerically zero against the block; a small diagonal entry whose
  // column is (nearly) decoupled -- e.g. a state that enters neither cost nor dynamics and sits far from its bounds, so
  // that only a tiny barrier term is on its diagonal -- is a perfectly good pivot and must not force the slow path.
  const double tiny = 1e-14 * scale;
  double d[N], l[N][N];
  bool ok = true;
  int np_ = 0, nn_ = 0;
  double minpiv = INFINITY;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double dk = a[k][k];
#pragma unroll
    for (int j = 0; j < k; ++j) dk -= l[k][j] * l[k][j] * d[j];
    d[k] = dk;
    minpiv = fmin(minpiv, fabs(dk));
    const bool masked = (mask >> k) & 1u;
    if (!masked) { if (dk > 0) ++np_; else ++nn_; }
    const double rk = 1.0 / dk;
    double colmax = 0.0;
#pragma unroll
    for (int i = k + 1; i < N; ++i) {
      double v = a[i][k];
#pragma unroll
      for (int j = 0; j < k; ++j) v -= l[i][j] * l[k][j] * d[j];
      colmax = fmax(colmax, fabs(v));
      l[i][k] = v * rk;
    }
    ok = ok && (fabs(dk) > 1e-7 * colmax) && (fabs(dk) > tiny);
  }
  if (!ok) return false;
  // Linv (unit lower): m = L^-1
  double m[N][N];
#pragma unroll
  for (int c = 0; c < N; ++c) {
#pragma unroll
    for (int i = c + 1; i < N; ++i) {
      double v = -l[i][c];
#pragma unroll
      for (int j = c + 1; j < i; ++j) v -= l[i][j] * m[j][c];
      m[i][c] = v;
    }
  }
  // inv = m^T D^-1 m
  double rd[N];
#pragma unroll
  for (int k = 0; k < N; ++k) rd[k] = 1.0 / d[k];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = i; j < N; ++j) {
      // sum over k >= max(i,j)=j: m[k][i] * rd[k] * m[k][j], with m[k][k] = 1
      double v = 0.0;
#pragma unroll
      for (int k = j; k < N; ++k) {
        const double mki = (k == i) ? 1.0 : m[k][i];
        const double mkj = (k == j) ? 1.0 : m[k][j];
        v += mki * rd[k] * mkj;
      }
      const bool mi = (mask >> i) & 1u, mj = (mask >> j) & 1u;
      v = (mi || mj) ? 0.0 : v;
      inv[i * N + j] = v;
      inv[j * N + i] = v;
    }
  npos = np_; nneg = nn_;
  piv_ratio = minpiv / fmax(scale, 1e-300);
  return true;
}

// inverse + inertia: fast path first, pivoted fallback
// The pivoted routine is out of line and takes pointers: it works on ITS OWN copies so that the caller's A / inv never
// have their address escape and stay in registers on the (overwhelmingly common) fast path.
template <int N>
MYR_HDI void sym_inverse(double* A, uint32_t mask, double* inv, int& npos, int& nneg, int& nzero, double* piv_ratio = nullptr) {
  nzero = 0;
  double pr = 0.0;
  if (sym_inverse_inertia_fast<N>(A, mask, inv, npos, nneg, pr)) {
    if (piv_ratio) *piv_ratio = pr;
    return;
  }
  if (piv_ratio) *piv_ratio = 0.0;  // pivoted fallback: treat as ill-conditioned
  double A2[N * N], inv2[N * N];
  int p2_, n2_, z2_;
#pragma unroll
  for (int i = 0; i < N * N; ++i) A2[i] = A[i];
  sym_inverse_inertia<N>(A2, mask, inv2, p2_, n2_, z2_);
#pragma unroll
  for (int i = 0; i < N * N; ++i) inv[i] = inv2[i];
  npos = p2_; nneg = n2_; nzero = z2_;
}

}  // namespace myr
