/* myriad_b200 -- C ABI of the B200-native batched trajectory-optimization engine.
 *
 * The reference (nikihowe/myriad) has no native FFI: its hot path is Python calling XLA and
 * cyipopt.  These entry points are what a binding for that path binds
 * instead; each comment names the reference interface the call replaces.  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every array is contiguous fp64 (or int32) and OWNED BY THE CALLER;
 *    device entry points take DEVICE pointers (e.g. torch.Tensor.data_ptr()) and a cudaStream_t passed
 *    as void*; the library never allocates or frees device memory (the workspace is passed in).
 *  - all batched arrays are instance-major: [B][...].
 *  - decision vector layout == the reference's ravel_pytree((x, u)): all states time-major, then all
 *    controls time-major (shooting.py:75, trapezoidal.py:51, hermite_simpson.py:47); constraint vector
 *    layout and SIGN conventions == the reference's constraints() of the same transcription.
 *  - every function returns 0 on success or a negative MYR_E_* code; myr_last_error() gives the
 *    thread-local message.  Nothing throws across the ABI.  Work is enqueued asynchronously on
 *    the given stream; the caller synchronises.
 *  - myr_host_* are CPU builds of the same templates taking HOST pointers (OpenMP over instances).  They
 *    exist for debugging / CI without a GPU and as bench.py's same-algorithm CPU baseline; the Python
 *    product path never calls them.
 */
#ifndef MYRIAD_B200_H
#define MYRIAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MYR_ABI_VERSION 4
#define MYR_MAX_PARAMS 16
#define MYR_MAX_NODE_LAYERS 5   /* Linear layers of a NODE MLP: up to 4 hidden + the output layer */
#define MYR_MAX_NODE_WIDTH 128  /* widest hidden layer */
/* Workspace of myr_kkt_solve / myr_ipm_solve: a header followed by SLOTS.  A slot belongs to a resident CTA (host
 * twin: to a thread), not to an instance, so the workspace does not grow with the batch:
 *     ws_doubles >= MYR_WS_HEADER + min(B, MyrSizes.ipm_workspace_slots) * MyrSizes.ipm_workspace_doubles */
#define MYR_WS_HEADER 16
#define MYR_WS_MAX_SLOTS 2048

/* SystemType members with a device implementation (myriad/systems/__init__.py:29-50). */
enum {
  MYR_SYS_SIMPLECASE = 0,
  MYR_SYS_CARTPOLE = 1,
  MYR_SYS_VANDERPOL = 2,
  MYR_SYS_CANCERTREATMENT = 3,
  MYR_SYS_MOULDFUNGICIDE = 4,
  MYR_SYS_BIOREACTOR = 5,
  MYR_SYS_SIMPLECASEWITHBOUNDS = 6,
  MYR_SYS_GLUCOSE = 7,
  MYR_SYS_HARVEST = 8,
  MYR_SYS_TIMBERHARVEST = 9,
  MYR_SYS_SEIR = 10,
  MYR_SYS_EPIDEMICSEIRN = 11,
  MYR_SYS_HIVTREATMENT = 12,
  MYR_SYS_BACTERIA = 13,
  MYR_SYS_TUMOUR = 14,
  MYR_SYS_PREDATORPREY = 15,
  MYR_SYS_BEARPOPULATIONS = 16,
  MYR_SYS_ROCKETLANDING = 17,
  MYR_SYS_PENDULUM = 18,     /* non-smooth: clip / angle_normalize with JAX's sub-gradient choices */
  MYR_SYS_MOUNTAINCAR = 19,  /* non-smooth: clipped force */
  MYR_SYS_INVASIVEPLANT = 20, /* DISCRETE dynamics: myr_fbsm_solve only (the reference's direct optimizers reject it too) */
  /* NodeSystem (myriad/systems/neural_ode/node_system.py:14-42) wrapping true system k: id = MYR_SYS_NODE_BASE + k.
   * Dynamics = the NODE MLP of myriad/neural_ode/create_node.py:110-117 (weights in MyrDesc.theta); cost, bounds,
   * horizon and the verification rollout are the true system's. */
  MYR_SYS_NODE_BASE = 100,
  /* systems registered at run time with myr_register_system (the reference's "add a SystemType member" plugin point,
   * myriad/systems/__init__.py:29-50): ids from MYR_SYS_USER_BASE up */
  MYR_SYS_USER_BASE = 1000
};
#define MYR_MAX_USER_SYSTEMS 64
/* OptimizerType x QuadratureRule (myriad/config.py:12-17,53-56; get_optimizer, trajectory_optimizers/__init__.py:12-28) */
enum { MYR_OPT_SHOOTING = 0, MYR_OPT_TRAPEZOIDAL = 1, MYR_OPT_HERMITE_SIMPSON = 2 };
/* IntegrationMethod (myriad/config.py:46-50) */
enum { MYR_INT_EULER = 0, MYR_INT_HEUN = 1, MYR_INT_MIDPOINT = 2, MYR_INT_RK4 = 3 };

enum {
  MYR_OK = 0,
  MYR_E_BADARG = -1,      /* unknown enum / inconsistent sizes: the reference raises KeyError / ValueError */
  MYR_E_UNSUPPORTED = -2, /* combination without a kernel yet */
  MYR_E_CUDA = -3,        /* a CUDA call failed */
  MYR_E_WORKSPACE = -4    /* workspace too small */
};

/* Per-instance solver exit codes written to status_out (IPOPT-style; the reference only reads
 * success == (status == 0), myriad/nlp_solvers/__init__.py:64). */
enum {
  MYR_ST_SOLVED = 0,
  MYR_ST_ACCEPTABLE = 1,
  MYR_ST_MAXITER = -1,
  MYR_ST_LINESEARCH = -2,
  MYR_ST_INERTIA = -3,
  MYR_ST_NAN = -13
};

/* What HParams + the system object determine (myriad/config.py:61-112, systems/base.py:11-36). */
typedef struct MyrDesc {
  int32_t system_id;
  int32_t optimizer;
  int32_t integration_method;
  int32_t intervals;
  int32_t controls_per_interval;
  int32_t n_params;            /* 0 => the system's constructor defaults */
  int32_t terminal_cost;
  int32_t reserved;
  double T;                    /* horizon; <= 0 => system default */
  double params[MYR_MAX_PARAMS];
  /* NODE systems only (system_id >= MYR_SYS_NODE_BASE), else zero.  The MLP is  Linear(h_1) sigmoid ... Linear(h_k)
   * sigmoid Linear(n)  applied to concat(x, u) (create_node.py:110-117; hp.hidden_layers = node_hidden[0..k-1]).
   * theta: DEVICE pointer for the device entry points (HOST pointer for myr_host_*), caller-owned, layer after layer
   * in haiku's order (linear, linear_1, ...: create_node.py:124-131): w as (in, out) row-major -- hk.Linear computes
   * x @ w + b -- followed by b (out). */
  int32_t node_num_hidden;     /* k, 1..MYR_MAX_NODE_LAYERS-1 */
  int32_t node_hidden[MYR_MAX_NODE_LAYERS - 1];
  const double* theta;
  int64_t theta_doubles;       /* length of theta, checked against the layer sizes */
} MyrDesc;

typedef struct MyrSizes {
  int32_t n, m;                /* state / control dimension */
  int32_t nx_nodes, nu_nodes;  /* rows of results['x'] / results['u'] */
  int32_t nvars, ncon;
  int32_t nodes, stages;       /* node/stage structure used by the block kernels */
  int32_t nw, nc;              /* variables per node block, constraint rows per stage */
  int32_t stage_nodes;         /* node blocks per stage row in Jblk */
  int32_t ipm_workspace_slots; /* most workspace slots a call ever uses (MYR_WS_MAX_SLOTS) */
  int64_t jac_block_doubles;   /* per instance: stages * stage_nodes * nc * nw */
  int64_t hess_block_doubles;  /* per instance: nodes * nw (nw + 1) / 2 (packed upper, row-major) */
  int64_t ipm_workspace_doubles; /* per workspace slot (upper bound: every array in global memory) */
} MyrSizes;

/* Options of the interior-point solve; zero / negative fields take the defaults noted (IPOPT's). */
typedef struct MyrIpmOpts {
  int32_t max_iter;        /* hp.max_iter (myriad/config.py:70); default 1000 */
  int32_t max_ls;          /* backtracking steps; default 40 */
  int32_t acceptable_iter; /* default 15 */
  int32_t max_soc;         /* second-order-correction attempts per iteration; 0 => default 4, negative => none */
  double tol;              /* default 1e-8 */
  double acceptable_tol;   /* default 1e-6 */
  double mu_init;          /* default 0.1 */
} MyrIpmOpts;

int myr_abi_version(void);
const char* myr_last_error(void);

/* Sizes implied by a descriptor.  Replaces what the optimizers' __init__ derive
 * (shooting.py:26-31, trapezoidal.py:25-29, hermite_simpson.py:28-30). */
int myr_problem_sizes(const MyrDesc* desc, MyrSizes* out);

/* K1.  Fused evaluation of objective, objective gradient, defect constraints and their block Jacobian
 * (and, if lam != NULL and Hblk != NULL, the block Hessian of f + lam.c) for B instances in one launch.
 * Replaces the four jitted callbacks of myriad/nlp_solvers/__init__.py:31-42 (fun, jac = jax.grad,
 * constraints fun, constraints jac = jax.jacrev); the dense ncon x nvars Jacobian is returned in compact
 * block form: Jblk[b][stage][k][r][i] = d c_{stage,r} / d v_{node(stage,k), i}.
 * Any of f, grad, c, Jblk, Hblk may be NULL. */
int myr_eval(const MyrDesc* desc, int B, const double* z, const double* lam,
             double* f, double* grad, double* c, double* Jblk, double* Hblk, void* stream);

/* K2.  Batched block-structured KKT solve
 *     [ H + diag(sigma) + delta_w I    J^T      ] [ dz   ]     [ rhs_z ]
 *     [ J                            -delta_c I ] [ dlam ] = - [ rhs_c ]
 * by node-block elimination, Schur complement and block cyclic reduction.  sigma[i] = +inf marks an
 * eliminated (fixed) variable.  inertia_ok[b] = 1 iff the matrix has exactly ncon negative and no zero
 * eigenvalues.  Replaces the linear solver inside IPOPT (MUMPS) for this problem class.
 * ws: device workspace, see MYR_WS_HEADER. */
int myr_kkt_solve(const MyrDesc* desc, int B, const double* Hblk, const double* Jblk, const double* sigma,
                  const double* rhs_z, const double* rhs_c, double delta_w, double delta_c,
                  double* dz, double* dlam, int32_t* inertia_ok, double* ws, size_t ws_doubles, void* stream);

/* K3.  Whole batched primal-dual interior-point solve: replaces cyipopt.minimize_ipopt at
 * myriad/nlp_solvers/__init__.py:56-58.  z0/lb/ub per instance as the optimizers build them
 * (guess, bounds; lb == ub fixes a variable).  Outputs: z (solution['x']), lam (solution.info['mult_g']),
 * obj (solution['fun']), status/iters, zL/zU bound multipliers, kkt_err (scaled optimality error at
 * exit), con_inf (max |c|). */
int myr_ipm_solve(const MyrDesc* desc, const MyrIpmOpts* opts, int B,
                  const double* z0, const double* lb, const double* ub,
                  double* z, double* lam, double* zL, double* zU,
                  double* obj, double* kkt_err, double* con_inf, int32_t* status, int32_t* iters,
                  double* ws, size_t ws_doubles, void* stream);

/* Post-solve verification rollout of the TRUE system under controls u with hp.integration_method:
 * replaces get_state_trajectory_and_cost (myriad/utils.py:258-298).  u: [B][nu_rows][m];
 * x0: [B][n]; xs: [B][num_steps+1][n] (may be NULL); cost: [B]. */
int myr_rollout_cost(const MyrDesc* desc, int B, int nu_rows, const double* u, const double* x0,
                     double* xs, double* cost, void* stream);

/* Point evaluation of a system for B (x, u, t) points in one launch: f = dynamics(x, u, t) and g = cost(x, u, t).
 * Replaces FiniteHorizonControlSystem.dynamics / .cost (myriad/systems/base.py:45-73) when host code calls them
 * (the solver itself evaluates them inside K1).  x: [B][n]; u: [B][m]; t: [B] or NULL (0); f: [B][n] or NULL; g: [B] or NULL.
 * NODE systems evaluate the MLP dynamics (node_system.py:36-38) and the true system's cost. */
int myr_dynamics(const MyrDesc* desc, int B, const double* x, const double* u, const double* t,
                 double* f, double* g, void* stream);

/* out[b] = J(z_b)^T lam[b] ([B][nvars], reference layout) from the compact block Jacobian Jblk that myr_eval returned:
 * the vector-Jacobian product the reference takes with jax.grad of the Lagrangian in its extragradient solver
 * (myriad/nlp_solvers/extra_gradient.py:21-33).  Works for all three transcriptions. */
int myr_jtvec(const MyrDesc* desc, int B, const double* Jblk, const double* lam, double* out, void* stream);

/* Options of the forward-backward sweep; zero / negative fields take the reference's values. */
typedef struct MyrFbsmOpts {
  int32_t max_iter;    /* sweeps per fixed-point solve; the reference loops without a cap; default 10000 */
  int32_t max_secant;  /* secant steps on the free terminal adjoint; default 100 */
  int32_t term_state;  /* index of the ONE state with a terminal value (system.x_T[i] is not None), or -1: none
                        * (forward_backward_sweep.py:59-70) */
  int32_t reserved;
  double delta;        /* stopping_criterion's delta (trajectory_optimizers/base.py:129); default 1e-3 */
  double secant_tol;   /* |x_T - target| at which the secant iteration stops (forward_backward_sweep.py:137); default 1e-10 */
  double term_value;   /* the terminal value of state term_state */
  double guess_a, guess_b; /* system.guess_a / guess_b: the two starting values of the secant iteration */
} MyrFbsmOpts;

/* Indirect optimizer: the Forward-Backward Sweep Method for B start states in one launch, one thread per instance.
 * Replaces FBSM.solve / sequencesolver (myriad/trajectory_optimizers/forward_backward_sweep.py:88-158) with its RK4 sweeps
 * (integrate_fbsm, myriad/utils.py:138-197), the systems' adj_ODE / optim_characterization (myriad/systems/lenhart) and
 * the stopping rule (trajectory_optimizers/base.py:128-141).  desc: system_id, intervals = hp.fbsm_intervals, T, params
 * (the other fields are ignored).  x0: [B][n] DEVICE.  adj_T: [n] HOST, system.adj_T, or NULL (zeros).  char_lb / char_ub:
 * [m] HOST, the bounds row(s) the system's optim_characterization clamps with.
 * Outputs (DEVICE) are TIME-MAJOR with the instance index fastest -- x: [N+1][n][B], u: [N+1][m][B], adj: [N+1][n][B] --
 * unlike the instance-major convention of the other calls: they are the sweep's working storage and this is the layout
 * in which a warp's accesses coalesce.  iters: [B] sweeps performed; status: [B] MYR_ST_SOLVED / MYR_ST_MAXITER / MYR_ST_NAN.
 * Systems: the 14 indirect (Lenhart) members of SystemType; others return MYR_E_UNSUPPORTED.  For the discrete one
 * (INVASIVEPLANT) intervals must equal T (h = 1) and u has N rows: u: [N][m][B]. */
int myr_fbsm_solve(const MyrDesc* desc, const MyrFbsmOpts* opts, int B, const double* x0, const double* adj_T,
                   const double* char_lb, const double* char_ub, double* x, double* u, double* adj,
                   int32_t* iters, int32_t* status, void* stream);

/* Registers a system that was compiled into its own shared library (myriad_b200/plugin.py builds it from symbolic
 * dynamics / cost with the same generator and kernel templates as the built-in systems).  vtable: the pointer returned by
 * the plugin library's myr_vtable_<Sys>() export; its id must be >= MYR_SYS_USER_BASE.  Replaces subclassing
 * FiniteHorizonControlSystem + adding a SystemType member (myriad/systems/base.py:11-111, systems/__init__.py:29-50). */
int myr_register_system(const void* vtable);

/* Measurement helper (no reference counterpart): launches `blocks` CTAs x 1024 threads, each doing `iters` rounds of 8
 * independent fp64 FMAs (2 * 8 * iters * 1024 * blocks flops); out: [blocks * 1024] doubles.  bench.py times it with
 * CUDA events to get the device's fp64 FMA peak, the denominator for the KKT / interior-point kernel's FLOP/s. */
int myr_bench_dfma(int blocks, int iters, double* out, void* stream);

/* Host twins (debug / CI and the "same algorithm on the CPU" baseline of bench.py; HOST pointers; OpenMP over
 * instances, one workspace slot per thread). */
int myr_host_eval(const MyrDesc* desc, int B, const double* z, const double* lam,
                  double* f, double* grad, double* c, double* Jblk, double* Hblk);
int myr_host_kkt_solve(const MyrDesc* desc, int B, const double* Hblk, const double* Jblk, const double* sigma,
                       const double* rhs_z, const double* rhs_c, double delta_w, double delta_c,
                       double* dz, double* dlam, int32_t* inertia_ok, double* ws, size_t ws_doubles);
int myr_host_ipm_solve(const MyrDesc* desc, const MyrIpmOpts* opts, int B,
                       const double* z0, const double* lb, const double* ub,
                       double* z, double* lam, double* zL, double* zU,
                       double* obj, double* kkt_err, double* con_inf, int32_t* status, int32_t* iters,
                       double* ws, size_t ws_doubles);
int myr_host_rollout_cost(const MyrDesc* desc, int B, int nu_rows, const double* u, const double* x0,
                          double* xs, double* cost);
int myr_host_dynamics(const MyrDesc* desc, int B, const double* x, const double* u, const double* t,
                      double* f, double* g);
int myr_host_jtvec(const MyrDesc* desc, int B, const double* Jblk, const double* lam, double* out);
int myr_host_fbsm_solve(const MyrDesc* desc, const MyrFbsmOpts* opts, int B, const double* x0, const double* adj_T,
                        const double* char_lb, const double* char_ub, double* x, double* u, double* adj,
                        int32_t* iters, int32_t* status);

#ifdef __cplusplus
}
#endif
#endif /* MYRIAD_B200_H */
