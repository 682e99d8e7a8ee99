/*
 * topo_b200.h -- C ABI of libtopo_b200.so: the B200 (sm_100a) implementation of the per-iteration
 * topology-optimisation hot path of SolidsPy-Opt (SIMP / BESO / ESO on structured quad4 meshes).
 *
 * The reference has no FFI layer: it is pure Python (numpy/scipy) and its boundary is the four
 * functions of src/solidspy_opt/optimize.py (SIMP :360, BESO :215, ESO_stiff :116, ESO_stress :16).
 * Each entry point below replaces the reference statement(s) cited next to it; the Python package
 * solidspy-opt_b200/solidspy_opt re-creates the four functions on top of this ABI through ctypes
 * (see INTEGRATION.md for the binding a maintainer of the reference would add).
 *
 * Conventions
 *   - every function returns 0 on success, a negative TOPO_E* code on failure; topo_last_error()
 *     returns a human-readable message for the calling thread's last failure;
 *   - all host arrays are caller-owned, C-contiguous; doubles are IEEE fp64, ids are int64/int32
 *     as declared; nothing is retained after the call returns;
 *   - a handle owns all device memory for one mesh on one GPU; handles are not re-entrant;
 *   - mesh numbering is the reference's (solidspy rect_grid): node id = row*(nx+1)+col, x fastest,
 *     y upwards; element id = row*nx+col; connectivity [n, n+1, n+nx+2, n+nx+1], n = e + row;
 *     element DOF order [ux0,uy0,ux1,uy1,ux2,uy2,ux3,uy3];
 *   - nodal vectors live in "full DOF space": 2*(nx+1)*(ny+1) doubles, (ux,uy) interleaved per node
 *     -- exactly the (N_n,2) array solidspy.postprocesor.complete_disp returns (optimize.py:438);
 *     constrained DOFs hold 0.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     TOPO_ENODEV.
 */
#ifndef TOPO_B200_H
#define TOPO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOPO_ABI_VERSION 1

enum {
  TOPO_OK = 0,
  TOPO_EINVAL = -1,    /* bad argument */
  TOPO_ENODEV = -2,    /* no usable CUDA device */
  TOPO_ECUDA = -3,     /* CUDA runtime error (message in topo_last_error) */
  TOPO_ENOCONV = -4,   /* linear solve hit maxit before reaching rtol (results still readable) */
  TOPO_ESTATE = -5     /* call order violated (e.g. solve before loads were set) */
};

/* preconditioners for topo_solve */
enum { TOPO_PRECOND_JACOBI = 0, TOPO_PRECOND_GMG = 1 };

/* device fields readable/writable through topo_get / topo_set */
enum {
  TOPO_F_U = 0,        /* displacement, 2*N_n doubles                      (optimize.py:438 UC)     */
  TOPO_F_LOAD = 1,     /* load vector in full DOF space, 2*N_n doubles     (optimize.py:434)         */
  TOPO_F_RHO = 2,      /* density, N_e doubles                             (optimize.py:399,449)     */
  TOPO_F_DC = 3,       /* (filtered) sensitivity, N_e doubles              (optimize.py:444-445)     */
  TOPO_F_SCALE = 4,    /* element stiffness scale E_e, N_e doubles         (optimize.py:430 mats[:,2]) */
  TOPO_F_CONS = 5,     /* BC flags, 2*N_n int8: -1 constrained, 0 free     (nodes[:,3:5])            */
  TOPO_F_SENS = 6,     /* BESO/ESO sensitivity number, N_e doubles         (optimize.py:188,304-311) */
  TOPO_F_KEEP = 7,     /* element mask, N_e uint8: 1 = present             (optimize.py:273,328)     */
  TOPO_F_RESID = 8,    /* last residual f - K u, 2*N_n doubles                                       */
  TOPO_F_SENS_PREV = 9,/* BESO history sensitivity (optimize.py:344 sensi_I), N_e doubles            */
  TOPO_F_NODAL = 10    /* BESO nodal sensitivity (optimize.py:305 sensi_nodes), N_n doubles          */
};

typedef struct topo_handle topo_t;

/* ---- life cycle -------------------------------------------------------------------------- */
int topo_abi_version(void);
const char *topo_last_error(void);
int topo_device_count(int *count);

/* Mesh of nx*ny quad4 elements sharing one 8x8 element matrix ke (row-major).  Replaces the arrays
 * built by beam()/rect_grid (Utils/beams.py:73-90) + DME (optimize.py:417): connectivity and DOF
 * numbering are implicit in (nx, ny).  `device` is the CUDA ordinal. */
int topo_create(topo_t **out, int nx, int ny, const double ke[64], int device);
int topo_destroy(topo_t *h);

/* ---- model state --------------------------------------------------------------------------- */
/* BC flags as in nodes[:,3:5] (-1 constrained / 0 free), 2*N_n int8.  Replaces the `cons` argument of
 * ass.DME (optimize.py:56,79,156,176,255,293,417). */
int topo_set_cons(topo_t *h, const int8_t *cons);
/* Point loads with ass.loadasem semantics (optimize.py:58,81,158,178,257,295,434): value ASSIGNED to
 * free DOFs in row order (later rows win), constrained DOFs skipped.  The (node,fx,fy) rows are kept
 * and re-applied whenever the BC flags change. */
int topo_set_loads(topo_t *h, const int64_t *node, const double *fx, const double *fy, int n_loads);
int topo_set(topo_t *h, int field, const void *host, size_t bytes);
int topo_get(topo_t *h, int field, void *host, size_t bytes);
/* scalar fill of an element field (TOPO_F_RHO / TOPO_F_SCALE / TOPO_F_KEEP) */
int topo_fill(topo_t *h, int field, double value);

/* ---- stage 1: matrix-free stiffness application --------------------------------------------- */
/* y = K(scale) u restricted to free DOFs (rows/cols of constrained DOFs removed, y = 0 there):
 * the product stiff_mat.dot(disp) of optimize.py:455 without sparse_assem (solver.py:364-410). */
int topo_matvec(topo_t *h, const double *u_host, double *y_host);
/* same, device-resident operands: y(field_y) = K u(field_u); used by bench for kernel timing */
int topo_matvec_device(topo_t *h, int repeats, float *ms_per_launch);

/* ---- stage 2: linear solve (replaces spsolve, optimize.py:61,84,161,181,260,298,437) -------- */
/* Preconditioned CG on K u = f.  warm != 0 starts from the current TOPO_F_U.  Stops when
 * ||f - K u||_2 <= rtol * ||f||_2.  Returns TOPO_ENOCONV if maxit is reached. */
int topo_solve(topo_t *h, int precond, double rtol, int maxit, int warm, int *iters, double *relres);
/* compliance f.u (optimize.py:440) */
int topo_compliance(topo_t *h, double *c);
/* The reference's guard np.allclose(K u / K.max(), f / K.max()) (optimize.py:72,172,286,455). */
int topo_equilibrium_check(topo_t *h, int *ok);

/* ---- stages 3+4 for SIMP ------------------------------------------------------------------- */
/* scale_e = emin + rho_e^penal (emax - emin)                                   (optimize.py:430) */
int topo_simp_scale(topo_t *h, double penal, double emin, double emax);
/* dc_e = -penal rho_e^(penal-1) (emax-emin) u_e^T ke u_e                       (optimize.py:443-444);
 * *objective receives sum_e scale_e u_e^T ke u_e (= u^T K u, the energy form of f.u) */
int topo_simp_sensitivity(topo_t *h, double penal, double emin, double emax, double *objective);
/* dc <- sum_j H_ij rho_j dc_j / (max(1e-3,rho_i) sum_j H_ij), H = max(0, rmin - dist)
 * (density_filter, solver.py:452-477; call optimize.py:445).  dx, dy: element size. */
int topo_filter_sensitivity(topo_t *h, double rmin, double dx, double dy);
/* optimality_criteria (solver.py:412-448; call optimize.py:449) + the change norm (optimize.py:452):
 * rho <- OC(rho, dc, *g); *g <- gt; *change <- max|rho_new - rho_old|; *nsteps bisection steps. */
int topo_oc_update(topo_t *h, double move, double *g, double *change, int *nsteps);
/* One whole SIMP design iteration on device-resident state (optimize.py:430-455). */
typedef struct {
  double penal, emin, emax, rmin, dx, dy, move;
  double rtol;
  int precond, maxit, warm;
} topo_simp_params;
typedef struct {
  double compliance, objective, g, change, relres;
  int cg_iters, oc_steps, equilibrium_ok, status;
} topo_simp_result;
int topo_simp_step(topo_t *h, const topo_simp_params *p, double *g_inout, topo_simp_result *res);
/* Same iteration through HOST buffers: uploads rho_in (N_e), runs the step, downloads rho_out (N_e). */
int topo_simp_step_host(topo_t *h, const topo_simp_params *p, const double *rho_in, double *rho_out,
                        double *g_inout, topo_simp_result *res);

/* ---- stages 3+4 for BESO / ESO -------------------------------------------------------------- */
/* sens_e = (1/2 u_e^T ke u_e) * keep_e, divided by its max     (sensitivity_elsBESO solver.py:90-129,
 * sensitivity_elsESO :245-279).  use_keep = 0 ignores the mask (ESO: absent elements have scale 0
 * and are excluded from max/threshold through keep anyway). */
int topo_energy_sensitivity(topo_t *h, int use_keep);
/* BESO smoothing: element -> node (sensitivity_nodes, solver.py:177-210) -> element
 * (sensitivity_filter, :212-243; r_min selecting the 4 corner nodes), / max; then, if
 * average_with_prev, (s + s_prev)/2 (optimize.py:310); / max (optimize.py:311).
 * w4/w2x/w2y/we: the reference's floating-point weights (computed by the host wrapper with the
 * reference's formula). */
int topo_beso_filter(topo_t *h, const double w4[4], const double w2x[2], const double w2y[2],
                     const double we[4], int average_with_prev);
/* The two halves of topo_beso_filter as separate calls (the reference's helper-level API):
 * nodal <- sensitivity_nodes(sens)   (solver.py:177-210)   and
 * sens  <- sensitivity_filter(nodal) (solver.py:212-243, including its division by the max). */
int topo_beso_nodes(topo_t *h, const double w4[4], const double w2x[2], const double w2y[2]);
int topo_beso_elements(topo_t *h, const double we[4]);
/* sensitivity_filter for ANY r_min (solver.py:212-243): the nodes within r_min of an element centre, cut off at the mesh
 * boundary, weighted (1 - r/sum r)/(count - 1) and normalised, then divided by the max; dx, dy: element size. */
int topo_beso_elements_radius(topo_t *h, double r_min, double dx, double dy);
/* alpha = sort(sens)[::-1][k] over ALL N_e values (optimize.py:323-325): device radix select. */
int topo_rank_select(topo_t *h, int64_t k, double *alpha);
/* BESO: keep = sens > alpha; keep |= protect_els(removed) ; del_node  (optimize.py:328-331,
 * solver.py:6-58).  protect: node ids (loads + BC).  Outputs kept count and #re-added elements. */
int topo_beso_apply(topo_t *h, double alpha, const int64_t *protect_nodes, int n_protect,
                    int64_t *n_kept, int64_t *n_protected);
/* ESO: delete present elements with sens < rr unless protected; scale_e <- 0, keep_e <- 0;
 * del_nodeESO (optimize.py:189-196, solver.py:318-362). */
int topo_eso_apply(topo_t *h, double rr, const int64_t *protect_nodes, int n_protect, int64_t *n_kept);
/* Number of present elements (keep = 1) that no chain of node-sharing present elements connects to an element touching
 * one of `support_nodes`: a floating part.  That is what turns the reference's stiffness matrix singular once ESO has
 * cut a piece loose (spsolve at optimize.py:84,181 fills the solution with NaN when SuperLU meets an exactly zero pivot,
 * the guard :72,:172 then ends the loop); the host loop uses it for its `stop_on_island` option. */
int topo_floating_elements(topo_t *h, const int64_t *support_nodes, int n_support, int64_t *n_floating);
/* sens_prev <- sens (optimize.py:344) */
int topo_beso_save_history(topo_t *h);
/* ESO_stress criterion: nodal stress averaging over present elements + element average + von Mises,
 * divided by max (strain_nodes contract; strain_els solver.py:281-316; optimize.py:86-92).
 * E, nu: material; dx, dy: element size. */
int topo_vonmises_sensitivity(topo_t *h, double E, double nu, double dx, double dy);
/* Nodal strains and stresses of the current displacements: corner-point values of the adjacent elements averaged per
 * node -- the (E_nodes, S_nodes) pair of pos.strain_nodes that the reference computes after every solve and hands to
 * its field plots (optimize.py:63,86,102-103,163,183,201-202,262,300,347-348).  present_only != 0: only elements
 * still present contribute (ESO, whose `els` shrinks); 0: all elements (BESO keeps the full `els`).  Nodes without a
 * contributing element hold NaN (0/0), as in the reference.  Outputs: 3*N_n doubles each, component-major
 * (exx | eyy | gxy planes, sxx | syy | txy planes). */
int topo_strain_nodes(topo_t *h, double E, double nu, double dx, double dy, int present_only,
                      double *strain_host, double *stress_host);

/* ---- multi-GPU building blocks (row slabs) --------------------------------------------
 * One process per GPU owns a slab of element rows as an ordinary handle (nx x ny_local); interface node
 * rows are duplicated on both neighbours.  These entry points work on CALLER-OWNED DEVICE pointers
 * (e.g. torch tensors) so that the host side can exchange / all-reduce them with NCCL on the same stream. */
/* run all of the handle's work on this CUDA stream (cudaStream_t passed as void*); default: private stream */
int topo_set_stream(topo_t *h, void *cuda_stream);
/* y_dev = K(scale) u_dev restricted to free DOFs; on a slab the interface rows of y hold PARTIAL sums
 * (local elements only) that the caller adds to the neighbour's partials (halo sum). */
int topo_matvec_dev(topo_t *h, const double *u_dev, double *y_dev);
/* diag_dev = diag(K(scale)) (2 per node, 0 at constrained DOFs); partial on interface rows, like topo_matvec_dev */
int topo_diag_dev(topo_t *h, double *diag_dev);
/* c = alpha*a + beta*b over n_nodes nodes (2 doubles each); a, b, c may alias */
int topo_vec_axpby_dev(topo_t *h, long long n_nodes, double alpha, const double *a_dev, double beta,
                       const double *b_dev, double *c_dev);
/* c = a (*) b elementwise, or c = (b != 0 ? a / b : 0) when divide != 0 */
int topo_vec_mul_dev(topo_t *h, long long n_nodes, const double *a_dev, const double *b_dev, double *c_dev, int divide);
/* out_dev[0] = sum over node rows [row_lo, row_hi) of a.b  (ownership-restricted dot; deterministic) */
int topo_vec_dot_dev(topo_t *h, const double *a_dev, const double *b_dev, int row_lo, int row_hi, double *out_dev);

/* ---- row-slab SIMP iteration over several GPUs (BASELINE configs[4]) ------------------------
 * The whole design iteration -- multigrid-preconditioned CG, sensitivity, filter, OC -- on a mesh split into
 * `world` equal slabs of element rows, one handle (one GPU, one process) per slab.  The library never
 * communicates: every call below runs one SEGMENT of the algorithm on the slab and names, in its comment, what
 * the host must move before the next segment.  Three kinds of movement, all on device memory, all on the
 * handle's stream (solidspy_opt/distributed.py does them with NCCL through torch.distributed):
 *   exchange   send my first interface row / block to rank-1 and my last one to rank+1, receive theirs into the
 *              handle's receive buffers (topo_dist_xchg_ptrs), then topo_dist_xchg_finish;
 *   gather     all-gather of an equally sized block per rank (topo_dist_gather_ptrs);
 *   all-reduce in place on a few device doubles (topo_dist_scalar_ptr).
 * Vectors are stored ACCUMULATED (interface rows hold the full value on both neighbours); the fused kernels
 * produce DISTRIBUTED results on interface rows (partial sums), which the exchange completes.  The multigrid
 * levels 1..K-1 live on the slab; levels >= K are small, are assembled from the gathered level-K Galerkin cell
 * matrices and run redundantly on every rank inside the tail kernel.  Mathematically the preconditioner is the
 * single-GPU one, so iteration counts agree with topo_solve on the whole mesh.
 * No counterpart in the reference (single process; scipy spsolve at optimize.py:437). */
enum { TOPO_X_AP = 0, TOPO_X_R0 = 1, TOPO_X_Z = 2, TOPO_X_LEVEL_B = 3, TOPO_X_LEVEL_X2 = 4, TOPO_X_DIAG = 5, TOPO_X_FILTER = 6 };
enum { TOPO_G_CELLS = 0, TOPO_G_BK = 1 };
enum { TOPO_S_PAP = 0, TOPO_S_RR_RZ = 1, TOPO_S_BB = 2, TOPO_S_RED = 3 };
/* Turns an ordinary handle (nx x ny_local, BC flags / loads / rho of ITS slab) into slab `rank` of `world`.
 * K_request <= 0: K = first level of the global hierarchy with <= 10 000 nodes.  ny_local must be a multiple
 * of 2^K. */
int topo_dist_init(topo_t *h, int rank, int world, int K_request);
int topo_dist_info(topo_t *h, int *rank, int *world, int *K, int *n_levels);
/* device pointers of an exchange: send_lo -> rank-1 (lands in its recv_hi), send_hi -> rank+1 (its recv_lo) */
int topo_dist_xchg_ptrs(topo_t *h, int what, int level, void **send_lo, void **send_hi, void **recv_lo,
                        void **recv_hi, long long *n_doubles);
int topo_dist_xchg_finish(topo_t *h, int what, int level);
int topo_dist_gather_ptrs(topo_t *h, int what, void **send, void **recv, long long *n_doubles_per_rank);
int topo_dist_scalar_ptr(topo_t *h, int what, void **ptr, int *n);
/* first n doubles of the TOPO_S_RED block -> host (synchronises the stream) */
int topo_dist_read_scalars(topo_t *h, double *out, int n);
/* multigrid set-up after topo_simp_scale.  phase 0: slab-local Galerkin products (then: exchange TOPO_X_DIAG,
 * gather TOPO_G_CELLS); phase 1: global levels. */
int topo_dist_setup(topo_t *h, int phase);
/* r0 = f - K u (then: exchange TOPO_X_R0, all-reduce TOPO_S_BB) */
int topo_dist_solve_begin(topo_t *h, double rtol, int maxit, int warm);
/* one PCG iteration = matvec, update, V-cycle (down 1..K-1, tail, up K-1..1), smooth, finish; `par` alternates
 * 1, 0, 1, ... (buffer parity, as in topo_solve); the first pass is update/V-cycle/smooth/finish with init = 1. */
int topo_dist_pcg_matvec(topo_t *h, int par);              /* then: exchange TOPO_X_AP, all-reduce TOPO_S_PAP */
int topo_dist_pcg_update(topo_t *h, int par, int init);    /* then: K > 1 ? exchange TOPO_X_LEVEL_B level 1 : gather TOPO_G_BK */
int topo_dist_vcycle_down(topo_t *h, int level);           /* then: level+1 < K ? exchange TOPO_X_LEVEL_B level+1 : gather TOPO_G_BK */
int topo_dist_vcycle_tail(topo_t *h, int p2p);             /* p2p != 0: pieces arrived through topo_dist_p2p_gather */
int topo_dist_vcycle_up(topo_t *h, int level);             /* then: exchange TOPO_X_LEVEL_X2 level */
int topo_dist_pcg_smooth(topo_t *h, int par, int init);    /* then: exchange TOPO_X_Z, all-reduce TOPO_S_RR_RZ */
int topo_dist_pcg_finish(topo_t *h, int init);
int topo_dist_solve_poll(topo_t *h, int *done, int *iters, double *relres);
int topo_dist_solve_end(topo_t *h, int iters);
/* TOPO_S_RED[1] = this slab's share of f.u (TOPO_S_RED[0] = objective share from topo_simp_sensitivity(.., NULL)) */
int topo_dist_compliance_partial(topo_t *h);
/* rho*dc of the slab's outer element rows -> send buffers (then: exchange TOPO_X_FILTER, topo_filter_sensitivity) */
int topo_dist_filter_pack(topo_t *h, double rmin, double dy);
/* OC bisection pieces (op table in csrc/simp.cu); between them the host all-reduces TOPO_S_RED */
int topo_dist_oc(topo_t *h, int op, double move, double value, long long ne_global);
int topo_dist_oc_state(topo_t *h, int *done, int *need_exact, int *steps, double *gt, double *change, double *lmid);
/* Peer-memory transport (NVLink / NVSwitch) for the movements of the PCG loop: every slab owns a region its peers
 * map (CUDA IPC between processes) and write into, payload first, then a sequence number (st.release.sys); the
 * consumer kernel spins on its own memory (ld.acquire.sys, bounded).  Replaces the NCCL launch of a node-row
 * exchange / level-K gather / scalar all-reduce by two small kernels: phase 0 pushes and never waits, phase 1
 * waits for the peers' data and completes the operation.  Every rank must issue the same sequence of calls. */
int topo_dist_p2p_export(topo_t *h, void *ipc_handle_64, void **local_ptr);
int topo_dist_p2p_attach(topo_t *h, const void *ipc_handles /* world x 64 bytes */, void *const *local_ptrs /* or NULL */);
int topo_dist_p2p_xchg(topo_t *h, int phase, int what, int level);
int topo_dist_p2p_gather(topo_t *h, int phase);              /* TOPO_G_BK; then topo_dist_vcycle_tail(h, 1) */
int topo_dist_p2p_allreduce(topo_t *h, int phase, int what, int n, int is_max);
/* One-launch forms for ONE slab per process (a kernel that waits for its peers must not share a stream with them):
 * interface rows of `what` pushed, awaited and completed by one kernel, optionally with the all-reduce (sum) of the
 * first scal_n doubles of scalar block scal_what (scal_n = 0: none) and the PCG bookkeeping that follows the
 * (rr, rz) all-reduce (finish = 1; 2 for the initial pass; replaces topo_dist_pcg_finish; finish = 3: the scalars are
 * reduced with max instead of sum; what = -1: scalars only);
 * level-K pieces pushed, awaited, assembled, then the tail kernel (replaces topo_dist_p2p_gather x2 + topo_dist_vcycle_tail). */
int topo_dist_p2p_fused(topo_t *h, int what, int level, int scal_what, int scal_n, int finish);
int topo_dist_p2p_gather_tail(topo_t *h);
/* mode bit 0: the restriction kernels (topo_dist_pcg_update, topo_dist_vcycle_down) and the up kernel
 * (topo_dist_vcycle_up) swap their interface rows with the neighbours themselves; the host then issues no
 * TOPO_X_LEVEL_B / TOPO_X_LEVEL_X2 exchange.  One slab per process only. */
int topo_dist_p2p_mode(topo_t *h, int mode);   /* bit 0: in-kernel row swaps; bit 1 (with bit 0): the small slab-local levels
                                                 * and the level-K gather inside one persistent kernel per direction */
int topo_dist_p2p_error(topo_t *h, int *error);              /* 1: a wait gave up (peer missing) */

/* ---- instrumentation ----------------------------------------------------------------------- */
/* CUDA-event time (ms) of the most recent call of each stage on this handle. */
typedef struct {
  float solve_ms, matvec_ms, sens_ms, filter_ms, oc_ms, setup_ms, total_ms;
  long long kernel_launches; /* cumulative count of kernels launched by this handle */
} topo_timing;
int topo_get_timing(topo_t *h, topo_timing *t);
/* Device-side stopwatch on the handle's stream (CUDA events): bench.py brackets its timed region. */
int topo_timer_start(topo_t *h);
int topo_timer_stop(topo_t *h, float *ms);
/* Per-launch CUDA-event timing of the stage-1 kernel (all variants of the fine-grid K application)
 * while enabled; read returns the number of launches recorded and their summed duration. */
int topo_profile_matvec(topo_t *h, int enable);   /* switching off discards the recorded data: read the report first */
int topo_profile_read(topo_t *h, int *launches, float *total_ms);
/* While profiling is enabled EVERY kernel launch of the handle is bracketed by events; the report is a
 * text table "kernel count total_ms" (one line per kernel, sorted by time) written into buf. */
int topo_profile_report(topo_t *h, char *buf, size_t buflen);
/* GMG hierarchy description: number of levels and the (nx, ny) of each (up to 32). */
int topo_gmg_info(topo_t *h, int *n_levels, int *nx_levels, int *ny_levels);
/* Solver knobs: omega = Jacobi damping of the V-cycle smoother, nu = sweeps pre/post. */
int topo_set_gmg(topo_t *h, double omega, int nu_pre, int nu_post);

#ifdef __cplusplus
}
#endif
#endif /* TOPO_B200_H */
