/* cslvr_b200.h -- C-ABI of the B200-native Blatter-Pattyn momentum path.
 *
 * The reference (pf4d/cslvr) has no FFI: its hot path leaves Python at ONE call,
 *
 *     dolfin.solve(F == 0, u, J=J, bcs=bcs, solver_parameters=params['solver'])
 *                                                   cslvr/physics.py:673-678
 *
 * (equivalently assemble_system + solve(A,x,b,method,pc), cslvr/model.py:2679,2698).
 * Everything below is what a ctypes binding at that seam needs: plain pointers
 * and sizes, an opaque context, integer return codes (0 = ok) and a message via
 * bp_last_error().  No torch types cross this boundary.  Host arrays are
 * borrowed for the duration of the call; device arrays passed in are owned by
 * the caller (PyTorch tensors in the Python glue) and never freed here.
 */
#ifndef CSLVR_B200_H
#define CSLVR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BP_OK              0
#define BP_ERR_ARG         1   /* bad argument / inconsistent mesh or dof map   */
#define BP_ERR_CUDA        2   /* CUDA runtime error (no GPU, OOM, launch fail) */
#define BP_ERR_STATE       3   /* call out of order (field not set, no J yet)   */
#define BP_ERR_COMM        4   /* NCCL could not be loaded / collective failed  */
#define BP_ERR_NONCONV     5   /* Newton (or a Krylov solve) did not converge and the matching
                                   error_on_nonconvergence flag is set                       */

/* field ids for bp_set_field.  All coefficient fields are VERTEX-indexed
 * (n = n_vertices): a P1 Function f on a dolfin space with dofmap dm is passed
 * as f_vertex[cells[c][a]] = f.vector()[dm.cell_dofs(c)[a]].                     */
#define BP_FIELD_S      0   /* surface height S      cslvr/model.py:2511, momentumbp.py:467 */
#define BP_FIELD_A      1   /* rate factor A         cslvr/model.py:2530, momentum.py:181   */
#define BP_FIELD_BETA   2   /* basal traction beta   cslvr/model.py:2527, momentumbp.py:473 */
#define BP_FIELD_B      3   /* bed height B          cslvr/model.py:2512, momentumbp.py:65 (u_z only)   */
#define BP_FIELD_B_RING 4   /* basal melt B_ring     momentumbp.py:66 (u_z only; default 0)             */
#define BP_FIELD_TP     5   /* pressure-adjusted temperature T'   } energy-dependent rate factor,      */
#define BP_FIELD_W      6   /* water content W (default 0)        } Model.form_energy_dependent_rate_  */
#define BP_FIELD_E      7   /* enhancement factor E (default 1)   } factor, cslvr/model.py:1815-1859   */
#define BP_FIELD_UZ     8   /* frozen vertical velocity u_z of MomentumDukowiczStokesReduced (default 0),
                               cslvr/momentumstokes.py:361, 406-418; read iff bp_params.reduced_stokes   */

/* what bp_assemble computes */
#define BP_ASSEMBLE_F   1
#define BP_ASSEMBLE_J   2
#define BP_TIME_CELL_ONLY 4 /* bp_time_assemble only: launch the cell kernel alone (no zero-fill,
                               no facet kernel); the resident Jacobian is invalid afterwards  */

/* preconditioners for the Krylov solve (reference: PETSc CG + hypre BoomerAMG,
 * cslvr/momentumbp.py:183-184; here a device-resident PCG) */
#define BP_PC_NONE          0
#define BP_PC_JACOBI        1   /* point Jacobi                                   */
#define BP_PC_BLOCK_JACOBI  2   /* 2x2 (u_x,u_y) node-block Jacobi                */
#define BP_PC_CHEBYSHEV     3   /* Chebyshev polynomial in D^-1 J (D = 2x2 node blocks), degree
                                   pc_degree, lmax by power iteration once per Newton step     */
#define BP_PC_COLUMN        4   /* vertical-line solve: exact inverse of the block-tridiagonal
                                   coupling inside every ice column (extruded meshes)          */

/* how the element tensors reach the resident Jacobian */
#define BP_ASM_AUTO     0   /* tiled on extruded prism meshes, else row gather unless a block row
                               is too long for shared memory                                    */
#define BP_ASM_SCATTER  1   /* one thread per tetrahedron, FP64 atomics into F and J           */
#define BP_ASM_GATHER   2   /* one thread per Jacobian block row, no atomics, deterministic    */
#define BP_ASM_TILED    3   /* extruded prism meshes: one thread per prism column, every tetrahedron
                               evaluated once, tile-local sums in shared memory, no atomics,
                               deterministic (what AUTO picks when the mesh qualifies; asking for
                               it on a mesh that does not is an error)                          */

#define BP_PC_MG            5   /* aggregation multigrid V-cycle: ice columns -> footprint nodes ->
                                   greedy 2-D aggregates, Chebyshev smoothing (pc_degree, pc_eig_ratio),
                                   Galerkin coarse operators rebuilt on the device per Newton step     */

typedef struct bp_ctx bp_ctx;

/* The mesh and the P1xP1 dof map: what the adapter pulls out of dolfin
 * (mesh.coordinates(), mesh.cells(), Q.dofmap().cell_dofs, model.ff) -- or what
 * cslvr_b200.D3Model generates for BoxMesh.  Replaces D3Model's mesh/marker
 * state consumed by the forms: cslvr/d3model.py:337-475, 477-626, 629-667.      */
typedef struct {
	int64_t        n_vertices;
	int64_t        n_cells;
	int64_t        n_dofs;          /* length of the caller's u vector                     */
	const double*  coords;          /* [n_vertices*3] x,y,z per vertex (deformed mesh)      */
	const int32_t* cells;           /* [n_cells*4] vertex ids                               */
	const int32_t* cell_dofs;       /* [n_cells*8] component-major (UFC mixed P1xP1:
	                                   0..3 = u_x, 4..7 = u_y), or NULL for the canonical
	                                   numbering dof = 2*vertex_to_node[v] + c              */
	const int32_t* vertex_to_node;  /* [n_vertices] periodic identification (d3model.py:
	                                   62-108), or NULL = identity; used iff cell_dofs NULL */
	int64_t        n_facets;        /* exterior facets carrying a marker                    */
	const int32_t* facet_cell;      /* [n_facets] cell of the facet                         */
	const int32_t* facet_local;     /* [n_facets] local facet = opposite local vertex       */
	const int32_t* facet_marker;    /* [n_facets] D3Model marker ids (d3model.py:16-27):
	                                   3/5 basal grounded/floating, 4/10 lateral; others
	                                   are ignored                                          */
	/* ---- partition (one context per GPU/rank); single rank: n_owned_nodes = -1 ---- */
	int64_t        n_owned_nodes;   /* nodes [0,n_owned) are owned, the rest are ghosts;
	                                   requires cell_dofs == NULL                           */
	int32_t        n_neighbors;
	const int32_t* neighbor_rank;   /* [n_neighbors]                                        */
	const int64_t* send_ptr;        /* [n_neighbors+1] into send_nodes                      */
	const int32_t* send_nodes;      /* owned local nodes whose values neighbour k needs     */
	const int64_t* recv_ptr;        /* [n_neighbors+1] ghost nodes from neighbour k are
	                                   n_owned + [recv_ptr[k], recv_ptr[k+1])               */
} bp_mesh;

/* Physical constants (cslvr/model.py:212-236) and solver parameters
 * (cslvr/momentumbp.py:177-203; dolfin NewtonSolver / PETSc KSP defaults).       */
typedef struct {
	double  n;                 /* Glen exponent                      model.py:215 */
	double  eps_reg;           /* strain-rate regularisation         model.py:212 */
	double  rho_i;             /* ice density                        model.py:221 */
	double  rho_sw;            /* sea-water density                  model.py:227 */
	double  g;                 /* gravitational acceleration         model.py:236 */
	int32_t periodic;          /* model.use_periodic                              */
	int32_t use_pressure_bc;   /* momentumbp.py:488: lateral water pressure iff
	                              !periodic && use_pressure_bc                    */
	/* quadrature for the only non-polynomial factor, A(x)^(-1/n):
	 * barycentric points [n_quad*4] and weights [n_quad] summing to 1            */
	int32_t       n_quad;
	const double* quad_points;
	const double* quad_weights;
	/* Newton (dolfin NewtonSolver semantics)                                    */
	double  newton_rtol;       /* 'relative_tolerance'   momentumbp.py:185 (1e-9) */
	double  newton_atol;       /* dolfin default 1e-10                            */
	int32_t newton_maxit;      /* 'maximum_iterations'   momentumbp.py:187 (25)   */
	double  newton_relax;      /* 'relaxation_parameter' momentumbp.py:186        */
	int32_t error_on_nonconvergence;          /* momentumbp.py:188 (False)       */
	/* Krylov (PETSc KSPCG defaults: rtol 1e-5, atol 1e-50, maxit 1e4)           */
	double  krylov_rtol;
	double  krylov_atol;
	int32_t krylov_maxit;
	int32_t pc;                /* BP_PC_*                                         */
	int32_t pc_degree;         /* Chebyshev degree                                */
	int32_t verbose;
	int32_t assembly;          /* BP_ASM_*                                        */
	double  pc_eig_ratio;      /* Chebyshev: interval [lmax/ratio, lmax] of D^-1 J (100)     */
	/* quadrature of the projection right-hand sides (solve_pressure: UFL degree 6)          */
	int32_t       n_quad_aux;
	const double* quad_points_aux;
	const double* quad_weights_aux;
	/* model.energy_dependent_rate_factor: A = E a_T (1 + 181.25 min(W, .01)) exp(-Q_T/(R T')) at
	 * the quadrature points instead of the P1 field A (constants cslvr/model.py:275-285, 263)   */
	int32_t energy_dependent;
	double  a_T_l, a_T_u, Q_T_l, Q_T_u, R_gas;
	/* CG stopping test: 0 = true residual |r|_2 <= max(rtol |b|_2, atol) (default: the Newton
	 * step is then resolved to rtol whatever the preconditioner), 1 = PETSc's default
	 * preconditioned norm |M^-1 r|_2 (only safe with a spectrally tight preconditioner)         */
	int32_t krylov_norm_preconditioned;
	int32_t mg_smoother;       /* BP_PC_MG level 0: 0 = Chebyshev on 2x2 node blocks, 1 = Chebyshev on exact
	                              ice-column solves (line smoother), -1 = line iff the tets are flat
	                              (median horizontal/vertical extent >= 3; default)              */
	/* MomentumDukowiczStokesReduced (cslvr/momentumstokes.py:300-400) instead of the Blatter-Pattyn
	 * forms: BP_FIELD_UZ enters the strain-rate tensor (eps_02 = (u_x,z + u_z,x)/2, eps_12 likewise),
	 * sliding acts on the grounded bed only (dGamma_bg) and carries u_z_b = (-B_ring - u_h.n_h)/n_z   */
	int32_t reduced_stokes;
	/* BP_PC_MG: factor on every coarse-grid correction (0 = default: 1.8, or 1.0 on the slabs of a
	 * partitioned mesh).  The piecewise-constant
	 * prolongation gives a coarse operator that is too stiff, so the plain correction under-shoots;
	 * any factor in (0, 2] keeps the V-cycle symmetric positive definite.                          */
	double  mg_coarse_scale;
	/* dolfin's PETScKrylovSolver raises when KSP stops without converging ('error_on_nonconvergence'
	 * of the krylov_solver parameter set, default True; cslvr/momentumbp.py:190 leaves it alone):
	 * != 0 -> a linear solve that reaches krylov_maxit returns BP_ERR_NONCONV; 0 -> the step is taken
	 * anyway and counted in bp_stats.krylov_unconverged                                            */
	int32_t krylov_error_on_nonconvergence;
} bp_params;

typedef struct {
	int32_t converged;
	int32_t newton_iterations;
	int32_t krylov_iterations;       /* total over all Newton steps               */
	double  residual_norm0;          /* ||F(u0)||_2                               */
	double  residual_norm;           /* final ||F(u)||_2                          */
	double  t_assemble_ms;           /* device time, CUDA events                  */
	double  t_krylov_ms;
	double  t_total_ms;
	int32_t n_assemblies;
	int32_t gpu_launches;            /* kernels launched by this call             */
	double  residual_history[64];
	int32_t krylov_unconverged;      /* linear solves that stopped at krylov_maxit            */
	double  krylov_last_relres;      /* relative residual of the last linear solve            */
	int32_t host_launches;           /* launches the host issued for them (a graph replay of two CG iterations counts once) */
} bp_stats;

/* ---- life cycle ---------------------------------------------------------- */
int  bp_create(bp_ctx** out, const bp_mesh* mesh, const bp_params* params, int device);
void bp_destroy(bp_ctx* ctx);
const char* bp_last_error(const bp_ctx* ctx);      /* ctx may be NULL: last create error */
int  bp_set_params(bp_ctx* ctx, const bp_params* params);
/* Stream contract.  bp_create gives the context its own non-blocking stream; every kernel and copy
 * of the library runs there, and every entry point returns only after its outputs are complete
 * (it synchronises that stream).  Device buffers handed in (on_device != 0) are READ on the context's
 * stream: work that produces them on another stream must have completed (or been ordered by the
 * caller) before the call.  bp_set_stream adopts the caller's stream instead (e.g. torch's current
 * stream), which makes that ordering automatic and lets the caller time / graph-capture the work.  */
int  bp_set_stream(bp_ctx* ctx, void* cuda_stream); /* run on the caller's stream        */

/* coefficient Functions: model.init_beta / init_A / init_S,
 * cslvr/model.py:486-497, 691-733 -> assign_variable model.py:2166-2221           */
int  bp_set_field(bp_ctx* ctx, int field, const double* values, int64_t n, int on_device);

/* Dirichlet conditions of the momentum solve: DirichletBC(Q.sub(0), u_x_lat, ff, GAMMA_L_DVD) and its
 * u_y twin, cslvr/momentumbp.py:101-112 (MomentumBP with use_lat_bcs on a mesh with mark_divide).  dofs
 * are in the caller's numbering (ghost dofs of a partitioned mesh included), values the prescribed
 * velocities.  bp_newton_solve / bp_linear_solve then treat those rows as dolfin's NewtonSolver does
 * (b_i = u_i - g_i, identity rows: the update carries u_i to g_i; columns eliminated as well so that CG
 * keeps a symmetric matrix).  bp_assemble keeps returning the unconstrained F and J, as assemble() does.
 * n = 0 removes the conditions.                                                                         */
int  bp_set_dirichlet(bp_ctx* ctx, int64_t n, const int32_t* dofs, const double* values);
/* ... and of bp_solve_vert_velocity: DirichletBC(Q, u_z_lat, ff, GAMMA_L_DVD) (cslvr/momentumbp.py:110-111), by mesh vertex;
 * applied after the basal conditions, so it wins on the nodes they share (bc.apply in list order).  n = 0 removes them. */
int  bp_set_dirichlet_uz(bp_ctx* ctx, int64_t n, const int32_t* vertices, const double* values);


/* sizes after the mesh has been digested */
int  bp_get_sizes(const bp_ctx* ctx, int64_t* n_nodes, int64_t* n_dofs, int64_t* nnz_csr,
                  int64_t* nnz_blocks, int64_t* n_cells, int64_t* n_basal_facets);

/* ---- sparsity: dolfin SparsityPatternBuilder / PETSc AIJ layout -------------
 * rows = caller's dofs, columns sorted ascending, explicit zeros kept.
 * The arrays are host memory owned by the context.                                */
int  bp_csr_pattern(bp_ctx* ctx, int64_t* nnz, const int64_t** indptr, const int32_t** indices);

/* ---- assemble(F), assemble(J): replaces FFC tabulate_tensor + dolfin Assembler
 * for cslvr/momentumbp.py:481-500 (action -> residual) and physics.py:75-77
 * (Jacobian).  u, F: caller's dof numbering, length n_dofs.  J_values: aligned
 * with bp_csr_pattern.  Any output may be NULL.  *_on_device: pointers are device
 * pointers on the context's GPU.  The Jacobian also stays resident in the context
 * (for bp_spmv / bp_krylov_solve).                                                 */
int  bp_assemble(bp_ctx* ctx, int what, const double* u, double* F, double* J_values,
                 int on_device);

/* y = J x with the resident Jacobian (PETSc MatMult)                              */
int  bp_spmv(bp_ctx* ctx, const double* x, double* y, int on_device);

/* solve J dx = b with the resident Jacobian (PETSc KSPCG + PC)                    */
int  bp_krylov_solve(bp_ctx* ctx, const double* b, double* dx, int on_device,
                     int32_t* iterations, double* rel_residual);

/* the whole dolfin.solve(F == 0, u, J=J) call: damped Newton from u_inout         */
int  bp_newton_solve(bp_ctx* ctx, double* u_inout, int on_device, bp_stats* stats);

/* ---- linear=True (cslvr/momentumbp.py:91-93, 396, 503; physics.py:655-661; used by
 * Momentum.linearize_viscosity, momentum.py:97-138): ONE linear solve lhs == rhs with the
 * viscosity frozen at eta(u_lin) (u_lin = model.u).  u_out = K(u_lin)^-1 (-f).            */
int  bp_linear_solve(bp_ctx* ctx, const double* u_lin, double* u_out, int on_device, int32_t* iterations);

/* ---- cslvr/momentumbp.py:205-229, MomentumBPBase.solve_pressure ------------------
 * p = project(rho g (S - z) - 2 eta(u) (u_x,x + u_y,y), model.Q): P1 mass-matrix solve
 * (dolfin project() = LU; here block-Jacobi PCG to 1e-13).  u: caller's dofs;
 * p_vertex: one value per mesh vertex.  The resident Jacobian is overwritten.          */
int  bp_solve_pressure(bp_ctx* ctx, const double* u, double* p_vertex, int on_device, int32_t* iterations);

/* ---- cslvr/momentumbp.py:60-85, 231-253, MomentumBPBase.solve_vert_velocity ----------
 * u_z from continuity exactly as the reference poses it: u_z_b = project(basal kinematic
 * condition, Q); A_ij = int dz(phi_j) phi_i, b_i = -int (u_x,x + u_y,y) phi_i; Dirichlet
 * rows u_z = u_z_b on the basal facets' dofs; the reference factorises A with MUMPS, here
 * BiCGStab right-preconditioned by an exact pivoted tridiagonal solve per ice column, to a
 * relative residual of 1e-12.  Needs BP_FIELD_B; uz_vertex: one value per mesh vertex.    */
int  bp_solve_vert_velocity(bp_ctx* ctx, const double* u, double* uz_vertex, int on_device, int32_t* iterations);

/* ---- MomentumDukowiczStokesReduced.solve, cslvr/momentumstokes.py:449-486 ------------------
 * Model.home_rolled_newton_method (cslvr/model.py:2610-2730) with solve_vert_velocity as the
 * per-iteration callback: loop { assemble J, b = -R at (u, u_z); solve J d = b; r = |b|;
 * converged = r < atol or r/r0 < rtol; u += relax d; u_z = solve_vert_velocity(u) }.
 * Needs reduced_stokes != 0 and BP_FIELD_B.  uz_vertex_inout: the starting u_z (NULL = keep the
 * one set with BP_FIELD_UZ) and, on return, the u_z of the last callback (per mesh vertex).
 * run_callback = 0 leaves u_z frozen (cb_ftn = None).  Partitioned contexts: one rank per process with a
 * communicator (bp_comm_init); in-process groups are not served.                                 */
int  bp_rs_newton_solve(bp_ctx* ctx, double* u_inout, double* uz_vertex_inout, int on_device,
                        double atol, int run_callback, bp_stats* stats);
/* solve_pressure as that solve() reaches it (momentumstokes.py:478 runs BEFORE update_model_var):
 * p = project(rho g (S - z) - 2 eta(u_model, u_z) (u_x,x + u_y,y), Q) with the viscosity of the
 * velocity the model holds (self.eta, cslvr/momentumbp.py:92, 219) and the divergence of u.    */
int  bp_rs_solve_pressure(bp_ctx* ctx, const double* u, const double* u_model, double* p_vertex,
                          int on_device, int32_t* iterations);

/* ---- timing of the kernels themselves (CUDA events on the context's stream) --- */
/* flush_l2 != 0: a 512 MiB buffer is written before every timed launch            */
int  bp_time_assemble(bp_ctx* ctx, int what, int repeats, int flush_l2, double* ms_per_launch);
int  bp_time_spmv(bp_ctx* ctx, int repeats, int flush_l2, double* ms_per_launch);
int64_t bp_launch_count(const bp_ctx* ctx);        /* kernels launched so far      */

/* ---- multi-GPU: one context per rank, NCCL only for halo + reductions --------- */
int  bp_comm_unique_id(const char* nccl_library_path, char id_out[128]);
int  bp_comm_init(bp_ctx* ctx, const char* nccl_library_path, const char id[128],
                  int rank, int n_ranks);
/* several contexts living in ONE process (one per device, or all on one device to
 * exercise the partitioned path on a single GPU): solved in lock-step, halo and
 * reductions by device copies.                                                     */
int  bp_newton_solve_group(bp_ctx** ctxs, int n, double** u_inout, int on_device, bp_stats* stats);
int  bp_assemble_group(bp_ctx** ctxs, int n, int what, const double** u, double** F, int on_device);

/* ---- debug hooks of the tiled cell kernel (BP_ASM_TILED) -- used by the CPU test suite, never by the
 * product path.  bp_debug_create_host digests a mesh WITHOUT a device (no compute entry point works on
 * such a context) and keeps the tile tables on the host; bp_debug_tiles_emulate walks them exactly as the
 * kernel does (same element math, same order of additions) and returns F (internal node order) and the
 * 2x2 blocks of J in BSR row order (bp_debug_bsr_pattern).  vert_xyzS: [n_vertices*4] x, y, z, S;
 * u: [2*n_nodes] in node order; ainv: [n_cells] sum_q w_q A_q^(-1/n).                                  */
int  bp_debug_create_host(bp_ctx** out, const bp_mesh* mesh, const bp_params* params);
int  bp_debug_bsr_pattern(const bp_ctx* ctx, int64_t* n_rows, const int32_t** rowptr, const int32_t** col);
int  bp_debug_tiles_emulate(bp_ctx* ctx, const double* vert_xyzS, const double* u, const double* ainv,
                            int picard, double* F, double* J_blocks);
/* ... plus the basal facets (markers 3, 5) the kernel folds into its first layer; beta: [n_vertices] */
int  bp_debug_tiles_emulate_facets(bp_ctx* ctx, const double* vert_xyzS, const double* u, const double* ainv, const double* beta,
                                   int picard, double* F, double* J_blocks);
int  bp_debug_tiles_info(const bp_ctx* ctx, int64_t* n_tiles, int64_t* n_prism_columns,
                         int64_t* n_threads_used, int64_t* n_entries, int64_t* n_contributions);
const char* bp_debug_tiles_why(const bp_ctx* ctx);   /* why the mesh is not served by the tiled kernel ("" if it is) */

#ifdef __cplusplus
}
#endif
#endif
