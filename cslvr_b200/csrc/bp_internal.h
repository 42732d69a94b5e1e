// bp_internal.h -- context layout shared by the translation units of libcslvr_b200.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "cslvr_b200.h"

#define BP_MAX_QUAD 128
#ifndef BP_ROWS_PER_CTA
#define BP_ROWS_PER_CTA 128       // block rows per CTA of the row-gather assembly kernel
#endif

// ---- device scalar block used by the Krylov loop (doubles) ------------------
enum {
	SC_T0 = 0,      // {r.z, z.z, r.r} written by even updates / the init   } one all-reduce each
	SC_T1 = 3,      // {r.z, z.z, r.r} written by odd updates               }
	SC_PAP = 6,     // p.Ap
	SC_NORM0,       // reference norm for the relative test
	SC_NORM,        // current norm
	SC_DONE,        // 1.0 once converged / broken down
	SC_ITERS,       // iterations performed
	SC_BREAKDOWN,   // 1.0 if p.Ap <= 0 or NaN
	SC_FF,          // F.F (Newton residual norm squared)
	SC_TMP,
	SC_LMAX,        // bound of lmax(D^-1 A) (multigrid levels: read by the Chebyshev kernels on the device)
	SC_COUNT = 16
};

// run-time switches, read from the environment ONCE per context at bp_create -- never per launch
struct bp_env {
	int    pipe_k = 8;              // BP_PIPE_K
	bool   no_pipeline = false;     // BP_NO_PIPELINE
	bool   facet_atomic = false;    // BP_FACET_ATOMIC=1: the atomic facet kernel (default: row gather, reproducible)
	bool   has_horizontal_first = false;
	double mg_horizontal_first = 1.0;                       // BP_MG_HORIZONTAL_FIRST
	bool   mg_fp32 = true;          // BP_MG_FP32
	bool   has_alpha0 = false, has_alpha = false;
	double mg_alpha0 = 0.0, mg_alpha = 0.0;                 // BP_MG_ALPHA0 / BP_MG_ALPHA
	bool   mg_coupled = true, mg_coupled_graph = true;      // BP_MG_COUPLED / BP_MG_COUPLED_GRAPH
	int    mg_gamma = 0;            // BP_MG_GAMMA (0 = the hierarchy's default)
	int    mg_tail = 4096;          // BP_MG_TAIL
	int    mg_repl = 32768;         // BP_MG_REPL: coupled hierarchy: levels with at most this many nodes (globally) are replicated on every rank (0: only the dense coarsest level)
	double mg_lmax_scale = 1.0;     // BP_MG_LMAX_SCALE: factor on the Gershgorin bound of lmax(D^-1 A) the multigrid smoothers use (experiments)
	int    asm_tiled = -1;          // BP_ASM_TILED: 1 / 0 force / forbid the tiled cell kernel (-1 = automatic)
	bool   cg_graph = true;         // BP_CG_GRAPH=0: launch the CG iteration directly instead of replaying its graph
	bool   pipe_graph = true;       // BP_PIPE_GRAPH=0: issue the pipelined host-buffer bp_assemble stream by stream instead of replaying its graph
	bool   p2p = true;              // BP_P2P=0: halo exchanges and scalar all-reduces through NCCL instead of peer memory
};
void bp_env_read(bp_env* e);

// BP_SETUP_TRACE=1: wall time of the host set-up stages on stderr
struct bp_stage_timer {
	double t0; bool on; const char* what;
	bp_stage_timer(const char* w);
	void lap(const char* stage);
};

// constants of the element kernels (built from bp_params per launch)
struct AsmParams {
	double n, eps_reg, rho_g, rho_sw_g;
	int    n_quad;
	int    n_is_3;
	int    picard;      // linear=True: viscosity frozen at the given state, no eta' term, F = u-independent part
	int    rs;          // MomentumDukowiczStokesReduced: u_z in the strain rate, u_z_b in the sliding term
};
struct bp_tiles;        // tiled cell assembly (bp_tiles.cu)

struct bp_comm_state;   // NCCL binding (bp_comm.cu)
struct bp_p2p_halo;     // peer-memory exchange of one partitioned context (bp_comm.cu)
struct bp_mg;           // multigrid hierarchy (bp_mg.cu)

struct bp_ctx {
	int          device = 0;
	cudaStream_t stream = nullptr;
	bool         own_stream = false;
	std::string  err;
	bp_params    prm{};
	bp_env       env;
	std::vector<double> quad_pts, quad_wts, quad_pts_aux, quad_wts_aux;

	// sizes
	int64_t nv = 0, nt = 0, nd_ext = 0;
	int64_t nn = 0;          // local nodes (owned + ghost)
	int64_t nn_owned = 0;
	int64_t nnzb = 0;        // 2x2 blocks in the resident Jacobian (owned rows)
	int64_t nf = 0;          // facets that integrate something (basal and/or pressure)
	int64_t n_basal = 0;
	bool    identity_dofs = true;    // caller's dof == 2*node + c
	bool    identity_v2n = true;
	std::vector<int32_t> h_v2n;          // vertex -> internal node (kept on the host for bp_set_dirichlet_uz; empty = identity)

	// host topology kept for the lazy CSR export
	std::vector<int32_t> h_ext_dof;      // [2*nn] caller's dof of (node, c): [2*i+c]
	std::vector<int32_t> h_rowptr, h_col;
	std::vector<int64_t> csr_indptr;
	std::vector<int32_t> csr_indices;
	bool    have_csr = false;

	// device: mesh + coefficients
	double4* d_geo = nullptr;            // per vertex {x, y, z, S}
	double*  d_A = nullptr;              // per vertex
	double*  d_beta = nullptr;           // per vertex
	double  *d_Tp = nullptr, *d_W = nullptr, *d_E = nullptr;   // energy-dependent rate factor inputs, per vertex
	double*  d_ainv = nullptr;           // [nt] sum_q w_q A_q^(-1/n) per tet (k_tet_ainv)
	bool     ainv_dirty = true, have_Tp = false, smem_attr_set = false;
	double  *d_Bv = nullptr, *d_Bring = nullptr;   // bed height / basal melt per vertex (u_z solve only)
	double*  d_uz = nullptr;             // frozen vertical velocity per vertex (reduced-Stokes model, BP_FIELD_UZ)
	double*  d_uvert2 = nullptr;         // [2*nv] second per-vertex velocity (viscosity state of bp_rs_solve_pressure)
	uint8_t* d_bcnode = nullptr;         // owned node lies on a basal facet (Dirichlet dof of the u_z system) or carries a lateral u_z condition
	std::vector<uint8_t> h_bcnode;       // the basal ones (bp_set_dirichlet_uz adds to a copy of it)
	int32_t* d_uzbc_node = nullptr; double* d_uzbc_val = nullptr; int64_t n_uzbc = 0;   // lateral Dirichlet nodes of the u_z solve and their values
	double*  d_aux = nullptr;            // [2*nn] u_z_b per node
	int4*    d_tets = nullptr;
	int32_t* d_v2n = nullptr;            // nullptr when identity
	int32_t* d_tet_blk = nullptr;        // [16][nt] block slot of (a,b), -1 = row not owned
	int32_t* d_inc_slice = nullptr;      // row-gather incidences, SELL-32: slice offsets
	int32_t* d_inc_code = nullptr;       //   tet*4 + local vertex, -1 = padding
	uint32_t* d_inc_pos = nullptr;       //   4 x uint8 block positions inside the row
	int32_t* d_inc_v = nullptr;          //   [4][slots] vertex ids, the row's own vertex first
	int64_t  n_inc_slots = 0;
	double*  d_uvert = nullptr;          // [2*nv] u per vertex (only when vertices != nodes)
	double*  d_soa = nullptr;            // [4][nv] x | y | z | S (struct-of-arrays copy of d_geo)
	double2* d_zS = nullptr;             // per vertex {z, S}: what changes along an ice column (tiled cell kernel)
	int32_t  max_row = 0;                // longest block row
	int32_t* d_fac_cell = nullptr;       // [nf]
	int32_t* d_fac_info = nullptr;       // [nf] local facet | basal<<8 | pressure<<9 | grounded<<10
	int32_t* d_ext_dof = nullptr;        // nullptr when identity
	// facet row gather: owned nodes touched by an integrating facet, CSR of their (facet * 4 + local vertex) pairs
	int64_t  n_fnodes = 0;
	int32_t *d_fn_node = nullptr, *d_fn_ptr = nullptr, *d_fn_inc = nullptr;
	// device: resident Jacobian, BSR 2x2, owned rows
	int32_t* d_sell_ptr = nullptr;       // [slices+1] first slot of every 32-row slice
	int32_t* d_rowlen = nullptr;         // [nn_owned] blocks in the row
	int64_t  n_slots = 0;                // slots incl. padding
	std::vector<int32_t> h_sell_ptr;
	int32_t* d_col = nullptr;
	int32_t* d_diag = nullptr;           // slot of the diagonal block of each owned row
	int32_t* d_mirror = nullptr;         // [n_slots] slot of the transposed block (row j, column i) of block (i, j), j > i owned; else -1
	double*  d_val = nullptr;            // [n_slots*4] 2x2 blocks {a00 a01 a10 a11}, SELL-32 slot order
	float*   d_valf = nullptr;           // FP32 copy of d_val for the multigrid smoother (allocated on first use)
	int32_t* d_csr_src = nullptr;        // CSR export gather map
	double*  d_csr_out = nullptr;        // CSR values staged for a host-side J_values (allocated on first use)
	// device: vectors in internal layout [2*node + c], length 2*nn
	double *d_u = nullptr, *d_F = nullptr, *d_dx = nullptr, *d_r = nullptr, *d_z = nullptr,
	       *d_p = nullptr, *d_Ap = nullptr, *d_io = nullptr, *d_io2 = nullptr;
	double  *d_z2 = nullptr, *d_cd = nullptr;   // Chebyshev ping-pong iterate and direction
	// vertical-line preconditioner: owned nodes grouped by ice column, bottom to top
	int64_t  n_cols = 0;
	int32_t  max_layers = 0, col_touch_max = 4;
	double   aspect = 1.0;               // median horizontal / vertical extent of the tets
	int32_t *d_colptr = nullptr, *d_colnode = nullptr, *d_tri_d = nullptr, *d_tri_u = nullptr, *d_tri_l = nullptr;
	double   cheb_lmax = 0.0;
	double*  d_dinv = nullptr;           // [nn_owned*4] preconditioner blocks
	double*  d_sc = nullptr;             // SC_COUNT scalars
	double*  d_part = nullptr;           // block partials [4][max_blocks]
	unsigned* d_counter = nullptr;       // last-block tickets
	double*  h_sc = nullptr;             // pinned mirror of d_sc
	// halo
	int32_t n_nbr = 0;
	std::vector<int32_t> nbr_rank;
	std::vector<int64_t> send_ptr, recv_ptr;
	int32_t* d_send_nodes = nullptr;
	std::vector<int32_t> h_send_nodes;
	double*  d_sendbuf = nullptr;
	int64_t  n_send = 0;
	// Dirichlet rows of the momentum solve (bp_set_dirichlet): 1.0 / 0.0 and the prescribed value per internal dof
	double  *d_bcmask = nullptr, *d_bcval = nullptr;
	bool     have_bc = false;
	bp_comm_state* comm = nullptr;
	bp_p2p_halo* p2p = nullptr;          // peer-memory exchange (nullptr: NCCL send / recv)
	bool     p2p_tried = false;          // bp_comm_p2p_attach ran for this context (it is collective: once)
	bp_mg*   mg = nullptr;
	std::vector<int32_t> h_fcell, h_finfo;   // integrating facets on the host (the tiled kernel folds the basal ones into its first layer)
	bp_tiles* tiles = nullptr;           // tiled cell assembly: nullptr when the mesh is not an extruded prism mesh (k_rows serves it)
	std::string tiles_why;               // why not
	bool     host_only = false;          // debug contexts without a device (index-table emulation in the CPU test suite)
	std::vector<int32_t> h_colptr, h_colnode;      // ice columns on the host (first multigrid aggregation)
	int      grid_spmv0 = 0, grid_spmv1 = 0, grid_cheb = 0;   // one-wave launch grids, per context

	bool have_S = false, have_A = false, have_beta = false, have_J = false, have_B = false;
	int64_t launches = 0;
	int64_t host_launches = 0;           // graph launches of the CG loop (each replays two iterations)
	cudaGraphExec_t cg_graph = nullptr;  // two CG iterations (both parities), captured once (bp_api.cu::group_pcg)
	unsigned char cg_key[96] = { 0 };    // what that graph was captured for
	int64_t cg_launches = 0;             // kernels inside it
	bool    cg_no_graph = false;         // the stream cannot be captured: launch directly
	bool    outer_capture = false;       // group_pcg is capturing: the V-cycle joins that capture instead of replaying its own graph
	int     krylov_unconverged = 0;      // linear solves that stopped at krylov_maxit (bp_stats)
	double  krylov_last_relres = 0.0;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	// host-buffer bp_assemble, pipelined: per CTA of the row-gather kernel, the vertex / node prefix its
	// incidences reference (prefix maxima) and whether a facet integral touches one of its rows
	std::vector<int32_t> h_blk_need_vert, h_blk_need_node;
	std::vector<uint8_t> h_blk_facet;
	std::vector<uint8_t> h_facet_node;   // owned node gets a facet integral (its F row is final after the facet kernel only)
	cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_aux = nullptr;
	std::vector<cudaEvent_t> pipe_ev;
	double* h_halo_stage = nullptr;      // pinned: boundary values of a host-side u for the neighbours (pipelined bp_assemble of a partitioned context)
};

// ---- helpers -------------------------------------------------------------
int bp_fail(bp_ctx* c, int code, const char* fmt, ...);
#define BP_CUDA(ctx, call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return bp_fail(ctx, BP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// setup (bp_setup.cu)
int bp_setup_topology(bp_ctx* c, const bp_mesh* m);
int bp_build_csr_export(bp_ctx* c);

// kernels (bp_kernels.cu)
void bp_launch_assemble(bp_ctx* c, int what);                 // d_u -> d_F / d_val
bool bp_rows_gather_mode(const bp_ctx* c);                    // the row-gather kernel serves this context
void bp_launch_rows_chunk(bp_ctx* c, int what, int blk0, int nblk, int v0, int v1);   // CTAs [blk0, blk0+nblk) of k_rows; u expanded for vertices [v0, v1) first
void bp_launch_facets(bp_ctx* c, int what);
void bp_update_ainv(bp_ctx* c);
void bp_launch_spmv(bp_ctx* c, const double* x, double* y, int with_dot);
void bp_launch_gather_in(bp_ctx* c, const double* ext, double* internal);   // caller -> internal
void bp_launch_scatter_out(bp_ctx* c, const double* internal, double* ext); // internal -> caller
void bp_launch_csr_export(bp_ctx* c, double* out);
void bp_launch_pc_setup(bp_ctx* c);
void bp_launch_set_S(bp_ctx* c, const double* S_dev);
void bp_launch_pack_halo(bp_ctx* c, const double* vec);
void bp_launch_uz_lateral_values(bp_ctx* c, double* node_vec);
void bp_launch_bc_rows(bp_ctx* c, double sign);   // constrained rows of d_F <- sign * u - g (d in d_io for bp_launch_bc_eliminate)
void bp_launch_bc_eliminate(bp_ctx* c);           // free rows lifted by J d, rows / columns of the constrained dofs -> identity
void bp_launch_dot_owned(bp_ctx* c, const double* a, const double* b, int sc_slot);
void bp_launch_axpy(bp_ctx* c, double alpha, const double* x, double* y);   // y += alpha x
void bp_launch_pcg_init(bp_ctx* c);
void bp_launch_pcg_update(bp_ctx* c, int par);
void bp_launch_pcg_p(bp_ctx* c, int first, int par);
void bp_launch_sum_scalars(bp_ctx** ctxs, int n, int slot, int count);
void bp_launch_pcg_init2(bp_ctx* c);                       // x = 0, r = b (preconditioner applied separately)
void bp_launch_pcg_update2(bp_ctx* c, int par);            // x += alpha p, r -= alpha Ap
void bp_launch_dots3(bp_ctx* c, int slot, int guard);      // {r.z, z.z, r.r} -> d_sc[slot..]
void bp_launch_col_solve(bp_ctx* c, const double* r, double* z);
// Chebyshev step on the column splitting: delta = T^-1 res; d = c1 d + c2 delta; znew = zold + d
void bp_launch_col_cheb(bp_ctx* c, double c1, double c2, const double* res, const double* zold, double* znew);
// lmax_dev != nullptr: c0 / c2 are given for lmax = 1 and scaled by 1 / *lmax_dev on the device
void bp_launch_cheb_first(bp_ctx* c, double c0, const double* r, double* z, const double* lmax_dev = nullptr);
void bp_launch_cheb_step(bp_ctx* c, double c1, double c2, const double* r, const double* zold, double* znew, const double* lmax_dev = nullptr, bool f32 = false);
int  bp_launch_val_to_float(bp_ctx* c);                    // d_val -> d_valf (FP32 copy read by the multigrid V-cycle)
void bp_launch_residual(bp_ctx* c, const double* b, const double* z, double* out, bool f32);   // out = b - A z
void bp_launch_apply_dinv(bp_ctx* c, double* v);           // v <- D^-1 v
void bp_launch_scale_inv_norm(bp_ctx* c, double* v, int sc_slot);   // v <- v / sqrt(d_sc[slot])
void bp_launch_fill_probe(bp_ctx* c, double* v);
void bp_launch_gershgorin(bp_ctx* c, int sc_slot);
void bp_launch_rows_aux(bp_ctx* c, int kind, const double2* u_eta_vertex = nullptr, const double* uz_vertex = nullptr);              // projection systems (mass matrix + rhs) -> d_val, d_F
void bp_launch_expand_to_vertex(bp_ctx* c, const double* node_vec2, double* vertex_vec2);   // [2*node+c] -> [2*vertex+c]
void bp_launch_node_to_vertex(bp_ctx* c, const double* node_vec, int comp, double* vertex_out);
void bp_launch_col_gtsv(bp_ctx* c, const double* r, double* z);   // scalar pivoted tridiagonal solve per column
void bp_launch_lincomb(bp_ctx* c, double* out, double a, const double* x, double b, const double* y, double cc, const double* z);        // upper bound of lmax(D^-1 J) -> d_sc[slot]

AsmParams bp_asm_params(const bp_ctx* c);

// tiled cell assembly (bp_tiles.cu)
int  bp_setup_tiles(bp_ctx* c, const bp_mesh* m, const std::vector<int32_t>& v2n);
bool bp_tiles_serve(const bp_ctx* c, int what);
bool bp_tiles_facet_lists(const bp_ctx* c, int64_t* n_fnodes, const int32_t** fn_node, const int32_t** fn_ptr, const int32_t** fn_inc);   // facets left for the facet kernel when the tiled kernel folds the basal ones in           // the tiled kernel serves this assembly request
void bp_launch_tiles(bp_ctx* c, int what);                // d_u -> d_F / d_val (cell integrals)
void bp_tiles_update_ainv(bp_ctx* c);                     // d_ainv -> tile order (after k_tet_ainv)
void bp_tiles_free(bp_ctx* c);
void bp_tiles_params_changed(bp_ctx* c);                  // drop the captured host-buffer pipeline (its kernels carry the constants)
bool bp_tiles_pipeline_ready(const bp_ctx* c, int what);  // host-buffer bp_assemble can run the tiled kernel pipelined
int  bp_tiles_assemble_pipelined(bp_ctx* c, int what, const double* u, double* F);

// multigrid (bp_mg.cu)
int  bp_mg_prepare(bp_ctx* c);      // numeric set-up for the current resident Jacobian
int  bp_mg_apply(bp_ctx* c);        // d_z = M^-1 d_r
void bp_mg_free(bp_ctx* c);
// coupled hierarchy of a partitioned mesh: all ranks (or all contexts of an in-process group) in lock-step
bool bp_mg_use_coupled(bp_ctx** cs, int n);
int  bp_mg_prepare_group(bp_ctx** cs, int n);
int  bp_mg_apply_group(bp_ctx** cs, int n);

// comm (bp_comm.cu)
int  bp_comm_halo(bp_ctx** ctxs, int n, int which_vec);   // 0:u 1:p
int  bp_comm_allreduce(bp_ctx** ctxs, int n, int slot, int count);
int  bp_comm_allreduce_max(bp_ctx* c, int slot, int count);
// element-wise sum (or max) of one device buffer per context over all ranks: NCCL for a single context with a
// communicator, a device kernel for an in-process group
int  bp_comm_allreduce_buf(bp_ctx** ctxs, int n, double** bufs, size_t count, int op_max);
int  bp_comm_rank(const bp_ctx* c);      // rank / number of ranks of the context's communicator (0 / 1 without one)
int  bp_comm_size(const bp_ctx* c);
void bp_comm_free(bp_ctx* c);
bool bp_mg_settled(bp_ctx** cs, int n, const void** id);   // the V-cycle's launch sequence is final: an enclosing stream capture is safe
int  bp_comm_p2p_attach(bp_ctx* c);   // collective: landing buffer of a partitioned context in the peer-mapped arena
void bp_comm_p2p_detach(bp_ctx* c);
double* bp_vec(bp_ctx* c, int which);
