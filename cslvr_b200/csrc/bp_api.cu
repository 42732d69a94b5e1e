// bp_api.cu -- the extern "C" boundary (include/cslvr_b200.h) and the host side of
// the Newton-Krylov loop.
//
// bp_newton_solve is the drop-in for `dolfin.solve(F == 0, u, J=J, ...)` at
// cslvr/physics.py:673-678 with dolfin NewtonSolver semantics:
//     b = F(u); r0 = |b|
//     while not (|b|/r0 < rtol or |b| < atol) and it < maxit:
//         assemble J(u); solve J dx = b; u <- u - relax dx; b = F(u)
// and a PETSc-KSPCG-like inner solve (zero initial guess, rtol 1e-5 by default; the stopping
// test is on the true residual norm unless krylov_norm_preconditioned asks for PETSc's default).
#include "bp_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

std::string g_bp_create_error;
void bp_upload_quadrature(bp_ctx* c);

void bp_env_read(bp_env* e)
{
	auto geti = [](const char* k, int d) { const char* v = getenv(k); return v ? atoi(v) : d; };
	auto getd = [](const char* k, double d) { const char* v = getenv(k); return v ? atof(v) : d; };
	e->pipe_k = geti("BP_PIPE_K", 8);
	e->no_pipeline = getenv("BP_NO_PIPELINE") != nullptr;
	e->facet_atomic = geti("BP_FACET_ATOMIC", 0) != 0;
	e->has_horizontal_first = getenv("BP_MG_HORIZONTAL_FIRST") != nullptr;
	e->mg_horizontal_first = getd("BP_MG_HORIZONTAL_FIRST", 1.0);
	e->mg_fp32 = geti("BP_MG_FP32", 1) != 0;
	e->has_alpha0 = getenv("BP_MG_ALPHA0") != nullptr; e->mg_alpha0 = getd("BP_MG_ALPHA0", 0.0);
	e->has_alpha = getenv("BP_MG_ALPHA") != nullptr;   e->mg_alpha = getd("BP_MG_ALPHA", 0.0);
	e->mg_coupled = geti("BP_MG_COUPLED", 1) != 0;
	e->mg_coupled_graph = geti("BP_MG_COUPLED_GRAPH", 1) != 0;
	e->mg_gamma = geti("BP_MG_GAMMA", 0);
	e->mg_tail = geti("BP_MG_TAIL", 4096);
	e->mg_lmax_scale = getd("BP_MG_LMAX_SCALE", 1.0);
	e->mg_repl = geti("BP_MG_REPL", 32768);
	e->asm_tiled = geti("BP_ASM_TILED", -1);
	e->cg_graph = geti("BP_CG_GRAPH", 1) != 0;
	e->p2p = geti("BP_P2P", 1) != 0;
	e->pipe_graph = geti("BP_PIPE_GRAPH", 1) != 0;
}

#include <chrono>
static double wall_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
bp_stage_timer::bp_stage_timer(const char* w) : t0(wall_now()), on(getenv("BP_SETUP_TRACE") != nullptr), what(w) {}
void bp_stage_timer::lap(const char* stage)
{
	if (!on) return;
	const double t = wall_now();
	fprintf(stderr, "[cslvr_b200] %s: %-28s %.3f s\n", what, stage, t - t0);
	t0 = t;
}

// ------------------------------------------------------------------ helpers
static int check_params(bp_ctx* c, const bp_params* p)
{
	if (!p) return bp_fail(c, BP_ERR_ARG, "bp_params is NULL");
	if (!(p->n > 0) || !(p->eps_reg >= 0) || !(p->rho_i > 0) || !(p->g > 0))
		return bp_fail(c, BP_ERR_ARG, "bp_params: n, rho_i, g must be positive and eps_reg non-negative");
	if (p->n_quad < 1 || p->n_quad > BP_MAX_QUAD || !p->quad_points || !p->quad_weights)
		return bp_fail(c, BP_ERR_ARG, "bp_params: n_quad must be in [1,%d] with points and weights", BP_MAX_QUAD);
	double s = 0;
	for (int q = 0; q < p->n_quad; ++q) s += p->quad_weights[q];
	if (std::fabs(s - 1.0) > 1e-12) return bp_fail(c, BP_ERR_ARG, "bp_params: quadrature weights sum to %.17g, not 1", s);
	if (p->n_quad_aux < 0 || p->n_quad_aux > BP_MAX_QUAD) return bp_fail(c, BP_ERR_ARG, "bp_params: n_quad_aux must be in [0,%d]", BP_MAX_QUAD);
	if (p->assembly < BP_ASM_AUTO || p->assembly > BP_ASM_TILED) return bp_fail(c, BP_ERR_ARG, "bp_params: unknown assembly mode %d", p->assembly);
	if (p->pc < BP_PC_NONE || p->pc > BP_PC_MG) return bp_fail(c, BP_ERR_ARG, "bp_params: unknown preconditioner %d", p->pc);
	return BP_OK;
}

static void store_params(bp_ctx* c, const bp_params* p)
{
	c->prm = *p;
	c->quad_pts.assign(p->quad_points, p->quad_points + 4 * p->n_quad);
	c->quad_wts.assign(p->quad_weights, p->quad_weights + p->n_quad);
	c->prm.quad_points = c->quad_pts.data();
	c->prm.quad_weights = c->quad_wts.data();
	if (p->n_quad_aux > 0 && p->quad_points_aux && p->quad_weights_aux) {
		c->quad_pts_aux.assign(p->quad_points_aux, p->quad_points_aux + 4 * p->n_quad_aux);
		c->quad_wts_aux.assign(p->quad_weights_aux, p->quad_weights_aux + p->n_quad_aux);
	} else { c->quad_pts_aux = c->quad_pts; c->quad_wts_aux = c->quad_wts; c->prm.n_quad_aux = c->prm.n_quad; }
	c->prm.quad_points_aux = c->quad_pts_aux.data();
	c->prm.quad_weights_aux = c->quad_wts_aux.data();
}

struct GroupStream {   // an in-process group runs on the first context's stream
	bp_ctx** cs; int n; std::vector<cudaStream_t> saved;
	GroupStream(bp_ctx** cs_, int n_) : cs(cs_), n(n_) { for (int r = 0; r < n; ++r) { saved.push_back(cs[r]->stream); cs[r]->stream = cs[0]->stream; } }
	~GroupStream() { for (int r = 0; r < n; ++r) cs[r]->stream = saved[r]; }
};

static int ready(bp_ctx* c)
{
	const bool haveA = c->prm.energy_dependent ? c->have_Tp : c->have_A;
	if (!c->have_S || !haveA || !c->have_beta)
		return bp_fail(c, BP_ERR_STATE, "fields not set: S=%d %s=%d beta=%d (bp_set_field)", (int)c->have_S,
		               c->prm.energy_dependent ? "T'" : "A", (int)haveA, (int)c->have_beta);
	return BP_OK;
}

// caller vector (host or device, caller numbering) -> internal vector
static int vec_in(bp_ctx* c, const double* src, int on_device, double* dst)
{
	const double* dev = src;
	if (!on_device) {
		BP_CUDA(c, cudaMemcpyAsync(c->d_io, src, (size_t)c->nd_ext * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		dev = c->d_io;
	}
	bp_launch_gather_in(c, dev, dst);
	return BP_OK;
}
static int vec_out(bp_ctx* c, const double* src, double* dst, int on_device)
{
	if (on_device) { bp_launch_scatter_out(c, src, dst); return BP_OK; }
	bp_launch_scatter_out(c, src, c->d_io2);
	BP_CUDA(c, cudaMemcpyAsync(dst, c->d_io2, (size_t)c->nd_ext * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	return BP_OK;
}

static int read_scalars(bp_ctx* c)
{
	BP_CUDA(c, cudaMemcpyAsync(c->h_sc, c->d_sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	return BP_OK;
}

// ------------------------------------------------------------------ group core
static int group_assemble(bp_ctx** cs, int n, int what)
{
	int rc = bp_comm_halo(cs, n, 0);
	if (rc) return rc;
	for (int r = 0; r < n; ++r) bp_launch_assemble(cs[r], what);
	return BP_OK;
}

static int group_norm_F(bp_ctx** cs, int n, double* out)
{
	for (int r = 0; r < n; ++r) bp_launch_dot_owned(cs[r], cs[r]->d_F, cs[r]->d_F, SC_FF);
	int rc = bp_comm_allreduce(cs, n, SC_FF, 1);
	if (rc) return rc;
	rc = read_scalars(cs[0]);
	if (rc) return rc;
	*out = std::sqrt(cs[0]->h_sc[SC_FF]);
	return BP_OK;
}

// ---- preconditioner set-up, once per Jacobian ---------------------------------------
static int group_pc_prepare(bp_ctx** cs, int n)
{
	int rc;
	for (int r = 0; r < n; ++r) bp_launch_pc_setup(cs[r]);
	if (cs[0]->prm.pc == BP_PC_CHEBYSHEV) {
		// upper bound of lmax(D^-1 J): Gershgorin row bound, max over all ranks
		double lmax = 0.0;
		for (int r = 0; r < n; ++r) bp_launch_gershgorin(cs[r], SC_TMP);
		if (n == 1 && (rc = bp_comm_allreduce_max(cs[0], SC_TMP, 1))) return rc;
		for (int r = 0; r < n; ++r) { if ((rc = read_scalars(cs[r]))) return rc; lmax = std::max(lmax, cs[r]->h_sc[SC_TMP]); }
		if (!(lmax > 0.0) || !std::isfinite(lmax)) return bp_fail(cs[0], BP_ERR_STATE, "Chebyshev: eigenvalue bound failed (%g)", lmax);
		for (int r = 0; r < n; ++r) cs[r]->cheb_lmax = lmax;
	}
	if (cs[0]->prm.pc == BP_PC_MG) {
		if (bp_mg_use_coupled(cs, n)) { if ((rc = bp_mg_prepare_group(cs, n))) return rc; }
		else for (int r = 0; r < n; ++r) if ((rc = bp_mg_prepare(cs[r]))) { if (r) cs[0]->err = cs[r]->err; return rc; }
	}
	if (cs[0]->prm.pc == BP_PC_COLUMN)
		for (int r = 0; r < n; ++r) if (cs[r]->max_layers > 64)
			return bp_fail(cs[0], BP_ERR_ARG, "column preconditioner supports at most 64 nodes per ice column (mesh has %d)", cs[r]->max_layers);
	return BP_OK;
}

// z = M^-1 r for the preconditioners that are not fused into the CG update kernel
static int group_pc_apply(bp_ctx** cs, int n)
{
	int rc;
	if (cs[0]->prm.pc == BP_PC_COLUMN) {
		for (int r = 0; r < n; ++r) bp_launch_col_solve(cs[r], cs[r]->d_r, cs[r]->d_z);
		return BP_OK;
	}
	if (cs[0]->prm.pc == BP_PC_MG) {
		if (bp_mg_use_coupled(cs, n)) return bp_mg_apply_group(cs, n);
		for (int r = 0; r < n; ++r) if ((rc = bp_mg_apply(cs[r]))) return rc;
		return BP_OK;
	}
	// Chebyshev on [b / ratio, b], b = 1.1 lmax
	const int deg = std::max(cs[0]->prm.pc_degree, 1);
	const double b = cs[0]->cheb_lmax, a = b / std::max(cs[0]->prm.pc_eig_ratio, 1.5);
	const double theta = 0.5 * (b + a), delta = 0.5 * (b - a), sigma = theta / delta;
	double rho = 1.0 / sigma;
	for (int r = 0; r < n; ++r) bp_launch_cheb_first(cs[r], 1.0 / theta, cs[r]->d_r, cs[r]->d_z);
	for (int k = 1; k < deg; ++k) {
		const double rho_n = 1.0 / (2.0 * sigma - rho);
		if ((rc = bp_comm_halo(cs, n, 2))) return rc;
		for (int r = 0; r < n; ++r) { bp_launch_cheb_step(cs[r], rho_n * rho, 2.0 * rho_n / delta, cs[r]->d_r, cs[r]->d_z, cs[r]->d_z2); std::swap(cs[r]->d_z, cs[r]->d_z2); }
		rho = rho_n;
	}
	return BP_OK;
}

// PCG on J dx = F (rhs in d_F, solution in d_dx)
static int group_pcg(bp_ctx** cs, int n, int* iters, double* relres)
{
	int rc;
	const bool fused = cs[0]->prm.pc <= BP_PC_BLOCK_JACOBI;       // pointwise: applied inside the update kernel
	if ((rc = group_pc_prepare(cs, n))) return rc;
	if (fused) {
		for (int r = 0; r < n; ++r) bp_launch_pcg_init(cs[r]);
	} else {
		for (int r = 0; r < n; ++r) bp_launch_pcg_init2(cs[r]);
		if ((rc = group_pc_apply(cs, n))) return rc;
		for (int r = 0; r < n; ++r) {
			cudaMemcpyAsync(cs[r]->d_p, cs[r]->d_z, (size_t)2 * cs[r]->nn_owned * sizeof(double), cudaMemcpyDeviceToDevice, cs[r]->stream);
			bp_launch_dots3(cs[r], SC_T0, 0);
		}
	}
	if ((rc = bp_comm_allreduce(cs, n, SC_T0, 3))) return rc;
	for (int r = 0; r < n; ++r) bp_launch_pcg_p(cs[r], 1, 0);
	const int maxit = cs[0]->prm.krylov_maxit > 0 ? cs[0]->prm.krylov_maxit : 10000;
	// one CG iteration; par = parity of the iteration (the device-resident scalars alternate between two sets)
	auto iterate = [&](const int par) -> int {
		int rc_;
		if ((rc_ = bp_comm_halo(cs, n, 1))) return rc_;
		for (int r = 0; r < n; ++r) bp_launch_spmv(cs[r], cs[r]->d_p, cs[r]->d_Ap, 1);
		if ((rc_ = bp_comm_allreduce(cs, n, SC_PAP, 1))) return rc_;
		if (fused) {
			for (int r = 0; r < n; ++r) bp_launch_pcg_update(cs[r], par);
		} else {
			for (int r = 0; r < n; ++r) bp_launch_pcg_update2(cs[r], par);
			if ((rc_ = group_pc_apply(cs, n))) return rc_;
			for (int r = 0; r < n; ++r) bp_launch_dots3(cs[r], par ? SC_T0 : SC_T1, 1);
		}
		if ((rc_ = bp_comm_allreduce(cs, n, par ? SC_T0 : SC_T1, 3))) return rc_;
		for (int r = 0; r < n; ++r) bp_launch_pcg_p(cs[r], 0, par);
		return BP_OK;
	};
	// The iteration is a fixed launch sequence (every scalar, the convergence flag and the Chebyshev bounds live on the
	// device; a converged solve turns the remaining kernels of a batch into no-ops), so two iterations (both parities) are
	// captured ONCE as a CUDA graph -- the V-cycle graph becomes a child node, the peer-memory / NCCL exchanges are captured
	// with it -- and the host issues one graph launch per two iterations instead of ~12 launches + 1 graph.  The first two
	// iterations of a solve run directly when the graph does not exist yet: the multigrid captures its own graph there.
	bp_ctx* c0 = cs[0];
	struct Key { const void* inner; const void* stream; double lmax, ratio, alpha, rtol, atol; int pc, deg, smoother, n, f32, norm_pc; } key;   // everything a captured kernel argument depends on
	memset(&key, 0, sizeof(key));
	const void* inner = nullptr;
	const bool settled = (c0->prm.pc != BP_PC_MG) || bp_mg_settled(cs, n, &inner);
	key.inner = inner; key.stream = (const void*)c0->stream; key.lmax = (c0->prm.pc == BP_PC_CHEBYSHEV) ? c0->cheb_lmax : 0.0;
	key.ratio = c0->prm.pc_eig_ratio; key.alpha = c0->prm.mg_coarse_scale; key.pc = c0->prm.pc; key.deg = c0->prm.pc_degree;
	key.smoother = c0->prm.mg_smoother; key.n = n; key.f32 = c0->env.mg_fp32 ? 1 : 0;
	key.rtol = c0->prm.krylov_rtol; key.atol = c0->prm.krylov_atol; key.norm_pc = c0->prm.krylov_norm_preconditioned;      // arguments of k_pcg_p
	if (c0->cg_graph && memcmp(&key, c0->cg_key, sizeof(key)) != 0) { cudaGraphExecDestroy(c0->cg_graph); c0->cg_graph = nullptr; }
	static_assert(sizeof(Key) <= sizeof(c0->cg_key), "bp_ctx::cg_key too small");
	int k = 0;
	bool done = false;
	while (k < maxit && !done) {
		const int batch = std::min(fused ? 32 : 8, maxit - k);
		int i = 0;
		while (i < batch) {
			const bool pair_ok = c0->env.cg_graph && !c0->cg_no_graph && ((k + i) & 1) == 0 && i + 2 <= batch;
			if (pair_ok && !c0->cg_graph && (k + i >= 2 || settled)) {
				// the preconditioner's own graph may have been (re)built by the direct iterations: look again
				const void* in2 = nullptr;
				if (c0->prm.pc != BP_PC_MG || bp_mg_settled(cs, n, &in2)) {
					key.inner = in2;
					std::vector<int64_t> l0(n);
					for (int r = 0; r < n; ++r) l0[r] = cs[r]->launches;
					cudaGraph_t g = nullptr;
					cudaError_t eb = cudaStreamBeginCapture(c0->stream, cudaStreamCaptureModeThreadLocal);
					if (eb == cudaSuccess) {
						for (int r = 0; r < n; ++r) cs[r]->outer_capture = true;
						rc = iterate(0);
						if (!rc) rc = iterate(1);
						for (int r = 0; r < n; ++r) cs[r]->outer_capture = false;
						eb = cudaStreamEndCapture(c0->stream, &g);
						if (rc == BP_OK && eb == cudaSuccess) eb = cudaGraphInstantiate(&c0->cg_graph, g, 0);
						if (g) cudaGraphDestroy(g);
						if (rc != BP_OK || eb != cudaSuccess) c0->cg_graph = nullptr;
					}
					if (!c0->cg_graph) {    // not capturable here: keep launching directly
						if (c0->prm.verbose) fprintf(stderr, "[cslvr_b200] CG iteration graph: capture failed (rc %d, %s); launching directly\n", rc, cudaGetErrorString(eb));
						cudaGetLastError(); c0->cg_no_graph = true; rc = BP_OK;
					}
					else {
						c0->cg_launches = 0;
						for (int r = 0; r < n; ++r) c0->cg_launches += cs[r]->launches - l0[r];
						memcpy(c0->cg_key, &key, sizeof(key));
					}
					for (int r = 0; r < n; ++r) cs[r]->launches = l0[r];
				}
			}
			if (pair_ok && c0->cg_graph) {
				BP_CUDA(c0, cudaGraphLaunch(c0->cg_graph, c0->stream));
				c0->launches += c0->cg_launches;
				c0->host_launches++;
				i += 2;
			} else {
				int64_t l0 = 0, l1 = 0;
				for (int r = 0; r < n; ++r) l0 += cs[r]->launches;
				if ((rc = iterate((k + i) & 1))) return rc;
				for (int r = 0; r < n; ++r) l1 += cs[r]->launches;
				c0->host_launches += l1 - l0;      // (a V-cycle graph inside counts with all its kernels: an upper bound)
				++i;
			}
		}
		k += batch;
		if ((rc = read_scalars(cs[0]))) return rc;
		done = cs[0]->h_sc[SC_DONE] != 0.0;
	}
	if (iters) *iters = (int)cs[0]->h_sc[SC_ITERS];
	const double rr = cs[0]->h_sc[SC_NORM0] > 0 ? cs[0]->h_sc[SC_NORM] / cs[0]->h_sc[SC_NORM0] : 0.0;
	if (relres) *relres = rr;
	cs[0]->krylov_last_relres = rr;
	if (cs[0]->h_sc[SC_BREAKDOWN] != 0.0)
		return bp_fail(cs[0], BP_ERR_STATE, "CG breakdown: p.Ap <= 0 (Jacobian not positive definite or NaN)");
	if (!done) {
		// PETSc KSP_DIVERGED_ITS: dolfin raises unless the krylov_solver's error_on_nonconvergence is off
		cs[0]->krylov_unconverged++;
		if (cs[0]->prm.krylov_error_on_nonconvergence)
			return bp_fail(cs[0], BP_ERR_NONCONV, "Krylov solver did not converge in %d iterations (relative residual %.3e, rtol %.3e)",
			               (int)cs[0]->h_sc[SC_ITERS], rr, cs[0]->prm.krylov_rtol);
	}
	return BP_OK;
}

static int group_newton(bp_ctx** cs, int n, bp_stats* st)
{
	bp_ctx* c0 = cs[0];
	const bp_params& P = c0->prm;
	bp_stats s;
	memset(&s, 0, sizeof(s));
	int rc;
	int64_t l0 = 0;
	const int unconv0 = c0->krylov_unconverged;
	const int64_t hl0 = c0->host_launches;
	for (int r = 0; r < n; ++r) { if ((rc = ready(cs[r]))) { if (r) c0->err = cs[r]->err; return rc; } l0 += cs[r]->launches; }
	cudaEvent_t eA, eB, eT0, eT1;
	cudaEventCreate(&eA); cudaEventCreate(&eB); cudaEventCreate(&eT0); cudaEventCreate(&eT1);
	float ms;
	cudaEventRecord(eT0, c0->stream);
	auto timed_assemble = [&](int what) -> int {
		cudaEventRecord(eA, c0->stream);
		int r_ = group_assemble(cs, n, what);
		cudaEventRecord(eB, c0->stream);
		if (r_) return r_;
		cudaEventSynchronize(eB); cudaEventElapsedTime(&ms, eA, eB); s.t_assemble_ms += ms; s.n_assemblies++;
		return BP_OK;
	};
	double rn = 0.0;
	// Dirichlet rows (bp_set_dirichlet): b_i = u_i - g_i before the norm, symmetric elimination before the linear solve
	auto bc_rows = [&]() { for (int r = 0; r < n; ++r) if (cs[r]->have_bc) bp_launch_bc_rows(cs[r], 1.0); };
	auto bc_eliminate = [&]() { for (int r = 0; r < n; ++r) if (cs[r]->have_bc) bp_launch_bc_eliminate(cs[r]); };
	rc = timed_assemble(BP_ASSEMBLE_F | BP_ASSEMBLE_J);
	if (!rc) { bc_rows(); rc = group_norm_F(cs, n, &rn); }
	if (rc) goto out;
	s.residual_norm0 = rn;
	s.residual_history[0] = rn;
	if (!std::isfinite(rn)) { rc = bp_fail(c0, BP_ERR_STATE, "initial residual is not finite"); goto out; }
	s.converged = (rn < P.newton_atol);
	while (!s.converged && s.newton_iterations < P.newton_maxit) {
		int kit = 0;
		bc_eliminate();
		cudaEventRecord(eA, c0->stream);
		rc = group_pcg(cs, n, &kit, nullptr);
		cudaEventRecord(eB, c0->stream);
		cudaEventSynchronize(eB); cudaEventElapsedTime(&ms, eA, eB); s.t_krylov_ms += ms;
		if (rc) goto out;
		s.krylov_iterations += kit;
		for (int r = 0; r < n; ++r) bp_launch_axpy(cs[r], -P.newton_relax, cs[r]->d_dx, cs[r]->d_u);
		s.newton_iterations++;
		if ((rc = timed_assemble(BP_ASSEMBLE_F | BP_ASSEMBLE_J))) goto out;
		bc_rows();
		if ((rc = group_norm_F(cs, n, &rn))) goto out;
		if (s.newton_iterations < 64) s.residual_history[s.newton_iterations] = rn;
		if (P.verbose)
			printf("Newton iteration %d: r (abs) = %.3e (tol = %.3e) r (rel) = %.3e (tol = %.3e)  [krylov %d]\n",
			       s.newton_iterations, rn, P.newton_atol, rn / s.residual_norm0, P.newton_rtol, kit);
		if (!std::isfinite(rn)) { rc = bp_fail(c0, BP_ERR_STATE, "residual became non-finite at Newton iteration %d", s.newton_iterations); goto out; }
		s.converged = (rn / s.residual_norm0 < P.newton_rtol) || (rn < P.newton_atol);
	}
	s.residual_norm = rn;
	if (!s.converged && P.error_on_nonconvergence)
		rc = bp_fail(c0, BP_ERR_NONCONV, "Newton solver did not converge in %d iterations (|F| = %.3e)", s.newton_iterations, rn);
out:
	cudaEventRecord(eT1, c0->stream);
	cudaEventSynchronize(eT1);
	cudaEventElapsedTime(&ms, eT0, eT1);
	s.t_total_ms = ms;
	cudaEventDestroy(eA); cudaEventDestroy(eB); cudaEventDestroy(eT0); cudaEventDestroy(eT1);
	int64_t l1 = 0;
	for (int r = 0; r < n; ++r) l1 += cs[r]->launches;
	s.gpu_launches = (int32_t)(l1 - l0);
	s.host_launches = (int32_t)(c0->host_launches - hl0);
	s.krylov_unconverged = c0->krylov_unconverged - unconv0;
	s.krylov_last_relres = c0->krylov_last_relres;
	if (st) *st = s;
	return rc;
}

// ------------------------------------------------------------------ C-ABI
extern "C" {

const char* bp_last_error(const bp_ctx* c) { return c ? c->err.c_str() : g_bp_create_error.c_str(); }

int bp_create(bp_ctx** out, const bp_mesh* mesh, const bp_params* params, int device)
{
	if (!out || !mesh) return bp_fail(nullptr, BP_ERR_ARG, "bp_create: NULL argument");
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0)
		return bp_fail(nullptr, BP_ERR_CUDA, "bp_create: no CUDA device (%s); this library has no CPU path", cudaGetErrorString(e));
	if (device < 0 || device >= ndev) return bp_fail(nullptr, BP_ERR_ARG, "bp_create: device %d of %d", device, ndev);
	bp_ctx* c = new bp_ctx();
	c->device = device;
	bp_env_read(&c->env);
	int rc = check_params(c, params);
	if (!rc) {
		store_params(c, params);
		e = cudaSetDevice(device);
		if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
		if (e != cudaSuccess) rc = bp_fail(c, BP_ERR_CUDA, "bp_create: %s", cudaGetErrorString(e));
		c->own_stream = true;
	}
	if (!rc) rc = bp_setup_topology(c, mesh);
	if (!rc) { cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1); bp_upload_quadrature(c); }
	if (rc) { g_bp_create_error = c->err; bp_destroy(c); return rc; }
	*out = c;
	return BP_OK;
}

void bp_destroy(bp_ctx* c)
{
	if (!c) return;
	if (c->host_only) { bp_tiles_free(c); delete c; return; }
	cudaSetDevice(c->device);
	bp_tiles_free(c);
	bp_mg_free(c);
	bp_comm_free(c);
	void* ptrs[] = { c->d_geo, c->d_A, c->d_beta, c->d_tets, c->d_v2n, c->d_tet_blk, c->d_inc_slice, c->d_inc_code, c->d_inc_pos, c->d_inc_v, c->d_soa, c->d_zS, c->d_uvert, c->d_fac_cell, c->d_fac_info, c->d_fn_node, c->d_fn_ptr, c->d_fn_inc, c->d_ext_dof,
		c->d_sell_ptr, c->d_rowlen, c->d_col, c->d_diag, c->d_mirror, c->d_val, c->d_valf, c->d_csr_src, c->d_csr_out, c->d_u, c->d_F, c->d_dx, c->d_r, c->d_z, c->d_p, c->d_Ap,
		c->d_io, c->d_io2, c->d_bcmask, c->d_bcval, c->d_uzbc_node, c->d_uzbc_val, c->d_z2, c->d_cd, c->d_aux, c->d_Bv, c->d_Bring, c->d_uz, c->d_uvert2, c->d_bcnode, c->d_Tp, c->d_W, c->d_E, c->d_ainv, c->d_colptr, c->d_colnode, c->d_tri_d, c->d_tri_u, c->d_tri_l, c->d_dinv, c->d_sc, c->d_part, c->d_counter, c->d_send_nodes, c->d_sendbuf };
	for (void* p : ptrs) if (p) cudaFree(p);
	if (c->h_sc) cudaFreeHost(c->h_sc);
	if (c->h_halo_stage) cudaFreeHost(c->h_halo_stage);
	if (c->ev0) cudaEventDestroy(c->ev0);
	if (c->ev1) cudaEventDestroy(c->ev1);
	if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
	if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
	if (c->s_aux) cudaStreamDestroy(c->s_aux);
	for (auto e : c->pipe_ev) if (e) cudaEventDestroy(e);
	if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

/* ---- debug: a context WITHOUT a device.  It digests the mesh like bp_create and keeps the index tables of
 * the tiled cell kernel on the host, so the CPU test suite can run them through bp_debug_tiles_emulate.
 * Nothing else works on such a context: every compute entry point needs the GPU.                        */
int bp_debug_create_host(bp_ctx** out, const bp_mesh* mesh, const bp_params* params)
{
	if (!out || !mesh) return bp_fail(nullptr, BP_ERR_ARG, "bp_debug_create_host: NULL argument");
	*out = nullptr;
	bp_ctx* c = new bp_ctx();
	c->host_only = true;
	bp_env_read(&c->env);
	int rc = check_params(c, params);
	if (!rc) { store_params(c, params); rc = bp_setup_topology(c, mesh); }
	if (rc) { g_bp_create_error = c->err; bp_destroy(c); return rc; }
	*out = c;
	return BP_OK;
}

int bp_debug_bsr_pattern(const bp_ctx* c, int64_t* n_rows, const int32_t** rowptr, const int32_t** col)
{
	if (!c) return BP_ERR_ARG;
	if (n_rows) *n_rows = c->nn_owned;
	if (rowptr) *rowptr = c->h_rowptr.data();
	if (col) *col = c->h_col.data();
	return BP_OK;
}

const char* bp_debug_tiles_why(const bp_ctx* c) { return c ? c->tiles_why.c_str() : ""; }

int bp_set_params(bp_ctx* c, const bp_params* p)
{
	if (!c) return BP_ERR_ARG;
	int rc = check_params(c, p);
	if (rc) return rc;
	if ((p->periodic != c->prm.periodic) || (p->use_pressure_bc != c->prm.use_pressure_bc))
		return bp_fail(c, BP_ERR_ARG, "bp_set_params: periodic / use_pressure_bc select the facet sets at bp_create and cannot change");
	if (p->assembly == BP_ASM_TILED && !c->tiles)
		return bp_fail(c, BP_ERR_ARG, "bp_set_params: assembly = BP_ASM_TILED but the tiled cell kernel cannot serve this mesh: %s", c->tiles_why.c_str());
	store_params(c, p);
	c->ainv_dirty = true;
	cudaSetDevice(c->device);
	bp_tiles_params_changed(c);
	bp_upload_quadrature(c);
	return BP_OK;
}

int bp_set_stream(bp_ctx* c, void* s)
{
	if (!c) return BP_ERR_ARG;
	if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
	c->stream = (cudaStream_t)s;
	c->own_stream = false;
	if (c->mg) bp_mg_free(c);          // graphs captured on the old stream
	return BP_OK;
}

int bp_set_field(bp_ctx* c, int field, const double* values, int64_t nvals, int on_device)
{
	if (!c || !values) return BP_ERR_ARG;
	if (nvals != c->nv) return bp_fail(c, BP_ERR_ARG, "bp_set_field: expected %lld vertex values, got %lld", (long long)c->nv, (long long)nvals);
	BP_CUDA(c, cudaSetDevice(c->device));
	const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
	switch (field) {
	case BP_FIELD_S:
		BP_CUDA(c, cudaMemcpyAsync(c->d_io, values, nvals * sizeof(double), kind, c->stream));
		bp_launch_set_S(c, c->d_io);
		c->have_S = true; break;
	case BP_FIELD_A:
		BP_CUDA(c, cudaMemcpyAsync(c->d_A, values, nvals * sizeof(double), kind, c->stream));
		c->have_A = true; c->ainv_dirty = true; break;
	case BP_FIELD_TP:
		BP_CUDA(c, cudaMemcpyAsync(c->d_Tp, values, nvals * sizeof(double), kind, c->stream));
		c->have_Tp = true; c->ainv_dirty = true; break;
	case BP_FIELD_W:
		BP_CUDA(c, cudaMemcpyAsync(c->d_W, values, nvals * sizeof(double), kind, c->stream));
		c->ainv_dirty = true; break;
	case BP_FIELD_E:
		BP_CUDA(c, cudaMemcpyAsync(c->d_E, values, nvals * sizeof(double), kind, c->stream));
		c->ainv_dirty = true; break;
	case BP_FIELD_BETA:
		BP_CUDA(c, cudaMemcpyAsync(c->d_beta, values, nvals * sizeof(double), kind, c->stream));
		c->have_beta = true; break;
	case BP_FIELD_B:
		BP_CUDA(c, cudaMemcpyAsync(c->d_Bv, values, nvals * sizeof(double), kind, c->stream));
		c->have_B = true; break;
	case BP_FIELD_B_RING:
		BP_CUDA(c, cudaMemcpyAsync(c->d_Bring, values, nvals * sizeof(double), kind, c->stream));
		break;
	case BP_FIELD_UZ:
		BP_CUDA(c, cudaMemcpyAsync(c->d_uz, values, nvals * sizeof(double), kind, c->stream));
		break;
	default:
		return bp_fail(c, BP_ERR_ARG, "bp_set_field: unknown field %d", field);
	}
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	c->have_J = false;
	return BP_OK;
}

int bp_get_sizes(const bp_ctx* c, int64_t* n_nodes, int64_t* n_dofs, int64_t* nnz_csr, int64_t* nnz_blocks, int64_t* n_cells, int64_t* n_basal)
{
	if (!c) return BP_ERR_ARG;
	if (n_nodes) *n_nodes = c->nn;
	if (n_dofs) *n_dofs = c->nd_ext;
	if (nnz_csr) *nnz_csr = 4 * c->nnzb;
	if (nnz_blocks) *nnz_blocks = c->nnzb;
	if (n_cells) *n_cells = c->nt;
	if (n_basal) *n_basal = c->n_basal;
	return BP_OK;
}

int bp_csr_pattern(bp_ctx* c, int64_t* nnz, const int64_t** indptr, const int32_t** indices)
{
	if (!c) return BP_ERR_ARG;
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc = bp_build_csr_export(c);
	if (rc) return rc;
	if (nnz) *nnz = c->csr_indptr.back();
	if (indptr) *indptr = c->csr_indptr.data();
	if (indices) *indices = c->csr_indices.data();
	return BP_OK;
}

// Host-buffer assembly, pipelined: u travels to the device in pieces, each slice of the row-gather pass
// starts as soon as the node prefix its incidences reference has arrived (h_blk_need_node, a prefix
// maximum computed at set-up: with any locality in the numbering the first slices need only the first
// part of u), and every slice of F starts its way back as soon as it is final -- right after the slice for
// rows no facet integral touches, after the facet kernel for the others.  Three streams: copies in,
// kernels, copies out (PCIe is full duplex).  Needs the caller's numbering to be the internal one.
static int assemble_pipelined(bp_ctx* c, int what, const double* u, double* F)
{
	const int64_t no = c->nn_owned, nn = c->nn;
	const int nblk = (int)((no + BP_ROWS_PER_CTA - 1) / BP_ROWS_PER_CTA);
	const int K = std::max(1, std::min(c->env.pipe_k, std::min(8, nblk / 64)));
	if (!c->s_h2d) {
		BP_CUDA(c, cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
		BP_CUDA(c, cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
		c->pipe_ev.resize(2 * 8 + 2);
		for (auto& e : c->pipe_ev) BP_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	}
	cudaEvent_t* e_h = c->pipe_ev.data();
	cudaEvent_t* e_c = c->pipe_ev.data() + 8;
	cudaEvent_t e_f = c->pipe_ev[16], e_0 = c->pipe_ev[17];
	bp_update_ainv(c);
	// the copy streams start after whatever the context's stream still has in flight
	BP_CUDA(c, cudaEventRecord(e_0, c->stream));
	BP_CUDA(c, cudaStreamWaitEvent(c->s_h2d, e_0, 0));
	BP_CUDA(c, cudaStreamWaitEvent(c->s_d2h, e_0, 0));
	int64_t prev = 0;
	for (int k = 0; k < K; ++k) {
		const int b1 = (int)((int64_t)(k + 1) * nblk / K);
		const int64_t hi = (k == K - 1) ? nn : std::min<int64_t>(nn, c->h_blk_need_node[b1 - 1]);
		if (hi > prev) {
			BP_CUDA(c, cudaMemcpyAsync(c->d_u + 2 * prev, u + 2 * prev, (size_t)(hi - prev) * 2 * sizeof(double), cudaMemcpyHostToDevice, c->s_h2d));
			prev = hi;
		}
		BP_CUDA(c, cudaEventRecord(e_h[k], c->s_h2d));
	}
	int vprev = 0;
	for (int k = 0; k < K; ++k) {
		const int b0 = (int)((int64_t)k * nblk / K), b1 = (int)((int64_t)(k + 1) * nblk / K);
		const int vhi = (k == K - 1) ? (int)c->nv : std::max(vprev, (int)c->h_blk_need_vert[b1 - 1]);
		BP_CUDA(c, cudaStreamWaitEvent(c->stream, e_h[k], 0));
		bp_launch_rows_chunk(c, what, b0, b1 - b0, vprev, vhi);
		vprev = vhi;
		BP_CUDA(c, cudaEventRecord(e_c[k], c->stream));
	}
	bp_launch_facets(c, what);
	BP_CUDA(c, cudaEventRecord(e_f, c->stream));
	if (what & BP_ASSEMBLE_J) c->have_J = true;
	if (F && (what & BP_ASSEMBLE_F)) {
		for (int pass = 0; pass < 2; ++pass)           // slices no facet touches first, the others after the facet kernel
			for (int k = 0; k < K; ++k) {
				const int b0 = (int)((int64_t)k * nblk / K), b1 = (int)((int64_t)(k + 1) * nblk / K);
				bool touched = false;
				for (int b = b0; b < b1 && !touched; ++b) touched = c->h_blk_facet[b] != 0;
				if ((int)touched != pass) continue;
				const int64_t r0 = (int64_t)b0 * BP_ROWS_PER_CTA, r1 = std::min<int64_t>(no, (int64_t)b1 * BP_ROWS_PER_CTA);
				BP_CUDA(c, cudaStreamWaitEvent(c->s_d2h, touched ? e_f : e_c[k], 0));
				BP_CUDA(c, cudaMemcpyAsync(F + 2 * r0, c->d_F + 2 * r0, (size_t)(r1 - r0) * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->s_d2h));
			}
		BP_CUDA(c, cudaStreamSynchronize(c->s_d2h));
	}
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	BP_CUDA(c, cudaGetLastError());
	return BP_OK;
}

int bp_assemble(bp_ctx* c, int what, const double* u, double* F, double* Jv, int on_device)
{
	if (!c || !u) return BP_ERR_ARG;
	if (!(what & (BP_ASSEMBLE_F | BP_ASSEMBLE_J))) return bp_fail(c, BP_ERR_ARG, "bp_assemble: nothing to assemble");
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc = ready(c);
	if (rc) return rc;
	if (Jv && (rc = bp_build_csr_export(c))) return rc;
	if (!on_device && !Jv && bp_tiles_pipeline_ready(c, what))
		return bp_tiles_assemble_pipelined(c, what, u, F);
	if (!on_device && !Jv && !bp_tiles_serve(c, what) && c->identity_dofs && !c->comm && c->nn == c->nn_owned && c->nd_ext == 2 * c->nn && bp_rows_gather_mode(c)
	    && !(what & ~(BP_ASSEMBLE_F | BP_ASSEMBLE_J)) && !c->env.no_pipeline)
		return assemble_pipelined(c, what, u, F);
	bp_ctx* cs[1] = { c };
	if (on_device && !Jv && c->identity_dofs && c->nn == c->nn_owned && c->nd_ext == 2 * c->nn && !(what & ~(BP_ASSEMBLE_F | BP_ASSEMBLE_J)) &&
	    bp_tiles_serve(c, what) && F) {
		// device buffers in the internal numbering: the tiled kernel reads the caller's u and writes the caller's F in place
		// (no copy into / out of the context's vectors: two launches and 64 MB of traffic per pass at 5 M tets)
		double *su = c->d_u, *sf = c->d_F;
		c->d_u = const_cast<double*>(u); c->d_F = F;
		rc = group_assemble(cs, 1, what);
		c->d_u = su; c->d_F = sf;
		if (rc) return rc;
		BP_CUDA(c, cudaStreamSynchronize(c->stream));
		BP_CUDA(c, cudaGetLastError());
		return BP_OK;
	}
	if ((rc = vec_in(c, u, on_device, c->d_u))) return rc;
	if ((rc = group_assemble(cs, 1, what))) return rc;
	if (F && (what & BP_ASSEMBLE_F)) { if ((rc = vec_out(c, c->d_F, F, on_device))) return rc; }
	if (Jv && (what & BP_ASSEMBLE_J)) {
		const int64_t nnz = c->csr_indptr.back();
		if (on_device) bp_launch_csr_export(c, Jv);
		else {
			// the export buffer is allocated once and kept (a caller who asks for J on the host does so every Newton step)
			if (!c->d_csr_out) BP_CUDA(c, cudaMalloc((void**)&c->d_csr_out, std::max<int64_t>(nnz, 1) * sizeof(double)));
			bp_launch_csr_export(c, c->d_csr_out);
			BP_CUDA(c, cudaMemcpyAsync(Jv, c->d_csr_out, nnz * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
			BP_CUDA(c, cudaStreamSynchronize(c->stream));
		}
	}
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	BP_CUDA(c, cudaGetLastError());
	return BP_OK;
}

int bp_spmv(bp_ctx* c, const double* x, double* y, int on_device)
{
	if (!c || !x || !y) return BP_ERR_ARG;
	if (!c->have_J) return bp_fail(c, BP_ERR_STATE, "bp_spmv: no resident Jacobian (bp_assemble with BP_ASSEMBLE_J first)");
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc;
	if ((rc = vec_in(c, x, on_device, c->d_p))) return rc;
	bp_ctx* cs[1] = { c };
	if ((rc = bp_comm_halo(cs, 1, 1))) return rc;
	bp_launch_spmv(c, c->d_p, c->d_Ap, 0);
	if ((rc = vec_out(c, c->d_Ap, y, on_device))) return rc;
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	BP_CUDA(c, cudaGetLastError());
	return BP_OK;
}

int bp_set_dirichlet(bp_ctx* c, int64_t n, const int32_t* dofs, const double* values)
{
	if (!c || n < 0 || (n > 0 && (!dofs || !values))) return BP_ERR_ARG;
	BP_CUDA(c, cudaSetDevice(c->device));
	if (n == 0) { c->have_bc = false; return BP_OK; }
	if (c->prm.reduced_stokes) return bp_fail(c, BP_ERR_ARG, "bp_set_dirichlet: not available for the reduced-Stokes solve");
	std::vector<double> m((size_t)c->nd_ext, 0.0), g((size_t)c->nd_ext, 0.0);
	for (int64_t k = 0; k < n; ++k) {
		if (dofs[k] < 0 || dofs[k] >= c->nd_ext) return bp_fail(c, BP_ERR_ARG, "bp_set_dirichlet: dof %d outside [0, %lld)", dofs[k], (long long)c->nd_ext);
		m[dofs[k]] = 1.0; g[dofs[k]] = values[k];
	}
	const size_t nb = (size_t)2 * c->nn * sizeof(double);
	if (!c->d_bcmask) { BP_CUDA(c, cudaMalloc((void**)&c->d_bcmask, nb)); BP_CUDA(c, cudaMalloc((void**)&c->d_bcval, nb)); }
	BP_CUDA(c, cudaMemcpyAsync(c->d_io, m.data(), m.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	bp_launch_gather_in(c, c->d_io, c->d_bcmask);
	BP_CUDA(c, cudaMemcpyAsync(c->d_io2, g.data(), g.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	bp_launch_gather_in(c, c->d_io2, c->d_bcval);
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	c->have_bc = true;
	return BP_OK;
}

int bp_set_dirichlet_uz(bp_ctx* c, int64_t n, const int32_t* vertices, const double* values)
{
	if (!c || n < 0 || (n > 0 && (!vertices || !values))) return BP_ERR_ARG;
	BP_CUDA(c, cudaSetDevice(c->device));
	std::vector<uint8_t> bc = c->h_bcnode;
	std::vector<int32_t> nodes;
	std::vector<double> vals;
	for (int64_t k = 0; k < n; ++k) {
		if (vertices[k] < 0 || vertices[k] >= c->nv) return bp_fail(c, BP_ERR_ARG, "bp_set_dirichlet_uz: vertex %d outside [0, %lld)", vertices[k], (long long)c->nv);
		const int32_t nd_ = c->h_v2n.empty() ? vertices[k] : c->h_v2n[vertices[k]];
		if (nd_ >= c->nn_owned) continue;                  // a ghost: its owner constrains it
		if (!(bc[nd_] & 2)) { bc[nd_] |= 2; nodes.push_back(nd_); vals.push_back(values[k]); }
	}
	if (c->d_uzbc_node) { cudaFree(c->d_uzbc_node); cudaFree(c->d_uzbc_val); c->d_uzbc_node = nullptr; c->d_uzbc_val = nullptr; }
	c->n_uzbc = (int64_t)nodes.size();
	if (c->n_uzbc) {
		BP_CUDA(c, cudaMalloc((void**)&c->d_uzbc_node, nodes.size() * sizeof(int32_t)));
		BP_CUDA(c, cudaMalloc((void**)&c->d_uzbc_val, vals.size() * sizeof(double)));
		BP_CUDA(c, cudaMemcpy(c->d_uzbc_node, nodes.data(), nodes.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
		BP_CUDA(c, cudaMemcpy(c->d_uzbc_val, vals.data(), vals.size() * sizeof(double), cudaMemcpyHostToDevice));
	}
	if (!bc.empty()) BP_CUDA(c, cudaMemcpy(c->d_bcnode, bc.data(), bc.size(), cudaMemcpyHostToDevice));
	return BP_OK;
}

int bp_krylov_solve(bp_ctx* c, const double* b, double* dx, int on_device, int32_t* iterations, double* rel_residual)
{
	if (!c || !b || !dx) return BP_ERR_ARG;
	if (!c->have_J) return bp_fail(c, BP_ERR_STATE, "bp_krylov_solve: no resident Jacobian");
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc;
	if ((rc = vec_in(c, b, on_device, c->d_F))) return rc;
	bp_ctx* cs[1] = { c };
	int it = 0;
	if ((rc = group_pcg(cs, 1, &it, rel_residual))) return rc;
	if (iterations) *iterations = it;
	if ((rc = vec_out(c, c->d_dx, dx, on_device))) return rc;
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	BP_CUDA(c, cudaGetLastError());
	return BP_OK;
}

int bp_newton_solve(bp_ctx* c, double* u, int on_device, bp_stats* st)
{
	if (!c || !u) return BP_ERR_ARG;
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc;
	if ((rc = vec_in(c, u, on_device, c->d_u))) return rc;
	bp_ctx* cs[1] = { c };
	rc = group_newton(cs, 1, st);
	if (rc && rc != BP_ERR_NONCONV) return rc;
	int rc2 = vec_out(c, c->d_u, u, on_device);
	if (rc2) return rc2;
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	BP_CUDA(c, cudaGetLastError());
	return rc;
}

int bp_newton_solve_group(bp_ctx** cs, int n, double** u, int on_device, bp_stats* st)
{
	if (!cs || n < 1 || n > 16 || !u) return BP_ERR_ARG;
	for (int r = 0; r < n; ++r) if (!cs[r] || cs[r]->device != cs[0]->device)
		return bp_fail(cs[0], BP_ERR_ARG, "bp_newton_solve_group: contexts must live on one device");
	BP_CUDA(cs[0], cudaSetDevice(cs[0]->device));
	GroupStream gs(cs, n);
	int rc;
	for (int r = 0; r < n; ++r) if ((rc = vec_in(cs[r], u[r], on_device, cs[r]->d_u))) return rc;
	rc = group_newton(cs, n, st);
	if (rc && rc != BP_ERR_NONCONV) return rc;
	for (int r = 0; r < n; ++r) { int rc2 = vec_out(cs[r], cs[r]->d_u, u[r], on_device); if (rc2) return rc2; }
	BP_CUDA(cs[0], cudaStreamSynchronize(cs[0]->stream));
	BP_CUDA(cs[0], cudaGetLastError());
	return rc;
}

int bp_assemble_group(bp_ctx** cs, int n, int what, const double** u, double** F, int on_device)
{
	if (!cs || n < 1 || n > 16 || !u) return BP_ERR_ARG;
	for (int r = 0; r < n; ++r) if (!cs[r] || cs[r]->device != cs[0]->device)
		return bp_fail(cs[0], BP_ERR_ARG, "bp_assemble_group: contexts must live on one device");
	BP_CUDA(cs[0], cudaSetDevice(cs[0]->device));
	GroupStream gs(cs, n);
	int rc;
	for (int r = 0; r < n; ++r) { if ((rc = ready(cs[r]))) return rc; if ((rc = vec_in(cs[r], u[r], on_device, cs[r]->d_u))) return rc; }
	if ((rc = group_assemble(cs, n, what))) return rc;
	if (F) for (int r = 0; r < n; ++r) if (F[r] && (rc = vec_out(cs[r], cs[r]->d_F, F[r], on_device))) return rc;
	BP_CUDA(cs[0], cudaStreamSynchronize(cs[0]->stream));
	BP_CUDA(cs[0], cudaGetLastError());
	return BP_OK;
}

int bp_linear_solve(bp_ctx* c, const double* u_lin, double* u_out, int on_device, int32_t* iterations)
{
	if (!c || !u_lin || !u_out) return BP_ERR_ARG;
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc = ready(c);
	if (rc) return rc;
	if ((rc = vec_in(c, u_lin, on_device, c->d_u))) return rc;
	bp_ctx* cs[1] = { c };
	if ((rc = group_assemble(cs, 1, BP_ASSEMBLE_F | BP_ASSEMBLE_J | 16))) return rc;     // K(u_lin), f
	if (c->have_bc) { bp_launch_bc_rows(c, 0.0); bp_launch_bc_eliminate(c); }             // the solve returns -u: d = -g
	int it = 0;
	if ((rc = group_pcg(cs, 1, &it, nullptr))) return rc;                                  // dx = K^-1 f
	if (iterations) *iterations = it;
	BP_CUDA(c, cudaMemsetAsync(c->d_u, 0, (size_t)2 * c->nn * sizeof(double), c->stream));
	bp_launch_axpy(c, -1.0, c->d_dx, c->d_u);                                              // u = -dx
	if ((rc = vec_out(c, c->d_u, u_out, on_device))) return rc;
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	BP_CUDA(c, cudaGetLastError());
	return BP_OK;
}

static int solve_pressure_impl(bp_ctx* c, const double* u, const double* u_model, double* p_vertex, int on_device, int32_t* iterations)
{
	if (!c || !u || !p_vertex) return BP_ERR_ARG;
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc = ready(c);
	if (rc) return rc;
	if (c->prm.assembly == BP_ASM_SCATTER || (size_t)4 * c->max_row * 128 * sizeof(double) > 200 * 1024)
		return bp_fail(c, BP_ERR_STATE, "bp_solve_pressure needs the row-gather assembly path");
	if (c->prm.energy_dependent)
		return bp_fail(c, BP_ERR_STATE, "bp_solve_pressure with the energy-dependent rate factor is not available yet");
	bp_ctx* cs[1] = { c };
	if (u_model) {                        // viscosity state of the reduced-Stokes solve(): per-vertex copy in d_uvert2
		if (c->nn != c->nn_owned && !c->comm) return bp_fail(c, BP_ERR_STATE, "bp_rs_solve_pressure on a partitioned context needs its communicator");
		if ((rc = vec_in(c, u_model, on_device, c->d_u))) return rc;
		if ((rc = bp_comm_halo(cs, 1, 0))) return rc;
		bp_launch_expand_to_vertex(c, c->d_u, c->d_uvert2);
	}
	if ((rc = vec_in(c, u, on_device, c->d_u))) return rc;
	if ((rc = bp_comm_halo(cs, 1, 0))) return rc;
	bp_launch_rows_aux(c, 0, u_model ? (const double2*)c->d_uvert2 : nullptr, c->prm.reduced_stokes ? c->d_uz : nullptr);
	c->have_J = false;
	// mass-matrix solve: block-Jacobi PCG to 1e-13 (dolfin's project() uses a direct solver)
	const bp_params saved = c->prm;
	c->prm.pc = BP_PC_BLOCK_JACOBI; c->prm.krylov_rtol = 1e-13; c->prm.krylov_atol = 0.0; c->prm.krylov_maxit = 2000;
	c->prm.krylov_error_on_nonconvergence = 1;       // dolfin's project() is a direct solve: an unconverged p is an error
	int it = 0;
	rc = group_pcg(cs, 1, &it, nullptr);
	c->prm = saved;
	if (rc) return rc;
	if (iterations) *iterations = it;
	if ((rc = bp_comm_halo(cs, 1, 5))) return rc;                 // ghost nodes get their owner's value
	double* out = on_device ? p_vertex : c->d_io;
	bp_launch_node_to_vertex(c, c->d_dx, 0, out);
	if (!on_device) BP_CUDA(c, cudaMemcpyAsync(p_vertex, c->d_io, (size_t)c->nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	BP_CUDA(c, cudaGetLastError());
	return BP_OK;
}

int bp_solve_pressure(bp_ctx* c, const double* u, double* p_vertex, int on_device, int32_t* iterations)
{
	return solve_pressure_impl(c, u, nullptr, p_vertex, on_device, iterations);
}

int bp_rs_solve_pressure(bp_ctx* c, const double* u, const double* u_model, double* p_vertex, int on_device, int32_t* iterations)
{
	if (!c || !u_model) return BP_ERR_ARG;
	if (!c->prm.reduced_stokes) return bp_fail(c, BP_ERR_STATE, "bp_rs_solve_pressure: bp_params.reduced_stokes is not set");
	return solve_pressure_impl(c, u, u_model, p_vertex, on_device, iterations);
}

// u_z from d_u (cslvr/momentumbp.py:231-253); the result is left per node in d_dx[2*i] (ghosts filled)
static int vert_velocity_core(bp_ctx* c, int32_t* iterations)
{
	int rc;
	if (!c->have_B) return bp_fail(c, BP_ERR_STATE, "bp_solve_vert_velocity: BP_FIELD_B (bed height) not set");
	if (c->max_layers > 64) return bp_fail(c, BP_ERR_ARG, "bp_solve_vert_velocity supports at most 64 nodes per ice column");
	if (c->prm.assembly == BP_ASM_SCATTER || (size_t)4 * c->max_row * 128 * sizeof(double) > 200 * 1024)
		return bp_fail(c, BP_ERR_STATE, "bp_solve_vert_velocity needs the row-gather assembly path");
	bp_ctx* cs[1] = { c };
	if ((rc = bp_comm_halo(cs, 1, 0))) return rc;
	// 1. u_z_b = project(kinematic basal condition, Q): mass-matrix PCG
	bp_launch_rows_aux(c, 1);
	c->have_J = false;
	{
		const bp_params saved = c->prm;
		c->prm.pc = BP_PC_BLOCK_JACOBI; c->prm.krylov_rtol = 1e-13; c->prm.krylov_atol = 0.0; c->prm.krylov_maxit = 2000;
		c->prm.krylov_error_on_nonconvergence = 1;   // project() is a direct solve in the reference
		rc = group_pcg(cs, 1, nullptr, nullptr);
		c->prm = saved;
		if (rc) return rc;
	}
	BP_CUDA(c, cudaMemcpyAsync(c->d_aux, c->d_dx, (size_t)2 * c->nn * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
	// 2. the continuity system with Dirichlet rows -> d_val (component 0 of the blocks), d_F
	bp_launch_uz_lateral_values(c, c->d_aux);      // DirichletBC(Q, u_z_lat, ff, GAMMA_L_DVD): applied last, it wins on shared nodes
	bp_launch_rows_aux(c, 2);
	// 3. BiCGStab, right-preconditioned by the pivoted tridiagonal solve of every column
	const size_t nb = (size_t)2 * c->nn * sizeof(double);
	double *x = c->d_dx, *r = c->d_r, *rh = c->d_z2, *p = c->d_p, *v = c->d_Ap, *y = c->d_z, *sv = c->d_cd, *t = c->d_io, *zz = c->d_io2;
	auto dot = [&](const double* a, const double* b, double* out) -> int {
		bp_launch_dot_owned(c, a, b, SC_TMP);
		int r_ = bp_comm_allreduce(cs, 1, SC_TMP, 1);
		if (!r_) r_ = read_scalars(c);
		*out = c->h_sc[SC_TMP];
		return r_;
	};
	BP_CUDA(c, cudaMemsetAsync(x, 0, nb, c->stream));
	BP_CUDA(c, cudaMemsetAsync(p, 0, nb, c->stream));
	BP_CUDA(c, cudaMemsetAsync(v, 0, nb, c->stream));
	BP_CUDA(c, cudaMemcpyAsync(r, c->d_F, nb, cudaMemcpyDeviceToDevice, c->stream));
	BP_CUDA(c, cudaMemcpyAsync(rh, c->d_F, nb, cudaMemcpyDeviceToDevice, c->stream));
	double bnorm2, rho = 1.0, alpha = 1.0, omega = 1.0, rnorm2;
	if ((rc = dot(r, r, &bnorm2))) return rc;
	rnorm2 = bnorm2;
	int it = 0;
	const int maxit = 500;
	while (it < maxit && bnorm2 > 0.0 && rnorm2 > 1e-24 * bnorm2) {
		double rho_n, rhv, ts, tt;
		if ((rc = dot(rh, r, &rho_n))) return rc;
		if (rho_n == 0.0 || !std::isfinite(rho_n)) return bp_fail(c, BP_ERR_STATE, "BiCGStab breakdown (rho) in the u_z solve at iteration %d", it);
		const double beta = (rho_n / rho) * (alpha / omega);
		bp_launch_lincomb(c, p, 1.0, r, beta, p, -beta * omega, v);          // p = r + beta (p - omega v)
		bp_launch_col_gtsv(c, p, y);                                          // y = M^-1 p
		if ((rc = bp_comm_halo(cs, 1, 2))) return rc;                         // ghosts of y (= d_z)
		bp_launch_spmv(c, y, v, 0);                                           // v = A y
		if ((rc = dot(rh, v, &rhv))) return rc;
		if (rhv == 0.0 || !std::isfinite(rhv)) return bp_fail(c, BP_ERR_STATE, "BiCGStab breakdown (alpha) in the u_z solve at iteration %d", it);
		alpha = rho_n / rhv;
		bp_launch_lincomb(c, sv, 1.0, r, -alpha, v, 0.0, v);                  // s = r - alpha v
		bp_launch_col_gtsv(c, sv, zz);                                        // z = M^-1 s
		if ((rc = bp_comm_halo(cs, 1, 4))) return rc;                         // ghosts of z (= d_io2)
		bp_launch_spmv(c, zz, t, 0);                                          // t = A z
		if ((rc = dot(t, sv, &ts))) return rc;
		if ((rc = dot(t, t, &tt))) return rc;
		omega = (tt > 0.0) ? ts / tt : 0.0;
		bp_launch_lincomb(c, x, 1.0, x, alpha, y, omega, zz);                 // x += alpha y + omega z
		bp_launch_lincomb(c, r, 1.0, sv, -omega, t, 0.0, t);                  // r = s - omega t
		if ((rc = dot(r, r, &rnorm2))) return rc;
		rho = rho_n;
		++it;
		if (omega == 0.0) break;
	}
	if (bnorm2 > 0.0 && !(rnorm2 <= 1e-20 * bnorm2))
		return bp_fail(c, BP_ERR_NONCONV, "u_z solve: BiCGStab stopped at relative residual %.3e after %d iterations", std::sqrt(rnorm2 / bnorm2), it);
	if (iterations) *iterations = it;
	if ((rc = bp_comm_halo(cs, 1, 5))) return rc;                 // x is d_dx: ghost nodes get their owner's value
	return BP_OK;
}

int bp_solve_vert_velocity(bp_ctx* c, const double* u, double* uz_vertex, int on_device, int32_t* iterations)
{
	if (!c || !u || !uz_vertex) return BP_ERR_ARG;
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc = ready(c);
	if (rc) return rc;
	if ((rc = vec_in(c, u, on_device, c->d_u))) return rc;
	if ((rc = vert_velocity_core(c, iterations))) return rc;
	double* out = on_device ? uz_vertex : c->d_io;
	bp_launch_node_to_vertex(c, c->d_dx, 0, out);
	if (!on_device) BP_CUDA(c, cudaMemcpyAsync(uz_vertex, c->d_io, (size_t)c->nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	BP_CUDA(c, cudaGetLastError());
	return BP_OK;
}

// ---- MomentumDukowiczStokesReduced.solve: Model.home_rolled_newton_method (cslvr/model.py:2676-2729)
int bp_rs_newton_solve(bp_ctx* c, double* u, double* uz_vertex, int on_device, double atol, int run_callback, bp_stats* st)
{
	if (!c || !u) return BP_ERR_ARG;
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc = ready(c);
	if (rc) return rc;
	if (!c->prm.reduced_stokes) return bp_fail(c, BP_ERR_STATE, "bp_rs_newton_solve: bp_params.reduced_stokes is not set");
	if (c->nn != c->nn_owned && !c->comm) return bp_fail(c, BP_ERR_STATE, "bp_rs_newton_solve on a partitioned context needs its communicator (one rank per process; in-process groups are not served)");
	if (run_callback && !c->have_B) return bp_fail(c, BP_ERR_STATE, "bp_rs_newton_solve: BP_FIELD_B (bed height) not set");
	const bp_params& P = c->prm;
	bp_stats s;
	memset(&s, 0, sizeof(s));
	const int64_t l0 = c->launches;
	const int unconv0 = c->krylov_unconverged;
	if (uz_vertex)
		BP_CUDA(c, cudaMemcpyAsync(c->d_uz, uz_vertex, (size_t)c->nv * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
	if ((rc = vec_in(c, u, on_device, c->d_u))) return rc;
	bp_ctx* cs[1] = { c };
	cudaEvent_t eA, eB, eT0, eT1;
	cudaEventCreate(&eA); cudaEventCreate(&eB); cudaEventCreate(&eT0); cudaEventCreate(&eT1);
	float ms;
	cudaEventRecord(eT0, c->stream);
	double rn = 0.0, r0 = 0.0;
	while (!s.converged && s.newton_iterations < P.newton_maxit) {
		// A, b = assemble_system(J, -R): the residual and the Jacobian at the current (u, u_z)
		cudaEventRecord(eA, c->stream);
		rc = group_assemble(cs, 1, BP_ASSEMBLE_F | BP_ASSEMBLE_J);
		cudaEventRecord(eB, c->stream);
		if (rc) break;
		cudaEventSynchronize(eB); cudaEventElapsedTime(&ms, eA, eB); s.t_assemble_ms += ms; s.n_assemblies++;
		if ((rc = group_norm_F(cs, 1, &rn))) break;
		if (!std::isfinite(rn)) { rc = bp_fail(c, BP_ERR_STATE, "residual became non-finite at Newton iteration %d", s.newton_iterations); break; }
		if (s.newton_iterations == 0) { r0 = rn; s.residual_norm0 = rn; }
		if (s.newton_iterations < 64) s.residual_history[s.newton_iterations] = rn;
		// solve(A, d, b): d = -J^-1 R; the step is taken in the converging iteration as well
		int kit = 0;
		if (rn > 0.0) {
			cudaEventRecord(eA, c->stream);
			rc = group_pcg(cs, 1, &kit, nullptr);
			cudaEventRecord(eB, c->stream);
			cudaEventSynchronize(eB); cudaEventElapsedTime(&ms, eA, eB); s.t_krylov_ms += ms;
			if (rc) break;
			s.krylov_iterations += kit;
			bp_launch_axpy(c, -P.newton_relax, c->d_dx, c->d_u);
		}
		s.converged = (rn < atol) || (r0 > 0.0 && rn / r0 < P.newton_rtol);
		s.newton_iterations++;
		if (P.verbose)
			printf("Newton iteration %d: r (abs) = %.3e (tol = %.3e) r (rel) = %.3e (tol = %.3e)  [krylov %d]\n",
			       s.newton_iterations, rn, atol, r0 > 0.0 ? rn / r0 : 0.0, P.newton_rtol, kit);
		if (run_callback) {            // cb_ftn = solve_vert_velocity (momentumstokes.py:468)
			if ((rc = vert_velocity_core(c, nullptr))) break;
			bp_launch_node_to_vertex(c, c->d_dx, 0, c->d_uz);
		}
	}
	s.residual_norm = rn;
	if (!rc && !s.converged && P.error_on_nonconvergence)
		rc = bp_fail(c, BP_ERR_NONCONV, "Newton solver did not converge in %d iterations (|F| = %.3e)", s.newton_iterations, rn);
	cudaEventRecord(eT1, c->stream);
	cudaEventSynchronize(eT1);
	cudaEventElapsedTime(&ms, eT0, eT1);
	s.t_total_ms = ms;
	cudaEventDestroy(eA); cudaEventDestroy(eB); cudaEventDestroy(eT0); cudaEventDestroy(eT1);
	s.gpu_launches = (int32_t)(c->launches - l0);
	s.krylov_unconverged = c->krylov_unconverged - unconv0;
	s.krylov_last_relres = c->krylov_last_relres;
	if (st) *st = s;
	if (rc && rc != BP_ERR_NONCONV) return rc;
	int rc2 = vec_out(c, c->d_u, u, on_device);
	if (rc2) return rc2;
	if (uz_vertex)
		BP_CUDA(c, cudaMemcpyAsync(uz_vertex, c->d_uz, (size_t)c->nv * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	BP_CUDA(c, cudaGetLastError());
	c->have_J = false;
	return rc;
}

// ---- timing on the context's stream with CUDA events -------------------------
static int flush_l2(bp_ctx* c, int enable)
{
	static void* buf = nullptr;
	static size_t bytes = 0;
	if (!enable) return BP_OK;
	if (!buf) { bytes = (size_t)512 << 20; BP_CUDA(c, cudaMalloc(&buf, bytes)); }
	BP_CUDA(c, cudaMemsetAsync(buf, 0, bytes, c->stream));
	return BP_OK;
}

int bp_time_assemble(bp_ctx* c, int what, int repeats, int flush, double* ms_per_launch)
{
	if (!c || repeats < 1 || !ms_per_launch) return BP_ERR_ARG;
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc = ready(c);
	if (rc) return rc;
	double total = 0;
	for (int i = 0; i < repeats; ++i) {
		if ((rc = flush_l2(c, flush))) return rc;
		BP_CUDA(c, cudaEventRecord(c->ev0, c->stream));
		bp_launch_assemble(c, what);
		BP_CUDA(c, cudaEventRecord(c->ev1, c->stream));
		BP_CUDA(c, cudaEventSynchronize(c->ev1));
		float ms; BP_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
		total += ms;
	}
	BP_CUDA(c, cudaGetLastError());
	*ms_per_launch = total / repeats;
	return BP_OK;
}

int bp_time_spmv(bp_ctx* c, int repeats, int flush, double* ms_per_launch)
{
	if (!c || repeats < 1 || !ms_per_launch) return BP_ERR_ARG;
	if (!c->have_J) return bp_fail(c, BP_ERR_STATE, "bp_time_spmv: no resident Jacobian");
	BP_CUDA(c, cudaSetDevice(c->device));
	int rc;
	double total = 0;
	for (int i = 0; i < repeats; ++i) {
		if ((rc = flush_l2(c, flush & 1))) return rc;
		BP_CUDA(c, cudaEventRecord(c->ev0, c->stream));
		bp_launch_spmv(c, c->d_u, c->d_Ap, (flush >> 1) & 1);
		BP_CUDA(c, cudaEventRecord(c->ev1, c->stream));
		BP_CUDA(c, cudaEventSynchronize(c->ev1));
		float ms; BP_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
		total += ms;
	}
	BP_CUDA(c, cudaGetLastError());
	*ms_per_launch = total / repeats;
	return BP_OK;
}

int64_t bp_launch_count(const bp_ctx* c) { return c ? c->launches : 0; }

}  // extern "C"
