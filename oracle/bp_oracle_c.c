/* bp_oracle_c.c -- C/OpenMP twin of oracle/bp_oracle.py (TEST INFRASTRUCTURE).
 *
 * PARITY UNPINNED (see oracle/__init__.py).  Same restatement of the reference as the
 * numpy oracle, written the way the reference's CPU path does the work so that it
 * can double as the reported CPU baseline ("port"):
 *   - tabulate_tensor: per cell, a loop over the quadrature points with tabulated P1
 *     basis values; cell-constant sub-expressions hoisted (cslvr/momentumbp.py:481-500
 *     -> UFL derivative -> FFC-generated code, third party);
 *   - Assembler: element tensors added into a precomputed CSR by binary search in the
 *     row, like MatSetValues on a preallocated AIJ matrix;
 *   - one thread per contiguous cell range (dolfin would run one MPI rank per range).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * may load this; the product never does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
	double n, eps_reg, rho_i, rho_sw, g;
	int periodic, use_pressure_bc;
} bpo_params;

static inline int64_t find_col(const int64_t* indptr, const int32_t* indices, int32_t row, int32_t col)
{
	int64_t lo = indptr[row], hi = indptr[row + 1] - 1;
	while (lo < hi) {
		int64_t mid = (lo + hi) >> 1;
		if (indices[mid] < col) lo = mid + 1; else hi = mid;
	}
	return lo;
}

static inline void atomic_add(double* p, double v)
{
	#pragma omp atomic
	*p += v;
}

/* torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the CPU baseline asks for its
 * thread count explicitly (bench.py) instead of inheriting that */
void bpo_set_num_threads(int n)
{
#ifdef _OPENMP
	if (n > 0) omp_set_num_threads(n);
#else
	(void)n;
#endif
}

int bpo_num_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* cell integrals: (V_d - P_e) dOmega, cslvr/momentumbp.py:471-481; V_d from
 * cslvr/momentum.py:218-221, strain rate from cslvr/momentumbp.py:122-135.          */
void bpo_assemble_cells(int64_t nt, const double* coords, const int32_t* cells, const int32_t* cell_dofs,
                        const double* S, const double* A, const double* u,
                        int nq, const double* qpts, const double* qw, const bpo_params* P,
                        const int64_t* indptr, const int32_t* indices, double* F, double* Jv)
{
	const double n = P->n, pw = (1.0 - n) / (2.0 * n), rho_g = P->rho_i * P->g;
	#pragma omp parallel for schedule(static)
	for (int64_t t = 0; t < nt; ++t) {
		const int32_t* cv = cells + 4 * t;
		const int32_t* cd = cell_dofs + 8 * t;
		double x[4][3];
		for (int a = 0; a < 4; ++a) for (int d = 0; d < 3; ++d) x[a][d] = coords[3 * (int64_t)cv[a] + d];
		/* Jacobian of the affine map and its inverse -> grad lambda */
		double Jm[3][3];
		for (int i = 0; i < 3; ++i) for (int d = 0; d < 3; ++d) Jm[i][d] = x[i + 1][d] - x[0][d];
		const double det = Jm[0][0] * (Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1])
		                 - Jm[0][1] * (Jm[1][0] * Jm[2][2] - Jm[1][2] * Jm[2][0])
		                 + Jm[0][2] * (Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0]);
		double G[4][3];
		G[1][0] = (Jm[1][1] * Jm[2][2] - Jm[1][2] * Jm[2][1]) / det;
		G[1][1] = (Jm[1][2] * Jm[2][0] - Jm[1][0] * Jm[2][2]) / det;
		G[1][2] = (Jm[1][0] * Jm[2][1] - Jm[1][1] * Jm[2][0]) / det;
		G[2][0] = (Jm[2][1] * Jm[0][2] - Jm[2][2] * Jm[0][1]) / det;
		G[2][1] = (Jm[2][2] * Jm[0][0] - Jm[2][0] * Jm[0][2]) / det;
		G[2][2] = (Jm[2][0] * Jm[0][1] - Jm[2][1] * Jm[0][0]) / det;
		G[3][0] = (Jm[0][1] * Jm[1][2] - Jm[0][2] * Jm[1][1]) / det;
		G[3][1] = (Jm[0][2] * Jm[1][0] - Jm[0][0] * Jm[1][2]) / det;
		G[3][2] = (Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0]) / det;
		for (int d = 0; d < 3; ++d) G[0][d] = -(G[1][d] + G[2][d] + G[3][d]);
		const double vol = fabs(det) / 6.0;
		/* cell-constant: grad u, grad S, strain rate and its variations */
		double g[2][3] = { { 0, 0, 0 }, { 0, 0, 0 } }, gS[2] = { 0, 0 };
		for (int a = 0; a < 4; ++a) {
			for (int c = 0; c < 2; ++c) for (int d = 0; d < 3; ++d) g[c][d] += u[cd[4 * c + a]] * G[a][d];
			gS[0] += S[cv[a]] * G[a][0]; gS[1] += S[cv[a]] * G[a][1];
		}
		const double e00 = g[0][0], e11 = g[1][1], e01 = 0.5 * (g[0][1] + g[1][0]);
		const double e02 = 0.5 * g[0][2], e12 = 0.5 * g[1][2], e22 = -e00 - e11;
		const double ed = 0.5 * (e00 * e00 + e11 * e11 + e22 * e22) + e01 * e01 + e02 * e02 + e12 * e12 + P->eps_reg;
		double de[8], H[8][8];
		for (int a = 0; a < 4; ++a) {
			de[a]     = (e00 - e22) * G[a][0] + e01 * G[a][1] + e02 * G[a][2];
			de[4 + a] = (e11 - e22) * G[a][1] + e01 * G[a][0] + e12 * G[a][2];
		}
		for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) {
			H[a][b]         = 2.0 * G[a][0] * G[b][0] + 0.5 * G[a][1] * G[b][1] + 0.5 * G[a][2] * G[b][2];
			H[4 + a][4 + b] = 2.0 * G[a][1] * G[b][1] + 0.5 * G[a][0] * G[b][0] + 0.5 * G[a][2] * G[b][2];
			H[a][4 + b]     = G[a][0] * G[b][1] + 0.5 * G[a][1] * G[b][0];
			H[4 + b][a]     = H[a][4 + b];
		}
		const double two_eta_e = pow(ed, pw), two_eta_p_e = pw * pow(ed, pw - 1.0);
		double Fe[8] = { 0 }, Je[8][8];
		memset(Je, 0, sizeof(Je));
		for (int q = 0; q < nq; ++q) {
			const double* lam = qpts + 4 * q;
			const double dx = qw[q] * vol;
			double Aq = 0.0;
			for (int a = 0; a < 4; ++a) Aq += lam[a] * A[cv[a]];
			const double Ainv = pow(Aq, -1.0 / n);
			const double te = Ainv * two_eta_e, tp = Ainv * two_eta_p_e;
			for (int i = 0; i < 8; ++i) {
				Fe[i] += dx * (te * de[i] + rho_g * gS[i / 4] * lam[i % 4]);
				if (Jv) for (int j = 0; j < 8; ++j) Je[i][j] += dx * (te * H[i][j] + tp * de[i] * de[j]);
			}
		}
		for (int i = 0; i < 8; ++i) {
			if (F) atomic_add(&F[cd[i]], Fe[i]);
			if (Jv) for (int j = 0; j < 8; ++j) atomic_add(&Jv[find_col(indptr, indices, cd[i], cd[j])], Je[i][j]);
		}
	}
}

/* exterior facet integrals: -S_l dGamma_b - P_b dGamma_bw [- P_b dGamma_lt],
 * cslvr/momentumbp.py:473-491; marker sets cslvr/d3model.py:576-598.                  */
void bpo_assemble_facets(int64_t nf, const int32_t* fcell, const int32_t* flocal, const int32_t* fmark,
                         const double* coords, const int32_t* cells, const int32_t* cell_dofs,
                         const double* S, const double* beta, const double* u,
                         int nq3, const double* q3pts, const double* q3w,
                         int nq2, const double* q2pts, const double* q2w, const bpo_params* P,
                         const int64_t* indptr, const int32_t* indices, double* F, double* Jv)
{
	static const int LOC[4][3] = { { 1, 2, 3 }, { 0, 2, 3 }, { 0, 1, 3 }, { 0, 1, 2 } };
	const double rho_g = P->rho_i * P->g, rsw_g = P->rho_sw * P->g;
	const int lateral = (!P->periodic) && P->use_pressure_bc;
	#pragma omp parallel for schedule(static)
	for (int64_t f = 0; f < nf; ++f) {
		const int mk = fmark[f];
		const int basal = (mk == 3 || mk == 5), press = (mk == 5) || (lateral && (mk == 4 || mk == 10));
		if (!basal && !press) continue;
		const int32_t c = fcell[f];
		const int* la = LOC[flocal[f]];
		const int32_t* cv = cells + 4 * (int64_t)c;
		const int32_t* cd = cell_dofs + 8 * (int64_t)c;
		double p[3][3], po[3];
		for (int a = 0; a < 3; ++a) for (int d = 0; d < 3; ++d) p[a][d] = coords[3 * (int64_t)cv[la[a]] + d];
		for (int d = 0; d < 3; ++d) po[d] = coords[3 * (int64_t)cv[flocal[f]] + d];
		double e1[3], e2[3], nr[3];
		for (int d = 0; d < 3; ++d) { e1[d] = p[1][d] - p[0][d]; e2[d] = p[2][d] - p[0][d]; }
		nr[0] = e1[1] * e2[2] - e1[2] * e2[1]; nr[1] = e1[2] * e2[0] - e1[0] * e2[2]; nr[2] = e1[0] * e2[1] - e1[1] * e2[0];
		const double ln = sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2]);
		const double area = 0.5 * ln;
		double side = 0.0;
		for (int d = 0; d < 3; ++d) side += nr[d] * (po[d] - p[0][d]);
		const double sg = (side > 0 ? -1.0 : 1.0) / ln;
		for (int d = 0; d < 3; ++d) nr[d] *= sg;
		double Fe[2][3] = { { 0 } }, Ke[3][3] = { { 0 } };
		if (basal) for (int q = 0; q < nq3; ++q) {
			const double* lam = q3pts + 3 * q;
			const double dx = q3w[q] * area;
			double bq = 0, uq[2] = { 0, 0 };
			for (int a = 0; a < 3; ++a) {
				bq += lam[a] * beta[cv[la[a]]];
				uq[0] += lam[a] * u[cd[la[a]]]; uq[1] += lam[a] * u[cd[4 + la[a]]];
			}
			for (int a = 0; a < 3; ++a) {
				Fe[0][a] += dx * bq * uq[0] * lam[a]; Fe[1][a] += dx * bq * uq[1] * lam[a];
				for (int b = 0; b < 3; ++b) Ke[a][b] += dx * bq * lam[a] * lam[b];
			}
		}
		if (press) for (int q = 0; q < nq2; ++q) {
			const double* lam = q2pts + 3 * q;
			const double dx = q2w[q] * area;
			double fq = 0;
			for (int a = 0; a < 3; ++a) {
				const double z = p[a][2], D = -fmin(0.0, z);
				fq += lam[a] * (rho_g * (S[cv[la[a]]] - z) - rsw_g * D);
			}
			for (int a = 0; a < 3; ++a) { Fe[0][a] -= dx * fq * nr[0] * lam[a]; Fe[1][a] -= dx * fq * nr[1] * lam[a]; }
		}
		for (int cc = 0; cc < 2; ++cc) for (int a = 0; a < 3; ++a) {
			const int32_t row = cd[4 * cc + la[a]];
			if (F) atomic_add(&F[row], Fe[cc][a]);
			if (Jv && basal) for (int b = 0; b < 3; ++b)
				atomic_add(&Jv[find_col(indptr, indices, row, cd[4 * cc + la[b]])], Ke[a][b]);
		}
	}
}

/* y = A x (CSR) and Jacobi-preconditioned CG (zero initial guess, stop on the true residual
 * norm): the Krylov part of cslvr/physics.py:673-678.                                  */
void bpo_spmv(int64_t n, const int64_t* indptr, const int32_t* indices, const double* v, const double* x, double* y)
{
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < n; ++i) {
		double s = 0.0;
		for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) s += v[k] * x[indices[k]];
		y[i] = s;
	}
}

static double dotp(int64_t n, const double* a, const double* b)
{
	double s = 0.0;
	#pragma omp parallel for reduction(+ : s) schedule(static)
	for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
	return s;
}

int bpo_pcg_jacobi(int64_t n, const int64_t* indptr, const int32_t* indices, const double* v,
                   const double* b, double* x, double rtol, double atol, int maxit, double* work /* 5n */)
{
	double *r = work, *z = work + n, *p = work + 2 * n, *Ap = work + 3 * n, *dinv = work + 4 * n;
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < n; ++i) {
		double d = 1.0;
		for (int64_t k = indptr[i]; k < indptr[i + 1]; ++k) if (indices[k] == i) d = v[k];
		dinv[i] = 1.0 / d; x[i] = 0.0; r[i] = b[i]; z[i] = dinv[i] * r[i]; p[i] = z[i];
	}
	double rz = dotp(n, r, z);
	const double norm0 = sqrt(dotp(n, r, r));          /* true residual norm, as the product */
	if (!(norm0 > atol)) return 0;
	int it = 0;
	while (it < maxit) {
		bpo_spmv(n, indptr, indices, v, p, Ap);
		const double alpha = rz / dotp(n, p, Ap);
		#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < n; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; z[i] = dinv[i] * r[i]; }
		const double rz1 = dotp(n, r, z), nz = sqrt(dotp(n, r, r));
		++it;
		if (!(nz > fmax(rtol * norm0, atol))) break;
		const double bt = rz1 / rz;
		rz = rz1;
		#pragma omp parallel for schedule(static)
		for (int64_t i = 0; i < n; ++i) p[i] = z[i] + bt * p[i];
	}
	return it;
}
