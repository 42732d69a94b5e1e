// bp_mg.cu -- aggregation multigrid preconditioner (BP_PC_MG) for the BP Jacobian.
//
// The reference preconditions KSPCG with hypre BoomerAMG (cslvr/momentumbp.py:183-184).
// This is the device-resident counterpart built for extruded ice meshes:
//   level 0 -> 1 : every ice column becomes one coarse node (the SSA-like, vertically
//                  constant space; thin-ice physics makes this the natural first coarsening);
//   level l -> l+1: greedy aggregation of the 2-D footprint graph (root + all its free
//                  neighbours), until <= 64 nodes remain;
//   prolongation : piecewise constant (2x2 identity per node), so the Galerkin product is a
//                  scatter-add of fine blocks into coarse blocks through a precomputed slot
//                  map -- no SpGEMM, and it is redone on the device every Newton step;
//   smoother     : Chebyshev polynomial in D^-1 A (same fused SpMV kernel as BP_PC_CHEBYSHEV),
//                  pre from a zero guess, post continuing from the corrected iterate;
//   coarsest     : dense inverse computed on the host (<= 128 x 128), applied by one kernel.
// All levels store 2x2 blocks in the same SELL-32 layout, so k_spmv / k_cheb_step / k_pc_setup
// / k_gershgorin are reused unchanged through light-weight level contexts.
// With several ranks the default is the COUPLED hierarchy further down (mgc_*: ghost nodes and a halo
// exchange on every level, global dense coarsest solve); BP_MG_COUPLED=0 selects the rank-local one
// (ghost columns ignored: block-Jacobi over the slabs with a V-cycle inside, no communication).
#include "bp_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <cstdlib>

struct bp_mg_level {
	bp_ctx* M = nullptr;                 // matrix + vectors of this level (level 0: the real context)
	int64_t n = 0;
	std::vector<int32_t> h_rowptr, h_col, h_sell_ptr;      // node-level pattern on the host
	// transfer to the next coarser level
	int32_t *d_agg = nullptr, *d_galmap = nullptr, *d_aggptr = nullptr, *d_aggidx = nullptr;
	double lmax = 0.0;
	// coupled hierarchy (partitioned meshes): ghost nodes of this level and their exchange lists
	int64_t n_ghost = 0;
	std::vector<int32_t> h_send_nodes;           // owned nodes whose values the neighbours need, grouped by neighbour
	std::vector<int64_t> h_send_ptr, h_recv_ptr; // [n_nbr + 1]
	std::vector<int32_t> ghost_owner_id;         // id of ghost g in its owner's numbering of this level
	std::vector<int32_t> ghost_nbr;              // neighbour index k that owns ghost g
};
struct bp_mg {
	std::vector<bp_mg_level> lv;
	double* d_dense = nullptr;           // inverse of the coarsest matrix, row-major
	// the V-cycle is a fixed sequence of ~50 mostly tiny launches: captured once as a CUDA graph
	cudaGraphExec_t graph = nullptr;
	const double *g_r = nullptr; double *g_z = nullptr, *g_z2 = nullptr;
	int g_deg = 0, g_smoother = -1; double g_ratio = 0.0, g_alpha = 0.0; int64_t g_launches = 0; bool g_f32 = false;
	std::vector<double> h_dense, h_val;
	int nc = 0;                          // coarsest dofs
	int gamma = 1;                       // coarse-grid visits per level below the finest
	bool no_graph = false;               // the context's stream cannot be captured
	int g_warm = 0;                      // direct applications before the coupled V-cycle is captured
	// coupled hierarchy: the coarsest level is gathered from all ranks into one dense system
	bool coupled = false, line0 = false;     // line0: Chebyshev on exact ice-column solves on level 0 (same decision on every rank)
	int my_rank = 0, n_ranks = 1, NG = 0;    // NG: nodes of the global coarsest level
	std::vector<int64_t> goff;               // first global coarsest node of every rank
	double *d_gmat = nullptr, *d_gvec = nullptr, *d_xs = nullptr;
	int32_t* d_gcol = nullptr;               // global id of every local (owned + ghost) coarsest node
	bool f32 = true;                     // the V-cycle reads FP32 copies of the level matrices (vectors and accumulation stay FP64)
	// coupled hierarchy, replicated tail: below a few ten thousand nodes (globally) a level is all exchange latency, so from
	// there down EVERY rank holds the whole operator (R: a context of its own with an ordinary single-rank hierarchy) and the
	// ranks only all-reduce the right-hand side of that level -- one collective instead of ~8 halo exchanges per level
	bp_ctx* R = nullptr;
	bp_mg_level Rlevel;                  // owns R's host pattern
	int32_t* d_rmap = nullptr;           // slot of the last coupled level -> slot of R
};

template <class T>
static int up(bp_ctx* c, T** dst, const T* src, size_t n)
{
	*dst = nullptr;
	if (n == 0) n = 1;
	BP_CUDA(c, cudaMalloc((void**)dst, n * sizeof(T)));
	if (src) BP_CUDA(c, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
	else BP_CUDA(c, cudaMemset(*dst, 0, n * sizeof(T)));
	return BP_OK;
}
#define UPM(dst, src, n) do { int rc_ = up(c, &(dst), (src), (size_t)(n)); if (rc_) return rc_; } while (0)

// ------------------------------------------------------------------ kernels
__global__ void k_mg_galerkin(int64_t nslots, const int32_t* __restrict__ galmap, const double* __restrict__ vf, double* vc)
{
	for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < nslots; s += (int64_t)gridDim.x * blockDim.x) {
		const int32_t cs = galmap[s];
		if (cs < 0) continue;
		const double2 a = ((const double2*)vf)[2 * s], b = ((const double2*)vf)[2 * s + 1];
		double* o = vc + 4 * (size_t)cs;
		atomicAdd(o, a.x); atomicAdd(o + 1, a.y); atomicAdd(o + 2, b.x); atomicAdd(o + 3, b.y);
	}
}

__global__ void __launch_bounds__(256)
k_mg_residual(int nrows, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ col, const double2* __restrict__ val,
              const double2* __restrict__ b, const double2* __restrict__ z, double2* __restrict__ out)
{
	const int lane = threadIdx.x & 31;
	const int nsl = (nrows + 31) >> 5;
	for (int slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slice < nsl; slice += (gridDim.x * blockDim.x) >> 5) {
		const int b0 = sell_ptr[slice], w = (sell_ptr[slice + 1] - b0) >> 5;
		const int row = (slice << 5) + lane;
		const int32_t* cp = col + b0 + lane;
		const double2* vp = val + 2 * ((size_t)b0 + lane);
		double s0 = 0.0, s1 = 0.0;
		#pragma unroll 4
		for (int k = 0; k < w; ++k) {
			const double2 a01 = vp[64 * (size_t)k], a23 = vp[64 * (size_t)k + 1];
			const double2 xv = z[cp[32 * k]];
			s0 += a01.x * xv.x + a01.y * xv.y; s1 += a23.x * xv.x + a23.y * xv.y;
		}
		if (row < nrows) { const double2 bv = b[row]; out[row] = make_double2(bv.x - s0, bv.y - s1); }
	}
}

__global__ void k_mg_restrict(int nc, const int32_t* __restrict__ aggptr, const int32_t* __restrict__ aggidx,
                              const double2* __restrict__ rf, double2* __restrict__ bc)
{
	for (int I = blockIdx.x * blockDim.x + threadIdx.x; I < nc; I += gridDim.x * blockDim.x) {
		double s0 = 0.0, s1 = 0.0;
		for (int k = aggptr[I]; k < aggptr[I + 1]; ++k) { const double2 v = rf[aggidx[k]]; s0 += v.x; s1 += v.y; }
		bc[I] = make_double2(s0, s1);
	}
}

__global__ void k_mg_prolong_add(int nf, const int32_t* __restrict__ agg, const double2* __restrict__ xc, double2* __restrict__ zf, double alpha)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) {
		const int I = agg[i];
		if (I < 0) continue;
		const double2 e = xc[I];
		double2 z = zf[i];
		z.x += alpha * e.x; z.y += alpha * e.y;
		zf[i] = z;
	}
}

__global__ void k_mg_dense(int n, const double* __restrict__ Ainv, const double* __restrict__ b, double* __restrict__ x)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	double s = 0.0;
	for (int j = 0; j < n; ++j) s += Ainv[(size_t)i * n + j] * b[j];
	x[i] = s;
}


// ------------------------------------------------------------------ the small levels in ONE kernel
// Below a few thousand nodes a level is one or two CTAs' worth of work and its ~12 launches per
// visit cost more than their arithmetic.  k_mg_tail runs the whole lower part of the V-cycle
// (pre-smoothing, residual, restriction ... dense coarsest solve ... prolongation, post-smoothing)
// for the levels >= l_tail in a single CTA of 1024 threads, __syncthreads() between the phases.
struct TailLevel {
	int n, nc;                           // nodes of this level / of the next coarser one
	const int32_t *sell_ptr, *col;
	const float4* valf;                  // FP32 copy (or nullptr -> val)
	const double2* val;
	const double* dinv;
	const double* lmax;                  // bound of lmax(D^-1 A), on the device
	double2 *F, *z, *z2, *cd, *io;       // rhs (written by the level above), iterates, direction, residual
	const int32_t *agg, *aggptr, *aggidx;
};
#define MG_TAIL_MAX 8
struct TailArgs { TailLevel lv[MG_TAIL_MAX]; int nl; int deg; double ratio, alpha; const double* dense; int ndense; };

__device__ __forceinline__ double2 tail_dinv(const double* dinv, int i, double2 r)
{
	const double2 d01 = ((const double2*)dinv)[2 * (size_t)i], d23 = ((const double2*)dinv)[2 * (size_t)i + 1];
	return make_double2(d01.x * r.x + d01.y * r.y, d23.x * r.x + d23.y * r.y);
}
// s = A x for the 32 rows of one slice (one warp)
__device__ __forceinline__ void tail_row(const TailLevel& L, int slice, int lane, const double2* x, double& s0, double& s1)
{
	const int b0 = L.sell_ptr[slice], w = (L.sell_ptr[slice + 1] - b0) >> 5;
	const int32_t* cp = L.col + b0 + lane;
	s0 = 0.0; s1 = 0.0;
	for (int k = 0; k < w; ++k) {
		double a0, a1, a2, a3;
		const size_t slot = (size_t)b0 + lane + 32 * (size_t)k;
		if (L.valf) { const float4 v = L.valf[slot]; a0 = v.x; a1 = v.y; a2 = v.z; a3 = v.w; }
		else { const double2 p = L.val[2 * slot], q = L.val[2 * slot + 1]; a0 = p.x; a1 = p.y; a2 = q.x; a3 = q.y; }
		const double2 xv = x[cp[32 * k]];
		s0 += a0 * xv.x + a1 * xv.y; s1 += a2 * xv.x + a3 * xv.y;
	}
}
__device__ void tail_cheb_step(const TailLevel& L, double c1, double c2, const double2* r, const double2* zold, double2* znew)
{
	const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31, nsl = (L.n + 31) >> 5;
	for (int slice = warp; slice < nsl; slice += nw) {
		double s0, s1;
		tail_row(L, slice, lane, zold, s0, s1);
		const int row = (slice << 5) + lane;
		if (row < L.n) {
			const double2 rv = r[row], zo = zold[row], dv = L.cd[row];
			const double2 pr = tail_dinv(L.dinv, row, make_double2(rv.x - s0, rv.y - s1));
			const double2 dn = make_double2(c1 * dv.x + c2 * pr.x, c1 * dv.y + c2 * pr.y);
			L.cd[row] = dn;
			znew[row] = make_double2(zo.x + dn.x, zo.y + dn.y);
		}
	}
	__syncthreads();
}
// Chebyshev smoothing exactly as cheb_smooth(): returns the buffer holding the result
__device__ double2* tail_smooth(const TailLevel& L, int deg, double ratio, const double2* r, double2* za, double2* zb, bool zero_init)
{
	const double a = 1.0 / ratio, theta = 0.5 * (1.0 + a), delta = 0.5 * (1.0 - a), sigma = theta / delta;
	const double ilm = 1.0 / *L.lmax;
	double rho = 1.0 / sigma;
	if (zero_init) {
		const double c0 = ilm / theta;
		for (int i = threadIdx.x; i < L.n; i += blockDim.x) {
			const double2 v = tail_dinv(L.dinv, i, r[i]);
			const double2 dv = make_double2(c0 * v.x, c0 * v.y);
			L.cd[i] = dv; za[i] = dv;
		}
		__syncthreads();
	} else {
		tail_cheb_step(L, 0.0, ilm / theta, r, za, zb);
		double2* t = za; za = zb; zb = t;
	}
	for (int k = 1; k < deg; ++k) {
		const double rho_n = 1.0 / (2.0 * sigma - rho);
		tail_cheb_step(L, rho_n * rho, 2.0 * rho_n / delta * ilm, r, za, zb);
		double2* t = za; za = zb; zb = t;
		rho = rho_n;
	}
	return za;
}
__global__ void __launch_bounds__(1024, 1)
k_mg_tail(const TailArgs A)
{
	__shared__ double2* zres[MG_TAIL_MAX];
	const int nl = A.nl;                  // levels in the tail; the last one is the dense coarsest level
	// ---- down
	for (int l = 0; l + 1 < nl; ++l) {
		const TailLevel& L = A.lv[l];
		double2* z = tail_smooth(L, A.deg, A.ratio, L.F, L.z, L.z2, true);
		if (threadIdx.x == 0) zres[l] = z;
		// residual -> io
		const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31, nsl = (L.n + 31) >> 5;
		for (int slice = warp; slice < nsl; slice += nw) {
			double s0, s1;
			tail_row(L, slice, lane, z, s0, s1);
			const int row = (slice << 5) + lane;
			if (row < L.n) { const double2 bv = L.F[row]; L.io[row] = make_double2(bv.x - s0, bv.y - s1); }
		}
		__syncthreads();
		// restriction -> rhs of the next level
		double2* bc = A.lv[l + 1].F;
		for (int I = threadIdx.x; I < L.nc; I += blockDim.x) {
			double s0 = 0.0, s1 = 0.0;
			for (int k = L.aggptr[I]; k < L.aggptr[I + 1]; ++k) { const double2 v = L.io[L.aggidx[k]]; s0 += v.x; s1 += v.y; }
			bc[I] = make_double2(s0, s1);
		}
		__syncthreads();
	}
	// ---- coarsest: dense inverse
	{
		const TailLevel& C = A.lv[nl - 1];
		const double* b = (const double*)C.F;
		double* x = (double*)C.z;
		for (int i = threadIdx.x; i < A.ndense; i += blockDim.x) {
			double sacc = 0.0;
			for (int j = 0; j < A.ndense; ++j) sacc += A.dense[(size_t)i * A.ndense + j] * b[j];
			x[i] = sacc;
		}
		__syncthreads();
	}
	// ---- up
	for (int l = nl - 2; l >= 0; --l) {
		const TailLevel& L = A.lv[l];
		double2* z = zres[l];
		const double2* xc = (l + 1 == nl - 1) ? A.lv[l + 1].z : zres[l + 1];
		for (int i = threadIdx.x; i < L.n; i += blockDim.x) {
			const int I = L.agg[i];
			if (I < 0) continue;
			const double2 e = xc[I];
			double2 v = z[i];
			v.x += A.alpha * e.x; v.y += A.alpha * e.y;
			z[i] = v;
		}
		__syncthreads();
		double2* other = (z == L.z) ? L.z2 : L.z;
		z = tail_smooth(L, A.deg, A.ratio, L.F, z, other, false);
		if (threadIdx.x == 0) zres[l] = z;
		__syncthreads();
	}
	// the level above reads the correction from lv[0].z
	if (nl > 1) {
		const TailLevel& L = A.lv[0];
		double2* z = zres[0];
		if (z != L.z) for (int i = threadIdx.x; i < L.n; i += blockDim.x) L.z[i] = z[i];
	}
}

static inline int blocks(int64_t n, int per)
{
	int64_t b = (n + per - 1) / per;
	return (int)std::min<int64_t>(std::max<int64_t>(b, 1), 148 * 8);
}

// ------------------------------------------------------------------ symbolic set-up (host, once)
static void greedy_aggregate(const std::vector<int32_t>& rp, const std::vector<int32_t>& ci, int64_t n, std::vector<int32_t>& agg, int32_t& na)
{
	agg.assign(n, -1);
	na = 0;
	for (int64_t i = 0; i < n; ++i) {                // roots: nodes whose whole neighbourhood is free
		if (agg[i] >= 0) continue;
		bool free_nb = true;
		for (int32_t k = rp[i]; k < rp[i + 1]; ++k) if (agg[ci[k]] >= 0) { free_nb = false; break; }
		if (!free_nb) continue;
		for (int32_t k = rp[i]; k < rp[i + 1]; ++k) agg[ci[k]] = na;
		agg[i] = na++;
	}
	for (int64_t i = 0; i < n; ++i) {                // leftovers join a neighbouring aggregate
		if (agg[i] >= 0) continue;
		for (int32_t k = rp[i]; k < rp[i + 1]; ++k) if (agg[ci[k]] >= 0 && ci[k] != i) { agg[i] = agg[ci[k]]; break; }
		if (agg[i] < 0) agg[i] = na++;
	}
}

static int make_level_ctx(bp_ctx* c, bp_mg_level& L, const std::vector<int32_t>& diag_slot, const std::vector<int32_t>& sell_col, int64_t n_slots)
{
	bp_ctx* M = new bp_ctx();
	L.M = M;
	M->device = c->device; M->stream = c->stream; M->prm = c->prm; M->env = c->env;
	M->prm.pc = BP_PC_BLOCK_JACOBI;
	M->nn_owned = L.n;
	M->nn = L.n + L.n_ghost;
	M->n_slots = n_slots;
	UPM(M->d_sell_ptr, L.h_sell_ptr.data(), L.h_sell_ptr.size());
	UPM(M->d_col, sell_col.data(), n_slots);
	UPM(M->d_diag, diag_slot.data(), L.n);
	UPM(M->d_val, (const double*)nullptr, (size_t)n_slots * 4);
	UPM(M->d_dinv, (const double*)nullptr, (size_t)L.n * 4);
	double** vecs[] = { &M->d_F, &M->d_z, &M->d_z2, &M->d_cd, &M->d_io };
	for (double** v : vecs) UPM(*v, (const double*)nullptr, (size_t)2 * (L.n + L.n_ghost));
	UPM(M->d_sc, (const double*)nullptr, SC_COUNT);
	UPM(M->d_part, (const double*)nullptr, 4 * 4096);
	UPM(M->d_counter, (const unsigned*)nullptr, 16);
	BP_CUDA(c, cudaMallocHost((void**)&M->h_sc, SC_COUNT * sizeof(double)));
	return BP_OK;
}

// One coarsening step: from the finest level built so far (mg->lv.back()) and the aggregate of each
// of its nodes -- agg[0 .. F.n) for the owned nodes and, in a coupled hierarchy, agg[F.n ..) for the
// ghost nodes (their aggregates are ghost nodes of the coarse level, numbered na .. na + n_cghost) --
// build the coarse pattern, the transfer maps and the level context.  `halo`: exchange lists of the
// coarse level (coupled hierarchy) or nullptr.
static int build_level(bp_ctx* c, bp_mg* mg, const std::vector<int32_t>& agg, int32_t na, int32_t n_cghost, const bp_mg_level* halo)
{
	bp_mg_level& F = mg->lv.back();
	const int64_t nf = F.n;
	// members of every aggregate
	std::vector<int32_t> aggptr(na + 1, 0), aggidx;
	for (int64_t i = 0; i < nf; ++i) if (agg[i] >= 0) aggptr[agg[i] + 1]++;
	for (int32_t I = 0; I < na; ++I) aggptr[I + 1] += aggptr[I];
	aggidx.resize(aggptr[na]);
	{
		std::vector<int32_t> fill(aggptr.begin(), aggptr.end() - 1);
		for (int64_t i = 0; i < nf; ++i) if (agg[i] >= 0) aggidx[fill[agg[i]]++] = (int32_t)i;
	}
	// coarse pattern: I ~ J iff some fine i in I has a column j in J (owned columns only)
	bp_mg_level C;
	C.n = na;
	C.n_ghost = n_cghost;
	C.h_rowptr.assign(na + 1, 0);
	std::vector<std::vector<int32_t>> rows(na);
	#pragma omp parallel for schedule(dynamic, 64)
	for (int32_t I = 0; I < na; ++I) {
		std::vector<int32_t>& r = rows[I];
		for (int32_t m = aggptr[I]; m < aggptr[I + 1]; ++m) {
			const int32_t i = aggidx[m];
			for (int32_t k = F.h_rowptr[i]; k < F.h_rowptr[i + 1]; ++k) {
				const int32_t j = F.h_col[k];
				if (j < (int64_t)agg.size() && agg[j] >= 0) r.push_back(agg[j]);
			}
		}
		std::sort(r.begin(), r.end());
		r.erase(std::unique(r.begin(), r.end()), r.end());
	}
	for (int32_t I = 0; I < na; ++I) C.h_rowptr[I + 1] = C.h_rowptr[I] + (int32_t)rows[I].size();
	C.h_col.resize(C.h_rowptr[na]);
	for (int32_t I = 0; I < na; ++I) std::copy(rows[I].begin(), rows[I].end(), C.h_col.begin() + C.h_rowptr[I]);
	// SELL-32 slots of the coarse level
	const int64_t nsl = ((int64_t)na + 31) / 32;
	C.h_sell_ptr.assign(nsl + 1, 0);
	for (int64_t s = 0; s < nsl; ++s) {
		int32_t w = 0;
		for (int64_t I = s * 32; I < std::min<int64_t>(na, s * 32 + 32); ++I) w = std::max(w, C.h_rowptr[I + 1] - C.h_rowptr[I]);
		C.h_sell_ptr[s + 1] = C.h_sell_ptr[s] + 32 * w;
	}
	const int64_t nslots_c = C.h_sell_ptr[nsl];
	auto slot_c = [&](int64_t I, int32_t k) -> int32_t { return C.h_sell_ptr[I >> 5] + 32 * k + (int32_t)(I & 31); };
	std::vector<int32_t> sell_col((size_t)std::max<int64_t>(nslots_c, 1), 0), diag(na, -1);
	for (int64_t I = 0; I < nsl * 32; ++I) {
		const int32_t w = (C.h_sell_ptr[(I >> 5) + 1] - C.h_sell_ptr[I >> 5]) >> 5;
		const int32_t len = I < na ? C.h_rowptr[I + 1] - C.h_rowptr[I] : 0;
		for (int32_t k = 0; k < w; ++k) {
			const int32_t cj = (k < len) ? C.h_col[C.h_rowptr[I] + k] : (I < na ? (int32_t)I : 0);
			sell_col[slot_c(I, k)] = cj;
			if (k < len && cj == I) diag[I] = slot_c(I, k);
		}
	}
	// fine slot -> coarse slot
	const int64_t nslots_f = F.h_sell_ptr.back();
	std::vector<int32_t> galmap((size_t)std::max<int64_t>(nslots_f, 1), -1);
	#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < nf; ++i) {
		const int32_t I = agg[i];
		if (I < 0) continue;
		const int32_t* lo = C.h_col.data() + C.h_rowptr[I];
		const int32_t* hi = C.h_col.data() + C.h_rowptr[I + 1];
		for (int32_t k = F.h_rowptr[i]; k < F.h_rowptr[i + 1]; ++k) {
			const int32_t j = F.h_col[k];
			if (j >= (int64_t)agg.size() || agg[j] < 0) continue;
			const int32_t kc = (int32_t)(std::lower_bound(lo, hi, agg[j]) - lo);
			galmap[F.h_sell_ptr[i >> 5] + 32 * (k - F.h_rowptr[i]) + (i & 31)] = slot_c(I, kc);
		}
	}
	UPM(F.d_agg, agg.data(), nf);
	UPM(F.d_galmap, galmap.data(), galmap.size());
	UPM(F.d_aggptr, aggptr.data(), aggptr.size());
	UPM(F.d_aggidx, aggidx.data(), aggidx.size());
	if (halo) { C.h_send_nodes = halo->h_send_nodes; C.h_send_ptr = halo->h_send_ptr; C.h_recv_ptr = halo->h_recv_ptr; C.ghost_owner_id = halo->ghost_owner_id; C.ghost_nbr = halo->ghost_nbr; }
	int rc = make_level_ctx(c, C, diag, sell_col, nslots_c);
	if (rc) return rc;
	mg->lv.push_back(std::move(C));
	return BP_OK;

}

static int mg_symbolic(bp_ctx* c)
{
	bp_mg* mg = new bp_mg();
	c->mg = mg;
	const int64_t no = c->nn_owned;
	mg->lv.emplace_back();
	{
		bp_mg_level& L0 = mg->lv.back();
		L0.M = c; L0.n = no;
		L0.h_rowptr = c->h_rowptr; L0.h_col = c->h_col; L0.h_sell_ptr = c->h_sell_ptr;
	}
	// Extruded structure of the current level: node -> (column, layer).  The order of the two kinds
	// of coarsening follows the anisotropy (c->aspect = horizontal / vertical extent of the tets):
	//   * flat elements (aspect >= 1, real ice sheets): the vertical coupling is the strong one, so
	//     every ice column collapses into one node first (the SSA-like, vertically constant space);
	//   * tall elements (ISMIP-HOM at high horizontal resolution): the strong coupling is horizontal;
	//     collapsing columns would leave the horizontally smooth, vertically varying error to the
	//     smoother, so the footprint is aggregated layer by layer (coarse node = 2-D aggregate x layer)
	//     until the aggregates are wide enough, and only then do the columns collapse.
	std::vector<int32_t> col_of(no, -1), lay_of(no, 0);
	int64_t ncol = c->n_cols;
	int32_t nlay = 0;
	bool uniform = true;
	for (int64_t cc = 0; cc < c->n_cols; ++cc) {
		const int32_t len = c->h_colptr[cc + 1] - c->h_colptr[cc];
		if (cc == 0) nlay = len; else if (len != nlay) uniform = false;
		for (int32_t j = c->h_colptr[cc]; j < c->h_colptr[cc + 1]; ++j) { col_of[c->h_colnode[j]] = (int32_t)cc; lay_of[c->h_colnode[j]] = j - c->h_colptr[cc]; }
	}
	double aspect = c->aspect;
	// slabs of a partitioned mesh keep the plain order: with the ghost columns ignored the layer-wise
	// levels and the over-relaxation both cost iterations there (measured on 4 and 8 slabs)
	const bool partitioned = c->nn > c->nn_owned;
	const double aspect_switch = c->env.has_horizontal_first ? c->env.mg_horizontal_first : (partitioned ? 0.0 : 1.0);
	std::vector<int32_t> agg;
	int32_t na = 0;
	auto next_aggregation = [&](const bp_mg_level& F) {
		const int64_t nf = F.n;
		if (nlay > 1 && uniform && aspect < aspect_switch) {
			// footprint graph of the columns, aggregated; layers kept
			std::vector<std::vector<int32_t>> nb(ncol);
			for (int64_t i = 0; i < nf; ++i)
				for (int32_t k = F.h_rowptr[i]; k < F.h_rowptr[i + 1]; ++k) {
					const int32_t j = F.h_col[k];
					if (j < nf && col_of[j] >= 0 && col_of[j] != col_of[i]) nb[col_of[i]].push_back(col_of[j]);
				}
			std::vector<int32_t> rp(ncol + 1, 0), ci;
			for (int64_t cc = 0; cc < ncol; ++cc) {
				std::sort(nb[cc].begin(), nb[cc].end());
				nb[cc].erase(std::unique(nb[cc].begin(), nb[cc].end()), nb[cc].end());
				rp[cc + 1] = rp[cc] + (int32_t)nb[cc].size();
			}
			ci.reserve(rp[ncol]);
			for (int64_t cc = 0; cc < ncol; ++cc) ci.insert(ci.end(), nb[cc].begin(), nb[cc].end());
			std::vector<int32_t> cagg;
			int32_t nca = 0;
			greedy_aggregate(rp, ci, ncol, cagg, nca);
			agg.assign(nf, -1);
			for (int64_t i = 0; i < nf; ++i) if (col_of[i] >= 0) agg[i] = cagg[col_of[i]] * nlay + lay_of[i];
			na = nca * nlay;
			aspect *= std::sqrt((double)ncol / std::max(nca, 1));
			ncol = nca;
			col_of.assign(na, 0); lay_of.assign(na, 0);
			for (int32_t I = 0; I < na; ++I) { col_of[I] = I / nlay; lay_of[I] = I % nlay; }
		} else if (nlay > 1 || !uniform) {
			// ice columns -> one node each
			agg.assign(col_of.begin(), col_of.begin() + nf);
			na = (int32_t)ncol;
			nlay = 1; uniform = true;
			col_of.assign(na, 0); lay_of.assign(na, 0);
			for (int32_t I = 0; I < na; ++I) col_of[I] = I;
		} else {
			greedy_aggregate(F.h_rowptr, F.h_col, nf, agg, na);
			ncol = na;
			col_of.assign(na, 0); lay_of.assign(na, 0);
			for (int32_t I = 0; I < na; ++I) col_of[I] = I;
		}
	};
	for (int guard = 0; guard < 24; ++guard) {
		{
			const int64_t n_before = mg->lv.back().n;
			next_aggregation(mg->lv.back());
			if (guard > 0 && na >= n_before) break;          // no coarsening possible
		}
		{ int rc_ = build_level(c, mg, agg, na, 0, nullptr); if (rc_) return rc_; }
		if (na <= 64) break;
	}
	mg->nc = (int)(2 * mg->lv.back().n);
	if (mg->nc > 4096) return bp_fail(c, BP_ERR_STATE, "multigrid: coarsest level has %d dofs (aggregation stalled)", mg->nc);
	mg->h_dense.resize((size_t)mg->nc * mg->nc);
	UPM(mg->d_dense, (const double*)nullptr, (size_t)mg->nc * mg->nc);
	if (c->prm.verbose) {
		printf("multigrid levels:");
		for (auto& L : mg->lv) printf(" %lld", (long long)L.n);
		printf(" nodes\n");
	}
	return BP_OK;
}


// ====================================================================== coupled hierarchy
// On a partitioned mesh the rank-local hierarchy above ignores every coupling across the slab
// interfaces on every level, and CG pays for it (2x the iterations on 2 slabs, 4-5x on 8).  The
// coupled hierarchy keeps the aggregates local to their rank but keeps the couplings: every level
// has ghost nodes (the neighbours' aggregates that touch this rank), its operator has the columns
// of those ghosts, and every smoothing step / residual is preceded by a halo exchange of the
// iterate on that level -- the same exchange the fine SpMV uses (bp_comm_halo: grouped
// ncclSend/ncclRecv, or device copies inside an in-process group).  The coarsest level is gathered
// into one dense system that every rank inverts (all-reduce of the matrix per Newton step and of the
// right-hand side per V-cycle).  All ranks walk through the set-up and the V-cycle in lock-step.
static int mgc_level_ctxs(bp_ctx** cs, int n, int l, bp_ctx** out)
{
	for (int r = 0; r < n; ++r) out[r] = cs[r]->mg->lv[l].M;
	return BP_OK;
}

// sum (or max) over all ranks of `count` doubles per rank (vals[r][0..count)), result back in vals
static int mgc_allreduce_host(bp_ctx** cs, int n, std::vector<std::vector<double>>& vals, size_t count, int op_max)
{
	double* bufs[16];
	for (int r = 0; r < n; ++r) {
		bufs[r] = cs[r]->mg->d_xs;
		BP_CUDA(cs[r], cudaMemcpyAsync(bufs[r], vals[r].data(), count * sizeof(double), cudaMemcpyHostToDevice, cs[r]->stream));
	}
	int rc = bp_comm_allreduce_buf(cs, n, bufs, count, op_max);
	if (rc) return rc;
	for (int r = 0; r < n; ++r) {
		BP_CUDA(cs[r], cudaMemcpyAsync(vals[r].data(), bufs[r], count * sizeof(double), cudaMemcpyDeviceToHost, cs[r]->stream));
		BP_CUDA(cs[r], cudaStreamSynchronize(cs[r]->stream));
	}
	return BP_OK;
}

// the owner's id of every ghost node: ids[r][i] for the owned nodes of level l travel through the halo lists
static int mgc_exchange_ids(bp_ctx** cs, int n, int l, const std::vector<std::vector<int32_t>>& ids, std::vector<std::vector<int32_t>>& ghost_ids)
{
	bp_ctx* Ls[16];
	mgc_level_ctxs(cs, n, l, Ls);
	std::vector<double> v;
	for (int r = 0; r < n; ++r) {
		bp_ctx* M = Ls[r];
		v.assign((size_t)2 * M->nn, 0.0);
		for (int64_t i = 0; i < M->nn_owned; ++i) v[2 * i] = (double)ids[r][i];
		BP_CUDA(M, cudaMemcpyAsync(M->d_z, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice, M->stream));
		BP_CUDA(M, cudaStreamSynchronize(M->stream));
	}
	int rc = bp_comm_halo(Ls, n, 2);
	if (rc) { if (Ls[0] != cs[0]) cs[0]->err = Ls[0]->err; return rc; }
	ghost_ids.assign(n, std::vector<int32_t>());
	for (int r = 0; r < n; ++r) {
		bp_ctx* M = Ls[r];
		const int64_t ng = M->nn - M->nn_owned;
		v.assign((size_t)2 * std::max<int64_t>(ng, 1), 0.0);
		if (ng) BP_CUDA(M, cudaMemcpyAsync(v.data(), M->d_z + 2 * M->nn_owned, (size_t)2 * ng * sizeof(double), cudaMemcpyDeviceToHost, M->stream));
		BP_CUDA(M, cudaStreamSynchronize(M->stream));
		ghost_ids[r].resize(ng);
		for (int64_t g = 0; g < ng; ++g) ghost_ids[r][g] = (int32_t)v[2 * g];
	}
	return BP_OK;
}

// give a level context the exchange lists of its level
static int mgc_attach_halo(bp_ctx* c, bp_mg_level& L)
{
	bp_ctx* M = L.M;
	M->n_nbr = c->n_nbr;
	M->nbr_rank = c->nbr_rank;
	M->send_ptr = L.h_send_ptr;
	M->recv_ptr = L.h_recv_ptr;
	M->n_send = L.h_send_ptr.empty() ? 0 : L.h_send_ptr.back();
	M->comm = c->comm;                       // shared, owned by the fine context
	UPM(M->d_send_nodes, L.h_send_nodes.data(), L.h_send_nodes.size());
	UPM(M->d_sendbuf, (const double*)nullptr, (size_t)2 * std::max<int64_t>(M->n_send, 1));
	M->device = c->device;
	M->env = c->env;
	return bp_comm_p2p_attach(M);            // collective: every rank attaches its levels in the same order
}

// ---- replicated tail (see bp_mg::R): the pattern of the last coupled level is gathered from all ranks (every rank adds its
// rows, in global numbering, to a zero-padded table that is all-reduced), every rank builds the same SELL layout from it and
// an ordinary single-rank hierarchy below it.
static int mgc_build_replicated(bp_ctx** cs, int n)
{
	int rc;
	const int NG = cs[0]->mg->NG;
	// widest row, over all ranks
	std::vector<std::vector<double>> wv(n, std::vector<double>(1, 0.0));
	for (int r = 0; r < n; ++r) {
		const bp_mg_level& Lc = cs[r]->mg->lv.back();
		for (int64_t i = 0; i < Lc.n; ++i) wv[r][0] = std::max(wv[r][0], (double)(Lc.h_rowptr[i + 1] - Lc.h_rowptr[i]));
	}
	if ((rc = mgc_allreduce_host(cs, n, wv, 1, 1))) return rc;
	const int maxw = (int)wv[0][0];
	const size_t cnt = (size_t)NG * maxw;
	// the table of global column ids (+1; 0 = none), one device buffer per context
	double* bufs[16];
	std::vector<double> tab(cnt);
	for (int r = 0; r < n; ++r) {
		bp_ctx* c = cs[r];
		bp_mg* mg = c->mg;
		const bp_mg_level& Lc = mg->lv.back();
		std::vector<int32_t> gcol(Lc.n + Lc.n_ghost);
		for (int64_t i = 0; i < Lc.n; ++i) gcol[i] = (int32_t)(mg->goff[mg->my_rank] + i);
		for (int64_t g = 0; g < Lc.n_ghost; ++g) gcol[Lc.n + g] = (int32_t)(mg->goff[c->nbr_rank[Lc.ghost_nbr[g]]] + Lc.ghost_owner_id[g]);
		std::fill(tab.begin(), tab.end(), 0.0);
		for (int64_t i = 0; i < Lc.n; ++i) {
			double* row = tab.data() + (size_t)(mg->goff[mg->my_rank] + i) * maxw;
			for (int32_t k = Lc.h_rowptr[i]; k < Lc.h_rowptr[i + 1]; ++k) row[k - Lc.h_rowptr[i]] = (double)gcol[Lc.h_col[k]] + 1.0;
		}
		BP_CUDA(c, cudaMalloc((void**)&bufs[r], std::max<size_t>(cnt, 1) * sizeof(double)));
		BP_CUDA(c, cudaMemcpyAsync(bufs[r], tab.data(), cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
		BP_CUDA(c, cudaStreamSynchronize(c->stream));
	}
	if ((rc = bp_comm_allreduce_buf(cs, n, bufs, cnt, 0))) return rc;
	for (int r = 0; r < n; ++r) {
		bp_ctx* c = cs[r];
		bp_mg* mg = c->mg;
		BP_CUDA(c, cudaMemcpyAsync(tab.data(), bufs[r], cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		BP_CUDA(c, cudaStreamSynchronize(c->stream));
		cudaFree(bufs[r]);
		bp_mg_level& G = mg->Rlevel;
		G.n = NG; G.n_ghost = 0;
		G.h_rowptr.assign(NG + 1, 0);
		G.h_col.clear();
		for (int I = 0; I < NG; ++I) {
			std::vector<int32_t> row;
			for (int k = 0; k < maxw; ++k) { const double v = tab[(size_t)I * maxw + k]; if (v > 0.0) row.push_back((int32_t)(v - 1.0)); }
			std::sort(row.begin(), row.end());
			row.erase(std::unique(row.begin(), row.end()), row.end());
			G.h_col.insert(G.h_col.end(), row.begin(), row.end());
			G.h_rowptr[I + 1] = (int32_t)G.h_col.size();
		}
		const int64_t nsl = ((int64_t)NG + 31) / 32;
		G.h_sell_ptr.assign(nsl + 1, 0);
		for (int64_t sl = 0; sl < nsl; ++sl) {
			int32_t w = 0;
			for (int64_t I = sl * 32; I < std::min<int64_t>(NG, sl * 32 + 32); ++I) w = std::max(w, G.h_rowptr[I + 1] - G.h_rowptr[I]);
			G.h_sell_ptr[sl + 1] = G.h_sell_ptr[sl] + 32 * w;
		}
		const int64_t nslots = G.h_sell_ptr[nsl];
		auto slot_g = [&](int64_t I, int32_t k) -> int32_t { return G.h_sell_ptr[I >> 5] + 32 * k + (int32_t)(I & 31); };
		std::vector<int32_t> sell_col((size_t)std::max<int64_t>(nslots, 1), 0), diag(NG, -1);
		for (int64_t I = 0; I < nsl * 32; ++I) {
			const int32_t w = (G.h_sell_ptr[(I >> 5) + 1] - G.h_sell_ptr[I >> 5]) >> 5;
			const int32_t len = I < NG ? G.h_rowptr[I + 1] - G.h_rowptr[I] : 0;
			for (int32_t k = 0; k < w; ++k) {
				const int32_t cj = (k < len) ? G.h_col[G.h_rowptr[I] + k] : (I < NG ? (int32_t)I : 0);
				sell_col[slot_g(I, k)] = cj;
				if (k < len && cj == I) diag[I] = slot_g(I, k);
			}
		}
		if ((rc = make_level_ctx(c, G, diag, sell_col, nslots))) return rc;
		bp_ctx* R = G.M;
		mg->R = R;
		// R doubles as the "fine context" of a single-rank hierarchy of its own (bp_mg_prepare / bp_mg_apply)
		R->nn = R->nn_owned = NG;
		R->h_rowptr = G.h_rowptr; R->h_col = G.h_col; R->h_sell_ptr = G.h_sell_ptr;
		R->n_cols = 0; R->max_layers = 0; R->aspect = 1.0; R->nnzb = G.h_rowptr[NG];
		R->prm.pc = BP_PC_MG; R->prm.mg_smoother = 0;
		R->h_colptr.assign(1, 0);
		UPM(R->d_r, (const double*)nullptr, (size_t)2 * NG);
		// slot of the last coupled level -> slot of R
		const bp_mg_level& Lc = mg->lv.back();
		std::vector<int32_t> gcol(Lc.n + Lc.n_ghost);
		for (int64_t i = 0; i < Lc.n; ++i) gcol[i] = (int32_t)(mg->goff[mg->my_rank] + i);
		for (int64_t g = 0; g < Lc.n_ghost; ++g) gcol[Lc.n + g] = (int32_t)(mg->goff[c->nbr_rank[Lc.ghost_nbr[g]]] + Lc.ghost_owner_id[g]);
		std::vector<int32_t> rmap((size_t)std::max<int64_t>(Lc.h_sell_ptr.back(), 1), -1);
		for (int64_t i = 0; i < Lc.n; ++i) {
			const int64_t I = mg->goff[mg->my_rank] + i;
			const int32_t* lo = G.h_col.data() + G.h_rowptr[I];
			const int32_t* hi = G.h_col.data() + G.h_rowptr[I + 1];
			for (int32_t k = Lc.h_rowptr[i]; k < Lc.h_rowptr[i + 1]; ++k) {
				const int32_t kc = (int32_t)(std::lower_bound(lo, hi, gcol[Lc.h_col[k]]) - lo);
				rmap[Lc.h_sell_ptr[i >> 5] + 32 * (k - Lc.h_rowptr[i]) + (i & 31)] = slot_g(I, kc);
			}
		}
		UPM(mg->d_rmap, rmap.data(), rmap.size());
		if (c->prm.verbose && mg->my_rank == 0)
			printf("coupled multigrid: replicated tail from %d global nodes (%lld blocks) on every rank\n", NG, (long long)G.h_rowptr[NG]);
	}
	return BP_OK;
}

static int mgc_symbolic(bp_ctx** cs, int n)
{
	int rc;
	const int n_ranks = (n > 1) ? n : bp_comm_size(cs[0]);
	if (n_ranks > 16) return bp_fail(cs[0], BP_ERR_ARG, "coupled multigrid supports at most 16 ranks");
	for (int r = 0; r < n; ++r) {
		bp_ctx* c = cs[r];
		bp_mg* mg = new bp_mg();
		c->mg = mg;
		mg->coupled = true;
		mg->n_ranks = n_ranks;
		mg->my_rank = (n > 1) ? r : bp_comm_rank(c);
		UPM(mg->d_xs, (const double*)nullptr, 64);
		mg->lv.emplace_back();
		bp_mg_level& L0 = mg->lv.back();
		L0.M = c; L0.n = c->nn_owned; L0.n_ghost = c->nn - c->nn_owned;
		L0.h_rowptr = c->h_rowptr; L0.h_col = c->h_col; L0.h_sell_ptr = c->h_sell_ptr;
		L0.h_send_nodes = c->h_send_nodes; L0.h_send_ptr = c->send_ptr; L0.h_recv_ptr = c->recv_ptr;
		if (L0.h_send_ptr.empty()) { L0.h_send_ptr.assign(c->n_nbr + 1, 0); L0.h_recv_ptr.assign(c->n_nbr + 1, 0); }
	}
	{	// the level-0 smoother must be the same on every rank: decide on the global minimum of the aspect ratio
		std::vector<std::vector<double>> v(n, std::vector<double>(2, 0.0));
		for (int r = 0; r < n; ++r) { v[r][0] = -cs[r]->aspect; v[r][1] = (cs[r]->max_layers <= 64 && cs[r]->n_cols < cs[r]->nn_owned) ? 0.0 : 1.0; }
		if ((rc = mgc_allreduce_host(cs, n, v, 2, 1))) return rc;
		for (int r = 0; r < n; ++r) {
			const double gaspect = -v[r][0];
			const bool want_line = cs[r]->prm.mg_smoother == 1 || (cs[r]->prm.mg_smoother < 0 && gaspect >= 3.0);
			cs[r]->mg->line0 = want_line && v[r][1] == 0.0;
		}
	}
	std::vector<std::vector<int32_t>> agg(n), gagg;
	std::vector<int32_t> na(n, 0);
	// extruded structure of the current level on every rank (as in mg_symbolic) and the global decision
	// about the order of the two kinds of coarsening
	std::vector<std::vector<int32_t>> col_of(n), lay_of(n);
	std::vector<int64_t> ncol(n, 0);
	int32_t nlay = 0;
	bool uniform = true;
	double aspect = 0.0;
	{
		std::vector<std::vector<double>> v(n, std::vector<double>(4, 0.0));
		for (int r = 0; r < n; ++r) {
			bp_ctx* c = cs[r];
			col_of[r].assign(c->nn_owned, -1); lay_of[r].assign(c->nn_owned, 0);
			ncol[r] = c->n_cols;
			int32_t lmin = 1 << 30, lmax = 0;
			for (int64_t cc = 0; cc < c->n_cols; ++cc) {
				const int32_t len = c->h_colptr[cc + 1] - c->h_colptr[cc];
				lmin = std::min(lmin, len); lmax = std::max(lmax, len);
				for (int32_t j = c->h_colptr[cc]; j < c->h_colptr[cc + 1]; ++j) { col_of[r][c->h_colnode[j]] = (int32_t)cc; lay_of[r][c->h_colnode[j]] = j - c->h_colptr[cc]; }
			}
			bool all_in = true;
			for (int64_t i = 0; i < c->nn_owned; ++i) if (col_of[r][i] < 0) all_in = false;
			v[r][0] = (double)lmax; v[r][1] = -(double)lmin; v[r][2] = all_in ? 0.0 : 1.0; v[r][3] = -c->aspect;
		}
		if ((rc = mgc_allreduce_host(cs, n, v, 4, 1))) return rc;
		nlay = (int32_t)v[0][0];
		uniform = (v[0][0] == -v[0][1]) && v[0][2] == 0.0;
		aspect = -v[0][3];
	}
	const double aspect_switch = cs[0]->env.mg_horizontal_first;
	for (int guard = 0; guard < 24; ++guard) {
		const int l = (int)cs[0]->mg->lv.size() - 1;
		// ---- A: aggregates of the owned nodes, rank by rank; the KIND of step is the same on every rank:
		//      0 = footprint aggregated layer by layer, 1 = ice columns collapse, 2 = greedy on the node graph
		const int kind = (nlay > 1 && uniform && aspect < aspect_switch) ? 0 : ((nlay > 1 || guard == 0) ? 1 : 2);
		std::vector<std::vector<double>> cnt(n, std::vector<double>(4, 0.0));
		for (int r = 0; r < n; ++r) {
			bp_ctx* c = cs[r];
			bp_mg_level& F = c->mg->lv.back();
			const int64_t nf = F.n;
			const int64_t ncol_before = ncol[r];
			if (kind == 0) {
				std::vector<std::vector<int32_t>> nb(ncol[r]);
				for (int64_t i = 0; i < nf; ++i)
					for (int32_t k = F.h_rowptr[i]; k < F.h_rowptr[i + 1]; ++k) {
						const int32_t j = F.h_col[k];
						if (j < nf && col_of[r][j] != col_of[r][i]) nb[col_of[r][i]].push_back(col_of[r][j]);
					}
				std::vector<int32_t> rp(ncol[r] + 1, 0), ci;
				for (int64_t cc = 0; cc < ncol[r]; ++cc) {
					std::sort(nb[cc].begin(), nb[cc].end());
					nb[cc].erase(std::unique(nb[cc].begin(), nb[cc].end()), nb[cc].end());
					rp[cc + 1] = rp[cc] + (int32_t)nb[cc].size();
				}
				for (int64_t cc = 0; cc < ncol[r]; ++cc) ci.insert(ci.end(), nb[cc].begin(), nb[cc].end());
				std::vector<int32_t> cagg;
				int32_t nca = 0;
				greedy_aggregate(rp, ci, ncol[r], cagg, nca);
				agg[r].assign(nf, -1);
				for (int64_t i = 0; i < nf; ++i) agg[r][i] = cagg[col_of[r][i]] * nlay + lay_of[r][i];
				na[r] = nca * nlay;
				ncol[r] = nca;
				col_of[r].assign(na[r], 0); lay_of[r].assign(na[r], 0);
				for (int32_t I = 0; I < na[r]; ++I) { col_of[r][I] = I / nlay; lay_of[r][I] = I % nlay; }
			} else if (kind == 1) {
				agg[r].assign(nf, -1);
				na[r] = (int32_t)ncol[r];
				for (int64_t i = 0; i < nf; ++i) agg[r][i] = col_of[r][i];
				for (int64_t i = 0; i < nf; ++i) if (agg[r][i] < 0) agg[r][i] = na[r]++;   // a node outside every column stays alone
				ncol[r] = na[r];
				col_of[r].assign(na[r], 0); lay_of[r].assign(na[r], 0);
				for (int32_t I = 0; I < na[r]; ++I) col_of[r][I] = I;
			} else {
				std::vector<int32_t> rp(nf + 1, 0), ci;
				ci.reserve(F.h_col.size());
				for (int64_t i = 0; i < nf; ++i) {
					for (int32_t k = F.h_rowptr[i]; k < F.h_rowptr[i + 1]; ++k) if (F.h_col[k] < nf) ci.push_back(F.h_col[k]);
					rp[i + 1] = (int32_t)ci.size();
				}
				greedy_aggregate(rp, ci, nf, agg[r], na[r]);
				ncol[r] = na[r];
				col_of[r].assign(na[r], 0); lay_of[r].assign(na[r], 0);
				for (int32_t I = 0; I < na[r]; ++I) col_of[r][I] = I;
			}
			cnt[r][0] = (double)nf; cnt[r][1] = (double)na[r]; cnt[r][2] = (double)ncol_before; cnt[r][3] = (double)ncol[r];
		}
		if ((rc = mgc_allreduce_host(cs, n, cnt, 4, 0))) return rc;
		if (kind == 0) aspect *= std::sqrt(cnt[0][2] / std::max(cnt[0][3], 1.0));
		if (kind == 1) nlay = 1;
		const double g_before = cnt[0][0], g_after = cnt[0][1];
		if (guard > 0 && g_after >= g_before) break;              // no rank can coarsen any further
		// ---- B: the aggregate (in the owner's numbering) of every ghost node
		if ((rc = mgc_exchange_ids(cs, n, l, agg, gagg))) return rc;
		// ---- C: ghost nodes and exchange lists of the coarse level, then the level itself
		for (int r = 0; r < n; ++r) {
			bp_ctx* c = cs[r];
			bp_mg* mg = c->mg;
			bp_mg_level& F = mg->lv.back();
			const int64_t nf = F.n, ng = F.n_ghost;
			bp_mg_level H;                                       // carries the coarse halo lists into build_level
			H.h_send_ptr.assign(c->n_nbr + 1, 0); H.h_recv_ptr.assign(c->n_nbr + 1, 0);
			std::vector<int32_t> aggfull(nf + ng, -1);
			std::copy(agg[r].begin(), agg[r].begin() + nf, aggfull.begin());
			int32_t ncg = 0;
			for (int k = 0; k < c->n_nbr; ++k) {
				// ghosts received from neighbour k, in the order they arrive: first appearance numbers the coarse ghost
				std::vector<std::pair<int32_t, int32_t>> seen;       // (owner id, coarse ghost) -- short lists, kept sorted
				for (int64_t g = F.h_recv_ptr[k]; g < F.h_recv_ptr[k + 1]; ++g) {
					const int32_t oid = gagg[r][g];
					auto it = std::lower_bound(seen.begin(), seen.end(), std::make_pair(oid, (int32_t)-1));
					if (it == seen.end() || it->first != oid) {
						it = seen.insert(it, std::make_pair(oid, ncg));
						H.ghost_owner_id.push_back(oid); H.ghost_nbr.push_back(k);
						++ncg;
					}
					aggfull[nf + g] = na[r] + it->second;
				}
				H.h_recv_ptr[k + 1] = ncg;
				// what neighbour k needs from this rank: the aggregates of the nodes sent to it, in the same first-appearance order
				std::vector<int32_t> sent;                            // sorted ids already listed
				for (int64_t q = F.h_send_ptr[k]; q < F.h_send_ptr[k + 1]; ++q) {
					const int32_t a = agg[r][F.h_send_nodes[q]];
					auto it = std::lower_bound(sent.begin(), sent.end(), a);
					if (it == sent.end() || *it != a) { sent.insert(it, a); H.h_send_nodes.push_back(a); }
				}
				H.h_send_ptr[k + 1] = (int64_t)H.h_send_nodes.size();
			}
			if ((rc = build_level(c, mg, aggfull, na[r], ncg, &H))) return rc;
			if ((rc = mgc_attach_halo(c, mg->lv.back()))) return rc;
		}
		if (g_after <= std::max(64.0, (double)cs[0]->env.mg_repl)) break;
	}
	// ---- the global coarsest system: rank r's nodes are [goff[r], goff[r+1])
	{
		std::vector<std::vector<double>> v(n, std::vector<double>(16, 0.0));
		for (int r = 0; r < n; ++r) v[r][cs[r]->mg->my_rank] = (double)cs[r]->mg->lv.back().n;
		if ((rc = mgc_allreduce_host(cs, n, v, 16, 0))) return rc;
		for (int r = 0; r < n; ++r) {
			bp_ctx* c = cs[r];
			bp_mg* mg = c->mg;
			mg->goff.assign(n_ranks + 1, 0);
			for (int q = 0; q < n_ranks; ++q) mg->goff[q + 1] = mg->goff[q] + (int64_t)v[r][q];
			mg->NG = (int)mg->goff[n_ranks];
			mg->nc = 2 * mg->NG;
			bp_mg_level& Lc = mg->lv.back();
			std::vector<int32_t> gcol(Lc.n + Lc.n_ghost);
			for (int64_t i = 0; i < Lc.n; ++i) gcol[i] = (int32_t)(mg->goff[mg->my_rank] + i);
			for (int64_t g = 0; g < Lc.n_ghost; ++g) gcol[Lc.n + g] = (int32_t)(mg->goff[c->nbr_rank[Lc.ghost_nbr[g]]] + Lc.ghost_owner_id[g]);
			UPM(mg->d_gcol, gcol.data(), gcol.size());
			UPM(mg->d_gvec, (const double*)nullptr, (size_t)mg->nc);
			if (mg->NG <= 64) {                            // small enough for the dense inverse
				mg->h_dense.resize((size_t)mg->nc * mg->nc);
				UPM(mg->d_dense, (const double*)nullptr, (size_t)mg->nc * mg->nc);
				UPM(mg->d_gmat, (const double*)nullptr, (size_t)mg->nc * mg->nc);
			}
			if (c->prm.verbose && mg->my_rank == 0) {
				printf("coupled multigrid levels (rank 0, owned+ghost):");
				for (auto& L : mg->lv) printf(" %lld+%lld", (long long)L.n, (long long)L.n_ghost);
				printf(" nodes; global coarsest %d\n", mg->NG);
			}
		}
	}
	if (cs[0]->mg->NG > 64 && (rc = mgc_build_replicated(cs, n))) return rc;
	return BP_OK;
}

// local rows of the coarsest operator into the global dense matrix (row-major, 2 NG x 2 NG)
__global__ void k_mgc_dense_rows(int n, int row0, int ndense, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ col,
                                 const int32_t* __restrict__ gcol, const double* __restrict__ val, double* __restrict__ G)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int b0 = sell_ptr[i >> 5], w = (sell_ptr[(i >> 5) + 1] - b0) >> 5;
	for (int k = 0; k < w; ++k) {
		const size_t slot = (size_t)b0 + 32 * (size_t)k + (i & 31);
		const double* v = val + 4 * slot;
		if (v[0] == 0.0 && v[1] == 0.0 && v[2] == 0.0 && v[3] == 0.0) continue;     // padding slots stay 0
		const int J = gcol[col[slot]];
		double* g = G + (size_t)(2 * (row0 + i)) * ndense + 2 * J;
		g[0] += v[0]; g[1] += v[1]; g[ndense] += v[2]; g[ndense + 1] += v[3];
	}
}
__global__ void k_mgc_rhs_in(int n, int row0, const double2* __restrict__ b, double2* __restrict__ g)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) g[row0 + i] = b[i];
}
__global__ void k_mgc_dense_apply(int n, int row0, int ndense, const double* __restrict__ Ainv, const double* __restrict__ b, double* __restrict__ x)
{
	// one warp per local dof row
	const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (w >= 2 * n) return;
	const double* a = Ainv + (size_t)(2 * row0 + w) * ndense;
	double s = 0.0;
	for (int j = lane; j < ndense; j += 32) s += a[j] * b[j];
	for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	if (lane == 0) x[w] = s;
}

static int mgc_prepare(bp_ctx** cs, int n)
{
	int rc;
	if (!cs[0]->mg && (rc = mgc_symbolic(cs, n))) return rc;
	const int nl = (int)cs[0]->mg->lv.size();
	for (int r = 0; r < n; ++r) {
		bp_ctx* c = cs[r];
		bp_mg* mg = c->mg;
		for (int l = 0; l < nl; ++l) mg->lv[l].M->stream = c->stream;
		for (int l = 0; l + 1 < nl; ++l) {
			bp_mg_level& F = mg->lv[l];
			bp_ctx* C = mg->lv[l + 1].M;
			BP_CUDA(c, cudaMemsetAsync(C->d_val, 0, (size_t)C->n_slots * 4 * sizeof(double), c->stream));
			const int64_t ns = F.h_sell_ptr.back();
			k_mg_galerkin<<<blocks(ns, 256), 256, 0, c->stream>>>(ns, F.d_galmap, F.M->d_val, C->d_val);
			c->launches++;
		}
		{
			mg->f32 = cs[0]->env.mg_fp32;
		}
		for (int l = 0; l < nl; ++l) {
			bp_ctx* M = mg->lv[l].M;
			if (l > 0) bp_launch_pc_setup(M);
			bp_launch_gershgorin(M, SC_LMAX);
			if (mg->f32 && l + 1 < nl && (rc = bp_launch_val_to_float(M))) return rc;
		}
	}
	// one eigenvalue bound per level for all ranks (the smoother must be the same polynomial everywhere)
	{
		std::vector<std::vector<double>> lm(n, std::vector<double>(32, 0.0));
		for (int r = 0; r < n; ++r)
			for (int l = 0; l < nl; ++l) {
				bp_ctx* M = cs[r]->mg->lv[l].M;
				BP_CUDA(cs[r], cudaMemcpyAsync(&lm[r][l], M->d_sc + SC_LMAX, sizeof(double), cudaMemcpyDeviceToHost, cs[r]->stream));
			}
		for (int r = 0; r < n; ++r) BP_CUDA(cs[r], cudaStreamSynchronize(cs[r]->stream));
		if ((rc = mgc_allreduce_host(cs, n, lm, 32, 1))) return rc;
		for (int r = 0; r < n; ++r)
			for (int l = 0; l < nl; ++l) {
				if (!(lm[r][l] > 0.0) || !std::isfinite(lm[r][l])) return bp_fail(cs[0], BP_ERR_STATE, "coupled multigrid: eigenvalue bound of level %d failed (%g)", l, lm[r][l]);
				cs[r]->mg->lv[l].lmax = lm[r][l];
				bp_ctx* M = cs[r]->mg->lv[l].M;
				BP_CUDA(cs[r], cudaMemcpyAsync(M->d_sc + SC_LMAX, &cs[r]->mg->lv[l].lmax, sizeof(double), cudaMemcpyHostToDevice, cs[r]->stream));
			}
	}
	if (cs[0]->mg->R) {
		// replicated tail: every rank adds its rows of the last coupled level to R (zero elsewhere), all-reduce, then the
		// ordinary single-rank numeric set-up below R -- the same numbers on every rank
		double* rv[16];
		for (int r = 0; r < n; ++r) {
			bp_ctx* c = cs[r];
			bp_mg* mg = c->mg;
			bp_ctx* R = mg->R;
			bp_mg_level& Lc = mg->lv.back();
			R->stream = c->stream;
			BP_CUDA(c, cudaMemsetAsync(R->d_val, 0, (size_t)R->n_slots * 4 * sizeof(double), c->stream));
			const int64_t ns = Lc.h_sell_ptr.back();
			if (ns) k_mg_galerkin<<<blocks(ns, 256), 256, 0, c->stream>>>(ns, mg->d_rmap, Lc.M->d_val, R->d_val);
			c->launches++;
			rv[r] = R->d_val;
		}
		if ((rc = bp_comm_allreduce_buf(cs, n, rv, (size_t)cs[0]->mg->R->n_slots * 4, 0))) return rc;
		for (int r = 0; r < n; ++r) {
			bp_ctx* R = cs[r]->mg->R;
			{ const bp_params keep = cs[r]->prm; R->prm = keep; R->prm.pc = BP_PC_MG; R->prm.mg_smoother = 0; }
			bp_launch_pc_setup(R);
			if ((rc = bp_mg_prepare(R))) { cs[0]->err = R->err; return rc; }
			R->mg->no_graph = true;                   // its launches are part of the coupled V-cycle's graph
		}
		return BP_OK;
	}
	// the global coarsest matrix: every rank adds its rows, all-reduce, dense inverse on every rank
	double* gm[16];
	for (int r = 0; r < n; ++r) {
		bp_ctx* c = cs[r];
		bp_mg* mg = c->mg;
		bp_mg_level& Lc = mg->lv.back();
		BP_CUDA(c, cudaMemsetAsync(mg->d_gmat, 0, (size_t)mg->nc * mg->nc * sizeof(double), c->stream));
		if (Lc.n) k_mgc_dense_rows<<<(int)((Lc.n + 127) / 128), 128, 0, c->stream>>>((int)Lc.n, (int)mg->goff[mg->my_rank], mg->nc, Lc.M->d_sell_ptr, Lc.M->d_col, mg->d_gcol, Lc.M->d_val, mg->d_gmat);
		c->launches++;
		gm[r] = mg->d_gmat;
	}
	if ((rc = bp_comm_allreduce_buf(cs, n, gm, (size_t)cs[0]->mg->nc * cs[0]->mg->nc, 0))) return rc;
	for (int r = 0; r < n; ++r) {
		bp_ctx* c = cs[r];
		bp_mg* mg = c->mg;
		const int nd = mg->nc;
		if (r > 0 && n > 1) {                      // in-process group: one inversion serves every context
			BP_CUDA(c, cudaMemcpyAsync(mg->d_dense, cs[0]->mg->d_dense, (size_t)nd * nd * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
			continue;
		}
		std::vector<double> A((size_t)nd * nd);
		BP_CUDA(c, cudaMemcpyAsync(A.data(), mg->d_gmat, A.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		BP_CUDA(c, cudaStreamSynchronize(c->stream));
		std::vector<double>& Inv = mg->h_dense;
		std::fill(Inv.begin(), Inv.end(), 0.0);
		for (int i = 0; i < nd; ++i) Inv[(size_t)i * nd + i] = 1.0;
		for (int col = 0; col < nd; ++col) {
			int piv = col;
			for (int q = col + 1; q < nd; ++q) if (std::fabs(A[(size_t)q * nd + col]) > std::fabs(A[(size_t)piv * nd + col])) piv = q;
			if (A[(size_t)piv * nd + col] == 0.0) return bp_fail(c, BP_ERR_STATE, "coupled multigrid: the global coarsest matrix is singular");
			if (piv != col) for (int j = 0; j < nd; ++j) { std::swap(A[(size_t)piv * nd + j], A[(size_t)col * nd + j]); std::swap(Inv[(size_t)piv * nd + j], Inv[(size_t)col * nd + j]); }
			const double ip = 1.0 / A[(size_t)col * nd + col];
			for (int j = 0; j < nd; ++j) { A[(size_t)col * nd + j] *= ip; Inv[(size_t)col * nd + j] *= ip; }
			for (int q = 0; q < nd; ++q) {
				if (q == col) continue;
				const double f = A[(size_t)q * nd + col];
				if (f == 0.0) continue;
				for (int j = 0; j < nd; ++j) { A[(size_t)q * nd + j] -= f * A[(size_t)col * nd + j]; Inv[(size_t)q * nd + j] -= f * Inv[(size_t)col * nd + j]; }
			}
		}
		BP_CUDA(c, cudaMemcpyAsync(mg->d_dense, Inv.data(), Inv.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	}
	return BP_OK;
}

// Chebyshev smoothing of level l on all ranks, a halo exchange of the iterate before every step that reads it.
// zid[r]: which of the two iterate buffers (2 = d_z, 3 = d_z2) holds the current iterate; updated.
static int mgc_smooth(bp_ctx** cs, int n, int l, int deg, double ratio, double** rhs, int* zid, bool zero_init)
{
	bp_ctx* Ls[16];
	mgc_level_ctxs(cs, n, l, Ls);
	const double a = 1.0 / ratio, theta = 0.5 * (1.0 + a), delta = 0.5 * (1.0 - a), sigma = theta / delta;
	double rho = 1.0 / sigma;
	int rc;
	const bool line = (l == 0) && cs[0]->mg->line0;
	if (line) {
		// Chebyshev on the column splitting J = T + N (col_cheb_smooth above): lmax(T^-1 J) <= columns a tet touches
		const double b = (double)cs[0]->col_touch_max, a2 = b / ratio, th = 0.5 * (b + a2), de = 0.5 * (b - a2), sg = th / de;
		double rh = 1.0 / sg;
		auto lstep = [&](double c1, double c2) -> int {
			if ((rc = bp_comm_halo(Ls, n, zid[0]))) return rc;
			for (int r = 0; r < n; ++r) {
				bp_ctx* M = Ls[r];
				double* zo = (zid[r] == 2) ? M->d_z : M->d_z2;
				double* zn = (zid[r] == 2) ? M->d_z2 : M->d_z;
				bp_launch_residual(M, rhs[r], zo, M->d_io, false);
				bp_launch_col_cheb(M, c1, c2, M->d_io, zo, zn);
				zid[r] = (zid[r] == 2) ? 3 : 2;
			}
			return BP_OK;
		};
		if (zero_init) { for (int r = 0; r < n; ++r) { bp_launch_col_cheb(Ls[r], 0.0, 1.0 / th, rhs[r], nullptr, Ls[r]->d_z); zid[r] = 2; } }
		else if ((rc = lstep(0.0, 1.0 / th))) return rc;
		for (int k = 1; k < deg; ++k) {
			const double rn = 1.0 / (2.0 * sg - rh);
			if ((rc = lstep(rn * rh, 2.0 * rn / de))) return rc;
			rh = rn;
		}
		return BP_OK;
	}
	auto step = [&](double c1, double c2) -> int {
		if ((rc = bp_comm_halo(Ls, n, zid[0]))) return rc;
		for (int r = 0; r < n; ++r) {
			bp_ctx* M = Ls[r];
			double* zo = (zid[r] == 2) ? M->d_z : M->d_z2;
			double* zn = (zid[r] == 2) ? M->d_z2 : M->d_z;
			bp_launch_cheb_step(M, c1, c2, rhs[r], zo, zn, M->d_sc + SC_LMAX, cs[r]->mg->f32);
			zid[r] = (zid[r] == 2) ? 3 : 2;
		}
		return BP_OK;
	};
	if (zero_init) {
		for (int r = 0; r < n; ++r) { bp_launch_cheb_first(Ls[r], 1.0 / theta, rhs[r], Ls[r]->d_z, Ls[r]->d_sc + SC_LMAX); zid[r] = 2; }
	} else if ((rc = step(0.0, 1.0 / theta))) return rc;
	for (int k = 1; k < deg; ++k) {
		const double rho_n = 1.0 / (2.0 * sigma - rho);
		if ((rc = step(rho_n * rho, 2.0 * rho_n / delta))) return rc;
		rho = rho_n;
	}
	return BP_OK;
}

// V-cycle on all ranks in lock-step; the correction of level l ends up in buffer zid[r] of its context
static int mgc_vcycle(bp_ctx** cs, int n, int l, double** rhs, int deg, double ratio, int* zid)
{
	int rc;
	const int nl = (int)cs[0]->mg->lv.size();
	bp_ctx* Ls[16];
	mgc_level_ctxs(cs, n, l, Ls);
	if (l == nl - 1 && cs[0]->mg->R) {
		// replicated tail: all-reduce the right-hand side of this level, one V-cycle of R's hierarchy on every rank, own part back
		double* gv[16];
		for (int r = 0; r < n; ++r) {
			bp_mg* mg = cs[r]->mg;
			bp_ctx* R = mg->R;
			const int64_t nloc = mg->lv[l].n;
			BP_CUDA(cs[r], cudaMemsetAsync(R->d_r, 0, (size_t)mg->nc * sizeof(double), cs[r]->stream));
			if (nloc) k_mgc_rhs_in<<<(int)((nloc + 127) / 128), 128, 0, cs[r]->stream>>>((int)nloc, (int)mg->goff[mg->my_rank], (const double2*)rhs[r], (double2*)R->d_r);
			cs[r]->launches++;
			gv[r] = R->d_r;
		}
		if ((rc = bp_comm_allreduce_buf(cs, n, gv, (size_t)cs[0]->mg->nc, 0))) return rc;
		for (int r = 0; r < n; ++r) {
			bp_mg* mg = cs[r]->mg;
			bp_ctx* R = mg->R;
			const int64_t nloc = mg->lv[l].n, l0 = R->launches;
			R->stream = cs[r]->stream;
			R->outer_capture = cs[0]->outer_capture;
			for (size_t q = 1; q < R->mg->lv.size(); ++q) R->mg->lv[q].M->stream = cs[r]->stream;
			if ((rc = bp_mg_apply(R))) { cs[0]->err = R->err; return rc; }
			cs[r]->launches += R->launches - l0;
			if (nloc) BP_CUDA(cs[r], cudaMemcpyAsync(Ls[r]->d_z, R->d_z + 2 * mg->goff[mg->my_rank], (size_t)2 * nloc * sizeof(double), cudaMemcpyDeviceToDevice, cs[r]->stream));
			zid[r] = 2;
		}
		return BP_OK;
	}
	if (l == nl - 1) {
		// gather the right-hand side, apply the rows of the dense inverse that belong to this rank
		double* gv[16];
		for (int r = 0; r < n; ++r) {
			bp_mg* mg = cs[r]->mg;
			const int64_t nloc = mg->lv[l].n;
			BP_CUDA(cs[r], cudaMemsetAsync(mg->d_gvec, 0, (size_t)mg->nc * sizeof(double), cs[r]->stream));
			if (nloc) k_mgc_rhs_in<<<(int)((nloc + 127) / 128), 128, 0, cs[r]->stream>>>((int)nloc, (int)mg->goff[mg->my_rank], (const double2*)rhs[r], (double2*)mg->d_gvec);
			cs[r]->launches++;
			gv[r] = mg->d_gvec;
		}
		if ((rc = bp_comm_allreduce_buf(cs, n, gv, (size_t)cs[0]->mg->nc, 0))) return rc;
		for (int r = 0; r < n; ++r) {
			bp_mg* mg = cs[r]->mg;
			const int64_t nloc = mg->lv[l].n;
			if (nloc) k_mgc_dense_apply<<<(int)((2 * nloc * 32 + 255) / 256), 256, 0, cs[r]->stream>>>((int)nloc, (int)mg->goff[mg->my_rank], mg->nc, mg->d_dense, mg->d_gvec, Ls[r]->d_z);
			cs[r]->launches++;
			zid[r] = 2;
		}
		return BP_OK;
	}
	if ((rc = mgc_smooth(cs, n, l, deg, ratio, rhs, zid, true))) return rc;
	// residual (needs the ghosts of the iterate), restriction.  Every level exchanges: an operator whose ghost
	// columns see stale zeros is no longer symmetric positive definite as a whole and CG breaks down (tried).
	if ((rc = bp_comm_halo(Ls, n, zid[0]))) return rc;
	double* crhs[16];
	for (int r = 0; r < n; ++r) {
		bp_ctx* M = Ls[r];
		bp_mg_level& L = cs[r]->mg->lv[l];
		bp_ctx* C = cs[r]->mg->lv[l + 1].M;
		bp_launch_residual(M, rhs[r], (zid[r] == 2) ? M->d_z : M->d_z2, M->d_io, cs[r]->mg->f32 && !(l == 0 && cs[r]->mg->line0));
		k_mg_restrict<<<blocks(C->nn_owned, 128), 128, 0, cs[r]->stream>>>((int)C->nn_owned, L.d_aggptr, L.d_aggidx, (const double2*)M->d_io, (double2*)C->d_F);
		cs[r]->launches++;
		crhs[r] = C->d_F;
	}
	int czid[16];
	if ((rc = mgc_vcycle(cs, n, l + 1, crhs, deg, ratio, czid))) return rc;
	for (int r = 0; r < n; ++r) {
		bp_ctx* M = Ls[r];
		bp_mg_level& L = cs[r]->mg->lv[l];
		bp_ctx* C = cs[r]->mg->lv[l + 1].M;
		const double alpha_def = (cs[r]->prm.mg_coarse_scale > 0.0 && cs[r]->prm.mg_coarse_scale <= 2.0) ? cs[r]->prm.mg_coarse_scale : 1.8;
		const bp_env& E = cs[r]->env;
		const double alpha = (l == 0 ? E.has_alpha0 : E.has_alpha) ? (l == 0 ? E.mg_alpha0 : E.mg_alpha) : alpha_def;
		k_mg_prolong_add<<<blocks(L.n, 256), 256, 0, cs[r]->stream>>>((int)L.n, L.d_agg, (const double2*)((czid[r] == 2) ? C->d_z : C->d_z2),
			(double2*)((zid[r] == 2) ? M->d_z : M->d_z2), alpha);
		cs[r]->launches++;
	}
	return mgc_smooth(cs, n, l, deg, ratio, rhs, zid, false);
}

static int mgc_apply_direct(bp_ctx** cs, int n, int deg, double ratio)
{
	int rc;
	double* rhs[16];
	int zid[16];
	for (int r = 0; r < n; ++r) { rhs[r] = cs[r]->d_r; for (size_t l = 1; l < cs[r]->mg->lv.size(); ++l) { cs[r]->mg->lv[l].M->launches = 0; cs[r]->mg->lv[l].M->stream = cs[r]->stream; } }
	if ((rc = mgc_vcycle(cs, n, 0, rhs, deg, ratio, zid))) return rc;
	for (int r = 0; r < n; ++r) {
		bp_ctx* c = cs[r];
		if (zid[r] != 2) BP_CUDA(c, cudaMemcpyAsync(c->d_z, c->d_z2, (size_t)2 * c->nn_owned * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
		for (size_t l = 1; l < c->mg->lv.size(); ++l) c->launches += c->mg->lv[l].M->launches;
	}
	return BP_OK;
}

// The V-cycle with its exchanges (NCCL send/recv and all-reduce, or the device copies of an in-process
// group) is a fixed sequence: the first application runs directly (NCCL sets up its peer connections
// there), the second is captured into a CUDA graph, all later ones replay it.
static int mgc_apply(bp_ctx** cs, int n)
{
	int rc;
	bp_mg* mg = cs[0]->mg;                  // the graph of an in-process group lives in its first context
	const int deg = std::max(cs[0]->prm.pc_degree, 1);
	const double ratio = std::max(cs[0]->prm.pc_eig_ratio, 1.5);
	if (cs[0]->outer_capture) return mgc_apply_direct(cs, n, deg, ratio);      // into the caller's capture, kernel by kernel
	const bool want_graph = cs[0]->env.mg_coupled_graph && !mg->no_graph;
	if (mg->graph && (mg->g_r != cs[0]->d_r || mg->g_z != cs[0]->d_z || mg->g_deg != deg || mg->g_ratio != ratio || mg->g_alpha != cs[0]->prm.mg_coarse_scale || mg->g_f32 != mg->f32)) {
		cudaGraphExecDestroy(mg->graph);
		mg->graph = nullptr;
	}
	if (!want_graph || (!mg->graph && mg->g_warm < 1)) {
		mg->g_warm++;
		return mgc_apply_direct(cs, n, deg, ratio);
	}
	if (!mg->graph) {
		std::vector<int64_t> l0(n);
		for (int r = 0; r < n; ++r) l0[r] = cs[r]->launches;
		cudaGraph_t g = nullptr;
		cudaError_t eb = cudaStreamBeginCapture(cs[0]->stream, cudaStreamCaptureModeThreadLocal);
		if (eb == cudaSuccess) {
			rc = mgc_apply_direct(cs, n, deg, ratio);
			eb = cudaStreamEndCapture(cs[0]->stream, &g);
			if (rc == BP_OK && eb == cudaSuccess) eb = cudaGraphInstantiate(&mg->graph, g, 0);
			if (g) cudaGraphDestroy(g);
			if (rc != BP_OK || eb != cudaSuccess) { mg->graph = nullptr; }
		}
		if (!mg->graph) {                       // not capturable here: keep launching directly
			cudaGetLastError();
			mg->no_graph = true;
			for (int r = 0; r < n; ++r) cs[r]->launches = l0[r];
			return mgc_apply_direct(cs, n, deg, ratio);
		}
		mg->g_launches = 0;
		for (int r = 0; r < n; ++r) { mg->g_launches += cs[r]->launches - l0[r]; cs[r]->launches = l0[r]; }
		mg->g_r = cs[0]->d_r; mg->g_z = cs[0]->d_z; mg->g_deg = deg; mg->g_ratio = ratio; mg->g_alpha = cs[0]->prm.mg_coarse_scale; mg->g_f32 = mg->f32;
	}
	BP_CUDA(cs[0], cudaGraphLaunch(mg->graph, cs[0]->stream));
	cs[0]->launches += mg->g_launches;
	return BP_OK;
}

// entry points for the partitioned path (bp_api.cu): coupled hierarchy unless BP_MG_COUPLED=0
bool bp_mg_use_coupled(bp_ctx** cs, int n)
{
	if (!cs[0]->env.mg_coupled) return false;
	if (cs[0]->mg) return cs[0]->mg->coupled;
	const bool partitioned = (n > 1) || cs[0]->n_nbr > 0;
	if (!partitioned) return false;
	for (int r = 0; r < n; ++r) if (cs[r]->n_cols <= 0) return false;
	return true;
}
// The V-cycle's launch sequence is final -- its graph exists, or it is launched directly for good -- so a stream capture
// around it (the CG iteration graph) will not run into a capture of its own.  id changes when the graph is rebuilt.
bool bp_mg_settled(bp_ctx** cs, int n, const void** id)
{
	(void)n;
	bp_mg* mg = cs[0]->mg;
	if (!mg) return false;
	*id = (const void*)mg->graph;
	if (mg->coupled) return mg->graph != nullptr || mg->no_graph || !cs[0]->env.mg_coupled_graph;
	return mg->graph != nullptr || mg->no_graph;
}
int bp_mg_prepare_group(bp_ctx** cs, int n) { return mgc_prepare(cs, n); }
int bp_mg_apply_group(bp_ctx** cs, int n) { return mgc_apply(cs, n); }

// ------------------------------------------------------------------ numeric set-up (every Newton step)
int bp_mg_prepare(bp_ctx* c)
{
	int rc;
	if (!c->mg && (rc = mg_symbolic(c))) return rc;
	bp_mg* mg = c->mg;
	const int nl = (int)mg->lv.size();
	for (int l = 0; l < nl; ++l) mg->lv[l].M->stream = c->stream;
	// ghosts of the fine iterates are zero: the hierarchy is rank-local
	if (c->nn > c->nn_owned) {
		const size_t off = (size_t)2 * c->nn_owned, nb = (size_t)2 * (c->nn - c->nn_owned) * sizeof(double);
		BP_CUDA(c, cudaMemsetAsync(c->d_z + off, 0, nb, c->stream));
		BP_CUDA(c, cudaMemsetAsync(c->d_z2 + off, 0, nb, c->stream));
	}
	for (int l = 0; l + 1 < nl; ++l) {
		bp_mg_level& F = mg->lv[l];
		bp_ctx* C = mg->lv[l + 1].M;
		BP_CUDA(c, cudaMemsetAsync(C->d_val, 0, (size_t)C->n_slots * 4 * sizeof(double), c->stream));
		const int64_t ns = F.h_sell_ptr.back();
		k_mg_galerkin<<<blocks(ns, 256), 256, 0, c->stream>>>(ns, F.d_galmap, F.M->d_val, C->d_val);
		c->launches++;
	}
	{
		mg->f32 = c->env.mg_fp32;
	}
	for (int l = 0; l < nl; ++l) {
		bp_ctx* M = mg->lv[l].M;
		if (l > 0) bp_launch_pc_setup(M);             // level 0: done by the caller
		bp_launch_gershgorin(M, SC_LMAX);
		// FP32 copy of the level matrix for the smoother (the line smoother of level 0 keeps FP64)
		const bool want_line = c->prm.mg_smoother == 1 || (c->prm.mg_smoother < 0 && c->aspect >= 3.0);
		const bool line0 = (l == 0) && want_line && c->max_layers <= 64 && c->n_cols < c->nn_owned;
		if (mg->f32 && !line0 && l + 1 < nl && (rc = bp_launch_val_to_float(M))) return rc;
	}
	for (int l = 0; l < nl; ++l) {
		bp_ctx* M = mg->lv[l].M;
		BP_CUDA(c, cudaMemcpyAsync(M->h_sc, M->d_sc, SC_COUNT * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	}
	// coarsest matrix -> host, dense inverse
	bp_mg_level& Lc = mg->lv.back();
	mg->h_val.resize((size_t)Lc.M->n_slots * 4);
	BP_CUDA(c, cudaMemcpyAsync(mg->h_val.data(), Lc.M->d_val, mg->h_val.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	for (int l = 0; l < nl; ++l) {
		mg->lv[l].lmax = mg->lv[l].M->h_sc[SC_LMAX];
		if (!(mg->lv[l].lmax > 0.0) || !std::isfinite(mg->lv[l].lmax))
			return bp_fail(c, BP_ERR_STATE, "multigrid: eigenvalue bound of level %d failed (%g)", l, mg->lv[l].lmax);
	}
	const int n = mg->nc;
	std::vector<double> A((size_t)n * n, 0.0);
	for (int64_t I = 0; I < Lc.n; ++I)
		for (int32_t k = Lc.h_rowptr[I]; k < Lc.h_rowptr[I + 1]; ++k) {
			const int32_t J = Lc.h_col[k];
			const double* v = mg->h_val.data() + 4 * (size_t)(Lc.h_sell_ptr[I >> 5] + 32 * (k - Lc.h_rowptr[I]) + (I & 31));
			A[(size_t)(2 * I) * n + 2 * J] = v[0]; A[(size_t)(2 * I) * n + 2 * J + 1] = v[1];
			A[(size_t)(2 * I + 1) * n + 2 * J] = v[2]; A[(size_t)(2 * I + 1) * n + 2 * J + 1] = v[3];
		}
	// Gauss-Jordan with partial pivoting
	std::vector<double>& Inv = mg->h_dense;
	std::fill(Inv.begin(), Inv.end(), 0.0);
	for (int i = 0; i < n; ++i) Inv[(size_t)i * n + i] = 1.0;
	for (int col = 0; col < n; ++col) {
		int piv = col;
		for (int r = col + 1; r < n; ++r) if (std::fabs(A[(size_t)r * n + col]) > std::fabs(A[(size_t)piv * n + col])) piv = r;
		if (A[(size_t)piv * n + col] == 0.0) return bp_fail(c, BP_ERR_STATE, "multigrid: coarsest matrix is singular");
		if (piv != col) for (int j = 0; j < n; ++j) { std::swap(A[(size_t)piv * n + j], A[(size_t)col * n + j]); std::swap(Inv[(size_t)piv * n + j], Inv[(size_t)col * n + j]); }
		const double ip = 1.0 / A[(size_t)col * n + col];
		for (int j = 0; j < n; ++j) { A[(size_t)col * n + j] *= ip; Inv[(size_t)col * n + j] *= ip; }
		for (int r = 0; r < n; ++r) {
			if (r == col) continue;
			const double f = A[(size_t)r * n + col];
			if (f == 0.0) continue;
			for (int j = 0; j < n; ++j) { A[(size_t)r * n + j] -= f * A[(size_t)col * n + j]; Inv[(size_t)r * n + j] -= f * Inv[(size_t)col * n + j]; }
		}
	}
	BP_CUDA(c, cudaMemcpyAsync(mg->d_dense, Inv.data(), Inv.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	return BP_OK;
}

// ------------------------------------------------------------------ V-cycle
// Chebyshev smoothing for the interval [lmax / ratio, lmax]; lmax is read on the device
// (d_sc[SC_LMAX] of the level), so the launch sequence does not depend on the matrix values
// and can be replayed as a graph.  Iterates ping-pong between za and zb; returns the buffer
// that holds the result.
static double* cheb_smooth(bp_ctx* M, int deg, double ratio, const double* r, double* za, double* zb, bool zero_init, bool f32)
{
	const double a = 1.0 / ratio, theta = 0.5 * (1.0 + a), delta = 0.5 * (1.0 - a), sigma = theta / delta;
	const double* lm = M->d_sc + SC_LMAX;
	double rho = 1.0 / sigma;
	if (zero_init) bp_launch_cheb_first(M, 1.0 / theta, r, za, lm);
	else { bp_launch_cheb_step(M, 0.0, 1.0 / theta, r, za, zb, lm, f32); std::swap(za, zb); }
	for (int k = 1; k < deg; ++k) {
		const double rho_n = 1.0 / (2.0 * sigma - rho);
		bp_launch_cheb_step(M, rho_n * rho, 2.0 * rho_n / delta, r, za, zb, lm, f32);
		std::swap(za, zb);
		rho = rho_n;
	}
	return za;
}

// level-0 line smoother: Chebyshev on the column splitting J = T + N, T = the exact coupling
// inside every ice column.  lmax(T^-1 J) <= (columns a tet touches) = 3 on extruded meshes
// (J is a sum of positive semi-definite element matrices), a fixed number: graph-safe.
static double* col_cheb_smooth(bp_ctx* c, int deg, double ratio, const double* rhs, double* za, double* zb, bool zero_init, int64_t n)
{
	const double b = (double)c->col_touch_max, a = b / ratio, theta = 0.5 * (b + a), delta = 0.5 * (b - a), sigma = theta / delta;
	double rho = 1.0 / sigma;
	if (zero_init) bp_launch_col_cheb(c, 0.0, 1.0 / theta, rhs, nullptr, za);
	else {
		k_mg_residual<<<blocks(n, 256), 256, 0, c->stream>>>((int)n, c->d_sell_ptr, c->d_col, (const double2*)c->d_val, (const double2*)rhs, (const double2*)za, (double2*)c->d_io);
		c->launches++;
		bp_launch_col_cheb(c, 0.0, 1.0 / theta, c->d_io, za, zb);
		std::swap(za, zb);
	}
	for (int k = 1; k < deg; ++k) {
		const double rho_n = 1.0 / (2.0 * sigma - rho);
		k_mg_residual<<<blocks(n, 256), 256, 0, c->stream>>>((int)n, c->d_sell_ptr, c->d_col, (const double2*)c->d_val, (const double2*)rhs, (const double2*)za, (double2*)c->d_io);
		c->launches++;
		bp_launch_col_cheb(c, rho_n * rho, 2.0 * rho_n / delta, c->d_io, za, zb);
		std::swap(za, zb);
		rho = rho_n;
	}
	return za;
}

// returns the buffer of level l that holds the correction
static double* vcycle(bp_ctx* c, bp_mg* mg, int l, const double* rhs, int deg, double ratio)
{
	bp_mg_level& L = mg->lv[l];
	bp_ctx* M = L.M;
	const int nl = (int)mg->lv.size();
	if (l == nl - 1) {
		k_mg_dense<<<(mg->nc + 127) / 128, 128, 0, c->stream>>>(mg->nc, mg->d_dense, rhs, M->d_z);
		c->launches++;
		return M->d_z;
	}
	{
		const int tail_nodes = c->env.mg_tail;
		const bool v_only = std::max(1, c->env.mg_gamma ? c->env.mg_gamma : mg->gamma) == 1;
		if (l > 0 && v_only && L.n <= tail_nodes && nl - l <= MG_TAIL_MAX && rhs == M->d_F) {
			// this level and everything below it: one launch
			TailArgs A;
			memset(&A, 0, sizeof(A));
			A.nl = nl - l; A.deg = deg; A.ratio = ratio;
			A.alpha = c->env.has_alpha ? c->env.mg_alpha : ((c->prm.mg_coarse_scale > 0.0 && c->prm.mg_coarse_scale <= 2.0) ? c->prm.mg_coarse_scale : (c->nn > c->nn_owned ? 1.0 : 1.8));
			A.dense = mg->d_dense; A.ndense = mg->nc;
			for (int k = l; k < nl; ++k) {
				bp_mg_level& Lk = mg->lv[k];
				bp_ctx* Mk = Lk.M;
				TailLevel& T = A.lv[k - l];
				T.n = (int)Lk.n; T.nc = (k + 1 < nl) ? (int)mg->lv[k + 1].n : 0;
				T.sell_ptr = Mk->d_sell_ptr; T.col = Mk->d_col;
				T.valf = (mg->f32 && k + 1 < nl) ? (const float4*)Mk->d_valf : nullptr;
				T.val = (const double2*)Mk->d_val; T.dinv = Mk->d_dinv; T.lmax = Mk->d_sc + SC_LMAX;
				T.F = (double2*)Mk->d_F; T.z = (double2*)Mk->d_z; T.z2 = (double2*)Mk->d_z2; T.cd = (double2*)Mk->d_cd; T.io = (double2*)Mk->d_io;
				T.agg = Lk.d_agg; T.aggptr = Lk.d_aggptr; T.aggidx = Lk.d_aggidx;
			}
			k_mg_tail<<<1, 1024, 0, c->stream>>>(A);
			c->launches++;
			return M->d_z;
		}
	}
	// line smoother when asked for, or (auto) when the elements are flat: dx >= 3 dz
	const bool want_line = c->prm.mg_smoother == 1 || (c->prm.mg_smoother < 0 && c->aspect >= 3.0);
	const bool line = (l == 0) && want_line && c->max_layers <= 64 && c->n_cols < c->nn_owned;
	double* z = line ? col_cheb_smooth(c, deg, ratio, rhs, M->d_z, M->d_z2, true, L.n)
	                 : cheb_smooth(M, deg, ratio, rhs, M->d_z, M->d_z2, true, mg->f32);
	bp_ctx* C = mg->lv[l + 1].M;
	// gamma coarse-grid visits per level below the finest (1 = V-cycle, 2 = W-cycle): plain
	// aggregation has a stiff coarse operator, a second visit of the (tiny) coarser levels is cheap
	const int gamma = (l == 0) ? 1 : std::max(1, c->env.mg_gamma ? c->env.mg_gamma : mg->gamma);
	const bool has_ev = (l == 0) ? c->env.has_alpha0 : c->env.has_alpha;
	const double ev_alpha = (l == 0) ? c->env.mg_alpha0 : c->env.mg_alpha;
	const double alpha_def = (c->prm.mg_coarse_scale > 0.0 && c->prm.mg_coarse_scale <= 2.0) ? c->prm.mg_coarse_scale : (c->nn > c->nn_owned ? 1.0 : 1.8);
	const double alpha = has_ev ? ev_alpha : alpha_def;
	for (int g = 0; g < gamma; ++g) {
		double* other = (z == M->d_z) ? M->d_z2 : M->d_z;
		if (line) {
			k_mg_residual<<<blocks(L.n, 256), 256, 0, c->stream>>>((int)L.n, M->d_sell_ptr, M->d_col, (const double2*)M->d_val,
				(const double2*)rhs, (const double2*)z, (double2*)M->d_io);
			c->launches++;
		} else bp_launch_residual(M, rhs, z, M->d_io, mg->f32);
		k_mg_restrict<<<blocks(C->nn_owned, 128), 128, 0, c->stream>>>((int)C->nn_owned, L.d_aggptr, L.d_aggidx, (const double2*)M->d_io, (double2*)C->d_F);
		c->launches += 1;
		const double* xc = vcycle(c, mg, l + 1, C->d_F, deg, ratio);
		// over-relaxed coarse correction (piecewise-constant prolongation under-corrects)
		k_mg_prolong_add<<<blocks(L.n, 256), 256, 0, c->stream>>>((int)L.n, L.d_agg, (const double2*)xc, (double2*)z, alpha);
		c->launches++;
		z = line ? col_cheb_smooth(c, deg, ratio, rhs, z, other, false, L.n) : cheb_smooth(M, deg, ratio, rhs, z, other, false, mg->f32);
	}
	return z;
}

// z = M^-1 r with r = c->d_r, z = c->d_z
int bp_mg_apply(bp_ctx* c)
{
	if (!c->mg) return bp_fail(c, BP_ERR_STATE, "multigrid preconditioner not prepared");
	bp_mg* mg = c->mg;
	const int deg = std::max(c->prm.pc_degree, 1);
	const double ratio = std::max(c->prm.pc_eig_ratio, 1.5);
	if (mg->graph && (mg->g_r != c->d_r || mg->g_z != c->d_z || mg->g_z2 != c->d_z2 || mg->g_deg != deg || mg->g_ratio != ratio || mg->g_smoother != c->prm.mg_smoother || mg->g_alpha != c->prm.mg_coarse_scale || mg->g_f32 != mg->f32)) {
		cudaGraphExecDestroy(mg->graph);
		mg->graph = nullptr;
	}
	if (c->outer_capture) {
		// the caller is capturing a graph of its own (the CG iteration): the V-cycle goes into it kernel by kernel
		for (size_t l = 1; l < mg->lv.size(); ++l) { mg->lv[l].M->launches = 0; mg->lv[l].M->stream = c->stream; }
		double* out = vcycle(c, mg, 0, c->d_r, deg, ratio);
		if (out != c->d_z) BP_CUDA(c, cudaMemcpyAsync(c->d_z, out, (size_t)2 * c->nn_owned * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
		for (size_t l = 1; l < mg->lv.size(); ++l) c->launches += mg->lv[l].M->launches;
		return BP_OK;
	}
	if (!mg->graph) {
		int64_t l0 = c->launches;
		for (size_t l = 1; l < mg->lv.size(); ++l) { mg->lv[l].M->launches = 0; mg->lv[l].M->stream = c->stream; }
		cudaGraph_t g = nullptr;
		// the legacy default stream (e.g. torch's default current stream handed over with bp_set_stream)
		// cannot be captured: launch the V-cycle directly there
		const cudaError_t eb = mg->no_graph ? cudaErrorStreamCaptureUnsupported : cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
		if (eb != cudaSuccess) {
			cudaGetLastError();
			mg->no_graph = true;
			double* out = vcycle(c, mg, 0, c->d_r, deg, ratio);
			if (out != c->d_z) BP_CUDA(c, cudaMemcpyAsync(c->d_z, out, (size_t)2 * c->nn_owned * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
			for (size_t l = 1; l < mg->lv.size(); ++l) c->launches += mg->lv[l].M->launches;
			return BP_OK;
		}
		double* out = vcycle(c, mg, 0, c->d_r, deg, ratio);
		if (out != c->d_z) cudaMemcpyAsync(c->d_z, out, (size_t)2 * c->nn_owned * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
		BP_CUDA(c, cudaStreamEndCapture(c->stream, &g));
		cudaError_t e = cudaGraphInstantiate(&mg->graph, g, 0);
		cudaGraphDestroy(g);
		BP_CUDA(c, e);
		mg->g_launches = c->launches - l0;
		for (size_t l = 1; l < mg->lv.size(); ++l) mg->g_launches += mg->lv[l].M->launches;
		c->launches = l0;
		mg->g_r = c->d_r; mg->g_z = c->d_z; mg->g_z2 = c->d_z2; mg->g_deg = deg; mg->g_ratio = ratio; mg->g_smoother = c->prm.mg_smoother; mg->g_alpha = c->prm.mg_coarse_scale; mg->g_f32 = mg->f32;
	}
	BP_CUDA(c, cudaGraphLaunch(mg->graph, c->stream));
	c->launches += mg->g_launches;
	return BP_OK;
}

void bp_mg_free(bp_ctx* c)
{
	if (c->cg_graph) { cudaGraphExecDestroy(c->cg_graph); c->cg_graph = nullptr; }      // it replays this hierarchy's buffers
	if (!c->mg) return;
	for (size_t l = 0; l < c->mg->lv.size(); ++l) {
		bp_mg_level& L = c->mg->lv[l];
		void* t[] = { L.d_agg, L.d_galmap, L.d_aggptr, L.d_aggidx };
		for (void* p : t) if (p) cudaFree(p);
		if (l > 0 && L.M) {
			bp_ctx* M = L.M;
			bp_comm_p2p_detach(M);
			void* ptrs[] = { M->d_send_nodes, M->d_sendbuf, M->d_sell_ptr, M->d_col, M->d_diag, M->d_val, M->d_valf, M->d_dinv, M->d_F, M->d_z, M->d_z2, M->d_cd, M->d_io, M->d_sc, M->d_part, M->d_counter };
			for (void* p : ptrs) if (p) cudaFree(p);
			if (M->h_sc) cudaFreeHost(M->h_sc);
			delete M;
		}
	}
	if (c->mg->R) {                                   // the replicated tail: a context with a hierarchy of its own
		bp_ctx* M = c->mg->R;
		bp_mg_free(M);
		void* ptrs[] = { M->d_sell_ptr, M->d_col, M->d_diag, M->d_val, M->d_valf, M->d_dinv, M->d_F, M->d_z, M->d_z2, M->d_cd, M->d_io, M->d_sc, M->d_part, M->d_counter, M->d_r };
		for (void* p : ptrs) if (p) cudaFree(p);
		if (M->h_sc) cudaFreeHost(M->h_sc);
		delete M;
	}
	if (c->mg->graph) cudaGraphExecDestroy(c->mg->graph);
	if (c->mg->d_dense) cudaFree(c->mg->d_dense);
	{ void* t[] = { c->mg->d_gmat, c->mg->d_gvec, c->mg->d_xs, c->mg->d_gcol, c->mg->d_rmap }; for (void* p : t) if (p) cudaFree(p); }
	delete c->mg;
	c->mg = nullptr;
}
