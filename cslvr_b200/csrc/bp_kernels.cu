// bp_kernels.cu -- sm_100a FP64 kernels of the Blatter-Pattyn momentum path.
//
// Replaces the FFC-generated tabulate_tensor + dolfin Assembler (residual:
// cslvr/momentumbp.py:481-500 / 372-375, Jacobian: cslvr/physics.py:75-77) and the
// PETSc MatMult / VecDot / VecAXPY inside KSPCG (cslvr/momentumbp.py:183-184).
//
// Element math is the closed form of SURVEY.md §8a: for P1 the strain rate is
// cell-constant, so the only quadrature is  wbar = |T| sum_q w_q A(x_q)^(-1/n).
#include "bp_internal.h"
#include <algorithm>
#include <cstdlib>

__constant__ double c_qw[BP_MAX_QUAD];
__constant__ double c_ql[BP_MAX_QUAD * 4];
__constant__ double c_qw2[BP_MAX_QUAD];          // projection right-hand sides (degree 6)
__constant__ double c_ql2[BP_MAX_QUAD * 4];

static AsmParams make_params_fwd(const bp_ctx* c);

#define TPB 128

// ---------------------------------------------------------------- math helpers
__device__ __forceinline__ double pow_m1n(double a, const AsmParams& P)
{	// a^(-1/n)
	return P.n_is_3 ? rcbrt(a) : pow(a, -1.0 / P.n);
}

// ---------------------------------------------------------------- rate-factor quadrature
// ainv[t] = sum_q w_q A(x_q)^(-1/n): the only quadrature of the BP cell integrals (the strain
// rate of a P1 velocity is cell-constant).  It does not depend on u, so it is evaluated once
// per coefficient change, not once per Newton iteration.
//   ENERGY = false: A is the P1 Function model.A (cslvr/model.py:2530);
//   ENERGY = true : A = E a_T (1 + 181.25 min(W, 0.01)) exp(-Q_T / (R T')) evaluated at the
//                   quadrature points from the P1 fields T', W, E with a_T, Q_T switching at
//                   T' = 263.15 K (Model.form_energy_dependent_rate_factor, model.py:1815-1859).
struct RateLaw { double a_T_l, a_T_u, Q_T_l, Q_T_u, R; };
template <bool ENERGY>
__global__ void __launch_bounds__(128)
k_tet_ainv(int nt, const int4* __restrict__ tets, const double* __restrict__ Av, const double* __restrict__ Tp,
           const double* __restrict__ W, const double* __restrict__ E, RateLaw L, AsmParams P, double* __restrict__ ainv)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nt) return;
	const int4 tv = tets[t];
	const int vid[4] = { tv.x, tv.y, tv.z, tv.w };
	if (!ENERGY) {
		double A[4];
		#pragma unroll
		for (int a = 0; a < 4; ++a) A[a] = Av[vid[a]];
		if (A[0] == A[1] && A[0] == A[2] && A[0] == A[3]) { ainv[t] = pow_m1n(A[0], P); return; }
		double s = 0.0;
		for (int q = 0; q < P.n_quad; ++q)
			s += c_qw[q] * pow_m1n(c_ql[4 * q] * A[0] + c_ql[4 * q + 1] * A[1] + c_ql[4 * q + 2] * A[2] + c_ql[4 * q + 3] * A[3], P);
		ainv[t] = s;
	} else {
		double T[4], Wv[4], Ev[4];
		#pragma unroll
		for (int a = 0; a < 4; ++a) { T[a] = Tp[vid[a]]; Wv[a] = W[vid[a]]; Ev[a] = E[vid[a]]; }
		double s = 0.0;
		for (int q = 0; q < P.n_quad; ++q) {
			const double l0 = c_ql[4 * q], l1 = c_ql[4 * q + 1], l2 = c_ql[4 * q + 2], l3 = c_ql[4 * q + 3];
			const double Tq = l0 * T[0] + l1 * T[1] + l2 * T[2] + l3 * T[3];
			const double Wq = l0 * Wv[0] + l1 * Wv[1] + l2 * Wv[2] + l3 * Wv[3];
			const double Eq = l0 * Ev[0] + l1 * Ev[1] + l2 * Ev[2] + l3 * Ev[3];
			const bool cold = Tq < 263.15;
			const double aT = cold ? L.a_T_l : L.a_T_u, QT = cold ? L.Q_T_l : L.Q_T_u;
			const double WT = (Wq < 0.01) ? Wq : 0.01;
			s += c_qw[q] * pow_m1n(Eq * aT * (1.0 + 181.25 * WT) * exp(-QT / (L.R * Tq)), P);
		}
		ainv[t] = s;
	}
}

// ---------------------------------------------------------------- cell kernel
// One thread per tetrahedron.  Loads: int4 connectivity, 4 x {x,y,z,S} (32 B each),
// 4 x {u_x,u_y} (16 B), 4 x A.  Builds F_cell (8) and J_cell (8x8, symmetric) in
// registers and scatter-adds into F / the resident BSR Jacobian.
template <bool WANT_F, bool WANT_J>
__global__ void __launch_bounds__(TPB)
k_cell(int nt, const int4* __restrict__ tets, const double4* __restrict__ geo,
       const int32_t* __restrict__ v2n, const double2* __restrict__ u,
       const double* __restrict__ Av, const int32_t* __restrict__ tet_blk,
       int nn_owned, double* __restrict__ F, double* __restrict__ val, AsmParams P,
       const double* __restrict__ uzv)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nt) return;
	const int4 tv = tets[t];
	const int vid[4] = { tv.x, tv.y, tv.z, tv.w };
	double4 p[4];
	int nd[4];
	double2 uu[4];
	#pragma unroll
	for (int a = 0; a < 4; ++a) {
		p[a] = geo[vid[a]];
		nd[a] = v2n ? v2n[vid[a]] : vid[a];
	}
	#pragma unroll
	for (int a = 0; a < 4; ++a) uu[a] = u[nd[a]];

	// gradients of the barycentric coordinates
	const double e1x = p[1].x - p[0].x, e1y = p[1].y - p[0].y, e1z = p[1].z - p[0].z;
	const double e2x = p[2].x - p[0].x, e2y = p[2].y - p[0].y, e2z = p[2].z - p[0].z;
	const double e3x = p[3].x - p[0].x, e3y = p[3].y - p[0].y, e3z = p[3].z - p[0].z;
	double X[4], Y[4], Z[4];
	X[1] = e2y * e3z - e2z * e3y; Y[1] = e2z * e3x - e2x * e3z; Z[1] = e2x * e3y - e2y * e3x;
	X[2] = e3y * e1z - e3z * e1y; Y[2] = e3z * e1x - e3x * e1z; Z[2] = e3x * e1y - e3y * e1x;
	X[3] = e1y * e2z - e1z * e2y; Y[3] = e1z * e2x - e1x * e2z; Z[3] = e1x * e2y - e1y * e2x;
	const double det = e1x * X[1] + e1y * Y[1] + e1z * Z[1];
	const double idet = 1.0 / det;
	#pragma unroll
	for (int a = 1; a < 4; ++a) { X[a] *= idet; Y[a] *= idet; Z[a] *= idet; }
	X[0] = -(X[1] + X[2] + X[3]); Y[0] = -(Y[1] + Y[2] + Y[3]); Z[0] = -(Z[1] + Z[2] + Z[3]);
	const double vol = fabs(det) * (1.0 / 6.0);

	// wbar = |T| * sum_q w_q A_q^(-1/n): the quadrature sum does not depend on u and is kept per tet (k_tet_ainv)
	const double wbar = vol * Av[t];

	// velocity gradient g[c][j] and surface gradient
	double g00 = 0, g01 = 0, g02 = 0, g10 = 0, g11 = 0, g12 = 0, sx = 0, sy = 0;
	#pragma unroll
	for (int a = 0; a < 4; ++a) {
		g00 += uu[a].x * X[a]; g01 += uu[a].x * Y[a]; g02 += uu[a].x * Z[a];
		g10 += uu[a].y * X[a]; g11 += uu[a].y * Y[a]; g12 += uu[a].y * Z[a];
		sx += p[a].w * X[a]; sy += p[a].w * Y[a];
	}
	double wx = 0.0, wy = 0.0;
	if (P.rs) {   // eps_02 = (u_x,z + u_z,x)/2, eps_12 = (u_y,z + u_z,y)/2 (cslvr/momentumstokes.py:406-418)
		#pragma unroll
		for (int a = 0; a < 4; ++a) { const double w = uzv[vid[a]]; wx += w * X[a]; wy += w * Y[a]; }
		g02 += wx; g12 += wy;
	}
	const double sh = 0.5 * (g01 + g10);
	const double ed = g00 * g00 + g11 * g11 + g00 * g11 + sh * sh + 0.25 * (g02 * g02 + g12 * g12) + P.eps_reg;
	// 2 eta = ed^((1-n)/(2n)),  2 eta' = (1-n)/(2n) ed^((1-n)/(2n) - 1)
	double two_eta, two_eta_p;
	if (P.n_is_3) {
		two_eta = rcbrt(ed);
		two_eta_p = (-1.0 / 3.0) * two_eta / ed;
	} else {
		const double e = (1.0 - P.n) / (2.0 * P.n);
		two_eta = pow(ed, e);
		two_eta_p = e * two_eta / ed;
	}
	const double alpha = wbar * two_eta, gamma = P.picard ? 0.0 : wbar * two_eta_p;
	const double cxx = 2.0 * g00 + g11, cyy = 2.0 * g11 + g00, hz0 = 0.5 * g02, hz1 = 0.5 * g12;
	double dx[4], dy[4];
	#pragma unroll
	for (int a = 0; a < 4; ++a) {
		dx[a] = cxx * X[a] + sh * Y[a] + hz0 * Z[a];
		dy[a] = cyy * Y[a] + sh * X[a] + hz1 * Z[a];
	}

	if (WANT_F) {
		const double grav = P.rho_g * vol * 0.25;
		#pragma unroll
		for (int a = 0; a < 4; ++a) {
			if (nd[a] < nn_owned) {
				// linear=True keeps the u-independent part: gravity and, in the reduced-Stokes model, the u_z part of d(epsdot)
				const double fx = P.picard ? alpha * 0.5 * wx * Z[a] : alpha * dx[a];
				const double fy = P.picard ? alpha * 0.5 * wy * Z[a] : alpha * dy[a];
				atomicAdd(&F[2 * nd[a]], fx + grav * sx);
				atomicAdd(&F[2 * nd[a] + 1], fy + grav * sy);
			}
		}
	}
	if (WANT_J) {
		#pragma unroll
		for (int a = 0; a < 4; ++a) {
			const double Pa = 2.0 * alpha * X[a], Qa = 0.5 * alpha * Y[a], Ra = 0.5 * alpha * Z[a];
			const double Sa = 2.0 * alpha * Y[a], Ta = 0.5 * alpha * X[a];
			const double Ua = alpha * X[a], Va = 0.5 * alpha * Y[a];
			const double Ea = gamma * dx[a], Ga = gamma * dy[a];
			#pragma unroll
			for (int b = 0; b < 4; ++b) {
				const int slot = tet_blk[(a * 4 + b) * nt + t];
				if (slot < 0) continue;
				const double jxx = Pa * X[b] + Qa * Y[b] + Ra * Z[b] + Ea * dx[b];
				const double jyy = Sa * Y[b] + Ta * X[b] + Ra * Z[b] + Ga * dy[b];
				const double jxy = Ua * Y[b] + Va * X[b] + Ea * dy[b];           // row (x,a), col (y,b)
				const double jyx = alpha * Y[a] * X[b] + 0.5 * alpha * X[a] * Y[b] + Ga * dx[b];   // row (y,a), col (x,b)
				double* v = val + 4 * (size_t)slot;
				atomicAdd(v + 0, jxx); atomicAdd(v + 1, jxy);
				atomicAdd(v + 2, jyx); atomicAdd(v + 3, jyy);
			}
		}
	}
}

#ifdef BP_EXPERIMENTS
// ---------------------------------------------------------------- EXPERIMENT (timing only)
// Thread-per-tet element computation with the 64 + 8 contributions accumulated by
// shared-memory atomicAdd(double) (ATOMS.CAST.SPIN.64) at slot-derived addresses in a
// 200 KB strip: measures what a tile kernel with shared-memory accumulation would cost.
__global__ void __launch_bounds__(384, 1)
k_proxy_smem(int nt, const int4* __restrict__ tets, const double4* __restrict__ geo,
             const int32_t* __restrict__ v2n, const double2* __restrict__ u,
             const double* __restrict__ Av, const int32_t* __restrict__ tet_blk, AsmParams P, int nacc, double* sink)
{
	extern __shared__ double acc[];
	for (int k = threadIdx.x; k < nacc; k += blockDim.x) acc[k] = 0.0;
	__syncthreads();
	const int per = (nt + gridDim.x - 1) / gridDim.x;
	const int t0 = blockIdx.x * per, t1 = min(nt, t0 + per);
	for (int t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
		const int4 tv = tets[t];
		const int vid[4] = { tv.x, tv.y, tv.z, tv.w };
		double4 p[4]; int nd[4]; double2 uu[4];
		#pragma unroll
		for (int a = 0; a < 4; ++a) { p[a] = geo[vid[a]]; nd[a] = v2n ? v2n[vid[a]] : vid[a]; }
		#pragma unroll
		for (int a = 0; a < 4; ++a) uu[a] = u[nd[a]];
		const double e1x = p[1].x - p[0].x, e1y = p[1].y - p[0].y, e1z = p[1].z - p[0].z;
		const double e2x = p[2].x - p[0].x, e2y = p[2].y - p[0].y, e2z = p[2].z - p[0].z;
		const double e3x = p[3].x - p[0].x, e3y = p[3].y - p[0].y, e3z = p[3].z - p[0].z;
		double X[4], Y[4], Z[4];
		X[1] = e2y * e3z - e2z * e3y; Y[1] = e2z * e3x - e2x * e3z; Z[1] = e2x * e3y - e2y * e3x;
		X[2] = e3y * e1z - e3z * e1y; Y[2] = e3z * e1x - e3x * e1z; Z[2] = e3x * e1y - e3y * e1x;
		X[3] = e1y * e2z - e1z * e2y; Y[3] = e1z * e2x - e1x * e2z; Z[3] = e1x * e2y - e1y * e2x;
		const double det = e1x * X[1] + e1y * Y[1] + e1z * Z[1];
		const double idet = 1.0 / det;
		#pragma unroll
		for (int a = 1; a < 4; ++a) { X[a] *= idet; Y[a] *= idet; Z[a] *= idet; }
		X[0] = -(X[1] + X[2] + X[3]); Y[0] = -(Y[1] + Y[2] + Y[3]); Z[0] = -(Z[1] + Z[2] + Z[3]);
		const double vol = fabs(det) * (1.0 / 6.0);
		const double wbar = vol * Av[t];
		double g00 = 0, g01 = 0, g02 = 0, g10 = 0, g11 = 0, g12 = 0, sx = 0, sy = 0;
		#pragma unroll
		for (int a = 0; a < 4; ++a) {
			g00 += uu[a].x * X[a]; g01 += uu[a].x * Y[a]; g02 += uu[a].x * Z[a];
			g10 += uu[a].y * X[a]; g11 += uu[a].y * Y[a]; g12 += uu[a].y * Z[a];
			sx += p[a].w * X[a]; sy += p[a].w * Y[a];
		}
		const double sh = 0.5 * (g01 + g10);
		const double ed = g00 * g00 + g11 * g11 + g00 * g11 + sh * sh + 0.25 * (g02 * g02 + g12 * g12) + P.eps_reg;
		const double two_eta = rcbrt(ed), two_eta_p = (-1.0 / 3.0) * (two_eta * two_eta) * (two_eta * two_eta);
		const double alpha = wbar * two_eta, gamma = wbar * two_eta_p;
		const double cxx = 2.0 * g00 + g11, cyy = 2.0 * g11 + g00, hz0 = 0.5 * g02, hz1 = 0.5 * g12;
		double dx[4], dy[4];
		#pragma unroll
		for (int a = 0; a < 4; ++a) { dx[a] = cxx * X[a] + sh * Y[a] + hz0 * Z[a]; dy[a] = cyy * Y[a] + sh * X[a] + hz1 * Z[a]; }
		const double grav = P.rho_g * vol * 0.25;
		#pragma unroll
		for (int a = 0; a < 4; ++a) {
			atomicAdd(&acc[(2 * nd[a]) % nacc], alpha * dx[a] + grav * sx);
			atomicAdd(&acc[(2 * nd[a] + 1) % nacc], alpha * dy[a] + grav * sy);
			const double Pa = 2.0 * alpha * X[a], Qa = 0.5 * alpha * Y[a], Ra = 0.5 * alpha * Z[a];
			const double Sa = 2.0 * alpha * Y[a], Ta = 0.5 * alpha * X[a];
			const double Ua = alpha * X[a], Wa = alpha * Y[a];
			const double Ea = gamma * dx[a], Ga = gamma * dy[a];
			#pragma unroll
			for (int b = 0; b < 4; ++b) {
				const int slot = tet_blk[(a * 4 + b) * nt + t];
				double* v = acc + (4 * (size_t)slot) % (size_t)(nacc - 4);
				atomicAdd(v + 0, Pa * X[b] + Qa * Y[b] + Ra * Z[b] + Ea * dx[b]);
				atomicAdd(v + 1, Ua * Y[b] + Qa * X[b] + Ea * dy[b]);
				atomicAdd(v + 2, Wa * X[b] + Ta * Y[b] + Ga * dx[b]);
				atomicAdd(v + 3, Sa * Y[b] + Ta * X[b] + Ra * Z[b] + Ga * dy[b]);
			}
		}
	}
	__syncthreads();
	double s = 0.0;
	for (int k = threadIdx.x; k < nacc; k += blockDim.x) s += acc[k];
	if (s == 1.2345e-300) sink[0] = s;
}

void bp_launch_proxy(bp_ctx* c)
{
	const AsmParams P = make_params_fwd(c);
	const int nacc = 25000;
	cudaFuncSetAttribute(k_proxy_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, nacc * 8);
	k_proxy_smem<<<148, 384, nacc * 8, c->stream>>>((int)c->nt, c->d_tets, c->d_geo, c->d_v2n, (const double2*)c->d_u, c->d_ainv, c->d_tet_blk, P, nacc, c->d_part);
	c->launches++;
}

#endif  // BP_EXPERIMENTS

// ---------------------------------------------------------------- row-gather kernel
// One thread per Jacobian block row (= owned node i).  The thread walks the tets that
// touch node i (SELL-32 incidence list), rebuilds each tet's cell-constant quantities
// from the vertex data with node i rotated to local position 0, and adds the two rows
// (u_x, u_y of node i) of that tet's element matrix into a private accumulator in
// shared memory (layout [entry][thread]: bank = thread, conflict-free).  Every J entry
// and every F entry is then written exactly once: no atomics, no zero-fill, bitwise
// reproducible.  The price is that the cell-constant part of a tet is recomputed by
// each of its four nodes.
#define RTPB BP_ROWS_PER_CTA
#ifndef ROWS_MIN_CTAS
#define ROWS_MIN_CTAS 3
#endif
template <bool WANT_F, bool WANT_J, bool RS>
__global__ void __launch_bounds__(RTPB, ROWS_MIN_CTAS)
k_rows(int nrows, const int32_t* __restrict__ inc_slice, const int32_t* __restrict__ inc_code,
       const uint32_t* __restrict__ inc_pos, const int32_t* __restrict__ inc_v, size_t nslot,
       const double* __restrict__ soa, size_t nv,
       const double2* __restrict__ u /* per VERTEX */, const double* __restrict__ Av,
       const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ rowlen,
       double* __restrict__ F, double* __restrict__ val, AsmParams P, const double* __restrict__ uzv, int blk0,
       const int32_t* __restrict__ mirror)
{
	extern __shared__ double acc[];                  // [4*(max_row-1)][RTPB]: the blocks right of the diagonal, strip slot = row position - diagonal's - 1
	const int tid = threadIdx.x;
	const int i = (blockIdx.x + blk0) * RTPB + tid;
	const bool active = i < nrows;
	const int len = active ? rowlen[i] : 0;
	double Fx = 0.0, Fy = 0.0;
	double dg0 = 0.0, dg1 = 0.0, dg2 = 0.0, dg3 = 0.0;    // diagonal block
	int kdiag = 1 << 30;
	const int slice = i >> 5, lane = i & 31;
	const int e0 = active ? inc_slice[slice] : 0, e1 = active ? inc_slice[slice + 1] : 0;
	// software pipeline, two stages deep on the index words (tet code, block positions, four
	// vertex ids: coalesced SELL-32 streams) and one stage on the gathered vertex data:
	//   iteration e computes on data(e), requests data(e+1) from idx(e+1) once its raw inputs
	//   are consumed, and requests idx(e+2) at the top.
	int e = e0 + lane;
	int code = (e < e1) ? inc_code[e] : -1;
	uint32_t pk = 0, pk1 = 0;
	int vid[4] = { 0, 0, 0, 0 }, vid1[4] = { 0, 0, 0, 0 };
	int code1 = -1;
	double4 p[4];
	double2 uu[4];
	double ww[4] = { 0.0, 0.0, 0.0, 0.0 };           // u_z at the vertices (reduced-Stokes model only)
	double ai = 0.0;                                 // sum_q w_q A_q^(-1/n) of the tet (k_tet_ainv)
	if (code >= 0) {
		pk = inc_pos[e];
		kdiag = (int)(pk & 255);                     // position of the diagonal block: the same in every incidence of the row
		#pragma unroll
		for (int j = 0; j < 4; ++j) vid[j] = inc_v[j * nslot + e];
		if (e + 32 < e1) {                           // padding slots hold code -1, vertex 0: safe to load
			code1 = inc_code[e + 32];
			pk1 = inc_pos[e + 32];
			#pragma unroll
			for (int j = 0; j < 4; ++j) vid1[j] = inc_v[j * nslot + e + 32];
		}
		#pragma unroll
		for (int j = 0; j < 4; ++j) {
			p[j] = make_double4(soa[vid[j]], soa[nv + vid[j]], soa[2 * nv + vid[j]], soa[3 * nv + vid[j]]);
			uu[j] = u[vid[j]];
			if (RS) ww[j] = uzv[vid[j]];
		}
		ai = Av[code >> 2];
	}
	int own = vid[0];                                // vertex whose data sits in p[0] / uu[0]
	// J is symmetric: this thread accumulates the diagonal block and the blocks right of it; their
	// transposes go to the rows of the column nodes at the end (mirror slots)
	if (WANT_J && code >= 0) for (int k = 0; k < 4 * (len - 1 - kdiag); ++k) acc[k * RTPB + tid] = 0.0;
	while (code >= 0) {
		const uint32_t pk_cur = pk;
		e += 32;
		// stage idx(e+1) -> current "next", request idx(e+2)
		const int code_n = code1;
		pk = pk1;
		#pragma unroll
		for (int j = 0; j < 4; ++j) vid[j] = vid1[j];
		code1 = -1;
		if (e + 32 < e1) {
			code1 = inc_code[e + 32];
			pk1 = inc_pos[e + 32];
			#pragma unroll
			for (int j = 0; j < 4; ++j) vid1[j] = inc_v[j * nslot + e + 32];
		}
		const double e1x = p[1].x - p[0].x, e1y = p[1].y - p[0].y, e1z = p[1].z - p[0].z;
		const double e2x = p[2].x - p[0].x, e2y = p[2].y - p[0].y, e2z = p[2].z - p[0].z;
		const double e3x = p[3].x - p[0].x, e3y = p[3].y - p[0].y, e3z = p[3].z - p[0].z;
		double X[4], Y[4], Z[4];
		X[1] = e2y * e3z - e2z * e3y; Y[1] = e2z * e3x - e2x * e3z; Z[1] = e2x * e3y - e2y * e3x;
		X[2] = e3y * e1z - e3z * e1y; Y[2] = e3z * e1x - e3x * e1z; Z[2] = e3x * e1y - e3y * e1x;
		X[3] = e1y * e2z - e1z * e2y; Y[3] = e1z * e2x - e1x * e2z; Z[3] = e1x * e2y - e1y * e2x;
		const double det = e1x * X[1] + e1y * Y[1] + e1z * Z[1];
		const double idet = __drcp_rn(det);
		#pragma unroll
		for (int j = 1; j < 4; ++j) { X[j] *= idet; Y[j] *= idet; Z[j] *= idet; }
		X[0] = -(X[1] + X[2] + X[3]); Y[0] = -(Y[1] + Y[2] + Y[3]); Z[0] = -(Z[1] + Z[2] + Z[3]);
		const double vol = fabs(det) * (1.0 / 6.0);
		const double wbar = vol * ai;
		double g00 = 0, g01 = 0, g02 = 0, g10 = 0, g11 = 0, g12 = 0, sx = 0, sy = 0;
		#pragma unroll
		for (int j = 0; j < 4; ++j) {
			g00 += uu[j].x * X[j]; g01 += uu[j].x * Y[j]; g02 += uu[j].x * Z[j];
			g10 += uu[j].y * X[j]; g11 += uu[j].y * Y[j]; g12 += uu[j].y * Z[j];
			sx += p[j].w * X[j]; sy += p[j].w * Y[j];
		}
		double wx = 0.0, wy = 0.0;
		if (RS) {   // eps_02 = (u_x,z + u_z,x)/2, eps_12 = (u_y,z + u_z,y)/2 (cslvr/momentumstokes.py:406-418)
			#pragma unroll
			for (int j = 0; j < 4; ++j) { wx += ww[j] * X[j]; wy += ww[j] * Y[j]; }
			g02 += wx; g12 += wy;
		}
		// raw vertex data of this incidence is consumed: request the next incidence's (the row's own
		// vertex only when it changes: periodic images)
		if (code_n >= 0) {
			#pragma unroll
			for (int j = 0; j < 4; ++j) {
				if (j == 0 && vid[0] == own) continue;
				p[j] = make_double4(soa[vid[j]], soa[nv + vid[j]], soa[2 * nv + vid[j]], soa[3 * nv + vid[j]]);
				uu[j] = u[vid[j]];
				if (RS) ww[j] = uzv[vid[j]];
			}
			own = vid[0];
			ai = Av[code_n >> 2];
		}
		const double sh = 0.5 * (g01 + g10);
		const double ed = g00 * g00 + g11 * g11 + g00 * g11 + sh * sh + 0.25 * (g02 * g02 + g12 * g12) + P.eps_reg;
		double two_eta, two_eta_p;
		if (P.n_is_3) {
			two_eta = rcbrt(ed);
			two_eta_p = (-1.0 / 3.0) * (two_eta * two_eta) * (two_eta * two_eta);    // ed^(-4/3)
		} else {
			const double ex = (1.0 - P.n) / (2.0 * P.n);
			two_eta = pow(ed, ex);
			two_eta_p = ex * two_eta / ed;
		}
		const double alpha = wbar * two_eta, gamma = P.picard ? 0.0 : wbar * two_eta_p;
		const double cxx = 2.0 * g00 + g11, cyy = 2.0 * g11 + g00, hz0 = 0.5 * g02, hz1 = 0.5 * g12;
		const double dx0 = cxx * X[0] + sh * Y[0] + hz0 * Z[0];
		const double dy0 = cyy * Y[0] + sh * X[0] + hz1 * Z[0];
		if (WANT_F) {
			const double grav = P.rho_g * vol * 0.25;
			// linear=True keeps the u-independent part: gravity and, in the reduced-Stokes model, the u_z part of d(epsdot)
			Fx += (P.picard ? alpha * 0.5 * wx * Z[0] : alpha * dx0) + grav * sx;
			Fy += (P.picard ? alpha * 0.5 * wy * Z[0] : alpha * dy0) + grav * sy;
		}
		if (WANT_J) {
			const double Pa = 2.0 * alpha * X[0], Qa = 0.5 * alpha * Y[0], Ra = 0.5 * alpha * Z[0];
			const double Sa = 2.0 * alpha * Y[0], Ta = 0.5 * alpha * X[0];
			const double Ua = alpha * X[0], Wa = alpha * Y[0];
			const double Ea = gamma * dx0, Ga = gamma * dy0;
			#pragma unroll
			for (int b = 0; b < 4; ++b) {
				// the vertices of an incidence are ordered "column node > row node first" (bp_setup.cu): the
				// first block left of the diagonal ends the visit
				const int pos = (int)((pk_cur >> (8 * b)) & 255);
				if (b > 0 && pos < kdiag) break;
				const double dxb = cxx * X[b] + sh * Y[b] + hz0 * Z[b];
				const double dyb = cyy * Y[b] + sh * X[b] + hz1 * Z[b];
				const double jxx = Pa * X[b] + Qa * Y[b] + Ra * Z[b] + Ea * dxb;
				const double jxy = Ua * Y[b] + Qa * X[b] + Ea * dyb;
				const double jyx = Wa * X[b] + Ta * Y[b] + Ga * dxb;
				const double jyy = Sa * Y[b] + Ta * X[b] + Ra * Z[b] + Ga * dyb;
				if (b == 0) {                             // the diagonal block gets a term from every tet: keep it in registers
					dg0 += jxx; dg1 += jxy; dg2 += jyx; dg3 += jyy;
				} else {
					double* s4 = acc + (4 * (pos - kdiag - 1)) * RTPB + tid;
					s4[0] += jxx; s4[RTPB] += jxy; s4[2 * RTPB] += jyx; s4[3 * RTPB] += jyy;
				}
			}
		}
		code = code_n;
	}
	if (!active) return;
	if (WANT_F) { F[2 * i] = Fx; F[2 * i + 1] = Fy; }
	if (WANT_J) {
		// SELL-32: block k of the 32 rows of a slice is contiguous -> coalesced stores for the diagonal and the
		// blocks right of it; the blocks left of the diagonal are written by the threads of their column nodes
		const size_t base = (size_t)sell_ptr[slice] + lane;
		double2* out = (double2*)val + 2 * base;
		if (kdiag < len) { out[64 * (size_t)kdiag] = make_double2(dg0, dg1); out[64 * (size_t)kdiag + 1] = make_double2(dg2, dg3); }
		for (int k = kdiag + 1; k < len; ++k) {
			const double* s4 = acc + (4 * (k - kdiag - 1)) * RTPB + tid;
			const double a00 = s4[0], a01 = s4[RTPB], a10 = s4[2 * RTPB], a11 = s4[3 * RTPB];
			out[64 * (size_t)k]     = make_double2(a00, a01);
			out[64 * (size_t)k + 1] = make_double2(a10, a11);
			const int ms = mirror[base + 32 * (size_t)k];
			if (ms >= 0) {
				double2* mo = (double2*)val + 2 * (size_t)ms;
				mo[0] = make_double2(a00, a10);
				mo[1] = make_double2(a01, a11);
			}
		}
	}
}

// ---------------------------------------------------------------- projection systems
// Same row-gather mapping as k_rows for the scalar P1 systems behind the follow-on solves
// of cslvr/momentumbp.py:205-229: the mass matrix int(phi_i phi_j) = |T|(1+delta)/20 goes
// into both diagonal entries of the 2x2 node blocks (so the block PCG solves two copies,
// the second with a zero right-hand side) and the right-hand side into F[2i].
// KIND 0: solve_pressure, rhs_i = int (rho g (S - z) - 2 eta(u) (u_x,x + u_y,y)) phi_i.
// KIND 1: projection of the basal kinematic condition (cslvr/momentumbp.py:68-72),
//         rhs_i = int ((-B_ring - u_x n_b0 - u_y n_b1) / n_b2) phi_i, n_b = (grad B - k)/|.|.
// KIND 2: the continuity system of solve_vert_velocity (momentumbp.py:75, 244-248):
//         A_ij = int dz(phi_j) phi_i = Z_j |T|/4 (NOT symmetric), b_i = -int (u_x,x + u_y,y) phi_i,
//         rows of the basal nodes replaced by identity with the projected u_z_b as value.
template <int KIND>
__global__ void __launch_bounds__(RTPB)
k_rows_aux(int nrows, const int32_t* __restrict__ inc_slice, const int32_t* __restrict__ inc_code,
           const uint32_t* __restrict__ inc_pos, const int32_t* __restrict__ inc_v, size_t nslot,
           const double* __restrict__ soa, size_t nv, const double2* __restrict__ u, const double* __restrict__ Av,
           const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ rowlen,
           double* __restrict__ F, double* __restrict__ val, AsmParams P, int nq,
           const double* __restrict__ Bv, const double* __restrict__ Bring,
           const double2* __restrict__ uzb, const uint8_t* __restrict__ bcnode,
           const double2* __restrict__ u_eta, const double* __restrict__ uzv)
{
	extern __shared__ double acc[];                  // [max_row][RTPB]
	const int tid = threadIdx.x;
	const int i = blockIdx.x * RTPB + tid;
	if (i >= nrows) return;
	const int len = rowlen[i];
	for (int k = 0; k < len; ++k) acc[k * RTPB + tid] = 0.0;
	int kd = 0;
	double rhs = 0.0;
	const int slice = i >> 5, lane = i & 31;
	const int e0 = inc_slice[slice], e1 = inc_slice[slice + 1];
	for (int e = e0 + lane; e < e1; e += 32) {
		const int code = inc_code[e];
		if (code < 0) break;
		const uint32_t pk = inc_pos[e];
		const int a = code & 3;
		double4 p[4]; double2 uu[4]; double A[4];
		#pragma unroll
		for (int j = 0; j < 4; ++j) {
			const int v = inc_v[j * nslot + e];
			p[j] = make_double4(soa[v], soa[nv + v], soa[2 * nv + v], soa[3 * nv + v]);
			A[j] = Av[v]; uu[j] = u[v];
		}
		const double e1x = p[1].x - p[0].x, e1y = p[1].y - p[0].y, e1z = p[1].z - p[0].z;
		const double e2x = p[2].x - p[0].x, e2y = p[2].y - p[0].y, e2z = p[2].z - p[0].z;
		const double e3x = p[3].x - p[0].x, e3y = p[3].y - p[0].y, e3z = p[3].z - p[0].z;
		double X[4], Y[4], Z[4];
		X[1] = e2y * e3z - e2z * e3y; Y[1] = e2z * e3x - e2x * e3z; Z[1] = e2x * e3y - e2y * e3x;
		X[2] = e3y * e1z - e3z * e1y; Y[2] = e3z * e1x - e3x * e1z; Z[2] = e3x * e1y - e3y * e1x;
		X[3] = e1y * e2z - e1z * e2y; Y[3] = e1z * e2x - e1x * e2z; Z[3] = e1x * e2y - e1y * e2x;
		const double det = e1x * X[1] + e1y * Y[1] + e1z * Z[1];
		const double idet = 1.0 / det;
		#pragma unroll
		for (int j = 1; j < 4; ++j) { X[j] *= idet; Y[j] *= idet; Z[j] *= idet; }
		X[0] = -(X[1] + X[2] + X[3]); Y[0] = -(Y[1] + Y[2] + Y[3]); Z[0] = -(Z[1] + Z[2] + Z[3]);
		const double vol = fabs(det) * (1.0 / 6.0);
		const double m = vol * (1.0 / 20.0);
		if (KIND != 2) {
			#pragma unroll
			for (int b = 0; b < 4; ++b) acc[((pk >> (8 * b)) & 255) * RTPB + tid] += (b == 0) ? 2.0 * m : m;
		}
		if (KIND == 1) {
			double bx = 0, by = 0, bz = 0, Bj[4], Rj[4];
			#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const int v = inc_v[j * nslot + e];
				Bj[j] = Bv[v]; Rj[j] = Bring[v];
				bx += Bj[j] * X[j]; by += Bj[j] * Y[j]; bz += Bj[j] * Z[j];
			}
			const double nz = bz - 1.0, imag = rsqrt(1.0 + bx * bx + by * by + bz * bz);   // n_b_mag = sqrt(1 + grad B . grad B), momentumbp.py:69
			const double n0 = bx * imag, n1 = by * imag, n2 = nz * imag;
			double fs = 0.0, f0 = 0.0;
			#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const double f = (-Rj[j] - uu[j].x * n0 - uu[j].y * n1) / n2;
				fs += f; if (j == 0) f0 = f;
			}
			rhs += m * (fs + f0);
		}
		if (KIND == 2) {
			double div = 0.0;
			#pragma unroll
			for (int j = 0; j < 4; ++j) div += uu[j].x * X[j] + uu[j].y * Y[j];
			#pragma unroll
			for (int b = 0; b < 4; ++b) acc[((pk >> (8 * b)) & 255) * RTPB + tid] += Z[b] * vol * 0.25;
			rhs -= div * vol * 0.25;
			kd = (int)(pk & 255);                                     // position of the diagonal block in the row
		}
		if (KIND == 0) {
			// viscosity state: u itself (BP: model.u was just updated) or, for the reduced-Stokes solve(), the
			// velocity the model still holds plus the frozen u_z (bp_rs_solve_pressure)
			double g00 = 0, g01 = 0, g02 = 0, g10 = 0, g11 = 0, g12 = 0, div = 0;
			#pragma unroll
			for (int j = 0; j < 4; ++j) {
				const int v = inc_v[j * nslot + e];
				const double2 ue = u_eta ? u_eta[v] : uu[j];
				const double w = uzv ? uzv[v] : 0.0;
				g00 += ue.x * X[j]; g01 += ue.x * Y[j]; g02 += ue.x * Z[j] + w * X[j];
				g10 += ue.y * X[j]; g11 += ue.y * Y[j]; g12 += ue.y * Z[j] + w * Y[j];
				div += uu[j].x * X[j] + uu[j].y * Y[j];
			}
			const double sh = 0.5 * (g01 + g10);
			const double ed = g00 * g00 + g11 * g11 + g00 * g11 + sh * sh + 0.25 * (g02 * g02 + g12 * g12) + P.eps_reg;
			const double two_eta_e = P.n_is_3 ? rcbrt(ed) : pow(ed, (1.0 - P.n) / (2.0 * P.n));
			double wq;                                   // sum_q w_q lambda_i(q) A_q^(-1/n)
			if (A[0] == A[1] && A[0] == A[2] && A[0] == A[3]) wq = 0.25 * pow_m1n(A[0], P);
			else {
				double Ao[4];
				#pragma unroll
				for (int j = 0; j < 4; ++j) Ao[(a + j) & 3] = A[j];
				wq = 0.0;
				for (int q = 0; q < nq; ++q) {
					const double aq = c_ql2[4 * q] * Ao[0] + c_ql2[4 * q + 1] * Ao[1] + c_ql2[4 * q + 2] * Ao[2] + c_ql2[4 * q + 3] * Ao[3];
					wq += c_qw2[q] * c_ql2[4 * q + a] * pow_m1n(aq, P);
				}
			}
			const double sz0 = p[0].w - p[0].z, szs = sz0 + (p[1].w - p[1].z) + (p[2].w - p[2].z) + (p[3].w - p[3].z);
			rhs += P.rho_g * m * (szs + sz0) - div * two_eta_e * vol * wq;
		}
	}
	if (KIND == 2 && bcnode[i]) {                  // DirichletBC row: identity, value u_z_b
		for (int k = 0; k < len; ++k) acc[k * RTPB + tid] = (k == kd) ? 1.0 : 0.0;
		rhs = uzb[i].x;
	}
	F[2 * i] = rhs; F[2 * i + 1] = 0.0;
	double2* out = (double2*)val + 2 * ((size_t)sell_ptr[slice] + lane);
	for (int k = 0; k < len; ++k) {
		const double mv = acc[k * RTPB + tid];
		out[64 * (size_t)k] = make_double2(mv, 0.0);
		out[64 * (size_t)k + 1] = make_double2(0.0, mv);
	}
}

__global__ void k_node_to_vertex(int nv, const int32_t* __restrict__ v2n, const double* __restrict__ nodevec, int comp, double* __restrict__ out)
{
	for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) out[v] = nodevec[2 * (v2n ? v2n[v] : v) + comp];
}

// ---------------------------------------------------------------- facet kernel
// Basal sliding  +1/2 beta u.u dGamma_b  and water pressure  -f_w u.n_h  on
// dGamma_bw / dGamma_lt (cslvr/momentumbp.py:473-491), exact P1 mass integrals.
template <bool WANT_F, bool WANT_J>
__global__ void __launch_bounds__(TPB)
k_facet(int nf, int nt, const int32_t* __restrict__ fcell, const int32_t* __restrict__ finfo,
        const int4* __restrict__ tets, const double4* __restrict__ geo,
        const int32_t* __restrict__ v2n, const double2* __restrict__ u,
        const double* __restrict__ beta, const int32_t* __restrict__ tet_blk,
        int nn_owned, double* __restrict__ F, double* __restrict__ val, AsmParams P,
        const double* __restrict__ Bring)
{
	const int f = blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= nf) return;
	const int cell = fcell[f], info = finfo[f];
	const int lf = info & 0xff;
	// sliding: dGamma_b (markers 3, 5) for the BP classes, dGamma_bg (3) for the reduced-Stokes model (momentumstokes.py:381)
	const bool basal = P.rs ? ((info >> 10) & 1) : ((info >> 8) & 1), press = (info >> 9) & 1;
	const int4 tv = tets[cell];
	const int vid[4] = { tv.x, tv.y, tv.z, tv.w };
	int la[3], k = 0;
	#pragma unroll
	for (int a = 0; a < 4; ++a) if (a != lf) la[k++] = a;
	double4 p[3];
	int nd[3];
	double2 uu[3];
	double bt[3], br[3] = { 0.0, 0.0, 0.0 };
	#pragma unroll
	for (int a = 0; a < 3; ++a) {
		const int v = vid[la[a]];
		p[a] = geo[v]; nd[a] = v2n ? v2n[v] : v; uu[a] = u[nd[a]]; bt[a] = beta[v];
		if (P.rs) br[a] = Bring[v];
	}
	const double4 po = geo[vid[lf]];
	const double ax = p[1].x - p[0].x, ay = p[1].y - p[0].y, az = p[1].z - p[0].z;
	const double bx = p[2].x - p[0].x, by = p[2].y - p[0].y, bz = p[2].z - p[0].z;
	double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
	const double nrm = sqrt(nx * nx + ny * ny + nz * nz);
	const double area = 0.5 * nrm;
	// outward: away from the opposite vertex
	const double side = nx * (po.x - p[0].x) + ny * (po.y - p[0].y) + nz * (po.z - p[0].z);
	const double sgn = (side > 0.0 ? -1.0 : 1.0) / nrm;
	nx *= sgn; ny *= sgn; nz *= sgn;

	double Fx[3] = { 0, 0, 0 }, Fy[3] = { 0, 0, 0 };
	if (basal) {
		// K[a][b] = |f| sum_d beta_d int(l_a l_b l_d),  int = {1/10, 1/30, 1/60}
		const double bs = bt[0] + bt[1] + bt[2];
		double K[3][3];
		#pragma unroll
		for (int a = 0; a < 3; ++a)
			#pragma unroll
			for (int b = 0; b < 3; ++b)
				K[a][b] = (a == b) ? area * (bt[a] * (1.0 / 10.0) + (bs - bt[a]) * (1.0 / 30.0))
				                   : area * ((bt[a] + bt[b]) * (1.0 / 30.0) + (bs - bt[a] - bt[b]) * (1.0 / 60.0));
		// reduced-Stokes: 1/2 beta u_z_b^2 with u_z_b = (-B_ring - u_h.n_h)/n_z (momentumstokes.py:369-370):
		// second variation beta (delta_cd + n_c n_d / n_z^2), first variation beta (u_c + (n_c/n_z)(B_ring/n_z + u_h.n_h/n_z))
		const double cx = P.rs ? nx / nz : 0.0, cy = P.rs ? ny / nz : 0.0, inz = P.rs ? 1.0 / nz : 0.0;
		#pragma unroll
		for (int a = 0; a < 3; ++a)
			#pragma unroll
			for (int b = 0; b < 3; ++b) {
				const double t = P.rs ? br[b] * inz + (P.picard ? 0.0 : cx * uu[b].x + cy * uu[b].y) : 0.0;
				if (!P.picard) { Fx[a] += K[a][b] * uu[b].x; Fy[a] += K[a][b] * uu[b].y; }
				if (P.rs) { Fx[a] += K[a][b] * cx * t; Fy[a] += K[a][b] * cy * t; }
				if (WANT_J) {
					const int slot = tet_blk[(la[a] * 4 + la[b]) * nt + cell];
					if (slot >= 0) {
						atomicAdd(val + 4 * (size_t)slot, K[a][b] * (1.0 + cx * cx)); atomicAdd(val + 4 * (size_t)slot + 3, K[a][b] * (1.0 + cy * cy));
						if (P.rs) { atomicAdd(val + 4 * (size_t)slot + 1, K[a][b] * cx * cy); atomicAdd(val + 4 * (size_t)slot + 2, K[a][b] * cx * cy); }
					}
				}
			}
	}
	if (press && WANT_F) {
		// f_w = rho g (S - z) - rho_sw g D,  D = -min(0, z)   (momentumbp.py:477, d3model.py:858-861)
		double fw[3];
		#pragma unroll
		for (int a = 0; a < 3; ++a) fw[a] = P.rho_g * (p[a].w - p[a].z) - P.rho_sw_g * (-fmin(0.0, p[a].z));
		const double fs = fw[0] + fw[1] + fw[2];
		#pragma unroll
		for (int a = 0; a < 3; ++a) {
			const double m = area * (fs + fw[a]) * (1.0 / 12.0);
			Fx[a] -= nx * m; Fy[a] -= ny * m;
		}
	}
	if (WANT_F) {
		#pragma unroll
		for (int a = 0; a < 3; ++a)
			if (nd[a] < nn_owned) { atomicAdd(&F[2 * nd[a]], Fx[a]); atomicAdd(&F[2 * nd[a] + 1], Fy[a]); }
	}
}

// ---------------------------------------------------------------- facet kernel, row gather
// The same integrals as k_facet without atomics: one thread per owned node that an integrating facet
// touches.  It walks the node's (facet, local vertex) pairs, rebuilds the facet's area / normal / mass
// integrals and adds ITS rows -- the two residual entries in registers, the Jacobian blocks by plain
// read-modify-write of the slots of its own block row, which no other thread of this kernel touches
// and which k_rows finished writing earlier in the stream.  Assembly is then free of atomics, so
// F and J are bitwise reproducible.
template <bool WANT_F, bool WANT_J>
__global__ void __launch_bounds__(TPB)
k_facet_rows(int nfn, const int32_t* __restrict__ fn_node, const int32_t* __restrict__ fn_ptr, const int32_t* __restrict__ fn_inc,
             int nt, const int32_t* __restrict__ fcell, const int32_t* __restrict__ finfo,
             const int4* __restrict__ tets, const double4* __restrict__ geo,
             const int32_t* __restrict__ v2n, const double2* __restrict__ u,
             const double* __restrict__ beta, const int32_t* __restrict__ tet_blk,
             double* __restrict__ F, double* __restrict__ val, AsmParams P, const double* __restrict__ Bring)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nfn) return;
	const int i = fn_node[r];
	double Fx = 0.0, Fy = 0.0;
	for (int q = fn_ptr[r]; q < fn_ptr[r + 1]; ++q) {
		const int inc = fn_inc[q], f = inc >> 2, a = inc & 3;          // a: my position among the facet's three vertices
		const int cell = fcell[f], info = finfo[f];
		const int lf = info & 0xff;
		const bool basal = P.rs ? ((info >> 10) & 1) : ((info >> 8) & 1), press = (info >> 9) & 1;
		const int4 tv = tets[cell];
		const int vid[4] = { tv.x, tv.y, tv.z, tv.w };
		int la[3], k = 0;
		#pragma unroll
		for (int j = 0; j < 4; ++j) if (j != lf) la[k++] = j;
		double4 p[3];
		double2 uu[3];
		double bt[3], br[3] = { 0.0, 0.0, 0.0 };
		#pragma unroll
		for (int j = 0; j < 3; ++j) {
			const int v = vid[la[j]];
			p[j] = geo[v]; uu[j] = u[v2n ? v2n[v] : v]; bt[j] = beta[v];
			if (P.rs) br[j] = Bring[v];
		}
		const double4 po = geo[vid[lf]];
		const double ax = p[1].x - p[0].x, ay = p[1].y - p[0].y, az = p[1].z - p[0].z;
		const double bx = p[2].x - p[0].x, by = p[2].y - p[0].y, bz = p[2].z - p[0].z;
		double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
		const double nrm = sqrt(nx * nx + ny * ny + nz * nz);
		const double area = 0.5 * nrm;
		const double side = nx * (po.x - p[0].x) + ny * (po.y - p[0].y) + nz * (po.z - p[0].z);
		const double sgn = (side > 0.0 ? -1.0 : 1.0) / nrm;
		nx *= sgn; ny *= sgn; nz *= sgn;
		if (basal) {
			const double bs = bt[0] + bt[1] + bt[2];
			const double cx = P.rs ? nx / nz : 0.0, cy = P.rs ? ny / nz : 0.0, inz = P.rs ? 1.0 / nz : 0.0;
			#pragma unroll
			for (int b = 0; b < 3; ++b) {
				// K[a][b] = |f| sum_d beta_d int(l_a l_b l_d),  int = {1/10, 1/30, 1/60}
				double bta = 0.0, btb = bt[b];
				#pragma unroll
				for (int j = 0; j < 3; ++j) if (j == a) bta = bt[j];
				const double Kab = (a == b) ? area * (bta * (1.0 / 10.0) + (bs - bta) * (1.0 / 30.0))
				                            : area * ((bta + btb) * (1.0 / 30.0) + (bs - bta - btb) * (1.0 / 60.0));
				const double t = P.rs ? br[b] * inz + (P.picard ? 0.0 : cx * uu[b].x + cy * uu[b].y) : 0.0;
				if (!P.picard) { Fx += Kab * uu[b].x; Fy += Kab * uu[b].y; }
				if (P.rs) { Fx += Kab * cx * t; Fy += Kab * cy * t; }
				if (WANT_J) {
					int lva = 0;
					#pragma unroll
					for (int j = 0; j < 3; ++j) if (j == a) lva = la[j];
					const int slot = tet_blk[(lva * 4 + la[b]) * nt + cell];
					if (slot >= 0) {
						double2* v2 = (double2*)val + 2 * (size_t)slot;
						double2 lo = v2[0], hi = v2[1];
						lo.x += Kab * (1.0 + cx * cx); hi.y += Kab * (1.0 + cy * cy);
						if (P.rs) { lo.y += Kab * cx * cy; hi.x += Kab * cx * cy; }
						v2[0] = lo; v2[1] = hi;
					}
				}
			}
		}
		if (press && WANT_F) {
			double fw[3], fwa = 0.0;
			#pragma unroll
			for (int j = 0; j < 3; ++j) { fw[j] = P.rho_g * (p[j].w - p[j].z) - P.rho_sw_g * (-fmin(0.0, p[j].z)); if (j == a) fwa = fw[j]; }
			const double mm = area * (fw[0] + fw[1] + fw[2] + fwa) * (1.0 / 12.0);
			Fx -= nx * mm; Fy -= ny * mm;
		}
	}
	if (WANT_F) { F[2 * i] += Fx; F[2 * i + 1] += Fy; }
}

static AsmParams make_params(const bp_ctx* c)
{
	AsmParams P;
	P.n = c->prm.n; P.eps_reg = c->prm.eps_reg;
	P.rho_g = c->prm.rho_i * c->prm.g; P.rho_sw_g = c->prm.rho_sw * c->prm.g;
	P.n_quad = (int)c->quad_wts.size();
	P.n_is_3 = (c->prm.n == 3.0);
	P.picard = 0;
	P.rs = c->prm.reduced_stokes ? 1 : 0;
	return P;
}

static AsmParams make_params_fwd(const bp_ctx* c) { return make_params(c); }
AsmParams bp_asm_params(const bp_ctx* c) { return make_params(c); }

void bp_upload_quadrature(bp_ctx* c)
{
	cudaMemcpyToSymbol(c_qw, c->quad_wts.data(), c->quad_wts.size() * sizeof(double));
	cudaMemcpyToSymbol(c_ql, c->quad_pts.data(), c->quad_pts.size() * sizeof(double));
	cudaMemcpyToSymbol(c_qw2, c->quad_wts_aux.data(), c->quad_wts_aux.size() * sizeof(double));
	cudaMemcpyToSymbol(c_ql2, c->quad_pts_aux.data(), c->quad_pts_aux.size() * sizeof(double));
}

// u per vertex for the row-gather kernel (periodic images read their master's value)
__global__ void k_expand_u(int nv, const int32_t* __restrict__ v2n, const double2* __restrict__ u, double2* __restrict__ uv)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) uv[i] = u[v2n[i]];
}
static inline int red_blocks_fwd(int64_t n_items, int per_block);
static const double2* bp_vertex_u(bp_ctx* c)
{
	if (!c->d_v2n) return (const double2*)c->d_u;
	k_expand_u<<<red_blocks_fwd(c->nv, 256), 256, 0, c->stream>>>((int)c->nv, c->d_v2n, (const double2*)c->d_u, (double2*)c->d_uvert);
	c->launches++;
	return (const double2*)c->d_uvert;
}

void bp_launch_expand_to_vertex(bp_ctx* c, const double* node_vec2, double* vertex_vec2)
{
	if (!c->d_v2n) { cudaMemcpyAsync(vertex_vec2, node_vec2, (size_t)2 * c->nv * sizeof(double), cudaMemcpyDeviceToDevice, c->stream); return; }
	k_expand_u<<<red_blocks_fwd(c->nv, 256), 256, 0, c->stream>>>((int)c->nv, c->d_v2n, (const double2*)node_vec2, (double2*)vertex_vec2);
	c->launches++;
}

void bp_launch_rows_aux(bp_ctx* c, int kind, const double2* u_eta_vertex, const double* uz_vertex)
{
	const AsmParams P = make_params(c);
	const int gr = (int)((c->nn_owned + RTPB - 1) / RTPB);
	const size_t sm = (size_t)c->max_row * RTPB * sizeof(double);
	const double2* uvert = bp_vertex_u(c);
	#define AUX(K) k_rows_aux<K><<<gr, RTPB, sm, c->stream>>>((int)c->nn_owned, c->d_inc_slice, c->d_inc_code, c->d_inc_pos, c->d_inc_v, (size_t)c->n_inc_slots, \
			c->d_soa, (size_t)c->nv, uvert, c->d_A, c->d_sell_ptr, c->d_rowlen, c->d_F, c->d_val, P, (int)c->quad_wts_aux.size(), \
			c->d_Bv, c->d_Bring, (const double2*)c->d_aux, c->d_bcnode, u_eta_vertex, uz_vertex)
	if (kind == 0) AUX(0); else if (kind == 1) AUX(1); else AUX(2);
	c->launches++;
}
__global__ void k_set_node_x(int n, const int32_t* __restrict__ nodes, const double* __restrict__ vals, double2* __restrict__ vec)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) vec[nodes[i]].x = vals[i];
}
// the lateral Dirichlet values of the u_z system replace the projected basal value on their nodes (bp_set_dirichlet_uz)
void bp_launch_uz_lateral_values(bp_ctx* c, double* node_vec)
{
	if (!c->n_uzbc) return;
	k_set_node_x<<<(int)((c->n_uzbc + 255) / 256), 256, 0, c->stream>>>((int)c->n_uzbc, c->d_uzbc_node, c->d_uzbc_val, (double2*)node_vec);
	c->launches++;
}
void bp_launch_node_to_vertex(bp_ctx* c, const double* node_vec, int comp, double* vertex_out)
{
	k_node_to_vertex<<<red_blocks_fwd(c->nv, 256), 256, 0, c->stream>>>((int)c->nv, c->d_v2n, node_vec, comp, vertex_out);
	c->launches++;
}

void bp_launch_proxy(bp_ctx* c);
void bp_update_ainv(bp_ctx* c)
{
	if (!c->ainv_dirty) return;
	const AsmParams P = make_params(c);
	RateLaw L = { c->prm.a_T_l, c->prm.a_T_u, c->prm.Q_T_l, c->prm.Q_T_u, c->prm.R_gas };
	const int g = (int)((c->nt + 127) / 128);
	if (c->prm.energy_dependent) k_tet_ainv<true><<<g, 128, 0, c->stream>>>((int)c->nt, c->d_tets, c->d_A, c->d_Tp, c->d_W, c->d_E, L, P, c->d_ainv);
	else k_tet_ainv<false><<<g, 128, 0, c->stream>>>((int)c->nt, c->d_tets, c->d_A, c->d_Tp, c->d_W, c->d_E, L, P, c->d_ainv);
	c->launches++;
	c->ainv_dirty = false;
	bp_tiles_update_ainv(c);
}

__global__ void k_expand_u_range(int v0, int v1, const int32_t* __restrict__ v2n, const double2* __restrict__ u, double2* __restrict__ uv)
{
	for (int i = v0 + blockIdx.x * blockDim.x + threadIdx.x; i < v1; i += gridDim.x * blockDim.x) uv[i] = u[v2n[i]];
}

bool bp_rows_gather_mode(const bp_ctx* c)
{
	const size_t row_smem = (size_t)4 * std::max(c->max_row - 1, 1) * RTPB * sizeof(double);
	return c->prm.assembly == BP_ASM_GATHER || (c->prm.assembly == BP_ASM_AUTO && row_smem <= 200 * 1024);
}

// one slice of the row-gather pass (the pipelined host-buffer bp_assemble): CTAs [blk0, blk0 + nblk)
void bp_launch_rows_chunk(bp_ctx* c, int what, int blk0, int nblk, int v0, int v1)
{
	AsmParams P = make_params(c);
	const bool wf = what & BP_ASSEMBLE_F, wj = what & BP_ASSEMBLE_J;
	const size_t row_smem = (size_t)4 * std::max(c->max_row - 1, 1) * RTPB * sizeof(double);
	if (!c->smem_attr_set) {
		cudaFuncSetAttribute(k_rows<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		cudaFuncSetAttribute(k_rows<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		cudaFuncSetAttribute(k_rows<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		cudaFuncSetAttribute(k_rows<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
		c->smem_attr_set = true;
	}
	const double2* uvert = (const double2*)c->d_u;
	if (c->d_v2n) {
		if (v1 > v0) { k_expand_u_range<<<red_blocks_fwd(v1 - v0, 256), 256, 0, c->stream>>>(v0, v1, c->d_v2n, (const double2*)c->d_u, (double2*)c->d_uvert); c->launches++; }
		uvert = (const double2*)c->d_uvert;
	}
	if (nblk <= 0) return;
	#define ROWSC(WF, WJ, SM) do { if (P.rs) k_rows<WF, WJ, true><<<nblk, RTPB, SM, c->stream>>>((int)c->nn_owned, c->d_inc_slice, c->d_inc_code, c->d_inc_pos, c->d_inc_v, (size_t)c->n_inc_slots, c->d_soa, (size_t)c->nv, uvert, c->d_ainv, c->d_sell_ptr, c->d_rowlen, c->d_F, c->d_val, P, c->d_uz, blk0, c->d_mirror); \
		else k_rows<WF, WJ, false><<<nblk, RTPB, SM, c->stream>>>((int)c->nn_owned, c->d_inc_slice, c->d_inc_code, c->d_inc_pos, c->d_inc_v, (size_t)c->n_inc_slots, c->d_soa, (size_t)c->nv, uvert, c->d_ainv, c->d_sell_ptr, c->d_rowlen, c->d_F, c->d_val, P, c->d_uz, blk0, c->d_mirror); } while (0)
	if (wf && wj) ROWSC(true, true, row_smem);
	else if (wf)  ROWSC(true, false, 0);
	else if (wj)  ROWSC(false, true, row_smem);
	c->launches++;
}

void bp_launch_facets(bp_ctx* c, int what)
{
	if (!c->nf) return;
	AsmParams P = make_params(c);
	const bool wf = what & BP_ASSEMBLE_F, wj = what & BP_ASSEMBLE_J;
	const int gf = (int)((c->nf + TPB - 1) / TPB);
	int64_t nfn = c->n_fnodes;
	const int32_t *fnn = c->d_fn_node, *fnp = c->d_fn_ptr, *fni = c->d_fn_inc;
	const bool folded = bp_tiles_serve(c, what) && bp_tiles_facet_lists(c, &nfn, &fnn, &fnp, &fni);   // the tiled kernel did the basal facets
	if (folded && nfn == 0) return;
	const int gfr = (int)((nfn + TPB - 1) / TPB);
	const bool use_fr = nfn > 0 && (folded || !c->env.facet_atomic);
	const double2* u2 = (const double2*)c->d_u;
	#define FACETC(WF, WJ) do { if (use_fr) k_facet_rows<WF, WJ><<<gfr, TPB, 0, c->stream>>>((int)nfn, fnn, fnp, fni, (int)c->nt, c->d_fac_cell, c->d_fac_info, c->d_tets, c->d_geo, c->d_v2n, u2, c->d_beta, c->d_tet_blk, c->d_F, c->d_val, P, c->d_Bring); \
		else k_facet<WF, WJ><<<gf, TPB, 0, c->stream>>>((int)c->nf, (int)c->nt, c->d_fac_cell, c->d_fac_info, c->d_tets, c->d_geo, c->d_v2n, u2, c->d_beta, c->d_tet_blk, (int)c->nn_owned, c->d_F, c->d_val, P, c->d_Bring); } while (0)
	if (wf && wj) FACETC(true, true);
	else if (wf)  FACETC(true, false);
	else if (wj)  FACETC(false, true);
	c->launches++;
}

void bp_launch_assemble(bp_ctx* c, int what)
{
	AsmParams P = make_params(c);
	P.picard = (what & 16) ? 1 : 0;
	bp_update_ainv(c);
	const bool wf = what & BP_ASSEMBLE_F, wj = what & BP_ASSEMBLE_J;
#ifdef BP_EXPERIMENTS
	if (what & 8) { bp_launch_proxy(c); return; }
#endif
	if ((what & BP_TIME_CELL_ONLY) && bp_tiles_serve(c, what)) { bp_launch_tiles(c, what); c->have_J = false; return; }
	if (what & BP_TIME_CELL_ONLY) {        // timing aid: the dominant kernel alone (accumulates into stale values)
		const size_t rs = (size_t)4 * std::max(c->max_row - 1, 1) * RTPB * sizeof(double);
		if (c->prm.assembly == BP_ASM_GATHER || (c->prm.assembly == BP_ASM_AUTO && rs <= 200 * 1024)) {
			cudaFuncSetAttribute(k_rows<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
			k_rows<true, true, false><<<(int)((c->nn_owned + RTPB - 1) / RTPB), RTPB, rs, c->stream>>>((int)c->nn_owned, c->d_inc_slice, c->d_inc_code, c->d_inc_pos,
				c->d_inc_v, (size_t)c->n_inc_slots, c->d_soa, (size_t)c->nv, bp_vertex_u(c), c->d_ainv, c->d_sell_ptr, c->d_rowlen, c->d_F, c->d_val, P, c->d_uz, 0, c->d_mirror);
		} else {
			const int gc = (int)((c->nt + TPB - 1) / TPB);
			k_cell<true, true><<<gc, TPB, 0, c->stream>>>((int)c->nt, c->d_tets, c->d_geo, c->d_v2n, (const double2*)c->d_u, c->d_ainv, c->d_tet_blk, (int)c->nn_owned, c->d_F, c->d_val, P, c->d_uz);
		}
		c->launches++;
		c->have_J = false;
		return;
	}
	const int gf = (int)((c->nf + TPB - 1) / TPB);
	int64_t nfn = c->n_fnodes;
	const int32_t *fnn = c->d_fn_node, *fnp = c->d_fn_ptr, *fni = c->d_fn_inc;
	const bool folded = bp_tiles_serve(c, what) && bp_tiles_facet_lists(c, &nfn, &fnn, &fnp, &fni);   // the tiled kernel does the basal facets
	const int gfr = (int)((nfn + TPB - 1) / TPB);
	const double2* u2 = (const double2*)c->d_u;
	const size_t row_smem = (size_t)4 * std::max(c->max_row - 1, 1) * RTPB * sizeof(double);   // the diagonal block lives in registers
	const bool gather = c->prm.assembly == BP_ASM_GATHER || (c->prm.assembly == BP_ASM_AUTO && row_smem <= 200 * 1024);
	// default: the atomic-free facet kernel (bitwise reproducible F and J); BP_FACET_ATOMIC=1 gives the atomic one back
	const bool use_fr = folded || ((gather || bp_tiles_serve(c, what)) && nfn > 0 && !c->env.facet_atomic);
	#define FACET(WF, WJ) do { if (use_fr) k_facet_rows<WF, WJ><<<gfr, TPB, 0, c->stream>>>((int)nfn, fnn, fnp, fni, (int)c->nt, c->d_fac_cell, c->d_fac_info, c->d_tets, c->d_geo, c->d_v2n, u2, c->d_beta, c->d_tet_blk, c->d_F, c->d_val, P, c->d_Bring); \
		else k_facet<WF, WJ><<<gf, TPB, 0, c->stream>>>((int)c->nf, (int)c->nt, c->d_fac_cell, c->d_fac_info, c->d_tets, c->d_geo, c->d_v2n, u2, c->d_beta, c->d_tet_blk, (int)c->nn_owned, c->d_F, c->d_val, P, c->d_Bring); } while (0)
	const bool tiled = bp_tiles_serve(c, what);
	if (tiled) {
		bp_launch_tiles(c, what);
		c->launches--;                               // counted below
	} else if (gather) {
		if (!c->smem_attr_set) {                     // function attributes are per device
			cudaFuncSetAttribute(k_rows<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
			cudaFuncSetAttribute(k_rows<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
			cudaFuncSetAttribute(k_rows<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
			cudaFuncSetAttribute(k_rows<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
			c->smem_attr_set = true;
		}
		const int gr = (int)((c->nn_owned + RTPB - 1) / RTPB);
		const double2* uvert = bp_vertex_u(c);
		#define ROWS(WF, WJ, SM) do { if (P.rs) k_rows<WF, WJ, true><<<gr, RTPB, SM, c->stream>>>((int)c->nn_owned, c->d_inc_slice, c->d_inc_code, c->d_inc_pos, c->d_inc_v, (size_t)c->n_inc_slots, c->d_soa, (size_t)c->nv, uvert, c->d_ainv, c->d_sell_ptr, c->d_rowlen, c->d_F, c->d_val, P, c->d_uz, 0, c->d_mirror); \
			else k_rows<WF, WJ, false><<<gr, RTPB, SM, c->stream>>>((int)c->nn_owned, c->d_inc_slice, c->d_inc_code, c->d_inc_pos, c->d_inc_v, (size_t)c->n_inc_slots, c->d_soa, (size_t)c->nv, uvert, c->d_ainv, c->d_sell_ptr, c->d_rowlen, c->d_F, c->d_val, P, c->d_uz, 0, c->d_mirror); } while (0)
		if (wf && wj) ROWS(true, true, row_smem);
		else if (wf)  ROWS(true, false, 0);
		else if (wj)  ROWS(false, true, row_smem);
	} else {
		if (wf) cudaMemsetAsync(c->d_F, 0, 2 * c->nn * sizeof(double), c->stream);
		if (wj) cudaMemsetAsync(c->d_val, 0, (size_t)c->n_slots * 4 * sizeof(double), c->stream);
		const int gc = (int)((c->nt + TPB - 1) / TPB);
		#define CELL(WF, WJ) k_cell<WF, WJ><<<gc, TPB, 0, c->stream>>>((int)c->nt, c->d_tets, c->d_geo, c->d_v2n, u2, c->d_ainv, c->d_tet_blk, (int)c->nn_owned, c->d_F, c->d_val, P, c->d_uz)
		if (wf && wj) CELL(true, true);
		else if (wf)  CELL(true, false);
		else if (wj)  CELL(false, true);
	}
	const bool run_facets = c->nf && !(folded && nfn == 0);
	if (run_facets) {
		if (wf && wj) FACET(true, true);
		else if (wf)  FACET(true, false);
		else if (wj)  FACET(false, true);
	}
	c->launches += 1 + (run_facets ? 1 : 0);
	if (wj) c->have_J = true;
}

// ---------------------------------------------------------------- reductions
// Block partials -> d_part[slot][block]; the last block to finish sums them in a
// fixed order (deterministic) and writes d_sc[slot...].
template <int NV>
__device__ __forceinline__ void block_reduce_finalize(double (&v)[NV], double* part, int pstride,
                                                      unsigned* counter, double* out)
{
	__shared__ double sm[NV][32];
	__shared__ bool last;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
	#pragma unroll
	for (int i = 0; i < NV; ++i) {
		double x = v[i];
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
		if (lane == 0) sm[i][w] = x;
	}
	__syncthreads();
	if (w == 0) {
		#pragma unroll
		for (int i = 0; i < NV; ++i) {
			double x = (lane < nw) ? sm[i][lane] : 0.0;
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
			if (lane == 0) part[i * pstride + blockIdx.x] = x;
		}
	}
	if (threadIdx.x == 0) {
		__threadfence();
		const unsigned ticket = atomicAdd(counter, 1u);
		last = (ticket == gridDim.x - 1);
	}
	__syncthreads();
	if (last) {
		__threadfence();
		#pragma unroll
		for (int i = 0; i < NV; ++i) {
			double x = 0.0;
			for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) x += ((volatile double*)part)[i * pstride + b];
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
			__syncthreads();
			if (lane == 0) sm[i][w] = x;
			__syncthreads();
			if (w == 0) {
				x = (lane < nw) ? sm[i][lane] : 0.0;
				#pragma unroll
				for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
				if (lane == 0) out[i] = x;
			}
		}
		if (threadIdx.x == 0) *counter = 0u;
	}
}

#define RED_BLOCKS_MAX 4096
static inline int red_blocks(int64_t n_items, int per_block);
static inline int red_blocks_fwd(int64_t n_items, int per_block) { return red_blocks(n_items, per_block); }
static inline int red_blocks(int64_t n_items, int per_block)
{
	int64_t b = (n_items + per_block - 1) / per_block;
	if (b < 1) b = 1;
	if (b > 148 * 8) b = 148 * 8;
	return (int)b;
}

// ---------------------------------------------------------------- SpMV (2x2 blocks, SELL-32)
// One thread per block row; block k of the 32 rows of a slice is contiguous, so every
// warp load of values (2 x 16 B per lane) and of column ids is a full-line access; the
// {x,y} pair of the column node is a 16 B gather.  No shuffles, no row pointer chase.
template <bool WITH_DOT>
__global__ void __launch_bounds__(256)
k_spmv(int nrows, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ col,
       const double2* __restrict__ val, const double2* __restrict__ x, double2* __restrict__ y,
       const double* __restrict__ sc, double* part, unsigned* counter, double* out)
{
	if (WITH_DOT && sc[SC_DONE] != 0.0) return;
	double acc[1] = { 0.0 };
	const int lane = threadIdx.x & 31;
	const int nsl = (nrows + 31) >> 5;
	for (int slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slice < nsl; slice += (gridDim.x * blockDim.x) >> 5) {
		const int b0 = sell_ptr[slice], w = (sell_ptr[slice + 1] - b0) >> 5;
		const int row = (slice << 5) + lane;
		const int32_t* cp = col + b0 + lane;
		const double2* vp = val + 2 * ((size_t)b0 + lane);
		double s0 = 0.0, s1 = 0.0;
		int k = 0;
		for (; k + 4 <= w; k += 4) {
			int cc[4]; double2 v01[4], v23[4], xv[4];
			#pragma unroll
			for (int j = 0; j < 4; ++j) { cc[j] = cp[32 * (k + j)]; v01[j] = vp[64 * (size_t)(k + j)]; v23[j] = vp[64 * (size_t)(k + j) + 1]; }
			#pragma unroll
			for (int j = 0; j < 4; ++j) xv[j] = x[cc[j]];
			#pragma unroll
			for (int j = 0; j < 4; ++j) { s0 += v01[j].x * xv[j].x + v01[j].y * xv[j].y; s1 += v23[j].x * xv[j].x + v23[j].y * xv[j].y; }
		}
		for (; k < w; ++k) {
			const double2 a01 = vp[64 * (size_t)k], a23 = vp[64 * (size_t)k + 1];
			const double2 xv = x[cp[32 * k]];
			s0 += a01.x * xv.x + a01.y * xv.y; s1 += a23.x * xv.x + a23.y * xv.y;
		}
		if (row < nrows) {
			y[row] = make_double2(s0, s1);
			if (WITH_DOT) { const double2 xr = x[row]; acc[0] += xr.x * s0 + xr.y * s1; }
		}
	}
	if (WITH_DOT) block_reduce_finalize<1>(acc, part, RED_BLOCKS_MAX, counter, out);
}

// grid = one wave of resident blocks (a second, partial wave costs a full block time)
template <class K>
static int one_wave(K kernel, int block, int64_t n_items, int per_block)
{
	int per_sm = 0, dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0);
	int64_t b = (n_items + per_block - 1) / per_block;
	const int64_t cap = (int64_t)std::max(per_sm, 1) * sms;
	if (b > cap) b = cap;
	if (b > RED_BLOCKS_MAX) b = RED_BLOCKS_MAX;
	return (int)std::max<int64_t>(b, 1);
}

void bp_launch_spmv(bp_ctx* c, const double* x, double* y, int with_dot)
{
	const int rows_per_block = 256;
	if (!c->grid_spmv0) {
		c->grid_spmv1 = one_wave(k_spmv<true>, 256, c->nn_owned, rows_per_block);
		c->grid_spmv0 = one_wave(k_spmv<false>, 256, c->nn_owned, rows_per_block);
	}
	const int g1 = c->grid_spmv1, g0 = c->grid_spmv0;
	if (with_dot)
		k_spmv<true><<<g1, 256, 0, c->stream>>>((int)c->nn_owned, c->d_sell_ptr, c->d_col, (const double2*)c->d_val,
			(const double2*)x, (double2*)y, c->d_sc, c->d_part, c->d_counter, c->d_sc + SC_PAP);
	else
		k_spmv<false><<<g0, 256, 0, c->stream>>>((int)c->nn_owned, c->d_sell_ptr, c->d_col, (const double2*)c->d_val,
			(const double2*)x, (double2*)y, c->d_sc, c->d_part, c->d_counter, c->d_sc + SC_PAP);
	c->launches++;
}

// ---------------------------------------------------------------- Dirichlet rows (momentumbp.py:101-112)
// dolfin's NewtonSolver applies a DirichletBC to the update (SURVEY A-N): on a constrained dof b_i = u_i - g_i and the
// row of J is an identity row, so the Newton step carries u_i to g_i.  The constrained dofs are eliminated
// symmetrically here (columns too, the free rows lifted by J d) -- the same solution, and CG keeps an SPD matrix.
// d = what the solve must return on the constrained dofs: u - g (Newton: sign = 1) or -g (linear=True: the solve
// returns -u, sign = 0); zero elsewhere.  All local dofs, ghosts included (their columns are eliminated as well).
__global__ void k_bc_d(int n2, const double* __restrict__ mask, const double* __restrict__ g, const double* __restrict__ u, double sign, double* __restrict__ d)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
		d[i] = mask[i] != 0.0 ? sign * u[i] - g[i] : 0.0;
}
// constrained rows of the right-hand side (owned dofs)
__global__ void k_bc_rows(int n2, const double* __restrict__ mask, const double* __restrict__ d, double* __restrict__ F)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
		if (mask[i] != 0.0) F[i] = d[i];
}
// free rows: F -= J d (Jd computed with the unconstrained J)
__global__ void k_bc_lift(int n2, const double* __restrict__ mask, const double* __restrict__ Jd, double* __restrict__ F)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
		if (mask[i] == 0.0) F[i] -= Jd[i];
}
// rows and columns of the constrained dofs -> identity (one thread per owned block row)
__global__ void k_bc_matrix(int n, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ rowlen, const int32_t* __restrict__ col,
                            const double* __restrict__ mask, double* __restrict__ val)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const bool r0 = mask[2 * i] != 0.0, r1 = mask[2 * i + 1] != 0.0;
		const int base = sell_ptr[i >> 5] + (i & 31);
		for (int k = 0; k < rowlen[i]; ++k) {
			const int s = base + 32 * k, j = col[s];
			const bool c0 = mask[2 * j] != 0.0, c1 = mask[2 * j + 1] != 0.0;
			if (!(r0 | r1 | c0 | c1)) continue;
			double* v = val + 4 * (size_t)s;
			if (r0 | c0) v[0] = (j == i && r0) ? 1.0 : 0.0;
			if (r0 | c1) v[1] = 0.0;
			if (r1 | c0) v[2] = 0.0;
			if (r1 | c1) v[3] = (j == i && r1) ? 1.0 : 0.0;
		}
	}
}
void bp_launch_bc_rows(bp_ctx* c, double sign)
{
	const int n2a = (int)(2 * c->nn), n2o = (int)(2 * c->nn_owned);
	k_bc_d<<<red_blocks(n2a, 256), 256, 0, c->stream>>>(n2a, c->d_bcmask, c->d_bcval, c->d_u, sign, c->d_io);
	k_bc_rows<<<red_blocks(n2o, 256), 256, 0, c->stream>>>(n2o, c->d_bcmask, c->d_io, c->d_F);
	c->launches += 2;
}
void bp_launch_bc_eliminate(bp_ctx* c)
{
	const int n2o = (int)(2 * c->nn_owned), no = (int)c->nn_owned;
	bp_launch_spmv(c, c->d_io, c->d_io2, 0);                 // J d with the unconstrained J (d: bp_launch_bc_rows)
	k_bc_lift<<<red_blocks(n2o, 256), 256, 0, c->stream>>>(n2o, c->d_bcmask, c->d_io2, c->d_F);
	k_bc_matrix<<<red_blocks(no, 128), 128, 0, c->stream>>>(no, c->d_sell_ptr, c->d_rowlen, c->d_col, c->d_bcmask, c->d_val);
	c->launches += 2;
}

// ---------------------------------------------------------------- vector glue
__global__ void k_gather_in(int n2, const int32_t* __restrict__ ext_dof, const double* __restrict__ ext, double* __restrict__ in)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
		in[i] = ext[ext_dof ? ext_dof[i] : i];
}
__global__ void k_scatter_out(int n2, const int32_t* __restrict__ ext_dof, const double* __restrict__ in, double* __restrict__ ext)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x)
		ext[ext_dof ? ext_dof[i] : i] = in[i];
}
void bp_launch_gather_in(bp_ctx* c, const double* ext, double* internal)
{
	const int n2 = (int)(2 * c->nn);
	k_gather_in<<<red_blocks(n2, 256), 256, 0, c->stream>>>(n2, c->d_ext_dof, ext, internal);
	c->launches++;
}
void bp_launch_scatter_out(bp_ctx* c, const double* internal, double* ext)
{
	const int n2 = (int)(2 * c->nn);
	k_scatter_out<<<red_blocks(n2, 256), 256, 0, c->stream>>>(n2, c->d_ext_dof, internal, ext);
	c->launches++;
}

__global__ void k_csr_export(int64_t nnz, const int32_t* __restrict__ src, const double* __restrict__ val, double* __restrict__ out)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
		out[i] = val[src[i]];
}
void bp_launch_csr_export(bp_ctx* c, double* out)
{
	const int64_t nnz = c->csr_indptr.back();
	k_csr_export<<<red_blocks(nnz, 256), 256, 0, c->stream>>>(nnz, c->d_csr_src, c->d_val, out);
	c->launches++;
}

__global__ void k_set_S(int nv, const double* __restrict__ S, double4* geo, double* soa_S, double2* zS)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) { geo[i].w = S[i]; soa_S[i] = S[i]; zS[i].y = S[i]; }
}
void bp_launch_set_S(bp_ctx* c, const double* S_dev)
{
	k_set_S<<<red_blocks(c->nv, 256), 256, 0, c->stream>>>((int)c->nv, S_dev, c->d_geo, c->d_soa + 3 * (size_t)c->nv, c->d_zS);
	c->launches++;
}

__global__ void k_pack_halo(int n, const int32_t* __restrict__ nodes, const double2* __restrict__ vec, double2* __restrict__ buf)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) buf[i] = vec[nodes[i]];
}
void bp_launch_pack_halo(bp_ctx* c, const double* vec)
{
	if (c->n_send == 0) return;
	k_pack_halo<<<red_blocks(c->n_send, 256), 256, 0, c->stream>>>((int)c->n_send, c->d_send_nodes, (const double2*)vec, (double2*)c->d_sendbuf);
	c->launches++;
}

// dot over the owned entries -> d_sc[slot]
__global__ void __launch_bounds__(256)
k_dot(int n, const double* __restrict__ a, const double* __restrict__ b, double* part, unsigned* counter, double* out)
{
	double acc[1] = { 0.0 };
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc[0] += a[i] * b[i];
	block_reduce_finalize<1>(acc, part, RED_BLOCKS_MAX, counter, out);
}
void bp_launch_dot_owned(bp_ctx* c, const double* a, const double* b, int sc_slot)
{
	const int n = (int)(2 * c->nn_owned);
	k_dot<<<red_blocks(n, 1024), 256, 0, c->stream>>>(n, a, b, c->d_part, c->d_counter, c->d_sc + sc_slot);
	c->launches++;
}

__global__ void k_axpy(int n, double alpha, const double* __restrict__ x, double* __restrict__ y)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] += alpha * x[i];
}
void bp_launch_axpy(bp_ctx* c, double alpha, const double* x, double* y)
{
	const int n = (int)(2 * c->nn_owned);
	k_axpy<<<red_blocks(n, 1024), 256, 0, c->stream>>>(n, alpha, x, y);
	c->launches++;
}

// ---------------------------------------------------------------- preconditioner
// dinv[row] = inverse of the 2x2 diagonal block (BLOCK_JACOBI), or diag(1/a00, 1/a11)
// (JACOBI), or identity (NONE).
__global__ void k_pc_setup(int nrows, int kind, const int32_t* __restrict__ diag, const double* __restrict__ val, double* __restrict__ dinv)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) {
		double a = 1.0, b = 0.0, cc = 0.0, d = 1.0;
		if (kind != BP_PC_NONE && diag[i] >= 0) {
			const double* v = val + 4 * (size_t)diag[i];
			if (kind == BP_PC_JACOBI) { a = 1.0 / v[0]; d = 1.0 / v[3]; }
			else {
				const double idet = 1.0 / (v[0] * v[3] - v[1] * v[2]);
				a = v[3] * idet; b = -v[1] * idet; cc = -v[2] * idet; d = v[0] * idet;
			}
		}
		dinv[4 * (size_t)i] = a; dinv[4 * (size_t)i + 1] = b; dinv[4 * (size_t)i + 2] = cc; dinv[4 * (size_t)i + 3] = d;
	}
}
void bp_launch_pc_setup(bp_ctx* c)
{
	int kind = c->prm.pc;
	if (kind != BP_PC_NONE && kind != BP_PC_JACOBI) kind = BP_PC_BLOCK_JACOBI;
	k_pc_setup<<<red_blocks(c->nn_owned, 256), 256, 0, c->stream>>>((int)c->nn_owned, kind, c->d_diag, c->d_val, c->d_dinv);
	c->launches++;
}

// ---------------------------------------------------------------- PCG kernels
// KSPCG-like: zero initial guess; norm = |r|_2 (default) or PETSc's preconditioned |z|_2,
// converged when ||z|| <= max(rtol * ||z_0||, atol).
__device__ __forceinline__ double2 apply_dinv(const double* __restrict__ dinv, int i, double2 r)
{
	const double2 d01 = ((const double2*)dinv)[2 * (size_t)i], d23 = ((const double2*)dinv)[2 * (size_t)i + 1];
	return make_double2(d01.x * r.x + d01.y * r.y, d23.x * r.x + d23.y * r.y);
}

// Scalars: iteration k (parity par = k & 1) reads r.z from triple T[par] and writes the
// new {r.z, z.z, r.r} into T[par ^ 1]; the init writes T[0].  No kernel rotates scalars.
// x = 0, r = b, z = M^-1 r, p = z;  (rz, zz, rr) -> T[0]
__global__ void __launch_bounds__(256)
k_pcg_init(int nrows, const double2* __restrict__ b, const double* __restrict__ dinv,
           double2* x, double2* r, double2* z, double2* p, double* part, unsigned* counter, double* sc)
{
	double acc[3] = { 0.0, 0.0, 0.0 };
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) {
		const double2 rv = b[i];
		const double2 zv = apply_dinv(dinv, i, rv);
		x[i] = make_double2(0.0, 0.0); r[i] = rv; z[i] = zv; p[i] = zv;
		acc[0] += rv.x * zv.x + rv.y * zv.y; acc[1] += zv.x * zv.x + zv.y * zv.y; acc[2] += rv.x * rv.x + rv.y * rv.y;
	}
	block_reduce_finalize<3>(acc, part, RED_BLOCKS_MAX, counter, sc + SC_T0);
}

// alpha = rz / pAp;  x += alpha p;  r -= alpha Ap;  z = M^-1 r;  (rz, zz, rr) -> T[par^1]
__global__ void __launch_bounds__(256)
k_pcg_update(int nrows, int par, const double* __restrict__ dinv, const double2* __restrict__ p, const double2* __restrict__ Ap,
             double2* x, double2* r, double2* z, double* part, unsigned* counter, double* sc)
{
	if (sc[SC_DONE] != 0.0) return;
	const double pAp = sc[SC_PAP];
	if (!(pAp > 0.0)) return;                        // breakdown: k_pcg_p flags it, x keeps the last good iterate
	const double alpha = sc[par ? SC_T1 : SC_T0] / pAp;
	double acc[3] = { 0.0, 0.0, 0.0 };
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) {
		const double2 pv = p[i], av = Ap[i];
		double2 xv = x[i], rv = r[i];
		xv.x += alpha * pv.x; xv.y += alpha * pv.y;
		rv.x -= alpha * av.x; rv.y -= alpha * av.y;
		const double2 zv = apply_dinv(dinv, i, rv);
		x[i] = xv; r[i] = rv; z[i] = zv;
		acc[0] += rv.x * zv.x + rv.y * zv.y; acc[1] += zv.x * zv.x + zv.y * zv.y; acc[2] += rv.x * rv.x + rv.y * rv.y;
	}
	block_reduce_finalize<3>(acc, part, RED_BLOCKS_MAX, counter, sc + (par ? SC_T0 : SC_T1));
}

// convergence test + p = z + beta p.  One designated thread does the bookkeeping; it is
// the only writer of SC_DONE, and blocks that already see it set skip the same work the
// others skip because they evaluate the same predicate.
__global__ void __launch_bounds__(256)
k_pcg_p(int nrows, int first, int par, double rtol, double atol, int precond_norm, const double2* __restrict__ z, double2* p, double* sc)
{
	if (!first && sc[SC_DONE] != 0.0) return;
	const double* Told = sc + (par ? SC_T1 : SC_T0);
	const double* Tnew = first ? sc + SC_T0 : sc + (par ? SC_T0 : SC_T1);
	const double rz = Told[0], rz_next = Tnew[0], pAp = sc[SC_PAP];
	const double norm = sqrt(precond_norm ? Tnew[1] : Tnew[2]);        // |M^-1 r| or |r|
	const double norm0 = first ? norm : sc[SC_NORM0];
	const bool bad = !first && !(pAp > 0.0);
	const bool conv = bad || !(norm > fmax(rtol * norm0, atol));     // also stops on NaN
	const double beta = first ? 0.0 : rz_next / rz;
	if (!conv && !first)
		for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) {
			const double2 zv = z[i], pv = p[i];
			p[i] = make_double2(zv.x + beta * pv.x, zv.y + beta * pv.y);
		}
	if (blockIdx.x == gridDim.x - 1 && threadIdx.x == blockDim.x - 1) {
		sc[SC_NORM] = bad ? sc[SC_NORM] : norm;
		if (first) { sc[SC_NORM0] = norm; sc[SC_ITERS] = 0.0; sc[SC_BREAKDOWN] = 0.0; }
		else if (!bad) sc[SC_ITERS] += 1.0;
		if (bad) sc[SC_BREAKDOWN] = 1.0;
		if (conv) sc[SC_DONE] = 1.0;
	}
}

void bp_launch_pcg_init(bp_ctx* c)
{
	const int n = (int)c->nn_owned;
	cudaMemsetAsync(c->d_sc + SC_DONE, 0, sizeof(double), c->stream);
	k_pcg_init<<<red_blocks(n, 512), 256, 0, c->stream>>>(n, (const double2*)c->d_F, c->d_dinv, (double2*)c->d_dx, (double2*)c->d_r,
		(double2*)c->d_z, (double2*)c->d_p, c->d_part, c->d_counter, c->d_sc);
	c->launches++;
}
void bp_launch_pcg_update(bp_ctx* c, int par)
{
	const int n = (int)c->nn_owned;
	k_pcg_update<<<red_blocks(n, 512), 256, 0, c->stream>>>(n, par, c->d_dinv, (const double2*)c->d_p, (const double2*)c->d_Ap,
		(double2*)c->d_dx, (double2*)c->d_r, (double2*)c->d_z, c->d_part, c->d_counter, c->d_sc);
	c->launches++;
}
void bp_launch_pcg_p(bp_ctx* c, int first, int par)
{
	const int n = (int)c->nn_owned;
	k_pcg_p<<<red_blocks(n, 512), 256, 0, c->stream>>>(n, first, par, c->prm.krylov_rtol, c->prm.krylov_atol, c->prm.krylov_norm_preconditioned, (const double2*)c->d_z, (double2*)c->d_p, c->d_sc);
	c->launches++;
}

// ---------------------------------------------------------------- general preconditioners
// CG steps with the preconditioner applied by separate kernels (Chebyshev, column solve)
__global__ void __launch_bounds__(256)
k_pcg_init2(int nrows, const double2* __restrict__ b, double2* x, double2* r)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) { x[i] = make_double2(0.0, 0.0); r[i] = b[i]; }
}
__global__ void __launch_bounds__(256)
k_pcg_update2(int nrows, int par, const double2* __restrict__ p, const double2* __restrict__ Ap, double2* x, double2* r, const double* sc)
{
	if (sc[SC_DONE] != 0.0) return;
	const double pAp = sc[SC_PAP];
	if (!(pAp > 0.0)) return;
	const double alpha = sc[par ? SC_T1 : SC_T0] / pAp;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) {
		const double2 pv = p[i], av = Ap[i];
		double2 xv = x[i], rv = r[i];
		xv.x += alpha * pv.x; xv.y += alpha * pv.y; rv.x -= alpha * av.x; rv.y -= alpha * av.y;
		x[i] = xv; r[i] = rv;
	}
}
__global__ void __launch_bounds__(256)
k_dots3(int nrows, const double2* __restrict__ r, const double2* __restrict__ z, double* part, unsigned* counter, double* out, const double* sc, int guard)
{
	if (guard && (sc[SC_DONE] != 0.0 || !(sc[SC_PAP] > 0.0))) return;
	double acc[3] = { 0.0, 0.0, 0.0 };
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) {
		const double2 rv = r[i], zv = z[i];
		acc[0] += rv.x * zv.x + rv.y * zv.y; acc[1] += zv.x * zv.x + zv.y * zv.y; acc[2] += rv.x * rv.x + rv.y * rv.y;
	}
	block_reduce_finalize<3>(acc, part, RED_BLOCKS_MAX, counter, out);
}
void bp_launch_pcg_init2(bp_ctx* c)
{
	const int n = (int)c->nn_owned;
	cudaMemsetAsync(c->d_sc + SC_DONE, 0, sizeof(double), c->stream);
	cudaMemsetAsync(c->d_sc + SC_PAP, 0, sizeof(double), c->stream);
	k_pcg_init2<<<red_blocks(n, 512), 256, 0, c->stream>>>(n, (const double2*)c->d_F, (double2*)c->d_dx, (double2*)c->d_r);
	c->launches++;
}
void bp_launch_pcg_update2(bp_ctx* c, int par)
{
	const int n = (int)c->nn_owned;
	k_pcg_update2<<<red_blocks(n, 512), 256, 0, c->stream>>>(n, par, (const double2*)c->d_p, (const double2*)c->d_Ap, (double2*)c->d_dx, (double2*)c->d_r, c->d_sc);
	c->launches++;
}
void bp_launch_dots3(bp_ctx* c, int slot, int guard)
{
	const int n = (int)c->nn_owned;
	k_dots3<<<red_blocks(n, 512), 256, 0, c->stream>>>(n, (const double2*)c->d_r, (const double2*)c->d_z, c->d_part, c->d_counter, c->d_sc + slot, c->d_sc, guard);
	c->launches++;
}

// vertical-line solve z = T^-1 r: T = the block-tridiagonal coupling of the nodes of one
// ice column (2x2 blocks), one thread per column, block Thomas algorithm.
#define BP_MAX_LAYERS 64
__device__ __forceinline__ void inv2(const double a, const double b, const double c, const double d, double& ia, double& ib, double& ic, double& id)
{
	const double idet = 1.0 / (a * d - b * c);
	ia = d * idet; ib = -b * idet; ic = -c * idet; id = a * idet;
}
__global__ void __launch_bounds__(128)
k_col_solve(int ncol, const int32_t* __restrict__ colptr, const int32_t* __restrict__ colnode,
            const int32_t* __restrict__ tri_d, const int32_t* __restrict__ tri_u, const int32_t* __restrict__ tri_l,
            const double* __restrict__ val, const double2* __restrict__ r, double2* __restrict__ z)
{
	const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
	if (cidx >= ncol) return;
	const int j0 = colptr[cidx], n = colptr[cidx + 1] - j0;
	double Di[BP_MAX_LAYERS][4];        // inverse of the eliminated diagonal blocks
	double2 rr[BP_MAX_LAYERS];
	for (int k = 0; k < n; ++k) {
		const int j = j0 + k;
		const double* D = val + 4 * (size_t)tri_d[j];
		double a = D[0], b = D[1], cc = D[2], d = D[3];
		double2 rv = r[colnode[j]];
		if (k > 0 && tri_l[j] >= 0 && tri_u[j - 1] >= 0) {
			const double* L = val + 4 * (size_t)tri_l[j];
			const double* U = val + 4 * (size_t)tri_u[j - 1];
			// M = L * Di[k-1]
			const double m0 = L[0] * Di[k - 1][0] + L[1] * Di[k - 1][2], m1 = L[0] * Di[k - 1][1] + L[1] * Di[k - 1][3];
			const double m2 = L[2] * Di[k - 1][0] + L[3] * Di[k - 1][2], m3 = L[2] * Di[k - 1][1] + L[3] * Di[k - 1][3];
			a -= m0 * U[0] + m1 * U[2]; b -= m0 * U[1] + m1 * U[3];
			cc -= m2 * U[0] + m3 * U[2]; d -= m2 * U[1] + m3 * U[3];
			rv.x -= m0 * rr[k - 1].x + m1 * rr[k - 1].y; rv.y -= m2 * rr[k - 1].x + m3 * rr[k - 1].y;
		}
		inv2(a, b, cc, d, Di[k][0], Di[k][1], Di[k][2], Di[k][3]);
		rr[k] = rv;
	}
	double2 zn = make_double2(0.0, 0.0);
	for (int k = n - 1; k >= 0; --k) {
		const int j = j0 + k;
		double2 rv = rr[k];
		if (k < n - 1 && tri_u[j] >= 0) {
			const double* U = val + 4 * (size_t)tri_u[j];
			rv.x -= U[0] * zn.x + U[1] * zn.y; rv.y -= U[2] * zn.x + U[3] * zn.y;
		}
		zn = make_double2(Di[k][0] * rv.x + Di[k][1] * rv.y, Di[k][2] * rv.x + Di[k][3] * rv.y);
		z[colnode[j]] = zn;
	}
}
void bp_launch_col_solve(bp_ctx* c, const double* r, double* z)
{
	k_col_solve<<<(int)((c->n_cols + 127) / 128), 128, 0, c->stream>>>((int)c->n_cols, c->d_colptr, c->d_colnode, c->d_tri_d, c->d_tri_u, c->d_tri_l,
		c->d_val, (const double2*)r, (double2*)z);
	c->launches++;
}

// Chebyshev step with the column (line) splitting: per ice column solve T delta = res with the
// block Thomas algorithm, then d = c1 d + c2 delta and znew = zold + d (multigrid level-0 smoother).
// MAXL bounds the column height at compile time so that the eliminated blocks stay in registers.
template <int MAXL>
__global__ void __launch_bounds__(128)
k_col_cheb(int ncol, double c1, double c2, const int32_t* __restrict__ colptr, const int32_t* __restrict__ colnode,
           const int32_t* __restrict__ tri_d, const int32_t* __restrict__ tri_u, const int32_t* __restrict__ tri_l,
           const double2* __restrict__ val, const double2* __restrict__ res, const double2* __restrict__ zold,
           double2* __restrict__ znew, double2* __restrict__ d)
{
	const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
	if (cidx >= ncol) return;
	const int j0 = colptr[cidx], n = colptr[cidx + 1] - j0;
	double Di[MAXL][4];
	double2 rr[MAXL];
	int nd[MAXL];
	#pragma unroll
	for (int k = 0; k < MAXL; ++k) {
		if (k < n) {
			const int j = j0 + k;
			nd[k] = colnode[j];
			const double2 D01 = val[2 * (size_t)tri_d[j]], D23 = val[2 * (size_t)tri_d[j] + 1];
			double a = D01.x, b = D01.y, cc = D23.x, dd = D23.y;
			double2 rv = res[nd[k]];
			const int sl = tri_l[j], su = (k > 0) ? tri_u[j - 1] : -1;
			if (k > 0 && sl >= 0 && su >= 0) {
				const double2 L01 = val[2 * (size_t)sl], L23 = val[2 * (size_t)sl + 1];
				const double2 U01 = val[2 * (size_t)su], U23 = val[2 * (size_t)su + 1];
				const double m0 = L01.x * Di[k - 1][0] + L01.y * Di[k - 1][2], m1 = L01.x * Di[k - 1][1] + L01.y * Di[k - 1][3];
				const double m2 = L23.x * Di[k - 1][0] + L23.y * Di[k - 1][2], m3 = L23.x * Di[k - 1][1] + L23.y * Di[k - 1][3];
				a -= m0 * U01.x + m1 * U23.x; b -= m0 * U01.y + m1 * U23.y;
				cc -= m2 * U01.x + m3 * U23.x; dd -= m2 * U01.y + m3 * U23.y;
				rv.x -= m0 * rr[k - 1].x + m1 * rr[k - 1].y; rv.y -= m2 * rr[k - 1].x + m3 * rr[k - 1].y;
			}
			inv2(a, b, cc, dd, Di[k][0], Di[k][1], Di[k][2], Di[k][3]);
			rr[k] = rv;
		}
	}
	double2 zn = make_double2(0.0, 0.0);
	#pragma unroll
	for (int k = MAXL - 1; k >= 0; --k) {
		if (k < n) {
			const int j = j0 + k;
			double2 rv = rr[k];
			const int su = tri_u[j];
			if (k < n - 1 && su >= 0) {
				const double2 U01 = val[2 * (size_t)su], U23 = val[2 * (size_t)su + 1];
				rv.x -= U01.x * zn.x + U01.y * zn.y; rv.y -= U23.x * zn.x + U23.y * zn.y;
			}
			zn = make_double2(Di[k][0] * rv.x + Di[k][1] * rv.y, Di[k][2] * rv.x + Di[k][3] * rv.y);
			double2 dv = make_double2(c2 * zn.x, c2 * zn.y);
			if (c1 != 0.0) { const double2 dold = d[nd[k]]; dv.x += c1 * dold.x; dv.y += c1 * dold.y; }
			d[nd[k]] = dv;
			if (zold) { const double2 zo = zold[nd[k]]; znew[nd[k]] = make_double2(zo.x + dv.x, zo.y + dv.y); }
			else znew[nd[k]] = dv;
		}
	}
}
void bp_launch_col_cheb(bp_ctx* c, double c1, double c2, const double* res, const double* zold, double* znew)
{
	const int g = (int)((c->n_cols + 127) / 128);
	#define COLCHEB(M) k_col_cheb<M><<<g, 128, 0, c->stream>>>((int)c->n_cols, c1, c2, c->d_colptr, c->d_colnode, c->d_tri_d, c->d_tri_u, c->d_tri_l, \
		(const double2*)c->d_val, (const double2*)res, (const double2*)zold, (double2*)znew, (double2*)c->d_cd)
	if (c->max_layers <= 8) COLCHEB(8);
	else if (c->max_layers <= 16) COLCHEB(16);
	else if (c->max_layers <= 32) COLCHEB(32);
	else COLCHEB(64);
	c->launches++;
}

// scalar tridiagonal solve with partial pivoting per ice column (LAPACK dgtsv recurrences) on
// component 0 of the blocks: the continuity matrix has zero diagonals in the interior, so the
// unpivoted Thomas algorithm of k_col_solve cannot be used.  z = T^-1 r, second component 0.
__global__ void __launch_bounds__(128)
k_col_gtsv(int ncol, const int32_t* __restrict__ colptr, const int32_t* __restrict__ colnode,
           const int32_t* __restrict__ tri_d, const int32_t* __restrict__ tri_u, const int32_t* __restrict__ tri_l,
           const double* __restrict__ val, const double2* __restrict__ r, double2* __restrict__ z)
{
	const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
	if (cidx >= ncol) return;
	const int j0 = colptr[cidx], n = colptr[cidx + 1] - j0;
	double d[BP_MAX_LAYERS], du[BP_MAX_LAYERS], dl[BP_MAX_LAYERS], b[BP_MAX_LAYERS];
	for (int k = 0; k < n; ++k) {
		const int j = j0 + k;
		d[k] = val[4 * (size_t)tri_d[j]];
		du[k] = (k < n - 1 && tri_u[j] >= 0) ? val[4 * (size_t)tri_u[j]] : 0.0;
		dl[k] = (k < n - 1 && tri_l[j + 1] >= 0) ? val[4 * (size_t)tri_l[j + 1]] : 0.0;   // entry (k+1, k)
		b[k] = r[colnode[j]].x;
	}
	for (int k = 0; k < n - 1; ++k) {
		if (dl[k] == 0.0) continue;
		if (fabs(d[k]) >= fabs(dl[k])) {
			const double mult = dl[k] / d[k];
			d[k + 1] -= mult * du[k]; b[k + 1] -= mult * b[k];
			if (k < n - 2) dl[k] = 0.0;
		} else {
			const double mult = d[k] / dl[k];
			d[k] = dl[k];
			double temp = d[k + 1];
			d[k + 1] = du[k] - mult * temp;
			if (k < n - 2) { dl[k] = du[k + 1]; du[k + 1] = -mult * dl[k]; }
			du[k] = temp;
			temp = b[k]; b[k] = b[k + 1]; b[k + 1] = temp - mult * b[k + 1];
		}
	}
	b[n - 1] /= d[n - 1];
	if (n > 1) b[n - 2] = (b[n - 2] - du[n - 2] * b[n - 1]) / d[n - 2];
	for (int k = n - 3; k >= 0; --k) b[k] = (b[k] - du[k] * b[k + 1] - dl[k] * b[k + 2]) / d[k];
	for (int k = 0; k < n; ++k) z[colnode[j0 + k]] = make_double2(b[k], 0.0);
}
void bp_launch_col_gtsv(bp_ctx* c, const double* r, double* z)
{
	k_col_gtsv<<<(int)((c->n_cols + 127) / 128), 128, 0, c->stream>>>((int)c->n_cols, c->d_colptr, c->d_colnode, c->d_tri_d, c->d_tri_u, c->d_tri_l,
		c->d_val, (const double2*)r, (double2*)z);
	c->launches++;
}
__global__ void k_lincomb(int n, double* out, double a, const double* x, double b, const double* y, double cc, const double* z)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		double v = a * x[i] + b * y[i];
		if (cc != 0.0) v += cc * z[i];
		out[i] = v;
	}
}
void bp_launch_lincomb(bp_ctx* c, double* out, double a, const double* x, double b, const double* y, double cc, const double* z)
{
	const int n = (int)(2 * c->nn_owned);
	k_lincomb<<<red_blocks(n, 1024), 256, 0, c->stream>>>(n, out, a, x, b, y, cc, z);
	c->launches++;
}

// Chebyshev polynomial in D^-1 J: first step d = c0 D^-1 r, z = d; further steps fused with
// the SpMV: res = r - J z_old; d = c1 d + c2 D^-1 res; z_new = z_old + d.
__global__ void __launch_bounds__(256)
k_cheb_first(int nrows, double c0, const double* __restrict__ dinv, const double2* __restrict__ r, double2* z, double2* d, const double* __restrict__ lmax_dev)
{
	if (lmax_dev) c0 /= *lmax_dev;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) {
		const double2 v = apply_dinv(dinv, i, r[i]);
		const double2 dv = make_double2(c0 * v.x, c0 * v.y);
		d[i] = dv; z[i] = dv;
	}
}
// 2x2 block {a00 a01 a10 a11} of the resident matrix: FP64 (double2 pairs) or the FP32 copy the
// multigrid V-cycle reads (float4; accumulation stays FP64, vectors stay FP64)
struct BlkD { typedef double2 T; static __device__ __forceinline__ void ld(const double2* v, size_t slot, double& a0, double& a1, double& a2, double& a3)
	{ const double2 p = v[2 * slot], q = v[2 * slot + 1]; a0 = p.x; a1 = p.y; a2 = q.x; a3 = q.y; } };
struct BlkF { typedef float4 T; static __device__ __forceinline__ void ld(const float4* v, size_t slot, double& a0, double& a1, double& a2, double& a3)
	{ const float4 p = v[slot]; a0 = p.x; a1 = p.y; a2 = p.z; a3 = p.w; } };

template <class B>
__global__ void __launch_bounds__(256)
k_cheb_step(int nrows, double c1, double c2, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ col,
            const typename B::T* __restrict__ val, const double* __restrict__ dinv, const double2* __restrict__ r,
            const double2* __restrict__ zold, double2* __restrict__ znew, double2* __restrict__ d, const double* __restrict__ lmax_dev)
{
	if (lmax_dev) c2 /= *lmax_dev;
	const int lane = threadIdx.x & 31;
	const int nsl = (nrows + 31) >> 5;
	for (int slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slice < nsl; slice += (gridDim.x * blockDim.x) >> 5) {
		const int b0 = sell_ptr[slice], w = (sell_ptr[slice + 1] - b0) >> 5;
		const int row = (slice << 5) + lane;
		const int32_t* cp = col + b0 + lane;
		const size_t s0_ = (size_t)b0 + lane;
		double s0 = 0.0, s1 = 0.0;
		#pragma unroll 4
		for (int k = 0; k < w; ++k) {
			double a0, a1, a2, a3;
			B::ld(val, s0_ + 32 * (size_t)k, a0, a1, a2, a3);
			const double2 xv = zold[cp[32 * k]];
			s0 += a0 * xv.x + a1 * xv.y; s1 += a2 * xv.x + a3 * xv.y;
		}
		if (row < nrows) {
			const double2 rv = r[row], zo = zold[row], dv = d[row];
			const double2 pr = apply_dinv(dinv, row, make_double2(rv.x - s0, rv.y - s1));
			const double2 dn = make_double2(c1 * dv.x + c2 * pr.x, c1 * dv.y + c2 * pr.y);
			d[row] = dn;
			znew[row] = make_double2(zo.x + dn.x, zo.y + dn.y);
		}
	}
}
// b - A z with either copy of the matrix (multigrid residual before restriction)
template <class B>
__global__ void __launch_bounds__(256)
k_residual(int nrows, const int32_t* __restrict__ sell_ptr, const int32_t* __restrict__ col, const typename B::T* __restrict__ val,
           const double2* __restrict__ b, const double2* __restrict__ z, double2* __restrict__ out)
{
	const int lane = threadIdx.x & 31;
	const int nsl = (nrows + 31) >> 5;
	for (int slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slice < nsl; slice += (gridDim.x * blockDim.x) >> 5) {
		const int b0 = sell_ptr[slice], w = (sell_ptr[slice + 1] - b0) >> 5;
		const int row = (slice << 5) + lane;
		const int32_t* cp = col + b0 + lane;
		const size_t s0_ = (size_t)b0 + lane;
		double s0 = 0.0, s1 = 0.0;
		#pragma unroll 4
		for (int k = 0; k < w; ++k) {
			double a0, a1, a2, a3;
			B::ld(val, s0_ + 32 * (size_t)k, a0, a1, a2, a3);
			const double2 xv = z[cp[32 * k]];
			s0 += a0 * xv.x + a1 * xv.y; s1 += a2 * xv.x + a3 * xv.y;
		}
		if (row < nrows) { const double2 bv = b[row]; out[row] = make_double2(bv.x - s0, bv.y - s1); }
	}
}
__global__ void k_val_to_float(size_t n4, const double2* __restrict__ v, float4* __restrict__ f)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
		const double2 p = v[2 * i], q = v[2 * i + 1];
		f[i] = make_float4((float)p.x, (float)p.y, (float)q.x, (float)q.y);
	}
}
// refresh the FP32 copy of the resident matrix (allocated on first use)
int bp_launch_val_to_float(bp_ctx* c)
{
	if (!c->d_valf) BP_CUDA(c, cudaMalloc((void**)&c->d_valf, (size_t)std::max<int64_t>(c->n_slots, 1) * sizeof(float4)));
	k_val_to_float<<<red_blocks(c->n_slots, 256), 256, 0, c->stream>>>((size_t)c->n_slots, (const double2*)c->d_val, (float4*)c->d_valf);
	c->launches++;
	return BP_OK;
}
void bp_launch_residual(bp_ctx* c, const double* b, const double* z, double* out, bool f32)
{
	const int g = red_blocks(c->nn_owned, 256);
	if (f32 && c->d_valf) k_residual<BlkF><<<g, 256, 0, c->stream>>>((int)c->nn_owned, c->d_sell_ptr, c->d_col, (const float4*)c->d_valf, (const double2*)b, (const double2*)z, (double2*)out);
	else k_residual<BlkD><<<g, 256, 0, c->stream>>>((int)c->nn_owned, c->d_sell_ptr, c->d_col, (const double2*)c->d_val, (const double2*)b, (const double2*)z, (double2*)out);
	c->launches++;
}
void bp_launch_cheb_first(bp_ctx* c, double c0, const double* r, double* z, const double* lmax_dev)
{
	const int n = (int)c->nn_owned;
	k_cheb_first<<<red_blocks(n, 512), 256, 0, c->stream>>>(n, c0, c->d_dinv, (const double2*)r, (double2*)z, (double2*)c->d_cd, lmax_dev);
	c->launches++;
}
void bp_launch_cheb_step(bp_ctx* c, double c1, double c2, const double* r, const double* zold, double* znew, const double* lmax_dev, bool f32)
{
	if (!c->grid_cheb) c->grid_cheb = one_wave(k_cheb_step<BlkD>, 256, c->nn_owned, 256);
	if (f32 && c->d_valf)
		k_cheb_step<BlkF><<<c->grid_cheb, 256, 0, c->stream>>>((int)c->nn_owned, c1, c2, c->d_sell_ptr, c->d_col, (const float4*)c->d_valf, c->d_dinv,
			(const double2*)r, (const double2*)zold, (double2*)znew, (double2*)c->d_cd, lmax_dev);
	else
		k_cheb_step<BlkD><<<c->grid_cheb, 256, 0, c->stream>>>((int)c->nn_owned, c1, c2, c->d_sell_ptr, c->d_col, (const double2*)c->d_val, c->d_dinv,
			(const double2*)r, (const double2*)zold, (double2*)znew, (double2*)c->d_cd, lmax_dev);
	c->launches++;
}

// Gershgorin bound of lmax(D^-1 J): max over block rows of sum_j |D_i^-1 J_ij|_inf.
// A guaranteed upper bound (power iteration converges from below and a Chebyshev interval
// that misses the top of the spectrum makes the preconditioner indefinite).
__global__ void __launch_bounds__(256)
k_gershgorin(int nrows, const int32_t* __restrict__ sell_ptr, const double2* __restrict__ val,
             const double* __restrict__ dinv, double* part, unsigned* counter, double* out, double scale)
{
	__shared__ double sm[32];
	__shared__ bool last;
	const int lane = threadIdx.x & 31, w_ = threadIdx.x >> 5;
	const int nsl = (nrows + 31) >> 5;
	double mx = 0.0;
	for (int slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; slice < nsl; slice += (gridDim.x * blockDim.x) >> 5) {
		const int b0 = sell_ptr[slice], w = (sell_ptr[slice + 1] - b0) >> 5;
		const int row = (slice << 5) + lane;
		if (row >= nrows) continue;
		const double2 d01 = ((const double2*)dinv)[2 * (size_t)row], d23 = ((const double2*)dinv)[2 * (size_t)row + 1];
		const double2* vp = val + 2 * ((size_t)b0 + lane);
		double g0 = 0.0, g1 = 0.0;
		for (int k = 0; k < w; ++k) {
			const double2 a01 = vp[64 * (size_t)k], a23 = vp[64 * (size_t)k + 1];
			g0 += fabs(d01.x * a01.x + d01.y * a23.x) + fabs(d01.x * a01.y + d01.y * a23.y);
			g1 += fabs(d23.x * a01.x + d23.y * a23.x) + fabs(d23.x * a01.y + d23.y * a23.y);
		}
		mx = fmax(mx, fmax(g0, g1));
	}
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
	if (lane == 0) sm[w_] = mx;
	__syncthreads();
	if (w_ == 0) {
		mx = (lane < (int)(blockDim.x >> 5)) ? sm[lane] : 0.0;
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
		if (lane == 0) part[blockIdx.x] = mx;
	}
	if (threadIdx.x == 0) { __threadfence(); last = (atomicAdd(counter, 1u) == gridDim.x - 1); }
	__syncthreads();
	if (last && w_ == 0) {
		__threadfence();
		mx = 0.0;
		for (int b = lane; b < (int)gridDim.x; b += 32) mx = fmax(mx, ((volatile double*)part)[b]);
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
		if (lane == 0) { *out = mx * scale; *counter = 0u; }
	}
}
void bp_launch_gershgorin(bp_ctx* c, int sc_slot)
{
	k_gershgorin<<<red_blocks(c->nn_owned, 256), 256, 0, c->stream>>>((int)c->nn_owned, c->d_sell_ptr, (const double2*)c->d_val, c->d_dinv,
		c->d_part, c->d_counter, c->d_sc + sc_slot, sc_slot == SC_LMAX ? c->env.mg_lmax_scale : 1.0);
	c->launches++;
}

// helpers for the power iteration that estimates lmax(D^-1 J)
__global__ void k_apply_dinv(int nrows, const double* __restrict__ dinv, double2* v)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) v[i] = apply_dinv(dinv, i, v[i]);
}
__global__ void k_scale_inv_norm(int n, double* v, const double* sc, int slot)
{
	const double s = rsqrt(sc[slot]);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] *= s;
}
__global__ void k_fill_probe(int n, double* v)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = 1.0 + 0.5 * sin(0.7 * (double)i) + 0.25 * cos(0.013 * (double)i);
}
void bp_launch_apply_dinv(bp_ctx* c, double* v)
{
	k_apply_dinv<<<red_blocks(c->nn_owned, 512), 256, 0, c->stream>>>((int)c->nn_owned, c->d_dinv, (double2*)v);
	c->launches++;
}
void bp_launch_scale_inv_norm(bp_ctx* c, double* v, int sc_slot)
{
	k_scale_inv_norm<<<red_blocks(2 * c->nn_owned, 1024), 256, 0, c->stream>>>((int)(2 * c->nn_owned), v, c->d_sc, sc_slot);
	c->launches++;
}
void bp_launch_fill_probe(bp_ctx* c, double* v)
{
	k_fill_probe<<<red_blocks(2 * c->nn_owned, 1024), 256, 0, c->stream>>>((int)(2 * c->nn_owned), v);
	c->launches++;
}

// sum d_sc[slot..slot+count) across the contexts of an in-process group (same device)
struct ScPtrs { double* p[16]; };
__global__ void k_sum_scalars(ScPtrs P, int n, int slot, int count)
{
	const int i = threadIdx.x;
	if (i >= count) return;
	double s = 0.0;
	for (int r = 0; r < n; ++r) s += P.p[r][slot + i];
	for (int r = 0; r < n; ++r) P.p[r][slot + i] = s;
}
void bp_launch_sum_scalars(bp_ctx** ctxs, int n, int slot, int count)
{
	ScPtrs P;
	for (int r = 0; r < n; ++r) P.p[r] = ctxs[r]->d_sc;
	k_sum_scalars<<<1, 32, 0, ctxs[0]->stream>>>(P, n, slot, count);
	ctxs[0]->launches++;
}
