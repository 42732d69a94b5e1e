// bp_tiles.cu -- compute-once cell assembly for extruded prism meshes (the D3Model mesh family).
//
// Replaces FFC tabulate_tensor + dolfin Assembler for the cell integrals of the BP residual and
// Jacobian (cslvr/momentumbp.py:481-500, cslvr/physics.py:75-77) on meshes that D3Model produces:
// a footprint triangulation extruded into layers, every prism split into three tetrahedra the
// same way along a column (BoxMesh, and the gmsh-extruded meshes of cslvr/meshing.py:771-860).
//
// k_rows (bp_kernels.cu) rebuilds every tetrahedron's cell-constant data once per node: 4x the
// FP64 work.  Here every tetrahedron is evaluated ONCE:
//   * a CTA owns a tile of ice columns (compact footprint patch) and all their layers;
//   * one THREAD per prism column (footprint triangle), marching bottom -> top.  With the column's
//     canonical vertex order (c0, c1, c2) the three tets of a prism are always
//         (b0 b1 b2 t2)  (b0 b1 t1 t2)  (b0 t0 t1 t2),        b = layer k, t = layer k + 1,
//     so the code is straight-line and the warp never diverges;
//   * contributions of consecutive prisms to the nodes of their shared interface are merged in
//     registers; per layer the thread stages 51 doubles in shared memory ([slot][thread]: bank =
//     thread, conflict-free): the interface blocks 3 x D (symmetric 2x2), 3 x E, 3 x F and the six
//     blocks between the two layers;
//   * 8 summing warps of the same CTA (warp-specialised, registers re-balanced with setmaxnreg) turn the staged values
//     into the block rows of the tile's own columns: a table (layer-independent) lists for every Jacobian block the
//     (thread, staged slot) pairs to add -- fixed order, no atomics, bitwise reproducible -- and the SELL-32 slot it goes
//     to (one 256-bit store per block; J is symmetric: the transposed block of the partner row from the same sum).
//     Rows are complete inside the tile: prism columns that touch a tile column but belong to a neighbouring tile are
//     recomputed (halo factor ~1.2).  The same warps feed the compute warps (cp.async vertex packages);
//   * the basal facet of a prism column is its bottom triangle: sliding and water-pressure integrals start the bottom
//     interface of the march, so the facet kernel only runs for lateral facets.
// Staging is double-buffered (224 prism columns per tile), hand-over by named barriers: summing layer k overlaps
// computing layer k + 1.  Every J block and F entry is written exactly once (checked at set-up).
//
// The index tables are also run by a host emulator (bp_debug_tiles_emulate): the CPU test suite
// checks tables + element math against the oracle without a GPU.  It is not reachable from
// bp_assemble.
#include "bp_internal.h"
#include <algorithm>
#include <parallel/algorithm>     // __gnu_parallel::sort for the million-element sorts of the set-up (OpenMP)
#include <cmath>
#include <cstring>
#include <numeric>

// Shape (compile time).  Default: the staging buffer is DOUBLE-buffered, so the sums of layer k run while the compute warps
// are already staging layer k + 1 (with one buffer the two kinds of warps take turns: measured 49 % of the compute warps'
// time and 46 % of the summing warps' time spent at the hand-over barriers).  Two buffers of 51 doubles per thread only fit
// into the 227 KB of an SM for 224 prism columns per tile: the 8 compute warps use 28 of their 32 lanes (the other four
// lanes shadow lane 0 and store nothing), which keeps two compute warps on every scheduler and the register split of
// setmaxnreg (a warpgroup = 4 warps must agree on it).  -DTILE_NBUF=1 -DTILE_LANES=32 is the single-buffered shape.
#ifndef TILE_NBUF
#define TILE_NBUF   2
#endif
#ifndef TILE_LANES
#define TILE_LANES  (TILE_NBUF == 2 ? 28 : 32)   // active lanes per compute warp
#endif
#define TILE_CW     8                            // compute warps
#define TILE_T      (TILE_CW * TILE_LANES)       // prism columns per tile (incl. halo) = compute threads that stage
#define TILE_TW     (TILE_CW * 32)               // threads of the compute warps
#define TILE_S      256                          // 8 summing warps: two per scheduler
#define TILE_CTAS   1
#define TILE_REG_C  "216"      // launch bound 512 threads -> 128 registers each; 256 x 88 move from the summing to the compute warps
#define TILE_REG_S  "40"
#define TILE_EPT    5          // table entries per summing thread (at most)
#define TILE_NSLOT  51         // staged doubles per compute thread and layer
#define TILE_MAXC   120        // ice columns owned by one tile
#define TILE_MAXE   (TILE_NBUF == 2 ? 1152 : TILE_EPT * TILE_S)   // table entries per tile
#define TILE_TABW   (8 * TILE_MAXE)       // contribution words (uint16) of one tile's table in shared memory

// staged slots:   0..8  D(b0) D(b1) D(b2): xx xy yy      9..20 E(b0b1) E(b0b2) E(b1b2): xx xy yx yy (row first, col second)
//                21..26 F(b0) F(b1) F(b2): x y           27..50 C00 C01 C02 C11 C12 C22: block (row b_i, col t_j)
// a contribution word: offset slot * TILE_T + thread into the staged buffer (bits 0-13) | type (bits 14-15); 0xffff = none
enum { CT_D = 0, CT_E = 1, CT_ET = 2, CT_F = 3 };
#define CW(t, slot, ty) ((uint16_t)(((slot) * TILE_T + (t)) | ((ty) << 14)))
#define CW_NONE 0xffffu

// old-style contribution codes used while the tables are built: 0-2 D(b_i), 3-5 E direct, 6-8 E transposed,
// 9-14 C direct, 15-20 C transposed, 21-23 F(b_i) -> (first slot, type)
static const uint8_t H_CODE_BASE[24] = { 0, 3, 6,  9, 13, 17,  9, 13, 17,  27, 31, 35, 39, 43, 47,  27, 31, 35, 39, 43, 47,  21, 23, 25 };
static const uint8_t H_CODE_TYPE[24] = { CT_D, CT_D, CT_D,  CT_E, CT_E, CT_E,  CT_ET, CT_ET, CT_ET,  CT_E, CT_E, CT_E, CT_E, CT_E, CT_E,
                                         CT_ET, CT_ET, CT_ET, CT_ET, CT_ET, CT_ET,  CT_F, CT_F, CT_F };

static inline int pair_E(int i, int j) { return (i == 0) ? (j - 1) : 2; }                 // (0,1) (0,2) (1,2) -> 0 1 2
static inline int pair_C(int i, int j) { return (i == 0) ? j : (i == 1 ? 2 + j : 5); }   // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2) -> 0..5

struct bp_tiles {
	int     n_tiles = 0, NL = 0, W = 8;       // W: contribution words per table entry (8 or 16)
	int64_t n_ent = 0, n_contrib = 0, n_prism_cols = 0, n_threads_used = 0;
	int32_t *d_vert = nullptr, *d_node = nullptr, *d_tet = nullptr, *d_nthr = nullptr, *d_ent0 = nullptr, *d_out = nullptr, *d_out2 = nullptr;
	uint16_t *d_tab = nullptr;                // [n_ent][W]
	uint16_t *d_tflag = nullptr;              // [n_tiles][TILE_T] which staged blocks a compute thread writes transposed
	// basal facets folded into the march: the bottom triangle of a prism column IS its basal facet, so the sliding and
	// water-pressure integrals of dGamma_b / dGamma_bw start the bottom-interface accumulators instead of being added by a
	// second kernel through read-modify-write of the finished rows.  Bit 0: sliding (markers 3, 5), bit 1: water pressure (5).
	uint8_t  *d_tfacet = nullptr;             // [n_tiles][TILE_T]
	std::vector<uint8_t> h_tfacet;
	// the pipelined host-buffer call as one graph (bp_tiles_assemble_pipelined)
	cudaGraphExec_t pipe_graph = nullptr;
	const void *pg_u = nullptr, *pg_F = nullptr; int pg_what = 0; cudaStream_t pg_stream = nullptr; int64_t pg_launches = 0; bool pipe_no_graph = false;
	bool      folded = false;                 // ... and the facet kernel only sees the facets that are left (the lateral ones)
	int32_t  *d_fn_node = nullptr, *d_fn_ptr = nullptr, *d_fn_inc = nullptr;   // their row-gather lists (as bp_ctx::d_fn_*)
	int64_t   n_fnodes = 0;
	std::vector<uint8_t> h_facet_node;        // owned nodes they add to (pipelined host-buffer path)
	uint8_t  *d_einfo = nullptr;              // [n_ent] contributions (bits 0-4) | mixed types (bit 5) | kind (bits 6-7: 0 block, 1 diagonal, 2 F)
	double  *d_ainv = nullptr;
	bool     attr_set = false;
	// host copies (kept by host-only debug contexts for the emulator)
	std::vector<int32_t> h_vert, h_node, h_tet, h_nthr, h_ent0, h_out, h_out2;   // out2: slot of the transposed block (same sum, second output)
	std::vector<uint16_t> h_tab, h_tflag;
	std::vector<uint8_t> h_einfo;
	// host-buffer bp_assemble, pipelined: groups of consecutive tiles with the node ranges of u they need (beyond what the
	// groups before them brought in) and the ranges of F rows they complete (those a facet integral adds to leave later)
	struct Chunk { int tile0 = 0, ntile = 0; std::vector<std::pair<int32_t, int32_t>> in, out_plain, out_facet; };
	std::vector<Chunk> chunks;
};

// ------------------------------------------------------------------------------------------------
// element math: one tetrahedron, evaluated once.  Closed form of SURVEY.md §8a (the strain rate of a
// P1 velocity is cell-constant; the quadrature of A^(-1/n) arrives as ai = sum_q w_q A_q^(-1/n)).
// F[2a + c]; D[3a + {xx, xy, yy}] (diagonal blocks are symmetric); E[4 pair + {xx, xy, yx, yy}] for the
// pairs (0,1) (0,2) (0,3) (1,2) (1,3) (2,3), block (row a, col b).
struct TetOut { double F[8]; double D[12]; double E[24]; };

__host__ __device__ __forceinline__ void tet_eval(const double4 p0, const double4 p1, const double4 p2, const double4 p3,
                                                   const double2 u0, const double2 u1, const double2 u2, const double2 u3,
                                                   const double ai, const AsmParams& P, TetOut& o)
{
	const double e1x = p1.x - p0.x, e1y = p1.y - p0.y, e1z = p1.z - p0.z;
	const double e2x = p2.x - p0.x, e2y = p2.y - p0.y, e2z = p2.z - p0.z;
	const double e3x = p3.x - p0.x, e3y = p3.y - p0.y, e3z = p3.z - p0.z;
	double X[4], Y[4], Z[4];
	X[1] = e2y * e3z - e2z * e3y; Y[1] = e2z * e3x - e2x * e3z; Z[1] = e2x * e3y - e2y * e3x;
	X[2] = e3y * e1z - e3z * e1y; Y[2] = e3z * e1x - e3x * e1z; Z[2] = e3x * e1y - e3y * e1x;
	X[3] = e1y * e2z - e1z * e2y; Y[3] = e1z * e2x - e1x * e2z; Z[3] = e1x * e2y - e1y * e2x;
	const double det = e1x * X[1] + e1y * Y[1] + e1z * Z[1];
#ifdef __CUDA_ARCH__
	const double idet = __drcp_rn(det);
#else
	const double idet = 1.0 / det;
#endif
	#pragma unroll
	for (int a = 1; a < 4; ++a) { X[a] *= idet; Y[a] *= idet; Z[a] *= idet; }
	X[0] = -(X[1] + X[2] + X[3]); Y[0] = -(Y[1] + Y[2] + Y[3]); Z[0] = -(Z[1] + Z[2] + Z[3]);
	const double vol = fabs(det) * (1.0 / 6.0);
	const double wbar = vol * ai;
	const double ux[4] = { u0.x, u1.x, u2.x, u3.x }, uy[4] = { u0.y, u1.y, u2.y, u3.y }, S[4] = { p0.w, p1.w, p2.w, p3.w };
	double g00 = 0, g01 = 0, g02 = 0, g10 = 0, g11 = 0, g12 = 0, sx = 0, sy = 0;
	#pragma unroll
	for (int a = 0; a < 4; ++a) {
		g00 += ux[a] * X[a]; g01 += ux[a] * Y[a]; g02 += ux[a] * Z[a];
		g10 += uy[a] * X[a]; g11 += uy[a] * Y[a]; g12 += uy[a] * Z[a];
		sx += S[a] * X[a]; sy += S[a] * Y[a];
	}
	const double sh = 0.5 * (g01 + g10);
	const double ed = g00 * g00 + g11 * g11 + g00 * g11 + sh * sh + 0.25 * (g02 * g02 + g12 * g12) + P.eps_reg;
	// 2 eta = ed^((1-n)/(2n)),  2 eta' = (1-n)/(2n) ed^((1-n)/(2n) - 1)
	double two_eta, two_eta_p;
	if (P.n_is_3) {
#ifdef __CUDA_ARCH__
		two_eta = rcbrt(ed);
#else
		two_eta = 1.0 / cbrt(ed);
#endif
		two_eta_p = (-1.0 / 3.0) * (two_eta * two_eta) * (two_eta * two_eta);           // ed^(-4/3)
	} else {
		const double ex = (1.0 - P.n) / (2.0 * P.n);
		two_eta = pow(ed, ex);
		two_eta_p = ex * two_eta / ed;
	}
	const double alpha = wbar * two_eta, gamma = P.picard ? 0.0 : wbar * two_eta_p;
	const double cxx = 2.0 * g00 + g11, cyy = 2.0 * g11 + g00, hz0 = 0.5 * g02, hz1 = 0.5 * g12;
	double dx[4], dy[4];
	#pragma unroll
	for (int a = 0; a < 4; ++a) {
		dx[a] = cxx * X[a] + sh * Y[a] + hz0 * Z[a];
		dy[a] = cyy * Y[a] + sh * X[a] + hz1 * Z[a];
	}
	const double gx = P.rho_g * vol * 0.25 * sx, gy = P.rho_g * vol * 0.25 * sy;
	#pragma unroll
	for (int a = 0; a < 4; ++a) {
		// linear=True keeps the u-independent part only (gravity)
		o.F[2 * a]     = (P.picard ? 0.0 : alpha * dx[a]) + gx;
		o.F[2 * a + 1] = (P.picard ? 0.0 : alpha * dy[a]) + gy;
	}
	#pragma unroll
	for (int a = 0; a < 4; ++a) {
		const double Pa = 2.0 * alpha * X[a], Qa = 0.5 * alpha * Y[a], Ra = 0.5 * alpha * Z[a];
		const double Sa = 2.0 * alpha * Y[a], Ta = 0.5 * alpha * X[a];
		const double Ua = alpha * X[a], Wa = alpha * Y[a];
		const double Ea = gamma * dx[a], Ga = gamma * dy[a];
		o.D[3 * a]     = Pa * X[a] + Qa * Y[a] + Ra * Z[a] + Ea * dx[a];
		o.D[3 * a + 1] = Ua * Y[a] + Qa * X[a] + Ea * dy[a];
		o.D[3 * a + 2] = Sa * Y[a] + Ta * X[a] + Ra * Z[a] + Ga * dy[a];
		#pragma unroll
		for (int b = a + 1; b < 4; ++b) {
			const int pr = (a == 0) ? (b - 1) : (a == 1 ? b + 1 : 5);
			o.E[4 * pr]     = Pa * X[b] + Qa * Y[b] + Ra * Z[b] + Ea * dx[b];
			o.E[4 * pr + 1] = Ua * Y[b] + Qa * X[b] + Ea * dy[b];
			o.E[4 * pr + 2] = Wa * X[b] + Ta * Y[b] + Ga * dx[b];
			o.E[4 * pr + 3] = Sa * Y[b] + Ta * X[b] + Ra * Z[b] + Ga * dy[b];
		}
	}
}

// per-thread state of the march: the accumulators of the interface between two prisms
struct Iface { double D[9]; double E[12]; double F[6]; };

__host__ __device__ __forceinline__ void iface_zero(Iface& a)
{
	#pragma unroll
	for (int i = 0; i < 9; ++i) a.D[i] = 0.0;
	#pragma unroll
	for (int i = 0; i < 12; ++i) a.E[i] = 0.0;
	#pragma unroll
	for (int i = 0; i < 6; ++i) a.F[i] = 0.0;
}

// One prism = three tets (b0 b1 b2 t2) (b0 b1 t1 t2) (b0 t0 t1 t2).  bot holds what the prism below added to the nodes
// of the bottom interface; every staged value goes out as soon as it is final (after the tet that touches it last), so
// few accumulators stay live; top is SET to this prism's part of the top interface.  st.gate() is called once, before
// the first store (the kernel waits there until the summing warps have released the staging buffer).
// stage a 2x2 block (xx xy yx yy) in the orientation its (single) consumer wants: block b of this thread is read
// transposed iff bit b of the thread's flag word is set (E01 E02 E12 C00 C01 C02 C11 C12 C22 = bits 0..8), so the summing
// warps never have to look at the type of a contribution
template <class ST>
__host__ __device__ __forceinline__ void put_blk(ST& st, const int slot, const int blk, const double xx, const double xy, const double yx, const double yy)
{
	// transposed = the two off-diagonal values change places: an address offset, no select on the values
	const int f = (st.flags >> blk) & 1u;
	st.put(slot, xx); st.put_off(slot + 1, f, xy); st.put_off(slot + 2, -f, yx); st.put(slot + 3, yy);
}

template <class ST>
__host__ __device__ __forceinline__ void prism_steps(const double4* pb, const double2* ub, const double4* pt, const double2* ut,
                                                      const double a0, const double a1, const double a2, const AsmParams& P,
                                                      Iface& bot, Iface& top, ST& st)
{
	TetOut o;
	double C01[4], C02[4], C12[4];
	// ---- tet A = (b0 b1 b2 t2): local 0 1 2 3.  Final afterwards: everything of b2, C22
	tet_eval(pb[0], pb[1], pb[2], pt[2], ub[0], ub[1], ub[2], ut[2], a0, P, o);
	// staged as soon as it is final: the gate (the kernel waits there until the summing warps have released this staging
	// buffer) comes after the element math of tet A -- with two staging buffers it is open long before
	st.gate();
	#pragma unroll
	for (int i = 0; i < 3; ++i) st.put(6 + i, bot.D[6 + i] + o.D[6 + i]);                  // D(b2)
	put_blk(st, 13, 1, bot.E[4] + o.E[4], bot.E[5] + o.E[5], bot.E[6] + o.E[6], bot.E[7] + o.E[7]);           // (0,2) -> E(b0b2)
	put_blk(st, 17, 2, bot.E[8] + o.E[12], bot.E[9] + o.E[13], bot.E[10] + o.E[14], bot.E[11] + o.E[15]);     // (1,2) -> E(b1b2)
	put_blk(st, 47, 8, o.E[20], o.E[21], o.E[22], o.E[23]);                                                   // (2,3) -> C22
	st.put(25, bot.F[4] + o.F[4]); st.put(26, bot.F[5] + o.F[5]);                          // F(b2)
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		bot.E[i] += o.E[i];                                                                // (0,1) -> E(b0b1)
		C02[i] = o.E[8 + i];                                                               // (0,3) -> C02
		C12[i] = o.E[16 + i];                                                              // (1,3) -> C12
	}
	#pragma unroll
	for (int i = 0; i < 6; ++i) bot.D[i] += o.D[i];                                        // D(b0) D(b1)
	#pragma unroll
	for (int i = 0; i < 4; ++i) bot.F[i] += o.F[i];
	top.D[6] = o.D[9]; top.D[7] = o.D[10]; top.D[8] = o.D[11];                             // D(t2)
	top.F[4] = o.F[6]; top.F[5] = o.F[7];
	// ---- tet B = (b0 b1 t1 t2).  Final afterwards: everything of b1, C11, C12
	tet_eval(pb[0], pb[1], pt[1], pt[2], ub[0], ub[1], ut[1], ut[2], a1, P, o);
	#pragma unroll
	for (int i = 0; i < 3; ++i) st.put(3 + i, bot.D[3 + i] + o.D[3 + i]);                 // D(b1)
	put_blk(st, 9, 0, bot.E[0] + o.E[0], bot.E[1] + o.E[1], bot.E[2] + o.E[2], bot.E[3] + o.E[3]);            // (0,1) -> E(b0b1)
	put_blk(st, 39, 6, o.E[12], o.E[13], o.E[14], o.E[15]);                                                   // (1,2) -> C11
	put_blk(st, 43, 7, C12[0] + o.E[16], C12[1] + o.E[17], C12[2] + o.E[18], C12[3] + o.E[19]);               // (1,3) -> C12
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		C01[i] = o.E[4 + i];                                                               // (0,2) -> C01
		C02[i] += o.E[8 + i];                                                              // (0,3) -> C02
		top.E[8 + i] = o.E[20 + i];                                                        // (2,3) -> E(t1t2)
	}
	st.put(23, bot.F[2] + o.F[2]); st.put(24, bot.F[3] + o.F[3]);                          // F(b1)
	bot.D[0] += o.D[0]; bot.D[1] += o.D[1]; bot.D[2] += o.D[2];
	bot.F[0] += o.F[0]; bot.F[1] += o.F[1];
	top.D[3] = o.D[6]; top.D[4] = o.D[7]; top.D[5] = o.D[8];                               // D(t1)
	top.D[6] += o.D[9]; top.D[7] += o.D[10]; top.D[8] += o.D[11];                          // D(t2)
	top.F[2] = o.F[4]; top.F[3] = o.F[5]; top.F[4] += o.F[6]; top.F[5] += o.F[7];
	// ---- tet C = (b0 t0 t1 t2).  Final afterwards: everything of b0, C00, C01, C02
	tet_eval(pb[0], pt[0], pt[1], pt[2], ub[0], ut[0], ut[1], ut[2], a2, P, o);
	#pragma unroll
	for (int i = 0; i < 3; ++i) st.put(i, bot.D[i] + o.D[i]);                             // D(b0)
	st.put(21, bot.F[0] + o.F[0]); st.put(22, bot.F[1] + o.F[1]);                          // F(b0)
	put_blk(st, 27, 3, o.E[0], o.E[1], o.E[2], o.E[3]);                                                       // (0,1) -> C00
	put_blk(st, 31, 4, C01[0] + o.E[4], C01[1] + o.E[5], C01[2] + o.E[6], C01[3] + o.E[7]);                   // (0,2) -> C01
	put_blk(st, 35, 5, C02[0] + o.E[8], C02[1] + o.E[9], C02[2] + o.E[10], C02[3] + o.E[11]);                 // (0,3) -> C02
	#pragma unroll
	for (int i = 0; i < 4; ++i) {
		top.E[i]      = o.E[12 + i];                                                       // (1,2) -> E(t0t1)
		top.E[4 + i]  = o.E[16 + i];                                                       // (1,3) -> E(t0t2)
		top.E[8 + i] += o.E[20 + i];                                                       // (2,3) -> E(t1t2)
	}
	top.D[0] = o.D[3]; top.D[1] = o.D[4]; top.D[2] = o.D[5];                               // D(t0)
	top.D[3] += o.D[6]; top.D[4] += o.D[7]; top.D[5] += o.D[8];                            // D(t1)
	top.D[6] += o.D[9]; top.D[7] += o.D[10]; top.D[8] += o.D[11];                          // D(t2)
	top.F[0] = o.F[2]; top.F[1] = o.F[3]; top.F[2] += o.F[4]; top.F[3] += o.F[5]; top.F[4] += o.F[6]; top.F[5] += o.F[7];
}

template <class ST>
__host__ __device__ __forceinline__ void iface_stage(const Iface& bot, ST& st)       // the top interface of the last prism
{
	st.gate();
	#pragma unroll
	for (int i = 0; i < 9; ++i) st.put(i, bot.D[i]);
	#pragma unroll
	for (int b = 0; b < 3; ++b) put_blk(st, 9 + 4 * b, b, bot.E[4 * b], bot.E[4 * b + 1], bot.E[4 * b + 2], bot.E[4 * b + 3]);
	#pragma unroll
	for (int i = 0; i < 6; ++i) st.put(21 + i, bot.F[i]);
}

// The basal facet of a prism column (its bottom triangle b0 b1 b2) into the accumulators of the bottom interface:
// sliding  1/2 beta u.u dGamma_b  and water pressure  -f_w u.n_h dGamma_bw  (cslvr/momentumbp.py:473-486), exact P1 mass
// integrals -- the expressions of k_facet_rows (bp_kernels.cu).  The outward normal points away from t2, which sits
// straight above b2: n_z < 0.
__host__ __device__ __forceinline__ void facet_fold(const double4* pb, const double2* ub, const double bt0, const double bt1, const double bt2,
                                                     const uint32_t fl, const AsmParams& P, Iface& bot)
{
	const double ax = pb[1].x - pb[0].x, ay = pb[1].y - pb[0].y, az = pb[1].z - pb[0].z;
	const double bx = pb[2].x - pb[0].x, by = pb[2].y - pb[0].y, bz = pb[2].z - pb[0].z;
	double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
	const double nrm = sqrt(nx * nx + ny * ny + nz * nz);
	const double area = 0.5 * nrm;
	const double sgn = (nz > 0.0 ? -1.0 : 1.0) / nrm;
	nx *= sgn; ny *= sgn;
	const double bt[3] = { bt0, bt1, bt2 };
	if (fl & 1u) {
		// K[a][b] = |f| sum_d beta_d int(l_a l_b l_d),  int = {1/10, 1/30, 1/60}
		const double bs = bt[0] + bt[1] + bt[2];
		#pragma unroll
		for (int a = 0; a < 3; ++a) {
			const double Kaa = area * (bt[a] * (1.0 / 10.0) + (bs - bt[a]) * (1.0 / 30.0));
			bot.D[3 * a] += Kaa; bot.D[3 * a + 2] += Kaa;
			if (!P.picard) { bot.F[2 * a] += Kaa * ub[a].x; bot.F[2 * a + 1] += Kaa * ub[a].y; }
			#pragma unroll
			for (int b = a + 1; b < 3; ++b) {
				const double Kab = area * ((bt[a] + bt[b]) * (1.0 / 30.0) + (bs - bt[a] - bt[b]) * (1.0 / 60.0));
				const int pr = (a == 0) ? (b - 1) : 2;
				bot.E[4 * pr] += Kab; bot.E[4 * pr + 3] += Kab;
				if (!P.picard) {
					bot.F[2 * a] += Kab * ub[b].x; bot.F[2 * a + 1] += Kab * ub[b].y;
					bot.F[2 * b] += Kab * ub[a].x; bot.F[2 * b + 1] += Kab * ub[a].y;
				}
			}
		}
	}
	if (fl & 2u) {
		// f_w = rho g (S - z) - rho_sw g D,  D = -min(0, z)   (momentumbp.py:477, d3model.py:858-861)
		double fw[3];
		#pragma unroll
		for (int a = 0; a < 3; ++a) fw[a] = P.rho_g * (pb[a].w - pb[a].z) - P.rho_sw_g * (-fmin(0.0, pb[a].z));
		const double fs = fw[0] + fw[1] + fw[2];
		#pragma unroll
		for (int a = 0; a < 3; ++a) {
			const double m = area * (fs + fw[a]) * (1.0 / 12.0);
			bot.F[2 * a] -= nx * m; bot.F[2 * a + 1] -= ny * m;
		}
	}
}

// add one contribution (a table word) from the staged buffer st[slot][TILE_T]
__host__ __device__ __forceinline__ void entry_add(const double* st, const uint32_t x, double& a0, double& a1, double& a2, double& a3)
{
	const int ty = x >> 14;
	const double* b = st + (x & 0x3fffu);
	const int o1 = (ty == CT_ET) ? 2 : 1, o2 = (ty == CT_E) ? 2 : 1, o3 = (ty == CT_D) ? 2 : (ty == CT_F ? 1 : 3);
	a0 += b[0]; a1 += b[o1 * TILE_T]; a2 += b[o2 * TILE_T]; a3 += b[o3 * TILE_T];
}

// the sum of one table entry: its contributions in table order.  Not inlined on the device: the summing warps call it
// for every entry, inlining it five to nine times only fills the instruction cache
__host__ __device__ __noinline__ double4 entry_total(const double* stage, const uint16_t* w, const uint32_t info)
{
	double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
	const int cnt = info & 31, kind = info >> 6;
	if (info & 32) {                                   // contributions of different types (rare: columns that meet themselves)
		for (int c = 0; c < cnt; ++c) entry_add(stage, w[c], a0, a1, a2, a3);
	} else if (kind == 0) {                            // 2x2 blocks, staged in the orientation this entry wants
		for (int c = 0; c < cnt; ++c) {
			const double* b = stage + (w[c] & 0x3fffu);
			a0 += b[0]; a1 += b[TILE_T]; a2 += b[2 * TILE_T]; a3 += b[3 * TILE_T];
		}
	} else if (kind == 1) {                            // diagonal blocks: symmetric, three numbers staged
		for (int c = 0; c < cnt; ++c) {
			const double* b = stage + (w[c] & 0x3fffu);
			const double xy = b[TILE_T];
			a0 += b[0]; a1 += xy; a2 += xy; a3 += b[2 * TILE_T];
		}
	} else {                                           // F rows
		for (int c = 0; c < cnt; ++c) {
			const double* b = stage + (w[c] & 0x3fffu);
			a0 += b[0]; a1 += b[TILE_T];
		}
	}
	return make_double4(a0, a1, a2, a3);
}

// ------------------------------------------------------------------------------------------------
// the kernel: 8 compute warps (one thread per prism column) + 8 summing warps, registers re-balanced with setmaxnreg
// (TILE_REG_C / TILE_REG_S: the CTA's register pool is conserved), hand-over through named barriers:
//   FULL[b]   compute warps arrive after staging a layer into buffer b, summing warps wait
//   EMPTY[b]  summing warps arrive after reading buffer b, compute warps wait before they stage into it again
//   VDATA     summing warps arrive when the vertex package of the next prism has landed, compute warps wait
//   VREAD     compute warps arrive when they hold a package in registers, summing warps wait before they refill the buffer
// so the sums of layer k overlap the element math of layer k + 1 (TILE_NBUF = 2: all of it).  The vertex data of the next
// layer arrives by cp.async (LDGSTS) into a per-thread shared-memory slot while the current prism is computed -- no global
// loads inside the march; the tile's table of (thread, slot) contributions is copied to shared memory once per CTA.
// No fence in front of the arrivals: bar.arrive / bar.sync order the shared-memory accesses of the threads that take part
// (the producer / consumer pattern of the PTX manual).
#define BAR_FULL  1            // + buffer
#define BAR_EMPTY 3            // + buffer
#define BAR_SUM   5
#define BAR_VDATA 6
#define BAR_VREAD 7
#define TILE_NTHR (TILE_TW + TILE_S)
__device__ __forceinline__ void bar_sync(const int id, const int n)   { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(const int id, const int n) { asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// a 2x2 block = one 32-byte sector of the SELL-32 Jacobian: ONE 256-bit store (STG.E.256, sm_100) instead of two 16-byte
// halves -- half the store instructions and tag look-ups in the LSU the staging traffic shares, no partial sectors
__device__ __forceinline__ void st_block(double* p, const double a, const double b, const double c, const double d)
{
	asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// the summing warps' version of entry_total for entries of one type: the eight table words of an entry arrive with one
// 16-byte load, two contributions are in flight at a time (eight independent loads), additions in table order
__device__ __forceinline__ double4 entry_total_dev(const double* stage, const uint16_t* w, const uint32_t info, const int W)
{
#ifdef TILE_EXP_NODATA
	const int cnt = (info & 31) ? 1 : 0, kind = info >> 6;
#else
	const int cnt = info & 31, kind = info >> 6;                    // (entries of mixed types, info & 32, do not come here)
#endif
	const int s2 = (kind == 0) ? 2 * TILE_T : TILE_T, s3 = (kind == 0) ? 3 * TILE_T : (kind == 1 ? 2 * TILE_T : TILE_T);
	double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
	for (int c0 = 0; c0 < cnt; c0 += 8) {
		const uint4 q = *(const uint4*)(w + c0);
		const uint32_t wd[8] = { q.x & 0x3fffu, (q.x >> 16) & 0x3fffu, q.y & 0x3fffu, (q.y >> 16) & 0x3fffu,
		                         q.z & 0x3fffu, (q.z >> 16) & 0x3fffu, q.w & 0x3fffu, (q.w >> 16) & 0x3fffu };
		#pragma unroll
		for (int c = 0; c < 8; c += 2) {
			if (c0 + c >= cnt) break;
			const bool two = c0 + c + 1 < cnt;
			const double* b0 = stage + wd[c];
			const double* b1 = stage + (two ? wd[c + 1] : 0u);
			const double x0 = b0[0], x1 = b0[TILE_T], x2 = b0[s2], x3 = b0[s3];
			const double y0 = b1[0], y1 = b1[TILE_T], y2 = b1[s2], y3 = b1[s3];
			a0 += x0; a1 += x1; a2 += x2; a3 += x3;
			if (two) { a0 += y0; a1 += y1; a2 += y2; a3 += y3; }
		}
	}
	return make_double4(a0, a1, a2, a3);
}

struct DevStage {
	double* sb; bool wait; int bar; uint32_t flags;
	__device__ __forceinline__ void gate() { if (wait) bar_sync(bar, TILE_NTHR); }
	__device__ __forceinline__ void put(const int slot, const double v) { sb[slot * TILE_T] = v; }
	__device__ __forceinline__ void put_off(const int slot, const int d, const double v) { sb[slot * TILE_T + d * TILE_T] = v; }
};

// -DTILE_TIMING (experiments): clock64 phase timers in one compute and one summing warp of every 400th tile, printed by the kernel
#ifdef TILE_TIMING
#define TILE_CLOCK_BEGIN(n) long long tm[n] = { 0 }, t_q = clock64(), t_start = t_q;
#define TICK(j) { const long long t_ = clock64(); tm[j] += t_ - t_q; t_q = t_; }
#else
#define TILE_CLOCK_BEGIN(n)
#define TICK(j)
#endif

#define TILE_STAGE_BYTES ((size_t)TILE_NSLOT * TILE_T * 8)
#define TILE_VBUF_BYTES  ((size_t)8 * TILE_T * 16)

__global__ void __launch_bounds__(TILE_NTHR, TILE_CTAS)
k_tile(const int tile0, const int NL, const int W, const int32_t* __restrict__ tvert, const int32_t* __restrict__ tnode, const double* __restrict__ tainv,
       const int32_t* __restrict__ ent0, const uint16_t* __restrict__ tab, const uint8_t* __restrict__ einfo, const uint16_t* __restrict__ tflag,
       const int32_t* __restrict__ out, const int32_t* __restrict__ out2, const int64_t n_ent,
       const uint8_t* __restrict__ tfacet, const double* __restrict__ beta,
       const double4* __restrict__ geo, const double2* __restrict__ zS, const double2* __restrict__ u, double* __restrict__ F, double* __restrict__ val, const AsmParams P)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	double* stage = (double*)smem_raw;                                              // [TILE_NBUF][TILE_NSLOT][TILE_T]
	double2* vbuf = (double2*)(smem_raw + TILE_NBUF * TILE_STAGE_BYTES);            // [8][TILE_T]: per prism {z, S}, {u_x, u_y} of the three top vertices, ainv of the three tets
	uint16_t* stab = (uint16_t*)(smem_raw + TILE_NBUF * TILE_STAGE_BYTES + TILE_VBUF_BYTES);   // [entries][W]
	uint8_t* sinfo = (uint8_t*)(stab + TILE_TABW);                                  // [entries]
	const int tile = tile0 + blockIdx.x;
	const int NZ = NL - 1;
	if (threadIdx.x < TILE_TW) {
		// ================================================================= compute warps: no global loads inside the march
		asm volatile("setmaxnreg.inc.sync.aligned.u32 " TILE_REG_C ";");
		const int lane = threadIdx.x & 31;
		// lanes beyond TILE_LANES shadow lane 0 of their warp: same inputs, same arithmetic, the same values stored to the same
		// addresses -- harmless, and no predicate in the march
		const int t = (threadIdx.x >> 5) * TILE_LANES + (lane < TILE_LANES ? lane : 0);
		const uint32_t tflags = tflag[(size_t)tile * TILE_T + t];
		double4 pb[3], pt[3];
		double2 ub[3], ut[3];
		Iface bot, top;
		iface_zero(bot);
		{
			const int32_t* tv = tvert + (size_t)tile * NL * 3 * TILE_T + t;
			const int32_t* tn = tnode + (size_t)tile * NL * 3 * TILE_T + t;
			#pragma unroll
			for (int i = 0; i < 3; ++i) { pb[i] = geo[tv[i * TILE_T]]; ub[i] = u[tn[i * TILE_T]]; }        // layer 0: straight into registers
			const uint32_t ffl = tfacet ? tfacet[(size_t)tile * TILE_T + t] : 0u;
			if (ffl) facet_fold(pb, ub, beta[tv[0]], beta[tv[TILE_T]], beta[tv[2 * TILE_T]], ffl, P, bot);          // the basal facet starts the bottom interface
		}
		TILE_CLOCK_BEGIN(4)
		for (int k = 0; k < NZ; ++k) {
			TICK(3)
			bar_sync(BAR_VDATA, TILE_NTHR);                     // the summing warps have brought in this prism's package
			TICK(0)
			const double2* vb = vbuf + t;
			#pragma unroll
			for (int i = 0; i < 3; ++i) {
				const double2 q0 = vb[(2 * i) * TILE_T];
				pt[i] = make_double4(pb[i].x, pb[i].y, q0.x, q0.y);  // x, y are the column's
				ut[i] = vb[(2 * i + 1) * TILE_T];
			}
			const double2 a01 = vb[6 * TILE_T], a2x = vb[7 * TILE_T];
			if (k + 1 < NZ) bar_arrive(BAR_VREAD, TILE_NTHR);   // the package is in registers: its buffer may be refilled
			const int b = (TILE_NBUF == 2) ? (k & 1) : 0;
			DevStage st = { stage + (size_t)b * TILE_NSLOT * TILE_T + t, k >= TILE_NBUF, BAR_EMPTY + b, tflags };
			TICK(1)
			prism_steps(pb, ub, pt, ut, a01.x, a01.y, a2x.x, P, bot, top, st);
			TICK(2)
			bar_arrive(BAR_FULL + b, TILE_NTHR);
			bot = top;
			#pragma unroll
			for (int i = 0; i < 3; ++i) { pb[i] = pt[i]; ub[i] = ut[i]; }
		}
		const int b = (TILE_NBUF == 2) ? (NZ & 1) : 0;
		DevStage st = { stage + (size_t)b * TILE_NSLOT * TILE_T + t, NZ >= TILE_NBUF, BAR_EMPTY + b, tflags };
		iface_stage(bot, st);
		bar_arrive(BAR_FULL + b, TILE_NTHR);
#ifdef TILE_TIMING
		if (threadIdx.x == 0 && (blockIdx.x % 400) == 7)
			printf("tile %d compute warp: total %lld | wait VDATA %lld  package read %lld  prism (incl. gate) %lld  loop tail %lld\n", (int)blockIdx.x, clock64() - t_start, tm[0], tm[1], tm[2], tm[3]);
#endif
	} else {
		// ================================================================= summing warps (and feeders of the compute warps)
		asm volatile("setmaxnreg.dec.sync.aligned.u32 " TILE_REG_S ";");
		const int sidx = threadIdx.x - TILE_TW;
		const int e_lo = ent0[tile], ne = ent0[tile + 1] - e_lo;
		// package of prism j for compute thread c: {z, S} and u of the three vertices of layer j + 1, ainv of the three tets
		auto feed = [&](const int j) {
			for (int c = sidx; c < TILE_T; c += TILE_S) {
				const int32_t* tv = tvert + ((size_t)tile * NL + j + 1) * 3 * TILE_T + c;
				const int32_t* tn = tnode + ((size_t)tile * NL + j + 1) * 3 * TILE_T + c;
				double2* vb = vbuf + c;
				#pragma unroll
				for (int i = 0; i < 3; ++i) { cp_async16(vb + (2 * i) * TILE_T, &zS[tv[i * TILE_T]]); cp_async16(vb + (2 * i + 1) * TILE_T, &u[tn[i * TILE_T]]); }
				const double* ta = tainv + (((size_t)tile * NZ + j) * TILE_T + c) * 4;
				cp_async16(vb + 6 * TILE_T, ta); cp_async16(vb + 7 * TILE_T, ta + 2);
			}
			asm volatile("cp.async.commit_group;" ::: "memory");
		};
		auto release = [&]() {                                  // the package requested last is complete: hand it to the compute warps
			asm volatile("cp.async.wait_group 0;" ::: "memory");
			bar_arrive(BAR_VDATA, TILE_NTHR);
		};
		if (NZ >= 1) feed(0);
		{	// the tile's table -> shared memory (16-byte pieces)
			const uint4* src = (const uint4*)(tab + (size_t)e_lo * W);
			uint4* dst = (uint4*)stab;
			const int n16 = ne * W / 8;
			for (int i = sidx; i < n16; i += TILE_S) dst[i] = src[i];
			for (int i = sidx; i < ne; i += TILE_S) sinfo[i] = einfo[e_lo + i];
		}
		bar_sync(BAR_SUM, TILE_S);
		if (NZ >= 1) release();                                 // package 0
		if (NZ >= 2) { bar_sync(BAR_VREAD, TILE_NTHR); feed(1); release(); }      // package 1, as soon as package 0 is in registers
		TILE_CLOCK_BEGIN(8)
		for (int k = 0; k <= NZ; ++k) {
			int o[TILE_EPT], o2[TILE_EPT];
			const int32_t* ok = out + (size_t)k * n_ent + e_lo;
			const int32_t* ok2 = out2 + (size_t)k * n_ent + e_lo;
			#pragma unroll
			for (int i = 0; i < TILE_EPT; ++i) { const int e = sidx + i * TILE_S; o[i] = (e < ne) ? ok[e] : -1; o2[i] = (e < ne) ? ok2[e] : -1; }
			const int b = (TILE_NBUF == 2) ? (k & 1) : 0;
			const double* stg = stage + (size_t)b * TILE_NSLOT * TILE_T;
			TICK(0)
			bar_sync(BAR_FULL + b, TILE_NTHR);
			TICK(1)
			// layer k is staged, so the compute warps are starting prism k + 1 and take its package into registers: after the
			// first entry request package k + 2 into the same buffer; it lands while the other entries are summed and is
			// handed over before the last one
			const bool feeding = k + 2 < NZ;
			auto emit = [&](const int oo, const int oo2, const double4 acc) {
#ifdef TILE_EXP_NOSTORE
				if (acc.x == 1.2345e301) F[0] = acc.y + acc.z + acc.w + oo2;
				return;
#endif
				if (oo >= 0) {
					st_block(val + 4 * (size_t)oo, acc.x, acc.y, acc.z, acc.w);
					if (oo2 >= 0) st_block(val + 4 * (size_t)oo2, acc.x, acc.z, acc.y, acc.w);      // J is symmetric: the transposed block of the partner row, same sum
				} else {
					*(double2*)(F + 2 * (size_t)(-oo - 2)) = make_double2(acc.x, acc.y);
				}
			};
			uint32_t mixed = 0;
			#pragma unroll
			for (int i = 0; i < TILE_EPT; ++i) {
				if (i == 1) TICK(2)
				if (i == 1 && feeding) { bar_sync(BAR_VREAD, TILE_NTHR); TICK(3) feed(k + 2); TICK(4) }
				if (i == TILE_EPT - 1) TICK(5)
				if (i == TILE_EPT - 1 && feeding) { release(); TICK(6) }
				if (o[i] == -1) continue;
				const int e = sidx + i * TILE_S;
				const uint32_t info = sinfo[e];
				if (info & 32) { mixed |= 1u << i; continue; }
				emit(o[i], o2[i], entry_total_dev(stg, stab + (size_t)e * W, info, W));
			}
			// rare (columns that meet themselves on tiny periodic meshes): entries whose contributions have different
			// types -- the generic sum, one copy of it, no call in the loop above (a call there spills the summing warps)
			if (mixed) {
				#pragma unroll 1
				for (int i = 0; i < TILE_EPT; ++i) {
					if (!((mixed >> i) & 1u)) continue;
					const int e = sidx + i * TILE_S;
					const uint16_t* w = stab + (size_t)e * W;
					const int cnt = sinfo[e] & 31;
					double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
					for (int c = 0; c < cnt; ++c) entry_add(stg, w[c], a0, a1, a2, a3);
					emit(ok[e], ok2[e], make_double4(a0, a1, a2, a3));
				}
			}
			if (k + TILE_NBUF <= NZ) bar_arrive(BAR_EMPTY + b, TILE_NTHR);
			TICK(7)
		}
#ifdef TILE_TIMING
		if (sidx == 0 && (blockIdx.x % 400) == 7)
			printf("tile %d summing warp: total %lld | top (index loads) %lld  wait FULL %lld  entry 0 %lld  wait VREAD %lld  feed %lld  entries 1.. %lld  release %lld  last entry + arrive %lld\n",
			       (int)blockIdx.x, clock64() - t_start, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7]);
#endif
	}
}

// d_ainv (per tet, mesh order) -> tile order [tile][layer][thread][4] (three tets + padding: two 16-byte pieces per thread)
__global__ void k_tile_ainv(int64_t n, const int32_t* __restrict__ ttet, const double* __restrict__ ainv, double* __restrict__ out)
{
	for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const int j = (int)(i & 3);
		const int64_t tk = i >> 2;                       // (tile * NZ + k) * TILE_T + t
		const int64_t t = tk % TILE_T, lk = tk / TILE_T;
		const int tt = (j < 3) ? ttet[(lk * 3 + j) * TILE_T + t] : -1;
		out[i] = tt >= 0 ? ainv[tt] : 0.0;
	}
}

void bp_tiles_update_ainv(bp_ctx* c)
{
	if (!c->tiles || c->host_only) return;
	bp_tiles* T = c->tiles;
	const int64_t n = (int64_t)T->n_tiles * (T->NL - 1) * TILE_T * 4;
	k_tile_ainv<<<(int)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, c->stream>>>(n, T->d_tet, c->d_ainv, T->d_ainv);
	c->launches++;
}

// the facet lists the facet kernel runs with when the tiled kernel serves the assembly (the basal facets are folded in)
bool bp_tiles_facet_lists(const bp_ctx* c, int64_t* n_fnodes, const int32_t** fn_node, const int32_t** fn_ptr, const int32_t** fn_inc)
{
	if (!c->tiles || !c->tiles->folded) return false;
	*n_fnodes = c->tiles->n_fnodes; *fn_node = c->tiles->d_fn_node; *fn_ptr = c->tiles->d_fn_ptr; *fn_inc = c->tiles->d_fn_inc;
	return true;
}

bool bp_tiles_serve(const bp_ctx* c, int what)
{
	if (!c->tiles || c->prm.reduced_stokes) return false;
	if ((what & (BP_ASSEMBLE_F | BP_ASSEMBLE_J)) != (BP_ASSEMBLE_F | BP_ASSEMBLE_J)) return false;
	if (c->prm.assembly == BP_ASM_SCATTER || c->prm.assembly == BP_ASM_GATHER) return false;
	return c->env.asm_tiled != 0;
}

static size_t tile_smem_bytes() { return TILE_NBUF * TILE_STAGE_BYTES + TILE_VBUF_BYTES + (size_t)TILE_TABW * 2 + TILE_MAXE; }

static void launch_tiles_range(bp_ctx* c, int what, int tile0, int ntile, cudaStream_t stream = nullptr)
{
	if (!stream) stream = c->stream;
	bp_tiles* T = c->tiles;
	AsmParams P = bp_asm_params(c);
	P.picard = (what & 16) ? 1 : 0;
	const size_t sm = tile_smem_bytes();
	if (!T->attr_set) {
		cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
		T->attr_set = true;
	}
	if (ntile <= 0) return;
	k_tile<<<ntile, TILE_NTHR, sm, stream>>>(tile0, T->NL, T->W, T->d_vert, T->d_node, T->d_ainv, T->d_ent0, T->d_tab, T->d_einfo, T->d_tflag, T->d_out, T->d_out2,
	                                                   T->n_ent, T->folded ? T->d_tfacet : nullptr, c->d_beta, c->d_geo, c->d_zS, (const double2*)c->d_u, c->d_F, c->d_val, P);
	c->launches++;
}

void bp_launch_tiles(bp_ctx* c, int what) { launch_tiles_range(c, what, 0, c->tiles->n_tiles); }

// Host-buffer assembly through the tiled kernel, pipelined over three streams (copies in, kernels, copies out; PCIe is
// full duplex): u travels in the node ranges each group of tiles needs, the group's kernel starts as soon as its
// ranges have arrived, and the F rows it completes start back right after it -- the rows a facet integral adds to
// after the facet kernel.  Needs the caller's numbering to be the internal one (no ghosts, identity dof map).
bool bp_tiles_pipeline_ready(const bp_ctx* c, int what)
{
	return bp_tiles_serve(c, what) && !(what & ~(BP_ASSEMBLE_F | BP_ASSEMBLE_J)) && !c->tiles->chunks.empty() && c->identity_dofs &&
	       (c->nn == c->nn_owned || (c->comm && c->n_nbr > 0)) && c->nd_ext == 2 * c->nn && !c->env.no_pipeline;
}

// copy node ranges [first, second) of a dof vector (2 doubles per node) between host and device.  On a layer-major
// numbering the ranges of a group of tiles are one range per layer -- equal length, equal stride: ONE strided (2-D) copy
// instead of six small ones (each small copy costs the copy engine ~5 us of set-up: 25 of them made 16 MB take 0.46 ms
// instead of 0.31 ms).
static cudaError_t copy_ranges(double* dst, const double* src, const std::vector<std::pair<int32_t, int32_t>>& rs, cudaMemcpyKind kind, cudaStream_t st)
{
	size_t i = 0;
	while (i < rs.size()) {
		const int64_t len = rs[i].second - rs[i].first;
		size_t j = i + 1;
		int64_t stride = 0;
		if (j < rs.size() && rs[j].second - rs[j].first == len) {
			stride = rs[j].first - rs[i].first;
			while (j < rs.size() && rs[j].second - rs[j].first == len && rs[j].first - rs[j - 1].first == stride) ++j;
		}
		cudaError_t e;
		if (j - i >= 2 && stride >= len)
			e = cudaMemcpy2DAsync(dst + 2 * (size_t)rs[i].first, (size_t)stride * 16, src + 2 * (size_t)rs[i].first, (size_t)stride * 16, (size_t)len * 16, j - i, kind, st);
		else { j = i + 1; e = cudaMemcpyAsync(dst + 2 * (size_t)rs[i].first, src + 2 * (size_t)rs[i].first, (size_t)len * 16, kind, st); }
		if (e != cudaSuccess) return e;
		i = j;
	}
	return cudaSuccess;
}

__global__ void k_scatter_nodes(int n, const int32_t* __restrict__ nodes, const double2* __restrict__ packed, double2* __restrict__ vec)
{
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) vec[nodes[i]] = packed[i];
}

// issue the pipeline (copies in / group kernels / facet kernel / copies out) on the four streams.  join = true (graph
// capture): the side streams are joined back into c->stream at the end instead of being synchronised by the host.
static int pipeline_issue(bp_ctx* c, int what, const double* u, double* F, bool join)
{
	bp_tiles* T = c->tiles;
	const int K = (int)T->chunks.size();
	cudaEvent_t* e_h = c->pipe_ev.data();
	cudaEvent_t* e_c = c->pipe_ev.data() + 8;
	cudaEvent_t e_f = c->pipe_ev[16], e_0 = c->pipe_ev[17], e_j0 = c->pipe_ev[18], e_j1 = c->pipe_ev[19], e_j2 = c->pipe_ev[20];
	BP_CUDA(c, cudaEventRecord(e_0, c->stream));
	BP_CUDA(c, cudaStreamWaitEvent(c->s_h2d, e_0, 0));
	BP_CUDA(c, cudaStreamWaitEvent(c->s_d2h, e_0, 0));
	const bool part = c->nn > c->nn_owned;
	cudaEvent_t e_halo = c->pipe_ev[21];
	if (part) {
		// partitioned: the boundary values the neighbours need were gathered from the caller's vector on the host
		// (bp_tiles_assemble_pipelined); they go ahead of everything else, the halo exchange runs while the bulk of u is
		// still on its way, and every group of tiles waits for the ghost rows as well as for its own ranges
		if (c->n_send) {
			BP_CUDA(c, cudaMemcpyAsync(c->d_sendbuf, c->h_halo_stage, (size_t)2 * c->n_send * sizeof(double), cudaMemcpyHostToDevice, c->stream));
			k_scatter_nodes<<<(int)((c->n_send + 255) / 256), 256, 0, c->stream>>>((int)c->n_send, c->d_send_nodes, (const double2*)c->d_sendbuf, (double2*)c->d_u);
			c->launches++;
		}
		bp_ctx* cs[1] = { c };
		int rc_ = bp_comm_halo(cs, 1, 0);
		if (rc_) return rc_;
		BP_CUDA(c, cudaEventRecord(e_halo, c->stream));
	}
	for (int k = 0; k < K; ++k) {
		BP_CUDA(c, copy_ranges(c->d_u, u, T->chunks[k].in, cudaMemcpyHostToDevice, c->s_h2d));
		BP_CUDA(c, cudaEventRecord(e_h[k], c->s_h2d));
	}
	// the groups' kernels alternate between two streams: the partial last wave of one overlaps the first wave of the next
	BP_CUDA(c, cudaStreamWaitEvent(c->s_aux, e_0, 0));
	if (part) BP_CUDA(c, cudaStreamWaitEvent(c->s_aux, e_halo, 0));
	for (int k = 0; k < K; ++k) {
		cudaStream_t sk = (k & 1) ? c->s_aux : c->stream;
		BP_CUDA(c, cudaStreamWaitEvent(sk, e_h[k], 0));
		launch_tiles_range(c, what, T->chunks[k].tile0, T->chunks[k].ntile, sk);
		BP_CUDA(c, cudaEventRecord(e_c[k], sk));
	}
	for (int k = 1; k < K; k += 2) BP_CUDA(c, cudaStreamWaitEvent(c->stream, e_c[k], 0));
	bp_launch_facets(c, what);
	BP_CUDA(c, cudaEventRecord(e_f, c->stream));
	if (F) {
		for (int k = 0; k < K; ++k) {
			BP_CUDA(c, cudaStreamWaitEvent(c->s_d2h, e_c[k], 0));
			BP_CUDA(c, copy_ranges(F, c->d_F, T->chunks[k].out_plain, cudaMemcpyDeviceToHost, c->s_d2h));
		}
		BP_CUDA(c, cudaStreamWaitEvent(c->s_d2h, e_f, 0));
		for (int k = 0; k < K; ++k) BP_CUDA(c, copy_ranges(F, c->d_F, T->chunks[k].out_facet, cudaMemcpyDeviceToHost, c->s_d2h));
	}
	if (join) {
		BP_CUDA(c, cudaEventRecord(e_j0, c->s_h2d)); BP_CUDA(c, cudaStreamWaitEvent(c->stream, e_j0, 0));
		BP_CUDA(c, cudaEventRecord(e_j1, c->s_d2h)); BP_CUDA(c, cudaStreamWaitEvent(c->stream, e_j1, 0));
		BP_CUDA(c, cudaEventRecord(e_j2, c->s_aux)); BP_CUDA(c, cudaStreamWaitEvent(c->stream, e_j2, 0));
	}
	return BP_OK;
}

static bool is_pinned_host(const void* p)
{
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost;
}

int bp_tiles_assemble_pipelined(bp_ctx* c, int what, const double* u, double* F)
{
	bp_tiles* T = c->tiles;
	if (!c->s_h2d) {
		BP_CUDA(c, cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
		BP_CUDA(c, cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
		BP_CUDA(c, cudaStreamCreateWithFlags(&c->s_aux, cudaStreamNonBlocking));
		c->pipe_ev.resize(2 * 8 + 6);
		for (auto& e : c->pipe_ev) BP_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	}
	bp_update_ainv(c);
	if (c->nn > c->nn_owned && c->n_send) {            // boundary values for the neighbours: a few thousand nodes, gathered here
		if (!c->h_halo_stage) BP_CUDA(c, cudaMallocHost((void**)&c->h_halo_stage, (size_t)2 * c->n_send * sizeof(double)));
		for (int64_t i = 0; i < c->n_send; ++i) {
			const size_t nd_ = (size_t)c->h_send_nodes[i];
			c->h_halo_stage[2 * i] = u[2 * nd_]; c->h_halo_stage[2 * i + 1] = u[2 * nd_ + 1];
		}
	}
	// The ~40 stream operations of one call cost the host ~0.1 ms -- a seventh of the call at 5 M tets.  With pinned
	// caller buffers the whole pipeline is captured once per (u, F, what, stream) as a CUDA graph over the four streams
	// and replayed with a single launch; pageable buffers (their copies cannot be captured) take the direct path.
	const bool graphable = c->env.pipe_graph && !T->pipe_no_graph && is_pinned_host(u) && (!F || is_pinned_host(F));
	if (graphable) {
		if (T->pipe_graph && (T->pg_u != u || T->pg_F != F || T->pg_what != what || T->pg_stream != c->stream)) { cudaGraphExecDestroy(T->pipe_graph); T->pipe_graph = nullptr; }
		if (!T->pipe_graph) {
			const int64_t l0 = c->launches;
			cudaGraph_t g = nullptr;
			cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
			int rc = BP_OK;
			if (e == cudaSuccess) {
				rc = pipeline_issue(c, what, u, F, true);
				e = cudaStreamEndCapture(c->stream, &g);
				if (rc == BP_OK && e == cudaSuccess) e = cudaGraphInstantiate(&T->pipe_graph, g, 0);
				if (g) cudaGraphDestroy(g);
				if (rc != BP_OK || e != cudaSuccess) T->pipe_graph = nullptr;
			}
			if (!T->pipe_graph) { cudaGetLastError(); T->pipe_no_graph = true; c->launches = l0; }
			else { T->pg_launches = c->launches - l0; c->launches = l0; T->pg_u = u; T->pg_F = F; T->pg_what = what; T->pg_stream = c->stream; }
		}
		if (T->pipe_graph) {
			BP_CUDA(c, cudaGraphLaunch(T->pipe_graph, c->stream));
			c->launches += T->pg_launches;
			if (what & BP_ASSEMBLE_J) c->have_J = true;
			BP_CUDA(c, cudaStreamSynchronize(c->stream));
			BP_CUDA(c, cudaGetLastError());
			return BP_OK;
		}
	}
	// BP_PIPE_TRACE=1 (experiments, direct path): when each stream of the pipeline finished, relative to its start
	static const bool trace = getenv("BP_PIPE_TRACE") != nullptr;
	cudaEvent_t t0 = nullptr, t1 = nullptr, t2 = nullptr, t3 = nullptr;
	if (trace) { cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventCreate(&t2); cudaEventCreate(&t3); cudaEventRecord(t0, c->stream); }
	int rc = pipeline_issue(c, what, u, F, false);
	if (rc) return rc;
	if (trace) {
		cudaEventRecord(t1, c->s_h2d); cudaEventRecord(t2, c->stream); cudaEventRecord(t3, c->s_d2h);
		cudaEventSynchronize(t1); cudaEventSynchronize(t2); cudaEventSynchronize(t3);
		float a = 0, b = 0, d = 0;
		cudaEventElapsedTime(&a, t0, t1); cudaEventElapsedTime(&b, t0, t2); cudaEventElapsedTime(&d, t0, t3);
		fprintf(stderr, "[cslvr_b200] pipeline trace: copies in done %.3f ms, kernels done %.3f ms, copies out done %.3f ms\n", a, b, d);
		cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(t2); cudaEventDestroy(t3);
	}
	if (what & BP_ASSEMBLE_J) c->have_J = true;
	if (F) BP_CUDA(c, cudaStreamSynchronize(c->s_d2h));
	BP_CUDA(c, cudaStreamSynchronize(c->s_aux));
	BP_CUDA(c, cudaStreamSynchronize(c->stream));
	BP_CUDA(c, cudaGetLastError());
	return BP_OK;
}

// bp_set_params: the captured pipeline carries the physical constants as kernel arguments
void bp_tiles_params_changed(bp_ctx* c)
{
	if (c->tiles && c->tiles->pipe_graph) { cudaGraphExecDestroy(c->tiles->pipe_graph); c->tiles->pipe_graph = nullptr; }
}

void bp_tiles_free(bp_ctx* c)
{
	if (!c->tiles) return;
	bp_tiles* T = c->tiles;
	void* ptrs[] = { T->d_vert, T->d_node, T->d_tet, T->d_nthr, T->d_ent0, T->d_out, T->d_out2, T->d_tab, T->d_tflag, T->d_einfo, T->d_ainv, T->d_tfacet, T->d_fn_node, T->d_fn_ptr, T->d_fn_inc };
	for (void* p : ptrs) if (p) cudaFree(p);
	if (T->pipe_graph) cudaGraphExecDestroy(T->pipe_graph);
	delete T;
	c->tiles = nullptr;
}

// ------------------------------------------------------------------------------------------------
// set-up (host)
template <class V>
static int up(bp_ctx* c, V** dst, const std::vector<V>& src)
{
	*dst = nullptr;
	if (c->host_only) return BP_OK;
	const size_t n = std::max<size_t>(src.size(), 1);
	BP_CUDA(c, cudaMalloc((void**)dst, n * sizeof(V)));
	if (!src.empty()) BP_CUDA(c, cudaMemcpy(*dst, src.data(), src.size() * sizeof(V), cudaMemcpyHostToDevice));
	return BP_OK;
}

static int no_tiles(bp_ctx* c, const char* why)
{
	c->tiles_why = why;
	if (c->prm.assembly == BP_ASM_TILED) return bp_fail(c, BP_ERR_ARG, "assembly = BP_ASM_TILED but the tiled cell kernel cannot serve this mesh: %s", why);
	if (c->env.asm_tiled == 1) return bp_fail(c, BP_ERR_ARG, "BP_ASM_TILED=1 but the tiled cell kernel cannot serve this mesh: %s", why);
	return BP_OK;
}

int bp_setup_tiles(bp_ctx* c, const bp_mesh* m, const std::vector<int32_t>& v2n)
{
	c->tiles = nullptr;
	if (c->env.asm_tiled == 0) { c->tiles_why = "disabled (BP_ASM_TILED=0)"; return BP_OK; }
	const int64_t nv = c->nv, nt = c->nt, nn = c->nn, no = c->nn_owned;
	if (no == 0) return no_tiles(c, "no owned nodes");

	bp_stage_timer tt("tiles");
	// ---- 1. vertex columns: vertices with equal (x, y), bottom -> top --------------------------
	double lo[2] = { 1e300, 1e300 }, hi[2] = { -1e300, -1e300 };
	for (int64_t v = 0; v < nv; ++v) for (int d = 0; d < 2; ++d) { lo[d] = std::min(lo[d], m->coords[3 * v + d]); hi[d] = std::max(hi[d], m->coords[3 * v + d]); }
	const double sx_ = (hi[0] > lo[0]) ? 1048576.0 / (hi[0] - lo[0]) : 0.0, sy_ = (hi[1] > lo[1]) ? 1048576.0 / (hi[1] - lo[1]) : 0.0;
	std::vector<int64_t> key(nv);
	for (int64_t v = 0; v < nv; ++v)
		key[v] = ((int64_t)std::llround((m->coords[3 * v] - lo[0]) * sx_) << 22) | (int64_t)std::llround((m->coords[3 * v + 1] - lo[1]) * sy_);
	std::vector<int32_t> ord(nv);
	std::iota(ord.begin(), ord.end(), 0);
	__gnu_parallel::sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) { return key[a] != key[b] ? key[a] < key[b] : (m->coords[3 * a + 2] != m->coords[3 * b + 2] ? m->coords[3 * a + 2] < m->coords[3 * b + 2] : a < b); });
	int NL = 1;
	while (NL < nv && key[ord[NL]] == key[ord[0]]) ++NL;
	if (NL < 2 || nv % NL) return no_tiles(c, "vertices do not form columns of equal length");
	const int NZ = NL - 1;
	const int64_t nvc = nv / NL;
	std::vector<int32_t> vcol(nv), vlay(nv);
	for (int64_t q = 0; q < nvc; ++q)
		for (int k = 0; k < NL; ++k) {
			const int32_t v = ord[q * NL + k];
			if (key[v] != key[ord[q * NL]] || (k && !(m->coords[3 * v + 2] > m->coords[3 * (int64_t)ord[q * NL + k - 1] + 2])))
				return no_tiles(c, "vertices do not form columns of equal length");
			// the kernel carries x, y once per column: they must be the same numbers on every layer
			if (m->coords[3 * (int64_t)v] != m->coords[3 * (int64_t)ord[q * NL]] || m->coords[3 * (int64_t)v + 1] != m->coords[3 * (int64_t)ord[q * NL] + 1])
				return no_tiles(c, "x, y are not bitwise equal along a vertex column");
			vcol[v] = (int32_t)q; vlay[v] = k;
		}
	if (nt % (3 * NZ)) return no_tiles(c, "cell count is not 3 x layers x triangles");

	tt.lap("vertex columns");
	// ---- 2. node columns (periodic images of a vertex column share one node column) -------------
	std::vector<int32_t> ncol(nn, -1), nlay(nn, -1), ncol_of_vcol(nvc), colnode;
	int32_t n_ncol = 0;
	for (int64_t q = 0; q < nvc; ++q) {
		const int32_t n0 = v2n[ord[q * NL]];
		if (ncol[n0] < 0) {
			for (int k = 0; k < NL; ++k) {
				const int32_t nk = v2n[ord[q * NL + k]];
				if (ncol[nk] >= 0) return no_tiles(c, "a node belongs to two columns");
				ncol[nk] = n_ncol; nlay[nk] = k;
				colnode.push_back(nk);
			}
			++n_ncol;
		} else {
			for (int k = 0; k < NL; ++k) {
				const int32_t nk = v2n[ord[q * NL + k]];
				if (ncol[nk] != ncol[n0] || nlay[nk] != k) return no_tiles(c, "periodic images of a column disagree");
			}
		}
		ncol_of_vcol[q] = ncol[n0];
	}
	for (int64_t i = 0; i < nn; ++i) if (ncol[i] < 0) return no_tiles(c, "a node is used by no vertex");
	std::vector<uint8_t> col_owned(n_ncol);
	for (int32_t q = 0; q < n_ncol; ++q) {
		const bool ow = colnode[(size_t)q * NL] < no;
		for (int k = 1; k < NL; ++k) if ((colnode[(size_t)q * NL + k] < no) != ow) return no_tiles(c, "an ice column is split between ranks");
		col_owned[q] = ow;
	}

	tt.lap("node columns");
	// ---- 3. prism columns: every tet lies in one (triangle of vertex columns, layer); 3 tets per prism -------
	struct Rec { int32_t a, b, cc, k, t; };
	std::vector<Rec> rec(nt);
	bool bad = false;
	#pragma omp parallel for schedule(static)
	for (int64_t t = 0; t < nt; ++t) {
		int32_t q[4], kmin = INT32_MAX, kmax = -1;
		for (int a = 0; a < 4; ++a) { const int32_t v = m->cells[t * 4 + a]; q[a] = vcol[v]; kmin = std::min(kmin, vlay[v]); kmax = std::max(kmax, vlay[v]); }
		std::sort(q, q + 4);
		int nu = 1;
		int32_t uq[4] = { q[0], 0, 0, 0 };
		for (int a = 1; a < 4; ++a) if (q[a] != q[a - 1]) uq[nu++] = q[a];
		if (nu != 3 || kmax != kmin + 1) { bad = true; rec[t] = { 0, 0, 0, 0, (int32_t)t }; continue; }
		rec[t] = { uq[0], uq[1], uq[2], kmin, (int32_t)t };
	}
	if (bad) return no_tiles(c, "a cell does not span three vertex columns and two adjacent layers");
	__gnu_parallel::sort(rec.begin(), rec.end(), [](const Rec& x, const Rec& y) {
		if (x.a != y.a) return x.a < y.a; if (x.b != y.b) return x.b < y.b; if (x.cc != y.cc) return x.cc < y.cc; if (x.k != y.k) return x.k < y.k; return x.t < y.t; });
	const int64_t ntri = nt / (3 * NZ);
	std::vector<int32_t> tri_col(3 * ntri), tri_tet((size_t)ntri * NZ * 3);       // canonical vertex columns; tets (A, B, C) per layer
	for (int64_t r = 0; r < ntri && !bad; ++r) {
		const Rec& r0 = rec[r * 3 * NZ];
		const int32_t s[3] = { r0.a, r0.b, r0.cc };
		int perm[3] = { -1, -1, -1 };
		for (int k = 0; k < NZ && !bad; ++k) {
			int cb[3] = { 0, 0, 0 }, ct[3] = { 0, 0, 0 }, bm[3], tm[3];
			for (int j = 0; j < 3; ++j) {
				const Rec& rr = rec[(r * NZ + k) * 3 + j];
				if (rr.a != s[0] || rr.b != s[1] || rr.cc != s[2] || rr.k != k) { bad = true; break; }
				bm[j] = tm[j] = 0;
				for (int a = 0; a < 4; ++a) {
					const int32_t v = m->cells[(int64_t)rr.t * 4 + a];
					const int i = (vcol[v] == s[0]) ? 0 : (vcol[v] == s[1] ? 1 : 2);
					if (vlay[v] == k) { if (bm[j] >> i & 1) bad = true; bm[j] |= 1 << i; cb[i]++; } else { if (tm[j] >> i & 1) bad = true; tm[j] |= 1 << i; ct[i]++; }
				}
			}
			if (bad) break;
			int p[3] = { -1, -1, -1 };
			for (int i = 0; i < 3; ++i) { if (cb[i] == 3 && ct[i] == 1) p[0] = i; else if (cb[i] == 2 && ct[i] == 2) p[1] = i; else if (cb[i] == 1 && ct[i] == 3) p[2] = i; }
			if (p[0] < 0 || p[1] < 0 || p[2] < 0) { bad = true; break; }
			if (k == 0) { perm[0] = p[0]; perm[1] = p[1]; perm[2] = p[2]; }
			else if (p[0] != perm[0] || p[1] != perm[1] || p[2] != perm[2]) { bad = true; break; }
			// the three tets as (b0 b1 b2 t2), (b0 b1 t1 t2), (b0 t0 t1 t2)
			const int mA_b = 7, mA_t = 1 << p[2], mB_b = (1 << p[0]) | (1 << p[1]), mB_t = (1 << p[1]) | (1 << p[2]), mC_b = 1 << p[0], mC_t = 7;
			int found[3] = { -1, -1, -1 };
			for (int j = 0; j < 3; ++j) {
				if (bm[j] == mA_b && tm[j] == mA_t) found[0] = j; else if (bm[j] == mB_b && tm[j] == mB_t) found[1] = j; else if (bm[j] == mC_b && tm[j] == mC_t) found[2] = j;
			}
			if (found[0] < 0 || found[1] < 0 || found[2] < 0) { bad = true; break; }
			for (int j = 0; j < 3; ++j) tri_tet[((size_t)r * NZ + k) * 3 + j] = rec[(r * NZ + k) * 3 + found[j]].t;
		}
		for (int i = 0; i < 3 && !bad; ++i) tri_col[3 * r + i] = s[perm[i]];
	}
	if (bad) return no_tiles(c, "prisms are not split into three tets the same way along a column");
	{ std::vector<Rec>().swap(rec); }

	tt.lap("prism columns");
	// ---- 3b. basal facets = bottom triangles of prism columns: folded into the march (see bp_tiles::d_tfacet) ----
	// facet (cell, local facet) is the bottom of column r iff the cell is r's tet A of layer 0 and the vertex opposite the
	// facet is its t2 (layer 1).  Everything else (the lateral water-pressure facets) stays with the facet kernel.
	std::vector<uint8_t> colflag(ntri, 0), facet_folded(c->h_fcell.size(), 0);
	const bool fold = !c->env.facet_atomic && !c->prm.reduced_stokes;
	if (fold) {
		std::vector<int32_t> tetA_col(nt, -1);
		for (int64_t r = 0; r < ntri; ++r) tetA_col[tri_tet[(size_t)r * NZ * 3]] = (int32_t)r;
		for (size_t f = 0; f < c->h_fcell.size(); ++f) {
			const int32_t cell = c->h_fcell[f], info = c->h_finfo[f], lf = info & 0xff;
			const int32_t r = tetA_col[cell];
			if (r < 0 || vlay[m->cells[(int64_t)cell * 4 + lf]] != 1) continue;
			colflag[r] |= (uint8_t)(((info >> 8) & 1) | (((info >> 9) & 1) << 1));
			facet_folded[f] = 1;
		}
	}

	tt.lap("facet fold");
	// ---- 4. tiles of owned ice columns: compact footprint patches, at most TILE_T prism columns touch one ----
	std::vector<int64_t> ctp(n_ncol + 1, 0);                   // node column -> prism columns touching it
	for (int64_t r = 0; r < ntri; ++r) {
		int32_t q[3] = { ncol_of_vcol[tri_col[3 * r]], ncol_of_vcol[tri_col[3 * r + 1]], ncol_of_vcol[tri_col[3 * r + 2]] };
		for (int i = 0; i < 3; ++i) { bool dup = false; for (int j = 0; j < i; ++j) dup |= q[j] == q[i]; if (!dup) ctp[q[i] + 1]++; }
	}
	for (int32_t q = 0; q < n_ncol; ++q) ctp[q + 1] += ctp[q];
	std::vector<int32_t> ctl(ctp[n_ncol]);
	{
		std::vector<int64_t> fill(ctp.begin(), ctp.end() - 1);
		for (int64_t r = 0; r < ntri; ++r) {
			int32_t q[3] = { ncol_of_vcol[tri_col[3 * r]], ncol_of_vcol[tri_col[3 * r + 1]], ncol_of_vcol[tri_col[3 * r + 2]] };
			for (int i = 0; i < 3; ++i) { bool dup = false; for (int j = 0; j < i; ++j) dup |= q[j] == q[i]; if (!dup) ctl[fill[q[i]]++] = (int32_t)r; }
		}
	}
	for (int32_t q = 0; q < n_ncol; ++q) if (col_owned[q] && ctp[q + 1] - ctp[q] > TILE_T) return no_tiles(c, "an ice column is touched by more prism columns than a tile holds");
	// relations of every owned column: (kind, neighbour column) pairs, kind 0 same layer / 1 layer above / 2 layer below.
	// A tile's table has one entry per diagonal block and per F row of its columns, one per relation that leaves the tile,
	// and ONE per pair of relations inside the tile (J is symmetric: same sum, two outputs).
	struct PK { int32_t p, q; int8_t kind; };
	std::vector<PK> pk;
	std::vector<int64_t> pkp(n_ncol + 1, 0);
	int max_val = 0;
	{
		pk.reserve((size_t)ntri * 21);
		for (int64_t r = 0; r < ntri; ++r) {
			int32_t q[3] = { ncol_of_vcol[tri_col[3 * r]], ncol_of_vcol[tri_col[3 * r + 1]], ncol_of_vcol[tri_col[3 * r + 2]] };
			for (int i = 0; i < 3; ++i) {
				if (!col_owned[q[i]]) continue;
				for (int j = 0; j < 3; ++j) {
					if (q[j] != q[i]) pk.push_back({ q[i], q[j], 0 });
					if (j >= i) pk.push_back({ q[i], q[j], 1 });
					if (j <= i) pk.push_back({ q[i], q[j], 2 });
				}
			}
		}
		__gnu_parallel::sort(pk.begin(), pk.end(), [](const PK& a, const PK& b) { if (a.p != b.p) return a.p < b.p; if (a.kind != b.kind) return a.kind < b.kind; return a.q < b.q; });
		pk.erase(std::unique(pk.begin(), pk.end(), [](const PK& a, const PK& b) { return a.p == b.p && a.kind == b.kind && a.q == b.q; }), pk.end());
		for (const PK& e : pk) pkp[e.p + 1]++;
		for (int32_t q = 0; q < n_ncol; ++q) pkp[q + 1] += pkp[q];
		for (int32_t q = 0; q < n_ncol; ++q) if (col_owned[q]) max_val = std::max(max_val, (int)(ctp[q + 1] - ctp[q]));
	}
	int W = (max_val <= 8) ? 8 : 16;
	const int max_ent = std::min<int>(TILE_MAXE, TILE_TABW / W);
	// position of a node column: its lowest vertex (the master of periodic images)
	std::vector<int32_t> rep(nn, -1);
	for (int64_t v = nv - 1; v >= 0; --v) rep[v2n[v]] = (int32_t)v;
	std::vector<int32_t> own_cols;
	double olo[2] = { 1e300, 1e300 }, ohi[2] = { -1e300, -1e300 };
	std::vector<double> cx(n_ncol), cy(n_ncol);
	for (int32_t q = 0; q < n_ncol; ++q) {
		const int32_t v = rep[colnode[(size_t)q * NL]];
		cx[q] = m->coords[3 * (int64_t)v]; cy[q] = m->coords[3 * (int64_t)v + 1];
		if (!col_owned[q]) continue;
		own_cols.push_back(q);
		olo[0] = std::min(olo[0], cx[q]); ohi[0] = std::max(ohi[0], cx[q]); olo[1] = std::min(olo[1], cy[q]); ohi[1] = std::max(ohi[1], cy[q]);
	}
	const double area = std::max(ohi[0] - olo[0], 1e-300) * std::max(ohi[1] - olo[1], 1e-300);
	const double rho = (double)own_cols.size() / area;
	const double strip_h = 7.0 / std::sqrt(rho);                  // ~7 rows of columns per strip: (7 + 1) x (15 + 1) x 2 = 256 prism columns (x 8 for 128)
	std::vector<int64_t> ybin(n_ncol, 0);
	for (int32_t q : own_cols) ybin[q] = (int64_t)std::floor((cy[q] - olo[1]) / strip_h + 1e-9);
	__gnu_parallel::sort(own_cols.begin(), own_cols.end(), [&](int32_t a, int32_t b) {
		if (ybin[a] != ybin[b]) return ybin[a] < ybin[b]; if (cx[a] != cx[b]) return cx[a] < cx[b]; if (cy[a] != cy[b]) return cy[a] < cy[b]; return a < b; });
	std::vector<int32_t> tile_c0(1, 0);                            // tile -> range of own_cols
	{
		std::vector<int32_t> stamp(ntri, -1), in_tile(n_ncol, -1);
		int ntile = 0, ntr = 0, ncl = 0, nen = 0;
		// entries the tile's table grows by when column q joins it (see above)
		auto grow = [&](int32_t q) {
			int d = 2;
			for (int64_t k = pkp[q]; k < pkp[q + 1]; ++k) {
				const bool inside = in_tile[pk[k].q] == ntile || pk[k].q == q;
				if (pk[k].kind == 0) d += inside ? 0 : 1;              // the pair entry was counted when the partner joined (as its relation leaving the tile)
				else if (pk[k].kind == 1) d += 1;                      // own "above" relations: always one entry
				else d += inside ? 0 : 1;                              // "below" relations inside the tile ride on the partner's "above" entry
			}
			// relations of columns already inside that pointed below at q were counted as leaving the tile: now they ride on q's "above" entries
			for (int64_t k = pkp[q]; k < pkp[q + 1]; ++k)
				if (pk[k].kind == 1 && pk[k].q != q && in_tile[pk[k].q] == ntile) d -= 1;
			return d;
		};
		for (size_t i = 0; i < own_cols.size(); ++i) {
			const int32_t q = own_cols[i];
			int add = 0;
			for (int64_t k = ctp[q]; k < ctp[q + 1]; ++k) if (stamp[ctl[k]] != ntile) ++add;
			const bool new_strip = i > 0 && ybin[q] != ybin[own_cols[i - 1]];
			int de = grow(q);
			if (pkp[q + 1] - pkp[q] + 2 > max_ent) return no_tiles(c, "an ice column has more Jacobian blocks than a tile's table holds");
			if (ncl > 0 && (ntr + add > TILE_T || ncl + 1 > TILE_MAXC || nen + de > max_ent || new_strip)) {
				tile_c0.push_back((int32_t)i); ++ntile; ntr = 0; ncl = 0; nen = 0;
				add = (int)(ctp[q + 1] - ctp[q]);
				de = grow(q);
			}
			for (int64_t k = ctp[q]; k < ctp[q + 1]; ++k) stamp[ctl[k]] = ntile;
			in_tile[q] = ntile;
			ntr += add; ++ncl; nen += de;
		}
		tile_c0.push_back((int32_t)own_cols.size());
	}
	const int n_tiles = (int)tile_c0.size() - 1;

	tt.lap("tiling");
	// ---- 5 + 6. per tile: thread list (prism columns), vertex / node / tet tables, sum tables -----------------
	bp_tiles* T = new bp_tiles();
	T->n_tiles = n_tiles; T->NL = NL; T->n_prism_cols = ntri; T->W = W;
	T->h_vert.assign((size_t)n_tiles * NL * 3 * TILE_T, 0);
	T->h_node.assign((size_t)n_tiles * NL * 3 * TILE_T, 0);
	T->h_tet.assign((size_t)n_tiles * NZ * 3 * TILE_T, -1);
	T->h_nthr.assign(n_tiles, 0);
	T->h_tfacet.assign((size_t)n_tiles * TILE_T, 0);
	T->folded = fold;
	T->h_ent0.assign(n_tiles + 1, 0);
	struct Con { int8_t kind; int64_t delta; int32_t p, q; int16_t t; uint8_t code; };
	struct Ent { int8_t kind, mirror; int32_t p, q; };       // mirror: -1 none, 0 / 2 the transposed block is the partner's same-layer / below block
	std::vector<std::vector<Ent>> tile_ent(n_tiles);
	std::vector<std::vector<uint32_t>> tile_eptr(n_tiles);
	std::vector<std::vector<uint16_t>> tile_con(n_tiles);
	std::vector<std::vector<int32_t>> tile_cols(n_tiles);
	int fail = 0;
	#pragma omp parallel
	{
		std::vector<int32_t> local(n_ncol, -1), tris;
		std::vector<Con> cons;
		#pragma omp for schedule(dynamic, 8)
		for (int tl = 0; tl < n_tiles; ++tl) {
			// the tile's columns, ordered by the id of their bottom node: consecutive rows -> consecutive SELL slots
			std::vector<int32_t>& cols = tile_cols[tl];
			cols.assign(own_cols.begin() + tile_c0[tl], own_cols.begin() + tile_c0[tl + 1]);
			std::sort(cols.begin(), cols.end(), [&](int32_t a, int32_t b) { return colnode[(size_t)a * NL] < colnode[(size_t)b * NL]; });
			for (size_t p = 0; p < cols.size(); ++p) local[cols[p]] = (int32_t)p;
			tris.clear();
			for (int32_t q : cols) for (int64_t k = ctp[q]; k < ctp[q + 1]; ++k) tris.push_back(ctl[k]);
			std::sort(tris.begin(), tris.end());
			tris.erase(std::unique(tris.begin(), tris.end()), tris.end());
			if ((int)tris.size() > TILE_T) { fail = 1; for (int32_t q : cols) local[q] = -1; continue; }
			// thread order: prism columns of one orientation class next to each other, then by position -- the lanes of a
			// warp then read staged slots of neighbouring threads when they sum neighbouring rows
			auto cls = [&](int32_t r) {
				const int32_t a = tri_col[3 * r], b = tri_col[3 * r + 1], d = tri_col[3 * r + 2];
				const double xa = m->coords[3 * (int64_t)ord[(int64_t)a * NL]], xb = m->coords[3 * (int64_t)ord[(int64_t)b * NL]], xd = m->coords[3 * (int64_t)ord[(int64_t)d * NL]];
				const double ya = m->coords[3 * (int64_t)ord[(int64_t)a * NL] + 1], yb = m->coords[3 * (int64_t)ord[(int64_t)b * NL] + 1], yd = m->coords[3 * (int64_t)ord[(int64_t)d * NL] + 1];
				return ((xb > xa) ? 1 : 0) | ((xd > xa) ? 2 : 0) | ((yb > ya) ? 4 : 0) | ((yd > ya) ? 8 : 0) | ((xd > xb) ? 16 : 0) | ((yd > yb) ? 32 : 0);
			};
			auto pos = [&](int32_t r, int d) { return m->coords[3 * (int64_t)ord[(int64_t)tri_col[3 * r] * NL] + d]; };
			std::sort(tris.begin(), tris.end(), [&](int32_t a, int32_t b) {
				const int ca = cls(a), cb2 = cls(b);
				if (ca != cb2) return ca < cb2;
				if (pos(a, 1) != pos(b, 1)) return pos(a, 1) < pos(b, 1);
				if (pos(a, 0) != pos(b, 0)) return pos(a, 0) < pos(b, 0);
				return a < b; });
			T->h_nthr[tl] = (int32_t)tris.size();
			cons.clear();
			for (size_t t = 0; t < tris.size(); ++t) {
				const int32_t r = tris[t];
				T->h_tfacet[(size_t)tl * TILE_T + t] = colflag[r];
				int32_t nc[3];
				for (int i = 0; i < 3; ++i) {
					const int32_t vc = tri_col[3 * r + i];
					nc[i] = ncol_of_vcol[vc];
					for (int k = 0; k < NL; ++k) {
						const int32_t v = ord[(int64_t)vc * NL + k];
						T->h_vert[(((size_t)tl * NL + k) * 3 + i) * TILE_T + t] = v;
						T->h_node[(((size_t)tl * NL + k) * 3 + i) * TILE_T + t] = v2n[v];
					}
				}
				for (int k = 0; k < NZ; ++k) for (int j = 0; j < 3; ++j)
					T->h_tet[(((size_t)tl * NZ + k) * 3 + j) * TILE_T + t] = tri_tet[((size_t)r * NZ + k) * 3 + j];
				for (int i = 0; i < 3; ++i) {
					const int32_t p = local[nc[i]];
					if (p < 0) continue;                                   // not a column of this tile
					auto dl = [&](int32_t q) { return (int64_t)colnode[(size_t)q * NL] - (int64_t)colnode[(size_t)nc[i] * NL]; };
					cons.push_back({ 0, 0, p, nc[i], (int16_t)t, (uint8_t)i });                            // D(b_i)
					cons.push_back({ 3, 0, p, nc[i], (int16_t)t, (uint8_t)(21 + i) });                     // F(b_i)
					for (int j = 0; j < 3; ++j) {
						if (j != i) cons.push_back({ 0, dl(nc[j]), p, nc[j], (int16_t)t, (uint8_t)(i < j ? 3 + pair_E(i, j) : 6 + pair_E(j, i)) });
						if (j >= i) cons.push_back({ 1, dl(nc[j]), p, nc[j], (int16_t)t, (uint8_t)(9 + pair_C(i, j)) });      // row (p, k), col (q, k+1)
						if (j <= i) cons.push_back({ 2, dl(nc[j]), p, nc[j], (int16_t)t, (uint8_t)(15 + pair_C(j, i)) });     // row (p, k+1), col (q, k)
					}
				}
			}
			std::sort(cons.begin(), cons.end(), [](const Con& a, const Con& b) {
				if (a.kind != b.kind) return a.kind < b.kind; if (a.delta != b.delta) return a.delta < b.delta; if (a.p != b.p) return a.p < b.p;
				if (a.q != b.q) return a.q < b.q; if (a.t != b.t) return a.t < b.t; return a.code < b.code; });
			std::vector<Ent>& E = tile_ent[tl];
			std::vector<uint32_t>& ep = tile_eptr[tl];
			std::vector<uint16_t>& cc = tile_con[tl];
			bool keep = false;
			for (size_t i = 0; i < cons.size(); ++i) {
				if (i == 0 || cons[i].kind != cons[i - 1].kind || cons[i].p != cons[i - 1].p || cons[i].q != cons[i - 1].q) {
					const int8_t kind = cons[i].kind;
					const int32_t lq = local[cons[i].q];
					int8_t mirror = -1;
					keep = true;
					if (kind == 0 && cons[i].q != cols[cons[i].p] && lq >= 0) { if (cons[i].p < lq) mirror = 0; else keep = false; }
					else if (kind == 1 && lq >= 0) mirror = 2;
					else if (kind == 2 && lq >= 0) keep = false;
					if (keep) { E.push_back({ kind, mirror, cons[i].p, cons[i].q }); ep.push_back((uint32_t)cc.size()); }
				}
				if (keep) cc.push_back((uint16_t)((uint16_t)cons[i].t | ((uint16_t)cons[i].code << 9)));
			}
			ep.push_back((uint32_t)cc.size());
			for (int32_t q : cols) local[q] = -1;
		}
	}
	if (fail) { delete T; return no_tiles(c, "a tile is touched by more prism columns than it holds"); }
	int64_t ne = 0, ncn = 0;
	for (int tl = 0; tl < n_tiles; ++tl) { T->h_ent0[tl] = (int32_t)ne; ne += (int64_t)tile_ent[tl].size(); ncn += (int64_t)tile_con[tl].size(); T->n_threads_used += T->h_nthr[tl]; }
	T->h_ent0[n_tiles] = (int32_t)ne;
	if (ne > INT32_MAX / 32) { delete T; return no_tiles(c, "too many table entries for one rank"); }
	T->n_ent = ne; T->n_contrib = ncn;
	T->h_out.assign((size_t)NL * ne, -1);
	T->h_out2.assign((size_t)NL * ne, -1);
	{	// the packed table: W words per entry, contributions in ascending thread order, padded with CW_NONE
		int maxc = 0;
		for (int tl = 0; tl < n_tiles; ++tl)
			for (size_t e = 0; e < tile_ent[tl].size(); ++e) maxc = std::max(maxc, (int)(tile_eptr[tl][e + 1] - tile_eptr[tl][e]));
		bool fits = maxc <= W;
		for (int tl = 0; tl < n_tiles && fits; ++tl) fits = (int64_t)tile_ent[tl].size() <= std::min<int64_t>(TILE_MAXE, TILE_TABW / W);
		if (!fits) { delete T; return no_tiles(c, "a Jacobian block gets more contributions than a table entry holds"); }
		T->h_tab.assign((size_t)ne * W, (uint16_t)CW_NONE);
		T->h_einfo.assign((size_t)ne, 0);
		T->h_tflag.assign((size_t)n_tiles * TILE_T, 0);
		std::vector<uint16_t> seen((size_t)TILE_T);
		for (int tl = 0; tl < n_tiles; ++tl) {
			std::fill(seen.begin(), seen.end(), 0);
			uint16_t* flg = T->h_tflag.data() + (size_t)tl * TILE_T;
			for (size_t e = 0; e < tile_ent[tl].size(); ++e) {
				uint16_t* dst = T->h_tab.data() + ((size_t)T->h_ent0[tl] + e) * W;
				const int cnt = (int)(tile_eptr[tl][e + 1] - tile_eptr[tl][e]);
				int types = 0;
				for (uint32_t q = tile_eptr[tl][e]; q < tile_eptr[tl][e + 1]; ++q) {
					const uint16_t w = tile_con[tl][q];
					const int t = w & 511, code = w >> 9;
					int ty = H_CODE_TYPE[code];
					if (ty == CT_E || ty == CT_ET) {
						// the staged block: E01 E02 E12 = 0..2, C00..C22 = 3..8; its first consumer decides how the thread stages it
						const int blk = (code < 9) ? (code - 3) % 3 : 3 + (code - 9) % 6;
						const bool want_t = (ty == CT_ET);
						if (!((seen[t] >> blk) & 1)) { seen[t] |= (uint16_t)(1u << blk); if (want_t) flg[t] |= (uint16_t)(1u << blk); }
						ty = (want_t == (bool)((flg[t] >> blk) & 1)) ? CT_E : CT_ET;
					}
					types |= 1 << ty;
					*dst++ = CW(t, H_CODE_BASE[code], ty);
				}
				const int kind = (types == (1 << CT_E)) ? 0 : (types == (1 << CT_D)) ? 1 : (types == (1 << CT_F)) ? 2 : 3;
				T->h_einfo[(size_t)T->h_ent0[tl] + e] = (uint8_t)(cnt | (kind == 3 ? 32 : 0) | ((kind & 3) << 6));
			}
		}
	}
	// output slots per layer, and the exactly-once check
	std::vector<uint8_t> hitJ((size_t)std::max<int64_t>(c->n_slots, 1), 0), hitF((size_t)no, 0);
	int miss = 0;
	#pragma omp parallel for schedule(dynamic, 8)
	for (int tl = 0; tl < n_tiles; ++tl) {
		const std::vector<int32_t>& cols = tile_cols[tl];
		for (size_t e = 0; e < tile_ent[tl].size(); ++e) {
			const Ent& en = tile_ent[tl][e];
			const int32_t pc = cols[en.p];
			for (int k = 0; k < NL; ++k) {
				int32_t o = -1;
				if (en.kind == 3) o = -(colnode[(size_t)pc * NL + k] + 2);
				else {
					if (en.kind != 0 && k == NL - 1) continue;
					const int32_t i = colnode[(size_t)pc * NL + k + (en.kind == 2 ? 1 : 0)];
					const int32_t j = colnode[(size_t)en.q * NL + k + (en.kind == 1 ? 1 : 0)];
					const int32_t* lo_ = c->h_col.data() + c->h_rowptr[i];
					const int32_t* hi_ = c->h_col.data() + c->h_rowptr[i + 1];
					const int32_t* it = std::lower_bound(lo_, hi_, j);
					if (it == hi_ || *it != j) { miss = 1; continue; }
					o = c->h_sell_ptr[i >> 5] + 32 * (int32_t)(it - lo_) + (i & 31);
				}
				T->h_out[(size_t)k * ne + T->h_ent0[tl] + e] = o;
				if (o >= 0) {
					#pragma omp atomic
					hitJ[o]++;
				} else {
					#pragma omp atomic
					hitF[-o - 2]++;
				}
				if (en.mirror >= 0 && o >= 0) {             // the transposed block: row = the partner's node, column = this column's node
					const int32_t i2 = colnode[(size_t)en.q * NL + k + (en.mirror == 2 ? 1 : 0)];
					const int32_t j2 = colnode[(size_t)pc * NL + k];
					const int32_t* lo2 = c->h_col.data() + c->h_rowptr[i2];
					const int32_t* hi2 = c->h_col.data() + c->h_rowptr[i2 + 1];
					const int32_t* it2 = std::lower_bound(lo2, hi2, j2);
					if (i2 >= no || it2 == hi2 || *it2 != j2) { miss = 1; continue; }
					const int32_t om = c->h_sell_ptr[i2 >> 5] + 32 * (int32_t)(it2 - lo2) + (i2 & 31);
					T->h_out2[(size_t)k * ne + T->h_ent0[tl] + e] = om;
					#pragma omp atomic
					hitJ[om]++;
				}
			}
		}
	}
	if (!miss) {
		for (int64_t i = 0; i < no && !miss; ++i) {
			if (hitF[i] != 1) miss = 1;
			for (int32_t k = 0; k < c->h_rowptr[i + 1] - c->h_rowptr[i]; ++k)
				if (hitJ[(size_t)c->h_sell_ptr[i >> 5] + 32 * k + (i & 31)] != 1) { miss = 1; break; }
		}
	}
	if (miss) { delete T; return no_tiles(c, "internal: the tile tables do not cover every Jacobian block exactly once"); }

	tt.lap("tables");
	// ---- the facets that are left for the facet kernel: row-gather lists as in bp_setup.cu, without the folded ones
	std::vector<int32_t> fn_node, fn_ptr(1, 0), fn_inc;
	T->h_facet_node.assign((size_t)std::max<int64_t>(no, 1), 0);
	if (fold) {
		static const int LOCG[4][3] = { { 1, 2, 3 }, { 0, 2, 3 }, { 0, 1, 3 }, { 0, 1, 2 } };
		std::vector<int32_t> cntn(no, 0), slot_of_node(no, -1);
		auto node_of = [&](size_t f, int k) { return v2n[m->cells[(int64_t)c->h_fcell[f] * 4 + LOCG[c->h_finfo[f] & 0xff][k]]]; };
		for (size_t f = 0; f < c->h_fcell.size(); ++f) if (!facet_folded[f])
			for (int k = 0; k < 3; ++k) { const int32_t nd_ = node_of(f, k); if (nd_ < no) cntn[nd_]++; }
		for (int64_t i = 0; i < no; ++i) if (cntn[i]) { slot_of_node[i] = (int32_t)fn_node.size(); fn_node.push_back((int32_t)i); fn_ptr.push_back(fn_ptr.back() + cntn[i]); T->h_facet_node[i] = 1; }
		fn_inc.assign((size_t)std::max<int32_t>(fn_ptr.back(), 1), 0);
		std::vector<int32_t> fillf(fn_node.size(), 0);
		for (size_t f = 0; f < c->h_fcell.size(); ++f) if (!facet_folded[f])
			for (int k = 0; k < 3; ++k) {
				const int32_t nd_ = node_of(f, k);
				if (nd_ >= no) continue;
				const int32_t r = slot_of_node[nd_];
				fn_inc[fn_ptr[r] + fillf[r]++] = (int32_t)(f * 4 + k);
			}
		T->n_fnodes = (int64_t)fn_node.size();
	}
	const std::vector<uint8_t>& facet_node = fold ? T->h_facet_node : c->h_facet_node;

	tt.lap("facet lists");
	// ---- groups of tiles for the pipelined host-buffer path
	// (a partitioned context too: its ghost nodes -- the tail of the numbering -- come from the neighbours, not from the host)
	if (!c->host_only && c->identity_dofs && (nn == no || c->n_nbr > 0)) {
		const int K = std::max(1, std::min(c->env.pipe_k, std::min(8, n_tiles / 148)));
		int n_sm = 0;
		{ int dev = 0; if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); cudaGetLastError(); }
		std::vector<uint8_t> have(nn, 0), mark(nn, 0);
		for (int64_t i = no; i < nn; ++i) have[i] = 1;
		const int GAP = 128;                                  // ranges closer than this are copied as one (the nodes between are needed later anyway)
		// group boundaries sit where a strip of tiles ends (tiles are ordered strip by strip): the node ranges of a group are
		// then whole rows of columns per layer -- a handful of large copies instead of one per row of a cut strip
		std::vector<int> bound(K + 1, 0);
		bound[K] = n_tiles;
		for (int k = 1; k < K; ++k) {
			int b = (int)((int64_t)k * n_tiles / K);
			// whole waves: a group of tiles that ends just below a multiple of the SM count wastes no partial wave
			if (n_sm > 0 && b >= n_sm) b = (int)std::lround((double)b / n_sm) * n_sm;
			int lo_b = b, hi_b = b;
			auto strip_of = [&](int tl) { return ybin[own_cols[tile_c0[tl]]]; };
			while (lo_b > bound[k - 1] + 1 && strip_of(lo_b) == strip_of(lo_b - 1)) --lo_b;
			while (hi_b < n_tiles && strip_of(hi_b) == strip_of(hi_b - 1)) ++hi_b;
			b = (lo_b > bound[k - 1] && (n_sm > 0 || b - lo_b <= hi_b - b)) ? lo_b : hi_b;      // at or just below the wave boundary
			bound[k] = std::min(std::max(b, bound[k - 1]), n_tiles);
		}
		T->chunks.resize(K);
		for (int k = 0; k < K; ++k) {
			bp_tiles::Chunk& ch = T->chunks[k];
			ch.tile0 = bound[k];
			ch.ntile = bound[k + 1] - bound[k];
			std::fill(mark.begin(), mark.end(), 0);
			for (size_t i = (size_t)ch.tile0 * NL * 3 * TILE_T; i < (size_t)(ch.tile0 + ch.ntile) * NL * 3 * TILE_T; ++i) mark[T->h_node[i]] = 1;
			for (int64_t i = 0; i < nn; ) {
				if (!mark[i] || have[i]) { ++i; continue; }
				int64_t j = i, last = i;                          // extend over needed or not-yet-present nodes, never over present ones
				while (j < nn && !have[j] && j - last <= GAP) { if (mark[j]) last = j; ++j; }
				ch.in.push_back({ (int32_t)i, (int32_t)(last + 1) });
				for (int64_t q = i; q <= last; ++q) have[q] = 1;
				i = last + 1;
			}
			std::fill(mark.begin(), mark.end(), 0);
			for (int tl = ch.tile0; tl < ch.tile0 + ch.ntile; ++tl)
				for (int32_t q : tile_cols[tl]) for (int kk = 0; kk < NL; ++kk) mark[colnode[(size_t)q * NL + kk]] = 1;
			for (int64_t i = 0; i < nn; ) {
				if (!mark[i]) { ++i; continue; }
				int64_t j = i;
				bool facet = false;
				while (j < nn && mark[j]) { facet |= (j < (int64_t)facet_node.size() && facet_node[j]); ++j; }
				(facet ? ch.out_facet : ch.out_plain).push_back({ (int32_t)i, (int32_t)j });
				i = j;
			}
		}
		size_t nr = 0;
		for (auto& ch : T->chunks) nr += ch.in.size() + ch.out_plain.size() + ch.out_facet.size();
		if (c->prm.verbose) {
			for (auto& ch : T->chunks) {
				int64_t ni = 0, no_ = 0;
				for (auto& r : ch.in) ni += r.second - r.first;
				for (auto& r : ch.out_plain) no_ += r.second - r.first;
				for (auto& r : ch.out_facet) no_ += r.second - r.first;
				fprintf(stderr, "[cslvr_b200] pipeline group: tiles %d+%d, in %zu ranges (%lld nodes), out %zu + %zu ranges (%lld nodes)\n", ch.tile0, ch.ntile,
				        ch.in.size(), (long long)ni, ch.out_plain.size(), ch.out_facet.size(), (long long)no_);
			}
		}
		if (nr > 600) T->chunks.clear();                      // a numbering without locality: too many small copies, the plain path serves it
	}

	tt.lap("pipeline groups");
	int rc = up(c, &T->d_vert, T->h_vert);
	if (!rc) rc = up(c, &T->d_node, T->h_node);
	if (!rc) rc = up(c, &T->d_tet, T->h_tet);
	if (!rc) rc = up(c, &T->d_nthr, T->h_nthr);
	if (!rc) rc = up(c, &T->d_ent0, T->h_ent0);
	if (!rc) rc = up(c, &T->d_out, T->h_out);
	if (!rc) rc = up(c, &T->d_out2, T->h_out2);
	if (!rc) rc = up(c, &T->d_tab, T->h_tab);
	if (!rc) rc = up(c, &T->d_tflag, T->h_tflag);
	if (!rc) rc = up(c, &T->d_einfo, T->h_einfo);
	if (!rc) rc = up(c, &T->d_tfacet, T->h_tfacet);
	if (!rc && fold) { rc = up(c, &T->d_fn_node, fn_node); if (!rc) rc = up(c, &T->d_fn_ptr, fn_ptr); if (!rc) rc = up(c, &T->d_fn_inc, fn_inc); }
	if (!rc && !c->host_only) {
		cudaError_t e = cudaMalloc((void**)&T->d_ainv, (size_t)n_tiles * NZ * 4 * TILE_T * sizeof(double));
		if (e != cudaSuccess) rc = bp_fail(c, BP_ERR_CUDA, "cudaMalloc(tile ainv) failed: %s", cudaGetErrorString(e));
	}
	c->tiles = T;
	if (rc) { bp_tiles_free(c); return rc; }
	if (!c->host_only) {                       // the emulator of a host-only context keeps the tables
		std::vector<int32_t>().swap(T->h_vert); std::vector<int32_t>().swap(T->h_node); std::vector<int32_t>().swap(T->h_tet);
		std::vector<int32_t>().swap(T->h_out); std::vector<int32_t>().swap(T->h_out2); std::vector<uint16_t>().swap(T->h_tab); std::vector<uint8_t>().swap(T->h_einfo);
	}
	if (c->prm.verbose)
		fprintf(stderr, "[cslvr_b200] tiled cell assembly: %d tiles, %lld prism columns x %d layers, halo factor %.3f, %lld table entries\n",
		        n_tiles, (long long)ntri, NZ, (double)T->n_threads_used / (double)ntri, (long long)ne);
	return BP_OK;
}

// ------------------------------------------------------------------------------------------------
// host emulation of k_tile: same tables, same element math, same order of additions.  Debug / CPU tests only.
static int tiles_emulate(bp_ctx* c, const double* vert_xyzS, const double* u, const double* ainv, const double* beta, int picard,
                         double* F, double* Jblocks)
{
	if (!c || !c->tiles) return bp_fail(c, BP_ERR_STATE, "no tile tables: %s", c ? c->tiles_why.c_str() : "");
	bp_tiles* T = c->tiles;
	if (T->h_vert.empty()) return bp_fail(c, BP_ERR_STATE, "tile tables were released after the upload (host-only contexts keep them)");
	AsmParams P = bp_asm_params(c);
	P.picard = picard;
	const int NL = T->NL, NZ = NL - 1;
	const double4* geo = (const double4*)vert_xyzS;
	const double2* u2 = (const double2*)u;
	std::vector<double> val((size_t)c->n_slots * 4, 0.0);
	struct HostStage {
		double* sb; uint32_t flags;
		void gate() {}
		void put(int slot, double v) { sb[(size_t)slot * TILE_T] = v; }
		void put_off(int slot, int d, double v) { sb[(size_t)(slot + d) * TILE_T] = v; }
	};
	const int W = T->W;
	#pragma omp parallel
	{
		std::vector<double> stage((size_t)TILE_NSLOT * TILE_T);
		std::vector<Iface> bot(TILE_T), top(TILE_T);
		#pragma omp for schedule(dynamic, 4)
		for (int tile = 0; tile < T->n_tiles; ++tile) {
			for (int t = 0; t < TILE_T; ++t) iface_zero(bot[t]);
			for (int k = 0; k <= NZ; ++k) {
				for (int t = 0; t < TILE_T; ++t) {           // every compute thread, padding threads (vertex 0, ainv 0) included
					HostStage st = { stage.data() + t, T->h_tflag[(size_t)tile * TILE_T + t] };
					if (k < NZ) {
						double4 pb[3], pt[3]; double2 ub[3], ut[3];
						for (int i = 0; i < 3; ++i) {
							pb[i] = geo[T->h_vert[(((size_t)tile * NL + k) * 3 + i) * TILE_T + t]];
							ub[i] = u2[T->h_node[(((size_t)tile * NL + k) * 3 + i) * TILE_T + t]];
							pt[i] = geo[T->h_vert[(((size_t)tile * NL + k + 1) * 3 + i) * TILE_T + t]];
							ut[i] = u2[T->h_node[(((size_t)tile * NL + k + 1) * 3 + i) * TILE_T + t]];
						}
						if (k == 0 && beta && T->folded) {           // the basal facet starts the bottom interface, as in the kernel
							const uint32_t ffl = T->h_tfacet[(size_t)tile * TILE_T + t];
							const int32_t* tv = T->h_vert.data() + (size_t)tile * NL * 3 * TILE_T + t;
							if (ffl) facet_fold(pb, ub, beta[tv[0]], beta[tv[TILE_T]], beta[tv[2 * TILE_T]], ffl, P, bot[t]);
						}
						double a[3];
						for (int j = 0; j < 3; ++j) { const int32_t tt = T->h_tet[(((size_t)tile * NZ + k) * 3 + j) * TILE_T + t]; a[j] = tt >= 0 ? ainv[tt] : 0.0; }
						prism_steps(pb, ub, pt, ut, a[0], a[1], a[2], P, bot[t], top[t], st);
						bot[t] = top[t];
					} else iface_stage(bot[t], st);
				}
				for (int e = T->h_ent0[tile]; e < T->h_ent0[tile + 1]; ++e) {
					const int o = T->h_out[(size_t)k * T->n_ent + e];
					if (o == -1) continue;
					const double4 acc = entry_total(stage.data(), T->h_tab.data() + (size_t)e * W, T->h_einfo[e]);
					const double a0 = acc.x, a1 = acc.y, a2 = acc.z, a3 = acc.w;
					if (o >= 0) {
						double* d = val.data() + 4 * (size_t)o; d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
						const int om = T->h_out2[(size_t)k * T->n_ent + e];
						if (om >= 0) { double* dt = val.data() + 4 * (size_t)om; dt[0] = a0; dt[1] = a2; dt[2] = a1; dt[3] = a3; }
					} else { F[2 * (size_t)(-o - 2)] = a0; F[2 * (size_t)(-o - 2) + 1] = a1; }
				}
			}
		}
	}
	for (int64_t i = 0; i < c->nn_owned; ++i)
		for (int32_t k = 0; k < c->h_rowptr[i + 1] - c->h_rowptr[i]; ++k)
			memcpy(Jblocks + 4 * ((size_t)c->h_rowptr[i] + k), val.data() + 4 * ((size_t)c->h_sell_ptr[i >> 5] + 32 * k + (i & 31)), 4 * sizeof(double));
	return BP_OK;
}

extern "C" int bp_debug_tiles_emulate(bp_ctx* c, const double* vert_xyzS, const double* u, const double* ainv, int picard,
                                      double* F, double* Jblocks)
{
	return tiles_emulate(c, vert_xyzS, u, ainv, nullptr, picard, F, Jblocks);
}

// ... with the basal facets the kernel folds into its first layer (beta: per vertex): cells + dGamma_b / dGamma_bw
extern "C" int bp_debug_tiles_emulate_facets(bp_ctx* c, const double* vert_xyzS, const double* u, const double* ainv, const double* beta,
                                             int picard, double* F, double* Jblocks)
{
	if (!beta) return BP_ERR_ARG;
	return tiles_emulate(c, vert_xyzS, u, ainv, beta, picard, F, Jblocks);
}

extern "C" int bp_debug_tiles_info(const bp_ctx* c, int64_t* n_tiles, int64_t* n_prism_cols, int64_t* n_threads_used, int64_t* n_entries, int64_t* n_contrib)
{
	if (!c || !c->tiles) return BP_ERR_STATE;
	if (n_tiles) *n_tiles = c->tiles->n_tiles;
	if (n_prism_cols) *n_prism_cols = c->tiles->n_prism_cols;
	if (n_threads_used) *n_threads_used = c->tiles->n_threads_used;
	if (n_entries) *n_entries = c->tiles->n_ent;
	if (n_contrib) *n_contrib = c->tiles->n_contrib;
	return BP_OK;
}
