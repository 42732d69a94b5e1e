// bp_setup.cu -- one-off host-side digestion of the mesh and dof map.
//
// Replaces what dolfin does once per FunctionSpace in the reference: DofMapBuilder
// for QM2 (cslvr/model.py:438-447), the periodic merge (cslvr/d3model.py:62-108) and
// SparsityPatternBuilder (SURVEY A-S).  Output: node-blocked (2x2) BSR pattern of the
// Jacobian for the owned rows, the per-tet block slots used by the assembly kernels,
// and -- lazily -- the scalar CSR pattern in the caller's dof numbering.
#include "bp_internal.h"
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <numeric>

int bp_fail(bp_ctx* c, int code, const char* fmt, ...)
{
	static thread_local char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	extern std::string g_bp_create_error;
	if (c) c->err = buf; else g_bp_create_error = buf;
	return code;
}

template <class T>
static int upload(bp_ctx* c, T** dst, const T* src, size_t n)
{
	*dst = nullptr;
	if (c->host_only) return BP_OK;            // debug contexts without a device keep the host tables only
	if (n == 0) n = 1;
	BP_CUDA(c, cudaMalloc((void**)dst, n * sizeof(T)));
	if (src) BP_CUDA(c, cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
	return BP_OK;
}
#define UP(dst, src, n) do { int rc_ = upload(c, &(dst), (src), (size_t)(n)); if (rc_) return rc_; } while (0)

int bp_setup_topology(bp_ctx* c, const bp_mesh* m)
{
	const int64_t nv = m->n_vertices, nt = m->n_cells;
	if (nv <= 0 || nt <= 0 || !m->coords || !m->cells)
		return bp_fail(c, BP_ERR_ARG, "bp_mesh: empty mesh or NULL coords/cells");
	if (nv > INT32_MAX / 4 || nt > INT32_MAX / 32)
		return bp_fail(c, BP_ERR_ARG, "bp_mesh: mesh too large for 32-bit slots on one rank");
	c->nv = nv; c->nt = nt; c->nd_ext = m->n_dofs;
	for (int64_t i = 0; i < nt * 4; ++i)
		if (m->cells[i] < 0 || m->cells[i] >= nv)
			return bp_fail(c, BP_ERR_ARG, "bp_mesh: cells[%lld] = %d out of range", (long long)i, m->cells[i]);

	bp_stage_timer tm("setup");
	// ---- nodes --------------------------------------------------------------
	std::vector<int32_t> v2n(nv, -1);
	if (m->cell_dofs) {
		if (m->n_owned_nodes >= 0)
			return bp_fail(c, BP_ERR_ARG, "bp_mesh: a partitioned mesh requires the canonical numbering (cell_dofs == NULL)");
		if (m->n_dofs <= 0 || (m->n_dofs & 1))
			return bp_fail(c, BP_ERR_ARG, "bp_mesh: n_dofs must be positive and even with cell_dofs");
		// a node is identified by its u_x dof; nodes are numbered by ascending u_x dof
		std::vector<int32_t> dof_node(m->n_dofs, -1), ydof(m->n_dofs, -1);
		for (int64_t t = 0; t < nt; ++t)
			for (int a = 0; a < 4; ++a) {
				int32_t dx = m->cell_dofs[t * 8 + a], dy = m->cell_dofs[t * 8 + 4 + a];
				if (dx < 0 || dx >= m->n_dofs || dy < 0 || dy >= m->n_dofs || dx == dy)
					return bp_fail(c, BP_ERR_ARG, "bp_mesh: cell_dofs of cell %lld out of range", (long long)t);
				if (ydof[dx] >= 0 && ydof[dx] != dy)
					return bp_fail(c, BP_ERR_ARG, "bp_mesh: cell_dofs is not a P1xP1 mixed dof map (u_x dof %d pairs with u_y dofs %d and %d)", dx, ydof[dx], dy);
				ydof[dx] = dy;
			}
		int32_t nn = 0;
		for (int64_t d = 0; d < m->n_dofs; ++d) if (ydof[d] >= 0) dof_node[d] = nn++;
		if ((int64_t)nn * 2 != m->n_dofs)
			return bp_fail(c, BP_ERR_ARG, "bp_mesh: cell_dofs references %d nodes but n_dofs = %lld", nn, (long long)m->n_dofs);
		c->nn = nn;
		c->h_ext_dof.resize((size_t)nn * 2);
		c->identity_dofs = true;
		for (int64_t d = 0; d < m->n_dofs; ++d) if (ydof[d] >= 0) {
			int32_t i = dof_node[d];
			c->h_ext_dof[2 * i] = (int32_t)d; c->h_ext_dof[2 * i + 1] = ydof[d];
			if (d != 2 * i || ydof[d] != 2 * i + 1) c->identity_dofs = false;
		}
		for (int64_t t = 0; t < nt; ++t)
			for (int a = 0; a < 4; ++a) {
				int32_t v = m->cells[t * 4 + a], i = dof_node[m->cell_dofs[t * 8 + a]];
				if (v2n[v] >= 0 && v2n[v] != i)
					return bp_fail(c, BP_ERR_ARG, "bp_mesh: vertex %d maps to two different dofs", v);
				v2n[v] = i;
			}
	} else {
		int32_t mx = -1;
		for (int64_t v = 0; v < nv; ++v) {
			v2n[v] = m->vertex_to_node ? m->vertex_to_node[v] : (int32_t)v;
			if (v2n[v] < 0) return bp_fail(c, BP_ERR_ARG, "bp_mesh: vertex_to_node[%lld] < 0", (long long)v);
			mx = std::max(mx, v2n[v]);
		}
		c->nn = (int64_t)mx + 1;
		if (m->n_dofs != 2 * c->nn)
			return bp_fail(c, BP_ERR_ARG, "bp_mesh: n_dofs = %lld but the node map has %lld nodes", (long long)m->n_dofs, (long long)c->nn);
		c->identity_dofs = true;
		c->h_ext_dof.resize((size_t)c->nn * 2);
		std::iota(c->h_ext_dof.begin(), c->h_ext_dof.end(), 0);
	}
	c->identity_v2n = true;
	for (int64_t v = 0; v < nv; ++v) {
		if (v2n[v] < 0) v2n[v] = 0;                 // vertex not used by any cell
		if (v2n[v] != v) c->identity_v2n = false;
	}
	const int64_t nn = c->nn;
	c->nn_owned = (m->n_owned_nodes >= 0) ? m->n_owned_nodes : nn;
	if (c->nn_owned > nn) return bp_fail(c, BP_ERR_ARG, "bp_mesh: n_owned_nodes > number of nodes");
	const int64_t no = c->nn_owned;

	tm.lap("nodes");
	// ---- node -> incident cells ------------------------------------------------
	std::vector<int32_t> cn((size_t)nt * 4);
	for (int64_t i = 0; i < nt * 4; ++i) cn[i] = v2n[m->cells[i]];
	std::vector<int64_t> nptr(nn + 1, 0);
	for (int64_t i = 0; i < nt * 4; ++i) nptr[cn[i] + 1]++;
	for (int64_t i = 0; i < nn; ++i) nptr[i + 1] += nptr[i];
	std::vector<int32_t> ncell(nptr[nn]);
	{
		std::vector<int64_t> fill(nptr.begin(), nptr.end() - 1);
		for (int64_t t = 0; t < nt; ++t)
			for (int a = 0; a < 4; ++a) ncell[fill[cn[t * 4 + a]]++] = (int32_t)t;
	}
	tm.lap("node->cells");
	// ---- BSR pattern of the owned rows: row i = sorted unique nodes sharing a cell ---
	c->h_rowptr.assign(no + 1, 0);
	#pragma omp parallel
	{
		std::vector<int32_t> tmp;
		#pragma omp for schedule(static)
		for (int64_t i = 0; i < no; ++i) {
			tmp.clear();
			for (int64_t k = nptr[i]; k < nptr[i + 1]; ++k)
				for (int a = 0; a < 4; ++a) tmp.push_back(cn[(int64_t)ncell[k] * 4 + a]);
			std::sort(tmp.begin(), tmp.end());
			c->h_rowptr[i + 1] = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
		}
	}
	for (int64_t i = 0; i < no; ++i) {
		int64_t s = (int64_t)c->h_rowptr[i] + c->h_rowptr[i + 1];
		if (s > INT32_MAX / 4) return bp_fail(c, BP_ERR_ARG, "bp_mesh: too many Jacobian blocks for one rank");
		c->h_rowptr[i + 1] = (int32_t)s;
	}
	c->nnzb = c->h_rowptr[no];
	c->h_col.resize(c->nnzb);
	std::vector<int32_t> diag(no, -1);
	#pragma omp parallel
	{
		std::vector<int32_t> tmp;
		#pragma omp for schedule(static)
		for (int64_t i = 0; i < no; ++i) {
			tmp.clear();
			for (int64_t k = nptr[i]; k < nptr[i + 1]; ++k)
				for (int a = 0; a < 4; ++a) tmp.push_back(cn[(int64_t)ncell[k] * 4 + a]);
			std::sort(tmp.begin(), tmp.end());
			size_t nu = std::unique(tmp.begin(), tmp.end()) - tmp.begin();
			int32_t* dst = c->h_col.data() + c->h_rowptr[i];
			for (size_t k = 0; k < nu; ++k) { dst[k] = tmp[k]; if (tmp[k] == i) diag[i] = c->h_rowptr[i] + (int32_t)k; }
		}
	}
	// rows of owned nodes that touch no cell get an empty row; give them a diagonal guard
	for (int64_t i = 0; i < no; ++i)
		if (diag[i] < 0 && nptr[i + 1] > nptr[i]) return bp_fail(c, BP_ERR_ARG, "internal: missing diagonal block");

	tm.lap("BSR pattern");
	// ---- SELL-32 slots of the resident Jacobian: block k of row i lives at
	// sell_ptr[i/32] + 32 k + (i%32); padding slots keep value 0 and point at the row itself
	const int64_t nsl_j = (no + 31) / 32;
	std::vector<int32_t> sell_ptr(nsl_j + 1, 0), row_len(no, 0);
	for (int64_t sidx = 0; sidx < nsl_j; ++sidx) {
		int32_t w = 0;
		for (int64_t i = sidx * 32; i < std::min<int64_t>(no, sidx * 32 + 32); ++i) {
			row_len[i] = c->h_rowptr[i + 1] - c->h_rowptr[i];
			w = std::max(w, row_len[i]);
		}
		const int64_t nxt = (int64_t)sell_ptr[sidx] + 32 * (int64_t)w;
		if (nxt > INT32_MAX / 4) return bp_fail(c, BP_ERR_ARG, "bp_mesh: too many Jacobian blocks for one rank");
		sell_ptr[sidx + 1] = (int32_t)nxt;
	}
	c->n_slots = sell_ptr[nsl_j];
	auto slot_of = [&](int64_t i, int32_t k) -> int32_t { return sell_ptr[i >> 5] + 32 * k + (int32_t)(i & 31); };
	std::vector<int32_t> sell_col((size_t)std::max<int64_t>(c->n_slots, 1), 0);
	for (int64_t i = 0; i < no; ++i) {
		const int32_t w = (sell_ptr[(i >> 5) + 1] - sell_ptr[i >> 5]) >> 5;
		for (int32_t k = 0; k < w; ++k)
			sell_col[slot_of(i, k)] = (k < row_len[i]) ? c->h_col[c->h_rowptr[i] + k] : (int32_t)i;
		if (diag[i] >= 0) diag[i] = slot_of(i, diag[i] - c->h_rowptr[i]);
	}
	for (int64_t i = no; i < nsl_j * 32; ++i) {          // lanes past the last row
		const int32_t w = (sell_ptr[(i >> 5) + 1] - sell_ptr[i >> 5]) >> 5;
		for (int32_t k = 0; k < w; ++k) sell_col[slot_of(i, k)] = 0;
	}

	tm.lap("SELL slots");
	// ---- per-tet block slots [16][nt] ------------------------------------------
	std::vector<int32_t> tet_blk((size_t)nt * 16);
	#pragma omp parallel for schedule(static)
	for (int64_t t = 0; t < nt; ++t)
		for (int a = 0; a < 4; ++a) {
			int32_t r = cn[t * 4 + a];
			for (int b = 0; b < 4; ++b) {
				int32_t slot = -1;
				if (r < no) {
					const int32_t* lo = c->h_col.data() + c->h_rowptr[r];
					const int32_t* hi = c->h_col.data() + c->h_rowptr[r + 1];
					slot = slot_of(r, (int32_t)(std::lower_bound(lo, hi, cn[t * 4 + b]) - lo));
				}
				tet_blk[(size_t)(a * 4 + b) * nt + t] = slot;
			}
		}

	tm.lap("tet block slots");
	// ---- row-gather incidences: for every owned node the (tet, local vertex) pairs that
	// touch it, SELL-32 layout (slice of 32 rows, entry e of lane l at base + 32 e + l),
	// with the block positions of the tet's four nodes inside the node's block row.
	{
		const int64_t nsl = (no + 31) / 32;
		std::vector<int32_t> cnt(no, 0);
		for (int64_t i = 0; i < nt * 4; ++i) if (cn[i] < no) cnt[cn[i]]++;
		std::vector<int32_t> sl_ptr(nsl + 1, 0);
		c->max_row = 0;
		for (int64_t s = 0; s < nsl; ++s) {
			int32_t w = 0;
			for (int64_t i = s * 32; i < std::min<int64_t>(no, s * 32 + 32); ++i) {
				w = std::max(w, cnt[i]);
				c->max_row = std::max(c->max_row, c->h_rowptr[i + 1] - c->h_rowptr[i]);
			}
			int64_t nxt = (int64_t)sl_ptr[s] + 32 * (int64_t)w;
			if (nxt > INT32_MAX) return bp_fail(c, BP_ERR_ARG, "bp_mesh: too many incidences for one rank");
			sl_ptr[s + 1] = (int32_t)nxt;
		}
		std::vector<int32_t> code(sl_ptr[nsl], -1);
		std::vector<uint32_t> pos(sl_ptr[nsl], 0u);
		const size_t nslot = (size_t)sl_ptr[nsl];
		std::vector<int32_t> incv(4 * nslot, 0);
		// The Jacobian is symmetric: the thread of row i computes only the blocks (i, j) with j > i (and
		// the diagonal) and writes their transposes into the rows j (d_mirror).  So that the lanes of a
		// warp agree on how many blocks a visit computes, the other three vertices of an incidence are
		// ordered "upper first" (column node > i) and the incidences of a row by descending number of
		// upper vertices.
		std::vector<int32_t> start(4 * (size_t)no, 0);          // [class 3 | 2 | 1 | 0] running offsets per row
		auto n_upper_of = [&](int64_t t, int a, int32_t i) {
			int nu = 0;
			for (int j = 1; j < 4; ++j) if (cn[t * 4 + ((a + j) & 3)] > i) ++nu;
			return nu;
		};
		for (int64_t t = 0; t < nt; ++t)
			for (int a = 0; a < 4; ++a) {
				const int32_t i = cn[t * 4 + a];
				if (i < no) start[4 * (size_t)i + (3 - n_upper_of(t, a, i))]++;
			}
		for (int64_t i = 0; i < no; ++i) {
			int32_t run = 0;
			for (int q = 0; q < 4; ++q) { const int32_t cq = start[4 * i + q]; start[4 * i + q] = run; run += cq; }
		}
		for (int64_t t = 0; t < nt; ++t)
			for (int a = 0; a < 4; ++a) {
				const int32_t i = cn[t * 4 + a];
				if (i >= no) continue;
				const int nu = n_upper_of(t, a, i);
				const int64_t at = (int64_t)sl_ptr[i / 32] + 32 * (int64_t)(start[4 * (size_t)i + (3 - nu)]++) + (i % 32);
				code[at] = (int32_t)(t * 4 + a);
				const int32_t* lo = c->h_col.data() + c->h_rowptr[i];
				const int32_t* hi = c->h_col.data() + c->h_rowptr[i + 1];
				int ord[4] = { a, 0, 0, 0 }, nu_seen = 0, nl_seen = 0;
				for (int j = 1; j < 4; ++j) {                    // upper vertices first, the others after them
					const int lv = (a + j) & 3;
					if (cn[t * 4 + lv] > i) ord[1 + nu_seen++] = lv; else ord[1 + nu + nl_seen++] = lv;
				}
				uint32_t pk = 0;
				for (int j = 0; j < 4; ++j) {
					const uint32_t k = (uint32_t)(std::lower_bound(lo, hi, cn[t * 4 + ord[j]]) - lo);
					if (k > 255) return bp_fail(c, BP_ERR_ARG, "bp_mesh: a Jacobian row has more than 255 node blocks");
					pk |= k << (8 * j);
					incv[(size_t)j * nslot + at] = m->cells[t * 4 + ord[j]];
				}
				pos[at] = pk;
			}
		{	// slot of the transposed block: (row j, column i) for every block (i, j) with j > i owned
			std::vector<int32_t> mirror((size_t)std::max<int64_t>(c->n_slots, 1), -1);
			#pragma omp parallel for schedule(static)
			for (int64_t i = 0; i < no; ++i)
				for (int32_t k = c->h_rowptr[i]; k < c->h_rowptr[i + 1]; ++k) {
					const int32_t j = c->h_col[k];
					if (j <= i || j >= no) continue;
					const int32_t* lo = c->h_col.data() + c->h_rowptr[j];
					const int32_t* hi = c->h_col.data() + c->h_rowptr[j + 1];
					const int32_t kj = (int32_t)(std::lower_bound(lo, hi, (int32_t)i) - lo);
					mirror[(size_t)sell_ptr[i >> 5] + 32 * (size_t)(k - c->h_rowptr[i]) + (i & 31)] = sell_ptr[j >> 5] + 32 * kj + (int32_t)(j & 31);
				}
			UP(c->d_mirror, mirror.data(), mirror.size());
		}
		{	// what the pipelined host-buffer bp_assemble needs per CTA of the row-gather kernel
			const int64_t nblk = (no + BP_ROWS_PER_CTA - 1) / BP_ROWS_PER_CTA;
			std::vector<int32_t> mxv(nblk, -1);
			for (int64_t t = 0; t < nt; ++t)
				for (int a = 0; a < 4; ++a) {
					const int32_t i = cn[t * 4 + a];
					if (i >= no) continue;
					int32_t& mv = mxv[i / BP_ROWS_PER_CTA];
					for (int j = 0; j < 4; ++j) mv = std::max(mv, m->cells[t * 4 + j]);
				}
			std::vector<int32_t> pmx(nv);                       // prefix maximum of the node of vertices 0..v
			for (int64_t v = 0; v < nv; ++v) pmx[v] = std::max(v2n[v], v ? pmx[v - 1] : -1);
			c->h_blk_need_vert.assign(nblk, 0); c->h_blk_need_node.assign(nblk, 0);
			int32_t run = -1;
			for (int64_t b = 0; b < nblk; ++b) {
				run = std::max(run, mxv[b]);
				c->h_blk_need_vert[b] = run + 1;
				c->h_blk_need_node[b] = run >= 0 ? pmx[run] + 1 : 0;
			}
		}
		UP(c->d_inc_slice, sl_ptr.data(), nsl + 1);
		UP(c->d_inc_code, code.data(), code.size());
		UP(c->d_inc_pos, pos.data(), pos.size());
		UP(c->d_inc_v, incv.data(), incv.size());
		c->n_inc_slots = (int64_t)nslot;
	}

	tm.lap("row-gather incidences");
	// ---- facets that integrate something ----------------------------------------
	// dGamma_b = {3,5}, dGamma_bw = {5}, dGamma_lt = {4,10}: cslvr/d3model.py:576-598
	std::vector<int32_t> fcell, finfo;
	const bool lateral = (!c->prm.periodic) && c->prm.use_pressure_bc;    // momentumbp.py:488
	c->n_basal = 0;
	for (int64_t f = 0; f < m->n_facets; ++f) {
		int32_t mk = m->facet_marker[f], cell = m->facet_cell[f], lf = m->facet_local[f];
		if (cell < 0 || cell >= nt || lf < 0 || lf > 3)
			return bp_fail(c, BP_ERR_ARG, "bp_mesh: facet %lld has cell %d / local facet %d", (long long)f, cell, lf);
		int basal = (mk == 3 || mk == 5), press = (mk == 5) || (lateral && (mk == 4 || mk == 10));
		if (!basal && !press) continue;
		c->n_basal += basal;
		fcell.push_back(cell);
		finfo.push_back(lf | (basal << 8) | (press << 9) | ((mk == 3) << 10));   // bit 10: dGamma_bg (grounded bed), the reduced-Stokes sliding set
	}
	c->nf = (int64_t)fcell.size();
	c->h_fcell = fcell; c->h_finfo = finfo;
	{	// facet row gather: every owned node touched by an integrating facet, with its (facet, local vertex) pairs
		static const int LOCG[4][3] = { { 1, 2, 3 }, { 0, 2, 3 }, { 0, 1, 3 }, { 0, 1, 2 } };
		std::vector<int32_t> cntn(no, 0);
		for (size_t f = 0; f < fcell.size(); ++f)
			for (int k = 0; k < 3; ++k) { const int32_t nd_ = cn[(int64_t)fcell[f] * 4 + LOCG[finfo[f] & 0xff][k]]; if (nd_ < no) cntn[nd_]++; }
		std::vector<int32_t> fn_node, fn_ptr(1, 0), slot_of_node(no, -1);
		for (int64_t i = 0; i < no; ++i) if (cntn[i]) { slot_of_node[i] = (int32_t)fn_node.size(); fn_node.push_back((int32_t)i); fn_ptr.push_back(fn_ptr.back() + cntn[i]); }
		std::vector<int32_t> fn_inc((size_t)std::max<int32_t>(fn_ptr.back(), 1), 0), fillf(fn_node.size(), 0);
		for (size_t f = 0; f < fcell.size(); ++f)
			for (int k = 0; k < 3; ++k) {
				const int32_t nd_ = cn[(int64_t)fcell[f] * 4 + LOCG[finfo[f] & 0xff][k]];
				if (nd_ >= no) continue;
				const int32_t r = slot_of_node[nd_];
				fn_inc[fn_ptr[r] + fillf[r]++] = (int32_t)(f * 4 + k);
			}
		c->n_fnodes = (int64_t)fn_node.size();
		c->h_facet_node.assign((size_t)std::max<int64_t>(no, 1), 0);
		for (int32_t nd_ : fn_node) c->h_facet_node[nd_] = 1;
		UP(c->d_fn_node, fn_node.data(), fn_node.size());
		UP(c->d_fn_ptr, fn_ptr.data(), fn_ptr.size());
		UP(c->d_fn_inc, fn_inc.data(), fn_inc.size());
	}
	{	// CTAs of the row-gather kernel with a row that a facet integral adds to
		static const int LOCF[4][3] = { { 1, 2, 3 }, { 0, 2, 3 }, { 0, 1, 3 }, { 0, 1, 2 } };
		c->h_blk_facet.assign((size_t)((no + BP_ROWS_PER_CTA - 1) / BP_ROWS_PER_CTA), 0);
		for (size_t f = 0; f < fcell.size(); ++f)
			for (int k = 0; k < 3; ++k) { const int32_t nd_ = cn[(int64_t)fcell[f] * 4 + LOCF[finfo[f] & 0xff][k]]; if (nd_ < no) c->h_blk_facet[nd_ / BP_ROWS_PER_CTA] = 1; }
	}
	{	// Dirichlet dofs of the u_z system: nodes of the basal facets (cslvr/momentumbp.py:79-85)
		static const int LOC[4][3] = { { 1, 2, 3 }, { 0, 2, 3 }, { 0, 1, 3 }, { 0, 1, 2 } };
		std::vector<uint8_t> bc((size_t)std::max<int64_t>(no, 1), 0);
		for (size_t f = 0; f < fcell.size(); ++f) if ((finfo[f] >> 8) & 1)
			for (int k = 0; k < 3; ++k) { const int32_t nd_ = cn[(int64_t)fcell[f] * 4 + LOC[finfo[f] & 0xff][k]]; if (nd_ < no) bc[nd_] = 1; }
		c->h_bcnode = bc;
		UP(c->d_bcnode, bc.data(), bc.size());
	}

	tm.lap("facets");
	// ---- ice columns (vertical-line preconditioner): owned nodes with the same (x, y),
	// ordered bottom to top; slots of the diagonal and of the up / down coupling blocks
	{
		std::vector<int32_t> rep(nn, -1);                       // lowest vertex id of every node
		for (int64_t v = nv - 1; v >= 0; --v) rep[v2n[v]] = (int32_t)v;
		double lo[2] = { 1e300, 1e300 }, hi[2] = { -1e300, -1e300 };
		for (int64_t v = 0; v < nv; ++v) for (int d = 0; d < 2; ++d) { lo[d] = std::min(lo[d], m->coords[3 * v + d]); hi[d] = std::max(hi[d], m->coords[3 * v + d]); }
		const double sx_ = (hi[0] > lo[0]) ? 1048576.0 / (hi[0] - lo[0]) : 0.0, sy_ = (hi[1] > lo[1]) ? 1048576.0 / (hi[1] - lo[1]) : 0.0;
		std::vector<std::pair<int64_t, std::pair<double, int32_t>>> key(no);
		for (int64_t i = 0; i < no; ++i) {
			const int32_t v = rep[i] < 0 ? 0 : rep[i];
			const int64_t kx = (int64_t)std::llround((m->coords[3 * v] - lo[0]) * sx_), ky = (int64_t)std::llround((m->coords[3 * v + 1] - lo[1]) * sy_);
			key[i] = { (kx << 22) | ky, { m->coords[3 * v + 2], (int32_t)i } };
		}
		std::sort(key.begin(), key.end());
		std::vector<int32_t> colptr, colnode(no), tri_d(no, -1), tri_u(no, -1), tri_l(no, -1);
		c->max_layers = 0;
		for (int64_t j = 0; j < no; ++j) {
			if (j == 0 || key[j].first != key[j - 1].first) colptr.push_back((int32_t)j);
			colnode[j] = key[j].second.second;
		}
		colptr.push_back((int32_t)no);
		c->n_cols = (int64_t)colptr.size() - 1;
		auto find_slot = [&](int32_t r, int32_t cnode) -> int32_t {
			const int32_t* lo_ = c->h_col.data() + c->h_rowptr[r];
			const int32_t* hi_ = c->h_col.data() + c->h_rowptr[r + 1];
			const int32_t* it = std::lower_bound(lo_, hi_, cnode);
			return (it != hi_ && *it == cnode) ? slot_of(r, (int32_t)(it - lo_)) : -1;
		};
		for (int64_t cc = 0; cc < c->n_cols; ++cc) {
			c->max_layers = std::max(c->max_layers, colptr[cc + 1] - colptr[cc]);
			for (int32_t j = colptr[cc]; j < colptr[cc + 1]; ++j) {
				tri_d[j] = diag[colnode[j]];
				if (j + 1 < colptr[cc + 1]) tri_u[j] = find_slot(colnode[j], colnode[j + 1]);
				if (j > colptr[cc]) tri_l[j] = find_slot(colnode[j], colnode[j - 1]);
			}
		}
		c->h_colptr = colptr; c->h_colnode = colnode;
		{	// largest number of ice columns one tet touches: lmax(T^-1 J) <= that number for the
			// column-block splitting (J is a sum of positive semi-definite element matrices)
			std::vector<int32_t> col_of(nn, -1);
			for (int64_t cc = 0; cc < c->n_cols; ++cc) for (int32_t j = colptr[cc]; j < colptr[cc + 1]; ++j) col_of[colnode[j]] = (int32_t)cc;
			int mx = 1;
			for (int64_t t = 0; t < nt; ++t) {
				int32_t cs[4]; int k = 0;
				for (int a = 0; a < 4; ++a) { const int32_t q = col_of[cn[t * 4 + a]]; bool seen = false; for (int b = 0; b < k; ++b) seen |= (cs[b] == q); if (!seen) cs[k++] = q; }
				mx = std::max(mx, k);
			}
			c->col_touch_max = mx;
			// mesh anisotropy: median over (a sample of) the tets of horizontal extent / vertical extent
			std::vector<double> asp;
			const int64_t stride = std::max<int64_t>(1, nt / 20000);
			for (int64_t t = 0; t < nt; t += stride) {
				double lo3[3] = { 1e300, 1e300, 1e300 }, hi3[3] = { -1e300, -1e300, -1e300 };
				for (int a = 0; a < 4; ++a) for (int d2 = 0; d2 < 3; ++d2) {
					const double v = m->coords[3 * (int64_t)m->cells[t * 4 + a] + d2];
					lo3[d2] = std::min(lo3[d2], v); hi3[d2] = std::max(hi3[d2], v);
				}
				asp.push_back(std::max(hi3[0] - lo3[0], hi3[1] - lo3[1]) / std::max(hi3[2] - lo3[2], 1e-300));
			}
			std::nth_element(asp.begin(), asp.begin() + asp.size() / 2, asp.end());
			c->aspect = asp[asp.size() / 2];
		}
		UP(c->d_colptr, colptr.data(), colptr.size());
		UP(c->d_colnode, colnode.data(), no);
		UP(c->d_tri_d, tri_d.data(), no);
		UP(c->d_tri_u, tri_u.data(), no);
		UP(c->d_tri_l, tri_l.data(), no);
	}

	tm.lap("ice columns");
	// ---- halo -----------------------------------------------------------------
	c->n_nbr = (m->n_owned_nodes >= 0) ? m->n_neighbors : 0;
	if (c->n_nbr > 0) {
		c->nbr_rank.assign(m->neighbor_rank, m->neighbor_rank + c->n_nbr);
		c->send_ptr.assign(m->send_ptr, m->send_ptr + c->n_nbr + 1);
		c->recv_ptr.assign(m->recv_ptr, m->recv_ptr + c->n_nbr + 1);
		c->n_send = c->send_ptr[c->n_nbr];
		if (c->recv_ptr[c->n_nbr] != nn - no)
			return bp_fail(c, BP_ERR_ARG, "bp_mesh: recv_ptr covers %lld ghost nodes, mesh has %lld", (long long)c->recv_ptr[c->n_nbr], (long long)(nn - no));
		for (int64_t k = 0; k < c->n_send; ++k)
			if (m->send_nodes[k] < 0 || m->send_nodes[k] >= no)
				return bp_fail(c, BP_ERR_ARG, "bp_mesh: send_nodes[%lld] is not an owned node", (long long)k);
		c->h_send_nodes.assign(m->send_nodes, m->send_nodes + c->n_send);
		UP(c->d_send_nodes, m->send_nodes, c->n_send);
		UP(c->d_sendbuf, (const double*)nullptr, 2 * c->n_send);
	} else if (nn != no) {
		return bp_fail(c, BP_ERR_ARG, "bp_mesh: ghost nodes without neighbours");
	}

	tm.lap("halo");
	// ---- upload ---------------------------------------------------------------
	{
		std::vector<double4> geo(nv);
		for (int64_t v = 0; v < nv; ++v)
			geo[v] = make_double4(m->coords[3 * v], m->coords[3 * v + 1], m->coords[3 * v + 2], 0.0);
		UP(c->d_geo, geo.data(), nv);
		std::vector<double> soa((size_t)4 * nv, 0.0);        // x | y | z | S, for the row-gather kernel
		for (int64_t v = 0; v < nv; ++v) for (int d = 0; d < 3; ++d) soa[(size_t)d * nv + v] = m->coords[3 * v + d];
		UP(c->d_soa, soa.data(), (size_t)4 * nv);
		std::vector<double2> zs(nv);
		for (int64_t v = 0; v < nv; ++v) zs[v] = make_double2(m->coords[3 * v + 2], 0.0);
		UP(c->d_zS, zs.data(), nv);
	}
	UP(c->d_A, (const double*)nullptr, nv);
	UP(c->d_beta, (const double*)nullptr, nv);
	UP(c->d_Tp, (const double*)nullptr, nv);
	UP(c->d_W, (const double*)nullptr, nv);
	if (!c->host_only) BP_CUDA(c, cudaMemset(c->d_W, 0, (size_t)nv * sizeof(double)));
	{ std::vector<double> ones(nv, 1.0); UP(c->d_E, ones.data(), nv); }
	UP(c->d_ainv, (const double*)nullptr, nt);
	UP(c->d_Bv, (const double*)nullptr, nv);
	UP(c->d_Bring, (const double*)nullptr, nv);
	if (!c->host_only) BP_CUDA(c, cudaMemset(c->d_Bring, 0, (size_t)nv * sizeof(double)));
	UP(c->d_uz, (const double*)nullptr, nv);
	if (!c->host_only) BP_CUDA(c, cudaMemset(c->d_uz, 0, (size_t)nv * sizeof(double)));
	UP(c->d_uvert2, (const double*)nullptr, 2 * nv);
	UP(c->d_tets, (const int4*)m->cells, nt);
	if (!c->identity_v2n) c->h_v2n = v2n;
	if (!c->identity_v2n) { UP(c->d_v2n, v2n.data(), nv); UP(c->d_uvert, (const double*)nullptr, (size_t)2 * nv); }
	UP(c->d_tet_blk, tet_blk.data(), (size_t)nt * 16);
	UP(c->d_fac_cell, fcell.data(), c->nf);
	UP(c->d_fac_info, finfo.data(), c->nf);
	if (!c->identity_dofs) UP(c->d_ext_dof, c->h_ext_dof.data(), 2 * nn);
	UP(c->d_sell_ptr, sell_ptr.data(), nsl_j + 1);
	UP(c->d_rowlen, row_len.data(), no);
	UP(c->d_col, sell_col.data(), c->n_slots);
	c->h_sell_ptr = sell_ptr;
	UP(c->d_diag, diag.data(), no);
	UP(c->d_val, (const double*)nullptr, (size_t)c->n_slots * 4);
	double** vecs[] = { &c->d_u, &c->d_F, &c->d_dx, &c->d_r, &c->d_z, &c->d_p, &c->d_Ap, &c->d_io, &c->d_io2, &c->d_z2, &c->d_cd, &c->d_aux };
	for (double** v : vecs) {
		size_t len = std::max<size_t>((size_t)2 * nn, (size_t)std::max<int64_t>(std::max<int64_t>(c->nd_ext, nv), 1));
		UP(*v, (const double*)nullptr, len);
		if (!c->host_only) BP_CUDA(c, cudaMemset(*v, 0, len * sizeof(double)));
	}
	UP(c->d_dinv, (const double*)nullptr, (size_t)no * 4);
	UP(c->d_sc, (const double*)nullptr, SC_COUNT);
	if (!c->host_only) BP_CUDA(c, cudaMemset(c->d_sc, 0, SC_COUNT * sizeof(double)));
	UP(c->d_part, (const double*)nullptr, 4 * 4096);
	UP(c->d_counter, (const unsigned*)nullptr, 16);
	if (!c->host_only) {
		BP_CUDA(c, cudaMemset(c->d_counter, 0, 16 * sizeof(unsigned)));
		BP_CUDA(c, cudaMallocHost((void**)&c->h_sc, SC_COUNT * sizeof(double)));
		BP_CUDA(c, cudaMemset(c->d_val, 0, (size_t)std::max<int64_t>(c->n_slots, 1) * 4 * sizeof(double)));
	}
	// tiled cell assembly for extruded prism meshes (bp_tiles.cu); other meshes are served by the row gather
	tm.lap("upload");
	{ const int rc_t = bp_setup_tiles(c, m, v2n); tm.lap("tiles"); return rc_t; }
}

// Scalar CSR in the caller's numbering: row = caller's dof, columns ascending,
// explicit zeros kept (dolfin SparsityPatternBuilder / PETSc AIJ, SURVEY A-S).
int bp_build_csr_export(bp_ctx* c)
{
	if (c->have_csr) return BP_OK;
	if (c->nn_owned != c->nn)
		return bp_fail(c, BP_ERR_STATE, "CSR export is defined for an unpartitioned context only");
	const int64_t nn = c->nn, nd = c->nd_ext;
	// internal row (node, comp) of every caller dof
	std::vector<int32_t> row_of(nd, -1);
	for (int64_t i = 0; i < 2 * nn; ++i) row_of[c->h_ext_dof[i]] = (int32_t)i;
	c->csr_indptr.assign(nd + 1, 0);
	for (int64_t d = 0; d < nd; ++d) {
		int32_t r = row_of[d];
		int64_t len = (r < 0) ? 0 : 2 * (int64_t)(c->h_rowptr[r / 2 + 1] - c->h_rowptr[r / 2]);
		c->csr_indptr[d + 1] = c->csr_indptr[d] + len;
	}
	const int64_t nnz = c->csr_indptr[nd];
	c->csr_indices.resize(nnz);
	std::vector<int32_t> src(nnz);
	#pragma omp parallel
	{
		std::vector<std::pair<int32_t, int32_t>> tmp;
		#pragma omp for schedule(static)
		for (int64_t d = 0; d < nd; ++d) {
			int32_t r = row_of[d];
			if (r < 0) continue;
			int32_t node = r / 2, comp = r % 2;
			tmp.clear();
			for (int32_t k = c->h_rowptr[node]; k < c->h_rowptr[node + 1]; ++k)
				for (int cc = 0; cc < 2; ++cc)
					tmp.push_back({ c->h_ext_dof[2 * c->h_col[k] + cc], (c->h_sell_ptr[node >> 5] + 32 * (k - c->h_rowptr[node]) + (node & 31)) * 4 + comp * 2 + cc });
			std::sort(tmp.begin(), tmp.end());
			int64_t o = c->csr_indptr[d];
			for (size_t k = 0; k < tmp.size(); ++k) { c->csr_indices[o + k] = tmp[k].first; src[o + k] = tmp[k].second; }
		}
	}
	int rc = upload(c, &c->d_csr_src, src.data(), (size_t)nnz);
	if (rc) return rc;
	c->have_csr = true;
	return BP_OK;
}
