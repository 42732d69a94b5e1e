"""Round-2 planning aid for the compute-once cell kernel: group the ice columns of a
mesh into tiles of C columns along a Morton curve of the footprint and count, per tile, the tets that touch
its rows (what a CTA would have to compute) against its share of the mesh -- the halo factor -- and the
shared memory a z-ring of two layers of staged element blocks (44 doubles per tet) would need.  CPU only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..', 'tests'))
os.environ.setdefault('CSLVR_B200_QUIET', '1')


def morton(ix, iy):
	def spread(v):
		v = v.astype(np.uint64) & np.uint64(0xffff)
		v = (v | (v << np.uint64(8))) & np.uint64(0x00ff00ff)
		v = (v | (v << np.uint64(4))) & np.uint64(0x0f0f0f0f)
		v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
		v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
		return v
	return spread(ix) | (spread(iy) << np.uint64(1))


def stats(model, C):
	X = model.flat_coordinates[model.node_vertex]
	cells = model.vertex_to_node[model.mesh.cells()]
	# columns = nodes with equal (x, y)
	key = np.round(X[:, :2] / (np.ptp(X[:, :2], axis=0).max() * 1e-9)).astype(np.int64)
	_, col_of = np.unique(key, axis=0, return_inverse=True)
	col_of = col_of.ravel()
	ncol = col_of.max() + 1
	cx = np.zeros(ncol); cy = np.zeros(ncol)
	cx[col_of] = X[:, 0]; cy[col_of] = X[:, 1]
	q = 1023.0 / max(np.ptp(cx), np.ptp(cy))
	order = np.argsort(morton(((cx - cx.min()) * q).astype(np.int64), ((cy - cy.min()) * q).astype(np.int64)), kind='stable')
	tile_of_col = np.empty(ncol, dtype=np.int64)
	tile_of_col[order] = np.arange(ncol) // C
	ntile = tile_of_col.max() + 1
	tt = tile_of_col[col_of[cells]]                       # [nt, 4] tile of every vertex's column
	tt.sort(axis=1)
	first = np.ones_like(tt, dtype=bool)
	first[:, 1:] = tt[:, 1:] != tt[:, :-1]
	touched = np.bincount(tt[first], minlength=ntile)    # tets touching each tile
	nt = cells.shape[0]
	nlay = X.shape[0] // ncol - 1
	halo = touched.sum() / nt
	per_layer = touched.max() / max(nlay, 1)
	return halo, touched.max(), per_layer, 2 * per_layer * 44 * 8 / 1024.0


def main():
	from cases import ismip_model, delaunay_extruded_model
	for name, m in (('BoxMesh 64x64x5 periodic', ismip_model('C', L=20000.0, n=(64, 64, 5))),
	                ('Delaunay footprint, 400 points x 5 layers', delaunay_extruded_model(n_pts=400, n_layers=5))):
		print(name, '-', m.num_cells, 'tets')
		for C in (8, 16, 32, 64, 128):
			h, tmax, pl, kb = stats(m, C)
			print('  %3d columns per tile: halo factor %.2f, at most %5d tets per tile (%4.0f per layer), z-ring of 2 layers %.0f KB' % (C, h, tmax, pl, kb))


if __name__ == '__main__':
	main()
