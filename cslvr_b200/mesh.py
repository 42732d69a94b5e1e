"""Structured extruded tetrahedral meshes and their topology (host side, numpy).

Stands in for the dolfin objects a cslvr script builds before the hot path:
``Point``, ``BoxMesh`` (docs/get_started.rst:15-18) -- same vertex numbering, the
same six tetrahedra per hexahedron and ``mesh.order()`` (SURVEY.md Appendix A-M;
the counts 1200 cells / 2680 facets / 363 vertices of ``BoxMesh(10,10,2)`` are the
ones quoted in simulations/verification/ismip_hom_a.py:5).
"""
import numpy as np

# local vertices of local facet i (the facet opposite local vertex i)
FACET_VERTS = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]], dtype=np.int64)


class Point(object):
	def __init__(self, x=0.0, y=0.0, z=0.0):
		self._x = (float(x), float(y), float(z))

	def x(self): return self._x[0]
	def y(self): return self._x[1]
	def z(self): return self._x[2]
	def array(self): return np.array(self._x)


class Mesh(object):
	"""vertex coordinates + cell connectivity with dolfin's accessor names."""

	def __init__(self, coords, cells, shape=None):
		self._coords = np.ascontiguousarray(coords, dtype=np.float64)
		self._cells = np.ascontiguousarray(cells, dtype=np.int32)
		self.shape = shape                     # (nx, ny, nz) for BoxMesh
		self._ext = None

	def coordinates(self): return self._coords
	def cells(self): return self._cells
	def num_vertices(self): return self._coords.shape[0]
	def num_cells(self): return self._cells.shape[0]
	def num_facets(self): return (4 * self.num_cells() + len(self.exterior_facets()[0])) // 2

	def exterior_facets(self):
		"""(cell, local facet) of the facets that belong to a single cell."""
		if self._ext is None:
			c = self._cells.astype(np.int64)
			f = np.sort(c[:, FACET_VERTS].reshape(-1, 3), axis=1)
			nv = np.int64(self.num_vertices())
			order = np.lexsort((f[:, 2], f[:, 0] * nv + f[:, 1]))
			g = f[order]
			dup = np.zeros(g.shape[0], dtype=bool)
			eq = np.all(g[1:] == g[:-1], axis=1)
			dup[1:] |= eq
			dup[:-1] |= eq
			ext = np.sort(order[~dup])
			self._ext = ((ext // 4).astype(np.int32), (ext % 4).astype(np.int32))
		return self._ext


def BoxMesh(p1, p2, nx, ny, nz):
	"""dolfin ``BoxMesh(p1, p2, nx, ny, nz)``: vertex id = iz (ny+1)(nx+1) + iy (nx+1) + ix,
	six tetrahedra per hexahedron around the v0-v7 diagonal, cells vertex-sorted."""
	a = p1.array() if isinstance(p1, Point) else np.asarray(p1, dtype=np.float64)
	b = p2.array() if isinstance(p2, Point) else np.asarray(p2, dtype=np.float64)
	ax = [a[d] + np.arange(n + 1) * ((b[d] - a[d]) / n) for d, n in enumerate((nx, ny, nz))]
	coords = np.empty(((nx + 1) * (ny + 1) * (nz + 1), 3))
	coords[:, 0] = np.tile(ax[0], (ny + 1) * (nz + 1))
	coords[:, 1] = np.tile(np.repeat(ax[1], nx + 1), nz + 1)
	coords[:, 2] = np.repeat(ax[2], (nx + 1) * (ny + 1))
	sx, sy, sz = 1, nx + 1, (nx + 1) * (ny + 1)
	base = (np.arange(nz)[:, None, None] * sz + np.arange(ny)[None, :, None] * sy
	        + np.arange(nx)[None, None, :] * sx).reshape(-1, 1)
	# hexahedron corner offsets v0..v7 and the six tetrahedra of each hexahedron
	off = np.array([0, sx, sy, sx + sy, sz, sz + sx, sz + sy, sz + sx + sy])
	split = np.array([[0, 1, 3, 7], [0, 1, 7, 5], [0, 5, 7, 4],
	                  [0, 3, 2, 7], [0, 6, 4, 7], [0, 2, 6, 7]])
	cells = (base[:, :, None] + off[split][None, :, :]).reshape(-1, 4)
	cells.sort(axis=1)
	return Mesh(coords, cells, shape=(nx, ny, nz))


def facet_geometry(coords, cells, fc, fl):
	"""vertex ids, area, outward unit normal and midpoint of facets (cell fc, local facet fl)."""
	fv = cells[fc][np.arange(len(fc))[:, None], FACET_VERTS[fl]]
	p = coords[fv]
	cr = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
	ln = np.linalg.norm(cr, axis=1)
	nrm = cr / ln[:, None]
	inward = np.einsum('ij,ij->i', nrm, coords[cells[fc, fl]] - p[:, 0]) > 0
	nrm[inward] *= -1.0
	return fv, 0.5 * ln, nrm, p.mean(axis=1)
