"""CPU checks of the isosurface restatement oracle/mc_oracle.py (the stage after the hot path, Reconstruction/main.cxx:151-189):
its generated 256-case table and the surfaces it produces.  VTK is absent: parity with VTK itself is unpinned; what is
checked are the properties any correct contouring has."""
from collections import Counter

import numpy as np

from oracle import mc_oracle as mc


def test_case_table_is_complete_and_small():
    t = mc.table()
    assert len(t) == 256 and t[0] == [] and t[255] == []
    assert max(len(x) for x in t) == 5
    for case in range(256):
        used = {e for tri in t[case] for e in tri}
        crossing = {e for e in range(12) if ((case >> mc.edge_corners(e)[0]) & 1) != ((case >> mc.edge_corners(e)[1]) & 1)}
        assert used == crossing                       # every crossing edge carries a vertex of some triangle, no other edge does
        assert len(t[case]) == len(crossing) - 2 * _loops(t[case])      # fan triangulation: n - 2 triangles per loop


def _loops(tris):
    # number of connected components of the triangle set (= loops, each fanned from one vertex)
    comp = []
    for tri in tris:
        s = set(tri)
        hit = [c for c in comp if c & s]
        for c in hit:
            s |= c
            comp.remove(c)
        comp.append(s)
    return len(comp)


def test_cell_to_point_averages_the_sharing_cells():
    cells = np.arange(2 * 3 * 4, dtype=np.float64)      # Nx=4, Ny=3, Nz=2
    P = mc.cell_to_point(cells, (4, 3, 2))
    c = cells.reshape(2, 3, 4)
    assert P.shape == (3, 4, 5)
    assert P[0, 0, 0] == c[0, 0, 0] and P[2, 3, 4] == c[1, 2, 3]                 # corners: one cell
    assert P[0, 0, 1] == 0.5 * (c[0, 0, 0] + c[0, 0, 1])                          # boundary edge: two cells
    assert P[1, 1, 1] == c[0:2, 0:2, 0:2].mean()                                  # interior: eight cells


def test_sphere_surface_is_closed_oriented_and_where_it_should_be():
    N = 14
    g = np.linspace(-1.2 + 1.2 / N, 1.2 - 1.2 / N, N)
    Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
    cells = (1.0 - np.sqrt(X * X + Y * Y + Z * Z)) * 4 + 1.0
    v, tr = mc.contour(cells.reshape(-1), (N, N, N), [-1.2] * 3, [2.4 / N] * 3, np.eye(4).reshape(16), 1.0)
    edges = Counter()
    for a, b, c in tr:
        for e in ((a, b), (b, c), (c, a)):
            edges[tuple(sorted(e))] += 1
    assert set(edges.values()) == {2}                                              # watertight
    assert len(v) - len(edges) + len(tr) == 2                                      # a sphere
    r = np.linalg.norm(v.astype(np.float64), axis=1)
    assert 0.97 < r.min() and r.max() < 1.01
    p = v.astype(np.float64)
    n = np.cross(p[tr[:, 1]] - p[tr[:, 0]], p[tr[:, 2]] - p[tr[:, 0]])
    assert ((n * p[tr].mean(1)).sum(1) > 0).all()                                  # normals point from inside (>= value) outwards


def test_random_field_with_ambiguous_faces_is_watertight():
    rng = np.random.RandomState(3)
    N = 7
    cells = rng.uniform(-1, 1, size=N ** 3)
    # pad with a shell of low values so that the surface does not reach the grid boundary
    c = cells.reshape(N, N, N); c[0] = c[-1] = -5; c[:, 0] = c[:, -1] = -5; c[:, :, 0] = c[:, :, -1] = -5
    v, tr = mc.contour(c.reshape(-1), (N, N, N), [0, 0, 0], [1, 1, 1], np.eye(4).reshape(16), 0.1)
    edges = Counter()
    for a, b, cc in tr:
        for e in ((a, b), (b, cc), (cc, a)):
            edges[(int(e[0]), int(e[1]))] += 1
    # every directed edge appears once and its reverse once: closed and consistently oriented
    assert all(n == 1 for n in edges.values())
    assert all((b, a) in edges for (a, b) in edges)
