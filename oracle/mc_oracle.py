"""oracle/mc_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy + plain Python loops, small grids only) of the stage that FOLLOWS the integration in the
reference's pipeline (Reconstruction/main.cxx:151-189):

    vtkCellDataToPointData   cell scalars "reconstruction_scalar" -> point scalars            (:151-154)
    vtkContourFilter         isosurface of the point scalars at --contour (default 1.0)       (:169-173)
    vtkTransformFilter       grid matrix applied to the surface's points                      (:177-181)

All three are VTK classes (un-vendored, version unpinned, absent from this image): PARITY UNPINNED for this stage.
What is restated is their published behaviour, and where VTK leaves freedom the choice is stated here:
  * cell -> point: a point's value is the average of the cells that share it, sum_j (1/n) * v_j accumulated in double
    in vtkStructuredData::GetPointCells' cell order (offsets (-1,0,0) (-1,-1,0) (-1,-1,-1) (-1,0,-1) (0,0,0) (0,-1,0)
    (0,-1,-1) (0,0,-1) to the point's index, cells outside the grid skipped).
  * a corner is INSIDE when its scalar >= value (VTK's contouring convention); a grid edge whose two ends differ
    carries exactly one surface vertex at t = (value - s0) / (s1 - s0), x = origin + (index + t * axis) * spacing,
    evaluated in double and stored as float32 (vtkPoints' default); the vertex SET of an isosurface is the same for
    marching cubes and for VTK's synchronized templates -- the triangulation is not, and is not claimed to match.
  * the grid matrix is applied to the float32 point in double (vtkLinearTransform on float points) and the result is
    stored as float32.
  * triangles: per cell, the crossing edges are joined face by face (a face with four crossings -- diagonal corners
    inside -- is cut so that each segment isolates one INSIDE corner; the rule depends on the face's own corners only,
    so neighbouring cells agree and the surface is watertight), the resulting closed loops are fan-triangulated from
    their first vertex (lowest edge id first, walking towards its lower-numbered face neighbour... see
    `case_triangles`) and oriented so that normals point from inside (>= value) to outside.
Vertices are numbered by owning point (k, j, i order), then axis (x, y, z); triangles by cell (k, j, i order).
"""
from __future__ import annotations

import numpy as np

# cube corners: bit 0 = x, bit 1 = y, bit 2 = z.  Edge id = 4 * axis + (a + 2 * b), (a, b) = the other two coordinates
# in increasing axis order; the edge runs from corner `lo` to corner `lo | (1 << axis)`.
AXES = [(0, 1, 2), (1, 0, 2), (2, 0, 1)]          # (axis, first other, second other)


def edge_corners(e):
    axis, o1, o2 = AXES[e // 4]
    a, b = (e % 4) & 1, (e % 4) >> 1
    lo = (a << o1) | (b << o2)
    return lo, lo | (1 << axis)


# faces: (fixed axis, side); each lists its 4 edges
def face_edges(axis, side):
    out = []
    for e in range(12):
        lo, hi = edge_corners(e)
        if ((lo >> axis) & 1) == side and ((hi >> axis) & 1) == side:
            out.append(e)
    return out


FACES = [(ax, sd) for ax in range(3) for sd in range(2)]


def corner_pos(c):
    return np.array([c & 1, (c >> 1) & 1, (c >> 2) & 1], dtype=np.float64)


def edge_mid(e):
    lo, hi = edge_corners(e)
    return 0.5 * (corner_pos(lo) + corner_pos(hi))


def case_triangles(case):
    """Triangles (as triples of edge ids) of one cube configuration; bit c of `case` = corner c is inside."""
    inside = [(case >> c) & 1 for c in range(8)]
    crossing = [e for e in range(12) if inside[edge_corners(e)[0]] != inside[edge_corners(e)[1]]]
    if not crossing:
        return []
    nbr = {e: [] for e in crossing}
    for axis, side in FACES:
        es = [e for e in face_edges(axis, side) if e in nbr]
        if len(es) == 2:
            nbr[es[0]].append(es[1]); nbr[es[1]].append(es[0])
        elif len(es) == 4:
            # ambiguous face: pair the two crossing edges that meet at each INSIDE corner of the face
            for c in range(8):
                if ((c >> axis) & 1) == side and inside[c]:
                    pair = [e for e in es if c in edge_corners(e)]
                    assert len(pair) == 2
                    nbr[pair[0]].append(pair[1]); nbr[pair[1]].append(pair[0])
    assert all(len(v) == 2 for v in nbr.values())
    tris, seen = [], set()
    for start in crossing:                                   # lowest edge id first
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            a, b = nbr[cur]
            nxt = min(a, b) if prev is None else (a if b == prev else b)     # first step: towards the lower edge id
            if prev is not None and a == b:
                nxt = a
            if nxt == start:
                break
            loop.append(nxt); seen.add(nxt)
            prev, cur = cur, nxt
        # orientation: normals from inside to outside
        pts = [edge_mid(e) for e in loop]
        n = np.zeros(3)
        for q in range(len(pts)):
            n += np.cross(pts[q], pts[(q + 1) % len(pts)])
        s = 0.0
        for e in loop:
            lo, hi = edge_corners(e)
            d = corner_pos(hi) - corner_pos(lo)
            s += float(np.dot(n, d if inside[lo] else -d))
        assert s != 0.0
        if s < 0:
            loop = [loop[0]] + loop[:0:-1]
        for q in range(1, len(loop) - 1):
            tris.append((loop[0], loop[q], loop[q + 1]))
    return tris


_TABLE = None


def table():
    global _TABLE
    if _TABLE is None:
        _TABLE = [case_triangles(c) for c in range(256)]
    return _TABLE


def cell_to_point(cells, n_cells):
    """cells: flat array in VTK cell order ((k*Ny + j)*Nx + i) -> point scalars [(Nz+1)][(Ny+1)][(Nx+1)] (double)."""
    Nx, Ny, Nz = n_cells
    c = np.asarray(cells, dtype=np.float64).reshape(Nz, Ny, Nx)
    offs = [(-1, 0, 0), (-1, -1, 0), (-1, -1, -1), (-1, 0, -1), (0, 0, 0), (0, -1, 0), (0, -1, -1), (0, 0, -1)]
    P = np.zeros((Nz + 1, Ny + 1, Nx + 1))
    cnt = np.zeros((Nz + 1, Ny + 1, Nx + 1), dtype=np.int64)
    ii, jj, kk = np.meshgrid(np.arange(Nx + 1), np.arange(Ny + 1), np.arange(Nz + 1), indexing="ij")
    for di, dj, dk in offs:
        ci, cj, ck = ii + di, jj + dj, kk + dk
        ok = (ci >= 0) & (ci < Nx) & (cj >= 0) & (cj < Ny) & (ck >= 0) & (ck < Nz)
        cnt[kk[ok], jj[ok], ii[ok]] += 1
    w = 1.0 / cnt
    for di, dj, dk in offs:                                   # accumulation in GetPointCells' order
        ci, cj, ck = ii + di, jj + dj, kk + dk
        ok = (ci >= 0) & (ci < Nx) & (cj >= 0) & (cj < Ny) & (ck >= 0) & (ck < Nz)
        P[kk[ok], jj[ok], ii[ok]] += w[kk[ok], jj[ok], ii[ok]] * c[ck[ok], cj[ok], ci[ok]]
    return P


def contour(cells, n_cells, origin, spacing, grid_matrix, value):
    """Returns (vertices float32 [nV,3] in world coordinates, triangles int32 [nT,3])."""
    Nx, Ny, Nz = n_cells
    P = cell_to_point(cells, n_cells)
    inside = P >= value
    origin = np.asarray(origin, dtype=np.float64); spacing = np.asarray(spacing, dtype=np.float64)
    M = np.asarray(grid_matrix, dtype=np.float64).reshape(4, 4)
    verts = []
    vid = {}
    for k in range(Nz + 1):
        for j in range(Ny + 1):
            for i in range(Nx + 1):
                for axis, (di, dj, dk) in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
                    i1, j1, k1 = i + di, j + dj, k + dk
                    if i1 > Nx or j1 > Ny or k1 > Nz:
                        continue
                    if inside[k, j, i] == inside[k1, j1, i1]:
                        continue
                    s0, s1 = P[k, j, i], P[k1, j1, i1]
                    t = (value - s0) / (s1 - s0)
                    idx = np.array([i, j, k], dtype=np.float64)
                    idx[axis] = idx[axis] + t
                    x = (origin + idx * spacing).astype(np.float32).astype(np.float64)       # vtkPoints: float32
                    wv = [np.float32(M[r, 0] * x[0] + M[r, 1] * x[1] + M[r, 2] * x[2] + M[r, 3]) for r in range(3)]
                    vid[(i, j, k, axis)] = len(verts)
                    verts.append(wv)
    tab = table()
    tris = []
    for k in range(Nz):
        for j in range(Ny):
            for i in range(Nx):
                case = 0
                for c in range(8):
                    if inside[k + ((c >> 2) & 1), j + ((c >> 1) & 1), i + (c & 1)]:
                        case |= 1 << c
                for tri in tab[case]:
                    out = []
                    for e in tri:
                        lo, _ = edge_corners(e)
                        out.append(vid[(i + (lo & 1), j + ((lo >> 1) & 1), k + ((lo >> 2) & 1), e // 4)])
                    tris.append(out)
    return (np.array(verts, dtype=np.float32).reshape(-1, 3), np.array(tris, dtype=np.int32).reshape(-1, 3))
