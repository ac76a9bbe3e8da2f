"""Mesh generator (csrc/mesh.cpp) against the reference's closed-form counts
(src/PolyMesh2d.f90:717-766) and structural invariants its unit tests check
(tests/IcosTriMeshTester.f90:72-89)."""
import numpy as np
import pytest

from lpm_v2_b200 import mesh as M


@pytest.mark.parametrize("L", [0, 1, 2, 3, 4, 5])
def test_icos_tri_counts_and_area(get_mesh, L):
    m = get_mesh(M.ICOS_TRI_SPHERE_SEED, L)
    F, V = 20 * 4 ** L, 2 + 10 * 4 ** L
    assert m.n == F + V
    assert m.n_active == F and m.n_leaf_faces == F
    assert m.n_leaf_edges == 30 * 4 ** L                     # leaf-edge count, IcosTriMeshTester.f90:83-89
    assert m.n_faces_total == sum(20 * 4 ** i for i in range(L + 1))
    assert abs(m.area.sum() - 4 * np.pi) < 1e-12            # Particles.f90:690-694
    assert np.all(m.area[m.is_active == 0] == 0.0)
    r = np.sqrt(m.x ** 2 + m.y ** 2 + m.z ** 2)
    assert np.max(np.abs(r - 1.0)) < 1e-14
    # the active particles are exactly the leaf-face centres, each once
    assert sorted(m.face_center.tolist()) == np.nonzero(m.is_active)[0].tolist()


def test_icos_tri_order(get_mesh):
    """Particle insertion order: 12 seed vertices, 20 seed centres, then per
    divided face its new edge midpoints followed by 3 new centres."""
    m0 = get_mesh(M.ICOS_TRI_SPHERE_SEED, 0)
    assert m0.is_active.tolist() == [0] * 12 + [1] * 20
    assert m0.z[0] == 1.0 and m0.z[11] == -1.0
    m1 = get_mesh(M.ICOS_TRI_SPHERE_SEED, 1)
    # refinement never moves existing particles
    assert np.array_equal(m1.x[:32], m0.x) and np.array_equal(m1.z[:32], m0.z)
    # face 1 divides first: 3 midpoints (passive) then 3 centres (active)
    assert m1.is_active[32:38].tolist() == [0, 0, 0, 1, 1, 1]
    # face 2 shares one edge with face 1: only 2 new midpoints
    assert m1.is_active[38:43].tolist() == [0, 0, 1, 1, 1]
    # level-0 centres stay active (child 4 keeps the parent's centre, Faces.f90:847-853)
    assert m1.is_active[12:32].all()
    m3 = get_mesh(M.ICOS_TRI_SPHERE_SEED, 3)
    m2 = get_mesh(M.ICOS_TRI_SPHERE_SEED, 2)
    assert np.array_equal(m3.y[:m2.n], m2.y)


@pytest.mark.parametrize("L", [0, 1, 2, 3, 4])
def test_cubed_sphere(get_mesh, L):
    m = get_mesh(M.CUBED_SPHERE_SEED, L)
    F = 6 * 4 ** L
    assert m.n_active == F
    assert m.n == F + 2 + 6 * 4 ** L        # PolyMesh2d.f90:739-740 (demoted parent centres are vertices)
    assert abs(m.area.sum() - 4 * np.pi) < 1e-12
    # divided faces' centres become passive vertices with zero area (Faces.f90:684-685)
    if L > 0:
        assert m.is_active[8:14].sum() == 0 and np.all(m.area[8:14] == 0)


@pytest.mark.parametrize("L", [0, 1, 2, 3, 4])
def test_quad_rect(get_mesh, L):
    m = get_mesh(M.QUAD_RECT_SEED, L, 7.0)
    nv = (3 + sum(2 ** i for i in range(1, L + 1))) ** 2     # PolyMesh2d.f90:731-736
    assert m.n == nv + 4 * 4 ** L
    assert m.n_active == 4 * 4 ** L
    assert abs(m.area.sum() - 196.0) < 1e-10
    assert m.x.min() == -7.0 and m.x.max() == 7.0
    assert np.all(m.z == 0.0)


def test_beta_plane_and_tri_hex(get_mesh):
    b = get_mesh(M.BETA_PLANE_SEED, 3)
    assert (b.x.min(), b.x.max(), b.y.min(), b.y.max()) == (0.0, 1.0, -0.5, 0.5)   # PolyMesh2d.f90:880-884
    assert abs(b.area.sum() - 1.0) < 1e-13
    h = get_mesh(M.TRI_HEX_SEED, 3)
    assert h.n_active == 6 * 4 ** 3
    assert abs(h.area.sum() - 1.5 * np.sqrt(3.0)) < 1e-13


def test_max_edge_length(get_mesh):
    """MaxEdgeLength (Edges.f90:260-275) feeds the PSE eps; halves per level."""
    h = [get_mesh(M.ICOS_TRI_SPHERE_SEED, L).max_edge_length for L in range(5)]
    assert abs(h[0] - np.arctan(2.0)) < 1e-14          # icosahedron edge angle
    for a, b in zip(h, h[1:]):
        assert 1.7 < a / b < 2.05


def _parse_vtk(path):
    toks = open(path).read().split("\n")
    assert toks[0] == "# vtk DataFile Version 2.0" and toks[2] == "ASCII" and toks[3] == "DATASET POLYDATA"
    out = {"title": toks[1], "fields": {}}
    i = 4
    while i < len(toks):
        ln = toks[i].split()
        if not ln:
            i += 1
            continue
        if ln[0] == "POINTS":
            n = int(ln[1])
            out["points"] = np.array([[float(v) for v in toks[i + 1 + k].split()] for k in range(n)])
            i += 1 + n
        elif ln[0] == "POLYGONS":
            nc, size = int(ln[1]), int(ln[2])
            out["polys"] = np.array([[int(v) for v in toks[i + 1 + k].split()] for k in range(nc)])
            assert size == 4 * nc
            i += 1 + nc
        elif ln[0] in ("POINT_DATA", "CELL_DATA"):
            section, count = ln[0], int(ln[1])
            i += 1
        elif ln[0] == "SCALARS":
            name, nd = ln[1], int(ln[3])
            assert toks[i + 1] == "LOOKUP_TABLE default"
            vals = np.array([[float(v) for v in toks[i + 2 + k].split()] for k in range(count)])
            assert vals.shape == (count, nd)
            out["fields"][(section, name)] = vals
            i += 2 + count
        else:
            raise AssertionError(f"unexpected line {i}: {toks[i]!r}")
    return out


@pytest.mark.parametrize("seed,L", [(M.ICOS_TRI_SPHERE_SEED, 2), (M.CUBED_SPHERE_SEED, 1), (M.QUAD_RECT_SEED, 2)])
def test_write_vtk_layout(tmp_path, seed, L):
    """OutputToVTK (src/SphereBVE.f90:283-328): POINTS, POLYGONS (leaf faces as triangles around their
    centre particle, 0-based), POINT_DATA lagParam + fields, CELL_DATA faceArea; exact round trip."""
    m = M.PolyMesh2d(seed, L, 2.0 if seed == M.QUAD_RECT_SEED else 1.0)
    sphere = seed != M.QUAD_RECT_SEED
    zeta = np.cos(3 * m.x) + m.y
    zeta[5] = 3e-15                                    # below ZERO_TOL: written as 0 (Field.f90:299-303)
    vel = (-m.y, m.x, 0.1 * m.z) if sphere else (-m.y, m.x)
    moved = (m.x + 0.01, m.y - 0.02, m.z)
    path = tmp_path / "mesh.vtk"
    m.write_vtk(path, [("relVort_1/s", zeta), ("velocity_m/s", vel)], title="unit test", positions=moved)
    v = _parse_vtk(path)
    assert v["title"] == "unit test"
    pts = np.stack([moved[0], moved[1], moved[2] if sphere else 0 * m.x], 1)
    assert np.array_equal(v["points"], pts)
    vpf = m.face_verts.shape[1]
    assert v["polys"].shape == (vpf * m.n_leaf_faces, 4) and np.all(v["polys"][:, 0] == 3)
    for f in range(m.n_leaf_faces):
        for j in range(vpf):
            row = v["polys"][f * vpf + j]
            assert row[1] == m.face_verts[f, j] and row[2] == m.face_verts[f, (j + 1) % vpf] and row[3] == m.face_center[f]
    lag = np.stack([m.x, m.y, m.z if sphere else 0 * m.x], 1)
    assert np.array_equal(v["fields"][("POINT_DATA", "lagParam")], lag)
    z0 = zeta.copy()
    z0[5] = 0.0
    assert np.array_equal(v["fields"][("POINT_DATA", "relVort_1/s")][:, 0], z0)
    assert np.array_equal(v["fields"][("POINT_DATA", "velocity_m/s")], np.stack(vel, 1))
    area = v["fields"][("CELL_DATA", "faceArea")][:, 0]
    assert np.array_equal(area, np.repeat(m.area[m.face_center], vpf))
    # the sub-triangles tile the mesh: their flat areas approach the panel areas
    if not sphere:
        p = v["points"]
        tri = v["polys"][:, 1:]
        a = 0.5 * np.abs((p[tri[:, 1], 0] - p[tri[:, 0], 0]) * (p[tri[:, 2], 1] - p[tri[:, 0], 1]) -
                         (p[tri[:, 2], 0] - p[tri[:, 0], 0]) * (p[tri[:, 1], 1] - p[tri[:, 0], 1]))
        assert abs(a.sum() - m.area[m.is_active != 0].sum()) <= 1e-12 * a.sum()
