"""CPU known-answer tests of the oracle's building blocks, each against a hand-derivable value or an independent
numpy evaluation of the reference formula (file:line in the oracle source)."""
import numpy as np


def test_dims_truncation_quirk(oracle):
    # OG.hpp:614-625: res passes through float, dims = (int)((max-min)/double(float(res)))
    assert oracle.OracleGrid((-0.25, 0.25) * 3, 0.001).dims == (499, 499, 499)
    assert oracle.OracleGrid((-0.25, 0.25) * 3, 0.005).dims == (100, 100, 100)     # 0.005f > 0.005 would give 99; it is below
    assert oracle.OracleGrid((-0.8, 1.8, -1.5, 1.5, 0.0, 1.0), 0.005).dims == (520, 600, 200)   # launch:8 box, SURVEY 8
    d = 0.5 / np.float64(np.float32(0.001))
    assert int(d) == 499 and d > 499.9999


def test_voxel_index_and_strict_box(oracle):
    g = oracle.OracleGrid((-0.25, 0.25) * 3, 0.001)
    res = np.float64(np.float32(0.001))
    pts = np.array([[-0.25, 0, 0],            # on xmin: rejected (x<=xmin)
                    [0.25, 0, 0],             # on xmax: rejected (x>=xmax)
                    [np.nextafter(np.float32(-0.25), np.float32(0)), 0, 0],   # first float inside
                    [0.2499, 0.2499, 0.2499],  # lands in the pad cell index == dim (OG.hpp:626)
                    [0.0, 0.0, 0.0],
                    [np.nan, 0, 0]], np.float32)
    ijk, valid = oracle.kat_voxel(g, pts)
    assert list(valid) == [0, 0, 1, 1, 1, 0]
    assert tuple(ijk[2]) == (0, 249, 249)
    assert tuple(ijk[3]) == (499, 499, 499) and g.dims[0] == 499
    want = np.floor((pts[4].astype(np.float64) + 0.25) / res).astype(int)
    assert tuple(ijk[4]) == tuple(want) == (249, 249, 249)
    # centre = float(min + res*i + res/2) in double (OG.hpp:131-135)
    c = oracle.kat_center(g, np.array([[0, 10, 499]], np.int32))[0]
    assert np.array_equal(c, np.float32(-0.25 + res * np.array([0, 10, 499]) + res / 2.0))


def test_transform_is_double_then_narrowed(oracle):
    rng = np.random.default_rng(3)
    T = np.eye(4); T[:3, :3] = np.linalg.qr(rng.normal(size=(3, 3)))[0]; T[:3, 3] = rng.normal(size=3)
    p = rng.uniform(-1, 1, (1000, 3)).astype(np.float32)
    got = oracle.kat_transform(T, p)
    pd = p.astype(np.float64)
    want = np.stack([((T[r, 0] * pd[:, 0] + T[r, 1] * pd[:, 1]) + T[r, 2] * pd[:, 2]) + T[r, 3] for r in range(3)], 1).astype(np.float32)
    assert np.array_equal(got, want)
    assert not np.array_equal(got, (p @ T[:3, :3].astype(np.float32).T + T[:3, 3].astype(np.float32)))   # float math differs


def test_projection_and_distance(oracle):
    # OG.hpp:40-49: projection onto the line through `c` along `n`
    c = np.array([[0.1, -0.05, 0.2]], np.float32)
    n = np.array([[0.0, 0.0, 1.0]], np.float32)
    pt = np.array([[0.1005, -0.05, 0.2031]], np.float32)
    proj, dist = oracle.kat_project(pt, c, n)
    assert abs(proj[0, 0] - 0.1) < 1e-7 and abs(proj[0, 1] + 0.05) < 1e-7 and abs(proj[0, 2] - pt[0, 2]) < 2e-7
    assert abs(dist[0] - 0.0005) < 1e-7


def test_pca_normal_of_planes(oracle):
    res = np.float64(np.float32(0.005))
    i, j = np.meshgrid(np.arange(5), np.arange(5), indexing="ij")
    for axis in range(3):
        idx = np.zeros((25, 3)); other = [a for a in range(3) if a != axis]
        idx[:, other[0]], idx[:, other[1]], idx[:, axis] = i.ravel(), j.ravel(), 2
        pts = (-0.25 + res * (idx + 40) + res / 2).astype(np.float32)
        cov, nrm, ev = oracle.kat_normal(pts)
        assert abs(abs(nrm[axis]) - 1) < 1e-3 and abs(ev) < 1e-6, (axis, nrm, ev)
    # tilted noisy plane: within a few degrees of the true normal, unit length
    rng = np.random.default_rng(5)
    true = np.array([0.3, -0.2, 0.93]); true /= np.linalg.norm(true)
    xy = rng.uniform(-0.01, 0.01, (60, 2))
    z = -(true[0] * xy[:, 0] + true[1] * xy[:, 1]) / true[2] + rng.normal(0, 2e-4, 60)
    pts = (np.c_[xy, z] + [0.1, 0.1, 0.1]).astype(np.float32)
    _, nrm, _ = oracle.kat_normal(pts)
    assert abs(np.linalg.norm(nrm) - 1) < 1e-5
    assert abs(np.dot(nrm, true)) > 0.99


def test_eigen33_against_numpy(oracle):
    rng = np.random.default_rng(6)
    for _ in range(200):
        a = rng.normal(size=(3, 3)); cov = (a @ a.T).astype(np.float32) * 1e-5
        nrm, ev = oracle.kat_eigen33(cov)
        w, v = np.linalg.eigh(cov.astype(np.float64))
        if (w[1] - w[0]) / w[2] < 1e-2:
            continue
        assert abs(abs(np.dot(nrm, v[:, 0])) - 1) < 1e-2
        assert abs(ev - w[0]) <= 2e-3 * w[2]


def test_cylinder_score_welford(oracle):
    # points on the axis +- offsets: accepted iff perpendicular distance < 1 mm (OG.hpp:262,426)
    c, n = np.array([0, 0, 0], np.float32), np.array([0, 0, 1], np.float32)
    pts = np.array([[0.0005, 0, 0.001], [0.002, 0, 0.0], [0, 0.0009, -0.002], [0.0011, 0, 0], [0, 0, 0.003]], np.float32)
    cen, sd, md, sdd, cnt = oracle.kat_score(pts, c, n)
    acc = pts[[0, 2, 4]].astype(np.float64)
    assert cnt == 3
    assert np.allclose(cen, [0, 0, acc[:, 2].mean()], atol=1e-7)
    d = np.array([0.0005, 0.0009, 0.0])
    assert abs(md - d.mean()) < 1e-7
    assert abs(sdd - d.var()) < 1e-9          # the recurrence yields the population variance
    assert np.allclose(sd, [0, 0, acc[:, 2].var()], atol=1e-9)


def test_update_gate_and_orientation(oracle):
    """A flat 1-voxel sheet: interior voxels see 25 neighbours (>20) and get a normal; normals face the viewpoint."""
    box, res = (-0.1, 0.1) * 3, 0.005
    g = oracle.OracleGrid(box, res)
    r = np.float64(np.float32(res))
    ii, jj = np.meshgrid(np.arange(8, 20), np.arange(8, 20), indexing="ij")
    pts = np.stack([-0.1 + r * (ii.ravel() + 0.5), -0.1 + r * (jj.ravel() + 0.5), np.full(ii.size, -0.1 + r * 20.5)], 1).astype(np.float32)
    g.add_points_world(pts, np.array([0, 0, -1.0], np.float32))       # viewpoint below the sheet
    g.update()
    out = g.download()
    # 12x12 sheet: voxels with a full 5x5 neighbourhood = 8x8; edge voxels with >=21 neighbours do not exist (5x4=20)
    assert len(out) == 64
    assert np.all(out.normal[:, 2] < -0.999)                           # flipped towards the viewpoint (OG.hpp:393-396)
    assert np.all(out.count >= 1)                                      # each voxel's own centre point is inside its cylinder
    s = g.state()
    assert s.normal_found.sum() == 64 and len(s.hash) == 144 and np.all(s.buffer_len == 1)
