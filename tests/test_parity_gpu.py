"""GPU parity tests: libpcfusion.so (through its C ABI) against the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from helpers import RESULT_FIELDS, STATE_FIELDS, assert_result_parity, assert_same, bits_equal, run_schedule

pytestmark = pytest.mark.gpu


def _synth(pcf):
    import importlib
    return importlib.import_module(pcf.__name__ + ".synth")


@pytest.fixture(scope="module")
def small(pcf):
    return _synth(pcf).small_sphere(6)


def test_kat_transform_clip_box_voxel(pcf, oracle):
    """a2/a3/a6: z clip, FP64 transform narrowed to float, strict box test, floor((p-min)/res)."""
    rng = np.random.default_rng(0)
    box, res = (-0.25, 0.25) * 3, 0.001
    fus = pcf.Fusion(box, res)
    og = oracle.OracleGrid(box, res)
    n = 200000
    pts = np.zeros((n, 4), np.float32)
    pts[:, :3] = rng.uniform(-0.3, 0.3, (n, 3))
    pts[:, 2] = rng.uniform(0.25, 0.65, n)
    # adversarial: exact clip limits, NaNs, points on voxel borders
    pts[:50, 2] = np.float32(0.28); pts[50:100, 2] = np.float32(0.6); pts[100:150, 2] = np.nan
    pts[150:160, 0] = np.nan
    ang = 0.3
    T = np.eye(4); T[:3, :3] = [[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]]; T[:3, 3] = [0.01, -0.02, -0.45]
    w, ijk, kept = fus.kat_transform_voxel(pts, T)
    clip = (pts[:, 2].astype(np.float64) < 0.6) & (pts[:, 2].astype(np.float64) > 0.28)
    w_ref = oracle.kat_transform(T, pts)
    ijk_ref, valid_ref = oracle.kat_voxel(og, w_ref)
    keep_ref = clip & (valid_ref == 1)
    assert np.array_equal(kept.astype(bool), keep_ref)
    assert bits_equal(w[keep_ref], w_ref[keep_ref])
    assert np.array_equal(ijk[keep_ref], ijk_ref[keep_ref])
    # identity pose with world points exactly on cell borders / box faces
    k = rng.integers(0, 500, (20000, 3))
    border = (-0.25 + k * np.float64(np.float32(res))).astype(np.float32)
    border[:, 2] = np.clip(border[:, 2], 0.2801, 0.5999)  # must survive the clip (pose is identity)
    og2 = oracle.OracleGrid((-0.25, 0.25, -0.25, 0.25, 0.0, 1.0), res)
    fus2 = pcf.Fusion((-0.25, 0.25, -0.25, 0.25, 0.0, 1.0), res)
    p4 = np.zeros((len(border), 4), np.float32); p4[:, :3] = border
    w2, ijk2, kept2 = fus2.kat_transform_voxel(p4, np.eye(4))
    ijk2_ref, valid2_ref = oracle.kat_voxel(og2, border)
    assert np.array_equal(kept2, valid2_ref)
    assert np.array_equal(ijk2[valid2_ref == 1], ijk2_ref[valid2_ref == 1])
    fus.close(); fus2.close()


def test_kat_pca_normal(pcf, oracle):
    """a10: float covariance in neighbour order + eigen33, both root branches."""
    rng = np.random.default_rng(1)
    fus = pcf.Fusion((-0.25, 0.25) * 3, 0.005)
    res = np.float64(np.float32(0.001))
    for trial in range(300):
        base = rng.integers(10, 400, 3)
        if trial % 3 == 0:      # perfect axis-aligned plane -> computeRoots2 branch
            occ = np.zeros((5, 5, 5), bool); occ[:, :, 2] = True
        elif trial % 3 == 1:    # noisy tilted sheet -> trigonometric branch
            i, j = np.meshgrid(np.arange(5), np.arange(5), indexing="ij")
            k = np.clip(np.round(2 + 0.4 * (i - 2) - 0.3 * (j - 2) + rng.normal(0, 0.5, (5, 5))), 0, 4).astype(int)
            occ = np.zeros((5, 5, 5), bool); occ[i, j, k] = True
            occ |= rng.random((5, 5, 5)) < 0.1
        else:
            occ = rng.random((5, 5, 5)) < 0.4
        idx = np.argwhere(occ)   # ascending (i, j, k) = the reference's d order
        if len(idx) < 21:
            continue
        centres = (-0.25 + res * (base + idx) + res / 2.0).astype(np.float32)
        _, want, _ = oracle.kat_normal(centres)
        got = fus.kat_normal(centres)
        assert bits_equal(got, want), (trial, got, want)
    fus.close()


def test_kat_cylinder_score(pcf, oracle):
    """a8/a11: projection, 1 mm test and the Welford recurrences, sequential order."""
    rng = np.random.default_rng(2)
    fus = pcf.Fusion((-0.25, 0.25) * 3, 0.005)
    for trial in range(50):
        c = rng.uniform(-0.2, 0.2, 3).astype(np.float32)
        n = rng.normal(size=3); n = (n / np.linalg.norm(n)).astype(np.float32)
        t = rng.uniform(-0.004, 0.004, 400)[:, None]
        pts = (c + t * n + rng.normal(0, 0.0007, (400, 3))).astype(np.float32)
        want = oracle.kat_score(pts, c, n)
        got = fus.kat_score(pts, c, n)
        assert got[4] == want[4] and got[4] > 0
        for g, w in zip(got[:4], want[:4]):
            assert bits_equal(np.asarray(g, np.float32), np.asarray(w, np.float32)), (trial, got, want)
    fus.close()


@pytest.mark.parametrize("update_every", [None, 1, 2])
def test_replay_parity_small(pcf, oracle, small, update_every):
    """End to end: frames -> update schedule -> extraction; grid state and output bit-exact."""
    g = small.grid
    og = oracle.OracleGrid(g.box, g.res)
    fus = pcf.Fusion(g.box, g.res)
    run_schedule(og, small, update_every)
    run_schedule(fus, small, update_every)
    assert fus.dims == og.dims
    assert_same(fus.state(), og.state(), STATE_FIELDS, "state.")
    want, got = og.download(), fus.extract()
    assert len(want) > 1000
    assert_result_parity(got, want, "result.")
    fus.close()


def test_device_batch_matches_host_frames(pcf, small):
    """pcf_push_frames_device (one launch for the whole batch) == per-frame host pushes."""
    import torch
    g = small.grid
    frames = [small.frame(i) for i in range(small.n_frames)]
    a = pcf.Fusion(g.box, g.res)
    for i, (pts, T) in enumerate(frames):
        a.push_frame(pts, T, i)
    a.update()
    b = pcf.Fusion(g.box, g.res)
    dev = torch.from_numpy(np.stack([f[0] for f in frames])).cuda()
    b.push_frames_device(dev, len(frames), frames[0][0].shape[0], 4, np.stack([f[1] for f in frames]), 0)
    b.update()
    assert_same(b.extract(), a.extract(), RESULT_FIELDS)
    assert_same(b.state(), a.state(), STATE_FIELDS)
    a.close(); b.close()


def test_stride3_and_clear_and_restart(pcf, oracle, small):
    g = small.grid
    og = oracle.OracleGrid(g.box, g.res)
    fus = pcf.Fusion(g.box, g.res, started=False)
    pts, T = small.frame(0)
    assert fus.push_frame(pts, T, 0) == 1          # dropped while stopped (node.cpp:329)
    fus.start()
    for rep in range(2):                           # second round after clear must equal the first
        for i in range(3):
            pts, T = small.frame(i)
            fus.push_frame(np.ascontiguousarray(pts[:, :3]), T, i)   # packed xyz
            if rep == 0:
                og.add_frame(pts, T)
        fus.update()
        if rep == 0:
            og.update()
            want = og.download()
        assert_result_parity(fus.extract(), want)
        fus.clear()
        assert len(fus.extract()) == 0
    fus.close()


@pytest.mark.parametrize("world", [2, 3])
def test_frame_sharded_merge_is_byte_identical(pcf, small, world):
    """N ranks emulated as N contexts on one GPU; merged extraction == 1-rank extraction."""
    import importlib
    sh = importlib.import_module(pcf.__name__ + ".sharded")
    g = small.grid
    frames = [small.frame(i) for i in range(small.n_frames)]
    one = pcf.Fusion(g.box, g.res)
    for i, (pts, T) in enumerate(frames):
        one.push_frame(pts, T, i)
    one.update()
    want = one.extract()
    ranks = [pcf.Fusion(g.box, g.res) for _ in range(world)]
    for r, f in enumerate(ranks):
        lo, hi = sh.frame_block(len(frames), r, world)
        for i in range(lo, hi):
            f.push_frame(frames[i][0], frames[i][1], i)
    got = sh.merge_and_extract_local(ranks)
    assert_same(got, want, RESULT_FIELDS, f"sharded x{world}: ")
    for f in ranks + [one]:
        f.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_exchange_v2_slab_routed_records_byte_identical(pcf, small, world):
    """Exchange v2 (pcfusion.h): records routed to the owner of their x-slab by ONE compaction+scatter kernel per rank
    (here into the other contexts' receive buffers on the same GPU); merged extraction == 1-rank extraction."""
    import importlib
    sh = importlib.import_module(pcf.__name__ + ".sharded")
    g = small.grid
    frames = [small.frame(i) for i in range(small.n_frames)]
    one = pcf.Fusion(g.box, g.res)
    for i, (pts, T) in enumerate(frames):
        one.push_frame(pts, T, i)
    one.update()
    want = one.extract()
    ranks = [pcf.Fusion(g.box, g.res) for _ in range(world)]
    for r, f in enumerate(ranks):
        lo, hi = sh.frame_block(len(frames), r, world)    # world 8 > 6 frames: some ranks own no frame at all
        for i in range(lo, hi):
            f.push_frame(frames[i][0], frames[i][1], i)
    got = sh.merge_and_extract_local_v2(ranks)
    assert_same(got, want, RESULT_FIELDS, f"exchange v2 x{world}: ")
    for f in ranks + [one]:
        f.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_device_resident_exchange_byte_identical_and_same_bounds_as_host_rule(pcf, small, world):
    """The device-resident exchange (pcf_exchange_hist / _plan / _scatter_async + region-keeping install): slab bounds computed
    by k_slab_bounds equal the host rule (sharded.choose_slabs on the summed histogram), a second round after clear() reuses
    every buffer, and the merged extraction is byte-identical to one context fed every frame."""
    import importlib
    sh = importlib.import_module(pcf.__name__ + ".sharded")
    g = small.grid
    frames = [small.frame(i) for i in range(small.n_frames)]
    one = pcf.Fusion(g.box, g.res)
    for i, (pts, T) in enumerate(frames):
        one.push_frame(pts, T, i)
    one.update()
    want = one.extract()
    ranks = [pcf.Fusion(g.box, g.res) for _ in range(world)]
    for rnd in range(2):
        for r, f in enumerate(ranks):
            lo, hi = sh.frame_block(len(frames), r, world)
            for i in range(lo, hi):
                f.push_frame(frames[i][0], frames[i][1], i)
        plane = sum(f.plane_point_counts().astype(np.int64) for f in ranks)
        got, bounds = sh.merge_and_extract_local_v3(ranks)
        assert bounds == sh.slab_bounds_from_points(plane, world)
        assert_same(got, want, RESULT_FIELDS, f"device-resident exchange x{world} round {rnd}: ")
        for f in ranks:
            f.clear()
    for f in ranks + [one]:
        f.close()


def test_pointcloud2_front_end_and_add_points(pcf, oracle, small):
    """a1 (node.cpp:182-216): a RealSense-style PointCloud2 (x y z at bytes 0/4/8, rgb at 16, point_step 20, padded
    rows) gives the same grid as the float4 path; pcf_add_points (OG.hpp:185 verbatim) equals the oracle's world-frame
    insertion with an explicit viewpoint."""
    g = small.grid
    a, b = pcf.Fusion(g.box, g.res), pcf.Fusion(g.box, g.res)
    W, H = small.width, small.height
    for i in range(3):
        pts, T = small.frame(i)
        a.push_frame(pts, T, i)
        step, row = 20, W * 20 + (0 if i == 0 else 12)               # frame 0: dense rows, later frames: padded rows
        msg = np.zeros(H * row, np.uint8)
        rows = msg.reshape(H, row)[:, :W * step].reshape(H, W, step)
        rows[:, :, :12] = pts[:, :3].reshape(H, W, 3).view(np.uint8).reshape(H, W, 12)
        rows[:, :, 16:20] = 0x7f                                     # rgb bytes: must be ignored
        b.push_pointcloud2(msg, W, H, step, row, (0, 4, 8), T, i)
    a.update(); b.update()
    assert_same(b.extract(), a.extract(), RESULT_FIELDS, "PointCloud2: ")
    assert_same(b.state(), a.state(), STATE_FIELDS, "PointCloud2 state: ")
    with pytest.raises(pcf.PcfError):
        b.push_pointcloud2(np.zeros(64, np.uint8), 2, 1, 18, 36, (0, 4, 8), np.eye(4), 10)   # point_step not a multiple of 4
    a.close(); b.close()

    og = oracle.OracleGrid(g.box, g.res)
    fus = pcf.Fusion(g.box, g.res)
    rng = np.random.default_rng(5)
    for i in range(3):
        pts, T = small.frame(i)
        world = oracle.kat_transform(T, pts[np.isfinite(pts[:, 2])])
        world[::97, 0] = np.nan                                      # D11: non-finite points are dropped
        vp = rng.uniform(-1, 1, 3).astype(np.float32)
        og.add_points_world(world[np.isfinite(world[:, 0])], vp)
        fus.add_points(np.ascontiguousarray(world), vp, i)
    og.update(); fus.update()
    assert_result_parity(fus.extract(), og.download(), "add_points: ")
    assert_same(fus.state(), og.state(), STATE_FIELDS, "add_points state: ")
    fus.close()


@pytest.mark.parametrize("kind", ["oracle", "ref_ordered"])
def test_anisotropic_resolution_and_offset_box(pcf, oracle, small, kind):
    """setResolution(x, y, z) with three different values and the launch-file style box (not centred, zmin = 0): the walk
    uses xres_ for every axis (OG.hpp:405), dims truncate per axis (OG.hpp:623-625).  Checked against the restatement
    AND against the reference's own header (oracle/_ref, D3-ordered build) when it is available."""
    if kind != "oracle" and not oracle.available(kind):
        pytest.skip("oracle/_ref not built on this box")
    box = (-0.21, 0.24, -0.2, 0.26, 0.0, 0.23)
    res = (0.005, 0.004, 0.006)
    og = oracle.OracleGrid(box, res, kind=kind)
    fus = pcf.Fusion(box, res)
    assert fus.dims == og.dims
    for i in range(small.n_frames):
        pts, T = small.frame(i)
        T = T.copy(); T[2, 3] += 0.1          # lift the sphere so that it straddles the z = 0 face of the box
        fus.push_frame(pts, T, i)
        og.add_frame(pts, T)
        if i == 2:
            fus.update(); og.update()
    fus.update(); og.update()
    want = og.download()
    assert len(want) > 500
    assert_result_parity(fus.extract(), want, f"anisotropic/{kind} result.")
    fus.close()


def test_error_behaviour(pcf, small):
    """Errors are status codes with a message, never exceptions across the ABI or silent acceptance."""
    g = small.grid
    with pytest.raises(pcf.PcfError, match="k_neighbourhood"):
        import ctypes as C
        lib = pcf.load_library()
        b = importlib_binding(pcf)
        cfg = b._Config()
        lib.pcf_default_config(C.byref(cfg))
        cfg.k_neighbourhood = 3
        h = C.c_void_p()
        rc = lib.pcf_create(C.byref(cfg), C.byref(h))
        assert rc == -1 and not h.value
        raise pcf.PcfError(rc, lib.pcf_last_error(None).decode())
    with pytest.raises(pcf.PcfError, match="20-bit|32 bit"):
        pcf.Fusion((-600.0, 600.0) * 3, 0.001)                       # 1.2e6 cells per axis
    fus = pcf.Fusion(g.box, g.res)
    pts, T = small.frame(0)
    fus.push_frame(pts, T, 5)
    with pytest.raises(pcf.PcfError, match="does not increase"):
        fus.push_frame(pts, T, 5)
    with pytest.raises(pcf.PcfError, match="bad frame arguments"):
        fus.push_frame(np.zeros((4, 2), np.float32), T, 6)           # stride < 3
    with pytest.raises(pcf.PcfError, match="max_frames"):
        fus.push_frame(pts, T, 1 << 20)
    fus.update()
    assert len(fus.extract()) >= 0                                   # the context is still usable after the errors
    with pytest.raises(pcf.PcfError, match="interleaved"):
        fus.exchange_counts([0, fus.dims[0] + 1])                    # sharded merge after an update pass is refused
    fus.close()


def importlib_binding(pcf):
    import importlib
    return importlib.import_module(pcf.__name__ + ".binding")


def test_edge_cases_empty_ragged_unaligned(pcf, oracle, small, tmp_path):
    """Empty clouds, 1-point clouds, lengths around the 256-point chunk size, all-NaN and all-outside clouds, packed xyz
    with a length that is not a multiple of 4 and a cloud that starts at an odd float offset (generic kernel instead of the
    bulk-copy one): the grid and the extraction must equal the oracle's; an empty grid extracts nothing and process()
    writes header-only files."""
    g = small.grid
    fus, og = pcf.Fusion(g.box, g.res), oracle.OracleGrid(g.box, g.res)
    assert len(fus.extract()) == 0                                   # nothing integrated yet
    fus.process(str(tmp_path / "empty.pcd"), str(tmp_path / "empty.csv"))
    assert (tmp_path / "empty.csv").read_text().count("\n") == 1 and "POINTS 0" in (tmp_path / "empty.pcd").read_text()
    pts, T = small.frame(0)
    fid = 0

    def both(cloud, pose):
        nonlocal fid
        fus.push_frame(cloud, pose, fid)
        og.add_frame(np.ascontiguousarray(cloud[:, :3]) if cloud.shape[1] != 4 else cloud, pose)
        fid += 1

    ok = np.isfinite(pts[:, 2])
    valid = pts[ok]
    both(np.zeros((0, 4), np.float32), T)                            # empty
    for n in (1, 255, 256, 257, 513):
        both(np.ascontiguousarray(valid[:n]), T)                     # ragged lengths
    both(np.full((300, 4), np.nan, np.float32), T)                   # all NaN
    far = valid[:300].copy(); far[:, :3] *= 50.0
    both(far, T)                                                     # all clipped / outside the box
    both(np.ascontiguousarray(valid[:1001, :3]), T)                  # packed xyz, n % 4 != 0 -> generic kernel
    both(np.ascontiguousarray(valid[:1000, :3]), T)                  # packed xyz, bulk-copy path
    backing = np.zeros(4 * len(valid) + 1, np.float32)
    odd = backing[1:].reshape(-1, 4)                                 # 4-byte aligned only
    odd[:] = valid
    both(odd, T)
    for i in range(1, small.n_frames):                               # enough surface for normals to appear
        both(*small.frame(i))
    fus.update(); og.update()
    assert_same(fus.state(), og.state(), STATE_FIELDS, "edge state.")
    want = og.download()
    assert len(want) > 100
    assert_result_parity(fus.extract(), want, "edge result.")
    fus.close()


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("PCF_FUZZ_SEEDS", "16"))))
def test_randomised_scenes_and_schedules(pcf, oracle, seed):
    """Seeded fuzzing of the whole path: random sphere size / stand-off / image size / resolution (anisotropic in half of the
    cases) / box offset / number of frames / update schedule (updates after random frames, sometimes twice in a row, sometimes
    none before the end) -- state and extraction bit-exact against the oracle every time."""
    synth = _synth(pcf)
    rng = np.random.default_rng(1000 + seed)
    radius = float(rng.uniform(0.06, 0.16))
    standoff = float(rng.uniform(0.30 + radius, 0.58))          # keeps the visible cap inside the 0.28-0.6 m depth clip
    w, h = int(rng.choice([96, 160, 200])), int(rng.choice([72, 120, 150]))
    res0 = float(rng.choice([0.003, 0.004, 0.005, 0.0065]))
    n_frames = int(rng.integers(3, 9))
    scene = synth.sphere_turntable(n_frames, w, h, res0, radius=radius, standoff=standoff, box_half=radius + 0.05,
                                   noise_sigma=float(rng.uniform(0.0002, 0.001)))
    off = rng.uniform(-0.02, 0.02, 3)
    half = radius + 0.05
    box = (-half + off[0], half + off[0], -half + off[1], half + off[1], -half * float(rng.uniform(0.3, 1.0)) + off[2], half + off[2])
    res = (res0, res0, res0) if seed % 2 == 0 else tuple(float(res0 * f) for f in rng.uniform(0.8, 1.25, 3))
    fus, og = pcf.Fusion(box, res), oracle.OracleGrid(box, res)
    assert fus.dims == og.dims
    updates = set(int(i) for i in rng.choice(n_frames, size=int(rng.integers(0, n_frames)), replace=False))
    for i in range(n_frames):
        pts, T = scene.frame(i)
        if rng.random() < 0.3:                                   # ragged: drop a random tail of the cloud
            pts = np.ascontiguousarray(pts[: int(rng.integers(1, len(pts)))])
        fus.push_frame(pts, T, i)
        og.add_frame(pts, T)
        if i in updates:
            fus.update(); og.update()
            if rng.random() < 0.25:
                fus.update(); og.update()                        # a second pass with no new frames in between
    fus.update(); og.update()
    assert_same(fus.state(), og.state(), STATE_FIELDS, f"fuzz {seed} state.")
    assert_result_parity(fus.extract(), og.download(), f"fuzz {seed} result.")
    fus.close()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("update_every", [1, 2, 4, None])
def test_interleaved_schedule_across_ranks_byte_identical(pcf, oracle, small, world, update_every):
    """an update pass between frames (the node's cleanGrid timer, node.cpp:301-325) combined with frame
    sharding.  Replicated-state mode (pcf_round_export / _install, pcf_update_local / _commit): the frames of every round are
    split over the ranks, records and normal records are gathered at every pass; the slabs' extractions concatenated in rank
    order equal the oracle (and so one GPU) bit for bit -- incremental dependants scoring (OG.hpp:244-277), unbuffered points
    (OG.hpp:210-216) and last-registrant holders (OG.hpp:443-449) included."""
    import importlib
    sh = importlib.import_module(pcf.__name__ + ".sharded")
    g = small.grid
    frames = [small.frame(i) for i in range(small.n_frames)]
    og = oracle.OracleGrid(g.box, g.res)
    run_schedule(og, small, update_every)
    want = og.download()
    assert len(want) > 1000
    ranks = [pcf.Fusion(g.box, g.res) for _ in range(world)]
    got = sh.interleaved_local(ranks, frames, update_every)
    assert_result_parity(got, want, f"interleaved x{world} every {update_every}: ")
    for f in ranks:
        f.close()
