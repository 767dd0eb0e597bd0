"""CPU tests of the parity oracle: known-answer tests and independent cross-checks
(scipy cKDTree, numpy brute force).  The reference repository holds no golden vectors for this path
(parity unpinned), so these analytic properties are what pins the restatement."""
import numpy as np
import pytest
from scipy.spatial import cKDTree


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def test_mt19937_known_answer(orc):
    # std::mt19937 / boost::mt19937: 10000th output of the default-seeded (5489) engine is 4123659995
    assert orc.mt19937_nth(5489, 10000) == 4123659995


def test_radius_search_matches_ckdtree(orc, synth):
    s = synth.make_scene(("y",), 20000, scene_id=3)
    q = s[::37]
    r = 0.03
    off, idx, d2 = orc.radius_search(s, q, r)
    tree = cKDTree(s.astype(np.float64))
    r2 = np.float32(r * r)
    eps = 1e-6
    for i in range(len(q)):
        got = idx[off[i]:off[i + 1]]
        gd = d2[off[i]:off[i + 1]]
        assert np.all(gd < r2)
        assert np.all(np.diff(gd) >= 0)                      # sorted ascending
        same = np.diff(gd) == 0
        assert np.all(np.diff(got)[same] > 0)               # ties by index
        dd = ((s.astype(np.float64) - q[i].astype(np.float64)) ** 2).sum(1)
        inner = set(np.flatnonzero(dd < r * r * (1 - eps)))
        outer = set(np.flatnonzero(dd < r * r * (1 + eps)))
        assert inner <= set(got.tolist()) <= outer
        assert inner <= set(tree.query_ball_point(q[i].astype(np.float64), r))
    # float32 L2_Simple distances are reproduced exactly
    i = 5
    got = idx[off[i]:off[i + 1]]
    d = s[got] - q[i]
    ref = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
    assert np.array_equal(ref.astype(np.float32), d2[off[i]:off[i + 1]])


def test_radius_search_empty_and_nan(orc):
    s = np.array([[0, 0, 0], [np.nan, 0, 0], [1, 0, 0], [0.005, 0, 0]], dtype=np.float32)
    q = np.array([[0, 0, 0], [5, 5, 5], [np.nan, 0, 0]], dtype=np.float32)
    off, idx, d2 = orc.radius_search(s, q, 0.01)
    assert off.tolist() == [0, 2, 2, 2]
    assert idx.tolist() == [0, 3]          # NaN surface row dropped, original indices kept
    # strict '<': a point exactly at the radius is excluded
    s2 = np.array([[0, 0, 0], [0.5, 0, 0]], dtype=np.float32)
    off, idx, _ = orc.radius_search(s2, s2[:1], 0.5)
    assert idx.tolist() == [0]


@pytest.mark.parametrize("k", [1, 10, 50])
def test_knn_matches_ckdtree(orc, synth, k):
    s = synth.make_scene(("diagonal",), 30000, scene_id=4)
    q = s[::53]
    idx, d2, kk = orc.knn_search(s, q, k)
    assert kk == k
    tree = cKDTree(s.astype(np.float64))
    dref, iref = tree.query(q.astype(np.float64), k=k)
    dref = dref.reshape(len(q), k)
    iref = iref.reshape(len(q), k)
    assert np.all(np.diff(d2, axis=1) >= 0)
    # identical sets except where the k-th / (k+1)-th distances are within eps (tie order is
    # traversal dependent in FLANN)
    dk1, _ = tree.query(q.astype(np.float64), k=k + 1)
    gap = (dk1[:, k] - dk1[:, k - 1]) > 1e-6
    bad = 0
    for i in np.flatnonzero(gap):
        if set(idx[i].tolist()) != set(iref[i].tolist()):
            bad += 1
    assert bad == 0
    np.testing.assert_allclose(np.sqrt(d2), dref, rtol=0, atol=2e-6)


def test_knn_clamps_k(orc):
    s = _rng(1).normal(size=(7, 3)).astype(np.float32)
    idx, d2, kk = orc.knn_search(s, s, 10)
    assert kk == 7
    assert np.all(idx[:, 7:] == -1) and np.all(np.isinf(d2[:, 7:]))
    assert np.all(idx[:, 0] == np.arange(7))


def test_normals_plane_and_cylinder(orc):
    rng = _rng(7)
    n = 4000
    # plane z = 1 (viewpoint at the origin → normal must point to -z), curvature 0
    p = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), np.ones(n)], 1).astype(np.float32)
    nm = orc.normals(p, k=10)
    # PCL 1.8's single-pass float covariance on raw coordinates loses digits (cov = E[xx]-E[x]^2);
    # the restatement keeps that behaviour, hence the loose tolerance
    assert np.all(nm[:, 2] < 0)
    assert np.median(np.abs(nm[:, 2])) > 0.999
    assert np.median(nm[:, 3]) < 1e-3
    # cylinder of radius 0.5 about the z axis, centred on the origin: normals are radial, inward
    # (towards the viewpoint on the axis)
    ang = rng.uniform(0, 2 * np.pi, n)
    c = np.stack([0.5 * np.cos(ang), 0.5 * np.sin(ang), rng.uniform(-0.2, 0.2, n)], 1).astype(np.float32)
    nc = orc.normals(c, k=20)
    radial = c[:, :2] / 0.5
    cosang = -(nc[:, 0] * radial[:, 0] + nc[:, 1] * radial[:, 1])
    assert np.median(cosang) > 0.995
    # k and radius both set / both unset → error like Feature::initCompute
    with pytest.raises(ValueError):
        orc.normals(p, k=10, radius=0.1)
    with pytest.raises(ValueError):
        orc.normals(p, k=0, radius=0.0)


def test_normals_follow_spec_arithmetic(orc, synth):
    """Recompute one normal in numpy float32 following Appendix A.2 step by step."""
    s = synth.make_model("y", 3000)
    k = 12
    idx, _, _ = orc.knn_search(s, s, k)
    nm = orc.normals(s, k=k)
    f = np.float32
    for i in (0, 17, 1234):
        acc = np.zeros(9, dtype=np.float32)
        for j in idx[i]:
            x, y, z = s[j]
            acc += np.array([x * x, x * y, x * z, y * y, y * z, z * z, x, y, z], dtype=np.float32)
        acc = acc / f(k)
        cov = np.array([[acc[0] - acc[6] * acc[6], acc[1] - acc[6] * acc[7], acc[2] - acc[6] * acc[8]],
                        [0, acc[3] - acc[7] * acc[7], acc[4] - acc[7] * acc[8]],
                        [0, 0, acc[5] - acc[8] * acc[8]]], dtype=np.float32)
        cov = cov + np.triu(cov, 1).T
        ev, vec = orc.eigen33_smallest(cov)
        if -(s[i] @ vec) < 0:
            vec = -vec
        np.testing.assert_allclose(nm[i, :3], vec, atol=1e-6)
        w, v = np.linalg.eigh(cov.astype(np.float64))
        assert abs(abs(v[:, 0] @ vec.astype(np.float64)) - 1) < 5e-3


def test_eigh3_and_umeyama(orc):
    rng = _rng(3)
    for _ in range(20):
        a = rng.normal(size=(3, 3))
        a = a @ a.T
        w, v = orc.eigh3(a)
        wr, _ = np.linalg.eigh(a)
        np.testing.assert_allclose(w, wr, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(a @ v, v * w, atol=1e-10)
    # rigid transform recovered from 3 points
    from scipy.spatial.transform import Rotation
    for seed in range(10):
        R = Rotation.random(random_state=seed).as_matrix()
        t = rng.normal(size=3)
        src = rng.normal(size=(3, 3))
        dst = src @ R.T + t
        T = orc.umeyama3(src, dst)
        np.testing.assert_allclose(T[:3, :3], R, atol=1e-9)
        np.testing.assert_allclose(T[:3, 3], t, atol=1e-9)
        assert abs(np.linalg.det(T[:3, :3]) - 1) < 1e-9


def _rigid(seed):
    from scipy.spatial.transform import Rotation
    R = Rotation.random(random_state=seed).as_matrix()
    t = _rng(seed).uniform(-1, 1, 3)
    return R, t


def test_shot_unit_norm_and_rigid_invariance(orc, synth):
    m = synth.make_model("y", 6000)
    nm = orc.normals(m, k=10)
    kp = synth.uniform_sampling(m, 0.03)
    d, rf = orc.shot352(m, nm, kp, 0.04)
    ok = np.isfinite(d[:, 0])
    assert ok.sum() > 0.9 * len(kp)
    np.testing.assert_allclose(np.linalg.norm(d[ok], axis=1), 1.0, atol=1e-5)
    # frames are orthonormal and right handed
    F = rf[ok].reshape(-1, 3, 3)
    np.testing.assert_allclose(np.einsum("nij,nkj->nik", F, F), np.tile(np.eye(3), (len(F), 1, 1)), atol=1e-5)
    assert np.all(np.linalg.det(F.astype(np.float64)) > 0.99)
    # rigid invariance: rotate cloud, normals and keypoints (normals are rotated, not re-estimated,
    # so the test isolates the LRF + histogram)
    R, t = _rigid(5)
    m2 = (m.astype(np.float64) @ R.T + t).astype(np.float32)
    kp2 = (kp.astype(np.float64) @ R.T + t).astype(np.float32)
    nm2 = nm.copy()
    nm2[:, :3] = (nm[:, :3].astype(np.float64) @ R.T).astype(np.float32)
    d2, _ = orc.shot352(m2, nm2, kp2, 0.04)
    both = ok & np.isfinite(d2[:, 0])
    err = np.linalg.norm(d[both] - d2[both], axis=1)
    # float32 re-rounding of the rotated coordinates moves a few neighbours across the radius /
    # sector boundaries; the bulk must agree tightly
    assert np.median(err) < 2e-3
    assert np.mean(err < 2e-2) > 0.9


def test_shot_nan_rows(orc):
    rng = _rng(9)
    s = rng.uniform(-1, 1, (200, 3)).astype(np.float32)
    nm = orc.normals(s, k=5)
    kp = np.array([[10, 10, 10], [np.nan, 0, 0]], dtype=np.float32)   # no neighbours / non-finite
    d, rf = orc.shot352(s, nm, kp, 0.05)
    assert np.all(np.isnan(d)) and np.all(np.isnan(rf))
    lrf = orc.shot_lrf(s, kp, 0.05)
    assert np.all(np.isnan(lrf))


def test_fpfh_blocks_sum_to_100(orc, synth):
    m = synth.make_model("horizontal", 3000)
    kp = synth.voxel_grid(m, 0.01)
    nm = orc.normals(kp, radius=0.04)
    f = orc.fpfh33(kp, nm, 0.04)
    ok = np.isfinite(f[:, 0])
    assert ok.sum() > 0.95 * len(kp)
    sums = f[ok].reshape(-1, 3, 11).sum(2)
    np.testing.assert_allclose(sums, 100.0, atol=2e-3)
    assert np.all(f[ok] >= 0)
    # general form (query cloud given explicitly) equals the every-point form
    f2 = orc.fpfh33(kp, nm, 0.04, q=kp[::7])
    np.testing.assert_array_equal(f2, f[::7])


def test_match_self_identity_and_bruteforce(orc):
    rng = _rng(11)
    a = rng.uniform(0, 1, (300, 352)).astype(np.float32)
    a /= np.linalg.norm(a, axis=1, keepdims=True)
    c = orc.match(a, a, mode=1, thr=0.25)
    assert np.array_equal(c["index_query"], np.arange(300)) and np.array_equal(c["index_match"], np.arange(300))
    assert np.all(c["distance"] == 0)
    b = a[:120] + rng.normal(0, 0.02, (120, 352)).astype(np.float32)
    b[7, 0] = np.nan                        # scene row skipped (pcl_isfinite(descriptor[0]))
    a2 = a.copy()
    a2[3, 100] = np.inf                     # model row dropped at tree build
    c = orc.match(a2, b, mode=1, thr=0.25)
    d = ((b[:, None, :].astype(np.float64) - a2[None].astype(np.float64)) ** 2).sum(2)
    d[:, 3] = np.inf
    ref_nn = np.argmin(d, axis=1)
    keep = [i for i in range(120) if i != 7 and d[i, ref_nn[i]] < 0.25]
    assert c["index_match"].tolist() == keep
    assert c["index_query"].tolist() == ref_nn[keep].tolist()
    c2 = orc.match(a2, b, mode=2)
    assert c2["index_match"].tolist() == [i for i in range(120) if i != 7]
    assert np.array_equal(orc.match(a2, b, mode=1, thr=0.25, omp=True), c)


def test_gc_recovers_known_transform(orc):
    rng = _rng(13)
    R, t = _rigid(21)
    model = rng.uniform(-0.3, 0.3, (60, 3)).astype(np.float32)
    scene_in = (model.astype(np.float64) @ R.T + t).astype(np.float32)
    clutter = rng.uniform(-2, 2, (80, 3)).astype(np.float32)
    scene = np.concatenate([scene_in, clutter])
    from oracle.pcl_oracle import CORR_DTYPE
    corrs = np.zeros(90, dtype=CORR_DTYPE)
    corrs["index_query"][:60] = np.arange(60)
    corrs["index_match"][:60] = np.arange(60)
    corrs["index_query"][60:] = rng.integers(0, 60, 30)
    corrs["index_match"][60:] = rng.integers(60, 140, 30)
    corrs["distance"] = rng.uniform(0, 0.2, 90).astype(np.float32)
    corrs = corrs[rng.permutation(90)]
    T, inst = orc.gc_recognize(model, scene, corrs, 0.01, 5)
    assert len(T) >= 1
    sizes = [len(i) for i in inst]
    b = int(np.argmax(sizes))
    assert sizes[b] >= 55
    np.testing.assert_allclose(T[b][:3, :3], R, atol=1e-4)
    np.testing.assert_allclose(T[b][:3, 3], t, atol=1e-4)
    assert np.all(inst[b]["index_query"] == inst[b]["index_match"])
    # too few correspondences → no instance
    T2, _ = orc.gc_recognize(model, scene, corrs[:3], 0.01, 5)
    assert len(T2) == 0


def test_keypoint_extractors_match_harness(orc, synth):
    """oracle UniformSampling / VoxelGrid (Appendix A.9 semantics) against the numpy harness versions that feed
    the benchmark, plus their defining properties."""
    cloud = synth.make_scene(("diagonal",), 30000, scene_id=6)
    cloud[17] = np.nan
    us, idx = orc.uniform_sampling(cloud, 0.02, return_index=True)
    ok = np.isfinite(cloud).all(1)
    assert np.array_equal(us, synth.uniform_sampling(cloud[ok], 0.02))
    assert np.array_equal(cloud[idx], us)                       # input points are kept, not averaged
    cells = np.floor(us / np.float32(0.02)).astype(np.int64)
    assert len(np.unique(cells, axis=0)) == len(us)             # one per leaf
    vg = orc.voxel_grid(cloud, 0.03)
    assert np.array_equal(vg, synth.voxel_grid(cloud[ok], 0.03))
    assert len(vg) == len(np.unique(np.floor(cloud[ok] / np.float32(0.03)).astype(np.int64), axis=0))
    vg3 = orc.voxel_grid(cloud, (0.03, 0.05, 0.02))
    assert 0 < len(vg3) < len(cloud)
    with pytest.raises(ValueError):
        orc.uniform_sampling(cloud, 1e-5)                       # lattice too fine (PCL: leaf size too small)
    assert len(orc.uniform_sampling(np.zeros((0, 3), np.float32), 0.01)) == 0


def test_hough3d_recovers_known_transform(orc):
    """Hough3DGrouping (restated): correspondences of a rigidly moved model, with consistent reference frames,
    all vote for the moved centroid — one bin holds them, RANSAC returns the transform; outliers do not."""
    from scipy.spatial.transform import Rotation
    rng = np.random.Generator(np.random.PCG64(3))
    model = rng.uniform(-0.3, 0.3, (60, 3)).astype(np.float32)
    R = Rotation.random(random_state=5).as_matrix()
    t = np.array([0.4, -0.2, 1.1])
    scene_in = (model.astype(np.float64) @ R.T + t).astype(np.float32)
    clutter = rng.uniform(-2, 2, (40, 3)).astype(np.float32)
    scene = np.concatenate([scene_in, clutter])
    mrf = np.stack([Rotation.random(random_state=100 + i).as_matrix() for i in range(60)]).astype(np.float32)
    srf_in = np.einsum("nij,kj->nik", mrf.astype(np.float64), R).astype(np.float32)     # rows (axes) rotated by R
    srf = np.concatenate([srf_in, np.stack([Rotation.random(random_state=900 + i).as_matrix() for i in range(40)])
                          .astype(np.float32)])
    corrs = np.zeros(80, dtype=orc.CORR_DTYPE)
    corrs["index_query"][:60] = np.arange(60)
    corrs["index_match"][:60] = np.arange(60)
    corrs["index_query"][60:] = rng.integers(0, 60, 20)
    corrs["index_match"][60:] = rng.integers(60, 100, 20)
    corrs["distance"] = rng.uniform(0, 0.2, 80).astype(np.float32)
    T, inst = orc.hough3d_recognize(model, mrf.reshape(-1, 9), scene, srf.reshape(-1, 9), corrs, 0.05, 5.0)
    assert len(T) >= 1
    b = int(np.argmax([len(i) for i in inst]))
    assert len(inst[b]) >= 50
    assert np.abs(T[b][:3, :3] - R).max() < 1e-3 and np.abs(T[b][:3, 3] - t).max() < 1e-3
    # a threshold above every bin yields nothing; a relative threshold keeps only the largest bin
    assert len(orc.hough3d_recognize(model, mrf.reshape(-1, 9), scene, srf.reshape(-1, 9), corrs, 0.05, 1000.0)[0]) == 0
    assert len(orc.hough3d_recognize(model, mrf.reshape(-1, 9), scene, srf.reshape(-1, 9), corrs, 0.05, -1.0)[0]) == 1


def test_icp_recovers_small_motion(orc):
    """IterativeClosestPoint (restated): a dense surface moved by a small rigid motion is pulled back; the
    convergence rules follow pcl::registration::DefaultConvergenceCriteria as ICP configures it."""
    from scipy.spatial.transform import Rotation
    from importlib import import_module
    synth = import_module("3d-object-detection-of-industrial-joints_b200.synth")
    model = synth.make_model("y", n=3000, seed=2)[:, :3]
    R = Rotation.from_rotvec([0.02, -0.03, 0.025]).as_matrix()
    t = np.array([0.004, -0.003, 0.002])
    moved = (model.astype(np.float64) @ R.T + t).astype(np.float32)
    # iteration cap: exactly max_iterations, "converged" (PCL reports ITERATIONS as convergence)
    r1 = orc.icp_align(moved[::3], model, max_iterations=1)
    assert r1["iterations"] == 1 and r1["converged"]
    r = orc.icp_align(moved[::3], model, max_iterations=100)
    assert r["converged"] and 1 < r["iterations"] <= 100
    assert r["fitness"] < r1["fitness"]
    Tinv = np.eye(4)
    Tinv[:3, :3], Tinv[:3, 3] = R.T, -R.T @ t
    assert np.abs(r["final_transform"] - Tinv).max() < 2e-3
    # aligned = source under the final transform
    ref = moved[::3].astype(np.float64) @ r["final_transform"][:3, :3].astype(np.float64).T + r["final_transform"][:3, 3]
    assert np.abs(r["aligned"] - ref).max() < 1e-5
    # a distance gate nothing passes: fewer than 3 correspondences, not converged, identity
    far = orc.icp_align(moved[::3] + 10.0, model, max_iterations=5, max_corr_dist=0.01)
    assert not far["converged"] and far["iterations"] == 0 and np.array_equal(far["final_transform"], np.eye(4))
    # the guess is applied first and is part of the final transformation
    g = np.eye(4, dtype=np.float32)
    g[:3, 3] = [-0.004, 0.003, -0.002]
    rg = orc.icp_align(moved[::3], model, max_iterations=100, guess=g)
    assert np.abs(rg["final_transform"] - Tinv).max() < 2e-3


def test_glibc_rand_known_answers(orc):
    """BOARD draws from rand(): the restated generator reproduces glibc's sequence (srand(1): 1804289383, ...)."""
    import ctypes
    assert [orc.glibc_rand_nth(1, i) for i in (1, 2, 3, 4, 5)] == [1804289383, 846930886, 1681692777, 1714636915,
                                                                  1957747793]
    libc = ctypes.CDLL(None)
    for seed in (1, 42, 12345):
        libc.srand(seed)
        ref = [libc.rand() for _ in range(400)]
        assert ref[0] == orc.glibc_rand_nth(seed, 1) and ref[99] == orc.glibc_rand_nth(seed, 100)
        assert ref[399] == orc.glibc_rand_nth(seed, 400)


def test_board_lrf_properties(orc):
    """BOARD frames (restated): orthonormal, z along the surface normal; on a plane with a straight border the x axis
    points into the hole (the empty side); fewer than 6 support points give NaN and draw no random numbers."""
    rng = np.random.Generator(np.random.PCG64(5))
    # half plane y <= 0.03 (border at y = 0.03), normals +z
    g = np.stack(np.meshgrid(np.arange(-0.1, 0.1001, 0.004), np.arange(-0.1, 0.0301, 0.004)), -1).reshape(-1, 2)
    surf = np.concatenate([g + rng.normal(0, 2e-4, g.shape), rng.normal(0, 1e-4, (len(g), 1))], 1).astype(np.float32)
    normals = np.zeros((len(surf), 4), np.float32)
    normals[:, :3] = [0.0, 0.0, 1.0] + rng.normal(0, 0.05, (len(surf), 3))     # (identical normals make PCL's
    normals[:, :3] /= np.linalg.norm(normals[:, :3], axis=1, keepdims=True)    # steepness ratio 0/0)
    kp = np.array([[0.0, 0.0, 0.0], [0.02, 0.005, 0.0], [5.0, 5.0, 5.0]], np.float32)
    # PCL's code leaves tangent_radius_ at 0 unless setTangentRadius is called (the reference does not call it): the
    # margin ring then degenerates to "every neighbour", and interior points fill all sectors.  With the tangent
    # radius set to the support radius the ring (0.85 r .. r) is cut by the border and the hole is found.
    rf0, used0 = orc.board_lrf(surf, normals, kp, 0.05)
    assert used0 == 4 and np.isnan(rf0[2]).all()
    for f in rf0[:2].reshape(2, 3, 3):
        assert np.abs(f @ f.T - np.eye(3)).max() < 1e-5 and f[2, 2] > 0.999
    tp = orc.board_params(tangent_radius=0.05)
    rf, used = orc.board_lrf(surf, normals, kp, 0.05, tp)
    assert used == 4 and np.isnan(rf[2]).all()
    for f in rf[:2].reshape(2, 3, 3):
        assert np.abs(f @ f.T - np.eye(3)).max() < 1e-5 and f[2, 2] > 0.999
        assert f[0, 1] > 0.55                      # x axis inside the empty sector (towards +y, where the surface ends)
        assert np.abs(np.cross(f[2], f[0]) - f[1]).max() < 1e-6
    # the frames do not depend on the random reference axis beyond the sector quantisation
    rf2, _ = orc.board_lrf(surf, normals, kp, 0.05, tp, rand_seed=99)
    assert np.abs(rf2[:2] - rf[:2]).max() < 0.3
    # without a hole (interior point, small radius) and without find_holes the x axis points to the most tilted normal
    normals2 = normals.copy()
    j = int(np.argmin(np.linalg.norm(surf[:, :2] - [0.01, -0.05], axis=1)))
    normals2[j, :3] = [0.6, 0.0, 0.8]
    kp2 = np.array([[0.0, -0.05, 0.0]], np.float32)
    for fh in (True, False):
        f = orc.board_lrf(surf, normals2, kp2, 0.03, orc.board_params(find_holes=fh))[0][0].reshape(3, 3)
        d = surf[j] - kp2[0]
        d[2] = 0
        assert np.dot(f[0], d / np.linalg.norm(d)) > 0.99


def test_cloud_utilities(orc):
    """removeNaNFromPointCloud keeps the finite rows in order; transformPointCloud applies the 4x4 in float32 and
    leaves non-finite rows alone."""
    from scipy.spatial.transform import Rotation
    rng = np.random.Generator(np.random.PCG64(11))
    c = rng.uniform(-1, 1, (50, 3)).astype(np.float32)
    c[3] = np.nan
    c[17, 1] = np.inf
    kept, idx = orc.remove_nan(c)
    assert idx.tolist() == [i for i in range(50) if i not in (3, 17)] and np.array_equal(kept, c[idx])
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = Rotation.from_rotvec([0.3, -0.2, 0.5]).as_matrix()
    T[:3, 3] = [0.1, 0.2, -0.3]
    out = orc.transform_points(c, T)
    ok = np.isfinite(c).all(1)
    ref = c[ok].astype(np.float64) @ T[:3, :3].astype(np.float64).T + T[:3, 3]
    assert np.abs(out[ok] - ref).max() < 1e-6
    assert np.array_equal(out[~ok], c[~ok], equal_nan=True)
    assert len(orc.remove_nan(np.zeros((0, 3), np.float32))[0]) == 0


def test_fpfh_plane_known_answer(orc):
    """Analytic known answer: on a plane with identical normals every pair has f1 = atan2(0, 1) = 0, f2 = v.n = 0,
    f3 = n.dp/|dp| = 0, i.e. the middle bin (5 of 0..10) of each of the three 11-bin blocks: every SPFH is
    [0,0,0,0,0,100,0,0,0,0,0] x 3, and so is every weighted, renormalised FPFH."""
    gx, gy = np.meshgrid(np.arange(0, 0.4, 0.01), np.arange(0, 0.4, 0.01))
    cloud = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], 1).astype(np.float32)
    normals = np.zeros((len(cloud), 4), np.float32)
    normals[:, 2] = 1.0
    f = orc.fpfh33(cloud, normals, 0.035)
    expected = np.zeros(33, np.float32)
    expected[[5, 16, 27]] = 100.0
    assert np.abs(f - expected).max() < 1e-3


def test_icp_translation_known_answer(orc):
    """A cloud shifted by less than half the point spacing is registered back exactly in one iteration: every nearest
    neighbour is the point's own original, so Umeyama returns the pure translation."""
    rng = np.random.Generator(np.random.PCG64(8))
    tgt = rng.uniform(-1, 1, (400, 3)).astype(np.float32)
    from scipy.spatial import cKDTree
    dmin = cKDTree(tgt).query(tgt, k=2)[0][:, 1].min()
    shift = np.array([0.2, -0.1, 0.15], np.float32) * np.float32(dmin)
    r = orc.icp_align(tgt + shift, tgt, max_iterations=3)
    assert np.abs(r["final_transform"][:3, :3] - np.eye(3)).max() < 1e-5
    assert np.abs(r["final_transform"][:3, 3] + shift).max() < 1e-6 and r["fitness"] < 1e-10


def test_icp_against_numpy_restatement(orc, synth):
    """Independent cross-check of the ICP loop: scipy cKDTree nearest neighbours + numpy SVD Umeyama (float64),
    same convergence rules; transforms agree to float32 rounding, iteration counts are equal."""
    from scipy.spatial import cKDTree
    from scipy.spatial.transform import Rotation
    model = synth.make_model("y", 2500, seed=4)
    R0 = Rotation.from_rotvec([0.015, 0.02, -0.01]).as_matrix()
    src = (model.astype(np.float64) @ R0.T + [0.002, -0.003, 0.001]).astype(np.float32)[::2]
    tree = cKDTree(model.astype(np.float64))

    def umeyama(a, b):
        ma, mb = a.mean(0), b.mean(0)
        S = (b - mb).T @ (a - ma) / len(a)
        U, _, Vt = np.linalg.svd(S)
        D = np.eye(3)
        if np.linalg.det(U) * np.linalg.det(Vt) < 0:
            D[2, 2] = -1
        Rm = U @ D @ Vt
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = Rm, mb - Rm @ ma
        return T

    for iters in (1, 4, 50):
        cur = src.copy()
        fin = np.eye(4, dtype=np.float32)
        prev, it, conv = np.finfo(np.float64).max, 0, False
        while True:
            # float32 L2_Simple distances like FLANN; the tree only proposes the neighbour
            nn = tree.query(cur.astype(np.float64))[1]
            d2 = ((cur - model[nn]) ** 2).astype(np.float32)
            d2 = (d2[:, 0] + d2[:, 1]) + d2[:, 2]
            T = umeyama(cur.astype(np.float64), model[nn].astype(np.float64)).astype(np.float32)
            cur = (cur @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
            fin = (T @ fin).astype(np.float32)
            it += 1
            if it >= iters:
                conv = True
                break
            if 0.5 * (np.trace(T[:3, :3].astype(np.float64)) - 1.0) >= 1.0 and float((T[:3, 3].astype(np.float64) ** 2).sum()) <= 0.0:
                conv = True
                break
            mse = float(d2.astype(np.float64).mean())
            if abs(mse - prev) < 1e-12:
                conv = True
                break
            prev = mse
        r = orc.icp_align(src, model, max_iterations=iters)
        assert r["iterations"] == it and r["converged"] == conv
        assert np.abs(r["final_transform"] - fin).max() < 2e-5


def test_board_against_numpy_restatement(orc, synth):
    """Independent cross-check of BOARD without the hole search (no random axis involved): z = eigenvector of the
    smallest eigenvalue of the support's scatter matrix (numpy eigh), signed by the mean support normal; x = unit
    projection on the tangent plane of the direction to the support point whose normal deviates most from z (first
    one in (distance, index) order on ties); y = z x x."""
    from scipy.spatial import cKDTree
    cloud = synth.make_model("y", 6000, seed=9)
    normals = orc.normals(cloud, k=12)
    kp = cloud[::150]
    r = 0.025
    rf, used = orc.board_lrf(cloud, normals, kp, r, orc.board_params(find_holes=False))
    assert used == 0
    tree = cKDTree(cloud.astype(np.float64))
    ok_rows = 0
    for i, c in enumerate(kp):
        idx = np.array(tree.query_ball_point(c.astype(np.float64), r * 1.0001))
        d2 = ((cloud[idx] - c) ** 2).astype(np.float32)
        d2 = (d2[:, 0] + d2[:, 1]) + d2[:, 2]
        keep = d2 < np.float32(r * r)
        idx, d2 = idx[keep], d2[keep]
        order = np.lexsort((idx, d2))
        idx, d2 = idx[order], d2[order]
        if len(idx) < 6:
            assert np.isnan(rf[i]).all()
            continue
        P = cloud[idx].astype(np.float64)
        w, V = np.linalg.eigh(np.cov((P - P.mean(0)).T, bias=True))
        z = V[:, 0]
        nm = normals[idx, :3].astype(np.float64)
        nm = nm[np.isfinite(nm).all(1)].sum(0)
        if z @ nm < 0:
            z = -z
        cosines = normals[idx, :3].astype(np.float64) @ z
        ring = d2 > 0                       # tangent_radius 0: every support point with d2 > 0 is a margin point
        cand = np.where(ring)[0] if ring.any() else np.arange(len(idx))
        j = cand[np.argmin(cosines[cand])]  # argmin returns the first minimum: (distance, index) order
        v = cloud[idx[j]].astype(np.float64) - c
        x = v - (v @ z) * z
        x /= np.linalg.norm(x)
        ref = np.concatenate([x, np.cross(z, x), z])
        if np.abs(rf[i] - ref).max() < 2e-5:
            ok_rows += 1
    assert ok_rows >= 0.97 * len(kp)       # the rest: float32 cosine ties / near-degenerate scatter matrices


def test_sensitivity_variants(orc, synth):
    """The oracle's variant switch (oracle/pcl_oracle.h): every variant is off by default and after a `with` block; the
    Appendix-A alternatives (Eigen 4-lane dot order, float32 Umeyama moments) stay far inside the north-star bars on a
    small case; the +-epsilon perturbations leave most rows bit-identical (they only move boundary cases)."""
    model = synth.make_model("y", 3000)
    kp = synth.voxel_grid(model, 0.02)
    nrm = orc.normals(model, k=10)
    base, rf = orc.shot352(model, nrm, kp, 0.03)
    with orc.variant("dot4"):
        alt, _ = orc.shot352(model, nrm, kp, 0.03)
    again, _ = orc.shot352(model, nrm, kp, 0.03)
    assert np.array_equal(base, again, equal_nan=True)            # the switch was restored
    ok = ~np.isnan(base[:, 0])
    assert np.linalg.norm(alt[ok] - base[ok], axis=1).max() < 1e-4
    with orc.variant("root_up"):
        nup = orc.normals(model, k=10)
    assert np.nanmax(np.abs(nup - nrm)) < 1e-3 and np.mean(np.abs(nup - nrm).max(1) <= 1e-6) > 0.95
    cloud = synth.voxel_grid(model, 0.01)
    nc = orc.normals(cloud, radius=0.05)
    f0 = orc.fpfh33(cloud, nc, 0.05)
    with orc.variant("fpfh_bin_up"):
        f1 = orc.fpfh33(cloud, nc, 0.05)
    with orc.variant("fpfh_skip"):
        f2 = orc.fpfh33(cloud, nc, 0.05)
    same = np.all((f0 == f1) | (np.isnan(f0) & np.isnan(f1)), axis=1)
    assert same.mean() > 0.9 and np.array_equal(f0, f2, equal_nan=True)
    # float32 Umeyama moments: a rigid fit moves by far less than the pose bars
    rng = np.random.default_rng(3)
    src = rng.normal(size=(20, 3))
    R = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    R *= np.sign(np.linalg.det(R))
    dst = src @ R.T + np.array([0.3, -0.2, 1.0])
    T0 = orc.umeyama3(src, dst)
    with orc.variant("umeyama_f32"):
        T1 = orc.umeyama3(src, dst)
    assert np.abs(T0 - T1).max() < 1e-5 and np.abs(T0[:3, :3] - R).max() < 1e-9
