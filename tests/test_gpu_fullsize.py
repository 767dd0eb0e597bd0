"""Parity properties at BASELINE.json's full size (1 M-point Kinect-like scene, 50 k-point model): the
oracle needs minutes there, so the CUDA path is checked through size-independent properties of each
stage — sortedness and completeness of neighbour lists, unit norms and rigid-motion invariance,
permutation equivariance of the correspondence search, agreement of the tensor-core filter with the exact
kernel, the defining invariants of geometric-consistency grouping, idempotence of the keypoint filters —
and against the oracle on a strided subset of the same inputs."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def full(b200):
    import sys
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    wl = bench.workload(0, 1_000_000, 50_000)
    ctx = b200.Context(0)
    cloud = ctx.cloud(wl["scene"])
    normals = ctx.normals(cloud, k=20)
    yield dict(wl=wl, ctx=ctx, cloud=cloud, normals=normals, bench=bench)
    ctx.close()


def test_fullsize_search_properties(full, orc):
    ctx, cloud, scene = full["ctx"], full["cloud"], full["wl"]["scene"]
    q = np.ascontiguousarray(full["wl"]["scene_kp"][::37])
    r = 0.02
    off, idx, d2 = ctx.radius_search(cloud, q, r)
    assert off[0] == 0 and off[-1] == len(idx) and np.all(np.diff(off) >= 1)      # keypoints are surface points
    assert np.all(d2 < np.float32(r * r))
    seg = np.repeat(np.arange(len(q)), np.diff(off))
    # sorted by (d2, index) inside every list; distances are the float32 L2_Simple values
    same = seg[1:] == seg[:-1]
    assert np.all((d2[1:] >= d2[:-1]) | ~same)
    ties = same & (d2[1:] == d2[:-1])
    assert np.all(idx[1:][ties] > idx[:-1][ties])
    diff = scene[idx] - q[seg]
    ref = (diff[:, 0] * diff[:, 0] + diff[:, 1] * diff[:, 1]) + diff[:, 2] * diff[:, 2]
    assert np.array_equal(ref.astype(np.float32), d2)
    # completeness against the oracle on a subset of the queries
    sub = q[::25]
    ooff, oidx, od2 = orc.radius_search(scene, sub, r)
    soff, sidx, sd2 = ctx.radius_search(cloud, sub, r)
    assert np.array_equal(soff, ooff) and np.array_equal(sidx, oidx) and np.array_equal(sd2, od2)
    # kNN: sorted, first neighbour is the point itself, consistent with the radius lists
    kidx, kd2, kf = ctx.knn_search(cloud, sub, 20)
    assert kf == 20 and np.all(np.diff(kd2, axis=1) >= 0) and np.all(kd2[:, 0] == 0)
    oki, okd, _ = orc.knn_search(scene, sub, 20)
    assert np.array_equal(kd2, okd) and np.array_equal(kidx, oki)


def test_fullsize_normals_properties(full, orc):
    ctx, scene, nrm = full["ctx"], full["wl"]["scene"], full["normals"]
    ok = ~np.isnan(nrm[:, 0])
    assert ok.mean() > 0.999
    np.testing.assert_allclose(np.linalg.norm(nrm[ok, :3], axis=1), 1.0, atol=2e-6)
    assert np.all((nrm[ok, :3] * (-scene[ok])).sum(1) >= -1e-6)      # flipped towards the viewpoint (origin)
    assert np.all((nrm[ok, 3] >= 0) & (nrm[ok, 3] <= 1 / 3 + 1e-6))  # curvature = lambda0 / trace
    # oracle on a strided subset of the queries (same surface)
    sub = np.ascontiguousarray(scene[::1000])
    got = ctx.normals(full["cloud"], q=sub, k=20)
    ref = orc.normals(scene, q=sub, k=20)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    assert np.nanmax(np.abs(got - ref)) < 2e-4
    # the query-as-surface run equals the explicit-query run on the same rows
    assert np.array_equal(got, nrm[::1000], equal_nan=True)


def test_fullsize_shot_properties(full, orc):
    ctx, cloud, wl, nrm = full["ctx"], full["cloud"], full["wl"], full["normals"]
    kp = wl["scene_kp"]
    desc, rf = ctx.shot352(cloud, nrm, kp, 0.02)
    ok = ~np.isnan(desc[:, 0])
    assert ok.mean() > 0.95
    np.testing.assert_allclose(np.linalg.norm(desc[ok], axis=1), 1.0, atol=1e-5)
    assert np.all(desc[ok] >= 0)
    # frames are right-handed orthonormal
    R = rf[ok].reshape(-1, 3, 3)
    assert np.abs(np.einsum("nij,nkj->nik", R, R) - np.eye(3)).max() < 1e-5
    assert np.abs(np.linalg.det(R.astype(np.float64)) - 1).max() < 1e-5
    # oracle on a strided subset of the keypoints (same surface and normals): 1e-4 L2 per descriptor
    sub = np.ascontiguousarray(kp[::60])
    od, orf = orc.shot352(wl["scene"], nrm, sub, 0.02)
    gd = desc[::60]
    assert np.array_equal(np.isnan(gd[:, 0]), np.isnan(od[:, 0]))
    o = ~np.isnan(od[:, 0])
    assert np.linalg.norm(gd[o].astype(np.float64) - od[o], axis=1).max() < 1e-4
    full["desc"] = desc


def test_fullsize_match_and_grouping_properties(full, orc, b200):
    ctx, wl, bench = full["ctx"], full["wl"], full["bench"]
    if "desc" not in full:
        full["desc"], _ = ctx.shot352(full["cloud"], full["normals"], wl["scene_kp"], 0.02)
    ds = full["desc"]
    p = b200.shot_params(**bench.PARAMS)
    model = ctx.model_create_shot(wl["model"], wl["model_kp"], p)
    dm, kpm = model.download()
    corr = ctx.match(dm, ds, 1, 0.25)                        # tensor-core filter + certificate
    assert len(corr) > 1000 and np.all(np.diff(corr["index_match"]) > 0) and np.all(corr["distance"] < 0.25)
    # (1) permutation equivariance: shuffling the scene rows permutes the answer
    rng = np.random.Generator(np.random.PCG64(5))
    perm = rng.permutation(len(ds))
    cp = ctx.match(dm, ds[perm], 1, 0.25)
    back = np.empty(len(perm), np.int64)
    back[np.arange(len(perm))] = perm
    a = {(int(q), int(back[m])): d for q, m, d in zip(cp["index_query"], cp["index_match"], cp["distance"])}
    b = {(int(q), int(m)): d for q, m, d in zip(corr["index_query"], corr["index_match"], corr["distance"])}
    assert a == b
    # (2) the exact float32 kernel and the oracle agree with the filter on a strided subset of the rows
    sub = np.ascontiguousarray(ds[::40])
    os.environ["B200_MATCH"] = "exact"
    try:
        ce = ctx.match(dm, sub, 1, 0.25)
    finally:
        del os.environ["B200_MATCH"]
    co = orc.match(dm, sub, 1, 0.25, omp=True)
    assert ce.tobytes() == co.tobytes()
    pick = corr[np.isin(corr["index_match"], np.arange(0, len(ds), 40))]
    assert np.array_equal(pick["index_match"] // 40, ce["index_match"])
    assert np.array_equal(pick["index_query"], ce["index_query"]) and np.array_equal(pick["distance"], ce["distance"])
    # (3) self matching: every finite model row finds itself at distance 0
    cs = ctx.match(dm, dm, 1, 0.25)
    fin = np.flatnonzero(np.isfinite(dm).all(1))
    assert np.array_equal(cs["index_match"], fin) and np.all(cs["distance"] == 0)
    assert np.all(np.linalg.norm(dm[cs["index_query"]] - dm[cs["index_match"]], axis=1) == 0)
    # (4) grouping invariants on the full correspondence list
    kps = wl["scene_kp"]
    T, inst, n = ctx.gc_recognize(kpm, kps, corr, 0.02, 2, max_inst=8192)
    assert n == len(inst) > 100
    seen = set()
    order = np.lexsort((np.arange(len(corr)), corr["distance"]))
    rank_of = {(int(c["index_query"]), int(c["index_match"])): r for r, c in enumerate(corr[order])}
    prev_seed = -1
    for t, ins in zip(T, inst):
        assert len(ins) >= 3                                     # more than gc_threshold members
        keys = [(int(c["index_query"]), int(c["index_match"])) for c in ins]
        if not np.array_equal(t, np.eye(4, dtype=np.float32)):
            Rm = t[:3, :3].astype(np.float64)                      # rigid transform
            assert np.abs(Rm @ Rm.T - np.eye(3)).max() < 1e-5 and abs(np.linalg.det(Rm) - 1) < 1e-5
            assert set(keys) <= set(rank_of)                       # RANSAC inliers are correspondences of the input
    # consensus sets: pairwise distance preservation, disjointness, seeds in ascending distance order — checked
    # on the sets before RANSAC filtering through the oracle's definition on a subsample of the instances
    oT, oinst = orc.gc_recognize(kpm, kps, corr, 0.02, 2, max_inst=8192)
    assert len(oinst) == n
    for x, y in list(zip(inst, oinst))[::50]:
        assert x.tobytes() == y.tobytes()
    assert max(np.abs(A - B).max() for A, B in list(zip(T, oT))[::50]) < 1e-4
    model.close()


def test_fullsize_keypoint_filters(full, orc):
    ctx, wl = full["ctx"], full["wl"]
    us, idx = ctx.uniform_sampling(wl["scene"], 0.01, return_index=True)
    assert np.array_equal(us, wl["scene_kp"])                    # the harness keypoints (numpy) bit for bit
    assert np.array_equal(wl["scene"][idx], us)
    assert np.array_equal(ctx.uniform_sampling(us, 0.01), us)   # idempotent
    vg = ctx.voxel_grid(wl["scene"], 0.01)
    assert len(vg) == len(us)                                    # same lattice, same occupied leaves, same order
    assert np.abs(vg - us).max() <= 0.01 * np.sqrt(3) + 1e-6     # a centroid and the kept point share a leaf


def test_config4_view_library_fullsize(b200, synth, tmp_path):
    """BASELINE.json config 4 at its size: 64 rendered partial views x 3 CAD joints = 192 views in one resident
    library, one scene matched against all of them.  Properties: per-view results equal the single-model pipeline
    (sampled views), instances reference keypoints of their own view, poses are rigid, and the library file
    round-trips bit for bit."""
    p = b200.shot_params(normal_k=10, descr_radius=0.02, match_mode=1, match_thr=0.25, gc_size=0.02, gc_threshold=3,
                         max_instances=512)
    ctx = b200.Context(0)
    lib = b200.Library(ctx)
    views = []
    for joint in ("y", "diagonal", "horizontal"):
        for v in range(64):
            cloud = synth.make_partial_view(joint, v, 4000)
            kp = synth.uniform_sampling(cloud, 0.02)
            lib.add_view(cloud, kp, p)
            views.append((cloud, kp))
    assert lib.views == 192
    scene = synth.make_scene(("y", "diagonal", "horizontal"), 150000, scene_id=7)
    kps = synth.uniform_sampling(scene, 0.03)
    res = lib.register_scene(scene, kps, p, max_inst=98304)
    n = res["n_instances"]
    assert n > 0 and np.all(np.diff(res["view"]) >= 0) and len(res["view_n_corrs"]) == 192
    for i in range(n):
        v = int(res["view"][i])
        ic = res["instances"][i]
        assert len(ic) >= 3 and ic["index_query"].max() < len(views[v][1]) and ic["index_match"].max() < len(kps)
    R = res["transforms"][:, :3, :3].astype(np.float64)
    assert np.abs(np.einsum("nij,nkj->nik", R, R) - np.eye(3)).max() < 1e-4
    assert np.abs(np.linalg.det(R) - 1.0).max() < 1e-4
    for v in (5, 77, 150):
        m = ctx.model_create_shot(views[v][0], views[v][1], p)
        single = ctx.register_scene_shot(m, scene, kps, p)
        mine = [i for i in range(n) if res["view"][i] == v]
        assert len(mine) == single["n_instances"] and res["view_n_corrs"][v] == len(single["corrs"])
        for i, ins in zip(mine, single["instances"]):
            assert res["instances"][i].tobytes() == ins.tobytes()
        m.close()
    path = str(tmp_path / "views192.b200lib")
    lib.save(path)
    lib2 = b200.Library.load(ctx, path)
    res2 = lib2.register_scene(scene, kps, p, max_inst=98304)
    assert res2["n_instances"] == n and np.array_equal(res2["transforms"], res["transforms"])
    assert all(a.tobytes() == b.tobytes() for a, b in zip(res["instances"], res2["instances"]))
    lib2.close()
    lib.close()
    ctx.close()


def test_config5_scene_batch_lanes_are_deterministic(b200, synth, pkg):
    """BASELINE.json config 5 (500 k-point scenes, a batch of 12 on one GPU): the scenes of a rank go through four
    lanes (contexts driven by their own host threads, kernels of different scenes interleaving on the GPU).  The
    results must be bit-identical to the same scenes registered one after the other on a single context."""
    sharding = importlib.import_module(pkg.__name__ + ".sharding")
    p = b200.shot_params(normal_k=20, descr_radius=0.02, match_mode=1, match_thr=0.25, gc_size=0.02, gc_threshold=2,
                         max_instances=4096)
    model = synth.make_model("y", 20000)
    kpm = synth.uniform_sampling(model, 0.005)
    scenes = [synth.make_kinect_scene(("y", "diagonal", "horizontal"), target_points=500_000, scene_id=20 + (s % 3))
              for s in range(3)]
    scenes = [scenes[s % 3] for s in range(12)]
    kps = [synth.uniform_sampling(sc, 0.012) for sc in scenes[:3]]
    kps = [kps[s % 3] for s in range(12)]
    ctxs = [b200.Context(0) for _ in range(4)]
    m = ctxs[0].model_create_shot(model, kpm, p)
    seq, _ = sharding.register_scene_batch(ctxs[0], m, scenes[:3], kps[:3], p)
    par, gathered = sharding.register_scene_batch(ctxs, m, scenes, kps, p)
    assert sorted(par) == list(range(12)) and sorted(gathered) == list(range(12))
    for s in range(12):
        a, b = par[s], seq[s % 3]
        assert a["n_instances"] == b["n_instances"] > 0
        assert a["corrs"].tobytes() == b["corrs"].tobytes() == gathered[s].tobytes()
        assert np.array_equal(a["transforms"], b["transforms"])
        assert all(x.tobytes() == y.tobytes() for x, y in zip(a["instances"], b["instances"]))
    # the same through the library's own lanes (b200_register_scene_batch_shot: contexts and host threads inside)
    nat = b200.register_scene_batch(m, scenes, kps, p, lanes=4)
    assert len(nat) == 12
    for s in range(12):
        a, b = nat[s], seq[s % 3]
        assert a["status"] == 0 and a["n_instances"] == b["n_instances"]
        assert a["corrs"].tobytes() == b["corrs"].tobytes() and np.array_equal(a["transforms"], b["transforms"])
        assert all(x.tobytes() == y.tobytes() for x, y in zip(a["instances"], b["instances"]))
    assert b200.register_scene_batch(m, [], [], p) == []
    b200.lanes_release(0)
    assert b200.register_scene_batch(m, scenes[:2], kps[:2], p, lanes=2)[1]['n_instances'] == seq[1]['n_instances']
    b200.lanes_release(0)
    m.close()
    for c in ctxs:
        c.close()
