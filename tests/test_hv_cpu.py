"""CPU checks of the hypothesis-verification restatement (oracle/hv_oracle.cpp; parity unpinned, see its header):
known answers of the cost function and of the streams it draws from, and the behaviour the reference relies on —
true poses survive, displaced and conflicting copies do not."""
import numpy as np

import hv_cases


def test_glibc_rand_and_mt19937_streams(orc):
    # srand(1): the first values of glibc's rand(); mt19937 default seed: the 10000th output (C++11 [rand.predef])
    assert [orc.glibc_rand_nth(1, k) for k in (1, 2, 3)] == [1804289383, 846930886, 1681692777]
    assert orc.mt19937_nth(5489, 10000) == 4123659995


def test_anneal_cost_of_known_lists(orc):
    # two hypotheses explaining the same three scene points with weight 1: each alone scores
    # -(3 - 0 - 0 - 0 - 1) = -2; both together 6 explained - 6 duplicity - 2 active = -(-2) = 2, so one is dropped
    eo = np.array([0, 3, 6], np.int32)
    ei = np.array([0, 1, 2, 0, 1, 2], np.int32)
    ew = np.ones(6, np.float32)
    oo = np.array([0, 0, 0], np.int32)
    p = orc.hv_params(detect_clutter=0)
    mask, cost, acc = orc.hv_optimize(3, eo, ei, ew, oo, np.zeros(0, np.int32), 1, np.ones(2, np.float32),
                                      np.zeros(2, np.int32), p)
    assert mask.sum() == 1 and cost == -2.0 and acc >= 1
    # disjoint sets: both stay, cost -(6 - 2) = -4
    ei2 = np.array([0, 1, 2, 3, 4, 5], np.int32)
    mask, cost, _ = orc.hv_optimize(6, eo, ei2, ew, oo, np.zeros(0, np.int32), 1, np.ones(2, np.float32),
                                    np.zeros(2, np.int32), p)
    assert mask.all() and cost == -4.0
    # outliers outweigh the explained points of the second hypothesis: regulariser 3 x 2 outliers > 3 - 1
    mask, cost, _ = orc.hv_optimize(6, eo, ei2, ew, oo, np.zeros(0, np.int32), 1, np.array([1, 3], np.float32),
                                    np.array([0, 2], np.int32), p)
    assert mask.tolist() == [True, False] and cost == -2.0
    # complete-model occupancy: two hypotheses sharing 2 cells pay 4 x (2 + 2) = 16 > the 2 the second one brings
    oo2 = np.array([0, 2, 4], np.int32)
    oi2 = np.array([5, 6, 5, 6], np.int32)
    mask, cost, _ = orc.hv_optimize(6, eo, ei2, ew, oo2, oi2, 8, np.ones(2, np.float32), np.zeros(2, np.int32), p)
    assert mask.sum() == 1 and cost == -2.0


def test_anneal_uniform_modes_and_seeds(orc):
    cues = hv_cases.random_cues(3)
    base = orc.hv_optimize(*cues, orc.hv_params(detect_clutter=0))
    again = orc.hv_optimize(*cues, orc.hv_params(detect_clutter=0))
    assert base[0].tolist() == again[0].tolist() and base[1] == again[1] and base[2] == again[2]
    greedy = orc.hv_optimize(*cues, orc.hv_params(detect_clutter=0, sa_uniform_mode=1))
    # with the raw-engine variate no uphill move is ever taken: every accepted move lowers the cost
    assert greedy[1] <= 0 or greedy[2] == 0
    # the best cost never exceeds the all-active start
    all_on = orc.hv_optimize(*cues, orc.hv_params(detect_clutter=0, max_iterations=0))
    assert all_on[2] == 0 and all_on[0].all()
    assert base[1] <= all_on[1] and greedy[1] <= all_on[1]


def test_verify_keeps_true_poses(orc, synth):
    scene, hyps, kind = hv_cases.cluttered(synth, 200000)
    p = orc.hv_params(detect_clutter=0, occlusion_reasoning=0, regularizer=3.0, radius_normals=0.02)
    r = orc.hv_verify(scene, hyps, p)
    info = r["info"]
    assert info["valid"].all() and (info["n_visible"] == 20000).all()
    for k, m, i in zip(kind, r["mask"], info):
        if k == "true":
            assert m and i["n_explained"] > 10 * i["n_outliers"]
        if k in ("displaced", "nowhere"):
            assert not m and i["n_outliers"] > i["n_explained"]
    # of a true pose and its near-duplicate exactly one survives (they explain the same scene points)
    for a in range(0, len(kind) - 1, 3):
        assert r["mask"][a] + r["mask"][a + 1] == 1
    assert len(r["expl_off"]) == len(hyps) + 1 and r["expl_off"][-1] == len(r["expl_idx"]) == info["n_explained"].sum()
    # explained lists ascend (std::map order) and stay inside the compacted scene
    for h in range(len(hyps)):
        seg = r["expl_idx"][r["expl_off"][h]:r["expl_off"][h + 1]]
        assert (np.diff(seg) > 0).all() and (len(seg) == 0 or seg[-1] < r["n_scene_points"])


def test_verify_occlusion_reasoning(orc, synth):
    scene, hyps, kind = hv_cases.kinect(synth, 200000)
    p = orc.hv_params(detect_clutter=0, occlusion_reasoning=1, regularizer=3.0, radius_normals=0.02)
    r = orc.hv_verify(scene, hyps, p)
    info = r["info"]
    # the z-buffers remove the far side of every tube and everything behind the scene's surface
    assert (info["n_visible"] < 0.3 * 20000).all()
    for k, m in zip(kind, r["mask"]):
        if k in ("displaced", "nowhere"):
            assert not m
    assert r["mask"][[i for i, k in enumerate(kind) if k in ("true", "near")]].any()
    # the reference's own setting: 5 mm normal radius on the 5 mm voxel grid leaves fewer points
    p2 = orc.hv_params(detect_clutter=0, occlusion_reasoning=1, regularizer=3.0, radius_normals=0.005)
    r2 = orc.hv_verify(scene, hyps, p2)
    assert r2["n_scene_points"] < r["n_scene_points"]
    assert (r2["info"]["n_points"] <= info["n_points"]).all()


def test_verify_edge_cases(orc, synth):
    scene, hyps, _ = hv_cases.cluttered(synth, 50000)
    p = orc.hv_params(detect_clutter=0, radius_normals=0.02)
    # no hypotheses
    r = orc.hv_verify(scene, [], p)
    assert len(r["mask"]) == 0
    # an empty hypothesis is invalid, the others are unaffected
    r1 = orc.hv_verify(scene, [hyps[0], np.zeros((0, 3), np.float32), hyps[3]], p)
    r0 = orc.hv_verify(scene, [hyps[0], hyps[3]], p)
    assert not r1["info"]["valid"][1] and not r1["mask"][1]
    assert r1["mask"][[0, 2]].tolist() == r0["mask"].tolist() and r1["best_cost"] == r0["best_cost"]
    # a scene far from every hypothesis: nothing explained, every point an outlier, nothing survives
    far = scene + np.float32(50.0)
    rf = orc.hv_verify(far, hyps[:2], p)
    assert (rf["info"]["n_explained"] == 0).all() and not rf["mask"].any()


def test_cues_against_an_independent_numpy_restatement(orc, synth):
    """The cue stage of the restatement (z-buffer visibility, voxel grids, explained points, outliers, occupancy cells)
    recomputed with numpy / scipy from the same description of PCL's code, written separately: visibility counts and
    occupancy counts exactly, explained / outlier counts up to the handful of pairs whose distance sits on the inlier
    radius (cKDTree's ball query is <=, FLANN's is <) or whose voxel centroid differs in the last bit."""
    from scipy.spatial import cKDTree
    scene, hyps, _ = hv_cases.kinect(synth, 150000)
    hyps = [h[::2] for h in hyps[:5]]
    p = orc.hv_params(detect_clutter=0, occlusion_reasoning=1, regularizer=3.0, radius_normals=0.03)
    r = orc.hv_verify(scene, hyps, p)
    f32 = np.float32

    def zbuffer(cloud, res):
        c = f32(res) / f32(2) - f32(0.5)
        bx, by = cloud[:, 0] / cloud[:, 2], cloud[:, 1] / cloud[:, 2]
        maxc = max(abs(bx.max()), abs(by.max()), abs(bx.min()), abs(by.min()))
        f = f32(c / maxc)

        def pix(pts):
            u = (f * pts[:, 0] / pts[:, 2] + c).astype(np.float32)
            v = (f * pts[:, 1] / pts[:, 2] + c).astype(np.float32)
            ui, vi = np.trunc(u).astype(np.int64), np.trunc(v).astype(np.int64)
            ok = (ui >= 0) & (vi >= 0) & (ui < res) & (vi < res)
            return ui, vi, ok
        ui, vi, ok = pix(cloud)
        depth = np.full((res, res), np.inf, np.float32)
        np.minimum.at(depth, (vi[ok], ui[ok]), cloud[ok, 2])

        def keeps(pts, thr):
            ui, vi, ok = pix(pts)
            d = np.full(len(pts), np.inf, np.float32)
            d[ok] = depth[vi[ok], ui[ok]]
            return ok & np.isfinite(d) & ~((pts[:, 2] - f32(thr)) > d)
        return keeps

    def voxel(cloud, leaf):
        ijk = np.floor(cloud / f32(leaf)).astype(np.int64)
        ijk -= ijk.min(0)
        key = (ijk[:, 2] * (ijk[:, 1].max() + 1) + ijk[:, 1]) * (ijk[:, 0].max() + 1) + ijk[:, 0]
        order = np.argsort(key, kind="stable")
        uk, start = np.unique(key[order], return_index=True)
        sums = np.add.reduceat(cloud[order].astype(np.float64), start)
        cnt = np.diff(np.append(start, len(cloud)))[:, None]
        return (sums / cnt).astype(np.float32)

    def with_normals(cloud, radius):
        t = cKDTree(cloud.astype(np.float64))
        n = np.array([len(x) for x in t.query_ball_point(cloud.astype(np.float64), radius * (1 - 1e-7))])
        return cloud[n >= 3]   # fewer than three neighbours: NaN normal, dropped

    scene_keeps = zbuffer(scene, 100)
    S = with_normals(voxel(scene, 0.005), 0.03)
    assert abs(len(S) - r["n_scene_points"]) <= 2
    tree = cKDTree(S.astype(np.float64))
    lo = np.min([h.min(0) for h in hyps], axis=0)
    for h, info in zip(hyps, r["info"]):
        vis = h[zbuffer(h, 75)(h, 0.005) & scene_keeps(h, 0.005)]
        assert len(vis) == info["n_visible"]
        M = with_normals(voxel(vis, 0.005), 0.03) if len(vis) else vis
        assert abs(len(M) - info["n_points"]) <= 2
        nb = tree.query_ball_point(M.astype(np.float64), 0.005 * (1 - 1e-7)) if len(M) else []
        outliers = sum(1 for x in nb if len(x) == 0)
        explained = len(set(j for x in nb for j in x))
        assert abs(outliers - info["n_outliers"]) <= 3 and abs(explained - info["n_explained"]) <= 3
        cells = np.floor((h - lo.astype(np.float32)) / f32(0.01)).astype(np.int64)
        assert len(np.unique(cells, axis=0)) == info["n_occupancy"]
