"""The C++ host side above the C ABI: apps/*.cpp drive the pipeline through the PCL-style adapters
(include/pcl_b200/pcl_b200.h) in the reference's call order (SHOT.cpp:298-482, FPFH_demo.cpp:405-538).
CPU: they build and fail loudly without a GPU.  GPU: their output equals the oracle chain."""
import importlib
import os
import subprocess

import numpy as np

import eps
import pytest

from conftest import PKG_NAME

CORR = np.dtype([("index_query", "<i4"), ("index_match", "<i4"), ("distance", "<f4")])


@pytest.fixture(scope="module")
def apps():
    builder = importlib.import_module(PKG_NAME + ".build")
    return {os.path.basename(p): p for p in builder.build_apps()}


def _write(path, a):
    np.ascontiguousarray(a[:, :3], dtype=np.float32).tofile(path)


def test_apps_build_and_need_a_gpu(apps, synth, tmp_path):
    import torch
    assert set(apps) >= {"shot_recognition", "fpfh_recognition", "batch_recognition", "sharded_recognition"}
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = synth.make_model("y", 500)
    for n in ("m", "mk", "s", "sk"):
        _write(tmp_path / (n + ".f32"), m)
    r = subprocess.run([apps["shot_recognition"], str(tmp_path / "m.f32"), str(tmp_path / "mk.f32"),
                        str(tmp_path / "s.f32"), str(tmp_path / "sk.f32"), str(tmp_path / "out")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr
    r = subprocess.run([apps["batch_recognition"], str(tmp_path / "m.f32"), str(tmp_path / "mk.f32"), str(tmp_path / "b"),
                        "2", str(tmp_path / "s.f32"), str(tmp_path / "sk.f32")], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def _read_instances(prefix):
    T = np.fromfile(prefix + ".T", dtype=np.float32).reshape(-1, 4, 4)
    raw = np.fromfile(prefix + ".inst", dtype=np.int32)
    inst, p = [], 0
    while p < len(raw):
        n = int(raw[p])
        inst.append(raw[p + 1:p + 1 + 3 * n].view(CORR).copy())
        p += 1 + 3 * n
    return T, inst


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["loop", "loop-single", "batch"])
def test_shot_recognition_app_matches_oracle(apps, orc, synth, tmp_path, mode):
    model = synth.make_model("y", 5000)
    scene = synth.make_scene(("y",), 30000, scene_id=3)
    kpm, kps = synth.uniform_sampling(model, 0.02), synth.uniform_sampling(scene, 0.03)
    for name, a in (("m", model), ("mk", kpm), ("s", scene), ("sk", kps)):
        _write(tmp_path / (name + ".f32"), a)
    prefix = str(tmp_path / "out")
    r = subprocess.run([apps["shot_recognition"]] + [str(tmp_path / (n + ".f32")) for n in ("m", "mk", "s", "sk")] +
                       [prefix, "10", "0.02", "0.25", "0.02", "2", mode], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "Model instances found" in r.stdout
    print([l for l in r.stdout.splitlines() if l.startswith("Correspondence search")])
    corr = np.fromfile(prefix + ".corr", dtype=CORR)
    T, inst = _read_instances(prefix)
    # oracle chain with the same parameters (descr_rad is a float in the reference: 0.02f)
    rad = float(np.float32(0.02))
    dm, _ = orc.shot352(model, orc.normals(model, k=10), kpm, rad)
    ds, _ = orc.shot352(scene, orc.normals(scene, k=10), kps, rad)
    oc = orc.match(dm, ds, 1, 0.25)
    assert len(oc) > 20
    # every entry that differs sits within the descriptor tolerance of the threshold or of a runner-up tie
    print("app correspondences:", eps.corr_check(dm, ds, corr, oc, 0.25, label="app correspondences"))
    # grouping on the app's own correspondences is bit-exact against the oracle
    oT, oinst = orc.gc_recognize(kpm, kps, corr, float(np.float32(0.02)), 2, max_inst=len(corr))
    assert len(oT) == len(T) == len(inst)
    for x, y in zip(inst, oinst):
        assert x.tobytes() == y.tobytes()
    if len(T):
        assert max(np.abs(A - B).max() for A, B in zip(T, oT)) < 1e-4


@pytest.mark.gpu
def test_fpfh_recognition_app_matches_oracle(apps, orc, synth, tmp_path):
    model = synth.make_model("y", 5000)
    scene = synth.make_scene(("y",), 30000, scene_id=4)
    kpm, kps = synth.voxel_grid(model, 0.01), synth.voxel_grid(scene, 0.02)
    _write(tmp_path / "mk.f32", kpm)
    _write(tmp_path / "sk.f32", kps)
    prefix = str(tmp_path / "out")
    r = subprocess.run([apps["fpfh_recognition"], str(tmp_path / "mk.f32"), str(tmp_path / "sk.f32"), prefix, "0.05",
                        "0.02", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    desc = np.fromfile(prefix + ".desc", dtype=np.float32).reshape(-1, 33)
    # stage by stage, each from identical inputs, each with an absolute bar and its boundary cases shown (tests/eps.py)
    nrm = np.fromfile(prefix + ".normals", dtype=np.float32).reshape(-1, 4)
    print("app normals:", eps.normals_check(orc, nrm, kps, radius=0.05, label="app normals"))
    print("app fpfh:", eps.fpfh_check(orc, desc, kps, nrm, 0.05, label="app fpfh"))
    corr = np.fromfile(prefix + ".corr", dtype=CORR)
    dm = orc.fpfh33(kpm, orc.normals(kpm, radius=0.05), 0.05)
    oc = orc.match(dm, desc, 2, 0.0)            # k = 2 ratio test on the app's own scene descriptors
    same = (corr["index_match"] == oc["index_match"]).all() and (corr["index_query"] == oc["index_query"]).mean() > 0.99
    assert len(corr) == len(oc) and same


@pytest.mark.gpu
def test_shot_recognition_app_extracts_keypoints(apps, orc, synth, tmp_path):
    """`us:<leaf>` keypoints: pcl::UniformSampling through the adapter, as SHOT.cpp:314-323 does."""
    model = synth.make_model("y", 5000)
    scene = synth.make_scene(("y",), 30000, scene_id=3)
    _write(tmp_path / "m.f32", model)
    _write(tmp_path / "s.f32", scene)
    prefix = str(tmp_path / "out")
    r = subprocess.run([apps["shot_recognition"], str(tmp_path / "m.f32"), "us:0.02", str(tmp_path / "s.f32"), "us:0.03",
                        prefix, "10", "0.02", "0.25", "0.02", "2", "batch"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    kpm, kps = orc.uniform_sampling(model, 0.02), orc.uniform_sampling(scene, 0.03)
    assert "Selected Keypoints: %d" % len(kpm) in r.stdout and "Selected Keypoints: %d" % len(kps) in r.stdout
    corr = np.fromfile(prefix + ".corr", dtype=CORR)
    rad = float(np.float32(0.02))
    dm, _ = orc.shot352(model, orc.normals(model, k=10), kpm, rad)
    ds, _ = orc.shot352(scene, orc.normals(scene, k=10), kps, rad)
    oc = orc.match(dm, ds, 1, 0.25)
    a = set(map(tuple, corr[["index_query", "index_match"]].tolist()))
    b = set(map(tuple, oc[["index_query", "index_match"]].tolist()))
    assert len(b) > 20 and len(a ^ b) <= max(2, 0.002 * len(b))


@pytest.mark.gpu
def test_shot_recognition_app_hough_branch(apps, orc, synth, tmp_path):
    """The reference's default grouping (SHOT.cpp:433-470) through the Hough3DGrouping adapter, frames = the SHOT
    frames of the descriptor stage; grouping on the app's own correspondences equals the restatement."""
    model = synth.make_model("y", 5000)
    scene = synth.make_scene(("y",), 30000, scene_id=3)
    kpm, kps = synth.uniform_sampling(model, 0.02), synth.uniform_sampling(scene, 0.03)
    for name, a in (("m", model), ("mk", kpm), ("s", scene), ("sk", kps)):
        _write(tmp_path / (name + ".f32"), a)
    prefix = str(tmp_path / "out")
    r = subprocess.run([apps["shot_recognition"]] + [str(tmp_path / (n + ".f32")) for n in ("m", "mk", "s", "sk")] +
                       [prefix, "10", "0.02", "0.25", "0.03", "3", "batch", "hough-shot"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    corr = np.fromfile(prefix + ".corr", dtype=CORR)
    T, inst = _read_instances(prefix)
    rad = float(np.float32(0.02))
    _, mrf = orc.shot352(model, orc.normals(model, k=10), kpm, rad)
    _, srf = orc.shot352(scene, orc.normals(scene, k=10), kps, rad)
    oT, oinst = orc.hough3d_recognize(kpm, mrf, kps, srf, corr, float(np.float32(0.03)), 3.0, max_inst=len(corr))
    # frames agree to ~1e-6, so a vote can change bin only on a bin boundary: allow a small mismatch
    assert abs(len(T) - len(oT)) <= max(1, len(oT) // 20)
    same = sum(1 for x, y in zip(inst, oinst) if x.tobytes() == y.tobytes())
    assert len(oT) == 0 or same >= 0.8 * min(len(inst), len(oinst))


@pytest.mark.gpu
def test_shot_recognition_app_icp_refinement(apps, orc, synth, tmp_path):
    """`icp:N`: the reference's icp_align (SHOT.cpp:177-192) through the IterativeClosestPoint adapter, on the model
    placed by each grouped pose; equals the restatement started from the app's own poses."""
    model = synth.make_model("y", 5000)
    scene = synth.make_scene(("y",), 30000, scene_id=3)
    kpm, kps = synth.uniform_sampling(model, 0.02), synth.uniform_sampling(scene, 0.03)
    for name, a in (("m", model), ("mk", kpm), ("s", scene), ("sk", kps)):
        _write(tmp_path / (name + ".f32"), a)
    prefix = str(tmp_path / "out")
    r = subprocess.run([apps["shot_recognition"]] + [str(tmp_path / (n + ".f32")) for n in ("m", "mk", "s", "sk")] +
                       [prefix, "10", "0.02", "0.25", "0.02", "2", "batch", "gc", "icp:4"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    T, _ = _read_instances(prefix)
    rec = np.fromfile(prefix + ".icp", dtype=np.float32).reshape(-1, 18)
    assert len(rec) == min(len(T), 8) and len(rec) >= 1
    assert "ICP has converged, score is" in r.stdout
    for i in range(min(len(rec), 3)):
        o = orc.icp_align(orc.transform_points(model, T[i]), scene, max_iterations=4)
        assert np.abs(rec[i, :16].reshape(4, 4) - o["final_transform"]).max() < 1e-4
        assert abs(rec[i, 16] - o["fitness"]) <= 1e-3 * o["fitness"] and rec[i, 17] == float(o["converged"])


@pytest.mark.gpu
def test_shot_recognition_app_hypothesis_verification(apps, orc, synth, tmp_path):
    """`hv:r`: the reference's hypothesis-verification block (SHOT_hypothesis.cpp:631-653) through the
    GlobalHypothesesVerification adapter on the ICP-registered instances; the mask equals the restatement run on the
    same instances (the app's refined poses applied to the model) with the reference's parameters in its call order
    (the occlusion threshold arrives after addModels: the constructor's 0.005 filters the models)."""
    model = synth.make_model("y", 8000)
    scene = synth.make_kinect_scene(("y", "diagonal", "horizontal"), 150000, scene_id=5)
    kpm, kps = synth.uniform_sampling(model, 0.015), synth.uniform_sampling(scene, 0.015)
    for name, a in (("m", model), ("mk", kpm), ("s", scene), ("sk", kps)):
        _write(tmp_path / (name + ".f32"), a)
    prefix = str(tmp_path / "out")
    r = subprocess.run([apps["shot_recognition"]] + [str(tmp_path / (n + ".f32")) for n in ("m", "mk", "s", "sk")] +
                       [prefix, "10", "0.03", "0.25", "0.02", "3", "batch", "gc", "icp:5", "hv:0.02"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rec = np.fromfile(prefix + ".icp", dtype=np.float32).reshape(-1, 18)
    T, _ = _read_instances(prefix)
    if len(rec) == 0:
        assert not os.path.exists(prefix + ".hv")
        pytest.skip("the grouping found no instance in this scene")
    assert "--- Hypotheses Verification ---" in r.stdout
    mask = np.fromfile(prefix + ".hv", dtype=np.uint8).astype(bool)
    assert len(mask) == len(rec)
    # the registered instance = the model placed by the grouped pose, then by ICP's final transformation
    inst = [orc.transform_points(orc.transform_points(model, T[i]), rec[i, :16].reshape(4, 4)) for i in range(len(rec))]
    p = orc.hv_params(detect_clutter=0, occlusion_reasoning=1, inlier_threshold=0.005, regularizer=0.001,
                      radius_clutter=0.003, clutter_regularizer=0.001, radius_normals=0.02)
    o = orc.hv_verify(scene, inst, p)
    print("hv app: %d instances, mask %s, restatement %s, visible %s" % (len(rec), mask.astype(int), o["mask"].astype(int),
                                                                         o["info"]["n_visible"]))
    assert mask.tolist() == o["mask"].tolist()
    for i, m in enumerate(mask):
        assert (("Instance %d is GOOD!" % i) in r.stdout) == bool(m)


@pytest.mark.gpu
def test_shot_recognition_app_hough_board_frames(apps, orc, synth, tmp_path):
    """The reference's Hough branch as written (SHOT.cpp:433-470): BOARD frames (find_holes, rf_rad 0.02) for model
    and scene keypoints through the BOARDLocalReferenceFrameEstimation adapter, then Hough3DGrouping.  Frames equal
    the restatement (the rand() stream runs on from the model call to the scene call); grouping on the app's own
    frames and correspondences equals the restatement."""
    model = synth.make_model("y", 5000)
    scene = synth.make_scene(("y",), 30000, scene_id=3)
    kpm, kps = synth.uniform_sampling(model, 0.02), synth.uniform_sampling(scene, 0.03)
    for name, a in (("m", model), ("mk", kpm), ("s", scene), ("sk", kps)):
        _write(tmp_path / (name + ".f32"), a)
    prefix = str(tmp_path / "out")
    r = subprocess.run([apps["shot_recognition"]] + [str(tmp_path / (n + ".f32")) for n in ("m", "mk", "s", "sk")] +
                       [prefix, "10", "0.02", "0.25", "0.03", "3", "batch", "hough"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr
    rf = np.fromfile(prefix + ".rf", dtype=np.float32).reshape(-1, 9)
    assert len(rf) == len(kpm) + len(kps)
    mrf, srf = rf[:len(kpm)], rf[len(kpm):]
    rad = float(np.float32(0.02))
    o_m, used = orc.board_lrf(model, orc.normals(model, k=10), kpm, rad)
    o_s, _ = orc.board_lrf(scene, orc.normals(scene, k=10), kps, rad, rand_skip=used)
    for a, b in ((mrf, o_m), (srf, o_s)):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        ok = ~np.isnan(b[:, 0])
        # the app's normals come from the device and differ from the restatement's in the last bit; BOARD's hole
        # direction divides differences of normal cosines by (1 - min cosine), which amplifies that to ~1e-4
        err = np.abs(a[ok] - b[ok]).max(axis=1)
        assert (err > 1e-5).mean() <= 0.15 and (err > 2e-3).mean() <= 0.005
    corr = np.fromfile(prefix + ".corr", dtype=CORR)
    T, inst = _read_instances(prefix)
    oT, oinst = orc.hough3d_recognize(kpm, mrf, kps, srf, corr, float(np.float32(0.03)), 3.0, max_inst=len(corr))
    assert len(T) == len(oT)
    for x, y in zip(inst, oinst):
        assert x.tobytes() == y.tobytes()
    if len(T):
        assert max(np.abs(A - B).max() for A, B in zip(T, oT)) < 1e-4


@pytest.mark.gpu
def test_batch_recognition_app_lanes(apps, b200, synth, tmp_path):
    """C++ host side of the lanes (b200_register_scene_batch_shot): six scenes, three lanes; every scene's
    correspondences and poses equal the scene registered alone."""
    model = synth.make_model("y", 20000)
    kpm = synth.uniform_sampling(model, 0.005)
    _write(tmp_path / "m.f32", model)
    _write(tmp_path / "mk.f32", kpm)
    args, scenes = [], []
    for s in range(6):
        sc = synth.make_scene(("y", "diagonal"), 60000, scene_id=30 + s % 2)
        kp = synth.uniform_sampling(sc, 0.02)
        _write(tmp_path / ("s%d.f32" % s), sc)
        _write(tmp_path / ("k%d.f32" % s), kp)
        args += [str(tmp_path / ("s%d.f32" % s)), str(tmp_path / ("k%d.f32" % s))]
        scenes.append((sc, kp))
    prefix = str(tmp_path / "out")
    r = subprocess.run([apps["batch_recognition"], str(tmp_path / "m.f32"), str(tmp_path / "mk.f32"), prefix, "3"] + args,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    p = b200.shot_params(normal_k=20, descr_radius=0.02, match_mode=1, match_thr=0.25, gc_size=0.02, gc_threshold=2,
                         max_instances=4096)
    ctx = b200.Context(0)
    m = ctx.model_create_shot(model, kpm, p)
    for s, (sc, kp) in enumerate(scenes):
        ref = ctx.register_scene_shot(m, sc, kp, p)
        corr = np.fromfile("%s.%d.corr" % (prefix, s), dtype=CORR)
        T = np.fromfile("%s.%d.T" % (prefix, s), dtype=np.float32).reshape(-1, 4, 4)
        assert "scene %d: %d correspondences, %d instances" % (s, len(ref["corrs"]), ref["n_instances"]) in r.stdout
        assert corr.tobytes() == ref["corrs"].tobytes() and np.array_equal(T, ref["transforms"])
    m.close()
    ctx.close()


@pytest.mark.gpu
def test_sharded_recognition_app(apps, b200, synth, tmp_path):
    """C++ host side of the multi-GPU path (b200_comm_* + b200_register_scene_shot_sharded, one host thread per GPU):
    the files it writes equal the single-GPU registration, with one rank and — on a box with two GPUs — with two."""
    import torch
    model = synth.make_model("y", 20000)
    kpm = synth.uniform_sampling(model, 0.005)
    scene = synth.make_scene(("y", "diagonal"), 120000, scene_id=31)
    kps = synth.uniform_sampling(scene, 0.015)
    for name, a in (("m", model), ("mk", kpm), ("s", scene), ("sk", kps)):
        _write(tmp_path / (name + ".f32"), a)
    p = b200.shot_params(normal_k=20, descr_radius=0.02, match_mode=1, match_thr=0.25, gc_size=0.02, gc_threshold=2,
                         max_instances=4096)
    ctx = b200.Context(0)
    m = ctx.model_create_shot(model, kpm, p)
    ref = ctx.register_scene_shot(m, scene, kps, p)
    m.close()
    ctx.close()
    assert ref["n_instances"] > 0
    for ranks in ([1, 2] if torch.cuda.device_count() >= 2 else [1]):
        prefix = str(tmp_path / ("out%d" % ranks))
        r = subprocess.run([apps["sharded_recognition"]] + [str(tmp_path / (n + ".f32")) for n in ("m", "mk", "s", "sk")] +
                           [prefix, str(ranks), "3"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        print(r.stdout.strip())
        corr = np.fromfile(prefix + ".corr", dtype=CORR)
        T = np.fromfile(prefix + ".T", dtype=np.float32).reshape(-1, 4, 4)
        assert corr.tobytes() == ref["corrs"].tobytes() and np.array_equal(T, ref["transforms"])
