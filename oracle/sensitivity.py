#!/usr/bin/env python
"""Sensitivity of the CPU restatement (oracle/pcl_oracle.cpp) to the details of upstream PCL that could not be
pinned (SURVEY.md Appendix A "[?]" items; PCL is absent here and on the GPU box, profiles/pcl_probe_r02.txt).

For BASELINE config 1 (5 k-pt model vs 100 k-pt scene) and for the north-star target scene (1 M points, reduced
keypoint set unless --full) the pipeline is run with the documented restatement and with each alternative; the
table reports how many outputs move past the north-star bars (descriptors 1e-4 L2, correspondence lists exact,
poses 1e-4 m / 0.01 deg).  An alternative that moves nothing is closed; one that does is a named parity risk.

    python oracle/sensitivity.py > profiles/sensitivity_r02.md        (test infrastructure; CPU only)
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pcl_oracle as orc  # noqa: E402

synth = importlib.import_module("3d-object-detection-of-industrial-joints_b200").synth


def pose_delta(Ta, Tb):
    """(max translation difference [m], max rotation difference [deg]) over paired 4x4 transforms; the angle from
    the chord |Ra - Rb|_F = 2 sqrt(2) sin(angle / 2), which stays accurate for tiny angles (arccos of the trace
    does not: float32 matrices alone put its floor at 0.02 deg)."""
    if len(Ta) == 0:
        return 0.0, 0.0
    Ta, Tb = np.asarray(Ta, np.float64).reshape(-1, 4, 4), np.asarray(Tb, np.float64).reshape(-1, 4, 4)
    dt = np.linalg.norm(Ta[:, :3, 3] - Tb[:, :3, 3], axis=1).max()
    chord = np.linalg.norm((Ta[:, :3, :3] - Tb[:, :3, :3]).reshape(-1, 9), axis=1)
    ang = np.degrees(2 * np.arcsin(np.clip(chord / (2 * np.sqrt(2)), 0, 1))).max()
    return float(dt), float(ang)


def pipeline(model, scene, kpm, kps, k, radius, thr, gc_size, gc_thr, reuse=None, stages=("normals", "shot", "match", "gc")):
    """SHOT pipeline; `reuse` = outputs of the documented run for the stages a variant cannot influence."""
    out = dict(reuse or {})
    if "normals" in stages:
        out["nm"], out["ns"] = orc.normals(model, k=k), orc.normals(scene, k=k)
    if "shot" in stages:
        out["dm"], out["fm"] = orc.shot352(model, out["nm"], kpm, radius)
        out["ds"], out["fs"] = orc.shot352(scene, out["ns"], kps, radius)
    if "match" in stages:
        out["corr"] = orc.match(out["dm"], out["ds"], 1, thr, omp=True)
    if "gc" in stages:
        out["T"], out["inst"] = orc.gc_recognize(kpm, kps, out["corr"], gc_size, gc_thr, max_inst=8192)
    return out


def compare(base, var):
    def moved(a, b, tol):
        ok = np.isfinite(a).all(1) & np.isfinite(b).all(1)
        nanflip = int((np.isfinite(a).all(1) != np.isfinite(b).all(1)).sum())
        d = np.linalg.norm(a[ok].astype(np.float64) - b[ok], axis=1)
        return int((d > tol).sum()) + nanflip, float(d.max()) if len(d) else 0.0
    r = {}
    r["normals>1e-5"], r["normals max"] = moved(np.concatenate([base["nm"], base["ns"]])[:, :3],
                                                np.concatenate([var["nm"], var["ns"]])[:, :3], 1e-5)
    r["SHOT>1e-4"], r["SHOT max L2"] = moved(np.concatenate([base["dm"], base["ds"]]),
                                             np.concatenate([var["dm"], var["ds"]]), 1e-4)
    pa = set(map(tuple, base["corr"][["index_query", "index_match"]].tolist()))
    pb = set(map(tuple, var["corr"][["index_query", "index_match"]].tolist()))
    r["corr sym.diff"] = len(pa ^ pb)
    r["corr total"] = len(pa)
    r["instances"] = "%d -> %d" % (len(base["T"]), len(var["T"]))
    def pairs(i):
        return np.stack([i["index_query"], i["index_match"]], 1).tobytes()
    same = len(base["T"]) == len(var["T"]) and all(pairs(a) == pairs(b) for a, b in zip(base["inst"], var["inst"]))
    r["instance lists identical"] = bool(same)
    if len(base["T"]) == len(var["T"]):
        r["pose dt [m]"], r["pose dR [deg]"] = pose_delta(base["T"], var["T"])
    else:
        r["pose dt [m]"], r["pose dR [deg]"] = float("nan"), float("nan")
    return r


VARIANT_STAGES = {   # the first stage a variant can influence
    "dot4": ("shot", "match", "gc"),
    "umeyama_f32": ("gc",),
    "root_up": ("normals", "shot", "match", "gc"),
    "root_down": ("normals", "shot", "match", "gc"),
    "shot_bin_up": ("shot", "match", "gc"),
    "shot_bin_down": ("shot", "match", "gc"),
}
WHAT = {
    "dot4": "Eigen 4-lane dot order in createBinDistanceShape / LRF votes: (x x'+z z')+(y y'+0) instead of sequential",
    "umeyama_f32": "RANSAC rigid fit from float32 moments (Matrix4f umeyama) instead of float64",
    "root_up": "eigen33 closed-form roots: theta * (1 + 2^-21) (libm ulp differences in atan2f/cosf/sinf)",
    "root_down": "eigen33 closed-form roots: theta * (1 - 2^-21)",
    "shot_bin_up": "SHOT cosine-bin coordinate + 1e-6 before floor(x + 0.5)",
    "shot_bin_down": "SHOT cosine-bin coordinate - 1e-6 before floor(x + 0.5)",
}


READING = '''
### Reading

* `dot4` (Eigen lane order of the float dot products feeding `floor(binDistance + 0.5)` and the LRF sign votes): no
  descriptor moves by more than 4e-7 L2, no correspondence and no instance list changes on either workload — closed.
* `umeyama_f32` (float32 vs float64 rigid fit in the RANSAC stage): instance lists identical, poses move by at most
  3.4e-6 m / 0.0012 deg, a factor 30 / 8 inside the north-star bars (1e-4 m, 0.01 deg) — closed.
* `fpfh_skip` (degenerate pairs vote with f = 0 or not at all): bit-identical on these clouds (no coincident points, no
  displacement parallel to a normal); it can only matter for clouds with duplicated points — open for such data, closed here.
* `root_*`, `shot_bin_*`, `fpfh_bin_*` are not alternatives but +-epsilon perturbations of discrete decisions (libm ulp
  differences): 1 normal in 1 M, 4-6 SHOT descriptors in 28 k and 87 FPFH rows in 61 k sit within epsilon of a decision
  boundary and then move by 0.02-0.3 L2 (one vote changes bin).  These are the only rows for which the CUDA path and any
  CPU evaluation (PCL included) can legitimately differ by more than 1e-4; the GPU tests identify them with the same
  perturbations (tests/eps.py) and hold every other row to the absolute bar.  None of them changes a correspondence or
  an instance on these workloads.
* What stays unpinned: anything a real PCL build would do differently from Appendix A in ways not listed here
  (e.g. a different neighbour order for equidistant points, a different RANSAC sample stream).  The table bounds the
  listed items only.
'''


def run_case(title, model, scene, kpm, kps, k, radius, thr, gc_size, gc_thr):
    print("\n### %s\n" % title)
    print("model %d pts / %d keypoints, scene %d pts / %d keypoints, normals k=%d, SHOT r=%g, d2<%g, GC %g/%d\n"
          % (len(model), len(kpm), len(scene), len(kps), k, radius, thr, gc_size, gc_thr))
    t0 = time.time()
    base = pipeline(model, scene, kpm, kps, k, radius, thr, gc_size, gc_thr)
    print("documented restatement: %d correspondences, %d instances (%.0f s)\n" % (len(base["corr"]), len(base["T"]),
                                                                                 time.time() - t0))
    cols = ["normals>1e-5", "SHOT>1e-4", "SHOT max L2", "corr sym.diff", "instances", "instance lists identical",
            "pose dt [m]", "pose dR [deg]"]
    print("| variant | " + " | ".join(cols) + " |")
    print("|---|" + "---|" * len(cols))
    for name, stages in VARIANT_STAGES.items():
        with orc.variant(name):
            var = pipeline(model, scene, kpm, kps, k, radius, thr, gc_size, gc_thr, reuse=base, stages=stages)
        r = compare(base, var)
        print("| %s | " % name + " | ".join(("%.3g" % r[c]) if isinstance(r[c], float) else str(r[c]) for c in cols) + " |")
        sys.stdout.flush()


def run_fpfh(model, scene):
    print("\n### FPFH33 (BASELINE config 2: voxel grid 0.01, radius normals and FPFH r = 0.05)\n")
    km, ks = synth.voxel_grid(model, 0.01), synth.voxel_grid(scene, 0.01)
    nm, ns = orc.normals(km, radius=0.05), orc.normals(ks, radius=0.05)
    base = np.concatenate([orc.fpfh33(km, nm, 0.05), orc.fpfh33(ks, ns, 0.05)])
    print("%d + %d descriptors; rows are L2 distances to the documented restatement (histogram blocks sum to 100)\n"
          % (len(km), len(ks)))
    print("| variant | rows > 1e-4 | rows > 1e-2 | max L2 |")
    print("|---|---|---|---|")
    for name, what in (("fpfh_skip", "degenerate pairs do not vote"),
                       ("fpfh_bin_up", "f1 (atan2f) bin coordinate + 1e-6"),
                       ("fpfh_bin_down", "f1 (atan2f) bin coordinate - 1e-6")):
        with orc.variant(name):
            v = np.concatenate([orc.fpfh33(km, nm, 0.05), orc.fpfh33(ks, ns, 0.05)])
        ok = np.isfinite(base[:, 0]) & np.isfinite(v[:, 0])
        d = np.linalg.norm(base[ok].astype(np.float64) - v[ok], axis=1)
        print("| %s (%s) | %d | %d | %.3g |" % (name, what, (d > 1e-4).sum(), (d > 1e-2).sum(), d.max()))


def main():
    full = "--full" in sys.argv
    orc.set_num_threads(os.cpu_count() or 1)
    print("# Sensitivity of the CPU restatement to unpinned PCL details (round 2)\n")
    print("Generated by `python oracle/sensitivity.py%s` on %d host threads.  PCL itself is not available "
          "(profiles/pcl_probe_r02.txt), so each Appendix-A \"[?]\" alternative is applied to the restatement and the "
          "outputs are compared with the documented choice at the north-star bars.\n" % (" --full" if full else "",
                                                                                       orc.num_threads()))
    print("Variants:\n")
    for k, v in WHAT.items():
        print("* `%s` — %s" % (k, v))
    model = synth.make_model("y", 5000)
    scene = synth.make_scene(("y",), 100000, scene_id=1)
    run_case("BASELINE config 1 (SHOT_demo parameters)", model, scene, synth.voxel_grid(model, 0.02),
             synth.voxel_grid(scene, 0.03), 10, 0.02, 0.25, 0.02, 2)
    run_fpfh(model, scene)
    model3 = synth.make_model("y", 50000)
    scene3 = synth.make_kinect_scene(("y", "diagonal", "horizontal"), target_points=1_000_000, scene_id=0)
    kpm3 = synth.uniform_sampling(model3, 0.005)
    kps3 = synth.uniform_sampling(scene3, 0.01)
    if not full:
        kps3 = np.ascontiguousarray(kps3[::6])
    run_case("north-star target scene (BASELINE config 3 parameters)%s" % ("" if full else ", every 6th scene keypoint"),
             model3, scene3, kpm3, kps3, 20, 0.02, 0.25, 0.02, 2)
    print(READING)


if __name__ == "__main__":
    main()
