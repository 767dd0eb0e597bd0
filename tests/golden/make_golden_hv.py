"""Generates tests/golden/golden_hv.npz from the CPU restatement of the hypothesis verification
(oracle/hv_oracle.cpp; GlobalHypothesesVerification, SHOT_hypothesis.cpp:631-653).

PARITY UNPINNED, like the other fixtures (make_golden.py): outputs of the restatement itself, a guard against drift
of the restatement and of the CUDA path, not a pin on PCL.  Re-run only when the restatement is deliberately changed:
    python tests/golden/make_golden_hv.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
from oracle import pcl_oracle as orc  # noqa: E402
import hv_cases  # noqa: E402

synth = importlib.import_module("3d-object-detection-of-industrial-joints_b200").synth

VERIFY_KW = dict(detect_clutter=0, occlusion_reasoning=1, regularizer=3.0, radius_normals=0.03)


def verify_case():
    scene, hyps, _ = hv_cases.kinect(synth, 60000, seed=2)
    return scene, [h[::4] for h in hyps[:6]] + [np.zeros((0, 3), np.float32)]


def main():
    out = {}
    # the annealing on fixed cue lists, both acceptance variates
    cues = hv_cases.random_cues(21, H=14, ns=4000, n_cells=3000)
    names = ("ns", "expl_off", "expl_idx", "expl_w", "occ_off", "occ_idx", "n_cells", "outliers_weight", "bad_information")
    for k, v in zip(names, cues):
        out["cue_" + k] = np.asarray(v)
    for mode in (0, 1):
        m, c, a = orc.hv_optimize(*cues, orc.hv_params(detect_clutter=0, sa_uniform_mode=mode))
        out["anneal_mask_%d" % mode] = m
        out["anneal_cost_accepted_%d" % mode] = np.array([c, a], dtype=np.float64)
    # a whole verification (the scene and the hypotheses are regenerated from their seeds by the test)
    scene, hyps = verify_case()
    r = orc.hv_verify(scene, hyps, orc.hv_params(**VERIFY_KW))
    out.update(hv_mask=r["mask"], hv_info=r["info"], hv_cost_accepted=np.array([r["best_cost"], r["accepted_moves"]]),
               hv_sizes=np.array([r["n_scene_points"], r["n_cells"]]), hv_expl_off=r["expl_off"], hv_expl_idx=r["expl_idx"],
               hv_expl_w=r["expl_w"], hv_occ_off=r["occ_off"], hv_occ_sorted=np.concatenate(
                   [np.sort(r["occ_idx"][a:b]) for a, b in zip(r["occ_off"][:-1], r["occ_off"][1:])] or [np.zeros(0, np.int32)]))
    path = os.path.join(HERE, "golden_hv.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; mask", r["mask"].astype(int), "info", r["info"]["n_explained"])


if __name__ == "__main__":
    main()
