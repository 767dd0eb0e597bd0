"""Generates tests/golden/golden_small.npz from the CPU oracle (oracle/pcl_oracle.cpp).

PARITY UNPINNED: the reference repository holds no golden vectors and libpcl cannot be built or imported
here, so these fixtures are outputs of the restatement itself.  They guard the oracle
and the CUDA path against drift; they do not pin PCL.  Re-run only when the oracle is deliberately
changed:  python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pcl_oracle as orc  # noqa: E402

synth = importlib.import_module("3d-object-detection-of-industrial-joints_b200").synth


def main():
    model = synth.make_model("y", 4000)
    scene = synth.make_scene(("y",), 12000, scene_id=11)
    kpm = synth.uniform_sampling(model, 0.02)[::3]
    kps = synth.uniform_sampling(scene, 0.03)[::16]
    out = dict(model=model, scene=scene, model_kp=kpm, scene_kp=kps)
    q = scene[::97]
    off, idx, d2 = orc.radius_search(scene, q, 0.03)
    out.update(rs_q=q, rs_off=off, rs_idx=idx, rs_d2=d2)
    kidx, kd2, _ = orc.knn_search(scene, q, 10)
    out.update(knn_idx=kidx, knn_d2=kd2)
    nm, ns = orc.normals(model, k=10), orc.normals(scene, k=10)
    out.update(model_normals=nm, scene_normals=ns[::4], scene_normals_r=orc.normals(kps, radius=0.15))
    dm, rfm = orc.shot352(model, nm, kpm, 0.03)
    ds, rfs = orc.shot352(scene, ns, kps, 0.03)
    out.update(model_shot=dm, model_rf=rfm, scene_shot=ds, scene_rf=rfs)
    vg = synth.voxel_grid(scene, 0.03)
    nvg = orc.normals(vg, radius=0.1)
    out.update(fpfh_cloud=vg, fpfh_normals=nvg, fpfh=orc.fpfh33(vg, nvg, 0.1))
    c1 = orc.match(dm, ds, 1, 0.35)
    c2 = orc.match(dm, ds, 2, 0.0)
    out.update(corr_k1=c1, corr_k2=c2)
    T, inst = orc.gc_recognize(kpm, kps, c2, 0.02, 2, max_inst=256)
    out.update(gc_T=T, gc_sizes=np.array([len(i) for i in inst], np.int32),
               gc_corrs=np.concatenate(inst) if inst else np.zeros(0, orc.CORR_DTYPE))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(c1), "k=1 corrs,", len(c2), "k=2 corrs,", len(T), "instances")
    make_next(model, scene, kpm, kps, nm, ns, c2, T)


def make_next(model, scene, kpm, kps, nm, ns, corrs, gc_T):
    """Second fixture, same inputs: the SURVEY 8(f) rows (keypoint filters, BOARD frames, Hough grouping, ICP, cloud
    utilities).  Kept in its own file so that golden_small.npz stays byte-identical."""
    out = {}
    us, us_idx = orc.uniform_sampling(scene, 0.03, return_index=True)
    out.update(us_xyz=us, us_index=us_idx, vg_xyz=orc.voxel_grid(scene, 0.03))
    bm, used = orc.board_lrf(model, nm, kpm, 0.03)
    bs, used2 = orc.board_lrf(scene, ns, kps, 0.03, rand_skip=used)
    out.update(board_model=bm, board_scene=bs, board_rand_used=np.array([used, used2], np.int32))
    hT, hinst = orc.hough3d_recognize(kpm, bm, kps, bs, corrs, 0.08, 3.0, max_inst=256)
    out.update(hough_T=hT, hough_sizes=np.array([len(i) for i in hinst], np.int32),
               hough_corrs=np.concatenate(hinst) if hinst else np.zeros(0, orc.CORR_DTYPE))
    if len(gc_T):
        placed = orc.transform_points(model, gc_T[0])
        r = orc.icp_align(placed[::2], scene, max_iterations=5)
        out.update(icp_source=placed[::2], icp_T=r["final_transform"], icp_fitness=np.array([r["fitness"]]),
                   icp_iterations=np.array([r["iterations"], int(r["converged"])], np.int32))
    c = scene[:500].copy()
    c[::50] = np.nan
    kept, idx = orc.remove_nan(c)
    out.update(nan_cloud=c, nan_kept=kept, nan_index=idx)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_next.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(hT), "Hough instances,", used + used2, "rand() draws")


if __name__ == "__main__":
    main()
