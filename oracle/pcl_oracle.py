"""ctypes loader for the CPU parity oracle (TEST INFRASTRUCTURE ONLY — see pcl_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
PARITY UNPINNED: the oracle restates PCL 1.8.x semantics; the reference repo
holds no golden vectors for this path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Corr(C.Structure):
    _fields_ = [("index_query", C.c_int), ("index_match", C.c_int), ("distance", C.c_float)]


class HvParams(C.Structure):
    _fields_ = [("resolution", C.c_float), ("inlier_threshold", C.c_float), ("occlusion_threshold", C.c_float),
                ("regularizer", C.c_float), ("radius_normals", C.c_float), ("res_occupancy_grid", C.c_float),
                ("w_occupied_multiple_cm", C.c_float), ("initial_temp", C.c_float), ("max_iterations", C.c_int),
                ("occlusion_reasoning", C.c_int), ("zbuffer_scene_resolution", C.c_int),
                ("zbuffer_self_resolution", C.c_int), ("self_occlusion_threshold", C.c_float),
                ("detect_clutter", C.c_int), ("radius_clutter", C.c_float), ("clutter_regularizer", C.c_float),
                ("rand_seed", C.c_uint), ("mt_seed", C.c_uint), ("sa_uniform_mode", C.c_int)]


class HvInfo(C.Structure):
    _fields_ = [("valid", C.c_int), ("n_visible", C.c_int), ("n_points", C.c_int), ("n_outliers", C.c_int),
                ("n_explained", C.c_int), ("n_occupancy", C.c_int), ("outliers_weight", C.c_float),
                ("explained_sum", C.c_float)]


HV_INFO_DTYPE = np.dtype([("valid", "<i4"), ("n_visible", "<i4"), ("n_points", "<i4"), ("n_outliers", "<i4"),
                          ("n_explained", "<i4"), ("n_occupancy", "<i4"), ("outliers_weight", "<f4"),
                          ("explained_sum", "<f4")])


class BoardParams(C.Structure):
    _fields_ = [("find_holes", C.c_int), ("tangent_radius", C.c_float), ("margin_thresh", C.c_float),
                ("check_margin_array_size", C.c_int), ("hole_size_prob_thresh", C.c_float), ("steep_thresh", C.c_float)]


CORR_DTYPE = np.dtype([("index_query", "<i4"), ("index_match", "<i4"), ("distance", "<f4")])


def build(force=False):
    so = os.path.join(_HERE, "libpcl_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("pcl_oracle.cpp", "hv_oracle.cpp", "pcl_oracle.h", "Makefile")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int)
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_radius_search.restype = C.c_int64
        L.orc_radius_search.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_double,
                                        C.POINTER(C.c_int64), ip, fp, C.c_int64]
        L.orc_knn_search.restype = C.c_int
        L.orc_knn_search.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int, ip, fp]
        L.orc_normals.restype = C.c_int
        L.orc_normals.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_double, fp, fp]
        L.orc_shot_lrf.restype = C.c_int
        L.orc_shot_lrf.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_double, fp]
        L.orc_shot352.restype = C.c_int
        L.orc_shot352.argtypes = [fp, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_double, fp, fp]
        L.orc_fpfh33.restype = C.c_int
        L.orc_fpfh33.argtypes = [fp, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_double, fp]
        for f in (L.orc_match, L.orc_match_omp):
            f.restype = C.c_int
            f.argtypes = [fp, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(Corr)]
        L.orc_gc_recognize.restype = C.c_int
        L.orc_gc_recognize.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.POINTER(Corr), C.c_int,
                                       C.c_double, C.c_int, fp, C.c_int, ip, C.POINTER(Corr), C.c_int]
        L.orc_board_lrf.restype = C.c_int
        L.orc_board_lrf.argtypes = [fp, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_double, C.POINTER(BoardParams),
                                    C.c_uint, C.c_int, fp]
        L.orc_glibc_rand_nth.restype = C.c_int
        L.orc_glibc_rand_nth.argtypes = [C.c_uint, C.c_int]
        L.orc_remove_nan.restype = C.c_int
        L.orc_remove_nan.argtypes = [fp, C.c_int, C.c_int, fp, ip]
        L.orc_transform_points.restype = None
        L.orc_transform_points.argtypes = [fp, C.c_int, C.c_int, fp, fp]
        L.orc_icp_align.restype = C.c_int
        L.orc_icp_align.argtypes = [fp, C.c_int, C.c_int, fp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                    C.c_double, fp, fp, fp, C.POINTER(C.c_double), ip, ip]
        L.orc_hough3d_recognize.restype = C.c_int
        L.orc_hough3d_recognize.argtypes = [fp, fp, C.c_int, C.c_int, fp, fp, C.c_int, C.c_int, C.POINTER(Corr), C.c_int,
                                            C.c_double, C.c_double, fp, C.c_int, ip, C.POINTER(Corr), C.c_int]
        L.orc_uniform_sampling.restype = C.c_int
        L.orc_uniform_sampling.argtypes = [fp, C.c_int, C.c_int, C.c_double, fp, ip]
        L.orc_voxel_grid.restype = C.c_int
        L.orc_voxel_grid.argtypes = [fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, fp]
        L.orc_eigen33_smallest.argtypes = [fp, fp, fp]
        L.orc_eigh3_f64.argtypes = [C.POINTER(C.c_double)] * 3
        L.orc_umeyama3.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double)]
        L.orc_mt19937_nth.restype = C.c_uint32
        L.orc_mt19937_nth.argtypes = [C.c_uint32, C.c_int]
        L.orc_hv_params_default.restype = None
        L.orc_hv_params_default.argtypes = [C.POINTER(HvParams)]
        L.orc_hv_verify.restype = C.c_int
        L.orc_hv_verify.argtypes = [fp, C.c_int, C.c_int, fp, ip, C.c_int, C.c_int, C.POINTER(HvParams),
                                    C.POINTER(C.c_ubyte), C.c_void_p, C.POINTER(C.c_double), ip, ip, ip]
        L.orc_hv_optimize.restype = C.c_int
        L.orc_hv_optimize.argtypes = [C.c_int, C.c_int, ip, ip, fp, ip, ip, C.c_int, fp, ip, C.POINTER(HvParams),
                                      C.POINTER(C.c_ubyte), C.POINTER(C.c_double), ip]
        L.orc_hv_last_size.restype = C.c_int
        L.orc_hv_last_size.argtypes = [C.c_int]
        L.orc_hv_last_copy.restype = C.c_int
        L.orc_hv_last_copy.argtypes = [C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _pts(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] >= 3
    return a


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


VARIANTS = {"dot4": 1, "umeyama_f32": 2, "fpfh_skip": 4, "fpfh_bin_up": 8, "fpfh_bin_down": 16, "root_up": 32,
            "root_down": 64, "shot_bin_up": 128, "shot_bin_down": 256, "board_angle_up": 512, "board_angle_down": 1024,
            "hv_closest": 2048}


def set_variant(*names):
    """Selects sensitivity variants of the restatement (see pcl_oracle.h); no argument = the documented one."""
    mask = 0
    for n in names:
        mask |= VARIANTS[n]
    lib().orc_set_variant.argtypes = [C.c_uint]
    lib().orc_set_variant(mask)


class variant:
    """with orc.variant("dot4"): ...  — restores the documented restatement afterwards."""

    def __init__(self, *names):
        self.names = names

    def __enter__(self):
        set_variant(*self.names)

    def __exit__(self, *exc):
        set_variant()
        return False


def radius_search(surf, q, radius):
    surf, q = _pts(surf), _pts(q)
    off = np.zeros(len(q) + 1, dtype=np.int64)
    total = lib().orc_radius_search(_f(surf), len(surf), surf.shape[1], _f(q), len(q), q.shape[1], float(radius),
                                    off.ctypes.data_as(C.POINTER(C.c_int64)), None, None, 0)
    idx = np.zeros(max(total, 1), dtype=np.int32)
    d2 = np.zeros(max(total, 1), dtype=np.float32)
    lib().orc_radius_search(_f(surf), len(surf), surf.shape[1], _f(q), len(q), q.shape[1], float(radius),
                            off.ctypes.data_as(C.POINTER(C.c_int64)), _i(idx), _f(d2), total)
    return off, idx[:total], d2[:total]


def knn_search(surf, q, k):
    surf, q = _pts(surf), _pts(q)
    idx = np.zeros((len(q), k), dtype=np.int32)
    d2 = np.zeros((len(q), k), dtype=np.float32)
    kk = lib().orc_knn_search(_f(surf), len(surf), surf.shape[1], _f(q), len(q), q.shape[1], int(k), _i(idx), _f(d2))
    return idx, d2, kk


def normals(surf, q=None, k=0, radius=0.0, viewpoint=(0.0, 0.0, 0.0)):
    surf = _pts(surf)
    q = surf if q is None else _pts(q)
    out = np.zeros((len(q), 4), dtype=np.float32)
    vp = np.asarray(viewpoint, dtype=np.float32)
    rc = lib().orc_normals(_f(surf), len(surf), surf.shape[1], _f(q), len(q), q.shape[1], int(k), float(radius),
                           _f(vp), _f(out))
    if rc != 0:
        raise ValueError("orc_normals: exactly one of k / radius must be non-zero")
    return out


def shot_lrf(surf, kp, radius):
    surf, kp = _pts(surf), _pts(kp)
    out = np.zeros((len(kp), 9), dtype=np.float32)
    lib().orc_shot_lrf(_f(surf), len(surf), surf.shape[1], _f(kp), len(kp), kp.shape[1], float(radius), _f(out))
    return out


def shot352(surf, nrm, kp, radius):
    surf, kp = _pts(surf), _pts(kp)
    nrm = np.ascontiguousarray(nrm, dtype=np.float32)
    assert nrm.shape == (len(surf), 4)
    desc = np.zeros((len(kp), 352), dtype=np.float32)
    rf = np.zeros((len(kp), 9), dtype=np.float32)
    lib().orc_shot352(_f(surf), _f(nrm), len(surf), surf.shape[1], _f(kp), len(kp), kp.shape[1], float(radius),
                      _f(desc), _f(rf))
    return desc, rf


def fpfh33(surf, nrm, radius, q=None):
    surf = _pts(surf)
    nrm = np.ascontiguousarray(nrm, dtype=np.float32)
    assert nrm.shape == (len(surf), 4)
    if q is None:
        out = np.zeros((len(surf), 33), dtype=np.float32)
        lib().orc_fpfh33(_f(surf), _f(nrm), len(surf), surf.shape[1], None, 0, 0, float(radius), _f(out))
    else:
        q = _pts(q)
        out = np.zeros((len(q), 33), dtype=np.float32)
        lib().orc_fpfh33(_f(surf), _f(nrm), len(surf), surf.shape[1], _f(q), len(q), q.shape[1], float(radius),
                         _f(out))
    return out


def match(model, scene, mode=1, thr=0.25, omp=False):
    model = np.ascontiguousarray(model, dtype=np.float32)
    scene = np.ascontiguousarray(scene, dtype=np.float32)
    assert model.shape[1] == scene.shape[1]
    out = np.zeros(max(len(scene), 1), dtype=CORR_DTYPE)
    fn = lib().orc_match_omp if omp else lib().orc_match
    c = fn(_f(model), len(model), _f(scene), len(scene), model.shape[1], int(mode), float(thr),
           out.ctypes.data_as(C.POINTER(Corr)))
    return out[:c].copy()


def gc_recognize(model_kp, scene_kp, corrs, gc_size, gc_threshold, max_inst=256):
    model_kp, scene_kp = _pts(model_kp), _pts(scene_kp)
    corrs = np.ascontiguousarray(corrs, dtype=CORR_DTYPE)
    T = np.zeros((max_inst, 16), dtype=np.float32)
    off = np.zeros(max_inst + 1, dtype=np.int32)
    cap = max(len(corrs), 1) * 2
    oc = np.zeros(cap, dtype=CORR_DTYPE)
    n = lib().orc_gc_recognize(_f(model_kp), len(model_kp), model_kp.shape[1], _f(scene_kp), len(scene_kp),
                               scene_kp.shape[1], corrs.ctypes.data_as(C.POINTER(Corr)), len(corrs),
                               float(gc_size), int(gc_threshold), _f(T), max_inst, _i(off),
                               oc.ctypes.data_as(C.POINTER(Corr)), cap)
    n = min(n, max_inst)
    return T[:n].reshape(n, 4, 4).copy(), [oc[off[i]:off[i + 1]].copy() for i in range(n)]


def hough3d_recognize(model_kp, model_rf, scene_kp, scene_rf, corrs, bin_size, threshold, max_inst=256):
    model_kp, scene_kp = _pts(model_kp), _pts(scene_kp)
    model_rf = np.ascontiguousarray(model_rf, dtype=np.float32).reshape(len(model_kp), 9)
    scene_rf = np.ascontiguousarray(scene_rf, dtype=np.float32).reshape(len(scene_kp), 9)
    corrs = np.ascontiguousarray(corrs, dtype=CORR_DTYPE)
    T = np.zeros((max_inst, 16), dtype=np.float32)
    off = np.zeros(max_inst + 1, dtype=np.int32)
    cap = max(len(corrs), 1) * 2
    oc = np.zeros(cap, dtype=CORR_DTYPE)
    n = lib().orc_hough3d_recognize(_f(model_kp), _f(model_rf), len(model_kp), model_kp.shape[1], _f(scene_kp),
                                    _f(scene_rf), len(scene_kp), scene_kp.shape[1],
                                    corrs.ctypes.data_as(C.POINTER(Corr)), len(corrs), float(bin_size), float(threshold),
                                    _f(T), max_inst, _i(off), oc.ctypes.data_as(C.POINTER(Corr)), cap)
    n = min(max(n, 0), max_inst)
    return T[:n].reshape(n, 4, 4).copy(), [oc[off[i]:off[i + 1]].copy() for i in range(n)]


def board_params(find_holes=True, tangent_radius=0.0, margin_thresh=0.85, check_margin_array_size=24,
                 hole_size_prob_thresh=0.2, steep_thresh=0.1):
    """PCL's constructor defaults, with find_holes as the reference sets it (SHOT.cpp:442)."""
    return BoardParams(int(find_holes), tangent_radius, margin_thresh, check_margin_array_size, hole_size_prob_thresh,
                       steep_thresh)


def board_lrf(surf, normals, kp, radius, params=None, rand_seed=1, rand_skip=0):
    """BOARDLocalReferenceFrameEstimation::compute.  Returns (frames K x 9, rand() values consumed)."""
    surf, kp = _pts(surf), _pts(kp)
    normals = np.ascontiguousarray(normals, dtype=np.float32).reshape(len(surf), 4)
    params = params or board_params()
    out = np.zeros((max(len(kp), 1), 9), dtype=np.float32)
    used = lib().orc_board_lrf(_f(surf), _f(normals), len(surf), surf.shape[1], _f(kp), len(kp), kp.shape[1],
                               float(radius), C.byref(params), int(rand_seed), int(rand_skip), _f(out))
    return out[:len(kp)], used


def glibc_rand_nth(seed, nth):
    return lib().orc_glibc_rand_nth(int(seed), int(nth))


def remove_nan(xyz):
    xyz = _pts(xyz)
    out = np.zeros((max(len(xyz), 1), 3), dtype=np.float32)
    idx = np.zeros(max(len(xyz), 1), dtype=np.int32)
    n = lib().orc_remove_nan(_f(xyz), len(xyz), xyz.shape[1], _f(out), _i(idx))
    return out[:n].copy(), idx[:n].copy()


def transform_points(xyz, transform):
    xyz = _pts(xyz)
    T = np.ascontiguousarray(transform, dtype=np.float32).reshape(16)
    out = np.zeros((max(len(xyz), 1), 3), dtype=np.float32)
    lib().orc_transform_points(_f(xyz), len(xyz), xyz.shape[1], _f(T), _f(out))
    return out[:len(xyz)]


def icp_align(source, target, max_iterations=10, max_corr_dist=0.0, transformation_epsilon=0.0,
              euclidean_fitness_epsilon=-1.7976931348623157e308, guess=None):
    source, target = _pts(source), _pts(target)
    T = np.zeros(16, dtype=np.float32)
    al = np.zeros((max(len(source), 1), 3), dtype=np.float32)
    g = None if guess is None else np.ascontiguousarray(guess, dtype=np.float32).reshape(16)
    fit, conv, it = C.c_double(), C.c_int(), C.c_int()
    lib().orc_icp_align(_f(source), len(source), source.shape[1], _f(target), len(target), target.shape[1],
                        int(max_iterations), float(max_corr_dist), float(transformation_epsilon),
                        float(euclidean_fitness_epsilon), None if g is None else _f(g), _f(T), _f(al), C.byref(fit),
                        C.byref(conv), C.byref(it))
    return {"final_transform": T.reshape(4, 4), "aligned": al[:len(source)], "fitness": fit.value,
            "converged": bool(conv.value), "iterations": it.value}


def uniform_sampling(xyz, leaf, return_index=False):
    xyz = _pts(xyz)
    out = np.zeros((max(len(xyz), 1), 3), dtype=np.float32)
    idx = np.zeros(max(len(xyz), 1), dtype=np.int32)
    n = lib().orc_uniform_sampling(_f(xyz), len(xyz), xyz.shape[1], float(leaf), _f(out), _i(idx))
    if n < 0:
        raise ValueError("leaf size is too small for the input dataset")
    return (out[:n].copy(), idx[:n].copy()) if return_index else out[:n].copy()


def voxel_grid(xyz, leaf):
    xyz = _pts(xyz)
    lx, ly, lz = (leaf, leaf, leaf) if np.isscalar(leaf) else leaf
    out = np.zeros((max(len(xyz), 1), 3), dtype=np.float32)
    n = lib().orc_voxel_grid(_f(xyz), len(xyz), xyz.shape[1], float(lx), float(ly), float(lz), _f(out))
    if n < 0:
        raise ValueError("leaf size is too small for the input dataset")
    return out[:n].copy()


def eigen33_smallest(cov):
    cov = np.ascontiguousarray(cov, dtype=np.float32).reshape(9)
    ev = np.zeros(1, dtype=np.float32)
    vec = np.zeros(3, dtype=np.float32)
    lib().orc_eigen33_smallest(_f(cov), _f(ev), _f(vec))
    return float(ev[0]), vec


def eigh3(a):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(9)
    w = np.zeros(3)
    v = np.zeros(9)
    dp = C.POINTER(C.c_double)
    lib().orc_eigh3_f64(a.ctypes.data_as(dp), w.ctypes.data_as(dp), v.ctypes.data_as(dp))
    return w, v.reshape(3, 3)


def umeyama3(src, dst):
    src = np.ascontiguousarray(src, dtype=np.float64)
    dst = np.ascontiguousarray(dst, dtype=np.float64)
    T = np.zeros(16)
    dp = C.POINTER(C.c_double)
    lib().orc_umeyama3(src.ctypes.data_as(dp), dst.ctypes.data_as(dp), len(src), T.ctypes.data_as(dp))
    return T.reshape(4, 4)


def mt19937_nth(seed, nth):
    return int(lib().orc_mt19937_nth(seed, nth))


def hv_params(**kw):
    """GlobalHypothesesVerification parameters: PCL's constructor defaults, overridden by keyword."""
    p = HvParams()
    lib().orc_hv_params_default(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown hypothesis-verification parameter %r" % k)
        setattr(p, k, v)
    return p


def _hv_models(models):
    models = [_pts(m) for m in models]
    offs = np.zeros(len(models) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(m) for m in models])
    flat = np.ascontiguousarray(np.concatenate([m[:, :3] for m in models]) if models else np.zeros((0, 3), np.float32),
                                dtype=np.float32)
    if len(flat) == 0:
        flat = np.zeros((1, 3), np.float32)
    return flat, offs


def hv_verify(scene, models, params):
    """GlobalHypothesesVerification (SHOT_hypothesis.cpp:631-653).  Returns a dict: mask, info, best_cost,
    accepted_moves, n_scene_points, n_cells and the cue lists (expl_off / expl_idx / expl_w / occ_off / occ_idx,
    CSR over the valid hypotheses)."""
    scene = _pts(scene)
    flat, offs = _hv_models(models)
    H = len(models)
    mask = np.zeros(max(H, 1), dtype=np.uint8)
    info = np.zeros(max(H, 1), dtype=HV_INFO_DTYPE)
    cost = C.c_double(0.0)
    acc, ns, nc = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = lib().orc_hv_verify(_f(scene), len(scene), scene.shape[1], _f(flat), _i(offs), H, 3, C.byref(params),
                             mask.ctypes.data_as(C.POINTER(C.c_ubyte)), info.ctypes.data, C.byref(cost), C.byref(acc),
                             C.byref(ns), C.byref(nc))
    if rc != 0:
        raise ValueError("orc_hv_verify failed: %d" % rc)
    out = {"mask": mask[:H].astype(bool), "info": info[:H], "best_cost": cost.value, "accepted_moves": acc.value,
           "n_scene_points": ns.value, "n_cells": nc.value}
    for which, (name, dt) in enumerate((("expl_off", np.int32), ("expl_idx", np.int32), ("expl_w", np.float32),
                                        ("occ_off", np.int32), ("occ_idx", np.int32))):
        a = np.zeros(max(lib().orc_hv_last_size(which), 1), dtype=dt)
        n = lib().orc_hv_last_size(which)
        lib().orc_hv_last_copy(which, a.ctypes.data)
        out[name] = a[:n]
    return out


def hv_optimize(ns, expl_off, expl_idx, expl_w, occ_off, occ_idx, n_cells, outliers_weight, bad_information, params):
    """SAOptimize on given cue lists.  Returns (mask, best_cost, accepted_moves)."""
    H = len(expl_off) - 1
    eo = np.ascontiguousarray(expl_off, dtype=np.int32)
    ei = np.ascontiguousarray(np.append(expl_idx, 0), dtype=np.int32)
    ew = np.ascontiguousarray(np.append(expl_w, 0), dtype=np.float32)
    oo = np.ascontiguousarray(occ_off, dtype=np.int32)
    oi = np.ascontiguousarray(np.append(occ_idx, 0), dtype=np.int32)
    ow = np.ascontiguousarray(outliers_weight, dtype=np.float32)
    bi = np.ascontiguousarray(bad_information, dtype=np.int32)
    mask = np.zeros(max(H, 1), dtype=np.uint8)
    cost = C.c_double(0.0)
    acc = C.c_int(0)
    lib().orc_hv_optimize(H, int(ns), _i(eo), _i(ei), _f(ew), _i(oo), _i(oi), int(n_cells), _f(ow), _i(bi),
                          C.byref(params), mask.ctypes.data_as(C.POINTER(C.c_ubyte)), C.byref(cost), C.byref(acc))
    return mask[:H].astype(bool), cost.value, acc.value
