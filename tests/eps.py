"""Epsilon statements for the parity tests: where the CUDA path and the CPU restatement may legitimately differ
(a value within rounding distance of a discrete decision), the tests do not accept a percentage of outliers — they
show that every differing output sits on such a boundary, and that everything else agrees at the north-star bar.

The boundaries are found with the oracle's perturbation variants (oracle/pcl_oracle.h): an output row that is
bit-identical under the +epsilon and the -epsilon variant cannot depend on any decision within epsilon of its
boundary ("stable"); rows that change are the boundary cases."""
import numpy as np

# north-star bars (BASELINE.json): descriptors 1e-4 L2 per descriptor, absolute
DESC_TOL = 1e-4
NORMAL_TOL = 1e-5


def _rows_equal(a, b):
    return np.all((a == b) | (np.isnan(a) & np.isnan(b)), axis=1)


def fpfh_check(orc, got, cloud, nrm, radius, q=None, label="fpfh"):
    """FPFH33: `got` (CUDA) against the restatement.  Only f1 goes through libm (atan2f, implementations differ by
    an ulp or two = up to 1e-6 of a bin): rows that no f1 value within 1e-6 of a bin border can reach must agree
    to 1e-4 ABSOLUTE L2; every other row may differ by at most what the border cases themselves move."""
    base = orc.fpfh33(cloud, nrm, radius, q=q)
    with orc.variant("fpfh_bin_up"):
        up = orc.fpfh33(cloud, nrm, radius, q=q)
    with orc.variant("fpfh_bin_down"):
        dn = orc.fpfh33(cloud, nrm, radius, q=q)
    assert np.array_equal(np.isnan(got), np.isnan(base)), label
    ok = ~np.isnan(base[:, 0])
    stable = ok & _rows_equal(base, up) & _rows_equal(base, dn)
    err = np.linalg.norm(got.astype(np.float64) - base, axis=1)
    assert stable.sum() > 0.5 * ok.sum(), (label, int(stable.sum()), int(ok.sum()))
    assert err[stable].max() < DESC_TOL, (label, "stable rows", float(err[stable].max()))
    unstable = ok & ~stable
    if unstable.any():
        reach = (np.linalg.norm(up.astype(np.float64) - base, axis=1) +
                 np.linalg.norm(dn.astype(np.float64) - base, axis=1))
        bad = unstable & (err > reach + DESC_TOL)
        assert not bad.any(), (label, "rows beyond what their border cases explain", int(bad.sum()),
                               float(err[bad].max()))
    return {"rows": int(ok.sum()), "stable": int(stable.sum()), "max_abs_l2_stable": float(err[stable].max()),
            "max_abs_l2_border": float(err[unstable].max()) if unstable.any() else 0.0,
            "rows_differing": int((err[ok] > DESC_TOL).sum())}


def normals_check(orc, got, cloud, k=0, radius=0.0, q=None, viewpoint=(0.0, 0.0, 0.0), label="normals"):
    """Normals: the float32 covariance sums are bit-identical on both sides; eigen33's closed-form roots go through
    atan2f / cosf / sinf.  Rows whose normal does not move by more than 1e-6 when theta is perturbed by +-2 ulp must
    agree to 1e-5; the others (ill-conditioned eigenvectors) at most by what the perturbation itself moves."""
    base = orc.normals(cloud, q=q, k=k, radius=radius, viewpoint=viewpoint)
    with orc.variant("root_up"):
        up = orc.normals(cloud, q=q, k=k, radius=radius, viewpoint=viewpoint)
    with orc.variant("root_down"):
        dn = orc.normals(cloud, q=q, k=k, radius=radius, viewpoint=viewpoint)
    assert np.array_equal(np.isnan(got), np.isnan(base)), label
    ok = ~np.isnan(base[:, 0])
    reach = np.maximum(np.abs(up - base).max(1), np.abs(dn - base).max(1))
    err = np.abs(got - base).max(1)
    stable = ok & (reach <= 1e-6)
    assert stable.sum() > 0.9 * ok.sum(), (label, int(stable.sum()), int(ok.sum()))
    assert err[stable].max() < NORMAL_TOL, (label, "stable rows", float(err[stable].max()))
    unstable = ok & ~stable
    if unstable.any():
        bad = unstable & (err > 4.0 * reach + NORMAL_TOL)
        assert not bad.any(), (label, "rows beyond what the root perturbation explains", int(bad.sum()),
                               float(err[bad].max()), float(reach[bad].min()))
    return {"rows": int(ok.sum()), "stable": int(stable.sum()), "max_err_stable": float(err[stable].max()),
            "max_err_border": float(err[unstable].max()) if unstable.any() else 0.0}


def corr_check(model_desc, scene_desc, got, ref, thr, delta=DESC_TOL, mode=1, label="correspondences"):
    """Correspondence lists computed from descriptors that differ by at most `delta` (L2) per row: every scene row
    whose entry differs between `got` and `ref` must have its best distance within eps of the threshold, or its
    runner-up within eps of the best, eps = 2 sqrt(d2) (2 delta) + (2 delta)^2 (both descriptors may move)."""
    def as_map(c):
        return {int(s): (int(m), float(d)) for m, s, d in zip(c["index_query"], c["index_match"], c["distance"])}
    g, r = as_map(got), as_map(ref)
    rows = sorted(s for s in set(g) | set(r) if g.get(s, (None,))[0] != r.get(s, (None,))[0])
    md = model_desc.astype(np.float64)
    valid = np.isfinite(md).all(1)
    worst = 0.0
    for s in rows:
        d2 = ((md[valid] - scene_desc[s].astype(np.float64)) ** 2).sum(1)
        o = np.sort(d2)[:2]
        eps = 2 * np.sqrt(o[0]) * (2 * delta) + (2 * delta) ** 2
        near_thr = mode == 1 and abs(o[0] - thr) <= eps
        near_tie = len(o) > 1 and (o[1] - o[0]) <= 2 * eps
        assert near_thr or near_tie, (label, "scene row %d differs away from any boundary" % s, float(o[0]),
                                      float(o[1]) if len(o) > 1 else None, float(eps))
        worst = max(worst, min(abs(o[0] - thr), (o[1] - o[0]) if len(o) > 1 else np.inf))
    return {"differing_rows": len(rows), "total": len(r), "largest_boundary_distance": worst}
