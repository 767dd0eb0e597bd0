"""Tensor-core (tcgen05) correspondence pre-filter: results must be bit-identical to the exact float32
kernel and to the oracle, whatever the filter certifies or hands back to the exact kernel."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


@pytest.fixture(scope="module")
def ctx(b200):
    c = b200.Context(0)
    c.set_profiling(True)
    yield c
    c.close()


def _with_mode(mode, fn):
    old = os.environ.get("B200_MATCH")
    os.environ["B200_MATCH"] = mode
    try:
        return fn()
    finally:
        if old is None:
            del os.environ["B200_MATCH"]
        else:
            os.environ["B200_MATCH"] = old


def _shot_like(rng, n, D=352, clusters=None):
    """non-negative, unit-norm rows; optionally near-duplicates of cluster centres (hard case)."""
    x = rng.gamma(0.3, 1.0, (n, D)).astype(np.float32)
    if clusters is not None:
        x = clusters[rng.integers(0, len(clusters), n)] + 0.05 * x
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x.astype(np.float32)


@pytest.mark.parametrize("terms", ["tc13", "tc3", "tc1"])
def test_tc_filter_matches_exact(ctx, orc, terms):
    rng = _rng(31)
    centres = _shot_like(rng, 300)
    for Km, Ks, D in ((3000, 2500, 352), (700, 300, 352), (1500, 1000, 33), (257, 129, 352)):
        cc = centres[:, :D] if D < 352 else centres
        model = _shot_like(rng, Km, D, cc)
        scene = _shot_like(rng, Ks, D, cc)
        scene[3, 0] = np.nan                    # skipped scene row
        model[4, D - 1] = np.inf                # dropped model row
        scene[5] = model[2]                     # exact duplicate → distance 0
        model[9] = model[8]                     # duplicated model rows → arg-min tie → lower index wins
        scene[6] = model[8]
        for mode, thr in ((1, 0.25), (1, 0.02), (2, 0.0)):
            ref = _with_mode("exact", lambda: ctx.match(model, scene, mode, thr))
            got = _with_mode(terms, lambda: ctx.match(model, scene, mode, thr))
            fb = ctx.match_fallback_rows()
            assert got.tobytes() == ref.tobytes(), (terms, Km, Ks, D, mode)
            assert 0 <= fb <= Ks
            # the approximation error actually observed must sit well inside the certificate's bound
            assert ctx.match_error_ratio() < 0.5, ctx.match_error_ratio()
        oref = orc.match(model, scene, 1, 0.25, omp=True)
        got = _with_mode(terms, lambda: ctx.match(model, scene, 1, 0.25))
        assert got.tobytes() == oref.tobytes()


def test_tc_filter_certifies_most_rows(ctx, orc, synth):
    """On real SHOT descriptors the three-term filter should certify nearly every row."""
    model = synth.make_model("y", 20000)
    scene = synth.make_scene(("y",), 60000, scene_id=7)
    kpm, kps = synth.uniform_sampling(model, 0.006), synth.voxel_grid(scene, 0.02)
    dm, _ = orc.shot352(model, orc.normals(model, k=10), kpm, 0.02)
    ds, _ = orc.shot352(scene, orc.normals(scene, k=10), kps, 0.02)
    ref = _with_mode("exact", lambda: ctx.match(dm, ds, 1, 0.25))
    for terms, max_frac in (("tc13", 0.05), ("tc3", 0.05), ("tc1", 0.6)):
        got = _with_mode(terms, lambda: ctx.match(dm, ds, 1, 0.25))
        fb = ctx.match_fallback_rows()
        print("%s: %d of %d rows fell back to the exact kernel (Km=%d); error/bound = %.3f; first pass left %d" %
              (terms, fb, len(ds), len(dm), ctx.match_error_ratio(), ctx.match_pass1_rows()))
        assert ctx.match_error_ratio() < 0.5
        assert got.tobytes() == ref.tobytes()
        assert fb <= max_frac * len(ds)


def test_tc_approximation_error_within_bound(ctx):
    """Adversarial magnitudes: large dynamic range and non-unit norms must still give exact results."""
    rng = _rng(5)
    model = (rng.gamma(0.2, 1.0, (2000, 352)) * rng.uniform(0.01, 100.0, (2000, 1))).astype(np.float32)
    scene = (model[rng.integers(0, 2000, 1500)] * rng.uniform(0.98, 1.02, (1500, 352))).astype(np.float32)
    ref = _with_mode("exact", lambda: ctx.match(model, scene, 1, 1e9))
    for terms in ("tc13", "tc3", "tc1"):
        got = _with_mode(terms, lambda: ctx.match(model, scene, 1, 1e9))
        assert got.tobytes() == ref.tobytes()
        print("%s adversarial: fallback %d, error/bound %.3f" % (terms, ctx.match_fallback_rows(),
                                                                 ctx.match_error_ratio()))
        assert ctx.match_error_ratio() < 0.5


@pytest.mark.parametrize("D", [640, 1024])
def test_tc_filter_long_descriptors(ctx, D):
    """The certificate's error terms scale with the descriptor length (float32 sum error ~ D 2^-23, tensor-core
    accumulation ~ K' 2^-22): long rows of near-duplicates must still give the exact kernel's answer."""
    rng = _rng(D)
    centres = rng.gamma(0.3, 1.0, (200, D)).astype(np.float32)
    model = (centres[rng.integers(0, 200, 2048)] + 0.02 * rng.gamma(0.3, 1.0, (2048, D))).astype(np.float32)
    scene = (model[rng.integers(0, 2048, 1500)] * rng.uniform(0.999, 1.001, (1500, D))).astype(np.float32)
    ref = _with_mode("exact", lambda: ctx.match(model, scene, 1, 1e9))
    for terms in ("tc13", "tc3", "tc1"):
        got = _with_mode(terms, lambda: ctx.match(model, scene, 1, 1e9))
        assert got.tobytes() == ref.tobytes(), (terms, D)
        print("D=%d %s: fallback %d, error/bound %.3f" % (D, terms, ctx.match_fallback_rows(), ctx.match_error_ratio()))
        assert ctx.match_error_ratio() < 0.5


def test_tc_filter_full_cross_check_at_bench_size(ctx):
    """Every row of a bench-sized problem (91 k x 13 k x 352) against the exact float32 kernel — not a sample."""
    rng = _rng(77)
    centres = _shot_like(rng, 2000)
    model = _shot_like(rng, 13049, 352, centres)
    scene = _shot_like(rng, 91076, 352, centres)
    ref = _with_mode("exact", lambda: ctx.match(model, scene, 1, 0.25))
    got = _with_mode("tc13", lambda: ctx.match(model, scene, 1, 0.25))
    print("bench size: %d correspondences, first pass left %d rows, exact kernel %d rows, error/bound %.3f" %
          (len(ref), ctx.match_pass1_rows(), ctx.match_fallback_rows(), ctx.match_error_ratio()))
    assert got.tobytes() == ref.tobytes()
    assert ctx.match_error_ratio() < 0.5
