"""Hypothesis-verification scenarios shared by the CPU and GPU tests (SHOT_hypothesis.cpp:631-653): a scene with
known joint poses and a set of hypotheses — the true poses, near-duplicates of them, displaced copies and one
placed nowhere — as the registered instances ICP would hand to GlobalHypothesesVerification."""
import numpy as np


def _place(m, T):
    return (m.astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)


def hypotheses(synth, joints, poses, seed, n_model=20000):
    rng = np.random.Generator(np.random.PCG64(77 + seed))
    hyps, kind = [], []
    for j, T in zip(joints, poses):
        m = synth.make_model(j, n_model)
        hyps.append(_place(m, T))
        kind.append("true")
        T2 = T.copy()
        T2[:3, 3] += rng.normal(0, 0.002, 3)
        hyps.append(_place(m, T2))
        kind.append("near")
        T3 = T.copy()
        T3[:3, 3] += np.array([0.04, -0.03, 0.05])
        hyps.append(_place(m, T3))
        kind.append("displaced")
    hyps.append(_place(synth.make_model(joints[0], n_model), synth.random_pose(rng)))
    kind.append("nowhere")
    return hyps, kind


def cluttered(synth, n_scene=300000, seed=0):
    """Area-uniform scene (not camera consistent: use without occlusion reasoning)."""
    joints = ("y", "diagonal")
    scene, poses = synth.make_scene(joints, n_scene, scene_id=seed, return_poses=True)
    hyps, kind = hypotheses(synth, joints, poses, seed)
    return scene, hyps, kind


def kinect(synth, target=400000, seed=0):
    """Depth-image scene seen from the origin: the case occlusion reasoning is made for."""
    joints = ("y", "diagonal", "horizontal")
    scene, poses = synth.make_kinect_scene(joints, target, scene_id=seed, return_poses=True)
    hyps, kind = hypotheses(synth, joints, poses, seed)
    return scene, hyps, kind


def random_cues(seed, H=12, ns=5000, n_cells=4000):
    """Random cue lists for the annealing alone: overlapping explained sets and occupancy cells."""
    rng = np.random.Generator(np.random.PCG64(seed))
    eo, ei, ew, oo, oi = [0], [], [], [0], []
    for h in range(H):
        centre = rng.integers(0, ns)
        n = int(rng.integers(50, 1500))
        idx = np.unique((centre + rng.integers(-1500, 1500, n)) % ns)
        ei.append(idx)
        ew.append(rng.uniform(0.0, 1.0, len(idx)).astype(np.float32))
        eo.append(eo[-1] + len(idx))
        c = np.unique((rng.integers(0, n_cells) + rng.integers(-300, 300, int(rng.integers(20, 400)))) % n_cells)
        rng.shuffle(c)
        oi.append(c)
        oo.append(oo[-1] + len(c))
    ow = rng.choice([1.0, 3.0, 0.5], H).astype(np.float32)
    bad = rng.integers(0, 800, H).astype(np.int32)
    return (ns, np.array(eo, np.int32), np.concatenate(ei).astype(np.int32), np.concatenate(ew),
            np.array(oo, np.int32), np.concatenate(oi).astype(np.int32), n_cells, ow, bad)
