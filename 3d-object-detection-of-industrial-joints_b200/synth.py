"""Synthetic pipe-joint clouds: analytic chord+stub joints, area-uniform models,
cluttered scenes and a Kinect-like ray-cast scene through the pinhole model of the reference's
depth-sensor bridge (ROS_server.cpp:2144-2157: f = (W/2)/tan(fov/2), x = -(i-W/2) z/f,
y = (j-H/2) z/f; fov 57 deg from render.cpp:26).

The reference's own CAD/scan files are not in its repository (README.md:9-12), so the shapes and
dimensions below are this build's choice and are recorded in every bench line.

Also holds the keypoint extractors the reference programs run between normals and descriptors
(pcl::VoxelGrid SHOT_demo.cpp:413-417, pcl::UniformSampling SHOT.cpp:314-323).  They are harness-side
numpy restatements; the same keypoints are fed to the CUDA path and to the oracle.
"""
import numpy as np

CHORD_R, CHORD_L = 0.060, 0.60
STUB_R, STUB_L = 0.045, 0.35
JOINTS = {"y": 60.0, "diagonal": 45.0, "horizontal": 90.0}
JOINT_IDS = {"y": 0, "diagonal": 1, "horizontal": 2}
SHAPE_INFO = {"chord_R": CHORD_R, "chord_L": CHORD_L, "stub_R": STUB_R, "stub_L": STUB_L,
              "theta_deg": JOINTS}


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def _stub_frame(theta_deg):
    th = np.deg2rad(theta_deg)
    a = np.array([np.cos(th), 0.0, np.sin(th)])          # stub axis
    u = np.array([0.0, 1.0, 0.0])
    v = np.cross(a, u)
    return a, u, v


def joint_surface(joint, n, rng):
    """Area-uniform samples (points, outward normals) of the joint surface, exactly n rows."""
    theta = JOINTS[joint]
    a, u, v = _stub_frame(theta)
    area_c = 2 * np.pi * CHORD_R * CHORD_L
    area_s = 2 * np.pi * STUB_R * STUB_L
    pts, nrm = [], []
    have = 0
    while have < n:
        m = int((n - have) * 1.6) + 64
        mc = rng.binomial(m, area_c / (area_c + area_s))
        ms = m - mc
        # chord: axis x, centred at the origin
        ang = rng.uniform(0, 2 * np.pi, mc)
        x = rng.uniform(-CHORD_L / 2, CHORD_L / 2, mc)
        pc = np.stack([x, CHORD_R * np.cos(ang), CHORD_R * np.sin(ang)], 1)
        nc = np.stack([np.zeros(mc), np.cos(ang), np.sin(ang)], 1)
        # remove chord points inside the stub
        t = pc @ a
        rad = np.linalg.norm(pc - np.outer(t, a), axis=1)
        keep = ~((t > 0) & (t < STUB_L) & (rad < STUB_R))
        pc, nc = pc[keep], nc[keep]
        # stub: axis a from the origin
        ang = rng.uniform(0, 2 * np.pi, ms)
        t = rng.uniform(0, STUB_L, ms)
        ns = np.outer(np.cos(ang), u) + np.outer(np.sin(ang), v)
        ps = np.outer(t, a) + STUB_R * ns
        keep = ~(((ps[:, 1] ** 2 + ps[:, 2] ** 2) < CHORD_R ** 2) & (np.abs(ps[:, 0]) <= CHORD_L / 2))
        ps, ns = ps[keep], ns[keep]
        p = np.concatenate([pc, ps])
        q = np.concatenate([nc, ns])
        perm = rng.permutation(len(p))
        pts.append(p[perm])
        nrm.append(q[perm])
        have += len(p)
    return np.concatenate(pts)[:n], np.concatenate(nrm)[:n]


def make_model(joint="y", n=5000, seed=None):
    """Model cloud: area-uniform, no noise, seed 1000 + joint id.  float32 (n, 3)."""
    seed = 1000 + JOINT_IDS[joint] if seed is None else seed
    p, _ = joint_surface(joint, n, _rng(seed))
    return p.astype(np.float32)


def view_directions(n_views=64):
    """Camera directions of the rendered partial views: a Fibonacci sphere (render.cpp:30-35 uses the 42
    vertices of a tessellated icosphere; BASELINE.json config 4 asks for 64)."""
    i = np.arange(n_views) + 0.5
    phi = np.arccos(1 - 2 * i / n_views)
    th = np.pi * (1 + 5 ** 0.5) * i
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)


def make_partial_view(joint="y", view=0, n=4000, n_views=64):
    """Partial view of a CAD joint as seen from view direction `view`: the surface points whose outward
    normal faces the camera (self-occlusion between chord and stub is ignored).  float32 (m, 3), m ~ n / 2."""
    p, q = joint_surface(joint, n, _rng(3000 + 97 * JOINT_IDS[joint] + view))
    d = view_directions(n_views)[view]
    return p[q @ d > 0.1].astype(np.float32)


def random_pose(rng, max_deg=60.0, trans=0.3, z=1.0):
    """Rigid pose: rotation within +-max_deg about each axis, translation U[-trans, trans]^3 + (0,0,z)."""
    ax, ay, az = np.deg2rad(rng.uniform(-max_deg, max_deg, 3))
    Rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    Ry = np.array([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]])
    Rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = rng.uniform(-trans, trans, 3) + np.array([0, 0, z])
    return T


def _box_surface(center, size, n, rng):
    """Area-uniform samples on the 5 visible faces (no bottom) of an axis-aligned box."""
    sx, sy, sz = size
    faces = [(0, +1, sy * sz), (0, -1, sy * sz), (1, +1, sx * sz), (1, -1, sx * sz), (2, +1, sx * sy)]
    areas = np.array([f[2] for f in faces])
    which = rng.choice(len(faces), n, p=areas / areas.sum())
    p = rng.uniform(-0.5, 0.5, (n, 3)) * np.array(size)
    nr = np.zeros((n, 3))
    for k, (axis, sign, _) in enumerate(faces):
        m = which == k
        p[m, axis] = sign * 0.5 * size[axis]
        nr[m, axis] = sign
    return p + np.array(center), nr


def make_scene(joints=("y",), n=100000, scene_id=0, noise=0.0005, return_poses=False):
    """Cluttered scene: joint(s) under random poses + 2x2 m ground plane + 4 boxes, area-uniform,
    Gaussian noise along the normal, seed 2000 + scene_id.  float32 (n, 3)."""
    rng = _rng(2000 + scene_id)
    poses = [random_pose(rng) for _ in joints]
    boxes = []
    for _ in range(4):
        size = rng.uniform(0.10, 0.25, 3)
        c = np.array([rng.uniform(-0.8, 0.8), rng.uniform(-0.8, 0.8), 0.0])
        boxes.append((c, size))
    ground_z = 0.45  # joints float above the ground plane (z ~ 1.0 +- 0.3), clutter sits on it
    area_joint = 2 * np.pi * (CHORD_R * CHORD_L + STUB_R * STUB_L)
    areas = [area_joint] * len(joints) + [4.0] + [2 * (s[1] * s[2] + s[0] * s[2]) + s[0] * s[1] for _, s in boxes]
    areas = np.array(areas)
    counts = rng.multinomial(n, areas / areas.sum())
    P, N = [], []
    for j, T, c in zip(joints, poses, counts[:len(joints)]):
        p, q = joint_surface(j, int(c), rng)
        P.append(p @ T[:3, :3].T + T[:3, 3])
        N.append(q @ T[:3, :3].T)
    c = int(counts[len(joints)])
    g = np.stack([rng.uniform(-1, 1, c), rng.uniform(-1, 1, c), np.full(c, ground_z)], 1)
    P.append(g)
    N.append(np.tile([0.0, 0.0, 1.0], (c, 1)))
    for (cen, size), c in zip(boxes, counts[len(joints) + 1:]):
        cen = cen + np.array([0, 0, ground_z + size[2] / 2])
        p, q = _box_surface(cen, size, int(c), rng)
        P.append(p)
        N.append(q)
    P = np.concatenate(P)
    N = np.concatenate(N)
    P = P + N * rng.normal(0.0, noise, (len(P), 1))
    perm = rng.permutation(len(P))
    P = P[perm].astype(np.float32)
    if return_poses:
        return P, poses
    return P


# ---------------------------------------------------------------------------------------------
# Kinect-like ray-cast scene
# ---------------------------------------------------------------------------------------------
def _ray_cylinder(d, c, a, R, L0, L1):
    """Rays o=0, directions d (M,3) against a tube |x - c - ((x-c).a)a| = R, t in [L0, L1] along a.
    Returns the nearest positive ray parameter s (inf where missed); the tube is open (no caps)."""
    da = d @ a
    dp = d - np.outer(da, a)
    oc = -c
    oa = oc @ a
    op = oc - oa * a
    A = np.einsum("ij,ij->i", dp, dp)
    B = 2 * dp @ op
    Cc = op @ op - R * R
    disc = B * B - 4 * A * Cc
    s = np.full(len(d), np.inf)
    ok = (disc >= 0) & (A > 1e-14)
    sq = np.sqrt(np.where(ok, disc, 0))
    for sign in (-1.0, 1.0):          # near root first, then far root (inside wall of the open tube)
        cand = np.where(ok, (-B + sign * sq) / (2 * np.where(ok, A, 1)), np.inf)
        t = oa + cand * da
        hit = ok & (cand > 1e-6) & (t >= L0) & (t <= L1) & (cand < s)
        s = np.where(hit, cand, s)
    return s


def _ray_box(d, lo, hi):
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        t0 = lo * inv
        t1 = hi * inv
    tmin = np.nanmax(np.minimum(t0, t1), axis=1)
    tmax = np.nanmin(np.maximum(t0, t1), axis=1)
    return np.where((tmax >= tmin) & (tmin > 1e-6), tmin, np.inf)


def make_kinect_scene(joints=("y", "diagonal", "horizontal"), target_points=1_000_000, scene_id=0,
                      fov_deg=57.0, depth_noise=0.0012, return_poses=False):
    """Organised depth image of an analytic scene (3 joints + floor + back wall + 4 boxes), NaN rows
    removed like removeNaNFromPointCloud (SHOT.cpp:298-299).  About target_points rows, float32."""
    rng = _rng(2000 + scene_id)
    H = int(round(np.sqrt(target_points * 3.0 / 4.0)))
    W = int(round(target_points / H))
    f = (W / 2.0) / np.tan(np.deg2rad(fov_deg) / 2.0)
    jj, ii = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    d = np.stack([-(ii - W / 2.0) / f, (jj - H / 2.0) / f, np.ones_like(ii, dtype=np.float64)], -1).reshape(-1, 3)
    best = np.full(len(d), np.inf)
    # back wall z = 2.0 and floor y = 0.45 (camera y points down)
    best = np.minimum(best, 2.0 / d[:, 2])
    with np.errstate(divide="ignore"):
        sf = np.where(d[:, 1] > 1e-9, 0.45 / d[:, 1], np.inf)
    best = np.minimum(best, sf)
    poses = []
    slots = [(-0.42, 1.05), (0.0, 1.25), (0.42, 1.05)]
    for k, j in enumerate(joints):
        T = random_pose(rng, max_deg=40.0, trans=0.05, z=0.0)
        sx, sz = slots[k % len(slots)]
        T[:3, 3] += np.array([sx, 0.05, sz])
        poses.append(T)
        R, t = T[:3, :3], T[:3, 3]
        a, _, _ = _stub_frame(JOINTS[j])
        best = np.minimum(best, _ray_cylinder(d, t, R @ np.array([1.0, 0, 0]), CHORD_R, -CHORD_L / 2, CHORD_L / 2))
        best = np.minimum(best, _ray_cylinder(d, t, R @ a, STUB_R, 0.0, STUB_L))
    for _ in range(4):
        size = rng.uniform(0.10, 0.25, 3)
        c = np.array([rng.uniform(-0.7, 0.7), 0.45 - size[1] / 2, rng.uniform(1.3, 1.9)])
        best = np.minimum(best, _ray_box(d, c - size / 2, c + size / 2))
    z = best * d[:, 2]
    z = z + rng.normal(0.0, 1.0, len(z)) * depth_noise * z * z
    P = d * (z / d[:, 2])[:, None]
    P = P[np.isfinite(P).all(1)].astype(np.float32)
    if return_poses:
        return P, poses
    return P


# ---------------------------------------------------------------------------------------------
# Keypoint extractors (harness side)
# ---------------------------------------------------------------------------------------------
def _cell_ids(p, leaf):
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(p[:, :3] * inv).astype(np.int64)
    mn = ijk.min(0)
    dims = ijk.max(0) - mn + 1
    rel = ijk - mn
    idx = rel[:, 0] + dims[0] * (rel[:, 1] + dims[1] * rel[:, 2])
    return ijk, idx


def voxel_grid(points, leaf):
    """pcl::VoxelGrid (SHOT_demo.cpp:413-417): centroid per voxel, ascending voxel index."""
    p = np.ascontiguousarray(points[:, :3], dtype=np.float32)
    _, idx = _cell_ids(p, leaf)
    order = np.argsort(idx, kind="stable")
    sidx = idx[order]
    starts = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]])
    sums = np.add.reduceat(p[order].astype(np.float64), starts, axis=0)
    cnt = np.diff(np.r_[starts, len(p)])[:, None]
    return (sums / cnt).astype(np.float32)


def uniform_sampling(points, leaf):
    """pcl::UniformSampling 1.8 filter (SHOT.cpp:314-323): per cell keep the input point minimising
    ||p - (float)ijk||^2 (sic: the integer index vector, not the metric centre); first wins on ties.
    Output order is defined as ascending cell id (PCL's is hash-map order)."""
    p = np.ascontiguousarray(points[:, :3], dtype=np.float32)
    ijk, idx = _cell_ids(p, leaf)
    diff = ((p - ijk.astype(np.float32)) ** 2).sum(1, dtype=np.float32) + np.float32(1.0)
    order = np.lexsort((np.arange(len(p)), diff, idx))
    sidx = idx[order]
    first = np.flatnonzero(np.r_[True, sidx[1:] != sidx[:-1]])
    return p[order[first]].copy()
