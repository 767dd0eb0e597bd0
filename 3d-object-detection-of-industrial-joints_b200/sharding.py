"""Multi-GPU layer: scenes (or views) are independent units, sharded round-robin across ranks with
the model descriptor library replicated.  The only data-path exchange is the gather
of the per-scene correspondence lists, which have a data-dependent length: one all_gather of the
counts, then one padded all_gather of the 12-byte records (NCCL over NVLink on the GPU box; the same
code runs on gloo/CPU tensors in the tests).  torch.distributed is plumbing only.
"""
import numpy as np

CORR_WORDS = 3  # b200_corr = {int index_query, int index_match, float distance} = 3 x 4 bytes


def scenes_for_rank(n_scenes, rank, world):
    """scene s → rank s mod world."""
    return list(range(rank, n_scenes, world))


def common_capacity(local_cap, device=None, group=None):
    """Largest per-rank correspondence capacity: the padded all_gather needs equally sized buffers on every
    rank (scenes differ in their keypoint count).  Call once at set-up."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(local_cap)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def gather_correspondences(corr_words, count, group=None):
    """corr_words: (cap, 3) int32 tensor viewing this rank's b200_corr buffer; count: (1,) int32 tensor
    with the number of valid rows; cap must be the same on every rank (common_capacity).  Returns (all_counts (world,), all_corrs (world, cap, 3)) on every
    rank.  Rows past a rank's count are padding."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    counts = torch.empty(world, dtype=count.dtype, device=count.device)
    dist.all_gather_into_tensor(counts, count.reshape(1), group=group)
    cap = corr_words.shape[0]
    # concatenated layout (world * cap, 3): the form both the NCCL and the gloo backend accept
    out = torch.empty((world * cap,) + tuple(corr_words.shape[1:]), dtype=corr_words.dtype, device=corr_words.device)
    dist.all_gather_into_tensor(out, corr_words.contiguous(), group=group)
    return counts, out.view((world, cap) + tuple(corr_words.shape[1:]))


def unpack_gathered(counts, words):
    """→ list (one per rank) of structured numpy correspondence arrays."""
    dt = np.dtype([("index_query", "<i4"), ("index_match", "<i4"), ("distance", "<f4")])
    res = []
    c = counts.cpu().numpy()
    w = words.cpu().numpy()
    for r in range(len(c)):
        res.append(np.ascontiguousarray(w[r, :int(c[r])]).view(dt).reshape(-1))
    return res


def register_scene_batch(ctx, model, scenes, keypoints, params, rank=0, world=1, group=None, device=None):
    """BASELINE.json config 5: a batch of scenes sharded round-robin over the ranks (scene s -> rank s mod world),
    every rank registering its scenes (ctx: one context, or a list of contexts = lanes working through the rank's
    scenes concurrently) against the replicated model library through the host-buffer C-ABI call,
    then ONE gather of the per-scene correspondence lists.  Returns (local, gathered): `local` maps scene id ->
    result dict of this rank's scenes (poses and instances stay on the rank that found them); `gathered` maps
    every scene id of the batch -> its correspondence list (structured numpy array), on every rank.
    world == 1 needs no process group."""
    import torch
    mine = scenes_for_rank(len(scenes), rank, world)
    lanes = list(ctx) if isinstance(ctx, (list, tuple)) else [ctx]   # several contexts = several scenes in flight
    local = {}

    def one(lane, k):
        s = mine[k]
        local[s] = lanes[lane].register_scene_shot(model, scenes[s], keypoints[s], params)

    run_lanes(len(lanes), len(mine), one)
    if world == 1:
        return local, {s: local[s]["corrs"] for s in mine}
    import torch.distributed as dist
    per_rank = (len(scenes) + world - 1) // world                      # slots per rank (padded)
    cap = common_capacity(max([len(keypoints[s]) for s in mine] + [1]), device=device, group=group)
    words = torch.zeros((per_rank * cap, CORR_WORDS), dtype=torch.int32, device=device)
    counts = torch.zeros(per_rank, dtype=torch.int32, device=device)
    for slot, s in enumerate(mine):
        c = local[s]["corrs"]
        if len(c):
            words[slot * cap:slot * cap + len(c)] = torch.from_numpy(
                np.ascontiguousarray(c).view(np.int32).reshape(-1, CORR_WORDS)).to(words.device)
        counts[slot] = len(c)
    all_counts = torch.empty(world * per_rank, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(all_counts, counts, group=group)
    all_words = torch.empty((world * per_rank * cap, CORR_WORDS), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(all_words, words, group=group)
    dt = np.dtype([("index_query", "<i4"), ("index_match", "<i4"), ("distance", "<f4")])
    ac = all_counts.cpu().numpy().reshape(world, per_rank)
    aw = all_words.cpu().numpy().reshape(world, per_rank, cap, CORR_WORDS)
    gathered = {}
    for r in range(world):
        for slot, s in enumerate(scenes_for_rank(len(scenes), r, world)):
            gathered[s] = np.ascontiguousarray(aw[r, slot, :int(ac[r, slot])]).view(dt).reshape(-1)
    return local, gathered


class Turnstile:
    """Runs callables in global step order no matter which host thread reaches its step first.  The ranks of a
    job must issue their collectives in the same order; with several lanes per rank the host threads race, so
    each lane takes its turn here before it enqueues the gather of step s."""

    def __init__(self):
        import threading
        self._cv = threading.Condition()
        self._next = 0
        self._failed = False

    def reset(self, first=0):
        with self._cv:
            self._next, self._failed = first, False

    def abort(self):
        with self._cv:
            self._failed = True
            self._cv.notify_all()

    def run(self, step, fn):
        with self._cv:
            while self._next != step and not self._failed:
                self._cv.wait()
            if self._failed:
                raise RuntimeError("turnstile aborted: another lane failed")
            try:
                return fn()
            finally:
                self._next = step + 1
                self._cv.notify_all()


def run_lanes(n_lanes, n_steps, step_fn, first_step=0, stagger_s=0.0):
    """Scenes within one rank are independent too: lane l (one host thread driving its own b200 context and
    CUDA stream) runs steps l, l + n_lanes, ... of `n_steps`, so that the latency-bound grouping stage of one
    scene (8 SMs) overlaps the wide stages of the next ones on the rest of the GPU.  step_fn(lane, step) is
    called on the lane's thread; exceptions are re-raised on the caller's thread.  Returns after every lane has
    issued (not necessarily completed) its steps.  stagger_s: lane l starts l * stagger_s late, so that lanes
    released together do not walk through the same stages in lockstep (all in the wide stages, then all in
    grouping with most of the GPU idle) until they drift apart by themselves."""
    import threading
    errors = []

    def work(lane):
        try:
            if stagger_s > 0.0 and lane > 0:
                import time
                time.sleep(lane * stagger_s)
            for s in range(first_step + lane, first_step + n_steps, n_lanes):
                if errors:
                    break
                step_fn(lane, s)
        except BaseException as e:  # noqa: BLE001 - re-raised below
            errors.append(e)

    if n_lanes == 1:
        work(0)
    else:
        threads = [threading.Thread(target=work, args=(l,), name="b200-lane-%d" % l) for l in range(n_lanes)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
