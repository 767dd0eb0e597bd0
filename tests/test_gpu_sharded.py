"""One scene sharded over the ranks of the library's own NCCL communicator (b200_register_scene_shot_sharded):
the result must be the single-GPU result bit for bit, whatever the number of ranks."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _workload(synth):
    model = synth.make_model("y", 20000)
    scene = synth.make_scene(("y", "diagonal"), 120000, scene_id=11)
    return model, scene, synth.uniform_sampling(model, 0.006), synth.voxel_grid(scene, 0.02)


def _same(a, b):
    return (a["corrs"].tobytes() == b["corrs"].tobytes() and a["n_instances"] == b["n_instances"] and
            np.array_equal(a["transforms"], b["transforms"]) and
            all(x.tobytes() == y.tobytes() for x, y in zip(a["instances"], b["instances"])))


def test_sharded_world1_equals_single(b200, synth):
    model, scene, kpm, kps = _workload(synth)
    p = b200.shot_params(normal_k=10, descr_radius=0.02, match_mode=1, match_thr=0.25, gc_size=0.02, gc_threshold=2,
                         max_instances=4096)
    ctx = b200.Context(0)
    m = ctx.model_create_shot(model, kpm, p)
    ref = ctx.register_scene_shot(m, scene, kps, p)
    got = ctx.register_scene_shot_sharded(m, p, scene, kps)
    assert ref["n_instances"] > 0 and _same(got, ref)
    m.close()
    ctx.close()


def test_sharded_two_gpus_equals_single(b200, synth):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    model, scene, kpm, kps = _workload(synth)
    p = b200.shot_params(normal_k=10, descr_radius=0.02, match_mode=1, match_thr=0.25, gc_size=0.02, gc_threshold=2,
                         max_instances=4096)
    tok = b200.comm_unique_id()
    out, errs = {}, []

    def rank_main(r):
        try:
            ctx = b200.Context(r)
            ctx.comm_init(tok, r, 2)
            m = ctx.model_create_shot(model, kpm, p)
            for _ in range(2):
                res = ctx.register_scene_shot_sharded(m, p, scene if r == 0 else None, kps if r == 0 else None)
            if r == 0:
                out["sharded"] = res
                out["single"] = ctx.register_scene_shot(m, scene, kps, p)
            m.close()
            ctx.close()
        except BaseException as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=rank_main, args=(r,)) for r in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    assert out["single"]["n_instances"] > 0 and _same(out["sharded"], out["single"])
