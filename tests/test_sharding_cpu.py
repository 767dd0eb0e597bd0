"""Multi-GPU layer on CPU: world_size-2 gloo run of the one exchange on the path (gather of the
variable-length correspondence lists) and of the scene → rank partition."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT

CORR = np.dtype([("index_query", "<i4"), ("index_match", "<i4"), ("distance", "<f4")])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_corrs(rank, cap):
    rng = np.random.Generator(np.random.PCG64(100 + rank))
    n = 5 + 7 * rank
    c = np.zeros(cap, dtype=CORR)
    c["index_query"][:n] = rng.integers(0, 1000, n)
    c["index_match"][:n] = np.sort(rng.integers(0, 5000, n))
    c["distance"][:n] = rng.uniform(0, 0.25, n).astype(np.float32)
    return c, n


def _worker(rank, world, port, root, pkg, out_dir):
    import sys
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharding = importlib.import_module(pkg + ".sharding")
    cap = sharding.common_capacity(40 + 24 * rank)      # ranks hold scenes of different size
    c, n = _make_corrs(rank, cap)
    words = torch.from_numpy(c.view(np.int32).reshape(cap, 3).copy())
    counts, allw = sharding.gather_correspondences(words, torch.tensor([n], dtype=torch.int32))
    lists = sharding.unpack_gathered(counts, allw)
    ok = len(lists) == world
    for r in range(world):
        cr, nr = _make_corrs(r, cap)
        ok = ok and cap == 64
        ok = ok and lists[r].tobytes() == cr[:nr].tobytes()
    # every scene is owned by exactly one rank
    mine = sharding.scenes_for_rank(11, rank, world)
    owned = torch.zeros(11, dtype=torch.int32)
    owned[mine] = 1
    dist.all_reduce(owned)
    ok = ok and bool((owned == 1).all()) and mine == list(range(rank, 11, world))
    np.save(os.path.join(out_dir, "ok%d.npy" % rank), np.array([int(ok)]))
    dist.destroy_process_group()


class _FakeCtx:
    """Stands in for binding.Context on CPU: returns a deterministic correspondence list per scene."""

    def register_scene_shot(self, model, scene, kp, params):
        n = int(scene[0, 0])
        c = np.zeros(n, dtype=CORR)
        c["index_query"] = np.arange(n) * 3
        c["index_match"] = np.arange(n)
        c["distance"] = np.float32(0.01) * np.arange(n, dtype=np.float32)
        return {"corrs": c, "n_instances": 0}


def _batch_worker(rank, world, port, root, pkg, out_dir):
    import sys
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharding = importlib.import_module(pkg + ".sharding")
    sizes = [5, 0, 17, 9, 30]                                  # scene s yields sizes[s] correspondences
    scenes = [np.full((4, 3), n, dtype=np.float32) for n in sizes]
    kps = [np.zeros((max(n, 1) + 2 * i, 3), np.float32) for i, n in enumerate(sizes)]
    local, gathered = sharding.register_scene_batch(_FakeCtx(), None, scenes, kps, None, rank=rank, world=world)
    ok = sorted(local) == list(range(rank, len(sizes), world)) and sorted(gathered) == list(range(len(sizes)))
    for s, n in enumerate(sizes):
        ref = _FakeCtx().register_scene_shot(None, scenes[s], None, None)["corrs"]
        ok = ok and gathered[s].tobytes() == ref.tobytes()
    np.save(os.path.join(out_dir, "bok%d.npy" % rank), np.array([int(ok)]))
    dist.destroy_process_group()


def test_scene_batch_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_batch_worker, args=(world, _free_port(), ROOT, PKG_NAME, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert int(np.load(tmp_path / ("bok%d.npy" % r))[0]) == 1


def test_gather_correspondences_gloo_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), ROOT, PKG_NAME, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert int(np.load(tmp_path / ("ok%d.npy" % r))[0]) == 1


def test_bench_reference_arm_only_rank0(tmp_path):
    """bench.py --impl reference under a 2-rank launch: rank 1 exits without work (no output)."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _lanes_worker(rank, world, port, root, pkg, out_dir):
    import random
    import sys
    import time
    sys.path.insert(0, root)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharding = importlib.import_module(pkg + ".sharding")
    n_lanes, n_steps, cap = 3, 10, 16
    gate = sharding.Turnstile()
    results = {}
    rnd = random.Random(17 * rank + 1)
    delays = [rnd.uniform(0.0, 0.02) for _ in range(n_steps)]   # lanes reach their steps in a different order per rank

    def step(lane, s):
        time.sleep(delays[s])
        n = (s * 5 + rank * 3) % cap
        words = torch.full((cap, 3), -1, dtype=torch.int32)
        words[:n] = 1000 * s + rank
        count = torch.tensor([n], dtype=torch.int32)
        results[s] = gate.run(s, lambda: sharding.gather_correspondences(words, count))

    sharding.run_lanes(n_lanes, n_steps, step)
    ok = sorted(results) == list(range(n_steps))
    for s in range(n_steps):
        counts, words = results[s]
        for r in range(world):
            n = (s * 5 + r * 3) % cap
            ok = ok and int(counts[r]) == n and bool((words[r, :n] == 1000 * s + r).all())
    # an exception on one lane is re-raised on the caller's thread
    try:
        sharding.run_lanes(2, 4, lambda lane, s: (_ for _ in ()).throw(ValueError("boom")) if s == 1 else None)
        ok = False
    except ValueError:
        pass
    np.save(os.path.join(out_dir, "lok%d.npy" % rank), np.array([int(ok)]))
    dist.destroy_process_group()


def test_lanes_keep_collective_order_gloo_world2(tmp_path):
    """Several lanes (host threads) per rank: the per-step gather is issued in global step order on every rank
    (Turnstile), whatever order the lanes reach it in."""
    world = 2
    mp.spawn(_lanes_worker, args=(world, _free_port(), ROOT, PKG_NAME, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert int(np.load(tmp_path / ("lok%d.npy" % r))[0]) == 1


def test_bench_reference_arm_json_contract():
    """bench.py --impl reference (the CPU arm) prints exactly one JSON line with the contract's keys; run on a reduced
    workload so that the CPU suite stays short."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--scene-points", "150000", "--model-points", "8000", "--cpu-sample", "200"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "descriptors/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
