#!/usr/bin/env python
"""bench.py — SHOT registration hot path on B200 (contract: see the task statement / DESIGN.md §5).

A "step" is one pass of the hot path over one scene per rank: normals (k=20) → SHOT352 (r=0.02) at the
scene keypoints → correspondence search against the resident model descriptor library (k=1,
d2<0.25) → geometric-consistency grouping + RANSAC poses (0.02 / 2); parameters of SHOT_scenes.cpp
(:50-55, :286, :360).  Workload = BASELINE.json's target run: a 1 M-point Kinect-like synthetic scene
with three joints against a 50 k-point Y-joint model.

  value  : SHOT descriptors/s through the whole step, inputs resident in HBM (device-timed)
  e2e    : the same through the host-buffer C-ABI call (b200_register_scene_shot), pinned host
           buffers in, host results out, copies inside the timed region
  --impl reference : the CPU restatement of the reference's PCL path (oracle/) on the host cores,
           on a bounded sample of the same workload
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "3d-object-detection-of-industrial-joints_b200"
METRIC = "SHOT descriptors/s + scene registrations/s at 1-8 B200 vs PCL OMP host CPU"

PARAMS = dict(normal_k=20, descr_radius=0.02, match_mode=1, match_thr=0.25, gc_size=0.02, gc_threshold=2,
              max_instances=4096)
MODEL_SS, SCENE_SS = 0.005, 0.01


_REAL_STDOUT = None

# DRAM read + write bytes per launch of each stage's main kernel, from the committed `ncu --set full` capture of the
# default workload (profiles/summary_r01.md)
TRAFFIC_NCU = {"match_filter": 332.2e6, "gc_group": 25.7e6, "gc_ransac": 1.4e6, "match": 332.2e6 + 142.8e6,
               "normals": 35.3e6, "shot": 145.9e6, "gc_adjacency": 50.4e6, "gc_sort": 1.0e6}


def _emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def workload(scene_id, scene_points, model_points):
    synth = importlib.import_module(PKG).synth
    model = synth.make_model("y", model_points)
    scene = synth.make_kinect_scene(("y", "diagonal", "horizontal"), target_points=scene_points, scene_id=scene_id)
    return dict(model=model, scene=scene, model_kp=synth.uniform_sampling(model, MODEL_SS),
                scene_kp=synth.uniform_sampling(scene, SCENE_SS))


def config_dict(args, wl, world):
    synth = importlib.import_module(PKG).synth
    return {
        "workload": "SHOT_scenes pipeline, north-star target: %d-pt Kinect-like scene (Y+diagonal+horizontal joints) "
                    "vs %d-pt Y-joint model; one scene per GPU per step" % (len(wl["scene"]), len(wl["model"])),
        "params": {"normals_k": 20, "uniform_sampling": {"model": MODEL_SS, "scene": SCENE_SS}, "shot_radius": 0.02,
                   "match": "k=1, d2<0.25", "gc": {"size": 0.02, "threshold": 2}},
        "N_scene": int(len(wl["scene"])), "N_model": int(len(wl["model"])), "K_scene": int(len(wl["scene_kp"])),
        "K_model": int(len(wl["model_kp"])), "shapes": synth.SHAPE_INFO, "scenes_per_step": world,
        "parallelism": "scene-sharded x%d (one scene per rank per step, same synthetic scene on every rank), model "
                       "library replicated, NCCL gather of the correspondence lists; %d scenes in flight per GPU "
                       "(lanes)" % (world, args.lanes),
        "lanes_per_gpu": args.lanes,
        "host_waits": "blocking events" if getattr(args, "blocking", False) else "spinning",
        "l2": "256 MiB buffer written on the step's stream before every step, inside the timed region (lanes pass); "
              "between steps, outside the timed intervals, in the single-lane pass",
    }


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
def cpu_pipeline_sample(wl, n_sample, threads=None):
    """The reference path (CPU restatement, oracle/) on a bounded sample: normals for the full scene,
    then SHOT / matching / grouping for an evenly strided subset of the scene keypoints against the
    full model library.  Returns (estimated full-workload seconds, detail dict)."""
    from oracle import pcl_oracle as orc
    if threads:
        orc.set_num_threads(threads)
    scene, kp = wl["scene"], wl["scene_kp"]
    stride = max(1, len(kp) // max(n_sample, 1))
    sub = np.ascontiguousarray(kp[::stride])
    f = len(sub) / len(kp)
    t0 = time.perf_counter()
    nrm = orc.normals(scene, k=PARAMS["normal_k"])
    t1 = time.perf_counter()
    desc, _ = orc.shot352(scene, nrm, sub, PARAMS["descr_radius"])
    t2 = time.perf_counter()
    corr = orc.match(wl["model_desc"], desc, PARAMS["match_mode"], PARAMS["match_thr"])  # serial, as SHOT.cpp:409-423
    t3 = time.perf_counter()
    orc.gc_recognize(wl["model_kp"], sub, corr, PARAMS["gc_size"], PARAMS["gc_threshold"], max_inst=4096)
    t4 = time.perf_counter()
    corr_omp = orc.match(wl["model_desc"], desc, PARAMS["match_mode"], PARAMS["match_thr"], omp=True)
    t5 = time.perf_counter()
    assert corr_omp.tobytes() == corr.tobytes()
    t_norm, t_shot, t_match, t_gc, t_match_omp = t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4
    # keypoint-proportional stages are scaled by 1/f (grouping is O(C^2): linear scaling favours the CPU)
    full = t_norm + (t_shot + t_match + t_gc) / f
    full_omp = t_norm + (t_shot + t_match_omp + t_gc) / f
    detail = {"sample_keypoints": int(len(sub)), "fraction": f, "t_normals_s": t_norm, "t_shot_s": t_shot,
              "t_match_serial_s": t_match, "t_match_omp_s": t_match_omp, "t_gc_s": t_gc, "sample_corrs": int(len(corr)),
              "est_full_s": full, "est_full_s_with_omp_match": full_omp}
    return full, detail


def model_descriptors_cpu(wl):
    from oracle import pcl_oracle as orc
    nrm = orc.normals(wl["model"], k=PARAMS["normal_k"])
    desc, _ = orc.shot352(wl["model"], nrm, wl["model_kp"], PARAMS["descr_radius"])
    return desc


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import pcl_oracle as orc
    wl = workload(0, args.scene_points, args.model_points)
    wl["model_desc"] = model_descriptors_cpu(wl)  # resident library: untimed setup, like the GPU arm
    cores = orc.num_threads()
    K = len(wl["scene_kp"])
    times, detail = [], None
    for s in range(args.warmup + args.steps):
        full, detail = cpu_pipeline_sample(wl, args.cpu_sample)
        if s >= args.warmup:
            times.append(full)
    est = float(np.mean(times))
    value = K / est
    sample = ("normals on the full %d-pt scene + SHOT/match/GC on %d of %d scene keypoints (every %d-th) vs the full "
              "%d-descriptor model library; keypoint-proportional stage times scaled by 1/fraction; serial matching "
              "loop as in the reference (SHOT.cpp:409-423)" %
              (len(wl["scene"]), detail["sample_keypoints"], K, max(1, K // max(args.cpu_sample, 1)),
               len(wl["model_kp"])))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "descriptors/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": est * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32/f64", "data": "synthetic",
        "config": config_dict(args, wl, 1),
        "registrations_per_s": 1.0 / est,
        "cpu_baseline": {"value": value, "unit": "descriptors/s", "cores": cores, "kind": "port", "sample": sample,
                         "detail": detail,
                         "value_with_omp_matching": K / detail["est_full_s_with_omp_match"]},
        "e2e": {"value": value, "unit": "descriptors/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "PCL-semantics restatement (oracle/), not libpcl: PCL is neither vendored in the reference nor "
                "installed here",
    }
    _emit(line)


# --------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    binding = importlib.import_module(PKG + ".binding")
    sharding = importlib.import_module(PKG + ".sharding")
    binding.lib()  # fail loudly if the CUDA library is missing
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # weak scaling: every rank registers one scene per step; all ranks use the same synthetic scene so that
    # the per-GPU work is identical (distinct scenes differ by up to 25 % in correspondences / instances and the
    # step time is the max over ranks)
    wl = workload(0, args.scene_points, args.model_points)
    p = binding.shot_params(**PARAMS)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    # Lanes: scenes are independent, so a rank keeps L of them in flight, each on its own context + stream +
    # host thread (sharding.run_lanes).  The grouping stage of one scene is a latency-bound chain on 8 SMs; the
    # other lanes' wide stages (normals, SHOT, matching) fill the rest of the GPU meanwhile.
    L = max(1, args.lanes)
    # every lane is a host thread that spins in stream synchronisations: keep lanes x ranks within the host cores
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    # more lane threads than cores: the contexts wait on blocking events (the threads sleep) instead of spinning
    blocking = world * L > cores or args.blocking_sync
    args.lanes = L
    args.blocking = blocking
    streams = [torch.cuda.Stream(device=dev) for _ in range(L)]
    ctxs = [binding.Context(local_rank, stream=st.cuda_stream) for st in streams]
    if blocking:
        for c in ctxs:
            c.set_blocking_sync(True)
    ctx = ctxs[0]
    with torch.cuda.stream(streams[0]):
        model = ctx.model_create_shot(wl["model"], wl["model_kp"], p)   # resident, replicated library (setup)
    ctx.sync()
    N, Ks, Km = len(wl["scene"]), len(wl["scene_kp"]), model.size
    d_scene = torch.from_numpy(wl["scene"]).to(dev)
    d_kp = torch.from_numpy(wl["scene_kp"]).to(dev)
    mi = PARAMS["max_instances"]
    # the gather needs equally sized correspondence buffers on every rank (scenes differ in K_s)
    corr_cap_all = sharding.common_capacity(Ks, device=dev) if world > 1 else Ks

    def make_out():
        return {"transforms": torch.zeros(mi * 16, dtype=torch.float32, device=dev),
                "inst_offsets": torch.zeros(mi + 1, dtype=torch.int32, device=dev),
                "inst_counts": torch.zeros(mi, dtype=torch.int32, device=dev),
                "inst_corrs": torch.zeros((Ks, 3), dtype=torch.int32, device=dev), "corr_cap": Ks,
                "n_inst": torch.zeros(1, dtype=torch.int32, device=dev),
                "corrs": torch.zeros((corr_cap_all, 3), dtype=torch.int32, device=dev),
                "n_corrs": torch.zeros(1, dtype=torch.int32, device=dev)}

    outs = [make_out() for _ in range(L)]
    flushes = [torch.empty(256 << 20, dtype=torch.uint8, device=dev) for _ in range(L)]
    out = outs[0]
    gate = sharding.Turnstile()   # collectives are issued in global step order on every rank
    torch.cuda.synchronize()

    def step(lane, s, flush=False):
        with torch.cuda.stream(streams[lane]):
            if flush:
                flushes[lane].zero_()
            ctxs[lane].dev_register_scene_shot(model, d_scene, N, 3, d_kp, Ks, 3, p, outs[lane])
            if world > 1:   # the path's one exchange: gather the correspondence lists (NCCL over NVLink)
                gate.run(s, lambda: sharding.gather_correspondences(outs[lane]["corrs"], outs[lane]["n_corrs"]))

    gate.reset()
    sharding.run_lanes(L, args.warmup * L, step)
    torch.cuda.synchronize()

    # ---- pass A, one lane: per-step latency and the per-stage device times (roofline attribution) ----
    stream = streams[0]
    ctx.set_profiling(True)
    ctx.reset_profiling()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_begin = time.time()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    gate.reset()
    for s in range(args.steps):
        with torch.cuda.stream(stream):
            flushes[0].zero_()
            starts[s].record(stream)
        step(0, s)
        ends[s].record(stream)
    torch.cuda.synchronize()
    step_ms = [a.elapsed_time(b) for a, b in zip(starts, ends)]
    stages = ctx.stage_times()
    ctx.set_profiling(False)
    mean_nbrs, max_nbrs = ctx.neighbor_stats()
    n_inst = int(out["n_inst"].item())
    n_corrs = int(out["n_corrs"].item())
    single_ms = float(np.mean(step_ms))

    # ---- pass B, L lanes: the throughput the metric is quoted on.  One start event when the device is idle,
    # one end event per lane stream; the L2 flush (256 MiB write) runs before every step INSIDE the timed region
    gate.reset()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = sum(c.launches for c in ctxs)
    ev_start = torch.cuda.Event(enable_timing=True)
    ev_ends = [torch.cuda.Event(enable_timing=True) for _ in range(L)]
    ev_start.record(streams[0])
    stagger = single_ms / 1e3 / L   # inside the timed region: lane l starts l/L of a scene latency late
    sharding.run_lanes(L, args.steps, lambda lane, s: step(lane, s, flush=True), stagger_s=stagger)
    for l in range(L):
        ev_ends[l].record(streams[l])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end)
    total_ms = max(ev_start.elapsed_time(e) for e in ev_ends)
    launches = sum(c.launches for c in ctxs) - launches0
    assert all(int(o["n_inst"].item()) == n_inst for o in outs[:min(L, args.steps)])

    ms = total_ms / args.steps
    t = torch.tensor([ms, float(Ks), single_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_max, total_desc, single_ms_max = float(tmax[0]), float(tsum[1]), float(tmax[2])
    else:
        ms_max, total_desc, single_ms_max = ms, float(Ks), single_ms

    # ---- end to end through the host-buffer C-ABI call: pinned host buffers, copies timed, same lanes ----
    h_scene = torch.from_numpy(wl["scene"]).pin_memory()
    h_kp = torch.from_numpy(wl["scene_kp"]).pin_memory()
    hs, hk = h_scene.numpy(), h_kp.numpy()
    results = [None] * L

    def e2e_step(lane, s):
        results[lane] = ctxs[lane].register_scene_shot(model, hs, hk, p)
        if world > 1:
            with torch.cuda.stream(streams[lane]):
                gate.run(s, lambda: sharding.gather_correspondences(outs[lane]["corrs"], outs[lane]["n_corrs"]))

    gate.reset()
    sharding.run_lanes(L, 2 * L, e2e_step)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e2e_steps = max(6 * L, args.steps)   # enough steps per lane for the lanes to fall out of lockstep
    gate.reset()
    t0 = time.perf_counter()
    sharding.run_lanes(L, e2e_steps, e2e_step, stagger_s=stagger)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    res = results[0]
    h2d = hs.nbytes + hk.nbytes
    d2h = int(res["transforms"].nbytes + 4 * (len(res["instances"]) + 2) + 12 * len(res["corrs"]) +
              12 * sum(len(i) for i in res["instances"]))
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms_max = float(te[0])

    if rank == 0:
        # ---- roofline of the dominant kernel (stage with the largest device time) ----
        work = {
            "normals": ("hbm", 32.0 * N, "GB/s"),
            "shot": ("hbm", 1444.0 * Ks + 32.0 * N, "GB/s"),
            "neighbor_count": ("hbm", 16.0 * N + 4.0 * Ks, "GB/s"),
            "grid_build": ("hbm", 40.0 * N, "GB/s"),
            "match": ("tensor", 2.0 * Ks * Km * 352, "TFLOP/s"),
            "match_filter": ("tensor", 2.0 * Ks * Km * 352, "TFLOP/s"),   # nested in "match": the tcgen05 kernel alone
            "gc_adjacency": ("hbm", 32.0 * n_corrs + n_corrs * (n_corrs / 8.0), "GB/s"),
            "gc_group": ("hbm", 12.0 * n_corrs + 16.0 * (Ks + Km), "GB/s"),
            "gc_sort": ("hbm", 12.0 * n_corrs * 2, "GB/s"),
            "gc_ransac": ("hbm", 12.0 * n_corrs, "GB/s"),
        }
        stage_ms = {k: (v[0] / max(args.steps, 1), v[1] // max(args.steps, 1)) for k, v in stages.items() if v[1] > 0}
        # The dominant kernel is the stage with the largest SM-time: with several scenes in flight the step rate is
        # set by the GPU-wide stages (their sum is the lanes-pass step time); the grouping stage is one 8-CTA
        # cluster (8 of the SMs) and runs beside the other lanes' kernels, so its wall time counts 8/SMs.
        sm_total = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_share = {"gc_group": 8.0 / sm_total}
        sm_ms = {k: stage_ms[k][0] * sm_share.get(k, 1.0) for k in stage_ms if k in work}
        # dominant STAGE among the top-level ones; when it is matching, the roofline line is its dominant KERNEL
        # (the tcgen05 filter, timed by its own nested event pair)
        dom = max((k for k in sm_ms if k != "match_filter"), key=lambda k: sm_ms[k])
        if dom == "match" and "match_filter" in sm_ms:
            dom = "match_filter"
        # DRAM read + write per launch of the stage's main kernel, from the committed `ncu --set full`
        # capture of this workload (profiles/summary_r01.md); None for stages not captured
        traffic_ncu = TRAFFIC_NCU

        def roof(k):
            bound, units, unit = work[k]
            dur_s = stage_ms[k][0] / 1e3
            if bound == "hbm":
                achieved, peak = units / dur_s / 1e9, peaks.get("hbm_gbs", 6650.0)
            else:
                achieved, peak = units / dur_s / 1e12, peaks.get("bf16_tflops_sustained", 1400.0)
            r = {"kernel": k, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                 "frac": achieved / peak, "traffic": traffic_ncu.get(k), "algorithmic_units": units,
                 "avg_stage_ms": stage_ms[k][0], "sm_ms": sm_ms[k]}
            if k in ("match", "match_filter"):
                # exact float32 results from fp16 tensor cores: each operand is split into hi + lo halves and three
                # of the four products are issued (the stage also holds the exact rescoring of 8 candidates per row)
                r["note"] = ("algorithmic flop = 2*K_s*K_m*352; the tcgen05 filter issues 3x that (fp16 hi/lo split, "
                             "error ~2^-22) and runs at 88 % tensor-pipe activity (ncu, profiles/summary_r01.md); "
                             "exact FP32 rescoring + certificate make the result bit-identical to the FP32 search")
                if k == "match_filter":
                    r["kernel"] = "tc_filter_kernel"
                    r["issued_tflops"] = 3.0 * achieved
            return r

        roofline = roof(dom)
        roofline.update({"traffic_source": "ncu --set full, profiles/summary_r01.md (default workload only)",
                         "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                         "selection": "largest SM-time (stage time x share of the SMs it occupies); stage times are "
                                      "CUDA-event pairs around the stage in the single-lane pass"})
        roofline_all = [roof(k) for k in sorted(sm_ms, key=lambda k: -sm_ms[k])]
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            wl["model_desc"], _ = model.download()
            from oracle import pcl_oracle as orc
            full, detail = cpu_pipeline_sample(wl, args.cpu_sample)
            K = len(wl["scene_kp"])
            cpu_baseline = {
                "value": K / full, "unit": "descriptors/s", "cores": orc.num_threads(), "kind": "port",
                "sample": "normals on the full scene + SHOT/match/GC on %d of %d scene keypoints vs the full model "
                          "library, keypoint-proportional stages scaled by 1/fraction; serial matching loop as in the "
                          "reference; PCL-semantics restatement, not libpcl" % (detail["sample_keypoints"], K),
                "detail": detail, "value_with_omp_matching": K / detail["est_full_s_with_omp_match"]}
        line = {
            "metric": METRIC, "value": total_desc / (ms_max / 1e3), "unit": "descriptors/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 for the SHOT frame/bin decisions)",
            "data": "synthetic", "config": config_dict(args, wl, world),
            "registrations_per_s": world / (ms_max / 1e3),
            "lanes": L,
            "single_lane": {"ms_per_step": single_ms_max, "value": total_desc / (single_ms_max / 1e3),
                            "note": "one scene in flight per GPU (per-scene latency); stage times and the roofline "
                                    "entry are measured in this pass"},
            "e2e": {"value": total_desc / (e2e_ms_max / 1e3), "unit": "descriptors/s", "ms_per_step": e2e_ms_max,
                    "registrations_per_s": world / (e2e_ms_max / 1e3), "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": d2h, "api": "b200_register_scene_shot (host buffers, pinned)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_all_stages": roofline_all,
            "cpu_baseline": cpu_baseline,
            "stages_ms_per_step": {k: round(v[0], 4) for k, v in stage_ms.items()},
            "measured": {"N": N, "K_scene": Ks, "K_model": Km, "mean_neighbors": mean_nbrs, "max_neighbors": max_nbrs,
                         "correspondences": n_corrs, "instances": n_inst, "step_ms": [round(x, 3) for x in step_ms]},
        }
        _emit(line)
    model.close()
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scene-points", type=int, default=1_000_000)
    ap.add_argument("--model-points", type=int, default=50_000)
    ap.add_argument("--cpu-sample", type=int, default=1500, help="scene keypoints in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--blocking-sync", action="store_true", help="contexts sleep in host waits instead of spinning")
    ap.add_argument("--lanes", type=int, default=6, help="scenes in flight per GPU (context + stream + host thread each)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner)
    # goes to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
