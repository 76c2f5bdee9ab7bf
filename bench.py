#!/usr/bin/env python
"""bench.py -- rays x bounces / s of the rayverb hot path on B200 (BASELINE.json metric).

A "step" is one pass of the whole hot path over one batch of synthetic rays:
  bounce kernels (nearest hit, specular bounces, diffuse + image-source emission, 8-band air
  attenuation) -> image-source dedupe -> speaker / HRTF attenuation -> deterministic flatten (radix
  sort by sample index + segmented reduction) -> Linkwitz-Riley band filters -> mixdown / normalize /
  trim.

Default workload at EVERY N: C3, the 100k-triangle hall with 1 Mi rays x 128 reflections, stereo --
the configuration BASELINE.json's scaling target is quoted on (it fits one GPU), so the driver's
1 -> 8 GPU curve is STRONG scaling of that one problem.

  value : whole-job rays*bounces/s with the directions already resident in HBM (CUDA events on
          the launching stream, max over ranks)
  e2e   : the same metric through the public C-ABI calls with HOST buffers: pinned-host
          directions -> device inside the timed region, final IR read back to the host
  N > 1 : rays are sharded; one small all-gather (time extrema + image sources), one NCCL reduce of
          the per-GPU histograms, rank 0 filters.  After the timed region rank 0 renders the whole
          ray set alone and the line carries `sharded_parity` (per-band relative L2 of the sharded IR
          against it).
  N = 1 : the line also carries `sub_blocks`: C2 (12-triangle shoebox, the brute-force kernel the
          FP32 roofline target is stated for) and the 4352-triangle brute-force room, each measured
          the same way with its own `roofline`.

`--impl reference` times the CPU restatement of the reference's own algorithm (oracle/) on the
host cores -- the reference's OpenCL build cannot run in this image (no OpenCL ICD).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "parallel-reverb-raytracer_b200", "python"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FLOP_PER_TEST = 52           # Moller-Trumbore, rayverb/kernel.cpp:63-88
FLOP_PER_BOX = 20            # slab test of one child box
SM_COUNT = 148
FP32_LANES = 128
SR = 44100.0
DEFAULT_WORKLOAD = "C3_hall_100ktri_1Mx128_stereo"
SUB_BLOCKS = ["C2_shoebox_12tri_1Mx128_omni", "X_room_4ktri_303kx64_bruteforce"]
PARITY_TOL = 2e-6            # sharded vs one-GPU IR, relative L2 per band (float summation order only)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


def _pairs_grid():
    """C5: 8 x 8 grid of (mic, source) positions on two planes of the hall."""
    jobs = []
    for i in range(8):
        for j in range(8):
            jobs.append(((8.0 + 3.0 * i, 1.5, 16.0 + 0.8 * j), (14.0 + 1.5 * j, 2.0, 3.0 + 0.6 * i)))
    return jobs


# scaling: "weak"   rays = per GPU        (total grows with N)
#          "strong" rays = total          (sharded over N)
#          "jobs"   independent (mic, src) pairs sharded over N, `rays` per pair
WORKLOADS = {
    "C2_shoebox_12tri_1Mx128_omni": dict(scene="shoebox", jobs=[((7, 2, 4), (2, 1.5, 2))], rays=1 << 20, refl=128, scaling="weak",
                                         speakers=[((0, 0, 1), 0.0)], seed=2015),
    # C1: the shipped vault room (3754 triangles after triangulation, tests/golden/vault_scene.npz),
    # vault.json's source / microphone, one omni speaker
    "C1_vault_3754tri_1024x64": dict(scene="vault", jobs=[((0, 1.75, 6), (0, 1.75, 0))], rays=1024, refl=64, scaling="weak",
                                     speakers=[((0, 0, -1), 0.0)], seed=2015, lo_cutoff=60.0),
    "C3_hall_100ktri_1Mx128_stereo": dict(scene="hall", jobs=[((24, 1.5, 18), (20, 2.0, 4))], rays=1 << 20, refl=128, scaling="strong",
                                          speakers=[((-1, 0, -1), 0.5), ((1, 0, -1), 0.5)], seed=2015),
    "C4_cathedral_1Mtri_4Mx256_hrtf": dict(scene="cathedral", jobs=[((15, 1.7, 60), (15, 2.0, 10))], rays=1 << 22, refl=256, scaling="strong",
                                           hrtf=dict(facing=(0, 0, -1), up=(0, 1, 0)), seed=2016),
    "C5_hall_64pairs_16kx128": dict(scene="hall", jobs=_pairs_grid(), rays=1 << 14, refl=128, scaling="jobs",
                                    speakers=[((-1, 0, -1), 0.5), ((1, 0, -1), 0.5)], seed=2015),
    # the brute-force intersector at a size where the tests dominate (roofline of the kernel proper)
    "X_room_4ktri_303kx64_bruteforce": dict(scene="room4k", jobs=[((12, 1.5, 9), (6, 2, 3))], rays=148 * 2048, refl=64, scaling="weak",
                                           speakers=[((0, 0, 1), 0.0)], seed=2015, accel=1),
}


def build_scene(name):
    from rayverb_b200 import scenes
    if name == "shoebox":
        return scenes.shoebox()
    if name == "vault":
        sc = np.load(os.path.join(ROOT, "tests", "golden", "vault_scene.npz"))
        return sc["triangles"], sc["vertices"], sc["surfaces"]
    if name == "room4k":
        tri, vtx = scenes.displaced_box(20.0, 8.0, 14.0, 4400, amp=0.1, fan=0.25, nsurfaces=4)
        return tri, vtx, scenes.hall_surfaces()
    if name == "hall":
        return scenes.concert_hall()
    if name == "cathedral":
        return scenes.cathedral()
    raise KeyError(name)


def render_kwargs(wl):
    from rayverb_b200 import scenes
    kw = dict(sr=SR, ft="linkwitz_riley", lo_cutoff=wl.get("lo_cutoff", 45.0))
    if "hrtf" in wl:
        kw["hrtf"] = dict(table=scenes.synthetic_hrtf_table(), facing=wl["hrtf"]["facing"], up=wl["hrtf"]["up"])
    else:
        kw["speakers"] = wl["speakers"]
    return kw


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report it rather than inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_sample_rays(wl):
    """Bounded ray sample of the workload for the CPU legs (brute force over T triangles)."""
    T = {"shoebox": 12, "vault": 3754, "room4k": 4400, "hall": 100_000, "cathedral": 1_000_000}[wl["scene"]]
    budget = 2.0e9   # ray-triangle tests per step: a few seconds on ~16 cores
    return int(max(16, min(wl["rays"], budget / (2.3 * wl["refl"] * T))))


def cpu_pipeline(wl, tri, vtx, srf, dirs, job):
    """The reference's CLI pipeline (cmd/main.cpp:241-336) on the CPU oracle."""
    from oracle import oracle as orc
    mic, src = job
    return orc.render_ir(tri, vtx, srf, mic, src, dirs, wl["refl"], **render_kwargs(wl))


def cpu_leg(wl, tri, vtx, srf, steps, warmup):
    from oracle import oracle as orc
    from rayverb_b200 import scenes
    orc.build()
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its ranks)
    try:
        orc.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        orc.set_num_threads(os.cpu_count() or 1)
    n = cpu_sample_rays(wl)
    dirs = scenes.sphere_directions(n, seed=wl["seed"])
    job = wl["jobs"][0]
    for _ in range(warmup):
        cpu_pipeline(wl, tri, vtx, srf, dirs[:max(1, n // 8)], job)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_pipeline(wl, tri, vtx, srf, dirs, job)
    dt = (time.perf_counter() - t0) / steps
    sample = (f"{n} rays x {wl['refl']} reflections of the workload per step, full pipeline "
              f"(trace+dedupe+attenuate+flatten+filter), {dt:.2f} s")
    return n * wl["refl"] / dt, dt, orc.num_threads(), sample


def run_reference(args, wl_name, wl):
    """--impl reference: the CPU restatement of the reference's algorithm on the host cores."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    tri, vtx, srf = build_scene(wl["scene"])
    value, dt, cores, sample = cpu_leg(wl, tri, vtx, srf, args.steps, args.warmup)
    line = {"impl": "reference", "metric": "rays_x_bounces_per_s", "value": value, "unit": "rays*bounces/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong" if wl["scaling"] == "strong" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl_name,
                       "note": "CPU oracle port of rayverb/kernel.cpp + host pipeline (OpenMP over rays); the reference's OpenCL build has no ICD here"},
            "cpu_baseline": {"value": value, "unit": "rays*bounces/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rays*bounces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def opencl_extra(wl, tri, vtx, srf, all_dirs, nrays, nrefl):
    """Same-hardware context: the reference's verbatim OpenCL `raytrace` kernel on this B200 through
    NVIDIA's OpenCL ICD (kernel only: no host pipeline, no flatten / filter).  Its brute-force loop
    limits this to small scenes."""
    try:
        from oracle import cl_reference as clr
        if not clr.available() or len(tri) > 5000:
            return None
        n = int(min(nrays, (1 << 30) // (64 * nrefl), 1 << 17))
        mic0, src0 = wl["jobs"][0]
        clr.raytrace(tri, vtx, srf, mic0, src0, all_dirs[:4096], nrefl, read_back=False)        # JIT + warm-up
        secs = min(clr.raytrace(tri, vtx, srf, mic0, src0, all_dirs[:n], nrefl, read_back=False)["kernel_seconds"] for _ in range(3))
        return {"value": n * nrefl / secs, "unit": "rays*bounces/s (raytrace kernel only)",
                "sample": f"{n} rays x {nrefl} reflections, best of 3, {secs * 1e3:.1f} ms",
                "note": "verbatim KERNEL_STRING of rayverb/kernel.cpp compiled by the driver's OpenCL 3.0 ICD"}
    except Exception as e:   # never let the optional extra break the bench line
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def _mean(k, ss):
    return float(np.mean([s[k] for s in ss])) if ss else 0.0


def measure(wl_name, wl, args, env, steps, warmup, parity=False):
    """Warm up, then time `steps` passes of workload `wl` on this rank's share (directions resident,
    then from host buffers).  Returns the rank-0 result block (None on other ranks)."""
    import torch
    import torch.distributed as dist
    from rayverb_b200 import api, scenes, sharding

    world, rank, local, stream = env["world"], env["rank"], env["local"], env["stream"]
    tri, vtx, srf = build_scene(wl["scene"])
    nrefl, scaling = wl["refl"], wl["scaling"]
    if scaling == "weak":
        total_rays = world * wl["rays"]
        lo, hi = rank * wl["rays"], (rank + 1) * wl["rays"]
        jobs = wl["jobs"]
    elif scaling == "strong":
        total_rays = wl["rays"]
        lo, hi = sharding.shard_slice(total_rays, rank, world)
        jobs = wl["jobs"]
    else:   # independent pairs: every rank traces ALL rays of its own pairs, no collective
        total_rays = wl["rays"]
        lo, hi = 0, total_rays
        jobs = wl["jobs"][rank::world]
    sharded = world > 1 and scaling != "jobs"
    all_dirs = scenes.sphere_directions(total_rays, seed=wl["seed"])
    shard = {}

    def set_shard(lo_, hi_):
        """(Re)bind this rank's contiguous ray slice: pinned host copy + device copy."""
        shard["lo"], shard["hi"], shard["n"] = lo_, hi_, hi_ - lo_
        shard["h"] = torch.from_numpy(all_dirs[lo_:hi_].copy()).pin_memory()
        shard["d"] = shard["h"].to("cuda", non_blocking=False)

    set_shard(lo, hi)
    accel = args.accel if args.accel is not None else wl.get("accel", 0)
    rt = api.Raytracer(nrefl, tri, vtx, srf, accel=accel)
    render_kw = render_kwargs(wl)
    cfg, keep = api.make_render_config(**render_kw)
    nch = 2 if "hrtf" in wl else len(wl["speakers"])
    state = {"trace": [], "render": []}

    # C5: independent pairs are small problems (16 k rays); several host threads, each with its own
    # CUDA stream and Raytracer, keep the GPU full.  The library's stream / error / stats state is
    # per thread, device blocks come from a mutex-protected cache.
    workers, pool = [], None
    if scaling == "jobs" and len(jobs) > 1:
        from concurrent.futures import ThreadPoolExecutor
        nworkers = min(int(os.environ.get("RVB_BENCH_WORKERS", "4")), len(jobs))
        pool = ThreadPoolExecutor(max_workers=nworkers)
        workers = [dict(stream=torch.cuda.Stream()) for _ in range(nworkers)]

    def render_on_worker(w, traces, ready):
        """Render this worker's share of a batched trace on its own stream."""
        out = []
        torch.cuda.set_device(local)
        api.set_device(local)
        api.set_stream(w["stream"].cuda_stream)
        w["stream"].wait_event(ready)
        for tr in traces:
            st_t = tr.stats()
            ir = tr.render_ir(**render_kw)
            out.append((st_t, api.render_stats(), ir))
            tr.close()
        w["stream"].synchronize()
        return out

    def step(resident, keep_bands=False):
        """One pass of the hot path over this rank's work.  Returns the IR(s) this rank produced."""
        irs = []
        if workers:
            # C5: ONE batched trace of all this rank's pairs (rvb_trace_run_batch), then the per-pair renders
            # (small sorts / filters) on the worker threads' streams
            if resident:
                d_ptr = shard["d"].data_ptr()
            else:   # host buffers: the directions go to the device inside the timed region
                state["e2e_dirs"] = shard["h"].to("cuda", non_blocking=True)
                d_ptr = state["e2e_dirs"].data_ptr()
            traces = rt.raytrace_batch(jobs, device_ptr=d_ptr, nrays=shard["n"])
            ready = torch.cuda.Event()
            ready.record(stream)
            futs = [pool.submit(render_on_worker, w, traces[i::len(workers)], ready) for i, w in enumerate(workers)]
            for f in futs:
                for st_t, st_r, ir in f.result():
                    state["trace"].append(st_t)
                    state["render"].append(st_r)
                    irs.append(ir)
            return irs
        for mic, src in jobs:
            t_host = time.perf_counter()
            if resident:
                tr = rt.raytrace(mic, src, None, ray_offset=shard["lo"], device_ptr=shard["d"].data_ptr(), nrays=shard["n"])
            else:
                tr = rt.raytrace(mic, src, shard["h"].numpy(), ray_offset=shard["lo"])
            state["trace"].append(tr.stats())
            timing = {"local_ms": 1e3 * (time.perf_counter() - t_host)}
            if not sharded:
                irs.append(tr.render_ir(keep_bands=keep_bands, **render_kw))
                state["render"].append(api.render_stats())
                continue
            # sharded: ONE all-gather (extrema + image sources), ONE NCCL reduce of the histogram, rank 0 filters
            hist, lengths, _ = sharding.sharded_histogram(tr, cfg, rank, world, "cuda", SR, False, timing)
            state["render"].append(api.render_stats())
            if rank == 0:
                # (the finish follows the last synchronisation point of the step -- the reduce -- so it is not
                # part of what the shard re-cut can balance: only rank 0's work BEFORE the reduce counts)
                irs.append(api.render_finish(cfg, hist.data_ptr(), lengths, keep_bands=keep_bands))
            state["timing"] = timing
        return irs

    def timed(resident, nsteps):
        state["trace"].clear()
        state["render"].clear()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0, d2h = api.launch_count(), 0
        e0.record(stream)
        for _ in range(nsteps):
            irs = step(resident)
            d2h = sum(c.nbytes for ir in irs for c in ir.channels)
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / nsteps
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, api.launch_count() - launches0, d2h

    # Warm-up; on a sharded run with several waves of rays per GPU the second warm-up step also MEASURES
    # the ranks: rank 0 merges the image sources on top of its share of the rays, so the contiguous ray
    # slices are re-cut once so that every rank reaches the reduce at the same time
    # (sharding.balance_weights; rank 0's filter + copy-out FOLLOW the reduce and cannot be balanced away:
    # counting them, as the first version did, took 7.5 k rays off rank 0 at 2 GPUs and made rank 1 the
    # last to arrive).  Small shards stay equal.
    balance = None
    for w_it in range(max(warmup, 3)):
        step(True)
        if sharded and w_it == 1 and not args.no_balance and shard["n"] >= (1 << 19):
            tm = state.get("timing", {})
            mine = torch.tensor([tm.get("local_ms", 0.0), tm.get("exclusive_ms", 0.0), float(shard["n"])], device="cuda", dtype=torch.float64)
            allm = torch.empty((world, 3), device="cuda", dtype=torch.float64)
            dist.all_gather_into_tensor(allm, mine.unsqueeze(0))
            allm = allm.cpu().numpy()
            ms_per_ray = float(allm[:, 0].sum() / max(1.0, allm[:, 2].sum()))
            weights = sharding.balance_weights(ms_per_ray, allm[:, 1], total_rays)
            cuts = sharding.shard_bounds(total_rays, world, weights)
            set_shard(cuts[rank], cuts[rank + 1])
            balance = {"rank0_share": (cuts[1] - cuts[0]) / total_rays, "equal_share": 1.0 / world,
                       "measured_exclusive_ms_rank0": float(allm[0, 1]), "measured_ms_per_kiloray": 1e3 * ms_per_ray}
    ms, launches, _ = timed(True, steps)
    tstats = [dict(s) for s in state["trace"]]
    rstats = [dict(s) for s in state["render"]]
    step(False)  # warm the host-buffer path (pinned H2D)
    e2e_ms, _, d2h_bytes = timed(False, steps)
    nrays = shard["n"]   # after the shard re-cut
    h2d_bytes = shard["h"].numel() * 4 * len(jobs)

    # sharded parity: the same total ray set on ONE GPU (rank 0) against the sharded result, per band
    sharded_parity = None
    if parity and sharded:
        irs = step(True, keep_bands=True)
        if rank == 0:
            mic, src = jobs[0]
            full = torch.from_numpy(all_dirs).to("cuda")
            tr1 = rt.raytrace(mic, src, None, ray_offset=0, device_ptr=full.data_ptr(), nrays=total_rays)
            one = tr1.render_ir(keep_bands=True, **render_kw)
            worst, where = 0.0, None
            same_len = all(len(a) == len(b) for a, b in zip(irs[0].channels, one.channels))
            for c in range(len(one.bands)):
                for b in range(one.bands[c].shape[0]):
                    ref = one.bands[c][b].astype(np.float64)
                    got = irs[0].bands[c][b].astype(np.float64)
                    if ref.shape != got.shape:
                        worst, where = float("inf"), (c, b)
                        continue
                    d = np.linalg.norm(ref)
                    e = float(np.linalg.norm(got - ref) / d) if d > 0 else float(np.linalg.norm(got))
                    if e > worst:
                        worst, where = e, (c, b)
            chan = max(float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-300))
                       for a, b in zip(irs[0].channels, one.channels)) if same_len else float("inf")
            sharded_parity = {"against": "the same %d rays rendered on one GPU (rank 0)" % total_rays,
                              "max_rel_l2_per_band": worst, "worst_channel_band": where, "max_rel_l2_per_channel": chan,
                              "tolerance": PARITY_TOL, "ok": bool(worst <= PARITY_TOL and chan <= PARITY_TOL)}
            del full
        if world > 1:
            dist.barrier()

    block = None
    if rank == 0:
        hbm_gbs, sm_max_mhz, peak_src = load_peaks()
        njobs_total = len(wl["jobs"])
        total_rb = total_rays * nrefl * (njobs_total if scaling == "jobs" else 1)
        fp32_peak_tf = SM_COUNT * FP32_LANES * 2 * sm_max_mhz * 1e6 / 1e12
        steady_ms, image_ms = _mean("steady_phase_ms", tstats), _mean("image_phase_ms", tstats)
        # counts per launch: a batched trace (C5) runs all of this rank's pairs in the same kernels
        per_launch = (lambda k: _mean(k, tstats) * len(jobs)) if scaling == "jobs" else (lambda k: _mean(k, tstats))
        steady_tests, steady_boxes = per_launch("steady_tri_tests"), per_launch("steady_box_tests")
        all_tests, all_boxes = per_launch("tri_tests"), per_launch("box_tests")
        brute = rt.accel == api.ACCEL_BRUTE_FORCE
        steady_flop = steady_tests * FLOP_PER_TEST + steady_boxes * FLOP_PER_BOX
        achieved_tf = steady_flop / (steady_ms * 1e-3) / 1e12 if steady_ms > 0 else 0.0
        split = min(9, nrefl)
        emit_gbs = nrays * (len(jobs) if scaling == "jobs" else 1) * (nrefl - split) * 52 / (steady_ms * 1e-3) / 1e9 if steady_ms > 0 else 0.0
        traffic, traffic_note = None, None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if world == 1 and scaling != "jobs" and os.path.exists(tp):
            # NOT measured in this run: DRAM bytes per ray*bounce of the steady kernel from the committed
            # ncu --set full capture (profiles/ncu_traffic.json), scaled to this run's ray*bounce count
            per_rb = json.load(open(tp)).get("steady_phase_dram_bytes_per_ray_bounce" if brute else "bvh_steady_phase_dram_bytes_per_ray_bounce")
            if per_rb:
                traffic = per_rb * nrays * (nrefl - split)
                traffic_note = "profile-derived constant (profiles/ncu_traffic.json) x this run's rays*bounces, not measured live"
        both_ms = steady_ms + image_ms
        bounces, emitted = per_launch("bounces"), per_launch("emitted")
        block = {
            "metric": "rays_x_bounces_per_s", "value": total_rb / (ms * 1e-3), "unit": "rays*bounces/s",
            "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong" if scaling == "strong" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl_name, "triangles": int(len(tri)), "rays_total": int(total_rays), "rays_this_gpu": int(nrays),
                       "reflections": nrefl, "source_receiver_pairs": njobs_total, "channels": nch, "sample_rate": int(SR),
                       "attenuation": "hrtf" if "hrtf" in wl else "speakers", "filter": "linkwitz_riley",
                       "accel": "brute_force_smem" if brute else "bvh+micmap",
                       "sharding": ("pairs/%d" % world if scaling == "jobs" else "rays/%d" % world) if world > 1 else "none",
                       "shard_balance": balance,
                       "l2": "inputs_larger_than_L2 (per-step working set %.2f GB per GPU)" % (nrays * nrefl * 52 / 1e9)},
            "e2e": {"value": total_rb / (e2e_ms * 1e-3), "unit": "rays*bounces/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes)},
            "ir_wall_ms": e2e_ms / max(1, len(jobs)),
            "gpu_launches": int(launches),
            # executed reflections on rank 0's shard (a ray that leaves the mesh stops, kernel.cpp:372-375) and
            # the share of them whose impulse the attenuators skip (zero volume, kernel.cpp:524,607)
            "bounces": {"executed": int(bounces), "of": int(nrays * nrefl * max(1, len(jobs) if scaling == "jobs" else 1)),
                        "emitted": int(emitted), "skipped_fraction": (1.0 - emitted / bounces) if bounces > 0 else None},
            "roofline": {"kernel": ("trace_bruteforce_kernel<false>" if brute else "trace_bvh_wavefront_kernel")
                         + " (steady phase, reflections %d..%d)" % (split, nrefl - 1),
                         "bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp32_peak_tf, "traffic": traffic, "traffic_note": traffic_note,
                         "peak_source": f"148 SM x 128 lanes x 2 x {sm_max_mhz:.0f} MHz ({peak_src} clock)",
                         "kernel_ms": steady_ms, "tests_per_launch": int(steady_tests), "flop_per_test": FLOP_PER_TEST,
                         "all_bounce_kernels": {"ms": both_ms, "tests": int(all_tests), "box_tests": int(all_boxes),
                                                "achieved": (all_tests * FLOP_PER_TEST + all_boxes * FLOP_PER_BOX) / (both_ms * 1e-3) / 1e12 if both_ms > 0 else 0.0},
                         "hbm_view": {"achieved": emit_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": emit_gbs / hbm_gbs,
                                      "bytes_per_impulse": 52}},
            "stages_ms": {"image_phase": image_ms, "steady_phase": steady_ms, "trace_total": _mean("total_ms", tstats),
                          "keys": _mean("keys_ms", rstats), "sort": _mean("sort_ms", rstats), "reduce": _mean("reduce_ms", rstats),
                          "filter": _mean("filter_ms", rstats), "finish": _mean("finish_ms", rstats)},
        }
        if not brute:
            block["roofline"]["box_tests_per_launch"] = int(steady_boxes)
            block["roofline"]["flop_per_box_test"] = FLOP_PER_BOX
            block["roofline"]["flop_model"] = "tri_tests x 52 + box_tests x 20 (SURVEY 8d); visibility queries go through the microphone map"
            block["roofline"]["binding_resource"] = ("L1 data pipe and issue slots: ncu l1tex__data_pipe_lsu_wavefronts 70 % of peak, issue active 58 %, "
                                                     "20.1 of 32 lanes per instruction on this kernel (profiles/r2_c3_shard_kernels_full.md) -- "
                                                     "divergent node fetches, not FP32 or HBM")
            block["roofline"]["micmap"] = {"cells_per_face_edge": int(_mean("micmap_cells", tstats)), "entries": int(_mean("micmap_entries", tstats))}
            try:   # what ncu measured on the binding resource (committed capture, not this run)
                block["roofline"]["ncu_utilisation"] = json.load(open(tp)).get("bvh_ncu_utilisation")
            except (OSError, ValueError):
                pass
        if sharded_parity is not None:
            block["sharded_parity"] = sharded_parity
    if pool is not None:
        pool.shutdown()
    rt.close()
    del shard["d"], shard["h"]
    return block, (wl, tri, vtx, srf, all_dirs, nrays, nrefl)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--rays", type=int, default=None, help="override the workload's ray count (testing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-blocks", action="store_true", help="N = 1: skip the C2 / brute-force-room sub-blocks")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded-vs-one-GPU IR comparison")
    ap.add_argument("--no-balance", action="store_true", help="multi-GPU: keep equal ray shards (no measured re-cut)")
    ap.add_argument("--accel", type=int, default=None, help="0 auto, 1 brute force, 2 BVH")
    args = ap.parse_args()
    wl_name = args.workload
    wl = dict(WORKLOADS[wl_name])
    if args.rays:
        wl["rays"] = args.rays
    if args.impl == "reference":
        return run_reference(args, wl_name, wl)

    import torch
    import torch.distributed as dist
    from rayverb_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the rayverb_b200 hot path has no CPU fallback")
    torch.cuda.set_device(local)
    api.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    api.set_stream(stream.cuda_stream)
    env = dict(world=world, rank=rank, local=local, stream=stream)

    sampler = ClockSampler(local)
    sampler.start()
    line, (wl, tri, vtx, srf, all_dirs, nrays, nrefl) = measure(wl_name, wl, args, env, args.steps, args.warmup, parity=not args.no_parity)
    clocks = sampler.stop()

    subs, sub_ctx = {}, {}
    if world == 1 and not args.no_sub_blocks and wl_name == DEFAULT_WORKLOAD and not args.rays:
        for name in SUB_BLOCKS:
            sub_args = argparse.Namespace(**vars(args))
            sub_args.accel = None
            s = ClockSampler(local)
            s.start()
            blk, ctx = measure(name, dict(WORKLOADS[name]), sub_args, env, args.steps, args.warmup)
            sub_ctx[name] = (blk, ctx)
            blk["clocks"] = s.stop()
            for k in ("metric", "unit", "n_gpus", "higher_is_better", "vs_baseline", "dtype", "data"):
                blk.pop(k, None)
            subs[name] = blk

    if rank == 0:
        line["clocks"] = clocks
        if subs:
            line["sub_blocks"] = subs
            c2 = subs.get("C2_shoebox_12tri_1Mx128_omni")
            if c2:   # the FP32 intersection roofline the north star states its 50 % target for
                line["intersection_roofline"] = {"workload": "C2_shoebox_12tri_1Mx128_omni (sub_blocks)", **{k: c2["roofline"][k] for k in
                                                 ("kernel", "bound", "achieved", "peak", "unit", "frac", "kernel_ms", "tests_per_launch")}}
        if not args.no_cpu_baseline:
            v, dt, cores, sample = cpu_leg(wl, tri, vtx, srf, 1, 0)
            line["cpu_baseline"] = {"value": v, "unit": "rays*bounces/s", "cores": cores, "kind": "port", "sample": sample}
            for name, (blk, ctx) in sub_ctx.items():   # extra context on the small scenes (see opencl_extra)
                extra = opencl_extra(*ctx)
                if extra is not None:
                    line["sub_blocks"][name]["reference_opencl_kernel_on_b200"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
