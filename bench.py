#!/usr/bin/env python
"""bench.py -- rays x bounces / s of the rayverb hot path on B200 (BASELINE.json metric).

A "step" is one pass of the whole hot path over one batch of synthetic rays:
  bounce kernel (nearest hit, specular bounces, diffuse + image-source emission, 8-band air
  attenuation) -> image-source dedupe -> speaker attenuation -> deterministic flatten (radix sort
  by sample index + segmented reduction) -> Linkwitz-Riley band filters -> mixdown/normalize/trim.

  value : whole-job rays*bounces/s with the directions already resident in HBM (CUDA events on
          the launching stream, max over ranks)
  e2e   : the same metric through the public C-ABI calls with HOST buffers: pinned-host
          directions -> device inside the timed region, final IR read back to the host
  N > 1 : rays are sharded (weak scaling: every GPU traces the full per-GPU ray count of its own
          contiguous slice of the direction set); per-GPU histograms are summed with one NCCL
          reduce, rank 0 filters.

`--impl reference` times the CPU restatement of the reference's own algorithm (oracle/) on the
host cores -- the reference's OpenCL build cannot run in this image (no OpenCL ICD).
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "parallel-reverb-raytracer_b200", "python"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FLOP_PER_TEST = 52           # Moller-Trumbore, rayverb/kernel.cpp:63-88
SM_COUNT = 148
FP32_LANES = 128


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), d.get("sm_max_mhz", 1965.0), "measured"
    return 6650.0, 1965.0, "fallback"


WORKLOADS = {
    # name: (scene builder, mic, source, rays per GPU, reflections, speakers)
    "C2_shoebox_12tri_1Mx128_omni": dict(scene="shoebox", mic=(7, 2, 4), src=(2, 1.5, 2), rays=1 << 20, refl=128,
                                         speakers=[((0, 0, 1), 0.0)], seed=2015),
    "C1_small_room_1024x64": dict(scene="room700", mic=(5.5, 1.2, 3.5), src=(2, 1.5, 2), rays=1024, refl=64,
                                  speakers=[((0, 0, -1), 0.0)], seed=2015),
    "C3_hall_100ktri_1Mx128_stereo": dict(scene="hall", mic=(24, 1.5, 18), src=(20, 2.0, 4), rays=1 << 20, refl=128,
                                          speakers=[((-1, 0, -1), 0.5), ((1, 0, -1), 0.5)], seed=2015),
}


def build_scene(name):
    from rayverb_b200 import scenes
    if name == "shoebox":
        return scenes.shoebox()
    if name == "room700":
        tri, vtx = scenes.displaced_box(8.0, 3.0, 5.0, 700, amp=0.08, fan=0.2, nsurfaces=4)
        return tri, vtx, scenes.hall_surfaces()
    if name == "hall":
        return scenes.concert_hall()
    raise KeyError(name)


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


def cpu_pipeline(wl, tri, vtx, srf, dirs):
    """The reference's CLI pipeline (cmd/main.cpp:241-336) on the CPU oracle."""
    from oracle import oracle as orc
    return orc.render_ir(tri, vtx, srf, wl["mic"], wl["src"], dirs, wl["refl"], speakers=wl["speakers"], sr=44100.0,
                         ft="linkwitz_riley", lo_cutoff=45.0)


def run_reference(args, wl_name, wl):
    """--impl reference: the CPU restatement of the reference's algorithm on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from rayverb_b200 import scenes
    orc.build()
    tri, vtx, srf = build_scene(wl["scene"])
    sample_rays = min(wl["rays"], 32768 if wl["scene"] == "shoebox" else 256)
    dirs = scenes.sphere_directions(sample_rays, seed=wl["seed"])
    for _ in range(args.warmup):
        cpu_pipeline(wl, tri, vtx, srf, dirs[:max(1, sample_rays // 8)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pipeline(wl, tri, vtx, srf, dirs)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample_rays * wl["refl"] / dt
    cores = orc.num_threads()
    sample = f"{sample_rays} rays x {wl['refl']} reflections of the workload per step, full pipeline (trace+dedupe+attenuate+flatten+filter)"
    line = {"impl": "reference", "metric": "rays_x_bounces_per_s", "value": value, "unit": "rays*bounces/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl_name, "note": "CPU oracle port of rayverb/kernel.cpp + host pipeline; the reference's OpenCL build has no ICD here"},
            "cpu_baseline": {"value": value, "unit": "rays*bounces/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "rays*bounces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2_shoebox_12tri_1Mx128_omni", choices=sorted(WORKLOADS))
    ap.add_argument("--rays", type=int, default=None, help="override rays per GPU (testing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--accel", type=int, default=0)
    args = ap.parse_args()
    wl_name = args.workload
    wl = dict(WORKLOADS[wl_name])
    if args.rays:
        wl["rays"] = args.rays
    if args.impl == "reference":
        return run_reference(args, wl_name, wl)

    import torch
    import torch.distributed as dist
    from rayverb_b200 import api, scenes, sharding

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

    tri, vtx, srf = build_scene(wl["scene"])
    nrays, nrefl = wl["rays"], wl["refl"]
    # one global direction set of world*nrays rays; rank r owns the contiguous slice [r*nrays, (r+1)*nrays)
    all_dirs = scenes.sphere_directions(world * nrays, seed=wl["seed"])
    h_dirs = torch.from_numpy(all_dirs[rank * nrays:(rank + 1) * nrays].copy()).pin_memory()
    d_dirs = h_dirs.to("cuda", non_blocking=False)
    rt = api.Raytracer(nrefl, tri, vtx, srf, accel=args.accel)
    render_kw = dict(speakers=wl["speakers"], sr=44100.0, ft="linkwitz_riley", lo_cutoff=45.0)
    cfg, keep = api.make_render_config(**render_kw)
    nch = len(wl["speakers"])
    h2d_bytes = h_dirs.numel() * 4
    state = {}

    def step(resident):
        """One pass of the hot path on this rank's ray shard.  Returns the IR on rank 0."""
        if resident:
            tr = rt.raytrace(wl["mic"], wl["src"], None, ray_offset=rank * nrays, device_ptr=d_dirs.data_ptr(), nrays=nrays)
        else:
            tr = rt.raytrace(wl["mic"], wl["src"], h_dirs.numpy(), ray_offset=rank * nrays)
        state["trace"] = tr.stats()
        if world == 1:
            ir = tr.render_ir(**render_kw)
            state["render"] = api.render_stats()
            return ir
        # ---- sharded: global predelay/lengths, cross-shard image merge, one NCCL reduce ----------
        hist, lengths, _ = sharding.sharded_histogram(tr, cfg, rank, world, "cuda", 44100.0, False)
        state["render"] = api.render_stats()
        if rank == 0:
            return api.render_finish(cfg, hist.data_ptr(), lengths)
        return None

    def timed(resident, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        kern_ms, launches0, d2h, tests = [], api.launch_count(), 0, 0
        for _ in range(steps):
            ir = step(resident)
            kern_ms.append(state["trace"]["trace_kernel_ms"])
            tests = state["trace"]["tri_tests"]
            if ir is not None:
                d2h = sum(c.nbytes for c in ir.channels)
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, float(np.mean(kern_ms)), api.launch_count() - launches0, d2h, tests

    for _ in range(max(args.warmup, 3)):
        step(True)
    sampler = ClockSampler(local)
    sampler.start()
    ms, kern_ms, launches, d2h_bytes, tests = timed(True, args.steps)
    render = dict(state["render"])
    trace_stats = dict(state["trace"])
    step(False)  # warm the host-buffer path (pinned H2D)
    e2e_ms, _, _, d2h_bytes, _ = timed(False, args.steps)
    clocks = sampler.stop()

    if rank == 0:
        hbm_gbs, sm_max_mhz, peak_src = load_peaks()
        total_rb = world * nrays * nrefl
        fp32_peak_tf = SM_COUNT * FP32_LANES * 2 * sm_max_mhz * 1e6 / 1e12
        achieved_tf = tests * FLOP_PER_TEST / (kern_ms * 1e-3) / 1e12
        # emission bytes of the bounce kernel: 48 useful bytes per impulse + 4 B sort key time
        emit_gbs = nrays * nrefl * 52 / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "rays_x_bounces_per_s", "value": total_rb / (ms * 1e-3), "unit": "rays*bounces/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl_name, "triangles": int(len(tri)), "rays_per_gpu": nrays, "reflections": nrefl,
                       "channels": nch, "sample_rate": 44100, "filter": "linkwitz_riley",
                       "accel": "brute_force_smem" if rt.accel == api.ACCEL_BRUTE_FORCE else "bvh",
                       "sharding": f"rays/{world}" if world > 1 else "none",
                       "l2": "inputs_larger_than_L2 (per-step working set %.1f GB)" % (nrays * nrefl * 52 / 1e9)},
            "e2e": {"value": total_rb / (e2e_ms * 1e-3), "unit": "rays*bounces/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": int(d2h_bytes)},
            "ir_wall_ms": e2e_ms,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "trace_bruteforce_kernel" if rt.accel == api.ACCEL_BRUTE_FORCE else "trace_bvh_kernel",
                         "bound": "fp32", "achieved": achieved_tf, "peak": fp32_peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp32_peak_tf, "traffic": None,
                         "peak_source": f"148 SM x 128 lanes x 2 x {sm_max_mhz:.0f} MHz ({peak_src} clock)",
                         "kernel_ms": kern_ms, "tests_per_launch": int(tests), "flop_per_test": FLOP_PER_TEST,
                         "hbm_view": {"achieved": emit_gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": emit_gbs / hbm_gbs,
                                      "bytes_per_impulse": 52}},
            "stages_ms": {"trace_kernel": kern_ms, "trace_total": trace_stats["total_ms"], "keys": render["keys_ms"],
                          "sort": render["sort_ms"], "reduce": render["reduce_ms"], "filter": render["filter_ms"],
                          "finish": render["finish_ms"]},
        }
        if not args.no_cpu_baseline:
            from oracle import oracle as orc
            orc.build()
            sample_rays = min(nrays, 32768 if wl["scene"] == "shoebox" else 256)
            d = all_dirs[:sample_rays]
            t0 = time.perf_counter()
            cpu_pipeline(wl, tri, vtx, srf, d)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample_rays * nrefl / dt, "unit": "rays*bounces/s", "cores": orc.num_threads(),
                                    "kind": "port",
                                    "sample": f"{sample_rays} rays x {nrefl} reflections of the same workload, full pipeline, {dt:.1f} s"}
        print(json.dumps(line))
    rt.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
