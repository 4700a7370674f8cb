#!/usr/bin/env python
"""bench.py — headline benchmark of the render hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl own|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one frame of the headline workload of BASELINE.json
("1080p 16spp 4-bounce": synthetic 64x64 skin seed 0, standing pose, 1920x1080,
16 spp, 4 bounces, reference defaults otherwise: soft shadows x8, gradient
background, 32x32 tiles).  Metric: M unique rays/s, where the unique-ray count of the
frame is the fixed integer the reference algorithm performs (intersectScene calls
minus the redundant re-test, counted by the CPU oracle: tests/golden/work_counts.json).

value    : frame time measured with CUDA events, scene resident on the device, output
           left in device memory (for N>1 the NCCL band gather to rank 0 is inside the step).
e2e      : the same frame through the public host API (host scene in, host image out,
           host<->device copies inside the timed region).
roofline : FP32 issue roofline: algorithmic lane-ops of the frame /
           event time / (SMs x 128 lanes x max SM clock).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = "headline_1080p_16spp_4b"
CPU_SAMPLE = "headline_1080p_4spp_4b"      # cpu_baseline leg of the own arm: one frame at 4 of the 16 spp
REF_STEP_SAMPLE = "headline_1080p_2spp_4b"  # --impl reference: each step is one frame at 2 of the 16 spp


def load_counts():
    return json.loads((ROOT / "tests" / "golden" / "work_counts.json").read_text())


def alg_ops(entry: dict) -> dict:
    """Algorithmic FP32 lane-ops of one frame (SURVEY.md §8d work model, DESIGN.md "Work model")."""
    c = entry["counters"]
    cfg = entry["config"]
    P, R = entry["n_boxes_plain"], entry["n_boxes_rotated"]
    S = c["n_primary_rays"]
    rays = entry["unique_rays"]
    n_shadow_samples = 8  # reference default (raytracer.h:19); the workloads do not override it
    per_ray = 6 + 28 * P + 74 * R
    primary = 35 * S + S * per_ray + 21 * c["n_background_primary"] + (4 * cfg["samples_per_pixel"] + 5) * cfg["width"] * cfg["height"]
    secondary = ((rays - S) * per_ray + 26 * (c["n_slab_pass"] + c["n_backface_eval"]) + 48 * c["n_rotated_hits"]
                 + 118 * c["n_shade"] + (60 + 42 * n_shadow_samples) * c["n_soft_shadow"] + 20 * c["n_hard_shadow"]
                 + 68 * c["n_reflect_rays"])
    return {"total": primary + secondary, "primary_pass": primary, "shade_pass": secondary}


def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text())
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "_fallback": True}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


# ------------------------------------------------------------------------- reference arm
def cpu_reference_run(entry_name: str, steps: int, warmup: int):
    """Times the reference's own CPU renderer (oracle/_ref when it was built, else the C port)."""
    from minecraftskin_raytracer_b200 import _abi
    from minecraftskin_raytracer_b200.scene import synth_skin
    from oracle.harness import Oracle, Reference
    counts = load_counts()
    entry = counts[entry_name]
    ref = Reference.load()
    kind = "reference" if ref is not None else "port"
    runner = ref if ref is not None else Oracle()
    cores = runner.hardware_threads()
    cfg = _abi.default_config(**entry["config"])
    # scene through the checker's own builder when it has one (reference), else the product's host builder
    if ref is not None:
        scene = ref.scene_from_atlas(synth_skin(entry["skin_seed"], entry["skin_kind"]), entry["pose"])
    else:
        from minecraftskin_raytracer_b200 import lib
        scene = lib.build_skin_scene(synth_skin(entry["skin_seed"], entry["skin_kind"]), entry["pose"])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        runner.render(scene, cfg)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    sec = sum(times) / len(times)
    return {"value": entry["unique_rays"] / sec / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
            "sample": f"{entry_name}: full 1920x1080 frame at {entry['config']['samples_per_pixel']} of the 16 spp "
                      f"({entry['unique_rays']} unique rays) through TileRenderer::render, threadCount=0, "
                      f"mean of {len(times)} run(s)",
            "ms_per_sample": sec * 1e3}


def run_reference_arm(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    res = cpu_reference_run(REF_STEP_SAMPLE, args.steps, args.warmup)
    counts = load_counts()
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": res["value"], "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_sample"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step_sample": REF_STEP_SAMPLE,
                   "unique_rays_per_step": counts[REF_STEP_SAMPLE]["unique_rays"]},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- own arm
def run_own_arm(args):
    import torch
    import torch.distributed as dist

    from minecraftskin_raytracer_b200 import _abi, build
    build.build()
    from minecraftskin_raytracer_b200 import bands, lib
    from minecraftskin_raytracer_b200.scene import synth_skin

    rank, local_rank, world = dist_env()
    # tuning aid: time ONE rank's share of an N-way split on a single GPU (no exchange): MCSKIN_BENCH_SPLIT=N
    emulate = int(os.environ.get("MCSKIN_BENCH_SPLIT", "0")) if world == 1 else 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    counts = load_counts()
    entry = counts[WORKLOAD]
    cfg = _abi.default_config(**entry["config"])
    W, H, ts = cfg.width, cfg.height, cfg.tile_size
    scene = lib.build_skin_scene(synth_skin(entry["skin_seed"], entry["skin_kind"]), entry["pose"])
    ops = alg_ops(entry)
    unique_rays = entry["unique_rays"]

    ctx = lib.Context(local_rank)
    lanes_default = int(os.environ.get("MCSKIN_FRAME_LANES", "0")) or lib.DEFAULT_FRAME_LANES
    ctx.set_scene(scene, cfg)
    split = emulate if emulate > 1 else world
    max_rows = bands.padded_band_rows(H, ts, split)  # padded band height, equal on all ranks
    assert ctx.band_rows(rank, world) == bands.band_pixel_rows(H, ts, rank, world)
    band = torch.zeros((max_rows, W, 4), dtype=torch.float32, device=dev)
    band_u8 = torch.zeros((max_rows, W, 4), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gathered = torch.empty((world * max_rows, W, 4), dtype=torch.float32, device=dev) if (world > 1 and rank == 0) else None
    row_index = bands.frame_row_index(H, ts, world, max_rows, dev) if rank == 0 else None
    frame = torch.zeros((H, W, 4), dtype=torch.float32, device=dev) if rank == 0 else None
    # a dedicated (non-default) stream: handle 0 would mean "the context's own stream" to the C ABI,
    # and torch.cuda.Event only sees the stream it is recorded on
    stream = torch.cuda.Stream(dev)
    torch.cuda.set_stream(stream)

    # N>1: every rank stores its tile rows straight into rank 0's frame over NVLink peer memory and the
    # exchange step is a barrier ("p2p", default); or rank 0 gathers compact bands with NCCL ("gather").
    exchange = os.environ.get("MCSKIN_EXCHANGE", "p2p") if world > 1 else "none"
    peer = None
    if exchange == "p2p":
        try:
            peer = bands.PeerFrame(lib, H, W, local_rank)
            frame_u8 = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)  # this rank's rows, quantised (stays local)
        except Exception as exc:  # noqa: BLE001  (no peer access between these devices)
            print(f"bench.py: peer frame unavailable ({exc}); falling back to the NCCL gather", file=sys.stderr)
            exchange, peer = "gather", None
        ok = torch.tensor([1.0 if peer is not None else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() < 1.0:
            exchange, peer = "gather", None

    def step_gather():
        ctx.render_bands(rank, split, band.data_ptr(), band_u8.data_ptr(), stream.cuda_stream)
        if world > 1:
            # gather the bands on rank 0 (NCCL over NVLink), rows back in order
            bands.gather_frame(band, frame, ts, gathered, row_index)

    def step_p2p():
        ctx.render_rows_into_frame(rank, world, peer.ptr, frame_u8.data_ptr(), stream.cuda_stream)
        peer.fence(stream.cuda_stream)  # ranks release a flag in the root's memory; the root's stream acquires them

    step = step_p2p if exchange == "p2p" else step_gather
    if exchange == "p2p":
        # one frame each way: the peer-written frame must equal the gathered one bit for bit
        step_gather()
        step_p2p()
        torch.cuda.synchronize(dev)
        dist.barrier()
        if rank == 0 and not torch.equal(peer.frame.view(torch.int32), frame.view(torch.int32)):
            raise SystemExit("bench.py: peer-written frame differs from the gathered frame")

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches_per_step = 0
    for _ in range(args.warmup):
        flush.fill_(1)
        step()
    torch.cuda.synchronize(dev)
    launches_per_step = ctx.sync()["n_kernel_launches"]

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    pass_ms = {"primary": 0.0, "shade": 0.0, "device": 0.0}
    for i in range(args.steps):
        flush.fill_(i & 0xff)  # L2 flush between timed iterations (outside the timed events)
        starts[i].record(stream)
        step()
        ends[i].record(stream)
        st = ctx.sync()  # the library's own events around this frame's launches (blocks on the frame; the next flush follows anyway)
        pass_ms["device"] += st["ms_device"]
        n_active = st["n_active_pixels"]
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    # Pass breakdown, outside the timed region: the same frame on ONE stream with direct launches
    # (the timed frames interleave several lanes inside a CUDA graph, where pass times overlap).
    ctx.set_option("use_graphs", 0)
    ctx.set_option("frame_lanes", 1)
    n_serial = max(1, min(args.steps, 5))
    for i in range(1 + n_serial):
        flush.fill_(i & 0xff)
        ctx.render_bands(rank, split, band.data_ptr(), band_u8.data_ptr(), stream.cuda_stream)
        st = ctx.sync()
        if i > 0:
            pass_ms["primary"] += st["ms_primary"] / n_serial
            pass_ms["shade"] += st["ms_shade"] / n_serial
    ctx.set_option("use_graphs", 1)
    ctx.set_option("frame_lanes", lanes_default)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps

    # ---- e2e: host scene in, host image out, copies inside the timed region
    h2d_bytes = scene.boxes.nbytes + scene.texels.nbytes + C.sizeof(_abi.McScene) + C.sizeof(_abi.McConfig)
    d2h_bytes = H * W * 16
    if args.kernel_only:
        e2e_ms = None
    elif world == 1:
        # the public host call: host scene in (re-flattened, re-uploaded every step), host float image
        # out, into a page-locked result buffer the caller reuses from frame to frame
        host_img = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True).numpy()
        for _ in range(max(1, min(args.warmup, 3))):
            lib.render(scene, cfg, device=local_rank, out_f32=host_img)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            lib.render(scene, cfg, device=local_rank, out_f32=host_img)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    else:
        host = torch.empty((H, W, 4), dtype=torch.float32, pin_memory=True) if rank == 0 else None

        def e2e_step():
            ctx.set_scene(scene, cfg)  # host -> device on every rank
            step()
            if rank == 0:
                host.copy_(peer.frame if exchange == "p2p" else frame, non_blocking=True)
            if exchange == "p2p":
                peer.fence_all()  # the peers may overwrite the root's frame only after it has left for the host
            torch.cuda.synchronize(dev)

        for _ in range(2):
            e2e_step()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        dist.barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
        h2d_bytes *= world

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    fp32_peak = sm_count * 128 * peaks["sm_max_mhz"] * 1e6 / 1e12  # T lane-ops/s, non-FMA issue rate
    try:
        fp32_peak_measured = lib.fp32_issue_peak(local_rank) / 1e12  # FADD/FMUL microbenchmark on this device
    except Exception:  # noqa: BLE001
        fp32_peak_measured = None
    # The frame is one pipeline of ~14 short kernels per lane, several lanes in flight at once, so
    # the roofline entry is that of the whole step: the frame's algorithmic lane-ops over the
    # event time of a frame's launches.  At N>1 each rank does ~1/N of the frame.
    device_ms = pass_ms["device"] / args.steps
    step_ops = ops["total"] / world
    achieved = step_ops / (device_ms * 1e-3) / 1e12 if device_ms > 0 else 0.0
    fb_bytes = W * H * (16 + 4)
    line = {
        "metric": "Mrays/s", "value": unique_rays / (ms_per_step * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "width": W, "height": H, "spp": cfg.samples_per_pixel,
                   "max_bounces": cfg.max_bounces, "shadow_samples": cfg.shadow_samples, "tile_size": ts,
                   "skin": "synthetic 64x64 seed 0", "unique_rays_per_frame": unique_rays,
                   "partition": "whole frame" if world == 1 else (
                       f"interleaved tile rows over {world} GPUs, stored into rank 0's frame over NVLink peer memory + barrier"
                       if exchange == "p2p" else f"interleaved tile rows over {world} GPUs + NCCL gather"),
                   "l2": "flushed between timed iterations (256 MiB fill outside the timed events)",
                   "frame_lanes": lanes_default, "launch": "CUDA graph replay of the frame's kernels"},
        "clocks": clocks,
        "e2e": {"value": unique_rays / (e2e_ms * 1e-3) / 1e6 if e2e_ms else None, "unit": "Mrays/s", "ms_per_frame": e2e_ms,
                "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes)},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": {
            "bound": "fp32", "kernel": f"frame pipeline ({launches_per_step} kernel launches: primary pass + wavefront shading)",
            "achieved": achieved, "peak": fp32_peak, "unit": "Tlane-op/s",
            "frac": achieved / fp32_peak,
            # DRAM bytes of one frame's kernels (dram__bytes_read.sum + dram__bytes_write.sum over the 16 launches of a
            # serial frame, ncu --set full): queue traffic, ~12 % of HBM bandwidth at this frame time
            "traffic": 0.93e9 if world == 1 else None, "traffic_source": "profiles/r01/final_ncu_summary.md",
            "peak_source": f"{sm_count} SMs x 128 FP32 lanes x {peaks['sm_max_mhz']:.0f} MHz (sm_max_mhz of MEASURED_PEAKS.json"
                           + (", fallback" if peaks.get("_fallback") else "") + "); non-FMA issue rate, SURVEY.md §8d",
            "peak_measured": fp32_peak_measured,
            "frac_of_measured_peak": (achieved / fp32_peak_measured) if fp32_peak_measured else None,
            "peak_measured_source": "mcskin_cuda_fp32_issue_peak: 8 independent unfused FMUL+FADD chains per thread, best of 6 launches",
            "alg_ops_per_launch": step_ops, "ms_per_launch": device_ms,
            "serial_breakdown": {
                "what": "one stream, direct launches, same frame (outside the timed region)",
                "ms_primary_pass": pass_ms["primary"], "ms_shade_pass": pass_ms["shade"],
                "frac_primary_pass": (ops["primary_pass"] / world / (pass_ms["primary"] * 1e-3) / 1e12 / fp32_peak) if pass_ms["primary"] > 0 else None,
                "frac_shade_pass": (ops["shade_pass"] / world / (pass_ms["shade"] * 1e-3) / 1e12 / fp32_peak) if pass_ms["shade"] > 0 else None},
            "hbm_framebuffer": {"bytes_per_frame": fb_bytes, "achieved_gbs": fb_bytes / (ms_per_step * 1e-3) / 1e9,
                                "peak_gbs": peaks["hbm_gbs"], "frac": fb_bytes / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"]},
        },
        "active_pixels": int(n_active),
    }
    if world == 1 and not args.no_cpu_baseline and not args.kernel_only:
        try:
            res = cpu_reference_run(CPU_SAMPLE, 1, 0)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as exc:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "Mrays/s", "cores": None, "kind": "unavailable", "sample": str(exc)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["own", "reference"], default="own")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="tuning runs: skip the e2e and cpu_baseline legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
