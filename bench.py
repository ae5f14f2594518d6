#!/usr/bin/env python
"""bench.py — post-processed images/sec (decode + NMS) on 1/2/4/8 B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg5|cfg1] [--batch B]
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

A step = one pass of the hot path (fused decode + candidate filter + top-k + NMS + gather) over one
batch of synthetic raw head logits resident in HBM.  Default workload = BASELINE.json's headline
configuration: 1280x1280 with the P2 stride-4 head (136 000 anchors), SAR posture JDE head (nc=1,
256-d embedding + 6 state logits, no = 327), val-mode thresholds conf 0.001 / IoU 0.7 / max_nms 30000 /
max_det 300, 16 images per GPU (cfg/default.yaml:15).  One batch of inputs is 2.85 GB (hot channels
566 MB) — larger than the 126 MB L2, so no flush is needed between steps.
Multi-GPU: images are independent -> weak scaling, batch sharded by rank, no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def ensure_built():
    """libsarpost.so normally exists (the driver runs __graft_entry__.build()); build it if a fresh checkout lacks it.
    Under torchrun only local rank 0 compiles, the others wait for the file."""
    lib = os.path.join(ROOT, "sar-yolo_b200", "libsarpost.so")
    if os.path.exists(lib):
        return
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        import __graft_entry__ as g

        g.build()
    else:
        t0 = time.time()
        while not os.path.exists(lib) and time.time() - t0 < 600:
            time.sleep(1.0)
        time.sleep(2.0)

WORKLOADS = {
    # name: (imgsz, strides, nc, embed_dim, state_classes, per-GPU batch, nms kwargs, cls_mean, description)
    "cfg3": (1280, (4, 8, 16, 32), 1, 256, 6, 16,
             dict(conf_thres=0.001, iou_thres=0.7, max_det=300, max_nms=30000, multi_label=False), -4.0,
             "1280x1280 P2 head (136000 anchors), JDE nc=1 no=327, val-mode conf 0.001 / iou 0.7 / max_nms 30000 / max_det 300"),
    "cfg2": (640, (8, 16, 32), 1, 256, 6, 64,
             dict(conf_thres=0.25, iou_thres=0.7, max_det=300, max_nms=30000, multi_label=False), -4.0,
             "batch 64 at 640x640 (8400 anchors), JDE nc=1 no=327, predict-mode conf 0.25 / iou 0.7"),
    "cfg1": (640, (8, 16, 32), 1, 256, 6, 1,
             dict(conf_thres=0.25, iou_thres=0.7, max_det=300, max_nms=30000, multi_label=False), -4.0,
             "batch 1 at 640x640 (8400 anchors), JDE nc=1 no=327, conf 0.25 / iou 0.7"),
    "cfg4": (640, (8, 16, 32), 1, 256, 6, 512,
             dict(conf_thres=0.25, iou_thres=0.7, max_det=300, max_nms=30000, multi_label=False), -4.0,
             "SAHI-style 640 tiles of 4000x3000 frames (48 tiles/frame), 512 tiles in total sharded over the GPUs "
             "(strong scaling), per-tile post-process + all-gather of counts/boxes + cross-tile merge per frame"),
    "cfg5": (640, (8, 16, 32), 6, 0, 0, 32,
             dict(conf_thres=0.001, iou_thres=0.7, max_det=300, max_nms=30000, multi_label=True), -4.0,
             "val-mode multi_label sweep at 640x640, Detect nc=6 no=70, conf 0.001, 32 images per GPU"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="sarpost", choices=["sarpost", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the workload's)")
    ap.add_argument("--streams", type=int, default=0,
                    help="CUDA streams the timed steps are issued on round-robin: 0 (default) = 2 for batches of >= 8 images, "
                         "else 1 (tiny batches are launch-bound and gain nothing); 2 = two independent batches "
                         "in flight, each with its own input buffers (the NMS kernel of one batch overlaps the fused "
                         "decode of the next); 1 = strictly one batch in flight.  The single-stream figure is always "
                         "measured too and reported as `single_stream`.")
    ap.add_argument("--quick", action="store_true", help="profiling run: no clock probe, no e2e, no CPU baseline")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-images", type=int, default=0, help="images in the CPU baseline sample (default: ~10-30 s of work)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_threads():
    # the reference's NUM_THREADS (ultralytics/utils/__init__.py:43), applied by select_device("cpu")
    return min(8, max(1, (os.cpu_count() or 1) - 1))


def cpu_reference_step(levels_cpu, strides, nc, ed, sc, kw):
    """decode + non_max_suppression exactly as the reference runs them on CPU (oracle port; the
    suppression call is torchvision.ops.nms like ops.py:296 when torchvision is importable)."""
    from oracle import postprocess_ref as R

    try:
        import torchvision  # noqa: F401
        nms_fn = R.nms_torchvision
    except Exception:
        nms_fn = R.nms_ref
    y = R.decode_ref(levels_cpu, strides, nc, 16, ed, sc)
    return R.non_max_suppression_ref(y, nc=nc, nms_fn=nms_fn, stable_topk=False, **kw)


def time_cpu_baseline(levels_cpu, strides, nc, ed, sc, kw, repeats=1):
    import torch

    torch.set_num_threads(cpu_threads())
    n_img = levels_cpu[0].shape[0]
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        cpu_reference_step(levels_cpu, strides, nc, ed, sc, kw)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    return n_img / best, best


def load_synth_standalone():
    """The synthetic-input generator (pure torch) loaded by file path, so the CPU reference arm does not import the
    `sarpost` package and therefore never maps libsarpost.so into its process."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_sarpost_synth", os.path.join(ROOT, "sar-yolo_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run_reference_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    imgsz, strides, nc, ed, sc, bs, kw, cls_mean, desc = WORKLOADS[args.workload]
    synth = load_synth_standalone()  # the product package (and libsarpost.so) is never imported by this arm

    shapes = synth.level_shapes(imgsz, strides)
    per_step = args.cpu_images or (1 if args.workload == "cfg3" else min(bs, 8))
    levels = synth.head_outputs(per_step, shapes, nc, ed, sc, cls_mean=cls_mean, seed=3000)
    torch.set_num_threads(cpu_threads())
    for _ in range(max(args.warmup, 1)):
        cpu_reference_step(levels, strides, nc, ed, sc, kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(levels, strides, nc, ed, sc, kw)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} image(s) of {args.workload} per step, oracle port of head.py:214-249 + ops.py:167-316 with torchvision.ops.nms CPU"
    line = {
        "impl": "reference", "metric": "post-processed images/sec (decode+NMS)", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "images_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cpu_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake_slowdown": 0x80}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._thr = [], set(), None, threading.Event(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


L2_BYTES = 126e6  # B200 L2

# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line: keep a private handle on the real stdout for it and point fd 1 at stderr,
    so that anything a library prints there (NCCL prints its version on stdout) cannot end up in front of the line."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _JSON_OUT


def emit(line: dict) -> None:
    out = claim_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":  # CPU arm: oracle port + torchvision only — libsarpost.so is neither built nor loaded here
        return run_reference_arm(args)
    ensure_built()
    if args.quick:
        args.no_e2e = args.no_cpu_baseline = True

    import torch
    import torch.distributed as dist

    import sarpost
    from sarpost import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:  # NUMA locality for the pinned host buffers of the e2e leg: run this rank on the CPUs next to its GPU
        import pynvml

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    imgsz, strides, nc, ed, sc, bs, kw, cls_mean, desc = WORKLOADS[args.workload]
    bs = args.batch or bs
    if args.streams <= 0:
        args.streams = 2 if bs >= 8 else 1
    sahi = args.workload == "cfg4"
    scaling = "weak"
    if sahi:  # strong scaling: a fixed total of tiles is sharded over the ranks
        scaling = "strong"
        total_tiles = bs
        if total_tiles % n_gpus:
            raise SystemExit("cfg4 needs the tile count to divide by the number of GPUs")
        bs = total_tiles // n_gpus
        args.no_e2e = True
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    shapes = synth.level_shapes(imgsz, strides)
    anchors = sum(h * w for h, w in shapes)
    levels = synth.head_outputs(bs, shapes, nc, ed, sc, cls_mean=cls_mean, seed=1000 * 3 + rank, device=dev)
    torch.cuda.synchronize()

    if sahi:
        origins1 = sarpost.dist.sahi_grid(4000, 3000, 640, 0.2).to(dev)  # (48, 2)
        tpf = origins1.shape[0]
        n_frames = -(-total_tiles // tpf)
        f_lo, f_hi = sarpost.dist.shard_range(n_frames, rank, world)
        # gathered detections of ALL tiles (6 columns), padded with empty tiles up to whole frames
        g_rows = torch.zeros((n_frames * tpf, kw["max_det"], 6), dtype=torch.float32, device=dev)
        g_cnt = torch.zeros((n_frames * tpf,), dtype=torch.int32, device=dev)
        origins = origins1.repeat(n_frames, 1)
        lo = rank * bs
        peer = None
        if world > 1 and not os.environ.get("SARPOST_BENCH_NCCL_GATHER"):
            # fused gather + exchange: K5 stores this rank's rows/counts into every rank's buffer over NVLink
            peer = sarpost.dist.PeerGatherBuffer(bs, kw["max_det"], 6, dev)

        def step():
            # boxes only: the extras of the few rows that survive the merge are fetched afterwards from the rank
            # that owns the tile (sarpost.gather_extras), not for 300 rows of every tile
            if peer is not None:
                rows_all, cnt_all = sarpost.postprocess_fused(levels, spec, with_extras=False, peer_out=peer.next(), **kw)
                peer.barrier()
                if f_hi > f_lo:
                    t0, t1 = f_lo * tpf, min(f_hi * tpf, total_tiles)
                    g_rows[t0:t1].copy_(rows_all[t0:t1])  # pad the last frame with empty tiles (g_cnt stays 0 there)
                    g_cnt[t0:t1].copy_(cnt_all[t0:t1])
                    return sarpost.merge_tiles(g_rows[f_lo * tpf:f_hi * tpf], g_cnt[f_lo * tpf:f_hi * tpf],
                                               origins[f_lo * tpf:f_hi * tpf], tpf, iou_thres=kw["iou_thres"],
                                               max_det=kw["max_det"], return_padded=True)
                return rows_all, cnt_all
            out, counts = sarpost.postprocess_fused(levels, spec, return_padded=True, with_extras=False, **kw)
            g_rows[lo:lo + bs].copy_(out)
            g_cnt[lo:lo + bs].copy_(counts)
            if world > 1:
                dist.all_gather_into_tensor(g_rows[:total_tiles], g_rows[lo:lo + bs])
                dist.all_gather_into_tensor(g_cnt[:total_tiles], g_cnt[lo:lo + bs])
            if f_hi > f_lo:
                return sarpost.merge_tiles(g_rows[f_lo * tpf:f_hi * tpf], g_cnt[f_lo * tpf:f_hi * tpf],
                                           origins[f_lo * tpf:f_hi * tpf], tpf, iou_thres=kw["iou_thres"],
                                           max_det=kw["max_det"], return_padded=True)
            return out, counts
    else:
        def step():
            return sarpost.postprocess_fused(next_levels(), spec, return_padded=True, **kw)

    # L2 rule: a batch whose hot channels do not clearly exceed the 126 MB L2 is rotated over enough identical copies
    # of the inputs that a buffer has been evicted by the time it is read again (no flush kernel in the timed region)
    hot_bytes = bs * anchors * (64 + nc) * 4
    n_sets = 1 if (sahi or hot_bytes >= 2 * L2_BYTES) else min(int(math.ceil(2 * L2_BYTES / hot_bytes)), 128)
    if not sahi:
        n_sets = max(n_sets, args.streams)
    level_sets = [levels] + [[x.clone() for x in levels] for _ in range(n_sets - 1)]
    set_iter = [0]

    def next_levels():
        set_iter[0] += 1
        return level_sets[set_iter[0] % n_sets]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out, counts = step()
    barrier()
    launches_per_step = sarpost.ops.last_launch_count() + (3 if sahi else 0)  # cfg4: fused (3) + merge (3)

    # ---- timed region: K steps, CUDA events on the launching (current) stream ----
    clocks = ClockSampler(local)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.start()
    barrier()
    # (a) strictly one batch in flight: K steps back to back on the current stream
    ev0.record()
    for _ in range(args.steps):
        out, counts = step()
    ev1.record()
    barrier()
    ms_single = ev0.elapsed_time(ev1)
    ms = ms_single
    if args.streams > 1 and not sahi:
        # (b) the same K steps issued round-robin on several streams: independent batches, each stream with its own
        # copy of the inputs and its own workspace; the NMS kernel of one batch (few SMs, latency-bound) overlaps
        # the fused decode of the next (HBM-bound).  Every step still does the whole path for one batch.
        streams = [torch.cuda.Stream() for _ in range(args.streams)]
        for s_ in streams:  # warm up each stream (workspace per stream)
            with torch.cuda.stream(s_):
                for _ in range(3):
                    sarpost.postprocess_fused(next_levels(), spec, return_padded=True, **kw)
        barrier()
        ev0.record()
        for s_ in streams:
            s_.wait_event(ev0)
        for i in range(args.steps):
            with torch.cuda.stream(streams[i % args.streams]):
                out, counts = sarpost.postprocess_fused(next_levels(), spec, return_padded=True, **kw)  # n_sets >= streams
        for s_ in streams:
            torch.cuda.current_stream().wait_stream(s_)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
    # second pass over the same K steps with CUDA events around every kernel (recorded by the library on the
    # launching stream, no host sync per step, mean read afterwards).  Kept out of the pass `value` comes from:
    # timing events between kernels cost ~8 % throughput by removing the overlap of consecutive launches.
    stage = None
    if not sahi:
        sarpost.ops.stage_timing(True, accumulate=True)
        for _ in range(3 if args.quick else args.steps):
            step()
        stage = list(sarpost.ops.stage_times())
        sarpost.ops.stage_timing(False)
    if len(clocks.samples) < 20 and not args.quick:  # short region: keep sampling over the same step to have clocks under load
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            for _ in range(20):
                step()
            torch.cuda.synchronize()
    clocks.stop()
    t_ms = torch.tensor([ms, ms_single], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms, ms_single = float(t_ms[0].item()), float(t_ms[1].item())
    value = bs * n_gpus * args.steps / (ms / 1e3)
    single = {"value": bs * n_gpus * args.steps / (ms_single / 1e3), "unit": "images/s", "ms_per_step": ms_single / args.steps,
              "note": "strictly one batch in flight (all K steps on one stream)"}
    n_det = int(counts.sum().item())
    if sahi:
        def step():  # stage timing / roofline below look at the per-tile fused call only
            return sarpost.postprocess_fused(levels, spec, return_padded=True, with_extras=False, **kw)

    # ---- per-stage durations (from the evented pass above); K1 roofline ----
    if stage is None:  # cfg4: time the per-tile fused call on its own
        sarpost.ops.stage_timing(True, accumulate=True)
        for _ in range(2 if args.quick else min(max(args.steps, 5), 50)):
            step()
        stage = list(sarpost.ops.stage_times())
        sarpost.ops.stage_timing(False)
    n_cand = 0
    for x in levels:  # candidates = anchors whose best class probability passes conf (bookkeeping, untimed)
        p = x[:, 64:64 + nc].sigmoid()
        n_cand += int(((p > kw["conf_thres"]).sum() if kw.get("multi_label") and nc > 1 else (p.amax(1) > kw["conf_thres"]).sum()).item())
    k1_bytes = bs * anchors * (64 + nc) * 4 + n_cand * 24
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = k1_bytes / (stage[0] * 1e-3) / 1e9 if stage[0] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k1_fused_tma (decode+score+compact)", "achieved": achieved, "peak": peak,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback (B200_PROFILING.md 6.65 TB/s)",
                "unit": "GB/s", "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0,
                "bytes_per_launch": k1_bytes, "ms_per_launch": stage[0], "traffic": None,
                "stage_ms": {"k1_candidates": stage[0], "k2_k4_select_sort_nms": stage[1], "k5_gather": stage[2], "whole_call": stage[3]}}
    prof = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(args.workload)
        except Exception:
            pass

    # ---- e2e: HOST buffers through the C-ABI host entry (H2D + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        host_levels = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x) for x in levels]
        ctx = sarpost.HostContext(local)
        out_host = torch.empty((bs, kw["max_det"], 6 + spec.nm), dtype=torch.float32, pin_memory=True)
        e2e_steps = max(3, min(args.steps, 20))
        for _ in range(2):
            ctx.postprocess(host_levels, spec, out=out_host, **kw)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.postprocess(host_levels, spec, out=out_host, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_e = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        h2d, d2h = ctx.last_traffic()
        e2e = {"value": bs * n_gpus * e2e_steps / float(t_e.item()), "unit": "images/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "api": "sarpost_fused_host (pinned host level tensors in, host rows out)"}
        ctx.close()
        del host_levels

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        n_img = args.cpu_images or (4 if args.workload == "cfg3" else min(bs, 16))
        n_img = min(n_img, bs)
        levels_cpu = [x[:n_img].cpu() for x in levels]
        cpu_reference_step([x[:1] for x in levels_cpu], strides, nc, ed, sc, kw)  # warm-up (lazy torchvision import)
        v, secs = time_cpu_baseline(levels_cpu, strides, nc, ed, sc, kw)
        cpu = {"value": v, "unit": "images/s", "cores": cpu_threads(), "kind": "port",
               "sample": f"first {n_img} images of the GPU batch, {secs:.1f} s; oracle port of head.py:214-249 + ops.py:167-316 "
                         f"(torch CPU ops + torchvision.ops.nms CPU, {cpu_threads()} torch threads, host has {os.cpu_count()} cpus)"}

    if rank == 0:
        line = {
            "metric": "post-processed images/sec (decode+NMS)", "value": value, "unit": "images/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "images_per_gpu": bs, "global_batch": bs * n_gpus,
                       "anchors": anchors, "channels": spec.no,
                       "streams": (1 if sahi else args.streams),
                       "in_flight": ("one batch" if (sahi or args.streams == 1) else
                                     f"{args.streams} independent batches, steps issued round-robin on {args.streams} CUDA streams, "
                                     "each with its own input buffers; `single_stream` holds the one-batch-in-flight figure"), "parallelism": ((f"tiles sharded x{n_gpus}, gather kernel stores counts+boxes into every rank over NVLink peer memory "
                                        f"(fused gather+exchange, no NCCL collective), frames merged by their owner rank"
                                        if (n_gpus > 1 and not os.environ.get("SARPOST_BENCH_NCCL_GATHER")) else
                                        f"tiles sharded x{n_gpus}, NCCL all-gather of counts+boxes, frames merged by their owner rank") if sahi
                                       else f"batch-sharded x{n_gpus}, no data-path collective"),
                       "l2": (("one batch of inputs (hot channels %.0f MB) exceeds the 126 MB L2; no flush" % (hot_bytes / 1e6))
                              if hot_bytes >= 2 * L2_BYTES else
                              ("hot channels %.1f MB per batch: steps rotate over %d identical input copies (%.0f MB in rotation > 2x the "
                               "126 MB L2), so every step reads its inputs from HBM; no flush" % (hot_bytes / 1e6, n_sets, n_sets * hot_bytes / 1e6))),
                       "input_sets": n_sets,
                       "candidates_per_image": n_cand / bs, "detections_per_image": n_det / bs},
            "single_stream": single, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "launches_per_step": launches_per_step, "clocks": clocks.summary(),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
