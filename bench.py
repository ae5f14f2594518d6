#!/usr/bin/env python
"""bench.py — post-processed images/sec (decode + NMS) on 1/2/4/8 B200, with roofline, the reference's own GPU and
CPU paths timed beside it, a clustered (suppression-heavy) leg, the host-buffer e2e figure and — at N > 1 — the sliced
inference (SAHI) exchange step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg5|cfg1|cfg4] [--batch B]
    python bench.py --impl reference ...        # the reference's CPU path on the host cores (no GPU, no libsarpost.so)

A step = one pass of the hot path (fused decode + candidate filter + top-k + NMS + gather) over one batch of synthetic
raw head logits resident in HBM.  Default workload = BASELINE.json's headline configuration: 1280x1280 with the P2
stride-4 head (136 000 anchors), SAR posture JDE head (nc=1, 256-d embedding + 6 state logits, no = 327), val-mode
thresholds conf 0.001 / IoU 0.7 / max_nms 30000 / max_det 300, 16 images per GPU (cfg/default.yaml:15).  One batch of
inputs is 2.85 GB (hot channels 566 MB) — larger than the 126 MB L2, so no flush is needed between steps.
Multi-GPU: images are independent -> weak scaling, batch sharded by rank, no data-path collective in the image step;
the one real exchange of the path (sliced inference: tiles of a frame on different ranks) is measured in the `sahi` block.

Keys of the JSON line beyond the contract:
  single_stream   the same K steps strictly one batch in flight
  roofline        K1 (fused decode) algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json
  two_streams     the same K steps round-robin on two independent streams (plain sarpost_fused calls)
  cat_layout      the same passes on the concatenated (B, no, H, W) level tensors (the headline input is the split layout:
                  separate branch tensors, channels_last embedding — what `patch(fused=True, split=True)` feeds the kernels)
  clustered       the same shapes/thresholds on inputs whose class logits carry 50 Gaussian blobs per image: neighbouring
                  anchors fire together and overlap, so NMS has to suppress (the regime of a trained detector)
  reference_gpu   the reference's own path as a `device=0` user runs it on this GPU: JDE._inference (torch CUDA ops) +
                  ops.non_max_suppression (torch CUDA ops + torchvision.ops.nms CUDA) — the UNMODIFIED reference code from
                  baseline/_ref when that install is present (kind "reference"), else the oracle's restatement of the same
                  op sequence on CUDA tensors (kind "port")
  cpu_baseline    the reference's CPU path on the host cores, bounded sample (N = 1 only)
  e2e             HOST buffers through the C-ABI host entry, H2D + D2H inside the timed region (+ the measured H2D ceiling)
  sahi            (N > 1) cfg4: 512 tiles of 640x640 sharded over the ranks, per-tile fused call -> exchange -> cross-tile merge
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
             "(strong scaling), per-tile post-process + exchange of counts/boxes + cross-tile merge per frame"),
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
    ap.add_argument("--quick", action="store_true", help="profiling run: no clock probe, no extra legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--no-clustered", action="store_true")
    ap.add_argument("--no-sahi", action="store_true")
    ap.add_argument("--no-split", action="store_true")
    ap.add_argument("--blobs", type=int, default=0, help="profiling: make the MAIN workload the clustered variant (Gaussian blobs per image)")
    ap.add_argument("--cpu-images", type=int, default=0, help="images in the CPU baseline sample (default: ~10-30 s of work)")
    return ap.parse_args()


def load_synth_standalone():
    """The synthetic-input generator (pure torch) loaded by file path, so the CPU reference arm does not import the
    `sarpost` package and therefore never maps libsarpost.so into its process."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("_sarpost_synth", os.path.join(ROOT, "sar-yolo_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------------
# The reference's own implementation of the path (CPU arm, cpu_baseline, reference_gpu)
# ------------------------------------------------------------------------------------------------
def cpu_threads():
    # the reference's NUM_THREADS (ultralytics/utils/__init__.py:43), applied by select_device("cpu")
    return min(8, max(1, (os.cpu_count() or 1) - 1))


class ReferencePath:
    """decode + non_max_suppression the way the reference runs them, on `device`.

    kind "reference": the UNMODIFIED reference imported from baseline/_ref (or /root/reference in the build container):
        head.JDE._inference / head.Detect._inference (nn/modules/head.py:100-131, :214-249) followed by
        utils/ops.non_max_suppression (:167-316), which calls torchvision.ops.nms (:296) — CPU or CUDA build, whichever
        device the tensors are on.  `max_time_img` is raised so the soft wall-clock limit (:312-314) never truncates a batch.
    kind "port": no reference install on this machine — the oracle's restatement of the same op sequence
        (oracle/postprocess_ref.py, pinned bit-equal to the live reference by tests/test_oracle.py) with torchvision.ops.nms.
    """

    def __init__(self, strides, nc, ed, sc, device):
        import torch

        self.strides, self.nc, self.ed, self.sc, self.device = strides, nc, ed, sc, torch.device(device)
        self.kind, self.source, self.module = "port", "oracle/postprocess_ref.py", None
        if not os.environ.get("SARPOST_BENCH_FORCE_PORT"):
            try:
                from oracle import ref_shim

                if ref_shim.available():
                    self.ops, _, head = ref_shim.load()
                    ch = tuple(64 for _ in strides)
                    with torch.no_grad():
                        m = (head.JDE(nc=nc, embed_dim=ed, state_classes=(sc or None), ch=ch) if ed else head.Detect(nc=nc, ch=ch))
                        m.stride = torch.tensor([float(s) for s in strides])
                        m.eval()
                        self.module = m.to(self.device)
                    self.kind, self.source = "reference", ref_shim.source()
            except Exception as e:  # noqa: BLE001  (a broken install must not take the bench down: fall back to the port)
                print(f"[bench] reference import failed ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
                self.module = None
                self.kind, self.source = "port", "oracle/postprocess_ref.py"

    def describe(self):
        import torch
        import torchvision

        what = ("unmodified reference code (" + self.source + "): head._inference + ops.non_max_suppression" if self.kind == "reference"
                else "oracle port of head.py:214-249 + ops.py:167-316 (pinned bit-equal to the live reference)")
        return f"{what}, torch {torch.__version__} ops on {self.device.type} + torchvision {torchvision.__version__} ops.nms ({self.device.type})"

    def step(self, levels, kw):
        import torch

        with torch.no_grad():
            if self.kind == "reference":
                y = self.module._inference(list(levels))
                return self.ops.non_max_suppression(y, nc=self.nc, max_time_img=1e9, **kw)
            import torchvision

            from oracle import postprocess_ref as R

            y = R.decode_ref(levels, self.strides, self.nc, 16, self.ed, self.sc, device=self.device)
            return R.non_max_suppression_ref(y, nc=self.nc, nms_fn=torchvision.ops.nms, stable_topk=False, device=self.device, **kw)


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
    ref = ReferencePath(strides, nc, ed, sc, "cpu")
    for _ in range(max(args.warmup, 1)):
        ref.step(levels, kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step(levels, kw)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} image(s) of {args.workload} per step; {ref.describe()}; {cpu_threads()} torch threads, host has {os.cpu_count()} cpus"
    line = {
        "impl": "reference", "impl_detail": "reference-cpu" if ref.kind == "reference" else "reference-cpu-port",
        "metric": "post-processed images/sec (decode+NMS)", "value": value, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "images_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cpu_threads(), "kind": ref.kind, "sample": sample},
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


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    """What every leg of the GPU arm needs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import sarpost

        self.args, self.torch, self.dist, self.sarpost = args, torch, dist, sarpost
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        try:  # NUMA locality for the pinned host buffers of the e2e leg: run this rank on the CPUs next to its GPU
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(self.local))
        except Exception:
            pass
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]


PIPE_DEPTH = int(os.environ.get("SARPOST_BENCH_PIPE_DEPTH", "3"))  # batches whose tails may be pending / running behind the decode stream


def guarded(name, fn):
    """An optional leg must not take the headline down with it: its failure is reported in its place in the line."""
    try:
        return fn()
    except Exception as e:  # noqa: BLE001
        print(f"bench: leg `{name}` failed: {type(e).__name__}: {e}", file=sys.stderr)
        return {"unavailable": f"{type(e).__name__}: {str(e)[:300]}"}


def rotating(level_sets):
    it = [0]

    def nxt():
        it[0] += 1
        return level_sets[it[0] % len(level_sets)]

    return nxt


def time_steps(cx, step, n, streams=None):
    """K steps timed with CUDA events on the launching stream (round-robin over `streams` when given), barrier +
    synchronize on both sides.  Returns this rank's milliseconds."""
    torch = cx.torch
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.barrier()
    ev0.record()
    if not streams:
        for _ in range(n):
            step()
    else:
        for s_ in streams:
            s_.wait_event(ev0)
        for i in range(n):
            with torch.cuda.stream(streams[i % len(streams)]):
                step()
        for s_ in streams:
            torch.cuda.current_stream().wait_stream(s_)
    ev1.record()
    cx.barrier()
    return ev0.elapsed_time(ev1)


def stage_means(cx, step, n):
    """Per-kernel CUDA-event durations recorded by the library on the launching stream (mean over n steps)."""
    ops = cx.sarpost.ops
    cx.torch.cuda.synchronize()
    # a spin kernel first, long enough for the host to get ahead of the device: the events then bracket back-to-back kernels
    # and an event-to-event span is the kernel's duration, not the host's launch cadence (small workloads: the host needs
    # longer per call than the device)
    cx.torch.cuda._sleep(int(min(n, 400) * 200_000))
    ops.stage_timing(True, accumulate=True)
    for _ in range(n):
        step()
    st = list(ops.stage_times())
    ops.stage_timing(False)
    return st


def count_candidates(torch, levels, nc, kw):
    n = 0
    for x in levels:  # candidates = anchors whose best class probability passes conf (bookkeeping, untimed)
        p = x[:, 64:64 + nc].sigmoid()
        n += int(((p > kw["conf_thres"]).sum() if kw.get("multi_label") and nc > 1 else (p.amax(1) > kw["conf_thres"]).sum()).item())
    return n


def leg_clustered(cx, spec, shapes, nc, ed, sc, bs, kw, cls_mean, steps, use_split=True):
    """Shapes / thresholds of the main workload on clustered inputs (synth.head_outputs(blobs=50)): the NMS kernel
    has to walk deep into the sorted candidates because most of them are suppressed."""
    torch, sarpost = cx.torch, cx.sarpost
    sets = []
    for i in range(2):
        lv = sarpost.synth.head_outputs(bs, shapes, nc, ed, sc, cls_mean=cls_mean, seed=3500 + cx.rank + 100 * i, device=cx.dev, blobs=50)
        sets.append(sarpost.split_levels(lv, spec, emb_channels_last=True) if use_split else lv)
        del lv
    levels = sets[0]
    stats = torch.zeros((bs, 4), dtype=torch.int64, device=cx.dev)

    def step():
        return sarpost.postprocess_fused(levels, spec, return_padded=True, **kw)

    for _ in range(3):
        out, counts = step()
    ms = time_steps(cx, step, steps)
    st = stage_means(cx, step, min(steps, 50))
    sarpost.postprocess_fused(levels, spec, return_padded=True, nms_stats=stats, **kw)
    torch.cuda.synchronize()
    # the same steps through the software pipeline (two input sets alternating)
    pl = sarpost.Pipeline(cx.dev, depth=PIPE_DEPTH)
    ring = [(torch.empty((bs, kw["max_det"], 6 + spec.nm), dtype=torch.float32, device=cx.dev), torch.empty((bs,), dtype=torch.int32, device=cx.dev))
            for _ in range(4)]
    for i in range(4):
        pl.submit(sets[i % 2], spec, out=ring[i % 4], **kw)
    pl.wait()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.barrier()
    ev0.record()
    for i in range(steps):
        pl.submit(sets[i % 2], spec, out=ring[i % 4], **kw)
    pl.wait()
    ev1.record()
    cx.barrier()
    ms_pipe = ev0.elapsed_time(ev1)
    pl.close()
    ms, ms_pipe = cx.max_over_ranks(ms, ms_pipe)
    s = stats.double().mean(0).tolist()
    return {"value": bs * cx.world * steps / (ms_pipe / 1e3), "unit": "images/s", "ms_per_step": ms_pipe / steps, "steps": steps,
            "single_stream": {"value": bs * cx.world * steps / (ms / 1e3), "ms_per_step": ms / steps},
            "k1_ms": st[0], "k4_ms": st[1], "k5_ms": st[2],
            "keeps_depth": s[0], "pair_tests": s[1], "sub_chunks": s[2], "collections": s[3],
            "detections_per_image": float(counts.sum().item()) / bs,
            "input": "synth.head_outputs(blobs=50): 50 Gaussian bumps (+6 logit) per image and level on the class logits, box logits of "
                     "fired anchors pulled to a common shape so neighbours overlap; `value` = software pipeline like the headline, "
                     "`single_stream` / k*_ms = one batch in flight",
            "note": "per image means on rank 0: keeps_depth = sorted candidates NMS consumed before it had max_det keeps (or ran out), "
                    "pair_tests = IoU tests executed, sub_chunks / collections = passes of the NMS / selection loops"}


def leg_reference_gpu(cx, levels, strides, nc, ed, sc, kw, n_img, reps, our_rows):
    """The reference's own path on this GPU (see ReferencePath) on the first n_img images of the batch."""
    torch = cx.torch
    ref = ReferencePath(strides, nc, ed, sc, cx.dev)
    sub = [x[:n_img].contiguous() for x in levels]
    for _ in range(2):
        rows = ref.step(sub, kw)
    torch.cuda.synchronize()
    clocks = ClockSampler(cx.local)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(reps):
        rows = ref.step(sub, kw)
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks.stop()
    ms = ev0.elapsed_time(ev1)
    # kept-set agreement with our rows for the same images (ours are bit-equal to the CPU oracle on identical decoded
    # input; torchvision's CUDA kernel may legitimately differ on 1-ulp IoU pairs and on score ties — SURVEY §7 hard part 1)
    only_ours = only_ref = total = 0
    for b in range(n_img):
        a, r = our_rows[b][:, :6].float().cpu(), rows[b][:, :6].float().cpu()
        total += max(a.shape[0], r.shape[0])
        if a.shape[0] and r.shape[0]:
            d = (a[:, None, :] - r[None, :, :]).abs()
            tol = 1e-4 * r[None, :, :].abs() + 2e-3
            hit = (d <= tol).all(-1)
            only_ours += int((~hit.any(1)).sum())
            only_ref += int((~hit.any(0)).sum())
        else:
            only_ours += a.shape[0]
            only_ref += r.shape[0]
    return {"value": n_img * reps / (ms / 1e3), "unit": "images/s", "ms_per_image": ms / (n_img * reps), "wall_ms_per_image": 1e3 * wall / (n_img * reps),
            "kind": ref.kind, "images_per_step": n_img, "steps": reps, "what": ref.describe(),
            "kept_set_diff_vs_sarpost": {"only_in_sarpost": only_ours, "only_in_reference_gpu": only_ref, "of": total},
            "clocks": clocks.summary()}


def leg_e2e(cx, levels, spec, bs, kw):
    """HOST buffers through the C-ABI host entry (H2D + D2H inside the timed region), and the H2D ceiling of this rank
    measured the same way (all ranks copying concurrently): the bare cudaMemcpyAsync of the same bytes, nothing else."""
    torch, sarpost = cx.torch, cx.sarpost
    host_levels = [torch.empty(x.shape, dtype=x.dtype, pin_memory=True).copy_(x) for x in levels]
    ctx = sarpost.HostContext(cx.local)
    out_host = torch.empty((bs, kw["max_det"], 6 + spec.nm), dtype=torch.float32, pin_memory=True)
    e2e_steps = max(3, min(cx.args.steps, 20))
    for _ in range(2):
        ctx.postprocess(host_levels, spec, out=out_host, **kw)
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.postprocess(host_levels, spec, out=out_host, **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    (dt,) = cx.max_over_ranks(dt)
    h2d, d2h = ctx.last_traffic()
    ctx.close()
    # H2D ceiling: the same hot channels (box + cls rows of every image and level), one contiguous copy per image and
    # level exactly like the host entry issues them, all ranks at once
    nch = 64 + spec.nc
    dst = [torch.empty((bs, nch) + tuple(x.shape[2:]), dtype=x.dtype, device=cx.dev) for x in levels]
    copy_stream = torch.cuda.Stream()

    def copy_all():
        with torch.cuda.stream(copy_stream):
            for d, h in zip(dst, host_levels):
                for b in range(bs):
                    d[b].copy_(h[b, :nch], non_blocking=True)

    copy_all()
    copy_stream.synchronize()
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        copy_all()
    copy_stream.synchronize()
    dt_c = time.perf_counter() - t0
    (dt_c,) = cx.max_over_ranks(dt_c)
    hot = sum(d.numel() * d.element_size() for d in dst)
    value = bs * cx.world * e2e_steps / dt
    ceiling = bs * cx.world * e2e_steps / dt_c
    del host_levels, dst
    return {"value": value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
            "api": "sarpost_fused_host (pinned host level tensors in, host rows out)",
            "h2d_ceiling": {"value": ceiling, "unit": "images/s", "gbytes_per_s_per_gpu": hot * e2e_steps / dt_c / 1e9,
                            "gbytes_per_s_aggregate": hot * cx.world * e2e_steps / dt_c / 1e9,
                            "what": f"bare pinned->device cudaMemcpyAsync of the same {hot / 1e6:.0f} MB of box+cls channels per step, "
                                    f"{cx.world} rank(s) concurrently, max over ranks"},
            "frac_of_h2d_ceiling": value / ceiling}


def leg_sahi(cx, steps):
    """cfg4 (BASELINE configs[3]): 512 SAHI tiles (640x640, 48 per 4000x3000 frame) sharded over the ranks — strong
    scaling.  Step = per-tile fused post-process -> exchange -> cross-tile merge per frame.  Two shardings:
      tiles   equal tile ranges per rank; K5 stores rows+counts into every rank's buffer over NVLink peer memory, one
              signal-pad barrier, every rank merges the frames it owns (frames straddle ranks);
      frames  whole frames per rank (dist.shard_frames): the merge is rank-local, only the merged per-frame rows are
              exchanged (again by peer stores, from the gather kernel of the merge).
    Both are timed; `value` is the better one and says which."""
    torch, sarpost, D = cx.torch, cx.sarpost, cx.sarpost.dist
    imgsz, strides, nc, ed, sc, total_tiles, kw, cls_mean, desc = WORKLOADS["cfg4"]
    world, rank, dev = cx.world, cx.rank, cx.dev
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    shapes = sarpost.synth.level_shapes(imgsz, strides)
    origins1 = D.sahi_grid(4000, 3000, 640, 0.2).to(dev)
    tpf = origins1.shape[0]
    n_frames = -(-total_tiles // tpf)
    max_det = kw["max_det"]
    res = {}
    n_in_flight = int(os.environ.get("SARPOST_BENCH_SAHI_INFLIGHT", "0")) or 4

    class Marks:
        def __init__(self, on):
            self.on, self.ev = on, []

        def __call__(self):
            if self.on:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                self.ev.append(e)

    def run(name, step_factory, n_local, phase_names):
        hot = n_local * sum(h * w for h, w in shapes) * (64 + nc) * 4
        n_sets = max(n_in_flight, min(int(math.ceil(2 * L2_BYTES / max(hot, 1))), 16))
        sets = [sarpost.synth.head_outputs(n_local, shapes, nc, ed, sc, cls_mean=cls_mean, seed=4000 + 17 * rank + 1000 * s_, device=dev)
                for s_ in range(n_sets)]
        nxt = rotating(sets)
        pipes = [step_factory(i) for i in range(n_in_flight)]
        streams = [torch.cuda.Stream() for _ in range(n_in_flight)]
        for i, s_ in enumerate(streams):  # warm-up: workspaces are cached per stream
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                for _ in range(3):
                    pipes[i](nxt())
        for _ in range(3):
            pipes[0](nxt())
        cx.barrier()
        ms1 = time_steps(cx, lambda: pipes[0](nxt()), steps)
        # steps alternate between the pipelines (own exchange buffers) on their own streams: the barrier + merge latency
        # chain of one step overlaps the per-tile kernels of the next
        k = [0]

        def one():
            k[0] += 1
            pipes[(k[0] - 1) % n_in_flight](nxt())

        ms2 = time_steps(cx, one, steps, streams) if n_in_flight > 1 else ms1
        # the same steps replayed from CUDA graphs (one graph per pipeline = two consecutive steps, i.e. both exchange
        # buffers of its double-buffered PeerGatherBuffer): every kernel, the signal-pad barrier and the allocations of a
        # step are captured once, so the host spends a graph launch per two steps instead of ~100 us of Python + launch
        # calls per step — which is what bounds the eager loop once a rank's share of the tiles takes less than that
        ms3, graph_err = None, None
        if os.environ.get("SARPOST_BENCH_SAHI_GRAPH", "1") != "0":
            graphs = []
            try:
                for i, s_ in enumerate(streams):
                    g = torch.cuda.CUDAGraph()
                    a, b = sets[(2 * i) % n_sets], sets[(2 * i + 1) % n_sets]
                    torch.cuda.synchronize()
                    with torch.cuda.graph(g, stream=s_):
                        pipes[i](a)
                        pipes[i](b)
                    graphs.append(g)
            except Exception as e:  # noqa: BLE001  (capture support of the symmetric-memory barrier varies with the torch build)
                graph_err = f"{type(e).__name__}: {str(e)[:200]}"
            ok = cx.max_over_ranks(0.0 if graph_err is None else 1.0)[0] == 0.0  # all ranks or none: the steps contain a barrier
            if ok:
                kg = [0]

                def two_steps():
                    kg[0] += 1
                    graphs[(kg[0] - 1) % len(graphs)].replay()

                for s_ in streams:
                    with torch.cuda.stream(s_):
                        two_steps()
                torch.cuda.synchronize()
                cx.barrier()
                ms3 = time_steps(cx, two_steps, max(steps // 2, 1), streams) * steps / (2 * max(steps // 2, 1))
            torch.cuda.synchronize()
            del graphs
        ms1, ms2, ms3v = cx.max_over_ranks(ms1, ms2, ms3 or 0.0)
        eager_ms = ms2
        if ms3 is not None and ms3v < ms2:
            ms2 = ms3v
        # per-phase breakdown (one step in flight, events between the phases, mean over a few steps)
        reps, acc = 20, None
        for _ in range(reps):
            torch.cuda._sleep(1_000_000)  # ~0.5 ms spin first: the step is fully enqueued before it starts, the events bracket device time
            ev = pipes[0](nxt(), phases=True)
            torch.cuda.synchronize()
            d = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(ev) - 1)]
            acc = d if acc is None else [a + b for a, b in zip(acc, d)]
        acc = cx.max_over_ranks(*[a / reps for a in acc])
        res[name] = {"tiles_per_s": total_tiles * steps / (ms2 / 1e3), "ms_per_step": ms2 / steps,
                     "launch": "cuda graphs (two steps per replay)" if ms2 != eager_ms else "eager",
                     "eager": {"tiles_per_s": total_tiles * steps / (eager_ms / 1e3), "ms_per_step": eager_ms / steps},
                     "graphs": ({"tiles_per_s": total_tiles * steps / (ms3v / 1e3), "ms_per_step": ms3v / steps} if ms3 is not None else
                                {"unavailable": graph_err or "capture failed on another rank"}),
                     "one_in_flight": {"tiles_per_s": total_tiles * steps / (ms1 / 1e3), "ms_per_step": ms1 / steps},
                     "tiles_on_rank0": n_local, "input_sets": n_sets, "hot_mb_in_rotation": n_sets * hot / 1e6,
                     "phase_ms_max_over_ranks": dict(zip(phase_names, acc))}
        del sets, pipes
        torch.cuda.empty_cache()

    # ---- sharding by tiles: equal ranges, fused gather + exchange, frames merged by their owner ----
    if total_tiles % world == 0:
        per = total_tiles // world
        lo = rank * per
        f_lo, f_hi = D.shard_range(n_frames, rank, world)
        origins = origins1.repeat(n_frames, 1)

        def factory_tiles(i):
            peer = None
            if world > 1:
                peer = D.PeerGatherBuffer(per, max_det, 6, dev, total_slots=n_frames * tpf, slot_offset=lo)
                for r_, c_, _, _ in peer._bufs:  # slots beyond total_tiles (padding of the last frame) stay empty tiles
                    c_.zero_()
            g_rows = torch.zeros((n_frames * tpf, max_det, 6), dtype=torch.float32, device=dev) if world == 1 else None
            g_cnt = torch.zeros((n_frames * tpf,), dtype=torch.int32, device=dev) if world == 1 else None

            def step(levels, phases=False):
                mark = Marks(phases)
                mark()
                if peer is not None:
                    rows_all, cnt_all = sarpost.postprocess_fused(levels, spec, with_extras=False, peer_out=peer.next(), nms_cluster=1, **kw)
                    mark()
                    peer.barrier()
                else:
                    rows_all, cnt_all = g_rows, g_cnt
                    sarpost.postprocess_fused(levels, spec, with_extras=False, return_padded=True, out=(g_rows[:per], g_cnt[:per]), nms_cluster=1, **kw)
                    mark()
                mark()
                if f_hi > f_lo:
                    sarpost.merge_tiles(rows_all[f_lo * tpf:f_hi * tpf], cnt_all[f_lo * tpf:f_hi * tpf], origins[f_lo * tpf:f_hi * tpf], tpf,
                                        iou_thres=kw["iou_thres"], max_det=max_det, return_padded=True)
                mark()
                return mark.ev

            return step

        run("tiles", factory_tiles, per, ("per_tile_fused", "exchange_barrier", "merge"))

    # ---- sharding by frames: merge is rank-local, merged frames exchanged by the merge's own gather kernel ----
    spans = [D.shard_range(n_frames, r, world) for r in range(world)]
    if world > 1 and all(hi > lo for lo, hi in spans):
        f_lo, f_hi = spans[rank]
        t_lo, t_hi = f_lo * tpf, min(f_hi * tpf, total_tiles)
        nf, n_local = f_hi - f_lo, t_hi - t_lo
        org = origins1.repeat(nf, 1)

        def factory_frames(i):
            peer = D.PeerGatherBuffer(nf, max_det, 6, dev, total_slots=n_frames, slot_offset=f_lo)
            rows = torch.zeros((nf * tpf, max_det, 6), dtype=torch.float32, device=dev)
            cnt = torch.zeros((nf * tpf,), dtype=torch.int32, device=dev)

            def step(levels, phases=False):
                mark = Marks(phases)
                mark()
                sarpost.postprocess_fused(levels, spec, with_extras=False, return_padded=True, out=(rows[:n_local], cnt[:n_local]), nms_cluster=1, **kw)
                mark()
                sarpost.merge_tiles(rows, cnt, org, tpf, iou_thres=kw["iou_thres"], max_det=max_det, peer_out=peer.next())
                mark()
                peer.barrier()
                mark()
                return mark.ev

            return step

        run("frames", factory_frames, n_local, ("per_tile_fused", "merge_with_peer_stores", "barrier"))

    best = max(res, key=lambda k_: res[k_]["tiles_per_s"])
    return {"workload": "cfg4: " + desc, "value": res[best]["tiles_per_s"], "unit": "tiles/s", "ms_per_step": res[best]["ms_per_step"],
            "sharding": best, "scaling": "strong", "total_tiles": total_tiles, "tiles_per_frame": tpf, "frames": n_frames, "steps": steps,
            "exchange": ("gather kernel stores rows+counts into every rank's symmetric-memory buffer over NVLink (P2P stores) + one signal-pad "
                         "barrier; no NCCL collective in the step") if world > 1 else "none (one rank)",
            "in_flight": f"{n_in_flight} step(s) in flight, each on its own stream with its own exchange buffers (`one_in_flight` = strictly sequential steps)",
            "by_sharding": res}


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":  # CPU arm: reference / oracle port + torchvision only — libsarpost.so is neither built nor loaded here
        return run_reference_arm(args)
    ensure_built()
    if args.quick:
        args.no_e2e = args.no_cpu_baseline = args.no_reference_gpu = args.no_clustered = args.no_sahi = True

    cx = Ctx(args)
    torch, dist, sarpost = cx.torch, cx.dist, cx.sarpost
    synth = sarpost.synth
    world, rank, dev = cx.world, cx.rank, cx.dev
    n_gpus = world

    imgsz, strides, nc, ed, sc, bs, kw, cls_mean, desc = WORKLOADS[args.workload]
    bs = args.batch or bs
    if args.workload == "cfg4":  # as a --workload the SAHI job is the whole line (strong scaling)
        blk = leg_sahi(cx, args.steps)
        if rank == 0:
            emit({"metric": "post-processed tiles/sec (decode+NMS+merge)", "value": blk["value"], "unit": "tiles/s", "n_gpus": n_gpus,
                  "steps": args.steps, "warmup": 3, "ms_per_step": blk["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                  "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": blk["workload"]}, "sahi": blk})
        if world > 1:
            dist.destroy_process_group()
        return 0
    if args.streams <= 0:
        args.streams = 2 if bs >= 8 else 1
    spec = sarpost.HeadSpec(nc=nc, strides=strides, embed_dim=ed, state_classes=sc)
    shapes = synth.level_shapes(imgsz, strides)
    anchors = sum(h * w for h, w in shapes)
    levels = synth.head_outputs(bs, shapes, nc, ed, sc, cls_mean=cls_mean, seed=1000 * 3 + rank, device=dev, blobs=args.blobs)
    torch.cuda.synchronize()

    # L2 rule: a batch whose hot channels do not clearly exceed the 126 MB L2 is rotated over enough identical copies
    # of the inputs that a buffer has been evicted by the time it is read again (no flush kernel in the timed region)
    hot_bytes = bs * anchors * (64 + nc) * 4
    n_sets = 1 if hot_bytes >= 2 * L2_BYTES else min(int(math.ceil(2 * L2_BYTES / hot_bytes)), 128)
    n_sets = max(n_sets, args.streams)
    level_sets = [levels] + [[x.clone() for x in levels] for _ in range(n_sets - 1)]
    # The same values in the split layout (SARPOST_LAYOUT_SPLIT): box / class / embedding / state branch outputs as separate
    # tensors, embedding channels_last — what `patch(fused=True, split=True)` hands over instead of the per-level torch.cat
    # of head.py:204-206.  This is the headline layout (the reference legs are fed the concatenated tensors holding the same
    # numbers); the concatenated layout is measured too (`cat_layout`).
    use_split = not args.no_split
    split_sets = [sarpost.split_levels(lv, spec, emb_channels_last=True) for lv in level_sets] if use_split else None

    def measure_layout(sets, steps, with_stage_steps):
        """single stream / two plain streams / software pipeline over the same K steps; per-kernel event pass."""
        nxt = rotating(sets)

        def step():
            return sarpost.postprocess_fused(nxt(), spec, return_padded=True, **kw)

        for _ in range(max(args.warmup, 3)):
            o_, c_ = step()
        cx.barrier()
        res = {"launches_per_step": sarpost.ops.last_launch_count(), "counts": c_}
        # (a) strictly one batch in flight: K steps back to back on the current stream, through a prepared plan
        # (sarpost.FusedPlan / sarpost_plan_*: the serving-loop form of the same call — geometry validated and tensor maps
        # encoded once, outputs rotating over 4 preallocated sets) and through the general entry point
        plan = sarpost.FusedPlan(sets[0], spec, **kw)
        plan_ring = [(torch.empty((bs, kw["max_det"], 6 + spec.nm), dtype=torch.float32, device=dev),
                      torch.empty((bs,), dtype=torch.int32, device=dev)) for _ in range(4)]
        qk = [0]

        def plan_step():
            qk[0] += 1
            return plan(nxt(), out=plan_ring[qk[0] & 3])

        for _ in range(4):
            o_p, c_p = plan_step()
        torch.cuda.synchronize()
        o_g, c_g = sarpost.postprocess_fused(sets[0], spec, return_padded=True, **kw)
        o_p, c_p = plan(sets[0])
        res["plan_same"] = bool(torch.equal(c_p, c_g)) and all(bool(torch.equal(o_p[b, :n], o_g[b, :n])) for b, n in enumerate(c_g.tolist()))
        res["ms_general"] = time_steps(cx, step, steps)
        res["ms_single"] = time_steps(cx, plan_step, steps)
        res["launches_per_step_plan"] = sarpost.ops.last_launch_count()
        res["ms_two"] = res["ms_pipe"] = None
        res["pipe_same"] = None
        if args.streams > 1:
            # (b) the same K steps issued round-robin on independent streams (plain sarpost_fused calls): the tails of two
            # batches run side by side, but each still queues behind the other batch's decode kernel
            streams = [torch.cuda.Stream() for _ in range(args.streams)]
            for s_ in streams:  # warm up each stream (workspace per stream)
                with torch.cuda.stream(s_):
                    for _ in range(3):
                        step()
            res["ms_two"] = time_steps(cx, step, steps, streams)
            # (c) software pipeline (sarpost_pipeline_*): the same K steps submitted back to back; batch i on stream i % 2, its
            # decode kernel chained to the previous batch's decode kernel, so NMS + gather of batch i run under the decode of
            # batch i+1.  Outputs rotate over 4 preallocated sets (a serving loop consumes set i while i+1.. are in flight).
            pl = sarpost.Pipeline(dev, depth=PIPE_DEPTH)
            out_ring = [(torch.empty((bs, kw["max_det"], 6 + spec.nm), dtype=torch.float32, device=dev),
                         torch.empty((bs,), dtype=torch.int32, device=dev)) for _ in range(4)]
            pk = [0]

            def pstep():
                pk[0] += 1
                return pl.submit(nxt(), spec, out=out_ring[pk[0] % len(out_ring)], **kw)

            for _ in range(4):
                pstep()
            pl.wait()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            cx.barrier()
            ev0.record()
            for _ in range(steps):
                pstep()
            pl.wait()  # the launching stream waits (device side) for every submitted batch
            ev1.record()
            cx.barrier()
            res["ms_pipe"] = ev0.elapsed_time(ev1)
            res["launches_per_step_pipelined"] = sarpost.ops.last_launch_count()
            p_out, p_counts = pl.submit(sets[0], spec, **kw)
            pl.wait()
            torch.cuda.synchronize()
            o_ref, c_ref = sarpost.postprocess_fused(sets[0], spec, return_padded=True, **kw)
            res["pipe_same"] = bool(torch.equal(p_counts, c_ref)) and all(bool(torch.equal(p_out[b, :n], o_ref[b, :n])) for b, n in enumerate(c_ref.tolist()))
            pl.close()
            del out_ring
        # pass over the same steps with CUDA events around every kernel (recorded by the library on the launching stream, no
        # host sync per step, mean read afterwards).  Kept out of the passes `value` comes from: timing events between
        # kernels cost a few % by removing the overlap of consecutive launches.
        res["stage"] = stage_means(cx, plan_step, with_stage_steps)
        res["step"] = plan_step
        res["plan"] = plan
        return res

    def summarize(res, steps):
        vals = cx.max_over_ranks(res["ms_single"], res["ms_two"] or 0.0, res["ms_pipe"] or 0.0, res["ms_general"])
        rate = lambda ms_: bs * n_gpus * steps / (ms_ / 1e3)  # noqa: E731
        out_ = {"single_stream": {"value": rate(vals[0]), "unit": "images/s", "ms_per_step": vals[0] / steps,
                                  "rows_identical_to_general_call": res["plan_same"],
                                  "note": "strictly one batch in flight (all K steps on one stream), each step one call of a prepared "
                                          "plan (sarpost.FusedPlan -> sarpost_plan_run)"},
                "single_stream_general": {"value": rate(vals[3]), "unit": "images/s", "ms_per_step": vals[3] / steps,
                                          "note": "the same through the general entry point (sarpost.postprocess_fused -> sarpost_fused: "
                                                  "arguments validated, tensor maps encoded, outputs allocated on every call)"}}
        if res["ms_two"] is not None:
            out_["two_streams"] = {"value": rate(vals[1]), "unit": "images/s", "ms_per_step": vals[1] / steps,
                                   "note": f"the same K steps issued round-robin on {args.streams} independent CUDA streams through sarpost_fused"}
            out_["pipelined"] = {"value": rate(vals[2]), "unit": "images/s", "ms_per_step": vals[2] / steps,
                                 "rows_identical_to_plain_call": res["pipe_same"],
                                 "note": "sarpost_pipeline_submit per step (depth 2), one sarpost_pipeline_wait at the end"}
        st_ = res["stage"]
        out_["stage_ms"] = {"k1_candidates": st_[0], "k2_k4_select_sort_nms": st_[1], "k5_gather": st_[2], "whole_call": st_[3]}
        return out_

    # ---- timed region: K steps, CUDA events on the launching (current) stream ----
    clocks = ClockSampler(cx.local)
    clocks.start()
    head_sets = split_sets if use_split else level_sets
    main = measure_layout(head_sets, args.steps, 3 if args.quick else args.steps)
    step, counts, stage = main["step"], main["counts"], main["stage"]
    if len(clocks.samples) < 20 and not args.quick:  # short region: keep sampling over the same step to have clocks under load
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            for _ in range(20):
                step()
            torch.cuda.synchronize()
    clocks.stop()
    main_sum = summarize(main, args.steps)
    best = main_sum.get("pipelined") or main_sum["single_stream"]
    headline_mode = "pipelined" if best is not main_sum["single_stream"] else "single_stream"
    if main_sum["single_stream"]["value"] > best["value"]:
        best, headline_mode = main_sum["single_stream"], "single_stream"  # small workloads: one prepared call per step beats the pipeline's stream hops
    value, ms = best["value"], best["ms_per_step"] * args.steps
    single, two_streams = main_sum["single_stream"], main_sum.get("two_streams")
    launches_per_step = (main.get("launches_per_step_pipelined") if headline_mode == "pipelined" else None) or main["launches_per_step_plan"]
    pipe_same_main = main["pipe_same"]
    n_det = int(counts.sum().item())
    # the concatenated layout (one (B, no, H, W) tensor per level, as the unpatched head returns them), same passes
    cat_layout = None
    if use_split and not args.quick:
        cat_steps = max(10, min(args.steps, 200))
        cat = measure_layout(level_sets, cat_steps, min(cat_steps, 100))
        cat_layout = summarize(cat, cat_steps)
        o_s, c_s = sarpost.postprocess_fused(split_sets[0], spec, return_padded=True, **kw)
        o_c, c_c = sarpost.postprocess_fused(level_sets[0], spec, return_padded=True, **kw)
        cat_layout["rows_identical_to_split_layout"] = bool(torch.equal(c_s, c_c)) and all(
            bool(torch.equal(o_s[b, :n], o_c[b, :n])) for b, n in enumerate(c_c.tolist()))
        cat_layout["steps"] = cat_steps
        cat_layout["layout"] = "per level one (B, no, H, W) tensor: box | cls | emb | state concatenated (head.py:204-206)"
        del cat

    # ---- K1 roofline from the evented pass ----
    n_cand = count_candidates(torch, levels, nc, kw)
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

    # ---- clustered leg (same shapes / thresholds, suppression-heavy inputs) ----
    clustered = None
    if not args.no_clustered and not args.blobs:
        clustered = guarded("clustered", lambda: leg_clustered(cx, spec, shapes, nc, ed, sc, bs, kw, cls_mean, max(10, min(args.steps, 100)), use_split))

    # ---- the reference's own GPU path on the same box (rank 0, N = 1 only) ----
    ref_gpu = None
    if rank == 0 and n_gpus == 1 and not args.no_reference_gpu:
        n_img = min(bs, 4)

        def _ref_gpu():
            ours = sarpost.postprocess_fused([x[:n_img].contiguous() for x in levels], spec, **kw)
            r_ = leg_reference_gpu(cx, levels, strides, nc, ed, sc, kw, n_img, 5, ours)
            r_["sarpost_over_reference_gpu"] = {"device_value": value / r_["value"], "single_stream": single["value"] / r_["value"]}
            return r_

        ref_gpu = guarded("reference_gpu", _ref_gpu)

    # ---- e2e: HOST buffers through the C-ABI host entry (H2D + D2H inside the timed region) ----
    e2e = None if args.no_e2e else guarded("e2e", lambda: leg_e2e(cx, levels, spec, bs, kw))

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        n_img = args.cpu_images or (4 if args.workload == "cfg3" else min(bs, 16))
        n_img = min(n_img, bs)

        def _cpu():
            levels_cpu = [x[:n_img].cpu() for x in levels]
            torch.set_num_threads(cpu_threads())
            ref = ReferencePath(strides, nc, ed, sc, "cpu")
            ref.step([x[:1] for x in levels_cpu], kw)  # warm-up (lazy torchvision import)
            t0 = time.perf_counter()
            ref.step(levels_cpu, kw)
            secs = time.perf_counter() - t0
            return {"value": n_img / secs, "unit": "images/s", "cores": cpu_threads(), "kind": ref.kind,
                    "sample": f"first {n_img} images of the GPU batch, {secs:.1f} s; {ref.describe()}; {cpu_threads()} torch threads, host has {os.cpu_count()} cpus"}

        cpu = guarded("cpu_baseline", _cpu)

    # ---- sliced inference (cfg4): the one exchange step of the path, measured whenever there is more than one rank ----
    sahi = None
    if not args.no_sahi:
        del levels, level_sets, split_sets, head_sets, main, step
        torch.cuda.empty_cache()
        sahi = guarded("sahi", lambda: leg_sahi(cx, max(10, min(args.steps, 100))))

    if rank == 0:
        line = {
            "metric": "post-processed images/sec (decode+NMS)", "value": value, "unit": "images/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}" + (f" [clustered inputs, blobs={args.blobs}]" if args.blobs else ""),
                       "images_per_gpu": bs, "global_batch": bs * n_gpus, "anchors": anchors, "channels": spec.no, "streams": args.streams,
                       "value_is": headline_mode,
                       "in_flight": ("one batch" if args.streams == 1 else
                                     f"software pipeline of depth {PIPE_DEPTH} (sarpost_pipeline_submit per step, one sarpost_pipeline_wait at the end): every "
                                     "step is the whole path for one batch; the decode kernels of successive batches run back to back on one "
                                     "stream, the NMS + gather kernels of batch i on a high-priority stream of their own under the decode kernel "
                                     f"of a later batch; steps rotate over their own input buffers; rows identical to a plain call: {pipe_same_main}; "
                                     f"`two_streams` = round-robin on {args.streams} plain streams, `single_stream` = one batch in flight"),
                       "input_layout": ("split (SARPOST_LAYOUT_SPLIT): per level box (B,64,H,W), cls (B,nc,H,W), emb (B,H,W,E) channels_last, state "
                                        "(B,S,H,W) — what patch(fused=True, split=True) hands over instead of torch.cat (head.py:204-206); the "
                                        "reference legs (reference_gpu, cpu_baseline, --impl reference) get the concatenated tensors holding the "
                                        "same values; `cat_layout` = this arm on those concatenated tensors") if use_split else
                                       "concatenated (B, no, H, W) tensor per level",
                       "parallelism": f"batch-sharded x{n_gpus}, no data-path collective",
                       "l2": (("one batch of inputs (hot channels %.0f MB) exceeds the 126 MB L2; no flush" % (hot_bytes / 1e6))
                              if hot_bytes >= 2 * L2_BYTES else
                              ("hot channels %.1f MB per batch: steps rotate over %d identical input copies (%.0f MB in rotation > 2x the "
                               "126 MB L2), so every step reads its inputs from HBM; no flush" % (hot_bytes / 1e6, n_sets, n_sets * hot_bytes / 1e6))),
                       "input_sets": n_sets,
                       "candidates_per_image": n_cand / bs, "detections_per_image": n_det / bs},
            "single_stream": single, "single_stream_general": main_sum.get("single_stream_general"), "two_streams": two_streams, "roofline": roofline, "cat_layout": cat_layout, "clustered": clustered, "reference_gpu": ref_gpu, "cpu_baseline": cpu, "e2e": e2e,
            "sahi": sahi, "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step, "clocks": clocks.summary(),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
