"""Benchmark of the rotated-box geometry hot path (BASELINE.json metric) -- one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Headline workload (named in ``config.workload``): BASELINE config 4, the anchor target-assignment sweep -- per step 16 frames
x boxes_iou_bev(211 200 KITTI 3-class anchors, 100 GT boxes) = 3.38e8 rotated-IoU pairs, 1.35 GB of float32 results.

  N = 1   one frame-batched launch per step, the matrix resident in HBM (the reference's frame loop as one grid).
  N > 1   STRONG scaling of the same 16-frame problem, one process per GPU: the anchors are row-sharded (glenet_b200.sharded),
          every rank writes its slab of the matrix and the step also delivers what the assigner needs from the OTHER ranks --
          the column maxima / first rows over all anchors (axis_aligned_target_assigner.py:141-165) -- exchanged INSIDE the
          IoU kernel (system-scope atomics into CUDA-IPC windows, no NCCL call on the data path).  ``extra.sharded`` reports,
          beside it, compute-only, the same reductions through NCCL, and the full gather (fused coordinate lists vs NCCL
          all-gather of the dense slabs, with the NVLink bound).

``value``  rotated-IoU pairs/s, whole job, inputs resident in HBM, CUDA events, max over ranks.
``e2e``    same metric through the public API with HOST buffers: boxes copied H2D from pinned memory and the IoU matrix
           (this rank's slab for N > 1) copied D2H into pinned memory, all inside the timed region; ``e2e.matrix_free`` is the
           same sweep through the reductions-only API (what the assigner consumes), host buffers both ways.
``roofline`` dominant kernel iou_tile_kernel: HBM-write bound, 4 B per pair; duration from CUDA events
           around back-to-back launches of exactly that kernel on this rank's share.
``cpu_baseline`` the reference's own CPU implementation (oracle/_ref, boxes_iou_bev_cpu) on the host cores of this box,
           rows split over processes; bounded sample, rank 0, N = 1 only.
``extra``  the other configs of the path (points_in_boxes pts/s with its own CPU baseline and e2e, NMS, dense IoU, cfg0
           latencies, variance-voting NMS, the reference's GPU kernels on the same box), each with its roofline fraction.
``--impl reference`` times only the reference CPU implementation on the same config/metric.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES = 16
N_ANCHORS, N_GT = 211200, 100
PAIRS_PER_FRAME = N_ANCHORS * N_GT
BYTES_PER_PAIR = 4.0           # the culled sweep is bound by the float32 result write
WORKLOAD = "cfg4 anchor sweep: 16 frames x boxes_iou_bev(211200 KITTI anchors x 100 GT)"
# dram__bytes_read.sum + dram__bytes_write.sum of the dominant launch from the ncu --set full capture in profiles/
TRAFFIC_PER_LAUNCH = 1.314e9
TRAFFIC_NOTE = ("profiles/r02_iou_frames_summary.txt: 1.305 GB written + 0.009 GB read per 16-frame launch on one GPU vs 1.352 GB "
                "algorithmic (the bulk zero fill writes whole lines once; the box loads carry an L2 evict-last policy)")
NVLINK_GBS = 770.0             # measured peer-copy bandwidth per direction (B200_PROFILING.md)
PIB_B, PIB_M, PIB_N = 128, 180000, 200


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


FP32_DERIVED_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # 74.4; the measured value comes from tools/cuda/ffma_peak.cu


def flop_model():
    """Algorithmic FLOPs of the seeded inputs (tools/flop_model.py, SURVEY 8d F_pair); constants of the benchmark."""
    path = os.path.join(ROOT, "profiles", "flop_model.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f)
    return {}


def measured_fp32_peak(device_index):
    exe = os.path.join(ROOT, "tools", "cuda", "bin", "ffma_peak")
    if not os.path.isfile(exe):
        return None
    try:
        out = subprocess.run([exe, str(device_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=60).stdout
        return json.loads(out.strip().splitlines()[-1])
    except Exception:
        return None


# ------------------------------------------------------------------ CPU reference arm / baselines
_G = {}   # inherited by the forked CPU workers (no per-job pickling of the 5.9 MB anchor table)


def _cpu_rows_worker(args):
    f, lo, hi = args
    import torch
    a_np, b_np = _G["anchors"], _G["gts"][f]
    if _G["kind"] == "reference":
        out = torch.zeros((hi - lo, b_np.shape[0]), dtype=torch.float32)
        _G["ext"].boxes_iou_bev_cpu(torch.from_numpy(a_np[lo:hi]), torch.from_numpy(b_np), out)
        return float(out.sum())
    from oracle import capi
    return float(capi.boxes_iou_bev(a_np[lo:hi], b_np, dialect=capi.CPU).sum())


def _cpu_pib_worker(args):
    f, lo, hi = args
    import torch
    boxes, pts = _G["pib_boxes"][f], _G["pib_pts"][f][lo:hi]
    if _G["kind"] == "reference":
        out = torch.zeros((boxes.shape[0], hi - lo), dtype=torch.int32)
        _G["pib_ext"].points_in_boxes_cpu(torch.from_numpy(boxes), torch.from_numpy(pts), out)
        return int(out.sum())
    from oracle import capi
    return int(capi.points_in_boxes_mask(pts, boxes, dialect=capi.CPU).sum())


class CpuArm:
    """The reference's CPU implementation of the path on all host cores (rows / points split over forked processes)."""

    def __init__(self, with_pib=False):
        import multiprocessing as mp
        import numpy as np
        import torch
        from oracle import capi, ref
        from glenet_b200 import synth
        torch.set_num_threads(1)
        self.kind = "reference" if ref.available() else "port"
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        _G["kind"] = self.kind
        _G["anchors"] = np.ascontiguousarray(synth.anchors_kitti3().numpy())
        _G["gts"] = [np.ascontiguousarray(synth.kitti_boxes(N_GT, 100 + f + 1).numpy()) for f in range(FRAMES)]
        self.pib_frames = 2
        if with_pib:
            bx = [synth.waymo_boxes(PIB_N, 100 + f) for f in range(self.pib_frames)]
            _G["pib_boxes"] = [np.ascontiguousarray(b.numpy()) for b in bx]
            _G["pib_pts"] = [np.ascontiguousarray(synth.points(PIB_M, bx[f], synth.WAYMO_RANGE, 0.05, seed=500 + f).numpy()) for f in range(self.pib_frames)]
        if self.kind == "reference":
            _G["ext"] = ref.iou3d_nms_cuda()      # the unmodified reference extension (oracle/_ref)
            _G["pib_ext"] = ref.roiaware_pool3d_cuda()
        else:
            capi.load()
        self.pool = mp.get_context("fork").Pool(self.cores)     # forked BEFORE any CUDA initialisation

    def run_frames(self, frames):
        """IoU of `frames` frames; returns seconds."""
        t0 = time.perf_counter()
        n = _G["anchors"].shape[0]
        chunks = self.cores * 4
        step = (n + chunks - 1) // chunks
        for f in frames:
            self.pool.map(_cpu_rows_worker, [(f % FRAMES, lo, min(n, lo + step)) for lo in range(0, n, step)], chunksize=1)
        return time.perf_counter() - t0

    def run_pib(self, frames):
        """points_in_boxes_cpu (roiaware_pool3d.cpp:143-168) of `frames` cfg2 frames, points split over the cores; seconds."""
        t0 = time.perf_counter()
        chunks = self.cores * 2
        step = (PIB_M + chunks - 1) // chunks
        for f in frames:
            self.pool.map(_cpu_pib_worker, [(f % self.pib_frames, lo, min(PIB_M, lo + step)) for lo in range(0, PIB_M, step)], chunksize=1)
        return time.perf_counter() - t0

    def cfg0_single_thread(self):
        """BASELINE config 0 as shipped (one thread): boxes_bev_iou_cpu 200 x 50, points_in_boxes_cpu 120k x 20; median ms of 5."""
        import torch
        from glenet_b200 import synth
        from oracle import ref, capi
        a, b = synth.kitti_boxes(200, 0), synth.kitti_boxes(50, 1)
        bx = synth.kitti_boxes(20, 2)
        pts = synth.points(120000, bx, synth.KITTI_RANGE, 0.05, seed=3)

        def med(fn):
            ts = []
            for _ in range(5):
                t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
            return 1e3 * statistics.median(ts)
        if self.kind == "reference":
            return {"boxes_bev_iou_cpu_200x50_ms": med(lambda: ref.boxes_bev_iou_cpu(a, b)),
                    "points_in_boxes_cpu_120k_x20_ms": med(lambda: ref.points_in_boxes_cpu(pts, bx)), "kind": "reference", "threads": 1}
        return {"boxes_bev_iou_cpu_200x50_ms": med(lambda: capi.boxes_iou_bev(a, b, dialect=capi.CPU)),
                "points_in_boxes_cpu_120k_x20_ms": med(lambda: capi.points_in_boxes_mask(pts, bx, dialect=capi.CPU)), "kind": "port", "threads": 1}

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference_arm(args, rank):
    if rank != 0:
        return
    arm = CpuArm()
    arm.run_frames([0])                          # one warm-up frame is enough for a CPU loop
    t = arm.run_frames(list(range(args.steps)))  # a "step" of this arm = one frame (bounded sample of the 16-frame step)
    arm.close()
    value = args.steps * PAIRS_PER_FRAME / t
    line = {
        "impl": "reference", "metric": "rotated_iou_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": "each step = 1 of the 16 frames (211200 x 100 pairs)"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": arm.cores, "kind": arm.kind,
                         "sample": f"{args.steps} frames of 211200x100 pairs, rows split over {arm.cores} processes"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks, NUMA
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa(torch, local_rank):
    """Run this rank (and first-touch its pinned buffers) on the NUMA node its GPU hangs off.  Best effort: a container whose
    cpuset does not contain that node's CPUs keeps its affinity; the outcome is reported in ``config.numa``."""
    info = {"node": None, "cpus_bound": None, "mempolicy": None}
    try:
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["node"] = node
        if node < 0:
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus_bound"] = len(allowed)
        try:   # set_mempolicy(MPOL_PREFERRED = 1, mask, maxnode): pinned allocations that follow land on this node
            mask = ctypes.c_ulong(1 << node)
            rc = ctypes.CDLL(None, use_errno=True).syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))
            info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
        except Exception as e:   # pragma: no cover
            info["mempolicy"] = f"unavailable ({type(e).__name__})"
    except Exception as e:
        info["error"] = f"{type(e).__name__}: {e}"
    return info


# ------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local_rank):
    cpu_base = cpu_pib = cfg0_cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(with_pib=not args.no_extra)     # fork the CPU workers before CUDA is touched
        arm.run_frames([0])
        nfr = 4
        t = arm.run_frames(list(range(nfr)))
        cpu_base = {"value": nfr * PAIRS_PER_FRAME / t, "unit": "pairs/s", "cores": arm.cores, "kind": arm.kind,
                    "sample": f"{nfr} of the 16 frames (211200x100 pairs each), reference boxes_iou_bev_cpu, rows split over {arm.cores} processes"}
        if not args.no_extra:
            arm.run_pib([0])
            npf = 6
            t = arm.run_pib(list(range(npf)))
            cpu_pib = {"value": npf * PIB_M / t, "unit": "points/s", "cores": arm.cores, "kind": arm.kind,
                       "sample": f"{npf} frames of 180000 points x 200 boxes, reference points_in_boxes_cpu (roiaware_pool3d.cpp:143-168), points split over {arm.cores} processes"}
            cfg0_cpu = arm.cfg0_single_thread()
        arm.close()

    import torch
    import torch.distributed as dist
    from glenet_b200 import iou3d_nms_utils as I, roiaware_pool3d_utils as R, sharded, synth
    import glenet_b200

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(torch, local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    glenet_b200.load()
    hbm_gbs, peak_src = measured_peaks()
    fp32 = measured_fp32_peak(local_rank) if (rank == 0 and world == 1 and not args.no_extra) else None

    anchors_h = synth.anchors_kitti3().pin_memory()
    gts_h = torch.stack([synth.kitti_boxes(N_GT, 100 + f + 1) for f in range(FRAMES)]).pin_memory()   # the same problem on every rank
    anchors, gts = anchors_h.to(dev), gts_h.to(dev)
    start, stop = sharded.shard_rows(N_ANCHORS, world, rank)
    rows = stop - start
    slab_anchors = anchors[start:stop].contiguous()
    out_slab = torch.empty((FRAMES, rows, N_GT), dtype=torch.float32, device=dev)      # this rank's share of the 1.35 GB matrix
    exchange = "none (single GPU: one frame-batched launch, matrix resident)"
    win = None
    if world > 1:
        try:
            win = sharded.ExchangeWindow(frames=FRAMES, nb=N_GT, list_cap=(FRAMES * sharded.slab_rows(N_ANCHORS, world) * N_GT) // 48)
            exchange = ("in-kernel: row-sharded slabs stay resident; every rank's column keys (max, first row) go into its slot of every "
                        "peer's CUDA-IPC exchange window as plain 16-byte stores over NVLink + a step flag, the maximum over the slots is "
                        "decoded locally (one PDL-chained exchange kernel behind the IoU kernel; no NCCL on the data path)")
        except Exception as e:   # IPC unavailable: the torch.distributed formulation carries the exchange
            win = None
            exchange = f"NCCL all_reduce of the column keys (CUDA IPC exchange windows unavailable: {type(e).__name__}: {e})"
    launches = [0]

    def step_dense():
        # compute only: this rank's slab, resident (the whole matrix on one GPU)
        I.boxes_iou_bev_frames(slab_anchors, gts, out=out_slab)
        launches[0] += 1

    def step_assign():
        # slab + the assigner's reductions, column keys exchanged inside the kernel
        sharded.anchor_assign_sharded(anchors, gts, win, out=out_slab)
        launches[0] += 2

    def step_assign_nccl():
        # the torch.distributed formulation of the same result: dense slab, torch reductions, two all_reduce calls
        I.boxes_iou_bev_frames(slab_anchors, gts, out=out_slab)
        row_max, row_arg = out_slab.max(dim=2)
        col_max, col_arg = out_slab.max(dim=1)
        col_arg = col_arg + start
        gmax = col_max.clone()
        dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
        cand = torch.where(col_max == gmax, col_arg, torch.full_like(col_arg, N_ANCHORS))
        dist.all_reduce(cand, op=dist.ReduceOp.MIN)
        return row_max, row_arg, gmax, cand

    headline = step_dense if world == 1 else (step_assign if win is not None else step_assign_nccl)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches[0] = 0
    ms = timed(headline, args.steps, args.warmup)
    timed_launches = (launches[0] * args.steps) // (args.steps + args.warmup)
    clk = clocks.stop() if rank == 0 else None
    pairs_step = FRAMES * PAIRS_PER_FRAME
    value = pairs_step * args.steps / (ms * 1e-3)

    # dominant kernel alone: back-to-back dense launches of this rank's share (events on the launching stream)
    side_steps = max(3, min(args.steps, 20))
    ms_dense = timed(step_dense, side_steps, 3) / side_steps
    slab_bytes = FRAMES * rows * N_GT * BYTES_PER_PAIR
    achieved = slab_bytes / (ms_dense * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "iou_tile_kernel<IOU_BEV, dense>", "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
                "frac": achieved / hbm_gbs, "traffic": TRAFFIC_PER_LAUNCH * rows / N_ANCHORS, "traffic_note": TRAFFIC_NOTE, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": slab_bytes, "units_per_launch": f"{FRAMES} frames x {rows} rows x {N_GT} columns (this rank's slab)",
                "avg_launch_ms": ms_dense}
    if world == 1:
        def step_per_frame():
            for f in range(FRAMES):
                I.boxes_iou_bev_frames(anchors, gts[f:f + 1], out=out_slab[f:f + 1])
        ms_pf = timed(step_per_frame, side_steps, 2) / side_steps
        roofline["per_frame_launches"] = {"ms_per_frame": ms_pf / FRAMES, "pairs_per_s": pairs_step / (ms_pf * 1e-3), "frac": pairs_step * 4 / (ms_pf * 1e-3) / 1e9 / hbm_gbs,
                                          "note": "the same work as 16 single-frame launches (the reference's call sequence), for comparison"}

    # ---- the sharded sweep, every way of getting results across (all ranks take part; also run at N = 1 for the scaling table)
    shard = {"rows_per_rank": rows, "compute_only_ms": ms_dense}
    if win is None and world == 1:
        win1 = sharded.ExchangeWindow(frames=FRAMES, nb=N_GT, list_cap=pairs_step // 48)
    else:
        win1 = win
    if win1 is not None:
        def step_assign_w():
            sharded.anchor_assign_sharded(anchors, gts, win1, out=out_slab)
        shard["gathered_reductions_fused_ms"] = timed(step_assign_w, side_steps, 3) / side_steps

        def step_keys_only():
            sharded.anchor_assign_sharded(anchors, gts, win1, dense=False)
        shard["gathered_reductions_matrix_free_ms"] = timed(step_keys_only, side_steps, 3) / side_steps
    if world > 1:
        shard["gathered_reductions_nccl_ms"] = timed(step_assign_nccl, side_steps, 3) / side_steps
        # weak-scaling form of the same exchange: the batch grows with the GPUs (16 frames per GPU, as DDP ranks bring them), the
        # anchors stay row-sharded, so every rank clips as many pairs as the single GPU does and the column keys of ALL
        # world x 16 frames cross NVLink inside the kernel
        try:
            gts_w = torch.stack([synth.kitti_boxes(N_GT, 100 + f + 1) for f in range(FRAMES * world)]).to(dev)
            win_w = sharded.ExchangeWindow(frames=FRAMES * world, nb=N_GT, list_cap=0)
            out_w = torch.empty((FRAMES * world, rows, N_GT), dtype=torch.float32, device=dev)

            def step_assign_weak():
                sharded.anchor_assign_sharded(anchors, gts_w, win_w, out=out_w)
            ms_w = timed(step_assign_weak, side_steps, 3) / side_steps
            shard["weak_batch_gathered_reductions_fused_ms"] = ms_w
            shard["weak_batch_pairs_per_s"] = FRAMES * world * PAIRS_PER_FRAME / (ms_w * 1e-3)
            shard["weak_batch_note"] = f"{FRAMES * world} frames (16 per GPU) x 211200 anchors row-sharded over {world} GPUs, slab + assigner reductions, in-kernel exchange"
            win_w.close()
            del gts_w, out_w
        except Exception as e:   # a side measurement must not take the headline down
            shard["weak_batch_error"] = f"{type(e).__name__}: {e}"
    if win1 is not None and not args.no_gather:
        full = torch.empty((FRAMES, N_ANCHORS, N_GT), dtype=torch.float32, device=dev)
        fill_stream = torch.cuda.Stream(device=dev)

        def step_gather():
            sharded.boxes_iou_gather_sharded(anchors, gts, win1, out=full, fill_stream=fill_stream)
        gsteps = max(3, min(args.steps, 10))
        shard["full_gather_fused_ms"] = timed(step_gather, gsteps, 2) / gsteps
        shard["full_gather_fused_note"] = "every rank zero-fills its own copy (HBM bound: 1.35 GB per GPU) and only the non-zero elements cross NVLink as coordinate lists"
        if world > 1:
            per = sharded.slab_rows(N_ANCHORS, world)
            padded = torch.zeros((FRAMES, per, N_GT), dtype=torch.float32, device=dev)
            gathered = full.view(-1)[: world * FRAMES * per * N_GT].view(world, FRAMES, per, N_GT) if world * per <= N_ANCHORS else \
                torch.empty((world, FRAMES, per, N_GT), dtype=torch.float32, device=dev)

            def step_gather_nccl():
                if rows == per:
                    I.boxes_iou_bev_frames(slab_anchors, gts, out=padded)
                else:   # the last rank's slab is shorter than the 64-row-aligned slab size: its rows go into the padded block
                    I.boxes_iou_bev_frames(slab_anchors, gts, out=out_slab)
                    padded[:, :rows].copy_(out_slab)
                dist.all_gather_into_tensor(gathered, padded)
            shard["full_gather_nccl_ms"] = timed(step_gather_nccl, 3, 1) / 3
            recv = (world - 1) / world * pairs_step * 4
            shard["full_gather_nvlink_bound_ms"] = recv / (NVLINK_GBS * 1e9) * 1e3
            shard["full_gather_nvlink_note"] = f"(N-1)/N x 1.35 GB received per GPU at the measured {NVLINK_GBS:.0f} GB/s per direction"
            del padded, gathered
        shard["exchange_status"] = win1.status()
        del full

    # ---- end to end through the public API, host buffers both ways
    out_h = torch.empty((FRAMES, rows, N_GT), dtype=torch.float32).pin_memory()
    up_stream = torch.cuda.Stream(device=dev)
    if world == 1:
        def step_e2e():
            # per frame: boxes host -> device, the drop-in call, matrix device -> host.  The upload of frame f+1 runs on a second
            # stream while frame f is computed and downloaded (PCIe is full duplex); every byte crosses inside the timed region.
            main = torch.cuda.current_stream()

            def upload(f):
                with torch.cuda.stream(up_stream):
                    a = anchors_h.to(dev, non_blocking=True)
                    g = gts_h[f].to(dev, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(up_stream)
                return a, g, done
            alive = []
            nxt = upload(0)
            for f in range(FRAMES):
                a_d, g_d, done = nxt
                if f + 1 < FRAMES:
                    nxt = upload(f + 1)
                main.wait_event(done)
                iou = I.boxes_iou_bev(a_d, g_d)
                out_h[f].copy_(iou, non_blocking=True)
                alive.append((a_d, g_d, iou))
            main.synchronize()
            up_stream.synchronize()
        h2d, d2h = FRAMES * (N_ANCHORS + N_GT) * 28, FRAMES * PAIRS_PER_FRAME * 4
        e2e_api = "glenet_b200.iou3d_nms_utils.boxes_iou_bev per frame, pinned host boxes in, pinned host (211200,100) matrix out"
    else:
        red_h = [torch.empty((FRAMES, rows), dtype=torch.float32).pin_memory(), torch.empty((FRAMES, rows), dtype=torch.int64).pin_memory(),
                 torch.empty((FRAMES, N_GT), dtype=torch.float32).pin_memory(), torch.empty((FRAMES, N_GT), dtype=torch.int64).pin_memory()]

        def step_e2e():
            a_d = anchors_h.to(dev, non_blocking=True)
            g_d = gts_h.to(dev, non_blocking=True)
            if win is not None:
                res = sharded.anchor_assign_sharded(a_d, g_d, win, out=out_slab)
                vecs = (res["row_max"], res["row_argmax"], res["col_max"], res["col_argmax"])
            else:
                vecs = step_assign_nccl()
            out_h.copy_(out_slab, non_blocking=True)
            for dst, src in zip(red_h, vecs):
                dst.copy_(src, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d = (N_ANCHORS + FRAMES * N_GT) * 28
        d2h = FRAMES * rows * N_GT * 4 + FRAMES * rows * 12 + FRAMES * N_GT * 12
        e2e_api = "glenet_b200.sharded.anchor_assign_sharded, pinned host boxes in, this rank's pinned host slab + reductions out (bytes per rank)"
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(step_e2e, e2e_steps, 1)
    e2e = {"value": pairs_step * e2e_steps / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": ms_e2e / e2e_steps, "api": e2e_api}
    # the same sweep through the reductions-only API: what the assigner consumes, nothing else crosses PCIe
    mf_h = [torch.empty((FRAMES, rows), dtype=torch.float32).pin_memory(), torch.empty((FRAMES, rows), dtype=torch.int64).pin_memory(),
            torch.empty((FRAMES, N_GT), dtype=torch.float32).pin_memory(), torch.empty((FRAMES, N_GT), dtype=torch.int64).pin_memory()]

    def step_e2e_mf():
        a_d = anchors_h.to(dev, non_blocking=True)
        g_d = gts_h.to(dev, non_blocking=True)
        if win1 is not None:
            res = sharded.anchor_assign_sharded(a_d, g_d, win1, dense=False)
            vecs = (res["row_max"], res["row_argmax"], res["col_max"], res["col_argmax"])
        else:
            vecs = I.iou_max_overlaps_frames(a_d[start:stop].contiguous(), g_d)
        for dst, src in zip(mf_h, vecs):
            dst.copy_(src, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_mf = timed(step_e2e_mf, max(3, e2e_steps), 1) / max(3, e2e_steps)
    e2e["matrix_free"] = {"value": pairs_step / (ms_mf * 1e-3), "unit": "pairs/s", "ms_per_step": ms_mf,
                          "h2d_bytes_per_step": (N_ANCHORS + FRAMES * N_GT) * 28, "d2h_bytes_per_step": FRAMES * rows * 12 + FRAMES * N_GT * 12,
                          "api": "anchor_assign_sharded(dense=False): row / column max + argmax only (axis_aligned_target_assigner.py:141-165), host buffers both ways"}

    # ---- the per-frame operations of the metric at N > 1: replicas only (frames are independent, no collective): every rank runs
    #      its own batch, time = max over ranks, throughput = all ranks' units / that time
    if world > 1 and not args.no_extra:
        try:
            rb = torch.stack([synth.waymo_boxes(PIB_N, 100 + PIB_B * rank + f) for f in range(PIB_B)])
            rp = torch.stack([synth.points(PIB_M, rb[f], synth.WAYMO_RANGE, 0.05, seed=500 + PIB_B * rank + f) for f in range(PIB_B)])
            rb, rp = rb.to(dev), rp.to(dev)
            ms_r = timed(lambda: R.points_in_boxes_gpu(rp, rb), 10, 3) / 10
            shard["points_in_boxes_replicas"] = {
                "workload": f"cfg2 on every GPU: {PIB_B} frames x {PIB_M} points x {PIB_N} boxes per rank (per-frame boxes and points, own seeds per rank)",
                "ms": ms_r, "value": world * PIB_B * PIB_M / (ms_r * 1e-3), "unit": "points/s", "scaling": "weak (replicas only: no data-path collective)",
                "frac_of_hbm_per_gpu": PIB_B * PIB_M * 16 / (ms_r * 1e-3) / 1e9 / hbm_gbs}
            del rb, rp
            nb_, ns_ = [], []
            for f in range(8):
                b_, s_ = synth.proposals(4096, 20, 20 + 8 * rank + f)
                nb_.append(b_); ns_.append(s_)
            nb_, ns_ = torch.stack(nb_).to(dev), torch.stack(ns_).to(dev)
            ms_n = timed(lambda: I.nms_gpu_batch(nb_, ns_, 0.7), 10, 3) / 10
            shard["nms_replicas"] = {"workload": "cfg1 on every GPU: 8 frames x nms_gpu(4096 proposals, thresh 0.7) per rank", "ms": ms_n,
                                     "value": world * 8 / (ms_n * 1e-3), "unit": "frames/s", "scaling": "weak (replicas only: no data-path collective)"}
            del nb_, ns_
        except Exception as e:   # a side measurement must not take the headline down
            shard["replicas_error"] = f"{type(e).__name__}: {e}"

    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        extra = side_runs(torch, I, R, synth, dev, hbm_gbs, fp32, cpu_pib, cfg0_cpu)
    if rank == 0:
        extra["sharded"] = shard
        line = {
            "metric": "rotated_iou_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD + (" on one GPU" if world == 1 else f", anchors row-sharded over {world} GPUs, assigner reductions gathered"),
                       "pairs_per_step": pairs_step, "frames": FRAMES, "rows_per_rank": rows,
                       "l2": f"each step writes {slab_bytes / 1e6:.0f} MB of results per GPU (> 126 MB L2), so successive launches stream through L2",
                       "exchange": exchange, "numa": numa},
            "clocks": clk, "gpu_launches": timed_launches,
            "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu_base, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    if win1 is not None:
        win1.close()
    if world > 1:
        dist.destroy_process_group()


def side_runs(torch, I, R, synth, dev, hbm_gbs, fp32, cpu_pib, cfg0_cpu):
    """Short device-timed runs of the other configs of the path (not the headline value)."""
    out = {}
    fm = flop_model()
    fp32_peak = fp32["fp32_tflops"] if fp32 else FP32_DERIVED_TFLOPS
    fp32_src = "measured: tools/cuda/ffma_peak.cu on this GPU (FFMA-only, burst)" if fp32 else "derived 148 SM x 128 lanes x 2 x 1.965 GHz"
    out["fp32_peak"] = {"measured": fp32, "derived_tflops": FP32_DERIVED_TFLOPS}

    def ev(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters

    def fp32_roof(flops, ms, note):
        t = flops / (ms * 1e-3) / 1e12
        return {"bound": "fp32", "achieved": t, "peak": fp32_peak, "unit": "TFLOP/s", "frac": t / fp32_peak, "peak_source": fp32_src,
                "frac_of_derived_74.4": t / FP32_DERIVED_TFLOPS, "algorithmic_flops": flops, "note": note}

    # cfg4 without the matrix: the assigner's reductions (row / column max + argmax) and the non-zero list
    anchors = synth.anchors_kitti3().to(dev)
    gts16 = torch.stack([synth.kitti_boxes(100, 101 + f) for f in range(16)]).to(dev)
    ms_sp = ev(lambda: I.boxes_iou_frames_sparse(anchors, gts16, "bev"), 10)
    ms_mx = ev(lambda: I.iou_max_overlaps_frames(anchors, gts16, "bev"), 10)
    npairs = 16 * anchors.shape[0] * 100
    out["anchor_sweep_sparse"] = {"workload": "cfg4 anchor sweep, 16 frames, non-zero IoU list instead of the dense matrix (incl. the host sync for its length)",
                                  "pairs_per_s_list": npairs / (ms_sp * 1e-3), "ms_list": ms_sp,
                                  "pairs_per_s_max_overlaps": npairs / (ms_mx * 1e-3), "ms_max_overlaps": ms_mx,
                                  "note": "row/col max+argmax (F,N)+(F,M) as consumed by axis_aligned_target_assigner.py:141-165; no 4 B/pair write"}
    del anchors, gts16
    # pcdet/ops/iou3d boxes_aligned_iou3d_gpu: predictions vs regression targets of the positive anchors (IoU-aware heads)
    from glenet_b200 import iou3d_utils as I1
    pred, tgt = synth.head_pairs(20000, 3)
    pred, tgt = pred.to(dev), tgt.to(dev)
    ms_v1 = ev(lambda: I1.boxes_aligned_iou3d_gpu(pred, tgt), 50)
    out["aligned_iou3d_heads"] = {"workload": "boxes_aligned_iou3d_gpu (pcdet/ops/iou3d), 20000 prediction/target pairs", "value": 20000 / (ms_v1 * 1e-3),
                                  "unit": "pairs/s", "ms": ms_v1}
    # cfg2: points_in_boxes, 128 frames x 180k points x 200 boxes, every frame its own boxes AND its own points (5 % resampled inside
    # that frame's boxes: SURVEY 8d generator); 276 MB of points > L2
    B, M, N = PIB_B, PIB_M, PIB_N
    boxes_h = torch.stack([synth.waymo_boxes(N, 100 + f) for f in range(B)])
    pts_h = torch.stack([synth.points(M, boxes_h[f], synth.WAYMO_RANGE, 0.05, seed=500 + f) for f in range(B)]).pin_memory()
    boxes_hp = boxes_h.pin_memory()
    boxes, pts = boxes_hp.to(dev), pts_h.to(dev)
    ms = ev(lambda: R.points_in_boxes_gpu(pts, boxes), 30)
    gbs = B * M * 16 / (ms * 1e-3) / 1e9
    inside = float((R.points_in_boxes_gpu(pts[:8], boxes[:8]) >= 0).float().mean())
    out_pib_h = torch.empty((B, M), dtype=torch.int32).pin_memory()

    def pib_e2e():
        r = R.points_in_boxes_gpu(pts_h.to(dev, non_blocking=True), boxes_hp.to(dev, non_blocking=True))
        out_pib_h.copy_(r, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms_pe = ev(pib_e2e, 3, 1)
    out["points_in_boxes"] = {"workload": "cfg2: 128 frames x 180000 points x 200 boxes, per-frame boxes and per-frame points (5 % resampled inside the frame's boxes)",
                              "value": B * M / (ms * 1e-3), "unit": "points/s", "ms": ms, "fraction_of_points_inside_a_box": inside,
                              "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs, "algorithmic_bytes_per_point": 16},
                              "e2e": {"value": B * M / (ms_pe * 1e-3), "unit": "points/s", "ms": ms_pe, "h2d_bytes": pts_h.numel() * 4 + boxes_hp.numel() * 4, "d2h_bytes": B * M * 4,
                                      "api": "roiaware_pool3d_utils.points_in_boxes_gpu, pinned host points + boxes in, pinned host (128, 180000) int32 out"},
                              "cpu_baseline": cpu_pib}
    del pts, boxes, pts_h, out_pib_h
    # cfg0: the KITTI-shape frame through the host-signature drop-ins (H2D + kernel + D2H inside), next to the reference on one host thread
    a0, b0 = synth.kitti_boxes(200, 0), synth.kitti_boxes(50, 1)
    bx0 = synth.kitti_boxes(20, 2)
    p0 = synth.points(120000, bx0, synth.KITTI_RANGE, 0.05, seed=3)

    def host_ms(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(9):
            torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return 1e3 * statistics.median(ts)
    out["cfg0_kitti_frame"] = {"workload": "cfg0: boxes_bev_iou_cpu 200 x 50 and points_in_boxes_cpu 120000 x 20, host tensors in and out (latency, PCIe / launch bound)",
                               "boxes_bev_iou_cpu_200x50_ms": host_ms(lambda: I.boxes_bev_iou_cpu(a0, b0)),
                               "points_in_boxes_cpu_120k_x20_ms": host_ms(lambda: R.points_in_boxes_cpu(p0, bx0)),
                               "reference_cpu": cfg0_cpu}
    # cfg1: NMS 4096 -> keep, thresh 0.7, batch 8 (batched launch pair and the per-frame drop-in loop)
    fb, fs = [], []
    for f in range(8):
        b, s = synth.proposals(4096, 20, 20 + f)
        fb.append(b); fs.append(s)
    fb, fs = torch.stack(fb).to(dev), torch.stack(fs).to(dev)
    ms_b = ev(lambda: I.nms_gpu_batch(fb, fs, 0.7), 10)
    ms_l = ev(lambda: [I.nms_gpu(fb[f], fs[f], 0.7)[0][:500] for f in range(8)], 5)
    out["nms"] = {"workload": "cfg1: 8 frames x nms_gpu(4096 proposals, thresh 0.7)", "frames_per_s_batched": 8 / (ms_b * 1e-3),
                  "frames_per_s_dropin_loop": 8 / (ms_l * 1e-3), "ms_batched": ms_b, "ms_dropin_loop": ms_l}
    if "cfg1_nms_8x4096_upper_triangle" in fm:
        out["nms"]["roofline"] = fp32_roof(fm["cfg1_nms_8x4096_upper_triangle"]["flops"], ms_b,
                                           "whole batched call (sort + spatial binning + mask kernel + deferred exact clips + per-component sweep; the mask kernel is ~40 % of it, "
                                           "profiles/r02_nms_summary.txt); FLOPs = SURVEY 8d model over the 8 x 8.39e6 upper-triangle pairs of the seeded input "
                                           "(profiles/flop_model.json) -- the work the reference does, not the pairs this path still examines")
    # GLENet's NMS_TYPE: variance-voting NMS on 4096 proposals (N x N CPU-dialect IoU + the voting loop, all on the device)
    vb, vs = synth.proposals(4096, 20, 31)
    vv = (torch.rand((4096, 7), generator=torch.Generator().manual_seed(5)) * 0.5 + 0.05)
    vb, vs, vv = vb.to(dev), vs.to(dev), vv.to(dev)
    ms_vn = host_ms(lambda: I.new_nms_gpu(vb, vs, 0.25, variance=vv))
    out["variance_voting_nms"] = {"workload": "new_nms_gpu (nms_func) on 4096 proposals, iou_threshold 0.25, with variances; wall clock of the call, tensors in, numpy out",
                                  "ms_per_frame": ms_vn, "kept": int(len(I.new_nms_gpu(vb, vs, 0.25, variance=vv)[0])),
                                  "note": "the reference spends ~3.4 s in boxes_bev_iou_cpu alone for this N (SURVEY 8f rank 1)"}
    # cfg3: CVAE 30 samples x 20k GT, 3D IoU (dense: every pair takes the clipping path)
    smp, gt = synth.cvae_samples(20000, 30, 0)
    smp, gt = smp.to(dev), gt.to(dev)
    ms = ev(lambda: I.boxes_iou3d_aligned(smp, gt, 30), 20)
    flops3 = fm.get("cfg3_aligned_600k", {}).get("flops", 600000 * 790.0)
    out["iou3d_cvae"] = {"workload": "cfg3: 600000 aligned pairs (30 samples x 20000 GT), all overlapping", "value": 600000 / (ms * 1e-3),
                         "unit": "pairs/s", "ms": ms,
                         "roofline": fp32_roof(flops3, ms, "FLOPs = SURVEY 8d model summed over the seeded 600000 pairs (675 per pair, profiles/flop_model.json); "
                                               "executed FP32 instructions and pipe utilisation: profiles/r02_iou_dense_summary.txt")}
    # cfg2: boxes_iou3d_gpu 4096 x 200
    g2 = synth.waymo_boxes(200, 2)
    pr, _ = synth.proposals(4096, seed=3, base=g2)
    pr, g2 = pr.to(dev), g2.to(dev)
    ms = ev(lambda: I.boxes_iou3d_gpu(pr, g2), 20)
    out["iou3d_4096x200"] = {"workload": "cfg2: boxes_iou3d_gpu 4096 x 200", "value": 4096 * 200 / (ms * 1e-3), "unit": "pairs/s", "ms": ms}
    out["reference_gpu_kernel"] = reference_gpu_kernels(torch, synth, dev, ev)
    out.update(next_rows(torch, R, synth, dev, ev, host_ms))
    return out


def next_rows(torch, R, synth, dev, ev, host_ms):
    """SURVEY 8f rank 4 + the second half of rank 2: KITTI-evaluator rotated IoU, GT-database crops, CVAE recall IoU."""
    import numpy as np
    from glenet_b200 import cvae_eval_utils as C, gt_database as G, rotate_iou as RI
    out = {}
    try:
        # one evaluation part of kitti eval (eval.py:344-400): ~75 frames, ~6 GT and ~11 detections each, dense (sum GT) x (sum det)
        rng = np.random.default_rng(3)
        frames = 75
        bc, qc = rng.integers(2, 11, frames), rng.integers(5, 18, frames)
        cen = [rng.uniform([0, -40], [70, 40], (int(bc[f]), 2)) for f in range(frames)]
        def mk(c, jit):
            n = c.shape[0]
            return np.concatenate([c + rng.normal(0, jit, (n, 2)), rng.uniform(3.5, 4.3, (n, 1)), rng.uniform(1.45, 1.75, (n, 1)), rng.uniform(-np.pi, np.pi, (n, 1))], 1).astype(np.float32)
        boxes = np.concatenate([mk(cen[f], 0.0) for f in range(frames)])
        query = np.concatenate([mk(cen[f][rng.integers(0, bc[f], int(qc[f]))], 0.4) for f in range(frames)])
        ms_dense = host_ms(lambda: RI.rotate_iou_gpu_eval(boxes, query, -1))
        ms_blocks = host_ms(lambda: RI.rotate_iou_gpu_eval_blocks(boxes, query, bc, qc, -1))
        lib = __import__("glenet_b200")._lib.load()
        d_b, d_q = torch.from_numpy(boxes).to(dev), torch.from_numpy(query).to(dev)
        d_o = torch.empty((boxes.shape[0], query.shape[0]), dtype=torch.float32, device=dev)
        ms_kernel = ev(lambda: lib.glenet_rotate_iou_eval_gpu(d_b.data_ptr(), boxes.shape[0], d_q.data_ptr(), query.shape[0], -1, d_o.data_ptr(),
                                                               torch.cuda.current_stream(dev).cuda_stream), 50)
        ri = {"workload": f"rotate_iou_gpu_eval on one KITTI evaluation part: {boxes.shape[0]} GT x {query.shape[0]} detections of {frames} frames (numpy in, numpy out, wall clock)",
              "ms_dense_call": ms_dense, "ms_blocks_call": ms_blocks, "ms_kernel_only": ms_kernel, "pairs": int(boxes.shape[0] * query.shape[0])}
        try:
            from oracle import ref
            mod = ref.rotate_iou_numba()
            if mod is not None:
                ri["reference_numba_ms_dense_call"] = host_ms(lambda: mod.rotate_iou_gpu_eval(boxes, query, -1))
        except Exception as e:   # the baseline must never take the bench down
            ri["reference_numba_unavailable"] = f"{type(e).__name__}: {e}"
        out["kitti_eval_rotate_iou"] = ri
        # GT-database crops: a KITTI frame (120k points x 20 objects, points_in_boxes_cpu rule) and a Waymo frame (180k x 200, points_in_boxes_gpu rule)
        crops = {}
        for rule, n_obj, n_pts, feats, gen, rg in (("kitti", 20, 120000, 4, synth.kitti_boxes, synth.KITTI_RANGE), ("waymo", 200, 180000, 5, synth.waymo_boxes, synth.WAYMO_RANGE)):
            bx = gen(n_obj, 7)
            pts = torch.cat([synth.points(n_pts, bx, rg, 0.05, seed=8), torch.rand((n_pts, feats - 3), generator=torch.Generator().manual_seed(1))], 1).numpy()
            bxn = bx.numpy().astype(np.float64 if rule == "kitti" else np.float32)
            ms_ours = host_ms(lambda: G.crop_gt_objects(pts, bxn, rule))
            def ref_path():
                # the reference's call sequence (kitti_dataset.py:248-254 / waymo_dataset.py:364-372) on the drop-in membership functions
                if rule == "kitti":
                    sel = R.points_in_boxes_cpu(torch.from_numpy(pts[:, 0:3]), torch.from_numpy(bxn)).numpy()
                    res = [pts[sel[i] > 0] for i in range(n_obj)]
                else:
                    idx = R.points_in_boxes_gpu(torch.from_numpy(pts[:, 0:3]).unsqueeze(0).float().to(dev), torch.from_numpy(bxn[:, 0:7]).unsqueeze(0).float().to(dev)).long().squeeze(0).cpu().numpy()
                    res = [pts[idx == i] for i in range(n_obj)]
                for i, r in enumerate(res):
                    r[:, :3] -= bxn[i, :3]
                return res
            crops[rule] = {"points": n_pts, "objects": n_obj, "ms_device_compaction": ms_ours, "ms_host_selection_after_dropin_membership": host_ms(ref_path)}
        crops["note"] = "per frame, host arrays in, per-object float32 rows out; the second figure keeps the reference's numpy selection loop on top of this package's membership kernels"
        out["gt_database_crops"] = crops
        # CVAE recall IoU: 20000 (GT, prediction) pairs
        smp, gt = synth.cvae_samples(20000, 1, 4)
        g_d, q_d = gt.to(dev), smp.reshape(-1, 7).to(dev)
        out["cvae_recall_iou3d"] = {"workload": "iou3d (cvae_uncertainty/eval_utils/eval_utils.py:14-65) on 20000 aligned pairs", "ms": ev(lambda: C.iou3d(g_d, q_d), 20),
                                    "note": "the reference runs Python loops over numpy float32 scalars on the host, milliseconds per pair (tests/golden/make_golden_cvae_iou3d.py)"}
    except Exception as e:   # side rows must never take the headline down
        out["next_rows_error"] = f"{type(e).__name__}: {e}"
    return out


def reference_gpu_kernels(torch, synth, dev, ev):
    """The reference's own CUDA kernels (oracle/_ref, compiled unmodified for sm_100a) timed on THIS GPU for the same
    inputs: the "existing kernel" bar of SURVEY 2.1.  A reported baseline next to the numbers above, never the product."""
    try:
        from oracle import ref
        if not ref.available():
            return {"unavailable": "oracle/_ref not built"}
        res = {"what": "reference kernels through oracle/ref.py (the reference's wrapper call sequence), CUDA events"}
        a = synth.anchors_kitti3().to(dev)
        g = synth.kitti_boxes(100, 101).to(dev)
        res["cfg4_boxes_iou_bev_one_frame_ms"] = ev(lambda: ref.boxes_iou_bev(a, g), 5, 2)
        g2 = synth.waymo_boxes(200, 2)
        pr = synth.proposals(4096, seed=3, base=g2)[0].to(dev)
        g2 = g2.to(dev)
        res["cfg2_boxes_iou3d_gpu_4096x200_ms"] = ev(lambda: ref.boxes_iou3d_gpu(pr, g2), 10, 2)
        b, s = synth.proposals(4096, 20, 20)
        b, s = b.to(dev), s.to(dev)
        res["cfg1_nms_gpu_4096_one_frame_ms"] = ev(lambda: ref.nms_gpu(b, s, 0.7), 5, 2)
        bx = torch.stack([synth.waymo_boxes(200, 100 + f) for f in range(8)])
        pts = torch.stack([synth.points(180000, bx[f], synth.WAYMO_RANGE, 0.05, seed=500 + f) for f in range(8)]).to(dev)
        bx = bx.to(dev)
        res["cfg2_points_in_boxes_gpu_8_frames_ms"] = ev(lambda: ref.points_in_boxes_gpu(pts, bx), 5, 2)
        smp, gt = synth.cvae_samples(200, 30, 0)
        smp, gt = smp.to(dev), gt.to(dev)
        res["cfg3_boxes_iou3d_gpu_6000x200_block_ms"] = ev(lambda: ref.boxes_iou3d_gpu(smp, gt), 10, 2)
        res["cfg3_note"] = "drop-in form of cfg3: 100 such blocks (6000 x 200, block diagonal taken) cover the 600000 pairs"
        return res
    except Exception as e:   # the baseline must never take the bench down
        return {"unavailable": f"{type(e).__name__}: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
