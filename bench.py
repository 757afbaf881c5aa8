#!/usr/bin/env python
"""Benchmark of the Mask3D hot path on B200 (contract: see the task brief; one JSON line on stdout from rank 0).

Workload (BASELINE.json configs[1]): full Mask3D inference — Res16UNet34C sparse-conv backbone + 12 masked-attention
query-refinement layers, 25 queries — on a synthetic single-floor 6-room Structured3D-shaped scene voxelised at 2 cm
(~0.89 M voxels), one scene per step per GPU, bf16 tensor-core mode (fp32 accumulate).  A step includes everything that
depends on the scene: coordinate hashing, 4 strided coordinate maps, 14 kernel maps, backbone, decoder.

  value   voxels/s with the scene's coordinates already resident in HBM (the next scene's coordinate phase is enqueued on a side
          stream behind the current scene's backbone: Mask3D.forward(x, prefetch=...); every scene's coordinate phase is inside
          the timed region)
  e2e     voxels/s through the public API from pinned HOST buffers, H2D of coordinates/features and D2H of the class
          logits + final mask logits inside the timed region
  train_fwd_bwd      (N = 1, beside the headline) Res16UNet34C forward + backward on the fused training path
  loss_side          (N = 1) SetCriterion forward + backward on the outputs of one inference step, and `train_step`: the whole
          training step Mask3D.forward_train -> SetCriterion -> backward -> AdamW on the same scene
  --impl reference   the CPU restatement of the reference's MinkowskiEngine + PyTorch path (oracle/, "port": ME 0.5.4 is
          not installable here) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_conv_umma launch from an `ncu --set full` capture (profiles/), next to the
# algorithmic bytes of that launch; filled in from the capture named in `note` (None until a capture of the current kernel exists)
NCU_TRAFFIC = {"bytes_per_launch": 229788416 + 139122176,
               "note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE k_conv_umma launch of the dominant shape (3^3 96->96 on the "
                       "889057-voxel level-0 map: 155.2 GFLOP, algorithmic minimum 437.9 MB incl. the 96 MB neighbour table, part of which "
                       "is L2-resident), ncu --set full, profiles/r02_ncu_conv96_summary.txt"}

METRIC = "Mask3D inference throughput (Res16UNet34C sparse-conv backbone + 12-layer masked-attention decoder)"
UNIT = "voxels/s"


def workload_config(voxel_m: float, n_scenes: int, scenes_per_step: int = 1, workload: str = "floor"):
    if workload == "building":
        what = ("BASELINE.json configs[3]: Mask3D full inference on synthetic multi-floor buildings (4 floors of 20x15x2.8 m, surfaces + "
                f"furniture), {voxel_m * 100:.1f} cm voxels, 25 queries, {scenes_per_step} building(s) per step per GPU, scenes sharded over GPUs")
    else:
        what = ("BASELINE.json configs[1]: Mask3D full inference, synthetic single-floor 6-room scenes "
                f"(10x8x2.8 m, surfaces + furniture), {voxel_m * 100:.1f} cm voxels, 25 queries, {scenes_per_step} scene(s) per step per GPU "
                "(one batch, as the reference batches scenes: conf/data/indoor.yaml:21)")
    return {"workload": what, "voxel_cm": voxel_m * 100, "scenes_per_step": scenes_per_step, "batches_cycled": n_scenes, "queries": 25,
            "decoder_layers": 12,
            "l2_policy": "inputs larger than L2: >= 0.9 M voxels x 96..384 channels per layer per scene (>170 MB activations per layer), "
                         "a different batch every step"}


# ----------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, as the profiling recipe prescribes)
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML (a thread polling every 50 ms: two cheap
    driver queries) — an `nvidia-smi -lms` subprocess re-initialises NVML and takes driver locks that showed up as occasional
    slow first timed regions; nvidia-smi remains the fallback when pynvml is unavailable."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.first, self.h, self._stop = index, [], None, 0, None, False

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            uuid = None
            try:
                import torch

                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            except Exception:
                pass
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
                except Exception:
                    h = None
            self.h = h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.h = None
        import subprocess

        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = [nv.nvmlClocksEventReasonHwSlowdown, nv.nvmlClocksEventReasonHwThermalSlowdown, nv.nvmlClocksEventReasonSwThermalSlowdown,
                nv.nvmlClocksEventReasonSwPowerCap]
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append([str(sm), str(self.mx), "0"] + ["Active" if (r & b) else "Not Active" for b in bits])
            except Exception:
                pass
            time.sleep(0.05)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Rows from here on belong to the timed region."""
        self.first = len(self.rows)

    def stop(self):
        if self.h is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        time.sleep(0.12)
        self._stop = True
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        rows = self.rows[self.first:] or self.rows  # inside the timed region (all rows if the region was shorter than a period)
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(self.NAMES, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.h is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's ME + PyTorch path)
# ----------------------------------------------------------------------------------------------------
def cpu_forward_factory(num_queries=25):
    import torch
    from oracle import mask3d_oracle, me_oracle
    from room_instance_segmentation_mask3d_b200 import mask3d_model, weights

    torch.set_num_threads(os.cpu_count() or 1)
    spec = [(k, tuple(v.shape)) for k, v in mask3d_model.Mask3D(num_queries=num_queries).state_dict().items()]
    sd = {k: torch.from_numpy(v) for k, v in weights.synth_state_dict(spec).items()}

    def run(coords, feats):
        x = me_oracle.SparseTensor(features=torch.from_numpy(feats), coordinates=torch.from_numpy(coords))
        with torch.no_grad():
            o = mask3d_oracle.mask3d_forward(sd, x, num_queries=num_queries)
        return o

    return run


def cpu_sample_scene(target_voxels: float):
    """The same floor-plan generator at the finest voxel size whose voxel count stays under `target_voxels`."""
    from room_instance_segmentation_mask3d_b200 import scenes

    n15 = 14633.0  # floor_plan(seed=1000) at 15 cm
    for v in (0.02, 0.03, 0.04, 0.05, 0.06, 0.08, 0.10, 0.15):
        if n15 * (0.15 / v) ** 2 * 1.1 <= target_voxels or v == 0.15:
            s = scenes.floor_plan(1000, voxel_m=v)
            return v, scenes.collate([s])


def time_cpu(budget_s: float, steps: int, warmup: int):
    run = cpu_forward_factory()
    from room_instance_segmentation_mask3d_b200 import scenes

    c, f = scenes.collate([scenes.floor_plan(1000, voxel_m=0.15)])
    t0 = time.perf_counter(); run(c, f); t_cal = time.perf_counter() - t0  # calibration: ~15 k voxels
    thr = len(c) / t_cal
    per_step = budget_s / max(steps + warmup, 1)
    v, (coords, feats) = cpu_sample_scene(max(thr * per_step * 1.5, 15000))  # larger scenes run at better voxels/s than the calibration
    for _ in range(warmup):
        run(coords, feats)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter(); run(coords, feats); ts.append(time.perf_counter() - t0)
    t = float(np.median(ts))
    return len(coords) / t, t, v, len(coords)


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    steps, warmup = args.steps, args.warmup
    val, t, v, n = time_cpu(150.0, steps, warmup)
    cores = os.cpu_count() or 1
    sample = f"floor_plan(seed=1000) at {v * 100:.0f} cm voxels = {n} voxels, one full Mask3D forward per step (median of {steps})"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(v, 1), note="the CPU arm runs the same generator at the finest voxel size that fits its time budget; "
                                                      "voxels/s normalises the scene size (the GPU arm's own 2 cm cpu_baseline is the slower figure)"),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "CPU restatement of the reference's MinkowskiEngine-CPU algorithm (hash maps, per-offset index_select -> "
                                     f"mm -> index_add_) + stock PyTorch decoder; torch threads = {torch.get_num_threads()}"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def main_ours(args):
    import torch
    import torch.distributed as dist

    from room_instance_segmentation_mask3d_b200 import lib, mask3d_model, me_api, ops, parallel, scenes, weights

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    ops.set_compute_mode(args.mode)
    voxel = args.voxel_cm / 100.0
    B = max(1, args.scenes_per_step)
    n_scenes = 3  # batches cycled
    # B + 2 distinct scenes per rank; batch i = scenes i .. i + B - 1 (cyclic): every step sees a different batch
    n_distinct = B + 2
    if args.workload == "building":
        gen = lambda i: scenes.building(2000 + rank * 64 + i, voxel_m=voxel)
    else:
        gen = lambda i: scenes.floor_plan(1000 + rank * 64 + i, voxel_m=voxel)
    distinct = [gen(i) for i in range(n_distinct)]
    host_scenes = [scenes.collate([distinct[(i + j) % n_distinct] for j in range(B)]) for i in range(n_scenes)]
    model = mask3d_model.Mask3D().eval()
    weights.load_synth(model, seed=0)
    model = model.to(dev)

    pinned = [(torch.from_numpy(c).pin_memory(), torch.from_numpy(f).pin_memory()) for c, f in host_scenes]
    resident = [(c.to(dev), f.to(dev)) for c, f in pinned]
    nvox = [c.shape[0] for c, _ in host_scenes]
    gb_per_mvox = 6.0  # caching-allocator segment per lane, per million voxels of a batch (maps + activations), measured ~4.5
    lane_gb = max(8.0, gb_per_mvox * max(nvox) / 1e6)
    lib.reserve_device_pool(4.0, dev)  # one splittable segment instead of cudaMalloc stalls while the allocator grows

    # Both loops hand the NEXT batch's inputs to the forward call (`prefetch=`): its coordinate phase (H2D in the e2e loop,
    # hashing, strided maps, kernel maps, tile plans: all the host round trips) runs on a side stream behind the current
    # batch's backbone instead of in front of an idle GPU.  Every batch still gets its own full coordinate phase inside the
    # timed region (step i prepares batch i+1; the first timed step consumes what the last warm-up step prepared).
    # Consecutive batches alternate between two streams ("lanes"): a batch's decoder is a chain of small latency-bound kernels, the
    # next batch's backbone a sequence of kernels that fill the GPU, and the two overlap.  Each lane has its own pinned result
    # buffers; the timed region ends when BOTH lanes have drained (see `timed`).
    n_lanes = 1 if args.lanes < 2 else 2
    lanes = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)]
    for st_ in lanes:  # the caching allocator keeps blocks per stream: one up-front segment for every stream a forward runs on
        lib.reserve_device_pool(lane_gb, dev, stream=st_)
    lib.reserve_device_pool(lane_gb / 2, dev, stream=model.side_stream(dev))
    compact = args.e2e_results == "compact"
    h_cls = [torch.empty((13, B, 25, 2), dtype=torch.float32).pin_memory() for _ in range(n_lanes)]
    # one pinned result buffer per lane, sized for the largest batch and sliced per batch (a buffer per (lane, batch) pinned 4 GB of
    # host memory per rank at 8 scenes per step)
    h_mask_buf = [torch.empty((max(nvox), 25), dtype=torch.float32).pin_memory() for _ in range(n_lanes)]
    h_bits_buf = [torch.empty(max(nvox), dtype=torch.int32).pin_memory() for _ in range(n_lanes)]
    h_mask = [[h_mask_buf[ln][:n] for n in nvox] for ln in range(n_lanes)]
    h_bits = [[h_bits_buf[ln][:n] for n in nvox] for ln in range(n_lanes)]
    h_score = [torch.empty((B, 25), dtype=torch.float32).pin_memory() for _ in range(n_lanes)]
    state = {"res": None, "e2e": None, "g_res": 0, "g_e2e": 0, "vox": 0, "compact": compact}

    def step_resident(_i=None):
        g = state["g_res"]
        if os.environ.get("MK3D_BENCH_STEP_DEBUG"):
            state.setdefault("t_calls", []).append(time.perf_counter())
        x = state["res"] if state["res"] is not None else model.prepare(*resident[g % n_scenes])
        with torch.cuda.stream(lanes[0 if state.get("one_lane") else g % n_lanes]):
            o = model(x, prefetch=resident[(g + 1) % n_scenes], aux_outputs=state.get("aux", True))
        state["res"] = o.pop("prefetched")
        state["g_res"] = g + 1
        state["vox"] += nvox[g % n_scenes]
        return o

    # e2e input path: the H2D copy of batch g + 1 (from pinned host memory, every step, inside the timed region) is issued on a copy
    # stream at the START of step g into a persistent device buffer of that batch, so that it runs under step g's backbone instead
    # of inside the coordinate phase the host waits for (`prepare` then sees device tensors).  A buffer is overwritten only after
    # the forward that last read it has finished (ev_done).
    copy_stream = torch.cuda.Stream(device=dev)
    dev_c = [torch.empty_like(c, device=dev) for c, _ in pinned]
    dev_f = [torch.empty_like(f, device=dev) for _, f in pinned]
    ev_done = [None] * n_scenes

    def step_e2e(_i=None):
        g = state["g_e2e"]
        if os.environ.get("MK3D_BENCH_ALLOC_DEBUG"):
            state.setdefault("t_calls", []).append(time.perf_counter())
        x = state["e2e"] if state["e2e"] is not None else model.prepare(*pinned[g % n_scenes])
        ln = g % n_lanes
        s_next = (g + 1) % n_scenes
        if ev_done[s_next] is not None:
            copy_stream.wait_event(ev_done[s_next])
        with torch.cuda.stream(copy_stream):
            dev_c[s_next].copy_(pinned[s_next][0], non_blocking=True)
            dev_f[s_next].copy_(pinned[s_next][1], non_blocking=True)
            ev_h2d = torch.cuda.Event()
            ev_h2d.record(copy_stream)
        model.side_stream(dev).wait_event(ev_h2d)
        with torch.cuda.stream(lanes[ln]):
            o = model(x, prefetch=(dev_c[s_next], dev_f[s_next]), compact=state["compact"])
            h_cls[ln].copy_(torch.stack(o["pred_class_all_decoders"]), non_blocking=True)
            if state["compact"]:  # the binary masks (one bit word per voxel) + per-query mask scores, what the post-processing keeps
                h_bits[ln][g % n_scenes].copy_(o["mask_bits"], non_blocking=True)
                h_score[ln].copy_(o["mask_score"], non_blocking=True)
            else:                 # the final fp32 mask logits [N, 25]
                h_mask[ln][g % n_scenes].copy_(o["pred_mask_all_decoders"][-1], non_blocking=True)
            ev_done[g % n_scenes] = torch.cuda.Event()
            ev_done[g % n_scenes].record(lanes[ln])  # everything that reads this step's input buffer is enqueued before this point
        state["e2e"] = o.pop("prefetched")
        state["g_e2e"] = g + 1
        state["vox"] += nvox[g % n_scenes]
        return o

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        for ln_ in lanes:  # the region ends when every lane has drained
            torch.cuda.current_stream().wait_stream(ln_)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # Python's cyclic collector is kept out of the timed regions (a generation-2 pass over the interpreter's ~10^5 objects is a
    # 50-100 ms pause, i.e. +5-10 ms per step on a 10-step measurement when it happens to fall inside one); the per-step
    # objects of this package are reference-count freed (no cycles: scripts/e2e_times.py with NOGC=1 shows flat memory)
    import gc

    gc.collect()
    gc.freeze()
    gc.disable()
    # the clock sampler is started BEFORE the warm-up: its start-up (NVML initialisation, first query) takes driver locks for tens of
    # milliseconds, which used to land inside the first timed region; the rows that fall inside the timed region are the ones reported
    sampler = ClockSampler(local)
    if rank == 0 and not args.no_clocks:
        sampler.start()
    # warm-up: at least every batch four times (CUDA-graph capture, weight images, and the caching allocator, which was still adding
    # multi-GB segments in the 7th-9th step of a leg: MK3D_BENCH_ALLOC_DEBUG=1) -- the line reports the --warmup
    # argument as `warmup` and the steps actually run as `warmup_run`
    n_warm = max(args.warmup, 4 * n_scenes)
    for i in range(n_warm):
        step_resident(i)
    sampler.mark()
    launches0 = lib.kernel_launches
    state["vox"] = 0
    mallocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    if os.environ.get("MK3D_BENCH_ALLOC_DEBUG"):
        torch.cuda.memory._record_memory_history(max_entries=200000)
    ms = timed(step_resident, args.steps)
    if os.environ.get("MK3D_BENCH_STEP_DEBUG"):
        tc = state["t_calls"][-args.steps:]
        print("[step debug] host ms between step calls in the timed region: " + " ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip(tc, tc[1:])), file=sys.stderr)
    if os.environ.get("MK3D_BENCH_ALLOC_DEBUG"):
        snap = torch.cuda.memory._snapshot()
        for tr in snap.get("device_traces", []):
            for ev in tr:
                if ev.get("action") in ("segment_alloc", "segment_free", "oom", "snapshot"):
                    print(f"[alloc debug] {ev.get('action')} size {ev.get('size', 0) / 2**20:.1f} MiB stream {ev.get('stream')}", file=sys.stderr)
        torch.cuda.memory._record_memory_history(enabled=None)
        st_ = torch.cuda.memory_stats(dev)
        print(f"[alloc debug] reserved {st_['reserved_bytes.all.current'] / 2**30:.1f} GiB, allocated peak {st_['allocated_bytes.all.peak'] / 2**30:.1f} GiB, "
              f"inactive split {st_['inactive_split_bytes.all.current'] / 2**30:.1f} GiB", file=sys.stderr)
    # cudaMalloc calls inside the timed region (0 when the caching allocator's pools were large enough: a cudaMalloc of a multi-GB
    # block stalls the enqueueing thread for milliseconds)
    mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs0
    launches = lib.kernel_launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    vox_per_step = state["vox"] / args.steps  # the voxels actually processed inside the timed region
    value = world * vox_per_step * args.steps / (ms * 1e-3)

    gc.collect()
    for i in range(max(args.warmup, 4 * n_scenes)):  # as many warm-up steps as the resident loop: the e2e path allocates its own result buffers
        step_e2e(i)
    state["vox"] = 0
    mallocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    if os.environ.get("MK3D_BENCH_ALLOC_DEBUG"):
        torch.cuda.memory._record_memory_history(max_entries=200000)
        state["t_calls"] = []
    ms_e2e = timed(step_e2e, args.steps)
    mallocs_e2e = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs0
    if os.environ.get("MK3D_BENCH_ALLOC_DEBUG"):
        for tr in torch.cuda.memory._snapshot().get("device_traces", []):
            for ev in tr:
                if ev.get("action") in ("segment_alloc", "segment_free", "oom"):
                    print(f"[alloc debug e2e] {ev.get('action')} size {ev.get('size', 0) / 2**20:.1f} MiB stream {ev.get('stream')}", file=sys.stderr)
        torch.cuda.memory._record_memory_history(enabled=None)
        tc = state.get("t_calls", [])
        print("[step debug e2e] host ms between e2e step calls: " + " ".join(f"{(b - a) * 1e3:.1f}" for a, b in zip(tc, tc[1:])), file=sys.stderr)
    e2e_value = world * state["vox"] / (ms_e2e * 1e-3)
    h2d = int(np.mean([c.numel() * 4 + f.numel() * 4 for c, f in pinned]))
    d2h_logits = int(h_cls[0].numel() * 4 + np.mean([m.numel() * 4 for m in h_mask[0]]))
    d2h_compact = int(h_cls[0].numel() * 4 + h_score[0].numel() * 4 + np.mean([m.numel() * 4 for m in h_bits[0]]))
    d2h = d2h_compact if compact else d2h_logits
    # the other result form, for comparison (fewer steps): full fp32 mask logits vs bit-packed masks + scores
    state["compact"] = not compact
    for i in range(n_scenes):
        step_e2e(i)
    state["vox"] = 0
    osteps = max(3, args.steps // 3)
    ms_other = timed(step_e2e, osteps)
    e2e_other = {"results": "logits" if compact else "compact", "value": world * state["vox"] / (ms_other * 1e-3), "unit": UNIT,
                 "d2h_bytes_per_step": d2h_logits if compact else d2h_compact, "ms_per_step": ms_other / osteps, "steps": osteps}
    state["compact"] = compact
    # resident steps WITHOUT the 12 intermediate full-resolution mask-logit products (Mask3D's aux_outputs: read by the criterion
    # only, the inference post-processing uses the final prediction) -- reported beside the headline, which computes everything
    # the reference's forward computes
    state["aux"] = False
    for i in range(max(4, n_scenes)):  # the allocator sees a new pattern of block sizes: let it settle outside the timed region
        step_resident(i)
    state["vox"] = 0
    ms_fo = timed(step_resident, args.steps)
    final_only = {"value": world * state["vox"] / (ms_fo * 1e-3), "unit": UNIT, "ms_per_step": ms_fo / args.steps, "steps": args.steps,
                  "note": "forward(aux_outputs=False): final prediction only, the intermediate decoder outputs' full-resolution mask logits are not materialised"}
    state["aux"] = True

    # per-launch CUDA events around the dominant kernel (k_conv_umma) over extra instrumented steps
    roofline = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        if args.mode == "tf32":  # tcgen05 kind::tf32 runs at half the bf16 rate; no measured tf32 figure exists in MEASURED_PEAKS.json
            peak *= 0.5
            peak_src = "half of " + peak_src + ": the dense tf32 rate of the tensor pipe"
        ops.USE_COMPOSITE = False  # time the conv launches on their own (the composite call also enqueues the norm kernels)
        state["one_lane"] = True  # per-launch events must not see the other lane's kernels interleaved
        for i in range(2):  # warm the instrumented configuration (separate conv / norm calls allocate differently) before timing it
            step_resident(i)
        torch.cuda.synchronize()
        conv_entry = "mk_conv_fwd_tf32" if args.mode == "tf32" else "mk_conv_fwd_bf16"
        ev = lib.instrument([conv_entry])
        isteps = min(args.steps, 3)
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(lanes[0]):
            t0.record()
        for i in range(isteps):
            step_resident(i)
        with torch.cuda.stream(lanes[0]):
            t1.record()
        state["one_lane"] = False
        torch.cuda.synchronize()
        lib.instrument(None)
        ops.USE_COMPOSITE = True
        recs = ev[conv_entry]
        step_ms = t0.elapsed_time(t1)

        def summarize(sel):
            ms_ = sum(s.elapsed_time(e) for s, e, t in recs if sel(t))
            fl_ = sum(t[0] for s, e, t in recs if sel(t))
            n_ = sum(1 for s, e, t in recs if sel(t))
            return ms_, fl_, n_

        tot_ms, tot_flops, n_launch = summarize(lambda t: True)
        by_shape = {}
        # t = (algorithmic FLOPs, K, cin, cout, rows, kind): kind "conv" = gather-GEMM over a spatial kernel map (3^3, 2^3, transposed
        # 2^3); "dense" = one input row per output row (1x1 convs, K / V projections, and the mask-logit product, whose "offsets"
        # are the scenes of the batch: K = B there)
        is_conv = lambda t: (t[5] == "conv") if len(t) > 5 else t[1] > 1
        for s, e, t in recs:
            k = ("K" if is_conv(t) else "dense_K") + f"{t[1]}_{t[2]}x{t[3]}"
            a = by_shape.setdefault(k, [0.0, 0.0, 0])
            a[0] += t[0]; a[1] += s.elapsed_time(e); a[2] += 1
        top = sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:24]
        # launch classes of the same kernel: the gather-GEMM proper (K = 27 / 8 offsets, tensor-bound by design) and the
        # K = 1 launches (1x1 convs, K/V projections, mask-logit product: one pass over [N, C], HBM-bound)
        g_ms, g_fl, g_n = summarize(is_conv)
        big_ms, big_fl, big_n = summarize(lambda t: is_conv(t) and t[4] >= 128 * 148 * 4)   # enough tiles to fill 148 SMs
        o_ms, o_fl, o_n = summarize(lambda t: not is_conv(t))
        hbm_peak = float(peaks.get("hbm_gbs", 6455.0))
        # algorithmic bytes of a K = 1 launch: read [N, Cin] bf16, write [N, Cout_stored] (bf16 activations; fp32 for the
        # 25-column mask logits and the fp32 `src` of the key projection) -- taken as 2 B in, 2 or 4 B out per element
        o_bytes = sum(t[4] * (2 * t[2] + (4 if (t[3] <= 32 or t[2] != t[3]) else 2) * min(t[3], 128)) for s, e, t in recs if not is_conv(t))
        tf = lambda fl, ms_: fl / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0.0
        achieved = tf(g_fl, g_ms)
        roofline = {"kernel": "k_conv_umma (tcgen05 gather-GEMM sparse conv): the launches over spatial kernel maps (3^3, 2^3 and transposed "
                              "2^3 convolutions); the dense launches of the same kernel (1x1 convs, K / V projections, mask-logit product) "
                              "are HBM-bound and reported under classes.K=1_hbm_bound",
                    "bound": "tensor", "achieved": achieved,
                    "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": NCU_TRAFFIC.get("bytes_per_launch"),
                    "traffic_note": NCU_TRAFFIC.get("note"), "peak_source": peak_src,
                    "launches_per_step": g_n / isteps, "flops_per_launch_avg": g_fl / max(g_n, 1),
                    "ms_per_launch_avg": g_ms / max(g_n, 1), "share_of_step": g_ms / step_ms,
                    "all_launches": {"launches_per_step": n_launch / isteps, "ms_per_step": tot_ms / isteps,
                                     "achieved_tflops": tf(tot_flops, tot_ms), "frac_of_tensor_peak": tf(tot_flops, tot_ms) / peak,
                                     "share_of_step": tot_ms / step_ms},
                    "classes": {
                        "gather_gemm_K>1": {"launches_per_step": g_n / isteps, "ms_per_step": g_ms / isteps, "achieved_tflops": tf(g_fl, g_ms),
                                            "frac_of_tensor_peak": tf(g_fl, g_ms) / peak},
                        "gather_gemm_K>1_levels_with_>=4_tiles_per_SM": {"launches_per_step": big_n / isteps, "ms_per_step": big_ms / isteps,
                                                                        "achieved_tflops": tf(big_fl, big_ms),
                                                                        "frac_of_tensor_peak": tf(big_fl, big_ms) / peak},
                        "K=1_hbm_bound": {"launches_per_step": o_n / isteps, "ms_per_step": o_ms / isteps, "achieved_tflops": tf(o_fl, o_ms),
                                          "achieved_gbs": o_bytes / (o_ms * 1e-3) / 1e9 if o_ms > 0 else 0.0, "hbm_peak_gbs": hbm_peak,
                                          "frac_of_hbm_peak": (o_bytes / (o_ms * 1e-3) / 1e9) / hbm_peak if o_ms > 0 else 0.0}},
                    "by_shape_tflops": {k: round(v[0] / (v[1] * 1e-3) / 1e12, 1) for k, v in top},
                    "by_shape_ms_per_step": {k: round(v[1] / isteps, 3) for k, v in top}}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, t, v, n = time_cpu(25.0, 1, 0)
        cpu_baseline = {"value": val, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"floor_plan(seed=1000) at {v * 100:.0f} cm voxels = {n} voxels, one full Mask3D forward ({t:.1f} s)"}

    # first clause of BASELINE.json's metric: Res16UNet34C forward + backward voxels/s (fused training path, bf16 tensor cores),
    # on the same scene generator; reported beside the headline, not part of the timed regions above
    Bt = max(1, min(B, args.train_scenes))
    c_t, f_t = (torch.from_numpy(x).to(dev) for x in scenes.collate([distinct[j] for j in range(Bt)]))
    n_t = int(c_t.shape[0])
    sizes_t = [len(distinct[j]) for j in range(Bt)]
    train = None
    if rank == 0 and world == 1 and not args.no_train:
        try:
            bb = mask3d_model.Res16UNet34C()
            weights.load_synth(bb, seed=0)
            bb = bb.to(dev).train()

            def tstep():
                for prm in bb.parameters():
                    prm.grad = None
                out, fm = bb.forward_train(me_api.SparseTensor(features=f_t, coordinates=c_t))
                (out.F.float().square().mean() + sum(t.F.float().square().mean() for t in fm[:-1])).backward()

            for _ in range(6):
                tstep()
            tsteps = 5
            t_ms = timed(lambda i: tstep(), tsteps) / tsteps
            train = {"what": "Res16UNet34C forward + backward, fused training path (bf16 activations, tcgen05 fwd/dgrad/wgrad), maps and "
                             f"plans rebuilt every step, {Bt} scene(s) per step", "voxels": n_t, "ms_fwd_bwd": t_ms,
                     "voxels_per_s": n_t / (t_ms * 1e-3)}
            del bb
        except Exception as exc:  # the headline line must survive a failure here
            train = {"error": repr(exc)[:200]}

    # BASELINE.json configs[2]: the training step, data-parallel over the N GPUs (one process per GPU, every rank its own scenes):
    # Mask3D.forward_train -> SetCriterion (incl. the reference's all_reduce(num_masks), criterion.py:244-246) -> backward with the
    # bucketed NCCL gradient all-reduce OVERLAPPED with it (parallel.GradientBuckets) -> AdamW (lr 1e-4).  Runs at every N.
    loss_side = None
    train_step = None
    if not args.no_train:
        try:
            from types import SimpleNamespace

            from room_instance_segmentation_mask3d_b200 import criterion as CR

            T = 24  # instances per scene = runs of voxel rows
            tgt = []
            for nb in sizes_t:
                seg = torch.div(torch.arange(nb, device=dev) * T, nb, rounding_mode="floor")
                tgt.append(SimpleNamespace(instances_labels=torch.zeros(T, dtype=torch.long, device=dev),
                                           instances_masks=torch.stack([seg == t for t in range(T)])))
            wts = {"loss_ce": 2.0, "loss_mask": 5.0, "loss_dice": 2.0}
            wd = dict(wts)
            for d in range(12):
                wd.update({k + f"_mask_module_{d}": v for k, v in wts.items()})
            crit = CR.SetCriterion(1, CR.HungarianMatcher(2.0, 5.0, 2.0, -1), wd, 0.1, ["labels", "masks"], -1, 3.0, 0.75, -1).to(dev)

            if rank == 0 and world == 1:
                # loss side alone on the outputs of one inference step of the training batch
                with torch.no_grad():
                    o = model(model.prepare(c_t, f_t))
                pm = [t.detach().clone().requires_grad_(True) for t in o["pred_mask_all_decoders"]]
                pc = [t.detach().clone().requires_grad_(True) for t in o["pred_class_all_decoders"]]
                mo = dict(pred_class=pc[-1], pred_class_all_decoders=pc, pred_mask=pm[-1], pred_mask_all_decoders=pm)

                def lstep():
                    for t in pm + pc:
                        t.grad = None
                    ls = crit(mo, tgt)
                    sum(ls[k] * wd[k] for k in ls).backward()

                for _ in range(3):
                    lstep()
                l_ms = timed(lambda i: lstep(), 5) / 5
                loss_side = {"what": "SetCriterion forward + backward (matcher costs, Hungarian on the host, mask/dice/class losses, gradients of "
                                     "all decoder outputs' mask logits) on the outputs of one inference step", "voxels": n_t,
                             "decoder_outputs": len(pm), "instances": T * Bt, "ms_fwd_bwd": l_ms}
                del pm, pc, mo, o

            model.train()
            params = [p_ for n_, p_ in model.named_parameters() if not n_.startswith("backbone.final")]
            gbk = parallel.GradientBuckets(params)
            opt = torch.optim.AdamW(params, lr=1e-4)
            ev_b = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            exposed = []

            def wstep(record=False):
                gbk.zero()
                out_t = model.forward_train(me_api.SparseTensor(features=f_t, coordinates=c_t))
                ls = crit(out_t, tgt)
                sum(ls[k] * wd[k] for k in ls).backward()
                ev_b[0].record()
                gbk.finish()  # the compute stream waits here for whatever part of the all-reduce backward did not hide
                ev_b[1].record()
                opt.step()
                if record:
                    torch.cuda.synchronize()
                    exposed.append(ev_b[0].elapsed_time(ev_b[1]))

            for _ in range(8):  # the caching allocator keeps growing (cudaMalloc) for the first handful of training steps
                wstep()
            wsteps = 5
            w_ms = timed(lambda i: wstep(), wsteps) / wsteps
            for _ in range(3):
                wstep(record=True)
            n_par = sum(p_.numel() for p_ in params)
            train_step = {"what": "BASELINE.json configs[2]: Mask3D.forward_train + SetCriterion + backward + bucketed gradient all-reduce "
                                  "(NCCL, overlapped with backward) + AdamW; maps and plans rebuilt every step; every rank its own "
                                  f"{Bt} scene(s) of the workload", "n_gpus": world, "scenes_per_step_per_gpu": Bt, "voxels_per_gpu": n_t,
                          "ms_step": w_ms, "voxels_per_s": world * n_t / (w_ms * 1e-3), "scenes_per_s": world * Bt / (w_ms * 1e-3),
                          "allreduce": {"parameters": n_par, "bytes": 4 * n_par, "buckets": len(gbk.buckets),
                                        "collectives_per_step": gbk.collectives // max(1, 8 + wsteps + 3) if world > 1 else 0,
                                        "exposed_ms": float(np.median(exposed)) if exposed else None,
                                        "note": "exposed = time the compute stream waits in GradientBuckets.finish() after backward "
                                                "(CUDA events); at N = 1 no collective runs"}}
            gbk.remove()
            model.eval()
            for p_ in model.parameters():
                p_.grad = None
            del opt, gbk
        except Exception as exc:
            train_step = {"error": repr(exc)[:300]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "warmup_run": n_warm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[args.mode], "data": "synthetic",
                "config": workload_config(voxel, n_scenes, B, args.workload),
                "scenes_per_sec": world * B * args.steps / (ms * 1e-3), "voxels_per_step_per_gpu": vox_per_step,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps,
                        "results": ("compact: class logits of all 13 decoder outputs + binary masks of the final output as one bit word per "
                                    "voxel + per-query mask scores (what trainer/trainer.py:373-402 derives from the logits), built on the "
                                    "device") if compact else "class logits of all 13 decoder outputs + final fp32 mask logits [N, 25]",
                        "other_result_form": e2e_other},
                "gpu_launches": launches, "cuda_mallocs_in_timed_region": mallocs, "cuda_mallocs_in_e2e_region": mallocs_e2e, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "train_fwd_bwd": train,
                "loss_side": loss_side, "train_step": train_step, "final_prediction_only": final_only}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class _StdoutToStderr:
    """Everything other than the final JSON line goes to stderr, also what native libraries write to file descriptor 1 (NCCL
    prints its version banner there at the first collective): the contract is ONE JSON line on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def restore(self):
        sys.stdout.flush()
        os.dup2(self.saved, 1)

    def __exit__(self, *exc):
        self.restore()
        os.close(self.saved)
        return False


_real_print = print


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "tf32", "fp32"],
                    help="bf16: the benchmarked tensor-core mode; tf32: the reference's own precision on the same tcgen05 kernel; fp32: CUDA-core checker")
    ap.add_argument("--voxel-cm", type=float, default=2.0)
    ap.add_argument("--scenes-per-step", type=int, default=8, help="scenes batched per step per GPU (the reference batches 16 small scenes, "
                    "conf/data/indoor.yaml:21; 8 x 0.9 M voxels fill the coarse U-Net levels: 208 tiles on level 4)")
    ap.add_argument("--train-scenes", type=int, default=2, help="scenes per training step per GPU")
    ap.add_argument("--workload", default="floor", choices=["floor", "building"], help="floor: configs[1] single-floor scenes; building: configs[3] multi-floor buildings")
    ap.add_argument("--e2e-results", default="compact", choices=["compact", "logits"],
                    help="what the end-to-end loop copies back per step: compact = bit-packed binary masks + mask scores + class logits "
                         "(default), logits = the final fp32 mask logits [N, 25] + class logits; the other form is measured beside it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=2, help="streams consecutive scenes alternate between (1 = a single stream)")
    ap.add_argument("--no-train", action="store_true", help="skip the backbone forward+backward measurement")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi during the timed region (debugging)")
    a = ap.parse_args()
    with _StdoutToStderr() as guard:
        lines = []
        print = lambda *args, **kw: lines.append(" ".join(str(x) for x in args))  # noqa: A001  (collect the JSON line)
        globals()["print"] = print
        if a.impl == "reference":
            main_reference(a)
        else:
            main_ours(a)
        guard.restore()
        for ln in lines:
            _real_print(ln, flush=True)
        os.dup2(2, 1)
