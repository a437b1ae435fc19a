#!/usr/bin/env python
"""bench.py -- KLT tracked points/s and frames/s of the Tracker2D hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3]

Workload (BASELINE.json configs[1], "c2"): PETS2009-shaped 4-view rig, 768x576 gray, 40 targets per view,
100 Shi-Tomasi seeds per target, 4-level pyramid, 21x21 window, forward + backward LK with the FB gate.
One STEP = one frame time of the rig on one GPU:
    psnw_push_frames   4 new camera frames -> Gaussian pyramid + Scharr derivatives (once per frame)
    psnw_klt_track     16 000 seeds slot 1 -> slot 0, forward + backward + FB check (32 000 LK solves), one launch
    psnw_box_update    160 targets, LocalSearchKLT box motion from the FB-validated pairs
    psnw_klt_detect    new seeds for the next step inside the 160 detection boxes of the new frame
Every rank (one process per GPU, torchrun) runs its own rig: cameras shard across GPUs with no collective on
the data path -> "scaling": "weak".  `value` = tracked points/s with frames, seeds and boxes
already resident in HBM; `e2e` = the same step through the C ABI with pinned HOST buffers, H2D/D2H inside the
timed region.  L2 is flushed (256 MiB memset) between timed steps; time = sum of per-step CUDA-event intervals
on the context stream, max over ranks.

--impl reference: the reference's CPU path (OpenCV calcOpticalFlowPyrLK via cv2 with all host threads + the C
restatement of LocalSearchKLT / selection), rank 0 only.  This leg and `cpu_baseline` are the only places that
touch oracle/.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, cams per GPU, targets per cam, seeds per target, levels)
    "c1": (768, 576, 1, 20, 100, 3),
    "c2": (768, 576, 4, 40, 100, 4),
    "c3": (1920, 1080, 1, 200, 256, 4),
    "c4": (3840, 2160, 2, 0, 0, 5),     # 16-camera rig over 8 GPUs = 2 cameras per GPU; dense grid KLT, stride 4 px, track + FB only
}
CYCLE = 6                       # distinct time steps rendered per camera (played ping-pong)
QUALITY, MIN_DIST = 0.01, 3.0   # psnw_klt_detect / goodFeaturesToTrack parameters
METRIC = "KLT tracked points/sec (forward+backward validated), Tracker2D hot path"


def pingpong(i):
    p = i % (2 * CYCLE - 2)
    return p if p < CYCLE else 2 * CYCLE - 2 - p


def render_rig(wl, rank):
    from mcmtt_opticalflow_b200 import synth
    w, h, ncam, ntgt, per, _ = WORKLOADS[wl]
    cams = [synth.Camera(w, h, ntgt, cam=rank * ncam + c) for c in range(ncam)]
    frames = [[c.frame(t) for t in range(CYCLE)] for c in cams]
    boxes = [[c.boxes(t) for t in range(CYCLE)] for c in cams]
    return frames, boxes


# ----------------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:          # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
class GpuRig:
    """One GPU's rig driven through the C ABI; all per-step arrays preallocated (device and pinned host)."""

    def __init__(self, wl, rank, device, pipeline=True):
        import torch
        self.pipeline = pipeline
        from mcmtt_opticalflow_b200 import klt
        self.torch, self.klt = torch, klt
        self.w, self.h, self.ncam, self.ntgt, self.per, self.levels = WORKLOADS[wl]
        self.grid = self.ntgt == 0
        if self.grid:
            from mcmtt_opticalflow_b200 import synth as _synth
            self.grid_pts = _synth.grid_seeds(self.w, self.h, 4, 2)
        self.npts_cam = len(self.grid_pts) if self.grid else self.ntgt * self.per
        self.npts = self.npts_cam * self.ncam
        self.ntargets = self.ntgt * self.ncam
        self.frames, self.boxes = render_rig(wl, rank)
        self.dev = torch.device("cuda", device)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.ctx = klt.KLTContext(self.w, self.h, n_cameras=self.ncam, levels=self.levels, win=(21, 21), fb_thresh=1.0,
                                  max_points=max(self.npts, 1024), max_targets=max(self.ntargets, 64), device=device,
                                  stream=self.stream.cuda_stream)
        t = torch
        with t.cuda.stream(self.stream):
            # frames: device copies and pinned host copies
            blk = np.stack([np.stack([self.frames[c][k] for c in range(self.ncam)]) for k in range(CYCLE)])   # [cycle][cam][h][w]
            self.h_block = t.from_numpy(blk).pin_memory()
            self.d_block = self.h_block.to(self.dev)
            self.h_frames = [[self.h_block[k, c] for k in range(CYCLE)] for c in range(self.ncam)]
            self.d_frames = [[self.d_block[k, c] for k in range(CYCLE)] for c in range(self.ncam)]
            # device state
            self.d_seeds = [t.zeros((self.npts, 2), dtype=t.float32, device=self.dev) for _ in range(2)]
            # results of a step live in ONE device blob (points | boxes | counts | status | mask) so that the end-to-end
            # arm reads them back with a single copy
            n8 = (self.npts + 7) // 8 * 8
            sizes = [self.npts * 8, self.ntargets * 32, (self.ntargets * 4 + 7) // 8 * 8, n8, n8]
            offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            self.d_blob = t.zeros(int(offs[-1]), dtype=t.uint8, device=self.dev)
            self.h_blob = t.zeros(int(offs[-1]), dtype=t.uint8).pin_memory()
            view = lambda o, nbytes, dt, shape: self.d_blob[int(o):int(o) + nbytes].view(dt).view(shape)
            self.d_out = view(offs[0], self.npts * 8, t.float32, (self.npts, 2))
            self.d_boxes_out = view(offs[1], self.ntargets * 32, t.float64, (self.ntargets, 4))
            self.d_counts = view(offs[2], self.ntargets * 4, t.int32, (self.ntargets,))
            self.d_status = view(offs[3], self.npts, t.uint8, (self.npts,))
            self.d_mask = view(offs[4], self.npts, t.uint8, (self.npts,))
            self.blob_bytes = int(offs[-1])
            self.d_err = t.zeros(self.npts, dtype=t.float32, device=self.dev)
            self.d_fb = t.zeros(self.npts, dtype=t.float32, device=self.dev)
            self.d_counts_det = t.zeros(self.ntargets, dtype=t.int32, device=self.dev)
            self.d_boxes_in = t.zeros((self.ntargets, 4), dtype=t.float64, device=self.dev)
            off = np.arange(self.ntargets + 1, dtype=np.int32) * max(self.per, 1)
            self.d_off = t.from_numpy(off).to(self.dev)
            self.h_off = t.from_numpy(off).pin_memory()
            # pinned host mirrors for the end-to-end arm
            pin = lambda *s, dt=t.float32: t.zeros(s, dtype=dt).pin_memory()
            self.h_seeds = [pin(self.npts, 2) for _ in range(2)]
            self.h_out, self.h_status, self.h_err, self.h_fb = pin(self.npts, 2), pin(self.npts, dt=t.uint8), pin(self.npts), pin(self.npts)
            self.h_counts_det = pin(self.ntargets, dt=t.int32)
            self.h_boxes_in = pin(self.ntargets, 4, dt=t.float64)
            self.h_boxes_out = pin(self.ntargets, 4, dt=t.float64)
            self.h_mask, self.h_counts = pin(self.npts, dt=t.uint8), pin(self.ntargets, dt=t.int32)
            self.flush = t.empty(256 << 20, dtype=t.uint8, device=self.dev)
        self.stream.synchronize()
        self.cams = np.arange(self.ncam, dtype=np.int32)
        self.which = 0
        self.step_idx = 0
        self._job_cache = {}
        self._det_cache = {}

    # -- helpers ---------------------------------------------------------------------------------------
    def _jobs(self, seeds_ptr, out_ptr, st_ptr, err_ptr, fb_ptr, slot_from=1, slot_to=0):
        key = (seeds_ptr, out_ptr, slot_from)
        if key in self._job_cache:
            return self._job_cache[key]
        klt = self.klt
        arr = (klt.TrackJob * self.ncam)()
        self._job_cache[key] = arr
        for c in range(self.ncam):
            o = c * self.npts_cam
            a = arr[c]
            a.cam, a.slot_from, a.slot_to, a.n_pts = c, slot_from, slot_to, self.npts_cam
            a.pts_xy, a.target_of_pt, a.n_targets, a.win_wh = seeds_ptr + o * 8, None, 1, None
            a.out_xy, a.status, a.err, a.fb_err = out_ptr + o * 8, st_ptr + o, err_ptr + o * 4, fb_ptr + o * 4
        return arr

    def _push(self, t_idx, host, overlap=False):
        src = self.h_frames if host else self.d_frames
        ptrs = (C.c_void_p * self.ncam)(*[src[c][t_idx].data_ptr() for c in range(self.ncam)])
        mem = (self.klt.MEM_PINNED if host else self.klt.MEM_DEVICE) | (self.klt.MEM_OVERLAP if overlap else 0)
        if overlap and not self.grid:
            mem |= self.klt.PUSH_EIG         # the next step detects on this frame: its Shi-Tomasi map is built with the pyramid
        self.ctx._ck(self.ctx.lib.psnw_push_frames(self.ctx._h, self.ncam, self.cams.ctypes.data, ptrs, self.w, t_idx, mem))

    def _detect(self, t_idx, seeds, counts, mem, slot=0):
        """Shi-Tomasi seeds inside the detection boxes of frame t_idx (ring slot `slot`), all cameras in one batched call."""
        key = (t_idx, seeds.data_ptr(), slot)
        jobs = self._det_cache.get(key)
        if jobs is None:
            jobs = (self.klt.DetectJob * self.ncam)()
            for c in range(self.ncam):
                j = jobs[c]
                j.cam, j.slot, j.boxes, j.n_targets = c, slot, self.boxes[c][t_idx].ctypes.data, self.ntgt
                j.pts_xy = seeds.data_ptr() + c * self.npts_cam * 8
                j.counts = counts.data_ptr() + c * self.ntgt * 4
            self._det_cache[key] = jobs
        self.ctx._ck(self.ctx.lib.psnw_klt_detect_batch(self.ctx._h, jobs, self.ncam, self.per, QUALITY, MIN_DIST, mem))

    def prime(self, host):
        """First frame + first seeds so that every timed step has a previous frame to track from."""
        self._push(pingpong(0), host)
        if self.grid:
            g = self.torch.from_numpy(np.tile(self.grid_pts, (self.ncam, 1))).to(self.dev)
            self.d_seeds[0].copy_(g); self.d_seeds[1].copy_(g)
        else:
            self._detect(pingpong(0), self.d_seeds[self.which], self.d_counts_det, self.klt.MEM_DEVICE)
            boxes0 = np.concatenate([self.boxes[c][pingpong(0)] for c in range(self.ncam)])
            self.d_boxes_in.copy_(self.torch.from_numpy(boxes0).to(self.dev))
        if self.pipeline:                    # one frame in flight: the frame a step tracks into was pushed by the step before
            self._push(pingpong(1), host)
        self.ctx.sync()
        self.step_idx = 1

    def step(self, host):
        """One frame time of the rig.
        host=False: frames already in HBM.  host=True (end to end): the step's inputs -- the 4 new frames, in pinned
        host memory, and the detection boxes, in pageable host memory -- are copied in, and the step's results
        (tracked points, status, updated boxes, inlier counts and mask) are copied out to pinned host memory.
        Seeds are internal state of the path (detect output -> next step's track input) and stay on the device.
        Detection of the next step's seeds runs on the auxiliary stream and overlaps with tracking."""
        # (grid workloads keep their seeds: self.which is not flipped)
        klt, ctx = self.klt, self.ctx
        t_idx = pingpong(self.step_idx)
        cur, nxt = self.which, self.which ^ 1
        seeds, seeds_next = self.d_seeds[cur], self.d_seeds[nxt]
        jobs = self._jobs(seeds.data_ptr(), self.d_out.data_ptr(), self.d_status.data_ptr(), self.d_err.data_ptr(), self.d_fb.data_ptr())
        if self.pipeline:
            # One frame in flight: frame t_idx is already on the device -- the previous step pushed it while
            # it was tracking.  This step tracks into it and ingests the NEXT frame on the upload stream
            # (PSNW_MEM_OVERLAP), so the H2D copy and the pyramid build run under the LK launch.  Every step still
            # uploads one rig frame, tracks one and returns one result.
            if not self.grid:
                self._detect(t_idx, seeds_next, self.d_counts_det, klt.MEM_DEVICE | klt.MEM_OVERLAP)
            ctx._ck(ctx.lib.psnw_klt_track_batch(ctx._h, jobs, self.ncam, klt.MEM_DEVICE))
            self._push(pingpong(self.step_idx + 1), host, overlap=True)
        else:
            self._push(t_idx, host)
            if not self.grid:
                self._detect(t_idx, seeds_next, self.d_counts_det, klt.MEM_DEVICE | klt.MEM_OVERLAP)
            ctx._ck(ctx.lib.psnw_klt_track_batch(ctx._h, jobs, self.ncam, klt.MEM_DEVICE))
        if not self.grid:                    # C4 (dense grid) is forward + backward LK + FB gate only
            ctx.box_update_raw(self.d_boxes_in.data_ptr(), self.ntargets, seeds.data_ptr(), self.d_out.data_ptr(), self.d_status.data_ptr(),
                               self.d_off.data_ptr(), self.d_boxes_out.data_ptr(), self.d_mask.data_ptr(), self.d_counts.data_ptr(), klt.MEM_DEVICE)
            self.which = nxt
        ctx.join()
        if host:
            self.h_blob.copy_(self.d_blob, non_blocking=True)
        self.step_idx += 1

    def e2e_bytes(self):
        n, nt = self.npts, self.ntargets
        if self.grid:
            return self.ncam * self.w * self.h, self.blob_bytes
        h2d = self.ncam * self.w * self.h + nt * 32          # frames + detection boxes (PSN_Rect)
        return h2d, self.blob_bytes                          # points, boxes, inlier counts, status, inlier mask (one blob)

    def timed(self, steps, warmup, host, sampler=None):
        """-> (total seconds over `steps`, launches) with L2 flushed between steps; events on the context stream."""
        t = self.torch
        self.prime(host)
        with t.cuda.stream(self.stream):
            for _ in range(warmup):
                self.step(host)
                if host:
                    self.ctx.sync()
            self.stream.synchronize()
            ev = [(t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)) for _ in range(steps)]
            l0 = self.ctx.launch_count
            if sampler:
                sampler.start()
            for i in range(steps):
                self.flush.zero_()
                ev[i][0].record(self.stream)
                self.step(host)
                ev[i][1].record(self.stream)
                if host:
                    self.ctx.sync()          # the step's result (boxes, points) is read on the host every step
            self.stream.synchronize()
        total_ms = sum(a.elapsed_time(b) for a, b in ev)
        return total_ms * 1e-3, self.ctx.launch_count - l0

    def time_pyramid(self, reps=30):
        """HBM-bound kernel group alone: pyramid + Scharr build of the rig's `ncam` resident frames."""
        t = self.torch
        with t.cuda.stream(self.stream):
            for _ in range(3):
                self.ctx.rebuild_pyramids(self.cams)
            ev = [(t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)) for _ in range(reps)]
            for i in range(reps):
                self.flush.zero_()
                ev[i][0].record(self.stream)
                self.ctx.rebuild_pyramids(self.cams)
                ev[i][1].record(self.stream)
            self.stream.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        return float(np.mean(ms)) * 1e-3, float(ms[0]) * 1e-3

    def time_track(self, reps=20):
        """LK kernel alone (forward+backward fused launch): the newest frame's seeds tracked into the previous frame."""
        t = self.torch
        self.step(False)
        if self.grid:
            self.step(False)
        seeds = self.d_seeds[self.which]
        # pipelined: slot 0 already holds the next frame, the seeds belong to slot 1 -- exactly the next step's launch;
        # otherwise the newest frame's seeds are tracked back into the previous frame
        sf, st_ = (1, 0) if self.pipeline else (0, 1)
        jobs = self._jobs(seeds.data_ptr(), self.d_out.data_ptr(), self.d_status.data_ptr(), self.d_err.data_ptr(), self.d_fb.data_ptr(), sf, st_)
        with t.cuda.stream(self.stream):
            ev = [(t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)) for _ in range(reps)]
            for i in range(reps):
                self.flush.zero_()
                ev[i][0].record(self.stream)
                self.ctx._ck(self.ctx.lib.psnw_klt_track_batch(self.ctx._h, jobs, self.ncam, self.klt.MEM_DEVICE))
                ev[i][1].record(self.stream)
            self.stream.synchronize()
        return float(np.mean([a.elapsed_time(b) for a, b in ev])) * 1e-3, float(self.d_status.float().mean().item())


# ----------------------------------------------------------------------------------------------------------
# CPU arm (reference path): cv2 LK with all host threads + C restatement of LocalSearchKLT / selection
# ----------------------------------------------------------------------------------------------------------
class CpuRig:
    def __init__(self, wl, rank=0):
        import cv2
        from oracle import cv2_ref, psnw_oracle
        self.cv2, self.ref, self.po = cv2, cv2_ref, psnw_oracle
        psnw_oracle.build()
        self.w, self.h, self.ncam, self.ntgt, self.per, self.levels = WORKLOADS[wl]
        self.frames, self.boxes = render_rig(wl, rank)
        self.threads = os.cpu_count() or 1
        cv2.setNumThreads(self.threads)
        self.seeds = [None] * self.ncam
        self.prev_boxes = [None] * self.ncam
        self.t = [0] * self.ncam
        self.grid = self.ntgt == 0
        if self.grid:
            from mcmtt_opticalflow_b200 import synth as _synth
            self.seeds = [_synth.grid_seeds(self.w, self.h, 4, 2) for _ in range(self.ncam)]
            self.t = [1] * self.ncam
            return
        for c in range(self.ncam):
            self.seeds[c], _ = self._detect(c, 0)
            self.prev_boxes[c] = self.boxes[c][0]
            self.t[c] = 1

    def _detect(self, c, t_idx):
        img = self.frames[c][t_idx]
        eig = self.cv2.cornerMinEigenVal(img, 3, ksize=3)          # once per frame, not once per box
        pts = np.full((self.ntgt, self.per, 2), -1e6, np.float32)
        counts = np.zeros(self.ntgt, np.int32)
        for k, b in enumerate(self.boxes[c][t_idx]):
            p = self.po.detect(img, b, self.per, QUALITY, MIN_DIST, eig=eig)
            pts[k, :len(p)] = p
            counts[k] = len(p)
        return pts.reshape(-1, 2), counts

    def step_cam(self, c):
        """One camera-frame of the step (bounded sample unit).  Returns the number of seeds carried."""
        t_prev, t_cur = pingpong(self.t[c] - 1), pingpong(self.t[c])
        a, b = self.frames[c][t_prev], self.frames[c][t_cur]
        seeds = self.seeds[c]
        out, st, err, fb = self.ref.track_fb(a, b, seeds, (21, 21), self.levels - 1, fb_thresh=1.0)
        if self.grid:
            self.t[c] += 1
            return len(seeds)
        for k in range(self.ntgt):
            s = slice(k * self.per, (k + 1) * self.per)
            self.po.local_search_klt(self.prev_boxes[c][k], seeds[s], out[s], st[s])
        self.seeds[c], _ = self._detect(c, t_cur)
        self.prev_boxes[c] = self.boxes[c][t_cur]
        self.t[c] += 1
        return len(seeds)


def cpu_measure(wl, steps, warmup, budget_s=20.0):
    """Times the CPU path on bounded samples: one camera-frame per sample.  -> dict for the JSON line."""
    rig = CpuRig(wl)
    for i in range(max(warmup, 1)):
        rig.step_cam(i % rig.ncam)
        if rig.grid:
            break
    n, pts, t0 = 0, 0, time.perf_counter()
    while n < steps and (time.perf_counter() - t0) < budget_s:
        pts += rig.step_cam(n % rig.ncam)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": pts / dt, "unit": "points/s", "cores": rig.threads, "kind": "port",
            "sample": "%d camera-frames of %dx%d (%d seeds each; c4 = stride-4 grid with LK + FB only; fwd+bwd LK 21x21/%d levels via cv2 %s with %d threads, "
                      "FB gate, C port of LocalSearchKLT per target, cornerMinEigenVal once per frame + C port of the GFTT "
                      "selection per box) in %.2f s" % (n, rig.w, rig.h, len(rig.seeds[0]), rig.levels, rig.cv2.__version__,
                                                       rig.threads, dt),
            "seeds_per_camera_frame": len(rig.seeds[0]),
            "frames_per_s": n / dt, "ms_per_camera_frame": 1e3 * dt / max(n, 1), "camera_frames": n}


def pyramid_sweep(rig, peak):
    """Same K1 launch group with more camera-frames per launch (BASELINE config 5: frames in flight): the kernel is
    latency-bound at 4 small frames and approaches the HBM roofline as the batch grows."""
    import torch
    from mcmtt_opticalflow_b200 import klt
    out = []
    for (w, h, levels, frames) in [(768, 576, 4, 16), (768, 576, 4, 64), (1920, 1080, 4, 32), (3840, 2160, 5, 16)]:
        ctx = klt.KLTContext(w, h, n_cameras=frames, levels=levels, ring_slots=1, stream=rig.stream.cuda_stream, max_points=1024, max_targets=64)
        rng = np.random.default_rng(0)
        base = [rng.integers(0, 256, size=(h, w), dtype=np.uint8) for _ in range(2)]
        ctx.push_frames([base[i % 2] for i in range(frames)])
        cams = np.arange(frames, dtype=np.int32)
        with torch.cuda.stream(rig.stream):
            for _ in range(3):
                ctx.rebuild_pyramids(cams)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
            for a, b in ev:
                rig.flush.zero_()
                a.record(rig.stream)
                ctx.rebuild_pyramids(cams)
                b.record(rig.stream)
            rig.stream.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        alg = ctx.push_frame_bytes * frames
        out.append({"frame": "%dx%d" % (w, h), "levels": levels, "frames_per_launch": frames, "algorithmic_bytes": alg,
                    "ms": ms, "achieved": alg / ms / 1e6, "frac": alg / ms / 1e6 / peak})
        ctx.close()
    return out


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="push each frame inside its own step (no frame in flight)")
    args = ap.parse_args()
    warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    w, h, ncam, ntgt, per, levels = WORKLOADS[args.workload]
    npts_cfg = ncam * ntgt * per if ntgt else ncam * (((w - 3) // 4) + 1) * (((h - 3) // 4) + 1)
    config = {"workload": "%s: %d-view rig %dx%d gray per GPU, %s, %d-level pyramid, 21x21 window, fwd+bwd LK + FB gate%s"
                          % (args.workload, ncam, w, h, ("%d targets/view x %d seeds" % (ntgt, per)) if ntgt else "dense grid stride 4 px",
                             levels, " + LocalSearchKLT box update + Shi-Tomasi re-detect" if ntgt else ""),
              "cameras_per_gpu": ncam, "points_per_step_per_gpu": npts_cfg, "lk_solves_per_step_per_gpu": 2 * npts_cfg,
              "l2": "flushed between timed steps (256 MiB memset)",
              "frames_in_flight": "1: each step tracks the frame pushed by the step before and ingests the next one (H2D + pyramid on "
                                  "the upload stream, PSNW_MEM_OVERLAP) under its LK launch" if not args.no_pipeline else "0: push, then track", "parallelism": "cameras sharded, %d rig(s), no collective" % world}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb = cpu_measure(args.workload, args.steps, warmup, budget_s=120.0)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "points/s", "n_gpus": args.gpus,
                "steps": cb["camera_frames"], "warmup": warmup, "ms_per_step": cb["ms_per_camera_frame"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16/f32 (OpenCV fixed point)",
                "data": "synthetic", "config": config, "frames_per_s": cb["frames_per_s"], "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    rig = GpuRig(args.workload, rank, local, pipeline=not args.no_pipeline)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    sec, launches = rig.timed(args.steps, warmup, host=False, sampler=sampler)
    barrier()
    clocks = sampler.stop() if sampler else None
    sec = max_over_ranks(sec)
    pts_step = rig.npts * world
    value = pts_step * args.steps / sec

    # ---- end to end through the C ABI with pinned host buffers -----------------------------------------
    barrier()
    sec_e2e, _ = rig.timed(args.steps, warmup, host=True)
    barrier()
    sec_e2e = max_over_ranks(sec_e2e)
    h2d, d2h = rig.e2e_bytes()
    e2e = {"value": pts_step * args.steps / sec_e2e, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": 1e3 * sec_e2e / args.steps, "frames_per_s": ncam * world * args.steps / sec_e2e}

    line = None
    if rank == 0:
        # ---- roofline of the HBM-bound kernel group (pyramid + Scharr), timed alone with events ------------
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
        rig.prime(False)
        n0 = rig.ctx.launch_count
        rig.ctx.rebuild_pyramids(rig.cams)
        pyr_launches = rig.ctx.launch_count - n0
        pyr_mean, pyr_best = rig.time_pyramid()
        alg = rig.ctx.push_frame_bytes * ncam
        ach = alg / pyr_mean / 1e9
        traffic = None
        if args.workload == "c2":
            try:      # DRAM bytes of the same launch group from the committed ncu --set full capture
                traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_pyr_traffic.json")))["traffic_bytes_per_group"]
            except Exception:
                pass
        pyr_kernel = ("pyr_scharr_kernel (%d PDL-chained launches = one per level, each writing its level's REFLECT_101 apron with extra CTAs; "
                      "%d frames batched in grid.z)" % (pyr_launches, ncam)) if pyr_launches >= levels else \
                     ("pyr_fused_kernel (levels 0-%d in one launch, TMA-staged, halo recompute) + apron_kernel%s; %d launches, %d frames "
                      "batched in grid.z" % (min(levels, 4) - 1, " + pyr_scharr_kernel for the last level" if pyr_launches > 2 else "",
                                             pyr_launches, ncam))
        roofline = {"kernel": pyr_kernel,
                    "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "algorithmic_bytes": alg, "ms": 1e3 * pyr_mean, "best_ms": 1e3 * pyr_best, "peak_source": peak_src,
                    "note": "the HBM-bound kernel group of the path (SURVEY 8d); the kernel that dominates the step, lk_w21_kernel, is "
                            "neither HBM- nor tensor-bound: its SM-issue roofline is in lk.issue_roofline"}
        roofline["frames_in_flight_sweep"] = pyramid_sweep(rig, peak) if args.workload == "c2" else None
        lk_sec, st_frac = rig.time_track()
        lk = {"kernel": "lk_w21_kernel (fwd+bwd fused, persistent warps, TMA-staged patches)", "ms": 1e3 * lk_sec,
              "lk_solves_per_s": 2 * rig.npts / lk_sec, "status_ok_frac": st_frac,
              "bound": "SM issue slots (not HBM, not tensor) -- see DESIGN.md"}
        if args.workload == "c2":
            try:      # issue-slot utilisation and instruction count of the same launch from the committed ncu --set full capture
                lk["ncu"] = json.load(open(os.path.join(ROOT, "profiles", "r01_lk_issue.json")))
                # instruction roofline of the launch: executed warp-instructions (ncu count per solve x solves of this launch)
                # over the live launch time, against SMs x 4 schedulers x 1 instruction per clock at the clock sampled above
                props = torch.cuda.get_device_properties(torch.cuda.current_device())
                mhz = ((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
                peak_issue = props.multi_processor_count * 4 * mhz * 1e6
                ach_issue = lk["ncu"]["warp_instructions_per_solve"] * 2 * rig.npts / lk_sec
                lk["issue_roofline"] = {"bound": "sm issue", "achieved": ach_issue / 1e9, "peak": peak_issue / 1e9,
                                        "unit": "G warp-instructions/s", "frac": ach_issue / peak_issue}
            except Exception:
                pass
        line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8/int16 fixed point, f32 solve, f64 box", "data": "synthetic", "config": config,
                "frames_per_s": ncam * world * args.steps / sec, "lk_solves_per_s": 2 * value,
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "lk": lk}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_measure(args.workload, 10 ** 9, 3, budget_s=15.0)
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
