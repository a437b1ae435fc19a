#!/usr/bin/env python
"""bench.py -- KLT tracked points/s and frames/s of the Tracker2D hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4]

Workload (BASELINE.json configs[1], "c2"): PETS2009-shaped 4-view rig, 768x576 gray, 40 targets per view,
100 Shi-Tomasi seeds per target, 4-level pyramid, 21x21 window, forward + backward LK with the FB gate.
One STEP = one frame time of the rig on one GPU:
    psnw_push_frames   4 new camera frames -> Gaussian pyramid + Scharr derivatives (once per frame)
    psnw_klt_track     16 000 seeds slot 1 -> slot 0, forward + backward + FB check (32 000 LK solves), one launch
    psnw_box_update    160 targets, LocalSearchKLT box motion from the FB-validated pairs
    psnw_klt_detect    new seeds for the next step inside the 160 detection boxes of the new frame
Every rank (one process per GPU, torchrun) runs its own rig: cameras shard across GPUs with no collective on
the data path -> "scaling": "weak".  `value` = tracked points/s with frames, seeds and boxes
already resident in HBM; `e2e` = the same step with the frames in pinned HOST memory and the results read back every
step; `e2e_host_abi` = the step as a CUDA-free C++ host drives it: EVERY input and output of every call is a host
pointer handed to the C ABI (PSNW_MEM_PINNED), no torch copy anywhere.  L2 is flushed (256 MiB memset) between
timed steps; time = sum of per-step CUDA-event intervals on the context stream, max over ranks.  N > 1 adds
`e2e_gather`: the per-camera boxes of every rank land in ONE shared pinned host segment that rank 0 reads every
step (north_star's only cross-GPU exchange, psn_where/PSNWhere.cpp:269), timed by the host clock.

Beside the headline the line carries `configs` (short legs of the other BASELINE configurations: c1, c3, c4 and a
C5 point, 16 frames in flight), `sustained` (the K-step block repeated for >= 1 s), `roofline` (pyramid + Scharr
kernels against the measured HBM bandwidth) and `lk` (the dominant kernel with its arithmetic roofline from
iteration counts measured in the run).

--impl reference: the reference's CPU path (OpenCV calcOpticalFlowPyrLK via cv2 with all host threads + the C
restatement of LocalSearchKLT / selection), rank 0 only.  This leg and `cpu_baseline` are the only places that
touch oracle/.
"""
import argparse
import ctypes as C
import hashlib
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
E2E_BLOCKS = 5          # end-to-end legs: blocks of `steps` steps, the median block is reported
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

    def __init__(self, wl, rank, device, pipeline=True, blob_host=None):
        """blob_host: optional page-locked uint8 numpy array (>= blob_bytes) that receives the step's result blob in the
        end-to-end arm instead of a private pinned tensor (the shared segment of the multi-GPU gather)."""
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
            self.blob_offs = [int(o) for o in offs]
            if blob_host is not None:
                assert blob_host.nbytes >= int(offs[-1])
                self.h_blob = t.from_numpy(blob_host[:int(offs[-1])])
            else:
                self.h_blob = t.zeros(int(offs[-1]), dtype=t.uint8).pin_memory()
            # (the end-to-end leg reads step i's result while step i+1 is already enqueued: two result buffers in turn)
            self.h_blobs = [self.h_blob, t.zeros(int(offs[-1]), dtype=t.uint8).pin_memory()]
            self.blob_sel = 0
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
            # detection boxes of every rendered time step (device-resident arm: box_update reads the step's own boxes, no copy)
            self.np_boxes_all = [np.ascontiguousarray(np.concatenate([self.boxes[c][k] for c in range(self.ncam)]).reshape(-1, 4))
                                 if self.ntgt else np.zeros((0, 4)) for k in range(CYCLE)]
            self.d_boxes_all = [t.from_numpy(b).to(self.dev) for b in self.np_boxes_all]
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

    def close(self):
        self.ctx.close()

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
        elif host == "abi":
            self._detect(pingpong(0), self.h_seeds[self.which], self.h_counts_det, self.klt.MEM_PINNED)
        else:
            self._detect(pingpong(0), self.d_seeds[self.which], self.d_counts_det, self.klt.MEM_DEVICE)
        if self.pipeline:                    # one frame in flight: the frame a step tracks into was pushed by the step before
            self._push(pingpong(1), host)
        self.ctx.sync()
        self.step_idx = 1

    def step(self, host):
        """One frame time of the rig.
        host=False: frames already in HBM.  host=True (end to end): the step's inputs -- the 4 new frames, in pinned
        host memory, and the detection boxes -- are copied in, and the step's results (tracked points, status, updated
        boxes, inlier counts and mask) are copied out to pinned host memory.  Seeds are internal state of the path (detect
        output -> next step's track input) and stay on the device.  host="abi": see step_abi.
        Detection of the next step's seeds runs on the auxiliary stream and overlaps with tracking."""
        if host == "abi":
            return self.step_abi()
        # (grid workloads keep their seeds: self.which is not flipped)
        klt, ctx = self.klt, self.ctx
        t_idx, t_prev = pingpong(self.step_idx), pingpong(self.step_idx - 1)
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
            # the boxes the seeds were detected in (frame t_prev) move with the flow; end to end they arrive from the host
            if host:
                self.h_boxes_in.numpy()[:] = self.np_boxes_all[t_prev]
                self.d_boxes_in.copy_(self.h_boxes_in, non_blocking=True)
                boxes_in = self.d_boxes_in
            else:
                boxes_in = self.d_boxes_all[t_prev]
            ctx.box_update_raw(boxes_in.data_ptr(), self.ntargets, seeds.data_ptr(), self.d_out.data_ptr(), self.d_status.data_ptr(),
                               self.d_off.data_ptr(), self.d_boxes_out.data_ptr(), self.d_mask.data_ptr(), self.d_counts.data_ptr(), klt.MEM_DEVICE)
            self.which = nxt
        ctx.join()
        if host:
            self.h_blobs[self.blob_sel].copy_(self.d_blob, non_blocking=True)
        self.step_idx += 1

    def step_abi(self):
        """The step as a CUDA-free host drives it: every array of every call is a HOST pointer given to the
        C ABI with PSNW_MEM_PINNED -- frames, detection boxes, seeds, tracked points, status, boxes, masks, counts -- and
        the library does every copy itself; the caller only calls psnw_sync before it reads the results.  Same calls and
        the same one-frame-in-flight order as step(); seeds travel detect -> host -> track like any other output."""
        klt, ctx = self.klt, self.ctx
        assert not self.grid
        t_idx, t_prev = pingpong(self.step_idx), pingpong(self.step_idx - 1)
        cur, nxt = self.which, self.which ^ 1
        seeds, seeds_next = self.h_seeds[cur], self.h_seeds[nxt]
        jobs = self._jobs(seeds.data_ptr(), self.h_out.data_ptr(), self.h_status.data_ptr(), self.h_err.data_ptr(), self.h_fb.data_ptr())
        self._detect(t_idx, seeds_next, self.h_counts_det, klt.MEM_PINNED | klt.MEM_OVERLAP)
        ctx._ck(ctx.lib.psnw_klt_track_batch(ctx._h, jobs, self.ncam, klt.MEM_PINNED))
        self._push(pingpong(self.step_idx + 1), True, overlap=True)
        self.h_boxes_in.numpy()[:] = self.np_boxes_all[t_prev]
        ctx.box_update_raw(self.h_boxes_in.data_ptr(), self.ntargets, seeds.data_ptr(), self.h_out.data_ptr(), self.h_status.data_ptr(),
                           self.h_off.data_ptr(), self.h_boxes_out.data_ptr(), self.h_mask.data_ptr(), self.h_counts.data_ptr(), klt.MEM_PINNED)
        self.which = nxt
        ctx.join()
        self.step_idx += 1

    def abi_bytes(self):
        n, nt = self.npts, self.ntargets
        h2d = self.ncam * self.w * self.h + nt * 32 + n * 8 + (nt * 32 + (nt + 1) * 4 + n * 8 + n * 8 + n)   # frames, detect targets, seeds | boxes, offsets, pre, cur, status
        d2h = n * 8 + n + n * 4 + n * 4 + (nt * 32 + n + nt * 4) + n * 8 + nt * 4                           # out, status, err, fb | boxes, mask, counts | next seeds, counts
        return h2d, d2h

    def e2e_bytes(self):
        n, nt = self.npts, self.ntargets
        if self.grid:
            return self.ncam * self.w * self.h, self.blob_bytes
        h2d = self.ncam * self.w * self.h + nt * 32          # frames + detection boxes (PSN_Rect)
        return h2d, self.blob_bytes                          # points, boxes, inlier counts, status, inlier mask (one blob)

    def timed(self, steps, warmup, host, sampler=None):
        """-> (total seconds over `steps`, launches) with L2 flushed between steps; events on the context stream.
        warmup < 0: continue from the rig's current state (no priming, no warm-up steps)."""
        t = self.torch
        if warmup >= 0:
            self.prime(host)
        with t.cuda.stream(self.stream):
            for _ in range(max(warmup, 0)):
                self.step(host)
                if host:
                    self.ctx.sync()
            self.stream.synchronize()
            ev = [(t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)) for _ in range(steps)]
            l0 = self.ctx.launch_count
            if sampler:
                sampler.start()
            # end to end (host=True) the host consumes step i's result while step i+1 is already enqueued -- the order a host that
            # feeds an associator works in (one frame in flight): the result lands in one of two pinned buffers, an event marks it
            # complete, and the host waits for that event (and reads the buffer) after it has enqueued the next step.  Every step's
            # interval still contains the H2D copy of its frames and the D2H copy of its result; what no longer enters is how
            # fast THIS host's Python issues the next step's calls.  host="abi" keeps one psnw_sync per step.
            lag = host is True
            done = [t.cuda.Event(), t.cuda.Event()]
            seen = 0
            for i in range(steps):
                self.flush.zero_()
                ev[i][0].record(self.stream)
                self.blob_sel = i & 1 if lag else 0
                self.step(host)
                ev[i][1].record(self.stream)
                if lag:
                    done[i & 1].record(self.stream)
                    if i > 0:
                        done[(i - 1) & 1].synchronize()
                        seen += int(self.h_blobs[(i - 1) & 1][-1])       # the host touches the result it waited for
                elif host:
                    self.ctx.sync()          # the step's result (boxes, points) is read on the host every step
            if lag:
                done[(steps - 1) & 1].synchronize()
                seen += int(self.h_blobs[(steps - 1) & 1][-1])
                self.blob_sel = 0
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

    def time_track(self, reps=20, prime=True):
        """LK kernel alone (forward+backward fused launch): the newest frame's seeds tracked into the previous frame."""
        t = self.torch
        if prime:
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


def _events(torch, n):
    return [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]


def event_floor(rig, reps=20):
    """What the protocol itself costs: an (almost) empty launch between the same pair of CUDA events, after the same L2
    flush.  No kernel timed this way can come out shorter."""
    t = rig.torch
    one = t.zeros(1, dtype=t.uint8, device=rig.dev)
    with t.cuda.stream(rig.stream):
        ev = _events(t, reps + 3)
        for a, b in ev:
            rig.flush.zero_()
            a.record(rig.stream)
            one.zero_()
            b.record(rig.stream)
        rig.stream.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev[3:]])) * 1e-3


def pyramid_modes(rig, w, h, levels, frames, isolated_reps=10, stream_launches=48):
    """K1 launch group for `frames` camera-frames of w x h per launch, timed two ways:
    isolated  one launch between two events, L2 flushed (256 MiB memset) before each: launch latency included;
    stream    `stream_launches` launches back to back between ONE pair of events, each on another ring slot, the ring
              sized so that the planes touched between two visits of a slot exceed the 126 MB L2 twice over (inputs larger
              than L2 instead of a flush): the kernel's average duration inside a stream of launches."""
    import torch
    from mcmtt_opticalflow_b200 import klt
    rng = np.random.default_rng(0)
    base = [rng.integers(0, 256, size=(h, w), dtype=np.uint8) for _ in range(2)]
    probe = klt.KLTContext(w, h, n_cameras=1, levels=levels, ring_slots=1, stream=rig.stream.cuda_stream, max_points=1024, max_targets=64)
    per_frame = probe.push_frame_bytes
    probe.close()
    slots = int(min(max(np.ceil((256 << 20) / (per_frame * frames)), 4), 32))
    ctx = klt.KLTContext(w, h, n_cameras=frames, levels=levels, ring_slots=slots, stream=rig.stream.cuda_stream, max_points=1024, max_targets=64)
    cams = np.arange(frames, dtype=np.int32)
    for _ in range(slots):
        ctx.push_frames([base[i % 2] for i in range(frames)])
    n0 = ctx.launch_count
    ctx.rebuild_pyramids(cams)
    launches = ctx.launch_count - n0
    with torch.cuda.stream(rig.stream):
        for _ in range(3):
            ctx.rebuild_pyramids(cams)
        ev = _events(torch, isolated_reps)
        for a, b in ev:
            rig.flush.zero_()
            a.record(rig.stream)
            ctx.rebuild_pyramids(cams)
            b.record(rig.stream)
        rig.stream.synchronize()
        iso = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        n = max(stream_launches, 2 * slots)
        for i in range(slots):
            ctx.rebuild_pyramids(cams, slot=i % slots)
        (a, b), = _events(torch, 1)
        # the launches are enqueued while the stream is still busy with a few L2-flush memsets (~60 us each), so that they
        # run back to back at the device's pace whatever the host's enqueue rate is (16 host threads may still be spinning
        # after the CPU arm); the start event is recorded behind the memsets
        for _ in range(16):
            rig.flush.zero_()
        a.record(rig.stream)
        for i in range(n):
            ctx.rebuild_pyramids(cams, slot=i % slots)
        b.record(rig.stream)
        rig.stream.synchronize()
        strm = a.elapsed_time(b) / n
    alg = per_frame * frames
    ctx.close()
    return {"frame": "%dx%d" % (w, h), "levels": levels, "frames_per_launch": frames, "algorithmic_bytes": alg, "kernels_per_launch_group": launches,
            "isolated_ms": iso, "stream_ms": strm, "ring_slots": slots, "working_set_bytes": int(per_frame * frames * slots)}


def sources_hash(names):
    h = hashlib.sha256()
    for n in names:
        with open(os.path.join(ROOT, "mcmtt_opticalflow_b200", "csrc", n), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def committed_profile(name, sources):
    """A number taken from a committed ncu capture is used only if the capture was made from the sources as they are now."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return None
    return d if d.get("sources_sha256") == sources_hash(sources) else None


PYR_SOURCES = ["pyramid.cu", "pyramid_fused.cu", "pyr_common.cuh", "common.cuh"]
LK_SOURCES = ["lk.cu", "pyr_common.cuh", "common.cuh"]


def lk_report(rig, clocks, props):
    """The dominant kernel: live launch time, and its ARITHMETIC roofline from work counted in the run -- SURVEY 8d's model
    (template pass 441 px x 15 lane-ops, iteration 441 px x 7) x the template passes and iterations the counting build of the
    kernel reports for this very batch, over launch time x SMs x 128 lanes x clock."""
    lk_sec, st_frac = rig.time_track()
    rig.ctx.lk_stats(True)                       # counters zeroed; exactly ONE launch of the timed batch is counted
    rig.time_track(reps=1, prime=False)
    st = rig.ctx.lk_stats(False)
    solves = max(st["solves"], 1)
    mhz = ((clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
    lane_ops = 441.0 * (15.0 * st["template_passes"] + 7.0 * st["iterations"])      # of that one launch
    peak_ops = props.multi_processor_count * 128 * mhz * 1e6
    lk = {"kernel": "lk_w21_kernel (fwd+bwd fused, persistent warps, TMA-staged patches)", "ms": 1e3 * lk_sec,
          "lk_solves_per_s": 2 * rig.npts / lk_sec, "status_ok_frac": st_frac,
          "bound": "SM issue / ALU lanes (not HBM, not tensor) -- see DESIGN.md",
          "counted": {"solves_run": st["solves"], "nominal_solves": 2 * rig.npts, "template_passes_per_solve": st["template_passes"] / solves,
                      "iterations_per_solve": st["iterations"] / solves, "iterations_per_s": st["iterations"] / lk_sec,
                      "note": "one launch; the backward solve of a seed the forward pass lost is not run"},
          "alg_roofline": {"bound": "alu lanes", "model": "441 x (15 x template passes + 7 x iterations) lane-ops (SURVEY 8d)",
                           "achieved": lane_ops / lk_sec / 1e12, "peak": peak_ops / 1e12, "unit": "T lane-ops/s",
                           "frac": lane_ops / lk_sec / peak_ops, "sm_mhz": mhz}}
    ncu = committed_profile("r02_lk_issue.json", LK_SOURCES)
    if ncu:       # executed warp-instructions of the same launch from the committed ncu capture of THESE sources
        peak_issue = props.multi_processor_count * 4 * mhz * 1e6
        ach_issue = ncu["warp_instructions_per_solve"] * 2 * rig.npts / lk_sec
        lk["ncu"] = ncu
        lk["issue_roofline"] = {"bound": "sm issue", "achieved": ach_issue / 1e9, "peak": peak_issue / 1e9,
                                "unit": "G warp-instructions/s", "frac": ach_issue / peak_issue}
    return lk


def config_leg(wl, local, steps, peak):
    """A short leg of another BASELINE configuration on the same GPU: resident and end-to-end step, LK launch, pyramid group."""
    w, h, ncam, ntgt, per, levels = WORKLOADS[wl]
    rig = GpuRig(wl, 0, local)
    try:
        sec, launches = rig.timed(steps, 3, host=False)
        sec_e, _ = rig.timed(steps, 3, host=True)
        rig.prime(False)
        lk_sec, st_frac = rig.time_track(reps=10)
        rig.ctx.lk_stats(True)                     # counting build of the LK kernel, ONE launch of the same batch
        rig.time_track(reps=1, prime=False)
        cnt = rig.ctx.lk_stats(False)
        pyr_mean, _ = rig.time_pyramid(reps=10)
        alg = rig.ctx.push_frame_bytes * ncam
        h2d, d2h = rig.e2e_bytes()
        return {"workload": wl, "frame": "%dx%d" % (w, h), "cameras_per_gpu": ncam, "levels": levels, "points_per_step": rig.npts, "steps": steps,
                "ms_per_step": 1e3 * sec / steps, "value": rig.npts * steps / sec, "frames_per_s": ncam * steps / sec,
                "e2e": {"value": rig.npts * steps / sec_e, "ms_per_step": 1e3 * sec_e / steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "lk_ms": 1e3 * lk_sec, "lk_solves_per_s": 2 * rig.npts / lk_sec, "status_ok_frac": st_frac,
                "lk_counted": {"solves_run": cnt["solves"], "nominal_solves": 2 * rig.npts,
                               "template_passes_per_nominal_solve": cnt["template_passes"] / (2.0 * rig.npts),
                               "iterations_per_nominal_solve": cnt["iterations"] / (2.0 * rig.npts),
                               "iterations_per_s": cnt["iterations"] / lk_sec,
                               "note": "lk_solves_per_s counts two solves per seed; the backward solve of a seed the forward pass lost is "
                                       "not run.  iterations_per_s compares configurations whose seeds need different numbers of iterations"},
                "pyramid": {"isolated_ms": 1e3 * pyr_mean, "algorithmic_bytes": alg, "frac_isolated": alg / pyr_mean / 1e9 / peak},
                "gpu_launches": int(launches)}
    finally:
        rig.close()


def c5_leg():
    """BASELINE config 5, one point: 16 frames in flight per camera (tools/frames_in_flight.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("frames_in_flight", os.path.join(ROOT, "tools", "frames_in_flight.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):          # (the tool prints its rows; bench.py prints ONE line)
        return mod.run("768x576", [16])[0]


class SharedSegment:
    """One page-locked host segment shared by the ranks of the node (a file in /dev/shm mapped by every rank and registered
    with the CUDA runtime): slot r = rank r's result blob + a sequence number; the last 64 bytes hold rank 0's acknowledgement.
    The multi-GPU exchange of north_star -- per-camera boxes to the host that runs Associator3D (psn_where/PSNWhere.cpp:269) --
    without pickling and without a collective."""

    def __init__(self, torch, dist, rank, world, blob_bytes):
        self.torch, self.rank, self.world = torch, rank, world
        self.slot = (blob_bytes + 63) // 64 * 64 + 64
        self.total = self.slot * world + 64
        self.path = "/dev/shm/psnw_gather_%s" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(self.total)
        dist.barrier()
        self.mm = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(self.total,))
        self.registered = int(torch.cuda.cudart().cudaHostRegister(self.mm.ctypes.data, self.total, 0)) == 0
        if not self.registered:
            raise RuntimeError("cudaHostRegister of the shared segment failed")
        self.seq = [self.mm[r * self.slot + self.slot - 64: r * self.slot + self.slot - 56].view(np.int64) for r in range(world)]
        self.ack = self.mm[self.total - 64: self.total - 56].view(np.int64)
        dist.barrier()

    def blob(self, r):
        return self.mm[r * self.slot: r * self.slot + self.slot - 64]

    def close(self, dist):
        try:
            self.torch.cuda.cudart().cudaHostUnregister(self.mm.ctypes.data)
        except Exception:
            pass
        dist.barrier()
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def gather_leg(rig, seg, steps, warmup, barrier):
    """End-to-end step + the gather: every rank's result blob lands in its slot of the shared segment (D2H straight into it);
    rank 0 waits for all slots of the step and reads every rank's boxes and inlier counts (what Associator3D consumes);
    a rank starts its next step when rank 0 has taken the previous one.  Host clock over `steps` steps between barriers."""
    torch = rig.torch
    rank, world = seg.rank, seg.world
    o = rig.blob_offs
    checksum = 0.0
    rig.prime(True)
    with torch.cuda.stream(rig.stream):
        for _ in range(warmup):
            rig.step(True)
            rig.ctx.sync()
        barrier()
        t0 = time.perf_counter()
        for i in range(1, steps + 1):
            while int(seg.ack[0]) < i - 1:        # rank 0 has consumed step i - 1
                pass
            rig.flush.zero_()
            rig.step(True)
            rig.ctx.sync()
            seg.seq[rank][0] = i
            if rank == 0:
                for r in range(world):
                    while int(seg.seq[r][0]) < i:
                        pass
                    b = seg.blob(r)
                    boxes = b[o[1]:o[1] + rig.ntargets * 32].view(np.float64)
                    counts = b[o[2]:o[2] + rig.ntargets * 4].view(np.int32)
                    checksum += float(boxes.sum()) + float(counts.sum())
                seg.ack[0] = i
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    return t1 - t0, checksum


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the legs of the other BASELINE configurations and the pyramid sweep")
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
    affinity = None
    if world > 1:
        import torch.distributed as dist
        # every rank on its own cores: eight Python hosts otherwise share the cores the launcher happened to leave them
        try:
            cores = sorted(os.sched_getaffinity(0))
            share = max(len(cores) // world, 1)
            mine = cores[local * share:(local + 1) * share] or cores
            os.sched_setaffinity(0, mine)
            affinity = [mine[0], mine[-1]]
        except Exception:
            affinity = None
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

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

    def all_ranks(x):
        if dist is None:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    seg = None
    rig = GpuRig(args.workload, rank, local, pipeline=not args.no_pipeline)
    if world > 1 and not rig.grid:
        seg = SharedSegment(torch, dist, rank, world, rig.blob_bytes)
        rig.close()
        rig = GpuRig(args.workload, rank, local, pipeline=not args.no_pipeline, blob_host=seg.blob(rank))

    # ---- device-resident throughput -------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    sec_own, launches = rig.timed(args.steps, warmup, host=False, sampler=sampler)
    barrier()
    clocks = sampler.stop() if sampler else None
    sec = max_over_ranks(sec_own)
    per_rank = all_ranks(1e3 * sec_own / args.steps)
    pts_step = rig.npts * world
    value = pts_step * args.steps / sec

    # ---- end to end: frames in pinned host memory, results read back every step -------------------------
    # (the host takes part in every step here -- copies enqueued, a sync, the result read -- so a single host hiccup shows in a
    # 20-step block: E2E_BLOCKS blocks of exactly `steps` steps are timed, each the max over ranks, and the MEDIAN block is
    # reported with all of them listed)
    def e2e_blocks(host):
        own, glob = [], []
        for b in range(E2E_BLOCKS):
            barrier()
            sec_b, _ = rig.timed(args.steps, warmup if b == 0 else -1, host=host)
            barrier()
            own.append(sec_b)
            glob.append(max_over_ranks(sec_b))
        k = sorted(range(E2E_BLOCKS), key=lambda i: glob[i])[E2E_BLOCKS // 2]
        return own[k], glob[k], [1e3 * g / args.steps for g in glob]

    sec_e2e_own, sec_e2e, e2e_all = e2e_blocks(True)
    per_rank_e2e = all_ranks(1e3 * sec_e2e_own / args.steps)
    h2d, d2h = rig.e2e_bytes()
    e2e = {"value": pts_step * args.steps / sec_e2e, "unit": "points/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": 1e3 * sec_e2e / args.steps, "frames_per_s": ncam * world * args.steps / sec_e2e,
           "blocks_ms_per_step": e2e_all, "of": "median of %d blocks of %d steps" % (E2E_BLOCKS, args.steps),
           "how": "every step copies its 4 frames and its detection boxes in from pinned host memory and its result blob (points, status, "
                  "boxes, inlier mask, counts) out to pinned host memory inside the timed interval; the host waits for step i's result "
                  "(an event behind its D2H copy) and reads it after it has enqueued step i+1 -- one frame in flight on the readback side "
                  "too, two result buffers in turn"}

    # ---- end to end with every array a host pointer handed to the C ABI (no torch copies) ----------------
    e2e_abi = None
    if not rig.grid:
        _, sec_abi, abi_all = e2e_blocks("abi")
        a_h2d, a_d2h = rig.abi_bytes()
        e2e_abi = {"value": pts_step * args.steps / sec_abi, "unit": "points/s", "ms_per_step": 1e3 * sec_abi / args.steps,
                   "blocks_ms_per_step": abi_all,
                   "h2d_bytes_per_step": a_h2d, "d2h_bytes_per_step": a_d2h, "frames_per_s": ncam * world * args.steps / sec_abi,
                   "how": "psnw_push_frames / psnw_klt_detect_batch / psnw_klt_track_batch / psnw_box_update with PSNW_MEM_PINNED host pointers for "
                          "every input and output (frames, boxes refreshed per step, seeds, points, status, err, fb_err, boxes, mask, counts), "
                          "psnw_sync once per step; the gap to e2e is the seeds and the tracked points crossing PCIe twice (detect -> host -> "
                          "track, track -> host -> box_update) instead of staying in HBM"}

    # ---- the gather (N > 1): per-camera boxes of every rank into one shared pinned segment, read by rank 0 every step ----
    e2e_gather = None
    if seg is not None:
        barrier()
        sec_g, checksum = gather_leg(rig, seg, args.steps, warmup, barrier)
        barrier()
        sec_g = max_over_ranks(sec_g)
        e2e_gather = {"value": pts_step * args.steps / sec_g, "unit": "points/s", "ms_per_step": 1e3 * sec_g / args.steps,
                      "frames_per_s": ncam * world * args.steps / sec_g, "clock": "host (perf_counter between barriers), max over ranks",
                      "gathered_bytes_per_step": world * (rig.ntargets * 36), "checksum": checksum,
                      "how": "each rank's D2H result blob lands in its slot of one /dev/shm segment registered with cudaHostRegister by every "
                             "rank; rank 0 polls the slots' sequence numbers and reads boxes + inlier counts of all ranks every step (lock step: a "
                             "rank starts step i+1 when rank 0 has taken step i); no pickle, no collective"}
        seg.close(dist)

    barrier()
    if dist is not None:          # the collective part is over: the other ranks leave, rank 0 finishes the single-GPU reports alone
        dist.destroy_process_group()
        if rank != 0:
            return 0

    # ---- sustained: the K-step block repeated until >= 1 s of device time has been timed ----------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    props = torch.cuda.get_device_properties(torch.cuda.current_device())
    sustained = None
    if world == 1:
        s2 = ClockSampler(local, period=0.02)
        blocks, total = [], 0.0
        rig.prime(False)
        s2.start()
        while total < 1.0 and len(blocks) < 500:
            bsec, _ = rig.timed(args.steps, -1, host=False)
            blocks.append(bsec)
            total += bsec
        c2 = s2.stop()
        rates = sorted(rig.npts * args.steps / b for b in blocks)
        sustained = {"blocks": len(blocks), "steps_per_block": args.steps, "timed_seconds": total, "value_median": rates[len(rates) // 2],
                     "value_min": rates[0], "value_max": rates[-1], "clocks": c2}

    # ---- roofline of the HBM-bound kernel group (pyramid + Scharr) -------------------------------------------------
    rig.prime(False)
    floor = event_floor(rig)
    main_modes = pyramid_modes(rig, w, h, levels, ncam)
    alg = main_modes["algorithmic_bytes"]
    strm_s, iso_s = main_modes["stream_ms"] * 1e-3, main_modes["isolated_ms"] * 1e-3
    prof = committed_profile("r02_pyr_traffic.json", PYR_SOURCES) if args.workload == "c2" else None
    roofline = {"kernel": "pyramid + Scharr + apron launch group of one rig frame (%d kernel(s) per group; builder chosen by launch size: "
                          "fused multi-level kernel / per-level kernels / hybrid, DESIGN.md section 3 K1), %d frames batched in grid.z"
                          % (main_modes["kernels_per_launch_group"], ncam),
                "bound": "hbm", "achieved": alg / strm_s / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / strm_s / 1e9 / peak,
                "traffic": prof["traffic_bytes_per_group"] if prof else None, "traffic_source": prof.get("source") if prof else
                "no committed ncu capture of the current pyramid sources", "algorithmic_bytes": alg, "ms": 1e3 * strm_s, "peak_source": peak_src,
                "timing": "average launch-group duration inside a stream of %d back-to-back launches walking a ring of %d slots (%.0f MB touched "
                          "between two visits of a slot: inputs larger than L2, no flush), one pair of CUDA events around the stream, the launches "
                          "enqueued behind 16 L2-flush memsets so that the host's enqueue rate does not enter; consecutive builds of other ring "
                          "slots overlap (late_wait, DESIGN.md K1)"
                          % (max(48, 2 * main_modes["ring_slots"]), main_modes["ring_slots"], main_modes["working_set_bytes"] / 1e6),
                "isolated": {"ms": 1e3 * iso_s, "achieved": alg / iso_s / 1e9, "frac": alg / iso_s / 1e9 / peak,
                             "event_floor_ms": 1e3 * floor,
                             "note": "one launch group between its own pair of events after an L2 flush; an empty launch measured the same way "
                                     "costs event_floor_ms, so 12 MB can not come out above %.2f of the roofline this way" % (alg / floor / 1e9 / peak)},
                "note": "the HBM-bound kernel group of the path (SURVEY 8d); the kernel that dominates the step, lk_w21_kernel, is "
                        "neither HBM- nor tensor-bound: its arithmetic roofline is in lk.alg_roofline"}
    if args.workload == "c2" and not args.no_configs:
        sweep = []
        for (sw, sh, sl, sf) in [(768, 576, 4, 16), (768, 576, 4, 64), (1920, 1080, 4, 32), (3840, 2160, 5, 16)]:
            m = pyramid_modes(rig, sw, sh, sl, sf, isolated_reps=8, stream_launches=12)
            m["frac_isolated"] = m["algorithmic_bytes"] / m["isolated_ms"] / 1e6 / peak
            m["frac_stream"] = m["algorithmic_bytes"] / m["stream_ms"] / 1e6 / peak
            sweep.append(m)
        roofline["frames_in_flight_sweep"] = sweep

    lk = lk_report(rig, clocks, props)
    line = {"metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/int16 fixed point, f32 solve, f64 box", "data": "synthetic", "config": config,
            "frames_per_s": ncam * world * args.steps / sec, "lk_solves_per_s": 2 * value,
            "e2e": e2e, "e2e_host_abi": e2e_abi, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "lk": lk,
            "per_rank_ms_per_step": {"resident": {"min": min(per_rank), "median": float(np.median(per_rank)), "max": max(per_rank)},
                                     "e2e": {"min": min(per_rank_e2e), "median": float(np.median(per_rank_e2e)), "max": max(per_rank_e2e)},
                                     "cpu_affinity_rank0": affinity}}
    if e2e_gather:
        line["e2e_gather"] = e2e_gather
    if sustained:
        line["sustained"] = sustained
    rig.close()
    if world == 1 and not args.no_configs:
        legs = {}
        for wl, st in (("c1", 50), ("c2", 30), ("c3", 20), ("c4", 5)):
            if wl == args.workload:
                continue
            try:
                legs[wl] = config_leg(wl, local, st, peak)
            except Exception as e:          # a leg that fails must not cost the headline line
                legs[wl] = {"error": repr(e)}
        try:
            legs["c5_16_frames_in_flight"] = c5_leg()
        except Exception as e:
            legs["c5_16_frames_in_flight"] = {"error": repr(e)}
        line["configs"] = legs
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_measure(args.workload, 10 ** 9, 3, budget_s=15.0)
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
