"""GPU tier: overlapped frame intake (PSNW_MEM_OVERLAP on psnw_push_frames) -- the next frame is uploaded and its pyramid
built on the upload stream while the current frame is still being tracked.  Results must be bit-identical to the serial
order, including when the ring is so short that the push recycles a slot the LK launch in flight still reads."""
import numpy as np
import pytest

from mcmtt_opticalflow_b200 import klt, synth

pytestmark = pytest.mark.gpu


def _run(w, h, frames, pts, ring_slots, overlap, levels=4):
    """Tracks pts through consecutive frame pairs (always from the original seeds); -> list of (out, status, err, fb)."""
    import torch
    dev = torch.device("cuda", 0)
    n = len(pts)
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    d_pts = torch.from_numpy(pts).to(dev)
    res = []
    with klt.KLTContext(w, h, levels=levels, ring_slots=ring_slots, fb_thresh=1.0, max_points=max(n, 1024)) as ctx:
        outs = [dict(out=torch.zeros((n, 2), dtype=torch.float32, device=dev), status=torch.zeros(n, dtype=torch.uint8, device=dev),
                     err=torch.zeros(n, dtype=torch.float32, device=dev), fb=torch.zeros(n, dtype=torch.float32, device=dev))
                for _ in range(len(frames) - 1)]
        push = lambda k, ov: ctx.push_frames([pinned[k].data_ptr()], [0], k, klt.MEM_PINNED | (klt.MEM_OVERLAP if ov else 0), pitch=w)
        push(0, False)
        if overlap:
            push(1, False)
        for k in range(1, len(frames)):
            o = outs[k - 1]
            job = dict(cam=0, slot_from=1, slot_to=0, n_pts=n, pts=d_pts.data_ptr(), out=o["out"].data_ptr(),
                       status=o["status"].data_ptr(), err=o["err"].data_ptr(), fb_err=o["fb"].data_ptr())
            if overlap:
                ctx.track_batch([job], mem=klt.MEM_DEVICE)            # frames k-1 -> k, both already on the device
                if k + 1 < len(frames):
                    push(k + 1, True)                                   # ingest k+1 under the LK launch
            else:
                push(k, False)
                ctx.track_batch([job], mem=klt.MEM_DEVICE)
        ctx.sync()
        for o in outs:
            res.append(tuple(o[key].cpu().numpy() for key in ("out", "status", "err", "fb")))
    return res


@pytest.mark.parametrize("ring_slots", [2, 3, 4])
def test_overlapped_push_equals_serial(ring_slots):
    w, h = 768, 576
    cam = synth.Camera(w, h, 12, cam=2)
    frames = [cam.frame(k) for k in range(6)]
    pts, _ = synth.seeds_in_boxes(cam.boxes(0), 400, w, h, np.random.default_rng(9))    # a launch long enough to overlap
    serial = _run(w, h, frames, pts, ring_slots, overlap=False)
    piped = _run(w, h, frames, pts, ring_slots, overlap=True)
    assert len(serial) == len(piped) == 5
    for k, (a, b) in enumerate(zip(serial, piped)):
        for x, y in zip(a, b):
            assert np.array_equal(x, y), ("step", k)
    assert serial[0][1].mean() > 0.5


def test_overlapped_push_then_readers_wait(oracle):
    """Every reader of ring slots (read_level, detect, track) first waits for a pending overlapped push."""
    import torch
    w, h = 384, 288
    cam = synth.Camera(w, h, 4, cam=1)
    f0, f1 = cam.frame(0), cam.frame(1)
    p1 = torch.from_numpy(f1).pin_memory()
    with klt.KLTContext(w, h, levels=3) as ctx:
        ctx.push_frame(f0)
        ctx.push_frames([p1.data_ptr()], [0], 1, klt.MEM_PINNED | klt.MEM_OVERLAP, pitch=w)
        img, dxy = ctx.read_level(0, 0, 1)
        assert np.array_equal(img, oracle.pyr_down(f1)) and np.array_equal(dxy, oracle.scharr(oracle.pyr_down(f1)))
        img0, _ = ctx.read_level(0, 1, 0)
        assert np.array_equal(img0, f0)
        seeds, counts = ctx.detect(cam.boxes(1), 30, 0.01, 3.0)
        eig = ctx.read_min_eig()
        assert np.array_equal(eig, oracle.min_eig_map(f1))
        for t in range(4):
            assert np.array_equal(seeds[t, :counts[t]], oracle.detect(f1, cam.boxes(1)[t], 30, 0.01, 3.0, eig=eig))


def test_eager_min_eig_under_overlapped_detect(oracle):
    """PSNW_PUSH_EIG: the Shi-Tomasi map of the frame ingested ahead is built on the upload stream while an overlapped
    detection may still read the map of the frame before (two maps per camera).  Corners must equal the oracle's on every
    frame, and the cached map must be the one of the right frame."""
    import torch
    w, h, ntgt, per = 384, 288, 6, 40
    cam = synth.Camera(w, h, ntgt, cam=5)
    frames = [cam.frame(k) for k in range(5)]
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    dev = torch.device("cuda", 0)
    seeds = [torch.zeros((ntgt, per, 2), dtype=torch.float32, device=dev) for _ in frames]
    counts = [torch.zeros(ntgt, dtype=torch.int32, device=dev) for _ in frames]
    with klt.KLTContext(w, h, levels=3) as ctx:
        ctx.push_frames([pinned[0].data_ptr()], [0], 0, klt.MEM_PINNED, pitch=w)
        for k in range(len(frames)):
            boxes = np.ascontiguousarray(cam.boxes(k), np.float64)
            job = (klt.DetectJob * 1)()
            job[0].cam, job[0].slot, job[0].boxes, job[0].n_targets = 0, 0, boxes.ctypes.data, ntgt
            job[0].pts_xy, job[0].counts = seeds[k].data_ptr(), counts[k].data_ptr()
            ctx._ck(ctx.lib.psnw_klt_detect_batch(ctx._h, job, 1, per, 0.01, 3.0, klt.MEM_DEVICE | klt.MEM_OVERLAP))
            if k + 1 < len(frames):           # ingest the next frame, with its map, while the detection runs
                ctx.push_frames([pinned[k + 1].data_ptr()], [0], k + 1, klt.MEM_PINNED | klt.MEM_OVERLAP | klt.PUSH_EIG, pitch=w)
        ctx.sync()
        eig_last = ctx.read_min_eig(0, 0)
        eig_prev = ctx.read_min_eig(0, 1)
    assert np.array_equal(eig_last, oracle.min_eig_map(frames[-1]))
    assert np.array_equal(eig_prev, oracle.min_eig_map(frames[-2]))
    for k, f in enumerate(frames):
        eig = oracle.min_eig_map(f)
        s, c = seeds[k].cpu().numpy(), counts[k].cpu().numpy()
        for t in range(ntgt):
            want = oracle.detect(f, cam.boxes(k)[t], per, 0.01, 3.0, eig=eig)
            assert c[t] == len(want) and np.array_equal(s[t, :c[t]], want), (k, t)


def test_pipelined_bench_steps_are_periodic():
    """Race hunt on the full pipelined step of bench.py (frame in flight, overlapped detection, eager Shi-Tomasi map,
    persistent LK): the rig plays its frames ping-pong, so 50 free-running steps (no host sync in between) must repeat
    bit for bit with the period of the frame sequence.  tools/soak.py runs the same check for thousands of steps."""
    import hashlib
    import torch
    import bench
    rig = bench.GpuRig("c1", 0, 0)
    rig.prime(False)
    period = 2 * bench.CYCLE - 2
    snaps = []
    with torch.cuda.stream(rig.stream):
        for _ in range(5 * period):
            rig.step(False)
            snaps.append((rig.d_blob.clone(), rig.d_seeds[rig.which].clone()))
        rig.ctx.sync()
        rig.stream.synchronize()
    sums = [hashlib.sha1(b.cpu().numpy().tobytes() + s.cpu().numpy().tobytes()).hexdigest() for b, s in snaps]
    for k in range(2 * period, len(sums)):
        assert sums[k] == sums[k - period], k
    assert len(set(sums[2 * period:])) == period
    rig.ctx.close()


def _fuzz_run(monkeypatch, seed, relaxed, mode, n_cmds=60):
    """A random sequence of pushes (subsets of cameras, overlapped or not, with or without the eager Shi-Tomasi map),
    pyramid rebuilds of random slots, Shi-Tomasi detections (on the auxiliary stream or not) and LK launches on one context.  relaxed: everything the library may overlap is allowed to
    (upload stream, consecutive builds beside each other, programmatic launches); strict: no overlap flag, every build waits,
    a sync after every call.  -> (list of track outputs, every level of every plane at the end)."""
    import torch
    monkeypatch.setenv("PSNW_PYR_MODE", mode)
    monkeypatch.setenv("PSNW_PYR_OVERLAP", "1" if relaxed else "0")
    dev = torch.device("cuda", 0)
    w, h, ncam, slots, levels, n = 256, 192, 3, 3, 3, 300
    rng = np.random.default_rng(seed)
    pool = [torch.from_numpy(rng.integers(0, 256, size=(h, w), dtype=np.uint8)).pin_memory() for _ in range(8)]
    # smooth content so that LK has something to track: blur by averaging shifted copies
    for k, p in enumerate(pool):
        a = p.numpy().astype(np.float32)
        for _ in range(3):
            a = (a + np.roll(a, 1, 0) + np.roll(a, 1, 1) + np.roll(a, -1, 0) + np.roll(a, -1, 1)) / 5.0
        p.numpy()[:] = np.clip(np.roll(a, k, 1), 0, 255).astype(np.uint8)
    pts = np.stack([rng.uniform(20, w - 20, n), rng.uniform(20, h - 20, n)], axis=1).astype(np.float32)
    d_pts = torch.from_numpy(pts).to(dev)
    held = [0] * ncam
    outs, res, dets = [], [], []
    with klt.KLTContext(w, h, n_cameras=ncam, levels=levels, ring_slots=slots, fb_thresh=1.0, max_points=1024) as ctx:
        for c in range(ncam):                                   # every camera starts with two frames
            for k in range(2):
                ctx.push_frames([pool[(c + k) % 8].data_ptr()], [c], k, klt.MEM_PINNED, pitch=w)
            held[c] = 2
        ctx.sync()
        for step in range(n_cmds):
            op = int(rng.integers(0, 4))
            if op == 3:                                         # Shi-Tomasi detection in a few boxes of one camera's slot
                c = int(rng.integers(0, ncam))
                slot = int(rng.integers(0, held[c]))
                nb, per = 4, 40
                bx = np.stack([rng.uniform(0, w - 80, nb), rng.uniform(0, h - 80, nb), rng.uniform(30, 80, nb), rng.uniform(30, 80, nb)], axis=1)
                ov = rng.random() < 0.6
                d = dict(pts=torch.zeros((nb, per, 2), dtype=torch.float32, device=dev), cnt=torch.zeros(nb, dtype=torch.int32, device=dev),
                         boxes=np.ascontiguousarray(bx, np.float64))
                dets.append(d)
                ctx._ck(ctx.lib.psnw_klt_detect(ctx._h, c, slot, d["boxes"].ctypes.data, nb, per, 0.01, 3.0, d["pts"].data_ptr(), d["cnt"].data_ptr(),
                                                klt.MEM_DEVICE | (klt.MEM_OVERLAP if (relaxed and ov) else 0)))
            elif op == 0:                                       # push a subset of cameras
                cams = sorted(rng.choice(ncam, size=int(rng.integers(1, ncam + 1)), replace=False).tolist())
                flags = klt.MEM_PINNED
                ov = rng.random() < 0.7                          # (drawn in both runs: the sequences must stay identical)
                if relaxed and ov:
                    flags |= klt.MEM_OVERLAP
                if rng.random() < 0.3:
                    flags |= klt.PUSH_EIG
                ctx.push_frames([pool[int(rng.integers(0, 8))].data_ptr() for _ in cams], cams, step, flags, pitch=w)
                for c in cams:
                    held[c] = min(held[c] + 1, slots)
            elif op == 1:                                       # rebuild some slot of some cameras
                cams = sorted(rng.choice(ncam, size=int(rng.integers(1, ncam + 1)), replace=False).tolist())
                slot = int(rng.integers(0, min(held[c] for c in cams)))
                ctx.rebuild_pyramids(np.array(cams, np.int32), slot=slot)
            else:                                               # LK between two slots of one camera
                c = int(rng.integers(0, ncam))
                a, b = rng.choice(held[c], size=2, replace=False).tolist()
                o = dict(out=torch.zeros((n, 2), dtype=torch.float32, device=dev), status=torch.zeros(n, dtype=torch.uint8, device=dev),
                         err=torch.zeros(n, dtype=torch.float32, device=dev), fb=torch.zeros(n, dtype=torch.float32, device=dev))
                outs.append(o)
                ctx.track_batch([dict(cam=c, slot_from=int(a), slot_to=int(b), n_pts=n, pts=d_pts.data_ptr(), out=o["out"].data_ptr(),
                                      status=o["status"].data_ptr(), err=o["err"].data_ptr(), fb_err=o["fb"].data_ptr())], mem=klt.MEM_DEVICE)
            if not relaxed:
                ctx.sync()
        ctx.sync()
        for o in outs:
            res.append(tuple(o[key].cpu().numpy() for key in ("out", "status", "err", "fb")))
        planes = [[ctx.read_level(c, s, l) for l in range(levels)] for c in range(ncam) for s in range(held[c])]
        for d in dets:
            res.append((d["pts"].cpu().numpy(), d["cnt"].cpu().numpy()))
    return res, planes


@pytest.mark.parametrize("mode", ["auto", "level", "hybrid"])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_command_sequences_equal_the_serialised_run(monkeypatch, seed, mode):
    """Stream-ordering fuzz: overlapped pushes, consecutive builds running beside each other (late_wait), programmatically
    launched LK kernels and rebuilds in random order give bit-identical tracks and pyramids to the same sequence run without
    any overlap and with a host sync after every call."""
    strict_tracks, strict_planes = _fuzz_run(monkeypatch, seed, False, mode)
    relaxed_tracks, relaxed_planes = _fuzz_run(monkeypatch, seed, True, mode)
    assert len(strict_tracks) == len(relaxed_tracks) > 5
    for k, (a, b) in enumerate(zip(strict_tracks, relaxed_tracks)):
        for x, y in zip(a, b):
            assert np.array_equal(x, y), ("track", k)
    for k, (pa, pb) in enumerate(zip(strict_planes, relaxed_planes)):
        for l, ((ia, da), (ib, db)) in enumerate(zip(pa, pb)):
            assert np.array_equal(ia, ib) and np.array_equal(da, db), ("plane", k, "level", l)
