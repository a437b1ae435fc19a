"""GPU tier: the Tracker2D numeric chain end to end -- detect -> backward LK chain with box motion
(Tracker2D.cpp:735-811) and forward tracking (:860-896) -- against the same flow on the oracle, per-target IoU."""
import numpy as np
import pytest

from mcmtt_opticalflow_b200 import klt, synth

pytestmark = pytest.mark.gpu
MIN_IOU = 0.99


def iou(a, b):
    ow = min(a[0] + a[2], b[0] + b[2]) - max(a[0], b[0])
    oh = min(a[1] + a[3], b[1] + b[3]) - max(a[1], b[1])
    if ow <= 0 or oh <= 0:
        return 0.0
    return ow * oh / (a[2] * a[3] + b[2] * b[3] - ow * oh)


def oracle_backward_chain(oracle, frames, box, seeds, levels):
    """Tracker2D.cpp:763-811 for one detection: frames[0] newest."""
    boxes, cur_pts = [np.array(box)], seeds
    pyr = [oracle.build_pyramid(f, levels - 1) for f in frames]
    win = (int(box[2]), int(box[2]))
    for k in range(1, len(frames)):
        out, st, err = oracle.lk_track(pyr[k - 1], pyr[k], cur_pts, win, levels - 1, exact_sums=True)
        nb, inl = oracle.local_search_klt(box, cur_pts, out)          # status ignored (:787); box = detection box (:770)
        if len(inl) < 4:
            break
        boxes.append(nb)
        cur_pts = out[inl]
    return boxes


def test_backward_chain_and_forward_iou(oracle):
    w, h, ntgt, levels = 768, 576, 20, 4
    cam = synth.Camera(w, h, ntgt, cam=6)
    frames = [cam.frame(t) for t in range(4)]                         # oldest .. newest
    det = cam.boxes(3)
    with klt.KLTContext(w, h, levels=levels, fb_thresh=-1.0) as c:
        for f in frames:
            c.push_frame(f)
        seeds, counts = c.detect(det, 100, 0.01, 3.0)
        # --- backward chain, all detections batched per step ---
        alive = counts >= 4
        cur = [seeds[t, :counts[t]].copy() for t in range(ntgt)]
        gboxes = [[det[t].copy()] for t in range(ntgt)]
        win = np.stack([det[:, 2].astype(np.int32), det[:, 2].astype(np.int32)], 1)
        for k in range(1, 4):
            idx = [t for t in range(ntgt) if alive[t]]
            if not idx:
                break
            pts = np.concatenate([cur[t] for t in idx])
            tid = np.concatenate([np.full(len(cur[t]), i, np.int32) for i, t in enumerate(idx)])
            out, st, err, _ = c.track(pts, slot_from=k - 1, slot_to=k, target_of_pt=tid, win_wh=win[idx])
            off = np.concatenate([[0], np.cumsum([len(cur[t]) for t in idx])]).astype(np.int32)
            nb, mask, cnt = c.box_update(det[idx], pts, out, off)
            for i, t in enumerate(idx):
                if cnt[i] < 4:
                    alive[t] = False
                    continue
                gboxes[t].append(nb[i])
                cur[t] = out[off[i]:off[i + 1]][mask[off[i]:off[i + 1]] == 1]
        # --- oracle flow on the same seeds ---
        worst = 1.0
        for t in range(ntgt):
            if counts[t] < 4:
                continue
            ob = oracle_backward_chain(oracle, frames[::-1], det[t], seeds[t, :counts[t]], levels)
            assert len(ob) == len(gboxes[t]), t
            for a, b in zip(ob, gboxes[t]):
                worst = min(worst, iou(a, b))
        assert worst >= MIN_IOU
        # --- forward tracking of live trackers: slot 1 -> slot 0, window (w, h), status honoured (:871-893) ---
        prev_boxes = cam.boxes(2)
        pts, tid = synth.seeds_in_boxes(prev_boxes, 100, w, h, np.random.default_rng(8))
        fwin = np.stack([prev_boxes[:, 2].astype(np.int32), prev_boxes[:, 3].astype(np.int32)], 1)
        out, st, err, _ = c.track(pts, slot_from=1, slot_to=0, target_of_pt=tid, win_wh=fwin)
        off = (np.arange(ntgt + 1) * 100).astype(np.int32)
        nb, mask, cnt = c.box_update(prev_boxes, pts, out, off, status=st)
        pa, pb = oracle.build_pyramid(frames[2], levels - 1), oracle.build_pyramid(frames[3], levels - 1)
        moved = 0
        for t in range(ntgt):
            m = tid == t
            g_o, g_st, _ = oracle.lk_track(pa, pb, pts[m], (int(fwin[t, 0]), int(fwin[t, 1])), levels - 1, exact_sums=True)
            ob, inl = oracle.local_search_klt(prev_boxes[t], pts[m], g_o, g_st)
            assert iou(ob, nb[t]) >= MIN_IOU, t
            moved += int(np.abs(ob[:2] - prev_boxes[t][:2]).max() > 0.5)
        assert moved >= ntgt // 2                                      # the estimator really moved the boxes


@pytest.mark.parametrize("use_status", [False, True])
def test_track_update_equals_the_two_calls(use_status):
    """psnw_klt_track_update (one upload, two launches, one download) == psnw_klt_track followed by psnw_box_update,
    bit for bit, with the reference's per-target windows (generic kernel) and with the default 21x21 window."""
    w, h, ntgt = 768, 576, 14
    cam = synth.Camera(w, h, ntgt, cam=6)
    f0, f1 = cam.frame(0), cam.frame(1)
    boxes = cam.boxes(0)
    pts, tid = synth.seeds_in_boxes(boxes, 60, w, h, np.random.default_rng(11))
    off = (np.arange(ntgt + 1) * 60).astype(np.int32)
    wins = np.stack([boxes[:, 2].astype(np.int32), boxes[:, 3].astype(np.int32)], 1)
    with klt.KLTContext(w, h, levels=4, fb_thresh=-1.0) as ctx:
        ctx.push_frame(f0); ctx.push_frame(f1)
        for win in (wins, None):
            o1, s1, e1, _ = ctx.track(pts, target_of_pt=tid, win_wh=win if win is not None else [[21, 21]] * ntgt)
            b1, m1, c1 = ctx.box_update(boxes, pts, o1, off, status=s1 if use_status else None)
            o2, s2, e2, b2, m2, c2 = ctx.track_update(pts, off, boxes, win_wh=win, use_status=use_status)
            assert np.array_equal(o1, o2) and np.array_equal(s1, s2) and np.array_equal(e1, e2)
            assert np.array_equal(b1, b2) and np.array_equal(m1, m2) and np.array_equal(c1, c2)
