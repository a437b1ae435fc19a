"""The C++ host mirror of CPSNWhere_Tracker2D (mcmtt_opticalflow_b200/host): assignment helper on CPU, the full
Run() flow on the GPU."""
import numpy as np
import pytest

from integration import tracker2d


def test_hungarian_matches_scipy():
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(0)
    for (r, c) in [(1, 1), (3, 3), (5, 8), (9, 4), (20, 20), (1, 6), (7, 1)]:
        for _ in range(5):
            cost = rng.uniform(0, 10, size=(r, c)).astype(np.float32)
            mr, mc = tracker2d.hungarian(cost)
            assert len(mr) == min(r, c) and len(set(mr)) == len(mr) and len(set(mc)) == len(mc)
            sr, sc = linear_sum_assignment(cost.astype(np.float64))
            assert abs(cost[mr, mc].astype(np.float64).sum() - cost[sr, sc].astype(np.float64).sum()) < 1e-4


@pytest.mark.gpu
def test_run_keeps_ids_and_matches_oracle_chain(oracle):
    from mcmtt_opticalflow_b200 import klt, synth
    w, h, ntgt = 768, 576, 12
    cam = synth.Camera(w, h, ntgt, cam=31, static_frac=0.0, leaving_frac=0.0, max_speed=3.0)
    with klt.KLTContext(w, h, levels=4, fb_thresh=-1.0) as ctx:
        trk = tracker2d.Tracker2D(ctx, 0, cam_id=7)
        hist, frames = [], []
        for t in range(6):
            f = cam.frame(t)
            frames.append(f)
            boxes = cam.boxes(t)
            heads = boxes.copy(); heads[:, 3] *= 0.2
            ids, ob, nf = trk.run(boxes, f, t, heads=heads)
            hist.append((ids, ob, nf))
            assert len(ids) == ntgt and (nf >= 4).all()
            # every reported box is the detection box it was matched to / created from (Tracker2D.cpp:1087,:1125)
            for b in ob:
                assert np.abs(boxes - b).sum(1).min() < 1e-9
            if t == 3:       # ring full: the backward chain has 3 steps -- compare with the oracle chain
                det_boxes = trk.detection_boxes()
                pyr = [oracle.build_pyramid(fr, 3) for fr in frames[::-1]]
                eig = ctx.read_min_eig(0, 0)
                checked = 0
                for d, box in enumerate(boxes):
                    seeds = oracle.detect(f, box, 100, 0.01, 3.0, eig=eig)
                    if len(seeds) < 4:
                        continue
                    ob_chain, cur = [np.array(box)], seeds
                    for k in range(1, 4):
                        out, st, er = oracle.lk_track(pyr[k - 1], pyr[k], cur, (int(box[2]), int(box[2])), 3, exact_sums=True)
                        nb, inl = oracle.local_search_klt(box, cur, out)
                        if len(inl) < 4:
                            break
                        ob_chain.append(nb); cur = out[inl]
                    got = det_boxes[checked]
                    assert len(got) == len(ob_chain)
                    assert np.allclose(got, np.array(ob_chain), atol=1e-6)
                    checked += 1
                assert checked == len(det_boxes)
        trk.close()
    # frame 0 creates ids 0..n-1; the same tracker follows each target while duration <= 3 (frames 1..3) and is
    # re-created afterwards (Tracker2D.cpp:1077): ids persist for four frames
    def by_box(t):
        ids, ob, _ = hist[t]
        gt = cam.boxes(t)
        return {int(np.abs(gt - b).sum(1).argmin()): int(i) for i, b in zip(ids, ob)}
    first = by_box(0)
    assert sorted(first.values()) == list(range(ntgt))

    def isolated(k):      # targets that never overlap another one: occlusions legitimately break matches (:937-1022)
        for t in range(5):
            b = cam.boxes(t)
            for j in range(ntgt):
                if j != k and (max(b[k, 0] + b[k, 2], b[j, 0] + b[j, 2]) - min(b[k, 0], b[j, 0]) < b[k, 2] + b[j, 2]) and \
                        (max(b[k, 1] + b[k, 3], b[j, 1] + b[j, 3]) - min(b[k, 1], b[j, 1]) < b[k, 3] + b[j, 3]):
                    return False
        return True
    iso = [k for k in range(ntgt) if isolated(k)]
    assert len(iso) >= ntgt // 3
    for t in (1, 2, 3):
        cur = by_box(t)
        assert all(cur[k] == first[k] for k in iso), (t, cur, first)
    later = by_box(4)
    assert all(later[k] != first[k] for k in first) and min(later.values()) >= ntgt


def test_detection_file_reader(tmp_path):
    """PETS.S2.L1 full-body detection dump, the format psn::ReadDetectionResultWithTxt parses (Utils.cpp:1021-1080)."""
    parts = ["HEAD", "F1", "S1", "GR", "S2", "A1", "A2", "F2"]
    rng = np.random.default_rng(5)
    boxes = rng.uniform(0, 500, size=(3, 4)).round(3)
    pb = rng.uniform(0, 500, size=(3, 8, 4)).round(3)
    txt = "numBoxes:3\n"
    for d in range(3):
        txt += "{\n\tROOT:{%f,%f,%f,%f}\n" % tuple(boxes[d])
        for p, name in enumerate(parts):
            txt += "\t%s:{%f,%f,%f,%f}\n" % ((name,) + tuple(pb[d, p]))
        txt += "}\n"
    f = tmp_path / "frame_0001.txt"
    f.write_text(txt)
    b, hd = tracker2d.read_detections(str(f))
    assert np.allclose(b, boxes, atol=1e-3) and np.allclose(hd, pb[:, 0], atol=1e-3)
    b, hd = tracker2d.read_detections(str(tmp_path / "missing.txt"))
    assert b.shape == (0, 4)


def test_result_file_roundtrip(tmp_path):
    """track2D_result_cam%d_frame%04d.txt: FilePrintResult (Tracker2D.cpp:1268-1342) -> Read2DTrackResultWithTxt (Utils.cpp:1148-1240)."""
    import ctypes as C
    lib = tracker2d.load()
    rng = np.random.default_rng(6)
    n, k = 5, 7
    ids = np.arange(10, 10 + n, dtype=np.uint32)
    boxes = rng.uniform(0, 700, size=(n, 4))
    pts = rng.uniform(0, 700, size=(n, k, 2)).astype(np.float32)
    ids_o, boxes_o, pts_o = np.zeros(n, np.uint32), np.zeros((n, 4)), np.zeros((n, k, 2), np.float32)
    path = str(tmp_path / "track2D_result_cam1_frame0042.txt")
    got = lib.psnt_result_roundtrip(path.encode(), 1, 42, n, ids.ctypes.data, boxes.ctypes.data, pts.ctypes.data, k,
                                    ids_o.ctypes.data, boxes_o.ctypes.data, pts_o.ctypes.data)
    assert got == n and np.array_equal(ids_o, ids)
    assert np.allclose(boxes_o, boxes, atol=1e-5) and np.allclose(pts_o, pts, atol=1e-4)        # %f keeps 6 decimals
    head = open(path).read().splitlines()[:3]
    assert head == ["camIdx:1", "frameIdx:42", "numObjectInfos:5{"]
