"""GPU tier: the device-resident backward chain (psnw_klt_backward_chain, Tracker2D.cpp:763-811) and the optical-flow
majority check (psnw_flow_majority, Tracker2D.cpp:931-944 / :981-1022) against the oracle flow driven step by step."""
import numpy as np
import pytest

from mcmtt_opticalflow_b200 import klt, synth

pytestmark = pytest.mark.gpu


def _oracle_chain(oracle, pyr, box, seeds, steps, min_inl=4):
    """The reference loop :763-811 with the oracle's LK (exact sums) and LocalSearchKLT."""
    boxes, pts_steps, cur, first = [], [], seeds, None
    for k in range(1, steps + 1):
        out, st, er = oracle.lk_track(pyr[k - 1], pyr[k], cur, (int(box[2]), int(box[2])), 3, exact_sums=True)
        nb, inl = oracle.local_search_klt(box, cur, out)              # always the detection box (:770), status ignored (:787)
        if len(inl) < min_inl:
            break
        if first is None:
            first = inl
        boxes.append(nb)
        cur = out[inl]
        pts_steps.append(cur)
    return boxes, pts_steps, first


@pytest.mark.parametrize("frames_held", [4, 2])
def test_backward_chain_equals_oracle_chain(oracle, frames_held):
    w, h, ntgt, cap = 768, 576, 14, 100
    cam = synth.Camera(w, h, ntgt, cam=41, static_frac=0.15, leaving_frac=0.1, max_speed=3.0)
    frames = [cam.frame(t) for t in range(frames_held)]
    boxes = cam.boxes(frames_held - 1)
    win = np.stack([boxes[:, 2].astype(np.int32), boxes[:, 2].astype(np.int32)], 1)
    with klt.KLTContext(w, h, levels=4, fb_thresh=-1.0, apron=64) as ctx:
        for f in frames:
            ctx.push_frame(f)
        seeds, counts = ctx.detect(boxes, cap, 0.01, 3.0)
        bs, sd, ps, cs, fm = ctx.backward_chain(seeds, counts, boxes, win_wh=win, n_steps=3, min_inliers=4)
    pyr = [oracle.build_pyramid(f, 3) for f in frames[::-1]]          # pyr[k] = k frames ago
    steps = frames_held - 1
    ran = 0
    for t in range(ntgt):
        s = seeds[t, :counts[t]]
        if counts[t] == 0:
            assert sd[t] == 0
            continue
        ob, op, first = _oracle_chain(oracle, pyr, boxes[t], s, steps)
        assert sd[t] == len(ob), (t, sd[t], len(ob))
        for k in range(len(ob)):
            assert np.allclose(bs[k, t], ob[k], atol=1e-6), (t, k)
            assert cs[k, t] == len(op[k])
            assert np.abs(ps[k, t, :cs[k, t]] - op[k]).max() <= 1e-4
        for k in range(len(ob), 3):
            assert cs[k, t] == 0 and not bs[k, t].any()
        if first is not None:
            assert np.array_equal(np.flatnonzero(fm[t]), first)
        ran += len(ob)
    assert ran >= ntgt * steps // 2


def test_backward_chain_device_memory_equals_host(oracle):
    torch = pytest.importorskip("torch")
    w, h, ntgt, cap = 384, 288, 6, 64
    cam = synth.Camera(w, h, ntgt, cam=42, static_frac=0.0, leaving_frac=0.0, max_speed=2.0)
    boxes = cam.boxes(3)
    with klt.KLTContext(w, h, levels=3, fb_thresh=-1.0) as ctx:
        for t in range(4):
            ctx.push_frame(cam.frame(t))
        seeds, counts = ctx.detect(boxes, cap, 0.01, 3.0)
        ref = ctx.backward_chain(seeds, counts, boxes, n_steps=3)
        dev = torch.device("cuda", 0)
        d_seeds, d_counts, d_boxes = torch.from_numpy(seeds).to(dev), torch.from_numpy(counts).to(dev), torch.from_numpy(boxes).to(dev)
        d_bs = torch.zeros((3, ntgt, 4), dtype=torch.float64, device=dev)
        d_sd = torch.zeros(ntgt, dtype=torch.int32, device=dev)
        d_ps = torch.zeros((3, ntgt, cap, 2), dtype=torch.float32, device=dev)
        d_cs = torch.zeros((3, ntgt), dtype=torch.int32, device=dev)
        d_fm = torch.zeros((ntgt, cap), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ctx._ck(ctx.lib.psnw_klt_backward_chain(ctx._h, 0, 3, d_seeds.data_ptr(), d_counts.data_ptr(), ntgt, cap, None, d_boxes.data_ptr(), 4,
                                                d_bs.data_ptr(), d_sd.data_ptr(), d_ps.data_ptr(), d_cs.data_ptr(), d_fm.data_ptr(), klt.MEM_DEVICE))
        ctx.sync()
        got = (d_bs.cpu().numpy(), d_sd.cpu().numpy(), d_ps.cpu().numpy(), d_cs.cpu().numpy(), d_fm.cpu().numpy())
    for a, b in zip(ref, got):
        # (slots past a target's count hold the dead-slot sentinel on the device path and zeros in a fresh host array)
        if a.ndim == 4:
            for k in range(3):
                for t in range(ntgt):
                    assert np.array_equal(a[k, t, :ref[3][k, t]], b[k, t, :ref[3][k, t]])
        else:
            assert np.array_equal(a, b)


def test_flow_majority_equals_reference_rule(oracle):
    rng = np.random.default_rng(9)
    with klt.KLTContext(320, 240, levels=2) as ctx:
        for trial in range(30):
            nt, nd = int(rng.integers(1, 12)), int(rng.integers(1, 9))
            sizes = rng.integers(0, 40, size=nt)
            off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
            centres = rng.uniform(40, 200, size=(nt, 2))
            pts = np.concatenate([rng.normal(centres[t], 12.0, size=(sizes[t], 2)) for t in range(nt)] + [np.zeros((0, 2))]).astype(np.float32)
            tb = np.concatenate([centres - 20, np.full((nt, 2), 40.0)], 1)
            det = np.concatenate([rng.uniform(20, 180, size=(nd, 2)), rng.uniform(20, 70, size=(nd, 2))], 1)
            ok = (rng.random(nt) < 0.8).astype(np.uint8)
            c, wipe = ctx.flow_majority(pts, off, tb, ok, det, 0.5)
            gc, gw = oracle.flow_majority(pts, off, tb, ok, det, 0.5)
            assert np.array_equal(c, gc) and np.array_equal(wipe, gw), trial
