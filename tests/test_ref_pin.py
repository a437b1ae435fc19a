"""The pin to the reference's OWN code: oracle/_ref/libpsnw_ref.so is LocalSearchKLT + PSN_Rect compiled unchanged from
/root/reference/psn_where/PSNWhere_Tracker2D.cpp:452-554 and PSNWhere_Types.h:18-182 (oracle/Makefile, oracle/ref_shim/).
CPU tier: the C restatement (oracle/klt_oracle.c) and the committed golden vectors against it.  GPU tier: the CUDA
box_update kernel against it through the C ABI, bit-exact doubles and identical inlier lists."""
import numpy as np
import pytest

from oracle import psnw_ref

needs_ref = pytest.mark.skipif(not psnw_ref.available(), reason="oracle/_ref not built and /root/reference absent")


def random_targets(seed=42, n_t=200):
    """The 200 random targets of test_gpu_box_detect.py::test_box_update_random_vs_oracle (ties, the 0.1-px gate,
    half outliers, empty / single-feature / 256-feature targets)."""
    rng = np.random.default_rng(seed)
    ns = rng.integers(0, 257, size=n_t)
    ns[:3] = [0, 1, 256]
    off = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
    pre = rng.uniform(0, 1900, size=(off[-1], 2)).astype(np.float32)
    cur = pre.copy()
    boxes = np.stack([rng.uniform(0, 1800, n_t), rng.uniform(0, 900, n_t), rng.uniform(5, 140, n_t), rng.uniform(20, 330, n_t)], 1)
    for t in range(n_t):
        s = slice(off[t], off[t + 1])
        kind = t % 4
        d = rng.uniform(-8, 8, size=2)
        if kind == 0:
            cur[s] = pre[s] + d + rng.normal(0, 0.4, size=(ns[t], 2))
        elif kind == 1:
            cur[s] = pre[s] + np.round(rng.uniform(-2, 2, size=(ns[t], 2)) * 4) / 4       # many exact ties
        elif kind == 2:
            cur[s] = pre[s] + rng.normal(0, 0.06, size=(ns[t], 2))                        # around the 0.1 px gate
        else:
            cur[s] = pre[s] + d
            k = ns[t] // 2
            cur[off[t]:off[t] + k] = pre[off[t]:off[t] + k] + rng.uniform(-40, 40, size=(k, 2))
    return boxes, pre, cur.astype(np.float32), off


@needs_ref
def test_golden_vectors_equal_the_reference_code(golden):
    """tests/golden/klt_golden.npz's LocalSearchKLT cases were made by a Python transcription (oracle/gen_golden.py); here the
    reference's compiled code reproduces every one of them bit for bit."""
    off = np.concatenate([[0], np.cumsum(golden["ls_n"])])
    ioff = np.concatenate([[0], np.cumsum(golden["ls_inl_n"])])
    for c in range(len(golden["ls_n"])):
        pre, cur = golden["ls_pre"][off[c]:off[c + 1]], golden["ls_cur"][off[c]:off[c + 1]]
        box, inl = psnw_ref.local_search_klt(golden["ls_box"][c], pre, cur)
        assert np.array_equal(box, golden["ls_box_out"][c]), c
        assert np.array_equal(inl, golden["ls_inl"][ioff[c]:ioff[c + 1]]), c


@needs_ref
def test_c_restatement_equals_the_reference_code(oracle):
    boxes, pre, cur, off = random_targets()
    for t in range(len(boxes)):
        s = slice(off[t], off[t + 1])
        b, inl = oracle.local_search_klt(boxes[t], pre[s], cur[s])
        rb, rinl = psnw_ref.local_search_klt(boxes[t], pre[s], cur[s])
        assert np.array_equal(b, rb), t
        assert np.array_equal(inl, rinl), t
    # forward caller: the status filter of Tracker2D.cpp:883-888 applied by the caller before the call
    rng = np.random.default_rng(5)
    st = (rng.random(off[-1]) < 0.8).astype(np.uint8)
    for t in range(0, len(boxes), 7):
        s = slice(off[t], off[t + 1])
        keep = st[s] == 1
        b, inl = oracle.local_search_klt(boxes[t], pre[s], cur[s], st[s])
        rb, rinl = psnw_ref.local_search_klt(boxes[t], pre[s][keep], cur[s][keep])
        assert np.array_equal(b, rb) and np.array_equal(inl, np.flatnonzero(keep)[rinl]), t


@needs_ref
def test_rect_helpers_equal_the_reference_code(oracle):
    """PSN_Rect::cropWithSize().cv() (mask ROI, Tracker2D.cpp:735), contain (:939-944), overlappedArea (IoU metric)."""
    rng = np.random.default_rng(11)
    for _ in range(300):
        box = np.array([rng.uniform(-60, 800), rng.uniform(-60, 600), rng.uniform(1, 200), rng.uniform(1, 400)])
        x, y, w, h = psnw_ref.crop_roi(box, 768, 576)
        x0, y0, x1, y1 = oracle.crop_roi(box, 768, 576)
        assert (x0, y0) == (x, y) and (x1, y1) == (x + max(w, 0), y + max(h, 0)), box


@pytest.mark.gpu
@needs_ref
def test_cuda_box_update_equals_the_reference_code():
    from mcmtt_opticalflow_b200 import klt
    boxes, pre, cur, off = random_targets()
    with klt.KLTContext(256, 192, levels=3, max_points=1 << 17, max_targets=1 << 10) as ctx:
        g_boxes, g_mask, g_counts = ctx.box_update(boxes, pre, cur, off)
    for t in range(len(boxes)):
        s = slice(off[t], off[t + 1])
        rb, rinl = psnw_ref.local_search_klt(boxes[t], pre[s], cur[s])
        assert np.array_equal(g_boxes[t], rb), t
        assert np.array_equal(np.flatnonzero(g_mask[s]), rinl), t
        assert g_counts[t] == len(rinl)


@pytest.mark.gpu
@needs_ref
def test_cuda_box_update_golden_equals_the_reference_code(golden):
    from mcmtt_opticalflow_b200 import klt
    off = np.concatenate([[0], np.cumsum(golden["ls_n"])]).astype(np.int32)
    with klt.KLTContext(256, 192, levels=3) as ctx:
        g_boxes, g_mask, g_counts = ctx.box_update(golden["ls_box"], golden["ls_pre"], golden["ls_cur"], off)
    for c in range(len(golden["ls_n"])):
        s = slice(off[c], off[c + 1])
        rb, rinl = psnw_ref.local_search_klt(golden["ls_box"][c], golden["ls_pre"][s], golden["ls_cur"][s])
        assert np.array_equal(g_boxes[c], rb) and np.array_equal(np.flatnonzero(g_mask[s]), rinl), c
