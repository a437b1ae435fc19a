// Host-side mirror of CPSNWhere_Tracker2D for the KLT hot path -- see psnw_tracker2d.h.
// All `:line` references are psn_where/PSNWhere_Tracker2D.cpp.  The control flow follows Run (:251-373); the OpenCV
// / LocalSearchKLT call sites are batched over all detections (or trackers) of the frame and sent through the C ABI.
#include "psnw_tracker2d.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace psn {

// parameters (:10-33)
static const unsigned PSN_2D_MAX_TRACKLET_LENGTH = 3;
static const size_t PSN_2D_FEATURE_MIN_NUM_TRACK = 4;
static const int PSN_2D_FEATURE_MAX_NUM_TRACK = 100;
static const int PSN_2D_BACKTRACKING_INTERVAL = 4;
static const double PSN_2D_MAX_HEIGHT = 2300, PSN_2D_MIN_HEIGHT = 1400;
static const double PSN_2D_BOX_MAX_DISTANCE = 1.0;
static const double PSN_2D_MAX_DETECTION_DISTANCE = 600, PSN_2D_MAX_HEIGHT_DIFFERENCE = 400;
static const double PSN_2D_MAX_BOX_CENTER_DIFFERENCE_RATIO = 0.5, PSN_2D_MIN_OVERLAPPED_RATIO = 0.3;
static const double PSN_2D_MIN_OPTICAL_FLOW_MAJORITY_RATIO = 0.5;
// detector parameters of the Shi-Tomasi replacement of GridFAST
static const float PSN_2D_DETECT_QUALITY = 0.01f, PSN_2D_DETECT_MIN_DIST = 3.0f;

// ---- PSN_Rect (psn_where/PSNWhere_Types.h:131-178) ----
bool PSN_Rect::overlap(const PSN_Rect &a) const
{
    return (std::max(x + w, a.x + a.w) - std::min(x, a.x) < w + a.w) && (std::max(y + h, a.y + a.h) - std::min(y, a.y) < h + a.h);
}
double PSN_Rect::distance(const PSN_Rect &a) const
{
    const double dx = (x + w / 2.0) - (a.x + a.w / 2.0), dy = (y + h / 2.0) - (a.y + a.h / 2.0), dw = w - a.w;
    return std::sqrt(dx * dx + dy * dy + dw * dw) / std::min(w, a.w);
}
double PSN_Rect::overlappedArea(const PSN_Rect &a) const
{
    const double ow = std::min(x + w, a.x + a.w) - std::max(x, a.x);
    if (0.0 >= ow) return 0.0;
    const double oh = std::min(y + h, a.y + a.h) - std::max(y, a.y);
    if (0.0 >= oh) return 0.0;
    return ow * oh;
}
void PSN_Rect::center(double &cx, double &cy) const { cx = x + std::ceil(w / 2.0); cy = y + std::ceil(h / 2.0); }

// ---- assignment (replaces helpers/PSNWhere_Hungarian.cpp) : Kuhn-Munkres with potentials, O(n^2 m) ----
void HungarianMatch(const std::vector<float> &cost, int rows, int cols, std::vector<int> &matchRows, std::vector<int> &matchCols)
{
    matchRows.clear(); matchCols.clear();
    if (rows <= 0 || cols <= 0) return;
    const bool transpose = rows > cols;                 // the algorithm needs n <= m
    const int n = transpose ? cols : rows, m = transpose ? rows : cols;
    auto at = [&](int i, int j) -> double { return transpose ? cost[(size_t)j * cols + i] : cost[(size_t)i * cols + j]; };
    const double INF = std::numeric_limits<double>::infinity();
    std::vector<double> u(n + 1, 0.0), v(m + 1, 0.0);
    std::vector<int> p(m + 1, 0), way(m + 1, 0);
    for (int i = 1; i <= n; ++i) {
        p[0] = i;
        int j0 = 0;
        std::vector<double> minv(m + 1, INF);
        std::vector<char> used(m + 1, 0);
        do {
            used[j0] = 1;
            const int i0 = p[j0];
            double delta = INF;
            int j1 = 0;
            for (int j = 1; j <= m; ++j) {
                if (used[j]) continue;
                const double cur = at(i0 - 1, j - 1) - u[i0] - v[j];
                if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
                if (minv[j] < delta) { delta = minv[j]; j1 = j; }
            }
            for (int j = 0; j <= m; ++j) {
                if (used[j]) { u[p[j]] += delta; v[j] -= delta; }
                else minv[j] -= delta;
            }
            j0 = j1;
        } while (p[j0] != 0);
        do { const int j1 = way[j0]; p[j0] = p[j1]; j0 = j1; } while (j0);
    }
    for (int j = 1; j <= m; ++j) {
        if (p[j] == 0) continue;
        const int a = p[j] - 1, b = j - 1;
        matchRows.push_back(transpose ? b : a);
        matchCols.push_back(transpose ? a : b);
    }
}

// ---------------------------------------------------------------------------------------------------------
CPSNWhere_Tracker2D::CPSNWhere_Tracker2D()
    : m_bInit(false), m_nCamID(0), m_nCurrentFrameIdx(0), m_nNewTrackerID(0), m_nInputWidth(0), m_nInputHeight(0),
      m_nKltCam(0), m_klt(nullptr), m_estimator(nullptr), m_estimatorUser(nullptr) {}

CPSNWhere_Tracker2D::~CPSNWhere_Tracker2D() {}

// :99-163 -- the FIFO, the GridFAST detector and the mask image are replaced by the psnw context
void CPSNWhere_Tracker2D::Initialize(unsigned int nCamID, psnw_ctx *ctx, int camIndex, int width, int height,
                                     HeightEstimator estimator, void *estimatorUser)
{
    if (m_bInit) Finalize();
    m_nCamID = nCamID; m_klt = ctx; m_nKltCam = camIndex;
    m_nInputWidth = width; m_nInputHeight = height;
    m_estimator = estimator; m_estimatorUser = estimatorUser;
    m_nCurrentFrameIdx = 0; m_nNewTrackerID = 0;
    m_listTracker2D.clear(); m_queueActiveTracker2D.clear(); m_vecDetection2D.clear();
    m_stTrack2DResult.camID = nCamID; m_stTrack2DResult.frameIdx = 0;
    m_stTrack2DResult.object2DInfos.clear(); m_stTrack2DResult.matMatchingCost.clear();
    m_stTrack2DResult.costRows = m_stTrack2DResult.costCols = 0;
    // The reference has no per-point forward-backward gate (its status is calcOpticalFlowPyrLK's alone, :883-888): this
    // mirror reproduces the reference only on a context created with fb_thresh < 0, and refuses any other.
    psnw_params kp;
    if (!ctx || psnw_get_params(ctx, &kp) != PSNW_OK || kp.fb_thresh >= 0.f) {
        m_strError = "CPSNWhere_Tracker2D::Initialize: the context must be created with fb_thresh < 0 (the reference has no forward-backward gate)";
        m_bInit = false;
        return;
    }
    m_bInit = true;
}

// :176-237
void CPSNWhere_Tracker2D::Finalize()
{
    if (!m_bInit) return;
    m_listTracker2D.clear(); m_queueActiveTracker2D.clear(); m_vecDetection2D.clear();
    m_bInit = false;
}

bool CPSNWhere_Tracker2D::Check(int rc, const char *what)
{
    if (rc == PSNW_OK) return true;
    m_strError = std::string(what) + ": " + psnw_last_error(m_klt);
    return false;
}

// :251-373
stTrack2DResult &CPSNWhere_Tracker2D::Run(std::vector<stDetection> curDetectionResult, const unsigned char *gray, size_t pitch,
                                          unsigned int frameIdx)
{
    m_strError.clear();
    m_nCurrentFrameIdx = frameIdx;
    m_stTrack2DResult.camID = m_nCamID;
    m_stTrack2DResult.frameIdx = frameIdx;
    m_stTrack2DResult.object2DInfos.clear();
    if (!m_bInit) { m_strError = "Run before Initialize"; return m_stTrack2DResult; }        // assert(m_bInit), :253
    // :256-263 + :311-316: the new frame becomes slot 0 of the device ring (pyramid + Scharr built once)
    if (gray && !Check(psnw_push_frame(m_klt, m_nKltCam, gray, pitch, frameIdx, PSNW_MEM_HOST), "psnw_push_frame")) return m_stTrack2DResult;

    Track2D_BackwardFeatureTracking(curDetectionResult);                                     // :294
    std::vector<float> matchingCost = Track2D_ForwardTrackingAndGetMatchingScore();          // :299
    Track2D_MatchingAndUpdating(matchingCost);                                               // :304
    return m_stTrack2DResult;
}

// :690-838, batched: one detect call for all detections, then <= 3 chain steps, each ONE track + ONE box_update call
size_t CPSNWhere_Tracker2D::Track2D_BackwardFeatureTracking(std::vector<stDetection> &curDetectionResult)
{
    m_vecDetection2D.clear();
    m_vecDetection2D.reserve(curDetectionResult.size());
    unsigned int detectionID = 0;
    for (size_t d = 0; d < curDetectionResult.size(); ++d) {
        // detection validation, :711-715
        PSN_Point3D loc = {0.0, 0.0, 0.0};
        const double height = m_estimator ? m_estimator(m_estimatorUser, curDetectionResult[d].box, &loc) : 1700.0;
        if (PSN_2D_MAX_HEIGHT < height || PSN_2D_MIN_HEIGHT > height) continue;
        stDetectedObject cur;
        cur.id = detectionID++;
        cur.detection = curDetectionResult[d];
        cur.bMatchedWithTracker = false;
        cur.bOverlapWithOtherDetection = false;
        cur.boxes.push_back(cur.detection.box);
        cur.location = loc;
        cur.height = height;
        m_vecDetection2D.push_back(cur);
    }
    const int D = (int)m_vecDetection2D.size();
    if (D == 0) return 0;

    // feature extraction, :735-757 (one batched call; the ROI crop of :735 happens inside the library)
    const int K = PSN_2D_FEATURE_MAX_NUM_TRACK;
    std::vector<double> boxes(4 * (size_t)D);
    for (int d = 0; d < D; ++d) {
        const PSN_Rect &b = m_vecDetection2D[d].detection.box;
        boxes[4 * d] = b.x; boxes[4 * d + 1] = b.y; boxes[4 * d + 2] = b.w; boxes[4 * d + 3] = b.h;
    }
    std::vector<float> seeds(2 * (size_t)D * K);
    std::vector<int> counts(D, 0);
    if (!Check(psnw_klt_detect(m_klt, m_nKltCam, 0, boxes.data(), D, K, PSN_2D_DETECT_QUALITY, PSN_2D_DETECT_MIN_DIST,
                               seeds.data(), counts.data(), PSNW_MEM_HOST), "psnw_klt_detect")) { m_vecDetection2D.clear(); return 0; }

    // detections with < 4 keypoints are skipped (:744)
    std::vector<stDetectedObject> kept;
    std::vector<std::vector<Point2f>> curPts;
    for (int d = 0; d < D; ++d) {
        if ((size_t)counts[d] < PSN_2D_FEATURE_MIN_NUM_TRACK) continue;
        const Point2f *p = reinterpret_cast<const Point2f *>(seeds.data()) + (size_t)d * K;
        curPts.emplace_back(p, p + counts[d]);
        kept.push_back(m_vecDetection2D[d]);
    }
    m_vecDetection2D.swap(kept);
    const int N = (int)m_vecDetection2D.size();

    // backward feature tracking, :763-811: every detection, every step, ONE call -- the three dependent steps (LK into the
    // older frame, LocalSearchKLT against the detection box, inliers carried on) stay on the device
    const int held = psnw_frames_held(m_klt, m_nKltCam);
    const int steps = std::min(PSN_2D_BACKTRACKING_INTERVAL - 1, held - 1);
    if (N > 0 && steps > 0) {
        std::vector<float> pts(2 * (size_t)N * K, -1.0e6f);
        std::vector<int> cnt(N), win(2 * (size_t)N), done(N, 0), cnt_steps((size_t)steps * N, 0);
        std::vector<double> bin(4 * (size_t)N), bsteps(4 * (size_t)steps * N, 0.0);
        std::vector<float> pts_steps(2 * (size_t)steps * N * K);
        std::vector<unsigned char> first((size_t)N * K, 0);
        for (int d = 0; d < N; ++d) {
            const PSN_Rect &b = m_vecDetection2D[d].detection.box;
            win[2 * d] = (int)b.w; win[2 * d + 1] = (int)b.w;                                      // cv::Size(w, w), :782
            bin[4 * d] = b.x; bin[4 * d + 1] = b.y; bin[4 * d + 2] = b.w; bin[4 * d + 3] = b.h;    // always the detection box, :770
            cnt[d] = (int)curPts[d].size();
            memcpy(&pts[2 * (size_t)d * K], curPts[d].data(), sizeof(Point2f) * curPts[d].size());
        }
        if (Check(psnw_klt_backward_chain(m_klt, m_nKltCam, steps, pts.data(), cnt.data(), N, K, win.data(), bin.data(),
                                          (int)PSN_2D_FEATURE_MIN_NUM_TRACK, bsteps.data(), done.data(), pts_steps.data(), cnt_steps.data(),
                                          first.data(), PSNW_MEM_HOST), "psnw_klt_backward_chain")) {
            for (int d = 0; d < N; ++d) {
                stDetectedObject &det = m_vecDetection2D[d];
                for (int k = 1; k <= done[d]; ++k) {
                    const double *bo = &bsteps[4 * ((size_t)(k - 1) * N + d)];
                    det.boxes.push_back(PSN_Rect(bo[0], bo[1], bo[2], bo[3]));                     // :789
                    if (k == 1) {                                                                  // :792-800
                        std::vector<Point2f> curInl;
                        for (int j = 0; j < cnt[d]; ++j)
                            if (first[(size_t)d * K + j]) curInl.push_back(curPts[d][j]);
                        det.vecvecTrackedFeatures.push_back(curInl);
                    }
                    const Point2f *pp = reinterpret_cast<const Point2f *>(&pts_steps[2 * (((size_t)(k - 1) * N + d) * K)]);
                    std::vector<Point2f> prevInl(pp, pp + cnt_steps[(size_t)(k - 1) * N + d]);     // :802-807
                    det.vecvecTrackedFeatures.push_back(prevInl);
                    curPts[d].swap(prevInl);
                }
            }
        }
    }
    for (int d = 0; d < N; ++d)                                                                    // :814-817
        if (m_vecDetection2D[d].vecvecTrackedFeatures.empty()) m_vecDetection2D[d].vecvecTrackedFeatures.push_back(curPts[d]);

    // check overlap, :824-835
    for (int a = 0; a < N; ++a) {
        if (m_vecDetection2D[a].bOverlapWithOtherDetection) continue;
        for (int b = a + 1; b < N; ++b)
            if (m_vecDetection2D[a].detection.box.overlap(m_vecDetection2D[b].detection.box)) { m_vecDetection2D[a].bOverlapWithOtherDetection = true; break; }
    }
    return m_vecDetection2D.size();
}

// :851-1025, batched over the active trackers
std::vector<float> CPSNWhere_Tracker2D::Track2D_ForwardTrackingAndGetMatchingScore()
{
    const size_t numDetection = m_vecDetection2D.size(), T = m_queueActiveTracker2D.size();
    std::vector<float> matchingCostArray(numDetection * T, std::numeric_limits<float>::infinity());
    if (T == 0 || psnw_frames_held(m_klt, m_nKltCam) < 2) return matchingCostArray;

    // forward feature tracking of every tracker in one launch: window (w, h) of its last box (:870-877)
    std::vector<Point2f> pts;
    std::vector<int> tgt, win(2 * T), off(T + 1, 0);
    std::vector<double> bin(4 * T);
    for (size_t t = 0; t < T; ++t) {
        const stTracker2D *trk = m_queueActiveTracker2D[t];
        const PSN_Rect &b = trk->boxes.back();
        win[2 * t] = std::max((int)b.w, 1); win[2 * t + 1] = std::max((int)b.h, 1);
        bin[4 * t] = b.x; bin[4 * t + 1] = b.y; bin[4 * t + 2] = b.w; bin[4 * t + 3] = b.h;
        pts.insert(pts.end(), trk->featurePoints.begin(), trk->featurePoints.end());
        tgt.insert(tgt.end(), trk->featurePoints.size(), (int)t);
        off[t + 1] = (int)pts.size();
    }
    const int n = (int)pts.size();
    std::vector<Point2f> out(std::max(n, 1));
    std::vector<unsigned char> st(std::max(n, 1)), inl(std::max(n, 1));
    std::vector<float> err(std::max(n, 1));
    std::vector<double> bout(4 * T);
    std::vector<int> ninl(T);
    if (n > 0) {
        // calcOpticalFlowPyrLK (:871-877) + motion estimation on the status-filtered pairs (:883-893) as one call
        if (!Check(psnw_klt_track_update(m_klt, m_nKltCam, 1, 0, &pts[0].x, off.data(), (int)T, win.data(), bin.data(),
                                         /*use_status=*/1, &out[0].x, st.data(), err.data(), bout.data(), inl.data(), ninl.data(),
                                         PSNW_MEM_HOST), "psnw_klt_track_update(forward)")) return matchingCostArray;
    }

    // points of every tracker inside every detection box + the majority rule (:931-944, :981-1022) on the device: the
    // tracked points of all trackers go up once more as one array, D x T counts and D flags come back
    std::vector<unsigned char> trkOk(T, 0), wipe(numDetection, 0);
    std::vector<int> inBox(numDetection * T, 0);
    for (size_t t = 0; t < T; ++t) {
        size_t okPts = 0;
        for (int j = off[t]; j < off[t + 1]; ++j) okPts += st[j] ? 1 : 0;
        trkOk[t] = okPts >= PSN_2D_FEATURE_MIN_NUM_TRACK ? 1 : 0;                                  // :889
    }
    if (n > 0 && numDetection > 0) {
        std::vector<double> dbox(4 * numDetection);
        for (size_t d = 0; d < numDetection; ++d) {
            const PSN_Rect &b = m_vecDetection2D[d].detection.box;
            dbox[4 * d] = b.x; dbox[4 * d + 1] = b.y; dbox[4 * d + 2] = b.w; dbox[4 * d + 3] = b.h;
        }
        if (!Check(psnw_flow_majority(m_klt, &out[0].x, off.data(), (int)T, bout.data(), trkOk.data(), dbox.data(), (int)numDetection,
                                      PSN_2D_MIN_OPTICAL_FLOW_MAJORITY_RATIO, inBox.data(), wipe.data(), PSNW_MEM_HOST),
                   "psnw_flow_majority")) return matchingCostArray;
    }

    for (size_t t = 0; t < T; ++t) {
        stTracker2D *curTracker = m_queueActiveTracker2D[t];
        std::vector<Point2f> vecPrev, vecCurr;
        for (int j = off[t]; j < off[t + 1]; ++j)
            if (st[j]) { vecPrev.push_back(pts[j]); vecCurr.push_back(out[j]); }
        // trackedPoints holds ALL LK outputs while the detections are compared (:874, :940-944)
        curTracker->trackedPoints.assign(out.begin() + off[t], out.begin() + off[t + 1]);
        if (PSN_2D_FEATURE_MIN_NUM_TRACK > vecCurr.size()) continue;                               // :889
        const PSN_Rect newBox(bout[4 * t], bout[4 * t + 1], bout[4 * t + 2], bout[4 * t + 3]);
        curTracker->boxes.push_back(newBox);                                                       // :895
        curTracker->heads.push_back(curTracker->heads.back());

        // compare with the detections' backtracking result, :931-974
        for (size_t d = 0; d < numDetection; ++d) {
            const stDetectedObject &det = m_vecDetection2D[d];
            const size_t costPos = d * T + t;
            if (!newBox.overlap(det.detection.box)) continue;                                      // (the point counts of :940-944: inBox)
            double boxCost = 0.0;
            const size_t len = std::min((size_t)PSN_2D_BACKTRACKING_INTERVAL, std::min(curTracker->boxes.size(), det.boxes.size()));
            size_t tb = (size_t)curTracker->duration;                                              // duration = # of boxes - 1
            for (size_t b = 0; b < len; ++b, --tb) {
                if (tb >= curTracker->boxes.size()) { boxCost = std::numeric_limits<double>::infinity(); break; }
                const PSN_Rect &db = det.boxes[b], &tbx = curTracker->boxes[tb];
                double dcx, dcy, tcx, tcy;
                db.center(dcx, dcy); tbx.center(tcx, tcy);
                if (!db.overlap(tbx) || PSN_2D_BOX_MAX_DISTANCE < db.distance(tbx) ||
                    PSN_2D_MIN_OVERLAPPED_RATIO > db.overlappedArea(tbx) / std::min(db.area(), tbx.area()) ||
                    PSN_2D_MAX_BOX_CENTER_DIFFERENCE_RATIO * std::max(db.w, tbx.w) < std::sqrt((dcx - tcx) * (dcx - tcx) + (dcy - tcy) * (dcy - tcy))) {
                    boxCost = std::numeric_limits<double>::infinity();
                    break;
                }
                boxCost += BoxMatchingCost(tbx, db);
            }
            if (std::numeric_limits<double>::infinity() == boxCost) continue;
            matchingCostArray[costPos] = (float)(boxCost / (double)len);
        }
        curTracker->featurePoints = vecPrev;                                                       // :977-978
        curTracker->trackedPoints = vecCurr;
    }

    // feature point majority check, :981-1022 (decided on the device above)
    for (size_t d = 0; d < numDetection; ++d) {
        if (!wipe[d]) continue;
        for (size_t t = 0; t < T; ++t) matchingCostArray[d * T + t] = std::numeric_limits<float>::infinity();
    }
    return matchingCostArray;
}

// :615-622
double CPSNWhere_Tracker2D::BoxMatchingCost(const PSN_Rect &box1, const PSN_Rect &box2)
{
    double ax, ay, bx, by;
    box1.center(ax, ay); box2.center(bx, by);
    const double nominator = std::sqrt((ax - bx) * (ax - bx) + (ay - by) * (ay - by)), denominator = (box1.w + box2.w) / 2.0;
    return (nominator * nominator) / (denominator * denominator);
}

// :1038-1182
void CPSNWhere_Tracker2D::Track2D_MatchingAndUpdating(std::vector<float> &matchingCostArray)
{
    float maxCost = -1000.0f;                                                                      // handling infinity, :1041-1052
    for (float c : matchingCostArray) if (std::isfinite(c) && maxCost < c) maxCost = c;
    maxCost += 100.0f;
    for (float &c : matchingCostArray) if (!std::isfinite(c)) c = maxCost;

    m_stTrack2DResult.object2DInfos.clear();
    std::deque<stTracker2D *> queueNewActiveTrackers;
    const int numDetection = (int)m_vecDetection2D.size(), T = (int)m_queueActiveTracker2D.size();
    std::vector<int> rows, cols;
    HungarianMatch(matchingCostArray, numDetection, T, rows, cols);
    for (size_t m = 0; m < rows.size(); ++m) {
        if (maxCost == matchingCostArray[(size_t)rows[m] * T + cols[m]]) continue;                 // :1063
        stDetectedObject *curDetection = &m_vecDetection2D[rows[m]];
        stTracker2D *curTracker = m_queueActiveTracker2D[cols[m]];
        // matching validation, :1071-1077
        const double dx = curDetection->location.x - curTracker->lastPosition.x, dy = curDetection->location.y - curTracker->lastPosition.y,
                     dz = curDetection->location.z - curTracker->lastPosition.z;
        if (std::sqrt(dx * dx + dy * dy + dz * dz) > PSN_2D_MAX_DETECTION_DISTANCE) continue;
        if (std::abs(curDetection->height - curTracker->height) > PSN_2D_MAX_HEIGHT_DIFFERENCE) continue;
        if (curTracker->duration > PSN_2D_MAX_TRACKLET_LENGTH) continue;
        // tracker update, :1082-1094
        curDetection->bMatchedWithTracker = true;
        curTracker->timeEnd = curTracker->timeLastUpdate = m_nCurrentFrameIdx;
        curTracker->duration = curTracker->timeEnd - curTracker->timeStart + 1;
        curTracker->numStatic = 0;
        curTracker->boxes.back() = curDetection->detection.box;
        curTracker->heads.back() = curDetection->detection.vecPartBoxes.empty() ? PSN_Rect() : curDetection->detection.vecPartBoxes.front();
        curTracker->confidence = 1.0;
        curTracker->lastPosition = curDetection->location;
        curTracker->height = curDetection->height;
        queueNewActiveTrackers.push_back(curTracker);
        stObject2DInfo info;                                                                       // result packaging, :1099-1101
        ResultWithTracker(curTracker, info);
        m_stTrack2DResult.object2DInfos.push_back(info);
        curTracker->featurePoints = curDetection->vecvecTrackedFeatures.front();                   // :1104-1105
        curTracker->trackedPoints.clear();
    }
    // tracker generation, :1112-1147
    for (stDetectedObject &det : m_vecDetection2D) {
        if (det.bMatchedWithTracker) continue;
        stTracker2D nt;
        nt.id = m_nNewTrackerID++;
        nt.timeStart = nt.timeEnd = nt.timeLastUpdate = m_nCurrentFrameIdx;
        nt.duration = 1; nt.numStatic = 0;
        nt.boxes.push_back(det.detection.box);
        nt.heads.push_back(det.detection.vecPartBoxes.empty() ? PSN_Rect() : det.detection.vecPartBoxes.front());
        nt.featurePoints = det.vecvecTrackedFeatures.front();
        nt.confidence = 1.0;
        nt.lastPosition = det.location;
        nt.height = det.height;
        m_listTracker2D.push_back(nt);
        queueNewActiveTrackers.push_back(&m_listTracker2D.back());
        stObject2DInfo info;
        ResultWithTracker(&nt, info);
        m_stTrack2DResult.object2DInfos.push_back(info);
    }
    // tracker termination, :1152-1164
    for (stTracker2D *trk : m_queueActiveTracker2D) {
        if (trk->timeLastUpdate == m_nCurrentFrameIdx) continue;
        trk->boxes.clear(); trk->heads.clear(); trk->featurePoints.clear(); trk->trackedPoints.clear();
    }
    m_queueActiveTracker2D = queueNewActiveTrackers;
    m_stTrack2DResult.matMatchingCost = matchingCostArray;          // the reference keeps a D x 0 matrix here (:1173); kept D x T for inspection
    m_stTrack2DResult.costRows = numDetection;
    m_stTrack2DResult.costCols = T;
}

// :1231-1257 (PSN_2D_DEBUG_DISPLAY_SCALE = 1)
void CPSNWhere_Tracker2D::ResultWithTracker(const stTracker2D *curTracker, stObject2DInfo &out)
{
    out.featurePointsPrev = curTracker->featurePoints;
    out.featurePointsCurr = curTracker->trackedPoints;
    out.id = curTracker->id;
    out.box = curTracker->boxes.back();
    out.head = curTracker->heads.back();
    out.score = 0;
}

}  // namespace psn

// ---------------------------------------------------------------------------------------------------------
// Flat C entry points so that the Python tests can drive the C++ class (not part of the product ABI).
// ---------------------------------------------------------------------------------------------------------
extern "C" {
#define PSNT_API __attribute__((visibility("default")))

PSNT_API void *psnt_create(psnw_ctx *ctx, unsigned cam_id, int cam_index, int width, int height)
{
    psn::CPSNWhere_Tracker2D *t = new psn::CPSNWhere_Tracker2D();
    t->Initialize(cam_id, ctx, cam_index, width, height);
    return t;
}

PSNT_API void psnt_destroy(void *h)
{
    psn::CPSNWhere_Tracker2D *t = static_cast<psn::CPSNWhere_Tracker2D *>(h);
    t->Finalize();
    delete t;
}

// boxes: [n][4] detection boxes, heads: [n][4] or NULL.  Returns the number of objects, or -1 on error.
PSNT_API int psnt_run(void *h, const double *boxes, const double *heads, int n, const unsigned char *gray, size_t pitch, unsigned frame_idx)
{
    psn::CPSNWhere_Tracker2D *t = static_cast<psn::CPSNWhere_Tracker2D *>(h);
    std::vector<psn::stDetection> dets(n);
    for (int i = 0; i < n; ++i) {
        dets[i].box = psn::PSN_Rect(boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]);
        if (heads) dets[i].vecPartBoxes.push_back(psn::PSN_Rect(heads[4 * i], heads[4 * i + 1], heads[4 * i + 2], heads[4 * i + 3]));
    }
    psn::stTrack2DResult &r = t->Run(dets, gray, pitch, frame_idx);
    if (!t->LastError().empty()) return -1;
    return (int)r.object2DInfos.size();
}

PSNT_API const char *psnt_last_error(void *h) { return static_cast<psn::CPSNWhere_Tracker2D *>(h)->LastError().c_str(); }

// objects of the last Run: ids[cap], boxes[cap][4], n_prev[cap] (feature counts); returns the number of objects
PSNT_API int psnt_get_objects(void *h, unsigned *ids, double *boxes, int *n_prev, int cap)
{
    const psn::stTrack2DResult &r = static_cast<psn::CPSNWhere_Tracker2D *>(h)->Result();
    const int n = (int)r.object2DInfos.size();
    for (int i = 0; i < n && i < cap; ++i) {
        const psn::stObject2DInfo &o = r.object2DInfos[i];
        ids[i] = o.id;
        boxes[4 * i] = o.box.x; boxes[4 * i + 1] = o.box.y; boxes[4 * i + 2] = o.box.w; boxes[4 * i + 3] = o.box.h;
        n_prev[i] = (int)o.featurePointsPrev.size();
    }
    return n;
}

// backward-chain boxes of the detections kept by the last Run (current -> past): n_boxes[cap], boxes[cap][4][4]
PSNT_API int psnt_get_detection_boxes(void *h, int *n_boxes, double *boxes, int cap)
{
    const std::vector<psn::stDetectedObject> &d = static_cast<psn::CPSNWhere_Tracker2D *>(h)->Detections();
    const int n = (int)d.size();
    for (int i = 0; i < n && i < cap; ++i) {
        n_boxes[i] = (int)d[i].boxes.size();
        for (int k = 0; k < n_boxes[i] && k < 4; ++k) {
            const psn::PSN_Rect &b = d[i].boxes[k];
            double *o = boxes + ((size_t)i * 4 + k) * 4;
            o[0] = b.x; o[1] = b.y; o[2] = b.w; o[3] = b.h;
        }
    }
    return n;
}

// assignment helper, exposed for the CPU-side unit test
PSNT_API int psnt_hungarian(const float *cost, int rows, int cols, int *match_rows, int *match_cols)
{
    std::vector<float> c(cost, cost + (size_t)rows * cols);
    std::vector<int> r, k;
    psn::HungarianMatch(c, rows, cols, r, k);
    for (size_t i = 0; i < r.size(); ++i) { match_rows[i] = r[i]; match_cols[i] = k[i]; }
    return (int)r.size();
}
}
