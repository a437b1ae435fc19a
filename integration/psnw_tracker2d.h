// Host-side mirror of CPSNWhere_Tracker2D (psn_where/PSNWhere_Tracker2D.h:132-216) for the KLT hot path: same
// Initialize / Run / Finalize interface, same result types, with the three numeric callees replaced by the C ABI
// of include/psnw_klt.h.  The reference's toolchain (MSVC + MFC + OpenCV 2.4.6) is absent, so this file restates
// the class portably: cv::Mat -> raw 8-bit gray pointer, cv::Point2f -> psn::Point2f, camera calibration -> an
// optional height/location callback (calibration is out of scope).
//
// What stays on the CPU here exactly as in the reference (scope row 1c): detection gating, cost matrix from the
// backward/forward box histories (:931-974), optical-flow majority check (:981-1022), assignment, tracker
// create / update / terminate (:1038-1164), result packaging (:1231-1257).
#pragma once
#include <deque>
#include <list>
#include <string>
#include <vector>

#include "psnw_klt.h"

namespace psn {

struct Point2f { float x, y; };
struct PSN_Point3D { double x, y, z; };

// psn_where/PSNWhere_Types.h:112-182
struct PSN_Rect {
    double x, y, w, h;
    PSN_Rect() : x(0), y(0), w(0), h(0) {}
    PSN_Rect(double x_, double y_, double w_, double h_) : x(x_), y(y_), w(w_), h(h_) {}
    double area() const { return w * h; }
    bool contain(const Point2f &a) const { return (double)a.x >= x && (double)a.x < x + w && (double)a.y >= y && (double)a.y < y + h; }
    bool overlap(const PSN_Rect &a) const;
    double distance(const PSN_Rect &a) const;
    double overlappedArea(const PSN_Rect &a) const;
    void center(double &cx, double &cy) const;
};

struct stDetection {                         // PSNWhere_Types.h:184-188
    PSN_Rect box;
    std::vector<PSN_Rect> vecPartBoxes;
};

struct stObject2DInfo {                      // PSNWhere_Types.h:190-198
    unsigned int id;
    PSN_Rect box, head;
    double score;
    std::vector<Point2f> featurePointsPrev, featurePointsCurr;
};

struct stTrack2DResult {                     // PSNWhere_Types.h:200-209 (matMatchingCost as a row-major float vector)
    unsigned int camID, frameIdx;
    std::vector<stObject2DInfo> object2DInfos;
    std::vector<float> matMatchingCost;
    int costRows, costCols;
};

struct stDetectedObject {                    // PSNWhere_Tracker2D.h:95-107
    unsigned int id;
    stDetection detection;
    bool bMatchedWithTracker, bOverlapWithOtherDetection;
    std::vector<std::vector<Point2f>> vecvecTrackedFeatures;   // current -> past order
    std::vector<PSN_Rect> boxes;                               // current -> past order
    PSN_Point3D location;
    double height;
};

struct stTracker2D {                         // PSNWhere_Tracker2D.h:109-126
    unsigned int id, timeStart, timeEnd, timeLastUpdate, duration, numStatic;
    double confidence;
    std::deque<PSN_Rect> boxes, heads;
    std::vector<Point2f> featurePoints, trackedPoints;
    PSN_Point3D lastPosition;
    double height;
};

// Replaces EstimateDetectionHeight (:1195-1220), which needs the Tsai camera model: returns the height in mm and the
// ground location of a detection.  NULL -> every detection passes the 1400..2300 mm gate (:715) with height 1700.
typedef double (*HeightEstimator)(void *user, const PSN_Rect &box, PSN_Point3D *location);

class CPSNWhere_Tracker2D {
public:
    CPSNWhere_Tracker2D();
    ~CPSNWhere_Tracker2D();
    // `ctx` may be shared by the trackers of all cameras that live on the same GPU; camIndex = index inside it
    void Initialize(unsigned int nCamID, psnw_ctx *ctx, int camIndex, int width, int height,
                    HeightEstimator estimator = nullptr, void *estimatorUser = nullptr);
    void Finalize();
    // gray: 8-bit single channel, `pitch` bytes per row (cvtColor/resize of :256-263 already applied, or use
    // psnw_push_frames_bgr beforehand and pass gray = NULL to skip the upload)
    stTrack2DResult &Run(std::vector<stDetection> curDetectionResult, const unsigned char *gray, size_t pitch, unsigned int frameIdx);

    const std::vector<stDetectedObject> &Detections() const { return m_vecDetection2D; }
    const stTrack2DResult &Result() const { return m_stTrack2DResult; }
    const std::string &LastError() const { return m_strError; }

private:
    size_t Track2D_BackwardFeatureTracking(std::vector<stDetection> &curDetectionResult);
    std::vector<float> Track2D_ForwardTrackingAndGetMatchingScore();
    void Track2D_MatchingAndUpdating(std::vector<float> &matchingCostArray);
    static double BoxMatchingCost(const PSN_Rect &box1, const PSN_Rect &box2);
    static void ResultWithTracker(const stTracker2D *curTracker, stObject2DInfo &outObjectInfo);
    bool Check(int rc, const char *what);

    bool m_bInit;
    unsigned int m_nCamID, m_nCurrentFrameIdx, m_nNewTrackerID;
    int m_nInputWidth, m_nInputHeight, m_nKltCam;
    psnw_ctx *m_klt;
    HeightEstimator m_estimator;
    void *m_estimatorUser;
    std::list<stTracker2D> m_listTracker2D;
    std::deque<stTracker2D *> m_queueActiveTracker2D;
    std::vector<stDetectedObject> m_vecDetection2D;
    stTrack2DResult m_stTrack2DResult;
    std::string m_strError;
};

// Minimum-cost assignment on a rows x cols float matrix (replaces CPSNWhere_Hungarian::Match,
// psn_where/helpers/PSNWhere_Hungarian.cpp:212-359): returns matched (row, col) pairs, min(rows, cols) of them.
void HungarianMatch(const std::vector<float> &cost, int rows, int cols, std::vector<int> &matchRows, std::vector<int> &matchCols);

// on-disk formats (psnw_formats.cpp)
std::vector<stDetection> ReadDetectionResultWithTxt(const std::string &path);
bool FilePrintResult(const std::string &path, const stTrack2DResult &result);
bool Read2DTrackResultWithTxt(const std::string &path, stTrack2DResult &result);

}  // namespace psn
