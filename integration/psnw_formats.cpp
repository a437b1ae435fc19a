// On-disk formats around the hot path ("next" row f-4), portable restatements of
//   psn::ReadDetectionResultWithTxt   psn_where/PSNWhere_Utils.cpp:996-1088  (PETS.S2.L1 full-body: "numBoxes:%d", ROOT + 8 parts)
//   CPSNWhere_Tracker2D::FilePrintResult   psn_where/PSNWhere_Tracker2D.cpp:1268-1342
//   psn::Read2DTrackResultWithTxt     psn_where/PSNWhere_Utils.cpp:1148-1240
// so that the drop-in class can be A/B-tested on real detection dumps and its results diffed against the reference's
// track2D_result_cam%d_frame%04d.txt files.  Same field order, separators and %f formatting as the reference.
#include <cstdio>
#include <cstring>
#include <string>

#include "psnw_tracker2d.h"

namespace psn {

static const int NUM_DETECTION_PART = 8;                                              // PSNWhere_Defines.h:61-62
static const char *DETCTION_PART_NAME[NUM_DETECTION_PART] = {"HEAD", "F1", "S1", "GR", "S2", "A1", "A2", "F2"};

// PETS.S2.L1 full-body detection file (PSN_DETECTION_TYPE 1, PSN_INPUT_TYPE 1); a missing file yields an empty vector (:1034)
std::vector<stDetection> ReadDetectionResultWithTxt(const std::string &path)
{
    std::vector<stDetection> result;
    FILE *fid = fopen(path.c_str(), "r");
    if (!fid) return result;
    int n = 0;
    float x, y, w, h;
    if (fscanf(fid, "numBoxes:%d\n", &n) != 1) { fclose(fid); return result; }
    for (int d = 0; d < n; ++d) {
        if (fscanf(fid, " { ROOT:{%f,%f,%f,%f}", &x, &y, &w, &h) != 4) break;
        stDetection det;
        det.box = PSN_Rect(x, y, w, h);
        bool ok = true;
        for (int p = 0; p < NUM_DETECTION_PART && ok; ++p) {
            char fmt[48];
            snprintf(fmt, sizeof fmt, " %s:{%%f,%%f,%%f,%%f}", DETCTION_PART_NAME[p]);
            ok = fscanf(fid, fmt, &x, &y, &w, &h) == 4;
            if (ok) det.vecPartBoxes.push_back(PSN_Rect(x, y, w, h));
        }
        if (!ok) break;
        if (fscanf(fid, " }") != 0) { /* closing brace consumed by the format */ }
        result.push_back(det);
    }
    fclose(fid);
    return result;
}

bool FilePrintResult(const std::string &path, const stTrack2DResult &r)
{
    FILE *fp = fopen(path.c_str(), "w");
    if (!fp) return false;
    fprintf(fp, "camIdx:%d\nframeIdx:%d\n", (int)r.camID, (int)r.frameIdx);
    fprintf(fp, "numObjectInfos:%d{\n", (int)r.object2DInfos.size());
    for (const stObject2DInfo &o : r.object2DInfos) {
        fprintf(fp, "\t{\n");
        fprintf(fp, "\t\tid:%d\n", o.id);
        fprintf(fp, "\t\tbox:(%f,%f,%f,%f)\n", o.box.x, o.box.y, o.box.w, o.box.h);
        fprintf(fp, "\t\thead:(%f,%f,%f,%f)\n", o.head.x, o.head.y, o.head.w, o.head.h);
        fprintf(fp, "\t\tscore:%f\n", o.score);
        fprintf(fp, "\t\tfeaturePointsPrev:%d,{", (int)o.featurePointsPrev.size());
        for (size_t f = 0; f < o.featurePointsPrev.size(); ++f)
            fprintf(fp, "(%f,%f)%s", o.featurePointsPrev[f].x, o.featurePointsPrev[f].y, f + 1 < o.featurePointsPrev.size() ? "," : "");
        fprintf(fp, "}\n");
        fprintf(fp, "\t\tfeaturePointsCurr:%d,{", (int)o.featurePointsCurr.size());
        for (size_t f = 0; f < o.featurePointsCurr.size(); ++f)
            fprintf(fp, "(%f,%f)%s", o.featurePointsCurr[f].x, o.featurePointsCurr[f].y, f + 1 < o.featurePointsCurr.size() ? "," : "");
        fprintf(fp, "}\n");
        fprintf(fp, "\t}\n");
    }
    fprintf(fp, "}\n");
    fprintf(fp, "detectionRects:0,{}\n");       // the reference never fills vecDetectionRects / vecTrackerRects
    fprintf(fp, "trackerRects:0,{}\n");
    fclose(fp);
    return true;
}

bool Read2DTrackResultWithTxt(const std::string &path, stTrack2DResult &result)
{
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) return false;
    int i1 = 0, i2 = 0, numObj = 0;
    float x, y, w, h, sc;
    bool ok = fscanf(fp, "camIdx:%d\nframeIdx:%d\n", &i1, &i2) == 2 && fscanf(fp, "numObjectInfos:%d{\n", &numObj) == 1;
    result.camID = (unsigned)i1; result.frameIdx = (unsigned)i2;
    result.object2DInfos.clear();
    for (int o = 0; ok && o < numObj; ++o) {
        stObject2DInfo obj;
        ok = fscanf(fp, " { id:%d", &i1) == 1;
        obj.id = (unsigned)i1;
        ok = ok && fscanf(fp, " box:(%f,%f,%f,%f)", &x, &y, &w, &h) == 4;
        obj.box = PSN_Rect(x, y, w, h);
        ok = ok && fscanf(fp, " head:(%f,%f,%f,%f)", &x, &y, &w, &h) == 4;
        obj.head = PSN_Rect(x, y, w, h);
        ok = ok && fscanf(fp, " score:%f", &sc) == 1;
        obj.score = sc;
        for (int which = 0; ok && which < 2; ++which) {
            int n = 0;
            ok = fscanf(fp, which ? " featurePointsCurr:%d,{" : " featurePointsPrev:%d,{", &n) == 1;
            std::vector<Point2f> &dst = which ? obj.featurePointsCurr : obj.featurePointsPrev;
            for (int f = 0; ok && f < n; ++f) {
                ok = fscanf(fp, "(%f,%f)", &x, &y) == 2;
                if (f + 1 < n) { if (fscanf(fp, ",") != 0) {} }
                dst.push_back(Point2f{x, y});
            }
            if (ok && fscanf(fp, "}") != 0) {}
        }
        if (ok && fscanf(fp, " }") != 0) {}
        if (ok) result.object2DInfos.push_back(obj);
    }
    fclose(fp);
    return ok;
}

}  // namespace psn

extern "C" {
#define PSNT_API __attribute__((visibility("default")))

// boxes[cap][4], heads[cap][4] (HEAD part); returns the number of detections in the file
PSNT_API int psnt_read_detections(const char *path, double *boxes, double *heads, int cap)
{
    const std::vector<psn::stDetection> d = psn::ReadDetectionResultWithTxt(path);
    for (int i = 0; i < (int)d.size() && i < cap; ++i) {
        boxes[4 * i] = d[i].box.x; boxes[4 * i + 1] = d[i].box.y; boxes[4 * i + 2] = d[i].box.w; boxes[4 * i + 3] = d[i].box.h;
        const psn::PSN_Rect hd = d[i].vecPartBoxes.empty() ? psn::PSN_Rect() : d[i].vecPartBoxes.front();
        heads[4 * i] = hd.x; heads[4 * i + 1] = hd.y; heads[4 * i + 2] = hd.w; heads[4 * i + 3] = hd.h;
    }
    return (int)d.size();
}

PSNT_API int psnt_write_result(void *tracker, const char *path)
{
    return psn::FilePrintResult(path, static_cast<psn::CPSNWhere_Tracker2D *>(tracker)->Result()) ? 0 : -1;
}

// round trip helper for the tests: write n objects, read them back; returns the number of objects read, ids/boxes filled
PSNT_API int psnt_result_roundtrip(const char *path, unsigned cam, unsigned frame, int n, const unsigned *ids, const double *boxes,
                                   const float *pts, int n_pts_each, unsigned *ids_out, double *boxes_out, float *pts_out)
{
    psn::stTrack2DResult r;
    r.camID = cam; r.frameIdx = frame; r.costRows = r.costCols = 0;
    for (int i = 0; i < n; ++i) {
        psn::stObject2DInfo o;
        o.id = ids[i];
        o.box = psn::PSN_Rect(boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3]);
        o.head = psn::PSN_Rect(boxes[4 * i], boxes[4 * i + 1], boxes[4 * i + 2], boxes[4 * i + 3] * 0.2);
        o.score = 0;
        for (int f = 0; f < n_pts_each; ++f) o.featurePointsPrev.push_back(psn::Point2f{pts[2 * (i * n_pts_each + f)], pts[2 * (i * n_pts_each + f) + 1]});
        r.object2DInfos.push_back(o);
    }
    if (!psn::FilePrintResult(path, r)) return -1;
    psn::stTrack2DResult q;
    if (!psn::Read2DTrackResultWithTxt(path, q)) return -2;
    if (q.camID != cam || q.frameIdx != frame) return -3;
    for (int i = 0; i < (int)q.object2DInfos.size() && i < n; ++i) {
        const psn::stObject2DInfo &o = q.object2DInfos[i];
        ids_out[i] = o.id;
        boxes_out[4 * i] = o.box.x; boxes_out[4 * i + 1] = o.box.y; boxes_out[4 * i + 2] = o.box.w; boxes_out[4 * i + 3] = o.box.h;
        for (int f = 0; f < n_pts_each && f < (int)o.featurePointsPrev.size(); ++f) {
            pts_out[2 * (i * n_pts_each + f)] = o.featurePointsPrev[f].x;
            pts_out[2 * (i * n_pts_each + f) + 1] = o.featurePointsPrev[f].y;
        }
    }
    return (int)q.object2DInfos.size();
}
}
