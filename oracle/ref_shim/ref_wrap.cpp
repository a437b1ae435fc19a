/* oracle/_ref: the reference's OWN LocalSearchKLT, compiled from the lines where they lie under /root/reference.
 * TEST INFRASTRUCTURE ONLY -- never linked into or loaded by the product library.
 *
 * oracle/Makefile (target _ref) extracts, at build time and only into oracle/_ref/ (git-ignored):
 *   ref_types.inc   psn_where/PSNWhere_Types.h  `class PSN_Point2D` .. end of `class PSN_Rect`   (:18-182)
 *   ref_params.inc  psn_where/PSNWhere_Tracker2D.cpp  the `#define PSN_2D_*` parameter block       (:10-33)
 *   ref_lsk.inc     psn_where/PSNWhere_Tracker2D.cpp  `#define PSN_LOCAL_SEARCH_KLT_*` + LocalSearchKLT (:452-554)
 * and this file wraps them, unmodified, behind a C entry point.  Nothing of the reference is copied into the repo. */
#include "cv_shim.h"

#include "ref_params.inc"
#include "ref_types.inc"

class CPSNWhere_Tracker2D
{
public:
    /* declaration as in psn_where/PSNWhere_Tracker2D.h:147 */
    static PSN_Rect LocalSearchKLT(PSN_Rect preBox, std::vector<cv::Point2f> &preFeatures, std::vector<cv::Point2f> &curFeatures,
                                   std::vector<size_t> &inlierFeatureIndex);
};

#include "ref_lsk.inc"

extern "C" __attribute__((visibility("default")))
int psnref_local_search_klt(const double *box, const float *pre_xy, const float *cur_xy, int n, double *box_out, int *inlier_idx)
{
    std::vector<cv::Point2f> pre(n), cur(n);
    for (int i = 0; i < n; ++i) {
        pre[i] = cv::Point2f(pre_xy[2 * i], pre_xy[2 * i + 1]);
        cur[i] = cv::Point2f(cur_xy[2 * i], cur_xy[2 * i + 1]);
    }
    std::vector<size_t> inl;
    PSN_Rect r = CPSNWhere_Tracker2D::LocalSearchKLT(PSN_Rect(box[0], box[1], box[2], box[3]), pre, cur, inl);
    box_out[0] = r.x; box_out[1] = r.y; box_out[2] = r.w; box_out[3] = r.h;
    for (size_t i = 0; i < inl.size(); ++i) inlier_idx[i] = (int)inl[i];
    return (int)inl.size();
}

/* PSN_Rect helpers the parity metrics and the mask ROI use (PSNWhere_Types.h:146-153, :158, :171-178) */
extern "C" __attribute__((visibility("default")))
double psnref_overlapped_area(const double *a, const double *b)
{
    return PSN_Rect(a[0], a[1], a[2], a[3]).overlappedArea(PSN_Rect(b[0], b[1], b[2], b[3]));
}
extern "C" __attribute__((visibility("default")))
int psnref_contain(const double *box, float x, float y)
{
    return PSN_Rect(box[0], box[1], box[2], box[3]).contain(cv::Point2f(x, y)) ? 1 : 0;
}
extern "C" __attribute__((visibility("default")))
void psnref_crop_roi(const double *box, double width, double height, int *roi)
{
    cv::Rect r = PSN_Rect(box[0], box[1], box[2], box[3]).cropWithSize(width, height).cv();
    roi[0] = r.x; roi[1] = r.y; roi[2] = r.width; roi[3] = r.height;
}
