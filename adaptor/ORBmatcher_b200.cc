// adaptor/ORBmatcher_b200.cc -- drop-in body for ORB_SLAM2::ORBmatcher::SearchByProjection(pF, vpMapPoints, th) on top of
// liborbslam2_dualcam_b200.so (replaces the loop of src/ORBmatcher.cc:539-624).  The other searches follow the same pattern --
// flatten what the loop reads into the PODs of include/orbslam2_dualcam_b200.h, one library call, scatter the result -- and are
// spelled out in INTEGRATION.md §2.  include/ORBmatcher.h stays unchanged.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#ifdef ORB_B200_USE_REFERENCE_HEADERS
#include "ORBmatcher.h"
#else
#include "orbslam_mirror.h"
#endif
#include "orbslam2_dualcam_b200.h"

namespace ORB_SLAM2 {

namespace {
orbm_t* matcher_handle() {                   // the reference's matcher is a stateless value object: one library handle per process
    static orbm_t* h = nullptr;
    if (!h && orbm_create(&h, /*device*/ 0, /*max_pairs*/ 1, /*max_query*/ 8192, /*max_train*/ 8192) != ORB_OK) {
        fprintf(stderr, "ORBmatcher (B200): %s\n", orb_last_error());
        exit(-1);
    }
    return h;
}

// The Frame as the searches see it (src/Frame.cc:141-199).  In a real build this is filled once at the end of Frame::Frame and cached.
struct FlatFrame {
    std::vector<int32_t> n_kp;
    std::vector<uint8_t> desc;
    std::vector<float> bounds;
    orbm_frame_t f;
    explicit FlatFrame(const Frame& F) {
        n_kp.assign(F.mvN.begin(), F.mvN.end());
        desc.resize((size_t)F.totalN * 32);
        size_t g = 0;
        for (int c = 0; c < F.mnCams; c++)                                     // vconcat(mvDescriptors[c]) = global index order
            for (int i = 0; i < F.mvN[c]; i++, g++) memcpy(&desc[g * 32], F.mvDescriptors[c].ptr<uint8_t>(i), 32);
        for (int c = 0; c < F.mnCams; c++) { bounds.push_back(F.mvMinX[c]); bounds.push_back(F.mvMaxX[c]); bounds.push_back(F.mvMinY[c]); bounds.push_back(F.mvMaxY[c]); }
        f.n_cams = F.mnCams; f.n_kp = n_kp.data();
        f.kps_un = reinterpret_cast<const orb_keypoint_t*>(F.mvTotalKeysUn.data());
        f.desc = desc.data(); f.bounds = bounds.data();
        f.n_levels = F.mnScaleLevels; f.scale_factors = F.mvScaleFactors.data();
    }
};
}  // namespace

int ORBmatcher::SearchByProjection(FramePtr pF, const std::vector<MapPointPtr>& vpMapPoints, const float th) {
    FlatFrame flat(*pF);
    std::vector<orbm_mp_t> mp(vpMapPoints.size());
    for (size_t i = 0; i < vpMapPoints.size(); i++) {                          // POD copy of what the loop reads (:547-573)
        MapPointPtr p = vpMapPoints[i];
        memset(&mp[i], 0, sizeof(orbm_mp_t));
        mp[i].valid = p && p->mbTrackInView && !p->isBad();
        if (!mp[i].valid) continue;
        mp[i].cam = p->mTrackProjCamera; mp[i].u = p->mTrackProjX; mp[i].v = p->mTrackProjY;
        mp[i].level = p->mnTrackScaleLevel; mp[i].view_cos = p->mTrackViewCos; mp[i].obs_positive = p->Observations() > 0;
        memcpy(mp[i].desc, p->GetDescriptor().data, 32);
    }
    std::vector<uint8_t> blocked((size_t)pF->totalN);
    for (int g = 0; g < pF->totalN; g++) blocked[g] = pF->mvpMapPoints[g] && pF->mvpMapPoints[g]->Observations() > 0;      // :589-591
    std::vector<int32_t> kp_to_mp((size_t)pF->totalN, -1);
    int32_t n = 0;
    if (orbm_search_by_projection(matcher_handle(), &flat.f, mp.data(), (int)mp.size(), th, mfNNratio, blocked.data(), kp_to_mp.data(), &n) != ORB_OK) {
        fprintf(stderr, "ORBmatcher::SearchByProjection (B200): %s\n", orb_last_error());
        exit(-1);
    }
    for (int g = 0; g < pF->totalN; g++)
        if (kp_to_mp[g] >= 0) pF->mvpMapPoints[g] = vpMapPoints[kp_to_mp[g]];  // :618
    return n;
}

}  // namespace ORB_SLAM2
