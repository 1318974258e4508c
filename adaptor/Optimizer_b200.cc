// adaptor/Optimizer_b200.cc -- drop-in body for ORB_SLAM2::Optimizer::LocalBundleAdjustment on top of liborbslam2_dualcam_b200.so.
//
// Replaces src/Optimizer.cc:407-696: the window selection (:409-460) and the write-back (:640-695) are the reference's own logic
// on the reference's own containers; the g2o graph construction and the two optimize() calls in between (:462-638) become one
// flattening pass into orbba_problem_f32_t (the values as the reference holds them: CV_32F poses / points, float key points,
// mvInvLevelSigma2[octave]) and one orbba_local_f32 call.  include/Optimizer.h and the g2o edge headers stay unchanged.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <list>
#include <map>
#include <vector>

#ifdef ORB_B200_USE_REFERENCE_HEADERS
#include "Optimizer.h"
#else
#include "orbslam_mirror.h"
#endif
#include "orbslam2_dualcam_b200.h"

namespace ORB_SLAM2 {

namespace {
orbba_t* ba_handle() {                        // LocalMapping runs one LocalBundleAdjustment at a time (src/LocalMapping.cc:97-104)
    static orbba_t* h = nullptr;
    if (!h && orbba_create(&h, /*device*/ 0, /*max_problems*/ 1) != ORB_OK) {
        fprintf(stderr, "Optimizer (B200): %s\n", orb_last_error());
        exit(-1);
    }
    return h;
}
void put_pose(const cv::Mat& Tcw, float* o) {  // 4x4 CV_32F -> row-major 3x4
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) o[4 * r + c] = Tcw.at<float>(r, c);
}
}  // namespace

void Optimizer::LocalBundleAdjustment(KeyFramePtr pKF, bool* pbStopFlag, MapPtr pMap, size_t fixId) {
    // ---- Local KeyFrames, local MapPoints, fixed KeyFrames: src/Optimizer.cc:409-460
    std::list<KeyFramePtr> lLocalKeyFrames;
    lLocalKeyFrames.push_back(pKF);
    pKF->mnBALocalForKF = pKF->mnId;
    const std::vector<KeyFramePtr> vNeighKFs = pKF->GetVectorCovisibleKeyFrames();
    for (size_t i = 0; i < vNeighKFs.size(); i++) {
        KeyFramePtr pKFi = vNeighKFs[i];
        pKFi->mnBALocalForKF = pKF->mnId;
        if (!pKFi->isBad()) lLocalKeyFrames.push_back(pKFi);
    }
    std::list<MapPointPtr> lLocalMapPoints;
    for (auto lit = lLocalKeyFrames.begin(); lit != lLocalKeyFrames.end(); lit++) {
        std::vector<MapPointPtr> vpMPs = (*lit)->GetMapPointMatches();
        for (auto vit = vpMPs.begin(); vit != vpMPs.end(); vit++) {
            MapPointPtr pMP = *vit;
            if (pMP && !pMP->isBad() && pMP->mnBALocalForKF != pKF->mnId) { lLocalMapPoints.push_back(pMP); pMP->mnBALocalForKF = pKF->mnId; }
        }
    }
    std::list<KeyFramePtr> lFixedCameras;
    for (auto lit = lLocalMapPoints.begin(); lit != lLocalMapPoints.end(); lit++) {
        std::map<KeyFramePtr, size_t> observations = (*lit)->GetObservations();
        for (auto mit = observations.begin(); mit != observations.end(); mit++) {
            KeyFramePtr pKFi = mit->first;
            if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId) {
                pKFi->mnBAFixedForKF = pKF->mnId;
                if (!pKFi->isBad()) lFixedCameras.push_back(pKFi);
            }
        }
    }
    // ---- flatten: poses (local key frames, then fixed ones), points, one 16-byte record per observation (:475-575)
    std::vector<KeyFramePtr> vKF;
    std::map<KeyFrame*, int> kfIndex;
    std::vector<float> poses;
    std::vector<uint8_t> fixed;
    auto add_kf = [&](KeyFramePtr k, bool fix) {
        kfIndex[k.get()] = (int)vKF.size();
        vKF.push_back(k);
        poses.resize(poses.size() + 12);
        put_pose(k->GetPose(), &poses[poses.size() - 12]);
        fixed.push_back(fix ? 1 : 0);
    };
    for (auto lit = lLocalKeyFrames.begin(); lit != lLocalKeyFrames.end(); lit++) add_kf(*lit, (*lit)->mnId == fixId);   // setFixed(pKFi->mnId == fixId)
    for (auto lit = lFixedCameras.begin(); lit != lFixedCameras.end(); lit++) add_kf(*lit, true);
    std::vector<MapPointPtr> vMP(lLocalMapPoints.begin(), lLocalMapPoints.end());
    std::vector<float> points(3 * vMP.size());
    std::vector<orbba_edge16_t> edges;
    std::vector<KeyFramePtr> vpEdgeKF;
    std::vector<MapPointPtr> vpEdgeMP;
    for (size_t l = 0; l < vMP.size(); l++) {
        const cv::Mat Xw = vMP[l]->GetWorldPos();
        for (int k = 0; k < 3; k++) points[3 * l + k] = Xw.at<float>(k, 0);
        const std::map<KeyFramePtr, size_t> observations = vMP[l]->GetObservations();
        for (auto mit = observations.begin(); mit != observations.end(); mit++) {
            KeyFramePtr pKFi = mit->first;
            if (pKFi->isBad()) continue;
            auto it = kfIndex.find(pKFi.get());
            if (it == kfIndex.end()) continue;          // (every observer is local or fixed by construction)
            const cv::KeyPoint& kpUn = pKFi->mvTotalKeysUn[mit->second];
            orbba_edge16_t e;
            e.point = (uint32_t)l; e.pose = (uint16_t)it->second;
            e.cam = (uint8_t)pKFi->keypointToCam[mit->second];
            e.octave = (uint8_t)kpUn.octave;             // weight = mvInvLevelSigma2[kpUn.octave]  (:553-554)
            e.u = kpUn.pt.x; e.v = kpUn.pt.y;
            edges.push_back(e);
            vpEdgeKF.push_back(pKFi);
            vpEdgeMP.push_back(vMP[l]);
        }
    }
    if (pbStopFlag && *pbStopFlag) return;              // :577-579
    // the rig: intrinsics per camera (taken from the key frame, as the edges do, :561-564), extrinsic and its 6x6 adjoint (:565-571)
    const int nC = pKF->mpCameras->getNCameras();
    std::vector<double> camK(4 * nC), camExt(12 * nC), camAdj(36 * nC);
    for (int c = 0; c < nC; c++) {
        camK[4 * c] = pKF->mvfx[c]; camK[4 * c + 1] = pKF->mvfy[c]; camK[4 * c + 2] = pKF->mvcx[c]; camK[4 * c + 3] = pKF->mvcy[c];
        const cv::Mat E = pKF->mpCameras->getExtrinsici(c), A = pKF->mpCameras->getExtrinsicAdji(c);
        for (int r = 0; r < 3; r++) for (int q = 0; q < 4; q++) camExt[12 * c + 4 * r + q] = E.at<float>(r, q);
        for (int r = 0; r < 6; r++) for (int q = 0; q < 6; q++) camAdj[36 * c + 6 * r + q] = A.at<float>(r, q);
    }
    orbba_problem_f32_t P;
    P.n_poses = (int32_t)vKF.size(); P.n_points = (int32_t)vMP.size(); P.n_edges = (int32_t)edges.size(); P.n_cams = nC;
    P.n_levels = (int32_t)pKF->mvInvLevelSigma2.size();
    P.poses = poses.data(); P.pose_fixed = fixed.data(); P.points = points.data(); P.edges = edges.data();
    P.inv_sigma2 = pKF->mvInvLevelSigma2.data();
    P.cam_K = camK.data(); P.cam_ext = camExt.data(); P.cam_adj = camAdj.data();
    // ---- optimize(5) with the Huber kernel, outlier pass at chi2 > 5.991, optimize(10) without (:581-638): one call
    std::vector<double> poses_out(12 * vKF.size()), points_out(3 * vMP.size());
    std::vector<uint8_t> outlier(edges.size());
    const int rc = orbba_local_f32(ba_handle(), &P, 5, 10, (double)sqrtf(5.991f), 5.991, reinterpret_cast<volatile uint8_t*>(pbStopFlag), poses_out.data(),
                                   points_out.data(), outlier.data(), nullptr);
    if (rc != ORB_OK && rc != ORB_E_ABORTED) { fprintf(stderr, "Optimizer::LocalBundleAdjustment (B200): %s\n", orb_last_error()); exit(-1); }
    // ---- write-back under the map mutex (:640-695)
    std::unique_lock<std::mutex> lock(pMap->mMutexMapUpdate);
    for (size_t i = 0; i < edges.size(); i++) {
        if (!outlier[i] || vpEdgeMP[i]->isBad()) continue;
        vpEdgeKF[i]->EraseMapPointMatch(vpEdgeMP[i]);
        vpEdgeMP[i]->EraseObservation(vpEdgeKF[i]);
    }
    size_t idx = 0;
    for (auto lit = lLocalKeyFrames.begin(); lit != lLocalKeyFrames.end(); lit++, idx++) {
        cv::Mat Tcw(4, 4, CV_32F);
        for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) Tcw.at<float>(r, c) = (float)poses_out[12 * idx + 4 * r + c];
        Tcw.at<float>(3, 0) = 0; Tcw.at<float>(3, 1) = 0; Tcw.at<float>(3, 2) = 0; Tcw.at<float>(3, 3) = 1;
        (*lit)->SetPose(Tcw);
    }
    for (size_t l = 0; l < vMP.size(); l++) {
        cv::Mat X(3, 1, CV_32F);
        for (int k = 0; k < 3; k++) X.at<float>(k, 0) = (float)points_out[3 * l + k];
        vMP[l]->SetWorldPos(X);
        vMP[l]->UpdateNormalAndDepth();
    }
}

}  // namespace ORB_SLAM2
