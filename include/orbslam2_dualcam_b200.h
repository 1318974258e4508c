/* orbslam2_dualcam_b200.h -- C-ABI of liborbslam2_dualcam_b200.so
 *
 * Drop-in boundary for the hot path of lixiny/ORB-SLAM2-DualCam.  The reference has no
 * FFI of its own: the path sits behind three C++ classes compiled into libORB_SLAM2_DualCam.so.  Each
 * entry point below names the reference interface it replaces; INTEGRATION.md shows the C++ adaptor a
 * maintainer adds inside ORBextractor / ORBmatcher / Optimizer to call them.
 *
 * Conventions: POD only, caller-allocated outputs with explicit capacities, int status return
 * (0 = ok, <0 = ORB_E_*), no exceptions, no exit().  One handle per GPU and per calling thread
 * (handles are not re-entrant -- same as the stateful reference ORBextractor).  All kernels are
 * sm_100a CUDA; there is NO CPU fallback: without a usable CUDA device every create() fails with
 * ORB_E_NO_DEVICE.
 */
#ifndef ORBSLAM2_DUALCAM_B200_H
#define ORBSLAM2_DUALCAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORB_OK             0
#define ORB_E_INVALID     -1   /* bad argument (null pointer, size out of range, capacity too small) */
#define ORB_E_NO_DEVICE   -2   /* no CUDA device / wrong architecture / driver error at create */
#define ORB_E_CUDA        -3   /* a CUDA call failed; see orb_last_error() */
#define ORB_E_OVERFLOW    -4   /* an internal capacity was exceeded (results would be truncated) */
#define ORB_E_ABORTED     -5   /* stop flag was raised (bundle adjustment) */

/* last error text of the calling thread (never NULL) */
const char* orb_last_error(void);
/* library version string */
const char* orb_version(void);

/* ------------------------------------------------------------------------------------------------
 * cv::KeyPoint memory layout (7 x 4 bytes) -- what ORBextractor::operator() fills
 * (reference include/ORBextractor.h:59-61, src/ORBextractor.cc:1043-1105).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float   x, y;        /* pt, in level-0 pixel units (pt * mvScaleFactor[octave]) */
    float   size;        /* (int)(31 * mvScaleFactor[octave]) */
    float   angle;       /* degrees [0,360], IC_Angle / fastAtan2 */
    float   response;    /* FAST corner score */
    int32_t octave;      /* pyramid level */
    int32_t class_id;    /* -1 */
} orb_keypoint_t;

/* ================================================================================================
 * EXTRACT -- replaces ORBextractor (include/ORBextractor.h:47-108, src/ORBextractor.cc:410-1132),
 * batched over B dual-frames x C cameras.  One reference ORBextractor object per camera with equal
 * parameters (src/Tracking.cc:204-207) maps onto ONE orbx handle.
 * ============================================================================================== */
typedef struct orbx orbx_t;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)  (src/ORBextractor.cc:410-470)
 * plus the static shape of the batch: image width/height, cameras per frame, max frames per call. */
int  orbx_create(orbx_t** out, int device, int width, int height, int cameras, int max_frames,
                 int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void orbx_destroy(orbx_t*);

/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (include/ORBextractor.h:63-83) and mnFeaturesPerLevel / umax (src/ORBextractor.cc:435-469).  Any pointer may be NULL. */
int  orbx_get_levels(const orbx_t*);
int  orbx_get_tables(const orbx_t*, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                     int32_t* features_per_level, int32_t* umax16);
/* Per-image keypoint capacity the caller must provide: the quadtree may return up to nfeatures + 2*nlevels
 * keypoints (src/ORBextractor.cc:730-731 breaks only AFTER reaching N). */
int  orbx_max_keypoints(const orbx_t*);
/* level geometry: width/height of pyramid level l (src/ORBextractor.cc:1111-1112) */
int  orbx_level_size(const orbx_t*, int level, int* w, int* h);

/* ORBextractor::operator()(image, mask, keypoints, descriptors) for frames*cameras images  (src/ORBextractor.cc:1043-1105).
 *   imgs      HOST  uint8 [frames][cameras][height][row_stride]   (CV_8UC1; mask is ignored as in the reference)
 *   kps       HOST  [frames][cameras][kp_capacity]                 kp_capacity >= orbx_max_keypoints()
 *   desc      HOST  uint8 [frames][cameras][kp_capacity][32]
 *   counts    HOST  int32 [frames][cameras]                        keypoints found per image (0 for an empty image)
 * Copies host->device, runs the kernels, copies device->host, synchronises. */
int  orbx_extract(orbx_t*, const uint8_t* imgs, int frames, size_t row_stride,
                  orb_keypoint_t* kps, uint8_t* desc, int32_t* counts, int kp_capacity);

/* Same computation with DEVICE pointers, asynchronous on the handle's stream (no copies, no sync).
 * imgs must be 16-byte aligned with row_stride % 16 == 0. */
int  orbx_extract_device(orbx_t*, const uint8_t* d_imgs, int frames, size_t row_stride,
                         orb_keypoint_t* d_kps, uint8_t* d_desc, int32_t* d_counts, int kp_capacity);
/* stream control for the device API */
int  orbx_set_stream(orbx_t*, void* cuda_stream);   /* cudaStream_t; NULL = the handle's own stream; cudaStreamLegacy (0x1) = the default stream */
int  orbx_synchronize(orbx_t*);
/* number of kernel launches issued by this handle since creation (bench.py's gpu_launches) */
long long orbx_launch_count(const orbx_t*);

/* Per-stage device timing for bench.py's roofline: when enabled, every orbx_extract_device call records CUDA events
 * between its stages on the launching stream (up to 256 calls are kept).  orbx_stage_ms synchronises the stream, returns
 * the summed milliseconds of {pyramid, fast cells, quadtree, describe} over the recorded calls, and resets. */
int  orbx_profile(orbx_t*, int enable);
int  orbx_stage_ms(orbx_t*, double* ms4, int* calls);

/* Stage taps for parity tests (device -> host copies of intermediate results of the LAST extract call).
 *   orbx_debug_level:      level image of image index `img` (dense w*h into out)
 *   orbx_debug_candidates: FAST+NMS survivors of (img, level) as (x, y, score) int triplets relative to the
 *                          16-px border, UNORDERED (the reference order is a function of (x,y))
 *   orbx_debug_selected:   quadtree output of (img, level) as (x, y, score) in level pixel coords, list order */
int  orbx_debug_level(orbx_t*, int img, int level, uint8_t* out, size_t out_bytes);
int  orbx_debug_candidates(orbx_t*, int img, int level, int32_t* xys, int cap);
int  orbx_debug_selected(orbx_t*, int img, int level, int32_t* xys, int cap);

/* ================================================================================================
 * MATCH -- replaces the Hamming searches of ORBmatcher (include/ORBmatcher.h, src/ORBmatcher.cc).
 * ============================================================================================== */
typedef struct orbm orbm_t;

#define ORBM_TH_LOW        50   /* ORBmatcher::TH_LOW        src/ORBmatcher.cc:58 */
#define ORBM_TH_HIGH      100   /* ORBmatcher::TH_HIGH       src/ORBmatcher.cc:57 */
#define ORBM_HISTO_LENGTH  30   /* ORBmatcher::HISTO_LENGTH  src/ORBmatcher.cc:59 */

int  orbm_create(orbm_t** out, int device, int max_pairs, int max_query, int max_train);
void orbm_destroy(orbm_t*);
int  orbm_set_stream(orbm_t*, void* cuda_stream);
int  orbm_synchronize(orbm_t*);
long long orbm_launch_count(const orbm_t*);

/* device time of the brute-force kernel over the recorded calls (same protocol as orbx_profile / orbx_stage_ms) */
int  orbm_profile(orbm_t*, int enable);
int  orbm_stage_ms(orbm_t*, double* ms1, int* calls);

/* ORBmatcher::DescriptorDistance(a, b) (src/ORBmatcher.cc:2015-2031) for n descriptor pairs (HOST buffers). */
int  orbm_descriptor_distance(orbm_t*, const uint8_t* a, const uint8_t* b, int n, int32_t* dist);

/* Brute-force 256-bit Hamming nearest / second-nearest (BASELINE.json configs[1]): for each of `pairs` independent
 * (query set, train set) pairs and each query descriptor, the train index with the smallest distance (lowest index
 * wins ties -- the reference's `if(dist<bestDist)` scan order, e.g. src/ORBmatcher.cc:575-610), that distance and the
 * second-smallest distance (256 when it does not exist; best_idx -1 when the train set is empty).
 *   dq   uint8 [pairs][q_capacity][32], nq int32 [pairs] valid queries per pair
 *   dt   uint8 [pairs][t_capacity][32], nt int32 [pairs]
 *   best_idx/best_d/second_d int32 [pairs][q_capacity]  (entries >= nq[p] are left untouched)
 * HOST buffers, synchronous. */
int  orbm_bruteforce(orbm_t*, const uint8_t* dq, const int32_t* nq, int q_capacity,
                     const uint8_t* dt, const int32_t* nt, int t_capacity, int pairs,
                     int32_t* best_idx, int32_t* best_d, int32_t* second_d);
/* DEVICE buffers, asynchronous on the handle's stream. */
int  orbm_bruteforce_device(orbm_t*, const uint8_t* d_dq, const int32_t* d_nq, int q_capacity,
                            const uint8_t* d_dt, const int32_t* d_nt, int t_capacity, int pairs,
                            int32_t* d_best_idx, int32_t* d_best_d, int32_t* d_second_d);

/* Same search over a POOL of descriptor sets (e.g. the [frames][cameras] output of orbx_extract_device): pair p matches
 * set q_set[p] (queries) against set t_set[p] (train), e.g. camera c of frame k against camera c of frame k+1.
 *   d_desc uint8 [n_sets][capacity][32], d_counts int32 [n_sets], d_q_set/d_t_set int32 [pairs]
 *   outputs int32 [pairs][capacity] (entries >= counts[q_set[p]] untouched).  DEVICE buffers, asynchronous. */
int  orbm_bruteforce_sets_device(orbm_t*, const uint8_t* d_desc, const int32_t* d_counts, int capacity, int n_sets,
                                 const int32_t* d_q_set, const int32_t* d_t_set, int pairs,
                                 int32_t* d_best_idx, int32_t* d_best_d, int32_t* d_second_d);

/* ------------------------------------------------------------------------------------------------
 * Guided searches.  The adaptor flattens what the reference loops read from Frame / KeyFrame / MapPoint into the PODs
 * below; the sequential "claim" semantics of the reference (a keypoint that received a map point is skipped by the map
 * points that follow, src/ORBmatcher.cc:589-591, :216) are reproduced exactly: distances are computed in parallel, the
 * claims are resolved in input order.  HOST buffers, synchronous.
 * ---------------------------------------------------------------------------------------------- */
#define ORBM_GRID_COLS 64       /* FRAME_GRID_COLS  include/Frame.h:40 */
#define ORBM_GRID_ROWS 48       /* FRAME_GRID_ROWS  include/Frame.h:39 */

/* a Frame as the searches see it (src/Frame.cc:141-199): undistorted keypoints of all cameras concatenated camera-major
 * (= the global index order of mvTotalKeysUn), their descriptors, per-camera image bounds and the pyramid scale factors.
 * The 64x48 grid (Frame::PosInGrid / mvGrids, src/Frame.cc:179-196,380-390) is built on the device. */
typedef struct {
    int32_t n_cams;
    const int32_t* n_kp;            /* [n_cams]      Frame::mvN */
    const orb_keypoint_t* kps_un;   /* [totalN]      Frame::mvTotalKeysUn (pt, angle, octave are read) */
    const uint8_t* desc;            /* [totalN][32]  rows of mvDescriptors[c] */
    const float* bounds;            /* [n_cams][4]   mvMinX, mvMaxX, mvMinY, mvMaxY */
    int32_t n_levels;
    const float* scale_factors;     /* [n_levels]    mvScaleFactors */
} orbm_frame_t;

/* what ORBmatcher::SearchByProjection(pF, vpMapPoints, th) reads from a MapPoint (src/ORBmatcher.cc:547-573) */
typedef struct {
    int32_t valid;          /* pMP && pMP->mbTrackInView && !pMP->isBad() */
    int32_t cam;            /* mTrackProjCamera */
    float   u, v;           /* mTrackProjX, mTrackProjY */
    int32_t level;          /* mnTrackScaleLevel */
    float   view_cos;       /* mTrackViewCos */
    int32_t obs_positive;   /* Observations() > 0: once assigned, the keypoint is skipped by later map points (:589-591) */
    uint8_t desc[32];       /* GetDescriptor() */
} orbm_mp_t;

/* ORBmatcher::SearchByProjection(FramePtr pF, const vector<MapPointPtr>& vpMapPoints, float th)  (TrackLocalMap; src/ORBmatcher.cc:539-624).
 *   blocked   uint8 [totalN]  1 iff pF->mvpMapPoints[g] && pF->mvpMapPoints[g]->Observations() > 0 on entry
 *   kp_to_mp  int32 [totalN]  in/out: entry g is set to i when vpMapPoints[i] is written into pF->mvpMapPoints[g] (:618)
 *   nmatches  the function's return value */
int  orbm_search_by_projection(orbm_t*, const orbm_frame_t* frame, const orbm_mp_t* mps, int n, float th, float nnratio,
                               const uint8_t* blocked, int32_t* kp_to_mp, int32_t* nmatches);

/* the last frame as ORBmatcher::SearchByProjectionOnCam(pFcurt, query, pFlast, th) reads it (src/ORBmatcher.cc:975-1036), one entry per
 * last-frame keypoint (global index) */
typedef struct {
    int32_t n;                    /* pFlast->totalN */
    const int32_t* cam;           /* [n]     keypointToCam */
    const uint8_t* valid;         /* [n]     mvpMapPoints[i] && !isBad() */
    const float*   pos;           /* [n][3]  GetWorldPos() */
    const uint8_t* desc;          /* [n][32] GetDescriptor() */
    const int32_t* octave;        /* [n]     mvTotalKeysUn[i].octave */
    const float*   angle;         /* [n]     mvTotalKeysUn[i].angle */
    const uint8_t* obs_positive;  /* [n]     Observations() > 0 */
} orbm_lastframe_t;

/* ORBmatcher::SearchByProjection(pCurrentFrame, pLastFrame, th, bMapScaled)  (TrackWithMotionModel; src/ORBmatcher.cc:634-690, 954-1113).
 *   Rsw, tsw   float [n_cams][9] / [n_cams][3]: rotation / translation of Tsw = mvExtrinsics[c] * mTcw (:962-967)
 *   K          float [n_cams][4]: mvfx, mvfy, mvcx, mvcy
 *   kp_to_last int32 [totalN] in/out: entry g is set to i when pFlast->mvpMapPoints[i] is written into pFcurt->mvpMapPoints[g],
 *              and back to -1 when the rotation-histogram check removes it (:1082-1101)
 *   per_cam    int32 [n_cams] (may be NULL): matches per camera; cameras after one with <= 20 matches are not searched and the
 *              total is REPLACED by that camera's count (:664-667, upstream behaviour) */
int  orbm_search_by_projection_last(orbm_t*, const orbm_frame_t* cur, const float* Rsw, const float* tsw, const float* K,
                                    const orbm_lastframe_t* last, float th, int check_orientation, int map_scaled,
                                    const uint8_t* blocked, int32_t* kp_to_last, int32_t* per_cam, int32_t* nmatches);

/* one side of SearchByBoW: descriptors, keypoint angles and the DBoW2::FeatureVector of every camera flattened to CSR
 * (std::map<NodeId, vector<unsigned>> -> node ids ascending, camera-local feature indices) */
typedef struct {
    int32_t n_cams;
    const int32_t* n_kp;          /* [n_cams] */
    const uint8_t* desc;          /* [totalN][32] */
    const float*   angle;         /* [totalN]  mvvkeysUnTemp[c][i].angle */
    const int32_t* node_first;    /* [n_cams + 1] range of camera c's nodes in node_id / node_off */
    const int32_t* node_id;       /* [n_nodes] ascending inside a camera */
    const int32_t* node_off;      /* [n_nodes + 1] into idx */
    const int32_t* idx;           /* camera-local feature indices */
} orbm_bowside_t;

/* ORBmatcher::SearchByBoW(pF, pKF, vpMapPointMatches, bMapScaled) -> SearchByBoWCrossCam(pF, c, pKF, c, ...)  (src/ORBmatcher.cc:102-294).
 *   kf_mp_valid uint8 [KF totalN]  pKF->GetMapPointMatches()[g] && !isBad()
 *   f_to_kf     int32 [F totalN]   out: global KF keypoint index whose map point lands in vpMapPointMatches[g], else -1 */
int  orbm_search_by_bow(orbm_t*, const orbm_bowside_t* F, const orbm_bowside_t* KF, const uint8_t* kf_mp_valid, float nnratio,
                        int check_orientation, int map_scaled, int32_t* f_to_kf, int32_t* nmatches);

/* Frame::isInFrustum(pMP, viewingCosLimit, bForAllCam) + MapPoint::PredictScale for n map points (src/Frame.cc:244-312,
 * src/MapPoint.cc:440-455): the producer of orbm_mp_t {cam, u, v, level, view_cos}. */
typedef struct {
    int32_t n_cams, n_levels;
    const float* Rsw;             /* [n_cams][9]  rotation of mvExtrinsics[c] * mTcw */
    const float* tsw;             /* [n_cams][3] */
    const float* Ow;              /* [n_cams][3]  Frame::GetCameraCenter(c) */
    const float* K;               /* [n_cams][4]  fx fy cx cy */
    const float* bounds;          /* [n_cams][4]  mvMinX, mvMaxX, mvMinY, mvMaxY */
    float log_scale_factor;       /* mfLogScaleFactor */
} orbm_frustum_t;
/*   pos, normal float [n][3]; max_dist, min_dist float [n] = mfMaxDistance, mfMinDistance (the 1.2 / 0.8 invariance factors are applied inside)
 *   out int32 [n][3] = {mbTrackInView, mTrackProjCamera, mnTrackScaleLevel};  uvc float [n][3] = {mTrackProjX, mTrackProjY, mTrackViewCos} */
int  orbm_is_in_frustum(orbm_t*, const orbm_frustum_t* frame, const float* pos, const float* normal, const float* max_dist,
                        const float* min_dist, int n, float viewing_cos_limit, int for_all_cams, int32_t* out, float* uvc);

/* ---- key-frame flavoured searches (SURVEY a14, a15 second half, f4).  The map mutations the reference performs inside some of these
 * loops (AddMapPoint / Replace / AddObservation / UpdateConnections) stay with the caller; the library returns the decisions. ---- */
typedef struct {                  /* candidate map points */
    int32_t n;
    const uint8_t* valid;         /* [n]     the caller's skip conditions: pMP && !isBad() && !sAlreadyFound.count(pMP) (resp. !IsInKeyFrame(pKF)) */
    const float*   pos;           /* [n][3]  GetWorldPos() */
    const float*   normal;        /* [n][3]  GetNormal(); may be NULL for the variants without the 60 degree test */
    const float*   max_dist;      /* [n]     mfMaxDistance (GetMaxDistanceInvariance() / 1.2) */
    const float*   min_dist;      /* [n]     mfMinDistance */
    const uint8_t* desc;          /* [n][32] GetDescriptor() */
    const float*   angle;         /* [n]     pKF->mvTotalKeysUn[i].angle; only read by the orientation check, may be NULL */
} orbm_points_t;

/* ORBmatcher::SearchByProjectionOnCam(pF, query, pKF, sAlreadyFound, th, ORBdist)  (relocalisation; src/ORBmatcher.cc:812-951).
 *   view        Rsw / tsw / Ow of mvExtrinsics[c] * mTcw for every camera of pF, K, log scale factor (bounds are taken from F)
 *   P           pKF->GetMapPointMatches(), index = global key point index of pKF
 *   blocked     uint8 [totalN]  pF->mvpMapPoints[g] != NULL
 *   kp_to_point int32 [totalN]  in/out: entry g is set to i when P[i] is written into pF->mvpMapPoints[g] (entries the orientation check
 *               resets to NULL are left untouched) */
int  orbm_search_by_projection_reloc(orbm_t*, const orbm_frame_t* F, const orbm_frustum_t* view, int cam, const orbm_points_t* P, float th,
                                     int orb_dist, int check_orientation, const uint8_t* blocked, int32_t* kp_to_point, int32_t* nmatches);

/* ORBmatcher::SearchByProjection(pKF, query, Scq_w, vpPoints, vpMatched, th)  (loop closing; src/ORBmatcher.cc:416-536).
 *   view           slot `cam` holds the decomposed similarity: Rsw = Rcqw, tsw = tcqw, Ow = Ocqw (:431-435)
 *   matched_local  uint8 [n_kp[cam]]  vpMatched[l] != NULL -- indexed by the CAMERA-LOCAL key point index, as the reference indexes it (:504, :523)
 *   local_to_point int32 [n_kp[cam]]  in/out: entry l is set to i when vpMatched[l] = vpPoints[i]
 *   kf_index_quirk 1 = KeyFrame::GetFeaturesInArea as upstream (window test on mvTotalKeysUn[camera-local index], src/KeyFrame.cc:757), 0 = the
 *                  camera's own key point; identical for camera 0 */
int  orbm_search_by_projection_sim3(orbm_t*, const orbm_frame_t* KF, const orbm_frustum_t* view, int cam, const orbm_points_t* P, int th,
                                    int kf_index_quirk, const uint8_t* matched_local, int32_t* local_to_point, int32_t* nmatches);

/* Search part of the three loops that mutate the map while they search: for every camera s of pKF and every map point i the key point with
 * the smallest descriptor distance inside the projected window (global index, -1 if none) and that distance (256 if none);
 * best_kp / best_dist are int32 [n_cams][n].  The caller applies the threshold and the Replace / AddMapPoint branch in the reference's order.
 *   ORBM_KF_SEARCH     SearchByProjection(pKF, vpMapPoints, sAlreadyFound, th, ORBdist)       src/ORBmatcher.cc:693-775  (threshold ORBdist)
 *   ORBM_KF_FUSE       Fuse(pKF, vpMapPoints, th)                                            src/ORBmatcher.cc:1431-1527 (threshold TH_LOW)
 *   ORBM_KF_FUSE_SIM3  Fuse(pKF, Scw, vpPoints, th, vpReplacePoint); view = vRsw/vtsw/vOsw   src/ORBmatcher.cc:1560-1668 (threshold TH_LOW) */
#define ORBM_KF_SEARCH    0
#define ORBM_KF_FUSE      1
#define ORBM_KF_FUSE_SIM3 2
int  orbm_project_best(orbm_t*, const orbm_frame_t* KF, const orbm_frustum_t* view, const orbm_points_t* P, float th, int variant,
                       int kf_index_quirk, int32_t* best_kp, int32_t* best_dist);
/* one camera of the above (best_kp / best_dist [n]).  The reference's loops change the map between cameras -- a point matched in camera 0
 * has its normal, depth range and descriptor refreshed before camera 1 projects it again (src/ORBmatcher.cc:783-787), Fuse skips in camera 1
 * what camera 0 added (:1452) -- so the exact adaptor is  for (s in cameras) { flatten the current state; orbm_project_best_cam(.., s, ..);
 * apply the reference's own `if (bestDist <= ...)` block in list order }. */
int  orbm_project_best_cam(orbm_t*, const orbm_frame_t* KF, const orbm_frustum_t* view, const orbm_points_t* points, float th, int variant,
                           int kf_index_quirk, int cam, int32_t* best_kp, int32_t* best_dist);

/* ORBmatcher::SearchByBoWCrossCam(pKF1, c1, pKF2, c2, vpMatches12)  (src/ORBmatcher.cc:297-414).  mp_valid1 / mp_valid2 are indexed by the
 * global key point index (pMP && !isBad()); matches12 int32 [K1->n_kp[c1]]: GLOBAL key point index of pKF2 whose map point is
 * vpMatches12[idx1local], else -1.  (SearchByBoWCrossCam(pF, cF, pKF, cKF, ...) with cF != cKF is orbm_search_by_bow on one-camera sides.) */
int  orbm_search_by_bow_kf(orbm_t*, const orbm_bowside_t* K1, int c1, const orbm_bowside_t* K2, int c2, const uint8_t* mp_valid1,
                           const uint8_t* mp_valid2, float nnratio, int check_orientation, int32_t* matches12, int32_t* nmatches);

/* ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, camS) + CheckDistEpipolarLine  (src/ORBmatcher.cc:1253-1427, 74-92).
 *   kps1 / kps2  global undistorted key point arrays; has_mp1 / has_mp2 uint8 (global): GetMapPoint(g) != NULL
 *   F12 float [9] row-major; C1sw = pKF1->GetCameraCenter(camS); R2sw, t2sw = pKF2->GetRotation / GetTranslation(camS); K2cam = fx fy cx cy of camS
 *   matches12 int32 [K1->n_kp[cam]]: camera-local index in pKF2 or -1 (vMatchedPairs = the entries >= 0, both made global by the caller) */
int  orbm_search_for_triangulation(orbm_t*, const orbm_bowside_t* K1, const orbm_bowside_t* K2, int cam, const orb_keypoint_t* kps1,
                                   const orb_keypoint_t* kps2, const uint8_t* has_mp1, const uint8_t* has_mp2, const float* F12, const float* C1sw,
                                   const float* R2sw, const float* t2sw, const float* K2cam, const float* scale_factors, int n_levels,
                                   int check_orientation, int32_t* matches12, int32_t* nmatches);

/* Frame::UndistortKeyPoints(c) (src/Frame.cc:410-442): cv::undistortPoints(mat, mat, K, distCoef, Mat(), K) on the keypoint centres, every other
 * field copied; a copy when distCoef[0] == 0.  K4 = fx fy cx cy, dist = k1 k2 p1 p2 [k3 ...] (n_dist <= 12).  kps_un may alias kps.  HOST buffers. */
int  orbm_undistort_keypoints(orbm_t*, const orb_keypoint_t* kps, int n, const float* K4, const float* dist, int n_dist, orb_keypoint_t* kps_un);
/* Frame::ComputeImageBounds(c) (src/Frame.cc:454-490): bounds = {mvMinX, mvMaxX, mvMinY, mvMaxY} from the undistorted image corners. */
int  orbm_image_bounds(orbm_t*, int width, int height, const float* K4, const float* dist, int n_dist, float* bounds);

/* ================================================================================================
 * BAG OF WORDS -- replaces ORBVocabulary::transform as Frame::ComputeBoW / KeyFrame::ComputeBoW call it (src/Frame.cc:393-408,
 * Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1149-1227, 1249-1292): the producer of the DBoW2::FeatureVector that orbm_search_by_bow*
 * and orbm_search_for_triangulation consume, and of the BowVector of the key-frame database.  TF_IDF weighting and L1 normalisation,
 * i.e. the ORB vocabulary's header "10 6 0 0".
 * ================================================================================================ */
typedef struct orbv orbv_t;

/* The node table of TemplatedVocabulary::loadFromTextFile (:1362-1447): row 0 is the root (ignored), row i >= 1 is line i of the text
 * file: parent id, leaf flag, 32 descriptor bytes, weight.  Word ids are assigned to the leaves in file order, children keep file order. */
int  orbv_create(orbv_t** out, int device, int k, int L, int n_nodes, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* desc,
                 const double* weight);
void orbv_destroy(orbv_t*);
int  orbv_set_stream(orbv_t*, void* cuda_stream);
int  orbv_words(const orbv_t*);
long long orbv_launch_count(const orbv_t*);

/* transform(features, v, fv, levelsup) for n_sets descriptor sets at once (one set = the descriptors of one camera image; set s is rows
 * set_off[s] .. set_off[s+1] of desc, at most 8192 rows).  HOST buffers, synchronous.  All per-feature outputs are laid out like desc:
 *   word_id, node_id int32 [n]   per feature: word, and ancestor at level L - levelsup (may be NULL)
 *   bow_ids int32 / bow_vals double [n], n_words int32 [n_sets]
 *                                BowVector of set s = entries set_off[s] .. + n_words[s]: word ids ascending, L1-normalised tf-idf values
 *   fv_node int32 [n], fv_off int32 [n + n_sets], fv_idx int32 [n], n_fv_nodes int32 [n_sets]
 *                                FeatureVector of set s: node ids fv_node[set_off[s] ..+ n_fv_nodes[s]] ascending; node j owns the features
 *                                fv_idx[set_off[s] + o[j] .. set_off[s] + o[j+1]) with o = fv_off + set_off[s] + s  (set-local indices, ascending)
 * Features whose word has weight 0 (stopped words) appear in neither vector. */
int  orbv_transform(orbv_t*, const uint8_t* desc, const int32_t* set_off, int n_sets, int levelsup, int32_t* word_id, int32_t* node_id,
                    int32_t* bow_ids, double* bow_vals, int32_t* n_words, int32_t* fv_node, int32_t* fv_off, int32_t* fv_idx, int32_t* n_fv_nodes);

/* ================================================================================================
 * BUNDLE ADJUSTMENT -- replaces Optimizer::LocalBundleAdjustment / BundleAdjustment / GlobalBundleAdjustemnt
 * (include/Optimizer.h:50-56, src/Optimizer.cc:62-248,407-696) together with the g2o machinery under them
 * (dual-camera EdgeSE3ProjectXYZ, Huber kernel, BlockSolver_6_3 Schur complement, Levenberg-Marquardt).
 *
 * The adaptor flattens the graph it used to build with `new g2o::Vertex/Edge` (src/Optimizer.cc:460-580):
 *   poses   rig poses Tcw = KeyFrame::GetPose(), row-major 3x4 [R|t] (CV_32F values widened to double), ascending mnId
 *   fixed   1 for lFixedCameras and for mnId == fixId (src/Optimizer.cc:480,495)
 *   points  MapPoint::GetWorldPos(), ascending mnId
 *   edges   one per observation: pose index, point index, camera of the rig (KeyFrame::keypointToCam), undistorted
 *           keypoint (mvTotalKeysUn[idx].pt), mvInvLevelSigma2[octave]
 *   cams    per camera fx fy cx cy, extrinsic Cameras::getExtrinsici (3x4) and the 6x6 Cameras::getExtrinsicAdji
 *           (src/Cameras.cc:17-40; passed through as opaque data, lower-left block as the caller holds it)
 * ============================================================================================== */
typedef struct {
    int32_t n_poses, n_points, n_edges, n_cams;
    const double*  poses;            /* [n_poses][12] */
    const uint8_t* pose_fixed;       /* [n_poses] */
    const double*  points;           /* [n_points][3] */
    const int32_t* edge_pose;        /* [n_edges] */
    const int32_t* edge_point;       /* [n_edges] */
    const int32_t* edge_cam;         /* [n_edges] */
    const double*  edge_obs;         /* [n_edges][2] */
    const double*  edge_inv_sigma2;  /* [n_edges] */
    const double*  cam_K;            /* [n_cams][4] */
    const double*  cam_ext;          /* [n_cams][12] */
    const double*  cam_adj;          /* [n_cams][36] */
} orbba_problem_t;

/* The same graph in the reference's own storage types (what an adaptor reads without widening anything): poses / points CV_32F
 * (KeyFrame::GetPose(), MapPoint::GetWorldPos()), one 16-byte record per observation (cv::KeyPoint::pt is float, the weight is
 * mvInvLevelSigma2[kpUn.octave], src/Optimizer.cc:549-571) and the extractor's per-level weight table.  Values are widened to
 * double on the device, so the results equal those of orbba_problem_t holding the same (float-representable) numbers, and the host ->
 * device copy carries 16 instead of 36 bytes per edge. */
typedef struct {
    uint32_t point;                  /* map point index */
    uint16_t pose;                   /* key frame index */
    uint8_t  cam;                    /* camera of the rig */
    uint8_t  octave;                 /* kpUn.octave: weight = inv_sigma2[octave] */
    float    u, v;                   /* kpUn.pt */
} orbba_edge16_t;
typedef struct {
    int32_t n_poses, n_points, n_edges, n_cams, n_levels;
    const float*          poses;       /* [n_poses][12] */
    const uint8_t*        pose_fixed;  /* [n_poses] */
    const float*          points;      /* [n_points][3] */
    const orbba_edge16_t* edges;       /* [n_edges] */
    const float*          inv_sigma2;  /* [n_levels] ORBextractor::GetInverseScaleSigmaSquares() */
    const double*         cam_K;       /* [n_cams][4] */
    const double*         cam_ext;     /* [n_cams][12] */
    const double*         cam_adj;     /* [n_cams][36] */
} orbba_problem_f32_t;

typedef struct {
    double  initial_chi2, final_chi2, final_lambda;   /* robust chi2 before, chi2 after the last accepted step, last lambda */
    int32_t iterations, trials, outliers;             /* LM outer iterations, linear solves, edges flagged at the end */
    int32_t status;                                    /* ORB_OK or ORB_E_ABORTED */
} orbba_stats_t;

typedef struct orbba orbba_t;

int  orbba_create(orbba_t** out, int device, int max_problems);
void orbba_destroy(orbba_t*);
int  orbba_set_stream(orbba_t*, void* cuda_stream);
/* optional second stream for orbba_upload (host->device copies + index kernels); orbba_run waits for the upload through an event.
 * Lets the upload of one handle overlap with the run of another handle that shares the compute stream.  NULL = use the compute stream. */
int  orbba_set_copy_stream(orbba_t*, void* cuda_stream);
int  orbba_synchronize(orbba_t*);
long long orbba_launch_count(const orbba_t*);

/* Optimizer::LocalBundleAdjustment(pKF, pbStopFlag, pMap, fixId) on one flattened problem: optimize(its1 = 5) with
 * Huber(delta = sqrt(5.991)), edges with chi2 > chi2_th (5.991) or non-positive depth leave the graph and the kernel is
 * dropped, optimize(its2 = 10), final outlier classification (the observations the caller erases, src/Optimizer.cc:641-660).
 * HOST buffers, synchronous.  `stop` = pbStopFlag (may be NULL): polled while the GPU works; returns ORB_E_ABORTED
 * with the inputs unchanged if it was already set on entry (src/Optimizer.cc:582-584).
 *   poses_out [n_poses][12], points_out [n_points][3], edge_outlier [n_edges], stats: any may be NULL. */
int  orbba_local(orbba_t*, const orbba_problem_t* problem, int its1, int its2, double huber_delta, double chi2_th,
                 const volatile uint8_t* stop, double* poses_out, double* points_out, uint8_t* edge_outlier, orbba_stats_t* stats);
/* the same from the compact form (orbba_problem_f32_t): what an adaptor reads out of KeyFrame / MapPoint without widening anything */
int  orbba_local_f32(orbba_t*, const orbba_problem_f32_t* problem, int its1, int its2, double huber_delta, double chi2_th,
                     const volatile uint8_t* stop, double* poses_out, double* points_out, uint8_t* edge_outlier, orbba_stats_t* stats);
/* Optimizer::BundleAdjustment(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust): one optimize(iterations),
 * Huber(huber_delta) when > 0, no outlier pass. */
int  orbba_global(orbba_t*, const orbba_problem_t* problem, int iterations, double huber_delta, const volatile uint8_t* stop,
                  double* poses_out, double* points_out, orbba_stats_t* stats);

/* Batched form: n independent problems (one per keyframe / per sequence) solved concurrently, one persistent CTA each.
 * upload = flatten + index + host->device; run = asynchronous on the handle's stream, always restarts from the uploaded
 * estimates (its2 < 0: single round); download = results of problem p (synchronises). */
int  orbba_upload(orbba_t*, const orbba_problem_t* problems, int n);
int  orbba_upload_f32(orbba_t*, const orbba_problem_f32_t* problems, int n);   /* same, from the compact form */
int  orbba_run(orbba_t*, int its1, int its2, double huber_delta, double chi2_th);
int  orbba_download(orbba_t*, int p, double* poses_out, double* points_out, uint8_t* edge_outlier, orbba_stats_t* stats);
/* results of every problem of the last run, concatenated in upload order: poses_out [sum n_poses][12], points_out
 * [sum n_points][3], edge_outlier [sum n_edges], stats [n].  Any pointer may be NULL.  Synchronises. */
int  orbba_download_batch(orbba_t*, double* poses_out, double* points_out, uint8_t* edge_outlier, orbba_stats_t* stats);
int  orbba_profile(orbba_t*, int enable);
int  orbba_stage_ms(orbba_t*, double* ms1, int* calls);
/* device time of the kernels of the LM step {k_lin, k_build (both only on the first step of a round), k_land (linearise + landmark
 * blocks + per-edge records), k_pairs (Schur products per pose pair), k_solve (reduced camera system), k_back (back-substitution +
 * trial errors + LM decision)}, summed over the steps recorded since the last call; *steps = LM steps summed. */
int  orbba_kernel_ms(orbba_t*, double* ms6, int* steps);

/* ------------------------------------------------------------------------------------------------
 * Optimizer::PoseOptimization(pFrame) (src/Optimizer.cc:250-405): pose-only LM of one frame against the map points its keypoints hold,
 * four rounds of optimize(10) with inlier / outlier re-classification.  Batched over frames (one per tracked sequence), one persistent
 * CTA per frame.  The adaptor lists, in keypoint order, every keypoint i with mvpMapPoints[i] != NULL.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const double*  pose;         /* [12]        Frame::mTcw, row-major 3x4 (CV_32F widened) */
    int32_t        n_obs;        /* nInitialCorrespondences */
    const double*  Xw;           /* [n_obs][3]  MapPoint::GetWorldPos() */
    const double*  obs;          /* [n_obs][2]  mvTotalKeysUn[i].pt */
    const double*  inv_sigma2;   /* [n_obs]     mvInvLevelSigma2[octave] */
    const int32_t* cam;          /* [n_obs]     keypointToCam[i] */
    int32_t        n_cams;
    const double*  cam_K;        /* [n_cams][4]  fx fy cx cy */
    const double*  cam_ext;      /* [n_cams][12] mvExtrinsics[c] */
    const double*  cam_adj;      /* [n_cams][36] mvExtAdj[c] */
} orbpo_frame_t;
/*   poses_out [n][12]: the pose SetPose() receives; outlier: mvbOutlier of the listed keypoints, frames concatenated [sum n_obs];
 *   n_inliers [n]: the return value (nInitialCorrespondences - nBad; 0 and nothing touched below 3 correspondences);
 *   lm_counts [n][2] (may be NULL): LM iterations and trials summed over the rounds.  HOST buffers, synchronous. */
int  orbba_pose_optimization(orbba_t*, const orbpo_frame_t* frames, int n, double* poses_out, uint8_t* outlier, int32_t* n_inliers,
                             int32_t* lm_counts);

/* ------------------------------------------------------------------------------------------------
 * GlobalBundleAdjustemnt for large maps, on one GPU or landmark-partitioned over the GPUs of a node (BASELINE.json configs[4]:
 * 2000 key frames x 2 cameras, 200k map points).  One process per GPU; key-frame poses are replicated, rank r owns the map
 * points `index mod world == r` and their edges; per LM trial ONE NCCL all-reduce (sum, FP64) of the reduced camera system
 * [Hschur | bschur] (block skyline: only the envelope of the reduced camera system is stored, built and shipped) and one of three
 * scalars; the hand-written skyline LDL^T of the reduced system and the LM policy are replicated.
 *   rank 0:      orbba_dist_unique_id(id)  -> broadcast the 128 bytes to the other ranks by any means (torch.distributed, MPI, a file)
 *   every rank:  orbba_dist_create(&h, device, rank, world, id)  (world == 1: id may be NULL, no NCCL needed)
 *                orbba_dist_optimize(h, shard, ...)   collective
 * ---------------------------------------------------------------------------------------------- */
typedef struct orbgba orbgba_t;
int  orbba_dist_unique_id(uint8_t* id128);
int  orbba_dist_create(orbgba_t** out, int device, int rank, int world, const uint8_t* id128);
void orbba_dist_destroy(orbgba_t*);
long long orbba_dist_launch_count(const orbgba_t*);
/* Optimizer::BundleAdjustment(vpKFs, vpMP, nIterations, pbStopFlag, nLoopKF, bRobust) (src/Optimizer.cc:70-248) on this rank's shard:
 * `shard` holds ALL poses (identical on every rank) and this rank's points / edges (edge_point indexes the local point array).
 * poses_out [n_poses][12] (identical on every rank), points_out [n_points][3] (local points).  huber_delta <= 0: no kernel. */
int  orbba_dist_optimize(orbgba_t*, const orbba_problem_t* shard, int iterations, double huber_delta, const volatile uint8_t* stop,
                         double* poses_out, double* points_out, orbba_stats_t* stats);
/* device milliseconds spent in the [Hschur | bschur] all-reduce and in the dense solve during the last optimize call, and the
 * bytes this rank contributed to all-reduces */
int  orbba_dist_timing(const orbgba_t*, double* allreduce_ms, double* solve_ms, double* allreduce_bytes);
/* device milliseconds of the whole LM loop of the last optimize call (events around it) and the 6x6 blocks of the reduced camera
 * system's skyline (what the all-reduce ships: 288 bytes each) */
int  orbba_dist_loop_ms(const orbgba_t*, double* loop_ms, long long* skyline_blocks);
/* segments the last optimize call cut the reduced camera system into (0: single-CTA band / envelope factorisation).  Long narrow
 * bands are substructured: one CTA per segment, a reduced band system for the separators.  ORBGBA_SEGMENTS=n overrides (0: off). */
int  orbba_dist_segments(const orbgba_t*);

#ifdef __cplusplus
}
#endif
#endif /* ORBSLAM2_DUALCAM_B200_H */
