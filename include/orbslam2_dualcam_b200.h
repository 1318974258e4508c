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

#ifdef __cplusplus
}
#endif
#endif /* ORBSLAM2_DUALCAM_B200_H */
