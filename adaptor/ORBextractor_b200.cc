// adaptor/ORBextractor_b200.cc -- drop-in bodies for ORB_SLAM2::ORBextractor on top of liborbslam2_dualcam_b200.so.
//
// Replaces src/ORBextractor.cc of the reference (constructor :410-470, operator() :1043-1105 and everything they call); the class
// declaration include/ORBextractor.h:47-108 is used UNCHANGED -- the library handle of an extractor lives in a side table keyed
// by the object, so no member has to be added.  Build:  g++ -c -I<reference>/include -I<repo>/include adaptor/ORBextractor_b200.cc
// and link liborbslam2_dualcam_b200.so instead of compiling src/ORBextractor.cc.
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>

#include <map>
#include <mutex>

#include "ORBextractor.h"                 // the reference's own header (or adaptor/shim/ORBextractor.h where OpenCV is absent)
#include "orbslam2_dualcam_b200.h"

namespace ORB_SLAM2 {

namespace {
std::mutex g_mu;
std::map<const ORBextractor*, orbx_t*> g_handles;     // one orbx_t per extractor object (the reference has one extractor per camera)
[[noreturn]] void die(const char* where) {
    fprintf(stderr, "ORBextractor (B200): %s: %s\n", where, orb_last_error());
    exit(-1);                              // the reference's own failure mode (exit(-1) in System / Tracking)
}
}  // namespace

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
    // The scale tables the getters return (include/ORBextractor.h:63-83) are filled from the library at the first operator() call, when
    // the image size is known; until then they hold the same float products the reference computes (src/ORBextractor.cc:415-431).
    mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) { mvScaleFactor[i] = mvScaleFactor[i - 1] * (float)scaleFactor; mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i]; }
    for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
    mnFeaturesPerLevel.resize(nlevels);
    umax.resize(16);
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask: ignored, as in the reference*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors) {
    if (_image.empty()) return;                                                // src/ORBextractor.cc:1046-1047
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);                                           // :1050
    orbx_t* h = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_mu);
        orbx_t*& slot = g_handles[this];
        if (!slot) {
            if (orbx_create(&slot, /*device*/ 0, image.cols, image.rows, /*cameras*/ 1, /*max_frames*/ 1, nfeatures, (float)scaleFactor, nlevels, iniThFAST,
                            minThFAST) != ORB_OK)
                die("orbx_create");
            if (orbx_get_tables(slot, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(), mnFeaturesPerLevel.data(),
                                umax.data()) != ORB_OK)
                die("orbx_get_tables");
        }
        h = slot;
    }
    static_assert(sizeof(cv::KeyPoint) == sizeof(orb_keypoint_t), "cv::KeyPoint is 7 x 4 bytes");
    const int cap = orbx_max_keypoints(h);
    _keypoints.resize(cap);
    cv::Mat desc(cap, 32, CV_8U);
    int32_t n = 0;
    if (orbx_extract(h, image.data, 1, image.step, reinterpret_cast<orb_keypoint_t*>(_keypoints.data()), desc.data, &n, cap) != ORB_OK) die("orbx_extract");
    _keypoints.resize(n);
    if (n == 0) _descriptors.release();                                        // :1072-1073
    else desc.rowRange(0, n).copyTo(_descriptors);
}

}  // namespace ORB_SLAM2
