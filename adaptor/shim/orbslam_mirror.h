// Declarations of the reference classes as far as the adaptor sources use them, for compiling adaptor/*.cc where the reference's own
// headers cannot be parsed (they pull in OpenCV, Eigen, DBoW2 and g2o).  TEST INFRASTRUCTURE: every member is declared with the
// reference's name and type and cites where the reference declares it (paths under /root/reference); a real build includes the
// reference's headers instead (define ORB_B200_USE_REFERENCE_HEADERS) and never sees this file.
#pragma once
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <unordered_map>
#include <vector>

#include "opencv2/core/core.hpp"

namespace ORB_SLAM2 {

class MapPoint; class KeyFrame; class Frame; class Map; class Cameras;
typedef std::shared_ptr<MapPoint> MapPointPtr;      // include/Frame.h:49
typedef std::shared_ptr<KeyFrame> KeyFramePtr;      // include/Frame.h:50
typedef std::shared_ptr<Cameras> CamerasPtr;        // include/Frame.h:54
typedef std::shared_ptr<Map> MapPtr;                // include/FrameDrawer.h:45
typedef std::shared_ptr<Frame> FramePtr;            // include/Initializer.h:32

class Cameras {                                      // include/Cameras.h:13-40
public:
    int getNCameras();
    cv::Mat getExtrinsici(int i);                    // 4x4 CV_32F, rig -> camera i
    cv::Mat getExtrinsicAdji(int i);                 // 6x6 CV_32F
};

class MapPoint {                                     // include/MapPoint.h
public:
    void SetWorldPos(const cv::Mat& Pos);            // :50
    cv::Mat GetWorldPos();                           // :52   3x1 CV_32F
    std::map<KeyFramePtr, size_t> GetObservations(); // :57
    int Observations();
    void EraseObservation(KeyFramePtr pKF);          // :61
    bool isBad();                                    // :67
    cv::Mat GetDescriptor();
    void UpdateNormalAndDepth();                     // :83
    long unsigned int mnId;                          // :92
    float mTrackProjX, mTrackProjY; int mTrackProjCamera; bool mbTrackInView; int mnTrackScaleLevel; float mTrackViewCos;   // :95-102
    long unsigned int mnBALocalForKF;                // :114
};

class KeyFrame {                                     // include/KeyFrame.h
public:
    void SetPose(const cv::Mat& Tcw);                // :68
    cv::Mat GetPose();                               // :70   4x4 CV_32F
    std::vector<KeyFramePtr> GetVectorCovisibleKeyFrames();   // :85
    void EraseMapPointMatch(MapPointPtr pMP);        // :105
    std::vector<MapPointPtr> GetMapPointMatches();   // :108
    bool isBad();                                    // :125
    long unsigned int mnId;                          // :150
    long unsigned int mnBALocalForKF, mnBAFixedForKF;   // :165-166
    const CamerasPtr mpCameras;                      // :183
    const std::vector<cv::KeyPoint> mvTotalKeysUn;   // :198
    std::unordered_map<size_t, int> keypointToCam;   // :200
    const std::vector<float> mvInvLevelSigma2;       // :221
    const std::vector<float> mvfx, mvfy, mvcx, mvcy; // :230-233
};

class Map {                                          // include/Map.h
public:
    std::mutex mMutexMapUpdate;                      // :69
};

class Frame {                                        // include/Frame.h
public:
    int mnCams;
    std::vector<int> mvN;                            // key points per camera
    int totalN;
    std::vector<cv::KeyPoint> mvTotalKeysUn;         // camera-major concatenation
    std::vector<cv::Mat> mvDescriptors;              // per camera [N_c][32] CV_8U
    std::vector<MapPointPtr> mvpMapPoints;           // [totalN]
    std::vector<float> mvMinX, mvMaxX, mvMinY, mvMaxY;   // per camera (ComputeImageBounds)
    int mnScaleLevels;
    std::vector<float> mvScaleFactors;
};

class ORBmatcher {                                   // include/ORBmatcher.h:49-309
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    int SearchByProjection(FramePtr pF, const std::vector<MapPointPtr>& vpMapPoints, const float th = 3);   // :64, src/ORBmatcher.cc:539-624
    static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
protected:
    float mfNNratio;
    bool mbCheckOrientation;
};

class Optimizer {                                    // include/Optimizer.h:50-56
public:
    static void LocalBundleAdjustment(KeyFramePtr pKF, bool* pbStopFlag, MapPtr pMap, size_t fixId);
};

}  // namespace ORB_SLAM2
