// adaptor/harness_extractor.cc -- runs the adaptor's ORB_SLAM2::ORBextractor (adaptor/ORBextractor_b200.cc, compiled against the
// reference's class declaration) exactly as Frame::ExtractORB does (src/Frame.cc:210-213) and compares it with a direct orbx_extract call.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ORBextractor.h"
#include "orbslam2_dualcam_b200.h"

int main() {
    const int W = 640, H = 480;
    cv::Mat im(H, W, CV_8UC1);
    unsigned s = 777u;
    memset(im.data, 0, (size_t)W * H);
    for (int b = 0; b < 500; b++) {
        s = s * 1664525u + 1013904223u; const int x0 = (s >> 8) % W;
        s = s * 1664525u + 1013904223u; const int y0 = (s >> 8) % H;
        s = s * 1664525u + 1013904223u; const int w = 8 + (s >> 8) % 70;
        s = s * 1664525u + 1013904223u; const int h = 8 + (s >> 8) % 70;
        s = s * 1664525u + 1013904223u; const int g = (s >> 8) % 256;
        for (int y = y0; y < y0 + h && y < H; y++)
            for (int x = x0; x < x0 + w && x < W; x++) im.at<unsigned char>(y, x) = (unsigned char)((g + (y - y0)) & 255);
    }
    ORB_SLAM2::ORBextractor extractor(1000, 1.2f, 8, 20, 7);       // src/Tracking.cc:204-207
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    extractor(im, cv::Mat(), keys, descriptors);                    // (*mpORBextractor)(im, cv::Mat(), mvKeys[c], mvDescriptors[c])
    if (keys.empty() || descriptors.rows != (int)keys.size() || descriptors.cols != 32) { fprintf(stderr, "adaptor returned %zu keys, %d x %d descriptors\n", keys.size(), descriptors.rows, descriptors.cols); return 2; }
    // the same image straight through the C-ABI
    orbx_t* h = nullptr;
    if (orbx_create(&h, 0, W, H, 1, 1, 1000, 1.2f, 8, 20, 7) != ORB_OK) { fprintf(stderr, "%s\n", orb_last_error()); return 1; }
    const int cap = orbx_max_keypoints(h);
    std::vector<orb_keypoint_t> k2(cap);
    std::vector<uint8_t> d2((size_t)cap * 32);
    int32_t n = 0;
    if (orbx_extract(h, im.data, 1, im.step, k2.data(), d2.data(), &n, cap) != ORB_OK) { fprintf(stderr, "%s\n", orb_last_error()); return 1; }
    orbx_destroy(h);
    bool same = n == (int)keys.size() && memcmp(k2.data(), keys.data(), sizeof(orb_keypoint_t) * n) == 0;
    for (int i = 0; same && i < n; i++) same = memcmp(descriptors.ptr<unsigned char>(i), &d2[(size_t)i * 32], 32) == 0;
    const std::vector<float> sf = extractor.GetScaleFactors();
    printf("ORBextractor adaptor: %zu keypoints, %d levels, scale[7] = %.6f, identical to orbx_extract: %s\n", keys.size(), extractor.GetLevels(), sf[7], same ? "yes" : "NO");
    return same ? 0 : 3;
}
