// Minimal stand-in for the part of OpenCV's core module that the adaptor sources touch (cv::Mat, cv::KeyPoint, cv::Point_,
// cv::InputArray / cv::OutputArray).  TEST INFRASTRUCTURE: it exists so that adaptor/*.cc can be compiled -- and the extractor adaptor
// run -- in an image without OpenCV headers; a real build uses OpenCV's own headers and never sees this file.
#pragma once
#include <stdint.h>
#include <string.h>

#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_64F 6

namespace cv {

template <class T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct KeyPoint {            // the 28-byte layout of cv::KeyPoint: pt, size, angle, response, octave, class_id
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

class _OutputArray;
class Mat {
public:
    int rows, cols;
    size_t step;
    unsigned char* data;
    Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    static size_t elem(int type) { return type == CV_8U ? 1 : type == CV_32F ? 4 : 8; }
    void create(int r, int c, int type) {
        type_ = type; rows = r; cols = c; step = (size_t)c * elem(type);
        buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * step);
        data = buf_->data();
    }
    int type() const { return type_; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; r++) memcpy(m.data + r * m.step, data + r * step, (size_t)cols * elem(type_)); return m; }
    template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <class T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    inline void copyTo(const _OutputArray& dst) const;
private:
    int type_;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};

class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(&m) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
private:
    const Mat* m_;
};
class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    Mat getMat() const { return *m_; }
    Mat& getMatRef() const { return *m_; }
    void release() const { m_->release(); }
private:
    Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline void Mat::copyTo(const _OutputArray& dst) const { dst.getMatRef() = clone(); }

}  // namespace cv
