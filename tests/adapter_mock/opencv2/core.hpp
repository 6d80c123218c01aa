// mock of the OpenCV declarations the adapter uses (opencv2/core.hpp)
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#define CV_8U 0
namespace cv {
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
struct _OutputArray;
struct Mat {
    unsigned char *data = nullptr;
    int rows = 0, cols = 0;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int type) : rows(r), cols(c), step((size_t)c) { (void)type; }
    bool empty() const { return data == nullptr; }
    Mat rowRange(int a, int b) const { (void)a; (void)b; return *this; }
    Mat clone() const { return *this; }
    void copyTo(const _OutputArray &o) const { (void)o; }
    void copyTo(Mat &o) const { (void)o; }
    void create(int r, int c, int type) { rows = r; cols = c; (void)type; }
    template <class T> T *ptr(int r = 0) { (void)r; return reinterpret_cast<T *>(data); }
};
struct _OutputArray { void release() const {} void create(int, int, int) const {} Mat getMat() const { return Mat(); } };
namespace line_descriptor {
struct KeyLine {  // feature/line_descriptor/descriptor_custom.hpp:105-199
    float angle; int class_id; int octave; Point2f pt; float response, size;
    float startPointX, startPointY, endPointX, endPointY, sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength; int numOfPixels;
    Point2f getStartPoint() const { return Point2f{startPointX, startPointY}; }
    Point2f getEndPoint() const { return Point2f{endPointX, endPointY}; }
};
}  // namespace line_descriptor
}  // namespace cv
