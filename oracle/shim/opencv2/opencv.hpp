// Stub of the OpenCV subset image.h:116-136 touches (cv::Mat(int,int,int), ptr<uchar>, CV_8UC3).
// Display/IO only — carries no path arithmetic. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <vector>
typedef unsigned char uchar;
#define CV_8UC3 16
namespace cv {
class Mat {
public:
    Mat() = default;
    Mat(int rows, int cols, int /*type*/) : rows(rows), cols(cols), data(size_t(rows) * cols * 3) {}
    template <typename T> T* ptr(int i) { return reinterpret_cast<T*>(data.data() + size_t(i) * cols * 3); }
    int rows = 0, cols = 0;
    std::vector<uchar> data;
};
} // namespace cv
