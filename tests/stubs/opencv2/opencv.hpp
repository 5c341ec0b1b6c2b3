// Minimal cv:: stub so the reference's example sources (which display through OpenCV highgui) can be COMPILED against
// include/xrt in tests/test_dropin_compile.py. Display only; no path arithmetic.
#pragma once
#include <string>
#include <vector>
typedef unsigned char uchar;
#define CV_8UC3 16
namespace cv {
class Mat {
public:
    Mat() = default;
    Mat(int rows, int cols, int) : rows(rows), cols(cols), data(size_t(rows) * cols * 3) {}
    template <typename T> T* ptr(int i) { return reinterpret_cast<T*>(data.data() + size_t(i) * cols * 3); }
    int rows = 0, cols = 0;
    std::vector<uchar> data;
};
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int) { return 0; }
} // namespace cv
