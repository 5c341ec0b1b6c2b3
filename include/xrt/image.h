// xrt/image.h — Image of the drop-in API (reference image.h:11-150): fp32 RGB framebuffer, row-major,
// index = j + width*i. The OpenCV export (image.h:116-136) is out of scope; PPM/PFM writers remain.
#pragma once
#include <fstream>
#include <string>
#include <vector>
#ifdef XRT_WITH_OPENCV
#include <opencv2/opencv.hpp>
#endif
#include "geometry.h"

class Image {
public:
    struct ImageIdx { int i; int j; };

    Image(uint32_t width, uint32_t height) : width(width), height(height) { pixels.resize(size_t(width) * height); }
    uint32_t getWidth() const { return width; }
    uint32_t getHeight() const { return height; }
    Vec3f getPixel(uint32_t i, uint32_t j) const { return pixels[getIndex(i, j)]; }
    void addPixel(uint32_t i, uint32_t j, const Vec3f& rgb) { pixels[getIndex(i, j)] += rgb; }
    void setPixel(uint32_t i, uint32_t j, const Vec3f& rgb) { pixels[getIndex(i, j)] = rgb; }
    Image& operator*=(const Vec3f& rgb) { for (auto& p : pixels) p = p * rgb; return *this; }
    Image& operator/=(const Vec3f& rgb) { for (auto& p : pixels) p = p / rgb; return *this; }
    void gammaCorrection(const float gamma)
    {
        for (auto& p : pixels) for (int c = 0; c < 3; ++c) p[c] = std::pow(p[c], 1.0f / gamma);
    }
    void writePPM(const std::string& filename) const
    {
        std::ofstream f(filename);
        f << "P3\n" << width << " " << height << "\n255\n";
        for (const auto& p : pixels) {
            for (int c = 0; c < 3; ++c) f << std::clamp(static_cast<uint32_t>(255.0f * p[c]), 0u, 255u) << (c == 2 ? "\n" : " ");
        }
    }
    void writePFM(const std::string& filename) const
    {
        std::ofstream f(filename, std::ios::binary);
        f << "PF\n" << width << " " << height << "\n-1.0\n";
        for (int i = int(height) - 1; i >= 0; --i) f.write(reinterpret_cast<const char*>(&pixels[size_t(i) * width]), sizeof(Vec3f) * width);
    }
#ifdef XRT_WITH_OPENCV
    // BGR 8-bit export exactly as the reference's Image::writeMat (image.h:116-136); only with an OpenCV (or stub) cv::Mat
    cv::Mat writeMat()
    {
        cv::Mat mat(height, width, CV_8UC3);
        for (uint32_t i = 0; i < height; ++i) {
            auto prow = mat.ptr<uchar>(i);
            for (uint32_t j = 0; j < width; ++j) {
                const Vec3f rgb = getPixel(i, j);
                for (int c = 0; c < 3; ++c)
                    prow[3 * j + (2 - c)] = uint8_t(std::clamp(static_cast<uint32_t>(255.0f * rgb[c]), 0u, 255u));
            }
        }
        return mat;
    }
#endif
    // additive: contiguous W*H*3 floats, the layout xrtg_render writes
    float* data() { return &pixels[0][0]; }
    const float* data() const { return pixels[0].getPtr(); }

private:
    uint32_t getIndex(uint32_t i, uint32_t j) const { return j + width * i; }
    uint32_t width, height;
    std::vector<Vec3f> pixels;
};
