// png++/png.hpp -- stand-in for png++ (test infrastructure, part of oracle/; see shim/opencv2/opencv.hpp).
//
// The reference's dataset reader (core/read_data.cpp:36-61) loads frames through png::image.  The hot path under
// test starts AFTER the coordinate CNN, so the pixel content of a frame is irrelevant: every image "loaded" here is
// a black frame of the configured size.  The shim calls shim_png_on_load(path) first, which is how the harness
// learns which frame the reference's driver is about to process (oracle/ref_harness).
#pragma once
#include <string>

extern "C" void shim_png_on_load(const char* path, int* width, int* height);

namespace png {
template <typename T> struct basic_rgb_pixel {
    T red, green, blue;
    basic_rgb_pixel() : red(0), green(0), blue(0) {}
};
template <typename Pixel> class image {
public:
    explicit image(const std::string& path) : w_(640), h_(480) { shim_png_on_load(path.c_str(), &w_, &h_); }
    int get_width() const { return w_; }
    int get_height() const { return h_; }
    Pixel get_pixel(int, int) const { return Pixel(); }

private:
    int w_, h_;
};
}  // namespace png
