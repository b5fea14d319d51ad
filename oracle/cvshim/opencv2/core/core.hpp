// Minimal stand-in for the three OpenCV headers /root/reference/CSNet_training/SalMetric/src/sal_metric.cpp includes, so the
// UNMODIFIED reference source compiles here without OpenCV (test infrastructure: oracle/_ref/salmetric pins oracle/salmetric.py).
// The reference only uses: cv::Mat {rows, cols, at<float>(h, w), convertTo(Mat&, CV_32F)} and cv::imread(path, 0) (8-bit grey).
// Here imread reads binary PGM ("P5") files — the tests write their 8-bit maps in that format; the arithmetic under test
// (sal_metric.cpp:86-120,164-185) is the reference's own, compiled from where it lies.
#ifndef CSNET_ORACLE_CVSHIM_CORE_HPP
#define CSNET_ORACLE_CVSHIM_CORE_HPP
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CV_32F 5

namespace cv {

class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c) : rows(r), cols(c), data_((size_t)r * c, 0.f) {}
  template <typename T> T& at(int h, int w) { return reinterpret_cast<T&>(data_[(size_t)h * cols + w]); }
  template <typename T> const T& at(int h, int w) const { return reinterpret_cast<const T&>(data_[(size_t)h * cols + w]); }
  // values are held as float from the start (8-bit integers are exact in float), so the conversion is the identity
  void convertTo(Mat& dst, int /*type*/) const { if (&dst != this) dst = *this; }
 private:
  std::vector<float> data_;
};

inline Mat imread(const std::string& path, int /*flags: 0 = greyscale*/) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return Mat();
  int w = 0, h = 0, maxv = 0;
  char magic[3] = {0, 0, 0};
  if (std::fscanf(f, "%2s %d %d %d", magic, &w, &h, &maxv) != 4 || std::strcmp(magic, "P5") != 0 || maxv != 255) { std::fclose(f); return Mat(); }
  std::fgetc(f);                                   // the single whitespace byte after the header
  Mat m(h, w);
  std::vector<unsigned char> row((size_t)w);
  for (int y = 0; y < h; ++y) {
    if (std::fread(row.data(), 1, (size_t)w, f) != (size_t)w) break;
    for (int x = 0; x < w; ++x) m.at<float>(y, x) = (float)row[(size_t)x];
  }
  std::fclose(f);
  return m;
}

}  // namespace cv
#endif
