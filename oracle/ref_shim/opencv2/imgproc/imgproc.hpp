// Minimal stand-in for <opencv2/imgproc/imgproc.hpp> -- TEST INFRASTRUCTURE (oracle/_ref), see ../core/core.hpp.
// Declarations only; cv::resize (INTER_LINEAR, 8U) and cv::pyrDown are defined in oracle/ref_driver.cpp from the oracle's
// bit-exact restatements of the OpenCV functions (pinned against cv2 by tests/test_oracle_pins.py).
#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
enum { INTER_LINEAR = 1, BORDER_CONSTANT = 0, BORDER_DEFAULT = 4 };
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void pyrDown(const Mat& src, Mat& dst, const Size& dstsize = Size(), int borderType = BORDER_DEFAULT);
Mat getGaussianKernel(int ksize, double sigma, int ktype = CV_64F);
void filter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernel, Point anchor = Point(-1, -1), double delta = 0, int borderType = BORDER_DEFAULT);
template <typename T, int m, int n> inline void filter2D(const Mat& src, Mat& dst, int ddepth, const Matx<T, m, n>&) { (void)src; (void)dst; (void)ddepth; throw Exception(-213, "shim: filter2D(Matx) is not implemented"); }
}  // namespace cv
