// plugin_demo.cpp -- drives the two stage plug-ins exactly as the reference's PartsBasedDetector<T>::detect does
// (src/PartsBasedDetector.cpp:69-83): features_->pyramid(im, pyramid); convolution_engine_->pdf(pyramid, responses), with
// CudaHOGFeatures / CudaConvolutionEngine behind the IFeatures / IConvolutionEngine pointers, and dumps every feature map and
// response map for the tests (tests/test_gpu_parity.py::test_stage_plugins_match_oracle).
//
//   g++ -std=c++17 -Iinclude -Ioracle/ref_shim examples/plugin_demo.cpp -Lpartsbaseddetector_b200 -lpbd_b200 -o plugin_demo
//   (with OpenCV installed drop -Ioracle/ref_shim; inside the reference tree add -I<reference>/include to use its own interface headers)
//   ./plugin_demo model.pbdm frame.raw H W C out.bin [f64]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <vector>

#include "pbd_b200_plugins.hpp"

using namespace pbd_b200;

int main(int argc, char** argv) {
  if (argc < 7) { printf("Usage: plugin_demo model.pbdm frame.raw H W C out.bin [f64]\n"); return 255; }
  const bool f64 = argc > 7 && !strcmp(argv[7], "f64");
  try {
    pbd_model* m = nullptr;
    check(pbd_model_load_bin(argv[1], &m));
    int32_t hdr[8]; float th;
    check(pbd_model_header(m, hdr, &th));                      // interval, sbin, norient, flen, nfilters, ...
    const int H = atoi(argv[3]), W = atoi(argv[4]), C = atoi(argv[5]);
    std::vector<uint8_t> px((size_t)H * W * C);
    std::ifstream(argv[2], std::ios::binary).read(reinterpret_cast<char*>(px.data()), (std::streamsize)px.size());
    cv::Mat im(H, W, CV_MAKETYPE(CV_8U, C), px.data());

    // distributeModel(), src/PartsBasedDetector.cpp:107-118
    std::unique_ptr<IFeatures> features;
    if (f64) features.reset(new CudaHOGFeatures<double>(hdr[1], hdr[0], hdr[3], hdr[2]));
    else features.reset(new CudaHOGFeatures<float>(hdr[1], hdr[0], hdr[3], hdr[2]));
    std::unique_ptr<IConvolutionEngine> convolution_engine(new CudaConvolutionEngine(f64 ? CV_64F : CV_32F, hdr[3]));
    vectorMat filters;
    for (int i = 0; i < hdr[4]; ++i) {
      int32_t rows, kw; const double* w = nullptr;
      check(pbd_model_filter(m, i, &rows, &kw, &w));
      cv::Mat f(rows, kw * hdr[3], CV_64F);
      memcpy(f.data, w, sizeof(double) * (size_t)rows * kw * hdr[3]);
      cv::Mat ft;
      f.convertTo(ft, f64 ? CV_64F : CV_32F);                   // `filters[i].convertTo(filters[i], DataType<T>::type)`, :115-117
      filters.push_back(ft);
    }
    convolution_engine->setFilters(filters);
    pbd_model_free(m);

    // detect(), :72-83
    vectorMat pyramid;
    features->pyramid(im, pyramid);
    vector2DMat responses;
    convolution_engine->pdf(pyramid, responses);

    std::ofstream out(argv[6], std::ios::binary);
    const vectorf scales = features->scales();
    const int32_t n = (int32_t)features->nscales(), nf = hdr[4], es = f64 ? 8 : 4;
    out.write((const char*)&n, 4); out.write((const char*)&nf, 4); out.write((const char*)&es, 4);
    for (int l = 0; l < n; ++l) {
      const int32_t oh = pyramid[l].rows, owf = pyramid[l].cols;
      out.write((const char*)&oh, 4); out.write((const char*)&owf, 4); out.write((const char*)&scales[l], 4);
      for (int y = 0; y < oh; ++y) out.write((const char*)pyramid[l].ptr(y), (std::streamsize)owf * es);
      for (int f = 0; f < nf; ++f) {
        const cv::Mat& r = responses[l][f];
        if (r.rows != oh || r.cols * hdr[3] != owf || r.depth() != (f64 ? CV_64F : CV_32F)) { printf("bad response shape\n"); return 250; }
        for (int y = 0; y < r.rows; ++y) out.write((const char*)r.ptr(y), (std::streamsize)r.cols * es);
      }
    }
    printf("%d levels, %d filters\n", n, nf);
    // error behaviour: pdf() before setFilters() is a state error, as the reference's contract demands ("must be called before pdf()")
    CudaConvolutionEngine fresh(CV_32F, hdr[3]);
    try { vector2DMat r2; fresh.pdf(pyramid, r2); printf("pdf without filters did not fail\n"); return 249; } catch (const Error& e) { if (e.code != PBD_E_STATE) return 248; }
  } catch (const Error& e) {
    printf("error %d: %s\n", e.code, e.what());
    return e.code == PBD_E_CUDA ? 251 : 252;
  }
  return 0;
}
