// demo.cpp -- the reference's src/demo.cpp:55-118 against the B200 path: load a model by extension, build a
// PartsBasedDetector<float>, read an image, detect, sort, print.  Images: binary PPM (P6) / PGM (P5), since
// OpenCV's imread is not assumed; with OpenCV present pass a cv::Mat to detect() instead.
//
//   g++ -std=c++17 -Iinclude examples/demo.cpp -Lpartsbaseddetector_b200 -lpbd_b200 -Wl,-rpath,$PWD/partsbaseddetector_b200 -o demo
//   ./demo model.xml|model.yaml|model.mat|model.pbdm image.ppm [thresh]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "pbd_b200.hpp"

using namespace pbd_b200;

struct BinModel : Model {       // .pbdm container (this library's compact copy of the XML fields)
  bool serialize(const std::string& f) const override { check(pbd_model_save_bin(h_, f.c_str())); return true; }
  bool deserialize(const std::string& f) override {
    pbd_model* m = nullptr;
    const int rc = pbd_model_load_bin(f.c_str(), &m);
    if (rc == PBD_E_IO) return false;
    check(rc);
    pbd_model_free(h_);
    h_ = m;
    return true;
  }
};

static bool read_pnm(const char* path, std::vector<uint8_t>& px, int& w, int& h, int& c) {
  std::ifstream f(path, std::ios::binary);
  std::string magic;
  int maxv = 0;
  if (!(f >> magic >> w >> h >> maxv) || (magic != "P6" && magic != "P5") || maxv != 255) return false;
  f.get();
  c = magic == "P6" ? 3 : 1;
  px.resize((size_t)w * h * c);
  f.read(reinterpret_cast<char*>(px.data()), (std::streamsize)px.size());
  if (!f) return false;
  if (c == 3) for (size_t i = 0; i < px.size(); i += 3) std::swap(px[i], px[i + 2]);   // RGB -> BGR as cv::imread
  return true;
}

int main(int argc, char** argv) {
  if (argc < 3) { printf("Usage: demo model_file image_file [thresh]\n"); return -1; }              // src/demo.cpp:58-61
  const std::string mfile = argv[1];
  const std::string ext = mfile.substr(mfile.find_last_of('.') + 1);
  Model* model = nullptr;
  if (ext == "xml" || ext == "yaml" || ext == "yml") model = new FileStorageModel;                   // :64-72
  else if (ext == "mat") model = new MatlabIOModel;                                                  // :69-70
  else if (ext == "pbdm") model = new BinModel;
  else { printf("Unsupported model format: %s\n", ext.c_str()); return -2; }                         // :73-76
  try {
    if (!model->deserialize(mfile)) { printf("Error deserializing file\n"); return -3; }             // :77-81
    PartsBasedDetector<float> pbd;                                                                   // :85-86
    pbd.distributeModel(*model);
    if (argc > 3) pbd.setOption("thresh", atof(argv[3]));
    std::vector<uint8_t> px;
    int w, h, c;
    if (!read_pnm(argv[2], px, w, h, c)) { printf("Could not find or open the image\n"); return -4; }   // :91-94
    vectorCandidate candidates;
    pbd.detect(Mat(h, w, c, px.data()), Mat(), candidates);                                          // :103
    printf("Number of candidates: %zu\n", candidates.size());                                        // :104
    if (!candidates.empty()) {
      Candidate::sort(candidates);                                                                   // :111
      const Candidate& best = candidates[0];
      const Rect bb = best.boundingBox();
      printf("best: score %.6f level %d box (%d,%d,%d,%d)\n", best.score(), best.level, bb.x, bb.y, bb.width, bb.height);
    }
  } catch (const Error& e) {
    printf("pbd_b200 error %d: %s\n", e.code, e.what());
    delete model;
    return -5;
  }
  delete model;
  return 0;
}
