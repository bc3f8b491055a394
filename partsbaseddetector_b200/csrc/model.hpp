// model.hpp -- host-side model container and (de)serialisers.
// Mirrors the reference `Model` fields (include/Model.hpp:49-122) and replaces
// FileStorageModel (src/FileStorageModel.cpp:42-159) without OpenCV.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pbd {

struct Part {
  int parentid = -1;
  std::vector<int> filterid, biasid, defid;
};

struct Model {
  std::string name;
  int interval = 0;          // XML "interval" (reference keeps it in Model::nscales_)
  float thresh = 0.f;
  int sbin = 0, norient = 18, flen = 32;
  std::vector<int> frows, fkw;                 // filter i is a Mat of frows[i] x (fkw[i]*flen), HWC
  std::vector<std::vector<double>> filters;    // as stored (f64); converted to T at distributeModel
  std::vector<float> biasw;
  std::vector<int> anchors;                    // x,y pairs (0-based, parent-relative)
  std::vector<float> defs;                     // 4 per entry: w0 (x^2), w1 (x), w2 (y^2), w3 (y)
  std::vector<std::vector<Part>> comps;

  int nfilters() const { return (int)filters.size(); }
  int ndefs() const { return (int)defs.size() / 4; }
  int ncomponents() const { return (int)comps.size(); }
  // throws std::runtime_error with a description if any index is out of range
  void validate() const;
};

// opencv_storage XML (what cv::FileStorage writes/reads for the reference's schema)
void load_xml(const std::string& path, Model& m);     // throws IoError / FormatError
void save_xml(const Model& m, const std::string& path);
// the same schema in cv::FileStorage's YAML flavour (%YAML:1.0); load_storage picks the parser from the content, save_storage
// the writer from the extension (.yml / .yaml -> YAML, anything else -> XML) as cv::FileStorage::open does
void load_yaml(const std::string& path, Model& m);
void save_yaml(const Model& m, const std::string& path);
void load_storage(const std::string& path, Model& m);
void save_storage(const Model& m, const std::string& path);
// MATLAB Level-5 MAT-file holding the training code's `model` struct (reference MatlabIOModel::deserialize,
// src/MatlabIOModel.cpp:71-188; native reader in matfile.cpp, no cvmatio)
void load_mat(const std::string& path, Model& m);     // throws IoError / FormatError
// compact little-endian container of the same fields ("PBDM" v1)
void load_bin(const std::string& path, Model& m);
void save_bin(const Model& m, const std::string& path);

struct IoError : std::exception {
  std::string msg;
  explicit IoError(std::string s) : msg(std::move(s)) {}
  const char* what() const noexcept override { return msg.c_str(); }
};
struct FormatError : std::exception {
  std::string msg;
  explicit FormatError(std::string s) : msg(std::move(s)) {}
  const char* what() const noexcept override { return msg.c_str(); }
};

}  // namespace pbd
