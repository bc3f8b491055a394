// matfile.cpp -- native MATLAB Level-5 MAT-file reader and the MatlabIOModel loader on top of it.
//
// Replaces the reference's MatlabIOModel::deserialize (src/MatlabIOModel.cpp:71-188), which parses the file with the
// un-vendored cvmatio library (MatlabIO::open/read/find).  Only what that loader needs is implemented: numeric arrays of
// every storage type (MATLAB stores integer-valued doubles in the smallest integer type that holds them), char arrays,
// cell arrays, struct arrays, small data elements, zlib-compressed variables (miCOMPRESSED), both byte orders.
//
// Index conventions follow the reference: cvmatio hands MATLAB's column-major arrays over as row-major cv::Mat, and the
// loader iterates them with Mat::begin/end, i.e. in ROW-MAJOR order of the MATLAB matrix; 1-based ids become 0-based
// (zeroIndex, src/MatlabIOModel.cpp:46-58); a filter w(m, n, c) lands at flat(m, n*C + c) (:112-118).
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "model.hpp"

namespace pbd {
namespace {

enum { miINT8 = 1, miUINT8 = 2, miINT16 = 3, miUINT16 = 4, miINT32 = 5, miUINT32 = 6, miSINGLE = 7, miDOUBLE = 9, miINT64 = 12,
       miUINT64 = 13, miMATRIX = 14, miCOMPRESSED = 15, miUTF8 = 16, miUTF16 = 17, miUTF32 = 18 };
enum { mxCELL = 1, mxSTRUCT = 2, mxOBJECT = 3, mxCHAR = 4, mxSPARSE = 5, mxDOUBLE = 6, mxSINGLE = 7, mxINT8 = 8, mxUINT8 = 9,
       mxINT16 = 10, mxUINT16 = 11, mxINT32 = 12, mxUINT32 = 13, mxINT64 = 14, mxUINT64 = 15 };

struct MatArray {
  std::string name;
  int cls = 0;
  std::vector<int> dims;
  std::vector<double> real;               // numeric classes, column-major as stored
  std::string text;                       // char arrays (row vectors)
  std::vector<MatArray> cells;            // cell arrays: one entry per element (column-major)
  std::vector<std::string> fields;        // struct arrays: field names ...
  std::vector<MatArray> fvals;            // ... and values, element-major: fvals[e * nfields + f]
  // product of the dimensions, saturating at SIZE_MAX (every caller rejects counts above 2^28)
  size_t numel() const {
    size_t n = 1;
    for (int d : dims) {
      if (d != 0 && n > SIZE_MAX / (size_t)d) return SIZE_MAX;
      n *= (size_t)d;
    }
    return dims.empty() ? 0 : n;
  }
  const MatArray* field(size_t elem, const char* fname) const {
    for (size_t f = 0; f < fields.size(); ++f)
      if (fields[f] == fname) return &fvals[elem * fields.size() + f];
    return nullptr;
  }
};

struct Reader {
  const uint8_t* p;
  size_t n;
  bool swap;
  uint32_t u32(size_t off) const {
    if (off + 4 > n) throw FormatError("MAT file truncated");
    uint32_t v; std::memcpy(&v, p + off, 4);
    return swap ? __builtin_bswap32(v) : v;
  }
};

template <typename T>
T load_swapped(const uint8_t* p, bool swap) {
  uint8_t b[sizeof(T)];
  std::memcpy(b, p, sizeof(T));
  if (swap) for (size_t i = 0; i < sizeof(T) / 2; ++i) std::swap(b[i], b[sizeof(T) - 1 - i]);
  T v; std::memcpy(&v, b, sizeof(T));
  return v;
}

// one data element: type, payload pointer/size, offset of the next element
struct Element { uint32_t type; const uint8_t* data; uint32_t size; size_t next; };
Element read_element(const Reader& r, size_t off) {
  const uint32_t w0 = r.u32(off);
  Element e;
  if (w0 >> 16) {                                   // small data element: size in the upper, type in the lower half word
    e.type = w0 & 0xFFFF; e.size = w0 >> 16;
    if (e.size > 4) throw FormatError("MAT file: bad small data element");
    e.data = r.p + off + 4; e.next = off + 8;
    if (e.next > r.n) throw FormatError("MAT file truncated");
  } else {
    e.type = w0; e.size = r.u32(off + 4);
    e.data = r.p + off + 8;
    if (off + 8 + (size_t)e.size > r.n) throw FormatError("MAT file truncated");
    e.next = off + 8 + (e.type == miCOMPRESSED ? (size_t)e.size : (((size_t)e.size + 7) & ~(size_t)7));
  }
  return e;
}

void numeric_to_double(const Element& e, bool swap, std::vector<double>& out) {
  size_t w = 0;
  switch (e.type) {
    case miINT8: case miUINT8: case miUTF8: w = 1; break;
    case miINT16: case miUINT16: case miUTF16: w = 2; break;
    case miINT32: case miUINT32: case miSINGLE: case miUTF32: w = 4; break;
    case miDOUBLE: case miINT64: case miUINT64: w = 8; break;
    default: throw FormatError("MAT file: unsupported numeric storage type " + std::to_string(e.type));
  }
  const size_t cnt = e.size / w;
  out.resize(cnt);
  for (size_t i = 0; i < cnt; ++i) {
    const uint8_t* q = e.data + i * w;
    switch (e.type) {
      case miINT8: out[i] = (double)(int8_t)q[0]; break;
      case miUINT8: case miUTF8: out[i] = (double)q[0]; break;
      case miINT16: out[i] = (double)load_swapped<int16_t>(q, swap); break;
      case miUINT16: case miUTF16: out[i] = (double)load_swapped<uint16_t>(q, swap); break;
      case miINT32: out[i] = (double)load_swapped<int32_t>(q, swap); break;
      case miUINT32: case miUTF32: out[i] = (double)load_swapped<uint32_t>(q, swap); break;
      case miSINGLE: out[i] = (double)load_swapped<float>(q, swap); break;
      case miDOUBLE: out[i] = load_swapped<double>(q, swap); break;
      case miINT64: out[i] = (double)load_swapped<int64_t>(q, swap); break;
      default: out[i] = (double)load_swapped<uint64_t>(q, swap); break;
    }
  }
}

void parse_matrix(const uint8_t* data, size_t size, bool swap, MatArray& a, int depth) {
  if (depth > 16) throw FormatError("MAT file: nesting too deep");
  if (size == 0) return;                             // empty matrix element ([] inside a cell / struct)
  Reader r{data, size, swap};
  Element flags = read_element(r, 0);
  if (flags.type != miUINT32 || flags.size < 8) throw FormatError("MAT file: bad array flags");
  a.cls = (int)(load_swapped<uint32_t>(flags.data, swap) & 0xFF);
  Element dims = read_element(r, flags.next);
  if (dims.type != miINT32) throw FormatError("MAT file: bad dimensions element");
  for (uint32_t i = 0; i < dims.size / 4; ++i) {
    const int d = load_swapped<int32_t>(dims.data + 4 * i, swap);
    if (d < 0) throw FormatError("MAT file: negative dimension");
    a.dims.push_back(d);
  }
  Element name = read_element(r, dims.next);
  a.name.assign((const char*)name.data, name.size);
  size_t off = name.next;
  const size_t cnt = a.numel();
  if (cnt > (size_t)1 << 28) throw FormatError("MAT file: array too large");
  // containers: every element needs at least an 8-byte tag inside this element's payload
  if ((a.cls == mxCELL || a.cls == mxSTRUCT || a.cls == mxOBJECT) && cnt > size / 8) throw FormatError("MAT file: container larger than its data element");
  switch (a.cls) {
    case mxCELL:
      a.cells.resize(cnt);
      for (size_t i = 0; i < cnt; ++i) {
        Element c = read_element(r, off);
        if (c.type != miMATRIX) throw FormatError("MAT file: cell element is not a matrix");
        parse_matrix(c.data, c.size, swap, a.cells[i], depth + 1);
        off = c.next;
      }
      break;
    case mxSTRUCT: case mxOBJECT: {
      if (a.cls == mxOBJECT) off = read_element(r, off).next;          // class name
      Element fl = read_element(r, off);
      if (fl.size < 4) throw FormatError("MAT file: bad struct field-name length element");
      const int flen = load_swapped<int32_t>(fl.data, swap);
      Element fn = read_element(r, fl.next);
      if (flen <= 0 || fn.size % (uint32_t)flen) throw FormatError("MAT file: bad struct field names");
      const size_t nf = fn.size / (uint32_t)flen;
      for (size_t f = 0; f < nf; ++f) {
        const char* s = (const char*)fn.data + f * flen;
        a.fields.emplace_back(s, strnlen(s, flen));
      }
      off = fn.next;
      if (nf > size / 8 || cnt * nf > size / 8) throw FormatError("MAT file: struct array larger than its data element");
      a.fvals.resize(cnt * nf);
      for (size_t i = 0; i < cnt * nf; ++i) {
        Element c = read_element(r, off);
        if (c.type != miMATRIX) throw FormatError("MAT file: struct field is not a matrix");
        parse_matrix(c.data, c.size, swap, a.fvals[i], depth + 1);
        off = c.next;
      }
      break;
    }
    case mxCHAR: {
      if (cnt == 0) break;
      Element d = read_element(r, off);
      std::vector<double> codes;
      numeric_to_double(d, swap, codes);
      for (double c : codes) a.text.push_back(c > 0 && c < 128 ? (char)c : '?');
      break;
    }
    case mxSPARSE:
      throw FormatError("MAT file: sparse arrays are not supported");
    default: {
      if (a.cls < mxDOUBLE || a.cls > mxUINT64) throw FormatError("MAT file: unknown array class " + std::to_string(a.cls));
      if (cnt == 0) break;
      Element d = read_element(r, off);
      numeric_to_double(d, swap, a.real);
      if (a.real.size() < cnt) throw FormatError("MAT file: numeric array shorter than its dimensions");
      a.real.resize(cnt);                            // an imaginary part, if any, is ignored
    }
  }
}

constexpr size_t kMaxInflated = (size_t)1 << 30;      // a model file holds a few MB; bound what a hostile stream can allocate
std::vector<uint8_t> inflate_all(const uint8_t* src, size_t n) {
  z_stream zs;
  std::memset(&zs, 0, sizeof(zs));
  if (inflateInit(&zs) != Z_OK) throw FormatError("zlib: inflateInit failed");
  std::vector<uint8_t> out(std::max<size_t>(4096, n * 4));
  zs.next_in = const_cast<Bytef*>(src); zs.avail_in = (uInt)n;
  for (;;) {
    if (zs.total_out == out.size()) {
      if (out.size() >= kMaxInflated) { inflateEnd(&zs); throw FormatError("MAT file: compressed variable expands beyond 1 GiB"); }
      out.resize(std::min(out.size() * 2, kMaxInflated));
    }
    zs.next_out = out.data() + zs.total_out;
    zs.avail_out = (uInt)std::min<size_t>(out.size() - zs.total_out, (size_t)1 << 30);
    const int rc = inflate(&zs, Z_NO_FLUSH);
    if (rc == Z_STREAM_END) break;
    if (rc != Z_OK) { inflateEnd(&zs); throw FormatError("MAT file: corrupt compressed variable"); }
    if (zs.avail_in == 0 && zs.avail_out != 0) { inflateEnd(&zs); throw FormatError("MAT file: truncated compressed variable"); }
  }
  const size_t have = (size_t)zs.total_out;
  inflateEnd(&zs);
  out.resize(have);
  return out;
}

std::vector<MatArray> read_mat(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw IoError("cannot open '" + path + "'");
  std::vector<uint8_t> buf;
  uint8_t chunk[1 << 16];
  size_t got;
  while ((got = std::fread(chunk, 1, sizeof(chunk), f)) > 0) buf.insert(buf.end(), chunk, chunk + got);
  std::fclose(f);
  if (buf.size() < 128) throw FormatError("'" + path + "' is not a MATLAB 5.0 MAT-file (too short)");
  if (std::memcmp(buf.data(), "MATLAB 5.0 MAT-file", 19) != 0) throw FormatError("'" + path + "' is not a MATLAB 5.0 MAT-file");
  bool swap;
  if (buf[126] == 'I' && buf[127] == 'M') swap = false;          // written little endian (read here on a little-endian host)
  else if (buf[126] == 'M' && buf[127] == 'I') swap = true;
  else throw FormatError("MAT file: bad endian indicator");
  Reader r{buf.data(), buf.size(), swap};
  std::vector<MatArray> vars;
  size_t off = 128;
  while (off + 8 <= buf.size()) {
    Element e = read_element(r, off);
    if (e.type == miCOMPRESSED) {
      std::vector<uint8_t> raw = inflate_all(e.data, e.size);
      Reader rr{raw.data(), raw.size(), swap};
      Element m = read_element(rr, 0);
      if (m.type == miMATRIX) { vars.emplace_back(); parse_matrix(m.data, m.size, swap, vars.back(), 0); }
    } else if (e.type == miMATRIX) {
      vars.emplace_back(); parse_matrix(e.data, e.size, swap, vars.back(), 0);
    }
    off = e.next;
  }
  return vars;
}

const MatArray& need_field(const MatArray& s, size_t elem, const char* name, const char* where) {
  const MatArray* a = s.field(elem, name);
  if (!a) throw FormatError(std::string("MAT model: field '") + name + "' missing in " + where);
  return *a;
}
double scalar_of(const MatArray& a, const char* what) {
  if (a.real.empty()) throw FormatError(std::string("MAT model: '") + what + "' is not numeric");
  return a.real[0];
}
// the reference iterates the cv::Mat that cvmatio made of the array: row-major over MATLAB's (rows, cols)
std::vector<int> ids_row_major_zero_based(const MatArray& a, const char* what) {
  if (a.real.size() != a.numel()) throw FormatError(std::string("MAT model: '") + what + "' is not numeric");
  const int rows = a.dims.size() > 0 ? a.dims[0] : 0, cols = rows ? (int)(a.numel() / rows) : 0;
  std::vector<int> out;
  out.reserve(a.numel());
  for (int rr = 0; rr < rows; ++rr)
    for (int cc = 0; cc < cols; ++cc) out.push_back((int)a.real[(size_t)cc * rows + rr] - 1);
  return out;
}

}  // namespace

void load_mat(const std::string& path, Model& m) {
  const std::vector<MatArray> vars = read_mat(path);
  const MatArray* model = nullptr;
  const MatArray* name = nullptr;
  for (const MatArray& v : vars) {
    if (v.name == "model" && v.cls == mxSTRUCT) model = &v;
    if (v.name == "name" && v.cls == mxCHAR) name = &v;
  }
  if (!model || model->numel() < 1) throw FormatError("MAT model: no struct variable named 'model'");
  m = Model();
  if (name) m.name = name->text;                      // :84-88: the variable `name`, else the file's stem
  else {
    size_t s = path.find_last_of("/\\");
    std::string stem = s == std::string::npos ? path : path.substr(s + 1);
    const size_t dot = stem.find_last_of('.');
    m.name = dot == std::string::npos ? stem : stem.substr(0, dot);
  }
  m.interval = (int)scalar_of(need_field(*model, 0, "interval", "model"), "interval");   // :99-102
  m.thresh = (float)scalar_of(need_field(*model, 0, "thresh", "model"), "thresh");
  m.sbin = (int)scalar_of(need_field(*model, 0, "sbin", "model"), "sbin");
  m.norient = 18;
  // filters (:106-124): w is M x N x C, flattened to M x (N*C) with the channel fastest
  const MatArray& filters = need_field(*model, 0, "filters", "model");
  if (filters.cls != mxSTRUCT) throw FormatError("MAT model: 'filters' is not a struct array");
  for (size_t f = 0; f < filters.numel(); ++f) {
    const MatArray& w = need_field(filters, f, "w", "filters");
    if (w.dims.size() < 2 || w.real.size() != w.numel()) throw FormatError("MAT model: filter weights are not a numeric array");
    const int M = w.dims[0], N = w.dims[1], Cn = w.dims.size() > 2 ? w.dims[2] : 1;
    std::vector<double> flat((size_t)M * N * Cn);
    for (int mm = 0; mm < M; ++mm)
      for (int c = 0; c < Cn; ++c)
        for (int n = 0; n < N; ++n) flat[((size_t)mm * N + n) * Cn + c] = w.real[(size_t)mm + (size_t)M * (n + (size_t)N * c)];
    m.flen = Cn;
    m.frows.push_back(M); m.fkw.push_back(N);
    m.filters.push_back(std::move(flat));
  }
  // components (:128-163): a cell array of struct arrays (one struct per part)
  const MatArray& comps = need_field(*model, 0, "components", "model");
  if (comps.cls != mxCELL) throw FormatError("MAT model: 'components' is not a cell array");
  for (const MatArray& comp : comps.cells) {
    if (comp.cls != mxSTRUCT) throw FormatError("MAT model: a component is not a struct array");
    std::vector<Part> parts(comp.numel());
    for (size_t p = 0; p < parts.size(); ++p) {
      parts[p].defid = ids_row_major_zero_based(need_field(comp, p, "defid", "component"), "defid");
      if (parts[p].defid.empty()) parts[p].defid.push_back(0);     // the root has no deformation: same convention as load_xml
      parts[p].filterid = ids_row_major_zero_based(need_field(comp, p, "filterid", "component"), "filterid");
      parts[p].biasid = ids_row_major_zero_based(need_field(comp, p, "biasid", "component"), "biasid");
      parts[p].parentid = (int)scalar_of(need_field(comp, p, "parent", "component"), "parent") - 1;
    }
    m.comps.push_back(std::move(parts));
  }
  // defs (:167-175): w (4 weights) and anchor (x, y[, level]) 1-based
  const MatArray& defs = need_field(*model, 0, "defs", "model");
  if (defs.cls != mxSTRUCT) throw FormatError("MAT model: 'defs' is not a struct array");
  for (size_t d = 0; d < defs.numel(); ++d) {
    const MatArray& w = need_field(defs, d, "w", "defs");
    const MatArray& an = need_field(defs, d, "anchor", "defs");
    if (w.real.size() < 4 || an.real.size() < 2) throw FormatError("MAT model: a deformation needs 4 weights and a 2-element anchor");
    for (int i = 0; i < 4; ++i) m.defs.push_back((float)w.real[i]);
    m.anchors.push_back((int)an.real[0] - 1); m.anchors.push_back((int)an.real[1] - 1);
  }
  // bias (:179-185)
  const MatArray& bias = need_field(*model, 0, "bias", "model");
  if (bias.cls != mxSTRUCT) throw FormatError("MAT model: 'bias' is not a struct array");
  for (size_t b = 0; b < bias.numel(); ++b) m.biasw.push_back((float)scalar_of(need_field(bias, b, "w", "bias"), "bias.w"));
  m.validate();
}

}  // namespace pbd
