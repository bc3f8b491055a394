// ingest.cpp -- the step in front of detect(): what the reference's callers do before they hand a cv::Mat to the detector.
//
//   demo:     Mat im = imread(argv[2]);  depth = imread(argv[3], IMREAD_ANYDEPTH) / 1000.0f      (reference src/demo.cpp:88-99)
//   ROS node: cv_bridge::toCvCopy(msg_rgb, enc::BGR8), toCvCopy(msg_d, enc::TYPE_32FC1)            (reference ros/Node.cpp:165-176)
//
// Decoders for the container formats that need no third-party codec (PNG through zlib, which the MAT reader already links, and
// binary PNM), the sensor_msgs/Image encodings cv_bridge converts to BGR8 / 32FC1, and pinned host allocation for the frame ring
// of pbd_submit_batch_u8.  Pure host code; JPEG is not decoded here (no codec in the image: decode upstream, e.g. nvJPEG).
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "engine.hpp"
#include "ingest.hpp"
#include "model.hpp"

namespace pbd {
namespace {

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

struct Raster {                    // decoded samples, interleaved, 8 or 16 bits (16: host byte order)
  int h = 0, w = 0, c = 0, bits = 8;
  std::vector<uint8_t> px;
};

// ---- PNG (ISO/IEC 15948): non-interlaced, colour types 0 / 2 / 3 / 4 / 6, bit depths 1..16 ----
void decode_png(const uint8_t* d, size_t n, Raster& R, bool header_only) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (n < 8 + 25 || memcmp(d, sig, 8)) throw FormatError("PNG: bad signature");
  size_t o = 8;
  int w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte;
  bool have_hdr = false, done = false;
  while (!done && o + 12 <= n) {
    const uint32_t len = be32(d + o);
    if (len > n - o - 12) throw FormatError("PNG: truncated chunk");
    const uint8_t* type = d + o + 4;
    const uint8_t* body = d + o + 8;
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) throw FormatError("PNG: bad IHDR");
      w = (int)be32(body); h = (int)be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
      if (w <= 0 || h <= 0 || w > 65535 || h > 65535) throw FormatError("PNG: image size out of range");
      if (body[10] != 0 || body[11] != 0) throw FormatError("PNG: unknown compression / filter method");
      have_hdr = true;
      if (header_only) break;
    } else if (!memcmp(type, "PLTE", 4)) {
      plte.assign(body, body + len);
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!memcmp(type, "IEND", 4)) {
      done = true;
    }
    o += 12 + (size_t)len;
  }
  if (!have_hdr) throw FormatError("PNG: no IHDR");
  int ch = 0;
  switch (ctype) {
    case 0: ch = 1; break;
    case 2: ch = 3; break;
    case 3: ch = 1; break;
    case 4: ch = 2; break;
    case 6: ch = 4; break;
    default: throw FormatError("PNG: bad colour type");
  }
  const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8) : (depth == 8 || depth == 16);
  if (!depth_ok) throw FormatError("PNG: bad bit depth");
  R.h = h; R.w = w; R.bits = depth == 16 ? 16 : 8;
  R.c = ctype == 3 ? 3 : ch;
  if (header_only) return;
  if (interlace) throw UnsupportedError("PNG: interlaced images are not supported");
  if (ctype == 3 && plte.size() < 3) throw FormatError("PNG: palette image without PLTE");
  const size_t bpp_bits = (size_t)ch * depth, stride = ((size_t)w * bpp_bits + 7) / 8, bpp = std::max<size_t>(1, bpp_bits / 8);
  std::vector<uint8_t> raw((stride + 1) * (size_t)h);
  {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit(&zs) != Z_OK) throw FormatError("zlib: inflateInit failed");
    zs.next_in = idat.data(); zs.avail_in = (uInt)idat.size();
    zs.next_out = raw.data(); zs.avail_out = (uInt)raw.size();
    const int rc = inflate(&zs, Z_FINISH);
    const size_t got = zs.total_out;
    inflateEnd(&zs);
    if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || got != raw.size()) throw FormatError("PNG: corrupt or truncated image data");
  }
  // undo the per-row filters in place
  std::vector<uint8_t> zero(stride, 0);
  for (int y = 0; y < h; ++y) {
    uint8_t* row = raw.data() + (size_t)y * (stride + 1) + 1;
    const uint8_t* up = y ? row - (stride + 1) : zero.data();
    const int ft = row[-1];
    for (size_t i = 0; i < stride; ++i) {
      const int a = i >= bpp ? row[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
      int pred = 0;
      switch (ft) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
        default: throw FormatError("PNG: bad filter type");
      }
      row[i] = (uint8_t)(row[i] + pred);
    }
  }
  // expand to interleaved 8 / 16-bit samples
  R.px.assign((size_t)h * w * R.c * (R.bits / 8), 0);
  for (int y = 0; y < h; ++y) {
    const uint8_t* row = raw.data() + (size_t)y * (stride + 1) + 1;
    uint8_t* out = R.px.data() + (size_t)y * w * R.c * (R.bits / 8);
    if (depth == 16) {
      uint16_t* o16 = reinterpret_cast<uint16_t*>(out);
      for (size_t i = 0; i < (size_t)w * ch; ++i) o16[i] = (uint16_t)((row[2 * i] << 8) | row[2 * i + 1]);
    } else if (depth == 8 && ctype != 3) {
      memcpy(out, row, (size_t)w * ch);
    } else {
      const int maxv = (1 << depth) - 1;
      for (int x = 0; x < w; ++x) {
        const size_t bit = (size_t)x * depth;
        const int v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & maxv;
        if (ctype == 3) {
          if ((size_t)v * 3 + 2 >= plte.size()) throw FormatError("PNG: palette index out of range");
          out[3 * x] = plte[3 * v]; out[3 * x + 1] = plte[3 * v + 1]; out[3 * x + 2] = plte[3 * v + 2];
        } else {
          out[x] = (uint8_t)(v * 255 / maxv);          // grey < 8 bit: scaled to 0..255 as OpenCV's decoder does
        }
      }
    }
  }
}

// ---- binary PNM: P5 (grey) / P6 (RGB), maxval <= 65535 ----
void decode_pnm(const uint8_t* d, size_t n, Raster& R, bool header_only) {
  size_t o = 2;
  auto number = [&]() -> long {
    for (;;) {
      while (o < n && (d[o] == ' ' || d[o] == '\n' || d[o] == '\r' || d[o] == '\t')) ++o;
      if (o < n && d[o] == '#') { while (o < n && d[o] != '\n') ++o; continue; }
      break;
    }
    long v = 0;
    bool any = false;
    while (o < n && d[o] >= '0' && d[o] <= '9') { v = v * 10 + (d[o++] - '0'); any = true; if (v > 1 << 20) break; }
    if (!any) throw FormatError("PNM: bad header");
    return v;
  };
  const long w = number(), h = number(), maxv = number();
  if (w <= 0 || h <= 0 || w > 65535 || h > 65535 || maxv <= 0 || maxv > 65535) throw FormatError("PNM: header values out of range");
  ++o;                                                  // the single whitespace byte after maxval
  R.w = (int)w; R.h = (int)h; R.c = d[1] == '6' ? 3 : 1; R.bits = maxv > 255 ? 16 : 8;
  if (header_only) return;
  const size_t bytes = (size_t)w * h * R.c * (R.bits / 8);
  if (o + bytes > n) throw FormatError("PNM: truncated pixel data");
  R.px.assign(d + o, d + o + bytes);
  if (R.bits == 16) { uint16_t* p = reinterpret_cast<uint16_t*>(R.px.data()); for (size_t i = 0; i < bytes / 2; ++i) p[i] = (uint16_t)((R.px[2 * i] << 8) | R.px[2 * i + 1]); }
}

void decode_any(const uint8_t* d, size_t n, Raster& R, bool header_only) {
  if (n >= 8 && d[0] == 0x89 && d[1] == 'P') return decode_png(d, n, R, header_only);
  if (n >= 7 && d[0] == 'P' && (d[1] == '5' || d[1] == '6')) { decode_pnm(d, n, R, header_only); if (d[1] == '6' && !header_only) R.c = -3; return; }   // -3: RGB order
  if (n >= 3 && d[0] == 0xff && d[1] == 0xd8) throw UnsupportedError("JPEG is not decoded by this library (no codec in the image): decode upstream");
  throw FormatError("unknown image container (PNG and binary PNM are supported)");
}

}  // namespace

void image_info(const uint8_t* bytes, size_t n, int* h, int* w, int* channels, int* bits) {
  Raster R;
  decode_any(bytes, n, R, true);
  if (h) *h = R.h;
  if (w) *w = R.w;
  if (channels) *channels = R.c;
  if (bits) *bits = R.bits;
}

// cv::imread(path, IMREAD_COLOR) semantics: 8-bit BGR, grey replicated, alpha dropped, 16-bit scaled down by 1/256 (>> 8)
void image_decode_bgr8(const uint8_t* bytes, size_t n, uint8_t* dst, size_t cap, int* h, int* w) {
  Raster R;
  decode_any(bytes, n, R, false);
  const bool rgb_pnm = R.c == -3;                     // PNM stores R, G, B like PNG
  const int c = R.c < 0 ? -R.c : R.c;
  if ((size_t)R.h * R.w * 3 > cap) throw ArgError("image_decode: destination too small");
  (void)rgb_pnm;
  for (size_t i = 0; i < (size_t)R.h * R.w; ++i) {
    int s[4] = {0, 0, 0, 0};
    for (int k = 0; k < c; ++k) s[k] = R.bits == 16 ? reinterpret_cast<const uint16_t*>(R.px.data())[i * c + k] >> 8 : R.px[i * c + k];
    uint8_t* o = dst + 3 * i;
    if (c <= 2) { o[0] = o[1] = o[2] = (uint8_t)s[0]; }            // grey (+ alpha)
    else { o[0] = (uint8_t)s[2]; o[1] = (uint8_t)s[1]; o[2] = (uint8_t)s[0]; }   // file order R, G, B -> B, G, R
  }
  if (h) *h = R.h;
  if (w) *w = R.w;
}

// cv::imread(path, IMREAD_ANYDEPTH) of a single-channel depth image, times `scale` (src/demo.cpp:96-99: mm -> m)
void image_decode_depth_f32(const uint8_t* bytes, size_t n, float scale, float* dst, size_t cap, int* h, int* w) {
  Raster R;
  decode_any(bytes, n, R, false);
  const int c = R.c < 0 ? -R.c : R.c;
  if ((size_t)R.h * R.w > cap) throw ArgError("image_decode: destination too small");
  for (size_t i = 0; i < (size_t)R.h * R.w; ++i) {
    const float v = R.bits == 16 ? (float)reinterpret_cast<const uint16_t*>(R.px.data())[i * c] : (float)R.px[i * c];
    dst[i] = v * scale;
  }
  if (h) *h = R.h;
  if (w) *w = R.w;
}

std::vector<uint8_t> slurp_bytes(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) throw IoError("cannot open '" + path + "'");
  std::vector<uint8_t> buf;
  uint8_t chunk[1 << 16];
  size_t got;
  while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) {
    buf.insert(buf.end(), chunk, chunk + got);
    if (buf.size() > ((size_t)1 << 30)) { fclose(f); throw FormatError("image file larger than 1 GiB"); }
  }
  fclose(f);
  return buf;
}

// sensor_msgs/Image payload -> packed BGR8 (cv_bridge::toCvCopy(msg, "bgr8"), ros/Node.cpp:169-170)
void ros_image_to_bgr8(const char* encoding, int h, int w, size_t step, int is_bigendian, const uint8_t* src, uint8_t* dst) {
  const std::string e(encoding ? encoding : "");
  int c = 0, bits = 8;
  enum { BGR, RGB, MONO } order = BGR;
  if (e == "bgr8" || e == "8UC3") { c = 3; order = BGR; }
  else if (e == "rgb8") { c = 3; order = RGB; }
  else if (e == "bgra8" || e == "8UC4") { c = 4; order = BGR; }
  else if (e == "rgba8") { c = 4; order = RGB; }
  else if (e == "mono8" || e == "8UC1") { c = 1; order = MONO; }
  else if (e == "mono16" || e == "16UC1") { c = 1; order = MONO; bits = 16; }
  else if (e == "bgr16") { c = 3; order = BGR; bits = 16; }
  else if (e == "rgb16") { c = 3; order = RGB; bits = 16; }
  else throw UnsupportedError("sensor_msgs/Image encoding '" + e + "' cannot be converted to bgr8");
  const size_t bps = bits / 8;
  if (step == 0) step = (size_t)w * c * bps;
  if (step < (size_t)w * c * bps) throw ArgError("sensor_msgs/Image: step smaller than a row");
  for (int y = 0; y < h; ++y) {
    const uint8_t* row = src + (size_t)y * step;
    uint8_t* o = dst + (size_t)y * w * 3;
    for (int x = 0; x < w; ++x) {
      int s[4] = {0, 0, 0, 0};
      for (int k = 0; k < c; ++k) {
        if (bits == 8) s[k] = row[(size_t)x * c + k];
        else {                                           // 16 -> 8 bit: cv_bridge scales by 1/256 (value * 255 / 65535, rounded like convertTo)
          const uint8_t* p = row + ((size_t)x * c + k) * 2;
          const int v = is_bigendian ? (p[0] << 8 | p[1]) : (p[1] << 8 | p[0]);
          s[k] = (int)lrint((double)v * (255.0 / 65535.0));
        }
      }
      if (order == MONO) { o[3 * x] = o[3 * x + 1] = o[3 * x + 2] = (uint8_t)s[0]; }
      else if (order == BGR) { o[3 * x] = (uint8_t)s[0]; o[3 * x + 1] = (uint8_t)s[1]; o[3 * x + 2] = (uint8_t)s[2]; }
      else { o[3 * x] = (uint8_t)s[2]; o[3 * x + 1] = (uint8_t)s[1]; o[3 * x + 2] = (uint8_t)s[0]; }
    }
  }
}

// sensor_msgs/Image depth payload -> packed 32FC1 (cv_bridge::toCvCopy(msg, TYPE_32FC1): plain value conversion, no unit scaling)
void ros_depth_to_f32(const char* encoding, int h, int w, size_t step, int is_bigendian, const uint8_t* src, float* dst) {
  const std::string e(encoding ? encoding : "");
  const bool f32 = e == "32FC1", u16 = e == "16UC1" || e == "mono16";
  if (!f32 && !u16) throw UnsupportedError("sensor_msgs/Image depth encoding '" + e + "' is not 32FC1 / 16UC1");
  const size_t bps = f32 ? 4 : 2;
  if (step == 0) step = (size_t)w * bps;
  if (step < (size_t)w * bps) throw ArgError("sensor_msgs/Image: step smaller than a row");
  for (int y = 0; y < h; ++y) {
    const uint8_t* row = src + (size_t)y * step;
    for (int x = 0; x < w; ++x) {
      const uint8_t* p = row + (size_t)x * bps;
      if (f32) {
        uint8_t b[4] = {p[0], p[1], p[2], p[3]};
        if (is_bigendian) { b[0] = p[3]; b[1] = p[2]; b[2] = p[1]; b[3] = p[0]; }
        float v;
        memcpy(&v, b, 4);
        dst[(size_t)y * w + x] = v;
      } else {
        dst[(size_t)y * w + x] = (float)(is_bigendian ? (p[0] << 8 | p[1]) : (p[1] << 8 | p[0]));
      }
    }
  }
}

void* pinned_alloc(size_t bytes) {
  void* p = nullptr;
  const cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) throw CudaError(std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return p;
}
void pinned_free(void* p) { if (p) cudaFreeHost(p); }

}  // namespace pbd
