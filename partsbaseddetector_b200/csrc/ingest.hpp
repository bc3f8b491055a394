// ingest.hpp -- host-side ingest helpers in front of detect(): image containers, sensor_msgs/Image encodings, pinned buffers.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace pbd {
void image_info(const uint8_t* bytes, size_t n, int* h, int* w, int* channels, int* bits);
void image_decode_bgr8(const uint8_t* bytes, size_t n, uint8_t* dst, size_t cap, int* h, int* w);
void image_decode_depth_f32(const uint8_t* bytes, size_t n, float scale, float* dst, size_t cap, int* h, int* w);
std::vector<uint8_t> slurp_bytes(const std::string& path);
void ros_image_to_bgr8(const char* encoding, int h, int w, size_t step, int is_bigendian, const uint8_t* src, uint8_t* dst);
void ros_depth_to_f32(const char* encoding, int h, int w, size_t step, int is_bigendian, const uint8_t* src, float* dst);
void* pinned_alloc(size_t bytes);
void pinned_free(void* p);
}  // namespace pbd
