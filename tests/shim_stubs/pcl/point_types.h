#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>
#include "../eigen_stub.h"
namespace boost { using std::shared_ptr; }              // PCL 1.8 hands out boost::shared_ptr
namespace pcl {
struct alignas(16) PointXYZ { float x, y, z, pad; };
struct alignas(16) PointXYZI { float x, y, z, pad; float intensity; float pad2[3]; };
struct alignas(16) PointXYZRGBL { float x, y, z, pad; std::uint32_t rgba; std::uint32_t label; float pad2[2]; };
static_assert(sizeof(PointXYZ) == 16 && sizeof(PointXYZI) == 32 && sizeof(PointXYZRGBL) == 32, "PCL point layouts");
}  // namespace pcl
