// Stub of the OpenVDB subset grid.h:22-85 names, enough for OpenVDBGrid to *compile*; it is never
// instantiated (no .vdb fixture ships with the reference and the library is absent). The oracle's
// DenseGrid (ref_harness.cpp) restates the lookup semantics instead. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <memory>
#include <string>
namespace openvdb {
struct Vec3f { float v[3]; Vec3f(float x = 0, float y = 0, float z = 0) : v{x, y, z} {} };
struct Vec3d { double v[3]; double operator[](int i) const { return v[i]; } };
struct Coord { int x = 0, y = 0, z = 0; };
struct CoordBBox { Coord getStart() const { return {}; } Coord getEnd() const { return {}; } };
struct GridBase { using Ptr = std::shared_ptr<GridBase>; virtual ~GridBase() = default; };
struct FloatGrid : GridBase {
    using Ptr = std::shared_ptr<FloatGrid>;
    CoordBBox evalActiveVoxelBoundingBox() const { return {}; }
    Vec3d indexToWorld(const Coord&) const { return {}; }
    void evalMinMax(float& a, float& b) const { a = b = 0; }
};
inline void initialize() {}
template <typename G> std::shared_ptr<G> gridPtrCast(const GridBase::Ptr& p) { return std::dynamic_pointer_cast<G>(p); }
namespace io {
struct File {
    explicit File(const std::string&) {}
    bool open() { return false; }
    GridBase::Ptr readGrid(const std::string&) { return nullptr; }
    void close() {}
};
} // namespace io
} // namespace openvdb
