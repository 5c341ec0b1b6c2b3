// Stub, see ../openvdb.h. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <openvdb/openvdb.h>
namespace openvdb { namespace tools {
struct BoxSampler {};
template <typename G, typename S> struct GridSampler {
    explicit GridSampler(const G&) {}
    float wsSample(const openvdb::Vec3f&) const { return 0.0f; }
};
}} // namespace openvdb::tools
