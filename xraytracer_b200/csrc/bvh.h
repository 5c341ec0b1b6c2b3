// bvh.h — host-side SAH BVH build for libxrtgpu.so. This is NEW functionality relative to the reference:
// Scene::intersect/occluded are brute-force loops (scene.cpp:190-211, primitive.cpp:83-138) and
// Scene::build() is an empty hook (scene.h:22-24). The BVH must therefore return exactly what brute force
// returns: boxes are padded conservatively, traversal never culls a node whose entry distance
// equals the current best, and ties in t resolve to the lowest primitive id.
#pragma once
#include <cstdint>
#include <vector>

namespace xrt {

// Two-child node with BOTH children's boxes stored in the parent (one 64-byte fetch tests both):
//   lo0.xyz hi0.xyz lo1.xyz hi1.xyz | child0 child1 count0 count1
// count > 0  : leaf, `child` is the first entry in the leaf-ordered triangle array
// count == 0 : inner node, `child` is a node index
// count <  0 : empty slot (never hit)
struct BvhNode {
    float lo0[3], hi0[3], lo1[3], hi1[3];
    int32_t child0, child1, count0, count1;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be one 64-byte record");

struct Bvh {
    std::vector<BvhNode> nodes;     // nodes[0] is the root
    std::vector<uint32_t> triOrder; // leaf-ordered -> index into the input triangle array
    int depth = 0;
    float sahCost = 0.f;            // SAH cost estimate (Ct = Ci = 1) relative to the root area
    float pad = 0.f;                // absolute padding added to every box
};

// Four-child node for the resumable traversal kernel of deep trees (k_trace): 128 bytes = one cache line, children's boxes in SoA
// so that two children share a packed fp32x2 instruction. Same child / count convention as BvhNode. One fetch decides among four
// children: half as many DEPENDENT node fetches per ray as the two-child form, and the wait for the node fetch is the hottest
// instruction of that kernel (profiles/r01_notes.md).
struct Bvh4Node {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    int32_t child[4], count[4];
};
static_assert(sizeof(Bvh4Node) == 128, "Bvh4Node must be one 128-byte record");

// Collapses a two-child tree (as produced by buildBvh or by the GPU LBVH builder) into four-child nodes: every wide node starts
// from the two children of a BvhNode and repeatedly replaces its largest-area inner child by that child's two children. The
// leaves and their (already padded) boxes are taken over unchanged. Returns the depth of the wide tree.
int collapseBvh4(const BvhNode* nodes, size_t nNodes, std::vector<Bvh4Node>& out);

// tri = n * 9 floats (v0 v1 v2). Binned SAH (16 bins, 3 axes), leaves of at most `maxLeaf` triangles.
void buildBvh(const float* tri, uint32_t n, int maxLeaf, Bvh& out);

} // namespace xrt
