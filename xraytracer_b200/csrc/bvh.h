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

// Eight-child node with 8-bit QUANTISED child boxes for the throughput instantiation's traversal of deep trees (k_trace8),
// 80 bytes = 5 x 16 B, after Ylitie, Karras & Laine, "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs"
// (HPG 2017), simplified: child box k, axis a = [p[a] + qlo[a][k] * 2^e[a], p[a] + qhi[a][k] * 2^e[a]] (lo rounded down, hi rounded
// up: a superset of the padded box of the two-child tree, so the traversal stays conservative and the triangle tests decide).
// Children sit in SLOTS chosen so that visiting slots in the order of decreasing (slot ^ octant-inverse) approximates front to
// back for a ray of that direction octant — no per-ray sorting. Inner children are stored consecutively from childBase in slot
// order (child = childBase + popcount(imask below the slot)); the triangles of all leaf children are stored consecutively from
// triBase in the node-ordered triangle array, in slot order. validTri has one NIBBLE per slot: (1 << count) - 1 for a leaf child
// (at most 4 triangles per leaf), 0 otherwise — the kernel ANDs it with "0xF per hit slot" to get the node's triangle hits in
// one instruction, and bit b of that mask is triangle triBase + popcount(validTri below b).
// 1 M triangles: ~75 k nodes = 6 MB instead of 25 MB of four-child nodes.
struct Bvh8Node {
    float p[3];
    uint8_t e[3];      // biased exponents: 2^e as a float is (e << 23)
    uint8_t imask;     // bit k: child k is an inner node
    uint32_t childBase;
    uint32_t triBase;
    uint32_t validTri; // nibble k: triangle bits of leaf child k
    uint32_t reserved;
    uint8_t qlo[3][8], qhi[3][8];
};
static_assert(sizeof(Bvh8Node) == 80, "Bvh8Node must be 80 bytes");

// Collapses a two-child tree into eight-child quantised nodes (largest-area inner child opened first, greedy slot assignment).
// triOrder8[k] = index into the LEAF-ordered triangle array of the triangle that the wide tree stores at position k.
// Returns the depth of the wide tree (0 if nNodes == 0).
int collapseBvh8(const BvhNode* nodes, size_t nNodes, std::vector<Bvh8Node>& out, std::vector<uint32_t>& triOrder8);

// tri = n * 9 floats (v0 v1 v2). Binned SAH (16 bins, 3 axes), leaves of at most `maxLeaf` triangles.
void buildBvh(const float* tri, uint32_t n, int maxLeaf, Bvh& out);

} // namespace xrt
