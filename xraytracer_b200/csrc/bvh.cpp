// bvh.cpp — binned-SAH BVH2 builder (host, OpenMP tasks), see bvh.h.
#include "bvh.h"
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace xrt {
namespace {

struct Box {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const float* p) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void grow(const Box& b) { for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], b.lo[a]); hi[a] = std::max(hi[a], b.hi[a]); } }
    float area() const
    {
        const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0) return 0.f;
        return 2.f * (dx * dy + dy * dz + dz * dx);
    }
};

struct BuildNode {
    Box box;
    int32_t left = -1, right = -1; // children (build-node indices) or -1
    uint32_t first = 0, count = 0; // leaf range in `order`
    int depth = 0;
};

struct Builder {
    const float* tri;
    std::vector<Box> tbox;
    std::vector<float> cent; // 3 per triangle
    std::vector<uint32_t> order;
    std::vector<BuildNode> pool;
    std::atomic<int32_t> next{0};
    int maxLeaf = 4;
    static constexpr int kBins = 16;
    static constexpr int kMaxDepth = 56;

    int32_t alloc() { return next.fetch_add(1); }

    void build(int32_t ni, uint32_t first, uint32_t count, int depth)
    {
        BuildNode& node = pool[ni];
        node.depth = depth;
        Box b, cb;
        for (uint32_t i = first; i < first + count; ++i) {
            b.grow(tbox[order[i]]);
            cb.grow(&cent[3 * order[i]]);
        }
        node.box = b;
        auto makeLeaf = [&]() { node.first = first; node.count = count; };
        if (count <= 1) { makeLeaf(); return; }

        // binned SAH over the three axes
        float bestCost = FLT_MAX;
        int bestAxis = -1, bestSplit = -1;
        if (depth < kMaxDepth) {
            for (int a = 0; a < 3; ++a) {
                const float ext = cb.hi[a] - cb.lo[a];
                if (!(ext > 0.f)) continue;
                Box bins[kBins];
                uint32_t cnt[kBins] = {};
                const float k = kBins * (1.f - 1e-6f) / ext;
                for (uint32_t i = first; i < first + count; ++i) {
                    const uint32_t t = order[i];
                    int bi = int((cent[3 * t + a] - cb.lo[a]) * k);
                    bi = std::min(std::max(bi, 0), kBins - 1);
                    bins[bi].grow(tbox[t]);
                    cnt[bi]++;
                }
                float rightArea[kBins];
                uint32_t rightCnt[kBins];
                Box acc;
                uint32_t c = 0;
                for (int i = kBins - 1; i > 0; --i) {
                    acc.grow(bins[i]); c += cnt[i];
                    rightArea[i] = acc.area(); rightCnt[i] = c;
                }
                acc = Box(); c = 0;
                for (int i = 0; i < kBins - 1; ++i) {
                    acc.grow(bins[i]); c += cnt[i];
                    if (c == 0 || rightCnt[i + 1] == 0) continue;
                    const float cost = acc.area() * float(c) + rightArea[i + 1] * float(rightCnt[i + 1]);
                    if (cost < bestCost) { bestCost = cost; bestAxis = a; bestSplit = i; }
                }
            }
        }
        uint32_t mid = 0;
        if (bestAxis >= 0) {
            const float leafCost = b.area() * float(count);
            if (count <= uint32_t(maxLeaf) && leafCost <= bestCost + b.area() /*traversal step*/) { makeLeaf(); return; }
            const int a = bestAxis;
            const float ext = cb.hi[a] - cb.lo[a];
            const float k = kBins * (1.f - 1e-6f) / ext;
            const float lo = cb.lo[a];
            auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t t) {
                int bi = int((cent[3 * t + a] - lo) * k);
                bi = std::min(std::max(bi, 0), kBins - 1);
                return bi <= bestSplit;
            });
            mid = uint32_t(it - order.begin());
        }
        if (bestAxis < 0 || mid == first || mid == first + count) {
            // identical centroids or depth limit: leaf if small enough, else split the index range in half
            if (count <= uint32_t(maxLeaf)) { makeLeaf(); return; }
            mid = first + count / 2;
        }
        const int32_t l = alloc(), r = alloc();
        pool[ni].left = l;
        pool[ni].right = r;
        const uint32_t lc = mid - first, rc = first + count - mid;
        if (count > 8192) {
#pragma omp task default(shared) firstprivate(l, first, lc, depth)
            build(l, first, lc, depth + 1);
#pragma omp task default(shared) firstprivate(r, mid, rc, depth)
            build(r, mid, rc, depth + 1);
#pragma omp taskwait
        }
        else {
            build(l, first, lc, depth + 1);
            build(r, mid, rc, depth + 1);
        }
    }
};

void setChild(BvhNode& n, int slot, const Box& b, float pad, int32_t child, int32_t count)
{
    float* lo = slot ? n.lo1 : n.lo0;
    float* hi = slot ? n.hi1 : n.hi0;
    for (int a = 0; a < 3; ++a) { lo[a] = b.lo[a] - pad; hi[a] = b.hi[a] + pad; }
    (slot ? n.child1 : n.child0) = child;
    (slot ? n.count1 : n.count0) = count;
}

void setEmpty(BvhNode& n, int slot)
{
    float* lo = slot ? n.lo1 : n.lo0;
    float* hi = slot ? n.hi1 : n.hi0;
    for (int a = 0; a < 3; ++a) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }
    (slot ? n.child1 : n.child0) = 0;
    (slot ? n.count1 : n.count0) = -1;
}

} // namespace

int collapseBvh4(const BvhNode* nodes, size_t nNodes, std::vector<Bvh4Node>& out)
{
    struct Slot { float lo[3], hi[3]; int32_t child, count; };
    auto slotsOf = [&](int32_t n, Slot s[2]) {
        const BvhNode& b = nodes[n];
        for (int a = 0; a < 3; ++a) { s[0].lo[a] = b.lo0[a]; s[0].hi[a] = b.hi0[a]; s[1].lo[a] = b.lo1[a]; s[1].hi[a] = b.hi1[a]; }
        s[0].child = b.child0; s[0].count = b.count0; s[1].child = b.child1; s[1].count = b.count1;
    };
    auto area = [](const Slot& s) {
        const float dx = s.hi[0] - s.lo[0], dy = s.hi[1] - s.lo[1], dz = s.hi[2] - s.lo[2];
        return dx * dy + dy * dz + dz * dx;
    };
    out.clear();
    if (nNodes == 0) return 0;
    struct Todo { int32_t bvh2, wide, depth; };
    std::vector<Todo> stack{{0, 0, 1}};
    out.emplace_back();
    int depth = 0;
    while (!stack.empty()) {
        const Todo t = stack.back();
        stack.pop_back();
        depth = std::max(depth, t.depth);
        Slot sl[4];
        int n = 0;
        Slot two[2];
        slotsOf(t.bvh2, two);
        for (int k = 0; k < 2; ++k) if (two[k].count >= 0) sl[n++] = two[k];
        while (n < 4) { // open the largest inner child
            int best = -1;
            float bestArea = -1.f;
            for (int k = 0; k < n; ++k)
                if (sl[k].count == 0 && area(sl[k]) > bestArea) { bestArea = area(sl[k]); best = k; }
            if (best < 0) break;
            slotsOf(sl[best].child, two);
            int added = 0;
            Slot repl[2];
            for (int k = 0; k < 2; ++k) if (two[k].count >= 0) repl[added++] = two[k];
            if (added == 0) { sl[best] = sl[--n]; continue; }
            if (n - 1 + added > 4) break;
            sl[best] = repl[0];
            if (added == 2) sl[n++] = repl[1];
        }
        Bvh4Node w;
        for (int k = 0; k < 4; ++k) {
            if (k < n) {
                w.lox[k] = sl[k].lo[0]; w.loy[k] = sl[k].lo[1]; w.loz[k] = sl[k].lo[2];
                w.hix[k] = sl[k].hi[0]; w.hiy[k] = sl[k].hi[1]; w.hiz[k] = sl[k].hi[2];
                w.count[k] = sl[k].count;
                if (sl[k].count == 0) { // inner: allocate the wide child now, fill it later
                    w.child[k] = int32_t(out.size());
                    out.emplace_back();
                    stack.push_back({sl[k].child, w.child[k], t.depth + 1});
                }
                else w.child[k] = sl[k].child;
            }
            else {
                w.lox[k] = w.loy[k] = w.loz[k] = FLT_MAX;
                w.hix[k] = w.hiy[k] = w.hiz[k] = -FLT_MAX;
                w.child[k] = 0; w.count[k] = -1;
            }
        }
        out[size_t(t.wide)] = w;
    }
    return depth;
}

int collapseBvh8(const BvhNode* nodes, size_t nNodes, std::vector<Bvh8Node>& out, std::vector<uint32_t>& triOrder8)
{
    struct Slot { float lo[3], hi[3]; int32_t child, count; };
    auto slotsOf = [&](int32_t n, Slot s[2]) {
        const BvhNode& b = nodes[n];
        for (int a = 0; a < 3; ++a) { s[0].lo[a] = b.lo0[a]; s[0].hi[a] = b.hi0[a]; s[1].lo[a] = b.lo1[a]; s[1].hi[a] = b.hi1[a]; }
        s[0].child = b.child0; s[0].count = b.count0; s[1].child = b.child1; s[1].count = b.count1;
    };
    auto area = [](const Slot& s) {
        const float dx = s.hi[0] - s.lo[0], dy = s.hi[1] - s.lo[1], dz = s.hi[2] - s.lo[2];
        return dx * dy + dy * dz + dz * dx;
    };
    out.clear();
    triOrder8.clear();
    if (nNodes == 0) return 0;
    struct Todo { int32_t bvh2, wide, depth; };
    std::vector<Todo> stack{{0, 0, 1}};
    out.emplace_back();
    int depth = 0;
    while (!stack.empty()) {
        const Todo t = stack.back();
        stack.pop_back();
        depth = std::max(depth, t.depth);
        // ---- gather up to eight children: open the largest inner child while there is room ----
        Slot ch[8];
        int n = 0;
        Slot two[2];
        slotsOf(t.bvh2, two);
        for (int k = 0; k < 2; ++k) if (two[k].count >= 0) ch[n++] = two[k];
        while (n < 8) {
            int best = -1;
            float bestArea = -1.f;
            for (int k = 0; k < n; ++k)
                if (ch[k].count == 0 && area(ch[k]) > bestArea) { bestArea = area(ch[k]); best = k; }
            if (best < 0) break;
            slotsOf(ch[best].child, two);
            int added = 0;
            Slot repl[2];
            for (int k = 0; k < 2; ++k) if (two[k].count >= 0) repl[added++] = two[k];
            if (added == 0) { ch[best] = ch[--n]; continue; }
            if (n - 1 + added > 8) break;
            ch[best] = repl[0];
            if (added == 2) ch[n++] = repl[1];
        }
        // a node may hold at most 32 triangles in its leaf children (one 32-bit hit mask): leaves are <= 4 triangles, 8 x 4 = 32
        // ---- node frame: origin = lower corner, one power-of-two scale per axis ----
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int k = 0; k < n; ++k)
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], ch[k].lo[a]); hi[a] = std::max(hi[a], ch[k].hi[a]); }
        Bvh8Node w{};
        float scale[3];
        for (int a = 0; a < 3; ++a) {
            // The kernel's slab FMA carries up to half a quantisation step of rounding error (wf_trace8.cuh: byteMagic), so every
            // child box gets ONE step of margin on each side: the frame starts one step below the children's lower corner and
            // spans 253 steps, which leaves q in [1, 254] before the margin and [0, 255] after it.
            const float ext = std::max(hi[a] - lo[a], 1e-30f);
            int ex = int(std::ceil(std::log2(double(ext) / 253.0)));
            while (std::ldexp(253.0, ex) < double(hi[a]) - double(lo[a])) ++ex;
            ex = std::min(std::max(ex, -100), 100);
            w.e[a] = uint8_t(ex + 127);
            scale[a] = std::ldexp(1.0f, ex);
            w.p[a] = lo[a] - scale[a];
            // (lo - scale rounds in float: make sure the frame origin really lies at least one step below the children)
            while (!((double(lo[a]) - double(w.p[a])) / double(scale[a]) >= 1.0)) w.p[a] = std::nextafter(w.p[a], -FLT_MAX);
        }
        // ---- slot assignment: child k prefers the slot whose octant direction its centre is displaced to (greedy) ----
        float cen[3];
        for (int a = 0; a < 3; ++a) cen[a] = 0.5f * (lo[a] + hi[a]);
        float cost[8][8];
        for (int k = 0; k < n; ++k)
            for (int s8 = 0; s8 < 8; ++s8) {
                float c = 0.f;
                for (int a = 0; a < 3; ++a) {
                    const float dirA = ((s8 >> a) & 1) ? 1.f : -1.f; // slot bit a set = the child lies on the + side of axis a
                    c += dirA * (0.5f * (ch[k].lo[a] + ch[k].hi[a]) - cen[a]);
                }
                cost[k][s8] = c;
            }
        int slotOf[8], childAt[8];
        for (int k = 0; k < 8; ++k) { slotOf[k] = -1; childAt[k] = -1; }
        for (int round = 0; round < n; ++round) {
            int bk = -1, bs = -1;
            float bc = -FLT_MAX;
            for (int k = 0; k < n; ++k) {
                if (slotOf[k] >= 0) continue;
                for (int s8 = 0; s8 < 8; ++s8)
                    if (childAt[s8] < 0 && cost[k][s8] > bc) { bc = cost[k][s8]; bk = k; bs = s8; }
            }
            slotOf[bk] = bs;
            childAt[bs] = bk;
        }
        // ---- emit: inner children get consecutive node indices, leaf triangles consecutive triangle slots, both in slot order ----
        w.childBase = uint32_t(out.size());
        w.triBase = uint32_t(triOrder8.size());
        for (int s8 = 0; s8 < 8; ++s8) {
            for (int a = 0; a < 3; ++a) { w.qlo[a][s8] = 255; w.qhi[a][s8] = 0; } // empty: inverted box
            const int k = childAt[s8];
            if (k < 0) continue;
            const Slot& c = ch[k];
            for (int a = 0; a < 3; ++a) {
                const double ql = std::floor((double(c.lo[a]) - double(w.p[a])) / double(scale[a])) - 1.0; // one step of margin
                const double qh = std::ceil((double(c.hi[a]) - double(w.p[a])) / double(scale[a])) + 1.0;
                w.qlo[a][s8] = uint8_t(std::min(std::max(ql, 0.0), 255.0));
                w.qhi[a][s8] = uint8_t(std::min(std::max(qh, 0.0), 255.0));
            }
            if (c.count == 0) {
                w.imask |= uint8_t(1u << s8);
                const int32_t idx = int32_t(out.size());
                out.emplace_back();
                stack.push_back({c.child, idx, t.depth + 1});
            }
            else {
                w.validTri |= ((1u << c.count) - 1u) << (4 * s8);
                for (int i = 0; i < c.count; ++i) triOrder8.push_back(uint32_t(c.child + i));
            }
        }
        out[size_t(t.wide)] = w;
    }
    return depth;
}

void buildBvh(const float* tri, uint32_t n, int maxLeaf, Bvh& out)
{
    out = Bvh();
    if (n == 0) {
        out.nodes.resize(1);
        setEmpty(out.nodes[0], 0);
        setEmpty(out.nodes[0], 1);
        return;
    }
    Builder B;
    B.tri = tri;
    B.maxLeaf = maxLeaf;
    B.tbox.resize(n);
    B.cent.resize(size_t(n) * 3);
    B.order.resize(n);
    Box scene;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < int64_t(n); ++i) {
        Box b;
        b.grow(tri + 9 * i); b.grow(tri + 9 * i + 3); b.grow(tri + 9 * i + 6);
        B.tbox[i] = b;
        for (int a = 0; a < 3; ++a) B.cent[3 * i + a] = 0.5f * (b.lo[a] + b.hi[a]);
        B.order[i] = uint32_t(i);
    }
    for (uint32_t i = 0; i < n; ++i) scene.grow(B.tbox[i]);
    // conservative padding: 2^-15 of the largest absolute coordinate / extent of the scene
    float mag = 0.f;
    for (int a = 0; a < 3; ++a) mag = std::max(mag, std::max(std::fabs(scene.lo[a]), std::fabs(scene.hi[a])));
    out.pad = std::max(mag * (1.f / 32768.f), 1e-30f);

    B.pool.resize(size_t(2) * n + 1);
    const int32_t root = B.alloc();
#pragma omp parallel
#pragma omp single nowait
    B.build(root, 0, n, 0);

    // flatten: every inner build node becomes one BvhNode holding its two children's boxes. Layout: the root at index 0,
    // then SIBLING PAIRS at even indices (two 64-byte records = one 128-byte line: the far sibling popped from the stack
    // later is usually still in L1), subtrees placed depth-first so a descent walks forward through memory.
    const int32_t nBuild = B.next.load();
    std::vector<int32_t> innerIndex(nBuild, -1);
    int32_t nInner = 0;
    if (B.pool[root].left >= 0) {
        innerIndex[root] = 0;
        nInner = 2; // slot 1 pads the root to a full line
        std::vector<int32_t> stack{root};
        while (!stack.empty()) {
            const int32_t bn = stack.back();
            stack.pop_back();
            const int32_t l = B.pool[bn].left, r = B.pool[bn].right;
            const bool li = B.pool[l].left >= 0, ri = B.pool[r].left >= 0;
            if (li || ri) {
                if (li) innerIndex[l] = nInner;
                if (ri) innerIndex[r] = nInner + (li ? 1 : 0);
                nInner += 2;
                if (ri) stack.push_back(r);
                if (li) stack.push_back(l); // left subtree first
            }
        }
    }
    out.triOrder = B.order;
    const float rootArea = std::max(B.pool[root].box.area(), 1e-30f);
    double cost = 0.0;
    int depth = 0;
    if (B.pool[root].left < 0) {
        // single leaf: root with one real child
        out.nodes.resize(1);
        setChild(out.nodes[0], 0, B.pool[root].box, out.pad, int32_t(B.pool[root].first), int32_t(B.pool[root].count));
        setEmpty(out.nodes[0], 1);
        out.depth = 1;
        out.sahCost = float(B.pool[root].count);
        return;
    }
    out.nodes.resize(nInner);
    for (auto& nd : out.nodes) { setEmpty(nd, 0); setEmpty(nd, 1); } // padding slots are never referenced
    for (int32_t i = 0; i < nBuild; ++i) {
        const BuildNode& bn = B.pool[i];
        depth = std::max(depth, bn.depth + 1);
        if (bn.left < 0) { cost += double(bn.box.area() / rootArea) * bn.count; continue; }
        cost += double(bn.box.area() / rootArea);
        BvhNode& o = out.nodes[innerIndex[i]];
        const int32_t ch[2] = {bn.left, bn.right};
        for (int s = 0; s < 2; ++s) {
            const BuildNode& c = B.pool[ch[s]];
            if (c.left >= 0) setChild(o, s, c.box, out.pad, innerIndex[ch[s]], 0);
            else setChild(o, s, c.box, out.pad, int32_t(c.first), int32_t(c.count));
        }
    }
    out.depth = depth;
    out.sahCost = float(cost);
}

} // namespace xrt
