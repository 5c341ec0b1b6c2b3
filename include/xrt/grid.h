// xrt/grid.h — DensityGrid of the drop-in API (reference grid.h:9-85). OpenVDB is an absent third-party
// dependency of the reference, so the concrete grid here is DenseGrid: a dense fp32 block with the
// semantics OpenVDBGrid exposes (index->world affine, trilinear BoxSampler lookup, active-voxel bounds,
// max value). The lookup itself runs on the GPU (kernels: delta tracking).
#pragma once
#include <vector>
#include "ray.h"
#include <xrtgpu.h>

class DensityGrid {
public:
    virtual ~DensityGrid() = default;
    virtual AABB getBounds() const = 0;
    virtual float getMaxDensity() const = 0;
    // C-ABI description; data pointer must stay valid while the grid lives
    virtual bool describe(xrtg_grid& out) const = 0;
};

class DenseGrid : public DensityGrid {
public:
    // voxels: nx*ny*nz floats, x fastest; world = origin + voxelSize * index
    DenseGrid(int nx, int ny, int nz, std::vector<float> voxels, Vec3f origin, float voxelSize, float background = 0.0f)
        : nx(nx), ny(ny), nz(nz), voxels(std::move(voxels)), origin(origin), voxelSize(voxelSize), background(background)
    {
        lo[0] = lo[1] = lo[2] = 0; hi[0] = nx - 1; hi[1] = ny - 1; hi[2] = nz - 1;
        int mn[3] = {nx, ny, nz}, mx[3] = {-1, -1, -1};
        maxDensity = background;
        for (int z = 0; z < nz; ++z) for (int y = 0; y < ny; ++y) for (int x = 0; x < nx; ++x) {
            const float v = this->voxels[(size_t(z) * ny + y) * nx + x];
            if (v != background) {
                mn[0] = std::min(mn[0], x); mn[1] = std::min(mn[1], y); mn[2] = std::min(mn[2], z);
                mx[0] = std::max(mx[0], x); mx[1] = std::max(mx[1], y); mx[2] = std::max(mx[2], z);
            }
            maxDensity = std::max(maxDensity, v);
        }
        if (mx[0] >= 0) for (int a = 0; a < 3; ++a) { lo[a] = mn[a]; hi[a] = mx[a]; }
    }
    AABB getBounds() const override
    {
        AABB r; // indexToWorld(bbox.getStart()/getEnd()), getEnd = max+1 (grid.h:58-69)
        for (int a = 0; a < 3; ++a) { r.pMin[a] = origin[a] + voxelSize * float(lo[a]); r.pMax[a] = origin[a] + voxelSize * float(hi[a] + 1); }
        return r;
    }
    float getMaxDensity() const override { return maxDensity; }
    bool describe(xrtg_grid& g) const override
    {
        g.nx = nx; g.ny = ny; g.nz = nz; g.data = voxels.data(); g.voxel_size = voxelSize; g.background = background;
        for (int a = 0; a < 3; ++a) { g.origin[a] = origin[a]; g.active_min[a] = lo[a]; g.active_max[a] = hi[a]; }
        g.max_density = maxDensity;
        return true;
    }

private:
    int nx, ny, nz;
    std::vector<float> voxels;
    Vec3f origin;
    float voxelSize, background, maxDensity;
    int lo[3], hi[3];
};
