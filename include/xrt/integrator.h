// xrt/integrator.h — the integrator menu of the drop-in API (reference integrator.h:9-636). On the GPU
// path an Integrator is a selector: GpuRenderer maps kind()/maxDepth() to the wavefront pipeline that
// implements the same estimator. FurnaceIntegrator is the furnace block that is dead code behind the early
// return of the reference's NormalIntegrator (integrator.h:36 vs :59-66).
#pragma once
#include "scene.h"
#include <xrtgpu.h>

class Integrator {
public:
    Integrator() = default;
    virtual ~Integrator() = default;
    virtual xrtg_integrator kind() const = 0;
    virtual uint32_t maxDepth() const { return 1; }
};

#define XRT_INTEGRATOR_0(Name, Kind)                                  \
    class Name : public Integrator {                                  \
    public:                                                           \
        Name() = default;                                             \
        xrtg_integrator kind() const override { return Kind; }        \
    };
#define XRT_INTEGRATOR_D(Name, Kind, Default)                         \
    class Name : public Integrator {                                  \
    public:                                                           \
        Name(uint32_t maxDepth Default) : m_maxDepth(maxDepth) {}     \
        xrtg_integrator kind() const override { return Kind; }        \
        uint32_t maxDepth() const override { return m_maxDepth; }     \
    private:                                                          \
        const uint32_t m_maxDepth;                                    \
    };

XRT_INTEGRATOR_0(NormalIntegrator, XRTG_INT_NORMAL)             // integrator.h:22-74
XRT_INTEGRATOR_0(FurnaceIntegrator, XRTG_INT_FURNACE)           // integrator.h:59-66
XRT_INTEGRATOR_0(DirectIntegrator, XRTG_INT_DIRECT)             // integrator.h:76-120
XRT_INTEGRATOR_D(IndirectIntegrator, XRTG_INT_INDIRECT, )       // integrator.h:122-190
XRT_INTEGRATOR_D(GIIntegrator, XRTG_INT_GI, )                   // integrator.h:198-291
XRT_INTEGRATOR_D(WhittedIntegrator, XRTG_INT_WHITTED, = 3)      // integrator.h:294-398
XRT_INTEGRATOR_D(VolumePathTracing, XRTG_INT_VOLUME, )          // integrator.h:401-478
XRT_INTEGRATOR_D(VolumePathTracingNEE, XRTG_INT_VOLUME_NEE, )   // integrator.h:481-636
#undef XRT_INTEGRATOR_0
#undef XRT_INTEGRATOR_D
