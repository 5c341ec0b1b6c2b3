// kernels_fast.cu — throughput instantiation of the wavefront kernels: FMA contraction allowed, counter-based
// Philox4x32-7 RNG keyed (seed,pixel)/(sample,block), several samples per pixel per wave.
#define XRT_EXACT 0
#define XRT_NS fast
#include <algorithm>
#include "wavefront.cuh"
#include "kernels.h"
namespace xrt {
const KernelTable& fastKernels()
{
    using namespace fast;
    static const KernelTable t = {launchSeedMt, launchRaygen, launchPrimary, launchPrimaryMasks, launchExtend, launchConnect, launchShadeSurface, launchBounceSmall, launchShadeVolume, launchVolumePaths,
                                  launchAccumulate, launchFinalize, launchTraceRays, launchScatterPrimaryHits, launchGenJitter};
    return t;
}
} // namespace xrt
