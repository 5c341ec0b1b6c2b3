// kernels_exact.cu — parity instantiation of the wavefront kernels. MUST be compiled with -fmad=false so that
// no multiply-add is contracted: the CPU reference (x86-64 baseline, no FMA) rounds every product and sum.
#define XRT_EXACT 1
#define XRT_NS exact
#include <algorithm>
#include "wavefront.cuh"
#include "kernels.h"
namespace xrt {
const KernelTable& exactKernels()
{
    using namespace exact;
    static const KernelTable t = {launchSeedMt, launchRaygen, launchPrimary, launchPrimaryMasks, launchExtend, launchConnect, launchShadeSurface, launchBounceSmall, launchShadeVolume, launchVolumePaths,
                                  launchAccumulate, launchFinalize, launchTraceRays, launchScatterPrimaryHits, launchGenJitter};
    return t;
}
} // namespace xrt
