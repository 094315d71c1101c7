// postprocess.h — host-visible interface of postprocess.cu.
#pragma once
#include <cuda_runtime.h>
#include "tracerboy_b200.h"
#include "launch.h"

namespace tbd {

// Auto exposure (when s.UseAutoExposure) + PostProcessCS on one image. `in` is float4 per pixel, or one float per
// pixel when scalarInput (AOVDepth). hist257: 256 histogram words + the averaged luminance (float bits); it is
// cleared here. out8 may be null.
cudaError_t postprocess(const void* in, bool scalarInput, const float4* aux, uint32_t width, uint32_t height, uint32_t outputType,
                        const TbPostProcessSettings& s, uint32_t* hist257, float4* out, uchar4* out8, int numSMs,
                        cudaStream_t stream, LaunchCounter& lc);

// Realtime temporal accumulation (TemporalAccumulationCS.hlsl) on device images, all float4 per pixel.
cudaError_t temporal_accumulate(const TbTemporalAccumulationParams& p, uint32_t width, uint32_t height, const float4* history,
                                const float4* current, const float4* worldPos, const float4* prevWorldPos, const float4* normals,
                                const float4* momentHistory, float4* outColor, float4* outMoment, cudaStream_t stream, LaunchCounter& lc);

} // namespace tbd
