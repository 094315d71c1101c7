// launch.h — kernel-launch bookkeeping shared by the CUDA translation units.
#pragma once
#include <cstdint>
namespace tbd {
struct LaunchCounter { uint64_t count = 0; }; // kernels of this library launched (TbRenderStats.KernelLaunches)
} // namespace tbd
