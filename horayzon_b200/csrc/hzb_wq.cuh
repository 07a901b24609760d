// hzb_wq.cuh -- constants shared by the warp-queue traversal kernels (block size, stack depth).
#pragma once
#include "hzb_geom.cuh"

namespace hzb {

constexpr int WQ_BLOCK = 128;                 // threads per CTA
constexpr int WQ_NWARPS = WQ_BLOCK / 32;
constexpr int WQ_STACK_N = 36;                // shared-memory stack entries per lane (a full stack ends the walk safely, see hzb_wq2.cuh)
constexpr uint32_t WQ_NONE = 0xFFFFFFFFu;

}  // namespace hzb
