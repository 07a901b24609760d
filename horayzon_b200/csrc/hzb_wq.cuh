// hzb_wq.cuh -- constants shared by the warp-queue traversal kernels (block size, stack depth).
#pragma once
#include "hzb_geom.cuh"

namespace hzb {

constexpr int WQ_BLOCK = 128;                 // threads per CTA
constexpr int WQ_NWARPS = WQ_BLOCK / 32;
#ifndef HZB_WQ_STACK_N
#define HZB_WQ_STACK_N 26
#endif
constexpr int WQ_STACK_N = HZB_WQ_STACK_N;    // shared-memory stack entries per lane.  26 (+3 spare rows) keeps six CTAs per SM inside the
                                              // 164 KB shared-memory carve-out (92 KB of L1 left); a full stack ends the walk safely (hzb_wq2.cuh):
                                              // the cell is recomputed by the fix-up kernel (2 of 1.7e10 packets of the 6000 x 6000 x 360 pass, none on the other scenes).
constexpr uint32_t WQ_NONE = 0xFFFFFFFFu;

}  // namespace hzb
