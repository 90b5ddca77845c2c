// Constants shared by the filter-bank planner (mfcc_plan.cu, host) and the fused MFCC kernel (stft512.cu).
#pragma once

#include <cstdint>

namespace dsb200 {

constexpr int kPlanSegLen = 8;        // bins per filter-bank segment
constexpr int kPlanMaxSeg = 128;      // slots (4 rounds of 32 lanes); 40 mel filters at 512 bins need ~85
constexpr int kPlanSlotsPerCh = 8;    // segments per filter: supports of up to 64 bins
constexpr int kPlanAmpPitch = 264;    // float2 units per staged amplitude row: 257 bins + zero padding

constexpr int mfcc_plan_ints(int C) { return 4 + 4 * kPlanMaxSeg + C * kPlanSlotsPerCh; }

// Fills plan[0 .. mfcc_plan_ints(C)); returns the number of slots (0: no plan, the kernel cuts the supports itself).
int mfcc_plan_build_host(const int32_t* col_begin, const int32_t* col_end, int C, int K, int32_t* plan);

}  // namespace dsb200
