// Host-side planner of the filter-bank stage of the fused waveform -> MFCC kernel (stft512.cu, kFmtMfcc).
//
// The kernel evaluates  mel[c] = sum_k H[k, c] amp[k]  (fbank.py:315-316) for the four frames of a quad with the
// non-zero support [begin_c, end_c) of every filter cut into SEGMENTS of <= 8 consecutive bins: a lane takes a
// segment, walks its 8 bins in the two staged amplitude rows (64-bit shared-memory loads) and leaves one partial
// sum per frame.  Shared memory serves a 64-bit load per half-warp, 16 lanes x 8 bytes over 32 banks: two lanes of
// the same half-warp collide when their segment starts differ by a non-zero multiple of 16 bins.  Round 1 cut the
// supports in order and measured 25 % of all shared-memory wavefronts of the kernel as bank conflicts
// (profiles/r1_mfcc_wave_v3.txt).  A segment may start anywhere at or before its first needed bin -- the bins in
// front get zero weights -- so this planner chooses every start and every (round, half-warp, lane) slot such that
// the starts of a half-warp are distinct modulo 16 (or equal: a broadcast), whenever that is possible.
//
// Plan layout (int32): [0] n_slots (multiple of 32, 0 = no plan), [1] n_channel, [2] conflicts left, [3] reserved,
// then seg_start[128], seg_first[128], seg_last[128] (bins [first, last) carry weights), seg_channel[128] (-1 =
// padding slot), channel_slots[n_channel][8] (slot ids, 128 = "no more").
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "mfcc_plan.h"

namespace dsb200 {

int mfcc_plan_build_host(const int32_t* cb, const int32_t* ce, int C, int K, int32_t* plan) {
  const int n_ints = mfcc_plan_ints(C);
  std::fill(plan, plan + n_ints, 0);
  int32_t* seg_start = plan + 4;
  int32_t* seg_first = seg_start + kPlanMaxSeg;
  int32_t* seg_last = seg_first + kPlanMaxSeg;
  int32_t* seg_ch = seg_last + kPlanMaxSeg;
  int32_t* ch_slots = seg_ch + kPlanMaxSeg;
  std::fill(seg_ch, seg_ch + kPlanMaxSeg, -1);
  std::fill(ch_slots, ch_slots + C * kPlanSlotsPerCh, kPlanMaxSeg);
  plan[1] = C;

  struct Piece { int c, a, b; };
  std::vector<Piece> pieces;
  for (int c = 0; c < C; ++c) {
    const int a = cb[c], b = ce[c];
    if (a < 0 || b > K || a > b) return 0;           // not a support: no plan (the kernel walks H densely)
    const int len = b - a;
    if (len == 0) continue;
    const int n = (len + kPlanSegLen - 1) / kPlanSegLen;
    if (n > kPlanSlotsPerCh) return 0;
    for (int i = 0; i < n; ++i)                        // near-equal pieces: every piece keeps some slack
      pieces.push_back({c, a + static_cast<int>(static_cast<int64_t>(len) * i / n),
                        a + static_cast<int>(static_cast<int64_t>(len) * (i + 1) / n)});
  }
  const int n_pieces = static_cast<int>(pieces.size());
  if (n_pieces == 0 || n_pieces > kPlanMaxSeg) return 0;
  const int n_slots = (n_pieces + 31) / 32 * 32;
  const int n_groups = n_slots / 16;                   // half-warp phases
  // least flexible pieces first
  std::stable_sort(pieces.begin(), pieces.end(), [](const Piece& x, const Piece& y) { return (x.b - x.a) > (y.b - y.a); });
  std::vector<int> fill(n_groups, 0);
  std::vector<int> owner(n_groups * 16, -1);           // start that occupies (group, start mod 16), -1 = free
  std::vector<int> per_ch(C, 0);
  int conflicts = 0;
  const int max_start = kPlanAmpPitch - kPlanSegLen;   // the staged rows are zero padded up to kPlanAmpPitch bins
  for (const Piece& p : pieces) {
    const int lo = std::max(0, p.b - kPlanSegLen), hi = std::min(p.a, max_start);
    int best_g = -1, best_s = -1;
    for (int pass = 0; pass < 2 && best_g < 0; ++pass) {
      // pass 0: a free residue class (or an equal start) in the emptiest group; pass 1: any free lane
      int best_fill = 17;
      for (int g = 0; g < n_groups; ++g) {
        if (fill[g] >= 16 || fill[g] >= best_fill) continue;
        for (int s = hi; s >= lo; --s) {
          const int o = owner[g * 16 + (s & 15)];
          if (pass == 1 || o < 0 || o == s) {
            best_fill = fill[g];
            best_g = g;
            best_s = s;
            break;
          }
        }
      }
      if (pass == 1 && best_g >= 0) ++conflicts;
    }
    if (best_g < 0) return 0;
    const int slot = best_g * 16 + fill[best_g]++;
    if (owner[best_g * 16 + (best_s & 15)] < 0) owner[best_g * 16 + (best_s & 15)] = best_s;
    seg_start[slot] = best_s;
    seg_first[slot] = p.a;
    seg_last[slot] = p.b;
    seg_ch[slot] = p.c;
    ch_slots[p.c * kPlanSlotsPerCh + per_ch[p.c]++] = slot;
  }
  // padding slots read bins [0, 8) of the rows with zero weights: give each the free residue class of its group
  for (int g = 0; g < n_groups; ++g)
    for (int i = fill[g]; i < 16; ++i) {
      int r = 0;
      while (r < 16 && owner[g * 16 + r] >= 0) ++r;
      if (r < 16) owner[g * 16 + r] = r;
      seg_start[g * 16 + i] = r < 16 ? r : 0;
    }
  plan[0] = n_slots;
  plan[2] = conflicts;
  return n_slots;
}

}  // namespace dsb200

extern "C" {

int32_t dsb200_mfcc_plan_ints(int32_t n_channel) { return n_channel > 0 ? dsb200::mfcc_plan_ints(n_channel) : 0; }

int dsb200_mfcc_plan_build(const int32_t* col_begin, const int32_t* col_end, int32_t n_channel, int32_t n_bins,
                           int32_t* plan) {
  using namespace dsb200;
  DSB_REQUIRE(col_begin != nullptr && col_end != nullptr && plan != nullptr, "NULL argument");
  DSB_REQUIRE(n_channel > 0 && n_bins > 0, "n_channel and n_bins must be positive");
  mfcc_plan_build_host(col_begin, col_end, n_channel, n_bins, plan);   // plan[0] == 0: the kernel plans for itself
  return DSB200_OK;
}

}  // extern "C"
