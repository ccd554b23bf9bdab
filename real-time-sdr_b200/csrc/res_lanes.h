// Lane tables of the 247/640 RDS resampler in k_rds_backend (host code, no CUDA types: sdr_chain.cu uploads them, the
// CPU test-suite runs the same function through tests/cpp/pllmath_host.cpp and replays the kernel's loop on them).
// /root/reference/src/filter.cpp:123-147 (convolveFIR with up/down), src/rds.cpp:61,130.
//
// A lane owns two ADJACENT output residues tp_lo, tp_hi = tp_lo + 1 (output n belongs to residue n % 247), whose input
// offsets floor(640 tp / 247) differ by delta = 2 or 3, and reads ONE input sample per loop step for both: at step j it
// holds x[base + 640 q - j], which is tap j - lag_hi of the `hi` residue (lag_hi = base - floor(640 tp_hi / 247)) and
// tap j - lag_lo of the `lo` residue (lag_lo = lag_hi + delta <= kResLagMax).  Each residue still accumulates its own 101
// branch taps h[phase + 247 j'] in ascending order, so the arithmetic is the reference's; the loop has kResLagMax extra
// steps, and in its first and last kResLagMax steps a residue whose tap index is outside 0..100 skips the MAC.
// lag_hi is the freedom used to give the 32 lanes of a warp bases that are distinct modulo 32, which makes the kernel's
// shared-memory loads conflict free.  One residue is left over (247 is odd): a lane with tp_lo = -1.
#pragma once

#include <algorithm>
#include <functional>
#include <vector>

namespace sdrb {

constexpr int kRdsUp = 247, kRdsDown = 640;
constexpr int kResTaps = 101;                   // = kTaps (checked in sdr_kernels.cuh)
constexpr int kResQ = 12;                       // outputs per residue and block: ceil(2836 / 247); the kernel requires n_out <= 12 * 247
constexpr int kResLagMax = 3;
constexpr int kResIter = kResTaps + kResLagMax; // loop steps
constexpr int kResLanes = 128;                  // lanes; threads t and t + 128 share a lane, t / 128 selects outputs q = 0..5 or 6..11

struct ResLane {   // uploaded as int4
    int tp_hi;     // -1: idle lane
    int tp_lo;     // -1: no partner
    int base;      // floor(640 tp_hi / 247) + lag_hi
    int lags;      // lag_hi | lag_lo << 8  (lag_lo = 0xFF without a partner)
};

// lh: the 101 * 247 taps of impulseResponseLPF(240000 * 247, 3e3, 101 * 247, h, 247).  lanes[kResLanes];
// rtaps[kResIter][kResLanes][2]: (.x, .y) = the taps of the hi / lo residue that loop step j uses (0 where there is none).
// Returns the number of residue pairs that could not be given a conflict-free lane (0 for 247/640), or -1 if the lanes ran out.
inline int build_res_lanes(const float* lh, ResLane* lanes, float* rtaps) {
    struct Item { int hi, lo, delta; };
    std::vector<Item> items;
    auto off = [](int tp) { return (kRdsDown * tp) / kRdsUp; };
    for (int tp = 0; tp + 1 < kRdsUp; tp += 2) items.push_back({tp + 1, tp, off(tp + 1) - off(tp)});
    if (kRdsUp % 2) items.push_back({kRdsUp - 1, -1, 0});
    constexpr int kWarps = kResLanes / 32;
    // Bipartite matching of items to (warp, bank) slots by augmenting paths: an item may sit in bank (off(hi) + e) % 32
    // for any lag e it can afford, a bank takes one item per warp.
    std::vector<std::vector<int>> in_bank(32);
    std::vector<int> lag(items.size(), -1);
    std::function<bool(int, int, std::vector<char>&)> place = [&](int i, int from, std::vector<char>& seen) -> bool {
        const int max_lag = kResLagMax - items[i].delta;
        for (int e = 0; e <= max_lag; e++) {
            const int bank = (off(items[i].hi) + e) % 32;
            if (bank == from || seen[bank]) continue;
            seen[bank] = 1;
            bool room = (int)in_bank[bank].size() < kWarps;
            for (size_t k = 0; !room && k < in_bank[bank].size(); k++) room = place(in_bank[bank][k], bank, seen);
            if (room) {
                if (from >= 0) in_bank[from].erase(std::find(in_bank[from].begin(), in_bank[from].end(), i));
                in_bank[bank].push_back(i);
                lag[i] = e;
                return true;
            }
        }
        return false;
    };
    std::vector<int> leftover;
    for (size_t i = 0; i < items.size(); i++) {
        std::vector<char> seen(32, 0);
        if (items[i].delta > kResLagMax || !place((int)i, -1, seen)) leftover.push_back((int)i);
    }
    auto lane_of = [&](const Item& it, int e) { return ResLane{it.hi, it.lo, off(it.hi) + e, e | ((it.lo >= 0 ? e + it.delta : 0xFF) << 8)}; };
    for (int tl = 0; tl < kResLanes; tl++) lanes[tl] = ResLane{-1, -1, 0, 0};
    for (int bank = 0; bank < 32; bank++)
        for (size_t w = 0; w < in_bank[bank].size(); w++) lanes[32 * w + bank] = lane_of(items[in_bank[bank][w]], lag[in_bank[bank][w]]);
    for (int i : leftover) {  // any free lane (bank conflicts, never a wrong result); a pair beyond the lag budget as two single residues
        std::vector<Item> parts;
        if (items[i].delta > kResLagMax) { parts.push_back({items[i].hi, -1, 0}); parts.push_back({items[i].lo, -1, 0}); }
        else parts.push_back(items[i]);
        for (const Item& it : parts) {
            ResLane* f = std::find_if(lanes, lanes + kResLanes, [](const ResLane& l) { return l.tp_hi < 0; });
            if (f == lanes + kResLanes) return -1;
            *f = lane_of(it, 0);
        }
    }
    std::fill(rtaps, rtaps + (size_t)kResIter * kResLanes * 2, 0.0f);
    for (int tl = 0; tl < kResLanes; tl++) {
        const ResLane& ln = lanes[tl];
        if (ln.tp_hi < 0) continue;
        const int lag_hi = ln.lags & 0xFF, lag_lo = (ln.lags >> 8) & 0xFF;
        for (int j = 0; j < kResTaps; j++) {  // tap j of a residue is used at loop step j + lag
            rtaps[((size_t)(j + lag_hi) * kResLanes + tl) * 2] = lh[(kRdsDown * ln.tp_hi) % kRdsUp + kRdsUp * j];
            if (ln.tp_lo >= 0) rtaps[((size_t)(j + lag_lo) * kResLanes + tl) * 2 + 1] = lh[(kRdsDown * ln.tp_lo) % kRdsUp + kRdsUp * j];
        }
    }
    return (int)leftover.size();
}

}  // namespace sdrb
