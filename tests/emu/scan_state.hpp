// Host-side stand-in for kicp_scan_next (kinematic-icp_b200/csrc/kicp_map.cu): the status words, the ticket counter and the launch
// number of the single-pass scans (kicp_scan.cuh), kept across launches exactly as a context keeps them — status words are never
// cleared between launches, the ticket is put back by the kernels themselves.  Test infrastructure.
#pragma once
#include <cstdint>
#include <vector>

#include "kicp_internal.h"

namespace emu {
struct ScanState {
    std::vector<unsigned long long> status;
    unsigned int ticket = 0;
    uint32_t launch = 0;
    kicp_scan_args next(int64_t items, int tile) {
        const size_t tiles = (size_t)((items + tile - 1) / tile) + 1;
        if (tiles > status.size()) status.assign(tiles + tiles / 2, 0ull), launch = 0;
        if (++launch == 0) std::fill(status.begin(), status.end(), 0ull), launch = 1;
        return kicp_scan_args{status.data(), &ticket, launch};
    }
};
inline ScanState &scan_state() {
    static ScanState s;
    return s;
}
}  // namespace emu
