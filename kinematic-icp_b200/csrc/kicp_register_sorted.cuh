// Interface between the voxel-sorted registration engine (kicp_register_sorted.cu) and the host-side API
// (kicp_register_api.cu).  SortedState stays private to the kernel file.
#pragma once
#include "kicp_register.cuh"

struct SortedState;

struct SortedArgs {
    SortedState *st;
    kicp_reg_result *result;        // the device-side result block of the context (the one kicp_device_result() names)
    double *dbg;                    // [KICP_MAX_ITERATIONS][6] per-pass timing probes (ns), or nullptr
    unsigned long long *stats;      // [4] work counters (option "stats")
    ScanView scan;
    MapView map;
    double *partials;               // [2][grid][8]
    UploadArgs up;
    RegArgs init;
    int pow2_voxel;
    int collect_stats;
    // scratch of the once-per-registration sort (all sized by the frame)
    unsigned long long *bin_key;    // [bin_mask + 1] voxel key of a slot, all ones = empty (left empty by every launch)
    unsigned int *bin_cnt;          // [bin_mask + 1] points of the slot's voxel, then their first sorted position (left zero)
    uint32_t bin_mask;              // slots - 1 (a power of two, at least two slots per point)
    uint2 *pslot;                   // [n] {slot, rank inside the slot's voxel} of every frame point
    double *sorted;                 // [n][4] the frame in voxel order {x, y, z, 0}
    unsigned int *nn_g;             // [n] by sorted position: the neighbour the previous pass found (0xFFFFFFFF = none)
    kicp_reg_result *result_host;   // optional: device-visible alias of the caller's page-locked result block
    unsigned long long timeout_ns;  // device-side waits (grid barriers, upload flags) give up after this long
};

size_t ks_state_bytes();
int ks_threads();
cudaError_t ks_prepare(int *ctas_per_sm);  // occupancy of k_register_sorted
cudaError_t ks_launch(int grid, SortedArgs &ka, cudaStream_t stream);
