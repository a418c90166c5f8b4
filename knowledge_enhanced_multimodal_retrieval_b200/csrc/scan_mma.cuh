// placeholder replaced by the tcgen05 kernel
#pragma once
#include "scan_warp.cuh"
namespace kemr {
struct MmaPlan { int parts = 0; int q_pad = 0; };
inline bool mma_built() { return false; }
inline bool mma_supported(int, int) { return false; }
inline int mma_make_plan(int, int64_t, int, int, int, int, int, MmaPlan*) { return 1; }
inline int mma_launch(const ScanArgs&, const MmaPlan&, cudaStream_t) { return 1; }
inline const char* mma_last_error() { return "tcgen05 path not built"; }
}
