// FP64 tensor-core (mma.sync DMMA) level kernel for 20- and 61-state models.
#pragma once
#include "pb_kernels.cuh"

namespace pbk {

inline bool dmma_supported(int) { return false; }   // built in a later milestone

inline int launch_level_dmma(int, const LevelNode*, const LevelChild*, int, int, const double*, int, int,
                             const uint8_t*, int64_t, double*, int64_t, int64_t, const double*, double, int,
                             double*, int, cudaStream_t)
{
    return -1;
}

}  // namespace pbk
