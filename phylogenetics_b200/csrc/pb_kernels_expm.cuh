// Batched Pade scaling-and-squaring matrix exponential (Matrix.expm, lib/linear_algebra.ml:308-404).
#pragma once
#include <cuda_runtime.h>

namespace pbk {

inline int launch_expm_batched(int, int, int, int, const double*, const double*, const double*, double*, size_t,
                               cudaStream_t)
{
    return -1;   // built in a later milestone
}

}  // namespace pbk
