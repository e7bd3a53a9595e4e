// futbol_selftest_arith: the guard-free IEEE sequences of ieee_fast.cuh against nvcc's builtins, element-wise.
#include <cuda_runtime.h>
#include <stdint.h>
#include "ieee_fast.cuh"

namespace futbol {

__global__ void selftest_arith_kernel(const double *a, const double *b, unsigned long long *mismatch, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long bad_div = 0, bad_div2 = 0, bad_sqrt = 0;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double x = a[i], y = b[i];
        const double ay = fabs(y);
        if (ay != 0.0) {
            bad_div += __double_as_longlong(fdiv(x, y)) != __double_as_longlong(__ddiv_rn(x, y)) && !(x == 0.0);
            bad_div += x == 0.0 && fdiv(x, y) != 0.0;
            double q1, q2;
            fdiv2(x, ay, y, q1, q2);
            bad_div2 += (__double_as_longlong(q1) != __double_as_longlong(__ddiv_rn(x, y)) && !(x == 0.0)) ||
                        __double_as_longlong(q2) != __double_as_longlong(__ddiv_rn(ay, y));
            bad_sqrt += __double_as_longlong(fsqrt(ay)) != __double_as_longlong(__dsqrt_rn(ay));
        }
    }
    if (bad_div) atomicAdd(&mismatch[0], bad_div);
    if (bad_div2) atomicAdd(&mismatch[1], bad_div2);
    if (bad_sqrt) atomicAdd(&mismatch[2], bad_sqrt);
}

cudaError_t launch_selftest_arith(const double *a, const double *b, unsigned long long *mismatch, size_t n, cudaStream_t st)
{
    selftest_arith_kernel<<<148 * 8, 256, 0, st>>>(a, b, mismatch, n);
    return cudaGetLastError();
}

}  // namespace futbol
