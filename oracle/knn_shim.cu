// TEST INFRASTRUCTURE — C-ABI shim around the UNMODIFIED reference simple-knn
// (/root/reference/submodules/simple-knn/simple_knn.cu), compiled by oracle/build.py into
// oracle/_ref/libref_knn.so.  Used only by tests/ as the bit-exact checker of csrc/knn.cu.
#include <cuda_runtime.h>
#include "simple_knn.h"

extern "C" int ref_knn(int P, const float* points_dev, float* mean_dist2_dev) {
    SimpleKNN::knn(P, (float3*)points_dev, mean_dist2_dev);
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : (int)e;
}
