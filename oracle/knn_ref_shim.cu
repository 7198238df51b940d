// knn_ref_shim.cu -- TEST INFRASTRUCTURE, not product code.
//
// C-ABI wrapper around the UNMODIFIED reference `SimpleKNN::knn` (submodules/simple-knn/simple_knn.cu:185-220), compiled
// from the source where it lies under /root/reference by oracle/Makefile (`make knn_ref`) into
// oracle/_ref/libsimpleknn_ref.so.  Nothing is copied: this file only includes the reference's header through the -I path.
// Used by tests/test_knn.py and bench.py's `init_knn` key on the GPU box (the built .so travels, /root/reference does not).
#include <cuda_runtime.h>

#include "simple_knn.h"

extern "C" int ref_knn_mean_dist2(int P, const float* points_device, float* mean_dist2_device) {
    if (P <= 0) return 0;
    SimpleKNN::knn(P, (float3*)points_device, mean_dist2_device);     // what spatial.cu:15-26 (distCUDA2) calls
    return (int)cudaDeviceSynchronize();
}
