// Tile row pass (R = 256) instantiations: Stokes I, coherence products, IQUV.
#include "b2f_fused.cuh"
#include "b2f_launch.h"

using namespace b2f;

template <int MODE>
static cudaError_t go(const KTParams& p, int grid, cudaStream_t st) {
    auto kern = kt_row_tiles<MODE>;
    const size_t smem = KTSmem::kBytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kKTThreads, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t b2f_launch_kt(int mode, const KTParams& p, int grid, cudaStream_t st) {
    switch (mode) {
        case B2F_POL_I: return go<B2F_POL_I>(p, grid, st);
        case B2F_POL_COHERENCE: return go<B2F_POL_COHERENCE>(p, grid, st);
        case B2F_POL_IQUV: return go<B2F_POL_IQUV>(p, grid, st);
    }
    return cudaErrorInvalidValue;
}
