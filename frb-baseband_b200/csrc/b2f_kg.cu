// Generic channeliser with compile-time geometry (b2f_generic.cuh): one translation unit per size, -DB2F_KG_LG=10..13
// (L = R = 1024 ... 8192, i.e. process_vdif --nchan 512 ... 4096 with its -F nchan:2*nchan rule).
#include "b2f_generic.cuh"
#include "b2f_launch.h"

using namespace b2f;

#ifndef B2F_KG_LG
#error "compile with -DB2F_KG_LG=10|11|12|13"
#endif
#define KG_CAT_(a, b) a##b
#define KG_CAT(a, b) KG_CAT_(a, b)

template <int MODE>
static cudaError_t row_go(const KGParams& p, int grid, cudaStream_t st, int* ctas) {
    using G = KGT<B2F_KG_LG>;
    auto kern = kgt_row_pass<B2F_KG_LG, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kRowSmem);
    if (e != cudaSuccess) return e;
    if (ctas) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, kern, G::kRowThreads, G::kRowSmem);
    kern<<<grid, G::kRowThreads, G::kRowSmem, st>>>(p);
    return cudaGetLastError();
}

// grid <= 0: only report how many CTAs fit on one SM (*ctas)
template <int NBIT>
static cudaError_t col_go(const KGParams& p, int grid, cudaStream_t st, int* ctas) {
    using G = KGT<B2F_KG_LG>;
    auto kern = kgt_column_pass<B2F_KG_LG, NBIT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kColSmem);
    if (e != cudaSuccess) return e;
    if (ctas) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, kern, G::kColThreads, G::kColSmem);
    kern<<<grid, G::kColThreads, G::kColSmem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t KG_CAT(b2f_launch_kgt_col_, B2F_KG_LG)(int in_nbit, const KGParams& p, int grid, cudaStream_t st, int* ctas) {
    // in_nbit: 8, 2 (1-bit input arrives as the same index bytes), or 22 = 2-bit input with JA98 levels
    return in_nbit == 8 ? col_go<8>(p, grid, st, ctas) : (in_nbit == 22 ? col_go<22>(p, grid, st, ctas) : col_go<2>(p, grid, st, ctas));
}

cudaError_t KG_CAT(b2f_launch_kgt_row_, B2F_KG_LG)(int mode, const KGParams& p, int grid, cudaStream_t st, int* ctas) {
    switch (mode) {
        case B2F_POL_P0: return row_go<B2F_POL_P0>(p, grid, st, ctas);
        case B2F_POL_P1: return row_go<B2F_POL_P1>(p, grid, st, ctas);
        case B2F_POL_I: return row_go<B2F_POL_I>(p, grid, st, ctas);
        case B2F_POL_I2: return row_go<B2F_POL_I2>(p, grid, st, ctas);
        case B2F_POL_PPQQ: return row_go<B2F_POL_PPQQ>(p, grid, st, ctas);
        case B2F_POL_COHERENCE: return row_go<B2F_POL_COHERENCE>(p, grid, st, ctas);
        case B2F_POL_IQUV: return row_go<B2F_POL_IQUV>(p, grid, st, ctas);
    }
    return cudaErrorInvalidValue;
}
