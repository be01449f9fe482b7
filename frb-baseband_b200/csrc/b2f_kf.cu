// Fused column + row kernel instantiations for one row length (-DB2F_FR=16|32|64|128|256|512).  Two translation units per
// row length halve the longest compile of the build (kf512: 7 minutes in one piece): -DB2F_KF_PART=1 holds the JA98
// instantiations behind b2f_launch_kfj_<R>, the default part the static-level ones and the dispatch.
#include "b2f_fused.cuh"
#include "b2f_launch.h"

using namespace b2f;

#define B2F_CAT2(a, b) a##b
#define B2F_CAT(a, b) B2F_CAT2(a, b)

template <int MODE, bool JA98>
static cudaError_t go2(const FParams& p, int grid, int cooperative, cudaStream_t st, int* max_ctas_per_sm) {
    auto kern = kf_fused<B2F_FR, MODE, JA98>;
    const size_t smem = FGeo<B2F_FR>::kBytes;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (max_ctas_per_sm) {                        // query only
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_ctas_per_sm, kern, kFThreads, smem);
    }
    if (cooperative) {
        FParams q = p;
        void* args[] = {&q};
        return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kFThreads), args, smem, st);
    }
    kern<<<grid, kFThreads, smem, st>>>(p);
    return cudaGetLastError();
}

// the dynamic-level decode is instantiated for the products the north star names (I, coherence, IQUV)
#if defined(B2F_KF_PART) && B2F_KF_PART == 1
cudaError_t B2F_CAT(b2f_launch_kfj_, B2F_FR)(int mode, const FParams& p, int grid, int cooperative, cudaStream_t st, int* max_ctas_per_sm) {
    switch (mode) {
        case B2F_POL_I: return go2<B2F_POL_I, true>(p, grid, cooperative, st, max_ctas_per_sm);
#ifndef B2F_KF_ONLY_I
        case B2F_POL_COHERENCE: return go2<B2F_POL_COHERENCE, true>(p, grid, cooperative, st, max_ctas_per_sm);
        case B2F_POL_IQUV: return go2<B2F_POL_IQUV, true>(p, grid, cooperative, st, max_ctas_per_sm);
#endif
    }
    return cudaErrorInvalidValue;
}
#else
cudaError_t B2F_CAT(b2f_launch_kfj_, B2F_FR)(int mode, const FParams& p, int grid, int cooperative, cudaStream_t st, int* max_ctas_per_sm);

template <int MODE>
static cudaError_t go(const FParams& p, int grid, int cooperative, cudaStream_t st, int* max_ctas_per_sm) {
    if (p.levels) return B2F_CAT(b2f_launch_kfj_, B2F_FR)(MODE, p, grid, cooperative, st, max_ctas_per_sm);
    return go2<MODE, false>(p, grid, cooperative, st, max_ctas_per_sm);
}

cudaError_t B2F_CAT(b2f_launch_kf_, B2F_FR)(int mode, const FParams& p, int grid, int cooperative, cudaStream_t st,
                                            int* max_ctas_per_sm) {
    switch (mode) {
#if !defined(B2F_KF_ONLY_I) && !defined(B2F_KF_MAIN_MODES)
        case B2F_POL_P0: return go<B2F_POL_P0>(p, grid, cooperative, st, max_ctas_per_sm);
        case B2F_POL_P1: return go<B2F_POL_P1>(p, grid, cooperative, st, max_ctas_per_sm);
#endif
        case B2F_POL_I: return go<B2F_POL_I>(p, grid, cooperative, st, max_ctas_per_sm);
#if !defined(B2F_KF_ONLY_I) && !defined(B2F_KF_MAIN_MODES)
        case B2F_POL_I2: return go<B2F_POL_I2>(p, grid, cooperative, st, max_ctas_per_sm);
        case B2F_POL_PPQQ: return go<B2F_POL_PPQQ>(p, grid, cooperative, st, max_ctas_per_sm);
#endif
#ifndef B2F_KF_ONLY_I
        case B2F_POL_COHERENCE: return go<B2F_POL_COHERENCE>(p, grid, cooperative, st, max_ctas_per_sm);
        case B2F_POL_IQUV: return go<B2F_POL_IQUV>(p, grid, cooperative, st, max_ctas_per_sm);
#endif
    }
    return cudaErrorInvalidValue;
}
#endif
