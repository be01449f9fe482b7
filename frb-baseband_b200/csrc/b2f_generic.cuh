// Generic channeliser, geometry as template parameters (round 2).
//
// The reference asks digifil for `-F nchan:2*nchan` whenever nchan > 128 (/root/reference/process_vdif.py:162), so the
// shapes its own callers produce are square: freq_res L = row length R = 2 nchan -- the CLI default --nchan 512
// (process_vdif.py:46) is 1024 x 1024, the 1024 channels per IF submit_job.py:58-76 picks at DM 560 are 2048 x 2048,
// and its upper limit of 2^13 band channels over 2 IFs is 8192 x 8192.  The run-time kernels kg_* (b2f_kernels.cuh)
// handle any power-of-two pair; ncu showed half of their issue slots going to index arithmetic (shifts by run-time
// log2 values, swizzles, loop bounds) and sincospif in the innermost step.  Here L = R = 2^LG is a template parameter:
// every shift, mask, trip count and shared-memory offset is a literal, the pass loops are unrolled at compile time,
// and the diagonal W_M^(k2 n1) of the innermost step is built from two sincospif per column (exact integer phase
// reduction) and a doubling tree instead of one sincospif per point.  Same algebra, same pass order, same tables
// (kg_tw_offset) and same intermediate layout as kg_*, so the two agree to rounding.
#pragma once
#include "b2f_kernels.cuh"

namespace b2f {

template <int LG>
struct KGT {
    static constexpr int N = 1 << LG;                       // L = R = N
    static constexpr int NF = (LG - 1) / 4;                 // radix-16 passes outside the innermost step
    static constexpr int LGI = LG - 4 * NF;                 // innermost radix 2^LGI
    static constexpr int RM = 1 << LGI;
    // Column pass: 256 threads on a strip of C columns = 64 KiB of column data, two CTAs per SM.  (Measured alternative:
    // one column pair per 128-thread CTA, four CTAs per SM, conflict-free float4 slots -- 2.1x slower: a warp then
    // touches 32 rows x 16 bytes per global access instead of 16 rows x 32 bytes and mio_throttle went from 3.1 to 8.5.)
    static constexpr int LGC = LG >= 12 ? 1 : 13 - LG;
    static constexpr int C = 1 << LGC;
    static constexpr bool kTwShared = LG <= 12;             // pass twiddles in shared memory (else read through L1)
    static constexpr int kColThreads = 256;
    static constexpr int kColCtas = (size_t)N * C * 8 <= 65536 ? 2 : 1;
    static constexpr size_t kColSmem = ((kTwShared ? (size_t)N : 0) + 32 + (size_t)N * C) * sizeof(float2);
    static constexpr int RB = N >= 4096 ? 1 : 4096 / N;     // rows per row-pass batch
    static constexpr int kRowThreads = LG >= 12 ? 512 : 256;
    static constexpr int CPT = (N / 2) / kRowThreads;       // channels per thread
    static constexpr bool kRowStaged = LG <= 12;            // cp.async double buffer in front of the first pass
    static constexpr int kRowCtas = LG <= 11 ? 2 : 1;
    static constexpr size_t kRowSmem = ((kTwShared ? (size_t)N : 0) + (size_t)RB * N + (kRowStaged ? 2 * (size_t)RB * N : 0)) * sizeof(float2);
};

// digit reversal of a segment number (NF base-16 digits)
template <int NF>
__device__ __forceinline__ int kgt_rev16(int seg) {
    int k = 0;
#pragma unroll
    for (int f = 0; f < NF; ++f) k |= ((seg >> (4 * (NF - 1 - f))) & 15) << (4 * f);
    return k;
}

// One radix-16 column pass over segments of 2^LGN points, two neighbouring columns per thread (see kg_pass16_pair).
// WM (warp-major): warp W works on the thread-iterations [W CNT/8, (W+1) CNT/8) -- the points [W N/8, (W+1) N/8) of every
// column of the strip, i.e. two whole segments of the second pass.  Every stage between the outermost forward and the
// outermost inverse pass stays inside those points, so with this mapping a warp only reads what it wrote itself and
// the stages are separated by __syncwarp instead of a block barrier: 3 block barriers per strip instead of 6 (8 for
// N = 8192), and the warps of a CTA drift into different stages (load burst / butterflies / store burst overlap).
// SRC: 0 shared, 1 index bytes of the 2-bit decode, 2 8-bit sample pairs (KG8: offset and stream-coordinate word masks),
// 3 index bytes with JA98 levels per window of 512 samples (KG8::levels).
template <int LG, int LGN, bool INV, int SRC, int DST, bool WM = false>
__device__ __forceinline__ void kgt_col_pass16(float2* sm, const float2* tw, const uint8_t* gsrc_b, float2* gdst, const float2* lut,
                                               const KG8* s8 = nullptr) {
    using G = KGT<LG>;
    constexpr int LGM = LGN - 4, M1 = 1 << LGM, LGP = G::LGC - 1, CNT = (G::N >> 4) << LGP;
    constexpr int NW = G::kColThreads / 32, PER = WM ? CNT / NW : CNT, STEP = WM ? 32 : G::kColThreads;
    static_assert(!WM || (CNT % (NW * 32) == 0), "whole warp iterations");
    const int t0 = WM ? (threadIdx.x >> 5) * PER : 0;
#pragma unroll 1
    for (int tl = WM ? (threadIdx.x & 31) : threadIdx.x; tl < PER; tl += STEP) {
        const int t = t0 + tl;
        const int c2 = (t & ((1 << LGP) - 1)) * 2, w = t >> LGP;
        const int lo = w & (M1 - 1), seg = w >> LGM;
        const int base = (seg << LGN) + lo;
        const float2* twl = tw + lo;
        float2 a[16], b[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int idx = base + j * M1;
            if (SRC == 1) {
                const uint32_t two = *reinterpret_cast<const uint16_t*>(gsrc_b + (int64_t)idx * G::N + c2);
                a[j] = *reinterpret_cast<const float2*>(reinterpret_cast<const uint8_t*>(lut) + (two & 255u));
                b[j] = *reinterpret_cast<const float2*>(reinterpret_cast<const uint8_t*>(lut) + (two >> 8));
            } else if (SRC == 2) {
                const int64_t o = ((int64_t)idx * G::N + c2) * 2;
                kg_decode8(*reinterpret_cast<const uint32_t*>(gsrc_b + o), *s8, o, a[j], b[j]);
            } else if (SRC == 3) {                                   // index bytes, JA98 levels of the sample's window
                const int64_t o = (int64_t)idx * G::N + c2;
                const uint32_t two = *reinterpret_cast<const uint16_t*>(gsrc_b + o);
                const float4 lv = __ldg(s8->levels + ((s8->byte0 + o) >> 9));
                a[j] = kg_ja98(*reinterpret_cast<const float2*>(reinterpret_cast<const uint8_t*>(lut) + (two & 255u)), lv);
                b[j] = kg_ja98(*reinterpret_cast<const float2*>(reinterpret_cast<const uint8_t*>(lut) + (two >> 8)), lv);
            } else {
                const float4 v4 = *reinterpret_cast<const float4*>(&sm[kg_phys<false>((idx << G::LGC) + c2)]);
                a[j] = make_float2(v4.x, v4.y);
                b[j] = make_float2(v4.z, v4.w);
            }
        }
        if (INV) {
#pragma unroll
            for (int q = 1; q < 16; ++q) {
                const float2 wq = G::kTwShared ? twl[(q - 1) << LGM] : __ldg(twl + ((q - 1) << LGM));
                a[q] = cmul_conj(a[q], wq);
                b[q] = cmul_conj(b[q], wq);
            }
            fft_inreg<16, true>(a);
            fft_inreg<16, true>(b);
        } else {
            fft_inreg<16, false>(a);
            fft_inreg<16, false>(b);
#pragma unroll
            for (int q = 1; q < 16; ++q) {
                const float2 wq = G::kTwShared ? twl[(q - 1) << LGM] : __ldg(twl + ((q - 1) << LGM));
                a[q] = cmul(a[q], wq);
                b[q] = cmul(b[q], wq);
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int idx = base + j * M1;
            const float4 v4 = make_float4(a[j].x, a[j].y, b[j].x, b[j].y);
            if (DST == 1) *reinterpret_cast<float4*>(gdst + (int64_t)idx * G::N + c2) = v4;
            else *reinterpret_cast<float4*>(&sm[kg_phys<false>((idx << G::LGC) + c2)]) = v4;
        }
    }
}

// tw[q] = exp(-2 pi i (base + q step) / 2^LGM), q = 0..RM-1: 1 + log2(RM) sincospif on exactly reduced integer phases
// (base, step, 2 step, 4 step ...) and one complex product per entry, so an entry is at most log2(RM) roundings away
// from exact.  (Squaring u = W^step instead of evaluating W^(2 step) saves two sincospif per column but pushed the
// cross-polarisation products of nchan 1024 to 1.02e-5 of the oracle; B2F_KGT_DIAG_SQUARE selects it.)
template <int RM, int LGM>
__device__ __forceinline__ void kgt_diag(float2 (&tw)[RM], uint32_t base, uint32_t step) {
    constexpr uint32_t MASK = (1u << LGM) - 1u, HALF = 1u << (LGM - 1);
    constexpr float kInvHalf = 1.0f / (float)(1u << (LGM - 1));
    // signed phase in [-M/2, M/2): |x| <= 1 keeps the float conversion exact up to M = 2^25
    auto w_of = [&](uint32_t ph) {
        const int p = (int)((ph + HALF) & MASK) - (int)HALF;
        float sn, cs;
        sincospif(-(float)p * kInvHalf, &sn, &cs);
        return make_float2(cs, sn);
    };
    tw[0] = w_of(base);
#ifdef B2F_KGT_DIAG_SQUARE
    float2 u = w_of(step);
#endif
#pragma unroll
    for (int h = 1; h < RM; h *= 2) {
#ifndef B2F_KGT_DIAG_SQUARE
        const float2 u = w_of(step * (uint32_t)h);
#endif
#pragma unroll
        for (int q = 0; q < h; ++q) tw[h + q] = cmul(tw[q], u);
#ifdef B2F_KGT_DIAG_SQUARE
        if (2 * h < RM) u = cmul(u, u);
#endif
    }
}

// innermost step of the column pass: FFT_RM, diagonal W_M^(k2 n1), IFFT_RM on the RM values of one segment, two columns
template <int LG>
__device__ __forceinline__ void kgt_col_inner(float2* sm, int n1_0, float2* colsum) {
    using G = KGT<LG>;
    constexpr int RM = G::RM, LGP = G::LGC - 1, CNT = (G::N >> G::LGI) << LGP, LGM = 2 * LG;
    constexpr int NW = G::kColThreads / 32, PER = CNT / NW;                 // warp-major, see kgt_col_pass16
    static_assert(CNT % (NW * 32) == 0, "whole warp iterations");
    const int t0 = (threadIdx.x >> 5) * PER;
#pragma unroll 1
    for (int tl = threadIdx.x & 31; tl < PER; tl += 32) {
        const int t = t0 + tl;
        const int c2 = (t & ((1 << LGP) - 1)) * 2, seg = t >> LGP;
        float2 a[RM], b[RM];
#pragma unroll
        for (int q = 0; q < RM; ++q) {
            const float4 v4 = *reinterpret_cast<const float4*>(&sm[kg_phys<false>(((seg * RM + q) << G::LGC) + c2)]);
            a[q] = make_float2(v4.x, v4.y);
            b[q] = make_float2(v4.z, v4.w);
        }
        fft_inreg<RM, false>(a);
        fft_inreg<RM, false>(b);
        if (seg == 0) {                                             // A[k2 = 0]
            colsum[c2] = a[0];
            colsum[c2 + 1] = b[0];
        }
        const uint32_t kseg = (uint32_t)kgt_rev16<G::NF>(seg), n1 = (uint32_t)(n1_0 + c2);
        {                                                           // k2 = kseg | q << 4 NF
            float2 tw[RM];
            kgt_diag<RM, LGM>(tw, kseg * n1, n1 << (4 * G::NF));
#pragma unroll
            for (int q = 0; q < RM; ++q) a[q] = cmul(a[q], tw[q]);
            kgt_diag<RM, LGM>(tw, kseg * (n1 + 1), (n1 + 1) << (4 * G::NF));
#pragma unroll
            for (int q = 0; q < RM; ++q) b[q] = cmul(b[q], tw[q]);
        }
        fft_inreg<RM, true>(a);
        fft_inreg<RM, true>(b);
#pragma unroll
        for (int q = 0; q < RM; ++q)
            *reinterpret_cast<float4*>(&sm[kg_phys<false>(((seg * RM + q) << G::LGC) + c2)]) = make_float4(a[q].x, a[q].y, b[q].x, b[q].y);
    }
}

template <int LG, int NBIT>
__global__ void __launch_bounds__(KGT<LG>::kColThreads, KGT<LG>::kColCtas) kgt_column_pass(const KGParams p) {
    using G = KGT<LG>;
    constexpr int N = G::N, C = G::C, NF = G::NF, NSTRIPS = N / C, LGS = LG - G::LGC;
    static_assert(NF == 2 || NF == 3, "two or three outer passes");
    extern __shared__ __align__(16) uint8_t kg_smem[];
    float2* tws = reinterpret_cast<float2*>(kg_smem);
    float2* lut = tws + (G::kTwShared ? N : 0);                           // [32]
    float2* data = lut + 32;                                              // [N][C], swizzled
    const int tid = threadIdx.x;
    if (G::kTwShared)
        for (int i = tid; i < N; i += G::kColThreads) tws[i] = p.tw_col[i];
    const float2* tw = G::kTwShared ? tws : p.tw_col;
    if (tid < 32) {
        const int c0 = tid & 3, c1 = (tid >> 2) & 3;
        const float m0 = (c0 == 0 || c0 == 3) ? kLevHi : kLevLo, m1 = (c1 == 0 || c1 == 3) ? kLevHi : kLevLo;
        lut[tid] = tid < 16 ? make_float2((c0 & 2) ? m0 : -m0, (c1 & 2) ? m1 : -m1) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const int64_t nwork = (p.gb_end - p.gb_begin) << LGS;
    for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int64_t lb = w >> LGS, gb = p.gb_begin + lb;
        const int strip = (int)(w & (NSTRIPS - 1));
        const int ifi = (int)(gb / p.nblk);
        const int64_t blk = gb - (int64_t)ifi * p.nblk;
        constexpr int BPS = NBIT == 8 ? 2 : 1;                            // stream bytes per time sample
        const int64_t off = (blk * p.step + (int64_t)strip * C) * BPS;           // step = M without overlap-save
        const uint8_t* src = p.compact + ifi * p.compact_stride + off;
        const KG8 s8{p.wmask + ifi * p.wmask_stride, off, p.in8_offset, NBIT == 8 && p.blkdirty[gb] != 0,
                     NBIT == 22 ? p.levels + ifi * p.levels_stride : nullptr};            // NBIT 22 = 2-bit input, JA98 levels
        float2* dst = p.inter + (lb << (2 * LG)) + strip * C;
        float2* colsum = p.colsum + gb * N + strip * C;
        // forward, outermost first
        kgt_col_pass16<LG, LG, false, NBIT == 8 ? 2 : (NBIT == 22 ? 3 : 1), 0>(data, tw, src, nullptr, lut, &s8);
        __syncthreads();
        kgt_col_pass16<LG, LG - 4, false, 0, 0, true>(data, tw + kg_tw_offset(LG, 1), nullptr, nullptr, lut);
        __syncwarp();
        if constexpr (NF == 3) {
            kgt_col_pass16<LG, LG - 8, false, 0, 0, true>(data, tw + kg_tw_offset(LG, 2), nullptr, nullptr, lut);
            __syncwarp();
        }
        kgt_col_inner<LG>(data, strip * C, colsum);
        __syncwarp();
        // inverse, innermost first
        if constexpr (NF == 3) {
            kgt_col_pass16<LG, LG - 8, true, 0, 0, true>(data, tw + kg_tw_offset(LG, 2), nullptr, nullptr, lut);
            __syncwarp();
        }
        kgt_col_pass16<LG, LG - 4, true, 0, 0, true>(data, tw + kg_tw_offset(LG, 1), nullptr, nullptr, lut);
        __syncthreads();
        kgt_col_pass16<LG, LG, true, 0, 1>(data, tw, nullptr, dst, lut);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------- rows
// One radix-16 row pass over the RB rows of a batch.  FIRST: reads the natural-order rows at `src` (the cp.async stage
// or, for the longest rows, global memory) instead of the swizzled buffer.
template <int LG, int LGN, bool FIRST>
__device__ __forceinline__ void kgt_row_pass16(float2* sm, const float2* tw, const float2* src) {
    using G = KGT<LG>;
    constexpr int LGM = LGN - 4, M1 = 1 << LGM, CNT = G::RB * (G::N >> 4);
#pragma unroll 1
    for (int t = threadIdx.x; t < CNT; t += G::kRowThreads) {
        const int lo = t & (M1 - 1), seg = t >> LGM;                  // seg runs over the rows of the batch as well
        const int base = (seg << LGN) + lo;
        const float2* twl = tw + lo;
        float2 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int idx = base + j * M1;
            v[j] = FIRST ? src[idx] : sm[kg_phys<true>(idx)];
        }
        fft_inreg<16, false>(v);
#pragma unroll
        for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], G::kTwShared ? twl[(q - 1) << LGM] : __ldg(twl + ((q - 1) << LGM)));
#pragma unroll
        for (int j = 0; j < 16; ++j) sm[kg_phys<true>(base + j * M1)] = v[j];
    }
}

template <int LG>
__device__ __forceinline__ void kgt_row_inner(float2* sm) {
    using G = KGT<LG>;
    constexpr int RM = G::RM, CNT = G::RB * (G::N >> G::LGI);
#pragma unroll 1
    for (int t = threadIdx.x; t < CNT; t += G::kRowThreads) {
        float2 v[RM];
#pragma unroll
        for (int q = 0; q < RM; ++q) v[q] = sm[kg_phys<true>((t << G::LGI) + q)];
        fft_inreg<RM, false>(v);
#pragma unroll
        for (int q = 0; q < RM; ++q) sm[kg_phys<true>((t << G::LGI) + q)] = v[q];
    }
}

// position of frequency k after the forward passes (compile-time geometry)
template <int LG>
__device__ __forceinline__ int kgt_pos_of_freq(int k) {
    using G = KGT<LG>;
    int seg = 0;
#pragma unroll
    for (int f = 0; f < G::NF; ++f) seg |= ((k >> (4 * f)) & 15) << (4 * (G::NF - 1 - f));
    return (seg << G::LGI) | (k >> (4 * G::NF));
}

// Row pass: a CTA holds RB rows of a block in shared memory, transforms them together, separates the polarisations
// between channel c and its mirror, detects (MODE: compile-time detection product) and integrates D rows.
template <int LG, int MODE>
__global__ void __launch_bounds__(KGT<LG>::kRowThreads, KGT<LG>::kRowCtas) kgt_row_pass(const KGParams p) {
    using G = KGT<LG>;
    constexpr int R = G::N, N = R / 2, RB = G::RB, CPT = G::CPT, NT = G::kRowThreads, NPROD = nprod_of_mode(MODE);
    constexpr int BATCH = RB * R;
    extern __shared__ __align__(16) uint8_t kg_smem[];
    float2* tws = reinterpret_cast<float2*>(kg_smem);
    float2* rows = tws + (G::kTwShared ? R : 0);                          // [RB][R], swizzled
    float2* stage = rows + BATCH;                                         // [2][RB][R], natural order (kRowStaged)
    const int tid = threadIdx.x, D = p.D;
    if (G::kTwShared)
        for (int i = tid; i < R; i += NT) tws[i] = p.tw_row[i];
    const float2* tw = G::kTwShared ? tws : p.tw_row;
    __syncthreads();
    const int U = max(D, RB);                                            // rows per work unit (whole output samples)
    const int lgU = 31 - __clz(U);
    const int lgUPB = LG - lgU;                                          // units per block = L / U
    const int64_t nunits = (p.gb_end - p.gb_begin) << lgUPB;
    int posA[CPT], posB[CPT];                                            // swizzled positions of channel c and its mirror
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
        const int c = tid + NT * k;
        posA[k] = kg_phys<true>(kgt_pos_of_freq<LG>(c));
        posB[k] = kg_phys<true>(kgt_pos_of_freq<LG>(R - 1 - c));
    }
    auto src_of = [&](int64_t g, int b0) {
        const int64_t lb = g >> lgUPB;
        return p.inter + (((lb << LG) + ((int)(g & ((1 << lgUPB) - 1)) << lgU) + b0) << LG);
    };
    auto prefetch = [&](int64_t g, int b0, int buf) {
        const float2* src = src_of(g, b0);
        float2* dst = stage + (size_t)buf * BATCH;
#pragma unroll
        for (int i = tid * 2; i < BATCH; i += 2 * NT) cp_async16(dst + i, src + i);
    };
    int it = 0;
    if (G::kRowStaged) {
        if (blockIdx.x < nunits) prefetch(blockIdx.x, 0, 0);
        cp_async_commit();
    }
    for (int64_t g = blockIdx.x; g < nunits; g += gridDim.x) {
        const int64_t lb = g >> lgUPB, gb = p.gb_begin + lb;
        const int r0 = (int)(g & ((1 << lgUPB) - 1)) << lgU;
        const int ifi = (int)(gb / p.nblk);
        const int64_t blk = gb - (int64_t)ifi * p.nblk;
        float acc[CPT][NPROD];
        float2 epsr[CPT];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
#pragma unroll
            for (int q = 0; q < NPROD; ++q) acc[k][q] = 0.f;
            epsr[k] = p.eps[gb * N + tid + NT * k];
        }
#pragma unroll 1
        for (int b0 = 0; b0 < U; b0 += RB, ++it) {
            const float2* src;
            if (G::kRowStaged) {
                int64_t gn = g;
                int bn = b0 + RB;
                if (bn >= U) { bn = 0; gn += gridDim.x; }
                if (gn < nunits) prefetch(gn, bn, (it + 1) & 1);
                cp_async_commit();
                cp_async_wait<1>();
                __syncthreads();
                src = stage + (size_t)(it & 1) * BATCH;
            } else {
                src = src_of(g, b0);
            }
            kgt_row_pass16<LG, LG, true>(rows, tw, src);
            __syncthreads();
            kgt_row_pass16<LG, LG - 4, false>(rows, tw + kg_tw_offset(LG, 1), nullptr);
            __syncthreads();
            if constexpr (G::NF == 3) {
                kgt_row_pass16<LG, LG - 8, false>(rows, tw + kg_tw_offset(LG, 2), nullptr);
                __syncthreads();
            }
            kgt_row_inner<LG>(rows);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                // the swizzle only looks at bits 0..10 of the element number: whole rows of 2048 or more keep it
                const float2* row = rows + (LG >= 11 ? (r << LG) : 0);
#pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    float2 a, b;
                    if (LG >= 11) {
                        a = row[posA[k]];
                        b = row[posB[k]];
                    } else {
                        const int c = tid + NT * k;
                        a = rows[kg_phys<true>((r << LG) + kgt_pos_of_freq<LG>(c))];
                        b = rows[kg_phys<true>((r << LG) + kgt_pos_of_freq<LG>(R - 1 - c))];
                    }
                    const float2 e = epsr[k];
                    const float2 bp = make_float2(b.x - e.x, -b.y - e.y);
                    detect_acc<MODE>(acc[k], make_float2(a.x + bp.x, a.y + bp.y), make_float2(a.x - bp.x, a.y - bp.y));
                }
                const int rr = r0 + b0 + r;                  // row of the block
                if (((rr + 1) & (D - 1)) == 0) {             // an output sample is complete
                    const int64_t t = p.row0 + (((blk << LG) + rr) >> (31 - __clz(D)));
                    float* dst = p.F + ifi * p.F_if_stride + t * (int64_t)(NPROD * N);
#pragma unroll
                    for (int k = 0; k < CPT; ++k) {
#pragma unroll
                        for (int q = 0; q < NPROD; ++q) {
                            dst[q * N + tid + NT * k] = acc[k][q];
                            acc[k][q] = 0.f;
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
    if (G::kRowStaged) cp_async_wait<0>();
}

}  // namespace b2f
