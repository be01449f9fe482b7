// Hand-written sm_100a kernels of the baseband -> filterbank path.  See DESIGN.md section 3
// for the algebra (column pass / row pass / block-constant correction) and section 4 for
// the HBM layout.  Replaces the arithmetic `digifil` performs for
// /root/reference/process_vdif.py:156-182 and `splice` for /root/reference/base2fil.sh:422.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2f.h"
#include "fft_inreg.cuh"

namespace b2f {

constexpr int kL = 512;                 // column FFT length (= digifil freq_res) of the fused path
constexpr int kStripCols = 16;          // columns per column-pass CTA
constexpr int kKAThreads = 256;
constexpr int kKBThreads = 256;
constexpr uint32_t kFillWord = 0x11223344u;
constexpr float kLevLo = 1.0f;          // standard VLBI optimal 2-bit reconstruction levels
constexpr float kLevHi = 3.3359f;

enum Counter {
    C_OK = 0, C_INVALID, C_FILLFRAMES, C_FILLWORDS, C_DROPPED, C_MISPLACED, C_BADHDR, C_MISSING,
    C_DIRTY, C_COUNT
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ================================================================== kernel 1: validate + de-frame
// One launch covers every IF (blockIdx.y).  Frames are staged whole into shared memory by
// TMA bulk copies (4-deep ring), the header is validated, every 32-bit payload word is
// compared with the recorder fill pattern, and the payload is written with 128-bit stores to
// the de-framed sample stream the channeliser reads.  One mask bit per payload word and one
// status byte per frame record what must decode to 0.0.
struct K0Params {
    const uint8_t* frames[B2F_MAX_IF];
    uint8_t* compact;  size_t compact_stride;
    uint8_t* wmask;    size_t wmask_stride;
    uint8_t* fstat;    size_t fstat_stride;
    unsigned long long* counters;
    int64_t nframes, nslots;
    int frame_bytes, header_bytes, payload_bytes, groups_per_slot;
    int in_nbit, time_mode, mask_faults, fps;
    uint32_t base_sec[B2F_MAX_IF], base_fnum[B2F_MAX_IF];
};

constexpr int kK0Threads = 128;
constexpr int kK0Stages = 4;

template <bool VEC>
__global__ void __launch_bounds__(kK0Threads) k0_validate_compact(const K0Params p) {
    extern __shared__ __align__(128) uint8_t k0_smem[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(k0_smem);
    const int stage_bytes = (p.frame_bytes + 127) & ~127;
    uint8_t* bufs = k0_smem + 128;
    const int ifi = blockIdx.y;
    const int tid = threadIdx.x;
    const uint8_t* src = p.frames[ifi];
    uint8_t* compact = p.compact + ifi * p.compact_stride;
    uint8_t* wmask = p.wmask + ifi * p.wmask_stride;
    uint8_t* fstat = p.fstat + ifi * p.fstat_stride;

    if (VEC) {
        if (tid == 0) {
            for (int s = 0; s < kK0Stages; ++s) mbar_init(&mbar[s], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            for (int s = 0; s < kK0Stages; ++s) {
                int64_t f = blockIdx.x + (int64_t)s * gridDim.x;
                if (f < p.nframes) {
                    mbar_expect_tx(&mbar[s], p.frame_bytes);
                    bulk_g2s(bufs + s * stage_bytes, src + f * p.frame_bytes, p.frame_bytes, &mbar[s]);
                }
            }
        }
    }

    unsigned long long c_ok = 0, c_inv = 0, c_fillf = 0, c_fillw = 0, c_drop = 0, c_mis = 0, c_bad = 0;
    int it = 0;
    for (int64_t f = blockIdx.x; f < p.nframes; f += gridDim.x, ++it) {
        const int stage = VEC ? (it % kK0Stages) : 0;
        uint8_t* buf = bufs + stage * stage_bytes;
        if (VEC) {
            mbar_wait(&mbar[stage], (it / kK0Stages) & 1);
        } else {
            const uint2* s8 = reinterpret_cast<const uint2*>(src + f * p.frame_bytes);
            for (int i = tid; i < p.frame_bytes / 8; i += kK0Threads) reinterpret_cast<uint2*>(buf)[i] = s8[i];
            __syncthreads();
        }
        const uint32_t* hw = reinterpret_cast<const uint32_t*>(buf);
        const uint32_t w0 = hw[0], w1 = hw[1], w2 = hw[2], w3 = hw[3];
        const bool invalid = (w0 >> 31) != 0;
        const bool bad = ((w2 & 0xFFFFFFu) * 8u != (uint32_t)p.frame_bytes) ||
                         ((int)((w3 >> 26) & 31u) + 1 != p.in_nbit) ||
                         ((int)((w0 >> 30) & 1u) != (p.header_bytes == 16 ? 1 : 0));
        const int64_t tslot = ((int64_t)(w0 & 0x3FFFFFFFu) - (int64_t)p.base_sec[ifi]) * p.fps +
                              ((int64_t)(w1 & 0xFFFFFFu) - (int64_t)p.base_fnum[ifi]);
        const int64_t slot = p.time_mode ? tslot : f;
        int any_fill = 0;
        unsigned nfill = 0;
        const bool in_range = (slot >= 0 && slot < p.nslots);
        if (in_range) {
            const bool dead = invalid || bad;
            const int ngroups = p.groups_per_slot;
            uint8_t* dst = compact + slot * (int64_t)p.payload_bytes;
            const uint8_t* pay = buf + p.header_bytes;
            for (int g = tid; g < ngroups; g += kK0Threads) {
                uint32_t m = 0;
                if (VEC) {
                    const uint4 a = *reinterpret_cast<const uint4*>(pay + 32 * g);
                    const uint4 b = *reinterpret_cast<const uint4*>(pay + 32 * g + 16);
                    m = (a.x == kFillWord) | ((a.y == kFillWord) << 1) | ((a.z == kFillWord) << 2) |
                        ((a.w == kFillWord) << 3) | ((b.x == kFillWord) << 4) | ((b.y == kFillWord) << 5) |
                        ((b.z == kFillWord) << 6) | ((b.w == kFillWord) << 7);
                    *reinterpret_cast<uint4*>(dst + 32 * g) = a;
                    *reinterpret_cast<uint4*>(dst + 32 * g + 16) = b;
                } else {
                    const int nw = min(8, (p.payload_bytes - 32 * g) / 4);
                    for (int k = 0; k < nw; ++k) {
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(pay + 32 * g + 4 * k);
                        m |= (w == kFillWord) << k;
                        *reinterpret_cast<uint32_t*>(dst + 32 * g + 4 * k) = w;
                    }
                }
                if (!dead) nfill += __popc(m);
                if (!p.mask_faults) m = 0;
                else if (dead) m = 0xFF;
                wmask[slot * (int64_t)ngroups + g] = (uint8_t)m;
                any_fill |= (m != 0);
            }
        }
        // barrier: everybody is done with buf; also reduces the fill flag
        any_fill = __syncthreads_or(any_fill);
        if (nfill) atomicAdd(&p.counters[C_FILLWORDS], (unsigned long long)nfill);
        if (tid == 0) {
            if (!in_range) {
                ++c_drop;
            } else {
                const bool dead = invalid || bad;
                fstat[slot] = dead ? 2 : (any_fill ? 3 : 1);
                if (bad) ++c_bad;
                else if (invalid) ++c_inv;
                else if (any_fill) { ++c_fillf; }
                else ++c_ok;
                if (tslot != f) ++c_mis;
            }
            if (VEC) {
                const int64_t fn = f + (int64_t)kK0Stages * gridDim.x;
                if (fn < p.nframes) {
                    fence_proxy_async();
                    mbar_expect_tx(&mbar[stage], p.frame_bytes);
                    bulk_g2s(buf, src + fn * p.frame_bytes, p.frame_bytes, &mbar[stage]);
                }
            }
        }
        (void)c_fillw;
    }
    if (tid == 0) {
        if (c_ok) atomicAdd(&p.counters[C_OK], c_ok);
        if (c_inv) atomicAdd(&p.counters[C_INVALID], c_inv);
        if (c_fillf) atomicAdd(&p.counters[C_FILLFRAMES], c_fillf);
        if (c_drop) atomicAdd(&p.counters[C_DROPPED], c_drop);
        if (c_mis) atomicAdd(&p.counters[C_MISPLACED], c_mis);
        if (c_bad) atomicAdd(&p.counters[C_BADHDR], c_bad);
    }
}

// slots no frame landed in are masked; blocks touched by any non-clean slot are flagged
struct K0bParams {
    uint8_t* wmask;    size_t wmask_stride;
    const uint8_t* fstat; size_t fstat_stride;
    uint8_t* blkdirty;                         // [nif][nblk]
    unsigned long long* counters;
    int64_t nslots; int nif, nblk, groups_per_slot, samples_per_frame; int64_t block_samples;
};
__global__ void k0b_finish_slots(const K0bParams p) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= p.nslots * p.nif) return;
    const int ifi = (int)(i / p.nslots);
    const int64_t slot = i % p.nslots;
    const uint8_t st = p.fstat[ifi * p.fstat_stride + slot];
    if (st == 1) return;
    if (st == 0) {
        uint8_t* m = p.wmask + ifi * p.wmask_stride + slot * (int64_t)p.groups_per_slot;
        for (int g = 0; g < p.groups_per_slot; ++g) m[g] = 0xFF;
        atomicAdd(&p.counters[C_MISSING], 1ull);
    }
    const int64_t s0 = slot * p.samples_per_frame;
    const int64_t b0 = s0 / p.block_samples, b1 = (s0 + p.samples_per_frame - 1) / p.block_samples;
    for (int64_t b = b0; b <= b1 && b < p.nblk; ++b) p.blkdirty[ifi * (int64_t)p.nblk + b] = 1;
}

// ================================================================== stand-alone decode
// compact payload (+ word mask) -> planar float samples out[2][nsamp].  HBM-bound expansion
// (2 bit -> 32 bit); the production path never runs it, the column pass decodes on the fly.
template <int NBIT>
__global__ void k_decode(const uint8_t* __restrict__ compact, const uint8_t* __restrict__ wmask,
                         int payload_bytes, int groups_per_slot, int64_t nwords, float* __restrict__ out,
                         int64_t nsamp) {
    const int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    const uint32_t v = reinterpret_cast<const uint32_t*>(compact)[w];
    const int words_per_slot = payload_bytes / 4;
    const int64_t slot = w / words_per_slot;
    const int ws = (int)(w % words_per_slot);
    const bool bad = (wmask[slot * groups_per_slot + (ws >> 3)] >> (ws & 7)) & 1;
    if (NBIT == 2) {
        float a[8], b[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t c0 = (v >> (4 * k)) & 3u, c1 = (v >> (4 * k + 2)) & 3u;
            const float m0 = (c0 == 0 || c0 == 3) ? kLevHi : kLevLo;
            const float m1 = (c1 == 0 || c1 == 3) ? kLevHi : kLevLo;
            a[k] = bad ? 0.f : ((c0 & 2) ? m0 : -m0);
            b[k] = bad ? 0.f : ((c1 & 2) ? m1 : -m1);
        }
        float4* o0 = reinterpret_cast<float4*>(out + w * 8);
        float4* o1 = reinterpret_cast<float4*>(out + nsamp + w * 8);
        o0[0] = make_float4(a[0], a[1], a[2], a[3]);
        o0[1] = make_float4(a[4], a[5], a[6], a[7]);
        o1[0] = make_float4(b[0], b[1], b[2], b[3]);
        o1[1] = make_float4(b[4], b[5], b[6], b[7]);
    } else {
        float2* o0 = reinterpret_cast<float2*>(out + w * 2);
        float2* o1 = reinterpret_cast<float2*>(out + nsamp + w * 2);
        const float x0 = bad ? 0.f : (float)(v & 255u) - 127.5f, y0 = bad ? 0.f : (float)((v >> 8) & 255u) - 127.5f;
        const float x1 = bad ? 0.f : (float)((v >> 16) & 255u) - 127.5f, y1 = bad ? 0.f : (float)(v >> 24) - 127.5f;
        o0[0] = make_float2(x0, x1);
        o1[0] = make_float2(y0, y1);
    }
}

// ================================================================== kernel 3a: column pass
// A block of M = R*512 dual-pol samples z[n] = xP[n] + i xQ[n] is viewed as a 512 x R matrix
// (n = n1 + R n2).  One CTA owns a strip of 16 columns: lane = column, so every shared-memory
// access is [index][lane] (conflict-free, no padding) and every twiddle except the two
// column-dependent tables is uniform across the 16 lanes.  Per column:
//   decode -> FFT_512 (16 x 32) -> * W_M^(k2 n1) -> IFFT_512 (32 x 16)  ->  B[m][n1]
// with only two shared-memory exchanges: the 32-point forward and inverse FFTs around the
// diagonal multiply act on the same 32 values, so they run back to back in registers.
struct KAParams {
    const uint8_t* compact;  size_t compact_stride;
    const uint8_t* wmask;    size_t wmask_stride;
    const uint8_t* blkdirty;
    float2* inter;           // [nif*nblk][512][R]
    float2* colsum;          // [nif*nblk][R]
    const float2* tab_g;     // [16][R]  W_M^(q n1)
    const float2* tab_h;     // [32][R]  W_M^(16 p n1)
    const float2* tab_w;     // [32][16] W_512^(l q)
    int R, nstrips, nblk, nif, payload_bytes, groups_per_slot;
};

template <int NBIT>
struct KASmem {
    static constexpr int kPiece = NBIT == 2 ? 8 : 32;          // raw bytes per (row, strip)
    static constexpr int kRawBytes = kL * kPiece;
    static constexpr size_t kBytes = (size_t)(kL * 16 + 512 + 512 + 512 + 256 + 16) * sizeof(float2) + 2 * kRawBytes;
};

template <int NBIT>
__global__ void __launch_bounds__(kKAThreads, 2) ka_column_pass(const KAParams p) {
    using S = KASmem<NBIT>;
    extern __shared__ __align__(16) uint8_t ka_smem[];
    float2* data = reinterpret_cast<float2*>(ka_smem);      // [512][16]
    float2* s_w = data + kL * 16;                            // [32][16]  W_512^(l q)
    float2* s_wT = s_w + 512;                                // [16][32]
    float2* s_h = s_wT + 512;                                // [32][16 lanes]
    float2* s_g = s_h + 512;                                 // [16][16 lanes]
    float2* s_lut = s_g + 256;                               // [16] nibble -> (pol0, pol1)
    uint8_t* s_raw = reinterpret_cast<uint8_t*>(s_lut + 16); // [2][512][kPiece]

    const int tid = threadIdx.x;
    const int lane16 = tid & 15;
    const int item = tid >> 4;
    const int R = p.R;
    const int strip = blockIdx.x % p.nstrips;
    const int n1 = strip * kStripCols + lane16;

    for (int i = tid; i < 512; i += kKAThreads) {
        const float2 w = p.tab_w[i];
        s_w[i] = w;
        s_wT[(i & 15) * 32 + (i >> 4)] = w;
        s_h[i] = p.tab_h[(i >> 4) * R + strip * kStripCols + (i & 15)];
    }
    s_g[tid] = p.tab_g[(tid >> 4) * R + strip * kStripCols + (tid & 15)];
    if (tid < 16) {
        const int c0 = tid & 3, c1 = tid >> 2;
        const float m0 = (c0 == 0 || c0 == 3) ? kLevHi : kLevLo;
        const float m1 = (c1 == 0 || c1 == 3) ? kLevHi : kLevLo;
        s_lut[tid] = make_float2((c0 & 2) ? m0 : -m0, (c1 & 2) ? m1 : -m1);
    }

    const int64_t nbt = (int64_t)p.nif * p.nblk;
    const int64_t first = blockIdx.x / p.nstrips;
    const int64_t step = gridDim.x / p.nstrips;
    const int bytes_per_samp4 = NBIT == 2 ? 1 : 4;            // bytes per 2 time samples (both pols)
    const int64_t row_bytes = (int64_t)R * bytes_per_samp4 / 2;
    const int64_t blk_bytes = row_bytes * kL;

    auto issue_raw = [&](int64_t gb, int buf) {
        const int ifi = (int)(gb / p.nblk);
        const int64_t blk = gb % p.nblk;
        const uint8_t* src = p.compact + ifi * p.compact_stride + blk * blk_bytes + (int64_t)strip * S::kPiece;
        uint8_t* dst = s_raw + buf * S::kRawBytes;
        if (NBIT == 2) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int row = tid + k * kKAThreads;
                cp_async8(dst + row * 8, src + row * row_bytes);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int j = tid + k * kKAThreads;             // 1024 pieces of 16 B
                const int row = j >> 1, h = j & 1;
                cp_async16(dst + row * 32 + h * 16, src + row * row_bytes + h * 16);
            }
        }
    };

    if (first < nbt) issue_raw(first, 0);
    cp_async_commit();

    int it = 0;
    for (int64_t gb = first; gb < nbt; gb += step, ++it) {
        const int buf = it & 1;
        if (gb + step < nbt) issue_raw(gb + step, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        const bool dirty = p.blkdirty[gb] != 0;
        const uint8_t* raw = s_raw + buf * S::kRawBytes;
        const int ifi = (int)(gb / p.nblk);
        const int64_t blk = gb % p.nblk;

        // ---- P1: decode + 16-point FFT over r (n2 = 32 r + l), twiddle W_512^(l q)
#pragma unroll 1
        for (int rnd = 0; rnd < 2; ++rnd) {
            const int l = item + 16 * rnd;
            float2 v[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int row = 32 * r + l;
                if (NBIT == 2) {
                    const uint32_t b = raw[row * 8 + (lane16 >> 1)];
                    v[r] = s_lut[(b >> ((lane16 & 1) * 4)) & 15u];
                } else {
                    const uint32_t b = *reinterpret_cast<const uint16_t*>(raw + row * 32 + lane16 * 2);
                    v[r] = make_float2((float)(b & 255u) - 127.5f, (float)(b >> 8) - 127.5f);
                }
            }
            if (dirty) {
                const uint8_t* wm = p.wmask + ifi * p.wmask_stride;
#pragma unroll 1
                for (int r = 0; r < 16; ++r) {
                    const int row = 32 * r + l;
                    const int64_t off = blk * blk_bytes + row * row_bytes + (int64_t)strip * S::kPiece +
                                        (NBIT == 2 ? (lane16 >> 1) : lane16 * 2);
                    const int64_t slot = off / p.payload_bytes;
                    const int w = (int)(off % p.payload_bytes) >> 2;
                    if ((wm[slot * p.groups_per_slot + (w >> 3)] >> (w & 7)) & 1) {
                        // cannot index v[] dynamically without spilling: select per r
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr)
                            if (rr == r) v[rr] = make_float2(0.f, 0.f);
                    }
                }
            }
            fft_inreg<16, false>(v);
            const float4* tw = reinterpret_cast<const float4*>(s_w + l * 16);
#pragma unroll
            for (int q = 0; q < 16; q += 2) {
                const float4 t = tw[q >> 1];
                if (q) v[q] = cmul(v[q], make_float2(t.x, t.y));
                v[q + 1] = cmul(v[q + 1], make_float2(t.z, t.w));
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) data[(q * 32 + l) * 16 + lane16] = v[q];
        }
        __syncthreads();

        // ---- Mid: FFT_32 over l -> p ; * W_M^(16 p n1) ; IFFT_32 over p -> m1 ; * conj W_512^(q m1)
        {
            const int q = item;
            float2 u[32];
#pragma unroll
            for (int l = 0; l < 32; ++l) u[l] = data[(q * 32 + l) * 16 + lane16];
            fft_inreg<32, false>(u);
            if (q == 0) p.colsum[gb * R + n1] = u[0];          // A[k2 = 0]: column sum
#pragma unroll
            for (int pp = 1; pp < 32; ++pp) u[pp] = cmul(u[pp], s_h[pp * 16 + lane16]);
            fft_inreg<32, true>(u);
            const float4* tw = reinterpret_cast<const float4*>(s_wT + q * 32);
#pragma unroll
            for (int m1 = 0; m1 < 32; m1 += 2) {
                const float4 t = tw[m1 >> 1];
                if (m1) u[m1] = cmul_conj(u[m1], make_float2(t.x, t.y));
                u[m1 + 1] = cmul_conj(u[m1 + 1], make_float2(t.z, t.w));
            }
#pragma unroll
            for (int m1 = 0; m1 < 32; ++m1) data[(q * 32 + m1) * 16 + lane16] = u[m1];
        }
        __syncthreads();

        // ---- P3: * W_M^(q n1) ; IFFT_16 over q -> m2 ; m = m1 + 32 m2
        float2* dst = p.inter + (gb * (int64_t)kL) * R + n1;
#pragma unroll 1
        for (int rnd = 0; rnd < 2; ++rnd) {
            const int m1 = item + 16 * rnd;
            float2 y[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) y[q] = data[(q * 32 + m1) * 16 + lane16];
#pragma unroll
            for (int q = 1; q < 16; ++q) y[q] = cmul(y[q], s_g[q * 16 + lane16]);
            fft_inreg<16, true>(y);
#pragma unroll
            for (int m2 = 0; m2 < 16; ++m2) dst[(int64_t)(m1 + 32 * m2) * R] = y[m2];
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}

// ================================================================== kernel 3b: eps
// eps_c = conj(G[R-1-c] - G[(R-c) mod R]),  G = FFT_R(column sums): the one block-constant
// term that separating the two real polarisations after the row pass needs.
__global__ void ke_eps(const float2* __restrict__ colsum, float2* __restrict__ eps, int R) {
    extern __shared__ float2 ke_s[];
    const int t = threadIdx.x;                      // R/2 threads
    const int lg = 31 - __clz(R);
    const float2* src = colsum + (int64_t)blockIdx.x * R;
    for (int i = t; i < R; i += blockDim.x) ke_s[__brev(i) >> (32 - lg)] = src[i];
    __syncthreads();
    for (int s = 1; s <= lg; ++s) {
        const int half = 1 << (s - 1);
        const int j = t & (half - 1);
        const int i0 = ((t >> (s - 1)) << s) + j, i1 = i0 + half;
        float sn, cs;
        sincospif(-(float)j / (float)half, &sn, &cs);
        const float2 a = ke_s[i0], b = ke_s[i1];
        const float2 wb = make_float2(b.x * cs - b.y * sn, b.x * sn + b.y * cs);
        ke_s[i0] = make_float2(a.x + wb.x, a.y + wb.y);
        ke_s[i1] = make_float2(a.x - wb.x, a.y - wb.y);
        __syncthreads();
    }
    const float2 g0 = ke_s[R - 1 - t], g1 = ke_s[(R - t) & (R - 1)];
    eps[(int64_t)blockIdx.x * (R / 2) + t] = make_float2(g0.x - g1.x, -(g0.y - g1.y));
}

// ================================================================== kernel 3c+4: row pass
// FFT_R across the R columns of every row m of the column-pass output, separation of the two
// polarisations (channel c with its mirror R-1-c, minus eps_c), detection and time
// integration over D rows.  TR lanes cooperate on one row (PT points each, R = TR*PT):
// radix-PT in registers, one transpose through padded shared memory, radix-TR in registers,
// mirror exchange by shuffle.
struct KBParams {
    const float2* inter;
    const float2* eps;          // [nif*nblk][R/2]
    const float2* tab_r;        // [PT][TR]  W_R^(s q)
    float* F;                   // [nif][cap_rows][nprod][R/2]
    int64_t F_if_stride;        // floats between IFs
    int64_t row0;               // first output row of this push inside F
    int nblk, nif, D, mode;
};

template <int NPROD>
__device__ __forceinline__ void detect_acc(float (&acc)[NPROD], float2 P, float2 Q, int mode) {
    const float pp = 0.25f * (P.x * P.x + P.y * P.y);
    const float qq = 0.25f * (Q.x * Q.x + Q.y * Q.y);
    if (NPROD == 1) {
        if (mode == B2F_POL_I) acc[0] += pp + qq;
        else if (mode == B2F_POL_P0) acc[0] += pp;
        else if (mode == B2F_POL_P1) acc[0] += qq;
        else { const float s = pp + qq; acc[0] = fmaf(s, s, acc[0]); }
    } else if (NPROD == 2) {
        acc[0] += pp;
        acc[1 % NPROD] += qq;
    } else {
        // yP yQ* = (i/4) P conj(Q):  Re = -Im(P conj Q)/4,  Im = Re(P conj Q)/4
        const float xr = P.x * Q.x + P.y * Q.y, xi = P.y * Q.x - P.x * Q.y;
        const float re = -0.25f * xi, im = 0.25f * xr;
        if (mode == B2F_POL_COHERENCE) {
            acc[0] += pp; acc[1 % NPROD] += qq; acc[2 % NPROD] += re; acc[3 % NPROD] += im;
        } else {
            acc[0] += pp + qq; acc[1 % NPROD] += 2.f * re; acc[2 % NPROD] += 2.f * im; acc[3 % NPROD] += pp - qq;
        }
    }
}

template <int TR, int PT, int NPROD>
struct KBSmem {
    static constexpr int NRS = kKBThreads / TR;
    static constexpr int N = TR * PT / 2;
    static constexpr size_t kXch = (size_t)NRS * TR * (PT + 1) * sizeof(float2);
    static constexpr size_t kRed = (size_t)NRS * NPROD * N * sizeof(float);
    static constexpr size_t kTw = (size_t)PT * TR * sizeof(float2);
    static constexpr size_t kBytes = kXch + kRed + kTw;
};

template <int TR, int PT, int NPROD>
__global__ void __launch_bounds__(kKBThreads, 2) kb_row_pass(const KBParams p) {
    constexpr int R = TR * PT, N = R / 2, NRS = kKBThreads / TR, QPT = PT / TR, HP = TR / 2;
    using S = KBSmem<TR, PT, NPROD>;
    extern __shared__ __align__(16) uint8_t kb_smem[];
    float2* xch = reinterpret_cast<float2*>(kb_smem);
    float* red = reinterpret_cast<float*>(kb_smem + S::kXch);
    float2* s_tw = reinterpret_cast<float2*>(kb_smem + S::kXch + S::kRed);

    const int tid = threadIdx.x;
    const int s = tid % TR, rs = tid / TR;
    for (int i = tid; i < PT * TR; i += kKBThreads) s_tw[i] = p.tab_r[i];
    __syncthreads();

    const int D = p.D;
    const int G = D > NRS ? D : NRS;                // rows per group
    const int passes = G / NRS;
    const int nout = G / D;
    const int rs_per_out = NRS / nout;
    const int groups_per_blk = kL / G;
    const int64_t ngroups = (int64_t)p.nif * p.nblk * groups_per_blk;
    float2* myx = xch + (size_t)rs * TR * (PT + 1);

    for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int64_t gb = grp / groups_per_blk;
        const int g0 = (int)(grp % groups_per_blk) * G;
        float2 e[QPT][HP];
#pragma unroll
        for (int j = 0; j < QPT; ++j)
#pragma unroll
            for (int pp = 0; pp < HP; ++pp) e[j][pp] = p.eps[gb * N + (s + TR * j) + PT * pp];
        float acc[QPT][HP][NPROD];
#pragma unroll
        for (int j = 0; j < QPT; ++j)
#pragma unroll
            for (int pp = 0; pp < HP; ++pp)
#pragma unroll
                for (int k = 0; k < NPROD; ++k) acc[j][pp][k] = 0.f;

#pragma unroll 1
        for (int pass = 0; pass < passes; ++pass) {
            const int row = g0 + pass * NRS + rs;
            const float2* src = p.inter + (gb * (int64_t)kL + row) * R;
            float2 v[PT];
#pragma unroll
            for (int a = 0; a < PT; ++a) v[a] = src[s + TR * a];
            fft_inreg<PT, false>(v);
#pragma unroll
            for (int q = 1; q < PT; ++q) v[q] = cmul(v[q], s_tw[q * TR + s]);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < PT; ++q) myx[s * (PT + 1) + q] = v[q];
            __syncwarp();
            float2 z[QPT][TR];
#pragma unroll
            for (int j = 0; j < QPT; ++j)
#pragma unroll
                for (int t = 0; t < TR; ++t) z[j][t] = myx[t * (PT + 1) + s + TR * j];
#pragma unroll
            for (int j = 0; j < QPT; ++j) fft_inreg<TR, false>(z[j]);
            // z[j][pp] = Z_c at c = (s + TR j) + PT pp.  Mirror R-1-c lives in lane s^(TR-1),
            // register [QPT-1-j][TR-1-pp].
#pragma unroll
            for (int j = 0; j < QPT; ++j)
#pragma unroll
                for (int pp = 0; pp < HP; ++pp) {
                    const float2 a = z[j][pp];
                    const float2 bs = z[QPT - 1 - j][TR - 1 - pp];
                    const float bx = __shfl_xor_sync(0xffffffffu, bs.x, TR - 1);
                    const float by = __shfl_xor_sync(0xffffffffu, bs.y, TR - 1);
                    const float2 bp = make_float2(bx - e[j][pp].x, -by - e[j][pp].y);
                    detect_acc<NPROD>(acc[j][pp], make_float2(a.x + bp.x, a.y + bp.y),
                                      make_float2(a.x - bp.x, a.y - bp.y), p.mode);
                }
        }
        // reduce the row slots that integrate into the same output sample
#pragma unroll
        for (int j = 0; j < QPT; ++j)
#pragma unroll
            for (int pp = 0; pp < HP; ++pp)
#pragma unroll
                for (int k = 0; k < NPROD; ++k)
                    red[(rs * NPROD + k) * N + (s + TR * j) + PT * pp] = acc[j][pp][k];
        __syncthreads();
        const int ifi = (int)(gb / p.nblk);
        const int64_t blk = gb % p.nblk;
        const int64_t t0 = p.row0 + (blk * kL + g0) / D;
        for (int idx = tid; idx < nout * NPROD * N; idx += kKBThreads) {
            const int o = idx / (NPROD * N), rem = idx % (NPROD * N);
            float sum = 0.f;
            for (int r = 0; r < rs_per_out; ++r) sum += red[(o * rs_per_out + r) * NPROD * N + rem];
            p.F[ifi * p.F_if_stride + (t0 + o) * (int64_t)(NPROD * N) + rem] = sum;
        }
        __syncthreads();
    }
}

// ================================================================== kernel 5a: statistics
// mean / sigma per (IF, product, channel) over the first rescale interval, fp64 accumulators,
// deterministic two-level reduction.
__global__ void ks_partial(const float* __restrict__ F, int64_t F_if_stride, int64_t rows, int ncol,
                           double2* __restrict__ partial) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const int ifi = blockIdx.y, split = blockIdx.z, nsplit = gridDim.z;
    const int64_t r0 = rows * split / nsplit, r1 = rows * (split + 1) / nsplit;
    const float* f = F + ifi * F_if_stride + col;
    double s = 0.0, ss = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        const double x = f[r * ncol];
        s += x;
        ss += x * x;
    }
    partial[((int64_t)ifi * nsplit + split) * ncol + col] = make_double2(s, ss);
}
__global__ void ks_final(const double2* __restrict__ partial, int nsplit, int64_t rows, int ncol,
                         float* __restrict__ mean, float* __restrict__ scale) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncol) return;
    const int ifi = blockIdx.y;
    double s = 0.0, ss = 0.0;
    for (int k = 0; k < nsplit; ++k) {
        const double2 v = partial[((int64_t)ifi * nsplit + k) * ncol + col];
        s += v.x;
        ss += v.y;
    }
    const double m = rows > 0 ? s / (double)rows : 0.0;
    const double var = rows > 0 ? ss / (double)rows - m * m : 0.0;
    mean[ifi * ncol + col] = (float)m;
    scale[ifi * ncol + col] = var > 0.0 ? (float)(1.0 / sqrt(var)) : 1.0f;
}

// ================================================================== kernel 5b: requantise + flip + splice
// One thread = 4 consecutive output channels of one output row.  The band flip of USB
// subbands and the splice order are index arithmetic on the load side.
struct KQParams {
    const float* F; int64_t F_if_stride;
    const float* mean; const float* scale;        // [nif][nprod*nchan]
    void* out;
    int64_t rows; int64_t out_row_elems;
    int nif, nprod, nchan, out_nbit, pol_major;
    int if_order[B2F_MAX_IF];
    int flip[B2F_MAX_IF];                          // 1: channel reversal (USB)
};

__device__ __forceinline__ float quant(float y, float dscale, float dmean, float dmax) {
    return fminf(fmaxf(floorf(fmaf(y, dscale, dmean + 0.5f)), 0.f), dmax);
}

__global__ void kq_quantise(const KQParams p) {
    const int64_t quads_per_row = p.out_row_elems / 4;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= p.rows * quads_per_row) return;
    const int64_t row = i / quads_per_row;
    const int j = (int)(i % quads_per_row) * 4;
    int tile, prod, k;
    if (p.pol_major) {
        prod = j / (p.nif * p.nchan);
        tile = (j / p.nchan) % p.nif;
        k = j % p.nchan;
    } else {
        tile = j / (p.nprod * p.nchan);
        prod = (j / p.nchan) % p.nprod;
        k = j % p.nchan;
    }
    const int ifi = p.if_order[tile];
    const int ncol = p.nprod * p.nchan;
    const float* f = p.F + ifi * p.F_if_stride + row * (int64_t)ncol + prod * p.nchan;
    const float* mu = p.mean + ifi * ncol + prod * p.nchan;
    const float* sc = p.scale + ifi * ncol + prod * p.nchan;
    float y[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int c = p.flip[ifi] ? (p.nchan - 1 - (k + a)) : (k + a);
        y[a] = (f[c] - mu[c]) * sc[c];
    }
    const int64_t o = row * p.out_row_elems + j;
    if (p.out_nbit == 8) {
        uint32_t w = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) w |= (uint32_t)quant(y[a], 127.5f / 6.0f, 127.5f, 255.f) << (8 * a);
        reinterpret_cast<uint32_t*>(p.out)[o / 4] = w;
    } else if (p.out_nbit == 16) {
        ushort4 w;
        w.x = (unsigned short)quant(y[0], 32768.0f / 6.0f, 32768.0f, 65535.f);
        w.y = (unsigned short)quant(y[1], 32768.0f / 6.0f, 32768.0f, 65535.f);
        w.z = (unsigned short)quant(y[2], 32768.0f / 6.0f, 32768.0f, 65535.f);
        w.w = (unsigned short)quant(y[3], 32768.0f / 6.0f, 32768.0f, 65535.f);
        reinterpret_cast<ushort4*>(p.out)[o / 4] = w;
    } else if (p.out_nbit == 2) {
        uint32_t w = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) w |= (uint32_t)quant(y[a], 1.0f, 1.5f, 3.f) << (2 * a);
        reinterpret_cast<uint8_t*>(p.out)[o / 4] = (uint8_t)w;
    } else {
        reinterpret_cast<float4*>(p.out)[o / 4] = make_float4(y[0], y[1], y[2], y[3]);
    }
}

}  // namespace b2f
