// Round-2 channeliser: warp-autonomous column pass fused with the row pass through an L2-resident ring.
//
// What it replaces: ka_column_pass + ke_eps + kb_row_pass of b2f_kernels.cuh for 2-bit input without
// dedispersion (the arithmetic of `digifil -F<nchan>:512 -d<n> -t<D>`, /root/reference/process_vdif.py:156-176).
// Same algebra, different mapping onto the machine:
//
//   * one WARP owns two neighbouring columns of a block: lane = 2 * item + col, item = 0..15.  The two
//     shared-memory exchanges of the 512-point FFT / diagonal / inverse FFT happen inside the warp
//     (__syncwarp only), so no block barrier exists anywhere and the 16 warps of an SM drift into
//     different phases: one warp's exchanges hide under another's butterflies.
//   * the exchange buffer is a padded 16 x 17 matrix of float4 per column: column access (P1 write,
//     P3 read) and row access (middle section) are both bank-conflict free with base + immediate
//     addresses.  The column-dependent twiddle table is 2 columns wide, i.e. a broadcast.
//   * input is the block-transposed index stream k0t_transpose writes: every lane gets its 32 samples
//     with two 128-bit loads, prefetched one work item ahead (no shared-memory staging, no cp.async).
//   * the column result goes to a ring of 1 MiB block slots that stays in L2 ([pair][row][2] float2, so a
//     warp's 8 KiB are contiguous); the warps of the same "lane" (R/2 warps = one block) then run that
//     block's rows one round later, reading 16-byte pieces with cp.async.  Per-lane arrival counters
//     (fence + relaxed atomic / acquire load) order the two; the [blk][512][R] intermediate never
//     reaches HBM.
//   * eps (the block-constant term) is computed by one rotating warp per lane and round.
//   * every warp integrates only its own 1024/R rows; if tscrunch spans more, kt_sum_partials adds the
//     partial rows in a fixed order (deterministic; no float atomics).
//
// phase 1 / phase 2 run the two halves as separate launches over a full-size intermediate (same device
// functions): kept as a debugging and profiling aid (B2F_PATH=split).
#pragma once
#include "b2f_kernels.cuh"

namespace b2f {

constexpr int kFWarps = 16;                 // one 512-thread CTA per SM: 16 autonomous warps, 128 registers each
constexpr int kFThreads = kFWarps * 32;
constexpr unsigned kFSpinLimit = 1u << 22;  // polls (with nanosleep) before a waiting warp gives up: ~1 s

enum FSync { FS_COL = 0, FS_ROW = 1, FS_EPS = 2, FS_STRIDE = 32 };   // counters of one lane, 128 bytes apart

template <int R> struct FShape;
template <> struct FShape<16>  { static constexpr int TR = 4,  PT = 4;  };
template <> struct FShape<32>  { static constexpr int TR = 4,  PT = 8;  };
template <> struct FShape<64>  { static constexpr int TR = 8,  PT = 8;  };
template <> struct FShape<128> { static constexpr int TR = 8,  PT = 16; };
#ifndef B2F_F256_TR
#define B2F_F256_TR 16
#endif
template <> struct FShape<256> { static constexpr int TR = B2F_F256_TR, PT = 256 / B2F_F256_TR; };
template <> struct FShape<512> { static constexpr int TR = 16, PT = 32; };

template <int R>
struct FGeo {
    static constexpr int TR = FShape<R>::TR, PT = FShape<R>::PT;
    static constexpr int N = R / 2, NPAIR = R / 2;
    static constexpr int RPW = 1024 / R;                       // rows of a block one warp transforms
    static constexpr int RW = 32 / TR;                         // rows per warp pass
    static constexpr int NPASS = RPW / RW;
    static constexpr int kPitch = R * 8 + (TR < 16 ? TR * 8 : 0);     // bytes between rows of the row tile
    static constexpr int kXch = RW * TR * (PT + 1) * 8;               // transposition buffer of one pass
    static constexpr int kRowBuf = RPW * kPitch + (kXch > RW * kPitch ? kXch - RW * kPitch : 0);
    static constexpr int kColBuf = 16 * 34 * 16;                      // 16 x 17 float4 per column, two columns
    static constexpr int kWarpBuf = ((kRowBuf > kColBuf ? kRowBuf : kColBuf) + 127) / 128 * 128;
    static constexpr int kOffW4 = 0;                                  // [8][32] float4  W_512^(l q), q pairs
    static constexpr int kOffLut = kOffW4 + 8 * 32 * 16;              // 17 x float2 (+ pad)
    static constexpr int kOffTw = kOffLut + 256;                      // [PT][TR] float2 W_R^(s q)
    static constexpr int kOffH = kOffTw + R * 8;                      // per warp [16][2] float4 (h^p, h^(p+16))
    static constexpr int kOffBuf = (kOffH + kFWarps * 512 + 127) / 128 * 128;
    static constexpr size_t kBytes = (size_t)kOffBuf + (size_t)kFWarps * kWarpBuf;
};

struct FParams {
    const uint8_t* tstream;        // [if][blk][pair][half][lane][16] index bytes (k0t_transpose)
    size_t tstream_if_stride;
    float2* ring;                  // fused: [(lag + 1) * lanes][M]; split: [blocks of the launch][M]; slot layout [32 row tiles][R/2 column pairs][16 rows][2 columns]
    float2* colsum;                // [nif*nblk][R]
    float2* eps;                   // [nif*nblk][R/2]
    const float2 *tab_h, *tab_w, *tab_beta, *tab_r;
    float* out;                    // F (Dp == D) or the partial-row buffer
    int64_t out_if_stride;         // floats between IFs
    int64_t out_row0;              // first row of this push inside `out`, in units of Dp input rows
    int Dp;                        // rows integrated per output row here: min(tscrunch, 1024 / R)
    int nblk, nif;
    int64_t gb_begin, gb_end;
    unsigned* sync;                // [lanes][FS_STRIDE]
    unsigned* abort_flag;
    int phase;                     // 0 fused, 1 column halves only, 2 row halves only
    int nslot;                     // ring slots per lane (fused): lag + 1
    int lag;                       // rounds between the column half and the row half of a block (fused): 1 or 2
    const float4* levels;          // JA98 decode: [nif*nblk][R windows] (lo0, hi0, lo1, hi1); NULL = static levels
    unsigned long long* prof;      // optional [warps][8] cycle counters (B2F_FUSED_PROF=1): where a warp's time goes
};

// ------------------------------------------------------------------ inter-warp ordering through L2
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// The whole warp calls; true when *ctr >= target was observed, false on abort / timeout.  The poll is a relaxed
// gpu-scope load: everything a waiter reads afterwards is read with .cg loads / cp.async.cg, i.e. from L2, the point
// of coherence, and is issued after the poll returned (control dependency), while the producer ordered its stores
// before the counter update with a gpu-scope fence.  An acquire load here costs an L1 invalidation (CCTL.IVALL) per
// poll: 12 % of the fused kernel's stall samples in the first version.
__device__ __forceinline__ bool warp_wait_ge(const unsigned* ctr, unsigned target, unsigned* abort_flag, int lane) {
    int ok = 1;
    if (lane == 0) {
        unsigned spins = 0;
        while (ld_relaxed_u32(ctr) < target) {
            if (++spins > kFSpinLimit || ((spins & 63u) == 0 && ld_relaxed_u32(abort_flag) != 0)) {
                atomicExch(abort_flag, 1u);
                ok = 0;
                break;
            }
            __nanosleep(spins < 64 ? 32 : 256);
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    return ok != 0;
}
// all lanes have finished their global stores; one release-ordered increment for the warp.  The fence waits until
// the warp's earlier stores are visible, so callers place this well after the stores (end of the round).
__device__ __forceinline__ void warp_arrive(unsigned* ctr, int lane) {
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
    }
}
// the warp's loads have completed (their data is in shared memory): nothing to publish, a relaxed increment is enough
__device__ __forceinline__ void warp_arrive_relaxed(unsigned* ctr, int lane) {
    __syncwarp();
    if (lane == 0) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ float2 ldcg_f2(const float2* p) {
    float2 v;
    asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {       // read-once input: do not keep it in L1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ------------------------------------------------------------------ column half of one work item
// lane = 2 * item + col.  xb: this warp's exchange buffer, E[a][b] of column c at float4 index a * 34 + 2 b + c.
// dst: this warp's 8 KiB of the block slot, float2 index 2 * row + col.
// sign and size class from the static lookup, magnitude from the window's levels (JA98 decode)
__device__ __forceinline__ float2 f_ja98(float2 v, float4 lv) {
    const float ax = fabsf(v.x), ay = fabsf(v.y);
    return make_float2(ax == 0.f ? 0.f : copysignf(ax > 2.f ? lv.y : lv.x, v.x), ay == 0.f ? 0.f : copysignf(ay > 2.f ? lv.w : lv.z, v.y));
}

template <int R, bool JA98>
__device__ __forceinline__ void f_col_front(const uint4 rawA, const uint4 rawB, float4* xb, const float4* s_w4,
                                            const uint8_t* s_lut, const float4* s_h4w, const int lane, float2* colsum_n1,
                                            const float4* levels_blk) {
    const int item = lane >> 1, col = lane & 1;
    // ---- P1: decode, FFT_16 over r (n2 = 32 r + l) for l = item (A) and item + 16 (B), twiddle W_512^(l q)
    {
        float2 vA[16], vB[16];
        const uint32_t wa[4] = {rawA.x, rawA.y, rawA.z, rawA.w}, wb[4] = {rawB.x, rawB.y, rawB.z, rawB.w};
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const uint32_t ia = __byte_perm(wa[r >> 2], 0u, 0x4440 + (r & 3));
            const uint32_t ib = __byte_perm(wb[r >> 2], 0u, 0x4440 + (r & 3));
            vA[r] = *reinterpret_cast<const float2*>(s_lut + ia);
            vB[r] = *reinterpret_cast<const float2*>(s_lut + ib);
        }
        if (JA98) {                 // a window of 512 time samples is 512 / R rows of the block (R <= 512)
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int rowA = 32 * r + item;
                vA[r] = f_ja98(vA[r], __ldg(levels_blk + ((rowA * R) >> 9)));
                vB[r] = f_ja98(vB[r], __ldg(levels_blk + (((rowA + 16) * R) >> 9)));
            }
        }
        fft_inreg<16, false>(vA);
        fft_inreg<16, false>(vB);
#pragma unroll
        for (int qp = 0; qp < 8; ++qp) {
            const float4 ta = s_w4[qp * 32 + item], tb = s_w4[qp * 32 + item + 16];
            if (qp) vA[2 * qp] = cmul(vA[2 * qp], make_float2(ta.x, ta.y));
            vA[2 * qp + 1] = cmul(vA[2 * qp + 1], make_float2(ta.z, ta.w));
            if (qp) vB[2 * qp] = cmul(vB[2 * qp], make_float2(tb.x, tb.y));
            vB[2 * qp + 1] = cmul(vB[2 * qp + 1], make_float2(tb.z, tb.w));
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) xb[q * 34 + lane] = make_float4(vA[q].x, vA[q].y, vB[q].x, vB[q].y);
    }
    __syncwarp();
    // ---- middle: FFT_32 over l -> p;  * W_M^(16 p n1);  IFFT_32 over p -> m1   (thread q = item, in place on row q)
    {
        float2 u[32];
        float4* row = xb + item * 34 + col;
#pragma unroll
        for (int l = 0; l < 16; ++l) {
            const float4 t = row[2 * l];
            u[l] = make_float2(t.x, t.y);
            u[l + 16] = make_float2(t.z, t.w);
        }
        fft_inreg<32, false>(u);
        if (item == 0) *colsum_n1 = u[0];                       // A[k2 = 0]: column sum
#pragma unroll
        for (int pp = 0; pp < 16; ++pp) {
            const float4 h = s_h4w[pp * 2 + col];
            if (pp) u[pp] = cmul(u[pp], make_float2(h.x, h.y));
            u[pp + 16] = cmul(u[pp + 16], make_float2(h.z, h.w));
        }
        fft_inreg<32, true>(u);
#pragma unroll
        for (int m = 0; m < 16; ++m) row[2 * m] = make_float4(u[m].x, u[m].y, u[m + 16].x, u[m + 16].y);
    }
    __syncwarp();
}

// ---- P3: * beta^q (W_M^(q n1) conj W_512^(q m1));  IFFT_16 over q -> m2;  rows m1 + 32 m2, m1 = item, item + 16
// Block slot layout: [32 row tiles][R/2 column pairs][16 rows][2 columns] float2.  A store instruction of the warp covers rows
// item + 32 m2 (item = 0..15) of its two columns = one 256-byte (tile, pair) chunk, fully coalesced; and a tile of 16 rows of
// ALL columns is 16 R contiguous bytes for the row pass (one bulk copy).  dst = slot + 32 * pair.
// (Measured alternatives: [pair][512 rows][2]: same store efficiency, but a row is R/2 pieces 8 KiB apart and a row pass that
// streams it from HBM takes twice the time; row-pair-major [256][R/2][2][2]: 32-byte store granules, column half 2.6x slower.)
template <int R>
__device__ __forceinline__ void f_col_back(const float4* xb, const float2 (&betaS)[4], const int lane, float2* dst) {
    {
        float2 pw[16];
        cpowers15(betaS, pw);
        float2 yA[16], yB[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float4 t = xb[q * 34 + lane];
            yA[q] = make_float2(t.x, t.y);
            yB[q] = make_float2(t.z, t.w);
        }
#pragma unroll
        for (int q = 1; q < 16; ++q) {
            yA[q] = cmul(yA[q], pw[q]);
            const float2 cb = make_float2(cos64(2 * q), sin64(2 * q));      // conj(W_32^q): m1 + 16
            yB[q] = cmul(cmul(yB[q], cb), pw[q]);
        }
        fft_inreg<16, true>(yA);
        fft_inreg<16, true>(yB);
        constexpr int TS = (R / 2) * 32;                        // float2 per row tile
#pragma unroll
        for (int m2 = 0; m2 < 16; ++m2) {
            dst[(2 * m2) * TS + lane] = yA[m2];                 // row item + 32 m2      -> tile 2 m2
            dst[(2 * m2 + 1) * TS + lane] = yB[m2];             // row item + 16 + 32 m2 -> tile 2 m2 + 1
        }
    }
}

// ------------------------------------------------------------------ row half
// FFT_R of RW rows held in shared memory (TR lanes per row, PT points per lane): radix PT in registers,
// transposition through `myx` (which may overwrite the rows once everybody has loaded them), radix TR.
// z[j][t] = Z_c at c = (s + TR j) + PT t.
template <int TR, int PT>
__device__ __forceinline__ void f_row_fft(const float2* tile, float2* myx, const float2* s_tw, const int s,
                                          float2 (&z)[PT / TR][TR]) {
    constexpr int QPT = PT / TR;
    float2 v[PT];
#pragma unroll
    for (int a = 0; a < PT; ++a) v[a] = tile[s + TR * a];
    fft_inreg<PT, false>(v);
#pragma unroll
    for (int q = 1; q < PT; ++q) v[q] = cmul(v[q], s_tw[q * TR + s]);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < PT; ++q) myx[s * (PT + 1) + q] = v[q];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < QPT; ++j)
#pragma unroll
        for (int t = 0; t < TR; ++t) z[j][t] = myx[t * (PT + 1) + s + TR * j];
#pragma unroll
    for (int j = 0; j < QPT; ++j) fft_inreg<TR, false>(z[j]);
}

// 16-byte pieces (one column pair of one row) of rows [m0, m0 + RPW) of a block slot -> this warp's row-major tile
template <int R>
__device__ __forceinline__ void f_row_load(const float2* slot, const int m0, uint8_t* wbuf, const int lane) {
    using G = FGeo<R>;
    constexpr int NP = G::NPAIR;
#pragma unroll
    for (int k = 0; k < 16; ++k) {                              // RPW * NPAIR = 512 pieces: lane -> (pair, row), row fastest
        const int idx = lane + 32 * k;
        const int r = idx % G::RPW, pp = idx / G::RPW;
        const int m = m0 + r;
        cp_async16(wbuf + r * G::kPitch + pp * 16, slot + ((((m >> 4) * NP + pp) * 16 + (m & 15)) * 2));
    }
    cp_async_commit();
}

// transform, un-mix, detect and integrate the RPW rows in the tile; write whole (partial) output rows
template <int R, int MODE>
__device__ __forceinline__ void f_row_compute(const FParams& p, uint8_t* wbuf, const float2* s_tw, const float2* eps_blk,
                                              const int64_t gb, const int m0, const int lane) {
    using G = FGeo<R>;
    constexpr int TR = G::TR, PT = G::PT, RW = G::RW, QPT = PT / TR, HP = TR / 2, N = G::N;
    constexpr int NPROD = nprod_of_mode(MODE);
    const int s = lane % TR, rsw = lane / TR;
    const int Dp = p.Dp;
    const int GW = Dp > RW ? Dp : RW;               // rows per integration group inside this task
    const int ppg = GW / RW;                        // passes per group
    const int nout = GW / Dp;                       // output rows per group (> 1 only if Dp < RW)
    const int spo = RW / nout;                      // row slots that add up to one output row
    float2 e[QPT][HP];
#pragma unroll
    for (int j = 0; j < QPT; ++j)
#pragma unroll
        for (int pp = 0; pp < HP; ++pp) e[j][pp] = ldcg_f2(eps_blk + (s + TR * j) + PT * pp);
    const int ifi = (int)(gb / p.nblk);
    const int64_t blk = gb % p.nblk;
    float acc[QPT][HP][NPROD];
    // passes run from the last rows to the first: the transposition buffer of pass k may then spill into the
    // (already consumed) rows of pass k + 1
#pragma unroll 1
    for (int k = G::NPASS - 1; k >= 0; --k) {
        if (k % ppg == ppg - 1) {
#pragma unroll
            for (int j = 0; j < QPT; ++j)
#pragma unroll
                for (int pp = 0; pp < HP; ++pp)
#pragma unroll
                    for (int c = 0; c < NPROD; ++c) acc[j][pp][c] = 0.f;
        }
        float2 z[QPT][TR];
        f_row_fft<TR, PT>(reinterpret_cast<const float2*>(wbuf + (k * RW + rsw) * G::kPitch),
                          reinterpret_cast<float2*>(wbuf + k * RW * G::kPitch) + rsw * TR * (PT + 1), s_tw, s, z);
        // mirror channel R-1-c lives in lane s ^ (TR-1), register [QPT-1-j][TR-1-pp]
#pragma unroll
        for (int j = 0; j < QPT; ++j)
#pragma unroll
            for (int pp = 0; pp < HP; ++pp) {
                const float2 a = z[j][pp];
                const float2 bs = z[QPT - 1 - j][TR - 1 - pp];
                const float bx = __shfl_xor_sync(0xffffffffu, bs.x, TR - 1);
                const float by = __shfl_xor_sync(0xffffffffu, bs.y, TR - 1);
                const float2 bp = make_float2(bx - e[j][pp].x, -by - e[j][pp].y);
                if (MODE == B2F_POL_I) {
                    float t = a.x * a.x;
                    t = fmaf(a.y, a.y, t);
                    t = fmaf(bp.x, bp.x, t);
                    t = fmaf(bp.y, bp.y, t);
                    acc[j][pp][0] = fmaf(0.5f, t, acc[j][pp][0]);
                } else {
                    detect_acc<MODE>(acc[j][pp], make_float2(a.x + bp.x, a.y + bp.y), make_float2(a.x - bp.x, a.y - bp.y));
                }
            }
        if (k % ppg == 0) {
            for (int m = TR; m < TR * spo; m <<= 1) {
#pragma unroll
                for (int j = 0; j < QPT; ++j)
#pragma unroll
                    for (int pp = 0; pp < HP; ++pp)
#pragma unroll
                        for (int c = 0; c < NPROD; ++c) acc[j][pp][c] += __shfl_xor_sync(0xffffffffu, acc[j][pp][c], m);
            }
            if (rsw % spo == 0) {
                const int g0 = m0 + (k / ppg) * GW;             // first row of this group inside the block
                const int64_t t = p.out_row0 + (blk * kL + g0) / Dp + rsw / spo;
                float* dst = p.out + ifi * p.out_if_stride + t * (int64_t)(NPROD * N);
#pragma unroll
                for (int c = 0; c < NPROD; ++c)
#pragma unroll
                    for (int j = 0; j < QPT; ++j)
#pragma unroll
                        for (int pp = 0; pp < HP; ++pp) dst[c * N + (s + TR * j) + PT * pp] = acc[j][pp][c];
            }
        }
    }
}

// eps_c = conj(G[R-1-c] - G[(R-c) mod R]), G = FFT_R(column sums), by one warp
template <int R>
__device__ __forceinline__ void f_eps(const float2* colsum_blk, float2* eps_blk, uint8_t* wbuf, const float2* s_tw,
                                      const int lane) {
    using G = FGeo<R>;
    constexpr int TR = G::TR, PT = G::PT, RW = G::RW, QPT = PT / TR;
    const int s = lane % TR, rsw = lane / TR;
    for (int i = lane; i < R; i += 32) {
        const float2 v = ldcg_f2(colsum_blk + i);
#pragma unroll
        for (int rs = 0; rs < RW; ++rs) reinterpret_cast<float2*>(wbuf + rs * G::kPitch)[i] = v;
    }
    __syncwarp();
    float2 z[QPT][TR];
    f_row_fft<TR, PT>(reinterpret_cast<const float2*>(wbuf + rsw * G::kPitch),
                      reinterpret_cast<float2*>(wbuf) + rsw * TR * (PT + 1), s_tw, s, z);
    __syncwarp();
    float2* Gs = reinterpret_cast<float2*>(wbuf);
    if (rsw == 0) {
#pragma unroll
        for (int j = 0; j < QPT; ++j)
#pragma unroll
            for (int t = 0; t < TR; ++t) Gs[(s + TR * j) + PT * t] = z[j][t];
    }
    __syncwarp();
    for (int c = lane; c < R / 2; c += 32) {
        const float2 g0 = Gs[R - 1 - c], g1 = Gs[(R - c) & (R - 1)];
        eps_blk[c] = make_float2(g0.x - g1.x, -(g0.y - g1.y));
    }
    __syncwarp();
}

// JA98 is a template parameter: as a run-time branch it cost the default decode 324 bytes of spills in the hot loop and
// 20 % of the column half's speed.
template <int R, int MODE, bool JA98>
__global__ void __launch_bounds__(kFThreads, 1) kf_fused(const FParams p) {
    using G = FGeo<R>;
    extern __shared__ __align__(128) uint8_t kf_smem[];
    float4* s_w4 = reinterpret_cast<float4*>(kf_smem + G::kOffW4);
    uint8_t* s_lut = kf_smem + G::kOffLut;
    float2* s_tw = reinterpret_cast<float2*>(kf_smem + G::kOffTw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4* s_h4w = reinterpret_cast<float4*>(kf_smem + G::kOffH + warp * 512);
    uint8_t* wbuf = kf_smem + G::kOffBuf + (size_t)warp * G::kWarpBuf;

    // ---- tables shared by the CTA
    for (int i = tid; i < 8 * 32; i += kFThreads) {
        const int qp = i >> 5, l = i & 31;
        const float2 w0 = p.tab_w[l * 16 + 2 * qp], w1 = p.tab_w[l * 16 + 2 * qp + 1];
        s_w4[i] = make_float4(w0.x, w0.y, w1.x, w1.y);
    }
    if (tid < 32) {
        const int c0 = tid & 3, c1 = (tid >> 2) & 3;
        const float m0 = (c0 == 0 || c0 == 3) ? kLevHi : kLevLo;
        const float m1 = (c1 == 0 || c1 == 3) ? kLevHi : kLevLo;
        if (tid <= 16)
            reinterpret_cast<float2*>(s_lut)[tid] = tid < 16 ? make_float2((c0 & 2) ? m0 : -m0, (c1 & 2) ? m1 : -m1) : make_float2(0.f, 0.f);
    }
    for (int i = tid; i < G::PT * G::TR; i += kFThreads) s_tw[i] = p.tab_r[i];

    // ---- this warp's static place: lane `lam` (one block per round), column pair `pr`.  Consecutive warps belong to
    // different lanes, so the 16 warps of an SM work on 16 different blocks: when one lane waits, the others keep the
    // SM busy (with whole lanes per SM every inter-warp wait idled the SM).
    const int gw = blockIdx.x * kFWarps + warp;
    const int nl = (int)(gridDim.x * kFWarps) / G::NPAIR;
    const int lam = gw % nl, pr = gw / nl;
    const bool active = pr < G::NPAIR;
    const int item = lane >> 1, col = lane & 1;
    const int n1 = 2 * (active ? pr : 0) + col;
    {
        const int pp = lane >> 1;                                // 16 entries x 2 columns
        const float2 h0 = p.tab_h[pp * R + n1], h1 = p.tab_h[(pp + 16) * R + n1];
        s_h4w[pp * 2 + col] = make_float4(h0.x, h0.y, h1.x, h1.y);
    }
    float2 betaS[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) betaS[k] = p.tab_beta[(size_t)(item * 4 + k) * R + n1];
    __syncthreads();
    if (!active) return;

    const int64_t nb = p.gb_end - p.gb_begin;
    const int64_t M = (int64_t)R * kL;
    unsigned* sy = p.sync + (size_t)lam * FS_STRIDE;
    const int phase = p.phase, ns = p.nslot;
    const int lag = phase == 0 ? p.lag : 0;                  // rounds between a block's columns and its rows

    auto raw_ptr = [&](int64_t lb) {
        const int64_t gb = p.gb_begin + lb;
        const int ifi = (int)(gb / p.nblk);
        const int64_t blk = gb % p.nblk;
        return p.tstream + ifi * p.tstream_if_stride + blk * M + (size_t)pr * 1024 + (size_t)lane * 16;
    };
    uint4 rawA = make_uint4(0, 0, 0, 0), rawB = rawA;
    if (phase != 2 && lam < nb) {
        const uint8_t* q = raw_ptr(lam);
        rawA = ldg_stream16(q);
        rawB = ldg_stream16(q + 512);
    }

    // per-section cycle counters: compiled in only with -DB2F_KF_PROF (16 live registers otherwise spill in the hot loop:
    // the first instrumented build made the column half 20 % slower)
#ifdef B2F_KF_PROF
    unsigned long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tprev = clock64();
    auto tick = [&](int k) {
        if (p.prof) {
            const long long t = clock64();
            tacc[k] += (unsigned long long)(t - tprev);
            tprev = t;
        }
    };
#else
    auto tick = [](int) {};
#endif
    // Round i of a lane: rows of block i - lag, then (one rotating warp) eps of block i - 1, then the columns of block i.
    // Every wait is for something the other warps of the lane did about a round earlier: the rows of block i - lag need
    // the columns published early in round i - lag + 1 and the eps computed in the middle of it; the ring slot the
    // columns of block i go to (lag + 1 slots per lane) was read at the start of round i - 1.  The three counters are
    // therefore read once at the top of the round, long before they are needed; a polling loop only runs if such a value
    // is not yet sufficient.  A round's own column result is published at the NEXT round's row stage, after a load wait,
    // when its stores have long left the SM: the gpu-scope fence in front of the counter update is then cheap (directly
    // after the stores it cost a store round trip per work item, 6 % of the kernel).
    bool col_pending = false;
#pragma unroll 1
    for (int i = 0;; ++i) {
        const int64_t lbc = (int64_t)i * nl + lam;                         // block whose columns run this round
        const int64_t lbr = lbc - (int64_t)lag * nl;                       // block whose rows run this round
        const int64_t lbe = lbc - nl;                                      // block whose eps is due this round
        const bool vc = phase != 2 && lbc < nb;
        const bool vr = phase != 1 && lbr >= 0 && lbr < nb;
        const bool ve = phase == 0 && lbe >= 0 && lbe < nb && pr == (i % G::NPAIR);
        if (lbr >= nb || lam >= nb) break;                                 // this lane's last block has had its rows
        __syncwarp();
        unsigned c_col = 0, c_row = 0, c_eps = 0;
        if (phase == 0) {                                                  // all lanes read the same words: one transaction each
            c_col = ld_relaxed_u32(sy + FS_COL);
            c_eps = ld_relaxed_u32(sy + FS_EPS);
            c_row = ld_relaxed_u32(sy + FS_ROW);
        }
        tick(7);
        if (vr) {
            const int64_t gb = p.gb_begin + lbr;
            const int64_t slot = phase == 0 ? (int64_t)lam * ns + ((i - lag) % ns) : lbr;
            if (phase == 0) {
                if (c_col < (unsigned)(G::NPAIR * (i - lag + 1)) &&
                    !warp_wait_ge(sy + FS_COL, (unsigned)(G::NPAIR * (i - lag + 1)), p.abort_flag, lane)) return;
                if (c_eps < (unsigned)(i - lag + 1) && !warp_wait_ge(sy + FS_EPS, (unsigned)(i - lag + 1), p.abort_flag, lane)) return;
            }
            tick(3);
            f_row_load<R>(p.ring + slot * M, pr * G::RPW, wbuf, lane);
            cp_async_wait<0>();
            __syncwarp();
            tick(4);
        }
        if (col_pending) { warp_arrive(sy + FS_COL, lane); col_pending = false; }     // last round's columns
        if (vr && phase == 0) warp_arrive_relaxed(sy + FS_ROW, lane);                  // the slot has been read
        if (vc && i > 0 && phase != 1) {                                   // input of this round's columns (round 0: loaded above)
            const uint8_t* q = raw_ptr(lbc);
            rawA = ldg_stream16(q);
            rawB = ldg_stream16(q + 512);
        }
        tick(5);
        if (vr) {
            const int64_t gb = p.gb_begin + lbr;
            f_row_compute<R, MODE>(p, wbuf, s_tw, p.eps + gb * G::N, gb, pr * G::RPW, lane);
            __syncwarp();
            tick(6);
        }
        if (ve) {
            if (!warp_wait_ge(sy + FS_COL, (unsigned)(G::NPAIR * i), p.abort_flag, lane)) return;
            const int64_t gb = p.gb_begin + lbe;
            f_eps<R>(p.colsum + gb * R, p.eps + gb * G::N, wbuf, s_tw, lane);
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                atomicMax(sy + FS_EPS, (unsigned)i);
            }
            tick(3);
        }
        if (vc) {
            const int64_t gb = p.gb_begin + lbc;
            const int64_t slot = phase == 0 ? (int64_t)lam * ns + (i % ns) : lbc;
            f_col_front<R, JA98>(rawA, rawB, reinterpret_cast<float4*>(wbuf), s_w4, s_lut, s_h4w, lane, p.colsum + gb * R + n1,
                                 JA98 ? p.levels + gb * R : nullptr);
            tick(0);
            // the slot written now was read by the row halves of the block `ns` rounds back
            if (phase == 0 && i >= ns && c_row < (unsigned)(G::NPAIR * (i - ns + 1))) {
                if (!warp_wait_ge(sy + FS_ROW, (unsigned)(G::NPAIR * (i - ns + 1)), p.abort_flag, lane)) return;
            }
            tick(1);
            if (phase == 1 && lbc + nl < nb) {      // columns only: the next work item's input, in flight during the back half
                const uint8_t* q = raw_ptr(lbc + nl);
                rawA = ldg_stream16(q);
                rawB = ldg_stream16(q + 512);
            }
            f_col_back<R>(reinterpret_cast<const float4*>(wbuf), betaS, lane, p.ring + slot * M + (size_t)pr * 32);
            col_pending = phase == 0;
            tick(2);
        }
        if (lag < 2 && col_pending) { warp_arrive(sy + FS_COL, lane); col_pending = false; }   // lag 1: next round's rows wait for it
    }
    if (col_pending) warp_arrive(sy + FS_COL, lane);
    // 0 column front, 1 wait for the ring slot, 2 column back + stores, 3 eps / publish / wait for the block's columns,
    // 4 row load (issue, eps wait, data), 5 arrivals, 6 row compute, 7 loop overhead
#ifdef B2F_KF_PROF
    if (p.prof && lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(&p.prof[(size_t)gw * 8 + k], tacc[k]);
    }
#endif
}

// ================================================================== row pass over 16-row tiles (R = 256)
// Consumer of the round-2 column kernel when the intermediate goes through HBM (B2F_PATH=split): a CTA streams whole 16-row
// tiles -- 32 KiB contiguous in the [tile][pair][16 rows][2] block layout, ONE bulk copy (TMA, mbarrier) per tile, double
// buffered -- so DRAM sees nothing but long sequential reads.  The tile arrives pair-major; instead of re-laying it out, the
// first radix-16 stage is spread so that the layout is conflict-free as it is: warp w, lane (row r, column c) takes the points
// n1 = (2w + c) + 16a, a = 0..15 (pairs w + 8a), i.e. the 32 lanes of a load instruction differ in (r, c) only: 32 different
// 8-byte bank slots.  The exchange between the two radix-16 stages is block-wide, X[q][s][r ^ 8 (s & 1)] (both sides
// conflict-free without padding).  Second stage: warp w owns q = w and 15 - w -- a channel and its mirror, one lane-xor-16
// shuffle apart -- for all 16 rows (lane = r + 16 j): FFT over s, detection, and the time integration over the rows of the
// tile is a butterfly of shuffles inside the warp; sums over several tiles (tscrunch > 16) stay in registers.  A tile buffer is
// dead after the first stage, so it is refilled right after the barrier: two bulk copies in flight per CTA.  One block
// barrier per tile (exchange written -> read); "everybody has read the exchange buffer" is a split mbarrier: arrive after the
// reads, wait only before the next tile's writes, with that tile's loads and first FFT stage in between.  tscrunch 1..512.
struct KTParams {
    const float2* inter;        // [blocks][32 tiles][128 pairs][16 rows][2 columns]
    const float2* eps;          // [blocks][128]
    const float2* tab_r;        // [16][16] W_256^(s q), q-major
    float* F; int64_t F_if_stride; int64_t row0;
    int nblk, nif, D;
    int64_t nb;                 // blocks of this launch
};
constexpr int kKTThreads = 256;
constexpr int kKTTile = 16 * 256 * (int)sizeof(float2);          // 32 KiB
constexpr int kKTEps = 128 * (int)sizeof(float2);
struct KTSmem {
    static constexpr int kOffTile = 128;
    static constexpr int kOffX = kOffTile + 2 * (kKTTile + kKTEps);
    static constexpr int kOffTw = kOffX + 16 * 256 * (int)sizeof(float2);
    static constexpr size_t kBytes = (size_t)kOffTw + 256 * sizeof(float2);
};
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(kKTThreads, 2) kt_row_tiles(const KTParams p) {
    constexpr int NPROD = nprod_of_mode(MODE), N = 128;
    using S = KTSmem;
    extern __shared__ __align__(128) uint8_t kt_smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(kt_smem);            // [0], [1]: tile landed; [2]: exchange buffer read
    float2* X = reinterpret_cast<float2*>(kt_smem + S::kOffX);
    float2* s_tw = reinterpret_cast<float2*>(kt_smem + S::kOffTw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rA = lane >> 1, cA = lane & 1, sA = 2 * warp + cA;     // first stage: row, column of the pair, s
    const int rB = lane & 15, qB = (lane >> 4) ? 15 - warp : warp;   // second stage: row, q
    s_tw[tid] = p.tab_r[tid];
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_init(&bar[2], kKTThreads);
        mbar_fence_init();
    }
    __syncthreads();

    const int D = p.D;
    const int lgG = D > 16 ? 31 - __clz(D / 16) : 0;         // 2^lgG tiles per integration unit
    const int G = 1 << lgG, lgU = 5 - lgG;                   // 2^lgU units per block
    const int Dt = D < 16 ? D : 16;                          // rows of one tile that add up to one output row
    const int lgD = 31 - __clz(D);
    int ifi = 0;
    const int nunits_cta = (int)((p.nb * (32 >> lgG) - blockIdx.x + gridDim.x - 1) / gridDim.x);   // units this CTA owns
    const int nseq = nunits_cta > 0 ? nunits_cta << lgG : 0;
    auto locate = [&](int seq, int& lb, int& rt) {
        const int64_t u = blockIdx.x + (int64_t)(seq >> lgG) * gridDim.x;
        lb = (int)(u >> lgU);
        rt = ((int)(u & ((1 << lgU) - 1)) << lgG) + (seq & (G - 1));
    };
    auto issue = [&](int lb, int rt, int buf) {              // thread 0 only
        uint8_t* dst = kt_smem + S::kOffTile + buf * (kKTTile + kKTEps);
        mbar_expect_tx(&bar[buf], kKTTile + kKTEps);
        bulk_g2s(dst, p.inter + ((int64_t)lb * 32 + rt) * (128 * 32), kKTTile, &bar[buf]);
        bulk_g2s(dst + kKTTile, p.eps + (int64_t)lb * N, kKTEps, &bar[buf]);
    };
    int lb, rt, lbn, rtn;
    if (tid == 0) {                                          // two tiles in flight per CTA
        for (int k = 0; k < 2 && k < nseq; ++k) {
            locate(k, lb, rt);
            issue(lb, rt, k);
        }
    }
    float acc[8][NPROD];
#pragma unroll
    for (int pp = 0; pp < 8; ++pp)
#pragma unroll
        for (int c = 0; c < NPROD; ++c) acc[pp][c] = 0.f;
    float2* xw = X + sA * 16 + (rA ^ (cA << 3));             // + 256 q
    const float2* xr0 = X + qB * 256 + rB;                   // + 16 t, even t
    const float2* xr1 = X + qB * 256 + (rB ^ 8);             //         odd t
#pragma unroll 1
    for (int seq = 0; seq < nseq; ++seq) {
        const int buf = seq & 1;
        locate(seq, lb, rt);
        mbar_wait(&bar[buf], (uint32_t)((seq >> 1) & 1));
        const float2* T = reinterpret_cast<const float2*>(kt_smem + S::kOffTile + buf * (kKTTile + kKTEps));
        const float2* s_eps = T + 16 * 256;
        // ---- first stage: FFT_16 over a of x[s + 16 a], twiddle W_256^(s q)
        {
            float2 v[16];
#pragma unroll
            for (int a = 0; a < 16; ++a) v[a] = T[((warp + 8 * a) * 16 + rA) * 2 + cA];
            fft_inreg<16, false>(v);
#pragma unroll
            for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], s_tw[q * 16 + sA]);
            if (seq > 0) mbar_wait(&bar[2], (uint32_t)((seq - 1) & 1));      // the previous tile's exchange values have been read
#pragma unroll
            for (int q = 0; q < 16; ++q) xw[q * 256] = v[q];
        }
        float2 e[8];                                         // eps of this tile's block, before the buffer is given away
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) e[pp] = s_eps[qB + 16 * pp];
        __syncthreads();
        // the tile has been consumed (rows in the first stage, eps just now): refill its buffer with the tile after next,
        // so that two bulk copies are in flight per CTA with two buffers
        if (tid == 0 && seq + 2 < nseq) {
            locate(seq + 2, lbn, rtn);
            fence_proxy_async();
            issue(lbn, rtn, buf);
        }
        // ---- second stage: FFT_16 over s -> channels c = q + 16 pp; mirror R-1-c in lane ^ 16, register 15 - pp
        {
            float2 z[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) z[t] = (t & 1) ? xr1[t * 16] : xr0[t * 16];
            mbar_arrive(&bar[2]);                            // this thread is done with the exchange buffer
            fft_inreg<16, false>(z);
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
                const float2 a = z[pp];
                const float2 bs = z[15 - pp];
                const float bx = __shfl_xor_sync(0xffffffffu, bs.x, 16);
                const float by = __shfl_xor_sync(0xffffffffu, bs.y, 16);
                const float2 bp = make_float2(bx - e[pp].x, -by - e[pp].y);
                if (MODE == B2F_POL_I) {
                    float t = a.x * a.x;
                    t = fmaf(a.y, a.y, t);
                    t = fmaf(bp.x, bp.x, t);
                    t = fmaf(bp.y, bp.y, t);
                    acc[pp][0] = fmaf(0.5f, t, acc[pp][0]);
                } else {
                    detect_acc<MODE>(acc[pp], make_float2(a.x + bp.x, a.y + bp.y), make_float2(a.x - bp.x, a.y - bp.y));
                }
            }
        }
        // ---- an output row is complete: add up its rows (lanes differ in r) and store
        if ((rt & (G - 1)) == G - 1) {
            while ((int64_t)(ifi + 1) * p.nblk <= lb) ++ifi;                 // lb only grows: no division
            const int blk = lb - ifi * p.nblk;
            if (Dt == 16) {
                // all 16 rows: reduce-scatter over the lanes (8 shuffles instead of a 32-shuffle butterfly).  After the steps
                // with lane masks 8, 4, 2 a lane holds the sum of one channel group pp = r >> 1 over 8 rows; mask 1 finishes it.
                const int64_t t = p.row0 + (((int64_t)blk * kL + ((rt >> lgG) << lgG) * 16) >> lgD);
                float* dst = p.F + ifi * p.F_if_stride + t * (int64_t)(NPROD * N) + qB + 16 * (rB >> 1);
#pragma unroll
                for (int c = 0; c < NPROD; ++c) {
                    float a4[4], a2[2], a1;
                    const bool h8 = (rB & 8) != 0, h4 = (rB & 4) != 0, h2 = (rB & 2) != 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float keep = h8 ? acc[k + 4][c] : acc[k][c], send = h8 ? acc[k][c] : acc[k + 4][c];
                        a4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float keep = h4 ? a4[k + 2] : a4[k], send = h4 ? a4[k] : a4[k + 2];
                        a2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                    }
                    {
                        const float keep = h2 ? a2[1] : a2[0], send = h2 ? a2[0] : a2[1];
                        a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                    }
                    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
                    if ((rB & 1) == 0) dst[c * N] = a1;
                }
            } else {
                for (int m = 1; m < Dt; m <<= 1) {
#pragma unroll
                    for (int pp = 0; pp < 8; ++pp)
#pragma unroll
                        for (int c = 0; c < NPROD; ++c) acc[pp][c] += __shfl_xor_sync(0xffffffffu, acc[pp][c], m);
                }
                if ((rB & (Dt - 1)) == 0) {
                    const int64_t t = p.row0 + (((int64_t)blk * kL + rt * 16 + rB) >> lgD);
                    float* dst = p.F + ifi * p.F_if_stride + t * (int64_t)(NPROD * N) + qB;
#pragma unroll
                    for (int c = 0; c < NPROD; ++c)
#pragma unroll
                        for (int pp = 0; pp < 8; ++pp) dst[c * N + 16 * pp] = acc[pp][c];
                }
            }
#pragma unroll
            for (int pp = 0; pp < 8; ++pp)
#pragma unroll
                for (int c = 0; c < NPROD; ++c) acc[pp][c] = 0.f;
        }
    }
}

// ================================================================== front end of the fused path
// Header-only validation: one thread per frame.  fstat = 1 usable, 2 dead (invalid bit or bad header).
struct K0HParams {
    const uint8_t* frames[B2F_MAX_IF];
    uint8_t* fstat; size_t fstat_stride;
    unsigned long long* counters;
    int64_t nframes;
    int frame_bytes, header_bytes, in_nbit, fps, nif;
    uint32_t base_sec[B2F_MAX_IF], base_fnum[B2F_MAX_IF];
};

// Validation + fill masking + block transposition: the 2-bit payload of one FFT block (512 rows of R time
// samples) becomes the index-byte stream the fused kernel reads, [pair][half][lane = 2 item + col][r]:
// byte r = sample (row 32 r + item + 16 half, column 2 pair + col), value (code1 << 2 | code0) << 3,
// 0x80 for samples that decode to 0.0.  One CTA handles a strip of SC = min(32, R) columns of one block:
// pieces of SC/2 payload bytes never straddle a frame.
struct K0TParams {
    const uint8_t* frames[B2F_MAX_IF];
    const uint8_t* fstat; size_t fstat_stride;
    int* fillflag; size_t fillflag_stride;                       // one int per frame: set when a fill word was seen
    uint8_t* tstream; size_t tstream_if_stride;
    unsigned long long* counters;
    int nblk, nif, R, frame_bytes, header_bytes, payload_bytes, mask_faults;
};

#ifdef B2F_API_TU
static __global__ void k0h_headers(const K0HParams p) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool in = i < p.nframes * p.nif;
    int bad = 0, inv = 0, mis = 0, alive = 0;
    if (in) {
        const int ifi = (int)(i / p.nframes);
        const int64_t f = i % p.nframes;
        const uint4 h = *reinterpret_cast<const uint4*>(p.frames[ifi] + f * p.frame_bytes);
        const bool invalid = (h.x >> 31) != 0;
        const bool badh = ((h.z & 0xFFFFFFu) * 8u != (uint32_t)p.frame_bytes) || ((int)((h.w >> 26) & 31u) + 1 != p.in_nbit) ||
                          ((int)((h.x >> 30) & 1u) != (p.header_bytes == 16 ? 1 : 0));
        const int64_t tslot = ((int64_t)(h.x & 0x3FFFFFFFu) - (int64_t)p.base_sec[ifi]) * p.fps +
                              ((int64_t)(h.y & 0xFFFFFFu) - (int64_t)p.base_fnum[ifi]);
        p.fstat[ifi * p.fstat_stride + f] = (invalid || badh) ? 2 : 1;
        bad = badh;
        inv = !badh && invalid;
        alive = !badh && !invalid;
        mis = tslot != f;
    }
    const unsigned nb = __popc(__ballot_sync(0xffffffffu, bad)), ni = __popc(__ballot_sync(0xffffffffu, inv));
    const unsigned nm = __popc(__ballot_sync(0xffffffffu, mis)), na = __popc(__ballot_sync(0xffffffffu, alive));
    if ((threadIdx.x & 31) == 0) {
        if (nb) atomicAdd(&p.counters[C_BADHDR], (unsigned long long)nb);
        if (ni) atomicAdd(&p.counters[C_INVALID], (unsigned long long)ni);
        if (nm) atomicAdd(&p.counters[C_MISPLACED], (unsigned long long)nm);
        if (na) atomicAdd(&p.counters[C_OK], (unsigned long long)na);      // frames with fill are moved out by k0t
    }
}

template <int SC>
static __global__ void __launch_bounds__(256) k0t_transpose(const K0TParams p) {
    constexpr int PB = SC / 2;                   // payload bytes per row piece (a whole 32-byte sector for SC = 64)
    constexpr int NW = PB / 4;                   // 32-bit words per piece
    constexpr int PITCH = SC + 8;                // bytes between rows of expanded index bytes: conflict-free word reads
    __shared__ __align__(16) uint8_t s_exp[kL * PITCH];
    const int tid = threadIdx.x;
    const int R = p.R, nstrip = R / SC;
    const int64_t M = (int64_t)R * kL;
    const int64_t nunits = (int64_t)p.nif * p.nblk * nstrip;
    unsigned long long nfill_total = 0;
    uint32_t w[2][NW];
    uint8_t dead[2];
    int64_t fr[2];
    int fif = 0;
    // the pieces of the next unit are loaded into registers while this unit's transposed read-out runs
    auto load_unit = [&](int64_t u) {
        const int strip = (int)(u % nstrip);
        const int64_t gb = u / nstrip;
        const int ifi = (int)(gb / p.nblk);
        const int64_t blk = gb % p.nblk;
        const uint8_t* frames = p.frames[ifi];
        fif = ifi;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int row = tid + 256 * k;
            const int64_t pb = (blk * M + (int64_t)R * row + (int64_t)strip * SC) >> 1;     // payload byte of the stream
            const int64_t f = pb / p.payload_bytes;
            const int off = (int)(pb - f * p.payload_bytes);
            const uint8_t* src = frames + f * p.frame_bytes + p.header_bytes + off;
            if (NW >= 4) {
#pragma unroll
                for (int v = 0; v < NW / 4; ++v) {
                    const uint4 x = ldg_stream16(src + 16 * v);
                    w[k][(4 * v) % NW] = x.x; w[k][(4 * v + 1) % NW] = x.y; w[k][(4 * v + 2) % NW] = x.z; w[k][(4 * v + 3) % NW] = x.w;
                }
            } else {
                const uint2 x = *reinterpret_cast<const uint2*>(src);
                w[k][0] = x.x; w[k][1 % NW] = x.y;
            }
            dead[k] = p.fstat[ifi * p.fstat_stride + f];
            fr[k] = f;
        }
    };
    if (blockIdx.x < nunits) load_unit(blockIdx.x);
    for (int64_t u = blockIdx.x; u < nunits; u += gridDim.x) {
        const int strip = (int)(u % nstrip);
        const int64_t gb = u / nstrip;
        const int ifi = (int)(gb / p.nblk);
        const int64_t blk = gb % p.nblk;
        // ---- validate + expand what was loaded: two rows per thread
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int row = tid + 256 * k;
            const bool isdead = dead[k] == 2;
            unsigned fm = 0;
#pragma unroll
            for (int j = 0; j < NW; ++j) fm |= (unsigned)(w[k][j] == kFillWord) << j;
            if (fm && !isdead) {
                nfill_total += __popc(fm);
                if (p.mask_faults && atomicExch(&p.fillflag[fif * p.fillflag_stride + fr[k]], 1) == 0) {
                    atomicAdd(&p.counters[C_FILLFRAMES], 1ull);
                    atomicAdd(&p.counters[C_OK], ~0ull);                    // -1: alive, but not clean
                }
            }
            const unsigned mask = !p.mask_faults ? 0u : (isdead ? 0xFFFFu : fm);
            uint2* d = reinterpret_cast<uint2*>(s_exp + row * PITCH);
#pragma unroll
            for (int j = 0; j < NW; ++j) d[j] = expand_word_2bit(w[k][j], (mask >> j) & 1);
        }
        __syncthreads();
        if (u + gridDim.x < nunits) load_unit(u + gridDim.x);
        // ---- transposed read-out: thread = (half, group of 8 columns, item); 16 rows x 8 columns (64-bit shared loads,
        // conflict-free at a row pitch of 9 x 8 bytes) -> 8 outputs of 16 bytes
        {
            const int lane = tid & 31, wrp = tid >> 5;
            const int item = lane & 15;
            const int half = wrp >> 2;
            const int c8 = (wrp & 3) * 2 + (lane >> 4);
            uint8_t* tb = p.tstream + ifi * p.tstream_if_stride + blk * M;
            if (c8 < SC / 8) {
                uint32_t o[8][4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint2 a[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        a[j] = *reinterpret_cast<const uint2*>(s_exp + (32 * (4 * g + j) + item + 16 * half) * PITCH + 8 * c8);
                    {
                        const uint32_t t0 = __byte_perm(a[0].x, a[1].x, 0x5140), t1 = __byte_perm(a[2].x, a[3].x, 0x5140);
                        const uint32_t t2 = __byte_perm(a[0].x, a[1].x, 0x7362), t3 = __byte_perm(a[2].x, a[3].x, 0x7362);
                        o[0][g] = __byte_perm(t0, t1, 0x5410);
                        o[1][g] = __byte_perm(t0, t1, 0x7632);
                        o[2][g] = __byte_perm(t2, t3, 0x5410);
                        o[3][g] = __byte_perm(t2, t3, 0x7632);
                    }
                    {
                        const uint32_t t0 = __byte_perm(a[0].y, a[1].y, 0x5140), t1 = __byte_perm(a[2].y, a[3].y, 0x5140);
                        const uint32_t t2 = __byte_perm(a[0].y, a[1].y, 0x7362), t3 = __byte_perm(a[2].y, a[3].y, 0x7362);
                        o[4][g] = __byte_perm(t0, t1, 0x5410);
                        o[5][g] = __byte_perm(t0, t1, 0x7632);
                        o[6][g] = __byte_perm(t2, t3, 0x5410);
                        o[7][g] = __byte_perm(t2, t3, 0x7632);
                    }
                }
                // A lane holds both columns of its item for four pairs; written as they are, a store instruction would touch
                // 32 half sectors (lanes 32 bytes apart).  Neighbouring lanes (items 2i, 2i + 1) swap one column each, so that
                // a lane pair writes the 32 contiguous bytes (item, col 0 | col 1) of one sector: 16 whole sectors per store.
                const bool odd = lane & 1;
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    uint32_t rcv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) rcv[k] = __shfl_xor_sync(0xffffffffu, odd ? o[2 * pp][k] : o[2 * pp + 1][k], 1);
                    const int pair = strip * (SC / 2) + 4 * c8 + pp;
                    uint8_t* pb = tb + (size_t)(pair * 2 + half) * 32 * 16;
                    *reinterpret_cast<uint4*>(pb + (2 * (item & ~1) + (int)odd) * 16) =
                        odd ? make_uint4(rcv[0], rcv[1], rcv[2], rcv[3]) : make_uint4(o[2 * pp][0], o[2 * pp][1], o[2 * pp][2], o[2 * pp][3]);
                    *reinterpret_cast<uint4*>(pb + (2 * (item | 1) + (int)odd) * 16) =
                        odd ? make_uint4(o[2 * pp + 1][0], o[2 * pp + 1][1], o[2 * pp + 1][2], o[2 * pp + 1][3]) : make_uint4(rcv[0], rcv[1], rcv[2], rcv[3]);
                }
            }
        }
        __syncthreads();
    }
    // fill words: one atomic per warp
    for (int m = 16; m; m >>= 1) nfill_total += __shfl_xor_sync(0xffffffffu, nfill_total, m);
    if ((tid & 31) == 0 && nfill_total) atomicAdd(&p.counters[C_FILLWORDS], nfill_total);
}

// JA98 decode (B2F_DECODE_JA98): one warp per window of 512 time samples of one IF counts, per polarisation, the samples
// between the thresholds (codes 1 and 2) among the samples that are not masked, and turns the fraction into the two
// output magnitudes (oracle: ja98_levels).  256 payload bytes per window; words never straddle a frame.
struct KJParams {
    const uint8_t* frames[B2F_MAX_IF];
    const uint8_t* fstat; size_t fstat_stride;
    float4* levels;                      // [nif*nblk][R]
    int nblk, nif, R, frame_bytes, header_bytes, payload_bytes, mask_faults;
};
static __global__ void kj_ja98_levels(const KJParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;          // window number over all IFs
    const int64_t nwin = (int64_t)p.nif * p.nblk * p.R;
    if (w >= nwin) return;
    const int ifi = (int)(w / ((int64_t)p.nblk * p.R));
    const int64_t wi = w % ((int64_t)p.nblk * p.R);                                    // window inside this IF's push
    unsigned lowP = 0, lowQ = 0, nval = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int64_t pb = wi * 256 + (lane + 32 * k) * 4;
        const int64_t f = pb / p.payload_bytes;
        const int off = (int)(pb - f * p.payload_bytes);
        const uint32_t x = *reinterpret_cast<const uint32_t*>(p.frames[ifi] + f * p.frame_bytes + p.header_bytes + off);
        const bool masked = p.mask_faults && (x == kFillWord || p.fstat[ifi * p.fstat_stride + f] == 2);
        if (!masked) {
            const uint32_t m = x ^ (x >> 1);                 // bit 0 of every 2-bit field: code is 1 or 2
            lowP += __popc(m & 0x11111111u);
            lowQ += __popc(m & 0x44444444u);
            nval += 8;
        }
    }
    for (int m = 16; m; m >>= 1) {
        lowP += __shfl_xor_sync(0xffffffffu, lowP, m);
        lowQ += __shfl_xor_sync(0xffffffffu, lowQ, m);
        nval += __shfl_xor_sync(0xffffffffu, nval, m);
    }
    if (lane == 0) {
        float lv[4];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            double phi = nval ? (double)(q ? lowQ : lowP) / (double)nval : 0.5;
            phi = fmin(fmax(phi, 1.0 / 512.0), 1.0 - 1.0 / 512.0);
            const double t = 1.4142135623730951 * erfinv(phi);
            const double e = exp(-0.5 * t * t);
            lv[2 * q] = (float)(0.7978845608028654 * (1.0 - e) / phi);
            lv[2 * q + 1] = (float)(0.7978845608028654 * e / (1.0 - phi));
        }
        p.levels[w] = make_float4(lv[0], lv[1], lv[2], lv[3]);
    }
}

// F[row][col] = sum of `ratio` consecutive partial rows, fixed order
static __global__ void kt_sum_partials(const float* __restrict__ part, int64_t part_if_stride, float* __restrict__ F,
                                       int64_t F_if_stride, int64_t F_row0, int64_t rows, int ncol, int ratio) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;        // one float4 of one output row
    const int q = ncol / 4;
    if (i >= rows * q) return;
    const int ifi = blockIdx.y;
    const int64_t row = i / q;
    const int c4 = (int)(i % q);
    const float4* src = reinterpret_cast<const float4*>(part + ifi * part_if_stride + row * ratio * (int64_t)ncol) + c4;
    float4 s = src[0];
    for (int k = 1; k < ratio; ++k) {
        const float4 v = src[(int64_t)k * q];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(F + ifi * F_if_stride + (F_row0 + row) * (int64_t)ncol)[c4] = s;
}
#endif

}  // namespace b2f
