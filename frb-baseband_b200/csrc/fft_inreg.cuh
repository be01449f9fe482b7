// In-register radix-2 DIT FFTs of length 2..32 with compile-time twiddles.
//
// Every loop is fully unrolled, so indices and twiddles are literals in SASS: trivial
// twiddles (1, -i) cost 4 FADD per butterfly, every other butterfly is the 6-FMA form
//     a' = a + w b  (4 FFMA),   b' = 2a - a'  (2 FFMA).
// Input and output are both in natural order; the bit reversal DIT needs is a compile-time
// register renaming.
#pragma once
#include <cuda_runtime.h>

namespace b2f {

// cos(2*pi*k/64), k = 0..16 -- a switch, not an array, so it has no storage and folds to
// literals on both host and device after unrolling
__host__ __device__ constexpr float cos64_q(int k) {
    switch (k) {
        case 0: return 1.00000000000000000000f;
        case 1: return 0.99518472667219692873f;
        case 2: return 0.98078528040323043058f;
        case 3: return 0.95694033573220882438f;
        case 4: return 0.92387953251128673848f;
        case 5: return 0.88192126434835504956f;
        case 6: return 0.83146961230254523567f;
        case 7: return 0.77301045336273699338f;
        case 8: return 0.70710678118654757274f;
        case 9: return 0.63439328416364548779f;
        case 10: return 0.55557023301960228867f;
        case 11: return 0.47139673682599780857f;
        case 12: return 0.38268343236508983729f;
        case 13: return 0.29028467725446233105f;
        case 14: return 0.19509032201612833135f;
        case 15: return 0.09801714032956077016f;
        case 16: return 0.00000000000000000000f;
        default: return 0.0f;
    }
}

__host__ __device__ constexpr float cos64(int k) {
    k &= 63;
    if (k > 32) k = 64 - k;                 // cos is even
    return k <= 16 ? cos64_q(k) : -cos64_q(32 - k);
}
__host__ __device__ constexpr float sin64(int k) { return cos64(k - 16); }

__host__ __device__ constexpr int bitrev(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
__host__ __device__ __forceinline__ float2 cmul_conj(float2 a, float2 w) {   // a * conj(w)
    return make_float2(fmaf(a.x, w.x, a.y * w.y), fmaf(a.y, w.x, -a.x * w.y));
}

// One DIT butterfly with twiddle W = exp(-+2 pi i k64/64) (INV: conjugate).
template <bool INV>
__host__ __device__ __forceinline__ void bfly(float2& a, float2& b, int k64) {
    k64 &= 63;
    if (k64 == 0) {
        float2 t = b;
        b = make_float2(a.x - t.x, a.y - t.y);
        a = make_float2(a.x + t.x, a.y + t.y);
    } else if (k64 == 16) {                 // w = -i (fwd) / +i (inv):  w*b = (b.y, -b.x) / (-b.y, b.x)
        float2 t = INV ? make_float2(-b.y, b.x) : make_float2(b.y, -b.x);
        b = make_float2(a.x - t.x, a.y - t.y);
        a = make_float2(a.x + t.x, a.y + t.y);
    } else {
        const float wr = cos64(k64);
        const float wi = INV ? sin64(k64) : -sin64(k64);
        float tr = fmaf(wr, b.x, a.x);
        tr = fmaf(-wi, b.y, tr);
        float ti = fmaf(wr, b.y, a.y);
        ti = fmaf(wi, b.x, ti);
        b = make_float2(fmaf(2.0f, a.x, -tr), fmaf(2.0f, a.y, -ti));
        a = make_float2(tr, ti);
    }
}

// v[n] natural order in -> v[k] natural order out.  Unnormalised; INV uses exp(+i...).
template <int N, bool INV>
__host__ __device__ __forceinline__ void fft_inreg(float2 (&v)[N]) {
    constexpr int LG = ilog2(N);
    float2 w[N];
#pragma unroll
    for (int i = 0; i < N; ++i) w[i] = v[bitrev(i, LG)];
#pragma unroll
    for (int s = 1; s <= LG; ++s) {
        const int half = 1 << (s - 1);
#pragma unroll
        for (int base = 0; base < N; base += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                // twiddle W_{2*half}^j  ->  k64 = j * 64 / (2*half)
                bfly<INV>(w[base + j], w[base + j + half], j * (32 / half));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = w[i];
}

}  // namespace b2f
