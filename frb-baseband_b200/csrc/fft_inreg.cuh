// In-register radix-2 DIT FFTs of length 2..32 with compile-time twiddles.
//
// Every loop is fully unrolled, so indices and twiddles are literals in SASS.  All arithmetic is
// on packed (re, im) pairs: trivial twiddles (1, -i) cost 2 FADD2 per butterfly, every other
// butterfly is 3 FFMA2:  a' = a + w b (2),  b' = 2a - a' (1).
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

// ---- packed FP32 pairs.  Blackwell executes add/mul/fma .f32x2 as one instruction per lane
// (SASS FADD2 / FMUL2 / FFMA2) whose operands can be half-swapped and negated per half for free
// (R.F32x2.LO_HI.NP) and whose multiplier can be a broadcast scalar or immediate.  A complex number
// is one (re, im) pair, so a general butterfly costs 3 issue slots instead of 6 and a twiddle
// multiply 2 instead of 4.  Every formula keeps the twiddle as a pure broadcast (s, s) and puts the
// swap / sign pattern on the data operand, which ptxas folds into the operand modifiers.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned long long f2pack(float2 a) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
    return r;
}
__device__ __forceinline__ float2 f2unpack(unsigned long long r) {
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r));
    return d;
}
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {      // a*b + c per component
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2pack(a)), "l"(f2pack(b)), "l"(f2pack(c)));
    return f2unpack(r);
}
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2pack(a)), "l"(f2pack(b)));
    return f2unpack(r);
}
__device__ __forceinline__ float2 f2add(float2 a, float2 b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2pack(a)), "l"(f2pack(b)));
    return f2unpack(r);
}
#else
inline float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
inline float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
inline float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
#endif

// a * w = w.x (a.x, a.y) + w.y (-a.y, a.x)
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return f2fma(make_float2(-a.y, a.x), make_float2(w.y, w.y), f2mul(a, make_float2(w.x, w.x)));
}
// a * conj(w) = w.x (a.x, a.y) + w.y (a.y, -a.x)
__host__ __device__ __forceinline__ float2 cmul_conj(float2 a, float2 w) {
    return f2fma(make_float2(a.y, -a.x), make_float2(w.y, w.y), f2mul(a, make_float2(w.x, w.x)));
}

// One DIT butterfly  a' = a + w b,  b' = a - w b  with W = exp(-+2 pi i k64/64) (INV: conjugate).
template <bool INV>
__host__ __device__ __forceinline__ void bfly(float2& a, float2& b, int k64) {
    k64 &= 63;
    if (k64 == 0) {
        const float2 t = b;
        b = f2add(a, make_float2(-t.x, -t.y));
        a = f2add(a, t);
    } else if (k64 == 16) {                 // w = -i (fwd): w b = (b.y, -b.x);  +i (inv): (-b.y, b.x)
        const float2 t = INV ? make_float2(-b.y, b.x) : make_float2(b.y, -b.x);
        b = f2add(a, make_float2(-t.x, -t.y));
        a = f2add(a, t);
    } else {
        const float wr = cos64(k64);
        const float wi = INV ? sin64(k64) : -sin64(k64);                       // w = wr + i wi
        float2 t = f2fma(b, make_float2(wr, wr), a);                            // a + wr b
        t = f2fma(make_float2(-b.y, b.x), make_float2(wi, wi), t);              //   + wi (i b)
        b = f2fma(a, make_float2(2.0f, 2.0f), make_float2(-t.x, -t.y));         // 2a - a'
        a = t;
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
