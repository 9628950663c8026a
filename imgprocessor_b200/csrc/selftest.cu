// Self-test kernels: properties of the arithmetic the hot kernels rely on, checked on the device itself (the host
// emulation in tests/host_emul cannot execute MUFU.RCP64H).  Reached through imgcorr_selftest_division.
#include "imgcorr_kernels.cuh"

namespace imgcorr {

__device__ __forceinline__ uint64_t st_mix(uint64_t z) {           // splitmix64
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// One thread per float32 significand pattern of the divisor b (2^23 patterns x sign x an exponent that cycles over the
// whole float32 range, denormals included); `nnum` numerators each, drawn from the shapes the kernels produce:
//   0: (double)uint16 - (double)float32          K1 on integer frames
//   1: (double)float32 - (double)float32         K1 on float32 frames
//   2: random float64 with a full 53-bit significand, |a| <= 1e300     K4's running mean (divisor = small integer)
//   3: RN(m * b) -+ 1 ulp for a random 53-bit m: quotients next to representable values and rounding midpoints
// The quotient of rcp_f32range + ddiv_rcp must equal IEEE a / b (__ddiv_rn) bit for bit.
__global__ void selftest_division_kernel(int nnum, uint64_t seed, unsigned long long* mismatches, unsigned long long* seed_err_bits) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;       // 0 .. 2^24 - 1
    const uint32_t mant = t & 0x7fffffu, sign = (t >> 23) & 1u;
    uint64_t r = st_mix(seed ^ ((uint64_t)t << 20));
    unsigned long long bad = 0;
    double worst = 0.0;
    for (int k = 0; k < nnum; ++k) {
        r = st_mix(r);
        uint32_t e = (uint32_t)(r % 255u);                          // 0 (denormal / zero) .. 254
        uint32_t bits = (sign << 31) | (e << 23) | mant;
        if ((bits & 0x7fffffffu) == 0) bits |= 1u;                  // b != 0 (the kernels divide by the zero-free copy)
        const float bf = __uint_as_float(bits);
        const double b = (double)bf;
        // small integers as divisors as well (K4): replace one draw in 16
        const double bd = ((r >> 8) & 15u) == 0 ? (double)(1 + (int)((r >> 12) % 4096u)) : b;
        double a;
        const uint64_t r2 = st_mix(r ^ 0x1234567ull);
        switch ((r >> 40) & 3u) {
            case 0: a = (double)(int)(r2 & 0xffffu) - (double)__uint_as_float((uint32_t)(r2 >> 16) & 0x4fffffffu); break;
            case 1: {
                float x = __uint_as_float((uint32_t)r2), y = __uint_as_float((uint32_t)(r2 >> 32));
                if (!(fabsf(x) <= FLT_MAX)) x = 1.0f;
                if (!(fabsf(y) <= FLT_MAX)) y = 2.0f;
                a = (double)x - (double)y;
                break;
            }
            case 2: {
                a = __longlong_as_double((long long)((r2 & 0x800fffffffffffffull) | ((uint64_t)(1023 - 200 + (r2 >> 52) % 400u) << 52)));
                break;
            }
            default: {
                const double m = __longlong_as_double((long long)((r2 & 0x000fffffffffffffull) | ((uint64_t)(1023 - 60 + (r2 >> 52) % 120u) << 52)));
                a = m * bd;
                const long long ab = __double_as_longlong(a);
                a = __longlong_as_double(ab + (long long)((r2 >> 60) % 3u) - 1);
                if (!(fabs(a) <= 1e300)) a = m;
                break;
            }
        }
        const double y = rcp_f32range(bd);
        const double q = ddiv_rcp(a, bd, y);
        const double want = __ddiv_rn(a, bd);
        if (__double_as_longlong(q) != __double_as_longlong(want)) ++bad;
        double y0;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(bd));
        const double e0 = fabs(fma(-bd, y0, 1.0));
        if (e0 > worst) worst = e0;
    }
    if (bad) atomicAdd(mismatches, bad);
    atomicMax(seed_err_bits, (unsigned long long)__double_as_longlong(worst));     // non-negative doubles order like integers
}

cudaError_t launch_selftest_division(int nnum, uint64_t seed, unsigned long long* dev2, cudaStream_t st) {
    selftest_division_kernel<<<(1u << 24) / 256, 256, 0, st>>>(nnum, seed, dev2, dev2 + 1);
    return cudaGetLastError();
}

}  // namespace imgcorr
