// wn_rng.cu -- the reference's Gaussian fill on the GPU (kernel K0 of DESIGN.md).
//
// Reference: WaveletNoise.cpp:74-77 / :146-147 draw n^d variates from std::normal_distribution<float>(0,1) over
// std::mt19937(seed).  With libstdc++ that is (bits/random.tcc:1811-1843, :3346-3382):
//   u = float(mt()) / 2^32 (nextafter(1,0) if it rounds to 1);  x = 2u-1, y = 2u'-1;  r2 = x*x + y*y;
//   reject while r2 > 1 or r2 == 0;  m = sqrt(-2*log(r2)/r2);  return y*m, then x*m on the next call.
// The sequence looks serial but is not: attempts are aligned to even draw indices, every accept/reject decision
// is independent, and the output slot of an accepted attempt is 2 * (number of accepted attempts before it).
//   k_mt19937      one CTA extends the MT recurrence 224 elements per step in a circular shared-memory buffer
//                  while a second thread group tempers and streams the previous step's words to global memory;
//   k_polar_count  accept flags per attempt -> per-block accept counts;
//   k_scan_blocks  exclusive scan of the block counts (single CTA);
//   k_polar_emit   re-evaluates the flags, ranks them (ballot/popc + block offset) and writes y*m, x*m.
// Every float operation is the un-fused IEEE operation the host performs, and logf is a restatement of glibc's
// logf (sysdeps/ieee754/flt-32/e_logf.c, the 16-entry table algorithm from ARM's optimized routines) in double
// arithmetic -- checked bit-for-bit against glibc 2.39 on all 2.5e8 floats in [2^-30, 1] (see tests) -- so the
// field, and therefore the tile, is BIT-IDENTICAL to the reference's host fill.
#include "wn_internal.h"

namespace {

constexpr int MT_N = 624, MT_M = 397;

__device__ __forceinline__ uint32_t mt_twist(uint32_t a, uint32_t b, uint32_t m)
{
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// draws[0 .. nblocks*624): the first nblocks*624 outputs of std::mt19937(seed).
// The generator is the linear recurrence x[p] = x[p-227] ^ twist(x[p-624], x[p-623]) over the seed expansion
// x[0..623], output q = temper(x[624+q]).  The shortest dependency is 227 elements, so 224 consecutive elements
// (7 warps) are independent: one step = 224 threads extend the sequence in a 2048-word circular buffer in shared
// memory while 224 other threads temper and store the elements of the PREVIOUS step, one barrier per step.  The chain
// of n^3 * 1.3 / 224 dependent steps is what bounds the kernel (latency, one CTA), not bandwidth.
constexpr int MT_STEP = 224, MT_RING = 2048;

__global__ void __launch_bounds__(512) k_mt19937(uint32_t seed, uint32_t *__restrict__ draws, int nblocks)
{
    __shared__ uint32_t x[MT_RING];
    const int t = threadIdx.x;
    if (t == 0) {
        uint32_t v = seed;
        x[0] = v;
        for (int i = 1; i < MT_N; ++i) { v = 1812433253u * (v ^ (v >> 30)) + (uint32_t)i; x[i] = v; }
    }
    __syncthreads();
    const long long end = (long long)MT_N * nblocks + MT_N;     // one past the last sequence element needed
    const bool producer = t < MT_STEP;
    const int u = t - 256;                                      // consumer lane 0..223 (threads 256..479)
    for (long long p0 = MT_N; p0 < end + MT_STEP; p0 += MT_STEP) {
        if (producer) {
            const long long p = p0 + t;
            if (p < end) {
                const unsigned i = (unsigned)p;
                x[i & (MT_RING - 1)] = mt_twist(x[(i - MT_N) & (MT_RING - 1)], x[(i - MT_N + 1) & (MT_RING - 1)],
                                                x[(i - (unsigned)(MT_N - MT_M)) & (MT_RING - 1)]);
            }
        } else if (u >= 0 && u < MT_STEP) {
            const long long q = p0 - MT_STEP + u;               // written by the previous step
            if (q >= MT_N && q < end) draws[q - MT_N] = mt_temper(x[(unsigned)q & (MT_RING - 1)]);
        }
        __syncthreads();
    }
}

// ---- glibc logf restated (normal positive inputs; r2 is never subnormal, zero or negative here) -------------
__constant__ double c_invc[16] = {
    0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0, 0x1.3c995b0b80385p+0, 0x1.30d190c8864a5p+0,
    0x1.25e227b0b8eap+0, 0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0, 0x1.0953f419900a7p+0, 0x1p+0,
    0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1, 0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1,
    0x1.767dcf5534862p-1 };
__constant__ double c_logc[16] = {
    -0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3, -0x1.6574f0ac07758p-3,
    -0x1.1aa2bc79c81p-3, -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4, -0x1.252f438e10c1ep-5, 0x0p+0,
    0x1.aa5aa5df25984p-5, 0x1.c5e53aa362eb4p-4, 0x1.526e57720db08p-3, 0x1.bc2860d22477p-3, 0x1.1058bc8a07ee1p-2,
    0x1.4043057b6ee09p-2 };

__device__ __forceinline__ float glibc_logf(float x)
{
    const uint32_t ix = __float_as_uint(x);
    if (ix == 0x3f800000u) return 0.0f;
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const int k = (int)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double z = (double)__uint_as_float(iz);
    const double r = __dsub_rn(__dmul_rn(z, c_invc[i]), 1.0);
    const double y0 = __dadd_rn(c_logc[i], __dmul_rn((double)k, 0x1.62e42fefa39efp-1));
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(0x1.5575b0be00b6ap-2, r), -0x1.ffffef20a4123p-2);
    y = __dadd_rn(__dmul_rn(-0x1.00ea348b88334p-2, r2), y);
    y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
    return __double2float_rn(y);
}

// generate_canonical<float,24>: float(draw) * 2^-32, clamped below 1
__device__ __forceinline__ float canonical(uint32_t d)
{
    float u = __fmul_rn(__uint2float_rn(d), 2.3283064365386963e-10f);
    if (u >= 1.0f) u = 0.99999994f;                         // nextafterf(1, 0)
    return u;
}

__device__ __forceinline__ bool attempt(const uint32_t *__restrict__ draws, size_t a, float &x, float &y, float &r2)
{
    const uint2 d = *reinterpret_cast<const uint2 *>(draws + 2 * a);
    x = __fsub_rn(__fmul_rn(2.0f, canonical(d.x)), 1.0f);
    y = __fsub_rn(__fmul_rn(2.0f, canonical(d.y)), 1.0f);
    r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));       // un-fused: the decision depends on every bit
    return !(r2 > 1.0f || r2 == 0.0f);
}

constexpr int PB = 1024;                                    // attempts per block

__global__ void __launch_bounds__(PB) k_polar_count(const uint32_t *__restrict__ draws, size_t attempts,
                                                    uint32_t *__restrict__ block_count)
{
    const size_t a = (size_t)blockIdx.x * PB + threadIdx.x;
    float x, y, r2;
    const bool ok = a < attempts && attempt(draws, a, x, y, r2);
    const int c = __syncthreads_count(ok);
    if (threadIdx.x == 0) block_count[blockIdx.x] = (uint32_t)c;
}

// exclusive scan of block_count[0..nb) in place; total[0] = sum (64-bit)
__global__ void __launch_bounds__(1024) k_scan_blocks(uint32_t *__restrict__ block_count, int nb,
                                                      unsigned long long *__restrict__ total)
{
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < nb ? block_count[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += n;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = warp_sum[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long n = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += n;
            }
            warp_sum[lane] = winc - w;                       // exclusive prefix of the warp sums
        }
        __syncthreads();
        const unsigned long long excl = carry + warp_sum[warp] + inc - v;
        if (i < nb) block_count[i] = (uint32_t)excl;         // fits: attempts < 2^32
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = carry;
}

__global__ void __launch_bounds__(PB) k_polar_emit(const uint32_t *__restrict__ draws, size_t attempts,
                                                   const uint32_t *__restrict__ block_offset, float *__restrict__ out,
                                                   size_t count, unsigned long long *__restrict__ info)
{
    __shared__ int warp_cnt[PB / 32];
    const size_t a = (size_t)blockIdx.x * PB + threadIdx.x;
    float x = 0.0f, y = 0.0f, r2 = 1.0f;
    const bool ok = a < attempts && attempt(draws, a, x, y, r2);
    const unsigned ballot = __ballot_sync(0xffffffffu, ok);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(ballot);
    __syncthreads();
    if (!ok) return;
    int before = __popc(ballot & ((1u << lane) - 1u));
    for (int w = 0; w < warp; ++w) before += warp_cnt[w];
    const size_t slot = 2 * ((size_t)block_offset[blockIdx.x] + (size_t)before);
    if (slot >= count) return;
    if (slot == ((count - 1) >> 1) * 2) info[1] = (unsigned long long)a;      // the attempt that produced the last output
    // mult = sqrt(-2*log(r2)/r2)
    const float mult = __fsqrt_rn(__fdiv_rn(__fmul_rn(-2.0f, glibc_logf(r2)), r2));
    out[slot] = __fadd_rn(__fmul_rn(__fmul_rn(y, mult), 1.0f), 0.0f);          // ret*stddev + mean
    if (slot + 1 < count) out[slot + 1] = __fadd_rn(__fmul_rn(__fmul_rn(x, mult), 1.0f), 0.0f);
}

} // namespace

// Fills out[0..count).  `accepted` (device, 2 x 8 bytes): [0] = number of accepted attempts (the fill is complete iff
// 2*accepted >= count), [1] = index of the attempt that produced the last output, i.e. the generator has consumed
// 2*([1]+1) raw draws.  margin_permille widens the number of attempts generated.
int wn_launch_gaussian_fill(unsigned seed, float *out, size_t count, unsigned long long *accepted, int margin_permille,
                            cudaStream_t st)
{
    if (count == 0) return 0;
    // expected attempts = (count/2) / (pi/4); generate margin on top, rounded up to whole 624-word blocks
    const double need = (double)((count + 1) / 2) * 1.2732395447351628;
    size_t attempts = (size_t)(need * (1.0 + margin_permille / 1000.0)) + 4096;
    const size_t nblocks = (2 * attempts + MT_N - 1) / MT_N;
    if (nblocks > 0x7fffffff) return -1;
    attempts = nblocks * MT_N / 2;
    if (attempts >= 0xffffffffull) return -1;
    uint32_t *draws = nullptr, *bc = nullptr;
    const size_t nb = (attempts + PB - 1) / PB;
    if (wn_scratch_alloc((void **)&draws, nblocks * MT_N * sizeof(uint32_t), st) != cudaSuccess) return -1;
    if (wn_scratch_alloc((void **)&bc, nb * sizeof(uint32_t), st) != cudaSuccess) { cudaFreeAsync(draws, st); return -1; }
    k_mt19937<<<1, 512, 0, st>>>(seed, draws, (int)nblocks);
    k_polar_count<<<(unsigned)nb, PB, 0, st>>>(draws, attempts, bc);
    k_scan_blocks<<<1, 1024, 0, st>>>(bc, (int)nb, accepted);
    k_polar_emit<<<(unsigned)nb, PB, 0, st>>>(draws, attempts, bc, out, count, accepted);
    cudaFreeAsync(bc, st);
    cudaFreeAsync(draws, st);
    return 4;
}
