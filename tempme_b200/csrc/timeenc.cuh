// TimeEncode cosine (reference models/explainer.py:55-58) for sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tmb {

// cos(x) for the TimeEncode arguments (they reach 1e8 and beyond, where the library cosf takes its slow path).
// Branch-free exact argument reduction in integer arithmetic: |x| = m * 2^e with a 24-bit integer m, so
// frac(|x| / 2pi) = frac(m * frac(2^e / 2pi)); kInv2Pi[e + 44] holds frac(2^e / 2pi) in 0.64 fixed point for every
// finite fp32 exponent (e = -44 .. 104; below that the angle is 0), of which the top 32 bits of the product are kept
// (error < 2^-32 turn = 1.5e-9 rad).  The turn fraction is shifted by a quarter turn (cos -> sin) and folded into [-pi/2, pi/2] for one
// odd polynomial.  Max error ~2 ulp of 1.0 (< 2e-7) against the exact cosine of the fp32 argument (cosf: 1-2 ulp); inf and nan give nan.  The table is read from shared memory (a copy of kInv2Pi): lanes index it with
// different exponents.
constexpr int kInv2PiN = 149;
__constant__ unsigned long long kInv2Pi[kInv2PiN] = {
    0x0000000000028be6ull, 0x00000000000517ccull, 0x00000000000a2f98ull, 0x0000000000145f30ull, 0x000000000028be60ull, 0x0000000000517cc1ull,
    0x0000000000a2f983ull, 0x000000000145f306ull, 0x00000000028be60dull, 0x000000000517cc1bull, 0x000000000a2f9836ull, 0x00000000145f306dull,
    0x0000000028be60dbull, 0x00000000517cc1b7ull, 0x00000000a2f9836eull, 0x0000000145f306dcull, 0x000000028be60db9ull, 0x0000000517cc1b72ull,
    0x0000000a2f9836e4ull, 0x000000145f306dc9ull, 0x00000028be60db93ull, 0x000000517cc1b727ull, 0x000000a2f9836e4eull, 0x00000145f306dc9cull,
    0x0000028be60db939ull, 0x00000517cc1b7272ull, 0x00000a2f9836e4e4ull, 0x0000145f306dc9c8ull, 0x000028be60db9391ull, 0x0000517cc1b72722ull,
    0x0000a2f9836e4e44ull, 0x000145f306dc9c88ull, 0x00028be60db93910ull, 0x000517cc1b727220ull, 0x000a2f9836e4e441ull, 0x00145f306dc9c882ull,
    0x0028be60db939105ull, 0x00517cc1b727220aull, 0x00a2f9836e4e4415ull, 0x0145f306dc9c882aull, 0x028be60db9391054ull, 0x0517cc1b727220a9ull,
    0x0a2f9836e4e44152ull, 0x145f306dc9c882a5ull, 0x28be60db9391054aull, 0x517cc1b727220a94ull, 0xa2f9836e4e441529ull, 0x45f306dc9c882a53ull,
    0x8be60db9391054a7ull, 0x17cc1b727220a94full, 0x2f9836e4e441529full, 0x5f306dc9c882a53full, 0xbe60db9391054a7full, 0x7cc1b727220a94feull,
    0xf9836e4e441529fcull, 0xf306dc9c882a53f8ull, 0xe60db9391054a7f0ull, 0xcc1b727220a94fe1ull, 0x9836e4e441529fc2ull, 0x306dc9c882a53f84ull,
    0x60db9391054a7f09ull, 0xc1b727220a94fe13ull, 0x836e4e441529fc27ull, 0x06dc9c882a53f84eull, 0x0db9391054a7f09dull, 0x1b727220a94fe13aull,
    0x36e4e441529fc275ull, 0x6dc9c882a53f84eaull, 0xdb9391054a7f09d5ull, 0xb727220a94fe13abull, 0x6e4e441529fc2757ull, 0xdc9c882a53f84eafull,
    0xb9391054a7f09d5full, 0x727220a94fe13abeull, 0xe4e441529fc2757dull, 0xc9c882a53f84eafaull, 0x9391054a7f09d5f4ull, 0x27220a94fe13abe8ull,
    0x4e441529fc2757d1ull, 0x9c882a53f84eafa3ull, 0x391054a7f09d5f47ull, 0x7220a94fe13abe8full, 0xe441529fc2757d1full, 0xc882a53f84eafa3eull,
    0x91054a7f09d5f47dull, 0x220a94fe13abe8faull, 0x441529fc2757d1f5ull, 0x882a53f84eafa3eaull, 0x1054a7f09d5f47d4ull, 0x20a94fe13abe8fa9ull,
    0x41529fc2757d1f53ull, 0x82a53f84eafa3ea6ull, 0x054a7f09d5f47d4dull, 0x0a94fe13abe8fa9aull, 0x1529fc2757d1f534ull, 0x2a53f84eafa3ea69ull,
    0x54a7f09d5f47d4d3ull, 0xa94fe13abe8fa9a6ull, 0x529fc2757d1f534dull, 0xa53f84eafa3ea69bull, 0x4a7f09d5f47d4d37ull, 0x94fe13abe8fa9a6eull,
    0x29fc2757d1f534ddull, 0x53f84eafa3ea69bbull, 0xa7f09d5f47d4d377ull, 0x4fe13abe8fa9a6eeull, 0x9fc2757d1f534ddcull, 0x3f84eafa3ea69bb8ull,
    0x7f09d5f47d4d3770ull, 0xfe13abe8fa9a6ee0ull, 0xfc2757d1f534ddc0ull, 0xf84eafa3ea69bb81ull, 0xf09d5f47d4d37703ull, 0xe13abe8fa9a6ee06ull,
    0xc2757d1f534ddc0dull, 0x84eafa3ea69bb81bull, 0x09d5f47d4d377036ull, 0x13abe8fa9a6ee06dull, 0x2757d1f534ddc0dbull, 0x4eafa3ea69bb81b6ull,
    0x9d5f47d4d377036dull, 0x3abe8fa9a6ee06dbull, 0x757d1f534ddc0db6ull, 0xeafa3ea69bb81b6cull, 0xd5f47d4d377036d8ull, 0xabe8fa9a6ee06db1ull,
    0x57d1f534ddc0db62ull, 0xafa3ea69bb81b6c5ull, 0x5f47d4d377036d8aull, 0xbe8fa9a6ee06db14ull, 0x7d1f534ddc0db629ull, 0xfa3ea69bb81b6c52ull,
    0xf47d4d377036d8a5ull, 0xe8fa9a6ee06db14aull, 0xd1f534ddc0db6295ull, 0xa3ea69bb81b6c52bull, 0x47d4d377036d8a56ull, 0x8fa9a6ee06db14acull,
    0x1f534ddc0db62959ull, 0x3ea69bb81b6c52b3ull, 0x7d4d377036d8a566ull, 0xfa9a6ee06db14accull, 0xf534ddc0db629599ull, 0xea69bb81b6c52b32ull,
    0xd4d377036d8a5664ull, 0xa9a6ee06db14acc9ull, 0x534ddc0db6295993ull, 0xa69bb81b6c52b327ull, 0x4d377036d8a5664full
};

__device__ __forceinline__ float cos_accurate(float x, const uint2 *tab) {
    const uint32_t bits = __float_as_uint(x) & 0x7fffffffu, ex = bits >> 23;              // |x| = m * 2^(ex - 150)
    const uint32_t m = ex < 106u ? 0u : ((bits & 0x7fffffu) | 0x800000u);                  // tiny |x|: angle 0
    const uint2 T = tab[min(max((int)ex - 106, 0), kInv2PiN - 1)];                          // {low, high} words of frac(2^e / 2pi)
    // cos(2 pi t) = sin(2 pi (t + 1/4)): turn fraction + a quarter turn, 0.32 fixed point, read as a signed angle in [-1/2, 1/2) turn
    const uint32_t s = m * T.y + __umulhi(m, T.x) + (1u << 30);
    // fold into [-1/4, 1/4] turn (sin(pi - a) = sin(a)): the two top bits differ exactly when |angle| >= 1/4 turn
    const int y = (int)(s ^ (s << 1)) < 0 ? (int)(0x80000000u - s) : (int)s;
    const float th = (float)y * 1.46291807926715968e-9f /* 2 pi / 2^32 */, z = th * th;
    // odd minimax polynomial of sin on [-pi/2, pi/2] (degree 11, fit error 2e-11; fp32 evaluation error < 1.2e-7)
    const float p = fmaf(z, fmaf(z, fmaf(z, fmaf(z, -2.3846693508744465e-8f, 2.752261934801936e-6f), -1.9840804452542216e-4f), 8.333330042660236e-3f), -1.666666716337204e-1f);
    const float v = fmaf(th * z, p, th);
    return ex == 255u ? __int_as_float(0x7fffffff) : v;
}

// copies the table into shared memory (kInv2PiN uint2 entries); call with all threads of the CTA, then synchronise
__device__ __forceinline__ void cos_table_to_smem(uint2 *ctab) {
    for (int i = threadIdx.x; i < kInv2PiN; i += blockDim.x) ctab[i] = make_uint2((uint32_t)kInv2Pi[i], (uint32_t)(kInv2Pi[i] >> 32));
}

}  // namespace tmb
