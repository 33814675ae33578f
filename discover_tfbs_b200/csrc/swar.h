// swar.h — SIMD-within-a-register ASCII -> 2-bit code conversion, 4 bases per 32-bit word.
// Pure integer code shared by the encode kernel (device) and tfbs_debug_swar4 (host, used by the
// CPU-side unit test so the bit tricks are checked exhaustively without a GPU).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define TFBS_HD __host__ __device__ __forceinline__
#else
#define TFBS_HD static inline
#endif

// Per-byte "is zero" flags (0x80 in every byte of z that is 0x00); exact, no cross-byte borrow.
TFBS_HD uint32_t tfbs_zero_bytes(uint32_t z)
{
    uint32_t t = (z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | z) & 0x80808080u;
}

// w holds 4 ASCII bytes (byte k = base k).  Returns the 4 2-bit codes packed into bits 0..7
// (base k at bits 2k, invalid bases forced to code 0); *valid4 gets bit k set iff base k is one
// of ACGTacgt.
TFBS_HD uint32_t tfbs_swar4(uint32_t w, uint32_t *valid4)
{
    const uint32_t u = w & 0xDFDFDFDFu;  // fold lower case onto upper case
    // A=0x41 C=0x43 G=0x47 share (u & 0xF9) == 0x41; the fourth byte with that property, 'E'
    // (0x45: bit2 set, bit1 clear), is excluded by OR-ing in a non-zero marker.
    const uint32_t acg = ((u & 0xF9F9F9F9u) ^ 0x41414141u) | (((u >> 1) & ~u) & 0x02020202u);
    const uint32_t t = u ^ 0x54545454u;  // 'T'
    const uint32_t v = tfbs_zero_bytes(acg) | tfbs_zero_bytes(t);  // 0x80 per valid byte
    const uint32_t vb = v >> 7;                                    // 0x01 per valid byte
    // (u >> 1) & 3 gives A=0 C=1 T=2 G=3; x ^ (x >> 1) swaps the last two: A=0 C=1 G=2 T=3.
    uint32_t c = (u >> 1) & 0x03030303u;
    c = (c ^ ((c >> 1) & 0x01010101u)) & (vb * 3u);
    *valid4 = (vb * 0x01020408u) >> 24;  // gather bits 0,8,16,24 -> bits 0..3
    return (c * 0x01041040u) >> 24;      // gather 2-bit fields at bits 0,8,16,24 -> bits 0..7
}
