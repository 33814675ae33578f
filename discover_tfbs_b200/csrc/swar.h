// swar.h — SIMD-within-a-register ASCII -> 2-bit code conversion, 4 bases per 32-bit word.
// Pure integer code shared by the encode kernel (device) and tfbs_debug_swar4 (host, used by the
// CPU-side unit test so the bit tricks are checked exhaustively without a GPU).
//
// ASCII facts used (after folding case, i.e. ignoring bit 5):
//   A = 0x41 = 0100 0001   C = 0x43 = 0100 0011   G = 0x47 = 0100 0111   T = 0x54 = 0101 0100
//   c' = bits 2..1 of the byte:  A=0  C=1  T=2  G=3      code = c' ^ (c' >> 1):  A=0 C=1 G=2 T=3
//   with bits 5, 2, 1 cleared (mask 0xD9) A, C, G collapse to 0x41 and T to 0x50 = 0x41 ^ 0x11...
//   ...but so do 'E' (0x45, c' = 2) and P/R/V (0x50/0x52/0x56, c' = 0/1/3).  Hence
//   valid  <=>  (byte & 0xD9) == 0x41 + 15 * [c' == 2].
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

// w holds 4 ASCII bytes (byte k = base k).  Returns the per-byte 2-bit codes (bits 0..1 of every
// byte, other bits 0; NOT yet zeroed for invalid bytes); *diff gets a non-zero byte exactly where
// the input byte is not one of ACGTacgt.  6 ALU ops + 1 multiply-add per 4 bases.
TFBS_HD uint32_t tfbs_swar_codes(uint32_t w, uint32_t *diff)
{
    const uint32_t s1 = w >> 1, s2 = w >> 2;
    const uint32_t is_t = s2 & ~s1 & 0x01010101u;           // [c' == 2] per byte
    const uint32_t expect = is_t * 15u + 0x41414141u;       // 0x41 or 0x50 per byte (no carries)
    *diff = (w & 0xD9D9D9D9u) ^ expect;
    const uint32_t c = s1 & 0x03030303u;
    return c ^ (s2 & 0x01010101u);
}

// Gather the four 2-bit fields at bits 0,8,16,24 into the TOP byte (bits 24..31, base k at
// bits 24+2k); the low 24 bits are garbage.
TFBS_HD uint32_t tfbs_gather_top(uint32_t codes) { return codes * 0x01041040u; }

// Reference composition used by the slow path and the host test: the 4 codes packed into bits
// 0..7 (base k at bits 2k, invalid bases forced to code 0); *valid4 gets bit k set iff base k is
// one of ACGTacgt.
TFBS_HD uint32_t tfbs_swar4(uint32_t w, uint32_t *valid4)
{
    uint32_t diff;
    uint32_t c = tfbs_swar_codes(w, &diff);
    const uint32_t vb = tfbs_zero_bytes(diff) >> 7;  // 0x01 per valid byte
    c &= vb * 3u;
    *valid4 = (vb * 0x01020408u) >> 24;              // gather bits 0,8,16,24 -> bits 0..3
    return tfbs_gather_top(c) >> 24;
}
