// kb_common.cuh — record layout, key mixing and bit-field helpers shared by all kernels.
//
// A record is one occurrence of a k-mer: the packed equivalent of one text line "left,mid,right"
// of the reference's k-mer table (krisp_fasta/Amplicon.py:298-348) plus the file it came from.
// It is a bit string, MSB first:
//
//   [ left : 2L bits ][ right : 2R bits ][ mid : 2D bits ][ zero pad ][ file id : 8 bits ]
//   \______ flank key, FB = 2(L+R) bits ___/
//
// Codes A=0 C=1 G=2 T=3 keep the reference's LC_ALL=C order (kstream/kstream.py:108) and make
// complement = code ^ 3.
//
// Two storage modes:
//   DIRECT   (2k + 8 <= 64, e.g. 25/1/2 -> 54+2+8 = 64 bits): the record is ONE 64-bit word and is
//            itself the sort element.  The flank key is stored *mixed* by a bijection on FB bits
//            (kb_mix) so that radix digits are uniform whatever the genome composition, and so
//            that sorting on a prefix of the key already separates almost all groups.
//   INDIRECT (longer k-mers, e.g. 32/60/32 -> 128+120+8 = 256 bits = 4 words): records are written
//            once, in extraction order, as W 64-bit words; the sort element is the 64-bit entry
//            [ hash32(flank) : 32 ][ record index : 32 ] and the group pass gathers the records.
// In both modes the radix sort orders elements by the top 8*P bits only; the group pass (kb_group.cuh)
// compares complete flank keys, so grouping is exact whatever P is.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define KB_MAX_W 8          // 64-bit words per record (k <= 252)
#define KB_MAX_FILES 256    // 8-bit file id
#define KB_IDBITS 8
#define KB_MAX_MW 16        // 32-bit mask words per side (D <= 128)

struct KbLayout {
    int L, D, R, k;
    int FB;        // flank bits = 2(L+R)
    int FW;        // 64-bit words holding flank bits (>= 1)
    int W;         // 64-bit words per record (1 in DIRECT mode)
    int direct;    // 1 = DIRECT, 0 = INDIRECT
    int mix;       // DIRECT only: 1 = flank key stored mixed (search), 0 = plain (sorted tables)
    int P;         // radix passes (8-bit digits over the top 8P bits of the sort element)
    int cmpbits;   // run boundary = change in the top cmpbits of the sort element
    int n_files;   // global number of input files
    int PW;        // presence words (32-bit) = ceil(n_files/32)
    int MW;        // mask words (32-bit) per side = ceil(D/8); column c = nibble (7 - c%8) of word c/8
    uint32_t shs;  // kb_mix xorshift distance
};

// ---- key mixing (bijection on nb-bit keys) -------------------------------------------------
#define KB_MIX_C1 0x9E3779B97F4A7C15ULL
#define KB_MIX_C2 0xD6E8FEB86659FD93ULL

__host__ __device__ constexpr uint64_t kb_inv64(uint64_t a) {   // inverse of odd a modulo 2^64 (Newton)
    uint64_t x = a;
    for (int i = 0; i < 6; i++) x *= 2 - a * x;
    return x;
}
__host__ __device__ __forceinline__ uint64_t kb_lowmask(int nb) { return nb >= 64 ? ~0ULL : ((1ULL << nb) - 1ULL); }

// Multiplicative mixing mod 2^nb (nb in 1..64): a bijection whose TOP bits depend on every input bit — exactly what the
// partition digits (top bits) and the hash-table slot (the bits below them) need; inverted only for survivors.
// (An earlier two-round multiply / xor-shift version cost K1 13 more instructions per record for no measurable gain in balance.)
__host__ __device__ __forceinline__ uint64_t kb_mix(uint64_t x, int nb, uint32_t /*s*/) {
    return (x * KB_MIX_C1) & kb_lowmask(nb);
}
__host__ __device__ __forceinline__ uint64_t kb_unmix(uint64_t x, int nb, uint32_t /*s*/) {
    return (x * kb_inv64(KB_MIX_C1)) & kb_lowmask(nb);
}

__host__ __device__ __forceinline__ uint64_t kb_mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

// Hash of the flank key (left, right) read straight from a 2-bit base stream, in 64-bit chunks of each field: used by the lazy
// multi-word path (K1 computes it without building the record, kb_materialize_kernel recomputes it from the sequence bytes).
// get(bit position, n bits <= 64) returns the bits right-aligned.  The TOP bits are the partition digits / table slots.
template <class G>
__device__ __forceinline__ uint64_t kb_flank_hash(G get, uint32_t lpos, uint32_t rpos, uint32_t L2, uint32_t R2) {
    uint64_t h = 0x243F6A8885A308D3ULL;
    for (uint32_t d = 0; d < L2; d += 64) { h = (h ^ get(lpos + d, min(64u, L2 - d))) * KB_MIX_C1; h ^= h >> 32; }
    for (uint32_t d = 0; d < R2; d += 64) { h = (h ^ get(rpos + d, min(64u, R2 - d))) * KB_MIX_C1; h ^= h >> 32; }
    return h * KB_MIX_C2;
}

// ---- MSB-first bit strings -------------------------------------------------------------------
// nbits (1..64) starting at bit `pos` of the word array `s` (word pos/64+1 must be readable).
__device__ __forceinline__ uint64_t kb_get_bits(const uint64_t* s, uint32_t pos, uint32_t nbits) {
    uint32_t w = pos >> 6, o = pos & 63;
    uint64_t v = s[w] << o;
    if (o) v |= s[w + 1] >> (64 - o);
    return v >> (64 - nbits);
}

// same on a register-resident record of WN words (unrolled selects, no local memory)
template <int WN>
__device__ __forceinline__ uint64_t kb_rec_bits(const uint64_t (&rec)[WN], uint32_t pos, uint32_t nbits) {
    uint32_t w = pos >> 6, o = pos & 63;
    uint64_t a = 0, b = 0;
#pragma unroll
    for (int j = 0; j < WN; j++) {
        if (j == (int)w) a = rec[j];
        if (j == (int)w + 1) b = rec[j];
    }
    uint64_t v = a << o;
    if (o) v |= b >> (64 - o);
    return v >> (64 - nbits);
}

// OR the low `nbits` of v into the record so that its MSB lands at bit `pos`.
template <int WN>
__device__ __forceinline__ void kb_put_bits(uint64_t (&rec)[WN], uint32_t pos, uint32_t nbits, uint64_t v) {
    uint32_t w = pos >> 6, o = pos & 63;
    uint64_t va = v << (64 - nbits);
    uint64_t hi = va >> o;
    uint64_t lo = o ? (va << (64 - o)) : 0ULL;
#pragma unroll
    for (int j = 0; j < WN; j++) {
        if (j == (int)w) rec[j] |= hi;
        if (j == (int)w + 1) rec[j] |= lo;
    }
}

// copy `n` bits of stream `s` starting at bit `src` into the record at bit `dst`
template <int WN>
__device__ __forceinline__ void kb_copy_bits(uint64_t (&rec)[WN], uint32_t dst, const uint64_t* s, uint32_t src, uint32_t n) {
    while (n > 0) {
        uint32_t c = n < 64 ? n : 64;
        kb_put_bits<WN>(rec, dst, c, kb_get_bits(s, src, c));
        dst += c; src += c; n -= c;
    }
}

// 8 two-bit codes (16 bits, first column in the top bits) -> 8 one-hot nibbles (first column in the top nibble)
__host__ __device__ __forceinline__ uint32_t kb_onehot8(uint32_t x) {
    uint32_t y = x & 0xFFFFu;
    y = (y | (y << 8)) & 0x00FF00FFu;
    y = (y | (y << 4)) & 0x0F0F0F0Fu;
    y = (y | (y << 2)) & 0x33333333u;
    uint32_t lo = y & 0x11111111u, hi = (y >> 1) & 0x11111111u;
    return (~(hi | lo) & 0x11111111u) | ((lo & ~hi) << 1) | ((hi & ~lo) << 2) | ((hi & lo) << 3);
}

// reverse complement of 32 packed bases (MSB-first 2-bit codes)
__device__ __forceinline__ uint64_t kb_rc64(uint64_t x) {
    uint64_t y = __brevll(~x);
    return ((y >> 1) & 0x5555555555555555ULL) | ((y & 0x5555555555555555ULL) << 1);
}

// ---- global memory access ----------------------------------------------------------------------
// streaming (read-once) loads: bypass L1 allocation
__device__ __forceinline__ uint64_t kb_ld_stream(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 kb_ld_stream128(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void kb_st_stream128(void* p, uint64_t a, uint64_t b) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void kb_st_stream64(void* p, uint64_t a) {
    asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" :: "l"(p), "l"(a) : "memory");
}

__device__ __forceinline__ uint32_t kb_lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
