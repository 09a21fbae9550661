// Device helpers shared by the per-direction aggregation kernels (sgbm_kernels.cu) and the fused
// wavefront sweeps (sweep_kernels.cu): streaming 16-byte accesses and one step of the A.4 recurrence.
#pragma once
#include "sgbm.cuh"

namespace wsg {

static constexpr unsigned FULL = 0xffffffffu;
static constexpr unsigned SAT2 = 0x7FFF7FFFu;

__device__ __forceinline__ uint4 ldg_stream(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_rw(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v)
{
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

template <int NL>
__device__ __forceinline__ unsigned group_min_u32(unsigned v)
{
    if constexpr (NL == 32) {
        return __reduce_min_sync(FULL, v);
    } else {
#pragma unroll
        for (int o = NL / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o, NL));
        return v;
    }
}

// consecutive pairs (0,1)(2,3)(4,5)(6,7) -> the interleaved vector (0,4)(1,5)(2,6)(3,7) of the C/S volumes (vec_pos)
__device__ __forceinline__ uint4 interleave8(unsigned p01, unsigned p23, unsigned p45, unsigned p67)
{
    return make_uint4(__byte_perm(p01, p45, 0x5410), __byte_perm(p01, p45, 0x7632), __byte_perm(p23, p67, 0x5410),
                      __byte_perm(p23, p67, 0x7632));
}

// One step of the recurrence for one direction.  R: normalised state of the predecessor (in/out),
// Cw: cost of this pixel, v: L of this pixel (out).  NR packed registers, 2 disparities each.
// a*one + b with `one` == 1 at run time: an integer add the compiler has to issue on the FMA pipe (IMAD), which is
// idle next to the integer ALU that bounds these kernels.
__device__ __forceinline__ unsigned add_on_fma(unsigned a, unsigned b, unsigned one)
{
    unsigned r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
}
// min(a + b, 32767) per 16-bit half for a, b <= 32767: the add on the FMA pipe, the full-rate two-input minimum on the ALU
__device__ __forceinline__ unsigned sat_add_split(unsigned a, unsigned b, unsigned one)
{
    return __vminu2(add_on_fma(a, b, one), SAT2);
}

// FASTL: max(C) + P2 <= 32767 is known (the domain in which cv2 is reproduced at all), so t + C cannot leave
// int16 and the saturating add becomes a plain add on the FMA pipe.
template <int NL, int NR, bool HASPAD, bool FASTL = false>
__device__ __forceinline__ void agg_step(unsigned (&R)[NR], const unsigned (&Cw)[NR], unsigned (&v)[NR], int l,
                                         unsigned P1p, unsigned P2mP1p, const unsigned* padm, unsigned one = 1u)
{
    // registers 4g..4g+3 of group g hold (a+8g+i, a+8g+4+i), i = 0..3.  d-1 of register 4g+i is register 4g+i-1, except
    // for i == 0: (a+8g-1, a+8g+3) = (high half of the previous group's last register, low half of this group's last);
    // d+1 of register 4g+i is register 4g+i+1, except for i == 3: (a+8g+4, a+8g+8) = (high half of this group's first
    // register, low half of the next group's first).  Previous / next group of the lane's first / last group live in the
    // neighbouring lane; beyond d = -1 and d = D the value is 32767.
    constexpr int NG = NR / 4;
    unsigned up = __shfl_up_sync(FULL, R[NR - 1], 1, NL);
    unsigned dn = __shfl_down_sync(FULL, R[0], 1, NL);
    if (l == 0) up = SAT2;
    if (l == NL - 1) dn = SAT2;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const unsigned prev3 = g == 0 ? up : R[4 * g - 1];
        const unsigned next0 = g == NG - 1 ? dn : R[4 * g + 4];
        const unsigned qlo = __byte_perm(prev3, R[4 * g + 3], 0x5432);      // (prev3.hi, own3.lo)
        const unsigned qhi = __byte_perm(R[4 * g], next0, 0x5432);          // (own0.hi, next0.lo)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int j = 4 * g + i;
            const unsigned below = i == 0 ? qlo : R[j - 1];
            const unsigned above = i == 3 ? qhi : R[j + 1];
            unsigned t = __vimin3_s16x2(below, above, P2mP1p);
            t = __viaddmin_s16x2(t, P1p, R[j]);
            v[j] = FASTL ? add_on_fma(t, Cw[j], one) : __viaddmin_u16x2(t, Cw[j], SAT2);
            if (HASPAD) v[j] |= padm[j / 4];
        }
    }
    unsigned m = v[0];
#pragma unroll
    for (int j = 1; j < NR; ++j) m = __vmins2(m, v[j]);
    unsigned mpk;
    if constexpr (NL == 32) {
        // both halves := min of the two halves; values are 0..32767, so the warp minimum of the packed words as unsigned
        // 32-bit numbers is the packed minimum (one REDUX, no unpack / repack)
        m = __vmins2(m, __byte_perm(m, m, 0x1032));
        mpk = __reduce_min_sync(FULL, m);
    } else {
        unsigned mm = min(m & 0xFFFFu, m >> 16);
        mm = group_min_u32<NL>(mm);
        mpk = mm * 0x10001u;
    }
    const unsigned minus_one = 0u - one;     // opaque like `one`: the subtraction is issued as IMAD on the FMA pipe
#pragma unroll
    for (int j = 0; j < NR; ++j) R[j] = add_on_fma(mpk, v[j], minus_one);  // v - mpk; both halves >= min: no borrow between them
}

}  // namespace wsg
