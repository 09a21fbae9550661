// Cost volume of cv::StereoSGBM (SURVEY.md Appendix A.2 + A.3; call site src/wass_stereo/wass_stereo.cpp:837),
// wide-tile form: one CTA owns XT (64) columns x 32 disparities and marches down a band of rows.
//
// Three stages run CONCURRENTLY on different warps of the CTA, one image row apart, with one barrier per row:
//   T   (10 warps) unpack the prefilter records of row r+2 into broadcast tables (img1 side) and a reversed,
//                  pre-negated table (img2 side) in shared memory; the global load is issued before P1, used after it
//   P1  (same)     Birchfield-Tomasi pixel cost of row r+1 for XT+2*SW2 columns x 32 disparities, two disparities
//                  per 32-bit register (VIADD.16x2 / VIADDMNMX.S16x2.RELU / VIMNMX.S16x2); also the sums of column pairs
//   P2  (4 warps)  row r: horizontal box sum (SW2 column pairs + one column, then sliding to the next column), a ring of
//                  2*SH2+1 row sums in shared memory for the vertical sliding sum, 16-byte stores of C
// The kernel is bound by the shared-memory pipe and by its longest stage (P2), so the layouts are chosen for zero bank
// conflicts (ncu: 1.43 -> 0.78 k wavefronts per row step): swizzled unpadded columns (wcol), chunk flip (pd_off), a P1 lane
// mapping whose two pair-parity table copies never meet in a bank, 32-byte uu records stored in two conflict-free halves.
// History at the benchmark size (2448x2048, D=256, window 13): 32x64 tile 4.21 ms -> this form 3.77 -> 2.49 ms.
// Shared memory grows with the window: windows above 17 use cost_kernel (sgbm_kernels.cu).
#include "sgbm_dev.cuh"
#include <cstdlib>

namespace wsg {

static constexpr int WDT = 32;             // disparities per CTA
static constexpr int WDTP = 32;            // u16 per column in shared memory (64 B, no padding: columns are swizzled, see wcol)
static constexpr int WRB = 256;            // rows per band
static constexpr int WMAXSW = 8;           // windows up to 17
static constexpr int WRVPAD = 40;           // puts the second table copy half a bank line (16 words) away from the first

// XT = output columns per CTA: 128 (one CTA per SM) or 64 (two CTAs per SM at the reference's window)
template <int XT> struct WideCfg {
    static constexpr int NCOL = XT + 2 * WMAXSW;
    static constexpr int P1 = (NCOL * 4 + 31) / 32 * 32;   // stage P1 + T threads: NCOL columns x 4 groups of 8 disparities
    static constexpr int CPT = XT == 64 ? 2 : 4;           // adjacent columns per P2 thread (sliding horizontal sum)
    static constexpr int GB = XT / CPT * 4;                // stage P2 threads: XT/CPT column groups x 4 groups of 8 disparities
    static constexpr int CT = P1 + GB;
    static constexpr int VT = (NCOL + WDT + 4) / 2 * 2;    // entries of the reversed img2 tables
    static constexpr int NTAB = 6;                         // img2 tables per set: V, Vlo, -Vhi for the two channels (-V is made in the ALU)
    static constexpr int RV = 2 * NTAB * VT + WRVPAD;      // s16 per img2 table set (two copies, one element apart)
    static constexpr int CTAS = XT == 64 ? 2 : 1;
    // the two parity copies are read by one LDS (even / odd lanes of a 24-entry span): their bank ranges must not overlap
    static_assert(((NTAB * VT + WRVPAD) / 2) % 32 >= 12 && ((NTAB * VT + WRVPAD) / 2) % 32 <= 20, "table copies share banks");
};

// Physical column of logical column c in the pd / ring arrays.  A quarter-warp of a 16-byte access covers two columns
// (P1: c, c+1; P2: c, c+CPT): flipping bit 0 with bit log2(CPT) puts those two in different 64-byte halves of the 128-byte
// bank line, so every LDS.128 / STS.128 of the kernel is conflict-free (the 80-byte padded layout was 2-way conflicting
// for CPT = 2: ncu showed the shared-memory pipe 90 % busy, the kernel's limiter).
template <int CPT> __device__ __forceinline__ int wcol(int c) { return c ^ ((c >> (CPT == 2 ? 1 : 2)) & 1); }

// u16 offset of the 8-disparity chunk g of logical column c inside one pd row buffer.  The chunk index is flipped with
// bit 1 of the column so that a P1 quarter-warp (4 columns x 2 chunks, see the lane mapping in stage P1) also covers all
// eight 16-byte slots of a bank line.
template <int CPT> __device__ __forceinline__ int pd_off(int c, int g)
{
    return wcol<CPT>(c) * WDTP + ((g ^ (c & 2)) << 3);
}

struct WideSmem { int pd, pp, uu, rv, ring, total; };
template <int XT> __host__ __device__ inline WideSmem wide_layout(int SH2)
{
    using Cfg = WideCfg<XT>;
    WideSmem s;
    s.pd = 0;                                          // u16 [2][NCOL][WDTP]
    s.pp = s.pd + 2 * Cfg::NCOL * WDTP * 2;            // u16 [2][NCOL/2][WDT]: pd[2j] + pd[2j+1] (used when CPT == 2)
    s.uu = s.pp + 2 * (Cfg::NCOL / 2) * WDT * 2;       // u32 [2][NCOL][8]
    s.rv = s.uu + 2 * Cfg::NCOL * 8 * 4;               // s16 [2][RV]
    s.ring = (s.rv + 2 * Cfg::RV * 2 + 127) & ~127;    // u16 [2*SH2+1][XT][WDTP], bank-line aligned
    s.total = s.ring + (2 * SH2 + 1) * XT * WDTP * 2;
    return s;
}

bool cost_wide_supported(const SgbmPlan& p) { return p.SW2 <= WMAXSW && p.SH2 <= WMAXSW && p.Dp % WDT == 0; }

template <int XT>
__global__ void __launch_bounds__(WideCfg<XT>::CT, WideCfg<XT>::CTAS) cost_wide_kernel(const uint2* __restrict__ pre1, const uint2* __restrict__ pre2,
                                                           int16_t* __restrict__ C, int* __restrict__ maxC, SgbmPlan p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    using Cfg = WideCfg<XT>;
    constexpr int WXT = XT, WNCOL = Cfg::NCOL, WP1 = Cfg::P1, WVT = Cfg::VT, WRV = Cfg::RV;
    const WideSmem L = wide_layout<XT>(p.SH2);
    uint16_t* pd = reinterpret_cast<uint16_t*>(smem + L.pd);
    uint16_t* pp = reinterpret_cast<uint16_t*>(smem + L.pp);
    unsigned* uu = reinterpret_cast<unsigned*>(smem + L.uu);
    int16_t* rv = reinterpret_cast<int16_t*>(smem + L.rv);
    uint16_t* ring = reinterpret_cast<uint16_t*>(smem + L.ring);

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * WXT;          // first output column (W1 space)
    const int d0 = blockIdx.y * WDT;          // first disparity slot (relative to minD)
    const int y0 = blockIdx.z * WRB;
    const int y1 = min(y0 + WRB, p.H);
    const int ncol = WXT + 2 * p.SW2;
    const int NR = 2 * p.SH2 + 1;
    const int win = 2 * p.SW2 + 1;
    const int xb = p.minX1 + min(max(x0 + WXT - 1 + p.SW2, 0), p.W1 - 1);   // image x of the last (clamped) halo column
    const int vtop = xb - (p.minD + d0);      // largest img2 column touched; table index i <-> x' = vtop - i
    const int nsteps = (y1 - y0) + 2 * p.SH2;

    // stage P2 role (group B): columns CPT*cg..+CPT-1, disparities d0+8g..+7
    constexpr int CPT = Cfg::CPT;
    const int tb = tid - WP1;
    const int g = tb & 3, cg = tb >> 2;
    const bool real_vec = (d0 + 8 * g) < p.D;
    unsigned acc[CPT][4];
#pragma unroll
    for (int i = 0; i < CPT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0;
    int vmax = 0;
    int rslot = 0;                            // ring slot of this row-step (idx % NR)
    // P2 addressing, constant over the rows: the swizzle of pd repeats every PER columns, so column c0+i sits at
    // poff[i % PER] + (i / PER) * PER * WDTP -- with the window loop unrolled the second term is an immediate.
    constexpr int PER = CPT == 2 ? 4 : 8;
    const int c0 = cg * CPT;
    int poff[PER], roff[CPT];
#pragma unroll
    for (int j = 0; j < PER; ++j) poff[j] = pd_off<CPT>(c0 + j, g);
#pragma unroll
    for (int j = 0; j < CPT; ++j) roff[j] = wcol<CPT>(c0 + j) * WDTP + g * 8;
    const int plast = pd_off<CPT>(c0 + win - 1, g), pnext = pd_off<CPT>(c0 + win, g);   // CPT == 2: last column of the window, and the next
    // first output row of this band, this thread's first column and disparity vector
    int16_t* dst0 = C + ((size_t)y0 * p.W1 + (x0 + cg * CPT)) * p.Dp + vec_slot((d0 >> 3) + g, p.NL, p.K) * 8;

    // P1 addressing, constant over the rows (see the lane mapping at stage P1)
    const int p1cc = (tid >> 6) * 16 + ((tid & 31) >> 1), p1gg = (tid & 1) + ((tid >> 4) & 2);
    const bool p1on = tid < WP1 && p1cc < ncol, p1real = d0 + 8 * p1gg < p.D;
    const int p1i0 = (xb - (p.minX1 + min(max(x0 - p.SW2 + p1cc, 0), p.W1 - 1))) + 8 * p1gg;
    const int p1rv = (p1i0 & 1) * (Cfg::NTAB * WVT + WRVPAD) + (p1i0 & ~1);      // s16 offset into one img2 table set
    const int p1pd = pd_off<Cfg::CPT>(p1cc, p1gg);

    // T addressing: threads [0,ncol) fetch an img1 column, [ncol, ncol+WVT) an img2 table entry
    const bool tOn = tid < ncol + WVT, isU = tid < ncol;
    const int te = isU ? tid : tid - ncol;
    const uint2* tsrc = isU ? pre1 + (p.minX1 + min(max(x0 - p.SW2 + te, 0), p.W1 - 1)) : pre2 + min(max(vtop - te, 0), p.W - 1);

    for (int s = 0; s < nsteps + 2; ++s) {
        if (tid < WP1) {
            // ---------------- T (first half): fetch this thread's prefilter record of row-step s; the latency hides
            //                  behind P1 below.  Threads [0,ncol): img1 column; [ncol, ncol+WVT): img2 table entry.
            const bool doT = s < nsteps && tOn;
            uint2 q = make_uint2(0, 0);
            if (doT) q = tsrc[(size_t)min(max(y0 - p.SH2 + s, 0), p.H - 1) * p.W];
            // ---------------- P1: pixel cost of row-step s-1 from tables[(s-1)&1] into pd[(s-1)&1]
            const int rs = s - 1;
            // lane mapping: a warp covers 16 columns x 2 disparity groups (warps 2k / 2k+1 of a 16-column block take
            // groups {0,1} / {2,3}).  The 16 lanes of one pair parity then read only 12 distinct table words (the two
            // groups overlap), so the two parity copies never meet in a bank; with 8 columns x 4 groups every LDS of
            // the img2 tables was a 2-way conflict (32 distinct words, ranges shifting by one with the tile's parity).
            if (rs >= 0 && rs < nsteps) {
                const int b = rs & 1;
                const int cc = p1cc;
                uint4 out = make_uint4(0, 0, 0, 0);
                if (p1on && p1real) {
                    const unsigned* rvw = reinterpret_cast<const unsigned*>(rv + b * WRV + p1rv);
                    const uint4 ua = *reinterpret_cast<const uint4*>(uu + (b * WNCOL + cc) * 8);
                    const uint4 ub = *reinterpret_cast<const uint4*>(uu + (b * WNCOL + cc) * 8 + 4);
                    unsigned res[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // six table words per disparity pair; -V costs two ALU instructions (~V + 1 per half), a seventh and
                        // eighth table would cost two more shared-memory wavefronts per warp -- and that pipe bounds the kernel
                        const unsigned V0 = rvw[0 * (WVT / 2) + k], Vl0 = rvw[1 * (WVT / 2) + k], nVh0 = rvw[2 * (WVT / 2) + k];
                        const unsigned V1 = rvw[3 * (WVT / 2) + k], Vl1 = rvw[4 * (WVT / 2) + k], nVh1 = rvw[5 * (WVT / 2) + k];
                        const unsigned nV0 = __vadd2(~V0, 0x00010001u), nV1 = __vadd2(~V1, 0x00010001u);
                        // per channel: c0 = max(0,u-vhi,vlo-u), c1 = max(0,v-uhi,ulo-v), c = min(c0,c1)
                        const unsigned e0 = __vimax_s16x2_relu(__vadd2(ua.x, nVh0), __vadd2(Vl0, ua.y));
                        const unsigned e1 = __vimax_s16x2_relu(__vadd2(V0, ua.w), __vadd2(ua.z, nV0));
                        const unsigned ca = __vmins2(e0, e1);
                        const unsigned f0 = __vimax_s16x2_relu(__vadd2(ub.x, nVh1), __vadd2(Vl1, ub.y));
                        const unsigned f1 = __vimax_s16x2_relu(__vadd2(V1, ub.w), __vadd2(ub.z, nV1));
                        const unsigned cb = __vmins2(f0, f1);
                        res[k] = ca + ((cb >> 2) & 0x3FFF3FFFu);
                    }
                    out = make_uint4(res[0], res[1], res[2], res[3]);
                }
                if (p1on) *reinterpret_cast<uint4*>(pd + b * WNCOL * WDTP + p1pd) = out;
                if (Cfg::CPT == 2) {
                    // sums of column pairs (2j, 2j+1): the neighbour column is two lanes up; halves the window loads of
                    // the box-sum warps, which are the longest stage of a row step
                    const uint4 nb = make_uint4(__shfl_down_sync(FULL, out.x, 2), __shfl_down_sync(FULL, out.y, 2),
                                                __shfl_down_sync(FULL, out.z, 2), __shfl_down_sync(FULL, out.w, 2));
                    if (p1on && !(cc & 1))
                        *reinterpret_cast<uint4*>(pp + (b * (WNCOL / 2) + (cc >> 1)) * WDT + p1gg * 8) =
                            make_uint4(__vadd2(out.x, nb.x), __vadd2(out.y, nb.y), __vadd2(out.z, nb.z), __vadd2(out.w, nb.w));
                }
            }
            // ---------------- T (second half): unpack into tables[s&1]
            if (doT) {
                const int b = s & 1;
                const int v0 = q.x & 255, l0 = (q.x >> 8) & 255, h0 = (q.x >> 16) & 255;
                const int v1 = q.x >> 24, l1 = q.y & 255, h1 = (q.y >> 8) & 255;
                if (isU) {
                    auto bc = [](int v) -> unsigned { return ((unsigned)v & 0xFFFFu) * 0x10001u; };
                    uint4* o = reinterpret_cast<uint4*>(uu + (b * WNCOL + te) * 8);
                    const uint4 c0v = make_uint4(bc(v0), bc(-v0), bc(l0), bc(-h0)), c1v = make_uint4(bc(v1), bc(-v1), bc(l1), bc(-h1));
                    const int first = (te >> 2) & 1;          // 32-byte records: lanes 4 apart would meet in a bank
                    o[first] = first ? c1v : c0v;
                    o[first ^ 1] = first ? c0v : c1v;
                } else {
                    const int16_t val[Cfg::NTAB] = {(int16_t)v0, (int16_t)l0, (int16_t)-h0, (int16_t)v1, (int16_t)l1, (int16_t)-h1};
                    int16_t* t = rv + b * WRV;
#pragma unroll
                    for (int qn = 0; qn < Cfg::NTAB; ++qn) {
                        t[(0 * Cfg::NTAB + qn) * WVT + te] = val[qn];                              // copy A: rv[i]
                        if (te > 0) t[WRVPAD + (1 * Cfg::NTAB + qn) * WVT + te - 1] = val[qn];     // copy B: rv[i+1]
                    }
                }
            }
        } else {
            // ---------------- P2: box sums of row-step s-2 from pd[(s-2)&1]
            const int idx = s - 2;
            if (idx >= 0) {
                const int b = idx & 1;
                unsigned hs[4] = {0, 0, 0, 0};
                uint4 head[CPT > 1 ? CPT - 1 : 1];
                const uint16_t* prow = pd + b * WNCOL * WDTP;
                if (CPT == 2) {
                    // window of column c0 (even): SW2 column pairs + the single column c0 + win - 1
                    const uint16_t* qrow = pp + (b * (WNCOL / 2) + (c0 >> 1)) * WDT + g * 8;
#pragma unroll
                    for (int m = 0; m < WMAXSW; ++m) {
                        if (m < p.SW2) {
                            const uint4 v = *reinterpret_cast<const uint4*>(qrow + m * WDT);
                            hs[0] = __vadd2(hs[0], v.x); hs[1] = __vadd2(hs[1], v.y);
                            hs[2] = __vadd2(hs[2], v.z); hs[3] = __vadd2(hs[3], v.w);
                        }
                    }
                    const uint4 v = *reinterpret_cast<const uint4*>(prow + plast);
                    hs[0] = __vadd2(hs[0], v.x); hs[1] = __vadd2(hs[1], v.y);
                    hs[2] = __vadd2(hs[2], v.z); hs[3] = __vadd2(hs[3], v.w);
                    head[0] = *reinterpret_cast<const uint4*>(prow + poff[0]);
                } else {
#pragma unroll
                    for (int i = 0; i < CPT - 1; ++i) {   // the columns that leave the window while sliding
                        head[i] = *reinterpret_cast<const uint4*>(prow + poff[i % PER] + (i / PER) * PER * WDTP);
                        if (i < win) {
                            hs[0] = __vadd2(hs[0], head[i].x); hs[1] = __vadd2(hs[1], head[i].y);
                            hs[2] = __vadd2(hs[2], head[i].z); hs[3] = __vadd2(hs[3], head[i].w);
                        }
                    }
#pragma unroll
                    for (int i = CPT - 1; i < 2 * WMAXSW + 1; ++i) {
                        if (i < win) {
                            const uint4 v = *reinterpret_cast<const uint4*>(prow + poff[i % PER] + (i / PER) * PER * WDTP);
                            hs[0] = __vadd2(hs[0], v.x); hs[1] = __vadd2(hs[1], v.y);
                            hs[2] = __vadd2(hs[2], v.z); hs[3] = __vadd2(hs[3], v.w);
                        }
                    }
                }
                const bool store = idx >= 2 * p.SH2;
                uint16_t* slot = ring + rslot * WXT * WDTP;
#pragma unroll
                for (int cc = 0; cc < CPT; ++cc) {
                    if (cc > 0) {
                        const uint4 vn = *reinterpret_cast<const uint4*>(prow + (CPT == 2 ? pnext : pd_off<CPT>(c0 + win - 1 + cc, g)));
                        const uint4 vo = head[cc - 1];
                        hs[0] = __vsub2(__vadd2(hs[0], vn.x), vo.x); hs[1] = __vsub2(__vadd2(hs[1], vn.y), vo.y);
                        hs[2] = __vsub2(__vadd2(hs[2], vn.z), vo.z); hs[3] = __vsub2(__vadd2(hs[3], vn.w), vo.w);
                    }
                    unsigned* ac = acc[cc];
                    if (idx >= NR) {
                        const uint4 o = *reinterpret_cast<const uint4*>(slot + roff[cc]);
                        ac[0] = __vsub2(ac[0], o.x); ac[1] = __vsub2(ac[1], o.y);
                        ac[2] = __vsub2(ac[2], o.z); ac[3] = __vsub2(ac[3], o.w);
                    }
                    *reinterpret_cast<uint4*>(slot + roff[cc]) = make_uint4(hs[0], hs[1], hs[2], hs[3]);
                    ac[0] = __vadd2(ac[0], hs[0]); ac[1] = __vadd2(ac[1], hs[1]);
                    ac[2] = __vadd2(ac[2], hs[2]); ac[3] = __vadd2(ac[3], hs[3]);
                    if (store && x0 + cg * CPT + cc < p.W1) {
                        uint4* dst = reinterpret_cast<uint4*>(dst0 + (size_t)cc * p.Dp);
                        if (real_vec) {
                            *dst = interleave8(ac[0], ac[1], ac[2], ac[3]);
                            const unsigned m = __vmaxs2(__vmaxs2(ac[0], ac[1]), __vmaxs2(ac[2], ac[3]));
                            vmax = max(vmax, max((int)(short)(m & 0xFFFF), (int)(short)(m >> 16)));
                        } else {
                            *dst = make_uint4(0, 0, 0, 0);
                        }
                    }
                }
                rslot = rslot + 1 == NR ? 0 : rslot + 1;
                if (store) dst0 += (size_t)p.W1 * p.Dp;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(FULL, vmax, o));
    if ((tid & 31) == 0 && vmax > 0) atomicMax(maxC, vmax);
}

template <int XT>
static void launch_cost_wide_t(const uint2* pre1, const uint2* pre2, int16_t* C, int* maxC, const SgbmPlan& p, cudaStream_t st)
{
    const int smem = wide_layout<XT>(p.SH2).total;
    cudaFuncSetAttribute(cost_wide_kernel<XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(cost_wide_kernel<XT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    dim3 g((p.W1 + XT - 1) / XT, p.Dp / WDT, (p.H + WRB - 1) / WRB);
    cost_wide_kernel<XT><<<g, WideCfg<XT>::CT, smem, st>>>(pre1, pre2, C, maxC, p);
}

void launch_cost_wide(const uint2* pre1, const uint2* pre2, int16_t* C, int* maxC, const SgbmPlan& p, cudaStream_t st)
{
    static int xt = -1;
    if (xt < 0) { const char* e = getenv("WSG_COST_XT"); xt = e ? atoi(e) : 64; }
    if (xt == 128) launch_cost_wide_t<128>(pre1, pre2, C, maxC, p, st);
    else           launch_cost_wide_t<64>(pre1, pre2, C, maxC, p, st);
}

}  // namespace wsg
