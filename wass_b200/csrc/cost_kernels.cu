// Cost volume of cv::StereoSGBM (SURVEY.md Appendix A.2 + A.3; call site src/wass_stereo/wass_stereo.cpp:837),
// wide-tile form: one CTA owns 64 columns x 32 disparities and marches down a band of rows.
//
// Three stages run CONCURRENTLY on different warps of the CTA, one image row apart, with one barrier per row:
//   T   (2 warps)  unpack the prefilter records of row r+2 into broadcast records (img1 side) and a reversed table
//                  (img2 side, two copies one element apart) in shared memory; the records of row r+3 are in flight meanwhile
//   P1  (5 warps)  Birchfield-Tomasi pixel cost of row r+1 for 64+2*SW2 columns x 32 disparities, two disparities per
//                  32-bit register (VIADD.16x2 / VIADDMNMX.S16x2.RELU / VIMNMX.S16x2).  One thread = columns c and c+2 x
//                  8 disparities: the two see the img2 table two entries (one word) apart, so five table words per table
//                  serve both instead of eight; the neighbouring lane owns c+1 and c+3 (the other parity copy of the
//                  table), and the sums of the column pairs (c,c+1), (c+2,c+3) that stage P2 reads take one exchange
//   P2  (4 warps)  row r: horizontal box sum (SW2 column pairs + one column, then sliding to the next column), a ring of
//                  2*SH2+1 row sums in shared memory for the vertical sliding sum, 16-byte stores of C
// The kernel is bound by the shared-memory pipe (wavefronts) and by instruction issue together, so (a) every layout is
// chosen for zero bank conflicts -- columns swizzled by wcol / ppcol so that the two columns of any quarter-warp access
// (always four apart) sit in different halves of a bank line -- and (b) the table loads, the largest single consumer
// of both, are shared between columns as described above.
// History at the benchmark size (2448x2048, D=256, window 13): 32x64 tile 4.21 ms -> wide tile, one column x 8
// disparities per P1 thread 3.77 -> 2.49 -> 2.32 ms -> this form (see DESIGN.md for the current figure).
// Shared memory grows with the window: windows above 17 use cost_kernel (sgbm_kernels.cu).
#include "sgbm_dev.cuh"
#include <cstdlib>

namespace wsg {

static constexpr int WXT = 64;             // output columns per CTA
static constexpr int WDT = 32;             // disparities per CTA
static constexpr int WDTP = 32;            // u16 per column in shared memory (64 B, no padding: columns are swizzled, see wcol)
static constexpr int WRB = 256;            // rows per band
static constexpr int WMAXSW = 8;           // windows up to 17
static constexpr int WNCOL = WXT + 2 * WMAXSW;     // columns of pixel costs per row (output columns + halo)
static constexpr int WQN = 26;             // 16-byte slots between the per-column sub-arrays of the img1 records (= 2 mod 8)
static constexpr int WVT = WNCOL + WDT;            // entries of the reversed img2 tables (largest index used: ncol + 30)
static constexpr int WNTAB = 6;            // img2 tables: V, Vlo, -Vhi for the two channels (-V is made in the ALU)
static constexpr int WBOFF = WNTAB * WVT + 10;     // s16 offset of copy B (entry i at i-1: odd-index pairs word-aligned); an ODD number of words
static constexpr int WRV = WBOFF + WNTAB * WVT;    // s16 per table set
static constexpr int WUSZ = 2 * 4 * WQN * 4;       // u32 per img1 record set: [half][column & 3][column >> 2][4]
static constexpr int WP1T = 160, WTT = 64, WP2T = 128;
static constexpr int WCT = WP1T + WTT + WP2T;
static_assert(WNCOL / 4 <= WQN && WQN % 8 == 2 && WBOFF % 4 == 2 && WNCOL / 4 * 8 <= WP1T && 3 * WTT >= WNCOL + WVT, "cost tile");

// Physical column of logical column c in the pd / ring arrays, and of column pair j in pp.  A quarter-warp of a 16-byte
// access covers two columns (P1: 4q+j and 4q+4+j; P2: c and c+4, see its lane mapping) or two pairs (two apart):
// flipping bit 0 with the bit that tells those two apart puts them in different 64-byte halves of the 128-byte bank
// line, so every LDS.128 / STS.128 of the kernel is conflict-free.
__device__ __forceinline__ int wcol(int c) { return c ^ ((c >> 2) & 1); }
__device__ __forceinline__ int ppcol(int j) { return j ^ ((j >> 1) & 1); }
__device__ __forceinline__ int pd_off(int c, int g) { return wcol(c) * WDTP + (g << 3); }
// s16 offset of the word holding table entries (i, i+1): copy A for even i, copy B for odd i
__device__ __forceinline__ int rv_off(int i) { return (i & 1) * WBOFF + (i & ~1); }

struct WideSmem { int pd, pp, uu, rv, ring, total; };
__host__ __device__ inline WideSmem wide_layout(int SH2)
{
    WideSmem s;
    s.pd = 0;                                          // u16 [2][NCOL][WDTP]
    s.pp = s.pd + 2 * WNCOL * WDTP * 2;                // u16 [2][NCOL/2][WDT]: pd[2j] + pd[2j+1]
    s.uu = s.pp + 2 * (WNCOL / 2) * WDT * 2;           // u32 [2][WUSZ]
    s.rv = s.uu + 2 * WUSZ * 4;                        // s16 [2][WRV]
    s.ring = (s.rv + 2 * WRV * 2 + 127) & ~127;        // u16 [2*SH2+1][WXT][WDTP], bank-line aligned
    s.total = s.ring + (2 * SH2 + 1) * WXT * WDTP * 2;
    return s;
}

bool cost_wide_supported(const SgbmPlan& p) { return p.SW2 <= WMAXSW && p.SH2 <= WMAXSW && p.Dp % WDT == 0; }

// Birchfield-Tomasi cost of one column against 8 disparities (4 packed words): T[t][k + OFF] are the img2 table words,
// ua / ub the img1 record of the column: (u, -u, ulo + 1, -uhi) per channel, each broadcast to both halves.
//   per channel: c0 = max(0, u - vhi, vlo - u), c1 = max(0, v - uhi, ulo - v), c = min(c0, c1);  ulo - v = (ulo + 1) + ~v
template <int OFF>
__device__ __forceinline__ uint4 bt_cost4(const unsigned* __restrict__ urec, const unsigned (&T)[WNTAB][5])
{
    const uint4 ua = *reinterpret_cast<const uint4*>(urec);
    const uint4 ub = *reinterpret_cast<const uint4*>(urec + 4 * WQN * 4);
    unsigned res[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned V0 = T[0][k + OFF], Vl0 = T[1][k + OFF], nVh0 = T[2][k + OFF];
        const unsigned V1 = T[3][k + OFF], Vl1 = T[4][k + OFF], nVh1 = T[5][k + OFF];
        const unsigned e0 = __vimax_s16x2_relu(__vadd2(ua.x, nVh0), __vadd2(Vl0, ua.y));
        const unsigned e1 = __vimax_s16x2_relu(__vadd2(V0, ua.w), __vadd2(ua.z, ~V0));
        const unsigned ca = __vmins2(e0, e1);
        const unsigned f0 = __vimax_s16x2_relu(__vadd2(ub.x, nVh1), __vadd2(Vl1, ub.y));
        const unsigned f1 = __vimax_s16x2_relu(__vadd2(V1, ub.w), __vadd2(ub.z, ~V1));
        const unsigned cb = __vmins2(f0, f1);
        res[k] = ca + ((cb >> 2) & 0x3FFF3FFFu);
    }
    return make_uint4(res[0], res[1], res[2], res[3]);
}

template <int NW>
__device__ __forceinline__ void load_tables(const unsigned* __restrict__ w, unsigned (&T)[WNTAB][5])
{
#pragma unroll
    for (int t = 0; t < WNTAB; ++t)
#pragma unroll
        for (int k = 0; k < NW; ++k) T[t][k] = w[t * (WVT / 2) + k];
}

__device__ __forceinline__ uint4 vadd2x4(const uint4& a, const uint4& b)
{
    return make_uint4(__vadd2(a.x, b.x), __vadd2(a.y, b.y), __vadd2(a.z, b.z), __vadd2(a.w, b.w));
}

__global__ void __launch_bounds__(WCT, 2) cost_wide_kernel(const uint2* __restrict__ pre1, const uint2* __restrict__ pre2,
                                                           int16_t* __restrict__ C, int* __restrict__ maxC, SgbmPlan p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const WideSmem L = wide_layout(p.SH2);
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
    int vmax = 0;

    if (tid < WP1T) {
        // ================= P1: pixel cost of row-step s-1 from tables[(s-1)&1] into pd[(s-1)&1], pp[(s-1)&1]
        // thread = (quad q of columns, e, g): columns 4q+e ("H") and 4q+e+2 ("L"), disparities 8g..8g+7; lane = e + 2g + 8(q&3).
        // One table LDS reads words W - 2q + 4g from one parity copy for the even lanes and from the other for the odd
        // lanes; the copies sit an odd number of words apart, so even word offsets of the two never share a bank.
        const int e = tid & 1, g = (tid >> 1) & 3, q = tid >> 3, cb = 4 * q;
        const int xu = x0 - p.SW2 + cb;                        // W1-space column of the quad's first column, unclamped
        const bool on = cb < ncol, real = d0 + 8 * g < p.D;
        // regular quad: no column clamped to the image -> table index of column cb+j = I0 - j
        const bool regular = xu >= 0 && xu + 3 <= p.W1 - 1 && cb + 3 < ncol;
        const int I = xb - (p.minX1 + xu) - e + 8 * g;         // table index of column H (regular quads)
        const int rvw = rv_off(max(I - 2, 0)) >> 1;            // word offset of entries I-2 .. I+7: H = words 1..4, L = words 0..3
        const int pdH = pd_off(cb + e, g), pdL = pd_off(cb + e + 2, g);
        const int ppo = ppcol(2 * q + e) * WDT + g * 8;        // pair (4q, 4q+1) for e = 0, (4q+2, 4q+3) for e = 1
        const int uH = (e * WQN + q) * 4, uL = ((e + 2) * WQN + q) * 4;
        for (int s = 0; s < nsteps + 2; ++s) {
            const int rs = s - 1;
            if (rs >= 0 && rs < nsteps) {
                const int b = rs & 1;
                uint4 oH = make_uint4(0, 0, 0, 0), oL = make_uint4(0, 0, 0, 0);
                if (on && real) {
                    const unsigned* tw = reinterpret_cast<const unsigned*>(rv + b * WRV);
                    const unsigned* ur = uu + b * WUSZ;
                    unsigned T[WNTAB][5];
                    if (regular) {
                        load_tables<5>(tw + rvw, T);
                        oH = bt_cost4<1>(ur + uH, T);
                        oL = bt_cost4<0>(ur + uL, T);
                    } else {
                        // image border (columns replicate the first / last one) or the partial last quad: column by column
                        if (cb + e < ncol) {
                            const int i0 = xb - (p.minX1 + min(max(xu + e, 0), p.W1 - 1)) + 8 * g;
                            load_tables<4>(tw + (rv_off(i0) >> 1), T);
                            oH = bt_cost4<0>(ur + uH, T);
                        }
                        if (cb + e + 2 < ncol) {
                            const int i0 = xb - (p.minX1 + min(max(xu + e + 2, 0), p.W1 - 1)) + 8 * g;
                            load_tables<4>(tw + (rv_off(i0) >> 1), T);
                            oL = bt_cost4<0>(ur + uL, T);
                        }
                    }
                }
                uint16_t* prow = pd + b * WNCOL * WDTP;
                if (on) {
                    *reinterpret_cast<uint4*>(prow + pdH) = oH;
                    *reinterpret_cast<uint4*>(prow + pdL) = oL;
                }
                // sums of the column pairs (4q, 4q+1), (4q+2, 4q+3): the partner lane (e ^ 1) holds the other column of each
                const uint4 snd = e ? oH : oL, mine = e ? oL : oH;
                const uint4 rcv = make_uint4(__shfl_xor_sync(FULL, snd.x, 1), __shfl_xor_sync(FULL, snd.y, 1),
                                             __shfl_xor_sync(FULL, snd.z, 1), __shfl_xor_sync(FULL, snd.w, 1));
                if (on) *reinterpret_cast<uint4*>(pp + b * (WNCOL / 2) * WDT + ppo) = vadd2x4(mine, rcv);
            }
            __syncthreads();
        }
    } else if (tid < WP1T + WTT) {
        // ================= T: prefilter records of row-step s into tables[s&1].  Records [0,ncol): img1 columns (clamped to
        // the image); [ncol, ncol+WVT): img2 table entries.  Three records per thread, the loads issued together.
        const int tt = tid - WP1T;
        const uint2* tsrc[3];
        int te[3];
        bool isU[3], onr[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int r = tt + k * WTT;
            isU[k] = r < ncol;
            onr[k] = r < ncol + WVT;
            te[k] = isU[k] ? r : r - ncol;
            tsrc[k] = isU[k] ? pre1 + (p.minX1 + min(max(x0 - p.SW2 + te[k], 0), p.W1 - 1)) : pre2 + min(max(vtop - te[k], 0), p.W - 1);
        }
        // the records of row-step s+1 are fetched while those of step s are unpacked: a global load's latency is longer than
        // a whole row step of the other two stages, and everybody waits for this one at the barrier
        uint2 nxt[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) nxt[k] = onr[k] ? tsrc[k][(size_t)min(max(y0 - p.SH2, 0), p.H - 1) * p.W] : make_uint2(0, 0);
        for (int s = 0; s < nsteps + 2; ++s) {
            if (s < nsteps) {
                const int b = s & 1;
                uint2 rec[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) rec[k] = nxt[k];
                if (s + 1 < nsteps) {
                    const size_t row = (size_t)min(max(y0 - p.SH2 + s + 1, 0), p.H - 1) * p.W;
#pragma unroll
                    for (int k = 0; k < 3; ++k) nxt[k] = onr[k] ? tsrc[k][row] : make_uint2(0, 0);
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (!onr[k]) continue;
                    const uint2 q = rec[k];
                    const int v0 = q.x & 255, l0 = (q.x >> 8) & 255, h0 = (q.x >> 16) & 255;
                    const int v1 = q.x >> 24, l1 = q.y & 255, h1 = (q.y >> 8) & 255;
                    if (isU[k]) {
                        auto bc = [](int v) -> unsigned { return ((unsigned)v & 0xFFFFu) * 0x10001u; };
                        unsigned* o = uu + b * WUSZ + ((te[k] & 3) * WQN + (te[k] >> 2)) * 4;
                        *reinterpret_cast<uint4*>(o) = make_uint4(bc(v0), bc(-v0), bc(l0 + 1), bc(-h0));
                        *reinterpret_cast<uint4*>(o + 4 * WQN * 4) = make_uint4(bc(v1), bc(-v1), bc(l1 + 1), bc(-h1));
                    } else {
                        const int16_t val[WNTAB] = {(int16_t)v0, (int16_t)l0, (int16_t)-h0, (int16_t)v1, (int16_t)l1, (int16_t)-h1};
                        int16_t* t = rv + b * WRV;
#pragma unroll
                        for (int qn = 0; qn < WNTAB; ++qn) {
                            t[qn * WVT + te[k]] = val[qn];                                   // copy A: entry i at i
                            if (te[k] > 0) t[WBOFF + qn * WVT + te[k] - 1] = val[qn];        // copy B: entry i at i-1
                        }
                    }
                }
            }
            __syncthreads();
        }
    } else {
        // ================= P2: box sums of row-step s-2 from pd[(s-2)&1], pp[(s-2)&1]: columns c0, c0+1, disparities d0+8g..+7.
        // Lane mapping: the two column groups of a quarter-warp are 2 apart (columns 4 apart), see wcol.
        const int tb = tid - WP1T - WTT;
        const int g = tb & 3, cgp = tb >> 2;
        const int cg = (cgp & ~3) | ((cgp & 1) << 1) | ((cgp >> 1) & 1);
        const int c0 = cg * 2;
        const bool real_vec = (d0 + 8 * g) < p.D;
        unsigned acc[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0;
        int rslot = 0;                            // ring slot of this row-step (idx % NR)
        // addressing, constant over the rows: the pair swizzle repeats every 4 pairs
        int ppo[4];
#pragma unroll
        for (int m = 0; m < 4; ++m) ppo[m] = ppcol(cg + m) * WDT + g * 8;
        const int phead = pd_off(c0, g), plast = pd_off(c0 + win - 1, g), pnext = pd_off(c0 + win, g);
        const int roff[2] = {wcol(c0) * WDTP + g * 8, wcol(c0 + 1) * WDTP + g * 8};
        // first output row of this band, this thread's first column and disparity vector
        int16_t* dst0 = C + ((size_t)y0 * p.W1 + (x0 + c0)) * p.Dp + vec_slot((d0 >> 3) + g, p.NL, p.K) * 8;
        for (int s = 0; s < nsteps + 2; ++s) {
            const int idx = s - 2;
            if (idx >= 0) {
                const int b = idx & 1;
                unsigned hs[4] = {0, 0, 0, 0};
                const uint16_t* prow = pd + b * WNCOL * WDTP;
                const uint16_t* qrow = pp + b * (WNCOL / 2) * WDT;
                // window of column c0 (even): SW2 column pairs + the single column c0 + win - 1
#pragma unroll
                for (int m = 0; m < WMAXSW; ++m) {
                    if (m < p.SW2) {
                        const uint4 v = *reinterpret_cast<const uint4*>(qrow + ppo[m & 3] + (m >> 2) * 4 * WDT);
                        hs[0] = __vadd2(hs[0], v.x); hs[1] = __vadd2(hs[1], v.y);
                        hs[2] = __vadd2(hs[2], v.z); hs[3] = __vadd2(hs[3], v.w);
                    }
                }
                {
                    const uint4 v = *reinterpret_cast<const uint4*>(prow + plast);
                    hs[0] = __vadd2(hs[0], v.x); hs[1] = __vadd2(hs[1], v.y);
                    hs[2] = __vadd2(hs[2], v.z); hs[3] = __vadd2(hs[3], v.w);
                }
                const uint4 head = *reinterpret_cast<const uint4*>(prow + phead);
                const bool store = idx >= 2 * p.SH2;
                uint16_t* slot = ring + rslot * WXT * WDTP;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    if (cc > 0) {
                        const uint4 vn = *reinterpret_cast<const uint4*>(prow + pnext);
                        hs[0] = __vsub2(__vadd2(hs[0], vn.x), head.x); hs[1] = __vsub2(__vadd2(hs[1], vn.y), head.y);
                        hs[2] = __vsub2(__vadd2(hs[2], vn.z), head.z); hs[3] = __vsub2(__vadd2(hs[3], vn.w), head.w);
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
                    if (store && x0 + c0 + cc < p.W1) {
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
            __syncthreads();
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(FULL, vmax, o));
    if ((tid & 31) == 0 && vmax > 0) atomicMax(maxC, vmax);
}

void launch_cost_wide(const uint2* pre1, const uint2* pre2, int16_t* C, int* maxC, const SgbmPlan& p, cudaStream_t st)
{
    const int smem = wide_layout(p.SH2).total;
    static bool once = false;
    if (!once) {
        cudaFuncSetAttribute(cost_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wide_layout(WMAXSW).total);
        cudaFuncSetAttribute(cost_wide_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        once = true;
    }
    dim3 g((p.W1 + WXT - 1) / WXT, p.Dp / WDT, (p.H + WRB - 1) / WRB);
    cost_wide_kernel<<<g, WCT, smem, st>>>(pre1, pre2, C, maxC, p);
}

}  // namespace wsg
