// sm_100a kernels of the dense matcher: the arithmetic of cv::StereoSGBM::compute as the reference
// calls it (src/wass_stereo/wass_stereo.cpp:775-782,837), specified in SURVEY.md Appendix A.
//
//   prefilter_kernel   A.1 + half-pixel bounds of A.2        image u8 -> 8-byte record per pixel
//   cost_kernel        A.2 pixel cost + A.3 box sum          -> C int16 [H][W1][Dp]
//   aggregate_kernel   A.4 one path direction per launch     C -> S (saturating int16 accumulate)
//   wta_kernel         A.5 + A.6 (one CTA per image row)     S -> raw disparity x16
//   median3_kernel     A.7
//
// All arithmetic is 16-bit integer, two disparities per 32-bit register, on the native packed
// instructions of sm_100a (VIADD.16x2, VIMNMX.S16x2, VIMNMX3.S16x2, VIADDMNMX.{S,U}16x2).
#include "sgbm_dev.cuh"
#include <cstdlib>

namespace wsg {

// ------------------------------------------------------------------------------------------------
// A.1 / A.2 : per-pixel record {p0, lo0, hi0, p1, lo1, hi1, 0, 0} (bytes)
// ------------------------------------------------------------------------------------------------
__global__ void prefilter_kernel(const uint8_t* __restrict__ img, size_t stride, uint2* __restrict__ pre,
                                 int H, int W, int ftzero)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= W) return;
    const uint8_t* r0 = img + (size_t)y * stride;
    const uint8_t* rm = img + (size_t)max(y - 1, 0) * stride;
    const uint8_t* rp = img + (size_t)min(y + 1, H - 1) * stride;
    auto sob = [&](int xx) -> int {
        if (xx <= 0 || xx >= W - 1) return ftzero;
        int g = 2 * ((int)r0[xx + 1] - (int)r0[xx - 1]) + ((int)rm[xx + 1] - (int)rm[xx - 1]) +
                ((int)rp[xx + 1] - (int)rp[xx - 1]);
        return min(max(g, -ftzero), ftzero) + ftzero;
    };
    auto raw = [&](int xx) -> int { return (xx <= 0 || xx >= W - 1) ? ftzero : (int)r0[xx]; };
    int c0 = sob(x), c1 = raw(x);
    int a0 = x > 0 ? (c0 + sob(x - 1)) / 2 : c0, b0 = x < W - 1 ? (c0 + sob(x + 1)) / 2 : c0;
    int a1 = x > 0 ? (c1 + raw(x - 1)) / 2 : c1, b1 = x < W - 1 ? (c1 + raw(x + 1)) / 2 : c1;
    unsigned lo0 = min(c0, min(a0, b0)), hi0 = max(c0, max(a0, b0));
    unsigned lo1 = min(c1, min(a1, b1)), hi1 = max(c1, max(a1, b1));
    uint2 o;
    o.x = (unsigned)c0 | (lo0 << 8) | (hi0 << 16) | ((unsigned)c1 << 24);
    o.y = lo1 | (hi1 << 8);
    pre[(size_t)y * W + x] = o;
}

void launch_prefilter(const uint8_t* img, size_t stride, uint2* pre, const SgbmPlan& p, cudaStream_t st)
{
    dim3 b(256), g((p.W + 255) / 256, p.H);
    prefilter_kernel<<<g, b, 0, st>>>(img, stride, pre, p.H, p.W, p.ftzero);
}

// ------------------------------------------------------------------------------------------------
// A.2 + A.3 : cost volume.  One CTA = XT columns x DT disparities, marching down a band of rows
// with a ring of (2*SH2+1) horizontal box sums in shared memory (vertical sliding window).
// ------------------------------------------------------------------------------------------------
static constexpr int XT = 32;     // output columns per CTA
static constexpr int DT = 64;     // disparities per CTA
static constexpr int CT = 256;    // threads: 32 columns x 8 groups of 8 disparities
static constexpr int RB = 128;    // rows per band
static constexpr int VT = 120;    // entries of the reversed img2 tables (>= XT+2*12+DT-1 .. rounded)
static constexpr int RVPAD = 4;   // s16 elements between table copy A and copy B (shared-memory bank skew)

struct CostSmem {
    // offsets into dynamic shared memory (bytes)
    int pd, uu, rv, ring, total;
};
__host__ __device__ inline CostSmem cost_smem_layout(int SW2, int SH2)
{
    CostSmem s;
    const int ncol = XT + 2 * SW2;
    s.pd = 0;                                  // u16 [ncol][DT]
    s.uu = s.pd + ncol * DT * 2;               // u32 [ncol][8]
    s.rv = s.uu + ncol * 8 * 4;                // s16 [2][8][VT] (+4 between the two copies: bank skew)
    s.ring = (s.rv + (2 * 8 * VT + RVPAD) * 2 + 15) & ~15;  // u16 [2*SH2+1][XT][DT]
    s.total = s.ring + (2 * SH2 + 1) * XT * DT * 2;
    return s;
}
int cost_smem_bytes(const SgbmPlan& p) { return cost_smem_layout(p.SW2, p.SH2).total; }

__global__ void __launch_bounds__(CT) cost_kernel(const uint2* __restrict__ pre1, const uint2* __restrict__ pre2,
                                                  int16_t* __restrict__ C, int* __restrict__ maxC, SgbmPlan p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const CostSmem L = cost_smem_layout(p.SW2, p.SH2);
    uint16_t* pdrow = reinterpret_cast<uint16_t*>(smem + L.pd);
    unsigned* uu = reinterpret_cast<unsigned*>(smem + L.uu);
    int16_t* rv = reinterpret_cast<int16_t*>(smem + L.rv);
    uint16_t* ring = reinterpret_cast<uint16_t*>(smem + L.ring);

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * XT;          // first output column (W1 space)
    const int d0 = blockIdx.y * DT;          // first disparity slot (logical, relative to minD)
    const int y0 = blockIdx.z * RB;
    const int y1 = min(y0 + RB, p.H);
    const int ncol = XT + 2 * p.SW2;
    const int NR = 2 * p.SH2 + 1;
    // absolute image x of the clamped halo columns
    const int xa = p.minX1 + min(max(x0 - p.SW2, 0), p.W1 - 1);
    const int xb = p.minX1 + min(max(x0 + XT - 1 + p.SW2, 0), p.W1 - 1);
    const int dlo = p.minD + d0;
    const int vtop = xb - dlo;               // largest img2 column touched; table index i <-> x' = vtop - i

    const int c4 = (tid >> 3) & 7, g = tid & 7;   // phase-2 role (tid<64): columns 4*c4..+3, disparities d0+8g..+7
    const bool real_vec = (d0 + 8 * g) < p.D;
    unsigned acc[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    int vmax = 0;

    const int nsteps = (y1 - y0) + 2 * p.SH2;
    for (int idx = 0; idx < nsteps; ++idx) {
        const int r = y0 - p.SH2 + idx;
        const int yy = min(max(r, 0), p.H - 1);
        // ---- tables for this row
        if (tid < ncol) {
            const int xx = p.minX1 + min(max(x0 - p.SW2 + tid, 0), p.W1 - 1);
            const uint2 q = pre1[(size_t)yy * p.W + xx];
            const int u0 = q.x & 255, l0 = (q.x >> 8) & 255, h0 = (q.x >> 16) & 255;
            const int u1 = q.x >> 24, l1 = q.y & 255, h1 = (q.y >> 8) & 255;
            auto bc = [](int v) -> unsigned { return ((unsigned)v & 0xFFFFu) * 0x10001u; };
            unsigned* o = uu + tid * 8;
            o[0] = bc(u0); o[1] = bc(-u0); o[2] = bc(l0); o[3] = bc(-h0);
            o[4] = bc(u1); o[5] = bc(-u1); o[6] = bc(l1); o[7] = bc(-h1);
        }
        if (tid < VT) {
            const int xp = min(max(vtop - tid, 0), p.W - 1);
            const uint2 q = pre2[(size_t)yy * p.W + xp];
            const int v0 = q.x & 255, l0 = (q.x >> 8) & 255, h0 = (q.x >> 16) & 255;
            const int v1 = q.x >> 24, l1 = q.y & 255, h1 = (q.y >> 8) & 255;
            const int16_t val[8] = {(int16_t)v0, (int16_t)-v0, (int16_t)l0, (int16_t)-h0,
                                    (int16_t)v1, (int16_t)-v1, (int16_t)l1, (int16_t)-h1};
#pragma unroll
            for (int qn = 0; qn < 8; ++qn) {
                rv[(0 * 8 + qn) * VT + tid] = val[qn];                           // copy A: rv[i]
                if (tid > 0) rv[RVPAD + (1 * 8 + qn) * VT + tid - 1] = val[qn];  // copy B: rv[i+1]
            }
        }
        __syncthreads();
        // ---- phase 1: pixel cost for ncol columns x DT disparities, two disparities per register
        for (int it = tid; it < ncol * 8; it += CT) {
            const int cc = it >> 3, gg = it & 7;
            uint4 out = make_uint4(0, 0, 0, 0);
            if (d0 + 8 * gg < p.D) {
                const int xx = p.minX1 + min(max(x0 - p.SW2 + cc, 0), p.W1 - 1);
                const int i0 = (xb - xx) + 8 * gg;
                const int par = i0 & 1;
                const unsigned* rvw = reinterpret_cast<const unsigned*>(rv + par * (8 * VT + RVPAD)) + ((i0 - par) >> 1);
                const uint4 ua = *reinterpret_cast<const uint4*>(uu + cc * 8);
                const uint4 ub = *reinterpret_cast<const uint4*>(uu + cc * 8 + 4);
                unsigned res[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned V0 = rvw[0 * (VT / 2) + k], nV0 = rvw[1 * (VT / 2) + k];
                    const unsigned Vl0 = rvw[2 * (VT / 2) + k], nVh0 = rvw[3 * (VT / 2) + k];
                    const unsigned V1 = rvw[4 * (VT / 2) + k], nV1 = rvw[5 * (VT / 2) + k];
                    const unsigned Vl1 = rvw[6 * (VT / 2) + k], nVh1 = rvw[7 * (VT / 2) + k];
                    // channel 0: c0 = max(0,u-vhi,vlo-u), c1 = max(0,v-uhi,ulo-v), c = min(c0,c1)
                    unsigned e0 = __vimax_s16x2_relu(__vadd2(ua.x, nVh0), __vadd2(Vl0, ua.y));
                    unsigned e1 = __vimax_s16x2_relu(__vadd2(V0, ua.w), __vadd2(ua.z, nV0));
                    unsigned ca = __vmins2(e0, e1);
                    unsigned f0 = __vimax_s16x2_relu(__vadd2(ub.x, nVh1), __vadd2(Vl1, ub.y));
                    unsigned f1 = __vimax_s16x2_relu(__vadd2(V1, ub.w), __vadd2(ub.z, nV1));
                    unsigned cb = __vmins2(f0, f1);
                    res[k] = ca + ((cb >> 2) & 0x3FFF3FFFu);
                }
                out = make_uint4(res[0], res[1], res[2], res[3]);
            }
            *reinterpret_cast<uint4*>(pdrow + cc * DT + gg * 8) = out;
        }
        __syncthreads();
        // ---- phase 2: horizontal box sum (sliding over 4 adjacent columns per thread), ring update,
        //      vertical sliding sum, store.  64 threads: 8 column groups x 8 disparity groups.
        if (tid < 64) {
            const int win = 2 * p.SW2 + 1;
            unsigned hs[4] = {0, 0, 0, 0};
            uint4 head[3];
            const uint16_t* prow = pdrow + (c4 * 4) * DT + g * 8;
#pragma unroll
            for (int i = 0; i < 3; ++i) {   // the three columns that leave the window while sliding
                head[i] = *reinterpret_cast<const uint4*>(prow + i * DT);
                if (i < win) {
                    hs[0] = __vadd2(hs[0], head[i].x); hs[1] = __vadd2(hs[1], head[i].y);
                    hs[2] = __vadd2(hs[2], head[i].z); hs[3] = __vadd2(hs[3], head[i].w);
                }
            }
            for (int i = 3; i < win; ++i) {
                const uint4 v = *reinterpret_cast<const uint4*>(prow + i * DT);
                hs[0] = __vadd2(hs[0], v.x); hs[1] = __vadd2(hs[1], v.y);
                hs[2] = __vadd2(hs[2], v.z); hs[3] = __vadd2(hs[3], v.w);
            }
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
                const int col = c4 * 4 + cc;
                if (cc > 0) {
                    const uint4 vn = *reinterpret_cast<const uint4*>(prow + (win - 1 + cc) * DT);
                    const uint4 vo = head[cc - 1];
                    hs[0] = __vsub2(__vadd2(hs[0], vn.x), vo.x); hs[1] = __vsub2(__vadd2(hs[1], vn.y), vo.y);
                    hs[2] = __vsub2(__vadd2(hs[2], vn.z), vo.z); hs[3] = __vsub2(__vadd2(hs[3], vn.w), vo.w);
                }
                uint4* slot = reinterpret_cast<uint4*>(ring + ((idx % NR) * XT + col) * DT + g * 8);
                unsigned* ac = acc[cc];
                if (idx >= NR) {
                    const uint4 o = *slot;
                    ac[0] = __vsub2(ac[0], o.x); ac[1] = __vsub2(ac[1], o.y);
                    ac[2] = __vsub2(ac[2], o.z); ac[3] = __vsub2(ac[3], o.w);
                }
                *slot = make_uint4(hs[0], hs[1], hs[2], hs[3]);
                ac[0] = __vadd2(ac[0], hs[0]); ac[1] = __vadd2(ac[1], hs[1]);
                ac[2] = __vadd2(ac[2], hs[2]); ac[3] = __vadd2(ac[3], hs[3]);
                if (idx >= 2 * p.SH2 && x0 + col < p.W1) {
                    const int y = r - p.SH2;
                    const int j = (d0 >> 3) + g;
                    int16_t* dst = C + ((size_t)y * p.W1 + (x0 + col)) * p.Dp + vec_slot(j, p.NL, p.K) * 8;
                    if (real_vec) {
                        *reinterpret_cast<uint4*>(dst) = interleave8(ac[0], ac[1], ac[2], ac[3]);
                        unsigned m = __vmaxs2(__vmaxs2(ac[0], ac[1]), __vmaxs2(ac[2], ac[3]));
                        vmax = max(vmax, max((int)(short)(m & 0xFFFF), (int)(short)(m >> 16)));
                    } else {
                        *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vmax = max(vmax, __shfl_xor_sync(FULL, vmax, o));
    if ((tid & 31) == 0 && vmax > 0) atomicMax(maxC, vmax);
}

void launch_cost(const uint2* pre1, const uint2* pre2, int16_t* C, int* maxC, const SgbmPlan& p, cudaStream_t st,
                 int* launches)
{
    static int impl = -1;
    if (impl < 0) { const char* e = getenv("WSG_COST_IMPL"); impl = e ? atoi(e) : 1; }
    if (impl != 0 && cost_wide_supported(p)) {
        launch_cost_wide(pre1, pre2, C, maxC, p, st);
        if (launches) *launches += 1;
        return;
    }
    const int smem = cost_smem_bytes(p);
    cudaFuncSetAttribute(cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    dim3 g((p.W1 + XT - 1) / XT, p.Dp / DT, (p.H + RB - 1) / RB);
    cost_kernel<<<g, CT, smem, st>>>(pre1, pre2, C, maxC, p);
    if (launches) *launches += 1;
}

// ------------------------------------------------------------------------------------------------
// A.4 : one aggregation direction.  A group of NL lanes walks one chain of pixels; lane l keeps the
// normalised path cost N(d) = L(d) - min_d L for its K*8 consecutive disparities in K*4 registers.
//   t(d)   = min( N(d), min(N(d-1), N(d+1), P2-P1) + P1 )           (== min(..., m+P2) - m)
//   L(d)   = min( t(d) + C(d), 32767 )
//   S(d)   = min( S(d) + L(d), 32767 )
// Diagonal chains wrap around the image edge and restart (out-of-image predecessor => N = 0), so
// every launch is W1 (or H) fully independent chains of equal length.
// ------------------------------------------------------------------------------------------------
struct AggArgs {
    int H, W1, Dp, D;
    int sx, sy;        // successor step
    int nchains, len;
    unsigned P1p, P2mP1p;
};

template <int NL, int K, int PF, bool FIRST, bool DIAG, bool HASPAD>
__global__ void __launch_bounds__(128) aggregate_kernel(const uint4* __restrict__ C, uint4* __restrict__ S, AggArgs a)
{
    constexpr int NR = 4 * K;  // packed registers of state per lane
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / NL;
    const int l = threadIdx.x % NL;
    const bool active = gid < a.nchains;
    const int chain = active ? gid : a.nchains - 1;

    // pad mask: logical vectors at or beyond D are forced to 32767 (== the L(d=D) border value)
    unsigned padm[K];
#pragma unroll
    for (int k = 0; k < K; ++k) padm[k] = ((l * K + k) * 8 >= a.D) ? SAT2 : 0u;

    // cursors in units of 16-byte vectors: idx(x,y) = (y*W1+x)*Dp/8 + lane slot
    const int Dp8 = a.Dp >> 3;
    const int dstep = (a.sx + a.sy * a.W1) * Dp8;
    const int wrapfix = a.W1 * Dp8;
    int x0, y0;
    if (a.sy == 0) { y0 = chain; x0 = a.sx > 0 ? 0 : a.W1 - 1; }
    else           { x0 = chain; y0 = a.sy > 0 ? 0 : a.H - 1; }
    int xp = x0, ip = (y0 * a.W1 + x0) * Dp8 + l;   // prefetch cursor
    int xc = xp, ic = ip;                           // compute cursor
    auto advance = [&](int& x, int& idx) {
        idx += dstep;
        if (DIAG) {
            x += a.sx;
            if (x >= a.W1) { x = 0; idx -= wrapfix; }
            else if (x < 0) { x = a.W1 - 1; idx += wrapfix; }
        }
    };

    uint4 cb[PF][K], sb[PF][K];
    int spf = 0;
#pragma unroll
    for (int u = 0; u < PF; ++u) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            cb[u][k] = ldg_stream(C + ip + k * NL);
            if (!FIRST) sb[u][k] = ldg_rw(S + ip + k * NL);
        }
        if (spf + 1 < a.len) advance(xp, ip);
        ++spf;
    }

    unsigned R[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) R[j] = 0;

    for (int base = 0; base < a.len; base += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int step = base + u;
            unsigned Cw[NR], v[NR];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                Cw[4 * k] = cb[u][k].x; Cw[4 * k + 1] = cb[u][k].y; Cw[4 * k + 2] = cb[u][k].z; Cw[4 * k + 3] = cb[u][k].w;
            }
            if (DIAG) {
                // a diagonal chain that just wrapped around the image edge restarts (out-of-image predecessor)
                const bool restart = (a.sx > 0 && xc == 0) || (a.sx < 0 && xc == a.W1 - 1);
#pragma unroll
                for (int j = 0; j < NR; ++j) R[j] = restart ? 0u : R[j];
            }
            agg_step<NL, NR, HASPAD>(R, Cw, v, l, a.P1p, a.P2mP1p, padm);
            const bool st = active && step < a.len;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                uint4 out;
                if (FIRST) {
                    out = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                } else {
                    out.x = __viaddmin_u16x2(sb[u][k].x, v[4 * k], SAT2);
                    out.y = __viaddmin_u16x2(sb[u][k].y, v[4 * k + 1], SAT2);
                    out.z = __viaddmin_u16x2(sb[u][k].z, v[4 * k + 2], SAT2);
                    out.w = __viaddmin_u16x2(sb[u][k].w, v[4 * k + 3], SAT2);
                }
                if (st) stg_stream(S + ic + k * NL, out);
            }
            advance(xc, ic);
            // refill this stage for step + PF (cursor frozen at the last pixel once the chain is exhausted)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                cb[u][k] = ldg_stream(C + ip + k * NL);
                if (!FIRST) sb[u][k] = ldg_rw(S + ip + k * NL);
            }
            if (spf + 1 < a.len) advance(xp, ip);
            ++spf;
        }
    }
}

template <int NL, int K, bool HASPAD>
static void launch_agg_t(const int16_t* C, int16_t* S, const AggArgs& a, bool first, cudaStream_t st)
{
    constexpr int PF = K == 1 ? 6 : (K == 2 ? 3 : 2);
    const int threads = 128;
    const long long total = (long long)a.nchains * NL;
    const int blocks = (int)((total + threads - 1) / threads);
    const uint4* c = reinterpret_cast<const uint4*>(C);
    uint4* s = reinterpret_cast<uint4*>(S);
    const bool diag = a.sx != 0 && a.sy != 0;
    if (first) {
        if (diag) aggregate_kernel<NL, K, PF, true, true, HASPAD><<<blocks, threads, 0, st>>>(c, s, a);
        else      aggregate_kernel<NL, K, PF, true, false, HASPAD><<<blocks, threads, 0, st>>>(c, s, a);
    } else {
        if (diag) aggregate_kernel<NL, K, PF, false, true, HASPAD><<<blocks, threads, 0, st>>>(c, s, a);
        else      aggregate_kernel<NL, K, PF, false, false, HASPAD><<<blocks, threads, 0, st>>>(c, s, a);
    }
}

void launch_aggregate_dir(const int16_t* C, int16_t* S, int dir, bool first, const SgbmPlan& p, cudaStream_t st)
{
    static const int pred[8][2] = {{-1, 0}, {-1, -1}, {0, -1}, {1, -1}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
    AggArgs a;
    a.H = p.H; a.W1 = p.W1; a.Dp = p.Dp; a.D = p.D;
    a.sx = -pred[dir][0]; a.sy = -pred[dir][1];
    a.nchains = a.sy == 0 ? p.H : p.W1;
    a.len = a.sy == 0 ? p.W1 : p.H;
    a.P1p = ((unsigned)p.P1 & 0xFFFFu) * 0x10001u;
    a.P2mP1p = ((unsigned)(p.P2 - p.P1) & 0xFFFFu) * 0x10001u;
    const bool pad = p.Dp != p.D;
#define WSG_AGG_CASE(nl, k)                                                    \
    if (p.NL == nl && p.K == k) {                                              \
        if (pad) launch_agg_t<nl, k, true>(C, S, a, first, st);               \
        else     launch_agg_t<nl, k, false>(C, S, a, first, st);              \
        return;                                                                \
    }
    WSG_AGG_CASE(8, 1) WSG_AGG_CASE(16, 1) WSG_AGG_CASE(32, 1) WSG_AGG_CASE(32, 2)
    WSG_AGG_CASE(32, 3) WSG_AGG_CASE(32, 4) WSG_AGG_CASE(32, 5) WSG_AGG_CASE(8, 4) WSG_AGG_CASE(16, 2)
#undef WSG_AGG_CASE
}

// ------------------------------------------------------------------------------------------------
// A.5 + A.6 : winner-take-all, uniqueness, sub-pixel, right-view map and LR check; one CTA per row.
// The sequential right-to-left scan of the reference (first writer with strictly smaller cost wins)
// becomes an atomicMin on the key (minS, W1-1-x, d).
// ------------------------------------------------------------------------------------------------
template <int NL, int K>
__global__ void __launch_bounds__(256) wta_kernel(const int16_t* __restrict__ S, int16_t* __restrict__ raw, SgbmPlan p)
{
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    int16_t* d1 = reinterpret_cast<int16_t*>(smem + (size_t)p.W * 8);
    const int y = blockIdx.x;
    const int tid = threadIdx.x;
    for (int x = tid; x < p.W; x += blockDim.x) { keys[x] = ~0ull; d1[x] = (int16_t)p.INVALID; }
    __syncthreads();

    constexpr int G = 256 / NL;
    constexpr int NV8 = K * 8;               // disparities held by one lane
    const int grp = tid / NL, l = tid % NL;
    const int iters = (p.W1 + G - 1) / G;
    const int dlane = l * NV8;               // first logical disparity of this lane
    // uniqueness test  S(d)*(100-uniq) < minS*100  <=>  S(d) <= Tm,  Tm = floor((minS*100-1)/(100-uniq))  (0<=uniq<100)
    const int udiv = 100 - p.uniq;
    const float urcp = udiv > 0 ? 1.0f / (float)udiv : 0.f;
    for (int it = 0; it < iters; ++it) {
        const int xh_raw = p.W1 - 1 - (it * G + grp);
        const bool act = xh_raw >= 0;
        const int xh = act ? xh_raw : 0;
        const int16_t* Sp = S + ((size_t)y * p.W1 + xh) * p.Dp;
        // 32-bit keys (S << 16 | d): their minimum is the FIRST disparity with minimal S
        unsigned key[NV8];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint4 v = *reinterpret_cast<const uint4*>(Sp + ((size_t)k * NL + l) * 8);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {   // interleaved vector: register e = (disparity e, disparity 4+e), see vec_pos
                key[k * 8 + e] = (w[e] << 16) + (unsigned)(dlane + k * 8 + e);
                key[k * 8 + 4 + e] = (w[e] & 0xFFFF0000u) + (unsigned)(dlane + k * 8 + 4 + e);
            }
        }
        unsigned kmin = 0xFFFFFFFFu;
#pragma unroll
        for (int e = 0; e < NV8; ++e)
            if (dlane + e < p.D) kmin = min(kmin, key[e]);
        kmin = group_min_u32<NL>(kmin);
        const int minS = (int)(kmin >> 16), best = (int)(kmin & 0xFFFFu);
        int bad = 0;
        if (udiv > 0) {
            const int n = minS * 100 - 1;                    // < 3.3e6: exact in float
            int Tm = n < 0 ? -1 : (int)((float)n * urcp);
            if (n >= 0) { if ((Tm + 1) * udiv <= n) ++Tm; else if (Tm * udiv > n) --Tm; }
            const unsigned Tkey = Tm < 0 ? 0u : (((unsigned)min(Tm, 65535) << 16) | 0xFFFFu);   // S <= 32767 < 65535
            const int rel = best - dlane;                   // position of best inside this lane (may be outside)
#pragma unroll
            for (int e = 0; e < NV8; ++e) {
                const bool near = (unsigned)(e - rel + 1) <= 2u;
                if (Tm >= 0 && key[e] <= Tkey && !near && dlane + e < p.D) bad = 1;
            }
        } else {
            // uniquenessRatio >= 100: S*(100-uniq) < minS*100 evaluated literally
#pragma unroll
            for (int e = 0; e < NV8; ++e) {
                const int sv = (int)(key[e] >> 16), d = dlane + e;
                if (d < p.D && sv * udiv < minS * 100 && abs(best - d) > 1) bad = 1;
            }
        }
        bad = __any_sync(FULL, bad && true) ? (NL == 32 ? 1 : bad) : 0;
        if (NL < 32) {
#pragma unroll
            for (int o = NL / 2; o > 0; o >>= 1) bad |= __shfl_xor_sync(FULL, bad, o, NL);
        }
        if (act && l == 0 && !bad) {
            const int x = xh + p.minX1;
            const int x2 = x - best - p.minD;
            const unsigned long long k64 = ((unsigned long long)minS << 40) |
                                           ((unsigned long long)(p.W1 - 1 - xh) << 16) | (unsigned long long)best;
            atomicMin(&keys[x2], k64);
            int dd = best * 16;
            if (best > 0 && best < p.D - 1) {
                const int jm = (best - 1) >> 3, jp = (best + 1) >> 3;
                const int sm = Sp[vec_slot(jm, NL, K) * 8 + vec_pos((best - 1) & 7)];
                const int sp = Sp[vec_slot(jp, NL, K) * 8 + vec_pos((best + 1) & 7)];
                const int den = max(sm + sp - 2 * minS, 1);
                dd += ((sm - sp) * 16 + den) / (2 * den);
            }
            d1[x] = (int16_t)(dd + p.minD * 16);
        }
    }
    __syncthreads();
    for (int x = tid; x < p.W; x += blockDim.x) {
        int dv = d1[x];
        if (x >= p.minX1 && x < p.maxX1 && dv != p.INVALID) {
            const int a = dv >> 4, b = (dv + 15) >> 4;
            const int xa = x - a, xb = x - b;
            bool ca = false, cbb = false;
            if (xa >= 0 && xa < p.W) {
                const unsigned long long k = keys[xa];
                const int d2 = (k == ~0ull) ? p.INVALID : (int)(k & 0xFFFFu) + p.minD;
                ca = d2 >= p.minD && abs(d2 - a) > p.d12;
            }
            if (xb >= 0 && xb < p.W) {
                const unsigned long long k = keys[xb];
                const int d2 = (k == ~0ull) ? p.INVALID : (int)(k & 0xFFFFu) + p.minD;
                cbb = d2 >= p.minD && abs(d2 - b) > p.d12;
            }
            if (ca && cbb) dv = p.INVALID;
        }
        raw[(size_t)y * p.W + x] = (int16_t)dv;
    }
}

void launch_wta(const int16_t* S, int16_t* raw, const SgbmPlan& p, cudaStream_t st)
{
    const int smem = p.W * 10 + 16;
#define WSG_WTA_CASE(nl, k)                                                                              \
    if (p.NL == nl && p.K == k) {                                                                        \
        cudaFuncSetAttribute(wta_kernel<nl, k>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);     \
        wta_kernel<nl, k><<<p.H, 256, smem, st>>>(S, raw, p);                                            \
        return;                                                                                          \
    }
    WSG_WTA_CASE(8, 1) WSG_WTA_CASE(16, 1) WSG_WTA_CASE(32, 1) WSG_WTA_CASE(32, 2)
    WSG_WTA_CASE(32, 3) WSG_WTA_CASE(32, 4) WSG_WTA_CASE(32, 5) WSG_WTA_CASE(8, 4) WSG_WTA_CASE(16, 2)
#undef WSG_WTA_CASE
}

// ------------------------------------------------------------------------------------------------
// A.7 : 3x3 median on int16, replicate border
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cswap(int& a, int& b) { const int t = min(a, b); b = max(a, b); a = t; }

__global__ void median3_kernel(const int16_t* __restrict__ src, int16_t* __restrict__ dst, int rows, int cols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= cols) return;
    int v[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int yy = min(max(y + j - 1, 0), rows - 1);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int xx = min(max(x + i - 1, 0), cols - 1);
            v[j * 3 + i] = src[(size_t)yy * cols + xx];
        }
    }
    // 19-exchange median-of-9 network
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
    cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]);
    cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
    cswap(v[4], v[2]);
    dst[(size_t)y * cols + x] = (int16_t)v[4];
}

void launch_median3(const int16_t* src, int16_t* dst, int rows, int cols, cudaStream_t st)
{
    dim3 b(256), g((cols + 255) / 256, rows);
    median3_kernel<<<g, b, 0, st>>>(src, dst, rows, cols);
}

// ------------------------------------------------------------------------------------------------
// A.7 (optional): cv::filterSpeckles as StereoSGBM::compute calls it when speckleWindowSize > 0
// (wass_stereo.cpp:781-782 sets the two parameters; off at the reference's defaults).  The flood fill of the
// reference visits 4-neighbours whose values differ by at most maxDiff = 16*speckleRange, so its regions are the
// connected components of that (symmetric) relation among valid pixels: lock-free union-find with atomicMin,
// component sizes, then every pixel of a component of at most speckleWindowSize pixels becomes invalid.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sp_find(int* L, int x)
{
    // path halving; the shortcut is installed with a compare-and-swap so that it can never undo a concurrent union
    int p = L[x];
    while (p != x) {
        const int g = L[p];
        if (g != p) atomicCAS(&L[x], p, g);
        x = p; p = g;
    }
    return x;
}
__device__ __forceinline__ void sp_union(int* L, int a, int b)
{
    while (true) {
        a = sp_find(L, a); b = sp_find(L, b);
        if (a == b) return;
        if (a > b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[b], a);
        if (old == b) return;
        b = old;
    }
}
__global__ void speckle_init_kernel(const int16_t* __restrict__ d, int* __restrict__ L, unsigned* __restrict__ cnt, int n, int invalid)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    L[i] = d[i] == invalid ? -1 : i;
    cnt[i] = 0;
}
__global__ void speckle_merge_kernel(const int16_t* __restrict__ d, int* L, int rows, int cols, int invalid, int maxDiff)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int i = y * cols + x;
    const int v = d[i];
    if (v == invalid) return;
    if (x + 1 < cols) { const int w = d[i + 1]; if (w != invalid && abs(v - w) <= maxDiff) sp_union(L, i, i + 1); }
    if (y + 1 < rows) { const int w = d[i + cols]; if (w != invalid && abs(v - w) <= maxDiff) sp_union(L, i, i + cols); }
}
__global__ void speckle_count_kernel(int* L, unsigned* cnt, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || L[i] < 0) return;
    const int r = sp_find(L, i);
    L[i] = r;
    const unsigned act = __activemask();
    const unsigned peers = __match_any_sync(act, r);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cnt[r], (unsigned)__popc(peers));
}
__global__ void speckle_apply_kernel(int16_t* d, const int* __restrict__ L, const unsigned* __restrict__ cnt, int n, int invalid,
                                     unsigned maxSize)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || L[i] < 0) return;
    if (cnt[L[i]] <= maxSize) d[i] = (int16_t)invalid;
}

void launch_filter_speckles(int16_t* disp, int rows, int cols, int invalid, int maxSpeckleSize, int maxDiff, int* labels,
                            unsigned* counts, cudaStream_t st)
{
    const int n = rows * cols;
    const int nb = (n + 255) / 256;
    speckle_init_kernel<<<nb, 256, 0, st>>>(disp, labels, counts, n, invalid);
    dim3 b(256), g((cols + 255) / 256, rows);
    speckle_merge_kernel<<<g, b, 0, st>>>(disp, labels, rows, cols, invalid, maxDiff);
    speckle_count_kernel<<<nb, 256, 0, st>>>(labels, counts, n);
    speckle_apply_kernel<<<nb, 256, 0, st>>>(disp, labels, counts, n, invalid, (unsigned)maxSpeckleSize);
}

}  // namespace wsg
