// Fused path aggregation of cv::StereoSGBM (call site
// src/wass_stereo/wass_stereo.cpp:837): one launch per SWEEP instead of one per direction, and one launch
// for a whole BATCH of frames instead of one per frame.
//
// A sweep runs the four directions whose predecessors are (x-1,y) (x-1,y-1) (x,y-1) (x+1,y-1)
// [sweep 1: r0..r3] -- or, rotated by 180 degrees, (x+1,y) (x+1,y+1) (x,y+1) (x-1,y+1)
// [sweep 2: r4,r7,r6,r5] -- as a skew-2 wavefront: one warp walks one image row, row y+1 trails
// row y by two columns, so every predecessor state already exists when a pixel is reached.  The
// cost vector C(p,.) is read from HBM exactly once per sweep, the four L_r(p,.) are summed in
// registers, and S is written once (sweep 1) or consumed on the spot by the winner-take-all
// stage (sweep 2): 4V of HBM traffic for MODE_HH instead of the 23V of eight separate launches.
//
//   CTA      = one persistent worker per SM: R row warps + one helper warp; it takes BANDS of R consecutive rows from
//              a ticket counter until none is left
//   ticket t = (frame t % nframes, band t / nframes): the bands of all frames of the batch are interleaved, so the
//              148 workers hold ~148/nframes consecutive bands of every frame.  A band therefore trails its
//              predecessor by ~16 * nframes columns instead of the minimum -- it never waits for it -- and the
//              wavefront of one frame fills while another drains: no SM idles until the batch runs out of bands.
//              A band only ever waits for bands with smaller tickets, which are running or done: no deadlock,
//              no co-residency assumption.
//   row->row = normalised states N_r = L_r - min L_r handed down through shared-memory rings (NS columns deep,
//              progress counters instead of CTA barriers: warps drift freely).  Every row warp runs the same code:
//              ring r in, ring r+1 out.
//   CTA->CTA = the helper warp drains ring R (the band's last row) into the global hand-off buffer and fills ring 0
//              from the previous band's part of it.  Every 32-bit word carries a 2-bit epoch tag in the two sign
//              bits a state never uses, so the consumer validates data word by word without fences or flags.
//
// All spin loops are bounded: on overrun the kernel raises an error flag and runs to completion.
#include "sgbm_dev.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <type_traits>

namespace wsg {

static constexpr unsigned TAGBITS = 0x80008000u;
static constexpr int SPIN_LIMIT = 1 << 22;
static constexpr int PROG_INF = 0x3fffffff;

// rows per band / ring depth for D <= 256 (K == 1) and D <= 512 (K == 2); experiments: -DWSG_SW_ROWS1=.. -DWSG_SW_NS1=..
#ifndef WSG_SW_ROWS1
#define WSG_SW_ROWS1 15
#endif
#ifndef WSG_SW_NS1
#define WSG_SW_NS1 8
#endif
#ifndef WSG_SW_ROWS2
#define WSG_SW_ROWS2 7
#endif
#ifndef WSG_SW_NS2
#define WSG_SW_NS2 6
#endif
#ifndef WSG_SW_TMA
#define WSG_SW_TMA 0         // 1: C and S staged through shared memory by TMA bulk copies (two pixels per copy, mbarrier completion)
#endif                       //    instead of register loads -- measured, not faster; needs -DWSG_SW_NS1=6 to fit
#ifndef WSG_SW_STAGGER
#define WSG_SW_STAGGER -1  // extra columns a row lets the row above get ahead before it starts; -1: half the ring's slack
#endif

// Shared-memory map of a worker (uint4 units).  K = 16-byte vectors per lane and pixel (D <= 256: 1, D <= 512: 2).
// The first sweep (MODE 0) streams only C and has no winner-take-all: it needs neither S staging nor scratch.
template <int K, int R, int NS, int MODE> struct SweepCfg {
    static constexpr int PIX_V = K * 32;                      // uint4 per pixel
    static constexpr int SLOT_V = 3 * PIX_V;                  // uint4 per ring slot: [dir][k][lane]
    static constexpr int RING_V = NS * SLOT_V;
    static constexpr int RINGS_V = (R + 1) * RING_V;          // ring r: states of the row above row r; ring R: out of the band
    static constexpr int SCR_V = MODE == 2 ? R * PIX_V : 0;   // per-row WTA scratch
    // TMA staging (WSG_SW_TMA): per row and stream two buffers of two pixels; per row two mbarriers (one per buffer)
    static constexpr int STREAMS = MODE == 0 ? 1 : 2;
    static constexpr int STG_V = WSG_SW_TMA ? STREAMS * 4 * PIX_V : 0;
    static constexpr int STGS_V = R * STG_V;
    static constexpr int MBAR_V = WSG_SW_TMA ? R : 0;         // 16 bytes per row
    static constexpr int SMEM = (RINGS_V + SCR_V + STGS_V + MBAR_V) * 16 + 64;  // (+64: the scratch of the last row is read one element beyond)
    static constexpr int THREADS = (R + 1) * 32;
    static constexpr int HD = NS - 2 < 4 ? NS - 2 : 4;        // boundary columns the helper polls per round trip
    // A row may run lag = 2 .. NS-2 columns behind the row above (2: it needs column x+1; NS-2: the ring is full).  Rows
    // that start at the minimum stay there -- every row polls for every column of the one above -- so each row first
    // lets its producer get STAGGER columns further ahead: the chain then sits in the middle of its slack.
    static constexpr int STAGGER = WSG_SW_STAGGER >= 0 ? WSG_SW_STAGGER : (NS - 4) / 2;
    static_assert(SMEM <= 227 * 1024, "worker does not fit an SM");
    static_assert(STAGGER >= 0 && STAGGER <= NS - 4 + 0 || NS < 4, "stagger beyond the ring's slack");
};

// TMA prefetch: one thread asks the copy engine to pull `bytes` (a multiple of 16) of global memory into L2.  A row warp
// issues one per stream and four pixels, eight pixels ahead of its register loads, which then hit L2 instead of HBM.
__device__ __forceinline__ void bulk_prefetch_l2(const void* gptr, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared with mbarrier completion, and the mbarrier operations it needs (WSG_SW_TMA)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef WSG_SW_CA
#define WSG_SW_CA 2          // pixels the register load of C runs ahead of its use (2..3; the horizontal path uses C(x+1))
#endif
#ifndef WSG_SW_SA
#define WSG_SW_SA 2          // the same for S (1..2)
#endif
static constexpr int SW_PF_AHEAD = 8;       // pixels between the L2 prefetch and the register load of the same pixel

struct SweepArgs {
    int H, W1, W, D;
    int Dp8;                 // 16-byte vectors per pixel (= 32*K)
    int flip;                // 0: top->bottom, left->right;  1: rotated by 180 degrees
    int nframes;             // frames of the batch (tickets interleave them)
    int rows;                // rows per band of THIS launch, <= R (the kernel's row warps): see sweep_rows_per_band
    const uint4* C;          // [nframes] cost volumes, vol_v uint4 apart (readable PFD pixels beyond either end)
    uint4* S;                // [nframes] aggregated volumes, same spacing
    size_t vol_v;
    unsigned P1p, P2mP1p;
    int P2;
    const int* maxC;         // max over the cost volume of frame f at maxC[f * scal_stride] (written by the cost kernel)
    int scal_stride;
    unsigned one;            // 1 (kept opaque to the compiler: see add_on_fma)
    unsigned tag;            // epoch tag of this launch (bits 15 and 31)
    uint4* bnd;              // [nframes][nbands-1][W1][3][K][32]
    size_t bnd_v;            // uint4 per frame of bnd
    int* ticket;             // [0] tickets handed out
    int eager;               // polls of a progress counter before the waiter starts to sleep between polls
    int* err;
    int* dbg;                // optional: per ticket {SM id, start ns, end ns}, or null
    // winner-take-all (MODE 2)
    unsigned long long* keys;   // [nframes][H][W]  (minS, W1-1-x, d) of the best match that lands on x2 (A.5)
    int16_t* d1;                // [nframes][H][W]  left-view disparity before the LR check
    int minD, minX1, uniq, INVALID;
    unsigned umagic;            // ceil(2^32 / (100 - uniq)); the fused WTA needs 100 - uniq >= 2
    int qm1;                    // 100 - uniq - 1
};

__device__ __forceinline__ uint4 ld_volatile(const uint4* p)
{
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_volatile(uint4* p, const uint4& v)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// Wait until *flag >= need (shared-memory progress counter of a neighbouring warp).  Bounded: on overrun (or when
// any other waiter has already given up) raise the error flag and stop waiting for good.
__device__ __forceinline__ void wait_prog(volatile int* flag, int need, int& seen, int* err, int eager_spins = 64)
{
    if (seen < need) {
        int spins = 0;
        while ((seen = *flag) < need) {
            // A warp that is not yet due (or is held back by back-pressure) must not steal issue slots.  Measured
            // alternatives, all ~5 % SLOWER: a tight load/compare/branch poll, the same with a 32 ns sleep per poll, and
            // sleeping sooner (a sleep on the row-to-row critical path costs more than the polling it saves).
            if (++spins > eager_spins) __nanosleep(spins > 4096 ? 400 : (eager_spins > 1 ? 40 : 150));
            if ((spins & 1023) == 0 && (spins > SPIN_LIMIT || *reinterpret_cast<volatile int*>(err) != 0)) {
                *err = 1;
                seen = PROG_INF;
                break;
            }
        }
    }
    // Shared-memory requests of one warp are served in issue order, so a counter read that saw the value is followed
    // by data reads that see the data; only the compiler has to be kept from moving them.
    asm volatile("" ::: "memory");
}

// A.5, split in two so that the scalar tail is paid once per 32 pixels instead of once per pixel.
//
// wta_eval (every pixel, warp-uniform result): s = final S of one pixel, 8*K consecutive disparities per lane (a lane is
// all real or all pad: numDisparities is a multiple of 16); `scratch` holds the same S in shared memory, lane-major.
//   winner      first d with minimal S: warp minimum of the 32-bit keys (S << 16 | d)
//   uniqueness  reject iff some d outside {best-1,best,best+1} has S(d)*(100-uniq) < minS*100, i.e. S(d) <= Tm with
//               Tm = floor((minS*100-1)/(100-uniq)).  Counted instead of searched: #(S <= Tm) over all d, by a packed
//               subtract whose sign bits are the comparison results, against the same count inside the window.
//   returns     key = minS << 16 | best, and nb = S(best-1) | S(best+1) << 16 | accepted << 31
// Branch-free, so the caller can run it one pixel late, interleaved with the next pixel's path steps.  Needs
// 0 <= uniquenessRatio < 99 (the host routes anything else to wta_kernel).
template <int K, bool HASPAD>
__device__ __forceinline__ void wta_eval(const unsigned (&s)[4 * K], int l, const SweepArgs& a, const uint16_t* scratch,
                                         unsigned& key, unsigned& nb)
{
    constexpr int NV8 = 8 * K;
    const int dlane = l * NV8;
    const bool padlane = HASPAD && dlane >= a.D;
    unsigned kmin = 0xFFFFFFFFu;
#pragma unroll
    for (int e = 0; e < 4 * K; ++e) {   // register 4k+i holds disparities (8k+i, 8k+4+i) of the lane, see vec_pos
        const unsigned dlo = (unsigned)(dlane + 8 * (e >> 2) + (e & 3)), dhi = dlo + 4u;
        const unsigned klo = s[e] * 65536u + dlo;
        const unsigned khi = (s[e] & 0xFFFF0000u) | dhi;
        kmin = __vimin3_u32(kmin, klo, khi);
    }
    if (padlane) kmin = 0xFFFFFFFFu;
    kmin = __reduce_min_sync(FULL, kmin);
    const int minS = (int)(kmin >> 16), best = (int)(kmin & 0xFFFFu);
    // Tm = floor((100 minS - 1) / q), q = 100 - uniq; for minS = 0 that is -1 (nothing can be below the winner).
    // floor((n - 1) / q) = ceil(n / q) - 1 = floor((n + q - 1) / q) - 1 has a non-negative numerator for every minS, so one
    // multiply-high by ceil(2^32 / q) does it (exact below 2^22; q = 1 has no such constant: the host routes it to wta_kernel).
    const int Tm = (int)__umulhi((unsigned)(minS * 100 + a.qm1), a.umagic) - 1;
    // packed count of S <= Tm: (0x8000 + T - S) keeps bit 15 iff T >= S (S <= 0x7fff; T = -1 gives 0x7fff - S: never).
    // The subtractions run on the FMA pipe; one byte gather per register pair collects the four flag bytes.
    const unsigned T2 = (unsigned)(min(Tm, 32767) + 0x8000) * 0x10001u;
    const unsigned minus_one = 0u - a.one;
    int cnt = 0;
#pragma unroll
    for (int e = 0; e < 4 * K; e += 2) {
        const unsigned t0 = add_on_fma(s[e], T2, minus_one), t1 = add_on_fma(s[e + 1], T2, minus_one);     // T2 - s
        cnt += __popc(__byte_perm(t0, t1, 0x7531) & 0x80808080u);
    }
    if (padlane) cnt = 0;
    const int total = __reduce_add_sync(FULL, cnt);
    // the winner's neighbours from the copy of S in shared memory (natural disparity order); the elements at -1 and D are
    // whatever lies beside the row's scratch area -- they only count where they exist
    const int sm = scratch[best - 1], sp = scratch[best + 1];
    const int inwin = (minS <= Tm) + (best > 0 && sm <= Tm) + (best < a.D - 1 && sp <= Tm);
    key = kmin;
    nb = (unsigned)sm | ((unsigned)(sp & 0x7FFF) << 16) | (total <= inwin ? 0x80000000u : 0u);   // (bit 31 is the flag: mask a stray sp)
}

// wta_flush (every 32 pixels): lane i holds the record of logical column xbase + i.  Sub-pixel parabola, right-view map
// (A.5) and the left-view map, all 32 lanes busy on 32 different pixels; the d1 row segment is one coalesced store.
__device__ __forceinline__ void wta_flush(unsigned key, unsigned nb, int xl, bool valid, const SweepArgs& a,
                                          unsigned long long* keys_row, int16_t* d1_row)
{
    if (!valid || !(nb & 0x80000000u)) return;
    const int minS = (int)(key >> 16), best = (int)(key & 0xFFFFu);
    const int sm = (int)(nb & 0xFFFFu), sp = (int)((nb >> 16) & 0x7FFFu);      // (a real S is <= 32767)
    const int xh = a.flip ? a.W1 - 1 - xl : xl;     // physical column in W1 space
    const int x = xh + a.minX1;
    const unsigned long long k64 = ((unsigned long long)minS << 40) | ((unsigned long long)(a.W1 - 1 - xh) << 16) |
                                   (unsigned long long)best;
    atomicMin(keys_row + (x - best - a.minD), k64);
    int dd = best * 16;
    if (best > 0 && best < a.D - 1) {
        // trunc(((sm-sp)*16 + den) / (2*den)): |numerator| < 2^24, so an approximate float quotient is off by at most one
        const int den = max(sm + sp - 2 * minS, 1), den2 = 2 * den;
        const int num = (sm - sp) * 16 + den, an = abs(num);
        int q = (int)__fdividef((float)an, (float)den2);
        const int rem = an - q * den2;
        q += rem >= den2 ? 1 : (rem < 0 ? -1 : 0);
        dd += num < 0 ? -q : q;
    }
    d1_row[x] = (int16_t)(dd + a.minD * 16);
}

// Per-row state of a sweep.  The C and S streams are prefetched into REGISTERS: a row warp takes ~1500 clocks per pixel
// (15 rows share an SM), so a load issued two pixels ahead has thousands of clocks to land, and the register sets rotate
// through a loop unrolled by four -- no staging in shared memory, no copies from one pixel to the next.
template <int K> struct RowState {
    const uint4* cpf; const uint4* spf; uint4* scur;      // C / S prefetch cursors, S store cursor
    long long dstep;
    const char* pfC; const char* pfS;                     // L2 prefetch cursors: lowest address of the next group of four pixels
    long long pfstep;
    const uint4* ring_in; uint4* ring_out;
    uint16_t* scratch;                                    // final S of the previous pixel, natural disparity order
    volatile int* prog_in; volatile int* prog_me; volatile int* prog_next;
    int seen_in, seen_next;
    int o_m1, o_0, o_p1;                                  // ring slots (uint4 offsets) of columns x-1, x, x+1
    unsigned Nh[4 * K];                                   // normalised state of the horizontal path
    unsigned Cs[4][4 * K];                                // C(x) lives in set x % 4: x and x+1 in use, x+2 and x+3 in flight
    unsigned Ls[2][4 * K];                                // horizontal L(x) in set x % 2 (made one pixel ahead)
    unsigned Vs[2][4 * K];                                // final S(x) in set (x+1) % 2, for the winner-take-all one pixel later
    unsigned Ss[2][4 * K];                                // S of the first sweep: S(x) in set x % 2, reloaded with S(x+2) once read
    unsigned padm[K];
    unsigned wkey, wnb, rkey, rnb;                        // winner-take-all: last evaluation, and this lane's kept record
    unsigned long long* keys_row; int16_t* d1_row;
    unsigned one, P1p, P2mP1p;
    int l;
    // WSG_SW_TMA: staging buffers of this row, their mbarriers, the global cursor of the next chunk to fetch
    uint4* stg; unsigned long long* mbar;
    const char* tmaC; const char* tmaS; long long tma_step;
    int off_even, off_odd;                                // uint4 offset inside a chunk of its first / second logical pixel
    int next_chunk, last_chunk;                           // next chunk to issue; last chunk of the row (W1 / 2)
    unsigned ph;                                          // phase parity of the waits of this group of four pixels
};

// (WSG_SW_TMA) lane 0 fetches chunk st.next_chunk -- logical pixels 2c, 2c+1 of both streams -- into buffer c & 1
template <int K, int MODE> __device__ __forceinline__ void tma_issue(RowState<K>& st)
{
    constexpr int PIX_V = K * 32;
    if (st.next_chunk <= st.last_chunk) {
        __syncwarp();                                     // every lane is done reading the buffer that is overwritten
        if (st.l == 0) {
            const int b = st.next_chunk & 1;
            constexpr unsigned bytes = 2u * PIX_V * 16u;
            fence_proxy_async();
            mbar_expect_tx(st.mbar + b, MODE == 0 ? bytes : 2u * bytes);
            bulk_copy_g2s(st.stg + b * 2 * PIX_V, st.tmaC, bytes, st.mbar + b);
            if (MODE != 0) bulk_copy_g2s(st.stg + (4 + b * 2) * PIX_V, st.tmaS, bytes, st.mbar + b);
        }
    }
    st.tmaC += st.tma_step; st.tmaS += st.tma_step;
    ++st.next_chunk;
}
__device__ __forceinline__ void tma_wait(unsigned long long* bar, unsigned parity, int* err)
{
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1 << 22)) { *err = 1; break; }
    }
}
template <int K> __device__ __forceinline__ void lds_set(unsigned (&dst)[4 * K], const uint4* p, int l)
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint4 c = p[k * 32 + l];
        dst[4 * k] = c.x; dst[4 * k + 1] = c.y; dst[4 * k + 2] = c.z; dst[4 * k + 3] = c.w;
    }
}

template <int K> __device__ __forceinline__ void load_set(unsigned (&dst)[4 * K], const uint4* p)
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint4 c = ldg_stream(p + k * 32);
        dst[4 * k] = c.x; dst[4 * k + 1] = c.y; dst[4 * k + 2] = c.z; dst[4 * k + 3] = c.w;
    }
}

// One pixel of a row; U = x % 4 selects the register sets at compile time.
template <int K, int R, int NS, int MODE, int NDIR, bool HASPAD, bool FAST, int U>
__device__ __forceinline__ void sweep_step(RowState<K>& st, const SweepArgs& a, const int x)
{
    using Cfg = SweepCfg<K, R, NS, MODE>;
    constexpr int NR = 4 * K;
    const int l = st.l;
    unsigned (&Cc)[NR] = st.Cs[U];                  // C(x)
    unsigned (&Cn)[NR] = st.Cs[(U + 1) & 3];        // C(x+1): issued two iterations ago
    unsigned (&Lh)[NR] = st.Ls[U & 1];              // horizontal L(x), made by the previous step
    unsigned (&Lhn)[NR] = st.Ls[(U + 1) & 1];       // horizontal L(x+1), made here
    unsigned (&vsp)[NR] = st.Vs[U & 1];             // S(x-1)
    unsigned (&vsn)[NR] = st.Vs[(U + 1) & 1];       // S(x), made here
    unsigned v[3][NR], Nd[3][NR];
#if WSG_SW_TMA
    {
        constexpr int PIX_V = K * 32;
        // C(x+1) sits in chunk (x+1)/2, buffer ((U+1)/2) & 1; an odd U opens a new chunk: wait for its copy
        if (U & 1) tma_wait(st.mbar + (((U + 1) >> 1) & 1), U == 1 ? st.ph : st.ph ^ 1u, a.err);
        lds_set<K>(st.Cs[(U + 1) & 3], st.stg + (((U + 1) >> 1) & 1) * 2 * PIX_V + (((U + 1) & 1) ? st.off_odd : st.off_even), l);
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < NR; ++j) vsn[j] = Lh[j];
            if (U & 1) tma_issue<K, MODE>(st);             // MODE 0: the chunk of C(x-1), C(x) is free (both are in registers)
        } else {
            unsigned sx[NR];
            lds_set<K>(sx, st.stg + (4 + ((U >> 1) & 1) * 2) * PIX_V + ((U & 1) ? st.off_odd : st.off_even), l);
#pragma unroll
            for (int j = 0; j < NR; ++j) vsn[j] = __viaddmin_u16x2(sx[j], Lh[j], SAT2);
            if (U & 1) tma_issue<K, MODE>(st);             // S(x) was the last read of chunk x/2: refill its buffer with chunk x/2 + 2
        }
        if (U == 3) st.ph ^= 1u;
    }
#else
    if (U == 0) {                                   // pixels x+3+AHEAD .. x+6+AHEAD into L2, one TMA prefetch per stream
        if (l == 0) {
            bulk_prefetch_l2(st.pfC, 4u * K * 512u);
            if (MODE != 0) bulk_prefetch_l2(st.pfS, 4u * K * 512u);
        }
        st.pfC += st.pfstep; st.pfS += st.pfstep;
    }
    load_set<K>(st.Cs[(U + WSG_SW_CA) & 3], st.cpf);   // C(x+CA) into a set that is free (past the row end: unused data)
    st.cpf += st.dstep;
    if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < NR; ++j) vsn[j] = Lh[j];
    } else {
#pragma unroll
        for (int j = 0; j < NR; ++j) vsn[j] = __viaddmin_u16x2(st.Ss[U & 1][j], Lh[j], SAT2);
        load_set<K>(st.Ss[(U + WSG_SW_SA) & 1], st.spf);          // S(x+SA)
        st.spf += st.dstep;
    }
#endif
    if (NDIR == 4) {
        // ---- states of the three directions that come from the row above: columns x-1, x, x+1 of ring r.  Columns -1
        // and W1 exist in the ring as zeros (zero-initialised slot NS-1, and one extra column written by the
        // producer): L = 0 for an out-of-image predecessor.
        wait_prog(st.prog_in, x + 2, st.seen_in, a.err, a.eager);
        const int sl[3] = {st.o_m1, st.o_0, st.o_p1};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const uint4 t = st.ring_in[sl[q] + (q * K + k) * 32];
                Nd[q][4 * k] = t.x; Nd[q][4 * k + 1] = t.y; Nd[q][4 * k + 2] = t.z; Nd[q][4 * k + 3] = t.w;
            }
        }
        // ---- independent chains: the winner-take-all of the PREVIOUS pixel, and the four path steps
        if (MODE == 2) wta_eval<K, HASPAD>(vsp, l, a, st.scratch, st.wkey, st.wnb);
        agg_step<32, NR, HASPAD, FAST>(st.Nh, Cn, Lhn, l, st.P1p, st.P2mP1p, st.padm, st.one);
#pragma unroll
        for (int q = 0; q < 3; ++q) agg_step<32, NR, HASPAD, FAST>(Nd[q], Cc, v[q], l, st.P1p, st.P2mP1p, st.padm, st.one);
        // ---- hand the new states down: slot of column x in ring r+1 is free once its reader has completed x-NS+1
        wait_prog(st.prog_next, x - NS + 2, st.seen_next, a.err, a.eager);
        uint4* dst = st.ring_out + st.o_0;
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int k = 0; k < K; ++k)
                dst[(q * K + k) * 32] = make_uint4(Nd[q][4 * k], Nd[q][4 * k + 1], Nd[q][4 * k + 2], Nd[q][4 * k + 3]);
        __syncwarp();                      // every lane's state stores are issued before the counter store
        asm volatile("" ::: "memory");
        if (l == 0) *st.prog_me = x + 1;
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int j = 0; j < NR; ++j) vsn[j] = __viaddmin_u16x2(vsn[j], v[q][j], SAT2);
        st.o_m1 = st.o_0; st.o_0 = st.o_p1;
        st.o_p1 = st.o_p1 + Cfg::SLOT_V == Cfg::RING_V ? 0 : st.o_p1 + Cfg::SLOT_V;
    } else {
        if (MODE == 2) wta_eval<K, HASPAD>(vsp, l, a, st.scratch, st.wkey, st.wnb);
        agg_step<32, NR, HASPAD, FAST>(st.Nh, Cn, Lhn, l, st.P1p, st.P2mP1p, st.padm, st.one);
    }
    // ---- S out, or kept (registers + shared memory) for the winner-take-all one iteration later
    if (MODE == 2) {
        // the evaluation above was for logical column x-1: lane (x-1)%32 keeps it; every 32 columns all lanes flush
        if (l == ((x - 1) & 31)) { st.rkey = st.wkey; st.rnb = st.wnb; }
        if (x > 0 && (x & 31) == 0) wta_flush(st.rkey, st.rnb, x - 32 + l, true, a, st.keys_row, st.d1_row);
        __syncwarp();                       // all lanes are done reading the previous pixel's S
#pragma unroll
        for (int k = 0; k < K; ++k)      // registers hold (d, d+4) pairs: undo the interleave, so that S(d) sits at scratch[d]
            reinterpret_cast<uint4*>(st.scratch)[l * K + k] =
                make_uint4(__byte_perm(vsn[4 * k], vsn[4 * k + 1], 0x5410), __byte_perm(vsn[4 * k + 2], vsn[4 * k + 3], 0x5410),
                           __byte_perm(vsn[4 * k], vsn[4 * k + 1], 0x7632), __byte_perm(vsn[4 * k + 2], vsn[4 * k + 3], 0x7632));
        __syncwarp();
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k)
            stg_stream(st.scur + k * 32, make_uint4(vsn[4 * k], vsn[4 * k + 1], vsn[4 * k + 2], vsn[4 * k + 3]));
    }
    st.scur += st.dstep;
}

// One image row of a sweep, walked by one warp (see the kernel below for the surrounding protocol).
//   prog[0]     columns the helper has put into ring 0          (ring r = states of the row above row r of the band)
//   prog[r+1]   columns row r has completed and written into ring r+1
//   prog[R+1]   columns the helper has drained from ring R
// MODE 0: S = sum of this sweep's L (no read);  1: S += sum (read-modify-write);  2: S += sum, then WTA (S not written)
// NDIR 4: full sweep;  1: horizontal direction only (the fifth path of MODE_SGBM): rows are independent, no rings.
template <int K, int R, int NS, int MODE, int NDIR, bool HASPAD, bool FAST>
__device__ __forceinline__ void sweep_row(const uint4* __restrict__ C, uint4* __restrict__ S, const SweepArgs& a, uint4* smem,
                                          volatile int* prog, int frame, int band, int r, int l)
{
    using Cfg = SweepCfg<K, R, NS, MODE>;
    constexpr int NR = 4 * K;
    const int W1 = a.W1;
    const int yl = band * a.rows + r;               // logical row (sweep order)
    if (r >= a.rows) return;                        // a launch may use fewer rows per band than the worker has row warps
    if (yl >= a.H) {
        if (NDIR == 4 && l == 0) prog[r + 1] = PROG_INF;         // the row above never waits for this one
        return;
    }
    const int yp = a.flip ? a.H - 1 - yl : yl;      // physical row
    const size_t first = ((size_t)yp * W1 + (a.flip ? W1 - 1 : 0)) * a.Dp8 + l;
    RowState<K> st;
    st.l = l; st.one = a.one; st.P1p = a.P1p; st.P2mP1p = a.P2mP1p;
    st.dstep = a.flip ? -(long long)a.Dp8 : (long long)a.Dp8;
    st.cpf = C + first; st.spf = S + first; st.scur = S + first;
    {   // the group of four pixels x+3+AHEAD .. x+6+AHEAD (logical) starts, in memory, at its first pixel (flip: at its last)
        const long long px0 = 3 + SW_PF_AHEAD + (a.flip ? 3 : 0);
        const long long off = ((long long)yp * W1 + (a.flip ? W1 - 1 - px0 : px0)) * a.Dp8;
        st.pfC = reinterpret_cast<const char*>(C + off);
        st.pfS = reinterpret_cast<const char*>(S + off);
        st.pfstep = (a.flip ? -4ll : 4ll) * a.Dp8 * 16;
    }
    st.ring_in = smem + (size_t)r * Cfg::RING_V + l;
    st.ring_out = smem + (size_t)(r + 1) * Cfg::RING_V + l;
    st.scratch = reinterpret_cast<uint16_t*>(smem + Cfg::RINGS_V + (size_t)r * Cfg::PIX_V);
    st.prog_in = &prog[r]; st.prog_me = &prog[r + 1]; st.prog_next = &prog[r + 2];
    st.seen_in = 0; st.seen_next = 0;
    st.wkey = st.wnb = st.rkey = st.rnb = 0;
    st.keys_row = a.keys + ((size_t)frame * a.H + yp) * a.W;
    st.d1_row = a.d1 + ((size_t)frame * a.H + yp) * a.W;
#pragma unroll
    for (int k = 0; k < K; ++k) st.padm[k] = ((l * K + k) * 8 >= a.D) ? SAT2 : 0u;

#if WSG_SW_TMA
    {   // chunk c = logical pixels 2c, 2c+1 of both streams, fetched by one bulk copy per stream into buffer c & 1; chunks 0
        // and 1 up front, chunk c+2 as soon as chunk c has been read.  A chunk may reach one pixel beyond the row end
        // (the volumes are readable there).  With flip the pair sits in memory in reverse order.
        st.stg = smem + Cfg::RINGS_V + Cfg::SCR_V + (size_t)r * Cfg::STG_V;
        st.mbar = reinterpret_cast<unsigned long long*>(smem + Cfg::RINGS_V + Cfg::SCR_V + Cfg::STGS_V + r);
        st.off_even = a.flip ? Cfg::PIX_V : 0; st.off_odd = a.flip ? 0 : Cfg::PIX_V;
        st.last_chunk = W1 / 2; st.next_chunk = 0; st.ph = 0;
        const long long off0 = ((long long)yp * W1 + (a.flip ? W1 - 2 : 0)) * a.Dp8;
        st.tmaC = reinterpret_cast<const char*>(C + off0);
        st.tmaS = reinterpret_cast<const char*>(S + off0);
        st.tma_step = (a.flip ? -2ll : 2ll) * a.Dp8 * 16;
        if (l == 0) {
            mbar_init(st.mbar, 1); mbar_init(st.mbar + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        tma_issue<K, MODE>(st);
        tma_issue<K, MODE>(st);
        tma_wait(st.mbar, 0, a.err);
        lds_set<K>(st.Cs[0], st.stg + st.off_even, l);
    }
#else
    // C(0), C(1), C(2) and S(0), S(1) up front; every step then issues C(x+3) and S(x+2).  The loads are not guarded at the
    // row end: the volumes are readable a few pixels beyond either end.
#pragma unroll
    for (int i = 0; i < WSG_SW_CA; ++i) { load_set<K>(st.Cs[i], st.cpf); st.cpf += st.dstep; }
    if (MODE != 0) {
#pragma unroll
        for (int i = 0; i < WSG_SW_SA; ++i) { load_set<K>(st.Ss[i], st.spf); st.spf += st.dstep; }
    }
#endif
#pragma unroll
    for (int j = 0; j < NR; ++j) { st.Vs[0][j] = 0; st.Vs[1][j] = 0; st.Nh[j] = 0; }
    // The horizontal direction runs ONE PIXEL AHEAD of the three directions that come from the row above: its step for
    // pixel x+1 and their steps for pixel x are four independent dependency chains in one basic block.
    agg_step<32, NR, HASPAD, FAST>(st.Nh, st.Cs[0], st.Ls[0], l, st.P1p, st.P2mP1p, st.padm, st.one);
    if (NDIR == 4 && Cfg::STAGGER > 0) wait_prog(st.prog_in, min(2 + Cfg::STAGGER, W1 + 1), st.seen_in, a.err, a.eager);
    st.o_m1 = (NS - 1) * Cfg::SLOT_V; st.o_0 = 0; st.o_p1 = Cfg::SLOT_V;      // column -1 is the zero-initialised slot NS-1

    int x = 0;
    for (; x + 3 < W1; x += 4) {
        sweep_step<K, R, NS, MODE, NDIR, HASPAD, FAST, 0>(st, a, x);
        sweep_step<K, R, NS, MODE, NDIR, HASPAD, FAST, 1>(st, a, x + 1);
        sweep_step<K, R, NS, MODE, NDIR, HASPAD, FAST, 2>(st, a, x + 2);
        sweep_step<K, R, NS, MODE, NDIR, HASPAD, FAST, 3>(st, a, x + 3);
    }
    if (x < W1) { sweep_step<K, R, NS, MODE, NDIR, HASPAD, FAST, 0>(st, a, x); ++x; }
    if (x < W1) { sweep_step<K, R, NS, MODE, NDIR, HASPAD, FAST, 1>(st, a, x); ++x; }
    if (x < W1) { sweep_step<K, R, NS, MODE, NDIR, HASPAD, FAST, 2>(st, a, x); ++x; }
    if (MODE == 2) {
        // last column (its S sits in set W1 % 2), then the columns still held in registers: xb .. W1-1, xb = 32*floor((W1-1)/32)
        if (W1 & 1) {
#pragma unroll
            for (int j = 0; j < NR; ++j) st.Vs[0][j] = st.Vs[1][j];
        }
        wta_eval<K, HASPAD>(st.Vs[0], l, a, st.scratch, st.wkey, st.wnb);
        if (l == ((W1 - 1) & 31)) { st.rkey = st.wkey; st.rnb = st.wnb; }
        const int xb = (W1 - 1) & ~31;
        wta_flush(st.rkey, st.rnb, xb + l, xb + l < W1, a, st.keys_row, st.d1_row);
    }
    if (NDIR == 4) {
        // the extra zero column: the out-of-image predecessor of the last pixel's (x+1,y-1) path in the row below
        wait_prog(st.prog_next, W1 - NS + 2, st.seen_next, a.err, a.eager);
#pragma unroll
        for (int j = 0; j < 3 * K; ++j) st.ring_out[st.o_0 + j * 32] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        asm volatile("" ::: "memory");
        if (l == 0) *st.prog_me = W1 + 1;
    }
#if WSG_SW_TMA
    __syncwarp();
    if (l == 0) {        // every issued chunk has been waited for: the barriers can be re-initialised by the next ticket
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(st.mbar)) : "memory");
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(st.mbar + 1)) : "memory");
    }
#endif
}

// The helper warp of a band.  PULL: polls the states the previous band's last row published (global memory, L2: a window
// of HD columns per round trip, its valid prefix is forwarded) into ring 0.  PUSH: drains ring R, the states of this
// band's last row, into the hand-off buffer, epoch-tagged.  Neither side blocks the other: with the bands of several
// frames interleaved a band trails its predecessor by many columns and the pull never has to wait, but a single frame
// is a tight wavefront in which every band follows the one above as closely as the data allows.
template <int K, int R, int NS>
__device__ __forceinline__ void sweep_helper(const SweepArgs& a, uint4* smem, volatile int* prog, int frame, int band, int nbands, int l)
{
    using Cfg = SweepCfg<K, R, NS, 0>;      // the rings come first in every mode's map
    constexpr int NV = 3 * K, HD = Cfg::HD;
    const size_t bstride = (size_t)a.W1 * Cfg::SLOT_V;      // uint4 per boundary row
    uint4* bnd_f = a.bnd + (size_t)frame * a.bnd_v;
    const uint4* src = bnd_f + (size_t)(band > 0 ? band - 1 : 0) * bstride + l;
    uint4* dst = bnd_f + (size_t)band * bstride + l;
    uint4* ring0 = smem + l;
    const int RR = a.rows;                                  // the band's last row is row RR-1: it fills ring RR
    const uint4* ringR = smem + (size_t)RR * Cfg::RING_V + l;
    int pulled = 0, pushed = 0, idle = 0;
    if (band == 0) {                         // image border: ring 0 stays all zeros (L = 0 for out-of-image predecessors)
        if (l == 0) prog[0] = a.W1 + 1;
        pulled = a.W1 + 1;
    }
    if (band + 1 >= nbands) {                // nothing below this band
        if (l == 0) prog[RR + 1] = PROG_INF;
        pushed = a.W1;
    }
    while (pulled <= a.W1 || pushed < a.W1) {
        bool did = false;
        if (pulled < a.W1) {
            // ring 0 slot of column c is free once row 0 has completed column c-NS+1
            const int room = prog[1] + NS - 1 - pulled;
            if (room > 0) {
                const int want = min(min(HD, room), a.W1 - pulled);
                uint4 hb[HD][NV];
#pragma unroll
                for (int u = 0; u < HD; ++u)
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        hb[u][j] = ld_volatile(src + (size_t)min(pulled + u, a.W1 - 1) * Cfg::SLOT_V + j * 32);
                int n = 0;
                bool prefix = true;
#pragma unroll
                for (int u = 0; u < HD; ++u) {
                    bool ok = u < want;
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        ok = ok && ((hb[u][j].x & TAGBITS) == a.tag) && ((hb[u][j].y & TAGBITS) == a.tag) &&
                             ((hb[u][j].z & TAGBITS) == a.tag) && ((hb[u][j].w & TAGBITS) == a.tag);
                    prefix = prefix && __all_sync(FULL, ok);
                    if (prefix) n = u + 1;
                }
                if (n > 0) {
#pragma unroll
                    for (int u = 0; u < HD; ++u) {
                        if (u < n) {
                            uint4* d = ring0 + ((pulled + u) % NS) * Cfg::SLOT_V;
#pragma unroll
                            for (int j = 0; j < NV; ++j)
                                d[j * 32] = make_uint4(hb[u][j].x & ~TAGBITS, hb[u][j].y & ~TAGBITS, hb[u][j].z & ~TAGBITS,
                                                       hb[u][j].w & ~TAGBITS);
                        }
                    }
                    __syncwarp();
                    asm volatile("" ::: "memory");
                    pulled += n;
                    if (l == 0) prog[0] = pulled;
                    did = true;
                }
            }
        } else if (pulled == a.W1) {
            // one more column of zeros: the out-of-image predecessor of the last pixel's (x+1,y-1) path
            if (prog[1] >= a.W1 - NS + 2) {
#pragma unroll
                for (int j = 0; j < NV; ++j) ring0[(a.W1 % NS) * Cfg::SLOT_V + j * 32] = make_uint4(0, 0, 0, 0);
                __syncwarp();
                asm volatile("" ::: "memory");
                pulled = a.W1 + 1;
                if (l == 0) prog[0] = pulled;
                did = true;
            }
        }
        if (pushed < a.W1) {
            const int avail = min((int)prog[RR], a.W1) - pushed;
            if (avail > 0) {
                asm volatile("" ::: "memory");
                const int n = min(avail, 3);
                for (int u = 0; u < n; ++u) {
                    const uint4* s = ringR + ((pushed + u) % NS) * Cfg::SLOT_V;
                    uint4* d = dst + (size_t)(pushed + u) * Cfg::SLOT_V;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        const uint4 t = s[j * 32];
                        st_volatile(d + j * 32, make_uint4(t.x | a.tag, t.y | a.tag, t.z | a.tag, t.w | a.tag));
                    }
                }
                __syncwarp();
                asm volatile("" ::: "memory");
                pushed += n;
                if (l == 0) prog[RR + 1] = pushed;
                did = true;
            }
        }
        if (did) {
            idle = 0;
        } else {
            ++idle;
            if ((idle & 255) == 0 && (idle > (SPIN_LIMIT >> 2) || *reinterpret_cast<volatile int*>(a.err) != 0)) {
                *a.err = 2;
                if (l == 0) { prog[0] = PROG_INF; prog[RR + 1] = PROG_INF; }
                return;
            }
            __nanosleep(idle > 64 ? 200 : 20);
        }
    }
}

template <int K, int R, int NS, int MODE, int NDIR, bool HASPAD>
__global__ void __launch_bounds__(SweepCfg<K, R, NS, MODE>::THREADS, 1)
sweep_kernel(SweepArgs a)
{
    using Cfg = SweepCfg<K, R, NS, MODE>;
    extern __shared__ __align__(16) uint4 smem[];
    __shared__ volatile int prog[R + 2];
    __shared__ int s_ticket;

    const int tid = threadIdx.x, warp = tid >> 5, l = tid & 31;
    const int nbands = (a.H + a.rows - 1) / a.rows;
    const int total = nbands * a.nframes;
    while (true) {
        __syncthreads();                         // the previous band is finished by every warp
        if (tid == 0) s_ticket = atomicAdd(a.ticket, 1);
        if (tid < R + 2) prog[tid] = 0;
        if (NDIR == 4)
            for (int i = tid; i < Cfg::RINGS_V; i += Cfg::THREADS) smem[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const int t = s_ticket;
        if (t >= total) break;
        const int frame = t % a.nframes, band = t / a.nframes;
        if (a.dbg && tid == 0) {
            unsigned sm; unsigned long long ns;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
            a.dbg[3 * t] = (int)sm;
            a.dbg[3 * t + 1] = (int)(ns & 0x7fffffff);
        }
        const uint4* C = a.C + (size_t)frame * a.vol_v;
        uint4* S = a.S + (size_t)frame * a.vol_v;
        // the verified domain (known since the cost kernel ran): cheaper arithmetic
        const bool fast = a.maxC[(size_t)frame * a.scal_stride] + a.P2 <= 32767;
        if (warp == R) {
            if (NDIR == 4) sweep_helper<K, R, NS>(a, smem, prog, frame, band, nbands, l);
        } else if (fast) {
            sweep_row<K, R, NS, MODE, NDIR, HASPAD, true>(C, S, a, smem, prog, frame, band, warp, l);
        } else {
            sweep_row<K, R, NS, MODE, NDIR, HASPAD, false>(C, S, a, smem, prog, frame, band, warp, l);
        }
        if (a.dbg && warp == a.rows - 1 && l == 0) {         // the band's last row is done
            unsigned long long ns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
            a.dbg[3 * t + 2] = (int)(ns & 0x7fffffff);
        }
    }
}

template <int K, int R, int NS, int MODE, int NDIR, bool HASPAD>
static void launch_sweep_t(const SweepArgs& a, int workers, cudaStream_t st)
{
    using Cfg = SweepCfg<K, R, NS, MODE>;
    auto kern = sweep_kernel<K, R, NS, MODE, NDIR, HASPAD>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    const int nbands = (a.H + a.rows - 1) / a.rows;
    const int grid = std::max(1, std::min(workers, nbands * a.nframes));
    kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a);
}

// Rows per band of a launch.  The kernel has RMAX row warps per worker (15 for D <= 256); a launch may use fewer of them.
// More rows per band = more warps per SM (throughput per SM grows like rows^0.6, measured) -- but the bands of a batch are
// handed out in WAVES of one per worker, and a launch whose last wave is mostly empty wastes more than a shorter band
// costs: 8 frames of 2048 rows are 8 x 137 = 7.4 waves of 15-row bands (8 waves of time) but 8 x 147 = 7.95 waves of
// 14-row bands.  Picks the rows in [RMAX-2, RMAX] with the least modelled time.
int sweep_rows_per_band(const SgbmPlan& p, int nframes, int workers)
{
    const int rmax = p.K == 1 ? WSG_SW_ROWS1 : WSG_SW_ROWS2;
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("WSG_SWEEP_ROWS"); forced = e ? atoi(e) : 0; }
    if (forced > 0) return std::min(std::max(forced, 1), rmax);
    // One or two frames are a tight wavefront: every row waits for the row above, so a sweep lasts
    // (2 H + bands x hand-off + W1) columns of one warp's LATENCY per column, which grows with the warps that share an SM.
    // Measured at 2448x2048x256 (profiles/probe_r2_rows_small_batches.txt): 8 rows beat 15 by 7 % (n = 1) and 5 % (n = 2).
    if (nframes <= 2 && p.K == 1) return std::min(rmax, 8);
    int best = rmax;
    double best_t = 1e300;
    for (int r = rmax; r >= std::max(rmax - 2, 1); --r) {
        const long long tickets = (long long)((p.H + r - 1) / r) * std::max(nframes, 1);
        const long long waves = (tickets + workers - 1) / std::max(workers, 1);
        const double t = (double)waves * std::pow((double)r / rmax, 0.4);       // band time ~ rows / rows^0.6
        if (t < best_t * 0.995) { best_t = t; best = r; }
    }
    return best;
}

size_t sweep_boundary_bytes(const SgbmPlan& p, int rows)
{
    const int nbands = (p.H + rows - 1) / rows;
    return (size_t)std::max(nbands - 1, 1) * p.W1 * 3 * p.K * 32 * 16;
}

size_t sweep_volume_pad_bytes() { return 32768; }    // the unguarded prefetches at the row ends reach 3 + 8 + 4 pixels of up to 1 KB beyond

bool sweep_supported(const SgbmPlan& p) { return p.NL == 32 && (p.K == 1 || p.K == 2); }

// mode: 0 first (write S), 1 accumulate (read+write S), 2 accumulate + winner-take-all.  ndir: 4 or 1.
void launch_sweep(const int16_t* C, int16_t* S, int flip, int mode, int ndir, const SgbmPlan& p, const SweepScratch& sc,
                  cudaStream_t st)
{
    SweepArgs a;
    a.H = p.H; a.W1 = p.W1; a.W = p.W; a.D = p.D; a.Dp8 = p.Dp / 8; a.flip = flip;
    a.nframes = std::max(sc.nframes, 1);
    a.C = reinterpret_cast<const uint4*>(C);
    a.S = reinterpret_cast<uint4*>(S);
    a.vol_v = sc.volume_stride_bytes / 16;
    a.P1p = ((unsigned)p.P1 & 0xFFFFu) * 0x10001u;
    a.P2mP1p = ((unsigned)(p.P2 - p.P1) & 0xFFFFu) * 0x10001u;
    a.P2 = p.P2; a.maxC = sc.maxC; a.scal_stride = sc.maxC_stride; a.one = 1u;
    a.tag = ((sc.epoch & 1) ? 0x8000u : 0u) | ((sc.epoch & 2) ? 0x80000000u : 0u);
    a.bnd = reinterpret_cast<uint4*>(sc.boundary);
    a.rows = sc.rows;
    a.bnd_v = sweep_boundary_bytes(p, sc.rows) / 16;
    a.ticket = sc.ticket; a.err = sc.err; a.dbg = sc.dbg;
    static int workers_env = -1;
    if (workers_env < 0) { const char* e = getenv("WSG_SWEEP_WORKERS"); workers_env = e ? atoi(e) : 0; }
    const int workers = std::min(workers_env > 0 ? workers_env : (sc.max_workers > 0 ? sc.max_workers : sc.num_sms), sc.num_sms);
    static int eager = -1;
    if (eager < 0) { const char* e = getenv("WSG_SWEEP_EAGER"); eager = e ? atoi(e) : 64; }
    a.eager = eager;
    a.keys = sc.keys; a.d1 = sc.d1;
    a.minD = p.minD; a.minX1 = p.minX1; a.uniq = p.uniq; a.INVALID = p.INVALID;
    a.umagic = p.uniq < 99 ? (unsigned)((0x100000000ull + (100 - p.uniq) - 1) / (unsigned)(100 - p.uniq)) : 0u;
    a.qm1 = 100 - p.uniq - 1;
    const bool pad = p.Dp != p.D;
#define WSG_SW_CASE(k, r, ns, m, n)                                                  \
    if (p.K == k && mode == m && ndir == n) {                                        \
        if (pad) launch_sweep_t<k, r, ns, m, n, true>(a, workers, st);               \
        else     launch_sweep_t<k, r, ns, m, n, false>(a, workers, st);              \
        return;                                                                      \
    }
    WSG_SW_CASE(1, WSG_SW_ROWS1, WSG_SW_NS1, 0, 4) WSG_SW_CASE(1, WSG_SW_ROWS1, WSG_SW_NS1, 1, 4)
    WSG_SW_CASE(1, WSG_SW_ROWS1, WSG_SW_NS1, 2, 4) WSG_SW_CASE(1, WSG_SW_ROWS1, WSG_SW_NS1, 1, 1)
    WSG_SW_CASE(1, WSG_SW_ROWS1, WSG_SW_NS1, 2, 1)
    WSG_SW_CASE(2, WSG_SW_ROWS2, WSG_SW_NS2, 0, 4) WSG_SW_CASE(2, WSG_SW_ROWS2, WSG_SW_NS2, 1, 4)
    WSG_SW_CASE(2, WSG_SW_ROWS2, WSG_SW_NS2, 2, 4) WSG_SW_CASE(2, WSG_SW_ROWS2, WSG_SW_NS2, 1, 1)
    WSG_SW_CASE(2, WSG_SW_ROWS2, WSG_SW_NS2, 2, 1)
#undef WSG_SW_CASE
}

// ------------------------------------------------------------------------------------------------
// Around the fused WTA: reset of the right-view keys / left-view map, and the LR check (A.6).
// ------------------------------------------------------------------------------------------------
__global__ void wta_reset_kernel(unsigned long long* __restrict__ keys, int16_t* __restrict__ d1, size_t n, int invalid)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = ~0ull; d1[i] = (int16_t)invalid; }
}

__global__ void lrcheck_kernel(const unsigned long long* __restrict__ keys, const int16_t* __restrict__ d1,
                               int16_t* __restrict__ raw, SgbmPlan p)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= p.W) return;
    const unsigned long long* krow = keys + (size_t)y * p.W;
    int dv = d1[(size_t)y * p.W + x];
    if (x >= p.minX1 && x < p.maxX1 && dv != p.INVALID) {
        const int a = dv >> 4, b = (dv + 15) >> 4;
        const int xa = x - a, xb = x - b;
        bool ca = false, cb = false;
        if (xa >= 0 && xa < p.W) {
            const unsigned long long k = krow[xa];
            const int d2 = (k == ~0ull) ? p.INVALID : (int)(k & 0xFFFFu) + p.minD;
            ca = d2 >= p.minD && abs(d2 - a) > p.d12;
        }
        if (xb >= 0 && xb < p.W) {
            const unsigned long long k = krow[xb];
            const int d2 = (k == ~0ull) ? p.INVALID : (int)(k & 0xFFFFu) + p.minD;
            cb = d2 >= p.minD && abs(d2 - b) > p.d12;
        }
        if (ca && cb) dv = p.INVALID;
    }
    raw[(size_t)y * p.W + x] = (int16_t)dv;
}

// keys / d1 of `nframes` frames are contiguous: one launch resets them all
void launch_wta_reset(const SweepScratch& sc, const SgbmPlan& p, cudaStream_t st)
{
    const size_t n = (size_t)p.H * p.W * std::max(sc.nframes, 1);
    wta_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sc.keys, sc.d1, n, p.INVALID);
}

// frame f of the batch: keys / d1 at f * H * W
void launch_lrcheck(const SweepScratch& sc, int frame, int16_t* raw, const SgbmPlan& p, cudaStream_t st)
{
    dim3 b(256), g((p.W + 255) / 256, p.H);
    const size_t off = (size_t)frame * p.H * p.W;
    lrcheck_kernel<<<g, b, 0, st>>>(sc.keys + off, sc.d1 + off, raw, p);
}

}  // namespace wsg
