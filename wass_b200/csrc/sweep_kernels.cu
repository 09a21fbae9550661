// Fused path aggregation of cv::StereoSGBM (call site
// src/wass_stereo/wass_stereo.cpp:837): one launch per SWEEP instead of one per direction.
//
// A sweep runs the four directions whose predecessors are (x-1,y) (x-1,y-1) (x,y-1) (x+1,y-1)
// [sweep 1: r0..r3] -- or, rotated by 180 degrees, (x+1,y) (x+1,y+1) (x,y+1) (x-1,y+1)
// [sweep 2: r4,r7,r6,r5] -- as a skew-2 wavefront: one warp walks one image row, row y+1 trails
// row y by two columns, so every predecessor state already exists when a pixel is reached.  The
// cost vector C(p,.) is read from HBM exactly once per sweep, the four L_r(p,.) are summed in
// registers, and S is written once (sweep 1) or consumed on the spot by the winner-take-all
// stage (sweep 2): 4V of HBM traffic for MODE_HH instead of the 23V of eight separate launches.
//
//   CTA      = R consecutive rows (one warp each) + one helper warp
//   row->row = normalised states N_r = L_r - min L_r handed down through a shared-memory ring
//              (NS columns deep, progress counters instead of CTA barriers: warps drift freely)
//   CTA->CTA = the last row of a band publishes its states to a global boundary buffer; every
//              32-bit word carries a 2-bit epoch tag in the two sign bits a state never uses, so
//              the consumer validates data word by word without fences or flags; the helper warp
//              of the next band polls them (L2) into that CTA's ring 0
//   order    = bands are handed out by an atomic ticket, so a band only ever waits for bands that
//              are already running: no co-residency assumption, no deadlock
//
// All spin loops are bounded: on overrun the kernel raises an error flag and runs to completion.
#include "sgbm_dev.cuh"
#include <algorithm>
#include <cstdlib>

namespace wsg {

static constexpr int SW_R = 7;             // rows (compute warps) per CTA: with the helper warp 8 warps, two per scheduler
                                           // (both sweeps, B200: R=4 10.1, 5 9.9, 6 9.0, 7 8.36, 8 8.58, 10 9.4, 12 9.9, 16 11.7 ms)
static constexpr int SW_THREADS = (SW_R + 1) * 32;
static constexpr unsigned TAGBITS = 0x80008000u;
static constexpr int SPIN_LIMIT = 1 << 22;


// Shared-memory budget of a CTA (K == 1; K == 2 always takes one SM to itself).
//   CFG 0 (default)  104 KB of rings and staging, padded to 120 KB: one sweep CTA per SM, and room beside it for a CTA of
//                    the cost kernel of ANOTHER frame (92 KB) -- measured best both for one frame at a time and for two
//                    frames in flight (bench.py --pipeline-depth 2)
//   CFG 1            200 KB: deeper rings and prefetch; same speed (neither depth is the limit)
//   CFG 2            104 KB unpadded: two sweep CTAs of different frames may share an SM; no faster, because the
//                    plateau of a sweep already keeps the integer ALU of its SM ~80 % busy
template <int K, int CFG> struct SweepCfg {
    static constexpr bool LEAN = K == 1 && CFG != 1;
    static constexpr int NS = K != 1 ? 6 : (LEAN ? 5 : 8);    // state ring depth in columns
    static constexpr int SLOT_V = 3 * K * 32;                 // uint4 per ring slot: [dir][k][lane]
    static constexpr int RING_V = NS * SLOT_V;
    static constexpr int PFD = K != 1 ? 4 : (CFG == 1 ? 12 : 6);   // pixels of C in flight per row (cp.async staging)
    static constexpr int PFS = K == 1 && CFG == 1 ? 12 : 4;   // pixels of S in flight per row
    static constexpr int PIX_V = K * 32;                      // uint4 per pixel
    static constexpr int RINGS_V = SW_R * RING_V;             // smem map (uint4 units): state rings
    static constexpr int SCR_V = SW_R * PIX_V;                //   per-warp WTA scratch
    static constexpr int STAGEC_V = SW_R * PFD * PIX_V;       //   per-warp staging of the C stream
    static constexpr int STAGES_V = SW_R * PFS * PIX_V;       //   per-warp staging of the S stream
    static constexpr int SMEM_USED = (RINGS_V + SCR_V + STAGEC_V + STAGES_V) * 16;
    static constexpr int CTAS_PER_SM = K == 1 && CFG == 2 ? 2 : 1;
    static constexpr int SMEM = CTAS_PER_SM == 1 && SMEM_USED < 120 * 1024 ? 120 * 1024 : SMEM_USED;   // (forces 1 CTA/SM)
    // boundary columns the helper polls per round trip.  At most NS-2: it may only overwrite ring-0 slots of columns
    // its consumer has completed, and the consumer can complete nothing beyond the columns already published.
    static constexpr int HD = NS - 2 < 4 ? NS - 2 : 4;
};

// 16-byte asynchronous global->shared copy (L2 only), one per lane; completion is tracked per thread in commit groups
__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct SweepArgs {
    int H, W1, W, D;
    int Dp8;                 // 16-byte vectors per pixel (= 32*K)
    int flip;                // 0: top->bottom, left->right;  1: rotated by 180 degrees
    unsigned P1p, P2mP1p;
    int P2;
    const int* maxC;         // max over the cost volume (written by the cost kernel)
    unsigned one;            // 1 (kept opaque to the compiler: see add_on_fma)
    unsigned tag;            // epoch tag of this launch (bits 15 and 31)
    uint4* bnd;              // [nbands-1][W1][3][K][32]
    int* ticket;             // [0] bands handed out, [1] workers elected, [2..] CTAs of this launch arrived per SM
    int num_sms;
    int max_workers;         // at most this many SMs work on the sweep (the others stay free for another frame's kernels)
    int eager;               // polls of a progress counter before the waiter starts to sleep between polls
    int* err;
    int* dbg;                // optional: SM id of every band (placement diagnostics), or null
    // winner-take-all (MODE 2)
    unsigned long long* keys;   // [H][W]  (minS, W1-1-x, d) of the best match that lands on x2 (A.5)
    int16_t* d1;                // [H][W]  left-view disparity before the LR check
    int minD, minX1, uniq, INVALID;
    unsigned umagic;            // ceil(2^32 / (100 - uniq)), or 0 when 100 - uniq == 1
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

// ---- thread-block cluster / distributed shared memory (CTA pairs: the even band hands its last row's states straight into
// the ring of the odd band on the neighbouring SM instead of going through the hand-off buffer in L2)
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned dsmem_addr(const void* local_smem, unsigned rank)
{
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"((unsigned)__cvta_generic_to_shared(local_smem)), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_st_v4(unsigned addr, const uint4& v)
{
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void dsmem_st_release_u32(unsigned addr, int v)   // (a plain volatile store: see WSG_SWEEP_PAIRS)
{
    asm volatile("st.volatile.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int dsmem_ld_u32(unsigned addr)
{
    int v;
    asm volatile("ld.volatile.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// wait_prog on a counter that lives in the shared memory of another CTA of the cluster
__device__ __forceinline__ void wait_prog_remote(unsigned addr, int need, int& seen, int* err)
{
    if (seen >= need) return;
    int spins = 0;
    while ((seen = dsmem_ld_u32(addr)) < need) {
        if (++spins > 16) __nanosleep(100);
        if ((spins & 1023) == 0 && (spins > SPIN_LIMIT || *reinterpret_cast<volatile int*>(err) != 0)) { *err = 4; seen = 0x7fffffff; break; }
    }
}

// Wait until *flag >= need (shared-memory progress counter of a neighbouring warp).  Bounded: on overrun (or when
// any other waiter has already given up) raise the error flag and stop waiting for good.
__device__ __forceinline__ void wait_prog(volatile int* flag, int need, int& seen, int* err, int eager_spins = 64)
{
    if (seen < need) {
        int spins = 0;
        while ((seen = *flag) < need) {
            // A warp that is not yet due (or is held back by back-pressure: small eager_spins) must not steal issue slots.
            // Measured alternatives, all ~5 % SLOWER: a tight load/compare/branch poll (it issues more often than this
            // loop and takes slots from the warps that do the work), the same with a 32 ns sleep per poll, and sleeping
            // sooner (a sleep on the band-to-band critical path costs more than the polling it saves).
            if (++spins > eager_spins) __nanosleep(spins > 4096 ? 400 : (eager_spins > 1 ? 40 : 150));
            if ((spins & 1023) == 0 && (spins > SPIN_LIMIT || *reinterpret_cast<volatile int*>(err) != 0)) {
                *err = 1;
                seen = 0x7fffffff;
                break;
            }
        }
    }
    // Shared-memory requests of one warp are served in issue order, so a counter read that saw the value is followed
    // by data reads that see the data; only the compiler has to be kept from moving them.
    asm volatile("" ::: "memory");
}

// ---- column publication through mbarriers (compile with -DWSG_SWEEP_MBAR=1; OFF by default: measured slower) ----------
// A row that has caught up with the row above POLLS that row's progress counter: ~11 polls per pixel, ~40 % of the
// kernel's executed instructions.  The experiment: each row also owns MBN one-arrival mbarriers; publishing column x
// completes one phase of barrier x % MBN, and the consumer's mbarrier.try_wait suspends the warp in hardware until that
// phase is over (the producer is never more than NS-2 <= MBN-2 columns ahead, so a waiter is never more than one phase
// behind and the parity test is unambiguous).  Result on B200, bit-exact: both sweeps 8.6 -> 9.6 ms.  Waking a suspended
// warp costs more than the polls it saves -- every one of the 2 H row-to-row hand-offs is on the critical path of the
// wavefront -- the same outcome as nanosleep back-off in wait_prog.
#ifndef WSG_SWEEP_MBAR
#define WSG_SWEEP_MBAR 0
#endif
static constexpr int MBN = 8;
__device__ __forceinline__ void mbar_init(unsigned addr, int count)
{
    asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned addr) { asm volatile("mbarrier.inval.shared.b64 [%0];" ::"r"(addr) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned addr)
{
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared.b64 st, [%0];\n\t}" ::"r"(addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned addr, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
    return ok != 0;
}
// Wait until column c of the row whose barriers start at `bars` is published.  Bounded like wait_prog.
__device__ __forceinline__ void wait_col(unsigned bars, int c, bool& dead, int* err)
{
    if (dead) return;
    const unsigned addr = bars + (unsigned)(c & (MBN - 1)) * 8u, parity = (unsigned)(c >> 3) & 1u;
    if (mbar_try_wait(addr, parity)) return;
    int spins = 0;
    while (!mbar_try_wait(addr, parity)) {
        if ((++spins & 15) == 0 && (spins > (SPIN_LIMIT >> 6) || *reinterpret_cast<volatile int*>(err) != 0)) { *err = 5; dead = true; break; }
    }
}
// (re)arm the barriers of all rows of a band; called by the whole CTA between two __syncthreads
__device__ __forceinline__ void mbar_reset_all(unsigned long long* bars, int tid, bool first)
{
    if (tid < (SW_R + 1) * MBN) {
        const unsigned addr = (unsigned)__cvta_generic_to_shared(bars + tid);
        if (!first) mbar_inval(addr);
        mbar_init(addr, 1);
    }
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
// 0 <= uniquenessRatio < 100 (the host routes anything else to wta_kernel).
template <int K, bool HASPAD>
__device__ __forceinline__ void wta_eval(const unsigned (&s)[4 * K], int l, const SweepArgs& a, const int16_t* scratch,
                                         unsigned& key, unsigned& nb)
{
    constexpr int NV8 = 8 * K;
    const int dlane = l * NV8;
    const bool padlane = HASPAD && dlane >= a.D;
    unsigned kmin = 0xFFFFFFFFu;
#pragma unroll
    for (int e = 0; e < 4 * K; ++e) {   // register 4k+i holds disparities (8k+i, 8k+4+i) of the lane, see vec_pos
        const int d = dlane + 8 * (e >> 2) + (e & 3);
        const unsigned klo = s[e] * 65536u + (unsigned)d;
        const unsigned khi = (s[e] & 0xFFFF0000u) | (unsigned)(d + 4);
        kmin = __vimin3_u32(kmin, klo, khi);
    }
    if (padlane) kmin = 0xFFFFFFFFu;
    kmin = __reduce_min_sync(FULL, kmin);
    const int minS = (int)(kmin >> 16), best = (int)(kmin & 0xFFFFu);
    const int n = minS * 100 - 1;                                   // < 2^22
    // floor(n / (100-uniq)): multiply-high by ceil(2^32/(100-uniq)) is exact for n < 2^22; a divisor of 1 has no such constant
    const int Tm = n < 0 ? -1 : (a.umagic ? (int)__umulhi((unsigned)n, a.umagic) : n);
    // packed count of S <= Tm: (0x8000 + T - S) keeps bit 15 iff T >= S (S, T <= 0x7fff: no borrow between the halves)
    const unsigned T2 = ((unsigned)min(max(Tm, 0), 32767) | 0x8000u) * 0x10001u;
    unsigned bits = 0;
#pragma unroll
    for (int e = 0; e < 4 * K; ++e) bits |= ((T2 - s[e]) & 0x80008000u) >> e;
    int cnt = __popc(bits);
    if (padlane || Tm < 0) cnt = 0;
    const int total = __reduce_add_sync(FULL, cnt);
    // the winner's neighbours (clamped addresses; the values only count where they exist)
    const int dm = max(best - 1, 0), dp = min(best + 1, a.D - 1);
    const int sm = scratch[(dm & ~7) + vec_pos(dm & 7)], sp = scratch[(dp & ~7) + vec_pos(dp & 7)];
    const int inwin = (minS <= Tm) + (best > 0 && sm <= Tm) + (best < a.D - 1 && sp <= Tm);
    key = kmin;
    nb = (unsigned)sm | ((unsigned)sp << 16) | (total <= inwin ? 0x80000000u : 0u);
}

// wta_flush (every 32 pixels): lane i holds the record of logical column xbase + i.  Sub-pixel parabola, right-view map
// (A.5) and the left-view map, all 32 lanes busy on 32 different pixels; the d1 row segment is one coalesced store.
__device__ __forceinline__ void wta_flush(unsigned key, unsigned nb, int xl, bool valid, const SweepArgs& a,
                                          unsigned long long* keys_row, int16_t* d1_row)
{
    if (!valid || !(nb & 0x80000000u)) return;
    const int minS = (int)(key >> 16), best = (int)(key & 0xFFFFu);
    const int sm = (int)(nb & 0xFFFFu), sp = (int)((nb >> 16) & 0x7FFFu);
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

// One image row of a sweep, walked by one warp (see the kernel below for the surrounding protocol).
template <int K, int MODE, int NDIR, bool HASPAD, int CFG, bool FAST>
__device__ __forceinline__ void sweep_row(const uint4* __restrict__ C, uint4* __restrict__ S, const SweepArgs& a, uint4* smem,
                                          volatile int* prog, int band, int warp, int l, bool to_peer, bool from_peer_cta = false,
                                          unsigned bars = 0)
{
    using Cfg = SweepCfg<K, CFG>;
    constexpr int NR = 4 * K;
    constexpr int NS = Cfg::NS;
    const size_t bstride = (size_t)a.W1 * Cfg::SLOT_V;
    const unsigned one = a.one;
    constexpr int NQ = NDIR == 4 ? 3 : 2;            // directions whose predecessor lies in the row above
    const int r = warp;
    const int yl = band * SW_R + r;                 // logical row (sweep order)
    if (yl >= a.H) return;
    const bool from_peer = from_peer_cta && r == 0;
    const bool top = NDIR == 1 || yl == 0;          // no predecessor row
    // 0: no row below; 1: ring of the next warp; 2: hand-off buffer in global memory; 3: ring 0 of the peer CTA (DSMEM)
    const int out_mode = (NDIR == 1 || yl == a.H - 1) ? 0 : (r < SW_R - 1 ? 1 : (to_peer ? 3 : 2));
    const int yp = a.flip ? a.H - 1 - yl : yl;      // physical row
    const long long dstep = a.flip ? -(long long)a.Dp8 : (long long)a.Dp8;
    const size_t first = ((size_t)yp * a.W1 + (a.flip ? a.W1 - 1 : 0)) * a.Dp8 + l;
    const uint4* cpf = C + first;                   // prefetch cursors
    const uint4* spf = S + first;
    uint4* scur = S + first;                        // compute cursor
    const uint4* ring_in = smem + (size_t)r * Cfg::RING_V + l;
    uint4* ring_out = smem + (size_t)(r + 1) * Cfg::RING_V + l;     // only used when out_mode == 1
    uint4* bnd_out = a.bnd + (size_t)band * bstride + l;            // only used when out_mode == 2
    int16_t* scratch = reinterpret_cast<int16_t*>(smem + Cfg::RINGS_V + (size_t)r * Cfg::PIX_V);
    uint4* stageC = smem + Cfg::RINGS_V + Cfg::SCR_V + (size_t)r * Cfg::PFD * Cfg::PIX_V + l;
    uint4* stageS = smem + Cfg::RINGS_V + Cfg::SCR_V + Cfg::STAGEC_V + (size_t)r * Cfg::PFS * Cfg::PIX_V + l;
    volatile int* prog_in = &prog[r];
    volatile int* prog_me = &prog[r + 1];
    volatile int* prog_next = &prog[r + 2 <= SW_R ? r + 2 : SW_R];
    int seen_in = 0, seen_next = 0;
    const bool use_mbar = WSG_SWEEP_MBAR && bars != 0 && !from_peer;        // (a peer CTA publishes through DSMEM counters only)
    const unsigned bars_in = bars + (unsigned)r * (MBN * 8), bars_me = bars + (unsigned)(r + 1) * (MBN * 8);
    bool dead = false;
    // peer CTA (cluster rank 1): its ring 0, its prog[0] (which this warp advances) and prog[1] (its first row: back-pressure)
    unsigned peer_ring = 0, peer_prog0 = 0, peer_prog1 = 0;
    if (out_mode == 3) {
        peer_ring = dsmem_addr(smem + l, 1);
        peer_prog0 = dsmem_addr(const_cast<int*>(&prog[0]), 1);
        peer_prog1 = dsmem_addr(const_cast<int*>(&prog[1]), 1);
    }

    unsigned padm[K];
#pragma unroll
    for (int k = 0; k < K; ++k) padm[k] = ((l * K + k) * 8 >= a.D) ? SAT2 : 0u;

    // one commit group per pixel, PFD pixels ahead; each lane copies and later reads back its own 16 bytes
    const unsigned stC = (unsigned)__cvta_generic_to_shared(stageC), stS = (unsigned)__cvta_generic_to_shared(stageS);
    static_assert(Cfg::PFS <= Cfg::PFD, "the S stream rides in the commit groups of the C stream");
    for (int i = 0; i < Cfg::PFD; ++i) {
        if (i < a.W1) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                cp_async16(stC + (i * Cfg::PIX_V + k * 32) * 16, cpf + k * 32);
                if (MODE != 0 && i < Cfg::PFS) cp_async16(stS + (i * Cfg::PIX_V + k * 32) * 16, spf + k * 32);
            }
        }
        cp_async_commit();
        cpf += dstep;
        if (i < Cfg::PFS) spf += dstep;
    }

    // The horizontal direction runs ONE PIXEL AHEAD of the three directions that come from the row above: its step for
    // pixel x+1 and their steps for pixel x are four independent dependency chains in one basic block.
    unsigned Nh[NR], Cc[NR], Lh[NR], vsp[NR];
    unsigned wkey = 0, wnb = 0, rkey = 0, rnb = 0;      // winner-take-all: last evaluation, and this lane's kept record
    unsigned long long* keys_row = a.keys + (size_t)yp * a.W;
    int16_t* d1_row = a.d1 + (size_t)yp * a.W;
#pragma unroll
    for (int j = 0; j < NR; ++j) vsp[j] = 0;
#pragma unroll
    for (int j = 0; j < NR; ++j) Nh[j] = 0;
    cp_async_wait<Cfg::PFD - 1>();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint4 c = stageC[k * 32];
        Cc[4 * k] = c.x; Cc[4 * k + 1] = c.y; Cc[4 * k + 2] = c.z; Cc[4 * k + 3] = c.w;
    }
    agg_step<32, NR, HASPAD, FAST>(Nh, Cc, Lh, l, a.P1p, a.P2mP1p, padm, one);

    int pslot = 0, sslot = 0;                       // x % PFD, x % PFS
    for (int x = 0; x < a.W1; ++x) {                // logical column
        const int nslot = pslot + 1 == Cfg::PFD ? 0 : pslot + 1;
        unsigned Cn[NR], Lhn[NR], vs[NR], v[3][NR], Nd[3][NR];
        cp_async_wait<Cfg::PFD - 2>();              // pixel x+1 has landed (past the row end: a stale slot, result unused)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint4 c = stageC[nslot * Cfg::PIX_V + k * 32];
            Cn[4 * k] = c.x; Cn[4 * k + 1] = c.y; Cn[4 * k + 2] = c.z; Cn[4 * k + 3] = c.w;
        }
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < NR; ++j) vs[j] = Lh[j];
        } else {
            // S(x) was committed PFS iterations ago; PFD + x groups exist by now
            if (Cfg::PFS < Cfg::PFD - 1) cp_async_wait<Cfg::PFS - 1>();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const uint4 sv = stageS[sslot * Cfg::PIX_V + k * 32];
                vs[4 * k] = sat_add_split(sv.x, Lh[4 * k], one);
                vs[4 * k + 1] = sat_add_split(sv.y, Lh[4 * k + 1], one);
                vs[4 * k + 2] = sat_add_split(sv.z, Lh[4 * k + 2], one);
                vs[4 * k + 3] = sat_add_split(sv.w, Lh[4 * k + 3], one);
            }
        }
        if (NDIR >= 3) {
            // ---- states of the NQ directions that come from the row above (3 with, 2 without the (x+1,y-1) path)
            if (top) {
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int j = 0; j < NR; ++j) Nd[q][j] = 0;
            } else {
                // columns -1 and W1 of the row above exist in the ring as zeros (zero-initialised slot NS-1, and
                // one extra column written by the producer): L = 0 for an out-of-image predecessor
                // NDIR 4 needs column x+1 of the row above (skew 2), NDIR 3 only column x (skew 1)
                if (use_mbar) wait_col(bars_in, x + NQ - 2, dead, a.err);
                else wait_prog(prog_in, x + NQ - 1, seen_in, a.err, a.eager);
                const int sl[3] = {(x + NS - 1) % NS, x % NS, (x + 1) % NS};
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const uint4 t = ring_in[sl[q] * Cfg::SLOT_V + (q * K + k) * 32];
                        Nd[q][4 * k] = t.x; Nd[q][4 * k + 1] = t.y; Nd[q][4 * k + 2] = t.z; Nd[q][4 * k + 3] = t.w;
                    }
                }
            }
            // ---- independent chains: the winner-take-all of the PREVIOUS pixel, and the four path steps
            if (MODE == 2) wta_eval<K, HASPAD>(vsp, l, a, scratch, wkey, wnb);
            agg_step<32, NR, HASPAD, FAST>(Nh, Cn, Lhn, l, a.P1p, a.P2mP1p, padm, one);
#pragma unroll
            for (int q = 0; q < NQ; ++q) agg_step<32, NR, HASPAD, FAST>(Nd[q], Cc, v[q], l, a.P1p, a.P2mP1p, padm, one);
            // ---- hand the new states down
            if (out_mode == 1) {
                wait_prog(prog_next, x - NS + 2, seen_next, a.err, a.eager);
                uint4* dst = ring_out + (x % NS) * Cfg::SLOT_V;
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        dst[(q * K + k) * 32] = make_uint4(Nd[q][4 * k], Nd[q][4 * k + 1], Nd[q][4 * k + 2], Nd[q][4 * k + 3]);
            } else if (out_mode == 2) {
                uint4* dst = bnd_out + (size_t)x * Cfg::SLOT_V;
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        st_volatile(dst + (q * K + k) * 32,
                                    make_uint4(Nd[q][4 * k] | a.tag, Nd[q][4 * k + 1] | a.tag, Nd[q][4 * k + 2] | a.tag,
                                               Nd[q][4 * k + 3] | a.tag));
            } else if (out_mode == 3) {
                wait_prog_remote(peer_prog1, x - NS + 2, seen_next, a.err);
                const unsigned dst = peer_ring + ((x % NS) * Cfg::SLOT_V) * 16;
#pragma unroll
                for (int q = 0; q < NQ; ++q)
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        dsmem_st_v4(dst + ((q * K + k) * 32) * 16, make_uint4(Nd[q][4 * k], Nd[q][4 * k + 1], Nd[q][4 * k + 2], Nd[q][4 * k + 3]));
                __syncwarp();
                if (l == 0) dsmem_st_release_u32(peer_prog0, x + 1);
            }
            __syncwarp();                      // every lane's state stores are issued before the counter store
            asm volatile("" ::: "memory");
            if (l == 0) {
                *prog_me = x + 1;
                if (WSG_SWEEP_MBAR && bars != 0 && out_mode == 1) mbar_arrive(bars_me + (unsigned)(x & (MBN - 1)) * 8u);
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
                for (int j = 0; j < NR; ++j) vs[j] = sat_add_split(vs[j], v[q][j], one);
        } else {
            if (MODE == 2) wta_eval<K, HASPAD>(vsp, l, a, scratch, wkey, wnb);
            agg_step<32, NR, HASPAD, FAST>(Nh, Cn, Lhn, l, a.P1p, a.P2mP1p, padm, one);
        }
        // ---- S out, or kept (registers + shared memory) for the winner-take-all one iteration later
        if (MODE == 2) {
            // the evaluation above was for logical column x-1: lane (x-1)%32 keeps it; every 32 columns all lanes flush
            if (l == ((x - 1) & 31)) { rkey = wkey; rnb = wnb; }
            if (x > 0 && (x & 31) == 0) wta_flush(rkey, rnb, x - 32 + l, true, a, keys_row, d1_row);
            __syncwarp();                       // all lanes are done reading the previous pixel's S
#pragma unroll
            for (int k = 0; k < K; ++k)
                reinterpret_cast<uint4*>(scratch)[l * K + k] = make_uint4(vs[4 * k], vs[4 * k + 1], vs[4 * k + 2], vs[4 * k + 3]);
#pragma unroll
            for (int j = 0; j < NR; ++j) vsp[j] = vs[j];
            __syncwarp();
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k)
                stg_stream(scur + k * 32, make_uint4(vs[4 * k], vs[4 * k + 1], vs[4 * k + 2], vs[4 * k + 3]));
        }
        scur += dstep;
        // ---- refill the staging slots just consumed with pixels x + PFD (C) and x + PFS (S)
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (x + Cfg::PFD < a.W1) cp_async16(stC + (pslot * Cfg::PIX_V + k * 32) * 16, cpf + k * 32);
            if (MODE != 0 && x + Cfg::PFS < a.W1) cp_async16(stS + (sslot * Cfg::PIX_V + k * 32) * 16, spf + k * 32);
        }
        cp_async_commit();
        cpf += dstep; spf += dstep;
        pslot = nslot;
        sslot = sslot + 1 == Cfg::PFS ? 0 : sslot + 1;
#pragma unroll
        for (int j = 0; j < NR; ++j) { Cc[j] = Cn[j]; Lh[j] = Lhn[j]; }
    }
    cp_async_wait<0>();
    if (MODE == 2) {
        // last column, then the columns still held in registers: xb .. W1-1 with xb = 32*floor((W1-1)/32)
        wta_eval<K, HASPAD>(vsp, l, a, scratch, wkey, wnb);
        if (l == ((a.W1 - 1) & 31)) { rkey = wkey; rnb = wnb; }
        const int xb = (a.W1 - 1) & ~31;
        wta_flush(rkey, rnb, xb + l, xb + l < a.W1, a, keys_row, d1_row);
    }
    if (NDIR >= 3 && out_mode == 1) {
        // the extra zero column (see above)
        wait_prog(prog_next, a.W1 - NS + 2, seen_next, a.err, a.eager);
#pragma unroll
        for (int j = 0; j < 3 * K; ++j) ring_out[(a.W1 % NS) * Cfg::SLOT_V + j * 32] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        asm volatile("" ::: "memory");
        if (l == 0) {
            *prog_me = a.W1 + 1;
            if (WSG_SWEEP_MBAR && bars != 0) mbar_arrive(bars_me + (unsigned)(a.W1 & (MBN - 1)) * 8u);
        }
    }
    if (NDIR >= 3 && out_mode == 3) {
        wait_prog_remote(peer_prog1, a.W1 - NS + 2, seen_next, a.err);
#pragma unroll
        for (int j = 0; j < 3 * K; ++j) dsmem_st_v4(peer_ring + ((a.W1 % NS) * Cfg::SLOT_V + j * 32) * 16, make_uint4(0, 0, 0, 0));
        __syncwarp();
        if (l == 0) dsmem_st_release_u32(peer_prog0, a.W1 + 1);
    }
}


// MODE 0: S = sum of this sweep's L (no read);  1: S += sum (read-modify-write);  2: S += sum, then WTA (S not written)
// NDIR 4: full sweep (skew 2);  3: without the (x+1,y-1) path (skew 1: rows trail each other by ONE column, which nearly
// halves the wavefront's critical path; that path then runs as its own HBM-bound launch);  1: horizontal direction only
// (the fifth path of MODE_SGBM): rows are independent.
// The helper warp of a band: polls the states the previous band's last row published (global memory, L2) into ring 0.
// MODE 0: S = sum of this sweep's L (no read);  1: S += sum (read-modify-write);  2: S += sum, then WTA (S not written)
// NDIR 4: full sweep (skew 2);  3: without the (x+1,y-1) path (skew 1: rows trail each other by ONE column, which nearly
// halves the wavefront's critical path; that path then runs as its own HBM-bound launch);  1: horizontal direction only
// (the fifth path of MODE_SGBM): rows are independent.
// The helper warp of a band: polls the states the previous band's last row published (global memory, L2) into ring 0.
template <int K, int CFG, int NQ>
__device__ __forceinline__ void sweep_helper(const SweepArgs& a, uint4* smem, volatile int* prog, int band, int l, unsigned bars = 0)
{
    using Cfg = SweepCfg<K, CFG>;
    constexpr int NS = Cfg::NS;
    const size_t bstride = (size_t)a.W1 * Cfg::SLOT_V;      // uint4 per boundary row
    {
        if (band == 0) return;
        const uint4* src = a.bnd + (size_t)(band - 1) * bstride + l;
        uint4* ring = smem;                                // ring 0
        uint4 hb[Cfg::HD][NQ * K];
        int seen = 0, spins = 0;
        // Poll a window of Cfg::HD columns per round trip to L2 and forward its valid prefix: the throughput adapts to the
        // producer (up to Cfg::HD columns per round trip) and the band ends up trailing it by about one window.
        for (int x = 0; x < a.W1;) {
#pragma unroll
            for (int u = 0; u < Cfg::HD; ++u)
#pragma unroll
                for (int j = 0; j < NQ * K; ++j)
                    hb[u][j] = ld_volatile(src + (size_t)min(x + u, a.W1 - 1) * Cfg::SLOT_V + j * 32);
            int n = 0;
            bool prefix = true;
#pragma unroll
            for (int u = 0; u < Cfg::HD; ++u) {
                bool ok = x + u < a.W1;
#pragma unroll
                for (int j = 0; j < NQ * K; ++j)
                    ok = ok && ((hb[u][j].x & TAGBITS) == a.tag) && ((hb[u][j].y & TAGBITS) == a.tag) &&
                         ((hb[u][j].z & TAGBITS) == a.tag) && ((hb[u][j].w & TAGBITS) == a.tag);
                prefix = prefix && __all_sync(FULL, ok);
                if (prefix) n = u + 1;
            }
            if (n == 0) {
                if ((++spins & 255) == 0 && (spins > (SPIN_LIMIT >> 2) || *reinterpret_cast<volatile int*>(a.err) != 0)) {
                    *a.err = 2;
                    return;
                }
                __nanosleep(spins > 64 ? 200 : 20);
                continue;
            }
            spins = 0;
            // ring 0 slot c is free once warp 0 has completed column c-NS+1
            wait_prog(&prog[1], x + n - 1 - NS + 2, seen, a.err, a.eager);
#pragma unroll
            for (int u = 0; u < Cfg::HD; ++u) {
                if (u < n) {
                    uint4* dst = ring + ((x + u) % NS) * Cfg::SLOT_V + l;
#pragma unroll
                    for (int j = 0; j < NQ * K; ++j)
                        dst[j * 32] = make_uint4(hb[u][j].x & ~TAGBITS, hb[u][j].y & ~TAGBITS, hb[u][j].z & ~TAGBITS,
                                                 hb[u][j].w & ~TAGBITS);
                }
            }
            __syncwarp();
            asm volatile("" ::: "memory");
            if (WSG_SWEEP_MBAR && bars != 0 && l < n) mbar_arrive(bars + (unsigned)((x + l) & (MBN - 1)) * 8u);   // row 0's barriers
            x += n;
            if (l == 0) prog[0] = x;
        }
        // one more column of zeros: the out-of-image predecessor of the last pixel's (x+1,y-1) path
        wait_prog(&prog[1], a.W1 - NS + 2, seen, a.err, a.eager);
#pragma unroll
        for (int j = 0; j < NQ * K; ++j) ring[(a.W1 % NS) * Cfg::SLOT_V + l + j * 32] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        asm volatile("" ::: "memory");
        if (l == 0) {
            prog[0] = a.W1 + 1;
            if (WSG_SWEEP_MBAR && bars != 0) mbar_arrive(bars + (unsigned)(a.W1 & (MBN - 1)) * 8u);
        }
        return;
    }
}

template <int K, int MODE, int NDIR, bool HASPAD, int CFG>
__global__ void __launch_bounds__(SW_THREADS, SweepCfg<K, CFG>::CTAS_PER_SM)
sweep_kernel(const uint4* __restrict__ C, uint4* __restrict__ S, SweepArgs a)
{
    using Cfg = SweepCfg<K, CFG>;
    extern __shared__ __align__(16) uint4 smem[];
    __shared__ volatile int prog[SW_R + 1];   // prog[0]: helper (row above the band); prog[r+1]: warp r
    __shared__ __align__(8) unsigned long long bars[(SW_R + 1) * MBN];   // column-publication barriers, same indexing as prog
    __shared__ int s_band;

    const int tid = threadIdx.x, warp = tid >> 5, l = tid & 31;
    const int nbands = (a.H + SW_R - 1) / SW_R;
    const unsigned bars_a = WSG_SWEEP_MBAR && NDIR >= 3 ? (unsigned)__cvta_generic_to_shared(bars) : 0u;
    const unsigned csize = cluster_nctarank(), crank = cluster_ctarank();
    if (csize == 2) {
        // ---- CTA pairs (thread-block cluster of two, one CTA per SM by its shared-memory size): the pair takes bands 2t and
        // 2t+1; rank 0's last row writes its states and its progress straight into rank 1's ring 0 (distributed shared
        // memory, release store on the counter), so every second band boundary costs an SM-to-SM hop instead of an L2
        // round trip.  Rank 1's last row uses the global hand-off buffer as before.
        const bool fast = *a.maxC + a.P2 <= 32767;
        while (true) {
            if (crank == 0 && tid == 0) {
                const int pair = atomicAdd(a.ticket, 1);
                s_band = pair;
                asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(dsmem_addr(&s_band, 1)), "r"(pair) : "memory");
            }
            if (tid <= SW_R) prog[tid] = 0;
            if (NDIR >= 3)
                for (int i = tid; i < Cfg::RINGS_V; i += SW_THREADS) smem[i] = make_uint4(0, 0, 0, 0);
            cluster_sync_all();                  // ticket visible in both CTAs; both rings clean before anyone writes into them
            const int band = 2 * s_band + (int)crank;
            if (2 * s_band >= nbands) break;
            if (band < nbands) {
                if (warp == SW_R) {
                    if (NDIR >= 3 && crank == 0) sweep_helper<K, CFG, NDIR == 4 ? 3 : 2>(a, smem, prog, band, l);
                } else if (fast) {
                    sweep_row<K, MODE, NDIR, HASPAD, CFG, true>(C, S, a, smem, prog, band, warp, l, crank == 0 && band + 1 < nbands, crank == 1);
                } else {
                    sweep_row<K, MODE, NDIR, HASPAD, CFG, false>(C, S, a, smem, prog, band, warp, l, crank == 0 && band + 1 < nbands, crank == 1);
                }
            }
            cluster_sync_all();                  // both bands done: nobody writes into a ring that is about to be reset
        }
        return;
    }
    // One WORKER per SM and launch: the first CTA of this launch to arrive on an SM stays and takes bands from a ticket
    // counter until none is left; every other CTA exits at once.  Bands therefore start in increasing order on
    // different SMs whatever the hardware's placement, nothing waits for a CTA that is not running, and the second
    // CTA slot of each SM stays free for the sweep (or cost kernel) of ANOTHER frame on another stream -- whose warps
    // fill the issue slots this latency-bound wavefront leaves empty.
    if (tid == 0) {
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        s_band = atomicAdd(a.ticket + 2 + min(sm, 250u), 1) == 0 ? 0 : -1;
        // A wavefront of H/7 bands cannot keep every SM busy (the skew between bands leaves ~40 % of the workers waiting at
        // any time), so the number of workers may be capped: the SMs given back run the sweep of another frame.
        if (s_band == 0 && atomicAdd(a.ticket + 1, 1) >= a.max_workers) s_band = -1;
    }
    __syncthreads();
    if (s_band < 0) return;
    const bool fast = *a.maxC + a.P2 <= 32767;   // the verified domain (known since the cost kernel ran): cheaper arithmetic
    bool first_band = true;
    while (true) {
        __syncthreads();                         // the previous band is finished by every warp
        if (tid == 0) {
            s_band = atomicAdd(a.ticket, 1);
            if (a.dbg && s_band < nbands) {
                unsigned sm; unsigned long long t;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                a.dbg[s_band] = (int)sm;
                a.dbg[4096 + 2 * s_band] = (int)(t & 0x7fffffff);      // start, ns (low bits)
            }
        }
        if (tid <= SW_R) prog[tid] = 0;
        if (bars_a) mbar_reset_all(bars, tid, first_band);
        first_band = false;
        if (NDIR >= 3)
            for (int i = tid; i < Cfg::RINGS_V; i += SW_THREADS) smem[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const int band = s_band;
        if (band >= nbands) break;
        if (warp == SW_R) {
            if (NDIR >= 3) sweep_helper<K, CFG, NDIR == 4 ? 3 : 2>(a, smem, prog, band, l, bars_a);
        } else if (fast) {
            sweep_row<K, MODE, NDIR, HASPAD, CFG, true>(C, S, a, smem, prog, band, warp, l, false, false, bars_a);
        } else {
            sweep_row<K, MODE, NDIR, HASPAD, CFG, false>(C, S, a, smem, prog, band, warp, l, false, false, bars_a);
        }
        if (a.dbg && warp == SW_R - 1 && l == 0) {           // the band's last row is done
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.dbg[4096 + 2 * band + 1] = (int)(t & 0x7fffffff);
        }
    }
}

template <int K, int MODE, int NDIR, bool HASPAD, int CFG>
static void launch_sweep_c(const int16_t* C, int16_t* S, const SweepArgs& a, cudaStream_t st)
{
    using Cfg = SweepCfg<K, CFG>;
    auto kern = sweep_kernel<K, MODE, NDIR, HASPAD, CFG>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    static int pairs = -1;
    // Opt-in experiment (WSG_SWEEP_PAIRS=1), not the default: 8.35 ms instead of 8.60 ms per frame for both sweeps, but the
    // counter store that follows the 32 lanes' remote data stores is only ordered behind them by the in-order delivery of
    // one warp's DSMEM stores (as with the CTA-local rings); with cluster-scope fences on both sides, which would make it
    // formally ordered, the last row of every even band takes so long that the frame needs 18.3 ms.
    if (pairs < 0) { const char* e = getenv("WSG_SWEEP_PAIRS"); pairs = e ? atoi(e) : 0; }
    const uint4* c = reinterpret_cast<const uint4*>(C);
    uint4* s = reinterpret_cast<uint4*>(S);
    if (pairs && NDIR >= 3 && Cfg::CTAS_PER_SM == 1) {
        // CTA pairs with a DSMEM hand-off between the two bands of a pair (see the kernel)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(a.num_sms & ~1));
        cfg.blockDim = dim3(SW_THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, kern, c, s, a) == cudaSuccess) return;
        cudaGetLastError();      // cluster launch not possible here: fall through to the plain launch
    }
    // enough CTAs for every SM to see one even when other kernels hold slots; all but one per SM exit immediately
    // (one CTA per SM by shared-memory size: launching max_workers CTAs puts them on that many different SMs)
    const int grid = Cfg::CTAS_PER_SM == 1 ? std::min(a.max_workers, a.num_sms) : Cfg::CTAS_PER_SM * a.num_sms;
    kern<<<grid, SW_THREADS, Cfg::SMEM, st>>>(c, s, a);
}

// ================================================================================================
// Two warps per image row (K == 1, four directions).  A sweep is latency-bound per warp (one SM runs 8 row-warps at
// ~0.55 IPC per scheduler; 16 rows per CTA showed +40 % SM throughput at twice the warps), but the number of rows in
// flight is fixed by the wavefront -- so the work of ONE row is split over two warps instead:
//   warp A   horizontal path (one pixel ahead) + the (x-1,y-1) path [+ the (x,y-1) path in the winner-take-all sweep]
//            -> partial sum of its L into a small exchange ring
//   warp B   the (x+1,y-1) path [+ the (x,y-1) path otherwise], S stream, total, S store or winner-take-all
// A rows trail each other by one column, B rows by two; A runs ahead of B by at most the exchange depth.  Each role has
// its own progress counters, its own part of every ring slot and of the band hand-off buffer, and its own helper warp.
// ================================================================================================
struct W2Cfg {
    static constexpr int NS = 5, NE = 4, PFA = 5, PFB = 4, PFS = 4, HD = 3;
    static constexpr int SLOT_V = 3 * 32, RING_V = NS * SLOT_V;
    static constexpr int RINGS_V = SW_R * RING_V;
    static constexpr int EX_V = SW_R * NE * 32;
    static constexpr int SCR_V = SW_R * 32;
    static constexpr int STA_V = SW_R * PFA * 32, STB_V = SW_R * PFB * 32, STS_V = SW_R * PFS * 32;
    static constexpr int SMEM = (RINGS_V + EX_V + SCR_V + STA_V + STB_V + STS_V) * 16;
    static constexpr int THREADS = (2 * SW_R + 2) * 32;
};

// helper warp of one role: forwards states q in [QLO, QLO+QN) of the previous band's last row into ring 0
template <int QLO, int QN>
__device__ __forceinline__ void w2_helper(const SweepArgs& a, uint4* smem, volatile int* prog, int band, int l)
{
    using Cfg = W2Cfg;
    constexpr int NS = Cfg::NS;
    if (band == 0) return;
    const uint4* src = a.bnd + (size_t)(band - 1) * ((size_t)a.W1 * Cfg::SLOT_V) + QLO * 32 + l;
    uint4* ring = smem + QLO * 32 + l;                     // ring 0
    uint4 hb[Cfg::HD][QN];
    int seen = 0, spins = 0;
    for (int x = 0; x < a.W1;) {
#pragma unroll
        for (int u = 0; u < Cfg::HD; ++u)
#pragma unroll
            for (int j = 0; j < QN; ++j) hb[u][j] = ld_volatile(src + (size_t)min(x + u, a.W1 - 1) * Cfg::SLOT_V + j * 32);
        int n = 0;
        bool prefix = true;
#pragma unroll
        for (int u = 0; u < Cfg::HD; ++u) {
            bool ok = x + u < a.W1;
#pragma unroll
            for (int j = 0; j < QN; ++j)
                ok = ok && ((hb[u][j].x & TAGBITS) == a.tag) && ((hb[u][j].y & TAGBITS) == a.tag) &&
                     ((hb[u][j].z & TAGBITS) == a.tag) && ((hb[u][j].w & TAGBITS) == a.tag);
            prefix = prefix && __all_sync(FULL, ok);
            if (prefix) n = u + 1;
        }
        if (n == 0) {
            if ((++spins & 255) == 0 && (spins > (SPIN_LIMIT >> 2) || *reinterpret_cast<volatile int*>(a.err) != 0)) {
                *a.err = 2;
                return;
            }
            __nanosleep(spins > 64 ? 200 : 20);
            continue;
        }
        spins = 0;
        wait_prog(&prog[1], x + n - 1 - NS + 2, seen, a.err, a.eager);
#pragma unroll
        for (int u = 0; u < Cfg::HD; ++u)
            if (u < n) {
#pragma unroll
                for (int j = 0; j < QN; ++j)
                    ring[((x + u) % NS) * Cfg::SLOT_V + j * 32] =
                        make_uint4(hb[u][j].x & ~TAGBITS, hb[u][j].y & ~TAGBITS, hb[u][j].z & ~TAGBITS, hb[u][j].w & ~TAGBITS);
            }
        __syncwarp();
        asm volatile("" ::: "memory");
        x += n;
        if (l == 0) prog[0] = x;
    }
    if (QLO + QN == 3) {      // the role that owns the (x+1,y-1) path: one more column of zeros (out-of-image predecessor)
        wait_prog(&prog[1], a.W1 - NS + 2, seen, a.err, a.eager);
#pragma unroll
        for (int j = 0; j < QN; ++j) ring[(a.W1 % NS) * Cfg::SLOT_V + j * 32] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        asm volatile("" ::: "memory");
        if (l == 0) prog[0] = a.W1 + 1;
    }
}

// role A of row r
template <int MODE, bool HASPAD, bool FAST>
__device__ __forceinline__ void w2_row_a(const uint4* __restrict__ C, const SweepArgs& a, uint4* smem, volatile int* progA,
                                         volatile int* progB, int band, int r, int l)
{
    using Cfg = W2Cfg;
    constexpr int NS = Cfg::NS, NE = Cfg::NE, PF = Cfg::PFA;
    constexpr int NA = MODE == 2 ? 2 : 1;            // paths from the row above handled here: q = 0 (x-1,y-1) [, q = 1 (x,y-1)]
    const unsigned one = a.one;
    const int yl = band * SW_R + r;
    if (yl >= a.H) return;
    const bool top = yl == 0;
    const int out_mode = yl == a.H - 1 ? 0 : (r < SW_R - 1 ? 1 : 2);
    const int yp = a.flip ? a.H - 1 - yl : yl;
    const long long dstep = a.flip ? -(long long)a.Dp8 : (long long)a.Dp8;
    const uint4* cpf = C + ((size_t)yp * a.W1 + (a.flip ? a.W1 - 1 : 0)) * a.Dp8 + l;
    const uint4* ring_in = smem + (size_t)r * Cfg::RING_V + l;
    uint4* ring_out = smem + (size_t)(r + 1) * Cfg::RING_V + l;
    uint4* bnd_out = a.bnd + (size_t)band * ((size_t)a.W1 * Cfg::SLOT_V) + l;
    uint4* ex = smem + Cfg::RINGS_V + (size_t)r * NE * 32 + l;
    uint4* stage = smem + Cfg::RINGS_V + Cfg::EX_V + Cfg::SCR_V + (size_t)r * PF * 32 + l;
    volatile int* prog_in = &progA[r];
    volatile int* prog_me = &progA[r + 1];
    volatile int* prog_next = &progA[r + 2 <= SW_R ? r + 2 : SW_R];
    volatile int* prog_b = &progB[r + 1];
    int seen_in = 0, seen_next = 0, seen_b = 0;
    unsigned padm[1] = {(l * 8 >= a.D) ? SAT2 : 0u};

    const unsigned st = (unsigned)__cvta_generic_to_shared(stage);
    for (int i = 0; i < PF; ++i) {
        if (i < a.W1) cp_async16(st + i * 32 * 16, cpf);
        cp_async_commit();
        cpf += dstep;
    }
    unsigned Nh[4] = {0, 0, 0, 0}, Cc[4], Lh[4];
    cp_async_wait<PF - 1>();
    { const uint4 c = stage[0]; Cc[0] = c.x; Cc[1] = c.y; Cc[2] = c.z; Cc[3] = c.w; }
    agg_step<32, 4, HASPAD, FAST>(Nh, Cc, Lh, l, a.P1p, a.P2mP1p, padm, one);

    int pslot = 0;
    for (int x = 0; x < a.W1; ++x) {
        const int nslot = pslot + 1 == PF ? 0 : pslot + 1;
        unsigned Cn[4], Lhn[4], v[NA][4], Nd[NA][4];
        cp_async_wait<PF - 2>();
        { const uint4 c = stage[nslot * 32]; Cn[0] = c.x; Cn[1] = c.y; Cn[2] = c.z; Cn[3] = c.w; }
        if (top) {
#pragma unroll
            for (int q = 0; q < NA; ++q) Nd[q][0] = Nd[q][1] = Nd[q][2] = Nd[q][3] = 0;
        } else {
            wait_prog(prog_in, x + 1, seen_in, a.err, a.eager);          // column x of the row above (its column -1 is a zero slot)
            const int sl[2] = {(x + NS - 1) % NS, x % NS};
#pragma unroll
            for (int q = 0; q < NA; ++q) {
                const uint4 t = ring_in[sl[q] * Cfg::SLOT_V + q * 32];
                Nd[q][0] = t.x; Nd[q][1] = t.y; Nd[q][2] = t.z; Nd[q][3] = t.w;
            }
        }
        agg_step<32, 4, HASPAD, FAST>(Nh, Cn, Lhn, l, a.P1p, a.P2mP1p, padm, one);
#pragma unroll
        for (int q = 0; q < NA; ++q) agg_step<32, 4, HASPAD, FAST>(Nd[q], Cc, v[q], l, a.P1p, a.P2mP1p, padm, one);
        if (out_mode == 1) {
            wait_prog(prog_next, x - NS + 2, seen_next, a.err, a.eager);
#pragma unroll
            for (int q = 0; q < NA; ++q)
                ring_out[(x % NS) * Cfg::SLOT_V + q * 32] = make_uint4(Nd[q][0], Nd[q][1], Nd[q][2], Nd[q][3]);
        } else if (out_mode == 2) {
#pragma unroll
            for (int q = 0; q < NA; ++q)
                st_volatile(bnd_out + (size_t)x * Cfg::SLOT_V + q * 32,
                            make_uint4(Nd[q][0] | a.tag, Nd[q][1] | a.tag, Nd[q][2] | a.tag, Nd[q][3] | a.tag));
        }
        // partial sum of this warp's L for warp B (B publishes column c before it reads the exchange slot of c: hence +2)
        unsigned pa[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            pa[j] = sat_add_split(Lh[j], v[0][j], one);
            if (NA == 2) pa[j] = sat_add_split(pa[j], v[1][j], one);
        }
        wait_prog(prog_b, x - NE + 2, seen_b, a.err, 0);       // back-pressure: warp A has slack, let it sleep
        ex[(x % NE) * 32] = make_uint4(pa[0], pa[1], pa[2], pa[3]);
        __syncwarp();
        asm volatile("" ::: "memory");
        if (l == 0) *prog_me = x + 1;
        if (x + PF < a.W1) cp_async16(st + pslot * 32 * 16, cpf);
        cp_async_commit();
        cpf += dstep;
        pslot = nslot;
#pragma unroll
        for (int j = 0; j < 4; ++j) { Cc[j] = Cn[j]; Lh[j] = Lhn[j]; }
    }
    cp_async_wait<0>();
}

// role B of row r
template <int MODE, bool HASPAD, bool FAST>
__device__ __forceinline__ void w2_row_b(const uint4* __restrict__ C, uint4* __restrict__ S, const SweepArgs& a, uint4* smem,
                                         volatile int* progA, volatile int* progB, int band, int r, int l)
{
    using Cfg = W2Cfg;
    constexpr int NS = Cfg::NS, NE = Cfg::NE, PF = Cfg::PFB, PFS = Cfg::PFS;
    constexpr int NB = MODE == 2 ? 1 : 2;            // paths from the row above handled here: [q = 1 (x,y-1),] q = 2 (x+1,y-1)
    constexpr int Q0 = 3 - NB;
    const unsigned one = a.one;
    const int yl = band * SW_R + r;
    if (yl >= a.H) return;
    const bool top = yl == 0;
    const int out_mode = yl == a.H - 1 ? 0 : (r < SW_R - 1 ? 1 : 2);
    const int yp = a.flip ? a.H - 1 - yl : yl;
    const long long dstep = a.flip ? -(long long)a.Dp8 : (long long)a.Dp8;
    const size_t first = ((size_t)yp * a.W1 + (a.flip ? a.W1 - 1 : 0)) * a.Dp8 + l;
    const uint4* cpf = C + first;
    const uint4* spf = S + first;
    uint4* scur = S + first;
    const uint4* ring_in = smem + (size_t)r * Cfg::RING_V + l;
    uint4* ring_out = smem + (size_t)(r + 1) * Cfg::RING_V + l;
    uint4* bnd_out = a.bnd + (size_t)band * ((size_t)a.W1 * Cfg::SLOT_V) + l;
    const uint4* ex = smem + Cfg::RINGS_V + (size_t)r * NE * 32 + l;
    int16_t* scratch = reinterpret_cast<int16_t*>(smem + Cfg::RINGS_V + Cfg::EX_V + (size_t)r * 32);
    uint4* stageC = smem + Cfg::RINGS_V + Cfg::EX_V + Cfg::SCR_V + Cfg::STA_V + (size_t)r * PF * 32 + l;
    uint4* stageS = smem + Cfg::RINGS_V + Cfg::EX_V + Cfg::SCR_V + Cfg::STA_V + Cfg::STB_V + (size_t)r * PFS * 32 + l;
    volatile int* prog_in = &progB[r];
    volatile int* prog_me = &progB[r + 1];
    volatile int* prog_next = &progB[r + 2 <= SW_R ? r + 2 : SW_R];
    volatile int* prog_a = &progA[r + 1];
    int seen_in = 0, seen_next = 0, seen_a = 0;
    unsigned padm[1] = {(l * 8 >= a.D) ? SAT2 : 0u};
    static_assert(PFS == PF, "one commit group per pixel carries C(x+PF) and S(x+PF)");

    const unsigned stC = (unsigned)__cvta_generic_to_shared(stageC), stS = (unsigned)__cvta_generic_to_shared(stageS);
    for (int i = 0; i < PF; ++i) {
        if (i < a.W1) {
            cp_async16(stC + i * 32 * 16, cpf);
            if (MODE != 0) cp_async16(stS + i * 32 * 16, spf);
        }
        cp_async_commit();
        cpf += dstep; spf += dstep;
    }
    unsigned vsp[4] = {0, 0, 0, 0};
    unsigned wkey = 0, wnb = 0, rkey = 0, rnb = 0;
    unsigned long long* keys_row = a.keys + (size_t)yp * a.W;
    int16_t* d1_row = a.d1 + (size_t)yp * a.W;

    int pslot = 0;
    for (int x = 0; x < a.W1; ++x) {
        unsigned Cc[4], vs[4], v[NB][4], Nd[NB][4];
        cp_async_wait<PF - 1>();
        { const uint4 c = stageC[pslot * 32]; Cc[0] = c.x; Cc[1] = c.y; Cc[2] = c.z; Cc[3] = c.w; }
        if (top) {
#pragma unroll
            for (int q = 0; q < NB; ++q) Nd[q][0] = Nd[q][1] = Nd[q][2] = Nd[q][3] = 0;
        } else {
            wait_prog(prog_in, x + 2, seen_in, a.err, a.eager);          // column x+1 of the row above (its column W1 is a zero slot)
            const int sl[3] = {0, x % NS, (x + 1) % NS};
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const uint4 t = ring_in[sl[Q0 + q] * Cfg::SLOT_V + (Q0 + q) * 32];
                Nd[q][0] = t.x; Nd[q][1] = t.y; Nd[q][2] = t.z; Nd[q][3] = t.w;
            }
        }
        if (MODE == 2) wta_eval<1, HASPAD>(vsp, l, a, scratch, wkey, wnb);      // previous pixel, an independent chain
#pragma unroll
        for (int q = 0; q < NB; ++q) agg_step<32, 4, HASPAD, FAST>(Nd[q], Cc, v[q], l, a.P1p, a.P2mP1p, padm, one);
        if (out_mode == 1) {
            wait_prog(prog_next, x - NS + 2, seen_next, a.err, a.eager);
#pragma unroll
            for (int q = 0; q < NB; ++q)
                ring_out[(x % NS) * Cfg::SLOT_V + (Q0 + q) * 32] = make_uint4(Nd[q][0], Nd[q][1], Nd[q][2], Nd[q][3]);
        } else if (out_mode == 2) {
#pragma unroll
            for (int q = 0; q < NB; ++q)
                st_volatile(bnd_out + (size_t)x * Cfg::SLOT_V + (Q0 + q) * 32,
                            make_uint4(Nd[q][0] | a.tag, Nd[q][1] | a.tag, Nd[q][2] | a.tag, Nd[q][3] | a.tag));
        }
        __syncwarp();
        asm volatile("" ::: "memory");
        if (l == 0) *prog_me = x + 1;               // rows below may go on; the exchange slot of x is read only now
        // ---- total: S_in + warp A's partial sum + this warp's L
        wait_prog(prog_a, x + 1, seen_a, a.err, a.eager);
        { const uint4 pa = ex[(x % NE) * 32]; vs[0] = pa.x; vs[1] = pa.y; vs[2] = pa.z; vs[3] = pa.w; }
        if (MODE != 0) {
            const uint4 sv = stageS[pslot * 32];
            vs[0] = sat_add_split(vs[0], sv.x, one); vs[1] = sat_add_split(vs[1], sv.y, one);
            vs[2] = sat_add_split(vs[2], sv.z, one); vs[3] = sat_add_split(vs[3], sv.w, one);
        }
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) vs[j] = sat_add_split(vs[j], v[q][j], one);
        if (MODE == 2) {
            if (l == ((x - 1) & 31)) { rkey = wkey; rnb = wnb; }
            if (x > 0 && (x & 31) == 0) wta_flush(rkey, rnb, x - 32 + l, true, a, keys_row, d1_row);
            __syncwarp();
            reinterpret_cast<uint4*>(scratch)[l] = make_uint4(vs[0], vs[1], vs[2], vs[3]);
#pragma unroll
            for (int j = 0; j < 4; ++j) vsp[j] = vs[j];
            __syncwarp();
        } else {
            stg_stream(scur, make_uint4(vs[0], vs[1], vs[2], vs[3]));
        }
        scur += dstep;
        if (x + PF < a.W1) {
            cp_async16(stC + pslot * 32 * 16, cpf);
            if (MODE != 0) cp_async16(stS + pslot * 32 * 16, spf);
        }
        cp_async_commit();
        cpf += dstep; spf += dstep;
        pslot = pslot + 1 == PF ? 0 : pslot + 1;
    }
    cp_async_wait<0>();
    if (MODE == 2) {
        wta_eval<1, HASPAD>(vsp, l, a, scratch, wkey, wnb);
        if (l == ((a.W1 - 1) & 31)) { rkey = wkey; rnb = wnb; }
        const int xb = (a.W1 - 1) & ~31;
        wta_flush(rkey, rnb, xb + l, xb + l < a.W1, a, keys_row, d1_row);
    }
    if (out_mode == 1) {      // the extra zero column for the (x+1,y-1) path of the row below
        wait_prog(prog_next, a.W1 - NS + 2, seen_next, a.err, a.eager);
#pragma unroll
        for (int q = 0; q < NB; ++q) ring_out[(a.W1 % NS) * Cfg::SLOT_V + (Q0 + q) * 32] = make_uint4(0, 0, 0, 0);
        __syncwarp();
        asm volatile("" ::: "memory");
        if (l == 0) *prog_me = a.W1 + 1;
    }
}

template <int MODE, bool HASPAD>
__global__ void __launch_bounds__(W2Cfg::THREADS, 1)
sweep2w_kernel(const uint4* __restrict__ C, uint4* __restrict__ S, SweepArgs a)
{
    using Cfg = W2Cfg;
    extern __shared__ __align__(16) uint4 smem[];
    __shared__ volatile int progA[SW_R + 1], progB[SW_R + 1];
    __shared__ int s_band;
    const int tid = threadIdx.x, warp = tid >> 5, l = tid & 31;
    if (tid == 0) {      // one worker per SM and launch (see sweep_kernel)
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        s_band = atomicAdd(a.ticket + 2 + min(sm, 250u), 1) == 0 ? 0 : -1;
    }
    __syncthreads();
    if (s_band < 0) return;
    const bool fast = *a.maxC + a.P2 <= 32767;
    const int nbands = (a.H + SW_R - 1) / SW_R;
    constexpr int NA = MODE == 2 ? 2 : 1;
    while (true) {
        __syncthreads();
        if (tid == 0) s_band = atomicAdd(a.ticket, 1);
        if (tid <= SW_R) { progA[tid] = 0; progB[tid] = 0; }
        for (int i = tid; i < Cfg::RINGS_V; i += Cfg::THREADS) smem[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        const int band = s_band;
        if (band >= nbands) break;
        if (warp == 2 * SW_R) {
            w2_helper<0, NA>(a, smem, progA, band, l);
        } else if (warp == 2 * SW_R + 1) {
            w2_helper<NA, 3 - NA>(a, smem, progB, band, l);
        } else if ((warp & 1) == 0) {
            if (fast) w2_row_a<MODE, HASPAD, true>(C, a, smem, progA, progB, band, warp >> 1, l);
            else      w2_row_a<MODE, HASPAD, false>(C, a, smem, progA, progB, band, warp >> 1, l);
        } else {
            if (fast) w2_row_b<MODE, HASPAD, true>(C, S, a, smem, progA, progB, band, warp >> 1, l);
            else      w2_row_b<MODE, HASPAD, false>(C, S, a, smem, progA, progB, band, warp >> 1, l);
        }
    }
}

template <int MODE, bool HASPAD>
static void launch_sweep2w_t(const int16_t* C, int16_t* S, const SweepArgs& a, cudaStream_t st)
{
    auto kern = sweep2w_kernel<MODE, HASPAD>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, W2Cfg::SMEM);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    kern<<<a.num_sms, W2Cfg::THREADS, W2Cfg::SMEM, st>>>(reinterpret_cast<const uint4*>(C), reinterpret_cast<uint4*>(S), a);
}

static int g_sweep_cfg = -1;
void set_sweep_cfg(int cfg) { g_sweep_cfg = cfg; }

template <int K, int MODE, int NDIR, bool HASPAD>
static void launch_sweep_t(const int16_t* C, int16_t* S, const SweepArgs& a, cudaStream_t st)
{
    if (g_sweep_cfg < 0) {
        const char* e = getenv("WSG_SWEEP_CFG");
        g_sweep_cfg = e ? atoi(e) : 0;
    }
    if (K == 1 && NDIR == 4) {
        switch (g_sweep_cfg) {
        case 1: launch_sweep_c<K, MODE, NDIR, HASPAD, (K == 1 && NDIR == 4) ? 1 : 0>(C, S, a, st); return;
        case 2: launch_sweep_c<K, MODE, NDIR, HASPAD, (K == 1 && NDIR == 4) ? 2 : 0>(C, S, a, st); return;
        default: break;
        }
    }
    launch_sweep_c<K, MODE, NDIR, HASPAD, 0>(C, S, a, st);
}

size_t sweep_boundary_bytes(const SgbmPlan& p)
{
    const int nbands = (p.H + SW_R - 1) / SW_R;
    return (size_t)std::max(nbands - 1, 1) * p.W1 * 3 * p.K * 32 * 16;
}

bool sweep_supported(const SgbmPlan& p) { return p.NL == 32 && (p.K == 1 || p.K == 2); }

// mode: 0 first (write S), 1 accumulate (read+write S), 2 accumulate + winner-take-all.  ndir: 4 or 1.
void launch_sweep(const int16_t* C, int16_t* S, int flip, int mode, int ndir, const SgbmPlan& p, const SweepScratch& sc,
                  cudaStream_t st)
{
    SweepArgs a;
    a.H = p.H; a.W1 = p.W1; a.W = p.W; a.D = p.D; a.Dp8 = p.Dp / 8; a.flip = flip;
    a.P1p = ((unsigned)p.P1 & 0xFFFFu) * 0x10001u;
    a.P2mP1p = ((unsigned)(p.P2 - p.P1) & 0xFFFFu) * 0x10001u;
    a.P2 = p.P2; a.maxC = sc.maxC; a.one = 1u;
    a.tag = ((sc.epoch & 1) ? 0x8000u : 0u) | ((sc.epoch & 2) ? 0x80000000u : 0u);
    a.bnd = reinterpret_cast<uint4*>(sc.boundary);
    a.ticket = sc.ticket; a.err = sc.err; a.dbg = sc.dbg;
    a.num_sms = sc.num_sms;
    static int workers = -1;
    if (workers < 0) { const char* e = getenv("WSG_SWEEP_WORKERS"); workers = e ? atoi(e) : 0; }
    a.max_workers = workers > 0 ? workers : (sc.max_workers > 0 ? sc.max_workers : sc.num_sms);
    static int eager = -1;
    if (eager < 0) { const char* e = getenv("WSG_SWEEP_EAGER"); eager = e ? atoi(e) : 64; }
    a.eager = eager;
    a.keys = sc.keys; a.d1 = sc.d1;
    a.minD = p.minD; a.minX1 = p.minX1; a.uniq = p.uniq; a.INVALID = p.INVALID;
    a.umagic = p.uniq < 99 ? (unsigned)((0x100000000ull + (100 - p.uniq) - 1) / (unsigned)(100 - p.uniq)) : 0u;
    const bool pad = p.Dp != p.D;
    // two warps per row (WSG_AGG_SWEEPS2W_WTA): measured equal to one warp per row at 2448x2048x256 (8.6 ms per frame for
    // both sweeps): the scheduler issue rate rises from 40 % to 58 %, but the split costs 25 % more instructions per
    // pixel and the extra progress-counter polling takes the rest.  Kept selectable, not the default.
    if (sc.two_warps && p.K == 1 && ndir == 4) {
        if (mode == 0) { if (pad) launch_sweep2w_t<0, true>(C, S, a, st); else launch_sweep2w_t<0, false>(C, S, a, st); }
        else if (mode == 1) { if (pad) launch_sweep2w_t<1, true>(C, S, a, st); else launch_sweep2w_t<1, false>(C, S, a, st); }
        else { if (pad) launch_sweep2w_t<2, true>(C, S, a, st); else launch_sweep2w_t<2, false>(C, S, a, st); }
        return;
    }
#define WSG_SW_CASE(k, m, n)                                                         \
    if (p.K == k && mode == m && ndir == n) {                                        \
        if (pad) launch_sweep_t<k, m, n, true>(C, S, a, st);                        \
        else     launch_sweep_t<k, m, n, false>(C, S, a, st);                       \
        return;                                                                      \
    }
    WSG_SW_CASE(1, 0, 4) WSG_SW_CASE(1, 1, 4) WSG_SW_CASE(1, 2, 4) WSG_SW_CASE(1, 1, 1) WSG_SW_CASE(1, 2, 1)
    WSG_SW_CASE(2, 0, 4) WSG_SW_CASE(2, 1, 4) WSG_SW_CASE(2, 2, 4) WSG_SW_CASE(2, 1, 1) WSG_SW_CASE(2, 2, 1)
    WSG_SW_CASE(1, 0, 3) WSG_SW_CASE(1, 1, 3) WSG_SW_CASE(1, 2, 3) WSG_SW_CASE(2, 0, 3) WSG_SW_CASE(2, 1, 3) WSG_SW_CASE(2, 2, 3)
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

void launch_wta_reset(const SweepScratch& sc, const SgbmPlan& p, cudaStream_t st)
{
    const size_t n = (size_t)p.H * p.W;
    wta_reset_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sc.keys, sc.d1, n, p.INVALID);
}

void launch_lrcheck(const SweepScratch& sc, int16_t* raw, const SgbmPlan& p, cudaStream_t st)
{
    dim3 b(256), g((p.W + 255) / 256, p.H);
    lrcheck_kernel<<<g, b, 0, st>>>(sc.keys, sc.d1, raw, p);
}

}  // namespace wsg
