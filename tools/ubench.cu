// Micro-benchmarks that size the aggregation design: packed-int16 ALU throughput, SHFL / REDUX
// throughput and latency, and CTA-to-CTA flag latency through L2.   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
template <int OP, int ILP>
__global__ void k_thr(unsigned* out, int iters, unsigned a0, unsigned b0)
{
    unsigned r[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) r[j] = a0 + threadIdx.x * 7 + j;
    unsigned b = b0, c = b0 * 3 + 1;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (OP == 0) r[j] = __vimin3_s16x2(r[j], b, c);
            if (OP == 1) r[j] = __viaddmin_s16x2(r[j], b, c);
            if (OP == 2) r[j] = __byte_perm(r[j], b, 0x5432);
            if (OP == 3) r[j] = __vadd2(r[j], b);
            if (OP == 4) r[j] = __shfl_up_sync(FULL, r[j], 1);
            if (OP == 5) r[j] = __reduce_min_sync(FULL, r[j]) + j;
            if (OP == 6) r[j] = r[j] * b + c;                       // IMAD
            if (OP == 7) { r[j] = __vimin3_s16x2(r[j], b, c); r[j] = r[j] * b + c; }  // alu + fma pipe mix
            if (OP == 8) r[j] = __viaddmin_u16x2(r[j], b, c);
            if (OP == 9) r[j] = __vmins2(r[j], b);
            if (OP == 10) r[j] = min(r[j], b) ;
            if (OP == 11) r[j] = __shfl_xor_sync(FULL, r[j], 1);
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += r[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP, int ILP>
void run(const char* name, int opsPerIter)
{
    unsigned* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps : {4, 16, 32}) {
        int iters = 4096;
        k_thr<OP, ILP><<<148, warps * 32>>>(out, 16, 1, 2);
        cudaEventRecord(e0);
        k_thr<OP, ILP><<<148, warps * 32>>>(out, iters, 1, 2);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double winstr = (double)148 * warps * iters * ILP * opsPerIter;
        printf("%-28s warps/SM=%2d ILP=%d : %.1f Gwarp-instr/s  (%.3f per clk per SM @1.9GHz)\n", name, warps, ILP, winstr / ms / 1e6,
               winstr / ms / 1e6 / 148 / 1.9);
    }
    cudaFree(out);
}
// dependent-chain latency (1 warp, ILP=1)
template <int OP>
__global__ void k_lat(unsigned* out, int iters, unsigned b, long long* cyc)
{
    unsigned r = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) r = __vimin3_s16x2(r, b, b + 1);
        if (OP == 4) r = __shfl_up_sync(FULL, r, 1);
        if (OP == 5) r = __reduce_min_sync(FULL, r) + 1;
        if (OP == 11) r = __shfl_xor_sync(FULL, r, 1);
    }
    long long t1 = clock64();
    out[threadIdx.x] = r;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
// flag ping-pong between two CTAs on different SMs
__global__ void k_pingpong(volatile int* flag, int iters, long long* cyc)
{
    if (threadIdx.x != 0) return;
    int me = blockIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (me == 0) {
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(2 * i + 1) : "memory");
            int v; do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag + 32) : "memory"); } while (v < 2 * i + 2);
        } else {
            int v; do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); } while (v < 2 * i + 1);
            asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag + 32), "r"(2 * i + 2) : "memory");
        }
    }
    long long t1 = clock64();
    if (me == 0) *cyc = t1 - t0;
}
int main()
{
    run<0, 8>("VIMNMX3.S16x2", 1); run<1, 8>("VIADDMNMX.S16x2", 1); run<8, 8>("VIADDMNMX.U16x2", 1);
    run<2, 8>("PRMT", 1); run<3, 8>("VIADD.16x2", 1); run<9, 8>("VIMNMX.S16x2", 1); run<10, 8>("IMNMX.U32", 1);
    run<6, 8>("IMAD", 1); run<7, 8>("VIMNMX3+IMAD (2 ops)", 2);
    run<4, 8>("SHFL.UP", 1); run<11, 8>("SHFL.BFLY", 1); run<5, 8>("REDUX.MIN (+IADD)", 1);
    unsigned* out; long long* cyc; cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 8);
    k_lat<0><<<1, 32>>>(out, 10000, 3, cyc); cudaDeviceSynchronize(); printf("latency VIMNMX3: %.1f clk\n", *cyc / 10000.0);
    k_lat<4><<<1, 32>>>(out, 10000, 3, cyc); cudaDeviceSynchronize(); printf("latency SHFL.UP: %.1f clk\n", *cyc / 10000.0);
    k_lat<11><<<1, 32>>>(out, 10000, 3, cyc); cudaDeviceSynchronize(); printf("latency SHFL.BFLY: %.1f clk\n", *cyc / 10000.0);
    k_lat<5><<<1, 32>>>(out, 10000, 3, cyc); cudaDeviceSynchronize(); printf("latency REDUX.MIN+IADD: %.1f clk\n", *cyc / 10000.0);
    int* flag; cudaMalloc(&flag, 4096); cudaMemset(flag, 0, 4096);
    k_pingpong<<<2, 32>>>(flag, 2000, cyc); cudaDeviceSynchronize();
    printf("flag round trip between 2 CTAs: %.1f clk (one way ~ half)\n", *cyc / 2000.0);
    return 0;
}
