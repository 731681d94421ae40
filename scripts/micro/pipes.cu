// pipes.cu -- instruction-throughput microbenchmark (development aid): warp-instructions per cycle per SM for the
// integer / half-precision operations the 8-bit filter is built from, alone and in pairs (do two op classes share a pipe?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

#define ITERS 2048
#define NACC 8

template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    if (OP == 0) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 1) r = __viaddmin_u16x2(a, b, c);
    else if (OP == 2) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 3) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 4) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 5) r = __vabsdiffu4(a, b) ^ c;      // VABSDIFF4.U8 + a LOP3
    else if (OP == 6) asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 7) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 8) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b + c));
    else if (OP == 9) r = __vimin3_u16x2(a, b, c);
    else if (OP == 10) asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(c));
    else if (OP == 11) asm volatile("{ .reg .pred p; setp.eq.f16x2 p|_, %1, %2; selp.u32 %0, %3, %1, p; }" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 12) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 13) r = __popc(a) + c;
    else if (OP == 14) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 15) asm volatile("min.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(c));
    else if (OP == 16) asm volatile("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    else if (OP == 17) asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(c));
    else r = a;
    return r;
}

template <int A, int B>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t x[NACC], y[NACC];
    for (int i = 0; i < NACC; i++) { x[i] = seed * (i + 1) + threadIdx.x; y[i] = seed ^ (i * 77 + threadIdx.x); }
    const uint32_t b = seed | 1, c = seed + 3;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            x[i] = op<A>(x[i], b, c);
            if (B >= 0) y[i] = op<B>(y[i], b, c);
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < NACC; i++) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int A, int B> void run(const char *name, uint32_t *out, long long *cyc) {
    const int grid = 148 * 4;      // 4 CTAs of 8 warps per SM: 32 warps, 8 per scheduler
    k<A, B><<<grid, 256>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    long long h[148 * 4];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; i++) avg += (double)h[i];
    avg /= grid;
    const double instr_per_warp = (double)ITERS * NACC * (B >= 0 ? 2 : 1);
    // per SM: 32 warps resident, all run concurrently for ~avg cycles
    printf("%-34s %7.3f warp-instr / cycle / SM   (%.0f cycles)\n", name, 32.0 * instr_per_warp / avg, avg);
}

int main() {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&cyc, 148 * 4 * 8);
    run<0, -1>("LOP3", out, cyc);
    run<1, -1>("VIADDMNMX.U16x2", out, cyc);
    run<9, -1>("VIMNMX3.U16x2", out, cyc);
    run<2, -1>("IMAD", out, cyc);
    run<3, -1>("PRMT", out, cyc);
    run<7, -1>("SHF", out, cyc);
    run<8, -1>("IADD3", out, cyc);
    run<15, -1>("VIMNMX.U32", out, cyc);
    run<4, -1>("VABSDIFF4.ACC (sad)", out, cyc);
    run<5, -1>("VABSDIFF4 + LOP3", out, cyc);
    run<6, -1>("HFMA2", out, cyc);
    run<16, -1>("HFMA2.BF16", out, cyc);
    run<10, -1>("HADD2", out, cyc);
    run<17, -1>("HMNMX2", out, cyc);
    run<11, -1>("HSETP2 + SEL", out, cyc);
    run<12, -1>("IDP4A", out, cyc);
    run<13, -1>("POPC + IADD", out, cyc);
    run<14, -1>("FFMA", out, cyc);
    run<1, 0>("VIADDMNMX + LOP3", out, cyc);
    run<1, 2>("VIADDMNMX + IMAD", out, cyc);
    run<1, 6>("VIADDMNMX + HFMA2", out, cyc);
    run<2, 6>("IMAD + HFMA2", out, cyc);
    run<0, 6>("LOP3 + HFMA2", out, cyc);
    run<1, 17>("VIADDMNMX + HMNMX2", out, cyc);
    run<1, 4>("VIADDMNMX + VABSDIFF4.ACC", out, cyc);
    run<14, 6>("FFMA + HFMA2", out, cyc);
    run<14, 2>("FFMA + IMAD", out, cyc);
    return 0;
}
