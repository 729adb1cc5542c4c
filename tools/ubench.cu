// tools/ubench.cu -- development probe: issue throughput (warp-instructions / clk / SM) of the
// instruction classes the fused kernel is made of, measured with clock64 on one resident block per SM.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint32_t seed, uint32_t* sink, long long* cyc)
{
    uint32_t a[ILP]; float f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = seed + threadIdx.x * 977u + i * 131u; f[i] = 1.0f + (float)(a[i] & 1023) * 1e-3f; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) { uint64_t p = (uint64_t)a[i] * 0xD2511F53u; a[i] = (uint32_t)p ^ (uint32_t)(p >> 32); }          // IMAD.WIDE + LOP3
            if (OP == 1) { a[i] = a[i] * 0xD2511F53u + seed; }                                                              // IMAD
            if (OP == 2) { a[i] = (a[i] ^ seed) & (a[(i + 1) % ILP] | 0x55u); }                                             // LOP3
            if (OP == 3) { f[i] = fmaf(f[i], 1.0000001f, 1e-9f); }                                                          // FFMA
            if (OP == 4) { asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }                                       // MUFU
            if (OP == 5) { a[i] = __byte_perm(a[i], seed, 0x7610 ^ (i & 1)); }                                              // PRMT
            if (OP == 6) { f[i] = fmaxf(f[i], f[(i + 1) % ILP] * 0.5f); }                                                    // FMNMX + FMUL
            if (OP == 7) { uint64_t p = (uint64_t)a[i] * 0xD2511F53u; a[i] = (uint32_t)(p >> 32) + (uint32_t)p; }            // IMAD.WIDE + IADD
            if (OP == 8) { asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }                                       // FMUL + MUFU.SIN
            if (OP == 9) { f[i] = fmaf(f[i], 1.0000001f, 1e-9f); a[i] = a[i] * 0xD2511F53u + seed; }                         // FFMA + IMAD mix
            if (OP == 10) { f[i] = fmaf(f[i], 1.0000001f, 1e-9f); a[i] = (a[i] ^ seed) & (a[(i + 1) % ILP] | 0x55u); }       // FFMA + LOP3 mix
            if (OP == 11) { f[i] = f[i] + 1e-9f; }                                                                          // FADD
            if (OP == 12) { f[i] = __uint2float_rn(__float_as_uint(f[i]) ^ a[i]); }                                         // I2F.U32 + LOP3
            if (OP == 13) { a[i] = __umulhi(a[i], 0xD2511F53u) + seed; }                                                    // IMAD.HI
            if (OP == 14) { uint64_t p = (uint64_t)a[i] * 0xD2511F53u + (uint64_t)a[(i + 1) % ILP]; a[i] = (uint32_t)(p >> 32); }  // IMAD.WIDE with add (chained on hi)
            if (OP == 15) { asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f[i])); }                                     // MUFU.SQRT
            if (OP == 17) { uint64_t p = (uint64_t)a[i] * 0xD2511F53u; a[i] = (uint32_t)p ^ (uint32_t)(p >> 32);            // 1 IMAD.WIDE + 1 LOP3 + 4 FFMA
                            f[i] = fmaf(f[i], 1.0000001f, 1e-9f); f[i] = fmaf(f[i], 1.0000002f, 2e-9f);
                            f[i] = fmaf(f[i], 1.0000003f, 3e-9f); f[i] = fmaf(f[i], 1.0000004f, 4e-9f); }
            if (OP == 18) { uint64_t p = (uint64_t)a[i] * 0xD2511F53u; a[i] = (uint32_t)p ^ (uint32_t)(p >> 32);            // 1 IMAD.WIDE + 1 LOP3 + 8 FFMA
                            f[i] = fmaf(f[i], 1.0000001f, 1e-9f); f[i] = fmaf(f[i], 1.0000002f, 2e-9f);
                            f[i] = fmaf(f[i], 1.0000003f, 3e-9f); f[i] = fmaf(f[i], 1.0000004f, 4e-9f);
                            f[i] = fmaf(f[i], 1.0000005f, 1e-9f); f[i] = fmaf(f[i], 1.0000006f, 2e-9f);
                            f[i] = fmaf(f[i], 1.0000007f, 3e-9f); f[i] = fmaf(f[i], 1.0000008f, 4e-9f); }
            if (OP == 19) { f[i] = fmaf(f[i], 1.0000001f, 1e-9f); f[i] = fmaf(f[i], 1.0000002f, 2e-9f);                      // 4 FFMA (baseline for 17)
                            f[i] = fmaf(f[i], 1.0000003f, 3e-9f); f[i] = fmaf(f[i], 1.0000004f, 4e-9f); }
            if (OP >= 20 && OP <= 23) {                                                                                     // 16-slot groups: FFMA only / with 1 MUFU / with 2 MUFU / 2 MUFU + 4 FADD|x| + 2 FMNMX
                float x = f[i];
                const int nf = OP == 20 ? 16 : OP == 21 ? 15 : OP == 22 ? 14 : 8;
#pragma unroll
                for (int r = 0; r < nf; r++) x = fmaf(x, 1.0000001f + r * 1e-7f, 1e-9f);
                if (OP >= 21) { float g = __uint_as_float(a[i]); float sn, cs = 0.f;
                                asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(g));
                                if (OP >= 22) asm volatile("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(g));
                                a[i] = __float_as_uint(sn) ^ __float_as_uint(cs); }
                if (OP == 23) { float y = fabsf(x) - 0.5f; float z = fabsf(y) - 0.25f; float w = fabsf(z) - 0.125f; float v = fabsf(w) - 0.0625f;
                                x = fmaxf(fmaxf(y, z), fmaxf(w, v)); }
                f[i] = x;
            }
            if (OP == 24) { f[i] = fmaf(f[i], f[(i + 1) % ILP], f[(i + 3) % ILP]); }                                         // FFMA, three register sources
            if (OP == 25) { f[i] = fmaf(f[i], f[(i + 1) % ILP], 1e-9f); }                                                   // FFMA, two register sources
            if (OP == 26) { f[i] = fmaf(f[(i + 2) % ILP], f[(i + 1) % ILP], f[(i + 3) % ILP]) + f[i] * 1e-9f; }              // FFMA 3 regs (dst != src) + FFMA
            if (OP == 30) { unsigned long long x = ((unsigned long long)__float_as_uint(f[(i + 1) % ILP]) << 32) | __float_as_uint(f[i]);      // FFMA2 (packed 2 x FP32)
                            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(0x3f8000013f800001ull), "l"(0x3089705f3089705full));
                            f[i] = __uint_as_float((unsigned)x); }
            if (OP == 31) { f[i] = fmaxf(fmaxf(f[i], f[(i + 1) % ILP]), f[(i + 3) % ILP]) * 1.0000001f; }                        // FMNMX3 + FMUL
            if (OP == 32) { a[i] = __funnelshift_l(__float_as_uint(f[i]), a[i], 1); }                                           // SHF
            if (OP == 33) { a[i] += (fabsf(f[i]) > f[(i + 1) % ILP]) ? 1u : 0u; }                                               // FSETP + predicated add
            if (OP == 16) { a[i] = __float_as_uint(f[i] = fmaf(f[i], 1.0000001f, 1e-9f)) >> 31; }                           // FFMA + SHF
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += a[i] + __float_as_uint(f[i]);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static int g_threads = 1024;

// packed FP32 (fma.rn.f32x2 -> FFMA2 on sm_100a): 8 independent 64-bit accumulators per thread
template <int OP>
__global__ void __launch_bounds__(1024, 1) k2(uint32_t seed, uint32_t* sink, long long* cyc)
{
    unsigned long long x[8]; float f[16]; uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = 0x3f8000003f800000ull + seed + threadIdx.x * 7u + i; a[i] = seed + i * 77u + threadIdx.x; }
#pragma unroll
    for (int i = 0; i < 16; i++) f[i] = 1.0f + (float)((seed + i * 13u + threadIdx.x) & 1023) * 1e-3f;
    const unsigned long long m = 0x3f8000013f800001ull, c = 0x3089705f3089705full;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0 || OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(m), "l"(c));
            if (OP == 1 || OP == 3) { f[2 * i] = fmaf(f[2 * i], 1.0000001f, 1e-9f); f[2 * i + 1] = fmaf(f[2 * i + 1], 1.0000001f, 1e-9f); }
            if (OP == 2 || OP == 3) a[i] = (a[i] ^ seed) & (a[(i + 1) % 8] | 0x55u);
            if (OP == 4) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(c));
            if (OP == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(x[i]) : "l"(m));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += (uint32_t)x[i] + (uint32_t)(x[i] >> 32) + a[i] + __float_as_uint(f[2 * i]) + __float_as_uint(f[2 * i + 1]);
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP> void run2(const char* name, int instr_per_iter)
{
    int sms = 148;
    uint32_t* sink; long long* cyc;
    cudaMalloc(&sink, sms * 1024 * 4); cudaMalloc(&cyc, sms * 8);
    k2<OP><<<sms, 1024>>>(12345u, sink, cyc);
    k2<OP><<<sms, 1024>>>(12345u, sink, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    const double warps_per_smsp = 8.0;
    printf("%-28s %8.1f cyc  %6.3f clk per SMSP per iteration (%d warp-instr): %5.3f warp-instr/clk/SMSP\n", name, avg,
           avg / ITERS / warps_per_smsp, instr_per_iter, instr_per_iter * ITERS * warps_per_smsp / avg);
    cudaFree(sink); cudaFree(cyc);
}

template <int OP> void run(const char* name, int instr_per_op)
{
    int sms = 148;
    uint32_t* sink; long long* cyc;
    cudaMalloc(&sink, sms * 1024 * 4); cudaMalloc(&cyc, sms * 8);
    k<OP><<<sms, g_threads>>>(12345u, sink, cyc);
    k<OP><<<sms, g_threads>>>(12345u, sink, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += h[i]; avg /= sms;
    double warp_ops = (g_threads / 32.0) * ITERS * ILP;            // per SM: g_threads / 32 warps
    printf("%-28s %8.1f cyc  %6.3f ops/clk/SM (=%5.2f per SMSP)  ~%.3f warp-instr/clk/SM\n", name, avg, warp_ops / avg, warp_ops / avg / 4,
           warp_ops * instr_per_op / avg);
    cudaFree(sink); cudaFree(cyc);
}

int main(int argc, char** argv)
{
    if (argc > 3) {                      // packed FP32 and the ALU-pipe instructions of the sweep loop
        run2<0>("FFMA2 x8 (16 FMA)", 8); run2<1>("FFMA x16", 16); run2<2>("FFMA2 x8 + LOP3 x8", 16); run2<3>("FFMA x16 + LOP3 x8", 24);
        run2<4>("FADD2 x8", 8); run2<5>("FMUL2 x8", 8);
        run<31>("FMNMX3+FMUL", 2); run<32>("SHF", 1); run<33>("FSETP+IADD", 2);
        return 0;
    }
    if (argc > 2) {                      // register-operand probe
        run<3>("FFMA r,imm,imm", 1); run<25>("FFMA r,r,imm", 1); run<24>("FFMA r,r,r", 1); run<26>("FFMA r,r,r + FFMA r,imm,r", 2);
        return 0;
    }
    if (argc > 1) {                      // occupancy probe: the screening mixes at fewer resident warps per scheduler
        for (int t : {1024, 768, 512, 384, 256, 128}) {
            g_threads = t; printf("-- %d warps per scheduler\n", t / 128);
            run<20>("16 FFMA", 16); run<22>("14 FFMA + SIN + COS", 18); run<23>("8 FFMA+4 FADD+3 FMNMX+SIN+COS", 19);
        }
        return 0;
    }
    run<3>("FFMA", 1); run<11>("FADD", 1); run<1>("IMAD", 1); run<0>("IMAD.WIDE+LOP3", 2); run<7>("IMAD.WIDE+IADD3", 2); run<2>("LOP3(x2)", 2);
    run<12>("I2F.U32+LOP3", 2); run<13>("IMAD.HI", 1); run<14>("IMAD.WIDE(+c)", 1); run<15>("MUFU.SQRT", 1);
    run<19>("4 FFMA", 4); run<17>("IMAD.WIDE+LOP3+4 FFMA", 6); run<18>("IMAD.WIDE+LOP3+8 FFMA", 10);
    run<5>("PRMT", 1); run<6>("FMNMX+FMUL", 2); run<4>("MUFU.LG2", 1); run<8>("FMUL+MUFU.SIN", 2); run<9>("FFMA+IMAD", 2); run<10>("FFMA+LOP3x2", 3);
    run<20>("16 FFMA", 16); run<21>("15 FFMA + MUFU.SIN", 17); run<22>("14 FFMA + SIN + COS", 18); run<23>("8 FFMA+4 FADD+3 FMNMX+SIN+COS", 19);
    return 0;
}
