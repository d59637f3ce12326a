// Pipe-throughput probes for the instruction mix of the bilateral kernel on sm_100a.
// Each kernel runs a long unrolled chain of one instruction kind with 8 independent streams per
// thread; reports warp-instructions per clock per SM.  Build: tools/build_microbench.sh
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 4096
#define ILP 8

template <int KIND>
__global__ void probe(uint32_t *out, uint32_t seed, float fs)
{
    uint32_t a[ILP];
    float f[ILP];
    unsigned long long p[ILP];
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        a[i] = seed * (i + 1) + threadIdx.x;
        f[i] = fs * (i + 1) + threadIdx.x * 1e-3f;
        p[i] = ((unsigned long long)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] * 0.5f);
    }
    const float c0 = fs * 0.999f, c1 = fs * 0.001f;
    unsigned long long pc = ((unsigned long long)__float_as_uint(c0) << 32) | __float_as_uint(c0);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) f[i] = fmaf(f[i], c0, c1);                                       // FFMA
            if (KIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pc)); // FFMA2
            if (KIND == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));           // MUFU.EX2
            if (KIND == 3) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a[i]) : "r"(seed), "r"(0x4B000000u));
            if (KIND == 4) a[i] = __byte_perm(a[i], seed, 0x7440);                           // PRMT
            if (KIND == 5) { float t; asm volatile("cvt.rn.f32.u8 %0, %1;" : "=f"(t) : "r"(a[i] & 0xff)); a[i] = __float_as_uint(t) >> 20; }  // I2F.U8 (+SHF)
            if (KIND == 6) a[i] = a[i] * seed + 12345u;                                       // IMAD
            if (KIND == 7) a[i] = (a[i] + seed) ^ 0x5bd1e995u;                                 // IADD3+LOP3
            if (KIND == 8) f[i] = sm[(__float_as_uint(f[i]) >> 3) & 1023] + c1;                 // LDS random + FADD
            if (KIND == 9) f[i] = f[i] + c1;                                                   // FADD
            if (KIND == 11) { double dd = (double)f[i]; f[i] = __uint_as_float((uint32_t)__double2hiint(dd)) ; }   // F2F.F64.F32 (+MOV)
            if (KIND == 12) { double dd = (double)a[i]; a[i] = (uint32_t)__double2hiint(dd) + (uint32_t)__double2loint(dd); } // I2F.F64.U32 (+IADD)
            if (KIND == 13) { double dd = __hiloint2double(a[i], a[i] ^ seed); f[i] = (float)dd; a[i] += 1; }    // F2F.F32.F64
            if (KIND == 14) { double dd = __hiloint2double(0x40000000 | (a[i] & 0xffff), a[i]); dd = dd * 1.0000001 + 0.5; a[i] = (uint32_t)__double2hiint(dd) ^ (uint32_t)__double2loint(dd); } // DFMA
            if (KIND == 15) { float t; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(t) : "r"(a[i])); a[i] = __float_as_uint(t) >> 3; }  // I2FP.F32.U32 (+SHF)
            if (KIND == 16) a[i] = __shfl_up_sync(0xffffffffu, a[i], 1) + 1;   // SHFL (+IADD)
            if (KIND == 17) { sm[(threadIdx.x + it) & 1023] = f[i]; f[i] += c1; }    // STS.32 (+FADD)
            if (KIND == 18) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(a[(i + 1) % ILP]), "r"(seed));  // IDP.4A
            if (KIND == 19) a[i] = __funnelshift_r(a[i], seed, 24) + 1;                        // SHF (+IADD)
            if (KIND == 20) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pc));   // FADD2
            if (KIND == 21) { a[i] = a[i] * seed + 12345u; f[i] = fmaf(f[i], c0, c1); }        // IMAD + FFMA
            if (KIND == 22) { asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(a[(i + 1) % ILP]), "r"(seed)); f[i] = fmaf(f[i], c0, c1); }  // IDP.4A + FFMA
            if (KIND == 23) { a[i] = __float_as_uint(sm[a[i] & 255]) + it; }                 // LDS.32, 256-entry table, chained random index
            if (KIND == 24) { a[i] = __float_as_uint(sm[(a[i] & 0) + (it & 255)]) + a[i]; }   // LDS.32 broadcast (all lanes one address)
            if (KIND == 25) { a[i] = __float_as_uint(sm[((threadIdx.x & 31) + it + i) & 255]) + a[i]; }  // LDS.32 consecutive lanes
            if (KIND == 26) { a[i] = __float_as_uint(sm[((threadIdx.x & 31) / 4 + it + i) & 255]) + a[i]; }  // LDS.32 4 lanes per address, 8 addresses
            if (KIND == 10) { float4 v = *reinterpret_cast<float4 *>(&sm[((threadIdx.x * 4) + (it & 7) * 128) & 1020]); f[i] += v.x + v.w; } // LDS.128 + 2 FADD
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r ^= a[i] ^ __float_as_uint(f[i]) ^ (uint32_t)p[i] ^ (uint32_t)(p[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int KIND>
void run(const char *name, int sms, double per_iter)
{
    uint32_t *d;
    const int blocks = sms * 2, threads = 512;
    cudaMalloc(&d, blocks * threads * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<KIND><<<blocks, threads, 4096>>>(d, 3u, 1.0f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        probe<KIND><<<blocks, threads, 4096>>>(d, 3u, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double warp_instr = (double)blocks * (threads / 32) * ITERS * ILP * per_iter;
    const double clocks = best * 1e-3 * clk_khz * 1e3;
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at nominal %d MHz)  %7.1f G thread-ops/s\n", name, best,
           warp_instr / clocks / sms, clk_khz / 1000, warp_instr * 32 / (best * 1e-3) / 1e9);
    cudaFree(d);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, cc %d.%d\n", p.name, p.multiProcessorCount, p.major, p.minor);
    const int sms = p.multiProcessorCount;
    run<0>("FFMA", sms, 1);
    run<1>("FFMA2 (fma.f32x2)", sms, 1);
    run<2>("MUFU.EX2", sms, 1);
    run<3>("VABSDIFF4.ACC", sms, 1);
    run<4>("PRMT", sms, 1);
    run<5>("I2F.U8 (+LOP,+SHF)", sms, 1);
    run<6>("IMAD", sms, 1);
    run<7>("IADD3+LOP3", sms, 2);
    run<8>("LDS.32 random (+FADD,+2)", sms, 1);
    run<9>("FADD", sms, 1);
    run<10>("LDS.128 (+2 FADD)", sms, 1);
    run<11>("F2F.F64.F32", sms, 1);
    run<12>("I2F.F64.U32 (+IADD)", sms, 1);
    run<13>("F2F.F32.F64 (+IADD)", sms, 1);
    run<14>("DFMA (+LOP..)", sms, 1);
    run<15>("I2FP.F32.U32 (+SHF)", sms, 1);
    run<16>("SHFL.UP (+IADD)", sms, 1);
    run<17>("STS.32 (+FADD)", sms, 1);
    run<23>("LDS.32 LUT256 random (+IADD,LOP)", sms, 1);
    run<24>("LDS.32 broadcast (+IADD)", sms, 1);
    run<25>("LDS.32 consecutive (+IADD)", sms, 1);
    run<26>("LDS.32 8 addr x 4 lanes (+IADD)", sms, 1);
    run<18>("IDP.4A", sms, 1);
    run<19>("SHF (+IADD)", sms, 2);
    run<20>("FADD2 (add.f32x2)", sms, 1);
    run<21>("IMAD + FFMA", sms, 2);
    run<22>("IDP.4A + FFMA", sms, 2);
    return 0;
}
