// Instruction-throughput micro-benchmarks for sm_100a (B200): the numbers behind the
// matcher's design choices in DESIGN.md (which pipe each exact-sum formulation lands on).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench scripts/ubench.cu
// Prints lane-ops per clock per SM for each instruction class.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed, float fa, double da) {
    uint32_t x[CHAINS];
    float f[CHAINS];
    double d[CHAINS];
    unsigned long long w[CHAINS];
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 256) sm[i] = i * seed;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { x[c] = threadIdx.x * 7 + c + seed; f[c] = (float)x[c]; d[c] = (double)x[c]; w[c] = x[c]; }
    unsigned long long wa;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(wa) : "f"(fa));
    if (OP >= 24) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(w[c]) : "f"(f[c]), "f"(f[(c + 3) % CHAINS]));
    }
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&sm[(threadIdx.x * 9) & 1023]) & ~7u;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(fa), "f"(f[(c + 1) % CHAINS]));
            if (OP == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(seed), "r"(x[(c + 1) % CHAINS]));
            if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[c]) : "r"(x[c]), "r"(seed));
            if (OP == 3) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x[c]) : "r"(x[(c + 1) % CHAINS]), "r"(seed));
            if (OP == 4) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(da));
            if (OP == 5) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[c]) : "d"(da));
            if (OP == 6) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[c])); asm volatile("" ::"d"(t)); d[c] = t; }
            if (OP == 7) { asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f[c]) : "r"(x[c])); }
            if (OP == 8) { asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(x[c]) : "f"(f[c])); }
            if (OP == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(seed), "r"(x[(c + 1) % CHAINS]));
            if (OP == 10) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(x[(c + 1) % CHAINS]), "r"(seed));
            if (OP == 11) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(x[(c + 1) % CHAINS]), "r"(seed));
            if (OP == 12) { uint32_t t; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"((uint32_t)__cvta_generic_to_shared(&sm[(x[c] + threadIdx.x) & 2047]))); x[c] += t; }
            if (OP == 13) { uint4 t; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"((uint32_t)__cvta_generic_to_shared(&sm[((x[c] + threadIdx.x) * 4) & 2047]))); x[c] += t.x ^ t.y ^ t.z ^ t.w; }
            if (OP == 14) x[c] = __shfl_xor_sync(0xffffffffu, x[c], 1);
            if (OP == 15) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[c]) : "f"(f[(c + 1) % CHAINS]));
            if (OP == 16) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(x[(c + 1) % CHAINS]));
            if (OP == 17) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[c]) : "f"(fa)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(seed), "r"(x[(c + 1) % CHAINS])); }
            if (OP == 18) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(fa), "f"(f[(c + 1) % CHAINS])); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(seed), "r"(x[(c + 1) % CHAINS])); }
            if (OP == 19) { asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x[c]) : "r"(x[(c + 1) % CHAINS]), "r"(seed)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[c]) : "f"(f[(c + 1) % CHAINS])); }
            if (OP == 20) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[c])); asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(t)); asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[c]) : "f"(fa)); }
            if (OP == 21) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[c]) : "r"(x[c])); }
            if (OP == 22) { unsigned long long t; asm volatile("cvt.rzi.s64.f32 %0, %1;" : "=l"(t) : "f"(f[c])); w[c] += t; }
            if (OP == 23) { asm volatile("popc.b32 %0, %0;" : "+r"(x[c])); }
            // packed FP32 (sm_100+): two lanes of FP32 per 64-bit register pair
            if (OP == 24) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[c]) : "l"(w[(c + 1) % CHAINS]));
            if (OP == 25) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(w[c]) : "l"(wa), "l"(w[(c + 1) % CHAINS]));
            if (OP == 26) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(w[c]) : "l"(wa));
            // the matcher's inner loop per pixel: LDS + FMUL + Fast2Sum (4 FADD), scalar ...
            if (OP == 27) {
                float sv, p, t, z, e;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(sv) : "r"(sbase + c * 4));
                asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(p) : "f"(fa), "f"(sv));
                asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(t) : "f"(f[c & 1]), "f"(p));
                asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(z) : "f"(t), "f"(f[c & 1]));
                asm volatile("sub.rn.f32 %0, %1, %2;" : "=f"(e) : "f"(p), "f"(z));
                asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[2 + (c & 1)]) : "f"(e));
                f[c & 1] = t;
            }
            // ... and packed, two pixels per instruction (two LDS.32 into a register pair)
            if (OP == 28 && (c & 1) == 0) {
                unsigned long long sv, p, t, z, e;
                float s0, s1;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(s0) : "r"(sbase + c * 4));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(s1) : "r"(sbase + c * 4 + 4));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(sv) : "f"(s0), "f"(s1));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(wa), "l"(sv));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(w[(c >> 1) & 1]), "l"(p));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(t), "l"(w[(c >> 1) & 1]));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(e) : "l"(p), "l"(z));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[2 + ((c >> 1) & 1)]) : "l"(e));
                w[(c >> 1) & 1] = t;
            }
            // same with one LDS.64 per pixel pair
            if (OP == 29 && (c & 1) == 0) {
                unsigned long long sv, p, t, z, e;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(sv) : "r"(sbase + c * 4));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(wa), "l"(sv));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(w[(c >> 1) & 1]), "l"(p));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(t), "l"(w[(c >> 1) & 1]));
                asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(e) : "l"(p), "l"(z));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(w[2 + ((c >> 1) & 1)]) : "l"(e));
                w[(c >> 1) & 1] = t;
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc += x[c] + (uint32_t)f[c] + (uint32_t)d[c] + (uint32_t)w[c];
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int OP>
void run(const char *name, int ops_per_iter, uint32_t *out, int sms, double mhz) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 8;
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        k<OP><<<blocks, 256>>>(out, 12345u + r, 1.0000001f, 1.0000001);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    double lane_ops = (double)blocks * 256 * ITERS * CHAINS * ops_per_iter;
    double per_clk_sm = lane_ops / (best * 1e-3) / (mhz * 1e6) / sms;
    printf("%-28s %8.3f ms  %7.1f lane-ops/clk/SM (at %.0f MHz)  %8.2f Tops/s\n", name, best, per_clk_sm, mhz, lane_ops / (best * 1e-3) / 1e12);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double mhz = khz / 1000.0;
    printf("%s  SMs=%d  clock=%.0f MHz\n", p.name, sms, mhz);
    uint32_t *out; cudaMalloc(&out, (size_t)sms * 8 * 256 * 4);
    run<0>("FFMA", 1, out, sms, mhz);
    run<15>("FADD", 1, out, sms, mhz);
    run<1>("IMAD.lo", 1, out, sms, mhz);
    run<2>("IMAD.WIDE.U32 (64b acc)", 1, out, sms, mhz);
    run<3>("IDP4A.u8.u8", 1, out, sms, mhz);
    run<16>("IADD", 1, out, sms, mhz);
    run<9>("LOP3", 1, out, sms, mhz);
    run<10>("SHF", 1, out, sms, mhz);
    run<11>("PRMT", 1, out, sms, mhz);
    run<23>("POPC", 1, out, sms, mhz);
    run<4>("DADD", 1, out, sms, mhz);
    run<5>("DFMA", 1, out, sms, mhz);
    run<6>("F2F.F64.F32", 1, out, sms, mhz);
    run<7>("I2F.u32", 1, out, sms, mhz);
    run<21>("I2F.s32", 1, out, sms, mhz);
    run<8>("F2I.u32", 1, out, sms, mhz);
    run<22>("F2I.s64+IADD64", 1, out, sms, mhz);
    run<12>("LDS.32 (+IADD)", 1, out, sms, mhz);
    run<13>("LDS.128 (+3LOP+IADD)", 1, out, sms, mhz);
    run<14>("SHFL", 1, out, sms, mhz);
    run<17>("FMUL+IMAD pair", 2, out, sms, mhz);
    run<18>("FFMA+LOP3 pair", 2, out, sms, mhz);
    run<19>("IDP4A+FADD pair", 2, out, sms, mhz);
    run<20>("FMUL+F2D+DADD triple", 3, out, sms, mhz);
    printf("packed FP32: lane-ops count each FP32 lane (2 per instruction)\n");
    run<24>("FADD2", 2, out, sms, mhz);
    run<25>("FFMA2", 2, out, sms, mhz);
    run<26>("FMUL2", 2, out, sms, mhz);
    printf("matcher inner loop, pixels/clk/SM\n");
    run<27>("px: LDS+FMUL+4FADD", 1, out, sms, mhz);
    run<28>("px: 2LDS+FMUL2+4FADD2 (/2px)", 1, out, sms, mhz);
    run<29>("px: LDS64+FMUL2+4FADD2 (/2px)", 1, out, sms, mhz);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
