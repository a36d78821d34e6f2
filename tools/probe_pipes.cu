// probe_pipes.cu -- measures the issue rates the planning kernel's roofline is quoted against (tools only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_probe_pipes tools/probe_pipes.cu && tools/_probe_pipes
// Each kernel runs 8 independent dependency chains per thread, 1024 threads per SM resident (8 warps per scheduler),
// so the loop is bound by the pipe's issue rate, not by latency.  Output: one JSON object (warp instructions per
// cycle per SM sub-partition).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 16384
template <int KIND>
__global__ void __launch_bounds__(256) probe(uint64_t *out, uint32_t m0, uint32_t m1) {
    uint64_t a[8];
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        a[k] = threadIdx.x * 977u + k;
        x[k] = threadIdx.x * 31u + k;
    }
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (KIND == 0) {  // IMAD.WIDE.U32: 32x32 -> 64 with a 64-bit addend
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(x[k]), "r"(m0));
            } else if (KIND == 1) {  // IMAD (low 32 bits)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(m0), "r"(m1));
            } else if (KIND == 2) {  // LOP3
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(m0), "r"(m1));
            } else if (KIND == 3) {  // the Philox mix: 4 wide multiplies + 3 integer ALU ops
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(x[k]), "r"(m0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(m0), "r"(m1));
            } else if (KIND == 5) {  // IMAD.WIDE.U32 without addend
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(a[k]) : "r"(x[k]), "r"(m0));
            } else if (KIND == 6) {  // wide multiply-add and an independent LOP3
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(m1), "r"(m0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(m0), "r"(m1));
            } else if (KIND == 7) {  // IMAD.HI
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(m0), "r"(m1));
            } else if (KIND == 8) {  // wide multiply without addend and an independent LOP3
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(a[k]) : "r"(m1), "r"(m0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(m0), "r"(m1));
            } else if (KIND == 9) {  // wide multiply-add and two independent LOP3
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(m1), "r"(m0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(m0), "r"(m1));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(m1), "r"(m0));
            } else {  // DFMA
                double d = __longlong_as_double((long long)a[k]);
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(1.0000001), "d"(0.5));
                a[k] = (uint64_t)__double_as_longlong(d);
            }
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += a[k] + x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
static double run(const char *name, int ops_per_iter, int sms, double mhz) {
    uint64_t *out;
    const int blocks = sms * 4;
    cudaMalloc(&out, sizeof(uint64_t) * blocks * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<KIND><<<blocks, 256>>>(out, 0xE14C6C93u, 0xD2E7470Eu);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        probe<KIND><<<blocks, 256>>>(out, 0xE14C6C93u, 0xD2E7470Eu);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaFree(out);
    const double warp_inst = (double)blocks * 8 * ITER * 8 * ops_per_iter;  // 8 warps per CTA, 8 chains
    const double per_s = warp_inst / (best * 1e-3);
    const double per_clk_smsp = per_s / (sms * 4.0 * mhz * 1e6);
    printf("  \"%s\": {\"ms\": %.4f, \"warp_inst_per_s\": %.4e, \"warp_inst_per_clk_per_smsp\": %.4f},\n", name, best, per_s,
           per_clk_smsp);
    return per_s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"sm_mhz_max\": %.0f,\n", p.name, p.multiProcessorCount, mhz);
    run<0>("imad_wide_u32", 1, p.multiProcessorCount, mhz);
    run<1>("imad_lo", 1, p.multiProcessorCount, mhz);
    run<2>("lop3", 1, p.multiProcessorCount, mhz);
    run<3>("imad_wide_plus_lop3", 2, p.multiProcessorCount, mhz);
    run<4>("dfma", 1, p.multiProcessorCount, mhz);
    run<6>("imad_wide_plus_independent_lop3", 2, p.multiProcessorCount, mhz);
    run<7>("imad_hi", 1, p.multiProcessorCount, mhz);
    run<9>("imad_wide_plus_two_independent_lop3", 3, p.multiProcessorCount, mhz);
    printf("  \"note\": \"8 independent chains per thread, 4 CTAs x 256 threads per SM, best of 5\"\n}\n");
    return 0;
}
