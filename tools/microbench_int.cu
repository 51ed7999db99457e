// Integer-pipe throughput probe for sm_100a (tools only, not part of the product): how many warp-instructions per clock
// and SM the multiply-high / wide-multiply / funnel-shift / LOP3 forms sustain, alone and mixed.  The gate kernel's hash
// is chosen from these numbers (gate section).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_int microbench_int.cu && ./microbench_int
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

constexpr int kIters = 4096;

template <int OP>
__global__ void __launch_bounds__(1024, 1) probe(uint32_t* out, uint32_t seed, unsigned long long* cycles) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
    uint32_t c = seed | 1u, k = seed * 2654435761u | 1u;
    __syncthreads();
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) a[i] = a[i] * k + c;                                   // IMAD
            if (OP == 1) a[i] = __umulhi(a[i], k) + c;                          // IMAD.HI (+ add folded?)
            if (OP == 2) { unsigned long long w = (unsigned long long)a[i] * k; a[i] = (uint32_t)(w >> 32) ^ (uint32_t)w; }  // IMAD.WIDE + LOP3
            if (OP == 3) a[i] = __funnelshift_r(a[i], c, a[i]);                 // SHF.R.W variable
            if (OP == 4) a[i] = (a[i] & k) ^ c;                                 // LOP3
            if (OP == 5) { a[i] = __umulhi(a[i], k); a[i] = (a[i] & 0x1FFFCu) ^ c; }  // IMAD.HI + LOP3: two pipes
            if (OP == 6) a[i] = __byte_perm(a[i], c, 0x4321);                   // PRMT
            if (OP == 7) { uint32_t h1 = __umulhi(a[i], k), h2 = __umulhi(a[i], c), h3 = __umulhi(a[i], k ^ c);
                           a[i] = (__funnelshift_r(h1, h1, h2) & __funnelshift_r(h1, h1, h3)) | (h1 & 0x1FFFCu); }  // 3 IMAD.HI + 2 SHF + 2 LOP3
            if (OP == 8) a[i] = a[i] + k + c;                                   // IADD3
            if (OP == 9) a[i] = __umulhi(a[i], k);                              // bare IMAD.HI chain
        }
    }
    const unsigned long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_inner, uint32_t* out, unsigned long long* cyc) {
    probe<OP><<<148, 1024>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    probe<OP><<<148, 1024>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto v : h) avg += (double)v;
    avg /= 148;
    const double warp_insts = 32.0 * kIters * 8 * ops_per_inner;  // per SM: 32 warps
    printf("%-44s %8.0f cycles  %.3f warp-instr/clk/SM (counting %d source ops per element)\n", name, avg, warp_insts / avg, ops_per_inner);
}

int main() {
    uint32_t* out;
    unsigned long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    run<0>("IMAD (mul.lo + add)", 1, out, cyc);
    run<1>("IMAD.HI + add", 1, out, cyc);
    run<9>("IMAD.HI bare", 1, out, cyc);
    run<2>("IMAD.WIDE + LOP3", 2, out, cyc);
    run<3>("SHF.R.W variable", 1, out, cyc);
    run<4>("LOP3", 1, out, cyc);
    run<8>("IADD3", 1, out, cyc);
    run<6>("PRMT", 1, out, cyc);
    run<5>("IMAD.HI + LOP3 (two pipes)", 2, out, cyc);
    run<7>("3 IMAD.HI + 2 SHF + 2 LOP3 (gate probe shape)", 7, out, cyc);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
