// Throughput of the float64 instructions the greedy kernel leans on (warp-instructions per clock per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dp_ops dp_ops.cu && ./dp_ops
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(double* out, long long* cyc, int iters, double seed) {
    double a = seed + threadIdx.x, b = seed * 0.5 + 1.0, c = 0.25 + threadIdx.x * 1e-3, d = 3.0;
    long long ia = threadIdx.x + 5, ib = 7;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (OP == 0) { a = __dadd_rn(a, b); c = __dadd_rn(c, d); }
            if (OP == 1) { a = __dmul_rn(a, b); c = __dmul_rn(c, d); }
            if (OP == 2) { a = __fma_rn(a, b, d); c = __fma_rn(c, d, b); }
            if (OP == 3) { a = floor(a * 1.0000001) ; c = floor(c + 0.7); }                 // FRND + 1 DP op each
            if (OP == 4) { ia += __double2ll_rd(a); a = a + 1.5; ib += __double2ll_rd(c); c = c + 2.5; }   // F2I.S64.F64 + DADD
            if (OP == 5) { a += (double)ia; ia += 3; c += (double)ib; ib += 5; }                           // I2F.F64.S64 + DADD
            if (OP == 6) { ia += (a < c) ? 1 : 2; a = __dadd_rn(a, 1.0); ib += (c < b) ? 1 : 3; c = __dadd_rn(c, 0.5); }   // DSETP + DADD
            if (OP == 7) { a = __ddiv_rn(a, b) + 3.0; c = __ddiv_rn(c, d) + 2.0; }
            if (OP == 8) { a = __dsqrt_rn(a) + 3.0; c = __dsqrt_rn(c) + 2.0; }
            if (OP == 9) { ia = ia * 3 + ib; ib = ib + (ia >> 3); }                                      // int64 baseline
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + c + (double)(ia + ib);
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
    const char* names[] = {"DADD x2", "DMUL x2", "DFMA x2", "floor(FRND)+DP x2", "F2I.S64.F64+DADD x2", "I2F.F64.S64+DADD x2",
                           "DSETP+DADD x2", "ddiv_rn+DADD x2", "dsqrt_rn+DADD x2", "int64 mul/add/shift"};
    const int iters = 2000;
    for (int threads : {32, 512}) {
        printf("threads per CTA = %d (1 CTA per SM)\n", threads);
        for (int op = 0; op < 10; ++op) {
            for (int rep = 0; rep < 2; ++rep) {
                switch (op) {
                    case 0: k<0><<<148, threads>>>(out, cyc, iters, 1.0); break;
                    case 1: k<1><<<148, threads>>>(out, cyc, iters, 1.0); break;
                    case 2: k<2><<<148, threads>>>(out, cyc, iters, 1.0); break;
                    case 3: k<3><<<148, threads>>>(out, cyc, iters, 1.0); break;
                    case 4: k<4><<<148, threads>>>(out, cyc, iters, 1.0); break;
                    case 5: k<5><<<148, threads>>>(out, cyc, iters, 1.0); break;
                    case 6: k<6><<<148, threads>>>(out, cyc, iters, 1.0); break;
                    case 7: k<7><<<148, threads>>>(out, cyc, iters, 1.0); break;
                    case 8: k<8><<<148, threads>>>(out, cyc, iters, 1.0); break;
                    case 9: k<9><<<148, threads>>>(out, cyc, iters, 1.0); break;
                }
                cudaDeviceSynchronize();
            }
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            const double pairs = (double)iters * 8;      // "op pairs" per thread
            printf("  %-24s %8.1f cycles per pair-of-ops per warp  (%.3f pairs/clk/SM)\n", names[op], (double)h / pairs,
                   pairs * (threads / 32) / (double)h);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
