// Issue / pipe throughput of the instructions the tile-stat kernel can be built from (sm_100a).
// Every test runs 8 independent chains per thread, 1024 threads per CTA (8 warps per SMSP), one CTA per SM,
// and reports cycles per warp-instruction per SMSP (1.0 = one instruction issued every clock).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_ops pipe_ops.cu && ./pipe_ops
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op {
    FFMA, FFMA_IMM, FADD, FMUL, FFMA2, FADD2, FMUL2, FMNMX, FMNMX_XS, FMNMX3,
    HFMA2, HFMA2_BF, HADD2, HADD2_BF, HMUL2, HMNMX2, HMNMX2_BF, HMNMX2_XS, HSET2,
    LOP3, SHL, IADD3, PRMT, IMAD, IMADW, VIMNMX, VIMNMX16, VIMNMX3_16, VIADD16, DP4A, DP2A, POPC, FLO,
    F2F64, F2FP_H, F2FP_BF, H2F, I2F, F2I, I2FP_pack, DADD, DFMA,
    SHFL, LDS,
    MIX_FFMA2_LOP3, MIX_HFMA2_LOP3, MIX_FFMA2_HFMA2, MIX_FFMA2_FMNMX, MIX_FFMA2_DFMA, MIX_HFMA2_DFMA, MIX_HFMA2_VIMNMX,
    MIX_FFMA2_DP4A, MIX_HFMA2_DP4A, MIX_LOP3_DP4A, MIX_FFMA_FFMA2, MIX_HFMA2_F2F64, MIX_3WAY,
    MIX_FFMA_LOP3, MIX_FADD_FMNMX, MIX_2FFMA_LOP3, MIX_FADD_LOP3_SHL, NOPS
};

template <int OP>
__device__ __forceinline__ void step(uint32_t (&r)[8], uint64_t (&q)[8], uint32_t c0, uint32_t c1, uint64_t d0, uint64_t d1, uint32_t* sm) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        uint32_t& a = r[u];
        uint64_t& b = q[u];
        if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == FFMA_IMM) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+r"(a) : "r"(c1));
        if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b) : "l"(d0), "l"(d1));
        if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b) : "l"(d0));
        if (OP == FMUL2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(b) : "l"(d0));
        if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == FMNMX_XS) asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == FMNMX3) asm volatile("max.NaN.abs.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == HFMA2_BF) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == HADD2) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == HADD2_BF) asm volatile("add.rn.bf16x2 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == HMUL2) asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == HMNMX2) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == HMNMX2_BF) asm volatile("min.bf16x2 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == HMNMX2_XS) asm volatile("min.xorsign.abs.f16x2 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == HSET2) asm volatile("set.gt.f16x2.f16x2 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == SHL) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == IADD3) asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == IMADW) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(b) : "r"(a), "r"(c1));
        if (OP == VIMNMX) asm volatile("max.s32 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == VIMNMX16) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == VIMNMX3_16) a = __vimax3_u16x2(a, c0, c1);
        if (OP == VIADD16) a = __vadd2(a, c0);
        if (OP == DP4A) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == DP2A) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == POPC) asm volatile("{.reg .b32 t; popc.b32 t, %0; add.s32 %0, t, %1;}" : "+r"(a) : "r"(c0));
        if (OP == FLO) asm volatile("{.reg .b32 t; clz.b32 t, %0; add.s32 %0, t, %1;}" : "+r"(a) : "r"(c0));
        if (OP == F2F64) asm volatile("{.reg .f64 t; cvt.f64.f32 t, %1; add.rn.f64 %0, %0, t;}" : "+l"(b) : "r"(a));
        if (OP == F2FP_H) asm volatile("cvt.rn.f16x2.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == F2FP_BF) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
        if (OP == H2F) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; cvt.f32.f16 %0, lo;}" : "+r"(a));
        if (OP == I2F) asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(a));
        if (OP == F2I) asm volatile("cvt.rni.s32.f32 %0, %0;" : "+r"(a));
        if (OP == I2FP_pack) asm volatile("cvt.pack.sat.s8.s32.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        if (OP == DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+l"(b) : "l"(d0));
        if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+l"(b) : "l"(d0), "l"(d1));
        if (OP == SHFL) a = __shfl_xor_sync(0xFFFFFFFFu, a, 1);
        if (OP == LDS) a = sm[(a & 1023u)];
        // two-instruction mixes: one of each per chain step
        if (OP == MIX_FFMA2_LOP3) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b) : "l"(d0), "l"(d1));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_HFMA2_LOP3) {
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
            uint32_t& a2 = reinterpret_cast<uint32_t*>(&b)[0];
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_FFMA2_HFMA2) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b) : "l"(d0), "l"(d1));
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_FFMA2_FMNMX) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b) : "l"(d0), "l"(d1));
            asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
        }
        if (OP == MIX_FFMA2_DFMA) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b) : "l"(d0), "l"(d1));
            uint64_t& b2 = q[(u + 4) & 7];
            if (u < 4) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+l"(b2) : "l"(d0), "l"(d1));
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_HFMA2_DFMA) {
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+l"(b) : "l"(d0), "l"(d1));
        }
        if (OP == MIX_HFMA2_VIMNMX) {
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
            uint32_t& a2 = reinterpret_cast<uint32_t*>(&b)[0];
            asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a2) : "r"(c0));
        }
        if (OP == MIX_FFMA2_DP4A) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b) : "l"(d0), "l"(d1));
            asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_HFMA2_DP4A) {
            uint32_t& a2 = reinterpret_cast<uint32_t*>(&b)[0];
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a2) : "r"(c0), "r"(c1));
            asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_LOP3_DP4A) {
            uint32_t& a2 = reinterpret_cast<uint32_t*>(&b)[0];
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(c0), "r"(c1));
            asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_FFMA_FFMA2) {
            asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b) : "l"(d0), "l"(d1));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_HFMA2_F2F64) {
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
            if ((u & 3) == 0) asm volatile("{.reg .f64 t; cvt.f64.f32 t, %1; add.rn.f64 %0, %0, t;}" : "+l"(b) : "r"(c0));
        }
        if (OP == MIX_FFMA_LOP3) {      // scalar fp32 + ALU: do these two pipes overlap when the fp32 side is not packed?
            uint32_t& a2 = reinterpret_cast<uint32_t*>(&b)[0];
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_FADD_FMNMX) {
            uint32_t& a2 = reinterpret_cast<uint32_t*>(&b)[0];
            asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
            asm volatile("min.xorsign.abs.f32 %0, %0, %1;" : "+r"(a2) : "r"(c0));
        }
        if (OP == MIX_2FFMA_LOP3) {     // two scalar fp32 per ALU instruction
            uint32_t& a2 = reinterpret_cast<uint32_t*>(&b)[0];
            uint32_t& a3 = reinterpret_cast<uint32_t*>(&b)[1];
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a3) : "r"(c0), "r"(c1));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_FADD_LOP3_SHL) {  // one scalar fp32 per two ALU instructions
            uint32_t& a2 = reinterpret_cast<uint32_t*>(&b)[0];
            uint32_t& a3 = reinterpret_cast<uint32_t*>(&b)[1];
            asm volatile("add.rn.f32 %0, %0, %1;" : "+r"(a) : "r"(c0));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(c0), "r"(c1));
            asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a3) : "r"(c0), "r"(c1));
        }
        if (OP == MIX_3WAY) {   // HFMA2 + LOP3 + (every other) DFMA
            asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(a) : "r"(c0), "r"(c1));
            uint32_t& a2 = reinterpret_cast<uint32_t*>(&q[(u + 1) & 7])[1];
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(c0), "r"(c1));
        }
    }
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, long long* cyc, int iters, uint32_t s0, uint32_t s1) {
    __shared__ uint32_t sm[1024];
    sm[threadIdx.x] = threadIdx.x * 7u;
    uint32_t r[8];
    uint64_t q[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        r[u] = 0x3C003C00u + threadIdx.x + u * s0;
        q[u] = ((uint64_t)(0x3F800000u + u) << 32) | (0x3F800000u + threadIdx.x * s1);
    }
    const uint32_t c0 = 0x3C013C01u + s0, c1 = 0x38003800u + s1;
    const uint64_t d0 = 0x3FF0000000000001ull + s0, d1 = 0x3F8000013F800001ull + s1;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) step<OP>(r, q, c0, c1, d0, d1, sm);
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc ^= r[u] ^ (uint32_t)q[u] ^ (uint32_t)(q[u] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct T { int op; const char* name; int instr_per_step; };

template <int OP>
void run(const char* name, int per_step, uint32_t* out, long long* cyc) {
    const int iters = 512;
    for (int rep = 0; rep < 2; ++rep) { k<OP><<<148, 1024>>>(out, cyc, iters, 0, 0); cudaDeviceSynchronize(); }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148;
    const double instr_per_smsp = (double)iters * 8 * per_step * 8;   // 8 warps per SMSP
    printf("%-22s %7.3f cycles/warp-instr/SMSP  (%d instr per chain step)  err=%s\n", name, avg / instr_per_smsp, per_step,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4 * 148 * 1024); cudaMalloc(&cyc, 8 * 148);
#define R(op, n) run<op>(#op, n, out, cyc)
    R(FFMA, 1); R(FFMA_IMM, 1); R(FADD, 1); R(FMUL, 1); R(FFMA2, 1); R(FADD2, 1); R(FMUL2, 1); R(FMNMX, 1); R(FMNMX_XS, 1); R(FMNMX3, 1);
    R(HFMA2, 1); R(HFMA2_BF, 1); R(HADD2, 1); R(HADD2_BF, 1); R(HMUL2, 1); R(HMNMX2, 1); R(HMNMX2_BF, 1); R(HMNMX2_XS, 1); R(HSET2, 1);
    R(LOP3, 1); R(SHL, 1); R(IADD3, 1); R(PRMT, 1); R(IMAD, 1); R(IMADW, 1); R(VIMNMX, 1); R(VIMNMX16, 1); R(VIMNMX3_16, 1); R(VIADD16, 1);
    R(DP4A, 1); R(DP2A, 1); R(POPC, 2); R(FLO, 2);
    R(F2F64, 2); R(F2FP_H, 1); R(F2FP_BF, 1); R(H2F, 1); R(I2F, 1); R(F2I, 1); R(I2FP_pack, 1); R(DADD, 1); R(DFMA, 1); R(SHFL, 1); R(LDS, 1);
    R(MIX_FFMA2_LOP3, 2); R(MIX_HFMA2_LOP3, 2); R(MIX_FFMA2_HFMA2, 2); R(MIX_FFMA2_FMNMX, 2); R(MIX_FFMA2_DFMA, 2); R(MIX_HFMA2_DFMA, 2);
    R(MIX_HFMA2_VIMNMX, 2); R(MIX_FFMA2_DP4A, 2); R(MIX_HFMA2_DP4A, 2); R(MIX_LOP3_DP4A, 2); R(MIX_FFMA_FFMA2, 2); R(MIX_HFMA2_F2F64, 1); R(MIX_3WAY, 2);
    R(MIX_FFMA_LOP3, 2); R(MIX_FADD_FMNMX, 2); R(MIX_2FFMA_LOP3, 3); R(MIX_FADD_LOP3_SHL, 3);
    return 0;
}
