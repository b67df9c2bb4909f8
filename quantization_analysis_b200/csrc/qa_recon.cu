// Reconstruction kernels: quantize->dequantize for a set of formats, and per-tile apply.
// One thread owns one 16-element shared-exponent group (32 B of bf16): the group maximum is a
// register reduction, loads/stores are 256-bit, and no lane ever needs another lane's data.
#include <cstdarg>
#include <cstdio>

#include "qa_common.cuh"

namespace qa {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

template <int DT, bool VEC>
__device__ __forceinline__ void load_group(const void* x, int64_t row, int64_t col0, int64_t cols,
                                           int64_t ld, uint32_t (&u)[GROUP]) {
    if (VEC) {
        if (DT == QA_DT_BF16) {
            uint32_t w[8];
            ldg256(reinterpret_cast<const uint16_t*>(x) + row * ld + col0, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                u[2 * i] = w[i] << 16;
                u[2 * i + 1] = w[i] & 0xFFFF0000u;
            }
        } else {
            uint32_t w[8];
            const uint32_t* p = reinterpret_cast<const uint32_t*>(x) + row * ld + col0;
            ldg256(p, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) u[i] = w[i];
            ldg256(p + 8, w);
#pragma unroll
            for (int i = 0; i < 8; ++i) u[8 + i] = w[i];
        }
    } else {
        load_group_scalar<DT>(x, row, col0, cols, ld, u);
    }
}

template <bool VEC>
__device__ __forceinline__ void store_group_bf16(void* out, int64_t row, int64_t col0, int64_t cols,
                                                 const uint32_t (&y)[GROUP]) {
    uint16_t* o = reinterpret_cast<uint16_t*>(out) + row * cols + col0;
    if (VEC) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = (y[2 * i] >> 16) | (y[2 * i + 1] & 0xFFFF0000u);
        stg256(o, w);
    } else {
#pragma unroll
        for (int i = 0; i < GROUP; ++i)
            if (col0 + i < cols) o[i] = (uint16_t)(y[i] >> 16);
    }
}

struct OutPtrs {
    void* p[QA_NFMT];
};

template <int DT, bool VEC>
__global__ void __launch_bounds__(256) recon_kernel(const void* __restrict__ x, int64_t rows,
                                                    int64_t cols, int64_t ld, int64_t gpr,
                                                    uint32_t fmt_mask, OutPtrs out) {
    const int64_t total = rows * gpr;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / gpr;
        const int64_t col0 = (idx - row * gpr) * GROUP;
        uint32_t u[GROUP], y[GROUP];
        load_group<DT, VEC>(x, row, col0, cols, ld, u);
        const uint32_t E = group_max_exp(u);
        if (fmt_mask & 1u) {
#pragma unroll
            for (int i = 0; i < GROUP; ++i) y[i] = bf16_rne_bits(u[i]);
            store_group_bf16<VEC>(out.p[0], row, col0, cols, y);
        }
        if (fmt_mask & 2u) {
#pragma unroll
            for (int i = 0; i < GROUP; ++i) y[i] = bfp_recon_bits<7>(u[i], E);
            store_group_bf16<VEC>(out.p[1], row, col0, cols, y);
        }
        if (fmt_mask & 4u) {
#pragma unroll
            for (int i = 0; i < GROUP; ++i) y[i] = bfp_recon_bits<3>(u[i], E);
            store_group_bf16<VEC>(out.p[2], row, col0, cols, y);
        }
        if (fmt_mask & 8u) {
#pragma unroll
            for (int i = 0; i < GROUP; ++i) y[i] = bfp_recon_bits<1>(u[i], E);
            store_group_bf16<VEC>(out.p[3], row, col0, cols, y);
        }
    }
}

template <int DT, bool VEC>
__global__ void __launch_bounds__(256) apply_kernel(const void* __restrict__ x, int64_t rows,
                                                    int64_t cols, int64_t ld, int64_t gpr,
                                                    int64_t tiles_w,
                                                    const int8_t* __restrict__ assignment,
                                                    void* __restrict__ out) {
    const int64_t total = rows * gpr;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / gpr;
        const int64_t g = idx - row * gpr;
        const int64_t col0 = g * GROUP;
        const int fmt = assignment[(row / TILE) * tiles_w + (g >> 1)];
        uint32_t u[GROUP], y[GROUP];
        load_group<DT, VEC>(x, row, col0, cols, ld, u);
        const uint32_t E = group_max_exp(u);
#pragma unroll
        for (int i = 0; i < GROUP; ++i) y[i] = fmt <= 0 ? bf16_rne_bits(u[i]) : recon_bits(fmt, u[i], E);
        store_group_bf16<VEC>(out, row, col0, cols, y);
    }
}

__global__ void __launch_bounds__(256) f32_to_bf16_checked_kernel(const float* __restrict__ x, int64_t n,
                                                                  uint16_t* __restrict__ out,
                                                                  unsigned long long* inexact) {
    unsigned int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t u = __float_as_uint(x[i]);
        bad += (u & 0xFFFFu) != 0u;
        out[i] = (uint16_t)(bf16_rne_bits(u) >> 16);
    }
    bad = __reduce_add_sync(0xFFFFFFFFu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(inexact, (unsigned long long)bad);
}

static inline bool vec_ok(const void* x, int dt, int64_t cols, int64_t ld, void* const* outs, int nout) {
    if (cols % GROUP || ld % GROUP) return false;
    if (reinterpret_cast<uintptr_t>(x) % 32) return false;
    (void)dt;
    for (int i = 0; i < nout; ++i)
        if (outs[i] && reinterpret_cast<uintptr_t>(outs[i]) % 32) return false;
    return true;
}

static inline int grid_for(int64_t items, int block) {
    int64_t g = cdiv(items, block);
    const int64_t cap = 148 * 16;  // 148 SMs x resident CTAs; grid-stride beyond that
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace qa

using namespace qa;

extern "C" int qa_version(void) { return 100; }
extern "C" const char* qa_last_error(void) { return g_err; }

extern "C" int qa_quant_recon(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ld,
                              uint32_t fmt_mask, void* const out[QA_NFMT], qa_stream_t stream) {
    if (rows < 0 || cols < 0 || ld < cols) { set_error("qa_quant_recon: bad shape"); return 1; }
    if (x_dtype != QA_DT_BF16 && x_dtype != QA_DT_F32) { set_error("qa_quant_recon: bad dtype"); return 1; }
    fmt_mask &= 0xFu;
    if (rows == 0 || cols == 0 || fmt_mask == 0) return 0;
    OutPtrs o;
    for (int f = 0; f < QA_NFMT; ++f) {
        o.p[f] = (fmt_mask >> f) & 1u ? out[f] : nullptr;
        if (((fmt_mask >> f) & 1u) && !out[f]) { set_error("qa_quant_recon: null output for format %d", f); return 1; }
    }
    const int64_t gpr = cdiv(cols, GROUP);
    const bool vec = vec_ok(x, x_dtype, cols, ld, o.p, QA_NFMT);
    const int grid = grid_for(rows * gpr, 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (x_dtype == QA_DT_BF16) {
        if (vec) recon_kernel<QA_DT_BF16, true><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, fmt_mask, o);
        else recon_kernel<QA_DT_BF16, false><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, fmt_mask, o);
    } else {
        if (vec) recon_kernel<QA_DT_F32, true><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, fmt_mask, o);
        else recon_kernel<QA_DT_F32, false><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, fmt_mask, o);
    }
    return check_launch("qa_quant_recon");
}

extern "C" int qa_apply_assignment(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ld,
                                   const int8_t* assignment, void* out_bf16, qa_stream_t stream) {
    if (rows < 0 || cols < 0 || ld < cols || !assignment || !out_bf16) { set_error("qa_apply_assignment: bad args"); return 1; }
    if (rows == 0 || cols == 0) return 0;
    const int64_t gpr = cdiv(cols, GROUP);
    const int64_t tiles_w = cdiv(cols, TILE);
    void* outs[1] = {out_bf16};
    const bool vec = vec_ok(x, x_dtype, cols, ld, outs, 1);
    const int grid = grid_for(rows * gpr, 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (x_dtype == QA_DT_BF16) {
        if (vec) apply_kernel<QA_DT_BF16, true><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, tiles_w, assignment, out_bf16);
        else apply_kernel<QA_DT_BF16, false><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, tiles_w, assignment, out_bf16);
    } else if (x_dtype == QA_DT_F32) {
        if (vec) apply_kernel<QA_DT_F32, true><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, tiles_w, assignment, out_bf16);
        else apply_kernel<QA_DT_F32, false><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, tiles_w, assignment, out_bf16);
    } else { set_error("qa_apply_assignment: bad dtype"); return 1; }
    return check_launch("qa_apply_assignment");
}

extern "C" int qa_f32_to_bf16_checked(const float* x, int64_t n, void* out_bf16,
                                      unsigned long long* inexact_count, qa_stream_t stream) {
    if (n < 0 || !inexact_count) { set_error("qa_f32_to_bf16_checked: bad args"); return 1; }
    if (n == 0) return 0;
    f32_to_bf16_checked_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        x, n, reinterpret_cast<uint16_t*>(out_bf16), inexact_count);
    return check_launch("qa_f32_to_bf16_checked");
}
