// Reconstruction kernels: quantize->dequantize for a set of formats, and per-tile apply.
// One thread owns one 16-element shared-exponent group (32 B of bf16): the group maximum is a
// register reduction, loads/stores are 256-bit, and no lane ever needs another lane's data.
#include <cstdarg>
#include <cstdio>

#include <algorithm>

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

// ---------------------------------------------------------------------------------------------
// bf16 fast path.  In "group units" X = x * 2^(134-E) every reconstruction is a small dyadic number:
// y = clamp(rne(X, step)) is (X + 1.5*2^23*step) - 1.5*2^23*step (ties to even, exact for bf16 inputs because no
// significand bit is shifted out before the rounding) followed by min(|y|, (2^mb - 1)*step) with the sign of y;
// scaling back by 2^(E-134) is exact and the upper 16 bits of the float32 pattern are the answer.
// Groups with E < 24 (the trick's constants would leave float32's normal range), E == 0 or E == 255 (inf/nan) use
// the integer recipe.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float clamp_sym(float y, float L) {
    float r;
    asm("min.xorsign.abs.f32 %0, %1, %2;" : "=f"(r) : "f"(y), "f"(L));
    return r;
}
__device__ __forceinline__ float max3abs(float m, float a, float b) {
    float r;
    asm("max.NaN.abs.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(m), "f"(a), "f"(b));   // a NaN must surface: its exponent is 255
    return r;
}
__device__ __forceinline__ uint32_t pack_hi16(float lo, float hi) {      // {bf16(lo), bf16(hi)} of exactly-representable values
    return __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
}
__device__ __forceinline__ void fmt_consts(int fmt, float& M, float& L) {
    M = fmt == 1 ? 25165824.f : fmt == 2 ? 402653184.f : 1610612736.f;     // 1.5 * 2^23 * step (2, 32, 128)
    L = fmt == 1 ? 254.f : fmt == 2 ? 224.f : 128.f;                       // (2^mb - 1) * step
}
// one format of one group: w = 8 packed words in, out = 8 packed words
__device__ __forceinline__ void group_recon_fast(const float2 (&X)[8], float M, float L, float back, uint32_t (&out)[8]) {
    const float2 Mv = make_float2(M, M), nMv = make_float2(-M, -M), bv = make_float2(back, back);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float2 y = __fadd2_rn(__fadd2_rn(X[i], Mv), nMv);
        y.x = clamp_sym(y.x, L);
        y.y = clamp_sym(y.y, L);
        y = __fmul2_rn(y, bv);
        out[i] = pack_hi16(y.x, y.y);
    }
}
// returns false when the group needs the integer recipe
__device__ __forceinline__ bool group_prepare_fast(const uint32_t (&w)[8], float2 (&X)[8], float& back, uint32_t& E) {
    float mabs = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        X[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
        mabs = max3abs(mabs, X[i].x, X[i].y);
    }
    E = __float_as_uint(mabs) >> 23;
    if (E < 24u || E == 255u) return false;
    const float inv = __uint_as_float((261u - E) << 23);   // 2^(134-E)
    back = __uint_as_float((E - 7u) << 23);                 // 2^(E-134)
    const float2 iv = make_float2(inv, inv);
#pragma unroll
    for (int i = 0; i < 8; ++i) X[i] = __fmul2_rn(X[i], iv);
    return true;
}
__device__ __forceinline__ void group_recon_slow(const uint32_t (&w)[8], uint32_t E, int fmt, uint32_t (&out)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t lo = recon_bits(fmt, w[i] << 16, E), hi = recon_bits(fmt, w[i] & 0xFFFF0000u, E);
        out[i] = (lo >> 16) | (hi & 0xFFFF0000u);
    }
}

__global__ void __launch_bounds__(256) recon_fast_kernel(const uint16_t* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                                                         int64_t gpr, uint32_t fmt_mask, OutPtrs out) {
    const int64_t total = rows * gpr;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / gpr;
        const int64_t col0 = (idx - row * gpr) * GROUP;
        uint32_t w[8], y[8];
        ldg256(x + row * ld + col0, w);
        if (fmt_mask & 1u) stg256(reinterpret_cast<uint16_t*>(out.p[0]) + row * cols + col0, w);   // bf16 of bf16 = identity
        float2 X[8];
        float back;
        uint32_t E;
        const bool fast = group_prepare_fast(w, X, back, E);
#pragma unroll
        for (int f = 1; f < QA_NFMT; ++f) {
            if (!((fmt_mask >> f) & 1u)) continue;
            if (fast) {
                float M, L;
                fmt_consts(f, M, L);
                group_recon_fast(X, M, L, back, y);
            } else {
                group_recon_slow(w, E, f, y);
            }
            stg256(reinterpret_cast<uint16_t*>(out.p[f]) + row * cols + col0, y);
        }
    }
}

__global__ void __launch_bounds__(256) apply_fast_kernel(const uint16_t* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                                                         int64_t gpr, int64_t tiles_w, const int8_t* __restrict__ assignment,
                                                         uint16_t* __restrict__ out) {
    const int64_t total = rows * gpr;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = idx / gpr;
        const int64_t g = idx - row * gpr;
        const int fmt = assignment[(row / TILE) * tiles_w + (g >> 1)];
        uint32_t w[8], y[8];
        ldg256(x + row * ld + g * GROUP, w);
        if (fmt >= 1) {                       // per-lane constants instead of per-format code: no divergence between tiles
            float2 X[8];
            float back, M, L;
            uint32_t E;
            fmt_consts(fmt, M, L);
            if (group_prepare_fast(w, X, back, E)) group_recon_fast(X, M, L, back, y);
            else group_recon_slow(w, E, fmt, y);
            stg256(out + row * cols + g * GROUP, y);
        } else {
            stg256(out + row * cols + g * GROUP, w);
        }
    }
}

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

// ---------------------------------------------------------------------------------------------
// Column groups (the `transpose` algorithm, compression_algorithms/transpose.py:13-33): the reference quantizes x.T and
// transposes back, i.e. the shared exponent runs over 16 consecutive ROWS of one column (axis 0 of the original array,
// all later axes flattened into `cols`).  A thread owns one column of one 16-row group: 16 loads that are coalesced
// across the warp (consecutive columns), register maximum, 16 coalesced stores per format - no data ever moves through
// a transposed copy.  Rows beyond `rows` read as +0 like the reference's zero padding.
// ---------------------------------------------------------------------------------------------
template <int DT>
__global__ void __launch_bounds__(256) recon_cols_kernel(const void* __restrict__ x, int64_t rows, int64_t cols, int64_t ld,
                                                         uint32_t fmt_mask, OutPtrs out) {
    const int64_t ngr = cdiv(rows, GROUP);
    const int64_t total = ngr * cols;
    for (int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = it / cols, c = it - g * cols;
        uint32_t u[GROUP];
#pragma unroll
        for (int i = 0; i < GROUP; ++i) {
            const int64_t r = g * GROUP + i;
            uint32_t v = 0;
            if (r < rows) {
                if (DT == QA_DT_BF16) v = (uint32_t)reinterpret_cast<const uint16_t*>(x)[r * ld + c] << 16;
                else v = reinterpret_cast<const uint32_t*>(x)[r * ld + c];
            }
            u[i] = v;
        }
        const uint32_t E = group_max_exp(u);
#pragma unroll
        for (int f = 0; f < QA_NFMT; ++f) {
            if (!((fmt_mask >> f) & 1u)) continue;
            uint16_t* o = reinterpret_cast<uint16_t*>(out.p[f]);
#pragma unroll
            for (int i = 0; i < GROUP; ++i) {
                const int64_t r = g * GROUP + i;
                if (r < rows) o[r * cols + c] = (uint16_t)(recon_bits(f, u[i], E) >> 16);
            }
        }
    }
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
        if (vec) recon_fast_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint16_t*>(x), rows, cols, ld, gpr, fmt_mask, o);
        else recon_kernel<QA_DT_BF16, false><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, fmt_mask, o);
    } else {
        if (vec) recon_kernel<QA_DT_F32, true><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, fmt_mask, o);
        else recon_kernel<QA_DT_F32, false><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, fmt_mask, o);
    }
    return check_launch("qa_quant_recon");
}

extern "C" int qa_quant_recon_cols(const void* x, int x_dtype, int64_t rows, int64_t cols, int64_t ld,
                                   uint32_t fmt_mask, void* const out[QA_NFMT], qa_stream_t stream) {
    if (rows < 0 || cols < 0 || ld < cols) { set_error("qa_quant_recon_cols: bad shape"); return 1; }
    if (x_dtype != QA_DT_BF16 && x_dtype != QA_DT_F32) { set_error("qa_quant_recon_cols: bad dtype"); return 1; }
    fmt_mask &= 0xFu;
    if (rows == 0 || cols == 0 || fmt_mask == 0) return 0;
    OutPtrs o;
    for (int f = 0; f < QA_NFMT; ++f) {
        o.p[f] = (fmt_mask >> f) & 1u ? out[f] : nullptr;
        if (((fmt_mask >> f) & 1u) && !out[f]) { set_error("qa_quant_recon_cols: null output for format %d", f); return 1; }
    }
    const int grid = grid_for(cdiv(rows, GROUP) * cols, 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (x_dtype == QA_DT_BF16) recon_cols_kernel<QA_DT_BF16><<<grid, 256, 0, s>>>(x, rows, cols, ld, fmt_mask, o);
    else recon_cols_kernel<QA_DT_F32><<<grid, 256, 0, s>>>(x, rows, cols, ld, fmt_mask, o);
    return check_launch("qa_quant_recon_cols");
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
        if (vec) apply_fast_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint16_t*>(x), rows, cols, ld, gpr, tiles_w, assignment,
                                                        reinterpret_cast<uint16_t*>(out_bf16));
        else apply_kernel<QA_DT_BF16, false><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, tiles_w, assignment, out_bf16);
    } else if (x_dtype == QA_DT_F32) {
        if (vec) apply_kernel<QA_DT_F32, true><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, tiles_w, assignment, out_bf16);
        else apply_kernel<QA_DT_F32, false><<<grid, 256, 0, s>>>(x, rows, cols, ld, gpr, tiles_w, assignment, out_bf16);
    } else { set_error("qa_apply_assignment: bad dtype"); return 1; }
    return check_launch("qa_apply_assignment");
}

// ---------------------------------------------------------------------------------------------
// mxfp4 / nvfp4 scalar proxies (quantization_formats.py:171-183 with :197-278): every magnitude is treated as the amax
// of a block of identical values, so the result is an elementwise function of |x|:
//   mxfp4: s = float32(|x| / 6.0 in float64), s_q = 2^ceil(log2 s), q = fp4_e2m1(|x| / s_q) * s_q
//   nvfp4: s = |x| / 6 (float32),             s_q = fp8_e4m3(s),   q = fp4_e2m1(|x| / s_q) * s_q   (0 when s_q == 0)
// in the reference's float32 arithmetic (IEEE division, first-argmin level search, round-half-even).  floor / ceil of
// log2 are taken from the exponent field, i.e. exactly; the reference's float32 np.log2 can round an argument within
// ~|k| * 4e-8 of 2^k onto k - never the case for quotients of bf16 values.
// ---------------------------------------------------------------------------------------------
namespace qa {

__device__ __forceinline__ float fp4_e2m1_nearest(float a) {        // a >= 0 (or nan); first argmin of |a - level|
    const float L[8] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, 4.f, 6.f};
    float best = fabsf(__fsub_rn(a, 0.f)), lv = 0.f;
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        const float d = fabsf(__fsub_rn(a, L[i]));
        if (d < best) { best = d; lv = L[i]; }
    }
    return lv;
}

__device__ __forceinline__ float fp8_e4m3_pos(float ax) {            // ax > 0, finite (quantization_formats.py:205-251)
    const uint32_t u = __float_as_uint(ax);
    const int eb = (int)(u >> 23);
    const int e = eb ? eb - 127 : -127;                              // floor(log2 ax); subnormals only need "< -6"
    if (e > 7) return 240.f;                                         // (1 + 7/8) * 2^7
    if (e < -6) return __fmul_rn(rintf(__fmul_rn(ax, 512.f)), 0.001953125f);     // round(ax / 2^-9) * 2^-9
    const float m = __fmul_rn(ax, __uint_as_float((uint32_t)(127 - e) << 23));    // ax / 2^e in [1, 2), exact
    float fq = __fmul_rn(rintf(__fmul_rn(__fsub_rn(m, 1.f), 8.f)), 0.125f);
    int en = e;
    if (fq >= 1.f) { fq = 0.f; en = min(e + 1, 7); }
    return __fmul_rn(__fadd_rn(1.f, fq), __uint_as_float((uint32_t)(127 + en) << 23));
}

template <int DT, int WHICH>
__global__ void __launch_bounds__(256) scalar_proxy_kernel(const void* __restrict__ x, int64_t n, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const float v = DT == QA_DT_BF16 ? __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(x)[i] << 16)
                                         : reinterpret_cast<const float*>(x)[i];
        const float am = fabsf(v);
        float q = 0.f;
        if (!(am <= 3.402823466e38f)) {                              // inf / nan propagate as nan
            out[i] = __uint_as_float(0x7FC00000u);
            continue;
        }
        if (am > 0.f) {
            float sq = 0.f;
            if (WHICH == 0) {
                const float s = __double2float_rn((double)am / 6.0);
                if (s > 0.f) {
                    const uint32_t u = __float_as_uint(s), mant = u & 0x7FFFFFu;
                    const int eb = (int)(u >> 23);
                    int k;                                           // ceil(log2 s)
                    if (eb) k = eb - 127 + (mant ? 1 : 0);
                    else { const int p = 31 - __clz((int)mant); k = -149 + p + ((mant & (mant - 1u)) ? 1 : 0); }
                    sq = k >= -126 ? __uint_as_float((uint32_t)(k + 127) << 23) : __uint_as_float(1u << (k + 149));
                }
            } else {
                const float s = __fdiv_rn(am, 6.f);
                if (s > 0.f) sq = fp8_e4m3_pos(s);
            }
            if (sq > 0.f) q = __fmul_rn(fp4_e2m1_nearest(__fdiv_rn(am, sq)), sq);
        }
        out[i] = v > 0.f ? q : (v < 0.f ? -q : 0.f);
    }
}

}  // namespace qa

extern "C" int qa_scalar_proxy(const void* x, int x_dtype, int64_t n, int which, float* out, qa_stream_t stream) {
    if (n < 0 || (n > 0 && (!x || !out)) || (which != 0 && which != 1)) { set_error("qa_scalar_proxy: bad args"); return 1; }
    if (n == 0) return 0;
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv(n, 256), 148 * 16);
    cudaStream_t s = (cudaStream_t)stream;
    if (x_dtype == QA_DT_BF16) {
        if (which == 0) scalar_proxy_kernel<QA_DT_BF16, 0><<<grid, 256, 0, s>>>(x, n, out);
        else scalar_proxy_kernel<QA_DT_BF16, 1><<<grid, 256, 0, s>>>(x, n, out);
    } else if (x_dtype == QA_DT_F32) {
        if (which == 0) scalar_proxy_kernel<QA_DT_F32, 0><<<grid, 256, 0, s>>>(x, n, out);
        else scalar_proxy_kernel<QA_DT_F32, 1><<<grid, 256, 0, s>>>(x, n, out);
    } else { set_error("qa_scalar_proxy: bad dtype"); return 1; }
    return check_launch("qa_scalar_proxy");
}

// ---------------------------------------------------------------------------------------------
// fp8 e4m3fn + per-block inverse scale -> float32 (hf_model_utils.py:199-215: tensor.float() * inv_scale over
// ceil(shape / scale shape) blocks), the step in front of the path for real checkpoints.  Also emits the bf16 pattern of
// every product and counts the products that are not bf16-exact, so the caller can pick the bf16 or the float32 kernels
// without another pass (qa_f32_to_bf16_checked fused in).
// ---------------------------------------------------------------------------------------------
namespace qa {

__device__ __forceinline__ float e4m3fn_to_f32(uint32_t b) {
    const uint32_t sgn = (b & 0x80u) << 24, e = (b >> 3) & 0xFu, m = b & 7u;
    if (e == 15u && m == 7u) return __uint_as_float(0x7FC00000u);
    if (e == 0u) return __uint_as_float(sgn | __float_as_uint((float)m * 0.001953125f));      // m * 2^-9
    return __uint_as_float(sgn | ((e + 120u) << 23) | (m << 20));
}

__global__ void __launch_bounds__(256) fp8_block_dequant_kernel(const uint8_t* __restrict__ w, const float* __restrict__ scale,
                                                                int64_t rows, int64_t cols, int64_t scols, int64_t br, int64_t bc,
                                                                float* __restrict__ out, uint16_t* __restrict__ out_bf16,
                                                                unsigned long long* __restrict__ inexact) {
    unsigned long long bad = 0;
    const int64_t n = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int64_t r = i / cols, c = i - r * cols;
        float v = __fmul_rn(e4m3fn_to_f32(w[i]), scale[(r / br) * scols + c / bc]);
        if (v != v) v = __uint_as_float(0x7FC00000u);                 // one nan pattern (bf16-exact: not counted)
        if (out) out[i] = v;
        const uint32_t u = __float_as_uint(v);
        if (out_bf16) out_bf16[i] = (uint16_t)(bf16_rne_bits(u) >> 16);
        bad += (u & 0xFFFFu) != 0u;
    }
    if (inexact) {
        bad = __reduce_add_sync(0xFFFFFFFFu, (unsigned)bad);
        if ((threadIdx.x & 31) == 0 && bad) atomicAdd(inexact, bad);
    }
}

}  // namespace qa

extern "C" int qa_fp8_block_dequant(const void* w_fp8, const float* scale_inv, int64_t rows, int64_t cols, int64_t scale_rows,
                                    int64_t scale_cols, float* out_f32, void* out_bf16, unsigned long long* inexact_count,
                                    qa_stream_t stream) {
    if (rows < 0 || cols < 0 || scale_rows <= 0 || scale_cols <= 0 || !w_fp8 || !scale_inv || (!out_f32 && !out_bf16)) {
        set_error("qa_fp8_block_dequant: bad args");
        return 1;
    }
    if (rows == 0 || cols == 0) return 0;
    const int64_t br = std::max<int64_t>(1, cdiv(rows, scale_rows)), bc = std::max<int64_t>(1, cdiv(cols, scale_cols));
    if (cdiv(rows, br) > scale_rows || cdiv(cols, bc) > scale_cols) { set_error("qa_fp8_block_dequant: scale shape too small"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    if (inexact_count && cudaMemsetAsync(inexact_count, 0, sizeof(unsigned long long), s) != cudaSuccess)
        return check_launch("qa_fp8_block_dequant (memset)");
    const unsigned grid = (unsigned)std::min<int64_t>(cdiv(rows * cols, 256), 148 * 32);
    fp8_block_dequant_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(w_fp8), scale_inv, rows, cols, scale_cols, br, bc,
                                                  out_f32, reinterpret_cast<uint16_t*>(out_bf16), inexact_count);
    return check_launch("qa_fp8_block_dequant");
}

extern "C" int qa_f32_to_bf16_checked(const float* x, int64_t n, void* out_bf16,
                                      unsigned long long* inexact_count, qa_stream_t stream) {
    if (n < 0 || !inexact_count) { set_error("qa_f32_to_bf16_checked: bad args"); return 1; }
    if (n == 0) return 0;
    f32_to_bf16_checked_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
        x, n, reinterpret_cast<uint16_t*>(out_bf16), inexact_count);
    return check_launch("qa_f32_to_bf16_checked");
}
