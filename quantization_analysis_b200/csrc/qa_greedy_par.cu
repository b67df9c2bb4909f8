// Parallel, bit-faithful greedy assignment: one 1024-thread block per tensor.
//
// The reference's greedy (mixed_tile_greedy.py:135-346) is three sequential chains:
//   (1) sequential float64 accumulation of the per-tile sums in tile order            (:165-170)
//   (2) numpy Generator.permutation of the candidate tiles, once per format           (:228-231)
//   (3) the accept/reject chain: cand = S + (new - old); accept iff metric(cand) passes (:234-346)
// All three are reproduced here with the same results as a one-thread loop, in parallel:
//
//  * Sequentially rounded float64 sums.  While a running sum S stays in one binade its ulp q is
//    fixed and S is an integer multiple m*q; adding t = (a + f)*q (a integer, 0 <= f < 1) gives
//    m + a rounded to nearest - up if f > 1/2, down if f < 1/2, and to the even neighbour if
//    f == 1/2.  The increment therefore depends on the running state only through the parity of
//    m at exact ties, so a chunk of additions composes as pairs (increment if m even, increment
//    if m odd): an associative operator, i.e. a block-wide prefix scan over int64 pairs.  The
//    first element that leaves the binade [2^52, 2^53) q is itself still exact (it is evaluated
//    with a real float64 add from the exact state before it); the chunk is cut after it and the
//    next chunk starts with the new ulp.
//  * Decisions.  Within a chunk the accept flags F are guessed, the exact states before every
//    element follow from the masked scan, every decision D is re-evaluated in parallel with the
//    reference's float64 formula, and F is corrected from the first mismatch on; the fixed
//    point is the sequential result.  (Typical passes are runs of accepts followed by runs of
//    rejects, so this converges in a couple of rounds.)
//  * numpy permutation.  The 32-bit draw stream of PCG64 is generated in parallel by jumping the
//    LCG; the masked-rejection acceptance (draw & mask <= i, i decreasing with every accept) is
//    resolved per 8192-draw round by alternating lower/upper bounds on the accept count until
//    they meet; the Fisher-Yates swap sequence is then applied in parallel by following, for
//    every step, the chain of earlier steps that last wrote the position it reads.
#include "qa_common.cuh"

namespace qa {

constexpr int GT = 1024;          // threads per block
constexpr int NW = GT / 32;
constexpr int DPT = 8;            // draws per thread per round (4 LCG outputs)
constexpr int EPT = 2;            // chain elements per thread per chunk
constexpr int CH = GT * EPT;      // chunk length
constexpr long long M_LO = 1ll << 52, M_HI = 1ll << 53;

struct ParOrder {
    int32_t fmt[QA_NFMT];
    int n;
};

struct ParWork {
    int32_t* order;   // [n]  candidates in visiting order
    int32_t* cand;    // [n]
    int32_t* jarr;    // [n]
    int32_t* off;     // [n+1]
    int32_t* cursor;  // [n]
    int32_t* bucket;  // [n]
    int32_t* succ;    // [n]
    int32_t* parent;  // [n]
    uint8_t* fixed;   // [n]
};

struct Sh {
    int i32[40];
    long long i64[NW * 6 * 2 + 16];
    double f64[16];
    u128 rs;            // LCG state at output index rk
    unsigned long long rk;
    unsigned long long pnext;
    int flag;
};

// ---------------------------------------------------------------------------------------------
// block-wide helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_scan_excl(int v, int& total, Sh& sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) sh.i32[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int x = sh.i32[lane];
        int xi = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xFFFFFFFFu, xi, o);
            if (lane >= o) xi += y;
        }
        sh.i32[lane] = xi - x;
        if (lane == 31) sh.i32[32] = xi;
    }
    __syncthreads();
    const int res = sh.i32[w] + inc - v;
    total = sh.i32[32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ int block_min(int v, Sh& sh) {
    v = __reduce_min_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31) == 0) sh.i32[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = sh.i32[threadIdx.x & 31];
    r = __reduce_min_sync(0xFFFFFFFFu, r);
    __syncthreads();
    return r;
}

__device__ __forceinline__ double block_sum_d(double v, Sh& sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1)
        v += __hiloint2double(__shfl_xor_sync(0xFFFFFFFFu, __double2hiint(v), o),
                              __shfl_xor_sync(0xFFFFFFFFu, __double2loint(v), o));
    if ((threadIdx.x & 31) == 0) sh.f64[0] = 0.0, sh.i64[threadIdx.x >> 5] = __double_as_longlong(v);
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < NW; ++i) r += __longlong_as_double(sh.i64[i]);
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------
// sequentially-rounded float64 accumulation as a scan
// ---------------------------------------------------------------------------------------------
struct P2 {          // increment of m if the incoming m is even / odd
    long long d0, d1;
};
__device__ __forceinline__ P2 p2_then(P2 a, P2 b) {   // apply a, then b
    P2 r;
    r.d0 = a.d0 + ((a.d0 & 1ll) ? b.d1 : b.d0);
    r.d1 = a.d1 + (((1ll + a.d1) & 1ll) ? b.d1 : b.d0);
    return r;
}

struct Grid {        // binade of the running sum at the start of a chunk
    double q;        // ulp (power of two); 0 => no grid (S == 0, inf/nan or denormal range)
    double invq;
    long long m0;    // |S| / q, in [2^52, 2^53)
    double sign;     // +1 / -1
};
__device__ __forceinline__ Grid make_grid(double S) {
    Grid g;
    const int e = (int)((__double2hiint(S) >> 20) & 0x7FF);
    if (e < 64 || e > 1900) { g.q = 0.0; g.invq = 0.0; g.m0 = 0; g.sign = 1.0; return g; }
    g.q = __hiloint2double((e - 52) << 20, 0);
    g.invq = __hiloint2double((1023 + 1023 + 52 - e) << 20, 0);
    g.sign = S < 0.0 ? -1.0 : 1.0;
    g.m0 = __double2ll_rn(fabs(S) * g.invq);
    return g;
}
// classify one addend: returns false when it cannot be expressed on the grid (caller cuts the chunk)
__device__ __forceinline__ bool classify(const Grid& g, double t, P2& out) {
    if (t == 0.0) { out = P2{0, 0}; return true; }      // adding zero never moves the sum
    if (g.q == 0.0) return false;
    const double v = (t * g.sign) * g.invq;             // exact scaling by a power of two
    if (!(fabs(v) < 4.0e18)) return false;
    const double fl = floor(v);
    const long long a = __double2ll_rd(v);
    if (fabs(v) >= 4503599627370496.0) { out.d0 = a; out.d1 = a; return true; }   // already an integer
    const double h = fl + 0.5;                          // exact; the comparisons below are exact too
    if (v < h) { out.d0 = a; out.d1 = a; }
    else if (v > h) { out.d0 = a + 1; out.d1 = a + 1; }
    else { out.d0 = a + (a & 1ll); out.d1 = a + 1 - (a & 1ll); }   // tie: to the even neighbour
    return true;
}
__device__ __forceinline__ long long m_after(const Grid& g, P2 p) { return g.m0 + ((g.m0 & 1ll) ? p.d1 : p.d0); }
__device__ __forceinline__ double s_of(const Grid& g, long long m) { return g.sign * ((double)m * g.q); }

// exclusive block scan of NS P2 streams, EPT elements per thread (element order = thread-major);
// element k of a thread takes part iff bit k of `on` is set.
template <int NS>
__device__ __forceinline__ void scan_p2(const P2 (&cls)[NS][EPT], unsigned on, P2 (&ex)[NS][EPT], Sh& sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    P2 inc[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        P2 run{0, 0};
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
            ex[s][k] = run;
            if ((on >> k) & 1u) run = p2_then(run, cls[s][k]);
        }
        inc[s] = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            P2 y;
            y.d0 = __shfl_up_sync(0xFFFFFFFFu, inc[s].d0, o);
            y.d1 = __shfl_up_sync(0xFFFFFFFFu, inc[s].d1, o);
            if (lane >= o) inc[s] = p2_then(y, inc[s]);
        }
        if (lane == 31) { sh.i64[(w * NS + s) * 2] = inc[s].d0; sh.i64[(w * NS + s) * 2 + 1] = inc[s].d1; }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        P2 pre{0, 0};
        for (int i = 0; i < w; ++i) {                    // totals of the warps before this one
            P2 t;
            t.d0 = sh.i64[(i * NS + s) * 2];
            t.d1 = sh.i64[(i * NS + s) * 2 + 1];
            pre = p2_then(pre, t);
        }
        P2 lanes_before;
        lanes_before.d0 = __shfl_up_sync(0xFFFFFFFFu, inc[s].d0, 1);
        lanes_before.d1 = __shfl_up_sync(0xFFFFFFFFu, inc[s].d1, 1);
        if (lane == 0) lanes_before = P2{0, 0};
        pre = p2_then(pre, lanes_before);
#pragma unroll
        for (int k = 0; k < EPT; ++k) ex[s][k] = p2_then(pre, ex[s][k]);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// metric evaluation (reference formula, float64, no contraction)
// ---------------------------------------------------------------------------------------------
struct Consts {
    double n, sx, sx2, thr;
    int metric;
};
__device__ __forceinline__ double pcc_value_par(const Consts& c, double sy, double sy2, double sxy, double sabs) {
    if (c.n == 0.0) return 1.0;
    const double mx = __ddiv_rn(c.sx, c.n);
    const double my = __ddiv_rn(sy, c.n);
    double am2 = __dsub_rn(c.sx2, __dmul_rn(__dmul_rn(c.n, mx), mx));
    double bm2 = __dsub_rn(sy2, __dmul_rn(__dmul_rn(c.n, my), my));
    if (am2 < 0.0) am2 = 0.0;
    if (bm2 < 0.0) bm2 = 0.0;
    const double den = __dsqrt_rn(__dmul_rn(am2, bm2));
    if (den == 0.0) return sabs == 0.0 ? 1.0 : 0.0;
    return __ddiv_rn(__dsub_rn(sxy, __dmul_rn(__dmul_rn(c.n, mx), my)), den);
}
__device__ __forceinline__ bool good_par(const Consts& c, const double (&S)[4]) {
    if (c.metric == QA_METRIC_PCC) return pcc_value_par(c, S[0], S[1], S[2], S[3]) >= c.thr;
    const double v = c.n != 0.0 ? __ddiv_rn(S[3], c.n) : 0.0;
    return v <= c.thr;
}

// ---------------------------------------------------------------------------------------------
// numpy permutation, parallel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t pcg_out(u128 s) {
    const uint64_t x = s.hi ^ s.lo;
    const uint32_t rot = (uint32_t)(s.hi >> 58);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}
__device__ __forceinline__ u128 pcg_step(u128 s, u128 inc) {
    return add128(mul128(s, u128{QA_PCG_MULT_HI, QA_PCG_MULT_LO}), inc);
}
__device__ void lcg_jump_consts(u128 inc, uint64_t delta, u128& am, u128& ap) {
    am = u128{0, 1};
    ap = u128{0, 0};
    u128 cm{QA_PCG_MULT_HI, QA_PCG_MULT_LO}, cp = inc;
    while (delta) {
        if (delta & 1ull) { am = mul128(am, cm); ap = add128(mul128(ap, cm), cp); }
        cp = mul128(add128(cm, u128{0, 1}), cp);
        cm = mul128(cm, cm);
        delta >>= 1;
    }
}

// count accepts among this thread's DPT draws given the accept count `c` before them
__device__ __forceinline__ int local_accepts(const uint32_t (&v)[DPT], uint32_t validmask, int c, int T, int L) {
    int a = 0;
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const bool ok = ((validmask >> j) & 1u) && (c + a < L) && ((int)v[j] <= T - (c + a));
        a += ok ? 1 : 0;
    }
    return a;
}

// Generates numpy's permutation(m) from the stream in `g` (uniform across the block; every thread
// holds the same copy), writes out[k] = cand ? cand[perm[k]] : perm[k].  apply == false only advances
// the stream (the visiting order is irrelevant when the running state cannot change).
__device__ void permutation_par(Pcg& g, int m, const int32_t* cand, int32_t* out, const ParWork& w, bool apply, Sh& sh) {
    const int tid = threadIdx.x;
    if (m <= 1) {
        if (m == 1 && apply && tid == 0) out[0] = cand ? cand[0] : 0;
        __syncthreads();
        return;
    }
    // per-thread jump constants for 4*tid LCG steps
    u128 am, ap;
    lcg_jump_consts(g.inc, 4ull * (uint64_t)tid, am, ap);
    if (tid == 0) {
        sh.rs = g.s;                 // state at output index rk (out_rk = pcg_out(rs)); out_0.hi == buffered half
        sh.rk = 0ull;
        sh.pnext = g.has32 ? 1ull : 2ull;
    }
    __syncthreads();
    int i_cur = m - 1;
    const int SEQ_TAIL = 96;
    while (i_cur > SEQ_TAIL) {
        const uint32_t mask = 0xFFFFFFFFu >> __clz((uint32_t)i_cur);
        const int seg_lo = (int)(mask >> 1) + 1;
        const int L = i_cur - max(seg_lo, SEQ_TAIL + 1) + 1;      // accepts wanted in this segment
        // advance the shared base to output index kb = pnext >> 1
        const unsigned long long pnext = sh.pnext;
        const unsigned long long kb = pnext >> 1;
        if (tid == 0 && kb != sh.rk) {
            u128 a2, p2;
            lcg_jump_consts(g.inc, kb - sh.rk, a2, p2);
            sh.rs = add128(mul128(a2, sh.rs), p2);
            sh.rk = kb;
        }
        __syncthreads();
        // this thread's 4 outputs: indices kb + 4 tid + j
        u128 s = add128(mul128(am, sh.rs), ap);
        uint32_t v[DPT];
        uint32_t validmask = 0;
#pragma unroll
        for (int j = 0; j < DPT / 2; ++j) {
            const uint64_t o = pcg_out(s);
            v[2 * j] = (uint32_t)o & mask;
            v[2 * j + 1] = (uint32_t)(o >> 32) & mask;
            s = pcg_step(s, g.inc);
        }
        const unsigned long long p0 = 2ull * (kb + 4ull * (unsigned long long)tid);
#pragma unroll
        for (int j = 0; j < DPT; ++j)
            if (p0 + j >= pnext) validmask |= 1u << j;
        // resolve the accept counts by alternating bounds
        int c_lo = 0, a_hi = 0, a_lo = 0, total = 0;
        for (int it = 0; it < 64; ++it) {
            a_hi = local_accepts(v, validmask, c_lo, i_cur, L);
            const int c_hi = block_scan_excl(a_hi, total, sh);
            a_lo = local_accepts(v, validmask, c_hi, i_cur, L);
            c_lo = block_scan_excl(a_lo, total, sh);
            const int a_chk = local_accepts(v, validmask, c_lo, i_cur, L);
            const int diff = __syncthreads_or(a_chk != a_lo);
            if (!diff) break;
        }
        // total = accepts in this round with exact c_lo per thread
        {
            int c = c_lo;
            unsigned long long last_pos = ~0ull;
#pragma unroll
            for (int j = 0; j < DPT; ++j) {
                const bool ok = ((validmask >> j) & 1u) && (c < L) && ((int)v[j] <= i_cur - c);
                if (ok) {
                    w.jarr[i_cur - c] = (int32_t)v[j];
                    ++c;
                    if (c == L) last_pos = p0 + j;      // the draw that completes the segment
                }
            }
            if (last_pos != ~0ull) sh.pnext = last_pos + 1ull;
        }
        __syncthreads();
        if (total < L) {            // segment not finished: every draw of the round was consumed
            if (tid == 0) sh.pnext = 2ull * (kb + 4ull * GT);
        }
        i_cur -= total;
        __syncthreads();
    }
    // sequential tail (and stream hand-back) on thread 0
    if (tid == 0) {
        const unsigned long long pnext = sh.pnext;
        const unsigned long long klast = (pnext - 1ull) >> 1;
        u128 a2, p2;
        lcg_jump_consts(g.inc, klast - sh.rk, a2, p2);
        const u128 s = add128(mul128(a2, sh.rs), p2);
        Pcg t;
        t.inc = g.inc;
        t.s = s;
        t.has32 = ((pnext - 1ull) & 1ull) == 0ull ? 1u : 0u;
        t.buf32 = (uint32_t)(pcg_out(s) >> 32);
        if (pnext == 2ull && !g.has32) { t.s = g.s; t.has32 = 0; t.buf32 = g.buf32; }   // nothing consumed yet
        if (pnext == 1ull) { t.s = g.s; t.has32 = 1; t.buf32 = g.buf32; }
        for (int i = i_cur; i >= 1; --i) w.jarr[i] = (int32_t)t.interval((uint32_t)i);
        sh.rs = t.s;
        sh.i32[34] = (int)t.has32;
        sh.i32[35] = (int)t.buf32;
    }
    __syncthreads();
    g.s = sh.rs;
    g.has32 = (uint32_t)sh.i32[34];
    g.buf32 = (uint32_t)sh.i32[35];
    __syncthreads();
    if (!apply) return;

    // ---- apply the swap sequence j[m-1..1] in parallel --------------------------------------
    for (int p = tid; p < m; p += GT) w.cursor[p] = 0;
    __syncthreads();
    for (int i = 1 + tid; i < m; i += GT) atomicAdd(&w.cursor[w.jarr[i]], 1);
    __syncthreads();
    {   // exclusive scan of the per-position counts -> off[], contiguous range per thread
        const int per = (m + GT - 1) / GT;
        const int b = min(m, tid * per), e = min(m, b + per);
        int local = 0;
        for (int p = b; p < e; ++p) local += w.cursor[p];
        int total;
        int run = block_scan_excl(local, total, sh);
        for (int p = b; p < e; ++p) {
            const int c = w.cursor[p];
            w.off[p] = run;
            w.cursor[p] = run;
            run += c;
        }
        if (tid == 0) w.off[m] = total;
    }
    __syncthreads();
    for (int i = 1 + tid; i < m; i += GT) {
        const int slot = atomicAdd(&w.cursor[w.jarr[i]], 1);
        w.bucket[slot] = i;
    }
    __syncthreads();
    for (int p = tid; p < m; p += GT) {
        const int b = w.off[p], e = w.off[p + 1];
        for (int a = b + 1; a < e; ++a) {          // insertion sort (buckets hold ~1 entry)
            const int key = w.bucket[a];
            int c = a - 1;
            while (c >= b && w.bucket[c] > key) { w.bucket[c + 1] = w.bucket[c]; --c; }
            w.bucket[c + 1] = key;
        }
        int par = -1;
        for (int a = b; a < e; ++a) {
            const int st = w.bucket[a];
            w.succ[st] = a + 1 < e ? w.bucket[a + 1] : -1;
            if (par < 0 && st != p) par = st;
        }
        w.parent[p] = par;       // first later step that writes position p
    }
    __syncthreads();
    for (int i = tid; i < m; i += GT) {
        // a[0] ends as the content of position 0 after all steps; a[i] (i >= 1) is what step i read
        const int start = i == 0 ? w.parent[0] : w.succ[i];
        int val;
        if (start < 0) val = i == 0 ? 0 : w.jarr[i];          // nobody wrote that position: initial content
        else {
            int cur = start;
            for (int pa = w.parent[cur]; pa >= 0; pa = w.parent[cur]) cur = pa;
            val = cur;
        }
        out[i] = cand ? cand[val] : val;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(GT) permutation_par_kernel(qa_pcg64* rng, int m, int32_t* out, ParWork w) {
    __shared__ Sh sh;
    Pcg g;
    g.load(rng);
    permutation_par(g, m, nullptr, out, w, true, sh);
    if (threadIdx.x == 0) g.store(rng);
}

// ---------------------------------------------------------------------------------------------
// the greedy kernel
// ---------------------------------------------------------------------------------------------
// Faithful sequential sum of table column(s) in tile order: S_k = fl(S_{k-1} + t_k).
// The first HEAD elements (where a sum starting from zero changes binade at almost every step)
// are added by one thread; the rest rides the scan.  A column whose running sum keeps leaving its
// binade (a zero-mean random walk around a power of two) is finished with a plain tree sum after
// MAX_EVENTS cuts and reported in `degraded` (bit per column).
template <int NC>
__device__ void faithful_init_sums(const double* const (&col)[NC], int nt, double (&S)[NC], unsigned& degraded,
                                   int max_events, Sh& sh) {
    const int tid = threadIdx.x;
    constexpr int HEAD = 192;
    int events[NC];
#pragma unroll
    for (int s = 0; s < NC; ++s) { S[s] = 0.0; events[s] = 0; }
    degraded = 0;
    int pos = min(nt, HEAD);
    if (tid < NC) {
        double acc = 0.0;
        const double* cp = col[0];
#pragma unroll
        for (int s = 0; s < NC; ++s) if (s == tid) cp = col[s];
        for (int i = 0; i < pos; ++i) acc = __dadd_rn(acc, cp[i]);
        sh.f64[tid] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NC; ++s) S[s] = sh.f64[s];
    __syncthreads();
    while (pos < nt) {
        const int len = min(CH, nt - pos);
        Grid g[NC];
        P2 cls[NC][EPT], ex[NC][EPT];
        double t[NC][EPT];
        int bad = len;                                  // first element that cannot ride the grid
#pragma unroll
        for (int s = 0; s < NC; ++s) {
            g[s] = make_grid(S[s]);
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const int idx = tid * EPT + k;
                t[s][k] = idx < len ? col[s][pos + idx] : 0.0;
                cls[s][k] = P2{0, 0};
                if (idx < len && !((degraded >> s) & 1u)) {
                    if (!classify(g[s], t[s][k], cls[s][k])) bad = min(bad, idx);
                }
            }
        }
        bad = block_min(bad, sh);
        unsigned on = 0;                                // elements before `bad` ride the scan
#pragma unroll
        for (int k = 0; k < EPT; ++k)
            if (tid * EPT + k < bad && tid * EPT + k < len) on |= 1u << k;
        scan_p2<NC>(cls, on, ex, sh);
        int cut = min(bad, len - 1);                    // last element handled this round
#pragma unroll
        for (int s = 0; s < NC; ++s) {
            if ((degraded >> s) & 1u) continue;
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                if ((on >> k) & 1u) {
                    const long long mi = m_after(g[s], p2_then(ex[s][k], cls[s][k]));
                    if (g[s].q != 0.0 && (mi < M_LO || mi >= M_HI)) cut = min(cut, tid * EPT + k);
                }
            }
        }
        cut = block_min(cut, sh);
        // state after element `cut` = fl(S_before(cut) + t_cut): a real add from the exact state before it
        if (cut / EPT == tid) {
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                if (k != cut % EPT) continue;
#pragma unroll
                for (int s = 0; s < NC; ++s) {
                    const double before = g[s].q != 0.0 ? s_of(g[s], m_after(g[s], ex[s][k])) : S[s];
                    sh.f64[s] = __dadd_rn(before, t[s][k]);
                }
            }
        }
        __syncthreads();
        const bool cut_short = cut < len - 1;
#pragma unroll
        for (int s = 0; s < NC; ++s) {
            if ((degraded >> s) & 1u) continue;
            const double nv = sh.f64[s];
            if (cut_short) {
                const Grid gn = make_grid(nv);
                if (gn.q != g[s].q || g[s].q == 0.0) ++events[s];      // this column left its grid here
            }
            S[s] = nv;
        }
        __syncthreads();
        pos += cut + 1;
#pragma unroll
        for (int s = 0; s < NC; ++s) {
            if (!((degraded >> s) & 1u) && events[s] > max_events) {
                double part = 0.0;
                for (int i = pos + tid; i < nt; i += GT) part += col[s][i];
                S[s] = S[s] + block_sum_d(part, sh);
                degraded |= 1u << s;
            }
        }
        bool all_deg = true;
#pragma unroll
        for (int s = 0; s < NC; ++s) all_deg = all_deg && ((degraded >> s) & 1u);
        if (all_deg) break;
    }
}

__global__ void __launch_bounds__(GT) greedy_par_kernel(const double* __restrict__ table, int nt, double numel, int metric,
                                                        double thr, ParOrder ord, qa_pcg64* rng, int8_t* assignment,
                                                        int64_t* counts, double* state, ParWork w) {
    __shared__ Sh sh;
    __shared__ int cnt_sh[QA_NFMT];
    const int tid = threadIdx.x;
    const int base = ord.fmt[0];
    const bool is_pcc = metric == QA_METRIC_PCC;
    for (int t = tid; t < nt; t += GT) { assignment[t] = (int8_t)base; w.fixed[t] = 0; }
    if (tid < QA_NFMT) cnt_sh[tid] = tid == base ? nt : 0;
    __syncthreads();

    const long long t_start = clock64();
    // ---- (1) initial sums, sequentially rounded in tile order ------------------------------
    Consts c;
    c.n = numel; c.thr = thr; c.metric = metric; c.sx = 0.0; c.sx2 = 0.0;
    double S[4] = {0.0, 0.0, 0.0, 0.0};      // sy, sy2, sxy, sabs
    unsigned degraded = 0;
    if (is_pcc) {
        {   // sums of non-negative terms: few binade changes, always carried faithfully
            const double* const cols[4] = {table + (size_t)QA_STAT_SX2 * nt, table + (size_t)QA_STAT_FMT(base, 1) * nt,
                                           table + (size_t)QA_STAT_FMT(base, 2) * nt, table + (size_t)QA_STAT_FMT(base, 3) * nt};
            double R[4];
            unsigned dg;
            faithful_init_sums<4>(cols, nt, R, dg, 1 << 30, sh);
            c.sx2 = R[0]; S[1] = R[1]; S[2] = R[2]; S[3] = R[3];
        }
        {   // signed sums (means): faithful unless they keep hopping between binades
            const double* const cols[2] = {table + (size_t)QA_STAT_SX * nt, table + (size_t)QA_STAT_FMT(base, 0) * nt};
            double R[2];
            faithful_init_sums<2>(cols, nt, R, degraded, 24, sh);
            c.sx = R[0]; S[0] = R[1];
        }
    } else {
        const double* const cols[1] = {table + (size_t)QA_STAT_FMT(base, 3) * nt};
        double R[1];
        unsigned dg;
        faithful_init_sums<1>(cols, nt, R, dg, 1 << 30, sh);
        S[3] = R[0];
    }
    Pcg g;
    g.load(rng);
    unsigned chain_rounds = 0;
    long long t_mark = clock64(), cyc_init = 0, cyc_perm = 0, cyc_chain = 0;
    cyc_init = t_mark - t_start;

    for (int fi = 0; fi < ord.n; ++fi) {
        const int fmt = ord.fmt[fi];
        // ---- candidates = not-fixed tiles in ascending order ------------------------------
        int m;
        {
            const int per = (nt + GT - 1) / GT;
            const int b = min(nt, tid * per), e = min(nt, b + per);
            int local = 0;
            for (int t = b; t < e; ++t) local += w.fixed[t] ? 0 : 1;
            int run = block_scan_excl(local, m, sh);
            for (int t = b; t < e; ++t)
                if (!w.fixed[t]) w.cand[run++] = t;
        }
        __syncthreads();
        if (m == 0) break;
        const bool base_pass = (fmt == base) && (fi == 0);
        // ---- (2) visiting order ----------------------------------------------------------
        t_mark = clock64();
        permutation_par(g, m, w.cand, w.order, w, !base_pass, sh);
        cyc_perm += clock64() - t_mark;
        t_mark = clock64();
        if (base_pass) {
            // every candidate already has this format: the state cannot change, so all of them
            // see the same test (mixed_tile_greedy.py:238-241)
            if (!good_par(c, S)) {
                for (int k = tid; k < m; k += GT) w.fixed[w.cand[k]] = 1;
            }
            __syncthreads();
            continue;
        }
        const double* fcol[4] = {table + (size_t)QA_STAT_FMT(fmt, 0) * nt, table + (size_t)QA_STAT_FMT(fmt, 1) * nt,
                                 table + (size_t)QA_STAT_FMT(fmt, 2) * nt, table + (size_t)QA_STAT_FMT(fmt, 3) * nt};
        const int s_lo = is_pcc ? 0 : 3;
        // ---- (3) accept / reject chain -------------------------------------------------------
        int pos = 0;
        bool guess = true;                      // initial guess for a chunk: accept everything
        while (pos < m) {
            const int len = min(CH, m - pos);
            int tile[EPT], prev[EPT];
            double d[4][EPT];
            P2 cls[4][EPT], ex[4][EPT];
            bool F[EPT], D[EPT], same[EPT];
            Grid gr[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) gr[s] = make_grid(S[s]);
            int bad = len;
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const int idx = tid * EPT + k;
                tile[k] = -1; prev[k] = 0; same[k] = false; F[k] = false; D[k] = false;
#pragma unroll
                for (int s = 0; s < 4; ++s) { d[s][k] = 0.0; cls[s][k] = P2{0, 0}; ex[s][k] = P2{0, 0}; }
                if (idx < len) {
                    const int t = w.order[pos + idx];
                    tile[k] = t;
                    prev[k] = assignment[t];
                    same[k] = prev[k] == fmt;
                    if (!same[k]) {
                        F[k] = guess;
                        for (int s = s_lo; s < 4; ++s) {
                            d[s][k] = __dsub_rn(fcol[s][t], table[(size_t)QA_STAT_FMT(prev[k], s) * nt + t]);
                            if (!classify(gr[s], d[s][k], cls[s][k])) bad = min(bad, idx);
                        }
                    }
                }
            }
            bad = block_min(bad, sh);            // the element at `bad` is added for real and ends the chunk
            int valid = min(len, bad + 1);
            for (int round = 0; round < 64; ++round) {
                ++chain_rounds;
                unsigned on = 0;
#pragma unroll
                for (int k = 0; k < EPT; ++k)
                    if (F[k] && (tid * EPT + k) < valid && (tid * EPT + k) != bad) on |= 1u << k;
                scan_p2<4>(cls, on, ex, sh);
                int cut = valid - 1;             // last element that may be committed this round
                int mism = 1 << 30;
#pragma unroll
                for (int k = 0; k < EPT; ++k) {
                    const int idx = tid * EPT + k;
                    D[k] = false;
                    if (idx < valid) {
                        double cnd[4];
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            const double sb = gr[s].q != 0.0 ? s_of(gr[s], m_after(gr[s], ex[s][k])) : S[s];
                            cnd[s] = same[k] ? sb : __dadd_rn(sb, d[s][k]);
                        }
                        D[k] = good_par(c, cnd);
                        if (!same[k]) {
                            if (D[k] != F[k]) mism = min(mism, idx);
                            if ((on >> k) & 1u) {
                                // an accepted element must leave every running sum inside its binade
                                for (int s = s_lo; s < 4; ++s) {
                                    const long long mi = m_after(gr[s], p2_then(ex[s][k], cls[s][k]));
                                    if (gr[s].q != 0.0 && (mi < M_LO || mi >= M_HI)) cut = min(cut, idx);
                                }
                            }
                        }
                    }
                }
                mism = block_min(mism, sh);
                cut = block_min(cut, sh);
                if (mism > cut) { valid = cut + 1; break; }       // flags are consistent up to the cut
                if (round == 63) { valid = mism + 1; }            // pathological: commit up to the first wrong flag
                // fix the first wrong flag; later ones take the freshly computed decisions as the new guess
#pragma unroll
                for (int k = 0; k < EPT; ++k) {
                    const int idx = tid * EPT + k;
                    if (idx >= mism && idx < valid && !same[k]) F[k] = D[k];
                }
            }
            // ---- commit [0, valid): D holds the decisions, ex the exact states before each element ----
            int loc[QA_NFMT] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < EPT; ++k) {
                const int idx = tid * EPT + k;
                if (idx < valid) {
                    const bool take = !same[k] && D[k];
                    if (same[k]) { if (!D[k]) w.fixed[tile[k]] = 1; }
                    else if (take) {
                        assignment[tile[k]] = (int8_t)fmt;
#pragma unroll
                        for (int f = 0; f < QA_NFMT; ++f) loc[f] += (f == fmt) - (f == prev[k]);
                    } else w.fixed[tile[k]] = 1;
                    if (idx == valid - 1) {
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            const double sb = gr[s].q != 0.0 ? s_of(gr[s], m_after(gr[s], ex[s][k])) : S[s];
                            sh.f64[s] = take ? __dadd_rn(sb, d[s][k]) : sb;
                        }
                        sh.flag = take ? 1 : 0;
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < QA_NFMT; ++f) {
                const int v = __reduce_add_sync(0xFFFFFFFFu, loc[f]);
                if ((tid & 31) == 0 && v) atomicAdd(&cnt_sh[f], v);
            }
            __syncthreads();
#pragma unroll
            for (int s = 0; s < 4; ++s) S[s] = sh.f64[s];
            guess = sh.flag != 0;
            pos += valid;
            __syncthreads();
        }
        cyc_chain += clock64() - t_mark;
    }
    if (tid == 0) {
        g.store(rng);
        for (int f = 0; f < QA_NFMT; ++f) counts[f] = cnt_sh[f];
        state[0] = c.sx; state[1] = c.sx2; state[2] = S[0]; state[3] = S[1]; state[4] = S[2]; state[5] = S[3];
        state[6] = (double)degraded + 65536.0 * (double)chain_rounds;
        state[7] = is_pcc ? pcc_value_par(c, S[0], S[1], S[2], S[3]) : (numel != 0.0 ? __ddiv_rn(S[3], numel) : 0.0);
        state[8] = (double)cyc_init; state[9] = (double)cyc_perm; state[10] = (double)cyc_chain;   // SM cycles per phase
    }
}

static inline int64_t al(int64_t v) { return (v + 255) / 256 * 256; }

}  // namespace qa

using namespace qa;

extern "C" int64_t qa_greedy_par_work_bytes(int64_t n) { return al(4 * (n + 1)) * 8 + al(n) + 512; }

static ParWork carve(void* work, int64_t n) {
    ParWork w;
    char* p = reinterpret_cast<char*>(work);
    auto take = [&](int64_t bytes) { char* r = p; p += al(bytes); return r; };
    w.order = reinterpret_cast<int32_t*>(take(4 * (n + 1)));
    w.cand = reinterpret_cast<int32_t*>(take(4 * (n + 1)));
    w.jarr = reinterpret_cast<int32_t*>(take(4 * (n + 1)));
    w.off = reinterpret_cast<int32_t*>(take(4 * (n + 1)));
    w.cursor = reinterpret_cast<int32_t*>(take(4 * (n + 1)));
    w.bucket = reinterpret_cast<int32_t*>(take(4 * (n + 1)));
    w.succ = reinterpret_cast<int32_t*>(take(4 * (n + 1)));
    w.parent = reinterpret_cast<int32_t*>(take(4 * (n + 1)));
    w.fixed = reinterpret_cast<uint8_t*>(take(n));
    return w;
}

extern "C" int qa_numpy_permutation_par(qa_pcg64* rng, int64_t n, int32_t* out_perm, void* work, qa_stream_t stream) {
    if (!rng || n < 0 || n > 0x3FFFFFFF || (n > 0 && (!out_perm || !work))) { set_error("qa_numpy_permutation_par: bad args"); return 1; }
    if (n == 0) return 0;
    permutation_par_kernel<<<1, GT, 0, (cudaStream_t)stream>>>(rng, (int)n, out_perm, carve(work, n));
    return check_launch("qa_numpy_permutation_par");
}

extern "C" int qa_greedy_assign_par(const double* table, int64_t ntiles, double numel, int metric, double threshold,
                                    const int32_t* fmt_order, int nfmt, qa_pcg64* rng, int8_t* assignment,
                                    int64_t* counts, double* state, void* work, qa_stream_t stream) {
    if (!table || ntiles <= 0 || ntiles > 0x3FFFFFFF || !fmt_order || nfmt < 1 || nfmt > QA_NFMT || !rng || !assignment ||
        !counts || !state || !work) {
        set_error("qa_greedy_assign_par: bad args");
        return 1;
    }
    if (metric != QA_METRIC_PCC && metric != QA_METRIC_MAE) { set_error("qa_greedy_assign_par: metric must be pcc or mae"); return 1; }
    ParOrder ord;
    ord.n = nfmt;
    for (int i = 0; i < QA_NFMT; ++i) ord.fmt[i] = i < nfmt ? fmt_order[i] : 0;
    for (int i = 0; i < nfmt; ++i)
        if (ord.fmt[i] < 0 || ord.fmt[i] >= QA_NFMT) { set_error("qa_greedy_assign_par: bad format index"); return 1; }
    greedy_par_kernel<<<1, GT, 0, (cudaStream_t)stream>>>(table, (int)ntiles, numel, metric, threshold, ord, rng, assignment,
                                                          counts, state, carve(work, ntiles));
    return check_launch("qa_greedy_assign_par");
}
