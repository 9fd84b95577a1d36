// csm_se.cuh -- SE90 / SE99, one thread BLOCK per matrix, matrix resident in shared memory as a
// packed lower triangle, thread = permuted position (row).  The block-level sibling of
// wsm_se.cuh for 64 < n <= kThreads (the whole triangle fits 227 KB of shared memory up to
// n = 236; kSpill keeps the first S rows -- the short ones, dead once their position is
// eliminated -- in a per-block global-memory buffer that stays in L1/L2, and addresses every
// row through 64-bit generic addresses):
//   * same storage (element (i,k) at T(i)+k, conflict-free column access), same branch-free
//     symmetric swap, same lower-triangle-only update, same STRICT / FAST arithmetic;
//   * warp collectives become two-level: REDUX inside each warp, one double-buffered exchange
//     line in shared memory and ONE __syncthreads per collective;
//   * n = 128 needs 66 KB instead of the 133 KB of the full-square CTA kernel: three blocks
//     per SM instead of one.
#pragma once
#include <type_traits>

#include "wsm_se.cuh"

namespace modchol {

constexpr int kCsmMaxWarps = 8;

// exchange area: 2 buffers x (8 hi, 8 lo, 8 pos, 8 flags) ints + 2 x 8 doubles + 8 scalars
struct CsmScratch {
    int* ints;        // [2][4][8]
    double* dbl;      // [2][8]
    double* scal;     // [8]
    int buf;
};
__host__ __device__ inline int csm_scratch_doubles() { return 32 + 16 + 8; }

__device__ __forceinline__ CsmScratch csm_scratch(double* base)
{
    CsmScratch s;
    s.ints = reinterpret_cast<int*>(base);
    s.dbl = base + 32;
    s.scal = base + 48;
    s.buf = 0;
    return s;
}

// Block-wide extreme with the NaN rules of warp_extreme; one __syncthreads.
template <bool kMax>
__device__ __forceinline__ double csm_extreme(CsmScratch& sc, int nwarps, double v, bool cand,
                                              double none, bool& win)
{
    int khi;
    unsigned klo;
    make_key(v, khi, klo);
    cand = cand && (v == v);
    if (!kMax) {
        khi = ~khi;
        klo = ~klo;
    }
    const int INT_MIN_ = (int)0x80000000;
    const int wh = __reduce_max_sync(kFull, cand ? khi : INT_MIN_);
    const unsigned wl = __reduce_max_sync(kFull, (cand && khi == wh) ? klo : 0u);
    int* I = sc.ints + sc.buf * 32;
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        I[warp] = wh;
        I[8 + warp] = (int)wl;
    }
    __syncthreads();
    int gh = INT_MIN_;
    unsigned gl = 0u;
    for (int w = 0; w < nwarps; ++w) {
        const int h = I[w];
        const unsigned l = (unsigned)I[8 + w];
        if (h > gh || (h == gh && l > gl)) {
            gh = h;
            gl = l;
        }
    }
    sc.buf ^= 1;
    win = cand && khi == gh && klo == gl;
    if (gh == INT_MIN_)
        return none;
    return kMax ? key_value(gh, gl) : key_value(~gh, ~gl);
}

// First-strict-maximum pivot search over rows [lo, n) (utils.rs:52-91); one __syncthreads.
__device__ __forceinline__ int csm_argmax_first(CsmScratch& sc, int nwarps, double v, bool active,
                                                int row, int lo, double& best)
{
    int khi;
    unsigned klo;
    make_key(v, khi, klo);
    const bool cand = active && (v == v);
    const int INT_MIN_ = (int)0x80000000;
    const int wh = __reduce_max_sync(kFull, cand ? khi : INT_MIN_);
    const bool c1 = cand && khi == wh;
    const unsigned wl = __reduce_max_sync(kFull, c1 ? klo : 0u);
    const unsigned wp = __reduce_min_sync(kFull, (c1 && klo == wl) ? (unsigned)row : 0xffffu);
    const unsigned hn = __ballot_sync(kFull, active && row == lo && (v != v));
    int* I = sc.ints + sc.buf * 32;
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        I[warp] = wh;
        I[8 + warp] = (int)wl;
        I[16 + warp] = (int)wp;
        I[24 + warp] = hn != 0u;
    }
    __syncthreads();
    int gh = INT_MIN_, gp = 0xffff, head_nan = 0;
    unsigned gl = 0u;
    for (int w = 0; w < nwarps; ++w) {
        const int h = I[w], p = I[16 + w];
        const unsigned l = (unsigned)I[8 + w];
        head_nan |= I[24 + w];
        if (h > gh || (h == gh && l > gl)) {
            gh = h;
            gl = l;
            gp = p;
        } else if (h == gh && l == gl && p < gp) {
            gp = p;
        }
    }
    sc.buf ^= 1;
    best = (gh == INT_MIN_) ? -INFINITY : key_value(gh, gl);
    if (head_nan || gh == INT_MIN_ || gp >= 0xffff)
        gp = lo;
    return gp;
}

// block-wide sum (tree order); one __syncthreads
__device__ __forceinline__ double csm_sum_tree(CsmScratch& sc, int nwarps, double x)
{
    x = warp_sum_tree<32>(kFull, x);
    double* D = sc.dbl + sc.buf * 8;
    if ((threadIdx.x & 31) == 0)
        D[threadIdx.x >> 5] = x;
    __syncthreads();
    double s = D[0];
    for (int w = 1; w < nwarps; ++w)
        s = dadd(s, D[w]);
    sc.buf ^= 1;
    return s;
}

// Row addressing of the block-level kernels.  kSpill = false: 32-bit shared byte addresses, the
// whole triangle at wb.  kSpill = true: rows [0, S) live in global memory at gb, rows [S, n) in
// shared memory, all reached through 64-bit generic byte addresses.
template <bool kSpill>
struct CsmRows {
    using addr_t = std::conditional_t<kSpill, unsigned long long, unsigned>;
    addr_t sbase;            // address of the first shared-memory row (row S)
    unsigned long long gb;   // address of row 0 in the spill buffer (kSpill only)
    int S, triS;
    __device__ __forceinline__ addr_t row(int i) const
    {
        if constexpr (kSpill)
            return i < S ? gb + 8ull * (unsigned)tri(i) : sbase + 8ull * (unsigned)(tri(i) - triS);
        else
            return sbase + 8u * (unsigned)tri(i);
    }
    // address of row i + 1 given the address of row i
    __device__ __forceinline__ addr_t next(addr_t a, int i) const
    {
        if constexpr (kSpill)
            return row(i + 1);
        else
            return a + 8u * (unsigned)(i + 1);
    }
};

// FAST mode: blocked trailing update on the FP64 tensor pipe (see wsm_rank4_update_dmma).  Every
// four steps the finished columns [sbase, j) are applied to the stored trailing block as 8x8
// tiles of one DMMA each, so that the per-thread left-looking dot product -- a serial chain of j
// dependent FMAs -- never spans more than three columns.
//   * the four columns are first copied into a compact shared-memory PANEL (row r at 4 r, rows
//     < j and columns >= j - sbase zero, padded to a multiple of eight rows): the A and B
//     fragments of every tile are plain 32-bit shared loads, and because finished rows
//     contribute zeros, a tile may rewrite finished entries with their own value -- only the
//     diagonal tiles (k <= i) and the last block row (i < n) need predicates;
//   * each warp owns the block rows bi = b0 + warp, + nwarps, ...: one row address and one A
//     fragment per block row, then 2 LDS + DMMA + 2 STS (+ 1 LDS for B) per tile.
__host__ __device__ inline int csm_panel_doubles(int n) { return 4 * 8 * ((n + 7) >> 3); }

template <bool kSpill>
__device__ __forceinline__ void csm_rank4_update_dmma(const CsmRows<kSpill>& rows, int n, int j,
                                                      unsigned panelb, int warp, int nwarps, int lane)
{
    using addr_t = typename CsmRows<kSpill>::addr_t;
    const int r8 = lane >> 2, c4 = lane & 3;
    const int b0 = j >> 3, nb = (n + 7) >> 3;
    for (int bi = b0 + warp; bi < nb; bi += nwarps) {
        const int i = 8 * bi + r8;
        const bool iv = i < n;
        const addr_t ra = rows.row(iv ? i : 0);
        const double na = -lds_f64(panelb + 8u * (unsigned)(4 * i + c4));
        const bool full = 8 * bi + 7 < n;
        for (int bj = b0; bj <= bi; ++bj) {
            const double fb = lds_f64(panelb + 8u * (unsigned)(4 * (8 * bj + r8) + c4));
            const unsigned k0 = (unsigned)(8 * bj + 2 * c4);
            double c0, c1, d0, d1;
            if (bj < bi && full) {   // interior tile: no predicates (warp-uniform branch)
                c0 = lds_f64(ra + 8u * k0);
                c1 = lds_f64(ra + 8u * k0 + 8u);
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
                             : "=d"(d0), "=d"(d1) : "d"(na), "d"(fb), "d"(c0), "d"(c1));
                sts_f64(ra + 8u * k0, d0);
                sts_f64(ra + 8u * k0 + 8u, d1);
            } else {
                const bool p0 = iv && (int)k0 <= i, p1 = iv && (int)k0 + 1 <= i;
                c0 = 0.0;
                c1 = 0.0;
                if (p0)
                    c0 = lds_f64(ra + 8u * k0);
                if (p1)
                    c1 = lds_f64(ra + 8u * k0 + 8u);
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%4, %5};"
                             : "=d"(d0), "=d"(d1) : "d"(na), "d"(fb), "d"(c0), "d"(c1));
                sts_f64_if(p0, ra + 8u * k0, d0);
                sts_f64_if(p1, ra + 8u * k0 + 8u, d1);
            }
        }
    }
}

// LEFT-looking evaluation of the pivot column (see wsm_se.cuh): columns >= j keep the
// permuted input until column j is needed,
//     col_i = a_ij - l_i,sbase l_j,sbase - ... - l_i,j-1 l_j,j-1      (in this order),
// the same roundings in the same order as the reference's right-looking updates, with
// read-only shared-memory traffic.  Two __syncthreads (line fill, line release).
// kDiag: the thread at position j evaluates its own diagonal element too (GMW81's c_jj).
template <bool kStrict, bool kDiag, typename addr_t>
__device__ __forceinline__ double csm_left_column(int j, addr_t tjb, int sbase, addr_t tib, int row,
                                                  bool rv)
{
    // row j of L is read in place (contiguous at tjb, block-wide broadcast loads, LDS.128
    // once 16-byte aligned); finished rows skip the dot product.  No barrier is needed
    // here: rows j and the thread's own row were last written before the swap's barrier,
    // and every consumer of the result goes through a block collective first.
    double col = 0.0;
    if (rv && row > (kDiag ? j - 1 : j)) {
        col = lds_f64(tib + 8u * j);
        addr_t rj = tjb + 8u * (unsigned)sbase, ri = tib + 8u * (unsigned)sbase;
        int left = j - sbase;
        if ((rj & 8u) && left > 0) {
            col = nmul_add<kStrict>(col, lds_f64(ri), lds_f64(rj));
            rj += 8;
            ri += 8;
            --left;
        }
        for (; left >= 8; left -= 8) {
            const double2 l01 = lds_v2f64(rj), l23 = lds_v2f64(rj + 16);
            const double2 l45 = lds_v2f64(rj + 32), l67 = lds_v2f64(rj + 48);
            const double lj[8] = {l01.x, l01.y, l23.x, l23.y, l45.x, l45.y, l67.x, l67.y};
            double w[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                w[u] = lds_f64(ri + 8u * u);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                col = nmul_add<kStrict>(col, w[u], lj[u]);
            rj += 64;
            ri += 64;
        }
        for (; left >= 2; left -= 2) {
            const double2 l01 = lds_v2f64(rj);
            const double w0 = lds_f64(ri), w1 = lds_f64(ri + 8);
            col = nmul_add<kStrict>(col, w0, l01.x);
            col = nmul_add<kStrict>(col, w1, l01.y);
            rj += 16;
            ri += 16;
        }
        if (left)
            col = nmul_add<kStrict>(col, lds_f64(ri), lds_f64(rj));
    }
    return col;
}

// FAST: the blocked update of the stored trailing block with columns [sbase, j) (j - sbase = 4):
// every live row copies its four entries into the panel, then the DMMA tiles run.
template <bool kSpill>
__device__ __forceinline__ void csm_blocked_update(const CsmRows<kSpill>& rows, int n, int j, int sbase,
                                                   unsigned panelb,
                                                   typename CsmRows<kSpill>::addr_t tib, int row,
                                                   bool rv, int tid, int nwarps)
{
    const int prow = 8 * ((n + 7) >> 3);
    if (tid < prow) {
        const bool live = rv && row >= j;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            double v = 0.0;
            if (live && sbase + t < j)
                v = lds_f64(tib + 8u * (unsigned)(sbase + t));
            sts_f64(panelb + 8u * (unsigned)(4 * tid + t), v);
        }
    }
    __syncthreads();
    csm_rank4_update_dmma<kSpill>(rows, n, j, panelb, tid >> 5, nwarps, tid & 31);
    __syncthreads();
}

template <int kAlgo, bool kStrict, int kThreads, int kMinBlocks, bool kSpill = false, bool kDmma = false>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
csm_se_kernel(int n, long long batch, const double* __restrict__ A, double* __restrict__ L,
              double* __restrict__ E, unsigned long long* __restrict__ P, int* __restrict__ K,
              double* __restrict__ spill = nullptr, int S = 0)
{
    extern __shared__ double sm[];
    double* W = sm;
    using addr_t = typename CsmRows<kSpill>::addr_t;
    CsmRows<kSpill> rows;
    rows.S = kSpill ? S : 0;
    rows.triS = tri(rows.S);
    const int tpad = (tri(n) - rows.triS + 1) & ~1;
    double* linep = W + tpad;                       // n + 8 entries (scaled pivot column)
    const int lpad = (n + 8 + 1) & ~1;
    CsmScratch sc = csm_scratch(linep + lpad);
    const unsigned wsh = (unsigned)__cvta_generic_to_shared(W);
    // kDmma: the 4-column panel of the blocked update follows the scratch area
    const unsigned panelb = wsh + 8u * (unsigned)(tpad + lpad + csm_scratch_doubles());
    if constexpr (kSpill) {
        rows.sbase = (unsigned long long)(size_t)W;
        rows.gb = (unsigned long long)(size_t)(spill + (size_t)blockIdx.x * rows.triS);
    } else {
        rows.sbase = wsh;
        rows.gb = 0;
    }
    const unsigned lineb = wsh + 8u * (unsigned)tpad;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nwarps = kThreads >> 5;
    const size_t nn = (size_t)n * n;
    const int row = tid;
    const bool rv = row < n;
    const addr_t tib = rows.row(rv ? row : 0);

    for (long long b = blockIdx.x; b < batch; b += gridDim.x) {
        const double* Ab = A + (size_t)b * nn;
        // ---- load the lower triangle (coalesced: consecutive threads, consecutive columns) ----
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += 4) {
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u;
                v[u] = (i < n && tid <= i) ? __ldg(Ab + i * n + tid) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u;
                sts_f64_if(i < n && tid <= i, rows.row(i < n ? i : 0) + 8u * (unsigned)tid, v[u]);
            }
        }
        __syncthreads();

        double dg = rv ? lds_f64(tib + 8u * row) : 0.0;
        double g = 0.0, ev = 0.0;
        int perm = row;
        bool win;
        // gamma = max |a_ii| folded from 0.0 (se90.rs:55-57, se99.rs:54-56)
        const double gamma = csm_extreme<true>(sc, nwarps, fabs(dg), rv, 0.0, win);
        const double taugamma = dmul(MODCHOL_TAU, gamma);
        double delta_prev = 0.0;
        int kstart = -1;

        // symmetric exchange of positions j < m (see wsm_swap_positions) + row scalars
        auto swap_positions = [&](int j, int m, addr_t tjb) {
            const addr_t tmb = rows.row(m);
            const int i = row;
            const bool act = rv && i != m;
            addr_t a1 = (i <= j) ? tjb + 8u * i : tib + 8u * j;
            addr_t a2 = (i == j) ? tmb + 8u * m : ((i < m) ? tmb + 8u * i : tib + 8u * m);
            if (!act) {
                a1 = tib;
                a2 = tib;
            }
            const double t1 = lds_f64(a1), t2 = lds_f64(a2);
            if (i == j || i == m) {
                double* q = sc.scal + (i == j ? 0 : 4);
                q[0] = dg;
                q[1] = g;
                q[2] = (double)perm;
            }
            __syncthreads();
            sts_f64_if(act, a1, t2);
            sts_f64_if(act, a2, t1);
            if (i == j || i == m) {
                const double* q = sc.scal + (i == j ? 4 : 0);
                dg = q[0];
                g = q[1];
                perm = (int)q[2];
            }
            __syncthreads();
        };
        auto scale = [&](double piv, double& sq, double& y, double& inv_piv) -> bool {
            const bool fast = !kStrict && fast_range_ok(piv);
            if (fast) {
                y = rsqrt_normal(piv);
                sq = dmul(piv, y);
                inv_piv = dmul(y, y);
            } else {
                sq = dsqrt(piv);
                y = 0.0;
                inv_piv = 0.0;
            }
            return fast;
        };
        auto left_column = [&](int j, addr_t tjb, int sbase) -> double {
            return csm_left_column<kStrict, false, addr_t>(j, tjb, sbase, tib, row, rv);
        };
        // write column j of L (se90.rs:84-87, se99.rs:96-99); the diagonal update is carried in dg
        auto eliminate = [&](int j, addr_t tjb, double sq, double cs, double dg_new) {
            const bool below = rv && row > j;
            sts_f64_if(below || row == j, below ? tib + 8u * j : tjb + 8u * j, below ? cs : sq);
            if (below)
                dg = dg_new;
            __syncthreads();
        };
        int sbase = 0;
        auto blocked_update = [&](int j) {
            csm_blocked_update<kSpill>(rows, n, j, sbase, panelb, tib, row, rv, tid, nwarps);
        };
        // blocked DMMA trailing update (kDmma, FAST only): measured +3..27 % for n >= 128 (more
        // with mixed phases), -10 % at n = 96 where five blocks per SM already hide the
        // dot-product latency and the extra registers cost one of them -- a separate instantiation
        // value of the row at position j, broadcast to the block (one sync)
        auto bcast_row = [&](double v, int j) -> double {
            double* D = sc.dbl + sc.buf * 8;
            if (row == j)
                D[0] = v;
            __syncthreads();
            sc.buf ^= 1;
            return D[0];
        };

        // ---- phase 1 (se90.rs:61-96, se99.rs:61-109) --------------------------------------
        int j = 0;
        addr_t tjb = rows.row(0);
        bool have_dmin = false;
        double dmin_carry = 0.0;
        for (; j < n; tjb = rows.next(tjb, j), ++j) {
            if constexpr (kDmma && !kStrict) {
                if (j - sbase == 4) {
                    blocked_update(j);
                    sbase = j;
                }
            }
            const bool active = rv && row >= j;
            double dmax;
            const int m = csm_argmax_first(sc, nwarps, dg, active, row, j, dmax);
            if (kAlgo == kSE99) {  // se99.rs:62-73
                // FAST: the previous step's look-ahead minimum is this minimum (see wsm_se.cuh)
                const double dmin = (!kStrict && have_dmin)
                                        ? dmin_carry
                                        : csm_extreme<false>(sc, nwarps, dg, active, INFINITY, win);
                if (dmax < taugamma || dmin < dmul(-MODCHOL_MU, dmax)) {
                    kstart = j;
                    break;
                }
            }
            if (m != j)
                swap_positions(j, m, tjb);
            // FAST: a winner other than the head is a non-NaN maximum and its value is the pivot
            // (saves the broadcast and its barrier); otherwise read the diagonal at position j
            const double piv = (!kStrict && m != j) ? dmax : bcast_row(dg, j);
            double sq, y, inv_piv;
            const bool fast = scale(piv, sq, y, inv_piv);
            const bool below = rv && row > j;
            const double col = left_column(j, tjb, sbase);
            const double cs = fast ? dmul(col, y) : ddiv(col, sq);
            const double dg_new = nmul_add<kStrict>(dg, cs, cs);
            const double nv = fast ? dg_new : dsub(dg, ddiv(dmul(col, col), piv));
            const double la = csm_extreme<false>(sc, nwarps, nv, below, INFINITY, win);
            have_dmin = fast;
            dmin_carry = la;
            const double thresh = (kAlgo == kSE99) ? dmul(-MODCHOL_MU, gamma) : taugamma;
            if (la < thresh) {
                kstart = j;
                break;
            }
            eliminate(j, tjb, sq, cs, dg_new);
        }

        // ---- phase 2 ----------------------------------------------------------------------
        if (kstart >= 0) {
            if (kAlgo == kSE99 && kstart == n - 1) {  // se99.rs:115-119
                const double ljj = bcast_row(dg, n - 1);
                const double ej = tail1x1_e(ljj, gamma);
                if (row == n - 1) {
                    ev = ej;
                    sts_f64(tib + 8u * row, dsqrt(dadd(ljj, ej)));
                }
            } else {
                const int k0 = kstart;
                if (k0 > sbase) {
                    // materialise the trailing block for the Gershgorin bounds: one right-looking
                    // sweep per finished column t, in order (same roundings as the reference)
                    const int kk0 = k0 & ~1;
                    const int c = rv ? row - k0 + 1 : 0;
                    const unsigned cnt = c > 0 ? (unsigned)c : 0u;   // columns k0 .. row
                    const int kend = min(n, (tid | 31) + 1);
                    for (int t = sbase; t < k0; ++t) {
                        const double lt = lds_f64(tib + 8u * t);
                        sts_f64_if(rv, lineb + 8u * row, lt);
                        __syncthreads();
                        addr_t wr = tib + 8u * kk0;
                        unsigned lk = lineb + 8u * kk0;
                        for (int k = kk0; k < kend; k += 4) {
                            const double2 c01 = lds_v2f64(lk);
                            const double2 c23 = lds_v2f64(lk + 16);
                            const double ck[4] = {c01.x, c01.y, c23.x, c23.y};
                            const unsigned off = (unsigned)(k - k0);
                            double w[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                w[u] = lds_f64(wr + 8u * u);
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                sts_f64_if(off + (unsigned)u < cnt, wr + 8u * u,
                                           nmul_add<kStrict>(w[u], lt, ck[u]));
                            wr += 32;
                            lk += 32;
                        }
                        __syncthreads();
                    }
                    sbase = k0;
                }
                // se90.rs:105-110, se99.rs:125-130
                if (rv && row >= k0) {
                    const int i = row;
                    if constexpr (kStrict) {
                        const double s1 =
                            wsm_nd_sum_abs_fn(i - k0, [&](int t) { return lds_f64(tib + 8u * (unsigned)(k0 + t)); });
                        const double s2 = wsm_nd_sum_abs_fn(
                            n - 1 - i, [&](int t) { return lds_f64(rows.row(i + 1 + t) + 8u * (unsigned)i); });
                        g = dsub(dsub(dg, s1), s2);
                    } else {
                        double s1 = 0.0, s2 = 0.0;
                        for (int c = k0; c < i; ++c)
                            s1 = dadd(s1, fabs(lds_f64(tib + 8u * c)));
                        addr_t tr = (i + 1 < n) ? rows.row(i + 1) : tib;
                        for (int r = i + 1; r < n; ++r) {
                            s2 = dadd(s2, fabs(lds_f64(tr + 8u * (unsigned)i)));
                            tr = rows.next(tr, r);
                        }
                        g = dsub(dg, dadd(s1, s2));
                    }
                }
                for (j = k0; j + 2 < n; tjb = rows.next(tjb, j), ++j) {
                    if constexpr (kDmma && !kStrict) {
                        if (j - sbase == 4) {
                            blocked_update(j);
                            sbase = j;
                        }
                    }
                    const bool active = rv && row >= j;
                    double gmax;  // se90.rs:115, se99.rs:135
                    const int m = csm_argmax_first(sc, nwarps, g, active, row, j, gmax);
                    if (m != j)
                        swap_positions(j, m, tjb);
                    double piv = bcast_row(dg, j);
                    const bool below = rv && row > j;
                    const double col = left_column(j, tjb, sbase);
                    const double acol = abs_bits(col);
                    double normj;  // se90.rs:124, se99.rs:144
                    if constexpr (kStrict) {
                        sts_f64_if(rv, lineb + 8u * row, acol);
                        __syncthreads();
                        double* D = sc.dbl + sc.buf * 8;
                        if (tid < 32) {
                            const double s = wsm_nd_sum_line(lineb, j + 1, n - j - 1, lane);
                            if (lane == 0)
                                D[0] = s;
                        }
                        __syncthreads();
                        sc.buf ^= 1;
                        normj = D[0];
                    } else {
                        normj = csm_sum_tree(sc, nwarps, acol);
                    }
                    const double ej =
                        max_b_not_nan(dadd(-piv, max_b_not_nan(normj, taugamma)), delta_prev);
                    if (kStrict) {
                        if (ej > 0.0) {
                            piv = dadd(piv, ej);
                            delta_prev = ej;
                        }
                    } else {
                        piv = dadd(piv, ej);
                        delta_prev = ej;
                    }
                    double sq, y, inv_piv;
                    const bool fast = scale(piv, sq, y, inv_piv);
                    const bool upd = fabs(dsub(piv, normj)) > MODCHOL_EPS;  // se90.rs:134-139
                    double t = 0.0;
                    if (upd)
                        t = fast ? __fma_rn(-normj, inv_piv, 1.0) : dsub(1.0, ddiv(normj, piv));
                    if (row == j)
                        ev = ej;
                    const double cs = fast ? dmul(col, y) : ddiv(col, sq);
                    const double dg_new = nmul_add<kStrict>(dg, cs, cs);
                    if (upd && below)
                        g = mul_add<kStrict>(g, acol, t);
                    eliminate(j, tjb, sq, cs, dg_new);
                }
                // final 2x2 block, never pivoted (se90.rs:155-166, se99.rs:175-186)
                double a00 = bcast_row(dg, n - 2);
                double a11 = bcast_row(dg, n - 1);
                const addr_t a10b = rows.row(n - 1) + 8u * (unsigned)(n - 2);
                const double c2 = left_column(n - 2, rows.row(n - 2), sbase);
                double a10 = bcast_row(c2, n - 1);
                double lhi, llo;
                eigenvalues_2x2(a00, a10, a10, a11, lhi, llo);
                const double en = tail2x2_e(kAlgo, lhi, llo, gamma, delta_prev);
                if (en > 0.0) {
                    a00 = dadd(a00, en);
                    a11 = dadd(a11, en);
                }
                a00 = dsqrt(a00);
                a10 = ddiv(a10, a00);
                a11 = dsqrt(dsub(a11, dmul(a10, a10)));
                __syncthreads();
                if (row == n - 2) {
                    ev = en;
                    sts_f64(tib + 8u * row, a00);
                }
                if (row == n - 1) {
                    ev = en;
                    sts_f64(a10b, a10);
                    sts_f64(tib + 8u * row, a11);
                }
            }
        }
        __syncthreads();

        // ---- epilogue ---------------------------------------------------------------------
        double* Lb = L + (size_t)b * nn;
        for (int i0 = 0; i0 < n; i0 += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u;
                if (i < n && tid < n)
                    Lb[i * n + tid] = (tid <= i) ? lds_f64(rows.row(i) + 8u * (unsigned)tid) : 0.0;
            }
        }
        if (rv) {
            E[(size_t)b * n + perm] = ev;                            // se99.rs:194-198
            P[(size_t)b * n + row] = (unsigned long long)perm;
        }
        if (K && tid == 0)
            K[b] = kstart;
    }
}

// shared memory of one block when rows [0, S) are spilled (S = 0: nothing spilled)
__host__ inline size_t csm_smem_bytes(int n, int S = 0, bool panel = false)
{
    return (size_t)(((tri(n) - tri(S) + 1) & ~1) + ((n + 8 + 1) & ~1) + csm_scratch_doubles() +
                    (panel ? csm_panel_doubles(n) : 0)) *
           sizeof(double);
}

}  // namespace modchol
