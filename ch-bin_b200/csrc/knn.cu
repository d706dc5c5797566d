// knn.cu -- label-segmented top-k with FP32 keys and an exact FP64 re-rank (default distance mode) (sm_100a).
//
// Replaces find_nearest_from_cluster (/root/reference/ch_bin/core/clustering/distance_matrix.py:47-62) for ALL
// bins of a query in one pass over the query's row.  The reference ranks on scipy's exact FP64 cdist values;
// here the streamed row holds FP32 approximations A of the SQUARED distances with |A - d^2| <= E (approx.cu), and
// exact values (scipy recipe: sequential sum, no FMA, sqrt) are formed only where they can matter:
//
//   pass 1  per bin keep the KR = min(k+3, 32) smallest (A, index) among the bin's current members, streaming the
//           row once (HBM/L2-bound: 4 B per point + packed labels);
//   re-rank every true top-k member j has d_j^2 <= d_(k)^2 <= A_(k) + E, hence A_j <= A_(k) + 2E: it is among the
//           kept entries unless more than KR entries fall inside that slack ("overflow").  For each bin whose kept
//           set changed, the entries with A <= A_(k) + 2E get their exact distance (one thread per entry, cached
//           members keep theirs) and the k smallest exact (distance, index) pairs are selected;
//   pass 2  (rare: duplicate contigs) overflowed bins are re-streamed with the fixed threshold A_(k) + 2E and
//           every passing point is ranked exactly.
//
// Only exact (distance, index) pairs are ever compared when choosing neighbours, so the selected sets are
// bit-identical to ranking the full exact row (ties at the k-th distance go to the lower index).
//
// Mode 0 (assignment rounds): the label the query at permutation position p sees for point i is
//        pos[i] < p ? tent_pt[i] : old_label[i]      (algorithm.py:46-60), the query itself removed;
//   lists are warm-started from the per-(query, bin) cache, so in steady state the scan is a pure threshold
//   filter, and a pair reaches the QP work list only if its neighbour list changed.
// Mode 1 (chb_knn_per_bin): snapshot labels, cold start, lists written out per item.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int NW = NT / 32;
constexpr int EPT = 4;
constexpr int CHUNK = NT * EPT;

struct Smem {
    double *le;   // C*KR exact distance of an entry, NaN = not evaluated
    double *q_d;  // CHUNK (pass 2)
    double *xq;   // d
    float *la;    // C*KR FP32 key
    float *q_a;   // CHUNK
    float *thr;   // C : entries with A <= thr can matter (A_(k) + 2E once k entries are known, +inf before)
    int *li;      // C*KR
    int *q_i, *q_c;
    int *cnt;     // C
    int *flags;   // C : bit0 kept set changed, bit1 overflow
    int *ctl;     // [0] queue length, [1] any overflow
};

__host__ __device__ inline size_t smem_layout(unsigned char *base, int C, int KR, int d, Smem *s)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
    const size_t o_le = take(sizeof(double) * (size_t)C * KR), o_qd = take(sizeof(double) * CHUNK),
                 o_xq = take(sizeof(double) * (size_t)d), o_la = take(sizeof(float) * (size_t)C * KR),
                 o_qa = take(sizeof(float) * CHUNK), o_thr = take(sizeof(float) * (size_t)C),
                 o_li = take(sizeof(int) * (size_t)C * KR), o_qi = take(sizeof(int) * CHUNK),
                 o_qc = take(sizeof(int) * CHUNK), o_cnt = take(sizeof(int) * (size_t)C),
                 o_fl = take(sizeof(int) * (size_t)C), o_ctl = take(sizeof(int) * 4);
    if (s) {
        s->le = reinterpret_cast<double *>(base + o_le);
        s->q_d = reinterpret_cast<double *>(base + o_qd);
        s->xq = reinterpret_cast<double *>(base + o_xq);
        s->la = reinterpret_cast<float *>(base + o_la);
        s->q_a = reinterpret_cast<float *>(base + o_qa);
        s->thr = reinterpret_cast<float *>(base + o_thr);
        s->li = reinterpret_cast<int *>(base + o_li);
        s->q_i = reinterpret_cast<int *>(base + o_qi);
        s->q_c = reinterpret_cast<int *>(base + o_qc);
        s->cnt = reinterpret_cast<int *>(base + o_cnt);
        s->flags = reinterpret_cast<int *>(base + o_fl);
        s->ctl = reinterpret_cast<int *>(base + o_ctl);
    }
    return off;
}

__device__ __forceinline__ double exact_distance(const double *__restrict__ xq_s, const double *__restrict__ xi, int d)
{
    // scipy cdist 'euclidean': s = 0; s += (u[t]-v[t])^2 in ascending t, separate multiply and add; sqrt
    double acc = 0.0;
    int t = 0;
    for (; t + 4 <= d; t += 4) {
        const double v0 = xi[t], v1 = xi[t + 1], v2 = xi[t + 2], v3 = xi[t + 3];
        const double d0 = __dsub_rn(xq_s[t], v0), d1 = __dsub_rn(xq_s[t + 1], v1);
        const double d2 = __dsub_rn(xq_s[t + 2], v2), d3 = __dsub_rn(xq_s[t + 3], v3);
        acc = __dadd_rn(acc, __dmul_rn(d0, d0));
        acc = __dadd_rn(acc, __dmul_rn(d1, d1));
        acc = __dadd_rn(acc, __dmul_rn(d2, d2));
        acc = __dadd_rn(acc, __dmul_rn(d3, d3));
    }
    for (; t < d; ++t) {
        const double d0 = __dsub_rn(xq_s[t], xi[t]);
        acc = __dadd_rn(acc, __dmul_rn(d0, d0));
    }
    return __dsqrt_rn(acc);
}

// pass 1: insert (ca, ci) [exact distance ce, NaN if unknown] into the FP32-keyed list of one bin (KR slots,
// lane l mirrors slot l); all lanes pass the same candidate
__device__ __forceinline__ void insert_key(float *la, int *li, double *le, int *cnt_p, float *thr_p, int *flag_p, int k, int KR,
                                           float slack2, float ca, int ci, double ce, int lane)
{
    const int n = *cnt_p;
    const bool have = lane < n;
    const float ma = have ? la[lane] : 0.f;
    const int mi = have ? li[lane] : -1;
    const double me = have ? le[lane] : 0.0;
    if (__any_sync(CHB_FULL, have && mi == ci)) return;
    const bool less = have && (ma < ca || (ma == ca && mi < ci));
    const int at = __popc(__ballot_sync(CHB_FULL, less));
    const float thr = *thr_p;
    if (at >= KR) {
        if (lane == 0 && ca <= thr) *flag_p |= 2; // a point inside the slack could not be kept
        return;
    }
    const float last = (n == KR) ? la[KR - 1] : FLT_MAX;
    __syncwarp();
    if (have && lane >= at && lane + 1 < KR) {
        la[lane + 1] = ma;
        li[lane + 1] = mi;
        le[lane + 1] = me;
    }
    if (lane == 0) {
        la[at] = ca;
        li[at] = ci;
        le[at] = ce;
        *cnt_p = n + 1 < KR ? n + 1 : KR;
        int f = *flag_p | 1;
        if (n == KR && last <= thr) f |= 2; // the evicted entry was inside the slack
        *flag_p = f;
    }
    __syncwarp();
    if (lane == 0 && *cnt_p >= k) *thr_p = __fadd_ru(la[k - 1], slack2);
    __syncwarp();
}

// pass 2: exact-keyed insert into the first k slots
__device__ __forceinline__ void insert_exact(double *le, int *li, int *cnt_p, int k, double cd, int ci, int lane)
{
    const int n = *cnt_p;
    const bool have = lane < n;
    const double md = have ? le[lane] : 0.0;
    const int mi = have ? li[lane] : -1;
    if (__any_sync(CHB_FULL, have && mi == ci)) return;
    const bool less = have && (md < cd || (md == cd && mi < ci));
    const int at = __popc(__ballot_sync(CHB_FULL, less));
    if (at >= k) return;
    __syncwarp();
    if (have && lane >= at && lane + 1 < k) {
        le[lane + 1] = md;
        li[lane + 1] = mi;
    }
    if (lane == 0) {
        le[at] = cd;
        li[at] = ci;
        *cnt_p = n + 1 < k ? n + 1 : k;
    }
    __syncwarp();
}

__device__ __forceinline__ int eff_label(const chb_knn_args &a, int64_t i, int p)
{
    if (a.mode == 0) {
        const int pi = a.pos[i];
        return (pi < p) ? a.tent_pt[i] : a.old_label[i];
    }
    return a.old_label[i];
}

__global__ void __launch_bounds__(NT) knn_scan_kernel(chb_knn_args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = a.C, k = a.k, d = a.d;
    const int KR = (k + 3 < 32) ? k + 3 : 32;
    Smem s;
    smem_layout(smem_raw, C, KR, d, &s);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t item = blockIdx.x;
    const int64_t n = a.n;

    int p, j;
    int64_t rowi, cache_row;
    if (a.mode == 0) {
        p = a.items[item];
        j = a.perm_pt[p];
        cache_row = (int64_t)a.qslot[j] - a.u0;
        rowi = a.row_is_item ? item : cache_row;
    } else {
        p = 0;
        j = a.items[item];
        cache_row = item;
        rowi = a.row_is_item ? item : ((int64_t)a.qslot[j] - a.u0);
    }
    const float *__restrict__ arow = a.arows + rowi * a.row_stride;
    int32_t *cidx = a.knn_idx + cache_row * (int64_t)C * k;
    int32_t *ccnt = a.knn_cnt + cache_row * (int64_t)C;
    double *cdst = a.knn_dist ? a.knn_dist + cache_row * (int64_t)C * k : nullptr;

    {
        const double *__restrict__ xj = a.X + (int64_t)j * a.ldx;
        for (int t = tid; t < d; t += NT) s.xq[t] = xj[t];
    }
    const double nmax = (double)__uint_as_float(*a.nrm_max_bits);
    const float slack2 = __double2float_ru(2.0 * (a.eps_rel * ((double)a.nrm[j] + nmax) + 1e-30));
    for (int c = tid; c < C; c += NT) {
        s.cnt[c] = 0;
        s.thr[c] = INFINITY;
        s.flags[c] = 0;
    }
    if (tid < 4) s.ctl[tid] = 0;
    __syncthreads();

    // ---- warm start: cached members that still belong to the bin keep their exact distance
    if (a.mode == 0) {
        for (int c = warp; c < C; c += NW) {
            const int m = ccnt[c];
            if (m < 0) { if (lane == 0) s.flags[c] = 1; continue; }
            int ci = -1;
            double ce = 0.0;
            float ca = 0.f;
            bool ok = false;
            if (lane < m) {
                ci = cidx[c * k + lane];
                ok = (eff_label(a, ci, p) == c) && (ci != j);
                if (ok) { ce = cdst[c * k + lane]; ca = arow[ci]; }
            }
            const unsigned keep = __ballot_sync(CHB_FULL, ok);
            const int kept = __popc(keep);
            if (lane == 0) {
                s.cnt[c] = 0;
                if (kept != m) s.flags[c] = 1;
            }
            __syncwarp();
            // FP32 keys are not monotone in the exact order: insert one by one
            unsigned rest = keep;
            while (rest) {
                const int src = __ffs(rest) - 1;
                rest &= rest - 1;
                const float ba = __shfl_sync(CHB_FULL, ca, src);
                const int bi = __shfl_sync(CHB_FULL, ci, src);
                const double be = __shfl_sync(CHB_FULL, ce, src);
                int dummy = 0;
                insert_key(s.la + c * KR, s.li + c * KR, s.le + c * KR, s.cnt + c, s.thr + c, &dummy, k, KR, slack2, ba, bi, be, lane);
            }
        }
        __syncthreads();
    }

    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    // ---- pass 1: stream the FP32 row
    for (int64_t base = 0; base < n; base += CHUNK) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int64_t i = base + (int64_t)e * NT + tid;
            bool pass = false;
            float fv = 0.f;
            int c = -1;
            if (i < n) {
                fv = arow[i];
                c = eff_label(a, i, p);
                pass = (c >= 0) && (c < C) && (i != j) && (fv <= s.thr[c]);
            }
            const unsigned m = __ballot_sync(CHB_FULL, pass);
            if (m) {
                int b = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) b = atomicAdd(&s.ctl[0], __popc(m));
                b = __shfl_sync(CHB_FULL, b, leader);
                if (pass) {
                    const int slot = b + __popc(m & ((1u << lane) - 1u));
                    s.q_a[slot] = fv;
                    s.q_i[slot] = (int)i;
                    s.q_c[slot] = c;
                }
            }
        }
        __syncthreads();
        const int qn = s.ctl[0];
        if (qn > 0) {
            for (int b = 0; b < qn; b += 32) {
                const int e = b + lane;
                float ca = 0.f;
                int ci = 0, cc = -1;
                if (e < qn) { ca = s.q_a[e]; ci = s.q_i[e]; cc = s.q_c[e]; }
                unsigned mine = __ballot_sync(CHB_FULL, (e < qn) && (cc % NW == warp));
                while (mine) {
                    const int src = __ffs(mine) - 1;
                    mine &= mine - 1;
                    const float ba = __shfl_sync(CHB_FULL, ca, src);
                    const int bi = __shfl_sync(CHB_FULL, ci, src);
                    const int bc = __shfl_sync(CHB_FULL, cc, src);
                    if (ba <= s.thr[bc]) // the threshold may have tightened since the element was queued
                        insert_key(s.la + bc * KR, s.li + bc * KR, s.le + bc * KR, s.cnt + bc, s.thr + bc, s.flags + bc, k, KR, slack2,
                                   ba, bi, qnan, lane);
                }
            }
            __syncthreads();
            if (tid == 0) s.ctl[0] = 0;
            __syncthreads();
        }
    }

    // ---- exact re-rank of the bins whose kept set changed (one warp per bin, one lane per kept entry)
    for (int c = warp; c < C; c += NW) {
        const int fl = s.flags[c];
        if (!(fl & 1)) continue;
        if (fl & 2) { if (lane == 0) s.ctl[1] = 1; continue; }
        const int m = s.cnt[c];
        const bool have = lane < m;
        const float ma = have ? s.la[c * KR + lane] : 0.f;
        int mi = have ? s.li[c * KR + lane] : INT32_MAX;
        double me = have ? s.le[c * KR + lane] : 0.0;
        const bool cand = have && (m < k || ma <= s.thr[c]);
        if (cand && !(me == me)) me = exact_distance(s.xq, a.X + (int64_t)mi * a.ldx, d);
        int rank = 0;
        const unsigned candm = __ballot_sync(CHB_FULL, cand);
        for (int t = 0; t < m; ++t) {
            const double oe = __shfl_sync(CHB_FULL, me, t);
            const int oi = __shfl_sync(CHB_FULL, mi, t);
            if (((candm >> t) & 1u) && (oe < me || (oe == me && oi < mi))) ++rank;
        }
        const int ncand = __popc(candm);
        __syncwarp();
        if (cand && rank < k) {
            s.le[c * KR + rank] = me;
            s.li[c * KR + rank] = mi;
        }
        if (lane == 0) s.cnt[c] = ncand < k ? ncand : k;
    }
    __syncthreads();

    // ---- pass 2 (rare): bins with more than KR entries inside the slack are re-streamed and ranked exactly
    if (s.ctl[1]) {
        for (int c = tid; c < C; c += NT)
            if (s.flags[c] & 2) s.cnt[c] = 0; // thr[c] keeps the fixed threshold A_(k) + 2E
        __syncthreads();
        for (int64_t base = 0; base < n; base += CHUNK) {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int64_t i = base + (int64_t)e * NT + tid;
                bool pass = false;
                int c = -1;
                if (i < n) {
                    const float fv = arow[i];
                    c = eff_label(a, i, p);
                    pass = (c >= 0) && (c < C) && (i != j) && (s.flags[c] & 2) && (fv <= s.thr[c]);
                }
                const unsigned m = __ballot_sync(CHB_FULL, pass);
                if (m) {
                    int b = 0;
                    const int leader = __ffs(m) - 1;
                    if (lane == leader) b = atomicAdd(&s.ctl[0], __popc(m));
                    b = __shfl_sync(CHB_FULL, b, leader);
                    if (pass) {
                        const int slot = b + __popc(m & ((1u << lane) - 1u));
                        s.q_i[slot] = (int)i;
                        s.q_c[slot] = c;
                    }
                }
            }
            __syncthreads();
            const int qn = s.ctl[0];
            if (qn > 0) {
                for (int e = tid; e < qn; e += NT) s.q_d[e] = exact_distance(s.xq, a.X + (int64_t)s.q_i[e] * a.ldx, d);
                __syncthreads();
                for (int b = 0; b < qn; b += 32) {
                    const int e = b + lane;
                    double cd = 0.0;
                    int ci = 0, cc = -1;
                    if (e < qn) { cd = s.q_d[e]; ci = s.q_i[e]; cc = s.q_c[e]; }
                    unsigned mine = __ballot_sync(CHB_FULL, (e < qn) && (cc % NW == warp));
                    while (mine) {
                        const int src = __ffs(mine) - 1;
                        mine &= mine - 1;
                        const double bd = __shfl_sync(CHB_FULL, cd, src);
                        const int bi = __shfl_sync(CHB_FULL, ci, src);
                        const int bc = __shfl_sync(CHB_FULL, cc, src);
                        insert_exact(s.le + bc * KR, s.li + bc * KR, s.cnt + bc, k, bd, bi, lane);
                    }
                }
                __syncthreads();
                if (tid == 0) s.ctl[0] = 0;
                __syncthreads();
            }
        }
    }

    // ---- write back: changed lists go to the cache and onto the QP work list
    for (int c = warp; c < C; c += NW) {
        const int m = s.cnt[c] < k ? s.cnt[c] : k;
        if (a.mode == 0) {
            if (!(s.flags[c] & 1)) continue;
            const int mo = ccnt[c];
            bool same = (mo == m);
            if (same) {
                const bool diff = (lane < m) && (cidx[c * k + lane] != s.li[c * KR + lane]);
                same = !__any_sync(CHB_FULL, diff);
            }
            if (same) continue;
            if (lane < k) cidx[c * k + lane] = lane < m ? s.li[c * KR + lane] : -1;
            if (lane < m) cdst[c * k + lane] = s.le[c * KR + lane];
            if (lane == 0) {
                ccnt[c] = m;
                const int w = atomicAdd(a.work_count, 1);
                a.work[w] = make_int2((int)cache_row, c);
            }
        } else {
            if (lane < k) cidx[c * k + lane] = lane < m ? s.li[c * KR + lane] : -1;
            if (lane == 0) ccnt[c] = m;
        }
    }
}

} // namespace

int chb_launch_knn_scan_exact(chb_ctx *ctx, const chb_knn_args &a);

int chb_launch_knn_scan(chb_ctx *ctx, const chb_knn_args &a)
{
    if (a.n_items <= 0) return CHB_OK;
    if (!a.filter) return chb_launch_knn_scan_exact(ctx, a);
    CHB_CHECK(ctx, a.k >= 1 && a.k <= CHB_KMAX, CHB_EINVAL, "num_neighbors must be in [1, %d]", CHB_KMAX);
    const int KR = (a.k + 3 < 32) ? a.k + 3 : 32;
    const size_t bytes = smem_layout(nullptr, a.C, KR, a.d, nullptr) + 16;
    CHB_CHECK(ctx, bytes <= 227 * 1024, CHB_EINVAL, "num_clusters*num_neighbors too large for the kNN kernel (%zu B smem)",
              bytes);
    static size_t configured = 0;
    if (bytes > configured) {
        CHB_CUDA(ctx, cudaFuncSetAttribute(knn_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = bytes;
    }
    {
        chb_stage_timer t(ctx, CHB_ST_KNN);
        knn_scan_kernel<<<(unsigned)a.n_items, NT, bytes, ctx->stream>>>(a);
    }
    CHB_CUDA(ctx, cudaGetLastError());
    ctx->tm.rows_scanned += a.n_items;
    return CHB_OK;
}
