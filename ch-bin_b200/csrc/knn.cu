// knn.cu -- label-segmented top-k with FP32 keys and an exact FP64 re-rank (default distance mode) (sm_100a).
//
// Replaces find_nearest_from_cluster (/root/reference/ch_bin/core/clustering/distance_matrix.py:47-62) for ALL
// bins of a query in one pass over the query's row.  The reference ranks on scipy's exact FP64 cdist values;
// here the streamed row holds FP32 approximations A of the SQUARED distances with |A - d^2| <= E (approx.cu), and
// exact values (scipy recipe: sequential sum, no FMA, sqrt) are formed only where they can matter:
//
//   pass 1  per bin keep the KR (= k+3, at most 32) smallest (A, index) keys among the bin's current members.
//           The row is streamed once, 16 B of keys + 32 B of packed labels per thread and step.  A point whose
//           key is within the bin's threshold is appended to a small per-bin bucket (one shared-memory atomic);
//           a warp "compacts" a bucket into the sorted list (rank by counting) only when it is half full, which
//           also tightens the threshold -- lazy selection instead of one-at-a-time insertion;
//   re-rank every true top-k member j has d_j^2 <= d_(k)^2 <= A_(k) + E, hence A_j <= A_(k) + 2E: it is among the
//           kept keys unless more than KR keys fall inside that slack ("overflow").  For each bin whose kept set
//           changed, the keys with A <= A_(k) + 2E that have no cached exact distance are evaluated exactly, one
//           THREAD per key CTA-wide, and the k smallest exact (distance, index) pairs are selected;
//   pass 2  (rare: duplicate contigs) overflowed bins are re-streamed with the fixed threshold A_(k) + 2E and
//           every passing point is ranked exactly.
//
// Only exact (distance, index) pairs are ever compared when choosing neighbours, so the selected sets are
// bit-identical to ranking the full exact row (ties at the k-th distance go to the lower index).
//
// Mode 0 (assignment rounds): the label the query at permutation position p sees for point i is
//        pos[i] < p ? tent[i] : old[i]      (algorithm.py:46-60), the query itself removed;
//   lists are warm-started from the per-(query, bin) cache, so in steady state the scan is a pure threshold
//   filter, and a pair reaches the QP work list only if its neighbour list changed.
// Mode 1 (chb_knn_per_bin): snapshot labels, cold start, lists written out per item.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int NW = NT / 32;
constexpr int EPT = 4;            // consecutive row elements per thread and step (one float4)
constexpr int CHUNK = NT * EPT;
typedef unsigned long long u64;
constexpr u64 KEY_MAX = ~0ull;

__device__ __forceinline__ unsigned sortable(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unsortable(unsigned s)
{
    return (s & 0x80000000u) ? __uint_as_float(s & 0x7fffffffu) : __uint_as_float(~s);
}
__device__ __forceinline__ u64 make_key(float a, int idx) { return ((u64)sortable(a) << 32) | (unsigned)idx; }
__device__ __forceinline__ float key_a(u64 k) { return unsortable((unsigned)(k >> 32)); }
__device__ __forceinline__ int key_i(u64 k) { return (int)(unsigned)(k & 0xffffffffu); }

struct Smem {
    u64 *kb;      // C*Wp keys per bin: [0,KR) sorted kept list | [KR,KR+B) unsorted bucket | KEY_MAX padding to Wp
    u64 *ko;      // C*KR compaction output
    double *le;   // C*KR exact distance of a list slot (re-rank), then the final exact lists
    double *q_d;  // CHUNK (pass 2)
    double *xq;   // d
    float *thr;   // C : keys with A <= thr can matter (A_(k) + 2E once k keys are known, +inf before)
    float *thr2;  // C : threshold being built by a compaction round
    int *incnt;   // C : keys inside the new slack (overflow detection)
    int *dirtyb;  // C : a bucket key entered the kept list
    int *li;      // C*KR final exact lists (indices)
    int *xw;      // C*KR exact-distance work items (c*KR + slot)
    int *q_i, *q_c; // CHUNK (pass 2)
    int *cnt, *bcnt, *flags; // C ; flags: bit0 kept set changed, bit1 overflow
    int *ctl;     // [0] pending/compaction request, [1] any overflow, [2] exact work count, [3] pass-2 queue length
};

__host__ __device__ inline size_t smem_layout(unsigned char *base, int C, int KR, int Wp, int d, Smem *s)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~size_t(15); return o; };
    const size_t o_kb = take(sizeof(u64) * (size_t)C * Wp), o_ko = take(sizeof(u64) * (size_t)C * KR),
                 o_t2 = take(sizeof(float) * (size_t)C), o_ic = take(sizeof(int) * (size_t)C),
                 o_db = take(sizeof(int) * (size_t)C), o_le = take(sizeof(double) * (size_t)C * KR),
                 o_qd = take(sizeof(double) * CHUNK), o_xq = take(sizeof(double) * (size_t)d),
                 o_thr = take(sizeof(float) * (size_t)C), o_li = take(sizeof(int) * (size_t)C * KR),
                 o_xw = take(sizeof(int) * (size_t)C * KR), o_qi = take(sizeof(int) * CHUNK),
                 o_qc = take(sizeof(int) * CHUNK), o_cnt = take(sizeof(int) * (size_t)C),
                 o_bc = take(sizeof(int) * (size_t)C), o_fl = take(sizeof(int) * (size_t)C), o_ctl = take(sizeof(int) * 4);
    if (s) {
        s->kb = reinterpret_cast<u64 *>(base + o_kb);
        s->ko = reinterpret_cast<u64 *>(base + o_ko);
        s->thr2 = reinterpret_cast<float *>(base + o_t2);
        s->incnt = reinterpret_cast<int *>(base + o_ic);
        s->dirtyb = reinterpret_cast<int *>(base + o_db);
        s->le = reinterpret_cast<double *>(base + o_le);
        s->q_d = reinterpret_cast<double *>(base + o_qd);
        s->xq = reinterpret_cast<double *>(base + o_xq);
        s->thr = reinterpret_cast<float *>(base + o_thr);
        s->li = reinterpret_cast<int *>(base + o_li);
        s->xw = reinterpret_cast<int *>(base + o_xw);
        s->q_i = reinterpret_cast<int *>(base + o_qi);
        s->q_c = reinterpret_cast<int *>(base + o_qc);
        s->cnt = reinterpret_cast<int *>(base + o_cnt);
        s->bcnt = reinterpret_cast<int *>(base + o_bc);
        s->flags = reinterpret_cast<int *>(base + o_fl);
        s->ctl = reinterpret_cast<int *>(base + o_ctl);
    }
    return off;
}

__device__ __forceinline__ double exact_distance(const double *__restrict__ xq_s, const double *__restrict__ xi, int d)
{
    // scipy cdist 'euclidean': s = 0; s += (u[t]-v[t])^2 in ascending t, separate multiply and add; sqrt.
    // The additions form one dependent chain (that IS the recipe); the row loads are software-pipelined 8 deep so
    // that chain, not memory latency, is the critical path.
    double acc = 0.0;
    const int d8 = d & ~7;
    double v[8];
    if (d8 > 0) {
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = xi[u];
    }
    for (int t = 0; t < d8; t += 8) {
        double sq[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double df = __dsub_rn(xq_s[t + u], v[u]);
            sq[u] = __dmul_rn(df, df);
        }
        if (t + 8 < d8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = xi[t + 8 + u];
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc = __dadd_rn(acc, sq[u]);
    }
    for (int t = d8; t < d; ++t) {
        const double df = __dsub_rn(xq_s[t], xi[t]);
        acc = __dadd_rn(acc, __dmul_rn(df, df));
    }
    return __dsqrt_rn(acc);
}

// One CTA-wide compaction round: every bin whose bucket holds at least `min_fill` keys merges it into its sorted
// kept list.  Thread-parallel "rank by counting": one (bin, slot) item per thread and step, so the critical path
// is one pass over a bin's Wp keys instead of a warp working through several bins one after the other.
// All threads of the CTA must call this (it contains barriers).
__device__ __forceinline__ void compact_round(const Smem &s, int C, int k, int KR, int B, int Wp, float slack2, int min_fill,
                                              bool mark_dirty, int tid)
{
    const int items = C * Wp;
    // phase A: rank every valid key of the participating bins, scatter the KR smallest, find the new threshold
    for (int it = tid; it < items; it += NT) {
        const int c = it / Wp, slot = it - c * Wp;
        const int nb = s.bcnt[c];
        if (nb < min_fill) continue;
        const u64 key = s.kb[it];
        if (key == KEY_MAX) continue;
        const u64 *row = s.kb + c * Wp;
        int rank = 0;
        for (int t = 0; t < Wp; t += 8) {
            u64 o[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) o[u] = row[t + u];
#pragma unroll
            for (int u = 0; u < 8; ++u) rank += (o[u] < key);
        }
        if (rank < KR) {
            s.ko[c * KR + rank] = key;
            if (slot >= KR) s.dirtyb[c] = 1;
        }
        if (rank == k - 1) s.thr2[c] = __fadd_ru(key_a(key), slack2);
    }
    __syncthreads();
    // phase B: count the keys inside the new slack, then rewrite the bin: kept list from `ko`, bucket cleared
    for (int it = tid; it < items; it += NT) {
        const int c = it / Wp, slot = it - c * Wp;
        const int nb = s.bcnt[c];
        if (nb < min_fill) continue;
        const int tot = s.cnt[c] + (nb < B ? nb : B);
        const u64 key = s.kb[it];
        if (key != KEY_MAX && tot >= k && key_a(key) <= s.thr2[c]) atomicAdd(&s.incnt[c], 1);
        const int ncnt = tot < KR ? tot : KR;
        s.kb[it] = (slot < ncnt) ? s.ko[c * KR + slot] : KEY_MAX;
    }
    __syncthreads();
    // phase C: per-bin bookkeeping
    for (int c = tid; c < C; c += NT) {
        const int nb = s.bcnt[c];
        if (nb < min_fill) continue;
        const int tot = s.cnt[c] + (nb < B ? nb : B);
        s.cnt[c] = tot < KR ? tot : KR;
        s.bcnt[c] = 0;
        s.thr[c] = (tot >= k) ? s.thr2[c] : INFINITY;
        int f = s.flags[c];
        if (s.dirtyb[c] && mark_dirty) f |= 1;
        if (s.incnt[c] > KR) f |= 3; // more keys inside the slack than the kept list can hold
        s.flags[c] = f;
        s.incnt[c] = 0;
        s.dirtyb[c] = 0;
    }
    __syncthreads();
}

// pass 2: exact-keyed insert into the first k slots of (le, li)
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

// effective labels of 4 consecutive points starting at i (i % 4 == 0, i + 3 < n)
__device__ __forceinline__ void eff_labels4(const chb_knn_args &a, int64_t i, int p, int (&c)[4])
{
    if (a.packed) {
        const int4 q0 = __ldg(reinterpret_cast<const int4 *>(a.packed + i));
        const int4 q1 = __ldg(reinterpret_cast<const int4 *>(a.packed + i + 2));
        const int ps[4] = {q0.x, q0.z, q1.x, q1.z};
        const int lb[4] = {q0.y, q0.w, q1.y, q1.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) c[e] = (ps[e] < p) ? (lb[e] >> 16) : (int)(short)(lb[e] & 0xffff);
    } else if (a.mode == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) c[e] = (a.pos[i + e] < p) ? a.tent_pt[i + e] : a.old_label[i + e];
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) c[e] = a.old_label[i + e];
    }
}
__device__ __forceinline__ int eff_label1(const chb_knn_args &a, int64_t i, int p)
{
    if (a.mode == 0) return (a.pos[i] < p) ? a.tent_pt[i] : a.old_label[i];
    return a.old_label[i];
}

__global__ void __launch_bounds__(NT, 4) knn_scan_kernel(chb_knn_args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = a.C, k = a.k, d = a.d;
    const int KR = (k + 3 < 32) ? k + 3 : 32;
    const int B = (KR <= 16) ? 16 : 32;
    const int Wp = (KR + B + 7) & ~7;
    Smem s;
    smem_layout(smem_raw, C, KR, Wp, d, &s);
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
    const bool vec_ok = ((a.row_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.arows) & 15) == 0);

    {
        const double *__restrict__ xj = a.X + (int64_t)j * a.ldx;
        for (int t = tid; t < d; t += NT) s.xq[t] = xj[t];
    }
    const double nmax = (double)__uint_as_float(*a.nrm_max_bits);
    const float slack2 = __double2float_ru(2.0 * (a.eps_rel * ((double)a.nrm[j] + nmax) + 1e-30));
    for (int c = tid; c < C; c += NT) {
        s.cnt[c] = 0;
        s.bcnt[c] = 0;
        s.thr[c] = INFINITY;
        s.flags[c] = 0;
        s.incnt[c] = 0;
        s.dirtyb[c] = 0;
    }
    for (int t = tid; t < C * Wp; t += NT) s.kb[t] = KEY_MAX;
    if (tid < 4) s.ctl[tid] = 0;
    __syncthreads();

    // ---- warm start: cached members that still belong to the bin seed the list (via the bucket)
    if (a.mode == 0) {
        for (int c = warp; c < C; c += NW) {
            const int m = ccnt[c];
            if (m < 0) { if (lane == 0) s.flags[c] = 1; continue; }
            bool ok = false;
            int ci = -1;
            if (lane < m) {
                ci = cidx[c * k + lane];
                ok = (eff_label1(a, ci, p) == c) && (ci != j);
            }
            const unsigned keep = __ballot_sync(CHB_FULL, ok);
            const int kept = __popc(keep);
            if (ok) s.kb[c * Wp + KR + __popc(keep & ((1u << lane) - 1u))] = make_key(arow[ci], ci); // kept <= k <= B
            if (lane == 0) {
                s.bcnt[c] = kept;
                if (kept != m) s.flags[c] = 1;
            }
        }
        __syncthreads();
        compact_round(s, C, k, KR, B, Wp, slack2, 1, false, tid);
    }

    // ---- pass 1: stream the FP32 row (next step's keys and labels are prefetched into registers)
    float4 nf = make_float4(0.f, 0.f, 0.f, 0.f);
    int4 nq0 = make_int4(0, 0, 0, 0), nq1 = make_int4(0, 0, 0, 0);
    const bool fast = vec_ok && a.packed != nullptr;
    if (fast) {
        const int64_t i0 = (int64_t)tid * EPT;
        if (i0 + EPT <= n) {
            nf = __ldg(reinterpret_cast<const float4 *>(arow + i0));
            nq0 = __ldg(reinterpret_cast<const int4 *>(a.packed + i0));
            nq1 = __ldg(reinterpret_cast<const int4 *>(a.packed + i0 + 2));
        }
    }
    for (int64_t base = 0; base < n; base += CHUNK) {
        const int64_t i0 = base + (int64_t)tid * EPT;
        float av[EPT];
        int cv[EPT];
        unsigned pend = 0;
        if (fast && i0 + EPT <= n) {
            av[0] = nf.x; av[1] = nf.y; av[2] = nf.z; av[3] = nf.w;
            const int ps[4] = {nq0.x, nq0.z, nq1.x, nq1.z};
            const int lb[4] = {nq0.y, nq0.w, nq1.y, nq1.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) cv[e] = (ps[e] < p) ? (lb[e] >> 16) : (int)(short)(lb[e] & 0xffff);
        } else if (i0 + EPT <= n && vec_ok) {
            const float4 f = __ldg(reinterpret_cast<const float4 *>(arow + i0));
            av[0] = f.x; av[1] = f.y; av[2] = f.z; av[3] = f.w;
            eff_labels4(a, i0, p, cv);
        } else {
#pragma unroll
            for (int e = 0; e < EPT; ++e) {
                const int64_t i = i0 + e;
                av[e] = i < n ? arow[i] : 0.f;
                cv[e] = i < n ? eff_label1(a, i, p) : -1;
            }
        }
        if (fast) {
            const int64_t i1 = i0 + CHUNK;
            if (i1 + EPT <= n) {
                nf = __ldg(reinterpret_cast<const float4 *>(arow + i1));
                nq0 = __ldg(reinterpret_cast<const int4 *>(a.packed + i1));
                nq1 = __ldg(reinterpret_cast<const int4 *>(a.packed + i1 + 2));
            }
        }
#pragma unroll
        for (int e = 0; e < EPT; ++e)
            if (cv[e] >= 0 && cv[e] < C && (i0 + e) != j && av[e] <= s.thr[cv[e]]) pend |= 1u << e;

        for (;;) {
            bool myreq = false;
            if (pend) {
#pragma unroll
                for (int e = 0; e < EPT; ++e) {
                    if (!((pend >> e) & 1u)) continue;
                    const int c = cv[e];
                    if (!(av[e] <= s.thr[c])) { pend &= ~(1u << e); continue; }
                    const int idx = (int)(i0 + e);
                    const u64 *row = s.kb + c * Wp;
                    bool dup = false;
                    for (int t = 0; t < KR; ++t) dup = dup || (key_i(row[t]) == idx); // empty slots hold index 0xffffffff
                    if (dup) { pend &= ~(1u << e); continue; }
                    const int slot = atomicAdd(&s.bcnt[c], 1);
                    if (slot < B) {
                        s.kb[c * Wp + KR + slot] = make_key(av[e], idx);
                        pend &= ~(1u << e);
                        if (slot >= B / 2) myreq = true;
                    } else {
                        myreq = true; // bucket full: retry after the compaction
                    }
                }
            }
            if (!__syncthreads_or(myreq)) break;
            compact_round(s, C, k, KR, B, Wp, slack2, B / 2, true, tid);
            if (!__syncthreads_or(pend != 0)) break;
        }
    }
    // final compaction of every non-empty bucket
    compact_round(s, C, k, KR, B, Wp, slack2, 1, true, tid);

    // ---- re-rank R1: per changed bin decide which kept keys need an exact distance
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    for (int c = warp; c < C; c += NW) {
        const int fl = s.flags[c];
        if (!(fl & 1)) continue;
        if (fl & 2) { if (lane == 0) s.ctl[1] = 1; continue; }
        const int m = s.cnt[c];
        const bool have = lane < m;
        const u64 key = have ? s.kb[c * Wp + lane] : KEY_MAX;
        const bool cand = have && (m < k || key_a(key) <= s.thr[c]);
        double e = qnan;
        if (cand && a.mode == 0) {
            const int mo = ccnt[c];
            const int idx = key_i(key);
            for (int t = 0; t < mo; ++t)
                if (cidx[c * k + t] == idx) e = cdst[c * k + t];
        }
        if (have) s.le[c * KR + lane] = cand ? e : INFINITY; // INFINITY = not a candidate
        const bool need = cand && !(e == e);
        const unsigned nm = __ballot_sync(CHB_FULL, need);
        if (nm) {
            int b = 0;
            if (lane == 0) b = atomicAdd(&s.ctl[2], __popc(nm));
            b = __shfl_sync(CHB_FULL, b, 0);
            if (need) s.xw[b + __popc(nm & ((1u << lane) - 1u))] = c * KR + lane;
        }
    }
    __syncthreads();
    // ---- re-rank R2: exact distances, one thread per key, CTA-wide
    {
        const int nx = s.ctl[2];
        for (int e = tid; e < nx; e += NT) {
            const int w = s.xw[e];
            s.le[w] = exact_distance(s.xq, a.X + (int64_t)key_i(s.kb[(w / KR) * Wp + (w % KR)]) * a.ldx, d);
        }
    }
    __syncthreads();
    // ---- re-rank R3: select the k smallest exact (distance, index) pairs of each changed bin
    for (int c = warp; c < C; c += NW) {
        const int fl = s.flags[c];
        if ((fl & 3) != 1) continue;
        const int m = s.cnt[c];
        const bool have = lane < m;
        const double me = have ? s.le[c * KR + lane] : INFINITY;
        const int mi = have ? key_i(s.kb[c * Wp + lane]) : INT32_MAX;
        const bool cand = have && (me < INFINITY);
        const unsigned candm = __ballot_sync(CHB_FULL, cand);
        int rank = 0;
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

    // ---- pass 2 (rare): bins with more than KR keys inside the slack are re-streamed and ranked exactly
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
                    c = eff_label1(a, i, p);
                    pass = (c >= 0) && (c < C) && (i != j) && (s.flags[c] & 2) && (fv <= s.thr[c]);
                }
                const unsigned m = __ballot_sync(CHB_FULL, pass);
                if (m) {
                    int b = 0;
                    const int leader = __ffs(m) - 1;
                    if (lane == leader) b = atomicAdd(&s.ctl[3], __popc(m));
                    b = __shfl_sync(CHB_FULL, b, leader);
                    if (pass) {
                        const int slot = b + __popc(m & ((1u << lane) - 1u));
                        s.q_i[slot] = (int)i;
                        s.q_c[slot] = c;
                    }
                }
            }
            __syncthreads();
            const int qn = s.ctl[3];
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
                if (tid == 0) s.ctl[3] = 0;
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
            const bool touched = s.flags[c] & 1;
            if (lane < k) cidx[c * k + lane] = (touched && lane < m) ? s.li[c * KR + lane] : -1;
            if (lane == 0) ccnt[c] = touched ? m : 0;
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
    const int B = (KR <= 16) ? 16 : 32;
    const int Wp = (KR + B + 7) & ~7;
    const size_t bytes = smem_layout(nullptr, a.C, KR, Wp, a.d, nullptr) + 16;
    CHB_CHECK(ctx, bytes <= 227 * 1024, CHB_EINVAL, "num_clusters*num_neighbors too large for the kNN kernel (%zu B smem)",
              bytes);
    // a per-device attribute: set per call, a context may live on any device of this process
    CHB_CUDA(ctx, cudaFuncSetAttribute(knn_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    {
        chb_stage_timer t(ctx, CHB_ST_KNN);
        knn_scan_kernel<<<(unsigned)a.n_items, NT, bytes, ctx->stream>>>(a);
    }
    CHB_CUDA(ctx, cudaGetLastError());
    ctx->tm.rows_scanned += a.n_items;
    return CHB_OK;
}
