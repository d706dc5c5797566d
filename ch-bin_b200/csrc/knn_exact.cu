// knn_exact.cu -- label-segmented top-k over one EXACT distance row per query (distance mode 0) (sm_100a).
// (The default path, FP32 keys + exact re-rank, is knn.cu.)
//
// Replaces find_nearest_from_cluster (/root/reference/ch_bin/core/clustering/distance_matrix.py:47-62) for ALL
// bins of a query in one pass: the reference calls it once per (query, bin), each call an O(n) np.where plus an
// argpartition; here one CTA streams the query's distance row once (HBM-bound: 8 B per point) and keeps, per
// bin, the k smallest (distance, index) pairs in shared memory.  "All members when |bin| <= k"
// (distance_matrix.py:58-59) is the same selection.  Ties at the k-th distance are broken by the lower point
// index (numpy's argpartition leaves them unspecified).
//
// Mode 0 (assignment rounds): the label a query at permutation position p sees for point i is
//        pos[i] < p ? tent_pt[i] : old_label[i]          (algorithm.py:46-60: earlier points of this iteration
//        are already re-assigned, later ones still carry last iteration's label; the query itself is removed).
//   Lists are warm-started from the per-(query, bin) cache of the previous round, so in steady state the scan
//   is a pure threshold filter; a pair is pushed to the QP work list only if its neighbour list changed.
// Mode 1 (chb_knn_per_bin): plain snapshot labels, cold start, lists written out per item.
//
// Row flavours.  filter = 0: the row holds the exact FP64 distances (distance.cu).  filter = 1 (default): the row
// holds FP32 approximations A of the SQUARED distance with |A - d^2| <= E (approx.cu); a point is a candidate for
// bin c when A <= thr_c^2 (1 + 2^-50) + E, which can never reject a point whose exact distance is <= thr_c; each
// candidate that is not already in the list gets its exact distance from the scipy recipe (sequential sum, no
// FMA, one thread per candidate) and only exact (distance, index) pairs are ever ranked, so the selected sets are
// bit-identical to ranking the full exact row.
#include <cfloat>

#include "common.cuh"

namespace {

constexpr int NT = 256;          // threads per CTA
constexpr int NW = NT / 32;      // warps per CTA
constexpr int EPT = 4;           // row elements per thread per chunk
constexpr int CHUNK = NT * EPT;  // queue capacity = worst case of one chunk

struct Smem {
    double *list_d; // C*k
    double *q_d;    // CHUNK
    int *list_i;    // C*k
    int *q_i;       // CHUNK
    int *q_c;       // CHUNK
    int *cnt;       // C
    float *thr;     // C   float upper bound of the current k-th distance (+inf while the list is not full)
    int *dirty;     // C
    int *qn;        // 1
    double *xq;     // d  (filter mode: the query's feature row)
};

__device__ __forceinline__ Smem carve(unsigned char *base, int C, int k, int d)
{
    Smem s;
    size_t off = 0;
    s.list_d = reinterpret_cast<double *>(base + off); off += sizeof(double) * (size_t)C * k;
    s.q_d = reinterpret_cast<double *>(base + off);    off += sizeof(double) * CHUNK;
    s.xq = reinterpret_cast<double *>(base + off);     off += sizeof(double) * (size_t)((d + 1) & ~1);
    s.list_i = reinterpret_cast<int *>(base + off);    off += sizeof(int) * (size_t)C * k;
    s.q_i = reinterpret_cast<int *>(base + off);       off += sizeof(int) * CHUNK;
    s.q_c = reinterpret_cast<int *>(base + off);       off += sizeof(int) * CHUNK;
    s.cnt = reinterpret_cast<int *>(base + off);       off += sizeof(int) * (size_t)C;
    s.thr = reinterpret_cast<float *>(base + off);     off += sizeof(float) * (size_t)C;
    s.dirty = reinterpret_cast<int *>(base + off);     off += sizeof(int) * (size_t)C;
    s.qn = reinterpret_cast<int *>(base + off);
    return s;
}

size_t smem_bytes(int C, int k, int d)
{
    return sizeof(double) * ((size_t)C * k + CHUNK + (size_t)((d + 1) & ~1)) +
           sizeof(int) * ((size_t)C * k + 2 * CHUNK + 3 * (size_t)C + 4) + 16;
}

// threshold the streamed value is compared with, from the exact k-th distance of a full list
__device__ __forceinline__ float make_thr(double kth, int filter, float slack)
{
    if (!filter) return __double2float_ru(kth);
    return __fadd_ru(__double2float_ru(kth * kth * (1.0 + 8.9e-16)), slack);
}

// Warp-cooperative insert of (cd, ci) into the sorted list of one bin; every lane passes the same candidate.
// Lane l mirrors list slot l (k <= 32).
__device__ __forceinline__ void warp_insert(double *ld, int *li, int *cnt_p, float *thr_p, int *dirty_p, int k, double cd,
                                            int ci, int lane, int filter, float slack)
{
    const int n = *cnt_p;
    const bool have = lane < n;
    const double md = have ? ld[lane] : 0.0;
    const int mi = have ? li[lane] : -1;
    if (__any_sync(CHB_FULL, have && mi == ci)) return; // already a member of the list
    const bool less = have && (md < cd || (md == cd && mi < ci));
    const int at = __popc(__ballot_sync(CHB_FULL, less));
    if (at >= k) return;
    __syncwarp();
    if (have && lane >= at && lane + 1 < k) {
        ld[lane + 1] = md;
        li[lane + 1] = mi;
    }
    if (lane == 0) {
        ld[at] = cd;
        li[at] = ci;
        *cnt_p = n + 1 < k ? n + 1 : k;
        *dirty_p = 1;
    }
    __syncwarp();
    if (lane == 0 && *cnt_p == k) *thr_p = make_thr(ld[k - 1], filter, slack);
    __syncwarp();
}

__global__ void __launch_bounds__(NT) knn_scan_exact_kernel(chb_knn_args a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = a.C, k = a.k;
    Smem s = carve(smem_raw, C, k, a.d);
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
    const int filter = a.filter;
    const double *__restrict__ row = filter ? nullptr : a.rows + rowi * a.row_stride;
    const float *__restrict__ arow = filter ? a.arows + rowi * a.row_stride : nullptr;
    int32_t *cidx = a.knn_idx + cache_row * (int64_t)C * k;
    int32_t *ccnt = a.knn_cnt + cache_row * (int64_t)C;
    double *cdst = a.knn_dist ? a.knn_dist + cache_row * (int64_t)C * k : nullptr;
    float slack = 0.f;
    if (filter) {
        const double *__restrict__ xj = a.X + (int64_t)j * a.ldx;
        for (int t = tid; t < a.d; t += NT) s.xq[t] = xj[t];
        const double nmax = (double)__uint_as_float(*a.nrm_max_bits);
        slack = __double2float_ru(a.eps_rel * ((double)a.nrm[j] + nmax) + 1e-30);
    }

    for (int c = tid; c < C; c += NT) {
        s.cnt[c] = 0;
        s.thr[c] = INFINITY;
        s.dirty[c] = 0;
    }
    if (tid == 0) *s.qn = 0;
    __syncthreads();

    // ---- warm start from the cached lists: keep the entries that are still members of the bin
    if (a.mode == 0) {
        for (int c = warp; c < C; c += NW) {
            const int m = ccnt[c];
            if (m < 0) { if (lane == 0) s.dirty[c] = 1; continue; } // never computed
            int ci = -1;
            double cd = 0.0;
            bool ok = false;
            if (lane < m) {
                ci = cidx[c * k + lane];
                const int pi = a.pos[ci];
                const int lab = (pi < p) ? a.tent_pt[ci] : a.old_label[ci];
                ok = (lab == c) && (ci != j);
                if (ok) cd = filter ? cdst[c * k + lane] : row[ci];
            }
            // cached order is canonical, so surviving entries stay sorted: compact them
            const unsigned keep = __ballot_sync(CHB_FULL, ok);
            const int dst = __popc(keep & ((1u << lane) - 1u));
            if (ok) {
                s.list_d[c * k + dst] = cd;
                s.list_i[c * k + dst] = ci;
            }
            const int kept = __popc(keep);
            if (lane == 0) {
                s.cnt[c] = kept;
                if (kept != m) s.dirty[c] = 1;
            }
            __syncwarp();
            if (lane == 0 && kept == k) s.thr[c] = make_thr(s.list_d[c * k + k - 1], filter, slack);
        }
        __syncthreads();
    }

    // ---- stream the row
    for (int64_t base = 0; base < n; base += CHUNK) {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
            const int64_t i = base + (int64_t)e * NT + tid;
            bool pass = false;
            double dv = 0.0;
            int c = -1;
            if (i < n) {
                float fv;
                if (filter) fv = arow[i];
                else { dv = row[i]; fv = __double2float_rd(dv); }
                if (a.mode == 0) {
                    const int pi = a.pos[i];
                    c = (pi < p) ? a.tent_pt[i] : a.old_label[i];
                } else {
                    c = a.old_label[i];
                }
                pass = (c >= 0) && (c < C) && (i != j) && (fv <= s.thr[c]);
            }
            const unsigned m = __ballot_sync(CHB_FULL, pass);
            if (m) {
                int b = 0;
                const int leader = __ffs(m) - 1;
                if (lane == leader) b = atomicAdd(s.qn, __popc(m));
                b = __shfl_sync(CHB_FULL, b, leader);
                if (pass) {
                    const int slot = b + __popc(m & ((1u << lane) - 1u));
                    s.q_d[slot] = dv;
                    s.q_i[slot] = (int)i;
                    s.q_c[slot] = c;
                }
            }
        }
        __syncthreads();
        const int qn = *s.qn;
        if (qn > 0 && filter) {
            // exact scipy-recipe distance for every candidate that is not already listed; one thread each
            for (int e = tid; e < qn; e += NT) {
                const int ci = s.q_i[e], cc = s.q_c[e];
                const int m = s.cnt[cc];
                bool dup = false;
                for (int t = 0; t < m; ++t) dup = dup || (s.list_i[cc * k + t] == ci);
                if (dup) { s.q_c[e] = -1; continue; }
                const double *__restrict__ xi = a.X + (int64_t)ci * a.ldx;
                double acc = 0.0;
                int t = 0;
                for (; t + 4 <= a.d; t += 4) {
                    const double v0 = xi[t], v1 = xi[t + 1], v2 = xi[t + 2], v3 = xi[t + 3];
                    const double d0 = __dsub_rn(s.xq[t], v0), d1 = __dsub_rn(s.xq[t + 1], v1);
                    const double d2 = __dsub_rn(s.xq[t + 2], v2), d3 = __dsub_rn(s.xq[t + 3], v3);
                    acc = __dadd_rn(acc, __dmul_rn(d0, d0));
                    acc = __dadd_rn(acc, __dmul_rn(d1, d1));
                    acc = __dadd_rn(acc, __dmul_rn(d2, d2));
                    acc = __dadd_rn(acc, __dmul_rn(d3, d3));
                }
                for (; t < a.d; ++t) {
                    const double d0 = __dsub_rn(s.xq[t], xi[t]);
                    acc = __dadd_rn(acc, __dmul_rn(d0, d0));
                }
                const double dv = __dsqrt_rn(acc);
                if (m == k) {
                    const double kd = s.list_d[cc * k + k - 1];
                    const int ki = s.list_i[cc * k + k - 1];
                    if (dv > kd || (dv == kd && ci > ki)) { s.q_c[e] = -1; continue; }
                }
                s.q_d[e] = dv;
            }
            __syncthreads();
        }
        if (qn > 0) {
            // each warp owns the bins c with c % NW == warp: no two warps ever touch the same list
            for (int b = 0; b < qn; b += 32) {
                const int e = b + lane;
                double cd = 0.0;
                int ci = 0, cc = -1;
                if (e < qn) {
                    cd = s.q_d[e];
                    ci = s.q_i[e];
                    cc = s.q_c[e];
                }
                unsigned mine = __ballot_sync(CHB_FULL, (e < qn) && (cc >= 0) && (cc % NW == warp));
                while (mine) {
                    const int src = __ffs(mine) - 1;
                    mine &= mine - 1;
                    const double bd = __shfl_sync(CHB_FULL, cd, src);
                    const int bi = __shfl_sync(CHB_FULL, ci, src);
                    const int bc = __shfl_sync(CHB_FULL, cc, src);
                    warp_insert(s.list_d + bc * k, s.list_i + bc * k, s.cnt + bc, s.thr + bc, s.dirty + bc, k, bd, bi, lane, filter, slack);
                }
            }
            __syncthreads();
            if (tid == 0) *s.qn = 0;
            __syncthreads();
        }
    }

    // ---- write back: lists that changed go to the cache and onto the QP work list
    for (int c = warp; c < C; c += NW) {
        const int m = s.cnt[c];
        if (a.mode == 0) {
            if (!s.dirty[c]) continue;
            // a dropped-and-reinserted entry leaves the list identical: compare before declaring it changed
            const int mo = ccnt[c];
            bool same = (mo == m);
            if (same) {
                const bool diff = (lane < m) && (cidx[c * k + lane] != s.list_i[c * k + lane]);
                same = !__any_sync(CHB_FULL, diff);
            }
            if (same) continue;
            if (lane < k) cidx[c * k + lane] = lane < m ? s.list_i[c * k + lane] : -1;
            if (cdst && lane < m) cdst[c * k + lane] = s.list_d[c * k + lane];
            if (lane == 0) {
                ccnt[c] = m;
                const int w = atomicAdd(a.work_count, 1);
                a.work[w] = make_int2((int)cache_row, c);
            }
        } else {
            if (lane < k) cidx[c * k + lane] = lane < m ? s.list_i[c * k + lane] : -1;
            if (lane == 0) ccnt[c] = m;
        }
    }
}

} // namespace

int chb_launch_knn_scan_exact(chb_ctx *ctx, const chb_knn_args &a)
{
    if (a.n_items <= 0) return CHB_OK;
    CHB_CHECK(ctx, a.k >= 1 && a.k <= CHB_KMAX, CHB_EINVAL, "num_neighbors must be in [1, %d]", CHB_KMAX);
    const size_t bytes = smem_bytes(a.C, a.k, a.d);
    CHB_CHECK(ctx, bytes <= 227 * 1024, CHB_EINVAL, "num_clusters*num_neighbors too large for the kNN kernel (%zu B smem)",
              bytes);
    // a per-device attribute: set per call, a context may live on any device of this process
    CHB_CUDA(ctx, cudaFuncSetAttribute(knn_scan_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    {
        chb_stage_timer t(ctx, CHB_ST_KNN);
        knn_scan_exact_kernel<<<(unsigned)a.n_items, NT, bytes, ctx->stream>>>(a);
    }
    CHB_CUDA(ctx, cudaGetLastError());
    ctx->tm.rows_scanned += a.n_items;
    return CHB_OK;
}
