// waafle_genecaller on the device: gene calls from BLAST hits, one warp per contig block.
//
// Reference: waafle/waafle_genecaller.py:107-170 (hits2ints, overlap_intervals, merge_inodes) with
// waafle/utils.py:455-500 (INode, calc_overlap).  Per contig: the hits that pass --min-scov become intervals; sorted by
// start (stable); two intervals are linked if they overlap by >= --min-overlap of the SHORTER one (the scan of later
// intervals stops at the first one that does not overlap at all); connected components are merged into genes
// [min start, max stop] with the strand of the longest member ('-' wins ties); genes shorter than --min-gene-length are
// dropped; components come out in the order of their first member.
//
// Here: bitonic sort of (start, file position) keys, label propagation with atomicMin until nothing changes (labels end as
// the smallest sorted index of the component, so the reference's output order is the ascending label order), per-root
// reductions, ballot compaction of the surviving genes.  All state lives in a global scratch the size of the hit arrays.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "waafle_b200.h"

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;
constexpr u32 FULL = 0xffffffffu;
constexpr int GC_WPC = 4;

struct GcArgs {
    int64_t n_blocks;
    const int64_t *block_off;
    const int32_t *qstart, *qend;
    const int8_t *strand;
    const uint8_t *keep;
    double thr, min_len;
    u64 *key;            // [2 * n_hits + 32 * n_blocks]: sort keys, padded to a power of two per block
    int *sa, *sb;        // [n_hits] sorted starts / stops
    int *lab;            // [n_hits] component label = smallest sorted index
    int *cs, *ce;        // [n_hits] per-root min start / max stop
    u64 *cbest;          // [n_hits] per-root max (length << 8 | strand char)
    int8_t *ss;          // [n_hits] sorted strands
    int32_t *g_start, *g_end;   // genes of block b at [block_off[b], block_off[b] + g_count[b])
    int8_t *g_strand;
    int32_t *g_count;
    unsigned long long *wq;
};

__global__ void __launch_bounds__(32 * GC_WPC) wfl_genecall(const GcArgs a) {
    const int lane = threadIdx.x & 31;
    const u32 lt = (1u << lane) - 1u;
    for (;;) {
        long long b = -1;
        if (lane == 0) {
            const unsigned long long w = atomicAdd(a.wq, 1ull);
            if (w < (unsigned long long)a.n_blocks) b = (long long)w;
        }
        b = __shfl_sync(FULL, b, 0);
        if (b < 0) break;
        const long long h0 = a.block_off[b], h1 = a.block_off[b + 1];
        u64 *key = a.key + 2 * h0 + 32 * b;
        int *sa = a.sa + h0, *sb = a.sb + h0, *lab = a.lab + h0, *cs = a.cs + h0, *ce = a.ce + h0;
        u64 *cbest = a.cbest + h0;
        int8_t *ss = a.ss + h0;
        // ---- intervals of the hits that passed the scov filter (hits2ints :107-113), keyed by (start, file position) ----
        int m = 0;
        for (long long base = h0; base < h1; base += 32) {
            const long long h = base + lane;
            const bool k = h < h1 && a.keep[h];
            const u32 bal = __ballot_sync(FULL, k);
            if (k) {
                const int q1 = a.qstart[h], q2 = a.qend[h];
                key[m + __popc(bal & lt)] = ((u64)(u32)min(q1, q2) << 32) | (u64)(u32)(h - h0);
            }
            m += __popc(bal);
        }
        int P = 32;
        while (P < m) P <<= 1;
        for (int j = m + lane; j < P; j += 32) key[j] = ~0ull;
        __syncwarp();
        // ---- sorted by start, file order among equal starts (sorted(inodes, key=start) :140) ----
        for (int k = 2; k <= P; k <<= 1)
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                for (int x = lane; x < (P >> 1); x += 32) {
                    const int lo = ((x & ~(jj - 1)) << 1) | (x & (jj - 1)), hi = lo | jj;
                    const u64 u = key[lo], w = key[hi];
                    if ((u > w) == ((lo & k) == 0)) { key[lo] = w; key[hi] = u; }
                }
                __syncwarp();
            }
        for (int j = lane; j < m; j += 32) {
            const long long h = h0 + (long long)(key[j] & 0xffffffffull);
            const int q1 = a.qstart[h], q2 = a.qend[h];
            sa[j] = min(q1, q2);
            sb[j] = max(q1, q2);
            ss[j] = a.strand[h];
            lab[j] = j;
        }
        __syncwarp();
        // ---- links (:141-152) + connected components (:154-157) by label propagation ----
        for (;;) {
            bool changed = false;
            for (int j = lane; j < m; j += 32) {
                const int aj = sa[j], bj = sb[j], lenj = bj - aj + 1;
                for (int k = j + 1; k < m; ++k) {
                    const int ak = sa[k], bk = sb[k];
                    double score = 0.0;   // calc_overlap (utils.py:487-500); sorted by start: aj <= ak
                    if (!(ak > bj)) score = (double)(min(bj, bk) - ak + 1) / (double)min(lenj, bk - ak + 1);
                    if (score >= a.thr) {
                        const int lj = lab[j], lk = lab[k], l = min(lj, lk);
                        if (lj > l) { atomicMin(&lab[j], l); changed = true; }
                        if (lk > l) { atomicMin(&lab[k], l); changed = true; }
                    } else if (score == 0.0) {
                        break;   // no further interval can overlap this one (:151-152)
                    }
                }
            }
            __syncwarp();
            for (int j = lane; j < m; j += 32) {   // pointer jumping
                int l = lab[j];
                while (lab[l] < l) l = lab[l];
                if (l < lab[j]) { lab[j] = l; changed = true; }
            }
            __syncwarp();
            if (!__any_sync(FULL, changed)) break;
        }
        // ---- merge_inodes (:121-135): [min start, max stop], strand of the longest member ('-' > '+' among ties) ----
        for (int j = lane; j < m; j += 32) {
            cs[j] = 0x7fffffff;
            ce[j] = (int)0x80000000;
            cbest[j] = 0ull;
        }
        __syncwarp();
        for (int j = lane; j < m; j += 32) {
            const int r = lab[j];
            atomicMin(&cs[r], sa[j]);
            atomicMax(&ce[r], sb[j]);
            atomicMax(&cbest[r], ((u64)(u32)(sb[j] - sa[j] + 1) << 8) | (u64)(unsigned char)ss[j]);
        }
        __syncwarp();
        // ---- genes in the order of their first member, length filter (:217-219) ----
        int ng = 0;
        for (int base = 0; base < m; base += 32) {
            const int j = base + lane;
            const bool g = j < m && lab[j] == j && (double)(ce[j] - cs[j] + 1) >= a.min_len;
            const u32 bal = __ballot_sync(FULL, g);
            if (g) {
                const long long o = h0 + ng + __popc(bal & lt);
                a.g_start[o] = cs[j];
                a.g_end[o] = ce[j];
                a.g_strand[o] = (int8_t)(cbest[j] & 0xffull);
            }
            ng += __popc(bal);
        }
        if (lane == 0) a.g_count[b] = ng;
        __syncwarp();
    }
}

}  // namespace

#define GCU(x)                                                                                      \
    do {                                                                                            \
        cudaError_t err__ = (x);                                                                    \
        if (err__ != cudaSuccess) {                                                                 \
            fprintf(stderr, "waafle_b200 genecaller: %s: %s\n", #x, cudaGetErrorString(err__));     \
            rc = WFL_ERR_CUDA;                                                                      \
            goto done;                                                                              \
        }                                                                                           \
    } while (0)

extern "C" int wfl_call_genes(int device, int64_t n_blocks, const int64_t *block_off, const int32_t *qstart, const int32_t *qend,
                              const int8_t *strand, const uint8_t *keep, double min_overlap, double min_gene_length,
                              int32_t *gene_start, int32_t *gene_end, int8_t *gene_strand, int32_t *gene_count, float *ms_kernel) {
    if (n_blocks < 0 || !block_off || !gene_count) return WFL_ERR_ARG;
    if (n_blocks == 0) return WFL_OK;
    const int64_t nh = block_off[n_blocks];
    if (nh > 0 && (!qstart || !qend || !strand || !keep || !gene_start || !gene_end || !gene_strand)) return WFL_ERR_ARG;
    int rc = WFL_OK;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return WFL_ERR_CUDA;
    char *pool = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    {
        auto al = [](size_t x) { return (x + 255) & ~size_t(255); };
        const size_t n1 = (size_t)n_blocks + 1, H = (size_t)std::max<int64_t>(nh, 1);
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += al(bytes); return o; };
        const size_t o_boff = take(n1 * 8), o_qs = take(H * 4), o_qe = take(H * 4), o_st = take(H), o_kp = take(H),
                     o_key = take((2 * H + 32 * (size_t)n_blocks) * 8), o_sa = take(H * 4), o_sb = take(H * 4), o_lab = take(H * 4),
                     o_cs = take(H * 4), o_ce = take(H * 4), o_cb = take(H * 8), o_ss = take(H), o_gs = take(H * 4),
                     o_ge = take(H * 4), o_gst = take(H), o_gc = take((size_t)n_blocks * 4), o_wq = take(8);
        GCU(cudaSetDevice(device));
        GCU(cudaMalloc(&pool, off));
        GCU(cudaEventCreate(&e0));
        GCU(cudaEventCreate(&e1));
        GCU(cudaMemcpy(pool + o_boff, block_off, n1 * 8, cudaMemcpyHostToDevice));
        if (nh > 0) {
            GCU(cudaMemcpy(pool + o_qs, qstart, (size_t)nh * 4, cudaMemcpyHostToDevice));
            GCU(cudaMemcpy(pool + o_qe, qend, (size_t)nh * 4, cudaMemcpyHostToDevice));
            GCU(cudaMemcpy(pool + o_st, strand, (size_t)nh, cudaMemcpyHostToDevice));
            GCU(cudaMemcpy(pool + o_kp, keep, (size_t)nh, cudaMemcpyHostToDevice));
        }
        GCU(cudaMemset(pool + o_wq, 0, 8));
        GcArgs a{};
        a.n_blocks = n_blocks;
        a.block_off = reinterpret_cast<const int64_t *>(pool + o_boff);
        a.qstart = reinterpret_cast<const int32_t *>(pool + o_qs);
        a.qend = reinterpret_cast<const int32_t *>(pool + o_qe);
        a.strand = reinterpret_cast<const int8_t *>(pool + o_st);
        a.keep = reinterpret_cast<const uint8_t *>(pool + o_kp);
        a.thr = min_overlap;
        a.min_len = min_gene_length;
        a.key = reinterpret_cast<u64 *>(pool + o_key);
        a.sa = reinterpret_cast<int *>(pool + o_sa);
        a.sb = reinterpret_cast<int *>(pool + o_sb);
        a.lab = reinterpret_cast<int *>(pool + o_lab);
        a.cs = reinterpret_cast<int *>(pool + o_cs);
        a.ce = reinterpret_cast<int *>(pool + o_ce);
        a.cbest = reinterpret_cast<u64 *>(pool + o_cb);
        a.ss = reinterpret_cast<int8_t *>(pool + o_ss);
        a.g_start = reinterpret_cast<int32_t *>(pool + o_gs);
        a.g_end = reinterpret_cast<int32_t *>(pool + o_ge);
        a.g_strand = reinterpret_cast<int8_t *>(pool + o_gst);
        a.g_count = reinterpret_cast<int32_t *>(pool + o_gc);
        a.wq = reinterpret_cast<unsigned long long *>(pool + o_wq);
        cudaDeviceProp prop;
        GCU(cudaGetDeviceProperties(&prop, device));
        const int grid = (int)std::min<int64_t>((n_blocks + GC_WPC - 1) / GC_WPC, (int64_t)prop.multiProcessorCount * 8);
        GCU(cudaEventRecord(e0));
        wfl_genecall<<<grid, 32 * GC_WPC>>>(a);
        GCU(cudaGetLastError());
        GCU(cudaEventRecord(e1));
        GCU(cudaMemcpy(gene_count, pool + o_gc, (size_t)n_blocks * 4, cudaMemcpyDeviceToHost));
        if (nh > 0) {
            GCU(cudaMemcpy(gene_start, pool + o_gs, (size_t)nh * 4, cudaMemcpyDeviceToHost));
            GCU(cudaMemcpy(gene_end, pool + o_ge, (size_t)nh * 4, cudaMemcpyDeviceToHost));
            GCU(cudaMemcpy(gene_strand, pool + o_gst, (size_t)nh, cudaMemcpyDeviceToHost));
        }
        if (ms_kernel) GCU(cudaEventElapsedTime(ms_kernel, e0, e1));
    }
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (pool) cudaFree(pool);
    return rc;
}
