// C ABI of the engine (include/waafle_b200.h): handle, device memory, H2D / launch / D2H.
// Plain CUDA runtime -- no torch types anywhere near the boundary.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "wfl_device.cuh"

using namespace wfl;

namespace {

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct wfl_engine {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr, stream2 = nullptr;
    cudaEvent_t ev_join = nullptr;
    int n_slots = 2;                     // plugin call: sub-batches alternate between two compute streams
    cudaEvent_t ev[8] = {};
    std::vector<cudaEvent_t> chunk_ev;
    size_t chunk_bytes = size_t(96) << 20;   // H2D chunk size of the pipelined plugin call (measured best: profiles/README.md)
    std::string err;
    bool have_params = false, have_tax = false, have_batch = false, have_results = false;
    DevParams P{};
    DevTax tax{};
    DevBatch b{};
    DevOut o{};
    int64_t n = 0, nh = 0, nl = 0;
    int S = 0;
    // knobs
    int threads = 32, smem_bytes = 10 * 1024, ctas_per_sm = 12;
    size_t slab_bytes = 256 * 1024;
    // device buffers (grow-only)
    Buf tx[4], in[12], out[18], slab, ctr, work, scratch, cm[5], dbg[4], plan_index, plan_data, plan_tree;
    int plan_nmax = 0;
    // multi-kernel pipeline state
    int mode = 2;                       // 0: v1 CTA-per-contig, 1: v2 monolithic warp kernel, 2: pipeline
    int tax_max_depth = 0;
    bool use_tree = false;              // WFL_K2=tree: K2 by tree walk with constant-subtree skipping (parity-tested,
                                        // but slower than the flat leaf plan on B200: more local-memory state)
    size_t pipe_pool_bytes = size_t(8192) << 20;
    Buf pipe_pool[2], pipe_ctg, pipe_lists[2], pipe_cnt[2], pipe_wq[2], k2_desc[2], k2_order[2], k2_keys[2], k2_meta[2];
    size_t k2_cap_override = 0;         // WFL_K2_CAP: descriptor capacity per sub-batch (test hook for the overflow path)
    bool k2_global = true;              // WFL_K2=contig: K2 per contig (wfl_pipe_scores) instead of the sorted global group list
    std::vector<int64_t> h_hit_off, h_locus_off;
    std::vector<int64_t> chunks;         // contig boundaries of the sub-batches of the current batch
    bool chunks_streamed = false;        // their H2D copies are in flight on copy_stream (plugin call)
    int resident_split = 1;              // WFL_SPLIT > 1: sub-batches of a resident batch on both compute streams (measured
                                         // slower: kernels of different phases then share the SMs and their instruction caches)
    bool chunk_fixed = false;            // WFL_CHUNK_MB given: use it as is
    double chunk_shrink = 1.0;           // < 1: geometric tail of the streaming schedule, each chunk >= shrink * its
                                         // predecessor (WFL_CHUNK_SHRINK; measured no better than equal chunks)
    wfl_stats stats{};
    int64_t members_total = 0;
};

namespace {

// WFL_TRACE=1: host-side timeline of a plugin call on stderr (milliseconds since the call started)
static bool g_trace = getenv("WFL_TRACE") != nullptr;
static std::chrono::steady_clock::time_point g_t0;
static void trace(const char *what) {
    if (!g_trace) return;
    if (!strcmp(what, "begin")) g_t0 = std::chrono::steady_clock::now();
    fprintf(stderr, "[wfl] %8.3f ms  %s\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_t0).count(), what);
}

bool set_err(wfl_engine *e, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    e->err = buf;
    return false;
}

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t _err = (call);                                                        \
        if (_err != cudaSuccess) {                                                        \
            set_err(e, "%s failed: %s", #call, cudaGetErrorString(_err));                 \
            return WFL_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

int ensure(wfl_engine *e, Buf &b, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return WFL_OK;
    if (b.p) CU(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8;
    CU(cudaMalloc(&b.p, want));
    b.cap = want;
    return WFL_OK;
}

template <class T>
int upload(wfl_engine *e, Buf &b, const T *src, size_t n, const T **dst, cudaStream_t st = nullptr) {
    int rc = ensure(e, b, n * sizeof(T));
    if (rc) return rc;
    if (n) CU(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st ? st : e->stream));
    *dst = static_cast<const T *>(b.p);
    return WFL_OK;
}

template <class T>
int outbuf(wfl_engine *e, Buf &b, size_t n, T **dst) {
    int rc = ensure(e, b, n * sizeof(T));
    if (rc) return rc;
    *dst = static_cast<T *>(b.p);
    return WFL_OK;
}


// Host twin of the kernel's build_plan(): numpy's pairwise split tree over n elements, flattened
// (numpy/_core/src/umath/loops_utils.h.src: leaves <= 128, split at n/2 - (n/2) % 8).
void host_build_plan(int n, std::vector<uint16_t> &data, PlanEntry &pe) {
    pe.off = (uint32_t)data.size();
    int sz[64], ch[64], sp = 0, nset = 0;
    uint8_t k8[5] = {0, 0, 0, 0, 0};
    sz[0] = n;
    ch[0] = 0;
    sp = 1;
    uint32_t cnt = 0;
    while (sp > 0) {
        int m = sz[--sp], c = ch[sp];
        if (m <= 128) {
            data.push_back((uint16_t)(m | (c << 8)));
            ++cnt;
            if (m >= 8 && nset <= 4) {
                uint8_t k = (uint8_t)(m >> 3);
                int j = 0;
                while (j < nset && j < 4 && k8[j] < k) ++j;
                if (j == nset || (j < 4 && k8[j] != k)) {
                    if (nset < 4) {
                        for (int q = nset; q > j; --q) k8[q] = k8[q - 1];
                        k8[j] = k;
                    }
                    ++nset;
                }
            }
        } else {
            int n2 = m / 2;
            n2 -= n2 % 8;
            sz[sp] = m - n2;
            ch[sp++] = c + 1;
            sz[sp] = n2;
            ch[sp++] = 0;
        }
    }
    pe.nleaf = cnt;
    if (nset > 4) k8[0] = k8[1] = k8[2] = k8[3] = 0;
    pe.k8 = (uint32_t)k8[0] | ((uint32_t)k8[1] << 8) | ((uint32_t)k8[2] << 16) | ((uint32_t)k8[3] << 24);
}


// Distinct node sizes of the split tree of an n-element sum, ascending, children by index.
void host_build_tree(int n, std::vector<TreeEntry> &data, PlanEntry &pe) {
    std::vector<int> sizes;
    std::vector<int> todo{n};
    while (!todo.empty()) {
        int m = todo.back();
        todo.pop_back();
        if (std::find(sizes.begin(), sizes.end(), m) != sizes.end()) continue;
        sizes.push_back(m);
        if (m > 128) {
            int n2 = m / 2;
            n2 -= n2 % 8;
            todo.push_back(n2);
            todo.push_back(m - n2);
        }
    }
    std::sort(sizes.begin(), sizes.end());
    pe.toff = (uint32_t)data.size();
    pe.nsz = (uint32_t)sizes.size();
    for (int m : sizes) {
        TreeEntry t;
        t.size = (uint16_t)m;
        t.li = t.ri = 255;
        if (m > 128) {
            int n2 = m / 2;
            n2 -= n2 % 8;
            t.li = (uint8_t)(std::find(sizes.begin(), sizes.end(), n2) - sizes.begin());
            t.ri = (uint8_t)(std::find(sizes.begin(), sizes.end(), m - n2) - sizes.begin());
        }
        data.push_back(t);
    }
}

// Leaf plans for every gene length up to `want` (capped): built once, kept on the device.
int ensure_plan_table(wfl_engine *e, int want) {
    const int cap = 16384;
    want = std::min(std::max(want, 4096), cap);
    if (e->plan_nmax >= want) return WFL_OK;
    std::vector<PlanEntry> index((size_t)want + 1);
    std::vector<uint16_t> data;
    data.reserve((size_t)want * want / 150 + 1024);
    index[0] = PlanEntry{0, 0, 0};
    std::vector<TreeEntry> tdata;
    tdata.reserve((size_t)want * 18);
    index[0].toff = index[0].nsz = 0;
    for (int n = 1; n <= want; ++n) {
        host_build_plan(n, data, index[n]);
        host_build_tree(n, tdata, index[n]);
    }
    const PlanEntry *di;
    const uint16_t *dd;
    int rc;
    if ((rc = upload(e, e->plan_index, index.data(), index.size(), &di))) return rc;
    if ((rc = upload(e, e->plan_data, data.data(), data.size(), &dd))) return rc;
    const TreeEntry *dt;
    if ((rc = upload(e, e->plan_tree, tdata.data(), tdata.size(), &dt))) return rc;
    CU(cudaStreamSynchronize(e->stream));
    e->plan_nmax = want;
    return WFL_OK;
}

int check_batch(wfl_engine *e, const wfl_batch *in) {
    if (!in || in->n_contigs < 0 || in->n_hits < 0 || in->n_loci < 0) {
        set_err(e, "bad batch sizes");
        return WFL_ERR_ARG;
    }
    if (in->n_hits >= (1ll << 31) || in->n_loci >= (1ll << 31) || in->n_contigs >= (1ll << 31)) {
        set_err(e, "batch too large: contigs, hits and loci must each be < 2^31 per batch");
        return WFL_ERR_ARG;
    }
    if (!in->hit_off || !in->locus_off || (in->n_hits && (!in->hit_qstart || !in->hit_qend ||
        !in->hit_taxon || !in->hit_score || !in->hit_scov || !in->hit_strand)) ||
        (in->n_loci && (!in->locus_start || !in->locus_end || !in->locus_strand))) {
        set_err(e, "null array in batch");
        return WFL_ERR_ARG;
    }
    if (e->P.p.n_systems > 0 && in->n_hits && !in->hit_sysmask) {
        set_err(e, "n_systems > 0 but hit_sysmask is null");
        return WFL_ERR_ARG;
    }
    if (in->hit_off[0] != 0 || in->locus_off[0] != 0 || in->hit_off[in->n_contigs] != in->n_hits ||
        in->locus_off[in->n_contigs] != in->n_loci) {
        set_err(e, "CSR offsets do not span the hit / locus arrays");
        return WFL_ERR_ARG;
    }
    for (int64_t c = 0; c < in->n_contigs; ++c)
        if (in->hit_off[c + 1] < in->hit_off[c] || in->locus_off[c + 1] < in->locus_off[c]) {
            set_err(e, "CSR offsets are not monotone at contig %lld", (long long)c);
            return WFL_ERR_ARG;
        }
    return WFL_OK;
}

int alloc_outputs(wfl_engine *e) {
    const size_t n = (size_t)e->n, nl = (size_t)e->nl, S = (size_t)e->S;
    int rc = 0;
    DevOut &o = e->o;
    rc |= outbuf(e, e->out[0], n, &o.call);
    rc |= outbuf(e, e->out[1], n, &o.direction);
    rc |= outbuf(e, e->out[2], n, &o.lifts);
    rc |= outbuf(e, e->out[3], n, &o.clade1);
    rc |= outbuf(e, e->out[4], n, &o.clade2);
    rc |= outbuf(e, e->out[5], n, &o.lca);
    rc |= outbuf(e, e->out[6], n, &o.best1);
    rc |= outbuf(e, e->out[7], n, &o.best2);
    rc |= outbuf(e, e->out[8], n, &o.crit);
    rc |= outbuf(e, e->out[9], n, &o.rank);
    rc |= outbuf(e, e->out[10], nl, &o.synteny);
    rc |= outbuf(e, e->out[11], nl, &o.locus_flags);
    rc |= outbuf(e, e->out[12], nl * S, &o.ann_winner);
    rc |= outbuf(e, e->out[13], n, &o.n_mem_a);
    rc |= outbuf(e, e->out[14], n, &o.n_mem_b);
    rc |= outbuf(e, e->out[15], n, &o.mem_pos);
    rc |= outbuf(e, e->out[17], n, &o.status);
    if (rc) return WFL_ERR_CUDA;
    size_t pool = std::max<size_t>(e->out[16].cap / sizeof(int32_t), std::max<size_t>(2 * n, 1 << 16));
    rc = outbuf(e, e->out[16], pool, &o.mem_pool);
    o.mem_pool_cap = (int64_t)(e->out[16].cap / sizeof(int32_t));
    return rc;
}

int stage_inputs(wfl_engine *e, const wfl_batch *in, bool copy_all);

// Cut the batch into sub-batches of whole contigs.
//  * pipeline mode: a sub-batch's intermediate state must fit the workspace pool;
//  * plugin call (streaming): a sub-batch is also the unit that crosses PCIe on the copy stream while
//    its predecessor is scored.  Byte schedule: a small first chunk (the kernels start early), full
//    chunks, then a geometric tail s_k = shrink * s_{k-1}: the kernels are faster than the link, so as
//    long as a chunk is not much smaller than its predecessor its copy hides the predecessor's kernels,
//    and only the kernels of the LAST (smallest) chunk are exposed behind the end of the H2D stream.
void plan_chunks(wfl_engine *e, const int64_t *hoff, const int64_t *loff, bool streaming) {
    e->chunks.clear();
    e->chunks.push_back(0);
    const int64_t n = e->n;
    if (n <= 0) return;
    const size_t hit_row = 29 + (e->S > 0 ? 4 : 0);
    std::vector<size_t> target;
    if (streaming) {
        // chunk size: 96 MB, but never more than ~12 chunks per call -- every chunk costs one pass through the
        // per-level kernel chain (20 launches x levels, each ending on its slowest contig)
        const size_t total = (size_t)hoff[n] * hit_row;
        const size_t full = e->chunk_fixed ? e->chunk_bytes : std::max(e->chunk_bytes, total / 12 + 1);
        const size_t first = std::min(full / 4, e->chunk_bytes);
        std::vector<size_t> tail;
        size_t acc = 0;
        for (size_t t = std::max<size_t>(full / 6, size_t(1) << 20); e->chunk_shrink < 0.999 && t < full && acc + t + first < total;
             t = (size_t)((double)t / e->chunk_shrink) + 1) {
            tail.push_back(t);
            acc += t;
        }
        size_t rem = total > acc ? total - acc : 0;
        target.push_back(std::min(rem, first));
        rem -= target.back();
        if (rem % full) { target.push_back(rem % full); rem -= rem % full; }   // the odd piece goes early
        for (; rem > 0; rem -= full) target.push_back(full);
        target.insert(target.end(), tail.rbegin(), tail.rend());
    }
    // resident batch (pipeline mode): a few sub-batches alternating between the two compute streams, so that
    // the tail of one sub-batch's kernel chain is filled by the other's kernels
    size_t split_hits = 0;
    if (!streaming && e->mode == 2 && e->n_slots > 1 && e->resident_split > 1 && n >= 8192 * (int64_t)e->resident_split)
        split_hits = (size_t)hoff[n] / (size_t)e->resident_split + 1;
    int64_t c0 = 0;
    size_t k = 0;
    while (c0 < n) {
        int64_t c1 = c0;
        size_t est = 0;
        const size_t goal = k < target.size() ? target[k] : (target.empty() ? e->chunk_bytes : target.back());
        for (;;) {
            // workspace estimate per contig (records + table + level arrays), see wfl_pipeline.cu
            const size_t h = (size_t)(hoff[c1 + 1] - hoff[c1]), g = (size_t)(loff[c1 + 1] - loff[c1]);
            est += 190 * h + 96 * g + 9000;
            ++c1;
            if (c1 >= n) break;
            if (streaming && (size_t)(hoff[c1 + 1] - hoff[c0]) * hit_row > goal) break;
            if (!streaming && split_hits && (size_t)(hoff[c1 + 1] - hoff[c0]) > split_hits) break;
            if (e->mode == 2 && est + 190 * (size_t)(hoff[c1 + 1] - hoff[c1]) + 9000 > e->pipe_pool_bytes) break;
        }
        e->chunks.push_back(c1);
        c0 = c1;
        ++k;
    }
}

// All H2D copies of the plugin call, chunk by chunk on the copy stream, one event per chunk.
int issue_chunk_copies(wfl_engine *e, const wfl_batch *src) {
    const int64_t *hoff = src->hit_off, *loff = src->locus_off;
    cudaStream_t cs = e->copy_stream;
    for (size_t k = 0; k + 1 < e->chunks.size(); ++k) {
        const int64_t c0 = e->chunks[k], c1 = e->chunks[k + 1];
        const size_t h0 = (size_t)hoff[c0], h1 = (size_t)hoff[c1];
        const size_t l0 = (size_t)loff[c0], l1 = (size_t)loff[c1];
#define CP(dst, srcp, lo, hi)                                                                         \
    if ((hi) > (lo))                                                                                  \
    CU(cudaMemcpyAsync(const_cast<void *>(static_cast<const void *>((dst) + (lo))), (srcp) + (lo),    \
                       ((hi) - (lo)) * sizeof(*(srcp)), cudaMemcpyHostToDevice, cs))
        CP(e->b.hit_qstart, src->hit_qstart, h0, h1);
        CP(e->b.hit_qend, src->hit_qend, h0, h1);
        CP(e->b.hit_taxon, src->hit_taxon, h0, h1);
        CP(e->b.hit_score, src->hit_score, h0, h1);
        CP(e->b.hit_scov, src->hit_scov, h0, h1);
        CP(e->b.hit_strand, src->hit_strand, h0, h1);
        if (e->S > 0) CP(e->b.hit_sysmask, src->hit_sysmask, h0, h1);
        CP(e->b.locus_start, src->locus_start, l0, l1);
        CP(e->b.locus_end, src->locus_end, l0, l1);
        CP(e->b.locus_strand, src->locus_strand, l0, l1);
#undef CP
        if (e->chunk_ev.size() <= k) {
            cudaEvent_t evn;
            CU(cudaEventCreateWithFlags(&evn, cudaEventDisableTiming));
            e->chunk_ev.push_back(evn);
        }
        CU(cudaEventRecord(e->chunk_ev[k], cs));
    }
    CU(cudaEventRecord(e->ev[5], cs));   // end of the last H2D chunk
    e->chunks_streamed = true;
    return WFL_OK;
}

// One sub-batch [c0, c1) through the multi-kernel pipeline (wfl_pipeline.cu).
int launch_pipeline_chunk(wfl_engine *e, const ScoreArgs &sa, int64_t c0, int64_t c1, int slot = 0) {
    cudaStream_t stream = slot ? e->stream2 : e->stream;
    const int L = std::min(std::max(e->tax_max_depth + 1 - e->P.p.jump_taxonomy, 1), 64);
    int rc;
    char *pool;
    PipeCtg *ctg;
    int *lists, *cnt;
    unsigned long long *wq;
    const size_t n = (size_t)e->n;
    if ((rc = outbuf(e, e->pipe_pool[slot], std::min<size_t>(e->pipe_pool_bytes, 200 * (size_t)e->nh + 9200 * n + 96 * (size_t)e->nl + (size_t(64) << 20)), &pool))) return rc;
    if ((rc = outbuf(e, e->pipe_ctg, n, &ctg))) return rc;
    if ((rc = outbuf(e, e->pipe_lists[slot], 4 * n + 16, &lists))) return rc;
    if ((rc = outbuf(e, e->pipe_cnt[slot], 3 * 66 + 8, &cnt))) return rc;
    if ((rc = outbuf(e, e->pipe_wq[slot], 6 * 66 + 8, &wq))) return rc;
    CU(cudaMemsetAsync(cnt, 0, (3 * 66 + 8) * sizeof(int), stream));
    CU(cudaMemsetAsync(wq, 0, (6 * 66 + 8) * sizeof(unsigned long long), stream));
    PipeArgs pa{};
    pa.b = sa.b; pa.t = sa.t; pa.o = sa.o; pa.P = sa.P; pa.ctr = sa.ctr;
    pa.pool = pool; pa.pool_used = wq; pa.pool_cap = e->pipe_pool[slot].cap;
    pa.ctg = ctg;
    pa.plan_nmax = sa.plan_nmax; pa.plan_index = sa.plan_index; pa.plan_data = sa.plan_data; pa.plan_tree = sa.plan_tree;
    K2Meta *k2meta = nullptr;
    pa.k2_desc = nullptr;
    if (e->k2_global && !e->use_tree) {
        const size_t cap = e->k2_cap_override ? e->k2_cap_override
                                             : (size_t)(e->h_hit_off[c1] - e->h_hit_off[c0]) * 5 / 4 + 4096;
        if ((rc = outbuf(e, e->k2_desc[slot], cap, &pa.k2_desc))) return rc;
        if ((rc = outbuf(e, e->k2_order[slot], cap, &pa.k2_order))) return rc;
        if ((rc = outbuf(e, e->k2_keys[slot], cap, &pa.k2_keys))) return rc;
        if ((rc = outbuf(e, e->k2_meta[slot], 66, &k2meta))) return rc;
        pa.k2_cap = cap;
        CU(cudaMemsetAsync(k2meta, 0, 66 * sizeof(K2Meta), stream));
    }
    pa.dbg_contig = -1;
    int *list[4] = {lists, lists + n, lists + 2 * n, lists + 3 * n};
    int *cnt_act = cnt, *cnt_two = cnt + 66, *cnt_lift = cnt + 2 * 66;
    const int grid = e->sm_count * pipe_ctas_per_sm();
    pa.wq = wq + 1;
    pa.work_base = c0; pa.n_work = c1 - c0;
    pa.list_act = list[0]; pa.cnt_act = &cnt_act[0];
    launch_pipe_prepare(pa, (int)std::min<int64_t>(grid, c1 - c0), stream);
    for (int lvl = 0; lvl < L; ++lvl) {
        pa.list_act = list[lvl & 1]; pa.cnt_act = &cnt_act[lvl];
        pa.list_next = list[(lvl + 1) & 1]; pa.cnt_next = &cnt_act[lvl + 1];
        pa.list_two = list[2]; pa.cnt_two = &cnt_two[lvl];
        pa.list_lift = list[3]; pa.cnt_lift = &cnt_lift[lvl];
        pa.k2_meta = k2meta ? k2meta + lvl : nullptr;
        pa.wq = wq + 2 + 6 * lvl;
        launch_pipe_regroup(pa, grid, stream);
        pa.wq = wq + 3 + 6 * lvl;
        if (pa.k2_desc != nullptr) {
            launch_pipe_k2sort(pa, grid, stream);
            launch_pipe_k2(pa, grid, stream);
        } else {
            launch_pipe_scores(pa, grid, stream);
        }
        pa.wq = wq + 4 + 6 * lvl;
        launch_pipe_masks(pa, grid, stream);
        pa.wq = wq + 5 + 6 * lvl;
        launch_pipe_one(pa, grid, stream);
        pa.wq = wq + 6 + 6 * lvl;
        launch_pipe_two(pa, grid, stream);
        pa.wq = wq + 7 + 6 * lvl;
        launch_pipe_lift(pa, grid, stream);
    }
    pa.list_act = list[L & 1]; pa.cnt_act = &cnt_act[L];
    launch_pipe_leftover(pa, stream);
    CU(cudaGetLastError());
    e->stats.kernel_launches += 2 + (pa.k2_desc != nullptr ? 9 : 6) * L;
    return WFL_OK;
}


int run_kernels(wfl_engine *e, const wfl_batch *src = nullptr) {
    if (!e->have_params || !e->have_tax || !e->have_batch) {
        set_err(e, "params, taxonomy and batch must be set before running");
        return WFL_ERR_STATE;
    }
    int rc = alloc_outputs(e);
    if (rc) return rc;
    DevCounters *ctr;
    if ((rc = outbuf(e, e->ctr, 1, &ctr))) return rc;
    int64_t *cm_off, *cm_counts, *cm_index, *cm_totals, *scan_tmp;
    int32_t *cm_na;
    if ((rc = outbuf(e, e->cm[0], (size_t)e->n + 1, &cm_off))) return rc;
    if ((rc = outbuf(e, e->cm[1], (size_t)e->n, &cm_na))) return rc;
    if ((rc = outbuf(e, e->cm[2], 8, &cm_counts))) return rc;
    if ((rc = outbuf(e, e->cm[3], (size_t)e->n, &cm_index))) return rc;
    cm_totals = cm_counts + 4;
    if ((rc = outbuf(e, e->scratch, compaction_scratch_elems(e->n), &scan_tmp))) return rc;

    e->stats = wfl_stats{};
    e->stats.contigs = e->n;
    e->stats.hits = e->nh;
    e->stats.loci = e->nl;
    DevCounters hc{};
    DevCounters acc{};
    int grid = e->sm_count * e->ctas_per_sm;
    size_t slab_bytes = e->slab_bytes;
    const int64_t *work_list = nullptr;
    int64_t n_work = e->n;
    std::vector<int64_t> replay;
    bool restart = false;
    CU(cudaEventRecord(e->ev[1], e->stream));
    for (int attempt = 0;; ++attempt) {
        char *slab;
        if ((rc = outbuf(e, e->slab, (size_t)grid * slab_bytes, &slab))) return rc;
        if (attempt == 0 || restart) {
            CU(cudaMemsetAsync(ctr, 0, sizeof(DevCounters), e->stream));
            restart = false;
        } else {
            // replay: keep the member-pool bump pointer and the statistics, reset the work queue
            hc.next_work = 0;
            hc.n_overflow = 0;
            hc.slab_need_max = 0;
            CU(cudaMemcpyAsync(ctr, &hc, sizeof hc, cudaMemcpyHostToDevice, e->stream));
        }
        ScoreArgs a{};
        a.b = e->b;
        a.t = e->tax;
        a.o = e->o;
        a.P = e->P;
        a.ctr = ctr;
        a.work_list = work_list;
        a.n_work = n_work;
        a.slab = slab;
        a.slab_bytes = slab_bytes;
        a.smem_bytes = e->smem_bytes;
        a.plan_nmax = e->plan_nmax;
        a.plan_index = static_cast<const PlanEntry *>(e->plan_index.p);
        a.plan_data = static_cast<const uint16_t *>(e->plan_data.p);
        a.plan_tree = e->use_tree ? static_cast<const TreeEntry *>(e->plan_tree.p) : nullptr;
        a.dbg_contig = -1;
        if (attempt == 0 && e->n > 0 && (src != nullptr || e->mode == 2)) {
            // The batch is cut into chunks of contigs.  Plugin call (src != nullptr): chunk k+1 crosses
            // PCIe on the copy stream while chunk k is scored on the compute stream.  Pipeline mode: a
            // chunk is also the sub-batch whose intermediate state shares the workspace pool.
            if (e->chunks.size() < 2 || e->chunks.back() != e->n)
                plan_chunks(e, e->h_hit_off.data(), e->h_locus_off.data(), false);
            bool used_slot1 = false;
            for (size_t k = 0; k + 1 < e->chunks.size(); ++k) {
                const int64_t c0 = e->chunks[k], c1 = e->chunks[k + 1];
                // plugin call: sub-batches alternate between two compute streams, so the tail of one
                // sub-batch's kernel chain (11+ dependent launches, each ending on its slowest contig)
                // is filled by the next sub-batch's kernels instead of idling the SMs
                const bool streamed = src != nullptr && e->chunks_streamed;
                const int slot = (e->mode == 2 && e->n_slots > 1 && e->chunks.size() > 2) ? (int)(k & 1) : 0;
                if (slot && !used_slot1) {
                    CU(cudaEventRecord(e->ev_join, e->stream));   // orders stream2 after the counter reset above
                    CU(cudaStreamWaitEvent(e->stream2, e->ev_join, 0));
                    used_slot1 = true;
                }
                if (streamed) CU(cudaStreamWaitEvent(slot ? e->stream2 : e->stream, e->chunk_ev[k], 0));
                if (e->mode == 2) {
                    if ((rc = launch_pipeline_chunk(e, a, c0, c1, slot))) return rc;
                } else {
                    CU(cudaMemsetAsync(&ctr->next_work, 0, sizeof(unsigned long long), e->stream));
                    a.work_base = c0;
                    a.n_work = c1 - c0;
                    if (e->mode == 1)
                        launch_score_kernel_warp(a, (int)std::min<int64_t>(grid, a.n_work), e->stream);
                    else
                        launch_score_kernel(a, (int)std::min<int64_t>(grid, a.n_work), e->threads, e->stream);
                    CU(cudaGetLastError());
                    e->stats.kernel_launches++;
                }
            }
            if (used_slot1) {
                CU(cudaEventRecord(e->ev_join, e->stream2));
                CU(cudaStreamWaitEvent(e->stream, e->ev_join, 0));
            }
        } else if (n_work > 0) {
            if (e->mode != 0)
                launch_score_kernel_warp(a, (int)std::min<int64_t>(grid, n_work), e->stream);
            else
                launch_score_kernel(a, (int)std::min<int64_t>(grid, n_work), e->threads, e->stream);
            CU(cudaGetLastError());
            e->stats.kernel_launches++;
        }
        if (attempt == 0) CU(cudaEventRecord(e->ev[2], e->stream));
        trace("kernels launched");
        CU(cudaMemcpyAsync(&hc, ctr, sizeof hc, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        trace("kernels finished");
        acc = hc;
        if (hc.n_badinput) {
            set_err(e, "%llu contig(s) carry a hit taxon index outside the taxonomy", hc.n_badinput);
            return WFL_ERR_ARG;
        }
        if (hc.n_runaway) {
            set_err(e, "Runaway taxonomic recursion in %llu contig(s)", hc.n_runaway);
            return WFL_ERR_RUNAWAY;
        }
        if ((int64_t)hc.mem_pool_used > e->o.mem_pool_cap) {
            // melded-member staging pool too small: grow it and redo the whole batch
            if (attempt > 8) { set_err(e, "member pool keeps overflowing"); return WFL_ERR_CUDA; }
            Buf nb;
            if ((rc = ensure(e, nb, (size_t)hc.mem_pool_used * sizeof(int32_t) * 2))) return rc;
            cudaFree(e->out[16].p);
            e->out[16] = nb;
            e->o.mem_pool = static_cast<int32_t *>(nb.p);
            e->o.mem_pool_cap = (int64_t)(nb.cap / sizeof(int32_t));
            work_list = nullptr;
            n_work = e->n;
            restart = true;
            e->stats.workspace_retries++;
            continue;
        }
        if (hc.n_overflow == 0) break;
        if (attempt > 8) { set_err(e, "workspace keeps overflowing (need %llu bytes)", hc.slab_need_max); return WFL_ERR_CUDA; }
        // replay the contigs that outgrew their workspace with a slab of the size they asked for
        std::vector<uint8_t> st((size_t)e->n);
        CU(cudaMemcpy(st.data(), e->o.status, (size_t)e->n, cudaMemcpyDeviceToHost));
        replay.clear();
        for (int64_t c = 0; c < e->n; ++c)
            if (st[c] == 1) replay.push_back(c);
        e->stats.workspace_retries += (int64_t)replay.size();
        slab_bytes = std::max<size_t>((size_t)hc.slab_need_max, 2 * slab_bytes);
        slab_bytes = (slab_bytes + 255) & ~size_t(255);
        size_t budget = size_t(8) << 30;
        grid = (int)std::max<size_t>(1, std::min<size_t>({(size_t)grid, replay.size(), budget / slab_bytes}));
        int64_t *wl;
        if ((rc = outbuf(e, e->work, replay.size(), &wl))) return rc;
        CU(cudaMemcpyAsync(wl, replay.data(), replay.size() * sizeof(int64_t), cudaMemcpyHostToDevice, e->stream));
        work_list = wl;
        n_work = (int64_t)replay.size();
    }
    CompactArgs ca{};
    ca.n = e->n;
    ca.o = e->o;
    ca.member_off = cm_off;
    ca.n_members_a = cm_na;
    ca.call_counts = cm_counts;
    ca.call_index = cm_index;
    ca.scan_tmp = scan_tmp;
    ca.totals = cm_totals;
    // members: compact into a buffer as large as the staging pool
    int32_t *cm_members;
    if ((rc = outbuf(e, e->cm[4], (size_t)e->o.mem_pool_cap, &cm_members))) return rc;
    ca.members = cm_members;
    ca.members_cap = e->o.mem_pool_cap;
    e->stats.kernel_launches += launch_compaction(ca, e->stream);
    CU(cudaGetLastError());
    CU(cudaEventRecord(e->ev[3], e->stream));
    int64_t totals[4];
    CU(cudaMemcpyAsync(totals, cm_totals, sizeof totals, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->members_total = totals[0];
    e->stats.matched_pairs = (int64_t)acc.matched_pairs;
    e->stats.groups = (int64_t)acc.groups;
    e->stats.levels = (int64_t)acc.levels;
    e->stats.pairs_tested = (int64_t)acc.pairs_tested;
    e->stats.pairs_scored = (int64_t)acc.pairs_scored;
    e->stats.smem_contigs = (int64_t)acc.smem_contigs;
    for (int q = 0; q < 12; ++q) e->stats.phase_cycles[q] = (int64_t)acc.phase_cycles[q];
    CU(cudaEventElapsedTime(&e->stats.ms_score_kernel, e->ev[1], e->ev[2]));
    CU(cudaEventElapsedTime(&e->stats.ms_kernels, e->ev[1], e->ev[3]));
    e->have_results = true;
    return WFL_OK;
}

int stage_inputs(wfl_engine *e, const wfl_batch *in, bool copy_all) {
    if (!e->have_params || !e->have_tax) {
        set_err(e, "set params and taxonomy before uploading a batch");
        return WFL_ERR_STATE;
    }
    int rc = check_batch(e, in);
    if (rc) return rc;
    e->have_batch = e->have_results = false;
    e->n = in->n_contigs;
    e->nh = in->n_hits;
    e->nl = in->n_loci;
    e->S = e->P.p.n_systems;
    DevBatch &b = e->b;
    b.n_contigs = e->n;
    b.n_hits = e->nh;
    b.n_loci = e->nl;
    const size_t n1 = (size_t)e->n + 1, nh = (size_t)e->nh, nl = (size_t)e->nl;
    e->chunks.clear();
    e->chunks_streamed = false;
    // The CSR offsets go first.  Plugin call: on the COPY stream, ahead of chunk 0 (two streams feeding the
    // same DMA queue are not ordered by issue time: on the compute stream they ended up behind the whole
    // bulk transfer and the first kernel with them); chunk 0's event covers them.
    cudaStream_t os = copy_all ? e->stream : e->copy_stream;
    CU(cudaEventRecord(e->ev[0], os));
    if ((rc = upload(e, e->in[0], in->hit_off, n1, &b.hit_off, os))) return rc;
    if ((rc = upload(e, e->in[1], in->locus_off, n1, &b.locus_off, os))) return rc;
    if (!copy_all) {
        // plugin call: device arrays only, and the chunked H2D copies start NOW on the copy stream -- the
        // host-side preparation below (plan table, offsets) and the kernels overlap with the transfer
        rc = 0;
        rc |= outbuf(e, e->in[2], nh, const_cast<int32_t **>(&b.hit_qstart));
        rc |= outbuf(e, e->in[3], nh, const_cast<int32_t **>(&b.hit_qend));
        rc |= outbuf(e, e->in[4], nh, const_cast<int32_t **>(&b.hit_taxon));
        rc |= outbuf(e, e->in[5], nh, const_cast<double **>(&b.hit_score));
        rc |= outbuf(e, e->in[6], nh, const_cast<double **>(&b.hit_scov));
        rc |= outbuf(e, e->in[7], nh, const_cast<int8_t **>(&b.hit_strand));
        b.hit_sysmask = nullptr;
        if (e->S > 0) rc |= outbuf(e, e->in[8], nh, const_cast<uint32_t **>(&b.hit_sysmask));
        rc |= outbuf(e, e->in[9], nl, const_cast<int32_t **>(&b.locus_start));
        rc |= outbuf(e, e->in[10], nl, const_cast<int32_t **>(&b.locus_end));
        rc |= outbuf(e, e->in[11], nl, const_cast<int8_t **>(&b.locus_strand));
        if (rc) return WFL_ERR_CUDA;
        if (e->n > 0) {
            plan_chunks(e, in->hit_off, in->locus_off, true);
            trace("chunks planned");
            if ((rc = issue_chunk_copies(e, in))) return rc;
            trace("chunk copies issued");
        }
    }
    {
        int maxlen = 0;
        for (int64_t i = 0; i < in->n_loci; ++i) {
            int d = in->locus_end[i] - in->locus_start[i];
            d = (d < 0 ? -d : d) + 1;
            maxlen = d > maxlen ? d : maxlen;
        }
        if ((rc = ensure_plan_table(e, maxlen))) return rc;
    }
    e->h_hit_off.assign(in->hit_off, in->hit_off + n1);
    e->h_locus_off.assign(in->locus_off, in->locus_off + n1);
    if (copy_all) {
        if ((rc = upload(e, e->in[2], in->hit_qstart, nh, &b.hit_qstart))) return rc;
        if ((rc = upload(e, e->in[3], in->hit_qend, nh, &b.hit_qend))) return rc;
        if ((rc = upload(e, e->in[4], in->hit_taxon, nh, &b.hit_taxon))) return rc;
        if ((rc = upload(e, e->in[5], in->hit_score, nh, &b.hit_score))) return rc;
        if ((rc = upload(e, e->in[6], in->hit_scov, nh, &b.hit_scov))) return rc;
        if ((rc = upload(e, e->in[7], in->hit_strand, nh, &b.hit_strand))) return rc;
        b.hit_sysmask = nullptr;
        if (e->S > 0 && (rc = upload(e, e->in[8], in->hit_sysmask, nh, &b.hit_sysmask))) return rc;
        if ((rc = upload(e, e->in[9], in->locus_start, nl, &b.locus_start))) return rc;
        if ((rc = upload(e, e->in[10], in->locus_end, nl, &b.locus_end))) return rc;
        if ((rc = upload(e, e->in[11], in->locus_strand, nl, &b.locus_strand))) return rc;
    }
    trace("host prep done");
    if (copy_all) {
        CU(cudaEventRecord(e->ev[1], e->stream));
        CU(cudaStreamSynchronize(e->stream));
        CU(cudaEventElapsedTime(&e->stats.ms_h2d, e->ev[0], e->ev[1]));
    } else if (e->n == 0) {
        CU(cudaStreamSynchronize(e->copy_stream));   // no chunk event will order the (empty) offsets
    }
    e->have_batch = true;
    return WFL_OK;
}

template <class T>
int d2h(wfl_engine *e, T *dst, const T *src, size_t n) {
    if (!dst || n == 0) return WFL_OK;
    CU(cudaMemcpyAsync(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost, e->stream));
    return WFL_OK;
}

int download(wfl_engine *e, wfl_results *out) {
    if (!e->have_results) {
        set_err(e, "no results to download");
        return WFL_ERR_STATE;
    }
    if (!out) { set_err(e, "null results"); return WFL_ERR_ARG; }
    out->members_used = e->members_total;
    if (out->members && out->members_capacity < e->members_total) {
        set_err(e, "members buffer too small: need %lld", (long long)e->members_total);
        return WFL_ERR_CAPACITY;
    }
    const size_t n = (size_t)e->n, nl = (size_t)e->nl, S = (size_t)e->S;
    const DevOut &o = e->o;
    int rc = 0;
    float h2d = e->stats.ms_h2d;
    CU(cudaEventRecord(e->ev[4], e->stream));
    rc |= d2h(e, out->call, o.call, n);
    rc |= d2h(e, out->direction, o.direction, n);
    rc |= d2h(e, out->lifts, o.lifts, n);
    rc |= d2h(e, out->clade1, o.clade1, n);
    rc |= d2h(e, out->clade2, o.clade2, n);
    rc |= d2h(e, out->lca, o.lca, n);
    rc |= d2h(e, out->best1, o.best1, n);
    rc |= d2h(e, out->best2, o.best2, n);
    rc |= d2h(e, out->crit, o.crit, n);
    rc |= d2h(e, out->rank, o.rank, n);
    rc |= d2h(e, out->synteny, o.synteny, nl);
    rc |= d2h(e, out->locus_flags, o.locus_flags, nl);
    rc |= d2h(e, out->ann_winner, o.ann_winner, nl * S);
    rc |= d2h(e, out->member_off, static_cast<const int64_t *>(e->cm[0].p), n + 1);
    rc |= d2h(e, out->n_members_a, static_cast<const int32_t *>(e->cm[1].p), n);
    rc |= d2h(e, out->members, static_cast<const int32_t *>(e->cm[4].p), (size_t)e->members_total);
    rc |= d2h(e, out->call_counts, static_cast<const int64_t *>(e->cm[2].p), 3);
    rc |= d2h(e, out->call_index, static_cast<const int64_t *>(e->cm[3].p), n);
    if (rc) return WFL_ERR_CUDA;
    CU(cudaEventRecord(e->ev[6], e->stream));
    CU(cudaStreamSynchronize(e->stream));
    CU(cudaEventElapsedTime(&e->stats.ms_d2h, e->ev[4], e->ev[6]));
    e->stats.ms_h2d = h2d;
    return WFL_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// exported entry points
// ---------------------------------------------------------------------------------------------

extern "C" {

int wfl_abi_version(void) { return WFL_ABI_VERSION; }

int wfl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int wfl_create(int device, wfl_engine **out) {
    if (!out) return WFL_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return WFL_ERR_CUDA;
    wfl_engine *e = new wfl_engine();
    e->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) != cudaSuccess) {
        delete e;
        return WFL_ERR_CUDA;
    }
    e->sm_count = prop.multiProcessorCount;
    if (const char *k = getenv("WFL_KERNEL")) {
        std::string m(k);
        e->mode = m == "v1" ? 0 : m == "v2" ? 1 : 2;
    }
    if (e->mode == 0) { e->threads = 128; e->smem_bytes = 36 * 1024; e->ctas_per_sm = 6; }
    if (const char *k = getenv("WFL_K2")) { e->use_tree = std::string(k) == "tree"; e->k2_global = std::string(k) != "contig"; }
    if (const char *k = getenv("WFL_K2_CAP")) e->k2_cap_override = (size_t)std::max<long long>(0, atoll(k));
    if (const char *k = getenv("WFL_POOL_MB")) e->pipe_pool_bytes = (size_t)atoll(k) << 20;
    if (const char *k = getenv("WFL_CHUNK_MB")) { e->chunk_bytes = (size_t)atoll(k) << 20; e->chunk_fixed = true; }
    if (const char *k = getenv("WFL_SPLIT")) e->resident_split = std::max(1, atoi(k));
    if (const char *k = getenv("WFL_STREAMS")) e->n_slots = atoi(k) >= 2 ? 2 : 1;
    if (const char *k = getenv("WFL_CHUNK_SHRINK")) e->chunk_shrink = std::min(1.0, std::max(0.05, atof(k)));
    e->smem_optin = prop.sharedMemPerBlockOptin;
    for (auto &ev : e->ev)
        if (cudaEventCreate(&ev) != cudaSuccess) {
            delete e;
            return WFL_ERR_CUDA;
        }
    *out = e;
    return WFL_OK;
}

void wfl_destroy(wfl_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    auto fr = [](Buf &b) { if (b.p) cudaFree(b.p); b.p = nullptr; };
    for (auto &b : e->tx) fr(b);
    for (auto &b : e->in) fr(b);
    for (auto &b : e->out) fr(b);
    for (auto &b : e->cm) fr(b);
    for (auto &b : e->dbg) fr(b);
    fr(e->slab); fr(e->ctr); fr(e->work); fr(e->scratch); fr(e->plan_index); fr(e->plan_data); fr(e->plan_tree);
    for (int q = 0; q < 2; ++q) { fr(e->pipe_pool[q]); fr(e->pipe_lists[q]); fr(e->pipe_cnt[q]); fr(e->pipe_wq[q]); }
    fr(e->pipe_ctg);
    for (int q = 0; q < 2; ++q) { fr(e->k2_desc[q]); fr(e->k2_order[q]); fr(e->k2_keys[q]); fr(e->k2_meta[q]); }
    for (auto &ev : e->ev) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->chunk_ev) cudaEventDestroy(ev);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->stream2) cudaStreamDestroy(e->stream2);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

const char *wfl_last_error(const wfl_engine *e) { return e ? e->err.c_str() : "null engine"; }

int wfl_set_params(wfl_engine *e, const wfl_params *p) {
    if (!e || !p) return WFL_ERR_ARG;
    if (p->n_systems < 0 || p->n_systems > WFL_MAX_SYSTEMS || p->disambiguate_one < 0 || p->disambiguate_one > 1 ||
        p->disambiguate_two < 0 || p->disambiguate_two > 2 || p->weak_loci < 0 || p->weak_loci > 2 ||
        p->ambiguous_threshold < 0 || p->ambiguous_threshold > 2 || p->sister_penalty < 0 ||
        p->sister_penalty > 2 || p->annotation_threshold < 0 || p->annotation_threshold > 2 ||
        p->jump_taxonomy < 0) {
        set_err(e, "parameter out of range");
        return WFL_ERR_ARG;
    }
    DevParams &P = e->P;
    P.p = *p;
    P.min_thr = std::min(p->k1, p->k2);   // waafle_orgscorer.py:338-339
    P.max_thr = std::max(p->k1, p->k2);
    const bool k1_is_min = p->k1 <= p->k2;
    const int sel_min = k1_is_min ? 0 : 1, sel_max = k1_is_min ? 1 : 0;
    const double tri[3] = {1e-6, P.min_thr, P.max_thr};   // off / lenient / strict
    P.ann_thr = tri[p->annotation_threshold];             // :341-346
    P.k_amb = tri[p->ambiguous_threshold];                // :515-516
    P.amb_sel = p->ambiguous_threshold == 0 ? 2 : (p->ambiguous_threshold == 1 ? sel_min : sel_max);
    P.sister_thr = p->sister_penalty == 1 ? P.max_thr : P.min_thr;   // :720-721
    P.sis_sel = p->sister_penalty == 1 ? sel_max : sel_min;
    e->have_params = true;
    e->have_batch = e->have_results = false;
    return WFL_OK;
}

int wfl_set_taxonomy(wfl_engine *e, int32_t n_nodes, const int32_t *parent, const int32_t *depth,
                     const int32_t *leaf_count, const uint8_t *listed, int32_t root_idx, int32_t unknown_idx) {
    if (!e) return WFL_ERR_ARG;
    if (n_nodes <= 0 || !parent || !depth || !leaf_count || !listed || root_idx < 0 || root_idx >= n_nodes ||
        unknown_idx < 0 || unknown_idx >= n_nodes) {
        set_err(e, "bad taxonomy arguments");
        return WFL_ERR_ARG;
    }
    if (parent[root_idx] != root_idx || depth[root_idx] != 0) {
        set_err(e, "root must be its own parent at depth 0");
        return WFL_ERR_ARG;
    }
    for (int32_t i = 0; i < n_nodes; ++i) {
        int32_t p = parent[i];
        if (p < 0 || p >= n_nodes || (i != root_idx && depth[i] != depth[p] + 1)) {
            set_err(e, "taxonomy node %d: parent / depth inconsistent", i);
            return WFL_ERR_ARG;
        }
    }
    CU(cudaSetDevice(e->device));
    int rc;
    if ((rc = upload(e, e->tx[0], parent, (size_t)n_nodes, &e->tax.parent))) return rc;
    if ((rc = upload(e, e->tx[1], depth, (size_t)n_nodes, &e->tax.depth))) return rc;
    if ((rc = upload(e, e->tx[2], leaf_count, (size_t)n_nodes, &e->tax.leaf_count))) return rc;
    if ((rc = upload(e, e->tx[3], listed, (size_t)n_nodes, &e->tax.listed))) return rc;
    CU(cudaStreamSynchronize(e->stream));
    e->tax_max_depth = 0;
    for (int32_t i = 0; i < n_nodes; ++i) e->tax_max_depth = std::max(e->tax_max_depth, (int)depth[i]);
    e->tax.n_nodes = n_nodes;
    e->tax.root = root_idx;
    e->tax.unknown = unknown_idx;
    e->have_tax = true;
    e->have_batch = e->have_results = false;
    return WFL_OK;
}

int wfl_upload_batch(wfl_engine *e, const wfl_batch *in) {
    if (!e) return WFL_ERR_ARG;
    CU(cudaSetDevice(e->device));
    return stage_inputs(e, in, true);
}

int wfl_run_resident(wfl_engine *e) {
    if (!e) return WFL_ERR_ARG;
    CU(cudaSetDevice(e->device));
    float h2d = e->stats.ms_h2d;
    int rc = run_kernels(e);
    e->stats.ms_h2d = h2d;
    return rc;
}

int wfl_download_results(wfl_engine *e, wfl_results *out) {
    if (!e) return WFL_ERR_ARG;
    CU(cudaSetDevice(e->device));
    return download(e, out);
}

int wfl_score_batch(wfl_engine *e, const wfl_batch *in, wfl_results *out) {
    if (!e) return WFL_ERR_ARG;
    CU(cudaSetDevice(e->device));
    trace("begin");
    int rc = stage_inputs(e, in, false);
    if (rc) return rc;
    trace("inputs staged, copies in flight");
    if ((rc = run_kernels(e, in))) return rc;
    trace("kernels + compaction done");
    float h2d = 0.f;
    if (e->n > 0 && cudaEventElapsedTime(&h2d, e->ev[0], e->ev[5]) == cudaSuccess) e->stats.ms_h2d = h2d;
    rc = download(e, out);
    trace("results downloaded");
    return rc;
}

int wfl_host_alloc(size_t bytes, void **out) {
    if (!out) return WFL_ERR_ARG;
    *out = nullptr;
    if (cudaHostAlloc(out, bytes ? bytes : 16, cudaHostAllocDefault) != cudaSuccess) return WFL_ERR_CUDA;
    return WFL_OK;
}

void wfl_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int wfl_get_stats(const wfl_engine *e, wfl_stats *out) {
    if (!e || !out) return WFL_ERR_ARG;
    *out = e->stats;
    return WFL_OK;
}

int wfl_configure(wfl_engine *e, int threads, int smem_bytes, int ctas_per_sm) {
    if (!e) return WFL_ERR_ARG;
    if (threads) {
        if (threads < 32 || threads > 1024 || threads % 32) { set_err(e, "threads must be a multiple of 32 in [32,1024]"); return WFL_ERR_ARG; }
        e->threads = threads;
    }
    if (smem_bytes) {
        if (smem_bytes < 0 || (size_t)smem_bytes + 2048 > e->smem_optin) { set_err(e, "smem_bytes too large"); return WFL_ERR_ARG; }
        e->smem_bytes = smem_bytes & ~15;
    }
    if (ctas_per_sm) {
        if (ctas_per_sm < 1 || ctas_per_sm > 32) { set_err(e, "ctas_per_sm out of range"); return WFL_ERR_ARG; }
        e->ctas_per_sm = ctas_per_sm;
    }
    return WFL_OK;
}

int64_t wfl_debug_gene_scores(wfl_engine *e, int64_t contig, int32_t *clade, int32_t *locus, double *score,
                              int64_t capacity) {
    if (!e) return WFL_ERR_ARG;
    if (!e->have_batch) { set_err(e, "no resident batch"); return WFL_ERR_STATE; }
    if (contig < 0 || contig >= e->n || capacity < 0) { set_err(e, "bad contig index"); return WFL_ERR_ARG; }
    CU(cudaSetDevice(e->device));
    int rc = alloc_outputs(e);
    if (rc) return rc;
    DevCounters *ctr;
    int32_t *dc, *dl;
    double *ds;
    long long *dn;
    int64_t *wl;
    char *slab;
    size_t cap = (size_t)std::max<int64_t>(capacity, 1);
    if ((rc = outbuf(e, e->ctr, 1, &ctr)) || (rc = outbuf(e, e->dbg[0], cap, &dc)) ||
        (rc = outbuf(e, e->dbg[1], cap, &dl)) || (rc = outbuf(e, e->dbg[2], cap, &ds)) ||
        (rc = outbuf(e, e->dbg[3], 1, &dn)) || (rc = outbuf(e, e->work, 1, &wl)))
        return rc;
    size_t slab_bytes = size_t(256) << 20;   // one CTA, generous
    if ((rc = outbuf(e, e->slab, slab_bytes, &slab))) return rc;
    CU(cudaMemsetAsync(ctr, 0, sizeof(DevCounters), e->stream));
    CU(cudaMemsetAsync(dn, 0, sizeof(long long), e->stream));
    CU(cudaMemcpyAsync(wl, &contig, sizeof contig, cudaMemcpyHostToDevice, e->stream));
    ScoreArgs a{};
    a.b = e->b; a.t = e->tax; a.o = e->o; a.P = e->P; a.ctr = ctr;
    a.work_list = wl; a.n_work = 1; a.slab = slab; a.slab_bytes = slab_bytes; a.smem_bytes = e->smem_bytes;
    a.plan_nmax = e->plan_nmax; a.plan_index = static_cast<const PlanEntry *>(e->plan_index.p);
    a.plan_data = static_cast<const uint16_t *>(e->plan_data.p);
    a.plan_tree = e->use_tree ? static_cast<const TreeEntry *>(e->plan_tree.p) : nullptr;
    a.dbg_contig = contig; a.dbg_clade = dc; a.dbg_locus = dl; a.dbg_score = ds; a.dbg_cap = capacity; a.dbg_count = dn;
    if (e->mode != 0)
        launch_score_kernel_warp(a, 1, e->stream);
    else
        launch_score_kernel(a, 1, e->threads, e->stream);
    CU(cudaGetLastError());
    long long cnt = 0;
    CU(cudaMemcpyAsync(&cnt, dn, sizeof cnt, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    size_t m = (size_t)std::min<long long>(cnt, capacity);
    if (m) {
        CU(cudaMemcpy(clade, dc, m * sizeof(int32_t), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(locus, dl, m * sizeof(int32_t), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(score, ds, m * sizeof(double), cudaMemcpyDeviceToHost));
    }
    e->have_results = false;
    return cnt;
}

}  // extern "C"
