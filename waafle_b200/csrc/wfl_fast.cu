// The fused fast path of the orgscorer engine: ONE kernel, one warp per contig, every piece of per-contig
// state in shared memory (a per-warp slice of the CTA's dynamic shared memory), all taxonomy levels in a
// data-dependent loop.  Replaces, for contigs with <= 32 retained loci, the pipeline's prepare / regroup /
// K2 sort / K2 / masks / one / two / lift launch chain and its global-memory workspace round trips.
//
// Reference path (waafle/waafle_orgscorer.py): attach_hits :359-392, update_gene_scores :394-429,
// raise_taxonomy :431-445, evaluate_contig :566-583, explain_one/two :585-619, meld_one/two :621-676,
// LGT checks :678-744.
//
// How it works (and differs from the exact pipeline, wfl_pipeline.cu):
//  * once per contig: the hit spans are staged in shared memory, one pass per locus appends the attached hits
//    (records: hit index + python slice, locus-major), and every locus' records are put in DESCENDING SCORE order
//    (skipped when they already are: the front end's packer delivers hits sorted by score);
//  * per taxonomy level, per locus: the records are streamed in score order; each record's clade (the level's row of
//    the ancestor table: a lift re-bins the site arrays by parent, waafle_orgscorer.py:431-445) gets a dense handle
//    from a shared-memory hash, and the clade's running ENVELOPE STATE for this locus -- the union of its better hits
//    as one interval and sum v * (newly covered sites) -- is updated in place.  That sum / n is the closed form of
//    np.mean of the site array (:371-382, :399-406); no grouping sort, no per-group scan.  A hit that leaves a gap
//    marks the group for a generic endpoint sweep;
//  * numpy's pairwise rounding is not reproduced, so the value is within ~1e-14 of np.mean; instead every DECISION
//    is protected by a guard band:
//      - a gene score within 1e-12 of a threshold it is compared with (k1, k2, 1e-6) is recomputed exactly, in
//        numpy's pairwise order, by the pipeline's own group_mean();
//      - a rank within 1e-12 of the best rank (arg-max) or of best - range (meld set) cannot be settled locally:
//        the contig is handed to the exact pipeline;
//    so calls / clades / synteny are bit-exact and crit / rank agree to <= 1e-12 (north_star tolerance);
//  * gene-score rows are per-clade linked lists of (locus, score) in locus order; every tie-break the reference
//    resolves by name order uses the node index explicitly (handles are arbitrary labels).
// Anything the slice cannot hold (too many loci / records / clades / groups / pairs) is retried by a second launch
// with a larger slice; long genes, > 32 loci, --min-overlap <= 0, guard-band trips and malformed input go to the
// exact pipeline.
#include "wfl_warp_common.cuh"

namespace wfl {

namespace {

constexpr int FAST_WPC = 4;     // warps (= contigs in flight) per CTA
constexpr int EMPTY_KEY = -1;
constexpr int GMAX = 32;        // retained loci per contig: gene bitmasks are one 32-bit word
constexpr int NONE16 = 0xffff;
#ifndef WFL_FAST_CPSM
#define WFL_FAST_CPSM 4         // resident CTAs per SM the kernel is compiled for (register budget)
#endif

#define SM(T, off) (reinterpret_cast<T *>(slice + (off)))

__device__ __forceinline__ u32 hash32(int key) {
    u32 x = (u32)key * 2654435761u;
    return x ^ (x >> 15);
}

// hit x locus test (waafle_orgscorer.py:365-367, utils.py:487-500) for OVERLAPPING intervals and min_overlap > 0.
// The quotient is only formed when the comparison is within 2^-40 of the threshold (fl is monotone, so outside
// that band num/den >= mo is decided by num vs mo*den).
__device__ __forceinline__ bool overlap_ok(int hmin, int hmax, int lmin, int lmax, int llen, double mo) {
    const double num = (double)(min(hmax, lmax) - max(hmin, lmin) + 1);
    const double den = (double)min(hmax - hmin + 1, llen);
    const double p = mo * den;
    if (num > p * (1.0 + 9.1e-13)) return true;
    if (num < p * (1.0 - 9.1e-13)) return false;
    return num / den >= mo;
}

// Per-level view of a contig for the search routines (all pointers into the warp's slice).
struct FLevel {
    int G, T, t_unk;         // t_unk: handle of the spiked Unknown (dense row unk_row), -1 if none
    u32 um;                  // non-ignored loci
    int nun;
    const u16 *cl_head;      // first group of each clade; groups of a clade are linked in locus order
    const u16 *g_next;
    const u8 *g_loc;
    const double *g_score;
    const double *unk_row;
    const int *cl_id;
    const u32 *mk[3];        // per clade gene bitmasks: score >= k1 / k2 / 1e-6
    const int *l_len;
    const int *cl_par;       // two-clade search with sister penalty: parent of every LISTED clade of the level, else -1
};

// gene-score row of one clade, walked in locus order (0 where the clade has no entry, waafle_orgscorer.py:404-405)
struct RowWalk {
    int p;
    bool dense;
    __device__ __forceinline__ void open(const FLevel &L, int t) {
        dense = t == L.t_unk;
        p = dense ? NONE16 : (int)L.cl_head[t];
    }
    __device__ __forceinline__ double at(const FLevel &L, int i) {
        if (dense) return L.unk_row[i];
#pragma unroll 1
        while (p != NONE16 && (int)L.g_loc[p] < i) p = L.g_next[p];
        return (p != NONE16 && (int)L.g_loc[p] == i) ? L.g_score[p] : 0.0;
    }
};

// Contig.score (waafle_orgscorer.py:447-461) over the non-ignored loci: crit = min, rank = np.mean (n <= 32 values:
// numpy sums n < 8 sequentially, otherwise eight strided accumulators, the fixed tree, then the tail).
__device__ __noinline__ double row_stats(const FLevel &L, int t1, int t2, double *crit_out) {
    const int n = L.nun;
    double r[8];
    double crit = __longlong_as_double(0x7ff0000000000000ll), res = 0.0;
    u32 bits = L.um;
    const int body = n < 8 ? 0 : (n & ~7);
    int idx = 0;
    RowWalk w1, w2;
    w1.open(L, t1);
    w2.open(L, t2 >= 0 ? t2 : t1);
#pragma unroll 1
    while (bits) {
        const int i = __ffs(bits) - 1;
        bits &= bits - 1;
        double v = w1.at(L, i);
        if (t2 >= 0) v = fmax(v, w2.at(L, i));
        crit = fmin(crit, v);
        if (idx < body) {
            const int j = idx & 7;
            // r[j] (+)= v without dynamic register indexing
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q == j) r[q] = idx < 8 ? v : r[q] + v;
            if (idx == body - 1) res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        } else {
            res += v;
        }
        ++idx;
    }
    *crit_out = crit;
    return res / (double)n;
}

struct FTwoEval {
    bool swap, dir, ok;
    int c1, c2, t1, t2;   // post-swap clade node ids / handles
    u32 A, B, amb;        // post-swap letters
};

// set_synteny_two + apply_lgt_checks for one option (waafle_orgscorer.py:511-545, 678-744); ta is the clade with the
// smaller node index (clade1 < clade2, :608).
__device__ __noinline__ void eval_two_fast(const FLevel &L, const DevTax &tax, const DevParams &P, int ta, int tb, FTwoEval &ev) {
    const u32 *mamb = L.mk[P.amb_sel], *msis = L.mk[P.sis_sel];
    const bool unk = L.cl_id[ta] == tax.unknown || L.cl_id[tb] == tax.unknown;
    const u32 amb = unk ? 0u : (mamb[ta] & mamb[tb] & L.um);
    u32 A = L.mk[1][ta] & L.um & ~amb;
    u32 B = L.mk[1][tb] & L.um & ~amb & ~A;
    const u32 ab = A | B;
    const bool swap = ab != 0u && (B & (ab & (~ab + 1u))) != 0u;   // "^[^A]*B": first clear letter is B
    if (swap) { const u32 t = A; A = B; B = t; }
    ev.swap = swap;
    ev.t1 = swap ? tb : ta;
    ev.t2 = swap ? ta : tb;
    ev.c1 = L.cl_id[ev.t1];
    ev.c2 = L.cl_id[ev.t2];
    ev.A = A; ev.B = B; ev.amb = amb;
    long long total_len = 0, amb_len = 0;
    int state = 0;
    u32 bits = L.um;
#pragma unroll 1
    while (bits) {   // lengths and "^A+B+A+$" over the non-ignored letters
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const u32 m = 1u << b;
        const int len = L.l_len[b];
        if (A & m) {
            total_len += len;
            state = (state == 0 || state == 1) ? 1 : (state == 2 || state == 3) ? 3 : -1;
        } else if (B & m) {
            total_len += len;
            state = (state == 1 || state == 2) ? 2 : -1;
        } else {
            if (amb & m) { total_len += len; amb_len += len; }
            state = -1;
        }
    }
    ev.dir = state == 3;
    bool ok = true;
    if (total_len > 0 && (double)amb_len / (double)total_len > P.p.ambiguous_fraction) ok = false;
    if (P.p.clade_genes >= 0 && min(__popc(A), __popc(B)) < P.p.clade_genes) ok = false;
    if (P.p.clade_leaves >= 0) {
        const int lc = ev.dir ? tax.leaf_count[ev.c2] : min(tax.leaf_count[ev.c1], tax.leaf_count[ev.c2]);
        if (lc < P.p.clade_leaves) ok = false;
    }
    if (P.p.sister_penalty != 0 && ok) {
        const int p1 = tax.parent[ev.c1], p2 = tax.parent[ev.c2];
#pragma unroll 1
        for (int t = 0; t < L.T && ok; ++t) {
            if (t == ev.t1 || t == ev.t2) continue;
            const int px = L.cl_par[t];   // get_sisters works on the taxonomy file's rows
            const bool s1 = px == p1, s2 = (px == p2) && !ev.dir;
            if (!s1 && !s2) continue;
            const u32 ms = msis[t];
            // a B locus is penalised by clade1's sisters, an A locus by clade2's
            if ((s1 && (ms & B)) || (s2 && (ms & A))) ok = false;
        }
    }
    ev.ok = ok;
}

// Dense handle of a clade in the level's shared-memory hash (get-or-insert).  Handles are arbitrary labels (insertion
// order of a race); nothing downstream depends on their order.  All lanes call it (inactive lanes pass active = false).
__device__ __forceinline__ int clade_handle(int *hkey, u16 *hval, int *cl_id, u32 *mk0, u32 *mk1, u32 *mk2, u16 *cl_head,
                                            u16 *cl_tail, int *t_count, int cmask, int Tcap, u32 i0, u32 i1, u32 i2, int key,
                                            bool active, bool &ovf) {
    u32 slot = hash32(key) & (u32)cmask;
    bool fail = false;
    if (active) {
        int probe = 0;
#pragma unroll 1
        for (; probe <= cmask; ++probe) {
            const int old = atomicCAS(&hkey[slot], EMPTY_KEY, key);
            if (old == EMPTY_KEY) {   // inserted: new clade of this level
                const int hnew = atomicAdd(t_count, 1);
                if (hnew < Tcap) {
                    cl_id[hnew] = key;
                    mk0[hnew] = i0; mk1[hnew] = i1; mk2[hnew] = i2;
                    cl_head[hnew] = cl_tail[hnew] = (u16)NONE16;
                    hval[slot] = (u16)hnew;
                } else {
                    hval[slot] = 0;
                    fail = true;
                }
                break;
            }
            if (old == key) break;
            slot = (slot + 1) & (u32)cmask;
        }
        if (probe > cmask) fail = true;
        if (fail) ovf = true;
    }
    __syncwarp();
    return (active && !fail) ? (int)hval[slot] : 0;
}

// Generic envelope integral of one (clade, locus) group whose better hits leave a gap: the group's records are the records
// [r0, r0 + k) of the locus whose clade of this level is `clade`; they are gathered into the warp's global scratch (slices and
// scores) and swept over their distinct endpoints.  Also serves the exact recomputation (numpy pairwise order) of a score
// that falls inside the guard band.  Cold.
template <class H>
__device__ __noinline__ double group_slow(const FastArgs &a, const H &hits, char *scratch, int Kcap, const u16 *rec_hit,
                                          const u32 *rec_ab, long long h0, int r0, int k, const int *anc, int clade, int n,
                                          bool exact) {
    int *ra = reinterpret_cast<int *>(scratch), *rb = ra + Kcap;
    double *rv = reinterpret_cast<double *>(rb + Kcap);
    int m = 0;
#pragma unroll 1
    for (int j = 0; j < k && m < Kcap; ++j) {
        const long long h = h0 + rec_hit[r0 + j];
        if (anc[hits.taxon(h)] != clade) continue;
        const u32 ab = rec_ab[r0 + j];
        const double sc = a.b.hit_score[h];
        ra[m] = (int)(ab & 0xffffu);
        rb[m] = (int)(ab >> 16);
        rv[m] = sc > 0.0 ? sc : 0.0;
        ++m;
    }
    if (exact) {
        const PlanEntry pe = a.plan_index[n];
        return group_mean(ra, rb, rv, 0, m, n, true, pe.k8, a.plan_data + pe.off, (int)pe.nleaf);   // descending scores
    }
    double sum = 0.0;
#pragma unroll 1
    for (int e2 = 0; e2 < 2 * m; ++e2) {
        const int e = (e2 & 1) ? rb[e2 >> 1] : ra[e2 >> 1];
        bool dup = false;
        int nx = 0x7fffffff;
        double mx = 0.0;
#pragma unroll 1
        for (int q = 0; q < m; ++q) {
            const int qa = ra[q], qb = rb[q];
            if (qa <= e && e < qb) mx = fmax(mx, rv[q]);
            if (qa > e) nx = min(nx, qa);
            if (qb > e) nx = min(nx, qb);
            dup |= (qa == e && 2 * q < e2) || (qb == e && 2 * q + 1 < e2);
        }
        if (!dup && mx > 0.0 && nx != 0x7fffffff) sum += mx * (double)(nx - e);
    }
    return sum / (double)n;
}

// ascending node index order for the melded member lists (the exact pipeline emits them in handle == name order)
__device__ __forceinline__ void emit_members_sorted(const FastArgs &a, const int *list, int n, long long dst, int lane) {
#pragma unroll 1
    for (int j = lane; j < n; j += 32) {
        const int id = list[j];
        int pos = 0;
#pragma unroll 1
        for (int q = 0; q < n; ++q) pos += list[q] < id;
        if (dst + pos < a.o.mem_pool_cap) a.o.mem_pool[dst + pos] = id;
    }
}

struct FOut {
    int call, dir, c1, c2, lca, b1, b2, na, nb, lifts;
    long long mem;
    double crit, rank;
};

__device__ __forceinline__ void write_result_fast(const FastArgs &a, long long c, const FOut &r) {
    a.o.call[c] = (uint8_t)r.call;
    a.o.direction[c] = (uint8_t)r.dir;
    a.o.lifts[c] = r.lifts;
    a.o.clade1[c] = r.c1;
    a.o.clade2[c] = r.c2;
    a.o.lca[c] = r.lca;
    a.o.best1[c] = r.b1;
    a.o.best2[c] = r.b2;
    a.o.crit[c] = r.crit;
    a.o.rank[c] = r.rank;
    a.o.n_mem_a[c] = r.na;
    a.o.n_mem_b[c] = r.nb;
    a.o.mem_pos[c] = r.mem;
    a.o.status[c] = 0;
}

// hit columns in the two wire formats (include/waafle_b200.h: wfl_batch / wfl_packed_batch)
// flags: bit 0 = scov_modified >= --min-scov (waafle_orgscorer.py:362), bit 1 = minus strand
template <bool PACKED>
struct Hits;
template <>
struct Hits<false> {
    const DevBatch &b;
    double min_scov;
    __device__ __forceinline__ void span(long long h, int &q1, int &q2) const { q1 = b.hit_qstart[h]; q2 = b.hit_qend[h]; }
    __device__ __forceinline__ u8 flags(long long h) const {
        return (u8)((b.hit_scov[h] >= min_scov ? 1 : 0) | (b.hit_strand[h] == '-' ? 2 : 0));
    }
    __device__ __forceinline__ int taxon(long long h) const { return b.hit_taxon[h]; }
    __device__ __forceinline__ u32 sysmask(long long h) const { return b.hit_sysmask[h]; }
};
template <>
struct Hits<true> {
    const DevBatch &b;
    double min_scov;
    __device__ __forceinline__ void span(long long h, int &q1, int &q2) const { q1 = b.hit_qstart16[h]; q2 = b.hit_qend16[h]; }
    __device__ __forceinline__ u8 flags(long long h) const {   // the host applied the scov filter while packing
        const u32 w = b.hit_tax16[h];
        return (u8)(((w >> 15) & 1u) | (((w >> 14) & 1u) << 1));
    }
    __device__ __forceinline__ int taxon(long long h) const { return (int)(b.hit_tax16[h] & 0x3fffu); }
    __device__ __forceinline__ u32 sysmask(long long h) const { return b.hit_sysmask8[h]; }
};

}  // namespace

template <bool PACKED>
__global__ void __launch_bounds__(32 * FAST_WPC, WFL_FAST_CPSM) wfl_fast_contigs(const FastArgs a) {
    extern __shared__ __align__(16) char fast_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    char *slice = fast_smem + (size_t)wid * a.cfg.slice_bytes;
    const DevParams &P = a.P;
    const DevTax &tax = a.t;
    const FastCfg &F = a.cfg;
    const int S = P.p.n_systems;
    const bool spike = P.p.weak_loci == 2;
    const Hits<PACKED> hits{a.b, P.p.min_scov};
    const double GUARD = a.guard;
    const double thr3[3] = {P.p.k1, P.p.k2, 1e-6};

    int *l_lo = SM(int, F.o_llo), *l_len = SM(int, F.o_llen), *l_raw = SM(int, F.o_lraw);
    signed char *l_str = SM(signed char, F.o_lstr);
    u16 *l_base = SM(u16, F.o_lbase);
    double *maxv = SM(double, F.o_maxv), *unk_row = SM(double, F.o_unk);
    // level-invariant records (hit x locus matches), locus-major, descending score inside a locus: hit index and python
    // slice [a, b) packed a | b << 16
    u16 *rec_hit = SM(u16, F.o_rhit);
    u32 *rec_ab = SM(u32, F.o_rab);
    char *xs = SM(char, F.o_x);           // scratch: rank cache / two-clade candidates + survivors
    // staging of the contig's hit spans / flags while the records are bucketed (aliases the per-level arrays)
    u32 *hsp = SM(u32, F.o_hkey);
    u32 *hmk = hsp + F.Hcap;              // loci each hit is attached to
    // per-level arrays: clade hash and table, (clade, locus) groups
    int *hkey = SM(int, F.o_hkey);
    u16 *hval = SM(u16, F.o_hval);
    int *cl_id = SM(int, F.o_clid);
    u32 *mk0 = SM(u32, F.o_mk0), *mk1 = SM(u32, F.o_mk1), *mk2 = SM(u32, F.o_mk2);
    u16 *cl_head = SM(u16, F.o_clhead), *cl_tail = SM(u16, F.o_cltail);
    double *g_score = SM(double, F.o_gscore);   // running sum v * (newly covered sites) while the locus streams, then the score
    u32 *g_u = SM(u32, F.o_gu);                 // [Kcap] union of the better hits of the CURRENT locus' groups: ua | ub << 16
    u16 *g_t = SM(u16, F.o_gt), *g_next = SM(u16, F.o_gnext);
    u8 *g_loc = SM(u8, F.o_gloc);               // locus of the group (bit 7: a hit left a gap -> generic sweep)
    unsigned long long *stat = SM(unsigned long long, F.o_stat);   // per-warp counters, flushed once at exit
    int *t_count = reinterpret_cast<int *>(stat + 8);              // clades of the current level
    char *scratch = a.scratch + ((size_t)blockIdx.x * FAST_WPC + wid) * (size_t)F.Kcap * 16;
    enum { ST_PAIRS, ST_GROUPS, ST_LEVELS, ST_PTEST, ST_PSCORE, ST_DONE, ST_REFINED, ST_TRIPS, ST_N };
    if (lane < ST_N) stat[lane] = 0;
    __syncwarp();
    const unsigned long long n_work = a.n_work_dev ? *a.n_work_dev : (unsigned long long)a.n_work;

#pragma unroll 1
    for (;;) {
        long long c = -1;
        if (lane == 0) {
            const unsigned long long w = atomicAdd(a.wq, 1ull);
            if (w < n_work) c = a.work_list ? (long long)a.work_list[w] : a.work_base + (long long)w;
        }
        c = __shfl_sync(FULL, c, 0);
        if (c < 0) break;
        const long long h0 = a.b.hit_off[c], l0 = a.b.locus_off[c];
        const int H = (int)(a.b.hit_off[c + 1] - h0), Graw = (int)(a.b.locus_off[c + 1] - l0);
        bool fallback = false, trip = false;
        int reason = 0;   // why the contig goes to the next pass: 0 loci, 1 hits, 2 coordinates, 3 records, 4 clades, 5 groups, 6 pairs, 7 guard
        FOut R{WFL_CALL_UNCLASSIFIED, 0, -1, -1, -1, -1, -1, 0, 0, H > 0 ? P.p.jump_taxonomy : 0, 0, 0.0, 0.0};
        bool finished = false;

        // ---- loci: --min-gene-length filter, GFF order kept (attach_loci, waafle_orgscorer.py:348-357) ----
        int G = 0;
#pragma unroll 1
        for (int base = 0; base < Graw; base += 32) {
            const int j = base + lane;
            int flag = 0, lo = 0, len = 0;
            signed char ls = 0;
            if (j < Graw) {
                const int s = a.b.locus_start[l0 + j], e = a.b.locus_end[l0 + j];
                lo = min(s, e);
                len = max(s, e) - lo + 1;
                flag = (double)len >= P.p.min_gene_length;
                ls = a.b.locus_strand[l0 + j];
                a.o.locus_flags[l0 + j] = flag ? WFL_LOCUS_RETAINED : 0;
                a.o.synteny[l0 + j] = 0;
#pragma unroll 1
                for (int s2 = 0; s2 < S; ++s2) a.o.ann_winner[(l0 + j) * S + s2] = -1;
            }
            const u32 m = __ballot_sync(FULL, flag);
            if (flag) {
                const int pos = G + __popc(m & lt_mask());
                if (pos < GMAX) {
                    l_lo[pos] = lo;
                    l_len[pos] = len;
                    l_raw[pos] = j;
                    l_str[pos] = ls;
                }
                if (len > 65535 || len > a.plan_nmax) fallback = true;   // slices are packed in 16 bits
            }
            G += __popc(m);
        }
        if (G > GMAX) fallback = true;
        fallback = __any_sync(FULL, fallback);
        __syncwarp();
        const u32 allG = G >= 32 ? 0xffffffffu : ((1u << G) - 1u);
        // a locus without an entry scores 0 (waafle_orgscorer.py:404-405): its mask bit is (0 >= threshold)
        const u32 init0 = thr3[0] <= 0.0 ? allG : 0u, init1 = thr3[1] <= 0.0 ? allG : 0u, init2 = 0u;

        // ---- K1, once per contig (level-invariant): which hit is attached to which locus (attach_hits :359-369) ----
        int M = 0;
        if (!fallback && H > 0 && G > 0) {
            if (H > F.Hcap) {
                fallback = true;
                reason = 1;
            } else {
                // (a) one pass over the hits: span, filter flags and the mask of the loci each hit is attached to
                bool big = false;
                const u8 fmask = P.p.stranded ? 3 : 1;
#pragma unroll 2
                for (int base = 0; base < H; base += 32) {
                    const int h = base + lane;
                    if (h < H) {
                        int q1, q2;
                        hits.span(h0 + h, q1, q2);
                        const int hmin = min(q1, q2), hmax = max(q1, q2);
                        big |= hmin < 0 || hmax > 65535;
                        const u8 fl = hits.flags(h0 + h);
                        u32 mb = 0;
                        if (fl & 1) {   // scov_modified >= --min-scov (:362)
#pragma unroll 1
                            for (int i = 0; i < G; ++i) {
                                const int lmin = l_lo[i], llen = l_len[i], lmax = lmin + llen - 1;
                                if (lmin > hmax || hmin > lmax) continue;
                                if (P.p.stranded) {   // hit.sstrand == locus.strand (:365)
                                    const signed char ls = l_str[i];
                                    if (!((ls == '-' && (fl & 2)) || (ls == '+' && !(fl & 2)))) continue;
                                }
                                if (overlap_ok(hmin, hmax, lmin, lmax, llen, P.p.min_overlap)) mb |= 1u << i;
                            }
                        }
                        hsp[h] = (u32)(hmin & 0xffff) | ((u32)(hmax & 0xffff) << 16);
                        hmk[h] = mb;
                    }
                }
                (void)fmask;
                if (__any_sync(FULL, big)) { fallback = true; reason = 2; }   // spans are staged in 16 bits
                __syncwarp();
            }
            // (b) one compaction pass per locus over the staged masks: records come out locus-major, hit order inside
#pragma unroll 1
            for (int i = 0; i < G && !fallback; ++i) {
                const int lmin = l_lo[i], llen = l_len[i];
                if (lane == 0) l_base[i] = (u16)M;
                const int m0 = M;
#pragma unroll 2
                for (int base = 0; base < H; base += 32) {
                    const int h = base + lane;
                    const bool mt = h < H && ((hmk[h] >> i) & 1u);
                    const u32 m = __ballot_sync(FULL, mt);
                    if (mt) {
                        const int slot = M + __popc(m & lt_mask());
                        if (slot < F.Mcap) {
                            const u32 sp = hsp[h];
                            const int hmin = (int)(sp & 0xffffu), hmax = (int)(sp >> 16);
                            rec_hit[slot] = (u16)h;
                            // python slice [h1 : h2+1] of the site array (:373-382)
                            rec_ab[slot] = (u32)max(0, hmin - lmin) | ((u32)(min(llen - 1, hmax - lmin) + 1) << 16);
                        }
                    }
                    M += __popc(m);
                }
                if (M > F.Mcap || M - m0 > F.Kcap) { fallback = true; reason = 3; }
            }
            if (lane == 0) l_base[G] = (u16)M;
            __syncwarp();
            // (c) every locus' records in DESCENDING SCORE order (ties: hit order), once per contig: the level loop streams
            // them in that order, so that a clade's envelope only ever grows by "newly covered sites x score".  Loci that
            // arrive sorted (the front end's packer sorts the hits of a contig by score) skip the rank-by-counting sort.
            if (!fallback) {
                double *tv = reinterpret_cast<double *>(hsp);                      // [Kcap] scores of the locus
                u16 *th = reinterpret_cast<u16 *>(tv + F.Kcap);                    // [Kcap] sorted hit indices
                u32 *tab = reinterpret_cast<u32 *>(th + F.Kcap + (F.Kcap & 1));    // [Kcap] sorted slices
#pragma unroll 1
                for (int i = 0; i < G; ++i) {
                    const int r0 = l_base[i], k = (int)l_base[i + 1] - r0;
                    if (k < 2) continue;
#pragma unroll 2
                    for (int j = lane; j < k; j += 32) tv[j] = a.b.hit_score[h0 + rec_hit[r0 + j]];
                    __syncwarp();
                    bool sorted = true;
#pragma unroll 1
                    for (int j = lane; j + 1 < k; j += 32) sorted &= tv[j] >= tv[j + 1];
                    if (__all_sync(FULL, sorted)) continue;
#pragma unroll 1
                    for (int j = lane; j < k; j += 32) {
                        const double v = tv[j];
                        int rk = 0;
#pragma unroll 4
                        for (int q = 0; q < k; ++q) {
                            const double vq = tv[q];
                            rk += (vq > v) || (vq == v && q < j);
                        }
                        th[rk] = rec_hit[r0 + j];
                        tab[rk] = rec_ab[r0 + j];
                    }
                    __syncwarp();
#pragma unroll 1
                    for (int j = lane; j < k; j += 32) {
                        rec_hit[r0 + j] = th[j];
                        rec_ab[r0 + j] = tab[j];
                    }
                    __syncwarp();
                }
            }
        }

        int iter = 0;
#pragma unroll 1
        while (!fallback && !finished && H > 0 && G > 0) {
            // ============================ one taxonomy level ============================
            const int *anc = a.anc + (size_t)min(R.lifts, a.anc_rows - 1) * (size_t)tax.n_nodes;
            int N = 0, ovf_reason = 4;
            bool ovf = false;
            if (lane == 0) *t_count = 0;
#pragma unroll 1
            for (int s = lane; s <= F.cmask; s += 32) hkey[s] = EMPTY_KEY;
            __syncwarp();
            if (spike)   // "Unknown" is a clade of every level (waafle_orgscorer.py:416-418): handle 0
                (void)clade_handle(hkey, hval, cl_id, mk0, mk1, mk2, cl_head, cl_tail, t_count, F.cmask, F.Tcap, 0u, 0u, 0u,
                                   tax.unknown, lane == 0, ovf);
            const int t_unk = spike ? 0 : -1;
            if (lane == 0) ++stat[ST_LEVELS];

#pragma unroll 1
            for (int i = 0; i < G; ++i) {
                const int llen = l_len[i];
                const int r0 = l_base[i], k = (int)l_base[i + 1] - r0;
                const int N0 = N;
                if (iter == 0 && lane == 0) stat[ST_PAIRS] += (unsigned long long)k;
                if (k == 0) { if (lane == 0) maxv[i] = 0.0; continue; }

                // ---- K3: annotation winners, level-independent (score_hit :384-392): last hit with the max score ----
                if (S > 0 && iter == 0) {
#pragma unroll 1
                    for (int s2 = 0; s2 < S; ++s2) {
                        u64 bb = 0;
                        long long bw = -1;
#pragma unroll 1
                        for (int r = lane; r < k; r += 32) {
                            const int hh = rec_hit[r0 + r];
                            const double sc = a.b.hit_score[h0 + hh];
                            if (((hits.sysmask(h0 + hh) >> s2) & 1u) && sc >= P.ann_thr) {
                                const u64 sb = dbits(sc);
                                if (sb > bb || (sb == bb && hh > bw)) { bb = sb; bw = hh; }
                            }
                        }
                        const u64 mx = warp_max_u64(bb);
                        const long long w = warp_max_ll((mx != 0 && bb == mx) ? bw : -1);
                        if (lane == 0) a.o.ann_winner[(l0 + l_raw[i]) * S + s2] = w >= 0 ? (int)(h0 + w) : -1;
                    }
                }

                // ---- K2: stream the locus' records in score order; every clade's envelope state grows in place ----
                // (the next tile's hit data is requested before the current tile is processed)
                u32 ab_n = 0;
                int tx_n = 0;
                double sc_n = 0.0;
                if (lane < k) {
                    const long long h = h0 + rec_hit[r0 + lane];
                    ab_n = rec_ab[r0 + lane];
                    tx_n = hits.taxon(h);
                    sc_n = a.b.hit_score[h];
                }
#pragma unroll 1
                for (int base = 0; base < k; base += 32) {
                    const int j = base + lane;
                    bool act = j < k;
                    const u32 ab = ab_n;
                    const int tx = tx_n;
                    const double sc = sc_n;
                    if (j + 32 < k) {
                        const long long h = h0 + rec_hit[r0 + j + 32];
                        ab_n = rec_ab[r0 + j + 32];
                        tx_n = hits.taxon(h);
                        sc_n = a.b.hit_score[h];
                    }
                    double v = 0.0;
                    int cl = tax.root;
                    if (act) {
                        if ((u32)tx >= (u32)tax.n_nodes) ovf = true;   // malformed input: the exact pipeline reports it
                        else cl = R.lifts ? anc[tx] : tx;              // row 0 of the ancestor table is the identity
                        v = sc > 0.0 ? sc : 0.0;                       // sites start at 0 (np.zeros, :381)
                    }
                    const int t = clade_handle(hkey, hval, cl_id, mk0, mk1, mk2, cl_head, cl_tail, t_count, F.cmask, F.Tcap,
                                               init0, init1, init2, cl, act, ovf);
                    // with "assign-unknown" the hits of a taxon NAMED Unknown carry no gene score: the spiked row replaces
                    // gene_scores["Unknown"] (waafle_orgscorer.py:416-418)
                    if (t == t_unk) act = false;
                    const u32 peers = __match_any_sync(FULL, act ? t : 0x10000 + lane);
                    const int leader = __ffs(peers) - 1;
                    int g = -1;
                    bool cont = false;
                    if (act && lane == leader) {   // the clade's group of THIS locus, if an earlier tile opened it
                        const int tail = cl_tail[t];
                        if (tail != NONE16 && (int)(g_loc[tail] & 0x3f) == i) { g = tail; cont = true; }
                    }
                    const bool newg = act && lane == leader && !cont;
                    const u32 nm = __ballot_sync(FULL, newg);
                    if (newg) {
                        g = N + __popc(nm & lt_mask());
                        if (g < F.Ncap) {
                            g_t[g] = (u16)t;
                            g_loc[g] = (u8)i;
                            g_next[g] = (u16)NONE16;
                            const int tail = cl_tail[t];   // rows: groups of a clade linked in locus order
                            if (tail != NONE16) g_next[tail] = (u16)g; else cl_head[t] = (u16)g;
                            cl_tail[t] = (u16)g;
                        }
                    }
                    N += __popc(nm);
                    if (act && N <= F.Ncap) {
                        g = __shfl_sync(peers, g, leader);
                        cont = __shfl_sync(peers, (int)cont, leader) != 0;
                        int ua, ub;
                        double sum;
                        bool cplx = false;
                        u32 rest = peers;
                        if (cont) {
                            const u32 u = g_u[g - N0];
                            ua = (int)(u & 0xffffu);
                            ub = (int)(u >> 16);
                            sum = g_score[g];
                        } else {   // the group opens with its best hit
                            const u32 ab0 = __shfl_sync(peers, ab, leader);
                            const double v0 = __shfl_sync(peers, v, leader);
                            ua = (int)(ab0 & 0xffffu);
                            ub = (int)(ab0 >> 16);
                            sum = v0 * (double)(ub - ua);
                            rest &= rest - 1;
                        }
#pragma unroll 1
                        while (rest) {   // the clade's other hits of this tile, in score order
                            if (ua == 0 && ub == llen) break;   // the union covers the gene: nothing can be added
                            const int p = __ffs(rest) - 1;
                            rest &= rest - 1;
                            const u32 abp = __shfl_sync(peers, ab, p);
                            const double vp = __shfl_sync(peers, v, p);
                            if (vp > 0.0) {
                                const int pa = (int)(abp & 0xffffu), pb = (int)(abp >> 16);
                                if (pa > ub || pb < ua) {
                                    cplx = true;   // a gap: the union is no longer one interval
                                } else {
                                    const int add = max(0, ua - pa) + max(0, pb - ub);
                                    if (add) sum += vp * (double)add;
                                    ua = min(ua, pa);
                                    ub = max(ub, pb);
                                }
                            }
                        }
                        if (lane == leader) {
                            g_u[g - N0] = (u32)ua | ((u32)ub << 16);
                            g_score[g] = sum;
                            if (cplx) g_loc[g] |= 0x80;
                        }
                    }
                    // the next tile reads this tile's state: clade_handle() synchronises the warp first
                }
                if (N > F.Ncap) { ovf = true; ovf_reason = 5; }
                ovf = __any_sync(FULL, ovf);
                if (ovf) break;
                __syncwarp();

                // ---- the locus' groups are complete: gene scores, guard band, masks, per-locus max ----
                u64 mxb = dbits(0.0);
                const double dn = (double)llen;
#pragma unroll 1
                for (int base = N0; base < N; base += 32) {
                    const int g = base + lane;
                    const bool act = g < N;
                    int t = 0;
                    bool cplx = false;
                    double sc = 0.0;
                    if (act) {
                        t = g_t[g];
                        cplx = (g_loc[g] & 0x80) != 0;
                        sc = g_score[g] / dn;
                    }
                    // groups whose hits left a gap: generic endpoint sweep; scores inside the guard band of a threshold:
                    // recomputed in numpy's summation order.  Both cold, one lane at a time (global scratch of the warp).
                    u32 cm = __ballot_sync(FULL, act && cplx);
#pragma unroll 1
                    while (cm) {
                        const int ln = __ffs(cm) - 1;
                        cm &= cm - 1;
                        if (lane == ln) sc = group_slow(a, hits, scratch, F.Kcap, rec_hit, rec_ab, h0, r0, k, anc, cl_id[t], llen, false);
                        __syncwarp();
                    }
                    const bool near = act && (fabs(sc - thr3[0]) <= GUARD || fabs(sc - thr3[1]) <= GUARD ||
                                              fabs(sc - thr3[2]) <= GUARD);
                    u32 nm = __ballot_sync(FULL, near);
#pragma unroll 1
                    while (nm) {
                        const int ln = __ffs(nm) - 1;
                        nm &= nm - 1;
                        if (lane == ln) {
                            sc = group_slow(a, hits, scratch, F.Kcap, rec_hit, rec_ab, h0, r0, k, anc, cl_id[t], llen, true);
                            atomicAdd(&stat[ST_REFINED], 1ull);
                        }
                        __syncwarp();
                    }
                    if (act) {
                        g_score[g] = sc;
                        g_loc[g] = (u8)i;
                        const u32 bit = 1u << i;
                        mk0[t] = sc >= thr3[0] ? (mk0[t] | bit) : (mk0[t] & ~bit);
                        mk1[t] = sc >= thr3[1] ? (mk1[t] | bit) : (mk1[t] & ~bit);
                        mk2[t] = sc >= thr3[2] ? (mk2[t] | bit) : (mk2[t] & ~bit);
                        if (cl_id[t] != tax.unknown) {   // waafle_orgscorer.py:409-411
                            const u64 sb = dbits(sc);
                            mxb = sb > mxb ? sb : mxb;
                        }
                    }
                }
                mxb = warp_max_u64(mxb);
                if (lane == 0) maxv[i] = dbits_inv(mxb);
                __syncwarp();
            }
            ovf = __any_sync(FULL, ovf);
            if (ovf) { fallback = true; reason = ovf_reason; break; }
            if (lane == 0) stat[ST_GROUPS] += (unsigned long long)N;
            __syncwarp();
            const int T = *t_count;

            // ---- K4: weak loci (update_gene_scores :407-429) ----
            bool ign = false;
            double mx = 0.0;
            if (lane < G) {
                mx = maxv[lane];
                ign = P.p.weak_loci == 0 ? !(mx >= P.min_thr) : false;
            }
            const u32 um = __ballot_sync(FULL, lane < G && !ign);
            const int nun = __popc(um);
            if (lane < G) a.o.locus_flags[l0 + l_raw[lane]] = WFL_LOCUS_RETAINED | (ign ? WFL_LOCUS_IGNORED : 0);
            if (spike) {
                // gene_scores["Unknown"] = 1 - maxes (:416-418): a dense row
                double u = 0.0;
                bool near = false;
                if (lane < G) {
                    u = 1.0 - mx;
                    unk_row[lane] = u;
                    near = fabs(u - thr3[0]) <= GUARD || fabs(u - thr3[1]) <= GUARD || fabs(u - thr3[2]) <= GUARD;
                }
                if (__any_sync(FULL, near)) { trip = true; break; }
                const u32 m0 = __ballot_sync(FULL, lane < G && u >= thr3[0]);
                const u32 m1 = __ballot_sync(FULL, lane < G && u >= thr3[1]);
                const u32 m2 = __ballot_sync(FULL, lane < G && u >= thr3[2]);
                if (lane == 0) { mk0[0] = m0; mk1[0] = m1; mk2[0] = m2; }
            }
            __syncwarp();
            if (iter == 0 && nun == 0) { finished = true; break; }   // "empty" contig (:959): unclassified

            int hasroot = 0;
#pragma unroll 1
            for (int t = lane; t < T; t += 32) hasroot |= cl_id[t] == tax.root;
            hasroot = __any_sync(FULL, hasroot);

            FLevel L{G, T, t_unk, um, nun, cl_head, g_next, g_loc, g_score, unk_row, cl_id, {mk0, mk1, mk2}, l_len, hkey};

            // ---- K6: one-clade search (explain_one :585-597) ----
            {
                // ranks of the passing clades are kept for the meld pass (in the scratch area, if they fit)
                double *rcache = reinterpret_cast<double *>(xs);
                const bool cached = T * 8 <= F.x_bytes;
                u64 bbits = 0;
                int bid = -1, btl = -1;
                double bcrit = 0.0;
#pragma unroll 1
                for (int t = lane; t < T; t += 32) {
                    if ((mk0[t] & um) != um) continue;   // crit >= k1
                    double crit;
                    const double rank = row_stats(L, t, -1, &crit);
                    if (cached) rcache[t] = rank;
                    const u64 b = dbits(rank);
                    if (b > bbits || (b == bbits && cl_id[t] > bid)) { bbits = b; bid = cl_id[t]; btl = t; bcrit = crit; }
                }
                const u64 wb = warp_max_u64(bbits);
                // ties: last in name order == largest node index (meld_one :623-624, canonical order)
                const long long wid2 = warp_max_ll((wb != 0 && bbits == wb) ? (long long)bid : -1);
                if (wb != 0 && wid2 >= 0) {
                    const int owner = __ffs(__ballot_sync(FULL, bbits == wb && (long long)bid == wid2)) - 1;
                    const int tb = __shfl_sync(FULL, btl, owner);
                    const double brank = dbits_inv(wb);
                    const double bcr = __shfl_sync(FULL, bcrit, owner);
                    // meld_one (:621-631): options within --range of the best; guard the arg-max and the range edge
                    int my = -1, nk = 0;
                    bool near = false;
                    int *klist = hkey;   // kept clades (node ids); the hash is dead after the loci loop
                    int kbase = 0;
#pragma unroll 1
                    for (int base = 0; base < T; base += 32) {
                        const int t = base + lane;
                        bool kept = false;
                        if (t < T && (mk0[t] & um) == um) {
                            double crit;
                            const double rank = t == tb ? brank : (cached ? rcache[t] : row_stats(L, t, -1, &crit));
                            const double d = brank - rank;
                            if (t != tb && fabs(d) <= GUARD) near = true;
                            if (P.p.disambiguate_one == 1 && fabs(d - P.p.range) <= GUARD) near = true;
                            kept = P.p.disambiguate_one == 1 && d <= P.p.range;
                        }
                        const u32 m = __ballot_sync(FULL, kept);
                        if (kept) {
                            my = lca2(tax, my, cl_id[t]);
                            klist[kbase + __popc(m & lt_mask())] = cl_id[t];
                            ++nk;
                        }
                        kbase += __popc(m);
                    }
                    if (__any_sync(FULL, near)) { trip = true; break; }
                    R.call = WFL_CALL_NO_LGT;
                    R.b1 = R.c1 = cl_id[tb];
                    R.crit = bcr;
                    R.rank = brank;
                    if (P.p.disambiguate_one == 1) {
                        R.c1 = warp_lca(tax, my);
                        R.na = warp_sum(nk);
                    }
                    if (lane < G)   // set_synteny_one (:495-509)
                        a.o.synteny[l0 + l_raw[lane]] = ign ? '~' : ((mk0[tb] >> lane) & 1u ? 'A' : '!');
                    __syncwarp();
                    if (R.na > 0) {
                        long long mb = 0;
                        if (lane == 0) mb = (long long)atomicAdd(&a.ctr->mem_pool_used, (unsigned long long)R.na);
                        R.mem = __shfl_sync(FULL, mb, 0);
                        emit_members_sorted(a, klist, R.na, R.mem, lane);
                    }
                    finished = true;
                    break;
                }
            }

            // ---- K7 / K8: two-clade search (explain_two :599-619, meld_two :633-669, LGT checks :678-744) ----
            {
                u16 *cand = reinterpret_cast<u16 *>(xs);
                int T2 = 0;
#pragma unroll 1
                for (int base = 0; base < T; base += 32) {
                    const int t = base + lane;
                    const bool f = t < T && mk1[t] != 0u;   // max(gene_scores[clade]) >= k2, unmasked (:603-605)
                    const u32 m = __ballot_sync(FULL, f);
                    if (f) cand[T2 + __popc(m & lt_mask())] = (u16)t;
                    T2 += __popc(m);
                }
                __syncwarp();
                const int NP = T2 * (T2 - 1) / 2;
                if (lane == 0) stat[ST_PTEST] += (unsigned long long)NP;
                // survivors of the mask prefilter follow the candidates in the scratch area
                u16 *s_a = reinterpret_cast<u16 *>(xs + ((2 * F.Tcap + 15) & ~15)), *s_b = s_a + F.Scap;
                double *s_rank = reinterpret_cast<double *>(s_b + F.Scap);
                int nsurv = 0;
                {
                    int pi = 0, po = 0;   // lane's pair: cand[pi] with cand[pi + 1 + po]
                    if (lane < NP) {
                        int jj;
                        pair_decode(lane, T2, pi, jj);
                        po = jj - pi - 1;
                    }
#pragma unroll 1
                    for (int pb = 0; pb < NP; pb += 32) {
                        const bool act = pb + lane < NP;
                        bool pass = false;
                        int ta = 0, tb = 0;
                        if (act) {
                            ta = cand[pi];
                            tb = cand[pi + 1 + po];
                            // crit >= k2 <=> every non-ignored locus is covered at k2 by one of the two clades (:610)
                            pass = ((mk1[ta] | mk1[tb]) & um) == um;
                        }
                        const u32 m = __ballot_sync(FULL, pass);
                        if (pass) {
                            const int dst = nsurv + __popc(m & lt_mask());
                            if (dst < F.Scap) {
                                const bool sw = cl_id[ta] > cl_id[tb];   // clade1 < clade2 by name (:608)
                                s_a[dst] = (u16)(sw ? tb : ta);
                                s_b[dst] = (u16)(sw ? ta : tb);
                            }
                        }
                        nsurv += __popc(m);
                        if (act) {
                            po += 32;
                            while (pi < T2 - 1 && po >= T2 - 1 - pi) { po -= T2 - 1 - pi; ++pi; }
                        }
                    }
                }
                if (nsurv > F.Scap) { fallback = true; reason = 6; break; }
                if (lane == 0) stat[ST_PSCORE] += (unsigned long long)nsurv;
                __syncwarp();
                // exact crit / rank of the survivors; best = last maximal rank in (clade1, clade2) iteration order
                u64 bbits = 0;
                int bx = -1, by = -1, bq = -1;
#pragma unroll 1
                for (int q = lane; q < nsurv; q += 32) {
                    double crit;
                    const double rank = row_stats(L, s_a[q], s_b[q], &crit);
                    s_rank[q] = rank;
                    const u64 b = dbits(rank);
                    const int x = cl_id[s_a[q]], y = cl_id[s_b[q]];
                    if (bq < 0 || b > bbits || (b == bbits && (x > bx || (x == bx && y > by)))) { bbits = b; bx = x; by = y; bq = q; }
                }
                const u64 wb = warp_max_u64(bq >= 0 ? bbits : 0ull);
                const long long wx = warp_max_ll((bq >= 0 && bbits == wb) ? (long long)bx : -1);
                const long long wy = warp_max_ll((bq >= 0 && bbits == wb && (long long)bx == wx) ? (long long)by : -1);
                __syncwarp();
                if (nsurv > 0) {
                    if (P.p.sister_penalty != 0) {   // parents of the level's listed clades, once (check_sister_penalty :717-744)
#pragma unroll 2
                        for (int t = lane; t < T; t += 32) {
                            const int id = cl_id[t];
                            hkey[t] = tax.listed[id] ? tax.parent[id] : -1;
                        }
                        __syncwarp();
                    }
                    const int owner = __ffs(__ballot_sync(FULL, bq >= 0 && bbits == wb && (long long)bx == wx && (long long)by == wy)) - 1;
                    const int bp = __shfl_sync(FULL, bq, owner);
                    const int bi = s_a[bp], bj = s_b[bp];
                    FTwoEval be;
                    eval_two_fast(L, tax, P, bi, bj, be);
                    double bcrit;
                    const double brank = row_stats(L, bi, bj, &bcrit);
                    // meld_two (:633-669) over the options within --range
                    u8 *memA = reinterpret_cast<u8 *>(hval), *memB = memA + F.Tcap;   // the hash is dead
#pragma unroll 1
                    for (int t = lane; t < T; t += 32) memA[t] = memB[t] = 0;
                    __syncwarp();
                    int nk = 0, nbad = 0, ndiff = 0, la = -1, lb = -1;
                    bool near = false;
#pragma unroll 1
                    for (int q = lane; q < nsurv; q += 32) {
                        const double d = brank - s_rank[q];
                        if (q != bp && fabs(d) <= GUARD) near = true;
                        if (P.p.disambiguate_two != 0 && fabs(d - P.p.range) <= GUARD) near = true;
                        if (!(d <= P.p.range)) continue;   // :636
                        FTwoEval ev;
                        eval_two_fast(L, tax, P, s_a[q], s_b[q], ev);
                        ++nk;
                        nbad += !ev.ok;
                        ndiff += !(ev.A == be.A && ev.B == be.B && ev.amb == be.amb);   // meld_precheck (:671-676)
                        la = lca2(tax, la, ev.c1);
                        lb = lca2(tax, lb, ev.c2);
                        memA[ev.t1] = 1;
                        memB[ev.t2] = 1;
                    }
                    if (__any_sync(FULL, near)) { trip = true; break; }
                    nk = warp_sum(nk);
                    nbad = warp_sum(nbad);
                    ndiff = warp_sum(ndiff);
                    la = warp_lca(tax, la);
                    lb = warp_lca(tax, lb);
                    __syncwarp();
                    bool have = true, melded = false;
                    int c1 = be.c1, c2 = be.c2;
                    if (nk == 1 || P.p.disambiguate_two == 0) {
                    } else if (P.p.disambiguate_two == 1) {
                        have = false;
                    } else if (nbad > 0 || ndiff > 0) {
                        have = false;
                    } else {
                        c1 = la;
                        c2 = lb;
                        melded = true;
                        if (!P.p.allow_lca) {   // post-meld LCA check (:661-665)
                            const int l = lca2(tax, c1, c2);
                            if (l == c1 || l == c2) have = false;
                        }
                    }
                    if (have && be.ok) {
                        R.call = WFL_CALL_LGT;
                        R.b1 = be.c1;
                        R.b2 = be.c2;
                        R.c1 = c1;
                        R.c2 = c2;
                        R.lca = lca2(tax, c1, c2);   // waafle_orgscorer.py:882
                        R.crit = bcrit;
                        R.rank = brank;
                        R.dir = be.dir;
                        if (lane < G) {
                            const u32 m = 1u << lane;
                            a.o.synteny[l0 + l_raw[lane]] = ign ? '~' : (be.amb & m) ? '*' : (be.A & m) ? 'A' : (be.B & m) ? 'B' : '!';
                        }
                        if (melded) {
                            // distinct melded clades per side, ascending node index (lists in the dead group scores)
                            int *klist = reinterpret_cast<int *>(g_score);
#pragma unroll 1
                            for (int side = 0; side < 2; ++side) {
                                const u8 *mem = side ? memB : memA;
                                int kb = 0;
#pragma unroll 1
                                for (int base = 0; base < T; base += 32) {
                                    const int t = base + lane;
                                    const bool f = t < T && mem[t];
                                    const u32 m = __ballot_sync(FULL, f);
                                    if (f) klist[side * F.Tcap + kb + __popc(m & lt_mask())] = cl_id[t];
                                    kb += __popc(m);
                                }
                                if (side) R.nb = kb; else R.na = kb;
                            }
                            __syncwarp();
                            long long mb = 0;
                            if (lane == 0) mb = (long long)atomicAdd(&a.ctr->mem_pool_used, (unsigned long long)(R.na + R.nb));
                            R.mem = __shfl_sync(FULL, mb, 0);
                            emit_members_sorted(a, klist, R.na, R.mem, lane);
                            emit_members_sorted(a, klist + F.Tcap, R.nb, R.mem + R.na, lane);
                        }
                        finished = true;
                        break;
                    }
                }
            }

            // ---- K9: not explained at this level: stop or lift (evaluate_contig :571-581) ----
            if (T == 0 || hasroot) { finished = true; break; }
            if (iter >= 100) { fallback = true; break; }   // runaway: the exact pipeline reports it
            ++R.lifts;
            ++iter;
            __syncwarp();
        }
        if (H == 0 || G == 0) finished = !fallback;

        if (trip) reason = 7;
        if (trip && lane == 0) ++stat[ST_TRIPS];
        if (fallback || trip || !finished) {
            if (lane == 0) {
                // capacity overflows are worth a second pass with a larger slice; the rest goes straight to the exact pipeline
                const bool retry = reason == 1 || reason == 3 || reason == 4 || reason == 5 || reason == 6;
                const unsigned long long s = atomicAdd(retry ? a.fb_count : a.fb_final_count, 1ull);
                (retry ? a.fb_list : a.fb_final)[s] = (int)c;
                atomicAdd(&a.ctr->fb_reason[reason & 7], 1ull);
                // placeholder record (the speculative compaction walks every contig): overwritten by the exact pipeline
                a.o.call[c] = WFL_CALL_UNCLASSIFIED;
                a.o.n_mem_a[c] = a.o.n_mem_b[c] = 0;
                a.o.mem_pos[c] = 0;
            }
        } else {
            if (lane == 0) write_result_fast(a, c, R);
            if (lane == 0) ++stat[ST_DONE];
        }
        __syncwarp();
    }
    __syncwarp();
    if (lane == 0) {
        if (stat[ST_PAIRS]) atomicAdd(&a.ctr->matched_pairs, stat[ST_PAIRS]);
        if (stat[ST_GROUPS]) atomicAdd(&a.ctr->groups, stat[ST_GROUPS]);
        if (stat[ST_LEVELS]) atomicAdd(&a.ctr->levels, stat[ST_LEVELS]);
        if (stat[ST_PTEST]) atomicAdd(&a.ctr->pairs_tested, stat[ST_PTEST]);
        if (stat[ST_PSCORE]) atomicAdd(&a.ctr->pairs_scored, stat[ST_PSCORE]);
        if (stat[ST_DONE]) atomicAdd(&a.ctr->smem_contigs, stat[ST_DONE]);
        if (stat[ST_TRIPS]) atomicAdd(&a.ctr->guard_trips, stat[ST_TRIPS]);
        if (stat[ST_REFINED]) atomicAdd(&a.ctr->refined_groups, stat[ST_REFINED]);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Slice layout for given capacities; returns the slice size in bytes (multiple of 16).
//   Kcap records of one locus, Mcap records of the contig, Ccap clade-hash slots (power of two), Tcap clades and
//   Ncap (clade, locus) groups per level.  Hcap (hits per contig) follows: the spans are staged over the per-level arrays.
int fast_layout(FastCfg &F, int Kcap, int Mcap, int Ccap, int Tcap, int Ncap, int Scap) {
    auto al = [](int x) { return (x + 15) & ~15; };
    int o = 0;
    Ncap = std::max(Ncap, Tcap);                           // the two-clade member lists (2 x Tcap ints) reuse g_score
    F.Kcap = Kcap; F.Mcap = Mcap; F.cmask = Ccap - 1; F.Tcap = Tcap; F.Ncap = Ncap;
    F.o_stat = o; o += al(8 * 8 + 8);                      // 8 counters + the level's clade count
    F.o_llo = o; o += al(4 * GMAX);
    F.o_llen = o; o += al(4 * GMAX);
    F.o_lraw = o; o += al(4 * GMAX);
    F.o_lstr = o; o += al(GMAX);
    F.o_lbase = o; o += al(2 * (GMAX + 2));
    F.o_maxv = o; o += al(8 * GMAX);
    F.o_unk = o; o += al(8 * GMAX);
    F.o_rhit = o; o += al(2 * Mcap);
    F.o_rab = o; o += al(4 * Mcap);
    // scratch: one-clade rank cache (Tcap doubles) | two-clade candidates (Tcap u16) + survivors (12 bytes each)
    F.Scap = (Scap + 1) & ~1;
    F.x_bytes = std::max(8 * Tcap, al(2 * Tcap) + 12 * F.Scap);
    F.o_x = o; o += al(F.x_bytes);
    const int lvl0 = o;
    F.o_hkey = o; o += al(4 * Ccap);                       // also: one-clade member list, parents of the clades (Tcap ints)
    F.o_hval = o; o += al(2 * Ccap);                       // also memA / memB (2 x Tcap bytes)
    F.o_clid = o; o += al(4 * Tcap);
    F.o_mk0 = o; o += al(4 * Tcap);
    F.o_mk1 = o; o += al(4 * Tcap);
    F.o_mk2 = o; o += al(4 * Tcap);
    F.o_clhead = o; o += al(2 * Tcap);
    F.o_cltail = o; o += al(2 * Tcap);
    F.o_gscore = o; o += al(8 * Ncap);
    F.o_gu = o; o += al(4 * Kcap);
    F.o_gt = o; o += al(2 * Ncap);
    F.o_gnext = o; o += al(2 * Ncap);
    F.o_gloc = o; o += al(Ncap);
    o = std::max(o, lvl0 + al(14 * Kcap + 16));            // the per-locus sort of the prologue is staged from o_hkey on
    F.Hcap = std::min(65535, (o - lvl0) / 8);              // as are the hits: u32 span + u32 locus mask each
    F.slice_bytes = o;
    return o;
}

int fast_warps_per_cta() { return FAST_WPC; }

cudaError_t launch_fast(const FastArgs &a, bool packed, int grid, cudaStream_t s) {
    const size_t smem = (size_t)FAST_WPC * a.cfg.slice_bytes;
    cudaError_t rc;
    if (packed) {
        rc = cudaFuncSetAttribute(wfl_fast_contigs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        wfl_fast_contigs<true><<<grid, 32 * FAST_WPC, smem, s>>>(a);
    } else {
        rc = cudaFuncSetAttribute(wfl_fast_contigs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        wfl_fast_contigs<false><<<grid, 32 * FAST_WPC, smem, s>>>(a);
    }
    return cudaGetLastError();
}

int fast_ctas_per_sm(const FastCfg &F, bool packed, size_t smem_per_sm) {
    int n = 0;
    const size_t smem = (size_t)FAST_WPC * F.slice_bytes;
    if (packed) {
        cudaFuncSetAttribute(wfl_fast_contigs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wfl_fast_contigs<true>, 32 * FAST_WPC, smem);
    } else {
        cudaFuncSetAttribute(wfl_fast_contigs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, wfl_fast_contigs<false>, 32 * FAST_WPC, smem);
    }
    (void)smem_per_sm;
    return n;
}

}  // namespace wfl
