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
//  * once per contig, ONE coalesced pass over the contig's hit columns: every hit is tested against the loci
//    (attach_hits); the hits attached to at least one locus are STAGED in shared memory (score, span, clade) together
//    with the mask of their loci -- hits attached to nothing never reach the site arrays and are dropped here;
//  * per taxonomy level the staged hits are visited in (clade ascending, score descending) order.  The front end's
//    packer delivers every contig's hits in exactly that order for level 0 (packing.Batch.sort_hits), so the order is
//    only CHECKED there; after a lift (clade := parent, raise_taxonomy :431-445) or for unsorted input a bitonic sort of
//    the hit permutation restores it.  In that order
//      - the distinct clades of the level are the runs of equal clade: dense handles in ascending node index == name
//        order, no hash table;
//      - one pass emits the records (hit x locus) locus-major, so the records of a (clade, locus) group are adjacent
//        and descending in score;
//  * gene scores, one LANE per (clade, locus) group: the group's records are streamed in score order while the union of
//    the better hits is kept as one interval [ua, ub): every record adds score x (newly covered sites).  That sum / n is
//    the closed form of np.mean of the site array (:371-382, :399-406).  A record that leaves a gap sends the group to
//    a generic endpoint sweep (cold);
//  * numpy's pairwise rounding is not reproduced, so the value is within ~1e-14 of np.mean; instead every DECISION
//    is protected by a guard band:
//      - a gene score within 1e-12 of a threshold it is compared with (k1, k2, 1e-6) is recomputed exactly, in
//        numpy's pairwise order, by the pipeline's own group_mean();
//      - a rank within 1e-12 of the best rank (arg-max) or of best - range (meld set) cannot be settled locally:
//        the contig is handed to the exact pipeline;
//    so calls / clades / synteny are bit-exact and crit / rank agree to <= 1e-12 (north_star tolerance);
//  * gene-score rows are a clade-major CSR addressed through each clade's locus-presence bitmask: row(t, i) is one
//    popc away, no list walks.
// Anything the slice cannot hold (too many attached hits / records / clades / groups / pairs) is retried by a second
// launch with a larger slice; long genes, > 32 loci, --min-overlap <= 0, guard-band trips and malformed input go to the
// exact pipeline.
#include "wfl_warp_common.cuh"

namespace wfl {

namespace {

constexpr int FAST_WPC = 2;     // warps (= contigs in flight) per CTA
constexpr int GMAX = 32;        // retained loci per contig: gene bitmasks are one 32-bit word
constexpr int NONE16 = 0xffff;
#ifndef WFL_FAST_CPSM
#define WFL_FAST_CPSM 6         // resident CTAs per SM the kernel is compiled for (register budget: 168; shared memory
                                // limits the residency before the registers do)
#endif

#define SM(T, off) (reinterpret_cast<T *>(slice + (off)))

// warp reductions on the redux unit (sm_80+): max of 64-bit keys in two 32-bit steps, signed 32-bit max, integer sum
__device__ __forceinline__ u64 rmax_u64(u64 v) {
    const u32 hi = (u32)(v >> 32);
    const u32 mh = __reduce_max_sync(FULL, hi);
    const u32 ml = __reduce_max_sync(FULL, hi == mh ? (u32)v : 0u);
    return ((u64)mh << 32) | (u64)ml;
}
__device__ __forceinline__ int rmax_i32(int v) { return __reduce_max_sync(FULL, v); }
__device__ __forceinline__ int rsum_i32(int v) { return __reduce_add_sync(FULL, v); }

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
    const u32 *pres;         // loci in which the clade has a gene score
    const u16 *cstart;       // first entry of the clade's row in `row` (entries in locus order)
    const double *row;
    const double *unk_row;
    const int *cl_id;
    const u32 *mk[3];        // per clade gene bitmasks: score >= k1 / k2 / 1e-6
    const int *l_len;
    const int *cl_par;       // two-clade search with sister penalty: parent of every LISTED clade of the level, else -1
};

// gene-score row of one clade: score at locus i, 0 where the clade has no entry (waafle_orgscorer.py:404-405)
struct RowRef {
    u32 p;
    int c;
    bool dense;
    __device__ __forceinline__ void open(const FLevel &L, int t) {
        dense = t == L.t_unk;
        p = dense ? 0u : L.pres[t];
        c = dense ? 0 : (int)L.cstart[t];
    }
    __device__ __forceinline__ double at(const FLevel &L, int i) const {
        if (dense) return L.unk_row[i];
        return ((p >> i) & 1u) ? L.row[c + __popc(p & ((1u << i) - 1u))] : 0.0;
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
    RowRef w1, w2;
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
    ev.ok = ok;   // the sister penalty is checked by the whole warp (sister_hit_warp)
}

// check_sister_penalty (waafle_orgscorer.py:717-744) for one option, all lanes together over the level's clades: is a B
// locus claimed by a sister of clade1, or (unless the direction is known) an A locus by a sister of clade2?
__device__ __noinline__ bool sister_hit_warp(const FLevel &L, const DevTax &tax, const DevParams &P, int t1, int t2, int c1, int c2,
                                             bool dir, u32 A, u32 B) {
    const u32 *msis = L.mk[P.sis_sel];
    const int p1 = tax.parent[c1], p2 = tax.parent[c2];
    bool hit = false;
#pragma unroll 1
    for (int t = lane_id(); t < L.T; t += 32) {
        if (t == t1 || t == t2) continue;
        const int px = L.cl_par[t];   // get_sisters works on the taxonomy file's rows
        const bool s1 = px == p1, s2 = (px == p2) && !dir;
        if (!s1 && !s2) continue;
        const u32 ms = msis[t];
        if ((s1 && (ms & B)) || (s2 && (ms & A))) hit = true;
    }
    return __any_sync(FULL, hit);
}

// python slice [h1 : h2+1] of the locus' site array covered by a hit (waafle_orgscorer.py:373-382), for overlapping
// intervals: a | b << 16
__device__ __forceinline__ void hit_slice(u32 span, int lmin, int llen, int &a, int &b) {
    a = max(0, (int)(span & 0xffffu) - lmin);
    b = min(llen - 1, (int)(span >> 16) - lmin) + 1;
}

// Cold paths of one (clade, locus) group whose records are rec[r0 ...] while their clade is `clade` (and < lend): they are
// gathered into the warp's global scratch (slices and scores, descending score order) and either swept over their distinct
// endpoints (a better hit left a gap) or -- `exact`, for a score inside the guard band -- summed in numpy's pairwise order.
__device__ __noinline__ double group_slow(const FastArgs &a, char *scratch, int cap, const u16 *rec, const u32 *hsp, const int *hcl,
                                          const double *hv, int r0, int lend, int clade, int lmin, int n, bool exact) {
    int *ra = reinterpret_cast<int *>(scratch), *rb = ra + cap;
    double *rv = reinterpret_cast<double *>(rb + cap);
    int m = 0;
    bool desc = true;
#pragma unroll 1
    for (int q = r0; q < lend && m < cap; ++q) {
        const int h = rec[q];
        if (hcl[h] != clade) break;
        int s, e;
        hit_slice(hsp[h], lmin, n, s, e);
        ra[m] = s;
        rb[m] = e;
        rv[m] = hv[h];
        desc &= m == 0 || rv[m - 1] >= rv[m];
        ++m;
    }
    if (exact) {
        const PlanEntry pe = a.plan_index[n];
        return group_mean(ra, rb, rv, 0, m, n, desc, pe.k8, a.plan_data + pe.off, (int)pe.nleaf);
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

// The writer prints the scores a contig reports (crit / rank of the chosen clade or pair, waafle_orgscorer.py:447-461) with
// four decimals (utils.py:137-139).  A fast-path score is within ~1e-14 of the reference's; when it lies within 1e-11 of a
// value that prints differently on either side (x.xxxx5 -- systematic for 3-decimal pident values), the gene scores of the
// chosen clade(s) -- at most one group per locus -- are recomputed in numpy's summation order, one lane per locus, and
// written back into the rows; the caller then re-runs row_stats, so the TSV bytes equal the reference's.
__device__ __forceinline__ bool near_print_edge(double v) {
    const double x = v * 1e4;
    return fabs(x - floor(x) - 0.5) <= 1e-7;
}
// Loci whose gene scores decide the reported scores of clade t1 (or of the pair t1, t2): all non-ignored loci if the rank is
// near a print edge, else only those attaining the minimum (crit).
__device__ __forceinline__ u32 edge_loci(const FLevel &L, int t1, int t2, double crit, double rank, int lane) {
    if (near_print_edge(rank)) return L.um;
    bool sel = false;
    if ((L.um >> lane) & 1u) {
        RowRef w1, w2;
        w1.open(L, t1);
        w2.open(L, t2 >= 0 ? t2 : t1);
        const double v = fmax(w1.at(L, lane), w2.at(L, lane));
        sel = v <= crit + 1e-11;
    }
    return __ballot_sync(FULL, sel);
}

// rows_exact: exact gene scores of clade t at the loci `loci`, lane per locus.  Records of a
// locus are sorted by (clade, score descending): the group is found by binary search.  Groups of more than LC records go
// through the warp's global scratch, one at a time.
constexpr int LC = 24;
__device__ __noinline__ void rows_exact(const FastArgs &a, char *cold, int coldcap, const FLevel &L, double *row, int t, u32 loci,
                                        int lane, const u16 *l_base, const int *l_lo, const u16 *rec, const u32 *hsp,
                                        const int *hcl, const double *hv) {
    if (t == L.t_unk) return;   // (dense row: the caller does not report it from here)
    const u32 pm = L.pres[t];
    const int clade = L.cl_id[t];
    const bool act = ((pm & loci) >> lane) & 1u;
    bool big = false;
    int r0 = 0, lend = 0, lmin = 0, n = 1;
    double sc = 0.0;
    if (act) {
        int lo = l_base[lane], hi = l_base[lane + 1];
        lend = hi;
        lmin = l_lo[lane];
        n = L.l_len[lane];
#pragma unroll 1
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (hcl[rec[mid]] < clade) lo = mid + 1; else hi = mid;
        }
        r0 = lo;
        int ra[LC], rb[LC];
        double rv[LC];
        int m = 0;
        bool desc = true;
#pragma unroll 1
        for (int q = r0; q < lend; ++q) {
            const int h = rec[q];
            if (hcl[h] != clade) break;
            if (m == LC) { big = true; break; }
            hit_slice(hsp[h], lmin, n, ra[m], rb[m]);
            rv[m] = hv[h];
            desc &= m == 0 || rv[m - 1] >= rv[m];
            ++m;
        }
        if (!big) {
            const PlanEntry pe = a.plan_index[n];
            sc = group_mean(ra, rb, rv, 0, m, n, desc, pe.k8, a.plan_data + pe.off, (int)pe.nleaf);
        }
    }
    u32 bm = __ballot_sync(FULL, act && big);
#pragma unroll 1
    while (bm) {
        const int ln = __ffs(bm) - 1;
        bm &= bm - 1;
        if (lane == ln) sc = group_slow(a, cold, coldcap, rec, hsp, hcl, hv, r0, lend, clade, lmin, n, true);
        __syncwarp();
    }
    if (act) row[(int)L.cstart[t] + __popc(pm & ((1u << lane) - 1u))] = sc;
    __syncwarp();
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
    __device__ __forceinline__ void all(long long h, int &q1, int &q2, u32 &fl, int &tx) const {
        span(h, q1, q2);
        fl = flags(h);
        tx = b.hit_taxon[h];
    }
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
    __device__ __forceinline__ void all(long long h, int &q1, int &q2, u32 &fl, int &tx) const {
        span(h, q1, q2);
        const u32 w = b.hit_tax16[h];
        fl = ((w >> 15) & 1u) | (((w >> 14) & 1u) << 1);
        tx = (int)(w & 0x3fffu);
    }
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
    const u32 le_mask = lt_mask() | (1u << lane);

    // loci
    int *l_lo = SM(int, F.o_llo), *l_len = SM(int, F.o_llen), *l_raw = SM(int, F.o_lraw);
    signed char *l_str = SM(signed char, F.o_lstr);
    u16 *l_base = SM(u16, F.o_lbase);            // records of locus i: rec[l_base[i] .. l_base[i + 1])
    u16 *l_cur = SM(u16, F.o_lcur);              // emission cursor / record count of the locus
    u16 *l_last = SM(u16, F.o_llast);            // clade handle of the locus' last emitted record
    u64 *maxb = SM(u64, F.o_maxb);               // per-locus max gene score (order-preserving bits)
    double *unk_row = SM(double, F.o_unk);
    // staged (attached) hits, level-invariant but for the clade
    double *hv = SM(double, F.o_hv);             // waafle score, clamped at 0 (sites start at 0, np.zeros :381)
    u32 *hsp = SM(u32, F.o_hsp);                 // span on the contig: min | max << 16
    int *hcl = SM(int, F.o_hcl);                 // clade of the hit at the current level
    u16 *hid = SM(u16, F.o_hid);                 // index of the hit in the contig (annotation winners; only with systems)
    u8 *hloc = SM(u8, F.o_hloc);                 // locus of the entry
    // per level
    u16 *rec = SM(u16, F.o_rec);                 // records (staged hit index), locus-major, (clade, score desc) inside
    u16 *gstart = SM(u16, F.o_gstart), *g_t = SM(u16, F.o_gt);   // (clade, locus) groups: first record, clade handle
    u8 *g_loc = SM(u8, F.o_gloc);
    double *row = SM(double, F.o_row);           // gene scores, clade-major CSR
    u16 *hp = SM(u16, F.o_row);                  // sort permutation of the entries (dead before the rows are written)
    int *cl_id = SM(int, F.o_clid);
    u32 *mk0 = SM(u32, F.o_mk0), *mk1 = SM(u32, F.o_mk1), *mk2 = SM(u32, F.o_mk2), *pres = SM(u32, F.o_pres);
    u16 *cstart = SM(u16, F.o_cstart);
    char *xs = SM(char, F.o_x);                  // search scratch (aliases gstart / g_t / g_loc)
    unsigned long long *stat = SM(unsigned long long, F.o_stat);   // per-warp counters, flushed once at exit
    char *gscr = a.scratch + ((size_t)blockIdx.x * FAST_WPC + wid) * (size_t)F.scratch_bytes;
    char *cold = gscr;                           // record arrays of the cold gene-score paths
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
        const int H = (int)min((long long)0x7fffffff, a.b.hit_off[c + 1] - h0), Graw = (int)(a.b.locus_off[c + 1] - l0);
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
        if (lane < GMAX) l_cur[lane] = 0;
        __syncwarp();
        const u32 allG = G >= 32 ? 0xffffffffu : ((1u << G) - 1u);
        // a locus without an entry scores 0 (waafle_orgscorer.py:404-405): its mask bit is (0 >= threshold)
        const u32 init0 = thr3[0] <= 0.0 ? allG : 0u, init1 = thr3[1] <= 0.0 ? allG : 0u;

        // ---- K1, once per contig: one pass over the hit columns (attach_hits :359-369); attached hits are staged ----
        int Hs = 0;   // staged entries == records of the contig
        if (!fallback && H > 0 && G > 0) {
            bool big = false, bad = false;
            const int *anc0 = a.anc + (size_t)min(R.lifts, a.anc_rows - 1) * (size_t)tax.n_nodes;
            // the columns of the NEXT tile of hits are requested before the current one is processed
            int n_q1 = 0, n_q2 = 0, n_tx = 0;
            u32 n_fl = 0;
            double n_sc = 0.0;
            if (lane < H) {
                hits.all(h0 + lane, n_q1, n_q2, n_fl, n_tx);
                n_sc = a.b.hit_score[h0 + lane];
            }
#pragma unroll 1
            for (int base = 0; base < H; base += 32) {
                const int h = base + lane;
                const int q1 = n_q1, q2 = n_q2, tx = n_tx;
                const u32 fl = n_fl;
                const double sc = n_sc;
                if (h + 32 < H) {
                    hits.all(h0 + h + 32, n_q1, n_q2, n_fl, n_tx);
                    n_sc = a.b.hit_score[h0 + h + 32];
                }
                u32 mb = 0;
                int hmin = 0, hmax = 0;
                if (h < H) {
                    hmin = min(q1, q2);
                    hmax = max(q1, q2);
                    big |= hmin < 0 || hmax > 65535;
                    if (fl & 1) {   // scov_modified >= --min-scov (:362)
                        u32 cand = 0;
#pragma unroll 1
                        for (int i = 0; i < G; ++i) {
                            const int lmin = l_lo[i], lmax = lmin + l_len[i] - 1;
                            bool ok = !(lmin > hmax || hmin > lmax);
                            if (P.p.stranded) {   // hit.sstrand == locus.strand (:365)
                                const signed char ls = l_str[i];
                                ok &= (ls == '-' && (fl & 2)) || (ls == '+' && !(fl & 2));
                            }
                            cand |= (ok ? 1u : 0u) << i;
                        }
                        mb = cand;
                    }
                }
                {   // calc_overlap >= --min-overlap on the candidates (usually one per hit: the lanes stay converged)
                    u32 cand = mb;
#pragma unroll 1
                    while (cand) {
                        const int i = __ffs(cand) - 1;
                        cand &= cand - 1u;
                        const int lmin = l_lo[i], llen = l_len[i];
                        if (!overlap_ok(hmin, hmax, lmin, lmin + llen - 1, llen, P.p.min_overlap)) mb &= ~(1u << i);
                    }
                }
                // one staged ENTRY per (hit, locus) match, in hit order (a hit is rarely attached to more than one locus)
                int tot, ex;
                if (__all_sync(FULL, (mb & (mb - 1u)) == 0u)) {
                    const u32 att = __ballot_sync(FULL, mb != 0u);
                    ex = __popc(att & lt_mask());
                    tot = __popc(att);
                } else {
                    ex = warp_excl_scan(__popc(mb), tot);
                }
                if (mb) {
                    int cl = tax.root;
                    if ((u32)tx >= (u32)tax.n_nodes) bad = true;   // malformed input: the exact pipeline reports it
                    else cl = R.lifts ? anc0[tx] : tx;             // row 0 of the ancestor table is the identity
                    int s = Hs + ex;
                    u32 m2 = mb;
#pragma unroll 1
                    while (m2) {
                        const int i = __ffs(m2) - 1;
                        m2 &= m2 - 1u;
                        if (s < F.Hcap) {
                            hv[s] = sc > 0.0 ? sc : 0.0;
                            hsp[s] = (u32)(hmin & 0xffff) | ((u32)(hmax & 0xffff) << 16);
                            hcl[s] = cl;
                            hloc[s] = (u8)i;
                            if (S > 0) hid[s] = (u16)h;
                        }
                        ++s;
                    }
                }
                Hs += tot;
                // entries per locus
                u32 m2 = mb;
#pragma unroll 1
                while (__any_sync(FULL, m2 != 0u)) {
                    const int i = m2 ? __ffs(m2) - 1 : -1;
                    const u32 peers = __match_any_sync(FULL, i >= 0 ? i : 64 + lane);
                    if (i >= 0 && (peers & lt_mask()) == 0u) l_cur[i] += (u16)__popc(peers);
                    m2 &= m2 - 1u;
                    __syncwarp();
                }
            }
            big = __any_sync(FULL, big);
            bad = __any_sync(FULL, bad);
            if (big) { fallback = true; reason = 2; }          // spans are staged in 16 bits
            else if (bad) { fallback = true; reason = 4; }
            else if (H > 65535) { fallback = true; reason = 1; }
            else if (Hs > F.Hcap) { fallback = true; reason = 3; }
            // locus-major record layout
            int cnt = lane < G ? (int)l_cur[lane] : 0;
            int tot;
            const int ex = warp_excl_scan(cnt, tot);
            if (lane <= G && lane < GMAX) l_base[lane] = (u16)min(ex, 0xffff);
            if (lane == 0 && G == GMAX) l_base[GMAX] = (u16)min(tot, 0xffff);
            __syncwarp();
            // ---- K3: annotation winners, level-independent (score_hit :384-392): the LAST hit attaining the max score
            //      per (locus, system), two phases of shared-memory atomics (score bits, then hit index) ----
            if (S > 0 && !fallback) {
                int *annw = reinterpret_cast<int *>(unk_row);
#pragma unroll 1
                for (int s2 = 0; s2 < S; ++s2) {
                    if (lane < GMAX) { maxb[lane] = 0ull; annw[lane] = -1; }
                    __syncwarp();
#pragma unroll 1
                    for (int phase = 0; phase < 2; ++phase) {
#pragma unroll 1
                        for (int j = lane; j < Hs; j += 32) {
                            const int hh = hid[j], i = hloc[j];
                            const double sc = a.b.hit_score[h0 + hh];
                            if (((hits.sysmask(h0 + hh) >> s2) & 1u) && sc >= P.ann_thr) {
                                const u64 sb = dbits(sc);
                                if (phase == 0) atomicMax(&maxb[i], sb);
                                else if (maxb[i] == sb) atomicMax(&annw[i], hh);
                            }
                        }
                        __syncwarp();
                    }
                    if (lane < G) a.o.ann_winner[(l0 + l_raw[lane]) * S + s2] = annw[lane] >= 0 ? (int)(h0 + annw[lane]) : -1;
                    __syncwarp();
                }
            }
        } else if (lane <= GMAX && lane <= G) {
            l_base[lane] = 0;
        }
        __syncwarp();

        int iter = 0;
#pragma unroll 1
        while (!fallback && !finished && H > 0 && G > 0) {
            // ============================ one taxonomy level ============================
            if (lane == 0) ++stat[ST_LEVELS];
            // ---- order of the level: staged hits by (clade, score descending); checked, sorted only if needed ----
            bool ident = true;
            {
                bool ok = true;
#pragma unroll 1
                for (int j = lane + 1; j < Hs; j += 32) {
                    const int c0 = hcl[j - 1], c1 = hcl[j];
                    ok &= c0 < c1 || (c0 == c1 && hv[j - 1] >= hv[j]);
                }
                ident = __all_sync(FULL, ok);
            }
            if (!ident) {
                // bitonic sort of 64-bit keys  clade | (qmax - quantised score) | entry index  in shared memory (aliasing the
                // per-level arrays).  The score is quantised to a.qbits bits: entries whose scores collide may come out in the
                // wrong order -- the gene-score walk checks the true order and sends such a group to the order-free sweep.
                int Pn = 32;
                while (Pn < Hs) Pn <<= 1;
                if (8 * Pn > F.sort_bytes) { fallback = true; reason = 3; break; }
                u64 *sk = reinterpret_cast<u64 *>(rec);
                const int qb = a.qbits;
                const double qs = (double)(1ull << (qb - 1));
                const u64 qmax = (1ull << qb) - 1ull;
#pragma unroll 1
                for (int j = lane; j < Pn; j += 32) {
                    u64 key = ~0ull;
                    if (j < Hs) {
                        const double x = hv[j] * qs;
                        const u64 q = x >= (double)qmax ? qmax : (u64)x;
                        key = ((u64)(u32)hcl[j] << (qb + 16)) | ((qmax - q) << 16) | (u64)j;
                    }
                    sk[j] = key;
                }
                __syncwarp();
#pragma unroll 1
                for (int k = 2; k <= Pn; k <<= 1) {
#pragma unroll 1
                    for (int jj = k >> 1; jj > 0; jj >>= 1) {
#pragma unroll 2
                        for (int x = lane; x < (Pn >> 1); x += 32) {
                            const int lo = ((x & ~(jj - 1)) << 1) | (x & (jj - 1)), hi = lo | jj;
                            const u64 u = sk[lo], w = sk[hi];
                            if ((u > w) == ((lo & k) == 0)) { sk[lo] = w; sk[hi] = u; }
                        }
                        __syncwarp();
                    }
                }
#pragma unroll 1
                for (int j = lane; j < Hs; j += 32) hp[j] = (u16)(sk[j] & 0xffffull);
                __syncwarp();
            }

            // ---- clades of the level (runs of equal clade: handles in ascending node index) and, in the same pass,
            //      the records (locus-major) and the (clade, locus) groups ----
            int T = 0, N = 0, t_unk = -1;
            bool ovf = false;
            int ovf_reason = 4;
            {
                if (lane < G) {
                    l_cur[lane] = l_base[lane];
                    l_last[lane] = (u16)NONE16;
                    maxb[lane] = dbits(0.0);
                }
                __syncwarp();
                int run = 0, n_lt = 0, carry = -1;
                bool seen_unk = false;
                const int unk_id = tax.unknown;
#pragma unroll 1
                for (int base = 0; base < Hs; base += 32) {
                    const int j = base + lane;
                    const bool act = j < Hs;
                    const int h = act ? (ident ? j : (int)hp[j]) : 0;
                    const int cl = act ? hcl[h] : 0x7fffffff;
                    int up = __shfl_up_sync(FULL, cl, 1);
                    if (lane == 0) up = carry;
                    const bool head = act && cl != up;
                    const u32 hb = __ballot_sync(FULL, head);
                    int t = run + __popc(hb & le_mask) - 1;
                    bool isunk = false;
                    if (spike) {
                        // "Unknown" is a clade of every level (waafle_orgscorer.py:416-418): its handle is inserted in
                        // node-index order if no hit carries it
                        const u32 ltb = __ballot_sync(FULL, act && cl < unk_id), eqb = __ballot_sync(FULL, act && cl == unk_id);
                        seen_unk |= eqb != 0u;
                        n_lt += __popc(hb & ltb);
                        if (!seen_unk && act && cl > unk_id) ++t;
                        isunk = act && cl == unk_id;
                    }
                    if (head) {
                        if (t < F.Tcap) {
                            cl_id[t] = cl;
                            mk0[t] = init0; mk1[t] = init1; mk2[t] = 0u;
                            pres[t] = 0u;
                        } else {
                            ovf = true;
                        }
                    }
                    run += __popc(hb);
                    carry = __shfl_sync(FULL, cl, 31);
                    if (__any_sync(FULL, ovf) || run + 1 > F.Tcap) { ovf = true; break; }
                    __syncwarp();
                    // records (locus-major, this order inside a locus) and groups of this tile
                    {
                        const int i = act ? (int)hloc[h] : -1;
                        const u32 peers = __match_any_sync(FULL, act ? i : 64 + lane);
                        const u32 below = peers & lt_mask();
                        const int pt_lane = __shfl_sync(FULL, t, below ? 31 - __clz(below) : lane);
                        bool gh = false;
                        int slot = 0;
                        if (act) {
                            slot = (int)l_cur[i] + __popc(below);
                            rec[slot] = (u16)h;
                            const int pt = below ? pt_lane : (int)l_last[i];
                            // with "assign-unknown" the hits of a taxon NAMED Unknown carry no gene score: the spiked row
                            // replaces gene_scores["Unknown"] (:416-418)
                            gh = pt != t && !isunk;
                        }
                        const u32 gb = __ballot_sync(FULL, gh);
                        if (gh) {
                            const int g = N + __popc(gb & lt_mask());
                            if (g < F.Ncap) {
                                gstart[g] = (u16)slot;
                                g_t[g] = (u16)t;
                                g_loc[g] = (u8)i;
                            }
                            atomicOr(&pres[t], 1u << i);
                        }
                        N += __popc(gb);
                        __syncwarp();
                        if (act && (peers >> lane) == 1u) {   // last of its locus in this tile
                            l_cur[i] = (u16)(slot + 1);
                            l_last[i] = (u16)t;
                        }
                        __syncwarp();
                    }
                    if (N > F.Ncap) { ovf = true; ovf_reason = 5; break; }
                }
                T = run;
                if (spike && !ovf) {
                    t_unk = n_lt;
                    if (!seen_unk) {
                        if (lane == 0) {
                            cl_id[t_unk] = unk_id;
                            mk0[t_unk] = mk1[t_unk] = mk2[t_unk] = 0u;
                            pres[t_unk] = 0u;
                        }
                        ++T;
                    }
                }
            }
            ovf = __any_sync(FULL, ovf);
            if (ovf) { fallback = true; reason = ovf_reason; break; }
            __syncwarp();
            if (lane == 0) {
                stat[ST_GROUPS] += (unsigned long long)N;
                if (iter == 0) stat[ST_PAIRS] += (unsigned long long)Hs;
            }

            // ---- rows: clade-major CSR through the presence masks ----
            {
                int acc = 0;
#pragma unroll 1
                for (int base = 0; base < T; base += 32) {
                    const int t = base + lane;
                    const int cnt = t < T ? __popc(pres[t]) : 0;
                    int tot;
                    const int ex = warp_excl_scan(cnt, tot);
                    if (t < T) cstart[t] = (u16)(acc + ex);
                    acc += tot;
                }
            }
            __syncwarp();

            // ---- K2: gene scores, one lane per (clade, locus) group; guard band, masks, per-locus max ----
#pragma unroll 1
            for (int base = 0; base < N; base += 32) {
                const int g = base + lane;
                const bool act = g < N;
                int t = 0, i = 0, r0 = 0, cl = 0, llen = 1, lmin = 0, lend = 0;
                bool cplx = false;
                double sc = 0.0;
                if (act) {
                    t = g_t[g];
                    i = g_loc[g];
                    r0 = gstart[g];
                    cl = cl_id[t];
                    llen = l_len[i];
                    lmin = l_lo[i];
                    lend = l_base[i + 1];
                    int h = rec[r0];
                    int ua, ub;
                    hit_slice(hsp[h], lmin, llen, ua, ub);
                    double vprev = hv[h];
                    double sum = vprev * (double)(ub - ua);   // the group opens with its best hit
#pragma unroll 1
                    for (int q = r0 + 1; q < lend; ++q) {
                        h = rec[q];
                        if (hcl[h] != cl) break;
                        const double vp = hv[h];
                        if (vp > vprev) { cplx = true; break; }   // scores that collided in the quantised sort: order-free sweep
                        vprev = vp;
                        if (ua == 0 && ub == llen) {           // the union covers the gene: nothing can be added
                            if (ident) break;
                            continue;                          // (after a quantised sort the rest of the order is still checked)
                        }
                        if (!(vp > 0.0)) break;                // descending: the rest contributes nothing
                        int pa, pb;
                        hit_slice(hsp[h], lmin, llen, pa, pb);
                        if (pa > ub || pb < ua) { cplx = true; break; }   // a gap: the union is no longer one interval
                        const int add = max(0, ua - pa) + max(0, pb - ub);
                        if (add) sum += vp * (double)add;
                        ua = min(ua, pa);
                        ub = max(ub, pb);
                    }
                    sc = sum / (double)llen;
                }
                // groups whose hits left a gap: generic endpoint sweep; scores inside the guard band of a threshold:
                // recomputed in numpy's summation order.  Both cold, one lane at a time (global scratch of the warp).
                u32 cm = __ballot_sync(FULL, act && cplx);
#pragma unroll 1
                while (cm) {
                    const int ln = __ffs(cm) - 1;
                    cm &= cm - 1;
                    if (lane == ln) sc = group_slow(a, cold, F.Hcap, rec, hsp, hcl, hv, r0, lend, cl, lmin, llen, false);
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
                        sc = group_slow(a, cold, F.Hcap, rec, hsp, hcl, hv, r0, lend, cl, lmin, llen, true);
                        atomicAdd(&stat[ST_REFINED], 1ull);
                    }
                    __syncwarp();
                }
                if (act) {
                    const u32 pm = pres[t], bit = 1u << i;
                    row[(int)cstart[t] + __popc(pm & (bit - 1u))] = sc;
                    if (sc >= thr3[0]) atomicOr(&mk0[t], bit); else if (init0) atomicAnd(&mk0[t], ~bit);
                    if (sc >= thr3[1]) atomicOr(&mk1[t], bit); else if (init1) atomicAnd(&mk1[t], ~bit);
                    if (sc >= thr3[2]) atomicOr(&mk2[t], bit);
                    if (cl != tax.unknown) atomicMax(&maxb[i], dbits(sc));   // waafle_orgscorer.py:409-411
                }
            }
            __syncwarp();

            // ---- K4: weak loci (update_gene_scores :407-429) ----
            bool ign = false;
            double mx = 0.0;
            if (lane < G) {
                mx = dbits_inv(maxb[lane]);
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
                if (lane == 0) { mk0[t_unk] = m0; mk1[t_unk] = m1; mk2[t_unk] = m2; }
            }
            __syncwarp();
            if (iter == 0 && nun == 0) { finished = true; break; }   // "empty" contig (:959): unclassified

            int hasroot = 0;
#pragma unroll 1
            for (int t = lane; t < T; t += 32) hasroot |= cl_id[t] == tax.root;
            hasroot = __any_sync(FULL, hasroot);

            int *cl_par = reinterpret_cast<int *>(xs + ((2 * F.Tcap + 15) & ~15));   // after the two-clade candidates
            FLevel L{G, T, t_unk, um, nun, pres, cstart, row, unk_row, cl_id, {mk0, mk1, mk2}, l_len, cl_par};

            // ---- K6: one-clade search (explain_one :585-597) ----
            {
                u64 bbits = 0;
                int bid = -1, btl = -1;
                double bcrit = 0.0;
                double *rcache = reinterpret_cast<double *>(xs + 4 * F.Tcap);   // after the kept-clade list
#pragma unroll 1
                for (int t = lane; t < T; t += 32) {
                    if ((mk0[t] & um) != um) continue;   // crit >= k1
                    double crit;
                    const double rank = row_stats(L, t, -1, &crit);
                    rcache[t] = rank;   // for the meld pass
                    const u64 b = dbits(rank);
                    if (b > bbits || (b == bbits && cl_id[t] > bid)) { bbits = b; bid = cl_id[t]; btl = t; bcrit = crit; }
                }
                const u64 wb = rmax_u64(bbits);
                // ties: last in name order == largest node index (meld_one :623-624, canonical order)
                const long long wid2 = wb != 0 ? (long long)rmax_i32(bbits == wb ? bid : -1) : -1;
                if (wb != 0 && wid2 >= 0) {
                    const int owner = __ffs(__ballot_sync(FULL, bbits == wb && (long long)bid == wid2)) - 1;
                    const int tb = __shfl_sync(FULL, btl, owner);
                    const double brank = dbits_inv(wb);
                    // meld_one (:621-631): options within --range of the best; guard the arg-max and the range edge
                    int my = -1, nk = 0;
                    bool near = false;
                    int *klist = reinterpret_cast<int *>(xs);   // kept clades (node ids), ascending
                    int kbase = 0;
#pragma unroll 1
                    for (int base = 0; base < T; base += 32) {
                        const int t = base + lane;
                        bool kept = false;
                        if (t < T && (mk0[t] & um) == um) {
                            const double d = brank - rcache[t];
                            // (the best option itself has d == 0 exactly: no guard, or --range 0 would trip every contig)
                            if (t != tb && fabs(d) <= GUARD) near = true;
                            if (t != tb && P.p.disambiguate_one == 1 && fabs(d - P.p.range) <= GUARD) near = true;
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
                    R.crit = __shfl_sync(FULL, bcrit, owner);
                    R.rank = brank;
                    if (near_print_edge(R.crit) || near_print_edge(R.rank)) {
                        // a reported score the writer would round the other way if it were 1e-12 off: recompute the clade's
                        // gene scores in numpy's summation order (the dense Unknown row: exact pipeline)
                        if (tb == t_unk) { trip = true; break; }
                        const u32 el = edge_loci(L, tb, -1, R.crit, R.rank, lane);
                        rows_exact(a, cold, F.Hcap, L, row, tb, el, lane, l_base, l_lo, rec, hsp, hcl, hv);
                        R.rank = row_stats(L, tb, -1, &R.crit);
                        if (lane == 0) ++stat[ST_REFINED];
                    }
                    R.call = WFL_CALL_NO_LGT;
                    R.b1 = R.c1 = cl_id[tb];
                    if (P.p.disambiguate_one == 1) {
                        R.c1 = warp_lca(tax, my);
                        R.na = rsum_i32(nk);
                    }
                    if (lane < G)   // set_synteny_one (:495-509)
                        a.o.synteny[l0 + l_raw[lane]] = ign ? '~' : ((mk0[tb] >> lane) & 1u ? 'A' : '!');
                    __syncwarp();
                    if (R.na > 0) {
                        long long mb = 0;
                        if (lane == 0) mb = (long long)atomicAdd(&a.ctr->mem_pool_used, (unsigned long long)R.na);
                        R.mem = __shfl_sync(FULL, mb, 0);
#pragma unroll 1
                        for (int j = lane; j < R.na; j += 32)
                            if (R.mem + j < a.o.mem_pool_cap) a.o.mem_pool[R.mem + j] = klist[j];
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
                // scratch after the candidates and the clade parents: melded-member flags, survivors of the mask prefilter
                u8 *memA = reinterpret_cast<u8 *>(cl_par + F.Tcap), *memB = memA + F.Tcap;
                u16 *s_a = reinterpret_cast<u16 *>(memB + ((F.Tcap + 15) & ~15)), *s_b = s_a + F.Scap;
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
                const u64 wb = rmax_u64(bq >= 0 ? bbits : 0ull);
                const long long wx = rmax_i32((bq >= 0 && bbits == wb) ? bx : -1);
                const long long wy = rmax_i32((bq >= 0 && bbits == wb && (long long)bx == wx) ? by : -1);
                __syncwarp();
                if (nsurv > 0) {
                    if (P.p.sister_penalty != 0) {   // parents of the level's listed clades, once (check_sister_penalty :717-744)
#pragma unroll 2
                        for (int t = lane; t < T; t += 32) {
                            const int id = cl_id[t];
                            cl_par[t] = tax.listed[id] ? tax.parent[id] : -1;
                        }
                        __syncwarp();
                    }
                    const int owner = __ffs(__ballot_sync(FULL, bq >= 0 && bbits == wb && (long long)bx == wx && (long long)by == wy)) - 1;
                    const int bp = __shfl_sync(FULL, bq, owner);
                    const int bi = s_a[bp], bj = s_b[bp];
                    FTwoEval be;
                    eval_two_fast(L, tax, P, bi, bj, be);
                    if (P.p.sister_penalty != 0 && be.ok && sister_hit_warp(L, tax, P, be.t1, be.t2, be.c1, be.c2, be.dir, be.A, be.B))
                        be.ok = false;
                    double bcrit;
                    const double brank = row_stats(L, bi, bj, &bcrit);
                    // meld_two (:633-669) over the options within --range
#pragma unroll 1
                    for (int t = lane; t < T; t += 32) memA[t] = memB[t] = 0;
                    __syncwarp();
                    int nk = 0, nbad = 0, ndiff = 0, la = -1, lb = -1;
                    bool near = false;
#pragma unroll 1
                    for (int qb = 0; qb < nsurv; qb += 32) {
                        const int q = qb + lane;
                        bool kept = false;
                        FTwoEval ev;
                        ev.ok = true;
                        ev.t1 = ev.t2 = ev.c1 = ev.c2 = 0;
                        ev.A = ev.B = 0u;
                        ev.dir = false;
                        if (q < nsurv) {
                            const double d = brank - s_rank[q];
                            if (q != bp && fabs(d) <= GUARD) near = true;
                            if (q != bp && P.p.disambiguate_two != 0 && fabs(d - P.p.range) <= GUARD) near = true;
                            kept = d <= P.p.range;   // :636
                            if (kept) eval_two_fast(L, tax, P, s_a[q], s_b[q], ev);
                        }
                        if (P.p.sister_penalty != 0) {   // one kept option at a time, all lanes over the clades
                            u32 pend = __ballot_sync(FULL, kept && ev.ok);
#pragma unroll 1
                            while (pend) {
                                const int src = __ffs(pend) - 1;
                                pend &= pend - 1;
                                const bool hit = sister_hit_warp(L, tax, P, __shfl_sync(FULL, ev.t1, src), __shfl_sync(FULL, ev.t2, src),
                                                                 __shfl_sync(FULL, ev.c1, src), __shfl_sync(FULL, ev.c2, src),
                                                                 __shfl_sync(FULL, (int)ev.dir, src) != 0, __shfl_sync(FULL, ev.A, src),
                                                                 __shfl_sync(FULL, ev.B, src));
                                if (lane == src && hit) ev.ok = false;
                            }
                        }
                        if (kept) {
                            ++nk;
                            nbad += !ev.ok;
                            ndiff += !(ev.A == be.A && ev.B == be.B && ev.amb == be.amb);   // meld_precheck (:671-676)
                            la = lca2(tax, la, ev.c1);
                            lb = lca2(tax, lb, ev.c2);
                            memA[ev.t1] = 1;
                            memB[ev.t2] = 1;
                        }
                    }
                    if (__any_sync(FULL, near)) { trip = true; break; }
                    nk = rsum_i32(nk);
                    nbad = rsum_i32(nbad);
                    ndiff = rsum_i32(ndiff);
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
                        R.crit = bcrit;
                        R.rank = brank;
                        if (near_print_edge(bcrit) || near_print_edge(brank)) {   // see the one-clade search
                            if (be.t1 == t_unk || be.t2 == t_unk) { trip = true; break; }
                            const u32 el = edge_loci(L, be.t1, be.t2, bcrit, brank, lane);
                            rows_exact(a, cold, F.Hcap, L, row, be.t1, el, lane, l_base, l_lo, rec, hsp, hcl, hv);
                            rows_exact(a, cold, F.Hcap, L, row, be.t2, el, lane, l_base, l_lo, rec, hsp, hcl, hv);
                            R.rank = row_stats(L, be.t1, be.t2, &R.crit);
                            if (lane == 0) ++stat[ST_REFINED];
                        }
                        R.call = WFL_CALL_LGT;
                        R.b1 = be.c1;
                        R.b2 = be.c2;
                        R.c1 = c1;
                        R.c2 = c2;
                        R.lca = lca2(tax, c1, c2);   // waafle_orgscorer.py:882
                        R.dir = be.dir;
                        if (lane < G) {
                            const u32 m = 1u << lane;
                            a.o.synteny[l0 + l_raw[lane]] = ign ? '~' : (be.amb & m) ? '*' : (be.A & m) ? 'A' : (be.B & m) ? 'B' : '!';
                        }
                        if (melded) {
                            // distinct melded clades per side, ascending node index == handle order (lists in the dead rows)
                            int *klist = reinterpret_cast<int *>(row);
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
#pragma unroll 1
                            for (int j = lane; j < R.na + R.nb; j += 32) {
                                const int id = j < R.na ? klist[j] : klist[F.Tcap + j - R.na];
                                if (R.mem + j < a.o.mem_pool_cap) a.o.mem_pool[R.mem + j] = id;
                            }
                        }
                        finished = true;
                        break;
                    }
                }
            }

            // ---- K9: not explained at this level: stop or lift (evaluate_contig :571-581, raise_taxonomy :431-445) ----
            if (T == 0 || hasroot) { finished = true; break; }
            if (iter >= 100) { fallback = true; break; }   // runaway: the exact pipeline reports it
            ++R.lifts;
            ++iter;
#pragma unroll 2
            for (int j = lane; j < Hs; j += 32) hcl[j] = tax.parent[hcl[j]];
            __syncwarp();
        }
        if (H == 0 || G == 0) finished = !fallback;

        if (trip) reason = 7;
        if (trip && lane == 0) ++stat[ST_TRIPS];
        if (fallback || trip || !finished) {
            if (lane == 0) {
                // capacity overflows are worth a second pass with a larger slice; the rest goes straight to the exact pipeline
                const bool retry = (reason == 1 && H <= 65535) || reason == 3 || reason == 4 || reason == 5 || reason == 6;
                if (retry && reason == 6 && a.fb_pairs) {
                    a.fb_pairs[atomicAdd(a.fb_pairs_count, 1ull)] = (int)c;
                } else {
                    const unsigned long long s = atomicAdd(retry ? a.fb_count : a.fb_final_count, 1ull);
                    (retry ? a.fb_list : a.fb_final)[s] = (int)c;
                }
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
//   Hcap attached hits and Mcap records of the contig, Tcap clades and Ncap (clade, locus) groups per level, Scap
//   two-clade pairs that survive the mask prefilter.
int fast_layout(FastCfg &F, int Hcap, int Mcap, int Tcap, int Ncap, int Scap, int n_systems) {
    auto al = [](int x) { return (x + 15) & ~15; };
    int o = 0;
    Hcap = std::min(std::max(Hcap, Mcap), 65534);         // staged entries ARE the records
    Mcap = Hcap;
    Tcap = std::min((Tcap + 15) & ~15, 65520);            // multiple of 16: the search scratch is carved in Tcap units
    Ncap = std::min(std::max(Ncap, Tcap), 65534);          // the two-clade member lists (2 x Tcap ints) reuse the rows
    int Pcap = 32;
    while (Pcap < Hcap) Pcap <<= 1;
    F.Hcap = Hcap; F.Pcap = Pcap; F.Mcap = Mcap; F.Tcap = Tcap; F.Ncap = Ncap;
    F.Scap = (Scap + 1) & ~1;
    F.o_stat = o; o += al(8 * 8 + 8);
    F.o_llo = o; o += al(4 * GMAX);
    F.o_llen = o; o += al(4 * GMAX);
    F.o_lraw = o; o += al(4 * GMAX);
    F.o_lstr = o; o += al(GMAX);
    F.o_lbase = o; o += al(2 * (GMAX + 2));
    F.o_lcur = o; o += al(2 * GMAX);
    F.o_llast = o; o += al(2 * GMAX);
    F.o_maxb = o; o += al(8 * GMAX);
    F.o_unk = o; o += al(8 * GMAX);
    F.o_hv = o; o += al(8 * Hcap);
    F.o_hsp = o; o += al(4 * Hcap);
    F.o_hcl = o; o += al(4 * Hcap);
    F.o_hid = o; o += n_systems > 0 ? al(2 * Hcap) : 0;
    F.o_hloc = o; o += al(Hcap);
    // per-level records, then the groups; the search scratch aliases the groups (dead once the gene scores are written):
    //   two-clade: candidates (Tcap u16) | clade parents (Tcap int) | member flags (2 x Tcap u8) | survivors (12 B each)
    //   one-clade: kept clades (Tcap int) | ranks (Tcap double)
    F.o_rec = o; o += al(2 * Mcap);
    const int groups = al(2 * Ncap) + al(2 * Ncap) + al(Ncap);
    F.x_bytes = std::max(std::max(groups, 12 * Tcap + 16), al(2 * Tcap) + 4 * Tcap + 2 * al(Tcap) + 12 * F.Scap + 16);
    F.o_x = o;
    F.o_gstart = o;
    F.o_gt = F.o_gstart + al(2 * Ncap);
    F.o_gloc = F.o_gt + al(2 * Ncap);
    o += al(F.x_bytes);
    F.o_clid = o; o += al(4 * Tcap);
    F.o_mk0 = o; o += al(4 * Tcap);
    F.o_mk1 = o; o += al(4 * Tcap);
    F.o_mk2 = o; o += al(4 * Tcap);
    F.o_pres = o; o += al(4 * Tcap);
    F.o_cstart = o; o += al(2 * Tcap);
    // the key array of the per-level sort (8 B x Pcap) spans the records / groups and the clade table, all rebuilt after it
    if (o - F.o_rec < 8 * Pcap) o = F.o_rec + 8 * Pcap;
    F.sort_bytes = o - F.o_rec;
    // the rows; until they are written the sort permutation (u16[Pcap]) lives here
    F.o_row = o; o += al(std::max(8 * Ncap, 2 * Pcap));
    F.o_hp = F.o_row;
    F.slice_bytes = o;
    F.scratch_bytes = 16 * Hcap + 16;
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
