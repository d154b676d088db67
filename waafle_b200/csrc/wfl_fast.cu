// The fused fast path of the orgscorer engine: ONE kernel, one warp per contig, every piece of per-contig
// state in shared memory (a per-warp slice of the CTA's dynamic shared memory), all taxonomy levels in a
// data-dependent loop.  Replaces, for contigs with <= 32 retained loci, the pipeline's prepare / regroup /
// K2 sort / K2 / masks / one / two / lift launch chain and its global-memory workspace round trips.
//
// Reference path (waafle/waafle_orgscorer.py): attach_hits :359-392, update_gene_scores :394-429,
// raise_taxonomy :431-445, evaluate_contig :566-583, explain_one/two :585-619, meld_one/two :621-676,
// LGT checks :678-744.
//
// How it differs from the exact pipeline (wfl_pipeline.cu):
//  * per taxonomy level the contig's hits are re-streamed ONE LOCUS AT A TIME from global memory (HBM on the
//    first pass, L2 afterwards) into a small record buffer; no per-hit or per-record state survives a locus, so
//    the slice is ~10 KB whatever the hit count, and a lift is "the same pass with the next row of the
//    ancestor table" (waafle_orgscorer.py:431-445 re-bins site arrays by parent: at level l the envelope of
//    clade c is over the hits whose l-th ancestor is c);
//  * gene scores are the CLOSED FORM of the envelope integral, sum_j v_j * |I_j \ union of better hits| / n
//    (records picked in descending score order, union kept as one interval; a gap sends the group to a
//    generic endpoint sweep).  numpy's pairwise rounding is not reproduced, so the value is within ~1e-14
//    of np.mean; instead every DECISION is protected by a guard band:
//      - a gene score within 1e-12 of a threshold it is compared with (k1, k2, 1e-6) is recomputed
//        exactly, in numpy's pairwise order, by the pipeline's own group_mean();
//      - a rank within 1e-12 of the best rank (arg-max) or of best - range (meld set) cannot be settled
//        locally: the contig is handed to the exact pipeline (fallback list);
//    so calls / clades / synteny are bit-exact and crit / rank agree to <= 1e-12 (north_star tolerance).
//  * clades are handled through a per-level open-addressing hash (dense handles in first-seen order); every
//    tie-break that the reference resolves by name order uses the node index explicitly.
// Anything the slice cannot hold (too many loci / records per locus / clades / groups / pairs), long genes,
// --min-overlap <= 0 and malformed input go to the fallback list and are scored by the exact pipeline.
#include "wfl_warp_common.cuh"

namespace wfl {

namespace {

constexpr int FAST_WPC = 4;     // warps (= contigs in flight) per CTA
constexpr int EMPTY_KEY = -1;
constexpr int GMAX = 32;        // retained loci per contig: gene bitmasks are one 32-bit word

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

// Exact np.mean of the group's site array (numpy pairwise order) through the pipeline's K2 code: the group's records are
// copied to a scratch area as the int / double arrays group_mean() walks.  Cold: near-threshold groups only.
__device__ __noinline__ double group_mean_exact(const FastArgs &a, char *scratch, int Kcap, const double *bv, const u32 *bab,
                                                const u16 *bord, int rs, int re, int n) {
    int *ra = reinterpret_cast<int *>(scratch), *rb = ra + Kcap;
    double *rv = reinterpret_cast<double *>(rb + Kcap);
    const int k = re - rs;
#pragma unroll 1
    for (int j = 0; j < k; ++j) {
        const int r = bord[rs + j];
        const u32 ab = bab[r];
        ra[j] = (int)(ab & 0xffffu);
        rb[j] = (int)(ab >> 16);
        rv[j] = bv[r];
    }
    const PlanEntry pe = a.plan_index[n];
    return group_mean(ra, rb, rv, 0, k, n, false, pe.k8, a.plan_data + pe.off, (int)pe.nleaf);
}

// Per-level view of a contig for the search routines (all pointers into the warp's slice).
struct FLevel {
    int G, T, t_unk;         // t_unk: handle of the spiked Unknown (dense row unk_row), -1 if none
    u32 um;                  // non-ignored loci
    int nun;
    const u16 *goff;         // groups of locus i: [goff[i], goff[i+1]), ascending handle
    const u16 *g_t;
    const double *g_score;
    const double *unk_row;
    const int *cl_id;
    const u32 *mk[3];        // per clade gene bitmasks: score >= k1 / k2 / 1e-6
    const int *l_len;
};

// gene score of clade handle t at locus i (0 where the clade has no entry, waafle_orgscorer.py:404-405)
__device__ __forceinline__ double score_at(const FLevel &L, int t, int i) {
    if (t == L.t_unk) return L.unk_row[i];
    int p = L.goff[i];
    const int e = L.goff[i + 1];
#pragma unroll 1
    while (p < e && (int)L.g_t[p] < t) ++p;
    return (p < e && (int)L.g_t[p] == t) ? L.g_score[p] : 0.0;
}

// Contig.score (waafle_orgscorer.py:447-461) over the non-ignored loci: crit = min, rank = np.mean (n <= 32 values:
// numpy sums n < 8 sequentially, otherwise eight strided accumulators, the fixed tree, then the tail).
__device__ __noinline__ double row_stats(const FLevel &L, int t1, int t2, double *crit_out) {
    const int n = L.nun;
    double r[8];
    double crit = __longlong_as_double(0x7ff0000000000000ll), res = 0.0;
    u32 bits = L.um;
    const int body = n < 8 ? 0 : (n & ~7);
    int idx = 0;
#pragma unroll 1
    while (bits) {
        const int i = __ffs(bits) - 1;
        bits &= bits - 1;
        double v = score_at(L, t1, i);
        if (t2 >= 0) v = fmax(v, score_at(L, t2, i));
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
            const int id = L.cl_id[t];
            const int px = tax.listed[id] ? tax.parent[id] : -1;   // get_sisters works on the taxonomy file's rows
            const bool s1 = px == p1, s2 = (px == p2) && !ev.dir;
            if (!s1 && !s2) continue;
            const u32 ms = msis[t];
            // a B locus is penalised by clade1's sisters, an A locus by clade2's
            if ((s1 && (ms & B)) || (s2 && (ms & A))) ok = false;
        }
    }
    ev.ok = ok;
}

// Warp-cooperative get-or-insert into the level's clade hash; returns the dense handle (first-seen order: lanes of one
// call are inserted in lane order, so handles -- and with them the group order and every sum -- are deterministic).
__device__ __forceinline__ int clade_handle(int *hkey, u16 *hval, int *cl_id, u32 *mk0, u32 *mk1, u32 *mk2, int cmask, int Tcap,
                                            int &T, u32 i0, u32 i1, u32 i2, int key, bool active, bool &ovf) {
    u32 slot = hash32(key) & (u32)cmask;
    int handle = 0;
    bool done = !active;
#pragma unroll 1
    for (int it = 0;; ++it) {
        if (__all_sync(FULL, done)) break;
        if (it > 2 * cmask + 2) ovf = true;
        if (__any_sync(FULL, ovf)) break;
        const int cur = done ? 0 : hkey[slot];
        if (!done && cur == key) { handle = hval[slot]; done = true; }
        const bool want = !done && cur == EMPTY_KEY;
        const u32 wm = __ballot_sync(FULL, want);
        if (wm) {
            u32 peers = 0;
            if (want) peers = __match_any_sync(wm, slot);
            const bool leader = want && (peers & lt_mask()) == 0;
            const u32 lm = __ballot_sync(FULL, leader);
            if (leader) {
                const int hnew = T + __popc(lm & lt_mask());
                if (hnew < Tcap) {
                    hkey[slot] = key;
                    hval[slot] = (u16)hnew;
                    cl_id[hnew] = key;
                    mk0[hnew] = i0; mk1[hnew] = i1; mk2[hnew] = i2;
                } else {
                    ovf = true;
                }
            }
            T += __popc(lm);
            __syncwarp();
        }
        if (!done && !want && cur != key) slot = (slot + 1) & (u32)cmask;   // occupied by another clade
    }
    return handle;
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
template <bool PACKED>
struct Hits;
template <>
struct Hits<false> {
    const DevBatch &b;
    double min_scov;
    __device__ __forceinline__ void span(long long h, int &q1, int &q2) const { q1 = b.hit_qstart[h]; q2 = b.hit_qend[h]; }
    // (scov ok, strand, taxon) of a hit whose span already overlaps the locus
    __device__ __forceinline__ bool rest(long long h, signed char &hs, int &tx) const {
        hs = b.hit_strand[h];
        tx = b.hit_taxon[h];
        return b.hit_scov[h] >= min_scov;   // waafle_orgscorer.py:362
    }
    __device__ __forceinline__ u32 sysmask(long long h) const { return b.hit_sysmask[h]; }
};
template <>
struct Hits<true> {
    const DevBatch &b;
    double min_scov;
    __device__ __forceinline__ void span(long long h, int &q1, int &q2) const { q1 = b.hit_qstart16[h]; q2 = b.hit_qend16[h]; }
    __device__ __forceinline__ bool rest(long long h, signed char &hs, int &tx) const {
        const u32 w = b.hit_tax16[h];
        hs = (w & 0x4000u) ? '-' : '+';
        tx = (int)(w & 0x3fffu);
        return (w & 0x8000u) != 0u;         // host applied scov_modified >= min_scov while packing
    }
    __device__ __forceinline__ u32 sysmask(long long h) const { return b.hit_sysmask8[h]; }
};

}  // namespace

template <bool PACKED>
__global__ void __launch_bounds__(32 * FAST_WPC, 4) wfl_fast_contigs(const FastArgs a) {
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
    u16 *goff = SM(u16, F.o_goff);
    double *maxv = SM(double, F.o_maxv), *unk_row = SM(double, F.o_unk);
    double *bv = SM(double, F.o_bv);
    u32 *bab = SM(u32, F.o_bab);
    int *bcl = SM(int, F.o_bcl);          // clade ids of the buffered records; dead once handles are known ...
    u16 *bord = SM(u16, F.o_bcl);         // ... then the multisplit order and
    u16 *grs = bord + F.Kcap;             // the group starts of the locus live there
    u16 *bt = SM(u16, F.o_bt);
    int *bh = SM(int, F.o_bh);
    int *hkey = SM(int, F.o_hkey);
    u16 *hval = SM(u16, F.o_hval);
    int *cl_id = SM(int, F.o_clid);
    u32 *mk0 = SM(u32, F.o_mk0), *mk1 = SM(u32, F.o_mk1), *mk2 = SM(u32, F.o_mk2);
    u16 *cur = SM(u16, F.o_cur);
    double *g_score = SM(double, F.o_gscore);
    u16 *g_t = SM(u16, F.o_gt);
    char *scratch = a.scratch + ((size_t)blockIdx.x * FAST_WPC + wid) * (size_t)F.Kcap * 16;

    unsigned long long st_pairs = 0, st_groups = 0, st_levels = 0, st_ptest = 0, st_pscore = 0, st_done = 0, st_refined = 0,
                       st_trips = 0;

#pragma unroll 1
    for (;;) {
        long long c = -1;
        if (lane == 0) {
            const unsigned long long w = atomicAdd(a.wq, 1ull);
            if ((long long)w < a.n_work) c = a.work_list ? (long long)a.work_list[w] : a.work_base + (long long)w;
        }
        c = __shfl_sync(FULL, c, 0);
        if (c < 0) break;
        const long long h0 = a.b.hit_off[c], l0 = a.b.locus_off[c];
        const int H = (int)(a.b.hit_off[c + 1] - h0), Graw = (int)(a.b.locus_off[c + 1] - l0);
        bool fallback = false, trip = false;
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

        int iter = 0;
#pragma unroll 1
        while (!fallback && !finished && H > 0 && G > 0) {
            // ============================ one taxonomy level ============================
            const int *anc = a.anc + (size_t)min(R.lifts, a.anc_rows - 1) * (size_t)tax.n_nodes;
            int T = 0, N = 0;
            bool ovf = false;
#pragma unroll 1
            for (int s = lane; s <= F.cmask; s += 32) hkey[s] = EMPTY_KEY;
            __syncwarp();
            if (spike)   // "Unknown" is a clade of every level (waafle_orgscorer.py:416-418): handle 0
                (void)clade_handle(hkey, hval, cl_id, mk0, mk1, mk2, F.cmask, F.Tcap, T, 0u, 0u, 0u, tax.unknown, lane == 0, ovf);
            const int t_unk = spike ? 0 : -1;
            ++st_levels;

#pragma unroll 1
            for (int i = 0; i < G && !ovf; ++i) {
                const int lmin = l_lo[i], llen = l_len[i], lmax = lmin + llen - 1;
                const signed char ls = l_str[i];
                // ---- K1: stream the contig's hits, keep the ones attached to locus i (attach_hits :359-369) ----
                int k = 0;
#pragma unroll 1
                for (int base = 0; base < H; base += 32) {
                    const int h = base + lane;
                    bool mt = false;
                    int hmin = 0, hmax = 0, tx = 0;
                    if (h < H) {
                        int q1, q2;
                        hits.span(h0 + h, q1, q2);
                        hmin = min(q1, q2);
                        hmax = max(q1, q2);
                        if (!(lmin > hmax || hmin > lmax)) {
                            signed char hs;
                            mt = hits.rest(h0 + h, hs, tx) && !(P.p.stranded && hs != ls) &&
                                 overlap_ok(hmin, hmax, lmin, lmax, llen, P.p.min_overlap);
                        }
                    }
                    const u32 m = __ballot_sync(FULL, mt);
                    if (!m) continue;
                    if (mt) {
                        const int slot = k + __popc(m & lt_mask());
                        if ((u32)tx >= (u32)tax.n_nodes) {
                            ovf = true;   // malformed input: the exact pipeline reports it
                        } else if (slot < F.Kcap) {
                            const double sc = a.b.hit_score[h0 + h];
                            // python slice [h1 : h2+1] of the site array (:373-382); sites start at 0 (np.zeros)
                            const int s1 = max(0, hmin - lmin), e1 = min(llen - 1, hmax - lmin) + 1;
                            bv[slot] = sc > 0.0 ? sc : 0.0;
                            bab[slot] = (u32)s1 | ((u32)e1 << 16);
                            bcl[slot] = anc[tx];
                            if (S > 0) bh[slot] = h;
                        }
                    }
                    k += __popc(m);
                }
                if (k > F.Kcap) ovf = true;
                ovf = __any_sync(FULL, ovf);
                if (ovf) break;
                __syncwarp();
                if (iter == 0) st_pairs += (unsigned long long)k;
                if (lane == 0) goff[i] = (u16)N;
                if (k == 0) { if (lane == 0) maxv[i] = 0.0; continue; }

                // ---- K3: annotation winners, level-independent (score_hit :384-392): last hit with the max score ----
                if (S > 0 && iter == 0) {
#pragma unroll 1
                    for (int s2 = 0; s2 < S; ++s2) {
                        u64 bb = 0;
                        long long bw = -1;
#pragma unroll 1
                        for (int r = lane; r < k; r += 32) {
                            const double sc = bv[r];
                            if (((hits.sysmask(h0 + bh[r]) >> s2) & 1u) && sc >= P.ann_thr) {
                                const u64 sb = dbits(sc);
                                if (sb >= bb) { bb = sb; bw = bh[r]; }
                            }
                        }
                        const u64 mx = warp_max_u64(bb);
                        const long long w = warp_max_ll((mx != 0 && bb == mx) ? bw : -1);
                        if (lane == 0) a.o.ann_winner[(l0 + l_raw[i]) * S + s2] = w >= 0 ? (int)(h0 + w) : -1;
                    }
                }

                // ---- clade handles of the records (dense, first-seen order) ----
#pragma unroll 1
                for (int base = 0; base < k; base += 32) {
                    const int r = base + lane;
                    const int hd = clade_handle(hkey, hval, cl_id, mk0, mk1, mk2, F.cmask, F.Tcap, T, init0, init1, init2,
                                                r < k ? bcl[r] : 0, r < k, ovf);
                    if (r < k) bt[r] = (u16)hd;
                }
                ovf = __any_sync(FULL, ovf);
                if (ovf) break;
                __syncwarp();

                // ---- group by clade: stable multisplit of the buffer by handle (bcl is dead: bord / grs take its place) ----
#pragma unroll 1
                for (int b = lane; b <= T; b += 32) cur[b] = 0;
                __syncwarp();
#pragma unroll 1
                for (int base = 0; base < k; base += 32) {
                    const int r = base + lane;
                    const int b = r < k ? (int)bt[r] : T;
                    const u32 peers = __match_any_sync(FULL, b);
                    if ((peers & lt_mask()) == 0) cur[b] += (u16)__popc(peers);
                    __syncwarp();
                }
                {
                    int carry = 0;
#pragma unroll 1
                    for (int base = 0; base < T; base += 32) {
                        const int b = base + lane;
                        const int cnt = b < T ? (int)cur[b] : 0;
                        int tot;
                        const int ex = warp_excl_scan(cnt, tot);
                        if (b < T) cur[b] = (u16)(carry + ex);
                        carry += tot;
                    }
                }
                __syncwarp();
#pragma unroll 1
                for (int base = 0; base < k; base += 32) {
                    const int r = base + lane;
                    const int b = r < k ? (int)bt[r] : T;
                    const u32 peers = __match_any_sync(FULL, b);
                    if (r < k) bord[cur[b] + __popc(peers & lt_mask())] = (u16)r;
                    __syncwarp();
                    if (r < k && (peers & lt_mask()) == 0) cur[b] += (u16)__popc(peers);
                    __syncwarp();
                }
                // ---- groups = runs of equal handle in bord ----
                int ng = 0;
#pragma unroll 1
                for (int base = 0; base < k; base += 32) {
                    const int q = base + lane;
                    bool head = false;
                    int t = 0;
                    if (q < k) {
                        t = bt[bord[q]];
                        head = q == 0 || t != (int)bt[bord[q - 1]];
                    }
                    const u32 m = __ballot_sync(FULL, head);
                    if (head) {
                        const int gid = ng + __popc(m & lt_mask());
                        grs[gid] = (u16)q;
                        if (N + gid < F.Ncap) g_t[N + gid] = (u16)t;
                    }
                    ng += __popc(m);
                }
                if (lane == 0) grs[ng] = (u16)k;
                if (N + ng > F.Ncap) { ovf = true; break; }
                __syncwarp();

                // ---- K2: gene score of every (clade, locus i) group; masks; per-locus max ----
                u64 mxb = dbits(0.0);
                const double dn = (double)llen;
#pragma unroll 1
                for (int base = 0; base < ng; base += 32) {
                    const int j = base + lane;
                    const bool act = j < ng;
                    int rs = 0, re = 0, t = 0;
                    if (act) {
                        rs = grs[j];
                        re = grs[j + 1];
                        t = g_t[N + j];
                    }
                    // with "assign-unknown" the hits of a taxon NAMED Unknown carry no gene score: the spiked row
                    // replaces gene_scores["Unknown"] (waafle_orgscorer.py:416-418)
                    const bool scored = act && t != t_unk;
                    double sc = 0.0;
                    if (scored) {
                        double sum;
                        if (re - rs == 1) {
                            const int r = bord[rs];
                            const u32 ab = bab[r];
                            sum = bv[r] * (double)((int)(ab >> 16) - (int)(ab & 0xffffu));
                        } else {
                            sum = group_integral(bv, bab, bord, rs, re);
                        }
                        sc = sum / dn;
                    }
                    // guard band: a score this close to a threshold is recomputed in numpy's summation order
                    const bool near = scored && (fabs(sc - thr3[0]) <= GUARD || fabs(sc - thr3[1]) <= GUARD ||
                                                 fabs(sc - thr3[2]) <= GUARD);
                    u32 nm = __ballot_sync(FULL, near);
#pragma unroll 1
                    while (nm) {
                        const int ln = __ffs(nm) - 1;
                        nm &= nm - 1;
                        if (lane == ln) {
                            sc = group_mean_exact(a, scratch, F.Kcap, bv, bab, bord, rs, re, llen);
                            ++st_refined;
                        }
                        __syncwarp();
                    }
                    if (act) g_score[N + j] = sc;
                    if (scored) {
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
                N += ng;
                __syncwarp();
            }
            ovf = __any_sync(FULL, ovf);
            if (ovf) { fallback = true; break; }
            if (lane == 0) goff[G] = (u16)N;
            st_groups += (unsigned long long)N;
            __syncwarp();

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

            FLevel L{G, T, t_unk, um, nun, goff, g_t, g_score, unk_row, cl_id, {mk0, mk1, mk2}, l_len};

            // ---- K6: one-clade search (explain_one :585-597) ----
            {
                u64 bbits = 0;
                int bid = -1, btl = -1;
                double bcrit = 0.0;
#pragma unroll 1
                for (int t = lane; t < T; t += 32) {
                    if ((mk0[t] & um) != um) continue;   // crit >= k1
                    double crit;
                    const double rank = row_stats(L, t, -1, &crit);
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
                            const double rank = t == tb ? brank : row_stats(L, t, -1, &crit);
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
                u16 *cand = cur;   // the multisplit cursors are dead
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
                st_ptest += (unsigned long long)NP;
                // survivors of the mask prefilter live in the (dead) record buffer
                u16 *s_a = reinterpret_cast<u16 *>(bv), *s_b = s_a + F.Scap;
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
                if (nsurv > F.Scap) { fallback = true; break; }
                st_pscore += (unsigned long long)nsurv;
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
                            // distinct melded clades per side, ascending node index
                            int *klist = hkey;
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

        if (trip) ++st_trips;
        if (fallback || trip || !finished) {
            if (lane == 0) {
                const unsigned long long s = atomicAdd(&a.ctr->n_fallback, 1ull);
                a.fb_list[s] = (int)c;
                // placeholder record (the speculative compaction walks every contig): overwritten by the exact pipeline
                a.o.call[c] = WFL_CALL_UNCLASSIFIED;
                a.o.n_mem_a[c] = a.o.n_mem_b[c] = 0;
                a.o.mem_pos[c] = 0;
            }
        } else {
            if (lane == 0) write_result_fast(a, c, R);
            ++st_done;
        }
        __syncwarp();
    }
    if (lane == 0) {
        if (st_pairs) atomicAdd(&a.ctr->matched_pairs, st_pairs);
        if (st_groups) atomicAdd(&a.ctr->groups, st_groups);
        if (st_levels) atomicAdd(&a.ctr->levels, st_levels);
        if (st_ptest) atomicAdd(&a.ctr->pairs_tested, st_ptest);
        if (st_pscore) atomicAdd(&a.ctr->pairs_scored, st_pscore);
        if (st_done) atomicAdd(&a.ctr->smem_contigs, st_done);
        if (st_trips) atomicAdd(&a.ctr->guard_trips, st_trips);
    }
    // st_refined lives on whichever lane refined: reduce over the warp
    {
        unsigned long long r = st_refined;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
        if (lane == 0 && r) atomicAdd(&a.ctr->refined_groups, r);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// Slice layout for given capacities; returns the slice size in bytes (multiple of 16).
int fast_layout(FastCfg &F, int Kcap, int Ccap, int Tcap, int Ncap, bool annotations) {
    auto al = [](int x) { return (x + 15) & ~15; };
    int o = 0;
    F.Kcap = Kcap; F.cmask = Ccap - 1; F.Tcap = Tcap; F.Ncap = Ncap;
    F.o_llo = o; o += al(4 * GMAX);
    F.o_llen = o; o += al(4 * GMAX);
    F.o_lraw = o; o += al(4 * GMAX);
    F.o_lstr = o; o += al(GMAX);
    F.o_goff = o; o += al(2 * (GMAX + 2));
    F.o_maxv = o; o += al(8 * GMAX);
    F.o_unk = o; o += al(8 * GMAX);
    const int buf0 = o;
    F.o_bv = o; o += al(8 * Kcap);
    F.o_bab = o; o += al(4 * Kcap);
    F.o_bt = o; o += al(2 * Kcap);
    // survivors of the two-clade prefilter reuse [o_bv, o_bcl): 12 bytes each
    F.Scap = ((o - buf0) / 12) & ~1;
    F.o_bcl = o; o += al(4 * Kcap + 8);   // int clade ids, then u16 bord[Kcap] + u16 grs[Kcap + 2]
    F.o_bh = o; o += annotations ? al(4 * Kcap) : 0;
    F.o_hkey = o; o += al(4 * std::max(Ccap, 2 * Tcap));   // also the melded-member lists (2 x Tcap ints)
    F.o_hval = o; o += al(2 * Ccap);                       // also memA / memB (2 x Tcap bytes)
    F.o_clid = o; o += al(4 * Tcap);
    F.o_mk0 = o; o += al(4 * Tcap);
    F.o_mk1 = o; o += al(4 * Tcap);
    F.o_mk2 = o; o += al(4 * Tcap);
    F.o_cur = o; o += al(2 * (Tcap + 2));
    F.o_gscore = o; o += al(8 * Ncap);
    F.o_gt = o; o += al(2 * Ncap);
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
