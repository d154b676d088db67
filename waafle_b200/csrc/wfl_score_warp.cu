// The per-contig scoring and clade-assignment kernel, revision 2 (sm_100a): ONE WARP PER CONTIG.
//
// Every CTA is a single warp that pulls contigs from a global work queue and runs the whole
// reference block waafle/waafle_orgscorer.py:952-960 for each of them with warp-synchronous
// primitives only (ballot / match / shuffle; no block barriers, no comparison sorts):
//
//   K1  hit x locus matching, records emitted locus-major by ballot ranking
//                                       attach_hits / calc_overlap (waafle_orgscorer.py:359-369,
//                                       559-564; utils.py:487-500)
//   K3  annotation arg-max              waafle_orgscorer.py:384-392
//   K5  taxonomy lift on the contig's distinct-clade table (parent gather + dedupe), records
//       regrouped by one stable warp multisplit          waafle_orgscorer.py:431-445
//   K2  envelope integral per (clade, locus) group in numpy's pairwise order, driven by a
//       per-locus leaf plan shared by all clades     waafle_orgscorer.py:371-382,399-406
//   K4  weak-loci mask / Unknown spike  waafle_orgscorer.py:407-429
//   K6  one-clade search                waafle_orgscorer.py:447-461,495-509,585-597,621-631
//   K7  two-clade pair search on gene bitmasks    waafle_orgscorer.py:511-545,599-619
//   K8  ranking / meld / LGT filters    waafle_orgscorer.py:633-744, utils.py:401-411
//   K9  level loop                      waafle_orgscorer.py:566-583
//
// Per-contig arrays come from a bump arena: the CTA's dynamic shared memory first, then a per-CTA
// global slab (L2-resident), so 2-20-gene contigs live in shared memory and 100-kb contigs still
// run.  fp64 throughout; the gene-score sums reproduce numpy's pairwise summation bit for bit
// (constant leaves cost one memoised add chain per envelope run, leaves containing a breakpoint
// are evaluated per accumulator column).
#include "wfl_warp_common.cuh"

namespace wfl {

// ---------------------------------------------------------------------------------------------
// the kernel: blockDim.x == 32
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(32) wfl_score_contigs_warp(const ScoreArgs a) {
    extern __shared__ __align__(16) char smem_dyn[];
    const int lane = threadIdx.x;
    const DevParams &P = a.P;
    const DevTax &tax = a.t;
    const int S = P.p.n_systems;
    const bool spike = P.p.weak_loci == 2;

#pragma unroll 1
    for (;;) {
        long long c = -1;
        if (lane == 0) {
            unsigned long long w = atomicAdd(&a.ctr->next_work, 1ull);
            c = (long long)w < a.n_work ? (a.work_list ? a.work_list[w] : a.work_base + (long long)w) : -1;
        }
        c = __shfl_sync(FULL, c, 0);
        if (c < 0) break;

        long long ph_last = clock64();
        unsigned long long ph[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PH(i) do { if (lane == 0) { long long t_ = clock64(); ph[i] += (unsigned long long)(t_ - ph_last); ph_last = t_; } } while (0)

        const long long h0 = a.b.hit_off[c], l0 = a.b.locus_off[c];
        const int H = (int)(a.b.hit_off[c + 1] - h0), Graw = (int)(a.b.locus_off[c + 1] - l0);
        Arena ar{smem_dyn, a.slab + (size_t)blockIdx.x * a.slab_bytes, (size_t)a.smem_bytes, a.slab_bytes,
                 0, 0, true, true};
        unsigned long long need_hint = 64ull * Graw + (1ull << 12);

        // ---- loci: --min-gene-length filter, GFF order kept (waafle_orgscorer.py:348-352) ----
        int *l_lo = ar.get<int>(Graw), *l_len = ar.get<int>(Graw), *l_raw = ar.get<int>(Graw);
        int *l_base = ar.get<int>(Graw + 2);   // record range of each locus (locus-major records)
        const u16 **l_plan = ar.get<const u16 *>(Graw + 1);   // leaf plan of each locus
        int *l_nleaf = ar.get<int>(Graw + 1);                 // leaves in it (fallback offset meanwhile)
        signed char *l_str = ar.get<signed char>(Graw);
        u32 *l_k8 = ar.get<u32>(Graw + 1);                     // packed m/8 values for the S_k memo
        bool overflow = !ar.ok;
        int G = 0;
#pragma unroll 1
        for (int base = 0; base < Graw && !overflow; base += 32) {
            int j = base + lane, flag = 0, lo = 0, len = 0;
            if (j < Graw) {
                int s = a.b.locus_start[l0 + j], e = a.b.locus_end[l0 + j];
                lo = min(s, e);
                len = max(s, e) - lo + 1;
                flag = (double)len >= P.p.min_gene_length;
                a.o.locus_flags[l0 + j] = flag ? WFL_LOCUS_RETAINED : 0;
                a.o.synteny[l0 + j] = 0;
#pragma unroll 1
                for (int s2 = 0; s2 < S; ++s2) a.o.ann_winner[(l0 + j) * S + s2] = -1;
            }
            u32 m = __ballot_sync(FULL, flag);
            if (flag) {
                int pos = G + __popc(m & lt_mask());
                l_lo[pos] = lo;
                l_len[pos] = len;
                l_raw[pos] = j;
                l_str[pos] = a.b.locus_strand[l0 + j];
            }
            G += __popc(m);
        }
        __syncwarp();
        const int W = (G + 63) >> 6;
        int lifts = H > 0 ? P.p.jump_taxonomy : 0;
        int r_call = WFL_CALL_UNCLASSIFIED, r_dir = 0, r_c1 = -1, r_c2 = -1, r_lca = -1, r_b1 = -1, r_b2 = -1,
            r_na = 0, r_nb = 0, r_status = 0;
        long long r_mem = 0;
        double r_crit = 0.0, r_rank = 0.0;
        bool bad_input = false;

        if (!overflow && H > 0 && G > 0) {
            // ---- leaf plans: host-built table for lengths <= plan_nmax, in-kernel build beyond -------
#pragma unroll 1
            for (int i = lane; i <= G + 1; i += 32) l_base[i] = 0;
            int np_tot = 0;
#pragma unroll 1
            for (int base = 0; base < G; base += 32) {
                int i = base + lane, np = (i < G && l_len[i] > a.plan_nmax) ? plan_cap(l_len[i]) : 0, tot;
                int ex = warp_excl_scan(np, tot);
                if (i < G) l_nleaf[i] = np_tot + ex;
                np_tot += tot;
            }
            __syncwarp();
            // ---- K1 pass 1: matches per locus (ballot counts) --------------------------------------
            const bool all_match = P.p.min_overlap <= 0.0;   // disjoint pairs "overlap" 0 >= min_overlap
#pragma unroll 1
            for (int base = 0; base < H; base += 32) {
                int h = base + lane;
                bool hok = false;
                int hmin = 0, hmax = 0;
                signed char hs = 0;
                if (h < H) {
                    hok = a.b.hit_scov[h0 + h] >= P.p.min_scov;   // waafle_orgscorer.py:362
                    int q1 = a.b.hit_qstart[h0 + h], q2 = a.b.hit_qend[h0 + h];
                    hmin = min(q1, q2);
                    hmax = max(q1, q2);
                    hs = a.b.hit_strand[h0 + h];
                }
                if (!__any_sync(FULL, hok)) continue;
#pragma unroll 1
                for (int i = 0; i < G; ++i) {
                    const int lmin = l_lo[i], llen = l_len[i];
                    bool cand = hok && (all_match || !(lmin > hmax || hmin > lmin + llen - 1));
                    if (!__any_sync(FULL, cand)) continue;
                    u32 m = __ballot_sync(FULL, cand && hit_matches(P, hmin, hmax, hs, lmin, llen, l_str[i]));
                    if (lane == 0 && m) l_base[i + 1] += __popc(m);
                }
            }
            __syncwarp();
            int M = 0;
#pragma unroll 1
            for (int base = 0; base < G; base += 32) {   // exclusive scan -> record base of each locus
                int i = base + lane, cnt = i < G ? l_base[i + 1] : 0, tot;
                int ex = warp_excl_scan(cnt, tot);
                __syncwarp();
                if (i < G) l_base[i + 1] = M + ex;   // cursor of locus i lives in l_base[i+1] during the fill
                M += tot;
            }
            __syncwarp();
            // worst case for this contig (groups <= M + G, clades <= groups + 1): one replay suffices
            need_hint = 96ull * Graw + 2ull * np_tot + 64ull * (unsigned long long)M +
                        48ull * ((unsigned long long)M + G + 2) +
                        (80ull + 24ull * W) * ((unsigned long long)M + G + 2) + 16ull * (2 * M + 64) +
                        (unsigned long long)G * (64 + 16 * S) + (1ull << 12);

            // ---- record arrays (locus-major) ------------------------------------------------------
            u16 *plan_fb = ar.get<u16>(np_tot);
            double *r_v = ar.get<double>(M);
            int *r_a = ar.get<int>(M), *r_b = ar.get<int>(M), *r_t = ar.get<int>(M), *r_loc = ar.get<int>(M),
                *r_hit = S > 0 ? ar.get<int>(M) : nullptr;
            int *base_ord = ar.get<int>(M);   // records in (locus, score descending) order
            double *maxv = ar.get<double>(G);
            u64 *maxb = ar.get<u64>(G);
            u8 *ign = ar.get<u8>(G + 1);
            u64 *um = ar.get<u64>(W);
            u64 *annb = S > 0 ? ar.get<u64>((size_t)G * S) : nullptr;
            int *annw = S > 0 ? ar.get<int>((size_t)G * S) : nullptr;
            overflow = !ar.ok;
            if (!overflow) {
#pragma unroll 1
                for (int i = lane; i < G; i += 32) {
                    const int len = l_len[i];
                    if (len <= a.plan_nmax) {
                        const PlanEntry pe = a.plan_index[len];
                        if (a.plan_tree != nullptr) {   // node-size table (tree walk); nleaf < 0 flags it
                            l_plan[i] = reinterpret_cast<const u16 *>(a.plan_tree + pe.toff);
                            l_nleaf[i] = -(int)pe.nsz;
                        } else {
                            l_plan[i] = a.plan_data + pe.off;
                            l_nleaf[i] = (int)pe.nleaf;
                        }
                        l_k8[i] = pe.k8;
                    } else {
                        u16 *dst = plan_fb + l_nleaf[i];
                        l_plan[i] = dst;
                        l_nleaf[i] = build_plan(len, dst, reinterpret_cast<u8 *>(&l_k8[i]));
                    }
                }
#pragma unroll 1
                for (int i = lane; i < G * S; i += 32) { annb[i] = 0; annw[i] = -1; }
                __syncwarp();
                // ---- K1 pass 2: emit records (score_hit, waafle_orgscorer.py:371-382) -------------
#pragma unroll 1
                for (int base = 0; base < H; base += 32) {
                    int h = base + lane;
                    bool hok = false;
                    int hmin = 0, hmax = 0, cl = 0;
                    signed char hs = 0;
                    double sc = 0.0;
                    u32 sys = 0;
                    if (h < H) {
                        hok = a.b.hit_scov[h0 + h] >= P.p.min_scov;
                        int q1 = a.b.hit_qstart[h0 + h], q2 = a.b.hit_qend[h0 + h];
                        hmin = min(q1, q2);
                        hmax = max(q1, q2);
                        hs = a.b.hit_strand[h0 + h];
                    }
                    if (!__any_sync(FULL, hok)) continue;
                    if (hok) {
                        cl = a.b.hit_taxon[h0 + h];
                        if ((u32)cl >= (u32)tax.n_nodes) { cl = tax.root; bad_input = true; }
#pragma unroll 1
                        for (int j = 0; j < P.p.jump_taxonomy; ++j) cl = tax.parent[cl];
                        sc = a.b.hit_score[h0 + h];
                        if (S > 0) sys = a.b.hit_sysmask[h0 + h];
                    }
#pragma unroll 1
                    for (int i = 0; i < G; ++i) {
                        const int lmin = l_lo[i], len = l_len[i];
                        bool cand = hok && (all_match || !(lmin > hmax || hmin > lmin + len - 1));
                        if (!__any_sync(FULL, cand)) continue;
                        bool mt = cand && hit_matches(P, hmin, hmax, hs, lmin, len, l_str[i]);
                        u32 m = __ballot_sync(FULL, mt);
                        if (!m) continue;
                        int cur = l_base[i + 1];
                        if (mt) {
                            int slot = cur + __popc(m & lt_mask());
                            // python slice [h1 : h2+1] of a length-len array (:376-382)
                            int s1 = max(0, hmin - lmin), e1 = min(len - 1, hmax - lmin) + 1;
                            if (e1 < 0) e1 = max(0, e1 + len);
                            s1 = min(s1, len);
                            if (e1 < s1) e1 = s1;
                            r_v[slot] = sc;
                            r_a[slot] = s1;
                            r_b[slot] = e1;
                            r_t[slot] = cl;
                            r_loc[slot] = i;
                            if (S > 0) {
                                r_hit[slot] = h;
                                // K3 phase 1: max annotated score per (locus, system)
                                u32 ms = sys;
                                if (ms && sc >= P.ann_thr) {
                                    u64 sb = dbits(sc);
                                    while (ms) {
                                        int s2 = __ffs(ms) - 1;
                                        ms &= ms - 1;
                                        atomicMax(&annb[(size_t)i * S + s2], sb);
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) l_base[i + 1] = cur + __popc(m);
                        __syncwarp();
                    }
                }
                __syncwarp();
                // after the fill, l_base[i+1] == end of locus i == start of locus i+1; l_base[0] == 0
                if (S > 0) {
                    // K3 phase 2: the LAST hit (file order) attaining the max wins (:389, '>=')
#pragma unroll 1
                    for (int r = lane; r < M; r += 32) {
                        u32 ms = a.b.hit_sysmask[h0 + r_hit[r]];
                        double sc = r_v[r];
                        if (ms && sc >= P.ann_thr) {
                            u64 sb = dbits(sc);
                            while (ms) {
                                int s2 = __ffs(ms) - 1;
                                ms &= ms - 1;
                                if (annb[(size_t)r_loc[r] * S + s2] == sb)
                                    atomicMax(&annw[(size_t)r_loc[r] * S + s2], r_hit[r]);
                            }
                        }
                    }
                    __syncwarp();
#pragma unroll 1
                    for (int i = lane; i < G * S; i += 32) {
                        int w = annw[i];
                        a.o.ann_winner[(l0 + l_raw[i / S]) * S + (i % S)] = w >= 0 ? (int)(h0 + w) : -1;
                    }
                }
                bool have_base_ord = false;
                PH(0);

                // ---- distinct clades of the contig: hash-dedupe + rank by counting ----------------
                // cl_id[0..T) ascending; r_t[r] becomes the rank of the record's clade.  Lifts then
                // work on this table (T parent gathers), not on the records.
                int T = 0;
                int *cl_id = ar.get<int>(M + 2);   // capacity for all levels (T never grows)
                int *ord = ar.get<int>(M);
                int *map_t = ar.get<int>(M + 2);
                u8 *fo = ar.get<u8>(M + 2);
                const size_t mark_smem = ar.smem_used, mark_slab = ar.slab_used;
                const int nwords = (tax.n_nodes + 31) >> 5;
                if (nwords <= BITMAP_MAX_WORDS) {
                    // small taxonomies: presence bitmap over the node indices; the rank of a clade among the
                    // contig's clades is a prefix popcount (ascending node index == ascending name)
                    u32 *bm = ar.get<u32>(nwords);
                    int *bpre = ar.get<int>(nwords + 1);
                    if (!ar.ok) overflow = true;
                    if (!overflow) {
#pragma unroll 1
                        for (int w = lane; w < nwords; w += 32) bm[w] = 0;
                        __syncwarp();
#pragma unroll 1
                        for (int r = lane; r < M + (spike ? 1 : 0); r += 32) {
                            const int key = r < M ? r_t[r] : tax.unknown;
                            atomicOr(&bm[key >> 5], 1u << (key & 31));
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int base = 0; base < nwords; base += 32) {
                            const int w = base + lane;
                            const u32 bits = w < nwords ? bm[w] : 0u;
                            int tot, ex = warp_excl_scan(__popc(bits), tot);
                            if (w < nwords) {
                                int pos = T + ex;
                                bpre[w] = pos;
                                u32 b = bits;
                                while (b) {
                                    cl_id[pos++] = (w << 5) + __ffs(b) - 1;
                                    b &= b - 1;
                                }
                            }
                            T += tot;
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int r = lane; r < M; r += 32) {
                            const int key = r_t[r];
                            r_t[r] = bpre[key >> 5] + __popc(bm[key >> 5] & ((1u << (key & 31)) - 1u));
                        }
                        __syncwarp();
                    }
                } else {
                    int cap = 64;
                    while (cap < 2 * (M + 1)) cap <<= 1;
                    int *hk = ar.get<int>(cap), *hv = ar.get<int>(cap);
                    int *dl = ar.get<int>(M + 1);
                    if (!ar.ok) overflow = true;
                    if (!overflow) {
#pragma unroll 1
                        for (int i = lane; i < cap; i += 32) hk[i] = -1;
                        __syncwarp();
#pragma unroll 1
                        for (int r = lane; r < M + (spike ? 1 : 0); r += 32) {
                            int key = r < M ? r_t[r] : tax.unknown;
                            u32 slot = ((u32)key * 2654435761u) & (cap - 1);
#pragma unroll 1
                            for (;;) {
                                int old = atomicCAS(&hk[slot], -1, key);
                                if (old == -1 || old == key) break;
                                slot = (slot + 1) & (cap - 1);
                            }
                            if (r < M) r_t[r] = (int)slot;
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int base = 0; base < cap; base += 32) {   // compact the distinct keys
                            int i = base + lane;
                            bool f = hk[i] >= 0;
                            u32 m = __ballot_sync(FULL, f);
                            if (f) dl[T + __popc(m & lt_mask())] = i;
                            T += __popc(m);
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int j = lane; j < T; j += 32) map_t[j] = hk[dl[j]];   // compact key list
                        __syncwarp();
#pragma unroll 1
                        for (int j = lane; j < T; j += 32) {   // rank = #distinct keys below
                            int key = map_t[j], rk = 0;
#pragma unroll 4
                            for (int q = 0; q < T; ++q) rk += map_t[q] < key;
                            hv[dl[j]] = rk;
                            cl_id[rk] = key;
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int r = lane; r < M; r += 32) r_t[r] = hv[r_t[r]];
                        __syncwarp();
                    }
                }
                PH(1);

                // ---- K9: level loop (evaluate_contig, waafle_orgscorer.py:566-583) ----------------
                int n_levels = 0;
                long long n_groups = 0, n_ptest = 0, n_pscore = 0;
#pragma unroll 1
                for (int iter = 0; !overflow; ++iter) {
                    ar.smem_used = mark_smem;
                    ar.slab_used = mark_slab;
                    ++n_levels;
                    if (!have_base_ord) {
                        // ---- level-invariant base order (built once):
                        // inside each locus, records by descending score
                        // (rank by counting; ties keep emission order).  Every later regrouping is a STABLE
                        // split of this sequence, so each (clade, locus) group arrives score-descending and
                        // the envelope scan can stop at the first record covering a site.
#pragma unroll 1
                        for (int r = lane; r < M; r += 32) {
                            const int loc = r_loc[r], s0 = l_base[loc], s1 = l_base[loc + 1];
                            const double v = r_v[r];
                            int rk = 0;
#pragma unroll 4
                            for (int q = s0; q < s1; ++q) {
                                double vq = r_v[q];
                                rk += (vq > v) || (vq == v && q < r);
                            }
                            base_ord[s0 + rk] = r;
                        }
                        __syncwarp();
                        have_base_ord = true;
                    }
                    // ---- regroup: stable multisplit of the locus-major records by clade rank ------
                    int *cur = ar.get<int>(T + 2);
                    int *s_a = ar.get<int>(M), *s_b = ar.get<int>(M);
                    double *s_v = ar.get<double>(M);
                    if (!ar.ok) { overflow = true; break; }
                    warp_multisplit(M, T, r_t, have_base_ord ? base_ord : nullptr, cur, ord);
                    __syncwarp();
                    // slices and scores copied into group order: the envelope scans read them contiguously
#pragma unroll 1
                    for (int r = lane; r < M; r += 32) {
                        const int i = ord[r];
                        s_a[r] = r_a[i];
                        s_b[r] = r_b[i];
                        s_v[r] = r_v[i];
                    }
                    // groups = maximal runs of equal (clade rank, locus) in ord
                    int t_unk = -1;
                    if (spike) {
                        int lo = 0, hi = T;   // cl_id ascending: binary search for Unknown
                        while (lo < hi) {
                            int mid = (lo + hi) >> 1;
                            if (cl_id[mid] < tax.unknown) lo = mid + 1; else hi = mid;
                        }
                        t_unk = lo;   // present by construction (inserted before ranking / kept by lifts)
                    }
                    int ng = 0, nlt = 0, nu = 0;
#pragma unroll 1
                    for (int base = 0; base < M; base += 32) {
                        int r = base + lane;
                        bool f = false;
                        int t = 0;
                        if (r < M) {
                            int i = ord[r];
                            t = r_t[i];
                            f = r == 0 || t != r_t[ord[r - 1]] || r_loc[i] != r_loc[ord[r - 1]];
                        }
                        ng += __popc(__ballot_sync(FULL, f));
                        if (spike) {
                            nlt += __popc(__ballot_sync(FULL, f && t < t_unk));
                            nu += __popc(__ballot_sync(FULL, f && t == t_unk));
                        }
                    }
                    const int Ngrp = spike ? ng - nu + G : ng;
                    n_groups += Ngrp;
                    double *g_score = ar.get<double>(Ngrp);
                    int *g_rs = ar.get<int>(Ngrp + 1), *g_re = ar.get<int>(Ngrp + 1), *g_loc = ar.get<int>(Ngrp),
                        *g_t = ar.get<int>(Ngrp), *gs = ar.get<int>(ng + 1),
                        *g_perm = ar.get<int>(Ngrp), *gcur = ar.get<int>(2 * G + 2);
                    if (!ar.ok) { overflow = true; break; }
                    int gbase = 0;
#pragma unroll 1
                    for (int base = 0; base < M; base += 32) {
                        int r = base + lane;
                        bool f = false;
                        int t = 0, loc = 0;
                        if (r < M) {
                            int i = ord[r];
                            t = r_t[i];
                            loc = r_loc[i];
                            f = r == 0 || t != r_t[ord[r - 1]] || loc != r_loc[ord[r - 1]];
                        }
                        u32 m = __ballot_sync(FULL, f);
                        if (f) {
                            int gid = gbase + __popc(m & lt_mask()), dst = gid;
                            if (spike) dst = t < t_unk ? gid : (t == t_unk ? -1 : gid - nu + G);
                            gs[gid] = r;
                            if (dst >= 0) {
                                g_rs[dst] = r;
                                g_re[dst] = gid;   // patched to the end position below
                                g_loc[dst] = loc;
                                g_t[dst] = t;
                            }
                        }
                        gbase += __popc(m);
                    }
                    if (lane == 0) gs[ng] = M;
                    __syncwarp();
#pragma unroll 1
                    for (int g = lane; g < Ngrp; g += 32)
                        if (!(spike && g >= nlt && g < nlt + G)) g_re[g] = gs[g_re[g] + 1];
                    if (spike)
                        for (int i = lane; i < G; i += 32) {
                            g_rs[nlt + i] = -1;
                            g_loc[nlt + i] = i;
                            g_t[nlt + i] = t_unk;
                        }
#pragma unroll 1
                    for (int i = lane; i < G; i += 32) maxb[i] = dbits(0.0);
                    __syncwarp();
                    PH(2);
                    // ---- K2: envelope integral per group, numpy-pairwise-exact ----------------------
                    // groups are visited multi-record groups first, then single-record ones, locus-major inside
                    // each class: lanes of a warp then walk the same leaf plan and meet groups of similar
                    // envelope complexity (leaves with several run boundaries are the expensive, divergent part)
                    {
                        int *gkey = reinterpret_cast<int *>(g_score);   // g_score is written only below
#pragma unroll 1
                        for (int g = lane; g < Ngrp; g += 32)
                            gkey[g] = g_loc[g] + ((g_rs[g] >= 0 && g_re[g] - g_rs[g] >= 2) ? 0 : G);
                        __syncwarp();
                        warp_multisplit(Ngrp, 2 * G, gkey, nullptr, gcur, g_perm);
                    }
                    __syncwarp();
#if WFL_K2_CONV
#pragma unroll 1
                    for (int base = 0; base < Ngrp; base += 32) {
                        // all 32 lanes call group_mean_warp together (it re-converges them before every leaf)
                        const int gi = base + lane;
                        int g = 0, rs = -1, re = 0, t = 0, loc = 0, nleaf = 0;
                        if (gi < Ngrp) {
                            g = g_perm[gi];
                            rs = g_rs[g];
                            if (rs >= 0) {
                                t = g_t[g];
                                loc = g_loc[g];
                                re = g_re[g];
                                nleaf = l_nleaf[loc];
                            }
                        }
                        double sc;
                        if (a.plan_tree != nullptr) {   // experimental tree walk: per lane
                            sc = rs >= 0 ? group_mean(s_a, s_b, s_v, rs, re, l_len[loc], have_base_ord, l_k8[loc],
                                                      l_plan[loc], nleaf) : 0.0;
                        } else {
                            sc = group_mean_warp(s_a, s_b, s_v, max(rs, 0), re, l_len[loc], have_base_ord, l_k8[loc],
                                                 l_plan[loc], rs >= 0 ? nleaf : 0);
                        }
                        if (rs >= 0) {
                            g_score[g] = sc;
                            if (cl_id[t] != tax.unknown)   // waafle_orgscorer.py:409-411
                                atomicMax(&maxb[loc], dbits(sc));
                        }
                    }
#else
#pragma unroll 1
                    for (int base = 0; base < Ngrp; base += 32) {
                        int gi = base + lane;
                        if (gi < Ngrp) {
                            int g = g_perm[gi];
                            int rs = g_rs[g];
                            if (rs >= 0) {
                                const int t = g_t[g], loc = g_loc[g];
                                const int re = g_re[g];
                                double sc = group_mean(s_a, s_b, s_v, rs, re, l_len[loc], have_base_ord, l_k8[loc],
                                                       l_plan[loc], l_nleaf[loc]);
                                g_score[g] = sc;
                                if (cl_id[t] != tax.unknown)   // waafle_orgscorer.py:409-411
                                    atomicMax(&maxb[loc], dbits(sc));
                            }
                        }
                    }
#endif
                    __syncwarp();
                    PH(3);
                    // ---- K4: weak loci (waafle_orgscorer.py:412-427) ---------------------------------
#pragma unroll 1
                    for (int i = lane; i < G; i += 32) {
                        double mx = dbits_inv(maxb[i]);
                        maxv[i] = mx;
                        ign[i] = (P.p.weak_loci == 0) ? !(mx >= P.min_thr) : 0;
                        if (spike) g_score[nlt + i] = 1.0 - mx;
                    }
                    if (lane == 0) ign[G] = 0;   // sentinel for ScoreSrc
                    __syncwarp();
                    int nun = 0;
#pragma unroll 1
                    for (int w = lane; w < W; w += 32) {
                        u64 m = 0;
#pragma unroll 1
                        for (int b = 0; b < 64 && w * 64 + b < G; ++b)
                            if (!ign[w * 64 + b]) m |= 1ull << b;
                        um[w] = m;
                        nun += __popcll(m);
                    }
#pragma unroll 1
                    for (int i = lane; i < G; i += 32)
                        a.o.locus_flags[l0 + l_raw[i]] = WFL_LOCUS_RETAINED | (ign[i] ? WFL_LOCUS_IGNORED : 0);
                    nun = warp_sum(nun);
                    if (iter == 0 && c == a.dbg_contig) {
#pragma unroll 1
                        for (int g = lane; g < Ngrp; g += 32)
                            if (g < a.dbg_cap) {
                                a.dbg_clade[g] = cl_id[g_t[g]];
                                a.dbg_locus[g] = g_loc[g];
                                a.dbg_score[g] = g_score[g];
                            }
                        if (lane == 0) *a.dbg_count = Ngrp;
                    }
                    if (iter == 0 && nun == 0) break;   // "empty" contig, waafle_orgscorer.py:959

                    // ---- clade rows + gene bitmasks ----------------------------------------------------
                    // every rank in [0, T) owns >= 1 group (the spiked Unknown owns G)
                    int *cl_go = ar.get<int>(T + 1), *cand = ar.get<int>(T);
                    double *cl_rank = ar.get<double>(T), *cl_crit = ar.get<double>(T);
                    u8 *cl_opt = ar.get<u8>(T), *memA = ar.get<u8>(T), *memB = ar.get<u8>(T);
                    u64 *mk0 = ar.get<u64>((size_t)T * W), *mk1 = ar.get<u64>((size_t)T * W),
                        *mk2 = ar.get<u64>((size_t)T * W);
                    u64 *bestm = ar.get<u64>(3 * (size_t)W);
                    int *cl_par = ar.get<int>(T);   // parent of each listed clade of the contig, -1 if unlisted
                    Level *Lp = ar.get<Level>(1);
                    if (!ar.ok) { overflow = true; break; }
#pragma unroll 1
                    for (int g = lane; g < Ngrp; g += 32)
                        if (g == 0 || g_t[g] != g_t[g - 1]) cl_go[g_t[g]] = g;
                    if (lane == 0) cl_go[T] = Ngrp;
                    int hasroot = 0;
#pragma unroll 1
                    for (int t = lane; t < T; t += 32) hasroot |= cl_id[t] == tax.root;
                    hasroot = __any_sync(FULL, hasroot);
                    __syncwarp();
#pragma unroll 1
                    for (int t = lane; t < T; t += 32) {
                        memA[t] = memB[t] = 0;
                        cl_par[t] = tax.listed[cl_id[t]] ? tax.parent[cl_id[t]] : -1;
#pragma unroll 1
                        for (int q = 0; q < 3; ++q) {
                            u64 *m = (q == 0 ? mk0 : q == 1 ? mk1 : mk2) + (size_t)t * W;
                            const double thr = q == 0 ? P.p.k1 : q == 1 ? P.p.k2 : 1e-6;
                            // a locus without an entry scores 0 (waafle_orgscorer.py:404-405)
#pragma unroll 1
                            for (int w = 0; w < W; ++w) {
                                int nb = min(64, G - w * 64);
                                m[w] = thr <= 0.0 ? (nb == 64 ? ~0ull : ((1ull << nb) - 1)) : 0ull;
                            }
#pragma unroll 1
                            for (int g = cl_go[t]; g < cl_go[t + 1]; ++g) {
                                int loc = g_loc[g];
                                u64 bit = 1ull << (loc & 63);
                                if (g_score[g] >= thr) m[loc >> 6] |= bit;
                                else m[loc >> 6] &= ~bit;
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) *Lp = Level{G, W, T, Ngrp, nun, g_loc, g_score, cl_id, cl_go, {mk0, mk1, mk2}, um, ign, l_len, cl_par};
                    __syncwarp();
                    const Level &L = *Lp;
                    PH(4);

                    // ---- K6: one-clade search (explain_one, waafle_orgscorer.py:585-597) -----------
                    u64 bbits = 0;
#pragma unroll 1
                    for (int t = lane; t < T; t += 32) {
                        bool pass = true;
#pragma unroll 1
                        for (int w = 0; w < W; ++w) pass &= (mk0[(size_t)t * W + w] & um[w]) == um[w];   // crit >= k1
                        cl_opt[t] = pass;
                        if (pass) {
                            double crit, rank;
                            rank = score_clades(Lp, t, -1, &crit);
                            cl_rank[t] = rank;
                            cl_crit[t] = crit;
                            u64 b = dbits(rank);
                            bbits = b > bbits ? b : bbits;
                        }
                    }
                    bbits = warp_max_u64(bbits);
                    __syncwarp();
                    long long bt = -1;
                    if (bbits)
                        for (int t = lane; t < T; t += 32)
                            if (cl_opt[t] && dbits(cl_rank[t]) == bbits) bt = t;   // ties: last in name order
                    bt = warp_max_ll(bt);
                    PH(5);
                    if (bt >= 0) {
                        // meld_one (waafle_orgscorer.py:621-631)
                        const int tb = (int)bt;
                        const double brank = cl_rank[tb];
                        r_call = WFL_CALL_NO_LGT;
                        r_b1 = r_c1 = cl_id[tb];
                        r_crit = cl_crit[tb];
                        r_rank = brank;
                        if (P.p.disambiguate_one == 1) {
                            int my = -1, nk = 0;
#pragma unroll 1
                            for (int t = lane; t < T; t += 32)
                                if (cl_opt[t] && brank - cl_rank[t] <= P.p.range) {
                                    my = lca2(tax, my, cl_id[t]);
                                    memA[t] = 1;
                                    ++nk;
                                }
                            r_c1 = warp_lca(tax, my);
                            r_na = warp_sum(nk);
                        }
#pragma unroll 1
                        for (int i = lane; i < G; i += 32)   // set_synteny_one (:495-509)
                            a.o.synteny[l0 + l_raw[i]] =
                                ign[i] ? '~' : ((mk0[(size_t)tb * W + (i >> 6)] >> (i & 63)) & 1 ? 'A' : '!');
                    } else {
                        // ---- K7: two-clade search (explain_two, waafle_orgscorer.py:599-619) --------
                        int T2 = 0;
#pragma unroll 1
                        for (int base = 0; base < T; base += 32) {
                            int t = base + lane;
                            bool f = false;
                            if (t < T)   // max(gene_scores[clade]) >= k2, unmasked (:603-605)
                                for (int w = 0; w < W; ++w) f |= mk1[(size_t)t * W + w] != 0;
                            u32 m = __ballot_sync(FULL, f);
                            if (f) cand[T2 + __popc(m & lt_mask())] = t;
                            T2 += __popc(m);
                        }
                        __syncwarp();
                        const long long NP = (long long)T2 * (T2 - 1) / 2;
                        n_ptest += NP;
                        // pass 1: mask prefilter over all pairs (crit >= k2 <=> every non-ignored locus is
                        // covered at k2 by one of the two clades); survivors are compacted in pair order
                        size_t room = (ar.smem_cap - ar.smem_used) > (ar.slab_cap - ar.slab_used)
                                          ? (ar.smem_cap - ar.smem_used) : (ar.slab_cap - ar.slab_used);
                        long long scap = (long long)(room / 16) - 8;
                        if (scap > NP) scap = NP;
                        if (scap < 0) scap = 0;
                        int *s_i = ar.get<int>((size_t)scap), *s_j = ar.get<int>((size_t)scap);
                        double *s_rank = ar.get<double>((size_t)scap);
                        if (!ar.ok) { overflow = true; break; }
                        int nsurv = 0;
                        {
                            int pi = 0, po = 0;   // lane's pair: clade cand[pi] with cand[pi + 1 + po]
                            if (lane < NP) {
                                int jj;
                                pair_decode(lane, T2, pi, jj);
                                po = jj - pi - 1;
                            }
#pragma unroll 1
                            for (long long pb = 0; pb < NP; pb += 32) {
                                const bool act = pb + lane < NP;
                                bool pass = false;
                                if (act) pass = pair_pass(L, cand[pi], cand[pi + 1 + po]);   // crit < k2 fails (:610)
                                u32 m = __ballot_sync(FULL, pass);
                                if (pass) {
                                    long long dst = (long long)nsurv + __popc(m & lt_mask());
                                    if (dst < scap) { s_i[dst] = pi; s_j[dst] = pi + 1 + po; }
                                }
                                nsurv += __popc(m);
                                if (act) {   // advance the lane's pair by 32 in i-major order
                                    po += 32;
                                    while (pi < T2 - 1 && po >= T2 - 1 - pi) { po -= T2 - 1 - pi; ++pi; }
                                }
                            }
                        }
                        if (nsurv > scap) {   // survivor list does not fit: replay with a slab sized for all pairs
                            overflow = true;
                            need_hint += 16ull * (unsigned long long)NP + 4096ull;
                            break;
                        }
                        n_pscore += nsurv;
                        __syncwarp();
                        // pass 2: exact crit / rank of the survivors; best rank, ties -> last pair in
                        // (clade1, clade2) iteration order
                        double my_rank = -1.0;
                        long long my_p = -1;
#pragma unroll 1
                        for (int q = lane; q < nsurv; q += 32) {
                            double crit;
                            double rank = score_clades(Lp, cand[s_i[q]], cand[s_j[q]], &crit);
                            s_rank[q] = rank;
                            if (my_p < 0 || rank >= my_rank) { my_rank = rank; my_p = q; }
                        }
                        u64 pb = warp_max_u64(my_p >= 0 ? dbits(my_rank) : 0ull);
                        const long long bp = warp_max_ll((my_p >= 0 && dbits(my_rank) == pb) ? my_p : -1);
                        __syncwarp();
                        if (bp >= 0) {
                            // meld_two (waafle_orgscorer.py:633-669)
                            const int bi = s_i[bp], bj = s_j[bp];
                            TwoEval be;
                            eval_two(L, tax, P, cand[bi], cand[bj], be);
                            double bcrit, brank;
                            brank = score_clades(Lp, cand[bi], cand[bj], &bcrit);
                            const bool bunk = be.c1 == tax.unknown || be.c2 == tax.unknown;
#pragma unroll 1
                            for (int w = lane; w < W; w += 32) {
                                u64 A, Bm, amb;
                                letters(L, L.mk[P.amb_sel], bunk, cand[bi], cand[bj], w, A, Bm, amb);
                                bestm[w] = be.swap ? Bm : A;
                                bestm[W + w] = be.swap ? A : Bm;
                                bestm[2 * W + w] = amb;
                            }
                            __syncwarp();
                            int nk = 0, nbad = 0, ndiff = 0, la = -1, lb = -1;
#pragma unroll 1
                            for (int q = lane; q < nsurv; q += 32) {
                                if (!(brank - s_rank[q] <= P.p.range)) continue;   // :636
                                const int t1 = cand[s_i[q]], t2 = cand[s_j[q]];
                                TwoEval ev;
                                eval_two(L, tax, P, t1, t2, ev);
                                ++nk;
                                nbad += !ev.ok;
                                const bool unk = ev.c1 == tax.unknown || ev.c2 == tax.unknown;
                                bool same = true;   // meld_precheck: same synteny string (:671-676)
#pragma unroll 1
                                for (int w = 0; w < W; ++w) {
                                    u64 A, Bm, amb;
                                    letters(L, L.mk[P.amb_sel], unk, t1, t2, w, A, Bm, amb);
                                    same &= (ev.swap ? Bm : A) == bestm[w] && (ev.swap ? A : Bm) == bestm[W + w] &&
                                            amb == bestm[2 * W + w];
                                }
                                ndiff += !same;
                                la = lca2(tax, la, ev.c1);
                                lb = lca2(tax, lb, ev.c2);
                                memA[ev.t1] = 1;
                                memB[ev.t2] = 1;
                            }
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
                                    int l = lca2(tax, c1, c2);
                                    if (l == c1 || l == c2) have = false;
                                }
                            }
                            if (have && be.ok) {
                                r_call = WFL_CALL_LGT;
                                r_b1 = be.c1;
                                r_b2 = be.c2;
                                r_c1 = c1;
                                r_c2 = c2;
                                r_lca = lca2(tax, c1, c2);   // waafle_orgscorer.py:882
                                r_crit = bcrit;
                                r_rank = brank;
                                r_dir = be.dir;
                                if (melded) {
                                    int na = 0, nb = 0;
#pragma unroll 1
                                    for (int t = lane; t < T; t += 32) { na += memA[t]; nb += memB[t]; }
                                    r_na = warp_sum(na);
                                    r_nb = warp_sum(nb);
                                }
#pragma unroll 1
                                for (int i = lane; i < G; i += 32) {
                                    u64 m = 1ull << (i & 63);
                                    int w = i >> 6;
                                    a.o.synteny[l0 + l_raw[i]] = ign[i] ? '~' : (bestm[2 * W + w] & m) ? '*'
                                                                 : (bestm[w] & m) ? 'A' : (bestm[W + w] & m) ? 'B' : '!';
                                }
                            }
                        }
                    }
                    PH(6);
                    if (r_call != WFL_CALL_UNCLASSIFIED) {
                        // ---- melded members -> staging pool (tails, waafle_orgscorer.py:630,658-659)
                        if (r_na + r_nb > 0) {
                            long long mb = 0;
                            if (lane == 0)
                                mb = (long long)atomicAdd(&a.ctr->mem_pool_used, (unsigned long long)(r_na + r_nb));
                            r_mem = __shfl_sync(FULL, mb, 0);
#pragma unroll 1
                            for (int side = 0; side < 2; ++side) {
                                const u8 *mem = side ? memB : memA;
                                long long off = r_mem + (side ? r_na : 0);
                                if ((side ? r_nb : r_na) == 0) continue;
                                int mbase = 0;
#pragma unroll 1
                                for (int base = 0; base < T; base += 32) {
                                    int t = base + lane;
                                    bool f = t < T && mem[t];
                                    u32 m = __ballot_sync(FULL, f);
                                    long long dst = off + mbase + __popc(m & lt_mask());
                                    if (f && dst < a.o.mem_pool_cap) a.o.mem_pool[dst] = cl_id[t];
                                    mbase += __popc(m);
                                }
                            }
                        }
                        break;
                    }
                    // not explained at this level: stop or lift (waafle_orgscorer.py:571-575)
                    if (T == 0 || hasroot) break;
                    if (iter >= 100) { r_status = 2; break; }   // :580-581
                    // ---- K5: lift the distinct-clade table (raise_taxonomy, :431-445) ----------------
                    // parents of the T clades, deduped and re-ranked (rank by counting); records follow
                    // through map_t.  "Unknown" spiked at this level is not a site-score clade: it is
                    // dropped here and re-inserted (:418, :433-443).
                    {
                        int *par = ar.get<int>(T + 1);
                        if (!ar.ok) { overflow = true; break; }
                        bool rec_unknown = false;   // does a record clade equal Unknown? (hit taxon named so)
                        if (spike) rec_unknown = nu > 0;
#pragma unroll 1
                        for (int t = lane; t < T; t += 32) {
                            bool drop = spike && t == t_unk && !rec_unknown;
                            par[t] = drop ? -1 : tax.parent[cl_id[t]];
                        }
                        if (lane == 0) par[T] = spike ? tax.unknown : -1;   // re-inserted spike key
                        __syncwarp();
                        int Tn = 0;
                        const int nwl = (tax.n_nodes + 31) >> 5;
                        if (nwl <= BITMAP_MAX_WORDS) {
                            // small taxonomies: presence bitmap of the parents, new rank = prefix popcount
                            u32 *bm = ar.get<u32>(nwl);
                            int *bpre = ar.get<int>(nwl + 1);
                            if (!ar.ok) { overflow = true; break; }
#pragma unroll 1
                            for (int w = lane; w < nwl; w += 32) bm[w] = 0;
                            __syncwarp();
#pragma unroll 1
                            for (int t = lane; t <= T; t += 32)
                                if (par[t] >= 0) atomicOr(&bm[par[t] >> 5], 1u << (par[t] & 31));
                            __syncwarp();
#pragma unroll 1
                            for (int base = 0; base < nwl; base += 32) {
                                const int w = base + lane;
                                int tot, ex = warp_excl_scan(w < nwl ? __popc(bm[w]) : 0, tot);
                                if (w < nwl) bpre[w] = Tn + ex;
                                Tn += tot;
                            }
                            __syncwarp();
#pragma unroll 1
                            for (int t = lane; t <= T; t += 32) {
                                const int key = par[t];
                                map_t[t] = key < 0 ? -1 : bpre[key >> 5] + __popc(bm[key >> 5] & ((1u << (key & 31)) - 1u));
                            }
                        } else {
#pragma unroll 1
                        for (int base = 0; base <= T; base += 32) {   // first occurrences of each key
                            int t = base + lane;
                            bool f = false;
                            if (t <= T && par[t] >= 0) {
                                f = true;
#pragma unroll 1
                                for (int z = 0; z < t; ++z)
                                    if (par[z] == par[t]) { f = false; break; }
                            }
                            if (t <= T) fo[t] = f;
                            Tn += __popc(__ballot_sync(FULL, f));
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int t = lane; t <= T; t += 32) {   // rank = #distinct keys below
                            int key = par[t], rk = -1;
                            if (key >= 0) {
                                rk = 0;
#pragma unroll 1
                                for (int q = 0; q <= T; ++q) rk += fo[q] && par[q] < key;
                            }
                            map_t[t] = rk;
                        }
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int r = lane; r < M; r += 32) r_t[r] = map_t[r_t[r]];
                        __syncwarp();
#pragma unroll 1
                        for (int t = lane; t <= T; t += 32)
                            if (map_t[t] >= 0) cl_id[map_t[t]] = par[t];   // equal keys write equal values
                        T = Tn;
                        __syncwarp();
                    }
                    ++lifts;
                    PH(7);
                }
                PH(8);
                if (lane == 0) {
#pragma unroll 1
                    for (int q = 0; q < 9; ++q) atomicAdd(&a.ctr->phase_cycles[q], ph[q]);
                    atomicAdd(&a.ctr->matched_pairs, (unsigned long long)M);
                    atomicAdd(&a.ctr->groups, (unsigned long long)n_groups);
                    atomicAdd(&a.ctr->levels, (unsigned long long)n_levels);
                    if (n_ptest) atomicAdd(&a.ctr->pairs_tested, (unsigned long long)n_ptest);
                    if (n_pscore) atomicAdd(&a.ctr->pairs_scored, (unsigned long long)n_pscore);
                    if (ar.all_smem && !overflow) atomicAdd(&a.ctr->smem_contigs, 1ull);
                }
            }
        }

        if (overflow) {
            // replay with a larger slab: report a worst-case byte count for this contig
            r_status = 1;
            r_call = WFL_CALL_UNCLASSIFIED;
            if (lane == 0) {
                atomicMax(&a.ctr->slab_need_max, need_hint);
                atomicAdd(&a.ctr->n_overflow, 1ull);
            }
        }
        if (__any_sync(FULL, bad_input)) r_status = 3;
        if (lane == 0) {
            if (r_status == 2) atomicAdd(&a.ctr->n_runaway, 1ull);
            if (r_status == 3) atomicAdd(&a.ctr->n_badinput, 1ull);
            a.o.call[c] = (uint8_t)r_call;
            a.o.direction[c] = (uint8_t)r_dir;
            a.o.lifts[c] = lifts;
            a.o.clade1[c] = r_c1;
            a.o.clade2[c] = r_c2;
            a.o.lca[c] = r_lca;
            a.o.best1[c] = r_b1;
            a.o.best2[c] = r_b2;
            a.o.crit[c] = r_crit;
            a.o.rank[c] = r_rank;
            a.o.n_mem_a[c] = r_na;
            a.o.n_mem_b[c] = r_nb;
            a.o.mem_pos[c] = r_mem;
            a.o.status[c] = (uint8_t)r_status;
        }
        __syncwarp();
    }
}

void launch_score_kernel_warp(const ScoreArgs &a, int grid, cudaStream_t s) {
    cudaFuncSetAttribute(wfl_score_contigs_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    wfl_score_contigs_warp<<<grid, 32, a.smem_bytes, s>>>(a);
}

}  // namespace wfl
