// The per-contig scoring and clade-assignment kernel (sm_100a).
//
// One CTA owns one contig at a time (persistent CTAs pull contigs from a global work queue) and
// runs the whole reference block waafle/waafle_orgscorer.py:952-960 for it:
//
//   K1  hit x locus matching            attach_hits / hit_locus_overlap / calc_overlap
//                                       (waafle_orgscorer.py:359-369,559-564, utils.py:487-500)
//   K2  envelope integral per (clade, locus) group, numpy-pairwise-exact
//                                       score_hit + np.mean (waafle_orgscorer.py:371-382,399-406)
//   K3  annotation arg-max              waafle_orgscorer.py:384-392
//   K4  weak-loci mask / Unknown spike  waafle_orgscorer.py:407-429
//   K5  taxonomy lift (parent gather + regroup)   waafle_orgscorer.py:431-445
//   K6  one-clade search                waafle_orgscorer.py:447-461,495-509,585-597,621-631
//   K7  two-clade pair search on gene bitmasks    waafle_orgscorer.py:511-545,599-619
//   K8  ranking / meld / LGT filters    waafle_orgscorer.py:633-744, utils.py:401-411
//   K9  level loop                      waafle_orgscorer.py:566-583
//
// Data layout: every per-contig array is carved from a bump arena that hands out dynamic shared
// memory first and a per-CTA global slab (L2-resident) once shared memory is exhausted, so small
// contigs (the 2-20 gene regime) live entirely in shared memory and 100-kb contigs still run.
//
// Arithmetic: fp64 throughout.  Gene scores are the mean over the gene's sites of the upper
// envelope of the matched hits' scores; the sum is accumulated in *numpy's pairwise order*
// (8 strided lanes per <=128-element leaf, leaves combined by the n/2 - (n/2)%8 split) directly
// over the piecewise-constant envelope, so every threshold decision is bit-identical to the
// reference without materialising per-site arrays.
#include "wfl_device.cuh"

namespace wfl {

namespace {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int KEY_LOCUS_BITS = 32;                 // group key = clade << 32 | locus
constexpr u64 KEY_LOCUS_MASK = (1ull << KEY_LOCUS_BITS) - 1;
constexpr int MAX_WARPS = 32;

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------

// Order-preserving map double -> u64 (so atomicMax on the bits is a max on the doubles).
__device__ __forceinline__ u64 dbits(double x) {
    if (x == 0.0) x = 0.0;   // -0.0 -> +0.0
    u64 u = (u64)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dbits_inv(u64 b) {
    u64 u = (b >> 63) ? (b & 0x7fffffffffffffffull) : ~b;
    return __longlong_as_double((long long)u);
}

struct Arena {
    char *smem, *slab;
    size_t smem_cap, slab_cap, smem_used, slab_used;
    bool ok, all_smem;
    template <class T>
    __device__ __forceinline__ T *get(size_t n) {
        size_t bytes = (n * sizeof(T) + 15) & ~size_t(15);
        if (smem_used + bytes <= smem_cap) {
            T *p = reinterpret_cast<T *>(smem + smem_used);
            smem_used += bytes;
            return p;
        }
        all_smem = false;
        if (slab_used + bytes <= slab_cap) {
            T *p = reinterpret_cast<T *>(slab + slab_used);
            slab_used += bytes;
            return p;
        }
        ok = false;
        return nullptr;
    }
};

struct Shared {
    long long c;
    int red_i[MAX_WARPS + 1];
    int red_j[MAX_WARPS + 1];
    int red_k[MAX_WARPS + 1];
    int fill;
    u64 best_bits;
    long long best_idx;
    int lca_a[MAX_WARPS], lca_b[MAX_WARPS];
    int flag_a, flag_b, flag_c;
    long long mem_base;
};

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum of three ints over the block, result to every thread.
__device__ __forceinline__ void block_sum3(int &a, int &b, int &c, Shared &sh) {
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) {
        sh.red_i[w] = a;
        sh.red_j[w] = b;
        sh.red_k[w] = c;
    }
    __syncthreads();
    a = b = c = 0;
    for (int i = 0; i < nw; ++i) {
        a += sh.red_i[i];
        b += sh.red_j[i];
        c += sh.red_k[i];
    }
    __syncthreads();
}
__device__ __forceinline__ int block_sum(int a, Shared &sh) {
    int b = 0, c = 0;
    block_sum3(a, b, c, sh);
    return a;
}

// Exclusive scan of one int per thread; total to every thread.
__device__ __forceinline__ int block_excl_scan(int v, int &total, Shared &sh) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh.red_i[w] = inc;
    __syncthreads();
    int base = 0;
    total = 0;
    for (int i = 0; i < nw; ++i) {
        int x = sh.red_i[i];
        if (i < w) base += x;
        total += x;
    }
    __syncthreads();
    return base + inc - v;
}

// Block-cooperative bitonic sort of (key, val) pairs; n is a power of two.
__device__ void bitonic_sort(u64 *key, u32 *val, int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    u64 a = key[i], b = key[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up && a != b) {
                        key[i] = b;
                        key[ixj] = a;
                        u32 t = val[i];
                        val[i] = val[ixj];
                        val[ixj] = t;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// LCA by depth-aligned parent walk == deepest common prefix of root-first lineages
// (waafle/utils.py:401-411).  -1 is the identity.
__device__ int lca2(const DevTax &t, int a, int b) {
    if (a < 0) return b;
    if (b < 0) return a;
    int da = t.depth[a], db = t.depth[b];
    while (da > db) { a = t.parent[a]; --da; }
    while (db > da) { b = t.parent[b]; --db; }
    while (a != b) { a = t.parent[a]; b = t.parent[b]; }
    return a;
}

__device__ int block_lca(const DevTax &t, int v, int *scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = lca2(t, v, __shfl_xor_sync(0xffffffffu, v, o));
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    int r = -1;
    for (int i = 0; i < nw; ++i) r = lca2(t, r, scratch[i]);
    __syncthreads();
    return r;
}

// ---------------------------------------------------------------------------------------------
// numpy pairwise summation over a sequential value source
// (numpy/_core/src/umath/loops_utils.h.src DOUBLE_pairwise_sum: n<8 plain loop; n<=128 eight
//  strided accumulators, ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the n%8 tail; else split at
//  n/2 - (n/2)%8.)  `Src` yields a[0], a[1], ... and may report constant runs.
// ---------------------------------------------------------------------------------------------

template <class Src>
__device__ double pw_leaf(Src &s, int m) {
    if (m < 8) {
        double r = 0.0;
        for (int i = 0; i < m; ++i) r += s.next();
        return r;
    }
    const int k8 = m >> 3, tail = m & 7;
    double res, v;
    if (s.take_const(m - tail, v)) {
        // all eight lanes see the same k8 values: r = v+v+...+v (sequential), res = 8r (exact)
        double r = v;
        if (v != 0.0)
            for (int i = 1; i < k8; ++i) r += v;
        res = 8.0 * r;
    } else {
        double r0 = s.next(), r1 = s.next(), r2 = s.next(), r3 = s.next();
        double r4 = s.next(), r5 = s.next(), r6 = s.next(), r7 = s.next();
        for (int c = 1; c < k8; ++c) {
            if (s.take_const(8, v)) {
                r0 += v; r1 += v; r2 += v; r3 += v; r4 += v; r5 += v; r6 += v; r7 += v;
            } else {
                r0 += s.next(); r1 += s.next(); r2 += s.next(); r3 += s.next();
                r4 += s.next(); r5 += s.next(); r6 += s.next(); r7 += s.next();
            }
        }
        res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    }
    for (int i = 0; i < tail; ++i) res += s.next();
    return res;
}

template <class Src>
__device__ double pairwise_sum(Src &s, long long n) {
    if (n <= 128) return pw_leaf(s, (int)n);
    long long sz[36];
    double acc[36];
    unsigned char st[36];
    int sp = 0;
    sz[0] = n;
    st[0] = 0;
    double ret = 0.0;
    bool returning = false;
    for (;;) {
        if (!returning) {
            long long m = sz[sp];
            if (m <= 128) {
                ret = pw_leaf(s, (int)m);
                returning = true;
                if (--sp < 0) break;
            } else {
                long long n2 = m / 2;
                n2 -= n2 % 8;
                st[sp] = 1;
                sz[sp + 1] = n2;
                st[++sp] = 0;
            }
        } else if (st[sp] == 1) {
            acc[sp] = ret;
            long long m = sz[sp], n2 = m / 2;
            n2 -= n2 % 8;
            st[sp] = 2;
            sz[sp + 1] = m - n2;
            st[++sp] = 0;
            returning = false;
        } else {
            ret = acc[sp] + ret;
            if (--sp < 0) break;
        }
    }
    return ret;
}

// Per-site upper envelope of one (clade, locus) group, streamed as constant runs.
struct SiteSrc {
    const u32 *sidx;
    const int *ra, *rb;
    const double *rv;
    int rs, re, n, pos, run_end;
    double run_v;
    // Records of a group arrive in DESCENDING score order, so the first record that covers
    // `pos` is the envelope value there; the run ends where that record ends or where a
    // higher-scoring record (all seen before it) starts.
    __device__ __forceinline__ void advance() {
        double v = 0.0;   // np.zeros(len(locus)), waafle_orgscorer.py:381
        int nx = n;
        for (int r = rs; r < re; ++r) {
            u32 i = sidx[r];
            int a = ra[i], b = rb[i];
            if (a <= pos && pos < b) {
                v = fmax(0.0, rv[i]);   // np.maximum(slice, score), waafle_orgscorer.py:382
                nx = min(nx, b);
                break;
            }
            if (a > pos) nx = min(nx, a);
        }
        run_v = v;
        run_end = nx;
    }
    __device__ __forceinline__ double next() {
        if (pos >= run_end) advance();
        ++pos;
        return run_v;
    }
    __device__ __forceinline__ bool take_const(int m, double &v) {
        if (pos >= run_end) advance();
        if (run_end - pos >= m) {
            v = run_v;
            pos += m;
            return true;
        }
        return false;
    }
};

// Gene-score rows are sparse: groups sorted by (clade, locus); a missing locus scores 0
// (waafle_orgscorer.py:404-405).
struct RowCursor {
    const int *g_loc;
    const double *g_score;
    int p, e;
    __device__ __forceinline__ double at(int locus) {
        while (p < e && g_loc[p] < locus) ++p;
        return (p < e && g_loc[p] == locus) ? g_score[p] : 0.0;
    }
};

// Values of Contig.score (waafle_orgscorer.py:447-461): max(s1, s2) at the non-ignored loci.
struct ScoreSrc {
    RowCursor r1, r2;
    const unsigned char *ign;
    int i;
    bool two;
    double crit;
    __device__ __forceinline__ double next() {
        while (ign[i]) ++i;
        double v = r1.at(i);
        if (two) v = fmax(v, r2.at(i));
        ++i;
        crit = fmin(crit, v);
        return v;
    }
    __device__ __forceinline__ bool take_const(int, double &) { return false; }
};

struct Level {   // per-level views shared by the search routines (all pointers arena-backed)
    int G, W, T, Ngrp, n_unmasked;
    const int *g_loc, *g_clade;
    const double *g_score;
    const int *cl_id, *cl_go;
    const u64 *mk[3];   // per clade gene bitmasks: score >= k1 / k2 / c_eps
    const u64 *um;      // non-ignored loci
    const unsigned char *ign;
    const int *l_len;
};

__device__ __forceinline__ void score_clades(const Level &L, int t1, int t2, double &crit,
                                             double &rank) {
    ScoreSrc s;
    s.r1 = RowCursor{L.g_loc, L.g_score, L.cl_go[t1], L.cl_go[t1 + 1]};
    s.two = t2 >= 0;
    s.r2 = s.two ? RowCursor{L.g_loc, L.g_score, L.cl_go[t2], L.cl_go[t2 + 1]} : s.r1;
    s.ign = L.ign;
    s.i = 0;
    s.crit = __longlong_as_double(0x7ff0000000000000ll);
    double sum = pairwise_sum(s, L.n_unmasked);
    crit = s.crit;
    rank = sum / (double)L.n_unmasked;   // np.mean = add.reduce / n
}

// Letters of a two-clade option on word w (waafle_orgscorer.py:524-534), before the A/B swap.
__device__ __forceinline__ void letters(const Level &L, const u64 *mamb, bool unknown_involved,
                                        int t1, int t2, int w, u64 &A, u64 &B, u64 &amb) {
    u64 um = L.um[w];
    amb = unknown_involved ? 0ull : (mamb[(size_t)t1 * L.W + w] & mamb[(size_t)t2 * L.W + w] & um);
    A = L.mk[1][(size_t)t1 * L.W + w] & um & ~amb;
    B = L.mk[1][(size_t)t2 * L.W + w] & um & ~amb & ~A;
}

struct TwoEval {
    bool swap, dir, ok;
    int c1, c2, t1, t2;   // post-swap clade node ids / table positions
};

// set_synteny_two + apply_lgt_checks for one option (waafle_orgscorer.py:511-545, 678-744).
__device__ void eval_two(const Level &L, const DevTax &tax, const DevParams &P, int ta, int tb,
                         TwoEval &ev) {
    const u64 *mamb = L.mk[P.amb_sel], *msis = L.mk[P.sis_sel];
    const bool unk = L.cl_id[ta] == tax.unknown || L.cl_id[tb] == tax.unknown;
    bool swap = false;
    for (int w = 0; w < L.W; ++w) {   // "^[^A]*B": first clear letter is B -> swap
        u64 A, B, amb;
        letters(L, mamb, unk, ta, tb, w, A, B, amb);
        u64 ab = A | B;
        if (ab) {
            u64 low = ab & (~ab + 1);
            swap = (B & low) != 0;
            break;
        }
    }
    ev.swap = swap;
    ev.t1 = swap ? tb : ta;
    ev.t2 = swap ? ta : tb;
    ev.c1 = L.cl_id[ev.t1];
    ev.c2 = L.cl_id[ev.t2];
    // one pass over the loci: lengths, counts, "^A+B+A+$" on the non-ignored letters
    long long total_len = 0, amb_len = 0;
    int nA = 0, nB = 0, state = 0;
    for (int w = 0; w < L.W; ++w) {
        u64 A, B, amb;
        letters(L, mamb, unk, ta, tb, w, A, B, amb);
        if (swap) { u64 t = A; A = B; B = t; }
        u64 um = L.um[w];
        nA += __popcll(A);
        nB += __popcll(B);
        u64 bits = um;
        while (bits) {
            int b = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            u64 m = 1ull << b;
            int len = L.l_len[w * 64 + b];
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
    }
    ev.dir = state == 3;
    bool ok = true;
    if (total_len > 0 && (double)amb_len / (double)total_len > P.p.ambiguous_fraction) ok = false;
    if (P.p.clade_genes >= 0 && min(nA, nB) < P.p.clade_genes) ok = false;
    if (P.p.clade_leaves >= 0) {
        int lc = ev.dir ? tax.leaf_count[ev.c2]
                        : min(tax.leaf_count[ev.c1], tax.leaf_count[ev.c2]);
        if (lc < P.p.clade_leaves) ok = false;
    }
    if (P.p.sister_penalty != 0 && ok) {
        const int p1 = tax.parent[ev.c1], p2 = tax.parent[ev.c2];
        for (int t = 0; t < L.T && ok; ++t) {
            int x = L.cl_id[t];
            if (x == ev.c1 || x == ev.c2 || !tax.listed[x]) continue;
            int px = tax.parent[x];
            bool s1 = px == p1, s2 = (px == p2) && !ev.dir;
            if (!s1 && !s2) continue;
            for (int w = 0; w < L.W; ++w) {
                u64 A, B, amb;
                letters(L, mamb, unk, ta, tb, w, A, B, amb);
                if (swap) { u64 tt = A; A = B; B = tt; }
                u64 ms = msis[(size_t)t * L.W + w];
                // a B locus is penalised by clade1's sisters, an A locus by clade2's
                if ((s1 && (ms & B)) || (s2 && (ms & A))) { ok = false; break; }
            }
        }
    }
    ev.ok = ok;
}

__device__ __forceinline__ void pair_decode(long long p, int n, int &i, int &j) {
    // pairs enumerated i-major: (0,1),(0,2),...,(0,n-1),(1,2),...; offset(i) = i*(2n-i-1)/2
    double nn = 2.0 * n - 1.0;
    long long ii = (long long)floor((nn - sqrt(nn * nn - 8.0 * (double)p)) * 0.5);
    if (ii < 0) ii = 0;
    if (ii > n - 2) ii = n - 2;
    while (ii > 0 && ii * (2LL * n - ii - 1) / 2 > p) --ii;
    while ((ii + 1) * (2LL * n - ii - 2) / 2 <= p) ++ii;
    i = (int)ii;
    j = (int)(p - ii * (2LL * n - ii - 1) / 2) + i + 1;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------

__global__ void wfl_score_contigs(const ScoreArgs a) {
    extern __shared__ __align__(16) char smem_dyn[];
    __shared__ Shared sh;
    const int tid = threadIdx.x, B = blockDim.x;
    const DevParams &P = a.P;
    const DevTax &tax = a.t;
    const int S = P.p.n_systems;
    const bool spike = P.p.weak_loci == 2;
    const double thr3[3] = {P.p.k1, P.p.k2, 1e-6};

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            unsigned long long w = atomicAdd(&a.ctr->next_work, 1ull);
            sh.c = (long long)w < a.n_work ? (a.work_list ? a.work_list[w] : a.work_base + (long long)w) : -1;
        }
        __syncthreads();
        const long long c = sh.c;
        if (c < 0) break;

        const long long h0 = a.b.hit_off[c], l0 = a.b.locus_off[c];
        const int H = (int)(a.b.hit_off[c + 1] - h0), Graw = (int)(a.b.locus_off[c + 1] - l0);
        Arena ar{smem_dyn, a.slab + (size_t)blockIdx.x * a.slab_bytes, (size_t)a.smem_bytes,
                 a.slab_bytes, 0, 0, true, true};

        long long ph_last = clock64();
        unsigned long long ph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#define PH(i) do { if (tid == 0) { long long t_ = clock64(); ph[i] += (unsigned long long)(t_ - ph_last); ph_last = t_; } } while (0)
        // ---- loci: --min-gene-length filter, GFF order kept (waafle_orgscorer.py:348-352) ----
        int *l_lo = ar.get<int>(Graw), *l_len = ar.get<int>(Graw), *l_raw = ar.get<int>(Graw);
        signed char *l_str = ar.get<signed char>(Graw);
        bool overflow = !ar.ok;
        int G = 0;
        for (int base = 0; base < Graw && !overflow; base += B) {
            int j = base + tid, flag = 0, lo = 0, len = 0;
            if (j < Graw) {
                int s = a.b.locus_start[l0 + j], e = a.b.locus_end[l0 + j];
                lo = min(s, e);
                len = max(s, e) - lo + 1;
                flag = (double)len >= P.p.min_gene_length;
                a.o.locus_flags[l0 + j] = flag ? WFL_LOCUS_RETAINED : 0;
                a.o.synteny[l0 + j] = 0;
                for (int s2 = 0; s2 < S; ++s2) a.o.ann_winner[(l0 + j) * S + s2] = -1;
            }
            int tot, pos = block_excl_scan(flag, tot, sh);
            if (flag) {
                l_lo[G + pos] = lo;
                l_len[G + pos] = len;
                l_raw[G + pos] = j;
                l_str[G + pos] = a.b.locus_strand[l0 + j];
            }
            G += tot;
        }
        const int W = (G + 63) >> 6;
        unsigned long long need_hint = 64ull * Graw + (1ull << 12);
        int lifts = H > 0 ? P.p.jump_taxonomy : 0;
        // result registers (thread 0 writes them out at the end)
        int r_call = WFL_CALL_UNCLASSIFIED, r_dir = 0, r_c1 = -1, r_c2 = -1, r_lca = -1, r_b1 = -1,
            r_b2 = -1, r_na = 0, r_nb = 0, r_status = 0;
        long long r_mem = 0;
        double r_crit = 0.0, r_rank = 0.0;

        if (!overflow && H > 0 && G > 0) {
            __syncthreads();
            // ---- K1 pass 1: count (hit, locus) matches -------------------------------------
            int cnt = 0;
            for (int h = tid; h < H; h += B) {
                if (!(a.b.hit_scov[h0 + h] >= P.p.min_scov)) continue;   // waafle_orgscorer.py:362
                int q1 = a.b.hit_qstart[h0 + h], q2 = a.b.hit_qend[h0 + h];
                int hmin = min(q1, q2), hmax = max(q1, q2), hlen = hmax - hmin + 1;
                signed char hs = a.b.hit_strand[h0 + h];
                for (int i = 0; i < G; ++i) {
                    if (P.p.stranded && hs != l_str[i]) continue;          // :365
                    int lmin = l_lo[i], lmax = lmin + l_len[i] - 1;
                    double ov = 0.0;                                        // utils.py:492-499
                    if (!(lmin > hmax || hmin > lmax))
                        ov = (double)(min(hmax, lmax) - max(hmin, lmin) + 1) /
                             (double)min(hlen, l_len[i]);
                    cnt += ov >= P.p.min_overlap;                           // :367
                }
            }
            const int M = block_sum(cnt, sh);
            PH(0);
            int Mp2 = 1;
            while (Mp2 < M) Mp2 <<= 1;
            // worst case for this contig (groups <= M + G, clades <= groups): one replay suffices
            need_hint = 64ull * Graw + 128ull * (unsigned long long)Mp2 + 32ull * ((unsigned long long)M + G) +
                        (48ull + 24ull * W) * ((unsigned long long)M + G + 1) +
                        (unsigned long long)G * (64 + 16 * S) + (1ull << 12);

            // ---- record arrays ---------------------------------------------------------------
            double *r_v = ar.get<double>(M);
            u64 *key = ar.get<u64>(Mp2);
            u32 *sidx = ar.get<u32>(Mp2);
            int *r_a = ar.get<int>(M), *r_b = ar.get<int>(M), *r_cl = ar.get<int>(M),
                *r_loc = ar.get<int>(M), *r_ord = ar.get<int>(M), *r_hit = S > 0 ? ar.get<int>(M) : nullptr;
            double *maxv = ar.get<double>(G);
            u64 *maxb = ar.get<u64>(G);
            unsigned char *ign = ar.get<unsigned char>(G + 1);
            u64 *um = ar.get<u64>(W);
            u64 *annb = S > 0 ? ar.get<u64>((size_t)G * S) : nullptr;
            int *annw = S > 0 ? ar.get<int>((size_t)G * S) : nullptr;
            overflow = !ar.ok;
            if (!overflow) {
                if (tid == 0) sh.fill = 0;
                for (int i = tid; i < G * S; i += B) { annb[i] = 0; annw[i] = -1; }
                __syncthreads();
                // ---- K1 pass 2: emit records (score_hit, waafle_orgscorer.py:371-382) -------
                for (int h = tid; h < H; h += B) {
                    if (!(a.b.hit_scov[h0 + h] >= P.p.min_scov)) continue;
                    int q1 = a.b.hit_qstart[h0 + h], q2 = a.b.hit_qend[h0 + h];
                    int hmin = min(q1, q2), hmax = max(q1, q2), hlen = hmax - hmin + 1;
                    signed char hs = a.b.hit_strand[h0 + h];
                    int cl = -1;
                    double sc = 0.0;
                    for (int i = 0; i < G; ++i) {
                        if (P.p.stranded && hs != l_str[i]) continue;
                        int lmin = l_lo[i], len = l_len[i], lmax = lmin + len - 1;
                        double ov = 0.0;
                        if (!(lmin > hmax || hmin > lmax))
                            ov = (double)(min(hmax, lmax) - max(hmin, lmin) + 1) /
                                 (double)min(hlen, len);
                        if (!(ov >= P.p.min_overlap)) continue;
                        if (cl < 0) {
                            cl = a.b.hit_taxon[h0 + h];
                            if ((u32)cl >= (u32)tax.n_nodes) { cl = tax.root; atomicAdd(&a.ctr->n_badinput, 1ull); }
                            for (int j = 0; j < P.p.jump_taxonomy; ++j) cl = tax.parent[cl];
                            sc = a.b.hit_score[h0 + h];
                        }
                        // python slice [h1 : h2+1] of a length-len array (:376-382)
                        int s1 = max(0, hmin - lmin), e1 = min(len - 1, hmax - lmin) + 1;
                        if (e1 < 0) e1 = max(0, e1 + len);
                        s1 = min(s1, len);
                        if (e1 < s1) e1 = s1;
                        int slot = atomicAdd(&sh.fill, 1);
                        r_v[slot] = sc;
                        r_a[slot] = s1;
                        r_b[slot] = e1;
                        r_cl[slot] = cl;
                        r_loc[slot] = i;
                        if (S > 0) {
                            r_hit[slot] = h;
                            // K3 phase 1: running max of the annotated hits' scores per (locus, system)
                            u32 m = a.b.hit_sysmask[h0 + h];
                            if (m && sc >= P.ann_thr) {
                                u64 sb = dbits(sc);
                                while (m) {
                                    int s2 = __ffs(m) - 1;
                                    m &= m - 1;
                                    atomicMax(&annb[(size_t)i * S + s2], sb);
                                }
                            }
                        }
                    }
                }
                __syncthreads();
                if (S > 0) {
                    // K3 phase 2: the LAST hit (file order) attaining the max wins (:389, '>=')
                    for (int r = tid; r < M; r += B) {
                        u32 m = a.b.hit_sysmask[h0 + r_hit[r]];
                        double sc = r_v[r];
                        if (m && sc >= P.ann_thr) {
                            u64 sb = dbits(sc);
                            while (m) {
                                int s2 = __ffs(m) - 1;
                                m &= m - 1;
                                if (annb[(size_t)r_loc[r] * S + s2] == sb)
                                    atomicMax(&annw[(size_t)r_loc[r] * S + s2], r_hit[r]);
                            }
                        }
                    }
                    __syncthreads();
                    for (int i = tid; i < G * S; i += B) {
                        int w = annw[i];
                        a.o.ann_winner[(l0 + l_raw[i / S]) * S + (i % S)] = w >= 0 ? (int)(h0 + w) : -1;
                    }
                }

                PH(1);
                // ---- level-invariant base order: (locus, score descending) -----------------
                // rank2[r] = position of record r in that order; every level then sorts by
                // (clade, rank2), which groups by (clade, locus) with descending scores inside.
                for (int r = tid; r < Mp2; r += B) {
                    key[r] = r < M ? ~dbits(r_v[r]) : ~0ull;
                    sidx[r] = (u32)r;
                }
                __syncthreads();
                bitonic_sort(key, sidx, Mp2);
                for (int r = tid; r < M; r += B) r_ord[sidx[r]] = r;
                __syncthreads();
                for (int r = tid; r < Mp2; r += B) {
                    key[r] = r < M ? (((u64)(u32)r_loc[r] << 32) | (u64)(u32)r_ord[r]) : ~0ull;
                    sidx[r] = (u32)r;
                }
                __syncthreads();
                bitonic_sort(key, sidx, Mp2);
                for (int r = tid; r < M; r += B) r_ord[sidx[r]] = r;
                __syncthreads();
                // ---- K9: level loop (evaluate_contig, waafle_orgscorer.py:566-583) ----------
                const size_t mark_smem = ar.smem_used, mark_slab = ar.slab_used;
                int n_levels = 0;
                long long n_groups = 0, n_ptest = 0, n_pscore = 0;
                for (int iter = 0;; ++iter) {
                    ar.smem_used = mark_smem;
                    ar.slab_used = mark_slab;
                    ++n_levels;
                    // ---- regroup: sort records by (clade, locus) ---------------------------
                    for (int r = tid; r < Mp2; r += B) {
                        key[r] = r < M ? (((u64)(u32)r_cl[r] << 32) | (u64)(u32)r_ord[r]) : ~0ull;
                        sidx[r] = (u32)r;
                    }
                    __syncthreads();
                    bitonic_sort(key, sidx, Mp2);
                    // the rank has done its job: turn the sort key into the group key (clade, locus)
                    for (int r = tid; r < M; r += B)
                        key[r] = (key[r] & ~KEY_LOCUS_MASK) | (u64)(u32)r_loc[sidx[r]];
                    __syncthreads();
                    PH(2);
                    const u64 unk_lo = (u64)(u32)tax.unknown << KEY_LOCUS_BITS;
                    int ng = 0, nlt = 0, nu = 0;
                    for (int r = tid; r < M; r += B) {
                        u64 k = key[r];
                        if (r == 0 || k != key[r - 1]) {
                            ++ng;
                            if (spike) {
                                nlt += k < unk_lo;
                                nu += (k >> KEY_LOCUS_BITS) == (u64)(u32)tax.unknown;
                            }
                        }
                    }
                    block_sum3(ng, nlt, nu, sh);
                    const int Ngrp = spike ? ng - nu + G : ng;
                    n_groups += Ngrp;
                    double *g_score = ar.get<double>(Ngrp);
                    int *g_rs = ar.get<int>(Ngrp), *g_loc = ar.get<int>(Ngrp),
                        *g_clade = ar.get<int>(Ngrp);
                    if (!ar.ok) { overflow = true; break; }
                    int gbase = 0;
                    for (int base = 0; base < M; base += B) {
                        int r = base + tid, flag = 0;
                        u64 k = 0;
                        if (r < M) {
                            k = key[r];
                            flag = r == 0 || k != key[r - 1];
                        }
                        int tot, gid = gbase + block_excl_scan(flag, tot, sh);
                        gbase += tot;
                        if (flag) {
                            int cl = (int)(k >> KEY_LOCUS_BITS);
                            int dst = gid;
                            if (spike)
                                dst = cl < tax.unknown ? gid : (cl == tax.unknown ? -1 : gid - nu + G);
                            if (dst >= 0) {
                                g_rs[dst] = r;
                                g_loc[dst] = (int)(k & KEY_LOCUS_MASK);
                                g_clade[dst] = cl;
                            }
                        }
                    }
                    if (spike)
                        for (int i = tid; i < G; i += B) {
                            g_rs[nlt + i] = -1;
                            g_loc[nlt + i] = i;
                            g_clade[nlt + i] = tax.unknown;
                        }
                    for (int i = tid; i < G; i += B) maxb[i] = dbits(0.0);
                    __syncthreads();
                    PH(3);
                    // ---- K2: envelope integral per group, numpy-pairwise-exact -------------
                    for (int g = tid; g < Ngrp; g += B) {
                        int rs = g_rs[g];
                        if (rs < 0) continue;
                        u64 k = key[rs];
                        int re = rs + 1;
                        while (re < M && key[re] == k) ++re;
                        int loc = g_loc[g], n = l_len[loc];
                        SiteSrc src{sidx, r_a, r_b, r_v, rs, re, n, 0, 0, 0.0};
                        double sc = pairwise_sum(src, n) / (double)n;
                        g_score[g] = sc;
                        if (g_clade[g] != tax.unknown)   // waafle_orgscorer.py:409-411
                            atomicMax(&maxb[loc], dbits(sc));
                    }
                    __syncthreads();
                    PH(4);
                    // ---- K4: weak loci (waafle_orgscorer.py:412-427) -------------------------
                    for (int i = tid; i < G; i += B) {
                        double mx = dbits_inv(maxb[i]);
                        maxv[i] = mx;
                        ign[i] = (P.p.weak_loci == 0) ? !(mx >= P.min_thr) : 0;
                        if (spike) g_score[nlt + i] = 1.0 - mx;
                    }
                    if (tid == 0) ign[G] = 0;   // sentinel for ScoreSrc
                    __syncthreads();
                    int nun = 0;
                    for (int w = tid; w < W; w += B) {
                        u64 m = 0;
                        for (int b = 0; b < 64 && w * 64 + b < G; ++b)
                            if (!ign[w * 64 + b]) m |= 1ull << b;
                        um[w] = m;
                        nun += __popcll(m);
                    }
                    for (int i = tid; i < G; i += B)
                        a.o.locus_flags[l0 + l_raw[i]] =
                            WFL_LOCUS_RETAINED | (ign[i] ? WFL_LOCUS_IGNORED : 0);
                    nun = block_sum(nun, sh);
                    if (iter == 0 && c == a.dbg_contig) {
                        for (int g = tid; g < Ngrp; g += B)
                            if (g < a.dbg_cap) {
                                a.dbg_clade[g] = g_clade[g];
                                a.dbg_locus[g] = g_loc[g];
                                a.dbg_score[g] = g_score[g];
                            }
                        if (tid == 0) *a.dbg_count = Ngrp;
                    }
                    if (iter == 0 && nun == 0) break;   // "empty" contig, waafle_orgscorer.py:959

                    // ---- clade table + gene bitmasks -------------------------------------------
                    int nt = 0, hasroot = 0, dummy = 0;
                    for (int g = tid; g < Ngrp; g += B)
                        if (g == 0 || g_clade[g] != g_clade[g - 1]) {
                            ++nt;
                            hasroot += g_clade[g] == tax.root;
                        }
                    block_sum3(nt, hasroot, dummy, sh);
                    const int T = nt;
                    int *cl_id = ar.get<int>(T), *cl_go = ar.get<int>(T + 1), *cand = ar.get<int>(T);
                    double *cl_rank = ar.get<double>(T), *cl_crit = ar.get<double>(T);
                    unsigned char *cl_opt = ar.get<unsigned char>(T), *memA = ar.get<unsigned char>(T),
                                  *memB = ar.get<unsigned char>(T);
                    u64 *mk0 = ar.get<u64>((size_t)T * W), *mk1 = ar.get<u64>((size_t)T * W),
                        *mk2 = ar.get<u64>((size_t)T * W);
                    u64 *bestm = ar.get<u64>(3 * (size_t)W);
                    if (!ar.ok) { overflow = true; break; }
                    int tbase = 0;
                    for (int base = 0; base < Ngrp; base += B) {
                        int g = base + tid, flag = 0;
                        if (g < Ngrp) flag = g == 0 || g_clade[g] != g_clade[g - 1];
                        int tot, t = tbase + block_excl_scan(flag, tot, sh);
                        tbase += tot;
                        if (flag) {
                            cl_id[t] = g_clade[g];
                            cl_go[t] = g;
                        }
                    }
                    if (tid == 0) cl_go[T] = Ngrp;
                    __syncthreads();
                    u64 *mks[3] = {mk0, mk1, mk2};
                    for (int t = tid; t < T; t += B) {
                        memA[t] = memB[t] = 0;
                        for (int q = 0; q < 3; ++q) {
                            u64 *m = mks[q] + (size_t)t * W;
                            const double thr = thr3[q];
                            // a locus without an entry scores 0 (waafle_orgscorer.py:404-405)
                            for (int w = 0; w < W; ++w) {
                                int nb = min(64, G - w * 64);
                                m[w] = thr <= 0.0 ? (nb == 64 ? ~0ull : ((1ull << nb) - 1)) : 0ull;
                            }
                            for (int g = cl_go[t]; g < cl_go[t + 1]; ++g) {
                                int loc = g_loc[g];
                                u64 bit = 1ull << (loc & 63);
                                if (g_score[g] >= thr) m[loc >> 6] |= bit;
                                else m[loc >> 6] &= ~bit;
                            }
                        }
                    }
                    __syncthreads();
                    Level L{G, W, T, Ngrp, nun, g_loc, g_clade, g_score, cl_id, cl_go,
                            {mk0, mk1, mk2}, um, ign, l_len};

                    PH(5);
                    // ---- K6: one-clade search (explain_one, waafle_orgscorer.py:585-597) ------
                    if (tid == 0) { sh.best_bits = 0; sh.best_idx = -1; }
                    __syncthreads();
                    for (int t = tid; t < T; t += B) {
                        bool pass = true;
                        for (int w = 0; w < W; ++w)
                            pass &= (mk0[(size_t)t * W + w] & um[w]) == um[w];   // crit >= k1
                        cl_opt[t] = pass;
                        if (pass) {
                            double crit, rank;
                            score_clades(L, t, -1, crit, rank);
                            cl_rank[t] = rank;
                            cl_crit[t] = crit;
                            atomicMax(&sh.best_bits, dbits(rank));
                        }
                    }
                    __syncthreads();
                    for (int t = tid; t < T; t += B)
                        if (cl_opt[t] && dbits(cl_rank[t]) == sh.best_bits)
                            atomicMax(&sh.best_idx, (long long)t);   // ties: last in name order
                    __syncthreads();
                    PH(6);
                    if (sh.best_idx >= 0) {
                        // meld_one (waafle_orgscorer.py:621-631)
                        const int tb = (int)sh.best_idx;
                        const double brank = cl_rank[tb];
                        r_call = WFL_CALL_NO_LGT;
                        r_b1 = r_c1 = cl_id[tb];
                        r_crit = cl_crit[tb];
                        r_rank = brank;
                        if (P.p.disambiguate_one == 1) {
                            int my = -1, nk = 0;
                            for (int t = tid; t < T; t += B)
                                if (cl_opt[t] && brank - cl_rank[t] <= P.p.range) {
                                    my = lca2(tax, my, cl_id[t]);
                                    memA[t] = 1;
                                    ++nk;
                                }
                            r_c1 = block_lca(tax, my, sh.lca_a);
                            r_na = block_sum(nk, sh);
                        }
                        for (int i = tid; i < G; i += B)   // set_synteny_one (:495-509)
                            a.o.synteny[l0 + l_raw[i]] =
                                ign[i] ? '~' : ((mk0[(size_t)tb * W + (i >> 6)] >> (i & 63)) & 1 ? 'A' : '!');
                    } else {
                        // ---- K7: two-clade search (explain_two, waafle_orgscorer.py:599-619) --
                        int T2 = 0;
                        for (int base = 0; base < T; base += B) {
                            int t = base + tid, flag = 0;
                            if (t < T)   // max(gene_scores[clade]) >= k2, unmasked (:603-605)
                                for (int w = 0; w < W; ++w) flag |= mk1[(size_t)t * W + w] != 0;
                            int tot, pos = block_excl_scan(flag, tot, sh);
                            if (flag) cand[T2 + pos] = t;
                            T2 += tot;
                        }
                        __syncthreads();
                        const long long NP = (long long)T2 * (T2 - 1) / 2;
                        n_ptest += NP;
                        // pass 1: best rank; ties -> last pair in (clade1, clade2) iteration order
                        double my_rank = -1.0;
                        long long my_p = -1;
                        int nsc = 0;
                        for (long long p = tid; p < NP; p += B) {
                            int i, j;
                            pair_decode(p, T2, i, j);
                            int t1 = cand[i], t2 = cand[j];
                            bool pass = true;
                            for (int w = 0; w < W; ++w)
                                pass &= ((mk1[(size_t)t1 * W + w] | mk1[(size_t)t2 * W + w]) & um[w]) == um[w];
                            if (!pass) continue;   // crit < k2 (:610)
                            ++nsc;
                            double crit, rank;
                            score_clades(L, t1, t2, crit, rank);
                            if (my_p < 0 || rank >= my_rank) { my_rank = rank; my_p = p; }
                        }
                        if (my_p >= 0) atomicMax(&sh.best_bits, dbits(my_rank));
                        n_pscore += block_sum(nsc, sh);
                        if (my_p >= 0 && dbits(my_rank) == sh.best_bits) atomicMax(&sh.best_idx, my_p);
                        __syncthreads();
                        const long long bp = sh.best_idx;
                        if (bp >= 0) {
                            // meld_two (waafle_orgscorer.py:633-669)
                            int bi, bj;
                            pair_decode(bp, T2, bi, bj);
                            TwoEval be;
                            eval_two(L, tax, P, cand[bi], cand[bj], be);
                            double bcrit, brank;
                            score_clades(L, cand[bi], cand[bj], bcrit, brank);
                            const bool bunk = be.c1 == tax.unknown || be.c2 == tax.unknown;
                            for (int w = tid; w < W; w += B) {
                                u64 A, Bm, amb;
                                letters(L, L.mk[P.amb_sel], bunk, cand[bi], cand[bj], w, A, Bm, amb);
                                bestm[w] = be.swap ? Bm : A;
                                bestm[W + w] = be.swap ? A : Bm;
                                bestm[2 * W + w] = amb;
                            }
                            __syncthreads();
                            int nk = 0, nbad = 0, ndiff = 0, la = -1, lb = -1;
                            for (long long p = tid; p < NP; p += B) {
                                int i, j;
                                pair_decode(p, T2, i, j);
                                int t1 = cand[i], t2 = cand[j];
                                bool pass = true;
                                for (int w = 0; w < W; ++w)
                                    pass &= ((mk1[(size_t)t1 * W + w] | mk1[(size_t)t2 * W + w]) & um[w]) == um[w];
                                if (!pass) continue;
                                double crit, rank;
                                score_clades(L, t1, t2, crit, rank);
                                if (!(brank - rank <= P.p.range)) continue;   // :636
                                TwoEval ev;
                                eval_two(L, tax, P, t1, t2, ev);
                                ++nk;
                                nbad += !ev.ok;
                                const bool unk = ev.c1 == tax.unknown || ev.c2 == tax.unknown;
                                bool same = true;   // meld_precheck: same synteny string (:671-676)
                                for (int w = 0; w < W; ++w) {
                                    u64 A, Bm, amb;
                                    letters(L, L.mk[P.amb_sel], unk, t1, t2, w, A, Bm, amb);
                                    same &= (ev.swap ? Bm : A) == bestm[w] &&
                                            (ev.swap ? A : Bm) == bestm[W + w] && amb == bestm[2 * W + w];
                                }
                                ndiff += !same;
                                la = lca2(tax, la, ev.c1);
                                lb = lca2(tax, lb, ev.c2);
                                memA[ev.t1] = 1;
                                memB[ev.t2] = 1;
                            }
                            block_sum3(nk, nbad, ndiff, sh);
                            la = block_lca(tax, la, sh.lca_a);
                            lb = block_lca(tax, lb, sh.lca_b);
                            bool have = true, ok = be.ok, melded = false;
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
                            if (have && ok) {
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
                                    int na = 0, nb = 0, z = 0;
                                    for (int t = tid; t < T; t += B) { na += memA[t]; nb += memB[t]; }
                                    block_sum3(na, nb, z, sh);
                                    r_na = na;
                                    r_nb = nb;
                                }
                                for (int i = tid; i < G; i += B) {
                                    u64 m = 1ull << (i & 63);
                                    int w = i >> 6;
                                    a.o.synteny[l0 + l_raw[i]] =
                                        ign[i] ? '~' : (bestm[2 * W + w] & m) ? '*' : (bestm[w] & m) ? 'A'
                                                 : (bestm[W + w] & m) ? 'B' : '!';
                                }
                            }
                        }
                    }
                    PH(7);
                    if (r_call != WFL_CALL_UNCLASSIFIED) {
                        // ---- melded members -> staging pool (tails, waafle_orgscorer.py:630,658-659)
                        if (r_na + r_nb > 0) {
                            if (tid == 0)
                                sh.mem_base = (long long)atomicAdd(&a.ctr->mem_pool_used,
                                                                   (unsigned long long)(r_na + r_nb));
                            __syncthreads();
                            r_mem = sh.mem_base;
                            for (int side = 0; side < 2; ++side) {
                                const unsigned char *mem = side ? memB : memA;
                                long long off = r_mem + (side ? r_na : 0);
                                int mbase = 0;
                                if ((side ? r_nb : r_na) == 0) continue;
                                for (int base = 0; base < T; base += B) {
                                    int t = base + tid, flag = t < T ? mem[t] : 0;
                                    int tot, pos = block_excl_scan(flag, tot, sh);
                                    if (flag && off + mbase + pos < a.o.mem_pool_cap)
                                        a.o.mem_pool[off + mbase + pos] = cl_id[t];
                                    mbase += tot;
                                }
                            }
                        }
                        break;
                    }
                    // not explained at this level: stop or lift (waafle_orgscorer.py:571-575)
                    if (T == 0 || hasroot) break;
                    if (iter >= 100) { r_status = 2; break; }   // :580-581
                    for (int r = tid; r < M; r += B) r_cl[r] = tax.parent[r_cl[r]];
                    ++lifts;
                    __syncthreads();
                }
                PH(8);
                if (tid == 0) {
                    for (int q = 0; q < 9; ++q) atomicAdd(&a.ctr->phase_cycles[q], ph[q]);
                    atomicAdd(&a.ctr->matched_pairs, (unsigned long long)M);
                    atomicAdd(&a.ctr->groups, (unsigned long long)n_groups);
                    atomicAdd(&a.ctr->levels, (unsigned long long)n_levels);
                    if (n_ptest) atomicAdd(&a.ctr->pairs_tested, (unsigned long long)n_ptest);
                    if (n_pscore) atomicAdd(&a.ctr->pairs_scored, (unsigned long long)n_pscore);
                    if (ar.all_smem) atomicAdd(&a.ctr->smem_contigs, 1ull);
                }
            }
        }

        if (overflow) {
            // replay with a larger slab: report a worst-case byte count for this contig
            r_status = 1;
            r_call = WFL_CALL_UNCLASSIFIED;
            if (tid == 0) {
                atomicMax(&a.ctr->slab_need_max, need_hint);
                atomicAdd(&a.ctr->n_overflow, 1ull);
            }
        }
        if (tid == 0) {
            if (r_status == 2) atomicAdd(&a.ctr->n_runaway, 1ull);
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
    }
}

void launch_score_kernel(const ScoreArgs &a, int grid, int threads, cudaStream_t s) {
    cudaFuncSetAttribute(wfl_score_contigs, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
    wfl_score_contigs<<<grid, threads, a.smem_bytes, s>>>(a);
}

}  // namespace wfl
