"""Generate waafle_b200/csrc/wfl_pipeline.cu from the monolithic warp kernel.

The pipeline runs the SAME per-phase code as wfl_score_warp.cu, cut at the phase markers into four
small kernels so that all warps of an SM execute the same phase at the same time (the monolithic
kernel is bound by instruction-cache misses: ncu gcc instruction-request throughput ~90 % of peak,
SM i-cache hit rate ~75 %).  Per-contig state lives in a global workspace pool between kernels.
Re-run this script after editing the phase code in wfl_score_warp.cu:
    python tools/gen_pipeline.py
"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = open(os.path.join(ROOT, "waafle_b200", "csrc", "wfl_score_warp.cu")).read()


def cut(start, end):
    i = SRC.index(start)
    j = SRC.index(end, i)
    return SRC[i:j]


def dedent(block, n):
    out = []
    for l in block.split("\n"):
        out.append(l[n:] if l.startswith(" " * n) else l)
    return "\n".join(out)


A = cut("        bool overflow = !ar.ok;\n        int G = 0;", "        const int W = (G + 63) >> 6;")
A = A.replace("        bool overflow = !ar.ok;\n", "        overflow = !arA.ok;\n")
B = cut("            // ---- leaf plans: host-built table", "            // ---- record arrays (locus-major)")
B = B.replace("            int M = 0;\n", "            M = 0;\n").replace("            int np_tot = 0;\n", "            np_tot = 0;\n")
C = cut("#pragma unroll 1\n                for (int i = lane; i < G; i += 32) {\n                    const int len = l_len[i];",
        "                bool have_base_ord = false;")
D = cut("                const int nwords = (tax.n_nodes + 31) >> 5;", "                PH(1);")
E = cut("                    if (!have_base_ord) {", "                    // ---- regroup: stable multisplit")
E = E.replace("                    if (!have_base_ord) {\n", "                    {\n").replace("                        have_base_ord = true;\n", "")
F = cut("                    // ---- regroup: stable multisplit", "                    PH(2);")
F = F.replace("have_base_ord ? base_ord : nullptr", "base_ord")
F = F.replace("                    const int Ngrp = spike ? ng - nu + G : ng;\n", "                    Ngrp = spike ? ng - nu + G : ng;\n")
F = F.replace("                    int t_unk = -1;\n", "                    t_unk = -1;\n")
F = F.replace("                    int ng = 0, nlt = 0, nu = 0;\n", "                    ng = 0; nlt = 0; nu = 0;\n")
F = re.sub(r"                    int \*cur = ar\.get<int>\(T \+ 2\);\n.*?double \*s_v = ar\.get<double>\(M\);\n",
           "                    DECL_CUR\n", F, flags=re.S)
assert "DECL_CUR" in F
F = re.sub(r"                    double \*g_score = ar\.get<double>\(Ngrp\);\n.*?\*gcur = ar\.get<int>\(2 \* G \+ 2\);\n",
           "                    DECL_GROUPS\n", F, flags=re.S)
G2 = cut("                    // ---- K2: envelope integral per group", "                    PH(3);")
G2 = G2.replace("have_base_ord", "true")
H = cut("                    // ---- K4: weak loci", "                    // ---- clade rows + gene bitmasks")
H = H.replace("                    int nun = 0;\n", "                    nun = 0;\n")
I = cut("                    // ---- clade rows + gene bitmasks", "                    PH(4);")
I = re.sub(r"                    int \*cl_go = ar\.get<int>\(T \+ 1\).*?Level \*Lp = ar\.get<Level>\(1\);\n",
           "                    DECL_CLADES\n", I, flags=re.S)
I = I.replace("                    int hasroot = 0;\n", "                    hasroot = 0;\n")
J = cut("                    // ---- K6: one-clade search", "                    PH(5);")
K = cut("                    if (bt >= 0) {\n                        // meld_one", "                    } else {\n                        // ---- K7: two-clade search")
K = K.replace("                    if (bt >= 0) {\n", "                    {\n") + "                    }\n"
L2 = cut("                        // ---- K7: two-clade search", "                    PH(6);")
L2 = L2[:L2.rindex("                    }\n")]          # drop the closing brace of the else
Mm = cut("                    if (r_call != WFL_CALL_UNCLASSIFIED) {\n                        // ---- melded members",
         "                    // not explained at this level: stop or lift")
N = cut("                    // not explained at this level: stop or lift", "                    PH(7);")

for name in ("F", "I", "L2", "N"):
    globals()[name] = globals()[name].replace("ar.", "al.")
L2 = L2.replace("""                        int *s_i = al.get<int>((size_t)scap), *s_j = al.get<int>((size_t)scap);
                        double *s_rank = al.get<double>((size_t)scap);
""", """                        Arena ax = al;
                        if (scap < NP) {   // not one slot per pair: take a fresh piece of the workspace pool
                            const unsigned long long want = 16ull * (unsigned long long)NP + 256ull;
                            unsigned long long off = 0;
                            if (lane == 0) off = atomicAdd(a.pool_used, (want + 255ull) & ~255ull);
                            off = __shfl_sync(FULL, off, 0);
                            if (off + want <= a.pool_cap) {
                                ax = Arena{a.pool + off, nullptr, (size_t)want, 0, 0, 0, true, true};
                                scap = NP;
                            }
                        }
                        int *s_i = ax.get<int>((size_t)scap), *s_j = ax.get<int>((size_t)scap);
                        double *s_rank = ax.get<double>((size_t)scap);
                        if (!ax.ok) { overflow = true; break; }
""")
L2 = L2.replace("""                        size_t room = (al.smem_cap - al.smem_used) > (al.slab_cap - al.slab_used)
                                          ? (al.smem_cap - al.smem_used) : (al.slab_cap - al.slab_used);
""", """                        size_t room = al.cap > al.used ? al.cap - al.used : 0;
""")
L2 = L2.replace("Arena ax = al;", "LinArena ax = al;").replace("ax = Arena{a.pool + off, nullptr, (size_t)want, 0, 0, 0, true, true};", "ax = LinArena{a.pool + off, (size_t)want, 0, true};")
L2 = L2.replace("long long scap = (long long)(room / 16) - 8;",
                "long long scap = ((long long)room - 4LL * (T + 1) - 8LL * (BITMAP_MAX_WORDS >= ((tax.n_nodes + 31) >> 5) ? ((tax.n_nodes + 31) >> 5) : 0) - 512) / 16;   // leave room for the lift's tables")
assert "LinArena ax = al;" in L2 and "al.cap - al.used" in L2 and "4LL * (T + 1)" in L2
Mm = Mm  # members block touches no arena

LEVEL_BLOCKS = dict(F=F, G2=G2, H=H, I=I, J=J, K=K, L2=L2, Mm=Mm, N=N)

TEMPLATE = r'''// GENERATED by tools/gen_pipeline.py from wfl_score_warp.cu -- do not edit by hand.
//
// The orgscorer path as a pipeline of seven small kernels (same per-phase code as the monolithic
// warp-per-contig kernel, one warp per contig in every kernel):
//
//   wfl_pipe_prepare : K1 match + record emission, K3 annotations, base order, distinct-clade table
//                      (the only kernel that streams the hit SoA from HBM)
//   wfl_pipe_regroup : K5 regroup (stable multisplit by clade rank, group table)
//   wfl_pipe_scores  : K2 envelope integrals (gene score per (clade, locus) group)
//   wfl_pipe_masks   : K4 weak loci, clade rows + gene bitmasks
//   wfl_pipe_one     : K6 one-clade search + meld; unresolved contigs go to the two-clade list
//   wfl_pipe_two     : K7/K8 two-clade search, meld, LGT filters; undecided contigs go to the lift list
//   wfl_pipe_lift    : K9 stop-or-lift (K5 lift of the clade table; next level's list)
//
// regroup/scores/masks/one/two/lift are launched once per taxonomy level over device-side work lists (no host sync in
// the level loop; an empty list makes the launch a no-op).  Between kernels a contig's state lives
// in a global workspace pool (regions A: loci, B: records + clade table, C: per-level arrays), carved
// by the same deterministic bump arena in every kernel.  Why: the monolithic kernel is bound by
// instruction fetch (profiles/r1_final_score_kernel_ncu.txt); here every kernel's hot code fits the
// SM instruction cache and all warps of an SM run the same phase.
#include "wfl_warp_common.cuh"

namespace wfl {

namespace {

#define PH(i) do { } while (0)

#ifndef WFL_PIPE_CPSM
#define WFL_PIPE_CPSM 32   // resident single-warp CTAs per SM the pipeline kernels are compiled for
#endif
// The small kernels (regroup, masks, one-clade, lift) wait on dependent workspace loads; they can be compiled
// with several independent warps per CTA to pass the 32-CTAs-per-SM limit (more warps, fewer registers each).
#ifndef WFL_LAT_WPC
#define WFL_LAT_WPC 1      // warps per CTA of the latency-bound kernels
#endif
#ifndef WFL_LAT_CPSM
#define WFL_LAT_CPSM 32    // their resident CTAs per SM
#endif

// Linear bump arena over one workspace region: region sizes are exact (loci_bytes / record_bytes) or
// bounds (level_bytes), so carving is a pointer increment plus one capacity compare.
struct LinArena {
    char *base;
    size_t cap, used;
    bool ok;
    template <class T>
    __device__ __forceinline__ T *get(size_t n) {
        T *p = reinterpret_cast<T *>(base + used);
        used += (n * sizeof(T) + 15) & ~size_t(15);
        ok = ok && used <= cap;
        return p;
    }
};

#define DECL_LOCI                                                                                   \
    int *l_lo = arA.get<int>(Graw), *l_len = arA.get<int>(Graw), *l_raw = arA.get<int>(Graw);       \
    int *l_base = arA.get<int>(Graw + 2);                                                           \
    const u16 **l_plan = arA.get<const u16 *>(Graw + 1);                                            \
    int *l_nleaf = arA.get<int>(Graw + 1);                                                          \
    signed char *l_str = arA.get<signed char>(Graw);                                                \
    u32 *l_k8 = arA.get<u32>(Graw + 1);

#define DECL_RECORDS                                                                                \
    u16 *plan_fb = ar.get<u16>(np_tot);                                                             \
    double *r_v = ar.get<double>(M);                                                                \
    int *r_a = ar.get<int>(M), *r_b = ar.get<int>(M), *r_t = ar.get<int>(M), *r_loc = ar.get<int>(M), \
        *r_hit = S > 0 ? ar.get<int>(M) : nullptr;                                                  \
    int *base_ord = ar.get<int>(M);                                                                 \
    double *maxv = ar.get<double>(G);                                                               \
    u64 *maxb = ar.get<u64>(G);                                                                     \
    u8 *ign = ar.get<u8>(G + 1);                                                                    \
    u64 *um = ar.get<u64>(W);                                                                       \
    u64 *annb = S > 0 ? ar.get<u64>((size_t)G * S) : nullptr;                                       \
    int *annw = S > 0 ? ar.get<int>((size_t)G * S) : nullptr;                                       \
    int *cl_id = ar.get<int>(M + 2);                                                                \
    int *ord = ar.get<int>(M);                                                                      \
    int *map_t = ar.get<int>(M + 2);                                                                \
    u8 *fo = ar.get<u8>(M + 2);

#define DECL_CUR                                                                                    \
    int *cur = al.get<int>(T + 2);                                                                  \
    int *s_a = al.get<int>(M), *s_b = al.get<int>(M);                                               \
    double *s_v = al.get<double>(M);

#define DECL_GROUPS                                                                                 \
    double *g_score = al.get<double>(Ngrp);                                                         \
    int *g_rs = al.get<int>(Ngrp + 1), *g_re = al.get<int>(Ngrp + 1), *g_loc = al.get<int>(Ngrp),   \
        *g_t = al.get<int>(Ngrp), *gs = al.get<int>(ng + 1), *g_perm = al.get<int>(Ngrp),           \
        *gcur = al.get<int>(2 * G + 2);

#define DECL_CLADES                                                                                 \
    int *cl_go = al.get<int>(T + 1), *cand = al.get<int>(T);                                        \
    double *cl_rank = al.get<double>(T), *cl_crit = al.get<double>(T);                              \
    u8 *cl_opt = al.get<u8>(T), *memA = al.get<u8>(T), *memB = al.get<u8>(T);                       \
    u64 *mk0 = al.get<u64>((size_t)T * W), *mk1 = al.get<u64>((size_t)T * W),                       \
        *mk2 = al.get<u64>((size_t)T * W);                                                          \
    u64 *bestm = al.get<u64>(3 * (size_t)W);                                                        \
    int *cl_par = al.get<int>(T);                                                                   \
    Level *Lp = al.get<Level>(1);

__device__ __forceinline__ size_t al16(size_t b) { return (b + 15) & ~size_t(15); }

// exact size of region A (the DECL_LOCI arrays)
__device__ __forceinline__ size_t loci_bytes(int Graw) {
    return 3 * al16(4 * (size_t)Graw) + al16(4 * ((size_t)Graw + 2)) + al16(8 * ((size_t)Graw + 1)) +
           2 * al16(4 * ((size_t)Graw + 1)) + al16((size_t)Graw);
}
// region B: DECL_RECORDS exactly, plus the temporary hash table of the clade-table build
__device__ __forceinline__ size_t record_bytes(int M, int G, int W, int S, int np_tot, int nwords, size_t *persist) {
    size_t m = (size_t)M, g = (size_t)G;
    size_t p = al16(2 * (size_t)np_tot) + al16(8 * m) + 4 * al16(4 * m) + (S > 0 ? al16(4 * m) : 0) + al16(4 * m) +
               2 * al16(8 * g) + al16(g + 1) + al16(8 * (size_t)W) +
               (S > 0 ? al16(8 * g * S) + al16(4 * g * S) : 0) + al16(4 * (m + 2)) + al16(4 * m) + al16(4 * (m + 2)) +
               al16(m + 2);
    *persist = p;
    size_t cap = 64;
    while (cap < 2 * (m + 1)) cap <<= 1;
    size_t hash_tmp = 2 * al16(4 * cap) + al16(4 * (m + 1)) + 64;
    size_t bitmap_tmp = nwords <= BITMAP_MAX_WORDS ? 2 * al16(4 * ((size_t)nwords + 1)) + 64 : 0;
    return p + (nwords <= BITMAP_MAX_WORDS ? bitmap_tmp : hash_tmp);
}
// region C: bound on the per-level arrays once the number of distinct clades T is known
__device__ __forceinline__ size_t level_bytes(int T, int M, int G, int W, int nwords) {
    size_t t = (size_t)T, ngb = (size_t)min((long long)M, (long long)T * G) + (size_t)G + 1;
    return al16(4 * (t + 2)) + 16 * (size_t)M + 48 + 40 * ngb + al16(4 * (2 * (size_t)G + 2)) + 160 + 27 * t + 24 * (size_t)W * t + 24 * (size_t)W +
           256 + sizeof(Level) + al16(4 * (t + 1)) + 4 * t + 16 * 25 + 6144 +
           (((size_t)nwords <= (size_t)BITMAP_MAX_WORDS) ? 8 * (size_t)nwords + 96 : 0);
}

struct ContigOut {
    int call, dir, c1, c2, lca, b1, b2, na, nb, status, lifts;
    long long mem;
    double crit, rank;
};

__device__ __noinline__ void write_result(const PipeArgs &a, long long c, const ContigOut &r) {
    if (r.status == 1) {
        atomicAdd(&a.ctr->n_overflow, 1ull);
        atomicMax(&a.ctr->slab_need_max, 1ull << 20);
    }
    if (r.status == 2) atomicAdd(&a.ctr->n_runaway, 1ull);
    if (r.status == 3) atomicAdd(&a.ctr->n_badinput, 1ull);
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
    a.o.status[c] = (uint8_t)r.status;
}

#define RESULT_LOCALS                                                                               \
    int r_call = WFL_CALL_UNCLASSIFIED, r_dir = 0, r_c1 = -1, r_c2 = -1, r_lca = -1, r_b1 = -1, r_b2 = -1, \
        r_na = 0, r_nb = 0, r_status = 0;                                                           \
    long long r_mem = 0;                                                                            \
    double r_crit = 0.0, r_rank = 0.0;

#define EMIT_RESULT(status_)                                                                        \
    do {                                                                                            \
        if (lane == 0) {                                                                            \
            ContigOut ro{r_call, r_dir, r_c1, r_c2, r_lca, r_b1, r_b2, r_na, r_nb, (status_), lifts, r_mem, r_crit, \
                         r_rank};                                                                   \
            write_result(a, c, ro);                                                                 \
            a.ctg[c].state = PIPE_DONE;                                                             \
        }                                                                                           \
    } while (0)

// fetch the next work item of a launch (lane 0 pops, the warp follows)
__device__ __forceinline__ long long pop_work(unsigned long long *wq, const int *list, const int *count, int lane) {
    long long c = -1;
    if (lane == 0) {
        unsigned long long w = atomicAdd(wq, 1ull);
        if ((long long)w < (long long)*count) c = list[w];
    }
    return __shfl_sync(FULL, c, 0);
}

// rebuild the arenas of a contig from its saved context (regions A+B persistent, B-leftover+C per level)
#define OPEN_CONTIG                                                                                 \
    const PipeCtg cx = a.ctg[c];                                                                    \
    const long long h0 = a.b.hit_off[c], l0 = a.b.locus_off[c];                                     \
    const int H = (int)(a.b.hit_off[c + 1] - h0), Graw = (int)(a.b.locus_off[c + 1] - l0);          \
    (void)H; (void)h0;                                                                              \
    const int G = cx.G, W = cx.W, M = cx.M, np_tot = cx.np_tot, iter = cx.iter;                     \
    int T = cx.T, lifts = cx.lifts;                                                                 \
    LinArena arA{a.pool + cx.offA, (size_t)cx.capA, 0, true};                                       \
    LinArena ar{a.pool + cx.offB, (size_t)cx.capB, 0, true};                                        \
    DECL_LOCI                                                                                       \
    DECL_RECORDS                                                                                    \
    LinArena al{a.pool + cx.offC, (size_t)cx.capC, 0, true};                                        \
    (void)l_lo; (void)l_base; (void)l_plan; (void)l_nleaf; (void)l_str; (void)l_k8; (void)plan_fb;  \
    (void)r_a; (void)r_b; (void)r_hit; (void)base_ord; (void)maxv; (void)annb; (void)annw; (void)map_t; (void)fo; \
    (void)r_v; (void)ord; (void)maxb; (void)iter;

}  // namespace

// ---------------------------------------------------------------------------------------------
// kernel 1: prepare
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32, WFL_PIPE_CPSM) wfl_pipe_prepare(const PipeArgs a) {
    const int lane = threadIdx.x;
    const DevParams &P = a.P;
    const DevTax &tax = a.t;
    const int S = P.p.n_systems;
    const bool spike = P.p.weak_loci == 2;
#pragma unroll 1
    for (;;) {
        long long c = -1;
        if (lane == 0) {
            unsigned long long w = atomicAdd(a.wq, 1ull);
            c = (long long)w < a.n_work ? a.work_base + (long long)w : -1;
        }
        c = __shfl_sync(FULL, c, 0);
        if (c < 0) break;
        const long long t_start = clock64();
        const long long h0 = a.b.hit_off[c], l0 = a.b.locus_off[c];
        const int H = (int)(a.b.hit_off[c + 1] - h0), Graw = (int)(a.b.locus_off[c + 1] - l0);
        RESULT_LOCALS
        bool bad_input = false, overflow = false;
        unsigned long long need_hint = 0;
        (void)need_hint;
        // region A
        const size_t capA = loci_bytes(Graw);
        unsigned long long offA = 0;
        if (lane == 0) offA = atomicAdd(a.pool_used, (unsigned long long)((capA + 255) & ~size_t(255)));
        offA = __shfl_sync(FULL, offA, 0);
        const bool fitsA = offA + capA <= a.pool_cap;
        LinArena arA{a.pool + (fitsA ? offA : 0), fitsA ? capA : 0, 0, true};
        LinArena ar{a.pool, 0, 0, true};
        DECL_LOCI
        if (!fitsA) arA.ok = false;
@A@
        __syncwarp();
        const int W = (G + 63) >> 6;
        int lifts = H > 0 ? P.p.jump_taxonomy : 0;
        int M = 0, np_tot = 0, T = 0;
        bool active = false;
        if (!overflow && H > 0 && G > 0) {
@B@
            // region B
            size_t persist = 0;
            const size_t capB = record_bytes(M, G, W, S, np_tot, (tax.n_nodes + 31) >> 5, &persist);
            unsigned long long offB = 0;
            if (lane == 0) offB = atomicAdd(a.pool_used, (unsigned long long)((capB + 255) & ~size_t(255)));
            offB = __shfl_sync(FULL, offB, 0);
            if (offB + capB <= a.pool_cap) {
                ar.base = a.pool + offB;
                ar.cap = capB;
            } else {
                ar.ok = false;
            }
            DECL_RECORDS
            overflow = !ar.ok;
            if (!overflow) {
@C@
@E@
@D@
            }
            const bool bad = __any_sync(FULL, bad_input);
            if (!overflow && !bad) {
                // region C: per-level arrays, sized now that T is known
                const size_t capC = level_bytes(T, M, G, W, (tax.n_nodes + 31) >> 5);
                unsigned long long offC = 0;
                if (lane == 0) offC = atomicAdd(a.pool_used, (unsigned long long)((capC + 255) & ~size_t(255)));
                offC = __shfl_sync(FULL, offC, 0);
                if (offC + capC > a.pool_cap) {
                    overflow = true;
                } else if (lane == 0) {
                    PipeCtg cx;
                    cx.offA = offA; cx.offB = offB; cx.offC = offC;
                    cx.capA = (unsigned)capA; cx.capB = (unsigned)capB; cx.capC = (unsigned)capC;
                    cx.G = G; cx.W = W; cx.M = M; cx.T = T; cx.np_tot = np_tot; cx.lifts = lifts; cx.iter = 0;
                    cx.Ngrp = cx.ng = cx.nlt = cx.nu = cx.nun = cx.hasroot = 0; cx.t_unk = -1;
                    cx.state = PIPE_ACTIVE;
                    a.ctg[c] = cx;
                    int slot = atomicAdd(a.cnt_act, 1);
                    a.list_act[slot] = (int)c;
                    atomicAdd(&a.ctr->matched_pairs, (unsigned long long)M);
                }
                active = !overflow;
            }
        }
        if (__any_sync(FULL, bad_input)) { overflow = false; r_status = 3; }
        if (!active) {
            if (overflow) r_status = 1;
            EMIT_RESULT(r_status);
        }
        if (lane == 0) atomicAdd(&a.ctr->phase_cycles[0], (unsigned long long)(clock64() - t_start));
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// kernel 2a: regroup (K5) -- records in (clade, locus) group order, group table
// ---------------------------------------------------------------------------------------------
// The scoring step is cut in three kernels (regroup | K2 | masks) for the same reason the pipeline
// exists: as one kernel it ran with an SM instruction-cache hit rate of 87 % and the GPC instruction
// cache at 92 % of its request throughput (profiles/r1_final2_pipeline_kernels_ncu.txt).
__global__ void __launch_bounds__(32 * WFL_LAT_WPC, WFL_LAT_CPSM) wfl_pipe_regroup(const PipeArgs a) {
    const int lane = threadIdx.x & 31;
    const DevParams &P = a.P;
    const DevTax &tax = a.t;
    const int S = P.p.n_systems;
    const bool spike = P.p.weak_loci == 2;
#pragma unroll 1
    for (;;) {
        const long long c = pop_work(a.wq, a.list_act, a.cnt_act, lane);
        if (c < 0) break;
        const long long t_start = clock64();
        OPEN_CONTIG
        RESULT_LOCALS
        bool overflow = false;
        int Ngrp = 0, ng = 0, nlt = 0, nu = 0, t_unk = -1;
        long long n_groups = 0;
#pragma unroll 1
        for (int once = 0; once < 1; ++once) {
@F@
            if (a.k2_desc != nullptr) {
                // ---- K2 work items: one descriptor per group with records, appended to the sub-batch's list
                const int nreal = spike ? Ngrp - G : Ngrp;   // the G spiked Unknown rows carry no records
                // reserve nreal slots with ONE atomic add (a compare-and-swap loop on this counter serialises the
                // whole grid).  A contig that does not fit fills its in-range slots with null items, which the K2
                // kernel skips, and is replayed by the warp kernel.
                unsigned long long dbase = 0;
                if (lane == 0) dbase = atomicAdd(&a.k2_meta->count, (unsigned long long)nreal);
                dbase = __shfl_sync(FULL, dbase, 0);
                if (dbase + (unsigned long long)nreal > a.k2_cap) {
                    K2Desc d{};
                    unsigned int nfill = 0;
#pragma unroll 1
                    for (unsigned long long i = dbase + lane; i < a.k2_cap; i += 32) {
                        a.k2_desc[i] = d;
                        a.k2_keys[i] = 0u;
                        ++nfill;
                    }
                    (void)nfill;
                    overflow = true;
                    break;
                }
                int wbase = 0;
#pragma unroll 1
                for (int gb = 0; gb < Ngrp; gb += 32) {
                    const int g = gb + lane;
                    const bool f = g < Ngrp && g_rs[g] >= 0;
                    const u32 m = __ballot_sync(FULL, f);
                    if (f) {
                        const int loc = g_loc[g], rs = g_rs[g], re = g_re[g], nleaf = l_nleaf[loc];
                        const int kc = re - rs <= 1 ? 0 : (re - rs <= 3 ? 1 : (re - rs <= 8 ? 2 : 3));
                        const int key = (min(nleaf, 127) << 2) | kc;
                        K2Desc d;
                        d.sa = s_a; d.plan = l_plan[loc]; d.out = &g_score[g];
                        d.maxb = cl_id[g_t[g]] != tax.unknown ? &maxb[loc] : nullptr;
                        d.d_sb = (int)(s_b - s_a);
                        d.d_sv = (int)(reinterpret_cast<const char *>(s_v) - reinterpret_cast<const char *>(s_a));
                        d.rs = rs; d.re = re; d.n = l_len[loc]; d.nleaf = nleaf; d.k8 = l_k8[loc]; d.key = (unsigned)key;
                        if (kc == 0) {
                            // a single record travels inside the descriptor (73 % of the groups at cfg2): K2 then
                            // never touches the record arrays of a contig it knows nothing else about
                            d.sa = reinterpret_cast<const int *>((size_t)__double_as_longlong(s_v[rs]));
                            d.d_sb = s_a[rs];
                            d.d_sv = s_b[rs];
                        }
                        const unsigned long long slot = dbase + wbase + __popc(m & lt_mask());
                        a.k2_desc[slot] = d;
                        a.k2_keys[slot] = (unsigned)key;
                    }
                    wbase += __popc(m);
                }
            }
        }
        if (lane == 0) {
            atomicAdd(&a.ctr->groups, (unsigned long long)n_groups);
            atomicAdd(&a.ctr->levels, 1ull);
            atomicAdd(&a.ctr->phase_cycles[2], (unsigned long long)(clock64() - t_start));
        }
        if (overflow) {
            EMIT_RESULT(1);
        } else if (lane == 0) {
            PipeCtg *cp = &a.ctg[c];
            cp->Ngrp = Ngrp; cp->ng = ng; cp->nlt = nlt; cp->nu = nu; cp->t_unk = t_unk;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// kernel 2b: K2 -- envelope integral of every (clade, locus) group, numpy-pairwise-exact
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32, WFL_PIPE_CPSM) wfl_pipe_scores(const PipeArgs a) {
    const int lane = threadIdx.x;
    const DevParams &P = a.P;
    const DevTax &tax = a.t;
    const int S = P.p.n_systems;
#pragma unroll 1
    for (;;) {
        const long long c = pop_work(a.wq, a.list_act, a.cnt_act, lane);
        if (c < 0) break;
        if (a.ctg[c].state != PIPE_ACTIVE) continue;
        const long long t_start = clock64();
        OPEN_CONTIG
        const int Ngrp = cx.Ngrp, ng = cx.ng;
        DECL_CUR
        DECL_GROUPS
        (void)cur; (void)gs; (void)T; (void)lifts; (void)l_raw; (void)l0; (void)ign; (void)um; (void)r_t; (void)r_loc;
@G2@
        if (lane == 0) atomicAdd(&a.ctr->phase_cycles[3], (unsigned long long)(clock64() - t_start));
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// kernels 2b': K2 over the sub-batch's GLOBAL group list.  Counting sort of the descriptors by
// (leaf count, record-count class), then lane per group: the lanes of a warp walk the same leaf plan
// whatever contig their group belongs to (per contig, lanes met different plans and a partly filled last
// round: 11 of 32 threads active per instruction).
// ---------------------------------------------------------------------------------------------
constexpr int K2_TILE = 4096;   // items per block pass of the counting sort (256 threads x 16)

// histogram of the sort keys: per-block counts in shared memory, one flush per block
__global__ void __launch_bounds__(256) wfl_pipe_k2hist(const PipeArgs a) {
    __shared__ unsigned int cnt[K2_KEYS];
    const unsigned long long n = a.k2_meta->count < a.k2_cap ? a.k2_meta->count : a.k2_cap;
    for (int k = threadIdx.x; k < K2_KEYS; k += 256) cnt[k] = 0;
    __syncthreads();
    for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256)
        atomicAdd(&cnt[a.k2_keys[i] & (K2_KEYS - 1)], 1u);
    __syncthreads();
    for (int k = threadIdx.x; k < K2_KEYS; k += 256)
        if (cnt[k]) atomicAdd(&a.k2_meta->hist[k], cnt[k]);
}

__global__ void wfl_pipe_k2scan(const PipeArgs a) {
    __shared__ unsigned int part[K2_KEYS];
    const int t = threadIdx.x;   // K2_KEYS threads
    const unsigned int h = a.k2_meta->hist[t];
    part[t] = h;
    __syncthreads();
    for (int o = 1; o < K2_KEYS; o <<= 1) {
        unsigned int v = t >= o ? part[t - o] : 0u;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    a.k2_meta->cursor[t] = part[t] - h;
}

// scatter: a block ranks a tile of items per key in shared memory, reserves one range per key with a single
// global atomic, and writes the item indices to their sorted positions
__global__ void __launch_bounds__(256) wfl_pipe_k2scatter(const PipeArgs a) {
    __shared__ unsigned int cnt[K2_KEYS], base[K2_KEYS];
    const unsigned long long n = a.k2_meta->count < a.k2_cap ? a.k2_meta->count : a.k2_cap;
    for (unsigned long long t0 = (unsigned long long)blockIdx.x * K2_TILE; t0 < n; t0 += (unsigned long long)gridDim.x * K2_TILE) {
        for (int k = threadIdx.x; k < K2_KEYS; k += 256) cnt[k] = 0;
        __syncthreads();
        unsigned int key[K2_TILE / 256], rk[K2_TILE / 256];
#pragma unroll
        for (int q = 0; q < K2_TILE / 256; ++q) {
            const unsigned long long i = t0 + (unsigned long long)q * 256 + threadIdx.x;
            key[q] = i < n ? (a.k2_keys[i] & (K2_KEYS - 1)) : 0xffffffffu;
            rk[q] = i < n ? atomicAdd(&cnt[key[q]], 1u) : 0u;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < K2_KEYS; k += 256)
            if (cnt[k]) base[k] = atomicAdd(&a.k2_meta->cursor[k], cnt[k]);
        __syncthreads();
#pragma unroll
        for (int q = 0; q < K2_TILE / 256; ++q) {
            const unsigned long long i = t0 + (unsigned long long)q * 256 + threadIdx.x;
            if (i < n) a.k2_order[base[key[q]] + rk[q]] = (unsigned int)i;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(32, WFL_PIPE_CPSM) wfl_pipe_k2(const PipeArgs a) {
    const int lane = threadIdx.x;
    const unsigned long long n = a.k2_meta->count < a.k2_cap ? a.k2_meta->count : a.k2_cap;
#pragma unroll 1
    for (;;) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&a.k2_meta->take, 32ull);
        base = __shfl_sync(FULL, base, 0);
        if (base >= n) break;
        const long long t_start = clock64();
        const unsigned long long i = base + lane;
        if (i < n) {
            const K2Desc d = a.k2_desc[a.k2_order[i]];
            if (d.out != nullptr) {
            double sc;
            if ((d.key & 3u) == 0u) {   // single record, inline: slice [d_sb, d_sv), score in the bits of `sa`
                int a1 = d.d_sb, b1 = d.d_sv;
                double v1 = __longlong_as_double((long long)(size_t)d.sa);
                sc = group_mean(&a1, &b1, &v1, 0, 1, d.n, true, d.k8, d.plan, d.nleaf);
            } else {
                sc = group_mean(d.sa, d.sa + d.d_sb,
                                reinterpret_cast<const double *>(reinterpret_cast<const char *>(d.sa) + d.d_sv),
                                d.rs, d.re, d.n, true, d.k8, d.plan, d.nleaf);
            }
            *d.out = sc;
            if (d.maxb != nullptr) atomicMax(d.maxb, dbits(sc));   // waafle_orgscorer.py:409-411
            }
        }
        if (lane == 0) atomicAdd(&a.ctr->phase_cycles[3], (unsigned long long)(clock64() - t_start));
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// kernel 2c: weak loci (K4), clade rows and gene bitmasks
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * WFL_LAT_WPC, WFL_LAT_CPSM) wfl_pipe_masks(const PipeArgs a) {
    const int lane = threadIdx.x & 31;
    const DevParams &P = a.P;
    const DevTax &tax = a.t;
    const int S = P.p.n_systems;
    const bool spike = P.p.weak_loci == 2;
#pragma unroll 1
    for (;;) {
        const long long c = pop_work(a.wq, a.list_act, a.cnt_act, lane);
        if (c < 0) break;
        if (a.ctg[c].state != PIPE_ACTIVE) continue;
        const long long t_start = clock64();
        OPEN_CONTIG
        RESULT_LOCALS
        const int Ngrp = cx.Ngrp, ng = cx.ng, nlt = cx.nlt;
        DECL_CUR
        DECL_GROUPS
        (void)cur; (void)s_a; (void)s_b; (void)s_v; (void)g_rs; (void)g_re; (void)gs; (void)g_perm; (void)gcur;
        (void)r_t; (void)r_loc;
        bool overflow = false, cont_ok = false;
        int nun = 0, hasroot = 0;
#pragma unroll 1
        for (int once = 0; once < 1; ++once) {
@H@
@I@
            cont_ok = true;
        }
        if (lane == 0) atomicAdd(&a.ctr->phase_cycles[4], (unsigned long long)(clock64() - t_start));
        if (overflow) {
            EMIT_RESULT(1);
        } else if (!cont_ok) {
            EMIT_RESULT(0);   // "empty" contig: every locus ignored at the first level (unclassified)
        } else if (lane == 0) {
            PipeCtg *cp = &a.ctg[c];
            cp->nun = nun;
            cp->hasroot = hasroot;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// kernel 3: one-clade search
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * WFL_LAT_WPC, WFL_LAT_CPSM) wfl_pipe_one(const PipeArgs a) {
    const int lane = threadIdx.x & 31;
    const DevParams &P = a.P;
    const DevTax &tax = a.t;
    const int S = P.p.n_systems;
#pragma unroll 1
    for (;;) {
        const long long c = pop_work(a.wq, a.list_act, a.cnt_act, lane);
        if (c < 0) break;
        if (a.ctg[c].state != PIPE_ACTIVE) continue;
        const long long t_start = clock64();
        OPEN_CONTIG
        RESULT_LOCALS
        const int Ngrp = cx.Ngrp, ng = cx.ng;
        DECL_CUR
        DECL_GROUPS
        DECL_CLADES
        (void)cur; (void)s_a; (void)s_b; (void)s_v; (void)g_score; (void)g_rs; (void)g_re; (void)g_loc; (void)g_t; (void)gs; (void)g_perm; (void)gcur;
        (void)cand; (void)memB; (void)mk1; (void)mk2; (void)bestm; (void)cl_go; (void)cl_par;
        const Level &L = *Lp;
        (void)L;
@J@
        if (bt >= 0) {
@K@
@Mm_ONE@
            EMIT_RESULT(0);
        } else if (lane == 0) {
            int slot = atomicAdd(a.cnt_two, 1);
            a.list_two[slot] = (int)c;
        }
        if (lane == 0) atomicAdd(&a.ctr->phase_cycles[5], (unsigned long long)(clock64() - t_start));
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// kernel 4a: two-clade search and decision; undecided contigs go to the lift list
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32, WFL_PIPE_CPSM) wfl_pipe_two(const PipeArgs a) {
    const int lane = threadIdx.x;
    const DevParams &P = a.P;
    const DevTax &tax = a.t;
    const int S = P.p.n_systems;
    const bool spike = P.p.weak_loci == 2;
#pragma unroll 1
    for (;;) {
        const long long c = pop_work(a.wq, a.list_two, a.cnt_two, lane);
        if (c < 0) break;
        const long long t_start = clock64();
        OPEN_CONTIG
        RESULT_LOCALS
        const int Ngrp = cx.Ngrp, ng = cx.ng, nu = cx.nu, t_unk = cx.t_unk, hasroot = cx.hasroot;
        DECL_CUR
        DECL_GROUPS
        DECL_CLADES
        (void)cur; (void)s_a; (void)s_b; (void)s_v; (void)g_score; (void)g_rs; (void)g_re; (void)g_loc; (void)g_t; (void)gs; (void)g_perm; (void)gcur;
        (void)cl_rank; (void)cl_crit; (void)cl_opt; (void)mk0; (void)mk2; (void)cl_go; (void)cl_par;
        (void)nu; (void)t_unk; (void)hasroot; (void)spike; (void)T; (void)r_t; (void)r_loc;
        const Level &L = *Lp;
        bool overflow = false, undecided = false;
        unsigned long long need_hint = 0;
        (void)need_hint;
        long long n_ptest = 0, n_pscore = 0;
#pragma unroll 1
        for (int once = 0; once < 1; ++once) {
@L2@
@Mm@
            undecided = true;
        }
        if (lane == 0) {
            if (n_ptest) atomicAdd(&a.ctr->pairs_tested, (unsigned long long)n_ptest);
            if (n_pscore) atomicAdd(&a.ctr->pairs_scored, (unsigned long long)n_pscore);
            atomicAdd(&a.ctr->phase_cycles[6], (unsigned long long)(clock64() - t_start));
        }
        if (overflow) {
            r_call = WFL_CALL_UNCLASSIFIED;
            r_na = r_nb = 0;
            EMIT_RESULT(1);
        } else if (!undecided) {
            EMIT_RESULT(r_status);
        } else if (lane == 0) {
            int slot = atomicAdd(a.cnt_lift, 1);
            a.list_lift[slot] = (int)c;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// kernel 4b: stop (root reached / nothing left) or lift the clade table to the parents (K9, K5)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * WFL_LAT_WPC, WFL_LAT_CPSM) wfl_pipe_lift(const PipeArgs a) {
    const int lane = threadIdx.x & 31;
    const DevParams &P = a.P;
    const DevTax &tax = a.t;
    const int S = P.p.n_systems;
    const bool spike = P.p.weak_loci == 2;
#pragma unroll 1
    for (;;) {
        const long long c = pop_work(a.wq, a.list_lift, a.cnt_lift, lane);
        if (c < 0) break;
        const long long t_start = clock64();
        OPEN_CONTIG
        RESULT_LOCALS
        const int Ngrp = cx.Ngrp, ng = cx.ng, nu = cx.nu, t_unk = cx.t_unk, hasroot = cx.hasroot;
        DECL_CUR
        DECL_GROUPS
        DECL_CLADES
        (void)cur; (void)s_a; (void)s_b; (void)s_v; (void)g_score; (void)g_rs; (void)g_re; (void)g_loc; (void)g_t; (void)gs; (void)g_perm; (void)gcur;
        (void)cl_rank; (void)cl_crit; (void)cl_opt; (void)mk0; (void)mk1; (void)mk2; (void)cl_go; (void)cl_par; (void)cand;
        (void)memA; (void)memB; (void)bestm; (void)Lp; (void)nu; (void)t_unk; (void)r_loc; (void)l_raw; (void)l0; (void)ign; (void)um;
        bool overflow = false, lifted = false;
#pragma unroll 1
        for (int once = 0; once < 1; ++once) {
@N@
            lifted = true;
        }
        if (lane == 0) atomicAdd(&a.ctr->phase_cycles[7], (unsigned long long)(clock64() - t_start));
        if (overflow) {
            r_call = WFL_CALL_UNCLASSIFIED;
            r_na = r_nb = 0;
            EMIT_RESULT(1);
        } else if (!lifted) {
            EMIT_RESULT(r_status);
        } else if (lane == 0) {
            PipeCtg *cp = &a.ctg[c];
            cp->T = T;
            cp->lifts = lifts;
            cp->iter = iter + 1;
            int slot = atomicAdd(a.cnt_next, 1);
            a.list_next[slot] = (int)c;
        }
        __syncwarp();
    }
}

// Contigs still on the work list after the last level (deeper than the host's level bound): hand
// them to the monolithic kernel by marking them for replay.
__global__ void wfl_pipe_leftover(const PipeArgs a) {
    const int n = *a.cnt_act;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = a.list_act[i];
        a.o.status[c] = 1;
        a.o.call[c] = WFL_CALL_UNCLASSIFIED;
        a.o.n_mem_a[c] = a.o.n_mem_b[c] = 0;
        atomicAdd(&a.ctr->n_overflow, 1ull);
        atomicMax(&a.ctr->slab_need_max, 1ull << 20);
    }
}

// ---------------------------------------------------------------------------------------------
// host side: launch sequence for one sub-batch
// ---------------------------------------------------------------------------------------------
void launch_pipe_prepare(const PipeArgs &a, int grid, cudaStream_t s) { wfl_pipe_prepare<<<grid, 32, 0, s>>>(a); }
void launch_pipe_regroup(const PipeArgs &a, int grid, cudaStream_t s) { wfl_pipe_regroup<<<grid / WFL_PIPE_CPSM * WFL_LAT_CPSM, 32 * WFL_LAT_WPC, 0, s>>>(a); }
void launch_pipe_scores(const PipeArgs &a, int grid, cudaStream_t s) { wfl_pipe_scores<<<grid, 32, 0, s>>>(a); }
void launch_pipe_k2sort(const PipeArgs &a, int grid, cudaStream_t s) {
    wfl_pipe_k2hist<<<grid / 8 + 1, 256, 0, s>>>(a);
    wfl_pipe_k2scan<<<1, K2_KEYS, 0, s>>>(a);
    wfl_pipe_k2scatter<<<grid / 8 + 1, 256, 0, s>>>(a);
}
void launch_pipe_k2(const PipeArgs &a, int grid, cudaStream_t s) { wfl_pipe_k2<<<grid, 32, 0, s>>>(a); }
void launch_pipe_masks(const PipeArgs &a, int grid, cudaStream_t s) { wfl_pipe_masks<<<grid / WFL_PIPE_CPSM * WFL_LAT_CPSM, 32 * WFL_LAT_WPC, 0, s>>>(a); }
void launch_pipe_one(const PipeArgs &a, int grid, cudaStream_t s) { wfl_pipe_one<<<grid / WFL_PIPE_CPSM * WFL_LAT_CPSM, 32 * WFL_LAT_WPC, 0, s>>>(a); }
void launch_pipe_two(const PipeArgs &a, int grid, cudaStream_t s) { wfl_pipe_two<<<grid, 32, 0, s>>>(a); }
void launch_pipe_lift(const PipeArgs &a, int grid, cudaStream_t s) { wfl_pipe_lift<<<grid / WFL_PIPE_CPSM * WFL_LAT_CPSM, 32 * WFL_LAT_WPC, 0, s>>>(a); }
int pipe_ctas_per_sm() { return WFL_PIPE_CPSM; }
void launch_pipe_leftover(const PipeArgs &a, cudaStream_t s) { wfl_pipe_leftover<<<64, 256, 0, s>>>(a); }

}  // namespace wfl
'''


def indent_to(block, cur, new):
    """Re-indent a block whose base indentation is `cur` spaces to `new` spaces."""
    out = []
    for l in block.rstrip("\n").split("\n"):
        if l.startswith("#pragma") or l.strip() == "":
            out.append(l)
        elif l.startswith(" " * cur):
            out.append(" " * new + l[cur:])
        else:
            out.append(l)
    return "\n".join(out)


Mm_one = Mm.replace("                        break;\n", "")
subs = {
    "@A@": indent_to(A, 8, 8),
    "@B@": indent_to(B, 12, 12),
    "@C@": indent_to(C, 16, 16),
    "@E@": indent_to(E, 20, 16),
    "@D@": indent_to(D, 16, 16),
    "@F@": indent_to(F, 20, 12),
    "@G2@": indent_to(G2, 20, 12),
    "@H@": indent_to(H, 20, 12),
    "@I@": indent_to(I, 20, 12),
    "@J@": indent_to(J, 20, 8),
    "@K@": indent_to(K, 20, 12),
    "@Mm_ONE@": indent_to(Mm_one, 20, 12),
    "@L2@": indent_to(L2, 24, 12),
    "@Mm@": indent_to(Mm, 20, 12),
    "@N@": indent_to(N, 20, 12),
}
out = TEMPLATE
for k, v in subs.items():
    assert k in out, k
    out = out.replace(k, v)
open(os.path.join(ROOT, "waafle_b200", "csrc", "wfl_pipeline.cu"), "w").write(out)
print("wrote wfl_pipeline.cu:", out.count("\n"), "lines")
