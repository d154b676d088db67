/*
 * TEST INFRASTRUCTURE ONLY -- literal C restatement of the waafle_orgscorer per-contig engine.
 *
 * Purpose: a FAST checker for parity tests at BASELINE.json's full sizes (100 k contigs in seconds on
 * the host cores), where the numpy oracle (oracle/orgscorer_oracle.py, the oracle of record, pinned
 * against the unmodified reference) would take half an hour.  It is deliberately literal: per-site
 * double arrays exactly like the reference (waafle/waafle_orgscorer.py:371-382), numpy's pairwise
 * summation written out (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum), no shortcuts
 * shared with the CUDA engine.  tests/test_c_oracle.py pins it against the numpy oracle.
 *
 * Nothing in the product (waafle_b200/) links or calls this file.  Cites: OS = waafle/waafle_orgscorer.py,
 * UT = waafle/utils.py under /root/reference.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -pthread; no fast-math: rounding must match numpy).
 * Threads: WFL_ORACLE_THREADS (default: online cores); contigs are independent (OS:943-960).
 */
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/waafle_b200.h"

typedef struct {
    int32_t n_nodes, root, unknown;
    const int32_t *parent, *depth, *leaf_count;
    const uint8_t *listed;
} otax;

/* numpy pairwise sum of n contiguous doubles */
static double pw(const double *a, long n) {
    if (n < 8) {
        double res = 0.;
        for (long i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        long i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pw(a, n2) + pw(a + n2, n - n2);
    }
}
static double np_mean(const double *a, long n) { return pw(a, n) / (double)n; }

static int lca2(const otax *t, int a, int b) { /* UT:401-411 */
    while (t->depth[a] > t->depth[b]) a = t->parent[a];
    while (t->depth[b] > t->depth[a]) b = t->parent[b];
    while (a != b) { a = t->parent[a]; b = t->parent[b]; }
    return a;
}

/* one (clade, locus) site array */
typedef struct {
    int clade, locus;
    double *site;
} entry;

typedef struct {
    int G;
    int *llo, *llen, *lraw;
    signed char *lstr;
    entry *e;
    int ne, cap;
    /* current level */
    int T;
    int *clade;      /* ascending */
    double *gene;    /* [T][G] */
    int spiked;      /* index of the spiked Unknown row or -1 */
    unsigned char *ign;
    int nmask;
    int *mask;       /* non-ignored loci, ascending */
} contig;

static int cmp_int(const void *a, const void *b) { return (*(const int *)a > *(const int *)b) - (*(const int *)a < *(const int *)b); }

static entry *get_entry(contig *C, int clade, int locus, int create) {
    for (int i = 0; i < C->ne; i++)
        if (C->e[i].clade == clade && C->e[i].locus == locus) return &C->e[i];
    if (!create) return NULL;
    if (C->ne == C->cap) {
        C->cap = C->cap ? 2 * C->cap : 64;
        C->e = (entry *)realloc(C->e, (size_t)C->cap * sizeof(entry));
    }
    entry *x = &C->e[C->ne++];
    x->clade = clade;
    x->locus = locus;
    x->site = (double *)calloc((size_t)C->llen[locus], sizeof(double)); /* np.zeros, OS:381 */
    return x;
}

/* update_gene_scores, OS:394-429 */
static void update_gene_scores(contig *C, const wfl_params *P, const otax *tax) {
    int G = C->G;
    int *cl = (int *)malloc((size_t)(C->ne + 1) * sizeof(int));
    int T = 0;
    for (int i = 0; i < C->ne; i++) cl[T++] = C->e[i].clade;
    qsort(cl, (size_t)T, sizeof(int), cmp_int);
    int u = 0;
    for (int i = 0; i < T; i++)
        if (i == 0 || cl[i] != cl[i - 1]) cl[u++] = cl[i];
    T = u;
    int has_unknown_row = 0;
    for (int i = 0; i < T; i++) has_unknown_row |= cl[i] == tax->unknown;
    int spike = P->weak_loci == 2;
    if (spike && !has_unknown_row) {
        int pos = 0;
        while (pos < T && cl[pos] < tax->unknown) pos++;
        memmove(cl + pos + 1, cl + pos, (size_t)(T - pos) * sizeof(int));
        cl[pos] = tax->unknown;
        T++;
    }
    free(C->clade);
    free(C->gene);
    C->clade = cl;
    C->T = T;
    C->gene = (double *)calloc((size_t)T * (size_t)(G > 0 ? G : 1), sizeof(double));
    for (int i = 0; i < C->ne; i++) {
        int t = 0;
        while (cl[t] != C->e[i].clade) t++;
        C->gene[(size_t)t * G + C->e[i].locus] = np_mean(C->e[i].site, C->llen[C->e[i].locus]);
    }
    double *maxes = (double *)calloc((size_t)(G > 0 ? G : 1), sizeof(double));
    for (int t = 0; t < T; t++)
        if (cl[t] != tax->unknown)
            for (int i = 0; i < G; i++)
                if (C->gene[(size_t)t * G + i] > maxes[i]) maxes[i] = C->gene[(size_t)t * G + i];
    C->spiked = -1;
    C->nmask = 0;
    double min_thr = P->k1 < P->k2 ? P->k1 : P->k2;
    if (P->weak_loci == 2) {
        int t = 0;
        while (cl[t] != tax->unknown) t++;
        for (int i = 0; i < G; i++) C->gene[(size_t)t * G + i] = 1 - maxes[i]; /* OS:417 */
        C->spiked = t;
    }
    for (int i = 0; i < G; i++) {
        if (P->weak_loci == 0) C->ign[i] = !(maxes[i] >= min_thr); /* OS:421-426 */
        else C->ign[i] = 0;
        if (!C->ign[i]) C->mask[C->nmask++] = i;
    }
    free(maxes);
}

/* raise_taxonomy, OS:431-445 */
static void raise_taxonomy(contig *C, const wfl_params *P, const otax *tax) {
    entry *old = C->e;
    int nold = C->ne;
    C->e = NULL;
    C->ne = C->cap = 0;
    for (int i = 0; i < nold; i++) {
        int parent = tax->parent[old[i].clade];
        entry *x = get_entry(C, parent, old[i].locus, 1);
        int len = C->llen[old[i].locus];
        for (int s = 0; s < len; s++)
            if (old[i].site[s] > x->site[s]) x->site[s] = old[i].site[s]; /* np.maximum */
        free(old[i].site);
    }
    free(old);
    update_gene_scores(C, P, tax);
}

/* Contig.score, OS:447-461 */
static void score(const contig *C, int t1, int t2, double *crit, double *rank, double *buf) {
    int G = C->G;
    for (int k = 0; k < C->nmask; k++) {
        int i = C->mask[k];
        double v = C->gene[(size_t)t1 * G + i];
        if (t2 >= 0 && C->gene[(size_t)t2 * G + i] > v) v = C->gene[(size_t)t2 * G + i];
        buf[k] = v;
    }
    double mn = buf[0];
    for (int k = 1; k < C->nmask; k++)
        if (buf[k] < mn) mn = buf[k];
    *crit = mn;
    *rank = np_mean(buf, C->nmask);
}

typedef struct {
    int ok, t1, t2, c1, c2, dir, recip;
    double crit, rank;
    char *syn;
} option;

static double tri(int mode, double k1, double k2) { return mode == 0 ? 1e-6 : (mode == 1 ? (k1 < k2 ? k1 : k2) : (k1 > k2 ? k1 : k2)); }

/* set_synteny_two, OS:511-545 */
static void synteny_two(const contig *C, const wfl_params *P, const otax *tax, option *o) {
    int G = C->G;
    double k_amb = tri(P->ambiguous_threshold, P->k1, P->k2);
    int unk = C->clade[o->t1] == tax->unknown || C->clade[o->t2] == tax->unknown;
    for (int i = 0; i < G; i++) {
        double s1 = C->gene[(size_t)o->t1 * G + i], s2 = C->gene[(size_t)o->t2 * G + i];
        char ch;
        if (C->ign[i]) ch = '~';
        else if ((s1 < s2 ? s1 : s2) >= k_amb && !unk) ch = '*';
        else if (s1 >= P->k2) ch = 'A';
        else if (s2 >= P->k2) ch = 'B';
        else ch = '!';
        o->syn[i] = ch;
    }
    o->syn[G] = 0;
    /* "^[^A]*B" */
    int swap = 0;
    for (int i = 0; i < G; i++) {
        if (o->syn[i] == 'A') break;
        if (o->syn[i] == 'B') { swap = 1; break; }
    }
    if (swap) {
        int t = o->t1; o->t1 = o->t2; o->t2 = t;
        for (int i = 0; i < G; i++) o->syn[i] = o->syn[i] == 'A' ? 'B' : (o->syn[i] == 'B' ? 'A' : o->syn[i]);
    }
    o->c1 = C->clade[o->t1];
    o->c2 = C->clade[o->t2];
    /* "^A+B+A+$" on the string without '~' */
    int st = 0;
    for (int i = 0; i < G && st >= 0; i++) {
        char ch = o->syn[i];
        if (ch == '~') continue;
        if (ch == 'A') st = (st == 0 || st == 1) ? 1 : 3;
        else if (ch == 'B') st = (st == 1 || st == 2) ? 2 : -1;
        else st = -1;
    }
    o->dir = st == 3;
    o->recip = o->dir ? o->c2 : -1;
}

/* apply_lgt_checks, OS:678-744 */
static void lgt_checks(const contig *C, const wfl_params *P, const otax *tax, option *o) {
    int G = C->G;
    long total = 0, amb = 0;
    int nA = 0, nB = 0;
    for (int i = 0; i < G; i++) {
        char ch = o->syn[i];
        if (ch == 'A' || ch == 'B' || ch == '*') {
            total += C->llen[i];
            if (ch == '*') amb += C->llen[i];
        }
        nA += ch == 'A';
        nB += ch == 'B';
    }
    if ((double)amb / (double)total > P->ambiguous_fraction) o->ok = 0;
    if (P->clade_genes >= 0 && (nA < nB ? nA : nB) < P->clade_genes) o->ok = 0;
    if (P->clade_leaves >= 0) {
        int lc = o->recip >= 0 ? tax->leaf_count[o->recip]
                               : (tax->leaf_count[o->c1] < tax->leaf_count[o->c2] ? tax->leaf_count[o->c1] : tax->leaf_count[o->c2]);
        if (lc < P->clade_leaves) o->ok = 0;
    }
    if (P->sister_penalty != 0) {
        double thr = P->sister_penalty == 1 ? (P->k1 > P->k2 ? P->k1 : P->k2) : (P->k1 < P->k2 ? P->k1 : P->k2);
        int badA = 0, badB = 0;
        for (int i = 0; i < G; i++) {
            char ch = o->syn[i];
            if (ch != 'A' && ch != 'B') continue;
            /* a B locus is penalised by clade1's sisters, an A locus by clade2's (OS:724-726) */
            int base = ch == 'B' ? o->c1 : o->c2, other = ch == 'B' ? o->c2 : o->c1;
            int p = tax->parent[base];
            for (int t = 0; t < C->T; t++) {
                int x = C->clade[t];
                if (x == base || x == other || !tax->listed[x] || tax->parent[x] != p) continue;
                if (C->gene[(size_t)t * G + i] >= thr) { if (ch == 'A') badA = 1; else badB = 1; }
            }
        }
        if (o->recip >= 0 ? badB : (badA || badB)) o->ok = 0;
    }
}

static void one_contig(const wfl_params *P, const otax *tax, const wfl_batch *in, wfl_results *out, int64_t c,
                       int32_t **mem_out, int *nmem_a, int *nmem_b) {
    const int64_t h0 = in->hit_off[c], h1 = in->hit_off[c + 1], l0 = in->locus_off[c], l1 = in->locus_off[c + 1];
    const int Graw = (int)(l1 - l0), S = P->n_systems;
    contig C;
    memset(&C, 0, sizeof C);
    C.llo = (int *)malloc((size_t)(Graw + 1) * sizeof(int));
    C.llen = (int *)malloc((size_t)(Graw + 1) * sizeof(int));
    C.lraw = (int *)malloc((size_t)(Graw + 1) * sizeof(int));
    C.lstr = (signed char *)malloc((size_t)(Graw + 1));
    C.ign = (unsigned char *)calloc((size_t)(Graw + 1), 1);
    C.mask = (int *)malloc((size_t)(Graw + 1) * sizeof(int));
    C.spiked = -1;
    for (int j = 0; j < Graw; j++) { /* attach_loci, OS:348-352 */
        int s = in->locus_start[l0 + j], e = in->locus_end[l0 + j];
        int len = abs(e - s) + 1;
        out->locus_flags[l0 + j] = 0;
        out->synteny[l0 + j] = 0;
        for (int q = 0; q < S; q++) out->ann_winner[(l0 + j) * S + q] = -1;
        if ((double)len >= P->min_gene_length) {
            C.llo[C.G] = s < e ? s : e;
            C.llen[C.G] = len;
            C.lraw[C.G] = j;
            C.lstr[C.G] = in->locus_strand[l0 + j];
            C.G++;
        }
    }
    const int G = C.G;
    double ann_thr = tri(P->annotation_threshold, P->k1, P->k2);
    double *ann_score = (double *)malloc((size_t)(G * S + 1) * sizeof(double));
    for (int i = 0; i < G * S; i++) ann_score[i] = ann_thr;
    int lifts = 0;
    option best;
    memset(&best, 0, sizeof best);
    int have_one = 0, have_two = 0;
    int *members = NULL;
    *nmem_a = *nmem_b = 0;
    char *bsyn = (char *)calloc((size_t)G + 2, 1);
    int out_c1 = -1, out_c2 = -1, out_b1 = -1, out_b2 = -1;

    if (h1 > h0) {
        /* attach_hits + score_hit, OS:359-392 */
        for (int64_t h = h0; h < h1; h++) {
            if (!(in->hit_scov[h] >= P->min_scov)) continue;
            int q1 = in->hit_qstart[h], q2 = in->hit_qend[h];
            int a1 = q1 < q2 ? q1 : q2, a2 = q1 < q2 ? q2 : q1;
            for (int i = 0; i < G; i++) {
                if (P->stranded && in->hit_strand[h] != C.lstr[i]) continue;
                int b1 = C.llo[i], b2 = b1 + C.llen[i] - 1;
                double ov = 0.0; /* UT:487-500 */
                if (!(b1 > a2 || a1 > b2)) {
                    int inl = a1 > b1 ? a1 : b1, inr = a2 < b2 ? a2 : b2;
                    int den = (a2 - a1 + 1) < (b2 - b1 + 1) ? (a2 - a1 + 1) : (b2 - b1 + 1);
                    ov = (double)(inr - inl + 1) / (double)den;
                }
                if (!(ov >= P->min_overlap)) continue;
                int len = C.llen[i];
                int s1 = a1 - b1 > 0 ? a1 - b1 : 0;
                int e1 = (a2 - b1 < len - 1 ? a2 - b1 : len - 1) + 1; /* python slice stop */
                if (e1 < 0) { e1 += len; if (e1 < 0) e1 = 0; }
                if (s1 > len) s1 = len;
                entry *x = get_entry(&C, in->hit_taxon[h], i, 1);
                double sc = in->hit_score[h];
                for (int s = s1; s < e1; s++)
                    if (sc > x->site[s]) x->site[s] = sc;
                if (S > 0) {
                    uint32_t m = in->hit_sysmask[h];
                    for (int q = 0; q < S; q++)
                        if ((m >> q & 1) && sc >= ann_score[i * S + q]) { /* OS:389-392 */
                            ann_score[i * S + q] = sc;
                            out->ann_winner[(l0 + C.lraw[i]) * S + q] = (int32_t)h;
                        }
                }
            }
        }
        update_gene_scores(&C, P, tax);
        for (int j = 0; j < P->jump_taxonomy; j++) { raise_taxonomy(&C, P, tax); lifts++; }
        int all_ign = 1;
        for (int i = 0; i < G; i++) all_ign &= C.ign[i];
        if (!all_ign) { /* evaluate_contig, OS:566-583 */
            double *buf = (double *)malloc((size_t)(G + 1) * sizeof(double));
            option *opts = NULL;
            for (int iter = 0;; iter++) {
                /* explain_one, OS:585-597 */
                int nopt = 0, cap = 0;
                for (int t = 0; t < C.T; t++) {
                    double crit, rank;
                    score(&C, t, -1, &crit, &rank, buf);
                    if (crit >= P->k1) {
                        if (nopt == cap) { cap = cap ? 2 * cap : 16; opts = (option *)realloc(opts, (size_t)cap * sizeof(option)); }
                        memset(&opts[nopt], 0, sizeof(option));
                        opts[nopt].t1 = t; opts[nopt].crit = crit; opts[nopt].rank = rank;
                        nopt++;
                    }
                }
                if (nopt > 0) { /* meld_one, OS:621-631: stable sort by rank, best = last */
                    int b = 0;
                    for (int k = 1; k < nopt; k++)
                        if (opts[k].rank >= opts[b].rank) b = k;
                    have_one = 1;
                    best = opts[b];
                    out_b1 = out_c1 = C.clade[best.t1];
                    for (int i = 0; i < G; i++)
                        bsyn[i] = C.ign[i] ? '~' : (C.gene[(size_t)best.t1 * G + i] >= P->k1 ? 'A' : '!');
                    if (P->disambiguate_one == 1) {
                        int l = -1, nm = 0;
                        members = (int *)malloc((size_t)nopt * sizeof(int));
                        for (int k = 0; k < nopt; k++)
                            if (best.rank - opts[k].rank <= P->range) {
                                l = l < 0 ? C.clade[opts[k].t1] : lca2(tax, l, C.clade[opts[k].t1]);
                                members[nm++] = C.clade[opts[k].t1];
                            }
                        out_c1 = l;
                        *nmem_a = nm;
                    }
                    break;
                }
                /* explain_two, OS:599-619 */
                int *pot = (int *)malloc((size_t)(C.T + 1) * sizeof(int));
                int np = 0;
                for (int t = 0; t < C.T; t++) {
                    double mx = C.gene[(size_t)t * G];
                    for (int i = 1; i < G; i++)
                        if (C.gene[(size_t)t * G + i] > mx) mx = C.gene[(size_t)t * G + i];
                    if (mx >= P->k2) pot[np++] = t;
                }
                nopt = 0;
                for (int x = 0; x < np; x++)
                    for (int y = x + 1; y < np; y++) {
                        double crit, rank;
                        score(&C, pot[x], pot[y], &crit, &rank, buf);
                        if (crit >= P->k2) {
                            if (nopt == cap) { cap = cap ? 2 * cap : 16; opts = (option *)realloc(opts, (size_t)cap * sizeof(option)); }
                            memset(&opts[nopt], 0, sizeof(option));
                            opts[nopt].t1 = pot[x]; opts[nopt].t2 = pot[y]; opts[nopt].crit = crit; opts[nopt].rank = rank;
                            opts[nopt].ok = 1;
                            nopt++;
                        }
                    }
                free(pot);
                int resolved = 0;
                if (nopt > 0) { /* meld_two, OS:633-669 */
                    int b = 0;
                    for (int k = 1; k < nopt; k++)
                        if (opts[k].rank >= opts[b].rank) b = k;
                    int nk = 0, all_ok = 1, same = 1, la = -1, lb = -1;
                    int *ma = (int *)malloc((size_t)nopt * sizeof(int)), *mb = (int *)malloc((size_t)nopt * sizeof(int));
                    char *s0 = NULL;
                    opts[b].syn = (char *)calloc((size_t)G + 2, 1);
                    synteny_two(&C, P, tax, &opts[b]);
                    lgt_checks(&C, P, tax, &opts[b]);
                    for (int k = 0; k < nopt; k++) {
                        if (!(opts[b].rank - opts[k].rank <= P->range)) continue;
                        if (k != b) {
                            opts[k].syn = (char *)calloc((size_t)G + 2, 1);
                            synteny_two(&C, P, tax, &opts[k]);
                            lgt_checks(&C, P, tax, &opts[k]);
                        }
                        if (!s0) s0 = opts[k].syn;
                        all_ok &= opts[k].ok;
                        same &= strcmp(opts[k].syn, s0) == 0;
                        la = la < 0 ? opts[k].c1 : lca2(tax, la, opts[k].c1);
                        lb = lb < 0 ? opts[k].c2 : lca2(tax, lb, opts[k].c2);
                        ma[nk] = opts[k].c1;
                        mb[nk] = opts[k].c2;
                        nk++;
                    }
                    int have = 1, melded = 0, c1 = opts[b].c1, c2 = opts[b].c2;
                    if (nk == 1 || P->disambiguate_two == 0) {
                    } else if (P->disambiguate_two == 1) {
                        have = 0;
                    } else if (!all_ok || !same) {
                        have = 0;
                    } else {
                        c1 = la; c2 = lb; melded = 1;
                        if (!P->allow_lca) {
                            int l = lca2(tax, c1, c2);
                            if (l == c1 || l == c2) have = 0;
                        }
                    }
                    if (have && opts[b].ok) {
                        resolved = 1;
                        have_two = 1;
                        best = opts[b];
                        memcpy(bsyn, opts[b].syn, (size_t)G + 1);
                        out_b1 = opts[b].c1; out_b2 = opts[b].c2; out_c1 = c1; out_c2 = c2;
                        if (melded) {
                            qsort(ma, (size_t)nk, sizeof(int), cmp_int);
                            qsort(mb, (size_t)nk, sizeof(int), cmp_int);
                            members = (int *)malloc(2 * (size_t)nk * sizeof(int));
                            int na = 0, nb = 0;
                            for (int k = 0; k < nk; k++)
                                if (k == 0 || ma[k] != ma[k - 1]) members[na++] = ma[k];
                            for (int k = 0; k < nk; k++)
                                if (k == 0 || mb[k] != mb[k - 1]) members[na + nb++] = mb[k];
                            *nmem_a = na;
                            *nmem_b = nb;
                        }
                    }
                    for (int k = 0; k < nopt; k++) free(opts[k].syn);
                    free(ma);
                    free(mb);
                }
                if (resolved) break;
                int hasroot = 0;
                for (int t = 0; t < C.T; t++) hasroot |= C.clade[t] == tax->root;
                if (C.T == 0 || hasroot) break; /* OS:571-574 */
                raise_taxonomy(&C, P, tax);
                lifts++;
            }
            free(opts);
            free(buf);
        }
    }
    /* results */
    out->lifts[c] = lifts;
    out->call[c] = have_one ? WFL_CALL_NO_LGT : (have_two ? WFL_CALL_LGT : WFL_CALL_UNCLASSIFIED);
    out->direction[c] = have_two ? (uint8_t)best.dir : 0;
    out->clade1[c] = out_c1; out->clade2[c] = out_c2; out->best1[c] = out_b1; out->best2[c] = out_b2;
    out->lca[c] = have_two ? lca2(tax, out_c1, out_c2) : -1;
    out->crit[c] = (have_one || have_two) ? best.crit : 0.0;
    out->rank[c] = (have_one || have_two) ? best.rank : 0.0;
    for (int i = 0; i < G; i++) {
        out->locus_flags[l0 + C.lraw[i]] = WFL_LOCUS_RETAINED | (C.ign[i] ? WFL_LOCUS_IGNORED : 0);
        if (have_one || have_two) out->synteny[l0 + C.lraw[i]] = (uint8_t)bsyn[i];
    }
    if (have_one && members) qsort(members, (size_t)*nmem_a, sizeof(int), cmp_int);
    *mem_out = members;
    for (int i = 0; i < C.ne; i++) free(C.e[i].site);
    free(C.e); free(C.llo); free(C.llen); free(C.lraw); free(C.lstr); free(C.ign); free(C.mask);
    free(C.clade); free(C.gene); free(ann_score); free(bsyn);
}

/* contigs are handed out in blocks of 64 from a shared counter */
typedef struct {
    const wfl_params *P; const otax *tax; const wfl_batch *in; wfl_results *out;
    int32_t **mems; int *na, *nb; int64_t n; int64_t next;
} job;
static void *worker(void *arg) {
    job *J = (job *)arg;
    for (;;) {
        int64_t b = __atomic_fetch_add(&J->next, 64, __ATOMIC_RELAXED);
        if (b >= J->n) break;
        int64_t e = b + 64 < J->n ? b + 64 : J->n;
        for (int64_t c = b; c < e; c++) one_contig(J->P, J->tax, J->in, J->out, c, &J->mems[c], &J->na[c], &J->nb[c]);
    }
    return NULL;
}

/* Exported: same arrays in and out as wfl_score_batch (host side). Returns 0, or -4 if `members` is too small
 * (members_used then holds the size needed). */
int wfl_oracle_score_batch(const wfl_params *P, int32_t n_nodes, const int32_t *parent, const int32_t *depth,
                           const int32_t *leaf_count, const uint8_t *listed, int32_t root_idx, int32_t unknown_idx,
                           const wfl_batch *in, wfl_results *out) {
    otax tax = {n_nodes, root_idx, unknown_idx, parent, depth, leaf_count, listed};
    const int64_t n = in->n_contigs;
    int32_t **mems = (int32_t **)calloc((size_t)(n > 0 ? n : 1), sizeof(int32_t *));
    int *na = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int)), *nb = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
    {
        job J = {P, &tax, in, out, mems, na, nb, n, 0};
        int nt = 0;
        const char *e = getenv("WFL_ORACLE_THREADS");
        if (e) nt = atoi(e);
        if (nt <= 0) nt = (int)sysconf(_SC_NPROCESSORS_ONLN);
        if (nt > 256) nt = 256;
        if ((int64_t)nt > (n + 63) / 64) nt = (int)((n + 63) / 64);
        if (nt <= 1) worker(&J);
        else {
            pthread_t th[256];
            int started = 0;
            for (int i = 0; i < nt; i++) if (pthread_create(&th[started], NULL, worker, &J) == 0) started++;
            if (!started) worker(&J);
            for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
        }
    }
    int64_t off = 0;
    int rc = 0;
    for (int64_t c = 0; c < n; c++) {
        out->member_off[c] = off;
        out->n_members_a[c] = na[c];
        int tot = na[c] + nb[c];
        if (off + tot <= out->members_capacity && mems[c]) memcpy(out->members + off, mems[c], (size_t)tot * sizeof(int32_t));
        off += tot;
        free(mems[c]);
    }
    out->member_off[n] = off;
    out->members_used = off;
    if (off > out->members_capacity) rc = -4;
    int64_t cnt[3] = {0, 0, 0};
    for (int64_t c = 0; c < n; c++) cnt[out->call[c] == WFL_CALL_LGT ? 0 : (out->call[c] == WFL_CALL_NO_LGT ? 1 : 2)]++;
    int64_t pos[3] = {0, cnt[0], cnt[0] + cnt[1]};
    for (int64_t c = 0; c < n; c++) {
        int k = out->call[c] == WFL_CALL_LGT ? 0 : (out->call[c] == WFL_CALL_NO_LGT ? 1 : 2);
        out->call_index[pos[k]++] = c;
    }
    out->call_counts[0] = cnt[0]; out->call_counts[1] = cnt[1]; out->call_counts[2] = cnt[2];
    free(mems); free(na); free(nb);
    return rc;
}
