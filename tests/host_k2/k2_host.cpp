// Host build of the K2 arithmetic of waafle_b200/csrc/wfl_warp_common.cuh (region between the markers
// "K2-HOST-BEGIN" / "K2-HOST-END", extracted by tests/test_k2_host.py) against a LITERAL evaluation:
// materialise the per-site array of a (clade, locus) group (waafle_orgscorer.py:371-382) and sum it
// in numpy's pairwise order.  Runs without a GPU; exit code 0 = every case bit-exact.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define __device__
#define __noinline__
#define __forceinline__ inline
typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;
using std::max;
using std::min;
constexpr int MAXDEPTH = 28;
constexpr int RMAX = 6;
#define WFL_K2_TWO 1

#include "k2_region.inc"

static double pw(const double *a, long n) {
    if (n < 8) { double r = 0.; for (long i = 0; i < n; i++) r += a[i]; return r; }
    if (n <= 128) {
        double r[8]; long i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    long n2 = n / 2; n2 -= n2 % 8;
    return pw(a, n2) + pw(a + n2, n - n2);
}

int main(int argc, char **argv) {
    const int cases = argc > 1 ? atoi(argv[1]) : 20000;
    std::mt19937_64 rng(12345);
    long bad = 0, done = 0;
    for (int it = 0; it < cases; ++it) {
        int n;
        switch (it % 6) {
            case 0: n = 1 + rng() % 40; break;
            case 1: n = 100 + rng() % 60; break;
            case 2: n = 200 + rng() % 2900; break;
            case 3: n = 128 * (1 + rng() % 20) + (int)(rng() % 3) - 1; break;
            default: n = 300 + rng() % 2200; break;
        }
        int k = (it % 5 == 0) ? 1 + rng() % 40 : 1 + rng() % 4;
        std::vector<int> ra(k), rb(k);
        std::vector<double> rv(k);
        const int style = rng() % 4;
        for (int i = 0; i < k; ++i) {
            int a, b;
            if (style == 0) {            // near-full-length hits (the cfg2 shape)
                a = (int)(rng() % 61); b = n - (int)(rng() % 61);
            } else if (style == 1) {     // arbitrary intervals
                a = rng() % n; b = a + 1 + rng() % (n - a);
            } else if (style == 2) {     // short islands
                a = rng() % n; b = std::min(n, a + 1 + (int)(rng() % 20));
            } else {                     // boundaries at multiples of 8 and around leaf edges
                a = (int)((rng() % (n / 8 + 1)) * 8) + (int)(rng() % 3) - 1; b = a + 1 + (int)(rng() % (n));
            }
            a = std::max(0, std::min(a, n - 1)); b = std::max(a + 1, std::min(b, n));
            if (rng() % 16 == 0) b = a;  // empty python slice
            ra[i] = a; rb[i] = b;
            double v = (rng() % 8 == 0) ? (double)(1 + rng() % 8) / 8.0 : (double)(rng() % 1000000) / 1e6 * 1.04;
            if (rng() % 32 == 0) v = 0.0;
            rv[i] = v;
        }
        // descending score order (stable), as the kernels keep the records
        std::vector<int> ord(k);
        for (int i = 0; i < k; ++i) ord[i] = i;
        std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return rv[x] > rv[y]; });
        std::vector<int> sa(k), sb(k); std::vector<double> sv(k);
        for (int i = 0; i < k; ++i) { sa[i] = ra[ord[i]]; sb[i] = rb[ord[i]]; sv[i] = rv[ord[i]]; }
        std::vector<double> site(n, 0.0);
        for (int i = 0; i < k; ++i) for (int p = sa[i]; p < sb[i]; ++p) site[p] = std::max(site[p], sv[i]);
        const double want = pw(site.data(), n) / (double)n;
        std::vector<u16> plan(plan_cap(n) + 4);
        u8 k8set[8] = {0};
        const int nleaf = build_plan(n, plan.data(), k8set);
        const u32 k8pack = k8set[0] | (k8set[1] << 8) | (k8set[2] << 16) | ((u32)k8set[3] << 24);
        for (int sorted = 1; sorted >= 0; --sorted) {
            const double got = sorted ? group_mean(sa.data(), sb.data(), sv.data(), 0, k, n, true, k8pack, plan.data(), nleaf)
                                      : group_mean(ra.data(), rb.data(), rv.data(), 0, k, n, false, k8pack, plan.data(), nleaf);
            ++done;
            if (memcmp(&got, &want, 8) != 0) {
                if (++bad <= 10) fprintf(stderr, "MISMATCH case %d n=%d k=%d sorted=%d got %a want %a\n", it, n, k, sorted, got, want);
            }
        }
    }
    // the fast path's closed-form integral (group_integral: records in descending score order, connected union, generic
    // endpoint sweep when a record leaves a gap) against the literal site array, within 1e-13 relative
    long fbad = 0, fdone = 0;
    for (int it = 0; it < cases; ++it) {
        const int n = 50 + rng() % 3000;
        const int k = (it % 7 == 0) ? 2 + rng() % 60 : 2 + rng() % 5;
        std::vector<double> bv(k);
        std::vector<u32> bab(k);
        std::vector<u16> bord(k);
        const int style = rng() % 4;
        for (int i = 0; i < k; ++i) {
            int a, b;
            if (style == 0) { a = (int)(rng() % 61); b = n - (int)(rng() % 61); }
            else if (style == 1) { a = rng() % n; b = a + 1 + rng() % (n - a); }
            else if (style == 2) { a = rng() % n; b = std::min(n, a + 1 + (int)(rng() % 40)); }
            else { a = (int)(rng() % (n / 2)); b = a + 1 + (int)(rng() % (n / 2)); }
            a = std::max(0, std::min(a, n - 1)); b = std::max(a + 1, std::min(b, n));
            bab[i] = (u32)a | ((u32)b << 16);
            double v = (rng() % 8 == 0) ? (double)(1 + rng() % 8) / 8.0 : (double)(rng() % 1000000) / 1e6 * 1.04;
            if (rng() % 32 == 0) v = 0.0;
            bv[i] = v;
            bord[i] = (u16)i;
        }
        // descending score order, ties in buffer order: what the fast path's per-locus sort delivers
        std::stable_sort(bord.begin(), bord.end(), [&](u16 x, u16 y) { return bv[x] > bv[y]; });
        std::vector<double> site(n, 0.0);
        for (int i = 0; i < k; ++i)
            for (int p = (int)(bab[i] & 0xffffu); p < (int)(bab[i] >> 16); ++p) site[p] = std::max(site[p], bv[i]);
        const double want = pw(site.data(), n);
        for (int general = 0; general < 2; ++general) {
            const double got = general ? group_integral_general(bv.data(), bab.data(), bord.data(), 0, k)
                                       : group_integral(bv.data(), bab.data(), bord.data(), 0, k, n);
            ++fdone;
            if (std::fabs(got - want) > 1e-13 * std::max(1.0, std::fabs(want))) {
                if (++fbad <= 10) fprintf(stderr, "FAST MISMATCH case %d n=%d k=%d general=%d got %a want %a\n", it, n, k, general, got, want);
            }
        }
    }
    printf("%ld fast-path evaluations, %ld out of tolerance\n", fdone, fbad);
    bad += fbad;
    printf("%ld evaluations, %ld mismatches\n", done, bad);
    return bad ? 1 : 0;
}
