// K10: on-device compaction of the per-contig results.
//
// (1) The melded-member lists that the scoring kernel bump-allocated in a staging pool (in
//     whatever order CTAs finished) are rewritten as a CSR in contig order.
// (2) Contig indices are stably partitioned into lgt | no_lgt | unclassified, i.e. the three
//     output tables of write_main_output_files (waafle/waafle_orgscorer.py:843-890) -- this is
//     also the compact record set a multi-GPU run gathers.
// Three launches: block totals, scan of the block totals, scatter.
#include "wfl_device.cuh"

namespace wfl {

namespace {

constexpr int CT = 256;          // threads per block
constexpr int CPB = 1024;        // contigs per block (4 per thread, blocked)

struct Quad {
    long long m, a, b, c;   // members, lgt, no_lgt, unclassified
};

__device__ __forceinline__ Quad quad_of(const DevOut &o, long long i) {
    uint8_t call = o.call[i];
    return Quad{(long long)o.n_mem_a[i] + o.n_mem_b[i], call == WFL_CALL_LGT, call == WFL_CALL_NO_LGT,
                call == WFL_CALL_UNCLASSIFIED};
}
__device__ __forceinline__ Quad operator+(const Quad &x, const Quad &y) {
    return Quad{x.m + y.m, x.a + y.a, x.b + y.b, x.c + y.c};
}
__device__ __forceinline__ Quad shfl_up(const Quad &x, int o) {
    return Quad{__shfl_up_sync(0xffffffffu, x.m, o), __shfl_up_sync(0xffffffffu, x.a, o),
                __shfl_up_sync(0xffffffffu, x.b, o), __shfl_up_sync(0xffffffffu, x.c, o)};
}

// Exclusive scan of one Quad per thread over the block; total returned to all threads.
__device__ Quad block_scan(Quad v, Quad &total) {
    __shared__ Quad ws[CT / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    Quad inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Quad t = shfl_up(inc, o);
        if (lane >= o) inc = inc + t;
    }
    if (lane == 31) ws[w] = inc;
    __syncthreads();
    Quad base{0, 0, 0, 0};
    total = base;
    for (int i = 0; i < CT / 32; ++i) {
        if (i < w) base = base + ws[i];
        total = total + ws[i];
    }
    __syncthreads();
    return Quad{base.m + inc.m - v.m, base.a + inc.a - v.a, base.b + inc.b - v.b, base.c + inc.c - v.c};
}

__global__ void wfl_compact_totals(CompactArgs a) {
    long long first = (long long)blockIdx.x * CPB + (long long)threadIdx.x * (CPB / CT);
    Quad v{0, 0, 0, 0};
    for (int k = 0; k < CPB / CT; ++k)
        if (first + k < a.n) v = v + quad_of(a.o, first + k);
    Quad tot;
    block_scan(v, tot);
    if (threadIdx.x == 0) {
        a.scan_tmp[4 * blockIdx.x + 0] = tot.m;
        a.scan_tmp[4 * blockIdx.x + 1] = tot.a;
        a.scan_tmp[4 * blockIdx.x + 2] = tot.b;
        a.scan_tmp[4 * blockIdx.x + 3] = tot.c;
    }
}

__global__ void wfl_compact_scan_blocks(CompactArgs a, int n_blocks) {
    // single block: chunked exclusive scan over the block totals, in place
    Quad carry{0, 0, 0, 0};
    for (int base = 0; base < n_blocks; base += CT) {
        int i = base + threadIdx.x;
        Quad v{0, 0, 0, 0};
        if (i < n_blocks)
            v = Quad{a.scan_tmp[4 * i], a.scan_tmp[4 * i + 1], a.scan_tmp[4 * i + 2], a.scan_tmp[4 * i + 3]};
        Quad tot, ex = block_scan(v, tot);
        if (i < n_blocks) {
            a.scan_tmp[4 * i + 0] = carry.m + ex.m;
            a.scan_tmp[4 * i + 1] = carry.a + ex.a;
            a.scan_tmp[4 * i + 2] = carry.b + ex.b;
            a.scan_tmp[4 * i + 3] = carry.c + ex.c;
        }
        carry = carry + tot;
    }
    if (threadIdx.x == 0) {
        a.totals[0] = carry.m;
        a.totals[1] = carry.a;
        a.totals[2] = carry.b;
        a.totals[3] = carry.c;
        a.call_counts[0] = carry.a;
        a.call_counts[1] = carry.b;
        a.call_counts[2] = carry.c;
        a.member_off[a.n] = carry.m;
    }
}

__global__ void wfl_compact_scatter(CompactArgs a) {
    long long first = (long long)blockIdx.x * CPB + (long long)threadIdx.x * (CPB / CT);
    Quad q[CPB / CT], v{0, 0, 0, 0};
    for (int k = 0; k < CPB / CT; ++k) {
        q[k] = first + k < a.n ? quad_of(a.o, first + k) : Quad{0, 0, 0, 0};
        v = v + q[k];
    }
    Quad tot, ex = block_scan(v, tot);
    Quad off{a.scan_tmp[4 * blockIdx.x] + ex.m, a.scan_tmp[4 * blockIdx.x + 1] + ex.a,
             a.scan_tmp[4 * blockIdx.x + 2] + ex.b, a.scan_tmp[4 * blockIdx.x + 3] + ex.c};
    const long long n_lgt = a.totals[1], n_no = a.totals[2];
    for (int k = 0; k < CPB / CT; ++k) {
        long long i = first + k;
        if (i >= a.n) break;
        a.member_off[i] = off.m;
        a.n_members_a[i] = a.o.n_mem_a[i];
        long long src = a.o.mem_pos[i];
        for (long long m = 0; m < q[k].m; ++m)
            if (off.m + m < a.members_cap && src + m < a.o.mem_pool_cap)
                a.members[off.m + m] = a.o.mem_pool[src + m];
        long long dst = q[k].a ? off.a : q[k].b ? n_lgt + off.b : n_lgt + n_no + off.c;
        a.call_index[dst] = i;
        off = off + q[k];
    }
}

}  // namespace

size_t compaction_scratch_elems(int64_t n) { return 4 * (size_t)((n + CPB - 1) / CPB + 1); }

int launch_compaction(const CompactArgs &a, cudaStream_t s) {
    int nb = (int)((a.n + CPB - 1) / CPB);
    if (nb == 0) nb = 1;
    wfl_compact_totals<<<nb, CT, 0, s>>>(a);
    wfl_compact_scan_blocks<<<1, CT, 0, s>>>(a, nb);
    wfl_compact_scatter<<<nb, CT, 0, s>>>(a);
    return 3;
}

}  // namespace wfl
