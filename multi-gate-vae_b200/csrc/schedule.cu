// Device-built level schedule: in/out-edge CSR (ascending original edge id per node), ASAP
// levelisation (Kahn frontier peel in one persistent kernel) and the (level, code)-segmented
// node lists.  Integer work only; results are bit-exact with the reference's top_sort /
// subgraph (utils/dag_utils.py:10-37, 91-105) -- see SURVEY.md Appendix D.
//
// Building blocks written here: exclusive scan (3-phase), stable LSD radix sort of
// (key32, value32) pairs (8-bit digits, warp match-any ranking), degree histograms.
#include "mgv_common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;   // 1024

// ------------------------------------------------------------------------------------ scan
__global__ void scan_tiles_kernel(const uint32_t* in, uint32_t* out, int n, uint32_t* tile_sums) {
    __shared__ uint32_t warp_tot[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int base = blockIdx.x * SCAN_TILE + tid * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t local = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        local += v[i];
    }
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[w] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int i = 0; i < w; ++i) woff += warp_tot[i];
    uint32_t run = woff + incl - local;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    if (tid == SCAN_THREADS - 1) tile_sums[blockIdx.x] = run;
}

// one block: exclusive scan of tile_sums[nt] in place (serial over chunks of blockDim with carry)
__global__ void scan_sums_kernel(uint32_t* __restrict__ tile_sums, int nt) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int c0 = 0; c0 < nt; c0 += blockDim.x) {
        uint32_t x = (c0 + tid < nt) ? tile_sums[c0 + tid] : 0u;
        uint32_t incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[w] = incl;
        __syncthreads();
        uint32_t woff = 0;
        for (int i = 0; i < w; ++i) woff += warp_tot[i];
        uint32_t carry = carry_s;
        if (c0 + tid < nt) tile_sums[c0 + tid] = carry + woff + incl - x;
        __syncthreads();
        if (tid == blockDim.x - 1) carry_s = carry + woff + incl;
        __syncthreads();
    }
}

__global__ void scan_add_kernel(uint32_t* __restrict__ out, int n, const uint32_t* __restrict__ tile_sums) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_sums[i / SCAN_TILE];
}

// out[i] = sum_{j<i} in[i]  (out may alias in).  tmp: >= ceil(n/1024) uint32.
int exclusive_scan(const uint32_t* in, uint32_t* out, int n, uint32_t* tmp, cudaStream_t st) {
    if (n <= 0) return MGV_OK;
    const int nt = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tiles_kernel<<<nt, SCAN_THREADS, 0, st>>>(in, out, n, tmp);
    mgv_count_launches(1);
    if (nt > 1) {
        scan_sums_kernel<<<1, 1024, 0, st>>>(tmp, nt);
        mgv_count_launches(1);
        scan_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(out, n, tmp);
        mgv_count_launches(1);
    }
    return mgv_check_cuda(cudaGetLastError(), "exclusive_scan");
}
inline size_t scan_tmp_count(int64_t n) { return (size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 1; }

// ------------------------------------------------------------------------------------ radix sort
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BINS = 256;
// items per warp: 256 (8 steps of 32) for large inputs; 64 up to a million keys, where 2048-item blocks would leave most
// SMs idle (100 k edges = 49 blocks) and each pass is a latency chain of 8 match/scatter steps per warp
inline int rs_per_warp(int64_t m) { return m <= (1 << 20) ? 64 : 256; }
inline int rs_tile(int64_t m) { return RS_WARPS * rs_per_warp(m); }

__global__ void radix_hist_kernel(const uint32_t* __restrict__ keys, int m, int shift,
                                  uint32_t* __restrict__ hist, int nblk, int RS_TILE) {
    __shared__ uint32_t h[RS_BINS];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * RS_TILE;
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        const int idx = base + i;
        if (idx < m) atomicAdd(&h[(keys[idx] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];     // digit-major
}

__global__ void radix_scatter_kernel(const uint32_t* __restrict__ kin, const uint32_t* __restrict__ vin,
                                     uint32_t* __restrict__ kout, uint32_t* __restrict__ vout, int m, int shift,
                                     const uint32_t* __restrict__ offs, int nblk, int RS_PER_WARP) {
    const int RS_TILE = RS_WARPS * RS_PER_WARP;
    __shared__ uint32_t wcnt[RS_WARPS][RS_BINS];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    const int base = blockIdx.x * RS_TILE + w * RS_PER_WARP;
    for (int s = 0; s < RS_PER_WARP / 32; ++s) {
        const int idx = base + s * 32 + lane;
        if (idx < m) atomicAdd(&wcnt[w][(kin[idx] >> shift) & (RS_BINS - 1)], 1u);
    }
    __syncthreads();
    {   // per digit: global offset of this block, then running prefix over the warps
        uint32_t run = offs[(size_t)tid * nblk + blockIdx.x];
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) {
            uint32_t c = wcnt[ww][tid];
            wcnt[ww][tid] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int s = 0; s < RS_PER_WARP / 32; ++s) {
        const int idx = base + s * 32 + lane;
        const bool valid = idx < m;
        const uint32_t key = valid ? kin[idx] : 0u;
        const uint32_t val = valid ? vin[idx] : 0u;
        const uint32_t digit = (key >> shift) & (RS_BINS - 1);
        const uint32_t tag = valid ? digit : (0x80000000u | (uint32_t)lane);   // invalid lanes: singleton groups
        const unsigned peers = __match_any_sync(0xffffffffu, tag);
        const int leader = __ffs(peers) - 1;
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t pos = 0;
        if (valid && lane == leader) {
            pos = wcnt[w][digit];
            wcnt[w][digit] = pos + __popc(peers);
        }
        pos = __shfl_sync(0xffffffffu, pos, leader);
        if (valid) {
            kout[pos + rank] = key;
            vout[pos + rank] = val;
        }
        __syncwarp();
    }
}

struct SortBufs {
    uint32_t *k0, *v0, *k1, *v1, *hist, *tmp;
};
inline size_t sort_hist_count(int64_t m) { return (size_t)RS_BINS * (size_t)((m + rs_tile(m) - 1) / rs_tile(m)) + 1; }

// Stable sort of (k0, v0)[m] by the low `bits` bits of the key.  Result pointers returned.
int radix_sort_pairs(SortBufs b, int m, int bits, uint32_t** kres, uint32_t** vres, cudaStream_t st) {
    uint32_t *ki = b.k0, *vi = b.v0, *ko = b.k1, *vo = b.v1;
    if (m > 0) {
        const int RS_TILE = rs_tile(m);
        const int nblk = (m + RS_TILE - 1) / RS_TILE;
        for (int shift = 0; shift < bits; shift += 8) {
            radix_hist_kernel<<<nblk, RS_THREADS, 0, st>>>(ki, m, shift, b.hist, nblk, RS_TILE);
            mgv_count_launches(1);
            int rc = exclusive_scan(b.hist, b.hist, RS_BINS * nblk, b.tmp, st);
            if (rc != MGV_OK) return rc;
            radix_scatter_kernel<<<nblk, RS_THREADS, 0, st>>>(ki, vi, ko, vo, m, shift, b.hist, nblk, rs_per_warp(m));
            mgv_count_launches(1);
            uint32_t* t;
            t = ki; ki = ko; ko = t;
            t = vi; vi = vo; vo = t;
        }
    }
    *kres = ki;
    *vres = vi;
    return mgv_check_cuda(cudaGetLastError(), "radix_sort_pairs");
}

inline int bits_for(uint64_t max_value) {
    int b = 1;
    while (b < 32 && (max_value >> b) != 0) ++b;
    return b;
}

// ------------------------------------------------------------------------------------ CSR kernels
// An end point outside [0, n) sets the flag and is read as node 0 by EVERY kernel below (keys, degrees, in_src, out_pack,
// code lookup), so the CSR stays self-consistent and every index the compute kernels later follow is in range; the caller
// raises when it reads the flag (deferred mode) or right away (validating mode).
__device__ __forceinline__ int64_t sane_node(int64_t v, int n) { return (v < 0 || v >= n) ? 0 : v; }

__global__ void edge_keys_kernel(const int64_t* __restrict__ ei, int64_t E, int row, uint32_t* __restrict__ keys,
                                 uint32_t* __restrict__ vals, uint32_t* __restrict__ deg, int n, int* __restrict__ bad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t v = ei[(int64_t)row * E + e];
    if (v < 0 || v >= n) {
        atomicOr(bad, 1);
        v = 0;
    }
    keys[e] = (uint32_t)v;
    vals[e] = (uint32_t)e;
    atomicAdd(&deg[v], 1u);
}

__global__ void fill_in_kernel(const uint32_t* __restrict__ eid_sorted, const int64_t* __restrict__ ei, int64_t E, int n,
                               int32_t* __restrict__ in_src, uint32_t* __restrict__ slot_of_eid) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= E) return;
    const uint32_t e = eid_sorted[s];
    in_src[s] = (int32_t)sane_node(ei[e], n);   // row 0 = src
    slot_of_eid[e] = (uint32_t)s;
}

__global__ void fill_out_kernel(const uint32_t* __restrict__ eid_sorted, const int64_t* __restrict__ ei, int64_t E, int n,
                                const uint32_t* __restrict__ slot_of_eid, const int32_t* __restrict__ code,
                                int32_t* __restrict__ out_pack, int32_t* __restrict__ out_slot) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= E) return;
    const uint32_t e = eid_sorted[p];
    const int32_t d = (int32_t)sane_node(ei[E + e], n);   // row 1 = dst
    int32_t c = 0;
    if (code != nullptr) {
        c = code[d];
        if (c < 0 || c > 6) c = 6;
    }
    out_pack[p] = d | (c << MGV_CODE_SHIFT);
    out_slot[p] = (int32_t)slot_of_eid[e];
}

// ------------------------------------------------------------------------------------ levelize
struct LevelizeParams {
    const int32_t* in_ptr;
    const int32_t* out_ptr;
    const int32_t* out_pack;
    int32_t n;
    int32_t* level;
    int32_t* indeg;
    int32_t* frontier[2];
    unsigned* counters;     // [0..2] frontier sizes (rotating), [3] grid barrier, [4] levels, [5] processed
};

__global__ void __launch_bounds__(256) levelize_kernel(LevelizeParams p) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gthreads = gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int gwarp = gtid >> 5, gwarps = gthreads >> 5;
    unsigned* bar = p.counters + 3;
    for (int v = gtid; v < p.n; v += gthreads) {
        const int d = p.in_ptr[v + 1] - p.in_ptr[v];
        p.indeg[v] = d;
        p.level[v] = 0;
        if (d == 0) {
            const unsigned slot = atomicAdd(&p.counters[0], 1u);
            p.frontier[0][slot] = v;
        }
    }
    mgv_grid_sync(bar, gridDim.x);
    unsigned processed = 0;
    int lvl = 0;
    while (true) {
        const unsigned ncur = mgv_ld_acquire(&p.counters[lvl % 3]);
        if (ncur == 0) break;
        processed += ncur;
        if (gtid == 0) p.counters[(lvl + 2) % 3] = 0;
        const int32_t* cur = p.frontier[lvl & 1];
        int32_t* nxt = p.frontier[(lvl + 1) & 1];
        unsigned* nxt_count = &p.counters[(lvl + 1) % 3];
        for (unsigned i = gwarp; i < ncur; i += gwarps) {
            const int v = cur[i];
            const int b = p.out_ptr[v], e = p.out_ptr[v + 1];
            for (int q = b + lane; q < e; q += 32) {
                const int d = p.out_pack[q] & ((1 << MGV_CODE_SHIFT) - 1);
                const int old = atomicSub(&p.indeg[d], 1);
                if (old == 1) {
                    p.level[d] = lvl + 1;
                    const unsigned slot = atomicAdd(nxt_count, 1u);
                    nxt[slot] = d;
                }
            }
        }
        ++lvl;
        mgv_grid_sync(bar, gridDim.x);
    }
    if (gtid == 0) {
        p.counters[4] = (unsigned)lvl;
        p.counters[5] = processed;
    }
}

// ------------------------------------------------------------------------------------ level lists
__global__ void node_keys_kernel(const int32_t* __restrict__ level, const int32_t* __restrict__ code,
                                 const int32_t* __restrict__ stream_of_node, int streams, int n, int L,
                                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ hist,
                                 int* __restrict__ bad) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int lv = level[v];
    if (lv < 0 || lv >= L) {
        atomicOr(bad, 2);
        lv = 0;
    }
    int c = code[v];
    if (c < 0 || c > 6) c = 6;
    int sn = stream_of_node ? stream_of_node[v] : 0;
    if (sn < 0 || sn >= streams) {
        atomicOr(bad, 2);
        sn = 0;
    }
    const uint32_t k = ((uint32_t)sn * (uint32_t)L + (uint32_t)lv) * MGV_NCODE + (uint32_t)c;
    keys[v] = k;
    vals[v] = (uint32_t)v;
    atomicAdd(&hist[k], 1u);
}

__global__ void copy_u32_to_i32_kernel(const uint32_t* __restrict__ src, int32_t* __restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (int32_t)src[i];
}

__global__ void code_count_kernel(const int32_t* __restrict__ seg_ptr, int L, int streams, unsigned long long* __restrict__ out) {
    const int c = threadIdx.x;
    if (c >= MGV_NCODE) return;
    unsigned long long s = 0;
    for (int l = 0; l < streams * L; ++l)
        if (l % L) s += (unsigned long long)(seg_ptr[l * MGV_NCODE + c + 1] - seg_ptr[l * MGV_NCODE + c]);
    out[c] = s;
}

// ------------------------------------------------------------------------------------ sweep row descriptors
__global__ void sweep_desc_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ in_ptr, const int32_t* __restrict__ in_src,
                                  const int32_t* __restrict__ out_ptr, int n, int4* __restrict__ desc) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int v = order[t];
    const int ib = in_ptr[v], ic = in_ptr[v + 1] - ib, ob = out_ptr[v], oc = out_ptr[v + 1] - ob;
    const int s0 = ic > 0 ? in_src[ib] : 0, s1 = ic > 1 ? in_src[ib + 1] : 0, s2 = ic > 2 ? in_src[ib + 2] : 0;
    desc[2 * (size_t)t] = make_int4(v, ib, ic, ob);
    desc[2 * (size_t)t + 1] = make_int4(oc, s0, s1, s2);
}

// ------------------------------------------------------------------------------------ degree order
// key = 255 - min(degree, 255): one stable 8-bit radix pass sorts nodes by DESCENDING degree, ascending id inside
// a degree (so the large equal-degree runs keep their memory locality).
__global__ void degree_keys_kernel(const int32_t* __restrict__ ptr, int n, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int d = ptr[v + 1] - ptr[v];
    keys[v] = 255u - (uint32_t)(d < 255 ? d : 255);
    vals[v] = (uint32_t)v;
}
// gdesc[r] = {node, first CSR slot, degree, first neighbour id} of row r of the degree order: the gather warps
// of the tile kernels start their row loads after ONE dependent load instead of order -> ptr -> idx.
__global__ void gather_desc_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const int32_t* __restrict__ order,
                                   int n, int4* __restrict__ gdesc) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int v = order[r];
    const int b = ptr[v], c = ptr[v + 1] - b;
    gdesc[r] = make_int4(v, b, c, c > 0 ? (idx[b] & ((1 << MGV_CODE_SHIFT) - 1)) : 0);
}
// cost[t] = MGV_TILE_FIXED_COST + rows + neighbours + MGV_TILE_WAVE_COST * (largest degree) of the 128-row tile t of the degree
// order (one warp per tile): a gather lane walks its rows' neighbours in dependent load waves, so a tile's time grows with
// its LARGEST degree, not only with its neighbour count (measured: 450 cycles per wave, 9 500 for a degree <= 2 tile)
__global__ void tile_cost_kernel(const int32_t* __restrict__ ptr, const int32_t* __restrict__ order, int n, int ntiles,
                                 uint32_t* __restrict__ cost) {
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (t > ntiles) return;
    uint32_t c = 0;
    if (t < ntiles) {
        uint32_t dmax = 0;
        for (int r = t * MGV_TILE_ROWS + lane; r < n && r < (t + 1) * MGV_TILE_ROWS; r += 32) {
            const int v = order[r];
            const uint32_t d = (uint32_t)(ptr[v + 1] - ptr[v]);
            c += 1u + d;
            dmax = d > dmax ? d : dmax;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c += __shfl_xor_sync(0xffffffffu, c, o);
            const uint32_t m = __shfl_xor_sync(0xffffffffu, dmax, o);
            dmax = m > dmax ? m : dmax;
        }
        c += MGV_TILE_FIXED_COST + MGV_TILE_WAVE_COST * (dmax < 4096u ? dmax : 4096u);
    }
    if (lane == 0) cost[t] = c;                      // cost[ntiles] = 0: the scan turns it into the total
}

}  // namespace

// ======================================================================================= C ABI
extern "C" size_t mgv_degree_order_workspace_bytes(int64_t N) {
    size_t b = 4 * mgv_align_up((size_t)N * 4 + 256, 256);
    b += mgv_align_up(sort_hist_count(N) * 4 + 256, 256);
    const size_t nt = (size_t)(N / MGV_TILE_ROWS) + 2;
    const size_t scan_n = sort_hist_count(N) > nt ? sort_hist_count(N) : nt;
    b += mgv_align_up(scan_tmp_count(scan_n) * 4 + 256, 256);
    return b + 1024;
}

extern "C" int mgv_build_degree_order(const int32_t* ptr, const int32_t* idx, int32_t N, int32_t* order, int32_t* gdesc,
                                      uint32_t* tile_cost, void* ws, size_t ws_bytes, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MGV_REQUIRE(N >= 0 && ptr && idx && order && gdesc && tile_cost, "mgv_build_degree_order: bad argument");
    if (ws_bytes < mgv_degree_order_workspace_bytes(N)) {
        mgv_set_error("mgv_build_degree_order: workspace too small");
        return MGV_ERR_WORKSPACE;
    }
    const int ntiles = (N + MGV_TILE_ROWS - 1) / MGV_TILE_ROWS;
    MgvArena a(ws, ws_bytes);
    SortBufs sb;
    sb.k0 = a.take<uint32_t>(N + 1); sb.v0 = a.take<uint32_t>(N + 1);
    sb.k1 = a.take<uint32_t>(N + 1); sb.v1 = a.take<uint32_t>(N + 1);
    sb.hist = a.take<uint32_t>(sort_hist_count(N));
    const size_t nt = (size_t)ntiles + 2;
    const size_t scan_n = sort_hist_count(N) > nt ? sort_hist_count(N) : nt;
    sb.tmp = a.take<uint32_t>(scan_tmp_count(scan_n));
    if (N > 0) {
        degree_keys_kernel<<<(N + 255) / 256, 256, 0, st>>>(ptr, N, sb.k0, sb.v0);
        mgv_count_launches(1);
    }
    uint32_t *ks, *vs;
    int rc = radix_sort_pairs(sb, N, 8, &ks, &vs, st);
    if (rc != MGV_OK) return rc;
    if (N > 0) {
        copy_u32_to_i32_kernel<<<(N + 255) / 256, 256, 0, st>>>(vs, order, N);
        gather_desc_kernel<<<(N + 255) / 256, 256, 0, st>>>(ptr, idx, order, N, reinterpret_cast<int4*>(gdesc));
        mgv_count_launches(2);
    }
    tile_cost_kernel<<<(ntiles + 1 + 7) / 8, 256, 0, st>>>(ptr, order, N, ntiles, tile_cost);
    mgv_count_launches(1);
    rc = exclusive_scan(tile_cost, tile_cost, ntiles + 1, sb.tmp, st);
    if (rc != MGV_OK) return rc;
    return mgv_check_cuda(cudaGetLastError(), "mgv_build_degree_order");
}

// Tile descriptors and tile costs of a degree order that already exists (built on the host at collate time, data.py): the part of
// mgv_build_degree_order behind the sort.
extern "C" int mgv_build_degree_tiles(const int32_t* ptr, const int32_t* idx, const int32_t* order, int32_t N, int32_t* gdesc,
                                      uint32_t* tile_cost, void* ws, size_t ws_bytes, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MGV_REQUIRE(N >= 0 && ptr && idx && order && gdesc && tile_cost, "mgv_build_degree_tiles: bad argument");
    if (ws_bytes < mgv_degree_order_workspace_bytes(N)) {
        mgv_set_error("mgv_build_degree_tiles: workspace too small");
        return MGV_ERR_WORKSPACE;
    }
    const int ntiles = (N + MGV_TILE_ROWS - 1) / MGV_TILE_ROWS;
    MgvArena a(ws, ws_bytes);
    uint32_t* tmp = a.take<uint32_t>(scan_tmp_count((size_t)ntiles + 2));
    if (N > 0) {
        gather_desc_kernel<<<(N + 255) / 256, 256, 0, st>>>(ptr, idx, order, N, reinterpret_cast<int4*>(gdesc));
        mgv_count_launches(1);
    }
    tile_cost_kernel<<<(ntiles + 1 + 7) / 8, 256, 0, st>>>(ptr, order, N, ntiles, tile_cost);
    mgv_count_launches(1);
    int rc = exclusive_scan(tile_cost, tile_cost, ntiles + 1, tmp, st);
    if (rc != MGV_OK) return rc;
    return mgv_check_cuda(cudaGetLastError(), "mgv_build_degree_tiles");
}

extern "C" size_t mgv_csr_workspace_bytes(int64_t N, int64_t E) {
    size_t b = 0;
    b += 4 * mgv_align_up((size_t)E * 4 + 256, 256);                 // k0 v0 k1 v1
    b += mgv_align_up(sort_hist_count(E) * 4 + 256, 256);            // hist
    b += mgv_align_up(scan_tmp_count(sort_hist_count(E) > (size_t)N + 1 ? sort_hist_count(E) : (size_t)N + 1) * 4 + 256, 256);
    b += mgv_align_up((size_t)E * 4 + 256, 256);                     // slot_of_eid
    b += mgv_align_up((size_t)(N + 1) * 4 + 256, 256);               // degree scratch
    b += 1024;
    return b;
}

extern "C" int mgv_build_csr(const int64_t* edge_index, int64_t E, int32_t N, const int32_t* code,
                             int32_t* in_ptr, int32_t* in_src, int32_t* out_ptr, int32_t* out_pack,
                             int32_t* out_slot, void* ws, size_t ws_bytes, int32_t* err_flag, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MGV_REQUIRE(N >= 0 && E >= 0, "mgv_build_csr: negative size");
    MGV_REQUIRE((int64_t)N < (1ll << MGV_CODE_SHIFT), "mgv_build_csr: N=%d exceeds 2^28 nodes", N);
    MGV_REQUIRE(E < (1ll << 31), "mgv_build_csr: E too large");
    MGV_REQUIRE(in_ptr && out_ptr, "mgv_build_csr: null output");
    if (ws_bytes < mgv_csr_workspace_bytes(N, E)) {
        mgv_set_error("mgv_build_csr: workspace %zu < %zu bytes", ws_bytes, mgv_csr_workspace_bytes(N, E));
        return MGV_ERR_WORKSPACE;
    }
    MgvArena a(ws, ws_bytes);
    SortBufs sb;
    sb.k0 = a.take<uint32_t>(E + 1); sb.v0 = a.take<uint32_t>(E + 1);
    sb.k1 = a.take<uint32_t>(E + 1); sb.v1 = a.take<uint32_t>(E + 1);
    sb.hist = a.take<uint32_t>(sort_hist_count(E));
    size_t scan_n = sort_hist_count(E) > (size_t)N + 1 ? sort_hist_count(E) : (size_t)N + 1;
    sb.tmp = a.take<uint32_t>(scan_tmp_count(scan_n));
    uint32_t* slot_of_eid = a.take<uint32_t>(E + 1);
    uint32_t* deg = a.take<uint32_t>(N + 1);
    int* bad = err_flag ? err_flag : a.take<int>(1);
    if (!err_flag) MGV_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
    const int bits = bits_for(N > 0 ? (uint64_t)(N - 1) : 0);
    const int eb = (int)((E + 255) / 256);
    for (int dir = 0; dir < 2; ++dir) {
        // dir 0: bucket by dst (row 1) -> in-CSR;  dir 1: bucket by src (row 0) -> out-CSR
        int32_t* ptr = dir == 0 ? in_ptr : out_ptr;
        MGV_CUDA(cudaMemsetAsync(deg, 0, (size_t)(N + 1) * 4, st));
        if (E > 0) {
            edge_keys_kernel<<<eb, 256, 0, st>>>(edge_index, E, dir == 0 ? 1 : 0, sb.k0, sb.v0, deg, N, bad);
            mgv_count_launches(1);
        }
        int rc = exclusive_scan(deg, (uint32_t*)ptr, N + 1, sb.tmp, st);
        if (rc != MGV_OK) return rc;
        uint32_t *ks, *vs;
        rc = radix_sort_pairs(sb, (int)E, bits, &ks, &vs, st);
        if (rc != MGV_OK) return rc;
        if (E > 0) {
            if (dir == 0) {
                fill_in_kernel<<<eb, 256, 0, st>>>(vs, edge_index, E, N, in_src, slot_of_eid);
                mgv_count_launches(1);
            }
            else {
                fill_out_kernel<<<eb, 256, 0, st>>>(vs, edge_index, E, N, slot_of_eid, code, out_pack, out_slot);
                mgv_count_launches(1);
            }
        }
    }
    MGV_CUDA(cudaGetLastError());
    if (err_flag) return MGV_OK;                 // deferred validation: the caller reads *err_flag later (no sync here)
    int bad_h = 0;
    MGV_CUDA(cudaMemcpyAsync(&bad_h, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    MGV_CUDA(cudaStreamSynchronize(st));
    MGV_REQUIRE(bad_h == 0, "mgv_build_csr: edge_index holds a node id outside [0, %d)", N);
    return MGV_OK;
}

extern "C" size_t mgv_levelize_workspace_bytes(int64_t N) {
    return 3 * mgv_align_up((size_t)(N + 1) * 4 + 256, 256) + 1024;
}

extern "C" int mgv_levelize(const int32_t* in_ptr, const int32_t* out_ptr, const int32_t* out_pack, int32_t N,
                            int32_t* level, int32_t* info_host, void* ws, size_t ws_bytes, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MGV_REQUIRE(N >= 0 && info_host, "mgv_levelize: bad argument");
    if (ws_bytes < mgv_levelize_workspace_bytes(N)) {
        mgv_set_error("mgv_levelize: workspace too small");
        return MGV_ERR_WORKSPACE;
    }
    if (N == 0) {
        info_host[0] = 0; info_host[1] = 0;
        return MGV_OK;
    }
    MgvArena a(ws, ws_bytes);
    LevelizeParams p;
    p.in_ptr = in_ptr; p.out_ptr = out_ptr; p.out_pack = out_pack; p.n = N; p.level = level;
    p.indeg = a.take<int32_t>(N + 1);
    p.frontier[0] = a.take<int32_t>(N + 1);
    p.frontier[1] = a.take<int32_t>(N + 1);
    p.counters = a.take<unsigned>(8);
    MGV_CUDA(cudaMemsetAsync(p.counters, 0, 8 * sizeof(unsigned), st));
    int dev = 0, sms = 0, occ = 0;
    MGV_CUDA(cudaGetDevice(&dev));
    MGV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MGV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, levelize_kernel, 256, 0));
    if (occ > 4) occ = 4;
    MGV_REQUIRE(occ >= 1, "mgv_levelize: kernel does not fit on an SM");
    int grid = sms * occ;
    const int need = (N + 255) / 256;
    if (grid > need) grid = need > 0 ? need : 1;
    void* args[] = {&p};
    MGV_CUDA(cudaLaunchCooperativeKernel((void*)levelize_kernel, dim3(grid), dim3(256), args, 0, st));
    mgv_count_launches(1);
    unsigned res[2] = {0, 0};
    MGV_CUDA(cudaMemcpyAsync(res, p.counters + 4, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    MGV_CUDA(cudaStreamSynchronize(st));
    info_host[0] = (int32_t)res[0];
    info_host[1] = (int32_t)res[1];
    if ((int32_t)res[1] != N) {
        mgv_set_error("mgv_levelize: graph has a cycle (%u of %d nodes levelised)", res[1], N);
        return MGV_ERR_CYCLE;
    }
    return MGV_OK;
}

extern "C" size_t mgv_level_lists_workspace_bytes(int64_t N, int32_t L) {
    size_t nkeys = (size_t)L * MGV_NCODE + 1;
    size_t b = 4 * mgv_align_up((size_t)N * 4 + 256, 256);
    b += mgv_align_up(sort_hist_count(N) * 4 + 256, 256);
    size_t scan_n = sort_hist_count(N) > nkeys ? sort_hist_count(N) : nkeys;
    b += mgv_align_up(scan_tmp_count(scan_n) * 4 + 256, 256);
    b += mgv_align_up(nkeys * 4 + 256, 256);
    b += 1024;
    return b;
}

extern "C" int mgv_build_level_lists(const int32_t* level, const int32_t* code, const int32_t* stream_of_node, int32_t streams,
                                     int32_t N, int32_t L, int32_t* order, int32_t* seg_ptr, int64_t* code_count_host,
                                     void* ws, size_t ws_bytes, int32_t* err_flag, mgv_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    MGV_REQUIRE(N >= 0 && L >= 1 && level && code && order && seg_ptr, "mgv_build_level_lists: bad argument");
    MGV_REQUIRE(streams >= 1 && (streams == 1 || stream_of_node), "mgv_build_level_lists: streams > 1 needs stream_of_node");
    const int L1 = L;
    L = streams * L1;                                     // keys: (stream, level, code)
    MGV_REQUIRE(code_count_host || err_flag, "mgv_build_level_lists: the asynchronous form (no code_count_host) needs err_flag");
    MGV_REQUIRE((int64_t)L * MGV_NCODE < (1ll << 31), "mgv_build_level_lists: too many levels");
    if (ws_bytes < mgv_level_lists_workspace_bytes(N, L)) {
        mgv_set_error("mgv_build_level_lists: workspace too small");
        return MGV_ERR_WORKSPACE;
    }
    const int nkeys = L * MGV_NCODE;
    MgvArena a(ws, ws_bytes);
    SortBufs sb;
    sb.k0 = a.take<uint32_t>(N + 1); sb.v0 = a.take<uint32_t>(N + 1);
    sb.k1 = a.take<uint32_t>(N + 1); sb.v1 = a.take<uint32_t>(N + 1);
    sb.hist = a.take<uint32_t>(sort_hist_count(N));
    size_t scan_n = sort_hist_count(N) > (size_t)nkeys + 1 ? sort_hist_count(N) : (size_t)nkeys + 1;
    sb.tmp = a.take<uint32_t>(scan_tmp_count(scan_n));
    uint32_t* khist = a.take<uint32_t>(nkeys + 1);
    unsigned long long* counts = a.take<unsigned long long>(MGV_NCODE);
    int* bad = err_flag ? err_flag : a.take<int>(1);
    if (!err_flag) MGV_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
    MGV_CUDA(cudaMemsetAsync(khist, 0, (size_t)(nkeys + 1) * 4, st));
    if (N > 0) {
        node_keys_kernel<<<(N + 255) / 256, 256, 0, st>>>(level, code, stream_of_node, streams, N, L1, sb.k0, sb.v0, khist, bad);
        mgv_count_launches(1);
    }
    int rc = exclusive_scan(khist, (uint32_t*)seg_ptr, nkeys + 1, sb.tmp, st);
    if (rc != MGV_OK) return rc;
    uint32_t *ks, *vs;
    rc = radix_sort_pairs(sb, N, bits_for((uint64_t)(nkeys > 0 ? nkeys - 1 : 0)), &ks, &vs, st);
    if (rc != MGV_OK) return rc;
    if (N > 0) {
        copy_u32_to_i32_kernel<<<(N + 255) / 256, 256, 0, st>>>(vs, order, N);
        mgv_count_launches(1);
    }
    if (!code_count_host) return mgv_check_cuda(cudaGetLastError(), "mgv_build_level_lists");   // asynchronous form
    code_count_kernel<<<1, 32, 0, st>>>(seg_ptr, L1, streams, counts);
    mgv_count_launches(1);
    MGV_CUDA(cudaGetLastError());
    unsigned long long ch[MGV_NCODE];
    int bad_h = 0;
    MGV_CUDA(cudaMemcpyAsync(ch, counts, sizeof(ch), cudaMemcpyDeviceToHost, st));
    MGV_CUDA(cudaMemcpyAsync(&bad_h, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    MGV_CUDA(cudaStreamSynchronize(st));
    MGV_REQUIRE(err_flag || bad_h == 0, "mgv_build_level_lists: level outside [0, %d) or stream outside [0, %d)", L1, streams);
    for (int c = 0; c < MGV_NCODE; ++c) code_count_host[c] = (int64_t)ch[c];
    return MGV_OK;
}

extern "C" int mgv_build_sweep_desc(const int32_t* order, const int32_t* in_ptr, const int32_t* in_src, const int32_t* out_ptr,
                                    int32_t N, int32_t* desc, mgv_stream_t stream) {
    MGV_REQUIRE(N >= 0 && order && in_ptr && in_src && out_ptr && desc, "mgv_build_sweep_desc: bad argument");
    MGV_REQUIRE((reinterpret_cast<uintptr_t>(desc) & 15) == 0, "mgv_build_sweep_desc: desc must be 16-byte aligned");
    if (N == 0) return MGV_OK;
    sweep_desc_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(order, in_ptr, in_src, out_ptr, N, reinterpret_cast<int4*>(desc));
    mgv_count_launches(1);
    return mgv_check_cuda(cudaGetLastError(), "mgv_build_sweep_desc");
}
