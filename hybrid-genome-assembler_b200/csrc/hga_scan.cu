// scan_probe: fused 2-bit pack + canonical k-mer windows + membership test + hit compaction.
//
// Replaces the per-read half of ReadClusteringEngine::construct_indices
// (clustering/ReadClusteringEngine.cpp:246-277) and KmerIterator (common/KmerIterator.cpp:23-76).
//
// Work decomposition (B200: 148 SMs, persistent CTAs pulling tile tickets):
//   * the concatenated base stream is cut into tiles of SCAN_TILE window-end positions;
//   * a CTA loads the tile's ASCII bytes (+32 bases of halo) with coalesced 16 B streaming loads and packs them
//     into TWO shared-memory streams: forward codes MSB-first and complement codes LSB-first. Bytes outside
//     {A,C,G,T} are stored as code 0 in BOTH streams, which is exactly the reference's rule
//     (KmerIterator.cpp:56,62: unordered_map::operator[] default-inserts 0 in both tables);
//   * every window is EXTRACTED from the two streams with funnel shifts (no rolling, no reverse-complement
//     arithmetic): the 32 lanes of a warp own 32 consecutive windows, nothing depends on the previous window;
//   * membership = one 8 B probe of the L2-resident blocked Bloom filter per window; windows that pass are
//     queued in shared memory and the 32 B key-group probes run DENSE, 32 candidates per warp instruction;
//   * hits are staged per warp in position order; a tile's hits are appended with ONE atomicAdd on a global
//     cursor (no inter-CTA waiting), and a copy kernel later moves the tile segments into global position
//     order (exclusive scan over tile counts). CSR row offsets are produced per tile and fixed up with the
//     tile's final offset: no per-hit atomics anywhere.
#include "hga_internal.cuh"

#include <cub/device/device_scan.cuh>

#define SCAN_THREADS 256
#define SCAN_WARPS (SCAN_THREADS / 32)
#define SCAN_TILE 4096
#define SCAN_SPAN (SCAN_TILE / SCAN_WARPS)      // windows per warp per tile (512)
#define SCAN_STEPS (SCAN_SPAN / 32)             // 16
#define SCAN_HALO 32
#define SCAN_CHUNKS ((SCAN_TILE + SCAN_HALO) / 16)   // 258 chunks of 16 bases
#define SCAN_UNROLL 2
#define SCAN_QCAP (32 * SCAN_UNROLL + 32)

struct ScanScalars {
    unsigned long long ticket;
    unsigned long long cursor;       // hits appended so far (= total when the kernel ends)
    unsigned int overflow;
    unsigned int pad;
};

namespace {

struct ScanParams {
    const char *bases;
    uint64_t n_bases;
    const uint64_t *read_off;
    uint64_t n_reads;
    int k;
    const uint64_t *keys;
    const uint64_t *filter;
    uint32_t n_groups, n_words;
    uint32_t *tmp_slot, *tmp_pos;     // tile segments in completion order
    uint64_t capacity;
    uint64_t *row_off;                // tile-local hit count before the read's first base (fixed up later)
    unsigned long long *tile_tmp_off; // where the tile's segment starts in tmp_*
    unsigned long long *tile_cnt;     // hits of the tile
    ScanScalars *scalars;
    uint64_t n_tiles;
    uint64_t tile_stride;             // > 0: sampling mode (count only, tile = ticket * stride)
};

// 4 ASCII bytes (little-endian in w, lowest address = first base) -> forward codes (8 bits, first base most
// significant) and complement codes (8 bits, first base LEAST significant). Codes follow KmerIterator.cpp:7-19
// (A0 C1 G2 T3 / complement A3 C2 G1 T0); any other byte gives 0 in both.
__device__ __forceinline__ void pack4(uint32_t w, uint32_t &fwd8, uint32_t &rc8) {
    const uint32_t valid = __vcmpeq4(w, 0x41414141u) | __vcmpeq4(w, 0x43434343u) | __vcmpeq4(w, 0x47474747u) | __vcmpeq4(w, 0x54545454u);
    uint32_t x = (w >> 1) & 0x03030303u;          // A0 C1 G3 T2
    x ^= (x >> 1) & 0x01010101u;                  // A0 C1 G2 T3
    const uint32_t xc = (x ^ 0x03030303u) & valid;
    x &= valid;
    fwd8 = (x * 0x40100401u) >> 24;               // b0<<6 | b1<<4 | b2<<2 | b3
    rc8 = (xc * 0x01041040u) >> 24;               // c0 | c1<<2 | c2<<4 | c3<<6
}

__device__ __forceinline__ void pack16(uint4 v, uint32_t &fwd_word, uint32_t &rc_word) {
    uint32_t f0, f1, f2, f3, r0, r1, r2, r3;
    pack4(v.x, f0, r0); pack4(v.y, f1, r1); pack4(v.z, f2, r2); pack4(v.w, f3, r3);
    fwd_word = (f0 << 24) | (f1 << 16) | (f2 << 8) | f3;
    rc_word = r0 | (r1 << 8) | (r2 << 16) | (r3 << 24);
}

// last r in [lo, hi] with read_off[r] <= pos (read_off[lo] <= pos is guaranteed by the caller)
__device__ __forceinline__ uint64_t find_read(const uint64_t *__restrict__ read_off, uint64_t lo, uint64_t hi, uint64_t pos) {
    while (lo < hi) {
        uint64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(&read_off[mid]) <= pos) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ int clamp_local(uint64_t glob, uint64_t tile_start) {
    const int64_t d = (int64_t) glob - (int64_t) tile_start;
    return (int) max((int64_t) -(1 << 30), min((int64_t) (1 << 30), d));
}

struct WarpStage {
    unsigned long long q_key[SCAN_QCAP];
    uint32_t q_pos[SCAN_QCAP];
    uint32_t st_slot[SCAN_SPAN];
    uint32_t st_pos[SCAN_SPAN];
    uint16_t q_w[SCAN_QCAP];
    uint16_t st_w[SCAN_SPAN];
};

// probe the key table for the first n (<= 32) queue entries, append the found ones to the staging area
__device__ __forceinline__ uint32_t drain_queue(WarpStage &ws, uint32_t n, uint32_t st_count, const ScanParams &p, int lane) {
    uint32_t slot = 0xFFFFFFFFu, pos = 0, widx = 0;
    if ((uint32_t) lane < n) {
        const unsigned long long key = ws.q_key[lane];
        pos = ws.q_pos[lane]; widx = ws.q_w[lane];
        uint32_t g = hga_scale(hga_hash(key).hi, p.n_groups);
        for (;;) {
            const ulonglong2 *gp = reinterpret_cast<const ulonglong2 *>(p.keys + 4ull * g);
            const ulonglong2 a = __ldg(gp), b = __ldg(gp + 1);
            if (a.x == key) { slot = 4 * g; break; }
            if (a.y == key) { slot = 4 * g + 1; break; }
            if (b.x == key) { slot = 4 * g + 2; break; }
            if (b.y == key) { slot = 4 * g + 3; break; }
            if (a.x == HGA_EMPTY_KEY || a.y == HGA_EMPTY_KEY || b.x == HGA_EMPTY_KEY || b.y == HGA_EMPTY_KEY) break;
            g = (g + 1 == p.n_groups) ? 0 : g + 1;
        }
    }
    const bool hit = slot != 0xFFFFFFFFu;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
    if (hit) {
        const uint32_t idx = st_count + __popc(bal & ((1u << lane) - 1));
        ws.st_slot[idx] = slot; ws.st_pos[idx] = pos; ws.st_w[idx] = (uint16_t) widx;
    }
    return st_count + __popc(bal);
}

struct ScanSmem {
    WarpStage stage[SCAN_WARPS];
    uint32_t fwd[SCAN_CHUNKS + 2];
    uint32_t rc[SCAN_CHUNKS + 2];
    uint32_t wcount[SCAN_WARPS + 1];
    unsigned long long tile, tile_off;
    uint64_t r_lo, r_hi, r_first;
};

__global__ void __launch_bounds__(SCAN_THREADS, 4) scan_probe_kernel(ScanParams p) {
    extern __shared__ __align__(16) unsigned char scan_smem_raw[];
    ScanSmem &S = *reinterpret_cast<ScanSmem *>(scan_smem_raw);
    uint32_t *s_fwd = S.fwd, *s_rc = S.rc, *s_wcount = S.wcount;
    WarpStage *s_stage = S.stage;
    unsigned long long &s_tile = S.tile, &s_tile_off = S.tile_off;
    uint64_t &s_r_lo = S.r_lo, &s_r_hi = S.r_hi, &s_r_first = S.r_first;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = p.k;
    const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    const uint32_t lane_lt = (1u << lane) - 1;
    WarpStage &ws = s_stage[warp];

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(&p.scalars->ticket, 1ull);
        __syncthreads();
        const uint64_t ticket = s_tile;
        if (ticket >= p.n_tiles) break;
        const uint64_t tile = p.tile_stride ? ticket * p.tile_stride : ticket;
        const uint64_t tile_start = tile * SCAN_TILE;
        const int n_loc = (int) min((uint64_t) SCAN_TILE, p.n_bases - tile_start);      // window ends in this tile

        // ---- load + pack (chunk c covers global bases [tile_start - 32 + 16c, +16)) -------------------
        for (int c = tid; c < SCAN_CHUNKS; c += SCAN_THREADS) {
            const int64_t g = (int64_t) tile_start - SCAN_HALO + 16 * (int64_t) c;
            uint4 v = make_uint4(0, 0, 0, 0);      // byte 0 is not a base: packs to 0/0 and never reaches a valid window
            if (g >= 0 && (uint64_t) g + 16 <= p.n_bases) {
                v = __ldcs(reinterpret_cast<const uint4 *>(p.bases + g));
            } else if (g >= 0 && (uint64_t) g < p.n_bases) {
                uint32_t t[4] = {0, 0, 0, 0};
                for (int i = 0; i < 16 && (uint64_t) g + i < p.n_bases; i++) t[i >> 2] |= (uint32_t) (unsigned char) p.bases[g + i] << (8 * (i & 3));
                v = make_uint4(t[0], t[1], t[2], t[3]);
            }
            uint32_t fw, rw;
            pack16(v, fw, rw);
            s_fwd[c] = fw; s_rc[c] = rw;
        }
        if (tid < 2) { s_fwd[SCAN_CHUNKS + tid] = 0; s_rc[SCAN_CHUNKS + tid] = 0; }
        if (tid == 0) s_r_lo = find_read(p.read_off, 0, p.n_reads - 1, tile_start);
        if (tid == 32) s_r_hi = find_read(p.read_off, 0, p.n_reads - 1, tile_start + n_loc - 1);
        if (tid == 64 && !p.tile_stride) {
            uint64_t lo = 0, hi = p.n_reads;     // first read whose offset is >= tile_start
            while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (__ldg(&p.read_off[mid]) < tile_start) lo = mid + 1; else hi = mid; }
            s_r_first = lo;
        }
        __syncthreads();

        // ---- windows: lane owns window end e_loc = warp*512 + step*32 + lane ---------------------------------
        const uint64_t r_hi = s_r_hi;
        const int e_first = warp * SCAN_SPAN + lane;
        uint64_t r = s_r_lo;
        int cur_start = 0, cur_end = 1 << 30;
        if (e_first < n_loc) {
            r = find_read(p.read_off, r, r_hi, tile_start + e_first);
            cur_start = clamp_local(__ldg(&p.read_off[r]), tile_start);
            cur_end = clamp_local(__ldg(&p.read_off[r + 1]), tile_start);
        }
        uint32_t q_count = 0, st_count = 0;   // warp uniform

        for (int s0 = 0; s0 < SCAN_STEPS; s0 += SCAN_UNROLL) {
            unsigned long long canon[SCAN_UNROLL];
            uint32_t posv[SCAN_UNROLL], m0[SCAN_UNROLL], m1[SCAN_UNROLL];
            uint2 fw[SCAN_UNROLL];
            bool valid[SCAN_UNROLL];
            #pragma unroll
            for (int u = 0; u < SCAN_UNROLL; u++) {
                const int e = e_first + 32 * (s0 + u);
                if (e < n_loc) {
                    while (e >= cur_end) {
                        r++;
                        cur_start = cur_end;
                        cur_end = clamp_local(__ldg(&p.read_off[r + 1]), tile_start);
                    }
                }
                valid[u] = (e < n_loc) && (e - cur_start + 1 >= k);
                posv[u] = (uint32_t) (e - cur_start + 1);   // exact: reads are < 2^30 bases (checked by hga_scan)
                // window [js, je] in staged coordinates (0 = tile_start - 32)
                const int je = e + SCAN_HALO, js = je - k + 1;
                const int w0 = js >> 4, o = (js & 15) * 2;
                const uint32_t F0 = s_fwd[w0], F1 = s_fwd[w0 + 1], F2 = s_fwd[w0 + 2];
                const uint32_t R0 = s_rc[w0], R1 = s_rc[w0 + 1], R2 = s_rc[w0 + 2];
                const unsigned long long fwd = ((((unsigned long long) __funnelshift_l(F1, F0, o)) << 32) | __funnelshift_l(F2, F1, o)) >> (64 - 2 * k);
                const unsigned long long rc = ((((unsigned long long) __funnelshift_r(R1, R2, o)) << 32) | __funnelshift_r(R0, R1, o)) & kmask;
                canon[u] = fwd < rc ? fwd : rc;                       // KmerIterator.cpp:69
                const KmerHash hs = hga_hash(canon[u]);
                m0[u] = (1u << (hs.lo >> 27)) | __funnelshift_l(0u, 1u, hs.lo >> 22);
                m1[u] = __funnelshift_l(0u, 1u, hs.lo >> 17) | __funnelshift_l(0u, 1u, hs.lo >> 12);
                fw[u] = make_uint2(0u, 0u);
                if (valid[u]) fw[u] = __ldg(reinterpret_cast<const uint2 *>(p.filter + hga_scale(hs.hi, p.n_words)));
            }
            #pragma unroll
            for (int u = 0; u < SCAN_UNROLL; u++) {
                const bool pass = valid[u] && ((fw[u].x & m0[u]) == m0[u]) && ((fw[u].y & m1[u]) == m1[u]);
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pass);
                if (pass) {
                    const uint32_t idx = q_count + __popc(bal & lane_lt);
                    ws.q_key[idx] = canon[u]; ws.q_pos[idx] = posv[u]; ws.q_w[idx] = (uint16_t) (e_first + 32 * (s0 + u));
                }
                q_count += __popc(bal);
            }
            __syncwarp();
            while (q_count >= 32) {
                st_count = drain_queue(ws, 32, st_count, p, lane);
                q_count -= 32;
                __syncwarp();
                // move the (< 32 * SCAN_UNROLL) leftovers to the front, 32 at a time
                for (uint32_t base = 0; base < q_count; base += 32) {
                    unsigned long long kk = 0; uint32_t pp = 0; uint16_t ww = 0;
                    const bool act = base + lane < q_count;
                    if (act) { kk = ws.q_key[32 + base + lane]; pp = ws.q_pos[32 + base + lane]; ww = ws.q_w[32 + base + lane]; }
                    __syncwarp();
                    if (act) { ws.q_key[base + lane] = kk; ws.q_pos[base + lane] = pp; ws.q_w[base + lane] = ww; }
                    __syncwarp();
                }
            }
        }
        if (q_count) st_count = drain_queue(ws, q_count, st_count, p, lane);
        if (lane == 0) s_wcount[warp] = st_count;
        __syncthreads();

        // ---- append the tile's hits (one atomic), tile directory, tile-local CSR row offsets -------------------
        uint32_t wbase = 0, total = 0;
        #pragma unroll
        for (int i = 0; i < SCAN_WARPS; i++) { if (i < warp) wbase += s_wcount[i]; total += s_wcount[i]; }
        if (p.tile_stride) {
            if (tid == 0 && total) atomicAdd(&p.scalars->cursor, (unsigned long long) total);
        } else {
            if (tid == 0) {
                const unsigned long long off = atomicAdd(&p.scalars->cursor, (unsigned long long) total);
                s_tile_off = off;
                p.tile_tmp_off[tile] = off; p.tile_cnt[tile] = total;
                if (off + total > p.capacity) p.scalars->overflow = 1;
            }
            // rows starting in this tile: number of tile hits whose window ends before the read's first base
            for (uint64_t rr = s_r_first + tid; rr <= r_hi; rr += SCAN_THREADS) {
                const uint64_t ro = __ldg(&p.read_off[rr]);
                if (ro >= tile_start && ro < tile_start + (uint64_t) n_loc) {
                    const uint32_t q = (uint32_t) (ro - tile_start);
                    const uint32_t wq = q / SCAN_SPAN;
                    uint32_t before = 0;
                    for (uint32_t i = 0; i < wq; i++) before += s_wcount[i];
                    const uint16_t *sw = s_stage[wq].st_w;
                    uint32_t lo = 0, hi = s_wcount[wq];
                    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (sw[mid] < q) lo = mid + 1; else hi = mid; }
                    p.row_off[rr] = before + lo;
                }
            }
            __syncthreads();
            const unsigned long long tile_off = s_tile_off;
            if (tile_off + total <= p.capacity) {
                for (uint32_t i = lane; i < st_count; i += 32) {
                    p.tmp_slot[tile_off + wbase + i] = ws.st_slot[i];
                    p.tmp_pos[tile_off + wbase + i] = ws.st_pos[i];
                }
            }
        }
        __syncthreads();
    }
}

// tile segments (completion order) -> global position order; one warp per tile
__global__ void scan_reorder_kernel(const uint32_t *__restrict__ tmp_slot, const uint32_t *__restrict__ tmp_pos,
                                    const unsigned long long *__restrict__ tile_tmp_off, const unsigned long long *__restrict__ tile_cnt,
                                    const unsigned long long *__restrict__ tile_off, uint64_t n_tiles, uint32_t *out_slot, uint32_t *out_pos) {
    const uint64_t warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    const uint64_t w = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t t = w; t < n_tiles; t += warps) {
        const unsigned long long src = tile_tmp_off[t], dst = tile_off[t], n = tile_cnt[t];
        for (unsigned long long i = lane; i < n; i += 32) { out_slot[dst + i] = tmp_slot[src + i]; out_pos[dst + i] = tmp_pos[src + i]; }
    }
}

// row_off[r] (tile-local count) += final offset of the tile holding the read's first base; rows at or past the end
// of the stream (trailing empty reads, terminal entry) get E
__global__ void scan_fix_rows_kernel(const uint64_t *__restrict__ read_off, uint64_t n_reads, uint64_t n_bases,
                                     const unsigned long long *__restrict__ tile_off, uint64_t n_tiles, uint64_t *row_off) {
    const unsigned long long E = tile_off[n_tiles];
    uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (; i <= n_reads; i += stride) {
        const uint64_t ro = (i == n_reads) ? n_bases : read_off[i];
        row_off[i] = (ro >= n_bases) ? E : row_off[i] + tile_off[ro / SCAN_TILE];
    }
}

}  // namespace

int hga_scan_run(hga_handle *h, const char *d_bases, const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases) {
    h->have_scan = h->have_index = h->have_pairs = h->have_selection = h->have_components = false;
    h->n_reads = n_reads; h->n_bases = n_bases; h->n_hits = 0;
    HGA_TRY(h->d_row_off.ensure((n_reads + 1) * 8));
    HGA_TRY(h->d_scan_scalars.ensure(sizeof(ScanScalars)));
    ScanScalars *d_sc = h->d_scan_scalars.as<ScanScalars>();
    const uint64_t n_tiles = (n_reads == 0) ? 0 : (n_bases + SCAN_TILE - 1) / SCAN_TILE;
    // tile directory: tmp offset | count | final offset  (n_tiles + 1 each)
    HGA_TRY(h->d_tile_state.ensure((n_tiles + 1) * 8 * 3));
    unsigned long long *tile_tmp_off = h->d_tile_state.as<unsigned long long>();
    unsigned long long *tile_cnt = tile_tmp_off + (n_tiles + 1), *tile_off = tile_cnt + (n_tiles + 1);

    ScanParams p;
    p.bases = d_bases; p.n_bases = n_bases; p.read_off = d_read_off; p.n_reads = n_reads; p.k = h->k;
    p.keys = h->table.keys; p.filter = h->table.filter; p.n_groups = h->table.n_groups; p.n_words = h->table.n_words;
    p.row_off = h->d_row_off.as<uint64_t>();
    p.tile_tmp_off = tile_tmp_off; p.tile_cnt = tile_cnt;
    p.scalars = d_sc;

    int occ = 0;
    const size_t smem = sizeof(ScanSmem);
    HGA_CUDA(cudaFuncSetAttribute(scan_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    HGA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, scan_probe_kernel, SCAN_THREADS, smem));
    if (occ < 1) occ = 1;
    const int grid_full = (int) std::min<uint64_t>((uint64_t) h->sm_count * occ, std::max<uint64_t>(n_tiles, 1));

    StageTimer timer(h, &h->metrics.scan_ms);
    ScanScalars sc;
    memset(&sc, 0, sizeof(sc));
    uint64_t capacity = 0;
    if (n_tiles > 0) {
        // capacity of the temporary hit arrays: exact upper bound for small inputs, otherwise estimated from a
        // strided count-only sample (1 tile in 64)
        const uint64_t small_limit = 32ull << 20;
        if (n_bases <= small_limit) {
            capacity = n_bases;
        } else {
            const uint64_t stride = 64;
            const uint64_t n_sample = (n_tiles + stride - 1) / stride;
            HGA_CUDA(cudaMemsetAsync(d_sc, 0, sizeof(ScanScalars), h->stream));
            ScanParams ps = p;
            ps.tile_stride = stride; ps.n_tiles = n_sample; ps.capacity = 0; ps.tmp_slot = ps.tmp_pos = nullptr;
            const int grid_s = (int) std::min<uint64_t>((uint64_t) h->sm_count * occ, n_sample);
            scan_probe_kernel<<<grid_s, SCAN_THREADS, smem, h->stream>>>(ps);
            h->metrics.kernel_launches++;
            HGA_CUDA(cudaGetLastError());
            HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            const double est = (double) sc.cursor * (double) n_tiles / (double) n_sample;
            capacity = (uint64_t) (est * 1.10) + (1ull << 20);
            if (capacity > n_bases) capacity = n_bases;
        }
    }

    uint64_t E = 0;
    for (int attempt = 0; attempt < 2 && n_tiles > 0; attempt++) {
        HGA_TRY(h->d_sort_a.ensure((capacity + 1) * 4));      // temporaries (reused by the index sort later)
        HGA_TRY(h->d_sort_b.ensure((capacity + 1) * 4));
        p.tmp_slot = h->d_sort_a.as<uint32_t>(); p.tmp_pos = h->d_sort_b.as<uint32_t>();
        p.capacity = capacity; p.n_tiles = n_tiles; p.tile_stride = 0;
        HGA_CUDA(cudaMemsetAsync(d_sc, 0, sizeof(ScanScalars), h->stream));
        scan_probe_kernel<<<grid_full, SCAN_THREADS, smem, h->stream>>>(p);
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
        HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        E = sc.cursor;
        if (!sc.overflow) break;
        if (attempt == 1) { hga_set_error("scan: hit buffer overflow after exact resize (internal error)"); return HGA_E_OVERFLOW; }
        capacity = E;   // exact; rerun once
    }

    HGA_TRY(h->d_hit_slot.ensure((E + 1) * 4));
    HGA_TRY(h->d_hit_pos.ensure((E + 1) * 4));
    if (n_tiles > 0) {
        size_t tmp_bytes = 0;
        HGA_CUDA(cudaMemsetAsync(tile_cnt + n_tiles, 0, 8, h->stream));
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, tile_cnt, tile_off, n_tiles + 1, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp_bytes + 16));
        HGA_CUDA(cub::DeviceScan::ExclusiveSum(h->d_sort_tmp.p, tmp_bytes, tile_cnt, tile_off, n_tiles + 1, h->stream));
        const int blocks = (int) std::min<uint64_t>((n_tiles * 32 + 255) / 256, (uint64_t) h->sm_count * 16);
        scan_reorder_kernel<<<blocks, 256, 0, h->stream>>>(h->d_sort_a.as<uint32_t>(), h->d_sort_b.as<uint32_t>(), tile_tmp_off, tile_cnt, tile_off, n_tiles,
                                                         h->d_hit_slot.as<uint32_t>(), h->d_hit_pos.as<uint32_t>());
        const int rblocks = (int) std::min<uint64_t>((n_reads + 256) / 256, 2048);
        scan_fix_rows_kernel<<<rblocks, 256, 0, h->stream>>>(d_read_off, n_reads, n_bases, tile_off, n_tiles, p.row_off);
        h->metrics.kernel_launches += 4;
        HGA_CUDA(cudaGetLastError());
    } else {
        HGA_CUDA(cudaMemsetAsync(p.row_off, 0, (n_reads + 1) * 8, h->stream));
    }
    timer.stop();
    h->n_hits = E;
    h->metrics.n_bases = n_bases; h->metrics.n_reads = n_reads; h->metrics.n_hits = E;
    h->have_scan = true;
    return HGA_OK;
}
