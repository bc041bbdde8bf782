// scan_probe: fused 2-bit pack + canonical k-mer windows + membership test + ordered hit compaction.
//
// Replaces the per-read half of ReadClusteringEngine::construct_indices
// (clustering/ReadClusteringEngine.cpp:246-277) and KmerIterator (common/KmerIterator.cpp:23-76).
//
// Work decomposition (B200: 148 SMs, persistent CTAs, tickets in tile order):
//   * the concatenated base stream is cut into tiles of SCAN_TILE window-end positions;
//   * a CTA loads the tile's ASCII bytes (+32 bases of halo) with coalesced 16 B loads, packs them to
//     2 bit/base (+1 exception bit for bytes outside {A,C,G,T}) into shared memory;
//   * every window is EXTRACTED from the packed words (funnel shifts) instead of rolled, so the 32 lanes of a
//     warp own 32 consecutive windows: ballots come out in position order and nothing depends on the
//     previous window (ILP for the probes);
//   * membership = blocked-Bloom word (L2 resident) then the 32 B key group;
//   * hits are staged in shared memory per warp, the tile total goes through a decoupled look-back
//     (single 64-bit status word per tile) and the tile's hits are written once, coalesced, in global
//     position order. Row offsets (CSR by read) fall out of the per-window hit bitmap: no atomics.
#include "hga_internal.cuh"

#define SCAN_THREADS 256
#define SCAN_WARPS (SCAN_THREADS / 32)
#define SCAN_TILE 4096
#define SCAN_SPAN (SCAN_TILE / SCAN_WARPS)      // windows per warp per tile (512)
#define SCAN_STEPS (SCAN_SPAN / 32)             // 16
#define SCAN_HALO 32
#define SCAN_CHUNKS ((SCAN_TILE + SCAN_HALO) / 16)   // 258 chunks of 16 bases
#define SCAN_UNROLL 4

#define TS_FLAG_AGG (1ull << 62)
#define TS_FLAG_PREFIX (2ull << 62)
#define TS_VALUE_MASK ((1ull << 62) - 1)

struct ScanScalars {
    unsigned long long ticket;
    unsigned long long total_hits;
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
    uint32_t *hit_slot, *hit_pos;
    uint64_t capacity;
    uint64_t *row_off;
    unsigned long long *tile_state;
    ScanScalars *scalars;
    uint64_t n_tiles;
    uint64_t tile_stride;   // > 0: sampling mode (count only, tile = ticket * stride)
};

// 4 ASCII bytes (little-endian in w, lowest address = first base) -> 8 bits of 2-bit codes, first base most
// significant, and 4 exception bits (bit 3 = first base). Codes follow KmerIterator.cpp:7-12 (A0 C1 G2 T3);
// any other byte gets code 0 and its exception bit set (both strands see 0, KmerIterator.cpp:56,62).
__device__ __forceinline__ void pack4(uint32_t w, uint32_t &codes8, uint32_t &exc4) {
    uint32_t valid = __vcmpeq4(w, 0x41414141u) | __vcmpeq4(w, 0x43434343u) | __vcmpeq4(w, 0x47474747u) | __vcmpeq4(w, 0x54545454u);
    uint32_t x = (w >> 1) & 0x03030303u;          // A0 C1 G3 T2
    x ^= (x >> 1) & 0x01010101u;                  // A0 C1 G2 T3
    x &= valid;
    codes8 = (x * 0x40100401u) >> 24;
    exc4 = (((~valid) & 0x01010101u) * 0x08040201u) >> 24 & 0xFu;
}

__device__ __forceinline__ void pack16(uint4 v, uint32_t &word, uint32_t &exc16) {
    uint32_t c0, c1, c2, c3, e0, e1, e2, e3;
    pack4(v.x, c0, e0); pack4(v.y, c1, e1); pack4(v.z, c2, e2); pack4(v.w, c3, e3);
    word = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    exc16 = (e0 << 12) | (e1 << 8) | (e2 << 4) | e3;
}

// reverse complement of a right-aligned 2k-bit k-mer (no exceptions)
__device__ __forceinline__ uint64_t revcomp(uint64_t fwd, int k) {
    uint64_t y = __brevll(~fwd);
    y = ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
    return y >> (64 - 2 * k);
}

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// last r in [lo, hi] with read_off[r] <= pos (read_off[lo] <= pos is guaranteed by the caller)
__device__ __forceinline__ uint64_t find_read(const uint64_t *__restrict__ read_off, uint64_t lo, uint64_t hi, uint64_t pos) {
    while (lo < hi) {
        uint64_t mid = (lo + hi + 1) >> 1;
        if (__ldg(&read_off[mid]) <= pos) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(SCAN_THREADS, 4) scan_probe_kernel(ScanParams p) {
    __shared__ uint32_t s_pack[SCAN_CHUNKS + 2];
    __shared__ uint16_t s_exc16[SCAN_CHUNKS + 2 + 4];
    __shared__ uint32_t s_hit_slot[SCAN_WARPS][SCAN_SPAN];
    __shared__ uint32_t s_hit_pos[SCAN_WARPS][SCAN_SPAN];
    __shared__ uint32_t s_hitmask[SCAN_TILE / 32];
    __shared__ uint32_t s_wordprefix[SCAN_TILE / 32 + 1];
    __shared__ unsigned long long s_tile;
    __shared__ unsigned long long s_tile_off;
    __shared__ uint64_t s_r_lo, s_r_hi, s_r_first;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = p.k;
    const uint32_t *s_exc32 = reinterpret_cast<const uint32_t *>(s_exc16);

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(&p.scalars->ticket, 1ull);
        __syncthreads();
        const uint64_t ticket = s_tile;
        if (ticket >= p.n_tiles) break;
        const uint64_t tile = p.tile_stride ? ticket * p.tile_stride : ticket;
        const uint64_t tile_start = tile * SCAN_TILE;
        const uint64_t tile_end = min(tile_start + (uint64_t) SCAN_TILE, p.n_bases);

        // ---- load + pack (chunk c covers global bases [tile_start - 32 + 16c, +16)) -------------------
        int any_exc = 0;
        for (int c = tid; c < SCAN_CHUNKS; c += SCAN_THREADS) {
            int64_t g = (int64_t) tile_start - SCAN_HALO + 16 * (int64_t) c;
            uint4 v = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);   // 'A' filler never reaches a valid window
            if (g >= 0 && (uint64_t) g + 16 <= p.n_bases) {
                v = __ldcs(reinterpret_cast<const uint4 *>(p.bases + g));
            } else if (g >= 0 && (uint64_t) g < p.n_bases) {
                unsigned char tmp[16];
                #pragma unroll
                for (int i = 0; i < 16; i++) tmp[i] = ((uint64_t) g + i < p.n_bases) ? (unsigned char) p.bases[g + i] : (unsigned char) 'A';
                v.x = tmp[0] | (tmp[1] << 8) | (tmp[2] << 16) | ((uint32_t) tmp[3] << 24);
                v.y = tmp[4] | (tmp[5] << 8) | (tmp[6] << 16) | ((uint32_t) tmp[7] << 24);
                v.z = tmp[8] | (tmp[9] << 8) | (tmp[10] << 16) | ((uint32_t) tmp[11] << 24);
                v.w = tmp[12] | (tmp[13] << 8) | (tmp[14] << 16) | ((uint32_t) tmp[15] << 24);
            }
            uint32_t word, exc;
            pack16(v, word, exc);
            s_pack[c] = word;
            s_exc16[c ^ 1] = (uint16_t) exc;      // u32 view: even chunk in the high half (MSB first)
            any_exc |= (exc != 0);
        }
        if (tid < 2) { s_pack[SCAN_CHUNKS + tid] = 0; s_exc16[(SCAN_CHUNKS + tid) ^ 1] = 0; }
        if (tid == 0) s_r_lo = find_read(p.read_off, 0, p.n_reads - 1, tile_start);
        if (tid == 32) s_r_hi = find_read(p.read_off, 0, p.n_reads - 1, tile_end - 1);
        any_exc = __syncthreads_or(any_exc);

        // ---- windows ------------------------------------------------------------------------------------
        const uint64_t r_lo = s_r_lo, r_hi = s_r_hi;
        const uint64_t e0 = tile_start + (uint64_t) warp * SCAN_SPAN + lane;
        uint64_t r = r_lo, cur_start = 0, cur_end = 0;
        if (e0 < p.n_bases) {
            r = find_read(p.read_off, r_lo, r_hi, e0);
            cur_start = __ldg(&p.read_off[r]);
            cur_end = __ldg(&p.read_off[r + 1]);
        }
        uint32_t wc = 0;   // hits staged by this warp so far (warp uniform)

        for (int s0 = 0; s0 < SCAN_STEPS; s0 += SCAN_UNROLL) {
            uint64_t canon[SCAN_UNROLL];
            uint32_t posv[SCAN_UNROLL];
            bool valid[SCAN_UNROLL], pass[SCAN_UNROLL];
            uint32_t hhi[SCAN_UNROLL];
            uint2 fw[SCAN_UNROLL];
            uint32_t m0[SCAN_UNROLL], m1[SCAN_UNROLL];

            #pragma unroll
            for (int u = 0; u < SCAN_UNROLL; u++) {
                const int s = s0 + u;
                const uint64_t e = e0 + 32 * (uint64_t) s;
                valid[u] = false;
                canon[u] = 0; posv[u] = 0; hhi[u] = 0; m0[u] = m1[u] = 0;
                if (e < p.n_bases) {
                    while (e >= cur_end) { r++; cur_start = cur_end; cur_end = __ldg(&p.read_off[r + 1]); }
                    valid[u] = (e + 1 >= cur_start + (uint64_t) k);
                    posv[u] = (uint32_t) (e - cur_start + 1);
                }
                // window [js, je] in staged coordinates (0 = tile_start - 32)
                const int je = warp * SCAN_SPAN + s * 32 + lane + SCAN_HALO;
                const int js = je - k + 1;
                const int w0 = js >> 4, o = (js & 15) * 2;
                const uint32_t W0 = s_pack[w0], W1 = s_pack[w0 + 1], W2 = s_pack[w0 + 2];
                const uint32_t top = __funnelshift_l(W1, W0, o), bot = __funnelshift_l(W2, W1, o);
                const uint64_t fwd = (((uint64_t) top << 32) | bot) >> (64 - 2 * k);
                uint64_t rc = revcomp(fwd, k);
                if (any_exc) {
                    // exception bits of the window, bit (k-1-i) <-> i-th base; such bases read 0 on both strands
                    const int x0 = js >> 5, xo = js & 31;
                    uint32_t ex = __funnelshift_l(s_exc32[x0 + 1], s_exc32[x0], xo) >> (32 - k);
                    while (ex) {
                        int b = 31 - __clz(ex);          // bit (k-1-i)
                        int i = k - 1 - b;
                        rc &= ~(3ull << (2 * i));
                        ex &= ~(1u << b);
                    }
                }
                canon[u] = fwd < rc ? fwd : rc;
                KmerHash hs = hga_hash(canon[u]);
                hhi[u] = hs.hi;
                hga_filter_mask(hs.lo, m0[u], m1[u]);
                fw[u] = make_uint2(0, 0);
                if (valid[u]) fw[u] = __ldg(reinterpret_cast<const uint2 *>(p.filter + hga_scale(hs.hi, p.n_words)));
            }
            #pragma unroll
            for (int u = 0; u < SCAN_UNROLL; u++)
                pass[u] = valid[u] && ((fw[u].x & m0[u]) == m0[u]) && ((fw[u].y & m1[u]) == m1[u]);

            uint32_t slot[SCAN_UNROLL];
            ulonglong2 ga[SCAN_UNROLL], gb[SCAN_UNROLL];
            uint32_t grp[SCAN_UNROLL];
            #pragma unroll
            for (int u = 0; u < SCAN_UNROLL; u++) {
                slot[u] = 0xFFFFFFFFu;
                grp[u] = hga_scale(hhi[u], p.n_groups);
                if (pass[u]) {
                    const ulonglong2 *gp = reinterpret_cast<const ulonglong2 *>(p.keys + 4ull * grp[u]);
                    ga[u] = __ldg(gp); gb[u] = __ldg(gp + 1);
                }
            }
            #pragma unroll
            for (int u = 0; u < SCAN_UNROLL; u++) {
                if (pass[u]) {
                    uint32_t g = grp[u];
                    ulonglong2 a = ga[u], b = gb[u];
                    const uint64_t key = canon[u];
                    for (;;) {
                        if (a.x == key) { slot[u] = 4 * g; break; }
                        if (a.y == key) { slot[u] = 4 * g + 1; break; }
                        if (b.x == key) { slot[u] = 4 * g + 2; break; }
                        if (b.y == key) { slot[u] = 4 * g + 3; break; }
                        if (a.x == HGA_EMPTY_KEY || a.y == HGA_EMPTY_KEY || b.x == HGA_EMPTY_KEY || b.y == HGA_EMPTY_KEY) break;
                        g = (g + 1 == p.n_groups) ? 0 : g + 1;
                        const ulonglong2 *gp = reinterpret_cast<const ulonglong2 *>(p.keys + 4ull * g);
                        a = __ldg(gp); b = __ldg(gp + 1);
                    }
                }
            }
            #pragma unroll
            for (int u = 0; u < SCAN_UNROLL; u++) {
                const bool hit = slot[u] != 0xFFFFFFFFu;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
                if (hit) {
                    const uint32_t idx = wc + __popc(bal & ((1u << lane) - 1));
                    s_hit_slot[warp][idx] = slot[u];
                    s_hit_pos[warp][idx] = posv[u];
                }
                if (lane == 0) s_hitmask[warp * SCAN_STEPS + s0 + u] = bal;
                wc += __popc(bal);
            }
        }
        __syncthreads();

        // ---- tile prefix over the hit bitmap + decoupled look-back (warp 0) ----------------------------
        if (warp == 0) {
            constexpr int WPL = (SCAN_TILE / 32) / 32;   // bitmap words per lane (4)
            uint32_t cnt[WPL], sum = 0;
            #pragma unroll
            for (int i = 0; i < WPL; i++) { cnt[i] = __popc(s_hitmask[lane * WPL + i]); sum += cnt[i]; }
            uint32_t incl = sum;
            #pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += t; }
            uint32_t run = incl - sum;
            #pragma unroll
            for (int i = 0; i < WPL; i++) { s_wordprefix[lane * WPL + i] = run; run += cnt[i]; }
            const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            if (lane == 31) s_wordprefix[SCAN_TILE / 32] = total;

            if (p.tile_stride) {
                if (lane == 0) atomicAdd(&p.scalars->total_hits, (unsigned long long) total);
            } else {
                unsigned long long exclusive = 0;
                if (tile == 0) {
                    if (lane == 0) atomicExch(&p.tile_state[0], TS_FLAG_PREFIX | (unsigned long long) total);
                } else {
                    if (lane == 0) atomicExch(&p.tile_state[tile], TS_FLAG_AGG | (unsigned long long) total);
                    int64_t look = (int64_t) tile - 1;
                    for (;;) {
                        const int64_t idx = look - lane;
                        unsigned long long st = TS_FLAG_PREFIX;   // tiles before 0 contribute an exclusive prefix of 0
                        if (idx >= 0) {
                            st = ld_state(&p.tile_state[idx]);
                            while ((st >> 62) == 0) { __nanosleep(20); st = ld_state(&p.tile_state[idx]); }
                        }
                        const uint32_t pm = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2);
                        unsigned long long v = st & TS_VALUE_MASK;
                        if (pm) {
                            const int first = __ffs(pm) - 1;
                            if (lane > first) v = 0;
                        }
                        #pragma unroll
                        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
                        exclusive += v;
                        if (pm) break;
                        look -= 32;
                    }
                    if (lane == 0) atomicExch(&p.tile_state[tile], TS_FLAG_PREFIX | (exclusive + total));
                }
                if (lane == 0) {
                    s_tile_off = exclusive;
                    if (tile == p.n_tiles - 1) p.scalars->total_hits = exclusive + total;
                    if (exclusive + total > p.capacity) p.scalars->overflow = 1;
                }
            }
        } else if (warp == 1 && !p.tile_stride) {
            // first read whose offset is >= tile_start (rows starting in this tile get their offset below)
            if (lane == 0) {
                uint64_t lo = 0, hi = p.n_reads;     // lower_bound over read_off[0..n_reads]
                while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (__ldg(&p.read_off[mid]) < tile_start) lo = mid + 1; else hi = mid; }
                s_r_first = lo;
            }
        }
        __syncthreads();

        if (!p.tile_stride) {
            const uint64_t tile_off = s_tile_off;
            const uint32_t tile_total = s_wordprefix[SCAN_TILE / 32];
            if (tile_off + tile_total <= p.capacity) {
                const uint32_t wbase = s_wordprefix[warp * SCAN_STEPS];
                for (uint32_t i = lane; i < wc; i += 32) {
                    p.hit_slot[tile_off + wbase + i] = s_hit_slot[warp][i];
                    p.hit_pos[tile_off + wbase + i] = s_hit_pos[warp][i];
                }
            }
            // CSR row offsets of the reads that start inside this tile: hits before local position q
            for (uint64_t rr = s_r_first + tid; rr <= r_hi; rr += SCAN_THREADS) {
                const uint64_t ro = __ldg(&p.read_off[rr]);
                if (ro >= tile_start && ro < tile_end) {
                    const uint32_t q = (uint32_t) (ro - tile_start);
                    const uint32_t wi = q >> 5, bi = q & 31;
                    p.row_off[rr] = tile_off + s_wordprefix[wi] + __popc(s_hitmask[wi] & ((1u << bi) - 1));
                }
            }
        }
        __syncthreads();
    }
}

// rows whose offset equals n_bases (trailing empty reads) and the terminal entry
__global__ void scan_finalize_rows_kernel(const uint64_t *__restrict__ read_off, uint64_t n_reads, uint64_t n_bases, uint64_t *row_off,
                                          const ScanScalars *scalars) {
    const uint64_t E = scalars->total_hits;
    uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (; i <= n_reads; i += stride)
        if (i == n_reads || read_off[i] >= n_bases) row_off[i] = E;
}

}  // namespace

int hga_scan_run(hga_handle *h, const char *d_bases, const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases) {
    h->have_scan = h->have_index = h->have_pairs = h->have_selection = h->have_components = false;
    h->n_reads = n_reads; h->n_bases = n_bases; h->n_hits = 0;
    HGA_TRY(h->d_row_off.ensure((n_reads + 1) * 8));
    HGA_TRY(h->d_scan_scalars.ensure(sizeof(ScanScalars)));
    ScanScalars *d_sc = h->d_scan_scalars.as<ScanScalars>();
    const uint64_t n_tiles = (n_bases + SCAN_TILE - 1) / SCAN_TILE;
    HGA_TRY(h->d_tile_state.ensure((n_tiles + 1) * 8));

    ScanParams p;
    p.bases = d_bases; p.n_bases = n_bases; p.read_off = d_read_off; p.n_reads = n_reads; p.k = h->k;
    p.keys = h->table.keys; p.filter = h->table.filter; p.n_groups = h->table.n_groups; p.n_words = h->table.n_words;
    p.row_off = h->d_row_off.as<uint64_t>();
    p.tile_state = h->d_tile_state.as<unsigned long long>();
    p.scalars = d_sc;

    int occ = 0;
    HGA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, scan_probe_kernel, SCAN_THREADS, 0));
    if (occ < 1) occ = 1;
    const int grid_full = (int) std::min<uint64_t>((uint64_t) h->sm_count * occ, std::max<uint64_t>(n_tiles, 1));

    StageTimer timer(h, &h->metrics.scan_ms);
    ScanScalars sc;
    uint64_t capacity = 0;
    if (n_tiles > 0 && n_reads > 0) {
        // capacity: exact upper bound for small inputs, otherwise estimated from a strided count-only sample
        const uint64_t small_limit = 32ull << 20;
        if (n_bases <= small_limit) {
            capacity = n_bases;
        } else {
            const uint64_t stride = 64;
            const uint64_t n_sample = (n_tiles + stride - 1) / stride;
            HGA_CUDA(cudaMemsetAsync(d_sc, 0, sizeof(ScanScalars), h->stream));
            ScanParams ps = p;
            ps.tile_stride = stride; ps.n_tiles = n_sample; ps.capacity = 0; ps.hit_slot = ps.hit_pos = nullptr;
            const int grid_s = (int) std::min<uint64_t>((uint64_t) h->sm_count * occ, n_sample);
            scan_probe_kernel<<<grid_s, SCAN_THREADS, 0, h->stream>>>(ps);
            h->metrics.kernel_launches++;
            HGA_CUDA(cudaGetLastError());
            HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            const double est = (double) sc.total_hits * (double) n_tiles / (double) n_sample;
            capacity = (uint64_t) (est * 1.10) + (1ull << 20);
            if (capacity > n_bases) capacity = n_bases;
        }
    }

    for (int attempt = 0; attempt < 2; attempt++) {
        HGA_TRY(h->d_hit_slot.ensure((capacity + 1) * 4));
        HGA_TRY(h->d_hit_pos.ensure((capacity + 1) * 4));
        p.hit_slot = h->d_hit_slot.as<uint32_t>(); p.hit_pos = h->d_hit_pos.as<uint32_t>();
        p.capacity = capacity; p.n_tiles = n_tiles; p.tile_stride = 0;
        HGA_CUDA(cudaMemsetAsync(d_sc, 0, sizeof(ScanScalars), h->stream));
        if (n_tiles > 0 && n_reads > 0) {
            HGA_CUDA(cudaMemsetAsync(p.tile_state, 0, (n_tiles + 1) * 8, h->stream));
            scan_probe_kernel<<<grid_full, SCAN_THREADS, 0, h->stream>>>(p);
            h->metrics.kernel_launches++;
            HGA_CUDA(cudaGetLastError());
        }
        {
            const int blocks = (int) std::min<uint64_t>((n_reads + 256) / 256, 1024);
            scan_finalize_rows_kernel<<<blocks, 256, 0, h->stream>>>(d_read_off, n_reads, n_bases, p.row_off, d_sc);
            h->metrics.kernel_launches++;
            HGA_CUDA(cudaGetLastError());
        }
        HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        if (!sc.overflow) break;
        if (attempt == 1) { hga_set_error("scan: hit buffer overflow after exact resize (internal error)"); return HGA_E_OVERFLOW; }
        capacity = sc.total_hits;   // exact; rerun once
    }
    timer.stop();
    h->n_hits = sc.total_hits;
    h->metrics.n_bases = n_bases; h->metrics.n_reads = n_reads; h->metrics.n_hits = h->n_hits;
    h->have_scan = true;
    return HGA_OK;
}
