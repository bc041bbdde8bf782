// scan_probe: fused 2-bit pack + canonical k-mer windows + membership test + hit compaction.
//
// Replaces the per-read half of ReadClusteringEngine::construct_indices
// (clustering/ReadClusteringEngine.cpp:246-277) and KmerIterator (common/KmerIterator.cpp:23-76).
//
// Work decomposition (B200: 148 SMs x 24 resident warps, every WARP is an independent worker):
//   * the concatenated base stream is cut into tiles of SCAN_TILE (960) window-end positions; a warp takes a tile
//     ticket, and nothing in the kernel synchronises two warps: no __syncthreads, no inter-tile waiting (a CTA-wide
//     tile with barriers and an in-order look-back were both tried; barrier stalls were a third of all samples);
//   * the warp loads the tile's ASCII bytes (+64 bases of halo = 1024 B) with two coalesced 16 B streaming loads per
//     lane and packs them into TWO shared-memory streams: forward codes MSB-first and complement codes LSB-first.
//     Bytes outside {A,C,G,T} are stored as code 0 in BOTH streams, which is exactly the reference's rule
//     (KmerIterator.cpp:56,62: unordered_map::operator[] default-inserts 0 in both tables);
//   * every window is EXTRACTED from the two streams with funnel shifts (no rolling dependency): the 32 lanes own
//     32 consecutive windows per step, the bit offset inside the word is loop invariant, consecutive steps share a word;
//   * membership probes are made LOCAL: the filter block / key bucket of a window is chosen by the window's
//     minimizer (sliding minimum of hashed canonical m-mers, computed across lanes with shuffles), which several
//     consecutive windows share, so a warp's 32 filter probes fall into a few sectors (hga_internal.cuh);
//   * the per-window loop does nothing but extract, hash and test the filter. Windows that pass are queued (window
//     index + locality hash) and their key bucket is prefetched into L2; everything else - validity against the read
//     boundaries, the position inside the read, the key-bucket probe (the whole 128 B bucket in one round trip) -
//     happens when the queue is drained, DENSE, 32 candidates per warp instruction, re-extracting the k-mer;
//   * hits are staged in position order; a tile's hits are appended with ONE atomicAdd on a global cursor and a copy
//     kernel moves the tile segments into stream order (exclusive scan over the tile counts). CSR row offsets are
//     produced per tile and fixed up with the tile's final offset: no per-hit atomics anywhere;
//   * windows that contain a non-ACGT byte (their two strands are not reverse complements, so the shared minimizer
//     would not be the one the table was built with) skip the filter and recompute the locality hash from the k-mer
//     value; tiles without such bytes never pay for the check.
#include "hga_internal.cuh"

#include <cub/device/device_scan.cuh>

#define SCAN_WARPS 4
#define SCAN_THREADS (SCAN_WARPS * 32)
#define SCAN_TILE 960                            // window-end positions per warp tile
#define SCAN_STEPS (SCAN_TILE / 32)              // 30
#define SCAN_HALO 64
#define SCAN_CHUNKS ((SCAN_TILE + SCAN_HALO) / 16)   // 64 chunks of 16 bases = 2 per lane
#define SCAN_UNROLL 3
#define SCAN_Q 128                               // candidate ring (power of two, >= 31 + 32 * SCAN_UNROLL)
#define SCAN_EXC_WORDS ((SCAN_TILE + SCAN_HALO) / 32 + 2)
#define SCAN_BND 64                              // read boundaries of a tile staged in shared memory (more: global search)

struct ScanScalars {
    unsigned long long ticket;
    unsigned long long total;        // hits appended so far (= all hits when the kernel ends)
    unsigned long long candidates;   // windows that passed the filter (diagnostic: filter false positives = candidates - hits)
    unsigned int overflow;
    unsigned int pad;
};

namespace {

struct ScanParams {
    const char *bases;
    uint64_t n_bases;
    const uint64_t *read_off;
    uint64_t n_reads;
    KmerTable t;
    uint32_t *out_slot, *out_pos;     // temporary hit arrays (tile segments in completion order)
    uint64_t capacity;
    uint64_t *row_off;
    const uint2 *tile_dir;            // per tile: x = first read starting at/after the tile start, y = read holding the tile start
    unsigned long long *tile_tmp_off; // where the tile's segment starts in the temporary arrays
    uint32_t *tile_cnt;               // hits of the tile
    ScanScalars *scalars;
    uint64_t n_tiles;                 // tickets of this launch
    uint64_t tile_begin;              // first tile of this launch (chunked launches while the bases are still arriving)
    uint64_t tile_stride;             // > 0: sampling mode (count only, tile = tile_begin + ticket * stride)
    int diag;                         // HGA_SCAN_DIAG timing experiments (results are WRONG when set): 1 = no key probes, 2 = no filter probes, 4 = no key-sector prefetch
};

// 4 ASCII bytes (little-endian in w, lowest address = first base) -> forward codes (8 bits, first base most
// significant) and complement codes (8 bits, first base LEAST significant). Codes follow KmerIterator.cpp:7-19
// (A0 C1 G2 T3 / complement A3 C2 G1 T0); any other byte gives 0 in both. valid: 0xFF per ACGT byte.
__device__ __forceinline__ void pack4(uint32_t w, uint32_t &fwd8, uint32_t &rc8, uint32_t &valid) {
    valid = __vcmpeq4(w, 0x41414141u) | __vcmpeq4(w, 0x43434343u) | __vcmpeq4(w, 0x47474747u) | __vcmpeq4(w, 0x54545454u);
    uint32_t x = (w >> 1) & 0x03030303u;          // A0 C1 G3 T2
    x ^= (x >> 1) & 0x01010101u;                  // A0 C1 G2 T3
    const uint32_t xc = (x ^ 0x03030303u) & valid;
    x &= valid;
    fwd8 = (x * 0x40100401u) >> 24;               // b0<<6 | b1<<4 | b2<<2 | b3
    rc8 = (xc * 0x01041040u) >> 24;               // c0 | c1<<2 | c2<<4 | c3<<6
}

// byte mask (0xFF / 0x00 per byte) -> 4 bits, bit j = byte j is NOT valid
__device__ __forceinline__ uint32_t invalid_nibble(uint32_t valid) {
    const uint32_t x = ~valid & 0x08040201u;
    return (x | (x >> 8) | (x >> 16) | (x >> 24)) & 0xFu;
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

// L2 eviction policies: the filter is the one structure every window touches (keep it: evict_last); key-table sectors are
// touched once per candidate and must not push it out (evict_first). The base stream uses ld.global.cs, the hits st.global.cs.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint32_t ldg_u32_policy(const uint32_t *a, uint64_t pol) {
    uint32_t v;
    asm("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ ulonglong2 ldg_u64x2_policy(const ulonglong2 *a, uint64_t pol) {
    ulonglong2 v;
    asm("ld.global.nc.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(a), "l"(pol));
    return v;
}

// Key-table lookup: slot of `key` or 0xFFFFFFFF. The home bucket (one 128 B line, prefetched into L2 when the candidate was
// queued) is read whole; a bucket with an empty slot and no match closes the search, a full one sends it to the next bucket of
// the chain (chain_buckets, default 1) and then to the overflow region.
__device__ __forceinline__ uint32_t probe_key(const KmerTable &t, unsigned long long key, uint32_t B) {
    uint32_t slot = 0xFFFFFFFFu;
    const uint32_t home = hga_scale(B, t.n_buckets) * HGA_BUCKET_SLOTS;
    const uint64_t pol_first = l2_policy_evict_first();
    bool open = true;             // chain not yet closed by an empty slot or a match
    // whole bucket (one 128 B line, eight independent 16 B loads) per round trip instead of a dependent sector-by-sector chain:
    // r3h, 10 Gbases: 72.5 ms against 76.9 ms for the sector loop (same table, one-bucket chains)
    #pragma unroll 1
    for (uint32_t b = 0; b < t.chain_buckets && open; b++) {
        const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(t.keys + home + b * HGA_BUCKET_SLOTS);
        ulonglong2 v[HGA_BUCKET_SLOTS / 2];
        #pragma unroll
        for (int i = 0; i < HGA_BUCKET_SLOTS / 2; i++) v[i] = ldg_u64x2_policy(bp + i, pol_first);
        bool empty = false;
        #pragma unroll
        for (int i = 0; i < HGA_BUCKET_SLOTS / 2; i++) {
            if (v[i].x == key) slot = home + b * HGA_BUCKET_SLOTS + 2 * i;
            if (v[i].y == key) slot = home + b * HGA_BUCKET_SLOTS + 2 * i + 1;
            empty |= (v[i].x == HGA_EMPTY_KEY) | (v[i].y == HGA_EMPTY_KEY);
        }
        if (slot != 0xFFFFFFFFu || empty) open = false;
    }
    if (open && t.n_over) {       // chain full: the key, if present, lives in the overflow region
        const uint32_t mask = t.n_over - 1;
        uint32_t q = hga_plain_hash(key) & mask;
        for (;;) {
            const unsigned long long a = __ldg(t.keys + t.n_main + q);
            if (a == key) { slot = t.n_main + q; break; }
            if (a == HGA_EMPTY_KEY) break;
            q = (q + 1) & mask;
        }
    }
    return slot;
}

// shared memory of one warp
struct WarpTile {
    uint32_t fwd[SCAN_CHUNKS + 4];
    uint32_t rc[SCAN_CHUNKS + 4];
    uint32_t exc[SCAN_EXC_WORDS];    // bit p: staged base p is not one of ACGT (written as u16 halves by the packers)
    uint2 q[SCAN_Q];                 // x = locality hash B, y = window index in the tile | 0x8000 when the window holds a non-ACGT byte
    uint32_t st_slot[SCAN_TILE];     // hits of the tile in position order
    uint16_t st_w[SCAN_TILE];        // their window index
    int32_t bnd[SCAN_BND];           // tile-local first base of reads r_lo, r_lo + 1, ... (n_bnd of them)
};

struct TileCtx {
    const ScanParams *p;
    WarpTile *T;
    uint64_t tile_start, r_lo, r_hi;
    uint32_t n_bnd;                  // 0: too many reads in this tile, search read_off in global memory
    int n_loc;
    unsigned long long kmask;
};

// one level of the cross-lane sliding minimum: value of `cur` d positions back (previous step's register for
// the first d lanes). __shfl_sync takes the source lane modulo 32.
__device__ __forceinline__ uint32_t back(uint32_t cur, uint32_t prev, int d, int lane) {
    return __shfl_sync(0xFFFFFFFFu, lane >= 32 - d ? prev : cur, lane - d);
}

struct MinState { uint32_t g, a1, a2, r; };

// minimum of the hashed m-mers ending at positions e-skip-W+1 .. e-skip (e = this lane's position in this step):
// log-step doubling (a1 = 2 positions, a2 = 4), then one overlapping combine for the W that are not powers of two
template<int W>
__device__ __forceinline__ uint32_t window_min(uint32_t g, MinState &ms, int skip, int lane) {
    const uint32_t a1 = min(g, back(g, ms.g, 1, lane));
    uint32_t a2 = 0, r;
    if (W >= 4) a2 = min(a1, back(a1, ms.a1, 2, lane));
    if (W == 2) r = a1;
    else if (W == 3) r = min(a1, back(g, ms.g, 2, lane));
    else if (W == 4) r = a2;
    else if (W == 5) r = min(a2, back(g, ms.g, 4, lane));
    else if (W == 6) r = min(a2, back(a1, ms.a1, 4, lane));
    else if (W == 7) r = min(a2, back(a2, ms.a2, 3, lane));
    else r = min(a2, back(a2, ms.a2, 4, lane));
    uint32_t gm = r;
    if (skip) gm = back(r, ms.r, skip, lane);
    ms.g = g; ms.a1 = a1; ms.a2 = a2; ms.r = r;
    return gm;
}

// canonical-k-mer pieces of the window ending at tile-local position e
__device__ __forceinline__ void extract_window(const uint32_t *s_fwd, const uint32_t *s_rc, int e, int k, unsigned long long kmask,
                                               unsigned long long &fwd, unsigned long long &rc) {
    const int je = e + SCAN_HALO, js = je - k + 1;     // staged coordinates (0 = tile_start - SCAN_HALO)
    const int w0 = js >> 4, o = (js & 15) * 2;
    const uint32_t F0 = s_fwd[w0], F1 = s_fwd[w0 + 1], F2 = s_fwd[w0 + 2];
    const uint32_t R0 = s_rc[w0], R1 = s_rc[w0 + 1], R2 = s_rc[w0 + 2];
    fwd = ((((unsigned long long) __funnelshift_l(F1, F0, o)) << 32) | __funnelshift_l(F2, F1, o)) >> (64 - 2 * k);
    rc = ((((unsigned long long) __funnelshift_r(R1, R2, o)) << 32) | __funnelshift_r(R0, R1, o)) & kmask;
}

// tile-local first base of the read that holds tile-local base e
__device__ __forceinline__ int read_start_of(const TileCtx &c, int e) {
    if (c.n_bnd) {
        uint32_t lo = 0, hi = c.n_bnd - 1;
        while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (c.T->bnd[mid] <= e) lo = mid; else hi = mid - 1; }
        return c.T->bnd[lo];
    }
    const uint64_t r = find_read(c.p->read_off, c.r_lo, c.r_hi, c.tile_start + (uint64_t) e);
    return clamp_local(__ldg(&c.p->read_off[r]), c.tile_start);
}

// n (<= 32) queued candidates starting at ring position head: re-extract the k-mer, check the window against the
// read boundaries, probe the key table; hits are appended to the staging area in position order
__device__ __forceinline__ uint32_t drain_queue(uint32_t head, uint32_t n, uint32_t st_count, int lane, WarpTile *Tp, const ScanParams *pp,
                                             uint64_t tile_start, uint64_t r_lo, uint64_t r_hi, uint32_t n_bnd, int n_loc) {
    TileCtx c;
    c.p = pp; c.T = Tp; c.tile_start = tile_start; c.r_lo = r_lo; c.r_hi = r_hi; c.n_bnd = n_bnd; c.n_loc = n_loc;
    const KmerTable &t = pp->t;
    WarpTile &T = *Tp;
    const int k = t.geom.k;
    c.kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
    uint32_t slot = 0xFFFFFFFFu, widx = 0;
    if ((uint32_t) lane < n) {
        const uint2 qe = T.q[(head + lane) & (SCAN_Q - 1)];
        uint32_t B = qe.x;
        widx = qe.y & 0x7FFFu;
        const int e = (int) widx;
        const int start = read_start_of(c, e);
        if (e < c.n_loc && e - start + 1 >= k && !(pp->diag & 1)) {
            unsigned long long fwd, rc;
            extract_window(T.fwd, T.rc, e, k, c.kmask, fwd, rc);
            const unsigned long long key = fwd < rc ? fwd : rc;                       // KmerIterator.cpp:69
            if (qe.y & 0x8000u) B = hga_locality_hash(key, t.geom);
            slot = probe_key(t, key, B);
        }
    }
    const bool hit = slot != 0xFFFFFFFFu;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
    if (hit) {
        const uint32_t idx = st_count + __popc(bal & ((1u << lane) - 1));
        T.st_slot[idx] = slot; T.st_w[idx] = (uint16_t) widx;
    }
    return st_count + __popc(bal);
}

// W = 0: locality hash = plain k-mer hash (small k); W = 2..8: minimizer over W m-mers
template<bool EXC, int W>
__device__ __forceinline__ uint32_t scan_tile_windows(const TileCtx &c, int lane, uint32_t &n_cand) {
    const ScanParams &p = *c.p;
    WarpTile &T = *c.T;
    const KmerGeom &geo = p.t.geom;
    const int k = geo.k;
    const unsigned long long kmask = c.kmask;
    const uint32_t kbits = (k == 32) ? 0xFFFFFFFFu : ((1u << k) - 1);
    const uint32_t lane_lt = (1u << lane) - 1;
    const uint32_t *filter = p.t.filter;
    const uint32_t n_blocks = p.t.n_blocks, n_buckets = p.t.n_buckets;
    const uint64_t *keys = p.t.keys;
    const bool diag2 = (p.diag & 2) != 0, diag4 = (p.diag & 4) != 0;
    const int skip = W ? geo.skip : 0;
    // staged coordinates (0 = tile_start - SCAN_HALO) of this lane's first window; every step moves 32 bases = 2 words,
    // so the bit offset inside the word is loop invariant and consecutive steps share a word
    const int js0 = lane + SCAN_HALO - k + 1;
    const uint32_t *pf = T.fwd + (js0 >> 4), *pr = T.rc + (js0 >> 4), *pe = T.exc + (js0 >> 5);
    const int o = (js0 & 15) * 2, oe = js0 & 31, fsh = 64 - 2 * k;
    uint32_t q_head = 0, q_tail = 0, st_count = 0;   // warp uniform

    MinState ms = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (W) {   // warm-up: the 32 positions before the tile feed the first sliding minima
        const unsigned long long fwd = ((((unsigned long long) __funnelshift_l(pf[-1], pf[-2], o)) << 32) | __funnelshift_l(pf[0], pf[-1], o)) >> fsh;
        const unsigned long long rc = ((((unsigned long long) __funnelshift_r(pr[-1], pr[0], o)) << 32) | __funnelshift_r(pr[-2], pr[-1], o)) & kmask;
        (void) window_min<W ? W : 2>(hga_mmer_hash((uint32_t) fwd & geo.mmask, (uint32_t) (rc >> geo.rc_shift)), ms, skip, lane);
    }
    uint32_t F0 = pf[0], R0 = pr[0];
    const uint64_t pol_last = l2_policy_evict_last();

    // (Fetching the next tile's bytes one tile ahead with cp.async was tried: slower, -5 % at 2 Gbases and -14 % at 10 Gbases:
    // cp.async has no evict-first path, so the base stream displaced the filter from L2.)
    // (Splitting the drain into "request the key sector now, compare when the queue has filled again" was tried as well: no gain,
    // r2c: 13.7 vs 13.2 ms at 2 Gbases, 76.4 vs 75.7 ms at 10 Gbases.)
    // (Requesting the filter words one group ahead of testing them was tried: no gain - the probes are bound by L1 wavefront
    // throughput, ~10 distinct lines per warp load, not by their latency - and the extra live registers spilled around the
    // drain call.)
    #pragma unroll 1
    for (int s0 = 0; s0 < SCAN_STEPS; s0 += SCAN_UNROLL, pf += 2 * SCAN_UNROLL, pr += 2 * SCAN_UNROLL, pe += SCAN_UNROLL) {
        uint32_t msk[SCAN_UNROLL], fw[SCAN_UNROLL], Bv[SCAN_UNROLL];
        bool exc[SCAN_UNROLL];
        #pragma unroll
        for (int u = 0; u < SCAN_UNROLL; u++) {
            const uint32_t F1 = pf[2 * u + 1], F2 = pf[2 * u + 2], R1 = pr[2 * u + 1], R2 = pr[2 * u + 2];
            const unsigned long long fwd = ((((unsigned long long) __funnelshift_l(F1, F0, o)) << 32) | __funnelshift_l(F2, F1, o)) >> fsh;
            const unsigned long long rc = ((((unsigned long long) __funnelshift_r(R1, R2, o)) << 32) | __funnelshift_r(R0, R1, o)) & kmask;
            F0 = F2; R0 = R2;
            const unsigned long long canon = fwd < rc ? fwd : rc;                     // KmerIterator.cpp:69
            if (W) Bv[u] = hga_locality_from_min(window_min<W ? W : 2>(hga_mmer_hash((uint32_t) fwd & geo.mmask, (uint32_t) (rc >> geo.rc_shift)), ms, skip, lane));
            else Bv[u] = hga_plain_hash(canon);
            exc[u] = false;
            if (EXC && W) exc[u] = (__funnelshift_r(pe[u], pe[u + 1], oe) & kbits) != 0;
            const uint32_t hb = hga_bits_hash(canon);
            msk[u] = hga_bits_mask(hb);
            fw[u] = diag2 ? 0u : ldg_u32_policy(filter + ((hga_scale(Bv[u], n_blocks) << 3) | hga_bits_word(hb)), pol_last);
        }
        #pragma unroll
        for (int u = 0; u < SCAN_UNROLL; u++) {
            const bool pass = exc[u] || (fw[u] & msk[u]) == msk[u];
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, pass);
            if (pass) {
                T.q[(q_tail + __popc(bal & lane_lt)) & (SCAN_Q - 1)] = make_uint2(Bv[u], (uint32_t) (lane + 32 * (s0 + u)) | (exc[u] ? 0x8000u : 0u));
                // start the key sector's trip from HBM now; the drain that reads it runs a few steps later
                if (!diag4) {
                    const uint64_t *kb = keys + (size_t) hga_scale(Bv[u], n_buckets) * HGA_BUCKET_SLOTS;
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(kb));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(kb + HGA_BUCKET_SLOTS / 2));
                }
            }
            q_tail += __popc(bal);
        }
        __syncwarp();
        while (q_tail - q_head >= 32) {
            st_count = drain_queue(q_head, 32, st_count, lane, c.T, c.p, c.tile_start, c.r_lo, c.r_hi, c.n_bnd, c.n_loc);
            q_head += 32;
        }
        __syncwarp();
    }
    if (q_tail != q_head) st_count = drain_queue(q_head, q_tail - q_head, st_count, lane, c.T, c.p, c.tile_start, c.r_lo, c.r_hi, c.n_bnd, c.n_loc);
    n_cand += q_tail;
    return st_count;
}

template<bool EXC>
__device__ __forceinline__ uint32_t scan_tile_dispatch(const TileCtx &c, int lane, uint32_t &n_cand) {
    const KmerGeom &geo = c.p->t.geom;
    switch (geo.use_min ? geo.W : 0) {
        case 2: return scan_tile_windows<EXC, 2>(c, lane, n_cand);
        case 3: return scan_tile_windows<EXC, 3>(c, lane, n_cand);
        case 4: return scan_tile_windows<EXC, 4>(c, lane, n_cand);
        case 5: return scan_tile_windows<EXC, 5>(c, lane, n_cand);
        case 6: return scan_tile_windows<EXC, 6>(c, lane, n_cand);
        case 7: return scan_tile_windows<EXC, 7>(c, lane, n_cand);
        case 8: return scan_tile_windows<EXC, 8>(c, lane, n_cand);
        default: return scan_tile_windows<EXC, 0>(c, lane, n_cand);
    }
}

// 63 registers, 30.9 KB of shared memory per CTA: 7 CTAs = 28 independent warps per SM. The parameters are __grid_constant__ and
// every helper is inlined so that NOTHING lives in local memory: a by-value parameter struct whose address is taken, and a
// tile context handed to a non-inlined drain function, cost 26 % of the scan time when they did (r1y / r1z / r2a).
__global__ void __launch_bounds__(SCAN_THREADS, 7) scan_probe_kernel(const __grid_constant__ ScanParams p) {
    __shared__ WarpTile s_tiles[SCAN_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpTile &T = s_tiles[warp];
    uint16_t *s_exc16 = reinterpret_cast<uint16_t *>(T.exc);
    uint32_t n_cand = 0;
    const int kk = p.t.geom.k;

    for (uint32_t i = lane; i < SCAN_EXC_WORDS; i += 32) T.exc[i] = 0;
    if (lane < 4) { T.fwd[SCAN_CHUNKS + lane] = 0; T.rc[SCAN_CHUNKS + lane] = 0; }

    for (;;) {
        unsigned long long ticket = 0;
        if (lane == 0) ticket = atomicAdd(&p.scalars->ticket, 1ull);
        ticket = __shfl_sync(0xFFFFFFFFu, ticket, 0);
        if (ticket >= p.n_tiles) break;
        const uint64_t tile = p.tile_begin + (p.tile_stride ? ticket * p.tile_stride : ticket);
        const uint64_t tile_start = tile * SCAN_TILE;
        const int n_loc = (int) min((uint64_t) SCAN_TILE, p.n_bases - tile_start);      // window ends in this tile

        // ---- load + pack (chunk c covers global bases [tile_start - HALO + 16c, +16)) --------------------------
        bool any_bad = false;
        #pragma unroll
        for (int it = 0; it < SCAN_CHUNKS / 32; it++) {
            const int c = lane + 32 * it;
            const int64_t g = (int64_t) tile_start - SCAN_HALO + 16 * (int64_t) c;
            uint4 v = make_uint4(0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u);   // outside the stream: 'A' (never inside a valid window)
            if (g >= 0 && (uint64_t) g + 16 <= p.n_bases) {
                v = __ldcs(reinterpret_cast<const uint4 *>(p.bases + g));
            } else if (g >= 0 && (uint64_t) g < p.n_bases) {
                uint32_t t[4] = {0x41414141u, 0x41414141u, 0x41414141u, 0x41414141u};
                for (int i = 0; i < 16 && (uint64_t) g + i < p.n_bases; i++) {
                    const int sh = 8 * (i & 3);
                    t[i >> 2] = (t[i >> 2] & ~(0xFFu << sh)) | ((uint32_t) (unsigned char) p.bases[g + i] << sh);
                }
                v = make_uint4(t[0], t[1], t[2], t[3]);
            }
            uint32_t f0, f1, f2, f3, r0, r1, r2, r3, v0, v1, v2, v3;
            pack4(v.x, f0, r0, v0); pack4(v.y, f1, r1, v1); pack4(v.z, f2, r2, v2); pack4(v.w, f3, r3, v3);
            T.fwd[c] = (f0 << 24) | (f1 << 16) | (f2 << 8) | f3;
            T.rc[c] = r0 | (r1 << 8) | (r2 << 16) | (r3 << 24);
            uint32_t bad = 0;
            if ((v0 & v1 & v2 & v3) != 0xFFFFFFFFu) {
                bad = invalid_nibble(v0) | (invalid_nibble(v1) << 4) | (invalid_nibble(v2) << 8) | (invalid_nibble(v3) << 12);
                any_bad = true;
            }
            s_exc16[c] = (uint16_t) bad;
        }
        // read boundaries of the tile (reads r_lo .. r_hi start at or before the tile's last base)
        const uint2 d0 = __ldg(&p.tile_dir[tile]), d1 = __ldg(&p.tile_dir[tile + 1]);
        TileCtx ctx;
        ctx.p = &p; ctx.T = &T; ctx.tile_start = tile_start; ctx.n_loc = n_loc;
        ctx.r_lo = d0.y; ctx.r_hi = min((uint64_t) d1.y, p.n_reads - 1);
        ctx.kmask = (kk == 32) ? ~0ull : ((1ull << (2 * kk)) - 1);
        const uint64_t nb = ctx.r_hi - ctx.r_lo + 1;
        ctx.n_bnd = nb <= SCAN_BND ? (uint32_t) nb : 0u;
        if (nb <= SCAN_BND) for (uint32_t i = lane; i < nb; i += 32) T.bnd[i] = clamp_local(__ldg(&p.read_off[ctx.r_lo + i]), tile_start);
        const bool exc_any = __any_sync(0xFFFFFFFFu, any_bad);
        __syncwarp();

        // ---- windows: lane owns window end e_loc = step * 32 + lane --------------------------------------------
        uint32_t total;
        if (exc_any) total = scan_tile_dispatch<true>(ctx, lane, n_cand);
        else total = scan_tile_dispatch<false>(ctx, lane, n_cand);
        __syncwarp();

        // ---- append the tile's hits (one atomic), tile directory, tile-local CSR row offsets ---------------------
        if (p.tile_stride) {
            if (lane == 0 && total) atomicAdd(&p.scalars->total, (unsigned long long) total);
            continue;
        }
        unsigned long long tile_off = 0;
        if (lane == 0) {
            tile_off = atomicAdd(&p.scalars->total, (unsigned long long) total);
            p.tile_tmp_off[tile] = tile_off; p.tile_cnt[tile] = total;
            if (tile_off + total > p.capacity) p.scalars->overflow = 1;
        }
        tile_off = __shfl_sync(0xFFFFFFFFu, tile_off, 0);
        // rows starting in this tile: number of tile hits whose window ends before the read's first base
        for (uint64_t rr = (uint64_t) d0.x + lane; rr <= ctx.r_hi; rr += 32) {
            const uint64_t ro = __ldg(&p.read_off[rr]);
            if (ro >= tile_start && ro < tile_start + (uint64_t) n_loc) {
                const uint32_t q = (uint32_t) (ro - tile_start);
                uint32_t lo = 0, hi = total;
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (T.st_w[mid] < q) lo = mid + 1; else hi = mid; }
                p.row_off[rr] = lo;       // tile-local; scan_fix_rows_kernel adds the tile's final offset
            }
        }
        if (tile_off + total <= p.capacity) {
            for (uint32_t i = lane; i < total; i += 32) {
                const int e = T.st_w[i];
                __stcs(&p.out_slot[tile_off + i], T.st_slot[i]);
                __stcs(&p.out_pos[tile_off + i], (uint32_t) (e - read_start_of(ctx, e) + 1));   // KmerIterator::position_in_sequence
            }
        }
        __syncwarp();
    }
    if (lane == 0 && n_cand) atomicAdd(&p.scalars->candidates, (unsigned long long) n_cand);
}

// per tile: x = first r in [0, n_reads] with read_off[r] >= tile start, y = last r in [0, n_reads) with read_off[r] <= tile start
__global__ void scan_tile_dir_kernel(const uint64_t *__restrict__ read_off, uint64_t n_reads, uint64_t n_tiles, uint2 *dir) {
    for (uint64_t t = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; t <= n_tiles; t += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t pos = t * SCAN_TILE;
        uint64_t lo = 0, hi = n_reads;        // lower bound over read_off[0 .. n_reads]
        while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (read_off[mid] < pos) lo = mid + 1; else hi = mid; }
        const uint64_t first = lo;
        lo = 0; hi = n_reads - 1;
        while (lo < hi) { const uint64_t mid = (lo + hi + 1) >> 1; if (read_off[mid] <= pos) lo = mid; else hi = mid - 1; }
        dir[t] = make_uint2((uint32_t) first, (uint32_t) lo);
    }
}

// tile segments (completion order) -> stream order; one warp per tile
__global__ void scan_reorder_kernel(const uint32_t *__restrict__ tmp_slot, const uint32_t *__restrict__ tmp_pos,
                                    const unsigned long long *__restrict__ tile_tmp_off, const uint32_t *__restrict__ tile_cnt,
                                    const unsigned long long *__restrict__ tile_off, uint64_t n_tiles, uint32_t *out_slot, uint32_t *out_pos) {
    const uint64_t warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    const uint64_t w = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint64_t t = w; t < n_tiles; t += warps) {
        const unsigned long long src = tile_tmp_off[t], dst = tile_off[t], n = tile_cnt[t];
        for (unsigned long long i = lane; i < n; i += 32) {
            __stcs(&out_slot[dst + i], __ldcs(&tmp_slot[src + i]));
            __stcs(&out_pos[dst + i], __ldcs(&tmp_pos[src + i]));
        }
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

int launch_scan(hga_handle *h, const ScanParams &p, int grid) {
    // experiment switch: HGA_SCAN_PAD_KB pads every CTA with unused dynamic shared memory (fewer resident CTAs per SM)
    size_t pad = 0;
    if (const char *e = getenv("HGA_SCAN_PAD_KB")) {
        pad = (size_t) atoi(e) * 1024;
        cudaFuncSetAttribute(scan_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pad);
        int occ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, scan_probe_kernel, SCAN_THREADS, pad);
        grid = std::min(grid, h->sm_count * std::max(occ, 1));
    }
    scan_probe_kernel<<<grid, SCAN_THREADS, pad, h->stream>>>(p);
    HGA_CUDA(cudaGetLastError());
    h->metrics.kernel_launches++;
    return HGA_OK;
}

}  // namespace

int hga_scan_run(hga_handle *h, const char *d_bases, const uint64_t *d_read_off, uint64_t n_reads, uint64_t n_bases, const char *h_bases) {
    h->have_scan = h->have_index = h->have_pairs = h->have_selection = h->have_components = h->have_enrichment = false;
    h->n_reads = n_reads; h->n_bases = n_bases; h->n_hits = 0;
    if (n_reads >= (1ull << 32) - 1) { hga_set_error("hga_scan: more than 2^32-2 reads per GPU"); return HGA_E_ARG; }
    HGA_TRY(h->d_row_off.ensure((n_reads + 1) * 8));
    HGA_TRY(h->d_scan_scalars.ensure(sizeof(ScanScalars)));
    ScanScalars *d_sc = h->d_scan_scalars.as<ScanScalars>();
    const uint64_t n_tiles = (n_reads == 0) ? 0 : (n_bases + SCAN_TILE - 1) / SCAN_TILE;
    HGA_TRY(h->d_tile_state.ensure((n_tiles + 2) * (8 + 8 + 4)));   // tile directory: tmp offset | final offset | count
    unsigned long long *tile_tmp_off = h->d_tile_state.as<unsigned long long>();
    unsigned long long *tile_off = tile_tmp_off + (n_tiles + 2);
    uint32_t *tile_cnt = reinterpret_cast<uint32_t *>(tile_off + (n_tiles + 2));
    HGA_TRY(h->d_tile_dir.ensure((n_tiles + 2) * sizeof(uint2)));

    ScanParams p;
    memset(&p, 0, sizeof(p));
    p.bases = d_bases; p.n_bases = n_bases; p.read_off = d_read_off; p.n_reads = n_reads;
    p.t = h->table;
    p.row_off = h->d_row_off.as<uint64_t>();
    p.tile_dir = h->d_tile_dir.as<uint2>();
    p.tile_tmp_off = tile_tmp_off; p.tile_cnt = tile_cnt;
    p.scalars = d_sc;
    if (const char *e = getenv("HGA_SCAN_DIAG")) p.diag = atoi(e);

    int occ = 0;
    HGA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, scan_probe_kernel, SCAN_THREADS, 0));
    if (occ < 1) occ = 1;
    const int grid_full = (int) std::min<uint64_t>((uint64_t) h->sm_count * occ, std::max<uint64_t>((n_tiles + SCAN_WARPS - 1) / SCAN_WARPS, 1));

    StageTimer timer(h, &h->metrics.scan_ms);
    ScanScalars sc;
    memset(&sc, 0, sizeof(sc));
    uint64_t capacity = 0, E = 0;
    // Host source (hga_scan): the bases travel in chunks on a second stream and every chunk is scanned as soon as it has
    // landed (a tile only needs bases at or before its own end, so chunk c can run while chunk c + 1 is in flight).
    uint64_t chunk_mb = 256;                                                      // bases per chunk; HGA_SCAN_CHUNK_MB lets the tests reach this path with small inputs
    if (const char *e = getenv("HGA_SCAN_CHUNK_MB")) chunk_mb = std::max(1, atoi(e));
    const uint64_t chunk_tiles = (chunk_mb << 20) / SCAN_TILE;
    const bool pipelined = h_bases != nullptr && n_tiles > 2 * chunk_tiles && h->copy_stream != nullptr;
    if (h_bases != nullptr && !pipelined && n_bases) {
        StageTimer t(h, &h->metrics.h2d_ms, true);
        HGA_CUDA(cudaMemcpyAsync(const_cast<char *>(d_bases), h_bases, n_bases, cudaMemcpyHostToDevice, h->stream));
        t.stop();
    }
    if (n_tiles > 0) {
        scan_tile_dir_kernel<<<(int) std::min<uint64_t>((n_tiles + 256) / 256, (uint64_t) h->sm_count * 8), 256, 0, h->stream>>>(d_read_off, n_reads, n_tiles,
                                                                                                                               h->d_tile_dir.as<uint2>());
        h->metrics.kernel_launches++;
        HGA_CUDA(cudaGetLastError());
    }
    bool done = false;
    if (pipelined) {
        const uint64_t n_chunks = (n_tiles + chunk_tiles - 1) / chunk_tiles;
        std::vector<cudaEvent_t> landed(n_chunks, nullptr);
        auto chunk_bytes = [&](uint64_t c, uint64_t &lo, uint64_t &hi) { lo = c * chunk_tiles * SCAN_TILE; hi = std::min<uint64_t>(n_bases, (c + 1) * chunk_tiles * SCAN_TILE); };
        cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
        HGA_CUDA(cudaEventCreate(&ev_begin)); HGA_CUDA(cudaEventCreate(&ev_end));
        HGA_CUDA(cudaEventRecord(ev_begin, h->stream));
        HGA_CUDA(cudaStreamWaitEvent(h->copy_stream, ev_begin, 0));               // the destination buffer was (re)allocated on h->stream
        for (uint64_t c = 0; c < n_chunks; c++) {
            uint64_t lo, hi;
            chunk_bytes(c, lo, hi);
            HGA_CUDA(cudaMemcpyAsync(const_cast<char *>(d_bases) + lo, h_bases + lo, hi - lo, cudaMemcpyHostToDevice, h->copy_stream));
            HGA_CUDA(cudaEventCreateWithFlags(&landed[c], cudaEventDisableTiming));
            HGA_CUDA(cudaEventRecord(landed[c], h->copy_stream));
        }
        HGA_CUDA(cudaEventRecord(ev_end, h->copy_stream));
        // capacity from a count-only sample of the first chunk (1 tile in 16)
        HGA_CUDA(cudaStreamWaitEvent(h->stream, landed[0], 0));
        {
            const uint64_t stride = 16, n_sample = (chunk_tiles + stride - 1) / stride;
            HGA_CUDA(cudaMemsetAsync(d_sc, 0, sizeof(ScanScalars), h->stream));
            ScanParams ps = p;
            ps.tile_stride = stride; ps.n_tiles = n_sample; ps.tile_begin = 0; ps.capacity = 0;
            HGA_TRY(launch_scan(h, ps, (int) std::min<uint64_t>((uint64_t) h->sm_count * occ, (n_sample + SCAN_WARPS - 1) / SCAN_WARPS)));
            HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            const double est = (double) sc.total * (double) n_tiles / (double) n_sample;
            capacity = std::min<uint64_t>(n_bases, (uint64_t) (est * 1.15) + (1ull << 20));
        }
        HGA_TRY(h->d_sort_a.ensure((capacity + 1) * 4));
        HGA_TRY(h->d_sort_b.ensure((capacity + 1) * 4));
        p.out_slot = h->d_sort_a.as<uint32_t>(); p.out_pos = h->d_sort_b.as<uint32_t>();
        p.capacity = capacity; p.tile_stride = 0;
        HGA_CUDA(cudaMemsetAsync(d_sc, 0, sizeof(ScanScalars), h->stream));
        for (uint64_t c = 0; c < n_chunks; c++) {
            HGA_CUDA(cudaStreamWaitEvent(h->stream, landed[c], 0));
            p.tile_begin = c * chunk_tiles; p.n_tiles = std::min<uint64_t>(chunk_tiles, n_tiles - p.tile_begin);
            HGA_CUDA(cudaMemsetAsync(&d_sc->ticket, 0, 8, h->stream));
            HGA_TRY(launch_scan(h, p, (int) std::min<uint64_t>((uint64_t) h->sm_count * occ, (p.n_tiles + SCAN_WARPS - 1) / SCAN_WARPS)));
        }
        HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->copy_stream));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev_begin, ev_end);
        h->metrics.h2d_ms = ms;
        for (cudaEvent_t e : landed) cudaEventDestroy(e);
        cudaEventDestroy(ev_begin); cudaEventDestroy(ev_end);
        E = sc.total;
        p.tile_begin = 0;
        if (!sc.overflow) done = true; else capacity = E;     // the bases are resident now: one plain rerun with the exact size
    } else if (n_tiles > 0) {
        // capacity of the hit arrays: exact upper bound for small inputs, otherwise estimated from a strided
        // count-only sample (1 tile in 64)
        const uint64_t small_limit = 32ull << 20;
        if (n_bases <= small_limit) {
            capacity = n_bases;
        } else {
            const uint64_t stride = 64;
            const uint64_t n_sample = (n_tiles + stride - 1) / stride;
            HGA_CUDA(cudaMemsetAsync(d_sc, 0, sizeof(ScanScalars), h->stream));
            ScanParams ps = p;
            ps.tile_stride = stride; ps.n_tiles = n_sample; ps.capacity = 0;
            const int grid_s = (int) std::min<uint64_t>((uint64_t) h->sm_count * occ, (n_sample + SCAN_WARPS - 1) / SCAN_WARPS);
            HGA_TRY(launch_scan(h, ps, grid_s));
            HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            const double est = (double) sc.total * (double) n_tiles / (double) n_sample;
            capacity = (uint64_t) (est * 1.10) + (1ull << 20);
            if (capacity > n_bases) capacity = n_bases;
        }
    }

    for (int attempt = 0; attempt < 2 && n_tiles > 0 && !done; attempt++) {
        HGA_TRY(h->d_sort_a.ensure((capacity + 1) * 4));      // temporaries (reused by the index sort later)
        HGA_TRY(h->d_sort_b.ensure((capacity + 1) * 4));
        p.out_slot = h->d_sort_a.as<uint32_t>(); p.out_pos = h->d_sort_b.as<uint32_t>();
        p.capacity = capacity; p.n_tiles = n_tiles; p.tile_begin = 0; p.tile_stride = 0;
        HGA_CUDA(cudaMemsetAsync(d_sc, 0, sizeof(ScanScalars), h->stream));
        HGA_TRY(launch_scan(h, p, grid_full));
        HGA_CUDA(cudaMemcpyAsync(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost, h->stream));
        HGA_CUDA(cudaStreamSynchronize(h->stream));
        E = sc.total;
        if (!sc.overflow) break;
        if (attempt == 1) { hga_set_error("scan: hit buffer overflow after exact resize (internal error)"); return HGA_E_OVERFLOW; }
        capacity = E;   // exact; rerun once
    }

    HGA_TRY(h->d_hit_slot.ensure((E + 1) * 4));
    HGA_TRY(h->d_hit_pos.ensure((E + 1) * 4));
    if (n_tiles > 0) {
        size_t tmp_bytes = 0;
        HGA_CUDA(cudaMemsetAsync(tile_cnt + n_tiles, 0, 4, h->stream));
        HGA_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, tmp_bytes, tile_cnt, tile_off, cub::Sum(), 0ull, n_tiles + 1, h->stream));   // 64-bit accumulator
        HGA_TRY(h->d_sort_tmp.ensure(tmp_bytes + 16));
        HGA_CUDA(cub::DeviceScan::ExclusiveScan(h->d_sort_tmp.p, tmp_bytes, tile_cnt, tile_off, cub::Sum(), 0ull, n_tiles + 1, h->stream));
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
    h->metrics.n_bases = n_bases; h->metrics.n_reads = n_reads; h->metrics.n_hits = E; h->metrics.n_candidates = sc.candidates;
    h->have_scan = true;
    return HGA_OK;
}
