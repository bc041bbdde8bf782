// components: connected components of the selected edge set by lock-free union-find on the GPU.
//
// Replaces union_find(connections, {}, min_size, -1) (clustering/ReadClusteringEngine.cpp:424-489, called at
// :763). With no restricted vertices and no size cap the reference's sequential Kruskal yields exactly the
// connected components of the selected edges, whatever the edge order; vertices that no selected edge touches
// are not part of any component (:427-431 "affected_vertices"), and only components with at least min_size
// members are returned (:482-487).
//
// Hooking always links the larger root under the smaller one (atomicCAS on the root), so after compression
// label[v] is the smallest row of v's component: deterministic and rank-independent. With several GPUs each
// rank hooks its own edges into a replicated parent array and the labels are merged by iterating
// { all-reduce(min) on labels ; hook (v, merged[v]) } to a fixed point.
#include "hga_internal.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace {

__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t v) {
    uint32_t p = parent[v];
    while (p != v) {
        const uint32_t gp = parent[p];
        if (gp != p) parent[v] = gp;     // path halving (benign race: only ever points further up)
        v = p; p = gp;
    }
    return v;
}

__device__ __forceinline__ void uf_union(uint32_t *parent, uint32_t a, uint32_t b) {
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        const uint32_t hi = max(a, b), lo = min(a, b);
        if (atomicCAS(&parent[hi], hi, lo) == hi) return;
    }
}

__global__ void cc_init_kernel(uint32_t *parent, uint32_t *touched, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) { parent[i] = (uint32_t) i; touched[i] = 0; }
}

__global__ void cc_hook_edges_kernel(const uint64_t *__restrict__ key, uint64_t n, uint32_t *parent, uint32_t *touched) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t kk = key[i];
        const uint32_t x = (uint32_t) (kk >> 32), y = (uint32_t) kk;
        touched[x] = 1; touched[y] = 1;
        uf_union(parent, x, y);
    }
}

__global__ void cc_compress_kernel(uint32_t *parent, uint32_t *label, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
        label[i] = uf_find(parent, (uint32_t) i);
}

// multi-GPU merge step: union v with the smallest label any rank proposed for it
__global__ void cc_hook_labels_kernel(const uint32_t *__restrict__ merged, uint32_t *parent, uint64_t n, uint32_t *changed) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t m = merged[i];
        if (uf_find(parent, (uint32_t) i) != uf_find(parent, m)) { uf_union(parent, (uint32_t) i, m); *changed = 1; }
    }
}

__global__ void cc_sizes_kernel(const uint32_t *__restrict__ label, const uint32_t *__restrict__ touched, uint64_t n, uint32_t *size) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
        if (touched[i]) atomicAdd(&size[label[i]], 1u);
}

__global__ void cc_collect_kernel(const uint32_t *__restrict__ size, uint64_t n, uint32_t min_size, uint32_t *out_label, uint32_t *out_size,
                                  unsigned long long *count) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t s = size[i];
        if (s > 0 && s >= min_size) {
            const unsigned long long at = atomicAdd(count, 1ull);
            out_label[at] = (uint32_t) i; out_size[at] = s;
        }
    }
}

}  // namespace

int hga_cc_run(hga_handle *h, int min_size) {
    if (!h->have_selection) { hga_set_error("hga_components: no selection (call hga_select_edges)"); return HGA_E_STATE; }
    h->have_components = false;
    const bool multi = h->comm && hga_comm_size(h) > 1;
    const uint64_t n = h->inc_rows;     // rows are global when a communicator is attached
    const uint64_t M = h->n_selected;
    StageTimer timer(h, &h->metrics.components_ms);

    // layout of d_parent: parent[n] | label[n] | size[n] | touched[n]
    HGA_TRY(h->d_parent.ensure((n + 1) * 4 * 4 + 64));
    uint32_t *parent = h->d_parent.as<uint32_t>(), *label = parent + (n + 1), *size = label + (n + 1), *touched = size + (n + 1);
    HGA_TRY(h->d_comp_scalars.ensure(64));
    unsigned long long *d_count = h->d_comp_scalars.as<unsigned long long>();
    uint32_t *d_changed = reinterpret_cast<uint32_t *>(d_count + 1);

    const int grid_n = (int) std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t) h->sm_count * 8));
    const int grid_m = (int) std::max<uint64_t>(1, std::min<uint64_t>((M + 255) / 256, (uint64_t) h->sm_count * 8));
    cc_init_kernel<<<grid_n, 256, 0, h->stream>>>(parent, touched, n);
    if (M) cc_hook_edges_kernel<<<grid_m, 256, 0, h->stream>>>(h->d_sel_key.as<uint64_t>(), M, parent, touched);
    cc_compress_kernel<<<grid_n, 256, 0, h->stream>>>(parent, label, n);
    h->metrics.kernel_launches += 3;
    HGA_CUDA(cudaGetLastError());

    if (multi) {
        for (int iter = 0; iter < 64; iter++) {
            HGA_TRY(hga_comm_allreduce_u32_min(h, label, n));
            HGA_CUDA(cudaMemsetAsync(d_changed, 0, 4, h->stream));
            cc_hook_labels_kernel<<<grid_n, 256, 0, h->stream>>>(label, parent, n, d_changed);
            cc_compress_kernel<<<grid_n, 256, 0, h->stream>>>(parent, label, n);
            h->metrics.kernel_launches += 2;
            HGA_TRY(hga_comm_allreduce_u32_max(h, d_changed, 1));
            uint32_t changed = 0;
            HGA_CUDA(cudaMemcpyAsync(&changed, d_changed, 4, cudaMemcpyDeviceToHost, h->stream));
            HGA_CUDA(cudaStreamSynchronize(h->stream));
            if (!changed) break;
        }
        // a vertex is "affected" (.cpp:427-431) if any rank saw a selected edge on it
        HGA_TRY(hga_comm_allreduce_u32_max(h, touched, n));
    }

    HGA_CUDA(cudaMemsetAsync(size, 0, (n + 1) * 4, h->stream));
    HGA_CUDA(cudaMemsetAsync(d_count, 0, 8, h->stream));
    cc_sizes_kernel<<<grid_n, 256, 0, h->stream>>>(label, touched, n, size);
    HGA_TRY(h->d_comp_label.ensure((n + 1) * 4 * 2));
    HGA_TRY(h->d_comp_size.ensure((n + 1) * 4 * 2));
    uint32_t *cl = h->d_comp_label.as<uint32_t>(), *cs = h->d_comp_size.as<uint32_t>();
    cc_collect_kernel<<<grid_n, 256, 0, h->stream>>>(size, n, (uint32_t) std::max(min_size, 0), cl + (n + 1), cs + (n + 1), d_count);
    h->metrics.kernel_launches += 2;
    HGA_CUDA(cudaGetLastError());
    unsigned long long nc = 0;
    HGA_CUDA(cudaMemcpyAsync(&nc, d_count, 8, cudaMemcpyDeviceToHost, h->stream));
    HGA_CUDA(cudaStreamSynchronize(h->stream));
    if (nc > 0) {
        size_t tmp_bytes = 0;
        const int bits = (int) std::max<uint32_t>(hga_ceil_log2(n + 1), 1);
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, cl + (n + 1), cl, cs + (n + 1), cs, nc, 0, bits, h->stream));
        HGA_TRY(h->d_sort_tmp.ensure(tmp_bytes + 16));
        HGA_CUDA(cub::DeviceRadixSort::SortPairs(h->d_sort_tmp.p, tmp_bytes, cl + (n + 1), cl, cs + (n + 1), cs, nc, 0, bits, h->stream));
        h->metrics.kernel_launches += 4;
    }
    timer.stop();
    h->n_components = nc;
    h->metrics.n_components = nc;
    h->have_components = true;
    return HGA_OK;
}
