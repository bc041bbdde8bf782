// Multi-GPU exchanges over NCCL (placeholder until the exchange kernels land: single-GPU handles never reach these).
#include "hga_internal.cuh"

int hga_comm_exchange_incidence(hga_handle *) { hga_set_error("multi-GPU exchange not built"); return HGA_E_NCCL; }
int hga_comm_reduce_pairs(hga_handle *) { hga_set_error("multi-GPU exchange not built"); return HGA_E_NCCL; }
int hga_comm_allreduce_u64_sum(hga_handle *, uint64_t *, size_t) { hga_set_error("multi-GPU exchange not built"); return HGA_E_NCCL; }
int hga_comm_allreduce_u32_min(hga_handle *, uint32_t *, size_t) { hga_set_error("multi-GPU exchange not built"); return HGA_E_NCCL; }
int hga_comm_allreduce_u32_max(hga_handle *, uint32_t *, size_t) { hga_set_error("multi-GPU exchange not built"); return HGA_E_NCCL; }
int hga_comm_rank(const hga_handle *) { return 0; }
int hga_comm_size(const hga_handle *) { return 1; }
void hga_comm_destroy(hga_handle *) {}

extern "C" int hga_comm_unique_id(void *) { hga_set_error("multi-GPU exchange not built"); return HGA_E_NCCL; }
extern "C" int hga_comm_init(hga_handle *, const void *, int, int, uint64_t) { hga_set_error("multi-GPU exchange not built"); return HGA_E_NCCL; }
