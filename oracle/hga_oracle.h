/* ORACLE — TEST INFRASTRUCTURE ONLY. See hga_oracle.c. */
#ifndef HGA_ORACLE_H
#define HGA_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* error codes shared by the record loader */
#define ORC_OK 0
#define ORC_E_NOFILE 1      /* std::invalid_argument  "File with path ... does not exist"  */
#define ORC_E_FORMAT 2      /* std::logic_error       "Unrecognized file format"           */
#define ORC_E_EMPTY 3       /* std::logic_error       "File is empty"                      */
#define ORC_E_HEADER 4      /* std::out_of_range from header.substr(1) on an empty header  */
#define ORC_E_K 5           /* std::invalid_argument  "Kmer size is too big"               */
#define ORC_E_NOMEM 6
#define ORC_E_DIVZERO 7     /* SIGFPE: a file (or the whole input) without a record, avg /= records (SequenceRecordIterator.cpp:64,70) */

typedef struct {
    uint64_t n_reads;
    uint64_t *seq_off;      /* n_reads+1 offsets into seq */
    char *seq;
    uint64_t *hdr_off;      /* n_reads+1 */
    char *hdr;
    uint64_t *qual_off;     /* n_reads+1 (all zero-length for FASTA) */
    char *qual;
    int32_t *file_index;    /* file the record's header line came from */
    /* per-file and aggregate metadata as load_meta_data computes them */
    int n_files;
    uint64_t *f_records, *f_min, *f_max, *f_avg, *f_total;
    int32_t *f_type;        /* 0 FASTA, 1 FASTQ */
    uint64_t a_records, a_min, a_max, a_avg, a_total;
} orc_reads;

int orc_load_reads(const char *const *paths, int n_paths, orc_reads *out);
void orc_free_reads(orc_reads *r);

/* KmerIterator restatement */
uint64_t orc_kmer_windows(const char *seq, uint64_t len, int k, uint64_t *out_kmer, uint32_t *out_pos);
/* load_text_file_kmers restatement: sorted unique canonical values, k = length of last line */
int orc_load_kmers(const char *path, uint64_t **out_sorted, uint64_t *out_n, int *out_k);

/* construct_indices restatement. kmers must be sorted ascending & unique; hit ids index into it.
 * Reads are rows 0..n_reads-1; read id = row + 1. Outputs malloc'd, in (read, position) order. */
int orc_scan(const char *seq, const uint64_t *seq_off, uint64_t n_reads, int k, const uint64_t *kmers, uint64_t n_kmers,
             uint64_t **row_off, uint32_t **hit_kid, uint32_t **hit_pos);
/* inverted index: per k-mer id the READ IDS (1-based), ascending, one entry per occurrence */
int orc_index(const uint64_t *row_off, const uint32_t *hit_kid, uint64_t n_reads, uint64_t n_kmers, uint64_t **inv_off, uint32_t **inv_read);
/* get_connections restatement over the given pivots (read ids; NULL = every read with hits).
 * Emits DIRECTED connections (x = pivot) with score >= min_score, ordered by (x asc, y asc). */
int orc_connections(const uint64_t *row_off, const uint32_t *hit_kid, uint64_t n_reads, const uint64_t *inv_off, const uint32_t *inv_read,
                    const uint32_t *pivots, uint64_t n_pivots, uint64_t min_score, uint64_t *n_conn, uint32_t **cx, uint32_t **cy, uint64_t **cs);
/* canonical order of SURVEY §8a-6 applied in place (score desc, min asc, max asc, x asc) */
void orc_canonical_sort(uint64_t n, uint32_t *cx, uint32_t *cy, uint64_t *cs);
/* union_find restatement (Kruskal in list order). restricted may be NULL. max_size -1 = unlimited.
 * Output: components with size >= min_size: comp_off[n_comp+1], comp_member (root first, then in the
 * reference's append order), tree_off[n_comp+1], tree_x/tree_y (edge endpoints as given). */
int orc_union_find(uint64_t n_edges, const uint32_t *ex, const uint32_t *ey, const uint32_t *restricted, uint64_t n_restricted, int min_size,
                   int max_size, uint64_t *n_comp, uint64_t **comp_off, uint32_t **comp_member, uint64_t **tree_off, uint32_t **tree_x,
                   uint32_t **tree_y);
void orc_free(void *p);

/* SURVEY §8f-1 — engine state after the scaffold stage: merge_components (:349-422), get_connections on the merged
 * state (:301-333), get_component_ids (:727-735). Component ids are read ids (1-based). */
typedef struct orc_engine orc_engine;
orc_engine *orc_engine_new(const uint64_t *row_off, const uint32_t *hit_kid, uint64_t n_reads, uint64_t n_kmers, const uint64_t *inv_off,
                           const uint32_t *inv_read);
void orc_engine_free(orc_engine *e);
int orc_engine_merge(orc_engine *e, uint64_t n_comp, const uint64_t *comp_off, const uint32_t *comp_member, uint32_t *merged_ids);
uint64_t orc_engine_ids(const orc_engine *e, uint64_t min_size, uint32_t *out);
uint64_t orc_engine_component_kmers(const orc_engine *e, uint32_t id, uint32_t *out);
uint64_t orc_engine_component_reads(const orc_engine *e, uint32_t id, uint32_t *out);
uint64_t orc_engine_index_list(const orc_engine *e, uint64_t kmer_id, uint32_t *out);
int orc_engine_connections(const orc_engine *e, const uint32_t *pivots, uint64_t n_pivots, uint64_t min_score, uint64_t *n_conn, uint32_t **cx,
                           uint32_t **cy, uint64_t **cs);

#ifdef __cplusplus
}
#endif
#endif
