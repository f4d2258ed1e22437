/*
 * ycnr_host.h — host-side (CPU, no CUDA) front end that feeds the ALS hot path.
 *
 * This is the data side of the drop-in boundary: it produces exactly the inputs
 * the reference master hands to its workers, from an in-memory ratings table
 * instead of PostgreSQL.  Everything here is integer/byte work and is compared
 * bit-exactly against oracle/front_end.py in tests/.
 *
 * Reference interfaces restated (file:line into the upstream repository):
 *   ratings table layout ........ data/db-schema.sql:887-893 (user_list_id,item_id,rating,dataset_type)
 *   split rule (Q9) ............. lib/emf/EmfLord.js:450-473 (+ knuth-shuffle)
 *   per-row stats ............... lib/emf/EmfLord.js:48-128
 *   portion planner (Q6) ........ lib/emf/EmfLord.js:510-612
 *   portion fetch filter/order .. lib/emf/EmfMaster.js:501-529
 *   portion CSR + last-rating drop (Q2) .. lib/emf/EmfMaster.js:571-614
 *
 * Ids are 0-based everywhere in this API (DB id = index + 1, EmfMaster.js:584-586).
 * All functions return 0 on success, non-zero on error (ycnr_host_last_error()).
 */
#ifndef YCNR_HOST_H
#define YCNR_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* ycnr_host_last_error(void);

/* Counter-based PRNG shared by generator, split and factor init (the reference is
 * unseeded: Math.random, EmfLord.js:463, EmfBase.js:486-497). */
uint64_t ycnr_mix64(uint64_t seed, uint64_t a, uint64_t b);
double ycnr_u01(uint64_t h); /* (h >> 11) * 2^-53 in [0,1) */

/* ---- synthetic ratings table ------------------------------------------------ */

/* Power-law (Pareto alpha) per-user rating counts in [1, items], rescaled so the
 * counts sum to target_nnz exactly (when 'users <= target_nnz <= users*items'). */
int ycnr_synth_user_counts(uint64_t seed, int32_t users, int32_t items, int64_t target_nnz,
                           double alpha, int32_t* counts_out);

/* Fills item ids (ascending inside each user, unique per user) and integer
 * ratings 1..max_rating drawn from a rank-'rank' model plus noise.
 * user_ptr[users+1] is the exclusive prefix sum of the counts. */
int ycnr_synth_fill(uint64_t seed, int32_t users, int32_t items, const int64_t* user_ptr,
                    int32_t max_rating, int32_t rank, double item_skew,
                    int32_t* item_ids_out, float* ratings_out, int32_t nthreads);

/* N(mean, dev) factor init, element e of matrix 'which' (0 = user, 1 = item)
 * (EmfBase.js:486-497: randomNormal(1/k) per element). */
int ycnr_init_factors(uint64_t seed, int32_t which, int64_t count, double mean, double dev,
                      float* out, int32_t nthreads);

/* ---- split to train/validate/test (dataset_type 1/2/3) ------------------------ */
int ycnr_split_sets(uint64_t seed, int32_t users, const int64_t* user_ptr,
                    const int32_t pcts[3], int8_t* dataset_type_out, int32_t nthreads);

/* ---- portion planner ---------------------------------------------------------- */
/* cnt_per_row[total_rows]: ratings_count over sets 1,2,3 (0 = row absent).
 * pct_plus1: 0 for byUser/byItem, dataSetDistr[1]+1 for rmseValidate,
 * dataSetDistr[2]+1 for rmseTest.  portions_row_id_to receives the inclusive
 * 1-based upper row id of every portion (== exclusive 0-based end). */
int ycnr_split_to_portions(const int32_t* cnt_per_row, int32_t total_rows,
                           int32_t ratings_in_portion_opt, int32_t num_threads_opt,
                           int32_t pct_plus1,
                           int32_t* portions_row_id_to, int32_t cap, int32_t* n_portions_out,
                           int32_t* max_ratings_in_portion_out, int32_t* max_rows_in_portion_out);

/* ---- fetch filters: CSR by user / by item over a dataset_type mask ------------ */
/* set_mask: bit t set <=> dataset_type t is selected (train step: (1<<1)|(1<<2)). */
int ycnr_count_by_user(int32_t users, const int64_t* user_ptr, const int8_t* dataset_type,
                       uint32_t set_mask, int64_t* out_ptr /*[users+1]*/);
int ycnr_fill_by_user(int32_t users, const int64_t* user_ptr, const int32_t* item_ids,
                      const float* ratings, const int8_t* dataset_type, uint32_t set_mask,
                      const int64_t* out_ptr, int32_t* out_idx, float* out_vals);
int ycnr_count_by_item(int32_t users, int32_t items, const int64_t* user_ptr,
                       const int32_t* item_ids, const int8_t* dataset_type, uint32_t set_mask,
                       int64_t* out_ptr /*[items+1]*/);
/* columns (user ids) ascending inside each item row */
int ycnr_fill_by_item(int32_t users, int32_t items, const int64_t* user_ptr,
                      const int32_t* item_ids, const float* ratings, const int8_t* dataset_type,
                      uint32_t set_mask, const int64_t* out_ptr, int32_t* out_idx, float* out_vals);

/* ---- portion conversion (reference wire format, incl. the Q2 drop) ------------ */
/* Rows [row_from, row_to) of the CSR -> buf_rows = [R, rowId0, n0, ...],
 * buf_indx / buf_vals = ALL fetched ratings of the portion in order (the dropped
 * last rating stays in the buffer but is not covered by any row, as upstream).
 * fetched_out = number of ratings fetched (data.length). */
int ycnr_build_portion(const int64_t* ptr, const int32_t* idx, const float* vals,
                       int32_t row_from, int32_t row_to,
                       int32_t* buf_rows, int32_t cap_rows_words,
                       int32_t* buf_indx, float* buf_vals, int32_t cap_ratings,
                       int32_t* fetched_out);

/* Bulk form of the same conversion: the concatenation of all portion headers as a
 * row list into the CSR's own idx/vals arrays (no copy of the ratings).
 * portion_first[p] = index into the row list of portion p's first row,
 * portion_first[n_portions] = number of rows. */
int ycnr_build_rowlist(const int64_t* ptr, const int32_t* portions_row_id_to, int32_t n_portions,
                       int32_t* row_ids, int64_t* row_start, int32_t* row_len, int32_t cap_rows,
                       int32_t* portion_first /*[n_portions+1]*/);

#ifdef __cplusplus
}
#endif
#endif /* YCNR_HOST_H */
