/*
 * ycnr_als.h — C ABI of the B200 (sm_100a) explicit-ALS factor-update library
 * (libycnr_als.so).  Plain pointers and sizes only; this is exactly the surface the
 * reference's native addon slot binds (N-API shim: you_can_not_recommend_b200/napi/,
 * ctypes: you_can_not_recommend_b200/native.py).  There is NO CPU fallback: every
 * compute entry point fails with a non-zero status when no CUDA device is usable.
 *
 * What each entry point replaces in the reference (file:line into upstream):
 *
 *   ycnr_create / ycnr_destroy ........ EmfWorker.mw_prepareToTrain / mw_endTrain
 *                                       (lib/emf/EmfWorker.js:119-130,160-164), options
 *                                       lib/emf/EmfBase.js:52-140
 *   ycnr_attach_factors ............... EmfBase.openSharedFactors (EmfBase.js:430-450):
 *                                       the row-major Float32 user/item matrices in SysV shm
 *   ycnr_host_register ................ the worker's portion buffers are fixed shm segments
 *                                       (EmfWorker.openWorkPortionBuffers, EmfWorker.js:66-89)
 *   ycnr_start_train_step ............. EmfWorker.mw_startTrainStep (EmfWorker.js:135-138)
 *   ycnr_als_portion .................. EmfWorker.mw_calcTrainAlsPortion (EmfWorker.js:169-261)
 *                                       incl. EmfBase.copySubFixedFactors (EmfBase.js:537-555),
 *                                       BLAS.gemm / Matrix.add / multiply / solveSquare / transpose
 *                                       (EmfWorker.js:231-247) and the in-place row write
 *                                       (EmfBase.js:518-532)
 *   ycnr_end_train_step ............... the 'stepComplete' barrier (EmfMaster.js:776-785)
 *   ycnr_start_calc_rmse .............. EmfWorker.mw_startCalcRmse (EmfWorker.js:144-148)
 *   ycnr_rmse_portion ................. EmfWorker.mw_calcRmsePortion (EmfWorker.js:266-315)
 *                                       incl. predictSync/_alsPredict (EmfBase.js:785-827)
 *   ycnr_s_als_build_sub_fixed_facts .. sAlsBuildSubFixedFacts (cpp_utils/als_utils.cc:22-38,
 *                                       registered cpp_utils/cpp_utils.cc:3-8)
 *   ycnr_rowset_* / ycnr_als_rowset / ycnr_rmse_rowset
 *                                       bulk form of the same two portion calls: all portions
 *                                       of a step resident on the device
 *   ycnr_device_factors, ycnr_ipc_*, ycnr_set_peers
 *                                       replica refresh after a half-step, replacing
 *                                       'alsSaveCalcedFactors' streams (EmfMaster.js:711-723)
 *
 * Conventions: ids 0-based; factor matrices row-major [rows x factors_count] float32
 * (EmfBase.js:403-412); every function returns 0 on success, otherwise a non-zero
 * code with a message in ycnr_last_error() (the N-API shim turns it into a thrown
 * Error so the worker's uncaughtException path fires, EmfWorkerProcess.js:30-45).
 * Calls are blocking with respect to their INPUT buffers (the caller may reuse
 * them on return) and one call is in flight per context, like the reference worker.
 */
#ifndef YCNR_ALS_H
#define YCNR_ALS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ycnr_ctx ycnr_ctx;

/* stepType (EmfWorker.js:135-148) */
enum {
  YCNR_BY_USER = 0,       /* solve user rows, item factors fixed */
  YCNR_BY_ITEM = 1,       /* solve item rows, user factors fixed */
  YCNR_RMSE_VALIDATE = 2,
  YCNR_RMSE_TEST = 3
};

/* which matrix */
enum { YCNR_USER_FACTORS = 0, YCNR_ITEM_FACTORS = 1 };

/* Gram-build code path (north star: tensor cores via 3xTF32 or FP32 FFMA) */
enum { YCNR_GRAM_AUTO = 0, YCNR_GRAM_FFMA = 1, YCNR_GRAM_TC3XTF32 = 2 };

typedef struct ycnr_options {
  int32_t factors_count;        /* options.factorsCount (EmfBase.js:90) */
  int32_t total_users;          /* user matrix height = max(list_id) (EmfLord.js:81) */
  int32_t total_items;          /* item matrix height = max(id) (EmfLord.js:82) */
  double user_fact_reg;         /* options.als.userFactReg (EmfBase.js:67) */
  double item_fact_reg;         /* options.als.itemFactReg (EmfBase.js:69) */
  int32_t use_double_precision; /* options.useDoublePrecision: must be 0 (rejected otherwise) */
  int32_t lowmem;               /* options.lowmem: must be 0 (rejected otherwise) */
  int32_t device;               /* CUDA device ordinal */
  int32_t gram_path;            /* YCNR_GRAM_* */
  int32_t dual_max_cols;        /* rows with <= this many ratings solve the n x n dual system;
                                   -1 = library default, 0 = never */
  int32_t split_cols;           /* ratings per partial-Gram work item for long rows; 0 = default */
  int32_t profile;              /* record CUDA events around every kernel class */
  int32_t tc_min_cols;          /* TC path: rows with at least this many ratings (and more than
                                   dual_max_cols) take the tensor-core Gram; 0 = all of them */
  int32_t tc_variant;           /* diagnostics only, keep 0 (16: raw fp32 in the TF32 head columns) */
  int32_t solve_chunks;         /* > 1: cut the split rows of a half-step into this many groups (<= 8) and run the
                                   reduce+solve of group i on a second stream under the Gram of group i+1
                                   (measured: no gain on B200, default 0 = stream order) */
  int32_t reserved[1];
} ycnr_options;

/* 'completedPortion' message fields (EmfWorker.js:254-260, 304-314) */
typedef struct ycnr_portion_info {
  int32_t rows_from;            /* rowsRange.from = first row id of the portion (-1 if empty) */
  int32_t rows_cnt;             /* rowsRange.cnt  = rowsInPortion */
  int64_t ratings_in_portion;   /* sum of cols */
  double time_ms;               /* host wall time spent inside the call */
  double r_sum_diff2;           /* RMSE only */
  double r_cnt;
  double r_sum;
} ycnr_portion_info;

/* accumulated device time per kernel class since the last ycnr_profile_reset */
enum {
  YCNR_K_PRIMAL_FUSED = 0,   /* gather + Gram + RHS + ridge + Cholesky solve, one CTA per row */
  YCNR_K_DUAL_FUSED = 1,     /* gather + n x n Gram + Cholesky solve + Y^T z */
  YCNR_K_GRAM_PARTIAL = 2,   /* gather + partial Gram/RHS of one slice of a long row */
  YCNR_K_REDUCE_SOLVE = 3,   /* sum partials + ridge + Cholesky solve */
  YCNR_K_RMSE_ROWS = 4,
  YCNR_K_RMSE_REDUCE = 5,
  YCNR_K_GATHER = 6,            /* data-movement kernels: row gather, header unpack, recommend */
  YCNR_K_GRAM_TC = 7,        /* tcgen05 3xTF32 Gram */
  YCNR_K_CLASSES = 8
};
typedef struct ycnr_profile {
  double ms[YCNR_K_CLASSES];
  int64_t launches[YCNR_K_CLASSES];
  int64_t rows[YCNR_K_CLASSES];      /* rows (or work items) processed */
  int64_t ratings[YCNR_K_CLASSES];   /* ratings gathered */
  int64_t total_launches;            /* every kernel launched by this context, profiled or not */
} ycnr_profile;

const char* ycnr_last_error(void);
int ycnr_device_count(int32_t* count_out);

int ycnr_create(const ycnr_options* opts, ycnr_ctx** ctx_out);
int ycnr_destroy(ycnr_ctx* ctx);

/* ---- factor store --------------------------------------------------------- */
/* Borrow the two host matrices (pointers into the shm segments), page-lock them and
 * upload both to the device replicas.  The pointers stay owned by the caller and must
 * outlive the context (or until the next attach). */
int ycnr_attach_factors(ycnr_ctx* ctx, float* user_factors, float* item_factors);
int ycnr_upload_factors(ycnr_ctx* ctx, int32_t which);                     /* host -> device */
int ycnr_download_factors(ycnr_ctx* ctx, int32_t which, int32_t row_from, int32_t row_cnt); /* device -> host */
/* Tell the library another process changed the host copy (next step re-uploads it). */
int ycnr_invalidate_device(ycnr_ctx* ctx, int32_t which);
/* Raw device pointer of a replica (for the NCCL all-gather done by the host layer). */
int ycnr_device_factors(ycnr_ctx* ctx, int32_t which, void** dptr_out);
int ycnr_stream(ycnr_ctx* ctx, void** cuda_stream_out);
int ycnr_synchronize(ycnr_ctx* ctx);

/* Page-lock caller memory that holds CACHED portion buffers (usePortionsCache: the master keeps converted
 * portions, EmfMaster.js:434-494, 656-658; segments of EmfMaster.createWorkPortionBuffers, EmfMaster.js:156-234).
 * Portions whose header / indx / vals lie inside a registered region are DMA'd straight from it,
 * ASYNCHRONOUSLY: such memory must stay unmodified until the step ends (ycnr_end_train_step, or the return of
 * ycnr_rmse_portion).  Buffers that are refilled between portions (the upstream work buffer) must NOT be
 * registered: unregistered buffers are copied before the call returns, as upstream expects. */
int ycnr_host_register(ycnr_ctx* ctx, void* ptr, size_t bytes);
int ycnr_host_unregister(ycnr_ctx* ctx, void* ptr);

/* ---- drop-in per-portion path (reference wire format) ------ */
int ycnr_start_train_step(ycnr_ctx* ctx, int32_t step_type);
int ycnr_als_portion(ycnr_ctx* ctx, const int32_t* als_rows, const int32_t* als_indx,
                     const float* als_vals, ycnr_portion_info* info_out);
/* Barrier: waits for every queued portion and makes the solved rows visible in the
 * attached host matrix. */
int ycnr_end_train_step(ycnr_ctx* ctx);

int ycnr_start_calc_rmse(ycnr_ctx* ctx, int32_t step_type, double global_avg_shift);
int ycnr_rmse_portion(ycnr_ctx* ctx, const int32_t* rmse_rows, const int32_t* rmse_indx,
                      const float* rmse_vals, ycnr_portion_info* info_out);

/* Portions (the reference default is 10 000 ratings, EmfBase.js:97-103) are QUEUED by ycnr_als_portion and launched
 * in batches of ~4 M ratings (or at ycnr_end_train_step): the call checks the header, returns the 'completedPortion'
 * fields at once, the rows are in the host segment after ycnr_end_train_step as before.  Exceptions: a portion of
 * 2^20 ratings or more in UNREGISTERED buffers (it would have to be staged) and portions of 2^24 rows take the
 * single-portion path at once.
 * The RMSE twin: ycnr_rmse_portion_async queues a portion under a caller tag (portionNo); ycnr_rmse_poll hands
 * back the completed portions in the order they were queued — wait != 0 flushes the queue and waits for all of
 * them (the worker's 'completedPortion' replies are asynchronous messages upstream as well, EmfWorker.js:304-314).
 * ycnr_rmse_portion itself stays synchronous. */
int ycnr_rmse_portion_async(ycnr_ctx* ctx, const int32_t* rmse_rows, const int32_t* rmse_indx, const float* rmse_vals,
                            int64_t tag);
int ycnr_rmse_poll(ycnr_ctx* ctx, int32_t wait, int32_t max_out, int64_t* tags_out, ycnr_portion_info* infos_out,
                   int32_t* n_out);
/* n per-portion calls issued from native code: the same entry points, without the binding's per-call cost. */
int ycnr_als_portions(ycnr_ctx* ctx, int32_t n, const int32_t* const* rows, const int32_t* const* indx,
                      const float* const* vals, ycnr_portion_info* infos_out);
int ycnr_rmse_portions_async(ycnr_ctx* ctx, int32_t n, const int32_t* const* rows, const int32_t* const* indx,
                             const float* const* vals, const int64_t* tags);

/* cpp_utils.sAlsBuildSubFixedFacts(sub, fixed, indx, cols, k): sub[c,:] = fixed[indx[c],:].
 * 'fixed' is a host matrix of fixed_rows x k floats. */
int ycnr_s_als_build_sub_fixed_facts(ycnr_ctx* ctx, float* sub, const float* fixed, int64_t fixed_rows,
                                     const int32_t* indx, int32_t cols, int32_t k);

/* The upstream export's own arity, (sub, fixed, indx, cols, k) (cpp_utils/cpp_utils.js:15-19) + the row count the
 * binding knows from the typed array: uses the most recently created live context of the process. */
int ycnr_s_als_build_sub_fixed_facts_noctx(float* sub, const float* fixed, int64_t fixed_rows, const int32_t* indx,
                                           int32_t cols, int32_t k);

/* Host-only check of a portion header against the lengths of the arrays that carry it (the typed arrays of the
 * N-API binding, EmfWorker.js:176-219 reads them unchecked): rows_len words of alsRows/rmseRows, indx_len /
 * vals_len entries.  Row and column ids are checked by the portion calls themselves (rows on the host, columns
 * on the device). */
int ycnr_check_portion(const int32_t* rows, int64_t rows_len, int64_t indx_len, int64_t vals_len);
/* total_users * factors_count (which = 0) or total_items * factors_count: what attachFactors must be given */
int ycnr_factor_elems(ycnr_ctx* ctx, int32_t which, int64_t* elems_out);
/* 'getMemoryUsage' (EmfWorker.js:43,109-113; EmfBase.js:880-934): out[4] = device bytes held by the context,
 * page-locked host bytes (slots + registered regions), free and total device memory. */
int ycnr_memory_usage(ycnr_ctx* ctx, int64_t out[4]);

/* ---- bulk path: all portions of a step resident on the device ------------- */
/* A row set is the concatenation of the portion headers of one step: row r covers
 * indx/vals[row_start[r] .. row_start[r]+row_len[r]).  span = number of entries of
 * indx/vals to upload.  portion_first[n_portions+1] indexes the row list (RMSE only
 * needs it, for the per-portion sums of quirk Q7); may be NULL with n_portions = 0. */
int ycnr_rowset_create(ycnr_ctx* ctx, int32_t step_type, int32_t n_rows, const int32_t* row_ids,
                       const int64_t* row_start, const int32_t* row_len,
                       const int32_t* indx, const float* vals, int64_t span,
                       const int32_t* portion_first, int32_t n_portions, int32_t* rowset_out);
int ycnr_rowset_destroy(ycnr_ctx* ctx, int32_t rowset);
/* One half-step over the row set, asynchronous on the context stream; device replicas
 * only (use ycnr_download_factors for the host copy). */
int ycnr_als_rowset(ycnr_ctx* ctx, int32_t rowset);
/* Queue the RMSE pass of a row set without waiting for it (a master that knows the next passes — validate and
 * test with shift 0, EmfLord.js:896-897 — starts them together); ycnr_rmse_rowset then only waits.  The sums of
 * the last pass are kept per row set together with the sum of the ratings, so a pass with ANOTHER shift over
 * unchanged factors (EmfLord.js:898) is derived on the host: sum (r-p-d)^2 = sum (r-p)^2 - 2d (sum r - sum p) + n d^2. */
int ycnr_rmse_rowset_begin(ycnr_ctx* ctx, int32_t rowset, double global_avg_shift);
/* Sum of the ratings of the row set / of its last portion, from the last pass (see ycnr_rmse_rowset_begin). */
int ycnr_rmse_rowset_ratings(ycnr_ctx* ctx, int32_t rowset, double* total_out, double* last_portion_out);
/* totals[3] = {rSumDiff2, rCnt, rSum}; portion_sums[n_portions*3] optional. Synchronous. */
int ycnr_rmse_rowset(ycnr_ctx* ctx, int32_t rowset, double global_avg_shift, double* totals,
                     double* portion_sums);

/* ---- device-side front end --------------------------------- */
/* The ratings table malrec_ratings (data/db-schema.sql:887-893) sorted by (user, item) as flat arrays:
 * user_ptr[total_users+1], then per rating the 0-based item id, the rating and its dataset_type
 * (EmfBase.js:229-247).  Uploaded once; the calls below replace the per-portion SQL fetch and the per-rating
 * conversion loop of the master (EmfMaster.js:501-614) for whole steps. */
int ycnr_table_upload(ycnr_ctx* ctx, const int64_t* user_ptr, const int32_t* item_ids, const float* ratings,
                      const int8_t* dataset_type);
/* First-time split of every user's ratings into train(1) / validate(2) / test(3) on the device, rule and order of
 * EmfLord.doSplitToSets' JS path (EmfLord.js:450-473) with Math.random() replaced by the library's counter-based
 * PRNG (seed, user, step) — the same bytes as ycnr_split_sets of the host front end (include/ycnr_host.h).
 * pcts = dataSetDistr.  Rewrites the uploaded table's dataset_type column; dataset_type_out (may be NULL)
 * receives it. */
int ycnr_table_split(ycnr_ctx* ctx, uint64_t seed, const int32_t pcts[3], int8_t* dataset_type_out);
/* Ratings per user (by_item = 0) or per item (by_item = 1) whose dataset_type bit is set in set_mask —
 * the planner's ratings_count (EmfLord.js:48-128, 255-397).  counts_out[total_users | total_items]. */
int ycnr_table_counts(ycnr_ctx* ctx, uint32_t set_mask, int32_t by_item, int32_t* counts_out);
/* The portions of a step as a device-resident row set, built on the device: the fetch (ratings of the set
 * grouped by user in item order, or by item in user order), then the concatenated portion headers for the
 * plan portions_row_id_to[n_portions] (exclusive 0-based upper row bounds, EmfLord.js:510-612; the first
 * portion starts at first_row — 0 for a whole step, the previous bound for a rank's slice of the plan) with
 * the conversion loop's quirk (the last rating of every portion is dropped, EmfMaster.js:582-609).  Same
 * rows, bit for bit, as ycnr_rowset_create on the host front end's arrays (row_start addresses the whole
 * step's fetch). */
int ycnr_rowset_from_table(ycnr_ctx* ctx, int32_t step_type, uint32_t set_mask, int32_t first_row,
                           const int32_t* portions_row_id_to, int32_t n_portions, int32_t* rowset_out);
int ycnr_rowset_info(ycnr_ctx* ctx, int32_t rowset, int32_t* n_rows, int64_t* span, int32_t* n_portions);
/* Copy a row set's arrays back (any pointer may be NULL): row_ids/row_start/row_len[n_rows],
 * portion_first[n_portions+1], indx/vals[span]. */
int ycnr_rowset_read(ycnr_ctx* ctx, int32_t rowset, int32_t* row_ids, int64_t* row_start, int32_t* row_len,
                     int32_t* portion_first, int32_t* indx, float* vals);

/* ---- multi-GPU replica refresh over NVLink peer memory --------------------- */
/* 64-byte CUDA IPC handle of a device replica; import returns a peer-mapped pointer. */
int ycnr_ipc_export(ycnr_ctx* ctx, int32_t which, uint8_t handle_out[64]);
int ycnr_ipc_import(ycnr_ctx* ctx, const uint8_t handle[64], void** dptr_out);
int ycnr_ipc_close(ycnr_ctx* ctx, void* dptr);
/* Peer replicas of matrix 'which': the solve kernels store every solved row into each of
 * them as well (fused all-gather).  n_peers = 0 clears. Max 7 peers. */
int ycnr_set_peers(ycnr_ctx* ctx, int32_t which, int32_t n_peers, void* const* peer_dptrs);

/* ---- serving: top-N recommendation ------------------------ */
/* YcnrController.recommendItemsForUser (lib/YcnrController.js:227-284) for a batch of users, from the device
 * replicas: every item that is not in the user's skip list (rated + "unrated" items, 244-251; 0-based ids,
 * skip_ptr[n_users+1] indexes skip_ids) gets predict = fp32 dot(U[u], V[i]) + global_avg_shift; items with
 * predict >= min_recommend_rating compete, best first, ties: lower item id first.  Upstream pops the last entry
 * whenever its list reaches `limit` (281-282), so at most limit-1 items come back: out_item_ids / out_predict
 * are [n_users][limit-1] (0-based item ids), out_count[n_users].  Synchronous. */
int ycnr_recommend_batch(ycnr_ctx* ctx, int32_t n_users, const int32_t* user_ids, const int64_t* skip_ptr,
                         const int32_t* skip_ids, int32_t limit, double min_recommend_rating,
                         double global_avg_shift, int32_t* out_item_ids, double* out_predict, int32_t* out_count);

/* ---- diagnostics ------------------------------------------------------------ */
/* The launch plan the library builds for a row list (host code only, no GPU needed): rows by kernel class.
 * summary[32]: [0..23] rows per dual bin (tile-row count 1..24), [24] fused rows, [25] split rows, [26] work items
 * (slices) of the split rows, [27..31] word offsets of the fused list, the split-row list, item_row, item_off and
 * item_order inside the packed plan; words_out (may be NULL) receives the packed plan (n_words_out words). */
int ycnr_debug_plan(const ycnr_options* opts, const int32_t* row_len, int32_t n_rows, int32_t* summary,
                    int32_t* words_out, int64_t cap_words, int64_t* n_words_out);
/* The row arrays of a batch of portion headers as the multi-portion entry points build them (host code only, no GPU
 * needed): counts_out[3 n] = entries, ratings, bad flag per portion; ids/len/start receive the concatenated rows
 * (kind 1) or RMSE work entries of at most 64 ratings (kind 2) of the valid portions; threads = host workers. */
int ycnr_debug_batch_rows(int32_t kind, int32_t n, const int32_t* const* rows, int64_t lim_rows, int32_t threads,
                          int64_t* counts_out, int32_t* ids_out, int32_t* len_out, int64_t* start_out, int64_t cap,
                          int64_t* entries_out);
/* Copy the tile partials ([items][tiles][16] floats) left by the last split-row launch. */
int ycnr_debug_read_partials(ycnr_ctx* ctx, float* out, int64_t n_floats);

/* ---- measurement ------------------------------------------------------------ */
int ycnr_profile_reset(ycnr_ctx* ctx);
int ycnr_profile_read(ycnr_ctx* ctx, ycnr_profile* out);   /* synchronises the stream */
/* YCNR_K_DUAL_FUSED broken down by dual bin (tile-row count 1..24): device ms and rows since the last reset */
int ycnr_profile_dual_bins(ycnr_ctx* ctx, double ms_out[24], int64_t rows_out[24]);

#ifdef __cplusplus
}
#endif
#endif /* YCNR_ALS_H */
