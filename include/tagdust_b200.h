/* tagdust_b200.h -- C ABI of the B200-native TagDust2 per-read HMM decode path.
 *
 * This library replaces exactly one seam of the reference (TagDust2 v2.33):
 *
 *     int run_pHMM(struct arch_bag* ab, struct model_bag* mb, struct read_info** ri,
 *                  struct parameters* param, struct fasta* reference_fasta,
 *                  int numseq, int mode);            barcode_hmm.h:342, barcode_hmm.c:1895
 *
 * and the per-read kernels underneath it (backward :3439, forward_max_posterior_decoding
 * :4128, the Q score in do_label_thread :2269 / do_probability_estimation :2174,
 * extract_reads :3172, do_arch_comparison :2111).  Everything is plain C: pointers,
 * sizes, POD structs.  No torch types, no CPU fallback: every compute entry point
 * returns TDG_ENODEV when no sm_100 device/driver is usable.
 *
 * Return codes follow kslib.h:13-17 (kslOK 0, kslFAIL 1, kslEMEM 2) for the values
 * the reference's callers test; extra codes are >= 16.
 */
#ifndef TAGDUST_B200_H
#define TAGDUST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDG_OK      0   /* kslOK   kslib.h:13 */
#define TDG_FAIL    1   /* kslFAIL kslib.h:14 */
#define TDG_EMEM    2   /* kslEMEM kslib.h:15 */
#define TDG_ENODEV  16  /* no usable CUDA device / driver */
#define TDG_EINVAL  17  /* bad argument */
#define TDG_ECUDA   18  /* CUDA runtime error, see tdg_last_error() */

/* run modes: same numbering as barcode_hmm.h:128-132 */
#define TDG_MODE_GET_LABEL 1
#define TDG_MODE_GET_PROB  4
#define TDG_MODE_ARCH_COMP 5
/* not a run_pHMM mode: run_rna_dust() (barcode_hmm.c:2043, do_rna_dust :2370) for input files whose architecture is a
 * single R segment -- every read starts as EXTRACT_SUCCESS, then the -ref artifact filter, then dust; no model */
#define TDG_MODE_RNA_DUST  6

/* read_type codes: io.h:36-52 (the numbering every reference TU actually sees) */
#define TDG_EXTRACT_SUCCESS                    0
#define TDG_EXTRACT_FAIL_ARCHITECTURE_MISMATCH 1
#define TDG_EXTRACT_FAIL_READ_TOO_SHORT        2
#define TDG_EXTRACT_FAIL_BAR_FINGER_NOT_FOUND  3
#define TDG_EXTRACT_FAIL_MATCHES_ARTIFACTS     5
#define TDG_EXTRACT_FAIL_LOW_COMPLEXITY        6

#define TDG_LOGSUM_SIZE 16000     /* misc.h:45 */
#define TDG_MAX_SEGMENTS 16       /* reference allows 10, interface.c:840 */
#define TDG_MAX_HMMS 255          /* labels are bytes; reference: char labels + total_prob[100] */

/* transition indices, barcode_hmm.h:87-96 */
enum { TDG_MM = 0, TDG_MI, TDG_MD, TDG_II, TDG_IM, TDG_DD, TDG_DM, TDG_MSKIP, TDG_ISKIP };

typedef struct tdg_context tdg_context;
typedef struct tdg_model   tdg_model;
typedef struct tdg_batch   tdg_batch;
typedef struct tdg_refset  tdg_refset;   /* -ref: the artifact sequences (struct fasta, io.h:59-69) on the devices */

/* ------------------------------------------------------------------------------------
 * Flattened `struct model_bag` (barcode_hmm.h:247-272).  Order everywhere is
 * segment -> hmm -> column.  H = total_hmms, C = total_columns = sum_s num_hmms*num_cols.
 * All floats are the log-space values the reference stores (float of a double log()).
 * ---------------------------------------------------------------------------------- */
typedef struct tdg_model_desc {
	int32_t num_segments;        /* mb->num_models */
	int32_t total_hmms;          /* mb->total_hmm_num */
	int32_t total_columns;
	int32_t average_raw_length;  /* mb->average_raw_length (random model, :4516-4523) */
	const char*    seg_type;     /* [S]  param->read_structure->type[] ('B','F','S','P','O','G','R') */
	const int32_t* seg_num_hmms; /* [S]  model[s]->num_hmms */
	const int32_t* seg_num_cols; /* [S]  model[s]->hmms[0]->num_columns */
	const float*   seg_skip;     /* [S]  model[s]->skip */
	const float*   background;   /* [5]  model[0]->background_nuc_frequency */
	const float*   transition;   /* [C*9] hmm_column.transition */
	const float*   m_emit;       /* [C*5] */
	const float*   i_emit;       /* [C*5] */
	const float*   silent_to_M;  /* [C]   model[s]->silent_to_M[f][g] */
	const float*   silent_to_I;  /* [C] */
	const int32_t* label;        /* [H]   mb->label */
	const float*   transition_matrix; /* [H*H] mb->transition_matrix, entries 0/1 */
} tdg_model_desc;

/* per-call knobs read from `struct parameters` by run_pHMM's callees */
typedef struct tdg_run_params {
	float   confidence_threshold; /* param->confidence_threshold  (extract_reads :3204) */
	int32_t minlen;               /* param->minlen */
	int32_t matchstart;           /* param->matchstart, -1 = unset */
	int32_t matchend;             /* param->matchend,   -1 = unset */
	int32_t dust;                 /* param->dust; 0 = off (dust_sequences :2407) */
	int32_t want_labels;          /* 1: ri->labels rows are returned (tdg_result.labels).  0: MODE_GET_PROB skips the label
	                                 DP/traceback altogether; MODE_GET_LABEL still runs it (extraction needs it) but keeps the
	                                 rows on the device -- most of the device->host bytes of a read */
	int32_t want_spans;           /* 1 (MODE_GET_LABEL): return the R-labelled runs of every extracted read (tdg_result.spans),
	                                 i.e. what make_extracted_read (barcode_hmm.c:3325-3356) leaves of it */
	/* -ref artifact filter, match_to_reference (barcode_hmm.c:2478-2583); order extract -> artifacts -> dust (:2345-2354).
	 * A read within filter_error edits of a reference sequence (either strand) gets read_type = (sequence number << 8) | 5.
	 * The batch stands for ONE run_pHMM / run_rna_dust call: the reference matches the reads in groups of four per thread
	 * slice of that call with one Myers variant and the remaining (slice length mod 4) reads with another, so the slicing
	 * (slice_threads = param->num_threads over the batch's reads) is part of the result. */
	const tdg_refset* refset;     /* NULL: no artifact filter */
	int32_t filter_error;         /* param->filter_error (-fe, default 2) */
	int32_t slice_threads;        /* param->num_threads */
} tdg_run_params;

/* Per-read results (host arrays owned by the batch; valid after tdg_wait/tdg_run). */
typedef struct tdg_result {
	int32_t        n_reads;
	int32_t        label_stride;  /* bytes per read in `labels` */
	const float*   mapq;          /* ri->mapq */
	const float*   bar_prob;      /* ri->bar_prob before it is overwritten with 100 (:2343) */
	const float*   f_score;       /* mb->f_score */
	const float*   b_score;       /* mb->b_score */
	const float*   r_score;       /* mb->r_score */
	const int32_t* read_type;     /* ri->read_type after extract_reads AND dust_sequences */
	const uint8_t* extracted;     /* 1 where extract_reads succeeded (make_extracted_read ran, :3325),
	                                 even if dust later overwrote read_type */
	const int32_t* barcode;       /* ri->barcode   (-1 if not set) */
	const int32_t* fingerprint;   /* ri->fingerprint (-1 if not set) */
	const uint8_t* labels;        /* ri->labels[0..len] at labels + r*label_stride; NULL unless want_labels */
	int32_t        span_stride;   /* (start, len) pairs per read in `spans` = R segments of the architecture + 1 */
	const uint16_t* spans;        /* want_spans: spans[(r*span_stride + k)*2] = first residue (0-based), [.. + 1] = length of the
	                                 k-th run of residues whose label lies in an R segment; len 0 ends the list.  Only filled for
	                                 reads with extracted == 1; NULL unless want_spans */
} tdg_result;

/* ---- lifetime ------------------------------------------------------------------- */
/* n_devices <= 0: use every visible device.  device_ids may be NULL (0..n-1). */
int  tdg_init(int n_devices, const int* device_ids, tdg_context** out);
void tdg_shutdown(tdg_context* ctx);
int  tdg_device_count(const tdg_context* ctx);
const char* tdg_last_error(void);
const char* tdg_version(void);

/* host-side numerics shared with the reference (misc.c:57-105); usable without a GPU */
void  tdg_logsum_table(float* out16000);      /* init_logsum */
float tdg_logsum_host(float a, float b);      /* logsum */

/* ---- model ------------------------------------------------------------------------ */
/* Copies the description, derives the dead-term masks and the label-DP source lists,
 * uploads to every device of the context.  max_len = longest read that will be
 * submitted (mb->current_dyn_length - 10). */
int  tdg_model_create(tdg_context* ctx, const tdg_model_desc* desc, int max_len, tdg_model** out);
void tdg_model_destroy(tdg_model* m);
/* Host-only validation/derivation (no GPU needed): writes the number of HMM columns
 * whose terms are statically dead etc.  Used by CPU tests. */
int  tdg_model_validate(const tdg_model_desc* desc, char* errbuf, size_t errbuf_len);

/* Model compiler: the host-side mirror of init_model_bag (barcode_hmm.c:5760-6011).
 * Builds the flat description from the user's segment strings ("B:ACGT,TTGA", "R:N", ...)
 * exactly as interface.c:489-598 + barcode_hmm.c:4689-5084,1710-1881 would.
 * background_logp[5] = ssi->background, the other scalars = the matching ssi fields.
 * The returned handle owns the arrays `desc` points into. */
typedef struct tdg_arch tdg_arch;
typedef struct tdg_arch_params {
	double background_logp[5];
	double average_length;
	int32_t max_seq_len;
	double expected_5_len, mean_5_len, stdev_5_len;
	double expected_3_len, mean_3_len, stdev_3_len;
	float  sequencer_error_rate;   /* param->sequencer_error_rate (-e, default 0.05) */
	float  indel_frequency;        /* param->indel_frequency (-i, default 0.1) */
	int32_t calibration_edit;      /* 1: zero the N-alternative priors, calibrateQ.c:67-86 */
} tdg_arch_params;
int  tdg_arch_compile(int num_segments, const char* const* segment_strings,
                      const tdg_arch_params* p, tdg_arch** out);
const tdg_model_desc* tdg_arch_desc(const tdg_arch* a);
void tdg_arch_destroy(tdg_arch* a);

/* ---- -ref artifact sequences ---------------------------------------------------------- */
/* codes = fasta->string (nuc_code[] of every sequence back to back), s_index[numseq + 1] = fasta->s_index
 * (get_fasta, io.c).  Copied to every device of the context. */
int  tdg_refset_create(tdg_context* ctx, const uint8_t* codes, const int32_t* s_index, int numseq, tdg_refset** out);
void tdg_refset_destroy(tdg_refset* r);

/* ---- read batches (pinned, structure-of-arrays, 4-bit packed) ---------------------- */
/* A batch owns pinned host staging (packed codes, lengths, results) and the matching
 * device buffers on the device(s) it is sharded over.  Two batches = double buffering. */
int  tdg_batch_create(tdg_context* ctx, int max_reads, int max_len, tdg_batch** out);
void tdg_batch_destroy(tdg_batch* b);
int  tdg_batch_clear(tdg_batch* b);
/* Append reads given as one byte code (0..4) per base, rows `stride` bytes apart.
 * The byte after the last base (codes[len]) is packed too: backward() reads it
 * (barcode_hmm.c:3516).  Pass NULL `terminators` to use 0 like the reference's NUL. */
int  tdg_batch_append_codes(tdg_batch* b, int n, const uint8_t* codes, size_t stride, const int32_t* len);
/* Append straight from an array of record pointers (`struct read_info**`):
 * seq pointer at byte offset seq_off, int length at len_off inside each record. */
int  tdg_batch_append_records(tdg_batch* b, int n, const void* const* records,
                              size_t seq_off, size_t len_off);
int  tdg_batch_size(const tdg_batch* b);

/* Host-only: how a batch of n_reads is split over n_devices (contiguous, 32-read-tile aligned,
 * keeps output order = input order like the reference's static slices, barcode_hmm.c:1911-1922). */
int  tdg_plan_shards(int n_reads, int n_devices, int32_t* first, int32_t* count);

/* ---- the hot path ----------------------------------------------------------------- */
/* Asynchronous: H2D copies, kernels and D2H copies are queued on the batch's streams
 * (reads sharded contiguously over the context's devices); returns immediately. */
int  tdg_submit(tdg_context* ctx, tdg_model* m, int mode, const tdg_run_params* p, tdg_batch* b);
/* Blocks until the batch's results are on the host; fills *out (pointers into b). */
int  tdg_wait(tdg_batch* b, tdg_result* out);
/* tdg_submit + tdg_wait */
int  tdg_run(tdg_context* ctx, tdg_model* m, int mode, const tdg_run_params* p, tdg_batch* b, tdg_result* out);

/* MODE_ARCH_COMP (do_arch_comparison :2111 + merge :1994-2017): backward() of every
 * read under each of the A models.  b_scores[a*n + r] (host, A*n floats, may be NULL)
 * and arch_posterior[a] (host, A floats) = per-architecture float sums accumulated in
 * read order over `num_threads` slices like the reference, added in slice order,
 * then log-normalised. */
int  tdg_arch_compare(tdg_context* ctx, tdg_model* const* models, int num_arch,
                      tdg_batch* b, int num_threads, float* b_scores, float* arch_posterior);

/* ---- device-resident entry point (benchmark `value`: inputs already in HBM) -------- */
/* Uploads the batch once; afterwards tdg_decode_resident() runs only the kernels on
 * `cuda_stream` (a cudaStream_t, may be 0) of device `device_index`. *n_launches
 * receives the number of kernel launches queued. */
int  tdg_batch_upload(tdg_context* ctx, tdg_batch* b);
int  tdg_decode_resident(tdg_context* ctx, tdg_model* m, int mode, const tdg_run_params* p,
                         tdg_batch* b, void* cuda_stream, int* n_launches);
int  tdg_batch_download(tdg_batch* b, tdg_result* out);

/* Per-kernel device timing (CUDA events around every launch on the launching stream).
 * kinds: 0 k_backward, 1 k_forward, 2 k_label.  Read after synchronising. */
int  tdg_profile_enable(tdg_context* ctx, int on);
int  tdg_profile_read(tdg_context* ctx, int device_index, float ms[3], int launches[3]);

/* Work the kernels execute per read position, summed over all HMMs of the model (host-only, no GPU needed):
 * out[0] logsums and out[1] float adds of the backward pass, out[2] / out[3] the same for forward + posterior.
 * Terms whose transition is log(0) are never evaluated (logsum(x, -inf) == x), so these "live" counts are lower
 * than SURVEY 8d's algorithmic 8 + 10 logsums per (column, position); bench.py's roofline.frac uses them. */
int  tdg_desc_live_ops(const tdg_model_desc* desc, double out[4]);

/* work accounting for the roofline: profile-column cells (2*L*C per read) */
double tdg_batch_cells(const tdg_model* m, const tdg_batch* b);

#ifdef __cplusplus
}
#endif
#endif /* TAGDUST_B200_H */
