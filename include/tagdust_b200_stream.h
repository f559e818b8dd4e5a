/* tagdust_b200_stream.h -- C ABI of the streaming demultiplexer: the callers and data formats
 * either side of the HMM decode path.
 *
 * Replaces, for FASTQ/FASTA input, the body of the labelling loop of
 *
 *     int hmm_controller_multiple(struct parameters* param)      barcode_hmm.c:51, loop :243-384
 *
 * i.e. per chunk: read_fasta_fastq (io.c:1684-1815) on every input file, run_pHMM
 * (barcode_hmm.c:1895) or run_rna_dust (:2043, :2370) per file, the cross-file merge of
 * read_type / barcode (:329-351), print_all (io.c:757-1016) and the log tallies (:356-384).
 * Output files are byte-identical to the reference's (names, `@name;FP:<int>;RQ:%0.2f`
 * headers, spacer-split multi-read records, the `_un` files, empty files for unused barcodes).
 *
 * What makes it fast where the reference is slow: block reads + one pass line splitting instead
 * of fgets and four mallocs per read, code conversion / 4-bit packing / record formatting on a
 * pool of host threads, output files opened once, and three pipeline stages (parse, GPU, write)
 * running concurrently on different chunks with double-buffered pinned batches.
 * SAM/BAM input is NOT handled here (the caller keeps the reference's own loop for it).
 */
#ifndef TAGDUST_B200_STREAM_H
#define TAGDUST_B200_STREAM_H

#include "tagdust_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TDG_EIO     19  /* cannot open / read / write a file */
#define TDG_EFORMAT 20  /* malformed input (io.c:1770 "Length of sequence and base qualities differ", unequal files) */

/* ---- FASTQ / FASTA reader (read_fasta_fastq, io.c:1684-1815; io_handler, io.c:382-608) ---- */
typedef struct tdg_fastq tdg_fastq;

/* One chunk of parsed reads; all pointers are owned by the reader and stay valid until the
 * next tdg_fastq_next / tdg_fastq_close on it. */
typedef struct tdg_fastq_chunk {
	int32_t         n;          /* reads in this chunk */
	int32_t         max_len;    /* longest read */
	const int32_t*  len;        /* [n]   ri->len */
	const uint64_t* seq_off;    /* [n]   offset of read r in `codes` / `qual` (len[r]+1 bytes each, 0-terminated) */
	const uint8_t*  codes;      /* ri->seq: nuc_code[] of every base (nuc_code.c:46-74), A,C,G,T/U = 0..3, '.' = 5, rest 4 */
	const uint8_t*  qual;       /* ri->qual bytes; NULL for FASTA input */
	const uint64_t* name_off;   /* [n+1] offsets into `names` */
	const char*     names;      /* ri->name: header line after '@' / '>' up to the first control character, 0-terminated */
} tdg_fastq_chunk;

/* path: plain file, or *.gz / *.bz2 (piped through zcat / bzcat like io_handler).
 * fasta: 1 = FASTA, 0 = FASTQ, -1 = decide from the suffix like io_handler (.fa/.fasta[.gz]). */
int  tdg_fastq_open(const char* path, int fasta, tdg_fastq** out);
/* Parses up to max_reads reads (param->num_query); chunk->n == 0 at end of input. */
int  tdg_fastq_next(tdg_fastq* f, int max_reads, int threads, tdg_fastq_chunk* chunk);
void tdg_fastq_close(tdg_fastq* f);

/* ---- sequence statistics (get_sequence_stats, io.c:52-300) ---------------------------------
 * The raw sums that function accumulates over the first reads of a file: chunks of num_query
 * reads until more than 1 000 000 have been seen (io.c:145-213).  All sums are sums of integers,
 * so the multi-threaded accumulation is exact; the caller derives ssi->average_length,
 * ssi->background[] and the 5'/3' partial-segment mean / stdev from them as io.c:216-270 does.
 * five / three: nuc_code[] of the P segment at the 5' / 3' end of the architecture, or NULL. */
typedef struct tdg_seq_stats {
	int64_t total_read;
	int32_t max_seq_len;
	double  sum_len;           /* ssi->average_length before the division */
	double  base_count[5];     /* ssi->background[] increments (without the initial 1.0) */
	double  five_s0, five_s1, five_s2;
	double  three_s0, three_s1, three_s2;
} tdg_seq_stats;
int  tdg_sequence_stats(const char* path, int fasta, int num_query,
                        const uint8_t* five, int five_len, const uint8_t* three, int three_len,
                        int threads, tdg_seq_stats* out);

/* Append a parsed chunk (ragged rows) to a batch, packing on `threads` host threads. */
int  tdg_batch_append_ragged(tdg_batch* b, int n, const uint8_t* codes, const uint64_t* seq_off,
                             const int32_t* len, int threads);

/* Label buffers (max_len + 1 bytes per read, pinned + device) are created by the first submit that asks for
 * labels; this creates them up front, e.g. on a set-up thread. */
int  tdg_batch_reserve_labels(tdg_batch* b);

/* `%0.2f` of a float exactly as fprintf prints ri->mapq (io.c:960-990); returns the length written. */
int  tdg_format_rq(float mapq, char* out);

/* ---- the demultiplexing job -------------------------------------------------------------- */
typedef struct tdg_demux_input {
	const char* path;              /* param->infile[i] */
	int32_t     fasta;             /* as tdg_fastq_open */
	tdg_model*  model;             /* flattened model_bag_container[i]; NULL when the file's architecture is a single
	                                  R segment (run_rna_dust path, barcode_hmm.c:312-318) */
	int32_t     num_read_segments; /* read_present[i]: R segments in the architecture = output reads of this file */
	float       confidence_threshold; /* param->confidence_thresholds[i] */
	int32_t     max_seq_len;       /* sequence_stats_info_container[i]->max_seq_len as the chunk loop sees it (threshold calibration
	                                  raises it to its longest simulated read): the "Long sequence found" count starts from it */
	int32_t     expected_len;      /* longest read get_sequence_stats saw in the file: sizes the staging batches (longer reads
	                                  later in the file still work, the batch of that slot is re-created); 0 = max_seq_len */
} tdg_demux_input;

typedef struct tdg_demux_job {
	int32_t n_inputs;                  /* param->infiles */
	const tdg_demux_input* inputs;
	int32_t barcode_input;             /* index of the input whose architecture holds a B segment, -1 = none (:329-341) */
	int32_t num_alternatives;          /* numseq_in_segment[first B segment] (barcodes + the N alternative); 2 without B */
	const char* const* barcode_names;  /* sequence_matrix[first B segment][0 .. num_alternatives-2]; NULL without B */
	const char* outfile;               /* param->outfile */
	int32_t minlen, dust;              /* param->minlen, param->dust */
	int32_t matchstart, matchend;      /* param->matchstart / matchend (-1 = unset) */
	int32_t print_seq_finger;          /* param->print_seq_finger (-show_finger_seq) */
	int32_t threads;                   /* host worker threads (param->num_threads) */
	int32_t chunk_reads;               /* reads per pipeline chunk; 0 = default */
	/* -ref artifact filter (param->reference_fasta; match_to_reference barcode_hmm.c:2478-2583).  With a refset every
	 * pipeline chunk is exactly one chunk of the reference's loop (ref_chunk_reads = param->num_query reads), because the
	 * reference's result depends on the thread slicing (`threads`) of each run_pHMM / run_rna_dust call. */
	const tdg_refset* refset;          /* NULL = no artifact filter */
	int32_t filter_error;              /* param->filter_error */
	int32_t ref_chunk_reads;           /* param->num_query */
	int64_t* artifact_counts;          /* out, [sequences of the refset] or NULL: reference_fasta->mer_hash (:381) */
} tdg_demux_job;

typedef struct tdg_demux_stats {       /* struct log_information, barcode_hmm.c:232-241, :356-384 */
	int64_t total_read;
	int64_t num_EXTRACT_SUCCESS;
	int64_t num_EXTRACT_FAIL_BAR_FINGER_NOT_FOUND;
	int64_t num_EXTRACT_FAIL_READ_TOO_SHORT;
	int64_t num_EXTRACT_FAIL_AMBIGIOUS_BARCODE;
	int64_t num_EXTRACT_FAIL_ARCHITECTURE_MISMATCH;
	int64_t num_EXTRACT_FAIL_MATCHES_ARTIFACTS;
	int64_t num_EXTRACT_FAIL_LOW_COMPLEXITY;
	int64_t long_sequence_events;      /* reads with len >= the running max_seq_len (:293-309: one model rebuild each) */
	double  seconds_split, seconds_parse, seconds_gpu_wait, seconds_write, seconds_total;  /* busy time per stage */
} tdg_demux_stats;

/* Runs the whole job; blocking.  Returns TDG_OK, or an error code with the message in the last-error string
 * ("Input File:%s and %s differ in number of entries." etc. carry the reference's wording). */
int  tdg_demux_run(tdg_context* ctx, const tdg_demux_job* job, tdg_demux_stats* stats);

/* model helpers the stream layer needs */
int  tdg_model_max_len(const tdg_model* m);
int  tdg_model_set_max_len(tdg_model* m, int max_len);   /* models do not depend on the read length; only scratch sizing does */
int  tdg_model_num_hmms(const tdg_model* m);
/* is_read[h] = 1 where HMM h belongs to an 'R' segment (make_extracted_read, barcode_hmm.c:3325-3356) */
int  tdg_model_read_hmms(const tdg_model* m, uint8_t* is_read);

#ifdef __cplusplus
}
#endif
#endif /* TAGDUST_B200_STREAM_H */
