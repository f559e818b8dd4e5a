/* controller_gpu.c -- the second reference-side binding: a drop-in definition of
 *
 *     int hmm_controller_multiple(struct parameters* param);      barcode_hmm.c:51
 *
 * that keeps the reference's own set-up (architecture selection, get_sequence_stats, threshold
 * calibration, init_model_bag -- all called as the reference's functions) and its summary
 * messages, and hands the per-chunk loop (read_fasta_fastq -> run_pHMM / run_rna_dust -> merge ->
 * print_all, barcode_hmm.c:243-384) to tdg_demux_run() of libtagdust_b200.so.
 *
 * Why the loop has to move: the reference rebuilds the whole model_bag for every read that is at
 * least as long as the longest read seen so far (barcode_hmm.c:293-309) -- i.e. for every read of a
 * fixed-length Illumina run -- and re-opens every output file per chunk; with the HMM on the GPU
 * those two dominate the run time by orders of magnitude.
 *
 * SAM/BAM input (read by `samtools view` through the reference's io_handler / read_sam_chunk, which the streaming
 * reader does not replace) takes a third route, chunk_loop() below: the reference's reader and print_all() around the
 * GPU run_pHMM, chunk by chunk like the reference's loop but WITHOUT the per-read model rebuild -- the model is
 * re-made at most once per chunk, and only when a read no longer fits it.  TDG_CONTROLLER=chunks forces that route for
 * any input (tests).  TDG_REFERENCE_CONTROLLER=1 calls the reference's own controller (found with dlsym(RTLD_NEXT)).
 * The -ref artifact filter runs on the GPU (k_artifact) on every route; the reference's get_fasta() reads the sequences.
 *
 * Documented differences, log file only: "Long sequence found. Need to realloc model..." is
 * written once with the number of reads it applies to instead of once per read, and the
 * per-read " %d %d" lines on stderr are not printed.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kslib.h"
#include "tagdust2.h"
#include "interface.h"
#include "misc.h"
#include "io.h"
#include "barcode_hmm.h"

#include "tagdust_b200_stream.h"
#include "shim.h"

typedef int (*controller_fn)(struct parameters*);

#include <time.h>
/* TDG_VERBOSE: wall-clock marks of the controller's phases on stderr */
static double phase_t0 = 0;
static void phase(const char* what)
{
	struct timespec ts;
	double t;
	if (!getenv("TDG_VERBOSE")) return;
	clock_gettime(CLOCK_REALTIME, &ts);
	t = ts.tv_sec + 1e-9 * ts.tv_nsec;
	if (phase_t0 == 0) phase_t0 = t;
	fprintf(stderr, "tagdust_b200: [%8.3f s, epoch %.3f] %s\n", t - phase_t0, t, what);
}

static void say(struct parameters* param, const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(param->buffer, MSG_BUFFER_SIZE, fmt, ap);
	va_end(ap);
	param->messages = append_message(param->messages, param->buffer);
}

static void die(struct parameters* param)
{
	tdg_shim_warmup_join();
	free_param(param);
	exit(EXIT_FAILURE);
}

static int ends_with(const char* s, const char* suf)
{
	size_t a = strlen(s), b = strlen(suf);
	return a >= b && !strcmp(s + a - b, suf);
}

static int needs_reference_controller(void)
{
	const char* e = getenv("TDG_REFERENCE_CONTROLLER");
	return e && atoi(e);
}

static int needs_chunk_loop(struct parameters* param)
{
	int i;
	const char* e = getenv("TDG_CONTROLLER");
	if (e && !strcmp(e, "chunks")) return 1;
	for (i = 0; i < param->infiles; i++) {
		const char* f = param->infile[i];
		if (ends_with(f, ".sam") || ends_with(f, ".bam") || ends_with(f, ".sam.gz") || ends_with(f, ".bam.gz")) return 1;
	}
	return 0;
}

/* The reference's chunk loop (barcode_hmm.c:216-432) with its own reader (read_sam_chunk / read_fasta_fastq through
 * io_handler) and writer (print_all), run_pHMM / run_rna_dust being the GPU definitions of run_phmm_gpu.c.  What is
 * different: a read at least as long as the longest one before it (:293-309) is counted, and the model is re-made once
 * per chunk only if such a read no longer fits its dynamic-programming length (the tables do not depend on that length). */
static int chunk_loop(struct parameters* param, struct sequence_stats_info** ssi, struct model_bag** bags, char* read_present, long long barcode_present)
{
	const int nf = param->infiles;
	int i, j, c, status = kslOK;
	long long long_events = 0;
	int total_read = 0, n_ok = 0, n_bar = 0, n_short = 0, n_arch = 0, n_art = 0, n_low = 0;
	struct fasta* reference_fasta = NULL;
	FILE** files = calloc(nf, sizeof *files);
	struct read_info*** ric = calloc(nf, sizeof *ric);
	int* numseqs = calloc(nf, sizeof(int));
	int (*reader)(struct read_info**, struct parameters*, FILE*, int*) = NULL;

	if (param->reference_fasta) {   /* :208-215 */
		reference_fasta = get_fasta(reference_fasta, param->reference_fasta);
		if (!reference_fasta) { status = kslFAIL; goto OUT; }
		reference_fasta->mer_hash = calloc(reference_fasta->numseq > 0 ? reference_fasta->numseq : 1, sizeof(int));
	}
	for (i = 0; i < nf; i++) {
		ric[i] = malloc_read_info(ric[i], param->num_query);
		files[i] = io_handler(files[i], i, param);
	}
	reader = param->sam == 0 ? &read_fasta_fastq : &read_sam_chunk;
	phase("chunk loop starts");
	for (;;) {
		c = 0;
		for (i = 0; i < nf; i++) {
			if (reader(ric[i], param, files[i], &numseqs[i]) != kslOK) {
				snprintf(param->errmsg, kslibERRBUFSIZE, "Failed to read data chunk from file: %s", param->infile[i]);
				status = kslFAIL; goto OUT;
			}
			c += numseqs[i];
		}
		if (!c) break;
		for (i = 0; i < nf - 1; i++)
			for (j = i + 1; j < nf; j++)
				if (numseqs[i] != numseqs[j]) { say(param, "Input File:%s and %s differ in number of entries.\n", param->infile[i], param->infile[j]); die(param); }
		if (!total_read)   /* :271-287 */
			for (i = 0; i < nf - 1; i++)
				for (j = i + 1; j < nf; j++)
					for (c = 0; c < (numseqs[0] < 1000 ? numseqs[0] : 1000); c++)
						if (compare_read_names(param, ric[i][c]->name, ric[j][c]->name)) {
							say(param, "Files seem to contain reads in different order:\n%s\n%s\n", ric[i][c]->name, ric[j][c]->name);
							die(param);
						}
		for (i = 0; i < nf; i++) {
			for (j = 0; j < numseqs[0]; j++)
				if (ric[i][j]->len >= ssi[i]->max_seq_len) { ssi[i]->max_seq_len = ric[i][j]->len; long_events++; }
			if (ssi[i]->max_seq_len + 10 > bags[i]->current_dyn_length) {   /* init_model_bag sizes for max_seq_len + 10 (:5778) */
				param->read_structure = param->read_structures[i];
				free_model_bag(bags[i]);
				bags[i] = init_model_bag(param, ssi[i]);
			}
		}
		for (i = 0; i < nf; i++) {
			param->read_structure = param->read_structures[i];
			param->confidence_threshold = param->confidence_thresholds[i];
			if (param->read_structure->num_segments == 1 && param->read_structure->type[0] == 'R') {
				if (run_rna_dust(ric[i], param, reference_fasta, numseqs[i]) != kslOK) { snprintf(param->errmsg, kslibERRBUFSIZE, "run_rna_dust failed\n"); status = kslFAIL; goto OUT; }
			} else {
				if (run_pHMM(0, bags[i], ric[i], param, reference_fasta, numseqs[i], MODE_GET_LABEL) != kslOK) { status = kslFAIL; goto OUT; }
			}
		}
		for (i = 0; i < nf; i++)   /* :329-341 */
			if (barcode_present & (1 << i)) {
				param->read_structure = param->read_structures[i];
				if (i) for (j = 0; j < numseqs[0]; j++) ric[0][j]->barcode = ric[i][j]->barcode;
				break;
			}
		for (i = 0; i < numseqs[0]; i++) {
			c = -100000;
			for (j = 0; j < nf; j++) if (ric[j][i]->read_type > c) c = ric[j][i]->read_type;
			ric[0][i]->read_type = c;
		}
		print_all(ric, param, numseqs[0], read_present);
		total_read += numseqs[0];
		for (i = 0; i < numseqs[0]; i++)   /* :358-382 */
			switch ((int)ric[0][i]->read_type) {
				case EXTRACT_SUCCESS: n_ok++; break;
				case EXTRACT_FAIL_BAR_FINGER_NOT_FOUND: n_bar++; break;
				case EXTRACT_FAIL_READ_TOO_SHORT: n_short++; break;
				case EXTRACT_FAIL_ARCHITECTURE_MISMATCH: n_arch++; break;
				case EXTRACT_FAIL_MATCHES_ARTIFACTS: n_art++; n_low++; break;   /* falls through in the reference */
				case EXTRACT_FAIL_LOW_COMPLEXITY: n_low++; break;
				default:
					n_art++;
					if (reference_fasta) reference_fasta->mer_hash[((int)(ric[0][i]->read_type) >> 8) - 1]++;
					break;
			}
	}
	phase("chunk loop done");
	if (long_events)
		say(param, "Long sequence found. Need to realloc model... (%lld reads at least as long as the longest seen before; the GPU model does not depend on the read length)\n",
		    long_events);
	say(param, "Done.\n\n");
	for (i = 0; i < nf; i++) say(param, "%s	Input file %d.\n", param->infile[i], i);
	say(param, "%d	total input reads\n", total_read);
	say(param, "%0.2f	selected threshold\n", param->confidence_threshold);
	say(param, "%d	successfully extracted\n", n_ok);
	say(param, "%0.1f%%	extracted\n", (float)n_ok / (float)total_read * 100.0f);
	say(param, "%d	problems with architecture\n", n_arch);
	say(param, "%d	barcode / UMI not found\n", n_bar);
	say(param, "%d	too short\n", n_short);
	say(param, "%d	low complexity\n", n_low);
	say(param, "%d	match artifacts:\n", n_art);
	if (reference_fasta)
		for (i = 0; i < reference_fasta->numseq; i++)
			if (reference_fasta->mer_hash[i]) say(param, "%d	%s\n", reference_fasta->mer_hash[i], reference_fasta->sn[i]);
OUT:
	if (reference_fasta) free_fasta(reference_fasta);
	for (i = 0; i < nf; i++) {
		if (ric[i]) free_read_info(ric[i], param->num_query);
		if (files[i]) pclose(files[i]);
	}
	free(files); free(ric); free(numseqs);
	return status;
}

int hmm_controller_multiple(struct parameters* param)
{
	const int nf = param->infiles;
	int i, j, status = kslOK;
	if (needs_reference_controller()) {
		controller_fn ref = (controller_fn)dlsym(RTLD_NEXT, "hmm_controller_multiple");
		if (!ref) { fprintf(stderr, "tagdust_b200: reference controller not found\n"); return kslFAIL; }
		return ref(param);
	}

	phase("controller start");
	tdg_shim_warmup();
	struct sequence_stats_info** ssi = calloc(nf, sizeof *ssi);
	struct model_bag** bags = calloc(nf, sizeof *bags);
	char* read_present = calloc(nf, 1);
	int* file_max_len = calloc(nf, sizeof(int));   /* longest read get_sequence_stats saw, before calibration raises ssi->max_seq_len */
	long long barcode_present = 0;
	int num_out_reads = 0;
	param->read_structures = calloc(nf, sizeof(struct read_structure*));
	param->confidence_thresholds = calloc(nf, sizeof(float));

	/* ---- one architecture per input file (barcode_hmm.c:101-138) */
	for (i = 0; i < nf; i++) {
		if (!i && param->read_structure->num_segments) {
			/* given on the command line */
		} else if (param->arch_file) {
			if (test_architectures(param, i) != kslOK) {
				snprintf(param->errmsg, kslibERRBUFSIZE, "Test architecture on file %s failed.\n", param->infile[i]);
				status = kslFAIL;
				goto DONE;
			}
		} else {
			if ((param->read_structure = malloc_read_structure()) == NULL) { status = kslEMEM; goto DONE; }
			if (assign_segment_sequences(param, "R:N", 0) != kslOK) { status = kslFAIL; goto DONE; }
			if (QC_read_structure(param)) { say(param, "Something wrong with architecture....\n"); die(param); }
		}
		param->read_structures[i] = param->read_structure;
		param->read_structure = NULL;
		for (j = 0; j < param->read_structures[i]->num_segments; j++) {
			if (param->read_structures[i]->type[j] == 'B') barcode_present |= (1 << i);
			if (param->read_structures[i]->type[j] == 'R') read_present[i]++;
		}
	}
	if (bitcount64(barcode_present) > 1) { say(param, "Barcodes seem to be in both architectures... \n"); die(param); }
	for (i = 0; i < nf; i++) num_out_reads += read_present[i];
	for (i = 0; i < nf; i++)
		if (barcode_present & (1 << i)) {
			param->read_structure = param->read_structures[i];
			j = check_for_existing_demultiplexed_files_multiple(param, num_out_reads);
			param->read_structure = NULL;
			if (j) { snprintf(param->errmsg, kslibERRBUFSIZE, "Error: some output files already exists.\n"); status = kslFAIL; goto DONE; }
		}

	init_logsum();
#if RTEST
	param->num_query = 1000;
#else
	param->num_query = 1000001;
#endif

	phase("architectures chosen");
	/* ---- sequence statistics, thresholds, models: the reference's own functions (:180-207) */
	{
		struct read_info** ri = NULL;
		ri = malloc_read_info(ri, param->num_query);
		for (i = 0; i < nf; i++) {
			param->read_structure = param->read_structures[i];
			ssi[i] = get_sequence_stats(param, ri, i);
			file_max_len[i] = ssi[i] ? ssi[i]->max_seq_len : 0;
		}
		free_read_info(ri, param->num_query);
	}
	phase("sequence statistics done");
	if (!param->confidence_threshold) {
		for (i = 0; i < nf; i++) {
			say(param, "Determining threshold for read%d.\n", i);
			param->read_structure = param->read_structures[i];
			if (estimateQthreshold(param, ssi[i]) != kslOK) {
				snprintf(param->errmsg, kslibERRBUFSIZE, "estimateQthreshold failed.\n");
				status = kslFAIL;
				goto DONE;
			}
			param->confidence_thresholds[i] = param->confidence_threshold;
		}
	}
	for (i = 0; i < nf; i++) {
		param->read_structure = param->read_structures[i];
		bags[i] = init_model_bag(param, ssi[i]);
	}

	phase("thresholds and models done");
	if (needs_chunk_loop(param)) {
		status = chunk_loop(param, ssi, bags, read_present, barcode_present);
		goto DONE;
	}
	/* ---- read-name order check of the first chunk (:271-287) on the first 1000 entries of every file */
	if (nf > 1) {
		struct read_info*** head = calloc(nf, sizeof *head);
		int* cnt = calloc(nf, sizeof(int));
		const int keep = param->num_query;
		param->num_query = 1000;
		for (i = 0; i < nf; i++) {
			FILE* fh = NULL;
			head[i] = malloc_read_info(head[i], 1000);
			fh = io_handler(fh, i, param);
			if (read_fasta_fastq(head[i], param, fh, &cnt[i]) != kslOK) { status = kslFAIL; }
			pclose(fh);
		}
		for (i = 0; i < nf - 1 && status == kslOK; i++)
			for (j = i + 1; j < nf; j++) {
				int c, lim = cnt[i] < cnt[j] ? cnt[i] : cnt[j];
				for (c = 0; c < lim; c++)
					if (compare_read_names(param, head[i][c]->name, head[j][c]->name)) {
						say(param, "Files seem to contain reads in different order:\n%s\n%s\n", head[i][c]->name, head[j][c]->name);
						die(param);
					}
			}
		for (i = 0; i < nf; i++) free_read_info(head[i], 1000);
		free(head); free(cnt);
		param->num_query = keep;
		if (status != kslOK) goto DONE;
	}
	/* ---- the streaming job */
	{
		struct fasta* reference_fasta = NULL;
		int64_t* artifact_counts = NULL;
		tdg_demux_input* in = calloc(nf, sizeof *in);
		tdg_demux_job job;
		tdg_demux_stats st;
		tdg_context* ctx = NULL;
		int need_gpu = 0, rc;
		memset(&job, 0, sizeof job);
		job.barcode_input = -1;
		job.num_alternatives = 2;
		for (i = 0; i < nf; i++) {
			struct read_structure* rs = param->read_structures[i];
			in[i].path = param->infile[i];
			in[i].fasta = -1;
			in[i].num_read_segments = read_present[i];
			in[i].confidence_threshold = param->confidence_thresholds[i];
			in[i].max_seq_len = ssi[i]->max_seq_len;
			in[i].expected_len = file_max_len[i];
			if (!(rs->num_segments == 1 && rs->type[0] == 'R')) need_gpu = 1;   /* :312 */
		}
		/* known contaminants (barcode_hmm.c:208-215): read with the reference's own get_fasta, matched on the device */
		if (param->reference_fasta) {
			reference_fasta = get_fasta(reference_fasta, param->reference_fasta);
			if (!reference_fasta) { status = kslFAIL; free(in); goto DONE; }
			need_gpu = 1;
		}
		if (need_gpu && !(ctx = tdg_shim_context(param))) { status = kslFAIL; free(in); goto DONE; }
		if (reference_fasta) {
			job.refset = tdg_shim_get_refset(reference_fasta, param);
			if (!job.refset) {
				snprintf(param->errmsg, kslibERRBUFSIZE, "tdg_refset_create: %s", tdg_last_error());
				status = kslFAIL; free(in); goto DONE;
			}
			job.filter_error = param->filter_error;
			job.ref_chunk_reads = param->num_query;
			artifact_counts = calloc(reference_fasta->numseq > 0 ? reference_fasta->numseq : 1, sizeof(int64_t));
			job.artifact_counts = artifact_counts;
		}
		for (i = 0; i < nf; i++) {
			struct read_structure* rs = param->read_structures[i];
			if (rs->num_segments == 1 && rs->type[0] == 'R') continue;
			param->read_structure = rs;
			/* sized for the reads of the file, not for the longest simulated calibration read
			 * (estimateQthreshold raises ssi->max_seq_len, calibrateQ.c:121-126) */
			in[i].model = tdg_shim_get_model_len(bags[i], param, file_max_len[i] + 10);
			if (!in[i].model) {
				snprintf(param->errmsg, kslibERRBUFSIZE, "tdg_model_create: %s", tdg_last_error());
				status = kslFAIL; free(in); goto DONE;
			}
		}
		/* file naming follows the architecture that holds the barcode (:329-341, io.c:822-838) */
		for (i = 0; i < nf; i++)
			if (barcode_present & (1 << i)) {
				struct read_structure* rs = param->read_structures[i];
				for (j = 0; j < rs->num_segments; j++)
					if (rs->type[j] == 'B') {
						job.barcode_input = i;
						job.num_alternatives = rs->numseq_in_segment[j];
						job.barcode_names = (const char* const*)rs->sequence_matrix[j];
						break;
					}
				break;
			}
		job.n_inputs = nf;
		job.inputs = in;
		job.outfile = param->outfile;
		job.minlen = param->minlen;
		job.dust = param->dust;
		job.matchstart = param->matchstart;
		job.matchend = param->matchend;
		job.print_seq_finger = param->print_seq_finger;
		job.threads = param->num_threads;
		job.chunk_reads = 0;
		{
			const char* e = getenv("TDG_CHUNK_READS");
			if (e && atoi(e) > 0) job.chunk_reads = atoi(e);
		}
		phase("streaming job starts");
		rc = tdg_demux_run(ctx, &job, &st);
		phase("streaming job done");
		free(in);
		if (rc != TDG_OK) {
			say(param, "%s\n", tdg_last_error());
			fprintf(stderr, "tagdust_b200: %s\n", tdg_last_error());
			die(param);
		}
		if (st.long_sequence_events)
			say(param, "Long sequence found. Need to realloc model... (%lld reads at least as long as the longest seen before; the GPU model does not depend on the read length)\n",
			    (long long)st.long_sequence_events);
		param->confidence_threshold = param->confidence_thresholds[nf - 1];

		/* ---- summary (barcode_hmm.c:386-428) */
		say(param, "Done.\n\n");
		for (i = 0; i < nf; i++) say(param, "%s	Input file %d.\n", param->infile[i], i);
		say(param, "%d	total input reads\n", (int)st.total_read);
		say(param, "%0.2f	selected threshold\n", param->confidence_threshold);
		say(param, "%d	successfully extracted\n", (int)st.num_EXTRACT_SUCCESS);
		say(param, "%0.1f%%	extracted\n", (float)st.num_EXTRACT_SUCCESS / (float)st.total_read * 100.0f);
		say(param, "%d	problems with architecture\n", (int)st.num_EXTRACT_FAIL_ARCHITECTURE_MISMATCH);
		say(param, "%d	barcode / UMI not found\n", (int)st.num_EXTRACT_FAIL_BAR_FINGER_NOT_FOUND);
		say(param, "%d	too short\n", (int)st.num_EXTRACT_FAIL_READ_TOO_SHORT);
		say(param, "%d	low complexity\n", (int)st.num_EXTRACT_FAIL_LOW_COMPLEXITY);
		say(param, "%d	match artifacts:\n", (int)st.num_EXTRACT_FAIL_MATCHES_ARTIFACTS);
		if (reference_fasta) {   /* :423-431 */
			for (i = 0; i < reference_fasta->numseq; i++)
				if (artifact_counts[i]) say(param, "%d	%s\n", (int)artifact_counts[i], reference_fasta->sn[i]);
			free(artifact_counts);
		}
		if (getenv("TDG_VERBOSE"))
			fprintf(stderr, "tagdust_b200: %lld reads in %.2f s (busy: line split %.2f s, convert+pack %.2f s, gpu wait %.2f s, write %.2f s)\n",
			        (long long)st.total_read, st.seconds_total, st.seconds_split, st.seconds_parse, st.seconds_gpu_wait, st.seconds_write);
	}

DONE:
	phase("controller returns");
	tdg_shim_warmup_join();
	for (i = 0; i < nf; i++) {
		if (bags[i]) free_model_bag(bags[i]);
		if (ssi[i]) free(ssi[i]);
	}
	param->read_structure = 0;
	free(bags); free(ssi); free(read_present); free(file_max_len);
	if (status != kslOK) fprintf(stderr, "%s\n", param->errmsg);
	return status;
}
