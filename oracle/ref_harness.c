/* TEST INFRASTRUCTURE -- reference harness.
 *
 * Compiled TOGETHER WITH the unmodified reference sources (see oracle/Makefile)
 * into oracle/_ref/libtagdust_ref[_rtest].so.  It contains no algorithm of its
 * own: every number it returns is produced by the reference's functions
 *   init_logsum               misc.c:57
 *   assign_segment_sequences  interface.c:489
 *   init_model_bag            barcode_hmm.c:5760
 *   backward                  barcode_hmm.c:3439
 *   forward_max_posterior_decoding  barcode_hmm.c:4128
 *   run_pHMM                  barcode_hmm.c:1895
 * It only (a) builds `struct parameters` / `struct sequence_stats_info` the way
 * main.c/interface.c/io.c would, (b) marshals flat arrays <-> `struct read_info`,
 * and (c) flattens `struct model_bag` into the tdg_model_desc layout of
 * include/tagdust_b200.h so tests can feed the same model to the CUDA path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * leg may load this library.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>
#include "kslib.h"
#include "tagdust2.h"
#include "interface.h"
#include "nuc_code.h"
#include "misc.h"
#include "io.h"
#include "barcode_hmm.h"

extern float logsum_lookup[LOGSUM_SIZE];

static int refh_inited = 0;

void refh_init(void)
{
	if(!refh_inited){
		init_logsum();
		init_nuc_code();
		refh_inited = 1;
	}
}

/* copy of the reference's logsum table, for bit-compare with the product's own */
void refh_logsum_table(float* out)
{
	refh_init();
	memcpy(out, logsum_lookup, sizeof(float) * LOGSUM_SIZE);
}

float refh_logsum(float a, float b){ refh_init(); return logsum(a,b); }

/* ---------------- parameters ---------------- */

struct parameters* refh_param_new(int num_segments, const char** segment_strings,
                                  float sequencer_error_rate, float indel_frequency,
                                  int minlen, int num_threads, float confidence_threshold,
                                  int dust, int matchstart, int matchend)
{
	struct parameters* param = calloc(1, sizeof(struct parameters));
	int i;
	refh_init();
	/* defaults: interface.c:66-127 */
	param->num_threads = num_threads;
	param->num_query = 1000000;
	param->matchstart = matchstart;
	param->matchend = matchend;
	param->minlen = minlen;
	param->dust = dust;
	param->sequencer_error_rate = sequencer_error_rate;
	param->indel_frequency = indel_frequency;
	param->average_read_length = 50;
	param->numbarcode = 8;
	param->confidence_threshold = confidence_threshold;
	param->filter_error = 2;
	param->buffer = calloc(MSG_BUFFER_SIZE + 16, 1);
	param->messages = NULL;
	param->quiet_flag = 1;
	param->read_structure = malloc_read_structure();
	for(i = 0; i < num_segments; i++){
		char* tmp = strdup(segment_strings[i]);
		if(assign_segment_sequences(param, tmp, i) != kslOK){
			free(tmp);
			return NULL;
		}
		free(tmp);
	}
	return param;
}

void refh_param_set(struct parameters* param, float confidence_threshold, int minlen, int dust, int num_threads)
{
	param->confidence_threshold = confidence_threshold;
	param->minlen = minlen;
	param->dust = dust;
	param->num_threads = num_threads;
}

void refh_param_free(struct parameters* param)
{
	if(!param) return;
	if(param->read_structure) free_read_structure(param->read_structure);
	free(param->buffer);
	if(param->messages) free(param->messages);
	free(param);
}

int refh_param_segment_info(struct parameters* param, int seg, char* type, int* numseq, int* seqlen)
{
	*type = param->read_structure->type[seg];
	*numseq = param->read_structure->numseq_in_segment[seg];
	*seqlen = (int)strlen(param->read_structure->sequence_matrix[seg][0]);
	return 0;
}

const char* refh_param_segment_seq(struct parameters* param, int seg, int idx)
{
	return param->read_structure->sequence_matrix[seg][idx];
}

/* ---------------- model ---------------- */

/* background_logp: the five log-space background values exactly as
 * get_sequence_stats leaves them in ssi->background (io.c:263-270). */
struct model_bag* refh_model_new(struct parameters* param, const double* background_logp,
                                 double average_length, int max_seq_len,
                                 double expected_5_len, double mean_5_len, double stdev_5_len,
                                 double expected_3_len, double mean_3_len, double stdev_3_len)
{
	struct sequence_stats_info ssi;
	int i;
	refh_init();
	memset(&ssi, 0, sizeof(ssi));
	for(i = 0; i < 5; i++) ssi.background[i] = background_logp[i];
	ssi.average_length = average_length;
	ssi.max_seq_len = max_seq_len;
	ssi.expected_5_len = expected_5_len;
	ssi.mean_5_len = mean_5_len;
	ssi.stdev_5_len = stdev_5_len;
	ssi.expected_3_len = expected_3_len;
	ssi.mean_3_len = mean_3_len;
	ssi.stdev_3_len = stdev_3_len;
	return init_model_bag(param, &ssi);
}

/* the model edit estimateQthreshold applies before emitting (calibrateQ.c:67-86) */
void refh_model_calibration_edit(struct model_bag* mb, struct parameters* param)
{
	int i,j;
	for(i = 0; i < mb->num_models;i++){
		if(param->read_structure->type[i] == 'B' || param->read_structure->type[i] == 'S'){
			for(j = 0 ; j < mb->model[i]->num_hmms-1;j++){
				mb->model[i]->silent_to_M[j][0] = prob2scaledprob(1.0 / (float)( mb->model[i]->num_hmms-1));
			}
			mb->model[i]->silent_to_M[mb->model[i]->num_hmms-1][0] = prob2scaledprob(0.0);
		}
	}
}

void refh_model_free(struct model_bag* mb){ if(mb) free_model_bag(mb); }

void refh_model_dims(struct model_bag* mb, int* num_segments, int* total_hmms, int* total_columns,
                     int* average_raw_length, int* current_dyn_length)
{
	int j, c = 0;
	for(j = 0; j < mb->num_models; j++){
		c += mb->model[j]->num_hmms * mb->model[j]->hmms[0]->num_columns;
	}
	*num_segments = mb->num_models;
	*total_hmms = mb->total_hmm_num;
	*total_columns = c;
	*average_raw_length = mb->average_raw_length;
	*current_dyn_length = mb->current_dyn_length;
}

/* Flatten in the order segment -> hmm -> column (tdg_model_desc layout). */
void refh_model_flatten(struct model_bag* mb,
                        int* seg_num_hmms, int* seg_num_cols, float* seg_skip,
                        float* background /*5*/,
                        float* transition /*C*9*/, float* m_emit /*C*5*/, float* i_emit /*C*5*/,
                        float* silent_to_M /*C*/, float* silent_to_I /*C*/,
                        int* label /*H*/, float* transition_matrix /*H*H*/)
{
	int j,f,g,k,c = 0;
	int H = mb->total_hmm_num;
	for(k = 0; k < 5; k++) background[k] = mb->model[0]->background_nuc_frequency[k];
	for(j = 0; j < mb->num_models; j++){
		struct model* m = mb->model[j];
		seg_num_hmms[j] = m->num_hmms;
		seg_num_cols[j] = m->hmms[0]->num_columns;
		seg_skip[j] = m->skip;
		for(f = 0; f < m->num_hmms; f++){
			for(g = 0; g < m->hmms[f]->num_columns; g++){
				struct hmm_column* col = m->hmms[f]->hmm_column[g];
				for(k = 0; k < 9; k++) transition[c*9+k] = col->transition[k];
				for(k = 0; k < 5; k++) m_emit[c*5+k] = col->m_emit[k];
				for(k = 0; k < 5; k++) i_emit[c*5+k] = col->i_emit[k];
				silent_to_M[c] = m->silent_to_M[f][g];
				silent_to_I[c] = m->silent_to_I[f][g];
				c++;
			}
		}
	}
	for(j = 0; j < H; j++){
		label[j] = mb->label[j];
		for(k = 0; k < H; k++) transition_matrix[j*H+k] = mb->transition_matrix[j][k];
	}
}

/* ---------------- reads ---------------- */

static struct read_info** refh_make_reads(int n, const unsigned char* codes, int stride, const int* lens)
{
	struct read_info** ri = 0;
	int i,j;
	ri = malloc_read_info(ri, n);
	for(i = 0; i < n; i++){
		int len = lens[i];
		/* io.c:1749-1762: seq and labels are len+1 long, NUL terminated */
		ri[i]->seq = malloc(len + 2);
		ri[i]->labels = malloc(len + 2);
		ri[i]->qual = malloc(len + 2);
		ri[i]->name = malloc(32);
		snprintf(ri[i]->name, 32, "r%d", i);
		for(j = 0; j < len; j++){
			ri[i]->seq[j] = (char)codes[(size_t)i*stride + j];
			ri[i]->labels[j] = 0;
			ri[i]->qual[j] = 'I';
		}
		ri[i]->seq[len] = 0; ri[i]->seq[len+1] = 0;
		ri[i]->labels[len] = 0; ri[i]->labels[len+1] = 0;
		ri[i]->qual[len] = 0; ri[i]->qual[len+1] = 0;
		ri[i]->len = len;
	}
	return ri;
}

/* Per-read scores straight from backward()/forward_max_posterior_decoding()
 * (the fields run_pHMM discards): f_score, b_score, r_score, bar_prob, labels.
 * labels_out is n*stride bytes; labels[0..len] are written.               */
int refh_decode_scores(struct model_bag* mb_in, int n, const unsigned char* codes, int stride,
                       const int* lens, float* f_score, float* b_score, float* r_score,
                       double* bar_prob, unsigned char* labels_out)
{
	struct model_bag* mb = copy_model_bag(mb_in);
	struct read_info** ri = refh_make_reads(n, codes, stride, lens);
	int i,j;
	for(i = 0; i < n; i++){
		mb = backward(mb, ri[i]->seq, ri[i]->len);
		mb = forward_max_posterior_decoding(mb, ri[i], ri[i]->seq, ri[i]->len);
		f_score[i] = mb->f_score;
		b_score[i] = mb->b_score;
		r_score[i] = mb->r_score;
		bar_prob[i] = ri[i]->bar_prob;
		if(labels_out){
			for(j = 0; j <= ri[i]->len; j++) labels_out[(size_t)i*stride + j] = (unsigned char)ri[i]->labels[j];
		}
	}
	free_read_info(ri, n);
	free_model_bag(mb);
	return 0;
}

/* b_score only (MODE_ARCH_COMP building block, barcode_hmm.c:2126-2135) */
int refh_backward_scores(struct model_bag* mb_in, int n, const unsigned char* codes, int stride,
                         const int* lens, float* b_score)
{
	struct model_bag* mb = copy_model_bag(mb_in);
	struct read_info** ri = refh_make_reads(n, codes, stride, lens);
	int i;
	for(i = 0; i < n; i++){
		mb = backward(mb, ri[i]->seq, ri[i]->len);
		b_score[i] = mb->b_score;
	}
	free_read_info(ri, n);
	free_model_bag(mb);
	return 0;
}

/* Full posterior matrix of ONE read after the exp() step but before the label
 * DP is not observable through the reference API; the matrix AFTER the DP is
 * (mb->dyn_prog_matrix).  Dump it for debugging parity of the label DP.      */
int refh_decode_matrix(struct model_bag* mb_in, const unsigned char* codes, int len,
                       float* dyn_out /* (len+1)*H */, int* path_out /* (len+1)*H */)
{
	struct model_bag* mb = copy_model_bag(mb_in);
	struct read_info** ri = refh_make_reads(1, codes, len + 1, &len);
	int i,j,H = mb->total_hmm_num;
	mb = backward(mb, ri[0]->seq, len);
	mb = forward_max_posterior_decoding(mb, ri[0], ri[0]->seq, len);
	for(i = 0; i <= len; i++){
		for(j = 0; j < H; j++){
			dyn_out[i*H+j] = mb->dyn_prog_matrix[i][j];
			path_out[i*H+j] = mb->path[i][j];
		}
	}
	free_read_info(ri, 1);
	free_model_bag(mb);
	return 0;
}

/* The seam itself: run_pHMM() in MODE_GET_LABEL (1) or MODE_GET_PROB (4).
 * seq_out/qual_out (n*stride, may be NULL) receive the in-place rewritten
 * sequence/quality (spacer 65) and len_out the post-extraction length.      */
int refh_run_phmm(struct model_bag* mb, struct parameters* param, int mode,
                  int n, const unsigned char* codes, int stride, const int* lens,
                  float* mapq, double* bar_prob, unsigned char* labels_out,
                  int* read_type, int* barcode, int* fingerprint,
                  unsigned char* seq_out, unsigned char* qual_out, int* len_out)
{
	struct read_info** ri = refh_make_reads(n, codes, stride, lens);
	int i,j,status;
	status = run_pHMM(0, mb, ri, param, 0, n, mode);
	for(i = 0; i < n; i++){
		mapq[i] = ri[i]->mapq;
		bar_prob[i] = ri[i]->bar_prob;
		read_type[i] = ri[i]->read_type;
		barcode[i] = ri[i]->barcode;
		fingerprint[i] = ri[i]->fingerprint;
		if(labels_out){
			for(j = 0; j <= lens[i]; j++) labels_out[(size_t)i*stride + j] = (unsigned char)ri[i]->labels[j];
		}
		if(seq_out){
			for(j = 0; j < lens[i]; j++) seq_out[(size_t)i*stride + j] = (unsigned char)ri[i]->seq[j];
		}
		if(qual_out){
			for(j = 0; j < lens[i]; j++) qual_out[(size_t)i*stride + j] = (unsigned char)ri[i]->qual[j];
		}
		if(len_out) len_out[i] = ri[i]->len;
	}
	free_read_info(ri, n);
	return status;
}

/* ---------------- -ref artifact filter (match_to_reference, barcode_hmm.c:2478-2583) ----------------
 * A struct fasta built from flat arrays (string = nuc codes of all sequences back to back, s_index[numseq+1]);
 * refh_set_reference(param, ...) installs it for the next refh_run_phmm / refh_run_rna_dust calls, numseq 0 removes it. */
static struct fasta* refh_fasta = 0;
static char refh_fasta_name[] = "refh.fa";

void refh_set_reference(struct parameters* param, const unsigned char* string, const int* s_index, int numseq, int filter_error)
{
	int i;
	if(refh_fasta){
		free(refh_fasta->string); free(refh_fasta->s_index); free(refh_fasta->mer_hash); free(refh_fasta);
		refh_fasta = 0;
	}
	param->reference_fasta = 0;
	if(numseq <= 0) return;
	refh_fasta = calloc(1, sizeof(struct fasta));
	refh_fasta->numseq = numseq;
	refh_fasta->string_len = s_index[numseq];
	refh_fasta->string = malloc(s_index[numseq] + 8);
	memcpy(refh_fasta->string, string, s_index[numseq]);
	refh_fasta->s_index = malloc(sizeof(int) * (numseq + 1));
	for(i = 0; i <= numseq; i++) refh_fasta->s_index[i] = s_index[i];
	refh_fasta->mer_hash = calloc(numseq, sizeof(int));
	param->reference_fasta = refh_fasta_name;
	param->filter_error = filter_error;
}

/* run_pHMM(MODE_GET_LABEL) with the installed reference: read_type comes back as (sequence_id << 8) | 5 for artifacts */
int refh_run_phmm_ref(struct model_bag* mb, struct parameters* param, int n, const unsigned char* codes, int stride,
                      const int* lens, float* mapq, int* read_type, int* barcode, int* fingerprint, unsigned char* seq_out)
{
	struct read_info** ri = refh_make_reads(n, codes, stride, lens);
	int i, j, status;
	status = run_pHMM(0, mb, ri, param, refh_fasta, n, MODE_GET_LABEL);
	for(i = 0; i < n; i++){
		mapq[i] = ri[i]->mapq; read_type[i] = ri[i]->read_type; barcode[i] = ri[i]->barcode; fingerprint[i] = ri[i]->fingerprint;
		if(seq_out) for(j = 0; j < lens[i]; j++) seq_out[(size_t)i*stride + j] = (unsigned char)ri[i]->seq[j];
	}
	free_read_info(ri, n);
	return status;
}

/* run_rna_dust() (barcode_hmm.c:2043, :2370): files whose architecture is a single R segment */
int refh_run_rna_dust(struct parameters* param, int n, const unsigned char* codes, int stride, const int* lens, int* read_type)
{
	struct read_info** ri = refh_make_reads(n, codes, stride, lens);
	int i, status;
	status = run_rna_dust(ri, param, refh_fasta, n);
	for(i = 0; i < n; i++) read_type[i] = ri[i]->read_type;
	free_read_info(ri, n);
	return status;
}

/* the two Myers variants on their own (misc.c:572-636 bpm_check_error; :718-796 validate_bpm_sse -> bmp_single) */
int refh_bpm_check_error(const unsigned char* t, const unsigned char* p, int n, int m){ return bpm_check_error(t, p, n, m); }
int refh_bmp_single(const unsigned char* t, const unsigned char* p, int n, int m)
{
	unsigned char* q[4]; int l[4];
	q[0] = q[1] = q[2] = q[3] = (unsigned char*)p; l[0] = l[1] = l[2] = l[3] = m;
	validate_bpm_sse(q, l, (unsigned char*)t, n, 4);
	return l[0];
}

/* run_pHMM() in MODE_ARCH_COMP over A models; returns the normalised
 * log-posteriors exactly as test_architectures.c:184 receives them.        */
int refh_run_arch_comp(struct model_bag** archs, int num_arch, struct parameters* param,
                       int n, const unsigned char* codes, int stride, const int* lens,
                       float* arch_posterior)
{
	struct arch_bag ab;
	struct read_info** ri = refh_make_reads(n, codes, stride, lens);
	int i,status;
	ab.num_arch = num_arch;
	ab.archs = archs;
	ab.command_line = NULL;
	ab.arch_posterior = malloc(sizeof(float) * num_arch);
	for(i = 0; i < num_arch; i++) ab.arch_posterior[i] = prob2scaledprob(1.0);
	status = run_pHMM(&ab, archs[0], ri, param, 0, n, MODE_ARCH_COMP);
	for(i = 0; i < num_arch; i++) arch_posterior[i] = ab.arch_posterior[i];
	free(ab.arch_posterior);
	free_read_info(ri, n);
	return status;
}

/* Calibration read emitters (barcode_hmm.c:2599-3046), for row 15 tests. */

int refh_emit(struct model_bag* mb, int n_model, int n_random, int average_length, unsigned int seed,
              unsigned char* codes, int stride, int* lens)
{
	struct read_info** ri = 0;
	int i,j,n = n_model + n_random;
	unsigned int s = seed;
	ri = malloc_read_info(ri, n);
	srand(seed);
	for(i = 0; i < n; i++){
		int st;
		if(i < n_model) st = emit_read_sequence(mb, ri[i], average_length, &s);
		else st = emit_random_sequence(mb, ri[i], average_length, &s);
		if(st != kslOK) return -1;
		if(ri[i]->len + 1 > stride) return -2;
		lens[i] = ri[i]->len;
		for(j = 0; j < ri[i]->len; j++) codes[(size_t)i*stride + j] = (unsigned char)ri[i]->seq[j];
		codes[(size_t)i*stride + ri[i]->len] = 0;
	}
	free_read_info(ri, n);
	return 0;
}

size_t refh_sizeof_read_info(void){ return sizeof(struct read_info); }

/* ---------------- FASTQ reader (io_handler io.c:382 + read_fasta_fastq io.c:1684) ----------------
 * Reads chunk number `chunk_index` (0-based) of `num_query` reads from `path` with the reference's own
 * reader and flattens it: lens[n], codes/quals rows of `stride` bytes (quals all 0 when the reference
 * left ri->qual NULL), names joined with '\n'.  Returns the number of reads in that chunk, -1 on error. */
int refh_read_file_chunk(const char* path, int num_query, int chunk_index, int stride,
                         int* lens, unsigned char* codes, unsigned char* quals, int* has_qual,
                         char* names, size_t names_cap)
{
	struct parameters* param = calloc(1, sizeof(struct parameters));
	struct read_info** ri = NULL;
	FILE* file = NULL;
	int numseq = 0, i, k;
	size_t np = 0;
	char* files[1];
	refh_init();
	files[0] = (char*)path;
	param->infile = files;
	param->infiles = 1;
	param->num_query = num_query;
	param->buffer = calloc(MSG_BUFFER_SIZE + 16, 1);
	param->quiet_flag = 1;
	ri = malloc_read_info(ri, num_query);
	file = io_handler(file, 0, param);
	for(k = 0; k <= chunk_index; k++){
		if(read_fasta_fastq(ri, param, file, &numseq) != kslOK){ numseq = -1; break; }
		if(!numseq) break;
	}
	*has_qual = 0;
	for(i = 0; i < numseq; i++){
		size_t nl = strlen(ri[i]->name);
		lens[i] = ri[i]->len;
		if(ri[i]->len + 1 > stride || np + nl + 1 > names_cap){ numseq = -1; break; }
		memcpy(codes + (size_t)i * stride, ri[i]->seq, ri[i]->len + 1);
		if(ri[i]->qual){ memcpy(quals + (size_t)i * stride, ri[i]->qual, ri[i]->len + 1); *has_qual = 1; }
		memcpy(names + np, ri[i]->name, nl); np += nl;
		names[np++] = '\n';
	}
	if(np < names_cap) names[np] = 0;
	pclose(file);
	free_read_info(ri, num_query);
	free(param->buffer);
	if(param->messages) free(param->messages);
	free(param);
	return numseq;
}

/* ---------------- get_sequence_stats (io.c:52-300) through the reference's own function ----------------
 * out[0..4] background, [5] expected_5, [6] expected_3, [7] mean_5, [8] stdev_5, [9] mean_3, [10] stdev_3,
 * [11] average_length, [12] max_seq_len */
int refh_sequence_stats(struct parameters* param, const char* path, int num_query, double* out)
{
	struct read_info** ri = NULL;
	struct sequence_stats_info* ssi;
	char* files[1];
	char** keep_files = param->infile;
	int keep_n = param->infiles, keep_q = param->num_query, i;
	files[0] = (char*)path;
	param->infile = files; param->infiles = 1; param->num_query = num_query;
	ri = malloc_read_info(ri, num_query);
	ssi = get_sequence_stats(param, ri, 0);
	free_read_info(ri, num_query);
	param->infile = keep_files; param->infiles = keep_n; param->num_query = keep_q;
	if(!ssi) return -1;
	for(i = 0; i < 5; i++) out[i] = ssi->background[i];
	out[5] = ssi->expected_5_len; out[6] = ssi->expected_3_len;
	out[7] = ssi->mean_5_len; out[8] = ssi->stdev_5_len; out[9] = ssi->mean_3_len; out[10] = ssi->stdev_3_len;
	out[11] = ssi->average_length; out[12] = ssi->max_seq_len;
	free(ssi);
	return 0;
}
