/* run_phmm_gpu.c -- the reference-side binding: a drop-in definition of
 *
 *     int run_pHMM(struct arch_bag* ab, struct model_bag* mb, struct read_info** ri,
 *                  struct parameters* param, struct fasta* reference_fasta, int numseq, int mode);
 *
 * (declared barcode_hmm.h:342, defined barcode_hmm.c:1895) that routes the three live modes
 * through libtagdust_b200.so.  It is compiled against the reference's own headers and linked
 * into the reference's executable in place of the pthread fan-out; nothing else of the
 * reference changes (CLI, FASTQ I/O, calibration driver, architecture detection, output
 * naming and logs stay the reference's C).  See INTEGRATION.md.
 *
 * Build: integration/Makefile (needs the reference tree for its headers; the resulting
 * binaries live in integration/_build/, git-ignored).
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kslib.h"
#include "tagdust2.h"
#include "interface.h"
#include "misc.h"
#include "io.h"
#include "barcode_hmm.h"

#include "tagdust_b200.h"
#include "shim.h"

static tdg_context* g_ctx = NULL;
static tdg_batch* g_batch = NULL;
static int g_batch_reads = 0, g_batch_len = 0;

#define MODEL_CACHE 8
static struct { unsigned long long key; int max_len; tdg_model* m; } g_models[MODEL_CACHE];
static int g_model_next = 0;

static int fail_msg(struct parameters* param, const char* what)
{
	snprintf(param->errmsg, kslibERRBUFSIZE, "%s: %s", what, tdg_last_error());
	fprintf(stderr, "tagdust_b200: %s: %s\n", what, tdg_last_error());
	return kslFAIL;
}

static int ensure_ctx(struct parameters* param)
{
	if (g_ctx) return kslOK;
	int n = 0; /* all visible devices */
	const char* e = getenv("TDG_NUM_DEVICES");
	if (e) n = atoi(e);
	if (tdg_init(n, NULL, &g_ctx) != TDG_OK) return fail_msg(param, "tdg_init");
	return kslOK;
}

/* FNV-1a over the flattened tables: a model_bag is rebuilt (same pointer or not) whenever the
 * threshold calibration / long-read path re-runs init_model_bag, so key on content. */
static unsigned long long fnv(unsigned long long h, const void* p, size_t n)
{
	const unsigned char* c = p;
	for (size_t i = 0; i < n; i++) { h ^= c[i]; h *= 1099511628211ULL; }
	return h;
}

tdg_context* tdg_shim_context(struct parameters* param)
{
	return ensure_ctx(param) == kslOK ? g_ctx : NULL;
}

/* struct model_bag -> tdg_model (flatten in segment -> hmm -> column order) */
tdg_model* tdg_shim_get_model(struct model_bag* mb, struct parameters* param)
{
	if (ensure_ctx(param) != kslOK) return NULL;
	int S = mb->num_models, H = mb->total_hmm_num, C = 0, j, f, g, k, c = 0;
	for (j = 0; j < S; j++) C += mb->model[j]->num_hmms * mb->model[j]->hmms[0]->num_columns;
	char* seg_type = malloc(S + 1);
	int32_t* nh = malloc(sizeof(int32_t) * S); int32_t* nc = malloc(sizeof(int32_t) * S);
	float* skip = malloc(sizeof(float) * S);
	float bg[5];
	float* tr = malloc(sizeof(float) * C * 9); float* me = malloc(sizeof(float) * C * 5); float* ie = malloc(sizeof(float) * C * 5);
	float* sM = malloc(sizeof(float) * C); float* sI = malloc(sizeof(float) * C);
	int32_t* label = malloc(sizeof(int32_t) * H); float* T = malloc(sizeof(float) * H * H);
	for (k = 0; k < 5; k++) bg[k] = mb->model[0]->background_nuc_frequency[k];
	for (j = 0; j < S; j++) {
		struct model* m = mb->model[j];
		seg_type[j] = param->read_structure->type[j];
		nh[j] = m->num_hmms; nc[j] = m->hmms[0]->num_columns; skip[j] = m->skip;
		for (f = 0; f < m->num_hmms; f++)
			for (g = 0; g < m->hmms[f]->num_columns; g++) {
				struct hmm_column* col = m->hmms[f]->hmm_column[g];
				for (k = 0; k < 9; k++) tr[c * 9 + k] = col->transition[k];
				for (k = 0; k < 5; k++) { me[c * 5 + k] = col->m_emit[k]; ie[c * 5 + k] = col->i_emit[k]; }
				sM[c] = m->silent_to_M[f][g]; sI[c] = m->silent_to_I[f][g];
				c++;
			}
	}
	seg_type[S] = 0;
	for (j = 0; j < H; j++) { label[j] = mb->label[j]; for (k = 0; k < H; k++) T[j * H + k] = mb->transition_matrix[j][k]; }
	const int max_len = mb->current_dyn_length;   /* >= max_seq_len + 10 (barcode_hmm.c:5778) */
	unsigned long long key = 1469598103934665603ULL;
	key = fnv(key, seg_type, S); key = fnv(key, nh, 4 * S); key = fnv(key, nc, 4 * S); key = fnv(key, skip, 4 * S);
	key = fnv(key, bg, 20); key = fnv(key, tr, 4 * C * 9); key = fnv(key, me, 4 * C * 5); key = fnv(key, ie, 4 * C * 5);
	key = fnv(key, sM, 4 * C); key = fnv(key, sI, 4 * C); key = fnv(key, label, 4 * H); key = fnv(key, T, 4 * H * H);
	key = fnv(key, &mb->average_raw_length, sizeof(int));
	tdg_model* out = NULL;
	for (k = 0; k < MODEL_CACHE; k++)
		/* same tables AND same length: a model sized for the long calibration reads would make every
		 * wave of the real run smaller (scratch per read grows with max_len) */
		if (g_models[k].m && g_models[k].key == key && g_models[k].max_len == max_len) out = g_models[k].m;
	if (!out) {
		tdg_model_desc d;
		d.num_segments = S; d.total_hmms = H; d.total_columns = C; d.average_raw_length = mb->average_raw_length;
		d.seg_type = seg_type; d.seg_num_hmms = nh; d.seg_num_cols = nc; d.seg_skip = skip; d.background = bg;
		d.transition = tr; d.m_emit = me; d.i_emit = ie; d.silent_to_M = sM; d.silent_to_I = sI; d.label = label;
		d.transition_matrix = T;
		if (tdg_model_create(g_ctx, &d, max_len, &out) != TDG_OK) out = NULL;
		else {
			if (g_models[g_model_next].m) tdg_model_destroy(g_models[g_model_next].m);
			g_models[g_model_next].m = out; g_models[g_model_next].key = key; g_models[g_model_next].max_len = max_len;
			g_model_next = (g_model_next + 1) % MODEL_CACHE;
		}
	}
	free(seg_type); free(nh); free(nc); free(skip); free(tr); free(me); free(ie); free(sM); free(sI); free(label); free(T);
	return out;
}

static int ensure_batch(struct parameters* param, int numseq, int max_len)
{
	if (g_batch && g_batch_reads >= numseq && g_batch_len >= max_len) return tdg_batch_clear(g_batch) == TDG_OK ? kslOK : kslFAIL;
	if (g_batch) tdg_batch_destroy(g_batch);
	g_batch = NULL;
	g_batch_reads = numseq > g_batch_reads ? numseq : g_batch_reads;
	g_batch_len = max_len > g_batch_len ? max_len : g_batch_len;
	if (tdg_batch_create(g_ctx, g_batch_reads, g_batch_len, &g_batch) != TDG_OK) return fail_msg(param, "tdg_batch_create");
	return kslOK;
}

static int load_batch(struct parameters* param, struct read_info** ri, int numseq, int max_len)
{
	int i, ml = max_len;
	for (i = 0; i < numseq; i++) if (ri[i]->len > ml) ml = ri[i]->len;
	if (ensure_batch(param, numseq, ml) != kslOK) return kslFAIL;
	if (tdg_batch_append_records(g_batch, numseq, (const void* const*)ri, offsetof(struct read_info, seq),
	                             offsetof(struct read_info, len)) != TDG_OK)
		return fail_msg(param, "tdg_batch_append_records");
	return kslOK;
}

int run_pHMM(struct arch_bag* ab, struct model_bag* mb, struct read_info** ri, struct parameters* param,
             struct fasta* reference_fasta, int numseq, int mode)
{
	int i, j;
	if (ensure_ctx(param) != kslOK) return kslFAIL;

	if (mode == MODE_ARCH_COMP) {
		if (!ab) return kslFAIL;
		tdg_model** models = malloc(sizeof(tdg_model*) * ab->num_arch);
		int max_len = 0;
		for (i = 0; i < ab->num_arch; i++) {
			models[i] = tdg_shim_get_model(ab->archs[i], param);   /* NB: seg types come from param->read_structure; only
			                                                 needed by extraction, not by backward() */
			if (!models[i]) { free(models); return fail_msg(param, "tdg_model_create"); }
			if (ab->archs[i]->current_dyn_length > max_len) max_len = ab->archs[i]->current_dyn_length;
		}
		if (load_batch(param, ri, numseq, 1) != kslOK) { free(models); return kslFAIL; }
		float* post = malloc(sizeof(float) * ab->num_arch);
		if (tdg_arch_compare(g_ctx, models, ab->num_arch, g_batch, param->num_threads, NULL, post) != TDG_OK) {
			free(models); free(post);
			return fail_msg(param, "tdg_arch_compare");
		}
		for (i = 0; i < ab->num_arch; i++) ab->arch_posterior[i] = post[i];
		free(models); free(post);
		return kslOK;
	}
	if (mode != MODE_GET_LABEL && mode != MODE_GET_PROB) return kslFAIL; /* MODE_TRAIN: no live caller */

	tdg_model* m = tdg_shim_get_model(mb, param);
	if (!m) return fail_msg(param, "tdg_model_create");
	if (load_batch(param, ri, numseq, 1) != kslOK) return kslFAIL;

	tdg_run_params rp;
	rp.confidence_threshold = param->confidence_threshold;
	rp.minlen = param->minlen;
	rp.matchstart = param->matchstart;
	rp.matchend = param->matchend;
	/* with -ref the reference order is extract -> match_to_reference -> dust (barcode_hmm.c:2345-2354):
	 * keep dust on the host then, after the artifact filter */
	rp.dust = reference_fasta ? 0 : param->dust;
	rp.want_labels = 1;
	tdg_result res;
	if (tdg_run(g_ctx, m, mode == MODE_GET_LABEL ? TDG_MODE_GET_LABEL : TDG_MODE_GET_PROB, &rp, g_batch, &res) != TDG_OK)
		return fail_msg(param, "tdg_run");

	for (i = 0; i < numseq; i++) {
		struct read_info* r = ri[i];
		const uint8_t* lab = res.labels + (size_t)i * res.label_stride;
		int wlen = r->len;
		if (param->matchstart != -1 || param->matchend != -1) wlen = param->matchend - param->matchstart;
		r->mapq = res.mapq[i];
		for (j = 0; j <= wlen; j++) r->labels[j] = (char)lab[j];
		if (mode == MODE_GET_PROB) { r->bar_prob = res.bar_prob[i]; continue; }
		r->bar_prob = 100;                                   /* barcode_hmm.c:2343 */
		r->read_type = res.read_type[i];
		if (res.barcode[i] != -1 || res.extracted[i]) { if (res.barcode[i] != -1) r->barcode = res.barcode[i]; }
		if (res.fingerprint[i] != -1) r->fingerprint = res.fingerprint[i];
		if (res.extracted[i]) {
			/* make_extracted_read (barcode_hmm.c:3325-3356): qualities never leave the host */
			int s_pos = 0;
			for (j = 0; j < r->len; j++) {
				const int c2 = mb->label[(int)r->labels[j + 1]] & 0xFFFF;
				if (param->read_structure->type[c2] == 'R') { r->seq[s_pos] = r->seq[j]; r->qual[s_pos] = r->qual[j]; }
				else { r->seq[s_pos] = 65; r->qual[s_pos] = 65; }
				s_pos++;
			}
			r->len = s_pos;
		}
		r->qual[r->len] = 0;                                 /* extract_reads :3308 */
	}
	if (mode == MODE_GET_LABEL && reference_fasta) {
		struct thread_data td;
		memset(&td, 0, sizeof td);
		td.ri = ri; td.mb = mb; td.param = param; td.fasta = reference_fasta; td.start = 0; td.end = numseq; td.numseq = numseq;
		ri = match_to_reference(&td);
		if (param->dust) ri = dust_sequences(&td);
	}
	return kslOK;
}
