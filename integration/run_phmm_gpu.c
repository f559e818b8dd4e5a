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
#include <pthread.h>
#include <time.h>
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

/* Content-keyed model cache.  A hit is only accepted after a memcmp of the whole flattened table blob
 * (the FNV key alone could collide); entries are evicted round robin, never while a caller holds them:
 * the arch-comparison path builds its models uncached (up to MAX_NUM_ARCH of them are alive at once). */
#define MODEL_CACHE 24
static struct { unsigned long long key; int max_len; size_t blob_n; unsigned char* blob; tdg_model* m; } g_models[MODEL_CACHE];
static int g_model_next = 0;

static int fail_msg(struct parameters* param, const char* what)
{
	snprintf(param->errmsg, kslibERRBUFSIZE, "%s: %s", what, tdg_last_error());
	fprintf(stderr, "tagdust_b200: %s: %s\n", what, tdg_last_error());
	return kslFAIL;
}

static pthread_mutex_t g_ctx_mu = PTHREAD_MUTEX_INITIALIZER;

static int init_ctx_locked(void)
{
	int rc = TDG_OK;
	pthread_mutex_lock(&g_ctx_mu);
	if (!g_ctx) {
		int n = 0; /* all visible devices */
		const char* e = getenv("TDG_NUM_DEVICES");
		if (e) n = atoi(e);
		rc = tdg_init(n, NULL, &g_ctx);
	}
	pthread_mutex_unlock(&g_ctx_mu);
	return rc;
}

static int ensure_ctx(struct parameters* param)
{
	if (init_ctx_locked() != TDG_OK) return fail_msg(param, "tdg_init");
	return kslOK;
}

/* CUDA start-up takes ~1.7 s on a B200 box: the controller starts it on a thread of its own while the
 * reference's host set-up (sequence statistics, simulated calibration reads) runs.  A failure here is
 * silent; the first real use retries on the calling thread and reports it. */
static void* warmup_fn(void* arg) { (void)arg; (void)init_ctx_locked(); return NULL; }
static pthread_t g_warm_thread;
static int g_warm_state = 0;   /* 0 not started, 1 running / joinable, 2 joined */
void tdg_shim_warmup(void)
{
	if (g_warm_state) return;
	if (pthread_create(&g_warm_thread, NULL, warmup_fn, NULL) == 0) g_warm_state = 1;
	else g_warm_state = 2;
}
/* before exit(): never tear the process down while the CUDA start-up is still in flight on the other thread */
void tdg_shim_warmup_join(void)
{
	if (g_warm_state == 1) { pthread_join(g_warm_thread, NULL); g_warm_state = 2; }
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

/* struct model_bag -> flat tables (segment -> hmm -> column order), all in one allocation */
struct flat_model {
	int S, H, C;
	unsigned char* blob; size_t blob_n;   /* everything below points into blob */
	char* seg_type; int32_t* nh; int32_t* nc; float* skip; float* bg; float* tr; float* me; float* ie; float* sM; float* sI;
	int32_t* label; float* T; int32_t* avg;
};

static int flatten_model(struct model_bag* mb, struct parameters* param, struct flat_model* fm)
{
	int S = mb->num_models, H = mb->total_hmm_num, C = 0, j, f, g, k, c = 0;
	size_t o = 0;
	for (j = 0; j < S; j++) C += mb->model[j]->num_hmms * mb->model[j]->hmms[0]->num_columns;
	fm->S = S; fm->H = H; fm->C = C;
#define TAKE(T, n) (o = (o + 7) & ~(size_t)7, o += sizeof(T) * (size_t)(n), o - sizeof(T) * (size_t)(n))
	const size_t o_nh = TAKE(int32_t, S), o_nc = TAKE(int32_t, S), o_skip = TAKE(float, S), o_bg = TAKE(float, 5),
	             o_tr = TAKE(float, C * 9), o_me = TAKE(float, C * 5), o_ie = TAKE(float, C * 5), o_sM = TAKE(float, C),
	             o_sI = TAKE(float, C), o_label = TAKE(int32_t, H), o_T = TAKE(float, (size_t)H * H), o_avg = TAKE(int32_t, 1),
	             o_type = TAKE(char, S + 1);
#undef TAKE
	fm->blob_n = o;
	fm->blob = calloc(1, o);   /* zeroed: alignment gaps take part in the hash / memcmp */
	if (!fm->blob) return kslEMEM;
	fm->nh = (int32_t*)(fm->blob + o_nh); fm->nc = (int32_t*)(fm->blob + o_nc); fm->skip = (float*)(fm->blob + o_skip);
	fm->bg = (float*)(fm->blob + o_bg); fm->tr = (float*)(fm->blob + o_tr); fm->me = (float*)(fm->blob + o_me);
	fm->ie = (float*)(fm->blob + o_ie); fm->sM = (float*)(fm->blob + o_sM); fm->sI = (float*)(fm->blob + o_sI);
	fm->label = (int32_t*)(fm->blob + o_label); fm->T = (float*)(fm->blob + o_T); fm->avg = (int32_t*)(fm->blob + o_avg);
	fm->seg_type = (char*)(fm->blob + o_type);
	for (k = 0; k < 5; k++) fm->bg[k] = mb->model[0]->background_nuc_frequency[k];
	for (j = 0; j < S; j++) {
		struct model* m = mb->model[j];
		fm->seg_type[j] = param->read_structure->type[j];
		fm->nh[j] = m->num_hmms; fm->nc[j] = m->hmms[0]->num_columns; fm->skip[j] = m->skip;
		for (f = 0; f < m->num_hmms; f++)
			for (g = 0; g < m->hmms[f]->num_columns; g++) {
				struct hmm_column* col = m->hmms[f]->hmm_column[g];
				for (k = 0; k < 9; k++) fm->tr[c * 9 + k] = col->transition[k];
				for (k = 0; k < 5; k++) { fm->me[c * 5 + k] = col->m_emit[k]; fm->ie[c * 5 + k] = col->i_emit[k]; }
				fm->sM[c] = m->silent_to_M[f][g]; fm->sI[c] = m->silent_to_I[f][g];
				c++;
			}
	}
	for (j = 0; j < H; j++) { fm->label[j] = mb->label[j]; for (k = 0; k < H; k++) fm->T[j * H + k] = mb->transition_matrix[j][k]; }
	*fm->avg = mb->average_raw_length;
	return kslOK;
}

static tdg_model* create_from_flat(const struct flat_model* fm, int max_len)
{
	tdg_model_desc d;
	tdg_model* out = NULL;
	d.num_segments = fm->S; d.total_hmms = fm->H; d.total_columns = fm->C; d.average_raw_length = *fm->avg;
	d.seg_type = fm->seg_type; d.seg_num_hmms = fm->nh; d.seg_num_cols = fm->nc; d.seg_skip = fm->skip; d.background = fm->bg;
	d.transition = fm->tr; d.m_emit = fm->me; d.i_emit = fm->ie; d.silent_to_M = fm->sM; d.silent_to_I = fm->sI; d.label = fm->label;
	d.transition_matrix = fm->T;
	if (tdg_model_create(g_ctx, &d, max_len, &out) != TDG_OK) return NULL;
	return out;
}

/* uncached: the caller owns the model and destroys it (arch comparison: all models of ab->archs[] are alive at once) */
static tdg_model* build_model_uncached(struct model_bag* mb, struct parameters* param, int max_len)
{
	struct flat_model fm;
	tdg_model* out;
	if (ensure_ctx(param) != kslOK) return NULL;
	if (flatten_model(mb, param, &fm) != kslOK) return NULL;
	out = create_from_flat(&fm, max_len);
	free(fm.blob);
	return out;
}

static tdg_model* get_model_len(struct model_bag* mb, struct parameters* param, int want_len)
{
	struct flat_model fm;
	int k;
	if (ensure_ctx(param) != kslOK) return NULL;
	if (flatten_model(mb, param, &fm) != kslOK) return NULL;
	const int max_len = want_len;
	const unsigned long long key = fnv(1469598103934665603ULL, fm.blob, fm.blob_n);
	tdg_model* out = NULL;
	for (k = 0; k < MODEL_CACHE && !out; k++)
		/* same tables AND same length: a model sized for the long calibration reads would make every
		 * wave of the real run smaller (scratch per read grows with max_len) */
		if (g_models[k].m && g_models[k].key == key && g_models[k].max_len == max_len && g_models[k].blob_n == fm.blob_n &&
		    memcmp(g_models[k].blob, fm.blob, fm.blob_n) == 0) out = g_models[k].m;
	if (out) { free(fm.blob); return out; }
	out = create_from_flat(&fm, max_len);
	if (!out) { free(fm.blob); return NULL; }
	if (g_models[g_model_next].m) { tdg_model_destroy(g_models[g_model_next].m); free(g_models[g_model_next].blob); }
	g_models[g_model_next].m = out; g_models[g_model_next].key = key; g_models[g_model_next].max_len = max_len;
	g_models[g_model_next].blob = fm.blob; g_models[g_model_next].blob_n = fm.blob_n;
	g_model_next = (g_model_next + 1) % MODEL_CACHE;
	return out;
}

/* struct fasta (get_fasta, io.c:1893-1998) -> tdg_refset; one reference file per run, cached by pointer */
static tdg_refset* g_refset = NULL;
static const struct fasta* g_refset_src = NULL;
tdg_refset* tdg_shim_get_refset(struct fasta* f, struct parameters* param)
{
	if (ensure_ctx(param) != kslOK) return NULL;
	if (g_refset && g_refset_src == f) return g_refset;
	if (g_refset) { tdg_refset_destroy(g_refset); g_refset = NULL; }
	if (tdg_refset_create(g_ctx, f->string, (const int32_t*)f->s_index, f->numseq, &g_refset) != TDG_OK) return NULL;
	g_refset_src = f;
	return g_refset;
}

tdg_model* tdg_shim_get_model_len(struct model_bag* mb, struct parameters* param, int max_len)
{
	return get_model_len(mb, param, max_len);
}

tdg_model* tdg_shim_get_model(struct model_bag* mb, struct parameters* param)
{
	return get_model_len(mb, param, mb->current_dyn_length);   /* >= max_seq_len + 10 (barcode_hmm.c:5778) */
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

static int by_length(const void* a, const void* b)
{
	const struct read_info* x = *(const struct read_info* const*)a;
	const struct read_info* y = *(const struct read_info* const*)b;
	return (x->len > y->len) - (x->len < y->len);
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

static double wall_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static int run_pHMM_impl(struct arch_bag* ab, struct model_bag* mb, struct read_info** ri, struct parameters* param,
                         struct fasta* reference_fasta, int numseq, int mode);

int run_pHMM(struct arch_bag* ab, struct model_bag* mb, struct read_info** ri, struct parameters* param,
             struct fasta* reference_fasta, int numseq, int mode)
{
	const double t0 = wall_s();
	const int rc = run_pHMM_impl(ab, mb, ri, param, reference_fasta, numseq, mode);
	if (getenv("TDG_VERBOSE"))
		fprintf(stderr, "tagdust_b200: run_pHMM mode %d, %d reads: %.3f s (called %.3f s after process start)\n", mode, numseq,
		        wall_s() - t0, t0 - (double)0);
	return rc;
}

static int run_pHMM_impl(struct arch_bag* ab, struct model_bag* mb, struct read_info** ri, struct parameters* param,
                         struct fasta* reference_fasta, int numseq, int mode)
{
	int i, j;
	{
		const double t0 = wall_s();
		const int first = (g_ctx == NULL);
		if (ensure_ctx(param) != kslOK) return kslFAIL;
		if (first && getenv("TDG_VERBOSE")) fprintf(stderr, "tagdust_b200: CUDA context + library start-up: %.3f s\n", wall_s() - t0);
	}

	if (mode == MODE_ARCH_COMP) {
		if (!ab) return kslFAIL;
		tdg_model** models = calloc(ab->num_arch, sizeof(tdg_model*));
		float* post = malloc(sizeof(float) * ab->num_arch);
		int rc = kslOK;
		/* uncached, owned here: test_architectures allows up to MAX_NUM_ARCH (100) candidates, all alive until
		 * tdg_arch_compare returns -- more than the cache holds */
		for (i = 0; i < ab->num_arch && rc == kslOK; i++) {
			models[i] = build_model_uncached(ab->archs[i], param, ab->archs[i]->current_dyn_length);   /* NB: seg types come from
			                                                 param->read_structure; only needed by extraction, not by backward() */
			if (!models[i]) rc = fail_msg(param, "tdg_model_create");
		}
		if (rc == kslOK && load_batch(param, ri, numseq, 1) != kslOK) rc = kslFAIL;
		if (rc == kslOK && tdg_arch_compare(g_ctx, models, ab->num_arch, g_batch, param->num_threads, NULL, post) != TDG_OK)
			rc = fail_msg(param, "tdg_arch_compare");
		if (rc == kslOK) for (i = 0; i < ab->num_arch; i++) ab->arch_posterior[i] = post[i];
		for (i = 0; i < ab->num_arch; i++) if (models[i]) tdg_model_destroy(models[i]);
		free(models); free(post);
		return rc;
	}
	if (mode != MODE_GET_LABEL && mode != MODE_GET_PROB) return kslFAIL; /* MODE_TRAIN: no live caller */

	tdg_model* m = tdg_shim_get_model(mb, param);
	if (!m) return fail_msg(param, "tdg_model_create");
	/* MODE_GET_PROB (threshold calibration, calibrateQ.c:136): the simulated reads have a geometric length
	 * tail (150 .. ~2000 nt at cfg2) and a warp walks its 32 reads to the longest of them, so the reads go to
	 * the GPU sorted by length; results are per read and come back through the same permutation.  Only
	 * ri->mapq and ri->bar_prob are observable after this call (labels are overwritten or freed), so the
	 * label DP is skipped. */
	struct read_info** order = ri;
	if (mode == MODE_GET_PROB && numseq > 1) {
		order = malloc(sizeof(struct read_info*) * numseq);
		memcpy(order, ri, sizeof(struct read_info*) * numseq);
		qsort(order, numseq, sizeof(struct read_info*), by_length);
	}
	if (mode == MODE_GET_PROB) {
		/* sorted chunks, each scored with a model sized for the chunk's longest read: the scratch per read
		 * (and with it the wave size) follows the length class instead of the longest simulated read */
		const int CH = 65536;
		int c0, rc = kslOK;
		if (numseq <= 0) return kslOK;
		tdg_run_params rq;
		memset(&rq, 0, sizeof rq);
		rq.confidence_threshold = param->confidence_threshold; rq.minlen = param->minlen;
		rq.matchstart = param->matchstart; rq.matchend = param->matchend; rq.dust = 0; rq.want_labels = 0; rq.want_spans = 0;
		if (ensure_batch(param, numseq < CH ? numseq : CH, order[numseq - 1]->len > 1 ? order[numseq - 1]->len : 1) != kslOK) rc = kslFAIL;
		for (c0 = 0; c0 < numseq && rc == kslOK; c0 += CH) {
			const int n = numseq - c0 < CH ? numseq - c0 : CH;
			const double tc = wall_s();
			int ml = (order[c0 + n - 1]->len + 10 + 63) / 64 * 64;
			tdg_result rs;
			tdg_model* mc;
			if (param->matchstart != -1 || param->matchend != -1) ml = mb->current_dyn_length;
			if (ml > mb->current_dyn_length) ml = mb->current_dyn_length;
			mc = get_model_len(mb, param, ml);
			if (!mc) { rc = fail_msg(param, "tdg_model_create"); break; }
			if (tdg_batch_clear(g_batch) != TDG_OK ||
			    tdg_batch_append_records(g_batch, n, (const void* const*)(order + c0), offsetof(struct read_info, seq),
			                             offsetof(struct read_info, len)) != TDG_OK) { rc = fail_msg(param, "tdg_batch_append_records"); break; }
			if (tdg_run(g_ctx, mc, TDG_MODE_GET_PROB, &rq, g_batch, &rs) != TDG_OK) { rc = fail_msg(param, "tdg_run"); break; }
			for (i = 0; i < n; i++) { order[c0 + i]->mapq = rs.mapq[i]; order[c0 + i]->bar_prob = rs.bar_prob[i]; }
			if (getenv("TDG_VERBOSE")) fprintf(stderr, "tagdust_b200:   calibration chunk %d reads, max length %d: %.3f s\n", n, order[c0 + n - 1]->len, wall_s() - tc);
		}
		if (order != ri) free(order);
		return rc;
	}
	if (load_batch(param, order, numseq, 1) != kslOK) { if (order != ri) free(order); return kslFAIL; }

	tdg_run_params rp;
	memset(&rp, 0, sizeof rp);
	rp.confidence_threshold = param->confidence_threshold;
	rp.minlen = param->minlen;
	rp.matchstart = param->matchstart;
	rp.matchend = param->matchend;
	/* with -ref the reference's order is extract -> match_to_reference -> dust (barcode_hmm.c:2345-2354); the library
	 * keeps that order on the device (k_artifact) and reproduces the thread slicing of this call */
	rp.dust = param->dust;
	rp.want_labels = (mode == MODE_GET_LABEL);
	rp.want_spans = 0;
	if (mode == MODE_GET_LABEL && reference_fasta) {
		rp.refset = tdg_shim_get_refset(reference_fasta, param);
		if (!rp.refset) { if (order != ri) free(order); return fail_msg(param, "tdg_refset_create"); }
		rp.filter_error = param->filter_error;
		rp.slice_threads = param->num_threads;
	}
	tdg_result res;
	if (tdg_run(g_ctx, m, mode == MODE_GET_LABEL ? TDG_MODE_GET_LABEL : TDG_MODE_GET_PROB, &rp, g_batch, &res) != TDG_OK)
	{
		if (order != ri) free(order);
		return fail_msg(param, "tdg_run");
	}
	for (i = 0; i < numseq; i++) {
		struct read_info* r = ri[i];
		const uint8_t* lab = res.labels + (size_t)i * res.label_stride;
		int wlen = r->len;
		if (param->matchstart != -1 || param->matchend != -1) wlen = param->matchend - param->matchstart;
		r->mapq = res.mapq[i];
		for (j = 0; j <= wlen; j++) r->labels[j] = (char)lab[j];
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
	return kslOK;
}
