/* stats_fast.c -- the third reference-side binding: a drop-in definition of
 *
 *     struct sequence_stats_info* get_sequence_stats(struct parameters* param, struct read_info** ri, int file_num);
 *                                                                                   io.h, io.c:52-300
 *
 * (callers: hmm_controller_multiple barcode_hmm.c:185, test_architectures test_architectures.c:114 --
 * once per candidate architecture).  The sums come from tdg_sequence_stats() of libtagdust_b200.so
 * (block reads, multi-threaded counting, no per-read malloc); the derived fields, warnings and their
 * wording follow io.c:216-270.  Raw sums are cached per (file, 5' sequence, 3' sequence) so that the
 * 64 candidate architectures of an -arch file do not re-read the input 64 times.
 * SAM/BAM input falls through to the reference's own function (dlsym RTLD_NEXT).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kslib.h"
#include "tagdust2.h"
#include "interface.h"
#include "nuc_code.h"
#include "misc.h"
#include "io.h"

#include "tagdust_b200_stream.h"

typedef struct sequence_stats_info* (*stats_fn)(struct parameters*, struct read_info**, int);

#define STATS_CACHE 16
static struct { char* key; tdg_seq_stats st; } g_cache[STATS_CACHE];
static int g_cache_next = 0;

static int is_sam(const char* f)
{
	size_t n = strlen(f);
	const char* suf[] = {".sam", ".bam", ".sam.gz", ".bam.gz"};
	int k;
	for (k = 0; k < 4; k++) { size_t m = strlen(suf[k]); if (n >= m && !strcmp(f + n - m, suf[k])) return 1; }
	return 0;
}

static void warn(struct parameters* param, const char* msg)
{
	snprintf(param->buffer, MSG_BUFFER_SIZE, "%s", msg);
	param->messages = append_message(param->messages, param->buffer);
}

struct sequence_stats_info* get_sequence_stats(struct parameters* param, struct read_info** ri, int file_num)
{
	const char* path = param->infile[file_num];
	struct read_structure* rs = param->read_structure;
	const char* five_s = NULL; const char* three_s = NULL;
	unsigned char five[1024], three[1024];
	int five_len = 0, three_len = 0, i;
	tdg_seq_stats st;
	struct sequence_stats_info* ssi;
	char key[4096];
	double sum;

	if (is_sam(path) || getenv("TDG_REFERENCE_CONTROLLER")) {
		stats_fn ref = (stats_fn)dlsym(RTLD_NEXT, "get_sequence_stats");
		return ref ? ref(param, ri, file_num) : NULL;
	}
	/* the reference still runs io_handler here: its "Cannot find input file" exit and the
	 * param->fasta / gzipped flags it leaves behind are observable */
	{
		FILE* fh = NULL;
		fh = io_handler(fh, file_num, param);
		if (fh) pclose(fh);
	}
	if (rs->type[0] == 'P') five_s = rs->sequence_matrix[0][0];
	if (rs->type[rs->num_segments - 1] == 'P') three_s = rs->sequence_matrix[rs->num_segments - 1][0];
	if (five_s) { five_len = (int)strlen(five_s); for (i = 0; i < five_len && i < 1024; i++) five[i] = nuc_code[(int)five_s[i]]; }
	if (three_s) { three_len = (int)strlen(three_s); for (i = 0; i < three_len && i < 1024; i++) three[i] = nuc_code[(int)three_s[i]]; }

	snprintf(key, sizeof key, "%s|%d|%s|%s", path, param->num_query, five_s ? five_s : "", three_s ? three_s : "");
	for (i = 0; i < STATS_CACHE; i++)
		if (g_cache[i].key && !strcmp(g_cache[i].key, key)) break;
	if (i < STATS_CACHE) st = g_cache[i].st;
	else {
		if (tdg_sequence_stats(path, param->fasta ? 1 : 0, param->num_query, five_len ? five : NULL, five_len,
		                       three_len ? three : NULL, three_len, param->num_threads, &st) != TDG_OK) {
			snprintf(param->buffer, MSG_BUFFER_SIZE, "%s\n", tdg_last_error());
			param->messages = append_message(param->messages, param->buffer);
			free_param(param);
			exit(EXIT_FAILURE);
		}
		free(g_cache[g_cache_next].key);
		g_cache[g_cache_next].key = strdup(key);
		g_cache[g_cache_next].st = st;
		g_cache_next = (g_cache_next + 1) % STATS_CACHE;
	}

	ssi = malloc(sizeof *ssi);
	ssi->expected_5_len = five_len;
	ssi->expected_3_len = three_len;
	ssi->max_seq_len = st.max_seq_len;
	ssi->average_length = st.sum_len;
	for (i = 0; i < 5; i++) ssi->background[i] = 1.0 + st.base_count[i];

	/* 5' / 3' partial segments: mean and standard deviation of the matched length (io.c:216-262) */
	if (five_len) {
		if (st.five_s0 <= 1) {
			warn(param, "WARNING: there seems to e not a single read containing the 5' partial sequence.\n");
			ssi->mean_5_len = ssi->expected_5_len;
			ssi->stdev_5_len = 1.0;
		} else {
			ssi->mean_5_len = st.five_s1 / st.five_s0;
			ssi->stdev_5_len = sqrt((st.five_s0 * st.five_s2 - pow(st.five_s1, 2.0)) / (st.five_s0 * (st.five_s0 - 1.0)));
			if (!ssi->stdev_5_len) ssi->stdev_5_len = 10000.0;
			if (ssi->mean_5_len <= 1) warn(param, "WARNING: 5' partial segment seems not to be present in the data (length < 1).\n");
		}
	} else { ssi->mean_5_len = -1.0; ssi->stdev_5_len = -1.0; }
	if (three_len) {
		if (st.three_s0 <= 1) {
			warn(param, "WARNING: 3' partial segment seems not to be present in the data.\n");
			ssi->mean_3_len = ssi->expected_3_len;
			ssi->stdev_3_len = 1.0;
		} else {
			ssi->mean_3_len = st.three_s1 / st.three_s0;
			ssi->stdev_3_len = sqrt((st.three_s0 * st.three_s2 - pow(st.three_s1, 2.0)) / (st.three_s0 * (st.three_s0 - 1.0)));
			if (!ssi->stdev_3_len) ssi->stdev_3_len = 10000.0;
			if (ssi->mean_3_len <= 1) warn(param, "WARNING: 3' partial segment seems not to be present in the data (length < 1).\n");
		}
	} else { ssi->mean_3_len = -1.0; ssi->stdev_3_len = -1.0; }

	if (param->matchstart != -1 || param->matchend != -1) ssi->average_length = (param->matchend - param->matchstart) * st.total_read;
	ssi->average_length = (int)floor((double)ssi->average_length / (double)st.total_read + 0.5);
	sum = 0.0;
	for (i = 0; i < 5; i++) sum += ssi->background[i];
	for (i = 0; i < 5; i++) ssi->background[i] = prob2scaledprob(ssi->background[i] / sum);
	return ssi;
}
