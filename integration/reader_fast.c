/* reader_fast.c -- the fifth reference-side binding: drop-in definitions of
 *
 *     FILE* io_handler(FILE* file, int file_num, struct parameters* param);                      io.c:382-608
 *     int   read_fasta_fastq(struct read_info** ri, struct parameters* param, FILE* file,
 *                            int* buffer_count);                                                  io.c:1684-1815
 *
 * so that every chunk the reference's own code still reads itself -- test_architectures' sample of 100 000 reads
 * (test_architectures.c:38-42, :169-182), the read-name order check of the controller -- is parsed by
 * tdg_fastq_next() (block reads, one memchr pass, multi-threaded conversion) instead of fgets + per-character loops,
 * and only the struct read_info array the callers expect is filled here (name / seq / labels / qual allocated per read
 * exactly as io.c:1716-1790 does, because free_read_info / clear_read_info release them one by one).
 *
 * io_handler keeps the reference's behaviour (it returns the FILE* of the reference's own pipe, so pclose() by the
 * caller works as before) and remembers which input file the FILE* belongs to; read_fasta_fastq looks the FILE* up and
 * parses the file itself with its own cursor.  SAM/BAM input is not touched (read_sam_chunk reads the pipe).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
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

typedef FILE* (*io_handler_fn)(FILE*, int, struct parameters*);
typedef int (*reader_fn)(struct read_info**, struct parameters*, FILE*, int*);

#define MAX_OPEN 32
static struct { FILE* fh; char* path; tdg_fastq* rd; } g_open[MAX_OPEN];

FILE* io_handler(FILE* file, int file_num, struct parameters* param)
{
	static io_handler_fn real = NULL;
	int k, slot = -1;
	if (!real) real = (io_handler_fn)dlsym(RTLD_NEXT, "io_handler");
	file = real(file, file_num, param);
	if (!file) return file;
	/* a FILE* value can be handed out again after pclose(): the new owner replaces the old entry */
	for (k = 0; k < MAX_OPEN; k++) if (g_open[k].fh == file) slot = k;
	for (k = 0; k < MAX_OPEN && slot < 0; k++) if (!g_open[k].fh) slot = k;
	if (slot < 0) slot = 0;
	if (g_open[slot].rd) tdg_fastq_close(g_open[slot].rd);
	free(g_open[slot].path);
	g_open[slot].fh = file;
	g_open[slot].path = strdup(param->infile[file_num]);
	g_open[slot].rd = NULL;
	return file;
}

int read_fasta_fastq(struct read_info** ri, struct parameters* param, FILE* file, int* buffer_count)
{
	static reader_fn real = NULL;
	int k, slot = -1, i;
	tdg_fastq_chunk ch;
	for (k = 0; k < MAX_OPEN; k++) if (g_open[k].fh == file && g_open[k].path) slot = k;
	if (slot < 0 || param->sam || getenv("TDG_REFERENCE_READER")) {
		if (!real) real = (reader_fn)dlsym(RTLD_NEXT, "read_fasta_fastq");
		return real(ri, param, file, buffer_count);
	}
	*buffer_count = 0;
	ri = clear_read_info(ri, param->num_query);
	if (!g_open[slot].rd && tdg_fastq_open(g_open[slot].path, param->fasta ? 1 : 0, &g_open[slot].rd) != TDG_OK) {
		snprintf(param->errmsg, kslibERRBUFSIZE, "%s", tdg_last_error());
		return kslFAIL;
	}
	if (tdg_fastq_next(g_open[slot].rd, param->num_query, param->num_threads > 0 ? param->num_threads : 1, &ch) != TDG_OK) {
		/* io.c:1770-1775: "ERROR: Length of sequence and base qualities differ!." ends the run */
		snprintf(param->buffer, MSG_BUFFER_SIZE, "%s\n", tdg_last_error());
		param->messages = append_message(param->messages, param->buffer);
		free_param(param);
		exit(EXIT_FAILURE);
	}
	for (i = 0; i < ch.n; i++) {
		const int len = ch.len[i];
		const char* name = ch.names + ch.name_off[i];
		const size_t nl = strlen(name);
		struct read_info* r = ri[i];
		r->name = malloc(nl + 2);
		memcpy(r->name, name, nl + 1);
		r->seq = malloc((size_t)len + 2);
		r->labels = malloc((size_t)len + 2);
		memcpy(r->seq, ch.codes + ch.seq_off[i], (size_t)len + 1);
		memset(r->labels, 0, (size_t)len + 1);
		r->len = len;
		if (ch.qual) {
			r->qual = malloc((size_t)len + 2);
			memcpy(r->qual, ch.qual + ch.seq_off[i], (size_t)len + 1);
		}
	}
	*buffer_count = ch.n;
	return kslOK;
}
