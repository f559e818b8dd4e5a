/* emit_fast.c -- the fourth reference-side binding: drop-in definitions of the two read emitters the
 * threshold calibration draws its 400 000 test reads from,
 *
 *     int emit_read_sequence  (struct model_bag* mb, struct read_info* ri, int average_length, unsigned int* seed);
 *     int emit_random_sequence(struct model_bag* mb, struct read_info* ri, int average_length, unsigned int* seed);
 *                                                   barcode_hmm.c:2708-3046, :2599-2689; caller calibrateQ.c:88-112
 *
 * Host-only C, no GPU: the emission is a strictly serial chain of rand() draws (one global generator
 * seeded by estimateQthreshold), so it cannot be spread over threads or moved to the device without
 * changing the reads.  What can change is the cost per draw: the reference re-derives every cumulative
 * threshold it compares a draw with -- a logsum() and a double exp() per candidate, up to
 * 2 x hmms x columns of them each time the silent state of the barcode segment is left -- on every step.
 * Here the same chains are evaluated ONCE per model with the reference's own logsum()/scaledprob2prob()
 * (same call order, same float/double conversions), stored, and each step is a scan over stored floats.
 * Same rand() calls in the same order, same comparisons `r < threshold`: the emitted reads are identical
 * (tests/test_emit_host.py compares them with the reference's emitters for several architectures and seeds).
 *
 * The tables are keyed by the model_bag pointer and dropped when that model_bag is freed
 * (free_model_bag is interposed for exactly that and forwards to the reference's own function).
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kslib.h"
#include "tagdust2.h"
#include "interface.h"
#include "misc.h"
#include "io.h"
#include "barcode_hmm.h"

#ifdef RTEST
#define MY_RAND_MAX 32768u
#else
#define MY_RAND_MAX ((unsigned int)RAND_MAX)
#endif

struct seg_tab {
	int nh, nc;
	float* sil;  /* [nh*nc*2]  cumulative P(silent -> M(i,j)), P(silent -> I(i,j)) in the reference's scan order */
	float* tM;   /* [nh*nc*3]  from M: MM, +MI, +MD      (rest: MSKIP) */
	float* tI;   /* [nh*nc*2]  from I: II, +IM           (rest: ISKIP) */
	float* tD;   /* [nh*nc]    from D: DD                (rest: DM)    */
	float* eM;   /* [nh*nc*5]  cumulative match emissions  */
	float* eI;   /* [nh*nc*5]  cumulative insert emissions */
};

static struct {
	struct model_bag* mb;
	int S;
	struct seg_tab* seg;
	float bg[5];
} g_tab;

static void drop_tables(void)
{
	int s;
	for (s = 0; s < g_tab.S; s++) {
		free(g_tab.seg[s].sil); free(g_tab.seg[s].tM); free(g_tab.seg[s].tI); free(g_tab.seg[s].tD);
		free(g_tab.seg[s].eM); free(g_tab.seg[s].eI);
	}
	free(g_tab.seg);
	memset(&g_tab, 0, sizeof g_tab);
}

/* one step of the reference's chain: `sum = logsum(sum, x); threshold = scaledprob2prob(sum);`
 * (sum is a double variable holding float values, barcode_hmm.c:2724) */
static float chain(double* sum, float x)
{
	*sum = logsum(*sum, x);
	return scaledprob2prob(*sum);
}

static void build_tables(struct model_bag* mb)
{
	int s, i, j, k;
	double sum;
	drop_tables();
	g_tab.mb = mb;
	g_tab.S = mb->num_models;
	g_tab.seg = calloc(g_tab.S, sizeof(struct seg_tab));
	for (s = 0; s < g_tab.S; s++) {
		struct model* m = mb->model[s];
		struct seg_tab* t = &g_tab.seg[s];
		const int nh = m->num_hmms, nc = m->hmms[0]->num_columns;   /* len = hmms[0]->num_columns, :2756 */
		t->nh = nh; t->nc = nc;
		t->sil = malloc(sizeof(float) * nh * nc * 2);
		t->tM = malloc(sizeof(float) * nh * nc * 3);
		t->tI = malloc(sizeof(float) * nh * nc * 2);
		t->tD = malloc(sizeof(float) * nh * nc);
		t->eM = malloc(sizeof(float) * nh * nc * 5);
		t->eI = malloc(sizeof(float) * nh * nc * 5);
		sum = prob2scaledprob(0.0f);
		for (i = 0; i < nh; i++)
			for (j = 0; j < nc; j++) {
				t->sil[(i * nc + j) * 2] = chain(&sum, m->silent_to_M[i][j]);
				t->sil[(i * nc + j) * 2 + 1] = chain(&sum, m->silent_to_I[i][j]);
			}
		for (i = 0; i < nh; i++)
			for (j = 0; j < nc; j++) {
				struct hmm_column* c = m->hmms[i]->hmm_column[j];
				const int o = i * nc + j;
				sum = prob2scaledprob(0.0f);
				t->tM[o * 3] = chain(&sum, c->transition[MM]);
				t->tM[o * 3 + 1] = chain(&sum, c->transition[MI]);
				t->tM[o * 3 + 2] = chain(&sum, c->transition[MD]);
				sum = prob2scaledprob(0.0f);
				t->tI[o * 2] = chain(&sum, c->transition[II]);
				t->tI[o * 2 + 1] = chain(&sum, c->transition[IM]);
				sum = prob2scaledprob(0.0f);
				t->tD[o] = chain(&sum, c->transition[DD]);
				sum = prob2scaledprob(0.0f);
				for (k = 0; k < 5; k++) t->eM[o * 5 + k] = chain(&sum, c->m_emit[k]);
				sum = prob2scaledprob(0.0f);
				for (k = 0; k < 5; k++) t->eI[o * 5 + k] = chain(&sum, c->i_emit[k]);
			}
	}
	sum = prob2scaledprob(0.0f);
	for (k = 0; k < 5; k++) g_tab.bg[k] = chain(&sum, mb->model[0]->background_nuc_frequency[k]);
}

typedef void (*free_bag_fn)(struct model_bag*);
void free_model_bag(struct model_bag* mb)
{
	static free_bag_fn real = NULL;
	if (!real) real = (free_bag_fn)dlsym(RTLD_NEXT, "free_model_bag");
	if (mb && mb == g_tab.mb) drop_tables();
	if (real) real(mb);
}

static double draw(void) { return (float)rand() / (float)MY_RAND_MAX; }

static void reset_read(struct read_info* ri)
{
	free(ri->seq); free(ri->name); free(ri->qual); free(ri->labels);
	ri->seq = 0; ri->name = 0; ri->qual = 0; ri->labels = 0;
	ri->len = 0;
	ri->read_type = 0;
}

static void finish_read(struct read_info* ri, int len, char name0, int keep_labels)
{
	int i;
	ri->seq = realloc(ri->seq, len + 1);
	ri->seq[len] = 0;
	if (keep_labels) { ri->labels = realloc(ri->labels, len + 1); ri->labels[len] = 0; }
	else ri->labels = malloc(len + 1);
	ri->qual = malloc(len + 1);
	for (i = 0; i < len; i++) ri->qual[i] = 'B';
	ri->qual[len] = 0;
	ri->len = len;
	ri->name = malloc(2);
	ri->name[0] = name0;
	ri->name[1] = 0;
}

int emit_random_sequence(struct model_bag* mb, struct read_info* ri, int average_length, unsigned int* seed)
{
	int current_length = 0, allocated_length = 100, nuc;
	double r = draw();                                           /* :2610 */
	const double stop = 1.0 - (1.0 / (float)average_length);     /* :2647 */
	(void)seed;
	if (mb != g_tab.mb) build_tables(mb);
	reset_read(ri);
	ri->seq = malloc(allocated_length);
	while (current_length < average_length) {
		for (;;) {
			for (nuc = 0; nuc < 5; nuc++)
				if (r < g_tab.bg[nuc]) { ri->seq[current_length++] = nuc; break; }
			if (current_length == allocated_length) { allocated_length *= 2; ri->seq = realloc(ri->seq, allocated_length); }
			r = draw();
			if (r > stop) break;
		}
		if (current_length < average_length) current_length = 0;
	}
	finish_read(ri, current_length, 'N', 0);
	return kslOK;
}

int emit_read_sequence(struct model_bag* mb, struct read_info* ri, int average_length, unsigned int* seed)
{
	int state, column, hmm, segment, nuc, k;
	int current_length = 0, allocated_length = 100;
	double r = draw();                                           /* :2720, the value is replaced before its first use */
	(void)seed;
	if (mb != g_tab.mb) build_tables(mb);
	reset_read(ri);
	ri->seq = malloc(allocated_length);
	ri->labels = malloc(allocated_length);
	while (current_length < average_length) {
		state = 0; column = 0; hmm = 0; segment = 0;             /* 0 silent, 1 M, 2 I, 3 D */
		for (;;) {
			const struct seg_tab* t = &g_tab.seg[segment];
			int o = hmm * t->nc + column;
			/* transition */
			r = draw();
			switch (state) {
				case 0: {
					const int n = t->nh * t->nc * 2;
					for (k = 0; k < n; k++)
						if (r < t->sil[k]) break;
					if (k < n) { state = 1 + (k & 1); hmm = (k >> 1) / t->nc; column = (k >> 1) % t->nc; }
					break;
				}
				case 1:
					if (r < t->tM[o * 3]) { column++; }
					else if (r < t->tM[o * 3 + 1]) { state = 2; }
					else if (r < t->tM[o * 3 + 2]) { state = 3; column++; }
					else { state = 0; segment++; column = 0; hmm = 0; }
					break;
				case 2:
					if (r < t->tI[o * 2]) { /* II */ }
					else if (r < t->tI[o * 2 + 1]) { state = 1; column++; }
					else { state = 0; segment++; column = 0; hmm = 0; }
					break;
				case 3:
					if (r < t->tD[o]) { column++; }
					else { state = 1; column++; }
					break;
				default:
					break;
			}
			/* emission */
			r = draw();
			if (state == 1 || state == 2) {
				const struct seg_tab* u = &g_tab.seg[segment];
				const float* e = (state == 1 ? u->eM : u->eI) + (size_t)(hmm * u->nc + column) * 5;
				for (nuc = 0; nuc < 5; nuc++)
					if (r < e[nuc]) {
						ri->seq[current_length] = nuc;
						ri->labels[current_length] = segment;
						current_length++;
						break;
					}
			}
			if (current_length == allocated_length) {
				allocated_length *= 2;
				ri->seq = realloc(ri->seq, allocated_length);
				ri->labels = realloc(ri->labels, allocated_length);
			}
			if (segment == g_tab.S) break;
		}
		if (current_length < average_length) current_length = 0;
	}
	finish_read(ri, current_length, 'P', 1);
	return kslOK;
}
