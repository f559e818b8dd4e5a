/* TEST INFRASTRUCTURE -- CPU oracle: plain-C restatement of the TagDust2 v2.33 per-read
 * HMM hot path on the flattened model (tdg_model_desc).  See oracle_hmm.h for who may
 * use it.  Parity: pinned against the unmodified reference (oracle/_ref) by
 * tests/test_oracle_vs_reference.py; compiled with -ffp-contract=off, no fast-math.
 *
 * Every function names the reference lines it restates (paths under src/).
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "oracle_hmm.h"

#define NEG_INF (-HUGE_VALF)

static float orc_table[TDG_LOGSUM_SIZE];
static int orc_table_ready = 0;

/* misc.c:57-63 init_logsum */
void orc_init_logsum(void)
{
	int i;
	if (orc_table_ready) return;
	for (i = 0; i < TDG_LOGSUM_SIZE; i++) {
		orc_table[i] = log(1. + exp((double)-i / 1000.0f));
	}
	orc_table_ready = 1;
}

void orc_logsum_table(float* out)
{
	orc_init_logsum();
	memcpy(out, orc_table, sizeof(orc_table));
}

#ifdef ORC_LS_TRACE
/* scripts/ls_index_hist.py: the table index of every logsum call, in call order (0xFFFF: an operand was -inf,
 * 0xFFFE: difference >= 15.7 -- the two cases in which the reference does not touch the table) */
unsigned short* orc_ls_trace = 0; long orc_ls_trace_n = 0, orc_ls_trace_cap = 0;
static void ls_trace(unsigned short v) { if (orc_ls_trace && orc_ls_trace_n < orc_ls_trace_cap) orc_ls_trace[orc_ls_trace_n++] = v; }
#endif

/* misc.c:72-78 logsum; HMMER3_MAX/MIN are ?: macros (misc.h:60,66) */
float orc_logsum(float a, float b)
{
	const float max = (a > b) ? a : b;
	const float min = (a < b) ? a : b;
#ifdef ORC_LS_TRACE
	ls_trace(min == -HUGE_VAL ? 0xFFFF : ((max - min) >= 15.7f ? 0xFFFE : (unsigned short)(int)((max - min) * 1000.0f)));
#endif
	return (min == -HUGE_VAL || (max - min) >= 15.7f) ? max : max + orc_table[(int)((max - min) * 1000.0f)];
}
#define LS orc_logsum

/* misc.c:85-92 / 98-105 */
static float p2s(float p) { if (p == 0.0) return -HUGE_VAL; return log(p); }
static float s2p(float p) { if (p == -HUGE_VAL) return 0.0; return exp(p); }

typedef struct orc_work {
	int C, H, S, L;            /* capacities */
	float *Mb, *Ib, *Db;       /* [C][L+2] */
	float *Mf, *If, *Df;       /* [C][L+2] */
	float *SB, *SF;            /* [S][L+2] */
	float *prev;               /* [L+2] previous_silent (aliased by next_silent) */
	float *P;                  /* [(L+1)][H] dyn_prog_matrix */
	int   *path;               /* [(L+1)][H] */
	float *TP;                 /* [H] total_prob */
	int   *colbase;            /* [S] first column index of segment */
	int   *hmmbase;            /* [S] first hmm index of segment */
} orc_work;

static orc_work* work_new(const tdg_model_desc* d, int L)
{
	orc_work* w = calloc(1, sizeof(orc_work));
	int s, c = 0, h = 0;
	size_t n = (size_t)d->total_columns * (L + 2);
	w->C = d->total_columns; w->H = d->total_hmms; w->S = d->num_segments; w->L = L;
	w->Mb = malloc(n * 4); w->Ib = malloc(n * 4); w->Db = malloc(n * 4);
	w->Mf = malloc(n * 4); w->If = malloc(n * 4); w->Df = malloc(n * 4);
	w->SB = malloc((size_t)w->S * (L + 2) * 4); w->SF = malloc((size_t)w->S * (L + 2) * 4);
	w->prev = malloc((size_t)(L + 2) * 4);
	w->P = malloc((size_t)(L + 1) * w->H * 4);
	w->path = malloc((size_t)(L + 1) * w->H * 4);
	w->TP = malloc((size_t)(w->H + 1) * 4);
	w->colbase = malloc(w->S * sizeof(int)); w->hmmbase = malloc(w->S * sizeof(int));
	for (s = 0; s < w->S; s++) {
		w->colbase[s] = c; w->hmmbase[s] = h;
		c += d->seg_num_hmms[s] * d->seg_num_cols[s];
		h += d->seg_num_hmms[s];
	}
	return w;
}

static void work_free(orc_work* w)
{
	if (!w) return;
	free(w->Mb); free(w->Ib); free(w->Db); free(w->Mf); free(w->If); free(w->Df);
	free(w->SB); free(w->SF); free(w->prev); free(w->P); free(w->path); free(w->TP);
	free(w->colbase); free(w->hmmbase); free(w);
}

#define T(c, k) (d->transition[(c) * 9 + (k)])
#define EM(c, x) (d->m_emit[(c) * 5 + (x)])
#define EI(c, x) (d->i_emit[(c) * 5 + (x)])
#define AT(arr, c, i) ((arr)[(size_t)(c) * W + (i)])

/* barcode_hmm.c:3439-3640 backward().  a[0..len] is read (a[len] is the terminator). */
static float backward(const tdg_model_desc* d, orc_work* w, const uint8_t* a, int len)
{
	const int W = w->L + 2;
	const int S = d->num_segments;
	const uint8_t* seqa = a - 1;
	int i, j, f, g, c;
	float* psilent; float* csilent;

	/* :3466-3485 */
	for (c = 0; c < d->total_columns; c++) {
		for (i = 0; i <= len + 1; i++) { AT(w->Mb, c, i) = NEG_INF; AT(w->Ib, c, i) = NEG_INF; AT(w->Db, c, i) = NEG_INF; }
	}
	for (j = 0; j < S; j++) for (i = 0; i <= len + 1; i++) w->SB[j * W + i] = NEG_INF;
	for (i = 0; i <= len + 1; i++) w->prev[i] = NEG_INF;
	w->prev[len + 1] = p2s(1.0f);
	/* :3487-3491 */
	w->SB[(S - 1) * W + len + 1] = p2s(1.0) + d->seg_skip[S - 1];
	for (j = S - 2; j >= 0; j--) w->SB[j * W + len + 1] = w->SB[(j + 1) * W + len + 1] + d->seg_skip[j];

	for (j = S - 1; j >= 0; j--) {
		const int nh = d->seg_num_hmms[j], nc = d->seg_num_cols[j];
		const int m = nc - 1;
		const float skip = d->seg_skip[j];
		psilent = (j == S - 1) ? w->prev : &w->SB[(j + 1) * W];
		csilent = &w->SB[j * W];
		for (f = 0; f < nh; f++) {
			const int c0 = w->colbase[j] + f * nc;     /* column index of g = 0 */
			for (i = len; i > 0; i--) {
				const int x1 = seqa[i + 1];  /* :3516 */
				const int x0 = seqa[i];
				int cc = c0 + m;
				/* last column :3518-3541 */
				AT(w->Mb, cc, i) = psilent[i + 1] + T(cc, TDG_MSKIP);
				AT(w->Ib, cc, i) = psilent[i + 1] + T(cc, TDG_ISKIP);
				AT(w->Ib, cc, i) = LS(AT(w->Ib, cc, i), AT(w->Mb, cc, i + 1) + T(cc, TDG_IM) + EM(cc, x1));
				AT(w->Ib, cc, i) = LS(AT(w->Ib, cc, i), AT(w->Ib, cc, i + 1) + T(cc, TDG_II) + EI(cc, x1));
				csilent[i] = LS(csilent[i], AT(w->Mb, cc, i) + d->silent_to_M[cc] + EM(cc, x0));
				csilent[i] = LS(csilent[i], AT(w->Ib, cc, i) + d->silent_to_I[cc] + EI(cc, x0));
				AT(w->Db, cc, i) = NEG_INF;
				for (g = m - 1; g >= 0; g--) {   /* :3545-3589 */
					const int cg = c0 + g, pg = c0 + g + 1;
					float v;
					v = AT(w->Mb, pg, i + 1) + EM(pg, x1) + T(cg, TDG_MM);
					v = LS(v, psilent[i + 1] + T(cg, TDG_MSKIP));
					v = LS(v, AT(w->Ib, cg, i + 1) + EI(cg, x1) + T(cg, TDG_MI));
					v = LS(v, AT(w->Db, pg, i) + T(cg, TDG_MD));
					AT(w->Mb, cg, i) = v;
					v = AT(w->Ib, cg, i + 1) + T(cg, TDG_II) + EI(cg, x1);
					v = LS(v, psilent[i + 1] + T(cg, TDG_ISKIP));
					v = LS(v, AT(w->Mb, pg, i + 1) + T(cg, TDG_IM) + EM(pg, x1));
					AT(w->Ib, cg, i) = v;
					v = AT(w->Db, pg, i) + T(cg, TDG_DD);
					v = LS(v, AT(w->Mb, pg, i) + EM(pg, x0) + T(cg, TDG_DM));
					AT(w->Db, cg, i) = v;
					csilent[i] = LS(csilent[i], AT(w->Mb, cg, i) + d->silent_to_M[cg] + EM(cg, x0));
					csilent[i] = LS(csilent[i], AT(w->Ib, cg, i) + d->silent_to_I[cg] + EI(cg, x0));
				}
				csilent[i] = LS(csilent[i], psilent[i] + skip);   /* :3604 */
			}
		}
	}
	return w->SB[1];  /* :3610 model[0]->silent_backward[1] */
}

/* barcode_hmm.c:4128-4525 forward_max_posterior_decoding() */
static void forward_decode(const tdg_model_desc* d, orc_work* w, const uint8_t* a, int len,
                           float b_score, int want_labels, orc_read_out* out, uint8_t* labels)
{
	const int W = w->L + 2;
	const int S = d->num_segments, H = d->total_hmms;
	const uint8_t* seqa = a - 1;
	int i, j, f, g, c, h = 0;
	float* psilent; float* csilent;
	float* prev = w->prev;          /* previous_silent; next_silent aliases it (:4151-4152) */
	float* P = w->P; float* TP = w->TP;

	for (c = 0; c < d->total_columns; c++)
		for (i = 0; i <= len; i++) { AT(w->Mf, c, i) = NEG_INF; AT(w->If, c, i) = NEG_INF; AT(w->Df, c, i) = NEG_INF; }
	for (j = 0; j < S; j++) for (i = 0; i <= len + 1; i++) w->SF[j * W + i] = NEG_INF;
	w->SF[0] = p2s(1.0) + d->seg_skip[0];                         /* :4171 */
	for (j = 1; j < S; j++) w->SF[j * W] = w->SF[(j - 1) * W] + d->seg_skip[j];
	for (i = 0; i <= len; i++) for (j = 0; j < H; j++) { P[i * H + j] = NEG_INF; w->path[i * H + j] = -1; }
	for (j = 0; j < H; j++) TP[j] = NEG_INF;
	for (i = 0; i <= len; i++) prev[i] = NEG_INF;
	prev[0] = p2s(1.0);
	prev[len + 1] = p2s(1.0f);

	for (j = 0; j < S; j++) {
		const int nh = d->seg_num_hmms[j], nc = d->seg_num_cols[j];
		const float skip = d->seg_skip[j];
		psilent = (j == 0) ? prev : &w->SF[(j - 1) * W];
		csilent = &w->SF[j * W];
		for (f = 0; f < nh; f++) {
			const int c0 = w->colbase[j] + f * nc;
			for (i = 1; i <= len; i++) {
				const int x = seqa[i];
				float v;
				int cc = c0;
				/* column 0 :4218-4266 */
				AT(w->Mf, cc, i) = psilent[i - 1] + d->silent_to_M[cc] + EM(cc, x);
				TP[h] = LS(TP[h], AT(w->Mf, cc, i) + AT(w->Mb, cc, i) - b_score);
				P[i * H + h] = LS(P[i * H + h], AT(w->Mf, cc, i) + AT(w->Mb, cc, i) - b_score);
				v = psilent[i - 1] + d->silent_to_I[cc];
				v = LS(v, AT(w->If, cc, i - 1) + T(cc, TDG_II));
				v = LS(v, AT(w->Mf, cc, i - 1) + T(cc, TDG_MI));
				v = v + EI(cc, x);
				AT(w->If, cc, i) = v;
				TP[h] = LS(TP[h], psilent[i - 1] + d->silent_to_I[cc] + EI(cc, x) + AT(w->Ib, cc, i) - b_score);
				P[i * H + h] = LS(P[i * H + h], AT(w->If, cc, i) + AT(w->Ib, cc, i) - b_score);
				AT(w->Df, cc, i) = NEG_INF;
				csilent[i] = LS(csilent[i], AT(w->Mf, cc, i) + T(cc, TDG_MSKIP));
				csilent[i] = LS(csilent[i], AT(w->If, cc, i) + T(cc, TDG_ISKIP));
				for (g = 1; g < nc; g++) {   /* :4270-4331 */
					const int cg = c0 + g, pg = c0 + g - 1;
					v = psilent[i - 1] + d->silent_to_M[cg];
					v = LS(v, AT(w->Mf, pg, i - 1) + T(pg, TDG_MM));
					v = LS(v, AT(w->If, pg, i - 1) + T(pg, TDG_IM));
					v = LS(v, AT(w->Df, pg, i) + T(pg, TDG_DM));
					v = v + EM(cg, x);
					AT(w->Mf, cg, i) = v;
					P[i * H + h] = LS(P[i * H + h], AT(w->Mf, cg, i) + AT(w->Mb, cg, i) - b_score);
					v = psilent[i - 1] + d->silent_to_I[cg];
					v = LS(v, AT(w->If, cg, i - 1) + T(cg, TDG_II));
					v = LS(v, AT(w->Mf, cg, i - 1) + T(cg, TDG_MI));
					v = v + EI(cg, x);
					AT(w->If, cg, i) = v;
					P[i * H + h] = LS(P[i * H + h], AT(w->If, cg, i) + AT(w->Ib, cg, i) - b_score);
					v = AT(w->Mf, pg, i) + T(pg, TDG_MD);
					v = LS(v, AT(w->Df, pg, i) + T(pg, TDG_DD));
					AT(w->Df, cg, i) = v;
					csilent[i] = LS(csilent[i], AT(w->Mf, cg, i) + T(cg, TDG_MSKIP));
					csilent[i] = LS(csilent[i], AT(w->If, cg, i) + T(cg, TDG_ISKIP));
				}
				csilent[i] = LS(csilent[i], psilent[i] + skip);   /* :4341 */
			}
			h++;
		}
	}
	out->f_score = w->SF[(S - 1) * W + len];   /* :4349 */

	/* :4354-4382 normalise total_prob per multi-HMM segment (next_silent aliases prev) */
	h = 0;
	prev[0] = NEG_INF; prev[1] = NEG_INF;
	for (j = 0; j < S; j++) {
		const int nh = d->seg_num_hmms[j];
		if (nh > 1) {
			g = h;
			prev[1] = NEG_INF;
			for (f = 0; f < nh; f++) { prev[1] = LS(prev[1], TP[h]); h++; }
			for (f = 0; f < nh; f++) { TP[g] = TP[g] - prev[1]; g++; }
		} else {
			h += nh;
		}
	}
	/* :4385-4429 bar_prob */
	h = 0; g = 1;
	prev[0] = NEG_INF; prev[1] = NEG_INF; prev[2] = p2s(1.0);
	for (j = 0; j < S; j++) {
		const int nh = d->seg_num_hmms[j];
		if (nh > 1) {
			g = 0;
			prev[1] = NEG_INF;
			for (f = 0; f < nh; f++) {
				if (TP[h] > prev[0] && f != nh - 1) prev[0] = TP[h];
				prev[1] = LS(prev[1], TP[h]);
				h++;
			}
			prev[0] = prev[0] - prev[1];
			prev[2] = prev[2] + prev[0];
		} else {
			h += nh;
		}
	}
	if (g) out->bar_prob = p2s(1.0);
	else out->bar_prob = (prev[2] > 0) ? p2s(1.0) : prev[2];

	/* :4516-4523 random model (moved before the optional label DP; independent of it) */
	{
		float r = p2s(1.0);
		for (i = 1; i <= len; i++) {
			c = seqa[i];
			r = r + d->background[c] + p2s(1.0 - (1.0 / (float)d->average_raw_length));
		}
		r += p2s(1.0 / (float)d->average_raw_length);
		out->r_score = r;
	}
	if (!want_labels) return;

	/* :4431-4440 */
	for (i = 0; i <= len; i++) for (j = 0; j < H; j++) P[i * H + j] = s2p(P[i * H + j]);
	/* :4447-4472 */
	{
		float max = 0, tmp; int move = -1;
		for (i = 1; i <= len; i++) {
			for (j = 0; j < H; j++) {
				max = -1;
				for (c = 0; c <= j; c++) {
					tmp = P[(i - 1) * H + c] * d->transition_matrix[c * H + j];
					if (tmp > max) { move = c; max = tmp; }
					if (tmp == max && c == j) { move = c; max = tmp; }
				}
				P[i * H + j] += max;
				w->path[i * H + j] = move;
			}
		}
		/* :4494-4514 */
		i = len; max = -1;
		for (j = 0; j < H; j++) if (P[i * H + j] > max) { max = P[i * H + j]; move = j; }
		for (i = 0; i <= len; i++) labels[i] = 0;
		labels[len] = (uint8_t)move;
		for (i = len; i > 0; i--) { move = w->path[i * H + move]; labels[i - 1] = (uint8_t)move; }
	}
}

/* do_label_thread :2316-2338 / do_probability_estimation :2216-2234 */
static float qscore(float bar_prob_f, float f_score, float r_score)
{
	double bar_prob = bar_prob_f;       /* ri->bar_prob is a double field */
	float pbest = -HUGE_VAL, Q;
	pbest = LS(pbest, f_score);
	pbest = LS(pbest, r_score);
	pbest = 1.0 - s2p((bar_prob + f_score) - pbest);
	if (!pbest) Q = 40.0;
	else if (pbest == 1.0) Q = 0.0;
	else Q = -10.0 * log10(pbest);
	return Q;
}

static int seg_of(const tdg_model_desc* d, int lab) { return d->label[lab] & 0xFFFF; }

/* extract_reads :3172-3313 + make_extracted_read :3325-3356.
 * seq = full read (ri->seq), rewritten in place on success; *plen = ri->len. */
static void extract(const tdg_model_desc* d, const tdg_run_params* p, const uint8_t* labels,
                    uint8_t* seq, int* plen, float mapq, orc_read_out* out)
{
	int j, c1, c2, c3, key = 0, bar = -1, mem = -1, fingerlen = 0, required_finger_len = 0;
	int s_pos = 0, offset = 0, len = *plen, hmm_has_barcode = 0, too_short = 0, in_read = 0, ok = 0;
	if (p->matchstart != -1 || p->matchend != -1) { offset = p->matchstart; len = p->matchend - p->matchstart; }
	for (j = 0; j < d->num_segments; j++) if (d->seg_type[j] == 'F') required_finger_len += d->seg_num_cols[j];
	if (p->confidence_threshold <= mapq) {
		for (j = 0; j < len; j++) {
			c1 = d->label[labels[j + 1]];
			c2 = c1 & 0xFFFF;
			c3 = (c1 >> 16) & 0x7FFF;
			if (d->seg_type[c2] == 'F') { fingerlen++; key = (key << 2) | (seq[j + offset] & 0x3); }
			if (d->seg_type[c2] == 'B') {
				hmm_has_barcode = 1; bar = c3;
				if (bar == d->seg_num_hmms[c2] - 1) hmm_has_barcode = -1;
				mem = c2;
			}
			if (d->seg_type[c2] == 'R') { s_pos++; if (!in_read) in_read = 1; }
			else {
				if (in_read) { if (s_pos < p->minlen) { too_short = 1; break; } }
				in_read = 0; s_pos = 0;
			}
		}
		if (in_read && s_pos < p->minlen) too_short = 1;
		if (!too_short) {
			if (hmm_has_barcode == -1) out->read_type = TDG_EXTRACT_FAIL_BAR_FINGER_NOT_FOUND;
			else if (hmm_has_barcode && required_finger_len) {
				if (fingerlen == required_finger_len && bar != -1) {
					ok = 1; out->barcode = (mem << 16) | bar;
					out->fingerprint = (key << 8) | (required_finger_len <= 255 ? required_finger_len : 255);
				} else out->read_type = TDG_EXTRACT_FAIL_BAR_FINGER_NOT_FOUND;
			} else if (hmm_has_barcode) {
				if (bar != -1) { ok = 1; out->barcode = (mem << 16) | bar; }
				else out->read_type = TDG_EXTRACT_FAIL_BAR_FINGER_NOT_FOUND;
			} else if (required_finger_len) {
				if (fingerlen == required_finger_len) {
					ok = 1;
					out->fingerprint = (key << 8) | (required_finger_len <= 255 ? required_finger_len : 255);
				} else out->read_type = TDG_EXTRACT_FAIL_BAR_FINGER_NOT_FOUND;
			} else ok = 1;
		} else out->read_type = TDG_EXTRACT_FAIL_READ_TOO_SHORT;
	} else out->read_type = TDG_EXTRACT_FAIL_ARCHITECTURE_MISMATCH;
	if (ok) {
		/* make_extracted_read :3325-3356 (uses ri->len, no window offset) */
		s_pos = 0;
		for (j = 0; j < *plen; j++) {
			c2 = seg_of(d, labels[j + 1]);
			if (d->seg_type[c2] == 'R') seq[s_pos] = seq[j]; else seq[s_pos] = 65;
			s_pos++;
		}
		*plen = s_pos;
		out->read_type = TDG_EXTRACT_SUCCESS;
	}
}

/* dust_sequences :2407-2467; seq must be readable up to index max(len, c+1) */
static int dust_low_complexity(const uint8_t* seq, int rlen, int dust_cut)
{
	double triplet[64]; double s = 0.0; int j, c = 0, key, len;
	for (j = 0; j < 64; j++) triplet[j] = 0.0;
	while (seq[c] == 65) c++;
	key = ((seq[c] & 0x3) << 2) | (seq[c + 1] & 0x3);
	len = rlen; if (len > 64) len = 64;
	c += 2;
	for (j = c; j < len; j++) {
		if (seq[j] == 65) break;
		key = key << 2 | (seq[j] & 0x3);
		triplet[key & 0x3F]++;
		c++;
	}
	for (j = 0; j < 64; j++) s += triplet[j] * (triplet[j] - 1.0) / 2.0;
	s = s / (double)(c - 3) * 10.0;
	return s > dust_cut;
}

/* ---- -ref artifact filter --------------------------------------------------------------------------
 * match_to_reference (barcode_hmm.c:2478-2583) with the two Myers bit-vector variants it calls:
 * groups of four reads of a thread slice -> validate_bpm_sse -> bmp_single (misc.c:718-765), best match over all
 * reference sequences, forward strand first; the last (slice length mod 4) reads of a slice -> bpm_check_error
 * (misc.c:572-636), first reference sequence within the error cut-off.  The pattern is the read as
 * make_extracted_read left it (spacer 65 outside R segments). */
static const uint8_t* g_ref_string = 0; static const int32_t* g_ref_index = 0; static int g_ref_numseq = 0, g_ref_cut = 2;
void orc_set_reference(const uint8_t* string, const int32_t* s_index, int numseq, int filter_error)
{
	g_ref_string = string; g_ref_index = s_index; g_ref_numseq = numseq; g_ref_cut = filter_error;
}

/* misc.c:718-765: semi-global edit distance of the first min(m, 63) pattern characters against the target */
int orc_bmp_single(const uint8_t* t, const uint8_t* p, int n, int m)
{
	uint64_t B[4] = {0, 0, 0, 0}, VP, VN = 0, D0, HN, HP, X, MASK;
	int64_t diff; int k, i;
	if (m > 63) m = 63;
	diff = m; k = m;
	for (i = 0; i < m; i++) if (p[i] != 65) B[p[i] & 3u] |= (uint64_t)1 << i;
	VP = ((uint64_t)1 << m) - 1;
	MASK = (uint64_t)1 << (m - 1);
	for (i = 0; i < n; i++) {
		X = B[t[i] & 3u] | VN;
		D0 = ((VP + (X & VP)) ^ VP) | X;
		HN = VP & D0;
		HP = VN | ~(VP | D0);
		X = HP << 1;
		VN = X & D0;
		VP = (HN << 1) | ~(X | D0);
		if (HP & MASK) diff++;
		if (HN & MASK) diff--;
		if (diff < k) k = (int)diff;
	}
	return k;
}

/* misc.c:572-636.  The reference shifts 1 by the character index and by (new_len - 1) without range checks; on x86-64
 * the shift count is taken modulo 64, which is what the compiled reference does and what is restated here. */
int orc_bpm_check_error(const uint8_t* t, const uint8_t* p, int n, int m)
{
	uint64_t B[4] = {0, 0, 0, 0}, VP = ~(uint64_t)0, VN = 0, D0, HN, HP, X, MASK, diff = (uint64_t)m, k;
	int i, new_len = 0, sh;
	for (i = 0; i < m; i++) if (p[i] != 65) { B[p[i] & 3] |= (uint64_t)1 << (i & 63); new_len++; }
	if (new_len > 31) new_len = 31;
	k = (uint64_t)new_len;
	sh = (new_len - 1) & 63;
	MASK = (uint64_t)1 << sh;
	for (i = 0; i < n; i++) {
		X = B[t[i] & 3] | VN;
		D0 = ((VP + (X & VP)) ^ VP) | X;
		HN = VP & D0;
		HP = VN | ~(VP | D0);
		X = HP << 1;
		VN = X & D0;
		VP = (HN << 1) | ~(X | D0);
		diff += (HP & MASK) >> sh;
		diff -= (HN & MASK) >> sh;
		if (diff < k) k = diff;
	}
	return (int)k;
}

/* reverse_complement (misc.c:829-848): spacers stay, A<->T, C<->G, N stays (nuc_code.c:68-72) */
static void revcomp(const uint8_t* p, int len, uint8_t* out)
{
	static const uint8_t rc[5] = {3, 2, 1, 0, 4};
	int i, c = 0;
	for (i = len - 1; i >= 0; i--) out[c++] = (p[i] == 65) ? 65 : rc[p[i] < 5 ? p[i] : 4];
	out[c] = 0;
}

/* one slice [start, end) of a run_pHMM / run_rna_dust call; seqs[i] = rewritten read i, 0-terminated */
static void match_slice(uint8_t** seqs, const int32_t* lens, int32_t* read_type, int start, int end)
{
	int i, j, c;
	int maxlen = 1;
	uint8_t* rev;
	for (i = start; i < end; i++) if (lens[i] > maxlen) maxlen = lens[i];
	rev = malloc(maxlen + 2);
	for (i = start; i <= end - 4; i += 4) {
		for (c = 0; c < 4; c++) {
			const int r = i + c;
			int best = 100000, id = 0;
			revcomp(seqs[r], lens[r], rev);
			for (j = 0; j < g_ref_numseq; j++) {
				const uint8_t* t = g_ref_string + g_ref_index[j];
				const int n = g_ref_index[j + 1] - g_ref_index[j];
				int e = lens[r] > 0 ? orc_bmp_single(t, seqs[r], n, lens[r]) : n;
				if (e < best) { best = e; id = j + 1; }
				e = lens[r] > 0 ? orc_bmp_single(t, rev, n, lens[r]) : n;
				if (e < best) { best = e; id = j + 1; }
			}
			if (best <= g_ref_cut && read_type[r] == TDG_EXTRACT_SUCCESS) read_type[r] = (id << 8) | TDG_EXTRACT_FAIL_MATCHES_ARTIFACTS;
		}
	}
	for (; i < end; i++) {
		int hit = 0;
		revcomp(seqs[i], lens[i], rev);
		for (j = 0; j < g_ref_numseq && !hit; j++) {
			const uint8_t* t = g_ref_string + g_ref_index[j];
			const int n = g_ref_index[j + 1] - g_ref_index[j];
			if (orc_bpm_check_error(t, seqs[i], n, lens[i]) <= g_ref_cut) hit = j + 1;
			else if (orc_bpm_check_error(t, rev, n, lens[i]) <= g_ref_cut) hit = j + 1;
		}
		if (hit && read_type[i] == TDG_EXTRACT_SUCCESS) read_type[i] = (hit << 8) | TDG_EXTRACT_FAIL_MATCHES_ARTIFACTS;
	}
	free(rev);
}

int orc_decode_read(const tdg_model_desc* d, const uint8_t* seq, int len, int want_labels,
                    orc_read_out* out, uint8_t* labels)
{
	orc_work* w;
	orc_init_logsum();
	w = work_new(d, len);
	out->b_score = backward(d, w, seq, len);
	forward_decode(d, w, seq, len, out->b_score, want_labels, out, labels);
	out->mapq = qscore(out->bar_prob, out->f_score, out->r_score);
	work_free(w);
	return 0;
}

float orc_backward_score(const tdg_model_desc* d, const uint8_t* seq, int len)
{
	float b; orc_work* w;
	orc_init_logsum();
	w = work_new(d, len);
	b = backward(d, w, seq, len);
	work_free(w);
	return b;
}

typedef struct orc_job {
	const tdg_model_desc* d; const tdg_run_params* p; int mode;
	const tdg_model_desc* const* archs; int num_arch; float* arch_sum;
	int start, end, maxlen;
	const uint8_t* codes; size_t stride; const int32_t* len;
	float *mapq, *bar_prob, *f_score, *b_score, *r_score;
	int32_t *read_type, *barcode, *fingerprint; uint8_t* labels; uint8_t* seq_out; int32_t* len_out;
	float* b_scores; int n;
	uint8_t** rw_seq; int32_t* rw_len;   /* -ref: the rewritten reads of the whole call */
} orc_job;

/* do_label_thread :2269 / do_probability_estimation :2174 */
static void* label_worker(void* arg)
{
	orc_job* jb = arg;
	const tdg_model_desc* d = jb->d; const tdg_run_params* p = jb->p;
	orc_work* w = work_new(d, jb->maxlen + 1);
	uint8_t* lab = calloc(jb->maxlen + 4, 1);
	uint8_t* seq = calloc(jb->maxlen + 4, 1);
	int i, j;
	const int windowed = (p->matchstart != -1 || p->matchend != -1);
	for (i = jb->start; i < jb->end; i++) {
		const uint8_t* a = jb->codes + (size_t)i * jb->stride;
		int rlen = jb->len[i];
		int len = rlen;
		orc_read_out o;
		memset(&o, 0, sizeof(o));
		o.barcode = -1; o.fingerprint = -1; o.read_type = 0;
		if (windowed) { a += p->matchstart; len = p->matchend - p->matchstart; }
		memset(lab, 0, rlen + 2);
		o.b_score = backward(d, w, a, len);
		forward_decode(d, w, a, len, o.b_score, (jb->mode == TDG_MODE_GET_LABEL) || p->want_labels, &o, lab);
		o.mapq = qscore(o.bar_prob, o.f_score, o.r_score);
		memcpy(seq, jb->codes + (size_t)i * jb->stride, rlen + 1);
		seq[rlen + 1] = 0;
		if (jb->mode == TDG_MODE_GET_LABEL) {
			extract(d, p, lab, seq, &rlen, o.mapq, &o);
			if (g_ref_numseq) {  /* do_label_thread :2345-2354: extract all, then match_to_reference, then dust */
				jb->rw_seq[i] = malloc(rlen + 2);
				memcpy(jb->rw_seq[i], seq, rlen + 1); jb->rw_seq[i][rlen + 1] = 0;
				jb->rw_len[i] = rlen;
			} else if (p->dust && dust_low_complexity(seq, rlen, p->dust)) o.read_type = TDG_EXTRACT_FAIL_LOW_COMPLEXITY;
		}
		if (jb->mapq) jb->mapq[i] = o.mapq;
		if (jb->bar_prob) jb->bar_prob[i] = o.bar_prob;
		if (jb->f_score) jb->f_score[i] = o.f_score;
		if (jb->b_score) jb->b_score[i] = o.b_score;
		if (jb->r_score) jb->r_score[i] = o.r_score;
		if (jb->read_type) jb->read_type[i] = o.read_type;
		if (jb->barcode) jb->barcode[i] = o.barcode;
		if (jb->fingerprint) jb->fingerprint[i] = o.fingerprint;
		if (jb->labels) for (j = 0; j <= jb->len[i]; j++) jb->labels[(size_t)i * jb->stride + j] = lab[j];
		if (jb->seq_out) for (j = 0; j < jb->len[i]; j++) jb->seq_out[(size_t)i * jb->stride + j] = seq[j];
		if (jb->len_out) jb->len_out[i] = rlen;
	}
	if (jb->mode == TDG_MODE_GET_LABEL && g_ref_numseq && jb->read_type) {
		match_slice(jb->rw_seq, jb->rw_len, jb->read_type, jb->start, jb->end);
		for (i = jb->start; i < jb->end; i++) {
			if (p->dust && dust_low_complexity(jb->rw_seq[i], jb->rw_len[i], p->dust)) jb->read_type[i] = TDG_EXTRACT_FAIL_LOW_COMPLEXITY;
			free(jb->rw_seq[i]);
		}
	}
	free(lab); free(seq); work_free(w);
	return NULL;
}

static int max_len(int n, const int32_t* len) { int i, m = 1; for (i = 0; i < n; i++) if (len[i] > m) m = len[i]; return m; }

/* run_pHMM :1895-2029, modes GET_LABEL / GET_PROB: static slices [t*(n/T), (t+1)*(n/T)), tail to last */
int orc_run(const tdg_model_desc* d, const tdg_run_params* p, int mode, int n,
            const uint8_t* codes, size_t stride, const int32_t* len, int num_threads,
            float* mapq, float* bar_prob, float* f_score, float* b_score, float* r_score,
            int32_t* read_type, int32_t* barcode, int32_t* fingerprint, uint8_t* labels,
            uint8_t* seq_out, int32_t* len_out)
{
	int t, interval, ml = max_len(n, len);
	pthread_t* th; orc_job* jobs;
	uint8_t** rw_seq = calloc(n > 0 ? n : 1, sizeof(uint8_t*)); int32_t* rw_len = calloc(n > 0 ? n : 1, sizeof(int32_t));
	if (num_threads < 1) num_threads = 1;
	orc_init_logsum();
	th = calloc(num_threads, sizeof(pthread_t)); jobs = calloc(num_threads, sizeof(orc_job));
	interval = (int)((double)n / (double)num_threads);
	for (t = 0; t < num_threads; t++) {
		orc_job* jb = &jobs[t];
		jb->rw_seq = rw_seq; jb->rw_len = rw_len;
		jb->d = d; jb->p = p; jb->mode = mode; jb->start = t * interval; jb->end = t * interval + interval;
		jb->maxlen = ml; jb->codes = codes; jb->stride = stride; jb->len = len;
		jb->mapq = mapq; jb->bar_prob = bar_prob; jb->f_score = f_score; jb->b_score = b_score; jb->r_score = r_score;
		jb->read_type = read_type; jb->barcode = barcode; jb->fingerprint = fingerprint; jb->labels = labels;
		jb->seq_out = seq_out; jb->len_out = len_out; jb->n = n;
	}
	jobs[num_threads - 1].end = n;
	for (t = 0; t < num_threads; t++) pthread_create(&th[t], NULL, label_worker, &jobs[t]);
	for (t = 0; t < num_threads; t++) pthread_join(th[t], NULL);
	free(th); free(jobs); free(rw_seq); free(rw_len);
	return 0;
}

/* run_rna_dust :2043 + do_rna_dust :2370: every read SUCCESS, then match_to_reference, then dust, per static slice */
int orc_run_rna_dust(int n, const uint8_t* codes, size_t stride, const int32_t* len, int num_threads, int dust, int32_t* read_type)
{
	int t, i, interval;
	uint8_t** seqs = calloc(n > 0 ? n : 1, sizeof(uint8_t*));
	if (num_threads < 1) num_threads = 1;
	for (i = 0; i < n; i++) { seqs[i] = (uint8_t*)codes + (size_t)i * stride; read_type[i] = TDG_EXTRACT_SUCCESS; }
	interval = (int)((double)n / (double)num_threads);
	for (t = 0; t < num_threads; t++) {
		const int start = t * interval, end = (t == num_threads - 1) ? n : t * interval + interval;
		if (g_ref_numseq) match_slice(seqs, len, read_type, start, end);
		for (i = start; i < end; i++)
			if (dust && dust_low_complexity(seqs[i], len[i], dust)) read_type[i] = TDG_EXTRACT_FAIL_LOW_COMPLEXITY;
	}
	free(seqs);
	return 0;
}

/* do_arch_comparison :2111-2148 */
static void* arch_worker(void* arg)
{
	orc_job* jb = arg;
	int i, a;
	orc_work** w = calloc(jb->num_arch, sizeof(orc_work*));
	for (a = 0; a < jb->num_arch; a++) { w[a] = work_new(jb->archs[a], jb->maxlen + 1); jb->arch_sum[a] = p2s(1.0); }
	for (i = jb->start; i < jb->end; i++) {
		for (a = 0; a < jb->num_arch; a++) {
			float b = backward(jb->archs[a], w[a], jb->codes + (size_t)i * jb->stride, jb->len[i]);
			if (jb->b_scores) jb->b_scores[(size_t)a * jb->n + i] = b;
			jb->arch_sum[a] += b;
		}
	}
	for (a = 0; a < jb->num_arch; a++) work_free(w[a]);
	free(w);
	return NULL;
}

/* run_pHMM MODE_ARCH_COMP :1924-1938, :1994-2017 */
int orc_arch_compare(const tdg_model_desc* const* archs, int num_arch, int n,
                     const uint8_t* codes, size_t stride, const int32_t* len, int num_threads,
                     float* b_scores, float* arch_posterior)
{
	int t, a, interval, ml = max_len(n, len);
	pthread_t* th; orc_job* jobs; float sum;
	if (num_threads < 1) num_threads = 1;
	orc_init_logsum();
	th = calloc(num_threads, sizeof(pthread_t)); jobs = calloc(num_threads, sizeof(orc_job));
	interval = (int)((double)n / (double)num_threads);
	for (t = 0; t < num_threads; t++) {
		orc_job* jb = &jobs[t];
		jb->archs = archs; jb->num_arch = num_arch; jb->arch_sum = calloc(num_arch, sizeof(float));
		jb->start = t * interval; jb->end = t * interval + interval; jb->maxlen = ml;
		jb->codes = codes; jb->stride = stride; jb->len = len; jb->b_scores = b_scores; jb->n = n;
	}
	jobs[num_threads - 1].end = n;
	for (t = 0; t < num_threads; t++) pthread_create(&th[t], NULL, arch_worker, &jobs[t]);
	for (t = 0; t < num_threads; t++) pthread_join(th[t], NULL);
	/* caller initialises ab->arch_posterior to log(1) = 0 (test_architectures.c) then adds per thread */
	for (a = 0; a < num_arch; a++) arch_posterior[a] = p2s(1.0);
	for (t = 0; t < num_threads; t++) for (a = 0; a < num_arch; a++) arch_posterior[a] += jobs[t].arch_sum[a];
	sum = arch_posterior[0];
	for (a = 1; a < num_arch; a++) sum = LS(sum, arch_posterior[a]);
	for (a = 0; a < num_arch; a++) arch_posterior[a] = arch_posterior[a] - sum;
	for (t = 0; t < num_threads; t++) free(jobs[t].arch_sum);
	free(th); free(jobs);
	return 0;
}
