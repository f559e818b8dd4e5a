/* TEST INFRASTRUCTURE -- CPU oracle (plain C restatement of the TagDust2 hot path).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  The product (tagdust_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every function
 * here bit-for-bit against the unmodified reference compiled into
 * oracle/_ref/libtagdust_ref.so, and tests/golden/ holds vectors generated from it.
 */
#ifndef ORACLE_HMM_H
#define ORACLE_HMM_H
#include <stdint.h>
#include "../include/tagdust_b200.h"   /* tdg_model_desc / tdg_run_params layouts only */

void  orc_init_logsum(void);
float orc_logsum(float a, float b);
void  orc_logsum_table(float* out);

/* one read: backward + forward/posterior + label DP; returns 0 */
typedef struct orc_read_out {
	float f_score, b_score, r_score;
	float bar_prob;
	float mapq;
	int32_t read_type, barcode, fingerprint;
} orc_read_out;

int orc_decode_read(const tdg_model_desc* d, const uint8_t* seq /* seq[0..len] */, int len,
                    int want_labels, orc_read_out* out, uint8_t* labels /* len+1 */);
float orc_backward_score(const tdg_model_desc* d, const uint8_t* seq, int len);

int orc_run(const tdg_model_desc* d, const tdg_run_params* p, int mode, int n,
            const uint8_t* codes, size_t stride, const int32_t* len, int num_threads,
            float* mapq, float* bar_prob, float* f_score, float* b_score, float* r_score,
            int32_t* read_type, int32_t* barcode, int32_t* fingerprint, uint8_t* labels,
            uint8_t* seq_out, int32_t* len_out);

/* -ref artifact filter (match_to_reference, barcode_hmm.c:2478-2583): installs the reference set (flat nuc codes +
 * s_index[numseq+1]) used by the following orc_run(MODE_GET_LABEL) / orc_run_rna_dust calls; numseq 0 removes it */
void orc_set_reference(const uint8_t* string, const int32_t* s_index, int numseq, int filter_error);
int  orc_bmp_single(const uint8_t* t, const uint8_t* p, int n, int m);       /* misc.c:718-765 */
int  orc_bpm_check_error(const uint8_t* t, const uint8_t* p, int n, int m);  /* misc.c:572-636 */
int  orc_run_rna_dust(int n, const uint8_t* codes, size_t stride, const int32_t* len, int num_threads, int dust, int32_t* read_type);

int orc_arch_compare(const tdg_model_desc* const* archs, int num_arch, int n,
                     const uint8_t* codes, size_t stride, const int32_t* len, int num_threads,
                     float* b_scores, float* arch_posterior);
#endif
