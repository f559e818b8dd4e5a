// tdg_device.h -- structures shared by the host runtime (tdg_host.cpp) and the kernels
// (tdg_kernels.cu).  Internal; the public surface is include/tagdust_b200.h.
#pragma once
#include <cstddef>
#include <cstdint>

namespace tdg {

constexpr int kMaxSegments = 16;
constexpr int kLogsumSize = 16000;
// Decode kernels: one CTA per SM, kBlock reads in flight per CTA (thread-per-read).
#ifndef TDG_BLOCK
#define TDG_BLOCK 512
#endif
constexpr int kBlock = TDG_BLOCK;  // experiments: 256 with two co-resident CTAs per SM (TDG_LANES=2)
// Label-DP kernel uses smaller CTAs (no 64 KB table in shared memory).
constexpr int kDpBlock = 128;
constexpr int kColRec = 12;    // floats per column record (3 x float4)
constexpr int kEmitRec = 10;   // eM[5], eI[5]
constexpr int kMaxSources = 4; // label-DP predecessor sources per HMM (structured form)
constexpr int kMaxStdCols = 16; // STDU segments with up to this many columns run fully unrolled kernels
constexpr int kMaxSegCols = 256; // longest segment (columns = nucleotides + 1): the column-loop paths keep 2 floats per column per thread

// Column record layout (floats), see barcode_hmm.h:87-96 for the transition indices:
//  0 MM  1 MI  2 MD  3 II | 4 IM  5 DD  6 DM  7 MSKIP | 8 ISKIP  9 sM  10 sI  11 live-mask (int bits)
enum ColField { F_MM = 0, F_MI, F_MD, F_II, F_IM, F_DD, F_DM, F_MSKIP, F_ISKIP, F_SM, F_SI, F_LIVE };
// live-mask bits: bit k set <=> field k is not -inf (term can contribute)
#define TDG_LIVE(k) (1u << (k))

#ifdef __CUDACC__
#define TDG_HD __host__ __device__
#else
#define TDG_HD
#endif

// Liveness of the terms of the "standard" profile column pattern that
// set_hmm_transition_parameters(len>=3, mean=-1, stdev=-1) produces (barcode_hmm.c:1787-1880)
// together with B/F/S entry (silent_to_M at column 0 only, :4897-4921):
//   col 0        : DD DM MSKIP ISKIP sI dead
//   cols 1..n-3  : MSKIP ISKIP sM sI dead
//   col n-2      : MD DD MSKIP ISKIP sM sI dead
//   col n-1      : everything dead but MSKIP
// The host only selects the STD code path when every term listed dead IS -inf in the model.
TDG_HD constexpr bool std_live(int nc, int g, int field)
{
	if (g == nc - 1) return field == F_MSKIP;
	if (field == F_MSKIP || field == F_ISKIP || field == F_SI) return false;
	if (field == F_SM) return g == 0;
	if (g == 0) return !(field == F_DD || field == F_DM);
	if (g == nc - 2) return !(field == F_MD || field == F_DD);
	return true;
}

struct SegInfo {
	int32_t nh;        // HMMs in segment
	int32_t nc;        // columns per HMM
	int32_t colbase;   // first global column
	int32_t hmmbase;   // first global HMM
	float   skip;      // model->skip
	int32_t skip_live; // skip != -inf
	int32_t kind;      // code path: 0 generic (runtime live masks), 1 STDU (standard pattern, uniform scalars)
	float   ta, tb, tb2, tc, td;  // STDU transition scalars (see trv<> in tdg_kernels.cu)
	int32_t use_win;   // 1: k_backward records the largest Mb/Ib of the segment per position (KArgs.mbmax) and k_forward
	                   //    skips the Mb/Ib loads and the posterior chain of cells that provably exp() to 0 (unrolled STDU only)
};

// Everything a kernel needs that is not in the model blob (passed by value).
struct KArgs {
	// model
	int32_t S, H, C;
	SegInfo seg[kMaxSegments];
	const float* model_blob;   // device: [colrec C*12][emit C*10]
	int32_t model_floats;      // floats of the model blob
	int32_t model_in_smem;     // 1: the blob is staged into shared memory behind the logsum table; 0: read from global memory
	int32_t dyn_cols;          // columns of shared-memory profile state reserved for the column-loop paths (0 = none)
	const float* logsum_tab;   // device: 16000 floats, entries >= 15700 zeroed
	float r_step;              // log(1 - 1/avg) as float   (barcode_hmm.c:4520)
	float r_end;               // log(1/avg) as float       (:4523)
	float bg[5];               // model[0]->background_nuc_frequency
	// batch (wave) geometry
	int32_t n_reads;           // reads in this wave
	int32_t lmax;              // scratch positions per read
	int32_t words;             // packed 32-bit words per read (8 codes each)
	int32_t win_start;         // matchstart or 0
	int32_t win_len;           // matchend - matchstart, or -1 = whole read
	const uint32_t* seq;       // device: tile layout [tile][word][lane]
	const int32_t* len;        // device: [read]
	// scratch (slot-major, see DESIGN.md "HBM layout")
	float2* bw;                // [cta][C*lmax][kBlock]  (Mb, Ib)
	float*  sb;                // [cta][S*(lmax+2)][kBlock] silent_backward
	float*  sf;                // [cta][S*(lmax+2)][kBlock] silent_forward
	float*  mbmax;             // [cta][S*(lmax+2)][kBlock] max over the segment's HMMs and columns of Mb / Ib at a position
	float*  post;              // [cta][lmax*H][kBlock] posterior matrix rows 1..L
	float*  tp;                // [cta][H][kBlock] total_prob
	uint32_t* prange;          // [cta][H][kBlock] (last<<16)|first position with posterior >= -104
	uint8_t* path;             // [cta][lmax*H][kBlock]
	// per-read outputs (device)
	float* b_score; float* f_score; float* r_score; float* bar_prob; float* mapq;
	int32_t* read_type; int32_t* barcode; int32_t* fingerprint; uint8_t* extracted;
	uint8_t* labels; int32_t label_stride;
	int32_t store_labels;      // 1: the label row is an output (flushed to `labels`); 0: labels are working storage only
	uint16_t* spans;           // [read][span_stride] (start, len) pairs of the R-labelled runs of extracted reads, or NULL
	int32_t span_stride;       // pairs per read (R segments + 1)
	int32_t dust;              // param->dust, 0 = off
	// label-DP tables (device)
	const int32_t* dp_src;     // [H][kMaxSources] source code: >=0 single hmm index, <0 = -(segment+1), INT_MIN = none
	const int32_t* hmm_label;  // [H] mb->label
	const uint8_t* seg_type;   // [S]
	const uint8_t* tmat;       // [H*H] 0/1 (generic label-DP fallback)
	int32_t dp_structured;     // 1: dp_src describes T exactly
	int32_t post_store_all;    // always 1 (see k_forward's posterior store)
	// extraction
	float confidence_threshold; int32_t minlen; int32_t required_finger_len; int32_t do_extract;
	int32_t want_labels;
	int32_t label_smem;        // set by launch_label: the read's labels are staged in shared memory
	// -ref artifact filter (k_artifact; match_to_reference barcode_hmm.c:2478-2583)
	const uint8_t* ref_codes;  // device: nuc codes of all reference sequences back to back (fasta->string)
	const int32_t* ref_index;  // device: [ref_numseq + 1] (fasta->s_index)
	int32_t ref_numseq;
	int32_t filter_error;      // param->filter_error
	int32_t slice_n;           // reads of the whole run_pHMM-equivalent call (the batch)
	int32_t slice_threads;     // param->num_threads: the reference matches groups of four reads per thread slice
	int32_t slice_interval;    // (int)((double)slice_n / (double)slice_threads)
	int32_t slice_base;        // index of this wave's first read inside the batch
	int32_t model_less;        // run_rna_dust path: no HMM ran, every read starts as EXTRACT_SUCCESS and is kept whole
	int32_t dust_after;        // param->dust applied by k_artifact (after the artifact match, barcode_hmm.c:2345-2354)
};

struct LaunchCfg { int ctas; };

// launchers (tdg_kernels.cu); all asynchronous on `stream`
int launch_backward(const KArgs& a, bool store, int ctas, void* stream);
int launch_forward(const KArgs& a, int ctas, void* stream);
int launch_label(const KArgs& a, int ctas_decode, void* stream);
int launch_artifact(const KArgs& a, void* stream);
int kernels_configure(int smem_bytes); // sets the dynamic shared memory attributes once per device
size_t decode_smem_bytes(int model_floats, int dyn_cols);

}  // namespace tdg
