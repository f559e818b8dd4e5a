// tdg_kernels.cu -- hand-written sm_100a kernels for the TagDust2 per-read HMM decode path.
//
// Reference behaviour reproduced (files under the reference's src/):
//   backward()                         barcode_hmm.c:3439-3640
//   forward_max_posterior_decoding()   barcode_hmm.c:4128-4525
//   Q score                            barcode_hmm.c:2316-2338 (do_label_thread), :2216-2234
//   extract_reads()                    barcode_hmm.c:3172-3313
//   logsum()                           misc.c:72-78
//
// Design: thread-per-read -- the reference's logsum() is not associative and
// the silent-state chain is strictly ordered, so one thread walks one read in the reference's
// exact operation order and 75 776 reads are in flight per GPU.  Three kernels per wave:
//   k_backward : backward pass; Mb/Ib of every (column, position) go to HBM scratch
//   k_forward  : forward pass + posterior accumulation + total_prob + bar_prob + r_score + Q
//   k_label    : exp() of the posterior matrix, constrained label DP, traceback, extraction
// The 16 000-entry logsum table and the compiled architecture live in shared memory.
// Terms whose transition is log(0) are never evaluated: logsum(x, -inf) == x exactly.
// Per segment the host picks one of three code paths (tdg_host.cu: derive_model):
//   KIND 1, 3..16 columns  standard B/F/S pattern, fully unrolled, columns in registers, five scalar transitions
//   KIND 0, 1..8 columns   any pattern (P/O/G/R, 1-2 column HMMs), unrolled with run-time live masks
//   column loop            longer segments; profile state in shared memory when it fits, else thread-local
//
// Compiled with -fmad=false: the reference binary contains no fused multiply-adds.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <cstdint>
#include <cstdio>
#include "tdg_device.h"

namespace tdg {

#define NEG_INF (-CUDART_INF_F)
constexpr int kDynMaxCols = tdg::kMaxSegCols;  // column-loop paths: thread-local profile state (local memory) for segments up to this long
constexpr int TDG_MAX_HMMS_DEV = 255;

// ------------------------------------------------------------------------------------------
// logsum (misc.c:72-78).  The device table is the host table with entries >= 15700 set to +0:
// (max-min) >= 15.7f  <=>  (int)((max-min)*1000.0f) >= 15700, so `max + tab[idx]` returns max
// exactly where the reference returns max, and the clamp maps +inf / NaN differences
// (min == -inf, or both -inf) to an entry that is 0 as well.  No branch, no select.
// The truncation (int)x is done on the FP32 pipe: x + 2^23 rounded toward zero has floor(x)
// in its mantissa (0 <= x < 2^23), so its bit pattern << 2 plus a pre-offset shared-memory
// base is the byte address of tab[(int)x]; no F2I (XU pipe) on the hot path.
// ------------------------------------------------------------------------------------------
#ifndef TDG_PREFETCH_DIST
#define TDG_PREFETCH_DIST 5
#endif
constexpr int kPrefetchDist = TDG_PREFETCH_DIST;  // 0 = no L1 prefetch of the silent-state lines
#ifndef TDG_AHEAD
#define TDG_AHEAD 1
#endif
constexpr int kAhead = TDG_AHEAD;  // register prefetch distance (positions) of the silent-state loads in k_backward
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Silent-state arrays (silent_backward / silent_forward, ~90 MB per wave at cfg2) are re-read and
// re-written once per HMM of a segment while 25 GB of Mb/Ib stream past them: they are accessed with
// an L2 evict_last policy so that they stay resident in the 126 MB L2 (the Mb/Ib stream uses
// evict-first __stcs/__ldcs).
#ifndef TDG_SILENT_EVICT_LAST
#define TDG_SILENT_EVICT_LAST 1
#endif
__device__ __forceinline__ uint64_t make_keep_policy()
{
	uint64_t pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
__device__ __forceinline__ float ld_keep(const float* p, uint64_t pol)
{
#if TDG_SILENT_EVICT_LAST
	float v;
	asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
	return v;
#else
	return *p;
#endif
}
__device__ __forceinline__ void st_keep(float* p, float v, uint64_t pol)
{
#if TDG_SILENT_EVICT_LAST
	asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
#else
	*p = v;
#endif
}

// Mb/Ib scratch load of k_forward.  -DTDG_EXP_FWD_NOLOAD (experiment builds only, scripts/build_variant.sh) replaces it
// with a constant: the kernel's compute-only time, i.e. what k_forward would cost if the backward values came for free.
__device__ __forceinline__ float2 ld_bw(const float2* p)
{
#ifdef TDG_EXP_FWD_NOLOAD
	return make_float2(-3.0f, -5.0f);
#else
	return __ldcs(p);
#endif
}

typedef uint32_t TabAddr;  // shared-memory byte address of the table minus (0x4B000000 << 2)

__device__ __forceinline__ TabAddr make_tab_addr(const float* tab)
{
	return (uint32_t)__cvta_generic_to_shared(tab) - (0x4B000000u << 2);
}

__device__ __forceinline__ float LS(float a, float b, TabAddr tab)
{
	// max - min == |a - b| bit for bit (IEEE negation is exact), and the |.| rides on the FMUL
	// as an operand modifier: one instruction less than max/min/sub.
	const float mx = fmaxf(a, b);
	const float p = fminf(fabsf(a - b) * 1000.0f, 15999.0f);
	const uint32_t bits = __float_as_uint(__fadd_rz(p, 8388608.0f));
	float t;
	asm("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(tab + (bits << 2)));
	return mx + t;
}

// Per-read profile state M[g], I[g]: registers (fully unrolled paths), or -- for the column-loop paths of
// long segments -- shared memory ([column][thread], conflict-free) when the host found room for it, else
// thread-local arrays (which spill to L1/L2: the 227 KB shared-memory carve-out leaves ~28 KB of L1).
template <int N>
struct RegVec {
	float v[N];
	__device__ __forceinline__ float& operator[](int g) { return v[g]; }
	__device__ __forceinline__ void bind(float*) {}
};
struct ShmVec {
	float* p;
	__device__ __forceinline__ float& operator[](int g) { return p[(size_t)g * kBlock]; }
	__device__ __forceinline__ void bind(float* q) { p = q; }
};
template <int N, bool SMS> struct StateVec { typedef RegVec<N> type; };
template <int N> struct StateVec<N, true> { typedef ShmVec type; };

template <int NC, bool STD>
struct Cols {
	static constexpr int N = NC > 0 ? NC : kDynMaxCols;
};

template <bool STD>
__device__ __forceinline__ bool live_of(int nc, int g, int field, uint32_t mask)
{
	if (STD) return std_live(nc, g, field);
	return (mask >> field) & 1u;
}

struct Smem {
	TabAddr tab;          // pre-offset shared address of the 16000 logsum entries
	const float* colrec;  // C * 12
	const float* emit;    // C * 10
	float* dyn;           // [2 * dyn_cols][kBlock] profile state of the column-loop paths (this thread's lane), or unused
};

// GM = the model tables stay in global memory (architectures too large for shared memory); the usual GM = false
// instantiation keeps the shared-memory pointers the compiler can address with LDS.
template <bool GM>
__device__ __forceinline__ Smem stage_smem(const KArgs& a, float* smem)
{
	for (int k = threadIdx.x; k < kLogsumSize; k += blockDim.x) smem[k] = a.logsum_tab[k];
	float* m = smem + kLogsumSize;
	const int staged = GM ? 0 : a.model_floats;
	for (int k = threadIdx.x; k < staged; k += blockDim.x) m[k] = a.model_blob[k];
	// The pre-offset table base is bounced through shared memory so that it reaches the inner
	// loops as an opaque register: ptxas otherwise re-splits it into (window base, constant) and
	// spends a second integer add per logsum on the constant.
	__shared__ volatile uint32_t s_tab_addr;
	if (threadIdx.x == 0) s_tab_addr = make_tab_addr(smem);
	__syncthreads();
	Smem s;
	s.tab = s_tab_addr;
	const float* tables = GM ? a.model_blob : m;
	s.colrec = tables;
	s.emit = tables + (size_t)a.C * kColRec;
	s.dyn = m + staged + threadIdx.x;
	return s;
}

// packed sequence access: tile layout [tile][word][lane], 8 codes of 4 bits per word.
struct SeqReader {
	const uint32_t* base;  // points at word 0 of this lane
	__device__ __forceinline__ int code(int pos) const
	{
		const uint32_t w = __ldg(base + (size_t)(pos >> 3) * 32);
		return (w >> ((pos & 7) * 4)) & 0xF;
	}
};

__device__ __forceinline__ SeqReader make_reader(const KArgs& a, int read)
{
	SeqReader r;
	r.base = a.seq + ((size_t)(read >> 5) * a.words) * 32 + (read & 31);
	return r;
}

// Register-resident cursors over the packed codes: the current 8-code word and the next one
// are kept in registers, so the per-position code is a shift+mask and the (coalesced, 128 B
// per warp) word load is issued 8 positions before its first use.
struct SeqDown {  // positions pos, pos-1, ...
	const uint32_t* base; uint32_t wcur, wnext; int pos;
	__device__ __forceinline__ void init(const SeqReader& rd, int p0)
	{
		base = rd.base; pos = p0; wcur = 0; wnext = 0;
		if (p0 >= 0) {
			const int wi = p0 >> 3;
			wcur = __ldg(base + (size_t)wi * 32);
			if (wi > 0) wnext = __ldg(base + (size_t)(wi - 1) * 32);
		}
	}
	__device__ __forceinline__ int get()
	{
		const int x = (wcur >> ((pos & 7) * 4)) & 0xF;
		if ((pos & 7) == 0) {
			wcur = wnext;
			const int wi = (pos >> 3) - 2;
			if (wi >= 0) wnext = __ldg(base + (size_t)wi * 32);
		}
		pos--;
		return x;
	}
};
struct SeqUp {  // positions pos, pos+1, ...
	const uint32_t* base; uint32_t wcur, wnext; int pos, words;
	__device__ __forceinline__ void init(const SeqReader& rd, int p0, int nwords)
	{
		base = rd.base; pos = p0; words = nwords;
		const int wi = p0 >> 3;
		wcur = __ldg(base + (size_t)wi * 32);
		wnext = (wi + 1 < words) ? __ldg(base + (size_t)(wi + 1) * 32) : 0u;
	}
	__device__ __forceinline__ int get()
	{
		const int x = (wcur >> ((pos & 7) * 4)) & 0xF;
		if ((pos & 7) == 7) {
			wcur = wnext;
			const int wi = (pos >> 3) + 2;
			if (wi < words) wnext = __ldg(base + (size_t)wi * 32);
		}
		pos++;
		return x;
	}
};

// ------------------------------------------------------------------------------------------
// Column parameters.  KIND 0: runtime live masks, values from the column record in shared memory.
// KIND 1 ("STDU"): the standard column pattern (std_live) AND every HMM of the segment carries the
// same five transition scalars the reference's set_hmm_transition_parameters() writes for it
// (barcode_hmm.c:1787-1880) -- verified bit for bit by the host (tdg_host.cu: derive_model):
//   MM = a (cols 0..n-2)          MI = b (cols 0..n-3), b2 (col n-2)      MD = b (cols 0..n-3)
//   II = c, IM = d (cols 0..n-2)  DD = c, DM = d (cols 1..n-3)            DM(col n-2) = +0, MSKIP(col n-1) = +0
// Then the values are registers / literal zeros and no parameter load remains in the inner loop.
// ------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ float trv(const float* r, const SegInfo& sg, int nc, int g, int field)
{
	if (KIND == 1) {
		switch (field) {
			case F_MM: return sg.ta;
			case F_MI: return g == nc - 2 ? sg.tb2 : sg.tb;
			case F_MD: return sg.tb;
			case F_II: return sg.tc;
			case F_IM: return sg.td;
			case F_DD: return sg.tc;
			case F_DM: return g == nc - 2 ? 0.0f : sg.td;
			default: return 0.0f;  // MSKIP of the last column (the only other live transition)
		}
	}
	return r[field];
}

// this thread's lane of the per-position maximum of Mb / Ib (same geometry as silent_backward)
__device__ __forceinline__ float* a_mbmax_lane(const KArgs& a)
{
	return a.mbmax + (size_t)blockIdx.x * ((size_t)a.S * (size_t)(a.lmax + 2)) * kBlock + threadIdx.x;
}

// ------------------------------------------------------------------------------------------
// backward, one segment (all HMMs f, all positions i).  barcode_hmm.c:3496-3607
// The silent-state values of the next iteration (cs[i-1], ps[i-1]) are loaded one position
// ahead (+ an L1 prefetch kPrefetchDist ahead): they come from L2/HBM and would otherwise
// stall the ordered chain.  Scratch layout of one HMM: [position][column][512 lanes] float2.
// ------------------------------------------------------------------------------------------
template <int NC, int KIND, bool STORE, bool SMS = false>
__device__ __forceinline__ void bwd_segment(const KArgs& a, const Smem& sm, const SegInfo sg, int j,
                                            const SeqReader& rd, int off, int len, int lw, int x_term, bool last_seg,
                                            float2* __restrict__ bw, float* __restrict__ sb, int f_begin = 0)
{
	constexpr bool STD = KIND == 1;
	constexpr int N = Cols<NC, STD>::N;
	constexpr int UN = NC > 0 ? N : 1;
	const int nc = NC > 0 ? NC : sg.nc;
	const int m = nc - 1;
	// STDU: the last column is not stored -- Mb[m][i] = ps[i+1] + 0 and Ib[m][i] = -inf (std_live);
	// k_forward rebuilds it from silent_backward of the next segment.
	const int ncs = STD ? nc - 1 : nc;
	const size_t W = (size_t)(a.lmax + 2);
	float* cs_arr = sb + ((size_t)j * W) * kBlock;
	const float* ps_arr = sb + ((size_t)(j + 1) * W) * kBlock;
	const TabAddr tab = sm.tab;
	const uint64_t keep = make_keep_policy();
	const bool win = STORE && STD && NC > 0 && !SMS && sg.use_win;   // see bwd_group_std
	float* mb_arr = win ? a_mbmax_lane(a) + ((size_t)j * W) * kBlock : nullptr;

	for (int f = f_begin; f < sg.nh; ++f) {
		const int c0 = sg.colbase + f * nc;
		const float* rec = sm.colrec + (size_t)c0 * kColRec;
		const float* em = sm.emit + (size_t)c0 * kEmitRec;
		typename StateVec<N, SMS>::type M, I;
		M.bind(sm.dyn); I.bind(sm.dyn + (size_t)a.dyn_cols * kBlock);
		// emissions of the residue right of the current position (seqa[i+1]): kept in registers on the
		// unrolled paths, looked up again from shared memory on the shared-state path (code xc)
		float eMc[SMS ? 1 : N];
		float eIc[(STD || SMS) ? 1 : N];  // STDU: insert emissions are the same for every column of the segment
		int xc = x_term;
		// state at i = len+1 : all -inf (:3466-3485); emissions of seqa[len+1] = a[len] (:3516)
#pragma unroll UN
		for (int g = 0; g < (NC > 0 ? N : nc); ++g) {
			if (g < nc) {
				M[g] = NEG_INF; I[g] = NEG_INF;
				if (!SMS) {
					eMc[SMS ? 0 : g] = em[g * kEmitRec + x_term];
					if (!STD) eIc[(STD || SMS) ? 0 : g] = em[g * kEmitRec + 5 + x_term];
				}
			}
		}
		if (STD && !SMS) eIc[0] = em[5 + x_term];
		const float sM0 = STD ? rec[F_SM] : 0.0f;
		float ps1 = last_seg ? 0.0f : ps_arr[(size_t)(len + 1) * kBlock];
		SeqDown sd;
		sd.init(rd, off + lw - 1);
		float* csp = cs_arr + (size_t)lw * kBlock;          // -> cs[i]
		const float* psp = ps_arr + (size_t)lw * kBlock;    // -> ps[i]
		float2* bwp = bw + ((size_t)c0 * a.lmax + (size_t)(lw - 1) * ncs) * kBlock;  // -> (position i, column 0)
		// silent-state values are loaded kAhead positions ahead of their use (they come from L2/HBM
		// under a saturating store stream: one iteration of lead is not enough)
		float cs_q[kAhead], ps_q[kAhead];
#pragma unroll
		for (int k = 0; k < kAhead; ++k) {
			cs_q[k] = (lw - k >= 1) ? ld_keep(csp - (size_t)k * kBlock, keep) : NEG_INF;
			ps_q[k] = (!last_seg && lw - k >= 1) ? ld_keep(psp - (size_t)k * kBlock, keep) : NEG_INF;
		}
		for (int i = lw; i >= 1; --i) {
			const int x0 = sd.get();  // seqa[i]
			float cs = cs_q[0];
			const float ps0 = ps_q[0];
#pragma unroll
			for (int k = 0; k + 1 < kAhead; ++k) { cs_q[k] = cs_q[k + 1]; ps_q[k] = ps_q[k + 1]; }
			if (i > kAhead) {
				cs_q[kAhead - 1] = ld_keep(csp - (size_t)kAhead * kBlock, keep);
				if (!last_seg) ps_q[kAhead - 1] = ld_keep(psp - (size_t)kAhead * kBlock, keep);
			}
			if (kPrefetchDist > 0 && i > kPrefetchDist) {  // pull the silent-state lines of iteration i-kPrefetchDist towards L1
				prefetch_l1(csp - (size_t)kPrefetchDist * kBlock);
				if (!last_seg) prefetch_l1(psp - (size_t)kPrefetchDist * kBlock);
			}
			if (i <= len) {
				float* mbp = win ? mb_arr + (size_t)i * kBlock : nullptr;
				float gmax = (win && f > 0) ? ld_keep(mbp, keep) : NEG_INF;
				float eM0[SMS ? 1 : N];
				float eI0[(STD || SMS) ? 1 : N];
				if (!SMS) {
#pragma unroll UN
					for (int g = 0; g < (NC > 0 ? N : nc); ++g) {
						if (g < nc) {
							eM0[SMS ? 0 : g] = em[g * kEmitRec + x0];
							if (!STD) eI0[(STD || SMS) ? 0 : g] = em[g * kEmitRec + 5 + x0];
						}
					}
					if (STD) eI0[0] = em[5 + x0];
				}
#define EMC(g) (SMS ? em[(g) * kEmitRec + xc] : eMc[SMS ? 0 : (g)])
#define EM0(g) (SMS ? em[(g) * kEmitRec + x0] : eM0[SMS ? 0 : (g)])
#define EIC(g) (SMS ? em[(STD ? 0 : (g)) * kEmitRec + 5 + xc] : eIc[(STD || SMS) ? 0 : (g)])
#define EI0(g) (SMS ? em[(STD ? 0 : (g)) * kEmitRec + 5 + x0] : eI0[(STD || SMS) ? 0 : (g)])
				// ---- last column (:3518-3541)
				float oldMp, newMp, D;
				{
					const float* r = rec + m * kColRec;
					const uint32_t lv = STD ? 0u : __float_as_uint(r[F_LIVE]);
					float nM = live_of<STD>(nc, m, F_MSKIP, lv) ? ps1 + trv<KIND>(r, sg, nc, m, F_MSKIP) : NEG_INF;
					float nI = live_of<STD>(nc, m, F_ISKIP, lv) ? ps1 + r[F_ISKIP] : NEG_INF;
					if (live_of<STD>(nc, m, F_IM, lv)) nI = LS(nI, M[m] + r[F_IM] + EMC(m), tab);
					if (live_of<STD>(nc, m, F_II, lv)) nI = LS(nI, I[m] + r[F_II] + EIC(m), tab);
					if (live_of<STD>(nc, m, F_SM, lv)) cs = LS(cs, nM + r[F_SM] + EM0(m), tab);
					if (live_of<STD>(nc, m, F_SI, lv)) cs = LS(cs, nI + r[F_SI] + EI0(m), tab);
					oldMp = M[m]; newMp = nM; D = NEG_INF;
					M[m] = nM; I[m] = nI;
					if (STORE && !STD) __stcs(&bwp[(size_t)m * kBlock], make_float2(nM, nI));
				}
				// ---- columns m-1 .. 0 (:3545-3589)
#pragma unroll UN
				for (int gg = 1; gg < (NC > 0 ? N : nc); ++gg) {
					const int g = m - gg;
					if (g >= 0) {
						const int p = g + 1;
						const float* r = rec + g * kColRec;
						const uint32_t lv = STD ? 0u : __float_as_uint(r[F_LIVE]);
						const float oldMg = M[g];
						float v;
						// M_backward[g][i]
						v = live_of<STD>(nc, g, F_MM, lv) ? oldMp + EMC(p) + trv<KIND>(r, sg, nc, g, F_MM) : NEG_INF;
						if (live_of<STD>(nc, g, F_MSKIP, lv)) v = LS(v, ps1 + r[F_MSKIP], tab);
						if (live_of<STD>(nc, g, F_MI, lv)) v = LS(v, I[g] + EIC(g) + trv<KIND>(r, sg, nc, g, F_MI), tab);
						if (live_of<STD>(nc, g, F_MD, lv)) v = LS(v, D + trv<KIND>(r, sg, nc, g, F_MD), tab);
						const float nM = v;
						// I_backward[g][i]
						v = live_of<STD>(nc, g, F_II, lv) ? I[g] + trv<KIND>(r, sg, nc, g, F_II) + EIC(g) : NEG_INF;
						if (live_of<STD>(nc, g, F_ISKIP, lv)) v = LS(v, ps1 + r[F_ISKIP], tab);
						if (live_of<STD>(nc, g, F_IM, lv)) v = LS(v, oldMp + trv<KIND>(r, sg, nc, g, F_IM) + EMC(p), tab);
						const float nI = v;
						// D_backward[g][i]
						{
							const bool ldd = live_of<STD>(nc, g, F_DD, lv);
							const bool ldm = live_of<STD>(nc, g, F_DM, lv);
							float dv = NEG_INF;
							if (ldd) dv = D + trv<KIND>(r, sg, nc, g, F_DD);
							if (ldm) {
								const float t = newMp + EM0(p) + trv<KIND>(r, sg, nc, g, F_DM);
								dv = ldd ? LS(dv, t, tab) : t;
							}
							D = dv;
						}
						if (live_of<STD>(nc, g, F_SM, lv)) cs = LS(cs, nM + (STD ? sM0 : r[F_SM]) + EM0(g), tab);
						if (live_of<STD>(nc, g, F_SI, lv)) cs = LS(cs, nI + r[F_SI] + EI0(g), tab);
						M[g] = nM; I[g] = nI;
						oldMp = oldMg; newMp = nM;
						if (STORE) __stcs(&bwp[(size_t)g * kBlock], make_float2(nM, nI));
						if (win) gmax = fmaxf(gmax, fmaxf(nM, nI));
					}
				}
				if (sg.skip_live) cs = LS(cs, ps0 + sg.skip, tab);  // once per HMM f (:3604)
				st_keep(csp, cs, keep);
				if (win) st_keep(mbp, gmax, keep);
#pragma unroll UN
				for (int g = 0; g < (NC > 0 ? N : nc); ++g) {
					if (!SMS && g < nc) { eMc[SMS ? 0 : g] = eM0[SMS ? 0 : g]; if (!STD) eIc[(STD || SMS) ? 0 : g] = eI0[(STD || SMS) ? 0 : g]; }
				}
				if (STD && !SMS) eIc[0] = eI0[0];
				xc = x0;
				ps1 = ps0;
#undef EMC
#undef EM0
#undef EIC
#undef EI0
			}
			csp -= kBlock; psp -= kBlock; bwp -= (size_t)ncs * kBlock;
		}
	}
}

// ------------------------------------------------------------------------------------------
// backward, standard-pattern segment, G HMMs per position loop (f .. f+G-1).  Same arithmetic and the same order of the
// silent-state chain as bwd_segment<NC, 1> run on f, then on f+1, ... -- cs[i] takes f's term, f's skip term, f+1's term,
// f+1's skip term, ... -- but cs[i] / ps[i] are loaded and stored once per group (they are 23 % of k_backward's HBM
// traffic at cfg2, and k_backward is HBM-bound), the sequence word, the insert emission and the loop bookkeeping are
// shared, and the G recurrences are independent instruction streams.  The HMMs that do not fill a group go through
// bwd_segment.
// ------------------------------------------------------------------------------------------
#ifndef TDG_NO_PAIRS
#ifndef TDG_BWD_GROUP
#define TDG_BWD_GROUP 4   // HMMs per position loop for segments of up to 6 columns; measured at cfg2 (6 columns, 49 HMMs), per wave:
                          // 1: 5.84 ms, 2: 5.03 ms, 3: 4.92-5.01 ms, 4: 4.69-4.76 ms, 5: 4.75-4.85 ms, 6: 4.93 ms, 7: 5.6 ms (spills)
#endif
template <int NC, bool STORE, int G>
__device__ __forceinline__ void bwd_group_std(const KArgs& a, const Smem& sm, const SegInfo sg, int j,
                                              const SeqReader& rd, int off, int len, int lw, int x_term, bool last_seg,
                                              float2* __restrict__ bw, float* __restrict__ sb)
{
	constexpr int m = NC - 1;
	constexpr int ncs = NC - 1;
	const size_t W = (size_t)(a.lmax + 2);
	float* cs_arr = sb + ((size_t)j * W) * kBlock;
	const float* ps_arr = sb + ((size_t)(j + 1) * W) * kBlock;
	const TabAddr tab = sm.tab;
	const uint64_t keep = make_keep_policy();
	const float ta = sg.ta, tb = sg.tb, tb2 = sg.tb2, tc = sg.tc, td = sg.td;
	const size_t hmm_stride = (size_t)NC * a.lmax * kBlock;   // scratch of one HMM
	const bool win = STORE && sg.use_win;                     // record max(Mb, Ib) of the segment per position for k_forward
	float* mb_arr = win ? a_mbmax_lane(a) + ((size_t)j * W) * kBlock : nullptr;

	int f = 0;
	for (; f + G <= sg.nh; f += G) {
		const int c0 = sg.colbase + f * NC;
		const float* em0 = sm.emit + (size_t)c0 * kEmitRec;
		float M[G][NC], I[G][NC], eMc[G][NC], sM0[G];
#pragma unroll
		for (int q = 0; q < G; ++q) {
			sM0[q] = (sm.colrec + (size_t)(c0 + q * NC) * kColRec)[F_SM];
#pragma unroll
			for (int g = 0; g < NC; ++g) {
				M[q][g] = NEG_INF; I[q][g] = NEG_INF;
				eMc[q][g] = em0[(q * NC + g) * kEmitRec + x_term];
			}
		}
		float eIc = em0[5 + x_term];   // one insert-emission row for the whole segment (checked by the host)
		float ps1 = last_seg ? 0.0f : ps_arr[(size_t)(len + 1) * kBlock];
		SeqDown sd;
		sd.init(rd, off + lw - 1);
		float* csp = cs_arr + (size_t)lw * kBlock;
		const float* psp = ps_arr + (size_t)lw * kBlock;
		float2* bwp = bw + ((size_t)c0 * a.lmax + (size_t)(lw - 1) * ncs) * kBlock;   // HMM f; HMM f + q is q * hmm_stride further
		float cs_n = (lw >= 1) ? ld_keep(csp, keep) : NEG_INF;
		float ps_n = (!last_seg && lw >= 1) ? ld_keep(psp, keep) : NEG_INF;
		for (int i = lw; i >= 1; --i) {
			const int x0 = sd.get();
			float cs = cs_n;
			const float ps0 = ps_n;
			if (i > 1) {
				cs_n = ld_keep(csp - kBlock, keep);
				if (!last_seg) ps_n = ld_keep(psp - kBlock, keep);
			}
			if (kPrefetchDist > 0 && i > kPrefetchDist) {
				prefetch_l1(csp - (size_t)kPrefetchDist * kBlock);
				if (!last_seg) prefetch_l1(psp - (size_t)kPrefetchDist * kBlock);
			}
			if (i <= len) {
				const float eI0 = em0[5 + x0];
				float* mbp = win ? mb_arr + (size_t)i * kBlock : nullptr;
				float gmax = (win && f > 0) ? ld_keep(mbp, keep) : NEG_INF;   // what the HMMs before this group recorded
#pragma unroll
				for (int q = 0; q < G; ++q) {
					// one HMM: columns m .. 0 at this position (the same expressions as bwd_segment<NC, 1>)
					const float* em = em0 + (size_t)q * NC * kEmitRec;
					float2* bq = bwp + (size_t)q * hmm_stride;
					float eM0[NC];
#pragma unroll
					for (int g = 0; g < NC; ++g) eM0[g] = em[g * kEmitRec + x0];
					float oldMp = M[q][m];
					float newMp = ps1 + 0.0f;   // last column: only MSKIP (= +0) is live
					float D = NEG_INF;
					M[q][m] = newMp; I[q][m] = NEG_INF;
#pragma unroll
					for (int gg = 1; gg < NC; ++gg) {
						const int g = m - gg, p = g + 1;
						const float oldMg = M[q][g];
						float v = oldMp + eMc[q][p] + ta;                                          // MM
						v = LS(v, I[q][g] + eIc + (g == NC - 2 ? tb2 : tb), tab);                   // MI
						if (g <= NC - 3) v = LS(v, D + tb, tab);                                    // MD
						const float nM = v;
						v = I[q][g] + tc + eIc;                                                     // II
						v = LS(v, oldMp + td + eMc[q][p], tab);                                     // IM
						const float nI = v;
						{
							const bool ldd = (g >= 1 && g <= NC - 3), ldm = (g >= 1);
							float dv = NEG_INF;
							if (ldd) dv = D + tc;
							if (ldm) {
								const float t = newMp + eM0[p] + (g == NC - 2 ? 0.0f : td);
								dv = ldd ? LS(dv, t, tab) : t;
							}
							D = dv;
						}
						if (g == 0) cs = LS(cs, nM + sM0[q] + eM0[0], tab);                         // SM: column 0 only
						M[q][g] = nM; I[q][g] = nI;
						oldMp = oldMg; newMp = nM;
						if (STORE) __stcs(&bq[(size_t)g * kBlock], make_float2(nM, nI));
						if (win) gmax = fmaxf(gmax, fmaxf(nM, nI));
					}
					if (sg.skip_live) cs = LS(cs, ps0 + sg.skip, tab);
#pragma unroll
					for (int g = 0; g < NC; ++g) eMc[q][g] = eM0[g];
				}
				st_keep(csp, cs, keep);
				if (win) st_keep(mbp, gmax, keep);
				eIc = eI0;
				ps1 = ps0;
			}
			csp -= kBlock; psp -= kBlock; bwp -= (size_t)ncs * kBlock;
		}
	}
	if (f < sg.nh) bwd_segment<NC, 1, STORE>(a, sm, sg, j, rd, off, len, lw, x_term, last_seg, bw, sb, f);
}
#endif

template <bool STORE, bool GM = false>
__global__ void __launch_bounds__(kBlock, 512 / kBlock) k_backward(const KArgs a)
{
	extern __shared__ float smem_f[];
	const Smem sm = stage_smem<GM>(a, smem_f);
	const int slot = blockIdx.x * kBlock + threadIdx.x;
	const int read = slot;
	const bool valid = read < a.n_reads;
	int len = 0, off = 0;
	if (valid) {
		len = a.len[read];
		if (a.win_len >= 0) { off = a.win_start; len = a.win_len; }
	}
	const int lw = __reduce_max_sync(0xffffffffu, len);
	const SeqReader rd = make_reader(a, valid ? read : 0);
	const size_t W = (size_t)(a.lmax + 2);
	float2* bw = a.bw + (size_t)blockIdx.x * ((size_t)a.C * a.lmax) * kBlock + threadIdx.x;
	float* sb = a.sb + (size_t)blockIdx.x * ((size_t)a.S * W) * kBlock + threadIdx.x;
	if (!valid) len = 0;
	const int x_term = rd.code(off + len);

	// init silent_backward (:3479-3491): -inf, then the len+1 chain of skips
	for (int j = 0; j < a.S; ++j)
		for (int i = 0; i <= len + 1; ++i) sb[((size_t)j * W + i) * kBlock] = NEG_INF;
	{
		float v = 0.0f + a.seg[a.S - 1].skip;
		sb[((size_t)(a.S - 1) * W + len + 1) * kBlock] = v;
		for (int j = a.S - 2; j >= 0; --j) {
			v = v + a.seg[j].skip;
			sb[((size_t)j * W + len + 1) * kBlock] = v;
		}
	}
	for (int j = a.S - 1; j >= 0; --j) {
		const SegInfo sg = a.seg[j];
		const bool last = (j == a.S - 1);
		const int kind = sg.kind;  // host-selected code path: 0 generic, 1 STD
		const int nc = sg.nc;
#define BWD_CASE(NCV, KINDV) bwd_segment<NCV, KINDV, STORE>(a, sm, sg, j, rd, off, len, lw, x_term, last, bw, sb)
#ifndef TDG_NO_PAIRS
// wider segments keep more state per HMM: two HMMs per loop from 7 columns on
#define BWD_PAIR(NCV) do { constexpr int G_ = (NCV) <= 6 ? TDG_BWD_GROUP : 2; \
		if (sg.nh >= G_) bwd_group_std<NCV, STORE, G_>(a, sm, sg, j, rd, off, len, lw, x_term, last, bw, sb); \
		else if (sg.nh >= 2) bwd_group_std<NCV, STORE, 2>(a, sm, sg, j, rd, off, len, lw, x_term, last, bw, sb); \
		else BWD_CASE(NCV, 1); } while (0)
#else
#define BWD_PAIR(NCV) BWD_CASE(NCV, 1)
#endif
// column-loop paths: profile state in shared memory when the host reserved room for this many columns
#define BWD_LOOP(KINDV)                                                                                          \
	do {                                                                                                         \
		if (a.dyn_cols >= nc) bwd_segment<0, KINDV, STORE, true>(a, sm, sg, j, rd, off, len, lw, x_term, last, bw, sb); \
		else BWD_CASE(0, KINDV);                                                                                 \
	} while (0)
		if (kind == 1) {
			switch (nc) {
				case 3: BWD_CASE(3, 1); break;
				case 4: BWD_PAIR(4); break;
				case 5: BWD_PAIR(5); break;
				case 6: BWD_PAIR(6); break;
				case 7: BWD_PAIR(7); break;
				case 8: BWD_PAIR(8); break;
				case 9: BWD_CASE(9, 1); break;
				case 10: BWD_CASE(10, 1); break;
				case 11: BWD_CASE(11, 1); break;
				case 12: BWD_CASE(12, 1); break;
				case 13: BWD_CASE(13, 1); break;
				case 14: BWD_CASE(14, 1); break;
				case 15: BWD_CASE(15, 1); break;
				case 16: BWD_CASE(16, 1); break;
				default: BWD_LOOP(1); break;  // standard pattern, more than kMaxStdCols columns: column loop
			}
		} else {
			switch (nc) {
				case 1: BWD_CASE(1, 0); break;
				case 2: BWD_CASE(2, 0); break;
				case 3: BWD_CASE(3, 0); break;
				case 4: BWD_CASE(4, 0); break;
				case 5: BWD_CASE(5, 0); break;
				case 6: BWD_CASE(6, 0); break;
				case 7: BWD_CASE(7, 0); break;
				case 8: BWD_CASE(8, 0); break;
				default: BWD_LOOP(0); break;
			}
		}
#undef BWD_CASE
#undef BWD_PAIR
#undef BWD_LOOP
	}
	if (valid) a.b_score[read] = sb[(size_t)1 * kBlock];  // model[0]->silent_backward[1] (:3610)
}

// ------------------------------------------------------------------------------------------
// forward + posterior, one segment.  barcode_hmm.c:4199-4345
// Mb/Ib of position i+1 (HBM), cs[i+1] and ps[i+1] are loaded one position ahead.
// ------------------------------------------------------------------------------------------
template <bool V> struct DeferTag { static constexpr bool value = V; };

template <int NC, int KIND, bool SMS = false>
__device__ __forceinline__ void fwd_segment(const KArgs& a, const Smem& sm, const SegInfo sg, int j,
                                            const SeqReader& rd, int off, int len, int lw, float B,
                                            const float2* __restrict__ bw, const float* __restrict__ sbk,
                                            float* __restrict__ sf, float* __restrict__ post, float* __restrict__ tp,
                                            uint32_t* __restrict__ prange, int f_begin = 0)
{
	constexpr bool STD = KIND == 1;
	constexpr int N = Cols<NC, STD>::N;
	constexpr int UN = NC > 0 ? N : 1;
	constexpr int NB = NC > 0 ? (STD ? NC - 1 : NC) : 1;  // prefetch registers only on the unrolled paths
	// Posterior window (unrolled standard segments): P[i][h] is only ever used through exp(), which is exactly 0 below
	// -104 (post_exp in k_label), and it is a logsum chain of 2 NC - 1 terms.  A logsum is at most its larger operand
	// + table[0] = ln 2, so if every term is below kPThr = -104 - (2 NC - 2) ln 2 the chain ends below -104 whatever its
	// bits: the 2 NC - 2 logsums (a third of k_forward's) are skipped and -inf is stored.  Otherwise the chain is run
	// afterwards on the same operands in the same order.  A NaN term fails the `<` and takes the chain.
#ifdef TDG_NO_PWIN
	constexpr bool PW = false;
#else
	constexpr bool PW = STD && NC > 0 && !SMS;
#endif
	constexpr float kPThr = -104.0f - 0.6932f * (float)(2 * (NC > 0 ? NC : 1) - 2) - 0.01f;
	const int nc = NC > 0 ? NC : sg.nc;
	const int ncs = STD ? nc - 1 : nc;  // stored columns (k_backward does not store the last STDU column)
	const bool last_seg = (j == a.S - 1);
	const size_t W = (size_t)(a.lmax + 2);
	float* cs_arr = sf + ((size_t)j * W) * kBlock;
	const float* ps_arr = sf + ((size_t)(j - 1) * W) * kBlock;  // only dereferenced when j > 0
	const TabAddr tab = sm.tab;
	const bool first_seg = (j == 0);
	const int skip_live = sg.skip_live;
	const size_t bstep = (size_t)ncs * kBlock;
	const uint64_t keep = make_keep_policy();

	for (int f = f_begin; f < sg.nh; ++f) {
		const int h = sg.hmmbase + f;
		const int c0 = sg.colbase + f * nc;
		const float* rec = sm.colrec + (size_t)c0 * kColRec;
		const float* em = sm.emit + (size_t)c0 * kEmitRec;
		const float2* bwq = bw + (size_t)c0 * a.lmax * kBlock;  // -> (position 1, column 0); walks ahead of i
		typename StateVec<N, SMS>::type M, I;
		M.bind(sm.dyn); I.bind(sm.dyn + (size_t)a.dyn_cols * kBlock);
		float2 bn[NB];
#pragma unroll UN
		for (int g = 0; g < (NC > 0 ? N : nc); ++g) {
			if (g < nc) { M[g] = NEG_INF; I[g] = NEG_INF; }
		}
		if (NC > 0) {
#pragma unroll
			for (int g = 0; g < NB; ++g) bn[g] = ld_bw(&bwq[(size_t)g * kBlock]);
		}
		const float sM0 = STD ? rec[F_SM] : 0.0f;
		// column-loop path: Mb/Ib are fetched one stored column ahead of their use (column 0 of the next position
		// while the last column of this one is computed), so the loads overlap the logsums instead of preceding them
		float2 b_ahead = make_float2(NEG_INF, NEG_INF);
		if (NC == 0) b_ahead = ld_bw(&bwq[0]);
		float TP = NEG_INF;
		int pfirst = 0xFFFF, plast = 0;  // positions whose posterior is >= -104 (exp != 0), for k_label
		float ps1 = first_seg ? 0.0f : ps_arr[0];  // psilent[0]
		SeqUp su;
		su.init(rd, off, a.words);
		float* csp = cs_arr + kBlock;         // -> cs[i]
		const float* psp = ps_arr + kBlock;   // -> ps[i]
		float* pp = post + (size_t)h * kBlock;  // -> posterior (position i, hmm h)
		float cs_n = ld_keep(csp, keep);
		float ps_n = first_seg ? NEG_INF : ld_keep(psp, keep);
		// STDU: silent_backward of the next segment at position i+1 (= Mb of the unstored last column)
		const float* qb = sbk + ((size_t)(j + 1) * W + 2) * kBlock;
		float q_n = (STD && !last_seg) ? ld_keep(qb, keep) : NEG_INF;
		bool defer = false;
		for (int i = 1; i <= lw; ++i) {
			const int x = su.get();  // seqa[i]
			float cs = cs_n;
			const float ps0 = ps_n;
			float q0 = NEG_INF;
			if (STD) {
				q0 = last_seg ? (i == len ? 0.0f : NEG_INF) : q_n;
				qb += kBlock;
				if (!last_seg && i < lw) q_n = ld_keep(qb, keep);
			}
			float2 bc[NB];
			if (NC > 0) {
				if (i < a.lmax) bwq += bstep;  // position i+1, clamped to the scratch
#pragma unroll
				for (int g = 0; g < NB; ++g) { bc[g] = bn[g]; bn[g] = ld_bw(&bwq[(size_t)g * kBlock]); }
			}
			cs_n = ld_keep(csp + kBlock, keep);
			if (!first_seg) ps_n = ld_keep(psp + kBlock, keep);
			// One position of this HMM.  DEFER = the posterior chain is decided after the recurrence (posterior window,
			// see kPThr); otherwise it is interleaved with the recurrence as the reference writes it.  The deferred chain,
			// when it has to run, is a bare dependent chain with nothing to overlap, so a warp only defers while all its
			// reads were outside the window at the previous position (`defer`, warp-uniform).
			auto position = [&](auto defer_tag) {
				constexpr bool DEFER = decltype(defer_tag)::value;
				float P;
				bool p_small = true;
				float oldMp, oldIp, newMp, D;
				const float eIu = STD ? em[5 + x] : 0.0f;  // STDU: one insert emission per position
				// ---- column 0 (:4218-4266)
				{
					const float* r = rec;
					const uint32_t lv = STD ? 0u : __float_as_uint(r[F_LIVE]);
					const float2 b = NC > 0 ? bc[0] : b_ahead;
					if (NC == 0 && ncs > 1) b_ahead = ld_bw(&bwq[((size_t)(i - 1) * ncs + 1) * kBlock]);
					const float eM = em[x], eI = STD ? eIu : em[5 + x];
					const bool lsm = live_of<STD>(nc, 0, F_SM, lv);
					const bool lsi = live_of<STD>(nc, 0, F_SI, lv);
					const float nM = lsm ? ps1 + (STD ? sM0 : r[F_SM]) + eM : NEG_INF;
					const float tM = nM + b.x - B;
					TP = LS(TP, tM, tab);
					P = tM;  // logsum(-inf, tM) == tM
					if (DEFER) p_small = tM < kPThr;
					float v;
					bool have = false;
					v = NEG_INF;
					if (lsi) { v = ps1 + r[F_SI]; have = true; }
					if (live_of<STD>(nc, 0, F_II, lv)) { const float t = I[0] + trv<KIND>(r, sg, nc, 0, F_II); v = have ? LS(v, t, tab) : t; have = true; }
					if (live_of<STD>(nc, 0, F_MI, lv)) { const float t = M[0] + trv<KIND>(r, sg, nc, 0, F_MI); v = have ? LS(v, t, tab) : t; have = true; }
					const float nI = v + eI;
					if (lsi) TP = LS(TP, ps1 + r[F_SI] + eI + b.y - B, tab);
					if (DEFER) p_small &= (nI + b.y - B) < kPThr;
					else P = LS(P, nI + b.y - B, tab);
					if (live_of<STD>(nc, 0, F_MSKIP, lv)) cs = LS(cs, nM + r[F_MSKIP], tab);
					if (live_of<STD>(nc, 0, F_ISKIP, lv)) cs = LS(cs, nI + r[F_ISKIP], tab);
					oldMp = M[0]; oldIp = I[0]; newMp = nM; D = NEG_INF;
					M[0] = nM; I[0] = nI;
				}
				// ---- columns 1 .. nc-1 (:4270-4331)
#pragma unroll UN
				for (int g = 1; g < (NC > 0 ? N : nc); ++g) {
					if (g < nc) {
						const int p = g - 1;
						const float* r = rec + g * kColRec;
						const float* rp = rec + p * kColRec;
						const uint32_t lv = STD ? 0u : __float_as_uint(r[F_LIVE]);
						const uint32_t lp = STD ? 0u : __float_as_uint(rp[F_LIVE]);
						float2 b;
						if (STD && g == nc - 1) b = make_float2(q0 + trv<KIND>(r, sg, nc, g, F_MSKIP), NEG_INF);  // as k_backward computes it
						else b = NC > 0 ? bc[(NC > 0 && !(STD && g == NC - 1)) ? g : 0] : b_ahead;
						if (NC == 0) {
							// next stored column of this position, or column 0 of the next position (clamped to the scratch)
							if (g + 1 < ncs) b_ahead = ld_bw(&bwq[((size_t)(i - 1) * ncs + g + 1) * kBlock]);
							else if (g + 1 == nc && i < a.lmax) b_ahead = ld_bw(&bwq[((size_t)i * ncs) * kBlock]);
						}
						const float eM = em[g * kEmitRec + x], eI = STD ? eIu : em[g * kEmitRec + 5 + x];
						const float oldMg = M[g], oldIg = I[g];
						float v; bool have;
						// M_forward[g][i]
						v = NEG_INF; have = false;
						if (live_of<STD>(nc, g, F_SM, lv)) { v = ps1 + r[F_SM]; have = true; }
						if (live_of<STD>(nc, p, F_MM, lp)) { const float t = oldMp + trv<KIND>(rp, sg, nc, p, F_MM); v = have ? LS(v, t, tab) : t; have = true; }
						if (live_of<STD>(nc, p, F_IM, lp)) { const float t = oldIp + trv<KIND>(rp, sg, nc, p, F_IM); v = have ? LS(v, t, tab) : t; have = true; }
						if (live_of<STD>(nc, p, F_DM, lp)) { const float t = D + trv<KIND>(rp, sg, nc, p, F_DM); v = have ? LS(v, t, tab) : t; have = true; }
						const float nM = v + eM;
						const bool m_reach = STD ? true : have;
						if (DEFER) p_small &= (nM + b.x - B) < kPThr;
						else if (m_reach) P = LS(P, nM + b.x - B, tab);
						// I_forward[g][i]
						v = NEG_INF; have = false;
						if (live_of<STD>(nc, g, F_SI, lv)) { v = ps1 + r[F_SI]; have = true; }
						if (live_of<STD>(nc, g, F_II, lv)) { const float t = oldIg + trv<KIND>(r, sg, nc, g, F_II); v = have ? LS(v, t, tab) : t; have = true; }
						if (live_of<STD>(nc, g, F_MI, lv)) { const float t = oldMg + trv<KIND>(r, sg, nc, g, F_MI); v = have ? LS(v, t, tab) : t; have = true; }
						const float nI = v + eI;
						if (DEFER) { if (have) p_small &= (nI + b.y - B) < kPThr; }
						else if (have) P = LS(P, nI + b.y - B, tab);
						// D_forward[g][i]
						{
							float dv = NEG_INF; bool dh = false;
							if (live_of<STD>(nc, p, F_MD, lp)) { dv = newMp + trv<KIND>(rp, sg, nc, p, F_MD); dh = true; }
							if (live_of<STD>(nc, p, F_DD, lp)) { const float t = D + trv<KIND>(rp, sg, nc, p, F_DD); dv = dh ? LS(dv, t, tab) : t; }
							D = dv;
						}
						if (live_of<STD>(nc, g, F_MSKIP, lv)) cs = LS(cs, nM + trv<KIND>(r, sg, nc, g, F_MSKIP), tab);
						if (live_of<STD>(nc, g, F_ISKIP, lv)) cs = LS(cs, nI + r[F_ISKIP], tab);
						oldMp = oldMg; oldIp = oldIg; newMp = nM;
						M[g] = nM; I[g] = nI;
					}
				}
				if (skip_live) cs = LS(cs, ps0 + sg.skip, tab);  // (:4341)
				st_keep(csp, cs, keep);
				if (DEFER) {
					if (p_small) P = NEG_INF;
					else {
						// the posterior chain (:4230-4330) on the states just computed, in the reference's order
						P = M[0] + bc[0].x - B;
						P = LS(P, I[0] + bc[0].y - B, tab);
#pragma unroll
						for (int g = 1; g < N; ++g) {
							const float bx = (g == NC - 1) ? q0 + 0.0f : bc[g < NB ? g : 0].x;
							P = LS(P, M[g] + bx - B, tab);
							if (g < NC - 1) P = LS(P, I[g] + bc[g < NB ? g : 0].y - B, tab);
						}
					}
				}
				// [pfirst, plast]: the window outside which exp(P) is exactly 0 (k_label skips HMMs without predecessors there)
				if (!(P < -104.0f)) { plast = i; pfirst = min(pfirst, i); }
				// k_label loads every posterior (no per-HMM window test), so all of them are stored.  The predicate is always
				// true (post_store_all = 1) but opaque to the compiler, and ptxas schedules the loop measurably better with
				// it than with a plain store: k_forward 7.24 vs 7.49 ms per wave in round 1, 7.47 vs 7.68 ms in round 2
				// (same box, back to back, scripts/gpu_ab.sh postalways; -DTDG_POST_ALWAYS builds the plain store).
#ifdef TDG_POST_ALWAYS
				__stcs(pp, P);
#else
				if (pfirst != 0xFFFF || a.post_store_all) __stcs(pp, P);
#endif
				ps1 = ps0;
				return P < -104.0f;
			};
			bool outside = true;   // this read is outside the posterior window at this position (or has ended)
			if (i <= len) {
				if (PW && defer) outside = position(DeferTag<PW>{});
				else outside = position(DeferTag<false>{});
			}
			if (PW) defer = __all_sync(0xffffffffu, outside) != 0;
			csp += kBlock; psp += kBlock; pp += (size_t)a.H * kBlock;
		}
		tp[(size_t)h * kBlock] = TP;
		prange[(size_t)h * kBlock] = ((uint32_t)plast << 16) | (uint32_t)pfirst;
	}
}

// ------------------------------------------------------------------------------------------
// forward + posterior, standard-pattern segment, TWO HMMs per position loop (see bwd_pair_std): cs[i], ps[i] and the
// silent_backward value of the next segment are loaded once per pair, the Mb/Ib of HMM f+1 are fetched while HMM f is
// computed and those of HMM f for the next position while HMM f+1 is computed (one buffer per HMM instead of two).
// MEASURED AND NOT ADOPTED (round 2, same box back to back): bit-identical, but k_forward 7.65-7.69 ms per wave against
// 7.39-7.40 ms for the one-HMM loop -- k_forward is bound by issue slots, not by HBM, the shared loads are a small part
// of its instructions and two HMMs' state plus two Mb/Ib buffers do not fit 128 registers without spills.  Built only with
// -DTDG_FWD_PAIRS (scripts/build_variant.sh); the backward pass, which IS bound by HBM, uses its pair loop by default.
// ------------------------------------------------------------------------------------------
#if defined(TDG_FWD_PAIRS) && !defined(TDG_NO_PAIRS)
template <int NC>
__device__ __forceinline__ void fwd_pair_std(const KArgs& a, const Smem& sm, const SegInfo sg, int j,
                                             const SeqReader& rd, int off, int len, int lw, float B,
                                             const float2* __restrict__ bw, const float* __restrict__ sbk,
                                             float* __restrict__ sf, float* __restrict__ post, float* __restrict__ tp,
                                             uint32_t* __restrict__ prange)
{
	constexpr int NB = NC - 1;   // stored columns
	const bool last_seg = (j == a.S - 1);
	const bool first_seg = (j == 0);
	const size_t W = (size_t)(a.lmax + 2);
	float* cs_arr = sf + ((size_t)j * W) * kBlock;
	const float* ps_arr = sf + ((size_t)(j - 1) * W) * kBlock;  // only dereferenced when j > 0
	const TabAddr tab = sm.tab;
	const int skip_live = sg.skip_live;
	const size_t bstep = (size_t)NB * kBlock;
	const uint64_t keep = make_keep_policy();
	const float ta = sg.ta, tb = sg.tb, tb2 = sg.tb2, tc = sg.tc, td = sg.td;

	for (int f = 0; f + 1 < sg.nh; f += 2) {
		const int h = sg.hmmbase + f;
		const int c0 = sg.colbase + f * NC;
		const float* emA = sm.emit + (size_t)c0 * kEmitRec;
		const float* emB = emA + (size_t)NC * kEmitRec;
		const float sM0A = (sm.colrec + (size_t)c0 * kColRec)[F_SM];
		const float sM0B = (sm.colrec + (size_t)(c0 + NC) * kColRec)[F_SM];
		const float2* bwqA = bw + (size_t)c0 * a.lmax * kBlock;          // -> (position 1, column 0) of HMM f
		const float2* bwqB = bwqA + (size_t)NC * a.lmax * kBlock;        //    ... of HMM f + 1
		float MA[NC], IA[NC], MB[NC], IB[NC];
		float2 bA[NB], bB[NB];
#pragma unroll
		for (int g = 0; g < NC; ++g) { MA[g] = NEG_INF; IA[g] = NEG_INF; MB[g] = NEG_INF; IB[g] = NEG_INF; }
#pragma unroll
		for (int g = 0; g < NB; ++g) bA[g] = ld_bw(&bwqA[(size_t)g * kBlock]);
		float TPA = NEG_INF, TPB = NEG_INF;
		int pfA = 0xFFFF, plA = 0, pfB = 0xFFFF, plB = 0;
		float ps1 = first_seg ? 0.0f : ps_arr[0];
		SeqUp su;
		su.init(rd, off, a.words);
		float* csp = cs_arr + kBlock;
		const float* psp = ps_arr + kBlock;
		float* pp = post + (size_t)h * kBlock;
		float cs_n = ld_keep(csp, keep);
		float ps_n = first_seg ? NEG_INF : ld_keep(psp, keep);
		const float* qb = sbk + ((size_t)(j + 1) * W + 2) * kBlock;
		float q_n = (!last_seg) ? ld_keep(qb, keep) : NEG_INF;
		for (int i = 1; i <= lw; ++i) {
			const int x = su.get();
			float cs = cs_n;
			const float ps0 = ps_n;
			const float q0 = last_seg ? (i == len ? 0.0f : NEG_INF) : q_n;
			qb += kBlock;
			if (!last_seg && i < lw) q_n = ld_keep(qb, keep);
			// HMM f + 1, this position: in flight while HMM f is computed
#pragma unroll
			for (int g = 0; g < NB; ++g) bB[g] = ld_bw(&bwqB[(size_t)g * kBlock]);
			cs_n = ld_keep(csp + kBlock, keep);
			if (!first_seg) ps_n = ld_keep(psp + kBlock, keep);
			const bool act = i <= len;
			const float eIu = emA[5 + x];
			// one HMM at this position (the same expressions as fwd_segment<NC, 1>)
			auto hmm = [&](float (&M)[NC], float (&I)[NC], const float2 (&b)[NB], const float* em, float sM0, float& TP, float& P) {
				float oldMp, oldIp, newMp, D;
				{
					const float nM = ps1 + sM0 + em[x];
					const float tM = nM + b[0].x - B;
					TP = LS(TP, tM, tab);
					P = tM;
					float v = I[0] + tc;                       // II
					v = LS(v, M[0] + tb, tab);                 // MI
					const float nI = v + eIu;
					P = LS(P, nI + b[0].y - B, tab);
					oldMp = M[0]; oldIp = I[0]; newMp = nM; D = NEG_INF;
					M[0] = nM; I[0] = nI;
				}
#pragma unroll
				for (int g = 1; g < NC; ++g) {
					const int p = g - 1;
					float2 bg;
					if (g == NC - 1) bg = make_float2(q0 + 0.0f, NEG_INF);   // the unstored last column, as k_backward computes it
					else bg = b[g < NB ? g : 0];
					const float eM = em[g * kEmitRec + x];
					const float oldMg = M[g], oldIg = I[g];
					float v = oldMp + ta;                                      // MM of column p
					v = LS(v, oldIp + td, tab);                                // IM
					if (p >= 1) v = LS(v, D + (p == NC - 2 ? 0.0f : td), tab);   // DM
					const float nM = v + eM;
					P = LS(P, nM + bg.x - B, tab);
					float nI = NEG_INF;
					if (g < NC - 1) {
						float w = oldIg + tc;                                  // II
						w = LS(w, oldMg + (g == NC - 2 ? tb2 : tb), tab);      // MI
						nI = w + eIu;
						P = LS(P, nI + bg.y - B, tab);
					} else {
						nI = NEG_INF + eIu;
					}
					{
						float dv = NEG_INF; bool dh = false;
						if (p <= NC - 3) { dv = newMp + tb; dh = true; }       // MD
						if (p >= 1 && p <= NC - 3) { const float t = D + tc; dv = dh ? LS(dv, t, tab) : t; }   // DD
						D = dv;
					}
					if (g == NC - 1) cs = LS(cs, nM + 0.0f, tab);              // MSKIP of the last column
					oldMp = oldMg; oldIp = oldIg; newMp = nM;
					M[g] = nM; I[g] = nI;
				}
				if (skip_live) cs = LS(cs, ps0 + sg.skip, tab);
			};
			float PA = NEG_INF, PB = NEG_INF;
			if (act) hmm(MA, IA, bA, emA, sM0A, TPA, PA);
			// HMM f, next position: in flight while HMM f + 1 is computed
			if (i < a.lmax) bwqA += bstep;
#pragma unroll
			for (int g = 0; g < NB; ++g) bA[g] = ld_bw(&bwqA[(size_t)g * kBlock]);
			if (act) {
				hmm(MB, IB, bB, emB, sM0B, TPB, PB);
				st_keep(csp, cs, keep);
				if (!(PA < -104.0f)) { plA = i; pfA = min(pfA, i); }
				if (!(PB < -104.0f)) { plB = i; pfB = min(pfB, i); }
				if (pfA != 0xFFFF || a.post_store_all) __stcs(pp, PA);
				if (pfB != 0xFFFF || a.post_store_all) __stcs(pp + kBlock, PB);
				ps1 = ps0;
			}
			if (i < a.lmax) bwqB += bstep;
			csp += kBlock; psp += kBlock; pp += (size_t)a.H * kBlock;
		}
		tp[(size_t)h * kBlock] = TPA;
		tp[(size_t)(h + 1) * kBlock] = TPB;
		prange[(size_t)h * kBlock] = ((uint32_t)plA << 16) | (uint32_t)pfA;
		prange[(size_t)(h + 1) * kBlock] = ((uint32_t)plB << 16) | (uint32_t)pfB;
	}
	if (sg.nh & 1) fwd_segment<NC, 1>(a, sm, sg, j, rd, off, len, lw, B, bw, sbk, sf, post, tp, prange, sg.nh - 1);
}
#endif

__device__ __forceinline__ float s2p_f(float p)
{
	// scaledprob2prob (misc.c:98-105): float argument, double exp, float result
	if (p == NEG_INF) return 0.0f;
	return (float)exp((double)p);
}

template <bool GM = false>
__global__ void __launch_bounds__(kBlock, 512 / kBlock) k_forward(const KArgs a)
{
	extern __shared__ float smem_f[];
	const Smem sm = stage_smem<GM>(a, smem_f);
	const int slot = blockIdx.x * kBlock + threadIdx.x;
	const int read = slot;
	const bool valid = read < a.n_reads;
	int len = 0, off = 0;
	if (valid) {
		len = a.len[read];
		if (a.win_len >= 0) { off = a.win_start; len = a.win_len; }
	}
	const int lw = __reduce_max_sync(0xffffffffu, len);
	const SeqReader rd = make_reader(a, valid ? read : 0);
	const size_t W = (size_t)(a.lmax + 2);
	const float2* bw = a.bw + (size_t)blockIdx.x * ((size_t)a.C * a.lmax) * kBlock + threadIdx.x;
	const float* sbk = a.sb + (size_t)blockIdx.x * ((size_t)a.S * W) * kBlock + threadIdx.x;
	float* sf = a.sf + (size_t)blockIdx.x * ((size_t)a.S * W) * kBlock + threadIdx.x;
	float* post = a.post + (size_t)blockIdx.x * ((size_t)a.lmax * a.H) * kBlock + threadIdx.x;
	float* tp = a.tp + (size_t)blockIdx.x * ((size_t)a.H) * kBlock + threadIdx.x;
	uint32_t* prange = a.prange + (size_t)blockIdx.x * ((size_t)a.H) * kBlock + threadIdx.x;
	const float B = valid ? a.b_score[read] : 0.0f;

	// init silent_forward (:4166-4175)
	for (int j = 0; j < a.S; ++j)
		for (int i = 0; i <= len + 1; ++i) sf[((size_t)j * W + i) * kBlock] = NEG_INF;
	{
		float v = 0.0f + a.seg[0].skip;
		sf[0] = v;
		for (int j = 1; j < a.S; ++j) {
			v = v + a.seg[j].skip;
			sf[((size_t)j * W) * kBlock] = v;
		}
	}
	for (int j = 0; j < a.S; ++j) {
		const SegInfo sg = a.seg[j];
		const int kind = sg.kind;
		const int nc = sg.nc;
#define FWD_CASE(NCV, KINDV) fwd_segment<NCV, KINDV>(a, sm, sg, j, rd, off, len, lw, B, bw, sbk, sf, post, tp, prange)
#if defined(TDG_FWD_PAIRS) && !defined(TDG_NO_PAIRS)
#define FWD_PAIR(NCV) do { if (sg.nh >= 2) fwd_pair_std<NCV>(a, sm, sg, j, rd, off, len, lw, B, bw, sbk, sf, post, tp, prange); else FWD_CASE(NCV, 1); } while (0)
#else
#define FWD_PAIR(NCV) FWD_CASE(NCV, 1)
#endif
#define FWD_LOOP(KINDV)                                                                              \
	do {                                                                                             \
		if (a.dyn_cols >= nc) fwd_segment<0, KINDV, true>(a, sm, sg, j, rd, off, len, lw, B, bw, sbk, sf, post, tp, prange);  \
		else FWD_CASE(0, KINDV);                                                                     \
	} while (0)
		if (kind == 1) {
			switch (nc) {
				case 3: FWD_CASE(3, 1); break;
				case 4: FWD_PAIR(4); break;
				case 5: FWD_PAIR(5); break;
				case 6: FWD_PAIR(6); break;
				case 7: FWD_PAIR(7); break;
				case 8: FWD_PAIR(8); break;
				case 9: FWD_CASE(9, 1); break;
				case 10: FWD_CASE(10, 1); break;
				case 11: FWD_CASE(11, 1); break;
				case 12: FWD_CASE(12, 1); break;
				case 13: FWD_CASE(13, 1); break;
				case 14: FWD_CASE(14, 1); break;
				case 15: FWD_CASE(15, 1); break;
				case 16: FWD_CASE(16, 1); break;
				default: FWD_LOOP(1); break;
			}
		} else {
			switch (nc) {
				case 1: FWD_CASE(1, 0); break;
				case 2: FWD_CASE(2, 0); break;
				case 3: FWD_CASE(3, 0); break;
				case 4: FWD_CASE(4, 0); break;
				case 5: FWD_CASE(5, 0); break;
				case 6: FWD_CASE(6, 0); break;
				case 7: FWD_CASE(7, 0); break;
				case 8: FWD_CASE(8, 0); break;
				default: FWD_LOOP(0); break;
			}
		}
#undef FWD_CASE
#undef FWD_PAIR
#undef FWD_LOOP
	}
	if (!valid) return;
	const TabAddr tab = sm.tab;
	const float f_score = sf[((size_t)(a.S - 1) * W + len) * kBlock];  // (:4349)

	// total_prob normalisation + bar_prob (:4354-4429); next_silent[0] is never reset.
	{
		int h = 0;
		for (int j = 0; j < a.S; ++j) {
			const int nh = a.seg[j].nh;
			if (nh > 1) {
				float s = NEG_INF;
				for (int f = 0; f < nh; ++f) s = LS(s, tp[(size_t)(h + f) * kBlock], tab);
				for (int f = 0; f < nh; ++f) tp[(size_t)(h + f) * kBlock] = tp[(size_t)(h + f) * kBlock] - s;
			}
			h += nh;
		}
	}
	float bar_prob;
	{
		int h = 0, g = 1;
		float n0 = NEG_INF, n2 = 0.0f;
		for (int j = 0; j < a.S; ++j) {
			const int nh = a.seg[j].nh;
			if (nh > 1) {
				g = 0;
				float n1 = NEG_INF;
				for (int f = 0; f < nh; ++f) {
					const float t = tp[(size_t)(h + f) * kBlock];
					if (t > n0 && f != nh - 1) n0 = t;
					n1 = LS(n1, t, tab);
				}
				n0 = n0 - n1;
				n2 = n2 + n0;
			}
			h += nh;
		}
		bar_prob = g ? 0.0f : ((n2 > 0.0f) ? 0.0f : n2);
	}
	// random model (:4516-4523)
	float r = 0.0f;
	for (int i = 1; i <= len; ++i) {
		const int c = rd.code(off + i - 1);
		r = r + a.bg[c < 5 ? c : 4] + a.r_step;
	}
	r += a.r_end;
	// Q (:2320-2338): pbest = logsum(logsum(-inf, f), r); the bar_prob sum is evaluated in double
	float pbest = LS(NEG_INF, f_score, tab);
	pbest = LS(pbest, r, tab);
	const float arg = (float)(((double)bar_prob + (double)f_score) - (double)pbest);
	pbest = (float)(1.0 - (double)s2p_f(arg));
	float Q;
	if (!pbest) Q = 40.0f;
	else if (pbest == 1.0f) Q = 0.0f;
	else Q = (float)(-10.0 * log10((double)pbest));
	a.f_score[read] = f_score;
	a.r_score[read] = r;
	a.bar_prob[read] = bar_prob;
	a.mapq[read] = Q;
}

// ------------------------------------------------------------------------------------------
// k_label: exp of posteriors (:4431-4440), constrained label DP (:4447-4472), traceback
// (:4494-4514), extract_reads (:3172-3313).  Thread-per-read; the posterior matrix is
// rewritten in place with the DP scores.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float post_exp(float x)
{
	// (float)exp((double)x) is exactly 0 below 2^-150 (x < -103.98): skip the double exp there.
	if (x < -104.0f) return 0.0f;
	return (float)exp((double)x);
}

__global__ void __launch_bounds__(kDpBlock) k_label(const KArgs a)
{
	__shared__ int s_src[TDG_MAX_HMMS_DEV * kMaxSources];
	__shared__ uint8_t s_segflag[kMaxSegments];  // bit 0: no HMM of the segment has a predecessor but itself; bit 1: all share one source list
	__shared__ int s_hlabel[TDG_MAX_HMMS_DEV];    // mb->label
	__shared__ uint8_t s_stype[kMaxSegments];     // read_structure->type
	extern __shared__ float dsm[];  // structured path: D row [H][bs] floats
	const int bs = blockDim.x;
	for (int k = threadIdx.x; k < a.H * kMaxSources; k += bs) s_src[k] = a.dp_src[k];
	for (int k = threadIdx.x; k < a.H; k += bs) s_hlabel[k] = a.hmm_label[k];
	if ((int)threadIdx.x < a.S) s_stype[threadIdx.x] = a.seg_type[threadIdx.x];
	__syncthreads();
	if ((int)threadIdx.x < a.S) {
		const int hb = a.seg[threadIdx.x].hmmbase, nh = a.seg[threadIdx.x].nh;
		bool self_only = true, uniform = true;
		for (int f = 0; f < nh; ++f)
			for (int q = 0; q < kMaxSources; ++q) {
				const int src = s_src[(hb + f) * kMaxSources + q];
				if (q == 0 && src != INT32_MIN) self_only = false;
				if (src != s_src[hb * kMaxSources + q]) uniform = false;
				if (src != INT32_MIN && src >= hb) uniform = false;  // a predecessor inside the segment is updated in place
			}
		s_segflag[threadIdx.x] = (self_only ? 1 : 0) | (uniform ? 2 : 0);
	}
	__syncthreads();
	const int slot = blockIdx.x * bs + threadIdx.x;
	const int read = slot;
	if (read >= a.n_reads) return;
	const int cta = slot / kBlock, t = slot % kBlock;
	int len = a.len[read], off = 0;
	const int rlen = len;
	if (a.win_len >= 0) { off = a.win_start; len = a.win_len; }
	const int H = a.H;
	float* post = a.post + (size_t)cta * ((size_t)a.lmax * H) * kBlock + t;
	uint8_t* path = a.path + (size_t)cta * ((size_t)a.lmax * H) * kBlock + t;
	uint8_t* labels = a.labels + (size_t)read * a.label_stride;
	// Labels of this read are staged in shared memory ([position][thread] bytes) when they fit: traceback and
	// extraction then run without global round trips, and the row goes out as 32-bit words at the end.
	uint8_t* slab = (uint8_t*)(dsm + (a.dp_structured ? (size_t)a.H * bs : 0)) + threadIdx.x;
	const bool lsm = a.label_smem != 0;
	auto lab_set = [&](int i, int v) { if (lsm) slab[(size_t)i * bs] = (uint8_t)v; else labels[i] = (uint8_t)v; };
	auto lab_get = [&](int i) -> int { return lsm ? (int)slab[(size_t)i * bs] : (int)labels[i]; };
	auto lab_flush = [&](int n) {  // n labels -> global row (label_stride is a multiple of 8, rows are 8-byte aligned)
		if (!lsm || !a.store_labels) return;
		for (int i = 0; i < n; i += 4) {
			uint32_t w = 0;
#pragma unroll
			for (int k = 0; k < 4; ++k) if (i + k < n) w |= (uint32_t)slab[(size_t)(i + k) * bs] << (8 * k);
			*reinterpret_cast<uint32_t*>(labels + i) = w;
		}
	};

	if (a.want_labels) {
		float segmax[kMaxSegments];
		int segarg[kMaxSegments];
		int move = -1;
		// row 0: exp(-inf) = 0 everywhere
		for (int s = 0; s < a.S; ++s) { segmax[s] = 0.0f; segarg[s] = a.seg[s].hmmbase; }
		if (a.dp_structured) {
			// The DP row lives in shared memory (4 H bytes per read: four 128-thread CTAs per SM) and is
			// updated in place from the highest HMM index down (every predecessor of j has a lower index,
			// so it still holds row i-1).  Posterior entries outside [first,last] of their HMM
			// (k_forward's prange) are < -104 and exp to exactly 0.
			//
			// HMMs whose only predecessor is themselves (the first segment): D[i][j] = D[i-1][j] + p
			// and path[i][j] = j, so outside their range nothing changes and no path byte is kept;
			// positions outside the union of the segment's ranges skip the segment altogether.  All D
			// entries are non-decreasing in i (p >= 0), so the segment's first-argmax is maintained
			// incrementally: an entry that rises above the maximum, or ties it at a lower index, takes over.
			constexpr int CH = 8;
			float* D = dsm + threadIdx.x;
			const uint32_t* prange = a.prange + (size_t)cta * H * kBlock + t;
			int sfirst[kMaxSegments], slast[kMaxSegments];  // union of the posterior ranges of a segment's HMMs
			for (int s = 0; s < a.S; ++s) {
				const int hb = a.seg[s].hmmbase, nh = a.seg[s].nh;
				int lo = 0xFFFF, hi = 0;
				for (int f = 0; f < nh; ++f) {
					const int j = hb + f;
					const uint32_t r = prange[(size_t)j * kBlock];
					D[(size_t)j * bs] = 0.0f;
					lo = min(lo, (int)(r & 0xFFFFu)); hi = max(hi, (int)(r >> 16));
				}
				sfirst[s] = lo; slast[s] = hi;
			}
			for (int i = 1; i <= len; ++i) {
				const float* row = post + ((size_t)(i - 1) * H) * kBlock;
				uint8_t* prow_path = path + ((size_t)(i - 1) * H) * kBlock;
				float nsegmax[kMaxSegments];
				int nsegarg[kMaxSegments];
				for (int s = a.S - 1; s >= 0; --s) {
					const int hb = a.seg[s].hmmbase, nh = a.seg[s].nh;
					const int flag = s_segflag[s];
					if (flag & 1) {
						float cm = segmax[s]; int ca = segarg[s];
						if (i >= sfirst[s] && i <= slast[s]) {
							for (int f1 = 0; f1 < nh; f1 += CH) {
								float pv[CH];
#pragma unroll
								for (int k = 0; k < CH; ++k) {
									const int f = f1 + k;
									if (f < nh) {
										pv[k] = __ldcs(&row[(size_t)(hb + f) * kBlock]);  // < -104 outside the HMM's own window
									}
								}
#pragma unroll
								for (int k = 0; k < CH; ++k) {
									const int f = f1 + k;
									if (f < nh && pv[k] >= -104.0f) {
										const int j = hb + f;
										const float nd = post_exp(pv[k]) + D[(size_t)j * bs];
										D[(size_t)j * bs] = nd;
										if (nd > cm || (nd == cm && j < ca)) { cm = nd; ca = j; }
									}
								}
							}
						}
						nsegmax[s] = cm; nsegarg[s] = ca;
						continue;
					}
					float ubest = -1.0f; int uarg = -1;
					if (flag & 2) {  // one source list for the whole segment: evaluate it once
						for (int q = 0; q < kMaxSources; ++q) {
							const int src = s_src[hb * kMaxSources + q];
							if (src == INT32_MIN) break;
							float v; int av;
							if (src < 0) { v = segmax[-src - 1]; av = segarg[-src - 1]; }
							else { v = D[(size_t)src * bs]; av = src; }
							if (v > ubest) { ubest = v; uarg = av; }
						}
					}
					float cm = -1.0f; int ca = hb;
					if (flag & 2) {
						// One source list for the whole segment (the usual case: "everything in the previous
						// segment").  HMMs are folded from the highest index down in full chunks of CH with one
						// base address per chunk (constant offsets, no per-element bounds tests); the posteriors of
						// the next chunk are fetched while this one is folded; where no posterior of a chunk
						// survives exp() the update is max(self, best) without the double-precision exp; outside
						// the union of the segment's posterior windows nothing is loaded at all.
						const bool inwin = (i >= sfirst[s] && i <= slast[s]);
						auto fold = [&](int j, float p, float self) {
							const bool keep = self >= ubest;
							const float mx = keep ? self : ubest;
							const float nd = post_exp(p) + mx;
							D[(size_t)j * bs] = nd;
							__stcs(&prow_path[(size_t)j * kBlock], (uint8_t)(keep ? j : uarg));
							if (nd >= cm) { cm = nd; ca = j; }  // descending scan: >= leaves the first (lowest) maximum
						};
						int f1 = nh;
						for (; f1 > (nh / CH) * CH; --f1) {  // the nh % CH highest HMMs one by one
							const int j = hb + f1 - 1;
							fold(j, inwin ? __ldcs(&row[(size_t)j * kBlock]) : NEG_INF, D[(size_t)j * bs]);
						}
						float pn[CH];
						if (f1 > 0) {
							const float* rb = row + (size_t)(hb + f1 - CH) * kBlock;
#pragma unroll
							for (int k = 0; k < CH; ++k) pn[k] = inwin ? __ldcs(&rb[(size_t)(CH - 1 - k) * kBlock]) : NEG_INF;
						}
						for (; f1 > 0; f1 -= CH) {
							const int jb = hb + f1 - CH;  // lowest HMM of the chunk; element k is HMM jb + CH-1-k
							float* Db = D + (size_t)jb * bs;
							uint8_t* pb = prow_path + (size_t)jb * kBlock;
							float pv[CH], sv[CH];
							bool any = false;
#pragma unroll
							for (int k = 0; k < CH; ++k) {
								pv[k] = pn[k];
								sv[k] = Db[(size_t)(CH - 1 - k) * bs];
								any |= !(pv[k] < -104.0f);  // NaN posteriors (degenerate reads) take the exact path too
							}
							if (f1 > CH) {
								const float* rb = row + (size_t)(jb - CH) * kBlock;
#pragma unroll
								for (int k = 0; k < CH; ++k) pn[k] = inwin ? __ldcs(&rb[(size_t)(CH - 1 - k) * kBlock]) : NEG_INF;
							}
							if (!any) {
#pragma unroll
								for (int k = 0; k < CH; ++k) {
									const int j = jb + CH - 1 - k;
									const bool keep = sv[k] >= ubest;
									const float nd = keep ? sv[k] : ubest;   // exp() == 0: 0.0f + mx is mx
									if (!keep) Db[(size_t)(CH - 1 - k) * bs] = nd;
									__stcs(&pb[(size_t)(CH - 1 - k) * kBlock], (uint8_t)(keep ? j : uarg));
									if (nd >= cm) { cm = nd; ca = j; }
								}
							} else {
#pragma unroll
								for (int k = 0; k < CH; ++k) fold(jb + CH - 1 - k, pv[k], sv[k]);
							}
						}
						nsegmax[s] = cm; nsegarg[s] = ca;
						continue;
					}
					for (int f1 = nh; f1 > 0; f1 -= CH) {
						const int f0 = f1 > CH ? f1 - CH : 0;
						float pv[CH];
#pragma unroll
						for (int k = 0; k < CH; ++k) {
							const int f = f1 - 1 - k;
							if (f >= f0) {
								const int j = hb + f;
								pv[k] = __ldcs(&row[(size_t)j * kBlock]);  // outside the HMM's window the posterior is < -104: post_exp() gives 0
							}
						}
#pragma unroll
						for (int k = 0; k < CH; ++k) {
							const int f = f1 - 1 - k;
							if (f >= f0) {
								const int j = hb + f;
								float best = ubest; int arg = uarg;
								if (!(flag & 2)) {
									for (int q = 0; q < kMaxSources; ++q) {
										const int src = s_src[j * kMaxSources + q];
										if (src == INT32_MIN) break;
										float v; int av;
										if (src < 0) { v = segmax[-src - 1]; av = segarg[-src - 1]; }
										else { v = D[(size_t)src * bs]; av = src; }
										if (v > best) { best = v; arg = av; }
									}
								}
								const float self = D[(size_t)j * bs];
								float mx; int mv;
								if (self >= best) { mx = self; mv = j; } else { mx = best; mv = arg; }
								const float nd = post_exp(pv[k]) + mx;
								D[(size_t)j * bs] = nd;
								__stcs(&prow_path[(size_t)j * kBlock], (uint8_t)mv);
								if (nd >= cm) { cm = nd; ca = j; }  // descending scan: >= leaves the first (lowest) maximum
							}
						}
					}
					nsegmax[s] = cm; nsegarg[s] = ca;
				}
				for (int s = 0; s < a.S; ++s) { segmax[s] = nsegmax[s]; segarg[s] = nsegarg[s]; }
			}
			// final argmax, first max wins (:4494-4501)
			float mx = -1.0f;
			for (int j = 0; j < H; ++j) {
				const float v = D[(size_t)j * bs];
				if (v > mx) { mx = v; move = j; }
			}
			// traceback (:4503-4514); an HMM without predecessors keeps the path on itself
			for (int i = 0; i <= rlen; ++i) lab_set(i, 0);
			if (move < 0) move = 0;
			lab_set(len, move);
			for (int i = len; i > 0; --i) {
				if (s_src[move * kMaxSources] != INT32_MIN) move = path[((size_t)(i - 1) * H + move) * kBlock];
				lab_set(i - 1, move);
			}
			lab_flush(rlen + 1);
		} else {
			// generic O(L*H^2) form, verbatim tie rules (:4453-4464)
			for (int i = 1; i <= len; ++i) {
				float* row = post + ((size_t)(i - 1) * H) * kBlock;
				const float* prow = post + ((size_t)(i - 2) * H) * kBlock;
				uint8_t* prow_path = path + ((size_t)(i - 1) * H) * kBlock;
				for (int j = 0; j < H; ++j) {
					float mx = -1.0f; int mv = -1;
					for (int c = 0; c <= j; ++c) {
						const float pv = (i >= 2) ? prow[(size_t)c * kBlock] : 0.0f;
						const float tmp = pv * (float)a.tmat[c * H + j];
						if (tmp > mx) { mv = c; mx = tmp; }
						if (tmp == mx && c == j) { mv = c; mx = tmp; }
					}
					row[(size_t)j * kBlock] = post_exp(row[(size_t)j * kBlock]) + mx;
					prow_path[(size_t)j * kBlock] = (uint8_t)mv;
				}
			}
			// final argmax, first max wins (:4494-4501)
			float mx = -1.0f;
			for (int j = 0; j < H; ++j) {
				const float v = (len >= 1) ? post[((size_t)(len - 1) * H + j) * kBlock] : 0.0f;
				if (v > mx) { mx = v; move = j; }
			}
			for (int i = 0; i <= rlen; ++i) lab_set(i, 0);
			if (move < 0) move = 0;
			lab_set(len, move);
			for (int i = len; i > 0; --i) {
				move = path[((size_t)(i - 1) * H + move) * kBlock];
				lab_set(i - 1, move);
			}
			lab_flush(rlen + 1);
		}
	}
	if (!a.do_extract) return;
	// ---- extract_reads (:3172-3313); make_extracted_read's rewrite is done on the host
	const SeqReader rd = make_reader(a, read);
	int key = 0, bar = -1, mem = -1, fingerlen = 0, s_pos = 0, hmm_has_barcode = 0, too_short = 0, in_read = 0;
	int read_type, barcode = -1, fingerprint = -1;
	const float mapq = a.mapq[read];
	if (a.confidence_threshold <= mapq) {
		for (int j = 0; j < len; ++j) {
			const int c1 = s_hlabel[lab_get(j + 1)];
			const int c2 = c1 & 0xFFFF;
			const int c3 = (c1 >> 16) & 0x7FFF;
			const uint8_t ty = s_stype[c2];
			if (ty == 'F') { fingerlen++; key = (key << 2) | (rd.code(j + off) & 0x3); }
			if (ty == 'B') {
				hmm_has_barcode = 1; bar = c3;
				if (bar == a.seg[c2].nh - 1) hmm_has_barcode = -1;
				mem = c2;
			}
			if (ty == 'R') { s_pos++; in_read = 1; }
			else {
				if (in_read && s_pos < a.minlen) { too_short = 1; break; }
				in_read = 0; s_pos = 0;
			}
		}
		if (in_read && s_pos < a.minlen) too_short = 1;
		const int rfl = a.required_finger_len;
		const int fp = (key << 8) | (rfl <= 255 ? rfl : 255);
		if (!too_short) {
			if (hmm_has_barcode == -1) read_type = 3;
			else if (hmm_has_barcode && rfl) {
				if (fingerlen == rfl && bar != -1) { barcode = (mem << 16) | bar; fingerprint = fp; read_type = 0; }
				else read_type = 3;
			} else if (hmm_has_barcode) {
				if (bar != -1) { barcode = (mem << 16) | bar; read_type = 0; }
				else read_type = 3;
			} else if (rfl) {
				if (fingerlen == rfl) { fingerprint = fp; read_type = 0; }
				else read_type = 3;
			} else read_type = 0;
		} else read_type = 2;
	} else read_type = 1;
	const bool extracted = (read_type == 0);
	// ---- dust_sequences (:2407-2467) on the sequence as make_extracted_read (:3325-3356)
	// leaves it: R-labelled residues keep their base, everything else is the spacer 65.
	if (a.dust) {
		auto e = [&](int j) -> int {
			if (j >= rlen) return 0;  // NUL terminator
			if (extracted) {
				const int c2 = s_hlabel[lab_get(j + 1)] & 0xFFFF;
				if (s_stype[c2] != 'R') return 65;
			}
			return rd.code(j);
		};
		uint8_t cnt[64];
		for (int j = 0; j < 64; ++j) cnt[j] = 0;
		int c = 0;
		while (e(c) == 65) c++;
		unsigned key = ((e(c) & 0x3) << 2) | (e(c + 1) & 0x3);
		int dl = rlen; if (dl > 64) dl = 64;
		c += 2;
		for (int j = c; j < dl; ++j) {
			const int v = e(j);
			if (v == 65) break;
			key = (key << 2) | (v & 0x3);
			cnt[key & 0x3F]++;
			c++;
		}
		int si = 0;
		for (int j = 0; j < 64; ++j) si += (int)cnt[j] * ((int)cnt[j] - 1) / 2;
		double s = (double)si;
		s = s / (double)(c - 3) * 10.0;
		if (s > (double)a.dust) read_type = 6;
	}
	// R-run spans: what make_extracted_read (:3325-3356) keeps of an extracted read -- residue j stays when the segment of
	// labels[j+1] is an R segment, everything else becomes the spacer.  Labels are non-decreasing in the HMM index (T is
	// upper triangular), so every R segment yields at most one run, plus one for a label-0 tail behind a -start/-end window.
	if (a.spans) {
		uint16_t* sp = a.spans + (size_t)read * a.span_stride * 2;
		int k = 0;
		if (extracted) {
			int j = 0;
			while (j < rlen && k < a.span_stride) {
				while (j < rlen && s_stype[s_hlabel[lab_get(j + 1)] & 0xFFFF] != 'R') j++;
				const int s0 = j;
				while (j < rlen && s_stype[s_hlabel[lab_get(j + 1)] & 0xFFFF] == 'R') j++;
				if (j > s0) { sp[2 * k] = (uint16_t)s0; sp[2 * k + 1] = (uint16_t)(j - s0); k++; }
			}
		}
		for (; k < a.span_stride; ++k) { sp[2 * k] = 0; sp[2 * k + 1] = 0; }
	}
	a.extracted[read] = extracted ? 1 : 0;
	a.read_type[read] = read_type;
	a.barcode[read] = barcode;
	a.fingerprint[read] = fingerprint;
}

// ------------------------------------------------------------------------------------------
// k_artifact: the -ref artifact filter, match_to_reference (barcode_hmm.c:2478-2583), then dust_sequences
// (:2407-2467) in the reference's order extract -> artifacts -> dust (:2345-2354).  Thread-per-read.
//
// The pattern is the read as make_extracted_read left it (spacer 65 outside the R-run spans), matched on both strands
// against every reference sequence with Myers' bit-vector algorithm in 64-bit words.  The reference has two variants and
// which one a read gets depends on its place in the thread slice of the run_pHMM call: reads in the groups of four at
// the start of a slice take the best match of validate_bpm_sse -> bmp_single (misc.c:718-765: first min(len, 63)
// pattern characters); the last (slice length mod 4) reads take the first reference sequence that bpm_check_error
// (misc.c:572-636: bits at index mod 64, score read at bit min(#bases, 31) - 1, score starting at len) puts within the
// cut-off.  Both are reproduced as they are.  Reference characters are staged through shared memory in tiles that all
// reads of the CTA walk together; both strands advance in the same loop (two independent dependency chains).
// ------------------------------------------------------------------------------------------
constexpr int kArtBlock = 128;
constexpr int kArtTile = 2048;

struct Myers {
	uint64_t VP, VN;
	__device__ __forceinline__ void step(uint64_t eq, uint64_t& HP, uint64_t& HN)
	{
		const uint64_t X = eq | VN;
		const uint64_t D0 = ((VP + (X & VP)) ^ VP) | X;
		HN = VP & D0;
		HP = VN | ~(VP | D0);
		const uint64_t Y = HP << 1;
		VN = Y & D0;
		VP = (HN << 1) | ~(Y | D0);
	}
};

__global__ void __launch_bounds__(kArtBlock) k_artifact(const KArgs a)
{
	__shared__ uint8_t s_t[kArtTile];
	const int read = blockIdx.x * kArtBlock + threadIdx.x;
	const bool valid = read < a.n_reads;
	const int rlen = valid ? a.len[read] : 0;
	int rt = 0;
	bool extracted = false;
	if (valid && !a.model_less) { rt = a.read_type[read]; extracted = a.extracted[read] != 0; }
	const SeqReader rd = make_reader(a, valid ? read : 0);
	const uint16_t* sp = (a.spans && valid) ? a.spans + (size_t)read * a.span_stride * 2 : nullptr;
	// residue j of the rewritten read
	auto e = [&](int j) -> int {
		if (j >= rlen) return 0;
		if (extracted) {
			bool in = false;
			for (int k = 0; k < a.span_stride; ++k) {
				const int s0 = sp[2 * k], sl = sp[2 * k + 1];
				if (sl == 0) break;
				if (j >= s0 && j < s0 + sl) { in = true; break; }
			}
			if (!in) return 65;
		}
		return rd.code(j);
	};
	auto rc_of = [](int c) -> int { return c == 65 ? 65 : (c < 4 ? 3 - c : 4); };   // rev_nuc_code, nuc_code.c:68-72

	bool active = valid && rt == 0 && a.ref_numseq > 0;   // only EXTRACT_SUCCESS reads can become artifacts (:2546, :2576)
	// place in the reference's thread slice (run_pHMM :1911-1922 / run_rna_dust :2063-2072)
	bool grouped = true;
	if (active) {
		const int gi = a.slice_base + read, T = a.slice_threads, iv = a.slice_interval;
		int t = iv > 0 ? gi / iv : T - 1;
		if (t > T - 1) t = T - 1;
		const int start = t * iv, end = (t == T - 1) ? a.slice_n : start + iv;
		grouped = (gi - start) < ((end - start) / 4) * 4;
	}
	// pattern bit-vectors, forward and reverse-complement strand
	uint64_t Bf[4] = {0, 0, 0, 0}, Br[4] = {0, 0, 0, 0};
	int m = 0, sh = 0;
	uint64_t mask = 0;
	if (active) {
		if (grouped) {
			m = rlen > 63 ? 63 : rlen;
			for (int i = 0; i < m; ++i) {
				const int c = e(i);
				if (c != 65) Bf[c & 3] |= (uint64_t)1 << i;
				const int r = rc_of(e(rlen - 1 - i));
				if (r != 65) Br[r & 3] |= (uint64_t)1 << i;
			}
			mask = m > 0 ? (uint64_t)1 << (m - 1) : 0;
		} else {
			int nf = 0;   // the number of bases is the same on both strands
			for (int i = 0; i < rlen; ++i) {
				const int c = e(i);
				if (c != 65) { Bf[c & 3] |= (uint64_t)1 << (i & 63); nf++; }
				const int r = rc_of(e(rlen - 1 - i));
				if (r != 65) Br[r & 3] |= (uint64_t)1 << (i & 63);
			}
			m = nf > 31 ? 31 : nf;
			sh = (m - 1) & 63;
			mask = (uint64_t)1 << sh;
		}
	}
	int best = 100000, best_id = 0, hit = 0;
	for (int j = 0; j < a.ref_numseq; ++j) {
		const int t0 = a.ref_index[j], n = a.ref_index[j + 1] - t0;
		Myers F, R;
		long long df, dr;   // running scores
		int kf, kr;
		if (grouped) { F.VP = R.VP = m > 0 ? (((uint64_t)1 << m) - 1) : 0; df = dr = m; kf = kr = m; }
		else { F.VP = R.VP = ~(uint64_t)0; df = dr = rlen; kf = kr = m; }
		F.VN = R.VN = 0;
		const bool run = active && !hit && (grouped ? rlen > 0 : true);
		for (int base = 0; base < n; base += kArtTile) {
			const int cnt = min(kArtTile, n - base);
			__syncthreads();
			for (int k = threadIdx.x; k < cnt; k += kArtBlock) s_t[k] = a.ref_codes[t0 + base + k] & 3;
			__syncthreads();
			if (!run) continue;
			for (int i = 0; i < cnt; ++i) {
				const int c = s_t[i];
				uint64_t HP, HN;
				F.step(Bf[c], HP, HN);
				if (grouped) { df += (HP & mask) ? 1 : 0; df -= (HN & mask) ? 1 : 0; if (df < kf) kf = (int)df; }
				else { df += (long long)((HP & mask) >> sh); df -= (long long)((HN & mask) >> sh); if ((unsigned long long)df < (unsigned long long)kf) kf = (int)df; }
				R.step(Br[c], HP, HN);
				if (grouped) { dr += (HP & mask) ? 1 : 0; dr -= (HN & mask) ? 1 : 0; if (dr < kr) kr = (int)dr; }
				else { dr += (long long)((HP & mask) >> sh); dr -= (long long)((HN & mask) >> sh); if ((unsigned long long)dr < (unsigned long long)kr) kr = (int)dr; }
			}
		}
		if (!active || hit) continue;
		if (grouped) {
			const int ef = rlen > 0 ? kf : n, er = rlen > 0 ? kr : n;   // validate_bpm_sse: an empty query scores n
			if (ef < best) { best = ef; best_id = j + 1; }
			if (er < best) { best = er; best_id = j + 1; }
		} else {
			if (kf <= a.filter_error || kr <= a.filter_error) hit = j + 1;   // forward is tried first, the id is the same
		}
	}
	if (!valid) return;
	if (active) {
		if (grouped && best <= a.filter_error) hit = best_id;
		if (hit) rt = (hit << 8) | 5;   // EXTRACT_FAIL_MATCHES_ARTIFACTS, io.h:36-52
	}
	if (a.dust_after) {   // dust_sequences (:2407-2467), same arithmetic as the tail of k_label
		uint8_t cnt[64];
		for (int j = 0; j < 64; ++j) cnt[j] = 0;
		int c = 0;
		while (e(c) == 65) c++;
		unsigned key = ((e(c) & 0x3) << 2) | (e(c + 1) & 0x3);
		int dl = rlen; if (dl > 64) dl = 64;
		c += 2;
		for (int j = c; j < dl; ++j) {
			const int v = e(j);
			if (v == 65) break;
			key = (key << 2) | (v & 0x3);
			cnt[key & 0x3F]++;
			c++;
		}
		int si = 0;
		for (int j = 0; j < 64; ++j) si += (int)cnt[j] * ((int)cnt[j] - 1) / 2;
		double s = (double)si;
		s = s / (double)(c - 3) * 10.0;
		if (s > (double)a.dust_after) rt = 6;
	}
	a.read_type[read] = rt;
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
size_t decode_smem_bytes(int model_floats, int dyn_cols)
{
	return (size_t)(kLogsumSize + model_floats) * sizeof(float) + (size_t)2 * dyn_cols * kBlock * sizeof(float);
}

// smem_bytes = the device's opt-in maximum per block; each kernel's own static shared memory
// is subtracted so that the dynamic limit requested is the largest the driver accepts.
template <class K>
static int set_dyn_smem(K kernel, int smem_bytes, int cap)
{
	cudaFuncAttributes fa;
	cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
	if (e != cudaSuccess) return (int)e;
	int dyn = smem_bytes - (int)fa.sharedSizeBytes;
	if (cap > 0 && dyn > cap) dyn = cap;
	return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
}

int kernels_configure(int smem_bytes)
{
	int e;
	if ((e = set_dyn_smem(k_backward<true, false>, smem_bytes, 0))) return e;
	if ((e = set_dyn_smem(k_backward<false, false>, smem_bytes, 0))) return e;
	if ((e = set_dyn_smem(k_backward<true, true>, smem_bytes, 0))) return e;
	if ((e = set_dyn_smem(k_backward<false, true>, smem_bytes, 0))) return e;
	if ((e = set_dyn_smem(k_forward<false>, smem_bytes, 0))) return e;
	if ((e = set_dyn_smem(k_forward<true>, smem_bytes, 0))) return e;
	if ((e = set_dyn_smem(k_label, smem_bytes, 110 * 1024))) return e;
	return 0;
}

int launch_backward(const KArgs& a, bool store, int ctas, void* stream)
{
	const size_t smem = decode_smem_bytes(a.model_in_smem ? a.model_floats : 0, a.dyn_cols);
	if (a.model_in_smem) {
		if (store) k_backward<true, false><<<ctas, kBlock, smem, (cudaStream_t)stream>>>(a);
		else k_backward<false, false><<<ctas, kBlock, smem, (cudaStream_t)stream>>>(a);
	} else {
		if (store) k_backward<true, true><<<ctas, kBlock, smem, (cudaStream_t)stream>>>(a);
		else k_backward<false, true><<<ctas, kBlock, smem, (cudaStream_t)stream>>>(a);
	}
	return (int)cudaGetLastError();
}

int launch_forward(const KArgs& a, int ctas, void* stream)
{
	const size_t smem = decode_smem_bytes(a.model_in_smem ? a.model_floats : 0, a.dyn_cols);
	if (a.model_in_smem) k_forward<false><<<ctas, kBlock, smem, (cudaStream_t)stream>>>(a);
	else k_forward<true><<<ctas, kBlock, smem, (cudaStream_t)stream>>>(a);
	return (int)cudaGetLastError();
}

int launch_label(const KArgs& a, int ctas_decode, void* stream)
{
	const int threads = ctas_decode * kBlock;
	int bs = kDpBlock;
	size_t smem = 0;
	if (a.dp_structured) {
		while (bs > 32 && (size_t)a.H * bs * 4 > 54 * 1024) bs >>= 1;  // keep >= 4 CTAs per SM
		smem = (size_t)a.H * bs * 4;
	}
	// labels staged in shared memory when they fit beside the DP row without costing a resident CTA
	KArgs b = a;
	const size_t lab_bytes = (size_t)((a.lmax + 1 + 3) / 4 * 4) * bs;
	b.label_smem = (b.want_labels && smem + lab_bytes <= 54 * 1024) ? 1 : 0;
	if (b.label_smem) smem += lab_bytes;
	const int ctas = (threads + bs - 1) / bs;
	k_label<<<ctas, bs, smem, (cudaStream_t)stream>>>(b);
	return (int)cudaGetLastError();
}

int launch_artifact(const KArgs& a, void* stream)
{
	const int ctas = (a.n_reads + kArtBlock - 1) / kArtBlock;
	if (ctas == 0) return 0;
	k_artifact<<<ctas, kArtBlock, 0, (cudaStream_t)stream>>>(a);
	return (int)cudaGetLastError();
}

}  // namespace tdg
