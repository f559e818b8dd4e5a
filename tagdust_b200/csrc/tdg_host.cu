// tdg_host.cu -- host runtime behind the C ABI of include/tagdust_b200.h.
//
// What it replaces in the reference: run_pHMM()'s thread fan-out and per-thread model copies
// (barcode_hmm.c:1895-2029, copy_model_bag :5262-5382).  Instead of pthreads over static
// slices it shards a batch contiguously over the GPUs of the context, stages reads in pinned
// 4-bit packed tiles, and queues H2D -> kernels -> D2H per device on streams.
#include <cuda_runtime.h>
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <thread>

#include <sys/mman.h>

#include "../../include/tagdust_b200.h"
#include "../../include/tagdust_b200_stream.h"
#include "tdg_device.h"
#include "tdg_pack.h"

using namespace tdg;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	g_err = buf;
	return code;
}
#define CK(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess)                                                                     \
			return fail(e_ == cudaErrorMemoryAllocation ? TDG_EMEM : TDG_ECUDA, "%s:%d %s -> %s", \
			            __FILE__, __LINE__, #call, cudaGetErrorString(e_));                        \
	} while (0)

extern "C" const char* tdg_last_error(void) { return g_err.c_str(); }
namespace tdg {
int set_last_error(int code, const char* msg) { g_err = msg ? msg : ""; return code; }  // used by tdg_stream.cpp
}
// tdg_model_set_max_len() may run on another host thread than tdg_submit(): both take this lock
static std::mutex g_model_mu;
extern "C" const char* tdg_version(void) { return "tagdust_b200 0.1 (sm_100a; reference TagDust 2.33)"; }

// ------------------------------------------------------------------------------------------
// host numerics identical to misc.c:57-105 (same libm, same expressions)
// ------------------------------------------------------------------------------------------
static float g_tab[kLogsumSize];
static std::once_flag g_tab_once;
static void init_tab()
{
	std::call_once(g_tab_once, [] {
		for (int i = 0; i < kLogsumSize; i++) g_tab[i] = (float)log(1. + exp((double)-i / 1000.0f));
	});
}
static inline float p2s(float p) { return p == 0.0 ? -HUGE_VALF : (float)log((double)p); }
extern "C" void tdg_logsum_table(float* out)
{
	init_tab();
	memcpy(out, g_tab, sizeof g_tab);
}
extern "C" float tdg_logsum_host(float a, float b)
{
	init_tab();
	const float mx = (a > b) ? a : b;
	const float mn = (a < b) ? a : b;
	return (mn == -HUGE_VAL || (mx - mn) >= 15.7f) ? mx : mx + g_tab[(int)((mx - mn) * 1000.0f)];
}

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct DeviceCtx {
	int dev = 0;
	int sms = 0;
	int ctas = 0;  // persistent decode CTAs = SM count
	cudaStream_t compute = nullptr;
	float* d_tab = nullptr;
	// scratch arena (grown on demand; kernels of one device are serialised on `compute`)
	void* scratch = nullptr;
	size_t scratch_bytes = 0;
	size_t smem_optin = 0;
	int configured_smem = 0;
	// experiment (TDG_LANES=2): waves alternate between `compute` and `lane2`, each with its own scratch half, so that
	// the HBM-bound k_backward of one wave shares the SMs with the issue-bound k_forward of another
	cudaStream_t lane2 = nullptr;
	cudaEvent_t lane_fork = nullptr, lane_join = nullptr;
	// optional per-kernel timing (tdg_profile_*): event pairs recorded around every launch
	bool profile = false;
	std::vector<cudaEvent_t> ev[3];  // [kernel kind] start/stop pairs
};

struct tdg_context {
	std::vector<DeviceCtx> devs;
	// staging batches handed back by the streaming layer: creating and, above all, freeing pinned staging costs more
	// than a whole chunk of work, so they are kept for the next job and only released by tdg_shutdown
	std::mutex pool_mu;
	std::vector<tdg_batch*> pool;
};

extern "C" int tdg_device_count(const tdg_context* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

extern "C" int tdg_init(int n_devices, const int* device_ids, tdg_context** out)
{
	if (!out) return fail(TDG_EINVAL, "tdg_init: out is NULL");
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		cudaGetLastError();
		return fail(TDG_ENODEV, "no CUDA device available (%s); tagdust_b200 has no CPU fallback",
		            e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
	}
	if (n_devices <= 0) n_devices = count;
	if (n_devices > count && !device_ids) return fail(TDG_EINVAL, "asked for %d devices, %d visible", n_devices, count);
	init_tab();
	auto* ctx = new tdg_context();
	// every device that got as far as push_back is released by tdg_shutdown on a later failure
	auto init_dev = [&](int k) -> int {
		DeviceCtx d;
		d.dev = device_ids ? device_ids[k] : k;
		cudaDeviceProp prop;
		CK(cudaGetDeviceProperties(&prop, d.dev));
		if (prop.major < 10)
			return fail(TDG_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", d.dev, prop.major, prop.minor);
		CK(cudaSetDevice(d.dev));
		d.sms = prop.multiProcessorCount;
		d.ctas = d.sms;
		d.smem_optin = prop.sharedMemPerBlockOptin;
		CK(cudaStreamCreateWithFlags(&d.compute, cudaStreamNonBlocking));
		ctx->devs.push_back(d);
		DeviceCtx& dd = ctx->devs.back();
		CK(cudaMalloc(&dd.d_tab, sizeof g_tab));
		std::vector<float> t(g_tab, g_tab + kLogsumSize);
		for (int i = 15700; i < kLogsumSize; i++) t[i] = 0.0f;  // see LS() in tdg_kernels.cu
		CK(cudaMemcpy(dd.d_tab, t.data(), sizeof g_tab, cudaMemcpyHostToDevice));
		return TDG_OK;
	};
	for (int k = 0; k < n_devices; k++) {
		const int rc = init_dev(k);
		if (rc != TDG_OK) {
			const std::string keep = g_err;
			tdg_shutdown(ctx);
			g_err = keep;
			return rc;
		}
	}
	*out = ctx;
	return TDG_OK;
}

extern "C" void tdg_shutdown(tdg_context* ctx)
{
	if (!ctx) return;
	for (tdg_batch* b : ctx->pool) tdg_batch_destroy(b);
	ctx->pool.clear();
	for (auto& d : ctx->devs) {
		cudaSetDevice(d.dev);
		cudaStreamSynchronize(d.compute);
		if (d.scratch) cudaFree(d.scratch);
		if (d.d_tab) cudaFree(d.d_tab);
		for (int k = 0; k < 3; k++) for (auto e : d.ev[k]) cudaEventDestroy(e);
		if (d.lane2) { cudaStreamSynchronize(d.lane2); cudaStreamDestroy(d.lane2); cudaEventDestroy(d.lane_fork); cudaEventDestroy(d.lane_join); }
		cudaStreamDestroy(d.compute);
	}
	delete ctx;
}

static void prof_mark(DeviceCtx& d, int kind, cudaStream_t st)
{
	if (!d.profile) return;
	cudaEvent_t e;
	if (cudaEventCreate(&e) != cudaSuccess) return;
	cudaEventRecord(e, st);
	d.ev[kind].push_back(e);
}

extern "C" int tdg_profile_enable(tdg_context* ctx, int on)
{
	if (!ctx) return fail(TDG_EINVAL, "NULL context");
	for (auto& d : ctx->devs) {
		cudaSetDevice(d.dev);
		for (int k = 0; k < 3; k++) {
			for (auto e : d.ev[k]) cudaEventDestroy(e);
			d.ev[k].clear();
		}
		d.profile = on != 0;
	}
	return TDG_OK;
}

// Sums the device time of every k_backward / k_forward / k_label launch recorded since
// tdg_profile_enable(ctx, 1) on device `devk`; the caller must have synchronised the stream.
extern "C" int tdg_profile_read(tdg_context* ctx, int devk, float ms[3], int launches[3])
{
	if (!ctx || devk < 0 || devk >= (int)ctx->devs.size()) return fail(TDG_EINVAL, "bad device index");
	DeviceCtx& d = ctx->devs[devk];
	CK(cudaSetDevice(d.dev));
	for (int k = 0; k < 3; k++) {
		ms[k] = 0.0f; launches[k] = (int)d.ev[k].size() / 2;
		for (size_t i = 0; i + 1 < d.ev[k].size(); i += 2) {
			float t = 0.0f;
			CK(cudaEventSynchronize(d.ev[k][i + 1]));
			CK(cudaEventElapsedTime(&t, d.ev[k][i], d.ev[k][i + 1]));
			ms[k] += t;
		}
	}
	return TDG_OK;
}

static int num_lanes()
{
	static const int v = [] { const char* e = getenv("TDG_LANES"); const int n = e ? atoi(e) : 1; return n == 2 ? 2 : 1; }();
	return v;
}

static int ensure_scratch(DeviceCtx& d, size_t bytes)
{
	if (bytes <= d.scratch_bytes) return TDG_OK;
	CK(cudaSetDevice(d.dev));
	CK(cudaStreamSynchronize(d.compute));
	if (d.scratch) CK(cudaFree(d.scratch));
	d.scratch = nullptr;
	d.scratch_bytes = 0;
	CK(cudaMalloc(&d.scratch, bytes));
	d.scratch_bytes = bytes;
	return TDG_OK;
}

// ------------------------------------------------------------------------------------------
// model
// ------------------------------------------------------------------------------------------
struct ModelDev {
	float* blob = nullptr;
	int32_t* dp_src = nullptr;
	int32_t* hmm_label = nullptr;
	uint8_t* seg_type = nullptr;
	uint8_t* tmat = nullptr;
};

struct HostModel {  // derived, GPU independent
	int S = 0, H = 0, C = 0, avg = 0;
	std::vector<SegInfo> seg;
	std::vector<float> blob;       // colrec C*12 then emit C*10
	std::vector<int32_t> dp_src;   // H * kMaxSources
	std::vector<int32_t> label;
	std::vector<uint8_t> seg_type;
	std::vector<uint8_t> tmat;
	int dp_structured = 0;
	int required_finger_len = 0;
	float bg[5];
	float r_step = 0, r_end = 0;
	int dead_terms = 0, std_segments = 0;
	int loop_cols = 0;   // longest segment that runs a column-loop kernel path (0 = every segment is unrolled)
	// logsums / float adds the kernels execute per read position over all HMMs (dead terms are never evaluated)
	double live_ls_bwd = 0, live_add_bwd = 0, live_ls_fwd = 0, live_add_fwd = 0;
};

static bool is_ninf(float v) { return std::isinf(v) && v < 0; }

static int derive_model(const tdg_model_desc* d, HostModel& hm, std::string& err)
{
	char buf[256];
	if (!d) { err = "desc is NULL"; return TDG_EINVAL; }
	const int S = d->num_segments, H = d->total_hmms, C = d->total_columns;
	if (S < 1 || S > kMaxSegments) { snprintf(buf, sizeof buf, "num_segments %d out of range 1..%d", S, kMaxSegments); err = buf; return TDG_EINVAL; }
	if (H < 1 || H > TDG_MAX_HMMS) { snprintf(buf, sizeof buf, "total_hmms %d out of range 1..%d", H, TDG_MAX_HMMS); err = buf; return TDG_EINVAL; }
	if (d->average_raw_length < 1) { err = "average_raw_length < 1"; return TDG_EINVAL; }
	hm.S = S; hm.H = H; hm.C = C; hm.avg = d->average_raw_length;
	hm.seg.resize(S);
	int cb = 0, hb = 0;
	for (int s = 0; s < S; s++) {
		SegInfo& g = hm.seg[s];
		g.nh = d->seg_num_hmms[s]; g.nc = d->seg_num_cols[s];
		if (g.nh < 1 || g.nc < 1 || g.nc > kMaxSegCols) { snprintf(buf, sizeof buf, "segment %d: %d HMMs x %d columns unsupported (columns 1..%d)", s, g.nh, g.nc, kMaxSegCols); err = buf; return TDG_EINVAL; }
		g.colbase = cb; g.hmmbase = hb; g.skip = d->seg_skip[s]; g.skip_live = !is_ninf(g.skip); g.kind = 0; g.use_win = 0;
		cb += g.nh * g.nc; hb += g.nh;
	}
	if (cb != C || hb != H) { err = "total_columns / total_hmms inconsistent with the segment tables"; return TDG_EINVAL; }
	hm.blob.assign((size_t)C * (kColRec + kEmitRec), 0.0f);
	float* rec = hm.blob.data();
	float* emit = rec + (size_t)C * kColRec;
	for (int c = 0; c < C; c++) {
		float* r = rec + (size_t)c * kColRec;
		for (int k = 0; k < 9; k++) r[k] = d->transition[c * 9 + k];
		r[F_SM] = d->silent_to_M[c];
		r[F_SI] = d->silent_to_I[c];
		uint32_t live = 0;
		for (int k = 0; k < 11; k++) {
			if (std::isnan(r[k]) || (std::isinf(r[k]) && r[k] > 0)) { err = "model contains NaN/+inf"; return TDG_EINVAL; }
			if (!is_ninf(r[k])) live |= 1u << k; else hm.dead_terms++;
		}
		memcpy(&r[F_LIVE], &live, 4);
		for (int k = 0; k < 5; k++) {
			emit[(size_t)c * kEmitRec + k] = d->m_emit[c * 5 + k];
			emit[(size_t)c * kEmitRec + 5 + k] = d->i_emit[c * 5 + k];
		}
	}
	// STDU check per segment (kernel code path 1): the standard liveness pattern, the five
	// transition scalars shared by every HMM of the segment, +0 for DM(n-2) and MSKIP(n-1), and
	// one insert-emission row for all columns.  Everything compared bit for bit.
	auto same = [](float x, float y) { return memcmp(&x, &y, 4) == 0; };
	for (int s = 0; s < S; s++) {
		SegInfo& g = hm.seg[s];
		g.ta = g.tb = g.tb2 = g.tc = g.td = 0.0f;
		if (g.nc < 3) continue;  // up to kMaxStdCols columns the kernels are fully unrolled, beyond that they loop over the columns
		if (getenv("TDG_NO_STDU")) continue;  // tests: every segment through the generic (run-time mask) paths
		const float* r0 = rec + (size_t)g.colbase * kColRec;
		const float* e0 = emit + (size_t)g.colbase * kEmitRec;
		const float ta = r0[F_MM], tb = r0[F_MI], tc = r0[F_II], td = r0[F_IM];
		const float tb2 = (r0 + (size_t)(g.nc - 2) * kColRec)[F_MI];
		bool ok = true;
		for (int f = 0; f < g.nh && ok; f++)
			for (int col = 0; col < g.nc && ok; col++) {
				const int c = g.colbase + f * g.nc + col;
				const float* r = rec + (size_t)c * kColRec;
				const float* e = emit + (size_t)c * kEmitRec;
				for (int k = 0; k < 11; k++)
					if (!std_live(g.nc, col, k) && !is_ninf(r[k])) ok = false;
				for (int k = 0; k < 5; k++)
					if (!same(e[5 + k], e0[5 + k])) ok = false;
				if (col == g.nc - 1) { if (!same(r[F_MSKIP], 0.0f)) ok = false; continue; }
				if (!same(r[F_MM], ta) || !same(r[F_II], tc) || !same(r[F_IM], td)) ok = false;
				if (!same(r[F_MI], col == g.nc - 2 ? tb2 : tb)) ok = false;
				if (col <= g.nc - 3 && !same(r[F_MD], tb)) ok = false;
				if (col >= 1 && col <= g.nc - 3 && (!same(r[F_DD], tc) || !same(r[F_DM], td))) ok = false;
				if (col == g.nc - 2 && !same(r[F_DM], 0.0f)) ok = false;
			}
		if (ok) { g.kind = 1; g.ta = ta; g.tb = tb; g.tb2 = tb2; g.tc = tc; g.td = td; hm.std_segments++; }
		g.use_win = (ok && g.nc >= 3 && g.nc <= kMaxStdCols && !getenv("TDG_NO_WIN")) ? 1 : 0;
	}
	for (int s = 0; s < S; s++) {
		const SegInfo& g = hm.seg[s];
		const bool unrolled = (g.kind == 1) ? (g.nc >= 3 && g.nc <= kMaxStdCols) : (g.nc <= 8);
		if (!unrolled) hm.loop_cols = std::max(hm.loop_cols, g.nc);
	}
	// Executed work per read position (the roofline's "live" op count): the same term-by-term walk as
	// bwd_segment / fwd_segment in tdg_kernels.cu, one logsum or add counted where the kernel issues one.
	for (int s = 0; s < S; s++) {
		const SegInfo& g = hm.seg[s];
		for (int f = 0; f < g.nh; f++) {
			const int c0 = g.colbase + f * g.nc, m = g.nc - 1;
			auto lv = [&](int col, int field) {
				uint32_t live;
				memcpy(&live, &rec[(size_t)(c0 + col) * kColRec + F_LIVE], 4);
				return (live >> field) & 1u;
			};
			double bl = 0, ba = 0, fl = 0, fa = 0;
			// backward, last column (:3518-3541)
			if (lv(m, F_MSKIP)) ba += 1;
			if (lv(m, F_ISKIP)) ba += 1;
			if (lv(m, F_IM)) { bl += 1; ba += 2; }
			if (lv(m, F_II)) { bl += 1; ba += 2; }
			if (lv(m, F_SM)) { bl += 1; ba += 2; }
			if (lv(m, F_SI)) { bl += 1; ba += 2; }
			for (int col = m - 1; col >= 0; col--) {  // (:3545-3589)
				if (lv(col, F_MM)) ba += 2;
				if (lv(col, F_MSKIP)) { bl += 1; ba += 1; }
				if (lv(col, F_MI)) { bl += 1; ba += 2; }
				if (lv(col, F_MD)) { bl += 1; ba += 1; }
				if (lv(col, F_II)) ba += 2;
				if (lv(col, F_ISKIP)) { bl += 1; ba += 1; }
				if (lv(col, F_IM)) { bl += 1; ba += 2; }
				if (lv(col, F_DD)) ba += 1;
				if (lv(col, F_DM)) { ba += 2; if (lv(col, F_DD)) bl += 1; }
				if (lv(col, F_SM)) { bl += 1; ba += 2; }
				if (lv(col, F_SI)) { bl += 1; ba += 2; }
			}
			if (g.skip_live) { bl += 1; ba += 1; }
			// forward + posterior, column 0 (:4218-4266)
			{
				bool have = false;
				if (lv(0, F_SM)) fa += 2;
				fa += 2; fl += 1;  // tM, TP
				if (lv(0, F_SI)) { fa += 1; have = true; }
				if (lv(0, F_II)) { fa += 1; if (have) fl += 1; have = true; }
				if (lv(0, F_MI)) { fa += 1; if (have) fl += 1; have = true; }
				fa += 1;  // + eI
				if (lv(0, F_SI)) { fl += 1; fa += 4; }
				fl += 1; fa += 2;  // P
				if (lv(0, F_MSKIP)) { fl += 1; fa += 1; }
				if (lv(0, F_ISKIP)) { fl += 1; fa += 1; }
			}
			for (int col = 1; col <= m; col++) {  // (:4270-4331)
				const int p = col - 1;
				bool have = false;
				if (lv(col, F_SM)) { fa += 1; have = true; }
				if (lv(p, F_MM)) { fa += 1; if (have) fl += 1; have = true; }
				if (lv(p, F_IM)) { fa += 1; if (have) fl += 1; have = true; }
				if (lv(p, F_DM)) { fa += 1; if (have) fl += 1; have = true; }
				fa += 1;
				if (have) { fl += 1; fa += 2; }
				have = false;
				if (lv(col, F_SI)) { fa += 1; have = true; }
				if (lv(col, F_II)) { fa += 1; if (have) fl += 1; have = true; }
				if (lv(col, F_MI)) { fa += 1; if (have) fl += 1; have = true; }
				fa += 1;
				if (have) { fl += 1; fa += 2; }
				bool dh = false;
				if (lv(p, F_MD)) { fa += 1; dh = true; }
				if (lv(p, F_DD)) { fa += 1; if (dh) fl += 1; }
				if (lv(col, F_MSKIP)) { fl += 1; fa += 1; }
				if (lv(col, F_ISKIP)) { fl += 1; fa += 1; }
			}
			if (g.skip_live) { fl += 1; fa += 1; }
			hm.live_ls_bwd += bl; hm.live_add_bwd += ba; hm.live_ls_fwd += fl; hm.live_add_fwd += fa;
		}
	}
	// labels, types
	hm.label.assign(d->label, d->label + H);
	hm.seg_type.assign((const uint8_t*)d->seg_type, (const uint8_t*)d->seg_type + S);
	for (int s = 0; s < S; s++)
		if (d->seg_type[s] == 'F') hm.required_finger_len += d->seg_num_cols[s];
	for (int h = 0; h < H; h++) {
		const int sg = d->label[h] & 0xFFFF;
		if (sg < 0 || sg >= S) { err = "label[] names a segment out of range"; return TDG_EINVAL; }
	}
	// label-DP transition matrix -> byte matrix + structured source lists
	hm.tmat.assign((size_t)H * H, 0);
	bool binary = true;
	for (int c = 0; c < H; c++)
		for (int j = 0; j < H; j++) {
			const float t = d->transition_matrix[c * H + j];
			if (t == 1.0f) hm.tmat[(size_t)c * H + j] = 1;
			else if (t != 0.0f) binary = false;
		}
	if (!binary) { err = "transition_matrix entries must be 0 or 1"; return TDG_EINVAL; }
	hm.dp_src.assign((size_t)H * kMaxSources, INT_MIN);
	hm.dp_structured = 1;
	for (int j = 0; j < H; j++) {
		if (!hm.tmat[(size_t)j * H + j]) hm.dp_structured = 0;  // reference always sets the diagonal
		int k = 0, c = 0;
		while (c < j) {
			if (!hm.tmat[(size_t)c * H + j]) { c++; continue; }
			// whole segment starting at c?
			int s = -1;
			for (int q = 0; q < S; q++) if (hm.seg[q].hmmbase == c) s = q;
			bool whole = false;
			if (s >= 0 && hm.seg[s].hmmbase + hm.seg[s].nh <= j) {
				whole = true;
				for (int f = 0; f < hm.seg[s].nh; f++) if (!hm.tmat[(size_t)(c + f) * H + j]) whole = false;
			}
			if (k >= kMaxSources) { hm.dp_structured = 0; break; }
			if (whole) { hm.dp_src[(size_t)j * kMaxSources + k++] = -(s + 1); c += hm.seg[s].nh; }
			else { hm.dp_src[(size_t)j * kMaxSources + k++] = c; c++; }
		}
	}
	if (getenv("TDG_GENERIC_LABEL_DP")) hm.dp_structured = 0;  // tests: the verbatim O(L*H^2) label DP
	for (int k = 0; k < 5; k++) hm.bg[k] = d->background[k];
	// random model constants (barcode_hmm.c:4520,4523): float argument, double log, float result
	hm.r_step = p2s((float)(1.0 - (1.0 / (float)d->average_raw_length)));
	hm.r_end = p2s((float)(1.0 / (float)d->average_raw_length));
	return TDG_OK;
}

struct tdg_model {
	tdg_context* ctx = nullptr;
	HostModel hm;
	int max_len = 0;
	std::vector<ModelDev> dev;
	size_t slot_bytes_full = 0, slot_bytes_bwd = 0;
	int dyn_cols = 0;  // shared-memory profile state for the column-loop paths, when it fits beside the table and the model
	int model_in_smem = 1;  // 0: the tables do not fit into shared memory beside the logsum table and are read from global memory
};

extern "C" int tdg_desc_live_ops(const tdg_model_desc* desc, double out[4])
{
	if (!out) return fail(TDG_EINVAL, "NULL argument");
	HostModel hm;
	std::string err;
	const int rc = derive_model(desc, hm, err);
	if (rc != TDG_OK) return fail(rc, "%s", err.c_str());
	out[0] = hm.live_ls_bwd; out[1] = hm.live_add_bwd; out[2] = hm.live_ls_fwd; out[3] = hm.live_add_fwd;
	return TDG_OK;
}

extern "C" int tdg_model_validate(const tdg_model_desc* desc, char* errbuf, size_t errbuf_len)
{
	HostModel hm;
	std::string err;
	const int rc = derive_model(desc, hm, err);
	if (errbuf && errbuf_len) {
		if (rc != TDG_OK) snprintf(errbuf, errbuf_len, "%s", err.c_str());
		else snprintf(errbuf, errbuf_len, "ok: S=%d H=%d C=%d dead_terms=%d std_segments=%d dp_structured=%d", hm.S, hm.H, hm.C,
		              hm.dead_terms, hm.std_segments, hm.dp_structured);
	}
	if (rc != TDG_OK) g_err = err;
	return rc;
}

template <class T>
static int upload(T** dst, const std::vector<T>& v)
{
	CK(cudaMalloc((void**)dst, std::max<size_t>(v.size(), 1) * sizeof(T)));
	if (!v.empty()) CK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
	return TDG_OK;
}

extern "C" void tdg_model_destroy(tdg_model* m)
{
	if (!m) return;
	for (size_t k = 0; k < m->dev.size(); k++) {
		cudaSetDevice(m->ctx->devs[k].dev);
		cudaStreamSynchronize(m->ctx->devs[k].compute);
		cudaFree(m->dev[k].blob); cudaFree(m->dev[k].dp_src); cudaFree(m->dev[k].hmm_label);
		cudaFree(m->dev[k].seg_type); cudaFree(m->dev[k].tmat);
	}
	delete m;
}

extern "C" int tdg_model_create(tdg_context* ctx, const tdg_model_desc* desc, int max_len, tdg_model** out)
{
	if (!ctx || !out) return fail(TDG_EINVAL, "tdg_model_create: NULL argument");
	*out = nullptr;
	if (max_len < 1) return fail(TDG_EINVAL, "max_len must be >= 1");
	if (max_len > 65000) return fail(TDG_EINVAL, "max_len %d too large (posterior windows are kept in 16 bits; limit 65000)", max_len);
	auto* m = new tdg_model();
	m->ctx = ctx;
	m->max_len = max_len;
	std::string err;
	int rc = derive_model(desc, m->hm, err);
	if (rc != TDG_OK) { delete m; return fail(rc, "%s", err.c_str()); }
	const HostModel& hm = m->hm;
	size_t smem = decode_smem_bytes((int)hm.blob.size(), 0);
	if (getenv("TDG_MODEL_IN_GLOBAL")) m->model_in_smem = 0;  // tests: force the global-memory tables
	const size_t W = (size_t)max_len + 2;
	m->slot_bytes_bwd = (size_t)hm.S * W * 4;
	m->slot_bytes_full = (size_t)hm.C * max_len * 8 + 3 * (size_t)hm.S * W * 4 + (size_t)max_len * hm.H * 4 + (size_t)hm.H * 8 +
	                     (size_t)max_len * hm.H;
	m->dev.resize(ctx->devs.size());
	for (size_t k = 0; k < ctx->devs.size(); k++) {
		DeviceCtx& d = ctx->devs[k];
		if (smem + 64 > d.smem_optin) {  // + the kernels' few bytes of static shared memory
			// too large to stage (more than ~1 880 columns): the kernels read the tables from global memory instead
			m->model_in_smem = 0;
			smem = decode_smem_bytes(0, 0);
		}
		if (k == 0 && hm.loop_cols > 0 && decode_smem_bytes(m->model_in_smem ? (int)hm.blob.size() : 0, hm.loop_cols) + 64 <= d.smem_optin && !getenv("TDG_NO_SMEM_STATE"))
			m->dyn_cols = hm.loop_cols;
		cudaSetDevice(d.dev);
		if ((int)smem > d.configured_smem) {
			// allow the maximum once so later (larger) models need no reconfiguration
			const int e = kernels_configure((int)d.smem_optin);
			if (e) { tdg_model_destroy(m); return fail(TDG_ECUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString((cudaError_t)e)); }
			d.configured_smem = (int)d.smem_optin;
		}
		ModelDev& md = m->dev[k];
		if ((rc = upload(&md.blob, hm.blob)) || (rc = upload(&md.dp_src, hm.dp_src)) || (rc = upload(&md.hmm_label, hm.label)) ||
		    (rc = upload(&md.seg_type, hm.seg_type)) || (rc = upload(&md.tmat, hm.tmat))) {
			tdg_model_destroy(m);
			return rc;
		}
	}
	*out = m;
	return TDG_OK;
}

// ------------------------------------------------------------------------------------------
// -ref artifact sequences
// ------------------------------------------------------------------------------------------
struct tdg_refset {
	tdg_context* ctx = nullptr;
	int numseq = 0;
	std::vector<uint8_t*> d_codes;
	std::vector<int32_t*> d_index;
};

extern "C" void tdg_refset_destroy(tdg_refset* r)
{
	if (!r) return;
	for (size_t k = 0; k < r->d_codes.size(); k++) {
		cudaSetDevice(r->ctx->devs[k].dev);
		cudaStreamSynchronize(r->ctx->devs[k].compute);
		cudaFree(r->d_codes[k]); cudaFree(r->d_index[k]);
	}
	delete r;
}

extern "C" int tdg_refset_create(tdg_context* ctx, const uint8_t* codes, const int32_t* s_index, int numseq, tdg_refset** out)
{
	if (!ctx || !out || !s_index || numseq < 0 || (numseq > 0 && !codes)) return fail(TDG_EINVAL, "tdg_refset_create: bad argument");
	*out = nullptr;
	for (int j = 0; j < numseq; j++)
		if (s_index[j + 1] < s_index[j]) return fail(TDG_EINVAL, "tdg_refset_create: s_index must be non-decreasing");
	auto* r = new tdg_refset();
	r->ctx = ctx; r->numseq = numseq;
	r->d_codes.assign(ctx->devs.size(), nullptr); r->d_index.assign(ctx->devs.size(), nullptr);
	std::vector<uint8_t> c(codes, codes + (numseq ? s_index[numseq] : 0));
	std::vector<int32_t> ix(s_index, s_index + numseq + 1);
	for (size_t k = 0; k < ctx->devs.size(); k++) {
		int rc;
		if (cudaSetDevice(ctx->devs[k].dev) != cudaSuccess) { tdg_refset_destroy(r); return fail(TDG_ECUDA, "cudaSetDevice failed"); }
		if ((rc = upload(&r->d_codes[k], c)) || (rc = upload(&r->d_index[k], ix))) { tdg_refset_destroy(r); return rc; }
	}
	*out = r;
	return TDG_OK;
}

// ------------------------------------------------------------------------------------------
// batches
// ------------------------------------------------------------------------------------------
struct Shard {  // one device's part of a batch
	int first = 0, n = 0;  // read range (first is a multiple of 32)
	cudaStream_t copy = nullptr;
	cudaEvent_t h2d_done = nullptr, k_done = nullptr, d2h_done = nullptr;
	int cap = 0;  // reads of device capacity
	void* slab = nullptr;  // one device allocation behind all the arrays below
	uint32_t* seq = nullptr; int32_t* len = nullptr;
	float *mapq = nullptr, *bar_prob = nullptr, *f = nullptr, *b = nullptr, *r = nullptr;
	int32_t *read_type = nullptr, *barcode = nullptr, *fingerprint = nullptr;
	uint8_t *extracted = nullptr, *labels = nullptr;
	uint16_t* spans = nullptr;
};

struct tdg_batch {
	tdg_context* ctx = nullptr;
	int max_reads = 0, max_len = 0, words = 0, label_stride = 0, n = 0;
	// pinned host staging: one allocation (pinning is the slow part of creating a batch), carved into the arrays below
	void* h_slab = nullptr;
	uint32_t* h_seq = nullptr; int32_t* h_len = nullptr;
	float *h_mapq = nullptr, *h_bar_prob = nullptr, *h_f = nullptr, *h_b = nullptr, *h_r = nullptr;
	int32_t *h_read_type = nullptr, *h_barcode = nullptr, *h_fingerprint = nullptr;
	uint8_t *h_extracted = nullptr, *h_labels = nullptr;
	uint16_t* h_spans = nullptr;
	int span_cap = 0;      // pairs per read the span buffers were allocated for
	int span_stride = 0;   // pairs per read of the last submit
	std::vector<Shard> shard;
	int pending_mode = 0; bool pending = false, want_labels = false, want_spans = false;
};

struct Carver {  // 256-byte aligned offsets into one slab
	size_t total = 0;
	size_t take(size_t bytes) { const size_t o = total; total += (std::max<size_t>(bytes, 1) + 255) / 256 * 256; return o; }
};

// Pinned host staging.  Large blocks are anonymous mappings with a huge-page hint, touched once and registered with
// the driver: pinning 2 MB pages runs at ~6 GB/s on this class of host, cudaHostAlloc's 4 KB pages at ~2 GB/s
// (scripts/micro/alloc_cost.cu, profiles/r02_alloc_cost2.txt) -- the start-up of a multi-GPU streaming job is made of
// exactly this.  Small blocks, and hosts where the mapping or the registration fails, use cudaHostAlloc.
struct PinnedBlock { void* p; size_t bytes; bool mapped; };
static std::mutex g_pin_mu;
static std::vector<PinnedBlock> g_pinned;   // mapped blocks only: how to release them

static int pinned_bytes(void** out, size_t bytes)
{
	bytes = std::max<size_t>(bytes, 1);
	static const bool no_thp = getenv("TDG_NO_THP") != nullptr;
	if (bytes >= ((size_t)4 << 20) && !no_thp) {
		const size_t huge = (size_t)2 << 20;
		const size_t al = (bytes + huge - 1) / huge * huge;
		void* q = mmap(nullptr, al, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
		if (q != MAP_FAILED) {
			madvise(q, al, MADV_HUGEPAGE);
			madvise(q, al, MADV_DONTFORK);   // the readers popen() zcat / bzcat: DMA-pinned pages stay out of the children
			memset(q, 0, al);   // first touch: the pages exist before the driver pins them
			if (cudaHostRegister(q, al, cudaHostRegisterPortable) == cudaSuccess) {
				std::lock_guard<std::mutex> l(g_pin_mu);
				g_pinned.push_back({q, al, true});
				*out = q;
				return TDG_OK;
			}
			cudaGetLastError();
			munmap(q, al);
		}
	}
	CK(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
	return TDG_OK;
}

static void pinned_free(void* p)
{
	if (!p) return;
	{
		std::lock_guard<std::mutex> l(g_pin_mu);
		for (size_t k = 0; k < g_pinned.size(); k++)
			if (g_pinned[k].p == p) {
				cudaHostUnregister(p);
				munmap(p, g_pinned[k].bytes);
				g_pinned.erase(g_pinned.begin() + (long)k);
				return;
			}
	}
	cudaFreeHost(p);
}

template <class T>
static int pinned(T** p, size_t n) { return pinned_bytes((void**)p, std::max<size_t>(n, 1) * sizeof(T)); }
template <class T>
static int devalloc(T** p, size_t n) { CK(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T))); return TDG_OK; }

extern "C" void tdg_batch_destroy(tdg_batch* b)
{
	if (!b) return;
	for (size_t k = 0; k < b->shard.size(); k++) {
		Shard& s = b->shard[k];
		cudaSetDevice(b->ctx->devs[k].dev);
		if (s.copy) cudaStreamSynchronize(s.copy);
		cudaStreamSynchronize(b->ctx->devs[k].compute);
		cudaFree(s.slab);
		cudaFree(s.labels);
		cudaFree(s.spans);
		if (s.h2d_done) cudaEventDestroy(s.h2d_done);
		if (s.k_done) cudaEventDestroy(s.k_done);
		if (s.d2h_done) cudaEventDestroy(s.d2h_done);
		if (s.copy) cudaStreamDestroy(s.copy);
	}
	pinned_free(b->h_slab);
	pinned_free(b->h_labels);
	pinned_free(b->h_spans);
	delete b;
}

extern "C" int tdg_batch_create(tdg_context* ctx, int max_reads, int max_len, tdg_batch** out)
{
	if (!ctx || !out) return fail(TDG_EINVAL, "tdg_batch_create: NULL argument");
	*out = nullptr;
	if (max_reads < 1 || max_len < 1) return fail(TDG_EINVAL, "max_reads and max_len must be >= 1");
	auto* b = new tdg_batch();
	b->ctx = ctx;
	b->max_reads = (max_reads + 31) / 32 * 32;
	b->max_len = max_len;
	b->words = (max_len + 1 + 7) / 8;
	b->label_stride = (max_len + 1 + 7) / 8 * 8;
	const size_t N = b->max_reads;
	int rc;
	{
		Carver cv;
		const size_t o_seq = cv.take(N * b->words * 4), o_len = cv.take(N * 4), o_mapq = cv.take(N * 4), o_bp = cv.take(N * 4),
		             o_f = cv.take(N * 4), o_b = cv.take(N * 4), o_r = cv.take(N * 4), o_rt = cv.take(N * 4), o_bc = cv.take(N * 4),
		             o_fp = cv.take(N * 4), o_ex = cv.take(N);
		char* base = nullptr;
		if ((rc = pinned(&base, cv.total))) { tdg_batch_destroy(b); return rc; }
		b->h_slab = base;
		b->h_seq = (uint32_t*)(base + o_seq); b->h_len = (int32_t*)(base + o_len); b->h_mapq = (float*)(base + o_mapq);
		b->h_bar_prob = (float*)(base + o_bp); b->h_f = (float*)(base + o_f); b->h_b = (float*)(base + o_b); b->h_r = (float*)(base + o_r);
		b->h_read_type = (int32_t*)(base + o_rt); b->h_barcode = (int32_t*)(base + o_bc); b->h_fingerprint = (int32_t*)(base + o_fp);
		b->h_extracted = (uint8_t*)(base + o_ex);
		// label buffers (the largest part: max_len + 1 bytes per read) are created by the first submit that asks for labels
	}
	memset(b->h_seq, 0, N * b->words * 4);
	const int nd = (int)ctx->devs.size();
	b->shard.resize(nd);
	const int tiles = (int)(N / 32);
	const int per = (tiles + nd - 1) / nd * 32;
	for (int k = 0; k < nd; k++) {
		Shard& s = b->shard[k];
		cudaSetDevice(ctx->devs[k].dev);
		s.cap = per;
		const size_t P = per;
		{
			Carver cv;
			const size_t o_seq = cv.take(P * b->words * 4), o_len = cv.take(P * 4), o_mapq = cv.take(P * 4), o_bp = cv.take(P * 4),
			             o_f = cv.take(P * 4), o_b = cv.take(P * 4), o_r = cv.take(P * 4), o_rt = cv.take(P * 4), o_bc = cv.take(P * 4),
			             o_fp = cv.take(P * 4), o_ex = cv.take(P);
			char* base = nullptr;
			if ((rc = devalloc(&base, cv.total))) { tdg_batch_destroy(b); return rc; }
			s.slab = base;
			s.seq = (uint32_t*)(base + o_seq); s.len = (int32_t*)(base + o_len); s.mapq = (float*)(base + o_mapq);
			s.bar_prob = (float*)(base + o_bp); s.f = (float*)(base + o_f); s.b = (float*)(base + o_b); s.r = (float*)(base + o_r);
			s.read_type = (int32_t*)(base + o_rt); s.barcode = (int32_t*)(base + o_bc); s.fingerprint = (int32_t*)(base + o_fp);
			s.extracted = (uint8_t*)(base + o_ex);
		}
		if (cudaStreamCreateWithFlags(&s.copy, cudaStreamNonBlocking) != cudaSuccess ||
		    cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreateWithFlags(&s.k_done, cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming) != cudaSuccess) {
			tdg_batch_destroy(b);
			return fail(TDG_ECUDA, "stream/event creation failed");
		}
	}
	*out = b;
	return TDG_OK;
}

extern "C" int tdg_batch_clear(tdg_batch* b)
{
	if (!b) return fail(TDG_EINVAL, "NULL batch");
	b->n = 0;
	return TDG_OK;
}
extern "C" int tdg_batch_size(const tdg_batch* b) { return b ? b->n : 0; }

static inline void pack_read(tdg_batch* b, int r, const uint8_t* codes, int len)
{
	// tile layout [tile][word][lane]; the code after the last base (the reference's NUL, or the
	// next base of a windowed read) is packed too: backward() reads it (barcode_hmm.c:3516)
	uint32_t* base = b->h_seq + ((size_t)(r >> 5) * b->words) * 32 + (r & 31);
	const int total = len + 1;
	int pos = 0;
	for (int w = 0; w < b->words; w++) {
		uint32_t v = 0;
		if (pos + 8 <= total) {
			// eight codes at once: keep the low nibble of every byte and fold the bytes together pairwise
			uint64_t x;
			memcpy(&x, codes + pos, 8);
			x &= 0x0F0F0F0F0F0F0F0FULL;
			x = (x | (x >> 4)) & 0x00FF00FF00FF00FFULL;
			x = (x | (x >> 8)) & 0x0000FFFF0000FFFFULL;
			x = (x | (x >> 16)) & 0x00000000FFFFFFFFULL;
			v = (uint32_t)x;
			pos += 8;
		} else {
			for (int k = 0; k < 8 && pos < total; k++, pos++) v |= (uint32_t)(codes[pos] & 0xF) << (4 * k);
		}
		base[(size_t)w * 32] = v;
	}
	b->h_len[r] = len;
}

extern "C" int tdg_batch_append_codes(tdg_batch* b, int n, const uint8_t* codes, size_t stride, const int32_t* len)
{
	if (!b || !codes || !len) return fail(TDG_EINVAL, "NULL argument");
	if (b->n + n > b->max_reads) return fail(TDG_EINVAL, "batch overflow: %d + %d > %d", b->n, n, b->max_reads);
	for (int i = 0; i < n; i++) {
		if (len[i] < 0 || len[i] > b->max_len) return fail(TDG_EINVAL, "read %d: length %d exceeds batch max_len %d", i, len[i], b->max_len);
		if ((size_t)len[i] + 1 > stride) return fail(TDG_EINVAL, "read %d: stride %zu too small for length %d + terminator", i, stride, len[i]);
	}
	// every read owns its own words of the tile layout: large appends are packed on a few host threads
	const int first = b->n;
	int T = n >= (1 << 17) ? 4 : 1;
	if (const char* e = getenv("TDG_PACK_THREADS")) T = std::max(1, atoi(e));
	T = std::min(T, std::max(1, n / 4096));
	auto work = [&](int lo, int hi) { for (int i = lo; i < hi; i++) pack_read(b, first + i, codes + (size_t)i * stride, len[i]); };
	if (T <= 1) work(0, n);
	else {
		std::vector<std::thread> th;
		const int per = (n + T - 1) / T;
		for (int t = 1; t < T; t++) th.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
		work(0, std::min(n, per));
		for (auto& x : th) x.join();
	}
	b->n += n;
	return TDG_OK;
}

extern "C" int tdg_batch_append_records(tdg_batch* b, int n, const void* const* records, size_t seq_off, size_t len_off)
{
	if (!b || !records) return fail(TDG_EINVAL, "NULL argument");
	if (b->n + n > b->max_reads) return fail(TDG_EINVAL, "batch overflow: %d + %d > %d", b->n, n, b->max_reads);
	for (int i = 0; i < n; i++) {
		const char* rec = (const char*)records[i];
		const uint8_t* seq = *(const uint8_t* const*)(rec + seq_off);
		const int len = *(const int*)(rec + len_off);
		if (len < 0 || len > b->max_len) return fail(TDG_EINVAL, "read %d: length %d exceeds batch max_len %d", i, len, b->max_len);
		pack_read(b, b->n + i, seq, len);
	}
	b->n += n;
	return TDG_OK;
}

// Ragged rows (one 0-terminated code string per read, as the FASTQ reader leaves them), packed on
// `threads` host threads: reads are independent and a tile's 32 lanes are written by one thread.
extern "C" int tdg_batch_append_ragged(tdg_batch* b, int n, const uint8_t* codes, const uint64_t* seq_off, const int32_t* len, int threads)
{
	if (!b || !codes || !seq_off || !len) return fail(TDG_EINVAL, "NULL argument");
	if (n < 0 || b->n + n > b->max_reads) return fail(TDG_EINVAL, "batch overflow: %d + %d > %d", b->n, n, b->max_reads);
	for (int i = 0; i < n; i++)
		if (len[i] < 0 || len[i] > b->max_len) return fail(TDG_EINVAL, "read %d: length %d exceeds batch max_len %d", i, len[i], b->max_len);
	const int first = b->n;
	// every read owns its own 32-bit words of the tile layout, so any split is race-free
	const int T = std::max(1, std::min(threads, n / 4096));
	auto work = [&](int lo, int hi) { for (int i = lo; i < hi; i++) pack_read(b, first + i, codes + seq_off[i], len[i]); };
	if (T <= 1) work(0, n);
	else {
		std::vector<std::thread> th;
		const int per = (n + T - 1) / T;
		for (int t = 1; t < T; t++) th.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
		work(0, std::min(n, per));
		for (auto& x : th) x.join();
	}
	b->n += n;
	return TDG_OK;
}

extern "C" int tdg_model_max_len(const tdg_model* m) { return m ? m->max_len : 0; }
extern "C" int tdg_model_num_hmms(const tdg_model* m) { return m ? m->hm.H : 0; }
extern "C" int tdg_model_set_max_len(tdg_model* m, int max_len)
{
	if (!m || max_len < 1 || max_len > 65000) return fail(TDG_EINVAL, "bad argument (max_len 1..65000)");
	std::lock_guard<std::mutex> model_lock(g_model_mu);
	if (max_len <= m->max_len) return TDG_OK;
	const HostModel& hm = m->hm;
	const size_t W = (size_t)max_len + 2;
	m->max_len = max_len;
	m->slot_bytes_bwd = (size_t)hm.S * W * 4;
	m->slot_bytes_full = (size_t)hm.C * max_len * 8 + 3 * (size_t)hm.S * W * 4 + (size_t)max_len * hm.H * 4 + (size_t)hm.H * 8 +
	                     (size_t)max_len * hm.H;
	return TDG_OK;
}
extern "C" int tdg_model_read_hmms(const tdg_model* m, uint8_t* is_read)
{
	if (!m || !is_read) return fail(TDG_EINVAL, "NULL argument");
	for (int h = 0; h < m->hm.H; h++) is_read[h] = m->hm.seg_type[m->hm.label[h] & 0xFFFF] == 'R';
	return TDG_OK;
}

extern "C" int tdg_plan_shards(int n_reads, int n_devices, int32_t* first, int32_t* count)
{
	if (n_reads < 0 || n_devices < 1 || !first || !count) return fail(TDG_EINVAL, "bad argument");
	const int tiles = (n_reads + 31) / 32;
	const int per = (tiles + n_devices - 1) / n_devices * 32;
	int f = 0;
	for (int k = 0; k < n_devices; k++) {
		first[k] = std::min(f, n_reads);
		count[k] = std::max(0, std::min(per, n_reads - first[k]));
		f += per;
	}
	return TDG_OK;
}

static void assign_shards(tdg_batch* b)
{
	const int nd = (int)b->shard.size();
	std::vector<int32_t> first(nd), count(nd);
	tdg_plan_shards(b->n, nd, first.data(), count.data());
	for (int k = 0; k < nd; k++) { b->shard[k].first = first[k]; b->shard[k].n = count[k]; }
}

// ------------------------------------------------------------------------------------------
// kernel queueing
// ------------------------------------------------------------------------------------------
static void fill_model_args(KArgs& a, const tdg_model* m, int devk, const DeviceCtx& d)
{
	const HostModel& hm = m->hm;
	memset(&a, 0, sizeof a);
	a.S = hm.S; a.H = hm.H; a.C = hm.C;
	for (int s = 0; s < hm.S; s++) a.seg[s] = hm.seg[s];
	a.model_blob = m->dev[devk].blob;
	a.model_floats = (int)hm.blob.size();
	a.model_in_smem = m->model_in_smem;
	a.dyn_cols = m->dyn_cols;
	a.logsum_tab = d.d_tab;
	a.r_step = hm.r_step; a.r_end = hm.r_end;
	for (int k = 0; k < 5; k++) a.bg[k] = hm.bg[k];
	a.lmax = m->max_len;
	a.dp_src = m->dev[devk].dp_src;
	a.hmm_label = m->dev[devk].hmm_label;
	a.seg_type = m->dev[devk].seg_type;
	a.tmat = m->dev[devk].tmat;
	a.dp_structured = hm.dp_structured;
	a.post_store_all = 1;
	a.required_finger_len = hm.required_finger_len;
}

// CTAs (of kBlock reads) per wave: one per SM, fewer when the per-read scratch of a long-read model
// (threshold calibration emits reads several times the average length) would not fit in HBM.
static int plan_wave_ctas(const tdg_model* m, const DeviceCtx& d, bool full, int n_reads)
{
	// the arena already holds what a full wave (or this whole shard) needs: no cudaMemGetInfo (it takes milliseconds per
	// device and waits behind concurrent allocations) on the per-chunk path
	{
		const long need = std::max(1, (n_reads + kBlock - 1) / kBlock);
		long want = std::min<long>(d.ctas, need);
		if (const char* e = getenv("TDG_WAVE_CTAS")) { const long cap = atol(e); if (cap > 0) want = std::min(want, cap); }
		const size_t bytes = (size_t)num_lanes() * ((size_t)want * kBlock * (full ? m->slot_bytes_full : m->slot_bytes_bwd) + 10 * 256);
		if (d.scratch && bytes <= d.scratch_bytes) return (int)want;
	}
	size_t free_b = 0, total_b = 0;
	if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return d.ctas; }
	const double budget = 0.90 * (double)(free_b + d.scratch_bytes);
	const double per_cta = (double)num_lanes() * (double)kBlock * (double)(full ? m->slot_bytes_full : m->slot_bytes_bwd);
	long fit = (long)(budget / per_cta);
	if (const char* e = getenv("TDG_WAVE_CTAS")) { const long cap = atol(e); if (cap > 0) fit = std::min(fit, cap); }  // tests: force small waves
	if (fit < 1) fit = 1;
	// a shard smaller than a wave needs scratch for its own CTAs only
	const long need = std::max(1, (n_reads + kBlock - 1) / kBlock);
	return (int)std::min<long>(std::min<long>(d.ctas, fit), need);
}

static size_t lane_scratch_bytes(const tdg_model* m, bool full, int wave_ctas)
{
	const size_t slots = (size_t)wave_ctas * kBlock;
	return slots * (full ? m->slot_bytes_full : m->slot_bytes_bwd) + 10 * 256;
}

static void carve_scratch(KArgs& a, const tdg_model* m, const DeviceCtx& d, bool full, int wave_ctas, int lane = 0)
{
	const HostModel& hm = m->hm;
	const size_t slots = (size_t)wave_ctas * kBlock;
	const size_t W = (size_t)m->max_len + 2;
	char* p = (char*)d.scratch + (size_t)lane * lane_scratch_bytes(m, full, wave_ctas);
	auto take = [&](size_t bytes) { char* q = p; p += (bytes + 255) / 256 * 256; return q; };
	a.sb = (float*)take(slots * hm.S * W * 4);
	if (full) {
		a.sf = (float*)take(slots * hm.S * W * 4);
		a.mbmax = (float*)take(slots * hm.S * W * 4);
		a.tp = (float*)take(slots * hm.H * 4);
		a.prange = (uint32_t*)take(slots * hm.H * 4);
		a.post = (float*)take(slots * (size_t)m->max_len * hm.H * 4);
		a.path = (uint8_t*)take(slots * (size_t)m->max_len * hm.H);
		a.bw = (float2*)take(slots * (size_t)hm.C * m->max_len * 8);
	}
}

static size_t scratch_need(const tdg_model* m, bool full, int wave_ctas)
{
	return (size_t)num_lanes() * lane_scratch_bytes(m, full, wave_ctas);
}

// Queue all waves of one shard on `stream`.  Returns kernel launches queued (<0 on error).
static int queue_decode(tdg_context* ctx, tdg_model* m, int mode, const tdg_run_params* p, tdg_batch* b, int devk,
                        cudaStream_t stream, float* b_score_override, int wave_ctas)
{
	DeviceCtx& d = ctx->devs[devk];
	Shard& s = b->shard[devk];
	if (s.n == 0) return 0;
	const bool bwd_only = (mode == TDG_MODE_ARCH_COMP);
	const bool want_labels = (mode == TDG_MODE_GET_LABEL) || (mode == TDG_MODE_GET_PROB && p && p->want_labels);  // run the label DP
	KArgs a;
	fill_model_args(a, m, devk, d);
	a.store_labels = (p && p->want_labels) ? 1 : 0;
	const tdg_refset* rs = (mode == TDG_MODE_GET_LABEL && p) ? p->refset : nullptr;
	const bool want_spans = (mode == TDG_MODE_GET_LABEL && p && (p->want_spans || rs) && s.spans);
	a.span_stride = want_spans ? b->span_stride : 0;
	carve_scratch(a, m, d, !bwd_only, wave_ctas);
	a.words = b->words;
	a.win_start = 0; a.win_len = -1;
	if (p && (p->matchstart != -1 || p->matchend != -1)) { a.win_start = p->matchstart; a.win_len = p->matchend - p->matchstart; }
	a.label_stride = b->label_stride;
	a.confidence_threshold = p ? p->confidence_threshold : 0.0f;
	a.minlen = p ? p->minlen : 0;
	a.dust = (p && mode == TDG_MODE_GET_LABEL) ? p->dust : 0;
	if (rs) {  // extract -> artifacts -> dust (barcode_hmm.c:2345-2354): dust moves behind the artifact match, into k_artifact
		a.dust_after = a.dust; a.dust = 0;
		a.ref_codes = rs->d_codes[devk]; a.ref_index = rs->d_index[devk]; a.ref_numseq = rs->numseq;
		a.filter_error = p->filter_error;
		a.slice_n = b->n; a.slice_threads = std::max(1, p->slice_threads);
		a.slice_interval = (int)((double)b->n / (double)a.slice_threads);
	}
	a.do_extract = (mode == TDG_MODE_GET_LABEL);
	a.want_labels = want_labels;
	const int wave = wave_ctas * kBlock;
	int launches = 0;
	const int lanes = num_lanes();
	cudaStream_t const main_stream = stream;
	if (lanes > 1) {
		if (!d.lane2) {
			if (cudaStreamCreateWithFlags(&d.lane2, cudaStreamNonBlocking) != cudaSuccess ||
			    cudaEventCreateWithFlags(&d.lane_fork, cudaEventDisableTiming) != cudaSuccess ||
			    cudaEventCreateWithFlags(&d.lane_join, cudaEventDisableTiming) != cudaSuccess) { fail(TDG_ECUDA, "lane stream creation failed"); return -1; }
		}
		cudaEventRecord(d.lane_fork, main_stream);
		cudaStreamWaitEvent(d.lane2, d.lane_fork, 0);
	}
	for (int w0 = 0, wi = 0; w0 < s.n; w0 += wave, wi++) {
		const int nw = std::min(wave, s.n - w0);
		if (lanes > 1) {
			stream = (wi % lanes) ? d.lane2 : main_stream;
			carve_scratch(a, m, d, !bwd_only, wave_ctas, wi % lanes);
		}
		a.n_reads = nw;
		a.seq = s.seq + (size_t)(w0 / 32) * b->words * 32;
		a.len = s.len + w0;
		a.b_score = (b_score_override ? b_score_override : s.b) + w0;
		a.f_score = s.f + w0; a.r_score = s.r + w0; a.bar_prob = s.bar_prob + w0; a.mapq = s.mapq + w0;
		a.read_type = s.read_type + w0; a.barcode = s.barcode + w0; a.fingerprint = s.fingerprint + w0;
		a.extracted = s.extracted + w0;
		a.labels = s.labels + (size_t)w0 * b->label_stride;
		a.spans = want_spans ? s.spans + (size_t)w0 * b->span_stride * 2 : nullptr;
		const int ctas = (nw + kBlock - 1) / kBlock;
		int e;
		prof_mark(d, 0, stream);
		if ((e = launch_backward(a, !bwd_only, ctas, stream))) { fail(TDG_ECUDA, "k_backward launch: %s", cudaGetErrorString((cudaError_t)e)); return -1; }
		prof_mark(d, 0, stream);
		launches++;
		if (!bwd_only) {
			prof_mark(d, 1, stream);
			if ((e = launch_forward(a, ctas, stream))) { fail(TDG_ECUDA, "k_forward launch: %s", cudaGetErrorString((cudaError_t)e)); return -1; }
			prof_mark(d, 1, stream);
			launches++;
			if (want_labels || a.do_extract) {
				prof_mark(d, 2, stream);
				if ((e = launch_label(a, ctas, stream))) { fail(TDG_ECUDA, "k_label launch: %s", cudaGetErrorString((cudaError_t)e)); return -1; }
				prof_mark(d, 2, stream);
				launches++;
			}
			if (rs) {
				a.slice_base = s.first + w0;
				if ((e = launch_artifact(a, stream))) { fail(TDG_ECUDA, "k_artifact launch: %s", cudaGetErrorString((cudaError_t)e)); return -1; }
				launches++;
			}
		}
	}
	if (lanes > 1) {
		cudaEventRecord(d.lane_join, d.lane2);
		cudaStreamWaitEvent(main_stream, d.lane_join, 0);
	}
	return launches;
}

// Device label rows: working storage of the label DP whenever it runs (the kernel stages them in shared memory when they fit,
// but that is decided per launch).  Pinned host rows: only when the caller wants the rows back -- pinning them is the slow part.
static int ensure_label_buffers(tdg_batch* b, bool host_rows)
{
	int rc;
	for (size_t k = 0; k < b->shard.size(); k++) {
		if (b->shard[k].labels) continue;
		CK(cudaSetDevice(b->ctx->devs[k].dev));
		if ((rc = devalloc(&b->shard[k].labels, (size_t)b->shard[k].cap * b->label_stride))) return rc;
	}
	if (host_rows && !b->h_labels && (rc = pinned(&b->h_labels, (size_t)b->max_reads * b->label_stride))) return rc;
	return TDG_OK;
}

static int span_stride_of(const tdg_model* m)
{
	int n = 1;
	for (int s = 0; s < m->hm.S; s++) n += m->hm.seg_type[s] == 'R';
	return n;
}

static int ensure_span_buffers(tdg_batch* b, int stride)
{
	b->span_stride = stride;
	if (b->span_cap >= stride) return TDG_OK;
	int rc;
	for (size_t k = 0; k < b->shard.size(); k++) {
		CK(cudaSetDevice(b->ctx->devs[k].dev));
		if (b->shard[k].spans) { CK(cudaStreamSynchronize(b->shard[k].copy)); CK(cudaFree(b->shard[k].spans)); b->shard[k].spans = nullptr; }
		if ((rc = devalloc(&b->shard[k].spans, (size_t)b->shard[k].cap * stride * 2))) return rc;
	}
	if (b->h_spans) { pinned_free(b->h_spans); b->h_spans = nullptr; }
	if ((rc = pinned(&b->h_spans, (size_t)b->max_reads * stride * 2))) return rc;
	b->span_cap = stride;
	return TDG_OK;
}

extern "C" int tdg_batch_reserve_labels(tdg_batch* b)
{
	if (!b) return fail(TDG_EINVAL, "NULL batch");
	return ensure_label_buffers(b, true);
}

static int check_compat(tdg_model* m, tdg_batch* b, const tdg_run_params* p)
{
	if (!m || !b) return fail(TDG_EINVAL, "NULL model or batch");
	if (m->ctx != b->ctx) return fail(TDG_EINVAL, "model and batch belong to different contexts");
	// k_label clears / stages labels over the whole read even when only a -start/-end window is decoded,
	// so the model must be sized for the longest read in either case
	int need = 0;
	for (int i = 0; i < b->n; i++) need = std::max(need, b->h_len[i]);
	if (p && (p->matchstart != -1 || p->matchend != -1)) {
		if (p->matchstart < 0 || p->matchend <= p->matchstart) return fail(TDG_EINVAL, "bad -start/-end window %d..%d", p->matchstart, p->matchend);
		for (int i = 0; i < b->n; i++)
			if (b->h_len[i] < p->matchend) return fail(TDG_EINVAL, "read %d (length %d) shorter than the -end window %d", i, b->h_len[i], p->matchend);
	}
	if (need > m->max_len) return fail(TDG_EINVAL, "read length %d exceeds the model's max_len %d (rebuild the model, cf. barcode_hmm.c:292-310)", need, m->max_len);
	return TDG_OK;
}

static int upload_shard(tdg_batch* b, int k, cudaStream_t st)
{
	Shard& s = b->shard[k];
	if (s.n == 0) return TDG_OK;
	const size_t tiles = (s.n + 31) / 32;
	CK(cudaMemcpyAsync(s.seq, b->h_seq + (size_t)(s.first / 32) * b->words * 32, tiles * b->words * 32 * 4, cudaMemcpyHostToDevice, st));
	CK(cudaMemcpyAsync(s.len, b->h_len + s.first, (size_t)s.n * 4, cudaMemcpyHostToDevice, st));
	return TDG_OK;
}

static int download_shard(tdg_batch* b, int k, cudaStream_t st, int mode, bool want_labels, bool want_spans = false)
{
	Shard& s = b->shard[k];
	if (s.n == 0) return TDG_OK;
	const size_t n = s.n, f = s.first;
	CK(cudaMemcpyAsync(b->h_b + f, s.b, n * 4, cudaMemcpyDeviceToHost, st));
	if (mode == TDG_MODE_ARCH_COMP) return TDG_OK;
	CK(cudaMemcpyAsync(b->h_mapq + f, s.mapq, n * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(b->h_bar_prob + f, s.bar_prob, n * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(b->h_f + f, s.f, n * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(b->h_r + f, s.r, n * 4, cudaMemcpyDeviceToHost, st));
	if (want_labels && b->h_labels) CK(cudaMemcpyAsync(b->h_labels + f * b->label_stride, s.labels, n * b->label_stride, cudaMemcpyDeviceToHost, st));
	if (mode == TDG_MODE_GET_LABEL && want_spans && b->h_spans)
		CK(cudaMemcpyAsync(b->h_spans + f * b->span_stride * 2, s.spans, n * b->span_stride * 2 * sizeof(uint16_t), cudaMemcpyDeviceToHost, st));
	if (mode == TDG_MODE_GET_LABEL) {
		CK(cudaMemcpyAsync(b->h_read_type + f, s.read_type, n * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(b->h_barcode + f, s.barcode, n * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(b->h_fingerprint + f, s.fingerprint, n * 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(b->h_extracted + f, s.extracted, n, cudaMemcpyDeviceToHost, st));
	}
	return TDG_OK;
}

static void fill_result(tdg_batch* b, tdg_result* out)
{
	if (!out) return;
	out->n_reads = b->n; out->label_stride = b->label_stride;
	out->mapq = b->h_mapq; out->bar_prob = b->h_bar_prob; out->f_score = b->h_f; out->b_score = b->h_b; out->r_score = b->h_r;
	out->read_type = b->h_read_type; out->extracted = b->h_extracted; out->barcode = b->h_barcode;
	out->fingerprint = b->h_fingerprint;
	out->labels = b->want_labels ? b->h_labels : nullptr;
	out->span_stride = b->want_spans ? b->span_stride : 0;
	out->spans = b->want_spans ? b->h_spans : nullptr;
}

// run_rna_dust (barcode_hmm.c:2043) for a batch of reads without a model: artifact filter + dust in one kernel
static int submit_rna_dust(tdg_context* ctx, const tdg_run_params* p, tdg_batch* b)
{
	if (!p || !b) return fail(TDG_EINVAL, "run params and batch required");
	if (b->ctx != ctx) return fail(TDG_EINVAL, "batch belongs to another context");
	if (p->refset && p->refset->ctx != ctx) return fail(TDG_EINVAL, "refset belongs to another context");
	if (b->pending) return fail(TDG_EINVAL, "batch already submitted; call tdg_wait first");
	assign_shards(b);
	for (size_t k = 0; k < ctx->devs.size(); k++) {
		DeviceCtx& d = ctx->devs[k];
		Shard& s = b->shard[k];
		if (s.n == 0) continue;
		CK(cudaSetDevice(d.dev));
		int rc;
		if ((rc = upload_shard(b, (int)k, s.copy))) return rc;
		CK(cudaEventRecord(s.h2d_done, s.copy));
		CK(cudaStreamWaitEvent(d.compute, s.h2d_done, 0));
		KArgs a;
		memset(&a, 0, sizeof a);
		a.n_reads = s.n; a.words = b->words; a.seq = s.seq; a.len = s.len; a.read_type = s.read_type;
		a.model_less = 1; a.dust_after = p->dust;
		if (p->refset) {
			a.ref_codes = p->refset->d_codes[k]; a.ref_index = p->refset->d_index[k]; a.ref_numseq = p->refset->numseq;
			a.filter_error = p->filter_error;
		}
		a.slice_n = b->n; a.slice_threads = std::max(1, p->slice_threads);
		a.slice_interval = (int)((double)b->n / (double)a.slice_threads);
		a.slice_base = s.first;
		const int e = launch_artifact(a, d.compute);
		if (e) return fail(TDG_ECUDA, "k_artifact launch: %s", cudaGetErrorString((cudaError_t)e));
		CK(cudaEventRecord(s.k_done, d.compute));
		CK(cudaStreamWaitEvent(s.copy, s.k_done, 0));
		CK(cudaMemcpyAsync(b->h_read_type + s.first, s.read_type, (size_t)s.n * 4, cudaMemcpyDeviceToHost, s.copy));
		CK(cudaEventRecord(s.d2h_done, s.copy));
	}
	b->pending = true; b->pending_mode = TDG_MODE_RNA_DUST; b->want_labels = false; b->want_spans = false;
	return TDG_OK;
}

extern "C" int tdg_submit(tdg_context* ctx, tdg_model* m, int mode, const tdg_run_params* p, tdg_batch* b)
{
	if (!ctx) return fail(TDG_EINVAL, "NULL context");
	if (mode == TDG_MODE_RNA_DUST) return submit_rna_dust(ctx, p, b);
	if (mode != TDG_MODE_GET_LABEL && mode != TDG_MODE_GET_PROB && mode != TDG_MODE_ARCH_COMP)
		return fail(TDG_EINVAL, "unsupported mode %d (MODE_TRAIN has no live caller in the reference)", mode);
	if (mode != TDG_MODE_ARCH_COMP && !p) return fail(TDG_EINVAL, "run params required");
	if (p && p->refset && p->refset->ctx != ctx) return fail(TDG_EINVAL, "refset belongs to another context");
	std::lock_guard<std::mutex> model_lock(g_model_mu);
	int rc = check_compat(m, b, p);
	if (rc) return rc;
	if (b->pending) return fail(TDG_EINVAL, "batch already submitted; call tdg_wait first");
	assign_shards(b);
	const bool run_dp = (mode == TDG_MODE_GET_LABEL) || (mode == TDG_MODE_GET_PROB && p->want_labels);
	const bool want_labels = mode != TDG_MODE_ARCH_COMP && p->want_labels;   // label rows come back to the host
	const bool want_spans = mode == TDG_MODE_GET_LABEL && p->want_spans;
	if (run_dp && (rc = ensure_label_buffers(b, want_labels))) return rc;
	if ((want_spans || (mode == TDG_MODE_GET_LABEL && p->refset)) && (rc = ensure_span_buffers(b, span_stride_of(m)))) return rc;
	for (size_t k = 0; k < ctx->devs.size(); k++) {
		DeviceCtx& d = ctx->devs[k];
		Shard& s = b->shard[k];
		if (s.n == 0) continue;
		CK(cudaSetDevice(d.dev));
		const int wc = plan_wave_ctas(m, d, mode != TDG_MODE_ARCH_COMP, b->shard[k].n);
		if ((rc = ensure_scratch(d, scratch_need(m, mode != TDG_MODE_ARCH_COMP, wc)))) return rc;
		if ((rc = upload_shard(b, (int)k, s.copy))) return rc;
		CK(cudaEventRecord(s.h2d_done, s.copy));
		CK(cudaStreamWaitEvent(d.compute, s.h2d_done, 0));
		if (queue_decode(ctx, m, mode, p, b, (int)k, d.compute, nullptr, wc) < 0) return TDG_ECUDA;
		CK(cudaEventRecord(s.k_done, d.compute));
		CK(cudaStreamWaitEvent(s.copy, s.k_done, 0));
		if ((rc = download_shard(b, (int)k, s.copy, mode, want_labels, want_spans))) return rc;
		CK(cudaEventRecord(s.d2h_done, s.copy));
	}
	b->pending = true; b->pending_mode = mode; b->want_labels = want_labels; b->want_spans = want_spans;
	return TDG_OK;
}

extern "C" int tdg_wait(tdg_batch* b, tdg_result* out)
{
	if (!b) return fail(TDG_EINVAL, "NULL batch");
	if (!b->pending) return fail(TDG_EINVAL, "nothing submitted on this batch");
	for (size_t k = 0; k < b->shard.size(); k++) {
		if (b->shard[k].n == 0) continue;
		CK(cudaSetDevice(b->ctx->devs[k].dev));
		CK(cudaEventSynchronize(b->shard[k].d2h_done));
	}
	b->pending = false;
	fill_result(b, out);
	return TDG_OK;
}

extern "C" int tdg_run(tdg_context* ctx, tdg_model* m, int mode, const tdg_run_params* p, tdg_batch* b, tdg_result* out)
{
	int rc = tdg_submit(ctx, m, mode, p, b);
	if (rc) return rc;
	return tdg_wait(b, out);
}

// MODE_ARCH_COMP: run_pHMM :1924-1938 + do_arch_comparison :2111-2148 + merge :1994-2017
extern "C" int tdg_arch_compare(tdg_context* ctx, tdg_model* const* models, int num_arch, tdg_batch* b, int num_threads,
                                float* b_scores, float* arch_posterior)
{
	if (!ctx || !models || !b || !arch_posterior || num_arch < 1) return fail(TDG_EINVAL, "bad argument");
	if (num_threads < 1) num_threads = 1;
	const int n = b->n;
	std::vector<float> all((size_t)num_arch * n);
	{
		// One upload of the reads, the backward kernels of all architectures queued back to back on the
		// device's compute stream (they share the scratch arena), one download of the A x n scores.
		std::lock_guard<std::mutex> model_lock(g_model_mu);
		if (b->pending) return fail(TDG_EINVAL, "batch already submitted; call tdg_wait first");
		int rc;
		for (int a = 0; a < num_arch; a++)
			if ((rc = check_compat(models[a], b, nullptr))) return rc;
		assign_shards(b);
		const size_t nd = ctx->devs.size();
		std::vector<float*> d_scores(nd, nullptr);
		auto body = [&]() -> int {  // every early return leaves through the cleanup below
			for (size_t k = 0; k < nd; k++) {
				DeviceCtx& d = ctx->devs[k];
				Shard& s = b->shard[k];
				if (s.n == 0) continue;
				CK(cudaSetDevice(d.dev));
				if ((rc = devalloc(&d_scores[k], (size_t)num_arch * s.n))) return rc;
				if ((rc = upload_shard(b, (int)k, s.copy))) return rc;
				CK(cudaEventRecord(s.h2d_done, s.copy));
				CK(cudaStreamWaitEvent(d.compute, s.h2d_done, 0));
				for (int a = 0; a < num_arch; a++) {
					const int wc = plan_wave_ctas(models[a], d, false, s.n);
					if ((rc = ensure_scratch(d, scratch_need(models[a], false, wc)))) return rc;
					if (queue_decode(ctx, models[a], TDG_MODE_ARCH_COMP, nullptr, b, (int)k, d.compute, d_scores[k] + (size_t)a * s.n, wc) < 0) return TDG_ECUDA;
				}
			}
			for (size_t k = 0; k < nd; k++) {
				Shard& s = b->shard[k];
				if (s.n == 0) continue;
				CK(cudaSetDevice(ctx->devs[k].dev));
				CK(cudaStreamSynchronize(ctx->devs[k].compute));
				for (int a = 0; a < num_arch; a++)
					CK(cudaMemcpy(all.data() + (size_t)a * n + s.first, d_scores[k] + (size_t)a * s.n, (size_t)s.n * 4, cudaMemcpyDeviceToHost));
			}
			return TDG_OK;
		};
		rc = body();
		for (size_t k = 0; k < nd; k++)
			if (d_scores[k]) { cudaSetDevice(ctx->devs[k].dev); cudaStreamSynchronize(ctx->devs[k].compute); cudaFree(d_scores[k]); }
		if (rc) return rc;
	}
	if (b_scores) memcpy(b_scores, all.data(), all.size() * 4);
	// per-"thread" float sums in read order over the reference's static slices, added in thread order
	const int interval = (int)((double)n / (double)num_threads);
	for (int a = 0; a < num_arch; a++) {
		float total = 0.0f;  // ab->arch_posterior[a] = prob2scaledprob(1.0)
		for (int t = 0; t < num_threads; t++) {
			const int s = t * interval, e = (t == num_threads - 1) ? n : t * interval + interval;
			float part = 0.0f;
			for (int i = s; i < e; i++) part += all[(size_t)a * n + i];
			total += part;
		}
		arch_posterior[a] = total;
	}
	float sum = arch_posterior[0];
	for (int a = 1; a < num_arch; a++) sum = tdg_logsum_host(sum, arch_posterior[a]);
	for (int a = 0; a < num_arch; a++) arch_posterior[a] = arch_posterior[a] - sum;
	return TDG_OK;
}

// ---- device-resident path ------------------------------------------------------------------
extern "C" int tdg_batch_upload(tdg_context* ctx, tdg_batch* b)
{
	if (!ctx || !b) return fail(TDG_EINVAL, "NULL argument");
	assign_shards(b);
	for (size_t k = 0; k < ctx->devs.size(); k++) {
		if (b->shard[k].n == 0) continue;
		CK(cudaSetDevice(ctx->devs[k].dev));
		int rc = upload_shard(b, (int)k, b->shard[k].copy);
		if (rc) return rc;
		CK(cudaStreamSynchronize(b->shard[k].copy));
	}
	return TDG_OK;
}

extern "C" int tdg_decode_resident(tdg_context* ctx, tdg_model* m, int mode, const tdg_run_params* p, tdg_batch* b,
                                   void* cuda_stream, int* n_launches)
{
	if (!ctx) return fail(TDG_EINVAL, "NULL context");
	std::lock_guard<std::mutex> model_lock(g_model_mu);
	int rc = check_compat(m, b, p);
	if (rc) return rc;
	if (mode != TDG_MODE_ARCH_COMP && (rc = ensure_label_buffers(b, true))) return rc;
	if (mode == TDG_MODE_GET_LABEL && p && p->want_spans && (rc = ensure_span_buffers(b, span_stride_of(m)))) return rc;
	b->want_labels = true; b->want_spans = (mode == TDG_MODE_GET_LABEL && p && p->want_spans);
	int total = 0;
	for (size_t k = 0; k < ctx->devs.size(); k++) {
		DeviceCtx& d = ctx->devs[k];
		if (b->shard[k].n == 0) continue;
		CK(cudaSetDevice(d.dev));
		const int wc = plan_wave_ctas(m, d, mode != TDG_MODE_ARCH_COMP, b->shard[k].n);
		if ((rc = ensure_scratch(d, scratch_need(m, mode != TDG_MODE_ARCH_COMP, wc)))) return rc;
		// a caller-provided stream is only meaningful for a single-device context
		cudaStream_t st = (ctx->devs.size() == 1) ? (cudaStream_t)cuda_stream : d.compute;
		const int l = queue_decode(ctx, m, mode, p, b, (int)k, st, nullptr, wc);
		if (l < 0) return TDG_ECUDA;
		total += l;
	}
	if (n_launches) *n_launches = total;
	return TDG_OK;
}

extern "C" int tdg_batch_download(tdg_batch* b, tdg_result* out)
{
	if (!b) return fail(TDG_EINVAL, "NULL batch");
	for (size_t k = 0; k < b->shard.size(); k++) {
		if (b->shard[k].n == 0) continue;
		CK(cudaSetDevice(b->ctx->devs[k].dev));
		CK(cudaDeviceSynchronize());
		int rc = download_shard(b, (int)k, b->shard[k].copy, TDG_MODE_GET_LABEL, true, b->want_spans);
		if (rc) return rc;
		CK(cudaStreamSynchronize(b->shard[k].copy));
	}
	fill_result(b, out);
	return TDG_OK;
}

extern "C" double tdg_batch_cells(const tdg_model* m, const tdg_batch* b)
{
	if (!m || !b) return 0.0;
	double cells = 0.0;
	for (int i = 0; i < b->n; i++) cells += 2.0 * (double)b->h_len[i] * (double)m->hm.C;
	return cells;
}

// ---- internal helpers of the streaming layer (tdg_stream.cpp) ---------------------------------------
namespace tdg {

// A pooled batch that is large enough (and not more than twice too large), or a new one.
int batch_acquire(tdg_context* ctx, int max_reads, int max_len, tdg_batch** out)
{
	{
		std::lock_guard<std::mutex> l(ctx->pool_mu);
		for (size_t k = 0; k < ctx->pool.size(); k++) {
			tdg_batch* b = ctx->pool[k];
			if (b->max_reads >= max_reads && b->max_len >= max_len && b->max_reads <= 2 * ((max_reads + 31) / 32 * 32) && b->max_len <= 2 * max_len + 16) {
				ctx->pool.erase(ctx->pool.begin() + (long)k);
				b->n = 0; b->pending = false;
				*out = b;
				return TDG_OK;
			}
		}
	}
	return tdg_batch_create(ctx, max_reads, max_len, out);
}

void batch_release(tdg_batch* b)
{
	if (!b) return;
	if (b->pending) { tdg_result r; tdg_wait(b, &r); }
	tdg_batch* victim = nullptr;
	{
		std::lock_guard<std::mutex> l(b->ctx->pool_mu);
		b->ctx->pool.push_back(b);
		if (b->ctx->pool.size() > 12) { victim = b->ctx->pool.front(); b->ctx->pool.erase(b->ctx->pool.begin()); }  // oldest goes
	}
	if (victim) tdg_batch_destroy(victim);
}

// The streaming layer packs reads straight from the characters of a FASTQ text into the pinned staging arrays
// (tdg_pack.h: pack_text_read, on its own worker pool, in the same pass that measures the lines): this hands out the
// arrays for the next n reads of the batch, batch_text_commit makes them part of it.
int batch_text_target(tdg_batch* b, int n, TextTarget* t)
{
	if (!b || !t) return fail(TDG_EINVAL, "NULL argument");
	if (n < 0 || b->n + n > b->max_reads) return fail(TDG_EINVAL, "batch overflow: %d + %d > %d", b->n, n, b->max_reads);
	t->seq = b->h_seq; t->len = b->h_len; t->words = b->words; t->max_len = b->max_len; t->first = b->n;
	return TDG_OK;
}

int batch_text_commit(tdg_batch* b, int n)
{
	if (!b || n < 0 || b->n + n > b->max_reads) return fail(TDG_EINVAL, "batch_text_commit: bad count");
	b->n += n;
	return TDG_OK;
}

// Everything the first tdg_submit(MODE_GET_LABEL, want_spans) of `b` under `m` would otherwise allocate on the GPU thread
// while kernels of earlier chunks are running (cudaMalloc waits for them): device label rows, span buffers.
int batch_prepare(tdg_batch* b, const tdg_model* m, bool host_labels)
{
	int rc = ensure_label_buffers(b, host_labels);
	if (!rc) rc = ensure_span_buffers(b, span_stride_of(m));
	return rc;
}

// The scratch arena of every device for full waves of `m`, allocated on one host thread per device.
int scratch_prepare(tdg_context* ctx, tdg_model* m)
{
	std::vector<std::thread> th;
	std::vector<int> rcs(ctx->devs.size(), TDG_OK);
	std::lock_guard<std::mutex> model_lock(g_model_mu);
	for (size_t k = 0; k < ctx->devs.size(); k++)
		th.emplace_back([&, k] {
			DeviceCtx& d = ctx->devs[k];
			if (cudaSetDevice(d.dev) != cudaSuccess) { rcs[k] = TDG_ECUDA; return; }
			const int wc = plan_wave_ctas(m, d, true, d.ctas * kBlock);
			rcs[k] = ensure_scratch(d, scratch_need(m, true, wc));
		});
	for (auto& t : th) t.join();
	for (int rc : rcs) if (rc) return rc;
	return TDG_OK;
}

}  // namespace tdg
