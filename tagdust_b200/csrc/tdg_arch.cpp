// tdg_arch.cpp -- architecture compiler: segment strings -> flattened model.
//
// Host-side mirror of the reference's model construction so the library can be driven
// without the reference's structs (bench, tests, new hosts).  It reproduces, expression by
// expression and with the same float/double typing,
//   assign_segment_sequences                 interface.c:489-598
//   init_model_bag                           barcode_hmm.c:5760-6011
//   init_model_according_to_read_structure   barcode_hmm.c:4689-5084
//   set_hmm_transition_parameters            barcode_hmm.c:1710-1881
//   gaussian_pdf                             misc.c:375-379
//   the calibration edit                     calibrateQ.c:67-86
// tests/test_capi_host.py (test_arch_compile_*) compares its output bit-for-bit with the reference's own
// init_model_bag (oracle/_ref) over every segment type.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tagdust_b200.h"

namespace {

inline float p2s(float p) { if (p == 0.0) return -HUGE_VAL; return (float)log((double)p); }   // misc.c:85-92
inline float s2p(float p) { if (p == -HUGE_VAL) return 0.0f; return (float)exp((double)p); }  // misc.c:98-105
const double kInvSqrt2Pi = 0.3989422804014327;                                                // misc.h
inline double gaussian_pdf(double x, double m, double s)
{
	double a = (x - m) / s;
	return kInvSqrt2Pi / s * exp(-0.5 * a * a);
}

struct Column {
	float t[9];
	float m_emit[5];
	float i_emit[5];
};
struct Hmm { std::vector<Column> col; };
struct Segment {
	char type = 0;
	std::vector<std::string> seqs;
	std::vector<Hmm> hmm;
	std::vector<std::vector<float>> sM, sI;  // [hmm][col]
	float skip = 0;
	float bg[5];
};

int nuc_code(unsigned char c)  // nuc_code.c:46-74
{
	switch (c) {
		case '.': return 5;
		case 'A': case 'a': return 0;
		case 'C': case 'c': return 1;
		case 'G': case 'g': return 2;
		case 'T': case 't': case 'U': case 'u': return 3;
		default: return 4;
	}
}

// barcode_hmm.c:1710-1881
void set_hmm_transition_parameters(Hmm& hmm, int len, double base_error, double indel_freq, double mean, double stdev)
{
	const float NI = p2s(0.0f);
	double sum_prob = 0.0;
	if (mean > 0.0 && stdev > 0.0)
		for (int i = 0; i <= len; i++) sum_prob += gaussian_pdf(i, mean, stdev);
	auto mskip = [&](double x) -> float {
		if (mean == -1.0 && stdev == -1.0) return p2s(0.0);
		else if (mean > -1.0 && stdev == -1.0) return p2s(mean / (float)(len - 1));
		else return p2s(gaussian_pdf(x, mean, stdev) / sum_prob);
	};
	auto rest = [&](const Column& c) -> float { return p2s(1.0 - s2p(c.t[TDG_MSKIP])); };
	if (len == 1) {
		Column& c = hmm.col[0];
		c.t[TDG_MM] = NI; c.t[TDG_MI] = NI; c.t[TDG_MD] = NI; c.t[TDG_MSKIP] = p2s(1.0f);
		c.t[TDG_II] = NI; c.t[TDG_IM] = NI; c.t[TDG_ISKIP] = NI; c.t[TDG_DD] = NI; c.t[TDG_DM] = NI;
	} else if (len == 2) {
		Column& c = hmm.col[0];
		c.t[TDG_MSKIP] = mskip(0);
		c.t[TDG_MM] = p2s(1.0 - base_error * indel_freq) + rest(c);
		c.t[TDG_MI] = p2s(base_error * indel_freq) + rest(c);
		c.t[TDG_MD] = p2s(base_error * indel_freq * 0.0) + rest(c);
		c.t[TDG_II] = p2s(1.0 - 0.999); c.t[TDG_IM] = p2s(0.999); c.t[TDG_ISKIP] = NI;
		c.t[TDG_DD] = NI; c.t[TDG_DM] = NI;
		Column& l = hmm.col[1];
		l.t[TDG_MM] = NI; l.t[TDG_MI] = NI; l.t[TDG_MD] = NI; l.t[TDG_MSKIP] = p2s(1.0);
		l.t[TDG_II] = NI; l.t[TDG_IM] = NI; l.t[TDG_ISKIP] = NI; l.t[TDG_DD] = NI; l.t[TDG_DM] = NI;
	} else {
		{
			Column& c = hmm.col[0];
			c.t[TDG_MSKIP] = mskip(0);
			c.t[TDG_MM] = p2s(1.0 - base_error * indel_freq) + rest(c);
			c.t[TDG_MI] = p2s(base_error * indel_freq * 0.5) + rest(c);
			c.t[TDG_MD] = p2s(base_error * indel_freq * 0.5) + rest(c);
			c.t[TDG_II] = p2s(1.0 - 0.999); c.t[TDG_IM] = p2s(0.999); c.t[TDG_ISKIP] = NI;
			c.t[TDG_DD] = NI; c.t[TDG_DM] = NI;
		}
		for (int i = 1; i < len - 2; i++) {
			Column& c = hmm.col[i];
			c.t[TDG_MSKIP] = mskip(i);
			c.t[TDG_MM] = p2s(1.0 - base_error * indel_freq) + rest(c);
			c.t[TDG_MI] = p2s(base_error * indel_freq * 0.5) + rest(c);
			c.t[TDG_MD] = p2s(base_error * indel_freq * 0.5) + rest(c);
			c.t[TDG_II] = p2s(1.0 - 0.999); c.t[TDG_IM] = p2s(0.999); c.t[TDG_ISKIP] = NI;
			c.t[TDG_DD] = p2s(1.0 - 0.999); c.t[TDG_DM] = p2s(0.999);
		}
		{
			Column& c = hmm.col[len - 2];
			c.t[TDG_MSKIP] = mskip(len - 1.0);
			c.t[TDG_MM] = p2s(1.0 - base_error * indel_freq) + rest(c);
			c.t[TDG_MI] = p2s(base_error * indel_freq * 1.0) + rest(c);
			c.t[TDG_MD] = p2s(base_error * indel_freq * 0.0) + rest(c);
			c.t[TDG_II] = p2s(1.0 - 0.999); c.t[TDG_IM] = p2s(0.999); c.t[TDG_ISKIP] = NI;
			c.t[TDG_DD] = p2s(0.0); c.t[TDG_DM] = p2s(1.0);
		}
		Column& l = hmm.col[len - 1];
		l.t[TDG_MM] = NI; l.t[TDG_MI] = NI; l.t[TDG_MD] = NI; l.t[TDG_MSKIP] = p2s(1.0);
		l.t[TDG_II] = NI; l.t[TDG_IM] = NI; l.t[TDG_ISKIP] = NI; l.t[TDG_DD] = NI; l.t[TDG_DM] = NI;
	}
}

// barcode_hmm.c:4689-5084
void init_segment(Segment& sg, float base_error, float indel_freq, const double* background, int assumed_length)
{
	const float NI = p2s(0.0f);
	const int nh = (int)sg.hmm.size();
	for (int i = 0; i < 5; i++) sg.bg[i] = (float)background[i];
	for (int i = 0; i < nh; i++) {
		const int len = (int)sg.hmm[i].col.size();
		const std::string& tmp = sg.seqs[i];
		for (int j = 0; j < len; j++) {
			Column& col = sg.hmm[i].col[j];
			int current_nuc = nuc_code((unsigned char)tmp[j]);
			if (current_nuc < 4) {
				for (int c = 0; c < 4; c++) {
					if (c == current_nuc) col.m_emit[c] = p2s(1.0 - s2p(background[4]) - base_error * (1.0 - indel_freq));
					else col.m_emit[c] = p2s(base_error * (1.0 - indel_freq) / 3.0);
					col.i_emit[c] = background[c];
				}
				col.m_emit[4] = background[4];
				col.i_emit[4] = background[4];
			} else if (current_nuc == 4) {
				for (int c = 0; c < 5; c++) { col.m_emit[c] = background[c]; col.i_emit[c] = background[c]; }
			} else {
				current_nuc = 4;
				for (int c = 0; c < 5; c++) {
					col.m_emit[c] = (c == current_nuc) ? p2s(1.0) : p2s(0.0);
					col.i_emit[c] = background[c];
				}
			}
		}
		set_hmm_transition_parameters(sg.hmm[i], len, base_error, indel_freq, -1.0, -1.0);
	}
	for (int i = 0; i < nh; i++) {
		sg.sM[i].assign(sg.hmm[i].col.size(), NI);
		sg.sI[i].assign(sg.hmm[i].col.size(), NI);
	}
	sg.skip = NI;
	const int len0 = (int)sg.hmm[0].col.size();
	if (sg.type == 'B' || sg.type == 'S') {
		for (int i = 0; i < nh; i++) { sg.sM[i][0] = p2s(1.0 / (float)nh); sg.sI[i][0] = p2s(0.0f); }
		sg.skip = p2s(0.0);
	}
	if (sg.type == 'F') {
		for (int i = 0; i < nh; i++) sg.sM[i][0] = p2s(1.0 / (float)nh);
		sg.skip = p2s(0.0);
	}
	if (sg.type == 'P') {
		for (int i = 0; i < nh; i++) {
			sg.sM[i][0] = p2s(1.0 / (float)nh) + p2s(1.0 - 0.01);
			for (int j = 0; j < len0; j++) {
				Column& col = sg.hmm[i].col[j];
				// here base_error and indel_freq are floats: the product is a float product
				col.t[TDG_MM] = p2s(1.0 - base_error * indel_freq) + p2s(0.99f);
				col.t[TDG_MI] = p2s(base_error * indel_freq) + p2s(0.5) + p2s(0.99f);
				col.t[TDG_MD] = p2s(base_error * indel_freq) + p2s(0.5) + p2s(0.99f);
				col.t[TDG_MSKIP] = p2s(0.01f);
				col.t[TDG_II] = p2s(1.0 - 0.999) + p2s(0.99f);
				col.t[TDG_IM] = p2s(0.999) + p2s(0.99f);
				col.t[TDG_ISKIP] = p2s(0.01f);
			}
		}
		sg.skip = p2s(0.01);
	}
	if (sg.type == 'O' || sg.type == 'G') {
		for (int i = 0; i < nh; i++) {
			if (sg.type == 'O') sg.sI[i][0] = p2s(1.0 / (float)nh) + p2s(0.5);
			else sg.sI[i][0] = p2s(0.8935878);
			for (int j = 0; j < len0; j++) {
				Column& col = sg.hmm[i].col[j];
				for (int c = 0; c < 5; c++) { col.i_emit[c] = col.m_emit[c]; col.m_emit[c] = p2s(0.0); }
			}
		}
		Column& col = sg.hmm[0].col[0];
		if (sg.type == 'O') {
			sg.skip = p2s(0.5);
			col.t[TDG_MM] = NI; col.t[TDG_MI] = NI; col.t[TDG_MD] = NI; col.t[TDG_MSKIP] = NI;
			col.t[TDG_II] = p2s(1.0 - 1.0 / (float)(len0 + 1));
			col.t[TDG_IM] = NI;
			col.t[TDG_ISKIP] = p2s(1.0 / (float)(len0 + 1));
			col.t[TDG_DD] = NI; col.t[TDG_DM] = NI;
		} else {
			sg.skip = p2s(1.0 - 0.8935878);
			col.t[TDG_MM] = NI; col.t[TDG_MI] = NI; col.t[TDG_MD] = NI;
			col.t[TDG_II] = p2s(0.195);
			col.t[TDG_IM] = NI;
			col.t[TDG_DD] = NI; col.t[TDG_DM] = NI;
		}
	}
	if (sg.type == 'R') {
		for (int i = 0; i < nh; i++) sg.sI[i][0] = p2s(1.0 / (float)nh);
		Column& col = sg.hmm[0].col[0];
		for (int c = 0; c < 5; c++) { col.m_emit[c] = background[c]; col.i_emit[c] = background[c]; }
		col.t[TDG_MM] = NI; col.t[TDG_MI] = NI; col.t[TDG_MD] = NI; col.t[TDG_MSKIP] = NI;
		col.t[TDG_II] = p2s(1.0 - 1.0 / (float)assumed_length);
		col.t[TDG_IM] = NI;
		col.t[TDG_ISKIP] = p2s(1.0 / (float)assumed_length);
		col.t[TDG_DD] = NI; col.t[TDG_DM] = NI;
		sg.skip = p2s(0.0);
	}
}

}  // namespace

struct tdg_arch {
	tdg_model_desc desc;
	std::string seg_type;
	std::vector<int32_t> seg_num_hmms, seg_num_cols, label;
	std::vector<float> seg_skip, background, transition, m_emit, i_emit, silent_to_M, silent_to_I, tmat;
};

extern "C" float tdg_logsum_host(float a, float b);

static thread_local std::string g_arch_err;
extern "C" const char* tdg_arch_last_error(void) { return g_arch_err.c_str(); }

extern "C" int tdg_arch_compile(int num_segments, const char* const* segment_strings, const tdg_arch_params* p, tdg_arch** out)
{
	char buf[256];
	if (!out || !segment_strings || !p) return TDG_EINVAL;
	*out = nullptr;
	if (num_segments < 1 || num_segments > TDG_MAX_SEGMENTS) return TDG_EINVAL;
	std::vector<Segment> segs(num_segments);
	// ---- assign_segment_sequences (interface.c:489-598)
	for (int s = 0; s < num_segments; s++) {
		const char* tmp = segment_strings[s];
		Segment& sg = segs[s];
		if (!tmp || !strchr("RGOPSFB", tmp[0]) || tmp[0] == 0 || tmp[1] != ':') {
			snprintf(buf, sizeof buf, "Segment type :%c not recognized.", tmp ? tmp[0] : '?');
			g_arch_err = buf;
			return TDG_FAIL;
		}
		sg.type = tmp[0];
		if (sg.type == 'R') sg.seqs.push_back("N");
		else {
			std::string cur;
			for (const char* q = tmp + 2; *q; q++) {
				if (*q != ',') cur.push_back(*q);
				else { sg.seqs.push_back(cur); cur.clear(); }
			}
			sg.seqs.push_back(cur);
			if (sg.type == 'B' || sg.type == 'S') sg.seqs.push_back(std::string(sg.seqs[0].size(), 'N'));
		}
		const size_t l0 = sg.seqs[0].size();
		if (l0 == 0) { g_arch_err = "empty segment sequence"; return TDG_FAIL; }
		for (auto& q : sg.seqs)
			if (q.size() != l0) { g_arch_err = "all sequences of a segment must have the same length"; return TDG_FAIL; }
	}
	// ---- init_model_bag (barcode_hmm.c:5760-6011)
	const int average_raw_length = (int)p->average_length;
	int read_length = (int)p->average_length;
	for (int s = 0; s < num_segments; s++) {
		const Segment& sg = segs[s];
		if (sg.type == 'G') read_length = read_length - 2;
		else if (sg.type == 'R') {}
		else if (sg.type == 'P') read_length = read_length - (int)sg.seqs[0].size() / 2;
		else read_length = read_length - (int)sg.seqs[0].size();
	}
	if (read_length < 20) read_length = 20;
	const float base_error = p->sequencer_error_rate, indel_freq = p->indel_frequency;
	for (int s = 0; s < num_segments; s++) {
		Segment& sg = segs[s];
		const int nh = (int)sg.seqs.size(), len = (int)sg.seqs[0].size();
		sg.hmm.assign(nh, Hmm());
		for (auto& h : sg.hmm) h.col.assign(len, Column());
		sg.sM.assign(nh, std::vector<float>());
		sg.sI.assign(nh, std::vector<float>());
		int segment_length = 0;
		if (sg.type == 'G') segment_length = 2;
		if (sg.type == 'R') segment_length = read_length;
		init_segment(sg, base_error, indel_freq, p->background_logp, segment_length);
	}
	// 5' partial segment (:5823-5884)
	if (p->expected_5_len) {
		Segment& mp = segs[0];
		double sum_prob = p2s(0.0);
		const int nh = (int)mp.hmm.size();
		if ((int)p->expected_5_len > (int)mp.hmm[0].col.size()) { g_arch_err = "expected_5_len longer than segment 0"; return TDG_EINVAL; }
		for (int i = 0; i < nh; i++) {
			for (int j = 0; j < p->expected_5_len; j++) {
				mp.sM[i][j] = p2s(1.0 / (float)nh) + p2s(gaussian_pdf(j, p->expected_5_len - p->mean_5_len, p->stdev_5_len));
				sum_prob = tdg_logsum_host((float)sum_prob, mp.sM[i][j]);
			}
			set_hmm_transition_parameters(mp.hmm[i], (int)p->expected_5_len, p->sequencer_error_rate, p->indel_frequency, -1.0, -1.0);
		}
		mp.skip = p2s(gaussian_pdf(p->expected_5_len, p->mean_5_len - p->expected_5_len, p->stdev_5_len));
		sum_prob = tdg_logsum_host((float)sum_prob, mp.skip);
		for (int i = 0; i < nh; i++)
			for (int j = 0; j < p->expected_5_len; j++) mp.sM[i][j] = mp.sM[i][j] - sum_prob;
		mp.skip = mp.skip - sum_prob;
	}
	// 3' partial segment (:5887-5901)
	if (p->expected_3_len) {
		Segment& mp = segs[num_segments - 1];
		double sum_prob = 0;
		if ((int)p->expected_3_len > (int)mp.hmm[0].col.size()) { g_arch_err = "expected_3_len longer than the last segment"; return TDG_EINVAL; }
		for (int i = 0; i < p->expected_3_len; i++) sum_prob += gaussian_pdf(i, p->mean_3_len, p->stdev_3_len);
		mp.skip = p2s(gaussian_pdf(0, p->mean_3_len, p->stdev_3_len) / sum_prob);
		const int nh = (int)mp.hmm.size();
		for (int i = 0; i < nh; i++) {
			mp.sM[i][0] = p2s(1.0 / (float)nh) + p2s(1.0 - gaussian_pdf(0, p->mean_3_len, p->stdev_3_len) / sum_prob);
			set_hmm_transition_parameters(mp.hmm[i], (int)p->expected_3_len, p->sequencer_error_rate, p->indel_frequency,
			                              p->mean_3_len, p->stdev_3_len);
		}
	}
	// internal P segments (:5903-5914)
	for (int c = 1; c < num_segments - 1; c++) {
		if (segs[c].type == 'P') {
			Segment& mp = segs[c];
			const int len = (int)mp.hmm[0].col.size();
			for (auto& h : mp.hmm) set_hmm_transition_parameters(h, len, p->sequencer_error_rate, p->indel_frequency, 0.1, -1.0);
		}
	}
	// calibration edit (calibrateQ.c:67-86)
	if (p->calibration_edit) {
		for (auto& sg : segs) {
			if (sg.type == 'B' || sg.type == 'S') {
				const int nh = (int)sg.hmm.size();
				for (int j = 0; j < nh - 1; j++) sg.sM[j][0] = p2s(1.0 / (float)(nh - 1));
				sg.sM[nh - 1][0] = p2s(0.0);
			}
		}
	}
	// ---- flatten
	auto* a = new tdg_arch();
	int H = 0, C = 0;
	for (auto& sg : segs) { H += (int)sg.hmm.size(); C += (int)(sg.hmm.size() * sg.hmm[0].col.size()); }
	if (H > TDG_MAX_HMMS) { delete a; g_arch_err = "too many HMMs"; return TDG_EINVAL; }
	a->background.assign(segs[0].bg, segs[0].bg + 5);
	for (int s = 0; s < num_segments; s++) {
		Segment& sg = segs[s];
		a->seg_type.push_back(sg.type);
		a->seg_num_hmms.push_back((int)sg.hmm.size());
		a->seg_num_cols.push_back((int)sg.hmm[0].col.size());
		a->seg_skip.push_back(sg.skip);
		for (size_t f = 0; f < sg.hmm.size(); f++) {
			int lab = (int)((f << 16) | (unsigned)s);
			if (sg.skip != p2s(0.0)) lab |= 0x80000000;
			a->label.push_back(lab);
			for (size_t g = 0; g < sg.hmm[f].col.size(); g++) {
				const Column& col = sg.hmm[f].col[g];
				a->transition.insert(a->transition.end(), col.t, col.t + 9);
				a->m_emit.insert(a->m_emit.end(), col.m_emit, col.m_emit + 5);
				a->i_emit.insert(a->i_emit.end(), col.i_emit, col.i_emit + 5);
				a->silent_to_M.push_back(sg.sM[f][g]);
				a->silent_to_I.push_back(sg.sI[f][g]);
			}
		}
	}
	// label-DP transition matrix (:5978-6006), same loop
	a->tmat.assign((size_t)H * H, 0.0f);
	for (int i = 0; i < H; i++) {
		int c = 1;
		for (int j = i + 1; j < H; j++) {
			a->tmat[(size_t)i * H + j] = 0;
			if ((a->label[i] & 0xFFFF) + 1 == (a->label[j] & 0xFFFF)) a->tmat[(size_t)i * H + j] = 1;
			if (((a->label[i] & 0xFFFF) < (a->label[j] & 0xFFFF)) && c) a->tmat[(size_t)i * H + j] = 1;
			if (!(a->label[j] & 0x80000000)) c = 0;
		}
		a->tmat[(size_t)i * H + i] = 1;
	}
	tdg_model_desc& d = a->desc;
	d.num_segments = num_segments; d.total_hmms = H; d.total_columns = C; d.average_raw_length = average_raw_length;
	d.seg_type = a->seg_type.c_str();
	d.seg_num_hmms = a->seg_num_hmms.data(); d.seg_num_cols = a->seg_num_cols.data(); d.seg_skip = a->seg_skip.data();
	d.background = a->background.data(); d.transition = a->transition.data(); d.m_emit = a->m_emit.data();
	d.i_emit = a->i_emit.data(); d.silent_to_M = a->silent_to_M.data(); d.silent_to_I = a->silent_to_I.data();
	d.label = a->label.data(); d.transition_matrix = a->tmat.data();
	*out = a;
	return TDG_OK;
}

extern "C" const tdg_model_desc* tdg_arch_desc(const tdg_arch* a) { return a ? &a->desc : nullptr; }
extern "C" void tdg_arch_destroy(tdg_arch* a) { delete a; }
