// MEASUREMENT HARNESS, not product code: the host stages of tdg_demux_run (line pass, lengths + pack, format + write)
// with the device stage replaced by stubs that answer at once -- how fast the host side alone moves a FASTQ file,
// on a machine without a GPU.  Includes the product's tdg_stream.cpp unchanged and supplies the handful of tdg_* entry
// points it calls.  The packing is the product's (tdg_pack.h, called from the stream layer) into plain
// host memory instead of pinned staging.
//
//   nvcc -x cu -O3 -std=c++17 -Xcompiler -O2 scripts/micro/host_pipeline.cpp -o /tmp/host_pipeline -lpthread
//   /tmp/host_pipeline in.fq outprefix threads devices [chunk_reads]
#include "../../tagdust_b200/csrc/tdg_stream.cpp"

struct tdg_model { int max_len = 160; int H = 49; };
struct tdg_batch {
	int max_reads = 0, max_len = 0, words = 0, n = 0;
	uint32_t* h_seq = nullptr; int32_t* h_len = nullptr;
	float* mapq = nullptr; int32_t *read_type = nullptr, *barcode = nullptr, *fingerprint = nullptr;
	uint8_t* extracted = nullptr; uint16_t* spans = nullptr;
};
static thread_local std::string g_err;
static int g_devices = 1;

namespace tdg {
int set_last_error(int code, const char* msg) { g_err = msg; return code; }
int batch_acquire(tdg_context*, int max_reads, int max_len, tdg_batch** out)
{
	auto* b = new tdg_batch();
	b->max_reads = max_reads; b->max_len = max_len; b->words = (max_len + 1 + 7) / 8;
	const size_t tiles = ((size_t)max_reads + 31) / 32;
	b->h_seq = (uint32_t*)calloc(tiles * b->words * 32, 4);
	b->h_len = (int32_t*)calloc(max_reads, 4);
	b->mapq = (float*)malloc((size_t)max_reads * 4);
	b->read_type = (int32_t*)calloc(max_reads, 4);
	b->barcode = (int32_t*)malloc((size_t)max_reads * 4);
	b->fingerprint = (int32_t*)malloc((size_t)max_reads * 4);
	b->extracted = (uint8_t*)malloc(max_reads);
	b->spans = (uint16_t*)malloc((size_t)max_reads * 2 * 2 * 2);
	for (int r = 0; r < max_reads; r++) {
		b->mapq[r] = 37.25f; b->barcode[r] = (int)(((unsigned)r * 2654435761u) >> 8) % 48; b->fingerprint[r] = -1; b->extracted[r] = 1;
		b->spans[(size_t)r * 4] = 6; b->spans[(size_t)r * 4 + 1] = 144; b->spans[(size_t)r * 4 + 2] = 0; b->spans[(size_t)r * 4 + 3] = 0;
	}
	*out = b;
	return TDG_OK;
}
void batch_release(tdg_batch* b) { tdg_batch_destroy(b); }
int batch_prepare(tdg_batch*, const tdg_model*, bool) { return TDG_OK; }
int scratch_prepare(tdg_context*, tdg_model*) { return TDG_OK; }

int batch_text_target(tdg_batch* b, int n, TextTarget* t)
{
	if (b->n + n > b->max_reads) return set_last_error(TDG_EINVAL, "batch overflow");
	t->seq = b->h_seq; t->len = b->h_len; t->words = b->words; t->max_len = b->max_len; t->first = b->n;
	return TDG_OK;
}
int batch_text_commit(tdg_batch* b, int n) { b->n += n; return TDG_OK; }
}  // namespace tdg

extern "C" {
int tdg_batch_clear(tdg_batch* b) { b->n = 0; return TDG_OK; }
void tdg_batch_destroy(tdg_batch* b)
{
	if (!b) return;
	free(b->h_seq); free(b->h_len); free(b->mapq); free(b->read_type); free(b->barcode); free(b->fingerprint); free(b->extracted); free(b->spans);
	delete b;
}
int tdg_device_count(const tdg_context*) { return g_devices; }
const char* tdg_last_error(void) { return g_err.c_str(); }
int tdg_model_max_len(const tdg_model* m) { return m->max_len; }
int tdg_model_set_max_len(tdg_model* m, int v) { m->max_len = v; return TDG_OK; }
int tdg_model_num_hmms(const tdg_model* m) { return m->H; }
int tdg_model_read_hmms(const tdg_model* m, uint8_t* is_read) { for (int h = 0; h < m->H; h++) is_read[h] = h == m->H - 1; return TDG_OK; }
int tdg_submit(tdg_context*, tdg_model*, int, const tdg_run_params*, tdg_batch*) { return TDG_OK; }
int tdg_wait(tdg_batch* b, tdg_result* r)
{
	memset(r, 0, sizeof *r);
	r->n_reads = b->n; r->mapq = b->mapq; r->read_type = b->read_type; r->extracted = b->extracted; r->barcode = b->barcode;
	r->fingerprint = b->fingerprint; r->span_stride = 2; r->spans = b->spans;
	return TDG_OK;
}
}

int main(int argc, char** argv)
{
	if (argc < 5) { fprintf(stderr, "usage: %s in.fq outprefix threads devices [chunk_reads]\n", argv[0]); return 2; }
	g_devices = atoi(argv[4]);
	tdg_model model;
	std::vector<std::string> names;
	std::vector<const char*> np;
	for (int k = 0; k < 48; k++) { char b[16]; snprintf(b, sizeof b, "BC%02d", k); names.push_back(b); }
	for (auto& s : names) np.push_back(s.c_str());
	tdg_demux_input in;
	memset(&in, 0, sizeof in);
	in.path = argv[1]; in.fasta = -1; in.model = &model; in.num_read_segments = 1; in.confidence_threshold = 20.0f; in.max_seq_len = 150; in.expected_len = 150;
	tdg_demux_job job;
	memset(&job, 0, sizeof job);
	job.n_inputs = 1; job.inputs = &in; job.barcode_input = 0; job.num_alternatives = 49; job.barcode_names = np.data();
	job.outfile = argv[2]; job.minlen = 16; job.dust = 100; job.matchstart = -1; job.matchend = -1; job.threads = atoi(argv[3]);
	job.chunk_reads = argc > 5 ? atoi(argv[5]) : 0;
	for (int rep = 0; rep < 2; rep++) {
		tdg_demux_stats st;
		memset(&st, 0, sizeof st);
		const double t0 = now_s();
		const int rc = tdg_demux_run((tdg_context*)&model, &job, &st);
		const double dt = now_s() - t0;
		printf("rc=%d reads=%lld  %.3f s  %.2f M reads/s   busy: split %.3f  convert %.3f  gpu-wait %.3f  write %.3f  (%s)\n", rc, (long long)st.total_read, dt,
		       st.total_read / dt / 1e6, st.seconds_split, st.seconds_parse, st.seconds_gpu_wait, st.seconds_write, rc ? g_err.c_str() : "ok");
	}
	return 0;
}
