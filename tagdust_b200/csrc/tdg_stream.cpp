// tdg_stream.cpp -- streaming FASTQ ingest + demultiplexed output around the GPU decode path
// (include/tagdust_b200_stream.h).  Host-only C++; the GPU is reached through the public C ABI.
//
// Reference behaviour reproduced (files under the reference's src/):
//   read_fasta_fastq()            io.c:1684-1815   line state machine, name/seq/qual up to the first
//                                                  control character, nuc_code[] conversion
//   io_handler()                  io.c:382-608     suffix rules, zcat/bzcat pipes
//   hmm_controller_multiple()     barcode_hmm.c:243-384   per-chunk loop, cross-file merge, tallies
//   run_rna_dust()/do_rna_dust()  barcode_hmm.c:2043, :2370   files whose architecture is a single R segment
//   dust_sequences()              barcode_hmm.c:2407-2467
//   make_extracted_read()         barcode_hmm.c:3325-3356
//   print_all()                   io.c:757-1016    output naming, headers, spacer-split records
//
// Design: three stages on their own threads -- parse (block reads, one memchr pass to split
// lines, then conversion/packing on a worker pool), GPU (tdg_submit of chunk k+1 is queued before
// tdg_wait of chunk k), write (extraction rewrite + formatting on the pool into per-thread,
// per-file buffers, flushed in thread order so every file keeps input order).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "../../include/tagdust_b200_stream.h"
#include "tdg_device.h"
#include "tdg_pack.h"

namespace tdg {
// tdg_host.cu
int set_last_error(int code, const char* msg);
int batch_acquire(tdg_context* ctx, int max_reads, int max_len, tdg_batch** out);
void batch_release(tdg_batch* b);
int batch_prepare(tdg_batch* b, const tdg_model* m, bool host_labels);
int scratch_prepare(tdg_context* ctx, tdg_model* m);
int batch_text_target(tdg_batch* b, int n, TextTarget* t);
int batch_text_commit(tdg_batch* b, int n);
}

namespace {

constexpr int kMaxLine = 10000;  // tagdust2.h:96; fgets() hands out at most kMaxLine-1 characters per call

double now_s()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int failf(int code, const char* fmt, ...)
{
	char buf[768];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	return tdg::set_last_error(code, buf);
}

// One set of worker threads for all the stages of a streaming job.  Every stage thread hands its loops to the same
// workers part by part, so the stages together never run more threads than the job was given (three stages that each
// start `threads` threads of their own oversubscribe the cores three times, and every loop then waits for its slowest,
// descheduled thread), and no loop pays for thread creation.
class WorkerPool {
public:
	explicit WorkerPool(int workers)
	{
		for (int k = 0; k < workers; k++) th_.emplace_back([this] { work(); });
	}
	~WorkerPool()
	{
		{ std::lock_guard<std::mutex> l(mu_); stop_ = true; }
		cv_work_.notify_all();
		for (auto& t : th_) t.join();
	}
	// fn(part) for every part in [0, parts), in any order, on the workers and on the caller; returns when all are done
	void run(int parts, const std::function<void(int)>& fn)
	{
		if (parts <= 0) return;
		Job j;
		j.fn = &fn; j.parts = parts;
		std::unique_lock<std::mutex> l(mu_);
		jobs_.push_back(&j);
		if (parts > 1) cv_work_.notify_all();
		while (j.next < j.parts) {
			const int p = take(&j);
			l.unlock();
			fn(p);
			l.lock();
			j.done++;
		}
		cv_done_.wait(l, [&] { return j.done == j.parts; });
	}
private:
	struct Job { const std::function<void(int)>* fn = nullptr; int parts = 0, next = 0, done = 0; };
	int take(Job* j)  // mu_ held, j->next < j->parts
	{
		const int p = j->next++;
		if (j->next == j->parts) jobs_.erase(std::find(jobs_.begin(), jobs_.end(), j));
		return p;
	}
	void work()
	{
		std::unique_lock<std::mutex> l(mu_);
		for (;;) {
			cv_work_.wait(l, [&] { return stop_ || !jobs_.empty(); });
			if (jobs_.empty()) return;  // stop_
			Job* j = jobs_[rr_++ % jobs_.size()];   // the stages take turns
			const int p = take(j);
			l.unlock();
			(*j->fn)(p);
			l.lock();
			if (++j->done == j->parts) cv_done_.notify_all();
		}
	}
	std::mutex mu_;
	std::condition_variable cv_work_, cv_done_;
	std::vector<Job*> jobs_;   // jobs with parts left to hand out
	std::vector<std::thread> th_;
	size_t rr_ = 0;
	bool stop_ = false;
};
thread_local WorkerPool* tl_pool = nullptr;   // set by the stage threads of tdg_demux_run

// fn(begin, end, part_index) over [0, n) in up to `threads` equal parts (part_index < threads): on the streaming job's
// worker pool when the calling thread belongs to one, else on threads of its own (the caller is one of them)
template <class F>
void parallel_for(int threads, size_t n, size_t grain, F fn)
{
	int T = std::max(1, threads);
	if (grain > 0) T = (int)std::min<size_t>(T, std::max<size_t>(1, n / grain));
	if (T <= 1) { fn((size_t)0, n, 0); return; }
	const size_t per = (n + T - 1) / T;
	// a part that runs out of memory on a worker must not take the process down: the caller gets the exception
	std::atomic<bool> oom{false};
	const std::function<void(int)> part = [&](int t) {
		try { fn(std::min(n, per * (size_t)t), std::min(n, per * (size_t)(t + 1)), t); }
		catch (const std::bad_alloc&) { oom.store(true); }
	};
	if (tl_pool) {
		tl_pool->run(T, part);
	} else {
		std::vector<std::thread> th;
		for (int t = 1; t < T; t++) th.emplace_back([&part, t] { part(t); });
		part(0);
		for (auto& x : th) x.join();
	}
	if (oom.load()) throw std::bad_alloc();
}

struct NucTable {
	uint8_t code[256];
	uint8_t ctl[256];
	NucTable()
	{
		for (int i = 0; i < 256; i++) { code[i] = 4; ctl[i] = (i < 32 || i == 127) ? 1 : 0; }  // iscntrl() in the C locale
		code[46] = 5;                                                                             // '.' (nuc_code.c:52)
		code['A'] = code['a'] = 0; code['C'] = code['c'] = 1; code['G'] = code['g'] = 2;
		code['T'] = code['t'] = 3; code['U'] = code['u'] = 3;
		for (int i = 0; i < 256; i++) out[i] = "ACGTNN"[code[i]];   // print_all's alphabet (io.c:757) of the code of a character
	}
	char out[256];
};
const NucTable kNuc;

// Length of the prefix of s[0..n) without a control character (iscntrl() in the C locale: < 32 or 127), eight bytes at
// a time: the reference stops names, sequences and qualities at the first one (io.c:1716-1790).
inline uint32_t ctl_span(const uint8_t* s, uint32_t n)
{
	uint32_t i = 0;
#if defined(__SSE2__)
	{
		const __m128i k1f = _mm_set1_epi8(0x1F), k7f = _mm_set1_epi8(0x7F);
		while (i + 16 <= n) {
			const __m128i x = _mm_loadu_si128((const __m128i*)(s + i));
			const __m128i c = _mm_or_si128(_mm_cmpeq_epi8(_mm_min_epu8(x, k1f), x), _mm_cmpeq_epi8(x, k7f));   // <= 0x1F, == 0x7F
			const int m = _mm_movemask_epi8(c);
			if (m) return i + (uint32_t)__builtin_ctz((unsigned)m);
			i += 16;
		}
	}
#endif
	const uint64_t k7f = 0x7F7F7F7F7F7F7F7FULL, k80 = 0x8080808080808080ULL, k01 = 0x0101010101010101ULL;
	while (i + 8 <= n) {
		uint64_t x;
		memcpy(&x, s + i, 8);
		const uint64_t ge20 = ((x & k7f) + 0x6060606060606060ULL) | x;   // high bit per byte: low 7 bits >= 0x20, or byte >= 0x80
		const uint64_t z = x ^ k7f;
		const uint64_t m = (~ge20 & k80) | ((z - k01) & ~z & k80);         // bytes < 0x20, bytes == 0x7F
		if (m) return i + (uint32_t)(__builtin_ctzll(m) >> 3);
		i += 8;
	}
	while (i < n && !kNuc.ctl[s[i]]) i++;
	return i;
}

bool has_suffix(const std::string& s, const char* suf)
{
	const size_t n = strlen(suf);
	return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// parsed chunk storage
// ------------------------------------------------------------------------------------------
// Growable POD array without value-initialisation: a std::vector would memset hundreds of MB on one
// thread (and take all the first-touch page faults there) every time a chunk buffer grows.
template <class T>
struct RawVec {
	T* p = nullptr;
	size_t n = 0, cap = 0;
	RawVec() = default;
	RawVec(const RawVec&) = delete;
	RawVec& operator=(const RawVec&) = delete;
	RawVec(RawVec&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
	~RawVec() { free(p); }
	void reserve(size_t want)
	{
		if (want <= cap) return;
		size_t nc = std::max(want, cap + cap / 2 + 64);
		T* q = (T*)realloc(p, nc * sizeof(T));
		if (!q) throw std::bad_alloc();
		p = q; cap = nc;
	}
	void resize(size_t m) { reserve(m); n = m; }
	void clear() { n = 0; }
	void push_back(const T& v) { if (n == cap) reserve(n + 1); p[n++] = v; }
	bool empty() const { return n == 0; }
	size_t size() const { return n; }
	T* data() { return p; }
	const T* data() const { return p; }
	T& operator[](size_t i) { return p[i]; }
	const T& operator[](size_t i) const { return p[i]; }
	T& back() { return p[n - 1]; }
};

struct ParsedChunk {
	int n = 0, max_len = 0;
	bool fasta = false;
	RawVec<int32_t> len;
	RawVec<uint64_t> seq_off, name_off;
	RawVec<uint8_t> codes, qual;
	RawVec<char> names;
	// lean form (the demultiplexer): no code / quality / name copies, everything is read from `text` again where it is
	// needed -- seq_pos / qual_pos / name_pos are text offsets, name_len the name's length up to its first control character
	RawVec<uint64_t> seq_pos, qual_pos, name_pos;
	RawVec<uint32_t> name_len;
	// line table of the sequential pass: offsets into `text`, piece lengths (without the newline)
	// `next`: text offset right behind the line that completed the entry (the first quality line of a FASTQ entry, the first
	// sequence line of a FASTA entry) -- where a chunk ends when this is its last entry (io.c:1797-1808)
	struct Rec { uint64_t name, seq, qual, next; uint32_t name_n, seq_n, qual_n; uint8_t has_seq, has_qual; };
	RawVec<Rec> recs;
	// the text the line table points into: the reader's mapping (plain files) or a buffer owned by
	// this chunk (pipes), so that the line pass of the next chunk can run while this one is converted
	const char* text = nullptr;
	RawVec<char> own_text;
	std::string path;
};

struct tdg_fastq {
	std::string path;
	bool fasta = false;
	// plain file: mmap; pipe: growing buffer
	const char* map = nullptr; size_t map_len = 0; int fd = -1;
	FILE* pipe = nullptr;
	RawVec<char> carry;        // pipe: bytes read but not consumed by the previous chunk
	size_t beg = 0, end = 0;   // plain file: unconsumed part [beg, end) of the mapping
	bool eof = false;
	int set = 0, seq_p = 0;    // read_fasta_fastq's flags; they persist across chunks like the FILE position does
	double bytes_per_entry = 0;  // running estimate (plain files): sizes the window the parallel line pass looks at
	ParsedChunk own;           // storage behind the public tdg_fastq_next
};

// pipe only: append another block to the chunk's own text; returns bytes added (0 at end of input)
static int reader_fill(tdg_fastq* f, ParsedChunk& pc, size_t* added)
{
	const size_t block = (size_t)8 << 20;
	const size_t have = pc.own_text.size();
	pc.own_text.reserve(have + block);
	const size_t got = fread(pc.own_text.data() + have, 1, block, f->pipe);
	if (got == 0) {
		if (ferror(f->pipe)) return failf(TDG_EIO, "read error on %s", f->path.c_str());
		f->eof = true;
	}
	pc.own_text.n = have + got;
	*added = got;
	return TDG_OK;
}

extern "C" int tdg_fastq_open(const char* path, int fasta, tdg_fastq** out)
{
	if (!path || !out) return failf(TDG_EINVAL, "tdg_fastq_open: NULL argument");
	*out = nullptr;
	const std::string p(path);
	if (access(path, R_OK) != 0) return failf(TDG_EIO, "Error: Cannot find input file: %s", path);  // io.c:403
	bool gz = false, bz = false, fa = false;
	// suffix rules of io_handler (io.c:410-457)
	if (has_suffix(p, ".gz")) gz = true;
	if (has_suffix(p, ".bz2")) bz = true;
	if (has_suffix(p, ".fa") || has_suffix(p, ".fasta") || has_suffix(p, ".fa.gz")) fa = true;
	if (has_suffix(p, ".sam") || has_suffix(p, ".bam") || has_suffix(p, ".sam.gz") || has_suffix(p, ".bam.gz"))
		return failf(TDG_EINVAL, "%s: SAM/BAM input is not handled by the streaming reader", path);
	auto* f = new tdg_fastq();
	f->path = p;
	f->fasta = fasta < 0 ? fa : fasta != 0;
	if (gz || bz) {
		const char* tool = bz ? "bzcat" : "zcat";
		if (gz && (access("/usr/bin/gzcat", X_OK) == 0 || access("/bin/gzcat", X_OK) == 0)) tool = "gzcat";
		// the path goes to sh -c inside single quotes: a quote in it is written as '\'' (close, escaped quote, reopen)
		std::string quoted;
		for (char ch : p) { if (ch == '\'') quoted += "'\\''"; else quoted += ch; }
		std::string cmd = std::string(tool) + " '" + quoted + "'";
		f->pipe = popen(cmd.c_str(), "r");
		if (!f->pipe) { delete f; return failf(TDG_EIO, "cannot run: %s", cmd.c_str()); }
	} else {
		f->fd = open(path, O_RDONLY);
		if (f->fd < 0) { delete f; return failf(TDG_EIO, "cannot open %s", path); }
		struct stat st;
		if (fstat(f->fd, &st) != 0) { close(f->fd); delete f; return failf(TDG_EIO, "cannot stat %s", path); }
		f->map_len = (size_t)st.st_size;
		if (f->map_len > 0) {
			void* m = mmap(nullptr, f->map_len, PROT_READ, MAP_PRIVATE, f->fd, 0);
			if (m == MAP_FAILED) { close(f->fd); delete f; return failf(TDG_EIO, "cannot mmap %s", path); }
			madvise(m, f->map_len, MADV_SEQUENTIAL);
			f->map = (const char*)m;
		} else {
			f->map = "";
		}
		f->end = f->map_len;
		f->eof = true;  // everything is addressable
	}
	*out = f;
	return TDG_OK;
}

extern "C" void tdg_fastq_close(tdg_fastq* f)
{
	if (!f) return;
	if (f->pipe) pclose(f->pipe);
	if (f->map && f->map_len) munmap((void*)f->map, f->map_len);
	if (f->fd >= 0) close(f->fd);
	delete f;
}

// ------------------------------------------------------------------------------------------
// line pass: read_fasta_fastq's state machine (io.c:1697-1796) over the text
// ------------------------------------------------------------------------------------------
// One fgets() piece starting at `pos`: up to the newline, or kMaxLine-1 characters of a longer line, or the unterminated
// rest of the input.  Returns false when [pos, end) holds no complete piece yet (pipes: more input is needed).
static inline bool next_piece(const char* base, size_t pos, size_t end, bool at_eof, size_t* piece, size_t* next)
{
	const char* nl = (pos < end) ? (const char*)memchr(base + pos, '\n', std::min(end - pos, (size_t)kMaxLine - 1)) : nullptr;
	if (nl) { *piece = (size_t)(nl - (base + pos)); *next = pos + *piece + 1; return true; }
	if (end - pos >= (size_t)kMaxLine - 1) { *piece = kMaxLine - 1; *next = pos + *piece; return true; }  // fgets() splits long lines
	if (at_eof && pos < end) { *piece = end - pos; *next = end; return true; }  // last line without newline
	return false;
}

// What one piece does to the machine: 1 = starts an entry ('@' / '>' with the flag clear), 2 = sequence line of the current
// entry, 3 = quality line of the current entry, 0 = nothing ('+' header, or a line the reference ignores).
struct LineState { int set, seq_p; };
static inline int line_step(LineState& st, const char* base, size_t pos, size_t piece)
{
	const char c0 = piece ? base[pos] : '\n';
	if ((c0 == '@' || c0 == '>') && !st.set) { st.seq_p = 1; st.set = 1; return 1; }
	if (c0 == '+' && !st.set) { st.seq_p = 0; st.set = 1; return 0; }
	const int ev = st.set ? (st.seq_p ? 2 : 3) : 0;
	st.set = 0;
	return ev;
}
static inline bool same_state(const LineState& a, const LineState& b) { return a.set == b.set && (!a.set || a.seq_p == b.seq_p); }

// The entries of one stretch of text plus the data lines that precede its first header ("orphans": they belong to the
// entry that was open when the stretch began).
struct SplitOut {
	RawVec<ParsedChunk::Rec> recs;
	struct Orphan { uint64_t pos, next; uint32_t n; uint8_t kind; };
	Orphan orphan[4]; int n_orphan = 0;
	bool dup = false;       // an entry received a second sequence / quality line: the sequential pass must decide where the chunk ends
	bool overflow = false;  // more orphans than fit (malformed input)
	void clear() { recs.clear(); n_orphan = 0; dup = false; overflow = false; }
	inline void apply(int ev, size_t pos, size_t piece, size_t next, bool fasta)
	{
		if (ev == 1) {
			ParsedChunk::Rec r;
			memset(&r, 0, sizeof r);
			r.name = pos; r.name_n = (uint32_t)piece;
			recs.push_back(r);
		} else if (ev) {
			if (recs.empty()) {
				if (n_orphan < 4) { orphan[n_orphan].pos = pos; orphan[n_orphan].next = next; orphan[n_orphan].n = (uint32_t)piece; orphan[n_orphan].kind = (uint8_t)ev; n_orphan++; }
				else overflow = true;
				return;
			}
			assign(recs.back(), ev, pos, piece, next, fasta, dup);
		}
	}
	static inline void assign(ParsedChunk::Rec& r, int ev, size_t pos, size_t piece, size_t next, bool fasta, bool& dup)
	{
		if (ev == 2) {
			if (r.has_seq) dup = true;
			if (fasta && !r.has_seq) r.next = next;
			r.seq = pos; r.seq_n = (uint32_t)piece; r.has_seq = 1;
		} else {
			if (r.has_qual) dup = true;
			if (!fasta && !r.has_qual) r.next = next;
			r.qual = pos; r.qual_n = (uint32_t)piece; r.has_qual = 1;
		}
	}
};

// Sequential pass (pipes, and the fallback of the parallel pass): split lines, run the machine, fill pc.recs.
static int split_lines_serial(tdg_fastq* f, int max_reads, ParsedChunk& pc)
{
	size_t pos, end;
	if (f->pipe) {  // start from what the previous chunk left over
		pc.own_text.clear();
		pc.own_text.reserve(f->carry.size() + 1);
		if (f->carry.size()) memcpy(pc.own_text.data(), f->carry.data(), f->carry.size());
		pc.own_text.n = f->carry.size();
		f->carry.clear();
		pos = 0; end = pc.own_text.size();
	} else {
		pc.text = f->map;
		pos = f->beg; end = f->end;
	}
	LineState st = {f->set, f->seq_p};
	bool done = false, dup = false;
	while (!done) {
		const char* base = f->pipe ? pc.own_text.data() : f->map;
		size_t piece, next;
		if (!next_piece(base, pos, end, f->eof, &piece, &next)) {
			if (f->eof) break;
			size_t added = 0;
			int rc = reader_fill(f, pc, &added);
			if (rc) return rc;
			end = pc.own_text.size();
			continue;
		}
		const int ev = line_step(st, base, pos, piece);
		if (ev == 1) {
			ParsedChunk::Rec r;
			memset(&r, 0, sizeof r);
			r.name = pos; r.name_n = (uint32_t)piece;
			pc.recs.push_back(r);
		} else if (ev && !pc.recs.empty()) SplitOut::assign(pc.recs.back(), ev, pos, piece, next, f->fasta, dup);
		pos = next;
		if ((int)pc.recs.size() == max_reads) {  // io.c:1797-1808: the chunk ends once its last entry is complete
			const ParsedChunk::Rec& r = pc.recs.back();
			if ((!f->fasta && r.has_qual) || (f->fasta && r.has_seq)) done = true;
		}
	}
	f->set = st.set; f->seq_p = st.seq_p;
	if (f->pipe) {
		pc.text = pc.own_text.data();
		f->carry.resize(end - pos);
		if (end > pos) memcpy(f->carry.data(), pc.own_text.data() + pos, end - pos);
	} else {
		f->beg = pos;
	}
	pc.n = (int)pc.recs.size();
	return TDG_OK;
}

// Parallel pass over a memory-mapped file.  The text is cut at line starts into one stretch per thread.  A thread does not
// know the machine's state at its first line, so it runs the three possible states (flag clear; flag set expecting a
// sequence line; flag set expecting a quality line) side by side until they agree -- which takes a handful of lines, because
// every data line clears the flag -- keeps what each of them produced up to there, and runs one machine from there on.
// The stretches are then stitched in order: the state the previous stretch ended in selects the prefix that was right.
// Returns 1 when the chunk was produced, 0 when the caller has to use the sequential pass (malformed or unusual input:
// an entry with two sequence / quality lines, a last entry that is not complete, stretches that never agree).
struct SplitPart {
	SplitOut pre[3], rest;
	LineState pre_end[3], end_state;
	bool converged = false;
	size_t beg = 0, end = 0;
};

static int split_lines_parallel(tdg_fastq* f, int max_reads, int threads, ParsedChunk& pc, std::vector<SplitPart>& parts)
{
	const char* base = f->map;
	const size_t file_end = f->end;
	const bool fasta = f->fasta;
	pc.text = base;
	size_t pos = f->beg;
	LineState st = {f->set, f->seq_p};
	bool dup = false;
	const int T = std::max(1, threads);
	if ((int)parts.size() < T) parts.resize((size_t)T);
	while ((int)pc.recs.size() < max_reads && pos < file_end) {
		// window: what the missing entries should need, a little more, at least 1 MB
		const double bpe = f->bytes_per_entry > 0 ? f->bytes_per_entry : 512.0;
		size_t want = (size_t)((double)(max_reads - (int)pc.recs.size()) * bpe * 1.02) + ((size_t)1 << 20);
		size_t wend = (file_end - pos <= want) ? file_end : pos + want;
		if (wend < file_end) {  // cut behind a newline
			const char* nl = (const char*)memchr(base + wend, '\n', file_end - wend);
			wend = nl ? (size_t)(nl - base) + 1 : file_end;
		}
		const size_t wbytes = wend - pos;
		const int P = (int)std::min<size_t>((size_t)T, std::max<size_t>(1, wbytes >> 18));  // >= 256 KB per stretch
		for (int k = 0; k < P; k++) {
			size_t b = pos + wbytes * (size_t)k / (size_t)P;
			if (k > 0) {
				const char* nl = (b < wend) ? (const char*)memchr(base + b - 1, '\n', wend - (b - 1)) : nullptr;
				b = nl ? (size_t)(nl - base) + 1 : wend;
			}
			parts[k].beg = b;
		}
		for (int k = 0; k < P; k++) parts[k].end = (k + 1 < P) ? parts[k + 1].beg : wend;
		parallel_for(P, (size_t)P, 0, [&](size_t kb, size_t ke, int) {
			for (size_t k = kb; k < ke; k++) {
				SplitPart& sp = parts[k];
				for (auto& o : sp.pre) o.clear();
				sp.rest.clear();
				sp.converged = false;
				const bool last_part = sp.end == file_end;  // only there can a line lack its newline
				LineState h[3] = {{0, 0}, {1, 1}, {1, 0}};
				size_t p = sp.beg, piece, next;
				int lines = 0;
				while (!(same_state(h[0], h[1]) && same_state(h[0], h[2]))) {
					if (lines++ >= 64 || !next_piece(base, p, sp.end, last_part, &piece, &next)) break;
					for (int q = 0; q < 3; q++) sp.pre[q].apply(line_step(h[q], base, p, piece), p, piece, next, fasta);
					p = next;
				}
				for (int q = 0; q < 3; q++) sp.pre_end[q] = h[q];
				if (!(same_state(h[0], h[1]) && same_state(h[0], h[2]))) { sp.end_state = h[0]; continue; }  // stitched serially
				sp.converged = true;
				LineState c = h[0];
				SplitOut& out = sp.rest;
				out.recs.reserve((size_t)((double)(sp.end - p) / bpe * 1.1) + 16);
				while (next_piece(base, p, sp.end, last_part, &piece, &next)) {
					out.apply(line_step(c, base, p, piece), p, piece, next, fasta);
					p = next;
				}
				sp.end_state = c;
			}
		});
		// stitch: the order of the stretches and the state each one starts in are settled one after the other; the line
		// tables themselves (48 bytes per entry) are copied into place by all workers afterwards, and the lines a stretch
		// found for the entry the stretch before it ended in (orphans) are applied to that entry last
		const size_t recs_before = pc.recs.size();
		struct Move { size_t at; const ParsedChunk::Rec* src; size_t n; };
		struct Late { size_t at; const SplitOut* o; int q; };
		std::vector<Move> moves;
		std::vector<Late> late;
		size_t total = pc.recs.size();
		auto merge = [&](SplitOut& o) -> bool {
			if (o.dup || o.overflow) return false;
			for (int q = 0; q < o.n_orphan; q++)
				if (total) late.push_back({total - 1, &o, q});
			if (o.recs.size()) { moves.push_back({total, o.recs.data(), o.recs.size()}); total += o.recs.size(); }
			return true;
		};
		auto settle = [&] {
			pc.recs.resize(total);
			ParsedChunk::Rec* dst = pc.recs.data();
			parallel_for(T, moves.size(), 0, [&](size_t mb, size_t me, int) {
				for (size_t m = mb; m < me; m++) memcpy(dst + moves[m].at, moves[m].src, moves[m].n * sizeof(ParsedChunk::Rec));
			});
			for (const Late& l : late) {
				const auto& x = l.o->orphan[l.q];
				SplitOut::assign(dst[l.at], x.kind, x.pos, x.n, x.next, fasta, dup);
			}
		};
		for (int k = 0; k < P; k++) {
			SplitPart& sp = parts[k];
			const int q = !st.set ? 0 : (st.seq_p ? 1 : 2);
			if (!sp.converged) {
				// the three machines did not agree within the stretch's first lines: run the right one over the whole stretch
				const bool last_part = sp.end == file_end;
				SplitOut& out = sp.rest;
				out.clear();
				size_t p = sp.beg, piece, next;
				while (next_piece(base, p, sp.end, last_part, &piece, &next)) {
					out.apply(line_step(st, base, p, piece), p, piece, next, fasta);
					p = next;
				}
				if (!merge(out)) return 0;
				continue;
			}
			if (!merge(sp.pre[q]) || !merge(sp.rest)) return 0;
			st = sp.end_state;
		}
		settle();
		if (dup) return 0;
		if (pc.recs.size() > recs_before) f->bytes_per_entry = (double)wbytes / (double)(pc.recs.size() - recs_before);
		pos = wend;
	}
	if ((int)pc.recs.size() >= max_reads) {
		const ParsedChunk::Rec& r = pc.recs[(size_t)max_reads - 1];
		if (!((!fasta && r.has_qual) || (fasta && r.has_seq)) || !r.next) return 0;
		// the chunk ends behind the line that completed its last entry; that was a data line, so the flag is clear there
		pc.recs.resize((size_t)max_reads);
		f->beg = r.next; f->set = 0; f->seq_p = fasta ? 1 : 0;
	} else {
		f->beg = pos; f->set = st.set; f->seq_p = st.seq_p;
	}
	pc.n = (int)pc.recs.size();
	return 1;
}

static int split_lines(tdg_fastq* f, int max_reads, ParsedChunk& pc, int threads = 1, std::vector<SplitPart>* scratch = nullptr)
{
	pc.n = 0; pc.max_len = 0; pc.fasta = f->fasta; pc.path = f->path;
	pc.recs.clear();
	if (max_reads < 1) return failf(TDG_EINVAL, "max_reads must be >= 1");
	static const bool no_parallel = getenv("TDG_SERIAL_SPLIT") != nullptr;   // tests: force the sequential pass
	if (!f->pipe && threads > 1 && !no_parallel && f->end - f->beg >= ((size_t)1 << 20)) {
		std::vector<SplitPart> local;
		const size_t beg0 = f->beg;
		const int set0 = f->set, seq0 = f->seq_p;
		if (split_lines_parallel(f, max_reads, threads, pc, scratch ? *scratch : local) == 1) return TDG_OK;
		if (getenv("TDG_TRACE")) fprintf(stderr, "[trace] %s: parallel line pass stepped aside at offset %zu, sequential pass takes this chunk\n", f->path.c_str(), beg0);
		f->beg = beg0; f->set = set0; f->seq_p = seq0;   // sequential pass from the same place
		pc.recs.clear();
	}
	return split_lines_serial(f, max_reads, pc);
}

// Conversion pass: names, codes, qualities on the worker pool.
static int convert_chunk(ParsedChunk& pc, int threads)
{
	struct { const char* path; bool fasta; } fv = {pc.path.c_str(), pc.fasta};
	auto* f = &fv;
	const double tt1 = now_s();
	const int n = pc.n;
	if (n == 0) return TDG_OK;
	// offsets (upper bounds: the piece lengths; the real lengths stop at the first control character)
	pc.len.resize(n); pc.seq_off.resize(n); pc.name_off.resize((size_t)n + 1);
	uint64_t so = 0, no = 0;
	for (int r = 0; r < n; r++) {
		const ParsedChunk::Rec& R = pc.recs[r];
		if (!R.has_seq) return failf(TDG_EFORMAT, "%s: entry %d has no sequence line", f->path, r);
		if (!f->fasta && !R.has_qual) return failf(TDG_EFORMAT, "%s: entry %d has no quality line", f->path, r);
		pc.seq_off[r] = so; so += (uint64_t)R.seq_n + 1;
		pc.name_off[r] = no; no += (uint64_t)R.name_n;  // '@' dropped, NUL added
	}
	pc.name_off[n] = no;
	pc.codes.reserve(so + 8);   // slack: the packer reads whole 8-byte words
	pc.codes.resize(so);
	if (!f->fasta) pc.qual.resize(so); else pc.qual.clear();
	pc.names.resize(no);
	const char* base = pc.text;
	std::atomic<int> bad{-1};
	std::vector<int> tmax((size_t)std::max(1, threads), 0);
	const bool fasta = f->fasta;
	parallel_for(threads, (size_t)n, 2048, [&](size_t b, size_t e, int t) {
		int mx = 0;
		for (size_t r = b; r < e; r++) {
			const ParsedChunk::Rec& R = pc.recs[r];
			// name: line[1..] up to the first control character (io.c:1723-1735)
			{
				const uint8_t* s = (const uint8_t*)base + R.name + 1;
				char* d = pc.names.data() + pc.name_off[r];
				const uint32_t i = ctl_span(s, R.name_n ? R.name_n - 1 : 0);
				memcpy(d, s, i);
				d[i] = 0;
			}
			uint8_t* c = pc.codes.data() + pc.seq_off[r];
			const uint8_t* s = (const uint8_t*)base + R.seq;
			const uint32_t i = ctl_span(s, R.seq_n);
			for (uint32_t k = 0; k < i; k++) c[k] = kNuc.code[s[k]];
			c[i] = 0;
			pc.len[r] = (int32_t)i;
			if ((int)i > mx) mx = (int)i;
			if (!fasta) {
				uint8_t* q = pc.qual.data() + pc.seq_off[r];
				const uint8_t* s2 = (const uint8_t*)base + R.qual;
				const uint32_t k = ctl_span(s2, R.qual_n);
				if (k != i) { int exp = -1; bad.compare_exchange_strong(exp, (int)r); }
				else memcpy(q, s2, i);
				q[i] = 0;
			}
		}
		tmax[t] = std::max(tmax[t], mx);
	});
	if (getenv("TDG_TRACE")) fprintf(stderr, "convert_chunk: %d reads, %.3f s\n", n, now_s() - tt1);
	if (bad.load() >= 0) return failf(TDG_EFORMAT, "ERROR: Length of sequence and base qualities differ!.");  // io.c:1770
	for (int v : tmax) pc.max_len = std::max(pc.max_len, v);
	return TDG_OK;
}

// Lean conversion (tdg_demux_run): lengths and validation only.  The 4-bit packing reads the sequence characters from
// the text (the per-read hook of this pass, tdg::pack_text_read), the writer formats names / bases / qualities from the text: the intermediate arrays
// of convert_chunk (about 1.5 x the input in writes, and as much again in reads by the writer) are never made.
template <class Hook>   // hook(r, first base, length): called once per read by the thread that measured it
static int lean_chunk(ParsedChunk& pc, int threads, Hook hook)
{
	const int n = pc.n;
	if (n == 0) return TDG_OK;
	const char* path = pc.path.c_str();
	pc.len.resize(n); pc.seq_pos.resize(n); pc.qual_pos.resize(n); pc.name_pos.resize(n); pc.name_len.resize(n);
	const char* base = pc.text;
	const bool fasta = pc.fasta;
	std::atomic<int> bad{-1}, noseq{-1}, noqual{-1};
	std::vector<int> tmax((size_t)std::max(1, threads), 0);
	parallel_for(threads, (size_t)n, 2048, [&](size_t b, size_t e, int t) {
		int mx = 0;
		for (size_t r = b; r < e; r++) {
			const ParsedChunk::Rec& R = pc.recs[r];
			if (!R.has_seq) { int exp = -1; noseq.compare_exchange_strong(exp, (int)r); continue; }
			if (!fasta && !R.has_qual) { int exp = -1; noqual.compare_exchange_strong(exp, (int)r); continue; }
			pc.name_pos[r] = R.name + 1;
			pc.name_len[r] = ctl_span((const uint8_t*)base + R.name + 1, R.name_n ? R.name_n - 1 : 0);
			const uint32_t i = ctl_span((const uint8_t*)base + R.seq, R.seq_n);
			pc.seq_pos[r] = R.seq;
			pc.len[r] = (int32_t)i;
			if ((int)i > mx) mx = (int)i;
			hook(r, (const uint8_t*)base + R.seq, (int)i);
			pc.qual_pos[r] = R.qual;
			if (!fasta && ctl_span((const uint8_t*)base + R.qual, R.qual_n) != i) { int exp = -1; bad.compare_exchange_strong(exp, (int)r); }
		}
		tmax[t] = std::max(tmax[t], mx);
	});
	if (noseq.load() >= 0) return failf(TDG_EFORMAT, "%s: entry %d has no sequence line", path, noseq.load());
	if (noqual.load() >= 0) return failf(TDG_EFORMAT, "%s: entry %d has no quality line", path, noqual.load());
	if (bad.load() >= 0) return failf(TDG_EFORMAT, "ERROR: Length of sequence and base qualities differ!.");  // io.c:1770
	for (int v : tmax) pc.max_len = std::max(pc.max_len, v);
	return TDG_OK;
}

extern "C" int tdg_fastq_next(tdg_fastq* f, int max_reads, int threads, tdg_fastq_chunk* chunk)
{
	if (!f || !chunk) return failf(TDG_EINVAL, "tdg_fastq_next: NULL argument");
	int rc;
	try {
		rc = split_lines(f, max_reads, f->own, threads);
		if (!rc) rc = convert_chunk(f->own, threads);
	} catch (const std::bad_alloc&) {
		rc = failf(TDG_EMEM, "out of host memory while reading %s", f->path.c_str());
	}
	memset(chunk, 0, sizeof *chunk);
	if (rc) return rc;
	const ParsedChunk& pc = f->own;
	chunk->n = pc.n; chunk->max_len = pc.max_len;
	if (pc.n) {
		chunk->len = pc.len.data(); chunk->seq_off = pc.seq_off.data(); chunk->codes = pc.codes.data();
		chunk->qual = pc.fasta ? nullptr : pc.qual.data();
		chunk->name_off = pc.name_off.data(); chunk->names = pc.names.data();
	}
	return TDG_OK;
}

// ------------------------------------------------------------------------------------------
// get_sequence_stats (io.c:52-300): raw sums over the first reads of a file
// ------------------------------------------------------------------------------------------
extern "C" int tdg_sequence_stats(const char* path, int fasta, int num_query, const uint8_t* five, int five_len,
                                  const uint8_t* three, int three_len, int threads, tdg_seq_stats* out)
{
	if (!path || !out || num_query < 1) return failf(TDG_EINVAL, "tdg_sequence_stats: bad argument");
	memset(out, 0, sizeof *out);
	tdg_fastq* f = nullptr;
	int rc = tdg_fastq_open(path, fasta, &f);
	if (rc) return rc;
	const int T = std::max(1, threads);
	struct Acc { int64_t n_len = 0, base[5] = {0, 0, 0, 0, 0}, f0 = 0, f1 = 0, f2 = 0, t0 = 0, t1 = 0, t2 = 0; int mx = 0; };
	ParsedChunk& pc = f->own;
	// the reference reads whole chunks of num_query reads until more than 1 000 000 have been seen (io.c:145-213)
	const int64_t target = ((int64_t)1000000 / num_query + 1) * (int64_t)num_query;
	while (out->total_read < target) {
		const int want = (int)std::min<int64_t>(target - out->total_read, (int64_t)1 << 20);
		if ((rc = split_lines(f, want, pc, T)) || (rc = convert_chunk(pc, T))) break;
		if (pc.n == 0) break;
		std::vector<Acc> acc((size_t)T);
		const uint8_t* end_of_codes = pc.codes.data() + pc.codes.size();
		parallel_for(T, (size_t)pc.n, 4096, [&](size_t b, size_t e, int t) {
			Acc a;
			for (size_t r = b; r < e; r++) {
				const uint8_t* seq = pc.codes.data() + pc.seq_off[r];
				const int len = pc.len[r];
				if (len > a.mx) a.mx = len;
				a.n_len += len;
				for (int j = 0; j < len; j++) if (seq[j] < 5) a.base[seq[j]]++;
				if (five_len) {  // longest exact match of a suffix of the 5' partial segment with the read start (io.c:160-175)
					for (int j = 0; j <= five_len; j++) {
						int c;
						for (c = 0; c < five_len - j; c++)
							if (seq + c >= end_of_codes || seq[c] != five[j + c]) break;
						if (c == five_len - j && c > 3) { a.f0++; a.f1 += five_len - j; a.f2 += (int64_t)(five_len - j) * (five_len - j); break; }
					}
				}
				if (three_len) {  // prefix of the 3' partial segment at the read end (io.c:177-193)
					for (int j = 0; j <= three_len; j++) {
						int c;
						for (c = 0; c < three_len - j; c++) {
							const int idx = len - (three_len - j - c);
							if (idx < 0 || seq[idx] != three[c]) break;
						}
						if (c == three_len - j && c > 3) { a.t0++; a.t1 += three_len - j; a.t2 += (int64_t)(three_len - j) * (three_len - j); break; }
					}
				}
			}
			acc[t] = a;
		});
		for (const Acc& a : acc) {
			out->max_seq_len = std::max(out->max_seq_len, a.mx);
			out->sum_len += (double)a.n_len;
			for (int k = 0; k < 5; k++) out->base_count[k] += (double)a.base[k];
			out->five_s0 += (double)a.f0; out->five_s1 += (double)a.f1; out->five_s2 += (double)a.f2;
			out->three_s0 += (double)a.t0; out->three_s1 += (double)a.t1; out->three_s2 += (double)a.t2;
		}
		out->total_read += pc.n;
	}
	tdg_fastq_close(f);
	return rc;
}

// ------------------------------------------------------------------------------------------
// formatting
// ------------------------------------------------------------------------------------------
// `%0.2f` of (double)mapq.  A float times 100 is exact in double (24 + 7 significant bits), so
// rint() under the default round-to-nearest-even mode gives the digits glibc prints.
extern "C" int tdg_format_rq(float mapq, char* out)
{
	const double v = (double)mapq;
	if (!std::isfinite(v) || std::fabs(v) >= 1e15) return sprintf(out, "%0.2f", v);
	char* p = out;
	if (std::signbit(v)) *p++ = '-';
	uint64_t n = (uint64_t)std::rint(std::fabs(v) * 100.0);
	const uint64_t ip = n / 100;
	const unsigned fr = (unsigned)(n % 100);
	char tmp[24];
	int k = 0;
	uint64_t x = ip;
	do { tmp[k++] = (char)('0' + x % 10); x /= 10; } while (x);
	while (k) *p++ = tmp[--k];
	*p++ = '.';
	*p++ = (char)('0' + fr / 10);
	*p++ = (char)('0' + fr % 10);
	*p = 0;
	return (int)(p - out);
}

namespace {

inline char* put_int(char* p, int v)
{
	unsigned u = (unsigned)v;
	if (v < 0) { *p++ = '-'; u = 0u - u; }
	char tmp[12];
	int k = 0;
	do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
	while (k) *p++ = tmp[--k];
	return p;
}

// print_all's characters (io.c:757: "ACGTNN"[code]) of n input characters: upper-case A, C, G, T and N stand for themselves
// -- sixteen at a time where a block holds nothing else -- everything else goes through the table.
inline char* put_bases(char* p, const uint8_t* s, int n)
{
	int q = 0;
#if defined(__SSE2__)
	const __m128i cA = _mm_set1_epi8('A'), cC = _mm_set1_epi8('C'), cG = _mm_set1_epi8('G'), cT = _mm_set1_epi8('T'), cN = _mm_set1_epi8('N');
	for (; q + 16 <= n; q += 16) {
		const __m128i x = _mm_loadu_si128((const __m128i*)(s + q));
		const __m128i ok = _mm_or_si128(_mm_or_si128(_mm_or_si128(_mm_cmpeq_epi8(x, cA), _mm_cmpeq_epi8(x, cC)), _mm_or_si128(_mm_cmpeq_epi8(x, cG), _mm_cmpeq_epi8(x, cT))),
		                                _mm_cmpeq_epi8(x, cN));
		if (_mm_movemask_epi8(ok) == 0xFFFF) _mm_storeu_si128((__m128i*)(p + q), x);
		else for (int k = 0; k < 16; k++) p[q + k] = kNuc.out[s[q + k]];
	}
#endif
	for (; q < n; q++) p[q] = kNuc.out[s[q]];
	return p + n;
}

// dust_sequences (barcode_hmm.c:2407-2467) on one read; `seq` is 0-terminated like ri->seq
bool dust_low_complexity(const uint8_t* seq, int rlen, int dust_cut)
{
	double triplet[64];
	for (int j = 0; j < 64; j++) triplet[j] = 0.0;
	int c = 0;
	while (seq[c] == 65) c++;
	int key = ((seq[c] & 0x3) << 2) | (seq[c + 1] & 0x3);
	int len = rlen;
	if (len > 64) len = 64;
	c += 2;
	for (int j = c; j < len; j++) {
		if (seq[j] == 65) break;
		key = key << 2 | (seq[j] & 0x3);
		triplet[key & 0x3F]++;
		c++;
	}
	double s = 0.0;
	for (int j = 0; j < 64; j++) s += triplet[j] * (triplet[j] - 1.0) / 2.0;
	s = s / (double)(c - 3) * 10.0;
	return s > dust_cut;
}

// the same on the characters of the input text (lean path): residue j is nuc_code[text[j]], 0 behind the read
bool dust_low_complexity_text(const uint8_t* s, int rlen, int dust_cut)
{
	uint8_t seq[72];
	const int m = rlen < 66 ? rlen : 66;
	for (int j = 0; j < m; j++) seq[j] = kNuc.code[s[j]];
	for (int j = m; j < 72; j++) seq[j] = 0;
	return dust_low_complexity(seq, rlen, dust_cut);
}

struct OutBuf {
	std::vector<char> d;
	size_t n = 0;
	char* grow(size_t need)
	{
		if (n + need > d.size()) d.resize(std::max(d.size() * 2, n + need + (1u << 16)));
		return d.data() + n;
	}
};

}  // namespace

// ------------------------------------------------------------------------------------------
// the pipeline
// ------------------------------------------------------------------------------------------
namespace {

struct Slot {
	std::vector<ParsedChunk> pc;          // per input
	std::vector<tdg_batch*> batch;        // per input (nullptr without model)
	std::vector<int> batch_reads, batch_len;
	std::vector<tdg_result> res;          // per input, valid after the GPU stage
	int n = 0;
	bool last = false;
};

template <class T>
struct Queue {
	std::mutex m; std::condition_variable cv; std::deque<T> q; bool closed = false;
	void push(T v) { { std::lock_guard<std::mutex> l(m); q.push_back(v); } cv.notify_all(); }
	void close() { { std::lock_guard<std::mutex> l(m); closed = true; } cv.notify_all(); }
	bool pop(T& v)
	{
		std::unique_lock<std::mutex> l(m);
		cv.wait(l, [&] { return !q.empty() || closed; });
		if (q.empty()) return false;
		v = q.front(); q.pop_front();
		return true;
	}
};

struct Shared {
	std::mutex m;
	int code = TDG_OK;
	std::string msg;
	std::atomic<bool> failed{false};
	void fail(int c, const char* s)
	{
		std::lock_guard<std::mutex> l(m);
		if (code == TDG_OK) { code = c; msg = s ? s : ""; }
		failed = true;
	}
};

}  // namespace

extern "C" int tdg_demux_run(tdg_context* ctx, const tdg_demux_job* job, tdg_demux_stats* stats)
{
	if (!job || !stats) return failf(TDG_EINVAL, "tdg_demux_run: NULL argument");
	memset(stats, 0, sizeof *stats);
	const int NI = job->n_inputs;
	if (NI < 1 || !job->inputs || !job->outfile) return failf(TDG_EINVAL, "tdg_demux_run: no inputs / no output prefix");
	bool any_model = false;
	for (int i = 0; i < NI; i++) any_model |= job->inputs[i].model != nullptr;
	if (any_model && !ctx) return failf(TDG_EINVAL, "tdg_demux_run: a GPU context is required");
	const int threads = std::max(1, job->threads);
	// default chunk: two waves (one with >= 4 devices) per device of the context; a chunk is sharded contiguously over the devices
	const int ndev = ctx ? std::max(1, tdg_device_count(ctx)) : 1;
	// with -ref a chunk is one chunk of the reference's own loop: its thread slices decide how each read is matched
	const int chunk_reads = job->refset ? (job->ref_chunk_reads > 0 ? job->ref_chunk_reads : 1000001)
	                                    : (job->chunk_reads > 0 ? job->chunk_reads : (ndev >= 4 ? 1 : 2) * 148 * 512 * ndev);
	// (one wave per device when there are many: the pinned staging of a slot is created at ~1.5 GB/s, and a chunk of
	//  1.2 M reads per slot would keep the pipeline waiting for its slots for most of a short job)
	if (job->refset && !ctx) return failf(TDG_EINVAL, "tdg_demux_run: a GPU context is required for the artifact filter");
	const int nalt = job->num_alternatives;
	if (nalt < 2) return failf(TDG_EINVAL, "num_alternatives must be >= 2");
	const double t_start = now_s();

	// ---- output files (print_all, io.c:822-915): all of them exist afterwards, even when empty
	int num_out_reads = 0;
	std::vector<int> base_file(NI, 0);
	for (int i = 0; i < NI; i++) { base_file[i] = nalt * num_out_reads; num_out_reads += job->inputs[i].num_read_segments; }
	const int num_outfiles = nalt * num_out_reads;
	if (num_outfiles == 0)
		return failf(TDG_EINVAL, "ERROR: No output files to create. Input sequences may not contain extractable reads or may not match the expected architecture.");
	// plain descriptors + pwrite at offsets kept here: every worker thread writes its own part of every file
	std::vector<int> files((size_t)num_outfiles, 0);     // 0 = no such file (fd + 1 otherwise)
	std::vector<int64_t> file_off((size_t)num_outfiles, 0);
	auto close_files = [&] { for (int f : files) if (f) close(f - 1); };
	{
		const bool bc = job->barcode_names != nullptr && job->barcode_input >= 0;
		char name[4096];
		int c = 0;
		for (int i = 0; i < num_out_reads; i++) {
			for (int j = 0; j < nalt; j++) {
				const bool un = (j == nalt - 1);
				if (bc) {
					if (un) { if (num_out_reads > 1) snprintf(name, sizeof name, "%s_un_READ%d.fq", job->outfile, i + 1); else snprintf(name, sizeof name, "%s_un.fq", job->outfile); }
					else if (num_out_reads > 1) snprintf(name, sizeof name, "%s_BC_%s_READ%d.fq", job->outfile, job->barcode_names[j], i + 1);
					else snprintf(name, sizeof name, "%s_BC_%s.fq", job->outfile, job->barcode_names[j]);
				} else {
					// two alternatives: the extracted file and the `un` file (io.c:889-914)
					if (un) { if (num_out_reads > 1) snprintf(name, sizeof name, "%s_un_READ%d.fq", job->outfile, i + 1); else snprintf(name, sizeof name, "%s_un.fq", job->outfile); }
					else if (j == 0) { if (num_out_reads > 1) snprintf(name, sizeof name, "%s_READ%d.fq", job->outfile, i + 1); else snprintf(name, sizeof name, "%s.fq", job->outfile); }
					else { c++; continue; }
				}
				const int fd = open(name, O_WRONLY | O_CREAT | O_TRUNC, 0666);
				if (fd < 0) { close_files(); return failf(TDG_EIO, "Failed to open file:%s", name); }
				files[c] = fd + 1;
				c++;
			}
		}
	}

	// ---- readers
	std::vector<tdg_fastq*> rd((size_t)NI, nullptr);
	auto close_readers = [&] { for (auto* r : rd) tdg_fastq_close(r); };
	for (int i = 0; i < NI; i++) {
		int rc = tdg_fastq_open(job->inputs[i].path, job->inputs[i].fasta, &rd[i]);
		if (rc) { close_readers(); close_files(); return rc; }
	}
	// per-model "is this HMM in an R segment" tables
	std::vector<std::vector<uint8_t>> is_read((size_t)NI);
	for (int i = 0; i < NI; i++)
		if (job->inputs[i].model) {
			is_read[i].resize((size_t)tdg_model_num_hmms(job->inputs[i].model));
			tdg_model_read_hmms(job->inputs[i].model, is_read[i].data());
		}

	// chunks in flight: one per stage (split, convert+pack, two on the GPU, write) so that no stage waits for a free slot
	// when the stages take about the same time (several devices); with one device the GPU stage dominates and 4 are enough
	const int NSLOT = ndev >= 4 ? 6 : 4;
	std::vector<Slot> slots((size_t)NSLOT);
	for (auto& s : slots) { s.pc.resize(NI); s.batch.assign(NI, nullptr); s.batch_reads.assign(NI, 0); s.batch_len.assign(NI, 0); s.res.resize(NI); }
	Queue<int> q_free, q_conv, q_gpu, q_write;
	for (int k = 0; k < NSLOT; k++) q_free.push(k);
	Shared sh;
	// Staging batches are slow to create (pinning the host pages, then buffers, a stream and events on every device:
	// 45-90 ms per batch on an 8-GPU box), so the slots' batches are made on set-up threads -- three at a time, slot k on
	// thread k mod 3 -- and every slot is released to the pipeline as soon as its own batch exists: chunk 0 is on the GPU
	// while the staging of later slots is still being pinned.  (A longer read later on re-creates the batch of that slot.)
	// The scratch arenas (tens of GB per device) are allocated meanwhile.
	std::vector<std::thread> t_alloc;
	std::mutex ready_mu;
	std::condition_variable ready_cv;
	std::vector<char> slot_ready((size_t)NSLOT, 0);
	bool scratch_ready = false;
	const int n_alloc = std::min(3, NSLOT);
	for (int a = 0; a < n_alloc; a++) t_alloc.emplace_back([&, a] {
		for (int k = a; k < NSLOT; k += n_alloc) {
			for (int i = 0; i < NI && !sh.failed; i++) {
				if (!((job->inputs[i].model || job->refset) && (job->inputs[i].expected_len > 0 || job->inputs[i].max_seq_len > 0))) continue;
				const double ta = now_s();
				Slot& s = slots[k];
				s.batch_reads[i] = std::min(chunk_reads, 1 << 24);
				s.batch_len[i] = job->inputs[i].expected_len > 0 ? job->inputs[i].expected_len : job->inputs[i].max_seq_len;
				// the label rows stay on the device (the writer works from the R-run spans): no pinned label staging
				if (tdg::batch_acquire(ctx, s.batch_reads[i], s.batch_len[i], &s.batch[i]) != TDG_OK ||
				    (job->inputs[i].model && tdg::batch_prepare(s.batch[i], job->inputs[i].model, false) != TDG_OK)) sh.fail(TDG_ECUDA, tdg_last_error());
				if (getenv("TDG_TRACE")) fprintf(stderr, "[trace] batch for slot %d input %d created in %.3f s\n", k, i, now_s() - ta);
			}
			{ std::lock_guard<std::mutex> l(ready_mu); slot_ready[k] = 1; }
			ready_cv.notify_all();
		}
	});
	t_alloc.emplace_back([&] {
		for (int i = 0; i < NI && !sh.failed; i++)
			if (job->inputs[i].model) {
				const double ta = now_s();
				if (tdg::scratch_prepare(ctx, job->inputs[i].model) != TDG_OK) sh.fail(TDG_ECUDA, tdg_last_error());
				if (getenv("TDG_TRACE")) fprintf(stderr, "[trace] scratch for input %d ready in %.3f s\n", i, now_s() - ta);
			}
		{ std::lock_guard<std::mutex> l(ready_mu); scratch_ready = true; }
		ready_cv.notify_all();
	});
	const bool trace = getenv("TDG_TRACE") != nullptr && getenv("TDG_TRACE")[0] != 0;
	auto tr = [&](const char* stage, int k, double t0) {
		if (trace) fprintf(stderr, "[trace] %-8s chunk-slot %d  %.3f -> %.3f s\n", stage, k, t0 - t_start, now_s() - t_start);
	};
	std::atomic<int64_t> long_events{0};
	double sec_split = 0, sec_parse = 0, sec_gpu = 0, sec_write = 0;
	// the workers all three host stages share; a stage thread works on its own loops too while it waits for them
	std::unique_ptr<WorkerPool> pool;
	if (threads > 1) pool.reset(new WorkerPool(threads));
	WorkerPool* const poolp = pool.get();

	// ---- stage 1a: line splitting (sequential per file)
	std::thread t_split([&] {
		tl_pool = poolp;
		try {
		std::vector<SplitPart> split_scratch;
		for (;;) {
			int k;
			if (!q_free.pop(k) || sh.failed) break;
			Slot& s = slots[k];
			const double t0 = now_s();
			bool ok = true;
			for (int i = 0; i < NI && ok; i++)
				if (split_lines(rd[i], chunk_reads, s.pc[i], threads, &split_scratch) != TDG_OK) { sh.fail(TDG_EFORMAT, tdg_last_error()); ok = false; }
			for (int i = 0; i + 1 < NI && ok; i++)
				for (int j = i + 1; j < NI && ok; j++)
					if (s.pc[i].n != s.pc[j].n) {  // barcode_hmm.c:258-268
						char b[1024];
						snprintf(b, sizeof b, "Input File:%s and %s differ in number of entries.", job->inputs[i].path, job->inputs[j].path);
						sh.fail(TDG_EFORMAT, b); ok = false;
					}
			sec_split += now_s() - t0;
			tr("split", k, t0);
			if (!ok) break;
			s.n = s.pc[0].n;
			s.last = (s.n == 0);
			q_conv.push(k);
			if (s.last) break;
		}
		} catch (const std::bad_alloc&) { sh.fail(TDG_EMEM, "out of host memory in the line-splitting stage"); }
		q_conv.close();
	});

	// ---- stage 1b: conversion + packing on the worker pool
	std::thread t_parse([&] {
		tl_pool = poolp;
		try {
		std::vector<int> run_max(NI);
		for (int i = 0; i < NI; i++) run_max[i] = job->inputs[i].max_seq_len;
		for (;;) {
			int k;
			if (!q_conv.pop(k) || sh.failed) break;
			Slot& s = slots[k];
			{ std::unique_lock<std::mutex> l(ready_mu); ready_cv.wait(l, [&] { return slot_ready[k] != 0; }); }
			if (sh.failed) break;
			const double t0 = now_s();
			bool ok = true;
			if (!s.last) {
				for (int i = 0; i < NI && ok; i++) {
					ParsedChunk& pc = s.pc[i];
					tdg_model* m = job->inputs[i].model;
					const bool to_gpu = m || job->refset;   // model-less files reach the GPU only for the artifact filter
					// the reads are packed into the staging batch by the pass that measures them (the bases are in cache then);
					// a chunk with a read longer than the batch was made for is packed again into a longer batch
					tdg::TextTarget tt;
					auto fresh_batch = [&](int min_len) -> bool {
						if (!s.batch[i] || s.batch_reads[i] < pc.n || s.batch_len[i] < min_len) {
							if (s.batch[i]) tdg_batch_destroy(s.batch[i]);
							s.batch[i] = nullptr;
							s.batch_reads[i] = std::max(s.batch_reads[i], std::max(pc.n, std::min(chunk_reads, 1 << 24)));
							s.batch_len[i] = std::max(s.batch_len[i], min_len);
							if (tdg::batch_acquire(ctx, s.batch_reads[i], s.batch_len[i], &s.batch[i]) != TDG_OK) { sh.fail(TDG_ECUDA, tdg_last_error()); return false; }
						}
						tdg_batch_clear(s.batch[i]);
						if (tdg::batch_text_target(s.batch[i], pc.n, &tt) != TDG_OK) { sh.fail(TDG_EINVAL, tdg_last_error()); return false; }
						return true;
					};
					if (to_gpu && !fresh_batch(s.batch_len[i])) { ok = false; break; }
					const uint8_t* code_of = kNuc.code;
					const bool nuc_std = tdg::is_nuc_code_table(kNuc.code);
					int rc;
					if (to_gpu) rc = lean_chunk(pc, threads, [&tt, code_of, nuc_std](size_t r, const uint8_t* seq, int len) {
						if (len <= tt.max_len) tdg::pack_text_read(tt, (int)r, seq, len, code_of, nuc_std);
					});
					else rc = lean_chunk(pc, threads, [](size_t, const uint8_t*, int) {});
					if (rc != TDG_OK) { sh.fail(TDG_EFORMAT, tdg_last_error()); ok = false; break; }
					tr("lengths", k, t0);
					// barcode_hmm.c:293-309: every read at least as long as the running maximum rebuilds the model
					int64_t ev = 0;
					int mx = run_max[i];
					for (int r = 0; r < pc.n; r++) if (pc.len[r] >= mx) { mx = pc.len[r]; ev++; }
					run_max[i] = mx;
					long_events += ev;
					if (!to_gpu) continue;
					int need = pc.max_len;
					if (job->matchstart != -1 || job->matchend != -1) need = std::max(need, job->matchend);
					if (m && need > tdg_model_max_len(m)) tdg_model_set_max_len(m, need + 10);
					if (pc.max_len > tt.max_len) {
						if (!fresh_batch(pc.max_len)) { ok = false; break; }
						const char* text = pc.text;
						parallel_for(threads, (size_t)pc.n, 2048, [&](size_t b, size_t e, int) {
							for (size_t r = b; r < e; r++) tdg::pack_text_read(tt, (int)r, (const uint8_t*)text + pc.seq_pos[r], pc.len[r], code_of, nuc_std);
						});
					}
					if (tdg::batch_text_commit(s.batch[i], pc.n) != TDG_OK) { sh.fail(TDG_EINVAL, tdg_last_error()); ok = false; break; }
				}
			}
			sec_parse += now_s() - t0;
			tr("convert", k, t0);
			if (!ok) break;
			q_gpu.push(k);
			if (s.last) break;
		}
		} catch (const std::bad_alloc&) { sh.fail(TDG_EMEM, "out of host memory in the conversion stage"); }
		q_gpu.close();
	});

	// ---- stage 2: GPU (submit of chunk k+1 is queued before the wait on chunk k)
	std::thread t_gpu([&] {
		int prev = -1;
		auto finish = [&](int k) -> bool {
			Slot& s = slots[k];
			const double t0 = now_s();
			for (int i = 0; i < NI; i++)
				if ((job->inputs[i].model || job->refset) && !s.last)
					if (tdg_wait(s.batch[i], &s.res[i]) != TDG_OK) { sh.fail(TDG_ECUDA, tdg_last_error()); return false; }
			sec_gpu += now_s() - t0;
			tr("gpu-wait", k, t0);
			q_write.push(k);
			return true;
		};
		for (;;) {
			int k;
			if (!q_gpu.pop(k)) break;
			if (sh.failed) break;
			{ std::unique_lock<std::mutex> l(ready_mu); ready_cv.wait(l, [&] { return scratch_ready; }); }
			if (sh.failed) break;
			Slot& s = slots[k];
			bool ok = true;
			const double ts = now_s();
			if (!s.last)
				for (int i = 0; i < NI && ok; i++) {
					tdg_model* m = job->inputs[i].model;
					if (!m && !job->refset) continue;
					tdg_run_params rp;
					memset(&rp, 0, sizeof rp);
					rp.confidence_threshold = job->inputs[i].confidence_threshold;
					rp.minlen = job->minlen; rp.matchstart = job->matchstart; rp.matchend = job->matchend;
					rp.dust = job->dust; rp.want_labels = 0; rp.want_spans = 1;
					rp.refset = job->refset; rp.filter_error = job->filter_error; rp.slice_threads = threads;
					// a file whose architecture is a single R segment: run_rna_dust (artifact filter + dust), no HMM
					if (tdg_submit(ctx, m, m ? TDG_MODE_GET_LABEL : TDG_MODE_RNA_DUST, &rp, s.batch[i]) != TDG_OK) { sh.fail(TDG_ECUDA, tdg_last_error()); ok = false; }
				}
			if (!ok) break;
			tr("submit", k, ts);
			if (prev >= 0 && !finish(prev)) { prev = -1; break; }
			prev = k;
		}
		if (prev >= 0 && !sh.failed) finish(prev);
		q_write.close();
	});

	// ---- stage 3: post-process + write
	std::thread t_write([&] {
		tl_pool = poolp;
		try {
		std::vector<std::vector<OutBuf>> ob((size_t)threads, std::vector<OutBuf>((size_t)num_outfiles));
		std::vector<std::vector<int64_t>> tally((size_t)threads, std::vector<int64_t>(8, 0));
		for (;;) {
			int k;
			if (!q_write.pop(k)) break;
			Slot& s = slots[k];
			if (s.last || sh.failed) { q_free.push(k); if (s.last) break; continue; }
			const double t0 = now_s();
			const int n = s.n;
			for (auto& v : ob) for (auto& b : v) b.n = 0;
			parallel_for(threads, (size_t)n, 1024, [&](size_t b, size_t e, int t) {
				std::vector<OutBuf>& out = ob[t];
				std::vector<int64_t>& tl = tally[t];
				std::vector<int32_t> rt((size_t)NI), fpv((size_t)NI);
				std::vector<float> mq((size_t)NI);
				std::vector<const uint16_t*> spv((size_t)NI);   // R-run spans of an extracted read, nullptr = the whole read
				std::vector<int> spn((size_t)NI);
				char fseq[260];
				for (size_t r = b; r < e; r++) {
					// per file: read_type / barcode / fingerprint / mapq and what make_extracted_read leaves of the read
					int merged = -100000, barcode = -1;
					for (int i = 0; i < NI; i++) {
						const ParsedChunk& pc = s.pc[i];
						spv[i] = nullptr; spn[i] = 0;
						if (job->inputs[i].model) {
							const tdg_result& R = s.res[i];
							rt[i] = R.read_type[r]; fpv[i] = R.fingerprint[r]; mq[i] = R.mapq[r];
							if (i == job->barcode_input) barcode = R.barcode[r];
							if (R.extracted[r]) { spv[i] = R.spans + r * (size_t)R.span_stride * 2; spn[i] = R.span_stride; }
						} else if (job->refset) {
							rt[i] = s.res[i].read_type[r]; fpv[i] = -1; mq[i] = -1.0f;  // run_rna_dust on the device (artifact filter, dust)
						} else {
							rt[i] = TDG_EXTRACT_SUCCESS; fpv[i] = -1; mq[i] = -1.0f;  // do_rna_dust + clear_read_info (io.c:2063-2093)
							if (job->dust && dust_low_complexity_text((const uint8_t*)pc.text + pc.seq_pos[r], pc.len[r], job->dust)) rt[i] = TDG_EXTRACT_FAIL_LOW_COMPLEXITY;
						}
						merged = std::max(merged, rt[i]);
					}
					switch (merged) {  // barcode_hmm.c:358-382
						case TDG_EXTRACT_SUCCESS: tl[0]++; break;
						case TDG_EXTRACT_FAIL_BAR_FINGER_NOT_FOUND: tl[1]++; break;
						case TDG_EXTRACT_FAIL_READ_TOO_SHORT: tl[2]++; break;
						case TDG_EXTRACT_FAIL_ARCHITECTURE_MISMATCH: tl[4]++; break;
						case TDG_EXTRACT_FAIL_MATCHES_ARTIFACTS: tl[5]++; tl[6]++; break;  // falls through in the reference
						case TDG_EXTRACT_FAIL_LOW_COMPLEXITY: tl[6]++; break;
						default:  // (sequence number << 8) | EXTRACT_FAIL_MATCHES_ARTIFACTS: counted per reference sequence (:379-382)
							tl[5]++;
							if (job->artifact_counts && (merged >> 8) >= 1) __atomic_fetch_add(&job->artifact_counts[(merged >> 8) - 1], (int64_t)1, __ATOMIC_RELAXED);
							break;
					}
					const int sel = (merged == TDG_EXTRACT_SUCCESS) ? (barcode != -1 ? (barcode & 0xFF) : 0) : nalt - 1;
					for (int i = 0; i < NI; i++) {
						if (!job->inputs[i].num_read_segments) continue;
						const ParsedChunk& pc = s.pc[i];
						const uint8_t* seq = (const uint8_t*)pc.text + pc.seq_pos[r];     // characters; codes through kNuc
						const uint8_t* ql = pc.fasta ? nullptr : (const uint8_t*)pc.text + pc.qual_pos[r];
						const int len = pc.len[r];
						const char* name = pc.text + pc.name_pos[r];
						const size_t name_len = pc.name_len[r];
						int f = base_file[i] + sel;
						// print_all (io.c:923-1001) walks the rewritten read: runs of bases (codes < 5) separated by spacers, one
						// output record per run, the run behind a spacer goes to the next READ file.  Here the spacers are never
						// written: a residue is "in" when it lies in an R-run span (extracted reads) and is a base.
						const uint16_t* sp = spv[i];
						const int nsp = sp ? spn[i] : 1;
						for (int k = 0; k < nsp; k++) {
							const int s0 = sp ? (int)sp[2 * k] : 0, sl = sp ? (int)sp[2 * k + 1] : len;
							if (sl == 0) break;
							const int s1 = std::min(len, s0 + sl);
							int g = s0;
							while (g < s1) {
								// the only character with a code >= 5 is '.' (nuc_code.c:52)
								while (g < s1 && seq[g] == '.') g++;
								if (g == s1) break;
								const void* dot = memchr(seq + g, '.', (size_t)(s1 - g));
								const int h = dot ? (int)((const uint8_t*)dot - seq) : s1;
								const bool more = h < len;  // something follows the run: the next run goes to the next READ file
								if (f >= 0 && f < num_outfiles && files[f]) {
									const int run = h - g;
									char* p = out[f].grow(name_len + 2 * (size_t)run + 320);
									char* p0 = p;
									*p++ = '@';
									memcpy(p, name, name_len); p += name_len;
									if (fpv[i] != -1) {
										memcpy(p, ";FP:", 4); p += 4;
										if (job->print_seq_finger) {  // get_finger_seq, io.c:1018-1029
											int key = fpv[i];
											const int fl = key & 0xFF;
											key >>= 8;
											for (int q = 0; q < fl; q++) { fseq[fl - q - 1] = "ACGTN"[key & 0x3]; key >>= 2; }
											memcpy(p, fseq, (size_t)fl); p += fl;
										} else p = put_int(p, fpv[i]);
									}
									memcpy(p, ";RQ:", 4); p += 4;
									p += tdg_format_rq(mq[i], p);
									*p++ = '\n';
									p = put_bases(p, seq + g, run);
									*p++ = '\n'; *p++ = '+'; *p++ = '\n';
									if (ql) { memcpy(p, ql + g, (size_t)run); p += run; }
									else { memset(p, '.', (size_t)run); p += run; }
									*p++ = '\n';
									out[f].n += (size_t)(p - p0);
								}
								if (more) f += nalt;
								g = h;
							}
						}
					}
				}
			});
			tr("format", k, t0);
			// Every file keeps input order: the part thread t formatted (reads [b_t, e_t) of the chunk) goes behind the parts
			// of the threads before it.  The offsets follow from the buffer sizes, so all threads write at once with
			// pwrite(), each its own (cache-warm) buffers, starting at a different file to stay off each other's inode locks.
			{
				std::vector<int64_t> pos((size_t)threads * num_outfiles);
				for (int f = 0; f < num_outfiles; f++) {
					int64_t off = file_off[f];
					for (int t = 0; t < threads; t++) { pos[(size_t)t * num_outfiles + f] = off; off += (int64_t)ob[t][f].n; }
					file_off[f] = off;
				}
				// one writer per file at a time (the kernel serialises writers of one inode anyway, and a queue of waiters on
				// its lock is slower than looking for another file): a thread walks its files starting at its own offset and
				// comes back later to the ones that were busy
				std::vector<std::atomic<char>> busy((size_t)num_outfiles);
				for (auto& b : busy) b.store(0);
				parallel_for(threads, (size_t)threads, 0, [&](size_t tb, size_t te, int) {
					for (size_t t = tb; t < te; t++) {
						const int f0 = (int)((t * (size_t)num_outfiles) / (size_t)threads);
						std::vector<int> todo;
						for (int q = 0; q < num_outfiles; q++) {
							const int f = (f0 + q) % num_outfiles;
							if (files[f] && ob[t][f].n) todo.push_back(f);
						}
						while (!todo.empty() && !sh.failed) {
							size_t kept = 0;
							for (size_t q = 0; q < todo.size(); q++) {
								const int f = todo[q];
								char expect = 0;
								if (!busy[f].compare_exchange_strong(expect, 1, std::memory_order_acquire)) { todo[kept++] = f; continue; }
								OutBuf& bf = ob[t][f];
								size_t done = 0;
								while (done < bf.n) {
									const ssize_t w = pwrite(files[f] - 1, bf.d.data() + done, bf.n - done, (off_t)(pos[t * num_outfiles + f] + (int64_t)done));
									if (w <= 0) { sh.fail(TDG_EIO, "write error on an output file"); break; }
									done += (size_t)w;
								}
								busy[f].store(0, std::memory_order_release);
							}
							if (kept == todo.size()) std::this_thread::yield();
							todo.resize(kept);
						}
					}
				});
			}
			stats->total_read += n;
			sec_write += now_s() - t0;
			tr("write", k, t0);
			q_free.push(k);
		}
		for (auto& tl : tally) {
			stats->num_EXTRACT_SUCCESS += tl[0];
			stats->num_EXTRACT_FAIL_BAR_FINGER_NOT_FOUND += tl[1];
			stats->num_EXTRACT_FAIL_READ_TOO_SHORT += tl[2];
			stats->num_EXTRACT_FAIL_ARCHITECTURE_MISMATCH += tl[4];
			stats->num_EXTRACT_FAIL_MATCHES_ARTIFACTS += tl[5];
			stats->num_EXTRACT_FAIL_LOW_COMPLEXITY += tl[6];
		}
		} catch (const std::bad_alloc&) { sh.fail(TDG_EMEM, "out of host memory in the writer stage"); }
		q_free.close();
	});

	t_split.join();
	for (auto& t : t_alloc) t.join();
	if (sh.failed) { q_conv.close(); q_free.close(); }
	t_parse.join();
	if (sh.failed) { q_gpu.close(); q_free.close(); }
	t_gpu.join();
	if (sh.failed) { q_write.close(); q_free.close(); }
	t_write.join();
	double tq = now_s();
	{
		// back to the context's pool (released by tdg_shutdown): freeing pinned staging costs ~0.3 s per batch
		for (auto& s : slots) for (auto* b : s.batch) if (b) tdg::batch_release(b);
	}
	tr("free-batches", -1, tq); tq = now_s();
	close_readers();
	tr("close-in", -1, tq); tq = now_s();
	close_files();
	tr("close-out", -1, tq);
	stats->long_sequence_events = long_events.load();
	stats->seconds_split = sec_split; stats->seconds_parse = sec_parse; stats->seconds_gpu_wait = sec_gpu; stats->seconds_write = sec_write;
	stats->seconds_total = now_s() - t_start;
	if (sh.failed) return tdg::set_last_error(sh.code, sh.msg.c_str());
	return TDG_OK;
}
