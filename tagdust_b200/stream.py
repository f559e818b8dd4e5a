"""ctypes front end of include/tagdust_b200_stream.h: the FASTQ reader and the streaming
demultiplexer (read_fasta_fastq io.c:1684, print_all io.c:757, the chunk loop of
hmm_controller_multiple barcode_hmm.c:243-384).  Used by tests and bench.py; the reference-side
binding is integration/controller_gpu.c."""
import ctypes as C

import numpy as np

from . import _capi
from .api import TagdustError, _check


class FastqReader:
    def __init__(self, path, fasta=-1):
        self.lib = _capi.load_library()
        h = C.c_void_p()
        _check(self.lib, self.lib.tdg_fastq_open(str(path).encode(), fasta, C.byref(h)))
        self.h = h

    def next(self, max_reads, threads=1):
        """Returns None at end of input, else dict(len, codes (list of arrays), qual, names)."""
        ch = _capi.FastqChunkC()
        _check(self.lib, self.lib.tdg_fastq_next(self.h, max_reads, threads, C.byref(ch)))
        n = ch.n
        if n == 0:
            return None
        lens = np.ctypeslib.as_array(ch.len, shape=(n,)).copy()
        off = np.ctypeslib.as_array(ch.seq_off, shape=(n,)).copy()
        total = int(off[-1] + lens[-1] + 1)
        codes = np.ctypeslib.as_array(ch.codes, shape=(total,)).copy()
        qual = np.ctypeslib.as_array(ch.qual, shape=(total,)).copy() if ch.qual else None
        noff = np.ctypeslib.as_array(ch.name_off, shape=(n + 1,)).copy()
        raw = C.string_at(ch.names, int(noff[-1]))
        names = [raw[noff[r]:noff[r + 1]].split(b"\0", 1)[0] for r in range(n)]
        return dict(n=n, max_len=ch.max_len, len=lens, off=off, codes=codes, qual=qual, names=names)

    def close(self):
        if self.h:
            self.lib.tdg_fastq_close(self.h)
            self.h = None


def format_rq(v):
    lib = _capi.load_library()
    buf = C.create_string_buffer(64)
    n = lib.tdg_format_rq(C.c_float(v), buf)
    return buf.raw[:n].decode()


def demux_run(ctx, inputs, outfile, *, barcode_input=-1, barcode_names=None, minlen=16, dust=100, matchstart=-1,
              matchend=-1, print_seq_finger=0, threads=8, chunk_reads=0, refset=None, filter_error=2, ref_chunk_reads=0,
              artifact_counts=None):
    """inputs: list of dict(path, model (api.Model or None), num_read_segments, threshold, max_seq_len, fasta=-1).
    barcode_names: sequences of the first B segment without the trailing N alternative."""
    lib = _capi.load_library()
    arr = (_capi.DemuxInputC * len(inputs))()
    for k, it in enumerate(inputs):
        arr[k].path = str(it["path"]).encode()
        arr[k].fasta = it.get("fasta", -1)
        arr[k].model = it["model"].h if it.get("model") is not None else None
        arr[k].num_read_segments = it.get("num_read_segments", 1)
        arr[k].confidence_threshold = it.get("threshold", 0.0)
        arr[k].max_seq_len = it.get("max_seq_len", 0)
        arr[k].expected_len = it.get("expected_len", 0)
    job = _capi.DemuxJobC()
    job.n_inputs = len(inputs)
    job.inputs = arr
    job.barcode_input = barcode_input
    if barcode_names:
        names = (C.c_char_p * len(barcode_names))(*[b.encode() for b in barcode_names])
        job.barcode_names = names
        job.num_alternatives = len(barcode_names) + 1
    else:
        job.barcode_names = None
        job.num_alternatives = 2
    job.outfile = str(outfile).encode()
    job.minlen, job.dust, job.matchstart, job.matchend = minlen, dust, matchstart, matchend
    job.print_seq_finger, job.threads, job.chunk_reads = print_seq_finger, threads, chunk_reads
    if refset is not None:
        job.refset = refset.h
        job.filter_error, job.ref_chunk_reads = filter_error, ref_chunk_reads
        if artifact_counts is not None:   # int64 numpy array, one counter per reference sequence
            job.artifact_counts = artifact_counts.ctypes.data_as(C.POINTER(C.c_int64))
    st = _capi.DemuxStatsC()
    _check(lib, lib.tdg_demux_run(ctx.h if ctx is not None else None, C.byref(job), C.byref(st)))
    return {k: getattr(st, k) for k, _ in st._fields_}
