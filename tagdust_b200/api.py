"""Host-side mirror of the reference's run_pHMM() seam (barcode_hmm.c:1895) over the C ABI.

    ctx   = Context()                       # tdg_init: all visible B200s
    model = ctx.model(desc, max_len)        # flattened struct model_bag -> device
    batch = ctx.batch(max_reads, max_len)   # pinned, 4-bit packed, double-bufferable
    batch.append(codes, lens)
    res = ctx.run_phmm(model, batch, MODE_GET_LABEL, threshold=..., minlen=...)

Names follow the reference (mapq, bar_prob, labels, read_type, barcode, fingerprint).
Everything computes on the GPU through libtagdust_b200.so; a missing library or device
raises -- there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import (MODE_ARCH_COMP, MODE_GET_LABEL, MODE_GET_PROB, MODE_RNA_DUST, ArchParamsC, ModelDesc, ResultC,
                    RunParamsC)


class TagdustError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tagdust_b200 error {code}: {msg}")
        self.code = code


def _check(lib, rc):
    if rc != 0:
        raise TagdustError(rc, lib.tdg_last_error().decode(errors="replace"))


def compile_architecture(segments, background_logp, average_length, max_seq_len, *, e=0.05, i=0.1,
                         five=(0.0, 0.0, 0.0), three=(0.0, -1.0, -1.0), calibration_edit=False) -> ModelDesc:
    """tdg_arch_compile: segment strings -> flattened model (host only, no GPU needed).
    Mirrors init_model_bag (barcode_hmm.c:5760-6011)."""
    lib = _capi.load_library()
    ap = ArchParamsC()
    for k in range(5):
        ap.background_logp[k] = float(background_logp[k])
    ap.average_length = float(average_length)
    ap.max_seq_len = int(max_seq_len)
    ap.expected_5_len, ap.mean_5_len, ap.stdev_5_len = [float(x) for x in five]
    ap.expected_3_len, ap.mean_3_len, ap.stdev_3_len = [float(x) for x in three]
    ap.sequencer_error_rate = e
    ap.indel_frequency = i
    ap.calibration_edit = 1 if calibration_edit else 0
    arr = (C.c_char_p * len(segments))(*[s.encode() for s in segments])
    h = C.c_void_p()
    _check(lib, lib.tdg_arch_compile(len(segments), arr, C.byref(ap), C.byref(h)))
    try:
        return ModelDesc.from_c(lib.tdg_arch_desc(h).contents)
    finally:
        lib.tdg_arch_destroy(h)


def live_ops(desc: ModelDesc):
    """tdg_desc_live_ops: logsums / adds the kernels execute per read position over all HMMs
    (dead log(0) terms excluded) -> dict(ls_bwd, add_bwd, ls_fwd, add_fwd)."""
    lib = _capi.load_library()
    out = (C.c_double * 4)()
    _check(lib, lib.tdg_desc_live_ops(C.byref(desc.c), out))
    return dict(ls_bwd=out[0], add_bwd=out[1], ls_fwd=out[2], add_fwd=out[3])


class Model:
    def __init__(self, ctx, desc: ModelDesc, max_len: int):
        self.ctx, self.desc, self.max_len = ctx, desc, int(max_len)
        self.h = C.c_void_p()
        _check(ctx.lib, ctx.lib.tdg_model_create(ctx.h, C.byref(desc.c), self.max_len, C.byref(self.h)))

    def close(self):
        if self.h:
            self.ctx.lib.tdg_model_destroy(self.h)
            self.h = C.c_void_p()


class RefSet:
    """tdg_refset: the -ref artifact sequences (struct fasta) on the devices of a context."""

    def __init__(self, ctx, codes, s_index):
        self.ctx = ctx
        codes = np.ascontiguousarray(codes, np.uint8)
        s_index = np.ascontiguousarray(s_index, np.int32)
        self.h = C.c_void_p()
        _check(ctx.lib, ctx.lib.tdg_refset_create(ctx.h, codes.ctypes.data_as(_capi.c_uint8_p), s_index.ctypes.data_as(_capi.c_int32_p),
                                                  len(s_index) - 1, C.byref(self.h)))

    def close(self):
        if self.h:
            self.ctx.lib.tdg_refset_destroy(self.h)
            self.h = C.c_void_p()


class Batch:
    def __init__(self, ctx, max_reads: int, max_len: int):
        self.ctx, self.max_reads, self.max_len = ctx, int(max_reads), int(max_len)
        self.h = C.c_void_p()
        _check(ctx.lib, ctx.lib.tdg_batch_create(ctx.h, self.max_reads, self.max_len, C.byref(self.h)))
        self._lens = []

    def clear(self):
        _check(self.ctx.lib, self.ctx.lib.tdg_batch_clear(self.h))
        self._lens = []

    def append(self, codes: np.ndarray, lens: np.ndarray):
        """codes[n, stride] uint8 (0..4 per base, codes[r, len] = terminator), lens[n] int32."""
        codes = np.ascontiguousarray(codes, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        n, stride = codes.shape
        _check(self.ctx.lib, self.ctx.lib.tdg_batch_append_codes(self.h, n, codes.ctypes.data, stride, lens.ctypes.data))
        self._lens.append(lens.copy())

    @property
    def size(self):
        return self.ctx.lib.tdg_batch_size(self.h)

    def lens(self):
        return np.concatenate(self._lens) if self._lens else np.zeros(0, np.int32)

    def close(self):
        if self.h:
            self.ctx.lib.tdg_batch_destroy(self.h)
            self.h = C.c_void_p()


def _result_to_numpy(res: ResultC, mode, want_labels, want_spans=False):
    n = res.n_reads

    def arr(ptr, dt, count=n):
        if count == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(ptr, shape=(count,)).astype(dt, copy=True)

    out = {"b_score": arr(res.b_score, np.float32)}
    if mode == MODE_ARCH_COMP:
        return out
    out.update(mapq=arr(res.mapq, np.float32), bar_prob=arr(res.bar_prob, np.float32),
               f_score=arr(res.f_score, np.float32), r_score=arr(res.r_score, np.float32))
    if want_labels:
        out["labels"] = arr(res.labels, np.uint8, n * res.label_stride).reshape(n, res.label_stride)
    if want_spans and mode == MODE_GET_LABEL:
        out["spans"] = arr(res.spans, np.uint16, n * res.span_stride * 2).reshape(n, res.span_stride, 2)
    if mode == MODE_GET_LABEL:
        out.update(read_type=arr(res.read_type, np.int32), barcode=arr(res.barcode, np.int32),
                   fingerprint=arr(res.fingerprint, np.int32), extracted=arr(res.extracted, np.uint8))
    return out


class Context:
    def __init__(self, n_devices=0, device_ids=None):
        self.lib = _capi.load_library()
        self.h = C.c_void_p()
        ids = None
        if device_ids is not None:
            ids = (C.c_int32 * len(device_ids))(*device_ids)
            n_devices = len(device_ids)
        _check(self.lib, self.lib.tdg_init(n_devices, ids, C.byref(self.h)))

    @property
    def device_count(self):
        return self.lib.tdg_device_count(self.h)

    def model(self, desc, max_len):
        return Model(self, desc, max_len)

    def batch(self, max_reads, max_len):
        return Batch(self, max_reads, max_len)

    @staticmethod
    def _params(threshold, minlen, matchstart, matchend, dust, want_labels, want_spans=False, refset=None, filter_error=2,
                slice_threads=1):
        return RunParamsC(float(threshold), int(minlen), int(matchstart), int(matchend), int(dust), int(want_labels),
                          int(want_spans), refset.h if refset is not None else None, int(filter_error), int(slice_threads))

    def refset(self, codes, s_index):
        return RefSet(self, codes, s_index)

    def submit(self, model, batch, mode, *, threshold=0.0, minlen=16, matchstart=-1, matchend=-1, dust=100,
               want_labels=True, want_spans=False, refset=None, filter_error=2, slice_threads=1):
        rp = self._params(threshold, minlen, matchstart, matchend, dust, want_labels, want_spans, refset, filter_error, slice_threads)
        _check(self.lib, self.lib.tdg_submit(self.h, model.h if model is not None else None, mode, C.byref(rp), batch.h))
        batch._pending = (mode, bool(want_labels) and mode != MODE_RNA_DUST, bool(want_spans))

    def rna_dust(self, batch, *, dust=100, refset=None, filter_error=2, slice_threads=1):
        """run_rna_dust() (barcode_hmm.c:2043): reads of a file whose architecture is a single R segment -> read_type."""
        self.submit(None, batch, MODE_RNA_DUST, dust=dust, want_labels=False, refset=refset, filter_error=filter_error,
                    slice_threads=slice_threads)
        res = ResultC()
        _check(self.lib, self.lib.tdg_wait(batch.h, C.byref(res)))
        return np.ctypeslib.as_array(res.read_type, shape=(res.n_reads,)).astype(np.int32, copy=True)

    def wait(self, batch, copy=True):
        res = ResultC()
        _check(self.lib, self.lib.tdg_wait(batch.h, C.byref(res)))
        mode, wl, ws = batch._pending
        return _result_to_numpy(res, mode, wl, ws) if copy else res

    def run_phmm(self, model, batch, mode, **kw):
        """run_pHMM(ab=0, mb, ri, param, 0, numseq, mode) for MODE_GET_LABEL / MODE_GET_PROB."""
        self.submit(model, batch, mode, **kw)
        return self.wait(batch)

    def arch_compare(self, models, batch, num_threads=1):
        """run_pHMM(ab, ..., MODE_ARCH_COMP): returns (b_scores[A, n], arch_posterior[A])."""
        A, n = len(models), batch.size
        arr = (C.c_void_p * A)(*[m.h for m in models])
        bs = np.zeros((A, n), np.float32)
        post = np.zeros(A, np.float32)
        _check(self.lib, self.lib.tdg_arch_compare(self.h, arr, A, batch.h, num_threads,
                                                   bs.ctypes.data_as(_capi.c_float_p), post.ctypes.data_as(_capi.c_float_p)))
        return bs, post

    # device-resident path (bench `value`)
    def upload(self, batch):
        _check(self.lib, self.lib.tdg_batch_upload(self.h, batch.h))

    def decode_resident(self, model, batch, mode, stream=0, **kw):
        rp = self._params(kw.get("threshold", 0.0), kw.get("minlen", 16), kw.get("matchstart", -1),
                          kw.get("matchend", -1), kw.get("dust", 100), kw.get("want_labels", True),
                          kw.get("want_spans", False))
        batch._resident_spans = bool(kw.get("want_spans", False))
        nl = C.c_int(0)
        _check(self.lib, self.lib.tdg_decode_resident(self.h, model.h, mode, C.byref(rp), batch.h,
                                                      C.c_void_p(stream), C.byref(nl)))
        return nl.value

    def download(self, batch, mode=MODE_GET_LABEL):
        res = ResultC()
        _check(self.lib, self.lib.tdg_batch_download(batch.h, C.byref(res)))
        return _result_to_numpy(res, mode, True, getattr(batch, "_resident_spans", False))

    def profile_enable(self, on=True):
        _check(self.lib, self.lib.tdg_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, device_index=0):
        ms = (C.c_float * 3)()
        nl = (C.c_int * 3)()
        _check(self.lib, self.lib.tdg_profile_read(self.h, device_index, ms, nl))
        names = ("k_backward", "k_forward", "k_label")
        return {names[k]: {"ms": float(ms[k]), "launches": int(nl[k])} for k in range(3)}

    def cells(self, model, batch):
        return self.lib.tdg_batch_cells(model.h, batch.h)

    def close(self):
        if self.h:
            self.lib.tdg_shutdown(self.h)
            self.h = C.c_void_p()
