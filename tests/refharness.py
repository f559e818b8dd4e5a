"""TEST INFRASTRUCTURE: ctypes access to the oracle.

* RefHarness  -> oracle/_ref/libtagdust_ref.so  (the UNMODIFIED reference + oracle/ref_harness.c)
* Oracle      -> oracle/liboracle.so            (plain-C restatement, oracle/oracle_hmm.c)

Only tests/, __graft_entry__.smoke() and bench.py's cpu baseline legs import this.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

from tagdust_b200._capi import ModelDesc, ModelDescC, RunParamsC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libtagdust_ref.so")
REF_RTEST_SO = os.path.join(ORACLE_DIR, "_ref", "libtagdust_ref_rtest.so")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)


def build_oracle():
    """(Re)build oracle/liboracle.so and, when /root/reference is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "all"], check=True)


def have_ref():
    return os.path.exists(REF_SO)


def _p(a, t):
    return a.ctypes.data_as(t)


def background_logp(counts=(1.0, 1.0, 1.0, 1.0, 1.0)):
    """ssi->background as get_sequence_stats leaves it (io.c:79-81,263-270):
    prob2scaledprob(count/sum) -- a float function result stored in a double."""
    counts = np.asarray(counts, dtype=np.float64)
    s = counts.sum()
    return np.array([float(np.float32(math.log(float(np.float32(c / s))))) for c in counts], dtype=np.float64)


class RefHarness:
    def __init__(self, rtest=False):
        path = REF_RTEST_SO if rtest else REF_SO
        if not os.path.exists(path):
            raise OSError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(path)
        self.L = L
        L.refh_init()
        L.refh_logsum.restype = C.c_float
        L.refh_logsum.argtypes = [C.c_float, C.c_float]
        L.refh_param_new.restype = C.c_void_p
        L.refh_param_new.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_float, C.c_float, C.c_int, C.c_int,
                                     C.c_float, C.c_int, C.c_int, C.c_int]
        L.refh_param_set.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int]
        L.refh_param_free.argtypes = [C.c_void_p]
        L.refh_model_new.restype = C.c_void_p
        L.refh_model_new.argtypes = [C.c_void_p, f64p, C.c_double, C.c_int] + [C.c_double] * 6
        L.refh_model_calibration_edit.argtypes = [C.c_void_p, C.c_void_p]
        L.refh_model_free.argtypes = [C.c_void_p]
        L.refh_model_dims.argtypes = [C.c_void_p] + [i32p] * 5
        L.refh_model_flatten.argtypes = [C.c_void_p, i32p, i32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, i32p, f32p]
        L.refh_param_segment_info.argtypes = [C.c_void_p, C.c_int, C.c_char_p, i32p, i32p]
        L.refh_decode_scores.argtypes = [C.c_void_p, C.c_int, u8p, C.c_int, i32p, f32p, f32p, f32p, f64p, u8p]
        L.refh_backward_scores.argtypes = [C.c_void_p, C.c_int, u8p, C.c_int, i32p, f32p]
        L.refh_decode_matrix.argtypes = [C.c_void_p, u8p, C.c_int, f32p, i32p]
        L.refh_run_phmm.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, u8p, C.c_int, i32p,
                                    f32p, f64p, u8p, i32p, i32p, i32p, u8p, u8p, i32p]
        L.refh_run_arch_comp.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, u8p, C.c_int, i32p, f32p]
        L.refh_emit.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint, u8p, C.c_int, i32p]
        L.refh_logsum_table.argtypes = [f32p]

    # -- numerics
    def logsum(self, a, b):
        return self.L.refh_logsum(a, b)

    def logsum_table(self):
        t = np.zeros(16000, dtype=np.float32)
        self.L.refh_logsum_table(_p(t, f32p))
        return t

    # -- parameters / model
    def param_new(self, segments, e=0.05, i=0.1, minlen=16, threads=1, threshold=0.0, dust=100,
                  matchstart=-1, matchend=-1):
        arr = (C.c_char_p * len(segments))(*[s.encode() for s in segments])
        p = self.L.refh_param_new(len(segments), arr, e, i, minlen, threads, threshold, dust, matchstart, matchend)
        if not p:
            raise ValueError(f"reference rejected architecture {segments}")
        return p

    def param_set(self, param, threshold, minlen=16, dust=100, threads=1):
        self.L.refh_param_set(param, threshold, minlen, dust, threads)

    def param_free(self, p):
        self.L.refh_param_free(p)

    def model_new(self, param, background=None, average_length=150.0, max_seq_len=150,
                  five=(0.0, 0.0, 0.0), three=(0.0, -1.0, -1.0)):
        bg = background_logp() if background is None else np.asarray(background, dtype=np.float64)
        mb = self.L.refh_model_new(param, _p(bg, f64p), average_length, max_seq_len,
                                   five[0], five[1], five[2], three[0], three[1], three[2])
        if not mb:
            raise RuntimeError("init_model_bag failed")
        return mb

    def model_calibration_edit(self, mb, param):
        self.L.refh_model_calibration_edit(mb, param)

    def model_free(self, mb):
        self.L.refh_model_free(mb)

    def flatten(self, mb, param) -> ModelDesc:
        d = [C.c_int32() for _ in range(5)]
        self.L.refh_model_dims(mb, *[C.byref(x) for x in d])
        S, H, Cn, avg, _dyn = [x.value for x in d]
        seg_num_hmms = np.zeros(S, np.int32)
        seg_num_cols = np.zeros(S, np.int32)
        seg_skip = np.zeros(S, np.float32)
        background = np.zeros(5, np.float32)
        transition = np.zeros((Cn, 9), np.float32)
        m_emit = np.zeros((Cn, 5), np.float32)
        i_emit = np.zeros((Cn, 5), np.float32)
        sM = np.zeros(Cn, np.float32)
        sI = np.zeros(Cn, np.float32)
        label = np.zeros(H, np.int32)
        T = np.zeros((H, H), np.float32)
        self.L.refh_model_flatten(mb, _p(seg_num_hmms, i32p), _p(seg_num_cols, i32p), _p(seg_skip, f32p),
                                  _p(background, f32p), _p(transition, f32p), _p(m_emit, f32p), _p(i_emit, f32p),
                                  _p(sM, f32p), _p(sI, f32p), _p(label, i32p), _p(T, f32p))
        types = b""
        for s in range(S):
            t = C.create_string_buffer(2)
            a, b = C.c_int32(), C.c_int32()
            self.L.refh_param_segment_info(param, s, t, C.byref(a), C.byref(b))
            types += t.raw[:1]
        return ModelDesc(types, seg_num_hmms, seg_num_cols, seg_skip, background, transition, m_emit, i_emit,
                         sM, sI, label, T, avg)

    # -- per-read
    def decode_scores(self, mb, codes, lens, want_labels=True):
        codes = np.ascontiguousarray(codes, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        n, stride = codes.shape
        f = np.zeros(n, np.float32); b = np.zeros(n, np.float32); r = np.zeros(n, np.float32)
        bp = np.zeros(n, np.float64)
        labels = np.zeros((n, stride), np.uint8)
        self.L.refh_decode_scores(mb, n, _p(codes, u8p), stride, _p(lens, i32p), _p(f, f32p), _p(b, f32p),
                                  _p(r, f32p), _p(bp, f64p), _p(labels, u8p) if want_labels else None)
        return dict(f_score=f, b_score=b, r_score=r, bar_prob=bp, labels=labels)

    def backward_scores(self, mb, codes, lens):
        codes = np.ascontiguousarray(codes, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        n, stride = codes.shape
        b = np.zeros(n, np.float32)
        self.L.refh_backward_scores(mb, n, _p(codes, u8p), stride, _p(lens, i32p), _p(b, f32p))
        return b

    def run_phmm(self, mb, param, mode, codes, lens):
        codes = np.ascontiguousarray(codes, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        n, stride = codes.shape
        out = dict(mapq=np.zeros(n, np.float32), bar_prob=np.zeros(n, np.float64),
                   labels=np.zeros((n, stride), np.uint8), read_type=np.zeros(n, np.int32),
                   barcode=np.zeros(n, np.int32), fingerprint=np.zeros(n, np.int32),
                   seq=np.zeros((n, stride), np.uint8), qual=np.zeros((n, stride), np.uint8),
                   len=np.zeros(n, np.int32))
        st = self.L.refh_run_phmm(mb, param, mode, n, _p(codes, u8p), stride, _p(lens, i32p),
                                  _p(out["mapq"], f32p), _p(out["bar_prob"], f64p), _p(out["labels"], u8p),
                                  _p(out["read_type"], i32p), _p(out["barcode"], i32p), _p(out["fingerprint"], i32p),
                                  _p(out["seq"], u8p), _p(out["qual"], u8p), _p(out["len"], i32p))
        assert st == 0
        return out

    # -- -ref artifact filter (match_to_reference, barcode_hmm.c:2478-2583)
    def set_reference(self, param, ref_codes, s_index, filter_error=2):
        """ref_codes: uint8 nuc codes of all reference sequences back to back; s_index[numseq+1]; empty = remove."""
        self.L.refh_set_reference.argtypes = [C.c_void_p, u8p, i32p, C.c_int, C.c_int]
        rc = np.ascontiguousarray(ref_codes, np.uint8)
        si = np.ascontiguousarray(s_index, np.int32)
        self.L.refh_set_reference(param, _p(rc, u8p), _p(si, i32p), len(si) - 1 if len(si) else 0, filter_error)

    def run_phmm_ref(self, mb, param, codes, lens):
        self.L.refh_run_phmm_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_int, u8p, C.c_int, i32p, f32p, i32p, i32p, i32p, u8p]
        codes = np.ascontiguousarray(codes, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        n, stride = codes.shape
        out = dict(mapq=np.zeros(n, np.float32), read_type=np.zeros(n, np.int32), barcode=np.zeros(n, np.int32),
                   fingerprint=np.zeros(n, np.int32), seq=np.zeros((n, stride), np.uint8))
        st = self.L.refh_run_phmm_ref(mb, param, n, _p(codes, u8p), stride, _p(lens, i32p), _p(out["mapq"], f32p),
                                      _p(out["read_type"], i32p), _p(out["barcode"], i32p), _p(out["fingerprint"], i32p),
                                      _p(out["seq"], u8p))
        assert st == 0
        return out

    def run_rna_dust(self, param, codes, lens):
        self.L.refh_run_rna_dust.argtypes = [C.c_void_p, C.c_int, u8p, C.c_int, i32p, i32p]
        codes = np.ascontiguousarray(codes, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        n, stride = codes.shape
        rt = np.zeros(n, np.int32)
        assert self.L.refh_run_rna_dust(param, n, _p(codes, u8p), stride, _p(lens, i32p), _p(rt, i32p)) == 0
        return rt

    def myers(self, which, t, p):
        fn = self.L.refh_bmp_single if which == "bmp_single" else self.L.refh_bpm_check_error
        fn.argtypes = [u8p, u8p, C.c_int, C.c_int]
        t = np.ascontiguousarray(t, np.uint8); p = np.ascontiguousarray(p, np.uint8)
        return fn(_p(t, u8p), _p(p, u8p), len(t), len(p))

    def run_arch_comp(self, mbs, param, codes, lens):
        codes = np.ascontiguousarray(codes, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        n, stride = codes.shape
        arr = (C.c_void_p * len(mbs))(*mbs)
        post = np.zeros(len(mbs), np.float32)
        st = self.L.refh_run_arch_comp(arr, len(mbs), param, n, _p(codes, u8p), stride, _p(lens, i32p), _p(post, f32p))
        assert st == 0
        return post

    def emit(self, mb, n_model, n_random, average_length, seed, stride):
        n = n_model + n_random
        codes = np.zeros((n, stride), np.uint8)
        lens = np.zeros(n, np.int32)
        st = self.L.refh_emit(mb, n_model, n_random, average_length, seed, _p(codes, u8p), stride, _p(lens, i32p))
        if st != 0:
            raise RuntimeError(f"refh_emit failed {st}")
        return codes, lens


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        L = C.CDLL(ORACLE_SO)
        self.L = L
        L.orc_init_logsum()
        L.orc_logsum.restype = C.c_float
        L.orc_logsum.argtypes = [C.c_float, C.c_float]
        L.orc_logsum_table.argtypes = [f32p]
        L.orc_backward_score.restype = C.c_float
        L.orc_backward_score.argtypes = [C.POINTER(ModelDescC), u8p, C.c_int]
        L.orc_run.argtypes = [C.POINTER(ModelDescC), C.POINTER(RunParamsC), C.c_int, C.c_int, u8p, C.c_size_t, i32p,
                              C.c_int, f32p, f32p, f32p, f32p, f32p, i32p, i32p, i32p, u8p, u8p, i32p]
        L.orc_arch_compare.argtypes = [C.POINTER(C.POINTER(ModelDescC)), C.c_int, C.c_int, u8p, C.c_size_t, i32p,
                                       C.c_int, f32p, f32p]

    def logsum(self, a, b):
        return self.L.orc_logsum(a, b)

    def logsum_table(self):
        t = np.zeros(16000, dtype=np.float32)
        self.L.orc_logsum_table(_p(t, f32p))
        return t

    def run(self, desc: ModelDesc, mode, codes, lens, threshold=0.0, minlen=16, dust=100, matchstart=-1,
            matchend=-1, want_labels=1, threads=1):
        codes = np.ascontiguousarray(codes, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        n, stride = codes.shape
        rp = RunParamsC(threshold, minlen, matchstart, matchend, dust, want_labels)
        out = dict(mapq=np.zeros(n, np.float32), bar_prob=np.zeros(n, np.float32), f_score=np.zeros(n, np.float32),
                   b_score=np.zeros(n, np.float32), r_score=np.zeros(n, np.float32),
                   read_type=np.zeros(n, np.int32), barcode=np.zeros(n, np.int32),
                   fingerprint=np.zeros(n, np.int32), labels=np.zeros((n, stride), np.uint8),
                   seq=np.zeros((n, stride), np.uint8), len=np.zeros(n, np.int32))
        st = self.L.orc_run(C.byref(desc.c), C.byref(rp), mode, n, _p(codes, u8p), stride, _p(lens, i32p), threads,
                            _p(out["mapq"], f32p), _p(out["bar_prob"], f32p), _p(out["f_score"], f32p),
                            _p(out["b_score"], f32p), _p(out["r_score"], f32p), _p(out["read_type"], i32p),
                            _p(out["barcode"], i32p), _p(out["fingerprint"], i32p), _p(out["labels"], u8p),
                            _p(out["seq"], u8p), _p(out["len"], i32p))
        assert st == 0
        return out

    def set_reference(self, ref_codes, s_index, filter_error=2):
        self.L.orc_set_reference.argtypes = [u8p, i32p, C.c_int, C.c_int]
        self._ref = (np.ascontiguousarray(ref_codes, np.uint8), np.ascontiguousarray(s_index, np.int32))   # keep alive
        self.L.orc_set_reference(_p(self._ref[0], u8p), _p(self._ref[1], i32p), max(len(self._ref[1]) - 1, 0), filter_error)

    def run_rna_dust(self, codes, lens, dust=100, threads=1):
        self.L.orc_run_rna_dust.argtypes = [C.c_int, u8p, C.c_size_t, i32p, C.c_int, C.c_int, i32p]
        codes = np.ascontiguousarray(codes, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        n, stride = codes.shape
        rt = np.zeros(n, np.int32)
        assert self.L.orc_run_rna_dust(n, _p(codes, u8p), stride, _p(lens, i32p), threads, dust, _p(rt, i32p)) == 0
        return rt

    def myers(self, which, t, p):
        fn = self.L.orc_bmp_single if which == "bmp_single" else self.L.orc_bpm_check_error
        fn.argtypes = [u8p, u8p, C.c_int, C.c_int]
        t = np.ascontiguousarray(t, np.uint8); p = np.ascontiguousarray(p, np.uint8)
        return fn(_p(t, u8p), _p(p, u8p), len(t), len(p))

    def arch_compare(self, descs, codes, lens, threads=1):
        codes = np.ascontiguousarray(codes, np.uint8)
        lens = np.ascontiguousarray(lens, np.int32)
        n, stride = codes.shape
        arr = (C.POINTER(ModelDescC) * len(descs))(*[C.pointer(d.c) for d in descs])
        bs = np.zeros((len(descs), n), np.float32)
        post = np.zeros(len(descs), np.float32)
        st = self.L.orc_arch_compare(arr, len(descs), n, _p(codes, u8p), stride, _p(lens, i32p), threads,
                                     _p(bs, f32p), _p(post, f32p))
        assert st == 0
        return bs, post
