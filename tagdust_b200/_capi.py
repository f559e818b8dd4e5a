"""ctypes binding of include/tagdust_b200.h (the C ABI of libtagdust_b200.so).

Nothing here computes: it declares the POD structs and prototypes, and loads the
in-tree shared library.  Loading fails loudly when the library is missing -- there is
no CPU fallback for the hot path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtagdust_b200.so")

# modes / codes (barcode_hmm.h:128-132, io.h:36-52)
MODE_GET_LABEL = 1
MODE_GET_PROB = 4
MODE_ARCH_COMP = 5
MODE_RNA_DUST = 6
EXTRACT_SUCCESS = 0
EXTRACT_FAIL_ARCHITECTURE_MISMATCH = 1
EXTRACT_FAIL_READ_TOO_SHORT = 2
EXTRACT_FAIL_BAR_FINGER_NOT_FOUND = 3
EXTRACT_FAIL_MATCHES_ARTIFACTS = 5
EXTRACT_FAIL_LOW_COMPLEXITY = 6
TDG_OK, TDG_FAIL, TDG_EMEM, TDG_ENODEV, TDG_EINVAL, TDG_ECUDA = 0, 1, 2, 16, 17, 18
LOGSUM_SIZE = 16000

c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)


class ModelDescC(C.Structure):
    _fields_ = [
        ("num_segments", C.c_int32),
        ("total_hmms", C.c_int32),
        ("total_columns", C.c_int32),
        ("average_raw_length", C.c_int32),
        ("seg_type", C.c_char_p),
        ("seg_num_hmms", c_int32_p),
        ("seg_num_cols", c_int32_p),
        ("seg_skip", c_float_p),
        ("background", c_float_p),
        ("transition", c_float_p),
        ("m_emit", c_float_p),
        ("i_emit", c_float_p),
        ("silent_to_M", c_float_p),
        ("silent_to_I", c_float_p),
        ("label", c_int32_p),
        ("transition_matrix", c_float_p),
    ]


class RunParamsC(C.Structure):
    _fields_ = [
        ("confidence_threshold", C.c_float),
        ("minlen", C.c_int32),
        ("matchstart", C.c_int32),
        ("matchend", C.c_int32),
        ("dust", C.c_int32),
        ("want_labels", C.c_int32),
        ("want_spans", C.c_int32),
        ("refset", C.c_void_p),
        ("filter_error", C.c_int32),
        ("slice_threads", C.c_int32),
    ]


class ResultC(C.Structure):
    _fields_ = [
        ("n_reads", C.c_int32),
        ("label_stride", C.c_int32),
        ("mapq", c_float_p),
        ("bar_prob", c_float_p),
        ("f_score", c_float_p),
        ("b_score", c_float_p),
        ("r_score", c_float_p),
        ("read_type", c_int32_p),
        ("extracted", c_uint8_p),
        ("barcode", c_int32_p),
        ("fingerprint", c_int32_p),
        ("labels", c_uint8_p),
        ("span_stride", C.c_int32),
        ("spans", C.POINTER(C.c_uint16)),
    ]


class ArchParamsC(C.Structure):
    _fields_ = [
        ("background_logp", C.c_double * 5),
        ("average_length", C.c_double),
        ("max_seq_len", C.c_int32),
        ("expected_5_len", C.c_double),
        ("mean_5_len", C.c_double),
        ("stdev_5_len", C.c_double),
        ("expected_3_len", C.c_double),
        ("mean_3_len", C.c_double),
        ("stdev_3_len", C.c_double),
        ("sequencer_error_rate", C.c_float),
        ("indel_frequency", C.c_float),
        ("calibration_edit", C.c_int32),
    ]


#: every symbol include/tagdust_b200.h declares -> (restype, argtypes)
PROTOTYPES = {
    "tdg_init": (C.c_int, [C.c_int, c_int32_p, C.POINTER(C.c_void_p)]),
    "tdg_shutdown": (None, [C.c_void_p]),
    "tdg_device_count": (C.c_int, [C.c_void_p]),
    "tdg_last_error": (C.c_char_p, []),
    "tdg_version": (C.c_char_p, []),
    "tdg_logsum_table": (None, [c_float_p]),
    "tdg_logsum_host": (C.c_float, [C.c_float, C.c_float]),
    "tdg_model_create": (C.c_int, [C.c_void_p, C.POINTER(ModelDescC), C.c_int, C.POINTER(C.c_void_p)]),
    "tdg_model_destroy": (None, [C.c_void_p]),
    "tdg_model_validate": (C.c_int, [C.POINTER(ModelDescC), C.c_char_p, C.c_size_t]),
    "tdg_arch_compile": (C.c_int, [C.c_int, C.POINTER(C.c_char_p), C.POINTER(ArchParamsC), C.POINTER(C.c_void_p)]),
    "tdg_arch_desc": (C.POINTER(ModelDescC), [C.c_void_p]),
    "tdg_arch_destroy": (None, [C.c_void_p]),
    "tdg_refset_create": (C.c_int, [C.c_void_p, c_uint8_p, c_int32_p, C.c_int, C.POINTER(C.c_void_p)]),
    "tdg_refset_destroy": (None, [C.c_void_p]),
    "tdg_batch_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "tdg_batch_destroy": (None, [C.c_void_p]),
    "tdg_batch_clear": (C.c_int, [C.c_void_p]),
    "tdg_batch_append_codes": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "tdg_batch_append_records": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]),
    "tdg_batch_size": (C.c_int, [C.c_void_p]),
    "tdg_plan_shards": (C.c_int, [C.c_int, C.c_int, c_int32_p, c_int32_p]),
    "tdg_submit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(RunParamsC), C.c_void_p]),
    "tdg_wait": (C.c_int, [C.c_void_p, C.POINTER(ResultC)]),
    "tdg_run": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(RunParamsC), C.c_void_p, C.POINTER(ResultC)]),
    "tdg_arch_compare": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int, c_float_p, c_float_p]),
    "tdg_batch_upload": (C.c_int, [C.c_void_p, C.c_void_p]),
    "tdg_decode_resident": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(RunParamsC), C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "tdg_batch_download": (C.c_int, [C.c_void_p, C.POINTER(ResultC)]),
    "tdg_batch_cells": (C.c_double, [C.c_void_p, C.c_void_p]),
    "tdg_desc_live_ops": (C.c_int, [C.POINTER(ModelDescC), C.POINTER(C.c_double)]),
    "tdg_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "tdg_profile_read": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
}


class FastqChunkC(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("max_len", C.c_int32),
        ("len", c_int32_p),
        ("seq_off", C.POINTER(C.c_uint64)),
        ("codes", c_uint8_p),
        ("qual", c_uint8_p),
        ("name_off", C.POINTER(C.c_uint64)),
        ("names", C.POINTER(C.c_char)),
    ]


class DemuxInputC(C.Structure):
    _fields_ = [
        ("path", C.c_char_p),
        ("fasta", C.c_int32),
        ("model", C.c_void_p),
        ("num_read_segments", C.c_int32),
        ("confidence_threshold", C.c_float),
        ("max_seq_len", C.c_int32),
        ("expected_len", C.c_int32),
    ]


class DemuxJobC(C.Structure):
    _fields_ = [
        ("n_inputs", C.c_int32),
        ("inputs", C.POINTER(DemuxInputC)),
        ("barcode_input", C.c_int32),
        ("num_alternatives", C.c_int32),
        ("barcode_names", C.POINTER(C.c_char_p)),
        ("outfile", C.c_char_p),
        ("minlen", C.c_int32),
        ("dust", C.c_int32),
        ("matchstart", C.c_int32),
        ("matchend", C.c_int32),
        ("print_seq_finger", C.c_int32),
        ("threads", C.c_int32),
        ("chunk_reads", C.c_int32),
        ("refset", C.c_void_p),
        ("filter_error", C.c_int32),
        ("ref_chunk_reads", C.c_int32),
        ("artifact_counts", C.POINTER(C.c_int64)),
    ]


class DemuxStatsC(C.Structure):
    _fields_ = [(k, C.c_int64) for k in (
        "total_read", "num_EXTRACT_SUCCESS", "num_EXTRACT_FAIL_BAR_FINGER_NOT_FOUND", "num_EXTRACT_FAIL_READ_TOO_SHORT",
        "num_EXTRACT_FAIL_AMBIGIOUS_BARCODE", "num_EXTRACT_FAIL_ARCHITECTURE_MISMATCH", "num_EXTRACT_FAIL_MATCHES_ARTIFACTS",
        "num_EXTRACT_FAIL_LOW_COMPLEXITY", "long_sequence_events")] + [(k, C.c_double) for k in (
            "seconds_split", "seconds_parse", "seconds_gpu_wait", "seconds_write", "seconds_total")]


class SeqStatsC(C.Structure):
    _fields_ = [("total_read", C.c_int64), ("max_seq_len", C.c_int32), ("sum_len", C.c_double), ("base_count", C.c_double * 5),
                ("five_s0", C.c_double), ("five_s1", C.c_double), ("five_s2", C.c_double),
                ("three_s0", C.c_double), ("three_s1", C.c_double), ("three_s2", C.c_double)]


#: every symbol include/tagdust_b200_stream.h declares
STREAM_PROTOTYPES = {
    "tdg_sequence_stats": (C.c_int, [C.c_char_p, C.c_int, C.c_int, c_uint8_p, C.c_int, c_uint8_p, C.c_int, C.c_int, C.POINTER(SeqStatsC)]),
    "tdg_fastq_open": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "tdg_fastq_next": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(FastqChunkC)]),
    "tdg_fastq_close": (None, [C.c_void_p]),
    "tdg_batch_append_ragged": (C.c_int, [C.c_void_p, C.c_int, c_uint8_p, C.POINTER(C.c_uint64), c_int32_p, C.c_int]),
    "tdg_format_rq": (C.c_int, [C.c_float, C.c_char_p]),
    "tdg_batch_reserve_labels": (C.c_int, [C.c_void_p]),
    "tdg_demux_run": (C.c_int, [C.c_void_p, C.POINTER(DemuxJobC), C.POINTER(DemuxStatsC)]),
    "tdg_model_max_len": (C.c_int, [C.c_void_p]),
    "tdg_model_set_max_len": (C.c_int, [C.c_void_p, C.c_int]),
    "tdg_model_num_hmms": (C.c_int, [C.c_void_p]),
    "tdg_model_read_hmms": (C.c_int, [C.c_void_p, c_uint8_p]),
}

_lib = None


def load_library(path=None):
    """Load libtagdust_b200.so (built in-tree by `make -C tagdust_b200/csrc` /
    __graft_entry__.build()).  Raises OSError when it is missing: no fallback."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("TDG_LIB") or LIB_PATH   # TDG_LIB: A/B runs of two builds on the same box
    if not os.path.exists(p):
        raise OSError(
            f"{p} not found: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in list(PROTOTYPES.items()) + list(STREAM_PROTOTYPES.items()):
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _fptr(a):
    return a.ctypes.data_as(c_float_p)


def _iptr(a):
    return a.ctypes.data_as(c_int32_p)


class ModelDesc:
    """Owns numpy arrays laid out as tdg_model_desc wants them and exposes `.c`."""

    FIELDS = ("seg_num_hmms", "seg_num_cols", "seg_skip", "background", "transition", "m_emit",
              "i_emit", "silent_to_M", "silent_to_I", "label", "transition_matrix")

    def __init__(self, seg_type, seg_num_hmms, seg_num_cols, seg_skip, background, transition,
                 m_emit, i_emit, silent_to_M, silent_to_I, label, transition_matrix,
                 average_raw_length):
        self.seg_type = bytes(seg_type)
        self.seg_num_hmms = np.ascontiguousarray(seg_num_hmms, dtype=np.int32)
        self.seg_num_cols = np.ascontiguousarray(seg_num_cols, dtype=np.int32)
        self.seg_skip = np.ascontiguousarray(seg_skip, dtype=np.float32)
        self.background = np.ascontiguousarray(background, dtype=np.float32)
        self.transition = np.ascontiguousarray(transition, dtype=np.float32).reshape(-1, 9)
        self.m_emit = np.ascontiguousarray(m_emit, dtype=np.float32).reshape(-1, 5)
        self.i_emit = np.ascontiguousarray(i_emit, dtype=np.float32).reshape(-1, 5)
        self.silent_to_M = np.ascontiguousarray(silent_to_M, dtype=np.float32)
        self.silent_to_I = np.ascontiguousarray(silent_to_I, dtype=np.float32)
        self.label = np.ascontiguousarray(label, dtype=np.int32)
        H = self.label.shape[0]
        self.transition_matrix = np.ascontiguousarray(transition_matrix, dtype=np.float32).reshape(H, H)
        self.average_raw_length = int(average_raw_length)
        self.num_segments = len(self.seg_type)
        self.total_hmms = H
        self.total_columns = int(self.transition.shape[0])
        assert int((self.seg_num_hmms * self.seg_num_cols).sum()) == self.total_columns
        assert int(self.seg_num_hmms.sum()) == H
        c = ModelDescC()
        c.num_segments = self.num_segments
        c.total_hmms = H
        c.total_columns = self.total_columns
        c.average_raw_length = self.average_raw_length
        c.seg_type = self.seg_type
        c.seg_num_hmms = _iptr(self.seg_num_hmms)
        c.seg_num_cols = _iptr(self.seg_num_cols)
        c.seg_skip = _fptr(self.seg_skip)
        c.background = _fptr(self.background)
        c.transition = _fptr(self.transition)
        c.m_emit = _fptr(self.m_emit)
        c.i_emit = _fptr(self.i_emit)
        c.silent_to_M = _fptr(self.silent_to_M)
        c.silent_to_I = _fptr(self.silent_to_I)
        c.label = _iptr(self.label)
        c.transition_matrix = _fptr(self.transition_matrix)
        self.c = c

    @classmethod
    def from_c(cls, cdesc):
        """Deep-copy a tdg_model_desc (e.g. the one tdg_arch_desc returns)."""
        S, H, Cn = cdesc.num_segments, cdesc.total_hmms, cdesc.total_columns

        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True)

        return cls(
            seg_type=cdesc.seg_type[:S],
            seg_num_hmms=arr(cdesc.seg_num_hmms, S, np.int32),
            seg_num_cols=arr(cdesc.seg_num_cols, S, np.int32),
            seg_skip=arr(cdesc.seg_skip, S, np.float32),
            background=arr(cdesc.background, 5, np.float32),
            transition=arr(cdesc.transition, Cn * 9, np.float32),
            m_emit=arr(cdesc.m_emit, Cn * 5, np.float32),
            i_emit=arr(cdesc.i_emit, Cn * 5, np.float32),
            silent_to_M=arr(cdesc.silent_to_M, Cn, np.float32),
            silent_to_I=arr(cdesc.silent_to_I, Cn, np.float32),
            label=arr(cdesc.label, H, np.int32),
            transition_matrix=arr(cdesc.transition_matrix, H * H, np.float32),
            average_raw_length=cdesc.average_raw_length,
        )

    def same_bits(self, other):
        """Bitwise comparison (NaN/-inf safe) of every array; returns list of differing fields."""
        bad = []
        if self.seg_type != other.seg_type:
            bad.append("seg_type")
        if self.average_raw_length != other.average_raw_length:
            bad.append("average_raw_length")
        for f in self.FIELDS:
            a, b = getattr(self, f), getattr(other, f)
            if a.shape != b.shape or a.tobytes() != b.tobytes():
                bad.append(f)
        return bad

    def cells_per_read(self, length):
        """Profile-column cells of one read: backward + forward = 2*L*C."""
        return 2 * int(length) * self.total_columns
