"""TEST INFRASTRUCTURE ONLY. ctypes bindings of the bag-of-words oracle (Frame::ComputeBoW): the CPU
restatement (oracle/liborb_oracle.so, orb_oracle_bow.cc) and the reference's own Thirdparty/DBoW2 compiled unmodified
(oracle/_ref/libmorb_ref_bow.so, ref_driver_bow.cc). Same import rules as oracle_py."""
import ctypes as C
import os

import numpy as np

from oracle.oracle_py import ORACLE_SO, HERE, _Lib, _p

REF_BOW_SO = os.path.join(HERE, "_ref", "libmorb_ref_bow.so")

_T_ARGS = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]


def have_reference():
    return os.path.exists(REF_BOW_SO)


def _run(fn, handle, desc, levelsup, extra):
    desc = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
    n = len(desc)
    cap = max(n, 1)
    bw, bv = np.zeros(cap, np.uint32), np.zeros(cap, np.float64)
    fn_, fo, ff = np.zeros(cap, np.uint32), np.zeros(cap + 1, np.int32), np.zeros(cap, np.uint32)
    nb, nf = C.c_int(0), C.c_int(0)
    rc = fn(handle, _p(desc), n, levelsup, cap, _p(bw), _p(bv), C.byref(nb), _p(fn_), _p(fo), _p(ff), C.byref(nf), *extra)
    assert rc == 0
    nb, nf = nb.value, nf.value
    return dict(bow_word=bw[:nb], bow_val=bv[:nb], fv_node=fn_[:nf], fv_off=fo[:nf + 1], fv_feat=ff[:fo[nf]])


class OracleVocabulary:
    """The restatement; takes the vocabulary as arrays (morb_slam_b200.synth.synth_vocabulary)."""

    def __init__(self, voc):
        self.lib = _Lib.load(ORACLE_SO)
        self.lib.oro_vocab_create.restype = C.c_void_p
        self.lib.oro_vocab_create.argtypes = [C.c_int] * 5 + [C.c_void_p] * 4
        self.lib.oro_vocab_free.argtypes = [C.c_void_p]
        self.lib.oro_bow_transform.argtypes = _T_ARGS + [C.c_void_p, C.c_void_p]
        parent = np.ascontiguousarray(voc["parent"], np.int32)
        leaf = np.ascontiguousarray(voc["is_leaf"], np.uint8)
        desc = np.ascontiguousarray(voc["desc"], np.uint8)
        weight = np.ascontiguousarray(voc["weight"], np.float64)
        self.h = self.lib.oro_vocab_create(voc["k"], voc["L"], voc["scoring"], voc["weighting"], len(parent), _p(parent), _p(leaf), _p(desc),
                                           _p(weight))

    def transform(self, desc, levelsup=4):
        n = len(np.asarray(desc).reshape(-1, 32))
        fw, fnode = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        r = _run(self.lib.oro_bow_transform, self.h, desc, levelsup, (_p(fw), _p(fnode)))
        r["feat_word"], r["feat_node"] = fw[:n], fnode[:n]
        return r

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oro_vocab_free(self.h)
            self.h = None


class ReferenceVocabulary:
    """The reference's DBoW2: loads the ORBvoc.txt text format with its own loadFromTextFile."""

    def __init__(self, path):
        self.lib = C.CDLL(REF_BOW_SO)
        self.lib.refb_vocab_load_text.restype = C.c_void_p
        self.lib.refb_vocab_load_text.argtypes = [C.c_char_p]
        self.lib.refb_vocab_free.argtypes = [C.c_void_p]
        self.lib.refb_vocab_info.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.refb_transform.argtypes = _T_ARGS
        self.h = self.lib.refb_vocab_load_text(path.encode())
        assert self.h, "loadFromTextFile failed"

    def info(self):
        a = np.zeros(5, np.int32)
        self.lib.refb_vocab_info(self.h, _p(a))
        return dict(k=int(a[0]), L=int(a[1]), scoring=int(a[2]), weighting=int(a[3]), words=int(a[4]))

    def transform(self, desc, levelsup=4):
        return _run(self.lib.refb_transform, self.h, desc, levelsup, ())

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.refb_vocab_free(self.h)
            self.h = None
