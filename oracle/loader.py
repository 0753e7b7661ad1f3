"""ctypes loaders for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, tools/gen_golden.py, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.  Nothing under
one_b200/ does.

  Oracle  - oracle/_build/libred_oracle.so  (red_oracle.c, our restatement)
  RefLib  - oracle/_ref/libred_ref.so       (the reference itself + ref_shim)
"""

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libred_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libred_ref.so")

# Matcher.h:62-69
STY_INSTANT, STY_FIRST, STY_TANGENT, STY_LAST, STY_FULL = 1, 2, 3, 4, 5
STYLES = {"instant": 1, "first": 2, "tangent": 3, "last": 4, "full": 5}
VERB_CHECK, VERB_MATCH, VERB_SCAN, VERB_SEARCH = 0, 1, 2, 3
VERBS = {"check": 0, "match": 1, "scan": 2, "search": 3}
# Consts.h:14-18
F_IGNORE_CASE, F_LOOSE_START, F_LOOSE_END = 1, 2, 4
# Parser.h:43-49
LANG_RAW, LANG_AUTO, LANG_GLOB, LANG_EXACT = 1, 2, 3, 4
FMT_AUTO = 255


def build(ref=True):
    """make the checkers (the reference half only where /root/reference is)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def have_ref():
    return os.path.exists(REF_SO)


class _Outcome(C.Structure):
    _fields_ = [("result", C.c_int32), ("start", C.c_uint64), ("end", C.c_uint64)]


def _u8ptr(b):
    return C.cast(C.c_char_p(b), C.c_void_p)


class Oracle:
    """red_oracle.c: one loaded serialized program."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(ORACLE_SO):
                build(ref=False)
            lib = C.CDLL(ORACLE_SO)
            lib.red_oracle_load.restype = C.c_void_p
            lib.red_oracle_load.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_char_p)]
            lib.red_oracle_free.argtypes = [C.c_void_p]
            lib.red_oracle_info.restype = C.c_uint32
            lib.red_oracle_info.argtypes = [C.c_void_p, C.c_int]
            lib.red_oracle_checksum.restype = C.c_uint32
            lib.red_oracle_checksum.argtypes = [C.c_void_p, C.c_uint64]
            lib.red_oracle_check_header.restype = C.c_char_p
            lib.red_oracle_check_header.argtypes = [C.c_void_p, C.c_uint64]
            lib.red_oracle_check.restype = C.c_int32
            lib.red_oracle_check.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
            lib.red_oracle_scan.restype = C.c_int32
            lib.red_oracle_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int]
            lib.red_oracle_match.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                             C.POINTER(_Outcome)]
            lib.red_oracle_search.argtypes = lib.red_oracle_match.argtypes
            lib.red_oracle_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_uint64, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
            lib.red_oracle_match_all.restype = C.c_uint64
            lib.red_oracle_match_all.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                                 C.c_void_p, C.c_void_p, C.c_void_p]
            lib.red_oracle_collect.restype = C.c_uint64
            lib.red_oracle_collect.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                               C.c_void_p, C.c_void_p, C.c_void_p]
            lib.red_oracle_replace.restype = C.c_uint64
            lib.red_oracle_replace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                               C.c_uint64, C.c_int, C.c_void_p, C.c_uint64,
                                               C.POINTER(C.c_uint64)]
            lib.red_oracle_advance.restype = C.c_int32
            lib.red_oracle_advance.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.c_uint64,
                                               C.c_void_p]
            cls._lib = lib
        return cls._lib

    FRESH = 0xffffffffffffffff

    def __init__(self, prog: bytes):
        lib = self.lib()
        why = C.c_char_p()
        self._h = lib.red_oracle_load(_u8ptr(prog), len(prog), C.byref(why))
        if not self._h:
            raise ValueError(why.value.decode())

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib().red_oracle_free(self._h)
            self._h = None

    @classmethod
    def check_header(cls, prog: bytes):
        msg = cls.lib().red_oracle_check_header(_u8ptr(prog), len(prog))
        return msg.decode() if msg else None

    def info(self):
        names = ["fmt", "max_char", "leader_len", "state_cnt", "initial_off", "leader_off"]
        return {n: self.lib().red_oracle_info(self._h, i) for i, n in enumerate(names)}

    def run(self, text: bytes, verb, style, do_leader=True):
        """-> (result, start, end); start/end are 0 for check/scan."""
        lib = self.lib()
        verb = VERBS.get(verb, verb)
        style = STYLES.get(style, style)
        if verb == VERB_CHECK:
            return (lib.red_oracle_check(self._h, _u8ptr(text), len(text), style, int(do_leader)), 0, 0)
        if verb == VERB_SCAN:
            return (lib.red_oracle_scan(self._h, _u8ptr(text), len(text), style, int(do_leader)), 0, 0)
        oc = _Outcome()
        fn = lib.red_oracle_match if verb == VERB_MATCH else lib.red_oracle_search
        fn(self._h, _u8ptr(text), len(text), style, int(do_leader), C.byref(oc))
        return (oc.result, oc.start, oc.end)

    def batch(self, text, starts, lens, verb=VERB_MATCH, style=STY_LAST, do_leader=True):
        """text: uint8 ndarray; starts: uint64; lens: uint32 -> (results, starts, ends)."""
        text = np.ascontiguousarray(text, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        n = len(starts)
        res = np.zeros(n, dtype=np.int32)
        so = np.zeros(n, dtype=np.uint32)
        eo = np.zeros(n, dtype=np.uint32)
        self.lib().red_oracle_batch(self._h, text.ctypes.data, starts.ctypes.data, lens.ctypes.data,
                                    n, VERBS.get(verb, verb), STYLES.get(style, style), int(do_leader),
                                    res.ctypes.data, so.ctypes.data, eo.ctypes.data)
        return res, so, eo

    def match_all(self, text: bytes, cap=4096):
        res = np.zeros(cap, dtype=np.int32)
        st = np.zeros(cap, dtype=np.uint64)
        en = np.zeros(cap, dtype=np.uint64)
        n = self.lib().red_oracle_match_all(self._h, _u8ptr(text), len(text), cap,
                                            res.ctypes.data, st.ctypes.data, en.ctypes.data)
        n = min(n, cap)
        return [(int(res[i]), int(st[i]), int(en[i])) for i in range(n)]

    def collect(self, text: bytes, cap=4096):
        res = np.zeros(cap, dtype=np.int32)
        st = np.zeros(cap, dtype=np.uint64)
        en = np.zeros(cap, dtype=np.uint64)
        n = self.lib().red_oracle_collect(self._h, _u8ptr(text), len(text), cap,
                                          res.ctypes.data, st.ctypes.data, en.ctypes.data)
        n = min(n, cap)
        return [(int(res[i]), int(st[i]), int(en[i])) for i in range(n)]


    def replace(self, text: bytes, repl: bytes, maxn, style):
        """-> (count, rewritten text)  (Matcher.h replace)"""
        cap = len(text) + (len(text) + 1) * max(1, len(repl))
        out = np.zeros(max(1, cap), dtype=np.uint8)
        olen = C.c_uint64()
        cnt = self.lib().red_oracle_replace(self._h, _u8ptr(text), len(text), _u8ptr(repl), len(repl),
                                            maxn, STYLES.get(style, style), out.ctypes.data, cap,
                                            C.byref(olen))
        assert olen.value <= cap
        return int(cnt), out[:olen.value].tobytes()

    def advance(self, state, data: bytes, per_byte=False):
        """StatefulMatcher: feed `data` from `state` (Oracle.FRESH = new matcher).
        -> (new state, result() after the last byte[, per-byte results])"""
        st = C.c_uint64(state)
        pb = np.zeros(max(1, len(data)), dtype=np.int32) if per_byte else None
        res = self.lib().red_oracle_advance(self._h, C.byref(st), _u8ptr(data), len(data),
                                            pb.ctypes.data if per_byte else None)
        if per_byte:
            return st.value, int(res), pb[:len(data)].tolist()
        return st.value, int(res)


class RefError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code  # -2 limit, -3 parse, -4 api, -1 other


class RefLib:
    """The reference itself (oracle/_ref/libred_ref.so)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
            lib = C.CDLL(REF_SO)
            lib.ref_last_error.restype = C.c_char_p
            lib.ref_compile.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_uint64,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
            lib.ref_free.argtypes = [C.c_void_p]
            lib.ref_load.restype = C.c_void_p
            lib.ref_load.argtypes = [C.c_void_p, C.c_uint64]
            lib.ref_unload.argtypes = [C.c_void_p]
            lib.ref_run.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                    C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
            lib.ref_run_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64,
                                          C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
            lib.ref_replace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                        C.c_uint64, C.c_int, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
            lib.ref_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_uint64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
            lib.ref_stateful.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
            cls._lib = lib
        return cls._lib

    @classmethod
    def _raise(cls, code):
        raise RefError(code, cls.lib().ref_last_error().decode(errors="replace"))

    @classmethod
    def compile(cls, rules, fmt=FMT_AUTO, max_states=0) -> bytes:
        """rules: iterable of (pattern: bytes|str, result: int, flags: int[, lang: int])."""
        lib = cls.lib()
        pats, results, flags, langs = [], [], [], []
        for r in rules:
            p = r[0].encode("latin-1") if isinstance(r[0], str) else bytes(r[0])
            pats.append(p)
            results.append(r[1])
            flags.append(r[2])
            langs.append(r[3] if len(r) > 3 else LANG_RAW)
        n = len(pats)
        arr = (C.c_char_p * n)(*pats)
        plens = np.array([len(p) for p in pats], dtype=np.uint64)
        res = np.array(results, dtype=np.int32)
        flg = np.array(flags, dtype=np.uint32)
        lng = np.array(langs, dtype=np.int32)
        out = C.c_void_p()
        outlen = C.c_uint64()
        rc = lib.ref_compile(n, arr, plens.ctypes.data, res.ctypes.data, flg.ctypes.data,
                             lng.ctypes.data, fmt, max_states, C.byref(out), C.byref(outlen))
        if rc != 0:
            cls._raise(rc)
        buf = C.string_at(out.value, outlen.value)
        lib.ref_free(out)
        return buf

    def __init__(self, prog: bytes):
        self._h = self.lib().ref_load(_u8ptr(prog), len(prog))
        if not self._h:
            self._raise(-1)

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib().ref_unload(self._h)
            self._h = None

    def run(self, text: bytes, verb, style):
        res, st, en = C.c_int32(), C.c_uint64(), C.c_uint64()
        rc = self.lib().ref_run(self._h, _u8ptr(text), len(text), VERBS.get(verb, verb),
                                STYLES.get(style, style), C.byref(res), C.byref(st), C.byref(en))
        if rc != 0:
            self._raise(rc)
        return (res.value, st.value, en.value)

    def batch(self, text, starts, lens, verb=VERB_MATCH, style=STY_LAST, nthreads=1):
        text = np.ascontiguousarray(text, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        n = len(starts)
        res = np.zeros(n, dtype=np.int32)
        so = np.zeros(n, dtype=np.uint32)
        eo = np.zeros(n, dtype=np.uint32)
        rc = self.lib().ref_run_batch(self._h, text.ctypes.data, starts.ctypes.data, lens.ctypes.data, n,
                                      VERBS.get(verb, verb), STYLES.get(style, style), nthreads,
                                      res.ctypes.data, so.ctypes.data, eo.ctypes.data)
        if rc != 0:
            self._raise(rc)
        return res, so, eo

    def replace(self, text: bytes, repl: bytes, maxn, style):
        out, outlen, cnt = C.c_void_p(), C.c_uint64(), C.c_uint64()
        rc = self.lib().ref_replace(self._h, _u8ptr(text), len(text), _u8ptr(repl), len(repl), maxn,
                                    STYLES.get(style, style), C.byref(out), C.byref(outlen), C.byref(cnt))
        if rc != 0:
            self._raise(rc)
        buf = C.string_at(out.value, outlen.value)
        self.lib().ref_free(out)
        return cnt.value, buf

    def stateful(self, text: bytes):
        """A fresh StatefulMatcher fed `text`: [result() before input, advance() per byte...]"""
        res = np.zeros(len(text) + 1, dtype=np.int32)
        rc = self.lib().ref_stateful(self._h, _u8ptr(text), len(text), res.ctypes.data)
        if rc != 0:
            self._raise(rc)
        return res.tolist()

    def multi(self, text: bytes, which, cap=4096):
        """which: 'match_all' (Matcher.h matchAll) or 'collect' (Red::collect)."""
        res = np.zeros(cap, dtype=np.int32)
        st = np.zeros(cap, dtype=np.uint64)
        en = np.zeros(cap, dtype=np.uint64)
        cnt = C.c_uint64()
        rc = self.lib().ref_multi(self._h, _u8ptr(text), len(text), 0 if which == "match_all" else 1,
                                  cap, res.ctypes.data, st.ctypes.data, en.ctypes.data, C.byref(cnt))
        if rc != 0:
            self._raise(rc)
        n = min(cnt.value, cap)
        return [(int(res[i]), int(st[i]), int(en[i])) for i in range(n)]
