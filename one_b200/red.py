"""Python face of libred_b200.so, mirroring RED's public API for the hot path.

Names follow the reference (quol/red/include): Parser.add/addAuto/addGlob/
addExact (Parser.h:64-95), compile (Compile.h:28), Executable -> Program,
check/match/scan/search with the five styles (Matcher.h:62-94), Outcome
(Outcome.h:33), and Red's orthogonal names matchFull()/searchLast()/...
(Red.h:113-132).  Exceptions mirror Except.h.

Everything computes on the GPU through the C ABI in include/red_b200.h.  There
is no CPU path here: without the built library or without an sm_100 device the
calls raise.  torch is used by callers for device buffers and streams only;
this module needs just ctypes and numpy.
"""

import ctypes as C
import os
from collections import namedtuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libred_b200.so")

# Matcher.h:62-69
styInstant, styFirst, styTangent, styLast, styFull = 1, 2, 3, 4, 5
STYLES = {"instant": 1, "first": 2, "tangent": 3, "last": 4, "full": 5}
VERB_CHECK, VERB_MATCH, VERB_SCAN, VERB_SEARCH = 0, 1, 2, 3
VERBS = {"check": 0, "match": 1, "scan": 2, "search": 3}
# Consts.h:14-18
fIgnoreCase, fLooseStart, fLooseEnd = 1, 2, 4
# Parser.h:43-49
langRegexRaw, langRegexAuto, langGlob, langExact = 1, 2, 3, 4
# Serializer.h:37-43
fmtDirect1, fmtDirect2, fmtDirect4, fmtDirectAuto = 1, 2, 4, 255

LINE_MATCH_DTYPE = np.dtype([("line_start", "<u8"), ("line_len", "<u4"), ("result", "<i4"),
                             ("match_start", "<u4"), ("match_end", "<u4")])

SPAN_DTYPE = np.dtype([("result", "<i4"), ("start", "<u4"), ("end", "<u4")])

STATE_FRESH = 0xffffffff  # RED_B200_STATE_FRESH: a StatefulMatcher that has seen no input
_E_LIMIT = -3

Outcome = namedtuple("Outcome", ["result", "start", "end"])


class RedExcept(RuntimeError):
    pass


class RedExceptInternal(RedExcept):
    pass


class RedExceptUser(RedExcept):
    pass


class RedExceptLimit(RedExcept):
    pass


class RedExceptExec(RedExceptInternal):
    pass


class RedExceptParse(RedExceptUser):
    pass


class RedExceptApi(RedExceptUser):
    pass


class RedExceptCuda(RedExceptInternal):
    pass


class RedExceptNoDevice(RedExcept):
    pass


_ERRORS = {-1: RedExceptApi, -2: RedExceptExec, -3: RedExceptLimit, -4: RedExceptParse,
           -5: RedExceptCuda, -6: RedExceptNoDevice, -7: RedExceptInternal}


class _Outcome(C.Structure):
    _fields_ = [("result", C.c_int32), ("start", C.c_uint64), ("end", C.c_uint64)]


class ProgramInfo(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("format", C.c_uint32), ("state_count", C.c_uint32),
                ("class_count", C.c_uint32), ("leader_len", C.c_uint32), ("table_mode", C.c_uint32),
                ("row_stride", C.c_uint32), ("accepting_states", C.c_uint32),
                ("pure_dead_states", C.c_uint32), ("max_result", C.c_uint32),
                ("table_bytes", C.c_uint64), ("device", C.c_int32), ("sm_count", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


_lib = None


def lib():
    """The loaded C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RedExceptInternal(
                "libred_b200.so is not built (%s); run `python -m one_b200.build`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
        L.red_b200_last_error.restype = C.c_char_p
        L.red_b200_abi_version.restype = i32
        L.red_b200_device_count.restype = i32
        L.red_b200_check_header.restype = C.c_char_p
        L.red_b200_check_header.argtypes = [vp, C.c_size_t]
        L.red_b200_program_upload.argtypes = [vp, C.c_size_t, i32, C.POINTER(vp)]
        L.red_b200_program_free.argtypes = [vp]
        L.red_b200_program_get_info.argtypes = [vp, C.POINTER(ProgramInfo)]
        L.red_b200_program_launch_count.restype = C.c_long
        L.red_b200_program_launch_count.argtypes = [vp]
        L.red_b200_program_hot_launch_count.restype = C.c_long
        L.red_b200_program_hot_launch_count.argtypes = [vp]
        L.red_b200_match_batch_device.argtypes = [vp, vp, vp, vp, u64, i32, i32, vp, vp, vp, vp]
        L.red_b200_match_batch_host.argtypes = [vp, vp, u64, vp, vp, u64, i32, i32, vp, vp, vp]
        L.red_b200_match.argtypes = [vp, vp, C.c_size_t, i32, i32, C.POINTER(_Outcome)]
        L.red_b200_split_lines_device.argtypes = [vp, u64, C.c_uint8, u64, vp, vp,
                                                  C.POINTER(u64), i32, vp]
        L.red_b200_compact_matches_device.argtypes = [vp, u64, vp, vp, i32, vp]
        L.red_b200_match_stream_device.argtypes = [vp, vp, u64, i32, i32, C.POINTER(_Outcome), vp]
        L.red_b200_multi_batch_device.argtypes = [vp, vp, vp, vp, u64, i32, C.c_uint32, vp, vp, vp]
        L.red_b200_multi_batch_host.argtypes = [vp, vp, u64, vp, vp, u64, i32, C.c_uint32, vp, vp]
        L.red_b200_replace_batch_device.argtypes = [vp, vp, vp, vp, u64, vp, C.c_uint32, u64, i32, i32,
                                                    vp, u64, vp, vp, vp, C.POINTER(u64), vp]
        L.red_b200_replace_batch_host.argtypes = [vp, vp, u64, vp, vp, u64, vp, C.c_uint32, u64, i32, i32,
                                                  vp, u64, vp, vp, vp, C.POINTER(u64)]
        L.red_b200_replace.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, u64, i32, vp, C.c_size_t,
                                       C.POINTER(C.c_size_t), C.POINTER(u64)]
        L.red_b200_advance_batch_device.argtypes = [vp, vp, vp, vp, u64, vp, vp, vp, vp, vp]
        L.red_b200_advance_batch_host.argtypes = [vp, vp, u64, vp, vp, u64, vp, vp, vp, vp]
        L.red_b200_grep_host.argtypes = [vp, vp, u64, C.c_uint8, i32, vp, u64, C.POINTER(u64), C.POINTER(u64)]
        L.red_b200_compile.argtypes = [i32, C.POINTER(C.c_char_p), vp, vp, vp, vp, i32, u64,
                                       C.POINTER(vp), C.POINTER(u64)]
        L.red_b200_free.argtypes = [vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        msg = lib().red_b200_last_error().decode(errors="replace")
        raise _ERRORS.get(rc, RedExceptInternal)(msg)


def device_count():
    return lib().red_b200_device_count()


def _style(s):
    if isinstance(s, str):
        s = STYLES[s.lower()]
    return int(s)


def _verb(v):
    if isinstance(v, str):
        v = VERBS[v.lower()]
    return int(v)


def _bytes(x):
    if isinstance(x, str):
        return x.encode("latin-1")
    return bytes(x)


class Parser:
    """Collects patterns; compile() turns them into one serialized DFA.

    Mirrors zezax::red::Parser (Parser.h:52-108).  The NFA/DFA work happens in
    the C++ pipeline behind red_b200_compile()."""

    def __init__(self, max_states=0):
        self._rules = []
        self._max_states = max_states  # Budget(states): Budget.h:33

    def add(self, regex, result, flags=0):
        self._rules.append((_bytes(regex), int(result), int(flags), langRegexRaw))

    def addAuto(self, regex, result, flags=0):
        self._rules.append((_bytes(regex), int(result), int(flags), langRegexAuto))

    def addGlob(self, glob, result, flags=0):
        self._rules.append((_bytes(glob), int(result), int(flags), langGlob))

    def addExact(self, text, result, flags=0):
        self._rules.append((_bytes(text), int(result), int(flags), langExact))

    def addAs(self, lang, text, result, flags=0):
        if lang not in (1, 2, 3, 4):
            raise RedExceptApi("unrecognized regex language")
        self._rules.append((_bytes(text), int(result), int(flags), int(lang)))

    @property
    def rules(self):
        return list(self._rules)


def compileToSerialized(parser, fmt=fmtDirectAuto) -> bytes:
    """Compile.h:30 - patterns -> serialized REDA bytes."""
    rules = parser.rules
    n = len(rules)
    pats = (C.c_char_p * max(n, 1))(*[r[0] for r in rules])
    plens = np.array([len(r[0]) for r in rules], dtype=np.uint64)
    res = np.array([r[1] for r in rules], dtype=np.int32)
    flg = np.array([r[2] for r in rules], dtype=np.uint32)
    lng = np.array([r[3] for r in rules], dtype=np.int32)
    out, outlen = C.c_void_p(), C.c_uint64()
    _check(lib().red_b200_compile(n, pats, plens.ctypes.data, res.ctypes.data, flg.ctypes.data,
                                  lng.ctypes.data, int(fmt), int(parser._max_states),
                                  C.byref(out), C.byref(outlen)))
    buf = C.string_at(out.value, outlen.value)
    lib().red_b200_free(out)
    return buf


class Program:
    """A serialized DFA uploaded to one GPU (the reference's Executable)."""

    def __init__(self, serialized: bytes, device=0):
        self._h = C.c_void_p()
        self.serialized = bytes(serialized)
        buf = C.create_string_buffer(self.serialized, len(self.serialized))
        _check(lib().red_b200_program_upload(C.cast(buf, C.c_void_p), len(self.serialized),
                                             int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            lib().red_b200_program_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def info(self):
        inf = ProgramInfo()
        _check(lib().red_b200_program_get_info(self._h, C.byref(inf)))
        return inf.as_dict()

    @property
    def launch_count(self):
        return lib().red_b200_program_launch_count(self._h)

    @property
    def hot_launch_count(self):
        """launches of the staged batch kernel (kernels_hot.cu) among them"""
        return lib().red_b200_program_hot_launch_count(self._h)

    # -- one text -------------------------------------------------------------

    def run(self, text, verb, style) -> Outcome:
        text = _bytes(text)
        oc = _Outcome()
        _check(lib().red_b200_match(self._h, C.cast(C.c_char_p(text), C.c_void_p), len(text),
                                    _verb(verb), _style(style), C.byref(oc)))
        return Outcome(oc.result, oc.start, oc.end)

    def check(self, text, style):
        return self.run(text, VERB_CHECK, style).result

    def match(self, text, style):
        return self.run(text, VERB_MATCH, style)

    def scan(self, text, style):
        return self.run(text, VERB_SCAN, style).result

    def search(self, text, style):
        return self.run(text, VERB_SEARCH, style)

    # -- batches ----------------------------------------------------------------

    def match_batch_host(self, text, starts, lens, verb=VERB_MATCH, style=styLast,
                         want_starts=True, want_ends=True, out=None):
        """numpy in, numpy out; copies to and from the GPU inside the call."""
        text = np.ascontiguousarray(text, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        n = len(starts)
        if out is None:
            res = np.zeros(n, dtype=np.int32)
            ms = np.zeros(n, dtype=np.uint32) if want_starts else None
            me = np.zeros(n, dtype=np.uint32) if want_ends else None
        else:
            res, ms, me = out
        _check(lib().red_b200_match_batch_host(
            self._h, text.ctypes.data, text.size, starts.ctypes.data, lens.ctypes.data, n,
            _verb(verb), _style(style), res.ctypes.data,
            ms.ctypes.data if ms is not None else None,
            me.ctypes.data if me is not None else None))
        return res, ms, me

    def match_batch_ptrs(self, text_ptr, starts_ptr, lens_ptr, n, verb, style,
                         results_ptr, mstarts_ptr, mends_ptr, stream=0):
        """Raw device pointers (ints); asynchronous on `stream`."""
        _check(lib().red_b200_match_batch_device(
            self._h, text_ptr, starts_ptr, lens_ptr, int(n), _verb(verb), _style(style),
            results_ptr, mstarts_ptr or None, mends_ptr or None, stream or None))

    def match_batch_device(self, text, starts, lens, verb=VERB_MATCH, style=styLast,
                           want_starts=False, want_ends=True, out=None, stream=None):
        """torch CUDA tensors in (uint8, int64/uint64, int32/uint32), torch out.

        Launches on torch's current stream unless `stream` is given."""
        import torch
        n = starts.numel()
        dev = text.device
        if out is None:
            res = torch.empty(n, dtype=torch.int32, device=dev)
            ms = torch.empty(n, dtype=torch.int32, device=dev) if want_starts else None
            me = torch.empty(n, dtype=torch.int32, device=dev) if want_ends else None
        else:
            res, ms, me = out
        st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        self.match_batch_ptrs(text.data_ptr(), starts.data_ptr(), lens.data_ptr(), n, verb, style,
                              res.data_ptr(), ms.data_ptr() if ms is not None else 0,
                              me.data_ptr() if me is not None else 0, st)
        return res, ms, me

    def multi_batch_host(self, text, starts, lens, which="collect", max_per_line=16):
        """Per line all matches: which = 'match_all' (matchAll, Matcher.h:127) or
        'collect' (Red::collect, Red.h:109).  -> (counts uint32[n], spans[n, max])."""
        text = np.ascontiguousarray(text, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        n = len(starts)
        counts = np.zeros(n, dtype=np.uint32)
        spans = np.zeros((n, max_per_line), dtype=SPAN_DTYPE)
        kind = {"match_all": 0, "collect": 1}[which] if isinstance(which, str) else int(which)
        _check(lib().red_b200_multi_batch_host(self._h, text.ctypes.data, text.size, starts.ctypes.data,
                                               lens.ctypes.data, n, kind, max_per_line,
                                               counts.ctypes.data, spans.ctypes.data))
        return counts, spans

    def collect(self, text, cap=256):
        """Red::collect on one text -> list of Outcome."""
        t = np.frombuffer(_bytes(text), dtype=np.uint8)
        c, s = self.multi_batch_host(t, [0], [len(t)], "collect", cap)
        return [Outcome(int(x["result"]), int(x["start"]), int(x["end"])) for x in s[0][:min(int(c[0]), cap)]]

    def matchAll(self, text, cap=256):
        """matchAll on one text -> list of Outcome."""
        t = np.frombuffer(_bytes(text), dtype=np.uint8)
        c, s = self.multi_batch_host(t, [0], [len(t)], "match_all", cap)
        return [Outcome(int(x["result"]), int(x["start"]), int(x["end"])) for x in s[0][:min(int(c[0]), cap)]]

    def replace_batch_host(self, text, starts, lens, repl, max_per_line=(1 << 63), style=styLast,
                           out_delim=-1):
        """replace() per line (Matcher.h:100-125).  -> (out_text uint8[], out_starts uint64[n],
        out_lens uint32[n], counts uint32[n]); every rewritten line is followed by the byte
        `out_delim` unless it is < 0."""
        text = np.ascontiguousarray(text, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        repl = np.frombuffer(_bytes(repl), dtype=np.uint8)
        n = len(starts)
        ostarts = np.zeros(n, dtype=np.uint64)
        olens = np.zeros(n, dtype=np.uint32)
        counts = np.zeros(n, dtype=np.uint32)
        need = C.c_uint64()
        cap = int(lens.sum()) + n + 64
        while True:
            out = np.zeros(max(1, cap), dtype=np.uint8)
            rc = lib().red_b200_replace_batch_host(self._h, text.ctypes.data, text.size, starts.ctypes.data,
                                                   lens.ctypes.data, n, repl.ctypes.data if repl.size else None,
                                                   repl.size, int(max_per_line), _style(style), int(out_delim),
                                                   out.ctypes.data, cap, ostarts.ctypes.data,
                                                   olens.ctypes.data, counts.ctypes.data, C.byref(need))
            if (rc == _E_LIMIT) and (need.value > cap):
                cap = need.value
                continue
            _check(rc)
            return out[:need.value], ostarts, olens, counts

    def replace_batch_device(self, text, starts, lens, repl, max_per_line=(1 << 63), style=styLast,
                             out_delim=-1, stream=None):
        """replace() per line on torch CUDA tensors.  -> (out_text, out_starts, out_lens, counts)
        as CUDA tensors.  Synchronous (the output size has to come back)."""
        import torch
        dev = text.device
        n = starts.numel()
        repl = np.frombuffer(_bytes(repl), dtype=np.uint8)
        ostarts = torch.empty(n, dtype=torch.int64, device=dev)
        olens = torch.empty(n, dtype=torch.int32, device=dev)
        counts = torch.empty(n, dtype=torch.int32, device=dev)
        need = C.c_uint64()
        st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
        cap = int(lens.sum().item()) + n + 64  # enough unless the replacement is longer than the matches
        out = torch.empty(cap, dtype=torch.uint8, device=dev)
        while True:
            rc = lib().red_b200_replace_batch_device(self._h, text.data_ptr(), starts.data_ptr(), lens.data_ptr(), n,
                                                     repl.ctypes.data if repl.size else None, repl.size,
                                                     int(max_per_line), _style(style), int(out_delim),
                                                     out.data_ptr(), cap, ostarts.data_ptr(), olens.data_ptr(),
                                                     counts.data_ptr(), C.byref(need), st or None)
            if (rc == _E_LIMIT) and (need.value > cap):
                cap = need.value
                out = torch.empty(cap, dtype=torch.uint8, device=dev)
                continue
            _check(rc)
            return out[:need.value], ostarts, olens, counts

    def replace(self, text, repl, max=(1 << 63), style=styLast):  # noqa: A002 (reference name)
        """replace(exec, text, repl, out, max, style) -> (count, rewritten text)."""
        t = np.frombuffer(_bytes(text), dtype=np.uint8)
        out, _, olens, counts = self.replace_batch_host(t, [0], [len(t)], repl, max, style)
        return int(counts[0]), out[:int(olens[0])].tobytes()

    def advance_batch_host(self, text, starts, lens, states=None):
        """StatefulMatcher::advance over one segment per stream (Matcher.h:770-794).
        states: uint32[n] from a previous call, or None for fresh matchers.
        -> (states, results int32[n], hit_results int32[n], hit_ends uint32[n])"""
        text = np.ascontiguousarray(text, dtype=np.uint8)
        starts = np.ascontiguousarray(starts, dtype=np.uint64)
        lens = np.ascontiguousarray(lens, dtype=np.uint32)
        n = len(starts)
        st = np.full(n, STATE_FRESH, dtype=np.uint32) if states is None \
            else np.array(states, dtype=np.uint32, copy=True)
        res = np.zeros(n, dtype=np.int32)
        hit = np.zeros(n, dtype=np.int32)
        hend = np.zeros(n, dtype=np.uint32)
        _check(lib().red_b200_advance_batch_host(self._h, text.ctypes.data if text.size else None, text.size,
                                                 starts.ctypes.data, lens.ctypes.data, n, st.ctypes.data,
                                                 res.ctypes.data, hit.ctypes.data, hend.ctypes.data))
        return st, res, hit, hend

    def advance_batch_device(self, text, starts, lens, states, results, hit_results=None,
                             hit_ends=None, stream=None):
        """As advance_batch_host on torch CUDA tensors; `states` (int32/uint32 view) is updated
        in place.  Asynchronous."""
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(text.device).cuda_stream
        _check(lib().red_b200_advance_batch_device(
            self._h, text.data_ptr(), starts.data_ptr(), lens.data_ptr(), starts.numel(),
            states.data_ptr(), results.data_ptr(),
            hit_results.data_ptr() if hit_results is not None else None,
            hit_ends.data_ptr() if hit_ends is not None else None, st or None))

    def grep_host(self, text, delim=b"\n", style=styInstant, capacity=None):
        """Log blob in (uint8 ndarray / bytes), matching lines out.

        -> (records, n_lines): records is a structured array with fields
        line_start, line_len, result, match_start, match_end, in text order."""
        text = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) \
            else np.ascontiguousarray(text, dtype=np.uint8)
        if capacity is None:
            capacity = max(16, text.size // 8)
        out = np.zeros(capacity, dtype=LINE_MATCH_DTYPE)
        nm, nl = C.c_uint64(), C.c_uint64()
        d = delim[0] if isinstance(delim, (bytes, bytearray)) else int(delim)
        _check(lib().red_b200_grep_host(self._h, text.ctypes.data, text.size, d, _style(style),
                                        out.ctypes.data, capacity, C.byref(nm), C.byref(nl)))
        return out[:nm.value], nl.value

    def match_stream_device(self, text, verb=VERB_MATCH, style=styLast, stream=None) -> Outcome:
        """One long torch uint8 CUDA tensor."""
        import torch
        st = stream if stream is not None else torch.cuda.current_stream(text.device).cuda_stream
        oc = _Outcome()
        _check(lib().red_b200_match_stream_device(self._h, text.data_ptr(), text.numel(),
                                                  _verb(verb), _style(style), C.byref(oc), st or None))
        return Outcome(oc.result, oc.start, oc.end)


class StatefulMatcher:
    """Matcher.h:770-794: one DFA state, fed piece by piece.  advance() takes a
    byte (as the reference) or a whole bytes object; result() is the result
    of the current state."""

    def __init__(self, program):
        self.program = program.program if isinstance(program, Red) else program
        self._state = None
        _, res, _, _ = self.program.advance_batch_host(np.zeros(0, np.uint8), [0], [0])
        self._result = int(res[0])

    def advance(self, data):
        if isinstance(data, int):
            data = bytes([data])
        t = np.frombuffer(_bytes(data), dtype=np.uint8)
        st, res, _, _ = self.program.advance_batch_host(t, [0], [len(t)], self._state)
        self._state = st
        self._result = int(res[0])
        return self._result

    def result(self):
        return self._result


def split_lines_device(text, delim=b"\n", capacity=None, stream=None):
    """Line index of a torch uint8 CUDA tensor -> (starts int64, lens int32)."""
    import torch
    dev = text.device
    n = text.numel()
    if capacity is None:
        capacity = n + 1
    starts = torch.empty(capacity, dtype=torch.int64, device=dev)
    lens = torch.empty(capacity, dtype=torch.int32, device=dev)
    cnt = C.c_uint64()
    st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    d = delim[0] if isinstance(delim, (bytes, bytearray)) else int(delim)
    _check(lib().red_b200_split_lines_device(text.data_ptr(), n, d, capacity, starts.data_ptr(),
                                             lens.data_ptr(), C.byref(cnt), dev.index or 0, st or None))
    return starts[:cnt.value], lens[:cnt.value]


def compact_matches_device(results, stream=None):
    """Indices (int64, ascending) of the lines whose result is > 0."""
    import torch
    dev = results.device
    n = results.numel()
    idx = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    st = stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream
    _check(lib().red_b200_compact_matches_device(results.data_ptr(), n, idx.data_ptr(),
                                                 cnt.data_ptr(), dev.index or 0, st or None))
    return idx[:int(cnt.item())]


def compile(parser, fmt=fmtDirectAuto, device=0) -> Program:  # noqa: A001 (reference name)
    """Compile.h:28 - patterns -> uploaded Program."""
    return Program(compileToSerialized(parser, fmt), device=device)


class Red:
    """Mainstream API (Red.h:82-258): one compiled regex or pattern set."""

    def __init__(self, regex=None, flags=0, parser=None, serialized=None, device=0):
        if serialized is not None:
            self.program = Program(serialized, device)
        else:
            if parser is None:
                parser = Parser()
                parser.add(regex, 1, flags)
            self.program = compile(parser, device=device)

    def serialized(self):
        return self.program.serialized

    def save(self, path):
        with open(path, "wb") as f:
            f.write(self.program.serialized)

    @classmethod
    def load(cls, path, device=0):
        with open(path, "rb") as f:
            return cls(serialized=f.read(), device=device)

    def getExec(self):
        return self.program


def _add_orthogonal(cls):
    # match<Style>() / search<Style>(): Red.h:113-132
    for sname, sty in (("Instant", 1), ("First", 2), ("Tangent", 3), ("Last", 4), ("Full", 5)):
        def m(self, text, _s=sty):
            return self.program.match(text, _s)

        def s(self, text, _s=sty):
            return self.program.search(text, _s)
        setattr(cls, "match" + sname, m)
        setattr(cls, "search" + sname, s)


_add_orthogonal(Red)
