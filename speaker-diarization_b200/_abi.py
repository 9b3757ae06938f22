"""ctypes binding of ``libspkdiar.so`` (the C-ABI of ``include/spkdiar.h``).

Thin by design: argument marshalling (numpy arrays -> pointers), error codes ->
``SpkdiarError``, handle lifetime.  No numeric work happens here and there is
no fallback: if the library is not built, or no sm_100-class GPU is present,
every entry point raises.
"""

import ctypes as C
import os
import os.path as op
import threading

import numpy as np

GLR, BIC, KL2 = 0, 1, 2
METRIC = {'GLR': GLR, 'BIC': BIC, 'KL2': KL2}
DIM = 39
RECORD = 820
NPROF = 5
PROF_NAMES = ('stats', 'score', 'gw', 'merge', 'h2d')

LIB_PATH = op.join(op.dirname(op.abspath(__file__)), 'csrc', 'libspkdiar.so')

# every symbol include/spkdiar.h declares (tests check the library exports them)
SYMBOLS = (
    'spkdiar_abi_version', 'spkdiar_create', 'spkdiar_destroy', 'spkdiar_last_error',
    'spkdiar_launch_count', 'spkdiar_sm_count', 'spkdiar_profile_enable',
    'spkdiar_profile_read', 'spkdiar_features_upload', 'spkdiar_features_adopt',
    'spkdiar_stats_build', 'spkdiar_features_free', 'spkdiar_features_frames',
    'spkdiar_stats_window', 'spkdiar_score_windows', 'spkdiar_score_sets',
    'spkdiar_gw_run', 'spkdiar_gw_run_multi', 'spkdiar_cluster_create', 'spkdiar_cluster_run',
    'spkdiar_cluster_run_sharded', 'spkdiar_cluster_run_sharded_nccl', 'spkdiar_nccl_unique_id',
    'spkdiar_cluster_run_sharded_p2p', 'spkdiar_mailbox_create', 'spkdiar_mailbox_open',
    'spkdiar_mailbox_close', 'spkdiar_mailbox_free',
    'spkdiar_cluster_free', 'spkdiar_cluster_matrix', 'spkdiar_cluster_rowlog', 'spkdiar_cluster_counters',
    'spkdiar_features_upload_batch', 'spkdiar_cluster_batch', 'spkdiar_selftest_stitch',
    'spkdiar_gw_multi_begin', 'spkdiar_gw_multi_wait', 'spkdiar_gw_multi_where', 'spkdiar_gw_multi_end',
    'spkdiar_ctx_exec',
    'spkdiar_replay_create', 'spkdiar_replay_free', 'spkdiar_replay_error', 'spkdiar_replay_info',
    'spkdiar_replay_chains', 'spkdiar_replay_segment', 'spkdiar_replay_turns', 'spkdiar_replay_cluster',
    'spkdiar_replay_text', 'spkdiar_features_upload_frames', 'spkdiar_cluster_inorder',
    'spkdiar_merge_chain',
)


class SpkdiarError(RuntimeError):
    def __init__(self, code, text):
        RuntimeError.__init__(self, 'libspkdiar error %d: %s' % (code, text))
        self.code = code


class GwParams(C.Structure):
    _fields_ = [('rate', C.c_double), ('winsize', C.c_double), ('winstep', C.c_double),
                ('deltaws', C.c_double), ('threshold', C.c_double), ('lambda_', C.c_double),
                ('metric', C.c_int32), ('max_groups', C.c_int32)]


GW_WINDOW_DTYPE = np.dtype([('start', '<f8'), ('end', '<f8'), ('maxi', '<f8'), ('maxd', '<f8'),
                            ('maxi_fine', '<f8'), ('maxd_fine', '<f8'), ('positive', '<i4'),
                            ('chain', '<i4'), ('ncand', '<i4'), ('ninf', '<i4'), ('seq', '<i4'),
                            ('pad', '<i4')])
MERGE_DTYPE = np.dtype([('a', '<i4'), ('b', '<i4'), ('d', '<f8')])
assert GW_WINDOW_DTYPE.itemsize == 72 and MERGE_DTYPE.itemsize == 16

EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)

_lib = None


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def load_library(path=None):
    """dlopen the library and declare its prototypes.  Raises ``SpkdiarError``
    when it has not been built (``python -c 'import __graft_entry__ as g;
    g.build()'``)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get('SPKDIAR_LIB', LIB_PATH)
    if not op.isfile(path):
        raise SpkdiarError(-6, 'libspkdiar.so not found at %s - build it first '
                           '(__graft_entry__.build()); there is no CPU fallback' % path)
    lib = C.CDLL(path)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    pi64, pdbl = C.POINTER(C.c_int64), C.POINTER(C.c_double)
    proto = {
        'spkdiar_abi_version': (C.c_int, []),
        'spkdiar_create': (C.c_int, [C.c_int, vp, C.POINTER(vp)]),
        'spkdiar_destroy': (None, [vp]),
        'spkdiar_last_error': (C.c_char_p, [vp]),
        'spkdiar_launch_count': (i64, [vp]),
        'spkdiar_sm_count': (C.c_int, [vp]),
        'spkdiar_profile_enable': (C.c_int, [vp, C.c_int]),
        'spkdiar_profile_read': (C.c_int, [vp, pdbl, pi64]),
        'spkdiar_features_upload': (C.c_int, [vp, vp, i64, i32, C.POINTER(vp)]),
        'spkdiar_features_adopt': (C.c_int, [vp, vp, i64, i32, C.POINTER(vp)]),
        'spkdiar_features_upload_frames': (C.c_int, [vp, vp, i64, i32, C.POINTER(vp)]),
        'spkdiar_stats_build': (C.c_int, [vp]),
        'spkdiar_features_free': (C.c_int, [vp]),
        'spkdiar_features_frames': (i64, [vp]),
        'spkdiar_stats_window': (C.c_int, [vp, i64, i64, pdbl, pdbl]),
        'spkdiar_score_windows': (C.c_int, [vp, pi64, pi64, pi64, i64, C.c_int, dbl, pdbl, pdbl]),
        'spkdiar_score_sets': (C.c_int, [vp, i64, pi64, pi64, pi64, pi64, pi64, pi64, C.c_int, dbl,
                                         pdbl, pdbl]),
        'spkdiar_gw_run': (C.c_int, [vp, C.POINTER(GwParams), pi64, pi64, i32, vp, i64, pi64]),
        'spkdiar_gw_run_multi': (C.c_int, [vp, i32, C.POINTER(GwParams), pi64, pi64, i32, C.POINTER(vp), pi64,
                                           C.POINTER(pi64)]),
        'spkdiar_cluster_create': (C.c_int, [vp, pi64, pi64, i64, C.c_int, dbl, C.POINTER(vp)]),
        'spkdiar_cluster_run': (C.c_int, [vp, dbl, i32, i32, vp, i64, pi64, pdbl]),
        'spkdiar_cluster_run_sharded': (C.c_int, [vp, dbl, i32, i32, i32, EXCHANGE_FN, vp, vp, i64,
                                                  pi64, pdbl]),
        'spkdiar_cluster_run_sharded_nccl': (C.c_int, [vp, dbl, i32, i32, i32, vp, vp, i64, pi64, pdbl]),
        'spkdiar_nccl_unique_id': (C.c_int, [vp]),
        'spkdiar_cluster_run_sharded_p2p': (C.c_int, [vp, dbl, i32, i32, i32, C.POINTER(vp), C.c_uint64, vp, i64, pi64,
                                                      pdbl]),
        'spkdiar_mailbox_create': (C.c_int, [vp, i32, C.POINTER(vp), vp]),
        'spkdiar_mailbox_open': (C.c_int, [vp, vp, C.POINTER(vp)]),
        'spkdiar_mailbox_close': (C.c_int, [vp, vp]),
        'spkdiar_mailbox_free': (C.c_int, [vp, vp]),
        'spkdiar_cluster_free': (C.c_int, [vp]),
        'spkdiar_cluster_matrix': (C.c_int, [vp, pdbl, C.POINTER(C.c_uint8)]),
        'spkdiar_cluster_rowlog': (C.c_int, [vp, pdbl, i64]),
        'spkdiar_cluster_counters': (C.c_int, [vp, C.POINTER(C.c_uint64)]),
        'spkdiar_features_upload_batch': (C.c_int, [vp, C.POINTER(vp), pi64, i32, i32, C.POINTER(vp), pi64]),
        'spkdiar_cluster_batch': (C.c_int, [vp, i32, pi64, pi64, pi64, C.c_int, dbl, dbl, i32, i32, vp, pi64, pdbl]),
        'spkdiar_selftest_stitch': (C.c_int, [C.c_uint64, dbl, i64, i32]),
        'spkdiar_gw_multi_begin': (C.c_int, [vp, i32, C.POINTER(GwParams), pi64, pi64, i32, C.POINTER(vp)]),
        'spkdiar_gw_multi_wait': (C.c_int, [vp, i32, vp, i64, pi64]),
        'spkdiar_gw_multi_where': (C.c_int, [vp, i32, C.POINTER(vp), C.POINTER(i32)]),
        'spkdiar_gw_multi_end': (C.c_int, [vp]),
        'spkdiar_ctx_exec': (C.c_int, [vp, vp, i32]),
        'spkdiar_cluster_inorder': (C.c_int, [vp, i32, pi64, pi64, pi64, i64, pi64, pi64, C.c_int, dbl, dbl, pdbl, i64, pi64,
                                              C.POINTER(C.c_int32)]),
        'spkdiar_merge_chain': (C.c_int, [vp, i64, pi64, pi64, C.c_int, dbl, dbl, i32, pdbl, pdbl, pdbl, C.POINTER(C.c_int32)]),
        'spkdiar_replay_create': (C.c_int, [dbl, C.c_char_p, i64, C.POINTER(vp)]),
        'spkdiar_replay_free': (C.c_int, [vp]),
        'spkdiar_replay_error': (C.c_char_p, [vp]),
        'spkdiar_replay_info': (C.c_int, [vp, pi64]),
        'spkdiar_replay_chains': (C.c_int, [vp, i64, i64, pi64, pi64, i64]),
        'spkdiar_replay_segment': (C.c_int, [vp, vp, pi64]),
        'spkdiar_replay_turns': (C.c_int, [vp, i64, i64, pi64, pi64, i64]),
        'spkdiar_replay_cluster': (C.c_int, [vp, vp, i64]),
        'spkdiar_replay_text': (vp, [vp, i32, pi64]),
    }
    for name, (res, args) in proto.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.spkdiar_abi_version() != 1:
        raise SpkdiarError(-5, 'ABI version mismatch: %d' % lib.spkdiar_abi_version())
    if path == os.environ.get('SPKDIAR_LIB', LIB_PATH):
        _lib = lib
    return lib


def nccl_unique_id():
    """128 bytes identifying a new NCCL communicator (call on rank 0, send to every rank)."""
    lib = load_library()
    buf = C.create_string_buffer(128)
    rc = lib.spkdiar_nccl_unique_id(buf)
    if rc != 0:
        raise SpkdiarError(rc, 'NCCL is not available (libnccl.so.2 could not be loaded)')
    return buf.raw


class ReplayUnsupported(Exception):
    """The native replay does not reproduce this recording to the byte (``SPKDIAR_E_UNSUPPORTED``):
    run the general Python replay (``Detector`` / ``Clusterer``) instead."""


class Replay(object):
    """``spkdiar_replay``: one recording through the text side of spk-diarization2.py:122-128 in
    native code - recipe text in, chains for the growing-window search, window records ->
    segmentation recipe, initial clusters, merge sequence -> clustered recipe.  Host only (no
    context, no device); the numeric results come from ``Features.gw_run`` / ``cluster_batch``."""

    def __init__(self, rate, lines):
        """``lines``: the recipe as a list of text lines (as ``readlines`` returns them)."""
        self.lib = load_library()
        self.h = None
        text = ''.join(lines)
        # an element that is not exactly one physical line is one "line" to the reference's searches
        if text.count('\n') != sum(1 for l in lines if l.endswith('\n')) or \
                any(not l.endswith('\n') for l in lines[:-1]):
            raise ReplayUnsupported('recipe elements are not single lines')
        try:
            raw = text.encode('ascii')
        except UnicodeEncodeError:
            raise ReplayUnsupported('non-ASCII recipe text')
        h = C.c_void_p()
        rc = self.lib.spkdiar_replay_create(float(rate), raw, len(raw), C.byref(h))
        self.h = h if h.value else None
        self._check(rc)
        info = self.info()
        self.nlines, self.nchains, self.single_wav = int(info[0]), int(info[1]), bool(info[2])

    def _check(self, rc):
        if rc == 0:
            return
        text = (self.lib.spkdiar_replay_error(self.h) or b'').decode() if self.h else 'no handle'
        if rc == -5:
            raise ReplayUnsupported(text)
        raise SpkdiarError(rc, text)

    def info(self):
        out = np.zeros(6, dtype=np.int64)
        self._check(self.lib.spkdiar_replay_info(self.h, _p(out, C.c_int64)))
        return out

    def chains(self, nframes, base=0, seg_a=None, seg_b=None):
        """-> (seg_a, seg_b) of this recording's chains, positions offset by ``base`` (the
        recording's first row in a packed batch); writes into the given int64 views if any."""
        if seg_a is None:
            seg_a, seg_b = np.zeros(self.nchains, dtype=np.int64), np.zeros(self.nchains, dtype=np.int64)
        self._check(self.lib.spkdiar_replay_chains(self.h, int(nframes), int(base), _p(seg_a, C.c_int64),
                                                   _p(seg_b, C.c_int64), seg_a.shape[0]))
        return seg_a, seg_b

    def segment(self, win, first):
        """``win``: window records (GW_WINDOW_DTYPE array), ``first``: int64 array with this
        recording's nchains + 1 offsets into ``win``.  -> number of turns written."""
        first = np.ascontiguousarray(first, dtype=np.int64)
        if first.shape[0] != self.nchains + 1:
            raise SpkdiarError(-2, 'win_first needs %d entries' % (self.nchains + 1))
        self._check(self.lib.spkdiar_replay_segment(self.h, win.ctypes.data_as(C.c_void_p), _p(first, C.c_int64)))
        return int(self.info()[3])

    def turns(self, nframes, nturns, base=0, seg_a=None, seg_b=None):
        if seg_a is None:
            seg_a, seg_b = np.zeros(nturns, dtype=np.int64), np.zeros(nturns, dtype=np.int64)
        self._check(self.lib.spkdiar_replay_turns(self.h, int(nframes), int(base), _p(seg_a, C.c_int64),
                                                  _p(seg_b, C.c_int64), seg_a.shape[0]))
        return seg_a, seg_b

    def cluster(self, merges):
        """``merges``: MERGE_DTYPE array (compacted indices).  -> number of speakers."""
        merges = np.ascontiguousarray(merges, dtype=MERGE_DTYPE)
        self._check(self.lib.spkdiar_replay_cluster(self.h, merges.ctypes.data_as(C.c_void_p), merges.shape[0]))
        return int(self.info()[4])

    def text(self, which):
        n = C.c_int64(0)
        ptr = self.lib.spkdiar_replay_text(self.h, int(which), C.byref(n))
        return C.string_at(ptr, n.value).decode('ascii') if ptr else ''

    def close(self):
        if getattr(self, 'h', None):
            self.lib.spkdiar_replay_free(self.h)
        self.h = None

    def __del__(self):
        self.close()


class _Serialised(object):
    """``include/spkdiar.h``: "calls on one context are serialised by the caller".  This proxy
    is that caller: every entry point reached through ``ctx.lib`` takes the context's lock, so a
    worker thread may queue device work (ctypes releases the GIL inside the call) while another
    thread replays records on the host (``corpus.run_corpus(..., overlap=True)``)."""

    def __init__(self, lib):
        self._lib = lib
        self._lock = threading.RLock()

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        lock = self._lock

        def call(*args):
            with lock:
                return fn(*args)
        setattr(self, name, call)
        return call


class Context(object):
    """One device context (``spkdiar_ctx``).  ``stream``: an existing
    ``cudaStream_t`` as an int (e.g. ``torch.cuda.current_stream().cuda_stream``)
    or None."""

    def __init__(self, device=0, stream=None):
        self.lib = _Serialised(load_library())
        h = C.c_void_p()
        rc = self.lib.spkdiar_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise SpkdiarError(rc, (self.lib.spkdiar_last_error(None) or b'').decode())
        self.h = h
        self.device = int(device)

    def _check(self, rc):
        if rc != 0:
            raise SpkdiarError(rc, (self.lib.spkdiar_last_error(self.h) or b'').decode())

    def close(self):
        for extra in getattr(self, '_lanes', []):
            extra.close()
        self._lanes = []
        if getattr(self, 'h', None):
            self.lib.spkdiar_destroy(self.h)
            self.h = None

    def lane_contexts(self, n):
        """``n`` further contexts (own streams) on this context's device, created on first use and
        closed with it: the extra lanes of ``corpus.diarize_batches``."""
        lanes = getattr(self, '_lanes', [])
        while len(lanes) < n:
            lanes.append(Context(self.device))
        self._lanes = lanes
        return lanes[:n]

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def launches(self):
        """Kernels launched on this context and on its lane contexts."""
        return int(self.lib.spkdiar_launch_count(self.h)) + sum(c.launches for c in getattr(self, '_lanes', []))

    @property
    def sm_count(self):
        return int(self.lib.spkdiar_sm_count(self.h))

    def profile(self, enable=True):
        self._check(self.lib.spkdiar_profile_enable(self.h, 1 if enable else 0))

    def profile_read(self):
        """-> {class: (ms, launches)}"""
        ms = np.zeros(NPROF)
        n = np.zeros(NPROF, dtype=np.int64)
        self._check(self.lib.spkdiar_profile_read(self.h, _p(ms, C.c_double), _p(n, C.c_int64)))
        return {PROF_NAMES[k]: (float(ms[k]), int(n[k])) for k in range(NPROF)}

    def exec_on(self, stream=None, sms=0):
        """Later calls run on ``stream`` (int) with at most ``sms`` SMs; no arguments restore the
        context's own stream and the whole device (``spkdiar_ctx_exec``)."""
        self._check(self.lib.spkdiar_ctx_exec(self.h, C.c_void_p(stream) if stream else None, int(sms)))

    def upload(self, frames):
        """(n, 39) float32 host matrix -> Features (copied to HBM, statistics built)."""
        frames = np.ascontiguousarray(frames, dtype=np.float32)
        if frames.ndim != 2:
            raise ValueError('frames must be (n, dim)')
        h = C.c_void_p()
        self._check(self.lib.spkdiar_features_upload(self.h, frames.ctypes.data_as(C.c_void_p),
                                                     frames.shape[0], frames.shape[1], C.byref(h)))
        return Features(self, h, frames.shape[0], frames.shape[1])

    def upload_frames(self, frames):
        """(n, 39) float32 host matrix -> Features WITHOUT window statistics (156 B per frame
        resident instead of 6,716): for clustering on its own.  Cluster records then come
        straight from the frames; the first call that scores windows builds the statistics."""
        frames = np.ascontiguousarray(frames, dtype=np.float32)
        if frames.ndim != 2:
            raise ValueError('frames must be (n, dim)')
        h = C.c_void_p()
        self._check(self.lib.spkdiar_features_upload_frames(self.h, frames.ctypes.data_as(C.c_void_p),
                                                            frames.shape[0], frames.shape[1], C.byref(h)))
        return Features(self, h, frames.shape[0], frames.shape[1])

    def upload_ptr(self, host_ptr, n, dim=DIM):
        """Same from a raw host pointer (e.g. pinned torch memory)."""
        h = C.c_void_p()
        self._check(self.lib.spkdiar_features_upload(self.h, C.c_void_p(host_ptr), n, dim, C.byref(h)))
        return Features(self, h, n, dim)

    def upload_batch(self, matrices):
        """A batch of recordings -> FeaturePack (one packed handle, statistics restarting per
        recording).  ``matrices``: (n_r, 39) float32 arrays, or (host pointer, n_r) pairs
        (e.g. pinned torch memory)."""
        keep, ptrs, ns = [], [], []
        dim = None                       # (host pointer, n) pairs carry DIM-dimensional rows
        for m in matrices:
            if isinstance(m, tuple):
                ptrs.append(int(m[0]))
                ns.append(int(m[1]))
            else:
                m = np.ascontiguousarray(m, dtype=np.float32)
                if m.ndim != 2 or not 1 <= m.shape[1] <= DIM:
                    raise SpkdiarError(-5, 'feature matrix of shape %r: the kernels handle 1..%d dimensions' % (m.shape, DIM))
                if dim is not None and m.shape[1] != dim:
                    raise SpkdiarError(-2, 'recordings of %d and %d dimensions in one batch' % (dim, m.shape[1]))
                dim = m.shape[1]
                keep.append(m)
                ptrs.append(m.ctypes.data)
                ns.append(m.shape[0])
        nrec = len(ptrs)
        arr = (C.c_void_p * max(nrec, 1))(*[C.c_void_p(p) for p in ptrs])
        n = np.array(ns, dtype=np.int64)
        base = np.zeros(max(nrec, 1), dtype=np.int64)
        h = C.c_void_p()
        self._check(self.lib.spkdiar_features_upload_batch(self.h, arr, _p(n, C.c_int64), nrec, dim or DIM,
                                                           C.byref(h), _p(base, C.c_int64)))
        return FeaturePack(self, h, base[:nrec].tolist(), ns, dim or DIM)

    def adopt(self, dev_ptr, n, dim=DIM):
        """Features over a matrix already in device memory (not copied)."""
        h = C.c_void_p()
        self._check(self.lib.spkdiar_features_adopt(self.h, C.c_void_p(dev_ptr), n, dim, C.byref(h)))
        return Features(self, h, n, dim)


class Features(object):
    """``spkdiar_feat``: frames + prefix statistics resident in HBM."""

    def __init__(self, ctx, handle, n, dim=DIM):
        self.ctx = ctx
        self.h = handle
        self.n = int(n)
        self.dim = int(dim)              # of the feature file (<= 39; the device keeps 39 zero-padded columns)
        ctx.live_features = getattr(ctx, 'live_features', 0) + 1
        ctx.peak_features = max(getattr(ctx, 'peak_features', 0), ctx.live_features)

    def close(self):
        if getattr(self, 'h', None):
            if self.ctx.h:
                self.ctx.lib.spkdiar_features_free(self.h)
            self.ctx.live_features -= 1
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def rebuild_stats(self):
        self.ctx._check(self.ctx.lib.spkdiar_stats_build(self.h))

    def stats_window(self, a, b):
        """-> (sum[39], packed lower-triangular second moments [780], shift[39])"""
        out = np.zeros(819)
        shift = np.zeros(39)
        self.ctx._check(self.ctx.lib.spkdiar_stats_window(self.h, int(a), int(b), _p(out, C.c_double),
                                                          _p(shift, C.c_double)))
        return out[:39], out[39:], shift

    def score_windows(self, a, m, b, metric, lambdac=1.3, terms=False):
        a, m, b = _i64(a), _i64(m), _i64(b)
        n = a.shape[0]
        d = np.empty(n)
        t = np.empty((n, 3)) if terms else None
        self.ctx._check(self.ctx.lib.spkdiar_score_windows(
            self.h, _p(a, C.c_int64), _p(m, C.c_int64), _p(b, C.c_int64), n, int(metric),
            float(lambdac), _p(d, C.c_double), _p(t, C.c_double) if terms else None))
        return (d, t) if terms else d

    def score_sets(self, sets1, sets2, metric, lambdac=1.3, terms=False):
        """sets1[p], sets2[p]: lists of (a, b) frame ranges; one distance per p."""
        def pack(sets):
            off = np.zeros(len(sets) + 1, dtype=np.int64)
            for k, s in enumerate(sets):
                off[k + 1] = off[k] + len(s)
            flat = [r for s in sets for r in s]
            ra = np.array([r[0] for r in flat], dtype=np.int64)
            rb = np.array([r[1] for r in flat], dtype=np.int64)
            return off, ra, rb
        n = len(sets1)
        assert len(sets2) == n
        o1, a1, b1 = pack(sets1)
        o2, a2, b2 = pack(sets2)
        d = np.empty(n)
        t = np.empty((n, 3)) if terms else None
        self.ctx._check(self.ctx.lib.spkdiar_score_sets(
            self.h, n, _p(o1, C.c_int64), _p(a1, C.c_int64), _p(b1, C.c_int64),
            _p(o2, C.c_int64), _p(a2, C.c_int64), _p(b2, C.c_int64), int(metric), float(lambdac),
            _p(d, C.c_double), _p(t, C.c_double) if terms else None))
        return (d, t) if terms else d

    def merge_chain(self, seg_a, seg_b, metric, lambdac, threshold, use_memo, memo_c1):
        """``merge_rec`` over the lines of one wav in ONE launch.  ``memo_c1``: None or the memo of
        the reference's first left BIC term.  -> (terms [n - 1][3], dist [n - 1], merged [n - 1],
        memo_c1 after the chain)."""
        seg_a, seg_b = _i64(seg_a), _i64(seg_b)
        n = seg_a.shape[0]
        terms = np.zeros((max(n - 1, 1), 3))
        dist = np.zeros(max(n - 1, 1))
        merged = np.zeros(max(n - 1, 1), dtype=np.int32)
        memo = np.array([np.nan if memo_c1 is None else float(memo_c1)])
        self.ctx._check(self.ctx.lib.spkdiar_merge_chain(
            self.h, n, _p(seg_a, C.c_int64), _p(seg_b, C.c_int64), int(metric), float(lambdac), float(threshold),
            1 if use_memo else 0, _p(memo, C.c_double), _p(terms, C.c_double), _p(dist, C.c_double),
            merged.ctypes.data_as(C.POINTER(C.c_int32))))
        return terms[:max(n - 1, 0)], dist[:max(n - 1, 0)], merged[:max(n - 1, 0)], (None if np.isnan(memo[0]) else memo[0])

    def cluster_inorder(self, speakers, seg_a, seg_b, metric, lambdac, threshold):
        """``spk_cluster_in`` over the lines [seg_a[l], seg_b[l]) in ONE launch.  ``speakers``: the
        range sets [[(a, b), ...], ...] of the speakers that exist before the first line.
        -> (dist, first, best): the distances of line l to the speakers that existed then are
        dist[first[l]:first[l + 1]], best[l] is the speaker joined or -1 for a new one."""
        seg_a, seg_b = _i64(seg_a), _i64(seg_b)
        nlines = seg_a.shape[0]
        off = np.zeros(len(speakers) + 1, dtype=np.int64)
        ra, rb = [], []
        for s, ranges in enumerate(speakers):
            for a, b in ranges:
                ra.append(a)
                rb.append(b)
            off[s + 1] = len(ra)
        ra, rb = _i64(ra), _i64(rb)
        # line l sees at most len(speakers) + l speakers; usually far fewer: start small, grow on demand
        full = nlines * len(speakers) + (nlines * (nlines - 1)) // 2 + 1
        cap = min(full, max(nlines * (len(speakers) + 64), 1))
        first = np.zeros(nlines + 1, dtype=np.int64)
        best = np.zeros(max(nlines, 1), dtype=np.int32)
        while True:
            dist = np.empty(cap)
            rc = self.ctx.lib.spkdiar_cluster_inorder(
                self.h, len(speakers), _p(off, C.c_int64), _p(ra, C.c_int64), _p(rb, C.c_int64), nlines,
                _p(seg_a, C.c_int64), _p(seg_b, C.c_int64), int(metric), float(lambdac), float(threshold),
                _p(dist, C.c_double), cap, _p(first, C.c_int64), best.ctypes.data_as(C.POINTER(C.c_int32)))
            if rc == -4 and cap < full:                 # SPKDIAR_E_CAPACITY
                cap = full
                continue
            self.ctx._check(rc)
            return dist[:int(first[nlines])], first, best[:nlines]

    def gw_run(self, seg_a, seg_b, rate, winsize, winstep, deltaws, threshold, lambdac, metric,
               max_groups=0, cap=None):
        """-> (window records [structured array], win_first [nchain + 1])"""
        seg_a, seg_b = _i64(seg_a), _i64(seg_b)
        nchain = seg_a.shape[0]
        prm = GwParams(float(rate), float(winsize), float(winstep), float(deltaws), float(threshold),
                       float(lambdac), int(metric), int(max_groups))
        if cap is None:
            unit = max(min(float(rate) / 2 - float(rate) / 10, float(winstep)), 1.0)
            cap = int(sum(2 * ((b - a) / unit + 2) for a, b in zip(seg_a, seg_b))) + 16
        first = np.zeros(nchain + 1, dtype=np.int64)
        while True:
            win = np.zeros(cap, dtype=GW_WINDOW_DTYPE)
            rc = self.ctx.lib.spkdiar_gw_run(self.h, C.byref(prm), _p(seg_a, C.c_int64), _p(seg_b, C.c_int64),
                                             nchain, win.ctypes.data_as(C.c_void_p), cap, _p(first, C.c_int64))
            if rc == -4 and first[0] > cap:            # SPKDIAR_E_CAPACITY: retry with the needed size
                cap = int(first[0])
                continue
            self.ctx._check(rc)
            return win[:int(first[nchain])], first

    def gw_run_multi(self, seg_a, seg_b, runs, max_groups=0):
        """Several growing-window searches over the same chains, side by side on disjoint SM
        subsets.  ``runs``: list of dicts with the keyword arguments of ``gw_run`` (rate, winsize,
        winstep, deltaws, threshold, lambdac, metric).  -> list of (records, win_first)."""
        seg_a, seg_b = _i64(seg_a), _i64(seg_b)
        nchain = seg_a.shape[0]
        nrun = len(runs)
        prm = (GwParams * nrun)(*[GwParams(float(r['rate']), float(r['winsize']), float(r['winstep']),
                                           float(r['deltaws']), float(r['threshold']), float(r['lambdac']),
                                           int(r['metric']), int(max_groups)) for r in runs])
        caps = np.zeros(nrun, dtype=np.int64)
        for k, r in enumerate(runs):
            unit = max(min(float(r['rate']) / 2 - float(r['rate']) / 10, float(r['winstep'])), 1.0)
            caps[k] = int(sum(2 * ((b - a) / unit + 2) for a, b in zip(seg_a, seg_b))) + 16
        firsts = [np.zeros(nchain + 1, dtype=np.int64) for _ in range(nrun)]
        while True:
            wins = [np.zeros(int(caps[k]), dtype=GW_WINDOW_DTYPE) for k in range(nrun)]
            wp = (C.c_void_p * nrun)(*[w.ctypes.data for w in wins])
            fp = (C.POINTER(C.c_int64) * nrun)(*[_p(f, C.c_int64) for f in firsts])
            rc = self.ctx.lib.spkdiar_gw_run_multi(self.h, nrun, prm, _p(seg_a, C.c_int64), _p(seg_b, C.c_int64),
                                                   nchain, wp, _p(caps, C.c_int64), fp)
            if rc == -4:                                # SPKDIAR_E_CAPACITY: retry with the needed sizes
                grown = False
                for k in range(nrun):
                    if firsts[k][0] > caps[k]:
                        caps[k] = int(firsts[k][0])
                        grown = True
                if grown:
                    continue
            self.ctx._check(rc)
            return [(wins[k][:int(firsts[k][nchain])], firsts[k]) for k in range(nrun)]

    def gw_multi_begin(self, seg_a, seg_b, runs, max_groups=0):
        """Asynchronous form of ``gw_run_multi``: -> GwMulti (``wait(r)`` per search, ``close()``)."""
        return GwMulti(self, seg_a, seg_b, runs, max_groups)

    def cluster(self, seg_a, seg_b, metric, lambdac=1.3):
        seg_a, seg_b = _i64(seg_a), _i64(seg_b)
        h = C.c_void_p()
        self.ctx._check(self.ctx.lib.spkdiar_cluster_create(
            self.h, _p(seg_a, C.c_int64), _p(seg_b, C.c_int64), seg_a.shape[0], int(metric),
            float(lambdac), C.byref(h)))
        return Clusters(self, h, seg_a.shape[0])


class GwMulti(object):
    """``spkdiar_gwm``: several growing-window searches over the same chains, launched together;
    ``wait(r)`` collects search r while the others keep running, ``where(r)`` names the stream
    and SM share it ran on (for ``Context.exec_on``: queue the clustering of its turns there)."""

    def __init__(self, feat, seg_a, seg_b, runs, max_groups=0):
        self.ctx = feat.ctx
        self.seg_a, self.seg_b = _i64(seg_a), _i64(seg_b)
        self.nchain = self.seg_a.shape[0]
        self.runs = runs
        prm = (GwParams * len(runs))(*[GwParams(float(r['rate']), float(r['winsize']), float(r['winstep']),
                                                float(r['deltaws']), float(r['threshold']), float(r['lambdac']),
                                                int(r['metric']), int(max_groups)) for r in runs])
        h = C.c_void_p()
        self.ctx._check(self.ctx.lib.spkdiar_gw_multi_begin(feat.h, len(runs), prm, _p(self.seg_a, C.c_int64),
                                                            _p(self.seg_b, C.c_int64), self.nchain, C.byref(h)))
        self.h = h

    def wait(self, r):
        """-> (window records, win_first) of search r, as ``Features.gw_run`` returns them."""
        if not 0 <= int(r) < len(self.runs):
            raise SpkdiarError(-2, 'search %d of %d' % (r, len(self.runs)))
        unit = max(min(float(self.runs[r]['rate']) / 2 - float(self.runs[r]['rate']) / 10,
                       float(self.runs[r]['winstep'])), 1.0)
        cap = int(sum(2 * ((b - a) / unit + 2) for a, b in zip(self.seg_a, self.seg_b))) + 16
        first = np.zeros(self.nchain + 1, dtype=np.int64)
        while True:
            win = np.zeros(cap, dtype=GW_WINDOW_DTYPE)
            rc = self.ctx.lib.spkdiar_gw_multi_wait(self.h, int(r), win.ctypes.data_as(C.c_void_p), cap,
                                                    _p(first, C.c_int64))
            if rc == -4 and first[0] > cap:
                cap = int(first[0])
                continue
            self.ctx._check(rc)
            return win[:int(first[self.nchain])], first

    def where(self, r):
        st, sms = C.c_void_p(), C.c_int32(0)
        self.ctx._check(self.ctx.lib.spkdiar_gw_multi_where(self.h, int(r), C.byref(st), C.byref(sms)))
        return st.value, int(sms.value)

    def close(self):
        if getattr(self, 'h', None) and self.ctx.h:
            self.ctx.lib.spkdiar_gw_multi_end(self.h)
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class FeaturePack(Features):
    """A batch of recordings in one handle (``spkdiar_features_upload_batch``).  Positions on
    the pack itself are packed frame rows; ``view(r)`` is recording r with its own positions."""

    def __init__(self, ctx, handle, base, n, dim=DIM):
        Features.__init__(self, ctx, handle, (base[-1] + (n[-1] // 128 + 1) * 128) if n else 0, dim)
        self.base = list(base)
        self.lengths = list(n)

    def __len__(self):
        return len(self.base)

    def view(self, r):
        return FeatureView(self, r)

    def gw_run_batch(self, chains, rate, winsize, winstep, deltaws, threshold, lambdac, metric):
        """``chains[r]``: list of (a, b) frame ranges of recording r (its recipe lines).  All
        chains of all recordings run in ONE launch (one CTA per chain when there are at least
        as many chains as SMs).  -> per recording (records, win_first) exactly as
        ``view(r).gw_run`` returns them."""
        seg_a, seg_b, owner = [], [], [0]
        for r, ch in enumerate(chains):
            for a, b in ch:
                seg_a.append(self.base[r] + int(a))
                seg_b.append(self.base[r] + int(b))
            owner.append(len(seg_a))
        win, first = self.gw_run(seg_a, seg_b, rate, winsize, winstep, deltaws, threshold, lambdac, metric)
        out = []
        for r in range(len(chains)):
            c0, c1 = owner[r], owner[r + 1]
            recs = win[int(first[c0]):int(first[c1])].copy()
            recs['chain'] -= c0
            out.append((recs, first[c0:c1 + 1] - first[c0]))
        return out

    def cluster_batch(self, problems, metric, lambdac=1.3, threshold=0.0, max_spk=0, variant=1):
        """``problems[r]``: list of (a, b) frame ranges of recording r, the initial clusters of
        one ``spk_cluster_hi`` run.  One launch, one CTA per problem.  -> per problem
        (merges, stats[4]) exactly as ``view(r).cluster(...).run(...)`` returns them."""
        rows = []
        for r, segs in enumerate(problems):
            rows.append((_i64([self.base[r] + int(a) for a, _ in segs]), _i64([self.base[r] + int(b) for _, b in segs])))
        return self.cluster_batch_rows(rows, metric, lambdac, threshold, max_spk, variant)

    def cluster_batch_rows(self, problems, metric, lambdac=1.3, threshold=0.0, max_spk=0, variant=1):
        """The same with ``problems[p]`` = (seg_a, seg_b) int64 arrays in PACKED frame rows (any
        recording of the pack, in any order)."""
        first = np.zeros(len(problems) + 1, dtype=np.int64)
        for p, (sa, _) in enumerate(problems):
            first[p + 1] = first[p] + len(sa)
        seg_a = _i64(np.concatenate([sa for sa, _ in problems])) if problems else np.zeros(0, dtype=np.int64)
        seg_b = _i64(np.concatenate([sb for _, sb in problems])) if problems else np.zeros(0, dtype=np.int64)
        out = np.zeros(max(len(seg_a), 1), dtype=MERGE_DTYPE)
        nm = np.zeros(max(len(problems), 1), dtype=np.int64)
        stats = np.zeros((max(len(problems), 1), 4))
        self.ctx._check(self.ctx.lib.spkdiar_cluster_batch(
            self.h, len(problems), _p(first, C.c_int64), _p(seg_a, C.c_int64), _p(seg_b, C.c_int64),
            int(metric), float(lambdac), float(threshold), int(max_spk), int(variant),
            out.ctypes.data_as(C.c_void_p), _p(nm, C.c_int64), _p(stats, C.c_double)))
        return [(out[int(first[r]):int(first[r]) + int(nm[r])], stats[r]) for r in range(len(problems))]


class FeatureView(object):
    """Recording r of a FeaturePack with the interface of ``Features`` (positions are the
    recording's own frame indices).  Closing a view does nothing: the pack owns the memory."""

    def __init__(self, pack, r):
        self.pack = pack
        self.ctx = pack.ctx
        self.r = int(r)
        self.off = int(pack.base[r])
        self.n = int(pack.lengths[r])
        self.dim = getattr(pack, 'dim', DIM)

    def close(self):
        pass

    def _sh(self, a):
        return _i64(a) + self.off

    def stats_window(self, a, b):
        return self.pack.stats_window(int(a) + self.off, int(b) + self.off)

    def score_windows(self, a, m, b, metric, lambdac=1.3, terms=False):
        return self.pack.score_windows(self._sh(a), self._sh(m), self._sh(b), metric, lambdac, terms)

    def score_sets(self, sets1, sets2, metric, lambdac=1.3, terms=False):
        sh = lambda sets: [[(a + self.off, b + self.off) for a, b in s] for s in sets]
        return self.pack.score_sets(sh(sets1), sh(sets2), metric, lambdac, terms)

    def gw_run(self, seg_a, seg_b, *args, **kw):
        return self.pack.gw_run(self._sh(seg_a), self._sh(seg_b), *args, **kw)

    def gw_run_multi(self, seg_a, seg_b, runs, max_groups=0):
        return self.pack.gw_run_multi(self._sh(seg_a), self._sh(seg_b), runs, max_groups)

    def gw_multi_begin(self, seg_a, seg_b, runs, max_groups=0):
        return self.pack.gw_multi_begin(self._sh(seg_a), self._sh(seg_b), runs, max_groups)

    def cluster(self, seg_a, seg_b, metric, lambdac=1.3):
        return self.pack.cluster(self._sh(seg_a), self._sh(seg_b), metric, lambdac)

    def cluster_inorder(self, speakers, seg_a, seg_b, metric, lambdac, threshold):
        sh = [[(a + self.off, b + self.off) for a, b in s] for s in speakers]
        return self.pack.cluster_inorder(sh, self._sh(seg_a), self._sh(seg_b), metric, lambdac, threshold)


class Clusters(object):
    """``spkdiar_clus``: per-cluster statistics + pair matrix resident in HBM."""

    def __init__(self, feat, handle, n):
        self.feat = feat
        self.ctx = feat.ctx
        self.h = handle
        self.n = int(n)

    def close(self):
        if getattr(self, 'h', None) and self.ctx.h:
            self.ctx.lib.spkdiar_cluster_free(self.h)
        self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def run(self, threshold, max_spk=0, variant=1):
        """-> (merges [structured array of (a, b, d), compacted indices], stats[4])"""
        out = np.zeros(max(self.n, 1), dtype=MERGE_DTYPE)
        nm = C.c_int64(0)
        stats = np.zeros(4)
        self.ctx._check(self.ctx.lib.spkdiar_cluster_run(
            self.h, float(threshold), int(max_spk), int(variant), out.ctypes.data_as(C.c_void_p),
            out.shape[0], C.byref(nm), _p(stats, C.c_double)))
        return out[:nm.value], stats

    def run_sharded(self, threshold, max_spk, rank, nranks, exchange):
        """``exchange(mine: bytes) -> bytes`` is an all-gather of 16-byte candidates."""
        out = np.zeros(max(self.n, 1), dtype=MERGE_DTYPE)
        nm = C.c_int64(0)
        stats = np.zeros(4)

        def _cb(user, mine, allp):
            try:
                got = exchange(C.string_at(mine, 16))
                C.memmove(allp, got, 16 * nranks)
                return 0
            except Exception:           # pragma: no cover - surfaced as an error code
                return -1
        cb = EXCHANGE_FN(_cb)
        self.ctx._check(self.ctx.lib.spkdiar_cluster_run_sharded(
            self.h, float(threshold), int(max_spk), int(rank), int(nranks), cb, None,
            out.ctypes.data_as(C.c_void_p), out.shape[0], C.byref(nm), _p(stats, C.c_double)))
        return out[:nm.value], stats

    def run_sharded_nccl(self, threshold, max_spk, rank, nranks, unique_id):
        """The sharded run with NCCL called by the library on its own stream; ``unique_id``:
        the 128 bytes of ``nccl_unique_id()`` made on rank 0 (None when nranks == 1)."""
        out = np.zeros(max(self.n, 1), dtype=MERGE_DTYPE)
        nm = C.c_int64(0)
        stats = np.zeros(4)
        idbuf = C.create_string_buffer(bytes(unique_id), 128) if unique_id is not None else None
        self.ctx._check(self.ctx.lib.spkdiar_cluster_run_sharded_nccl(
            self.h, float(threshold), int(max_spk), int(rank), int(nranks), idbuf,
            out.ctypes.data_as(C.c_void_p), out.shape[0], C.byref(nm), _p(stats, C.c_double)))
        return out[:nm.value], stats

    def run_sharded_p2p(self, threshold, max_spk, rank, nranks, mailbox_ptrs, seq_base):
        """The sharded run as one persistent kernel per rank with the exchange through peer
        memory; ``mailbox_ptrs``: device pointers (ints) of all ranks' mailboxes as mapped here."""
        out = np.zeros(max(self.n, 1), dtype=MERGE_DTYPE)
        nm = C.c_int64(0)
        stats = np.zeros(4)
        arr = (C.c_void_p * max(nranks, 1))(*[C.c_void_p(p) for p in (mailbox_ptrs or [0])])
        self.ctx._check(self.ctx.lib.spkdiar_cluster_run_sharded_p2p(
            self.h, float(threshold), int(max_spk), int(rank), int(nranks), arr, int(seq_base),
            out.ctypes.data_as(C.c_void_p), out.shape[0], C.byref(nm), _p(stats, C.c_double)))
        return out[:nm.value], stats

    def counters(self):
        """Phase cycle counters of the last persistent run: dict of cycles PER MERGE (CTA 0's clock)."""
        out = np.zeros(6, dtype=np.uint64)
        self.ctx._check(self.ctx.lib.spkdiar_cluster_counters(self.h, _p(out, C.c_uint64)))
        m = float(out[5]) if out[5] else 1.0
        names = ('scan', 'barrier1', 'pick_exchange', 'rescore', 'barrier2')
        d = {k: float(out[i]) / m for i, k in enumerate(names)}
        d['merges'] = int(out[5])
        return d

    def rowlog(self, cap_rows):
        """Test hook: the next run() also records row ``a`` of the pair matrix after every merge
        (original indices); returns the (cap_rows, n) array the run will fill."""
        self._rowlog = np.zeros((cap_rows, self.n)) if cap_rows > 0 else None
        self.ctx._check(self.ctx.lib.spkdiar_cluster_rowlog(
            self.h, _p(self._rowlog, C.c_double) if cap_rows > 0 else None, cap_rows))
        return self._rowlog

    def matrix(self):
        m = np.empty((self.n, self.n))
        alive = np.empty(self.n, dtype=np.uint8)
        self.ctx._check(self.ctx.lib.spkdiar_cluster_matrix(self.h, _p(m, C.c_double), _p(alive, C.c_uint8)))
        return m, alive.astype(bool)
