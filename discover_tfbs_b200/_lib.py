"""ctypes binding of libtfbs_cuda.so (include/tfbs_cuda.h) plus thin handle classes.

This is the only place Python touches the C ABI.  There is no CPU fallback: if the shared
library is missing or no B200 is visible, every compute call raises RuntimeError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtfbs_cuda.so")

HIT_DTYPE = np.dtype([("pos", "<i8"), ("motif", "<i4"), ("score", "<f4")])  # tfbs_hit, 16 bytes

TFBS_MAX_MOTIF_LEN = 32
PER_NUCLEOTIDE, PER_STEP = 0, 1
_ERR_OVERFLOW = -5

_vp, _i64, _i32, _u64, _u32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/tfbs_cuda.h declaration by declaration
_PROTOTYPES = {
    "tfbs_last_error": (C.c_char_p, []),
    "tfbs_version": (C.c_char_p, []),
    "tfbs_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "tfbs_ctx_create": (C.c_int, [C.c_int, _pp]),
    "tfbs_ctx_destroy": (C.c_int, [_vp]),
    "tfbs_ctx_sync": (C.c_int, [_vp]),
    "tfbs_last_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "tfbs_last_kernel_launches": (C.c_int, [_vp, C.POINTER(_i64)]),
    "tfbs_total_kernel_launches": (C.c_int, [_vp, C.POINTER(_i64)]),
    "tfbs_last_scan_candidates": (C.c_int, [_vp, C.POINTER(_i64)]),
    "tfbs_last_hits_ms": (C.c_int, [_vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "tfbs_last_direct_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "tfbs_timer_start": (C.c_int, [_vp]),
    "tfbs_timer_stop": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "tfbs_host_alloc": (C.c_int, [_pp, _i64]),
    "tfbs_host_free": (C.c_int, [_vp]),
    "tfbs_device_pci_bus_id": (C.c_int, [C.c_int, C.c_char_p, _i32]),
    "tfbs_host_register": (C.c_int, [_vp, _i64]),
    "tfbs_host_unregister": (C.c_int, [_vp]),
    "tfbs_flush_l2": (C.c_int, [_vp]),
    "tfbs_measure_smem_gbs": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "tfbs_encode": (C.c_int, [_vp, _vp, _i64, _pp]),
    "tfbs_encode_device": (C.c_int, [_vp, _vp, _i64, _pp]),
    "tfbs_encode_into": (C.c_int, [_vp, _vp, _i64, _vp]),
    "tfbs_encode_device_into": (C.c_int, [_vp, _vp, _i64, _vp]),
    "tfbs_seq_create": (C.c_int, [_vp, _i64, _pp]),
    "tfbs_seq_destroy": (C.c_int, [_vp]),
    "tfbs_seq_length": (C.c_int, [_vp, C.POINTER(_i64)]),
    "tfbs_seq_download": (C.c_int, [_vp, _vp, _vp, _vp]),
    "tfbs_seq_base_counts": (C.c_int, [_vp, _vp, _vp]),
    "tfbs_gc_counts": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "tfbs_synth_ascii": (C.c_int, [_vp, _u64, _i64, _i64, C.c_double, _i32, _i32, _pp]),
    "tfbs_device_free": (C.c_int, [_vp, _vp]),
    "tfbs_device_download": (C.c_int, [_vp, _vp, _vp, _i64]),
    "tfbs_pwm_upload": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _pp]),
    "tfbs_pwms_destroy": (C.c_int, [_vp]),
    "tfbs_scan_dense": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "tfbs_scan_best": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "tfbs_scan_hits": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, C.POINTER(_i64)]),
    "tfbs_scan_hits_host": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _i32, C.POINTER(_i64)]),
    "tfbs_scan_hits_host_shard": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i64, _i64, C.POINTER(_i64)]),
    "tfbs_shapes_upload": (C.c_int, [_vp, _vp, _vp, _i32, _pp]),
    "tfbs_shapes_destroy": (C.c_int, [_vp]),
    "tfbs_shape_tracks": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "tfbs_shape_sites": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "tfbs_track_sites": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "tfbs_similarity_sums": (C.c_int, [_vp] + [_vp] * 6 + [_i64] + [_vp] * 6 + [_i64, _vp]),
    "tfbs_svc_decision": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _i32, C.c_double, C.c_double, _vp]),
    "tfbs_debug_swar4": (C.c_int, [_u32, C.POINTER(_u32), C.POINTER(_u32)]),
    "tfbs_debug_dense_tables": (C.c_int, [_vp, _i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), _vp, _i64]),
    "tfbs_debug_hits_group_plan": (C.c_int, [_i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "tfbs_debug_hits_block_plan": (C.c_int, [C.POINTER(_i32), _i32, _i32, C.POINTER(_i32), C.POINTER(_i32)]),
    "tfbs_pwms_hits_direct": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64)]),
    "tfbs_pwms_hits_blocks": (C.c_int, [_vp, C.POINTER(_i32), _i32, C.POINTER(_i32)]),
    "tfbs_debug_direct_lookup": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _i64, _vp, _vp]),
    "tfbs_debug_direct_motifs": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "tfbs_debug_hits_plan_checksum": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "tfbs_debug_hits_filter": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _i32, _vp, _i64, _vp, _vp, _vp, C.POINTER(_i32)]),
}

_lib = None


def load():
    """Load libtfbs_cuda.so and set every prototype.  Raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or make -C discover_tfbs_b200/csrc).  There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_PROTOTYPES)


def _check(rc, allow=()):
    if rc != 0 and rc not in allow:
        msg = load().tfbs_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libtfbs_cuda error {rc}: {msg}")
    return rc


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)


def hits_group_plan(L: int):
    """(A, B): 4-column and 3-column groups the hits kernel uses for a block of longest motif L."""
    a, b = _i32(0), _i32(0)
    _check(load().tfbs_debug_hits_group_plan(int(L), C.byref(a), C.byref(b)))
    return a.value, b.value


def dense_tables(pssm):
    """(K, G, e, neg_thr, entries int32[G][4**K][2]): the dense kernel's fixed-point k-mer tables of one
    motif (log-odds [L][4]); a window score is float32(sum of its G entries) * 2**-e, or -inf when
    the sum is below neg_thr.  Host-only test hook."""
    m = np.ascontiguousarray(pssm, dtype=np.float64)
    L = int(m.shape[0])
    k, g, e, neg = _i32(0), _i32(0), _i32(0), _i32(0)
    lib = load()
    _check(lib.tfbs_debug_dense_tables(_ptr(m), L, C.byref(k), C.byref(g), C.byref(e), C.byref(neg), None, 0))
    ent = np.zeros((g.value, 4 ** k.value, 2), dtype=np.int32)
    _check(lib.tfbs_debug_dense_tables(_ptr(m), L, C.byref(k), C.byref(g), C.byref(e), C.byref(neg), _ptr(ent), ent.size))
    return k.value, g.value, e.value, neg.value, ent


def hits_block_plan(lens, narrow_max_groups: int = 9):
    """[(A, B, narrow, n_motifs), ...]: the blocks the hits kernel cuts a library into (longest
    motifs first) when every block of at most `narrow_max_groups` column groups is taken in narrow
    form (64 motifs, 8-bit filter fields; wide blocks hold 32 with 16-bit fields).  A scan decides
    narrow/wide per block from the thresholds: see MotifLibrary.hits_blocks()."""
    a = np.ascontiguousarray(lens, dtype=np.int32)
    out = np.zeros((max(a.size, 1), 4), dtype=np.int32)
    nb = _i32(0)
    _check(load().tfbs_debug_hits_block_plan(a.ctypes.data_as(C.POINTER(_i32)), int(a.size), int(narrow_max_groups),
                                             out.ctypes.data_as(C.POINTER(_i32)), C.byref(nb)))
    return [tuple(int(v) for v in row) for row in out[: nb.value]]


def hits_filter_host(logodds, lens, thr, form: int, codes, with_plan: bool = False):
    """Host run of the hits kernel's candidate filter (test hook, no GPU): uint8[n][M][2], 1 where
    window start p of motif m passes on strand s.  form 0 wide, 1 narrow, 2 narrow + sticky blocks,
    3 the plan a scan would choose.  with_plan: also return [(A, B, form, n_motifs), ...] and the
    candidates per window start the table builder expects of every narrow / sticky block."""
    lo = np.ascontiguousarray(logodds, dtype=np.float64)
    ln = np.ascontiguousarray(lens, dtype=np.int32)
    th = np.ascontiguousarray(thr, dtype=np.float32)
    cd = np.ascontiguousarray(codes, dtype=np.uint8)
    M, Lmax = int(lo.shape[0]), int(lo.shape[1])
    out = np.zeros((cd.size, M, 2), dtype=np.uint8)
    plan = np.zeros((M, 4), dtype=np.int32)
    expected = np.zeros(M, dtype=np.float64)
    nb = _i32(0)
    _check(load().tfbs_debug_hits_filter(_ptr(lo), _ptr(ln), M, Lmax, _ptr(th), int(form), _ptr(cd), int(cd.size),
                                         _ptr(out), _ptr(plan), _ptr(expected), C.byref(nb)))
    if not with_plan:
        return out
    return out, [tuple(int(v) for v in row) for row in plan[: nb.value]], expected[: nb.value].copy()


def direct_lookup_host(logodds, lens, thr, codes):
    """Host run of the hits scan's DIRECT path (test hook, no GPU): (cand uint8[n][M][2], info dict)."""
    lo = np.ascontiguousarray(logodds, dtype=np.float64)
    ln = np.ascontiguousarray(lens, dtype=np.int32)
    th = np.ascontiguousarray(thr, dtype=np.float32)
    cd = np.ascontiguousarray(codes, dtype=np.uint8)
    M, Lmax = int(lo.shape[0]), int(lo.shape[1])
    out = np.zeros((cd.size, M, 2), dtype=np.uint8)
    info = np.zeros(4, dtype=np.int64)
    _check(load().tfbs_debug_direct_lookup(_ptr(lo), _ptr(ln), M, Lmax, _ptr(th), _ptr(cd), int(cd.size), _ptr(out), _ptr(info)))
    return out, {"motifs": int(info[0]), "max_len": int(info[1]), "keys": int(info[2]), "bitmap_bits": int(info[3])}


def hits_plan_checksum_host(logodds, lens, thr):
    """The complete plan a hits scan makes for a threshold vector, built on the host by the scan's own code
    (test hook, no GPU): ({blocks, words, tables, descriptors, direct_motifs, direct_keys}, [(A, B, form, motifs)])."""
    lo = np.ascontiguousarray(logodds, dtype=np.float64)
    ln = np.ascontiguousarray(lens, dtype=np.int32)
    th = np.ascontiguousarray(thr, dtype=np.float32)
    M, Lmax = int(lo.shape[0]), int(lo.shape[1])
    out = np.zeros(6, dtype=np.uint64)
    plan = np.zeros((M, 4), dtype=np.int32)
    _check(load().tfbs_debug_hits_plan_checksum(_ptr(lo), _ptr(ln), M, Lmax, _ptr(th), _ptr(out), _ptr(plan)))
    keys = ("blocks", "words", "tables", "descriptors", "direct_motifs", "direct_keys")
    info = {k: int(v) for k, v in zip(keys, out)}
    return info, [tuple(int(v) for v in row) for row in plan[: info["blocks"]]]


def direct_motifs_host(logodds, lens, thr) -> np.ndarray:
    """Which motifs the hits scan's direct path takes for a threshold vector (test hook, no GPU): bool[M]."""
    lo = np.ascontiguousarray(logodds, dtype=np.float64)
    ln = np.ascontiguousarray(lens, dtype=np.int32)
    th = np.ascontiguousarray(thr, dtype=np.float32)
    M, Lmax = int(lo.shape[0]), int(lo.shape[1])
    out = np.zeros(M, dtype=np.uint8)
    _check(load().tfbs_debug_direct_motifs(_ptr(lo), _ptr(ln), M, Lmax, _ptr(th), _ptr(out)))
    return out.astype(bool)


def device_count() -> int:
    n = C.c_int(0)
    rc = load().tfbs_device_count(C.byref(n))
    return n.value if rc == 0 else 0


def device_local_cpus(device: int):
    """CPUs on the NUMA node the GPU hangs off (from /sys/bus/pci/devices/<bus id>/local_cpulist),
    intersected with the CPUs this process may use; None when unknown."""
    buf = C.create_string_buffer(32)
    if load().tfbs_device_pci_bus_id(int(device), buf, 32) != 0:
        return None
    try:
        with open(f"/sys/bus/pci/devices/{buf.value.decode().lower()}/local_cpulist") as fh:
            text = fh.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except (OSError, ValueError, AttributeError):
        return None


def bind_thread_near(device: int):
    """Pin the CALLING thread to the CPUs local to `device` (pinned buffers it allocates and touches then
    live in that socket's memory).  Returns the previous affinity, or None when nothing was changed
    (unknown topology, or TFBS_NUMA_BIND=0)."""
    if os.environ.get("TFBS_NUMA_BIND", "1") == "0":
        return None
    cpus = device_local_cpus(device)
    if not cpus:
        return None
    before = os.sched_getaffinity(0)
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return None
    return before


def as_ascii(seq) -> np.ndarray:
    """str / bytes / uint8 array -> contiguous uint8 array (no copy when possible)."""
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    if isinstance(seq, (bytes, bytearray, memoryview)):
        return np.frombuffer(seq, dtype=np.uint8)
    return np.ascontiguousarray(seq, dtype=np.uint8)


class PinnedBuffer:
    """Pinned host memory exposed as a NumPy array (end-to-end path)."""

    def __init__(self, nbytes: int):
        self._p = C.c_void_p()
        _check(load().tfbs_host_alloc(C.byref(self._p), int(nbytes)))
        self.nbytes = int(nbytes)
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.uint8, count=self.nbytes)

    def view(self, dtype, count=None):
        a = self.array.view(dtype)
        return a if count is None else a[:count]

    def close(self):
        if self._p:
            self.array = None
            load().tfbs_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RegisteredHost:
    """Page-locks a caller-owned NumPy array in place for the lifetime of the object (or `with` block)."""

    def __init__(self, array: np.ndarray):
        if not array.flags["C_CONTIGUOUS"]:
            raise ValueError("only a C-contiguous array can be page-locked in place")
        self.array = array
        self._ptr = array.ctypes.data
        _check(load().tfbs_host_register(C.c_void_p(self._ptr), int(array.nbytes)))

    def close(self):
        if self._ptr:
            load().tfbs_host_unregister(C.c_void_p(self._ptr))
            self._ptr = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One B200 + one stream.  All compute goes through a Context."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self._lib = load()
        _check(self._lib.tfbs_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    # -- lifetime --------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.tfbs_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- bookkeeping -----------------------------------------------------------------------
    def sync(self):
        _check(self._lib.tfbs_ctx_sync(self._h))

    def last_kernel_ms(self) -> float:
        ms = C.c_float(0)
        _check(self._lib.tfbs_last_kernel_ms(self._h, C.byref(ms)))
        return float(ms.value)

    def last_kernel_launches(self) -> int:
        n = _i64(0)
        _check(self._lib.tfbs_last_kernel_launches(self._h, C.byref(n)))
        return int(n.value)

    def total_kernel_launches(self) -> int:
        n = _i64(0)
        _check(self._lib.tfbs_total_kernel_launches(self._h, C.byref(n)))
        return int(n.value)

    def last_scan_candidates(self) -> int:
        n = _i64(0)
        _check(self._lib.tfbs_last_scan_candidates(self._h, C.byref(n)))
        return int(n.value)

    def last_hits_ms(self):
        """(scan_ms, verify_ms) of the filter scan and the exact verify pass of the last scan_hits."""
        a, b = C.c_float(0), C.c_float(0)
        _check(self._lib.tfbs_last_hits_ms(self._h, C.byref(a), C.byref(b)))
        return float(a.value), float(b.value)

    def last_direct_ms(self) -> float:
        """Part of the last scan_hits' scan time spent in the direct-path kernel."""
        a = C.c_float(0)
        _check(self._lib.tfbs_last_direct_ms(self._h, C.byref(a)))
        return float(a.value)

    def timer_start(self):
        _check(self._lib.tfbs_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0)
        _check(self._lib.tfbs_timer_stop(self._h, C.byref(ms)))
        return float(ms.value)

    def measure_smem_gbs(self) -> float:
        """Measured shared-memory load bandwidth (GB/s) of this GPU: conflict-free LDS.32 probe kernel."""
        g = C.c_float(0)
        _check(self._lib.tfbs_measure_smem_gbs(self._h, C.byref(g)))
        return float(g.value)

    def flush_l2(self):
        _check(self._lib.tfbs_flush_l2(self._h))

    # -- (1) encode ------------------------------------------------------------------------
    def encode(self, seq) -> "EncodedSeq":
        a = as_ascii(seq)
        h = C.c_void_p()
        _check(self._lib.tfbs_encode(self._h, _ptr(a), a.size, C.byref(h)))
        return EncodedSeq(self, h, a.size)

    def alloc_seq(self, n: int) -> "EncodedSeq":
        """An empty handle for n bases (device buffers only; filled by reencode / scan_hits_host)."""
        h = C.c_void_p()
        _check(self._lib.tfbs_seq_create(self._h, int(n), C.byref(h)))
        return EncodedSeq(self, h, int(n))

    def encode_device(self, dev_ptr: int, n: int) -> "EncodedSeq":
        h = C.c_void_p()
        _check(self._lib.tfbs_encode_device(self._h, C.c_void_p(dev_ptr), int(n), C.byref(h)))
        return EncodedSeq(self, h, int(n))

    def synth_ascii(self, seed: int, start: int, n: int, gc: float = 0.41, n_runs: bool = True,
                    lowercase: bool = True) -> "DeviceBytes":
        p = C.c_void_p()
        _check(self._lib.tfbs_synth_ascii(self._h, int(seed), int(start), int(n), float(gc),
                                          int(n_runs), int(lowercase), C.byref(p)))
        return DeviceBytes(self, p, int(n))

    # -- (2) PWM ---------------------------------------------------------------------------
    def upload_pwms(self, logodds, lens=None) -> "MotifLibrary":
        """logodds: float64[M][Lmax][4] (ACGT), or a list of [L_m][4] matrices."""
        if isinstance(logodds, (list, tuple)):
            mats = [np.asarray(m, dtype=np.float64) for m in logodds]
            lens = np.array([m.shape[0] for m in mats], dtype=np.int32)
            lmax = int(lens.max())
            arr = np.zeros((len(mats), lmax, 4), dtype=np.float64)
            for i, m in enumerate(mats):
                arr[i, : m.shape[0]] = m
        else:
            arr = np.ascontiguousarray(logodds, dtype=np.float64)
            if arr.ndim == 2:
                arr = arr[None]
            lens = (np.full(arr.shape[0], arr.shape[1], dtype=np.int32) if lens is None
                    else np.ascontiguousarray(lens, dtype=np.int32))
        h = C.c_void_p()
        _check(self._lib.tfbs_pwm_upload(self._h, _ptr(arr), _ptr(lens), arr.shape[0], arr.shape[1],
                                         C.byref(h)))
        return MotifLibrary(self, h, arr, lens)

    def scan_dense(self, seq: "EncodedSeq", pwms: "MotifLibrary", to_host: bool = True):
        """-> (fwd, rev) float32[M][n] (None, None when to_host is False: results stay in HBM)."""
        if not to_host:
            _check(self._lib.tfbs_scan_dense(self._h, seq._h, pwms._h, None, None))
            return None, None
        fwd = np.empty((pwms.M, seq.n), dtype=np.float32)
        rev = np.empty((pwms.M, seq.n), dtype=np.float32)
        _check(self._lib.tfbs_scan_dense(self._h, seq._h, pwms._h, _ptr(fwd), _ptr(rev)))
        return fwd, rev

    def scan_best(self, seq: "EncodedSeq", pwms: "MotifLibrary", rec_start, rec_len, with_pos: bool = False):
        """Per-record best window: float32[S][M][2] (fwd, rev; NaN = no valid window), reduced on the
        device; with_pos also returns int32[S][M][2] offsets of the first best window (-1 = none)."""
        rs = np.ascontiguousarray(rec_start, dtype=np.int64)
        rl = np.ascontiguousarray(rec_len, dtype=np.int64)
        best = np.empty((rs.size, pwms.M, 2), dtype=np.float32)
        pos = np.empty((rs.size, pwms.M, 2), dtype=np.int32) if with_pos else None
        _check(self._lib.tfbs_scan_best(self._h, seq._h, pwms._h, _ptr(rs), _ptr(rl), rs.size, _ptr(best), _ptr(pos)))
        return (best, pos) if with_pos else best

    def scan_hits(self, seq: "EncodedSeq", pwms: "MotifLibrary", thresholds, cap: int = 1 << 20,
                  sort: bool = True, to_host: bool = True, out: np.ndarray = None,
                  grow: bool = True):
        """-> structured array of HIT_DTYPE (to_host) or the hit count (device only)."""
        thr = np.ascontiguousarray(np.broadcast_to(np.asarray(thresholds, dtype=np.float32), (pwms.M,)))
        n = _i64(0)
        while True:
            if to_host:
                buf = out if out is not None and out.size >= cap else np.empty(cap, dtype=HIT_DTYPE)
                rc = self._lib.tfbs_scan_hits(self._h, seq._h, pwms._h, _ptr(thr), _ptr(buf), cap,
                                              int(sort), C.byref(n))
            else:
                buf = None
                rc = self._lib.tfbs_scan_hits(self._h, seq._h, pwms._h, _ptr(thr), None, cap, 0,
                                              C.byref(n))
            if rc == _ERR_OVERFLOW and grow:
                cap = int(n.value) + 1024
                out = None
                continue
            _check(rc)
            break
        return buf[: n.value] if to_host else int(n.value)

    def scan_hits_host(self, ascii_host: np.ndarray, seq: "EncodedSeq", pwms: "MotifLibrary", thresholds,
                       out: np.ndarray, chunks: int = 6, sort: bool = False):
        """Chunk-pipelined end-to-end scan of a host buffer (H2D, encode+scan, D2H overlap).
        `out` is a preallocated HIT_DTYPE array (its size is the capacity); returns out[:n_hits]."""
        thr = np.ascontiguousarray(np.broadcast_to(np.asarray(thresholds, dtype=np.float32), (pwms.M,)))
        a = as_ascii(ascii_host)
        n = _i64(0)
        _check(self._lib.tfbs_scan_hits_host(self._h, _ptr(a), a.size, seq._h, pwms._h, _ptr(thr), _ptr(out),
                                             out.size, int(chunks), int(sort), C.byref(n)))
        return out[: n.value]

    def scan_hits_host_shard(self, ascii_host: np.ndarray, seq: "EncodedSeq", pwms: "MotifLibrary", thresholds,
                             out: np.ndarray, owned: int, shard_start: int, chunks: int = 6, sort: bool = False):
        """scan_hits_host for one shard of a longer sequence (buffer = shard + right halo): only windows
        starting in the first `owned` bases are reported, positions are shifted by `shard_start` — both
        on the device."""
        thr = np.ascontiguousarray(np.broadcast_to(np.asarray(thresholds, dtype=np.float32), (pwms.M,)))
        a = as_ascii(ascii_host)
        n = _i64(0)
        _check(self._lib.tfbs_scan_hits_host_shard(self._h, _ptr(a), a.size, seq._h, pwms._h, _ptr(thr), _ptr(out),
                                                   out.size, int(chunks), int(sort), int(owned), int(shard_start),
                                                   C.byref(n)))
        return out[: n.value]

    # -- (3) shape -------------------------------------------------------------------------
    def upload_shapes(self, tables, kinds) -> "ShapeTables":
        tab = np.ascontiguousarray(tables, dtype=np.float32)
        kinds = np.ascontiguousarray(kinds, dtype=np.int32)
        assert tab.shape == (kinds.size, 2, 1024), tab.shape
        h = C.c_void_p()
        _check(self._lib.tfbs_shapes_upload(self._h, _ptr(tab), _ptr(kinds), kinds.size, C.byref(h)))
        return ShapeTables(self, h, kinds)

    def shape_tracks(self, seq: "EncodedSeq", shapes: "ShapeTables", rec_start=None, rec_len=None,
                     to_host: bool = True):
        S = 0 if rec_start is None else len(rec_start)
        rs = None if S == 0 else np.ascontiguousarray(rec_start, dtype=np.int64)
        rl = None if S == 0 else np.ascontiguousarray(rec_len, dtype=np.int64)
        out = np.empty((shapes.T, seq.n), dtype=np.float32) if to_host else None
        _check(self._lib.tfbs_shape_tracks(self._h, seq._h, shapes._h, _ptr(rs), _ptr(rl), S, _ptr(out)))
        return out

    def shape_sites(self, seq: "EncodedSeq", shapes: "ShapeTables", rec_start, rec_len, start=None,
                    end=None, strand=None, width: int = 1, mode: int = 0, trailing_empty: bool = True):
        rs = np.ascontiguousarray(rec_start, dtype=np.int64)
        rl = np.ascontiguousarray(rec_len, dtype=np.int64)
        S = rs.size
        st = None if start is None else np.ascontiguousarray(start, dtype=np.int64)
        en = None if end is None else np.ascontiguousarray(end, dtype=np.int64)
        sd = None if strand is None else np.ascontiguousarray(strand, dtype=np.int8)
        out = np.empty((S, shapes.T, int(width)), dtype=np.float32)
        _check(self._lib.tfbs_shape_sites(self._h, seq._h, shapes._h, _ptr(rs), _ptr(rl), _ptr(st),
                                          _ptr(en), _ptr(sd), S, int(width), int(mode),
                                          int(trailing_empty), _ptr(out)))
        return out

    def track_sites(self, values, rec_off, rec_ntok, start, end, strand, width: int):
        """Python-slice gather over already-parsed token values (legacy dict input)."""
        val = np.ascontiguousarray(values, dtype=np.float32)
        ro = np.ascontiguousarray(rec_off, dtype=np.int64)
        rn = np.ascontiguousarray(rec_ntok, dtype=np.int64)
        st = np.ascontiguousarray(start, dtype=np.int64)
        en = np.ascontiguousarray(end, dtype=np.int64)
        sd = np.ascontiguousarray(strand, dtype=np.int8)
        out = np.empty((ro.size, int(width)), dtype=np.float32)
        _check(self._lib.tfbs_track_sites(self._h, _ptr(val), val.size, _ptr(ro), _ptr(rn), _ptr(st),
                                          _ptr(en), _ptr(sd), ro.size, int(width), _ptr(out)))
        return out

    def similarity_sums(self, cand: dict, true: dict) -> np.ndarray:
        """cand / true: dicts with words(u4) counts(u2) woff(i4) nk(i4) hash(u4) hoff(i4).
        -> float64[S][3] sums of (Jaccard, PAS, PPS) over the true set."""
        def arr(d, k, dt):
            return np.ascontiguousarray(d[k], dtype=dt)
        c = [arr(cand, "words", np.uint32), arr(cand, "counts", np.uint16), arr(cand, "woff", np.int32),
             arr(cand, "nk", np.int32), arr(cand, "hash", np.uint32), arr(cand, "hoff", np.int32)]
        t = [arr(true, "words", np.uint32), arr(true, "counts", np.uint16), arr(true, "woff", np.int32),
             arr(true, "nk", np.int32), arr(true, "hash", np.uint32), arr(true, "hoff", np.int32)]
        S, U = c[3].size, t[3].size
        out = np.zeros((S, 3), dtype=np.float64)
        _check(self._lib.tfbs_similarity_sums(self._h, *[_ptr(a) for a in c], S, *[_ptr(a) for a in t], U, _ptr(out)))
        return out


class _Handle:
    _destroy = None

    def __init__(self, ctx: Context, h):
        self.ctx = ctx
        self._h = h

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            getattr(self.ctx._lib, self._destroy)(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EncodedSeq(_Handle):
    _destroy = "tfbs_seq_destroy"

    def __init__(self, ctx, h, n):
        super().__init__(ctx, h)
        self.n = int(n)

    def reencode(self, seq):
        a = as_ascii(seq)
        _check(self.ctx._lib.tfbs_encode_into(self.ctx._h, _ptr(a), a.size, self._h))

    def reencode_device(self, dev_ptr: int):
        _check(self.ctx._lib.tfbs_encode_device_into(self.ctx._h, C.c_void_p(dev_ptr), self.n, self._h))

    def download(self):
        packed = np.empty((self.n + 15) // 16, dtype=np.uint32)
        mask = np.empty((self.n + 31) // 32, dtype=np.uint32)
        _check(self.ctx._lib.tfbs_seq_download(self.ctx._h, self._h, _ptr(packed), _ptr(mask)))
        return packed, mask

    def base_counts(self) -> np.ndarray:
        out = np.zeros(5, dtype=np.int64)
        _check(self.ctx._lib.tfbs_seq_base_counts(self.ctx._h, self._h, _ptr(out)))
        return out

    def gc_counts(self, start, length) -> np.ndarray:
        st = np.ascontiguousarray(start, dtype=np.int64)
        ln = np.ascontiguousarray(length, dtype=np.int64)
        out = np.zeros(st.size, dtype=np.int64)
        _check(self.ctx._lib.tfbs_gc_counts(self.ctx._h, self._h, _ptr(st), _ptr(ln), st.size, _ptr(out)))
        return out


class MotifLibrary(_Handle):
    _destroy = "tfbs_pwms_destroy"

    def __init__(self, ctx, h, logodds, lens):
        super().__init__(ctx, h)
        self.logodds = logodds
        self.lens = lens
        self.M = int(logodds.shape[0])

    def hits_blocks(self):
        """[(A, B, narrow, n_motifs), ...] of the last hits scan of this library (empty before the
        first): per block the column groups, whether it ran in narrow form (64 motifs, 8-bit filter
        fields) or wide (32 motifs, 16-bit), and its motif count."""
        out = np.zeros((max((self.M + 31) // 32, 1), 4), dtype=np.int32)
        nb = _i32(0)
        _check(self.ctx._lib.tfbs_pwms_hits_blocks(self._h, out.ctypes.data_as(C.POINTER(_i32)), out.shape[0],
                                                   C.byref(nb)))
        return [tuple(int(v) for v in row) for row in out[: nb.value]]


    def hits_direct(self):
        """(n_motifs, max_len, keys) of the direct path of the last hits scan: the shortest motifs, looked up
        in an enumerated 12-mer hash instead of the table filter; (0, 0, 0) when the path was not taken."""
        n, ln, k = _i32(0), _i32(0), _i64(0)
        _check(self.ctx._lib.tfbs_pwms_hits_direct(self._h, C.byref(n), C.byref(ln), C.byref(k)))
        return int(n.value), int(ln.value), int(k.value)


class ShapeTables(_Handle):
    _destroy = "tfbs_shapes_destroy"

    def __init__(self, ctx, h, kinds):
        super().__init__(ctx, h)
        self.kinds = kinds
        self.T = int(kinds.size)


class DeviceBytes:
    """A library-owned ASCII buffer in HBM (tfbs_synth_ascii)."""

    def __init__(self, ctx, p, n):
        self.ctx, self._p, self.n = ctx, p, n

    @property
    def ptr(self) -> int:
        return self._p.value

    def download(self) -> np.ndarray:
        out = np.empty(self.n, dtype=np.uint8)
        _check(self.ctx._lib.tfbs_device_download(self.ctx._h, self._p, _ptr(out), self.n))
        return out

    def close(self):
        if self._p and getattr(self.ctx, "_h", None):
            self.ctx._lib.tfbs_device_free(self.ctx._h, self._p)
        self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
