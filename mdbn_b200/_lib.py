"""ctypes binding of the C ABI in include/mdbn_b200.h (libmdbn_b200.so).

There is no CPU fallback: importing this module without the built library, or
creating a context without a B200, raises."""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# MDBN_B200_LIB: an alternative build of the same library (e.g. the instrumented one of scripts/build_debug.sh)
LIB_PATH = os.environ.get("MDBN_B200_LIB") or os.path.join(_HERE, "csrc", "libmdbn_b200.so")

RBM, GRBM = 0, 1
RNG_NONE, RNG_BUFFER, RNG_PHILOX = 0, 1, 2
PATH_AUTO, PATH_GENERIC, PATH_SKINNY, PATH_TENSOR, PATH_TINY, PATH_MID = 0, 1, 2, 3, 4, 5
PHASE_FULL, PHASE_STATS, PHASE_APPLY = 0, 1, 2
PATHS = {"auto": PATH_AUTO, "generic": PATH_GENERIC, "skinny": PATH_SKINNY, "tensor": PATH_TENSOR, "tiny": PATH_TINY, "mid": PATH_MID}


class MdbnError(RuntimeError):
    pass


class Rng(C.Structure):
    _fields_ = [("mode", C.c_int), ("buffer", C.c_void_p), ("seed", C.c_ulonglong), ("offset", C.c_ulonglong)]


class CdArgs(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("noisy", C.c_int), ("B", C.c_int), ("B_nom", C.c_int),
        ("V", C.c_int), ("H", C.c_int), ("k", C.c_int),
        ("W", C.c_void_p), ("ldw", C.c_int), ("hbias", C.c_void_p), ("vbias", C.c_void_p),
        ("W_speed", C.c_void_p), ("hbias_speed", C.c_void_p), ("vbias_speed", C.c_void_p),
        ("W_snap", C.c_void_p), ("data", C.c_void_p), ("ld_data", C.c_longlong),
        ("indices", C.c_void_p), ("persistent", C.c_void_p), ("bit_i_idx", C.c_void_p),
        ("lr", C.c_float), ("momentum", C.c_float), ("lambda_1", C.c_float), ("lambda_2", C.c_float),
        ("weightcost", C.c_float), ("rng", Rng), ("cost_out", C.c_void_p),
        ("path", C.c_int), ("tf32", C.c_int), ("phase", C.c_int), ("stats_buf", C.c_void_p),
        ("B_total", C.c_int), ("comm", C.c_void_p),
    ]


EXPORTS = ("mdbn_abi_version", "mdbn_last_error", "mdbn_create", "mdbn_destroy", "mdbn_launch_count",
           "mdbn_set_tf32_phases",
           "mdbn_propup", "mdbn_forward", "mdbn_propdown", "mdbn_free_energy", "mdbn_cd_step", "mdbn_cd_steps", "mdbn_copy_async", "mdbn_stats_size",
           "mdbn_comm_unique_id", "mdbn_comm_init", "mdbn_comm_destroy", "mdbn_comm_all_reduce")
COMM_ID_BYTES = 128

_lib = None
_lock = threading.Lock()


def load():
    """Load libmdbn_b200.so (built by __graft_entry__.build()); raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise MdbnError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        lib.mdbn_abi_version.restype = i
        lib.mdbn_last_error.restype = C.c_char_p
        lib.mdbn_create.argtypes = [C.POINTER(vp), i]
        lib.mdbn_destroy.argtypes = [vp]
        lib.mdbn_launch_count.argtypes = [vp]
        lib.mdbn_launch_count.restype = C.c_ulonglong
        lib.mdbn_set_tf32_phases.argtypes = [vp, i]
        lib.mdbn_stats_size.argtypes = [i, i]
        lib.mdbn_stats_size.restype = C.c_longlong
        lib.mdbn_propup.argtypes = [vp, vp, i, vp, vp, i, i, i, i, vp, vp, vp, C.POINTER(Rng), vp]
        lib.mdbn_forward.argtypes = [vp, vp, i, vp, vp, i, i, i, i, vp, vp]
        lib.mdbn_propdown.argtypes = [vp, vp, i, vp, vp, i, i, i, i, i, i, vp, vp, vp, C.POINTER(Rng), vp]
        lib.mdbn_free_energy.argtypes = [vp, vp, i, vp, vp, vp, i, i, i, i, i, vp, vp]
        lib.mdbn_cd_step.argtypes = [vp, C.POINTER(CdArgs), vp]
        lib.mdbn_cd_steps.argtypes = [vp, C.POINTER(CdArgs), C.c_int, vp]
        lib.mdbn_copy_async.argtypes = [vp, vp, C.c_ulonglong, vp]
        lib.mdbn_comm_unique_id.argtypes = [C.c_char_p]
        lib.mdbn_comm_init.argtypes = [C.POINTER(vp), C.c_char_p, i, i, i]
        lib.mdbn_comm_destroy.argtypes = [vp]
        lib.mdbn_comm_all_reduce.argtypes = [vp, vp, C.c_ulonglong, vp]
        for n in EXPORTS:
            getattr(lib, n)
        if lib.mdbn_abi_version() != 2:
            raise MdbnError("ABI version mismatch")
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise MdbnError(load().mdbn_last_error().decode("utf-8", "replace"))


class Context:
    """Per-device scratch + launch accounting (mdbn_ctx)."""

    def __init__(self, device):
        self.lib = load()
        self.device = int(device)
        h = C.c_void_p()
        check(self.lib.mdbn_create(C.byref(h), self.device))
        self.handle = h

    def set_tf32_phases(self, enable):
        """Single-phase calls (propup/propdown/sample_*/free_energy) run on the tcgen05 path in fp32-exact
        split-TF32 arithmetic by default; enable=True selects plain TF32 (tolerance 2e-3)."""
        check(self.lib.mdbn_set_tf32_phases(self.handle, int(bool(enable))))

    @property
    def launches(self):
        return int(self.lib.mdbn_launch_count(self.handle))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mdbn_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Comm:
    """NCCL communicator owned by the library (mdbn_comm): the data-parallel CD step runs entirely behind the C ABI.
    `exchange_id(id_bytes_or_None) -> id_bytes` carries rank 0's unique id to the other ranks (the host's channel:
    torch.distributed, MPI, a file ...)."""

    def __init__(self, rank, world, device, exchange_id):
        self.lib = load()
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        buf = C.create_string_buffer(COMM_ID_BYTES)
        if self.rank == 0:
            check(self.lib.mdbn_comm_unique_id(buf))
        uid = exchange_id(bytes(buf.raw) if self.rank == 0 else None)
        h = C.c_void_p()
        check(self.lib.mdbn_comm_init(C.byref(h), uid, self.rank, self.world, self.device))
        self.handle = h

    def all_reduce(self, tensor, stream):
        check(self.lib.mdbn_comm_all_reduce(self.handle, tensor.data_ptr(), tensor.numel(), stream))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mdbn_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_ctxs = {}


def context(device_index):
    c = _ctxs.get(device_index)
    if c is None:
        c = _ctxs[device_index] = Context(device_index)
    return c
