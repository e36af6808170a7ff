"""ctypes binding of the C ABI (include/memflow_b200.h) -> memflow_b200/lib/libmemflow_b200.so.

There is NO CPU fallback: if the library is missing, or a tensor is not on a CUDA device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
#: MEMFLOW_B200_LIB points the loader at another build of the same library (A/B timing of two builds on one box)
LIB_PATH = os.environ.get("MEMFLOW_B200_LIB") or os.path.join(_HERE, "lib", "libmemflow_b200.so")

#: every symbol include/memflow_b200.h declares (tests check the .so exports all of them)
SYMBOLS = (
    "mf_abi_version", "mf_error_string", "mf_device_info",
    "mf_rambo_forward_f64", "mf_rambo_forward_f32", "mf_massvar_solve_f64", "mf_rambo_forward_bwd_f64", "mf_rambo_inverse_bwd_f64", "mf_rambo_inverse_f64", "mf_rambo_inverse_f32", "mf_get_ps_f64",
    "mf_rqs_forward_f32", "mf_rqs_forward_f64", "mf_rqs_inverse_f32", "mf_rqs_inverse_f64",
    "mf_rqs_forward_strided_f32", "mf_rqs_forward_strided_f64", "mf_rqs_inverse_strided_f32", "mf_rqs_inverse_strided_f64",
    "mf_searchsorted_f32", "mf_searchsorted_f64",
    "mf_rqs_forward_bwd_f32", "mf_rqs_forward_bwd_f64", "mf_rqs_inverse_bwd_f32", "mf_rqs_inverse_bwd_f64",
    "mf_is_reduce_workspace_bytes", "mf_is_reduce_f64", "mf_is_chain_f64", "mf_flow_sample_f32", "mf_flow_sample_f64", "mf_rambo_is_f64", "mf_higgs_decay_f64", "mf_top_decay_f64",
    "mf_decay_partons_f64", "mf_decay_partons_bwd_f64", "mf_sampling_partons_f64", "mf_sampling_partons_bwd_f64",
)


class MemflowB200Error(RuntimeError):
    pass


class DecayPartonsCfg(ctypes.Structure):
    """mf_decay_partons_cfg of include/memflow_b200.h."""
    _fields_ = [("mean_prop", ctypes.c_double * 3), ("std_prop", ctypes.c_double * 3),
                ("mean_lab", ctypes.c_double * 3), ("std_lab", ctypes.c_double * 3),
                ("mean_boost", ctypes.c_double * 2), ("std_boost", ctypes.c_double * 2),
                ("higgs_mass", ctypes.c_double), ("thad_mass", ctypes.c_double), ("tlep_mass", ctypes.c_double),
                ("w_had_mass", ctypes.c_double), ("w_lep_mass", ctypes.c_double), ("b_mass", ctypes.c_double),
                ("eps", ctypes.c_double),
                ("ptetaphi", ctypes.c_int32), ("unscale_phi", ctypes.c_int32), ("final_scaling", ctypes.c_int32),
                ("reserved", ctypes.c_int32)]


_lib = None

#: C type -> ctypes type of the scalar arguments of the ABI; every pointer (device or host array, struct, stream) travels as
#: c_void_p, which accepts None, integers, c_void_p, ctypes arrays and byref() results
_CTYPES = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "uint32_t": ctypes.c_uint32,
           "uint64_t": ctypes.c_uint64, "double": ctypes.c_double, "float": ctypes.c_float, "mf_stream_t": ctypes.c_void_p}
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "memflow_b200.h")


def header_signatures(path=None):
    """{symbol: (restype, [argtypes])} parsed from the declarations of include/memflow_b200.h -- the ONE table of the
    boundary: the binding sets argtypes from it (a wrong argument count or width is a TypeError / ArgumentError instead of
    silent memory corruption) and tests/test_abi_cpu.py checks the exports against it."""
    import re
    src = open(path or HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"typedef\s+struct[^{;]*\{.*?\}[^;]*;", " ", src, flags=re.S)
    src = " ".join(l for l in src.splitlines() if not l.lstrip().startswith("#") and "\\" not in l[-2:])
    out = {}
    for stmt in src.split(";"):
        stmt = stmt.strip()
        m = re.search(r"(mf_[a-z0-9_]+)\s*\((.*)\)\s*$", stmt, flags=re.S)
        if not m:
            continue
        name, args = m.group(1), m.group(2)
        ret = stmt[: m.start(1)]
        restype = ctypes.c_char_p if "char" in ret else (ctypes.c_int64 if "int64_t" in ret else ctypes.c_int)
        argtypes = []
        for a in (x.strip() for x in args.split(",")):
            if a in ("", "void"):
                continue
            if "*" in a:
                argtypes.append(ctypes.c_void_p)
                continue
            toks = a.replace("const", " ").split()
            argtypes.append(_CTYPES[toks[0]])
        out[name] = (restype, argtypes)
    return out


def lib():
    """Load the shared library (once). Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MemflowB200Error(
                "memflow_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C memflow_b200/csrc`. There is no CPU / PyTorch fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.mf_error_string.restype = ctypes.c_char_p
        L.mf_error_string.argtypes = [ctypes.c_int]
        L.mf_is_reduce_workspace_bytes.restype = ctypes.c_int64
        for s in SYMBOLS:
            if s not in ("mf_error_string", "mf_is_reduce_workspace_bytes"):
                getattr(L, s).restype = ctypes.c_int
        if not os.path.exists(HEADER_PATH):
            raise MemflowB200Error("memflow_b200: %s not found: the binding takes the argument types of every entry point from "
                                   "the header (device pointers travel as plain integers)" % HEADER_PATH)
        for name, (restype, argtypes) in header_signatures().items():
            f = getattr(L, name)
            f.restype, f.argtypes = restype, argtypes
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        raise MemflowB200Error("%s failed: %s (code %d)" % (what, lib().mf_error_string(rc).decode(), rc))


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL), as the int / None the c_void_p argtypes of the binding accept."""
    if t is None:
        return None
    if not t.is_cuda:
        raise MemflowB200Error("memflow_b200 kernels need CUDA tensors (got device %s); there is no CPU fallback"
                               % t.device)
    return t.data_ptr()


def stream_ptr(device):
    """Raw cudaStream_t of torch's current stream on `device` (an int: the bound argtype is c_void_p)."""
    idx = device.index
    return torch._C._cuda_getCurrentRawStream(idx if idx is not None else torch.cuda.current_device())


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """`with device_guard(dev):` makes `dev` the current CUDA device for the launch -- a no-op object when it already is
    (torch.cuda.device costs several microseconds per call; training calls the kernels with B ~ 1e3)."""
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def needs_dispatcher(*tensors):
    """True when a call has to go through the registered torch op (memflow_b200/ops.py) rather than straight to the C ABI:
    autograd is recording for one of the tensors, torch.compile is tracing, or a tensor is not a plain torch.Tensor
    (FakeTensor, functorch wrappers...). Plain eager inference calls skip the dispatcher's ~15 us."""
    if torch.compiler.is_compiling():
        return True
    grad = torch.is_grad_enabled()
    for t in tensors:
        if t is None:
            continue
        if type(t) is not torch.Tensor and type(t) is not torch.nn.Parameter:
            return True
        if grad and t.requires_grad:
            return True
    return False


def plain_eager(*tensors):
    """True when every tensor is a plain torch.Tensor / Parameter and nothing is tracing (torch.compile, functorch
    transforms, torch.func): such a call may use a directly-applied torch.autograd.Function instead of the dispatcher."""
    if torch.compiler.is_compiling():
        return False
    try:
        if torch._C._functorch.peek_interpreter_stack() is not None:   # inside vmap / grad / functionalize
            return False
    except Exception:
        return False
    for t in tensors:
        if t is not None and type(t) is not torch.Tensor and type(t) is not torch.nn.Parameter:
            return False
    return True


def host_array(values, ctype):
    vals = [float(v) for v in values] if ctype in (ctypes.c_double, ctypes.c_float) else [int(v) for v in values]
    return (ctype * len(vals))(*vals)


def require_cuda(t, name):
    if not torch.is_tensor(t) or not t.is_cuda:
        raise MemflowB200Error(
            "memflow_b200: `%s` must be a CUDA tensor (device is %s). This package has no CPU fallback."
            % (name, getattr(t, "device", type(t))))


_workspaces = {}


def reduce_workspace(device):
    """Zero-initialised per-(device, stream) workspace for the estimator reductions."""
    key = (torch.device(device).index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        n = int(lib().mf_is_reduce_workspace_bytes())
        ws = torch.zeros(n, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws
