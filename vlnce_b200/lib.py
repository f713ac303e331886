"""ctypes binding of ``libvlnce_b200.so`` — the ONLY compute backend of this package.

There is no CPU / eager-PyTorch fallback: if the CUDA library is missing or a call
returns a non-zero status a ``VlnceError`` is raised.  Signatures are parsed from
``include/vlnce_b200.h`` so the header stays the single source of truth for the C ABI.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "vlnce_b200.h")
LIB_PATH = os.path.join(HERE, "libvlnce_b200.so")

F32, BF16, U8, I64, F16 = 0, 1, 2, 3, 4
_DTYPE_CODE = {torch.float32: F32, torch.bfloat16: BF16, torch.uint8: U8, torch.int64: I64, torch.float16: F16}


class VlnceError(RuntimeError):
    pass


_CTYPE = {
    "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double,
    "long long": ctypes.c_longlong, "size_t": ctypes.c_size_t, "cudaStream_t": ctypes.c_void_p,
}


def parse_header(path: str = HEADER) -> Dict[str, Tuple[str, List[str]]]:
    """{function name: (return type, [argument C types])} for every prototype."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = "\n".join(ln for ln in src.splitlines() if not ln.lstrip().startswith("#"))
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int|long\s+long|void)\s+(vlnce_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    types.append("ptr")
                else:
                    types.append(" ".join(a.split(" ")[:-1]).replace("const ", "").replace("unsigned ", ""))
        protos[name] = (" ".join(ret.split()), types)
    return protos


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise VlnceError(
                f"{LIB_PATH} is missing: build it with `python -m vlnce_b200.build` "
                "(nvcc, sm_100a).  vlnce_b200 has no CPU or eager fallback.")
        dll = ctypes.CDLL(LIB_PATH)
        for name, (ret, types) in self.protos.items():
            fn = getattr(dll, name)  # AttributeError = header/library mismatch: fail loudly
            fn.argtypes = [ctypes.c_void_p if t == "ptr" else _CTYPE[t] for t in types]
            fn.restype = {"int": ctypes.c_int, "long long": ctypes.c_longlong,
                          "const char *": ctypes.c_char_p, "const char*": ctypes.c_char_p,
                          "void": None}[ret]
        self._dll = dll
        return dll

    def __getattr__(self, name):
        if not name.startswith("vlnce_"):
            raise AttributeError(name)
        fn = getattr(self.load(), name)
        ret = self.protos[name][0]
        if ret != "int":
            return fn

        def call(*args):
            status = fn(*args)
            if status != 0:
                msg = self._dll.vlnce_last_error().decode()
                raise VlnceError(f"{name} failed with status {status}: {msg}")

        call.__name__ = name
        setattr(self, name, call)
        return call


C = _Lib()


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def code(t_or_dtype) -> int:
    dt = t_or_dtype.dtype if torch.is_tensor(t_or_dtype) else t_or_dtype
    return _DTYPE_CODE[dt]


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class guard:
    """Make ``device`` the current CUDA device for the duration of a public entry point.  Kernels are enqueued on
    ``torch.cuda.current_stream()`` and scratch is allocated on the current device, while the reference's trainers
    never call ``torch.cuda.set_device`` (they build ``torch.device("cuda", TORCH_GPU_ID / local_rank)``,
    ``base_il_trainer.py:46``, ``ddppo_waypoint_trainer.py:328-341``): without this a policy on cuda:1 would launch on
    device 0's stream with device-1 pointers.  No-op for CPU tensors (dry-run walks) and when the device is current."""

    def __init__(self, device):
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        self.ctx = None
        if dev.type == "cuda" and torch.cuda.is_available():
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            if idx != torch.cuda.current_device():
                self.ctx = torch.cuda.device(idx)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def guarded(fn):
    """Decorator for methods of an ``nn.Module`` (or of an object with a ``.policy`` / ``.actor_critic`` module):
    run with the module's device current (see ``guard``)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        mod = self
        for attr in ("policy", "actor_critic"):
            if not isinstance(mod, torch.nn.Module) or attr in getattr(mod, "__dict__", {}):
                mod = getattr(self, attr, mod)
                break
        try:
            dev = next(mod.parameters()).device
        except (StopIteration, AttributeError):
            return fn(self, *args, **kwargs)
        with guard(dev):
            return fn(self, *args, **kwargs)

    return wrapper


def require_cuda(t: torch.Tensor, what: str = "tensor"):
    if not t.is_cuda:
        raise VlnceError(f"{what} is on {t.device}: vlnce_b200 computes on CUDA only (no CPU fallback)")


class dry_run:
    """Test helper: inside this context every C-ABI call is validated against the header
    prototype (argument count and kinds) and recorded, but NOT executed, and the CUDA-device
    check is waived — it lets the CPU test-suite walk the whole launch sequence (shapes,
    pointer arithmetic, attribute names) without a GPU.  Never used by the product path.

    ``executor``: optional test-supplied callable ``(name, args)`` run after the validation of each
    call (``tests/abi_emulator.py`` passes a numpy model of the head kernels so that the launch
    plans can be checked numerically on host memory).  The package itself ships no executor."""

    def __init__(self, executor=None):
        self.calls = []
        self.executor = executor

    def __enter__(self):
        global C, require_cuda, stream
        self._saved = (C, require_cuda, stream)
        outer = self

        class _Checker:
            protos = self._saved[0].protos

            def __getattr__(self, name):
                if name not in self.protos:
                    raise AttributeError(name)
                ret, types = self.protos[name]

                def call(*args):
                    if len(args) != len(types):
                        raise TypeError(f"{name}: {len(args)} arguments for {len(types)} parameters")
                    for i, (a, t) in enumerate(zip(args, types)):
                        if t in ("ptr", "cudaStream_t"):
                            ok = a is None or isinstance(a, int)
                        elif t in ("int", "long long", "size_t"):
                            ok = isinstance(a, int) and not isinstance(a, bool)
                        else:
                            ok = isinstance(a, (int, float))
                        if not ok:
                            raise TypeError(f"{name}: argument {i} ({t}) got {type(a).__name__} {a!r}")
                    outer.calls.append(name)
                    if outer.executor is not None:
                        outer.executor(name, args)
                    return 0

                return call

        C = _Checker()
        require_cuda = lambda t, what="tensor": None  # noqa: E731
        stream = lambda: 0  # noqa: E731
        import sys
        for mod in list(sys.modules.values()):
            if getattr(mod, "__name__", "").startswith("vlnce_b200") and mod is not sys.modules[__name__]:
                if getattr(mod, "C", None) is self._saved[0]:
                    mod.C = C
        return self

    def __exit__(self, *exc):
        global C, require_cuda, stream
        checker = C
        C, require_cuda, stream = self._saved
        import sys
        for mod in list(sys.modules.values()):
            if getattr(mod, "__name__", "").startswith("vlnce_b200") and getattr(mod, "C", None) is checker:
                mod.C = C
        return False


class profile:
    """Device-time profile of the library calls issued inside the context: a CUDA-event pair
    around every C-ABI call on its launching stream (run with VLNCE_OVERLAP=0 semantics for
    clean serial numbers).  ``summary()`` -> {function: {calls, us}} after a device sync."""

    def __init__(self):
        self.events = []

    def __enter__(self):
        global C
        self._saved = C
        real, outer = C, self

        class _Timed:
            protos = real.protos

            def __getattr__(self, name):
                fn = getattr(real, name)
                if real.protos.get(name, ("",))[0] != "int":
                    return fn

                def call(*args):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    fn(*args)
                    b.record()
                    outer.events.append((name, a, b))

                return call

        C = _Timed()
        import sys
        for mod in list(sys.modules.values()):
            if getattr(mod, "__name__", "").startswith("vlnce_b200") and getattr(mod, "C", None) is real:
                mod.C = C
        return self

    def __exit__(self, *exc):
        global C
        timed = C
        C = self._saved
        import sys
        for mod in list(sys.modules.values()):
            if getattr(mod, "__name__", "").startswith("vlnce_b200") and getattr(mod, "C", None) is timed:
                mod.C = C
        return False

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, a, b in self.events:
            d = out.setdefault(name, {"calls": 0, "us": 0.0})
            d["calls"] += 1
            d["us"] += a.elapsed_time(b) * 1e3
        return dict(sorted(out.items(), key=lambda kv: -kv[1]["us"]))
