"""ctypes binding of liblasr_b200.so (the C ABI declared in include/lasr.h).

There is deliberately NO fallback: if the shared library is missing, or the device is not an
sm_100 part, every op raises.  Signatures are spelled with one character per argument:
  p = device pointer (torch.Tensor / int / None), i = int, f = float, z = size_t, q = uint64_t, s = cudaStream_t
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblasr_b200.so")

LASR_F32 = 0
LASR_BF16 = 1
ACT_NONE = 0
ACT_RELU = 1
ABI_VERSION = 4

def _parse_header(path):
    """Derive the ctypes signatures from include/lasr.h so the binding cannot drift from the C ABI."""
    import re

    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    sigs = {}
    for m in re.finditer(r"\n(const char\*|int|size_t)\s+(lasr_\w+)\s*\(([^)]*)\)\s*;", text):
        res, name, args = m.group(1), m.group(2), m.group(3).strip()
        codes = ""
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "lasr_stream_t" in a:
                    codes += "s"
                elif "*" in a:
                    codes += "p"
                elif a.startswith("size_t"):
                    codes += "z"
                elif a.startswith("uint64_t"):
                    codes += "q"
                elif a.startswith("float"):
                    codes += "f"
                elif a.startswith("int") or a.startswith("int32_t"):
                    codes += "i"
                else:
                    raise ValueError(f"unparsed argument {a!r} in {name}")
        sigs[name] = ({"const char*": "str", "int": "int", "size_t": "size_t"}[res], codes)
    return sigs


HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "lasr.h")
# name -> (restype, argument codes)
SIGNATURES = _parse_header(HEADER_PATH)

_CODES = {
    "p": ctypes.c_void_p,
    "i": ctypes.c_int,
    "f": ctypes.c_float,
    "z": ctypes.c_size_t,
    "q": ctypes.c_uint64,
    "s": ctypes.c_void_p,
}

_lib = None

# Optional instrumentation (bench.py / tools only): when PROFILE is a list, every call() appends
# (name, int/float args, pointer-non-null flags, start_event, end_event) recorded on the launching stream; CALLS counts entry-point calls.
PROFILE = None
PROFILE_EXTERNAL = False  # True: the events become event-record NODES of a CUDA graph being captured (timed per replay)
PROFILE_ONLY = None  # optional callable(name) -> bool: only these entry points get events (the rest are listed untimed)
CALLS = {"n": 0}


class LasrError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LasrError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C lightning-asr_b200/csrc`. There is no CPU / eager fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, codes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = [_CODES[c] for c in codes]
        fn.restype = {"int": ctypes.c_int, "str": ctypes.c_char_p, "size_t": ctypes.c_size_t}[res]
    _lib = lib
    return lib


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.data_ptr()
    return int(x)


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke an int-returning entry point on the current torch CUDA stream; raise on a non-zero code.

    The trailing stream argument is appended automatically.
    """
    lib = load()
    codes = SIGNATURES[name][1]
    assert codes.endswith("s"), name
    if len(args) != len(codes) - 1:
        raise TypeError(f"{name}: expected {len(codes) - 1} arguments, got {len(args)}")
    conv = []
    for c, a in zip(codes, args):
        if c == "p":
            conv.append(_ptr(a))
        elif c == "f":
            conv.append(float(a))
        elif c == "q":
            conv.append(int(a) & 0xFFFFFFFFFFFFFFFF)
        else:
            conv.append(int(a))
    conv.append(stream_ptr())
    CALLS["n"] += 1
    if PROFILE is not None and PROFILE_ONLY is not None and not PROFILE_ONLY(name):
        rc = getattr(lib, name)(*conv)
        PROFILE.append((name, tuple(a for c, a in zip(codes, conv) if c in "if"),
                        tuple(a is not None for c, a in zip(codes, conv) if c == "p"), None, None))
    elif PROFILE is not None:
        e0 = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        e1 = torch.cuda.Event(enable_timing=True, external=PROFILE_EXTERNAL)
        e0.record()
        rc = getattr(lib, name)(*conv)
        e1.record()
        PROFILE.append((name, tuple(a for c, a in zip(codes, conv) if c in "if"),
                        tuple(a is not None for c, a in zip(codes, conv) if c == "p"), e0, e1))
    else:
        rc = getattr(lib, name)(*conv)
    if rc != 0:
        msg = lib.lasr_strerror(rc).decode()
        raise LasrError(f"{name} failed: {msg} (code {rc})")
    return rc


def dtype_code(dt):
    if dt == torch.float32:
        return LASR_F32
    if dt == torch.bfloat16:
        return LASR_BF16
    raise LasrError(f"unsupported activation dtype {dt}")


def require_device():
    lib = load()
    if not torch.cuda.is_available():
        raise LasrError("lightning_asr_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    rc = lib.lasr_check_device()
    if rc != 0:
        raise LasrError("lightning_asr_b200 kernels are built for sm_100a only: " + lib.lasr_strerror(rc).decode())
