"""ctypes binding of csrc/libdv3b200.so.  Signatures are parsed from include/dv3b200.h, the single
source of truth of the C ABI.  There is NO fallback: a missing library or a failing call raises."""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DV3_LIB") or os.path.join(HERE, "csrc", "libdv3b200.so")   # DV3_LIB: A/B runs of two builds
HEADER = os.path.join(os.path.dirname(HERE), "include", "dv3b200.h")

_CTYPES = {"int": ctypes.c_int, "unsigned": ctypes.c_uint, "float": ctypes.c_float,
           "long long": ctypes.c_longlong, "double": ctypes.c_double}


class Dv3Error(RuntimeError):
    pass


def parse_header(path=HEADER):
    """-> {name: (restype, [(ctype, argname), ...])} for every function the header declares."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"#[^\n]*", " ", src)
    src = src.replace('extern "C" {', " ").replace("}", " ")
    decls = {}
    for m in re.finditer(r"([\w\s\*]+?)\b(dv3_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                ty = mm.group(1).strip()
                params.append((ctypes.c_void_p if "*" in ty else _CTYPES[ty.replace("const ", "")],
                               mm.group(2)))
        restype = ctypes.c_char_p if "char" in ret else _CTYPES[ret.replace("const ", "")]
        decls[name] = (restype, params)
    return decls


class _Lib:
    def __init__(self):
        self._dll = None
        self.decls = None

    def load(self):
        if self._dll is not None:
            return self
        if not os.path.exists(LIB_PATH):
            raise Dv3Error(
                "dv3b200 CUDA library not built: %s is missing. Run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (or python -m deepvoice3_pytorch_b200._build). There is no CPU fallback." % LIB_PATH)
        self._dll = ctypes.CDLL(LIB_PATH)
        self.decls = parse_header()
        for name, (restype, params) in self.decls.items():
            fn = getattr(self._dll, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = [t for t, _ in params]
        return self

    def call(self, name, *args):
        """Call an int-returning entry point; raise Dv3Error with dv3_last_error() on failure."""
        self.load()
        rc = getattr(self._dll, name)(*args)
        if rc != 0:
            raise Dv3Error("%s failed (%d): %s" % (name, rc, self._dll.dv3_last_error().decode()))

    def raw(self, name):
        self.load()
        return getattr(self._dll, name)


lib = _Lib()
