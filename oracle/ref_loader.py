"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference implementation.

Two sources, in this order:
  * ``/root/reference`` (the build container): the files are read from where they lie;
  * ``oracle/_ref/*.pycode`` (the GPU box, where ``/root/reference`` does not exist): the
    reference's own modules COMPILED to CPython bytecode by ``build_ref()`` below, which
    ``__graft_entry__.build()`` calls while the reference tree is present.  ``oracle/_ref/``
    is git-ignored build output (like a compiled C reference's ``.so``) and travels to the
    GPU box with the snapshot; no reference source text enters the repository.
It is used by ``tests/golden/make_golden.py`` to generate the committed golden vectors, by
the ``not gpu`` tests that pin ``oracle/gmm2d_oracle.py`` / ``oracle/image_oracle.py``
against the real thing, and by ``bench.py``'s CPU legs (``cpu_baseline.kind ==
"reference"``).  The product never imports it.

How the reference is loaded:
  * ``restoration_algorithms.py`` imports unchanged once ``matplotlib``,
    ``matplotlib.pyplot`` and ``deepinv.optim.data_fidelity`` (attribute ``L2``)
    are stubbed in ``sys.modules`` (restoration_algorithms.py:4,9).
  * ``utils_2D.py`` imports unchanged with stubs for ``matplotlib{,.pyplot,.cm,
    .patches}``, ``bm3d`` and ``ot`` (utils_2D.py:4,6,15,16,21).
  * ``sampling_2D.py`` runs its whole experiment at import time
    (sampling_2D.py:74-251), so only its two ``FunctionDef`` nodes ``PnP_ULA``
    (sampling_2D.py:21-45) and ``SnoPnP_ULA`` (sampling_2D.py:48-72) are
    compiled, in ``utils_2D``'s namespace.
No reference source is copied: the files are read from where they lie.
"""
from __future__ import annotations

import ast
import importlib.machinery
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PSGLA_REFERENCE_ROOT", "/root/reference")
REF_BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
EXT = ".pycode"  # CPython bytecode in the .pyc container format; the snapshot that ships the repo to the GPU box drops *.pyc
_COMPILED = ("utils_2D", "restoration_algorithms", "sampling_2D_functions")


def _source_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sampling_2D.py"))


def _compiled_available() -> bool:
    return all(os.path.isfile(os.path.join(REF_BUILD, m + EXT)) for m in _COMPILED)


def reference_available() -> bool:
    return _source_available() or _compiled_available()


def reference_kind() -> str:
    """"source" (read from /root/reference), "compiled" (oracle/_ref/*.pycode) or "" (absent)."""
    return "source" if _source_available() else ("compiled" if _compiled_available() else "")


def _sampling_2D_functions_ast():
    path = os.path.join(REFERENCE_ROOT, "sampling_2D.py")
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("PnP_ULA", "SnoPnP_ULA")]
    assert len(keep) == 2, "reference sampling_2D.py no longer defines PnP_ULA / SnoPnP_ULA"
    return path, ast.Module(body=keep, type_ignores=[])


def build_ref() -> list:
    """Compiles the reference's modules, from the sources where they lie under /root/reference, into oracle/_ref/*.pycode
    (outputs only; nothing is copied).  No-op without the reference tree.  Returns the files written."""
    if not _source_available():
        return []
    import importlib._bootstrap_external as be
    import py_compile
    os.makedirs(REF_BUILD, exist_ok=True)
    out = []
    for mod in ("utils_2D", "restoration_algorithms"):
        out.append(py_compile.compile(os.path.join(REFERENCE_ROOT, mod + ".py"), cfile=os.path.join(REF_BUILD, mod + EXT),
                                      doraise=True, quiet=2))
    path, tree = _sampling_2D_functions_ast()  # sampling_2D.py runs its experiment at import: only its two samplers
    st = os.stat(path)
    data = be._code_to_timestamp_pyc(compile(tree, path, "exec"), int(st.st_mtime), st.st_size)
    target = os.path.join(REF_BUILD, "sampling_2D_functions" + EXT)
    with open(target, "wb") as fh:
        fh.write(data)
    out.append(target)
    return out


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = sys.modules.get(name)
    if mod is None:
        mod = types.ModuleType(name)
        mod.__dict__["__psgla_stub__"] = True
        sys.modules[name] = mod
    for k, v in attrs.items():
        if not hasattr(mod, k):
            setattr(mod, k, v)
    return mod


def _install_stubs() -> None:
    class _Nothing:  # placeholder for symbols that are imported but never called here
        def __init__(self, *a, **k):
            pass

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.patches",
                 "bm3d", "ot", "deepinv", "deepinv.optim", "deepinv.optim.data_fidelity"):
        try:
            if name not in sys.modules:
                importlib.import_module(name)
        except Exception:
            _stub(name)
    if getattr(sys.modules["matplotlib.patches"], "__psgla_stub__", False):
        _stub("matplotlib.patches", Ellipse=_Nothing)
    if getattr(sys.modules["matplotlib"], "__psgla_stub__", False):
        _stub("matplotlib", cm=sys.modules["matplotlib.cm"], pyplot=sys.modules["matplotlib.pyplot"],
              patches=sys.modules["matplotlib.patches"])
    if getattr(sys.modules["bm3d"], "__psgla_stub__", False):
        _stub("bm3d", bm3d=_Nothing, BM3DProfile=_Nothing)
    if getattr(sys.modules["deepinv.optim.data_fidelity"], "__psgla_stub__", False):
        _stub("deepinv.optim.data_fidelity", L2=_Nothing)
        _stub("deepinv.optim", data_fidelity=sys.modules["deepinv.optim.data_fidelity"])
        _stub("deepinv", optim=sys.modules["deepinv.optim"])


def _load_file(modname: str, filename: str) -> types.ModuleType:
    path = os.path.join(REFERENCE_ROOT, filename)
    if os.path.isfile(path):
        spec = importlib.util.spec_from_file_location(modname, path)
    else:  # the GPU box: the module as compiled by build_ref()
        cfile = os.path.join(REF_BUILD, filename[:-3] + EXT)
        spec = importlib.util.spec_from_loader(modname, importlib.machinery.SourcelessFileLoader(modname, cfile))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_CACHE: dict[str, types.ModuleType] = {}


def load_utils_2D() -> types.ModuleType:
    """The reference ``utils_2D`` module, unmodified (utils_2D.py)."""
    if "utils_2D" not in _CACHE:
        _install_stubs()
        _CACHE["utils_2D"] = _load_file("_psgla_ref_utils_2D", "utils_2D.py")
    return _CACHE["utils_2D"]


def load_sampling_2D() -> types.SimpleNamespace:
    """``PnP_ULA`` and ``SnoPnP_ULA`` exactly as written in sampling_2D.py:21-72.

    ``tqdm`` is replaced by the identity so that loading is silent; nothing else in
    the function bodies is touched.
    """
    if "sampling_2D" not in _CACHE:
        u2d = load_utils_2D()
        ns = dict(u2d.__dict__)
        ns["tqdm"] = lambda it, *a, **k: it
        if _source_available():
            path, tree = _sampling_2D_functions_ast()
            exec(compile(tree, path, "exec"), ns)
        else:
            import marshal
            with open(os.path.join(REF_BUILD, "sampling_2D_functions" + EXT), "rb") as fh:
                exec(marshal.loads(fh.read()[16:]), ns)
        _CACHE["sampling_2D"] = types.SimpleNamespace(PnP_ULA=ns["PnP_ULA"], SnoPnP_ULA=ns["SnoPnP_ULA"], namespace=ns)
    return _CACHE["sampling_2D"]


def load_restoration_algorithms() -> types.ModuleType:
    """The reference ``restoration_algorithms`` module (psgla :163-285, pnpula :38-160)."""
    if "restoration_algorithms" not in _CACHE:
        _install_stubs()
        mod = _load_file("_psgla_ref_restoration_algorithms", "restoration_algorithms.py")
        mod.tqdm = lambda it, *a, **k: it  # silence progress bars only
        _CACHE["restoration_algorithms"] = mod
    return _CACHE["restoration_algorithms"]
