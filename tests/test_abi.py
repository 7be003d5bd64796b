"""The C-ABI shared library loads and exports every symbol include/dl4ds_b200.h declares, with the
argument list the ctypes binding assumes (no compute calls -- runs without a GPU)."""
import ctypes
import os
import re

from dl4ds_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_prototypes():
    src = open(os.path.join(ROOT, 'include', 'dl4ds_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(const char\*|int64_t|int)\s+(dl4ds_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        codes = ''
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                if '*' in a:
                    codes += 'p'
                elif a.startswith('int64_t'):
                    codes += 'l'
                elif a.startswith('float'):
                    codes += 'f'
                elif a.startswith('int'):
                    codes += 'i'
                else:
                    raise AssertionError('unparsed argument %r of %s' % (a, name))
        protos[name] = ({'const char*': 's', 'int64_t': 'l', 'int': 'i'}[ret], codes)
    return protos


def test_library_exports_every_declared_symbol():
    protos = _header_prototypes()
    assert len(protos) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(lib, name), 'libdl4ds_b200.so does not export %s' % name


def test_binding_signatures_match_header():
    protos = _header_prototypes()
    assert set(protos) == set(_lib.SIGNATURES), set(protos) ^ set(_lib.SIGNATURES)
    for name, sig in protos.items():
        assert _lib.SIGNATURES[name] == sig, (name, _lib.SIGNATURES[name], sig)


def test_load_and_version():
    lib = _lib.load()
    assert lib.dl4ds_version() == 100
    assert isinstance(_lib.last_error(), str)


def test_bad_arguments_are_reported_not_crashed():
    lib = _lib.load()
    rc = lib.dl4ds_avgpool_coarsen(None, None, 1, 4, 4, 1, 2, None)
    assert rc == -1 and 'null' in _lib.last_error()
    rc = lib.dl4ds_adam_step(None, None, None, None, 10, 1e-3, 0.9, 0.999, 1e-7, 1, 1.0, None)
    assert rc == -1
