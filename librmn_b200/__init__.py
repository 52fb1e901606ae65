"""librmn_b200 -- B200-native EZ interpolator behind librmn's c_ezdefset / c_ezsint / c_ezuvint / c_ezsetopt API.

The product is ``libezb200.so`` (CUDA kernels for sm_100a + a C-ABI host layer, see ``include/ezscint_b200.h``).
This module is the thin host-side mirror a Python caller uses: it binds the C ABI with ctypes and keeps the
reference's function names, argument order and return codes (``/root/reference/src/PUBLIC_INCLUDES/rmn/ezscint.h``),
so calling code reads like code written against librmn / rpnpy.

Arrays may be numpy arrays (host memory; the library stages them through HBM) or torch CUDA tensors (device
memory, used in place).  There is no CPU implementation here: if the shared library is missing or no CUDA
device is usable, calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libezb200.so")

_i32 = C.c_int32
_f32p = C.c_void_p   # host or device pointer to float32
_i32p = C.c_void_p


def _ptr(a):
    """Raw address of a numpy array or a torch tensor (host or CUDA); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous (Fortran order of the field = C order of shape (nj, ni))")
        return a.ctypes.data
    if hasattr(a, "data_ptr"):   # torch.Tensor
        if not a.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return a.data_ptr()
    raise TypeError("expected a numpy array or a torch tensor, got %r" % type(a))


def _check_f32(*arrs):
    for a in arrs:
        if a is None:
            continue
        dt = str(a.dtype)
        if "float32" not in dt:
            raise TypeError("fields must be float32, got %s" % dt)


class EzLib:
    """ctypes binding of one shared library that exports the reference's EZ C API."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError("%s not found -- build it first (python -m librmn_b200.build)" % path)
        self.path = path
        self.lib = L = C.CDLL(path)
        L.c_ezdefset.argtypes = [_i32, _i32]
        L.c_ezsint.argtypes = [_f32p, _f32p]
        L.c_ezuvint.argtypes = [_f32p] * 4
        L.c_ezwdint.argtypes = [_f32p] * 4
        L.c_ezsetopt.argtypes = [C.c_char_p, C.c_char_p]
        L.c_ezsetval.argtypes = [C.c_char_p, C.c_float]
        L.c_ezsetival.argtypes = [C.c_char_p, _i32]
        L.c_ezgetopt.argtypes = [C.c_char_p, C.c_char_p]
        L.c_ezgetval.argtypes = [C.c_char_p, C.POINTER(C.c_float)]
        L.c_ezgetival.argtypes = [C.c_char_p, C.POINTER(_i32)]
        L.c_ezgdef_fmem.argtypes = [_i32, _i32, C.c_char_p, C.c_char_p] + [_i32] * 4 + [_f32p, _f32p]
        L.c_ezqkdef.argtypes = [_i32, _i32, C.c_char_p] + [_i32] * 5
        L.c_ezgdef_supergrid.argtypes = [_i32, _i32, C.c_char_p, C.c_char_p, _i32, _i32, C.POINTER(_i32)]
        L.c_gdrls.argtypes = [_i32]
        L.c_gdll.argtypes = [_i32, _f32p, _f32p]
        L.c_gdxyfll.argtypes = [_i32, _f32p, _f32p, _f32p, _f32p, _i32]
        L.c_gdxysval.argtypes = [_i32, _f32p, _f32p, _f32p, _f32p, _i32]
        for nm in ('c_gdllfxy', 'c_gdllsval'):
            if hasattr(L, nm):
                getattr(L, nm).argtypes = [_i32, _f32p, _f32p, _f32p, _f32p, _i32]
        for nm in ('c_gdxyvval', 'c_gdllvval', 'c_gdxywdval', 'c_gdllwdval', 'c_gdwdfuv', 'c_gduvfwd'):
            if hasattr(L, nm):
                getattr(L, nm).argtypes = [_i32] + [_f32p] * 6 + [_i32]
        L.c_ezgprm.argtypes = [_i32, C.c_char_p] + [C.POINTER(_i32)] * 6
        L.c_ezgxprm.argtypes = [_i32, C.POINTER(_i32), C.POINTER(_i32), C.c_char_p] + [C.POINTER(_i32)] * 4 + [C.c_char_p] + [C.POINTER(_i32)] * 4
        L.c_ezget_nsubgrids.argtypes = [_i32]
        L.c_ezget_subgridids.argtypes = [_i32, C.POINTER(_i32)]
        L.c_ezsint_mask.argtypes = [_i32p, _i32p]
        L.c_gdsetmask.argtypes = [_i32, _i32p]
        L.c_gdgetmask.argtypes = [_i32, _i32p]
        L.cxgaig_.argtypes = [C.c_char_p] + [C.POINTER(_i32)] * 4 + [C.POINTER(C.c_float)] * 4 + [_i32]
        L.cigaxg_.argtypes = [C.c_char_p] + [C.POINTER(C.c_float)] * 4 + [C.POINTER(_i32)] * 4 + [_i32]
        L.cxgaig_.restype = None
        L.cigaxg_.restype = None
        self.native = hasattr(L, "c_ezb200_version")
        if self.native:
            L.c_ezsint_batch.argtypes = [_f32p, _f32p, _i32]
            L.c_ezuvint_batch.argtypes = [_f32p] * 4 + [_i32]
            L.c_ezb200_set_stream.argtypes = [C.c_void_p]
            L.c_ezb200_set_arith.argtypes = [_i32]
            L.c_ezb200_get_gridset.argtypes = [_i32, _i32, _f32p, _f32p, C.c_void_p, C.c_void_p]
            L.c_ezb200_launch_count.argtypes = [_i32]
            L.c_ezb200_launch_count.restype = C.c_int64
            L.c_ezb200_version.restype = C.c_char_p
            L.c_ezb200_last_path.argtypes = [C.POINTER(_i32)]
            L.c_ezb200_set_timing.argtypes = [_i32]
            L.c_ezb200_set_rows.argtypes = [_i32, _i32]
            L.c_ezb200_set_devices.argtypes = [_i32]
            L.c_ezb200_set_direct_max_fields.argtypes = [_i32]
            L.c_ezb200_last_kernel_ms.argtypes = [C.POINTER(C.c_float), C.POINTER(_i32)]

    # ---- descriptors -------------------------------------------------------------------------
    def cxgaig(self, grtyp, xg1, xg2, xg3, xg4):
        ig = [_i32() for _ in range(4)]
        xg = [C.c_float(v) for v in (xg1, xg2, xg3, xg4)]
        self.lib.cxgaig_(grtyp.encode(), *[C.byref(i) for i in ig], *[C.byref(x) for x in xg], 1)
        return tuple(i.value for i in ig)

    def cigaxg(self, grtyp, ig1, ig2, ig3, ig4):
        ig = [_i32(v) for v in (ig1, ig2, ig3, ig4)]
        xg = [C.c_float() for _ in range(4)]
        self.lib.cigaxg_(grtyp.encode(), *[C.byref(x) for x in xg], *[C.byref(i) for i in ig], 1)
        return tuple(x.value for x in xg)

    # ---- grids -------------------------------------------------------------------------------
    def ezqkdef(self, ni, nj, grtyp, ig1, ig2, ig3, ig4, iunit=0):
        return self.lib.c_ezqkdef(ni, nj, grtyp.encode(), ig1, ig2, ig3, ig4, iunit)

    def ezgdef_fmem(self, ni, nj, grtyp, grref, ig1, ig2, ig3, ig4, ax, ay):
        if ax is not None:
            ax = np.ascontiguousarray(ax, dtype=np.float32)
            ay = np.ascontiguousarray(ay, dtype=np.float32)
        return self.lib.c_ezgdef_fmem(ni, nj, grtyp.encode(), grref.encode(), ig1, ig2, ig3, ig4, _ptr(ax), _ptr(ay))

    def ezgdef_supergrid(self, ni, nj, grtyp, grref, vercode, subgrids):
        arr = (_i32 * len(subgrids))(*subgrids)
        return self.lib.c_ezgdef_supergrid(ni, nj, grtyp.encode(), grref.encode(), vercode, len(subgrids), arr)

    def gdrls(self, gdid):
        return self.lib.c_gdrls(gdid)

    def ezgprm(self, gdid):
        t = C.create_string_buffer(4)
        v = [_i32() for _ in range(6)]
        rc = self.lib.c_ezgprm(gdid, t, *[C.byref(x) for x in v])
        return rc, t.value[:1].decode(), tuple(x.value for x in v)

    def ezgxprm(self, gdid):
        t, r = C.create_string_buffer(4), C.create_string_buffer(4)
        ni, nj = _i32(), _i32()
        ig = [_i32() for _ in range(4)]
        igr = [_i32() for _ in range(4)]
        rc = self.lib.c_ezgxprm(gdid, C.byref(ni), C.byref(nj), t, *[C.byref(x) for x in ig], r, *[C.byref(x) for x in igr])
        return rc, ni.value, nj.value, t.value[:1].decode(), tuple(x.value for x in ig), r.value[:1].decode(), tuple(x.value for x in igr)

    def gdgaxes(self, gdid, ax, ay):
        _check_f32(ax, ay)
        self.lib.c_gdgaxes.argtypes = [_i32, C.c_void_p, C.c_void_p]
        return self.lib.c_gdgaxes(gdid, _ptr(ax), _ptr(ay))

    def ezget_nsubgrids(self, gdid):
        return self.lib.c_ezget_nsubgrids(gdid)

    def ezget_subgridids(self, gdid):
        n = max(1, self.ezget_nsubgrids(gdid))
        arr = (_i32 * n)()
        self.lib.c_ezget_subgridids(gdid, arr)
        return list(arr)

    def gdll(self, gdid, lat, lon):
        _check_f32(lat, lon)
        return self.lib.c_gdll(gdid, _ptr(lat), _ptr(lon))

    def gdxyfll(self, gdid, x, y, lat, lon):
        _check_f32(x, y, lat, lon)
        return self.lib.c_gdxyfll(gdid, _ptr(x), _ptr(y), _ptr(lat), _ptr(lon), int(lat.size if isinstance(lat, np.ndarray) else lat.numel()))

    def gdxysval(self, gdin, zout, zin, x, y):
        _check_f32(zout, zin, x, y)
        return self.lib.c_gdxysval(gdin, _ptr(zout), _ptr(zin), _ptr(x), _ptr(y), int(x.size if isinstance(x, np.ndarray) else x.numel()))

    @staticmethod
    def _n(a):
        return int(a.size if isinstance(a, np.ndarray) else a.numel())

    def gdllfxy(self, gdid, lat, lon, x, y):
        _check_f32(lat, lon, x, y)
        return self.lib.c_gdllfxy(gdid, _ptr(lat), _ptr(lon), _ptr(x), _ptr(y), self._n(x))

    def gdllsval(self, gdid, zout, zin, lat, lon):
        _check_f32(zout, zin, lat, lon)
        return self.lib.c_gdllsval(gdid, _ptr(zout), _ptr(zin), _ptr(lat), _ptr(lon), self._n(lat))

    def _pts6(self, fn, gdid, a, b, c, d, e, f):
        _check_f32(a, b, c, d, e, f)
        return fn(gdid, _ptr(a), _ptr(b), _ptr(c), _ptr(d), _ptr(e), _ptr(f), self._n(e))

    def gdxyvval(self, gdin, uuout, vvout, uuin, vvin, x, y):
        return self._pts6(self.lib.c_gdxyvval, gdin, uuout, vvout, uuin, vvin, x, y)

    def gdllvval(self, gdid, uuout, vvout, uuin, vvin, lat, lon):
        return self._pts6(self.lib.c_gdllvval, gdid, uuout, vvout, uuin, vvin, lat, lon)

    def gdxywdval(self, gdin, uuout, vvout, uuin, vvin, x, y):
        return self._pts6(self.lib.c_gdxywdval, gdin, uuout, vvout, uuin, vvin, x, y)

    def gdllwdval(self, gdid, uuout, vvout, uuin, vvin, lat, lon):
        return self._pts6(self.lib.c_gdllwdval, gdid, uuout, vvout, uuin, vvin, lat, lon)

    def gdwdfuv(self, gdid, spd, wd, uuin, vvin, lat, lon):
        return self._pts6(self.lib.c_gdwdfuv, gdid, spd, wd, uuin, vvin, lat, lon)

    def gduvfwd(self, gdid, uuout, vvout, spd, wd, lat, lon):
        return self._pts6(self.lib.c_gduvfwd, gdid, uuout, vvout, spd, wd, lat, lon)

    def gdsetmask(self, gdid, mask):
        return self.lib.c_gdsetmask(gdid, _ptr(np.ascontiguousarray(mask, dtype=np.int32)))

    # ---- options -----------------------------------------------------------------------------
    def ezsetopt(self, option, value):
        return self.lib.c_ezsetopt(option.encode(), value.encode())

    def ezsetval(self, option, value):
        return self.lib.c_ezsetval(option.encode(), float(value))

    def ezsetival(self, option, value):
        return self.lib.c_ezsetival(option.encode(), int(value))

    def ezgetopt(self, option):
        buf = C.create_string_buffer(64)
        self.lib.c_ezgetopt(option.encode(), buf)
        return buf.value.decode()

    def ezgetval(self, option):
        v = C.c_float()
        self.lib.c_ezgetval(option.encode(), C.byref(v))
        return v.value

    def ezgetival(self, option):
        v = _i32()
        self.lib.c_ezgetival(option.encode(), C.byref(v))
        return v.value

    # ---- hot path ----------------------------------------------------------------------------
    def ezdefset(self, gdout, gdin):
        return self.lib.c_ezdefset(gdout, gdin)

    def ezsint(self, zout, zin):
        _check_f32(zout, zin)
        return self.lib.c_ezsint(_ptr(zout), _ptr(zin))

    def ezuvint(self, uuout, vvout, uuin, vvin):
        _check_f32(uuout, vvout, uuin, vvin)
        return self.lib.c_ezuvint(_ptr(uuout), _ptr(vvout), _ptr(uuin), _ptr(vvin))

    def ezwdint(self, spdout, wdout, uuin, vvin):
        _check_f32(spdout, wdout, uuin, vvin)
        return self.lib.c_ezwdint(_ptr(spdout), _ptr(wdout), _ptr(uuin), _ptr(vvin))

    def ezsint_mask(self, mask_out, mask_in):
        return self.lib.c_ezsint_mask(_ptr(mask_out), _ptr(mask_in))

    def ezsint_mdm(self, zout, mask_out, zin, mask_in):
        _check_f32(zout, zin)
        return self.lib.c_ezsint_mdm(C.c_void_p(_ptr(zout)), C.c_void_p(_ptr(mask_out)), C.c_void_p(_ptr(zin)), C.c_void_p(_ptr(mask_in)))

    def ezuvint_mdm(self, uuout, vvout, mask_out, uuin, vvin, mask_in):
        _check_f32(uuout, vvout, uuin, vvin)
        p = [C.c_void_p(_ptr(a)) for a in (uuout, vvout, mask_out, uuin, vvin, mask_in)]
        return self.lib.c_ezuvint_mdm(*p)

    # ---- additive (libezb200 only) -----------------------------------------------------------
    def ezsint_batch(self, zout, zin, nfields):
        _check_f32(zout, zin)
        return self.lib.c_ezsint_batch(_ptr(zout), _ptr(zin), nfields)

    def ezuvint_batch(self, uuout, vvout, uuin, vvin, nfields):
        _check_f32(uuout, vvout, uuin, vvin)
        return self.lib.c_ezuvint_batch(_ptr(uuout), _ptr(vvout), _ptr(uuin), _ptr(vvin), nfields)

    def set_stream(self, cuda_stream_handle):
        return self.lib.c_ezb200_set_stream(C.c_void_p(cuda_stream_handle))

    def set_arith(self, mode):
        return self.lib.c_ezb200_set_arith(mode)

    def set_devices(self, n=1):
        """Fan host-array batch calls of this thread out over n CUDA devices (one gridset replica each); returns the number used."""
        return self.lib.c_ezb200_set_devices(n)

    def set_direct_max_fields(self, n=4):
        """Calls of at most n fields take the direct kernel, larger ones the tile kernels (0 = always the tile kernels); returns the previous n."""
        return self.lib.c_ezb200_set_direct_max_fields(n)

    def set_rows(self, j0=-1, j1=-1):
        """Output-row band [j0, j1) of the target for this thread's following calls; no arguments = whole target again."""
        return self.lib.c_ezb200_set_rows(j0, j1)

    def get_gridset(self, gdout, gdin, npts):
        x = np.empty(npts, np.float32)
        y = np.empty(npts, np.float32)
        zb = np.empty(npts, np.uint8)
        counts = np.zeros(6, np.int32)
        rc = self.lib.c_ezb200_get_gridset(gdout, gdin, _ptr(x), _ptr(y), _ptr(zb), _ptr(counts))
        if rc != 0:
            raise RuntimeError("c_ezb200_get_gridset failed (rc=%d)" % rc)
        return x, y, zb, counts

    def launch_count(self, reset=False):
        return self.lib.c_ezb200_launch_count(1 if reset else 0)

    def last_path(self):
        """(bitmask, nlist) of the kernels the last interpolation launch used: 1 per-point, 2 window tile kernel, 4 per-point on
        the compacted list, 8 row-separable tile kernel."""
        n = _i32(0)
        return self.lib.c_ezb200_last_path(C.byref(n)), n.value

    def set_timing(self, on=True):
        return self.lib.c_ezb200_set_timing(1 if on else 0)

    def last_kernel_ms(self):
        """((rowsep_ms, window_ms, list_ms) or None, (tiles_rowsep, tiles_window, list_points)) of the last tiled launch."""
        ms = (C.c_float * 3)()
        tiles = (_i32 * 3)()
        rc = self.lib.c_ezb200_last_kernel_ms(ms, tiles)
        return (tuple(ms) if rc == 0 else None), tuple(tiles)

    def sync(self):
        return self.lib.c_ezb200_sync()

    def version(self):
        return self.lib.c_ezb200_version().decode()


_lib = None


def load(build_if_missing=False):
    """Load libezb200.so.  Raises if it is missing: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if build_if_missing and not os.path.exists(LIB_PATH):
            from . import build as _b
            _b.build()
        _lib = EzLib(os.environ.get("EZB200_LIB", LIB_PATH))   # EZB200_LIB: experiment builds (librmn_b200/build.py --out=...)
    return _lib
