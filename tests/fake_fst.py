"""An in-memory stand-in for the FST layer (c_fstinf / c_fstsui / c_fstprm / c_fstluk, rmn/fstd98.h:42-67), for tests of the
file-backed grid definitions.  Records are dicts; the same table is served to the product
(c_ezb200_set_fst_hooks) and to the oracle (oracle_set_fst_hooks) through ctypes callbacks."""
import ctypes as C

import numpy as np

_i = C.c_int
_ip = C.POINTER(C.c_int)
FSTINF = C.CFUNCTYPE(_i, _i, _ip, _ip, _ip, _i, C.c_char_p, _i, _i, _i, C.c_char_p, C.c_char_p)
FSTSUI = C.CFUNCTYPE(_i, _i, _ip, _ip, _ip)
FSTPRM = C.CFUNCTYPE(_i, _i, *([_ip] * 11), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, *([_ip] * 11))
FSTLUK = C.CFUNCTYPE(_i, C.c_void_p, _i, _ip, _ip, _ip)


class FakeFst:
    def __init__(self):
        self.recs = []          # key = index + 1
        self.cursor = None      # (criteria, next index) of the last c_fstinf, for c_fstsui
        self.inf = FSTINF(self._inf); self.sui = FSTSUI(self._sui); self.prm = FSTPRM(self._prm); self.luk = FSTLUK(self._luk)

    def add(self, nomvar, data, ip1, ip2, ip3, grtyp, ig):
        data = np.ascontiguousarray(data, np.float32)
        nj, ni = (1, data.size) if data.ndim == 1 else data.shape
        self.recs.append(dict(nomvar=nomvar.ljust(4), data=data, ni=ni, nj=nj, ip1=ip1, ip2=ip2, ip3=ip3, grtyp=grtyp, ig=tuple(ig)))
        return len(self.recs)

    def add_column(self, nomvar, data, ip1, ip2, ip3, grtyp, ig):
        """a '^^' record: ni = 1, nj = len(data)"""
        k = self.add(nomvar, data, ip1, ip2, ip3, grtyp, ig)
        self.recs[-1]["ni"], self.recs[-1]["nj"] = 1, int(np.asarray(data).size)
        return k

    def _match(self, r, crit):
        nomvar, ip1, ip2, ip3 = crit
        return r["nomvar"].strip() == nomvar.strip() and (ip1 == -1 or r["ip1"] == ip1) and (ip2 == -1 or r["ip2"] == ip2) and (ip3 == -1 or r["ip3"] == ip3)

    def _search(self, crit, start, ni, nj, nk):
        for k in range(start, len(self.recs)):
            if self._match(self.recs[k], crit):
                ni[0], nj[0], nk[0] = self.recs[k]["ni"], self.recs[k]["nj"], 1
                self.cursor = (crit, k + 1)
                return k + 1
        self.cursor = (crit, len(self.recs))
        return -1

    def _inf(self, iun, ni, nj, nk, datev, etiket, ip1, ip2, ip3, typvar, nomvar):
        return self._search((nomvar.decode(), ip1, ip2, ip3), 0, ni, nj, nk)

    def _sui(self, iun, ni, nj, nk):
        if self.cursor is None:
            return -1
        return self._search(self.cursor[0], self.cursor[1], ni, nj, nk)

    def _prm(self, key, dateo, deet, npas, ni, nj, nk, nbits, datyp, ip1, ip2, ip3, typvar, nomvar, etiket, grtyp, ig1, ig2, ig3, ig4, swa, lng,
             dltf, ubc, e1, e2, e3):
        if key < 1 or key > len(self.recs):
            return -1
        r = self.recs[key - 1]
        for p, v in ((dateo, 0), (deet, 0), (npas, 0), (ni, r["ni"]), (nj, r["nj"]), (nk, 1), (nbits, 32), (datyp, 5), (ip1, r["ip1"]), (ip2, r["ip2"]),
                     (ip3, r["ip3"]), (ig1, r["ig"][0]), (ig2, r["ig"][1]), (ig3, r["ig"][2]), (ig4, r["ig"][3]), (swa, 0), (lng, 0), (dltf, 0), (ubc, 0),
                     (e1, 0), (e2, 0), (e3, 0)):
            p[0] = v
        C.memmove(typvar, b"X \0", 3)
        C.memmove(nomvar, r["nomvar"].encode() + b"\0", 5)
        C.memmove(etiket, b"            \0", 13)
        C.memmove(grtyp, r["grtyp"].encode() + b"\0", 2)
        return 0

    def _luk(self, field, key, ni, nj, nk):
        if key < 1 or key > len(self.recs):
            return -1
        r = self.recs[key - 1]
        C.memmove(field, r["data"].ctypes.data, r["data"].nbytes)
        ni[0], nj[0], nk[0] = r["ni"], r["nj"], 1
        return key

    def hooks(self):
        return [C.cast(f, C.c_void_p) for f in (self.inf, self.sui, self.prm, self.luk)]

    def install(self, lib):
        """lib: an EzLib (product: c_ezb200_set_fst_hooks ; oracle: oracle_set_fst_hooks)"""
        L = lib.lib
        if hasattr(L, "c_ezb200_set_fst_hooks"):
            L.c_ezb200_set_fst_hooks.argtypes = [C.c_void_p] * 4
            return L.c_ezb200_set_fst_hooks(*self.hooks())
        L.oracle_set_fst_hooks.argtypes = [C.c_void_p] * 4
        L.oracle_set_fst_hooks.restype = None
        L.oracle_set_fst_hooks(*self.hooks())
        return 0
