"""Thermodynamic methods shared by the grid classes -- the dispatch layer of
meteotools/grids/thermaldynamic_calculation.py:7-133 with the callee swapped for the device
kernels.  Which formula runs is decided, as in the reference, by which attributes the grid
currently holds (``'T' in dir(self)`` ...), including its quirks:

* ``calc_rho`` ALWAYS recomputes Tv (the reference tests ``'self.Tv' not in dir(self)``, :91);
* ``calc_moist_dTdz`` returns the DRY rate when ``qvs`` is present (:119-121).

Operands are handed to ``meteotools_b200.calc`` as CUDA tensors, so a chain of calls never
leaves the device; ``calc_all_thermaldynamic`` runs the whole set in one fused pass.
"""
from .. import _device as D
from .. import _lib
from ..calc import thermaldynamic as td


class calc_thermaldynamic:
    def _d(self, name):
        return self.device(name)

    def calc_theta(self):
        if self._has('T', 'p'):
            self._put('Th', td.calc_theta(self._d('T'), self._d('p')))
        else:
            raise AttributeError('T and p must be set first')

    def calc_T(self):
        if self._has('Th', 'p'):
            self._put('T', td.calc_T(self._d('Th'), self._d('p')))
        else:
            raise AttributeError('Th and p must be set first')

    def calc_saturated_vapor(self):
        if not self._has('T'):
            self.calc_T()
        self._put('vapor_s', td.calc_saturated_vapor(self._d('T')))

    def calc_saturated_qv(self):
        if not self._has('T'):
            self.calc_T()
        if self._has('p'):
            self._put('qvs', td.calc_saturated_qv(self._d('T'), self._d('p')))
        else:
            raise AttributeError('p must be set first')

    def calc_vapor(self):
        if self._has('Td'):
            self._put('vapor', td.calc_vapor(self._d('Td')))
        elif self._has('qv', 'p'):
            self._put('vapor', td.qv_to_vapor(self._d('p'), self._d('qv')))
        elif self._has('T', 'RH'):
            self._put('vapor', td.calc_vapor(self._d('T'), self._d('RH')))
        elif self._has('Th', 'RH', 'p'):
            self.calc_T()
            self._put('vapor', td.calc_vapor(self._d('T'), self._d('RH')))
        else:
            raise AttributeError('one of (Th, p, RH), (p, qv), (Td) or (T, RH) must be set first')

    def calc_Td(self):
        if not self._has('vapor'):
            self.calc_vapor()
        self._put('Td', td.calc_Td(es=self._d('vapor')))

    def calc_qv(self):
        if not self._has('vapor'):
            self.calc_vapor()
        if self._has('p'):
            self._put('qv', td.calc_qv(self._d('p'), vapor=self._d('vapor')))
        else:
            raise AttributeError('p must be set first')

    def calc_theta_es(self):
        if not self._has('T'):
            self.calc_T()
        if self._has('p'):
            self._put('Th_es', td.calc_theta_es(self._d('T'), self._d('p')))
        else:
            raise AttributeError('p must be set first')

    def calc_theta_v(self):
        if not self._has('Th'):
            self.calc_theta()
        if not self._has('qv'):
            self.calc_qv()
        if self._has('qc', 'qr'):
            ql = self._add(self._d('qc'), self._d('qr'))
        elif self._has('qc'):
            ql = self._d('qc')
        elif self._has('qr'):
            ql = self._d('qr')
        else:
            ql = 0
        self._put('Th_v', td.calc_theta_v(theta=self._d('Th'), qv=self._d('qv'), ql=ql))

    @staticmethod
    def _add(a, b):
        """ql = qc + qr as one rounded IEEE sum on the device (:70-72)."""
        return td._pw("ADD", a, b)

    def calc_Tv(self):
        if not self._has('T'):
            self.calc_T()
        if not self._has('qv'):
            self.calc_qv()
        self._put('Tv', td.calc_Tv(self._d('T'), qv=self._d('qv')))

    def calc_rho(self):
        if 'self.Tv' not in dir(self):  # sic (thermaldynamic_calculation.py:91): always true
            self.calc_Tv()
        if self._has('p'):
            self._put('rho', td.calc_rho(self._d('p'), Tv=self._d('Tv')))
        else:
            raise AttributeError('p must be set first')

    def calc_Tc(self):
        if not self._has('T'):
            self.calc_T()
        if not self._has('qv'):
            self.calc_qv()
        if self._has('p'):
            self._put('Tc', td.calc_Tc(self._d('T'), self._d('p'), qv=self._d('qv')))
        else:
            raise AttributeError('p must be set first')

    def calc_theta_e(self):
        if not self._has('Th'):
            self.calc_theta()
        if not self._has('qv'):
            self.calc_qv()
        if not self._has('Tc'):
            self.calc_Tc()
        self._put('Th_e', td.calc_theta_e(theta=self._d('Th'), qv=self._d('qv'), Tc=self._d('Tc')))

    def calc_moist_dTdz(self):
        if not self._has('T'):
            self.calc_T()
        if self._has('p', 'qvs'):
            self.moist_dTdz = td.calc_dTdz(qvs=self._d('qvs'))  # dry rate, as the reference (:120-121)
        else:
            self._put('moist_dTdz', td.calc_dTdz(T=self._d('T'), P=self._d('p')))

    def calc_dry_dTdz(self):
        self.dry_dTdz = td.calc_dTdz()

    def calc_RH(self):
        if not self._has('T'):
            self.calc_T()
        if not self._has('vapor'):
            self.calc_vapor()
        self._put('RH', td.calc_RH(T=self._d('T'), vapor=self._d('vapor')))

    # ---- fused set (F4): every thermodynamic diagnostic in one pass over T, p, qv(, qc, qr) ----
    def calc_all_thermaldynamic(self, outputs=None):
        """Compute the cyl_td set (test_cyl_thermaldynamics.py:96-260) with ``mt_cyl_td``:
        Th, vapor_s, qvs, vapor, RH, Td, Th_es, Th_v, Tv, rho, Tc, Th_e, moist_dTdz.
        Requires T, p, qv; uses qc / qr when present.  Results equal the per-method chain."""
        if not self._has('T', 'p', 'qv'):
            raise AttributeError('T, p and qv must be set first')
        names = list(outputs) if outputs is not None else list(_lib.TD_OUT)
        for dep in ('Tc', 'Th'):   # Th_e is produced from the Tc and Th buffers of the same call
            if 'Th_e' in names and dep not in names:
                names.append(dep)
        tin = _lib.mt_td_in_t()
        for k in _lib.TD_IN:
            setattr(tin, k, self._d(k).data_ptr() if self._has(k) else None)
        tout = _lib.mt_td_out_t()
        bufs = {}
        for k in names:
            if k not in _lib.TD_OUT:
                raise ValueError(f'unknown thermodynamic output {k!r}')
            bufs[k] = self._new()
            setattr(tout, k, bufs[k].data_ptr())
        n = self._d('T').numel()
        scratch = D.empty((16 + n,))
        _lib.check(_lib.load().mt_cyl_td(tin, tout, n, D.ptr(scratch), D.stream()), 'mt_cyl_td')
        for k, b in bufs.items():
            self._put(k, b)
        self._tc_scratch = scratch
        return self
