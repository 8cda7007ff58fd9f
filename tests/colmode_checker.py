"""CPU checker with the methods rrtmg_lw_b200.colmode.run_case asks of a backend
(clear / get_alpha / mcica / nomcica), answered by the oracle.  Test infrastructure only."""
import numpy as np


class OracleBackend:
    def __init__(self, orc):
        self.orc = orc

    def heatfac(self):
        return float(self.orc.get_table("heatfac")[0])

    def band_range(self, istart=1, iend=16):
        self.orc.set_band_range(istart, iend)

    def clear(self, b, idrv):
        o = self.orc.rrtmg_lw(b, icld=0, idrv=idrv)
        assert o["rc"] == 0
        return o

    def get_alpha(self, b, icld, idcor, decorr_con, juldat):
        return self.orc.get_alpha(b, icld, idcor=idcor, decorr_con=decorr_con, juldat=juldat)

    def nomcica(self, b, icld, idrv, flags):
        o = self.orc.rrtmg_lw_nomcica(b, icld=icld, idrv=idrv, inflglw=flags[0], iceflglw=flags[1], liqflglw=flags[2])
        assert o["rc"] == 0
        return o

    def mcica(self, b, icld, idrv, permuteseed, irng, flags, alpha):
        mc = self.orc.mcica_subcol_lw(b, icld, permuteseed=permuteseed, irng=irng, alpha=alpha)
        assert mc["rc"] == 0
        o = self.orc.rrtmg_lw(b, icld=icld, idrv=idrv, inflglw=flags[0], iceflglw=flags[1], liqflglw=flags[2], mc=mc)
        assert o["rc"] == 0
        return o
