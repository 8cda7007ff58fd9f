"""rrtmg_lw_b200 -- B200-native (sm_100a) implementation of the RRTMG_LW per-column longwave
hot path behind the reference's GCM entry points.  See DESIGN.md."""
from . import api, kdata, synth  # noqa: F401
from .api import (get_alpha, mcica_subcol_lw, rrtmg_lw, rrtmg_lw_ini, rrtmg_lw_mcica)  # noqa: F401

__all__ = ["rrtmg_lw_ini", "rrtmg_lw", "rrtmg_lw_mcica", "mcica_subcol_lw", "get_alpha", "api", "kdata", "synth"]
