"""``supy_driver.suews_driver``: the f2py extension module of the reference, served by libsuews_b200.so."""
from suews_b200.supy_driver import (  # noqa: F401
    suews_driver, narp_module, atmmoiststab_module, resist_module, evap_module, snow_module, dailystate_module,
    lumps_module, anemsn_module, rsl_module, co2_module, waterdist_module, meteo)
