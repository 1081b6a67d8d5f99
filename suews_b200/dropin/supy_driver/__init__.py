"""Drop-in for the reference's f2py package ``supy_driver`` (supy-driver/supy_driver/__init__.py:13-27).

Put this directory's parent on PYTHONPATH ahead of (or instead of) the f2py build:

    PYTHONPATH=/path/to/suews_b200/dropin python -c "import supy_driver as sd; print(sd.suews_driver.output_name_n(3))"

and ``from supy_driver import suews_driver as sd`` -- the import supy makes -- resolves to the CUDA library
(libsuews_b200.so through ctypes; no f2py, no Fortran, no PyTorch).  Same names as the reference package exports.
"""
from suews_b200.supy_driver import (  # noqa: F401
    suews_driver,
    narp_module,
    atmmoiststab_module,
    resist_module,
    evap_module,
    snow_module,
    dailystate_module,
    lumps_module,
    anemsn_module,
    rsl_module,
    co2_module,
    waterdist_module,
    meteo,
)
from .version import __version__  # noqa: F401
