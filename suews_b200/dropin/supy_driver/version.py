# the reference's supy_driver/version.py carries the SUEWS version the driver was built from ('2020b8'); this drop-in
# restates 2020b (Test/BaseRun/2020b goldens) and says which engine serves it
__version__ = '2020b8+b200'
