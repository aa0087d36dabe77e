"""Inputs shared by tests/golden/make_gfortran_io.py and tests/test_gfortran_io.py."""
import numpy as np


def F16_VALUES():
    """the values of tests/test_gpu_output.py::test_f16_4_torture (same seed, same specials, same ties)"""
    m, n, l = 64, 16, 16
    rng = np.random.default_rng(11)
    vals = rng.standard_normal(m * n * l) * 10.0 ** rng.integers(-9, 13, m * n * l)
    special = [0.0, -0.0, 0.03125, -0.03125, 0.09375, 0.00005, 0.00015, 0.00025, -0.00005, 1e-300, -1e-300, 5e-324,
               0.99995, 0.999949999, 9.99995, 99999999999.0, 99999999999.99994, 99999999999.99996, -9999999999.99995,
               -99999999999.0, 1e11, 1.0e15, 1e300, -1e300, np.inf, -np.inf, np.nan, 2.0 ** -20, 12345.67895, 0.5, 1.5,
               2.5e-5, 7.5e-5, 123456789.12345, -0.00004999, 4.9999999e-5, 5.0000001e-5]
    ties = (np.arange(1, 4001) * 2 + 1) / 32.0 * 1e-3
    ties2 = np.arange(-2000, 2000) / 8.0 + 0.03125
    vals[:len(special)] = special
    vals[100:100 + len(ties)] = ties
    vals[5000:5000 + len(ties2)] = ties2
    return [float(v) for v in vals]


SELFTEST_REALS = [0.5, 1.0, 0.0, -0.0, 123456.789, 1e16, 9.9999999999999999e16, 1e17, 0.1, 0.099999, -2.5, 1e-3,
                  -1e-300, 1e300, 3.0e-5, 12.0, 100.0, 0.25, 1.0 / 3.0, 20.0 / 3.0, 1e15 + 0.5, 9.9999999999999995,
                  0.99999999999999999, 99999999999999990.0, 5e-324, 1.7976931348623157e308, 2.0e-4, 5.0e-5,
                  0.063 / 63.0, 1.7, 1300.0, 3.0e-2]
SELFTEST_INTS = [0, 1, -1, 100, 5000, 2147483647, -2147483647, 64]


def LIST_RECORDS():
    """the records of `pixelflow_driver --format-selftest`, as the reference's write(*,*) statements state them"""
    recs = [("# xnue =", v) for v in SELFTEST_REALS]                                # lib/global.f90:68
    recs += [("# SOR max iteration steps =", v) for v in SELFTEST_INTS]             # :88
    recs.append(("--- time_steps= ", 7, " --  time = ", 7 * 5.0e-5))                # ibm_3d_uniform_omp_cpu.f90:102
    recs.append(("SOR iteration no.", 100, "-- p error:", 1.2345678901234567e-3))   # :608
    recs.append(("# m, n, l =", 64, 64, 64))                                        # lib/grid.f90:311
    recs.append(("# dx, dy, dz =", 1e-3, 0.5, 12.5))                                # :313
    recs.append(("Fp =", -1.5e-3, 2.25))                                            # lib/output.f90:299
    recs.append(("Cd =", 1.25, "Cl =", -3.5e-7))                                    # :302
    recs.append((-0.315, 0.0, 0.315))                                               # rows of etc/grid.dat (:57)
    recs.append(("# istep_max= ", 2000, "   istep_out= ", 100))                     # ibm_3d_uniform_omp_cpu.f90:65
    return recs


def READ_RECORDS():
    """porosity records in the forms the reference's tools write (stl2poro '.6E' through csv.writer, voxel2poro
    'i, j, k, %.10f') and a few list-directed oddities"""
    rng = np.random.default_rng(3)
    lines = []
    for q in range(40):
        i, j, k = (int(x) for x in rng.integers(1, 600, 3))
        v = float(rng.random()) * 10.0 ** int(rng.integers(-7, 1))
        lines.append(f"{i},{j},{k},{v:.6E}")
        lines.append(f"{i}, {j}, {k}, {np.float32(v):.10f}")
    lines += ["1,1,1,1.0d0", "2,3,4,5.000000D-01", "1,1,1,+0.25", "1,1,1,.125", "1,1,1,1.", "1,1,1,7", "  3 ,\t4   5 , 0.5  ",
              "1,1,1,1e-30", "1,1,1,0.12345678901234567890", "1,1,1,0.30000000000000004", "1,1,1,6.02214076e23",
              "1,1,1,1e22", "1,1,1,1e23", "1,1,1,8.5e-23", "1,1,1,4.9e-324", "1,1,1,1.7976931348623157e+308",
              "1,1,1,abc", "1,1,1", "1.0,1,1,0.5", "1,1,1,0.5,7"]
    return lines


ODD_CONTROLDICT = """! odd but legal namelist forms
&PHYSICAL
 XNUE = 1.0d-3, xlambda=2.5E-1 ,
 density = 1 ! integer literal for a real
 width   = .63, height = +0.63 , depth = 6.3e-1
 time = 1.
 inlet_velocity = 1.500000 outlet_pressure = 0.
 AoA = -12.5
/
&file_control  istep_out = 50 /
&grid_control
istep_max = 2000
/
&porosity_control
thickness = 1.5, threshold = 1.0e-6, radius = 0.63
center_x = 0.5
center_y = 0.5, center_z = 0.25
/
&calculation_method
nonslip = F
/
&directory_control
output_folder   = 'room out'
csv_file        = "data/room.csv"
/
&solver_control
iter_max        = 100
relux_factor    = 1.7d0
/
"""


def echo_records(r):
    """the header echo of read_settings (lib/global.f90:66-90) as write(*,*) item lists, from a dict of settings"""
    return [("#",), ("# --- Physical conditions",), ("# xnue =", r["xnue"]), ("# xlambda =", r["xlambda"]),
            ("# density =", r["density"]), ("# width =", r["width"]), ("# height =", r["height"]),
            ("# depth =", r["depth"]), ("# time =", r["time"]), ("# inlet_velocity =", r["inlet_velocity"]),
            ("# outlet_pressure =", r["outlet_pressure"]), ("# Angle of inlet_velocity (AoA) =", r["AoA"]), ("#",),
            ("# --- Porosity information",), ("# thickness =", r["thickness"]), ("# threshold =", r["threshold"]),
            ("# radius =", r["radius"]), ("#",), ("# --- Directory information",),
            ("# output_folder =", r["output_folder"].ljust(50)), ("# input_porosity_file =", r["csv_file"].ljust(50)),
            ("#",), ("# --- Solver information",), ("# SOR max iteration steps =", int(r["iter_max"])),
            ("# SOR reluxation factor =", r["relux_factor"])]
