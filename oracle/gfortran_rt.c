/* gfortran_rt.c -- TEST INFRASTRUCTURE ONLY (like everything under oracle/): the reference's own Fortran runtime,
 * driven from C.
 *
 * The reference cannot be compiled here (no Fortran compiler), but libgfortran.so.5 -- the runtime its gfortran
 * build links -- IS in the image (bundled by numpy / scipy).  The formatted and list-directed I/O statements on
 * either side of the hot path are runtime calls, so their exact behaviour can be observed by issuing the same
 * calls gfortran generates:
 *     write(65,"(3(f16.4,1x))") a, b, c           lib/output.f90:1009 ...   -> gfrt_formatted_write
 *     write(*,*) 'SOR iteration no.', iter_max, '-- p error:', error        -> gfrt_list_write
 *     read(52,*) x, y, z, poro_val                lib/grid.f90:288          -> gfrt_list_read_record
 *     read(11,nml=physical) ... read(11,nml=solver_control)   lib/global.f90:47-62   -> gfrt_read_settings
 * The st_parameter_dt / st_parameter_open blocks are filled at the offsets of libgfortran's io.h for the
 * GFORTRAN_8 ABI (gfc_charlen_type = size_t, x86-64):
 *   common: flags@0 unit@4 filename@8 line@16 iomsg_len@24 iomsg@32 iostat@40
 *   dt    : rec@48 size@56 iolength@64 internal_unit_desc@72 format@80 format_len@88 advance_len@96 advance@104
 *           internal_unit@112 internal_unit_len@120 ...
 *   open  : recl_in@48 file_len@56 file@64 ...
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define GFRT_EXPORT __attribute__((visibility("default")))

typedef void (*st_fn)(void *);
typedef void (*tr_fn)(void *, void *, int);
typedef void (*trc_fn)(void *, void *, size_t);

static struct {
  void *lib;
  st_fn st_open, st_close, st_write, st_write_done, st_read, st_read_done;
  tr_fn real_w, int_w, logical_w, real_r, int_r;
  trc_fn char_w;
  void *set_nml;
} G;

static unsigned char blk[4096] __attribute__((aligned(16)));

enum { DT_NAMELIST_READ_MODE = 1 << 8, DT_HAS_NAMELIST_NAME = 1 << 15,
       DT_LIST_FORMAT = 1 << 7, DT_HAS_FORMAT = 1 << 12, DT_HAS_INTERNAL_UNIT = 1 << 14, HAS_IOSTAT = 1 << 5,
       OPEN_HAS_FILE = 1 << 8 };

static void common(int flags, int unit) {
  memset(blk, 0, sizeof blk);
  *(int32_t *)(blk + 0) = flags;
  *(int32_t *)(blk + 4) = unit;
  *(const char **)(blk + 8) = "gfortran_rt.c";
  *(int32_t *)(blk + 16) = 1;
}

GFRT_EXPORT int gfrt_open(const char *libgfortran_path) {
  if (G.lib) return 0;
  G.lib = dlopen(libgfortran_path, RTLD_NOW | RTLD_GLOBAL);
  if (!G.lib) return 1;
#define SYM(field, type, name) G.field = (type)dlsym(G.lib, name); if (!G.field) return 2;
  SYM(st_open, st_fn, "_gfortran_st_open") SYM(st_close, st_fn, "_gfortran_st_close")
  SYM(st_write, st_fn, "_gfortran_st_write") SYM(st_write_done, st_fn, "_gfortran_st_write_done")
  SYM(st_read, st_fn, "_gfortran_st_read") SYM(st_read_done, st_fn, "_gfortran_st_read_done")
  SYM(real_w, tr_fn, "_gfortran_transfer_real_write") SYM(int_w, tr_fn, "_gfortran_transfer_integer_write")
  SYM(logical_w, tr_fn, "_gfortran_transfer_logical_write") SYM(char_w, trc_fn, "_gfortran_transfer_character_write")
  SYM(real_r, tr_fn, "_gfortran_transfer_real") SYM(int_r, tr_fn, "_gfortran_transfer_integer")
  SYM(set_nml, void *, "_gfortran_st_set_nml_var")
#undef SYM
  return 0;
}

static char tmpname[64];
static void unit_open(void) {
  snprintf(tmpname, sizeof tmpname, "/tmp/gfrt_%d.txt", (int)getpid());
  remove(tmpname);
  common(OPEN_HAS_FILE, 65);
  *(size_t *)(blk + 56) = strlen(tmpname);
  *(const char **)(blk + 64) = tmpname;
  G.st_open(blk);
}
static int unit_close_and_fetch(char *out, int cap) {
  common(0, 65);
  G.st_close(blk);
  FILE *f = fopen(tmpname, "rb");
  if (!f) return -1;
  const int n = (int)fread(out, 1, (size_t)cap, f);
  fclose(f);
  remove(tmpname);
  return n;
}

/* nrec records, each `write(65, fmt) vals[r*per_rec .. +per_rec)`; returns the file's bytes (newlines included) */
GFRT_EXPORT int gfrt_formatted_write(const char *fmt, const double *vals, int per_rec, int nrec, char *out, int cap) {
  if (!G.lib) return -2;
  unit_open();
  for (int r = 0; r < nrec; ++r) {
    common(DT_HAS_FORMAT, 65);
    *(const char **)(blk + 80) = fmt;
    *(size_t *)(blk + 88) = strlen(fmt);
    G.st_write(blk);
    for (int q = 0; q < per_rec; ++q) G.real_w(blk, (void *)&vals[(size_t)r * per_rec + q], 8);
    G.st_write_done(blk);
  }
  return unit_close_and_fetch(out, cap);
}

/* one `write(65,*) item, item, ...`: kinds[i] = 0 character (next of strs), 1 integer(4) (next of ints),
 * 2 real(8) (next of reals), 3 logical(4) (next of ints) */
GFRT_EXPORT int gfrt_list_write(const int *kinds, int nitems, const char *const *strs, const int *ints,
                                const double *reals, char *out, int cap) {
  if (!G.lib) return -2;
  unit_open();
  common(DT_LIST_FORMAT, 65);
  G.st_write(blk);
  int is = 0, ii = 0, ir = 0;
  for (int q = 0; q < nitems; ++q) {
    if (kinds[q] == 0) { G.char_w(blk, (void *)strs[is], strlen(strs[is])); ++is; }
    else if (kinds[q] == 1) { G.int_w(blk, (void *)&ints[ii++], 4); }
    else if (kinds[q] == 2) { G.real_w(blk, (void *)&reals[ir++], 8); }
    else { G.logical_w(blk, (void *)&ints[ii++], 4); }
  }
  G.st_write_done(blk);
  return unit_close_and_fetch(out, cap);
}

/* `read(line,*) x, y, z, v` on an internal unit; returns iostat (0 = ok) */
GFRT_EXPORT int gfrt_list_read_record(const char *line, int len, int *xyz, double *v) {
  if (!G.lib) return -2;
  int ios = 0;
  common(DT_LIST_FORMAT | DT_HAS_INTERNAL_UNIT | HAS_IOSTAT, -1);
  *(int32_t **)(blk + 40) = &ios;
  *(char **)(blk + 112) = (char *)line;
  *(size_t *)(blk + 120) = (size_t)len;
  G.st_read(blk);
  G.int_r(blk, &xyz[0], 4);
  G.int_r(blk, &xyz[1], 4);
  G.int_r(blk, &xyz[2], 4);
  G.real_r(blk, v, 8);
  G.st_read_done(blk);
  return ios;
}

/* ---- namelist input: the reference's read_settings (lib/global.f90:47-62) ---------------------------------------
 * gfortran registers every namelist object with _gfortran_st_set_nml_var(dtp, addr, name, kind, string_length, dtype)
 * BEFORE calling _gfortran_st_read (trans-io.c: build_dt); dtype is passed by value (GFORTRAN_8: elem_len, version,
 * rank, type, attribute).  namelist_name_len@128, namelist_name@136 in st_parameter_dt. */
typedef struct { size_t elem_len; int version; signed char rank; signed char type; signed short attribute; } gfrt_dtype;
typedef void (*nml_fn)(void *, void *, char *, int32_t, size_t, gfrt_dtype);
enum { BT_INTEGER = 1, BT_LOGICAL = 2, BT_REAL = 3, BT_CHARACTER = 6 };

static int nml_ios;
static void nml_begin(const char *group) {
  common(DT_HAS_NAMELIST_NAME | DT_NAMELIST_READ_MODE | HAS_IOSTAT, 11);
  *(int32_t **)(blk + 40) = &nml_ios;
  *(size_t *)(blk + 128) = strlen(group);
  *(const char **)(blk + 136) = group;
}
static void nml_var(void *addr, const char *name, int type, int kind, size_t slen) {
  gfrt_dtype d;
  memset(&d, 0, sizeof d);
  d.elem_len = slen ? slen : (size_t)kind;
  d.type = (signed char)type;
  ((nml_fn)G.set_nml)(blk, addr, (char *)name, kind, slen, d);
}
static int nml_end(void) {
  G.st_read(blk);
  G.st_read_done(blk);
  return nml_ios;
}

/* reals[19] = xnue xlambda density width height depth time inlet_velocity outlet_pressure AoA thickness threshold
 * radius center_x center_y center_z relux_factor (17 used), ints[4] = istep_out istep_max nonslip iter_max,
 * folder / csv = 50 blank-padded characters each.  Returns 0, or 100*group + 1 if a READ fails (iostat in *iostat). */
GFRT_EXPORT int gfrt_read_settings(const char *path, double *reals, int *ints, char *folder, char *csv, int *iostat) {
  if (!G.lib) return -2;
  common(OPEN_HAS_FILE, 11);
  *(size_t *)(blk + 56) = strlen(path);
  *(const char **)(blk + 64) = path;
  G.st_open(blk);
  int rc = 0, g = 0;
#define DONE() do { ++g; if (!rc && nml_end() != 0) { rc = 100 * g + 1; *iostat = nml_ios; } } while (0)
  static const char *phys[] = {"xnue", "xlambda", "density", "width", "height", "depth", "time", "inlet_velocity",
                               "outlet_pressure", "aoa"};
  nml_begin("physical");
  for (int q = 0; q < 10; ++q) nml_var(&reals[q], phys[q], BT_REAL, 8, 0);
  DONE();
  if (!rc) { nml_begin("file_control"); nml_var(&ints[0], "istep_out", BT_INTEGER, 4, 0); DONE(); }
  if (!rc) { nml_begin("grid_control"); nml_var(&ints[1], "istep_max", BT_INTEGER, 4, 0); DONE(); }
  if (!rc) {
    static const char *por[] = {"thickness", "threshold", "radius", "center_x", "center_y", "center_z"};
    nml_begin("porosity_control");
    for (int q = 0; q < 6; ++q) nml_var(&reals[10 + q], por[q], BT_REAL, 8, 0);
    DONE();
  }
  if (!rc) { nml_begin("calculation_method"); nml_var(&ints[2], "nonslip", BT_LOGICAL, 4, 0); DONE(); }
  if (!rc) {
    nml_begin("directory_control");
    nml_var(folder, "output_folder", BT_CHARACTER, 1, 50);
    nml_var(csv, "csv_file", BT_CHARACTER, 1, 50);
    DONE();
  }
  if (!rc) {
    nml_begin("solver_control");
    nml_var(&ints[3], "iter_max", BT_INTEGER, 4, 0);
    nml_var(&reals[16], "relux_factor", BT_REAL, 8, 0);
    DONE();
  }
#undef DONE
  common(0, 11);
  G.st_close(blk);
  return rc;
}
