/* Runtime for the C that oracle/f90toc.py generates from the reference's Fortran sources.
 * TEST INFRASTRUCTURE (oracle side): never linked into the product library.
 * It supplies what a gfortran executable gets from libgfortran for the statements the translated
 * units use: open/close, list-directed read, namelist read, list-directed write (captured, not
 * formatted like gfortran: the byte-exact log format is pinned elsewhere, tests/test_gfortran_io.py),
 * and stubs for the subroutines that are not translated (output_*, get_now_time, system). */
#ifndef PF_REF_RUNTIME_H
#define PF_REF_RUNTIME_H
#include <math.h>
#include <stdlib.h>

typedef struct {
  const char *name;
  void *ptr;
  char type;      /* 'd' double, 'f' float, 'i' int, 'l' logical (int), 'c' character */
  int rank;
  int lo[3], hi[3];
} rt_var;

static inline double rt_sq(double x) { return x * x; }
static inline float rt_sqf(float x) { return x * x; }
static inline int rt_sqi(int x) { return x * x; }
static inline int rt_ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }
/* MAX/MIN of gfortran: compare-and-select */
static inline double rt_maxd(double a, double b) { return a > b ? a : b; }
static inline double rt_mind(double a, double b) { return a < b ? a : b; }
static inline float rt_maxf(float a, float b) { return a > b ? a : b; }
static inline float rt_minf(float a, float b) { return a < b ? a : b; }
static inline int rt_maxi(int a, int b) { return a > b ? a : b; }
static inline int rt_mini(int a, int b) { return a < b ? a : b; }

void rt_open(int unit, const char *name, int len);
void rt_str_begin(void);
void rt_str_add(const char *p, int len, int trim);
void rt_open_str(int unit, int for_write);
void rt_close(int unit);
void rt_read_begin(int unit);
void rt_read_int(int *v);
void rt_read_real(double *v);
void rt_read_real4(float *v);
void rt_read_end(void);
void rt_nml_begin(int unit, const char *group);
void rt_nml_item(const char *name, char type, void *ptr, int charlen);
void rt_nml_end(void);
void rt_write_begin(int unit);
void rt_write_begin_fmt(int unit, const char *fmt, int fmtlen);
void rt_write_begin_internal(char *buf, int buflen, const char *fmt, int fmtlen);
void rt_write_str(const char *s);
void rt_write_chars(const char *p, int len, int trim);
void rt_write_int(int v);
void rt_write_real(double v);
void rt_write_logical(int v);
void rt_write_end(void);
void rt_stub(const char *name);
void rt_stop(void);
#endif
