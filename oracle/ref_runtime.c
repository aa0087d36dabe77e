/* See ref_runtime.h.  TEST INFRASTRUCTURE — the product never links this.
 *
 * Exports for the Python harness (oracle/ref_translated.py):
 *   int   ref_run(const char *workdir)      chdir(workdir), run the translated `program main`, chdir back;
 *                                           returns 0, or 1 if the program stopped on an I/O error
 *   const rt_var *ref_lookup(name)          program / module variable by its Fortran name
 *   int   ref_perr_count(); double ref_perr(i)   the reals written on the 'p error' log lines, in order
 *   const char *ref_log()                   everything the program wrote to unit * (plain, not gfortran-spaced)
 *   void  ref_set_threads(n); int ref_max_threads()   OpenMP flavour: omp_set_num_threads / omp_get_max_threads (a
 *                                           launcher such as torchrun exports OMP_NUM_THREADS=1)
 *   void  ref_set_verbose(int)              echo the log to stdout while running
 *   int   ref_stub_count(name)              how often an untranslated subroutine was called
 *   int   ref_step_count(); double ref_step_time(i); double ref_end_time()
 *                                           CLOCK_MONOTONIC seconds at which step i (0-based) announced itself and at
 *                                           which the program returned: per-step wall times for the CPU baseline
 *   int   ref_use_libgfortran(path)         from now on every OPEN, READ (list-directed, namelist) and WRITE
 *                                           (list-directed, formatted, internal) is executed by libgfortran.so.5 — the runtime library a
 *                                           gfortran build of the reference links (it ships inside numpy/scipy) — by
 *                                           issuing the calls gfortran generates (_gfortran_st_write, ...; parameter
 *                                           blocks at the GFORTRAN_8 offsets, see oracle/gfortran_rt.c).  Unit * is
 *                                           connected to <workdir>/stdout.log.  Needed by the flavour that translates
 *                                           lib/output.f90 (formatted writes); without it such writes stop the run.
 *   void  ref_set_step_limit(n)             leave the time loop when step n+1 announces itself ('--- time_steps=' line):
 *                                           a shipped deck (5000 steps) can be run unmodified for its first n steps;
 *                                           the fields are then those at the end of step n (0 = no limit)
 */
#include "ref_runtime.h"
#include <ctype.h>
#include <dlfcn.h>
#include <setjmp.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

extern const rt_var rt_registry[];
void f_MAIN(void);
void rt_reset_statics(void);

static FILE *units[100];
static jmp_buf stop_env;
static char errmsg[512];
static int verbose = 0;
static int step_limit = 0, steps_seen = 0, w_is_step = 0, stopped_by_limit = 0;
#define MAXSTEPT 4096
static double step_t[MAXSTEPT], end_t;
static int nstep_t;
static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

static char *logbuf = 0;
static size_t loglen = 0, logcap = 0;
static double *perr = 0;
static int nperr = 0, perrcap = 0;

static void fail(const char *msg) {
  snprintf(errmsg, sizeof errmsg, "%s", msg);
  longjmp(stop_env, 1);
}

/* ------------------------------------------------------------------ libgfortran backend (optional) */
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
/* _gfortran_st_set_nml_var(dtp, addr, name, kind, string_length, dtype): dtype by value (GFORTRAN_8) */
typedef struct { size_t elem_len; int version; signed char rank; signed char type; signed short attribute; } gf_dtype;
typedef void (*nml_fn)(void *, void *, char *, int32_t, size_t, gf_dtype);
enum { BT_INTEGER = 1, BT_LOGICAL = 2, BT_REAL = 3, BT_CHARACTER = 6 };
static int gf_active = 0;
static unsigned char dtblk[4096] __attribute__((aligned(16)));
static unsigned char opblk[1024] __attribute__((aligned(16)));
static unsigned char gf_unit[100];      /* units opened through libgfortran */
static int w_gf = 0;                    /* the write statement in progress goes through libgfortran */
enum { DT_LIST_FORMAT = 1 << 7, DT_HAS_FORMAT = 1 << 12, DT_HAS_INTERNAL_UNIT = 1 << 14, OPEN_HAS_FILE = 1 << 8,
       DT_NAMELIST_READ_MODE = 1 << 8, DT_HAS_NAMELIST_NAME = 1 << 15, HAS_IOSTAT = 1 << 5 };
static int r_gf = 0;        /* the read statement in progress goes through libgfortran */
static int32_t r_ios;

static void gf_common(unsigned char *blk, size_t size, int flags, int unit) {
  memset(blk, 0, size);
  *(int32_t *)(blk + 0) = flags;
  *(int32_t *)(blk + 4) = unit;
  *(const char **)(blk + 8) = "translated reference";
  *(int32_t *)(blk + 16) = 1;
}

int ref_use_libgfortran(const char *path) {
  if (G.lib) { gf_active = 1; return 0; }
  G.lib = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!G.lib) return 1;
#define SYM(field, type, name) G.field = (type)dlsym(G.lib, name); if (!G.field) return 2;
  SYM(st_open, st_fn, "_gfortran_st_open") SYM(st_close, st_fn, "_gfortran_st_close")
  SYM(st_write, st_fn, "_gfortran_st_write") SYM(st_write_done, st_fn, "_gfortran_st_write_done")
  SYM(real_w, tr_fn, "_gfortran_transfer_real_write") SYM(int_w, tr_fn, "_gfortran_transfer_integer_write")
  SYM(logical_w, tr_fn, "_gfortran_transfer_logical_write") SYM(char_w, trc_fn, "_gfortran_transfer_character_write")
  SYM(st_read, st_fn, "_gfortran_st_read") SYM(st_read_done, st_fn, "_gfortran_st_read_done")
  SYM(real_r, tr_fn, "_gfortran_transfer_real") SYM(int_r, tr_fn, "_gfortran_transfer_integer")
  SYM(set_nml, void *, "_gfortran_st_set_nml_var")
#undef SYM
  gf_active = 1;
  return 0;
}

static void gf_open(int unit, const char *path) {
  gf_common(opblk, sizeof opblk, OPEN_HAS_FILE, unit);
  *(size_t *)(opblk + 56) = strlen(path);
  *(const char **)(opblk + 64) = path;
  G.st_open(opblk);
  if (unit >= 0 && unit < 100) gf_unit[unit] = 1;
}

static void gf_close(int unit) {
  gf_common(opblk, sizeof opblk, 0, unit);
  G.st_close(opblk);
  if (unit >= 0 && unit < 100) gf_unit[unit] = 0;
}

/* ------------------------------------------------------------------ open / close */
static char strbuf[2048];
static int strlen_;

void rt_str_begin(void) { strlen_ = 0; }
void rt_str_add(const char *p, int len, int trim) {
  if (trim) while (len > 0 && p[len - 1] == ' ') len--;
  if (strlen_ + len >= (int)sizeof strbuf) fail("character expression too long");
  memcpy(strbuf + strlen_, p, (size_t)len);
  strlen_ += len;
}
void rt_open_str(int unit, int for_write) {
  if (!for_write) { rt_open(unit, strbuf, strlen_); return; }
  /* OPEN trims trailing blanks of FILE= (F2008 9.5.6.10) */
  while (strlen_ > 0 && strbuf[strlen_ - 1] == ' ') strlen_--;
  strbuf[strlen_] = 0;
  if (!gf_active) fail("open for writing needs the libgfortran backend (ref_use_libgfortran)");
  gf_open(unit, strbuf);
}

void rt_open(int unit, const char *name, int len) {
  static char path[1024];
  while (len > 0 && name[len - 1] == ' ') len--;      /* trailing blanks of a character variable */
  int s = 0;
  while (s < len && name[s] == ' ') s++;
  if (len - s >= (int)sizeof path) fail("rt_open: name too long");
  memcpy(path, name + s, (size_t)(len - s));
  path[len - s] = 0;
  if (unit < 0 || unit >= 100) fail("rt_open: bad unit");
  if (gf_active) {
    /* the reference's runtime opens and reads the file; a missing file is reported here (libgfortran would abort) */
    if (access(path, R_OK) != 0) {
      char m[1200];
      snprintf(m, sizeof m, "rt_open: cannot open '%s'", path);
      fail(m);
    }
    gf_open(unit, path);
    return;
  }
  units[unit] = fopen(path, "r");
  if (!units[unit]) {
    char m[1200];
    snprintf(m, sizeof m, "rt_open: cannot open '%s'", path);
    fail(m);
  }
}

void rt_close(int unit) {
  if (unit >= 0 && unit < 100 && units[unit]) { fclose(units[unit]); units[unit] = 0; }
  if (unit >= 0 && unit < 100 && gf_unit[unit]) gf_close(unit);
}

/* ------------------------------------------------------------------ list-directed read
 * Items are separated by a comma (with optional blanks) or by blanks; a read statement starts on a
 * new record and continues onto following records while items are missing (Fortran 2008 10.10.3). */
static FILE *rd;
static char *rline = 0;
static size_t rcap = 0;
static char *rpos;

static int next_record(void) {
  ssize_t n = getline(&rline, &rcap, rd);
  if (n < 0) return 0;
  rpos = rline;
  return 1;
}

void rt_read_begin(int unit) {
  if (gf_active && unit >= 0 && unit < 100 && gf_unit[unit]) {
    gf_common(dtblk, sizeof dtblk, DT_LIST_FORMAT | HAS_IOSTAT, unit);
    r_ios = 0;
    *(int32_t **)(dtblk + 40) = &r_ios;
    G.st_read(dtblk);
    r_gf = 1;
    return;
  }
  r_gf = 0;
  if (unit < 0 || unit >= 100 || !units[unit]) fail("read: unit not open");
  rd = units[unit];
  if (!next_record()) fail("read: end of file");
}

static char *next_item(void) {
  for (;;) {
    while (*rpos == ' ' || *rpos == '\t' || *rpos == '\r' || *rpos == '\n') rpos++;
    if (*rpos) break;
    if (!next_record()) fail("read: end of file inside a record list");
  }
  char *start = rpos;
  while (*rpos && *rpos != ',' && *rpos != ' ' && *rpos != '\t' && *rpos != '\r' && *rpos != '\n') rpos++;
  char *end = rpos;
  /* consume the separator: blanks, then at most one comma */
  while (*rpos == ' ' || *rpos == '\t') rpos++;
  if (*rpos == ',') rpos++;
  *end = 0;   /* if the item ran to the end of the record, rpos == end and the next call fetches a record */
  return start;
}

void rt_read_int(int *v) {
  if (r_gf) { G.int_r(dtblk, v, 4); return; }
  char *t = next_item(), *e;
  long x = strtol(t, &e, 10);
  if (e == t || *e) fail("read: bad integer");
  *v = (int)x;
}

void rt_read_real(double *v) {
  if (r_gf) { G.real_r(dtblk, v, 8); return; }
  char *t = next_item(), *e;
  for (char *c = t; *c; c++) if (*c == 'd' || *c == 'D') *c = 'e';
  double x = strtod(t, &e);
  if (e == t || *e) fail("read: bad real");
  *v = x;
}

void rt_read_real4(float *v) {
  if (r_gf) { G.real_r(dtblk, v, 4); return; }
  char *t = next_item(), *e;
  for (char *c = t; *c; c++) if (*c == 'd' || *c == 'D') *c = 'e';
  float x = strtof(t, &e);       /* decimal -> binary32 in one rounding, as a real(4) read does */
  if (e == t || *e) fail("read: bad real");
  *v = x;
}

void rt_read_end(void) {
  if (!r_gf) return;
  G.st_read_done(dtblk);
  r_gf = 0;
  if (r_ios != 0) fail("read: libgfortran reports an error (iostat /= 0)");
}

/* ------------------------------------------------------------------ namelist read
 * read(u,nml=g): search forward for '&g', then 'name = value' pairs up to '/'.  Names are case-insensitive;
 * values: reals (d exponents allowed), integers, .true./.false./T/F, quoted strings. */
static char *nml_text = 0;

static int nml_gf = 0;

void rt_nml_begin(int unit, const char *group) {
  if (gf_active && unit >= 0 && unit < 100 && gf_unit[unit]) {
    gf_common(dtblk, sizeof dtblk, DT_HAS_NAMELIST_NAME | DT_NAMELIST_READ_MODE | HAS_IOSTAT, unit);
    r_ios = 0;
    *(int32_t **)(dtblk + 40) = &r_ios;
    *(size_t *)(dtblk + 128) = strlen(group);
    *(const char **)(dtblk + 136) = group;
    nml_gf = 1;
    return;
  }
  nml_gf = 0;
  if (unit < 0 || unit >= 100 || !units[unit]) fail("namelist read: unit not open");
  FILE *f = units[unit];
  size_t cap = 1024, len = 0;
  free(nml_text);
  nml_text = malloc(cap);
  int c, state = 0;   /* 0: looking for &group, 1: inside */
  size_t glen = strlen(group);
  /* find "&group" */
  for (;;) {
    c = fgetc(f);
    if (c == EOF) fail("namelist read: group not found");
    if (c == '!') { while ((c = fgetc(f)) != EOF && c != '\n') {} continue; }
    if (c != '&') continue;
    char name[128];
    size_t k = 0;
    while ((c = fgetc(f)) != EOF && (isalnum(c) || c == '_') && k < sizeof name - 1) name[k++] = (char)tolower(c);
    name[k] = 0;
    if (k == glen && strncmp(name, group, glen) == 0) { state = 1; break; }
    /* gfortran reads namelist groups in file order: a different group is skipped */
  }
  (void)state;
  int q = 0;
  for (;;) {
    if (c == EOF) fail("namelist read: unterminated group");
    if (q) { if (c == q) q = 0; }
    else if (c == '\'' || c == '"') q = c;
    else if (c == '!') { while ((c = fgetc(f)) != EOF && c != '\n') {} c = '\n'; }
    else if (c == '/') break;
    if (len + 2 > cap) { cap *= 2; nml_text = realloc(nml_text, cap); }
    nml_text[len++] = (char)c;
    c = fgetc(f);
  }
  nml_text[len] = 0;
}

void rt_nml_item(const char *name, char type, void *ptr, int charlen) {
  if (nml_gf) {
    /* what gfortran-generated code does before _gfortran_st_read: register every object of the group */
    gf_dtype d;
    memset(&d, 0, sizeof d);
    const int bt = type == 'c' ? BT_CHARACTER : type == 'l' ? BT_LOGICAL : type == 'i' ? BT_INTEGER : BT_REAL;
    const int kind = type == 'c' ? 1 : type == 'd' ? 8 : 4;
    const size_t slen = type == 'c' ? (size_t)charlen : 0;
    d.elem_len = slen ? slen : (size_t)kind;
    d.type = (signed char)bt;
    ((nml_fn)G.set_nml)(dtblk, ptr, (char *)name, kind, slen, d);
    return;
  }
  /* find `name` followed by optional blanks and '=' outside quotes, case-insensitively */
  size_t nl = strlen(name);
  int q = 0;
  for (char *p = nml_text; *p; p++) {
    if (q) { if (*p == q) q = 0; continue; }
    if (*p == '\'' || *p == '"') { q = *p; continue; }
    if ((p == nml_text || !(isalnum((unsigned char)p[-1]) || p[-1] == '_')) && strncasecmp(p, name, nl) == 0) {
      char *e = p + nl;
      while (*e == ' ' || *e == '\t') e++;
      if (*e != '=') continue;
      e++;
      while (*e == ' ' || *e == '\t') e++;
      if (type == 'c') {
        char qq = *e;
        if (qq != '\'' && qq != '"') fail("namelist: character value must be quoted");
        char *s = e + 1, *t = strchr(s, qq);
        if (!t) fail("namelist: unterminated string");
        int n = (int)(t - s);
        memset(ptr, ' ', (size_t)charlen);
        memcpy(ptr, s, (size_t)(n < charlen ? n : charlen));
      } else if (type == 'l') {
        if (*e == '.') e++;
        *(int *)ptr = (*e == 't' || *e == 'T');
      } else {
        char tok[128];
        size_t k = 0;
        while (*e && *e != ',' && !isspace((unsigned char)*e) && k < sizeof tok - 1) {
          tok[k++] = (*e == 'd' || *e == 'D') ? 'e' : *e;
          e++;
        }
        tok[k] = 0;
        char *end;
        if (type == 'i') { long x = strtol(tok, &end, 10); if (end == tok || *end) fail("namelist: bad integer"); *(int *)ptr = (int)x; }
        else if (type == 'f') { float x = strtof(tok, &end); if (end == tok || *end) fail("namelist: bad real"); *(float *)ptr = x; }
        else { double x = strtod(tok, &end); if (end == tok || *end) fail("namelist: bad real"); *(double *)ptr = x; }
      }
      return;
    }
  }
  /* an object that is absent from the group keeps its value (Fortran semantics) */
}

void rt_nml_end(void) {
  if (!nml_gf) return;
  G.st_read(dtblk);
  G.st_read_done(dtblk);
  nml_gf = 0;
  if (r_ios != 0) fail("namelist read: libgfortran reports an error (iostat /= 0)");
}

/* ------------------------------------------------------------------ list-directed write (captured) */
static char wline[4096];
static size_t wlen;
static int w_is_perr;
static double w_last_real;
static int w_has_real;

static void wappend(const char *s) {
  size_t n = strlen(s);
  if (wlen + n + 1 < sizeof wline) { memcpy(wline + wlen, s, n); wlen += n; wline[wlen] = 0; }
}

static void capture_begin(void) { wlen = 0; wline[0] = 0; w_is_perr = 0; w_has_real = 0; w_is_step = 0; }
static int w_capture = 1;   /* writes to unit * are captured (log, p errors, step marks) */

void rt_write_begin(int unit) {
  capture_begin();
  w_capture = unit < 0 || unit == 6;
  w_gf = 0;
  if (gf_active) {
    gf_common(dtblk, sizeof dtblk, DT_LIST_FORMAT, unit < 0 ? 6 : unit);
    G.st_write(dtblk);
    w_gf = 1;
  } else if (!w_capture) {
    fail("write to a file needs the libgfortran backend (ref_use_libgfortran)");
  }
}

void rt_write_begin_fmt(int unit, const char *fmt, int fmtlen) {
  capture_begin();
  w_capture = unit < 0 || unit == 6;
  if (!gf_active) fail("formatted write needs the libgfortran backend (ref_use_libgfortran)");
  gf_common(dtblk, sizeof dtblk, DT_HAS_FORMAT, unit < 0 ? 6 : unit);
  *(const char **)(dtblk + 80) = fmt;
  *(size_t *)(dtblk + 88) = (size_t)fmtlen;
  G.st_write(dtblk);
  w_gf = 1;
}

void rt_write_begin_internal(char *buf, int buflen, const char *fmt, int fmtlen) {
  capture_begin();
  w_capture = 0;
  if (!gf_active) fail("internal write needs the libgfortran backend (ref_use_libgfortran)");
  gf_common(dtblk, sizeof dtblk, DT_HAS_FORMAT | DT_HAS_INTERNAL_UNIT, -1);
  *(const char **)(dtblk + 80) = fmt;
  *(size_t *)(dtblk + 88) = (size_t)fmtlen;
  *(char **)(dtblk + 112) = buf;
  *(size_t *)(dtblk + 120) = (size_t)buflen;
  G.st_write(dtblk);
  w_gf = 1;
}
void rt_write_str(const char *s) {
  if (w_gf) G.char_w(dtblk, (void *)s, strlen(s));
  if (!w_capture) return;
  if (strstr(s, "p error")) w_is_perr = 1;
  if (strstr(s, "time_steps=")) w_is_step = 1;
  wappend(" ");
  wappend(s);
}
void rt_write_chars(const char *p, int len, int trim) {
  char b[512];
  if (len > 511) len = 511;
  memcpy(b, p, (size_t)len);
  if (trim) while (len > 0 && b[len - 1] == ' ') len--;
  b[len] = 0;
  if (w_gf) G.char_w(dtblk, (void *)b, (size_t)len);
  if (!w_capture) return;
  wappend(" ");
  wappend(b);
}
void rt_write_int(int v) {
  if (w_gf) G.int_w(dtblk, &v, 4);
  if (!w_capture) return;
  char b[32];
  snprintf(b, sizeof b, " %11d", v);
  wappend(b);
}
void rt_write_real(double v) {
  if (w_gf) G.real_w(dtblk, &v, 8);
  if (!w_capture) return;
  char b[48];
  snprintf(b, sizeof b, " %.17g", v);
  wappend(b);
  w_last_real = v;
  w_has_real = 1;
}
void rt_write_logical(int v) {
  if (w_gf) G.logical_w(dtblk, &v, 4);
  if (w_capture) wappend(v ? " T" : " F");
}
void rt_write_end(void) {
  if (w_gf) { G.st_write_done(dtblk); w_gf = 0; }
  if (!w_capture) return;
  if (w_is_step && nstep_t < MAXSTEPT) step_t[nstep_t++] = now();
  if (w_is_step && step_limit > 0 && ++steps_seen > step_limit) {
    stopped_by_limit = 1;
    longjmp(stop_env, 1);
  }
  if (w_is_perr && w_has_real) {
    if (nperr == perrcap) { perrcap = perrcap ? 2 * perrcap : 256; perr = realloc(perr, sizeof(double) * (size_t)perrcap); }
    perr[nperr++] = w_last_real;
  }
  if (loglen + wlen + 2 > logcap) { logcap = 2 * (logcap + wlen + 2); logbuf = realloc(logbuf, logcap); }
  memcpy(logbuf + loglen, wline, wlen);
  loglen += wlen;
  logbuf[loglen++] = '\n';
  logbuf[loglen] = 0;
  if (verbose) puts(wline);
}

/* ------------------------------------------------------------------ stubs */
#define MAXSTUB 32
static struct { char name[64]; int count; } stubs[MAXSTUB];
static int nstubs;

void rt_stub(const char *name) {
  for (int i = 0; i < nstubs; i++) if (!strcmp(stubs[i].name, name)) { stubs[i].count++; return; }
  if (nstubs < MAXSTUB) { snprintf(stubs[nstubs].name, sizeof stubs[nstubs].name, "%s", name); stubs[nstubs++].count = 1; }
}

void rt_stop(void) { fail("stop"); }

/* ------------------------------------------------------------------ harness exports */
#ifdef _OPENMP
#include <omp.h>
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int ref_max_threads(void) { return omp_get_max_threads(); }
#else
void ref_set_threads(int n) { (void)n; }
int ref_max_threads(void) { return 1; }
#endif

int ref_run(const char *workdir) {
  char cwd[4096];
  if (!getcwd(cwd, sizeof cwd)) return 2;
  if (workdir && chdir(workdir) != 0) { snprintf(errmsg, sizeof errmsg, "chdir(%s) failed", workdir); return 2; }
  nperr = 0;
  loglen = 0;
  if (logbuf) logbuf[0] = 0;
  nstubs = 0;
  errmsg[0] = 0;
  steps_seen = 0;
  stopped_by_limit = 0;
  nstep_t = 0;
  static int ran = 0;
  if (ran) rt_reset_statics();   /* zero-initialised static storage, as at program start (first run: fresh BSS) */
  ran = 1;
  int rc = 0;
  if (gf_active) gf_open(6, "stdout.log");     /* unit *: the program's log, formatted by libgfortran */
  if (setjmp(stop_env) == 0) f_MAIN(); else rc = stopped_by_limit ? 0 : 1;
  if (gf_active) { w_gf = 0; for (int u = 0; u < 100; u++) if (gf_unit[u]) gf_close(u); }
  end_t = step_limit > 0 && nstep_t > step_limit ? step_t[step_limit] : now();
  for (int u = 0; u < 100; u++) rt_close(u);
  if (chdir(cwd) != 0) return 2;
  return rc;
}

/* porosity CSV in the format of template/data/.porosity:1-3 (`m,n,l` then `i, j, k, value`, i fastest); %.17g
 * round-trips every double.  eps is [l][n][m].  Only here because a Python writer needs 15 s for 4M records. */
int ref_write_csv(const char *path, const double *eps, int m, int n, int l) {
  FILE *f = fopen(path, "w");
  if (!f) return 1;
  static char buf[1 << 20];
  setvbuf(f, buf, _IOFBF, sizeof buf);
  fprintf(f, "%d,%d,%d\n", m, n, l);
  for (int k = 1; k <= l; k++)
    for (int j = 1; j <= n; j++)
      for (int i = 1; i <= m; i++)
        fprintf(f, "%d, %d, %d, %.17g\n", i, j, k, eps[(size_t)(i - 1) + (size_t)m * ((size_t)(j - 1) + (size_t)n * (size_t)(k - 1))]);
  return fclose(f) != 0;
}

const rt_var *ref_lookup(const char *name) {
  for (const rt_var *v = rt_registry; v->name; v++) if (!strcmp(v->name, name)) return v;
  return 0;
}
int ref_perr_count(void) { return nperr; }
double ref_perr(int i) { return (i >= 0 && i < nperr) ? perr[i] : 0.0; }
const char *ref_log(void) { return logbuf ? logbuf : ""; }
const char *ref_error(void) { return errmsg; }
void ref_set_verbose(int v) { verbose = v; }
void ref_set_step_limit(int n) { step_limit = n; }
int ref_step_count(void) { return nstep_t; }
double ref_step_time(int i) { return (i >= 0 && i < nstep_t) ? step_t[i] : 0.0; }
double ref_end_time(void) { return end_t; }
int ref_stub_count(const char *name) {
  for (int i = 0; i < nstubs; i++) if (!strcmp(stubs[i].name, name)) return stubs[i].count;
  return 0;
}
