// pixelflow_driver.cpp -- C++ twin of the reference's Fortran `program main` drivers.
//
// The reference's five programs (src/omp_parallel/ibm_*_omp_cpu.f90: program main) read
// config/controlDict.txt and a porosity CSV, run the time loop and write logs / VTK snapshots.  The
// Fortran drivers keep doing that and call libpixelflow_gpu.so through iso_c_binding
// (pixelflow_b200/fortran/pixelflow_gpu_mod.f90); this file is the same driver in C++ for machines
// without a Fortran compiler (this image has none).  It owns NO numerics: every field value comes
// from the C ABI (include/pixelflow_gpu.h).
//
// Mirrors, by reference line:
//   read_settings + echo            lib/global.f90:28-92
//   grid_conditions*                lib/grid.f90:6-110 (2D), :116-248 (wall), :250-382 (y/z periodic)
//   output_grid_*                   lib/output.f90:42-61, :591-613
//   output_paraview_temp_*          lib/output.f90:421-537, :968-1088  (ASCII legacy VTK, f16.4)
//   time loop + log lines           ibm_3d_uniform_omp_cpu.f90:33-143
//
// Usage: run from a project directory (config/controlDict.txt, data/*.csv), no arguments, under one
// of the reference's executable names (scripts/build/buildAll.sh:20-24, README.md:176-179):
//   ibm2_uniform_omp | ibm2_omp, ibm2_drag_omp, ibm2_backstep_omp, ibm3_uniform_omp | ibm3_omp,
//   ibm3_air_condition_omp           (symlinks to this binary; or `pixelflow_driver --case NAME`)
// Several GPUs: `--gpus N` or the environment variable PIXELFLOW_GPUS=N (the analogue of OMP_NUM_THREADS in the
// reference's config/omp_config.conf) runs the 3D cases z-slab decomposed over N GPUs, one forked process per GPU
// (pf_ranks_launch; rank 0 = this process keeps the log, every rank formats and writes its own planes of a snapshot).
// Extra, optional flags: --csv PATH (override csv_file), --steps N (override istep_max),
// --no-output (skip VTK files), --project DIR (chdir first), --cache (keep a binary copy <csv>.pfbin of the parsed CSV); --echo-settings and --format-selftest need no GPU
// (they exercise the namelist reader and the list-directed writer against libgfortran in the CPU tests).
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "pixelflow_gpu.h"

namespace {

// ------------------------------------------------------------------------------------------------
// gfortran list-directed output (write(*,*)), as observed from libgfortran.so.5 itself
// (tests/test_gfortran_io.py): every item is preceded by one blank -- the record's leading blank for the first --
// except a character item that follows a character item; integer(4) = I11; real(8) = a 25-column field with 17
// significant digits: F form right-justified in 20 columns + 5 blanks when the rounded value is in [0.1, 1e17)
// or zero, else d.dddddddddddddddddE+ddd right-justified; Infinity / NaN right-justified.
// f_real / f_int return the item WITH its separating blank (26 / 12 characters).
// ------------------------------------------------------------------------------------------------
std::string f_real(double x) {
  char b[64];
  if (std::isnan(x)) { snprintf(b, sizeof b, " %25s", "NaN"); return b; }
  if (std::isinf(x)) { snprintf(b, sizeof b, " %25s", x < 0 ? "-Infinity" : "Infinity"); return b; }
  char t[64];
  snprintf(t, sizeof t, "%.16E", x);                    // [-]d.ddddddddddddddddE[+-]XX, correctly rounded
  std::string m(t);
  const size_t epos = m.find('E');
  const int ex = atoi(m.c_str() + epos + 1);
  const bool neg = m[0] == '-';
  std::string digits;                                   // the 17 significant digits
  for (size_t q = neg ? 1 : 0; q < epos; ++q) if (m[q] != '.') digits += m[q];
  const bool zero = digits.find_first_not_of('0') == std::string::npos;
  std::string body;
  if (zero || (ex >= -1 && ex <= 16)) {
    if (zero) body = "0.0000000000000000";
    else if (ex == -1) body = "0." + digits;
    else body = digits.substr(0, (size_t)ex + 1) + "." + digits.substr((size_t)ex + 1);
    if (neg) body = "-" + body;
    snprintf(b, sizeof b, " %20s     ", body.c_str());
  } else {
    char u[64];
    snprintf(u, sizeof u, "%s%c.%sE%c%03d", neg ? "-" : "", digits[0], digits.substr(1).c_str(), ex < 0 ? '-' : '+', abs(ex));
    snprintf(b, sizeof b, " %25s", u);
  }
  return b;
}
std::string f_int(long long v) {
  char b[32];
  snprintf(b, sizeof b, " %11lld", v);
  return b;
}
// --format-selftest: the records the logs and etc/*.dat files are made of, for a fixed battery of values
// (compared with the runtime's own output on CPU: tests/test_gfortran_io.py)
int format_selftest() {
  const double vals[] = {0.5, 1.0, 0.0, -0.0, 123456.789, 1e16, 9.9999999999999999e16, 1e17, 0.1, 0.099999, -2.5, 1e-3,
                         -1e-300, 1e300, 3.0e-5, 12.0, 100.0, 0.25, 1.0 / 3.0, 20.0 / 3.0, 1e15 + 0.5, 9.9999999999999995,
                         0.99999999999999999, 99999999999999990.0, 5e-324, 1.7976931348623157e308, 2.0e-4, 5.0e-5,
                         0.063 / 63.0, 1.7, 1300.0, 3.0e-2};
  for (double v : vals) printf(" # xnue =%s\n", f_real(v).c_str());
  const int ints[] = {0, 1, -1, 100, 5000, 2147483647, -2147483647, 64};
  for (int v : ints) printf(" # SOR max iteration steps =%s\n", f_int(v).c_str());
  printf(" --- time_steps= %s  --  time = %s\n", f_int(7).c_str(), f_real(7 * 5.0e-5).c_str());
  printf(" SOR iteration no.%s -- p error:%s\n", f_int(100).c_str(), f_real(1.2345678901234567e-3).c_str());
  printf(" # m, n, l =%s%s%s\n", f_int(64).c_str(), f_int(64).c_str(), f_int(64).c_str());
  printf(" # dx, dy, dz =%s%s%s\n", f_real(1e-3).c_str(), f_real(0.5).c_str(), f_real(12.5).c_str());
  printf(" Fp =%s%s\n", f_real(-1.5e-3).c_str(), f_real(2.25).c_str());
  printf(" Cd =%s Cl =%s\n", f_real(1.25).c_str(), f_real(-3.5e-7).c_str());
  printf("%s%s%s\n", f_real(-0.315).c_str(), f_real(0.0).c_str(), f_real(0.315).c_str());
  printf(" # istep_max= %s    istep_out= %s\n", f_int(2000).c_str(), f_int(100).c_str());
  return 0;
}
void now_time() {  // lib/utils.f90:7-16
  time_t t = time(nullptr);
  struct tm tmv;
  localtime_r(&t, &tmv);
  char b[32];
  strftime(b, sizeof b, "%Y-%m-%d %H:%M:%S", &tmv);
  printf(" # --- TIME: %s\n", b);
}

// ------------------------------------------------------------------------------------------------
// controlDict.txt : seven namelist groups in fixed order (lib/global.f90:47-62)
// ------------------------------------------------------------------------------------------------
struct Settings {
  double xnue = 0, xlambda = 0, density = 0, width = 0, height = 0, depth = 0, time = 0;
  double inlet_velocity = 0, outlet_pressure = 0, AoA = 0;
  int istep_max = 0, istep_out = 0;
  double thickness = 0, threshold = 0, radius = 0, center_x = 0, center_y = 0, center_z = 0;
  bool nonslip = false;
  std::string output_folder, csv_file;
  int iter_max = 0;
  double relux_factor = 0;
};

std::string lower(std::string s) {
  std::transform(s.begin(), s.end(), s.begin(), ::tolower);
  return s;
}

double fortran_double(std::string v) {
  for (char &c : v)
    if (c == 'd' || c == 'D') c = 'e';
  return strtod(v.c_str(), nullptr);
}

Settings read_settings(const std::string &path) {
  std::ifstream in(path);
  if (!in) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(1); }
  static const char *order[] = {"physical", "file_control", "grid_control", "porosity_control",
                                "calculation_method", "directory_control", "solver_control"};
  std::map<std::string, std::string> kv;
  std::string line, group;
  int next_group = 0;
  while (std::getline(in, line)) {
    std::string t;
    char q = 0;
    for (char c : line) {  // strip ! comments outside quotes
      if (q) { t += c; if (c == q) q = 0; }
      else if (c == '"' || c == '\'') { q = c; t += c; }
      else if (c == '!') break;
      else t += c;
    }
    size_t a = t.find_first_not_of(" \t\r"), b = t.find_last_not_of(" \t\r");
    if (a == std::string::npos) continue;
    t = t.substr(a, b - a + 1);
    if (t[0] == '&') {
      group = lower(t.substr(1));
      size_t sp = group.find_first_of(" \t");
      if (sp != std::string::npos) group = group.substr(0, sp);
      // the reference reads the groups in order from one unit: a group is only found at or after
      // the current position
      int gi = -1;
      for (int i = 0; i < 7; ++i) if (group == order[i]) gi = i;
      if (gi >= 0) {
        if (gi < next_group) { fprintf(stderr, "namelist &%s out of order\n", group.c_str()); exit(1); }
        next_group = gi + 1;
      }
      continue;
    }
    if (t[0] == '/') { group.clear(); continue; }
    if (group.empty()) continue;
    std::stringstream ss(t);
    std::string item;
    while (std::getline(ss, item, ',')) {
      size_t eq = item.find('=');
      if (eq == std::string::npos) continue;
      std::string k = item.substr(0, eq), v = item.substr(eq + 1);
      auto trim = [](std::string s) {
        size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r/");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
      };
      kv[lower(trim(k))] = trim(v);
    }
  }
  Settings s;
  auto D = [&](const char *k, double &dst) { if (kv.count(k)) dst = fortran_double(kv[k]); };
  auto I = [&](const char *k, int &dst) { if (kv.count(k)) dst = atoi(kv[k].c_str()); };
  auto S = [&](const char *k, std::string &dst) {
    if (!kv.count(k)) return;
    std::string v = kv[k];
    if (v.size() >= 2 && (v[0] == '"' || v[0] == '\'')) v = v.substr(1, v.size() - 2);
    dst = v.substr(0, 50);
  };
  D("xnue", s.xnue); D("xlambda", s.xlambda); D("density", s.density); D("width", s.width);
  D("height", s.height); D("depth", s.depth); D("time", s.time); D("inlet_velocity", s.inlet_velocity);
  D("outlet_pressure", s.outlet_pressure); D("aoa", s.AoA);
  I("istep_out", s.istep_out); I("istep_max", s.istep_max);
  D("thickness", s.thickness); D("threshold", s.threshold); D("radius", s.radius);
  D("center_x", s.center_x); D("center_y", s.center_y); D("center_z", s.center_z);
  if (kv.count("nonslip")) {
    std::string v = lower(kv["nonslip"]);
    s.nonslip = v.find('t') != std::string::npos && v.find('t') <= 1;
  }
  S("output_folder", s.output_folder); S("csv_file", s.csv_file);
  I("iter_max", s.iter_max); D("relux_factor", s.relux_factor);
  // echo (lib/global.f90:66-90)
  printf(" #\n # --- Physical conditions\n");
  printf(" # xnue =%s\n # xlambda =%s\n # density =%s\n # width =%s\n # height =%s\n # depth =%s\n # time =%s\n",
         f_real(s.xnue).c_str(), f_real(s.xlambda).c_str(), f_real(s.density).c_str(), f_real(s.width).c_str(),
         f_real(s.height).c_str(), f_real(s.depth).c_str(), f_real(s.time).c_str());
  printf(" # inlet_velocity =%s\n # outlet_pressure =%s\n # Angle of inlet_velocity (AoA) =%s\n",
         f_real(s.inlet_velocity).c_str(), f_real(s.outlet_pressure).c_str(), f_real(s.AoA).c_str());
  printf(" #\n # --- Porosity information\n # thickness =%s\n # threshold =%s\n # radius =%s\n",
         f_real(s.thickness).c_str(), f_real(s.threshold).c_str(), f_real(s.radius).c_str());
  printf(" #\n # --- Directory information\n # output_folder =%-50s\n # input_porosity_file =%-50s\n",
         s.output_folder.c_str(), s.csv_file.c_str());
  printf(" #\n # --- Solver information\n # SOR max iteration steps =%s\n # SOR reluxation factor =%s\n",
         f_int(s.iter_max).c_str(), f_real(s.relux_factor).c_str());
  return s;
}

// ------------------------------------------------------------------------------------------------
// porosity CSV: "m,n,l" then m*n*l records "ix, iy, iz, value" (lib/grid.f90:281-294).  The reader
// honours the explicit indices; values are clamped to `threshold` on read (:289).
// ------------------------------------------------------------------------------------------------
struct Grid {
  int m = 0, n = 0, l = 1;
  bool d3 = false;
  double dx = 0, dy = 0, dz = 1, dt = 0;
  std::vector<double> xp, yp, zp, eps;  // eps: (l+2 | 1) x (n+2) x (m+2)
  size_t LX() const { return m + 2; }
  size_t LY() const { return n + 2; }
  size_t idx(int i, int j, int k) const { return i + LX() * (j + LY() * (size_t)k); }
};

// Binary cache of a parsed CSV, OPT-IN (`--cache`): `<csv>.pfbin` = header {magic, m, n, l, d3, threshold, csv size,
// csv mtime to the nanosecond} + the clamped interior-with-halo-slots array as raw fp64.  It is used only if every
// header field still matches the CSV next to it AND its dimensions equal the CSV's own header line, so the CSV stays
// the contract (template/data/.porosity); without --cache nothing is read from or written into the data directory.
struct CacheHeader {
  char magic[8];
  int m, n, l, d3;
  double threshold;
  long long csv_size, csv_mtime_s, csv_mtime_ns;
};

// m, n[, l] of the CSV's first line (lib/grid.f90:283 / :38); false if it cannot be read
bool csv_dims(const std::string &csv, bool d3, int &m, int &n, int &l) {
  FILE *f = fopen(csv.c_str(), "rb");
  if (!f) return false;
  char line[256];
  const bool got = fgets(line, sizeof line, f) != nullptr;
  fclose(f);
  if (!got) return false;
  for (char *c = line; *c; ++c)
    if (*c == ',') *c = ' ';
  m = n = 0; l = 1;
  const int k = sscanf(line, "%d %d %d", &m, &n, &l);
  if (!d3) l = 1;
  return k >= (d3 ? 3 : 2) && m > 0 && n > 0 && l > 0;
}

bool read_cache(const std::string &csv, bool d3, double threshold, Grid &g) {
  struct stat sc;
  int m = 0, n = 0, l = 1;
  if (stat(csv.c_str(), &sc) != 0 || !csv_dims(csv, d3, m, n, l)) return false;
  FILE *f = fopen((csv + ".pfbin").c_str(), "rb");
  if (!f) return false;
  CacheHeader h;
  bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, "PFBIN02", 8) == 0 && h.d3 == (int)d3 &&
            h.threshold == threshold && h.csv_size == (long long)sc.st_size &&
            h.csv_mtime_s == (long long)sc.st_mtim.tv_sec && h.csv_mtime_ns == (long long)sc.st_mtim.tv_nsec &&
            h.m == m && h.n == n && (!d3 || h.l == l);
  if (ok) {
    g.d3 = d3; g.m = h.m; g.n = h.n; g.l = d3 ? h.l : 1;
    const size_t ne = (size_t)(d3 ? g.l + 2 : 1) * g.LX() * g.LY();
    g.eps.resize(ne);
    ok = fread(g.eps.data(), sizeof(double), ne, f) == ne;
  }
  fclose(f);
  return ok;
}

void write_cache(const std::string &csv, double threshold, const Grid &g) {
  struct stat sc;
  if (stat(csv.c_str(), &sc) != 0) return;
  FILE *f = fopen((csv + ".pfbin").c_str(), "wb");
  if (!f) return;   // read-only data directory: just run without a cache
  CacheHeader h;
  memset(&h, 0, sizeof h);
  memcpy(h.magic, "PFBIN02", 8);
  h.m = g.m; h.n = g.n; h.l = g.l; h.d3 = g.d3; h.threshold = threshold;
  h.csv_size = sc.st_size; h.csv_mtime_s = sc.st_mtim.tv_sec; h.csv_mtime_ns = sc.st_mtim.tv_nsec;
  if (fwrite(&h, sizeof h, 1, f) != 1 || fwrite(g.eps.data(), sizeof(double), g.eps.size(), f) != g.eps.size()) {
    fclose(f);
    remove((csv + ".pfbin").c_str());
    return;
  }
  fclose(f);
}

// cache_read / cache_write: --cache (every rank may read the cache, rank 0 alone writes it); device: the GPU that
// parses the records (-1 = the current one)
Grid read_porosity(const std::string &path, bool d3, double threshold, bool cache_read, bool cache_write, int device) {
  if (cache_read) {
    Grid cached;
    if (read_cache(path, d3, threshold, cached)) return cached;
  }
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open porosity file %s\n", path.c_str()); exit(1); }
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<char> buf(sz + 1);
  if (fread(buf.data(), 1, sz, f) != (size_t)sz) { fprintf(stderr, "short read\n"); exit(1); }
  buf[sz] = 0;
  fclose(f);
  char *p = buf.data();
  auto skip = [&]() { while (*p && (*p == ',' || *p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) ++p; };
  Grid g;
  g.d3 = d3;
  skip(); g.m = (int)strtol(p, &p, 10);
  skip(); g.n = (int)strtol(p, &p, 10);
  // the 2D reader takes only m,n from the header (lib/grid.f90:38); skip the rest of the line
  if (d3) { skip(); g.l = (int)strtol(p, &p, 10); }
  while (*p && *p != '\n') ++p;
  if (*p == '\n') ++p;
  const size_t planes = d3 ? g.l + 2 : 1;
  g.eps.assign(planes * g.LX() * g.LY(), 0.0);
  // the records (`read(52,*) x, y, z, poro_val`, lib/grid.f90:281-294) are parsed on the GPU
  const long long nrec = (long long)g.m * g.n * (d3 ? g.l : 1);
  long long got = 0;
  if (pf_parse_porosity_csv(p, (size_t)(buf.data() + sz - p), g.m, g.n, d3 ? g.l : 0, threshold, g.eps.data(), &got, device)) {
    fprintf(stderr, " %s\n", pf_last_error(nullptr)); exit(1);
  }
  if (got < nrec) { fprintf(stderr, "porosity file %s: %lld records, %lld expected\n", path.c_str(), got, nrec); exit(1); }
  if (cache_write) write_cache(path, threshold, g);
  return g;
}

// halo rules of lib/grid.f90 (input preparation)
void porosity_halo(Grid &g, int scase) {
  const int m = g.m, n = g.n, l = g.l;
  auto &e = g.eps;
  if (!g.d3) {                                   // :92-106
    for (int j = 1; j <= n + 1; ++j) { e[g.idx(0, j, 0)] = e[g.idx(1, j, 0)]; e[g.idx(m + 1, j, 0)] = e[g.idx(m, j, 0)]; }
    for (int i = 0; i <= m + 1; ++i) { e[g.idx(i, 0, 0)] = e[g.idx(i, n, 0)]; e[g.idx(i, n + 1, 0)] = e[g.idx(i, 1, 0)]; }
  } else if (scase == PF_IBM3_AIRCOND) {         // :215-243
    for (int j = 0; j <= n + 1; ++j) for (int k = 0; k <= l + 1; ++k) { e[g.idx(0, j, k)] = e[g.idx(1, j, k)]; e[g.idx(m + 1, j, k)] = e[g.idx(m, j, k)]; }
    for (int i = 0; i <= m + 1; ++i) for (int k = 0; k <= l + 1; ++k) { e[g.idx(i, 0, k)] = e[g.idx(i, 1, k)]; e[g.idx(i, n + 1, k)] = e[g.idx(i, n, k)]; }
    for (int i = 0; i <= m + 1; ++i) for (int j = 0; j <= n + 1; ++j) { e[g.idx(i, j, 0)] = e[g.idx(i, j, 1)]; e[g.idx(i, j, l + 1)] = e[g.idx(i, j, l)]; }
  } else {                                       // :349-378
    for (int j = 1; j <= n + 1; ++j) for (int k = 1; k <= l + 1; ++k) { e[g.idx(0, j, k)] = e[g.idx(1, j, k)]; e[g.idx(m + 1, j, k)] = e[g.idx(m, j, k)]; }
    for (int i = 0; i <= m + 1; ++i) for (int k = 0; k <= l + 1; ++k) { e[g.idx(i, 0, k)] = e[g.idx(i, n, k)]; e[g.idx(i, n + 1, k)] = e[g.idx(i, 1, k)]; }
    for (int i = 0; i <= m + 1; ++i) for (int j = 0; j <= n + 1; ++j) { e[g.idx(i, j, 0)] = e[g.idx(i, j, l)]; e[g.idx(i, j, l + 1)] = e[g.idx(i, j, 1)]; }
  }
}

void grid_conditions(Grid &g, const Settings &s) {  // lib/grid.f90:297-347
  g.dx = s.width / (double)(g.m - 1);
  g.dy = s.height / (double)(g.n - 1);
  if (g.d3) g.dz = s.depth / (double)(g.l - 1);
  g.dt = s.time / (double)s.istep_max;
  const double cfl = s.inlet_velocity * g.dt / g.dx, pec = s.inlet_velocity * g.dx / s.xnue;
  const double dif = s.xnue * g.dt / g.dy / g.dy, re = s.inlet_velocity * s.radius * 2.0 / s.xnue;
  printf("\n # --- Grid conditions\n");
  if (g.d3) {
    printf(" # m, n, l =%s%s%s\n # istep_max =%s\n # dx, dy, dz =%s%s%s\n", f_int(g.m).c_str(), f_int(g.n).c_str(),
           f_int(g.l).c_str(), f_int(s.istep_max).c_str(), f_real(g.dx).c_str(), f_real(g.dy).c_str(), f_real(g.dz).c_str());
  } else {
    printf(" # m, n =%s%s\n # dx, dy =%s%s\n", f_int(g.m).c_str(), f_int(g.n).c_str(), f_real(g.dx).c_str(),
           f_real(g.dy).c_str());
  }
  printf(" # dt =%s\n # cfl_no =%s\n # pecret_no =%s\n # diffusion_factor =%s\n # reynolds_no =%s\n",
         f_real(g.dt).c_str(), f_real(cfl).c_str(), f_real(pec).c_str(), f_real(dif).c_str(), f_real(re).c_str());
  if (g.d3) printf(" # thickness =%s\n # threshold =%s\n", f_real(s.thickness).c_str(), f_real(s.threshold).c_str());
  printf("\n");
  g.xp.resize(g.m + 2); g.yp.resize(g.n + 2); g.zp.resize(g.l + 2);
  for (int i = 0; i <= g.m + 1; ++i) g.xp[i] = g.dx * (double)(i - 1) - s.width * s.center_x;
  for (int j = 0; j <= g.n + 1; ++j) g.yp[j] = g.dy * (double)(j - 1) - s.height * s.center_y;
  for (int k = 0; k <= g.l + 1; ++k) g.zp[k] = g.dz * (double)(k - 1) - s.depth * s.center_z;
}

// ------------------------------------------------------------------------------------------------
// outputs
// ------------------------------------------------------------------------------------------------
void output_grid(const Grid &g) {  // lib/output.f90:42-61, :591-613
  FILE *f = fopen("etc/grid.dat", "w");
  if (!f) return;
  if (g.d3) fprintf(f, " m, n, l =%s%s%s\n", f_int(g.m).c_str(), f_int(g.n).c_str(), f_int(g.l).c_str());
  else      fprintf(f, " m, n =%s%s\n", f_int(g.m).c_str(), f_int(g.n).c_str());
  fprintf(f, " grid points =\n");
  auto row = [&](const std::vector<double> &a, int cnt) {
    for (int i = 1; i <= cnt; ++i) fprintf(f, "%s", f_real(a[i]).c_str());
    fprintf(f, "\n");
  };
  row(g.xp, g.m); row(g.yp, g.n);
  if (g.d3) row(g.zp, g.l);
  fclose(f);
}

// ASCII legacy VTK, f16.4 (lib/output.f90:421-537 2D, :968-1088 3D).  The header lines are written here; the
// bodies -- 210 bytes per cell in 3D -- are formatted on the GPU from the device-resident fields
// (pf_vtk_section) in chunks of planes and copied straight to the file.
double g_output_seconds = 0.0;
void die(pf_solver *s, const char *what);
// `final_file`: output_paraview_3d (lib/output.f90:795-912) orders its scalars pressure, VelocityDivergent, porosity;
// the per-step snapshots and both 2D routines write porosity, pressure, VelocityDivergent.
// `s == nullptr` (--replay, no device): the header lines only, no bodies.
// Several ranks: the records have a fixed width, so every byte offset of the file is known in advance -- rank 0 writes
// the header lines, every rank formats the records of ITS planes on its GPU and writes them at their place (pwrite).
void output_paraview(pf_solver *s, const Grid &g, const std::string &fname, bool final_file = false) {
  const auto t0 = std::chrono::steady_clock::now();
  const int m = g.m, n = g.n, l = g.d3 ? g.l : 1;
  const long long np = (long long)m * n * l;
  const int rank = pf_ranks_rank();
  int k_first = 1, k_count = l;
  if (s && g.d3) pf_local_slab(s, &k_first, &k_count);
  // the file as a list of (header text, section) items; section < 0: header only
  struct Item { std::string head; int section; };
  std::vector<Item> items;
  char line[256];
  snprintf(line, sizeof line, "# vtk DataFile Version 3.0\n%s\nASCII \nDATASET STRUCTURED_GRID\nDIMENSIONS  %4d %4d %4d\n"
           "POINTS %9lld float\n", g.d3 ? "3D flow" : "2D flow", m, n, l, np);
  items.push_back({line, PF_VTK_POINTS});
  snprintf(line, sizeof line, "POINT_DATA %9lld\nVECTORS velocity float\n", np);
  items.push_back({line, PF_VTK_VELOCITY});
  items.push_back({"VECTORS velocityInFluid float\n", PF_VTK_VELOCITY_IN_FLUID});
  if (!g.d3) items.push_back({"VECTORS dimless_v float\n", PF_VTK_DIMLESS_V});                  // lib/output.f90:468-474
  const bool porosity_last = final_file && g.d3;
  if (!porosity_last) items.push_back({"SCALARS porosity float\nLOOKUP_TABLE default\n", PF_VTK_POROSITY});
  items.push_back({"SCALARS pressure float\nLOOKUP_TABLE default\n", PF_VTK_PRESSURE});
  items.push_back({"SCALARS VelocityDivergent float\nLOOKUP_TABLE default\n", PF_VTK_DIVERGENT});
  if (porosity_last) items.push_back({"SCALARS porosity float\nLOOKUP_TABLE default\n", PF_VTK_POROSITY});
  if (!g.d3) items.push_back({"SCALARS abs_dimless_v float\nLOOKUP_TABLE default\n", PF_VTK_ABS_DIMLESS_V});   // :518-526

  int fd = -1;
  if (rank == 0) {
    fd = open(fname.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0666);
    if (fd < 0) fprintf(stderr, "cannot write %s\n", fname.c_str());
  }
  if (pf_ranks_barrier()) die(s, "pf_ranks_barrier");   // the file exists (and is empty) before anybody else opens it
  if (rank != 0) fd = open(fname.c_str(), O_WRONLY);
  if (fd < 0) return;
  auto put = [&](const void *data, size_t bytes, long long off) {
    const char *c = static_cast<const char *>(data);
    while (bytes) {
      const ssize_t w = pwrite(fd, c, bytes, off);
      if (w <= 0) { fprintf(stderr, "write to %s failed\n", fname.c_str()); exit(1); }
      c += w; bytes -= (size_t)w; off += w;
    }
  };
  std::vector<char> buf;
  long long off = 0;
  for (const Item &it : items) {
    if (rank == 0) put(it.head.data(), it.head.size(), off);
    off += (long long)it.head.size();
    if (!s) continue;                                       // --replay: header lines only
    const long long per_plane = (long long)pf_vtk_section_bytes(s, it.section, 1);
    // this rank's planes, in chunks of at most ~256 MB of text
    const int chunk = g.d3 ? (int)std::max(1LL, std::min((long long)k_count, (256LL << 20) / std::max(per_plane, 1LL))) : 1;
    for (int c0 = 0; c0 < k_count; c0 += chunk) {
      const int nk = std::min(chunk, k_count - c0);
      const size_t bytes = pf_vtk_section_bytes(s, it.section, nk);
      if (buf.size() < bytes) buf.resize(bytes);
      if (pf_vtk_section(s, it.section, g.d3 ? c0 + 1 : 0, nk, g.xp.data(), g.yp.data(), g.d3 ? g.zp.data() : nullptr, buf.data()))
        die(s, "pf_vtk_section");
      put(buf.data(), bytes, off + (long long)(k_first - 1 + c0) * per_plane);
    }
    off += per_plane * l;
  }
  close(fd);
  if (pf_ranks_barrier()) die(s, "pf_ranks_barrier");   // complete before the run goes on (or the process ends)
  g_output_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// output_solution_post_2d / _3d (lib/output.f90:64-200 / :648-792): etc/solution_uvp.dat, list-directed, and
// etc/surface_profile.dat.  In 3D every k-plane is ONE record ((a(i,j,k), i=1,m), j=1,n); in 2D every row j is.
void output_solution(const Grid &g, const std::vector<double> &u, const std::vector<double> &v,
                     const std::vector<double> &w, const std::vector<double> &p) {
  const int l = g.d3 ? g.l : 1, K0 = g.d3 ? 1 : 0;
  const double small = 1.e-6, pmin = 0.25, pmax = 0.75;
  auto p_cnt = [&](size_t c) { return g.eps[c] > small ? p[c] : 0.0; };     // :673-677
  FILE *f = fopen("etc/solution_uvp.dat", "w");
  if (f) {
    auto block = [&](const char *title, auto &&value) {
      fprintf(f, " %s\n", title);
      for (int k = K0; k < K0 + l; ++k) {
        for (int j = 1; j <= g.n; ++j) {
          for (int i = 1; i <= g.m; ++i) fprintf(f, "%s", f_real(value(g.idx(i, j, k))).c_str());
          if (!g.d3) fprintf(f, "\n");
        }
        if (g.d3) fprintf(f, "\n");
      }
    };
    fprintf(f, " m, n%s =%s%s%s\n", g.d3 ? ", l" : "", f_int(g.m).c_str(), f_int(g.n).c_str(), g.d3 ? f_int(g.l).c_str() : "");
    block("velocity u_bulk ", [&](size_t c) { return u[c] * g.eps[c]; });
    block("velocity v_bulk ", [&](size_t c) { return v[c] * g.eps[c]; });
    if (g.d3) block("velocity w_bulk ", [&](size_t c) { return w[c] * g.eps[c]; });
    block("velocity u_inst ", [&](size_t c) { return u[c]; });
    block("velocity v_inst ", [&](size_t c) { return v[c]; });
    if (g.d3) block("velocity w_inst ", [&](size_t c) { return w[c]; });
    block("pressure p_fluid", p_cnt);
    block("pressure P_all", [&](size_t c) { return p[c]; });
    block("porosity", [&](size_t c) { return g.eps[c]; });
    fclose(f);
  }
  f = fopen("etc/surface_profile.dat", "w");                                 // :181-195 / :771-789
  if (f) {
    for (int k = K0; k < K0 + l; ++k)
      for (int j = 1; j <= g.n; ++j)
        for (int i = 1; i <= g.m; ++i) {
          const size_t c = g.idx(i, j, k);
          if (g.eps[c] < pmax && g.eps[c] > pmin) {
            fprintf(f, "%s%s", f_real(g.xp[i]).c_str(), f_real(g.yp[j]).c_str());
            // zp(i), sic (:783): beyond l+1 the reference reads the zero-initialised tail of its static array
            if (g.d3) fprintf(f, "%s", f_real(i <= g.l + 1 ? g.zp[i] : 0.0).c_str());
            fprintf(f, "%s%s\n", f_real(p_cnt(c)).c_str(), f_real(g.eps[c]).c_str());
          }
        }
    fclose(f);
  }
}

// output_divergent_2d / _3d (lib/output.f90:202-242 / :912-966): porosity-weighted divergence, etc/divergent.dat
void output_divergent(const Grid &g, const std::vector<double> &u, const std::vector<double> &v,
                      const std::vector<double> &w) {
  FILE *f = fopen("etc/divergent.dat", "w");
  if (!f) return;
  const int l = g.d3 ? g.l : 1, K0 = g.d3 ? 1 : 0;
  const size_t sx = 1, sy = g.LX(), sz = g.LX() * g.LY();
  const auto &e = g.eps;
  auto rows = [&](auto &&value) {
    for (int k = K0; k < K0 + l; ++k)
      for (int j = 1; j <= g.n; ++j) {
        for (int i = 1; i <= g.m; ++i) fprintf(f, "%s", f_real(value(g.idx(i, j, k))).c_str());
        fprintf(f, "\n");
      }
  };
  fprintf(f, "\n porosity\n");
  rows([&](size_t c) { return e[c]; });
  fprintf(f, "\n divergent velocity\n");
  rows([&](size_t c) {
    double d = ((e[c + sx] * u[c] + e[c] * u[c + sx]) / 2 - (e[c - sx] * u[c] + e[c] * u[c - sx]) / 2) / g.dx +
               ((e[c + sy] * v[c] + e[c] * v[c + sy]) / 2 - (e[c - sy] * v[c] + e[c] * v[c - sy]) / 2) / g.dy;
    if (g.d3) d = d + ((e[c + sz] * w[c] + e[c] * w[c + sz]) / 2 - (e[c - sz] * w[c] + e[c] * w[c - sz]) / 2) / g.dz;
    return d;
  });
  fprintf(f, "\n");
  fclose(f);
}

struct CaseName { const char *exe; int scase; };
const CaseName kNames[] = {
    {"ibm2_uniform_omp", PF_IBM2_UNIFORM}, {"ibm2_omp", PF_IBM2_UNIFORM}, {"ibm2", PF_IBM2_UNIFORM},
    {"ibm_2d_uniform_omp_cpu", PF_IBM2_UNIFORM}, {"ibm2_uniform", PF_IBM2_UNIFORM},
    {"ibm2_backstep_omp", PF_IBM2_BACKSTEP}, {"ibm_2d_backstep_omp_cpu", PF_IBM2_BACKSTEP}, {"ibm2_backstep", PF_IBM2_BACKSTEP},
    {"ibm2_drag_omp", PF_IBM2_DRAG}, {"ibm_2d_drag_omp_cpu", PF_IBM2_DRAG}, {"ibm2_drag", PF_IBM2_DRAG},
    {"ibm3_uniform_omp", PF_IBM3_UNIFORM}, {"ibm3_omp", PF_IBM3_UNIFORM}, {"ibm3", PF_IBM3_UNIFORM},
    {"ibm_3d_uniform_omp_cpu", PF_IBM3_UNIFORM}, {"ibm3_uniform", PF_IBM3_UNIFORM},
    {"ibm3_air_condition_omp", PF_IBM3_AIRCOND}, {"ibm_3d_air_condition_omp_cpu", PF_IBM3_AIRCOND},
    {"ibm3_air_condition", PF_IBM3_AIRCOND},
};

int case_from_name(const std::string &name) {
  for (const CaseName &c : kNames) if (name == c.exe) return c.scase;
  return -1;
}

void die(pf_solver *s, const char *what) {
  fprintf(stderr, " pixelflow_gpu error in %s (rank %d): %s\n", what, pf_ranks_rank(), pf_last_error(s));
  pf_ranks_finish(1);   // ranks > 0 leave here; rank 0 collects them
  exit(1);
}

}  // namespace

// --replay FILE: TEST HOOK, not a solver.  FILE is a recorded run (the per-step p errors, the force log values and
// the final u, v, [w,] p in this driver's array layout); the driver then prints its log and writes its host-side
// files (etc/*.dat, and the VTK files' header lines without their device-formatted bodies) from that record, with no
// device and no computation.  It exists so that the log and the file writers can be diffed byte for byte against
// the reference's own output routines on a machine without a GPU (tests/test_ref_output_files.py).
struct Replay {
  bool on = false;
  std::vector<double> perr, force, u, v, w, p;
};

bool load_replay(const std::string &path, size_t nelem, bool d3, Replay &r) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  char magic[8];
  int nsteps = 0, has_force = 0;
  bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "PFREPLAY", 8) == 0 && fread(&nsteps, 4, 1, f) == 1 &&
            fread(&has_force, 4, 1, f) == 1 && nsteps >= 0;
  auto rd = [&](std::vector<double> &a, size_t n) {
    a.resize(n);
    return n == 0 || fread(a.data(), sizeof(double), n, f) == n;
  };
  ok = ok && rd(r.perr, (size_t)nsteps) && rd(r.force, has_force ? 8 * (size_t)nsteps : 0) && rd(r.u, nelem) &&
       rd(r.v, nelem) && rd(r.w, d3 ? nelem : 0) && rd(r.p, nelem);
  fclose(f);
  r.on = ok;
  return ok;
}

// --ranks-selftest N FILE: the multi-process plumbing of pf_ranks.cu without a GPU (CPU test): N ranks are forked, meet
// at a barrier, each writes one fixed-width record at ITS offset of FILE (the way the VTK snapshots are written), meet
// again, and rank 0 collects the others.  With a third argument R, rank R fails: every rank must notice.
int ranks_selftest(int n, const char *file, int failing) {
  int rank = -1;
  if (pf_ranks_launch(n, &rank)) { fprintf(stderr, "pf_ranks_launch: %s\n", pf_last_error(nullptr)); return 1; }
  printf("rank %d of %d prints\n", rank, pf_ranks_count());   // only rank 0's line reaches stdout
  fflush(stdout);
  int fd = -1;
  if (rank == 0) fd = open(file, O_CREAT | O_TRUNC | O_WRONLY, 0666);
  if (rank == failing) return pf_ranks_finish(3);
  if (pf_ranks_barrier()) return pf_ranks_finish(4);
  if (rank != 0) fd = open(file, O_WRONLY);
  char rec[16];
  snprintf(rec, sizeof rec, "rank %3d ok\n", rank);   // 12 bytes
  const bool wrote = fd >= 0 && pwrite(fd, rec, 12, 12 * (off_t)rank) == 12;
  if (fd >= 0) close(fd);
  if (pf_ranks_barrier()) return pf_ranks_finish(5);
  return pf_ranks_finish(wrote ? 0 : 6);
}

int main(int argc, char **argv) {
  std::string exe = argv[0];
  size_t slash = exe.find_last_of('/');
  if (slash != std::string::npos) exe = exe.substr(slash + 1);
  int scase = case_from_name(exe);
  std::string csv_override, project, replay_path;
  int steps_override = -1, gpus = 0;
  bool no_output = false, echo_only = false, use_cache = false;
  for (int a = 1; a < argc; ++a) {
    std::string o = argv[a];
    if (o == "--case" && a + 1 < argc) scase = case_from_name(argv[++a]);
    else if (o == "--csv" && a + 1 < argc) csv_override = argv[++a];
    else if (o == "--steps" && a + 1 < argc) steps_override = atoi(argv[++a]);
    else if (o == "--gpus" && a + 1 < argc) gpus = atoi(argv[++a]);
    else if (o == "--format-selftest") return format_selftest();
    else if (o == "--ranks-selftest" && a + 2 < argc)
      return ranks_selftest(atoi(argv[a + 1]), argv[a + 2], a + 3 < argc ? atoi(argv[a + 3]) : -1);
    else if (o == "--project" && a + 1 < argc) project = argv[++a];
    else if (o == "--no-output") no_output = true;
    else if (o == "--cache") use_cache = true;   // keep / use <csv>.pfbin next to the porosity CSV
    else if (o == "--replay" && a + 1 < argc) replay_path = argv[++a];
    else if (o == "--echo-settings") echo_only = true;   // read config/controlDict.txt, print the header echo, stop
    else { fprintf(stderr, "unknown option %s\n", o.c_str()); return 2; }
  }
  if (echo_only) {
    if (!project.empty() && chdir(project.c_str()) != 0) { perror("chdir"); return 2; }
    read_settings("config/controlDict.txt");
    return 0;
  }
  if (scase < 0) {
    fprintf(stderr, "cannot tell the solver from the executable name '%s'; use --case ibm3_uniform_omp etc.\n", exe.c_str());
    return 2;
  }
  if (!project.empty() && chdir(project.c_str()) != 0) { perror("chdir"); return 2; }
  const bool d3 = scase >= PF_IBM3_UNIFORM;

  // Several GPUs (3D cases): one forked process per GPU from here on, before anything touches CUDA -- the CSV is
  // parsed on the GPU, so every rank reads the deck itself.  Rank 0 is this process and keeps the log; the others
  // write no file of their own except their planes of the VTK snapshots.  `--gpus 0` (the default) takes the count
  // from PIXELFLOW_GPUS.
  int rank = 0;
  if (replay_path.empty()) {
    if (!d3 && (gpus > 1 || (gpus == 0 && getenv("PIXELFLOW_GPUS") && atoi(getenv("PIXELFLOW_GPUS")) > 1))) {
      fprintf(stderr, " the 2D cases run on one GPU (there is no z to decompose)\n");
      return 2;
    }
    if (d3 && pf_ranks_launch(gpus, &rank)) { fprintf(stderr, " pf_ranks_launch: %s\n", pf_last_error(nullptr)); return 1; }
  }
  const int nranks = pf_ranks_count();

  now_time();
  Settings st = read_settings("config/controlDict.txt");
  if (!csv_override.empty()) st.csv_file = csv_override;
  // --steps shortens the loop only; dt stays time/istep_max as in the deck
  const int nloop = steps_override >= 0 ? std::max(steps_override, 1) : st.istep_max;
  if (rank == 0) {
    mkdir(st.output_folder.c_str(), 0777);
    mkdir("etc", 0777);
  }
  Grid g = read_porosity(st.csv_file, d3, st.threshold, use_cache, use_cache && rank == 0, nranks > 1 ? rank : -1);
  grid_conditions(g, st);
  porosity_halo(g, scase);
  if (rank == 0) output_grid(g);
  printf(" # istep_max= %s    istep_out= %s\n", f_int(st.istep_max).c_str(), f_int(st.istep_out).c_str());

  pf_config cfg;
  pf_config_init(&cfg);
  cfg.solver_case = scase;
  cfg.m = g.m; cfg.n = g.n; cfg.l = g.l;
  cfg.dx = g.dx; cfg.dy = g.dy; cfg.dz = g.dz; cfg.dt = g.dt;
  cfg.xnue = st.xnue; cfg.xlambda = st.xlambda; cfg.density = st.density; cfg.thickness = st.thickness;
  cfg.nonslip = st.nonslip ? 1 : 0;
  cfg.iter_max = st.iter_max;
  cfg.relux_factor = st.relux_factor;
  cfg.inlet_velocity = st.inlet_velocity; cfg.outlet_pressure = st.outlet_pressure; cfg.AoA = st.AoA;
  cfg.rank = rank;
  cfg.nranks = nranks;
  if (nranks > 1) {
    cfg.device = rank;                              // one GPU per rank
    cfg.nccl_unique_id = pf_ranks_unique_id();      // made by rank 0, awaited by the others
    if (!cfg.nccl_unique_id) { fprintf(stderr, " pf_ranks_unique_id: %s\n", pf_last_error(nullptr)); pf_ranks_finish(1); return 1; }
  }
  pf_solver *s = nullptr;
  // the fields live on the devices; only rank 0 ever holds them on the host (pf_gather), for the end-of-run files
  const size_t nelem = rank == 0 ? g.eps.size() : 0;
  std::vector<double> u(nelem, 0.0), v(nelem, 0.0), w(d3 ? nelem : 0, 0.0), p(nelem, 0.0);
  Replay replay;
  if (!replay_path.empty()) {
    if (!load_replay(replay_path, nelem, d3, replay)) { fprintf(stderr, " cannot read the replay record %s\n", replay_path.c_str()); return 2; }
  } else {
    if (pf_create(&s, &cfg)) { fprintf(stderr, " pf_create (rank %d): %s\n", rank, pf_last_error(nullptr)); pf_ranks_finish(1); return 1; }
    if (pf_set_porosity(s, g.eps.data())) die(s, "pf_set_porosity");
    // u = v = w = p = 0 (the reference's static arrays): the device arrays are created zeroed; rank 0 uploads its
    // host zeros all the same, so that one rank runs exactly the sequence of calls it always did
    if (rank == 0 && pf_upload(s, u.data(), v.data(), d3 ? w.data() : nullptr, p.data())) die(s, "pf_upload");
    if (pf_initial_conditions(s)) die(s, "pf_initial_conditions");   // initial_conditions + boundary (:68-71)
  }
  auto snapshot = [&](int istep) {   // the fields stay on the device: the snapshot text is produced there
    if (no_output) return;
    char name[512];
    snprintf(name, sizeof name, "%s/output_%05d.vtk", st.output_folder.c_str(), istep);
    output_paraview(s, g, name);
  };
  snapshot(0);

  now_time();
  printf(" # --- MAC algorithm start\n");
  double total_ms = 0, sor_ms = 0;
  for (int istep = 1; istep <= nloop; ++istep) {
    const double time = istep * g.dt;
    printf(" --- time_steps= %s  --  time = %s\n", f_int(istep).c_str(), f_real(time).c_str());
    double perr = 0;
    if (replay.on) {
      if ((size_t)istep > replay.perr.size()) { fprintf(stderr, " replay record too short\n"); return 2; }
      perr = replay.perr[istep - 1];
    } else if (pf_step(s, 1, &perr)) die(s, "pf_step");
    printf(" SOR iteration no.%s -- p error:%s\n", f_int(st.iter_max).c_str(), f_real(perr).c_str());
    if (!replay.on) {
      double a, b; long long nl;
      pf_last_timing(s, &a, &b, &nl);
      total_ms += a; sor_ms += b;
    }
    if (scase == PF_IBM2_DRAG) {   // call output_force_log_2d, ibm_2d_drag_omp_cpu.f90:121 (lib/output.f90:244-305)
      double F[8];
      if (replay.on) {
        if (replay.force.size() < 8 * (size_t)istep) { fprintf(stderr, " replay record holds no force log\n"); return 2; }
        memcpy(F, &replay.force[8 * (size_t)(istep - 1)], sizeof F);
      } else if (pf_force_log_2d(s, st.radius, F)) die(s, "pf_force_log_2d");
      printf(" Fp =%s%s\n Fv =%s%s\n F  =%s%s\n Cd =%s Cl =%s\n", f_real(F[0]).c_str(), f_real(F[1]).c_str(),
             f_real(F[2]).c_str(), f_real(F[3]).c_str(), f_real(F[4]).c_str(), f_real(F[5]).c_str(),
             f_real(F[6]).c_str(), f_real(F[7]).c_str());
    }
    if (st.istep_out > 0 && istep % st.istep_out == 0) snapshot(istep);
  }
  now_time();
  if (replay.on) { u = replay.u; v = replay.v; w = replay.w; p = replay.p; }
  else if (pf_gather(s, u.data(), v.data(), d3 ? w.data() : nullptr, p.data())) die(s, "pf_gather");
  if (!no_output) {
    if (rank == 0) {
      output_solution(g, u, v, w, p);
      output_divergent(g, u, v, w);
    }
    output_paraview(s, g, st.output_folder + "/output_paraview.vtk", true);
  }
  const double cells = (double)g.m * g.n * (d3 ? g.l : 1);
  if (!replay.on && rank == 0) fprintf(stderr, " [pixelflow_gpu] %d GPU(s), %d steps, %.3f ms/step on the device (%.3f ms in SOR), %.1f M cell-updates/s; "
                  "%.3f s in VTK snapshots\n",
          nranks, nloop, total_ms / nloop, sor_ms / nloop, cells * nloop / (total_ms * 1e-3) / 1e6, g_output_seconds);
  if (s) pf_destroy(s);
  if (pf_ranks_finish(0)) { fprintf(stderr, " a GPU rank failed\n"); return 1; }   // ranks > 0 end here
  printf(" program finished\n");
  now_time();
  return 0;
}
