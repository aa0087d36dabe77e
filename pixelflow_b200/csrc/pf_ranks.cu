// pf_ranks.cu -- one process per GPU WITHOUT a launcher: the helpers a plain driver program (the reference's
// `program main`, pixelflow_b200/fortran/*.f90, or the C++ twin) uses to run its deck z-slab decomposed over N GPUs.
//
// The reference is one process in one address space; BASELINE's north star wants ITS driver slab-
// decomposed over 1, 2, 4, 8 GPUs.  With these five entry points the driver stays one program started once, with no
// mpirun / torchrun around it:
//
//     call pf_ranks_launch(ngpus, rank)      ! fork: from here on `ngpus` copies of the program run, rank = 0 .. ngpus-1
//                                            ! (ngpus = 0: the count is the environment variable PIXELFLOW_GPUS)
//     cfg%rank = rank; cfg%nranks = ngpus; cfg%device = rank; cfg%nccl_unique_id = pf_ranks_unique_id()
//     ... pf_create / pf_set_porosity / pf_step as on one GPU (global-shaped host arrays: every rank reads the deck) ...
//     call pf_gather(s, u, v, w, p)          ! the whole fields in rank 0's arrays, for the reference's output routines
//     call pf_ranks_finish(0)                ! ranks > 0 end here; rank 0 returns once they have
//
// pf_ranks_launch must run before anything touches CUDA (a forked child cannot inherit a CUDA context).  Rank 0 is
// the original process: its stdout stays the log; the children's stdout goes to /dev/null so that the reference's
// write(*,*) lines are printed once.  The ranks meet through one page of shared memory (the NCCL id, a barrier, a
// failure flag) mapped before the fork.  See INTEGRATION.md, "Several GPUs from one driver".
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <atomic>

#include "pf_internal.cuh"

namespace {

struct RankPage {
  std::atomic<int> uid_state;      // 0 = not made, 1 = ready, -1 = rank 0 failed to make it
  unsigned char uid[128];
  std::atomic<int> arrived;        // sense-reversing barrier
  std::atomic<int> generation;
  std::atomic<int> failed;         // a rank called pf_ranks_finish with a non-zero status, or died
  int nranks;
};

RankPage *g_page = nullptr;
int g_rank = 0, g_nranks = 1;
pid_t g_children[64];

double now_s() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

// rank 0 notices children that died without reaching pf_ranks_finish
void reap(bool block) {
  if (g_rank != 0) return;
  for (int r = 1; r < g_nranks; ++r) {
    if (g_children[r] <= 0) continue;
    int st = 0;
    const pid_t w = waitpid(g_children[r], &st, block ? 0 : WNOHANG);
    if (w == g_children[r]) {
      g_children[r] = 0;
      if (!(WIFEXITED(st) && WEXITSTATUS(st) == 0)) g_page->failed.store(1);
    }
  }
}

}  // namespace

extern "C" {

int pf_ranks_launch(int nranks, int *rank) {
  if (!rank) return 1;
  *rank = 0;
  if (nranks == 0) {   // the count comes from the environment, like OMP_NUM_THREADS does for the reference
    const char *e = getenv("PIXELFLOW_GPUS");
    nranks = e ? atoi(e) : 1;
  }
  if (nranks <= 1) { g_nranks = 1; return 0; }
  if (g_page) { pf_set_global_error("pf_ranks_launch was already called"); return 1; }
  if (nranks > 64) { pf_set_global_error("pf_ranks_launch: at most 64 ranks"); return 1; }
  void *mem = mmap(nullptr, sizeof(RankPage), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (mem == MAP_FAILED) { pf_set_global_error(std::string("pf_ranks_launch: mmap: ") + strerror(errno)); return 1; }
  g_page = new (mem) RankPage();
  g_page->uid_state.store(0);
  g_page->arrived.store(0);
  g_page->generation.store(0);
  g_page->failed.store(0);
  g_page->nranks = nranks;
  g_nranks = nranks;
  fflush(stdout);
  fflush(stderr);
  for (int r = 1; r < nranks; ++r) {
    const pid_t pid = fork();
    if (pid < 0) {
      pf_set_global_error(std::string("pf_ranks_launch: fork: ") + strerror(errno));
      g_page->failed.store(1);
      return 1;
    }
    if (pid == 0) {   // child = rank r: the log is rank 0's
      g_rank = r;
      const int devnull = open("/dev/null", O_WRONLY);
      if (devnull >= 0) { dup2(devnull, STDOUT_FILENO); close(devnull); }
      *rank = r;
      return 0;
    }
    g_children[r] = pid;
  }
  g_rank = 0;
  return 0;
}

int pf_ranks_rank(void) { return g_rank; }
int pf_ranks_count(void) { return g_nranks; }

// the 128-byte NCCL id of this run: made by rank 0 on first use, read by the others (NULL on one rank or on failure)
const void *pf_ranks_unique_id(void) {
  if (!g_page) return nullptr;
  if (g_rank == 0) {
    if (g_page->uid_state.load() == 0) {
      std::string err;
      if (pf_comm_get_unique_id(g_page->uid, err)) {
        pf_set_global_error(err);
        g_page->uid_state.store(-1);
        g_page->failed.store(1);
        return nullptr;
      }
      g_page->uid_state.store(1);
    }
    return g_page->uid_state.load() == 1 ? g_page->uid : nullptr;
  }
  const double t0 = now_s();
  while (g_page->uid_state.load() == 0) {
    if (g_page->failed.load() || now_s() - t0 > 300.) { pf_set_global_error("pf_ranks_unique_id: rank 0 did not publish the id"); return nullptr; }
    usleep(200);
  }
  return g_page->uid_state.load() == 1 ? g_page->uid : nullptr;
}

// all ranks arrive, or the call fails (non-zero) on every rank that is still alive once one rank has failed
int pf_ranks_barrier(void) {
  if (!g_page) return 0;
  const int gen = g_page->generation.load();
  if (g_page->arrived.fetch_add(1) + 1 == g_nranks) {
    g_page->arrived.store(0);
    g_page->generation.fetch_add(1);
    return g_page->failed.load() ? 1 : 0;
  }
  const double t0 = now_s();
  while (g_page->generation.load() == gen) {
    reap(false);
    if (g_page->failed.load()) { pf_set_global_error("pf_ranks_barrier: another rank failed"); return 1; }
    if (now_s() - t0 > 3600.) { pf_set_global_error("pf_ranks_barrier: timeout"); return 1; }
    usleep(100);
  }
  return g_page->failed.load() ? 1 : 0;
}

// End of the run.  Ranks > 0 never return: they leave with `status`.  Rank 0 waits for them and returns 0 only if
// every rank finished with status 0.
int pf_ranks_finish(int status) {
  if (!g_page) return status;
  if (status) g_page->failed.store(1);
  if (g_rank != 0) {
    fflush(stderr);
    _exit(status ? 1 : 0);
  }
  reap(true);
  const int bad = g_page->failed.load() || status;
  munmap(g_page, sizeof(RankPage));
  g_page = nullptr;
  g_nranks = 1;
  return bad ? 1 : 0;
}

}  // extern "C"
