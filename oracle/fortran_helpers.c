/* The two contained procedures of pixelflow_b200/fortran/pixelflow_gpu_mod.f90 that lie outside the subset
 * oracle/f90toc.py translates (allocatable deferred-length character results, c_f_pointer, transfer), written in C
 * by hand, statement for statement.  TEST INFRASTRUCTURE: used only when the Fortran driver is run through the
 * translator.  Error path only — no numerics.
 *
 *   function pf_error_message(handle) result(msg)     : the C string pf_last_error returns, as text
 *   subroutine pf_check(ierr, handle, what)           : if (ierr /= 0) write(*,*) 'pixelflow_gpu: ', what,
 *                                                       ' failed: ', pf_error_message(handle); stop 1
 */
#include "ref_runtime.h"

const char *pf_last_error(void *handle);

const char *ft_pf_error_message(void *handle) {
  const char *cp = pf_last_error(handle);
  return cp ? cp : "";
}

void ft_pf_check(int ierr, void *handle, const char *what) {
  if (ierr != 0) {
    rt_write_begin(-1);
    rt_write_str("pixelflow_gpu: ");
    rt_write_str(what);
    rt_write_str(" failed: ");
    rt_write_str(ft_pf_error_message(handle));
    rt_write_end();
    rt_stop();
  }
}
