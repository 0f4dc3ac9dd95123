/* integration/c_env.c — an environment written in C against the reference's C interface for Fortran/C apps
 * (include/smarties_extern.h:28-102: an opaque communicator pointer, `const double*` + length arguments, no return
 * codes), the same entry points apps/cart_pole_f90/app_main.f90 binds through iso_c_binding.  gfortran is not in
 * this image, so this C file stands in for a Fortran app: identical ABI, identical calls.  Ours, not a copy of a
 * reference app: a 1-D double integrator that has to be parked at the origin with a bounded force. */
#include <math.h>

void smarties_setStateActionDims(void* comm, int state_dim, int action_dim, int agent_id);
void smarties_setActionScales(void* comm, const double* upper, const double* lower, int are_bounds, int action_dim, int agent_id);
void smarties_sendInitState(void* comm, const double* S, int state_dim, int agentID);
void smarties_sendState(void* comm, const double* S, int state_dim, double R, int agentID);
void smarties_sendTermState(void* comm, const double* S, int state_dim, double R, int agentID);
void smarties_sendLastState(void* comm, const double* S, int state_dim, double R, int agentID);
void smarties_recvAction(void* comm, double* A, int action_dim, int agentID);
void smarties_getUniformRandom(void* comm, double begin, double end, double* sampled);

void app_main(void* comm, int f_mpicomm)
{
  (void)f_mpicomm;
  const double up[1] = {2.0}, lo[1] = {-2.0};
  smarties_setStateActionDims(comm, 2, 1, 0);
  smarties_setActionScales(comm, up, lo, 1, 1, 0);
  for (;;) {
    double s[2], a[1];
    smarties_getUniformRandom(comm, -1.0, 1.0, &s[0]);
    smarties_getUniformRandom(comm, -0.5, 0.5, &s[1]);
    smarties_sendInitState(comm, s, 2, 0);
    for (int step = 1; ; ++step) {
      smarties_recvAction(comm, a, 1, 0);
      s[1] += 0.1 * a[0];
      s[0] += 0.1 * s[1];
      const double r = -(s[0] * s[0]) - 0.1 * s[1] * s[1] - 0.01 * a[0] * a[0];
      if (fabs(s[0]) > 4.0) { smarties_sendTermState(comm, s, 2, r - 5.0, 0); break; }
      if (step == 150) { smarties_sendLastState(comm, s, 2, r, 0); break; }
      smarties_sendState(comm, s, 2, r, 0);
    }
  }
}
