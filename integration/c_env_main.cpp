// integration/c_env_main.cpp — the C++ `main` of a C/Fortran app, in the shape of apps/cart_pole_f90/main.cpp:
// Engine + a trampoline that hands the communicator to the foreign-language `app_main` (integration/c_env.c).
#include "smarties.h"

extern "C" void app_main(void* smarties_comm, int f_mpicomm);

inline void app_main_interface(smarties::Communicator* const comm, int argc, char** argv) { app_main(comm, 0); }

int main(int argc, char** argv)
{
  smarties::Engine e(argc, argv);
  if (e.parse()) return 1;
  e.run(app_main_interface);
  return 0;
}
