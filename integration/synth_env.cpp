// integration/synth_env.cpp — a synthetic environment with the shape of BASELINE.json configs[3]
// (apps/OpenAI_gym HalfCheetah-v3: 17 state components, 6 actions bounded to [-1, 1], episodes
// truncated after 1000 steps), written against the reference's PUBLIC app-side API only
// (include/smarties.h: smarties::Engine, smarties::Communicator — Communicator.h:41-216).  It
// exists because gym / mujoco are not in this image; the learner, the Communicator, the sockets
// and the worker threads it exercises are the reference's own.
//
// Dynamics: a stable random linear system x' = A x + B a + noise observed through a fixed random
// projection; reward = forward "velocity" x[0] minus a control cost, so there is something to learn.
//
//   synth_env --nEnvironments 64 --nTrainSteps 50000 --nThreads 8   (settings.json in the run dir)
#include "smarties.h"

#include <cmath>
#include <random>
#include <vector>

static constexpr int DS = 17, DA = 6, HORIZON = 1000;

struct Dyn {
  double A[DS][DS], B[DS][DA];
  Dyn() {
    std::mt19937 g(1234);                    // the same system in every environment process
    std::normal_distribution<double> n(0, 1);
    for (int i = 0; i < DS; ++i) {
      for (int j = 0; j < DS; ++j) A[i][j] = (i == j ? 0.9 : 0.0) + 0.05 * n(g) / std::sqrt((double)DS);
      for (int j = 0; j < DA; ++j) B[i][j] = 0.3 * n(g);
    }
  }
};

inline void app_main(smarties::Communicator* const comm, int argc, char** argv)
{
  comm->setStateActionDims(DS, DA);
  std::vector<double> up(DA, 1.0), lo(DA, -1.0);
  comm->setActionScales(up, lo, true);       // bounded -> SquashedNormalPolicy on the learner side
  static const Dyn dyn;
  std::normal_distribution<double> noise(0, 0.05);
  std::vector<double> x(DS), xn(DS);

  while (true) {
    std::mt19937& gen = comm->getPRNG();
    for (int i = 0; i < DS; ++i) x[i] = 0.1 * noise(gen) / 0.05;
    comm->sendInitState(x);
    for (int t = 1; ; ++t) {
      const std::vector<double> a = comm->recvAction();
      if (comm->terminateTraining()) return;
      double cost = 0;
      for (int j = 0; j < DA; ++j) cost += a[j] * a[j];
      for (int i = 0; i < DS; ++i) {
        double v = noise(gen);
        for (int j = 0; j < DS; ++j) v += dyn.A[i][j] * x[j];
        for (int j = 0; j < DA; ++j) v += dyn.B[i][j] * a[j];
        xn[i] = v;
      }
      x = xn;
      const double reward = x[0] - 0.1 * cost;
      if (t >= HORIZON) { comm->sendLastState(x, reward); break; }   // time-out, not a terminal state
      comm->sendState(x, reward);
    }
  }
}

int main(int argc, char** argv)
{
  smarties::Engine e(argc, argv);
  if (e.parse()) return 1;
  e.run(app_main);
  return 0;
}
