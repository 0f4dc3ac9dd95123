// oracle/ref_harness.cpp — TEST INFRASTRUCTURE (parity oracle + CPU baseline), never shipped.
//
// Learner-only driver around the UNMODIFIED reference library (oracle/_ref/libsmarties.so,
// compiled from /root/reference by oracle/Makefile).  It fills the reference MemoryBuffer with
// a synthetic replay buffer read from a file, then runs the reference's own learner loop
//     spawnTrainTasks(); processMemoryBuffer(); applyGradient(); globalGradCounterUpdate();
// (reference: Learners/RACER.cpp:81-109, Learners/Learner_approximator.cpp:36-105,
//  Learners/Learner.cpp:74-100,130-133) and either times it (CPU baseline) or dumps golden
// vectors (sampled indices, net outputs, importance weights, output gradients, parameter
// gradient, weights after Adam, Retrace estimates, ReF-ER coefficients) for the parity tests.
//
// All arithmetic is the reference's; this file only constructs objects, copies data in and
// reads results out (compiled with -fno-access-control for that purpose).
#include "smarties/Learners/RACER.h"
#include "smarties/Learners/Learner_approximator.h"
#include "smarties/Network/Approximator.h"
#include "smarties/Network/Optimizer.h"
#include "smarties/Math/Continuous_policy.h"
#include "smarties/Math/Zero_advantage.h"
#include "smarties/Math/Gaus_advantage.h"
#include "smarties/Math/Discrete_policy.h"
#include "smarties/Math/Discrete_advantage.h"
#include "smarties/ReplayMemory/MemoryProcessing.h"
#include "smarties/Utils/Profiler.h"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

using namespace smarties;

// ---------------------------------------------------------------------------------------------
// record stream: [u32 name_len][name][u32 dtype: 0=f32 1=f64 2=i64][u32 ndim][u64 dims..][data]
struct Dump {
  FILE* f = nullptr;
  bool on() const { return f != nullptr; }
  void open(const std::string& path) { f = fopen(path.c_str(), "wb"); if(!f) { perror("dump"); exit(1);} }
  void close() { if(f) fclose(f); f = nullptr; }
  void put(const std::string& name, uint32_t dtype, const std::vector<uint64_t>& dims, const void* data, size_t bytes) {
    if(!f) return;
    uint32_t nl = name.size(), nd = dims.size();
    fwrite(&nl, 4, 1, f); fwrite(name.data(), 1, nl, f);
    fwrite(&dtype, 4, 1, f); fwrite(&nd, 4, 1, f);
    fwrite(dims.data(), 8, nd, f); fwrite(data, 1, bytes, f);
  }
  void f32(const std::string& n, const std::vector<float>& v, std::vector<uint64_t> d = {}) {
    if(d.empty()) d = {v.size()}; put(n, 0, d, v.data(), v.size()*4); }
  void f64(const std::string& n, const std::vector<double>& v, std::vector<uint64_t> d = {}) {
    if(d.empty()) d = {v.size()}; put(n, 1, d, v.data(), v.size()*8); }
  void i64(const std::string& n, const std::vector<int64_t>& v, std::vector<uint64_t> d = {}) {
    if(d.empty()) d = {v.size()}; put(n, 2, d, v.data(), v.size()*8); }
};

struct SynthData {
  int64_t dS = 0, dA = 0, nEp = 0;
  std::vector<int64_t> N, term, start;
  std::vector<float> S, A, MU, R;
  void load(const std::string& path, int64_t dPolicy = 0) {   // dPolicy: columns of MU (0 = 2*dA, Gaussian mean | stdev)
    FILE* f = fopen(path.c_str(), "rb");
    if(!f) { perror(path.c_str()); exit(1); }
    int64_t hdr[4];
    if(fread(hdr, 8, 4, f) != 4 || hdr[0] != 0x31424D53) { fprintf(stderr, "bad data file\n"); exit(1); }
    dS = hdr[1]; dA = hdr[2]; nEp = hdr[3];
    N.resize(nEp); term.resize(nEp); start.resize(nEp);
    int64_t tot = 0;
    for(int64_t i=0; i<nEp; ++i) {
      int64_t v[2]; if(fread(v, 8, 2, f) != 2) exit(1);
      N[i] = v[0]; term[i] = v[1]; start[i] = tot; tot += v[0];
    }
    S.resize(tot*dS); A.resize(tot*dA); MU.resize(tot*(dPolicy > 0 ? dPolicy : 2*dA)); R.resize(tot);
    if(fread(S.data(), 4, S.size(), f) != S.size()) exit(1);
    if(fread(A.data(), 4, A.size(), f) != A.size()) exit(1);
    if(fread(MU.data(), 4, MU.size(), f) != MU.size()) exit(1);
    if(fread(R.data(), 4, R.size(), f) != R.size()) exit(1);
    fclose(f);
  }
};

struct Args {
  std::string data, settings, dump, weights, restart;
  int steps = 10, threads = 1, bounded = 0, dumpAll = 0, quiet = 0, reps = 1, save = 0;
  int warmup = 0;     // learner steps run before the timed repetitions (thread team, caches and page tables warm)
  int discrete = 0;   // > 0: one discrete action component with that many options (RACER<Discrete_advantage, Discrete_policy, Uint>)
  long startStep = 0;
  unsigned long seed = 42, sampleSeed = 0;
  std::set<long> dumpSteps;
};

static Args parse(int argc, char** argv) {
  Args a;
  for(int i=1; i<argc; ++i) {
    std::string k = argv[i];
    auto next = [&]() { if(i+1>=argc) { fprintf(stderr, "missing value for %s\n", k.c_str()); exit(1);} return std::string(argv[++i]); };
    if(k=="--data") a.data = next();
    else if(k=="--settings") a.settings = next();
    else if(k=="--dump") a.dump = next();
    else if(k=="--weights") a.weights = next();
    else if(k=="--steps") a.steps = std::stoi(next());
    else if(k=="--threads") a.threads = std::stoi(next());
    else if(k=="--bounded") a.bounded = std::stoi(next());
    else if(k=="--discrete") a.discrete = std::stoi(next());
    else if(k=="--startStep") a.startStep = std::stol(next());
    else if(k=="--seed") a.seed = std::stoul(next());
    else if(k=="--sampleSeed") a.sampleSeed = std::stoul(next());
    else if(k=="--dumpAll") a.dumpAll = 1;
    else if(k=="--quiet") a.quiet = 1;
    else if(k=="--save") a.save = 1;                 // Learner_approximator::save() after the last step (files agent_00_* in the cwd)
    else if(k=="--restart") a.restart = next();      // Learner_approximator::restart() from that directory instead of filling the buffer
    else if(k=="--reps") a.reps = std::stoi(next());
    else if(k=="--warmup") a.warmup = std::stoi(next());
    else if(k=="--dumpSteps") { std::stringstream ss(next()); std::string tok; while(std::getline(ss, tok, ',')) a.dumpSteps.insert(std::stol(tok)); }
    else { fprintf(stderr, "unknown arg %s\n", k.c_str()); exit(1); }
  }
  return a;
}

// ---------------------------------------------------------------------------------------------
template<typename Base>
struct Probe : public Base
{
  using Base::networks; using Base::data; using Base::settings; using Base::profiler;
  mutable bool recording = false;
  mutable std::vector<double> recO;
  mutable std::vector<float> recG, recS;
  mutable std::vector<int64_t> recT, recEp;
  Uint nOut = 0, dS = 0;

  Probe(MDPdescriptor& M, HyperParameters& S, ExecutionInfo& D) : Base(M, S, D) {
    nOut = networks[0]->nOutputs(); dS = M.dimStateObserved;
  }

  void Train(const MiniBatch& MB, const Uint wID, const Uint bID) const override {
    Base::Train(MB, wID, bID);
    if(!recording) return;
    const Approximator& NET = * networks[0];
    const Uint t = MB.sampledTstep(bID);
    const Rvec O = NET.forward(bID, t); // cached activation, no recompute
    const auto& C = NET.getContext(bID);
    const std::vector<nnReal> g = C.activation(t, 0)->getOutputDelta();
    const NNvec& s = MB.state(bID, t);
    for(Uint i=0; i<nOut; ++i) { recO[bID*nOut+i] = O[i]; recG[bID*nOut+i] = g[i]; }
    for(Uint i=0; i<dS; ++i) recS[bID*dS+i] = s[i];
    recT[bID] = t; recEp[bID] = MB.getEpisode(bID).ID;
  }

  AdamOptimizer* adam() const { return dynamic_cast<AdamOptimizer*>(networks[0]->opt.get()); }
  std::vector<float> blob(const Parameters* P) const { return std::vector<float>(P->params, P->params + P->nParams); }

  void dumpTransitions(Dump& D, const std::string& pre) const {
    std::vector<float> V, A, Q, dlt, rho, kl, agg;
    std::vector<int64_t> ids, lens;
    for(long i=0; i<data->nStoredEps(); ++i) {
      const Episode& EP = data->get(i);
      ids.push_back(EP.ID); lens.push_back(EP.nsteps());
      V.insert(V.end(), EP.stateValue.begin(), EP.stateValue.end());
      A.insert(A.end(), EP.actionAdvantage.begin(), EP.actionAdvantage.end());
      Q.insert(Q.end(), EP.returnEstimator.begin(), EP.returnEstimator.end());
      dlt.insert(dlt.end(), EP.deltaValue.begin(), EP.deltaValue.end());
      rho.insert(rho.end(), EP.offPolicImpW.begin(), EP.offPolicImpW.end());
      kl.insert(kl.end(), EP.KullbLeibDiv.begin(), EP.KullbLeibDiv.end());
      const float a[10] = { EP.avgKLDivergence, EP.fracFarPolSteps, EP.avgSquaredErr, EP.maxAbsError,
                            EP.sumSquaredQ, EP.sumQ, EP.maxQ, EP.minQ, EP.totR, (float) EP.just_sampled };
      agg.insert(agg.end(), a, a+10);
    }
    D.i64(pre+"/epID", ids); D.i64(pre+"/epLen", lens);
    D.f32(pre+"/V", V); D.f32(pre+"/A", A); D.f32(pre+"/Qret", Q);
    D.f32(pre+"/delta", dlt); D.f32(pre+"/rho", rho); D.f32(pre+"/KL", kl);
    D.f32(pre+"/epAgg", agg, {(uint64_t) ids.size(), 10});
  }

  void dumpScaling(Dump& D, const std::string& pre) const {
    const MDPdescriptor& M = data->MDP;
    D.f32(pre+"/stateMean", M.stateMean); D.f32(pre+"/stateScale", M.stateScale);
    D.f32(pre+"/stateStdDev", M.stateStdDev);
    D.f32(pre+"/rewards", {M.rewardsMean, M.rewardsScale, M.rewardsStdDev});
  }

  void dumpRefer(Dump& D, const std::string& pre) const {
    D.f64(pre+"/refer", { data->beta, data->CmaxRet, data->CinvRet, (double) data->stats.nFarPolicySteps,
      data->stats.avgKLdivergence, data->stats.avgSquaredErr, data->stats.maxAbsError, data->stats.avgReturn,
      data->stats.stdevQ, data->stats.avgQ, data->stats.maxQ, data->stats.minQ,
      (double) data->stats.countReturnsEstimateUpdates, data->stats.sumReturnsEstimateErrors });
  }

  int run(const Args& args, const SynthData& SD, ExecutionInfo& distrib)
  {
    MDPdescriptor& MDP = data->MDP;
    const Uint dA = MDP.dimAction, dP = MDP.policyVecDim;
    // ---- fill the replay memory (recipe: SURVEY.md §8c) ----
    const bool restarted = args.restart.size() > 0;
    if(restarted) { distrib.restart = args.restart; this->restart(); }    // Learner_approximator.cpp:118-131
    for(int64_t e=0; e<SD.nEp && !restarted; ++e) {
      std::unique_ptr<Episode> EP = std::make_unique<Episode>(MDP);
      const int64_t N = SD.N[e], o = SD.start[e];
      for(int64_t t=0; t<N; ++t) {
        const float* s = &SD.S[(o+t)*dS];
        EP->states.push_back(Fvec(s, s+dS));
        EP->latent_states.push_back(Fvec());
        const bool last = t+1 == N;
        Rvec a(dA, 0), p(dP, 0);
        if(!last) {
          for(Uint i=0;i<dA;++i) a[i] = SD.A[(o+t)*dA+i];
          for(Uint i=0;i<dP;++i) p[i] = SD.MU[(o+t)*dP+i];
        }
        EP->actions.push_back(a); EP->policies.push_back(p);
        EP->rewards.push_back(t==0 ? 0 : (Real) SD.R[o+t]);
        EP->totR += t==0 ? 0 : SD.R[o+t];
      }
      EP->bReachedTermState = SD.term[e] != 0;
      EP->agentID = 0;
      EP->finalize(e);
      MemoryProcessing::computeReturnEstimator(* data.get(), * EP.get());
      data->counters.nSeenTransitions_loc += N-1;
      data->counters.nSeenEpisodes_loc ++;
      data->pushBackEpisode(std::move(EP));
      // pushBackEpisode stamps ID=max(nLocTimeStepsTrain,0)=0 for pre-training data; give each
      // episode a distinct, insertion-ordered ID so FIFO ordering is well defined:
      data->episodes.back()->ID = e;
    }

    if(args.weights.size()) {
      FILE* f = fopen(args.weights.c_str(), "rb");
      Parameters* W = adam()->weights.get();
      if(!f || fread(W->params, 4, W->nParams, f) != W->nParams) { fprintf(stderr, "bad weights file\n"); return 1; }
      fclose(f);
    }

    Dump D; if(args.dump.size()) D.open(args.dump);
    const Uint B = settings.batchSize_local, nParams = adam()->weights->nParams;
    {
      std::vector<int64_t> lsz;
      for(const auto& l : networks[0]->net->layers) lsz.push_back(l->size);
      D.i64("meta/dims", {(int64_t) dS, (int64_t) dA, (int64_t) nOut, (int64_t) B, (int64_t) nParams,
                          (int64_t) SD.nEp, (int64_t) data->nStoredSteps(), (int64_t) args.startStep});
      D.i64("meta/layerSizes", lsz);
      D.f64("meta/hyper", {settings.gamma, settings.lambda, settings.clipImpWeight, settings.penalTol,
                           settings.epsAnneal, settings.learnrate, settings.nnLambda, settings.explNoise,
                           (double) settings.maxTotObsNum, (double) settings.batchSize});
      D.f32("init/weights", blob(adam()->weights.get()));
      dumpTransitions(D, "preinit");
    }

    this->initializeLearner();
    if(D.on()) { dumpScaling(D, "init"); dumpTransitions(D, "init"); dumpRefer(D, "init"); }

    // start the gradient-step counter where asked (so that short runs cover the every-1000-steps
    // sweeps); Adam's own step counter follows as in a restart (Approximator.h:64)
    if(!restarted) {
      data->counters.nGradSteps = args.startStep;
      networks[0]->setNgradSteps(args.startStep);
    }
    if(args.sampleSeed) distrib.generators[0].seed(args.sampleSeed);

    recO.assign(B*nOut, 0); recG.assign(B*nOut, 0); recS.assign(B*dS, 0); recT.assign(B, 0); recEp.assign(B, 0);
    std::vector<double> trBeta, trCmax, trWnorm; std::vector<int64_t> trNfar;

    for(int s=0; s<args.warmup; ++s) {   // untimed warm-up steps of the same loop
      this->spawnTrainTasks(); this->processMemoryBuffer(); this->applyGradient(); this->globalGradCounterUpdate();
    }
    double bestSec = 1e300, totSec = 0;
    std::vector<double> repSec;
    for(int rep=0; rep<args.reps; ++rep)
    {
      const auto t0 = std::chrono::steady_clock::now();
      for(int s=0; s<args.steps; ++s)
      {
        const bool dumpThis = D.on() && rep==0 && (args.dumpAll || args.dumpSteps.count(s));
        recording = dumpThis;
        const std::string pre = "s" + std::to_string(s);
        if(dumpThis) dumpRefer(D, pre+"/pre");
        this->spawnTrainTasks();
        if(dumpThis) {
          D.i64(pre+"/sampledEpID", recEp); D.i64(pre+"/sampledT", recT);
          D.f64(pre+"/O", recO, {B, nOut}); D.f32(pre+"/g", recG, {B, nOut}); D.f32(pre+"/S", recS, {B, dS});
          D.f32(pre+"/gradSum", blob(adam()->gradSum.get()));
        }
        this->processMemoryBuffer();
        this->applyGradient();
        this->globalGradCounterUpdate();
        // "targetDelay" > 0: AdamOptimizer::target_weights after every update (Optimizer.cpp:162-177)
        if(settings.targetDelay > 0) D.f32(pre+"/tgt", blob(adam()->target_weights.get()));
        if(dumpThis) {
          D.f32(pre+"/weights", blob(adam()->weights.get()));
          D.f32(pre+"/m1", blob(adam()->_1stMom.get()));
          D.f32(pre+"/m2", blob(adam()->_2ndMom.get()));
          dumpRefer(D, pre+"/post"); dumpScaling(D, pre+"/post");
          dumpTransitions(D, pre+"/post");
        }
        if(D.on() && rep==0) {
          trBeta.push_back(data->beta); trCmax.push_back(data->CmaxRet);
          trNfar.push_back(data->stats.nFarPolicySteps);
          trWnorm.push_back((double) adam()->weights->compute_weight_norm());
        }
      }
      const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      bestSec = std::min(bestSec, sec); totSec += sec; repSec.push_back(sec);
    }
    if(D.on()) {
      D.f64("trace/beta", trBeta); D.f64("trace/Cmax", trCmax); D.i64("trace/nFar", trNfar);
      D.f64("trace/wnorm", trWnorm);
      D.f32("final/weights", blob(adam()->weights.get()));
      dumpTransitions(D, "final"); dumpScaling(D, "final"); dumpRefer(D, "final");
      D.close();
    }
    if(args.save) this->save();     // Learner_approximator::save (Learner_approximator.cpp:133-142)
    if(!args.quiet) printf("%s\n", profiler->printStatAndReset().c_str());
    const double medSec = totSec / args.reps;      // mean over the repetitions (name kept)
    std::sort(repSec.begin(), repSec.end());
    const double median = repSec.empty() ? medSec : (repSec.size() % 2 ? repSec[repSec.size()/2]
                                                                        : 0.5 * (repSec[repSec.size()/2 - 1] + repSec[repSec.size()/2]));
    printf("{\"harness\": \"reference\", \"steps\": %d, \"reps\": %d, \"threads\": %d, \"batch\": %lu, "
           "\"seconds_mean\": %.6f, \"seconds_best\": %.6f, \"seconds_median\": %.6f, \"steps_per_s\": %.3f, "
           "\"transitions_per_s\": %.3f, \"transitions_per_s_median\": %.3f, \"nTransitions\": %ld, \"nEpisodes\": %ld}\n",
           args.steps, args.reps, args.threads, (unsigned long) B, medSec, bestSec, median, args.steps/medSec,
           B*args.steps/medSec, B*args.steps/median, data->nStoredSteps(), data->nStoredEps());
    return 0;
  }
};

int main(int argc, char** argv)
{
  Args args = parse(argc, argv);
  if(args.data.empty()) { fprintf(stderr, "usage: ref_harness --data FILE [--settings JSON] [--steps K] [--threads T] ...\n"); return 1; }
  omp_set_num_threads(args.threads);
  SynthData SD; SD.load(args.data, args.discrete);

  std::vector<std::string> av = {"ref_harness"};
  ExecutionInfo distrib(av);
  distrib.nThreads = args.threads;
  distrib.randSeed = args.seed;
  distrib.initialze();
  distrib.nAgents = 1; distrib.bIsMaster = true;
  distrib.nOwnedEnvironments = 1; distrib.nEnvironments = 1;
  distrib.logAllSamples = 0;
  distrib.learners_train_comm = MPI_COMM_WORLD;

  MDPdescriptor MDP;
  MDP.dimState = SD.dS; MDP.dimAction = SD.dA;
  MDP.bActionSpaceBounded = std::vector<bool>(SD.dA, args.bounded != 0);
  if(args.discrete > 0) {      // what Communicator::setNumberOfOptions sets up for an app (Communicator.cpp: discreteActionValues)
    if(SD.dA != 1) { fprintf(stderr, "--discrete needs a one-component action\n"); return 1; }
    MDP.discreteActionValues = std::vector<Uint>(1, (Uint) args.discrete);
  }
  MDP.synchronize([](void*, size_t){});

  HyperParameters settings(MDP.dimObs(), MDP.dimAct());
  std::ifstream ifs(args.settings);
  settings.initializeOpts(ifs, distrib);
  if(settings.returnsEstimator == "default") settings.returnsEstimator = "retrace";
  const ActionInfo aInfo(MDP);

  int ret = 1;
  if(args.discrete > 0 && settings.learner == "RACER") {
    using L = RACER<Discrete_advantage, Discrete_policy, Uint>;
    MDP.policyVecDim = L::getnDimPolicy(aInfo);
    Probe<L> learner(MDP, settings, distrib);
    learner.setLearnerName("agent_00", 0);
    ret = learner.run(args, SD, distrib);
  } else if(settings.learner == "VRACER") {
    using L = RACER<Zero_advantage, Continuous_policy, Rvec>;
    MDP.policyVecDim = L::getnDimPolicy(aInfo);
    Probe<L> learner(MDP, settings, distrib);
    learner.setLearnerName("agent_00", 0);
    ret = learner.run(args, SD, distrib);
  } else if(settings.learner == "RACER") {
    using L = RACER<Param_advantage, Continuous_policy, Rvec>;
    MDP.policyVecDim = L::getnDimPolicy(aInfo);
    Probe<L> learner(MDP, settings, distrib);
    learner.setLearnerName("agent_00", 0);
    ret = learner.run(args, SD, distrib);
  } else { fprintf(stderr, "unsupported learner %s\n", settings.learner.c_str()); }
  fflush(0);
  return ret;
}
