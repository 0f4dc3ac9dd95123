// integration/RACER_B200.cpp — the reference-side binding of libsmarties_b200.so.
//
// This is the ONE translation unit a maintainer of the reference adds to libsmarties (see
// INTEGRATION.md).  It is written against the reference's own headers and is compiled here by
// integration/Makefile from the sources under /root/reference (nothing of the reference is
// copied into this repository).  What it does:
//
//   * RACER_B200<Advantage> derives from the reference's RACER<Advantage, Continuous_policy, Rvec>
//     (Learners/RACER.h:44-98).  Everything actor-side stays the reference's: Learner::select
//     (Learners/Learner.cpp:31-45), RACER::selectAction / processTerminal (RACER.cpp:30-59) on the
//     host Approximator, MemoryBuffer::storeState/storeAction/terminateCurrentEpisode, the
//     DataCoordinator, the Communicator / Worker / Master plumbing and every app.
//   * setupTasks (RACER.cpp:61-110) is overridden: the three reference tasks
//     {initializeLearner} / {spawnTrainTasks, processMemoryBuffer} / {applyGradient,
//     globalGradCounterUpdate} become {smb200_initialize_learner} / {mirror finished episodes into
//     the HBM replay, smb200_train_steps(1), copy the new weights into the host network used by the
//     actors}.  The host MemoryBuffer keeps receiving episodes and is pruned first-in-first-out
//     exactly like MemoryProcessing::applyEpisodesRemovalAlgo (MemoryProcessing.cpp:327-351).
//   * smarties::createLearner (Learners/AlgoFactory.cpp:60-340) is wrapped: AlgoFactory.cpp is
//     compiled with -DcreateLearner=createLearner_reference and this file provides
//     createLearner, which returns a RACER_B200 when the environment variable SMARTIES_B200 is set
//     and the settings are ones the device path covers, and the reference learner otherwise.
//
// Errors of the C-ABI become die() (Utils/Warnings.h:36-44), like every other fatal error of the
// reference.
#include "smarties/Learners/AlgoFactory.h"
#include "smarties/Learners/RACER.h"
#include "smarties/Math/Continuous_policy.h"
#include "smarties/Math/Zero_advantage.h"
#include "smarties/Math/Gaus_advantage.h"
#include "smarties/Math/Discrete_policy.h"
#include "smarties/Math/Discrete_advantage.h"
#include "smarties/Network/Approximator.h"
#include "smarties/Network/Optimizer.h"
#include "smarties/ReplayMemory/MemoryProcessing.h"
#include "smarties/Utils/Warnings.h"
#include "smarties/Utils/SstreamUtilities.h"

#include <smarties_b200.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <cstdlib>
#include <fstream>
#include <unistd.h>
#include <unordered_map>
#include <unordered_set>

namespace smarties
{

// the reference's own factory (AlgoFactory.cpp built with -DcreateLearner=createLearner_reference)
std::unique_ptr<Learner> createLearner_reference(const Uint learnerID, MDPdescriptor& MDP, ExecutionInfo& distrib);

template<typename Advantage_t, typename Policy_t = Continuous_policy, typename Action_t = Rvec>
class RACER_B200 : public RACER<Advantage_t, Policy_t, Action_t>
{
  using Base = RACER<Advantage_t, Policy_t, Action_t>;
  using Base::data; using Base::settings; using Base::distrib; using Base::MDP; using Base::networks;
  using Base::algoSubStepID; using Base::nObsB4StartTraining; using Base::bTrain; using Base::aInfo;
  using Base::profiler; using Base::learn_rank; using Base::learn_size;
  using Base::pol_start; using Base::adv_start; using Base::VsID;

  smb200_learner* gpu = nullptr;
  const bool isRacer;
  std::unordered_set<const Episode*> mirrored;      // episodes already resident in the HBM replay
  // Episode::ID (a time stamp) is not unique; the device gets ID * 2^20 + sequence number, which keeps the
  // first-in-first-out order of the reference (ID descending) and identifies the episode for save()
  std::unordered_map<int64_t, Episode*> byTag;
  std::unordered_map<const Episode*, int64_t> tagOf;
  int64_t nextSeq = 0;
  std::vector<float> bufS, bufA, bufMU, bufR, bufV, bufADV, wblob;
  smb200_step_stats last{};
  int maxStepsPerCall = 16;
  std::vector<smb200_step_stats> stepStats;
  double secPush = 0, secStep = 0, secSync = 0, tTrainStart = 0;
  long nPushed = 0;

  // ---- actors on the device (SMARTIES_B200_ACTORS=1): RACER::selectAction / processTerminal evaluate the policy with
  //      smb200_forward_seq instead of the host Approximator.  The worker threads of Master::waitForStateActionCallers
  //      (Core/Master.cpp:88-145) call Learner::select concurrently, one agent each: requests that arrive while a device call
  //      is in flight are answered together by the next call (the first waiting thread leads it), so n pending agents
  //      cost one launch and one round trip, not n. ----
  bool deviceActors = false;
  struct FwdReq { const float* states; int len; float* out; bool done; };
  std::mutex fwdMutex; std::condition_variable fwdCv; std::vector<FwdReq*> fwdPending; bool fwdLeader = false;
  std::vector<float> fwdS, fwdO; std::vector<int32_t> fwdL;      // the leader's packing buffers (one leader at a time)
  long nFwdCalls = 0, nFwdAgents = 0, nFwdMax = 0;

  // scaleNet2V (Learners/RACER_common.cpp:23-27; a template in a .cpp of the reference, not reachable from here)
  static Real net2V(const Real x) { return x > 0 ? 100 * (x + 51) - 100 * std::sqrt(2601 + 100 * x) : 100 * (x - 51) + 100 * std::sqrt(2601 - 100 * x); }

  Rvec deviceForward(const MiniBatch& MB)
  {
    const Episode& EP = * MB.episodes[0];
    const Uint dS = MDP.dimStateObserved, nOut = (Uint) smb200_n_outputs(gpu);
    const Sint t0 = MB.begTimeStep[0], t1 = MB.endTimeStep[0];       // the window of MemoryBuffer::agentToMinibatch
    const int len = (int) (t1 - t0);
    std::vector<float> st((size_t) len * dS), out(nOut);
    for (Sint t = t0; t < t1; ++t)
      for (Uint k = 0; k < dS; ++k) st[(size_t) (t - t0) * dS + k] = (float) EP.states[t][k];     // raw: the device standardises
    FwdReq rq{st.data(), len, out.data(), false};
    std::unique_lock<std::mutex> lk(fwdMutex);
    fwdPending.push_back(&rq);
    while (!rq.done) {
      if (fwdLeader) { fwdCv.wait(lk); continue; }
      fwdLeader = true;                       // lead one device call for everything that is pending now (this request included)
      std::vector<FwdReq*> batch; batch.swap(fwdPending);
      lk.unlock();
      const int n = (int) batch.size();
      int maxLen = 1;
      for (const FwdReq* r : batch) maxLen = std::max(maxLen, r->len);
      fwdS.assign((size_t) n * maxLen * dS, 0.f); fwdL.resize(n); fwdO.resize((size_t) n * nOut);
      for (int i = 0; i < n; ++i) {
        std::copy(batch[i]->states, batch[i]->states + (size_t) batch[i]->len * dS, fwdS.begin() + (size_t) i * maxLen * dS);
        fwdL[i] = batch[i]->len;
      }
      check(smb200_forward_seq(gpu, fwdS.data(), fwdL.data(), n, maxLen, fwdO.data()), "forward_seq");
      for (int i = 0; i < n; ++i) std::copy(fwdO.begin() + (size_t) i * nOut, fwdO.begin() + (size_t) (i + 1) * nOut, batch[i]->out);
      lk.lock();
      ++nFwdCalls; nFwdAgents += n; nFwdMax = std::max<long>(nFwdMax, n);
      for (FwdReq* r : batch) r->done = true;
      fwdLeader = false;
      fwdCv.notify_all();
    }
    return Rvec(out.begin(), out.end());
  }

  void check(const int rc, const char* what) const {
    if (rc) _die("smarties_b200 %s: %s", what, smb200_last_error());
  }

  Parameters* hostWeights() const { return networks[0]->opt->weights.get(); }

  void createDeviceLearner()
  {
    smb200_config c;
    check(smb200_default_config(&c, (int32_t) MDP.dimStateObserved, (int32_t) MDP.dimAction), "default_config");
    const char* dev = std::getenv("SMARTIES_B200_DEVICE");
    c.device = dev ? std::atoi(dev) : 0;
    c.algo = isRacer ? SMB200_RACER : SMB200_VRACER;
    // discrete action space (RACER<Discrete_advantage, Discrete_policy, Uint>): one component, its number of options
    if (MDP.bDiscreteActions()) c.discrete_options = (int32_t) MDP.discreteActionValues[0];
    // RACER::setupNet (RACER_common.cpp:82-91): createEncoder's layers and nnLayerSizes are stacked in ONE network ("encodr" is renamed "net")
    std::vector<Uint> hidden;
    for (const Uint n : settings.encoderLayerSizes) if (n > 0) hidden.push_back(n);
    for (const Uint n : settings.nnLayerSizes) if (n > 0) hidden.push_back(n);
    if (MDP.dimAction > SMB200_MAX_ACTION || hidden.size() > SMB200_MAX_HIDDEN) die("network too large for smarties_b200");
    for (Uint i = 0; i < MDP.dimAction; ++i) c.action_bounded[i] = MDP.bActionSpaceBounded[i] ? 1 : 0;
    c.n_hidden = (int32_t) hidden.size();
    for (int i = 0; i < c.n_hidden; ++i) c.hidden[i] = (int32_t) hidden[i];
    c.batch_size = (int32_t) settings.batchSize_local;  c.batch_size_global = (int32_t) settings.batchSize;
    c.max_tot_obs = settings.maxTotObsNum_local;         c.max_tot_obs_global = settings.maxTotObsNum;
    c.gamma = settings.gamma; c.lambda = settings.lambda; c.clip_imp_weight = settings.clipImpWeight;
    c.penal_tol = settings.penalTol; c.eps_anneal = settings.epsAnneal; c.learnrate = settings.learnrate;
    c.nn_lambda = settings.nnLambda; c.expl_noise = settings.explNoise; c.out_weights_prefac = settings.outWeightsPrefac;
    c.refer_reduce_threads = (int32_t) distrib.nThreads;
    c.world_rank = (int32_t) learn_rank; c.world_size = (int32_t) learn_size;
    c.seed = distrib.randSeed;
    // a partially observable MDP turns a feed-forward request into MGU layers (Network/Approximator.cpp:219-223)
    const bool mgu = settings.nnType == "MGU" || settings.nnType == "GRU" || (MDP.isPartiallyObservable && !settings.bRecurrent);
    c.nn_type = settings.nnType == "LSTM" ? SMB200_LSTM : (mgu ? SMB200_MGU : SMB200_FFNN);
    c.nn_bptt_seq = (int32_t) settings.nnBPTTseq;
    c.data_sampling = settings.dataSamplingAlgo == "PERrank" ? SMB200_SAMPLE_PER_RANK : settings.dataSamplingAlgo == "PERerr" ? SMB200_SAMPLE_PER_ERR
                    : settings.dataSamplingAlgo == "PERseq" ? SMB200_SAMPLE_PER_SEQ : SMB200_SAMPLE_UNIFORM;
    c.er_filter = settings.ERoldSeqFilter == "farpolfrac" ? SMB200_FILTER_FARPOLFRAC : settings.ERoldSeqFilter == "maxkldiv" ? SMB200_FILTER_MAXKLDIV
                : settings.ERoldSeqFilter == "minerror" ? SMB200_FILTER_MINERROR : SMB200_FILTER_OLDEST;
    c.nn_func = settings.nnFunc == "SoftSign" ? SMB200_SOFTSIGN : settings.nnFunc == "HardSign" ? SMB200_HARDSIGN : settings.nnFunc == "Sigm" ? SMB200_SIGM
              : settings.nnFunc == "Relu" ? SMB200_RELU : settings.nnFunc == "LRelu" ? SMB200_LRELU : settings.nnFunc == "ExpPlus" ? SMB200_EXPPLUS
              : settings.nnFunc == "SoftPlus" ? SMB200_SOFTPLUS : settings.nnFunc == "Exp" ? SMB200_EXP : settings.nnFunc == "Linear" ? SMB200_LINEAR : SMB200_TANH;
    c.target_delay = settings.targetDelay;       // RACER never evaluates the target weights; they are kept for the checkpoint
    c.returns_estimator = settings.returnsEstimator == "GAE" ? SMB200_GAE
                        : (settings.returnsEstimator == "retraceExplore" ? SMB200_RETRACE_EXPLORE : SMB200_RETRACE);
    // an episode occupies nsteps() = ndata()+1 rows and is at least two rows long: room for the worst case,
    // plus the episodes that arrive between two pruning passes
    c.capacity_rows = 2 * (int64_t) settings.maxTotObsNum_local + 65536;
    c.max_episodes = (int32_t) std::min<long>(1 << 20, (long) settings.maxTotObsNum_local + 4096);
    check(smb200_create(&c, &gpu), "create");
    Parameters* W = hostWeights();
    if ((int64_t) W->nParams != smb200_n_params(gpu))
      _die("parameter blob mismatch: reference %lu vs device %ld floats", (unsigned long) W->nParams, (long) smb200_n_params(gpu));
    // same initial policy on both sides: the device learner starts from the host network's weights
    check(smb200_set_weights(gpu, W->params, W->nParams), "set_weights");
    wblob.resize(W->nParams);
    // the actors read this blob while the device writes the next policy into it: page-locked, so that the copy is one DMA
    if (smb200_pin_host_buffer(W->params, (int64_t) (sizeof(nnReal) * W->nParams))) warn("smarties_b200: host weights stay pageable");
  }

  // the reference's output-gradient statistics file <learner>_<net>_outGrad_stats.raw (Approximator::updateGradStats,
  // Network/Approximator.h:65-68; written every 1000 steps by StatsTracker::printToFile); SMARTIES_B200_GRADSTATS=0: off.
  // Named on the first learner step: the factory sets learner_name after the constructor (Learner::setLearnerName).
  bool gradStatsNamed = false;
  void nameGradStats()
  {
    if (gradStatsNamed) return;
    gradStatsNamed = true;
    const char* gs = std::getenv("SMARTIES_B200_GRADSTATS");
    if (gs && std::atoi(gs) == 0) return;
    check(smb200_set_grad_stats(gpu, (this->learner_name + "_" + networks[0]->name).c_str()), "set_grad_stats");
  }

  // Episode (ReplayMemory/Episode.h:40-110) -> the row-major f32 arrays of smb200_push_episode
  void pushToDevice(const Episode& EP, const bool restoredValues = false)
  {
    const Uint N = EP.nsteps(), dS = MDP.dimStateObserved, dA = MDP.dimAction, dP = MDP.policyVecDim;
    bufS.assign((size_t) N * dS, 0.f); bufA.assign((size_t) N * dA, 0.f); bufMU.assign((size_t) N * dP, 0.f);
    bufR.assign(N, 0.f); bufV.assign(N, 0.f); bufADV.assign(N, 0.f);
    for (Uint t = 0; t < N; ++t) {
      for (Uint k = 0; k < dS && k < EP.states[t].size(); ++k) bufS[(size_t) t * dS + k] = EP.states[t][k];
      if (t < EP.actions.size())  for (Uint k = 0; k < dA && k < EP.actions[t].size(); ++k)  bufA[(size_t) t * dA + k]  = (float) EP.actions[t][k];
      if (t < EP.policies.size()) for (Uint k = 0; k < dP && k < EP.policies[t].size(); ++k) bufMU[(size_t) t * dP + k] = (float) EP.policies[t][k];
      bufR[t] = (float) EP.rewards[t];
      if (t < EP.stateValue.size())      bufV[t]   = EP.stateValue[t];
      if (t < EP.actionAdvantage.size()) bufADV[t] = EP.actionAdvantage[t];
    }
    const int64_t tag = (int64_t) std::max<Sint>(EP.ID, 0) * (1 << 20) + (nextSeq++ & ((1 << 20) - 1));
    byTag[tag] = const_cast<Episode*>(&EP); tagOf[&EP] = tag;
    if (restoredValues) {   // an episode read from a checkpoint: keep its return estimates, TD errors, importance weights
      std::vector<float> q(EP.returnEstimator.begin(), EP.returnEstimator.end()), d(EP.deltaValue.begin(), EP.deltaValue.end()),
                         w(EP.offPolicImpW.begin(), EP.offPolicImpW.end()), kl(EP.KullbLeibDiv.begin(), EP.KullbLeibDiv.end());
      q.resize(N, 0.f); d.resize(N, 0.f); w.resize(N, 0.f); kl.resize(N, 0.f);
      check(smb200_push_episode_restored(gpu, tag, (int32_t) N, EP.bReachedTermState ? 1 : 0, bufS.data(), bufA.data(), bufMU.data(),
                                         bufR.data(), bufV.data(), bufADV.data(), q.data(), d.data(), w.data(), kl.data(),
                                         (double) data->CmaxRet), "push_episode_restored");
    } else
    check(smb200_push_episode(gpu, tag, (int32_t) N, EP.bReachedTermState ? 1 : 0, bufS.data(), bufA.data(),
                              bufMU.data(), bufR.data(), bufV.data(), bufADV.data()), "push_episode");
    ++nPushed;
  }

  // new episodes -> HBM; then the first-in-first-out pruning of the host copy
  // (MemoryProcessing::applyEpisodesRemovalAlgo with ERoldSeqFilter "oldest")
  void mirrorEpisodes()
  {
    std::lock_guard<std::mutex> lock(data->dataset_mutex);
    for (const auto& e : data->episodes)
      if (mirrored.insert(e.get()).second) pushToDevice(*e);
    // the device's total order: tag = ID * 2^20 + arrival number, newest first.  Episode::ID alone has ties (every episode
    // collected before training has ID 0) and std::sort is not stable: both copies must evict the same episode.
    std::sort(data->episodes.begin(), data->episodes.end(),
              [this](const std::unique_ptr<Episode>& a, const std::unique_ptr<Episode>& b) {
                const auto ta = tagOf.find(a.get()), tb = tagOf.find(b.get());
                const int64_t xa = ta == tagOf.end() ? (int64_t) std::max<Sint>(a->ID, 0) * (1 << 20) : ta->second;
                const int64_t xb = tb == tagOf.end() ? (int64_t) std::max<Sint>(b->ID, 0) * (1 << 20) : tb->second;
                return xa > xb; });
    const long maxTotObs = settings.maxTotObsNum_local;
    while (data->episodes.size() > 1 && data->nStoredSteps() - (long) data->episodes.back()->nsteps() > maxTotObs) {
      const Episode* gone = data->episodes.back().get();
      mirrored.erase(gone);
      const auto it = tagOf.find(gone);
      if (it != tagOf.end()) { byTag.erase(it->second); tagOf.erase(it); }
      data->removeBackEpisode();
      data->stats.nPrunedEps++;
    }
  }

  void pullScaling()
  {
    const Uint dS = MDP.dimStateObserved;
    std::vector<float> mean(dS), scale(dS), stdev(dS); float rew[3];
    check(smb200_get_scaling(gpu, mean.data(), scale.data(), stdev.data(), rew), "get_scaling");
    for (Uint k = 0; k < dS; ++k) { MDP.stateMean[k] = mean[k]; MDP.stateScale[k] = scale[k]; MDP.stateStdDev[k] = stdev[k]; }
    MDP.rewardsMean = rew[0]; MDP.rewardsScale = rew[1]; MDP.rewardsStdDev = rew[2];
  }

  void pullWeights()
  {
    check(smb200_get_weights(gpu, wblob.data(), (int64_t) wblob.size()), "get_weights");
    std::copy(wblob.begin(), wblob.end(), hostWeights()->params);   // read by the actors' forward passes
  }

  void publishStats()
  {
    data->beta = last.beta; data->CmaxRet = last.cmax; data->CinvRet = last.cinv;
    ReplayStats& s = data->stats;
    s.nFarPolicySteps = (Uint) last.n_far_policy; s.avgKLdivergence = last.avg_kl; s.avgSquaredErr = last.avg_sq_err;
    s.maxAbsError = last.max_abs_err; s.avgReturn = last.avg_return; s.stdevQ = last.stdev_q; s.avgQ = last.avg_q;
    s.maxQ = last.max_q; s.minQ = last.min_q;
    s.countReturnsEstimateUpdates = (Sint) last.cnt_ret; s.sumReturnsEstimateErrors = last.sum_ret_err;
  }

  static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

 public:
  RACER_B200(MDPdescriptor& M, HyperParameters& S, ExecutionInfo& D, const bool racer) : Base(M, S, D), isRacer(racer)
  {
    if (D.world_rank == 0) printf("smarties_b200: learner steps of this agent run on the GPU (libsmarties_b200.so)\n");
    if (const char* m = std::getenv("SMARTIES_B200_MAXSTEPS")) maxStepsPerCall = std::max(1, std::min(256, std::atoi(m)));
    if (const char* m = std::getenv("SMARTIES_B200_ACTORS")) deviceActors = std::atoi(m) != 0;
    if (deviceActors && D.world_rank == 0) printf("smarties_b200: the actors' policy evaluations run on the GPU as well (smb200_forward_seq)\n");
    stepStats.resize(maxStepsPerCall);
    createDeviceLearner();
  }
  ~RACER_B200() override
  {
    if (gpu && distrib.world_rank == 0)
      printf("smarties_b200: %ld gradient steps, %ld episodes mirrored; seconds in push %.3f, device steps %.3f, weight sync %.3f; "
             "%.3f s of wall clock since training started\n",
             (long) data->nGradSteps(), nPushed, secPush, secStep, secSync, tTrainStart > 0 ? now() - tTrainStart : 0.0);
    if (gpu && deviceActors && distrib.world_rank == 0)
      printf("smarties_b200: %ld policy evaluations in %ld device calls (largest call %ld agents)\n", nFwdAgents, nFwdCalls, nFwdMax);
    if (gpu) smb200_pin_host_buffer(hostWeights()->params, 0);
    smb200_destroy(gpu);
  }

  // RACER::selectAction (Learners/RACER.cpp:30-47) with the network outputs from the device
  void selectAction(const MiniBatch& MB, Agent& agent) override
  {
    if (!deviceActors) { Base::selectAction(MB, agent); return; }
    const Rvec output = deviceForward(MB);
    const Policy_t pol(pol_start, aInfo, output);
    auto action = pol.selectAction(agent, distrib.bTrain);
    const Advantage_t adv(adv_start, aInfo, output, &pol);
    const Real V = net2V(output[VsID]);
    MB.appendValues(V, V + adv.computeAdvantage(action));
    agent.setAction(action, pol.getVector());
  }

  // RACER::processTerminal (Learners/RACER.cpp:49-59)
  void processTerminal(const MiniBatch& MB, Agent& agent) override
  {
    if (!deviceActors) { Base::processTerminal(MB, agent); return; }
    if (agent.agentStatus == LAST) MB.appendValues(net2V(deviceForward(MB)[VsID]));   // truncated: not a terminal state
    else MB.appendValues(0);
  }

  // Learner_approximator::save (Learner_approximator.cpp:133-142): the reference's own writers produce the files;
  // what lives on the device (Adam moments, per-transition values of the replay) is copied into the host objects first.
  void save() override
  {
    AdamOptimizer* const adam = dynamic_cast<AdamOptimizer*>(networks[0]->opt.get());
    Parameters* M1 = adam ? adam->_1stMom.get() : nullptr, * M2 = adam ? adam->_2ndMom.get() : nullptr;
    if (M1 && M2) check(smb200_get_adam(gpu, M1->params, M2->params, (int64_t) M1->nParams), "get_adam");
    if (settings.targetDelay > 0) {   // AdamOptimizer::target_weights are maintained on the device (Optimizer.cpp:162-177)
      Parameters* T = networks[0]->opt->target_weights.get();
      check(smb200_get_target_weights(gpu, T->params, (int64_t) T->nParams), "get_target_weights");
    }
    {
      std::lock_guard<std::mutex> lock(data->dataset_mutex);
      const int64_t nEp = smb200_n_episodes(gpu), nRows = smb200_n_rows(gpu);
      std::vector<int64_t> ids(nEp), rows(nEp);
      check(smb200_read_episodes(gpu, ids.data(), rows.data(), nullptr, nEp), "read_episodes");
      std::vector<float> F[6];
      const int fid[6] = {SMB200_F_V, SMB200_F_ADV, SMB200_F_QRET, SMB200_F_DELTA, SMB200_F_RHO, SMB200_F_KL};
      for (int k = 0; k < 6; ++k) { F[k].resize(nRows); check(smb200_read_field(gpu, fid[k], F[k].data(), nRows), "read_field"); }
      int64_t o = 0;
      for (int64_t e = 0; e < nEp; ++e) {
        const auto it = byTag.find(ids[e]);
        const int64_t N = rows[e];
        if (it != byTag.end() && (int64_t) it->second->nsteps() == N) {
          Episode& EP = * it->second;
          EP.stateValue.assign(F[0].begin() + o, F[0].begin() + o + N);
          EP.actionAdvantage.assign(F[1].begin() + o, F[1].begin() + o + N);
          EP.returnEstimator.assign(F[2].begin() + o, F[2].begin() + o + N);
          EP.deltaValue.assign(F[3].begin() + o, F[3].begin() + o + N);
          EP.offPolicImpW.assign(F[4].begin() + o, F[4].begin() + o + N);
          EP.KullbLeibDiv.assign(F[5].begin() + o, F[5].begin() + o + N);
        }
        o += N;
      }
    }
    Base::save();
  }

  // Learner_approximator::restart (Learner_approximator.cpp:118-131): the reference's own readers fill the host network,
  // optimiser, MDP scaling, counters and episodes; the device learner is then loaded from those objects.
  void restart() override
  {
    Base::restart();
    Parameters* W = hostWeights();
    check(smb200_set_weights(gpu, W->params, W->nParams), "set_weights");
    if (data->nGradSteps() <= 0 && data->nStoredEps() == 0) return;      // nothing but (maybe) a policy was found
    AdamOptimizer* const adam = dynamic_cast<AdamOptimizer*>(networks[0]->opt.get());
    Parameters* M1 = adam ? adam->_1stMom.get() : nullptr, * M2 = adam ? adam->_2ndMom.get() : nullptr;
    if (M1 && M2) check(smb200_set_adam(gpu, M1->params, M2->params, (int64_t) M1->nParams, data->nGradSteps()), "set_adam");
    if (settings.targetDelay > 0) {
      Parameters* T = networks[0]->opt->target_weights.get();
      check(smb200_set_target_weights(gpu, T->params, (int64_t) T->nParams), "set_target_weights");
    }
    const Uint dS = MDP.dimStateObserved;
    std::vector<float> mean(MDP.stateMean.begin(), MDP.stateMean.end()), scale(MDP.stateScale.begin(), MDP.stateScale.end()),
                       stdev(MDP.stateStdDev.begin(), MDP.stateStdDev.end());
    mean.resize(dS, 0.f); scale.resize(dS, 1.f); stdev.resize(dS, 1.f);
    const float rew[3] = {(float) MDP.rewardsMean, (float) MDP.rewardsScale, (float) MDP.rewardsStdDev};
    check(smb200_set_scaling(gpu, mean.data(), scale.data(), stdev.data(), rew), "set_scaling");
    {
      std::lock_guard<std::mutex> lock(data->dataset_mutex);
      for (const auto& e : data->episodes)
        if (mirrored.insert(e.get()).second) pushToDevice(*e, true);
    }
    check(smb200_set_grad_step(gpu, data->nGradSteps()), "set_grad_step");
    check(smb200_set_refer(gpu, data->beta, data->CmaxRet), "set_refer");
    if (distrib.world_rank == 0)
      printf("smarties_b200: restarted the device learner at gradient step %ld with %ld episodes (%ld transitions)\n",
             (long) data->nGradSteps(), (long) smb200_n_episodes(gpu), (long) smb200_n_transitions(gpu));
  }

  void setupTasks(TaskQueue& tasks) override
  {
    if (not bTrain) return;
    algoSubStepID = -1;

    auto stepInit = [&]()      // Learner::initializeLearner (Learner.cpp:47-72)
    {
      if (algoSubStepID >= 0) return;
      if (data->nStoredSteps() < nObsB4StartTraining) return;
      mirrorEpisodes();
      check(smb200_initialize_learner(gpu), "initialize_learner");
      pullScaling();
      data->counters.nGatheredB4Startup = nObsB4StartTraining;
      tTrainStart = now();
      algoSubStepID = 0;
    };
    tasks.add(stepInit);

    auto stepMain = [&]()      // spawnTrainTasks + processMemoryBuffer + applyGradient (RACER.cpp:81-109)
    {
      if (algoSubStepID not_eq 0) return;
      if (this->blockGradientUpdates()) return;
      // The reference runs one step per pass of the task loop; when the actors are ahead by k
      // observations it runs k passes back to back (Learner::blockGradientUpdates, Learner.cpp:119-126).
      // Those k steps go to the device as ONE call (one persistent launch); SMARTIES_B200_MAXSTEPS=1
      // restores the strict per-step hand-off.
      const long ahead = (long) std::floor(this->nLocTimeStepsTrain() / std::max((Real) 1e-9, this->obsPerStep_loc)) - this->nGradSteps();
      const long toSweep = 1000 - (this->nGradSteps() + 1) % 1000;          // stop at the every-1000-steps boundary
      // a checkpoint (Learner::logStats: `currStep % saveFreq == 0`) must see the device state of ITS step: end the call there
      const long sf = (long) std::max<Uint>(1, settings.saveFreq);
      const long toSave = (sf - (this->nGradSteps() + 1) % sf) % sf;
      const int k = (int) std::max<long>(1, std::min<long>({(long) maxStepsPerCall, ahead, toSweep + 1, toSave + 1}));
      const double t0 = now();
      nameGradStats();
      mirrorEpisodes();
      const double t1 = now();
      // the new weights land in the host network (page-locked once, below) behind the last step of the same call
      check(smb200_train_steps_weights(gpu, k, stepStats.data(), hostWeights()->params, (int64_t) hostWeights()->nParams), "train_steps_weights");
      const double t2 = now();
      const double t3 = t2;
      secPush += t1 - t0; secStep += t2 - t1; secSync += t3 - t2;
      for (int i = 0; i < k; ++i) {
        last = stepStats[i];
        if ((this->nGradSteps() + 1) % 1000 == 0) pullScaling();   // the every-1000-steps moment sweep moved the normalisers
        publishStats();
        this->logStats();
        this->globalGradCounterUpdate();
      }
    };
    tasks.add(stepMain);
  }
};

static std::ifstream openSettings(ExecutionInfo& D, const Uint ID)   // same lookup order as the reference factory
{
  char cwd[512];
  if (!getcwd(cwd, 512)) cwd[0] = 0;
  if (chdir(D.initial_runDir)) {}
  char name[256];
  snprintf(name, 256, "settings_%02u.json", (unsigned) ID);
  std::ifstream ret(name, std::ifstream::in);
  if (!ret.is_open()) ret.open("settings.json", std::ifstream::in);
  if (chdir(cwd)) {}
  return ret;
}

std::unique_ptr<Learner> createLearner(const Uint learnerID, MDPdescriptor& MDP, ExecutionInfo& distrib)
{
  if (!std::getenv("SMARTIES_B200")) return createLearner_reference(learnerID, MDP, distrib);

  HyperParameters settings(MDP.dimObs(), MDP.dimAct());
  std::ifstream ifs = openSettings(distrib, learnerID);
  settings.initializeOpts(ifs, distrib);
  // "V-RACER makes little sense with discrete action-spaces": the reference's factory overrides the user (AlgoFactory.cpp:78-83)
  if (settings.learner == "VRACER" && MDP.bDiscreteActions()) settings.learner = "RACER";
  const bool covered =
      (settings.learner == "VRACER" || settings.learner == "RACER") &&
      // discrete actions: RACER<Discrete_advantage, Discrete_policy, Uint> (AlgoFactory.cpp:100-113), one component, feed-forward net
      (!MDP.bDiscreteActions() || (settings.learner == "RACER" && MDP.dimAction == 1 && MDP.discreteActionValues[0] >= 2 &&
                                   MDP.discreteActionValues[0] <= 64 && settings.nnType == "FFNN" && !MDP.isPartiallyObservable)) &&
      // prioritized samplers run on the device learner (one launch per step); the non-FIFO episode filters do too, but this binding
      // prunes its host copy of the episodes first-in-first-out (mirrorEpisodes), so they stay with the reference learner here
      (settings.dataSamplingAlgo == "uniform" || settings.dataSamplingAlgo == "PERrank" || settings.dataSamplingAlgo == "PERerr" ||
       settings.dataSamplingAlgo == "PERseq") && (settings.returnsEstimator == "default" || settings.returnsEstimator == "retrace" || settings.returnsEstimator == "GAE" ||
                                                  settings.returnsEstimator == "retraceExplore") &&
      (settings.ERoldSeqFilter == "oldest" || settings.ERoldSeqFilter == "default") &&
      (settings.nnType == "FFNN" || settings.nnType == "LSTM" || settings.nnType == "MGU" || settings.nnType == "GRU") &&
      // hidden-layer functions the device evaluates (makeFunction, Functions.h:643-668); recurrent cells keep Tanh.
      // settings/default.json asks for SoftSign.
      (settings.nnFunc == "Tanh" || (settings.nnType == "FFNN" && !MDP.isPartiallyObservable &&
                                     (settings.nnFunc == "SoftSign" || settings.nnFunc == "HardSign" || settings.nnFunc == "Sigm" ||
                                      settings.nnFunc == "Relu" || settings.nnFunc == "LRelu" || settings.nnFunc == "ExpPlus" ||
                                      settings.nnFunc == "SoftPlus" || settings.nnFunc == "Exp" || settings.nnFunc == "Linear"))) &&
      settings.nnOutputFunc == "Linear" &&
      // several learner ranks: the device learners of the ranks would have to exchange CUDA-IPC handles over
      // distrib.learners_train_comm (smb200_comm_init / smb200_comm_attach) — not wired into the binding: reference learner
      MPICommSize(distrib.learners_train_comm) == 1 &&
      settings.ESpopSize == 1 && MDP.nAppendedObs == 0 && MDP.conv2dDescriptors.size() == 0 &&
      // encoder layers are the first layers of the one network; in a partially observable MDP the reference gives them another
      // cell type than the layers after them ("RNN" vs "MGU", Approximator.cpp:219-223,265-267): reference learner
      (!MDP.isPartiallyObservable || settings.bRecurrent ||
       std::all_of(settings.encoderLayerSizes.begin(), settings.encoderLayerSizes.end(), [](Uint n) { return n == 0; }));
  if (!covered) {
    warn("SMARTIES_B200 is set but these settings are outside the device path: using the reference CPU learner.");
    return createLearner_reference(learnerID, MDP, distrib);
  }
  if (settings.returnsEstimator == "default") settings.returnsEstimator = "retrace";   // AlgoFactory.cpp:134-136
  const ActionInfo aInfo = ActionInfo(MDP);
  std::unique_ptr<Learner> ret;
  std::ostringstream o;
  o << MDP.dimState << " ";
  if (settings.learner == "RACER" && MDP.bDiscreteActions()) {
    using R = RACER<Discrete_advantage, Discrete_policy, Uint>;
    MDP.policyVecDim = R::getnDimPolicy(aInfo);
    ret = std::make_unique<RACER_B200<Discrete_advantage, Discrete_policy, Uint>>(MDP, settings, distrib, true);
  } else if (settings.learner == "RACER") {
    using R = RACER<Param_advantage, Continuous_policy, Rvec>;
    MDP.policyVecDim = R::getnDimPolicy(aInfo);
    ret = std::make_unique<RACER_B200<Param_advantage>>(MDP, settings, distrib, true);
  } else {
    using R = RACER<Zero_advantage, Continuous_policy, Rvec>;
    MDP.policyVecDim = R::getnDimPolicy(aInfo);
    ret = std::make_unique<RACER_B200<Zero_advantage>>(MDP, settings, distrib, false);
  }
  o << MDP.dimAction << " " << MDP.policyVecDim;
  if (distrib.world_rank == 0) { std::ofstream fout("problem_size.log", std::ios::app); fout << o.str() << std::endl; }
  char lName[256];
  snprintf(lName, 256, "agent_%02u", (unsigned) learnerID);
  ret->setLearnerName(std::string(lName), learnerID);
  return ret;
}

}  // namespace smarties
