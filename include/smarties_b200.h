/*
 * smarties_b200.h — C-ABI of the B200-native V-RACER / RACER learner hot path.
 *
 * This is the drop-in boundary for ONE path of cselab/smarties: the off-policy learner step
 *   sample -> gather -> network forward -> ReF-ER/Retrace loss -> backward -> Adam -> replay stats
 * i.e. what `smarties::Learner_approximator::spawnTrainTasks`, `Learner::processMemoryBuffer`,
 * `Learner_approximator::applyGradient` and the `MemoryBuffer`/`MemoryProcessing` functions they
 * call compute on host cores in the reference.  A `smarties::Learner` subclass (see
 * INTEGRATION.md, "RACER_B200") forwards to these entry points; everything above it
 * (Engine, Communicator, Worker, Master, sockets/MPI) stays the reference's.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative smb200_status and never throws; the caller owns host buffers, the library owns
 * device memory behind the opaque handle; calls on one handle must be serialised by the
 * caller.  Reference file:line citations are relative to /root/reference/source/smarties/.
 */
#ifndef SMARTIES_B200_H
#define SMARTIES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMB200_MAX_HIDDEN 8
#define SMB200_MAX_ACTION 64

typedef struct smb200_learner smb200_learner; /* opaque */

enum smb200_status {
  SMB200_OK = 0,
  SMB200_ERR_INVALID = -1,   /* bad argument / unsupported configuration */
  SMB200_ERR_CUDA = -2,      /* CUDA runtime error (message via smb200_last_error) */
  SMB200_ERR_CAPACITY = -3,  /* replay ring / episode table full */
  SMB200_ERR_STATE = -4      /* call not valid in the current state */
};

enum smb200_algo { SMB200_VRACER = 0, SMB200_RACER = 1 };
enum smb200_nn_type { SMB200_FFNN = 0, SMB200_LSTM = 1,
                      SMB200_MGU = 2 /* "MGU" and "GRU" build the same MGULayer (Builder.cpp:67-72, Layers/Layer_GRU.h); it is also
                                        what a partially observable MDP gets for a feed-forward request (Approximator.cpp:219-223) */ };
/* "returnsEstimator" (createReturnEstimator, ReplayMemory/MemoryProcessing.cpp:419-450): Retrace (:391-400; also what
 * "default" means for RACER / V-RACER, Learners/AlgoFactory.cpp:134-136) or GAE (:411-417). */
enum smb200_returns_estimator { SMB200_RETRACE = 0, SMB200_GAE = 1,
                                SMB200_RETRACE_EXPLORE = 2 /* computeRetraceExplBonus, MemoryProcessing.cpp:402-409 */ };
/* "dataSamplingAlgo" (Sampling::prepareSampler, ReplayMemory/Sampling.cpp:298-335): uniform, or prioritized by the rank of the
 * squared TD error (TSample_impRank, :101-166), by the TD error (TSample_impErr, :169-226), by the episode's mean squared error
 * (Sample_impSeq, :230-296).  "ERoldSeqFilter" (getERfilterAlgo, MemoryProcessing.cpp:261-298): which episodes leave a full
 * buffer — the oldest, or the ones with the largest far-policy fraction / the largest KL divergence / the smallest error.
 * Everything but uniform + oldest makes every learner step a separate launch with one host round trip (the sampler and the
 * episode order of step k+1 depend on values step k wrote), like the reference re-prepares its sampler every step. */
enum smb200_sampling { SMB200_SAMPLE_UNIFORM = 0, SMB200_SAMPLE_PER_RANK = 1, SMB200_SAMPLE_PER_ERR = 2, SMB200_SAMPLE_PER_SEQ = 3 };
enum smb200_er_filter { SMB200_FILTER_OLDEST = 0, SMB200_FILTER_FARPOLFRAC = 1, SMB200_FILTER_MAXKLDIV = 2, SMB200_FILTER_MINERROR = 3 };
/* "nnFunc": function of the hidden dense layers (makeFunction, Network/Layers/Functions.h:643-668) with its own weight
 * initialisation factor (Function::initFactor); feed-forward nets (recurrent cells keep Tanh). */
enum smb200_nn_func { SMB200_TANH = 0, SMB200_SOFTSIGN = 1 /* settings/default.json */, SMB200_HARDSIGN = 2, SMB200_SIGM = 3,
                      SMB200_RELU = 4, SMB200_LRELU = 5, SMB200_EXPPLUS = 6, SMB200_SOFTPLUS = 7, SMB200_EXP = 8, SMB200_LINEAR = 9 };
enum smb200_field {          /* per-transition replay arrays, Episode.h:66-75 */
  SMB200_F_V = 0, SMB200_F_ADV = 1, SMB200_F_QRET = 2, SMB200_F_DELTA = 3, SMB200_F_RHO = 4, SMB200_F_KL = 5,
  SMB200_F_REWARD = 6
};

/* Everything the reference reads from settings.json (Settings/HyperParameters.h:37-73), from
 * the MDP descriptor (Core/StateAction.h:46-125) and from Bund.h that shapes the arithmetic.
 * Zero-initialise, then call smb200_default_config() and override. */
typedef struct smb200_config {
  int32_t device;                       /* CUDA device ordinal */
  int32_t algo;                         /* smb200_algo: "learner": "VRACER" | "RACER" */
  int32_t dim_state;                    /* MDP.dimStateObserved */
  int32_t dim_action;                   /* MDP.dimAction (continuous) */
  uint8_t action_bounded[SMB200_MAX_ACTION]; /* MDP.bActionSpaceBounded -> SquashedNormalPolicy */
  int32_t n_hidden;                     /* nnLayerSizes.size() */
  int32_t hidden[SMB200_MAX_HIDDEN];    /* nnLayerSizes */
  int32_t batch_size;                   /* batchSize_local */
  int32_t batch_size_global;            /* batchSize (== local for one learner rank) */
  int64_t max_tot_obs;                  /* maxTotObsNum_local */
  int64_t max_tot_obs_global;           /* maxTotObsNum */
  int64_t capacity_rows;                /* rows of HBM ring to allocate; 0 = derive from max_tot_obs */
  int32_t max_episodes;                 /* episode-table slots; 0 = derive */
  double gamma, lambda, clip_imp_weight, penal_tol, eps_anneal;
  double learnrate, nn_lambda, expl_noise, out_weights_prefac;
  int32_t refer_reduce_threads;         /* how many OpenMP threads' reduction order the far-policy
                                           count emulates (MemoryProcessing.cpp:202-227); 0 = 32 */
  int32_t world_rank, world_size;       /* learner ranks sharing the gradient (nMasters) */
  uint64_t seed;                        /* ExecutionInfo::randSeed (sampler + weight init) */
  int32_t nn_type;                      /* smb200_nn_type: "nnType": "FFNN" | "LSTM" (Layers/Layer_LSTM.h) | "MGU" / "GRU" (Layers/Layer_GRU.h) */
  int32_t nn_bptt_seq;                  /* "nnBPTTseq": recurrent window = min(nnBPTTseq, t) past steps
                                           (ReplayMemory/MemoryBuffer.cpp:393-402) */
  int64_t min_tot_obs;                  /* minTotObsNum_local = nObsB4StartTraining: only recorded in checkpoints
                                           ("nInitialData", MemoryBuffer.cpp:300); 0 = max_tot_obs */
  int32_t returns_estimator;            /* smb200_returns_estimator: "returnsEstimator": "retrace" | "GAE" | ("retraceExplore") */
  int32_t discrete_options;             /* 0: continuous actions.  K > 0: one discrete action with K options
                                           (ActionInfo::dimDiscrete, Core/StateAction.h; RACER<Discrete_advantage, Discrete_policy, Uint>,
                                           Math/Discrete_policy.h, Discrete_advantage.h): dim_action = 1, the stored action is the
                                           option label (+0.1, StateAction.h:320-341), the behaviour policy has K columns, the net
                                           outputs [V | advantages(K) | policy(K)] and has no ParamLayer.  Feed-forward nets. */
  int32_t data_sampling;                /* smb200_sampling: "dataSamplingAlgo" */
  int32_t er_filter;                    /* smb200_er_filter: "ERoldSeqFilter" */
  int32_t nn_func;                      /* smb200_nn_func: "nnFunc" (default Tanh, HyperParameters.h:72) */
  double target_delay;                  /* "targetDelay" (AdamOptimizer::tgtUpdateAlpha, Network/Optimizer.cpp:162-177): 0 = the target weights
                                           stay what they were at construction / restart; in (0, 1): exponential average target += a (w - target)
                                           after every update; >= 1: copy of the weights every floor(targetDelay) updates.  RACER never evaluates
                                           the target network: the weights only reach the checkpoint (<name>_net_tgt_weights.raw). */
} smb200_config;

/* Per-step scalars the reference prints / feeds back (MemoryBuffer::getMetrics,
 * MemoryBuffer.cpp:522-575; MemoryProcessing.cpp:46-92,187-259), valid after the step. */
typedef struct smb200_step_stats {
  double beta, cmax, cinv;
  int64_t n_far_policy;       /* reference-formula far-policy count (drives beta) */
  int64_t n_far_exact;        /* exact integer count of per-transition far flags */
  double avg_kl, avg_sq_err, max_abs_err, avg_return, stdev_q, avg_q, max_q, min_q;
  double sum_ret_err; int64_t cnt_ret;
  int64_t grad_step;          /* nGradSteps after the step */
} smb200_step_stats;

/* HyperParameters defaults for (dim_state, dim_action) — Settings/HyperParameters.h:37-73. */
int smb200_default_config(smb200_config* cfg, int32_t dim_state, int32_t dim_action);

/* Learner construction: RACER ctor + setupNet (Learners/RACER_common.cpp:70-115,
 * Network/Builder.cpp:119-170) incl. weight initialisation from mt19937(seed). */
int smb200_create(const smb200_config* cfg, smb200_learner** out);
void smb200_destroy(smb200_learner* h);
const char* smb200_last_error(void);

int64_t smb200_n_params(const smb200_learner* h);   /* padded blob size, Parameters.h:159-176 */
int32_t smb200_n_outputs(const smb200_learner* h);  /* 1 + 2*dA (V-RACER) */

/* Weights / Adam moments as the reference's padded parameter blob (Parameters.h:28-177). */
int smb200_set_weights(smb200_learner* h, const float* blob, int64_t n);
int smb200_get_weights(smb200_learner* h, float* blob, int64_t n);
/* AdamOptimizer::target_weights (padded blob like the weights). */
int smb200_get_target_weights(smb200_learner* h, float* blob, int64_t n);
int smb200_set_target_weights(smb200_learner* h, const float* blob, int64_t n);
int smb200_set_adam(smb200_learner* h, const float* m1, const float* m2, int64_t n, int64_t n_step);
int smb200_get_adam(smb200_learner* h, float* m1, float* m2, int64_t n);
/* Last summed parameter gradient (AdamOptimizer::gradSum before apply_update, Optimizer.cpp:110-120). */
int smb200_get_grad(smb200_learner* h, float* blob, int64_t n);

/* State / reward normalisers (MDPdescriptor, Core/StateAction.h:56-58;
 * agent_XX_scaling.raw order of MemoryBuffer.cpp:277-287). rewards = {mean, scale, stdev}. */
int smb200_set_scaling(smb200_learner* h, const float* mean, const float* scale, const float* stdev, const float rewards[3]);
int smb200_get_scaling(smb200_learner* h, float* mean, float* scale, float* stdev, float rewards[3]);

/* MemoryBuffer::pushBackEpisode + Episode::finalize + computeReturnEstimator
 * (MemoryBuffer.cpp:133-170,479-520; Episode.cpp:244-273).  n_rows = nsteps() including the
 * terminal/truncated row.  value/advantage may be NULL (zeros).  Thread-safe against the train / forward / weight
 * calls of the same learner (one lock per learner serialises them: actor threads may push finished episodes and
 * evaluate the policy while the learner thread trains — MemoryBuffer::dataset_mutex of the reference). */
int smb200_push_episode(smb200_learner* h, int64_t id, int32_t n_rows, int32_t terminated,
                        const float* states, const float* actions, const float* policies,
                        const float* rewards, const float* value, const float* advantage);
int64_t smb200_n_transitions(const smb200_learner* h);  /* MemoryBuffer::nStoredSteps */
int64_t smb200_n_episodes(const smb200_learner* h);     /* MemoryBuffer::nStoredEps */

/* Learner::initializeLearner (Learners/Learner.cpp:47-72). */
int smb200_initialize_learner(smb200_learner* h);
/* counters.nGradSteps / AdamOptimizer::nStep after a restart (Approximator.h:64). */
int smb200_set_grad_step(smb200_learner* h, int64_t n_grad_steps);
int smb200_seed_sampler(smb200_learner* h, uint64_t seed);

/* Sample_uniform::sample + Sampling::IDtoSeqStep (ReplayMemory/Sampling.cpp:26-47,82-96):
 * B unique ascending transition ids from the handle's std::mt19937 -> (episode position in the
 * buffer's current order, time step).  Advances the generator exactly like the reference. */
int smb200_sample(smb200_learner* h, int64_t* episode_pos, int64_t* tstep);

/* n learner steps: {spawnTrainTasks; processMemoryBuffer; applyGradient; globalGradCounterUpdate}
 * (Learners/RACER.cpp:81-109) with the internal sampler.  stats (may be NULL) receives n entries. */
int smb200_train_steps(smb200_learner* h, int32_t n, smb200_step_stats* stats);
/* The same, plus the padded parameter blob after the last step (the layout of smb200_get_weights) in the same call: the copy
 * shares the call's stream synchronisation.  This is how the binding refreshes the host network its actors evaluate
 * (RACER::selectAction, Learners/RACER.cpp:30-47) without a second device round trip per learner call. */
int smb200_train_steps_weights(smb200_learner* h, int32_t n, smb200_step_stats* stats, float* weights, int64_t n_weights);
/* cudaHostRegister / cudaHostUnregister (bytes = 0) of a caller-owned buffer that the calls above copy into. */
int smb200_pin_host_buffer(void* ptr, int64_t bytes);
/* One step on caller-supplied samples (episode position, time step), e.g. the output of
 * smb200_sample or of the reference's own sampler. */
int smb200_train_step_on(smb200_learner* h, const int64_t* episode_pos, const int64_t* tstep,
                         int32_t batch, smb200_step_stats* stats);

/* Diagnostics of the last step, batch-major: net outputs O[B][nOut] (f32), output gradient
 * g[B][nOut] (f32), standardized inputs X[B][dS].  NULL pointers are skipped. */
int smb200_get_last_batch(smb200_learner* h, float* outputs, float* out_grad, float* inputs);

/* Output-gradient statistics file of the reference: StatsTracker::track_vector / reduce_stats / printToFile
 * (Utils/StatsTracker.cpp:28-107, called from Approximator::setGradient, Network/Approximator.h:197, and
 * Learner_approximator::spawnTrainTasks, Learners/Learner_approximator.cpp:89).  Once `base` is set, every learner step
 * that starts with nGradSteps % 1000 == 0 appends the mean and root-mean-square over the mini-batch of each network
 * output's gradient (2*nOut floats) to "<base>_outGrad_stats.raw" (base = "<learner_name>_<net name>", e.g.
 * "agent_00_net"); the file begins with the float nOut + 0.1 when that step is the learner's first.  Rank 0 writes,
 * from its own samples, like the reference.  NULL or "" switches it off (the default).  Not written by the
 * benchmark-only smb200_train_presampled path. */
int smb200_set_grad_stats(smb200_learner* h, const char* base);

/* Stand-alone sweeps (also run internally every 1000 steps):
 * updateReturnEstimator over all episodes (MemoryProcessing.cpp:23-44,452-481) ... */
int smb200_retrace_sweep(smb200_learner* h, double* sum_err2);
/* ... and the reward/state moments of updateRewardsStats (MemoryProcessing.cpp:94-185):
 * out[2*dS+3] = {sum(s-mean)[dS], sum((s-mean)^2)[dS], count, sum(r-mean), sum((r-mean)^2)}. */
int smb200_reward_state_moments(smb200_learner* h, double* out);
/* Both of them, plus the exact recompute of the per-episode aggregates (Episode::updateCumulative, Episode.cpp:213-242), as
 * the ONE pass over the buffer the learner runs every 1000 steps (MemoryProcessing.cpp:187-259 followed by :94-185; both read
 * the normalisers of before the pass): k_sweep_fused, state rows streamed through shared memory by cp.async.bulk.
 * moments[2*dS+3] as above.  Needs dim_state % 4 == 0 with dim_state / 4 a power of two, estimator retrace or GAE. */
int smb200_fused_sweep(smb200_learner* h, double* sum_err2, double* moments);

/* Read one per-transition array for every stored episode, concatenated in the buffer's current
 * episode order (rows incl. the terminal row), and the episode table in the same order:
 * ids[nEp], n_rows[nEp], aggregates[nEp][9] = {avgKL, fracFar, avgSqErr, maxAbsErr, sumQ2,
 * sumQ, maxQ, minQ, totR} (Episode.h:77-81). */
int smb200_read_field(smb200_learner* h, int32_t field, float* out, int64_t n);
/* Inverse of smb200_read_field (restores per-transition arrays, e.g. from a reference checkpoint). */
int smb200_write_field(smb200_learner* h, int32_t field, const float* in, int64_t n);
int smb200_read_episodes(smb200_learner* h, int64_t* ids, int64_t* n_rows, float* aggregates, int64_t n_ep);
int64_t smb200_n_rows(const smb200_learner* h);
int smb200_get_stats(smb200_learner* h, smb200_step_stats* out);

/* Checkpoints in the reference's own file formats (byte-compatible both ways; fully observed states):
 * Learner_approximator::save / restart (Learners/Learner_approximator.cpp:118-142) =
 *   <base>_net_{weights,tgt_weights,1stMom,2ndMom}.raw   Network::save, padding stripped (Network/Network.cpp:22-67)
 *   <base>_scaling.raw                                   3*dS + 3 doubles (ReplayMemory/MemoryBuffer.cpp:277-287)
 *   <base>_rank_XXX_learner_status.raw / _data.raw       counters + packed episodes (MemoryBuffer.cpp:289-324,
 *                                                        Episode::packEpisode, ReplayMemory/Episode.cpp:24-93)
 * base = "<dir>/agent_00".  smb200_save also writes the *_backup.raw twins the reference writes.  smb200_restart
 * needs a freshly created learner; like the reference it accepts a directory that only holds some of the files. */
int smb200_save(smb200_learner* h, const char* base);
/* The two pieces of MemoryBuffer::restart a binding needs when the reference itself has read the files
 * (integration/RACER_B200.cpp): an episode with its stored per-transition values (Episode::unpackEpisode,
 * Episode.cpp:95-128; aggregates recomputed with cmax like Episode::updateCumulative, return estimates kept),
 * and the ReF-ER scalars of the status file ("CmaxReFER", "beta", MemoryBuffer.cpp:252-254). */
int smb200_push_episode_restored(smb200_learner* h, int64_t id, int32_t n_rows, int32_t terminated,
                                 const float* states, const float* actions, const float* policies, const float* rewards,
                                 const float* value, const float* advantage, const float* q_ret, const float* delta,
                                 const float* rho, const float* kl, double cmax);
int smb200_set_refer(smb200_learner* h, double beta, double cmax);
int smb200_restart(smb200_learner* h, const char* base);

/* Actor-side policy evaluation (RACER::selectAction / processTerminal, Learners/RACER.cpp:30-59): raw states in,
 * net outputs out[n][nOut].  Feed-forward networks only (a recurrent network needs the window: smb200_forward_seq).
 * One call answers n agents (the agents of one Master::waitForStateActionCallers poll, Core/Master.cpp:88-145). */
int smb200_forward(smb200_learner* h, const float* states, int32_t n, float* outputs);
/* The same for any network, on the window MemoryBuffer::agentToMinibatch builds for an agent (ReplayMemory/MemoryBuffer.cpp:440-467):
 * states[n][max_len][dS] raw, the first lengths[i] rows of agent i are the states of its episode in progress, oldest first
 * (1 <= lengths[i] <= max_len).  Recurrent networks are evaluated from a zero recurrent state on the newest
 * min(lengths[i], nnBPTTseq + 1) rows (Approximator.h:129-139); feed-forward networks on the newest row.  outputs[n][nOut]
 * are the outputs at the newest row. */
int smb200_forward_seq(smb200_learner* h, const float* states, const int32_t* lengths, int32_t n, int32_t max_len, float* outputs);

/* Device-side timing of the last smb200_train_steps call (CUDA events on the library's
 * stream), and the number of kernels it launched. */
int smb200_last_timing(smb200_learner* h, double* ms_device, int64_t* kernel_launches);
/* Which kernels run the learner steps of this learner: 0 two kernels per step, 1 persistent tile kernel, 2 cluster kernel,
 * 3 wide step (tcgen05 tiles of 128 sampled transitions; feed-forward V-RACER / RACER with continuous actions, batch_size >= 2048 or SMB200_WIDE=1). */
int smb200_step_kernel(const smb200_learner* h);
/* Raw access for benchmarks: run n steps on ids already resident in HBM (uploaded by
 * smb200_presample) without any host<->device copy. */
int smb200_presample(smb200_learner* h, int32_t n_steps);
int smb200_train_presampled(smb200_learner* h, int32_t first, int32_t n);
int smb200_sync(smb200_learner* h);

/* Several learner ranks (one process per GPU of one node) sharing the gradient: replaces the
 * MPI_Iallreduce over learners_train_comm (Network/Optimizer.cpp:114-118) and the delayed
 * reductions of counters / moments (Utils/DelayedReductor.cpp:73-82) by exchanges through
 * peer memory inside the kernels.  comm_init allocates this rank's exchange block and returns
 * its CUDA IPC handle (handle_bytes >= 64); the caller all-gathers the handles of all ranks
 * (rank order) and passes them to comm_attach.  comm_error reports a timed-out peer. */
int smb200_comm_init(smb200_learner* h, int32_t world, int32_t rank, uint8_t* handle_out, int32_t handle_bytes);
int smb200_comm_attach(smb200_learner* h, const uint8_t* handles, int32_t handle_bytes);
int smb200_comm_error(smb200_learner* h);
/* Diagnostics: n presampled steps in one persistent launch with per-CTA phase timestamps
 * (SM clock cycles), out[n][grid][8]; *grid_out = CTAs of the persistent grid. */
int smb200_profile_phases(smb200_learner* h, int32_t n, int64_t* out, int64_t capacity, int32_t* grid_out);
/* Diagnostics, host only (no GPU needed): the statistics phase's emulation of the reference's
 * `Uint nOffPol += float` (ReplayMemory/MemoryProcessing.cpp:202-227) with x86-64 conversion semantics —
 * the same inline function the device code calls, compiled for the host.  Returns the new count. */
uint64_t smb200_uint_plus_float(uint64_t n, float x);
/* The chain `n += xs[first], xs[first + stride], ...` (positions < n_pos) of ONE virtual OpenMP thread of that count as the statistics
 * phase evaluates it: while the count is below 2^24 and the terms are non-negative as a float truncation per term (equal to the
 * integer round trip there), otherwise term by term with smb200_uint_plus_float's emulation.  Host build of the device function. */
uint64_t smb200_host_far_chain(uint64_t n0, const float* xs, int32_t first, int32_t n_pos, int32_t stride);
/* Diagnostics, host only (no GPU needed): the padded parameter blob smb200_create starts from — the layout of
 * Network/Layers/Parameters.h:159-176 initialised like Builder::build does from generators[0] of a run with randSeed =
 * cfg->seed (Network/Builder.cpp:133-137, Layer_Base.h:115-141, Layer_LSTM.h:168-188, ExecutionInfo.cpp:391).
 * Returns the blob size in floats (blob may be NULL to query it), negative on error. */
int64_t smb200_host_init_weights(const smb200_config* cfg, float* blob, int64_t n);
/* Diagnostics, host only (no GPU needed): the wide step's plan for cfg's network (csrc/wide_step.cuh: index maps and pre-split
 * operand images of the tensor-core kernels).  Returns 1 if the wide step covers the network, 0 if not, negative on error.
 * info[16] = {dense layers, forward image floats, transposed image floats, vector block floats, partial-record floats, TMEM
 * columns of the weight-gradient kernel, shared memory of k_wide_fwd / k_wide_bwd / k_wide_wgrad (bytes), image stages,
 * NpG, 0...};  per dense layer d: dense[8 d ...] = {K, Kp, N, Np, forward image offset, transposed image offset, gN, gPart}.
 * With blob (n_blob = smb200_n_params floats): idx[5][n_blob] receives the maps (record position, tile-image, forward-image,
 * transposed-image and vector-block position of every parameter, -1 = none) and img_f / img_b / vec the images
 * wide_fill_images builds from blob (sizes info[1], info[2], info[3]).  NULL pointers are skipped. */
int smb200_host_wide_plan(const smb200_config* cfg, int32_t* info, int32_t* dense, const float* blob, int64_t n_blob, int32_t* idx,
                          float* img_f, float* img_b, float* vec);
/* Diagnostics, host only (no GPU needed): the library's conversion between the padded parameter blob and the order
 * Network::save writes to <name>_net_{weights,tgt_weights,1stMom,2ndMom}.raw (Network/Network.cpp:22-67; padding stripped per
 * layer: Layer_Base.h:143-169, Layers.h:401-418,554-566, Layer_LSTM.h:189-211) — what smb200_save / smb200_restart use.
 * dir = +1: blob -> flat, -1: flat -> blob.  Returns the stripped size in floats (flat may be NULL to query it). */
int64_t smb200_host_strip_weights(const smb200_config* cfg, float* blob, int64_t n_blob, float* flat, int64_t n_flat, int32_t dir);
/* Diagnostics, host only (no GPU needed): the writer of <base>_outGrad_stats.raw (StatsTracker::reduce_stats / printToFile,
 * Utils/StatsTracker.cpp:28-107) that smb200_set_grad_stats switches on, fed with per-sample output gradients
 * g[batch][n_out]; first_tracker_step != 0 starts the file with the n_out + 0.1 header, otherwise the row is appended. */
int smb200_host_write_grad_stats(const char* base, int32_t batch, int32_t n_out, const float* g, int32_t first_tracker_step);
/* Diagnostics, host only (no GPU needed): the episode format of <name>_rank_XXX_learner_data.raw (MemoryBuffer::save /
 * restart, ReplayMemory/MemoryBuffer.cpp:172-324; Episode::packEpisode / unpackEpisode, Episode.cpp:24-130).  Every episode of
 * the file image `in` is read with the parser smb200_restart uses and written to `out` with the packer smb200_save uses;
 * returns the bytes written.  ids / n_rows / terminated (optional, capacity max_eps) report the episodes read. */
int64_t smb200_host_repack_episodes(int32_t dim_state, int32_t dim_action, const uint8_t* in, int64_t n_in, uint8_t* out, int64_t capacity,
                                    int64_t max_eps, int64_t* n_episodes, int64_t* ids, int32_t* n_rows, int32_t* terminated);
/* Diagnostics, host only (no GPU needed): host builds of scalar device functions of the step kernel, same source lines.
 * smb200_host_adam: n elements of the Adam variant of Network/Optimizer.cpp:61-108,122-161 (Nesterov + safe + AdamW) as the
 * weight-gradient epilogue applies it, after adam_step_done completed updates with running beta powers bt1, bt2.
 * smb200_host_value_scaling: scaleNet2V and scaleVdiff (Learners/RACER_common.cpp:23-32). */
int smb200_host_adam(int64_t n, const float* G, float* W, float* M1, float* M2, double learnrate, double eps_anneal, int64_t adam_step_done,
                     double bt1, double bt2, double nn_lambda, int32_t batch_global);
int smb200_host_value_scaling(int64_t n, const double* x, double* v, double* dvdx);
/* Groundwork for discrete actions: RACER<Discrete_advantage, Discrete_policy, Uint>::Train per
 * sample (Learners/RACER_train.cpp:12-67, Math/Discrete_policy.h:64-167, Math/Discrete_advantage.h:44-75) as the
 * __host__ __device__ function the device loss stage will call, run on the host over a batch.  O [B][1+2K], act [B] (label + 0.1),
 * mu [B][K], qret [B]; g [B][1+2K] output gradients, out [B][6] = {rho, D_KL, isFar, V, A, deltaQ}. */
int smb200_host_discrete_loss(int32_t B, int32_t K, const float* O, const float* act, const float* mu, const float* qret, double beta,
                              double cmax, double cinv, double* g, double* out);
/* updateReturnEstimator(EP, N-2) of one episode (ReplayMemory/MemoryProcessing.cpp:23-44) for estimator = retrace / GAE /
 * retraceExplore (:391-417), sequentially on the host with the scalar functions the sweep kernels call (reward scaling,
 * clipped importance weight, the recursion's expression order).  Q is updated in place; returns the sum of squared changes. */
double smb200_host_return_estimator(int32_t n_rows, int32_t terminated, int32_t estimator, const float* R, const float* V,
                                    const float* ADV, const float* RHO, float* Q, double gamma, double lambda,
                                    float reward_mean, float reward_scale, double max_abs_err);
/* Diagnostics, host only (no GPU needed): the host half of n_steps learner steps with no device work — the library's own
 * Sample_uniform::sample + Sampling::IDtoSeqStep (ReplayMemory/Sampling.cpp:26-47,82-93), FIFO applyEpisodesRemovalAlgo
 * (ReplayMemory/MemoryProcessing.cpp:327-351), ring allocator and the Adam update's draw from the sampler's generator
 * (Network/Optimizer.cpp:139) — on n_ep episodes given as (id, rows incl. the terminal row, terminated) in push order.
 * ep_id_out / t_out [n_steps][batch_size]: sampled episode id and time step; n_ep_after [n_steps] and order_out
 * [n_steps][n_ep] (padded with -1): the episode vector after each step (both optional). */
int smb200_host_replay_trace(int32_t batch_size, int64_t max_tot_obs, int64_t capacity_rows, int32_t n_ep, const int64_t* ids,
                             const int32_t* n_rows, const int32_t* terminated, uint64_t seed, int32_t n_steps,
                             int64_t* ep_id_out, int64_t* t_out, int32_t* n_ep_after, int64_t* order_out,
                             const int32_t* push_before_step /* optional, non-decreasing: episode e arrives before that learner step */,
                             int64_t* start_out /* optional [n_steps][n_ep]: first ring row of each live episode after the step */);

#ifdef __cplusplus
}
#endif
#endif /* SMARTIES_B200_H */
