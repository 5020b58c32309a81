/* rddl_b200.h -- C ABI of the B200 batched RDDL stepping backend.
 *
 * This is the boundary a pyRDDLGym maintainer binds (ctypes) to replace
 * the NumPy evaluator behind RDDLSimulator.step and the check_* methods. Each entry point
 * names the reference interface it replaces (paths relative to the pyRDDLGym repository).
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a
 * negative error code, and rddl_last_error() describes the failure; no exceptions cross the
 * ABI. All tensor pointers are DEVICE pointers owned by the caller (C-order, dtype per
 * pyRDDLGym/core/compiler/initializer.py:16-23: bool = uint8, int = int64, real = float64),
 * fluents carry a leading env axis [n_envs, ...], non-fluents are shared by all envs.
 * Exception: bool state fluents that the op table marks with dtype code 3 (OpTable.packed --
 * those only the bit-plane kernel writes) are stored BIT-PACKED, uint32[n_envs, cells / 32], cell i in
 * bit (i % 32) of word i / 32; the host mirror converts to the byte layout on demand.
 * Calls are asynchronous on the given CUDA stream.
 */
#ifndef RDDL_B200_H
#define RDDL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rddl_program rddl_program_t;

enum { RDDL_PLAN_STEP = 0, RDDL_PLAN_TERMINAL = 1, RDDL_PLAN_INVARIANT = 2, RDDL_PLAN_PRECONDITION = 3 };
enum { RDDL_OK = 0, RDDL_ERR_FORMAT = -1, RDDL_ERR_CUDA = -2, RDDL_ERR_ARG = -3 };

/* Replaces RDDLSimulator._compile's product (pyRDDLGym/core/simulator.py:136-174): takes the
 * lowered op table (pyrddlgym_b200/optable.py) and uploads it to `device`. */
int rddl_program_create(const void* optable, size_t nbytes, int device, rddl_program_t** out);
void rddl_program_destroy(rddl_program_t* prog);

/* Text of the last error on this program (or of the last failed create if prog == NULL). */
const char* rddl_last_error(const rddl_program_t* prog);

/* what: 0 = #tensors, 1 = #variate slots, 2 = #launches of the step plan, 3 = total #instructions,
 * 4 = kernel launches issued since create, 5-7 = specialisation counters (below), 9 = fused env-tile
 * scalar kernels in use (runs of scalar launches executed by one kernel), 1000 + tensor id =
 * 1 if a device-side policy has to materialise that action tensor in the caller's buffer each step. */
int64_t rddl_program_info(const rddl_program_t* prog, int what);

/* Replaces RDDLSimulator.step (pyRDDLGym/core/simulator.py:442-502) for n_envs environments:
 * evaluates every CPF in level order, the reward and the terminations on the new state.
 *   tensors[id]  device pointer per op-table tensor id (state, next-state, action, interm,
 *                non-fluent, then __reward f64[n], __done u8[n], __chk u8[n, width]);
 *                next-state tensors are written, the caller swaps state/next pointers after.
 *   injected     NULL, or per variate slot a device pointer to pre-drawn base variates
 *                f64[n_envs, *scope] (NULL entries draw Philox4x32-10 keyed by
 *                (seed, env_offset + env, step, node id, element)).
 *   status       u32[n_envs] error bits OR-ed in (replaces the exceptions raised by
 *                _check_positive/_check_bounds/_check_range, simulator.py:229-253). */
int rddl_step(rddl_program_t* prog, void* const* tensors, const double* const* injected,
              uint32_t* status, int64_t n_envs, int64_t env_offset, uint64_t seed, uint64_t step,
              void* cuda_stream);

/* Replaces check_state_invariants / check_action_preconditions / check_terminal_states
 * (pyRDDLGym/core/simulator.py:362-399). which = RDDL_PLAN_INVARIANT | _PRECONDITION writes one
 * bool per expression into __chk[n, width]; RDDL_PLAN_TERMINAL writes __done[n]. Random variables
 * inside a constraint (legal RDDL) draw Philox with (seed, step) like rddl_step. */
int rddl_check(rddl_program_t* prog, int which, void* const* tensors, uint32_t* status,
               int64_t n_envs, int64_t env_offset, uint64_t seed, uint64_t step, void* cuda_stream);

/* Replaces the rollout loop of BaseAgent.evaluate (pyRDDLGym/core/policy.py:85-117: for each
 * step, env.step(action); total_reward += reward * gamma^t; stop accumulating once done) for
 * n_envs environments and n_steps steps with pre-drawn action sequences:
 *   step_stride_bytes[id]  byte offset from one step's buffer of tensor id to the next step's
 *                          (action sequences laid out [n_steps, n_envs, ...]); 0 = same buffer
 *                          every step; NULL = all 0
 *   returns                f64[n_envs], += discounted return of this rollout
 *   rewards                NULL or f64[n_steps, n_envs], per-step rewards
 *   state_in_next          out: 1 if the final state is in the next-state buffers (caller swaps
 *                          state/next pointers, as after rddl_step), 0 if it is in the state buffers
 * Step t draws Philox with step index step0 + t, so the result is identical to n_steps calls of
 * rddl_step. When the whole step plan is one single-tile bit-plane launch, all steps run inside
 * one kernel with the state held on chip (only the action planes stream from HBM). */
int rddl_rollout(rddl_program_t* prog, void* const* tensors, uint32_t* status, int64_t n_envs,
                 int64_t env_offset, uint64_t seed, uint64_t step0, int32_t n_steps,
                 const int64_t* step_stride_bytes, double gamma, double* returns, double* rewards,
                 int32_t* state_in_next, void* cuda_stream);

/* Device-side random policy: the batched counterpart of RandomAgent / NoOpAgent
 * (pyRDDLGym/core/policy.py:129-195). RandomAgent.sample_action draws every action grounding uniformly
 * from its range and then keeps `num_actions` randomly chosen groundings (policy.py:166-178); here the
 * draws happen inside the kernels (no action tensor crosses PCIe or HBM):
 *   kind RDDL_POLICY_BERNOULLI     bool fluent, 1 with probability lo (1/2 = the agent's Discrete(2))
 *        RDDL_POLICY_UNIFORM_REAL  real fluent, lo + (hi - lo) * U,  U in [0, 1)
 *        RDDL_POLICY_UNIFORM_INT   int fluent, uniform on {lo, ..., hi}
 *        RDDL_POLICY_DEFAULT       the fluent keeps its default (NoOpAgent)
 *   dflt        the fluent's default (noop) value
 *   max_nondef  < 0 or >= #groundings: every grounding keeps its draw; otherwise only max_nondef
 *               groundings per (env, step), chosen uniformly without replacement over the groundings of
 *               all entries in entry order (at most 32), keep theirs and the rest take the default --
 *               the max-nondef-actions rule of check_default_action_count (simulator.py:329-360).
 * Every draw is a function of (seed, global env id, step, fluent, grounding) only; the mapping is
 * restated in oracle/philox_np.py (policy_actions). Action fluents not listed are read from `tensors`. */
enum { RDDL_POLICY_BERNOULLI = 1, RDDL_POLICY_UNIFORM_REAL = 2, RDDL_POLICY_UNIFORM_INT = 3, RDDL_POLICY_DEFAULT = 4 };
typedef struct { int32_t tensor, kind; double lo, hi, dflt; } rddl_policy_entry_t;
typedef struct { int32_t n_entries, max_nondef; rddl_policy_entry_t entries[8]; } rddl_policy_t;

/* rddl_rollout with the actions of `policy` drawn on the device (the loop of BaseAgent.evaluate with a
 * RandomAgent, policy.py:85-117): same arguments and results as rddl_rollout minus the action
 * sequences; n_steps = 1 is one policy-driven transition. Identical to materialising every step's
 * actions with rddl_policy_sample and calling rddl_rollout with them. */
int rddl_rollout_policy(rddl_program_t* prog, const rddl_policy_t* policy, void* const* tensors, uint32_t* status,
                        int64_t n_envs, int64_t env_offset, uint64_t seed, uint64_t step0, int32_t n_steps,
                        double gamma, double* returns, double* rewards, int32_t* state_in_next, void* cuda_stream);

/* Writes the actions the policy takes at `step` into the action tensors (tensors[entry.tensor],
 * [n_envs, ...]): RandomAgent.sample_action for every env (policy.py:166-178). */
int rddl_policy_sample(rddl_program_t* prog, const rddl_policy_t* policy, void* const* tensors, int64_t n_envs,
                       int64_t env_offset, uint64_t seed, uint64_t step, void* cuda_stream);

/* Replaces the counting part of check_default_action_count (simulator.py:329-360) as a device
 * reduction: counts[e] (int32[n_envs], device) = number of groundings of the n_actions action tensors
 * action_tensors[i] whose value differs from default_values[i]; the caller compares the maximum with
 * max_allowed_actions and raises RDDLInvalidActionError. */
int rddl_action_count(rddl_program_t* prog, void* const* tensors, const int32_t* action_tensors,
                      const double* default_values, int32_t n_actions, int32_t* counts, int64_t n_envs,
                      void* cuda_stream);

/* Replaces the statistics of BaseAgent.evaluate (pyRDDLGym/core/policy.py:120-126: np.mean / std / min
 * / max of the per-episode returns) for this rank's shard: out5 = {sum, sum of squares, count, min,
 * max} of returns f64[n] (device pointers), one deterministic block reduction. Ranks combine their
 * five numbers with one NCCL all-gather (pyrddlgym_b200/parallel.py). */
int rddl_return_stats(const double* returns, int64_t n, double* out5, void* cuda_stream);

/* Run-time specialisation of the kernels (no reference counterpart; the closest analogue is the
 * reference's JAX backend jit-compiling a model once, pyRDDLGym-jax, docs/jax.rst). The library
 * recompiles its own kernel sources for sm_100a with NVRTC with a launch's program, pools and
 * geometry as constants (pyrddlgym_b200/csrc/rddl_jit.h): plane launches on their first run, the
 * scalar launches of a program (one module) once it has issued a number of scalar launches. Results
 * are bit-identical to the interpreter kernels', which run instead whenever NVRTC is unavailable.
 *   rddl_program_option(prog, 0, v)  v = 0: interpreter kernels only; 1 (default unless the
 *                                    environment has RDDL_B200_JIT=0): specialise
 *   rddl_program_option(prog, 1, n)  scalar kernels are specialised after n scalar launches
 *                                    (default 24, $RDDL_B200_JIT_HOT; 0 = on first use)
 *   rddl_program_option(prog, 2, v)  sum-product contractions (sum_{?a} W(?a,?b) * x(?a), simulator.py:584-637,766-779):
 *                                    v = 0 (default) exact FP64 tensor-core kernel (DMMA); v = 1 the Blackwell
 *                                    tcgen05 kernel on three-term bfloat16 splits of both operands with float32
 *                                    accumulation in tensor memory (TMA operand loads) -- within 1e-5 relative
 *   rddl_program_info(prog, 5 | 6 | 7)  specialised kernels in use | launches that fell back to the
 *                                    interpreter (rddl_jit_error has the compiler log) | cubins
 *                                    taken from the disk cache (<library dir>/jit_cache or
 *                                    $RDDL_B200_JIT_CACHE)
 *   rddl_plane_precompile            host only (no device needed): fills the disk cache with the
 *                                    modules an op table needs for batches of n_envs (one per plane
 *                                    launch + one for all scalar launches); returns their number,
 *                                    < 0 on error (rddl_last_error(NULL))
 *   rddl_plane_source                inspection aid: the generated constants of plane launch `launch`;
 *                                    returns the text length, writes at most cap - 1 bytes + NUL */
int rddl_program_option(rddl_program_t* prog, int what, int64_t value);
const char* rddl_jit_error(const rddl_program_t* prog);
int rddl_plane_precompile(const void* optable, size_t nbytes, int64_t n_envs);
int rddl_policy_precompile(const void* optable, size_t nbytes, int64_t n_envs, const rddl_policy_t* policy);
int64_t rddl_plane_source(const void* optable, size_t nbytes, int launch, int64_t n_envs, char* out, size_t cap);

/* Measurement aid (no reference counterpart): when enabled, every kernel launch of this
 * program is bracketed by CUDA events on its stream; rddl_profile_read synchronises them and
 * returns, per launch index of the op table (n entries), the summed milliseconds and the
 * number of launches since the last read. */
int rddl_profile(rddl_program_t* prog, int enable);
int rddl_profile_read(rddl_program_t* prog, double* ms, int64_t* count, int n);

#ifdef __cplusplus
}
#endif
#endif
