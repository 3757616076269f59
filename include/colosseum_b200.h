/* colosseum_b200.h -- C ABI of libcolosseum_b200.so: batched, bit-packed, B200-native (sm_100a) game
 * dynamics for ColosseumRL's Tron, Blokus and 2/3/4-player Tic Tac Toe.
 *
 * The reference has no FFI for this path: its boundary is the Python ABC `BaseEnvironment`
 * (colosseumrl/BaseEnvironment.py:10-283) plus one Cython entry point
 * (`next_state_inplace`, colosseumrl/envs/tron/CyTronGrid.pyx:3-7).  Each entry point below names the
 * reference function(s) it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *  - every `state`, `actions`, `result`, `stats`, ... pointer is a DEVICE pointer into caller-owned memory
 *    (e.g. torch.Tensor.data_ptr()); the library never allocates, frees or retains them;
 *  - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and never synchronise;
 *  - return value 0 = OK, non-zero = error (crl_last_error() gives a thread-local message);
 *  - per-environment data errors (illegal action id) never fail the call: they set the error bit of that
 *    environment's result record and are applied as a pass / no-op (match_server.py:192-198 behaviour);
 *  - B = number of environments in the batch; all per-environment arrays are dense over [0, B).
 *  - there is NO CPU fallback: without a CUDA device every compute entry point returns CRL_ERR_CUDA.
 */
#ifndef COLOSSEUM_B200_H
#define COLOSSEUM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRL_OK 0
#define CRL_ERR_ARG 1
#define CRL_ERR_CUDA 2
#define CRL_ERR_UNSUPPORTED 3

#define CRL_FLAG_AUTO_RESET 1 /* an environment whose stored terminal flag is set is replaced by new_state() before the step */

/* statistics vector: int64[CRL_NSTAT] on the device, accumulated (+=) by step / rollout kernels */
#define CRL_NSTAT 32
#define CRL_ST_STEPS 0     /* env-steps                                   */
#define CRL_ST_EPISODES 1  /* finished episodes (terminal transitions)    */
#define CRL_ST_EPLEN 2     /* sum of finished-episode lengths             */
#define CRL_ST_WINS 3      /* [3..6]  wins per seat                       */
#define CRL_ST_NOWIN 7     /* finished episodes without a winner          */
#define CRL_ST_SCORE 8     /* [8..11] sum of final scores per seat        */
#define CRL_ST_ERRORS 12   /* illegal actions                             */
#define CRL_ST_NVALID 13   /* sum of valid-action counts                  */
#define CRL_ST_RANK 14     /* [14..17] sum of final ranks per seat        */
#define CRL_ST_REWARD 18   /* sum over steps of sum_p (p+1) * reward_p    */

typedef void *crl_stream_t; /* cudaStream_t */

int crl_version(void);
const char *crl_last_error(void);
/* Select + validate the device (compute capability 10.x required). Idempotent. */
int crl_init(int device);

/* Philox4x32-10 words for environments [first_env, first_env+B) at `step`: out = uint32[B][4].
 * counter = (env_lo, env_hi, step, tag), key = (seed_lo, seed_hi). */
int crl_philox_words(uint32_t *out, uint64_t seed, uint64_t first_env, uint32_t step, uint32_t tag, int64_t B,
                     crl_stream_t stream);

/* ------------------------------------------------------------------------------------------- Tron
 * Packed state: 208 bytes per environment, SoA [13][B] of 16-byte vectors (csrc/tron.cuh).
 * Supported: 5 <= N <= 19 (N*N <= 384), 2 <= P <= 4.
 * actions: int8[B][4]  (0 forward, +1 right, -1 left; TronGridEnvironment.STRING_TO_ACTION :62-67),
 *          entries of dead / absent players are ignored.
 * result:  8 bytes per environment: int8 reward[4] | u8 terminal | u8 alive mask | u8 winners mask |
 *          u8 ranking (2 bits per player, competition ranking of compute_ranking).              */
int64_t crl_tron_state_bytes(int N, int P, int64_t B);
/* HOST function. generate_start_positions (TronGridEnvironment.py:183-226) with new_state's defaults:
 * heads[p] = y*N + x, directions[p] in {0 N, 1 E, 2 S, 3 W}. */
int crl_tron_start_positions(int N, int P, int32_t *heads, int32_t *directions);
/* new_state (TronGridEnvironment.py:228-263) for every environment, or where mask[e] != 0 */
int crl_tron_reset(void *state, const uint8_t *mask_or_null, int64_t B, int N, int P, crl_stream_t stream);
/* next_state (TronGridEnvironment.py:265-323 -> CyTronGrid.pyx:3-62) + compute_ranking (:483-508).
 * state_out may equal state_in (in place). stats may be NULL. */
int crl_tron_step(const void *state_in, void *state_out, const int8_t *actions, uint8_t *result,
                  int64_t *stats_or_null, int64_t B, int N, int P, int flags, crl_stream_t stream);
/* uniform random policy: action of player p = {0,+1,-1}[philox(env, step, tag 1)[p] % 3] */
int crl_tron_policy_random(int8_t *actions, uint64_t seed, uint64_t first_env, uint32_t step, int64_t B,
                           crl_stream_t stream);
/* K fused random-policy steps with auto-reset, state kept in registers (identical to K x policy+step) */
int crl_tron_rollout(void *state, uint8_t *result_or_null, int64_t *stats_or_null, uint64_t seed,
                     uint64_t first_env, uint32_t step0, int K, int64_t B, int N, int P, crl_stream_t stream);
/* state_to_observation (TronGridEnvironment.py:363-405, CyTronGrid.pyx:65-71). player < 0: absolute unpack.
 * board int8[B][N][N]; heads (y*N+x) / directions / deaths int32[B][P] (may be NULL); terminal u8[B] (may be NULL) */
int crl_tron_observe(const void *state, int player, int8_t *board, int32_t *heads, int32_t *directions,
                     int32_t *deaths, uint8_t *terminal_or_null, int64_t B, int N, int P, crl_stream_t stream);
/* import a reference-layout state */
int crl_tron_pack(void *state, const int8_t *board, const int32_t *heads, const int32_t *directions,
                  const int32_t *deaths, int64_t B, int N, int P, crl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
