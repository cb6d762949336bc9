// gmz_slices.cu -- trajectory post-processing on the device (reference workers.py:144-152, 183-222,
// 430-433): final-reward pattern, n-step value targets, and TrainingSlice batches assembled straight
// from the resident games.  A "slice" is never materialised per move (the reference stores U+1
// observations and policies per move, a 6x duplication): it is the pair (slot, t), and
// gmz_build_batch rebuilds its tensors by replaying the game's moves on bitboards.
#include <stdint.h>
#include <stdio.h>

#include "../../include/gmz.h"
#include "gmz_common.cuh"

extern "C" void gmz_set_error_(const char *msg);
static int sl_fail(const char *m) { gmz_set_error_(m); return 1; }

// final_rewards (workers.py:183-187): r[T-1] = +1, r[T-2] = -1, r[i] = -r[i+2]; zeros on a draw.
__device__ __forceinline__ float final_reward(int i, int T, int winner)
{
    if (winner == 0 || i < 0 || i >= T) return 0.0f;
    const int j = (T - 1 - i) & 3;
    return (j == 0 || j == 3) ? 1.0f : -1.0f;
}

// compute_n_step_returns as called from self-play (workers.py:144-152, 205): double reward sum
// (rewards are Python floats there), float32 bootstrap = float32(value) * float32(discount**n),
// float32 add.  dpow[i] = discount**i computed by the host exactly as Python does.
__global__ void k_value_targets(const double *value, int max_moves, const int32_t *slots, const int32_t *lengths,
                                const int32_t *winners, int n_games, const double *dpow, int n_steps, float *targets)
{
    const int gi = blockIdx.x;
    if (gi >= n_games) return;
    const int slot = slots[gi], T = min(lengths[gi], max_moves), w = winners[gi];
    const float gn = (float)dpow[n_steps];
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double acc = 0.0;
        for (int i = 0; i < n_steps; ++i)
            if (t + i < T) acc = __dadd_rn(acc, __dmul_rn(dpow[i], (double)final_reward(t + i, T, w)));
        const int b = t + n_steps;
        float out;
        if (b < T) out = __fadd_rn((float)acc, __fmul_rn((float)value[(size_t)slot * max_moves + b], gn));
        else out = (float)acc;
        targets[(size_t)slot * max_moves + t] = out;
    }
}

// One warp per sampled slice: replay the game to move t, then emit U+1 observations / policies /
// values and U actions / rewards with the reference's padding (zeros, -1) past the end of the game.
// D4 symmetry of calculate_loss (loss.py:37-44): planes and policies turn rk quarter turns the way
// torch.rot90 does (cell (r,c) -> (N-1-c, r) per turn), then flip left-right.
__device__ __forceinline__ int sym_cell(int c, int N, int rk, int fl)
{
    const int r = c / N, q = c - r * N;
    int i = r, j = q;
    if (rk == 1) { i = N - 1 - q; j = r; }
    else if (rk == 2) { i = N - 1 - r; j = N - 1 - q; }
    else if (rk == 3) { i = q; j = N - 1 - r; }
    if (fl) j = N - 1 - j;
    return i * N + j;
}
// The action index follows the reference's own formula (loss.py:46-51) -- for rk = 1, 3 that is the
// opposite quarter turn of the planes'; kept as written.  The -1 padding stays -1 (the reference masks
// on the un-augmented actions, loss.py:85, and never reads what its formula makes of a padded entry).
__device__ __forceinline__ int sym_action(int a, int N, int rk, int fl)
{
    if (a < 0) return a;
    int rows = a / N, cols = a - rows * N;
    if (rk == 1) { const int t = rows; rows = cols; cols = N - 1 - t; }
    else if (rk == 2) { rows = N - 1 - rows; cols = N - 1 - cols; }
    else if (rk == 3) { const int t = rows; rows = N - 1 - cols; cols = t; }
    if (fl) cols = N - 1 - cols;
    return rows * N + cols;
}

__global__ void __launch_bounds__(128)
k_build_batch(int N, int A, int max_moves, const double *policy, const int32_t *action, const u64 *start_board,
              const int32_t *start_info, const float *targets, const int32_t *len_by_slot, const int32_t *win_by_slot,
              const int32_t *s_slot, const int32_t *s_t, int B, int U, int rk, int fl,
              float *obs, int32_t *act, float *rew, double *pi, float *val)
{
    const int lane = threadIdx.x & 31, b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int slot = s_slot[b], t0 = s_t[b];
    const int T = min(len_by_slot[slot], max_moves), w = win_by_slot[slot];
    u64 P = lane < GMZ_WORDS ? start_board[((size_t)slot * 2 + 0) * GMZ_WORDS + lane] : 0ull;
    u64 M = lane < GMZ_WORDS ? start_board[((size_t)slot * 2 + 1) * GMZ_WORDS + lane] : 0ull;
    int colour = start_info[(size_t)slot * 4 + 0], last = start_info[(size_t)slot * 4 + 2];
    const int32_t *acts = action + (size_t)slot * max_moves;
    for (int m = 0; m < t0 && m < T; ++m) {            // positions before move t0
        const int a = acts[m];
        if (lane == (a >> 6)) { const u64 bit = 1ull << (a & 63); if (colour > 0) { P |= bit; M &= ~bit; } else { M |= bit; P &= ~bit; } }
        colour = -colour; last = a;
    }
    for (int k = 0; k <= U; ++k) {
        const int m = t0 + k;
        float *o = obs + ((size_t)b * (U + 1) + k) * 3 * A;
        double *pk = pi + ((size_t)b * (U + 1) + k) * A;
        if (m < T) {
            const u64 own = colour > 0 ? P : M, opp = colour > 0 ? M : P;
            for (int c0 = 0; c0 < A; c0 += 32) {       // all 32 lanes take part in the shuffles
                const int c = c0 + lane, cw = min(c, A - 1) >> 6;
                const u64 ow = __shfl_sync(GMZ_FULL, own, cw), pw = __shfl_sync(GMZ_FULL, opp, cw);
                if (c < A) {
                    const int d = sym_cell(c, N, rk, fl);
                    o[d] = (float)((ow >> (c & 63)) & 1ull);
                    o[A + d] = (float)((pw >> (c & 63)) & 1ull);
                    o[2 * A + d] = c == last ? 1.0f : 0.0f;
                    pk[d] = policy[((size_t)slot * max_moves + m) * A + c];
                }
            }
            if (lane == 0) val[(size_t)b * (U + 1) + k] = targets[(size_t)slot * max_moves + m];
            const int a = acts[m];
            if (k < U && lane == 0) { act[(size_t)b * U + k] = sym_action(a, N, rk, fl); rew[(size_t)b * U + k] = final_reward(m, T, w); }
            if (lane == (a >> 6)) { const u64 bit = 1ull << (a & 63); if (colour > 0) { P |= bit; M &= ~bit; } else { M |= bit; P &= ~bit; } }
            colour = -colour; last = a;
        } else {
            for (int c = lane; c < A; c += 32) { o[c] = 0.f; o[A + c] = 0.f; o[2 * A + c] = 0.f; pk[c] = 0.0; }
            if (lane == 0) {
                val[(size_t)b * (U + 1) + k] = 0.0f;
                if (k < U) { act[(size_t)b * U + k] = sym_action(-1, N, rk, fl); rew[(size_t)b * U + k] = 0.0f; }
            }
        }
    }
}

extern "C" int gmz_value_targets(const gmz_traj *traj, const int32_t *slots, const int32_t *lengths, const int32_t *winners,
                                 int n_games, const double *discount_pow, int n_steps, float *targets, gmz_stream stream)
{
    if (!traj || !slots || !lengths || !winners || !discount_pow || !targets) return sl_fail("gmz_value_targets: null argument");
    if (n_games <= 0) return 0;
    if (n_steps < 0) return sl_fail("gmz_value_targets: n_steps < 0");
    k_value_targets<<<n_games, 128, 0, (cudaStream_t)stream>>>(traj->value, traj->max_moves, slots, lengths, winners, n_games,
                                                               discount_pow, n_steps, targets);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : sl_fail(cudaGetErrorString(e));
}

extern "C" int gmz_build_batch_aug(const gmz_traj *traj, int board_size, const float *targets, const int32_t *len_by_slot,
                                   const int32_t *win_by_slot, const int32_t *sample_slot, const int32_t *sample_t, int batch,
                                   int unroll, int rot_k, int flip, float *obs, int32_t *act, float *rew, double *pi,
                                   float *val, gmz_stream stream)
{
    if (!traj || !targets || !len_by_slot || !win_by_slot || !sample_slot || !sample_t || !obs || !act || !rew || !pi || !val)
        return sl_fail("gmz_build_batch: null argument");
    if (batch <= 0) return 0;
    if (board_size < 1 || board_size > GMZ_MAX_BOARD || unroll < 0) return sl_fail("gmz_build_batch: bad size");
    if (rot_k < 0 || rot_k > 3) return sl_fail("gmz_build_batch: rot_k must be 0..3");
    k_build_batch<<<(batch + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
        board_size, board_size * board_size, traj->max_moves, traj->policy, traj->action, (const u64 *)traj->start_board,
        traj->start_info, targets, len_by_slot, win_by_slot, sample_slot, sample_t, batch, unroll, rot_k, flip ? 1 : 0,
        obs, act, rew, pi, val);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : sl_fail(cudaGetErrorString(e));
}

extern "C" int gmz_build_batch(const gmz_traj *traj, int board_size, const float *targets, const int32_t *len_by_slot,
                               const int32_t *win_by_slot, const int32_t *sample_slot, const int32_t *sample_t, int batch,
                               int unroll, float *obs, int32_t *act, float *rew, double *pi, float *val, gmz_stream stream)
{
    return gmz_build_batch_aug(traj, board_size, targets, len_by_slot, win_by_slot, sample_slot, sample_t, batch, unroll,
                               0, 0, obs, act, rew, pi, val, stream);
}
