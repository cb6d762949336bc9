// gmz_records.cu -- device-side trajectory hand-off (reference workers.py:172-230, 399-433).
//
// A finished game leaves the trajectory store as PACKED MOVE RECORDS: one fixed-stride record per move holding
// everything the reference's self-play loop emits for that move -- the observation planes (game.py:12-17), the
// search policy (float64, as the reference keeps it), the board before the move, and the scalars (action, root
// search value, final reward workers.py:183-187, n-step value target workers.py:144-152).  The same bytes serve
//   * the host: GameRecord / TrainingSlice objects are numpy views of the copied block (no per-move arithmetic),
//   * the wire: ranks gather them with one collective of raw bytes (replaces data_queue, workers.py:230, 399),
//   * the replay ring on the device: a TrainingSlice is U+1 consecutive records of one game, so a training batch
//     (workers.py:430-433) is gathered straight from the ring, D4 augmentation (loss.py:37-51) included.
// Record layout (gmz_move_record_bytes): 64-byte header | policy f64[A] | obs f32[3A] | board i8[A] | pad to 16.
#include <stdint.h>

#include "../../include/gmz.h"
#include "gmz_common.cuh"

extern "C" void gmz_set_error_(const char *msg);
static int rc_fail(const char *m) { gmz_set_error_(m); return 1; }

struct RecHeader {            // 64 bytes, see include/gmz.h gmz_move_record
    int32_t game_seq, t, length, winner;
    int32_t action, to_move, last_move, move_count;
    float reward, value_target;
    double search_value;
    int32_t game, slot, pad0, pad1;
};
static_assert(sizeof(RecHeader) == 64, "record header must be 64 bytes");

static inline size_t rec_stride(int A) { return ((size_t)64 + (size_t)21 * A + 15) / 16 * 16; }
extern "C" size_t gmz_move_record_bytes(int board_size)
{
    if (board_size < 1 || board_size > GMZ_MAX_BOARD) return 0;
    return rec_stride(board_size * board_size);
}

__device__ __forceinline__ float rec_final_reward(int i, int T, int winner)   // workers.py:183-187
{
    if (winner == 0 || i < 0 || i >= T) return 0.0f;
    const int j = (T - 1 - i) & 3;
    return (j == 0 || j == 3) ? 1.0f : -1.0f;
}
__device__ __forceinline__ void rec_apply(u64 &P, u64 &M, int colour, int a, int lane)
{
    if (lane == (a >> 6)) { const u64 bit = 1ull << (a & 63); if (colour > 0) { P |= bit; M &= ~bit; } else { M |= bit; P &= ~bit; } }
}

// One CTA per finished game, warp w emits moves w, w + 4, ... (each warp replays the game on its own bitboards).
__global__ void __launch_bounds__(128)
k_traj_pack(int N, int A, size_t stride, int max_moves, const double *policy, const double *value, const int32_t *action,
            const u64 *start_board, const int32_t *start_info, const int32_t *fin, int n_games, const int64_t *move_offset,
            const double *dpow, int n_steps, unsigned char *out)
{
    const int gi = blockIdx.x, lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    if (gi >= n_games) return;
    const int slot = fin[4 * gi + 0], game = fin[4 * gi + 1], T = min(fin[4 * gi + 2], max_moves), winner = fin[4 * gi + 3];
    u64 P = lane < GMZ_WORDS ? start_board[((size_t)slot * 2 + 0) * GMZ_WORDS + lane] : 0ull;
    u64 M = lane < GMZ_WORDS ? start_board[((size_t)slot * 2 + 1) * GMZ_WORDS + lane] : 0ull;
    int colour = start_info[(size_t)slot * 4 + 0], mc = start_info[(size_t)slot * 4 + 1], last = start_info[(size_t)slot * 4 + 2];
    const int32_t *acts = action + (size_t)slot * max_moves;
    const double *vals = value + (size_t)slot * max_moves;
    const float gn = (float)dpow[n_steps];
    int m = 0;
    for (int t = wi; t < T; t += 4) {
        for (; m < t; ++m) { const int a = acts[m]; rec_apply(P, M, colour, a, lane); colour = -colour; last = a; ++mc; }
        unsigned char *rec = out + (size_t)(move_offset[gi] + t) * stride;
        double *pol = reinterpret_cast<double *>(rec + 64);
        float *obs = reinterpret_cast<float *>(rec + 64 + (size_t)8 * A);
        int8_t *brd = reinterpret_cast<int8_t *>(rec + 64 + (size_t)20 * A);
        const double *psrc = policy + ((size_t)slot * max_moves + t) * A;
        const u64 own = colour > 0 ? P : M, opp = colour > 0 ? M : P;
        for (int c0 = 0; c0 < A; c0 += 32) {           // all 32 lanes take part in the shuffles
            const int c = c0 + lane, cw = min(c, A - 1) >> 6;
            const u64 ow = __shfl_sync(GMZ_FULL, own, cw), pw = __shfl_sync(GMZ_FULL, opp, cw);
            const u64 bp = __shfl_sync(GMZ_FULL, P, cw), bm = __shfl_sync(GMZ_FULL, M, cw);
            if (c < A) {
                obs[c] = (float)((ow >> (c & 63)) & 1ull);
                obs[A + c] = (float)((pw >> (c & 63)) & 1ull);
                obs[2 * A + c] = c == last ? 1.0f : 0.0f;
                brd[c] = (int8_t)((int)((bp >> (c & 63)) & 1ull) - (int)((bm >> (c & 63)) & 1ull));
                pol[c] = psrc[c];
            }
        }
        if (lane == 0) {
            // compute_n_step_returns (workers.py:144-152): Python-float reward sum, float32 bootstrap, float32 add
            double acc = 0.0;
            for (int i = 0; i < n_steps; ++i)
                if (t + i < T) acc = __dadd_rn(acc, __dmul_rn(dpow[i], (double)rec_final_reward(t + i, T, winner)));
            const int b = t + n_steps;
            RecHeader h;
            h.game_seq = gi; h.t = t; h.length = T; h.winner = winner;
            h.action = acts[t]; h.to_move = colour; h.last_move = last; h.move_count = mc;
            h.reward = rec_final_reward(t, T, winner);
            h.value_target = b < T ? __fadd_rn((float)acc, __fmul_rn((float)vals[b], gn)) : (float)acc;
            h.search_value = vals[t];
            h.game = game; h.slot = slot; h.pad0 = 0; h.pad1 = 0;
            *reinterpret_cast<RecHeader *>(rec) = h;
        }
    }
}

extern "C" int gmz_traj_pack(const gmz_traj *traj, int board_size, const int32_t *fin, int n_games, const int64_t *move_offset,
                             const double *discount_pow, int n_steps, void *out_records, gmz_stream stream)
{
    if (!traj || !fin || !move_offset || !discount_pow || !out_records) return rc_fail("gmz_traj_pack: null argument");
    if (n_games <= 0) return 0;
    if (board_size < 1 || board_size > GMZ_MAX_BOARD || n_steps < 0) return rc_fail("gmz_traj_pack: bad size");
    const int A = board_size * board_size;
    k_traj_pack<<<n_games, 128, 0, (cudaStream_t)stream>>>(board_size, A, rec_stride(A), traj->max_moves, traj->policy, traj->value,
                                                           traj->action, (const u64 *)traj->start_board, traj->start_info, fin, n_games,
                                                           move_offset, discount_pow, n_steps, (unsigned char *)out_records);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : rc_fail(cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------------------------------------
// Training batches from a ring of records (DeviceReplayBuffer): sample b is the record at ring position pos[b]
// (= the reference's data index, replay_buffer.py:50-55, 80); its slice is that record and the next U of the same
// game (they sit at the following ring positions: records are appended game by game in move order, and a ring
// overwrites oldest first, so the successors of a live record are live), padded past the game's end like
// workers.py:208-222 (zero planes / policies / values / rewards, action -1).
__device__ __forceinline__ int rec_sym_cell(int c, int N, int rk, int fl)      // loss.py:39-44, as torch.rot90 / flip
{
    const int r = c / N, q = c - r * N;
    int i = r, j = q;
    if (rk == 1) { i = N - 1 - q; j = r; }
    else if (rk == 2) { i = N - 1 - r; j = N - 1 - q; }
    else if (rk == 3) { i = q; j = N - 1 - r; }
    if (fl) j = N - 1 - j;
    return i * N + j;
}
__device__ __forceinline__ int rec_sym_action(int a, int N, int rk, int fl)     // loss.py:46-51, as written
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
k_records_batch(int N, int A, size_t stride, int64_t capacity, const unsigned char *ring, const int64_t *pos, int B, int U,
                int rk, int fl, float *obs, int32_t *act, float *rew, double *pi, float *val)
{
    const int lane = threadIdx.x & 31, b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int64_t p0 = pos[b];
    const RecHeader h0 = *reinterpret_cast<const RecHeader *>(ring + (size_t)p0 * stride);
    for (int k = 0; k <= U; ++k) {
        float *o = obs + ((size_t)b * (U + 1) + k) * 3 * A;
        double *pk = pi + ((size_t)b * (U + 1) + k) * A;
        if (h0.t + k < h0.length) {
            const unsigned char *rec = ring + (size_t)((p0 + k) % capacity) * stride;
            const RecHeader *h = reinterpret_cast<const RecHeader *>(rec);
            const double *pol = reinterpret_cast<const double *>(rec + 64);
            const float *ob = reinterpret_cast<const float *>(rec + 64 + (size_t)8 * A);
            for (int c = lane; c < A; c += 32) {
                const int d = rec_sym_cell(c, N, rk, fl);
                o[d] = ob[c]; o[A + d] = ob[A + c]; o[2 * A + d] = ob[2 * A + c];
                pk[d] = pol[c];
            }
            if (lane == 0) {
                val[(size_t)b * (U + 1) + k] = h->value_target;
                if (k < U) { act[(size_t)b * U + k] = rec_sym_action(h->action, N, rk, fl); rew[(size_t)b * U + k] = h->reward; }
            }
        } else {
            for (int c = lane; c < A; c += 32) { o[c] = 0.f; o[A + c] = 0.f; o[2 * A + c] = 0.f; pk[c] = 0.0; }
            if (lane == 0) {
                val[(size_t)b * (U + 1) + k] = 0.0f;
                if (k < U) { act[(size_t)b * U + k] = -1; rew[(size_t)b * U + k] = 0.0f; }
            }
        }
    }
}

extern "C" int gmz_records_batch(const void *ring, int64_t capacity, int board_size, const int64_t *positions, int batch, int unroll,
                                 int rot_k, int flip, float *obs, int32_t *act, float *rew, double *pi, float *val, gmz_stream stream)
{
    if (!ring || !positions || !obs || !act || !rew || !pi || !val) return rc_fail("gmz_records_batch: null argument");
    if (batch <= 0) return 0;
    if (board_size < 1 || board_size > GMZ_MAX_BOARD || unroll < 0 || capacity < 1) return rc_fail("gmz_records_batch: bad size");
    if (rot_k < 0 || rot_k > 3) return rc_fail("gmz_records_batch: rot_k must be 0..3");
    const int A = board_size * board_size;
    k_records_batch<<<(batch + 3) / 4, 128, 0, (cudaStream_t)stream>>>(board_size, A, rec_stride(A), capacity, (const unsigned char *)ring,
                                                                       positions, batch, unroll, rot_k, flip ? 1 : 0, obs, act, rew, pi, val);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : rc_fail(cudaGetErrorString(e));
}
