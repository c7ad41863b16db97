/* transgo_b200.h -- C ABI of libtransgo_b200.so (B200 / sm_100a batched self-play engine for Transgo).
 *
 * Two groups of entry points, both plain C (pointers + sizes, no torch / C++ types):
 *
 *  (1) The reference's own 15 symbols (GoEnv/cpp_src/go_env.h:23-71), so that the unmodified
 *      GoEnv/environment.py can `ctypes.cdll.LoadLibrary` this library in place of go_env.so
 *      (environment.py:42).  They take ONE host-side state blob (the caller's >=1188-byte c_GoState;
 *      this library uses its first tg_state_bytes(n) bytes) and run the same CUDA kernels as (2) with
 *      G = 1: copy in, launch, copy out.  There is no CPU implementation behind them.
 *
 *  (2) Batched entry points `tg_*` over DEVICE memory + a CUDA stream, called from PyTorch host code
 *      with tensor.data_ptr().  These are the hot path.
 *
 * Error convention for tg_*: return 0 on success, otherwise a cudaError_t (or -1 for bad arguments);
 * tg_last_error() returns a static description.  The 15 reference symbols keep the reference convention
 * (bool results, messages printed to stdout: go_env.cc:23,53,77, board.cc:449-460).
 *
 * State layout (n = 9: 48 B, n = 19: 128 B), little-endian u32 words:
 *   [black bitboard][white bitboard][pad to a multiple of 16 B][4 meta words]
 *   bitboard word = RPW rows of (n+1) bits each (guard bit per row); n=9: 3 x u32, n=19: 7 x u64
 *   meta[0] = last_move1 (i16) | last_move2 (i16) << 16        (board.h:43-44; -1 pass, -3 none)
 *   meta[1] = ko_location (i16) | ko_age (i16) << 16           (board.h:47,49)
 *   meta[2] = step_count (u16) | next_player (u8) << 16 | ko_color (u8) << 24   (board.h:41-42,48)
 *   meta[3] = terminated (bit 0)                               (go_env.h:16)
 * A batch of G states is stored as SQ = bytes/16 planes of uint4[G] (plane k holds the k-th 16 B of every
 * state) so that consecutive lanes read consecutive 128-bit words.
 *
 * Legal masks are "dense": bit a of a little-endian u32 string is action a; n=9: 4 x u32 (16 B),
 * n=19: 12 x u32 (48 B).  Bit P (PASS) is set only when no board point is legal -- the rule
 * GoEnv/environment.py:121-129 applies on top of go_env.cc:154-171 (which always appends PASS).
 */
#ifndef TRANSGO_B200_H_
#define TRANSGO_B200_H_
#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- (1) reference symbols, go_env.h:23-71 */
typedef void TgState;   /* opaque host blob, >= tg_state_bytes(board size) bytes */

bool    Init(int history_dim, int encode_dim, int max_step, float komi);     /* go_env.h:27  */
bool    Reset(TgState* state);                                               /* go_env.h:30  */
bool    Step(const TgState* state, TgState* next_state, int16_t action);     /* go_env.h:34  */
bool    Step_(TgState* state, int16_t action);                               /* go_env.h:35  */
bool    checkAction(const TgState* state, int16_t action);                   /* go_env.h:38  */
bool    isTerminated(const TgState* state);                                  /* go_env.h:41  */
bool    Encode(const TgState* state, float* encode_state);                   /* go_env.h:44  */
float   getScore(const TgState* state);                                      /* go_env.h:47  */
float   getTerritory(const TgState* state, float* territory);                /* go_env.h:50  */
int     getLegalAction(const TgState* state, int* actions);                  /* go_env.h:53  */
int     getLegalNoEye(const TgState* state, int* actions);                   /* go_env.h:56  */
void    Show(const TgState* state);                                          /* go_env.h:59  */
uint8_t getPlayer(const TgState* state);                                     /* go_env.h:62  */
int     getStep(const TgState* state);                                       /* go_env.h:65  */
void    getSubEncode(int* encode_state, int* sub_encode_state, int sub_board_size,
                     int encode_state_channels, int cut_num);                /* go_env.h:68  */

/* The reference fixes the board size at compile time (go_comm.h:20); here it is a process setting
 * for the single-state symbols above (default 9).  Returns false for an unsupported size. */
bool    tg_set_board_size(int n);

/* ---------------------------------------------------------------- (2) batched device entry points */
int         tg_version(void);
const char* tg_last_error(void);
int         tg_supported_board(int n);          /* 1 for n in {9, 19} */
int         tg_state_bytes(int n);              /* 48 / 128 */
int         tg_legal_words(int n);              /* u32 words of a dense legal mask, padded to 16 B: 4 / 12 */

/* All pointers are device pointers; `stream` is a cudaStream_t (NULL = default stream).  The playout entry points also accept
 * PINNED HOST pointers for `states`, `trace` and `score_out` (zero-copy: the kernel reads / writes them across PCIe once).
 * `tile` = lanes cooperating on one game: 32 (lowest latency, small batches) ... 1 (highest throughput,
 * very large batches); 0 lets the library choose from G (for playouts: the incremental engine below, measured fastest at
 * every batch size).  tg_env_playout also takes TG_TILE_ROWS: one board ROW per lane, 31 / (n + 1) games side by side in a
 * warp (csrc/tg_rows.cuh) -- an independent formulation kept as a cross-check. */
#define TG_TILE_ROWS (-1)
/* tg_env_playout: OR into a tile width of 4, 8 or 16 to run the tiles that share a warp in lock step (warp-uniform control
 * flow: one warp never executes two games' branches one after the other) */
#define TG_TILE_UNIFORM 64
/* tg_env_playout: OR into a tile width of 1, 4, 8, 16 or 32 to run the incremental engine (csrc/tg_inc.cuh): dense bit-string
 * boards and an atari set kept in registers across plies; the lanes of a tile take the <= 5 groups around the new stone */
#define TG_TILE_INC 128

/* board.cc:13-26 + go_env.cc:34-41 */
int tg_env_reset(int n, void* states, int64_t G, void* stream);

/* go_env.cc:44-80 for every game: states are updated in place.
 *   actions   int16[G]         0..P-1 point, P or -1 pass, -2 resign
 *   legal_out u32[G][LW]       dense legal mask of the NEW position (may be NULL)
 *   done_out  u8[G]            the bool Step returns (may be NULL)
 *   acc_out   u8[G]            0 where the reference prints "Fail to Step: invalid action" (may be NULL) */
int tg_env_step(int n, void* states, int64_t G, const int16_t* actions, uint32_t* legal_out,
                uint8_t* done_out, uint8_t* acc_out, int max_step, int tile, void* stream);

/* go_env.cc:154-171 (no_eye = 0) / :178-188 (no_eye = 1) as dense masks */
/* tg_env_step on the incremental engine (csrc/tg_inc.cuh): the atari set of every position travels beside it in `atari`
 * (tg_env_atari_words(n) u32 per game: 4 for 9x9, 16 for 19x19; start from all zeros) so that a step does not re-derive it.
 * The record carries a digest of the stones it belongs to: a record that does not match its position (never initialised, or the
 * position was changed by something else) is rebuilt inside the same launch, so the buffer is a cache and can never make a result
 * wrong.  Same arguments, outputs and semantics as tg_env_step otherwise. */
int tg_env_atari_words(int n);
int tg_env_step_inc(int n, void* states, uint32_t* atari, int64_t G, const int16_t* actions, uint32_t* legal_out, uint8_t* done_out,
                    uint8_t* accepted_out, int max_step, void* stream);

int tg_env_legal(int n, const void* states, int64_t G, uint32_t* legal_out, int no_eye, int tile, void* stream);

/* go_env.cc:96-115 with encode_dim 10: planes float32[G][10][n][n], written in place (e.g. the net's input) */
int tg_env_encode(int n, const void* states, int64_t G, float* planes, int tile, void* stream);

/* go_env.cc:126-149: score float32[G] = area(black) - area(white) - komi; territory float32[G][P] in
 * {+1 black, 0 dame, -1 white} (may be NULL) */
int tg_env_score(int n, const void* states, int64_t G, float komi, float* score, float* territory,
                 int tile, void* stream);

/* SURVEY.md 8d C2 workload: every game plays up to `plies` moves in ONE launch.  Move rule: the k-th legal
 * board point in ascending order, k = (mix32(seed, game_id0 + g, step_count) * n_legal) >> 32; PASS iff none
 * (environment.py:121-129).  no_eye = 1 draws from getLegalNoEye's points instead.
 *   trace_actions int16[plies][G]  chosen action per ply, -3 after the game ended (may be NULL)
 *   plies_done    u64[1]           += number of moves actually played (may be NULL) */
int tg_env_playout(int n, void* states, int64_t G, int plies, int max_step, uint64_t seed, int64_t game_id0,
                   int no_eye, int16_t* trace_actions, unsigned long long* plies_done, int tile, void* stream);

/* The same launch with the two steps that frame a playout folded in (one launch instead of three for "reset, play to the end,
 * score"): from_reset != 0 starts every game from the empty board (tg_env_reset) without reading `states` first; score_out
 * (float32[G], may be NULL) receives getScore of the final positions (area(black) - area(white) - komi, go_env.cc:126-130). */
int tg_env_playout_scored(int n, void* states, int64_t G, int plies, int max_step, uint64_t seed, int64_t game_id0,
                          int no_eye, int16_t* trace_actions, unsigned long long* plies_done, int from_reset,
                          float* score_out, float komi, int tile, void* stream);

/* the same with separate input and output positions (states_in is only read, states_out only written; they may be the same
 * buffer): a host that keeps its start positions in pinned memory needs no refill between launches.  score_out may be NULL. */
int tg_env_playout_io(int n, const void* states_in, void* states_out, int64_t G, int plies, int max_step, uint64_t seed, int64_t game_id0,
                      int no_eye, int16_t* trace_actions, unsigned long long* plies_done, float* score_out, float komi, int tile,
                      void* stream);

/* the layout tg_env_playout uses for (n, G, tile): TG_TILE_ROWS or the number of lanes per game (tile = 0: measured choice) */
int tg_env_playout_tile(int n, int64_t G, int tile);
/* diagnostics: games per warp of the incremental engine's tiled layouts (0 = fill the warp: 32 / lanes per game); returns the
 * previous setting */
int tg_env_debug_games_per_warp(int gw);

/* board.cc:1166-1271: float32[B][C][n][n] -> float32[B][cuts][C][sub][sub] */
int tg_sub_encode(int n, const float* enc, int64_t B, int channels, int sub, int cuts, float* out, void* stream);

/* ---------------------------------------------------------------- (3) batched MCTS (self_play.py: WP_MCTS :575-875)
 *
 * G independent searches, one per game, advanced in lock-step "waves" exactly like WP_MCTS.run (self_play.py:607-654):
 * up to `readouts` leaves per game are selected one after another (PUCT + variance bonus + ongoing-simulation counts,
 * :706-725, :767-774), terminal leaves are scored and backed up at once (:638-642), the remaining leaves are encoded
 * straight into the net's input tensor, and after the net ran they are expanded and backed up (:727-764).
 * The tree is a per-game arena in the caller-provided workspace (device memory): 16-byte "hot" edge records
 * {prior f32, W f32, var f32, N:16|ons:7|action:9} scanned with 128-bit loads, 8-byte "cold" records {mean f32, child},
 * and node records {state, edge base, edge count}.  Re-rooting keeps the chosen subtree (self_play.py:857-860) by
 * compacting it into the arena's other half.
 *
 * The library is stateless: every call takes the same TgMctsConfig and workspace pointer.  All other pointers are
 * device pointers owned by the caller (PyTorch tensors).  Calls on one stream are ordered; nothing synchronises.
 */
typedef struct TgMctsConfig {
    int32_t  n;                 /* board size: 9 or 19 */
    int32_t  num_games;         /* G */
    int32_t  readouts;          /* configure.py:33 parallel_readouts (leaves per wave and game), <= 8 */
    int32_t  num_simulation;    /* configure.py:29 */
    int32_t  wu_loss;           /* configure.py:32 */
    int32_t  max_step;          /* environment.py:37 */
    int32_t  node_cap;          /* expanded nodes per game and arena half */
    int32_t  edge_cap;          /* edges per game and arena half */
    int32_t  max_depth;         /* longest recorded path (edges) */
    int32_t  tie_mode;          /* 0: first maximal child in ascending action order; 1: hashed draw (see tg_mcts_select) */
    double   c_puct1;           /* configure.py:26 */
    double   c_puct2;           /* configure.py:27 */
    float    komi;              /* environment.py:38 */
    int32_t  reserved0;
    uint64_t tie_seed;          /* hashed tie-break: index = (mix32(tie_seed, game_id0 + g, tie_counter++) * n_ties) >> 32 */
    int64_t  game_id0;          /* global id of game 0 of this shard (multi-GPU: identical results for any sharding) */
} TgMctsConfig;

/* per-game status bits reported by tg_mcts_status */
#define TG_MCTS_ACTIVE        1u    /* game not finished */
#define TG_MCTS_NEED_EXPAND   2u    /* root has no children yet (new game, or re-rooted onto an unexpanded child) */
#define TG_MCTS_NOISY         4u    /* root priors carry Dirichlet noise (float64, self_play.py:90-95) */
#define TG_MCTS_ERR_NODES    16u    /* arena overflow: raise node_cap */
#define TG_MCTS_ERR_EDGES    32u    /* arena overflow: raise edge_cap */
#define TG_MCTS_ERR_DEPTH    64u    /* path longer than max_depth */
#define TG_MCTS_ERR_COUNT   128u    /* visit count above 65535 */
#define TG_MCTS_ERR_NAN     512u    /* a node without a finite selectable score: the evaluator returned NaN (256 = bad action, see tg_mcts_advance) */
#define TG_MCTS_ERR_ANY     (16u | 32u | 64u | 128u | 512u)

int64_t tg_mcts_workspace_bytes(const TgMctsConfig* cfg);
/* every game: empty board, empty tree, ACTIVE | NEED_EXPAND (WP_MCTS.reset_root before the net call, :595-597) */
int tg_mcts_init(const TgMctsConfig* cfg, void* ws, void* stream);
/* re-initialise the games whose mask byte is non-zero (start the next self-play game in that slot) */
int tg_mcts_reset_games(const TgMctsConfig* cfg, void* ws, const uint8_t* mask, void* stream);
/* encode10 of every root state -> obs_root float32 [G][10][n][n] (computeValuePolicy([root.state]), :598) */
int tg_mcts_root_obs(const TgMctsConfig* cfg, void* ws, float* obs_root, void* stream);
/* expand roots flagged NEED_EXPAND with the RAW policy over legal actions and W = -value (:599-605, :862-870) */
int tg_mcts_root_expand(const TgMctsConfig* cfg, void* ws, const float* policy_root /*[G][A]*/,
                        const float* value_root /*[G]*/, void* stream);
/* start of get_action_probs (:657-663): noise float64 [G][A] in root-edge order (first K entries used per game; NULL = no
 * noise, is_selfplay False); records N0 = root visits so far */
int tg_mcts_begin_move(const TgMctsConfig* cfg, void* ws, const double* noise, void* stream);
/* wave, first half: selection + terminal backups + leaf step/legal/encode.  obs float32 [G*readouts][10][n][n];
 * rows of unused slots are left untouched. */
int tg_mcts_wave_begin(const TgMctsConfig* cfg, void* ws, float* obs, void* stream);
/* number of active games still below N0 + num_simulation (the loop condition of get_action_probs, :660-664) -> out int32[1];
 * a separate small launch so that the waves themselves carry no counter traffic (ask every few waves) */
int tg_mcts_pending(const TgMctsConfig* cfg, void* ws, int32_t* out, void* stream);
/* test hook: 1 = run the selection of every wave with one warp per game in the reference's descent order (the default kernel
 * pipelines the descents of a wave over one warp each and falls back to that order by itself where needed); returns the old value */
int tg_mcts_debug_sequential_select(int on);
/* diagnostics: device buffer of 4 x u64 per game (start, selection done, end in globaltimer ns; deepest path) written by every
 * pipelined selection launch; NULL stops the recording */
int tg_mcts_debug_times(unsigned long long* times);
/* diagnostics: 1 / 0 (default) = speculative L2 prefetch of the most-visited child's edge records during selection (never changes
 * a result) */
int tg_mcts_debug_prefetch(int on);

/* wave, second half: policy float32 [G*readouts][A] (softmaxed, model.py:111), value float32 [G*readouts];
 * pending (may be NULL; prefer tg_mcts_pending) is INCREMENTED by the number of active games still below N0 + num_simulation */
int tg_mcts_wave_end(const TgMctsConfig* cfg, void* ws, const float* policy, const float* value, int32_t* pending,
                     void* stream);
/* root statistics in ACTION order, zero where the action has no edge.  Any pointer may be NULL.
 *   counts i32 [G][A]  (:666)   W f32 [G][A]   var f32 [G][A]   prior f32 [G][A]   root_n i32 [G] = root total_visit_count
 *   root_k i32 [G] = number of root edges   step i32 [G] = getStep(root.state) */
int tg_mcts_root_stats(const TgMctsConfig* cfg, void* ws, int32_t* counts, float* W, float* var, float* prior,
                       int32_t* root_n, int32_t* root_k, int32_t* step, void* stream);
/* update_with_action (:857-872): play actions int16 [G] on the root states, re-root onto the chosen child keeping its
 * subtree; done_out u8 [G] as returned by env.step.  Finished / inactive games are left alone. */
int tg_mcts_advance(const TgMctsConfig* cfg, void* ws, const int16_t* actions, uint8_t* done_out, void* stream);
/* status u32 [G] (TG_MCTS_* bits), sims u64 [1] += simulations (net-evaluated leaves + terminal backups) since last call;
 * either may be NULL */
int tg_mcts_status(const TgMctsConfig* cfg, void* ws, uint32_t* status, unsigned long long* sims, void* stream);
/* root game states as a batch in the layout of section (2): returns the device pointer inside the workspace */
void* tg_mcts_root_states(const TgMctsConfig* cfg, void* ws);
/* selection statistics since the previous call, summed over the games (and reset): out3[0] descents, out3[1] edges walked
 * (sum of path lengths), out3[2] children scanned by PUCT -- mean depth and fan-out for the traffic model of SURVEY.md 8d.
 * out3 must be zeroed by the caller. */
int tg_mcts_search_stats(const TgMctsConfig* cfg, void* ws, unsigned long long* out3, void* stream);

/* final labels of finished games (self_play.py:931-937): winner u8 [G] (1|2), score f32 [G], territory f32 [G][P] */
int tg_mcts_outcome(const TgMctsConfig* cfg, void* ws, uint8_t* winner, float* score, float* territory, void* stream);
/* a device-side synthetic policy/value provider for tree-only throughput runs (bench.py): policy[row] = normalised
 * heavy-tailed hash weights of (seed, row, wave), value in (-1, 1).  NOT the reference net. */
int tg_synth_provider(int n, int64_t rows, uint64_t seed, uint64_t wave, float* policy, float* value, void* stream);
/* same, with the wave index read from (and then incremented in) device memory, so that the call can sit inside a replayed CUDA
 * graph.  wave_counter: u64[2], zero-initialised ([0] the wave index, [1] scratch: the block that finishes last advances [0]) */
int tg_synth_provider_dev(int n, int64_t rows, uint64_t seed, unsigned long long* wave_counter, float* policy, float* value, void* stream);

/* ---------------------------------------------------------------- (3b) trajectory assembly + root noise
 * self_play.py:931-965 for a batch of games in ONE launch: every recorded position (t < length[g]) of every finished game
 * (fin[g] != 0; NULL = all) is labelled -- z = +1 iff the mover is the winner, ownership = territory from the mover's view --
 * expanded with the eight board symmetries in the reference's order (for i in 1..4: rot90 by i, then its fliplr;
 * symmetries = 1: as played) and written at row (ring_index + r) % capacity of the output tensors, r numbering rows
 * game-major / position / symmetry.  hist_obs float32 [G][T][C][n][n], hist_pi float32 [G][T][A], hist_player int32 [G][T]
 * (1 black, 2 white), length int64 [G] (NULL = one position per game), winner uint8 [G], territory float32 [G][P] (tg_mcts_outcome);
 * out_obs [capacity][C][n][n], out_pi [capacity][A], out_z [capacity], out_own [capacity][P]; rows_out int64[1] (may be NULL)
 * receives the number of rows written. */
int tg_traj_assemble(int n, int channels, int64_t G, int64_t T, const float* hist_obs, const float* hist_pi, const int32_t* hist_player,
                     const int64_t* length, const uint8_t* fin, const uint8_t* winner, const float* territory, float* out_obs,
                     float* out_pi, float* out_z, float* out_own, int64_t capacity, int64_t ring_index, int symmetries, int64_t* rows_out,
                     void* stream);
/* Node_V.dirichlet_prior's noise (self_play.py:90-95) on the device: noise float64 [G][A], row g ~ Dirichlet(alpha) over its
 * first root_k[g] entries (root-edge order), zeros beyond.  Counter-based (Philox4x32-10 keyed by seed; counter = child, global
 * game id, move, draw), so a shard of games draws what the full batch would. */
int tg_dirichlet(double* noise, const int32_t* root_k, int64_t G, int A, double alpha, uint64_t seed, uint32_t move, int64_t game_id0,
                 void* stream);

/* ---------------------------------------------------------------- (4) inference epilogue of the policy/value net
 * One pass over NHWC bf16 activations [rows][channels] (channels % 8 == 0):
 *     sum = x + alpha * (y + ybias[c])   (y may be NULL: sum = x; ybias may be NULL)   -> sum_out (may be NULL)
 *     act = relu(sum * scale[c] + shift[c])                 -> act_out (may be NULL)
 * i.e. a residual junction of model.py:208-262 / :286-314 with the next eval-mode BatchNorm folded to scale/shift (f32). */
int tg_net_add_bn_relu(const void* x, const void* y, float alpha, const float* ybias, const float* scale, const float* shift,
                       void* sum_out, void* act_out, int64_t rows, int channels, void* stream);
/* the same junction on float32 activations (channels % 4 == 0): the folded float32 schedule of fused_net.py */
int tg_net_add_bn_relu_f32(const float* x, const float* y, float alpha, const float* ybias, const float* scale, const float* shift,
                           float* sum_out, float* act_out, int64_t rows, int channels, void* stream);
/* diagnostics: 1 walks the junction's tensors back to front (the tail of y is what L2 may still hold), 0 (default) front to back;
 * returns the previous setting */
int tg_net_debug_reverse(int on);


/* Fused single-head spatial self-attention block of model.py:286-314 in inference, NHWC bf16:
 *   qkv  bf16 [B][P][dq + dq + C]   the three 1x1 convolutions (query | key | value) of x, one GEMM upstream
 *   out  bf16 [B][P][C] = relu((x + gamma * o) * scale[c] + shift[c]),  o[j] = sum_i softmax_j(q_i . k_j) * v_i
 * Supports P <= 368 (one kernel with scores in registers up to 96 points, a two-pass one above: 19x19), dq == 32, C == 128;
 * returns -1 otherwise (callers then use the library GEMM path). */
int tg_net_attention(const void* qkv, const void* x, float gamma, const float* scale, const float* shift, void* out,
                     int64_t B, int P, int dq, int channels, void* stream);
/* ... with the second output out2 = relu(out * scale2[c] + shift2[c]) of tg_net_attention_fused2 (out2 may be NULL) */
int tg_net_attention2(const void* qkv, const void* x, float gamma, const float* scale, const float* shift, void* out,
                      const float* scale2, const float* shift2, void* out2, int64_t B, int P, int dq, int channels, void* stream);


/* The stem of the net (model.py:43, :81): 3x3 convolution from the feature planes to 128 channels, folded BatchNorm, ReLU, 9x9 boards.
 *   obs float32 [B][cin][81] (the planes as the leaf kernels write them, cin <= 12)    w bf16 [128][9][16] = [output channel][tap][input
 *   channel, zero beyond cin]    bias f32 [128]    out bf16 [B][81][128] (NHWC) = relu(conv3x3(obs) + bias)
 *   out2 (may be NULL) = relu(out * scale2[c] + shift2[c]): the opening activation of the first residual block (model.py:216-217) */
int tg_net_stem(const float* obs, int cin, const void* w, const float* bias, const float* scale2, const float* shift2, void* out, void* out2,
                int64_t B, int board, int channels, void* stream);

/* The head convolutions of the net (model.py:93-94, :105-106): 3x3, 128 -> nout (<= 4) channels, folded BatchNorm, ReLU, 9x9 boards.
 *   x bf16 [B][81][128]   w bf16 [8][9][128] = [output channel][tap ky * 3 + kx][input channel] (rows >= nout zero)   bias f32 [8]
 *   out bf16 [B][81][nout] = relu(conv3x3(x) + bias)                                   (one pass over x: mma.sync, csrc/tg_net.cu) */
int tg_net_conv3x3_head(const void* x, const void* w, const float* bias, void* out, int64_t B, int board, int channels, int nout,
                        void* stream);

/* The same block INCLUDING the q | k | v projection, as one tcgen05 kernel (csrc/tg_net_tc.cu; 9x9 boards: P <= 96):
 *   x     bf16 [B][P][C]          wqkv bf16 [dq + dq + C][C] (query | key | gamma-scaled value rows, the three 1x1 convolutions
 *   of model.py:290-292 stacked)   bqkv bf16 [dq + dq + C]
 *   out   bf16 [B][P][C] = relu((x + gamma * o) * scale[c] + shift[c]),  q|k|v = x . wqkv^T + bqkv (rounded to bf16),
 *                           o[j] = sum_i softmax_j(q_i . k_j) * v_i
 * q, k, v and the scores stay in shared / tensor memory.  debug (may be NULL): float32 [128*192 + 128*96 + 128*128], the
 * accumulator tiles of position 0 (projection + bias, scores, o) for stage-by-stage tests.  Returns -1 for unsupported shapes. */
int tg_net_attention_fused(const void* x, const void* wqkv, const void* bqkv, float gamma, const float* scale, const float* shift,
                           void* out, int64_t B, int P, int dq, int channels, float* debug, void* stream);
/* ... with a second output out2 = relu(out * scale2[c] + shift2[c]) (out2 may be NULL): the BatchNorm + ReLU that opens the residual
 * block after the attention block (model.py:216-217), so that no separate pass over `out` is needed */
int tg_net_attention_fused2(const void* x, const void* wqkv, const void* bqkv, float gamma, const float* scale, const float* shift,
                            void* out, const float* scale2, const float* shift2, void* out2, int64_t B, int P, int dq, int channels,
                            float* debug, void* stream);

#ifdef __cplusplus
}
#endif
#endif
