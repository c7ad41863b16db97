// tg_compat.cu -- the reference's 15 extern "C" symbols (GoEnv/cpp_src/go_env.h:23-71) for ONE host-side state.
// Each call stages the blob to the GPU, runs the batched sm_100a kernels of tg_env.cu with G = 1 and copies the
// answer back.  There is deliberately no CPU rules engine here: without a CUDA device these calls fail loudly.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include "../../include/transgo_b200.h"
#include "tg_board.cuh"

namespace {

// process-global settings, as in go_env.cc:9-12 (same defaults); board size added (go_comm.h:20 is a constant there)
int g_history_dim = 4;
int g_encode_dim = 10;
int g_max_step = 300;
float g_komi = 7.5f;
int g_n = 9;

std::mutex g_mu;
struct Scratch {
    bool ready = false;
    cudaStream_t st = nullptr;
    void* d_state = nullptr; int16_t* d_action = nullptr; uint32_t* d_legal = nullptr;
    uint8_t* d_flags = nullptr; float* d_planes = nullptr; float* d_score = nullptr; float* d_terr = nullptr;
    float* d_sub_in = nullptr; float* d_sub_out = nullptr; size_t sub_in_cap = 0, sub_out_cap = 0;
} S;

void die(const char* what, cudaError_t e) {
    fprintf(stderr, "transgo_b200: %s failed: %s -- this library has no CPU fallback, a CUDA device is required\n",
            what, cudaGetErrorString(e));
    abort();
}
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) die(#x, e_); } while (0)
#define TGCK(x) do { int e_ = (x); if (e_ != 0) { fprintf(stderr, "transgo_b200: %s -> %s\n", #x, tg_last_error()); abort(); } } while (0)

void ensure() {
    if (S.ready) return;
    CK(cudaStreamCreateWithFlags(&S.st, cudaStreamNonBlocking));
    CK(cudaMalloc(&S.d_state, 256));
    CK(cudaMalloc(&S.d_action, 16));
    CK(cudaMalloc(&S.d_legal, 64));
    CK(cudaMalloc(&S.d_flags, 16));
    CK(cudaMalloc(&S.d_planes, 10 * 361 * sizeof(float)));
    CK(cudaMalloc(&S.d_score, 16));
    CK(cudaMalloc(&S.d_terr, 361 * sizeof(float)));
    S.ready = true;
}
int sbytes() { return tg_state_bytes(g_n); }
void up(const void* blob) { CK(cudaMemcpyAsync(S.d_state, blob, sbytes(), cudaMemcpyHostToDevice, S.st)); }
void sync() { CK(cudaStreamSynchronize(S.st)); }

const uint32_t* meta(const void* blob) { return reinterpret_cast<const uint32_t*>(blob) + sbytes() / 4 - 4; }

int legal_list(const void* state, int* actions, int no_eye) {
    std::lock_guard<std::mutex> lk(g_mu);
    ensure(); up(state);
    TGCK(tg_env_legal(g_n, S.d_state, 1, S.d_legal, no_eye, 32, S.st));
    uint32_t d[16];
    const int lw = tg_legal_words(g_n);
    CK(cudaMemcpyAsync(d, S.d_legal, lw * 4, cudaMemcpyDeviceToHost, S.st));
    sync();
    const int P = g_n * g_n;
    int k = 0;
    for (int a = 0; a < P; ++a) if ((d[a >> 5] >> (a & 31)) & 1u) actions[k++] = a;
    actions[k++] = P;                       // go_env.cc:167-169: PASS is always appended at the C level
    return k;
}

}  // namespace

extern "C" {

bool tg_set_board_size(int n) {
    if (!tg_supported_board(n)) return false;
    std::lock_guard<std::mutex> lk(g_mu);
    g_n = n;
    return true;
}

bool Init(int history_dim, int encode_dim, int max_step, float komi) {      // go_env.cc:20-31
    if (history_dim > 1) { printf("history_dim is too large\n"); return false; }
    std::lock_guard<std::mutex> lk(g_mu);
    g_history_dim = history_dim; g_encode_dim = encode_dim; g_max_step = max_step; g_komi = komi;
    return true;
}

bool Reset(TgState* state) {                                                // go_env.cc:34-41
    std::lock_guard<std::mutex> lk(g_mu);
    ensure();
    TGCK(tg_env_reset(g_n, S.d_state, 1, S.st));
    CK(cudaMemcpyAsync(state, S.d_state, sbytes(), cudaMemcpyDeviceToHost, S.st));
    sync();
    return true;
}

bool Step_(TgState* state, int16_t action) {                                // go_env.cc:51-80
    if (isTerminated(state)) { printf("Fail to Step: game is done!!\n\n"); return true; }
    std::lock_guard<std::mutex> lk(g_mu);
    ensure(); up(state);
    CK(cudaMemcpyAsync(S.d_action, &action, sizeof(action), cudaMemcpyHostToDevice, S.st));
    TGCK(tg_env_step(g_n, S.d_state, 1, S.d_action, nullptr, S.d_flags, S.d_flags + 1, g_max_step, 32, S.st));
    uint8_t fl[2];
    CK(cudaMemcpyAsync(fl, S.d_flags, 2, cudaMemcpyDeviceToHost, S.st));
    CK(cudaMemcpyAsync(state, S.d_state, sbytes(), cudaMemcpyDeviceToHost, S.st));
    sync();
    if (!fl[1]) {
        printf("Invalid move on Coord %d\n", (int)action);
        printf("Fail to Step: invalid action\n\n");
    }
    return fl[0] != 0;
}

bool Step(const TgState* state, TgState* next_state, int16_t action) {     // go_env.cc:44-47
    memcpy(next_state, state, sbytes());
    return Step_(next_state, action);
}

bool checkAction(const TgState* state, int16_t action) {                    // go_env.cc:84-88 (no P -> PASS remap here)
    if (action == tg::kPass || action == tg::kResign) return true;
    const int P = g_n * g_n;
    bool ok = false;
    if (action >= 0 && action < P) {
        int acts[362];
        int k = legal_list(state, acts, 0);
        for (int i = 0; i + 1 < k; ++i) ok |= (acts[i] == action);
    }
    if (!ok) printf("Invalid move on Coord %d\n", (int)action);
    return ok;
}

bool isTerminated(const TgState* state) { return (meta(state)[3] & 1u) != 0; }    // go_env.cc:91-93

bool Encode(const TgState* state, float* encode_state) {                    // go_env.cc:96-115
    if (g_encode_dim != 9 && g_encode_dim != 10) {
        printf("Fail to Encode: encode_dim %d not implemented\n", g_encode_dim);
        return false;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    ensure(); up(state);
    TGCK(tg_env_encode(g_n, S.d_state, 1, S.d_planes, 32, S.st));
    // encode9 is the first nine planes of encode10 (board_feature.cc:199-223)
    CK(cudaMemcpyAsync(encode_state, S.d_planes, (size_t)g_encode_dim * g_n * g_n * sizeof(float), cudaMemcpyDeviceToHost, S.st));
    sync();
    return true;
}

float getScore(const TgState* state) {                                      // go_env.cc:126-130
    std::lock_guard<std::mutex> lk(g_mu);
    ensure(); up(state);
    TGCK(tg_env_score(g_n, S.d_state, 1, g_komi, S.d_score, nullptr, 32, S.st));
    float sc;
    CK(cudaMemcpyAsync(&sc, S.d_score, 4, cudaMemcpyDeviceToHost, S.st));
    sync();
    return sc;
}

float getTerritory(const TgState* state, float* territory) {                // go_env.cc:136-149
    std::lock_guard<std::mutex> lk(g_mu);
    ensure(); up(state);
    TGCK(tg_env_score(g_n, S.d_state, 1, g_komi, S.d_score, S.d_terr, 32, S.st));
    float sc;
    CK(cudaMemcpyAsync(&sc, S.d_score, 4, cudaMemcpyDeviceToHost, S.st));
    CK(cudaMemcpyAsync(territory, S.d_terr, (size_t)g_n * g_n * sizeof(float), cudaMemcpyDeviceToHost, S.st));
    sync();
    return sc;
}

int getLegalAction(const TgState* state, int* actions) { return legal_list(state, actions, 0); }    // go_env.cc:154-171
int getLegalNoEye(const TgState* state, int* actions) { return legal_list(state, actions, 1); }     // go_env.cc:178-188

uint8_t getPlayer(const TgState* state) { return (uint8_t)((meta(state)[2] >> 16) & 0xff); }       // go_env.cc:208-210
int getStep(const TgState* state) { return (int)(meta(state)[2] & 0xffff); }                       // go_env.cc:213-215

void Show(const TgState* state) {                                           // go_env.cc:192-205, board.cc:32-44 (formatting only)
    const int n = g_n, stride = n + 1;
    const uint32_t* u = reinterpret_cast<const uint32_t*>(state);
    auto stone = [&](int c, int x, int y) -> bool {
        if (n == 9) { const int rpw = 3; uint32_t w = u[c * 3 + y / rpw]; return (w >> ((y % rpw) * stride + x)) & 1u; }
        const int rpw = 3, W = 7; int wi = y / rpw;
        uint64_t w = (uint64_t)u[(c * W + wi) * 2] | ((uint64_t)u[(c * W + wi) * 2 + 1] << 32);
        return (w >> ((y % rpw) * stride + x)) & 1ull;
    };
    printf("     A B C D E F G F I J K L M N O P Q R S\n");
    printf("    ——————————————————————————————————————\n");
    for (int y = 0; y < n; ++y) {
        if (y < 9) printf(" ");
        printf("%d | ", y + 1);
        for (int x = 0; x < n; ++x) printf("%s ", stone(0, x, y) ? "X" : (stone(1, x, y) ? "O" : "."));
        printf("\n");
    }
    const int player = getPlayer(state);
    printf("step: %d\n", getStep(state));
    printf("next_player: %s(%s)\n", player == 1 ? "Black" : "White", player == 1 ? "X" : "O");
    printf(isTerminated(state) ? "done: true\n" : "done: false\n");
    printf("\n");
}

void getSubEncode(int* encode_state, int* sub_encode_state, int sub_board_size, int encode_state_channels, int cut_num) {   // go_env.cc:118-121
    std::lock_guard<std::mutex> lk(g_mu);
    ensure();
    const size_t in_b = (size_t)encode_state_channels * g_n * g_n * 4;
    const size_t out_b = (size_t)cut_num * encode_state_channels * sub_board_size * sub_board_size * 4;
    if (in_b > S.sub_in_cap) { if (S.d_sub_in) cudaFree(S.d_sub_in); CK(cudaMalloc(&S.d_sub_in, in_b)); S.sub_in_cap = in_b; }
    if (out_b > S.sub_out_cap) { if (S.d_sub_out) cudaFree(S.d_sub_out); CK(cudaMalloc(&S.d_sub_out, out_b)); S.sub_out_cap = out_b; }
    CK(cudaMemcpyAsync(S.d_sub_in, encode_state, in_b, cudaMemcpyHostToDevice, S.st));   // 4-byte bit copy, as the reference does with int*
    TGCK(tg_sub_encode(g_n, S.d_sub_in, 1, encode_state_channels, sub_board_size, cut_num, S.d_sub_out, S.st));
    CK(cudaMemcpyAsync(sub_encode_state, S.d_sub_out, out_b, cudaMemcpyDeviceToHost, S.st));
    sync();
}

}  // extern "C"
