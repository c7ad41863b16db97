/* oracle/go_rules.c -- TEST INFRASTRUCTURE ONLY (see go_rules.h).
 *
 * A deliberately simple array + flood-fill restatement of what the reference engine makes
 * observable.  The reference keeps incremental block lists and liberty counters
 * (board.cc:217-428, 546-653); their net effect is "standard capture rules with exact
 * liberty counts" (SURVEY.md 8a-2), which is what is restated here from first principles.
 * Every function names the reference lines whose behaviour it follows.
 */
#include "go_rules.h"
#include <string.h>

static const int DX[4] = {-1, 0, 1, 0};
static const int DY[4] = {0, -1, 0, 1};
static const int GX[4] = {-1, -1, 1, 1};   /* diagonals, go_comm.h:47-48 */
static const int GY[4] = {-1, 1, 1, -1};

#define P_OF(s) ((s)->n * (s)->n)
#define OPP(c)  ((uint8_t)(3 - (c)))

static int on_xy(const OgState* s, int x, int y) { return x >= 0 && x < s->n && y >= 0 && y < s->n; }

/* Flood one group from p.  mark[] gets `tag` on every stone of the group; lib[] gets `tag`
 * on every liberty.  Returns liberties, *stones = group size. */
static int group_scan(const OgState* s, int p, int* mark, int* lib, int tag, int* stones, int* members) {
    int n = s->n, sp = 0, cnt = 0, libs = 0;
    int stack[OG_MAX_P];
    uint8_t col = s->color[p];
    stack[sp++] = p; mark[p] = tag;
    while (sp) {
        int c = stack[--sp];
        if (members) members[cnt] = c;
        cnt++;
        for (int d = 0; d < 4; ++d) {
            int x = c % n + DX[d], y = c / n + DY[d];
            if (!on_xy(s, x, y)) continue;
            int q = y * n + x;
            if (s->color[q] == OG_EMPTY) { if (lib[q] != tag) { lib[q] = tag; libs++; } }
            else if (s->color[q] == col && mark[q] != tag) { mark[q] = tag; stack[sp++] = q; }
        }
    }
    if (stones) *stones = cnt;
    return libs;
}

typedef struct { int libs[OG_MAX_P]; int size[OG_MAX_P]; int gid[OG_MAX_P]; } Groups;

/* liberties / size / group id of every stone (the observable content of Block, board.h:18-23) */
static void analyse(const OgState* s, Groups* g) {
    int P = P_OF(s);
    int mark[OG_MAX_P], lib[OG_MAX_P], members[OG_MAX_P];
    for (int i = 0; i < P; ++i) { mark[i] = -1; lib[i] = -1; g->libs[i] = 0; g->size[i] = 0; g->gid[i] = -1; }
    for (int p = 0; p < P; ++p) {
        if (s->color[p] == OG_EMPTY || g->gid[p] >= 0) continue;
        int cnt = 0;
        int l = group_scan(s, p, mark, lib, p, &cnt, members);
        for (int i = 0; i < cnt; ++i) { g->libs[members[i]] = l; g->size[members[i]] = cnt; g->gid[members[i]] = p; }
    }
}

/* board.cc:90-158  StoneLibertyAnalysis + isSuicideMove */
static int is_suicide(const OgState* s, const Groups* g, int p, uint8_t player) {
    int n = s->n;
    for (int d = 0; d < 4; ++d) {
        int x = p % n + DX[d], y = p / n + DY[d];
        if (!on_xy(s, x, y)) continue;
        int q = y * n + x;
        if (s->color[q] == OG_EMPTY) return 0;                       /* condition 1 */
        if (s->color[q] == player) { if (g->libs[q] > 1) return 0; } /* condition 2 */
        else if (g->libs[q] == 1) return 0;                          /* condition 3 */
    }
    return 1;
}

/* board.cc:198-200 */
static int ko_violation(const OgState* s, int p, uint8_t player) {
    return s->ko_loc == p && s->ko_age == 0 && s->ko_color == player;
}

/* board.cc:437-464 TryPlay for a board point */
static int point_legal(const OgState* s, const Groups* g, int p, uint8_t player) {
    if (p < 0 || p >= P_OF(s)) return 0;
    if (s->color[p] != OG_EMPTY) return 0;
    if (ko_violation(s, p, player)) return 0;
    return !is_suicide(s, g, p, player);
}

/* board.cc:665-714 isEye / isFakeEye / isTrueEye */
static int true_eye(const OgState* s, int p, uint8_t player) {
    int n = s->n;
    if (s->color[p] != OG_EMPTY) return 0;
    for (int d = 0; d < 4; ++d) {
        int x = p % n + DX[d], y = p / n + DY[d];
        if (on_xy(s, x, y) && s->color[y * n + x] != player) return 0;
    }
    int opp = 0, wall = 0;
    for (int d = 0; d < 4; ++d) {
        int x = p % n + GX[d], y = p / n + GY[d];
        if (!on_xy(s, x, y)) wall++;
        else if (s->color[y * n + x] == OPP(player)) opp++;
    }
    int fake = (wall > 0 && opp >= 1) || (wall == 0 && opp >= 2);
    return !fake;
}

void og_reset(OgState* s, int n, int max_step, float komi) {   /* board.cc:13-26, go_env.cc:34-41 */
    memset(s, 0, sizeof(*s));
    s->n = n; s->max_step = max_step; s->komi = komi;
    s->next_player = OG_BLACK;
    s->last1 = s->last2 = s->ko_loc = OG_INVALID;
    s->step_count = 1;
}

int og_sizeof_state(void) { return (int)sizeof(OgState); }

int og_check_action(const OgState* s, int action) {            /* go_env.cc:84-88 (no PASS remap there) */
    if (action == OG_PASS || action == OG_RESIGN) return 1;
    Groups g; analyse(s, &g);
    return point_legal(s, &g, action, s->next_player);
}

static void advance(OgState* s, int c) {                       /* board.cc:536-542 */
    s->next_player = OPP(s->next_player);
    s->last2 = s->last1; s->last1 = c; s->step_count++;
}

int og_step(OgState* s, int action, int* accepted) {           /* go_env.cc:51-80 */
    int n = s->n, P = P_OF(s);
    if (accepted) *accepted = 0;
    if (s->terminated) return 1;                               /* :52-55 */
    if (action == P) action = OG_PASS;                         /* :56-57 */
    uint8_t me = s->next_player;
    int done = 0;
    if (action == OG_PASS || action == OG_RESIGN) {            /* board.cc:554-558, 656-661 */
        advance(s, action);
        done = s->step_count > 1 && ((s->last1 == OG_PASS && s->last2 == OG_PASS) || s->last1 == OG_RESIGN);
    } else {
        Groups g; analyse(s, &g);
        if (!point_legal(s, &g, action, me)) return 0;          /* go_env.cc:75-79: state kept, done=false */
        /* ko bookkeeping BEFORE captures: board.cc:163-194, 561-570 */
        int self_lib = 0, own_nb = 0, n11 = 0, ko_at = OG_INVALID, seen[4], ns = 0;
        for (int d = 0; d < 4; ++d) {
            int x = action % n + DX[d], y = action / n + DY[d];
            if (!on_xy(s, x, y)) continue;
            int q = y * n + x;
            if (s->color[q] == OG_EMPTY) self_lib++;
            else if (s->color[q] == me) own_nb++;
            else if (g.libs[q] == 1 && g.size[q] == 1) {
                int dup = 0;
                for (int i = 0; i < ns; ++i) dup |= (seen[i] == g.gid[q]);
                if (!dup) { seen[ns++] = g.gid[q]; n11++; ko_at = q; }
            }
        }
        if (self_lib == 0 && own_nb == 0 && n11 == 1) { s->ko_loc = ko_at; s->ko_color = OPP(me); s->ko_age = 0; }
        else s->ko_age++;
        /* place, then lift every enemy neighbour group whose only liberty was this point (board.cc:606-630) */
        s->color[action] = me;
        for (int d = 0; d < 4; ++d) {
            int x = action % n + DX[d], y = action / n + DY[d];
            if (!on_xy(s, x, y)) continue;
            int q = y * n + x;
            if (s->color[q] == OPP(me) && g.libs[q] == 1) {
                int gid = g.gid[q];
                for (int i = 0; i < P; ++i) if (g.gid[i] == gid) s->color[i] = OG_EMPTY;
            }
        }
        advance(s, action);
    }
    if (accepted) *accepted = 1;
    if (done || s->step_count > s->max_step) { s->terminated = 1; return 1; }   /* go_env.cc:67-73 */
    return done;
}

int og_legal(const OgState* s, int32_t* out) {                 /* board.cc:467-489 + go_env.cc:154-171 */
    int P = P_OF(s), k = 0;
    Groups g; analyse(s, &g);
    for (int p = 0; p < P; ++p) if (point_legal(s, &g, p, s->next_player)) out[k++] = p;
    out[k++] = P;
    return k;
}

int og_legal_noeye(const OgState* s, int32_t* out) {           /* board.cc:492-517 + go_env.cc:178-188 */
    int P = P_OF(s), k = 0;
    Groups g; analyse(s, &g);
    for (int p = 0; p < P; ++p)
        if (point_legal(s, &g, p, s->next_player) && !true_eye(s, p, s->next_player)) out[k++] = p;
    out[k++] = P;
    return k;
}

/* board.cc:731-817 GivenBlockLives for the group whose id (lowest point) is gid */
static int group_lives(const OgState* s, const Groups* g, int gid) {
    int n = s->n, P = P_OF(s);
    uint8_t col = s->color[gid];
    if (g->libs[gid] <= 1) return 0;
    uint8_t cand[OG_MAX_P];
    int eyes[OG_MAX_P], ne = 0;
    memset(cand, 0, sizeof(cand));
    for (int p = 0; p < P; ++p) {
        if (g->gid[p] != gid) continue;
        for (int d = 0; d < 4; ++d) {
            int x = p % n + DX[d], y = p / n + DY[d];
            if (!on_xy(s, x, y)) continue;
            int q = y * n + x;
            if (!cand[q] && true_eye(s, q, col)) { cand[q] = 1; eyes[ne++] = q; }
        }
    }
    if (ne <= 1) return 0;
    int good = 0;
    for (int i = 0; i < ne; ++i) {
        int nb = 0, nt = 0;
        for (int d = 0; d < 4; ++d) {
            int x = eyes[i] % n + GX[d], y = eyes[i] / n + GY[d];
            if (!on_xy(s, x, y)) { nb++; continue; }
            int q = y * n + x;
            if (s->color[q] == OG_EMPTY) { if (cand[q]) nt++; }
            else if (s->color[q] == col) nt++;
        }
        if ((nb >= 1 && nb + nt == 4) || (nb == 0 && nt >= 3)) good++;
        if (good >= 2) return 1;
    }
    return 0;
}

void og_encode10(const OgState* s, float* f) {                 /* board_feature.cc:17-41,69-100,142-182,213-223 */
    int P = P_OF(s);
    uint8_t me = s->next_player, op = OPP(me);
    Groups g; analyse(s, &g);
    memset(f, 0, sizeof(float) * 10 * P);
    for (int p = 0; p < P; ++p) {
        uint8_t c = s->color[p];
        if (c != OG_EMPTY) {
            int base = (c == me) ? 0 : 3;
            int cls = g.libs[p] == 1 ? 0 : (g.libs[p] == 2 ? 1 : (g.libs[p] >= 3 ? 2 : -1));
            if (cls >= 0) f[(base + cls) * P + p] = 1.0f;
        } else {
            if (is_suicide(s, &g, p, me)) f[7 * P + p] = 1.0f;          /* FindAllSuicideMoves board.cc:520-533 */
            if (true_eye(s, p, me)) f[8 * P + p] = 1.0f;               /* getTrueEyeMap */
        }
    }
    if (s->last1 >= 0 && s->last1 < P) f[6 * P + s->last1] = 1.0f;    /* getHistoryMap1 */
    if (s->ko_age == 0 && s->ko_loc >= 0 && s->ko_loc < P) f[7 * P + s->ko_loc] = 1.0f;  /* board.cc:205-213: colour ignored */
    (void)op;
    for (int p = 0; p < P; ++p) {                                      /* getBlockLives */
        if (s->color[p] != me || g.gid[p] != p) continue;
        if (group_lives(s, &g, p))
            for (int q = 0; q < P; ++q) if (g.gid[q] == p) f[9 * P + q] = 1.0f;
    }
}

/* board.cc:822-958: Tromp-Taylor area.  own[] in {0 dame, 1 black, 2 white}.  Returns black-white, and
 * reproduces the early `return 0.0` on a stone-less board (board.cc:932-935). */
static int area(const OgState* s, uint8_t* own) {
    int n = s->n, P = P_OF(s), nb = 0, nw = 0, any = 0;
    uint8_t seen[OG_MAX_P];
    memset(seen, 0, sizeof(seen));
    for (int p = 0; p < P; ++p) {
        own[p] = 0;
        if (s->color[p] == OG_BLACK) { own[p] = 1; nb++; any = 1; }
        else if (s->color[p] == OG_WHITE) { own[p] = 2; nw++; any = 1; }
    }
    if (!any) return 0;
    for (int p = 0; p < P; ++p) {
        if (s->color[p] != OG_EMPTY || seen[p]) continue;
        int q[OG_MAX_P], qh = 0, qt = 0, touch = 0;
        q[qt++] = p; seen[p] = 1;
        while (qh < qt) {
            int c = q[qh++];
            for (int d = 0; d < 4; ++d) {
                int x = c % n + DX[d], y = c / n + DY[d];
                if (!on_xy(s, x, y)) continue;
                int r = y * n + x;
                if (s->color[r] == OG_EMPTY) { if (!seen[r]) { seen[r] = 1; q[qt++] = r; } }
                else touch |= s->color[r];
            }
        }
        if (touch == 1 || touch == 2) {
            for (int i = 0; i < qt; ++i) own[q[i]] = (uint8_t)touch;
            if (touch == 1) nb += qt; else nw += qt;
        }
    }
    return nb - nw;
}

float og_score(const OgState* s) { uint8_t own[OG_MAX_P]; return (float)area(s, own) - s->komi; }

float og_territory(const OgState* s, float* t) {               /* go_env.cc:136-149: +1 black, 0 dame, -1 white */
    uint8_t own[OG_MAX_P];
    int raw = area(s, own), P = P_OF(s);
    for (int p = 0; p < P; ++p) t[p] = own[p] == 1 ? 1.0f : (own[p] == 2 ? -1.0f : 0.0f);
    return (float)raw - s->komi;
}

void og_sub_encode(const float* enc, float* sub, int n, int sub_n, int channels, int cuts) {   /* board.cc:1166-1271 */
    int iv = n - sub_n, ter = iv / 2;
    for (int i = 0; i < cuts; ++i) {
        int x0 = (i == 1 || i == 3) ? iv : (i == 4 ? ter : 0);
        int y0 = (i == 2 || i == 3) ? iv : (i == 4 ? ter : 0);
        for (int j = 0; j < channels; ++j)
            for (int y = 0; y < sub_n; ++y)
                for (int x = 0; x < sub_n; ++x)
                    sub[(i * channels + j) * sub_n * sub_n + y * sub_n + x] = enc[j * n * n + (y + y0) * n + (x + x0)];
    }
}

/* ---- batch trace replay (checker for tg_env_playout; SURVEY.md 8d C2 move rule) ---------------------------------
 * mix32: the hash the playout kernels draw moves with (tests/helpers.py: mix32, provider.mix32 are its twins). */
static uint32_t og_mix32(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t z = seed + a * 0x9E3779B97F4A7C15ull + b * 0xBF58476D1CE4E5B9ull;
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (uint32_t)(z >> 32);
}

/* Replays trace[plies][G] (device-chosen actions, -3 after the game ended) for games [0, G): every action must be the
 * k-th legal point, k = (mix32(seed, game_id0 + g, step_count) * n_legal) >> 32, PASS iff there is none
 * (environment.py:121-129 rule).  Fills finals[G] (nullable).  Returns the number of plies replayed, or -(1 + index of
 * the first mismatch) with bad[0..3] = game, ply, expected, got. */
long long og_replay_playout(int n, int max_step, float komi, const int16_t* trace, long long plies, long long G,
                            uint64_t seed, long long game_id0, int no_eye, OgState* finals, int32_t* bad) {
    int P = n * n;
    long long total = 0;
    int32_t legal[OG_MAX_P + 1];
    for (long long g = 0; g < G; ++g) {
        OgState s;
        og_reset(&s, n, max_step, komi);
        for (long long t = 0; t < plies; ++t) {
            int got = trace[t * G + g], want;
            if (s.terminated) want = OG_INVALID;
            else {
                int k = (no_eye ? og_legal_noeye(&s, legal) : og_legal(&s, legal)) - 1;   /* drop the appended PASS */
                if (k == 0) want = P;
                else want = legal[((uint64_t)og_mix32(seed, (uint64_t)(game_id0 + g), (uint64_t)s.step_count) * (uint64_t)k) >> 32];
            }
            if (got != want) {
                if (bad) { bad[0] = (int32_t)g; bad[1] = (int32_t)t; bad[2] = want; bad[3] = got; }
                return -(1 + g * plies + t);
            }
            if (!s.terminated) { int acc; og_step(&s, want, &acc); total++; }
        }
        if (finals) finals[g] = s;
    }
    return total;
}

/* observables of many states at once: planes [G][10P], score [G], territory [G][P], legal (python rule) [G][P+1] u8 */
void og_batch_observe(const OgState* st, long long G, float* planes, float* score, float* terr, uint8_t* legal, uint8_t* noeye) {
    int32_t buf[OG_MAX_P + 1];
    for (long long g = 0; g < G; ++g) {
        const OgState* s = &st[g];
        int P = P_OF(s);
        if (planes) og_encode10(s, planes + g * 10 * P);
        if (terr) { float sc = og_territory(s, terr + g * P); if (score) score[g] = sc; }
        else if (score) score[g] = og_score(s);
        if (legal) {
            int k = og_legal(s, buf);
            memset(legal + g * (P + 1), 0, (size_t)P + 1);
            if (k == 1) legal[g * (P + 1) + P] = 1;
            else for (int i = 0; i < k - 1; ++i) legal[g * (P + 1) + buf[i]] = 1;
        }
        if (noeye) {
            int k = og_legal_noeye(s, buf);
            memset(noeye + g * (P + 1), 0, (size_t)P + 1);
            if (k == 1) noeye[g * (P + 1) + P] = 1;
            else for (int i = 0; i < k - 1; ++i) noeye[g * (P + 1) + buf[i]] = 1;
        }
    }
}
