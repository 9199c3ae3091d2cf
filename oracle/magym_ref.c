/*
 * CPU restatement in plain C of ma-gym 0.0.14 Checkers-v0 and Switch{2,4}-v0 -- TEST INFRASTRUCTURE.
 *
 * Oracle for SURVEY.md section 8 rows a1/a2; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may link or call it.  PARITY UNPINNED: the algorithm lives in
 * the un-vendored third-party package ma-gym==0.0.14 (see oracle/ma_gym_ref.py for the evidence this
 * restatement is anchored on); tests/test_oracle_env.py checks this file cell-for-cell against that
 * Python restatement.  Deliberately written the way the original is -- a grid of cell codes mutated
 * agent by agent -- and NOT the way the CUDA kernel is (bitmaps + shuffles), so the two are
 * independent statements of the same rules.
 *
 * Reference call sites: vdn/main.py:83,93; mappo/runner/shared/magym_runner.py:52-57;
 * offpolicy/runner/shared/magym_runner.py:94-97; qmix/main.py:103,115.
 */
#include <stdint.h>
#include <string.h>

/* ---------------------------------------------------------------- Checkers ---- */
enum { C_EMPTY = 0, C_LEMON = 1, C_APPLE = 2, C_WALL = 3, C_A1 = 4, C_A2 = 5 };
#define CK_ROWS 3
#define CK_COLS 8

typedef struct ck_env {
    int8_t grid[CK_ROWS][CK_COLS];
    int32_t pos[2][2];
    int32_t step_count;
    int32_t lemons, apples;
    int32_t dones;          /* all agents share one flag in Checkers */
    int32_t beyond_done;    /* steps_beyond_done is not None */
} ck_env;

/* Python: round(r / 3, 2), round(c / 7, 2) -- the doubles nearest to these decimals */
static const double CK_POS_ROW[3] = {0.0, 0.33, 0.67};
static const double CK_POS_COL[8] = {0.0, 0.14, 0.29, 0.43, 0.57, 0.71, 0.86, 1.0};

int ck_env_size(void) { return (int)sizeof(ck_env); }

void ck_reset(ck_env* e) {
    for (int r = 0; r < CK_ROWS; ++r)
        for (int c = 0; c < CK_COLS; ++c) {
            if (c >= CK_COLS - 2) e->grid[r][c] = C_EMPTY;
            else if (r % 2 == 0) e->grid[r][c] = (c % 2 == 0) ? C_APPLE : C_LEMON;
            else e->grid[r][c] = (c % 2 != 0) ? C_APPLE : C_LEMON;
        }
    e->pos[0][0] = 0; e->pos[0][1] = CK_COLS - 2;
    e->pos[1][0] = 2; e->pos[1][1] = CK_COLS - 2;
    e->grid[0][CK_COLS - 2] = C_A1;
    e->grid[2][CK_COLS - 2] = C_A2;
    e->step_count = 0;
    e->lemons = e->apples = ((CK_COLS - 2) / 2) * CK_ROWS;
    e->dones = 0;
    e->beyond_done = 0;
}

/* obs: [2][47] (full_observable == 0) or [2][94] */
void ck_obs(const ck_env* e, int full_observable, float* obs) {
    float one[2][47];
    for (int i = 0; i < 2; ++i) {
        const int r0 = e->pos[i][0], c0 = e->pos[i][1];
        float* o = one[i];
        memset(o, 0, sizeof(float) * 47);
        o[0] = (float)CK_POS_ROW[r0];
        o[1] = (float)CK_POS_COL[c0];
        int k = 2;
        for (int r = r0 - 1; r <= r0 + 1; ++r)
            for (int c = c0 - 1; c <= c0 + 1; ++c, k += 5) {
                if (r < 0 || r >= CK_ROWS || c < 0 || c >= CK_COLS) continue;
                switch (e->grid[r][c]) {
                    case C_LEMON: o[k + 0] = 1.f; break;
                    case C_APPLE: o[k + 1] = 1.f; break;
                    case C_A1:    o[k + 2] = 1.f; break;
                    case C_A2:    o[k + 3] = 1.f; break;
                    case C_WALL:  o[k + 4] = 1.f; break;
                    default: break;
                }
            }
    }
    if (!full_observable) {
        memcpy(obs, one, sizeof(one));
    } else {
        for (int i = 0; i < 2; ++i) memcpy(obs + i * 94, one, sizeof(one));
    }
}

/* returns 0, or -1 on an action outside [0,5) (the original raises) */
int ck_step(ck_env* e, const int32_t* actions, double step_cost, int max_steps, float* rewards,
            uint8_t* dones) {
    static const double food_reward[2][2] = {{-10, 10}, {-1, 1}};   /* [agent][lemon, apple] */
    double rew[2] = {step_cost, step_cost};
    e->step_count += 1;
    for (int i = 0; i < 2; ++i) {
        const int r = e->pos[i][0], c = e->pos[i][1];
        int nr = r, nc = c, wants = 1;
        switch (actions[i]) {
            case 0: nr = r + 1; break;
            case 1: nc = c - 1; break;
            case 2: nr = r - 1; break;
            case 3: nc = c + 1; break;
            case 4: wants = 0; break;
            default: return -1;
        }
        if (wants && nr >= 0 && nr < CK_ROWS && nc >= 0 && nc < CK_COLS && e->grid[nr][nc] != C_A1 &&
            e->grid[nr][nc] != C_A2) {
            const int8_t cell = e->grid[nr][nc];
            if (cell == C_LEMON) { rew[i] += food_reward[i][0]; e->lemons -= 1; }
            else if (cell == C_APPLE) { rew[i] += food_reward[i][1]; e->apples -= 1; }
            e->grid[r][c] = C_EMPTY;
            e->grid[nr][nc] = (int8_t)(C_A1 + i);
            e->pos[i][0] = nr; e->pos[i][1] = nc;
        }
    }
    if (e->step_count >= max_steps || e->apples == 0) e->dones = 1;
    if (!e->beyond_done && e->dones) {
        e->beyond_done = 1;
    } else if (e->beyond_done) {
        rew[0] = rew[1] = 0;
    }
    rewards[0] = (float)rew[0]; rewards[1] = (float)rew[1];
    dones[0] = dones[1] = (uint8_t)e->dones;
    return 0;
}

/* lockstep batch helpers (what the CUDA kernel is compared with at scale) */
void ck_batch_reset(ck_env* envs, int n, int full_observable, float* obs) {
    const int d = 2 * 47 * (full_observable ? 2 : 1);
    for (int i = 0; i < n; ++i) {
        ck_reset(&envs[i]);
        if (obs) ck_obs(&envs[i], full_observable, obs + (size_t)i * d);
    }
}

int ck_batch_step(ck_env* envs, int n, const int32_t* actions, double step_cost, int max_steps,
                  int full_observable, int auto_reset, float* obs, float* rewards, uint8_t* dones) {
    const int d = 2 * 47 * (full_observable ? 2 : 1);
    int rc = 0;
    for (int i = 0; i < n; ++i) {
        rc |= ck_step(&envs[i], actions + 2 * i, step_cost, max_steps, rewards + 2 * i, dones + 2 * i);
        if (auto_reset && dones[2 * i]) ck_reset(&envs[i]);
        if (obs) ck_obs(&envs[i], full_observable, obs + (size_t)i * d);
    }
    return rc;
}

/* positions [n][2][2], step counts [n], food left [n][2] (lemons, apples) for state parity */
void ck_batch_export(const ck_env* envs, int n, int32_t* pos, int32_t* steps, int32_t* food) {
    for (int i = 0; i < n; ++i) {
        memcpy(pos + 4 * i, envs[i].pos, sizeof(int32_t) * 4);
        steps[i] = envs[i].step_count;
        food[2 * i] = envs[i].lemons;
        food[2 * i + 1] = envs[i].apples;
    }
}

/* ------------------------------------------------------------------ Switch ---- */
#define SW_ROWS 3
#define SW_COLS 7

typedef struct sw_env {
    int8_t grid[SW_ROWS][SW_COLS];   /* -1 wall, 0 free, i+1 agent i */
    int32_t pos[4][2];
    int32_t done[4];
    int32_t step_count;
    int32_t n_agents;
} sw_env;

static const double SW_POS_ROW[3] = {0.0, 0.5, 1.0};                       /* round(r / 2, 2) */
static const double SW_POS_COL[7] = {0.0, 0.17, 0.33, 0.5, 0.67, 0.83, 1.0}; /* round(c / 6, 2) */
static const int SW_INIT[4][2] = {{0, 1}, {0, SW_COLS - 2}, {2, 1}, {2, SW_COLS - 2}};
static const int SW_GOAL[4][2] = {{0, SW_COLS - 1}, {0, 0}, {2, SW_COLS - 1}, {2, 0}};

int sw_env_size(void) { return (int)sizeof(sw_env); }

void sw_reset(sw_env* e, int n_agents) {
    e->n_agents = n_agents;
    for (int r = 0; r < SW_ROWS; ++r)
        for (int c = 0; c < SW_COLS; ++c)
            e->grid[r][c] = (r == SW_ROWS / 2 || c <= 1 || c >= SW_COLS - 2) ? 0 : -1;
    for (int i = 0; i < n_agents; ++i) {
        e->pos[i][0] = SW_INIT[i][0]; e->pos[i][1] = SW_INIT[i][1];
        e->grid[e->pos[i][0]][e->pos[i][1]] = (int8_t)(i + 1);
        e->done[i] = 0;
    }
    e->step_count = 0;
}

void sw_obs(const sw_env* e, int full_observable, int clock, int max_steps, float* obs) {
    const int d = 2 + (clock ? 1 : 0), A = e->n_agents;
    float one[4][3];
    for (int i = 0; i < A; ++i) {
        one[i][0] = (float)SW_POS_ROW[e->pos[i][0]];
        one[i][1] = (float)SW_POS_COL[e->pos[i][1]];
        one[i][2] = (float)((double)e->step_count / (double)max_steps);
    }
    if (!full_observable) {
        for (int i = 0; i < A; ++i) memcpy(obs + i * d, one[i], sizeof(float) * d);
    } else {
        for (int i = 0; i < A; ++i)
            for (int j = 0; j < A; ++j) memcpy(obs + (i * A + j) * d, one[j], sizeof(float) * d);
    }
}

int sw_step(sw_env* e, const int32_t* actions, double step_cost, int max_steps, float* rewards,
            uint8_t* dones) {
    const int A = e->n_agents;
    e->step_count += 1;
    for (int i = 0; i < A; ++i) {
        double rew = step_cost;
        if (!e->done[i]) {
            const int r = e->pos[i][0], c = e->pos[i][1];
            int nr = r, nc = c, wants = 1;
            switch (actions[i]) {
                case 0: nr = r + 1; break;
                case 1: nc = c - 1; break;
                case 2: nr = r - 1; break;
                case 3: nc = c + 1; break;
                case 4: wants = 0; break;
                default: return -1;
            }
            if (wants && nr >= 0 && nr < SW_ROWS && nc >= 0 && nc < SW_COLS && e->grid[nr][nc] == 0) {
                e->grid[r][c] = 0;
                e->grid[nr][nc] = (int8_t)(i + 1);
                e->pos[i][0] = nr; e->pos[i][1] = nc;
            }
            e->done[i] = (e->pos[i][0] == SW_GOAL[i][0] && e->pos[i][1] == SW_GOAL[i][1]);
            if (e->done[i]) rew = 5;
        }
        rewards[i] = (float)rew;
    }
    if (e->step_count >= max_steps)
        for (int i = 0; i < A; ++i) e->done[i] = 1;
    for (int i = 0; i < A; ++i) dones[i] = (uint8_t)e->done[i];
    return 0;
}

void sw_batch_reset(sw_env* envs, int n, int n_agents, int full_observable, int clock, int max_steps,
                    float* obs) {
    const int d = (2 + (clock ? 1 : 0)) * n_agents * (full_observable ? n_agents : 1);
    for (int i = 0; i < n; ++i) {
        sw_reset(&envs[i], n_agents);
        if (obs) sw_obs(&envs[i], full_observable, clock, max_steps, obs + (size_t)i * d);
    }
}

int sw_batch_step(sw_env* envs, int n, const int32_t* actions, double step_cost, int max_steps,
                  int full_observable, int clock, float* obs, float* rewards, uint8_t* dones) {
    const int A = envs[0].n_agents;
    const int d = (2 + (clock ? 1 : 0)) * A * (full_observable ? A : 1);
    int rc = 0;
    for (int i = 0; i < n; ++i) {
        rc |= sw_step(&envs[i], actions + A * i, step_cost, max_steps, rewards + A * i, dones + A * i);
        if (obs) sw_obs(&envs[i], full_observable, clock, max_steps, obs + (size_t)i * d);
    }
    return rc;
}

void sw_batch_export(const sw_env* envs, int n, int32_t* pos /* [n][A][2] */) {
    for (int i = 0; i < n; ++i)
        memcpy(pos + (size_t)i * envs[i].n_agents * 2, envs[i].pos, sizeof(int32_t) * 2 * envs[i].n_agents);
}
