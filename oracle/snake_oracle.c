/* snake_oracle.c -- plain-C CPU oracle for the batched multi-snake env step.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load the library built from this file (oracle/Makefile ->
 * oracle/libsnake_oracle.so).  The product (libsnk.so) never links or calls it.
 *
 * It restates, for N envs held in the canonical state layout of include/snk.h
 * (snk_state_layout), the algorithm of the reference's per-instance Python env.  Paths below are
 * relative to /root/reference/src/gym-snake/gym_snake/:
 *   so_reset_env      envs/snake_multiple_test.py:219-232 (reset), :199-200 (choose_cell)
 *   so_spawn_cell     envs/snake_multiple_test.py:202-217 (safe_choose_cell, incl. the
 *                     un-bounds-checked y*dim+x aliasing of out-of-board heads)
 *   so_advance        envs/snake_multiple_test.py:97-145 (update_snake);
 *                     envs/snake_adversarial_env.py:137-141 (spare_fruits on eat)
 *   so_step_env       envs/snake_multiple_test.py:166-197 (step), :147-164 (is_snake_alive);
 *                     envs/snake_adversarial_env.py:180-186 (dead body -> fruit, spare += len^2)
 *   so_encode_obs     envs/snake_multiple_test.py:24-58 (draw_snake / get_ob_for_snake), :93-95;
 *                     core/new_world.py:206-214 (K = S views)
 *   auto-reset        ../../baselines/common/vec_env/subproc_vec_env.py:13-16
 *   episode stats     ../../baselines/bench/monitor.py:57-78
 * The `cut` rule-set has no reference code (README.md:11 only); it follows DESIGN.md and its
 * parity is UNPINNED.
 *
 * Pinning: validated against recordings of the reference itself (tests/golden/, produced by
 * oracle/make_golden.py from the unmodified reference run under oracle/gym_shim) by
 * tests/test_oracle_golden.py, and step-for-step against oracle/snake_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/snk.h"

typedef struct so_vec {
  snk_config cfg;
  snk_state_layout lay;
  int D, V, S, F, K, cap, n_actions;
  int64_t N;
  /* canonical state, exactly the dump layout */
  uint8_t* blob;
  int32_t* t;
  uint32_t* spare;
  uint32_t* draw_ctr;
  float* ep_ret;
  int32_t* ep_len;
  uint16_t* len;
  uint16_t* grow;
  uint8_t* vel;
  uint16_t* body;
  uint16_t* fruit;  /* classic */
  uint8_t* grid;    /* adversarial / cut */
  /* draws */
  const uint32_t* tape_vals;
  const uint32_t* tape_bounds;
  const uint64_t* tape_off;
  uint32_t* tape_store;
  uint64_t* tape_off_store;
  int rng_mode;
  uint32_t errors;
  double stats[SNK_NSTATS];
} so_vec;

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void so_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }

static uint32_t philox_bounded(uint64_t seed, uint64_t env_id, uint32_t stream, uint64_t index, uint32_t n) {
  /* one Philox block (counter = index >> 2) serves four consecutive draws (word = index & 3) */
  uint64_t blk = index >> 2;
  uint32_t ctr[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), stream, (uint32_t)(seed >> 32)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)env_id};
  uint32_t out[4];
  philox4x32_10(ctr, key, out);
  return (uint32_t)(((uint64_t)out[index & 3] * n) >> 32);
}

static uint32_t so_draw(so_vec* v, int64_t e, uint32_t n) {
  uint32_t i = v->draw_ctr[e]++;
  v->stats[SNK_STAT_DRAWS] += 1;
  if (v->rng_mode == SNK_RNG_TAPE) {
    uint64_t pos = v->tape_off[e] + i;
    if (pos >= v->tape_off[e + 1]) { v->errors |= SNK_DEVERR_TAPE_UNDERRUN; return 0; }
    if (v->tape_bounds && v->tape_bounds[pos] != n) v->errors |= SNK_DEVERR_TAPE_BOUND;
    return v->tape_vals[pos];
  }
  return philox_bounded(v->cfg.seed, (uint64_t)(v->cfg.env_id_base + e), 0, i, n);
}

/* ------------------------------------------------------------------ layout / lifetime */
static size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

int so_state_layout_of(const snk_config* c, snk_state_layout* o) {
  if (!c || !o) return SNK_EINVAL;
  size_t N = (size_t)c->num_envs, S = (size_t)c->n_snakes, F = (size_t)c->n_fruits;
  size_t D = (size_t)c->size, V = D + 2;
  int cap = (int)((D * D + 1 + 7) & ~(size_t)7);
  size_t off = 0;
  o->off_t = off;        off = align16(off + 4 * N);
  o->off_spare = off;    off = align16(off + 4 * N);
  o->off_draw_ctr = off; off = align16(off + 4 * N);
  o->off_ep_ret = off;   off = align16(off + 4 * N);
  o->off_ep_len = off;   off = align16(off + 4 * N);
  o->off_len = off;      off = align16(off + 2 * N * S);
  o->off_grow_to = off;  off = align16(off + 2 * N * S);
  o->off_vel = off;      off = align16(off + N * S);
  o->off_body = off;     off = align16(off + 2 * N * S * (size_t)cap);
  o->off_fruit = off;
  o->fruit_is_grid = c->rules != SNK_RULES_CLASSIC;
  off = align16(off + (o->fruit_is_grid ? N * V * V : 2 * N * F));
  o->total_bytes = off;
  o->cap = cap;
  return SNK_OK;
}

static int cfg_ok(const snk_config* c) {
  return c && c->size >= 2 && c->size <= 254 && c->n_snakes >= 1 && c->n_snakes <= 32 && c->n_fruits >= 0 &&
         c->n_fruits <= 64 && c->n_views >= 0 && c->n_views <= 32 && c->rules >= 0 && c->rules <= 2 && c->num_envs >= 1;
}

so_vec* so_create(const snk_config* cfg) {
  if (!cfg_ok(cfg)) return NULL;
  so_vec* v = (so_vec*)calloc(1, sizeof(so_vec));
  if (!v) return NULL;
  v->cfg = *cfg;
  if (v->cfg.n_views == 0) v->cfg.n_views = cfg->n_snakes;
  if (v->cfg.max_steps == 0) v->cfg.max_steps = 2000;
  so_state_layout_of(&v->cfg, &v->lay);
  v->D = cfg->size; v->V = v->D + 2; v->S = cfg->n_snakes; v->F = cfg->n_fruits; v->K = v->cfg.n_views;
  v->cap = v->lay.cap; v->N = cfg->num_envs; v->rng_mode = cfg->rng_mode;
  v->n_actions = cfg->rules == SNK_RULES_CUT ? 6 : 5;
  v->blob = (uint8_t*)calloc(1, v->lay.total_bytes);
  if (!v->blob) { free(v); return NULL; }
  v->t = (int32_t*)(v->blob + v->lay.off_t);
  v->spare = (uint32_t*)(v->blob + v->lay.off_spare);
  v->draw_ctr = (uint32_t*)(v->blob + v->lay.off_draw_ctr);
  v->ep_ret = (float*)(v->blob + v->lay.off_ep_ret);
  v->ep_len = (int32_t*)(v->blob + v->lay.off_ep_len);
  v->len = (uint16_t*)(v->blob + v->lay.off_len);
  v->grow = (uint16_t*)(v->blob + v->lay.off_grow_to);
  v->vel = v->blob + v->lay.off_vel;
  v->body = (uint16_t*)(v->blob + v->lay.off_body);
  if (v->lay.fruit_is_grid) v->grid = v->blob + v->lay.off_fruit; else v->fruit = (uint16_t*)(v->blob + v->lay.off_fruit);
  return v;
}

void so_destroy(so_vec* v) {
  if (!v) return;
  free(v->tape_store); free(v->tape_off_store); free(v->blob); free(v);
}

int so_set_draw_tape(so_vec* v, const uint32_t* vals, const uint32_t* bounds, const uint64_t* off) {
  if (!v || !vals || !off) return SNK_EINVAL;
  uint64_t n = off[v->N];
  free(v->tape_store); free(v->tape_off_store);
  v->tape_store = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(2 * n + 2));
  v->tape_off_store = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(v->N + 1));
  if (!v->tape_store || !v->tape_off_store) return SNK_ENOMEM;
  memcpy(v->tape_store, vals, sizeof(uint32_t) * (size_t)n);
  if (bounds) memcpy(v->tape_store + n, bounds, sizeof(uint32_t) * (size_t)n);
  memcpy(v->tape_off_store, off, sizeof(uint64_t) * (size_t)(v->N + 1));
  v->tape_vals = v->tape_store;
  v->tape_bounds = bounds ? v->tape_store + n : NULL;
  v->tape_off = v->tape_off_store;
  v->rng_mode = SNK_RNG_TAPE;
  return SNK_OK;
}

int so_dump_state(so_vec* v, void* dst, size_t bytes) {
  if (!v || !dst || bytes < v->lay.total_bytes) return SNK_EINVAL;
  memcpy(dst, v->blob, v->lay.total_bytes);
  return SNK_OK;
}

int so_load_state(so_vec* v, const void* src, size_t bytes) {
  if (!v || !src || bytes < v->lay.total_bytes) return SNK_EINVAL;
  memcpy(v->blob, src, v->lay.total_bytes);
  return SNK_OK;
}

int so_get_stats(so_vec* v, double* out) { memcpy(out, v->stats, sizeof(v->stats)); return SNK_OK; }
int so_reset_stats(so_vec* v) { memset(v->stats, 0, sizeof(v->stats)); return SNK_OK; }
uint32_t so_check_errors(so_vec* v) { uint32_t e = v->errors; v->errors = 0; return e; }

/* ------------------------------------------------------------------ game logic */
static inline uint16_t* body_of(so_vec* v, int64_t e, int s) { return v->body + ((size_t)e * v->S + s) * v->cap; }

static void add_fruit_grid(so_vec* v, int64_t e, int pid) {
  uint8_t* g = v->grid + (size_t)e * v->V * v->V;
  if (g[pid] == 255) v->errors |= SNK_DEVERR_FRUIT_OVERFLOW; else g[pid]++;
}

static void so_reset_env(so_vec* v, int64_t e) {
  const int D = v->D, V = v->V, S = v->S, F = v->F;
  memset(body_of(v, e, 0), 0, sizeof(uint16_t) * (size_t)S * v->cap);
  if (v->grid) memset(v->grid + (size_t)e * V * V, 0, (size_t)V * V);
  int n = S > F ? S : F;
  for (int i = 0; i < n; ++i) {
    if (i < S) {
      int x = (int)so_draw(v, e, (uint32_t)D);
      int y = (int)so_draw(v, e, (uint32_t)D);
      body_of(v, e, i)[0] = (uint16_t)((x + 1) * V + (y + 1));
      v->len[e * S + i] = 1; v->grow[e * S + i] = 3; v->vel[e * S + i] = 0;
    }
    if (i < F) {
      int x = (int)so_draw(v, e, (uint32_t)D);
      int y = (int)so_draw(v, e, (uint32_t)D);
      int pid = (x + 1) * V + (y + 1);
      if (v->grid) add_fruit_grid(v, e, pid); else v->fruit[e * F + i] = (uint16_t)pid;
    }
  }
  v->t[e] = 0; v->ep_ret[e] = 0.f; v->ep_len[e] = 0;
  /* `spare` deliberately survives: snake_adversarial_env.py:14 sets it in __init__ only */
}

/* k-th free board index in y-major order (idx = y*D + x), bodies not bounds-checked */
static int so_spawn_cell(so_vec* v, int64_t e) {
  const int D = v->D, V = v->V, S = v->S, DD = D * D;
  uint8_t used[254 * 254];
  memset(used, 0, (size_t)DD);
  for (int s = 0; s < S; ++s) {
    const uint16_t* b = body_of(v, e, s);
    for (int i = 0; i < v->len[e * S + s]; ++i) {
      int x = b[i] / V - 1, y = b[i] % V - 1;
      int idx = y * D + x;
      if (idx >= 0 && idx < DD) used[idx] = 1;
    }
  }
  int n_free = 0;
  for (int i = 0; i < DD; ++i) n_free += !used[i];
  if (n_free == 0) return 1 * V + 1; /* (0,0), no draw */
  int k = (int)so_draw(v, e, (uint32_t)n_free);
  for (int i = 0; i < DD; ++i)
    if (!used[i] && k-- == 0) return (i % D + 1) * V + (i / D + 1);
  return 1 * V + 1;
}

static int so_advance(so_vec* v, int64_t e, int s, int action, uint8_t* strike, uint8_t* moved) {
  const int V = v->V, S = v->S, F = v->F, rules = v->cfg.rules;
  static const int opposite[5] = {0, 3, 4, 1, 2};
  const int delta[5] = {0, V, 1, -V, -1};
  int L = v->len[e * S + s];
  if (L == 0) return 0;
  int vel = v->vel[e * S + s];
  if (action >= 1 && action <= 4 && vel != opposite[action]) vel = action;
  if (rules == SNK_RULES_CUT && action == 5) strike[s] = 1;
  if (vel == 0) return 0;
  uint16_t* b = body_of(v, e, s);
  int head = b[0] + delta[vel];
  int n_eat = 0, hit[64];
  if (v->grid) {
    n_eat = v->grid[(size_t)e * V * V + head];
  } else {
    for (int i = 0; i < F; ++i) if (v->fruit[e * F + i] == head) hit[n_eat++] = i;
  }
  int grow = v->grow[e * S + s] + 2 * n_eat;
  if (L >= grow) { L--; b[L] = 0; }
  if (L + 1 > v->cap) { v->errors |= SNK_DEVERR_BODY_OVERFLOW; L = v->cap - 1; }
  memmove(b + 1, b, sizeof(uint16_t) * (size_t)L);
  b[0] = (uint16_t)head; L++;
  v->len[e * S + s] = (uint16_t)L;
  if (rules == SNK_RULES_CLASSIC) {
    for (int i = 0; i < n_eat; ++i) v->fruit[e * F + hit[i]] = (uint16_t)so_spawn_cell(v, e);
  } else {
    for (int i = 0; i < n_eat; ++i) {
      if (rules == SNK_RULES_ADVERSARIAL && v->spare[e] > 0) {
        v->spare[e]--;
      } else {
        v->grid[(size_t)e * V * V + head]--;
        add_fruit_grid(v, e, so_spawn_cell(v, e));
      }
    }
  }
  v->vel[e * S + s] = (uint8_t)vel;
  v->grow[e * S + s] = (uint16_t)grow;
  moved[s] = 1;
  return n_eat;
}

static void so_encode_obs(so_vec* v, int64_t e, uint8_t* ob) {
  const int V = v->V, S = v->S, F = v->F, K = v->K, C = 3 * K;
  memset(ob, 0, (size_t)V * V * C);
  if (v->grid) {
    const uint8_t* g = v->grid + (size_t)e * V * V;
    for (int p = 0; p < V * V; ++p) if (g[p]) for (int k = 0; k < K; ++k) ob[p * C + 3 * k] = 255;
  } else {
    for (int i = 0; i < F; ++i) for (int k = 0; k < K; ++k) {
      uint8_t* px = ob + v->fruit[e * F + i] * C + 3 * k; px[0] = 255; px[1] = 0; px[2] = 0;
    }
  }
  static const uint8_t col[4][3] = {{0, 204, 0}, {191, 242, 191}, {0, 51, 204}, {128, 154, 230}};
  for (int s = 0; s < S; ++s) {
    const uint16_t* b = body_of(v, e, s);
    int L = v->len[e * S + s];
    for (int k = 0; k < K; ++k) {
      const uint8_t* cb = col[s == k ? 0 : 2];
      const uint8_t* ch = col[s == k ? 1 : 3];
      for (int i = L - 1; i >= 0; --i) memcpy(ob + b[i] * C + 3 * k, i ? cb : ch, 3);
    }
  }
  for (int i = 0; i < V; ++i) {
    memset(ob + (0 * V + i) * C, 255, (size_t)C);
    memset(ob + ((V - 1) * V + i) * C, 255, (size_t)C);
    memset(ob + (i * V + 0) * C, 255, (size_t)C);
    memset(ob + (i * V + V - 1) * C, 255, (size_t)C);
  }
}

static void so_step_env(so_vec* v, int64_t e, const int8_t* act, uint8_t* ob, float* reward, float* reward_all,
                        uint8_t* done, uint8_t* num_alive, float* fin_ret, int32_t* fin_len) {
  const int V = v->V, D = v->D, S = v->S, rules = v->cfg.rules;
  uint8_t strike[32] = {0}, moved[32] = {0}, was_alive[32], dead[32] = {0}, saved[32] = {0};
  int eaten[32], cut_at[32];
  for (int s = 0; s < S; ++s) was_alive[s] = v->len[e * S + s] > 0;
  for (int s = 0; s < S; ++s) eaten[s] = so_advance(v, e, s, act[s], strike, moved);

  /* death test on the post-move bodies of all snakes, before anything is cleared */
  for (int i = 0; i < S; ++i) cut_at[i] = -1;
  for (int i = 0; i < S; ++i) {
    int Li = v->len[e * S + i];
    if (Li == 0) { dead[i] = 1; continue; }
    int head = body_of(v, e, i)[0];
    int x = head / V - 1, y = head % V - 1;
    int oob = x < 0 || x >= D || y < 0 || y >= D;
    int hit_own = 0, hit_head = 0, hit_body = 0;
    for (int j = 0; j < S; ++j) {
      const uint16_t* b = body_of(v, e, j);
      for (int k = 0; k < v->len[e * S + j]; ++k) {
        if (b[k] != head || (j == i && k == 0)) continue;
        if (j == i) hit_own = 1; else if (k == 0) hit_head = 1; else hit_body = 1;
      }
    }
    if (oob) dead[i] = 1;
    else if (hit_own || hit_head || hit_body) {
      if (rules == SNK_RULES_CUT && strike[i] && moved[i] && !hit_own && !hit_head) saved[i] = 1;
      else dead[i] = 1;
    }
  }
  if (rules == SNK_RULES_CUT) {
    for (int i = 0; i < S; ++i) {
      if (!saved[i]) continue;
      int head = body_of(v, e, i)[0];
      for (int j = 0; j < S; ++j) {
        if (j == i) continue;
        const uint16_t* b = body_of(v, e, j);
        for (int k = 1; k < v->len[e * S + j]; ++k)
          if (b[k] == head) { if (cut_at[j] < 0 || k < cut_at[j]) cut_at[j] = k; break; }
      }
    }
    for (int j = 0; j < S; ++j) {
      if (cut_at[j] < 0) continue;
      uint16_t* b = body_of(v, e, j);
      for (int k = cut_at[j]; k < v->len[e * S + j]; ++k) {
        int on_saved_head = 0;
        for (int i = 0; i < S; ++i) if (saved[i] && body_of(v, e, i)[0] == b[k]) on_saved_head = 1;
        if (!on_saved_head) add_fruit_grid(v, e, b[k]);
        b[k] = 0;
      }
      v->len[e * S + j] = (uint16_t)cut_at[j];
      v->grow[e * S + j] = (uint16_t)cut_at[j];
    }
  }
  if (rules == SNK_RULES_ADVERSARIAL) {
    for (int i = 0; i < S; ++i) {
      if (!dead[i]) continue;
      int L = v->len[e * S + i];
      for (int k = 0; k < L; ++k) { add_fruit_grid(v, e, body_of(v, e, i)[k]); v->spare[e] += (uint32_t)L; }
    }
  }
  int alive = 0, cells = 0, deaths = 0, fruits = 0;
  for (int i = 0; i < S; ++i) {
    if (dead[i]) {
      memset(body_of(v, e, i), 0, sizeof(uint16_t) * (size_t)v->len[e * S + i]);
      v->len[e * S + i] = 0;
      deaths += was_alive[i];
    } else {
      alive++;
    }
    cells += v->len[e * S + i];
    fruits += eaten[i];
  }
  int main_dead = v->len[e * S] == 0;
  float r = main_dead ? -1.f : (float)eaten[0];
  reward[e] = r;
  if (reward_all) {
    reward_all[e * S] = r;
    for (int s = 1; s < S; ++s) reward_all[e * S + s] = dead[s] ? (was_alive[s] ? -1.f : 0.f) : (float)eaten[s];
  }
  v->t[e] += 1;
  int d = v->t[e] >= v->cfg.max_steps || main_dead;
  done[e] = (uint8_t)d;
  num_alive[e] = (uint8_t)alive;
  v->ep_ret[e] += r;
  v->ep_len[e] += 1;
  v->stats[SNK_STAT_ENV_STEPS] += 1;
  v->stats[SNK_STAT_FRUITS] += fruits;
  v->stats[SNK_STAT_DEATHS] += deaths;
  v->stats[SNK_STAT_BODY_CELLS] += cells;
  if (fin_ret) fin_ret[e] = d ? v->ep_ret[e] : 0.f;
  if (fin_len) fin_len[e] = d ? v->ep_len[e] : 0;
  if (d) {
    v->stats[SNK_STAT_EPISODES] += 1;
    v->stats[SNK_STAT_RETURN_SUM] += v->ep_ret[e];
    v->stats[SNK_STAT_LENGTH_SUM] += v->ep_len[e];
    if (v->cfg.auto_reset) so_reset_env(v, e);
  }
  if (ob) so_encode_obs(v, e, ob + (size_t)e * V * V * 3 * v->K);
}

int so_reset(so_vec* v, const uint8_t* mask, uint8_t* ob) {
  if (!v) return SNK_EINVAL;
  for (int64_t e = 0; e < v->N; ++e) {
    if (mask && !mask[e]) continue;
    so_reset_env(v, e);
    if (ob) so_encode_obs(v, e, ob + (size_t)e * v->V * v->V * 3 * v->K);
  }
  return SNK_OK;
}

/* steps envs [begin, end): lets the harness fan a batch out over host threads */
int so_step_range(so_vec* v, int64_t begin, int64_t end, const int8_t* actions, uint8_t* ob, float* reward,
                  float* reward_all, uint8_t* done, uint8_t* num_alive, float* fin_ret, int32_t* fin_len) {
  if (!v || !actions || !reward || !done || !num_alive || begin < 0 || end > v->N) return SNK_EINVAL;
  for (int64_t e = begin; e < end; ++e)
    so_step_env(v, e, actions + e * v->S, ob, reward, reward_all, done, num_alive, fin_ret, fin_len);
  return SNK_OK;
}

int so_step(so_vec* v, const int8_t* actions, uint8_t* ob, float* reward, float* reward_all, uint8_t* done,
            uint8_t* num_alive, float* fin_ret, int32_t* fin_len) {
  return so_step_range(v, 0, v ? v->N : 0, actions, ob, reward, reward_all, done, num_alive, fin_ret, fin_len);
}

int so_observe(so_vec* v, uint8_t* ob) {
  for (int64_t e = 0; e < v->N; ++e) so_encode_obs(v, e, ob + (size_t)e * v->V * v->V * 3 * v->K);
  return SNK_OK;
}

void so_gen_actions(const snk_config* c, int8_t* actions, uint64_t step, uint64_t seed, int32_t n_actions) {
  for (int64_t e = 0; e < c->num_envs; ++e)
    for (int s = 0; s < c->n_snakes; ++s)
      actions[e * c->n_snakes + s] = (int8_t)philox_bounded(seed, (uint64_t)(c->env_id_base + e), 1,
                                                            step * (uint64_t)c->n_snakes + (uint64_t)s, (uint32_t)n_actions);
}
