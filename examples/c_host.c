/* A plain-C host on the C-ABI (include/ss_env.h): no Python, no torch.  It does what a ShmemVecEnv user does --
 * reset(), then step() with host action / observation buffers (common/envs_utils.py:542-558) -- and prints the
 * end-to-end rate.  Tables come from `python tools/dump_tables.py DIR walker3d N SEED`.
 *
 *   gcc -O2 -I include -o c_host examples/c_host.c -L steppingstone_b200 -lss_b200 -Wl,-rpath,$PWD/steppingstone_b200 -lm
 *   ./c_host DIR N STEPS
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "ss_env.h"

static void* slurp(const char* dir, const char* name, size_t want) {
    char path[512]; snprintf(path, sizeof path, "%s/%s", dir, name);
    FILE* f = fopen(path, "rb"); if (!f) { perror(path); exit(1); }
    void* p = malloc(want);
    if (fread(p, 1, want, f) != want) { fprintf(stderr, "%s: short read (want %zu bytes)\n", path, want); exit(1); }
    fclose(f); return p;
}
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s DIR N STEPS\n", argv[0]); return 2; }
    const char* dir = argv[1]; int n = atoi(argv[2]), steps = atoi(argv[3]);
    ss_model* M = slurp(dir, "model.bin", sizeof(ss_model));
    ss_params* P = slurp(dir, "params.bin", sizeof(ss_params));
    uint32_t* mt_env = slurp(dir, "mt_env.bin", (size_t)n * SS_MT_WORDS * 4);
    uint32_t* mt_aux = slurp(dir, "mt_aux.bin", (size_t)n * SS_MT_WORDS * 4);
    ss_env* env = NULL;
    if (ss_create(M, P, n, 0, mt_env, mt_aux, &env)) { fprintf(stderr, "ss_create: %s\n", ss_last_error()); return 1; }
    float* act = calloc((size_t)n * SS_ACT, 4); float* obs = malloc((size_t)n * SS_OBS * 4);
    double* rew = malloc((size_t)n * 8); uint8_t* done = malloc(n); uint8_t* to = malloc(n);
    float* ep_ret = malloc((size_t)n * 4); int32_t* ep_len = malloc((size_t)n * 4);
    if (ss_reset_host(env, obs)) { fprintf(stderr, "ss_reset_host: %s\n", ss_last_error()); return 1; }
    unsigned s = 12345; double sum_r = 0; long episodes = 0;
    double t0 = 0;
    for (int t = 0; t < steps + 8; t++) {
        if (t == 8) t0 = now();                                   /* 8 warm-up steps */
        for (size_t i = 0; i < (size_t)n * SS_ACT; i++) {         /* cheap uniform actions in [-0.4, 0.4] */
            s = s * 1664525u + 1013904223u; act[i] = ((s >> 8) * (1.0f / 16777216.0f) - 0.5f) * 0.8f;
        }
        if (ss_step_host(env, act, obs, rew, done, to, ep_ret, ep_len)) { fprintf(stderr, "ss_step_host: %s\n", ss_last_error()); return 1; }
        for (int i = 0; i < n; i++) { sum_r += rew[i]; episodes += done[i]; }
    }
    double dt = now() - t0;
    printf("%d envs x %d steps through ss_step_host: %.2f M env-steps/s end to end (incl. action generation), "
           "mean reward %.3f, %ld episodes finished, obs[0][0..2] = %.4f %.4f %.4f\n",
           n, steps, (double)n * steps / dt / 1e6, sum_r / ((double)n * (steps + 8)), episodes, obs[0], obs[1], obs[2]);
    ss_destroy(env);
    return isfinite(sum_r) ? 0 : 1;
}
