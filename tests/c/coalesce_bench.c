/* 100 concurrent per-request callers against libzkgpu's C ABI — the call shape of the reference's prover server
 * (/root/reference/tee/crates/shielder-prover-tee/src/server.rs:157-195: one task per client, each calling generate_proof;
 * /root/reference/tee/crates/shielder-prover-server/src/command_line_args.rs:24-27: at most 100 in flight).
 *
 * Plain C + pthreads, no Python in the measured process.  Reads one input file written by tests/test_gpu_service.py:
 *   u32 k | u64 blob_len | blob | g (n x 64 B) | g_lagrange (n x 64 B) | u32 num_wit | u32 num_advice | u32 num_pi |
 *   num_wit x { advice (num_advice x n x 32 B) | instance (num_pi x 32 B) }
 * and measures (a) zkgpu_prove_batch_rng over `total` proofs in one call, (b) `threads` callers looping over zkgpu_prove until
 * `total` proofs are done.  Prints one JSON line; the proofs of (b) are written to argv[5] for the Python side to verify.
 * usage: coalesce_bench <input> <threads> <total> <device_mask> <proofs_out> */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "zkgpu.h"

static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
#define CHECK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, zkgpu_last_error()); exit(1); } } while (0)

static uint64_t pk;
static size_t n, num_advice, num_pi, num_wit, proof_len, total;
static uint64_t* wit_advice;   /* num_wit x num_advice x n x 4 */
static uint64_t* wit_inst;     /* num_wit x num_pi x 4 */
static uint8_t* proofs;        /* total x proof_len */
static volatile long next_job = 0;
static int failures = 0;

static void seed_of(size_t job, uint8_t seed[32]) { for (int i = 0; i < 32; ++i) seed[i] = (uint8_t)(job * 131 + i * 7 + 1); }

static void* caller(void* arg) {
    (void)arg;
    /* every caller owns its request buffer, as a request handler does (pageable memory, filled per request) */
    uint64_t* adv = malloc(num_advice * n * 32);
    for (;;) {
        long job = __sync_fetch_and_add(&next_job, 1);
        if ((size_t)job >= total) break;
        size_t w = (size_t)job % num_wit;
        memcpy(adv, wit_advice + w * num_advice * n * 4, num_advice * n * 32);
        uint8_t seed[32]; seed_of((size_t)job, seed);
        int rc = zkgpu_prove(pk, adv, wit_inst + w * num_pi * 4, num_pi, ZKGPU_RNG_CHACHA20_SEED, seed, proofs + (size_t)job * proof_len, proof_len);
        if (rc != 0) { fprintf(stderr, "zkgpu_prove job %ld -> %d: %s\n", job, rc, zkgpu_last_error()); __sync_fetch_and_add(&failures, 1); }
    }
    free(adv);
    return NULL;
}

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s <input> <threads> <total> <device_mask> <proofs_out>\n", argv[0]); return 2; }
    int threads = atoi(argv[2]);
    total = (size_t)atol(argv[3]);
    int mask = atoi(argv[4]);
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("input"); return 2; }
    uint32_t k, nw, na, np_; uint64_t blob_len;
    if (fread(&k, 4, 1, f) != 1 || fread(&blob_len, 8, 1, f) != 1) return 2;
    uint8_t* blob = malloc(blob_len);
    n = (size_t)1 << k;
    uint64_t* g = malloc(n * 64); uint64_t* gl = malloc(n * 64);
    if (fread(blob, 1, blob_len, f) != blob_len || fread(g, 64, n, f) != n || fread(gl, 64, n, f) != n) return 2;
    if (fread(&nw, 4, 1, f) != 1 || fread(&na, 4, 1, f) != 1 || fread(&np_, 4, 1, f) != 1) return 2;
    num_wit = nw; num_advice = na; num_pi = np_;
    wit_advice = malloc(num_wit * num_advice * n * 32); wit_inst = malloc(num_wit * num_pi * 32);
    for (size_t w = 0; w < num_wit; ++w)
        if (fread(wit_advice + w * num_advice * n * 4, 32, num_advice * n, f) != num_advice * n || fread(wit_inst + w * num_pi * 4, 32, num_pi, f) != num_pi) return 2;
    fclose(f);

    CHECK(zkgpu_init(mask));
    uint64_t srs, info[16];
    CHECK(zkgpu_srs_register(g, gl, k, &srs));
    CHECK(zkgpu_pk_create(srs, blob, blob_len, &pk));
    CHECK(zkgpu_pk_info(pk, info));
    proof_len = info[9];
    proofs = calloc(total, proof_len);

    /* (a) the batched call over the same requests, laid out contiguously */
    uint64_t* adv_all = malloc(total * num_advice * n * 32);
    uint64_t* inst_all = malloc(total * num_pi * 32 + 32);
    uint8_t* seeds = malloc(total * 32);
    int32_t* status = malloc(total * sizeof(int32_t));
    for (size_t j = 0; j < total; ++j) {
        size_t w = j % num_wit;
        memcpy(adv_all + j * num_advice * n * 4, wit_advice + w * num_advice * n * 4, num_advice * n * 32);
        memcpy(inst_all + j * num_pi * 4, wit_inst + w * num_pi * 4, num_pi * 32);
        seed_of(j, seeds + 32 * j);
    }
    uint8_t* proofs_batched = calloc(total, proof_len);
    CHECK(zkgpu_prove_batch_rng(pk, adv_all, inst_all, num_pi, total, ZKGPU_RNG_CHACHA20_SEED, seeds, proofs_batched, proof_len, status));   /* warm-up */
    double t0 = now();
    CHECK(zkgpu_prove_batch_rng(pk, adv_all, inst_all, num_pi, total, ZKGPU_RNG_CHACHA20_SEED, seeds, proofs_batched, proof_len, status));
    double t_batched = now() - t0;
    free(adv_all);

    /* (b) concurrent single-proof callers; one untimed round first (dispatcher threads, their workspaces) */
    pthread_t* th = malloc(sizeof(pthread_t) * threads);
    double t_conc = 0;
    for (int round = 0; round < 2; ++round) {
        next_job = 0;
        t0 = now();
        for (int i = 0; i < threads; ++i) pthread_create(&th[i], NULL, caller, NULL);
        for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
        t_conc = now() - t0;
    }
    uint64_t stats[4];
    CHECK(zkgpu_prove_stats(pk, stats));
    int same = memcmp(proofs, proofs_batched, total * proof_len) == 0;
    f = fopen(argv[5], "wb"); fwrite(proofs, proof_len, total, f); fclose(f);
    printf("{\"threads\": %d, \"total\": %zu, \"devices\": %d, \"batched_proofs_per_s\": %.2f, \"concurrent_proofs_per_s\": %.2f, \"ratio\": %.4f, "
           "\"coalesced_requests\": %llu, \"coalesced_batches\": %llu, \"largest_batch\": %llu, \"dispatchers\": %llu, \"failures\": %d, "
           "\"identical_to_batched\": %s}\n",
           threads, total, zkgpu_device_count(), total / t_batched, total / t_conc, t_batched / t_conc, (unsigned long long)stats[0],
           (unsigned long long)stats[1], (unsigned long long)stats[2], (unsigned long long)stats[3], failures, same ? "true" : "false");
    zkgpu_shutdown();
    return failures ? 1 : 0;
}
