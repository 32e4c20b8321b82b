/*
 * ps_oracle.c -- CPU restatement of the parameter-server update semantics that
 * douban/tfmesos selects for its data path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this file's shared object.  Nothing under
 * tfmesos_b200/ links, imports or executes it: the product path is the CUDA
 * library (include/psx.h) and it fails loudly when that library is missing.
 *
 * PARITY UNPINNED upstream: the reference ships no test, golden vector or
 * known-answer value for this path (tox.ini:8 runs flake8 only; the sole
 * expected output in the tree is "42", README.rst:65).  The arithmetic lives in
 * TensorFlow (requirements.txt:10 pins tensorflow-gpu==0.12.0), which is not
 * vendored under /root/reference and cannot be installed here.  What is
 * restated below is TF 0.12's published kernel algebra:
 *
 *   ApplyGradientDescent   var -= grad * lr
 *   ApplyAdam              alpha = lr * sqrt(1 - b2^t) / (1 - b1^t)
 *                          m   += (g - m)   * (1 - b1)
 *                          v   += (g*g - v) * (1 - b2)
 *                          var -= (m * alpha) / (sqrt(v) + eps)
 *   AdamOptimizer._finish  b1^t *= b1 ; b2^t *= b2   (once per minimize())
 *
 * anchored on the reference's own call sites:
 *   examples/mnist/mnist.py:55               GradientDescentOptimizer(0.005)
 *   examples/mnist/mnist_replica.py:147-157  AdamOptimizer(lr) [+ SyncReplicas]
 *   examples/matrix_factorization.py:39-41   GradientDescentOptimizer(0.1)
 *   tfmesos/server.py:52-61                  tf.train.Server  (PS on host CPU)
 *
 * All arithmetic is IEEE-754 binary32, one rounding per operation, never
 * contracted to FMA (build with -ffp-contract=off), which is what Eigen's
 * un-fused CPU expressions produce and what the CUDA kernels reproduce with
 * -fmad=false -prec-div=true -prec-sqrt=true.  Hence the GPU parity bar for
 * these functions is BIT-EXACT, not a tolerance.
 *
 * Update disciplines:
 *   PSX_ORACLE_ASYNC_ORDERED  every worker's push is applied on its own, in
 *                             worker-index order (the serialisable schedule of
 *                             the reference's default async mode)
 *   PSX_ORACLE_SUM            g = ((g0 + g1) + g2) + ...   then one apply
 *   PSX_ORACLE_SYNC_MEAN      g = sum / (float)W           then one apply
 *                             (SyncReplicasOptimizer, mnist_replica.py:148-154)
 */
#define _GNU_SOURCE
#include <math.h>
#include <sched.h>
#include <stdatomic.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

enum { PSX_ORACLE_ASYNC_ORDERED = 0, PSX_ORACLE_SUM = 1, PSX_ORACLE_SYNC_MEAN = 2 };

/* ---- single applies (mnist.py:55 / matrix_factorization.py:39) ------------ */

void psx_oracle_sgd(float *var, const float *g, size_t n, float lr)
{
    for (size_t i = 0; i < n; ++i) {
        float step = g[i] * lr;
        var[i] = var[i] - step;
    }
}

/* alpha exactly as TF's functor spells it: (lr * sqrt(1-b2p)) / (1-b1p) */
float psx_oracle_adam_alpha(float lr, float b1p, float b2p)
{
    float s = sqrtf(1.0f - b2p);
    float num = lr * s;
    return num / (1.0f - b1p);
}

/* one ApplyAdam over n elements with the STORED powers; does not advance them
 * (mnist_replica.py:147) */
void psx_oracle_adam(float *var, float *m, float *v, const float *g, size_t n,
                     float lr, float b1, float b2, float eps, float b1p, float b2p)
{
    const float alpha = psx_oracle_adam_alpha(lr, b1p, b2p);
    const float omb1 = 1.0f - b1;
    const float omb2 = 1.0f - b2;
    for (size_t i = 0; i < n; ++i) {
        float gi = g[i];
        float mi = m[i];
        float vi = v[i];
        float dm = (gi - mi) * omb1;
        mi = mi + dm;
        float g2 = gi * gi;
        float dv = (g2 - vi) * omb2;
        vi = vi + dv;
        float num = mi * alpha;
        float den = sqrtf(vi) + eps;
        var[i] = var[i] - num / den;
        m[i] = mi;
        v[i] = vi;
    }
}

/* ---- one PS round over W gradient slots ----------------------------------- */
/* slots: W contiguous gradients of n floats each (slot w at slots + w*stride).
 * state[0]=b1^t, state[1]=b2^t (Adam only), advanced here exactly as
 * AdamOptimizer._finish does; *step is the global_step counter
 * (mnist.py:46,55; mnist_replica.py:121,156-157). */

static void reduce_slots(float *dst, const float *slots, size_t stride, int W,
                         size_t n, int mean)
{
    for (size_t i = 0; i < n; ++i) {
        float acc = slots[i];
        for (int w = 1; w < W; ++w)
            acc = acc + slots[(size_t)w * stride + i];
        if (mean)
            acc = acc / (float)W;
        dst[i] = acc;
    }
}

int psx_oracle_round_sgd(float *var, const float *slots, size_t stride, int W,
                         size_t n, float lr, int mode, float *scratch,
                         int64_t *step)
{
    if (W < 1)
        return -1;
    if (mode == PSX_ORACLE_ASYNC_ORDERED) {
        for (int w = 0; w < W; ++w)
            psx_oracle_sgd(var, slots + (size_t)w * stride, n, lr);
        *step += W;
        return 0;
    }
    reduce_slots(scratch, slots, stride, W, n, mode == PSX_ORACLE_SYNC_MEAN);
    psx_oracle_sgd(var, scratch, n, lr);
    *step += 1;
    return 0;
}

int psx_oracle_round_adam(float *var, float *m, float *v, const float *slots,
                          size_t stride, int W, size_t n, float lr, float b1,
                          float b2, float eps, int mode, float *state,
                          float *scratch, int64_t *step)
{
    if (W < 1)
        return -1;
    if (mode == PSX_ORACLE_ASYNC_ORDERED) {
        for (int w = 0; w < W; ++w) {
            psx_oracle_adam(var, m, v, slots + (size_t)w * stride, n, lr, b1, b2,
                            eps, state[0], state[1]);
            state[0] = state[0] * b1;
            state[1] = state[1] * b2;
        }
        *step += W;
        return 0;
    }
    reduce_slots(scratch, slots, stride, W, n, mode == PSX_ORACLE_SYNC_MEAN);
    psx_oracle_adam(var, m, v, scratch, n, lr, b1, b2, eps, state[0], state[1]);
    state[0] = state[0] * b1;
    state[1] = state[1] * b2;
    *step += 1;
    return 0;
}

/* ---- index-list (IndexedSlices) round: second restatement of ps_oracle.py rows_round
 * The shard is a [n_rows, d] matrix.  Worker w contributes k[w] rows with STRICTLY
 * ASCENDING row indices idx[w][0..k[w]) and gradients rows[w] (k[w] x d).  A row
 * pushed by several workers gets ((g_w + g_w') + ...) in worker order; mean != 0
 * divides by W; SGD / Adam (stored powers, not advanced here) are applied ONCE to
 * every touched row; untouched rows keep var / m / v.  (the NMF row
 * blocks of examples/matrix_factorization.py:21-28,43-49.)  Returns -1 on a
 * non-ascending list or an index outside the matrix. */
static long rows_find(const int64_t *idx, long n, int64_t key)
{
    long lo = 0, hi = n;
    while (lo < hi) {
        long mid = (lo + hi) / 2;
        if (idx[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return (lo < n && idx[lo] == key) ? lo : -1;
}

int psx_oracle_rows_round(float *var, float *m, float *v, size_t n_rows, size_t d, int W,
                          const int64_t *const *idx, const float *const *rows, const size_t *k,
                          int opt_adam, int mean, float lr, float b1, float b2, float eps,
                          float b1p, float b2p, float *scratch /* d floats */)
{
    for (int w = 0; w < W; ++w)
        for (size_t j = 0; j < k[w]; ++j) {
            if (idx[w][j] < 0 || (size_t)idx[w][j] >= n_rows) return -1;
            if (j > 0 && idx[w][j] <= idx[w][j - 1]) return -1;
        }
    for (int w = 0; w < W; ++w) {
        for (size_t j = 0; j < k[w]; ++j) {
            const int64_t r = idx[w][j];
            int first = 1;                       /* is w the lowest worker holding row r? */
            for (int w2 = 0; w2 < w && first; ++w2)
                if (rows_find(idx[w2], (long)k[w2], r) >= 0) first = 0;
            if (!first) continue;
            for (size_t e = 0; e < d; ++e) scratch[e] = rows[w][j * d + e];
            for (int w2 = w + 1; w2 < W; ++w2) {
                long p = rows_find(idx[w2], (long)k[w2], r);
                if (p < 0) continue;
                for (size_t e = 0; e < d; ++e) scratch[e] = scratch[e] + rows[w2][(size_t)p * d + e];
            }
            if (mean)
                for (size_t e = 0; e < d; ++e) scratch[e] = scratch[e] / (float)W;
            if (opt_adam)
                psx_oracle_adam(var + (size_t)r * d, m + (size_t)r * d, v + (size_t)r * d, scratch, d,
                                lr, b1, b2, eps, b1p, b2p);
            else
                psx_oracle_sgd(var + (size_t)r * d, scratch, d, lr);
        }
    }
    return 0;
}

/* ---- bf16 wire format (BASELINE config #4: grads pushed / params pulled bf16)
 * round-to-nearest-even float -> bf16, NaN kept quiet; matches
 * __float2bfloat16_rn on the device. */
uint16_t psx_oracle_f32_to_bf16(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u)
        return (uint16_t)((u >> 16) | 0x0040u);
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
}

float psx_oracle_bf16_to_f32(uint16_t h)
{
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

void psx_oracle_cast_f32_bf16(uint16_t *dst, const float *src, size_t n)
{
    for (size_t i = 0; i < n; ++i)
        dst[i] = psx_oracle_f32_to_bf16(src[i]);
}

void psx_oracle_cast_bf16_f32(float *dst, const uint16_t *src, size_t n)
{
    for (size_t i = 0; i < n; ++i)
        dst[i] = psx_oracle_bf16_to_f32(src[i]);
}

/* ---- multi-threaded CPU-PS round: the timed CPU baseline ------------------
 * The reference's CPU path per worker-step: PULL = the PS
 * copies every variable out to the worker (one RecvTensor per variable),
 * PUSH = the worker's gradient is copied into the PS, APPLY on PS host cores
 * (Eigen thread pool).  This is the best case for that path: both transfers
 * are plain memcpy (no protobuf, no TCP) and the apply is spread over every
 * host thread, element range by element range; per element the arithmetic is
 * the scalar code above, so results equal the single-thread functions bit for
 * bit.  worker_grad[w] / worker_param[w] are the workers' private buffers. */

static void range_of(size_t n, int part, int parts, size_t *lo, size_t *hi)
{
    size_t chunk = (n + (size_t)parts - 1) / (size_t)parts;
    chunk = (chunk + 15) & ~(size_t)15;
    *lo = (size_t)part * chunk;
    *hi = *lo + chunk;
    if (*lo > n) *lo = n;
    if (*hi > n) *hi = n;
}

int psx_oracle_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    if (n < 1) n = 1;
    if (n > 256) n = 256;
    return (int)n;
}

typedef struct {
    float *var, *m, *v, *slots, *scratch;
    size_t stride, n;
    float *const *worker_grad;
    float *const *worker_param;
    int W, opt_adam, mode, part, parts;
    float lr, b1, b2, eps;
    const float *b1p, *b2p;
    int init;                 /* 1: first-touch + fill instead of a round */
} cpu_ps_job;

/* deterministic fill values (exact integer hash -> (-0.01, 0.01)) */
static float synth(size_t i, unsigned seed)
{
    uint64_t h = ((uint64_t)i * 2654435761ull + (uint64_t)seed * 40503ull + 12345ull) & 0xFFFFFFull;
    return ((float)h / 16777216.0f - 0.5f) * 0.02f;
}

static void cpu_ps_range(cpu_ps_job *j, size_t lo, size_t hi)
{
    size_t cnt = hi - lo;
    if (cnt == 0)
        return;
    if (j->init) {
        /* FIRST TOUCH by the thread that will own this range in every round: the
         * pages land on that thread's NUMA node (the arrays arrive untouched from
         * the allocator), so placement -- and with it the measured number -- does
         * not depend on which socket the launching thread happened to run on. */
        for (size_t i = lo; i < hi; ++i) {
            j->var[i] = synth(i, 7u) * 100.0f;
            j->m[i] = 0.0f;
            j->v[i] = 0.0f;
            j->scratch[i] = 0.0f;
        }
        for (int w = 0; w < j->W; ++w) {
            float *g = j->worker_grad[w], *p = j->worker_param[w];
            float *s = j->slots + (size_t)w * j->stride;
            for (size_t i = lo; i < hi; ++i) {
                g[i] = synth(i, 100u + (unsigned)w);
                p[i] = 0.0f;
                s[i] = 0.0f;
            }
        }
        return;
    }
    /* PUSH: worker -> PS receive buffers */
    for (int w = 0; w < j->W; ++w)
        memcpy(j->slots + (size_t)w * j->stride + lo, j->worker_grad[w] + lo, cnt * 4);
    /* APPLY on the PS */
    if (j->mode == PSX_ORACLE_ASYNC_ORDERED) {
        for (int w = 0; w < j->W; ++w) {
            const float *g = j->slots + (size_t)w * j->stride + lo;
            if (j->opt_adam)
                psx_oracle_adam(j->var + lo, j->m + lo, j->v + lo, g, cnt, j->lr,
                                j->b1, j->b2, j->eps, j->b1p[w], j->b2p[w]);
            else
                psx_oracle_sgd(j->var + lo, g, cnt, j->lr);
        }
    } else {
        reduce_slots(j->scratch + lo, j->slots + lo, j->stride, j->W, cnt,
                     j->mode == PSX_ORACLE_SYNC_MEAN);
        if (j->opt_adam)
            psx_oracle_adam(j->var + lo, j->m + lo, j->v + lo, j->scratch + lo, cnt,
                            j->lr, j->b1, j->b2, j->eps, j->b1p[0], j->b2p[0]);
        else
            psx_oracle_sgd(j->var + lo, j->scratch + lo, cnt, j->lr);
    }
    /* PULL: PS -> every worker */
    for (int w = 0; w < j->W; ++w)
        memcpy(j->worker_param[w] + lo, j->var + lo, cnt * 4);
}

/* ---- persistent thread pool ------------------------------------------------
 * Created once (psx_oracle_pool_start); every round is two barrier crossings.
 * Thread p is pinned to the p-th CPU of the affinity mask the process had when
 * the pool started, so a range is always touched from the same core (and, with
 * the first-touch initialisation above, from the socket its pages live on).
 * (TF's PS runs its Eigen thread pool the same way: long-lived workers, one per
 * core -- tfmesos/server.py:52-61 sizes it from the task's `cpus`.) */
#define PSX_ORACLE_MAX_THREADS 512
/* Work distribution: thread p owns range p (the one it first-touched) and walks it
 * in 64 Ki-element chunks claimed from the range's atomic cursor; a thread that
 * has finished its own range STEALS chunks from the others' cursors.  On a quiet
 * box nothing is stolen and every byte stays NUMA-local; when a core is busy with
 * somebody else's work (the GPU boxes' host cores are shared) its range is
 * finished by the idle threads instead of stalling the whole round. */
#define PSX_ORACLE_CHUNK ((size_t)65536)
typedef struct {
    _Atomic size_t next;
    size_t hi;
    char pad[48];
} range_cursor;
static range_cursor g_cursor[PSX_ORACLE_MAX_THREADS];

static void cpu_ps_work(cpu_ps_job *j, int p)
{
    for (int k = 0; k < j->parts; ++k) {
        range_cursor *c = &g_cursor[(p + k) % j->parts];   /* own range first, then steal */
        for (;;) {
            size_t lo = atomic_fetch_add_explicit(&c->next, PSX_ORACLE_CHUNK, memory_order_relaxed);
            if (lo >= c->hi)
                break;
            size_t hi = lo + PSX_ORACLE_CHUNK;
            cpu_ps_range(j, lo, hi < c->hi ? hi : c->hi);
        }
    }
}

static struct {
    int n;
    pthread_t tid[PSX_ORACLE_MAX_THREADS];
    int cpu[PSX_ORACLE_MAX_THREADS];
    pthread_barrier_t start, done;
    cpu_ps_job job;             /* template for the current round */
    volatile int quit;
} g_pool;

static void *pool_main(void *arg)
{
    int p = (int)(intptr_t)arg;
    if (g_pool.cpu[p] >= 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        CPU_SET(g_pool.cpu[p], &set);
        pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
    }
    for (;;) {
        pthread_barrier_wait(&g_pool.start);
        if (g_pool.quit)
            return NULL;
        cpu_ps_job j = g_pool.job;
        j.part = p;
        if (p < j.parts) {
            if (j.init) {              /* first touch: strictly the owner, no stealing */
                size_t lo, hi;
                range_of(j.n, p, j.parts, &lo, &hi);
                cpu_ps_range(&j, lo, hi);
            } else {
                cpu_ps_work(&j, p);
            }
        }
        pthread_barrier_wait(&g_pool.done);
    }
}

int psx_oracle_pool_threads(void) { return g_pool.n; }

/* threads <= 0: one per CPU this process may run on.  Returns the pool size. */
int psx_oracle_pool_start(int threads)
{
    if (g_pool.n > 0)
        return g_pool.n;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    int ncpu = 0, cpus[PSX_ORACLE_MAX_THREADS];
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0) {
        for (int c = 0; c < CPU_SETSIZE && ncpu < PSX_ORACLE_MAX_THREADS; ++c)
            if (CPU_ISSET(c, &allowed))
                cpus[ncpu++] = c;
    }
    if (ncpu == 0) {
        ncpu = psx_oracle_threads();
        for (int c = 0; c < ncpu; ++c) cpus[c] = -1;
    }
    int n = threads > 0 ? threads : ncpu;
    if (n > PSX_ORACLE_MAX_THREADS) n = PSX_ORACLE_MAX_THREADS;
    g_pool.quit = 0;
    pthread_barrier_init(&g_pool.start, NULL, (unsigned)n + 1);
    pthread_barrier_init(&g_pool.done, NULL, (unsigned)n + 1);
    for (int p = 0; p < n; ++p) {
        g_pool.cpu[p] = cpus[p % ncpu];
        if (pthread_create(&g_pool.tid[p], NULL, pool_main, (void *)(intptr_t)p) != 0) {
            g_pool.n = p;       /* the barriers expect n+1: unusable, report failure */
            return -3;
        }
    }
    g_pool.n = n;
    return n;
}

void psx_oracle_pool_stop(void)
{
    if (g_pool.n <= 0)
        return;
    g_pool.quit = 1;
    pthread_barrier_wait(&g_pool.start);
    for (int p = 0; p < g_pool.n; ++p)
        pthread_join(g_pool.tid[p], NULL);
    pthread_barrier_destroy(&g_pool.start);
    pthread_barrier_destroy(&g_pool.done);
    g_pool.n = 0;
}

static int pool_run(const cpu_ps_job *job)
{
    if (g_pool.n <= 0 && psx_oracle_pool_start(0) <= 0)
        return -3;
    g_pool.job = *job;
    size_t cap = job->n / 65536 + 1;           /* never less than 64 Ki elements each */
    g_pool.job.parts = (size_t)g_pool.n > cap ? (int)cap : g_pool.n;
    for (int p = 0; p < g_pool.job.parts; ++p) {
        size_t lo, hi;
        range_of(job->n, p, g_pool.job.parts, &lo, &hi);
        atomic_store_explicit(&g_cursor[p].next, lo, memory_order_relaxed);
        g_cursor[p].hi = hi;
    }
    pthread_barrier_wait(&g_pool.start);
    pthread_barrier_wait(&g_pool.done);
    return g_pool.job.parts;
}

/* first-touch + deterministic fill of every array of a CPU-PS instance, each
 * range by its owning pool thread */
int psx_oracle_cpu_ps_init(float *var, float *m, float *v, float *slots, size_t stride,
                           float *const *worker_grad, float *const *worker_param, int W,
                           size_t n, float *scratch)
{
    if (W < 1 || W > 64)
        return -1;
    cpu_ps_job j;
    memset(&j, 0, sizeof(j));
    j.var = var; j.m = m; j.v = v; j.slots = slots; j.scratch = scratch;
    j.stride = stride; j.n = n; j.worker_grad = worker_grad; j.worker_param = worker_param;
    j.W = W; j.init = 1;
    return pool_run(&j);
}

/* One CPU-PS round on the persistent pool.  Returns the number of threads that
 * took part (> 0) or a negative error. */
int psx_oracle_cpu_ps_round(float *var, float *m, float *v, float *slots,
                            size_t stride, float *const *worker_grad,
                            float *const *worker_param, int W, size_t n,
                            int opt_adam, float lr, float b1, float b2, float eps,
                            int mode, float *state, float *scratch, int64_t *step,
                            int threads)
{
    if (W < 1)
        return -1;
    if (W > 64)
        return -2;
    if (g_pool.n <= 0 && psx_oracle_pool_start(threads) <= 0)
        return -3;
    float b1p[64], b2p[64];
    b1p[0] = opt_adam ? state[0] : 0.0f;
    b2p[0] = opt_adam ? state[1] : 0.0f;
    for (int w = 1; w < W; ++w) {
        b1p[w] = b1p[w - 1] * b1;
        b2p[w] = b2p[w - 1] * b2;
    }
    cpu_ps_job j = { var, m, v, slots, scratch, stride, n, worker_grad,
                     worker_param, W, opt_adam, mode, 0, 0,
                     lr, b1, b2, eps, b1p, b2p, 0 };
    int parts = pool_run(&j);
    if (parts <= 0)
        return parts;
    int applies = (mode == PSX_ORACLE_ASYNC_ORDERED) ? W : 1;
    if (opt_adam) {
        for (int k = 0; k < applies; ++k) {
            state[0] = state[0] * b1;
            state[1] = state[1] * b2;
        }
    }
    *step += applies;
    return parts;
}
